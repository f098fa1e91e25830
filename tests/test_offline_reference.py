"""The offline barrel-roll path on the host: config_barrel_roll mirrors the reference's values and
mpc_utils.reference_barell_roll (numpy, product) equals the oracle's statement-by-statement restatement of reference
mpc_utils.py:316-398, plus first-principles checks of the planned trajectory."""
import numpy as np
import pytest

import mpx_b200.config.config_barrel_roll as cfg
from mpx_b200.utils import mpc_utils as product
from oracle import mpc_utils as oracle_utils


@pytest.mark.parametrize("N", [101, 90, 140])
def test_reference_barell_roll_equals_oracle(N):
    a = product.reference_barell_roll(N, cfg.dt, cfg.n_joints, cfg.n_contact, cfg.p_legs0, cfg.q0)
    b = oracle_utils.reference_barell_roll(N, cfg.dt, cfg.n_joints, cfg.n_contact, cfg.p_legs0, cfg.q0)
    assert a[0].shape == (N, 13 + cfg.n_joints + 7 * cfg.n_contact) and a[1].shape == (N, 4 * cfg.n_contact)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-12


def test_reference_barell_roll_known_answers():
    N, nj, nc = cfg.N + 1, cfg.n_joints, cfg.n_contact
    ref, par = product.reference_barell_roll(N, cfg.dt, nj, nc, cfg.p_legs0, cfg.q0)
    contact = ref[:, 13 + nj + 3 * nc:13 + nj + 4 * nc]
    # phases of 0.2 / 0.2 / 0.3 / 0.1 s at dt = 0.01, then rest (mpc_utils.py:317-320, 335-386)
    assert (contact[:20] == 1).all() and (contact[20:40] == [0, 1, 0, 1]).all() and (contact[40:70] == 0).all()
    assert (contact[70:] == 1).all() and (par[:, :nc] == contact).all()
    quat = ref[:, 3:7]
    assert np.abs(np.linalg.norm(quat, axis=1) - 1).max() < 1e-12
    # one full turn (with the reference's pi = 3.14) over the 60 knots of phases 2-4: knot 79 has turned 59/60 of it
    ang = 2 * np.arctan2(quat[79, 1], quat[79, 0])
    assert abs(ang - 2 * 3.14 * 59 / 60) < 1e-9 and (quat[80:] == [1, 0, 0, 0]).all()
    # flight: ballistic arc from z = 0.4 landing at 0.28 after 0.3 s, vertical speed continuous with the push-off
    z, vz = ref[:, 2], ref[:, 7 + nj + 2]
    v0 = (0.28 - 0.4 + 0.5 * 9.81 * 0.09) / 0.3
    assert abs(z[40] - 0.4) < 1e-12 and abs(vz[40] - v0) < 1e-12 and abs(vz[39] - v0 * 19 / 20) < 1e-12
    assert abs((0.4 + v0 * 0.3 - 0.5 * 9.81 * 0.09) - 0.28) < 1e-12 and (z[70:] == 0.28).all()
    # feet: nominal footholds under the base, on the ground
    foot = ref[:, 13 + nj:13 + nj + 3 * nc]
    assert (foot[:, 2::3] == 0).all() and np.abs(foot[:, 0::3] - np.asarray(cfg.p_legs0)[0::3]).max() < 1e-12


def test_config_barrel_roll_values():
    """reference mpx/config/config_barrel_roll.py, the values the solve reads."""
    assert (cfg.dt, cfg.N, cfg.mpc_frequency, cfg.n, cfg.m, cfg.n_joints, cfg.n_contact) == (0.01, 100, 50, 61, 12, 12, 4)
    assert cfg.grf_as_state and int(1 / (cfg.dt * cfg.mpc_frequency)) == 2
    d = np.diag(cfg.W)
    assert cfg.W.shape == (72, 72) and np.count_nonzero(cfg.W - np.diag(d)) == 0
    assert d[:6].tolist() == [0, 0, 5e4, 100, 100, 100] and (d[6:18] == 1e2).all() and (d[18:21] == 100).all()
    assert (d[21:24] == 1e2).all() and (d[24:36] == 1e-1).all() and d[36:39].tolist() == [1e3, 1e3, 5e4]
    assert (d[48:60] == 1e-1).all() and (d[60:] == 0).all()
    assert cfg.cost.args == (False,) and cfg.hessian_approx.args == (False,)  # no swing tracking: contact-masked feet
    assert cfg.reference is product.reference_barell_roll
