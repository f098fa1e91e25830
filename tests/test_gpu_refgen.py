"""GPU parity of mpxb_reference_generator (gait timer + swing/foothold reference generator kernel) against the float64
oracle restatement of reference mpx/utils/mpc_utils.py:7-270, for every robot configuration, several gait phases
(so lift-off / touch-down events fall inside the horizon), measured-contact patterns and both timer conventions."""
import importlib

import numpy as np
import pytest
import torch

from oracle import mpc_utils as omu

pytestmark = pytest.mark.gpu

TOL = 2e-5  # float32 kernel vs float64 oracle on O(1) quantities (positions, velocities, forces / mass*g)


def _case(robot, E, seed):
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    rng = np.random.default_rng(seed)
    srbd = robot == "srbd"
    nj, nc = (0 if srbd else cfg.n_joints), cfg.n_contact
    n = cfg.n
    x = np.zeros((E, n))
    x[:, :3] = np.asarray(cfg.p0) + rng.uniform(-0.05, 0.05, (E, 3))
    q = rng.normal(size=(E, 4)) * 0.1 + np.array([1.0, 0, 0, 0])
    x[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    x[:, 7:] = rng.uniform(-0.3, 0.3, (E, n - 7))
    foot = np.tile(np.asarray(cfg.p_legs0, dtype=np.float64), (E, 1)) + rng.uniform(-0.03, 0.03, (E, 3 * nc))
    foot[:, 0::3] += x[:, 0:1]; foot[:, 1::3] += x[:, 1:2]
    if not srbd:
        x[:, 13 + 2 * nj:13 + 2 * nj + 3 * nc] = foot
    cmd = np.stack([rng.uniform(-0.2, 0.5, E), rng.uniform(-0.1, 0.1, E), np.zeros(E), np.zeros(E), np.zeros(E),
                    rng.uniform(-0.3, 0.3, E), np.full(E, cfg.robot_height)], 1)
    cmd[0, :3] = 0.0  # zero commanded velocity: the 1e-6 guard of initial_speed
    timer = np.mod(np.asarray(cfg.timer_t)[None] + rng.uniform(0, 1, (E, 1)), 1.0)
    liftoff = foot + rng.uniform(-0.02, 0.02, foot.shape)
    contact = (rng.uniform(size=(E, nc)) < 0.7).astype(np.float64)
    return cfg, x, foot, cmd, timer, liftoff, contact


@pytest.mark.parametrize("robot", ["srbd", "go2", "aliengo", "h1", "talos"])
@pytest.mark.parametrize("advance_first", [False, True])
def test_reference_generator_matches_oracle(robot, advance_first):
    from mpx_b200 import _lib
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = 37
    cfg, x, foot, cmd, timer, liftoff, contact = _case(robot, E, seed=5)
    srbd = robot == "srbd"
    nc = cfg.n_contact
    terrain = cfg.use_terrain_estimator if srbd else cfg.use_terrain_estimation
    mass, clr = (cfg.mass, cfg.clearence_speed) if srbd else (23.7, 0.33)
    solver = MpcSolver(cfg, E, limited_memory=True)
    gait = _lib.make_gait(cfg, mass, cfg.duty_factor, cfg.step_freq, cfg.step_height, clr, terrain, advance_first)
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous().cuda()
    xd, footd, timerd, liftd = T(x), T(foot), T(timer), T(liftoff)
    timer_out = torch.empty_like(timerd)
    if srbd:
        fview, stride = footd, 3 * nc
    else:
        off = 13 + 2 * cfg.n_joints
        fview, stride = xd[:, off:], cfg.n
    out = solver.reference_generator(gait, xd, fview, stride, T(cmd), T(contact), timerd, timer_out, liftd,
                                     want_foot_ref_dot=True)
    torch.cuda.synchronize()
    ref_d, par_d, frd_d = (t.cpu().numpy() for t in out)
    mpc_dt = 1.0 / cfg.mpc_frequency
    for e in range(E):
        _, t_adv = omu.timer_run(cfg.duty_factor, cfg.step_freq, timer[e], mpc_dt)
        t_gen = t_adv if advance_first else timer[e]
        kw = dict(t_timer=t_gen, x=x[e], foot=foot[e], input=cmd[e], duty_factor=cfg.duty_factor, step_freq=cfg.step_freq,
                  step_height=cfg.step_height, liftoff=liftoff[e], contact=contact[e], clearence_speed=clr)
        if srbd:
            ref_o, par_o, lift_o, frd_o = omu.reference_generator_srbd(terrain, cfg.N, cfg.dt, nc, mass, cfg.p_legs0, **kw)
            assert np.abs(frd_d[e] - frd_o).max() < 1e-4
        else:
            ref_o, par_o, lift_o = omu.reference_generator(terrain, cfg.N, cfg.dt, cfg.n_joints, nc, mass, cfg.p_legs0,
                                                           cfg.q0, **kw)
        scale = np.maximum(1.0, np.abs(ref_o))
        assert (np.abs(ref_d[e] - ref_o) / scale).max() < TOL, (robot, e)
        assert np.abs(par_d[e] - par_o).max() < TOL
        assert np.abs(liftd[e].cpu().numpy() - lift_o).max() < TOL
        assert np.abs(timer_out[e].cpu().numpy() - t_adv).max() < 1e-6
