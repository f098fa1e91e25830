"""GPU parity of mpxb_whole_body_interface (mpc_utils.py:275-313 behind
mpc_wrapper_srbd.BatchedMPCControllerWrapper.whole_body_run) against the float64 oracle restatement."""
import numpy as np
import pytest
import torch

import mpx_b200.config.config_aliengo as cfg_wb
import mpx_b200.config.config_srbd as cfg
from oracle import mpc_utils as oracle_utils
from oracle.problems import Problem

pytestmark = pytest.mark.gpu


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(1.0, np.abs(b).max()))


def test_whole_body_interface_matches_oracle():
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper
    E, nj, nc = 4, cfg.n_joints, cfg.n_contact
    rng = np.random.default_rng(21)
    w = BatchedMPCControllerWrapper(cfg, E)
    pr = Problem(cfg_wb)  # same robot (aliengo) as a whole-body tree
    body_id = pr.body_id
    qpos = np.zeros((E, 7 + nj)); qvel = rng.uniform(-0.5, 0.5, (E, 6 + nj))
    quat = rng.normal(0, 1, (E, 4)) * 0.05 + np.array([1.0, 0, 0, 0])
    qpos[:, :3] = np.asarray(cfg.p0) + rng.uniform(-0.05, 0.05, (E, 3))
    qpos[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    qpos[:, 7:] = np.asarray(cfg.q0) + rng.uniform(-0.2, 0.2, (E, nj))
    grf = rng.uniform(-10, 60, (E, 3 * nc))
    contact = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [0, 0, 0, 0]], dtype=np.float64)
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).cuda()
    feet = np.stack([pr.foot_fk(qpos[e]) for e in range(E)])
    foot_ref = feet + rng.uniform(-0.03, 0.03, feet.shape)
    foot_ref_dot = rng.uniform(-0.3, 0.3, feet.shape)
    w.grf, w.foot_ref, w.foot_ref_dot, w.contact = T(grf), T(foot_ref), T(foot_ref_dot), T(contact)
    tau, J = w.whole_body_run(T(qpos), T(qvel), return_jacobian=True)
    torch.cuda.synchronize()
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64)
    for e in range(E):
        tau_o, J_o = oracle_utils.whole_body_interface(pr.m, pr.contact_id, body_id, t64(cfg.Kp), t64(cfg.Kd), t64(qpos[e]),
                                                       t64(qvel[e]), t64(grf[e]), t64(foot_ref[e]), t64(foot_ref_dot[e]),
                                                       t64(contact[e]))
        assert _rel(J[e].cpu().numpy(), J_o.numpy()) < 1e-5
        assert _rel(tau[e].cpu().numpy(), tau_o.numpy()) < 1e-4, (e, tau[e].cpu().numpy(), tau_o.numpy())


def test_whole_body_interface_argument_errors():
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper
    w = BatchedMPCControllerWrapper(cfg, 2)
    with pytest.raises(ValueError):
        w.whole_body_run(torch.zeros(2, 5).cuda(), torch.zeros(2, 18).cuda())


def test_whole_body_interface_at_a_leg_singularity():
    """Straight knee: the 3x3 leg Jacobian loses rank, but J [nv, 3 nc] does not -- its six floating-base rows carry
    every foot force -- so pinv(J') (mpc_utils.py:304) is the plain left inverse the kernel's Cholesky computes.
    Checked here: the leg block IS singular, J'J stays well conditioned, and the torques match the oracle's pinv."""
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper
    E, nj, nc = 2, cfg.n_joints, cfg.n_contact
    rng = np.random.default_rng(4)
    w = BatchedMPCControllerWrapper(cfg, E)
    pr = Problem(cfg_wb)
    qpos = np.zeros((E, 7 + nj)); qvel = rng.uniform(-0.3, 0.3, (E, 6 + nj))
    qpos[:, :3] = np.asarray(cfg.p0); qpos[:, 3] = 1.0
    qpos[:, 7:] = np.asarray(cfg.q0)
    qpos[0, 7 + 5] = 0.0        # FR calf joint: knee fully extended
    qpos[1, 7 + 2::3] = 0.0     # every knee extended
    grf = rng.uniform(-10, 60, (E, 3 * nc))
    contact = np.zeros((E, nc))  # all legs in swing: tau = the feedback branch through pinv(J')
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).cuda()
    feet = np.stack([pr.foot_fk(qpos[e]) for e in range(E)])
    foot_ref = feet + rng.uniform(-0.03, 0.03, feet.shape)
    foot_ref_dot = rng.uniform(-0.3, 0.3, feet.shape)
    w.grf, w.foot_ref, w.foot_ref_dot, w.contact = T(grf), T(foot_ref), T(foot_ref_dot), T(contact)
    tau, J = w.whole_body_run(T(qpos), T(qvel), return_jacobian=True)
    torch.cuda.synchronize()
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64)
    for e in range(E):
        Je = J[e].double().cpu().numpy()
        leg = Je[6 + 3:6 + 6, 3:6] if e == 0 else Je[6:9, 0:3]   # joint rows x foot columns of one straight leg
        assert np.linalg.svd(leg, compute_uv=False)[-1] < 1e-5           # the leg itself IS singular
        assert np.linalg.eigvalsh(Je.T @ Je)[0] > 1e-3                   # J'J is not (0.04 / 0.004 here)
        tau_o, J_o = oracle_utils.whole_body_interface(pr.m, pr.contact_id, pr.body_id, t64(cfg.Kp), t64(cfg.Kd), t64(qpos[e]),
                                                       t64(qvel[e]), t64(grf[e]), t64(foot_ref[e]), t64(foot_ref_dot[e]),
                                                       t64(contact[e]))
        assert _rel(tau[e].cpu().numpy(), tau_o.numpy()) < 1e-4, (e, tau[e].cpu().numpy(), tau_o.numpy())
