"""GPU parity of the whole-body solve (configs 2-5 of BASELINE.json) through the C ABI against the committed
golden fixtures (tests/golden/*.npz, produced from the float64 oracle by tests/golden/make_golden.py).

Tolerance: north_star's 1e-4 relative in FP32, widened only where the float32 oracle's OWN error against float64
(stored in the fixture as budget_<name>) says FP32 cannot do better: tol = max(1e-4, 4 x budget)."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("go2", "warm"), ("go2", "second"), ("go2", "cold"), ("aliengo", "warm"), ("aliengo", "second"),
         ("h1", "warm"), ("h1", "second"), ("talos", "warm"), ("talos", "second")]


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(1.0, np.abs(b).max()))


def _load(robot, case):
    path = os.path.join(GOLDEN, f"{robot}_{case}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing (run tests/golden/make_golden.py)")
    return np.load(path)


def _solve(cfg, g, limited_memory, eager=True):
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = g["x0"].shape[0]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    solver = MpcSolver(cfg, E, limited_memory=limited_memory, eager_line_search=eager)
    X, U, V = solver.solve(T(g["reference"]), T(g["parameter"]), T(cfg.W), T(g["x0"]), T(g["X0"]), T(g["U0"]), T(g["V0"]))
    torch.cuda.synchronize()
    return solver, X, U, V


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("robot,case", CASES)
def test_wb_step_matches_golden(robot, case, eager):
    """Linearisation (A,B,Q,R,M,q,r,c,g), sequential-LQ gains / value function / search direction, line-search
    merit data, accepted step and the new iterate of one optimizers.mpc call.  eager: all 11 trial points are
    evaluated like the reference's vmap; lazy (the default): alpha = 1 first, the rest only where it is rejected --
    the accepted step and the new iterate must not change."""
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    g = _load(robot, case)
    solver, X, U, V = _solve(cfg, g, True, eager)
    E = g["x0"].shape[0]
    for e in range(E):
        tol = lambda nm: max(REL_TOL, 4 * float(g["budget_" + nm][e]))
        for nm in ("A", "B", "Q", "R", "M", "q", "r", "c", "K", "k", "P", "p", "dX", "dU", "dV"):
            err = _rel(solver.view(nm)[e].cpu().numpy(), g[nm][e])
            assert err < tol(nm), (nm, err, tol(nm))
        assert abs(float(solver.view("g")[e]) - float(g["g"][e])) / max(1.0, abs(float(g["g"][e]))) < REL_TOL
        npts = 11 if (eager or int(g["alpha_index"][e]) > 0) else 1  # lazy: points 1..10 only if alpha = 1 is rejected
        assert _rel(solver.view("ls_phi")[e, :npts].cpu().numpy(), g["phi"][e][:npts]) < tol("phi")
        assert _rel(solver.view("ls_theta")[e, :npts].cpu().numpy(), g["theta"][e][:npts]) < max(1e-3, tol("theta"))
        assert int(solver.alpha_index[e]) == int(g["alpha_index"][e])
        for nm, got in (("X", X), ("U", U), ("V", V)):
            err = _rel(got[e].cpu().numpy(), g[nm][e])
            assert err < tol(nm), (nm, err, tol(nm))
        assert int(solver.status[e]) == 0 and int(solver.iters[e]) == 1


def test_wb_foot_positions_match_golden():
    """mpxb_wb_foot_positions (device FK) reproduces the foot block of x0, which the oracle's FK produced."""
    from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper
    for robot in ("go2", "h1", "talos"):
        cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
        g = _load(robot, "warm")
        E, nj, nc = g["x0"].shape[0], cfg.n_joints, cfg.n_contact
        w = BatchedMPCControllerWrapper(cfg, E, limited_memory=True)
        feet = w.foot_positions(torch.as_tensor(g["x0"][:, :7 + nj], dtype=torch.float32).cuda())
        assert _rel(feet.cpu().numpy(), g["x0"][:, 13 + 2 * nj:13 + 2 * nj + 3 * nc]) < 1e-5


def test_wb_batched_wrapper_closed_loop():
    """BatchedMPCControllerWrapper: 8 closed-loop ticks on the model stay finite, carry the robot and track height."""
    import mpx_b200.config.config_go2 as cfg
    from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper
    g = _load("go2", "warm")
    E, nj, nc = 16, cfg.n_joints, cfg.n_contact
    mpc = BatchedMPCControllerWrapper(cfg, E, limited_memory=True)
    x0 = torch.as_tensor(np.tile(g["x0"][:1], (E, 1)), dtype=torch.float32).cuda()
    data = mpc.reset(mpc.make_data(), x0[:, :7 + nj], x0[:, 7 + nj:13 + 2 * nj], x0[:, 13 + 2 * nj:13 + 2 * nj + 3 * nc])
    cmd = torch.tensor([[0.3, 0, 0, 0, 0, 0, cfg.robot_height]] * E, dtype=torch.float32).cuda()
    for _ in range(8):
        data, tau = mpc.run(data, x0, cmd)
        x0 = mpc.X[:, 1, :].clone()
    torch.cuda.synchronize()
    assert tau.shape == (E, nj)
    assert torch.isfinite(mpc.X).all() and torch.isfinite(mpc.U).all() and torch.isfinite(tau).all()
    assert (mpc.X[:, 1, 2] - cfg.robot_height).abs().max() < 0.1
    assert (data.grf[:, 2::3].sum(1) > 0.5 * 15.2 * 9.81).all()   # vertical GRFs carry most of the 15.2 kg robot
    assert torch.equal(tau[0], tau[-1])                            # identical environments give identical results


def test_per_env_W_equals_shared():
    import mpx_b200.config.config_go2 as cfg
    g = _load("go2", "warm")
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = g["x0"].shape[0]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    s = MpcSolver(cfg, E, limited_memory=True)
    args = (T(g["x0"]), T(g["X0"]), T(g["U0"]), T(g["V0"]))
    a = [t.clone() for t in s.solve(T(g["reference"]), T(g["parameter"]), T(cfg.W), *args)]
    b = s.solve(T(g["reference"]), T(g["parameter"]), T(np.tile(cfg.W, (E, 1, 1))), *args)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
