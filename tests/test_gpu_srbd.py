"""GPU parity of the SRBD solve (config 1 of BASELINE.json) through the C ABI against the float64 oracle."""
import numpy as np
import pytest
import torch

import mpx_b200.config.config_srbd as cfg
from oracle.problems import Problem, synthetic_inputs

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star: trajectories, gains and multipliers within 1e-4 relative in FP32


def _rel(a, b):
    b = np.asarray(b)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(1.0, np.abs(b).max()))


@pytest.fixture(scope="module")
def problem():
    return Problem(cfg)


def _run(solver, inp, W=None):
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous().cuda()
    out = solver.solve(T(inp["reference"]), T(inp["parameter"]), T(cfg.W if W is None else W), T(inp["x0"]),
                       T(inp["X0"]), T(inp["U0"]), T(inp["V0"]))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("lq", ["sequential", "scan_parallel", "scan_serial"])
@pytest.mark.parametrize("case", ["warm", "second", "cold", "far"])
def test_srbd_step_matches_golden(case, lq, eager):
    """One optimizers.mpc call of the SRBD model (config 1) against tests/golden/srbd_<case>.npz: every intermediate,
    the accepted step and X, U, V, block by block (tests/parity.py), for the sequential LQ kernel and both schedules of
    the associative scan (mpc_wrapper_srbd.py:68 binds limited_memory=False)."""
    from mpx_b200 import _lib
    from golden_check import load, solve, check_step
    g = load("srbd", case)
    mode = {"sequential": _lib.LQ_SEQUENTIAL, "scan_parallel": _lib.LQ_SCAN_PARALLEL, "scan_serial": _lib.LQ_SCAN_SERIAL}[lq]
    solver, X, U, V = solve(cfg, g, mode, eager)
    check_step(cfg, g, solver, X, U, V, "" if lq == "sequential" else "_scan", eager)


def test_srbd_per_env_W_equals_shared(problem):
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = 4
    inp = synthetic_inputs(problem, E, seed=5)
    solver = MpcSolver(cfg, E, limited_memory=True)
    a = [t.clone() for t in _run(solver, inp)]
    b = _run(solver, inp, W=np.tile(cfg.W, (E, 1, 1)))
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_srbd_closed_loop_wrapper_runs(problem):
    """mpc_wrapper_srbd.BatchedMPCControllerWrapper: 10 closed-loop ticks on the model stay finite and track height."""
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper
    E = 8
    mpc = BatchedMPCControllerWrapper(cfg, E)
    inp = synthetic_inputs(problem, E, seed=2)
    x0 = torch.as_tensor(inp["x0"], dtype=torch.float32).cuda()
    feet = torch.as_tensor(inp["parameter"][:, 0, 4:], dtype=torch.float32).cuda()
    cmd = torch.tensor([[0.3, 0, 0, 0, 0, 0, cfg.robot_height]] * E, dtype=torch.float32).cuda()
    contact = torch.ones(E, 4).cuda()
    mpc.liftoff = feet.clone()
    for _ in range(10):
        mpc.run(x0, cmd, feet, contact)
        x0 = mpc.X[:, 1, :].clone()
    torch.cuda.synchronize()
    assert torch.isfinite(mpc.X).all() and torch.isfinite(mpc.U).all()
    assert (mpc.grf[:, 2::3].sum(1) > 50).all()          # carrying a 24.6 kg robot
    assert (mpc.X[:, 1, 2] - cfg.robot_height).abs().max() < 0.1


def test_multi_iteration_device_loop(problem):
    """max_iter > 1 runs the SQP loop on the device and equals repeated single calls."""
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = 2
    inp = synthetic_inputs(problem, E, seed=7)
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous().cuda()
    one = MpcSolver(cfg, E, limited_memory=True)
    X, U, V = T(inp["X0"]), T(inp["U0"]), T(inp["V0"])
    for _ in range(3):
        X, U, V = one.solve(T(inp["reference"]), T(inp["parameter"]), T(cfg.W), T(inp["x0"]), X, U, V)
    three = MpcSolver(cfg, E, limited_memory=True, max_iter=3)
    X3, U3, V3 = three.solve(T(inp["reference"]), T(inp["parameter"]), T(cfg.W), T(inp["x0"]), T(inp["X0"]),
                             T(inp["U0"]), T(inp["V0"]))
    torch.cuda.synchronize()
    assert torch.equal(X, X3) and torch.equal(U, U3) and torch.equal(V, V3)
    assert (three.iters == 3).all()


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("cold", [False, True])
def test_srbd_merit_line_search_matches_oracle(problem, cold, eager):
    """MPXB_LS_MERIT: parallel_line_search on the augmented-Lagrangian merit function with merit_rho / slope
    (optimizers.py:185-217, 295-354; the call the reference keeps commented out at 637-669)."""
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = 3
    inp = synthetic_inputs(problem, E, seed=17, cold_start=cold)
    # start away from the warm start so that the multipliers and the constraint violation are not trivial
    rng = np.random.default_rng(5)
    inp["V0"] = rng.normal(0, 0.5, inp["V0"].shape)
    inp["X0"] = inp["X0"] + rng.normal(0, 0.02, inp["X0"].shape)
    solver = MpcSolver(cfg, E, limited_memory=True, eager_line_search=eager, line_search="merit")
    X, U, V = _run(solver, inp)
    for e in range(E):
        det = {}
        Xo, Uo, Vo = problem.mpc(True, inp["reference"][e], inp["parameter"][e], inp["x0"][e], inp["X0"][e],
                                 inp["U0"][e], inp["V0"][e], detail=det, line_search="merit")
        assert int(solver.alpha_index[e]) == det["alpha_index"], (int(solver.alpha_index[e]), det["alpha_index"], det["rejected"])
        npts = 11 if (eager or det["rejected"][0]) else 1
        assert _rel(solver.view("ls_vc")[e, :npts].cpu().numpy(), det["vc"].numpy()[:npts]) < 1e-3
        assert _rel(solver.view("ls_theta")[e, :npts].cpu().numpy(), det["theta"].numpy()[:npts]) < 1e-3
        assert _rel(X[e].cpu().numpy(), Xo.numpy()) < REL_TOL
        assert _rel(U[e].cpu().numpy(), Uo.numpy()) < REL_TOL
        assert _rel(V[e].cpu().numpy(), Vo.numpy()) < REL_TOL
