"""The legacy SQP loop on the device: sequential Armijo `line_search` on the augmented-Lagrangian merit
function (old_optimazers.py:173-251 == optimizers.py:219-293) with the legacy slope (old_optimazers.py:154-170) and the
legacy loop's `continuation_criterion` (old_optimazers.py:435-447), selected by MPXB_LS_ARMIJO + MPXB_STOP_LEGACY,
against the oracle's restatement of that loop (oracle/legacy.py, pinned on the reference's own acrobot / LQ tests in
tests/test_oracle_mjx.py) run around the MPC path's Gauss-Newton direction.

Iteration counts are compared exactly where the decision is not at the resolution of FP32: the criterion's slope
threshold is 1e-4 on a merit of ~3e3, i.e. a few float32 ulps, so in the last iterations before convergence the
float32 loop may stop one or two iterations earlier or later than the float64 oracle (checked as such below)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import legacy
from oracle.problems import Problem, synthetic_inputs

T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()


def _solve(cfg, inp, E, max_iter):
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    s = MpcSolver(cfg, E, limited_memory=True, max_iter=max_iter, line_search="armijo", stop_rule="legacy")
    X, U, V = s.solve(T(inp["reference"]), T(inp["parameter"]), T(cfg.W), T(inp["x0"]), T(inp["X0"]), T(inp["U0"]), T(inp["V0"]))
    torch.cuda.synchronize()
    return s, X, U, V


def _inputs(robot, E):
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    pr = Problem(cfg)
    inp = synthetic_inputs(pr, E, seed=11)
    if robot == "srbd":   # start off the warm start so that multipliers and defects are not trivial
        rng = np.random.default_rng(3)
        inp["V0"] = rng.normal(0, 0.5, inp["V0"].shape)
        inp["X0"] = inp["X0"] + rng.normal(0, 0.02, inp["X0"].shape)
    return cfg, pr, inp


def _oracle(pr, inp, e, max_iter, start=None):
    trace = []
    X0, U0, V0 = (inp["X0"][e], inp["U0"][e], inp["V0"][e]) if start is None else start
    out = legacy.legacy_loop_on_problem(pr, inp["reference"][e], inp["parameter"][e], inp["x0"][e], X0, U0, V0,
                                        max_iterations=max_iter, trace=trace)
    return out, trace


rel = lambda a, b: float(np.abs(a.cpu().numpy() - b.numpy()).max() / max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("robot,E,max_iter", [("srbd", 3, 6), ("go2", 1, 3)])
def test_legacy_loop_matches_oracle(robot, E, max_iter):
    """Away from convergence: the same number of steps, the same accepted step size in every iteration (read back by
    running the device loop with max_iter = 1, 2, ...), the same iterates."""
    cfg, pr, inp = _inputs(robot, E)
    alphas = []
    for k in range(1, max_iter + 1):
        s, X, U, V = _solve(cfg, inp, E, k)
        assert (s.iters == k).all()
        alphas.append(s.alpha_index.cpu().numpy().copy())
    for e in range(E):
        (Xo, Uo, Vo, iters, g, c, no_errors), trace = _oracle(pr, inp, e, max_iter)
        assert iters == max_iter == int(s.iters[e])
        assert [int(a[e]) for a in alphas] == [int(round(-np.log2(t["alpha"]))) for t in trace], (e, trace)
        assert int(s.status[e]) == 0
        assert rel(X[e], Xo) < 2e-3 and rel(U[e], Uo) < 2e-3, (e, rel(X[e], Xo), rel(U[e], Uo))


def test_legacy_loop_stops_by_the_continuation_criterion():
    """SRBD with room to converge: device and oracle both leave the loop through the criterion (not max_iterations),
    within two iterations of each other, at the same optimum; and from a converged iterate no step is taken at all
    (the criterion is tested BEFORE the first step, old_optimazers.py:449-471)."""
    cfg, pr, inp = _inputs("srbd", 3)
    s, X, U, V = _solve(cfg, inp, 3, 40)
    conv = []
    for e in range(3):
        (Xo, Uo, Vo, iters, g, c, no_errors), trace = _oracle(pr, inp, e, 40)
        assert iters < 40 and int(s.iters[e]) < 40, (iters, int(s.iters[e]))
        assert abs(int(s.iters[e]) - iters) <= 2, (e, int(s.iters[e]), iters, [t["merit_slope"] for t in trace])
        assert rel(X[e], Xo) < 2e-3 and rel(U[e], Uo) < 2e-3
        # ten more float64 iterations: decisively converged (slope ~1e-7)
        (Xc, Uc, Vc, it2, *_), _ = _oracle(pr, inp, e, 10, start=(Xo, Uo, Vo))
        conv.append((Xc.numpy(), Uc.numpy(), Vc.numpy()))
    inp2 = dict(inp, X0=np.stack([c[0] for c in conv]), U0=np.stack([c[1] for c in conv]), V0=np.stack([c[2] for c in conv]))
    s2, X2, U2, V2 = _solve(cfg, inp2, 3, 40)
    assert (s2.iters == 0).all() and (s2.status == 0).all()
    assert torch.equal(X2, T(inp2["X0"])) and torch.equal(U2, T(inp2["U0"])) and torch.equal(V2, T(inp2["V0"]))


def test_legacy_stop_rule_needs_a_merit_search():
    import mpx_b200.config.config_srbd as cfg
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    with pytest.raises(ValueError):
        MpcSolver(cfg, 1, limited_memory=True, max_iter=3, line_search="filter", stop_rule="legacy")
