"""The vectorised oracle (oracle/batched.py, used for the golden fixtures and the CPU baseline) against the per-knot
loop form (oracle/primal_dual_ilqr.py) it restates, to float64 rounding; and the committed fixtures against a fresh
run of the script that made them."""
import importlib
import os

import numpy as np
import pytest
import torch
from torch.func import jacrev

from oracle.batched import BatchedProblem
from oracle.problems import Problem, synthetic_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("limited_memory", [True, False])
def test_batched_mpc_equals_loop_mpc_srbd(limited_memory):
    cfg = importlib.import_module("mpx_b200.config.config_srbd")
    pr = Problem(cfg)
    bp = BatchedProblem(pr)
    E = 3
    inp = synthetic_inputs(pr, E, seed=11)
    detb = {}
    Xb, Ub, Vb = bp.mpc(limited_memory, inp["reference"], inp["parameter"], inp["x0"], inp["X0"], inp["U0"], inp["V0"],
                        detail=detb)
    for e in range(E):
        det = {}
        X, U, V = pr.mpc(limited_memory, inp["reference"][e], inp["parameter"][e], inp["x0"][e], inp["X0"][e],
                         inp["U0"][e], inp["V0"][e], detail=det)
        for nm in ("A", "B", "Q", "R", "M", "q", "r", "c", "K", "k", "P", "p", "dX", "dU", "dV", "theta", "phi"):
            assert _rel(detb[nm][e], det[nm]) < 1e-11, nm
        assert int(detb["alpha_index"][e]) == det["alpha_index"]
        assert _rel(Xb[e], X) < 1e-10 and _rel(Ub[e], U) < 1e-10 and _rel(Vb[e], V) < 1e-9


@pytest.mark.parametrize("robot", ["go2", "talos"])
def test_batched_derivatives_equal_per_knot_jacrev(robot):
    """Whole-body: the (env x knot)-vmapped derivatives against per-knot jacrev of the same restated functions at a
    few knots (the full loop form costs ~40 s per environment)."""
    from oracle import primal_dual_ilqr as pdi
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    pr = Problem(cfg)
    bp = BatchedProblem(pr)
    inp = synthetic_inputs(pr, 2, seed=5)
    rng = np.random.default_rng(0)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64)
    X = t(inp["X0"] + rng.normal(0, 1e-2, inp["X0"].shape))
    U = t(inp["U0"] + rng.normal(0, 1.0, inp["U0"].shape))
    V = t(rng.normal(0, 1.0, inp["V0"].shape))
    ref, par, x0 = t(inp["reference"]), t(inp["parameter"]), t(inp["x0"])
    W, N = pr.W, cfg.N
    Q, R, M, q, r, A, B = bp.linearize_quadratize(W, ref, par, x0, X, U, V)
    g, c, _ = bp.evaluate(W, ref, par, x0, X, U)
    e = 1
    cost = lambda x, u, k: pr.cost(W, ref[e], x, u, k)
    dyn = lambda x, u, k: pr.dynamics(x, u, k, par[e])
    hes = lambda x, u, k: pr.hessian_approx(W, ref[e], x, u, k)
    go, co = pdi.model_evaluator(cost, dyn, x0[e], X[e], U[e])
    assert _rel(g[e], go) < 1e-12 and _rel(c[e], co) < 1e-12
    zero_u = torch.zeros_like(U[e, 0])
    for k in (0, 7, N - 1, N):
        x, u = X[e, k], (U[e, k] if k < N else zero_u)
        Qo, Ro, Mo = hes(x, u, k)
        assert _rel(Q[e, k], Qo) < 1e-12
        gx = jacrev(lambda x_: cost(x_, u, k))(x)
        if k < N:
            Ao, Bo = jacrev(lambda x_: dyn(x_, u, k))(x), jacrev(lambda u_: dyn(x, u_, k))(u)
            gu = jacrev(lambda u_: cost(x, u_, k))(u)
            assert _rel(A[e, k], Ao) < 1e-11 and _rel(B[e, k], Bo) < 1e-11
            assert _rel(R[e, k], Ro) < 1e-12 and _rel(M[e, k], Mo) < 1e-12
            assert _rel(q[e, k], gx + Ao.T @ V[e, k + 1] - V[e, k]) < 1e-11
            assert _rel(r[e, k], gu + Bo.T @ V[e, k + 1]) < 1e-11
        else:
            assert _rel(q[e, k], gx - V[e, k]) < 1e-11


def test_committed_fixture_is_what_the_script_makes():
    """tests/golden/srbd_warm.npz equals a fresh evaluation of the oracle on the fixture's own inputs."""
    cfg = importlib.import_module("mpx_b200.config.config_srbd")
    g = np.load(os.path.join(GOLDEN, "srbd_warm.npz"))
    bp = BatchedProblem(Problem(cfg))
    for lm, sfx in ((True, ""), (False, "_scan")):
        det = {}
        X, U, V = bp.mpc(lm, g["reference"], g["parameter"], g["x0"], g["X0"], g["U0"], g["V0"], detail=det)
        assert det["alpha_index"].tolist() == g["alpha_index" + sfx].tolist()
        for nm, got in (("X", X), ("U", U), ("V", V), ("K", det["K"]), ("P", det["P"]), ("dV", det["dV"])):
            ref = torch.as_tensor(g[nm + sfx], dtype=torch.float64)
            assert _rel(got, ref) < 1e-6, nm   # large arrays are stored as float32


def test_fixture_coverage():
    """>= 4 environments per fixture; every robot has an environment whose accepted step is alpha <= 1/4; a cold start
    (identity quaternion) per robot; the float32 oracle agreed with the float64 one on every accepted index."""
    for robot in ("srbd", "go2", "aliengo", "h1", "talos"):
        amax = 0
        for case in ("warm", "second", "cold", "far"):
            g = np.load(os.path.join(GOLDEN, f"{robot}_{case}.npz"))
            assert g["x0"].shape[0] >= 4
            assert "P_scan" in g.files and "alpha_index_scan" in g.files
            amax = max(amax, int(g["alpha_index"].max()))
        assert amax >= 2, (robot, amax)
        cold = np.load(os.path.join(GOLDEN, f"{robot}_cold.npz"))
        assert np.all(cold["x0"][:, 3:7] == np.array([1.0, 0, 0, 0]))
