"""Parity at BASELINE.json's full size (Go2 whole-body x 1024 environments) through size-independent properties:
the oracle cannot evaluate a thousand whole-body solves in test time, so this checks what must hold for ANY batch --
the KKT conditions of the LQ sub-problem the kernels claim to have solved (kkt_helpers.py:71-109 /
old_optimazers.py:126-131), batch independence (an environment's result does not depend on the batch it is solved
in: the reference's vmap semantics, examples/multi_env.py:46-49), and lazy == eager line search."""
import numpy as np
import pytest
import torch

import mpx_b200.config.config_go2 as cfg

pytestmark = pytest.mark.gpu

E_FULL = 1024


def _inputs(E, seed=0):
    """The benchmark's synthetic closed-loop inputs (bench.synthetic_states + device reference generator), after two
    closed-loop ticks so that the warm start, multipliers and constraint violation are non-trivial."""
    import bench
    from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper
    st = bench.synthetic_states(cfg, E, seed=seed)
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).cuda().contiguous()
    w = BatchedMPCControllerWrapper(cfg, E, limited_memory=True)
    qpos, qvel = T(st["qpos"]), T(st["qvel"])
    feet = w.foot_positions(qpos)
    d = w.reset(w.make_data(), qpos, qvel, feet).replace(contact_time=T(st["timer"]))
    x0 = d.X0[:, 0, :].clone()
    shift = int(1 / (cfg.dt * cfg.mpc_frequency))
    for _ in range(2):
        d, _ = w.run(d, x0, T(st["cmd"]))
        x0 = w.X[:, shift, :].clone()
    d, _ = w.run(d, x0, T(st["cmd"]))
    # inputs of the third tick: reference / parameter it used, x0, and the warm start it started from
    return dict(reference=w.reference.clone(), parameter=w.parameter.clone(), x0=x0.clone(), W=T(cfg.W)), w


def _solve(E, inp, warm, **kw):
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    s = MpcSolver(cfg, E, limited_memory=True, **kw)
    X, U, V = s.solve(inp["reference"][:E], inp["parameter"][:E], inp["W"], inp["x0"][:E], warm[0][:E].contiguous(),
                      warm[1][:E].contiguous(), warm[2][:E].contiguous())
    torch.cuda.synchronize()
    return s, X, U, V


@pytest.fixture(scope="module")
def full():
    inp, w = _inputs(E_FULL)
    N, n, m = cfg.N, cfg.n, cfg.m
    gen = torch.Generator(device="cuda").manual_seed(3)
    # warm start: the closed-loop iterate, perturbed so that the step is not (numerically) zero
    X0 = w.X + 0.01 * torch.randn(w.X.shape, device="cuda", generator=gen)
    U0 = w.U + 0.1 * torch.randn(w.U.shape, device="cuda", generator=gen)
    V0 = w.V.clone()
    return inp, (X0, U0, V0)


def test_lq_kkt_conditions_full_batch(full):
    inp, warm = full
    s, X, U, V = _solve(E_FULL, inp, warm)
    assert int(s.status.sum()) == 0 and bool(torch.isfinite(X).all())
    A, B, Q, R, M = (s.view(k).double() for k in ("A", "B", "Q", "R", "M"))
    q, r, c, dX, dU, dV = (s.view(k).double() for k in ("q", "r", "c", "dX", "dU", "dV"))
    mv = lambda Mx, v: torch.einsum("etij,etj->eti", Mx, v)
    mtv = lambda Mx, v: torch.einsum("etji,etj->eti", Mx, v)
    # primal feasibility: dX_0 = c_0, dX_{t+1} = A dX_t + B dU_t + c_{t+1}
    dyn = mv(A, dX[:, :-1]) + mv(B, dU) + c[:, 1:] - dX[:, 1:]
    sx = mv(Q[:, :-1], dX[:, :-1]) + mv(M, dU) + mtv(A, dV[:, 1:]) - dV[:, :-1] + q[:, :-1]
    sxN = mv(Q[:, -1:], dX[:, -1:]) + q[:, -1:] - dV[:, -1:]
    su = mv(R, dU) + mtv(M, dX[:, :-1]) + mtv(B, dV[:, 1:]) + r
    # FP32 budget: dV_t = P_t dX_t + p_t is formed from a value function with entries up to ~1e7 (W spans
    # 1e-3 .. 1e5 and the hard contact rows are stiff), so the residuals are measured against the magnitude of the
    # products that cancel, |P_t| |dX_t| + |p_t|, not against the (much smaller) result
    P, pv = s.view("P").double(), s.view("p").double()
    amax = lambda *ts: max(1.0, *(float(t.abs().max()) for t in ts))
    scale_x = amax(dV, q, mv(P.abs(), dX.abs()) + pv.abs())
    scale_u = amax(r, mv(R, dU), mtv(M, dX[:, :-1]), mtv(B, dV[:, 1:]), mtv(B.abs(), (mv(P.abs(), dX.abs()) + pv.abs())[:, 1:]))
    res = dict(init=float((dX[:, 0] - c[:, 0]).abs().max()), dyn=float(dyn.abs().max()) / max(1.0, float(dX.abs().max())),
               stat_x=float(sx.abs().max()) / scale_x, stat_xN=float(sxN.abs().max()) / scale_x,
               stat_u=float(su.abs().max()) / scale_u)
    print("KKT residuals (relative):", res)
    assert res["init"] == 0.0
    assert res["dyn"] < 1e-4 and res["stat_x"] < 1e-4 and res["stat_xN"] < 1e-4 and res["stat_u"] < 1e-4, res


def test_batch_independence_and_lazy_equals_eager(full):
    inp, warm = full
    s_full, X, U, V = _solve(E_FULL, inp, warm)
    s_eager, Xe, Ue, Ve = _solve(E_FULL, inp, warm, eager_line_search=True)
    assert torch.equal(s_full.alpha_index, s_eager.alpha_index)
    assert torch.equal(X, Xe) and torch.equal(U, Ue) and torch.equal(V, Ve)
    # the first 8 environments solved on their own: same kernels, same per-env arithmetic -> bitwise equal
    s_small, Xs, Us, Vs = _solve(8, inp, warm)
    assert torch.equal(s_small.alpha_index, s_full.alpha_index[:8])
    assert torch.equal(Xs, X[:8]) and torch.equal(Us, U[:8]) and torch.equal(Vs, V[:8])
    hist = torch.bincount(s_full.alpha_index.long(), minlength=11).tolist()
    print("alpha index histogram:", hist)
    assert sum(hist) == E_FULL


def test_three_pass_line_search_equals_eager_h1():
    """Large H1 / Talos batches run the lazy search in three passes (alpha = 1, alpha = 1/2, the rest; mpxb_mpc_step):
    on a batch whose environments need different step sizes the result must equal the eager evaluation bitwise."""
    import importlib
    import os
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    cfg_h1 = importlib.import_module("mpx_b200.config.config_h1")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h1_warm.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture missing")
    g = np.load(path)
    E = 256
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda().contiguous()
    rep = lambda a: T(np.repeat(np.asarray(a)[:1], E, axis=0))
    gen = torch.Generator(device="cuda").manual_seed(11)
    # per-environment perturbation of the warm start, growing with the env index: from "alpha = 1 accepted" to rejected
    scale = torch.linspace(0.0, 0.3, E, device="cuda")[:, None, None]
    X0 = rep(g["X0"]); U0 = rep(g["U0"]); V0 = rep(g["V0"])
    X0 = X0 + scale * torch.randn(X0.shape, device="cuda", generator=gen)
    U0 = U0 + 10.0 * scale * torch.randn(U0.shape, device="cuda", generator=gen)
    out = {}
    for eager in (True, False):
        s = MpcSolver(cfg_h1, E, limited_memory=True, eager_line_search=eager)
        X, U, V = s.solve(rep(g["reference"]), rep(g["parameter"]), T(cfg_h1.W), rep(g["x0"]), X0, U0, V0)
        torch.cuda.synchronize()
        out[eager] = (s.alpha_index.clone(), X, U, V)
    hist = torch.bincount(out[True][0].long(), minlength=11).tolist()
    print("alpha index histogram (H1 x256):", hist)
    assert sum(1 for h in hist if h > 0) >= 3, hist  # the batch really exercises all three passes
    assert torch.equal(out[True][0], out[False][0])
    for a, b in zip(out[True][1:], out[False][1:]):
        assert torch.equal(a, b) or (torch.isnan(a) == torch.isnan(b)).all() and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
