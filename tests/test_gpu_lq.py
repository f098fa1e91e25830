"""GPU parity of the LQ kernels (mpxb_tvlqr through the C ABI) against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import primal_dual_ilqr as pdi
from lq_problems import simple_lqr_problem, random_lq

pytestmark = pytest.mark.gpu


def _to_gpu(prob, E):
    Q, q, R, r, M, A, B, c, x0 = prob
    cc = torch.cat([x0[None], c], 0)
    f = lambda t: t.to(torch.float32).unsqueeze(0).repeat(E, *([1] * t.dim())).contiguous().cuda()
    return tuple(f(t) for t in (Q, q, R, r, M, A, B, cc))


def _oracle(prob, scan):
    Q, q, R, r, M, A, B, c, x0 = prob
    if scan:
        K, k, P, p = pdi.tvlqr_gpu(Q, q, R, r, M, A, B, c)
        dX, dU = pdi.rollout_gpu(K, k, x0, A, B, c)
    else:
        K, k, P, p = pdi.tvlqr(Q, q, R, r, M, A, B, c)
        dX, dU = pdi.rollout(K, k, x0, A, B, c)
    dV = pdi.dual_lqr(dX, P, p)
    slope = torch.sum(q * dX) + torch.sum(r * dU)
    return K, k, P, p, dX, dU, dV, slope


def _check(got, want, tol):
    """Per-knot relative error (tests/parity.py; one block per axis: these LQ problems have no physical layout)."""
    import parity
    names = ["K", "k", "P", "p", "dX", "dU", "dV", "slope"]
    n, m = want[0].shape[2], want[0].shape[1]
    segs = parity.segments_generic(n, m)
    for nm, g, w in zip(names, got, want):
        g = g[0].double().cpu().numpy()
        if nm == "slope":
            err = parity.scalar_rel_err(g, float(w))
        else:
            err = parity.block_rel_err(nm, g, w.numpy(), segs)
        assert err < tol, f"{nm}: rel err {err:.3e}"


@pytest.mark.parametrize("limited_memory", [True, False])
def test_reference_simple_lqr_problem(limited_memory):
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    prob = simple_lqr_problem()
    got = tvlqr_solve(*_to_gpu(prob, 3), limited_memory=limited_memory)
    _check(got, _oracle(prob, not limited_memory), 2e-5)


@pytest.mark.parametrize("n,m,T", [(4, 1, 50), (8, 2, 30), (13, 12, 25), (61, 12, 25), (75, 19, 25), (81, 46, 25), (81, 46, 60)])
@pytest.mark.parametrize("lq_mode", [0, 1, 2, 3], ids=["sequential", "scan", "scan_serial", "scan_parallel"])
def test_random_lq(n, m, T, lq_mode):
    """Every LQ kernel: sequential Riccati, and the associative scan in both schedules (parallel over the horizon,
    one CTA per knot and log2 levels; serial, one CTA per environment) -- `scan` picks the parallel one at E = 2."""
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    prob = random_lq(n, m, T, seed=n + T)
    got = tvlqr_solve(*_to_gpu(prob, 2), lq_mode=lq_mode)
    # FP32 budget: the float32 oracle's own error against float64 is ~1e-5 on these problems
    _check(got, _oracle(prob, lq_mode != 0), 2e-4)
    # batch entries are independent and identical
    for g in got:
        assert torch.equal(g[0], g[1])


def test_small_system_kernel_and_its_fallbacks():
    """Sequential LQ of the SRBD size: the whole-horizon-in-shared-memory kernel (csrc/lq_small.cuh) serves horizons up
    to N = 47 and batches up to 6 waves of CTAs, the streaming kernel (csrc/lq.cuh) everything else.  Every route must
    meet the oracle, and a batch must not depend on which route its size selects beyond rounding."""
    from mpx_b200 import _lib
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    for T, E in [(1, 1), (25, 3), (47, 2), (48, 2), (90, 1)]:      # 47 is the last horizon that fits, 48 streams
        prob = random_lq(13, 12, T, seed=100 + T)
        got = tvlqr_solve(*_to_gpu(prob, E), lq_mode=_lib.LQ_SEQUENTIAL)
        _check(got, _oracle(prob, False), 2e-4)
        for g in got:
            assert torch.equal(g[0], g[E - 1])
    prob = random_lq(13, 12, 25, seed=7)
    small = tvlqr_solve(*_to_gpu(prob, 4), lq_mode=_lib.LQ_SEQUENTIAL)         # lq_small
    large = tvlqr_solve(*_to_gpu(prob, 6 * 148 + 1), lq_mode=_lib.LQ_SEQUENTIAL)   # streaming kernel
    _check(large, _oracle(prob, False), 2e-4)
    for a, b in zip(small, large):
        assert torch.allclose(a[0], b[0], rtol=2e-4, atol=2e-4 * float(a[0].abs().max()))


def test_parallel_scan_long_horizon_and_odd_lengths():
    """Hillis-Steele levels with N + 1 not a power of two, and a long horizon (Talos sizes, N = 200)."""
    from mpx_b200 import _lib
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    for n, m, T in [(8, 2, 1), (8, 2, 2), (8, 2, 7), (13, 12, 33), (81, 46, 200)]:
        prob = random_lq(n, m, T, seed=3 * n + T)
        got = tvlqr_solve(*_to_gpu(prob, 1), lq_mode=_lib.LQ_SCAN_PARALLEL)
        _check(got, _oracle(prob, True), 5e-4 if T > 100 else 2e-4)


def test_parallel_scan_rejected_for_large_batches():
    from mpx_b200 import _lib
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    prob = random_lq(4, 1, 3, seed=1)
    with pytest.raises(RuntimeError):
        tvlqr_solve(*_to_gpu(prob, 65), lq_mode=_lib.LQ_SCAN_PARALLEL)


def test_unsupported_dims_rejected():
    from mpx_b200.primal_dual_ilqr.optimizers import tvlqr_solve
    prob = random_lq(5, 3, 4, seed=1)
    with pytest.raises(RuntimeError, match="UNSUPPORTED"):
        tvlqr_solve(*_to_gpu(prob, 1), limited_memory=True)
