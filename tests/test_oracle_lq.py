"""Pin the oracle's LQ solvers: closed-form problem of the reference's own test + dense-KKT known answers."""
import numpy as np
import pytest
import torch

from oracle import primal_dual_ilqr as pdi
from lq_problems import simple_lqr_problem, random_lq


def _solve(Q, q, R, r, M, A, B, c, x0, scan, order="sequential"):
    if scan:
        K, k, P, p = pdi.tvlqr_gpu(Q, q, R, r, M, A, B, c, order=order)
        dX, dU = pdi.rollout_gpu(K, k, x0, A, B, c)
    else:
        K, k, P, p = pdi.tvlqr(Q, q, R, r, M, A, B, c)
        dX, dU = pdi.rollout(K, k, x0, A, B, c)
    dV = pdi.dual_lqr(dX, P, p)
    return K, k, P, p, dX, dU, dV


@pytest.mark.parametrize("scan", [False, True])
def test_reference_simple_lqr_problem_matches_kkt(scan):
    """reference tests/simple_tests.py:70-113: the LQ problem has a unique solution; Riccati must equal dense KKT."""
    Q, q, R, r, M, A, B, c, x0 = simple_lqr_problem()
    K, k, P, p, X, U, V = _solve(Q, q, R, r, M, A, B, c, x0, scan)
    Xk, Uk, Vk = pdi.tvlqr_kkt(Q, q, R, r, M, A, B, c, x0)
    assert torch.allclose(X, Xk, rtol=1e-9, atol=1e-9)
    assert torch.allclose(U, Uk, rtol=1e-9, atol=1e-9)
    # multipliers: the reference's V pairs with constraints c_t = f - x_{t+1}; KKT rows use the same sign
    assert torch.allclose(V, Vk, rtol=1e-8, atol=1e-8)
    # dynamics are satisfied exactly (the reference test's `no_errors`)
    for t in range(2):
        assert torch.allclose(X[t + 1], A[t] @ X[t] + B[t] @ U[t] + c[t], atol=1e-9)


@pytest.mark.parametrize("n,m,T", [(2, 1, 2), (4, 1, 50), (8, 2, 30), (13, 12, 25)])
@pytest.mark.parametrize("scan", [False, True])
def test_random_lq_kkt(n, m, T, scan):
    prob = random_lq(n, m, T, seed=n * 100 + T)
    Q, q, R, r, M, A, B, c, x0 = prob
    K, k, P, p, X, U, V = _solve(*prob, scan)
    Xk, Uk, Vk = pdi.tvlqr_kkt(*prob)
    assert torch.allclose(X, Xk, rtol=1e-7, atol=1e-7)
    assert torch.allclose(U, Uk, rtol=1e-7, atol=1e-7)
    assert torch.allclose(V, Vk, rtol=1e-6, atol=1e-6)


def test_scan_orders_and_sequential_agree():
    prob = random_lq(8, 2, 21, seed=5)
    a = _solve(*prob, False)
    b = _solve(*prob, True, "sequential")
    c = _solve(*prob, True, "tree")
    for x, y, z in zip(a, b, c):
        assert torch.allclose(x, y, rtol=1e-8, atol=1e-8)
        assert torch.allclose(x, z, rtol=1e-8, atol=1e-8)


def test_two_by_two_closed_form_branch():
    """linalg_helpers.py:149-154,213-218: m == 2 uses the adjugate inverse; must agree with Cholesky."""
    G = torch.tensor([[4.0, 1.0], [1.0, 3.0]], dtype=torch.float64)
    b = torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=torch.float64)
    x = pdi.solve_spd(G, b)
    assert torch.allclose(G @ x, b, atol=1e-12)
