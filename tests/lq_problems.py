"""Shared LQ test problems (test infrastructure)."""
import numpy as np
import torch


def simple_lqr_problem(dtype=torch.float64):
    """The time-varying LQ problem of the reference's tests/simple_tests.py:70-113 (T=2, n=2, m=1), as LQ data."""
    T, n, m = 2, 2, 1
    A = np.stack([6 * t * np.ones([2, 2]) + np.arange(4).reshape([2, 2]) for t in range(T)])
    B = np.stack([6 * t * np.ones([2, 1]) + np.arange(4, 6).reshape([2, 1]) for t in range(T)])
    c = np.stack([6 * t * np.ones(2) + np.arange(6, 8) for t in range(T)])
    Q = np.stack([(t + 1) * np.diag(np.arange(1, 3)) for t in range(T + 1)]).astype(float)
    R = np.stack([(t + 1) * np.array([3.0]).reshape([1, 1]) for t in range(T)])
    q = np.stack([-(t + 1) * np.arange(91, 93) for t in range(T + 1)]).astype(float)
    r = np.stack([-(t + 1) * np.array([99.0]) for t in range(T)])
    M = np.stack([(t + 1) / 100.0 * np.arange(55, 57).reshape([2, 1]) for t in range(T)])
    x0 = np.array([-11.0, -22.0])
    tt = lambda a: torch.as_tensor(a, dtype=dtype)
    return tuple(tt(a) for a in (Q, q, R, r, M, A, B, c, x0))


def random_lq(n, m, T, seed, dtype=torch.float64, scale_a=0.6):
    """Random well-posed LQ problem: SPD [Q M; M^T R], contractive-ish A."""
    rng = np.random.default_rng(seed)
    Q, R, M = [], [], []
    for t in range(T + 1):
        G = rng.normal(size=(n + m, n + m))
        H = G @ G.T / (n + m) + 0.5 * np.eye(n + m)
        Q.append(H[:n, :n]); R.append(H[n:, n:]); M.append(0.3 * H[:n, n:])
    A = rng.normal(size=(T, n, n)) * scale_a / np.sqrt(n) + np.eye(n) * 0.5
    B = rng.normal(size=(T, n, m)) / np.sqrt(n)
    c = rng.normal(size=(T, n)) * 0.1
    q = rng.normal(size=(T + 1, n))
    r = rng.normal(size=(T, m))
    x0 = rng.normal(size=n)
    tt = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    return tt(Q), tt(q), tt(R[:T]), tt(r), tt(M[:T]), tt(A), tt(B), tt(c), tt(x0)
