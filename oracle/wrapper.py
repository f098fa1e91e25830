"""Restatement of the wrapper-level steps either side of the solve (TEST INFRASTRUCTURE, see oracle/__init__.py).

* ``update_and_extract``  -- reference mpx/utils/mpc_wrapper.py:79-98 (batched wrapper) == 261-282 (single wrapper):
  warm-start shift, NaN guard, command extraction.
* ``srbd_shift``          -- reference mpx/utils/mpc_wrapper_srbd.py:135-139.
* ``offline_loop``        -- the host-side SQP loop of ``runOffline`` (mpc_wrapper.py:422-467): repeat ``mpc`` from
  the previous iterate; after every iteration evaluate (g, c) at the new iterate and stop when
  ``last_cost - g*g < 1e-3 and sum(c*c) < 1e-5`` or after ``max_iter + 1`` iterations (``i > max_iter``).
"""
import numpy as np
import torch


def update_and_extract(U, X, V, x0, X0, U0, shift, n_joints, u_ref):
    """mpc_wrapper.py:79-98.  Returns (new_U0, new_X0, new_V0, tau, q, dq) as numpy float64."""
    U, X, V, x0, X0, U0 = (np.asarray(a, dtype=np.float64) for a in (U, X, V, x0, X0, U0))
    N, n = U.shape[0], X.shape[1]
    nj = n_joints
    if np.isnan(U[0, 0]):
        new_U0 = np.tile(np.asarray(u_ref, dtype=np.float64), (N, 1))
        new_X0 = np.tile(x0, (N + 1, 1))
        new_V0 = np.zeros((N + 1, n))
        tau = U0[1, :nj]
        q = X0[1, 7:nj + 7]
        dq = X0[1, 13 + nj:2 * nj + 13]
    else:
        new_U0 = np.concatenate([U[shift:], np.tile(U[-1:], (shift, 1))])
        new_X0 = np.concatenate([X[shift:], np.tile(X[-1:], (shift, 1))])
        new_V0 = np.concatenate([V[shift:], np.tile(V[-1:], (shift, 1))])
        tau = U[0, :nj]
        q = X[0, 7:nj + 7]
        dq = X[1, 13 + nj:2 * nj + 13]
    return new_U0, new_X0, new_V0, tau, q, dq


def srbd_shift(U, X, V, shift=1):
    """mpc_wrapper_srbd.py:135-139 (per environment; the reference tiles the last row ``shift`` times)."""
    U, X, V = (np.asarray(a, dtype=np.float64) for a in (U, X, V))
    sh = lambda a: np.concatenate([a[shift:], np.tile(a[-1:], (shift, 1))])
    return sh(U), sh(X), sh(V)


def offline_loop(bp, limited_memory, reference, parameter, x0, X, U, V, max_iter, W=None):
    """mpc_wrapper.py:422-467 for a batch of environments (each with its own stop decision).

    ``bp`` is an ``oracle.batched.BatchedProblem``.  Returns X, U, V [E, ...] (torch), iters [E] = number of ``mpc``
    calls each environment made, and the per-iteration (g, sum c^2) history."""
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=bp.dtype)
    Wt = bp.pr.W if W is None else t(W)
    reference, parameter, x0, X, U, V = map(t, (reference, parameter, x0, X, U, V))
    E = X.shape[0]
    active = torch.ones(E, dtype=torch.bool)
    iters = torch.zeros(E, dtype=torch.int64)
    last_cost = torch.full((E,), 1e10, dtype=bp.dtype)
    hist = []
    i = 0
    while bool(active.any()):
        idx = torch.nonzero(active).flatten()
        Xn, Un, Vn = bp.mpc(limited_memory, reference[idx], parameter[idx], x0[idx], X[idx], U[idx], V[idx], W=Wt)
        X, U, V = X.clone(), U.clone(), V.clone()
        X[idx], U[idx], V[idx] = Xn, Un, Vn
        g, c, _ = bp.evaluate(Wt, reference[idx], parameter[idx], x0[idx], Xn, Un)
        l2 = g * g
        c2 = (c * c).sum(dim=(1, 2))
        hist.append((idx.clone(), g.clone(), c2.clone()))
        iters[idx] += 1
        i += 1
        stop = (last_cost[idx] - l2 < 1e-3) & (c2 < 1e-5)
        if i > max_iter:
            stop = torch.ones_like(stop)
        last_cost[idx] = l2
        active[idx[stop]] = False
    return X, U, V, iters, hist
