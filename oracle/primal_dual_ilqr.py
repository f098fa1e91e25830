"""Restatement of the reference's primal-dual iLQR step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows mpx/primal_dual_ilqr/primal_dual_ilqr/{optimizers,primal_tvlqr,dual_tvlqr,linalg_helpers,kkt_helpers}.py.
``cost(x,u,t)``, ``dynamics(x,u,t)``, ``hessian_approx(x,u,t)`` take a python int ``t``.
The trajax combinators the reference imports (``vectorize`` / ``linearize`` / ``quadratize`` -- trajax is an
unpinned third-party dependency absent from /root/reference) are restated as python loops over knots with
``torch.func.jacrev``: linearize = (d/d arg0, d/d arg1), quadratize = (hess_xx, hess_uu, d(grad_x)/du).
"""
import torch
from torch.func import jacrev, hessian


# ----------------------------------------------------------------------------- LQ solvers
def _sym(x):
    return 0.5 * (x + x.T)


def solve_spd(A, b):
    """linalg_helpers.py:213-218: 2x2 closed form, else Cholesky (cho_factor/cho_solve)."""
    if A.shape[0] == 2:
        a, bb, c, d = A.flatten()
        det = a * d - bb * c
        inv = (1.0 / det) * torch.stack([torch.stack([d, -bb]), torch.stack([-c, a])])
        return inv @ b
    L = torch.linalg.cholesky(A)
    return torch.cholesky_solve(b, L)


def lqr_step(P, p, Q, q, R, r, M, A, B, c):
    """primal_tvlqr.py:11-52."""
    AtP = A.T @ P
    AtPA = _sym(AtP @ A)
    BtP = B.T @ P
    BtPA = BtP @ A
    H = BtPA + M.T
    h = B.T @ p + BtP @ c + r
    G = _sym(R + BtP @ B)
    K_k = solve_spd(G, -torch.cat([H, h.reshape(-1, 1)], dim=1))
    K, k = K_k[:, :-1], K_k[:, -1]
    P = _sym(Q + AtPA + K.T @ H)
    p = q + A.T @ p + AtP @ c + K.T @ h
    return K, k, P, p


def tvlqr(Q, q, R, r, M, A, B, c):
    """primal_tvlqr.py:103-144: sequential Riccati recursion."""
    T = Q.shape[0] - 1
    P, p = Q[T], q[T]
    Ks, ks, Ps, ps = [None] * T, [None] * T, [None] * (T + 1), [None] * (T + 1)
    Ps[T], ps[T] = P, p
    for t in range(T - 1, -1, -1):
        K, k, P, p = lqr_step(P, p, Q[t], q[t], R[t], r[t], M[t], A[t], B[t], c[t])
        Ks[t], ks[t], Ps[t], ps[t] = K, k, P, p
    return torch.stack(Ks), torch.stack(ks), torch.stack(Ps), torch.stack(ps)


def _scan_combine(nxt, prev, n):
    """primal_tvlqr.py:364-394: combine ``prev`` (earlier, 'l') with ``nxt`` (later, 'r')."""
    A_l, c_l, C_l, p_l, P_l = prev
    A_r, c_r, C_r, p_r, P_r = nxt
    eye = torch.eye(n, dtype=A_l.dtype)
    ArI = A_r @ torch.linalg.inv(eye + C_l @ P_r)
    AlT = A_l.T @ torch.linalg.inv(eye + P_r @ C_l)
    return (ArI @ A_l,
            ArI @ (c_l - C_l @ p_r) + c_r,
            ArI @ C_l @ A_r.T + C_r,
            AlT @ (p_r + P_r @ c_l) + p_l,
            AlT @ P_r @ A_l + P_l)


def tvlqr_gpu(Q, q, R, r, M, A, B, c, order="sequential"):
    """primal_tvlqr.py:337-475: associative-scan form.  The suffix combine is evaluated right-to-left
    (``order='sequential'``) or as a balanced tree (``'tree'``); both orders are the same associative
    product and differ only in rounding, like lax.associative_scan's own schedule."""
    T = Q.shape[0] - 1
    n = Q.shape[1]
    elems = []
    for t in range(T):
        L = torch.linalg.cholesky(R[t])
        Rinv = torch.cholesky_solve(torch.eye(R[t].shape[0], dtype=R.dtype), L)
        BRinv = B[t] @ Rinv
        MRinv = M[t] @ Rinv
        elems.append((A[t] - BRinv @ M[t].T, c[t] - BRinv @ r[t], BRinv @ B[t].T,
                      q[t] - MRinv @ r[t], Q[t] - MRinv @ M[t].T))
    z = torch.zeros(n, n, dtype=Q.dtype)
    elems.append((z, torch.zeros(n, dtype=Q.dtype), z, q[T], Q[T]))
    res = [None] * (T + 1)
    if order == "sequential":
        acc = elems[T]
        res[T] = acc
        for t in range(T - 1, -1, -1):
            acc = _scan_combine(acc, elems[t], n)
            res[t] = acc
    else:
        # Hillis-Steele suffix scan (log-depth), the schedule the CUDA scan kernel uses
        cur = list(elems)
        d = 1
        while d <= T:
            nxt = list(cur)
            for t in range(T + 1 - d):
                nxt[t] = _scan_combine(cur[t + d], cur[t], n)
            cur = nxt
            d *= 2
        res = cur
    P = torch.stack([e[4] for e in res])
    p = torch.stack([e[3] for e in res])
    Ks, ks = [], []
    for t in range(T):
        BtP = B[t].T @ P[t + 1]
        H = BtP @ A[t] + M[t].T
        h = B[t].T @ p[t + 1] + BtP @ c[t] + r[t]
        K_k = torch.linalg.solve(R[t] + BtP @ B[t], -torch.cat([H, h.reshape(-1, 1)], dim=1))
        Ks.append(K_k[:, :-1])
        ks.append(K_k[:, -1])
    return torch.stack(Ks), torch.stack(ks), P, p


def rollout(K, k, x0, A, B, c):
    """primal_tvlqr.py:478-498."""
    T = c.shape[0]
    X, U = [x0], []
    x = x0
    for t in range(T):
        u = K[t] @ x + k[t]
        x = A[t] @ x + B[t] @ u + c[t]
        X.append(x)
        U.append(u)
    return torch.stack(X), torch.stack(U)


def rollout_gpu(K, k, x0, A, B, c):
    """primal_tvlqr.py:523-549: prefix composition of affine maps (F, f)."""
    T = c.shape[0]
    F = [A[t] + B[t] @ K[t] for t in range(T)]
    f = [c[t] + B[t] @ k[t] for t in range(T)]
    comp = [(F[0], f[0])]
    for t in range(1, T):
        Fp, fp = comp[-1]
        comp.append((F[t] @ Fp, f[t] + F[t] @ fp))
    X = [x0] + [comp[t][0] @ x0 + comp[t][1] for t in range(T)]
    U = [K[t] @ X[t] + k[t] for t in range(T)]
    return torch.stack(X), torch.stack(U)


def dual_lqr(X, P, p):
    """dual_tvlqr.py:6-19."""
    return torch.stack([P[t] @ X[t] + p[t] for t in range(X.shape[0])])


def tvlqr_kkt(Q, q, R, r, M, A, B, c, x0):
    """Dense KKT solve of the LQ sub-problem (idea of kkt_helpers.py:71-109), used as a known-answer check.

    minimise  sum_t 1/2 x'Qx + 1/2 u'Ru + x'Mu + q'x + r'u   s.t.  x_0 = x0,  x_{t+1} = A x_t + B u_t + c_t.
    Returns (X, U, V) with V the multipliers of [x0 - x_0 ; A x + B u + c - x_{t+1}].
    """
    T = R.shape[0]
    n, m = B.shape[1], B.shape[2]
    nz = (T + 1) * n + T * m
    nl = (T + 1) * n
    ix = lambda t: slice(t * n, (t + 1) * n)
    iu = lambda t: slice((T + 1) * n + t * m, (T + 1) * n + (t + 1) * m)
    H = torch.zeros(nz, nz, dtype=Q.dtype)
    g = torch.zeros(nz, dtype=Q.dtype)
    C = torch.zeros(nl, nz, dtype=Q.dtype)
    d = torch.zeros(nl, dtype=Q.dtype)
    for t in range(T + 1):
        H[ix(t), ix(t)] = Q[t]
        g[ix(t)] = q[t]
        if t < T:
            H[iu(t), iu(t)] = R[t]
            H[ix(t), iu(t)] = M[t]
            H[iu(t), ix(t)] = M[t].T
            g[iu(t)] = r[t]
    # constraint rows: row 0: x0 - x_0 = 0 ; row t+1: A x_t + B u_t + c_t - x_{t+1} = 0
    eye = torch.eye(n, dtype=Q.dtype)
    C[ix(0), ix(0)] = -eye
    d[ix(0)] = x0
    for t in range(T):
        C[ix(t + 1), ix(t)] = A[t]
        C[ix(t + 1), iu(t)] = B[t]
        C[ix(t + 1), ix(t + 1)] = -eye
        d[ix(t + 1)] = c[t]
    KKT = torch.cat([torch.cat([H, C.T], dim=1), torch.cat([C, torch.zeros(nl, nl, dtype=Q.dtype)], dim=1)])
    rhs = -torch.cat([g, d])
    sol = torch.linalg.solve(KKT, rhs)
    X = sol[:(T + 1) * n].reshape(T + 1, n)
    U = sol[(T + 1) * n:nz].reshape(T, m)
    V = sol[nz:].reshape(T + 1, n)
    return X, U, V


# ----------------------------------------------------------------------------- SQP step
def model_evaluator(cost, dynamics, x0, X, U):
    """optimizers.py:580-602."""
    T = U.shape[0]
    zero_u = torch.zeros_like(U[0])
    g = sum(cost(X[t], U[t] if t < T else zero_u, t) for t in range(T + 1))
    c = torch.stack([x0 - X[0]] + [dynamics(X[t], U[t], t) - X[t + 1] for t in range(T)])
    return g, c


def linearize_quadratize(cost, dynamics, hessian_approx, x0, X, U, V):
    """The derivative half of compute_search_direction (optimizers.py:150-173): Q,R,M,q,r,A,B."""
    T = U.shape[0]
    zero_u = torch.zeros_like(U[0])
    zero_v = torch.zeros_like(V[0])
    Qs, Rs, Ms, qs, rs, As, Bs = [], [], [], [], [], [], []
    for t in range(T + 1):
        x, u = X[t], (U[t] if t < T else zero_u)
        if hessian_approx is None:
            ct = lambda x_, u_: cost(x_, u_, t)
            Qt = hessian(ct, 0)(x, u)
            Rt = hessian(ct, 1)(x, u)
            Mt = jacrev(jacrev(ct, 0), 1)(x, u)
        else:
            Qt, Rt, Mt = hessian_approx(x, u, t)
        v = V[t + 1] if t < T else zero_v
        v_prev = V[t]

        def lag(x_, u_):
            c3 = torch.dot(v_prev, (x0 - x_) if t == 0 else -x_)
            # at t == T the reference multiplies the dynamics by the zero-padded multiplier (optimizers.py:168)
            c2 = torch.dot(v, dynamics(x_, u_, t)) if t < T else 0.0
            return cost(x_, u_, t) + c2 + c3
        qt = jacrev(lag, 0)(x, u)
        rt = jacrev(lag, 1)(x, u)
        Qs.append(Qt); qs.append(qt)
        if t < T:
            f = lambda x_, u_: dynamics(x_, u_, t)
            As.append(jacrev(f, 0)(x, u)); Bs.append(jacrev(f, 1)(x, u))
            Rs.append(Rt); Ms.append(Mt); rs.append(rt)
    st = torch.stack
    return st(Qs), st(Rs), st(Ms), st(qs), st(rs), st(As), st(Bs)


def compute_search_direction(cost, dynamics, hessian_approx, limited_memory, x0, X, U, V, c, detail=None):
    """optimizers.py:121-184."""
    Q, R, M, q, r, A, B = linearize_quadratize(cost, dynamics, hessian_approx, x0, X, U, V)
    if limited_memory:
        K, k, P, p = tvlqr(Q, q, R, r, M, A, B, c[1:])
        dX, dU = rollout(K, k, c[0], A, B, c[1:])
    else:
        K, k, P, p = tvlqr_gpu(Q, q, R, r, M, A, B, c[1:])
        dX, dU = rollout_gpu(K, k, c[0], A, B, c[1:])
    dV = dual_lqr(dX, P, p)
    if detail is not None:
        detail.update(Q=Q, R=R, M=M, q=q, r=r, A=A, B=B, K=K, k=k, P=P, p=p)
    return dX, dU, dV, q, r


def parallel_filter_line_search(evaluator, X_in, U_in, V_in, dX, dU, dV, current_cost, current_c, q, r,
                                theta_max=1e-2, theta_min=1e-6, eta=1e-4, gamma_phi=1e-6, gamma_theta=1e-6,
                                detail=None):
    """optimizers.py:468-578: 11 step sizes 2^0..2^-10, first accepted else the last."""
    theta_k = torch.sum(current_c * current_c)
    phi_k = current_cost
    slope = torch.sum(q * dX) + torch.sum(r * dU)
    accepted, trials, thetas, phis = [], [], [], []
    for i in range(11):
        alpha = 2.0 ** (-i)
        X_new, U_new, V_new = X_in + alpha * dX, U_in + alpha * dU, V_in + alpha * dV
        new_cost, new_c = evaluator(X_new, U_new)
        theta_new = torch.sum(new_c * new_c)
        phi_new = new_cost
        condition1 = theta_new > theta_max
        case1 = condition1 & (theta_new < (1 - gamma_theta) * theta_k)
        condition2 = (torch.maximum(theta_new, theta_k) < theta_min) & (slope < 0)
        case2 = condition2 & (phi_new < phi_k + eta * alpha * slope)
        condition3 = ~(condition1 | condition2)
        case3 = condition3 & ((phi_new < phi_k - gamma_phi * phi_k) | (theta_new < (1 - gamma_theta) * theta_k))
        accepted.append(bool(case1 | case2 | case3))
        trials.append((X_new, U_new, V_new))
        thetas.append(theta_new); phis.append(phi_new)
    best = accepted.index(True) if any(accepted) else 10  # index -1 of 11 entries
    if detail is not None:
        detail.update(alpha_index=best, theta=torch.stack(thetas), phi=torch.stack(phis),
                      theta_k=theta_k, phi_k=phi_k, slope=slope, accepted=accepted)
    return trials[best]


def merit_rho(c, dV):
    """optimizers.py:185-198."""
    c2 = torch.sum(c * c)
    dV2 = torch.sum(dV * dV)
    return 2.0 * torch.sqrt(dV2 / c2) if bool(c2 > 1e-12) else torch.as_tensor(1e-2, dtype=c.dtype)


def merit_slope(dX, dU, dV, c, q, r, rho):
    """optimizers.py:201-217 (`slope`; note the factor 2 on sum(dV*c), unlike old_optimazers.py:170)."""
    return torch.sum(q * dX) + torch.sum(r * dU) + 2 * torch.sum(dV * c) - rho * torch.sum(c * c)


def parallel_line_search(evaluator, X_in, U_in, V_in, dX, dU, dV, g, c, q, r, armijo_factor=1e-4, detail=None):
    """optimizers.py:295-354 with the merit function / rho / slope of the (commented-out) call site
    optimizers.py:637-669: merit = g + sum((V + 0.5 rho c) c), 11 step sizes 2^0..2^-10, Armijo test
    `merit_new > merit + alpha * armijo * slope` (True = rejected), best = argmin(rejected) if any rejected else 0;
    a NaN merit is replaced by the current merit."""
    rho = merit_rho(c, dV)
    merit_fn = lambda V, g_, c_: g_ + torch.sum((V + 0.5 * rho * c_) * c_)
    current = merit_fn(V_in, g, c)
    slope = merit_slope(dX, dU, dV, c, q, r, rho)
    rejected, trials, merits, thetas, phis, vcs = [], [], [], [], [], []
    for i in range(11):
        alpha = 2.0 ** (-i)
        X_new, U_new, V_new = X_in + alpha * dX, U_in + alpha * dU, V_in + alpha * dV
        new_g, new_c = evaluator(X_new, U_new)
        m_new = merit_fn(V_new, new_g, new_c)
        if bool(torch.isnan(m_new)):
            m_new = current
        rejected.append(bool(m_new > current + alpha * armijo_factor * slope))
        trials.append((X_new, U_new, V_new))
        merits.append(m_new); thetas.append(torch.sum(new_c * new_c)); phis.append(new_g)
        vcs.append(torch.sum(V_new * new_c))
    best = rejected.index(False) if (any(rejected) and not all(rejected)) else 0
    if detail is not None:
        detail.update(alpha_index=best, merit=torch.stack(merits), theta=torch.stack(thetas), phi=torch.stack(phis),
                      vc=torch.stack(vcs), rho=rho, merit_k=current, merit_slope=slope, rejected=rejected)
    return trials[best]


def mpc(cost, dynamics, hessian_approx, limited_memory, reference, parameter, W, x0, X_in, U_in, V_in, detail=None,
        line_search="filter"):
    """optimizers.py:603-683: one real-time SQP iteration.

    ``line_search="merit"`` selects the augmented-Lagrangian merit line search that the reference keeps (commented
    out) beside the filter search: optimizers.py:637-669.

    ``cost(W, reference, x, u, t)``, ``hessian_approx(W, reference, x, u, t)``,
    ``dynamics(x, u, t, parameter)`` as the reference's partials leave them.
    """
    _cost = lambda x, u, t: cost(W, reference, x, u, t)
    _hess = None if hessian_approx is None else (lambda x, u, t: hessian_approx(W, reference, x, u, t))
    _dyn = lambda x, u, t: dynamics(x, u, t, parameter)
    evaluator = lambda X, U: model_evaluator(_cost, _dyn, x0, X, U)
    g, c = evaluator(X_in, U_in)
    dX, dU, dV, q, r = compute_search_direction(_cost, _dyn, _hess, limited_memory, x0, X_in, U_in, V_in, c,
                                                detail=detail)
    if detail is not None:
        detail.update(g=g, c=c, dX=dX, dU=dU, dV=dV)
    if line_search == "merit":
        return parallel_line_search(evaluator, X_in, U_in, V_in, dX, dU, dV, g, c, q, r, detail=detail)
    return parallel_filter_line_search(evaluator, X_in, U_in, V_in, dX, dU, dV, g, c, q, r, detail=detail)
