"""Restatement of the reference's ORIGINAL multi-iteration solver ``primal_dual_ilqr`` (TEST INFRASTRUCTURE).

Follows mpx/primal_dual_ilqr/primal_dual_ilqr/old_optimazers.py: ``lagrangian`` 21-30, ``regularize`` 33-63,
``compute_search_direction`` 66-135 (exact Lagrangian Hessians by ``quadratize``, eigenvalue projection
``project_psd_cone`` linalg_helpers.py:8-23), ``merit_rho`` 138-151, ``slope`` 154-170 (``sum(dV*c)``, not the
``2 sum(dV*c)`` of the newer optimizers.py:217), sequential Armijo ``line_search`` 173-251, the device-side
``while_loop`` with ``continuation_criterion`` 435-447, return tuple 473.  This is the only entry point of the reference
with an iteration count and the one its own tests (tests/simple_tests.py) exercise.

``cost(x, u, t)`` / ``dynamics(x, u, t)`` take a python int ``t``; trajax's ``linearize`` / ``quadratize`` are python
loops over knots with ``torch.func`` (see oracle/primal_dual_ilqr.py).
"""
import torch
from torch.func import hessian, jacrev

from . import primal_dual_ilqr as pdi


def project_psd_cone(Q, delta=0.0):
    """linalg_helpers.py:8-23."""
    S, V = torch.linalg.eigh(Q)
    S = torch.clamp(S, min=delta)
    Qp = V @ torch.diag(S) @ V.T
    return 0.5 * (Qp + Qp.T)


def regularize(Q, R, M, make_psd, psd_delta):
    """old_optimazers.py:33-63."""
    T = R.shape[0]
    if make_psd:
        R = torch.stack([project_psd_cone(R[t], psd_delta) for t in range(T)])
    Qn = []
    for t in range(T):
        Rinv = pdi.solve_spd(R[t], torch.eye(R.shape[1], dtype=R.dtype))
        MRM = M[t] @ Rinv @ M[t].T
        S = Q[t] - MRM
        if make_psd:
            S = project_psd_cone(S, psd_delta)
        Qn.append(S + MRM)
    QT = project_psd_cone(Q[T], psd_delta) if make_psd else Q[T]
    return torch.stack(Qn + [QT]), R


def compute_search_direction(cost, dynamics, x0, X, U, V, c, make_psd, psd_delta):
    """old_optimazers.py:66-135: exact Hessians of the Lagrangian, regularised, sequential Riccati."""
    T = U.shape[0]
    zero_u = torch.zeros_like(U[0])
    zero_v = torch.zeros_like(V[0])
    Qs, Rs, Ms, qs, rs, As, Bs = [], [], [], [], [], [], []
    for t in range(T + 1):
        x, u = X[t], (U[t] if t < T else zero_u)
        v = V[t + 1] if t < T else zero_v
        v_prev = V[t]

        def lag(x_, u_):
            c3 = torch.dot(v_prev, (x0 - x_) if t == 0 else -x_)
            return cost(x_, u_, t) + torch.dot(v, dynamics(x_, u_, t)) + c3
        Qs.append(hessian(lag, 0)(x, u))
        qs.append(jacrev(lag, 0)(x, u))
        if t < T:
            Rs.append(hessian(lag, 1)(x, u))
            Ms.append(jacrev(jacrev(lag, 0), 1)(x, u))
            rs.append(jacrev(lag, 1)(x, u))
            As.append(jacrev(lambda x_: dynamics(x_, u, t))(x))
            Bs.append(jacrev(lambda u_: dynamics(x, u_, t))(u))
    st = torch.stack
    Q, R, M, q, r, A, B = st(Qs), st(Rs), st(Ms), st(qs), st(rs), st(As), st(Bs)
    Q, R = regularize(Q, R, M, make_psd, psd_delta)
    K, k, P, p = pdi.tvlqr(Q, q, R, r, M, A, B, c[1:])
    dX, dU = pdi.rollout(K, k, c[0], A, B, c[1:])
    dV = pdi.dual_lqr(dX, P, p)
    return dX, dU, dV, q, r


def slope(dX, dU, dV, c, q, r, rho):
    """old_optimazers.py:154-170."""
    return torch.sum(q * dX) + torch.sum(r * dU) + torch.sum(dV * c) - rho * torch.sum(c * c)


def line_search(merit_function, evaluator, X_in, U_in, V_in, dX, dU, dV, current_merit, current_g, current_c,
                merit_slope, armijo_factor, alpha_0, alpha_mult, alpha_min):
    """old_optimazers.py:173-251: sequential backtracking; returns (..., no_errors, alpha, trials)."""
    X, U, V, new_g, new_c = X_in, U_in, V_in, current_g, current_c
    new_merit = torch.as_tensor(float("inf"), dtype=X_in.dtype)
    alpha = alpha_0 / alpha_mult
    trials = 0
    while bool(new_merit > current_merit + alpha * armijo_factor * merit_slope) and alpha > alpha_min:
        alpha = alpha * alpha_mult
        X, U, V = X_in + alpha * dX, U_in + alpha * dU, V_in + alpha * dV
        new_g, new_c = evaluator(X, U)
        new_merit = merit_function(V, new_g, new_c)
        if bool(torch.isnan(new_merit)):
            new_merit = current_merit
        trials += 1
    return X, U, V, new_g, new_c, alpha > alpha_min, alpha, trials


def continuation_criterion(dX, dU, dV, iteration, no_errors, c, merit_slope, max_iterations, slope_threshold,
                           var_threshold, c_sq_threshold):
    """old_optimazers.py:435-447."""
    c_sq = torch.sum(c * c)
    slope_ok = bool(torch.abs(merit_slope) > slope_threshold)
    delta_ok = bool(torch.sum(dX * dX) + torch.sum(dU * dU) + torch.sum(dV * dV) > var_threshold ** 2)
    progress_ok = (slope_ok and delta_ok) or bool(c_sq > c_sq_threshold)
    return bool(no_errors) and iteration < max_iterations and progress_ok


def primal_dual_ilqr(cost, dynamics, x0, X_in, U_in, V_in, max_iterations=100, slope_threshold=1e-4, var_threshold=0.0,
                     c_sq_threshold=1e-4, make_psd=True, psd_delta=1e-6, armijo_factor=1e-4, alpha_0=1.0,
                     alpha_mult=0.5, alpha_min=5e-5, direction=None, trace=None):
    """old_optimazers.py:280-474.  Returns (X, U, V, iterations, g, c, no_errors).

    ``direction(x0, X, U, V, c) -> dX, dU, dV, q, r`` overrides the search direction (default: the exact-Hessian one of
    old_optimazers.py:66-135); the MPC path's Gauss-Newton direction (optimizers.py:121-184) can be plugged in to run
    the legacy loop / line search / convergence test on the robot models."""
    evaluator = lambda X, U: pdi.model_evaluator(cost, dynamics, x0, X, U)
    if direction is None:
        direction = lambda x0_, X, U, V, c: compute_search_direction(cost, dynamics, x0_, X, U, V, c, make_psd, psd_delta)

    def direction_and_merit(X, U, V, g, c):
        dX, dU, dV, q, r = direction(x0, X, U, V, c)
        rho = pdi.merit_rho(c, dV)
        merit = g + torch.sum((V + 0.5 * rho * c) * c)
        return dX, dU, dV, rho, merit, slope(dX, dU, dV, c, q, r, rho)

    X, U, V = X_in, U_in, V_in
    g, c = evaluator(X, U)
    dX, dU, dV, rho, merit, merit_slope = direction_and_merit(X, U, V, g, c)
    iteration, no_errors = 0, True
    while continuation_criterion(dX, dU, dV, iteration, no_errors, c, merit_slope, max_iterations, slope_threshold,
                                 var_threshold, c_sq_threshold):
        merit_fn = lambda V_, g_, c_, rho_=rho: g_ + torch.sum((V_ + 0.5 * rho_ * c_) * c_)
        X, U, V, g, c, no_errors, alpha, trials = line_search(merit_fn, evaluator, X, U, V, dX, dU, dV, merit, g, c,
                                                              merit_slope, armijo_factor, alpha_0, alpha_mult, alpha_min)
        dX, dU, dV, rho, merit, merit_slope = direction_and_merit(X, U, V, g, c)
        iteration += 1
        if trace is not None:
            trace.append(dict(iteration=iteration, alpha=alpha, trials=trials, g=float(g), c_sq=float(torch.sum(c * c)),
                              merit_slope=float(merit_slope), no_errors=bool(no_errors)))
    no_errors = bool(no_errors) and iteration < max_iterations
    return X, U, V, iteration, g, c, no_errors


def legacy_loop_on_problem(pr, reference, parameter, x0, X_in, U_in, V_in, max_iterations, alpha_min=2.0 ** -10, trace=None):
    """The legacy loop -- sequential Armijo ``line_search``, ``continuation_criterion``, legacy merit slope
    (old_optimazers.py:154-251, 387-473) -- around the MPC path's Gauss-Newton search direction (optimizers.py:121-184)
    for one environment of an ``oracle.problems.Problem``.  This is what ``MPXB_LS_ARMIJO`` + ``MPXB_STOP_LEGACY`` run
    on the device; ``alpha_min = 2^-10`` puts the backtracking on the library's 11 step sizes."""
    t = lambda a: torch.as_tensor(a, dtype=pr.dtype)
    reference, parameter, x0, X_in, U_in, V_in = map(t, (reference, parameter, x0, X_in, U_in, V_in))
    cost = lambda x, u, k: pr.cost(pr.W, reference, x, u, k)
    hess = lambda x, u, k: pr.hessian_approx(pr.W, reference, x, u, k)
    dyn = lambda x, u, k: pr.dynamics(x, u, k, parameter)
    direction = lambda x0_, X, U, V, c: pdi.compute_search_direction(cost, dyn, hess, True, x0_, X, U, V, c)
    return primal_dual_ilqr(cost, dyn, x0, X_in, U_in, V_in, max_iterations=max_iterations, alpha_min=alpha_min,
                            direction=direction, trace=trace)
