"""Vectorised evaluation of the oracle's SQP step (TEST INFRASTRUCTURE, see oracle/__init__.py).

Same arithmetic as ``oracle/primal_dual_ilqr.mpc`` -- it calls the very same restated cost / dynamics / Hessian
functions (oracle/models.py, oracle/objectives.py) -- but the per-knot Python loops of the trajax combinators
(``vectorize(fun)`` = ``jax.vmap`` over knots, reference optimizers.py:158-173, 596-600) are evaluated with
``torch.func.vmap`` over (environment, knot) instead of one knot at a time, exactly as the reference evaluates
them (``jax.vmap`` over knots inside ``jax.vmap`` over environments, examples/multi_env.py:46-49).  This is what makes
(i) golden fixtures with several environments per robot affordable and (ii) ``bench.py``'s ``cpu_baseline`` a sane
stand-in for "the reference on the CPU backend" (all host cores, batched over the same environments).

``tests/test_oracle_batched.py`` checks it against the loop form to float64 rounding.

The functions take a python int knot index ``t`` and index ``reference[t]`` / ``parameter[t]``; under vmap each knot
gets its own row, passed as a one-row table with ``t = 0`` (stage knots; N > 0, so ``t == N`` is false) or as an
(N+1)-row table with ``t = N`` (terminal knot).
"""
import torch
from torch.func import jacrev, vmap

from . import primal_dual_ilqr as pdi


class BatchedProblem:
    """Batched (env x knot) evaluation of a ``oracle.problems.Problem``."""

    def __init__(self, problem):
        self.pr = problem
        self.N = problem.config.N
        self.dtype = problem.dtype
        cost, dyn, hess = problem.cost, problem.dynamics, problem.hessian_approx
        N = self.N
        # per-knot closures: (W, ref_row, par_row, x, u)
        self._stage_cost = lambda W, ref, x, u: cost(W, ref[None], x, u, 0)
        self._term_cost = lambda W, ref, x, u: cost(W, ref[None].expand(N + 1, -1), x, u, N)
        self._dyn = lambda par, x, u: dyn(x, u, 0, par[None])
        self._hess = lambda W, ref, x, u: hess(W, ref[None], x, u, 0)

    # ------------------------------------------------------------------ evaluation (optimizers.py:580-602)
    def evaluate(self, W, reference, parameter, x0, X, U):
        """g [B], c [B, N+1, n] for a batch of trajectories (leading axis B; W [nw,nw] shared)."""
        B, N = X.shape[0], self.N
        zero_u = torch.zeros_like(U[:, :1])
        Upad = torch.cat([U, zero_u], dim=1)
        stage = vmap(vmap(self._stage_cost, in_dims=(None, 0, 0, 0)), in_dims=(None, 0, 0, 0))(
            W, reference[:, :N], X[:, :N], Upad[:, :N])
        term = vmap(self._term_cost, in_dims=(None, 0, 0, 0))(W, reference[:, N], X[:, N], Upad[:, N])
        g = stage.sum(dim=1) + term
        f = vmap(vmap(self._dyn))(parameter[:, :N], X[:, :N], U)
        c = torch.cat([(x0 - X[:, 0])[:, None], f - X[:, 1:]], dim=1)
        return g, c, torch.cat([stage, term[:, None]], dim=1)

    # ------------------------------------------------------------------ derivatives (optimizers.py:150-173)
    def linearize_quadratize(self, W, reference, parameter, x0, X, U, V):
        B, N = X.shape[0], self.N
        zero_u = torch.zeros_like(U[:, :1])
        Upad = torch.cat([U, zero_u], dim=1)
        vv = lambda f, dims: vmap(vmap(f, in_dims=dims), in_dims=dims)
        Q, R, M = vv(self._hess, (None, 0, 0, 0))(W, reference, X, Upad)
        gx_s, gu_s = vv(jacrev(self._stage_cost, argnums=(2, 3)), (None, 0, 0, 0))(W, reference[:, :N], X[:, :N], U)
        gx_t = vmap(jacrev(self._term_cost, argnums=2), in_dims=(None, 0, 0, 0))(W, reference[:, N], X[:, N], Upad[:, N])
        A, Bm = vv(jacrev(self._dyn, argnums=(1, 2)), (0, 0, 0))(parameter[:, :N], X[:, :N], U)
        # lagrangian (optimizers.py:111-120): q_t = grad_x cost_t + A_t' V_{t+1} - V_t, r_t = grad_u cost_t + B_t' V_{t+1};
        # V_{N+1} := 0 (optimizers.py:168 pads V[1:])
        q = torch.cat([gx_s + torch.einsum("btij,bti->btj", A, V[:, 1:]), gx_t[:, None]], dim=1) - V
        r = gu_s + torch.einsum("btij,bti->btj", Bm, V[:, 1:])
        return Q, R[:, :N], M[:, :N], q, r, A, Bm

    # ------------------------------------------------------------------ one mpc call per environment
    def mpc(self, limited_memory, reference, parameter, x0, X, U, V, W=None, detail=None, lq_order="sequential",
            max_alpha=11):
        """optimizers.py:603-683 for E environments at once.  Returns X, U, V [E, ...]; ``detail`` (dict) receives
        every intermediate stacked over environments, like ``oracle.primal_dual_ilqr.mpc`` fills it per environment."""
        t = lambda a: torch.as_tensor(a, dtype=self.dtype)
        W = self.pr.W if W is None else t(W)
        reference, parameter, x0, X, U, V = map(t, (reference, parameter, x0, X, U, V))
        E, N = X.shape[0], self.N
        g, c, _ = self.evaluate(W, reference, parameter, x0, X, U)
        Q, R, M, q, r, A, Bm = self.linearize_quadratize(W, reference, parameter, x0, X, U, V)
        Ks, ks, Ps, ps, dXs, dUs, dVs = [], [], [], [], [], [], []
        for e in range(E):
            if limited_memory:
                K, k, P, p = pdi.tvlqr(Q[e], q[e], R[e], r[e], M[e], A[e], Bm[e], c[e, 1:])
                dX, dU = pdi.rollout(K, k, c[e, 0], A[e], Bm[e], c[e, 1:])
            else:
                K, k, P, p = pdi.tvlqr_gpu(Q[e], q[e], R[e], r[e], M[e], A[e], Bm[e], c[e, 1:], order=lq_order)
                dX, dU = pdi.rollout_gpu(K, k, c[e, 0], A[e], Bm[e], c[e, 1:])
            dV = pdi.dual_lqr(dX, P, p)
            for lst, v in zip((Ks, ks, Ps, ps, dXs, dUs, dVs), (K, k, P, p, dX, dU, dV)):
                lst.append(v)
        st = torch.stack
        K, k, P, p, dX, dU, dV = map(st, (Ks, ks, Ps, ps, dXs, dUs, dVs))
        # parallel_filter_line_search (optimizers.py:468-578): all 11 trial points of every environment in one batch
        alphas = torch.tensor([2.0 ** (-i) for i in range(11)], dtype=self.dtype)
        tr = lambda a, d: (a[:, None] + alphas.view(1, -1, *([1] * (a.dim() - 1))) * d[:, None])
        Xt, Ut, Vt = tr(X, dX), tr(U, dU), tr(V, dV)
        rep = lambda a: a[:, None].expand(E, 11, *a.shape[1:]).reshape(E * 11, *a.shape[1:])
        flat = lambda a: a.reshape(E * 11, *a.shape[2:])
        gt, ct, _ = self.evaluate(W, rep(reference), rep(parameter), rep(x0), flat(Xt), flat(Ut))
        phi = gt.reshape(E, 11)
        theta = (ct * ct).sum(dim=(1, 2)).reshape(E, 11)
        theta_k = (c * c).sum(dim=(1, 2))
        phi_k = g
        slope = (q * dX).sum(dim=(1, 2)) + (r * dU).sum(dim=(1, 2))
        theta_max, theta_min, eta, gamma_phi, gamma_theta = 1e-2, 1e-6, 1e-4, 1e-6, 1e-6
        cond1 = theta > theta_max
        case1 = cond1 & (theta < (1 - gamma_theta) * theta_k[:, None])
        cond2 = (torch.maximum(theta, theta_k[:, None]) < theta_min) & (slope[:, None] < 0)
        case2 = cond2 & (phi < phi_k[:, None] + eta * alphas[None] * slope[:, None])
        cond3 = ~(cond1 | cond2)
        case3 = cond3 & ((phi < phi_k[:, None] - gamma_phi * phi_k[:, None]) | (theta < (1 - gamma_theta) * theta_k[:, None]))
        accepted = case1 | case2 | case3
        best = torch.where(accepted.any(dim=1), accepted.to(torch.int64).argmax(dim=1), torch.full((E,), 10))
        idx = torch.arange(E)
        Xn, Un, Vn = Xt[idx, best], Ut[idx, best], Vt[idx, best]
        if detail is not None:
            detail.update(Q=Q, R=R, M=M, q=q, r=r, A=A, B=Bm, K=K, k=k, P=P, p=p, g=g, c=c, dX=dX, dU=dU, dV=dV,
                          alpha_index=best, theta=theta, phi=phi, theta_k=theta_k, phi_k=phi_k, slope=slope,
                          accepted=accepted)
        return Xn, Un, Vn
