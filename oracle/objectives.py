"""Restatement of reference ``mpx/utils/objectives.py`` (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import torch
from torch.func import grad, jacrev, vmap

from . import mjx_math as math


def penalty(constraint, alpha=0.1, sigma=5):
    """reference objectives.py:7-13 (relaxed log barrier, clipped to [0, 1e8])."""
    def safe_log(x):
        return torch.log(torch.clamp(x, 1e-10, 1e6))
    quadratic_barrier = alpha / 2 * (torch.square((constraint - 2 * sigma) / sigma) - torch.ones_like(constraint))
    log_barrier = -alpha * safe_log(constraint)
    return torch.clamp(torch.where(constraint > sigma, log_barrier, quadratic_barrier + log_barrier), 0, 1e8)


def _hessian_penalty(alpha=0.1, sigma=5):
    f = lambda c: penalty(c, alpha, sigma)
    return grad(grad(f))


def _qform(a, W, b=None):
    b = a if b is None else b
    return a @ W @ b


def _cone(grf, mu, eps, nc):
    return mu * grf[2::3] - torch.sqrt(torch.square(grf[1::3]) + torch.square(grf[::3])
                                       + torch.ones(nc, dtype=grf.dtype) * eps)


# ------------------------------------------------------------------ SRBD
def quadruped_srbd_obj(n_contact, N, W, reference, x, u, t):
    """reference objectives.py:15-37."""
    p, quat, dp, omega, grf = x[:3], x[3:7], x[7:10], x[10:13], u
    p_ref, quat_ref = reference[t, :3], reference[t, 3:7]
    dp_ref, omega_ref = reference[t, 7:10], reference[t, 10:13]
    friction_cone = _cone(grf, 0.7, 1e-2, n_contact)
    friction_cone = torch.clamp(penalty(friction_cone), 1e-6, 1e6)
    qs = math.quat_sub(quat, quat_ref)
    term_cost = (_qform(p - p_ref, W[:3, :3]) + _qform(qs, W[3:6, 3:6])
                 + _qform(dp - dp_ref, W[6:9, 6:9]) + _qform(omega - omega_ref, W[9:12, 9:12]))
    stage_cost = term_cost + _qform(grf, W[12:12 + 3 * n_contact, 12:12 + 3 * n_contact]) + torch.sum(friction_cone)
    return 0.5 * term_cost if t == N else 0.5 * stage_cost


def quadruped_srbd_hessian_gn(n_contact, W, reference, x, u, t):
    """reference objectives.py:39-75."""
    def residual(x, u):
        p, quat, dp, omega, grf = x[:3], x[3:7], x[7:10], x[10:13], u
        return torch.cat([p - reference[t, :3], math.quat_sub(quat, reference[t, 3:7]),
                          dp - reference[t, 7:10], omega - reference[t, 10:13], grf])

    def friction_constraint(u):
        return _cone(u, 0.7, 1e-2, n_contact)
    Jx = jacrev(residual, 0)(x, u)
    Ju = jacrev(residual, 1)(x, u)
    contact = reference[t, 13:13 + n_contact]
    hp = _hessian_penalty()
    Jf = jacrev(friction_constraint)(u)
    H_penalty = torch.diag(torch.clamp(vmap(hp)(friction_constraint(u)), 1e-6, 1e6) * contact)
    H_constraint = Jf.T @ H_penalty @ Jf
    return Jx.T @ W @ Jx, Ju.T @ W @ Ju + H_constraint, Jx.T @ W @ Ju


# ------------------------------------------------------------------ whole-body common pieces
def _wb_split(x, u, nj, nc, grf_in_u=False):
    p, quat, q = x[:3], x[3:7], x[7:7 + nj]
    dp, omega, dq = x[7 + nj:10 + nj], x[10 + nj:13 + nj], x[13 + nj:13 + 2 * nj]
    p_leg = x[13 + 2 * nj:13 + 2 * nj + 3 * nc]
    grf = u[nj:] if grf_in_u else x[13 + 2 * nj + 3 * nc:]
    tau = u[:nj]
    return p, quat, q, dp, omega, dq, p_leg, grf, tau


def _wb_ref(reference, t, nj, nc):
    r = reference[t]
    return (r[:3], r[3:7], r[7:7 + nj], r[7 + nj:10 + nj], r[10 + nj:13 + nj],
            r[13 + nj:13 + nj + 3 * nc], r[13 + nj + 3 * nc:13 + nj + 4 * nc],
            r[13 + nj + 4 * nc:13 + nj + 7 * nc])


def _wb_quadratic(W, nj, nc, p, quat, q, dp, omega, dq, p_leg, grf, tau,
                  p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, grf_ref, contact_map):
    qs = math.quat_sub(quat, quat_ref)
    a = 12 + 2 * nj
    term = (_qform(p - p_ref, W[:3, :3]) + _qform(qs, W[3:6, 3:6]) + _qform(q - q_ref, W[6:6 + nj, 6:6 + nj])
            + _qform(dp - dp_ref, W[6 + nj:9 + nj, 6 + nj:9 + nj])
            + _qform(omega - omega_ref, W[9 + nj:12 + nj, 9 + nj:12 + nj])
            + _qform(dq, W[12 + nj:12 + 2 * nj, 12 + nj:12 + 2 * nj]))
    leg = contact_map * (p_leg - p_leg_ref)
    stage = (term + _qform(leg, W[a:a + 3 * nc, a:a + 3 * nc])
             + _qform(tau, W[a + 3 * nc:a + nj + 3 * nc, a + 3 * nc:a + nj + 3 * nc])
             + _qform(grf - grf_ref, W[a + nj + 3 * nc:a + nj + 6 * nc, a + nj + 3 * nc:a + nj + 6 * nc]))
    return term, stage


def _wb_residual(reference, t, nj, nc, contact_map_fn, grf_in_u):
    def residual(x, u):
        p, quat, q, dp, omega, dq, p_leg, grf, tau = _wb_split(x, u, nj, nc, grf_in_u)
        p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, contact, grf_ref = _wb_ref(reference, t, nj, nc)
        cm = contact_map_fn(contact, x.dtype)
        return torch.cat([p - p_ref, math.quat_sub(quat, quat_ref), q - q_ref, dp - dp_ref, omega - omega_ref,
                          dq, cm * (p_leg - p_leg_ref), tau, grf - grf_ref])
    return residual


def _limits(vec, lim):
    """kron(eye(n), [-1, 1]).T @ vec + lim + 1e-2 -> interleaved (-v_i + lim, +v_i + lim)."""
    n = vec.shape[0]
    K = torch.kron(torch.eye(n, dtype=vec.dtype), torch.tensor([[-1.0, 1.0]], dtype=vec.dtype)).T
    return K @ vec + lim + torch.ones_like(lim) * 1e-2


# ------------------------------------------------------------------ quadruped whole-body
def quadruped_wb_obj(swing_tracking, n_joints, n_contact, N, W, reference, x, u, t):
    """reference objectives.py:77-125."""
    nj, nc = n_joints, n_contact
    p, quat, q, dp, omega, dq, p_leg, grf, tau = _wb_split(x, u, nj, nc)
    p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, contact, grf_ref = _wb_ref(reference, t, nj, nc)
    friction_cone = penalty(_cone(grf, 0.5, 1e-2, nc))
    torque_limits = _limits(tau, torch.full((24,), 44.0, dtype=x.dtype))
    joint_speed_limits = _limits(dq, torch.ones(2 * nj, dtype=x.dtype) * 10)
    if swing_tracking:
        contact_map = torch.ones(3 * nc, dtype=x.dtype)
    else:
        contact_map = contact.repeat_interleave(3)
    term, stage = _wb_quadratic(W, nj, nc, p, quat, q, dp, omega, dq, p_leg, grf, tau,
                                p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, grf_ref, contact_map)
    stage = (stage + torch.sum(penalty(torque_limits, 1, 1)) + torch.sum(friction_cone * contact)
             + torch.sum(penalty(joint_speed_limits, 1, 1)))
    return 0.5 * term if t == N else 0.5 * stage


def quadruped_wb_hessian_gn(swing_tracking, n_joints, n_contact, W, reference, x, u, t):
    """reference objectives.py:127-202 (cone eps 1e-1 here vs 1e-2 in the cost; speed penalty with default alpha/sigma)."""
    nj, nc = n_joints, n_contact
    contact = reference[t, 13 + nj + 3 * nc:13 + nj + 4 * nc]
    if swing_tracking:
        cm_fn = lambda c, dt: torch.ones(3 * nc, dtype=dt)
    else:
        cm_fn = lambda c, dt: c.repeat_interleave(3)
    residual = _wb_residual(reference, t, nj, nc, cm_fn, False)

    def friction_constraint(x):
        return _cone(x[13 + 2 * nj + 3 * nc:], 0.5, 1e-1, nc)

    def torque_constraint(u):
        return _limits(u[:nj], torch.full((24,), 44.0, dtype=u.dtype))

    def speed_constraint(x):
        return _limits(x[13 + nj:13 + 2 * nj], torch.ones(2 * nj, dtype=x.dtype) * 10)
    Jx = jacrev(residual, 0)(x, u)
    Ju = jacrev(residual, 1)(x, u)
    hp = _hessian_penalty()
    hp_torque = _hessian_penalty(1, 1)
    Jf = jacrev(friction_constraint)(x)
    Jt = jacrev(torque_constraint)(u)
    Js = jacrev(speed_constraint)(x)
    H_penalty = torch.diag(torch.clamp(vmap(hp)(friction_constraint(x)), -1e6, 1e6) * contact)
    H_penalty_torque = torch.diag(torch.clamp(vmap(hp_torque)(torque_constraint(u)), -1e6, 1e6))
    H_penalty_speed = torch.diag(torch.clamp(vmap(hp)(speed_constraint(x)), -1e6, 1e6))
    H_constraint = Jf.T @ H_penalty @ Jf
    H_constraint_u = Jt.T @ H_penalty_torque @ Jt
    H_constraint = H_constraint + Js.T @ H_penalty_speed @ Js
    return Jx.T @ W @ Jx + H_constraint, Ju.T @ W @ Ju + H_constraint_u, Jx.T @ W @ Ju


# ------------------------------------------------------------------ H1 / Talos
def _humanoid_obj(grf_in_u, n_joints, n_contact, N, W, reference, x, u, t):
    nj, nc = n_joints, n_contact
    p, quat, q, dp, omega, dq, p_leg, grf, tau = _wb_split(x, u, nj, nc, grf_in_u)
    p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, contact, grf_ref = _wb_ref(reference, t, nj, nc)
    contact_map = torch.ones(3 * nc, dtype=x.dtype)
    term, stage = _wb_quadratic(W, nj, nc, p, quat, q, dp, omega, dq, p_leg, grf, tau,
                                p_ref, quat_ref, q_ref, dp_ref, omega_ref, p_leg_ref, grf_ref, contact_map)
    return 0.5 * term if t == N else 0.5 * stage


def _humanoid_hessian(grf_in_u, n_joints, n_contact, W, reference, x, u, t):
    nj, nc = n_joints, n_contact
    residual = _wb_residual(reference, t, nj, nc, lambda c, dt: torch.ones(3 * nc, dtype=dt), grf_in_u)
    Jx = jacrev(residual, 0)(x, u)
    Ju = jacrev(residual, 1)(x, u)
    # the friction-cone Hessian is computed by the reference but never added (objectives.py:316-323, 443-450)
    return Jx.T @ W @ Jx, Ju.T @ W @ Ju, Jx.T @ W @ Ju


def h1_wb_obj(n_joints, n_contact, N, W, reference, x, u, t):
    """reference objectives.py:204-250."""
    return _humanoid_obj(False, n_joints, n_contact, N, W, reference, x, u, t)


def h1_wb_hessian_gn(n_joints, n_contact, W, reference, x, u, t):
    """reference objectives.py:252-323."""
    return _humanoid_hessian(False, n_joints, n_contact, W, reference, x, u, t)


def talos_wb_obj(n_joints, n_contact, N, W, reference, x, u, t):
    """reference objectives.py:325-376 (grf = u[nj:])."""
    return _humanoid_obj(True, n_joints, n_contact, N, W, reference, x, u, t)


def talos_wb_hessian_gn(n_joints, n_contact, W, reference, x, u, t):
    """reference objectives.py:378-450."""
    return _humanoid_hessian(True, n_joints, n_contact, W, reference, x, u, t)
