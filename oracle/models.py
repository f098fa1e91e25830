"""Restatement of reference ``mpx/utils/models.py`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

``parameter`` is the full [N+1, npar] array and ``t`` a python int, exactly as the reference's
``dynamics(x, u, t, parameter)`` signature.
"""
import torch

from . import mjx_math as math
from . import mjx_smooth as mjx


def quadruped_srbd_dynamics(mass, inertia, inertia_inv, dt, x, u, t, parameter):
    """reference mpx/utils/models.py:6-37."""
    p = x[:3]
    quat = x[3:7]
    dp = x[7:10]
    omega = x[10:13]
    grf = u
    contact = parameter[t, :4]
    p_legs = parameter[t, 4:]
    g = torch.tensor([0.0, 0.0, -9.81], dtype=x.dtype)
    dp_next = dp + (g + (1 / mass) * (grf[:3] * contact[0] + grf[3:6] * contact[1]
                                      + grf[6:9] * contact[2] + grf[9:12] * contact[3])) * dt
    p0, p1, p2, p3 = p_legs[:3], p_legs[3:6], p_legs[6:9], p_legs[9:]
    b_R_w = math.quat_to_mat(quat)
    wrench = (math.cross(p0 - p, grf[:3]) * contact[0] + math.cross(p1 - p, grf[3:6]) * contact[1]
              + math.cross(p2 - p, grf[6:9]) * contact[2] + math.cross(p3 - p, grf[9:12]) * contact[3])
    omega_next = omega + inertia_inv @ (b_R_w.T @ wrench - math.cross(omega, inertia @ omega)) * dt
    p_new = p + dp_next * dt
    quat_new = math.quat_integrate(quat, omega_next, dt)
    return torch.cat([p_new, quat_new, dp_next, omega_next])


def _wb_common(m, contact_id, body_id_per_foot, n_joints, x):
    qpos = x[:n_joints + 7]
    qvel = x[n_joints + 7:2 * n_joints + 13]
    d = mjx.fwd_position(m, qpos)
    d = mjx.fwd_velocity(m, d, qvel)
    feet = [d.geom_xpos[g] for g in contact_id]
    Js = [mjx.jac(m, d, feet[i], body_id_per_foot[i])[0] for i in range(len(feet))]
    J = torch.cat(Js, dim=1)
    return d, qvel, feet, J


def _mask(grf, contact, nc):
    return torch.cat([grf[3 * i:3 * i + 3] * contact[i] for i in range(nc)])


def quadruped_wb_dynamics(m, contact_id, body_id, n_joints, dt, x, u, t, parameter):
    """reference mpx/utils/models.py:39-117 (alpha=25, no regulariser, feet FL,FR,RL,RR)."""
    d, qvel, feet, J = _wb_common(m, contact_id, body_id, n_joints, x)
    M, D = d.qLD, d.qfrc_bias
    contact = parameter[t, :4]
    tau = torch.cat([torch.zeros(6, dtype=x.dtype), u])
    current_leg = torch.cat(feet)
    alpha = 25
    g_dot = J.T @ qvel
    baumgarte_term = -2 * alpha * g_dot
    JT_M_invJ = J.T @ mjx.cho_solve_upper(M, J)
    rhs = -J.T @ mjx.cho_solve_upper(M, tau - D) + baumgarte_term
    cho = torch.linalg.cholesky(JT_M_invJ, upper=True)
    grf = mjx.cho_solve_upper(cho, rhs)
    grf = _mask(grf, contact, 4)
    v = qvel + mjx.cho_solve_upper(M, tau - D + J @ grf) * dt
    p = x[:3] + v[:3] * dt
    quat = math.quat_integrate(x[3:7], v[3:6], dt)
    q = x[7:7 + n_joints] + v[6:6 + n_joints] * dt
    return torch.cat([p, quat, q, v, current_leg, grf])


def h1_wb_dynamics(m, contact_id, body_id, n_joints, dt, x, u, t, parameter):
    """reference mpx/utils/models.py:119-169 (alpha=5, +1e-3 I, feet FL,RL on body_id[0]; FR,RR on body_id[1])."""
    bodies = [body_id[0], body_id[0], body_id[1], body_id[1]]
    d, qvel, feet, J = _wb_common(m, contact_id, bodies, n_joints, x)
    M, D = d.qLD, d.qfrc_bias
    contact = parameter[t, :4]
    tau = torch.cat([torch.zeros(6, dtype=x.dtype), u])
    g_dot = J.T @ qvel
    alpha = 5
    baumgarte_term = -2 * alpha * g_dot
    JT_M_invJ = J.T @ mjx.cho_solve_upper(M, J)
    rhs = -J.T @ mjx.cho_solve_upper(M, tau - D) + baumgarte_term
    epsilon = 1e-3
    reg = JT_M_invJ + epsilon * torch.eye(JT_M_invJ.shape[0], dtype=x.dtype)
    cho = torch.linalg.cholesky(reg, upper=True)
    grf = mjx.cho_solve_upper(cho, rhs)
    grf = _mask(grf, contact, 4)
    v = qvel + mjx.cho_solve_upper(M, tau - D + J @ grf) * dt
    p = x[:3] + v[:3] * dt
    quat = math.quat_integrate(x[3:7], v[3:6], dt)
    q = x[7:7 + n_joints] + v[6:6 + n_joints] * dt
    return torch.cat([p, quat, q, v] + feet + [grf])


def talos_wb_dynamics(m, contact_id, body_id, n_joints, dt, x, u, t, parameter):
    """reference mpx/utils/models.py:170-217 (GRFs are controls u[nj:], 8 foot points on 2 bodies)."""
    bodies = [body_id[0]] * 4 + [body_id[1]] * 4
    d, qvel, feet, J = _wb_common(m, contact_id, bodies, n_joints, x)
    M, D = d.qLD, d.qfrc_bias
    contact = parameter[t, :8]
    tau = torch.cat([torch.zeros(6, dtype=x.dtype), u[:n_joints]])
    grf = _mask(u[n_joints:], contact, 8)
    v = qvel + mjx.cho_solve_upper(M, tau - D + J @ grf) * dt
    p = x[:3] + v[:3] * dt
    quat = math.quat_integrate(x[3:7], v[3:6], dt)
    q = x[7:7 + n_joints] + v[6:6 + n_joints] * dt
    return torch.cat([p, quat, q, v] + feet)
