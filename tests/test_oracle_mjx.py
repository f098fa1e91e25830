"""Independent pins of the oracle: the restated MJX subset and dynamics against first-principles
mechanics and float64 finite differences, and the restated upstream tests of the reference's own solver.

The reference cannot be run here (PARITY UNPINNED, oracle/__init__.py), so the oracle's rigid-body terms are pinned by
what they must equal physically, using NOTHING of the restated CRBA / RNE code -- only forward kinematics (body COM
positions and orientations) and finite differences of it:

* mass matrix      qM_ij  = sum_b m_b Jv_bi . Jv_bj + Jw_bi' I_b Jw_bj  (+ armature)      [kinetic-energy Hessian]
* bias force       qfrc_bias_i = sum_b m_b (a_b - g) . Jv_bi + (I_b alpha_b + w_b x I_b w_b) . Jw_bi   [Kane / virtual power]
  with Jv, Jw the body COM / angular velocity Jacobians w.r.t. the quasi-velocities (central differences of FK along
  each dof) and a_b, alpha_b the accelerations along the constant-velocity motion (second differences of FK);
  at rest this is the gradient of the potential energy;
* foot Jacobians   mjx.jac == d(foot position)/d(dof) by central differences;
* dynamics Jacobians (jacrev, what the fixtures hold) == float64 central differences for all four robots and the SRBD
  model; and the safe-norm ZERO-Jacobian branch at exactly zero angular velocity, which finite differences do not see;
* the reference's own tests: acrobot swing-up `g < 51` (tests/simple_tests.py:115-139) and the T = 2 LQ problem
  (70-113) through the restated legacy solver `primal_dual_ilqr` (old_optimazers.py:280-474).
"""
import importlib
import math

import numpy as np
import pytest
import torch
from torch.func import jacrev, vmap

from oracle import legacy, mjx_math, mjx_smooth
from oracle.problems import Problem, synthetic_inputs

ROBOTS = ["aliengo", "go2", "h1", "talos"]
F64 = torch.float64


def _integrate(m, qpos, v, h):
    """qpos (+) h v with MuJoCo's conventions: free joint = world-frame linear velocity, LOCAL-frame angular velocity
    (quat <- quat * exp(h w)); hinges add."""
    q = qpos.clone()
    q[:3] = qpos[:3] + h * v[:3]
    w = v[3:6]
    ang = float(torch.linalg.norm(w)) * h
    if ang != 0.0:
        axis = w / torch.linalg.norm(w)
        dq = torch.cat([torch.tensor([math.cos(ang / 2)], dtype=F64), axis * math.sin(ang / 2)])
        q[3:7] = mjx_math.quat_mul(qpos[3:7], dq)
    q[7:] = qpos[7:] + h * v[6:]
    return q


def _fk(m, qpos):
    d = mjx_smooth.kinematics(m, qpos)
    return torch.stack(d.xipos), torch.stack(d.ximat), d


def _vee(S):
    return torch.stack([S[..., 2, 1] - S[..., 1, 2], S[..., 0, 2] - S[..., 2, 0], S[..., 1, 0] - S[..., 0, 1]], dim=-1) * 0.5


def _body_jacobians(m, qpos, h=1e-6):
    """Jv [nbody, nv, 3], Jw [nbody, nv, 3] by central differences of forward kinematics along each quasi-velocity."""
    Jv, Jw = [], []
    for i in range(m.nv):
        e = torch.zeros(m.nv, dtype=F64)
        e[i] = 1.0
        xp, Rp, _ = _fk(m, _integrate(m, qpos, e, h))
        xm, Rm, _ = _fk(m, _integrate(m, qpos, e, -h))
        Jv.append((xp - xm) / (2 * h))
        Jw.append(_vee(Rp @ Rm.transpose(-1, -2)) / (2 * h))   # R+ R-' = exp(2h [w]x)
    return torch.stack(Jv, dim=1), torch.stack(Jw, dim=1)


def _world_inertia(m, ximat):
    return ximat @ torch.diag_embed(m.body_inertia.to(F64)) @ ximat.transpose(-1, -2)


@pytest.fixture(scope="module", params=ROBOTS)
def robot(request):
    cfg = importlib.import_module(f"mpx_b200.config.config_{request.param}")
    pr = Problem(cfg)
    inp = synthetic_inputs(pr, 1, seed=9, scale=3.0)
    nj = cfg.n_joints
    x0 = torch.as_tensor(inp["x0"][0], dtype=F64)
    return cfg, pr, x0[:7 + nj].clone(), x0[7 + nj:13 + 2 * nj].clone()


def test_mass_matrix_is_the_kinetic_energy_hessian(robot):
    cfg, pr, qpos, qvel = robot
    m = pr.m
    d = mjx_smooth.fwd_position(m, qpos)
    Jv, Jw = _body_jacobians(m, qpos)
    _, ximat, _ = _fk(m, qpos)
    Iw = _world_inertia(m, ximat)
    mass = m.body_mass.to(F64)
    M = torch.einsum("b,bix,bjx->ij", mass, Jv, Jv) + torch.einsum("bix,bxy,bjy->ij", Jw, Iw, Jw)
    M = M + torch.diag(m.dof_armature.to(F64))
    assert float((d.qM - M).abs().max()) < 1e-7 * float(M.abs().max())
    assert abs(float(d.qM[0, 0]) - float(mass.sum())) < 1e-9          # total mass
    assert float(torch.linalg.eigvalsh(d.qM).min()) > 0


@pytest.mark.parametrize("at_rest", [True, False])
def test_bias_force_is_newton_euler_by_finite_differences(robot, at_rest):
    """qfrc_bias against Kane's equations evaluated from FK alone; at rest it is the potential-energy gradient."""
    cfg, pr, qpos, qvel = robot
    m = pr.m
    v = torch.zeros_like(qvel) if at_rest else qvel
    d = mjx_smooth.fwd_position(m, qpos)
    d = mjx_smooth.fwd_velocity(m, d, v)
    Jv, Jw = _body_jacobians(m, qpos)
    x0, R0, _ = _fk(m, qpos)
    Iw = _world_inertia(m, R0)
    mass = m.body_mass.to(F64)
    h = 1e-4
    xs = {s: _fk(m, _integrate(m, qpos, v, s * h)) for s in (-2, -1, 1, 2)}
    acc = (xs[1][0] - 2 * x0 + xs[-1][0]) / h ** 2
    omega = _vee(xs[1][1] @ xs[-1][1].transpose(-1, -2)) / (2 * h)
    om_p = _vee(xs[2][1] @ R0.transpose(-1, -2)) / (2 * h)       # angular velocity at t = +h
    om_m = _vee(R0 @ xs[-2][1].transpose(-1, -2)) / (2 * h)      # at t = -h
    alpha = (om_p - om_m) / (2 * h)
    g = m.gravity.to(F64)
    lin = mass[:, None] * (acc - g[None])
    ang = torch.einsum("bxy,by->bx", Iw, alpha) + torch.cross(omega, torch.einsum("bxy,by->bx", Iw, omega), dim=-1)
    bias = torch.einsum("bx,bix->i", lin, Jv) + torch.einsum("bx,bix->i", ang, Jw)
    scale = float(bias.abs().max())
    assert float((d.qfrc_bias - bias).abs().max()) < 2e-5 * scale, float((d.qfrc_bias - bias).abs().max()) / scale
    if at_rest:
        # potential energy V = -sum_b m_b g . x_b: bias = dV/dq along the quasi-velocities
        grad_v = -torch.einsum("b,x,bix->i", mass, g, Jv)
        assert float((d.qfrc_bias - grad_v).abs().max()) < 1e-7 * scale


def test_foot_jacobians_are_derivatives_of_foot_positions(robot):
    cfg, pr, qpos, qvel = robot
    m = pr.m
    d = mjx_smooth.fwd_position(m, qpos)
    nc = cfg.n_contact
    bodies = pr.body_id if len(pr.body_id) == nc else [pr.body_id[0]] * (nc // 2) + [pr.body_id[1]] * (nc // 2)
    h = 1e-6
    feet = lambda q: torch.stack([mjx_smooth.kinematics(m, q).geom_xpos[g] for g in pr.contact_id])
    for i in range(m.nv):
        e = torch.zeros(m.nv, dtype=F64)
        e[i] = 1.0
        fd = (feet(_integrate(m, qpos, e, h)) - feet(_integrate(m, qpos, e, -h))) / (2 * h)
        for k in range(nc):
            jp, _ = mjx_smooth.jac(m, d, d.geom_xpos[pr.contact_id[k]], bodies[k])
            assert float((jp[i] - fd[k]).abs().max()) < 1e-7


def _central_jacobians(f, x, u, h=1e-6):
    n, m = x.shape[0], u.shape[0]
    eye_x, eye_u = torch.eye(n, dtype=F64), torch.eye(m, dtype=F64)
    fx = (vmap(lambda dx: f(x + h * dx, u))(eye_x) - vmap(lambda dx: f(x - h * dx, u))(eye_x)) / (2 * h)
    fu = (vmap(lambda du: f(x, u + h * du))(eye_u) - vmap(lambda du: f(x, u - h * du))(eye_u)) / (2 * h)
    return fx.T, fu.T


@pytest.mark.parametrize("name", ["srbd"] + ROBOTS)
def test_dynamics_jacobians_equal_central_differences(name):
    """A_t, B_t as the fixtures hold them (jacrev of the restated dynamics == jax.jacobian in the reference,
    optimizers.py:166,171) against float64 central differences at a generic state."""
    cfg = importlib.import_module(f"mpx_b200.config.config_{name}")
    pr = Problem(cfg)
    inp = synthetic_inputs(pr, 1, seed=13, scale=2.0)
    rng = np.random.default_rng(2)
    x = torch.as_tensor(inp["x0"][0], dtype=F64)
    u = torch.as_tensor(np.asarray(cfg.u_ref, dtype=np.float64) + rng.normal(0, 3.0, cfg.m), dtype=F64)
    par = torch.as_tensor(inp["parameter"][0], dtype=F64)
    for t in (0, 11):
        f = lambda x_, u_: pr.dynamics(x_, u_, t, par)
        A, B = jacrev(f, argnums=(0, 1))(x, u)
        Afd, Bfd = _central_jacobians(f, x, u)
        sa, sb = float(A.abs().max()), float(B.abs().max())
        assert float((A - Afd).abs().max()) < 2e-6 * sa, (name, t, float((A - Afd).abs().max()) / sa)
        assert float((B - Bfd).abs().max()) < 2e-6 * max(sb, 1e-3), (name, t)


def test_safe_norm_zero_jacobian_branch():
    """mjx math.norm is gradient-safe: at exactly zero angular velocity quat_integrate has ZERO Jacobian w.r.t. the
    velocity (the reference inherits this, SURVEY.md 7 'hard parts'), although the true derivative -- which central
    differences see -- is not zero.  SRBD with zero ground-reaction forces keeps omega+ = omega = 0 exactly."""
    cfg = importlib.import_module("mpx_b200.config.config_srbd")
    pr = Problem(cfg)
    inp = synthetic_inputs(pr, 1, seed=1, cold_start=True)
    x = torch.as_tensor(inp["x0"][0], dtype=F64)
    u = torch.zeros(cfg.m, dtype=F64)
    par = torch.as_tensor(inp["parameter"][0], dtype=F64)
    f = lambda x_, u_: pr.dynamics(x_, u_, 0, par)
    A, B = jacrev(f, argnums=(0, 1))(x, u)
    assert float(A[3:7, 10:13].abs().max()) == 0.0 and float(B[3:7].abs().max()) == 0.0
    Afd, _ = _central_jacobians(f, x, u)
    assert float(Afd[3:7, 10:13].abs().max()) > 1e-3      # d quat / d omega = dt/2 in truth
    # off the singular point the two agree again
    x2 = x.clone()
    x2[10:13] = torch.tensor([1e-3, -2e-3, 5e-4], dtype=F64)
    A2 = jacrev(lambda x_: f(x_, u))(x2)
    A2fd, _ = _central_jacobians(f, x2, u)
    assert float((A2 - A2fd).abs().max()) < 1e-6


# ------------------------------------------------------------------ the reference's own tests, restated
def _acrobot(x, u, t, params=(1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 1.0)):
    """The acrobot of the reference's test (tests/simple_tests.py:31-68): two links, torque at the elbow."""
    m1, m2, l1, lc1, lc2, I1, I2 = params
    grav = 9.8
    th1, th2, w1, w2 = x[0], x[1], x[2], x[3]
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * torch.cos(th2)) + I1 + I2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * torch.cos(th2)) + I2
    phi2 = m2 * lc2 * grav * torch.cos(th1 + th2 - math.pi / 2.0)
    phi1 = (-m2 * l1 * lc2 * w2 ** 2 * torch.sin(th2) - 2 * m2 * l1 * lc2 * w2 * w1 * torch.sin(th2)
            + (m1 * lc1 + m2 * l1) * grav * torch.cos(th1 - math.pi / 2) + phi2)
    a2 = (u[0] + d2 / d1 * phi1 - m2 * l1 * lc2 * w1 ** 2 * torch.sin(th2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
    a1 = -(d2 * a2 + phi1) / d1
    return torch.stack([w1, w2, a1, a2])


def test_reference_acrobot_solve():
    """tests/simple_tests.py:115-139: swing-up over T = 50 Euler steps of 0.1 s; asserts g < 51 and no_errors."""
    T = 50
    goal = torch.tensor([math.pi, 0.0, 0.0, 0.0], dtype=F64)
    dynamics = lambda x, u, t: x + 0.1 * _acrobot(x, u, t)      # trajax.integrators.euler(acrobot, dt=0.1)
    params = (1000.0, 0.1, 0.01)

    def cost(x, u, t):
        delta = x - goal
        if t == T:
            return 0.5 * params[0] * torch.dot(delta, delta)
        return 0.5 * params[1] * torch.dot(delta, delta) + 0.5 * params[2] * torch.dot(u, u)
    z = lambda *s: torch.zeros(*s, dtype=F64)
    trace = []
    X, U, V, iterations, g, c, no_errors = legacy.primal_dual_ilqr(cost, dynamics, z(4), z(T + 1, 4), z(T, 1), z(T + 1, 4),
                                                                  trace=trace)
    assert no_errors
    assert float(g) < 51.0
    assert float(torch.sum(c * c)) <= 1e-4 and 0 < iterations < 100
    # the swing-up reaches the upright goal
    assert float((X[-1] - goal).abs().max()) < 0.1


def test_reference_simple_lqr_problem_legacy_solver():
    """tests/simple_tests.py:70-113 through the legacy loop: an LQ problem converges in one full Newton step."""
    t64 = lambda a: torch.as_tensor(a, dtype=F64)

    def dynamics(x, u, t):
        A = 6 * t * torch.ones(2, 2, dtype=F64) + t64(np.arange(4).reshape(2, 2))
        B = 6 * t * torch.ones(2, 1, dtype=F64) + t64(np.arange(4, 6).reshape(2, 1))
        c = 6 * t * torch.ones(2, dtype=F64) + t64(np.arange(6, 8))
        return A @ x + B @ u + c

    def cost(x, u, t):
        Q = (t + 1) * torch.diag(t64(np.arange(1, 3)))
        R = (t + 1) * t64([[3.0]])
        q = -(t + 1) * t64(np.arange(91, 93))
        r = -(t + 1) * t64([99.0])
        M = (t + 1) / 100.0 * t64(np.arange(55, 57).reshape(2, 1))
        return 0.5 * x @ Q @ x + 0.5 * u @ R @ u + x @ M @ u + q @ x + r @ u
    x0 = t64([-11.0, -22.0])
    z = lambda *s: torch.zeros(*s, dtype=F64)
    trace = []
    X, U, V, iterations, g, c, no_errors = legacy.primal_dual_ilqr(cost, dynamics, x0, z(3, 2), z(2, 1), z(3, 2), trace=trace)
    assert no_errors and iterations <= 3
    assert trace[0]["alpha"] == 1.0                      # the full step is accepted on an LQ problem
    assert float(torch.sum(c * c)) < 1e-16               # and lands on the constraint manifold
