"""Restatement of ``mujoco.mjx._src.math`` (mujoco-mjx 3.3.0) -- the quaternion helpers the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  mujoco-mjx is a third-party dependency that is NOT in
/root/reference (pinned by reference pyproject.toml:15-21 as ``mujoco-mjx==3.3.0``); the functions below
restate its published algorithm.  Reference call sites: mpx/utils/models.py:26,32,112,165,213;
mpx/utils/objectives.py:31,34,54,115,151; mpx/utils/mpc_utils.py:26-28.

Quaternions are (w, x, y, z).
"""
import torch


def _stack(xs):
    return torch.stack([x if torch.is_tensor(x) else torch.as_tensor(x) for x in xs])


def cross(a, b):
    return torch.stack([
        a[1] * b[2] - a[2] * b[1],
        a[2] * b[0] - a[0] * b[2],
        a[0] * b[1] - a[1] * b[0],
    ])


def norm(x):
    """Gradient-safe norm: zero value AND zero gradient when ``allclose(x, 0)`` (atol 1e-8)."""
    is_zero = torch.all(torch.abs(x) <= 1e-8)
    x = torch.where(is_zero, torch.ones_like(x), x)
    n = torch.sqrt(torch.sum(x * x))
    n = torch.where(is_zero, torch.zeros_like(n), n)
    return n


def normalize_with_norm(x):
    n = norm(x)
    x = x / (n + 1e-6 * (n == 0.0).to(x.dtype))
    return x, n


def normalize(x):
    return normalize_with_norm(x)[0]


def quat_mul(u, v):
    return torch.stack([
        u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3],
        u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
        u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1],
        u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0],
    ])


def quat_inv(q):
    return q * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)


def rotate(vec, quat):
    s, u = quat[0], quat[1:]
    r = 2 * (torch.dot(u, vec) * u) + (s * s - torch.dot(u, u)) * vec
    r = r + 2 * s * cross(u, vec)
    return r


def axis_angle_to_quat(axis, angle):
    s, c = torch.sin(angle * 0.5), torch.cos(angle * 0.5)
    return torch.cat([c.reshape(1), axis * s])


def quat_to_mat(q):
    w, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)]),
        torch.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)]),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]),
    ])


def quat_to_axis_angle(q):
    axis, sin_a_2 = normalize_with_norm(q[1:])
    angle = 2 * torch.atan2(sin_a_2, q[0])
    angle = torch.where(angle > torch.pi, angle - 2 * torch.pi, angle)
    return axis, angle


def quat_sub(u, v):
    """(u - v) as a 3D rotation vector: axis*angle of v^-1 * u."""
    q = quat_mul(quat_inv(v), u)
    axis, angle = quat_to_axis_angle(q)
    return axis * angle


def quat_integrate(q, v, dt):
    v, norm_ = normalize_with_norm(v)
    angle = dt * norm_
    q_res = axis_angle_to_quat(v, angle)
    q_res = quat_mul(q, q_res)
    return normalize(q_res)


def inert_mul(i, v):
    """cinert (6 triu rot inertia [xx,yy,zz,xy,xz,yz], 3 mass*offset, mass) times motion [ang, lin]."""
    inr = torch.stack([
        torch.stack([i[0], i[3], i[4]]),
        torch.stack([i[3], i[1], i[5]]),
        torch.stack([i[4], i[5], i[2]]),
    ])
    pos, mass = i[6:9], i[9]
    ang = inr @ v[:3] + cross(pos, v[3:])
    vel = mass * v[3:] - cross(pos, v[:3])
    return torch.cat([ang, vel])


def motion_cross(u, v):
    ang = cross(u[:3], v[:3])
    vel = cross(u[3:], v[:3]) + cross(u[:3], v[3:])
    return torch.cat([ang, vel])


def motion_cross_force(v, f):
    ang = cross(v[:3], f[:3]) + cross(v[3:], f[3:])
    vel = cross(v[:3], f[3:])
    return torch.cat([ang, vel])
