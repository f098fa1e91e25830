"""CPU parity of the kernels' __host__ __device__ arithmetic (tests/cpu_harness, g++ build of wb.cuh) against the
float64 oracle: MJX terms (qM, qfrc_bias, foot Jacobians), whole-body dynamics, their forward-mode Jacobians and
the cost gradient, for every robot of the reference configs."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest
import torch
from torch.func import jacrev

from mpx_b200 import _lib
from oracle import mjx_smooth
from oracle.problems import Problem, synthetic_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness"))
import build as harness_build  # noqa: E402

ROBOTS = ["aliengo", "go2", "h1", "talos"]


@pytest.fixture(scope="module")
def H():
    return C.CDLL(harness_build.build())


@pytest.fixture(scope="module", params=ROBOTS)
def robot(request):
    cfg = importlib.import_module(f"mpx_b200.config.config_{request.param}")
    pr = Problem(cfg)
    inp = synthetic_inputs(pr, 2, seed=4)
    return cfg, pr, inp, _lib.make_desc(cfg)


_P32 = {}


def _problem32(cfg):
    if cfg.__name__ not in _P32:
        _P32[cfg.__name__] = Problem(cfg, dtype=torch.float32)
    return _P32[cfg.__name__]


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(1.0, np.abs(b).max()))


def _state(cfg, inp, rng):
    """A state off the warm start: perturbed x0 with non-zero GRF tail, random torques."""
    x = inp["x0"][0].copy()
    x[7:7 + cfg.n_joints] += rng.uniform(-0.2, 0.2, cfg.n_joints)
    if cfg.grf_as_state:
        x[-3 * cfg.n_contact:] = rng.uniform(-5, 30, 3 * cfg.n_contact)
    u = rng.uniform(-5, 5, cfg.m)
    return x, u


@pytest.mark.parametrize("fn", ["h_wb_terms", "h_wb_terms_dfs"])
def test_mjx_terms(H, robot, fn):
    """qM, qfrc_bias, foot Jacobians and foot positions: the COM-centred arrays form (trial kernel) and the
    single-pass path-stack form about the base position (linearise kernel) against the restated MJX."""
    cfg, pr, inp, desc = robot
    rng = np.random.default_rng(0)
    x, _ = _state(cfg, inp, rng)
    nq, nv, nc = 7 + cfg.n_joints, 6 + cfg.n_joints, cfg.n_contact
    qpos, qvel = x[:nq], x[nq:nq + nv]
    M, bias, J, feet = (np.zeros(s, np.float32) for s in ((nv, nv), (nv,), (nv, 3 * nc), (3 * nc,)))
    assert getattr(H, fn)(C.byref(desc), _p(_f(qpos)), _p(_f(qvel)), _p(M), _p(bias), _p(J), _p(feet)) == 0
    d = mjx_smooth.fwd_position(pr.m, torch.as_tensor(qpos))
    d = mjx_smooth.fwd_velocity(pr.m, d, torch.as_tensor(qvel))
    per_foot = [desc.contact_body[i] for i in range(nc)]
    fo = [d.geom_xpos[g] for g in pr.contact_id]
    Jo = torch.cat([mjx_smooth.jac(pr.m, d, fo[i], per_foot[i])[0] for i in range(nc)], dim=1)
    assert _rel(M, d.qM.numpy()) < 2e-6
    assert _rel(bias, d.qfrc_bias.numpy()) < 5e-6
    assert _rel(J, Jo.numpy()) < 2e-6
    assert _rel(feet, torch.cat(fo).numpy()) < 2e-6
    # physical invariant: the translational block of qM is total mass * I
    assert abs(M[0, 0] - pr.robot_mass) / pr.robot_mass < 1e-5


@pytest.mark.parametrize("fn", ["h_wb_dynamics", "h_wb_dynamics_hybrid", "h_wb_dynamics_tree"])
@pytest.mark.parametrize("contact_pattern", ["all", "trot"])
def test_dynamics_and_jacobians(H, robot, contact_pattern, fn):
    """x+ = f(x,u), A = df/dx, B = df/du: plain forward mode through the whole dynamics (h_wb_dynamics) and the
    kernel's path-stack + analytic contact-solve derivative (h_wb_dynamics_hybrid) against jacrev of the oracle."""
    cfg, pr, inp, desc = robot
    rng = np.random.default_rng(1)
    x, u = _state(cfg, inp, rng)
    n, m, nc = cfg.n, cfg.m, cfg.n_contact
    contact = np.ones(nc) if contact_pattern == "all" else (np.arange(nc) % 2).astype(np.float64)
    par = np.tile(contact, (cfg.N + 1, 1))
    xn, A, B = np.zeros(n, np.float32), np.zeros((n, n), np.float32), np.zeros((n, m), np.float32)
    assert getattr(H, fn)(C.byref(desc), _p(_f(x)), _p(_f(u)), _p(_f(contact)), _p(xn), _p(A), _p(B)) == 0
    xt, ut, pt = torch.as_tensor(x), torch.as_tensor(u), torch.as_tensor(par)
    f = lambda x_, u_: pr.dynamics(x_, u_, 0, pt)
    Ao, Bo = jacrev(f, 0)(xt, ut), jacrev(f, 1)(xt, ut)
    # FP32 budget: the oracle's own float32 evaluation against float64 (H1's regularised 12x12 contact system of
    # two line contacts is ill-conditioned: models.py:155-157), floor 2e-5 / 1e-4
    pr32 = _problem32(cfg)
    f32 = lambda x_, u_: pr32.dynamics(x_, u_, 0, pt.float())
    x32, u32 = xt.float(), ut.float()
    bx = max(2e-5, 4 * _rel(f32(x32, u32).numpy(), f(xt, ut).numpy()))
    bA = max(1e-4, 4 * _rel(jacrev(f32, 0)(x32, u32).numpy(), Ao.numpy()))
    bB = max(1e-4, 4 * _rel(jacrev(f32, 1)(x32, u32).numpy(), Bo.numpy()))
    assert _rel(xn, f(xt, ut).numpy()) < bx, (bx,)
    assert _rel(A, Ao.numpy()) < bA, (bA,)
    assert _rel(B, Bo.numpy()) < bB, (bB,)
    # the foot / GRF tail of the state never enters f (models.py:61,95,109-115)
    ndyn = 13 + 2 * cfg.n_joints
    assert np.all(A[:, ndyn:] == 0) and float(Ao[:, ndyn:].abs().max()) == 0.0


@pytest.mark.parametrize("terminal", [0, 1])
def test_cost_and_gradient(H, robot, terminal):
    cfg, pr, inp, desc = robot
    rng = np.random.default_rng(2)
    x, u = _state(cfg, inp, rng)
    if terminal:
        u = np.zeros_like(u)
    t = cfg.N if terminal else 3
    ref = inp["reference"][0]
    cost, grad = np.zeros(1, np.float32), np.zeros(cfg.n + cfg.m, np.float32)
    assert H.h_wb_cost(C.byref(desc), _p(_f(cfg.W)), _p(_f(ref[t])), _p(_f(x)), _p(_f(u)), terminal, _p(cost), _p(grad)) == 0
    xt, ut, rt = torch.as_tensor(x), torch.as_tensor(u), torch.as_tensor(ref)
    c = lambda x_, u_: pr.cost(pr.W, rt, x_, u_, t)
    co = float(c(xt, ut))
    assert abs(float(cost[0]) - co) / max(1.0, abs(co)) < 2e-5
    go = torch.cat([jacrev(c, 0)(xt, ut), jacrev(c, 1)(xt, ut)]).numpy()
    assert _rel(grad, go) < 2e-5
