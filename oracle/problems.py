"""Problem builders for the oracle: bind a config module to the restated cost/dynamics/Hessian, and the
synthetic-input protocol of SURVEY.md section 8(d).  TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from functools import partial

import numpy as np
import torch

from mpx_b200.mjcf import load_model
from . import models, objectives, mpc_utils, mjx_smooth
from . import mjx_math as math
from . import primal_dual_ilqr as pdi

_DYN = {"quadruped_srbd_dynamics": models.quadruped_srbd_dynamics,
        "quadruped_wb_dynamics": models.quadruped_wb_dynamics,
        "h1_wb_dynamics": models.h1_wb_dynamics,
        "talos_wb_dynamics": models.talos_wb_dynamics}
_OBJ = {"quadruped_srbd_obj": objectives.quadruped_srbd_obj,
        "quadruped_wb_obj": objectives.quadruped_wb_obj,
        "h1_wb_obj": objectives.h1_wb_obj,
        "talos_wb_obj": objectives.talos_wb_obj}
_HES = {"quadruped_srbd_hessian_gn": objectives.quadruped_srbd_hessian_gn,
        "quadruped_wb_hessian_gn": objectives.quadruped_wb_hessian_gn,
        "h1_wb_hessian_gn": objectives.h1_wb_hessian_gn,
        "talos_wb_hessian_gn": objectives.talos_wb_hessian_gn}


def _unwrap(desc):
    """config attribute -> (descriptor name, bound leading args) (configs use functools.partial(fn, True))."""
    args = ()
    while isinstance(desc, partial):
        args = desc.args + args
        desc = desc.func
    return desc.name, args


class Problem:
    """cost/dynamics/hessian closures with the reference's partial-application order."""

    def __init__(self, config, dtype=torch.float64):
        self.config, self.dtype = config, dtype
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
        self.is_srbd = _unwrap(config.dynamics)[0] == "quadruped_srbd_dynamics"
        self.rm = load_model(config.model_path)
        self.W = t(config.W)
        cname, cargs = _unwrap(config.cost)
        hname, hargs = _unwrap(config.hessian_approx)
        dname, _ = _unwrap(config.dynamics)
        if self.is_srbd:
            inertia = t(config.inertia)
            # mpc_wrapper_srbd.py:60-66
            self.cost = partial(_OBJ[cname], config.n_contact, config.N)
            self.hessian_approx = partial(_HES[hname], config.n_contact)
            self.dynamics = partial(_DYN[dname], config.mass, inertia, torch.linalg.inv(inertia), config.dt)
            self.robot_mass = config.mass
        else:
            self.m = mjx_smooth.put_model(self.rm, dtype)
            self.contact_id = [self.rm.geom_id(nm) for nm in config.contact_frame]
            self.body_id = [self.rm.body_id(nm) for nm in config.body_name]
            # mpc_wrapper.py:66-72
            self.cost = partial(_OBJ[cname], *cargs, config.n_joints, config.n_contact, config.N)
            self.hessian_approx = partial(_HES[hname], *hargs, config.n_joints, config.n_contact)
            self.dynamics = partial(_DYN[dname], self.m, self.contact_id, self.body_id, config.n_joints, config.dt)
            # mpc_wrapper.py:45-46: robot_mass = qM[0] at qpos0 = total mass
            self.robot_mass = float(self.rm.total_mass)

    def foot_fk(self, qpos):
        d = mjx_smooth.kinematics(self.m, torch.as_tensor(qpos, dtype=torch.float64))
        return torch.cat([d.geom_xpos[g] for g in self.contact_id]).numpy()

    def mpc(self, limited_memory, reference, parameter, x0, X, U, V, W=None, detail=None, line_search="filter"):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=self.dtype)
        W = self.W if W is None else t(W)
        out = pdi.mpc(self.cost, self.dynamics, self.hessian_approx, limited_memory, t(reference), t(parameter), W,
                      t(x0), t(X), t(U), t(V), detail=detail, line_search=line_search)
        return out


def rpy_to_quat(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


def synthetic_inputs(problem, n_env, seed=0, cold_start=False, scale=1.0):
    """per-env random robot state + command -> (reference, parameter, x0, X0, U0, V0) float64 numpy.

    Draw order per env: height, rpy, q, v_lin, omega, dq, command, gait phase.  ``cold_start`` gives the exact
    identity-quaternion / zero-velocity state of ``reset`` (the safe-norm discontinuity case).  ``scale`` multiplies
    the ranges of the random robot state (not of the command): states far from the nominal stance, where the line
    search has to cut the step.
    """
    cfg = problem.config
    rng = np.random.default_rng(seed)
    N, nc = cfg.N, cfg.n_contact
    out = {k: [] for k in ("reference", "parameter", "x0", "X0", "U0", "V0")}
    for _ in range(n_env):
        dz = scale * rng.uniform(-0.02, 0.02)
        rpy = scale * rng.uniform(-0.1, 0.1, 3)
        nj = 0 if problem.is_srbd else cfg.n_joints
        dq_pos = scale * rng.uniform(-0.05, 0.05, cfg.n_joints)
        v_lin = scale * rng.uniform(-0.2, 0.2, 3)
        omega = scale * rng.uniform(-0.2, 0.2, 3)
        dq = scale * rng.uniform(-0.5, 0.5, cfg.n_joints)
        cmd = np.array([rng.uniform(-0.2, 0.5), rng.uniform(-0.1, 0.1), 0, 0, 0, rng.uniform(-0.3, 0.3),
                        cfg.robot_height])
        phase = rng.uniform(0, 1)
        if cold_start:
            dz, rpy, dq_pos, v_lin, omega, dq = 0.0, np.zeros(3), 0 * dq_pos, 0 * v_lin, 0 * omega, 0 * dq
        p = np.asarray(cfg.p0, dtype=np.float64) + np.array([0, 0, dz])
        quat = rpy_to_quat(*rpy) if not cold_start else np.array([1.0, 0, 0, 0])
        timer_t = np.mod(np.asarray(cfg.timer_t, dtype=np.float64) + phase, 1.0)
        contact = np.ones(nc)
        if problem.is_srbd:
            yaw = rpy[2]
            Ryaw = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
            feet = (np.array([p[0], p[1], 0.0]) + np.asarray(cfg.p_legs0).reshape(nc, 3) @ Ryaw.T).reshape(-1)
            x0 = np.concatenate([p, quat, v_lin, omega])
            ref, par, _, _ = mpc_utils.reference_generator_srbd(
                cfg.use_terrain_estimator, N, cfg.dt, nc, cfg.mass, cfg.p_legs0, timer_t, x0, feet, cmd,
                cfg.duty_factor, cfg.step_freq, cfg.step_height, feet.copy(), contact, cfg.clearence_speed)
        else:
            q = np.asarray(cfg.q0, dtype=np.float64) + dq_pos
            qpos = np.concatenate([p, quat, q])
            qvel = np.concatenate([v_lin, omega, dq])
            feet = problem.foot_fk(qpos)
            tail = [feet, np.zeros(3 * nc)] if cfg.grf_as_state else [feet]
            x0 = np.concatenate([qpos, qvel] + tail)
            ref, par, _ = mpc_utils.reference_generator(
                cfg.use_terrain_estimation, N, cfg.dt, cfg.n_joints, nc, problem.robot_mass, cfg.p_legs0, cfg.q0,
                timer_t, x0, feet, cmd, cfg.duty_factor, cfg.step_freq, cfg.step_height, feet.copy(), contact, 0.2)
        out["reference"].append(ref)
        out["parameter"].append(par)
        out["x0"].append(x0)
        out["X0"].append(np.tile(x0, (N + 1, 1)))
        out["U0"].append(np.tile(np.asarray(cfg.u_ref, dtype=np.float64), (N, 1)))
        out["V0"].append(np.zeros((N + 1, cfg.n)))
    return {k: np.stack(v) for k, v in out.items()}
