"""Whole-body MPC wrappers -- same classes / constructors / ``run`` / ``reset`` as reference
``mpx/utils/mpc_wrapper.py`` (mpx_data 16-28, BatchedMPCControllerWrapper 29-192, MPCControllerWrapper 194-468) with
the jitted JAX solve replaced by libmpxb.so.

Differences that follow from dropping JAX (documented in INTEGRATION.md):
  * the reference's ``BatchedMPCControllerWrapper.run`` is a per-environment function that the caller wraps in
    ``jax.jit(jax.vmap(...))`` (examples/multi_env.py:46-49); here the env axis is native: ``make_data`` returns the
    batched ``mpx_data`` ([n_env, ...] CUDA tensors) and ``run(data, x0, input)`` takes ``x0 [n_env, n]``,
    ``input [n_env, 7]`` and returns ``(data, tau [n_env, n_joints])`` without synchronising;
  * the host-side CPU-MuJoCo forward kinematics of ``MPCControllerWrapper`` (mpc_wrapper.py:313-318) is the
    ``mpxb_wb_foot_positions`` kernel; ``robot_mass = qM[0]`` (mpc_wrapper.py:45-46) is the model table's total mass.
There is no CPU fallback: construction raises without a CUDA device / libmpxb.so.
"""
from dataclasses import dataclass, replace as _dc_replace
from functools import partial

import numpy as np
import torch

import mpx_b200.utils.mpc_utils as mpc_utils
import mpx_b200.primal_dual_ilqr.optimizers as optimizers
from mpx_b200 import _lib
from mpx_b200.mjcf import load_model


@dataclass
class mpx_data:
    """reference mpc_wrapper.py:16-28 (batched: every tensor carries the leading env axis except W when shared)."""
    dt: float
    duty_factor: float
    step_freq: float
    step_height: float
    contact_time: torch.Tensor
    liftoff: torch.Tensor
    X0: torch.Tensor
    U0: torch.Tensor
    V0: torch.Tensor
    W: torch.Tensor
    grf: torch.Tensor

    def replace(self, **kw):
        return _dc_replace(self, **kw)


def _initial_state(config):
    parts = [config.p0, config.quat0, config.q0, np.zeros(6 + config.n_joints), config.p_legs0]
    if config.grf_as_state:
        parts.append(np.zeros(3 * config.n_contact))
    return np.concatenate([np.asarray(p, dtype=np.float64) for p in parts])


class _Base:
    def _setup(self, config, n_env, limited_memory, device, max_iter=1):
        if not torch.cuda.is_available():
            raise RuntimeError("mpx_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)
        self.config = config
        self.n_env = n_env
        self.model = load_model(config.model_path)
        self.robot_mass = float(self.model.total_mass)   # == qM[0] at qpos0 (mpc_wrapper.py:45-46)
        self.mpc_frequency = config.mpc_frequency
        self.shift = int(1 / (config.dt * config.mpc_frequency))
        self.q0 = np.array(config.q0, dtype=np.float64)
        self.initial_state = self._T(_initial_state(config))
        self.contact_id = [self.model.geom_id(nm) for nm in config.contact_frame]
        self.body_id = [self.model.body_id(nm) for nm in config.body_name]
        # cost / hessian / dynamics partials bound like the reference (mpc_wrapper.py:66-74, 237-244)
        self.cost = partial(config.cost, config.n_joints, config.n_contact, config.N)
        hessian_approx = partial(config.hessian_approx, config.n_joints, config.n_contact)
        self.dynamics = partial(config.dynamics, self.model, None, config.contact_frame, config.body_name,
                                config.n_joints, config.dt)
        self.work = partial(optimizers.mpc, self.cost, self.dynamics, hessian_approx, limited_memory)
        self._solver = optimizers.MpcSolver(config, n_env, device=self.device,
                                            lq_mode=optimizers.wrapper_lq_mode(limited_memory, n_env, config.n, config.N),
                                            robot_model=self.model, max_iter=max_iter)
        # limited_memory=False does not mean "associative scan" here: the wrapper binds the fastest LQ kernel for the
        # shape (optimizers.wrapper_lq_mode); this says which one
        self.lq_mode_name = self._solver.lq_mode_name
        self.u_ref = self._T(config.u_ref)
        self.W = self._T(config.W)

    def _gait(self, duty_factor, step_freq, step_height, clearence_speed, advance_first):
        key = (float(duty_factor), float(step_freq), float(step_height), float(clearence_speed), bool(advance_first))
        if getattr(self, "_gait_key", None) != key:
            self._gait_key = key
            self._gait_struct = _lib.make_gait(self.config, self.robot_mass, key[0], key[1], key[2], key[3],
                                               self.config.use_terrain_estimation, advance_first)
        return self._gait_struct

    def _as(self, t):
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t), dtype=torch.float32)
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True)

    def foot_positions(self, qpos):
        """geom_xpos of the contact spheres, qpos [E, 7+nj] -> [E, 3*nc] (mpxb_wb_foot_positions)."""
        qpos = self._as(qpos).contiguous()
        E = qpos.shape[0]
        feet = torch.empty(E, 3 * self.config.n_contact, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._solver.lib.mpxb_wb_foot_positions(self._solver.handle, stream, E, qpos.data_ptr(),
                                                           feet.data_ptr()), "mpxb_wb_foot_positions")
        return feet


class BatchedMPCControllerWrapper(_Base):
    def __init__(self, config, n_env, limited_memory=False, device=None):
        """reference mpc_wrapper.py:30-100."""
        self._setup(config, n_env, limited_memory, device)

    def reference_generator(self, *, duty_factor, step_freq, step_height, t_timer, x, foot, input, liftoff, contact=None):
        """The reference's ``self.reference_generator`` (mpc_wrapper.py:76-78: mpc_utils.reference_generator with
        use_terrain_estimation, N, dt, n_joints, n_contact, robot_mass, foot0, q0 and clearence_speed = 0.2 bound) for
        every environment, on the device (mpxb_reference_generator).  Keyword arguments as at the call site
        mpc_wrapper.py:137-147; returns (reference, parameter, liftoff) and leaves its arguments untouched."""
        x, foot = self._as(x).contiguous(), self._as(foot).contiguous()
        gait = self._gait(duty_factor, step_freq, step_height, 0.2, False)
        timer_in = self._as(t_timer).contiguous()
        liftoff = self._as(liftoff).clone()
        reference, parameter = self._solver.reference_generator(
            gait, x, foot, foot.shape[1], self._as(input).contiguous(), None if contact is None else self._as(contact).contiguous(),
            timer_in, torch.empty_like(timer_in), liftoff)
        return reference, parameter, liftoff

    def make_data(self):
        """reference mpc_wrapper.py:101-114, batched over n_env."""
        cfg, E = self.config, self.n_env
        return mpx_data(
            dt=cfg.dt, duty_factor=cfg.duty_factor, step_freq=cfg.step_freq, step_height=cfg.step_height,
            contact_time=self._T(cfg.timer_t).repeat(E, 1),
            liftoff=torch.zeros(E, 3 * cfg.n_contact, device=self.device),
            X0=self.initial_state.repeat(E, cfg.N + 1, 1).contiguous(),
            U0=self.u_ref.repeat(E, cfg.N, 1).contiguous(),
            V0=torch.zeros(E, cfg.N + 1, cfg.n, device=self.device),
            W=self.W,
            grf=torch.zeros(E, 3 * cfg.n_contact, device=self.device))

    def run(self, data, x0, input):
        """One MPC update for every environment (reference mpc_wrapper.py:116-170). Returns (data, tau)."""
        cfg = self.config
        nj, nc = cfg.n_joints, cfg.n_contact
        x0, input = self._as(x0), self._as(input)
        # gait timer + reference generator: one kernel (mpc_wrapper.py:132-147; the batched wrapper feeds the generator
        # the phase BEFORE this tick's timer advance, contact = 0 and clearence_speed = 0.2)
        x0 = x0.contiguous()
        gait = self._gait(data.duty_factor, data.step_freq, data.step_height, 0.2, False)
        contact_time = torch.empty_like(data.contact_time)
        liftoff = data.liftoff.clone()
        reference, parameter = self._solver.reference_generator(
            gait, x0, x0[:, 13 + 2 * nj:], x0.shape[1], input.contiguous(), None, data.contact_time.contiguous(),
            contact_time, liftoff)
        self.reference, self.parameter = reference, parameter
        X, U, V = self._solver.solve(reference, parameter, data.W, x0, data.X0, data.U0, data.V0)
        X0, U0, V0 = torch.empty_like(data.X0), torch.empty_like(data.U0), torch.empty_like(data.V0)
        X0.copy_(data.X0); U0.copy_(data.U0); V0.copy_(data.V0)
        tau, _, _ = self._solver.shift_extract(X, U, V, x0, self.u_ref, X0, U0, V0, self.shift)
        g0 = 13 + 2 * nj + 3 * nc
        data = data.replace(X0=X0, U0=U0, V0=V0, contact_time=contact_time, liftoff=liftoff,
                            grf=X0[:, 0, g0:g0 + 3 * nc])
        self.X, self.U, self.V = X, U, V
        return data, tau

    def reset(self, data, qpos, qvel, foot):
        """reference mpc_wrapper.py:172-192, batched: qpos [E,7+nj], qvel [E,6+nj], foot [E,3nc]."""
        cfg, E = self.config, self.n_env
        qpos, qvel, foot = self._as(qpos), self._as(qvel), self._as(foot)
        parts = [qpos, qvel, foot]
        if cfg.grf_as_state:
            parts.append(torch.zeros(E, 3 * cfg.n_contact, device=self.device))
        initial_state = torch.cat(parts, dim=1)
        return data.replace(
            U0=self.u_ref.repeat(E, cfg.N, 1).contiguous(),
            X0=initial_state[:, None, :].repeat(1, cfg.N + 1, 1).contiguous(),
            V0=torch.zeros(E, cfg.N + 1, cfg.n, device=self.device),
            contact_time=self._T(cfg.timer_t).repeat(E, 1), liftoff=foot.clone(),
            grf=torch.zeros(E, 3 * cfg.n_contact, device=self.device))


class MPCControllerWrapper(_Base):
    def __init__(self, config, limited_memory=False, device=None):
        """reference mpc_wrapper.py:195-300 (single instance, stateful; numpy in / numpy out)."""
        self._setup(config, 1, limited_memory, device)
        self.foot0 = np.array(config.p_legs0, dtype=np.float64)
        self._solve = self._solver.solve
        self.U0 = self.u_ref.repeat(1, config.N, 1).contiguous()
        self.X0 = self.initial_state.repeat(1, config.N + 1, 1).contiguous()
        self.V0 = torch.zeros(1, config.N + 1, config.n, device=self.device)
        self.contact_time = self._T(config.timer_t)[None]
        self.liftoff = self._T(config.p_legs0)[None].clone()
        self.tau = np.zeros(config.n_joints)
        self.q = np.zeros(config.n_joints)
        self.dq = np.zeros(config.n_joints)
        self.duty_factor = config.duty_factor
        self.step_freq = config.step_freq
        self.step_height = config.step_height
        self.robot_height = config.initial_height
        self.tau0 = np.zeros(config.n_joints)
        self.start_time = 0
        self.contact = np.zeros(config.n_contact)
        self.clearence_speed = 0.4
        self.p_collision = np.zeros(3 * config.n_contact)
        self.collision = [0, 0, 0, 0]
        self.collision_cycle = np.zeros(config.n_contact)

    def _x0(self, qpos, qvel):
        qpos_t = self._as(np.asarray(qpos, dtype=np.float32))[None]
        qvel_t = self._as(np.asarray(qvel, dtype=np.float32))[None]
        foot_op = self.foot_positions(qpos_t)
        parts = [qpos_t, qvel_t, foot_op]
        if self.config.grf_as_state:
            parts.append(torch.zeros(1, 3 * self.config.n_contact, device=self.device))
        return torch.cat(parts, dim=1), foot_op

    def run(self, qpos, qvel, input, contact):
        """reference mpc_wrapper.py:302-364: returns (tau, q, dq) as numpy arrays (blocks on the D2H read)."""
        cfg = self.config
        self.contact = np.array(contact).copy()
        x0, foot_op = self._x0(qpos, qvel)
        input = np.array(input, dtype=np.float64)
        input[6] = self.robot_height
        contact_t = self._as(np.asarray(contact, dtype=np.float32))[None]
        # mpc_wrapper.py:329-343: the timer advances first and the generator starts from the advanced phase
        gait = self._gait(self.duty_factor, self.step_freq, self.step_height, self.clearence_speed, True)
        reference, parameter = self._solver.reference_generator(
            gait, x0, foot_op, foot_op.shape[1], self._as(input)[None].contiguous(), contact_t.contiguous(),
            self.contact_time, self.contact_time, self.liftoff)
        self.reference, self.parameter = reference, parameter
        X, U, V = self._solve(reference, parameter, self.W, x0, self.X0, self.U0, self.V0)
        tau, q, dq = self._solver.shift_extract(X, U, V, x0, self.u_ref, self.X0, self.U0, self.V0, self.shift)
        self.X, self.U, self.V = X, U, V
        cmd = self._solver.cmd_pack[:, 0, :].cpu().numpy()     # (tau, q, dq): the one blocking D2H read of the tick
        tau = np.clip(cmd[0], cfg.min_torque, cfg.max_torque)
        return tau, cmd[1], cmd[2]

    def reset(self, qpos, qvel):
        """reference mpc_wrapper.py:366-385."""
        cfg = self.config
        x0, foot_op = self._x0(qpos, qvel)
        self.contact_time = self._T(cfg.timer_t)[None]
        self.liftoff = foot_op.clone()
        self.U0 = self.u_ref.repeat(1, cfg.N, 1).contiguous()
        self.X0 = x0[:, None, :].repeat(1, cfg.N + 1, 1).contiguous()
        self.V0 = torch.zeros(1, cfg.N + 1, cfg.n, device=self.device)
        print("MPC Controller Reset")
        return

    def runOffline(self, qpos, qvel, reference=None, parameter=None, max_iter=100, keep_iterates=True):
        """reference mpc_wrapper.py:387-468: solve the whole horizon offline from ``config.reference`` (the barrel roll,
        config_barrel_roll.py:91).  The SQP loop runs `mpc` until ``last_cost - g*g < 1e-3 and sum(c*c) < 1e-5`` (g, c at
        the NEW iterate) or until ``i > max_iter``, i.e. at most ``max_iter + 1`` iterations.

        keep_iterates=True (the reference's behaviour): a host loop like the reference's, one `mpxb_mpc_step` per
        iteration and one small read-back of the accepted trial point's (g, sum c^2) -- which the line search has
        already evaluated at the new iterate -- for the stop decision; returns ``(X, U, reference, output)`` with
        ``output`` = [X0, X_1, X_2, ...] exactly as the reference returns it (examples/barel_roll.py renders them).
        keep_iterates=False: the same loop on the device (`mpxb_mpc_step` with max_iter + 1 iterations, per-environment
        done flag, no host round trip); ``output`` is then [X0, X_final] and ``self.offline_iterations`` the count.
        ``reference`` / ``parameter`` default to ``config.reference(N + 1, dt, n_joints, n_contact, p_legs0, q0)``."""
        cfg = self.config
        x0, _ = self._x0(qpos, qvel)
        if reference is None:
            reference, parameter = cfg.reference(cfg.N + 1, cfg.dt, cfg.n_joints, cfg.n_contact, cfg.p_legs0, cfg.q0)
        reference = self._as(np.asarray(reference, dtype=np.float32))[None].contiguous()
        # the dynamics read parameter[t, :n_contact] only (models.py:72); the library takes exactly those columns
        parameter = self._as(np.asarray(parameter, dtype=np.float32)[:, :self._solver.npar])[None].contiguous()
        nq = 13 + cfg.n_joints
        self.X0 = self.X0.clone()
        self.X0[:, :, :nq] = reference[:, :, :nq]  # warm start (mpc_wrapper.py:420)
        output = [self.X0[0].clone()]
        if not keep_iterates:
            key = ("offline", int(max_iter))
            if getattr(self, "_offline_key", None) != key:
                self._offline_key = key
                self._offline_solver = optimizers.MpcSolver(cfg, 1, lq_mode=self._solver.lq_mode, device=self.device,
                                                            robot_model=self.model, max_iter=int(max_iter) + 1)
            s = self._offline_solver
            X, U, V = s.solve(reference, parameter, self.W, x0, self.X0, self.U0, self.V0)
            self.offline_iterations = int(s.iters[0])
            output.append(X[0].clone())
        else:
            s = self._solver
            last_cost, i = 1e10, 0
            while True:
                X, U, V = s.solve(reference, parameter, self.W, x0, self.X0, self.U0, self.V0)
                self.X0, self.U0, self.V0 = X.clone(), U.clone(), V.clone()
                output.append(self.X0[0].clone())
                # (g, sum c^2) of the new iterate = the accepted trial point of this step's line search
                ai = s.alpha_index[:1].long()
                gc = torch.stack([s.view("ls_phi")[0, ai[0]], s.view("ls_theta")[0, ai[0]]]).cpu().numpy()
                l2_cost, viol = float(gc[0]) * float(gc[0]), float(gc[1])
                i += 1
                if i > max_iter or (last_cost - l2_cost < 1e-3 and viol < 1e-5) or not np.isfinite(l2_cost):
                    break
                last_cost = l2_cost
            self.offline_iterations = i
        self.X0, self.U0, self.V0 = X.clone(), U.clone(), V.clone()
        return self.X0[0], self.U0[0], reference[0], output
