"""SRBD batched MPC wrapper -- same class/constructor/``run``/``reset`` as reference
``mpx/utils/mpc_wrapper_srbd.py:19-175`` with the jitted JAX solve replaced by libmpxb.so.

State lives in torch CUDA tensors; ``run`` enqueues the gait timer, the reference generator, the solve and the
warm-start shift on the current CUDA stream and never synchronises.
"""
from functools import partial

import numpy as np
import torch

import mpx_b200.utils.mpc_utils as mpc_utils
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives
import mpx_b200.primal_dual_ilqr.optimizers as optimizers
from mpx_b200 import _lib


class BatchedMPCControllerWrapper:
    def __init__(self, config, n_env, device=None):
        """reference mpc_wrapper_srbd.py:20-96."""
        if not torch.cuda.is_available():
            raise RuntimeError("mpx_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)
        self._T = T
        self.n_env = n_env
        self.config = config
        self.mpc_frequency = config.mpc_frequency
        self.shift = int(1 / (config.dt * config.mpc_frequency))
        self.foot0 = np.array(config.p_legs0, dtype=np.float64)
        self.q0 = np.array(config.q0, dtype=np.float64)
        self.initial_state = T(np.concatenate([config.p0, config.quat0, np.zeros(6)]))

        # cost / hessian / dynamics partials, bound exactly like the reference (mpc_wrapper_srbd.py:60-68)
        cost = partial(mpc_objectives.quadruped_srbd_obj, config.n_contact, config.N)
        hessian_approx = partial(mpc_objectives.quadruped_srbd_hessian_gn, config.n_contact)
        inertia = np.asarray(config.inertia, dtype=np.float64)
        dynamics = partial(mpc_dyn_model.quadruped_srbd_dynamics, config.mass, inertia, np.linalg.inv(inertia), config.dt)
        self.work = partial(optimizers.mpc, cost, dynamics, hessian_approx, False)
        # the batch axis of the CUDA solver replaces jax.jit(jax.vmap(work)) (mpc_wrapper_srbd.py:77)
        self._solver = optimizers.MpcSolver(config, n_env, device=self.device,
                                            lq_mode=optimizers.wrapper_lq_mode(False, n_env, config.n, config.N))
        self._solve = self._solver.solve
        self.lq_mode_name = self._solver.lq_mode_name   # the LQ kernel bound for this shape (optimizers.wrapper_lq_mode)
        self._gait = _lib.make_gait(config, config.mass, config.duty_factor, config.step_freq, config.step_height,
                                    config.clearence_speed, config.use_terrain_estimator, True)
        self.W = T(config.W)
        self.u_ref = T(config.u_ref)
        self.reset(verbose=False)
        self.foot_ref = torch.zeros(n_env, 3 * config.n_contact, device=self.device)
        self.foot_ref_dot = torch.zeros(n_env, 3 * config.n_contact, device=self.device)
        self.grf = torch.zeros(n_env, 3 * config.n_contact, device=self.device)

    def run(self, x0, input, foot_op, contact):
        """One MPC update (mpc_wrapper_srbd.py:91-141). All arguments [n_env, ...] CUDA tensors."""
        x0, input, foot_op, contact = (self._as(t) for t in (x0, input, foot_op, contact))
        cfg = self.config
        x0, input, foot_op, contact = (t.contiguous() for t in (x0, input, foot_op, contact))
        # mpc_wrapper_srbd.py:103-121: timer advance, then the generator from the advanced phase -- one kernel
        reference, parameter, foot_ref_dot = self._solver.reference_generator(
            self._gait, x0, foot_op, foot_op.shape[1], input, contact, self.contact_time, self.contact_time, self.liftoff,
            want_foot_ref_dot=True)
        self.contact = (self.contact_time < cfg.duty_factor).to(torch.float32)
        self.reference, self.parameter = reference, parameter
        self.foot_ref = parameter[:, 0, 4:]
        self.foot_ref_dot = foot_ref_dot[:, 0, :]
        X, U, V = self._solve(reference, parameter, self.W, x0, self.batch_X0, self.batch_U0, self.batch_V0)
        # warm start for the next call (mpc_wrapper_srbd.py:135-137), fused shift kernel
        self._solver.shift_extract(X, U, V, x0, self.u_ref, self.batch_X0, self.batch_U0, self.batch_V0, self.shift)
        self.grf = U[:, 0, :]
        self.X, self.U, self.V = X, U, V
        return 0

    def graph_run(self, x0, input, foot_op, contact):
        """``run`` replayed from a CUDA graph (one launch per tick instead of one per kernel): the first call captures
        reference generator + solve + shift on static input buffers, later calls copy the arguments into them and replay.
        Same results and the same state updates as ``run``; ``reset`` drops the graph."""
        args = tuple(self._as(t).contiguous() for t in (x0, input, foot_op, contact))
        if getattr(self, "_graph", None) is None:
            self._graph_in = tuple(t.clone() for t in args)
            state = (self.contact_time, self.liftoff, self.batch_X0, self.batch_U0, self.batch_V0)
            saved = tuple(t.clone() for t in state)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.run(*self._graph_in)          # warm-up outside the capture (module load, kernel attributes)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self.run(*self._graph_in)
                for t, s in zip(state, saved):     # neither the warm-up nor the capture may advance the controller
                    t.copy_(s)
            cur.wait_stream(side)
            self._graph = graph
        else:
            for dst, src in zip(self._graph_in, args):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return 0

    def _as(self, t):
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t), dtype=torch.float32)
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True)

    def torch_run(self, x0_torch, input_torch, foot_op_torch, contact_torch=None):
        """mpc_wrapper_srbd.py:143-151 (the reference passes 3 arguments to a 4-argument run; contact defaults to 1)."""
        if contact_torch is None:
            contact_torch = torch.ones(self.n_env, self.config.n_contact, device=self.device)
        return self.run(x0_torch, input_torch, foot_op_torch, contact_torch)

    def _whole_body_model(self):
        """Whole-body (MJCF tree) handle of the same robot for whole_body_interface, created on first use."""
        if getattr(self, "_wb_handle", None) is None:
            import ctypes as C
            from types import SimpleNamespace
            cfg = self.config
            nj, nc = cfg.n_joints, cfg.n_contact
            wb = SimpleNamespace(model_path=cfg.model_path, contact_frame=cfg.contact_frame, body_name=cfg.body_name,
                                 n_joints=nj, n_contact=nc, n=13 + 2 * nj + 6 * nc, m=nj, N=cfg.N, dt=cfg.dt,
                                 W=np.zeros((12 + 3 * nj + 6 * nc,) * 2), dynamics=mpc_dyn_model.quadruped_wb_dynamics,
                                 cost=mpc_objectives.quadruped_wb_obj)
            self._wb_desc = _lib.make_desc(wb)
            handle = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().mpxb_model_create(C.byref(self._wb_desc), C.byref(handle)), "mpxb_model_create")
            self._wb_handle = handle
            self._Kp, self._Kd = self._T(cfg.Kp).contiguous(), self._T(cfg.Kd).contiguous()
        return self._wb_handle

    def whole_body_run(self, qpos, qvel, return_jacobian=False):
        """mpc_wrapper_srbd.py:153-154 -> mpc_utils.whole_body_interface (275-313) on the device: joint torques
        [n_env, n_joints] from the last solve's GRFs, foot references and contact schedule."""
        handle = self._whole_body_model()
        E, nj, nc = self.n_env, self.config.n_joints, self.config.n_contact
        qpos, qvel = self._as(qpos).contiguous(), self._as(qvel).contiguous()
        if tuple(qpos.shape) != (E, 7 + nj) or tuple(qvel.shape) != (E, 6 + nj):
            raise ValueError(f"qpos / qvel: expected {(E, 7 + nj)} / {(E, 6 + nj)}, got {tuple(qpos.shape)} / {tuple(qvel.shape)}")
        tau = torch.empty(E, nj, dtype=torch.float32, device=self.device)
        J = torch.empty(E, 6 + nj, 3 * nc, dtype=torch.float32, device=self.device) if return_jacobian else None
        grf, fr, frd = self.grf.contiguous(), self.foot_ref.contiguous(), self.foot_ref_dot.contiguous()
        contact = self._as(self.contact).contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().mpxb_whole_body_interface(
            handle, stream, E, self._Kp.data_ptr(), self._Kd.data_ptr(), qpos.data_ptr(), qvel.data_ptr(), grf.data_ptr(),
            fr.data_ptr(), frd.data_ptr(), contact.data_ptr(), tau.data_ptr(), J.data_ptr() if J is not None else None),
            "mpxb_whole_body_interface")
        return (tau, J) if return_jacobian else tau

    def torch_whole_body_run(self, qpos, qvel):
        """mpc_wrapper_srbd.py:156-160 (the tensors already are torch CUDA tensors here)."""
        return self.whole_body_run(qpos, qvel)

    def reset(self, verbose=True):
        """mpc_wrapper_srbd.py:162-174."""
        cfg, E = self.config, self.n_env
        self._graph = None
        self.contact_time = self._T(cfg.timer_t).repeat(E, 1)
        self.liftoff = torch.zeros(E, 3 * cfg.n_contact, device=self.device)
        self.contact = torch.ones(E, cfg.n_contact, device=self.device)
        self.batch_U0 = self.u_ref.repeat(E, cfg.N, 1).contiguous()
        self.batch_X0 = self.initial_state.repeat(E, cfg.N + 1, 1).contiguous()
        self.batch_V0 = torch.zeros(E, cfg.N + 1, cfg.n, device=self.device)
        if verbose:
            print("MPC Controller Reset")
        return
