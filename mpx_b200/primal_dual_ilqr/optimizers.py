"""``optimizers.mpc`` -- the reference's solver entry point, served by libmpxb.so.

Reference: mpx/primal_dual_ilqr/primal_dual_ilqr/optimizers.py:603-683.  The positional signature is kept::

    work = partial(optimizers.mpc, cost, dynamics, hessian_approx, limited_memory)
    X, U, V = work(reference, parameter, W, x0, X_in, U_in, V_in)

``cost`` / ``dynamics`` / ``hessian_approx`` are the descriptor partials the wrappers build exactly like the
reference builds its function partials (mpc_wrapper.py:66-74, mpc_wrapper_srbd.py:60-68); their bound arguments
carry the dimensions and model tables from which the CUDA model handle is created (once, cached).
Inputs are torch CUDA float32 tensors, either per-environment ([N+1, n] ...) or batched with a leading env
axis ([E, N+1, n] ...) -- the batch axis replaces the reference's ``jax.vmap(work)`` (mpc_wrapper_srbd.py:77,
examples/multi_env.py:46-49).  There is no CPU path: non-CUDA tensors raise.
"""
import ctypes as C
from functools import partial

import numpy as np
import os

import torch

from .. import _lib
from ..utils.models import DynamicsKind
from ..utils.objectives import ObjectiveKind


class _Cfg:
    """Minimal config namespace rebuilt from descriptor partials."""


def _unwrap(p):
    args = ()
    while isinstance(p, partial):
        args = p.args + args
        p = p.func
    return p, args


# measured on B200 at E = 1 (profiles/r01ap_launches_*_e1_pscan.csv, r01ah_talos_horizon_sweep.jsonl), microseconds:
# n -> (sequential kernel per knot, parallel-scan combine per level, parallel-scan fixed part, its rollout per knot)
_LQ_COST_US = {13: (8.9, 15.3, 24.5, 1.2), 61: (18.0, 92.6, 64.0, 2.7), 75: (30.5, 131.0, 89.0, 3.4),
               81: (79.0, 158.6, 230.0, 4.1)}


# the small-system sequential kernel (csrc/lq_small.cuh, SRBD): 71 us for N = 25 at E = 1 (r02w); it holds the whole
# horizon in shared memory, which for n = 13, m = 12 fits up to N = 47
_LQ_SMALL_KNOT_US, _LQ_SMALL_MAX_N = 2.9, 47


LQ_MODE_NAMES = {_lib.LQ_SEQUENTIAL: "sequential", _lib.LQ_SCAN: "scan", _lib.LQ_SCAN_SERIAL: "scan_serial",
                 _lib.LQ_SCAN_PARALLEL: "scan_parallel"}


def _sm_count():
    """SMs of the current CUDA device (one parallel-scan CTA per SM and wave); 148 (B200) when no device is visible."""
    try:
        if torch.cuda.is_available():
            return int(torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count)
    except Exception:
        pass
    return 148


def wrapper_lq_mode(limited_memory, n_env, n, N):
    """Which LQ kernel the controller wrappers bind (an ``_lib.LQ_*`` value).

    The reference picks the associative-scan Riccati (`tvlqr_gpu`) when ``limited_memory=False`` because on an XLA
    GPU the sequential recursion is launch-latency bound.  Here every form solves the same LQ problem to rounding,
    so for ``limited_memory=False`` the wrappers bind whichever is faster on B200 for the shape at hand:

    * ``limited_memory=True`` -> the sequential kernel, literally (examples/multi_env.py:44);
    * ``limited_memory=False`` -> the scan scheduled in parallel over the horizon (one CTA per knot, log2 levels) when
      its modelled time  fixed + levels x waves x combine  beats  N x knot  of the sequential kernel, with the
      per-size constants measured at E = 1 (``_LQ_COST_US``; waves = ceil(n_env (N + 1) / 148)): single instances of
      Talos (N = 25: LQ 1.04 vs 1.99 ms; N = 200: 3.7 vs 15.3 ms), long horizons in general (the SRBD model beyond
      N = 47; up to there its sequential solve runs in the small-system kernel, LQ 0.07 vs 0.13 ms); Aliengo / Go2 at N = 25 (0.56 vs 0.45 ms), H1 (0.77 vs 0.76 ms) and every large batch stay on the sequential
      kernel, where one CTA per environment already fills the GPU and the scan's ~9x FLOPs buy nothing.

    ``MPXB_LQ=scan`` in the environment restores the literal mapping (``LQ_SCAN``), ``MPXB_LQ=sequential`` forces the
    sequential kernel; ``MpcSolver`` itself always maps ``limited_memory`` literally (the parity tests exercise every
    kernel through it)."""
    import os
    forced = os.environ.get("MPXB_LQ", "").lower()
    if forced == "scan":
        return _lib.LQ_SEQUENTIAL if limited_memory else _lib.LQ_SCAN
    if forced == "sequential" or limited_memory or n_env > 64 or n not in _LQ_COST_US:
        return _lib.LQ_SEQUENTIAL
    knot, comb, fixed, roll = _LQ_COST_US[n]
    if n == 13 and N <= _LQ_SMALL_MAX_N:
        knot = _LQ_SMALL_KNOT_US  # whole horizon in shared memory (csrc/lq_small.cuh)
    levels = int(N).bit_length()
    waves = -(-n_env * (N + 1) // _sm_count())
    t_scan = fixed + roll * N + levels * waves * comb
    t_seq = N * knot
    return _lib.LQ_SCAN_PARALLEL if t_scan < t_seq else _lib.LQ_SEQUENTIAL


class MpcSolver:
    """One CUDA model handle + workspace for a fixed (config, n_env, lq_mode)."""

    def __init__(self, config, n_env, limited_memory=False, device=None, max_iter=1, robot_model=None,
                 eager_line_search=False, line_search="filter", lq_mode=None, cost_structure="auto",
                 stop_rule="offline"):
        if not torch.cuda.is_available():
            raise RuntimeError("mpx_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_env = int(n_env)
        self.lq_mode = (_lib.LQ_SEQUENTIAL if limited_memory else _lib.LQ_SCAN) if lq_mode is None else int(lq_mode)
        if self.lq_mode == _lib.LQ_SCAN:
            # MPXB_SCAN_SCHEDULE=serial|parallel forces a schedule of the associative scan (benchmarks / profiling)
            import os
            forced = {"serial": _lib.LQ_SCAN_SERIAL, "parallel": _lib.LQ_SCAN_PARALLEL}.get(
                os.environ.get("MPXB_SCAN_SCHEDULE", "").lower())
            if forced is not None:
                self.lq_mode = forced
        self.lq_mode_name = LQ_MODE_NAMES[self.lq_mode]   # which LQ kernel this solver runs (see wrapper_lq_mode)
        self.max_iter = int(max_iter)
        self.desc = _lib.make_desc(config, robot_model)
        self.N, self.n, self.m = self.desc.N, self.desc.n, self.desc.m
        self.nref, self.npar, self.nw = self.desc.nref, self.desc.npar, self.desc.nw
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mpxb_model_create(C.byref(self.desc), C.byref(handle)), "mpxb_model_create")
        self.handle = handle
        # lazy (default): alpha = 1 first, the ten smaller steps only where it is rejected -- same X, U, V, alpha_index;
        # eager: all 11 trial points for every environment (fills the ls_theta / ls_phi views completely)
        self.eager_line_search = bool(eager_line_search)
        _lib.check(self.lib.mpxb_model_set_option(self.handle, _lib.OPT_EAGER_LINE_SEARCH, int(self.eager_line_search)),
                   "mpxb_model_set_option")
        # "filter": parallel_filter_line_search (what the reference's mpc runs); "merit": parallel_line_search on the
        # augmented-Lagrangian merit function (optimizers.py:185-217, 295-354, 637-669)
        # "armijo": the legacy solver's sequential backtracking line_search (old_optimazers.py:173-251) on the same 11
        # step sizes.  stop_rule (max_iter > 1): "offline" = runOffline's rule (mpc_wrapper.py:456-467), "legacy" = the
        # legacy loop's continuation_criterion (old_optimazers.py:435-447); "legacy" needs "merit" or "armijo"
        modes = {"filter": _lib.LS_FILTER, "merit": _lib.LS_MERIT, "armijo": _lib.LS_ARMIJO}
        if line_search not in modes:
            raise ValueError(f"line_search must be 'filter', 'merit' or 'armijo', got {line_search!r}")
        if stop_rule not in ("offline", "legacy") or (stop_rule == "legacy" and line_search == "filter"):
            raise ValueError(f"stop_rule must be 'offline' or 'legacy' (with a merit line search), got {stop_rule!r}")
        self.line_search, self.stop_rule = line_search, stop_rule
        _lib.check(self.lib.mpxb_model_set_option(self.handle, _lib.OPT_LINE_SEARCH, modes[line_search]),
                   "mpxb_model_set_option")
        _lib.check(self.lib.mpxb_model_set_option(self.handle, _lib.OPT_STOP_RULE,
                                                  _lib.STOP_LEGACY if stop_rule == "legacy" else _lib.STOP_OFFLINE),
                   "mpxb_model_set_option")
        # cost_structure: "dense" reads the whole weight matrix W; "diagonal" promises a diagonal W and runs the
        # structured kernels (MPXB_OPT_COST_STRUCTURE, csrc/lq3.cuh; whole-body models with limited_memory=True);
        # "auto" (default) looks at W the first time a given tensor is passed to solve() and picks for itself
        if cost_structure not in ("auto", "dense", "diagonal"):
            raise ValueError(f"cost_structure must be 'auto', 'dense' or 'diagonal', got {cost_structure!r}")
        self.cost_structure = cost_structure
        self._w_seen = None
        self._structured = False
        self._set_structure(cost_structure == "diagonal")
        self._expanded = True
        self.ws_bytes = self.lib.mpxb_workspace_bytes(self.handle, self.n_env, self.lq_mode)
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        E = self.n_env
        self.alpha_index = torch.zeros(E, dtype=torch.int32, device=self.device)
        self.iters = torch.zeros(E, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(E, dtype=torch.int32, device=self.device)
        # kernels per SQP iteration: linearise, LQ, select + the trial kernels (eager: one; lazy: alpha = 1, then the rest;
        # the cooperative whole-body form adds its reduce kernel, and its second pass launches both forms -- see
        # wb_class_trial in csrc/wb_kernels.cuh)
        wb = self.desc.kind != 0
        # passes: alpha = 1, [alpha = 1/2 for large H1 / Talos batches,] the rest (mpxb_mpc_step); the cooperative form
        # (quadrupeds and H1 at any batch size, every robot for small batches) is kernel + reduce in the first pass and
        # warp form + cooperative form + reduce in the later ones
        npass = 3 if (wb and self.desc.nv > 18 and self.n_env >= 256) else 2
        coop = wb and (self.desc.nv <= 25 or self.n_env * (self.N + 1) <= 148 * 8)
        if self.eager_line_search:
            trial = 1
        elif coop:
            trial = 2 + 3 * (npass - 1)
        else:
            trial = npass
        # the parallel-in-time scan is elements + ceil(log2(N+1)) levels + gains + rollout; the other LQ kernels are one launch
        parallel_scan = self.lq_mode == _lib.LQ_SCAN_PARALLEL or (self.lq_mode == _lib.LQ_SCAN and self.n_env <= 64)
        lq = 3 + int(self.N).bit_length() if parallel_scan else 1
        self._launches_base = 2 + lq + trial

    @property
    def launches_per_call(self):
        """Kernel launches of one solve().  Batches above one wave of LQ CTAs whose remainder leaves the last wave partly
        empty run the structured LQ kernel and the first line-search pass as two slices each (csrc/mpxb.cu, `EA`):
        one more LQ launch, one more trial kernel + reduction."""
        extra = 0
        if getattr(self, "_structured", False) and not self.eager_line_search and self.lq_mode == _lib.LQ_SEQUENTIAL \
                and os.environ.get("MPXB_SPLIT_LQ", "1")[:1] != "0":
            resident = {18: 4, 25: 2}.get(int(self.desc.nv), 1)
            slots = _sm_count() * resident
            rem = self.n_env % slots
            if self.n_env > slots and rem > 0 and 10 * rem <= 9 * slots:
                extra = 3
        return (self._launches_base + extra) * self.max_iter

    def _set_structure(self, diagonal):
        self._structured = bool(diagonal)
        _lib.check(self.lib.mpxb_model_set_option(self.handle, _lib.OPT_COST_STRUCTURE,
                                                  _lib.COST_DIAGONAL_W if diagonal else _lib.COST_DENSE),
                   "mpxb_model_set_option")

    def _inspect_W(self, W):
        """cost_structure="auto": one look (a device reduction + one scalar read) per distinct W tensor."""
        key = (W.data_ptr(), W._version, tuple(W.shape))
        if key == self._w_seen:
            return
        self._w_seen = key
        Wm = W if W.dim() == 3 else W[None]
        off = Wm - torch.diag_embed(torch.diagonal(Wm, dim1=-2, dim2=-1))
        self._set_structure(not bool(torch.any(off != 0).item()))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mpxb_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _chk(self, t, shape, name):
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
        if t.device != self.device:
            raise RuntimeError(f"{name}: tensor on {t.device}, solver on {self.device}")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)}")
        return t

    def solve(self, reference, parameter, W, x0, X_in, U_in, V_in, out=None):
        """Batched call: every argument carries the leading env axis except W ([nw,nw] shared or [E,nw,nw])."""
        E, N, n, m = self.n_env, self.N, self.n, self.m
        reference = self._chk(reference, (E, N + 1, self.nref), "reference")
        parameter = self._chk(parameter, (E, N + 1, self.npar), "parameter")
        if W.dim() == 2:
            W = self._chk(W, (self.nw, self.nw), "W")
            wstride = 0
        else:
            W = self._chk(W, (E, self.nw, self.nw), "W")
            wstride = self.nw * self.nw
        if self.cost_structure == "auto" and self.desc.kind != _lib.KIND_SRBD and self.lq_mode == _lib.LQ_SEQUENTIAL:
            self._inspect_W(W)
        x0 = self._chk(x0, (E, n), "x0")
        X_in = self._chk(X_in, (E, N + 1, n), "X_in")
        U_in = self._chk(U_in, (E, N, m), "U_in")
        V_in = self._chk(V_in, (E, N + 1, n), "V_in")
        if out is None:
            X = torch.empty_like(X_in)
            U = torch.empty_like(U_in)
            V = torch.empty_like(V_in)
        else:
            X, U, V = out
        self._expanded = False
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpxb_mpc_step(
            self.handle, stream, E, self.lq_mode, self.max_iter,
            reference.data_ptr(), parameter.data_ptr(), W.data_ptr(), wstride,
            x0.data_ptr(), X_in.data_ptr(), U_in.data_ptr(), V_in.data_ptr(),
            X.data_ptr(), U.data_ptr(), V.data_ptr(),
            self.alpha_index.data_ptr(), self.iters.data_ptr(), self.status.data_ptr(),
            self.workspace.data_ptr(), self.ws_bytes)
        _lib.check(rc, "mpxb_mpc_step")
        return X, U, V

    def view(self, name):
        """Intermediate of the last call (A, B, Q, R, M, q, r, c, K, k, P, p, dX, dU, dV, ls_theta, ls_phi, g, cost, slope)."""
        if not self._expanded:
            # the structured path keeps the LQ data packed; write the dense arrays out for inspection
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.mpxb_workspace_expand(self.handle, stream, self.n_env, self.lq_mode,
                                                      self.workspace.data_ptr()), "mpxb_workspace_expand")
            self._expanded = True
        cnt = C.c_size_t()
        ptr = self.lib.mpxb_workspace_view(self.handle, self.n_env, self.lq_mode, self.workspace.data_ptr(),
                                           name.encode(), C.byref(cnt))
        if not ptr:
            raise KeyError(name)
        off = ptr - self.workspace.data_ptr()
        flat = self.workspace[off:off + 4 * cnt.value].view(torch.float32)
        E, N, n, m = self.n_env, self.N, self.n, self.m
        shapes = {"A": (E, N, n, n), "B": (E, N, n, m), "Q": (E, N + 1, n, n), "R": (E, N, m, m), "M": (E, N, n, m),
                  "q": (E, N + 1, n), "r": (E, N, m), "c": (E, N + 1, n), "K": (E, N, m, n), "k": (E, N, m),
                  "P": (E, N + 1, n, n), "p": (E, N + 1, n), "dX": (E, N + 1, n), "dU": (E, N, m),
                  "dV": (E, N + 1, n), "ls_theta": (E, 11), "ls_phi": (E, 11), "ls_vc": (E, 11), "g": (E,), "cost": (E, N + 1),
                  "slope": (E,)}
        return flat.view(shapes[name])

    def reference_generator(self, gait, x, foot, foot_stride, input, contact, timer_in, timer_out, liftoff,
                            want_foot_ref_dot=False):
        """mpxb_reference_generator: gait timer + reference generator on the device (mpc_utils.py:7-14, 38-270).

        ``gait`` is a ``_lib.Gait``; ``foot`` may be a strided view into ``x`` (pass its row stride).  ``liftoff`` is
        updated in place, ``timer_out`` receives the advanced gait phase.  Returns (reference, parameter[, foot_ref_dot])."""
        E, N = self.n_env, self.N
        nc = self.desc.n_contact
        mk = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        reference, parameter = mk(E, N + 1, self.nref), mk(E, N + 1, self.npar)
        frd = mk(E, N + 1, 3 * nc) if want_foot_ref_dot else None
        p = lambda t: t.data_ptr() if t is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpxb_reference_generator(self.handle, stream, E, C.byref(gait), p(x), p(foot), int(foot_stride),
                                               p(input), p(contact), p(timer_in), p(timer_out), p(liftoff),
                                               p(reference), p(parameter), p(frd))
        _lib.check(rc, "mpxb_reference_generator")
        return (reference, parameter, frd) if want_foot_ref_dot else (reference, parameter)

    def shift_extract(self, X, U, V, x0, u_ref, X0, U0, V0, shift, want_cmd=True):
        """update_and_extract_helper (mpc_wrapper.py:79-98): in-place update of the warm start (X0,U0,V0)."""
        E, nj = self.n_env, self.desc.n_joints
        wb = self.desc.kind != _lib.KIND_SRBD
        tau = q = dq = None
        if wb and want_cmd:
            # one [3, E, nj] buffer so that a caller that wants all three on the host pays a single D2H copy
            self.cmd_pack = torch.empty(3, E, nj, dtype=torch.float32, device=self.device)
            tau, q, dq = self.cmd_pack[0], self.cmd_pack[1], self.cmd_pack[2]
        p = lambda t: t.data_ptr() if t is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mpxb_shift_extract(self.handle, stream, E, int(shift), p(X), p(U), p(V), p(x0), p(u_ref),
                                         p(X0), p(U0), p(V0), p(tau), p(q), p(dq))
        _lib.check(rc, "mpxb_shift_extract")
        return tau, q, dq


_SOLVERS = {}      # insertion-ordered: least recently used first
_SOLVERS_MAX = 8


def _config_from_partials(cost, dynamics, hessian_approx):
    """Rebuild the numeric config from the bound arguments of the descriptor partials."""
    dyn, dargs = _unwrap(dynamics)
    obj, cargs = _unwrap(cost)
    if not isinstance(dyn, DynamicsKind) or not isinstance(obj, ObjectiveKind):
        raise TypeError("mpc() expects the descriptor partials built by the mpx_b200 wrappers "
                        "(arbitrary Python cost/dynamics callables cannot run on the CUDA path)")
    cfg = _Cfg()
    cfg.dynamics, cfg.cost, cfg.hessian_approx = dyn, obj, hessian_approx
    if dyn.kind == _lib.KIND_SRBD:
        # partial(quadruped_srbd_dynamics, mass, inertia, inertia_inv, dt); partial(srbd_obj, n_contact, N)
        cfg.mass, cfg.inertia, _, cfg.dt = dargs[:4]
        cfg.n_contact, cfg.N = cargs[-2:]
        cfg.n_joints = 3 * cfg.n_contact
        cfg.n, cfg.m = 13, 3 * cfg.n_contact
        cfg.W = np.zeros((12 + 3 * cfg.n_contact,) * 2)
        robot_model = None
    else:
        # partial(dynamics, model, mjx_model, contact_id, body_id, n_joints, dt); partial(cost, [flag], nj, nc, N)
        robot_model, _, contact_names, body_names, cfg.n_joints, cfg.dt = dargs[:6]
        cfg.n_joints, cfg.n_contact, cfg.N = cargs[-3:]
        cfg.contact_frame, cfg.body_name = contact_names, body_names
        if len(cargs) > 3:
            cfg.cost = partial(obj, *cargs[:-3])
        nj, nc = cfg.n_joints, cfg.n_contact
        grf_in_u = dyn.kind == _lib.KIND_TALOS_WB
        cfg.n = 13 + 2 * nj + (3 * nc if grf_in_u else 6 * nc)
        cfg.m = nj + (3 * nc if grf_in_u else 0)
        cfg.W = np.zeros((12 + 3 * nj + 6 * nc,) * 2)
    return cfg, robot_model


def mpc(cost, dynamics, hessian_approx, limited_mempory, reference, parameter, W, x0, X_in, U_in, V_in):
    """Drop-in for reference optimizers.mpc (optimizers.py:603-683); see the module docstring."""
    batched = reference.dim() == 3
    if not batched:
        reference, parameter, x0, X_in, U_in, V_in = (t.unsqueeze(0) for t in (reference, parameter, x0, X_in, U_in, V_in))
    E = reference.shape[0]
    key = (id(_unwrap(cost)[0]), _unwrap(cost)[1][-3:], id(_unwrap(dynamics)[0]), id(_unwrap(dynamics)[1][0]),
           bool(limited_mempory), E, reference.device.index)
    entry = _SOLVERS.pop(key, None)
    if entry is None:
        cfg, rm = _config_from_partials(cost, dynamics, hessian_approx)
        solver = MpcSolver(cfg, E, bool(limited_mempory), device=reference.device, robot_model=rm)
        # the key holds id()s: keep the keyed objects alive beside the solver so that an id cannot be reused for another
        # model while the entry exists
        entry = (solver, (_unwrap(cost)[0], _unwrap(dynamics)[0], _unwrap(dynamics)[1][0]))
    _SOLVERS[key] = entry            # most recently used last
    while len(_SOLVERS) > _SOLVERS_MAX:
        _SOLVERS.pop(next(iter(_SOLVERS)))   # evict the least recently used handle + workspace
    solver = entry[0]
    X, U, V = solver.solve(reference, parameter, W, x0, X_in, U_in, V_in)
    if not batched:
        X, U, V = X[0], U[0], V[0]
    return X, U, V


def tvlqr_solve(Q, q, R, r, M, A, B, c, limited_memory=True, lq_mode=None):
    """Batched LQ sub-problem on the GPU: tvlqr+rollout+dual_lqr (limited_memory=True) or the scan forms.

    Reference: compute_search_direction's LQ half (optimizers.py:175-182).  All inputs [E, ...] CUDA float32;
    ``c`` is [E, N+1, n] (c[:,0] = rollout start, c[:,1:] = offsets).  Returns K, k, P, p, dX, dU, dV, slope.
    ``lq_mode`` overrides ``limited_memory`` with one of ``_lib.LQ_*`` (LQ_SCAN_SERIAL / LQ_SCAN_PARALLEL force a
    schedule of the associative scan; LQ_SCAN picks the parallel-in-time one for E <= 64).
    """
    lib = _lib.lib()
    for t in (Q, q, R, r, M, A, B, c):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("tvlqr_solve expects contiguous float32 CUDA tensors")
    E, N, n, m = B.shape
    dev = Q.device
    mk = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    K, k, P, p = mk(E, N, m, n), mk(E, N, m), mk(E, N + 1, n, n), mk(E, N + 1, n)
    dX, dU, dV, slope = mk(E, N + 1, n), mk(E, N, m), mk(E, N + 1, n), mk(E)
    mode = (_lib.LQ_SEQUENTIAL if limited_memory else _lib.LQ_SCAN) if lq_mode is None else int(lq_mode)
    scan = None
    if mode != _lib.LQ_SEQUENTIAL:
        scan = torch.empty(max(lib.mpxb_tvlqr_scan_bytes(E, N, n, m), 4), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.mpxb_tvlqr(stream, E, N, n, m, mode, Q.data_ptr(), q.data_ptr(), R.data_ptr(), r.data_ptr(),
                        M.data_ptr(), A.data_ptr(), B.data_ptr(), c.data_ptr(), K.data_ptr(), k.data_ptr(),
                        P.data_ptr(), p.data_ptr(), dX.data_ptr(), dU.data_ptr(), dV.data_ptr(), slope.data_ptr(),
                        scan.data_ptr() if scan is not None else None)
    _lib.check(rc, "mpxb_tvlqr")
    return K, k, P, p, dX, dU, dV, slope
