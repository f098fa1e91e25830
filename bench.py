#!/usr/bin/env python
"""bench.py -- batched MPC solves/sec of the B200-native primal-dual iLQR path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # native arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (restated reference on host cores)

A "step" is one pass of the hot path over one batch of synthetic robot states: one `optimizers.mpc` call
(evaluate + linearise/quadratise + LQ solve + 11-alpha filter line search) for every environment, followed by
the warm-start shift, in closed loop on the model (x0 <- X[shift]).  Workload at N=1: Go2 whole-body MPC,
1024 environments per GPU (BASELINE.json configs[2]); with N>1 every rank owns its own 1024-environment slice
(weak scaling, no collective on the solve path; the end-to-end tick ends with the NCCL gather of the controls).
The same line carries, under "strong", the north-star split: Go2 x4096 and H1 x4096 (configs[4]) sharded as
4096 / N environments per GPU.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched MPC solves/sec"
UNIT = "solves/s"
WORKLOADS = {
    "go2_wb": ("go2", "Go2 whole-body batched MPC (examples/multi_env.py with config_go2), sequential LQ (limited_memory=True)"),
    "aliengo_wb": ("aliengo", "Aliengo whole-body MPC (examples/mjx_quad.py)"),
    "h1_wb": ("h1", "H1 whole-body batched MPC"),
    "talos_wb": ("talos", "Talos whole-body MPC (examples/mjx_talos.py)"),
    "srbd": ("srbd", "Aliengo SRBD MPC (examples/srbd_quad.py)"),
}


def f_lqr(n, m, N):
    """Algorithmic FLOPs of one LQ solve: sequential Riccati + rollout + dual, FMA = 2."""
    return (N * (4 * n ** 3 + 6 * m * n ** 2 + 2 * m ** 2 * n + m ** 3 / 3 + 2 * m ** 2 * (n + 1) + 4 * n ** 2 + 6 * m * n)
            + N * (2 * n ** 2 + 4 * m * n) + (N + 1) * 2 * n ** 2)


def lq_bytes(n, m, N):
    """Algorithmic HBM bytes of the LQ kernel per solve: read A,B,Q,R,M,q,r,c once, write K,k,P,p,dX,dU,dV once."""
    rd = N * (n * n + n * m + m * m + n * m + m) + (N + 1) * (n * n + 2 * n)
    wr = N * (m * n + m + m) + (N + 1) * (n * n + 3 * n)
    return 4 * (rd + wr)


def synthetic_states(cfg, E, seed):
    """SURVEY.md 8(d) synthetic robot states + commands (numpy float32)."""
    rng = np.random.default_rng(seed)
    nj = cfg.n_joints
    dz = rng.uniform(-0.02, 0.02, E)
    rpy = rng.uniform(-0.1, 0.1, (E, 3))
    cr, sr = np.cos(rpy[:, 0] / 2), np.sin(rpy[:, 0] / 2)
    cp, sp = np.cos(rpy[:, 1] / 2), np.sin(rpy[:, 1] / 2)
    cy, sy = np.cos(rpy[:, 2] / 2), np.sin(rpy[:, 2] / 2)
    quat = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 1)
    p = np.tile(np.asarray(cfg.p0, dtype=np.float64), (E, 1))
    p[:, 2] += dz
    q = np.asarray(cfg.q0, dtype=np.float64)[None] + rng.uniform(-0.05, 0.05, (E, nj))
    v = rng.uniform(-0.2, 0.2, (E, 3))
    w = rng.uniform(-0.2, 0.2, (E, 3))
    dq = rng.uniform(-0.5, 0.5, (E, nj))
    cmd = np.stack([rng.uniform(-0.2, 0.5, E), rng.uniform(-0.1, 0.1, E), np.zeros(E), np.zeros(E), np.zeros(E),
                    rng.uniform(-0.3, 0.3, E), np.full(E, cfg.robot_height)], 1)
    phase = rng.uniform(0, 1, (E, 1))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(qpos=f32(np.concatenate([p, quat, q], 1)), qvel=f32(np.concatenate([v, w, dq], 1)), cmd=f32(cmd),
                timer=f32(np.mod(np.asarray(cfg.timer_t)[None] + phase, 1.0)), yaw=rpy[:, 2])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if "Active" in val and "Not" not in val:
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_arm(robot, steps, warmup, envs, seed=1000):
    """Times the restated reference on the host cores: each step = ONE batched `optimizers.mpc` call over `envs`
    environments, float32, evaluated by oracle/batched.py the way the reference evaluates it -- vmap over knots inside
    vmap over environments (examples/multi_env.py:46-49), one process, every core through torch's intra-op threads.
    Returns (solves/s, ms per step, threads used)."""
    import torch
    from oracle.batched import BatchedProblem
    from oracle.problems import Problem, synthetic_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    bp = BatchedProblem(Problem(cfg, dtype=torch.float32))
    inp = synthetic_inputs(bp.pr, envs, seed=seed)
    args = (inp["reference"], inp["parameter"], inp["x0"])
    X, U, V = inp["X0"], inp["U0"], inp["V0"]
    lm = robot != "srbd"
    for _ in range(warmup):
        bp.mpc(lm, *args, X, U, V)
    t0 = time.perf_counter()
    for _ in range(steps):
        # closed loop on the same batch: the next call starts from this call's iterate (no shift: the CPU arm times
        # the solve, which is >99.9 % of a tick)
        X, U, V = (t.numpy() for t in bp.mpc(lm, *args, X, U, V))
    dt = time.perf_counter() - t0
    return envs * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    robot, desc = WORKLOADS[args.workload]
    envs = args.ref_envs
    # bounded sample: the timed steps are capped so that the run ends within minutes (stated in `sample`)
    steps = max(1, min(args.steps, args.ref_max_steps))
    warmup = min(args.warmup, 1)
    value, ms, cores = cpu_arm(robot, steps, warmup, envs)
    sample = (f"{envs} environments per step in one batched call (the native arm runs {args.envs} per GPU), {steps} timed "
              f"steps after {warmup} warm-up; float32 torch restatement of the reference vectorised over environments and "
              f"knots like jax.vmap (JAX / MJX are not installable in this image)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "robot": robot, "envs_per_step": envs},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- latency leg
def _pct(a, q):
    return float(np.percentile(np.asarray(a, dtype=np.float64), q))


def latency_leg(dev, ticks=1000, warm=20):
    """p50 / p99 single-solve latency (BASELINE.json metric, second half): E = 1, `ticks` consecutive
    closed-loop ticks through the public wrapper call with HOST buffers, host wall clock around
    H2D + launch + D2H + sync, and CUDA-event time of the same region.

    * Aliengo SRBD (BASELINE configs[0], north-star target p50 < 1 ms): mpc_wrapper_srbd.BatchedMPCControllerWrapper
      `.run` (one launch per kernel) and `.graph_run` (the same tick replayed from a CUDA graph);
    * Aliengo whole-body (configs[1]): MPCControllerWrapper.run, numpy in / numpy out like the reference."""
    import torch
    from mpx_b200.utils.mpc_wrapper import MPCControllerWrapper
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper as SrbdWrapper
    out = {"ticks": ticks, "timing": "host wall clock around H2D + wrapper call + D2H + stream sync; device = CUDA events around the same region"}
    stream = torch.cuda.current_stream(dev)

    # ---- Aliengo SRBD, E = 1
    cfg = importlib.import_module("mpx_b200.config.config_srbd")
    nc = cfg.n_contact
    st = synthetic_states(cfg, 1, seed=123)
    yaw = float(st["yaw"][0])
    Ry = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    base = np.array([st["qpos"][0, 0], st["qpos"][0, 1], 0.0])
    feet0 = (base[None] + np.asarray(cfg.p_legs0).reshape(nc, 3) @ Ry.T).reshape(1, -1).astype(np.float32)
    x00 = np.concatenate([st["qpos"][:, :7], st["qvel"][:, :6]], 1).astype(np.float32)
    for mode in ("eager", "graph"):
        wrap = SrbdWrapper(cfg, 1, device=dev)
        shift = wrap.shift
        wrap.contact_time.copy_(torch.as_tensor(st["timer"]))
        wrap.liftoff.copy_(torch.as_tensor(feet0))
        x0_h, cmd_h = torch.as_tensor(x00).pin_memory(), torch.as_tensor(st["cmd"]).pin_memory()
        feet_h, con_h = torch.as_tensor(feet0).pin_memory(), torch.ones(1, nc).pin_memory()
        grf_h = torch.empty(1, 3 * nc).pin_memory()
        x0_d, cmd_d, feet_d, con_d = (torch.empty_like(t, device=dev) for t in (x0_h, cmd_h, feet_h, con_h))
        call = wrap.run if mode == "eager" else wrap.graph_run
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall, devt = [], []
        for i in range(warm + ticks):
            t0 = time.perf_counter()
            e0.record()
            x0_d.copy_(x0_h, non_blocking=True); cmd_d.copy_(cmd_h, non_blocking=True)
            feet_d.copy_(feet_h, non_blocking=True); con_d.copy_(con_h, non_blocking=True)
            call(x0_d, cmd_d, feet_d, con_d)
            grf_h.copy_(wrap.grf, non_blocking=True)
            e1.record()
            stream.synchronize()
            t1 = time.perf_counter()
            if i >= warm:
                wall.append((t1 - t0) * 1e3); devt.append(e0.elapsed_time(e1))
            # simulator stand-in (not timed): the next measured state is the model's own prediction
            x0_h.copy_(wrap.X[:, shift, :]); feet_h.copy_(wrap.parameter[:, 0, 4:])
        ok = bool(torch.isfinite(wrap.X).all().item())
        out["srbd_" + mode] = {"workload": "Aliengo SRBD MPC, single instance (mpc_wrapper_srbd %s)" % ("run" if mode == "eager" else "graph_run, CUDA-graph replay"),
                               "p50_ms": _pct(wall, 50), "p99_ms": _pct(wall, 99), "p50_device_ms": _pct(devt, 50),
                               "p99_device_ms": _pct(devt, 99), "finite": ok}
        del wrap

    # ---- Aliengo whole-body, E = 1, numpy API
    cfg = importlib.import_module("mpx_b200.config.config_aliengo")
    nj, ncw = cfg.n_joints, cfg.n_contact
    st = synthetic_states(cfg, 1, seed=124)
    wrap = MPCControllerWrapper(cfg, limited_memory=False, device=dev)
    qpos, qvel = st["qpos"][0].astype(np.float64), st["qvel"][0].astype(np.float64)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        wrap.reset(qpos, qvel)
    cmd, contact = st["cmd"][0].astype(np.float64), np.ones(ncw)
    wall = []
    for i in range(warm + ticks):
        t0 = time.perf_counter()
        tau, q, dq = wrap.run(qpos, qvel, cmd, contact)
        t1 = time.perf_counter()
        if i >= warm:
            wall.append((t1 - t0) * 1e3)
        x1 = wrap.X[0, wrap.shift].cpu().numpy().astype(np.float64)
        qpos, qvel = x1[:7 + nj], x1[7 + nj:13 + 2 * nj]
    out["aliengo_wb"] = {"workload": "Aliengo whole-body MPC, single instance (MPCControllerWrapper.run, numpy in / numpy out)",
                         "p50_ms": _pct(wall, 50), "p99_ms": _pct(wall, 99), "finite": bool(np.isfinite(tau).all())}
    best = min(out["srbd_eager"], out["srbd_graph"], key=lambda d: d["p50_ms"])
    out["p50_ms"], out["p99_ms"], out["workload"] = best["p50_ms"], best["p99_ms"], best["workload"]
    return out


# --------------------------------------------------------------------------------------------- native arm
LQ_NAMES = {0: "sequential", 1: "scan", 2: "scan-serial", 3: "scan-parallel"}


class Loop:
    """One closed-loop workload on one GPU: wrapper, pre-generated input tape resident in HBM, and the timed step."""

    def __init__(self, args, robot, E, K, n_warm, rank, dev, lm, horizon=0):
        import torch
        from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper
        from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper as SrbdWrapper
        cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
        if horizon and horizon != cfg.N:
            # horizon sweep (BASELINE configs[3]): the same config with N replaced; every kernel takes N at run time
            from types import SimpleNamespace
            cfg = SimpleNamespace(**{k: getattr(cfg, k) for k in dir(cfg) if not k.startswith("__")})
            cfg.N = int(horizon)
        self.cfg, self.E, self.K, self.dev, self.robot = cfg, E, K, dev, robot
        self.srbd = srbd = robot == "srbd"
        st = self.st = synthetic_states(cfg, E, seed=rank)
        T = self.T = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
        N, n, m, nj, nc = cfg.N, cfg.n, cfg.m, cfg.n_joints, cfg.n_contact
        if srbd:
            wrap = SrbdWrapper(cfg, E, device=dev)
            yaw = st["yaw"]
            Ry = np.stack([np.stack([np.cos(yaw), -np.sin(yaw), 0 * yaw], 1), np.stack([np.sin(yaw), np.cos(yaw), 0 * yaw], 1),
                           np.stack([0 * yaw, 0 * yaw, 1 + 0 * yaw], 1)], 1)
            base = np.concatenate([st["qpos"][:, :2], np.zeros((E, 1))], 1)
            feet = T((base[:, None, :] + np.einsum("eij,fj->efi", Ry, np.asarray(cfg.p_legs0).reshape(nc, 3))).reshape(E, -1))
            x0 = T(np.concatenate([st["qpos"][:, :7], st["qvel"][:, :6]], 1))
            wrap.contact_time = T(st["timer"])
            wrap.liftoff = feet.clone()
        else:
            wrap = BatchedMPCControllerWrapper(cfg, E, limited_memory=lm, device=dev)
            self.qpos, self.qvel = T(st["qpos"]), T(st["qvel"])
            feet = wrap.foot_positions(self.qpos)
            data = wrap.reset(wrap.make_data(), self.qpos, self.qvel, feet)
            data = data.replace(contact_time=T(st["timer"]))
            x0 = data.X0[:, 0, :].clone()
        self.wrap, self.solver, self.feet = wrap, wrap._solver, feet
        self.Wt, self.u_ref = T(cfg.W), T(cfg.u_ref)
        self.shift = shift = int(1 / (cfg.dt * cfg.mpc_frequency))
        # ---- input tape: the real closed loop (gait timer + reference generator + solve + shift, x0 <- X[shift]) is
        # run once, untimed, for n_warm + K ticks and its per-tick solve inputs (reference, parameter, x0) are kept in
        # HBM.  The timed loop replays the tape through the solve + shift kernels only; every kernel is deterministic,
        # so the replay reproduces the closed-loop iterates exactly.
        tape = []
        if srbd:
            for _ in range(n_warm + K):
                wrap.run(x0, T(st["cmd"]), feet, torch.ones(E, nc, device=dev))
                tape.append((wrap.reference, wrap.parameter, x0))
                x0 = self._advance(x0, wrap.X[:, shift, :])
                feet = wrap.parameter[:, 0, 4:].clone()
            self.x_init = tape[0][2]
            X0 = wrap.initial_state.repeat(E, N + 1, 1).contiguous()
        else:
            cmd = T(st["cmd"])
            d = data
            for _ in range(n_warm + K):
                d, _ = wrap.run(d, x0, cmd)
                tape.append((wrap.reference, wrap.parameter, x0))
                x0 = self._advance(x0, wrap.X[:, shift, :])
            self.x_init = tape[0][2]
            X0 = self.x_init[:, None, :].repeat(1, N + 1, 1).contiguous()
        self.tape, self.pos = tape, 0
        self.X0, self.U0 = X0, self.u_ref.repeat(E, N, 1).contiguous()
        self.V0 = torch.zeros(E, N + 1, n, device=dev)
        self.X, self.U, self.V = torch.empty_like(self.X0), torch.empty_like(self.U0), torch.empty_like(self.V0)
        self.launches_per_step = self.solver.launches_per_call + 1

    @staticmethod
    def _advance(x0, x_next):
        """Closed loop on the model: x0 <- X[shift].  An environment whose solve came back non-finite (FP32 loses the
        positive definiteness of R + B'PB on a badly conditioned iterate -- the float32 oracle fails on the same inputs,
        scripts/check_nonfinite_oracle.py; the reference reports it in-band as NaN and its wrapper falls back to the
        initial guess, mpc_wrapper.py:96) keeps its previous state, as a robot whose controller skipped a tick would."""
        import torch
        ok = torch.isfinite(x_next).all(1, keepdim=True)
        return torch.where(ok, x_next, x0).clone()

    def step(self):
        reference, parameter, x0c = self.tape[self.pos]
        self.pos += 1
        self.solver.solve(reference, parameter, self.Wt, x0c, self.X0, self.U0, self.V0, out=(self.X, self.U, self.V))
        self.solver.shift_extract(self.X, self.U, self.V, x0c, self.u_ref, self.X0, self.U0, self.V0, self.shift, want_cmd=False)


def run_native(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from mpx_b200 import _lib
    from mpx_b200.utils.sharding import gather_controls, max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py native arm needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    # stdout carries the one JSON line and nothing else: everything any library prints to file descriptor 1 during the
    # run (NCCL's version banner and its INFO log: communicator ranks, transports, NVLS) is sent to stderr, where the
    # driver can read it; the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        # the GPU boxes export NCCL_DEBUG=VERSION (only the banner is printed): raise it to INFO so the communicator
        # lines ("... rank r nranks N ...", transports, NVLS) reach stderr; a caller's INFO / TRACE setting is kept
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
    robot, desc = WORKLOADS[args.workload]
    E, K, Wm = args.envs, args.steps, args.warmup
    srbd = robot == "srbd"
    lib = _lib.lib()
    # limited_memory as the reference's examples pass it: True for the batched whole-body runs (multi_env.py:44), the
    # default False for the single-instance controllers and the SRBD wrapper; the wrappers then bind the LQ kernel
    # that is fastest for the shape (optimizers.wrapper_lq_mode); --scan binds the associative scan literally
    lm = args.workload in ("go2_wb", "h1_wb") and not args.scan
    if args.scan:
        if args.scan_schedule:
            os.environ["MPXB_SCAN_SCHEDULE"] = args.scan_schedule
        os.environ["MPXB_LQ"] = "scan"  # bind the associative-scan kernel literally (wrappers default to the sequential one)
    n_warm = max(Wm, 3)
    loop = Loop(args, robot, E, K, n_warm, rank, dev, lm, args.horizon)
    cfg, solver, wrap, st, T = loop.cfg, loop.solver, loop.wrap, loop.st, loop.T
    N, n, m, nj, nc = cfg.N, cfg.n, cfg.m, cfg.n_joints, cfg.n_contact
    shift = loop.shift

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # bring the device out of its idle power state before anything is timed (a fresh box runs the first ~2 s of
    # kernels measurably slower); this is the FP32 peak micro-benchmark the roofline uses, run for a fixed time
    t_pre = time.perf_counter()
    fp32_peak = 0.0
    while time.perf_counter() - t_pre < args.prewarm_s:
        fp32_peak = max(fp32_peak, float(lib.mpxb_fp32_peak_tflops(None, 2048)))
    for _ in range(n_warm):
        loop.step()
    prof = C.c_void_p()
    _lib.check(lib.mpxb_profile_create(K, C.byref(prof)), "mpxb_profile_create")
    _lib.check(lib.mpxb_profile_attach(solver.handle, prof), "mpxb_profile_attach")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        loop.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    kms = (C.c_double * 4)()
    ncalls = C.c_int()
    _lib.check(lib.mpxb_profile_read(prof, C.byref(kms), C.byref(ncalls)), "mpxb_profile_read")
    lib.mpxb_profile_attach(solver.handle, None)
    lib.mpxb_profile_destroy(prof)
    nonfinite_envs = int((~torch.isfinite(loop.X).flatten(1).all(1)).sum().item())
    alpha_hist = torch.bincount(solver.alpha_index.to(torch.int64), minlength=11).tolist()
    ms_total = max_over_ranks(ms_total, device=dev)

    # ---- end to end through the wrapper API with HOST buffers: H2D of x0 / command, the whole tick, the NCCL gather of
    # the controls onto rank 0 (N > 1; north_star: "only a final NCCL gather of controls") and the D2H read of the
    # torques, every step
    gather_ms = None
    if not srbd:
        x0_h = torch.empty(E, n, dtype=torch.float32).pin_memory()
        cmd_h = torch.as_tensor(st["cmd"]).pin_memory()
        tau_h = torch.empty(E * (world if rank == 0 else 1), nj, dtype=torch.float32).pin_memory()
        x0_h.copy_(loop.x_init.cpu())
        d2 = wrap.reset(wrap.make_data(), loop.qpos, loop.qvel, loop.feet).replace(contact_time=T(st["timer"]))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather_acc = []

        def tick(d, timed=False):
            xd = x0_h.to(dev, non_blocking=True)
            cd = cmd_h.to(dev, non_blocking=True)
            d, tau = wrap.run(d, xd, cd)
            if world > 1:
                if timed:
                    g0.record()
                tau_all = gather_controls(tau.contiguous(), n_env_total=E * world, dst=0)   # same stream as the solve
                if timed:
                    g1.record()
                if rank == 0:
                    tau_h.copy_(tau_all, non_blocking=True)
            else:
                tau_h.copy_(tau, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # the caller reads tau
            if timed and world > 1:
                gather_acc.append(g0.elapsed_time(g1))
            x0_h.copy_(wrap.X[:, shift, :], non_blocking=True)   # simulator stand-in: next measured state (not counted)
            return d
        for _ in range(max(Wm, 3)):
            d2 = tick(d2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            d2 = tick(d2, True)
        e1.record()
        barrier()
        t_e2e = max_over_ranks(e0.elapsed_time(e1), device=dev)
        if gather_acc:
            gather_ms = max_over_ranks(float(np.mean(gather_acc)), device=dev)
        e2e = {"value": E * world * K / (t_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(E * (n + 7) * 4), "d2h_bytes_per_step": int(E * (world if world > 1 else 1) * nj * 4),
               "ms_per_step": t_e2e / K, "gather_ms_per_step": gather_ms,
               "api": "BatchedMPCControllerWrapper.run (gait timer + reference generator + solve + shift)"
                      + (", NCCL gather of tau [E/N, nj] onto rank 0 on the solve stream" if world > 1 else "")
                      + ", pinned host buffers"}
    else:
        x0_h = torch.empty(E, n, dtype=torch.float32).pin_memory()
        x0_h.copy_(loop.x_init.cpu())
        cmd_h = torch.as_tensor(st["cmd"]).pin_memory()
        feet_h = loop.tape[0][1][:, 0, 4:].cpu().pin_memory()
        con_h = torch.ones(E, nc).pin_memory()
        grf_h = torch.empty(E, 3 * nc, dtype=torch.float32).pin_memory()
        wrap.reset(verbose=False)
        wrap.contact_time = T(st["timer"])
        wrap.liftoff = feet_h.to(dev)

        def tick_srbd():
            wrap.run(x0_h.to(dev, non_blocking=True), cmd_h.to(dev, non_blocking=True), feet_h.to(dev, non_blocking=True),
                     con_h.to(dev, non_blocking=True))
            grf_h.copy_(wrap.grf, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # the caller reads the GRFs
            x0_h.copy_(wrap.X[:, shift, :], non_blocking=True)   # simulator stand-in (not counted)
            feet_h.copy_(wrap.parameter[:, 0, 4:], non_blocking=True)
        for _ in range(max(Wm, 3)):
            tick_srbd()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            tick_srbd()
        e1.record()
        barrier()
        t_e2e = max_over_ranks(e0.elapsed_time(e1), device=dev)
        e2e = {"value": E * world * K / (t_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(E * (n + 7 + 3 * nc + nc) * 4), "d2h_bytes_per_step": int(E * 3 * nc * 4),
               "ms_per_step": t_e2e / K,
               "api": "mpc_wrapper_srbd.BatchedMPCControllerWrapper.run (gait timer + reference generator + solve + shift), pinned host buffers"}

    # ---- strong scaling (north_star; BASELINE configs[4]): a FIXED batch of `strong_envs` environments sharded over
    # the ranks, 4096 / N per GPU, for Go2 and H1; device-timed like the main line, max over ranks
    strong = None
    if not args.no_strong and args.workload == "go2_wb" and not args.scan and args.strong_envs % world == 0:
        strong = {"total_envs": args.strong_envs, "envs_per_gpu": args.strong_envs // world, "steps": max(4, K // 2)}
        del loop.tape
        for rb in ("go2", "h1"):
            Ks = strong["steps"]
            lp = Loop(args, rb, args.strong_envs // world, Ks, 3, rank, dev, True)
            for _ in range(3):
                lp.step()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(Ks):
                lp.step()
            s1.record()
            barrier()
            ms = max_over_ranks(s0.elapsed_time(s1), device=dev) / Ks
            bad = torch.tensor([float((~torch.isfinite(lp.X).flatten(1).all(1)).sum().item())], device=dev)
            if world > 1:
                dist.all_reduce(bad)
            strong[rb] = {"ms_per_step": ms, "value": args.strong_envs / (ms * 1e-3), "unit": UNIT,
                          "nonfinite_envs": int(bad.item())}
            del lp
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp32_peak = max(fp32_peak, float(lib.mpxb_fp32_peak_tflops(None, 2048)))
        names = ["linearize", "lq_solve", "trial", "select"]
        kernel_ms = {nm: float(kms[i]) for i, nm in enumerate(names)}
        dom = max(kernel_ms, key=kernel_ms.get)
        flq = f_lqr(n, m, N) * E
        blq = lq_bytes(n, m, N) * E
        step_s = ms_total / K * 1e-3
        lq_s = kernel_ms["lq_solve"] * 1e-3
        lq_name = LQ_NAMES[solver.lq_mode]
        if solver.lq_mode == _lib.LQ_SCAN:
            lq_name = "scan-parallel" if E <= 64 else "scan-serial"
        structured = bool(getattr(solver, "_structured", False)) and lq_name == "sequential"
        # sequential, dense: the small-system kernel holds the whole horizon in shared memory (n, m <= 16, N <= 47 for SRBD,
        # up to 6 waves of environments; csrc/mpxb.cu launch_lq_seq), the streaming kernel otherwise
        seq_kernel = "lq_small_kernel" if (n <= 16 and m <= 16 and N <= 47 and E <= 6 * 148) else "lq_seq_kernel"
        lq_kernel = "lq3_kernel" if structured else {"sequential": seq_kernel, "scan-serial": "lq_scan_kernel"}.get(
            lq_name, "pscan_init + pscan_combine x levels + pscan_gains + pscan_rollout")
        # DRAM traffic per launch of the two largest kernels from the newest ncu --set full capture of this command
        # (profiles/traffic.json names the capture); the run itself cannot read DRAM counters
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
            traffic = tr.get(dom + "_dram_bytes_per_launch")
            traffic_src = tr.get("capture")
        except Exception:
            tr = {}
        # achieved = F_lqr x solves/s over the WHOLE step (linearise + LQ + line search + shift) against
        # the FP32 FFMA peak measured in this run; the LQ kernel's own fraction and every kernel's share sit beside it
        roofline = {"bound": "fp32", "achieved": flq / step_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": flq / step_s / 1e12 / fp32_peak, "traffic": traffic,
                    "scope": "whole step: F_lqr x solves per step / step time (SURVEY.md 8d)",
                    "traffic_of": dom, "traffic_source": traffic_src,
                    "peak_source": "FP32 FFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                    "algorithmic": f"F_lqr(n={n},m={m},N={N}) = {f_lqr(n, m, N) / 1e6:.2f} MFLOP per solve x {E} solves per step",
                    "dominant_kernel": dom, "kernel_ms": kernel_ms,
                    "share_of_step": {k: v / (ms_total / K) for k, v in kernel_ms.items()},
                    "lq_kernel": {"kernel": lq_kernel, "achieved": flq / lq_s / 1e12, "frac": flq / lq_s / 1e12 / fp32_peak,
                                  "unit": "TFLOP/s", "traffic": tr.get("lq_solve_dram_bytes_per_launch"),
                                  "hbm": {"achieved": blq / lq_s / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                          "frac": (blq / lq_s / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                          "algorithmic": f"{lq_bytes(n, m, N) / 1e6:.3f} MB per solve (dense LQ data read once + outputs written once)"}}}
        latency = None
        if world == 1 and not args.no_latency:
            latency = latency_leg(dev, ticks=args.latency_ticks)
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = cpu_arm(robot, 2, 1, args.ref_envs)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{args.ref_envs} environments of the same workload per batched call, 2 timed calls after 1 warm-up "
                                      "(float32 torch restatement of the reference, vmap over environments and knots, all host cores)"}
        line = {"metric": METRIC, "value": E * world * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(Wm, 3), "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "robot": robot, "envs_per_gpu": E, "n": n, "m": m, "horizon": N,
                           "lq": lq_name, "limited_memory": lm, "cost_structure": "diagonal W (structured kernels)" if structured else "dense W",
                           "sqp_iterations_per_step": 1,
                           "l2": "per-step working set (workspace %.0f MB) exceeds the 126 MB L2; no flush" % (solver.ws_bytes / 1e6),
                           "parallelism": f"env-sharded x{world}, no collective on the solve path"},
                "e2e": e2e, "gpu_launches": loop.launches_per_step * K, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "latency": latency, "clocks": clocks, "nonfinite_envs": nonfinite_envs,
                "alpha_index_hist_last_step": alpha_hist, "strong": strong}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="go2_wb", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=1024, help="environments per GPU")
    ap.add_argument("--scan", action="store_true", help="associative-scan LQ (limited_memory=False)")
    ap.add_argument("--scan-schedule", default="", choices=["", "serial", "parallel"],
                    help="force the schedule of the associative scan (default: parallel over the horizon for E <= 64)")
    ap.add_argument("--horizon", type=int, default=0, help="override config.N (horizon sweep, e.g. Talos 25/50/100/200)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the E = 1 p50 / p99 latency leg")
    ap.add_argument("--latency-ticks", type=int, default=1000)
    ap.add_argument("--prewarm-s", type=float, default=2.0, help="seconds of untimed device pre-warm before the warm-up steps")
    ap.add_argument("--ref-max-steps", type=int, default=4)
    ap.add_argument("--ref-envs", type=int, default=64, help="environments per step of the CPU arm / cpu_baseline")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (Go2 / H1 x4096 over the ranks)")
    ap.add_argument("--strong-envs", type=int, default=4096)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
