#!/usr/bin/env python
"""bench.py -- batched MPC solves/sec of the B200-native primal-dual iLQR path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # native arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (restated reference on host cores)

A "step" is one pass of the hot path over one batch of synthetic robot states: one `optimizers.mpc` call
(evaluate + linearise/quadratise + LQ solve + 11-alpha filter line search) for every environment, followed by
the warm-start shift, in closed loop on the model (x0 <- X[shift]).  Workload at N=1: Go2 whole-body MPC,
1024 environments per GPU (BASELINE.json configs[2]); with N>1 every rank owns its own 1024-environment slice
(weak scaling, no collective on the solve path, one final gather of the controls).
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
def _cpu_solve_worker(args):
    """One restated-reference `optimizers.mpc` call (float32 torch oracle) for one environment."""
    robot, seed = args
    import torch
    torch.set_num_threads(1)
    from oracle.problems import Problem, synthetic_inputs
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    key = ("pr", robot)
    if key not in _cpu_solve_worker.__dict__:
        _cpu_solve_worker.__dict__[key] = Problem(cfg, dtype=torch.float32)
    pr = _cpu_solve_worker.__dict__[key]
    inp = synthetic_inputs(pr, 1, seed=seed)
    t0 = time.perf_counter()
    pr.mpc(True, inp["reference"][0], inp["parameter"][0], inp["x0"][0], inp["X0"][0], inp["U0"][0], inp["V0"][0])
    return time.perf_counter() - t0


def cpu_arm(robot, steps, warmup, cores):
    """Times the oracle port on the host cores: each step = `cores` independent solves, one per process."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        seed = 1000
        for _ in range(warmup):
            pool.map(_cpu_solve_worker, [(robot, seed + i) for i in range(cores)])
            seed += cores
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_solve_worker, [(robot, seed + i) for i in range(cores)])
            seed += cores
        dt = time.perf_counter() - t0
    return cores * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    robot, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # bounded sample: the pure-Python/torch restatement needs ~10-20 s per whole-body solve per core, so the
    # number of timed steps is capped to keep the run within minutes (stated in `sample`)
    steps = max(1, min(args.steps, args.ref_max_steps))
    warmup = min(args.warmup, 1)
    value, ms = cpu_arm(robot, steps, warmup, cores)
    sample = (f"{cores} independent solves per step (one per process), {steps} timed steps after {warmup} warm-up; "
              f"float32 torch restatement of the reference (JAX/MJX not installable here)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "robot": robot, "envs_per_step": cores},
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
def run_native(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from mpx_b200 import _lib
    from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper as SrbdWrapper
    from mpx_b200.utils.sharding import gather_controls, max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py native arm needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION / INFO print to stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("MPXB_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    robot, desc = WORKLOADS[args.workload]
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    if args.horizon and args.horizon != cfg.N:
        # horizon sweep (BASELINE configs[3]): the same config with N replaced; every kernel takes N at run time
        from types import SimpleNamespace
        cfg = SimpleNamespace(**{k: getattr(cfg, k) for k in dir(cfg) if not k.startswith("__")})
        cfg.N = int(args.horizon)
    E, K, Wm = args.envs, args.steps, args.warmup
    srbd = robot == "srbd"
    lib = _lib.lib()
    st = synthetic_states(cfg, E, seed=rank)
    T = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
    N, n, m, nj, nc = cfg.N, cfg.n, cfg.m, cfg.n_joints, cfg.n_contact
    # limited_memory as the reference's examples pass it: True for the batched whole-body runs (multi_env.py:44), the
    # default False for the single-instance controllers and the SRBD wrapper; the wrappers then bind the LQ kernel
    # that is fastest for the shape (optimizers.wrapper_lq_mode); --scan binds the associative scan literally
    lm = args.workload in ("go2_wb", "h1_wb") and not args.scan
    if args.scan:
        if args.scan_schedule:
            os.environ["MPXB_SCAN_SCHEDULE"] = args.scan_schedule
        os.environ["MPXB_LQ"] = "scan"  # bind the associative-scan kernel literally (wrappers default to the sequential one)

    if srbd:
        wrap = SrbdWrapper(cfg, E, device=dev)
        solver = wrap._solver
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
        solver = wrap._solver
        qpos, qvel = T(st["qpos"]), T(st["qvel"])
        feet = wrap.foot_positions(qpos)
        data = wrap.reset(wrap.make_data(), qpos, qvel, feet)
        data = data.replace(contact_time=T(st["timer"]))
        x0 = data.X0[:, 0, :].clone()
    Wt = T(cfg.W)
    u_ref = T(cfg.u_ref)
    shift = int(1 / (cfg.dt * cfg.mpc_frequency))
    n_warm = max(Wm, 3)
    # ---- input tape: the real closed loop (gait timer + reference generator + solve + shift, x0 <- X[shift]) is run
    # once, untimed, for n_warm + K ticks and its per-tick solve inputs (reference, parameter, x0) are kept in HBM.
    # The timed loop replays the tape through the solve + shift kernels only; every kernel is deterministic, so the
    # replay reproduces the closed-loop iterates exactly.
    tape = []
    if srbd:
        for _ in range(n_warm + K):
            wrap.run(x0, T(st["cmd"]), feet, torch.ones(E, nc, device=dev))
            tape.append((wrap.reference, wrap.parameter, x0))
            x0 = wrap.X[:, shift, :].clone()
            feet = wrap.parameter[:, 0, 4:].clone()
        x_init = tape[0][2]
        X0 = wrap.initial_state.repeat(E, N + 1, 1).contiguous()
    else:
        cmd = T(st["cmd"])
        d = data
        for _ in range(n_warm + K):
            d, _ = wrap.run(d, x0, cmd)
            tape.append((wrap.reference, wrap.parameter, x0))
            x0 = wrap.X[:, shift, :].clone()
        x_init = tape[0][2]
        X0 = x_init[:, None, :].repeat(1, N + 1, 1).contiguous()
    U0 = u_ref.repeat(E, N, 1).contiguous()
    V0 = torch.zeros(E, N + 1, n, device=dev)
    X, U, V = torch.empty_like(X0), torch.empty_like(U0), torch.empty_like(V0)
    tape_pos = [0]

    def step():
        reference, parameter, x0c = tape[tape_pos[0]]
        tape_pos[0] += 1
        solver.solve(reference, parameter, Wt, x0c, X0, U0, V0, out=(X, U, V))
        solver.shift_extract(X, U, V, x0c, u_ref, X0, U0, V0, shift, want_cmd=False)
    launches_per_step = solver.launches_per_call + 1

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
        step()
    prof = C.c_void_p()
    _lib.check(lib.mpxb_profile_create(K, C.byref(prof)), "mpxb_profile_create")
    _lib.check(lib.mpxb_profile_attach(solver.handle, prof), "mpxb_profile_attach")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    kms = (C.c_double * 4)()
    ncalls = C.c_int()
    _lib.check(lib.mpxb_profile_read(prof, C.byref(kms), C.byref(ncalls)), "mpxb_profile_read")
    lib.mpxb_profile_attach(solver.handle, None)
    lib.mpxb_profile_destroy(prof)
    nonfinite_envs = int((~torch.isfinite(X).flatten(1).all(1)).sum().item())
    alpha_hist = torch.bincount(solver.alpha_index.to(torch.int64), minlength=11).tolist()
    ms_total = max_over_ranks(ms_total, device=dev)

    # ---- end to end through the wrapper API with HOST buffers (H2D of x0 / command, D2H of the torques every step)
    e2e = None
    if not srbd:
        x0_h = torch.empty(E, n, dtype=torch.float32).pin_memory()
        cmd_h = torch.as_tensor(st["cmd"]).pin_memory()
        tau_h = torch.empty(E, nj, dtype=torch.float32).pin_memory()
        x0_h.copy_(x_init.cpu())
        d2 = wrap.reset(wrap.make_data(), qpos, qvel, feet).replace(contact_time=T(st["timer"]))

        def tick(d):
            xd = x0_h.to(dev, non_blocking=True)
            cd = cmd_h.to(dev, non_blocking=True)
            d, tau = wrap.run(d, xd, cd)
            tau_h.copy_(tau, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # the caller reads tau
            x0_h.copy_(wrap.X[:, shift, :], non_blocking=True)   # simulator stand-in: next measured state (not counted)
            return d
        for _ in range(max(Wm, 3)):
            d2 = tick(d2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            d2 = tick(d2)
        e1.record()
        barrier()
        t_e2e = max_over_ranks(e0.elapsed_time(e1), device=dev)
        e2e = {"value": E * world * K / (t_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(E * (n + 7) * 4), "d2h_bytes_per_step": int(E * nj * 4),
               "ms_per_step": t_e2e / K,
               "api": "BatchedMPCControllerWrapper.run (gait timer + reference generator + solve + shift), pinned host buffers"}

    else:
        x0_h = torch.empty(E, n, dtype=torch.float32).pin_memory()
        x0_h.copy_(x_init.cpu())
        cmd_h = torch.as_tensor(st["cmd"]).pin_memory()
        feet_h = tape[0][1][:, 0, 4:].cpu().pin_memory()
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

    # final gather of the controls (the only collective; north_star), outside the timed solve path
    if world > 1:
        gather_controls(U[:, 0, :nj].contiguous(), n_env_total=E * world, dst=0)

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
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.workload, {}).get("lq_solve_dram_bytes_per_launch")
        except Exception:
            pass
        lq_s = kernel_ms["lq_solve"] * 1e-3
        lq_name = {_lib.LQ_SEQUENTIAL: "sequential", _lib.LQ_SCAN: "scan", _lib.LQ_SCAN_SERIAL: "scan-serial",
                   _lib.LQ_SCAN_PARALLEL: "scan-parallel"}[solver.lq_mode]
        if solver.lq_mode == _lib.LQ_SCAN:
            lq_name = "scan-parallel" if E <= 64 else "scan-serial"
        roofline = {"kernel": {"sequential": "lq_seq_kernel", "scan-serial": "lq_scan_kernel"}.get(lq_name, "pscan_init + pscan_combine x levels + pscan_gains + pscan_rollout"), "bound": "fp32",
                    "achieved": flq / lq_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": flq / lq_s / 1e12 / fp32_peak, "traffic": traffic,
                    "peak_source": "FP32 FFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                    "algorithmic": f"F_lqr(n={n},m={m},N={N}) = {f_lqr(n, m, N) / 1e6:.2f} MFLOP per solve x {E} solves per launch",
                    "hbm": {"achieved": blq / lq_s / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                            "frac": (blq / lq_s / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                            "algorithmic": f"{lq_bytes(n, m, N) / 1e6:.3f} MB per solve (LQ data read once + outputs written once)"},
                    "whole_step": {"achieved": flq / (ms_total / K * 1e-3) / 1e12, "unit": "TFLOP/s",
                                   "frac": flq / (ms_total / K * 1e-3) / 1e12 / fp32_peak,
                                   "note": "SURVEY.md 8(d): F_lqr x solves/s / FP32 peak over the whole tick (linearise + LQ + line search + shift)"},
                    "kernel_ms": kernel_ms, "dominant_kernel": dom,
                    "share_of_step": {k: v / (ms_total / K) for k, v in kernel_ms.items()}}
        latency = None
        if world == 1 and not args.no_latency:
            latency = latency_leg(dev, ticks=args.latency_ticks)
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, ms = cpu_arm(robot, 1, 0, cores)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{cores} solves of the same workload, one per process (float32 torch restatement of the reference)"}
        line = {"metric": METRIC, "value": E * world * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(Wm, 3), "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "robot": robot, "envs_per_gpu": E, "n": n, "m": m, "horizon": N,
                           "lq": lq_name, "limited_memory": lm, "sqp_iterations_per_step": 1,
                           "l2": "per-step working set (workspace %.0f MB) exceeds the 126 MB L2; no flush" % (solver.ws_bytes / 1e6),
                           "parallelism": f"env-sharded x{world}, no collective on the solve path"},
                "e2e": e2e, "gpu_launches": launches_per_step * K, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "latency": latency, "clocks": clocks, "nonfinite_envs": nonfinite_envs,
                "alpha_index_hist_last_step": alpha_hist}
        print(json.dumps(line), flush=True)
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
