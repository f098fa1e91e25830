"""Offline barrel-roll solve on the GPU against the float64 golden (tests/golden/barrel_roll_offline.npz):
    python scripts/barrel_roll_gpu.py            # per-iteration report for both limited_memory settings"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import mpx_b200.config.config_barrel_roll as cfg  # noqa: E402
from mpx_b200.utils.mpc_wrapper import MPCControllerWrapper  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "barrel_roll_offline.npz"))
segs = parity.segments(cfg)
for lm in (True, False):
    w = MPCControllerWrapper(cfg, limited_memory=lm)
    w.reset(g["qpos"], g["qvel"])
    t0 = time.perf_counter()
    X, U, ref, out = w.runOffline(g["qpos"], g["qvel"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"limited_memory={lm} lq={w._solver.lq_mode} iterations={w.offline_iterations} (oracle {int(g['iterations'])}) wall {dt:.2f}s "
          f"finite={bool(torch.isfinite(X).all())}")
    for i in range(min(6, len(out) - 1)):
        print("   iterate", i, "X err vs f64", parity.block_rel_err("X", out[i + 1].cpu().numpy(), g["first_X"][i], segs))
    print("   final X err", parity.block_rel_err("X", X.cpu().numpy(), g["X"], segs),
          "U err", parity.block_rel_err("U", U.cpu().numpy(), g["U"], segs))
    s = w._solver
    ai = int(s.alpha_index[0])
    print("   final g", float(s.view("ls_phi")[0, ai]), "(oracle", float(g["g_hist"][-1]), ") sum c^2", float(s.view("ls_theta")[0, ai]),
          "(oracle", float(g["theta_hist"][-1]), ")")
    t0 = time.perf_counter()
    w.reset(g["qpos"], g["qvel"])
    X2, U2, _, out2 = w.runOffline(g["qpos"], g["qvel"], keep_iterates=False)
    torch.cuda.synchronize()
    print(f"   device loop: iterations={w.offline_iterations} wall {time.perf_counter() - t0:.2f}s max|X - X_hostloop| {float((X2 - X).abs().max()):.3e}")
