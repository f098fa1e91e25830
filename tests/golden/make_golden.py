"""Generate the golden fixtures tests/golden/<robot>_<case>.npz from the CPU oracle (TEST INFRASTRUCTURE).

    python tests/golden/make_golden.py [robot ...]

The reference (JAX + MuJoCo-MJX) cannot be imported in this environment, so the vectors come from the float64
oracle (oracle/, a restatement of the reference's arithmetic -- see oracle/__init__.py: PARITY UNPINNED).  Each
fixture holds the solve's inputs, every intermediate of one `optimizers.mpc` call (A, B, Q, R, M, q, r, c, g, K, k,
P, p, dX, dU, dV, line-search theta/phi, accepted alpha index) and its outputs (X, U, V), plus `budget_<name>`:
the float32 oracle's own relative error against float64 for that quantity, which the GPU tests use as the FP32
error budget (tolerance = max(1e-4, 4 x budget)).  Large arrays are stored as float32.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.problems import Problem, synthetic_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("A", "B", "Q", "R", "M", "q", "r", "c", "K", "k", "P", "p", "dX", "dU", "dV")
CASES = {"srbd": 3, "go2": 2, "aliengo": 1, "h1": 1, "talos": 1}   # environments per fixture
SEED = 21


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(1.0, np.abs(b).max()))


def make(robot, case):
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    pr, pr32 = Problem(cfg), Problem(cfg, dtype=torch.float32)
    E = CASES[robot]
    inp = synthetic_inputs(pr, E, seed=SEED, cold_start=(case == "cold"))
    if case == "second":
        # a generic iterate: the oracle's own first SQP step, so X, U and V are all non-trivial
        for e in range(E):
            X, U, V = pr.mpc(True, inp["reference"][e], inp["parameter"][e], inp["x0"][e], inp["X0"][e],
                             inp["U0"][e], inp["V0"][e])
            inp["X0"][e], inp["U0"][e], inp["V0"][e] = X.numpy(), U.numpy(), V.numpy()
    out = {k: v for k, v in inp.items()}
    acc = {}
    for e in range(E):
        args = (inp["reference"][e], inp["parameter"][e], inp["x0"][e], inp["X0"][e], inp["U0"][e], inp["V0"][e])
        det, det32, dets = {}, {}, {}
        Xo, Uo, Vo = pr.mpc(True, *args, detail=det)
        X32, U32, V32 = pr32.mpc(True, *args, detail=det32)
        Xs, Us, Vs = pr.mpc(False, *args, detail=dets)   # associative-scan LQ (same search direction in exact arithmetic)
        row = {nm: det[nm].numpy() for nm in NAMES}
        row.update(g=np.float64(det["g"]), theta=det["theta"].numpy(), phi=det["phi"].numpy(),
                   theta_k=np.float64(det["theta_k"]), phi_k=np.float64(det["phi_k"]), slope=np.float64(det["slope"]),
                   alpha_index=np.int32(det["alpha_index"]), alpha_index_scan=np.int32(dets["alpha_index"]),
                   X=Xo.numpy(), U=Uo.numpy(), V=Vo.numpy(), X_scan=Xs.numpy(), U_scan=Us.numpy(), V_scan=Vs.numpy(),
                   K_scan=dets["K"].numpy(), P_scan=dets["P"].numpy(), dX_scan=dets["dX"].numpy(),
                   dV_scan=dets["dV"].numpy())
        bud = {nm: rel(det32[nm].numpy(), det[nm].numpy()) for nm in NAMES}
        bud.update(theta=rel(det32["theta"].numpy(), det["theta"].numpy()), phi=rel(det32["phi"].numpy(), det["phi"].numpy()),
                   X=rel(X32.numpy(), Xo.numpy()), U=rel(U32.numpy(), Uo.numpy()), V=rel(V32.numpy(), Vo.numpy()))
        for k, v in row.items():
            acc.setdefault(k, []).append(v)
        for k, v in bud.items():
            acc.setdefault("budget_" + k, []).append(v)
        print(f"{robot}/{case} env {e}: alpha_index={det['alpha_index']} (scan {dets['alpha_index']}) "
              f"f32 budgets: " + " ".join(f"{k}={v:.1e}" for k, v in bud.items()), flush=True)
    for k, v in acc.items():
        a = np.stack(v)
        if a.dtype == np.float64 and a.size > 4096:
            a = a.astype(np.float32)
        out[k] = a
    path = os.path.join(HERE, f"{robot}_{case}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    robots = sys.argv[1:] or list(CASES)
    for r in robots:
        for case in (("warm", "second", "cold") if r in ("srbd", "go2") else ("warm", "second")):
            make(r, case)
