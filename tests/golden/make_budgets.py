"""FP32 error budgets of the golden fixtures -> tests/golden/budgets_<robot>.json (TEST INFRASTRUCTURE).

    python tests/golden/make_budgets.py [robot ...]

north_star asks for 1e-4 relative in FP32.  Several quantities of the whole-body solves cannot be computed to 1e-4 by
ANY float32 implementation (value functions of 5e6 next to multipliers of 1e2: dV = P dX + p is cancellation), so each
fixture quantity gets a budget: the error of the float32 ORACLE against the float64 oracle under the tests' own metric
(tests/parity.py).  One float32 run is one sample of rounding error, and an equivalent algorithm that merely sums in
another order lands elsewhere, so the budget is the worst of several equivalent float32 evaluations of the same
reference algorithm on the same inputs:

* the plain float32 oracle (oracle/batched.py, torch float32);
* the LQ solve on the same float32 linearisation with the state and control axes reversed (every dot product is
  summed in the opposite order), un-permuted afterwards;
* for the associative scan: the suffix products combined right-to-left AND as a balanced (Hillis-Steele) tree --
  lax.associative_scan evaluates a tree, the CUDA kernels evaluate either, the result is the same product in exact
  arithmetic;
* the whole float32 oracle (linearisation + LQ solve) on inputs that differ from the fixture's by ONE FLOAT32 ULP of
  relative noise (three draws): the iterate X (knots >= 1), U, V and the reference trajectory each move by at most
  2^-23 relative.  Any float32 caller already rounds its inputs this much, so whatever this moves in the
  outputs is accuracy no float32 implementation can claim (forward error = condition x unit roundoff).

The GPU tests accept  max(1e-4, 4 x budget).
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle.problems import Problem  # noqa: E402
from oracle.batched import BatchedProblem  # noqa: E402
from oracle import primal_dual_ilqr as pdi  # noqa: E402
import parity  # noqa: E402

LIN = ("A", "B", "Q", "R", "M", "q", "r", "c")
LQ = ("K", "k", "P", "p", "dX", "dU", "dV")
CASES = ("warm", "second", "cold", "far")


def lq32(det32, e, scan, order, reverse):
    """One float32 LQ solve of environment e on the float32 linearisation; optionally with reversed axes."""
    Q, R, M, q, r, A, B, c = (det32[k][e] for k in ("Q", "R", "M", "q", "r", "A", "B", "c"))
    if reverse:
        fx = lambda a, *dims: torch.flip(a, dims)
        Q, R, M = fx(Q, 1, 2), fx(R, 1, 2), fx(M, 1, 2)
        q, r, c = fx(q, 1), fx(r, 1), fx(c, 1)
        A, B = fx(A, 1, 2), fx(B, 1, 2)
    if scan:
        K, k, P, p = pdi.tvlqr_gpu(Q, q, R, r, M, A, B, c[1:], order=order)
        dX, dU = pdi.rollout_gpu(K, k, c[0], A, B, c[1:])
    else:
        K, k, P, p = pdi.tvlqr(Q, q, R, r, M, A, B, c[1:])
        dX, dU = pdi.rollout(K, k, c[0], A, B, c[1:])
    dV = pdi.dual_lqr(dX, P, p)
    if reverse:
        fx = lambda a, *dims: torch.flip(a, dims)
        K, k, P, p = fx(K, 1, 2), fx(k, 1), fx(P, 1, 2), fx(p, 1)
        dX, dU, dV = fx(dX, 1), fx(dU, 1), fx(dV, 1)
    return dict(K=K, k=k, P=P, p=p, dX=dX, dU=dU, dV=dV)


def budgets_of(robot, case):
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    g = np.load(os.path.join(HERE, f"{robot}_{case}.npz"))
    bp32 = BatchedProblem(Problem(cfg, dtype=torch.float32))
    segs = parity.segments(cfg)
    E = g["x0"].shape[0]
    args = (g["reference"], g["parameter"], g["x0"], g["X0"], g["U0"], g["V0"])
    out = {}
    put = lambda key, e, v: out.setdefault(key, [0.0] * E).__setitem__(e, max(out.setdefault(key, [0.0] * E)[e], float(v)))
    rel_arr = lambda a, b: float(np.abs(a - b).max() / max(1e-3, np.abs(b).max()))
    npy = lambda t: t.detach().double().numpy()
    t32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
    for scan, sfx in ((False, ""), (True, "_scan")):
        det32 = {}
        X32, U32, V32 = bp32.mpc(not scan, *args, detail=det32)
        for e in range(E):
            if not scan:
                for nm in LIN:
                    put(nm, e, parity.block_rel_err(nm, npy(det32[nm][e]), g[nm][e], segs,
                                                    magnitude=g["X0"][e] if nm == "c" else None))
            variants = [{nm: det32[nm][e] for nm in LQ}]
            variants.append(lq32(det32, e, scan, "sequential", True))
            if scan:
                variants.append(lq32(det32, e, scan, "tree", False))
            alpha = 2.0 ** (-int(g["alpha_index" + sfx][e]))
            for i, v in enumerate(variants):
                for nm in LQ:
                    put(nm + sfx, e, parity.block_rel_err(nm, npy(v[nm]), g[nm + sfx][e], segs))
                # the new iterate along that direction with the fixture's accepted step
                Xn = t32(g["X0"][e]) + alpha * v["dX"]
                Un = t32(g["U0"][e]) + alpha * v["dU"]
                Vn = t32(g["V0"][e]) + alpha * v["dV"]
                for nm, a in (("X", Xn), ("U", Un), ("V", Vn)):
                    put(nm + sfx, e, parity.block_rel_err(nm, npy(a), g[nm + sfx][e], segs))
            for nm, a in (("X", X32), ("U", U32), ("V", V32)):
                if int(det32["alpha_index"][e]) == int(g["alpha_index" + sfx][e]):
                    put(nm + sfx, e, parity.block_rel_err(nm, npy(a[e]), g[nm + sfx][e], segs))
            for nm in ("theta", "phi"):
                put(nm + sfx, e, rel_arr(npy(det32[nm][e]), g[nm + sfx][e]))
        # one float32 ulp of input noise, whole pipeline
        for draw in range(3):
            rng = np.random.default_rng(1000 + draw)
            jit = lambda a: np.asarray(a, dtype=np.float64) * (1.0 + 2.0 ** -23 * rng.uniform(-1, 1, np.shape(a)))
            detp = {}
            X0p = jit(g["X0"])
            X0p[:, 0] = g["X0"][:, 0]   # the measured state and knot 0 of the iterate stay as they are (c_0 = x0 - X_0 exactly)
            bp32.mpc(not scan, jit(g["reference"]), g["parameter"], g["x0"], X0p, jit(g["U0"]), jit(g["V0"]), detail=detp)
            for e in range(E):
                if not scan:
                    for nm in LIN:
                        put(nm, e, parity.block_rel_err(nm, npy(detp[nm][e]), g[nm][e], segs,
                                                        magnitude=g["X0"][e] if nm == "c" else None))
                alpha = 2.0 ** (-int(g["alpha_index" + sfx][e]))
                for nm in LQ:
                    put(nm + sfx, e, parity.block_rel_err(nm, npy(detp[nm][e]), g[nm + sfx][e], segs))
                for nm, base, d in (("X", "X0", "dX"), ("U", "U0", "dU"), ("V", "V0", "dV")):
                    a = t32(g[base][e]) + alpha * detp[d][e]
                    put(nm + sfx, e, parity.block_rel_err(nm, npy(a), g[nm + sfx][e], segs))
    return out


if __name__ == "__main__":
    robots = sys.argv[1:] or ["srbd", "go2", "aliengo", "h1", "talos"]
    for r in robots:
        OUT = os.path.join(HERE, f"budgets_{r}.json")
        table = {}
        for case in (("iter",) if r == "barrel_roll" else CASES):
            b = budgets_of(r, case)
            table[f"{r}_{case}"] = {k: [float(f"{x:.4e}") for x in v] for k, v in sorted(b.items())}
            big = {k: f"{max(v):.1e}" for k, v in b.items() if max(v) > 2.5e-5}
            print(r, case, big, flush=True)
            json.dump(table, open(OUT, "w"), indent=0, sort_keys=True)
