"""Generate the golden fixtures tests/golden/<robot>_<case>.npz from the CPU oracle (TEST INFRASTRUCTURE).

    python tests/golden/make_golden.py [robot ...]          # srbd go2 aliengo h1 talos

The reference (JAX + MuJoCo-MJX) cannot be imported in this environment, so the vectors come from the float64
oracle (oracle/, a restatement of the reference's arithmetic -- see oracle/__init__.py: PARITY UNPINNED), evaluated
batched over (environment, knot) by oracle/batched.py (checked against the per-knot loop form in
tests/test_oracle_batched.py).

Each fixture holds E = 4 environments: the solve's inputs, every intermediate of one `optimizers.mpc` call with the
sequential LQ solver (A, B, Q, R, M, q, r, c, g, K, k, P, p, dX, dU, dV, line-search theta/phi, accepted alpha
index) and its outputs (X, U, V); the same call with the associative-scan LQ solver (`*_scan`: K, k, P, p, dX, dU,
dV, theta, phi, alpha_index, X, U, V -- the linearisation is shared); and `budget_<name>[_scan]`: the float32
oracle's own per-knot / per-block relative error (tests/parity.py) against float64 for that quantity, which the GPU
tests use as the FP32 error budget (tolerance = max(1e-4, 4 x budget)).

Cases: `warm` (the reset warm start X0 = tile(x0), U0 = u_ref, V0 = 0 at a random state), `second` (the oracle's own
first iterate, so X, U, V are all non-trivial) and `cold` (exact identity quaternion / zero velocity: the safe-norm
zero-Jacobian branch) and `far` (a random state drawn from 6x the nominal ranges: joint speeds up to 3 rad/s, base
tilt up to 0.6 rad -- the line search has to cut the step).  The environments of a fixture are picked from 16 seeded candidates so that the accepted
step sizes differ and at least one environment accepts alpha <= 1/4 (alpha index >= 2) wherever a candidate does,
and so that the float32 and float64 oracles agree on the accepted index (no near-tie in the filter test).
Large arrays are stored as float32.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle.problems import Problem, synthetic_inputs  # noqa: E402
from oracle.batched import BatchedProblem  # noqa: E402
import parity  # noqa: E402

NAMES = ("A", "B", "Q", "R", "M", "q", "r", "c", "K", "k", "P", "p", "dX", "dU", "dV")
SCAN_NAMES = ("K", "k", "P", "p", "dX", "dU", "dV")
E_FIX, E_CAND = 4, 16
SEED = 21
CASE_NAMES = tuple(os.environ.get("GOLDEN_CASES", "warm,second,cold,far").split(","))


def pick(alpha64, alpha32, alpha64s, alpha32s, k):
    """k candidate indices: one per distinct accepted index (largest first), then in order; only candidates on which
    the float32 and float64 oracles agree in both LQ modes."""
    ok = [i for i in range(len(alpha64)) if alpha64[i] == alpha32[i] and alpha64s[i] == alpha32s[i]]
    chosen = []
    for a in sorted({alpha64[i] for i in ok}, reverse=True):
        chosen.append(next(i for i in ok if alpha64[i] == a))
    for i in ok:
        if len(chosen) >= k:
            break
        if i not in chosen:
            chosen.append(i)
    return sorted(chosen[:k])


def barrel_roll_candidates():
    """Iterates of the offline barrel-roll solve (tests/golden/barrel_roll_offline.npz, made by
    make_golden_barrel_roll.py; N = 100): the start (accepted step 1/2) and the iterate after 6 SQP iterations (accepted
    step 1/16).  Two environments keep the fixture at the size of the others."""
    g = np.load(os.path.join(HERE, "barrel_roll_offline.npz"))
    its = [(g["X0"], g["U0"], g["V0"]), (g["first_X"][5], g["first_U"][5], g["first_V"][5])]
    E = len(its)
    tile = lambda a: np.tile(np.asarray(a, dtype=np.float64)[None], (E,) + (1,) * np.ndim(a))
    # the dynamics read parameter[t, :4] only (models.py:72); the library takes exactly those columns
    return {"reference": tile(g["reference"]), "parameter": tile(g["parameter"][:, :4]), "x0": tile(g["x0"]),
            "X0": np.stack([i[0] for i in its]), "U0": np.stack([i[1] for i in its]), "V0": np.stack([i[2] for i in its])}


def make(robot, case):
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    pr, pr32 = Problem(cfg), Problem(cfg, dtype=torch.float32)
    bp, bp32 = BatchedProblem(pr), BatchedProblem(pr32)
    segs = parity.segments(cfg)
    if robot == "barrel_roll":
        inp = barrel_roll_candidates()
    else:
        inp = synthetic_inputs(pr, E_CAND, seed=SEED, cold_start=(case == "cold"), scale=6.0 if case == "far" else 1.0)
    if case == "cold" and robot == "talos":
        # Talos at the exact cold start is ill-posed in FP32 in the reference itself: with U = u_ref = 0 the robot is in
        # free fall, the new angular velocity is 0 in exact arithmetic (1e-17 in float64, 2e-8 in float32), and MJX's
        # safe norm switches quat_integrate's Jacobian to zero below 1e-8 -- the float32 oracle takes the other branch
        # and differs from the float64 one by O(1) in A, K, dX.  The Talos cold case therefore keeps the identity
        # quaternion (the zero branch of quat_sub in the cost, which IS robust) but starts the base with 1e-4 rad/s.
        nj = cfg.n_joints
        inp["x0"][:, 10 + nj:13 + nj] = 1e-4
        inp["X0"][:, :, 10 + nj:13 + nj] = 1e-4
    if case == "second":
        X, U, V = bp.mpc(True, inp["reference"], inp["parameter"], inp["x0"], inp["X0"], inp["U0"], inp["V0"])
        inp["X0"], inp["U0"], inp["V0"] = X.numpy(), U.numpy(), V.numpy()
    args = lambda d: (d["reference"], d["parameter"], d["x0"], d["X0"], d["U0"], d["V0"])
    det, det32, dets, dets32 = {}, {}, {}, {}
    Xo, Uo, Vo = bp.mpc(True, *args(inp), detail=det)
    X32, U32, V32 = bp32.mpc(True, *args(inp), detail=det32)
    Xs, Us, Vs = bp.mpc(False, *args(inp), detail=dets)
    Xs32, Us32, Vs32 = bp32.mpc(False, *args(inp), detail=dets32)
    sel = pick(det["alpha_index"].tolist(), det32["alpha_index"].tolist(), dets["alpha_index"].tolist(),
               dets32["alpha_index"].tolist(), E_FIX)
    print(f"{robot}/{case}: candidates alpha {det['alpha_index'].tolist()} (f32 {det32['alpha_index'].tolist()}, "
          f"scan {dets['alpha_index'].tolist()}, scan f32 {dets32['alpha_index'].tolist()}) -> envs {sel}", flush=True)
    out = {k: v[sel] for k, v in inp.items()}
    out["candidate_index"] = np.asarray(sel, dtype=np.int32)
    npy = lambda t: t.detach().numpy()
    for nm in NAMES:
        out[nm] = npy(det[nm])[sel]
    for nm in SCAN_NAMES:
        out[nm + "_scan"] = npy(dets[nm])[sel]
    for nm in ("g", "theta", "phi", "theta_k", "phi_k", "slope"):
        out[nm] = npy(det[nm])[sel]
    out.update(theta_scan=npy(dets["theta"])[sel], phi_scan=npy(dets["phi"])[sel], slope_scan=npy(dets["slope"])[sel],
               alpha_index=npy(det["alpha_index"])[sel].astype(np.int32),
               alpha_index_scan=npy(dets["alpha_index"])[sel].astype(np.int32),
               X=npy(Xo)[sel], U=npy(Uo)[sel], V=npy(Vo)[sel], X_scan=npy(Xs)[sel], U_scan=npy(Us)[sel], V_scan=npy(Vs)[sel])
    # FP32 budgets: the float32 oracle against the float64 oracle under the tests' own metric
    rel_arr = lambda a, b: float(np.abs(a - b).max() / max(1e-3, np.abs(b).max()))
    bud = {}
    for j, e in enumerate(sel):
        for nm in NAMES:
            mag = inp["X0"][e] if nm == "c" else None   # the defect c is judged against the iterate it is a defect of
            bud.setdefault("budget_" + nm, []).append(
                parity.block_rel_err(nm, npy(det32[nm])[e], npy(det[nm])[e], segs, magnitude=mag))
        for nm in SCAN_NAMES:
            bud.setdefault("budget_" + nm + "_scan", []).append(
                parity.block_rel_err(nm, npy(dets32[nm])[e], npy(dets[nm])[e], segs))
        for nm, a, b in (("X", X32, Xo), ("U", U32, Uo), ("V", V32, Vo)):
            bud.setdefault("budget_" + nm, []).append(parity.block_rel_err(nm, npy(a)[e], npy(b)[e], segs))
        for nm, a, b in (("X", Xs32, Xs), ("U", Us32, Us), ("V", Vs32, Vs)):
            bud.setdefault("budget_" + nm + "_scan", []).append(parity.block_rel_err(nm, npy(a)[e], npy(b)[e], segs))
        for nm in ("theta", "phi"):
            bud.setdefault("budget_" + nm, []).append(rel_arr(npy(det32[nm])[e], npy(det[nm])[e]))
            bud.setdefault("budget_" + nm + "_scan", []).append(rel_arr(npy(dets32[nm])[e], npy(dets[nm])[e]))
        print(f"  env {e}: alpha {int(det['alpha_index'][e])} / scan {int(dets['alpha_index'][e])}; f32 budgets "
              + " ".join(f"{k[7:]}={v[-1]:.1e}" for k, v in bud.items()), flush=True)
    for k, v in bud.items():
        out[k] = np.asarray(v, dtype=np.float64)
    for k, v in list(out.items()):
        if v.dtype == np.float64 and v.size > 4096:
            out[k] = v.astype(np.float32)
    path = os.path.join(HERE, f"{robot}_{case}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    robots = sys.argv[1:] or ["srbd", "go2", "aliengo", "h1", "talos"]
    for r in robots:
        for case in (("iter",) if r == "barrel_roll" else CASE_NAMES):
            make(r, case)
