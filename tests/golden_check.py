"""Shared checker of the golden-fixture parity tests (TEST INFRASTRUCTURE): load a fixture, run one mpc call through
the C ABI, compare every intermediate and output block by block (tests/parity.py)."""
import json
import os

import numpy as np
import torch

import parity

REL_TOL = 1e-4
# no scan-mode budget, however large, relaxes a check beyond 5 %: the associative scan in float32 on whole-body LQ data
# is unstable enough that equivalent float32 evaluations of the reference algorithm differ by O(1) (parity;
# tests/golden/budgets_*.json), and a budget of that size would make the comparison vacuous
TOL_CAP = 5e-2
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIN = ("A", "B", "Q", "R", "M", "q", "r", "c")
LQ = ("K", "k", "P", "p", "dX", "dU", "dV")


def tolerance(g, key, e):
    """max(1e-4, 4 x FP32 budget); the part of a scan-mode budget that exceeds the sequential budget of the same
    quantity is capped at TOL_CAP (see above)."""
    if key.endswith("_scan"):
        seq = g.budget(key[:-5], e) if key[:-5] in g.budgets else 0.0
        return max(REL_TOL, min(4 * g.budget(key, e), max(TOL_CAP, 4 * seq)))
    return max(REL_TOL, 4 * g.budget(key, e))


class Fixture:
    """A golden fixture (tests/golden/<robot>_<case>.npz) with its FP32 budgets (tests/golden/budgets_<robot>.json,
    made by tests/golden/make_budgets.py): fx["A"] reads an array, fx.budget("A", e) the budget of environment e."""

    def __init__(self, robot, case):
        path = os.path.join(GOLDEN, f"{robot}_{case}.npz")
        assert os.path.exists(path), f"{path} missing (run tests/golden/make_golden.py)"
        self.npz = np.load(path)
        bpath = os.path.join(GOLDEN, f"budgets_{robot}.json")
        assert os.path.exists(bpath), f"{bpath} missing (run tests/golden/make_budgets.py)"
        with open(bpath) as f:
            self.budgets = json.load(f)[f"{robot}_{case}"]
        self.files = self.npz.files

    def __getitem__(self, key):
        return self.npz[key]

    def budget(self, key, e):
        return float(self.budgets[key][e])


def load(robot, case):
    return Fixture(robot, case)


def solve(cfg, g, lq_mode, eager=True, cost_structure="dense"):
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    E = g["x0"].shape[0]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    solver = MpcSolver(cfg, E, lq_mode=lq_mode, eager_line_search=eager, cost_structure=cost_structure)
    X, U, V = solver.solve(T(g["reference"]), T(g["parameter"]), T(cfg.W), T(g["x0"]), T(g["X0"]), T(g["U0"]), T(g["V0"]))
    torch.cuda.synchronize()
    return solver, X, U, V


def check_step(cfg, g, solver, X, U, V, sfx, eager):
    """Compare one mpc call with the fixture; sfx = "" (sequential LQ) or "_scan" (associative scan)."""
    segs = parity.segments(cfg)
    E = g["x0"].shape[0]
    fails = []
    for e in range(E):
        tol = lambda key: tolerance(g, key, e)
        for nm in LIN:
            err = parity.block_rel_err(nm, solver.view(nm)[e].cpu().numpy(), g[nm][e], segs,
                                       magnitude=g["X0"][e] if nm == "c" else None)
            if not err < tol(nm):
                fails.append((e, nm, err, tol(nm)))
        for nm in LQ:
            err = parity.block_rel_err(nm, solver.view(nm)[e].cpu().numpy(), g[nm + sfx][e], segs)
            if not err < tol(nm + sfx):
                fails.append((e, nm + sfx, err, tol(nm + sfx)))
        if not parity.scalar_rel_err(float(solver.view("g")[e]), float(g["g"][e])) < REL_TOL:
            fails.append((e, "g", float(solver.view("g")[e]), float(g["g"][e])))
        a_ref = int(g["alpha_index" + sfx][e])
        npts = 11 if (eager or a_ref > 0) else 1  # lazy: points 1..10 only if alpha = 1 is rejected
        for nm, view, floor in (("phi", "ls_phi", REL_TOL), ("theta", "ls_theta", 1e-3)):
            got = solver.view(view)[e, :npts].cpu().numpy().astype(np.float64)
            ref = g[nm + sfx][e][:npts]
            err = float(np.abs(got - ref).max() / max(1e-3, np.abs(ref).max()))
            if not err < max(floor, tol(nm + sfx)):
                fails.append((e, nm + sfx, err, max(floor, tol(nm + sfx))))
        if int(solver.alpha_index[e]) != a_ref:
            fails.append((e, "alpha_index" + sfx, int(solver.alpha_index[e]), a_ref))
        for nm, got in (("X", X), ("U", U), ("V", V)):
            err = parity.block_rel_err(nm, got[e].cpu().numpy(), g[nm + sfx][e], segs)
            if not err < tol(nm + sfx):
                fails.append((e, nm + sfx, err, tol(nm + sfx)))
        if int(solver.status[e]) != 0 or int(solver.iters[e]) != 1:
            fails.append((e, "status/iters", int(solver.status[e]), int(solver.iters[e])))
    assert not fails, fails
