"""Print err / tolerance of every quantity of every golden fixture for each LQ mode (diagnostics for the -m gpu tests)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from golden_check import load, solve, LIN, LQ, REL_TOL, TOL_CAP, tolerance  # noqa: E402
from mpx_b200 import _lib  # noqa: E402

robots = sys.argv[1].split(",") if len(sys.argv) > 1 else ["srbd", "go2", "aliengo", "h1", "talos"]
modes = {"seq": (_lib.LQ_SEQUENTIAL, ""), "seq3": (_lib.LQ_SEQUENTIAL, ""), "pscan": (_lib.LQ_SCAN_PARALLEL, "_scan"), "sscan": (_lib.LQ_SCAN_SERIAL, "_scan")}
for robot in robots:
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    segs = parity.segments(cfg)
    for case in ("warm", "second", "cold", "far"):
        g = load(robot, case)
        for mname, (mode, sfx) in modes.items():
            if mname == "seq3" and robot == "srbd":
                continue
            solver, X, U, V = solve(cfg, g, mode, True, "diagonal" if mname == "seq3" else "dense")
            E = g["x0"].shape[0]
            worst = {}
            for e in range(E):
                tol = lambda key: tolerance(g, key, e)
                items = [(nm, solver.view(nm)[e], g[nm][e], nm) for nm in (LIN if mname in ("seq", "seq3") else ())]
                items += [(nm, solver.view(nm)[e], g[nm + sfx][e], nm + sfx) for nm in LQ]
                items += [(nm, got[e], g[nm + sfx][e], nm + sfx) for nm, got in (("X", X), ("U", U), ("V", V))]
                for nm, got, ref, key in items:
                    err = parity.block_rel_err(nm, got.cpu().numpy(), ref, segs, magnitude=g["X0"][e] if nm == "c" else None)
                    r = err / tol(key)
                    if r > worst.get(key, (0, 0, 0, 0))[0]:
                        worst[key] = (r, e, err, tol(key))
                a_ok = int(solver.alpha_index[e]) == int(g["alpha_index" + sfx][e])
                if not a_ok:
                    worst["alpha_e%d" % e] = (99.0, e, int(solver.alpha_index[e]), int(g["alpha_index" + sfx][e]))
            bad = {k: v for k, v in worst.items() if v[0] > 0.5}
            print(f"{robot:8s}{case:7s}{mname:6s} " + ("ok" if not any(v[0] >= 1 for v in worst.values()) else "FAIL") + " "
                  + " ".join(f"{k}:e{v[1]}:{v[2]:.1e}/{v[3]:.1e}" for k, v in sorted(bad.items(), key=lambda kv: -kv[1][0])), flush=True)
