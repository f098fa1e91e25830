"""Run the oracle on the inputs dumped by scripts/dump_nonfinite.py (CPU):  python scripts/check_nonfinite_oracle.py go2"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.problems import Problem  # noqa: E402
robot = sys.argv[1] if len(sys.argv) > 1 else "go2"
cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
d = np.load(os.path.join(ROOT, "gpurun_out", f"nonfinite_{robot}.npz"))
pr = Problem(cfg)
for e in d["envs"]:
    g = lambda k: d[f"e{e}_{k}"]
    print("env", e, "tick", g("tick"), "GPU alpha_index", g("alpha_index"), "GPU finite X", np.isfinite(g("X")).all(),
          "inputs finite", all(np.isfinite(g(k)).all() for k in ("reference", "parameter", "x0", "X0", "U0", "V0")))
    print("  x0", np.round(g("x0")[:19], 3))
    for nm in ("dX", "dU", "c", "ls_theta", "ls_phi"):
        if f"e{e}_{nm}" in d.files:
            a = g(nm)
            print("  GPU", nm, "finite", np.isfinite(a).all(), "max|.|", np.nanmax(np.abs(a)) if np.isfinite(a).any() else None)
    det = {}
    Xo, Uo, Vo = pr.mpc(True, g("reference"), g("parameter"), g("x0"), g("X0"), g("U0"), g("V0"), detail=det)
    print("  oracle f64: finite X", bool(torch.isfinite(Xo).all()), "alpha_index", det.get("alpha_index"),
          "theta", np.asarray(det.get("theta"))[:4] if det.get("theta") is not None else None)
    for k in ("dX", "dU", "c"):
        if k in det:
            a = np.asarray(det[k])
            print("  oracle", k, "finite", np.isfinite(a).all(), "max|.|", np.nanmax(np.abs(a)))
