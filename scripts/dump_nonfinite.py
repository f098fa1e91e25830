"""Debug helper (GPU box): replay the bench's closed loop and dump the solve inputs of every environment whose solve
output is non-finite at some tick, so the float64 oracle can be run on the same inputs on the CPU
(scripts/check_nonfinite_oracle.py).   python scripts/dump_nonfinite.py go2 2048 [ticks] -> gpurun_out/nonfinite_<robot>.npz"""
import os, sys
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

robot = sys.argv[1] if len(sys.argv) > 1 else "go2"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
K = int(sys.argv[3]) if len(sys.argv) > 3 else 13
dev = torch.device("cuda:0")
args = SimpleNamespace()
lp = bench.Loop(args, robot, E, K, 0, 0, dev, True)
found = {}
for k in range(K):
    reference, parameter, x0c = lp.tape[lp.pos]
    X0, U0, V0 = lp.X0.clone(), lp.U0.clone(), lp.V0.clone()
    lp.step()
    torch.cuda.synchronize()
    bad = (~torch.isfinite(lp.X).flatten(1).all(1)).nonzero().flatten().tolist()
    s = lp.solver
    print("tick", k, "nonfinite envs", bad, "alpha hist", torch.bincount(s.alpha_index.to(torch.int64), minlength=11).tolist(), flush=True)
    for e in bad:
        if e in found:
            continue
        d = {"tick": k, "env": e, "reference": reference[e], "parameter": parameter[e], "x0": x0c[e], "X0": X0[e], "U0": U0[e],
             "V0": V0[e], "X": lp.X[e], "U": lp.U[e], "V": lp.V[e], "alpha_index": s.alpha_index[e]}
        for name in ("dX", "dU", "dV", "c", "ls_theta", "ls_phi", "g"):
            try:
                d[name] = s.view(name)[e]
            except Exception as ex:  # structured mode does not materialise every view
                print("view", name, ex)
        found[e] = {kk: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for kk, v in d.items()}
out = {}
for e, d in found.items():
    for kk, v in d.items():
        out[f"e{e}_{kk}"] = v
out["envs"] = np.asarray(sorted(found), dtype=np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"nonfinite_{robot}.npz"), **out)
print("saved", sorted(found))
