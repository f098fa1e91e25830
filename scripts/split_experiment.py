"""Measurement: does running the batch as two half-batches on two streams (kernels of the two halves overlap, tails of one
fill with the other's CTAs) beat one full batch?  Go2 whole-body, synthetic warm inputs, solve only."""
import sys, time, importlib
import numpy as np, torch
sys.path.insert(0, ".")
from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
from oracle.problems import Problem, synthetic_inputs

robot, E = sys.argv[1], int(sys.argv[2])
cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
dev = torch.device("cuda:0")
pr = Problem(cfg)
inp = synthetic_inputs(pr, 8, seed=3)
T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev).contiguous()
rep = lambda a: T(a).repeat((E // 8,) + (1,) * (np.asarray(a).ndim - 1)).contiguous()
ref, par, x0, X0, U0, V0 = (rep(inp[k]) for k in ("reference", "parameter", "x0", "X0", "U0", "V0"))
W = T(cfg.W)

def run(parts, iters=30):
    sl = [slice(i * E // parts, (i + 1) * E // parts) for i in range(parts)]
    solvers = [MpcSolver(cfg, E // parts, limited_memory=True, device=dev) for _ in sl]
    streams = [torch.cuda.Stream(dev) for _ in sl]
    outs = [(torch.empty_like(X0[s]), torch.empty_like(U0[s]), torch.empty_like(V0[s])) for s in sl]
    def once():
        main = torch.cuda.current_stream(dev)
        for st in streams: st.wait_stream(main)
        for s, so, st, o in zip(sl, solvers, streams, outs):
            with torch.cuda.stream(st):
                so.solve(ref[s], par[s], W, x0[s], X0[s], U0[s], V0[s], out=o)
        for st in streams: main.wait_stream(st)
    for _ in range(5): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): once()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, torch.cat([o[0] for o in outs])

t1, X1 = run(1)
t2, X2 = run(2)
t4, X4 = run(4)
print(f"{robot} E={E}: one batch {t1:.3f} ms, two halves {t2:.3f} ms, four quarters {t4:.3f} ms; bitwise equal: {bool((X1 == X2).all())} {bool((X1 == X4).all())}")
