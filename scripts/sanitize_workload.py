"""A small pass over every kernel family of libmpxb for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_workload.py [robots...]
E = 2 environments, one mpc call per (robot, LQ mode, cost structure), the device SQP loop, reference generator, shift,
foot FK, whole-body interface and the stand-alone LQ solve."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpx_b200 import _lib  # noqa: E402
from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver, tvlqr_solve  # noqa: E402

robots = sys.argv[1:] or ["srbd", "go2", "h1", "talos"]
T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
for robot in robots:
    cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{robot}_warm.npz"))
    E = 2
    args = [T(g[k][:E]) for k in ("reference", "parameter")] + [T(cfg.W)] + [T(g[k][:E]) for k in ("x0", "X0", "U0", "V0")]
    combos = [(_lib.LQ_SEQUENTIAL, "dense"), (_lib.LQ_SCAN_PARALLEL, "dense"), (_lib.LQ_SCAN_SERIAL, "dense")]
    if robot != "srbd":
        combos.insert(1, (_lib.LQ_SEQUENTIAL, "diagonal"))
    for mode, structure in combos:
        for eager in (True, False):
            s = MpcSolver(cfg, E, lq_mode=mode, cost_structure=structure, eager_line_search=eager)
            X, U, V = s.solve(*args)
            s.view("P")
            torch.cuda.synchronize()
            assert torch.isfinite(X).all(), (robot, mode, structure)
    s = MpcSolver(cfg, E, limited_memory=True, max_iter=3)
    X, U, V = s.solve(*args)
    X0, U0, V0 = args[4].clone(), args[5].clone(), args[6].clone()
    s.shift_extract(X, U, V, args[3], T(cfg.u_ref), X0, U0, V0, 1, want_cmd=(robot != "srbd"))
    torch.cuda.synchronize()
    print(robot, "mpc ok", flush=True)
    if robot == "srbd":
        from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper as W2
        w = W2(cfg, E)
        feet = args[1][:, 0, 4:].contiguous()
        w.liftoff = feet.clone()
        w.run(args[3], T([[0.3, 0, 0, 0, 0, 0, cfg.robot_height]] * E), feet, torch.ones(E, 4).cuda())
        nq = 19
        qpos = torch.cat([args[3][:, :7], T(np.tile(np.asarray([0, 0.8, -1.6] * 4), (E, 1)))], 1)
        w.whole_body_run(qpos, torch.zeros(E, 18).cuda())
    else:
        from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper as W1
        w = W1(cfg, E, limited_memory=True)
        nj, nc = cfg.n_joints, cfg.n_contact
        x0 = args[3]
        d = w.reset(w.make_data(), x0[:, :7 + nj], x0[:, 7 + nj:13 + 2 * nj], x0[:, 13 + 2 * nj:13 + 2 * nj + 3 * nc])
        d, tau = w.run(d, x0, T([[0.3, 0, 0, 0, 0, 0, cfg.robot_height]] * E))
    torch.cuda.synchronize()
    print(robot, "wrapper ok", flush=True)
# stand-alone LQ solve, small generic size
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lq_problems import random_lq  # noqa: E402
Q, q, R, r, M, A, B, c, x0 = random_lq(8, 2, 7, seed=3)
cc = torch.cat([x0[None], c], 0)
f = lambda t: t.to(torch.float32).unsqueeze(0).repeat(2, *([1] * t.dim())).contiguous().cuda()
for mode in (0, 2, 3):
    tvlqr_solve(*(f(t) for t in (Q, q, R, r, M, A, B, cc)), lq_mode=mode)
torch.cuda.synchronize()
print("tvlqr ok")
