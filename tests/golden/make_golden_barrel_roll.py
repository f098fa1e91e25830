"""Golden fixture of the offline barrel-roll solve (reference examples/barel_roll.py:68 ->
MPCControllerWrapper.runOffline, mpc_wrapper.py:387-468, config_barrel_roll.py): the float64 oracle's host loop
(oracle/wrapper.py::offline_loop) on the oracle's reference_barell_roll, limited_memory=False as the example
constructs the wrapper.  Stores the inputs, the first iterates, the per-iteration (g, sum c^2) history, the iteration
count and the final trajectory.      python tests/golden/make_golden_barrel_roll.py   (about 10 minutes on 8 cores)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mpx_b200.config.config_barrel_roll as cfg  # noqa: E402
from oracle import mpc_utils as omu, wrapper as owrap  # noqa: E402
from oracle.batched import BatchedProblem  # noqa: E402
from oracle.problems import Problem  # noqa: E402

MAX_ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 100
KEEP = 6
pr = Problem(cfg)
bp = BatchedProblem(pr)
ref, par = omu.reference_barell_roll(cfg.N + 1, cfg.dt, cfg.n_joints, cfg.n_contact, cfg.p_legs0, cfg.q0)
qpos = np.concatenate([cfg.p0, cfg.quat0, cfg.q0])
qvel = np.zeros(6 + cfg.n_joints)
foot = np.asarray(pr.foot_fk(qpos)).reshape(-1)
x0 = np.concatenate([qpos, qvel, foot, np.zeros(3 * cfg.n_contact)])
nq = 13 + cfg.n_joints
X0 = np.tile(x0, (cfg.N + 1, 1))
X0[:, :nq] = ref[:, :nq]
U0 = np.tile(np.asarray(cfg.u_ref, dtype=np.float64), (cfg.N, 1))
V0 = np.zeros((cfg.N + 1, cfg.n))

# the first KEEP iterates one by one (for the per-iteration parity test), then the whole loop
t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
X, U, V = t(X0[None]), t(U0[None]), t(V0[None])
firstX, firstU, firstV, alphas = [], [], [], []
for i in range(KEEP):
    det = {}
    X, U, V = bp.mpc(False, t(ref[None]), t(par[None]), t(x0[None]), X, U, V, detail=det)
    firstX.append(X[0].numpy().copy()); firstU.append(U[0].numpy().copy()); firstV.append(V[0].numpy().copy())
    alphas.append(int(np.asarray(det["alpha_index"]).reshape(-1)[0]))
    print("iterate", i, "alpha_index", alphas[-1], flush=True)
Xf, Uf, Vf, iters, hist = owrap.offline_loop(bp, False, ref[None], par[None], x0[None], X0[None], U0[None], V0[None], MAX_ITER)
g_hist = np.array([float(h[1][0]) for h in hist])
th_hist = np.array([float(h[2][0]) for h in hist])
print("iterations", int(iters[0]), "final g", g_hist[-1], "final sum c^2", th_hist[-1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "barrel_roll_offline.npz"),
                    qpos=qpos, qvel=qvel, reference=ref, parameter=par, x0=x0, X0=X0, U0=U0, V0=V0,
                    first_X=np.stack(firstX), first_U=np.stack(firstU), first_V=np.stack(firstV),
                    first_alpha=np.asarray(alphas), iterations=int(iters[0]), g_hist=g_hist, theta_hist=th_hist,
                    X=Xf[0].numpy(), U=Uf[0].numpy(), V=Vf[0].numpy(), max_iter=MAX_ITER)
