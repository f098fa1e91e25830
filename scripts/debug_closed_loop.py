"""Debug helper: closed-loop ticks on synthetic states, reporting alpha indices / non-finite environments per tick."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_states
from mpx_b200.utils.mpc_wrapper import BatchedMPCControllerWrapper

robot = sys.argv[1] if len(sys.argv) > 1 else "go2"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
regen = len(sys.argv) > 3 and sys.argv[3] == "regen"
cfg = importlib.import_module(f"mpx_b200.config.config_{robot}")
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
st = synthetic_states(cfg, E, 0)
wrap = BatchedMPCControllerWrapper(cfg, E, limited_memory=True, device=dev)
qpos, qvel = T(st["qpos"]), T(st["qvel"])
feet = wrap.foot_positions(qpos)
data = wrap.reset(wrap.make_data(), qpos, qvel, feet).replace(contact_time=T(st["timer"]))
x0 = data.X0[:, 0, :].clone()
cmd = T(st["cmd"])
s = wrap._solver
if regen:
    for k in range(12):
        data, tau = wrap.run(data, x0, cmd)
        x0 = wrap.X[:, 1, :].clone()
        bad = (~torch.isfinite(wrap.X).flatten(1).all(1)).sum().item()
        print(k, "alpha hist", torch.bincount(s.alpha_index, minlength=11).tolist(), "nonfinite envs", bad,
              "theta0", float(s.view("ls_theta")[:, 0].median()), "g", float(s.view("g").median()))
else:
    nc = cfg.n_contact
    reference, parameter, _ = wrap.reference_generator(
        duty_factor=data.duty_factor, step_freq=data.step_freq, step_height=data.step_height, t_timer=data.contact_time,
        x=x0, foot=feet, input=cmd, liftoff=data.liftoff, contact=torch.zeros(E, nc, device=dev))
    X0, U0, V0 = data.X0.clone(), data.U0.clone(), data.V0.clone()
    for k in range(12):
        X, U, V = s.solve(reference.contiguous(), parameter.contiguous(), wrap.W, x0, X0, U0, V0)
        bad = (~torch.isfinite(X).flatten(1).all(1))
        print(k, "alpha hist", torch.bincount(s.alpha_index, minlength=11).tolist(), "nonfinite envs", int(bad.sum()),
              "status", int(s.status.sum()), "theta0 med", float(s.view("ls_theta")[:, 0].median()),
              "max|dX|", float(s.view("dX").abs().max()), "nonfinite P", int((~torch.isfinite(s.view("P"))).sum()),
              "nonfinite A", int((~torch.isfinite(s.view("A"))).sum()))
        s.shift_extract(X, U, V, x0, wrap.u_ref, X0, U0, V0, 1, want_cmd=False)
        x0 = X[:, 1, :].clone()
