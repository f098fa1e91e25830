"""Host-side mirror of reference ``mpx/utils/mpc_utils.py``.

The per-tick generators of the reference -- ``timer_run`` (7-14), ``terrain_orientation`` (15-36),
``reference_generator`` (38-144), ``reference_generator_srbd`` (146-270) and ``whole_body_interface`` (275-313) -- run
as CUDA kernels here (``mpxb_reference_generator`` in csrc/refgen.cuh, ``mpxb_whole_body_interface`` in
csrc/wb_kernels.cuh); the wrappers call them through ``MpcSolver.reference_generator`` /
``whole_body_run``.  The names below are kept so that configs and wrappers read like the reference's; calling one
directly raises and says which device entry point does the work (there is no host implementation of the solve path).

``reference_barell_roll`` (316-398) is different: it is evaluated ONCE, on the host, before the offline solve
(``MPCControllerWrapper.runOffline``, mpc_wrapper.py:416), so it is plain numpy here.
"""
import numpy as np


class _DeviceOnly:
    def __init__(self, name, where, entry):
        self.__name__, self._where, self._entry = name, where, entry

    def __call__(self, *a, **k):
        raise RuntimeError(f"mpc_utils.{self.__name__} (reference mpc_utils.py:{self._where}) runs on the device here: "
                           f"use {self._entry}")


timer_run = _DeviceOnly("timer_run", "7-14", "the wrapper's run() / MpcSolver.reference_generator (gait timer fused in)")
reference_generator = _DeviceOnly("reference_generator", "38-144",
                                  "BatchedMPCControllerWrapper.reference_generator / MpcSolver.reference_generator")
reference_generator_srbd = _DeviceOnly("reference_generator_srbd", "146-270", "MpcSolver.reference_generator")
whole_body_interface = _DeviceOnly("whole_body_interface", "275-313", "mpc_wrapper_srbd.BatchedMPCControllerWrapper.whole_body_run")


def reference_barell_roll(N, dt, n_joints, n_contact, foot0, q0):
    """Reference trajectory of the barrel roll (reference mpc_utils.py:316-398; config_barrel_roll.py:91).

    Phases, in knots of ``dt``: stance 0.2 s | lateral support on the right legs 0.2 s, pushing up with v0/t2 | flight
    0.3 s on the ballistic arc from z = 0.4 to z = 0.28 | landing 0.1 s | rest.  Through phases 2-4 the base rolls about
    x at 2*3.14 / 0.6 rad/s (the reference's 3.14).  The lateral coordinate is ``k dt v_lateral`` with k restarting in
    every phase, as in the reference.  Returns (reference [N, 13 + nj + 7 nc], parameter [N, 4 nc]) in float64:
    reference = p | quat | q | dp | omega | foot | contact | grf, parameter = contact | foot."""
    g = 9.81
    t_stance, t_push, t_fly, t_land = 0.2, 0.2, 0.3, 0.1
    z0, z_up, z_down = 0.33, 0.4, 0.28
    vy = -0.25 / (t_push + t_fly)
    v0 = (z_down - z_up + 0.5 * g * t_fly * t_fly) / t_fly
    roll_rate = 2 * 3.14 / (t_push + t_fly + t_land)
    counts = [int(t_stance / dt), int(t_push / dt), int(t_fly / dt), int(t_land / dt)]
    counts.append(N - sum(counts))
    foot0 = np.asarray(foot0, dtype=np.float64).reshape(-1)
    nref = 13 + n_joints + 7 * n_contact
    ref = np.zeros((N, nref))
    contact = np.zeros((N, n_contact))
    p, quat, dp, omega = ref[:, 0:3], ref[:, 3:7], ref[:, 7 + n_joints:10 + n_joints], ref[:, 10 + n_joints:13 + n_joints]
    quat[:, 0] = 1.0
    ref[:, 7:7 + n_joints] = np.asarray(q0, dtype=np.float64).reshape(-1)
    start = np.concatenate([[0], np.cumsum(counts)])
    patterns = [(1, 1, 1, 1), (0, 1, 0, 1), (0, 0, 0, 0), (1, 1, 1, 1), (1, 1, 1, 1)]
    y_hold = 0.0
    for ph, (a, cnt) in enumerate(zip(start[:-1], counts)):
        k = np.arange(cnt)
        rows = slice(a, a + cnt)
        contact[rows] = patterns[ph][:n_contact]
        if ph < 3:
            p[rows, 1] = k * dt * vy
            dp[rows, 1] = vy
            y_hold = p[a + cnt - 1, 1] if cnt else y_hold
        else:
            p[rows, 1] = y_hold  # landing and rest hold the last lateral position of the flight
        if ph == 0:
            p[rows, 2] = z0
        elif ph == 1:
            p[rows, 2] = 0.5 * (k * dt) ** 2 * (v0 / t_push) + z0
            dp[rows, 2] = k * dt * (v0 / t_push)
        elif ph == 2:
            p[rows, 2] = z_up - 0.5 * g * (k * dt) ** 2 + v0 * k * dt
            dp[rows, 2] = -g * k * dt + v0
        else:
            p[rows, 2] = z_down
    # roll: knot j of phases 2-4 has turned by j * roll_rate * dt about x (quat_integrate of a constant body rate)
    a, b = start[1], start[4]
    half = 0.5 * roll_rate * dt * np.arange(b - a)
    quat[a:b, 0], quat[a:b, 1] = np.cos(half), np.sin(half)
    omega[a:b, 0] = roll_rate
    f0 = 13 + n_joints
    foot = np.tile(foot0, (N, 1)) + np.tile(p, (1, n_contact))
    foot[:, 2::3] = 0.0
    ref[:, f0:f0 + 3 * n_contact] = foot
    ref[:, f0 + 3 * n_contact:f0 + 4 * n_contact] = contact
    return ref, np.concatenate([contact, foot], axis=1)
