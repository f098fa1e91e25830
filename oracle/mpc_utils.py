"""Restatement of reference ``mpx/utils/mpc_utils.py`` gait timer and reference generators (TEST INFRASTRUCTURE).

Per-environment numpy float64 loops that follow the reference line by line (mpc_utils.py:7-270); they produce
the ``reference`` / ``parameter`` inputs of the solve for the tests and benchmarks.
"""
import numpy as np


def timer_run(duty_factor, step_freq, leg_time, dt):
    """mpc_utils.py:7-14."""
    leg_time = leg_time + dt * step_freq
    leg_time = np.where(leg_time > 1, leg_time - 1, leg_time)
    contact = np.where(leg_time < duty_factor, 1, 0)
    return contact, leg_time


def _safe_norm(x):
    return 0.0 if np.all(np.abs(x) <= 1e-8) else float(np.linalg.norm(x))


def _mat_to_quat_xyzw(R):
    """scipy Rotation.from_matrix(R).as_quat() for a proper rotation (x,y,z,w), Shepperd's method."""
    t = np.trace(R)
    decision = np.array([R[0, 0], R[1, 1], R[2, 2], t])
    choice = int(np.argmax(decision))
    q = np.empty(4)
    if choice != 3:
        i = choice
        j = (i + 1) % 3
        k = (j + 1) % 3
        q[i] = 1 - decision[3] + 2 * R[i, i]
        q[j] = R[j, i] + R[i, j]
        q[k] = R[k, i] + R[i, k]
        q[3] = R[k, j] - R[j, k]
    else:
        q[0] = R[2, 1] - R[1, 2]
        q[1] = R[0, 2] - R[2, 0]
        q[2] = R[1, 0] - R[0, 1]
        q[3] = 1 + decision[3]
    return q / np.linalg.norm(q)


def terrain_orientation(liftoff_pos, Ryaw):
    """mpc_utils.py:15-36 (jax Rotation.from_matrix orthonormalises; the three vectors here already are
    orthogonal up to the front/back tilt, so the nearest rotation is taken by SVD as scipy does)."""
    vfb = (liftoff_pos[:3] + liftoff_pos[3:6] - liftoff_pos[6:9] - liftoff_pos[9:12]) / 2
    vlr = Ryaw @ np.array([0.0, 1.0, 0.0])
    nrm = np.cross(vfb, vlr)
    vfb = vfb / _safe_norm(vfb)
    vlr = vlr / _safe_norm(vlr)
    nrm = nrm / _safe_norm(nrm)
    Rm = np.stack([vfb, vlr, nrm], axis=1)
    q = _mat_to_quat_xyzw(Rm)
    return np.roll(q, 1)


def _splines(liftoff, foothold_x, foothold_y, initial_speed, step_height, val):
    def xy(cur, fh, v0):
        a0, a1 = cur, v0
        a2 = 3 * (fh - cur) - 2 * v0
        a3 = v0 - 2 * (fh - cur)
        return a0 + a1 * val + a2 * val ** 2 + a3 * val ** 3, 2 * a2 * val + 3 * a3 * val ** 2 + a1

    def z(cur, fh, sh):
        s0 = 0.7
        a = 16 * sh - 8 * fh - 8 * cur - 2 * s0
        b = 5 * s0 + 14 * fh + 18 * cur - 32 * sh
        c = 16 * sh - 5 * fh - 11 * cur - 4 * s0
        return (a * val ** 4 + b * val ** 3 + c * val ** 2 + s0 * val + cur,
                4 * a * val ** 3 + 3 * b * val ** 2 + 2 * c * val + s0)
    lx, ly, lz = liftoff
    (sx, sxd) = xy(lx, foothold_x, initial_speed[0])
    (sy, syd) = xy(ly, foothold_y, initial_speed[1])
    (sz, szd) = z(lz, lz, lz + step_height)
    return (sx, sy, sz), (sxd, syd, szd)


def _generator(srbd, use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot, input,
               duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    x = np.asarray(x, dtype=np.float64)
    foot = np.asarray(foot, dtype=np.float64)
    input = np.asarray(input, dtype=np.float64)
    liftoff = np.array(liftoff, dtype=np.float64)
    t_timer = np.asarray(t_timer, dtype=np.float64)
    contact = np.asarray(contact)
    p = x[:3]
    quat = x[3:7]
    dp = x[7:10] if srbd else x[7 + n_joints:10 + n_joints]
    yaw = np.arctan2(2 * (quat[0] * quat[3] + quat[1] * quat[2]), 1 - 2 * (quat[2] * quat[2] + quat[3] * quat[3]))
    Ryaw = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    proprio_height = input[6] + np.sum(liftoff[2::3]) / n_contact
    p = np.array([p[0], p[1], proprio_height])
    if use_terrain_estimator:
        quat_ref = np.tile(terrain_orientation(liftoff, Ryaw), (N + 1, 1))
    else:
        quat_ref = np.tile(np.array([1.0, 0, 0, 0]), (N + 1, 1))
    contact_sequence = np.zeros((N + 1, n_contact))
    pitch = np.arcsin(2 * (quat_ref[0, 0] * quat_ref[0, 2] - quat_ref[0, 3] * quat_ref[0, 1]))
    Rpitch = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    ref_lin_vel = Ryaw @ Rpitch @ input[:3]
    ref_ang_vel = input[3:6]
    k = np.arange(N + 1)
    p_ref = np.stack([k * dt * ref_lin_vel[0] + p[0], k * dt * ref_lin_vel[1] + p[1],
                      np.ones(N + 1) * proprio_height], axis=1)
    dp_ref = np.tile(ref_lin_vel, (N + 1, 1))
    omega_ref = np.tile(ref_ang_vel, (N + 1, 1))
    foot_ref = np.tile(foot, (N + 1, 1))
    foot_ref_dot = np.zeros((N + 1, 3 * n_contact))
    bd = np.kron(np.eye(n_contact), Ryaw)
    foot0_projected = np.tile(p, n_contact) + np.asarray(foot0, dtype=np.float64) @ bd.T
    grf_ref = np.zeros((N + 1, 3 * n_contact))

    des_contact, current_timer = timer_run(duty_factor, step_freq, t_timer, dt)
    early_contact = np.where((des_contact == 0) & (contact == 1) & (current_timer > 0.5 + 0.5 * duty_factor), 1, 0)

    initial_speed = -ref_lin_vel / (np.linalg.norm(ref_lin_vel) + 1e-6) * clearence_speed
    lx, ly, lz = liftoff[::3].copy(), liftoff[1::3].copy(), liftoff[2::3].copy()
    # WB generator keeps a [N+1, nc] timer table initialised with t_timer and reads row t-1 (row -1 at t=0);
    # the SRBD generator carries a single running timer.  Both give timer(t) = advance^(t+1)(t_timer).
    timer = t_timer.copy()
    for t in range(N + 1):
        prev = foot_ref[t - 1].copy()          # t = 0 reads row -1 (still the tiled measured foot)
        prev_contact = contact_sequence[t - 1].copy()
        new_contact, timer = timer_run(duty_factor, step_freq, timer, dt)
        contact_sequence[t] = new_contact
        lift = (contact_sequence[t] == 0) & (prev_contact != 0)
        lx = np.where(lift, prev[::3], lx)
        ly = np.where(lift, prev[1::3], ly)
        lz = np.where(lift, prev[2::3], lz)

        def foothold(direction):
            f1 = 0.5 * ref_lin_vel[direction] * duty_factor / step_freq
            f2 = np.sqrt(input[6] / 9.81) * (dp[direction] - ref_lin_vel[direction])
            return f1 + f2 + foot0_projected[direction::3]
        val = (timer - duty_factor) / (1 - duty_factor)
        (sx, sy, sz), (sxd, syd, szd) = _splines((lx, ly, lz), foothold(0), foothold(1), initial_speed, step_height, val)
        hold = (new_contact > 0) | (early_contact == 1)
        foot_ref[t, ::3] = np.where(hold, prev[::3], sx)
        foot_ref[t, 1::3] = np.where(hold, prev[1::3], sy)
        foot_ref[t, 2::3] = np.where(hold, prev[2::3], sz)
        foot_ref_dot[t, ::3] = np.where(hold, 0, sxd)
        foot_ref_dot[t, 1::3] = np.where(hold, 0, syd)
        foot_ref_dot[t, 2::3] = np.where(hold, 0, szd)
        grf_ref[t, 2::3] = new_contact * mass * 9.81 / (np.sum(new_contact) + 1e-5)
    liftoff[::3], liftoff[1::3], liftoff[2::3] = lx, ly, lz
    if srbd:
        reference = np.concatenate([p_ref, quat_ref, dp_ref, omega_ref, contact_sequence], axis=1)
        parameter = np.concatenate([contact_sequence, foot_ref], axis=1)
        return reference, parameter, liftoff, foot_ref_dot
    q_ref = np.tile(np.asarray(q0, dtype=np.float64), (N + 1, 1))
    reference = np.concatenate([p_ref, quat_ref, q_ref, dp_ref, omega_ref, foot_ref, contact_sequence, grf_ref], axis=1)
    return reference, contact_sequence.copy(), liftoff


def reference_generator(use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot, input,
                        duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    """mpc_utils.py:38-144."""
    return _generator(False, use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot,
                      input, duty_factor, step_freq, step_height, liftoff, contact, clearence_speed)


def reference_generator_srbd(use_terrain_estimator, N, dt, n_contact, mass, foot0, t_timer, x, foot, input,
                             duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    """mpc_utils.py:146-270."""
    return _generator(True, use_terrain_estimator, N, dt, 0, n_contact, mass, foot0, None, t_timer, x, foot, input,
                      duty_factor, step_freq, step_height, liftoff, contact, clearence_speed)


def whole_body_interface(m, contact_id, body_id, Kp, Kd, qpos, qvel, grf, foot_ref, foot_ref_dot, contact):
    """mpc_utils.py:275-313: SRBD ground-reaction forces -> joint torques (torch; ``m`` is the oracle's MJX-style
    model of oracle/mjx_smooth.py).  Returns (tau [nj], J [nv, 3 nc])."""
    import torch
    from . import mjx_smooth
    d = mjx_smooth.fwd_position(m, qpos)
    d = mjx_smooth.fwd_velocity(m, d, qvel)
    M, D = d.qM, d.qfrc_bias
    legs = [d.geom_xpos[g] for g in contact_id]
    J = torch.cat([mjx_smooth.jac(m, d, legs[i], body_id[i])[0] for i in range(len(legs))], dim=1)
    current_leg = torch.cat(legs)
    current_leg_dot = J.T @ qvel
    csa = Kp @ (foot_ref - current_leg) + Kd @ (foot_ref_dot - current_leg_dot)
    tau_fb_lin = D[6:] + (M @ torch.linalg.pinv(J.T) @ csa)[6:]
    tau_mpc = -(J @ grf)[6:]
    mask = torch.repeat_interleave(contact, 3)
    return tau_mpc * mask + (1 - mask) * tau_fb_lin, J


def reference_barell_roll(N, dt, n_joints, n_contact, foot0, q0):
    """mpc_utils.py:316-398, statement by statement (float64 numpy; examples/barel_roll.py:2 enables x64).

    Five phases: full stance (t1), lateral support on FR/RR with a vertical acceleration (t2), flight on the ballistic
    arc (t3), landing (t4), rest.  The base rolls at 2*3.14/(t2+t3+t4) rad/s through phases 2-4 (quat_integrate knot
    by knot, :361-377).  The lateral position restarts at zero in every phase and ``3.14`` is the reference's pi --
    both kept.  Returns (reference [N, 13+nj+7nc], parameter [N, 4nc])."""
    from . import mjx_math
    import torch
    t1, t2, t3, t4 = 0.2, 0.2, 0.3, 0.1
    z_start, z_land = 0.4, 0.28
    v_lateral = -0.25 / (t2 + t3)
    v0 = (z_land - z_start + 0.5 * 9.81 * t3 * t3) / t3
    roll_speed = 2 * 3.14 / (t2 + t3 + t4)
    acc = v0 / t2
    foot0 = np.asarray(foot0, dtype=np.float64).reshape(-1)
    q0 = np.asarray(q0, dtype=np.float64).reshape(-1)
    n1, n2, n3, n4 = int(t1 / dt), int(t2 / dt), int(t3 / dt), int(t4 / dt)
    # :335-341
    p1 = np.tile(np.array([0, 0, 0.33]), (n1, 1))
    p1[:, 1] = np.arange(n1) * dt * v_lateral
    dp1 = np.tile(np.array([0, v_lateral, 0]), (n1, 1))
    contact1 = np.tile(np.array([1, 1, 1, 1]), (n1, 1))
    quat1 = np.tile(np.array([1.0, 0, 0, 0]), (n1, 1))
    omega1 = np.zeros((n1, 3))
    # :343-349
    p2 = np.tile(np.array([0, p1[-1, 1], 0.33]), (n2, 1))
    p2[:, 2] = 0.5 * np.arange(n2) * dt * np.arange(n2) * dt * acc + 0.33
    p2[:, 1] = np.arange(n2) * dt * v_lateral
    dp2 = np.tile(np.array([0, v_lateral, 0]), (n2, 1))
    dp2[:, 2] = np.arange(n2) * dt * acc
    contact2 = np.tile(np.array([0, 1, 0, 1]), (n2, 1))
    # :354-367
    p3 = np.tile(np.array([0, p2[-1, 1], p2[-1, 2]]), (n3, 1))
    p3[:, 1] = np.arange(n3) * dt * v_lateral
    dp3 = np.tile(np.array([0, v_lateral, 0]), (n3, 1))
    for i in range(n3):
        t = i * dt
        p3[i, 2] = z_start - 0.5 * 9.81 * t ** 2 + v0 * t
        dp3[i, 2] = -9.81 * t + v0
    contact3 = np.zeros((n3, 4))
    # :369-374
    p4 = np.tile(np.array([0, p3[-1, 1], z_land]), (n4, 1))
    dp4 = np.zeros((n4, 3))
    contact4 = np.ones((n4, 4))
    # :376-378
    quat234 = np.tile(np.array([1.0, 0.0, 0, 0]), (n2 + n3 + n4, 1))
    w = torch.tensor([roll_speed, 0.0, 0.0], dtype=torch.float64)
    for t in range(1, n2 + n3 + n4):
        quat234[t] = mjx_math.quat_integrate(torch.as_tensor(quat234[t - 1]), w, dt).numpy()
    omega234 = np.tile(np.array([roll_speed, 0, 0]), (n2 + n3 + n4, 1))
    # :380-386
    n5 = N - (n1 + n2 + n3 + n4)
    p5 = np.tile(np.array([0, p4[-1, 1], z_land]), (n5, 1))
    dp5 = np.zeros((n5, 3))
    quat5 = np.tile(np.array([1.0, 0, 0, 0]), (n5, 1))
    omega5 = np.zeros((n5, 3))
    contact5 = np.ones((n5, 4))
    # :388-398
    p_ref = np.concatenate([p1, p2, p3, p4, p5])
    quat_ref = np.concatenate([quat1, quat234, quat5])
    q_ref = np.tile(q0, (N, 1))
    dp_ref = np.concatenate([dp1, dp2, dp3, dp4, dp5])
    omega_ref = np.concatenate([omega1, omega234, omega5])
    foot_ref = np.tile(foot0, (N, 1)) + np.tile(p_ref, n_contact)
    foot_ref[:, 2::3] = 0.0
    contact_sequence = np.concatenate([contact1, contact2, contact3, contact4, contact5])
    grf_ref = np.zeros((N, 3 * n_contact))
    return (np.concatenate([p_ref, quat_ref, q_ref, dp_ref, omega_ref, foot_ref, contact_sequence, grf_ref], axis=1),
            np.concatenate([contact_sequence, foot_ref], axis=1))
