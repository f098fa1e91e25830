"""Gait timer and reference generators, batched over environments on the solver's device.

Host-side mirror of reference ``mpx/utils/mpc_utils.py`` (timer_run 7-14, terrain_orientation 15-36,
reference_generator 38-144, reference_generator_srbd 146-270).  The reference vmaps per-environment JAX
functions; here every function takes a leading env axis and runs as torch ops on whatever device the inputs
live on (the wrappers keep everything on the GPU, so a tick never leaves the device).  These functions produce
the solve's inputs (``reference``, ``parameter``); they are not part of the solve itself.
"""
import math

import torch


def timer_run(duty_factor, step_freq, leg_time, dt):
    """mpc_utils.py:7-14, batched: leg_time [E, nc] -> (contact [E, nc], leg_time)."""
    leg_time = leg_time + dt * step_freq
    leg_time = torch.where(leg_time > 1, leg_time - 1, leg_time)
    contact = torch.where(leg_time < duty_factor, 1.0, 0.0).to(leg_time.dtype)
    return contact, leg_time


def _safe_norm(x):
    """mjx math.norm on the last axis."""
    is_zero = torch.all(torch.abs(x) <= 1e-8, dim=-1, keepdim=True)
    n = torch.linalg.norm(torch.where(is_zero, torch.ones_like(x), x), dim=-1, keepdim=True)
    return torch.where(is_zero, torch.zeros_like(n), n)


def _mat_to_quat_xyzw(R):
    """jax.scipy Rotation.from_matrix(R).as_quat() (x, y, z, w), batched [E,3,3]."""
    E = R.shape[0]
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    decision = torch.stack([R[:, 0, 0], R[:, 1, 1], R[:, 2, 2], tr], dim=1)
    choice = torch.argmax(decision, dim=1)
    quats = []
    for i in range(3):
        j, k = (i + 1) % 3, (i + 2) % 3
        q = torch.zeros(E, 4, dtype=R.dtype, device=R.device)
        q[:, i] = 1 - decision[:, 3] + 2 * R[:, i, i]
        q[:, j] = R[:, j, i] + R[:, i, j]
        q[:, k] = R[:, k, i] + R[:, i, k]
        q[:, 3] = R[:, k, j] - R[:, j, k]
        quats.append(q)
    q3 = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1], 1 + decision[:, 3]], dim=1)
    quats.append(q3)
    allq = torch.stack(quats, dim=1)  # [E,4,4]
    q = allq[torch.arange(E, device=R.device), choice]
    return q / torch.linalg.norm(q, dim=1, keepdim=True)


def terrain_orientation(liftoff_pos, Ryaw):
    """mpc_utils.py:15-36, batched: liftoff [E,12], Ryaw [E,3,3] -> quat (w,x,y,z) [E,4]."""
    vfb = (liftoff_pos[:, :3] + liftoff_pos[:, 3:6] - liftoff_pos[:, 6:9] - liftoff_pos[:, 9:12]) / 2
    vlr = Ryaw[:, :, 1]
    nrm = torch.linalg.cross(vfb, vlr)
    vfb = vfb / _safe_norm(vfb)
    vlr = vlr / _safe_norm(vlr)
    nrm = nrm / _safe_norm(nrm)
    Rm = torch.stack([vfb, vlr, nrm], dim=2)
    return torch.roll(_mat_to_quat_xyzw(Rm), 1, dims=1)


def _generator(srbd, use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot, input,
               duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    E = x.shape[0]
    dev, dtp = x.device, x.dtype
    nc = n_contact
    T = lambda a: torch.as_tensor(a, dtype=dtp, device=dev)
    foot0 = T(foot0)
    p = x[:, :3]
    quat = x[:, 3:7]
    dp = x[:, 7:10] if srbd else x[:, 7 + n_joints:10 + n_joints]
    yaw = torch.atan2(2 * (quat[:, 0] * quat[:, 3] + quat[:, 1] * quat[:, 2]),
                      1 - 2 * (quat[:, 2] * quat[:, 2] + quat[:, 3] * quat[:, 3]))
    cy, sy, z, o = torch.cos(yaw), torch.sin(yaw), torch.zeros_like(yaw), torch.ones_like(yaw)
    Ryaw = torch.stack([torch.stack([cy, -sy, z], 1), torch.stack([sy, cy, z], 1), torch.stack([z, z, o], 1)], 1)
    proprio_height = input[:, 6] + liftoff[:, 2::3].sum(1) / nc
    p = torch.stack([p[:, 0], p[:, 1], proprio_height], 1)
    if use_terrain_estimator:
        quat0 = terrain_orientation(liftoff, Ryaw)
    else:
        quat0 = T([1.0, 0, 0, 0]).expand(E, 4)
    quat_ref = quat0[:, None, :].expand(E, N + 1, 4)
    pitch = torch.asin(2 * (quat0[:, 0] * quat0[:, 2] - quat0[:, 3] * quat0[:, 1]))
    cp, sp = torch.cos(pitch), torch.sin(pitch)
    Rpitch = torch.stack([torch.stack([cp, z, sp], 1), torch.stack([z, o, z], 1), torch.stack([-sp, z, cp], 1)], 1)
    ref_lin_vel = torch.einsum("eij,ejk,ek->ei", Ryaw, Rpitch, input[:, :3])
    ref_ang_vel = input[:, 3:6]
    k = torch.arange(N + 1, dtype=dtp, device=dev)
    p_ref = torch.stack([k[None, :] * dt * ref_lin_vel[:, 0:1] + p[:, 0:1],
                         k[None, :] * dt * ref_lin_vel[:, 1:2] + p[:, 1:2],
                         proprio_height[:, None].expand(E, N + 1)], 2)
    dp_ref = ref_lin_vel[:, None, :].expand(E, N + 1, 3)
    omega_ref = ref_ang_vel[:, None, :].expand(E, N + 1, 3)
    foot_ref = foot[:, None, :].repeat(1, N + 1, 1)
    foot_ref_dot = torch.zeros(E, N + 1, 3 * nc, dtype=dtp, device=dev)
    # foot0 @ block_diag(Ryaw...).T : per foot  Ryaw @ foot0_f
    foot0_projected = p.repeat(1, nc) + torch.einsum("eij,fj->efi", Ryaw, foot0.view(nc, 3)).reshape(E, 3 * nc)
    grf_ref = torch.zeros(E, N + 1, 3 * nc, dtype=dtp, device=dev)
    contact_sequence = torch.zeros(E, N + 1, nc, dtype=dtp, device=dev)

    des_contact, current_timer = timer_run(duty_factor, step_freq, t_timer, dt)
    early = (des_contact == 0) & (contact == 1) & (current_timer > 0.5 + 0.5 * duty_factor)
    initial_speed = -ref_lin_vel / (torch.linalg.norm(ref_lin_vel, dim=1, keepdim=True) + 1e-6) * clearence_speed

    lx, ly, lz = liftoff[:, ::3].clone(), liftoff[:, 1::3].clone(), liftoff[:, 2::3].clone()
    timer = t_timer
    prev_contact = torch.zeros(E, nc, dtype=dtp, device=dev)   # contact_sequence[-1] is still zero at t = 0
    prev = foot.clone()                                        # foot_ref[-1] is still the measured foot at t = 0

    def foothold(direction):
        f1 = 0.5 * ref_lin_vel[:, direction:direction + 1] * duty_factor / step_freq
        f2 = torch.sqrt(input[:, 6:7] / 9.81) * (dp[:, direction:direction + 1] - ref_lin_vel[:, direction:direction + 1])
        return f1 + f2 + foot0_projected[:, direction::3]
    fh_x, fh_y = foothold(0), foothold(1)

    for t in range(N + 1):
        new_contact, timer = timer_run(duty_factor, step_freq, timer, dt)
        contact_sequence[:, t] = new_contact
        lift = (new_contact == 0) & (prev_contact != 0)
        lx = torch.where(lift, prev[:, ::3], lx)
        ly = torch.where(lift, prev[:, 1::3], ly)
        lz = torch.where(lift, prev[:, 2::3], lz)
        val = (timer - duty_factor) / (1 - duty_factor)

        def xy(cur, fh, v0):
            a2 = 3 * (fh - cur) - 2 * v0
            a3 = v0 - 2 * (fh - cur)
            return cur + v0 * val + a2 * val ** 2 + a3 * val ** 3, 2 * a2 * val + 3 * a3 * val ** 2 + v0
        sx, sxd = xy(lx, fh_x, initial_speed[:, 0:1])
        sy_, syd = xy(ly, fh_y, initial_speed[:, 1:2])
        s0 = 0.7
        sh = lz + step_height
        a = 16 * sh - 8 * lz - 8 * lz - 2 * s0
        b = 5 * s0 + 14 * lz + 18 * lz - 32 * sh
        c = 16 * sh - 5 * lz - 11 * lz - 4 * s0
        sz = a * val ** 4 + b * val ** 3 + c * val ** 2 + s0 * val + lz
        szd = 4 * a * val ** 3 + 3 * b * val ** 2 + 2 * c * val + s0
        hold = (new_contact > 0) | early
        cur = torch.empty_like(prev)
        cur[:, ::3] = torch.where(hold, prev[:, ::3], sx)
        cur[:, 1::3] = torch.where(hold, prev[:, 1::3], sy_)
        cur[:, 2::3] = torch.where(hold, prev[:, 2::3], sz)
        foot_ref[:, t] = cur
        zero = torch.zeros_like(sx)
        foot_ref_dot[:, t, ::3] = torch.where(hold, zero, sxd)
        foot_ref_dot[:, t, 1::3] = torch.where(hold, zero, syd)
        foot_ref_dot[:, t, 2::3] = torch.where(hold, zero, szd)
        grf_ref[:, t, 2::3] = new_contact * mass * 9.81 / (new_contact.sum(1, keepdim=True) + 1e-5)
        prev, prev_contact = cur, new_contact
    liftoff = torch.stack([lx, ly, lz], 2).reshape(E, 3 * nc)
    if srbd:
        reference = torch.cat([p_ref, quat_ref, dp_ref, omega_ref, contact_sequence], 2)
        parameter = torch.cat([contact_sequence, foot_ref], 2)
        return reference, parameter, liftoff, foot_ref_dot
    q_ref = T(q0)[None, None, :].expand(E, N + 1, n_joints)
    reference = torch.cat([p_ref, quat_ref, q_ref, dp_ref, omega_ref, foot_ref, contact_sequence, grf_ref], 2)
    return reference, contact_sequence.clone(), liftoff


def reference_generator(use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot, input,
                        duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    """mpc_utils.py:38-144, batched over the leading env axis."""
    return _generator(False, use_terrain_estimator, N, dt, n_joints, n_contact, mass, foot0, q0, t_timer, x, foot,
                      input, duty_factor, step_freq, step_height, liftoff, contact, clearence_speed)


def reference_generator_srbd(use_terrain_estimator, N, dt, n_contact, mass, foot0, t_timer, x, foot, input,
                             duty_factor, step_freq, step_height, liftoff, contact, clearence_speed):
    """mpc_utils.py:146-270, batched over the leading env axis."""
    return _generator(True, use_terrain_estimator, N, dt, 0, n_contact, mass, foot0, None, t_timer, x, foot, input,
                      duty_factor, step_freq, step_height, liftoff, contact, clearence_speed)
