"""Restatement of the mujoco-mjx 3.3.0 subset the reference dynamics consume (TEST INFRASTRUCTURE).

The reference calls ``mjx.make_data`` / ``mjx.fwd_position`` / ``mjx.fwd_velocity`` / ``mjx.jac`` and then
reads ``qLD`` (dense upper Cholesky factor of qM), ``qfrc_bias`` and ``geom_xpos``
(reference mpx/utils/models.py:59-92, 121-142, 172-203).  mujoco-mjx is absent from /root/reference
(third-party, pinned ``mujoco-mjx==3.3.0``); this file restates its published algorithm
(``mjx/_src/smooth.py``: kinematics, com_pos, crb, factor_m, com_vel, rne; ``mjx/_src/support.py``: jac, make_m)
in the same COM-centred spatial-vector convention: motion vectors are [angular, linear], cinert is
(6 triu inertia, 3 mass*offset, mass) about the root subtree COM.

Only free + hinge joints are supported (all four robots of the reference configs use nothing else).
"""
from types import SimpleNamespace

import torch

from . import mjx_math as math

JNT_FREE, JNT_HINGE = 0, 3
MJ_MINVAL = 1e-15


def put_model(rm, dtype=torch.float64):
    """RobotModel (numpy, mpx_b200.mjcf) -> namespace of torch tensors + python index lists."""
    t = lambda a: torch.as_tensor(a, dtype=dtype)
    m = SimpleNamespace()
    m.dtype = dtype
    m.nbody, m.njnt, m.nq, m.nv = rm.nbody, rm.njnt, rm.nq, rm.nv
    m.gravity = t(rm.gravity)
    for f in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia",
              "jnt_pos", "jnt_axis", "qpos0", "dof_armature", "geom_pos"):
        setattr(m, f, t(getattr(rm, f)))
    for f in ("body_parentid", "body_rootid", "body_jntadr", "body_jntnum", "jnt_type", "jnt_bodyid",
              "jnt_qposadr", "jnt_dofadr", "dof_bodyid", "dof_jntid", "dof_parentid", "geom_bodyid"):
        setattr(m, f, [int(v) for v in getattr(rm, f)])
    m.geom_names, m.body_names = list(rm.geom_names), list(rm.body_names)
    m.children = [[] for _ in range(m.nbody)]
    for b in range(1, m.nbody):
        m.children[m.body_parentid[b]].append(b)
    # dofs of each body
    m.body_dofs = [[d for d in range(m.nv) if m.dof_bodyid[d] == b] for b in range(m.nbody)]
    return m


def kinematics(m, qpos):
    """mjx smooth.kinematics: body frames, joint anchors/axes in the world frame."""
    xpos = [None] * m.nbody
    xquat = [None] * m.nbody
    xanchor = [None] * m.njnt
    xaxis = [None] * m.njnt
    z = torch.zeros(3, dtype=qpos.dtype)
    xpos[0] = z
    xquat[0] = torch.tensor([1.0, 0, 0, 0], dtype=qpos.dtype)
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        pos = xpos[p] + math.rotate(m.body_pos[b].to(qpos.dtype), xquat[p])
        quat = math.quat_mul(xquat[p], m.body_quat[b].to(qpos.dtype))
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            qa = m.jnt_qposadr[j]
            if m.jnt_type[j] == JNT_FREE:
                anchor = qpos[qa:qa + 3]
                axis = torch.tensor([0.0, 0.0, 1.0], dtype=qpos.dtype)
                pos = qpos[qa:qa + 3]
                quat = math.normalize(qpos[qa + 3:qa + 7])
            elif m.jnt_type[j] == JNT_HINGE:
                anchor = math.rotate(m.jnt_pos[j].to(qpos.dtype), quat) + pos
                axis = math.rotate(m.jnt_axis[j].to(qpos.dtype), quat)
                angle = qpos[qa] - m.qpos0[qa].to(qpos.dtype)
                qloc = math.axis_angle_to_quat(m.jnt_axis[j].to(qpos.dtype), angle)
                quat = math.quat_mul(quat, qloc)
            else:
                raise NotImplementedError("joint type")
            xanchor[j], xaxis[j] = anchor, axis
            # correct for off-center rotation
            pos = anchor - math.rotate(m.jnt_pos[j].to(qpos.dtype), quat)
        xpos[b], xquat[b] = pos, quat
    d = SimpleNamespace()
    d.xpos, d.xquat, d.xanchor, d.xaxis = xpos, xquat, xanchor, xaxis
    d.xmat = [math.quat_to_mat(q) for q in xquat]
    d.xipos = [xpos[b] + math.rotate(m.body_ipos[b].to(qpos.dtype), xquat[b]) for b in range(m.nbody)]
    d.ximat = [math.quat_to_mat(math.quat_mul(xquat[b], m.body_iquat[b].to(qpos.dtype))) for b in range(m.nbody)]
    d.geom_xpos = [xpos[m.geom_bodyid[g]] + math.rotate(m.geom_pos[g].to(qpos.dtype), xquat[m.geom_bodyid[g]])
                   for g in range(len(m.geom_bodyid))]
    return d


def com_pos(m, d):
    """mjx smooth.com_pos: subtree COM, cinert, cdof."""
    dtype = d.xpos[0].dtype
    pos = [d.xipos[b] * m.body_mass[b].to(dtype) for b in range(m.nbody)]
    mass = [m.body_mass[b].to(dtype) for b in range(m.nbody)]
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        pos[p] = pos[p] + pos[b]
        mass[p] = mass[p] + mass[b]
    subtree_com = []
    for b in range(m.nbody):
        if float(mass[b]) < MJ_MINVAL:
            subtree_com.append(d.xipos[b])
        else:
            subtree_com.append(pos[b] / mass[b])
    d.subtree_com = subtree_com
    eye = torch.eye(3, dtype=dtype)
    cinert = []
    for b in range(m.nbody):
        root_com = subtree_com[m.body_rootid[b]]
        off = d.xipos[b] - root_com
        h = torch.stack([math.cross(off, -eye[i]) for i in range(3)])
        inert = (d.ximat[b] * m.body_inertia[b].to(dtype)) @ d.ximat[b].T + h @ h.T * m.body_mass[b].to(dtype)
        tri = torch.stack([inert[0, 0], inert[1, 1], inert[2, 2], inert[0, 1], inert[0, 2], inert[1, 2]])
        cinert.append(torch.cat([tri, off * m.body_mass[b].to(dtype), m.body_mass[b].to(dtype).reshape(1)]))
    d.cinert = cinert
    cdof = [None] * m.nv
    dof_com = lambda a, o: torch.cat([a, math.cross(a, o)])
    for j in range(m.njnt):
        b = m.jnt_bodyid[j]
        root_com = subtree_com[m.body_rootid[b]]
        offset = root_com - d.xanchor[j]
        da = m.jnt_dofadr[j]
        if m.jnt_type[j] == JNT_FREE:
            for i in range(3):
                cdof[da + i] = torch.cat([torch.zeros(3, dtype=dtype), eye[i]])
            for i in range(3):
                cdof[da + 3 + i] = dof_com(d.xmat[b][:, i], offset)
        else:
            cdof[da] = dof_com(d.xaxis[j], offset)
    d.cdof = cdof
    return d


def crb(m, d):
    """mjx smooth.crb + support.make_m (dense): composite rigid body inertia -> qM."""
    dtype = d.cinert[0].dtype
    crb_body = [c for c in d.cinert]
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        crb_body[p] = crb_body[p] + crb_body[b]
    crb_body[0] = torch.zeros(10, dtype=dtype)
    crb_cdof = [math.inert_mul(crb_body[m.dof_bodyid[i]], d.cdof[i]) for i in range(m.nv)]
    rows = []
    M = [[None] * m.nv for _ in range(m.nv)]
    zero = torch.zeros((), dtype=dtype)
    for i in range(m.nv):
        for j in range(m.nv):
            M[i][j] = zero
    for i in range(m.nv):
        j = i
        while j >= 0:
            val = torch.dot(crb_cdof[i], d.cdof[j])
            M[i][j] = val
            M[j][i] = val
            j = m.dof_parentid[j]
        M[i][i] = M[i][i] + m.dof_armature[i].to(dtype)
    for i in range(m.nv):
        rows.append(torch.stack(M[i]))
    d.qM = torch.stack(rows)
    return d


def factor_m(m, d):
    """Dense path of mjx smooth.factor_m: qLD = cho_factor(qM) (upper factor, jax.scipy default lower=False)."""
    d.qLD = torch.linalg.cholesky(d.qM, upper=True)
    return d


def com_vel(m, d, qvel):
    """mjx smooth.com_vel: cvel per body and cdof_dot per dof."""
    dtype = qvel.dtype
    cvel = [None] * m.nbody
    cdof_dot = [None] * m.nv
    cvel[0] = torch.zeros(6, dtype=dtype)
    for b in range(1, m.nbody):
        cv = cvel[m.body_parentid[b]]
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            da = m.jnt_dofadr[j]
            if m.jnt_type[j] == JNT_FREE:
                for i in range(3):
                    cv = cv + d.cdof[da + i] * qvel[da + i]
                for i in range(3):
                    cdof_dot[da + i] = torch.zeros(6, dtype=dtype)
                for i in range(3):
                    cdof_dot[da + 3 + i] = math.motion_cross(cv, d.cdof[da + 3 + i])
                for i in range(3):
                    cv = cv + d.cdof[da + 3 + i] * qvel[da + 3 + i]
            else:
                cdof_dot[da] = math.motion_cross(cv, d.cdof[da])
                cv = cv + d.cdof[da] * qvel[da]
        cvel[b] = cv
    d.cvel, d.cdof_dot = cvel, cdof_dot
    return d


def rne(m, d, qvel):
    """mjx smooth.rne: qfrc_bias (Coriolis/centrifugal + gravity)."""
    dtype = qvel.dtype
    cacc = [None] * m.nbody
    cacc[0] = torch.cat([torch.zeros(3, dtype=dtype), -m.gravity.to(dtype)])
    for b in range(1, m.nbody):
        a = cacc[m.body_parentid[b]]
        for dd in m.body_dofs[b]:
            a = a + d.cdof_dot[dd] * qvel[dd]
        cacc[b] = a
    loc = []
    for b in range(m.nbody):
        f = math.inert_mul(d.cinert[b], cacc[b])
        f = f + math.motion_cross_force(d.cvel[b], math.inert_mul(d.cinert[b], d.cvel[b]))
        loc.append(f)
    cfrc = list(loc)
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        cfrc[p] = cfrc[p] + cfrc[b]
    d.qfrc_bias = torch.stack([torch.dot(d.cdof[i], cfrc[m.dof_bodyid[i]]) for i in range(m.nv)])
    return d


def fwd_position(m, qpos):
    d = kinematics(m, qpos)
    d = com_pos(m, d)
    d = crb(m, d)
    d = factor_m(m, d)
    return d


def fwd_velocity(m, d, qvel):
    d = com_vel(m, d, qvel)
    d = rne(m, d, qvel)
    return d


def jac(m, d, point, body_id):
    """mjx support.jac: (jacp [nv,3], jacr [nv,3]) of a world point attached to ``body_id``."""
    dtype = point.dtype
    anc = set()
    b = body_id
    while b > 0:
        anc.add(b)
        b = m.body_parentid[b]
    offset = point - d.subtree_com[m.body_rootid[body_id]]
    jacp, jacr = [], []
    for i in range(m.nv):
        if m.dof_bodyid[i] in anc:
            jacp.append(d.cdof[i][3:] + math.cross(d.cdof[i][:3], offset))
            jacr.append(d.cdof[i][:3])
        else:
            jacp.append(torch.zeros(3, dtype=dtype))
            jacr.append(torch.zeros(3, dtype=dtype))
    return torch.stack(jacp), torch.stack(jacr)


def cho_solve_upper(U, b):
    """jax.scipy.linalg.cho_solve((U, False), b): solve (U^T U) x = b."""
    y = torch.linalg.solve_triangular(U.T, b.reshape(b.shape[0], -1), upper=False)
    x = torch.linalg.solve_triangular(U, y, upper=True)
    return x.reshape(b.shape)
