"""ctypes binding of libmpxb.so (include/mpxb.h).  No fallback: a missing library raises at import of the solver."""
import ctypes as C
import os
from functools import partial

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPXB_LIB_PATH: another build of the same library (A/B measurements of kernel variants); never a different backend
LIB_PATH = os.environ.get("MPXB_LIB_PATH") or os.path.join(_HERE, "libmpxb.so")

MAX_BODY, MAX_DOF, MAX_CONTACT = 32, 32, 8
KIND_SRBD, KIND_QUAD_WB, KIND_H1_WB, KIND_TALOS_WB = 0, 1, 2, 3
OPT_EAGER_LINE_SEARCH = 1
OPT_LINE_SEARCH = 2
OPT_COST_STRUCTURE = 3
OPT_STOP_RULE = 4
COST_DENSE, COST_DIAGONAL_W = 0, 1
LS_FILTER, LS_MERIT, LS_ARMIJO = 0, 1, 2
STOP_OFFLINE, STOP_LEGACY = 0, 1
LQ_SEQUENTIAL, LQ_SCAN, LQ_SCAN_SERIAL, LQ_SCAN_PARALLEL = 0, 1, 2, 3
ERRORS = {0: "ok", -1: "MPXB_EINVAL", -2: "MPXB_EUNSUPPORTED", -3: "MPXB_EWORKSPACE", -4: "MPXB_ECUDA"}


class ModelDesc(C.Structure):
    _fields_ = [
        ("version", C.c_int32), ("kind", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("N", C.c_int32),
        ("n_joints", C.c_int32), ("n_contact", C.c_int32), ("nref", C.c_int32), ("npar", C.c_int32),
        ("nw", C.c_int32), ("swing_tracking", C.c_int32), ("dt", C.c_float),
        ("mass", C.c_float), ("inertia", C.c_float * 9), ("inertia_inv", C.c_float * 9),
        ("nbody", C.c_int32), ("nv", C.c_int32), ("gravity", C.c_float * 3),
        ("baumgarte_alpha", C.c_float), ("contact_reg", C.c_float),
        ("body_parentid", C.c_int32 * MAX_BODY), ("body_dofadr", C.c_int32 * MAX_BODY),
        ("body_jnttype", C.c_int32 * MAX_BODY),
        ("body_pos", (C.c_float * 3) * MAX_BODY), ("body_quat", (C.c_float * 4) * MAX_BODY),
        ("body_ipos", (C.c_float * 3) * MAX_BODY), ("body_iquat", (C.c_float * 4) * MAX_BODY),
        ("body_mass", C.c_float * MAX_BODY), ("body_inertia", (C.c_float * 3) * MAX_BODY),
        ("jnt_pos", (C.c_float * 3) * MAX_BODY), ("jnt_axis", (C.c_float * 3) * MAX_BODY),
        ("jnt_qpos0", C.c_float * MAX_BODY), ("dof_armature", C.c_float * MAX_DOF),
        ("contact_body", C.c_int32 * MAX_CONTACT), ("contact_pos", (C.c_float * 3) * MAX_CONTACT),
    ]


class Gait(C.Structure):
    _fields_ = [("use_terrain_estimation", C.c_int32), ("advance_timer_first", C.c_int32), ("mpc_dt", C.c_float),
                ("duty_factor", C.c_float), ("step_freq", C.c_float), ("step_height", C.c_float), ("mass", C.c_float),
                ("clearence_speed", C.c_float), ("foot0", C.c_float * (3 * MAX_CONTACT)), ("q0", C.c_float * MAX_DOF)]


_lib = None


def lib():
    """Load libmpxb.so once; raise loudly if it has not been built (python __graft_entry__.py / make -C mpx_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is the only implementation of the solve "
            "(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32p, i32p, sz = C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t
    L.mpxb_version_string.restype = C.c_char_p
    L.mpxb_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    L.mpxb_model_create.restype = i32
    L.mpxb_model_set_option.argtypes = [vp, i32, i32]
    L.mpxb_model_set_option.restype = i32
    L.mpxb_model_destroy.argtypes = [vp]
    L.mpxb_model_destroy.restype = None
    L.mpxb_workspace_bytes.argtypes = [vp, i32, i32]
    L.mpxb_workspace_bytes.restype = sz
    L.mpxb_mpc_step.argtypes = [vp, vp, i32, i32, i32, f32p, f32p, f32p, sz, f32p, f32p, f32p, f32p,
                                f32p, f32p, f32p, i32p, i32p, i32p, vp, sz]
    L.mpxb_mpc_step.restype = i32
    L.mpxb_workspace_view.argtypes = [vp, i32, i32, vp, C.c_char_p, C.POINTER(sz)]
    L.mpxb_workspace_view.restype = vp
    L.mpxb_workspace_expand.argtypes = [vp, vp, i32, i32, vp]
    L.mpxb_workspace_expand.restype = i32
    L.mpxb_shift_extract.argtypes = [vp, vp, i32, i32] + [f32p] * 11
    L.mpxb_shift_extract.restype = i32
    L.mpxb_tvlqr.argtypes = [vp, i32, i32, i32, i32, i32] + [f32p] * 16 + [vp]
    L.mpxb_tvlqr.restype = i32
    L.mpxb_tvlqr_scan_bytes.argtypes = [i32, i32, i32, i32]
    L.mpxb_tvlqr_scan_bytes.restype = sz
    L.mpxb_wb_foot_positions.argtypes = [vp, vp, i32, f32p, f32p]
    L.mpxb_wb_foot_positions.restype = i32
    L.mpxb_whole_body_interface.argtypes = [vp, vp, i32] + [f32p] * 10
    L.mpxb_whole_body_interface.restype = i32
    L.mpxb_reference_generator.argtypes = [vp, vp, i32, C.POINTER(Gait), f32p, f32p, i32] + [f32p] * 8
    L.mpxb_reference_generator.restype = i32
    L.mpxb_profile_create.argtypes = [i32, C.POINTER(vp)]
    L.mpxb_profile_create.restype = i32
    L.mpxb_profile_destroy.argtypes = [vp]
    L.mpxb_profile_destroy.restype = None
    L.mpxb_profile_attach.argtypes = [vp, vp]
    L.mpxb_profile_attach.restype = i32
    L.mpxb_profile_read.argtypes = [vp, C.POINTER(C.c_double * 4), C.POINTER(i32)]
    L.mpxb_profile_read.restype = i32
    L.mpxb_fp32_peak_tflops.argtypes = [vp, i32]
    L.mpxb_fp32_peak_tflops.restype = C.c_double
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}")


def _unwrap(desc):
    args = ()
    while isinstance(desc, partial):
        args = desc.args + args
        desc = desc.func
    return desc, args


def _fill(arr, values):
    values = np.asarray(values)
    if values.ndim == 1:
        for i, v in enumerate(values):
            arr[i] = v
    else:
        for i, row in enumerate(values):
            for j, v in enumerate(row):
                arr[i][j] = v


def make_desc(config, robot_model=None):
    """config module (+ parsed MJCF tables for whole-body kinds) -> ModelDesc."""
    dyn, _ = _unwrap(config.dynamics)
    cost, cargs = _unwrap(config.cost)
    d = ModelDesc()
    d.version = 1
    d.kind = dyn.kind
    d.n, d.m, d.N = int(config.n), int(config.m), int(config.N)
    d.n_joints, d.n_contact = int(config.n_joints), int(config.n_contact)
    d.dt = float(config.dt)
    d.nw = int(np.asarray(config.W).shape[0])
    nj, nc = d.n_joints, d.n_contact
    if d.kind == KIND_SRBD:
        d.nref, d.npar = 13 + nc, 4 * nc
        d.mass = float(config.mass)
        inertia = np.asarray(config.inertia, dtype=np.float64)
        _fill(d.inertia, inertia.reshape(-1))
        _fill(d.inertia_inv, np.linalg.inv(inertia).reshape(-1))
        return d
    d.nref, d.npar = 13 + nj + 7 * nc, nc
    d.swing_tracking = int(bool(cargs[0])) if cargs else 1
    d.baumgarte_alpha, d.contact_reg = float(dyn.baumgarte_alpha), float(dyn.contact_reg)
    rm = robot_model
    if rm is None:
        from .mjcf import load_model
        rm = load_model(config.model_path)
    if rm.nbody > MAX_BODY or rm.nv > MAX_DOF or nc > MAX_CONTACT:
        raise ValueError("model exceeds the compiled table sizes")
    d.nbody, d.nv = rm.nbody, rm.nv
    _fill(d.gravity, rm.gravity)
    _fill(d.body_parentid, rm.body_parentid)
    dofadr = -np.ones(rm.nbody, dtype=np.int32)
    jtype = -np.ones(rm.nbody, dtype=np.int32)
    jpos = np.zeros((rm.nbody, 3)); jaxis = np.zeros((rm.nbody, 3)); q0 = np.zeros(rm.nbody)
    for b in range(rm.nbody):
        if rm.body_jntnum[b] > 1:
            raise ValueError("bodies with more than one joint are not supported")
        if rm.body_jntnum[b] == 1:
            j = rm.body_jntadr[b]
            if rm.jnt_type[j] not in (0, 3):
                raise ValueError("only free and hinge joints are supported")
            dofadr[b], jtype[b] = rm.jnt_dofadr[j], rm.jnt_type[j]
            jpos[b], jaxis[b] = rm.jnt_pos[j], rm.jnt_axis[j]
            q0[b] = rm.qpos0[rm.jnt_qposadr[j]] if rm.jnt_type[j] == 3 else 0.0
    _fill(d.body_dofadr, dofadr); _fill(d.body_jnttype, jtype)
    _fill(d.body_pos, rm.body_pos); _fill(d.body_quat, rm.body_quat)
    _fill(d.body_ipos, rm.body_ipos); _fill(d.body_iquat, rm.body_iquat)
    _fill(d.body_mass, rm.body_mass); _fill(d.body_inertia, rm.body_inertia)
    _fill(d.jnt_pos, jpos); _fill(d.jnt_axis, jaxis); _fill(d.jnt_qpos0, q0)
    _fill(d.dof_armature, rm.dof_armature)
    # foot point -> body, following each reference dynamics' pairing (models.py:78-87, 131-138, 183-201)
    gids = [rm.geom_id(nm) for nm in config.contact_frame]
    bids = [rm.body_id(nm) for nm in config.body_name]
    if d.kind == KIND_QUAD_WB:
        per_foot = bids
    elif d.kind == KIND_H1_WB:
        per_foot = [bids[0], bids[0], bids[1], bids[1]]
    else:
        per_foot = [bids[0]] * 4 + [bids[1]] * 4
    _fill(d.contact_body, per_foot)
    _fill(d.contact_pos, rm.geom_pos[gids])
    for g, b in zip(gids, per_foot):
        if rm.geom_bodyid[g] != b:
            raise ValueError("contact geom is not attached to the body the dynamics pair it with")
    return d


def make_gait(config, mass, duty_factor, step_freq, step_height, clearence_speed, use_terrain, advance_timer_first):
    """config + the wrapper's (mutable) gait attributes -> Gait for mpxb_reference_generator."""
    g = Gait()
    g.use_terrain_estimation = int(bool(use_terrain))
    g.advance_timer_first = int(bool(advance_timer_first))
    g.mpc_dt = 1.0 / float(config.mpc_frequency)
    g.duty_factor, g.step_freq, g.step_height = float(duty_factor), float(step_freq), float(step_height)
    g.mass, g.clearence_speed = float(mass), float(clearence_speed)
    _fill(g.foot0, np.asarray(config.p_legs0, dtype=np.float64).reshape(-1))
    q0 = getattr(config, "q0", None)
    if q0 is not None and len(np.asarray(q0).reshape(-1)) <= MAX_DOF:
        _fill(g.q0, np.asarray(q0, dtype=np.float64).reshape(-1))
    return g
