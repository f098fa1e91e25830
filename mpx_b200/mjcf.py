"""Minimal MJCF loader: the subset of MuJoCo's model compiler the MPC path needs.

The reference builds its model with ``mujoco.MjModel.from_xml_path(config.model_path)``
followed by ``mjx.put_model`` (reference mpx/utils/mpc_wrapper.py:43-47,207-211).  Neither
``mujoco`` nor ``mjx`` exists in this environment, and the solve only ever consumes the
kinematic tree, the body inertials, joint axes/anchors/armature, gravity and the named foot
geoms.  This module parses exactly that out of an MJCF file into flat
numpy arrays that use MuJoCo's own field names (``body_parentid``, ``jnt_axis`` ...), so the
oracle can restate MJX against it and the CUDA side can be fed from it.

Supported: ``<include>``, ``<compiler angle=...>``, ``<option gravity=...>``, nested
``<default class=...>`` trees with ``childclass`` / ``class`` resolution for ``joint`` and
``geom``, ``<body pos quat euler>``, ``<inertial pos quat mass diaginertia|fullinertia>``,
``<joint type pos axis ref armature>``, ``<freejoint>``, named ``<geom pos>``.
Anything else in the file is ignored (meshes, actuators, sensors, contacts, visuals).

A model can also be stored as / loaded from a small JSON "model table" (``save_tables`` /
``load_tables``) so that the robot descriptions ship with the package without the meshes.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3


@dataclass
class RobotModel:
    """MuJoCo-named flat model arrays (body 0 is the world)."""

    nbody: int
    njnt: int
    nq: int
    nv: int
    gravity: np.ndarray  # [3]
    body_names: list
    body_parentid: np.ndarray  # [nbody]
    body_rootid: np.ndarray  # [nbody]
    body_pos: np.ndarray  # [nbody,3]
    body_quat: np.ndarray  # [nbody,4]
    body_ipos: np.ndarray  # [nbody,3]
    body_iquat: np.ndarray  # [nbody,4]
    body_mass: np.ndarray  # [nbody]
    body_inertia: np.ndarray  # [nbody,3] principal moments
    body_jntadr: np.ndarray  # [nbody] first joint or -1
    body_jntnum: np.ndarray  # [nbody]
    jnt_names: list
    jnt_type: np.ndarray  # [njnt]
    jnt_bodyid: np.ndarray
    jnt_pos: np.ndarray  # [njnt,3]
    jnt_axis: np.ndarray  # [njnt,3]
    jnt_qposadr: np.ndarray
    jnt_dofadr: np.ndarray
    qpos0: np.ndarray  # [nq]
    dof_bodyid: np.ndarray  # [nv]
    dof_jntid: np.ndarray  # [nv]
    dof_parentid: np.ndarray  # [nv]
    dof_armature: np.ndarray  # [nv]
    geom_names: list
    geom_bodyid: np.ndarray  # [ngeom]  (named geoms only)
    geom_pos: np.ndarray  # [ngeom,3]
    source: str = ""
    extra: dict = field(default_factory=dict)

    # -- name lookups (mirror mjx.name2id use at reference mpc_wrapper.py:59-64) --
    def geom_id(self, name: str) -> int:
        return self.geom_names.index(name)

    def body_id(self, name: str) -> int:
        return self.body_names.index(name)

    @property
    def total_mass(self) -> float:
        return float(self.body_mass.sum())


# --------------------------------------------------------------------------------------
# small math helpers (host side, float64)
# --------------------------------------------------------------------------------------

def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(t) for t in s.replace(",", " ").split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got '{s}'")
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def _euler_to_quat(e, seq="xyz"):
    q = np.array([1.0, 0, 0, 0])
    for ax, ang in zip(seq, e):
        h = 0.5 * ang
        r = np.array([math.cos(h), 0, 0, 0])
        r["xyz".index(ax.lower()) + 1] = math.sin(h)
        # lower-case = intrinsic (rotating frame): post-multiply; upper = extrinsic: pre-multiply
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q


# --------------------------------------------------------------------------------------
# parser
# --------------------------------------------------------------------------------------

class _Defaults:
    """Resolved <default> tree: class name -> {element tag -> attrib dict}."""

    def __init__(self):
        self.classes = {"main": {}}
        self.parent = {"main": None}

    def add_tree(self, node, parent_name=None):
        name = node.get("class") or "main"
        if name not in self.classes:
            self.classes[name] = {}
            self.parent[name] = parent_name
        base = self.classes[name]
        if parent_name is not None and parent_name != name:
            # inherit a copy of the parent's settings first
            for tag, attrs in self.classes[parent_name].items():
                merged = dict(attrs)
                merged.update(base.get(tag, {}))
                base[tag] = merged
        for child in node:
            if child.tag == "default":
                continue
            d = base.setdefault(child.tag, {})
            d.update(child.attrib)
        for child in node:
            if child.tag == "default":
                self.add_tree(child, name)

    def get(self, cls, tag):
        cls = cls or "main"
        if cls not in self.classes:
            raise KeyError(f"unknown default class '{cls}'")
        return self.classes[cls].get(tag, {})


def _expand_includes(root, base_dir):
    """Replace every <include file=...> by the children of the included file's root."""
    changed = True
    while changed:
        changed = False
        for parent in list(root.iter()):
            for i, child in enumerate(list(parent)):
                if child.tag == "include":
                    inc = ET.parse(os.path.join(base_dir, child.get("file"))).getroot()
                    parent.remove(child)
                    for j, sub in enumerate(list(inc)):
                        parent.insert(i + j, sub)
                    changed = True
    return root


def load_mjcf(path: str) -> RobotModel:
    """Parse an MJCF file into a RobotModel (see module docstring for the supported subset)."""
    root = ET.parse(path).getroot()
    base_dir = os.path.dirname(os.path.abspath(path))
    _expand_includes(root, base_dir)

    angle_scale = 1.0
    eulerseq = "xyz"
    for comp in root.findall("compiler"):
        # MuJoCo's default is degrees
        ang = comp.get("angle")
        if ang is not None:
            angle_scale = 1.0 if ang == "radian" else math.pi / 180.0
        eulerseq = comp.get("eulerseq", eulerseq)
    if not root.findall("compiler") or all(c.get("angle") is None for c in root.findall("compiler")):
        angle_scale = math.pi / 180.0

    gravity = np.array([0.0, 0.0, -9.81])
    for opt in root.findall("option"):
        if opt.get("gravity") is not None:
            gravity = _vec(opt.get("gravity"), 3)

    defaults = _Defaults()
    for d in root.findall("default"):
        defaults.add_tree(d, None)

    def frame_quat(el):
        if el.get("quat") is not None:
            q = _vec(el.get("quat"), 4)
            return q / np.linalg.norm(q)
        if el.get("euler") is not None:
            return _euler_to_quat(_vec(el.get("euler"), 3) * angle_scale, eulerseq)
        for unsupported in ("axisangle", "xyaxes", "zaxis"):
            if el.get(unsupported) is not None:
                raise NotImplementedError(f"orientation spec '{unsupported}' on <{el.tag}>")
        return np.array([1.0, 0, 0, 0])

    B = dict(names=["world"], parent=[0], pos=[np.zeros(3)], quat=[np.array([1.0, 0, 0, 0])],
             ipos=[np.zeros(3)], iquat=[np.array([1.0, 0, 0, 0])], mass=[0.0], inertia=[np.zeros(3)],
             jntadr=[-1], jntnum=[0])
    J = dict(names=[], type=[], body=[], pos=[], axis=[], ref=[], armature=[])
    G = dict(names=[], body=[], pos=[])

    def merged(el, childclass):
        cls = el.get("class") or childclass
        attrs = dict(defaults.get(cls, el.tag if el.tag != "freejoint" else "__none__"))
        attrs.update(el.attrib)
        return attrs

    def visit(body_el, parent_id, childclass):
        childclass = body_el.get("childclass") or childclass
        bid = len(B["names"])
        B["names"].append(body_el.get("name", f"body{bid}"))
        B["parent"].append(parent_id)
        B["pos"].append(_vec(body_el.get("pos"), 3, [0, 0, 0]))
        B["quat"].append(frame_quat(body_el))
        inert = body_el.find("inertial")
        if inert is None:
            # the MPC models all carry explicit inertials; a mass-less
            # frame body is legal and contributes nothing.
            B["ipos"].append(np.zeros(3)); B["iquat"].append(np.array([1.0, 0, 0, 0]))
            B["mass"].append(0.0); B["inertia"].append(np.zeros(3))
        else:
            B["ipos"].append(_vec(inert.get("pos"), 3, [0, 0, 0]))
            B["mass"].append(float(inert.get("mass")))
            if inert.get("fullinertia") is not None:
                f = _vec(inert.get("fullinertia"), 6)  # xx yy zz xy xz yz
                M = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                w, V = np.linalg.eigh(M)
                # descending principal moments, right-handed frame (as MuJoCo's compiler does)
                order = np.argsort(-w)
                w, V = w[order], V[:, order]
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                q_full = mat_to_quat(V)
                B["iquat"].append(quat_mul(frame_quat(inert), q_full))
                B["inertia"].append(w)
            else:
                B["iquat"].append(frame_quat(inert))
                B["inertia"].append(_vec(inert.get("diaginertia"), 3))
        jnts = [c for c in body_el if c.tag in ("joint", "freejoint")]
        B["jntadr"].append(len(J["names"]) if jnts else -1)
        B["jntnum"].append(len(jnts))
        for j in jnts:
            if j.tag == "freejoint":
                a = dict(j.attrib); a["type"] = "free"
            else:
                a = merged(j, childclass)
            typ = a.get("type", "hinge")
            J["names"].append(a.get("name", f"jnt{len(J['names'])}"))
            J["type"].append({"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[typ])
            J["body"].append(bid)
            J["pos"].append(_vec(a.get("pos"), 3, [0, 0, 0]))
            ax = _vec(a.get("axis"), 3, [0, 0, 1])
            J["axis"].append(ax / max(np.linalg.norm(ax), 1e-300))
            ref = float(a.get("ref", 0.0))
            J["ref"].append(ref * (angle_scale if typ == "hinge" else 1.0))
            J["armature"].append(0.0 if typ == "free" and j.tag == "freejoint" else float(a.get("armature", 0.0)))
        for g in body_el.findall("geom"):
            a = merged(g, childclass)
            if a.get("name") is None:
                continue
            if a.get("fromto") is not None:
                ft = _vec(a.get("fromto"), 6)
                pos = 0.5 * (ft[:3] + ft[3:])
            else:
                pos = _vec(a.get("pos"), 3, [0, 0, 0])
            G["names"].append(a["name"]); G["body"].append(bid); G["pos"].append(pos)
        for child in body_el.findall("body"):
            visit(child, bid, childclass)

    for wb in root.findall("worldbody"):
        wb_class = wb.get("childclass")
        for body_el in wb.findall("body"):
            visit(body_el, 0, wb_class)

    nbody = len(B["names"])
    njnt = len(J["names"])
    qposadr, dofadr, qpos0 = [], [], []
    dof_body, dof_jnt, dof_arm = [], [], []
    nq = nv = 0
    for j in range(njnt):
        qposadr.append(nq); dofadr.append(nv)
        t = J["type"][j]
        if t == JNT_FREE:
            b = J["body"][j]
            qpos0 += list(B["pos"][b]) + list(B["quat"][b])
            nq += 7; w = 6
        elif t == JNT_BALL:
            qpos0 += [1.0, 0, 0, 0]; nq += 4; w = 3
        else:
            qpos0.append(J["ref"][j]); nq += 1; w = 1
        for _ in range(w):
            dof_body.append(J["body"][j]); dof_jnt.append(j); dof_arm.append(J["armature"][j])
        nv += w

    # dof_parentid: previous dof in the same body, else last dof of the nearest ancestor with dofs
    body_lastdof = [-1] * nbody
    dof_parent = []
    for d in range(nv):
        b = dof_body[d]
        if d > 0 and dof_body[d - 1] == b:
            dof_parent.append(d - 1)
        else:
            p = B["parent"][b]
            while p > 0 and body_lastdof[p] < 0:
                p = B["parent"][p]
            dof_parent.append(body_lastdof[p] if p > 0 else -1)
        body_lastdof[b] = d

    rootid = [0] * nbody
    for b in range(1, nbody):
        rootid[b] = b if B["parent"][b] == 0 else rootid[B["parent"][b]]

    return RobotModel(
        nbody=nbody, njnt=njnt, nq=nq, nv=nv, gravity=gravity,
        body_names=B["names"], body_parentid=np.array(B["parent"], dtype=np.int32),
        body_rootid=np.array(rootid, dtype=np.int32),
        body_pos=np.array(B["pos"]), body_quat=np.array(B["quat"]),
        body_ipos=np.array(B["ipos"]), body_iquat=np.array(B["iquat"]),
        body_mass=np.array(B["mass"]), body_inertia=np.array(B["inertia"]),
        body_jntadr=np.array(B["jntadr"], dtype=np.int32), body_jntnum=np.array(B["jntnum"], dtype=np.int32),
        jnt_names=J["names"], jnt_type=np.array(J["type"], dtype=np.int32),
        jnt_bodyid=np.array(J["body"], dtype=np.int32),
        jnt_pos=np.array(J["pos"]).reshape(njnt, 3), jnt_axis=np.array(J["axis"]).reshape(njnt, 3),
        jnt_qposadr=np.array(qposadr, dtype=np.int32), jnt_dofadr=np.array(dofadr, dtype=np.int32),
        qpos0=np.array(qpos0), dof_bodyid=np.array(dof_body, dtype=np.int32),
        dof_jntid=np.array(dof_jnt, dtype=np.int32), dof_parentid=np.array(dof_parent, dtype=np.int32),
        dof_armature=np.array(dof_arm),
        geom_names=G["names"], geom_bodyid=np.array(G["body"], dtype=np.int32),
        geom_pos=np.array(G["pos"]).reshape(len(G["names"]), 3),
        source=os.path.basename(path),
    )


# --------------------------------------------------------------------------------------
# JSON model tables
# --------------------------------------------------------------------------------------

_ARRAY_FIELDS = [
    "gravity", "body_parentid", "body_rootid", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass",
    "body_inertia", "body_jntadr", "body_jntnum", "jnt_type", "jnt_bodyid", "jnt_pos", "jnt_axis", "jnt_qposadr",
    "jnt_dofadr", "qpos0", "dof_bodyid", "dof_jntid", "dof_parentid", "dof_armature", "geom_bodyid", "geom_pos",
]
_INT_FIELDS = {"body_parentid", "body_rootid", "body_jntadr", "body_jntnum", "jnt_type", "jnt_bodyid",
               "jnt_qposadr", "jnt_dofadr", "dof_bodyid", "dof_jntid", "dof_parentid", "geom_bodyid"}


def save_tables(model: RobotModel, path: str, keep_geoms=None) -> None:
    d = {"format": "mpx_b200.model_tables.v1", "source": model.source,
         "nbody": model.nbody, "njnt": model.njnt, "nq": model.nq, "nv": model.nv,
         "body_names": model.body_names, "jnt_names": model.jnt_names}
    gsel = list(range(len(model.geom_names)))
    if keep_geoms is not None:
        gsel = [i for i in gsel if model.geom_names[i] in keep_geoms]
    d["geom_names"] = [model.geom_names[i] for i in gsel]
    for f in _ARRAY_FIELDS:
        a = getattr(model, f)
        if f in ("geom_bodyid", "geom_pos"):
            a = a[gsel]
        d[f] = a.tolist()
    with open(path, "w") as fh:
        json.dump(d, fh, indent=1)


def load_tables(path: str) -> RobotModel:
    with open(path) as fh:
        d = json.load(fh)
    if d.get("format") != "mpx_b200.model_tables.v1":
        raise ValueError(f"{path}: not an mpx_b200 model table")
    kw = {}
    for f in _ARRAY_FIELDS:
        kw[f] = np.array(d[f], dtype=np.int32 if f in _INT_FIELDS else np.float64)
    ng = len(d["geom_names"])
    kw["geom_pos"] = kw["geom_pos"].reshape(ng, 3)
    kw["jnt_pos"] = kw["jnt_pos"].reshape(d["njnt"], 3)
    kw["jnt_axis"] = kw["jnt_axis"].reshape(d["njnt"], 3)
    return RobotModel(nbody=d["nbody"], njnt=d["njnt"], nq=d["nq"], nv=d["nv"],
                      body_names=d["body_names"], jnt_names=d["jnt_names"], geom_names=d["geom_names"],
                      source=d.get("source", os.path.basename(path)), **kw)


def load_model(path: str) -> RobotModel:
    """``config.model_path`` entry point: ``.xml`` -> MJCF parser, ``.json`` -> model tables."""
    if path.endswith(".json"):
        return load_tables(path)
    return load_mjcf(path)
