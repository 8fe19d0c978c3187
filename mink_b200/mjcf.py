"""MJCF subset compiler: XML -> `Model` (MjModel-shaped arrays).

Setup-time host code.  mink itself never parses MJCF -- it receives a compiled
``mujoco.MjModel`` (reference ``mink/configuration.py:37-51``).  MuJoCo is not
installable in this environment, so the front end needs its own way to obtain
the kinematic tree.  This module compiles the subset of MJCF that the example
robots use into a `Model` whose field names follow MjModel, so
that the rest of the package is indifferent to whether it was handed a real
``mujoco.MjModel`` or one of these.

Supported: <include>, nested <default class> + childclass + class=, compiler
angle/autolimits/eulerseq, body pos/quat/euler/axisangle/xyaxes/zaxis, mocap,
<joint> hinge/slide/ball/free, <freejoint>, <inertial pos mass>, <geom>
(type/size/pos/quat/fromto/contype/conaffinity), <site>, <keyframe>.
Not supported (ignored): meshes, tendons, actuators, equality, mesh-derived
inertia (bodies without <inertial> get mass 0).
"""

from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

# Joint type codes follow mjtJoint (reference mink/constants.py:27-34).
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
_JNT_CODE = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}
QPOS_WIDTH = {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}
DOF_WIDTH = {JNT_FREE: 6, JNT_BALL: 3, JNT_SLIDE: 1, JNT_HINGE: 1}

# Geom type codes follow mjtGeom.
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID = 0, 1, 2, 3, 4
GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 5, 6, 7
_GEOM_CODE = {
    "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE,
    "capsule": GEOM_CAPSULE, "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER,
    "box": GEOM_BOX, "mesh": GEOM_MESH, "sdf": 8,
}

MJ_MAXVAL = 1e10
MJ_MINVAL = 1e-15


# --------------------------------------------------------------------------- #
# Small quaternion helpers (wxyz, Hamilton) used only at compile time.
# --------------------------------------------------------------------------- #
def _qmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def _qnormalize(q: np.ndarray) -> np.ndarray:
    n = float(np.linalg.norm(q))
    if n < MJ_MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def _quat_from_axis_angle(axis: np.ndarray, angle: float) -> np.ndarray:
    n = float(np.linalg.norm(axis))
    if n < MJ_MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    s = math.sin(0.5 * angle) / n
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def _quat_from_mat(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion (largest-component branch)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        w = 0.5 * math.sqrt(1.0 + t)
        k = 0.25 / w
        q = np.array([w, (R[2, 1] - R[1, 2]) * k, (R[0, 2] - R[2, 0]) * k, (R[1, 0] - R[0, 1]) * k])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        x = 0.5 * math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2])
        k = 0.25 / x
        q = np.array([(R[2, 1] - R[1, 2]) * k, x, (R[0, 1] + R[1, 0]) * k, (R[0, 2] + R[2, 0]) * k])
    elif R[1, 1] > R[2, 2]:
        y = 0.5 * math.sqrt(1.0 - R[0, 0] + R[1, 1] - R[2, 2])
        k = 0.25 / y
        q = np.array([(R[0, 2] - R[2, 0]) * k, (R[0, 1] + R[1, 0]) * k, y, (R[1, 2] + R[2, 1]) * k])
    else:
        z = 0.5 * math.sqrt(1.0 - R[0, 0] - R[1, 1] + R[2, 2])
        k = 0.25 / z
        q = np.array([(R[1, 0] - R[0, 1]) * k, (R[0, 2] + R[2, 0]) * k, (R[1, 2] + R[2, 1]) * k, z])
    return _qnormalize(q)


def _quat_z_to(vec: np.ndarray) -> np.ndarray:
    """Minimal rotation taking +z to `vec` (MJCF zaxis / fromto semantics)."""
    v = vec / max(float(np.linalg.norm(vec)), MJ_MINVAL)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, v)
    s = float(np.linalg.norm(axis))
    ang = math.atan2(s, float(v[2]))
    if s < MJ_MINVAL:
        axis = np.array([1.0, 0.0, 0.0])
    return _quat_from_axis_angle(axis, ang)


def _floats(text: Optional[str]) -> Optional[np.ndarray]:
    if text is None:
        return None
    return np.array([float(t) for t in text.split()], dtype=np.float64)


# --------------------------------------------------------------------------- #
# Named element views (model.body("x").id etc.; reference uses these accessors
# in mink/lie/se3.py:88, mink/utils.py:26, mink/limits/velocity_limit.py:52).
# --------------------------------------------------------------------------- #
class _ElemView:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Model:
    """Compiled kinematic model with MjModel-compatible field names."""

    # ---- construction -------------------------------------------------- #
    @classmethod
    def from_xml_path(cls, path: str) -> "Model":
        path = os.path.abspath(path)
        root = _load_with_includes(path)
        return _compile(root, os.path.dirname(path))

    @classmethod
    def from_xml_string(cls, xml: str, base_dir: Optional[str] = None) -> "Model":
        root = ET.fromstring(xml)
        base = base_dir or os.getcwd()
        _splice_includes(root, base)
        return _compile(root, base)

    # ---- named access --------------------------------------------------- #
    def _lookup(self, names: Sequence[str], key: Union[int, str], kind: str) -> int:
        if isinstance(key, (int, np.integer)):
            if not 0 <= int(key) < len(names):
                raise IndexError(f"Invalid {kind} index {key}")
            return int(key)
        try:
            return names.index(key)
        except ValueError:
            raise KeyError(f"Invalid name '{key}'. Valid names: {list(names)}") from None

    def name2id(self, kind: str, name: str) -> int:
        names = {"body": self.body_names, "joint": self.joint_names, "geom": self.geom_names,
                 "site": self.site_names, "key": self.key_names}[kind]
        try:
            return names.index(name)
        except ValueError:
            return -1

    def body(self, key):
        i = self._lookup(self.body_names, key, "body")
        return _ElemView(id=i, name=self.body_names[i], mocapid=self.body_mocapid[i:i + 1],
                         parentid=self.body_parentid[i:i + 1], pos=self.body_pos[i], quat=self.body_quat[i])

    def joint(self, key):
        i = self._lookup(self.joint_names, key, "joint")
        return _ElemView(id=i, name=self.joint_names[i], type=self.jnt_type[i:i + 1],
                         range=self.jnt_range[i], qposadr=self.jnt_qposadr[i:i + 1],
                         dofadr=self.jnt_dofadr[i:i + 1])

    def geom(self, key):
        i = self._lookup(self.geom_names, key, "geom")
        return _ElemView(id=i, name=self.geom_names[i], bodyid=self.geom_bodyid[i:i + 1])

    def site(self, key):
        i = self._lookup(self.site_names, key, "site")
        return _ElemView(id=i, name=self.site_names[i], bodyid=self.site_bodyid[i:i + 1])

    def key(self, key):
        i = self._lookup(self.key_names, key, "key")
        return _ElemView(id=i, name=self.key_names[i], qpos=self.key_qpos[i])


# --------------------------------------------------------------------------- #
# XML loading with <include> splicing.
# --------------------------------------------------------------------------- #
def _load_with_includes(path: str) -> ET.Element:
    root = ET.parse(path).getroot()
    _splice_includes(root, os.path.dirname(path))
    return root


def _splice_includes(elem: ET.Element, base_dir: str) -> None:
    i = 0
    while i < len(elem):
        child = elem[i]
        if child.tag == "include":
            inc_path = os.path.join(base_dir, child.attrib["file"])
            inc_root = ET.parse(inc_path).getroot()
            _splice_includes(inc_root, os.path.dirname(inc_path))
            elem.remove(child)
            for k, sub in enumerate(list(inc_root)):
                elem.insert(i + k, sub)
            i += len(inc_root)
        else:
            _splice_includes(child, base_dir)
            i += 1


# --------------------------------------------------------------------------- #
# Defaults.
# --------------------------------------------------------------------------- #
class _Defaults:
    """Default classes: class name -> {element tag -> attrib dict}, with parent chain resolved."""

    def __init__(self):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}

    def add_tree(self, elem: ET.Element, parent: Optional[str]) -> None:
        name = elem.attrib.get("class", "main" if parent is None else None)
        if name is None:
            raise ValueError("nested <default> requires a class attribute")
        base = {t: dict(a) for t, a in self.classes.get(parent, {}).items()} if parent else \
            {t: dict(a) for t, a in self.classes.get("main", {}).items()}
        if name in self.classes and name != "main":
            raise ValueError(f"repeated default class '{name}'")
        table = base
        for child in elem:
            if child.tag == "default":
                continue
            table.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = table
        for child in elem:
            if child.tag == "default":
                self.add_tree(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.attrib.get("class", childclass) or "main"
        if cls not in self.classes:
            raise ValueError(f"unknown default class '{cls}'")
        out = dict(self.classes[cls].get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


# --------------------------------------------------------------------------- #
# Compilation.
# --------------------------------------------------------------------------- #
class _Ctx:
    def __init__(self):
        self.angle_scale = math.pi / 180.0  # MJCF default is degrees
        self.autolimits = True              # MuJoCo >= 3.0 default
        self.eulerseq = "xyz"
        self.defaults = _Defaults()
        self.bodies: List[dict] = []
        self.joints: List[dict] = []
        self.geoms: List[dict] = []
        self.sites: List[dict] = []


def _orientation(attrib: Dict[str, str], ctx: _Ctx) -> np.ndarray:
    if "quat" in attrib:
        return _qnormalize(_floats(attrib["quat"]))
    if "axisangle" in attrib:
        v = _floats(attrib["axisangle"])
        return _quat_from_axis_angle(v[:3], v[3] * ctx.angle_scale)
    if "euler" in attrib:
        e = _floats(attrib["euler"]) * ctx.angle_scale
        q = np.array([1.0, 0.0, 0.0, 0.0])
        for ch, ang in zip(ctx.eulerseq, e):
            ax = {"x": [1.0, 0, 0], "y": [0, 1.0, 0], "z": [0, 0, 1.0]}[ch.lower()]
            r = _quat_from_axis_angle(np.array(ax), ang)
            # lower-case = intrinsic (rotating axes): post-multiply; upper = extrinsic.
            q = _qmul(q, r) if ch.islower() else _qmul(r, q)
        return _qnormalize(q)
    if "xyaxes" in attrib:
        v = _floats(attrib["xyaxes"])
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * float(x @ v[3:])
        y = y / np.linalg.norm(y)
        z = np.cross(x, y)
        return _quat_from_mat(np.stack([x, y, z], axis=1))
    if "zaxis" in attrib:
        return _quat_z_to(_floats(attrib["zaxis"]))
    return np.array([1.0, 0.0, 0.0, 0.0])


def _parse_bool(text: Optional[str], auto: Optional[bool] = None) -> Optional[bool]:
    if text is None or text == "auto":
        return auto
    return text == "true"


def _walk_body(elem: ET.Element, parent_id: int, childclass: Optional[str], ctx: _Ctx) -> None:
    battr = dict(elem.attrib)
    childclass = battr.get("childclass", childclass)
    bid = len(ctx.bodies)
    body = dict(
        name=battr.get("name", ""), parent=parent_id,
        pos=_floats(battr.get("pos", "0 0 0")), quat=_orientation(battr, ctx),
        mocap=battr.get("mocap", "false") == "true",
        ipos=np.zeros(3), mass=0.0, joints=[], geoms=[], sites=[], has_inertial=False,
    )
    ctx.bodies.append(body)
    for child in elem:
        tag = child.tag
        if tag == "inertial":
            body["has_inertial"] = True
            body["ipos"] = _floats(child.attrib.get("pos", "0 0 0"))
            body["mass"] = float(child.attrib.get("mass", "0"))
        elif tag in ("joint", "freejoint"):
            a = dict(child.attrib) if tag == "freejoint" else ctx.defaults.resolve("joint", child, childclass)
            jtype = JNT_FREE if tag == "freejoint" else _JNT_CODE[a.get("type", "hinge")]
            rng = _floats(a.get("range", "0 0"))
            if jtype == JNT_HINGE or jtype == JNT_BALL:
                rng = rng * ctx.angle_scale
            limited = _parse_bool(a.get("limited"), None)
            if limited is None:
                limited = bool(ctx.autolimits and "range" in a)
            if jtype == JNT_FREE:
                limited = False
            axis = _floats(a.get("axis", "0 0 1"))
            n = float(np.linalg.norm(axis))
            axis = axis / n if n > MJ_MINVAL else np.array([0.0, 0.0, 1.0])
            ref = float(a.get("ref", "0"))
            if jtype == JNT_HINGE:
                ref *= ctx.angle_scale
            ctx.joints.append(dict(
                name=a.get("name", ""), body=bid, type=jtype, axis=axis,
                pos=_floats(a.get("pos", "0 0 0")), range=rng, limited=limited, ref=ref))
            body["joints"].append(len(ctx.joints) - 1)
        elif tag == "geom":
            a = ctx.defaults.resolve("geom", child, childclass)
            gtype = _GEOM_CODE[a.get("type", "sphere")]
            size = np.zeros(3)
            s = _floats(a.get("size"))
            if s is not None:
                size[: len(s)] = s[:3]
            pos = _floats(a.get("pos", "0 0 0"))
            quat = _orientation(a, ctx)
            if "fromto" in a:
                ft = _floats(a["fromto"])
                pos = 0.5 * (ft[:3] + ft[3:])
                quat = _quat_z_to(ft[3:] - ft[:3])
                size[1] = 0.5 * float(np.linalg.norm(ft[3:] - ft[:3]))
            ctx.geoms.append(dict(
                name=a.get("name", ""), body=bid, type=gtype, size=size, pos=pos, quat=quat,
                contype=int(a.get("contype", "1")), conaffinity=int(a.get("conaffinity", "1"))))
            body["geoms"].append(len(ctx.geoms) - 1)
        elif tag == "site":
            a = ctx.defaults.resolve("site", child, childclass)
            ctx.sites.append(dict(name=a.get("name", ""), body=bid,
                                  pos=_floats(a.get("pos", "0 0 0")), quat=_orientation(a, ctx)))
            body["sites"].append(len(ctx.sites) - 1)
    for child in elem:
        if child.tag == "body":
            _walk_body(child, bid, childclass, ctx)


def _compile(root: ET.Element, base_dir: str) -> Model:
    ctx = _Ctx()
    for comp in root.iter("compiler"):
        if "angle" in comp.attrib:
            ctx.angle_scale = 1.0 if comp.attrib["angle"] == "radian" else math.pi / 180.0
        if "autolimits" in comp.attrib:
            ctx.autolimits = comp.attrib["autolimits"] == "true"
        if "eulerseq" in comp.attrib:
            ctx.eulerseq = comp.attrib["eulerseq"]
    for d in root.findall("default"):
        ctx.defaults.add_tree(d, None)

    # World body = id 0; all <worldbody> sections merge into it.
    world = dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mocap=False,
                 ipos=np.zeros(3), mass=0.0, joints=[], geoms=[], sites=[])
    ctx.bodies.append(world)
    for wb in root.findall("worldbody"):
        for child in wb:
            if child.tag == "geom":
                a = ctx.defaults.resolve("geom", child, None)
                size = np.zeros(3)
                s = _floats(a.get("size"))
                if s is not None:
                    size[: len(s)] = s[:3]
                ctx.geoms.append(dict(name=a.get("name", ""), body=0, type=_GEOM_CODE[a.get("type", "sphere")],
                                      size=size, pos=_floats(a.get("pos", "0 0 0")), quat=_orientation(a, ctx),
                                      contype=int(a.get("contype", "1")), conaffinity=int(a.get("conaffinity", "1"))))
                world["geoms"].append(len(ctx.geoms) - 1)
            elif child.tag == "site":
                a = ctx.defaults.resolve("site", child, None)
                ctx.sites.append(dict(name=a.get("name", ""), body=0,
                                      pos=_floats(a.get("pos", "0 0 0")), quat=_orientation(a, ctx)))
                world["sites"].append(len(ctx.sites) - 1)
    for wb in root.findall("worldbody"):
        for child in wb:
            if child.tag == "body":
                _walk_body(child, 0, None, ctx)

    # The parser appends joints/geoms/sites in document order *per body visit*, which for a
    # DFS pre-order walk equals body order, except world geoms/sites which were added first (body 0).
    m = Model()
    nb = len(ctx.bodies)
    m.nbody = nb
    m.body_names = [b["name"] for b in ctx.bodies]
    m.body_parentid = np.array([b["parent"] for b in ctx.bodies], dtype=np.int32)
    m.body_pos = np.array([b["pos"] for b in ctx.bodies], dtype=np.float64).reshape(nb, 3)
    m.body_quat = np.array([b["quat"] for b in ctx.bodies], dtype=np.float64).reshape(nb, 4)
    m.body_ipos = np.array([b["ipos"] for b in ctx.bodies], dtype=np.float64).reshape(nb, 3)
    m.body_mass = np.array([b["mass"] for b in ctx.bodies], dtype=np.float64)
    # MuJoCo derives mass and inertia from the geoms of a body that has no <inertial>; this compiler does not (no meshes, no
    # geom densities), so such bodies carry mass 0 here.  Remember them: a ComTask on such a model must be refused, not wrong.
    m.body_mass_missing = np.array([(not b.get("has_inertial", False)) and len(b["geoms"]) > 0 and not b.get("mocap", False)
                                    for b in ctx.bodies], dtype=bool)
    mocapid = np.full(nb, -1, dtype=np.int32)
    nmocap = 0
    for i, b in enumerate(ctx.bodies):
        if b["mocap"]:
            mocapid[i] = nmocap
            nmocap += 1
    m.body_mocapid = mocapid
    m.nmocap = nmocap

    nj = len(ctx.joints)
    m.njnt = nj
    m.joint_names = [j["name"] for j in ctx.joints]
    m.jnt_type = np.array([j["type"] for j in ctx.joints], dtype=np.int32)
    m.jnt_bodyid = np.array([j["body"] for j in ctx.joints], dtype=np.int32)
    m.jnt_axis = np.array([j["axis"] for j in ctx.joints], dtype=np.float64).reshape(nj, 3)
    m.jnt_pos = np.array([j["pos"] for j in ctx.joints], dtype=np.float64).reshape(nj, 3)
    m.jnt_range = np.array([j["range"] for j in ctx.joints], dtype=np.float64).reshape(nj, 2)
    m.jnt_limited = np.array([j["limited"] for j in ctx.joints], dtype=bool)
    qadr = np.zeros(nj, dtype=np.int32)
    dadr = np.zeros(nj, dtype=np.int32)
    nq = nv = 0
    for k, j in enumerate(ctx.joints):
        qadr[k], dadr[k] = nq, nv
        nq += QPOS_WIDTH[j["type"]]
        nv += DOF_WIDTH[j["type"]]
    m.jnt_qposadr, m.jnt_dofadr, m.nq, m.nv = qadr, dadr, nq, nv

    m.body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in ctx.bodies], dtype=np.int32)
    m.body_jntnum = np.array([len(b["joints"]) for b in ctx.bodies], dtype=np.int32)
    m.body_dofnum = np.array([sum(DOF_WIDTH[ctx.joints[j]["type"]] for j in b["joints"]) for b in ctx.bodies],
                             dtype=np.int32)
    m.body_dofadr = np.array([dadr[b["joints"][0]] if b["joints"] else -1 for b in ctx.bodies], dtype=np.int32)

    # dof tables.
    dof_body = np.zeros(nv, dtype=np.int32)
    dof_jnt = np.zeros(nv, dtype=np.int32)
    for k, j in enumerate(ctx.joints):
        for d in range(DOF_WIDTH[j["type"]]):
            dof_body[dadr[k] + d] = j["body"]
            dof_jnt[dadr[k] + d] = k
    m.dof_bodyid, m.dof_jntid = dof_body, dof_jnt
    # dof_parentid: previous dof on the same body, else last dof of the nearest ancestor with dofs.
    last_dof_of_body = np.full(nb, -1, dtype=np.int32)
    for i in range(nb):
        if m.body_dofnum[i] > 0:
            last_dof_of_body[i] = m.body_dofadr[i] + m.body_dofnum[i] - 1
    dof_parent = np.full(nv, -1, dtype=np.int32)
    for d in range(nv):
        b = dof_body[d]
        if d > m.body_dofadr[b]:
            dof_parent[d] = d - 1
        else:
            p = m.body_parentid[b]
            while p != 0 and last_dof_of_body[p] < 0:
                p = m.body_parentid[p]
            dof_parent[d] = last_dof_of_body[p] if p != 0 else -1
    m.dof_parentid = dof_parent

    # weld / root ids.
    weld = np.zeros(nb, dtype=np.int32)
    rootid = np.zeros(nb, dtype=np.int32)
    for i in range(1, nb):
        p = m.body_parentid[i]
        weld[i] = i if m.body_jntnum[i] > 0 else weld[p]
        rootid[i] = i if p == 0 else rootid[p]
    m.body_weldid, m.body_rootid = weld, rootid

    # geoms / sites. Re-sort into body order (world entries first already; stable by body id).
    gorder = sorted(range(len(ctx.geoms)), key=lambda g: ctx.geoms[g]["body"])
    geoms = [ctx.geoms[g] for g in gorder]
    ng = len(geoms)
    m.ngeom = ng
    m.geom_names = [g["name"] for g in geoms]
    m.geom_bodyid = np.array([g["body"] for g in geoms], dtype=np.int32)
    m.geom_type = np.array([g["type"] for g in geoms], dtype=np.int32)
    m.geom_size = np.array([g["size"] for g in geoms], dtype=np.float64).reshape(ng, 3)
    m.geom_pos = np.array([g["pos"] for g in geoms], dtype=np.float64).reshape(ng, 3)
    m.geom_quat = np.array([g["quat"] for g in geoms], dtype=np.float64).reshape(ng, 4)
    m.geom_contype = np.array([g["contype"] for g in geoms], dtype=np.int32)
    m.geom_conaffinity = np.array([g["conaffinity"] for g in geoms], dtype=np.int32)
    m.body_geomnum = np.array([int(np.sum(m.geom_bodyid == i)) for i in range(nb)], dtype=np.int32)
    m.body_geomadr = np.array([int(np.argmax(m.geom_bodyid == i)) if m.body_geomnum[i] else -1
                               for i in range(nb)], dtype=np.int32)

    sorder = sorted(range(len(ctx.sites)), key=lambda s: ctx.sites[s]["body"])
    sites = [ctx.sites[s] for s in sorder]
    ns = len(sites)
    m.nsite = ns
    m.site_names = [s["name"] for s in sites]
    m.site_bodyid = np.array([s["body"] for s in sites], dtype=np.int32)
    m.site_pos = np.array([s["pos"] for s in sites], dtype=np.float64).reshape(ns, 3)
    m.site_quat = np.array([s["quat"] for s in sites], dtype=np.float64).reshape(ns, 4)

    # qpos0: free joint takes the body frame, ball identity, scalar joints `ref`.
    qpos0 = np.zeros(nq)
    for k, j in enumerate(ctx.joints):
        a = qadr[k]
        if j["type"] == JNT_FREE:
            b = ctx.bodies[j["body"]]
            qpos0[a:a + 3] = b["pos"]
            qpos0[a + 3:a + 7] = b["quat"]
        elif j["type"] == JNT_BALL:
            qpos0[a:a + 4] = [1.0, 0.0, 0.0, 0.0]
        else:
            qpos0[a] = j["ref"]
    m.qpos0 = qpos0

    # subtree masses.
    sub = m.body_mass.copy()
    for i in range(nb - 1, 0, -1):
        sub[m.body_parentid[i]] += sub[i]
    m.body_subtreemass = sub

    # keyframes.
    keys = [k for kf in root.findall("keyframe") for k in kf.findall("key")]
    m.nkey = len(keys)
    m.key_names = [k.attrib.get("name", "") for k in keys]
    key_qpos = np.tile(qpos0, (max(len(keys), 1), 1))[: len(keys)].copy() if keys else np.zeros((0, nq))
    for i, k in enumerate(keys):
        if "qpos" in k.attrib:
            v = _floats(k.attrib["qpos"])
            if len(v) != nq:
                raise ValueError(f"keyframe {i}: expected {nq} qpos values, got {len(v)}")
            key_qpos[i] = v
    m.key_qpos = key_qpos

    # mocap initial poses (MjData.mocap_pos/quat are initialised from these).
    m.mocap_pos0 = np.array([ctx.bodies[i]["pos"] for i in range(nb) if mocapid[i] >= 0]).reshape(nmocap, 3)
    m.mocap_quat0 = np.array([ctx.bodies[i]["quat"] for i in range(nb) if mocapid[i] >= 0]).reshape(nmocap, 4)
    return m
