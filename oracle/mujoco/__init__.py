"""ORACLE / TEST INFRASTRUCTURE -- numpy restatement of the `mujoco` calls mink makes.

This is NOT the product and is never imported by `mink_b200/`.  It exists so that the
*unmodified* reference package at /root/reference/mink can be imported in this container
(where the real `mujoco` wheel is not installable) and used to (a) run the reference's own
property tests against this restatement and (b) generate golden vectors for tests/golden/.

Third-party dependency restated: `mujoco >= 3.1.6` (reference pyproject.toml:26-29); its C
source is not under /root/reference.  The restatement follows MuJoCo's published algorithms
(engine_core_smooth.c: mj_kinematics, mj_comPos; engine_support.c: mj_jac, mj_jacSubtreeCom,
mj_integratePos, mj_differentiatePos; engine_util_spatial.c: mju_mat2Quat, mju_quat2Mat,
mju_mulQuat, mju_quatIntegrate, mju_subQuat) and is anchored on every reference call site:

  mj_kinematics / mj_comPos ........ mink/configuration.py:63-64
  mj_jacBody/Geom/Site ............. mink/constants.py:10-14 -> mink/configuration.py:144-145
  mj_integratePos .................. mink/configuration.py:225,235
  mj_differentiatePos .............. mink/tasks/posture_task.py:107, mink/limits/configuration_limit.py:100,110
  mj_jacSubtreeCom, subtree_com .... mink/tasks/com_task.py:69,82,96
  mj_geomDistance, mj_jac .......... mink/limits/collision_avoidance_limit.py:219,69,71
  mju_mat2Quat/quat2Mat/mulQuat .... mink/lie/so3.py:83,113,150
  mj_name2id ....................... mink/configuration.py:72,133,170

PARITY PINNING: the reference's tests hold no absolute golden numbers for this path
(SURVEY.md 8c).  The restatement is pinned by running the reference's own tests
(tests/test_jacobians.py finite differences, tests/test_solve_ik.py convergence,
tests/test_configuration*.py, tests/test_lie_*.py ...) on top of it: see oracle/run_reference_tests.py.
`mj_geomDistance` is restated only for plane/sphere/capsule pairs and is NOT pinned by any
reference test that can run here (tests/test_collision_avoidance_limit.py:65-111 needs real MuJoCo).

Model loading (MJCF -> arrays) is delegated to the product's setup-time compiler
`mink_b200.mjcf`; all per-step arithmetic below is independent of the product.
"""

from __future__ import annotations

import enum
import os
import sys

import numpy as np

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from mink_b200 import mjcf as _mjcf  # noqa: E402  (setup-time model compiler only)

mjMAXVAL = 1e10
mjMINVAL = 1e-15


class mjtJoint(enum.IntEnum):
    mjJNT_FREE = 0
    mjJNT_BALL = 1
    mjJNT_SLIDE = 2
    mjJNT_HINGE = 3


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_XBODY = 2
    mjOBJ_JOINT = 3
    mjOBJ_DOF = 4
    mjOBJ_GEOM = 5
    mjOBJ_SITE = 6
    mjOBJ_KEY = 24


class mjtGeom(enum.IntEnum):
    mjGEOM_PLANE = 0
    mjGEOM_HFIELD = 1
    mjGEOM_SPHERE = 2
    mjGEOM_CAPSULE = 3
    mjGEOM_ELLIPSOID = 4
    mjGEOM_CYLINDER = 5
    mjGEOM_BOX = 6
    mjGEOM_MESH = 7


class MjModel(_mjcf.Model):
    """MjModel stand-in.  Arrays are produced by mink_b200.mjcf (shared setup-time compiler)."""

    @classmethod
    def from_xml_path(cls, path: str) -> "MjModel":
        m = _mjcf.Model.from_xml_path(path)
        m.__class__ = cls
        return m

    @classmethod
    def from_xml_string(cls, xml: str, assets=None) -> "MjModel":
        m = _mjcf.Model.from_xml_string(xml)
        m.__class__ = cls
        return m


class _DataElem:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class MjData:
    """MjData stand-in holding only the fields mink reads."""

    def __init__(self, model: MjModel):
        object.__setattr__(self, "_model", model)
        m = model
        self._qpos = m.qpos0.copy()
        self.qvel = np.zeros(m.nv)
        self.xpos = np.zeros((m.nbody, 3))
        self.xquat = np.tile([1.0, 0, 0, 0], (m.nbody, 1))
        self.xmat = np.tile(np.eye(3).ravel(), (m.nbody, 1))
        self.xipos = np.zeros((m.nbody, 3))
        self.xanchor = np.zeros((m.njnt, 3))
        self.xaxis = np.zeros((m.njnt, 3))
        self.geom_xpos = np.zeros((m.ngeom, 3))
        self.geom_xmat = np.tile(np.eye(3).ravel(), (m.ngeom, 1))
        self.site_xpos = np.zeros((m.nsite, 3))
        self.site_xmat = np.tile(np.eye(3).ravel(), (m.nsite, 1))
        self.subtree_com = np.zeros((m.nbody, 3))
        self.cdof = np.zeros((m.nv, 6))
        self.mocap_pos = m.mocap_pos0.copy()
        self.mocap_quat = m.mocap_quat0.copy()

    # `data.qpos = q` must copy INTO the buffer (real MuJoCo semantics).
    @property
    def qpos(self):
        return self._qpos

    @qpos.setter
    def qpos(self, value):
        self._qpos[:] = np.asarray(value, dtype=np.float64)

    def body(self, key):
        i = self._model._lookup(self._model.body_names, key, "body")
        return _DataElem(id=i, name=self._model.body_names[i], xpos=self.xpos[i], xquat=self.xquat[i],
                         xmat=self.xmat[i], xipos=self.xipos[i], subtree_com=self.subtree_com[i])

    def site(self, key):
        i = self._model._lookup(self._model.site_names, key, "site")
        return _DataElem(id=i, name=self._model.site_names[i], xpos=self.site_xpos[i], xmat=self.site_xmat[i])

    def geom(self, key):
        i = self._model._lookup(self._model.geom_names, key, "geom")
        return _DataElem(id=i, name=self._model.geom_names[i], xpos=self.geom_xpos[i], xmat=self.geom_xmat[i])


# --------------------------------------------------------------------------- #
# mju_* helpers
# --------------------------------------------------------------------------- #
def _mulquat(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def _normquat(q):
    n = np.sqrt(q @ q)
    if n < mjMINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def _quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def _axisangle2quat(axis, angle):
    if angle == 0.0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def mju_mulQuat(res, a, b):
    res[:] = _mulquat(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))


def mju_quat2Mat(res, quat):
    res[:] = _quat2mat(np.asarray(quat, dtype=np.float64)).ravel()


def mju_mat2Quat(quat, mat):
    """Row-major 3x3 -> quaternion, largest-diagonal branch selection, then normalise."""
    m = np.asarray(mat, dtype=np.float64).reshape(3, 3)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if tr > 0:
        q[0] = 0.5 * np.sqrt(1 + tr)
        k = 0.25 / q[0]
        q[1] = k * (m[2, 1] - m[1, 2])
        q[2] = k * (m[0, 2] - m[2, 0])
        q[3] = k * (m[1, 0] - m[0, 1])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        q[1] = 0.5 * np.sqrt(1 + m[0, 0] - m[1, 1] - m[2, 2])
        k = 0.25 / q[1]
        q[0] = k * (m[2, 1] - m[1, 2])
        q[2] = k * (m[1, 0] + m[0, 1])
        q[3] = k * (m[0, 2] + m[2, 0])
    elif m[1, 1] > m[2, 2]:
        q[2] = 0.5 * np.sqrt(1 - m[0, 0] + m[1, 1] - m[2, 2])
        k = 0.25 / q[2]
        q[0] = k * (m[0, 2] - m[2, 0])
        q[1] = k * (m[1, 0] + m[0, 1])
        q[3] = k * (m[2, 1] + m[1, 2])
    else:
        q[3] = 0.5 * np.sqrt(1 - m[0, 0] - m[1, 1] + m[2, 2])
        k = 0.25 / q[3]
        q[0] = k * (m[1, 0] - m[0, 1])
        q[1] = k * (m[0, 2] + m[2, 0])
        q[2] = k * (m[2, 1] + m[1, 2])
    quat[:] = _normquat(q)


def mju_normalize3(vec):
    n = np.sqrt(vec @ vec)
    if n < mjMINVAL:
        vec[:] = [1.0, 0.0, 0.0]
        return 0.0
    vec /= n
    return float(n)


def _quat_integrate(q, w, dt):
    """quat <- normalize(quat) * axisangle(w/|w|, dt|w|), renormalised (mju_quatIntegrate)."""
    q = _normquat(np.asarray(q, dtype=np.float64))
    n = np.sqrt(w @ w)
    if n < mjMINVAL:
        return q
    return _normquat(_mulquat(q, _axisangle2quat(w / n, dt * n)))


def _sub_quat(qa, qb):
    """Body-frame rotation vector taking qb to qa: vel(conj(qb) * qa) (mju_subQuat)."""
    qb_conj = np.array([qb[0], -qb[1], -qb[2], -qb[3]])
    d = _mulquat(qb_conj, qa)
    axis = d[1:].copy()
    s = np.sqrt(axis @ axis)
    if s < mjMINVAL:
        return np.zeros(3)
    axis /= s
    ang = 2.0 * np.arctan2(s, d[0])
    if ang > np.pi:
        ang -= 2.0 * np.pi
    return axis * ang


# --------------------------------------------------------------------------- #
# Kinematics
# --------------------------------------------------------------------------- #
def mj_name2id(model, objtype, name):
    kind = {mjtObj.mjOBJ_BODY: "body", mjtObj.mjOBJ_JOINT: "joint", mjtObj.mjOBJ_GEOM: "geom",
            mjtObj.mjOBJ_SITE: "site", mjtObj.mjOBJ_KEY: "key"}[mjtObj(int(objtype))]
    return model.name2id(kind, name)


def mj_kinematics(m, d):
    d.xpos[0] = 0.0
    d.xquat[0] = [1.0, 0.0, 0.0, 0.0]
    d.xmat[0] = np.eye(3).ravel()
    d.xipos[0] = 0.0
    for i in range(1, m.nbody):
        jadr, jnum = m.body_jntadr[i], m.body_jntnum[i]
        if jnum == 1 and m.jnt_type[jadr] == mjtJoint.mjJNT_FREE:
            a = m.jnt_qposadr[jadr]
            xpos = d.qpos[a:a + 3].copy()
            xquat = _normquat(d.qpos[a + 3:a + 7].copy())
            d.qpos[a + 3:a + 7] = xquat  # mj_kinematics normalises the quaternion in place
            d.xanchor[jadr] = xpos
            d.xaxis[jadr] = [0.0, 0.0, 1.0]
        else:
            mid = m.body_mocapid[i]
            p = m.body_parentid[i]
            if mid >= 0:
                xpos = d.mocap_pos[mid].copy()
                xquat = _normquat(d.mocap_quat[mid].copy())
            else:
                Rp = d.xmat[p].reshape(3, 3)
                xpos = d.xpos[p] + Rp @ m.body_pos[i]
                xquat = _mulquat(d.xquat[p], m.body_quat[i])
            for j in range(jadr, jadr + jnum):
                R = _quat2mat(xquat)
                d.xanchor[j] = xpos + R @ m.jnt_pos[j]
                d.xaxis[j] = R @ m.jnt_axis[j]
                a = m.jnt_qposadr[j]
                t = m.jnt_type[j]
                if t == mjtJoint.mjJNT_SLIDE:
                    xpos = xpos + d.xaxis[j] * (d.qpos[a] - m.qpos0[a])
                elif t == mjtJoint.mjJNT_HINGE:
                    xquat = _mulquat(xquat, _axisangle2quat(m.jnt_axis[j], d.qpos[a] - m.qpos0[a]))
                    xpos = d.xanchor[j] - _quat2mat(xquat) @ m.jnt_pos[j]
                elif t == mjtJoint.mjJNT_BALL:
                    qloc = _normquat(d.qpos[a:a + 4].copy())
                    d.qpos[a:a + 4] = qloc
                    xquat = _mulquat(xquat, qloc)
                    xpos = d.xanchor[j] - _quat2mat(xquat) @ m.jnt_pos[j]
                else:
                    raise ValueError("free joint must be the only joint of a body")
        xquat = _normquat(xquat)
        d.xpos[i] = xpos
        d.xquat[i] = xquat
        R = _quat2mat(xquat)
        d.xmat[i] = R.ravel()
        d.xipos[i] = xpos + R @ m.body_ipos[i]
    for g in range(m.ngeom):
        b = m.geom_bodyid[g]
        R = d.xmat[b].reshape(3, 3)
        d.geom_xpos[g] = d.xpos[b] + R @ m.geom_pos[g]
        d.geom_xmat[g] = _quat2mat(_normquat(_mulquat(d.xquat[b], m.geom_quat[g]))).ravel()
    for s in range(m.nsite):
        b = m.site_bodyid[s]
        R = d.xmat[b].reshape(3, 3)
        d.site_xpos[s] = d.xpos[b] + R @ m.site_pos[s]
        d.site_xmat[s] = _quat2mat(_normquat(_mulquat(d.xquat[b], m.site_quat[s]))).ravel()


def mj_comPos(m, d):
    """subtree_com by backward accumulation; cdof = [omega; v] about the root subtree CoM."""
    acc = d.xipos * m.body_mass[:, None]
    for i in range(m.nbody - 1, 0, -1):
        acc[m.body_parentid[i]] += acc[i]
    for i in range(m.nbody):
        if m.body_subtreemass[i] < mjMINVAL:
            d.subtree_com[i] = d.xipos[i]
        else:
            d.subtree_com[i] = acc[i] / m.body_subtreemass[i]
    for j in range(m.njnt):
        b = m.jnt_bodyid[j]
        c = d.subtree_com[m.body_rootid[b]]
        off = c - d.xanchor[j]
        da = m.jnt_dofadr[j]
        t = m.jnt_type[j]
        if t == mjtJoint.mjJNT_FREE:
            for k in range(3):
                d.cdof[da + k] = 0.0
                d.cdof[da + k, 3 + k] = 1.0
            R = d.xmat[b].reshape(3, 3)
            for k in range(3):
                ax = R[:, k]
                d.cdof[da + 3 + k, :3] = ax
                d.cdof[da + 3 + k, 3:] = np.cross(ax, off)
        elif t == mjtJoint.mjJNT_BALL:
            R = d.xmat[b].reshape(3, 3)
            for k in range(3):
                ax = R[:, k]
                d.cdof[da + k, :3] = ax
                d.cdof[da + k, 3:] = np.cross(ax, off)
        elif t == mjtJoint.mjJNT_SLIDE:
            d.cdof[da, :3] = 0.0
            d.cdof[da, 3:] = d.xaxis[j]
        else:
            d.cdof[da, :3] = d.xaxis[j]
            d.cdof[da, 3:] = np.cross(d.xaxis[j], off)


def mj_forward(m, d):
    mj_kinematics(m, d)
    mj_comPos(m, d)


def mj_fwdPosition(m, d):
    mj_forward(m, d)


def mj_jac(m, d, jacp, jacr, point, body):
    point = np.asarray(point, dtype=np.float64)
    if jacp is not None:
        jacp[:] = 0.0
    if jacr is not None:
        jacr[:] = 0.0
    body = int(body)
    while body and m.body_dofnum[body] == 0:
        body = int(m.body_parentid[body])
    if not body:
        return
    off = point - d.subtree_com[m.body_rootid[body]]
    i = int(m.body_dofadr[body] + m.body_dofnum[body] - 1)
    while i >= 0:
        if jacr is not None:
            jacr[:, i] = d.cdof[i, :3]
        if jacp is not None:
            jacp[:, i] = d.cdof[i, 3:] + np.cross(d.cdof[i, :3], off)
        i = int(m.dof_parentid[i])


def mj_jacBody(m, d, jacp, jacr, body):
    mj_jac(m, d, jacp, jacr, d.xpos[body], body)


def mj_jacBodyCom(m, d, jacp, jacr, body):
    mj_jac(m, d, jacp, jacr, d.xipos[body], body)


def mj_jacGeom(m, d, jacp, jacr, geom):
    mj_jac(m, d, jacp, jacr, d.geom_xpos[geom], m.geom_bodyid[geom])


def mj_jacSite(m, d, jacp, jacr, site):
    mj_jac(m, d, jacp, jacr, d.site_xpos[site], m.site_bodyid[site])


def mj_jacSubtreeCom(m, d, jacp, body):
    jacp[:] = 0.0
    tmp = np.zeros((3, m.nv))
    for b in range(m.nbody - 1, int(body) - 1, -1):
        # is b inside subtree(body)?
        a = b
        while a > body:
            a = int(m.body_parentid[a])
        if a != body:
            continue
        mj_jac(m, d, tmp, None, d.xipos[b], b)
        jacp += tmp * m.body_mass[b]
    jacp /= m.body_subtreemass[body]


def mj_integratePos(m, qpos, qvel, dt):
    qvel = np.asarray(qvel, dtype=np.float64)
    for j in range(m.njnt):
        a, v = m.jnt_qposadr[j], m.jnt_dofadr[j]
        t = m.jnt_type[j]
        if t == mjtJoint.mjJNT_FREE:
            qpos[a:a + 3] += dt * qvel[v:v + 3]
            qpos[a + 3:a + 7] = _quat_integrate(qpos[a + 3:a + 7], qvel[v + 3:v + 6], dt)
        elif t == mjtJoint.mjJNT_BALL:
            qpos[a:a + 4] = _quat_integrate(qpos[a:a + 4], qvel[v:v + 3], dt)
        else:
            qpos[a] += dt * qvel[v]


def mj_differentiatePos(m, qvel, dt, qpos1, qpos2):
    with np.errstate(all="ignore"):
        for j in range(m.njnt):
            a, v = m.jnt_qposadr[j], m.jnt_dofadr[j]
            t = m.jnt_type[j]
            if t == mjtJoint.mjJNT_FREE:
                qvel[v:v + 3] = (qpos2[a:a + 3] - qpos1[a:a + 3]) / dt
                qvel[v + 3:v + 6] = _sub_quat(qpos2[a + 3:a + 7], qpos1[a + 3:a + 7]) / dt
            elif t == mjtJoint.mjJNT_BALL:
                qvel[v:v + 3] = _sub_quat(qpos2[a:a + 4], qpos1[a:a + 4]) / dt
            else:
                qvel[v] = (qpos2[a] - qpos1[a]) / dt


def mj_resetData(m, d):
    d.qpos = m.qpos0
    d.qvel[:] = 0.0
    d.mocap_pos[:] = m.mocap_pos0
    d.mocap_quat[:] = m.mocap_quat0


def mj_resetDataKeyframe(m, d, key):
    mj_resetData(m, d)
    d.qpos = m.key_qpos[key]


# --------------------------------------------------------------------------- #
# Geom distance (primitive pairs only).
# --------------------------------------------------------------------------- #
def _seg_closest(p1, d1, p2, d2):
    """Closest points between segments p1 + s*d1, p2 + t*d2 with s,t in [-1, 1]."""
    a, e, f = d1 @ d1, d2 @ d2, None
    r = p1 - p2
    b = d1 @ d2
    c = d1 @ r
    f = d2 @ r
    den = a * e - b * b
    s = np.clip((b * f - c * e) / den, -1.0, 1.0) if den > 1e-14 else 0.0
    t = (b * s + f) / e if e > 1e-14 else 0.0
    if t < -1.0 or t > 1.0:
        t = np.clip(t, -1.0, 1.0)
        s = np.clip((b * t - c) / a, -1.0, 1.0) if a > 1e-14 else 0.0
    return p1 + s * d1, p2 + t * d2


def _primitive_core(m, d, g):
    """Return (kind, centre-or-segment, radius) reducing sphere/capsule to a point/segment + radius."""
    t = m.geom_type[g]
    pos = d.geom_xpos[g]
    R = d.geom_xmat[g].reshape(3, 3)
    if t == mjtGeom.mjGEOM_SPHERE:
        return "pt", (pos, np.zeros(3)), m.geom_size[g, 0]
    if t == mjtGeom.mjGEOM_CAPSULE:
        return "seg", (pos, R[:, 2] * m.geom_size[g, 1]), m.geom_size[g, 0]
    if t == mjtGeom.mjGEOM_PLANE:
        return "plane", (pos, R[:, 2]), 0.0
    if t == mjtGeom.mjGEOM_BOX:
        return "box", (pos, R, np.array(m.geom_size[g], dtype=np.float64)), 0.0
    raise NotImplementedError(f"oracle mj_geomDistance: geom type {int(t)} not restated")


def _seg_box_param(c, d, s):
    """t in [-1, 1] minimising the squared distance from c + t d to the box |x_k| <= s_k (box frame).

    The squared distance is piecewise quadratic in t with breakpoints where a coordinate crosses a face plane:
    minimise each piece in closed form and keep the best (exact, no iteration)."""
    def f(t):
        p = c + t * d
        e = np.maximum(np.abs(p) - s, 0.0)
        return float(e @ e)

    # segment through the box: the middle of the part inside (the distance is zero on that whole interval)
    t0, t1, hit = -1.0, 1.0, True
    for k in range(3):
        if d[k] == 0.0:
            hit = hit and abs(c[k]) <= s[k]
        else:
            a, b = sorted(((-s[k] - c[k]) / d[k], (s[k] - c[k]) / d[k]))
            t0, t1 = max(t0, a), min(t1, b)
    if hit and t0 <= t1:
        return 0.5 * (t0 + t1)

    cuts = [-1.0, 1.0]
    for k in range(3):
        if abs(d[k]) > 0.0:
            for sg in (-1.0, 1.0):
                t = (sg * s[k] - c[k]) / d[k]
                if -1.0 < t < 1.0:
                    cuts.append(float(t))
    cuts = sorted(set(cuts))
    best_t, best_f = -1.0, f(-1.0)
    for a, b in zip(cuts[:-1], cuts[1:]):
        pm = c + 0.5 * (a + b) * d
        num = den = 0.0
        for k in range(3):
            if abs(pm[k]) > s[k]:   # coordinate k is outside on this piece: term (sg*(c_k + t d_k) - s_k)^2
                sg = np.sign(pm[k])
                num += sg * d[k] * (sg * c[k] - s[k])
                den += d[k] * d[k]
        t = float(np.clip(-num / den, a, b)) if den > 0.0 else 0.5 * (a + b)
        for cand in (t, b):
            fc = f(cand)
            if fc < best_f:
                best_t, best_f = cand, fc
    return best_t


def _box_distance(box, k2, c2, r2):
    """Box against a sphere / capsule core (point or segment + radius): (dist, point on box, point on the other geom)."""
    pos, R, s = box
    c = R.T @ (c2[0] - pos)
    dd = R.T @ c2[1]
    t = _seg_box_param(c, dd, s) if k2 == "seg" else 0.0
    p = c + t * dd
    qc = np.clip(p, -s, s)
    v = p - qc
    L = float(np.sqrt(v @ v))
    if L > mjMINVAL:
        n = v / L
    else:   # core point inside the box: leave through the nearest face
        gap = s - np.abs(p)
        k = int(np.argmin(gap))
        n = np.zeros(3)
        n[k] = -1.0 if p[k] < 0 else 1.0
        L = -float(gap[k])
        qc = p - L * n
    return L - r2, pos + R @ qc, pos + R @ (p - r2 * n)


def mj_geomDistance(m, d, geom1, geom2, distmax, fromto):
    """Signed distance between two primitive geoms; `distmax` and zero fromto when farther.

    Box pairs (box vs plane / sphere / capsule) restate what MuJoCo's convex-distance routine returns for non-penetrating
    shapes (the witness points of the minimum distance); MuJoCo itself is not available to pin them against."""
    k1, c1, r1 = _primitive_core(m, d, geom1)
    k2, c2, r2 = _primitive_core(m, d, geom2)
    swap = False
    if k2 == "plane" or (k2 == "box" and k1 != "plane"):
        k1, c1, r1, k2, c2, r2 = k2, c2, r2, k1, c1, r1
        swap = True
    if k1 == "plane" and k2 == "box":
        p0, n = c1
        pos, R, s = c2
        corner = pos - R @ (np.where(R.T @ n >= 0.0, 1.0, -1.0) * s)
        dist = float((corner - p0) @ n)
        on2 = corner
        on1 = corner - n * dist
    elif k1 == "box":
        if k2 == "box":
            raise NotImplementedError("box-box")
        dist, on1, on2 = _box_distance(c1, k2, c2, r2)
    elif k1 == "plane":
        if k2 == "plane":
            raise NotImplementedError("plane-plane")
        p0, n = c1
        ends = [c2[0]] if k2 == "pt" else [c2[0] + c2[1], c2[0] - c2[1]]
        hs = [float((e - p0) @ n) for e in ends]
        i = int(np.argmin(hs))
        dist = hs[i] - r2
        on2 = ends[i] - n * r2
        on1 = ends[i] - n * hs[i]
    else:
        a, b = _seg_closest(c1[0], c1[1], c2[0], c2[1])
        v = b - a
        L = float(np.sqrt(v @ v))
        nrm = v / L if L > mjMINVAL else np.array([1.0, 0.0, 0.0])
        dist = L - r1 - r2
        on1 = a + nrm * r1
        on2 = b - nrm * r2
    if swap:
        on1, on2 = on2, on1
    if dist >= distmax:
        if fromto is not None:
            fromto[:] = 0.0
        return float(distmax)
    if fromto is not None:
        fromto[:3] = on1
        fromto[3:] = on2
    return float(dist)


__all__ = [n for n in dir() if n.startswith(("mj", "Mj"))]
