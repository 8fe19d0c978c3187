"""SO(3) / SE(3) value types with mink's method names (host side, numpy, optionally batched).

These exist for API parity with `mink.lie` (reference mink/lie/{base,so3,se3}.py): examples build
targets with them (`SE3.from_mocap_id`, `SE3.from_rotation_and_translation`, `T @ SE3.from_translation(..)`).
The per-step Lie algebra of the hot path does NOT run here -- it lives in csrc/bik_math.h.  The
implementation is array-first: every operation works on a leading batch shape, a single transform
being the batch shape ().

Parameterisation (same as the reference): SO3.wxyz = (qw,qx,qy,qz); SE3.wxyz_xyz = (qw,qx,qy,qz,x,y,z);
tangents are (omega) and (v, omega).
"""

from __future__ import annotations

from typing import Union

import numpy as np

_EPS = 1e-10


def _skew(w: np.ndarray) -> np.ndarray:
    z = np.zeros_like(w[..., 0])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1),
                     np.stack([w[..., 2], z, -w[..., 0]], -1),
                     np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def _qmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def _qmat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = np.moveaxis(q, -1, 0)
    return np.stack([np.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1)], -2)


def _mat_to_quat(R: np.ndarray) -> np.ndarray:
    """Largest-pivot conversion, normalised (same branch rule as mju_mat2Quat)."""
    R = np.asarray(R, dtype=np.float64)
    flat = R.reshape(-1, 3, 3)
    out = np.empty((flat.shape[0], 4))
    for n, m in enumerate(flat):
        tr = m[0, 0] + m[1, 1] + m[2, 2]
        if tr > 0:
            w = 0.5 * np.sqrt(1 + tr); k = 0.25 / w
            q = [w, k * (m[2, 1] - m[1, 2]), k * (m[0, 2] - m[2, 0]), k * (m[1, 0] - m[0, 1])]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            x = 0.5 * np.sqrt(1 + m[0, 0] - m[1, 1] - m[2, 2]); k = 0.25 / x
            q = [k * (m[2, 1] - m[1, 2]), x, k * (m[1, 0] + m[0, 1]), k * (m[0, 2] + m[2, 0])]
        elif m[1, 1] > m[2, 2]:
            y = 0.5 * np.sqrt(1 - m[0, 0] + m[1, 1] - m[2, 2]); k = 0.25 / y
            q = [k * (m[0, 2] - m[2, 0]), k * (m[1, 0] + m[0, 1]), y, k * (m[2, 1] + m[1, 2])]
        else:
            z = 0.5 * np.sqrt(1 - m[0, 0] - m[1, 1] + m[2, 2]); k = 0.25 / z
            q = [k * (m[1, 0] - m[0, 1]), k * (m[0, 2] + m[2, 0]), k * (m[2, 1] + m[1, 2]), z]
        q = np.asarray(q)
        out[n] = q / np.linalg.norm(q)
    return out.reshape(R.shape[:-2] + (4,))


def _coef_A(t2):
    """(1 - (t/2) cot(t/2)) / t^2, series near zero."""
    t2 = np.asarray(t2, dtype=np.float64)
    small = t2 < 1e-4
    ts = np.where(small, 1.0, t2)
    th = np.sqrt(ts)
    big = (1.0 - 0.5 * th * np.cos(0.5 * th) / np.sin(0.5 * th)) / ts
    return np.where(small, 1.0 / 12.0 + t2 / 720.0 + t2 * t2 / 30240.0, big)


class MatrixLieGroup:
    """Common interface (reference mink/lie/base.py:8-156)."""

    matrix_dim: int
    parameters_dim: int
    tangent_dim: int
    space_dim: int

    def __matmul__(self, other):
        if isinstance(other, np.ndarray):
            return self.apply(other)
        return self.multiply(other)

    # right / left plus and minus (base.py:108-130)
    def rplus(self, tangent):
        return self @ type(self).exp(tangent)

    def rminus(self, other):
        return (other.inverse() @ self).log()

    def lplus(self, tangent):
        return type(self).exp(tangent) @ self

    def lminus(self, other):
        return (self @ other.inverse()).log()

    plus = rplus
    minus = rminus

    @classmethod
    def rjac(cls, tangent):
        return cls.ljac(-np.asarray(tangent))

    @classmethod
    def rjacinv(cls, tangent):
        return cls.ljacinv(-np.asarray(tangent))

    def jlog(self):
        return self.rjacinv(self.log())


class SO3(MatrixLieGroup):
    matrix_dim, parameters_dim, tangent_dim, space_dim = 3, 4, 3, 3

    def __init__(self, wxyz: np.ndarray):
        wxyz = np.asarray(wxyz, dtype=np.float64)
        if wxyz.shape[-1:] != (4,):
            raise ValueError(f"Expeced wxyz to be a length 4 vector but got {wxyz.shape[-1] if wxyz.ndim else 0}.")
        self.wxyz = wxyz

    def __repr__(self):
        return f"SO3(wxyz={np.round(self.wxyz, 5)})"

    def parameters(self):
        return self.wxyz

    def copy(self):
        return SO3(self.wxyz.copy())

    @classmethod
    def identity(cls):
        return SO3(np.array([1.0, 0.0, 0.0, 0.0]))

    @classmethod
    def from_matrix(cls, matrix):
        matrix = np.asarray(matrix, dtype=np.float64)
        assert matrix.shape[-2:] == (3, 3)
        return SO3(_mat_to_quat(matrix))

    @classmethod
    def from_x_radians(cls, theta):
        return SO3.exp(np.array([theta, 0.0, 0.0]))

    @classmethod
    def from_y_radians(cls, theta):
        return SO3.exp(np.array([0.0, theta, 0.0]))

    @classmethod
    def from_z_radians(cls, theta):
        return SO3.exp(np.array([0.0, 0.0, theta]))

    @classmethod
    def from_rpy_radians(cls, roll, pitch, yaw):
        return SO3.from_z_radians(yaw) @ SO3.from_y_radians(pitch) @ SO3.from_x_radians(roll)

    @classmethod
    def sample_uniform(cls):
        u1, u2, u3 = np.random.uniform(0.0, [1.0, 2 * np.pi, 2 * np.pi])
        a, b = np.sqrt(1 - u1), np.sqrt(u1)
        return SO3(np.array([a * np.sin(u2), a * np.cos(u2), b * np.sin(u3), b * np.cos(u3)]))

    def as_matrix(self):
        return _qmat(self.wxyz)

    def compute_roll_radians(self):
        q0, q1, q2, q3 = np.moveaxis(self.wxyz, -1, 0)
        return np.arctan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 ** 2 + q2 ** 2))

    def compute_pitch_radians(self):
        q0, q1, q2, q3 = np.moveaxis(self.wxyz, -1, 0)
        return np.arcsin(2 * (q0 * q2 - q3 * q1))

    def compute_yaw_radians(self):
        q0, q1, q2, q3 = np.moveaxis(self.wxyz, -1, 0)
        return np.arctan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 ** 2 + q3 ** 2))

    def as_rpy_radians(self):
        from collections import namedtuple
        return namedtuple("RollPitchYaw", "roll pitch yaw")(self.compute_roll_radians(), self.compute_pitch_radians(),
                                                             self.compute_yaw_radians())

    def inverse(self):
        return SO3(self.wxyz * np.array([1.0, -1.0, -1.0, -1.0]))

    def normalize(self):
        return SO3(self.wxyz / np.linalg.norm(self.wxyz, axis=-1, keepdims=True))

    def apply(self, target):
        target = np.asarray(target, dtype=np.float64)
        assert target.shape[-1] == 3
        pad = np.concatenate([np.zeros(target.shape[:-1] + (1,)), target], -1)
        return _qmul(_qmul(self.wxyz, pad), self.inverse().wxyz)[..., 1:]

    def multiply(self, other):
        return SO3(_qmul(self.wxyz, other.wxyz))

    @classmethod
    def exp(cls, tangent):
        tangent = np.asarray(tangent, dtype=np.float64)
        assert tangent.shape[-1] == 3
        t2 = np.sum(tangent * tangent, -1, keepdims=True)
        small = t2 < _EPS
        th = np.sqrt(np.where(small, 1.0, t2))
        real = np.where(small, 1.0 - t2 / 8.0 + t2 * t2 / 384.0, np.cos(0.5 * th))
        imag = np.where(small, 0.5 - t2 / 48.0 + t2 * t2 / 3840.0, np.sin(0.5 * th) / th)
        return SO3(np.concatenate([real, imag * tangent], -1))

    def log(self):
        w = self.wxyz[..., :1]
        v = self.wxyz[..., 1:]
        n2 = np.sum(v * v, -1, keepdims=True)
        small = n2 < _EPS
        n = np.sqrt(np.where(small, 1.0, n2))
        wsafe = np.where(small, w, 1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            taylor = 2.0 / wsafe - 2.0 / 3.0 * n2 / wsafe ** 3
            at = np.arctan2(np.where(w < 0, -n, n), np.abs(w))
            full = np.where(np.abs(w) < _EPS, np.where(w > 0, 1.0, -1.0) * np.pi / n, 2.0 * at / n)
        return np.where(small, taylor, full) * v

    def adjoint(self):
        return self.as_matrix()

    @classmethod
    def ljac(cls, other):
        other = np.asarray(other, dtype=np.float64)
        t2 = np.sum(other * other, -1)[..., None, None]
        small = t2 < 1e-8
        ts = np.where(small, 1.0, t2); th = np.sqrt(ts)
        A = np.where(small, 0.5 - t2 / 24.0, (1 - np.cos(th)) / ts)
        B = np.where(small, 1.0 / 6.0 - t2 / 120.0, (th - np.sin(th)) / (ts * th))
        S = _skew(other)
        return np.eye(3) + A * S + B * (S @ S)

    @classmethod
    def ljacinv(cls, other):
        other = np.asarray(other, dtype=np.float64)
        A = _coef_A(np.sum(other * other, -1))[..., None, None]
        S = _skew(other)
        return np.eye(3) - 0.5 * S + A * (S @ S)


class SE3(MatrixLieGroup):
    matrix_dim, parameters_dim, tangent_dim, space_dim = 4, 7, 6, 3

    def __init__(self, wxyz_xyz: np.ndarray):
        self.wxyz_xyz = np.asarray(wxyz_xyz, dtype=np.float64)
        if self.wxyz_xyz.shape[-1:] != (7,):
            raise ValueError("Expected wxyz_xyz to have 7 trailing entries")

    def __repr__(self):
        return f"SE3(wxyz={np.round(self.wxyz_xyz[..., :4], 5)}, xyz={np.round(self.wxyz_xyz[..., 4:], 5)})"

    def copy(self):
        return SE3(np.array(self.wxyz_xyz))

    def parameters(self):
        return self.wxyz_xyz

    @classmethod
    def identity(cls):
        return SE3(np.array([1.0, 0, 0, 0, 0, 0, 0]))

    @classmethod
    def from_rotation_and_translation(cls, rotation: SO3, translation: np.ndarray):
        translation = np.asarray(translation, dtype=np.float64)
        assert translation.shape[-1] == 3
        shape = np.broadcast_shapes(rotation.wxyz.shape[:-1], translation.shape[:-1])
        return SE3(np.concatenate([np.broadcast_to(rotation.wxyz, shape + (4,)), np.broadcast_to(translation, shape + (3,))], -1))

    @classmethod
    def from_rotation(cls, rotation: SO3):
        return SE3.from_rotation_and_translation(rotation, np.zeros(3))

    @classmethod
    def from_translation(cls, translation):
        return SE3.from_rotation_and_translation(SO3.identity(), translation)

    @classmethod
    def from_matrix(cls, matrix):
        matrix = np.asarray(matrix, dtype=np.float64)
        assert matrix.shape[-2:] == (4, 4)
        return SE3.from_rotation_and_translation(SO3.from_matrix(matrix[..., :3, :3]), matrix[..., :3, 3])

    @classmethod
    def from_mocap_id(cls, data, mocap_id: int):
        return SE3.from_rotation_and_translation(SO3(np.array(data.mocap_quat[mocap_id])), np.array(data.mocap_pos[mocap_id]))

    @classmethod
    def from_mocap_name(cls, model, data, mocap_name: str):
        from .exceptions import InvalidMocapBody
        mocap_id = model.body(mocap_name).mocapid[0]
        if mocap_id == -1:
            raise InvalidMocapBody(mocap_name, model)
        return SE3.from_mocap_id(data, mocap_id)

    @classmethod
    def sample_uniform(cls):
        return SE3.from_rotation_and_translation(SO3.sample_uniform(), np.random.uniform(-1.0, 1.0, size=3))

    def rotation(self) -> SO3:
        return SO3(self.wxyz_xyz[..., :4])

    def translation(self) -> np.ndarray:
        return self.wxyz_xyz[..., 4:]

    def as_matrix(self):
        out = np.zeros(self.wxyz_xyz.shape[:-1] + (4, 4))
        out[..., :3, :3] = self.rotation().as_matrix()
        out[..., :3, 3] = self.translation()
        out[..., 3, 3] = 1.0
        return out

    def inverse(self):
        Ri = self.rotation().inverse()
        return SE3.from_rotation_and_translation(Ri, -Ri.apply(self.translation()))

    def normalize(self):
        return SE3.from_rotation_and_translation(self.rotation().normalize(), self.translation())

    def apply(self, target):
        return self.rotation().apply(np.asarray(target, dtype=np.float64)) + self.translation()

    def multiply(self, other):
        return SE3.from_rotation_and_translation(self.rotation() @ other.rotation(),
                                                 self.rotation().apply(other.translation()) + self.translation())

    @classmethod
    def exp(cls, tangent):
        tangent = np.asarray(tangent, dtype=np.float64)
        assert tangent.shape[-1] == 6
        w = tangent[..., 3:]
        R = SO3.exp(w)
        V = SO3.ljac(w)   # V(w) is the SO(3) left Jacobian
        return SE3.from_rotation_and_translation(R, np.einsum("...ij,...j->...i", V, tangent[..., :3]))

    def log(self):
        w = self.rotation().log()
        Vi = SO3.ljacinv(w)
        return np.concatenate([np.einsum("...ij,...j->...i", Vi, self.translation()), w], -1)

    def adjoint(self):
        R = self.rotation().as_matrix()
        out = np.zeros(R.shape[:-2] + (6, 6))
        out[..., :3, :3] = R
        out[..., 3:, 3:] = R
        out[..., :3, 3:] = _skew(self.translation()) @ R
        return out

    @staticmethod
    def _Q(c):
        v, w = c[..., :3], c[..., 3:]
        t2 = np.sum(w * w, -1)[..., None, None]
        small = t2 < 1e-6
        ts = np.where(small, 1.0, t2); th = np.sqrt(ts); s, co = np.sin(th), np.cos(th)
        B = np.where(small, 1.0 / 6.0 - t2 / 120.0, (th - s) / (ts * th))
        C = np.where(small, -1.0 / 24.0 + t2 / 720.0, (1.0 - ts / 2.0 - co) / (ts * ts))
        D = np.where(small, 1.0 / 120.0 - t2 / 2520.0, (2 * th - 3 * s + th * co) / (2 * ts * ts * th))
        V, W = _skew(v), _skew(w)
        VW = V @ W; WV = np.swapaxes(VW, -1, -2); WVW = WV @ W; VWW = VW @ W
        return 0.5 * V + B * (WV + VW + WVW) - C * (VWW - np.swapaxes(VWW, -1, -2) - 3 * WVW) + D * (WVW @ W + W @ WVW)

    @classmethod
    def ljac(cls, other):
        other = np.asarray(other, dtype=np.float64)
        J = SO3.ljac(other[..., 3:])
        out = np.zeros(other.shape[:-1] + (6, 6))
        ident = (np.sum(other[..., 3:] ** 2, -1) < _EPS)[..., None, None]   # reference quirk: se3.py:202-203
        out[..., :3, :3] = J; out[..., 3:, 3:] = J; out[..., :3, 3:] = cls._Q(other)
        return np.where(ident, np.eye(6), out)

    @classmethod
    def ljacinv(cls, other):
        other = np.asarray(other, dtype=np.float64)
        Ji = SO3.ljacinv(other[..., 3:])
        out = np.zeros(other.shape[:-1] + (6, 6))
        ident = (np.sum(other[..., 3:] ** 2, -1) < _EPS)[..., None, None]   # reference quirk: se3.py:212-214
        out[..., :3, :3] = Ji; out[..., 3:, 3:] = Ji; out[..., :3, 3:] = -Ji @ cls._Q(other) @ Ji
        return np.where(ident, np.eye(6), out)


Transform = Union[SO3, SE3]
