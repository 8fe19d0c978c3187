"""ORACLE / TEST INFRASTRUCTURE -- ctypes wrapper around oracle/_build/libikoracle.so (ik_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
It shares the *interface definitions* (struct layouts) with the product via mink_b200._abi, nothing else.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mink_b200._abi import BikFrame, BikLimitDesc, BikTaskDesc, ProblemSpec, c_frames

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libikoracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """fp64 CPU evaluation of a (model blob, ProblemSpec) pair."""

    def __init__(self, blob: bytes, spec: ProblemSpec, nq: int, nv: int):
        self.blob = C.create_string_buffer(blob, len(blob))
        self.spec = spec
        self.nq, self.nv = nq, nv
        self.tasks, self.ntasks, self.limits, self.nlimits, self._keep = spec.to_c()

    def fk(self, q, frames, want_com=True):
        q = _f64(q); B = q.shape[0]
        arr = c_frames(frames)
        poses = np.zeros((B, len(frames), 7)); com = np.zeros((B, 3)) if want_com else None
        rc = lib().iko_fk(self.blob, B, _d(q), arr, len(frames), _d(poses), _d(com))
        assert rc == 0
        return poses, com

    def frame_jacobian(self, q, frames):
        q = _f64(q); B = q.shape[0]
        J = np.zeros((B, len(frames), 6, self.nv))
        assert lib().iko_frame_jacobian(self.blob, B, _d(q), c_frames(frames), len(frames), _d(J)) == 0
        return J

    def fk_jac(self, q, frame_targets=None, posture_target=None, com_targets=None):
        q = _f64(q); B = q.shape[0]; s = self.spec
        ft, ct = _f64(frame_targets), _f64(com_targets)
        pt = _f64(posture_target)
        batched = int(pt is not None and pt.ndim >= 2 and pt.shape[0] == B and pt.size == B * s.nposture * self.nq and B > 1)
        J = np.zeros((B, s.nrows, self.nv)); e = np.zeros((B, s.nrows)); ep = np.zeros((B, max(s.nposture, 1), self.nv))
        assert lib().iko_fk_jac(self.blob, self.tasks, self.ntasks, B, _d(q), _d(ft), _d(pt), batched, _d(ct),
                                _d(J), _d(e), _d(ep)) == 0
        return J, e, ep[:, :s.nposture]

    def objective(self, J, e, ep, damping):
        B = J.shape[0]
        H = np.zeros((B, self.nv, self.nv)); c = np.zeros((B, self.nv))
        ep = _f64(ep) if ep is not None and ep.size else np.zeros((B, 1, self.nv))
        assert lib().iko_objective(self.blob, self.tasks, self.ntasks, B, _d(_f64(J)), _d(_f64(e)), _d(ep),
                                   C.c_double(damping), _d(H), _d(c)) == 0
        return H, c

    def box(self, q, dt):
        q = _f64(q); B = q.shape[0]
        lo = np.zeros((B, self.nv)); hi = np.zeros((B, self.nv))
        assert lib().iko_box(self.blob, self.limits, self.nlimits, B, _d(q), C.c_double(dt), _d(lo), _d(hi)) == 0
        return lo, hi

    def collision(self, q, dt):
        q = _f64(q); B = q.shape[0]; P = self.spec.npairs
        G = np.zeros((B, max(P, 1), self.nv)); h = np.zeros((B, max(P, 1)))
        assert lib().iko_collision(self.blob, self.limits, self.nlimits, B, _d(q), C.c_double(dt), _d(G), _d(h)) == 0
        return G[:, :P], h[:, :P]

    def step(self, q, frame_targets=None, posture_target=None, com_targets=None, dt=1e-2, damping=1e-12,
             nsteps=1, integrate=False, nthreads=0):
        """Returns (dq [B,nv], q_after [B,nq], status [B], n_active [B])."""
        q = _f64(q).copy(); B = q.shape[0]; s = self.spec
        ft, ct, pt = _f64(frame_targets), _f64(com_targets), _f64(posture_target)
        batched = int(pt is not None and pt.ndim >= 2 and pt.size == B * s.nposture * self.nq and B > 1)
        dq = np.zeros((B, self.nv)); st = np.zeros(B, dtype=np.int32); na = np.zeros(B, dtype=np.int32)
        rc = lib().iko_step(self.blob, self.tasks, self.ntasks, self.limits, self.nlimits, B, _d(q), _d(ft), _d(pt),
                            batched, _d(ct), C.c_double(dt), C.c_double(damping), int(nsteps), int(bool(integrate)),
                            _d(dq), st.ctypes.data_as(C.POINTER(C.c_int32)), na.ctypes.data_as(C.POINTER(C.c_int32)),
                            int(nthreads))
        assert rc == 0
        return dq, q, st, na

    def converge(self, q, frame_targets=None, posture_target=None, com_targets=None, dt=1e-2, damping=1e-12, max_iters=20,
                 pos_threshold=1e-4, ori_threshold=1e-4, nthreads=0):
        """The examples' solve+integrate-until-threshold loop.  Returns (q_final [B,nq], iters [B], converged [B], status [B])."""
        q = _f64(q).copy(); B = q.shape[0]; s = self.spec
        ft, ct, pt = _f64(frame_targets), _f64(com_targets), _f64(posture_target)
        batched = int(pt is not None and pt.ndim >= 2 and pt.size == B * s.nposture * self.nq and B > 1)
        it = np.zeros(B, dtype=np.int32); cv = np.zeros(B, dtype=np.int32); st = np.zeros(B, dtype=np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = lib().iko_converge(self.blob, self.tasks, self.ntasks, self.limits, self.nlimits, B, _d(q), _d(ft), _d(pt), batched, _d(ct),
                                C.c_double(dt), C.c_double(damping), int(max_iters), C.c_double(pos_threshold), C.c_double(ori_threshold),
                                i32(it), i32(cv), i32(st), int(nthreads))
        assert rc == 0
        return q, it, cv, st

    def integrate(self, q, dq):
        q = _f64(q).copy()
        assert lib().iko_integrate(self.blob, q.shape[0], _d(q), _d(_f64(dq))) == 0
        return q


def num_threads() -> int:
    return int(lib().iko_num_threads())
