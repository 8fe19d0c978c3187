"""TEST INFRASTRUCTURE -- ctypes wrapper for tests/host_emu (kernel headers compiled for the host, 1 lane)."""

import ctypes as C
import os
import subprocess

import numpy as np

from mink_b200._abi import c_frames

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")
_SO = os.path.join(_HERE, "_build", "libbikemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "emu.cpp")
        csrc = os.path.join(os.path.dirname(_HERE), "..", "mink_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", _SO, src])
        _lib = C.CDLL(_SO)
        _lib.emu_problem_create.restype = C.c_void_p
        _lib.emu_last_error.restype = C.c_char_p
    return _lib


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Emu:
    def __init__(self, blob: bytes, spec, nq, nv):
        self.spec, self.nq, self.nv = spec, nq, nv
        t, nt, l, nl, self._keep = spec.to_c()
        self.h = lib().emu_problem_create(blob, len(blob), t, nt, l, nl)
        if not self.h:
            raise RuntimeError(lib().emu_last_error().decode())
        self.h = C.c_void_p(self.h)

    def header(self):
        """nv, nu (coupled dofs), nfree (leading coupled dofs without bounds), nnode, nneeded, nslots, G, nsteps of the image."""
        out = np.zeros(8, np.int32)
        lib().emu_header(self.h, _p(out, C.c_int32))
        return dict(zip(("nv", "nu", "nfree", "nnode", "nneeded", "nslots", "G", "nsteps"), (int(v) for v in out)))

    def set_knobs(self, sweeps: int, rule: int):
        lib().emu_set_knobs(self.h, int(sweeps), int(rule))

    def fk_jac(self, q, ftgt=None, ptgt=None, ctgt=None, dt=1e-2):
        q = _f32(q); B = q.shape[0]; s = self.spec
        ftgt, ptgt, ctgt = _f32(ftgt), _f32(ptgt), _f32(ctgt)
        J = np.zeros((B, s.nrows, self.nv), np.float32); e = np.zeros((B, max(s.nrows, 1)), np.float32)
        ep = np.zeros((B, max(s.nposture, 1), self.nv), np.float32)
        Gc = np.zeros((B, max(s.npairs, 1), self.nv), np.float32); hc = np.zeros((B, max(s.npairs, 1)), np.float32)
        batched = int(ptgt is not None and ptgt.size == B * s.nposture * self.nq and B > 1)
        lib().emu_fk_jac(self.h, B, _p(q), _p(ftgt), _p(ptgt), batched, _p(ctgt), C.c_float(dt), _p(J), _p(e), _p(ep), _p(Gc), _p(hc))
        return J, e[:, :s.nrows], ep[:, :s.nposture], Gc[:, :s.npairs], hc[:, :s.npairs]

    def solve(self, q, J, e, ep, Gc, hc, dt, damping, use_double=True, want_objective=False):
        q = _f32(q); B = q.shape[0]; s = self.spec
        pad = lambda a, n: np.ascontiguousarray(a if a.shape[1] else np.zeros((B, 1) + a.shape[2:], np.float32))
        dq = np.zeros((B, self.nv), np.float32); st = np.zeros(B, np.int32); it = np.zeros(B, np.int32)
        H = np.zeros((B, self.nv, self.nv)) if want_objective else None
        c = np.zeros((B, self.nv)) if want_objective else None
        lo = np.zeros((B, self.nv), np.float32); hi = np.zeros((B, self.nv), np.float32)
        self.last_rc = lib().emu_solve(self.h, B, _p(q), _p(_f32(J)), _p(_f32(pad(e, 1))), _p(_f32(pad(ep, 1))), _p(_f32(pad(Gc, 1))), _p(_f32(pad(hc, 1))),
                        C.c_float(dt), C.c_double(damping), int(use_double), _p(dq), _p(st, C.c_int32), _p(it, C.c_int32),
                        _p(H, C.c_double), _p(c, C.c_double), _p(lo), _p(hi))
        return dq, st, it, H, c, lo, hi

    def solve_warm(self, q, J, e, ep, dt, damping, dq_prev, warm):
        """One step of a rollout on the small-group fp64 path: `warm` ([B, nu] int8, zeros before the first step) and
        `dq_prev` (the previous step's dq) are updated in place, as bik_step does between its steps."""
        q = _f32(q); B = q.shape[0]
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32)
        assert dq_prev.dtype == np.float32 and warm.dtype == np.int8 and dq_prev.flags.c_contiguous and warm.flags.c_contiguous
        rc = lib().emu_solve_warm(self.h, B, _p(q), _p(_f32(J)), _p(_f32(e)), _p(_f32(ep)), C.c_float(dt), C.c_double(damping),
                                  _p(dq_prev), _p(st, C.c_int32), _p(it, C.c_int32), warm.ctypes.data_as(C.POINTER(C.c_byte)))
        assert rc == 0, lib().emu_last_error()
        return dq_prev, st, it

    def fk(self, q, frames, want_J=False):
        q = _f32(q); B = q.shape[0]
        poses = np.zeros((B, len(frames), 7), np.float32); com = np.zeros((B, 3), np.float32)
        J = np.zeros((B, len(frames), 6, self.nv), np.float32) if want_J else None
        assert lib().emu_fk(self.h, B, _p(q), c_frames(frames), len(frames), _p(poses), _p(com), _p(J)) == 0
        return poses, com, J

    def integrate(self, q, dq):
        q = _f32(q).copy()
        lib().emu_integrate(self.h, q.shape[0], _p(q), _p(_f32(dq)))
        return q

    def check_limits(self, q, tol=1e-6):
        q = _f32(q); st = np.zeros(q.shape[0], np.int32)
        lib().emu_check_limits(self.h, q.shape[0], _p(q), C.c_float(tol), _p(st, C.c_int32))
        return st
