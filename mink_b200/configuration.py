"""Configuration: model + current q, FK and frame queries -- mink's `Configuration` on the B200 engine.

Mirrors reference mink/configuration.py:21-253 (same methods, argument meaning and error types).
`q` may be one configuration ([nq], numpy in / numpy out, like the reference) or a batch ([B, nq],
torch CUDA tensors out).  FK, Jacobians and integration run in libbik (bik_fk, bik_frame_jacobian,
bik_integrate, bik_check_limits); nothing here computes kinematics on the host.
"""

from __future__ import annotations

import logging
from types import SimpleNamespace
from typing import Optional

import numpy as np

from . import exceptions
from .flatten import FlatModel, flatten
from .lie import SE3

SUPPORTED_FRAMES = ("body", "geom", "site")


def as_flat(model) -> FlatModel:
    if isinstance(model, FlatModel):
        return model
    cached = getattr(model, "_bik_flat", None)
    if cached is None:
        cached = flatten(model)
        try:
            model._bik_flat = cached
        except Exception:
            pass
    return cached


def device_model(model, device=None):
    """One DeviceModel per (model, device), created on first use."""
    import torch

    from .engine import DeviceModel

    flat = as_flat(model)
    dev = torch.cuda.current_device() if device is None else int(device)
    cache = flat.__dict__.setdefault("_device_models", {})
    if dev not in cache:
        cache[dev] = DeviceModel(flat, dev)
    return cache[dev]


class Configuration:
    def __init__(self, model, q=None, device: Optional[int] = None):
        self.model = model
        self.flat = as_flat(model)
        self._device = device
        self._dm = None
        self.batched = False
        self._q = None          # torch [B, nq] on the device: fp64 for a single (numpy) configuration and for float64
                                # tensors -- the reference's precision, fp64 kernels -- fp32 for every other batch
        self._q64 = None        # float64 host copy for the single-instance case (exact `.q` round trip)
        self.data = SimpleNamespace(qpos=None)   # minimal stand-in for the reference's `configuration.data`
        self.update(q=q if q is not None else self.flat.qpos0)

    # ---- engine ----------------------------------------------------------------------------------
    @property
    def dm(self):
        if self._dm is None:
            self._dm = device_model(self.model, self._device)
        return self._dm

    def _to_device(self, q):
        import torch

        if isinstance(q, torch.Tensor):
            if q.shape[-1] != self.nq:
                raise ValueError(f"Expected q with trailing dimension {self.nq}, got {tuple(q.shape)}")
            batched = q.ndim == 2
            dtype = torch.float64 if (q.dtype == torch.float64 or not batched) else torch.float32
            t = q.to(device=f"cuda:{self.dm.device}", dtype=dtype).reshape(-1, self.nq).contiguous().clone()
            return t, batched, (None if batched else t[0].cpu().numpy().astype(np.float64))
        a = np.asarray(q, dtype=np.float64)
        if a.shape[-1] != self.nq:
            raise ValueError(f"Expected q with trailing dimension {self.nq}, got {a.shape}")
        batched = a.ndim == 2
        t = torch.tensor(a.reshape(-1, self.nq), dtype=torch.float32 if batched else torch.float64, device=f"cuda:{self.dm.device}")
        return t, batched, (None if batched else a.copy())

    def update(self, q=None) -> None:
        """Set the configuration (reference configuration.py:53-64).  FK is evaluated lazily by the
        kernels that need it; there is no host-side MjData to refresh."""
        if q is not None:
            self._q, self.batched, self._q64 = self._to_device(q)
            self.data.qpos = self._q64 if self._q64 is not None else self._q

    def update_from_keyframe(self, key_name: str) -> None:
        names = self.flat.names["key"]
        if key_name not in names:
            raise exceptions.InvalidKeyframe(key_name, self.flat)
        self.update(q=self.flat.key_qpos[names.index(key_name)])

    # ---- limits ----------------------------------------------------------------------------------
    def check_limits(self, tol: float = 1e-6, safety_break: bool = True) -> None:
        """reference configuration.py:77-110: raise (safety_break) or warn when a limited scalar joint
        is outside [range - tol, range + tol]; free joints are skipped."""
        st = self.dm.check_limits(self._q, tol)
        bad = (st & 1).nonzero().flatten()
        if bad.numel() == 0:
            return
        b = int(bad[0])
        qb = self._q[b].cpu().numpy().astype(np.float64) if self._q64 is None else self._q64
        fm = self.flat
        for d in range(fm.nv):
            a = int(fm.dof_qadr[d])
            if a < 0 or not fm.dof_limited[d]:
                continue
            if qb[a] < fm.dof_lo[d] - tol or qb[a] > fm.dof_hi[d] + tol:
                jnt = int(fm.dof_node[d])
                if safety_break:
                    raise exceptions.NotWithinConfigurationLimits(jnt, qb[a], fm.dof_lo[d], fm.dof_hi[d], fm,
                                                                  instance=b if self.batched else None)
                logging.warning(f"Value {qb[a]:.2f} at index {jnt} is outside of its limits: "
                                f"[{fm.dof_lo[d]:.2f}, {fm.dof_hi[d]:.2f}]"
                                + (f" ({bad.numel()} of {self._q.shape[0]} instances out of limits)" if self.batched else ""))
                return

    # ---- frames ----------------------------------------------------------------------------------
    def _frame(self, frame_name: str, frame_type: str):
        if frame_type not in SUPPORTED_FRAMES:
            raise exceptions.UnsupportedFrame(frame_type, SUPPORTED_FRAMES)
        try:
            return self.flat.frame(frame_name, frame_type)
        except KeyError:
            raise exceptions.InvalidFrame(frame_name, frame_type, self.flat) from None

    def _out(self, t):
        return t if self.batched else t[0].cpu().numpy().astype(np.float64)

    def get_frame_jacobian(self, frame_name: str, frame_type: str):
        """Body-frame Jacobian (6, nv) [or (B, 6, nv)] -- reference configuration.py:112-155."""
        J = self.dm.frame_jacobian(self._q, [self._frame(frame_name, frame_type)])[:, 0]
        return self._out(J)

    def get_transform_frame_to_world(self, frame_name: str, frame_type: str) -> SE3:
        """reference configuration.py:157-185."""
        poses, _ = self.dm.fk(self._q, [self._frame(frame_name, frame_type)])
        p = poses[:, 0].cpu().numpy().astype(np.float64)
        return SE3(p if self.batched else p[0])

    def get_transform(self, source_name: str, source_type: str, dest_name: str, dest_type: str) -> SE3:
        """Pose of `source` in `dest` (reference configuration.py:187-212)."""
        a = self.get_transform_frame_to_world(source_name, source_type)
        b = self.get_transform_frame_to_world(dest_name, dest_type)
        return b.inverse() @ a

    def get_com(self):
        """CoM of subtree(body 1): the reference's `configuration.data.subtree_com[1]` (com_task.py:69)."""
        _, com = self.dm.fk(self._q, [], want_com=True)
        return self._out(com)

    # ---- integration -----------------------------------------------------------------------------
    def _dq(self, velocity, dt):
        import torch

        v = velocity if isinstance(velocity, torch.Tensor) else torch.tensor(np.asarray(velocity, dtype=np.float64))
        v = v.to(device=self._q.device, dtype=self._q.dtype).reshape(-1, self.nv) * float(dt)
        if v.shape[0] != self._q.shape[0]:
            if v.shape[0] != 1:
                raise ValueError(f"velocity batch {v.shape[0]} does not match the configuration batch {self._q.shape[0]}")
            v = v.expand(self._q.shape[0], -1)   # one velocity for the whole batch
        return v.contiguous()

    def integrate(self, velocity, dt: float):
        """q (+) v dt without changing the configuration (reference configuration.py:214-226)."""
        out = self._q.clone()
        self.dm.integrate(out, self._dq(velocity, dt))
        return self._out(out)

    def integrate_inplace(self, velocity, dt: float) -> None:
        """reference configuration.py:228-236."""
        self.dm.integrate(self._q, self._dq(velocity, dt))
        self._sync_from_device()

    def _sync_from_device(self) -> None:
        """Refresh the host mirror of a single configuration after the device copy changed in place."""
        if self._q64 is not None:
            self._q64 = self._q[0].cpu().numpy().astype(np.float64)
            self.data.qpos = self._q64

    # ---- aliases ---------------------------------------------------------------------------------
    @property
    def q(self):
        """Copy of the configuration (reference configuration.py:240-243)."""
        if self.batched:
            return self._q.clone()
        return self._q64.copy()

    @property
    def q_device(self):
        """[B, nq] CUDA tensor used by the kernels (no copy; fp64 for a single configuration, else the batch's dtype)."""
        return self._q

    @property
    def nv(self) -> int:
        return self.flat.nv

    @property
    def nq(self) -> int:
        return self.flat.nq
