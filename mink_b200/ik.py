"""build_ik / solve_ik with the reference signatures (mink/solve_ik.py:43-105) on the B200 engine.

`solve_ik` lowers the task and limit lists to ONE bik_problem (cached by layout) and issues one
bik_step: check_limits -> FK + Jacobians (K1) -> QP assembly + exact active-set solve (K2).
The `solver` string is accepted for drop-in compatibility and ignored: the device solver returns
the exact optimum of the same strictly convex QP that quadprog / daqp solve.
"""

from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np

from ._abi import ProblemSpec
from .configuration import Configuration
from .limits import ConfigurationLimit, Limit
from .tasks import ComTask, FrameTask, PostureTask, Task, problem_for


class Problem(NamedTuple):
    """Stand-in for qpsolvers.Problem: min 1/2 x^T P x + q^T x  s.t.  G x <= h."""

    P: object
    q: object
    G: Optional[object]
    h: Optional[object]


def _lower(configuration: Configuration, tasks: Sequence[Task], limits: Optional[Sequence[Limit]]):
    flat = configuration.flat
    if limits is None:   # default of the reference (solve_ik.py:28-29)
        limits = [ConfigurationLimit(configuration.model)]
    # libbik stacks rows by task kind order of appearance; keep frame / com targets in list order
    tspecs = [t._spec(flat) for t in tasks]
    lspecs = [s for s in (l._spec(flat) for l in limits) if s is not None]
    spec = ProblemSpec(tspecs, lspecs)
    B = configuration.q_device.shape[0]
    ft = [np.broadcast_to(np.asarray(t._target()).reshape(-1, 7), (B, 7)) for t in tasks if isinstance(t, FrameTask)]
    pt = [np.asarray(t._target()).reshape(-1, flat.nq) for t in tasks if isinstance(t, PostureTask)]
    ct = [np.broadcast_to(np.asarray(t._target()).reshape(-1, 3), (B, 3)) for t in tasks if isinstance(t, ComTask)]
    kw = {}
    if ft:
        kw["frame_targets"] = np.stack(ft, axis=1)
    if pt:
        nb = max(p.shape[0] for p in pt)
        kw["posture_targets"] = np.stack([np.broadcast_to(p, (nb, flat.nq)) for p in pt], axis=1)
    if ct:
        kw["com_targets"] = np.stack(ct, axis=1)
    return problem_for(configuration, spec), kw, limits


def build_ik(configuration: Configuration, tasks: Sequence[Task], dt: float, damping: float = 1e-12,
             limits: Optional[Sequence[Limit]] = None) -> Problem:
    """QP of the IK problem (reference solve_ik.py:43-65): P = damping I + sum H_t, q = sum c_t,
    G / h stacked over the limits in list order (None when there is no inequality)."""
    prob, kw, limits = _lower(configuration, tasks, limits)
    J, e, ep, _, _ = prob.fk_jac(configuration.q_device, dt=dt, **kw)
    H, c = prob.objective(J, e, ep, damping)
    G_list, h_list = [], []
    for limit in limits:
        con = limit.compute_qp_inequalities(configuration, dt)
        if not con.inactive:
            G_list.append(con.G)
            h_list.append(con.h)
    if configuration.batched:
        import torch

        G = torch.cat([g.to(h_list[0].dtype) for g in G_list], dim=1) if G_list else None
        h = torch.cat(list(h_list), dim=1) if h_list else None
        return Problem(H, c, G, h)
    G = np.vstack(G_list) if G_list else None
    h = np.hstack(h_list) if h_list else None
    return Problem(H[0].cpu().numpy(), c[0].cpu().numpy(), G, h)


def solve_ik(configuration: Configuration, tasks: Sequence[Task], dt: float, solver: str = "b200", damping: float = 1e-12,
             safety_break: bool = False, limits: Optional[Sequence[Limit]] = None, **kwargs):
    """Velocity v = dq / dt tangent to the configuration (reference solve_ik.py:68-105).

    Returns a numpy (nv,) vector for a single configuration, a torch CUDA [B, nv] tensor for a batch.
    Raises NotWithinConfigurationLimits when `safety_break` and some instance is out of limits;
    raises AssertionError (like the reference's `assert dq is not None`) when the QP of any instance
    could not be solved.
    """
    prob, kw, limits = _lower(configuration, tasks, limits)
    q = configuration.q_device
    dq, status = prob.step(q, dt=dt, damping=damping, nsteps=1, integrate=False, **kw)
    st = status.cpu().numpy()
    if (st & 1).any():
        if safety_break:
            configuration.check_limits(safety_break=True)
        else:
            configuration.check_limits(safety_break=False)
    assert not (st & (2 | 4 | 8)).any(), f"QP not solved for {int(((st & 14) != 0).sum())} instance(s) (status bits {np.unique(st & 14)})"
    v = dq / float(dt)
    if configuration.batched:
        return v
    return v[0].cpu().numpy().astype(np.float64)


def converge_ik(configuration: Configuration, tasks: Sequence[Task], dt: float, solver: str = "b200", damping: float = 1e-12,
                limits: Optional[Sequence[Limit]] = None, max_iters: int = 20, pos_threshold: float = 1e-4,
                ori_threshold: float = 1e-4, check_every: int = 1):
    """The inner loop the reference's examples wrap around solve_ik (examples/arm_iiwa.py:63-70,
    examples/quadruped_spot.py:89-104), run on the device per instance:

        for i in range(max_iters):
            vel = solve_ik(configuration, tasks, dt, solver, damping, limits=limits)
            configuration.integrate_inplace(vel, dt)
            if every FrameTask/RelativeFrameTask: |err[:3]| <= pos_threshold and |err[3:]| <= ori_threshold: break

    The configuration is updated in place; returns (iters, converged): steps taken and whether the thresholds were
    met, numpy scalars for a single configuration, [B] numpy arrays for a batch."""
    prob, kw, limits = _lower(configuration, tasks, limits)
    q = configuration.q_device
    iters, status = prob.converge(q, dt=dt, damping=damping, max_iters=max_iters, pos_threshold=pos_threshold,
                                  ori_threshold=ori_threshold, check_every=check_every, **kw)
    configuration._sync_from_device()
    st = status.cpu().numpy()
    assert not (st & (2 | 4 | 8)).any(), f"QP not solved for {int(((st & 14) != 0).sum())} instance(s) (status bits {np.unique(st & 14)})"
    it, ok = iters.cpu().numpy(), (st & 16) == 0
    if configuration.batched:
        return it, ok
    return int(it[0]), bool(ok[0])
