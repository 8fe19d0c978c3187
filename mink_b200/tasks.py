"""Kinematic tasks with mink's interface (reference mink/tasks/*.py) evaluated by libbik.

`compute_error` / `compute_jacobian` / `compute_qp_objective` keep the reference signatures and
return numpy arrays for a single configuration, torch CUDA tensors with a leading batch dimension
otherwise.  Every evaluation is a K1 (+K2 assembly) kernel launch on a one-task problem; `solve_ik`
does not go through these methods -- it lowers the whole task list to one problem.
"""

from __future__ import annotations

import abc
from typing import NamedTuple, Optional

import numpy as np

from ._abi import TASK_COM, TASK_FRAME, TASK_POSTURE, TASK_RELATIVE_FRAME, ProblemSpec, TaskSpec
from .configuration import SUPPORTED_FRAMES, Configuration, as_flat
from .exceptions import (InvalidDamping, InvalidFrame, InvalidGain, InvalidTarget, TargetNotSet, TaskDefinitionError,
                         UnsupportedFrame)
from .lie import SE3


class Objective(NamedTuple):
    """Quadratic objective 1/2 x^T H x + c^T x (reference tasks/task.py:12-22)."""

    H: object
    c: object

    def value(self, x):
        return x.T @ self.H @ x + self.c @ x


def problem_for(configuration: Configuration, spec: ProblemSpec):
    """bik_problem handles cached on the flat model by static layout."""
    from .engine import Problem

    dm = configuration.dm
    cache = dm.__dict__.setdefault("_problems", {})
    key = spec.key()
    if key not in cache:
        if len(cache) > 64:
            cache.clear()
        cache[key] = Problem(dm, spec)
    return cache[key]


class Task(abc.ABC):
    """Base class: cost vector, gain in [0,1], Levenberg-Marquardt damping >= 0 (tasks/task.py:54-79)."""

    def __init__(self, cost: np.ndarray, gain: float = 1.0, lm_damping: float = 0.0):
        if not 0.0 <= gain <= 1.0:
            raise InvalidGain("`gain` must be in the range [0, 1]")
        if lm_damping < 0.0:
            raise InvalidDamping("`lm_damping` must be >= 0")
        self.cost = cost
        self.gain = gain
        self.lm_damping = lm_damping

    # -- lowering to the C ABI ---------------------------------------------------------------------
    @abc.abstractmethod
    def _spec(self, flat) -> TaskSpec:
        ...

    @abc.abstractmethod
    def _target_kwargs(self, configuration: Configuration) -> dict:
        """Keyword arguments (frame_targets / posture_targets / com_targets) for Problem.fk_jac."""

    def _evaluate(self, configuration: Configuration):
        prob = problem_for(configuration, ProblemSpec([self._spec(configuration.flat)], []))
        J, e, ep, _, _ = prob.fk_jac(configuration.q_device, dt=1.0, **self._target_kwargs(configuration))
        return prob, J, e, ep

    @abc.abstractmethod
    def compute_error(self, configuration: Configuration):
        ...

    @abc.abstractmethod
    def compute_jacobian(self, configuration: Configuration):
        ...

    def compute_qp_objective(self, configuration: Configuration) -> Objective:
        """(H, c) of this task alone (reference tasks/task.py:105-138), fp64."""
        prob, J, e, ep = self._evaluate(configuration)
        H, c = prob.objective(J, e, ep, 0.0)
        if configuration.batched:
            return Objective(H, c)
        return Objective(H[0].cpu().numpy(), c[0].cpu().numpy())


def _out(configuration, t):
    return t if configuration.batched else t[0].cpu().numpy().astype(np.float64)


class FrameTask(Task):
    """Regulate the pose of a body / geom / site frame in the world (reference tasks/frame_task.py)."""

    k: int = 6

    def __init__(self, frame_name: str, frame_type: str, position_cost, orientation_cost, gain: float = 1.0,
                 lm_damping: float = 0.0):
        super().__init__(cost=np.zeros((self.k,)), gain=gain, lm_damping=lm_damping)
        self.frame_name = frame_name
        self.frame_type = frame_type
        self.position_cost = position_cost
        self.orientation_cost = orientation_cost
        self.transform_target_to_world: Optional[SE3] = None
        self.set_position_cost(position_cost)
        self.set_orientation_cost(orientation_cost)

    def _set(self, cost, what: str, sl: slice):
        cost = np.atleast_1d(cost)
        if cost.ndim != 1 or cost.shape[0] not in (1, 3):
            raise TaskDefinitionError(f"{self.__class__.__name__} {what} cost should be a vector of shape 1 "
                                      f"(aka identical cost for all coordinates) or (3,) but got {cost.shape}")
        if not np.all(cost >= 0.0):
            raise TaskDefinitionError(f"{self.__class__.__name__} position cost should be >= 0")
        self.cost[sl] = cost

    def set_position_cost(self, position_cost) -> None:
        self._set(position_cost, "position", slice(0, 3))

    def set_orientation_cost(self, orientation_cost) -> None:
        self._set(orientation_cost, "orientation", slice(3, 6))

    def set_target(self, transform_target_to_world: SE3) -> None:
        self.transform_target_to_world = transform_target_to_world.copy()   # copied, as frame_task.py:83

    def set_target_from_configuration(self, configuration: Configuration) -> None:
        self.set_target(configuration.get_transform_frame_to_world(self.frame_name, self.frame_type))

    def _spec(self, flat) -> TaskSpec:
        if self.frame_type not in SUPPORTED_FRAMES:
            raise UnsupportedFrame(self.frame_type, SUPPORTED_FRAMES)
        try:
            frame = flat.frame(self.frame_name, self.frame_type)
        except KeyError:
            raise InvalidFrame(self.frame_name, self.frame_type, flat) from None
        return TaskSpec(TASK_FRAME, frame=frame, cost=self.cost.copy(), gain=self.gain, lm_damping=self.lm_damping)

    def _target(self):
        if self.transform_target_to_world is None:
            raise TargetNotSet(self.__class__.__name__)
        return self.transform_target_to_world.wxyz_xyz

    def _target_kwargs(self, configuration):
        t = np.asarray(self._target())
        B = configuration.q_device.shape[0]
        return dict(frame_targets=np.broadcast_to(t.reshape(-1, 1, 7), (B, 1, 7)))

    def compute_error(self, configuration: Configuration):
        """e = log(T_wb^-1 T_wt), a body twist (frame_task.py:95-122)."""
        _, J, e, _ = self._evaluate(configuration)
        return _out(configuration, e)

    def compute_jacobian(self, configuration: Configuration):
        """J = -jlog(T_wt^-1 T_wb) J_b (frame_task.py:124-146)."""
        _, J, e, _ = self._evaluate(configuration)
        return _out(configuration, J)


class RelativeFrameTask(FrameTask):
    """Regulate the pose of a frame relative to another (root) frame -- reference
    mink/tasks/relative_frame_task.py.  The target is T_root<-target; e = log(T_rt^-1 T_rf),
    J = jlog(T_tf)(J_f - Ad(T_rf^-1) J_r)."""

    def __init__(self, frame_name: str, frame_type: str, root_name: str, root_type: str, position_cost, orientation_cost,
                 gain: float = 1.0, lm_damping: float = 0.0):
        super().__init__(frame_name, frame_type, position_cost, orientation_cost, gain=gain, lm_damping=lm_damping)
        self.root_name = root_name
        self.root_type = root_type
        self.transform_target_to_root: Optional[SE3] = None

    def set_target(self, transform_target_to_root: SE3) -> None:
        self.transform_target_to_root = transform_target_to_root.copy()

    def set_target_from_configuration(self, configuration: Configuration) -> None:
        self.set_target(configuration.get_transform(self.frame_name, self.frame_type, self.root_name, self.root_type))

    def _spec(self, flat) -> TaskSpec:
        base = super()._spec(flat)
        if self.root_type not in SUPPORTED_FRAMES:
            raise UnsupportedFrame(self.root_type, SUPPORTED_FRAMES)
        try:
            root = flat.frame(self.root_name, self.root_type)
        except KeyError:
            raise InvalidFrame(self.root_name, self.root_type, flat) from None
        return TaskSpec(TASK_RELATIVE_FRAME, frame=base.frame, root=root, cost=self.cost.copy(), gain=self.gain,
                        lm_damping=self.lm_damping)

    def _target(self):
        if self.transform_target_to_root is None:
            raise TargetNotSet(self.__class__.__name__)
        return self.transform_target_to_root.wxyz_xyz


class PostureTask(Task):
    """Regulate joint angles towards a posture; floating-base coordinates are not affected
    (reference tasks/posture_task.py)."""

    def __init__(self, model, cost, gain: float = 1.0, lm_damping: float = 0.0):
        flat = as_flat(model)
        super().__init__(cost=np.zeros((flat.nv,)), gain=gain, lm_damping=lm_damping)
        self.target_q: Optional[np.ndarray] = None
        self.k = flat.nv
        self.nq = flat.nq
        self._flat = flat
        self.set_cost(cost)

    def set_cost(self, cost) -> None:
        cost = np.atleast_1d(cost)
        if cost.ndim != 1 or cost.shape[0] not in (1, self.k):
            raise TaskDefinitionError(f"{self.__class__.__name__} cost must be a vector of shape (1,) "
                                      f"(aka identical cost for all dofs) or ({self.k},). Got {cost.shape}")
        if not np.all(cost >= 0.0):
            raise TaskDefinitionError(f"{self.__class__.__name__} cost should be >= 0")
        self.cost[: self.k] = cost

    def set_target(self, target_q) -> None:
        target_q = np.atleast_1d(np.asarray(target_q.detach().cpu() if hasattr(target_q, "detach") else target_q, dtype=np.float64))
        if target_q.shape[-1] != self.nq or target_q.ndim > 2:
            raise InvalidTarget(f"Expected target posture to have shape ({self.nq},) but got {target_q.shape}")
        self.target_q = target_q.copy()

    def set_target_from_configuration(self, configuration: Configuration) -> None:
        q = configuration.q
        self.set_target(q.cpu().numpy() if hasattr(q, "cpu") else q)

    def _spec(self, flat) -> TaskSpec:
        return TaskSpec(TASK_POSTURE, dof_cost=self.cost.copy(), gain=self.gain, lm_damping=self.lm_damping)

    def _target(self):
        if self.target_q is None:
            raise TargetNotSet(self.__class__.__name__)
        return self.target_q

    def _target_kwargs(self, configuration):
        return dict(posture_targets=self._target())

    def compute_error(self, configuration: Configuration):
        """e = q* (-) q with free-joint dofs zeroed (posture_task.py:87-118)."""
        _, J, e, ep = self._evaluate(configuration)
        return _out(configuration, ep[:, 0])

    def compute_jacobian(self, configuration: Configuration):
        """-I with free-joint columns zeroed (posture_task.py:120-142).  Constant, so built on the host."""
        self._target()
        jac = -np.eye(configuration.nv)
        free = [d for d in range(self._flat.nv) if self._flat.node_type[self._flat.dof_node[d]] == 0]
        jac[:, free] = 0.0
        if not configuration.batched:
            return jac
        import torch

        B = configuration.q_device.shape[0]
        return torch.tensor(jac, dtype=torch.float32, device=configuration.q_device.device).expand(B, -1, -1)


class DampingTask(PostureTask):
    """Minimise joint velocities: PostureTask with gain 0 and target qpos0 (tasks/damping_task.py:11-20)."""

    def __init__(self, model, cost):
        super().__init__(model=model, cost=cost, gain=0.0, lm_damping=0.0)
        self.target_q = np.asarray(self._flat.qpos0, dtype=np.float64).copy()


class ComTask(Task):
    """Regulate the centre of mass of the robot (subtree of body 1) -- reference tasks/com_task.py."""

    k: int = 3

    def __init__(self, cost, gain: float = 1.0, lm_damping: float = 0.0):
        super().__init__(cost=np.zeros((self.k,)), gain=gain, lm_damping=lm_damping)
        self.target_com: Optional[np.ndarray] = None
        self.set_cost(cost)

    def set_cost(self, cost) -> None:
        cost = np.atleast_1d(cost)
        if cost.ndim != 1 or cost.shape[0] not in (1, self.k):
            raise TaskDefinitionError(f"{self.__class__.__name__} cost must be a vector of shape (1,) "
                                      f"(aka identical cost for all coordinates) or ({self.k},). Got {cost.shape}")
        if not np.all(cost >= 0.0):
            raise TaskDefinitionError(f"{self.__class__.__name__} cost must be >= 0")
        self.cost[:] = cost

    def set_target(self, target_com) -> None:
        target_com = np.atleast_1d(np.asarray(target_com.detach().cpu() if hasattr(target_com, "detach") else target_com,
                                              dtype=np.float64))
        if target_com.shape[-1] != self.k or target_com.ndim > 2:
            raise InvalidTarget(f"Expected target CoM to have shape ({self.k},) but got {target_com.shape}")
        self.target_com = target_com.copy()

    def set_target_from_configuration(self, configuration: Configuration) -> None:
        self.set_target(configuration.get_com())

    def _spec(self, flat) -> TaskSpec:
        if getattr(flat, "com_missing", None):
            raise TaskDefinitionError(
                f"{self.__class__.__name__}: the model was compiled from MJCF without mesh/geom-derived inertia and "
                f"{len(flat.com_missing)} bodies of the robot have no <inertial> (e.g. {flat.com_missing[:3]}): their mass would be "
                "taken as 0 and the centre of mass would be wrong.  Pass a real mujoco.MjModel (its body_mass is complete).")
        cost = np.zeros(6)
        cost[:3] = self.cost
        return TaskSpec(TASK_COM, cost=cost, gain=self.gain, lm_damping=self.lm_damping)

    def _target(self):
        if self.target_com is None:
            raise TargetNotSet(self.__class__.__name__)
        return self.target_com

    def _target_kwargs(self, configuration):
        B = configuration.q_device.shape[0]
        return dict(com_targets=np.broadcast_to(np.asarray(self._target()).reshape(-1, 1, 3), (B, 1, 3)))

    def compute_error(self, configuration: Configuration):
        """e = com - target (com_task.py:71-82)."""
        _, J, e, _ = self._evaluate(configuration)
        return _out(configuration, e)

    def compute_jacobian(self, configuration: Configuration):
        """mj_jacSubtreeCom of body 1 (com_task.py:84-97)."""
        _, J, e, _ = self._evaluate(configuration)
        return _out(configuration, J)
