"""Exception types with the reference's names (mink/exceptions.py, tasks/exceptions.py, limits/exceptions.py).

Inside a *batch* the library reports per-instance status bits instead (include/bik.h); the front
end turns them into these exceptions / warnings where the reference would have raised.
"""

from __future__ import annotations


class MinkError(Exception):
    """Base class of every error raised by this package."""


def _names(model, kind: str):
    attr = {"body": "body_names", "site": "site_names", "geom": "geom_names", "key": "key_names"}[kind]
    if hasattr(model, attr):
        return list(getattr(model, attr))
    if hasattr(model, "names"):
        return list(model.names[kind])
    return []


class UnsupportedFrame(MinkError):
    def __init__(self, frame_type, supported_types):
        super().__init__(f"{frame_type} is not supported.Supported frame types are: {supported_types}")


class InvalidFrame(MinkError):
    def __init__(self, frame_name, frame_type, model):
        super().__init__(f"{frame_type} '{frame_name}' does not exist in the model. "
                         f"Available {frame_type} names: {_names(model, frame_type)}")


class InvalidKeyframe(MinkError):
    def __init__(self, keyframe_name, model):
        super().__init__(f"Keyframe {keyframe_name} does not exist in the model. "
                         f"Available keyframe names: {_names(model, 'key')}")


class InvalidMocapBody(MinkError):
    def __init__(self, mocap_name, model):
        super().__init__(f"Body '{mocap_name}' is not a mocap body.")


class NotWithinConfigurationLimits(MinkError):
    def __init__(self, joint_id, value, lower, upper, model, instance=None):
        names = getattr(model, "joint_names", None) or (model.names["joint"] if hasattr(model, "names") else [])
        name = names[joint_id] if joint_id < len(names) else "?"
        where = "" if instance is None else f" (instance {instance})"
        super().__init__(f"Joint {joint_id} ({name}) violates configuration limits {lower} <= {value} <= {upper}{where}")


class TaskDefinitionError(MinkError):
    """Ill-formed task definition."""


class TargetNotSet(MinkError):
    def __init__(self, cls_name: str):
        super().__init__(f"No target set for {cls_name}")


class InvalidTarget(MinkError):
    """Target of the wrong shape."""


class InvalidGain(MinkError):
    """Gain outside [0, 1]."""


class InvalidDamping(MinkError):
    """Negative Levenberg-Marquardt damping."""


class LimitDefinitionError(MinkError):
    """Ill-formed limit definition."""
