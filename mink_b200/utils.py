"""Model-introspection helpers with the reference's names (mink/utils.py:38-174): what its examples call to set up
PostureTask free-joint masks and CollisionAvoidanceLimit geom groups.  Host-side, set-up time only; `model` is a
`mink_b200.Model` (the MJCF compiler's output).  `move_mocap_to_frame` belongs to the interactive viewer loop and is not
rebuilt."""

from typing import List, Optional, Tuple

import numpy as np

from .exceptions import InvalidKeyframe

_JNT_FREE, _JNT_BALL = 0, 1


def get_freejoint_dims(model) -> Tuple[List[int], List[int]]:
    """All floating-joint indices in configuration and tangent space (utils.py:38-56)."""
    q_ids: List[int] = []
    v_ids: List[int] = []
    for j in range(model.njnt):
        if int(model.jnt_type[j]) == _JNT_FREE:
            qadr, vadr = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
            q_ids.extend(range(qadr, qadr + 7))
            v_ids.extend(range(vadr, vadr + 6))
    return q_ids, v_ids


def custom_configuration_vector(model, key_name: Optional[str] = None, **kwargs) -> np.ndarray:
    """qpos0 (or a keyframe) with the named joints set to the given values (utils.py:59-97)."""
    if key_name is not None:
        if key_name not in list(model.key_names):
            raise InvalidKeyframe(key_name, model)
        q = np.array(model.key_qpos[list(model.key_names).index(key_name)], dtype=np.float64)
    else:
        q = np.array(model.qpos0, dtype=np.float64)
    for name, value in kwargs.items():
        jid = model.joint(name).id
        t = int(model.jnt_type[jid])
        dim = 7 if t == _JNT_FREE else (4 if t == _JNT_BALL else 1)
        value = np.atleast_1d(value)
        if value.shape != (dim,):
            raise ValueError(f"Joint {name} should have a qpos value of {dim,} but got {value.shape}")
        qid = int(model.jnt_qposadr[jid])
        q[qid:qid + dim] = value
    return q


def _children(model) -> List[List[int]]:
    """children[b] = bodies whose parent is b, ascending (the world body is its own parent in MuJoCo's tables)."""
    table: List[List[int]] = [[] for _ in range(model.nbody)]
    for i in range(1, model.nbody):
        table[int(model.body_parentid[i])].append(i)
    return table


def _preorder_last_child_first(children: List[List[int]], root: int) -> List[int]:
    """Depth-first order in which the reference's explicit stack visits a subtree: a body, then its children from the highest
    id down, each with its whole subtree (utils.py:126-134 pops the most recently pushed child)."""
    order = [root]
    for c in reversed(children[root]):
        order += _preorder_last_child_first(children, c)
    return order


def get_body_body_ids(model, body_id: int) -> List[int]:
    """Immediate children of a body (utils.py:100-115)."""
    return list(_children(model)[body_id])


def get_subtree_body_ids(model, body_id: int) -> List[int]:
    """The body and all its descendants, in the reference's visiting order (utils.py:118-134)."""
    return _preorder_last_child_first(_children(model), body_id)


def get_body_geom_ids(model, body_id: int) -> List[int]:
    """Geoms attached to a body (utils.py:137-152)."""
    first, count = int(model.body_geomadr[body_id]), int(model.body_geomnum[body_id])
    return [first + k for k in range(count)]


def get_subtree_geom_ids(model, body_id: int) -> List[int]:
    """Geoms of a body and of all its descendants (utils.py:155-174)."""
    return [g for b in get_subtree_body_ids(model, body_id) for g in get_body_geom_ids(model, b)]
