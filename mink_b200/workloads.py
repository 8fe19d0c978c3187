"""The five BASELINE.json configurations as data, plus the synthetic-input generator.

Each workload names the robot, the task list, the limits and the solver settings taken from the
reference examples (file:line cited per entry), following SURVEY.md 8(d).  The same dictionaries
drive (a) the golden-vector generator, which instantiates the *reference's* task/limit classes,
(b) this package's own task/limit classes, (c) bench.py.
"""

from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Tuple

import numpy as np

PI = float(np.pi)

# Example robots of the reference beyond the BASELINE configs (goldens tests/golden/<name>.npz, models mink_b200/models/<name>.*)
EXAMPLE_ROBOTS = ["iiwa", "h1", "go1", "stretch", "tidybot", "aloha", "leap"]

WORKLOADS: Dict[str, dict] = {
    # BASELINE config 1: UR5e single instance, FrameTask + PostureTask + ConfigurationLimit
    # (examples/arm_ur5e.py:20-28,66,74: dt = 1/500, damping 1e-3, lm_damping 1).
    "ur5e": dict(
        robot="ur5e", scene="universal_robots_ur5e/scene.xml", key="home",
        frames=[dict(name="attachment_site", type="site", position_cost=1.0, orientation_cost=1.0,
                     lm_damping=1.0)],
        posture=dict(cost=1e-2), com=None,
        limits=[dict(kind="configuration", gain=0.95)],
        dt=2e-3, damping=1e-3, batch=1,
    ),
    # BASELINE config 2: UR5e batch 4096, same tasks, no inequality limits (pure damped LS).
    "ur5e_dls": dict(
        robot="ur5e", scene="universal_robots_ur5e/scene.xml", key="home",
        frames=[dict(name="attachment_site", type="site", position_cost=1.0, orientation_cost=1.0,
                     lm_damping=1.0)],
        posture=dict(cost=1e-2), com=None, limits=[],
        dt=2e-3, damping=1e-3, batch=4096,
    ),
    # BASELINE config 3 (headline): G1, 3 FrameTasks + PostureTask + configuration/velocity limits
    # (examples/humanoid_g1.py:22-40,80,88: foot tasks 200/10 lm 1, pelvis 0/10, posture 1,
    #  dt = 1/200, damping 1e-1; velocity limit pi rad/s as tests/test_velocity_limit.py:23-25).
    "g1": dict(
        robot="g1", scene="unitree_g1/scene.xml", key="stand",
        frames=[dict(name="left_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="right_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="pelvis", type="body", position_cost=0.0, orientation_cost=10.0, lm_damping=0.0)],
        posture=dict(cost=1.0), com=None,
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI)],
        dt=5e-3, damping=1e-1, batch=65536,
    ),
    # BASELINE config 4: Shadow Hand, 5 fingertip FrameTasks + PostureTask, joint limits
    # (examples/hand_shadow.py:19-35,53,63: position cost 1, lm 1, posture 1e-2, dt 1/500, damping 1e-5).
    "shadow": dict(
        robot="shadow", scene="shadow_hand/scene_left.xml", key="grasp hard",
        frames=[dict(name=f, type="site", position_cost=1.0, orientation_cost=0.0, lm_damping=1.0)
                for f in ("thumb", "first", "middle", "ring", "little")],
        posture=dict(cost=1e-2), com=None,
        limits=[dict(kind="configuration", gain=0.95)],
        dt=2e-3, damping=1e-5, batch=16384,
    ),
    # BASELINE config 5: Spot, 4 foot FrameTasks (geoms) + ComTask + CollisionAvoidanceLimit
    # (examples/quadruped_spot.py:37-42; ComTask cost as examples/humanoid_g1.py:29).
    "spot": dict(
        robot="spot", scene="boston_dynamics_spot/scene.xml", key="home",
        frames=[dict(name=f, type="geom", position_cost=1.0, orientation_cost=0.0, lm_damping=0.0)
                for f in ("FL", "FR", "HR", "HL")],
        posture=None, com=dict(cost=200.0),
        limits=[dict(kind="collision", pairs=[(["FL", "FR", "HR", "HL"], ["floor"]),
                                              (["FL", "FR", "HR", "HL"], ["FL", "FR", "HR", "HL"])],
                     gain=0.85, minimum_distance=0.005, detection_distance=0.3, bound_relaxation=0.0)],
        dt=2e-3, damping=1e-3, batch=32768,
    ),
    # Not a BASELINE config: exercises RelativeFrameTask (SURVEY.md 8f "next" row 1) -- left palm regulated
    # relative to the right palm, as the bimanual examples do (examples/arm_hand_iiwa_allegro.py:75-83).
    "g1_rel": dict(
        robot="g1", scene="unitree_g1/scene.xml", key="stand",
        frames=[dict(name="pelvis", type="body", position_cost=0.0, orientation_cost=10.0, lm_damping=0.0)],
        relative_frames=[dict(name="left_palm", type="site", root_name="right_palm", root_type="site",
                              position_cost=5.0, orientation_cost=1.0, lm_damping=0.5)],
        posture=dict(cost=1.0), com=None,
        limits=[dict(kind="configuration", gain=0.95)],
        dt=5e-3, damping=1e-2, batch=4096,
    ),
    # Not a BASELINE config: the reference's humanoid example as written (examples/humanoid_g1.py:20-50) -- pelvis
    # orientation, posture, CoM, both feet AND both palms -- with the limits of the headline config.  The CoM task couples
    # every dof (43 coupled: general warp-per-problem path).
    "g1_full": dict(
        robot="g1", scene="unitree_g1/scene.xml", key="stand",
        frames=[dict(name="pelvis", type="body", position_cost=0.0, orientation_cost=10.0, lm_damping=0.0),
                dict(name="right_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="left_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="right_palm", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0),
                dict(name="left_palm", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0)],
        posture=dict(cost=1.0), com=dict(cost=200.0),
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI)],
        dt=5e-3, damping=1e-1, batch=4096,
    ),
    # Same example without the CoM task: 31 coupled dofs (base 6 + legs 12 + waist 1 + arms 12), 6 of them unbounded --
    # the largest block the small-group path takes (active sets are 32-bit masks), 25 dofs after the elimination.
    "g1_hands": dict(
        robot="g1", scene="unitree_g1/scene.xml", key="stand",
        frames=[dict(name="pelvis", type="body", position_cost=0.0, orientation_cost=10.0, lm_damping=0.0),
                dict(name="right_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="left_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="right_palm", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0),
                dict(name="left_palm", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0)],
        posture=dict(cost=1.0), com=None,
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI)],
        dt=5e-3, damping=1e-1, batch=4096,
    ),
    # Not a BASELINE config: DampingTask (reference mink/tasks/damping_task.py:11-20 = PostureTask with gain 0
    # and target qpos0, as examples/mobile_tidybot.py:60 uses it) next to a frame task, with both box limits.
    "ur5e_damp": dict(
        robot="ur5e", scene="universal_robots_ur5e/scene.xml", key="home",
        frames=[dict(name="attachment_site", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=None, damping_task=dict(cost=0.3), com=None,
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=2 * PI)],
        dt=2e-2, damping=1e-4, batch=1024,
    ),
    # Not a BASELINE config: the limit set of examples/arm_ur5e.py:29-47 (wrist capsule against the floor plane and the wall BOX,
    # configuration and velocity limits); detection distance widened so that sampled configurations have active rows.
    "ur5e_wall": dict(
        robot="ur5e", scene="universal_robots_ur5e/scene.xml", key="home",
        frames=[dict(name="attachment_site", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=dict(cost=1e-2), com=None,
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI),
                dict(kind="collision", pairs=[(["wrist_3_link"], ["floor", "wall"])],
                     gain=0.85, minimum_distance=0.005, detection_distance=0.5, bound_relaxation=0.0)],
        dt=2e-3, damping=1e-3, batch=1024,
    ),
    # ---- the reference's other example robots (not BASELINE configs): task sets as in examples/*.py, default limits -------
    # examples/arm_iiwa.py:24-33,64 (7-dof arm).
    "iiwa": dict(
        robot="iiwa", scene="kuka_iiwa_14/scene.xml", key="home",
        frames=[dict(name="attachment_site", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=dict(cost=1e-2), com=None, limits=[dict(kind="configuration", gain=0.95)],
        dt=5e-3, damping=1e-3, batch=4096,
    ),
    # examples/humanoid_h1.py:20-54,88 (floating base, CoM task, feet and wrists).
    "h1": dict(
        robot="h1", scene="unitree_h1/scene.xml", key="stand",
        frames=[dict(name="pelvis", type="body", position_cost=0.0, orientation_cost=10.0, lm_damping=0.0),
                dict(name="right_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="left_foot", type="site", position_cost=200.0, orientation_cost=10.0, lm_damping=1.0),
                dict(name="right_wrist", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0),
                dict(name="left_wrist", type="site", position_cost=200.0, orientation_cost=0.0, lm_damping=1.0)],
        posture=dict(cost=1.0), com=dict(cost=200.0), limits=[dict(kind="configuration", gain=0.95)],
        dt=5e-3, damping=1e-1, batch=4096,
    ),
    # examples/quadruped_go1.py:19-38,70 (floating base, trunk + four feet, posture cost 1e-5, damping 1e-5).
    "go1": dict(
        robot="go1", scene="unitree_go1/scene.xml", key="home",
        frames=[dict(name="trunk", type="body", position_cost=1.0, orientation_cost=1.0, lm_damping=0.0)] +
               [dict(name=f, type="site", position_cost=1.0, orientation_cost=0.0, lm_damping=0.0) for f in ("FL", "FR", "RR", "RL")],
        posture=dict(cost=1e-5), com=None, limits=[dict(kind="configuration", gain=0.95)],
        dt=5e-3, damping=1e-5, batch=4096,
    ),
    # examples/mobile_stretch.py:18-33,73 (mobile base + lift / telescoping SLIDE joints; no posture task).
    "stretch": dict(
        robot="stretch", scene="hello_robot_stretch_3/scene.xml", key="home",
        frames=[dict(name="base_link", type="body", position_cost=0.1, orientation_cost=1.0, lm_damping=0.0),
                dict(name="link_grasp_center", type="site", position_cost=1.0, orientation_cost=1e-4, lm_damping=0.0)],
        posture=None, com=None, limits=[dict(kind="configuration", gain=0.95)],
        dt=1e-2, damping=1e-3, batch=4096,
    ),
    # examples/mobile_tidybot.py:45-70,111 (planar base + 7-dof arm + gripper; per-dof posture cost: 0 on the base dofs).
    "tidybot": dict(
        robot="tidybot", scene="stanford_tidybot/scene.xml", key="home",
        frames=[dict(name="pinch_site", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=dict(cost=[0.0] * 3 + [1e-3] * 15), com=None, limits=[dict(kind="configuration", gain=0.95)],
        dt=1e-2, damping=1e-3, batch=4096,
    ),
    # examples/arm_aloha.py:75-115,153 (two 6-dof arms + grippers; configuration + velocity limits; the example's collision
    # pairs include mesh geoms and are left out).
    "aloha": dict(
        robot="aloha", scene="aloha/scene.xml", key="neutral_pose",
        frames=[dict(name="left/gripper", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0),
                dict(name="right/gripper", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=dict(cost=1e-4), com=None, limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI)],
        dt=5e-3, damping=1e-5, batch=4096,
    ),
    # examples/arm_hand_xarm_leap.py:75-85: the LEAP hand's four fingertip tasks (position only) + posture, on the hand alone (the
    # example attaches it to an arm with MjSpec, which the MJCF subset here does not do).
    "leap": dict(
        robot="leap", scene="leap_hand/scene_right.xml", key=None,
        frames=[dict(name=f, type="site", position_cost=1.0, orientation_cost=0.0, lm_damping=1.0) for f in ("tip_1", "tip_2", "tip_3", "th_tip")],
        posture=dict(cost=5e-2), com=None, limits=[dict(kind="configuration", gain=0.95)],
        q_spread=0.3, dt=5e-3, damping=1e-3, batch=16384,
    ),
    # examples/arm_aloha.py as written: the same tasks and limits PLUS its CollisionAvoidanceLimit (:95-110) -- wrist subtree
    # against wrist subtree, both arm subtrees against the metal frame and the table: 1 104 geom pairs after the reference's
    # filtering (capsule-capsule, sphere-capsule, sphere-sphere, box-capsule, box-sphere).
    "aloha_coll": dict(
        robot="aloha", scene="aloha/scene.xml", key="neutral_pose",
        frames=[dict(name="left/gripper", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0),
                dict(name="right/gripper", type="site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)],
        posture=dict(cost=1e-4), com=None,
        limits=[dict(kind="configuration", gain=0.95), dict(kind="velocity", vmax=PI),
                dict(kind="collision", pairs=[(["subtree:left/wrist_link"], ["subtree:right/wrist_link"]),
                                              (["subtree:left/upper_arm_link", "subtree:right/upper_arm_link"], ["body:metal_frame", "table"])],
                     gain=0.85, minimum_distance=0.05, detection_distance=0.1, bound_relaxation=0.0)],
        q_spread=0.35, dt=5e-3, damping=1e-5, batch=1024,
    ),
    # Not a BASELINE config: edge-case model authored for this repository (mink_b200/models/edge.xml):
    # ball joint with off-centre anchor, slide joint with ref, two joints on one body, a second floating
    # root, capsule/sphere/plane collision pairs, every task and limit kind at once.
    "edge": dict(
        robot="edge", scene="@mink_b200/models/edge.xml", key="home",
        frames=[dict(name="tool", type="site", position_cost=[3.0, 2.0, 1.0], orientation_cost=0.7, lm_damping=0.3, gain=0.8),
                dict(name="float_site", type="site", position_cost=1.0, orientation_cost=[0.5, 0.0, 0.2], lm_damping=0.0)],
        relative_frames=[dict(name="side_tip", type="site", root_name="tool", root_type="site",
                              position_cost=2.0, orientation_cost=0.0, lm_damping=0.1)],
        posture=dict(cost=0.5), com=dict(cost=2.0),
        limits=[dict(kind="configuration", gain=0.9), dict(kind="velocity", vmax=2.0),
                dict(kind="collision", pairs=[(["g_hand", "g_fore", "g_upper"], ["g_side"]), (["g_float"], ["floor", "g_hand"])],
                     gain=0.85, minimum_distance=0.01, detection_distance=0.6, bound_relaxation=0.0)],
        dt=1e-2, damping=1e-3, batch=1024,
    ),
}


def resolve_geom_groups(model, groups):
    """Collision groups of a workload: plain geom names / ids, or "subtree:<body>" / "body:<body>" for what the reference's
    examples build with get_subtree_geom_ids / get_body_geom_ids (examples/arm_aloha.py:95-102).  `model` is a FlatModel, a
    mink_b200.Model or an MjModel-like object."""
    from .utils import get_body_geom_ids, get_subtree_geom_ids

    def body_id(name):
        names = getattr(model, "names", None)
        return names["body"].index(name) if isinstance(names, dict) else model.body(name).id

    def expand(group):
        out = []
        for g in group:
            if isinstance(g, str) and g.startswith("subtree:"):
                out += get_subtree_geom_ids(model, body_id(g[8:]))
            elif isinstance(g, str) and g.startswith("body:"):
                out += get_body_geom_ids(model, body_id(g[5:]))
            else:
                out.append(g)
        return out

    return [(expand(a), expand(b)) for a, b in groups]


def _quat_exp(w: np.ndarray) -> np.ndarray:
    """exp of rotation vectors [...,3] -> unit quaternions wxyz."""
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.maximum(th, 1e-300), 0.5)
    return np.concatenate([np.cos(half), k * w], axis=-1)


def sample_q(fm, key_q: np.ndarray, B: int, rng: np.random.Generator) -> np.ndarray:
    """q per SURVEY.md 8(d): scalar joints uniform in the inner 80 % of their range (unlimited:
    +-0.8 pi), free joint = keyframe position + U(-0.2,0.2)^3 with orientation exp(N(0,0.2^2)^3)."""
    q = np.tile(key_q, (B, 1)).astype(np.float64)
    for n in range(fm.nnode):
        t, a, d = int(fm.node_type[n]), int(fm.node_qadr[n]), int(fm.node_dadr[n])
        if t == 0:
            q[:, a:a + 3] = key_q[a:a + 3] + rng.uniform(-0.2, 0.2, size=(B, 3))
            q[:, a + 3:a + 7] = _quat_exp(rng.normal(0.0, 0.2, size=(B, 3)))
        elif t == 1:
            q[:, a:a + 4] = _quat_exp(rng.normal(0.0, 0.2, size=(B, 3)))
        else:
            if fm.dof_limited[d]:
                lo, hi = fm.dof_lo[d], fm.dof_hi[d]
            else:
                lo, hi = -PI, PI
            q[:, a] = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo), size=B)
    return q


def perturb_q(fm, q: np.ndarray, sigma: float, rng: np.random.Generator) -> np.ndarray:
    """q' = q (+) delta, delta ~ N(0, sigma^2) on the scalar (hinge/slide) dofs only."""
    qp = q.copy()
    for n in range(fm.nnode):
        if int(fm.node_type[n]) >= 2:
            a = int(fm.node_qadr[n])
            qp[:, a] += rng.normal(0.0, sigma, size=q.shape[0])
    return qp


FkFn = Callable[[np.ndarray], Tuple[np.ndarray, Optional[np.ndarray]]]

MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def load_flat(robot: str):
    """Flattened model of a BASELINE robot (mink_b200/models/*.bikm + *.json, written by oracle/gen_golden.py from the
    reference's vendored MJCFs): what bench.py and the tests build their problems from on a box without /root/reference."""
    from .flatten import FlatModel

    with open(os.path.join(MODELS_DIR, robot + ".bikm"), "rb") as f:
        blob = f.read()
    with open(os.path.join(MODELS_DIR, robot + ".json")) as f:
        meta = f.read()
    return FlatModel.from_blob(blob, meta)


def task_frames(wl: dict, fm):
    """Frames of the workload's (absolute) frame tasks, in task order."""
    return [fm.frame(f["name"], f["type"]) for f in wl["frames"]]


def make_inputs(fm, wl: dict, B: int, fk: FkFn, seed: int = 0, sigma: float = 0.1,
                aligned_fraction: float = 0.01) -> dict:
    """Synthetic batch for a workload.

    `fk(q[B,nq]) -> (frame poses [B,F,7] wxyz_xyz, com [B,3] or None)` is supplied by the caller
    (the oracle in tests/golden generation, the CUDA FK in bench.py).  A fraction of the instances
    gets a target orientation *exactly* equal to the current one, which exercises the reference's
    identity branch of SE3.ljacinv (mink/lie/se3.py:213).
    """
    rng = np.random.default_rng(seed)
    key_q = fm.key(wl["key"]) if wl.get("key") else np.asarray(fm.qpos0, dtype=np.float64).copy()   # no keyframe: the model's qpos0
    if wl.get("q_spread") is not None:   # configurations around the keyframe (an arm workspace, not the whole joint range)
        q = perturb_q(fm, np.tile(key_q, (B, 1)).astype(np.float64), float(wl["q_spread"]), rng)
        for d in range(fm.nv):
            qa = int(fm.dof_qadr[d])
            if qa >= 0 and fm.dof_limited[d]:
                q[:, qa] = np.clip(q[:, qa], fm.dof_lo[d] + 1e-3, fm.dof_hi[d] - 1e-3)
    else:
        q = sample_q(fm, key_q, B, rng)
    qp = perturb_q(fm, q, sigma, rng)
    poses_now, _ = fk(q)
    poses_tgt, com_tgt = fk(qp)
    targets = np.array(poses_tgt, dtype=np.float64, copy=True)
    n_al = int(round(aligned_fraction * B))
    if B >= 8 and n_al == 0 and aligned_fraction > 0:
        n_al = 1
    if n_al:
        idx = rng.choice(B, size=n_al, replace=False)
        targets[idx, :, :4] = poses_now[idx, :, :4]
    # a DampingTask's "target" is qpos0 (damping_task.py:20); it travels in the posture-target slot
    ptgt = np.asarray(fm.qpos0, dtype=np.float64).copy() if (wl.get("damping_task") is not None and wl.get("posture") is None) else key_q.copy()
    out = dict(q=q, frame_targets=targets, posture_target=ptgt)
    if wl.get("com") is not None:
        out["com_target"] = np.array(com_tgt, dtype=np.float64, copy=True)
    return out
