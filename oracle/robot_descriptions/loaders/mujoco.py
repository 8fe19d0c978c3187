"""ORACLE / TEST INFRASTRUCTURE -- maps `load_robot_description(name)` to the MJCFs vendored
under /root/reference/examples (the real package git-clones mujoco_menagerie; no network here).

UR5e: the menagerie model the reference tests use names joints `shoulder_pan_joint`, ... while the
vendored copy uses `shoulder_pan` -- the `_joint` suffix is added on load.
"""

import os

import mujoco

_EX = os.environ.get("MINK_REFERENCE_EXAMPLES", "/root/reference/examples")
_MAP = {
    "ur5e_mj_description": ("universal_robots_ur5e/ur5e.xml", "_joint"),
    "g1_mj_description": ("unitree_g1/g1.xml", None),
    "stretch_mj_description": ("hello_robot_stretch_3/stretch.xml", None),
    "stretch_3_mj_description": ("hello_robot_stretch_3/stretch.xml", None),
    "shadow_hand_mj_description": ("shadow_hand/left_hand.xml", None),
    "spot_mj_description": ("boston_dynamics_spot/spot_arm.xml", None),
    "go1_mj_description": ("unitree_go1/go1.xml", None),
    "h1_mj_description": ("unitree_h1/h1.xml", None),
    "iiwa14_mj_description": ("kuka_iiwa_14/iiwa14.xml", None),
    "talos_mj_description": ("unitree_h1/h1.xml", None),
}


def load_robot_description(name: str, variant=None):
    rel, suffix = _MAP[name]
    model = mujoco.MjModel.from_xml_path(os.path.join(_EX, rel))
    if suffix:
        model.joint_names = [n + suffix if n and not n.endswith(suffix) else n for n in model.joint_names]
    return model
