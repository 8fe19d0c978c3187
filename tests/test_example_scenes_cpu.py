"""Every MJCF scene under the reference's examples/ compiles with mink_b200's MJCF subset and flattens into a kinematic tree
(DESIGN.md section 0, row f3).  Needs the reference checkout; skipped on machines without it (the GPU box)."""

import glob
import os

import numpy as np
import pytest

import mink_b200 as mink

EXAMPLES = "/root/reference/examples"
# keyframe-only / actuator-only fragments meant to be <include>d into a model, not models of their own
FRAGMENTS = {"aloha/keyframe_no_act.xml", "aloha/keyframe_ctrl.xml", "shadow_hand/keyframes.xml",
             "aloha/filtered_cartesian_actuators.xml", "aloha/joint_position_actuators.xml"}
SCENES = sorted(p for p in glob.glob(os.path.join(EXAMPLES, "*", "*.xml")) if os.path.relpath(p, EXAMPLES) not in FRAGMENTS)


@pytest.mark.skipif(not SCENES, reason="the reference's example scenes are not on this machine")
@pytest.mark.parametrize("path", SCENES, ids=[os.path.relpath(p, EXAMPLES) for p in SCENES])
def test_scene_compiles_and_flattens(path):
    model = mink.Model.from_xml_path(path)
    assert model.nq >= model.nv > 0 and model.nbody > 1
    fm = mink.flatten(model)
    assert fm.nq == model.nq and fm.nv == model.nv
    # parents come first, every dof belongs to exactly one node, the blob round-trips
    assert all(int(fm.node_parent[n]) < n for n in range(fm.nnode))
    assert len(fm.dof_node) == fm.nv and all(0 <= int(d) < fm.nnode for d in fm.dof_node)
    back = mink.FlatModel.from_blob(fm.to_blob(), fm.to_meta_json())
    np.testing.assert_array_equal(back.node_parent, fm.node_parent)
    np.testing.assert_array_equal(back.qpos0, fm.qpos0)
    assert back.names["body"] == fm.names["body"]
