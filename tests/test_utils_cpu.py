"""mink_b200.utils: the reference's model-introspection helpers (mink/utils.py:38-174) on the in-repo edge-case model, with
expectations read off mink_b200/models/edge.xml by hand, and -- where the reference's example scenes are present -- against
check sums of what the reference's own functions return for them (computed once with /root/reference on the numpy shims)."""

import os
import zlib

import numpy as np
import pytest

import mink_b200 as mink
from mink_b200.workloads import MODELS_DIR

EXAMPLES = "/root/reference/examples"


def _edge():
    return mink.Model.from_xml_path(os.path.join(MODELS_DIR, "edge.xml"))


def test_body_and_geom_id_helpers_on_the_edge_model():
    m = _edge()
    # bodies: 0 world, 1 base, 2 shoulder, 3 fore, 4 hand, 5 second, 6 floater;  geoms: 0 floor, 1 g_base, 2 g_upper, 3 g_fore,
    # 4 g_hand, 5 g_side, 6 g_float
    assert [m.body(n).id for n in ("base", "shoulder", "fore", "hand", "second", "floater")] == [1, 2, 3, 4, 5, 6]
    assert mink.get_body_body_ids(m, 0) == [1, 6]
    assert mink.get_body_body_ids(m, 1) == [2, 5]
    assert mink.get_body_body_ids(m, 4) == []
    # the reference pops the most recently pushed child first: base, second, shoulder, fore, hand
    assert mink.get_subtree_body_ids(m, 1) == [1, 5, 2, 3, 4]
    assert mink.get_subtree_body_ids(m, 0) == [0, 6, 1, 5, 2, 3, 4]
    assert mink.get_body_geom_ids(m, 1) == [1] and mink.get_body_geom_ids(m, 0) == [0]
    assert mink.get_subtree_geom_ids(m, 2) == [2, 3, 4]
    assert mink.get_subtree_geom_ids(m, 1) == [1, 5, 2, 3, 4]


def test_freejoint_dims_and_custom_configuration_vector():
    m = _edge()
    q_ids, v_ids = mink.get_freejoint_dims(m)
    assert len(q_ids) == 7 and len(v_ids) == 6
    j = [i for i in range(m.njnt) if int(m.jnt_type[i]) == 0][0]
    assert q_ids[0] == int(m.jnt_qposadr[j]) and v_ids[0] == int(m.jnt_dofadr[j])
    hinge = [i for i in range(m.njnt) if int(m.jnt_type[i]) == 3][0]
    name = m.joint_names[hinge]
    q = mink.custom_configuration_vector(m, **{name: 0.25})
    expect = np.array(m.qpos0, dtype=np.float64)
    expect[int(m.jnt_qposadr[hinge])] = 0.25
    np.testing.assert_array_equal(q, expect)
    qk = mink.custom_configuration_vector(m, "home")
    np.testing.assert_array_equal(qk, m.key_qpos[list(m.key_names).index("home")])
    with pytest.raises(mink.InvalidKeyframe):
        mink.custom_configuration_vector(m, "no such key")
    with pytest.raises(ValueError):
        mink.custom_configuration_vector(m, **{name: [0.1, 0.2]})


@pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="the reference's example scenes are not on this machine")
@pytest.mark.parametrize("scene,body,sizes,crc", [
    ("aloha/scene.xml", "left/wrist_link", [1, 5, 2, 19], 3698714588),
    ("aloha/scene.xml", "right/upper_arm_link", [1, 8, 2, 25], 714669284),
    ("aloha/scene.xml", "metal_frame", [0, 1, 62, 62], 3025689596),
    ("aloha/scene.xml", "world", [5, 24, 4, 126], 1806653886),
    ("unitree_g1/scene.xml", "pelvis", [3, 38, 3, 89], 671185076),
    ("unitree_g1/scene.xml", "torso_link", [2, 25, 6, 56], 1027126641),
])
def test_id_helpers_reproduce_the_reference_on_its_example_scenes(scene, body, sizes, crc):
    m = mink.Model.from_xml_path(os.path.join(EXAMPLES, scene))
    bid = m.body(body).id
    r = (mink.get_body_body_ids(m, bid), mink.get_subtree_body_ids(m, bid), mink.get_body_geom_ids(m, bid), mink.get_subtree_geom_ids(m, bid))
    assert [len(x) for x in r] == sizes
    assert zlib.crc32(repr(r).encode()) == crc
