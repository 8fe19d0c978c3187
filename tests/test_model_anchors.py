"""MODEL PINNING -- the flattened models every parity test builds on (mink_b200/models/*.bikm, produced by the MJCF compiler
the oracle's mujoco shim SHARES with the product) checked against numbers that compiler did not produce:

  * SURVEY.md 8(c) anchors, derived by the surveyor with an independent throw-away shim: UR5e `home` site pose, G1 `stand`
    site positions, subtree CoM and mass;
  * hand derivations straight from the vendored MJCF text (link offsets and joint angles quoted below): Spot's front-left foot
    at the `home` keyframe, and length invariants of the Shadow Hand's first finger.

Evaluated with the C oracle's FK (CPU); the GPU parity tests then tie the kernels to the same blobs.
"""

import numpy as np

from mink_b200._abi import ProblemSpec
from mink_b200.workloads import load_flat
from oracle.ikoracle import Oracle


def _fk(robot, q, frames):
    fm = load_flat(robot)
    orc = Oracle(fm.to_blob(), ProblemSpec([], []), fm.nq, fm.nv)
    poses, com = orc.fk(np.atleast_2d(q), [fm.frame(n, t) for n, t in frames])
    return fm, poses[0], com[0]


def test_ur5e_home_site_pose_survey_anchor():
    # attachment_site wxyz_xyz at keyframe `home` (examples/universal_robots_ur5e/ur5e.xml:131)
    fm, poses, _ = _fk("ur5e", load_flat("ur5e").key("home"), [("attachment_site", "site")])
    ref = np.array([1.8366e-06, 1.0, -1.8366e-06, 1.8366e-06, 0.491999298412, 0.133997825466, 0.488000367319])
    p = poses[0] * np.sign(poses[0][1])   # quaternion sign is not canonical
    np.testing.assert_allclose(p[4:], ref[4:], atol=1e-9)
    np.testing.assert_allclose(p[:4], ref[:4], atol=2e-9)


def test_g1_stand_survey_anchors():
    # G1 `stand` (examples/unitree_g1/g1.xml:422-432)
    fm = load_flat("g1")
    _, poses, com = _fk("g1", fm.key("stand"), [("left_foot", "site"), ("right_palm", "site")])
    np.testing.assert_allclose(poses[0][4:], [5.874e-07, 0.1178713, 0.030324088559], atol=1e-9)
    np.testing.assert_allclose(poses[0][:4] * np.sign(poses[0][0]), [1, 0, 0, 0], atol=1e-5)      # "quat ~ identity"
    np.testing.assert_allclose(poses[1][4:], [0.26396, -0.162078039168, 0.834913034466], atol=1e-8)
    np.testing.assert_allclose(com, [0.031266998892, 0.000535253418, 0.659741641071], atol=1e-9)
    np.testing.assert_allclose(float(np.sum(fm.com_mass[fm.com_node >= 0])), 32.2389206, atol=1e-6)
    assert (fm.nq, fm.nv, fm.nnode) == (44, 43, 38)


def test_spot_front_left_foot_hand_derived():
    """examples/boston_dynamics_spot/spot_arm.xml:94-121,290-296.  body at (0, 0, 0.46), identity; fl_hip +(0.29785, 0.055, 0),
    hx = 0; fl_uleg +(0, 0.1108, 0), hy = 1.04 about y; fl_lleg +(0.025, 0, -0.32) in the upper-leg frame, kn = -1.8 about y;
    foot (class `foot` default, :80-83) at (0, 0, -0.3365) in the lower-leg frame."""
    ry = lambda t, v: np.array([v[0] * np.cos(t) + v[2] * np.sin(t), v[1], -v[0] * np.sin(t) + v[2] * np.cos(t)])
    p = np.array([0.0, 0.0, 0.46]) + [0.29785, 0.055, 0.0] + np.array([0.0, 0.1108, 0.0])
    p = p + ry(1.04, np.array([0.025, 0.0, -0.32]))
    p = p + ry(1.04 - 1.8, np.array([0.0, 0.0, -0.3365]))
    fm = load_flat("spot")
    _, poses, _ = _fk("spot", fm.key("home"), [("FL", "geom"), ("FL", "site"), ("FR", "geom")])
    np.testing.assert_allclose(poses[0][4:], p, atol=1e-12)
    np.testing.assert_allclose(poses[1][4:], p, atol=1e-12)
    np.testing.assert_allclose(poses[2][4:], p * [1, -1, 1], atol=1e-12)     # the right leg mirrors the left one in y
    assert (fm.nq, fm.nv) == (26, 25)


def test_shadow_first_finger_length_invariants():
    """examples/shadow_hand/left_hand.xml:128-150: lh_ffknuckle -> lh_ffproximal (+0) -> lh_ffmiddle (+0.045 z) -> lh_ffdistal
    (+0.025 z) -> site `first` (+0.025 z), all child frames unrotated; FFJ3/FFJ2/FFJ1 turn about x (class defaults :6-60).
    Distances between the knuckle origin and the fingertip do not depend on anything above the knuckle."""
    fm = load_flat("shadow")
    names = fm.names["joint"]
    q = np.zeros(fm.nq)

    def dist(qq):
        _, poses, _ = _fk("shadow", qq, [("lh_ffknuckle", "body"), ("first", "site")])
        return np.linalg.norm(poses[1][4:] - poses[0][4:])

    assert abs(dist(q) - 0.095) < 1e-12                                   # straight finger
    q2 = q.copy(); q2[int(fm.node_qadr[names.index("lh_FFJ2")])] = np.pi / 2
    assert abs(dist(q2) - np.hypot(0.045, 0.05)) < 1e-12                  # bent 90 deg at the middle joint
    q3 = q.copy(); q3[int(fm.node_qadr[names.index("lh_FFJ3")])] = 0.7    # rotating at the base keeps the length
    assert abs(dist(q3) - 0.095) < 1e-12
    q4 = q.copy(); q4[int(fm.node_qadr[names.index("lh_FFJ1")])] = np.pi / 2
    assert abs(dist(q4) - np.hypot(0.07, 0.025)) < 1e-12                  # only the distal phalanx folds
