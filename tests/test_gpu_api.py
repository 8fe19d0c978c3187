"""The mink-compatible Python API on the GPU engine, written the way the reference's own tests are
(tests/test_solve_ik.py, test_frame_task.py, test_configuration.py, test_configuration_limit.py ...)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mink_b200 as mink  # noqa: E402
from mink_b200.lie import SE3  # noqa: E402
from tests.helpers import load_case, load_flat  # noqa: E402


@pytest.fixture(autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64) if hasattr(t, "detach") else np.asarray(t)


def _g1_tasks(fm, wl):
    tasks = [mink.FrameTask(f["name"], f["type"], f["position_cost"], f["orientation_cost"], lm_damping=f["lm_damping"])
             for f in wl["frames"]]
    tasks.append(mink.PostureTask(fm, cost=wl["posture"]["cost"]))
    limits = [mink.ConfigurationLimit(fm), mink.VelocityLimit(fm, {n: np.pi for n, t in zip(fm.names["joint"], fm.node_type) if t >= 2})]
    return tasks, limits


def test_single_instance_matches_reference_objects():
    """Single configuration, numpy in / numpy out: every public quantity vs the reference (golden row 0)."""
    wl, fm, spec, g = load_case("g1")
    cfg = mink.Configuration(fm, g["q"][0])
    assert cfg.nq == 44 and cfg.nv == 43 and not cfg.batched
    np.testing.assert_array_equal(cfg.q, g["q"][0])
    tasks, limits = _g1_tasks(fm, wl)
    for k, t in enumerate(tasks[:3]):
        t.set_target(SE3(g["frame_targets"][0, k]))
        e, J = t.compute_error(cfg), t.compute_jacobian(cfg)
        assert e.shape == (6,) and J.shape == (6, 43)
        np.testing.assert_allclose(e, g["e_frame"][0, k], atol=2e-5)
        np.testing.assert_allclose(J, g["J_frame"][0, k], atol=5e-5)
        T = cfg.get_transform_frame_to_world(wl["frames"][k]["name"], wl["frames"][k]["type"])
        np.testing.assert_allclose(T.as_matrix(), SE3(g["frame_pose"][0, k]).as_matrix(), atol=5e-6)
        np.testing.assert_allclose(cfg.get_frame_jacobian(wl["frames"][k]["name"], wl["frames"][k]["type"]), g["J_body"][0, k], atol=1e-5)
    tasks[3].set_target(g["posture_target"])
    np.testing.assert_allclose(tasks[3].compute_error(cfg), g["e_posture"][0], atol=1e-6)
    Jp = tasks[3].compute_jacobian(cfg)
    assert np.array_equal(Jp[:, :6], np.zeros((43, 6))) and np.array_equal(Jp[6:, 6:], -np.eye(37))
    # unit-cost frame task: H = J^T J, c = e^T J ... here with the example costs against the golden total
    prob = mink.build_ik(cfg, tasks, float(g["dt"]), float(g["damping"]), limits)
    np.testing.assert_allclose(prob.P, g["H"][0], atol=1e-6 * np.abs(g["H"][0]).max())
    np.testing.assert_allclose(prob.q, g["c"][0], atol=1e-5 * np.abs(g["c"][0]).max())
    assert prob.G.shape == g["G"][0].shape == (148, 43)
    np.testing.assert_array_equal(prob.G, g["G"][0])
    np.testing.assert_allclose(prob.h, g["h"][0], atol=1e-6)
    v = mink.solve_ik(cfg, tasks, float(g["dt"]), "quadprog", float(g["damping"]), limits=limits)
    assert isinstance(v, np.ndarray) and v.shape == (43,)
    np.testing.assert_allclose(v * float(g["dt"]), g["dq"][0], atol=1e-4)
    np.testing.assert_allclose(cfg.integrate(v, float(g["dt"])), g["q_next"][0], atol=2e-4)
    cfg.integrate_inplace(v, float(g["dt"]))
    np.testing.assert_allclose(cfg.q, g["q_next"][0], atol=2e-4)


def test_batched_solve_ik_matches_reference():
    wl, fm, spec, g = load_case("g1")
    cfg = mink.Configuration(fm, g["q"])
    assert cfg.batched
    tasks, limits = _g1_tasks(fm, wl)
    for k in range(3):
        tasks[k].set_target(SE3(g["frame_targets"][:, k]))
    tasks[3].set_target(g["posture_target"])
    v = mink.solve_ik(cfg, tasks, float(g["dt"]), "daqp", float(g["damping"]), limits=limits)
    assert v.is_cuda and tuple(v.shape) == (32, 43)
    np.testing.assert_allclose(_np(v) * float(g["dt"]), g["dq"], atol=1e-4)
    e = tasks[0].compute_error(cfg)
    np.testing.assert_allclose(_np(e), g["e_frame"][:, 0], atol=2e-5)
    obj = tasks[0].compute_qp_objective(cfg)
    assert tuple(obj.H.shape) == (32, 43, 43)
    con = limits[0].compute_qp_inequalities(cfg, float(g["dt"]))
    np.testing.assert_allclose(_np(con.h), g["h"][:, :74], atol=1e-6)


def test_solve_ik_converges_like_reference_test():
    """reference tests/test_solve_ik.py:95-148: UR5e `home`, +0.1 m along the site's local z, dt 5e-3,
    Configuration + Velocity(pi) limits: error strictly decreasing, converged in < 20 steps."""
    fm = load_flat("ur5e")
    cfg = mink.Configuration(fm)
    cfg.update_from_keyframe("home")
    task = mink.FrameTask("attachment_site", "site", 1.0, 1.0)
    T0 = cfg.get_transform_frame_to_world("attachment_site", "site")
    task.set_target(T0 @ SE3.from_translation(np.array([0.0, 0.0, 0.1])))
    limits = [mink.ConfigurationLimit(fm), mink.VelocityLimit(fm, {n: np.pi for n in fm.names["joint"]})]
    dt, errs = 5e-3, []
    v = mink.solve_ik(cfg, [task], dt, "quadprog", limits=limits)
    # first step of the reference (SURVEY.md 8c anchor): three joints sit on the velocity bound pi*dt
    np.testing.assert_allclose(v * dt, [0.002654728895, 0.015707963268, 0.015707963268, -0.015707963268, -0.000129333879,
                                        0.002654787088], atol=2e-5)
    for it in range(20):
        err = np.linalg.norm(task.compute_error(cfg))
        errs.append(err)
        if err < 1e-5:
            break
        v = mink.solve_ik(cfg, [task], dt, "quadprog", limits=limits)
        cfg.integrate_inplace(v, dt)
    assert errs[0] == pytest.approx(0.1, abs=1e-5)
    assert all(b < a for a, b in zip(errs, errs[1:]))
    assert len(errs) < 20 and errs[-1] < 1e-5


def test_no_tasks_and_task_at_target_give_zero_velocity():
    """reference tests/test_solve_ik.py:74-93."""
    fm = load_flat("ur5e")
    cfg = mink.Configuration(fm)
    cfg.update_from_keyframe("home")
    np.testing.assert_allclose(mink.solve_ik(cfg, [], 1e-3, "quadprog", limits=[]), np.zeros(6), atol=1e-12)
    task = mink.FrameTask("attachment_site", "site", 1.0, 1.0)
    task.set_target_from_configuration(cfg)
    np.testing.assert_allclose(mink.solve_ik(cfg, [task], 5e-3, "quadprog"), np.zeros(6), atol=1e-4)
    assert mink.build_ik(cfg, [task], 1e-3, limits=[]).G is None                     # tests/test_solve_ik.py:62-66
    assert mink.build_ik(cfg, [task], 1e-3, limits=None).G.shape == (12, 6)          # default ConfigurationLimit :68-72


def test_safety_break_and_errors():
    """reference tests/test_solve_ik.py:33-60, tests/test_configuration.py."""
    fm = load_flat("ur5e")
    q = fm.key("home").copy()
    q[1] = 100.0
    cfg = mink.Configuration(fm, q)
    task = mink.FrameTask("attachment_site", "site", 1.0, 1.0)
    task.set_target(SE3.identity())
    with pytest.raises(mink.NotWithinConfigurationLimits):
        mink.solve_ik(cfg, [task], 1e-3, "quadprog", safety_break=True)
    mink.solve_ik(cfg, [task], 1e-3, "quadprog", safety_break=False)      # warns, continues
    with pytest.raises(mink.NotWithinConfigurationLimits):
        cfg.check_limits()
    cfg.check_limits(safety_break=False)
    with pytest.raises(mink.InvalidKeyframe):
        cfg.update_from_keyframe("nope")
    with pytest.raises(mink.InvalidFrame):
        cfg.get_frame_jacobian("nope", "site")
    with pytest.raises(mink.UnsupportedFrame):
        cfg.get_transform_frame_to_world("attachment_site", "joint")
    with pytest.raises(mink.TargetNotSet):
        mink.solve_ik(mink.Configuration(fm), [mink.FrameTask("attachment_site", "site", 1.0, 1.0)], 1e-3)
    # get_transform: pose of source in dest
    cfg2 = mink.Configuration(fm, fm.key("home"))
    T = cfg2.get_transform("attachment_site", "site", "wrist_3_link", "body")
    np.testing.assert_allclose(T.translation(), [0.0, 0.1, 0.0], atol=1e-6)


def test_com_task_and_collision_limit_spot():
    wl, fm, spec, g = load_case("spot")
    cfg = mink.Configuration(fm, g["q"])
    com = mink.ComTask(cost=200.0)
    com.set_target(g["com_target"])
    np.testing.assert_allclose(_np(com.compute_error(cfg)), g["e_com"], atol=5e-6)
    np.testing.assert_allclose(_np(com.compute_jacobian(cfg)), g["J_com"], atol=5e-6)
    np.testing.assert_allclose(_np(cfg.get_com()), g["com"], atol=5e-6)
    l = wl["limits"][0]
    lim = mink.CollisionAvoidanceLimit(fm, l["pairs"], gain=l["gain"], minimum_distance_from_collisions=l["minimum_distance"],
                                       collision_detection_distance=l["detection_distance"], bound_relaxation=l["bound_relaxation"])
    con = lim.compute_qp_inequalities(cfg, float(g["dt"]))
    fin = np.isfinite(g["h"])
    assert np.array_equal(np.isfinite(_np(con.h)), fin)
    np.testing.assert_allclose(_np(con.G), g["G"], atol=2e-5)
    tasks = [mink.FrameTask(f["name"], f["type"], f["position_cost"], f["orientation_cost"]) for f in wl["frames"]]
    for k, t in enumerate(tasks):
        t.set_target(SE3(g["frame_targets"][:, k]))
    v = mink.solve_ik(cfg, tasks + [com], float(g["dt"]), "quadprog", float(g["damping"]), limits=[lim])
    np.testing.assert_allclose(_np(v) * float(g["dt"]), g["dq"], atol=5e-3)


def test_collision_limit_against_box_wall_ur5e():
    """The limit set of the reference's examples/arm_ur5e.py:29-47: wrist capsule against the floor plane and the wall BOX."""
    wl, fm, spec, g = load_case("ur5e_wall")
    cfg = mink.Configuration(fm, g["q"])
    l = wl["limits"][2]
    lim = mink.CollisionAvoidanceLimit(fm, l["pairs"], gain=l["gain"], minimum_distance_from_collisions=l["minimum_distance"],
                                       collision_detection_distance=l["detection_distance"], bound_relaxation=l["bound_relaxation"])
    assert lim.max_num_contacts == 2
    con = lim.compute_qp_inequalities(cfg, float(g["dt"]))
    Gr, hr = g["G"][:, -2:], g["h"][:, -2:]
    fin = np.isfinite(hr)
    assert np.array_equal(np.isfinite(_np(con.h)), fin)
    np.testing.assert_allclose(_np(con.h)[fin], hr[fin], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(_np(con.G), Gr, atol=2e-5)
    task = mink.FrameTask("attachment_site", "site", 1.0, 1.0, lm_damping=1.0)
    task.set_target(SE3(g["frame_targets"][:, 0]))
    post = mink.PostureTask(fm, cost=1e-2)
    post.set_target(g["posture_target"])
    limits = [mink.ConfigurationLimit(fm, gain=0.95), mink.VelocityLimit(fm, {n: np.pi for n in
              ("shoulder_pan", "shoulder_lift", "elbow", "wrist_1", "wrist_2", "wrist_3")}), lim]
    v = mink.solve_ik(cfg, [task, post], float(g["dt"]), "quadprog", float(g["damping"]), limits=limits)
    np.testing.assert_allclose(_np(v) * float(g["dt"]), g["dq"], atol=2e-5)
    # box-box pairs have no device distance: refused when the limit is compiled into a problem
    bad = mink.CollisionAvoidanceLimit(fm, [(["wall"], [fm.names["geom"].index("wall") - 1])])
    if bad.geom_id_pairs:
        with pytest.raises(mink.LimitDefinitionError):
            bad.compute_qp_inequalities(cfg, 0.01)


def test_aloha_example_collision_limit_through_the_python_api():
    """examples/arm_aloha.py:95-110 as a user would write it against this package: subtree / body geom groups from the utils
    helpers, CollisionAvoidanceLimit over the resulting 1 104 pairs, rows against the reference's golden."""
    wl, fm, spec, g = load_case("aloha_coll")
    body = fm.names["body"].index
    l_wrist = mink.get_subtree_geom_ids(fm, body("left/wrist_link"))
    r_wrist = mink.get_subtree_geom_ids(fm, body("right/wrist_link"))
    arms = mink.get_subtree_geom_ids(fm, body("left/upper_arm_link")) + mink.get_subtree_geom_ids(fm, body("right/upper_arm_link"))
    frame = mink.get_body_geom_ids(fm, body("metal_frame"))
    lim = mink.CollisionAvoidanceLimit(model=fm, geom_pairs=[(l_wrist, r_wrist), (arms, frame + ["table"])],
                                       minimum_distance_from_collisions=0.05, collision_detection_distance=0.1)
    assert lim.max_num_contacts == 1104
    cfg = mink.Configuration(fm, g["q"])
    con = lim.compute_qp_inequalities(cfg, float(g["dt"]))
    Gr, hr = g["G"][:, -1104:], g["h"][:, -1104:]
    fin = np.isfinite(hr)
    assert fin.sum() >= 10 and np.array_equal(np.isfinite(_np(con.h)), fin)
    np.testing.assert_allclose(_np(con.h)[fin], hr[fin], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(_np(con.G), Gr, atol=5e-5)


def test_relative_frame_task_matches_reference_and_world_root_identity():
    """RelativeFrameTask vs the reference golden (g1_rel), and the reference's own cross-check
    (tests/test_relative_frame_task.py:128-154): with root = world it equals minus the FrameTask."""
    wl, fm, spec, g = load_case("g1_rel")
    cfg = mink.Configuration(fm, g["q"])
    f = wl["relative_frames"][0]
    rel = mink.RelativeFrameTask(f["name"], f["type"], f["root_name"], f["root_type"], f["position_cost"], f["orientation_cost"],
                                 lm_damping=f["lm_damping"])
    with pytest.raises(mink.TargetNotSet):
        rel.compute_error(cfg)
    rel.set_target(SE3(g["frame_targets"][:, 1]))
    np.testing.assert_allclose(_np(rel.compute_error(cfg)), g["e_frame"][:, 1], atol=5e-5)
    np.testing.assert_allclose(_np(rel.compute_jacobian(cfg)), g["J_frame"][:, 1], atol=1e-4)
    pel = mink.FrameTask("pelvis", "body", 0.0, 10.0)
    pel.set_target(SE3(g["frame_targets"][:, 0]))
    post = mink.PostureTask(fm, cost=1.0)
    post.set_target(g["posture_target"])
    v = mink.solve_ik(cfg, [pel, rel, post], float(g["dt"]), "quadprog", float(g["damping"]), limits=[mink.ConfigurationLimit(fm)])
    np.testing.assert_allclose(_np(v) * float(g["dt"]), g["dq"], atol=1e-4 * max(1.0, np.abs(g["dq"]).max()))
    # root = world  ==>  error and Jacobian are minus the FrameTask's
    one = mink.Configuration(fm, g["q"][3])
    T = SE3(g["frame_targets"][3, 0])
    a = mink.RelativeFrameTask("pelvis", "body", "world", "body", 1.0, 1.0)
    b = mink.FrameTask("pelvis", "body", 1.0, 1.0)
    a.set_target(T); b.set_target(T)
    np.testing.assert_allclose(a.compute_error(one), -b.compute_error(one), atol=2e-6)
    np.testing.assert_allclose(a.compute_jacobian(one), -b.compute_jacobian(one), atol=2e-5)
    a.set_target_from_configuration(one)
    np.testing.assert_allclose(a.compute_error(one), np.zeros(6), atol=2e-6)


def test_converge_ik_matches_the_explicit_loop():
    """converge_ik == the examples' loop written with solve_ik + integrate_inplace (examples/arm_iiwa.py:63-70)."""
    wl, fm, spec, g = load_case("ur5e")

    def make(b):
        cfg = mink.Configuration(fm, g["q"][b])
        ee = mink.FrameTask("attachment_site", "site", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)
        ee.set_target(SE3(wxyz_xyz=g["frame_targets"][b, 0]))
        post = mink.PostureTask(fm, cost=1e-2)
        post.set_target(g["posture_target"])
        return cfg, [ee, post], ee, [mink.ConfigurationLimit(fm, gain=0.95)]

    max_iters, pos, ori = int(g["conv_params"][0]), float(g["conv_params"][1]), float(g["conv_params"][2])
    dt, damping = float(g["dt"]), float(g["damping"])
    for b in range(4):
        cfg, tasks, ee, limits = make(b)
        it, ok = mink.converge_ik(cfg, tasks, dt, "quadprog", damping, limits=limits, max_iters=max_iters, pos_threshold=pos,
                                  ori_threshold=ori)
        assert it == int(g["conv_iters"][b]) and ok == bool(g["conv_ok"][b])
        np.testing.assert_allclose(cfg.q, g["conv_q"][b], atol=1e-3)
        cfg2, tasks2, ee2, limits2 = make(b)
        steps = 0
        for i in range(max_iters):
            v = mink.solve_ik(cfg2, tasks2, dt, "quadprog", damping, limits=limits2)
            cfg2.integrate_inplace(v, dt)
            steps = i + 1
            err = ee2.compute_error(cfg2)
            if np.linalg.norm(err[:3]) <= pos and np.linalg.norm(err[3:]) <= ori:
                break
        assert steps == it
        np.testing.assert_allclose(cfg.q, cfg2.q, atol=2e-5)


def test_damping_task_objective_matches_reference_test():
    """reference tests/test_damping_task.py:21-26: DampingTask(cost=1) on UR5e gives H = I, c = 0 -- single configuration
    (fp64 kernels) and a batch (fp32 kernels)."""
    fm = load_flat("ur5e")
    task = mink.DampingTask(fm, cost=1.0)
    cfg = mink.Configuration(fm)
    H, c = task.compute_qp_objective(cfg)
    np.testing.assert_allclose(H, np.eye(cfg.nv))
    np.testing.assert_allclose(c, np.zeros(cfg.nv))
    q = np.tile(fm.key("home"), (7, 1)) + np.linspace(0, 0.3, 7)[:, None]
    cfgb = mink.Configuration(fm, q)
    Hb, cb = task.compute_qp_objective(cfgb)
    np.testing.assert_allclose(_np(Hb), np.tile(np.eye(cfg.nv), (7, 1, 1)))
    np.testing.assert_allclose(_np(cb), 0.0)
    # the error is q0 - q whatever the configuration, and the gain is 0: it never enters c
    np.testing.assert_allclose(task.compute_error(mink.Configuration(fm, q[3])), fm.qpos0 - q[3], atol=1e-12)


def test_damping_task_in_solve_ik_matches_reference_golden():
    """FrameTask + DampingTask + Configuration/Velocity limits through solve_ik against the golden made by running the
    reference's DampingTask (oracle/gen_golden.py, case ur5e_damp)."""
    wl, fm, spec, g = load_case("ur5e_damp")
    f = wl["frames"][0]
    frame = mink.FrameTask(f["name"], f["type"], f["position_cost"], f["orientation_cost"], lm_damping=f["lm_damping"])
    damp = mink.DampingTask(fm, cost=wl["damping_task"]["cost"])
    limits = [mink.ConfigurationLimit(fm), mink.VelocityLimit(fm, {n: 2 * np.pi for n, t in zip(fm.names["joint"], fm.node_type) if t >= 2})]
    dt, damping = float(g["dt"]), float(g["damping"])
    for b in range(4):
        cfg = mink.Configuration(fm, g["q"][b])
        frame.set_target(SE3(g["frame_targets"][b, 0]))
        v = mink.solve_ik(cfg, [frame, damp], dt, "quadprog", damping, limits=limits)
        np.testing.assert_allclose(v * dt, g["dq"][b], atol=2e-7)
        prob = mink.build_ik(cfg, [frame, damp], dt, damping, limits)
        np.testing.assert_allclose(prob.P, g["H"][b], atol=1e-10 * np.abs(g["H"][b]).max())
        np.testing.assert_allclose(prob.q, g["c"][b], atol=1e-10 * max(1.0, np.abs(g["c"][b]).max()))
    cfgb = mink.Configuration(fm, g["q"])
    frame.set_target(SE3(g["frame_targets"][:, 0]))
    vb = mink.solve_ik(cfgb, [frame, damp], dt, "quadprog", damping, limits=limits)
    np.testing.assert_allclose(_np(vb) * dt, g["dq"], atol=1e-4)
