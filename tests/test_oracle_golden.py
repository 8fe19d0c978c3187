"""The C oracle (oracle/ik_oracle.c) against the golden vectors produced by the unmodified reference.

Golden = /root/reference/mink running on the numpy shims (oracle/gen_golden.py).  fp64 vs fp64, so
tolerances are tight; they are not bit-exact because the two restatements order operations differently.
"""

import numpy as np
import pytest

from oracle.ikoracle import Oracle
from tests.helpers import load_case, quat_align, task_frames

CASES = ["ur5e", "ur5e_dls", "g1", "shadow", "spot", "g1_rel", "edge", "g1_full", "g1_hands", "ur5e_damp", "ur5e_wall",
         "iiwa", "h1", "go1", "stretch", "tidybot", "aloha", "aloha_coll", "leap"]   # the last eight: the reference's other example robots


def _oracle(name):
    wl, fm, spec, g = load_case(name)
    return wl, fm, spec, g, Oracle(fm.to_blob(), spec, fm.nq, fm.nv)


@pytest.mark.parametrize("name", CASES)
def test_fk_and_body_jacobian(name):
    wl, fm, spec, g, orc = _oracle(name)
    frames = task_frames(wl, fm)
    poses, com = orc.fk(g["q"], frames)
    ref = g["frame_pose"][:, :len(frames)]
    np.testing.assert_allclose(poses[..., 4:], ref[..., 4:], atol=1e-12)
    np.testing.assert_allclose(quat_align(poses[..., :4], ref[..., :4]), ref[..., :4], atol=1e-12)
    if fm.ncom:
        np.testing.assert_allclose(com, g["com"], atol=1e-12)
    Jb = orc.frame_jacobian(g["q"], frames)
    np.testing.assert_allclose(Jb, g["J_body"][:, :len(frames)], atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_task_errors_and_jacobians(name):
    wl, fm, spec, g, orc = _oracle(name)
    J, e, ep = orc.fk_jac(g["q"], g["frame_targets"], g["posture_target"], g.get("com_target"))
    F = spec.nframe
    np.testing.assert_allclose(e[:, :6 * F].reshape(-1, F, 6), g["e_frame"], atol=1e-11)
    np.testing.assert_allclose(J[:, :6 * F].reshape(-1, F, 6, fm.nv), g["J_frame"], atol=1e-10)
    if spec.nposture:
        np.testing.assert_allclose(ep[:, 0], g["e_posture"], atol=1e-12)
    if spec.ncom:
        np.testing.assert_allclose(e[:, 6 * F:], g["e_com"], atol=1e-12)
        np.testing.assert_allclose(J[:, 6 * F:], g["J_com"], atol=1e-12)
    H, c = orc.objective(J, e, ep, float(g["damping"]))
    scale = np.abs(g["H"]).max()
    np.testing.assert_allclose(H, g["H"], atol=1e-11 * scale)
    np.testing.assert_allclose(c, g["c"], atol=1e-11 * max(1.0, np.abs(g["c"]).max()))


@pytest.mark.parametrize("name", CASES)
def test_limits(name):
    wl, fm, spec, g, orc = _oracle(name)
    lo, hi = orc.box(g["q"], float(g["dt"]))
    np.testing.assert_allclose(lo, g["box_lo"], atol=1e-13)
    np.testing.assert_allclose(hi, g["box_hi"], atol=1e-13)
    if spec.npairs:
        G, h = orc.collision(g["q"], float(g["dt"]))
        Gr, hr = g["G"][:, -spec.npairs:], g["h"][:, -spec.npairs:]   # collision rows are stacked last
        fin = np.isfinite(hr)
        assert np.array_equal(np.isfinite(h), fin)
        np.testing.assert_allclose(h[fin], hr[fin], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(G, Gr, atol=1e-10)


@pytest.mark.parametrize("name", CASES)
def test_solve_and_integrate(name):
    wl, fm, spec, g, orc = _oracle(name)
    dq, qn, st, nact = orc.step(g["q"], g["frame_targets"], g["posture_target"], g.get("com_target"),
                                dt=float(g["dt"]), damping=float(g["damping"]), nsteps=1, integrate=True)
    assert not st.any()
    # exact-QP optimum: both sides are fp64 active-set solves
    tol = {"spot": 1e-6, "g1_rel": 1e-8}.get(name, 1e-9)   # spot: cond(H) ~ 4e7 (posture-free, damping 1e-3)
    np.testing.assert_allclose(dq, g["dq"], atol=tol)
    np.testing.assert_array_equal(nact, g["n_active"])
    np.testing.assert_allclose(qn, g["q_next"], atol=tol)


@pytest.mark.parametrize("name", CASES)
def test_rollout(name):
    wl, fm, spec, g, orc = _oracle(name)
    traj = g["rollout_q"]
    T, RB = traj.shape[0] - 1, traj.shape[1]
    ct = g["com_target"][:RB] if "com_target" in g else None
    _, q, st, _ = orc.step(traj[0], g["frame_targets"][:RB], g["posture_target"], ct, dt=float(g["dt"]),
                           damping=float(g["damping"]), nsteps=T, integrate=True)
    assert not st.any()
    tol = {"spot": 1e-5, "g1_rel": 1e-6}.get(name, 1e-8)
    np.testing.assert_allclose(q, traj[-1], atol=tol)


@pytest.mark.parametrize("name", CASES)
def test_converge_loop(name):
    """The examples' solve+integrate-until-threshold loop (examples/arm_iiwa.py:63-70): iterations and final q per instance."""
    wl, fm, spec, g, orc = _oracle(name)
    CB = g["conv_q"].shape[0]
    max_iters, pos, ori = int(g["conv_params"][0]), float(g["conv_params"][1]), float(g["conv_params"][2])
    ct = g["com_target"][:CB] if "com_target" in g else None
    q, it, ok, st = orc.converge(g["q"][:CB], g["frame_targets"][:CB], g["posture_target"], ct, dt=float(g["dt"]), damping=float(g["damping"]),
                                 max_iters=max_iters, pos_threshold=pos, ori_threshold=ori)
    assert not st.any()
    np.testing.assert_array_equal(it, g["conv_iters"])
    np.testing.assert_array_equal(ok, g["conv_ok"])
    tol = {"spot": 1e-4, "g1_rel": 1e-5, "edge": 1e-6}.get(name, 1e-7)
    np.testing.assert_allclose(q, g["conv_q"], atol=tol)
