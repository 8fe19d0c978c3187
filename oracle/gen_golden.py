"""ORACLE / TEST INFRASTRUCTURE -- generate tests/golden/ by running the UNMODIFIED reference.

Imports /root/reference/mink on top of the oracle shims (oracle/mujoco, oracle/qpsolvers) and,
for every BASELINE workload, records per instance: inputs (q, targets), FK poses, task errors and
Jacobians (Task.compute_error / compute_jacobian), the QP (build_ik -> P, q, G, h), the solution
dq = solve_ik(...) * dt, the integrated configuration, and a short solve+integrate rollout.
Also writes the flattened model blobs (mink_b200/models/*.bikm + *.json) so that the GPU box,
which has no /root/reference, can rebuild every problem from numbers alone.

Run here (needs /root/reference):  python oracle/gen_golden.py
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("MINK_REFERENCE", "/root/reference")
sys.path[:0] = [HERE, REF, REPO]

import numpy as np  # noqa: E402
import mujoco  # noqa: E402  (oracle shim)
import mink  # noqa: E402  (the unmodified reference)

from mink_b200.flatten import flatten  # noqa: E402
from mink_b200.workloads import WORKLOADS, make_inputs, resolve_geom_groups  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
GOLDEN_B = {"leap": 8, "aloha_coll": 6, "iiwa": 8, "h1": 8, "go1": 8, "stretch": 8, "tidybot": 8, "aloha": 8, "ur5e_wall": 24, "ur5e_damp": 12, "ur5e": 8, "ur5e_dls": 16, "g1": 32, "shadow": 24, "spot": 24, "g1_rel": 16, "edge": 48, "g1_full": 16, "g1_hands": 24}
ROLLOUT_T, ROLLOUT_B = 8, 4
CONV_B, CONV_MAX_ITERS, CONV_POS, CONV_ORI = 8, 20, 2e-3, 2e-3


def build_reference_problem(model, wl):
    tasks, frame_tasks = [], []
    for f in wl["frames"]:
        t = mink.FrameTask(f["name"], f["type"], f["position_cost"], f["orientation_cost"],
                           gain=f.get("gain", 1.0), lm_damping=f["lm_damping"])
        frame_tasks.append(t)
    for f in wl.get("relative_frames", []):
        t = mink.RelativeFrameTask(f["name"], f["type"], f["root_name"], f["root_type"], f["position_cost"],
                                   f["orientation_cost"], lm_damping=f["lm_damping"])
        frame_tasks.append(t)
    tasks.extend(frame_tasks)
    posture = com = None
    if wl["posture"] is not None:
        posture = mink.PostureTask(model, cost=wl["posture"]["cost"])
        tasks.append(posture)
    if wl.get("damping_task") is not None:   # takes the posture slot of the record (its target is qpos0)
        assert posture is None
        posture = mink.DampingTask(model, cost=wl["damping_task"]["cost"])
        tasks.append(posture)
    if wl["com"] is not None:
        com = mink.ComTask(cost=wl["com"]["cost"])
        tasks.append(com)
    limits = []
    for l in wl["limits"]:
        if l["kind"] == "configuration":
            limits.append(mink.ConfigurationLimit(model, gain=l["gain"]))
        elif l["kind"] == "velocity":
            vel = {model.joint_names[j]: (np.full(3, l["vmax"]) if model.jnt_type[j] == 1 else l["vmax"])
                   for j in range(model.njnt) if model.jnt_type[j] in (1, 2, 3)}
            limits.append(mink.VelocityLimit(model, vel))
        elif l["kind"] == "collision":
            limits.append(mink.CollisionAvoidanceLimit(
                model, resolve_geom_groups(model, l["pairs"]), gain=l["gain"],
                minimum_distance_from_collisions=l["minimum_distance"],
                collision_detection_distance=l["detection_distance"],
                bound_relaxation=l["bound_relaxation"]))
    return tasks, frame_tasks, posture, com, limits


def main():
    MODELS = os.path.join(REPO, "mink_b200", "models")
    os.makedirs(MODELS, exist_ok=True)
    done_models = set()
    only = set(sys.argv[1:])   # `python oracle/gen_golden.py g1_hands` regenerates just that case
    for name, wl in WORKLOADS.items():
        if only and name not in only:
            continue
        scene = os.path.join(REPO, wl["scene"][1:]) if wl["scene"].startswith("@") else os.path.join(REF, "examples", wl["scene"])
        model = mujoco.MjModel.from_xml_path(scene)
        fm = flatten(model)
        if wl["robot"] not in done_models and (not only or not os.path.exists(os.path.join(MODELS, wl["robot"] + ".bikm"))):
            with open(os.path.join(MODELS, wl["robot"] + ".bikm"), "wb") as f:
                f.write(fm.to_blob())
            with open(os.path.join(MODELS, wl["robot"] + ".json"), "w") as f:
                f.write(fm.to_meta_json())
            done_models.add(wl["robot"])
        tasks, frame_tasks, posture, com, limits = build_reference_problem(model, wl)
        cfg = mink.Configuration(model)
        nv, nq, F = model.nv, model.nq, len(frame_tasks)

        rel = wl.get("relative_frames", [])
        nabs = len(wl["frames"])

        def fk(qb):
            poses = np.zeros((qb.shape[0], F, 7))
            coms = np.zeros((qb.shape[0], 3))
            for b in range(qb.shape[0]):
                cfg.update(qb[b])
                for k, f in enumerate(wl["frames"]):
                    poses[b, k] = cfg.get_transform_frame_to_world(f["name"], f["type"]).wxyz_xyz
                for k, f in enumerate(rel):   # relative tasks are targeted in the root frame
                    poses[b, nabs + k] = cfg.get_transform(f["name"], f["type"], f["root_name"], f["root_type"]).wxyz_xyz
                coms[b] = cfg.data.subtree_com[1]
            return poses, coms

        B = GOLDEN_B[name]
        inp = make_inputs(fm, wl, B, fk, seed=0)
        q = inp["q"]
        # mj_kinematics normalises free-joint quaternions in place; inputs are already unit.
        dt, damping = wl["dt"], wl["damping"]
        rec = dict(q=q, frame_targets=inp["frame_targets"], posture_target=inp["posture_target"],
                   dt=np.array(dt), damping=np.array(damping))
        if com is not None:
            rec["com_target"] = inp["com_target"]
        poses = np.zeros((B, F, 7)); comp = np.zeros((B, 3))
        eF = np.zeros((B, F, 6)); JF = np.zeros((B, F, 6, nv)); JB = np.zeros((B, F, 6, nv))
        eP = np.zeros((B, nv)); eC = np.zeros((B, 3)); JC = np.zeros((B, 3, nv))
        H = np.zeros((B, nv, nv)); c = np.zeros((B, nv)); dq = np.zeros((B, nv)); qn = np.zeros((B, nq))
        lo = np.full((B, nv), -np.inf); hi = np.full((B, nv), np.inf)
        Gs, hs, nact = [], [], np.zeros(B, dtype=np.int32)
        for b in range(B):
            cfg.update(q[b])
            for k, t in enumerate(frame_tasks):
                t.set_target(mink.SE3(wxyz_xyz=inp["frame_targets"][b, k]))
                eF[b, k] = t.compute_error(cfg)
                JF[b, k] = t.compute_jacobian(cfg)
                if k < nabs:
                    f = wl["frames"][k]
                    poses[b, k] = cfg.get_transform_frame_to_world(f["name"], f["type"]).wxyz_xyz
                    JB[b, k] = cfg.get_frame_jacobian(f["name"], f["type"])
                else:
                    f = rel[k - nabs]
                    poses[b, k] = cfg.get_transform(f["name"], f["type"], f["root_name"], f["root_type"]).wxyz_xyz
            comp[b] = cfg.data.subtree_com[1]
            if posture is not None:
                if not isinstance(posture, mink.DampingTask):
                    posture.set_target(inp["posture_target"])
                eP[b] = posture.compute_error(cfg)
            if com is not None:
                com.set_target(inp["com_target"][b])
                eC[b] = com.compute_error(cfg)
                JC[b] = com.compute_jacobian(cfg)
            prob = mink.build_ik(cfg, tasks, dt, damping, limits)
            H[b], c[b] = prob.P, prob.q
            if prob.G is not None:
                Gs.append(prob.G); hs.append(prob.h)
                for r in range(prob.G.shape[0]):
                    nz = np.nonzero(prob.G[r])[0]
                    if len(nz) == 1 and abs(prob.G[r, nz[0]]) == 1.0:   # box row +-e_i
                        i = nz[0]
                        if prob.G[r, i] > 0:
                            hi[b, i] = min(hi[b, i], prob.h[r])
                        else:
                            lo[b, i] = max(lo[b, i], -prob.h[r])
            import qpsolvers
            sol = qpsolvers.solve_problem(prob, solver="quadprog")
            assert sol.x is not None
            nact[b] = len(sol.active)
            v = mink.solve_ik(cfg, tasks, dt, "quadprog", damping, limits=limits)
            dq[b] = v * dt
            assert np.allclose(dq[b], sol.x, atol=1e-12)
            qn[b] = cfg.integrate(v, dt)
        rec.update(frame_pose=poses, com=comp, e_frame=eF, J_frame=JF, J_body=JB, e_posture=eP,
                   e_com=eC, J_com=JC, H=H, c=c, dq=dq, q_next=qn, box_lo=lo, box_hi=hi, n_active=nact)
        if Gs:
            rec["G"] = np.stack(Gs); rec["h"] = np.stack(hs)
        # Rollout: ROLLOUT_T solve+integrate steps with targets held (reference examples' loop,
        # e.g. examples/humanoid_g1.py:81-94).
        RB = min(ROLLOUT_B, B)
        traj = np.zeros((ROLLOUT_T + 1, RB, nq)); traj[0] = q[:RB]
        for b in range(RB):
            cfg.update(q[b])
            for k, t in enumerate(frame_tasks):
                t.set_target(mink.SE3(wxyz_xyz=inp["frame_targets"][b, k]))
            if com is not None:
                com.set_target(inp["com_target"][b])
            for s in range(ROLLOUT_T):
                v = mink.solve_ik(cfg, tasks, dt, "quadprog", damping, limits=limits)
                cfg.integrate_inplace(v, dt)
                traj[s + 1, b] = cfg.q
        rec["rollout_q"] = traj
        # Converge-until-threshold loop of the examples (examples/arm_iiwa.py:63-70, arm_ur5e_actuators.py:88-97): solve_ik +
        # integrate, then every frame task's error at the new configuration against the thresholds.
        CB = min(CONV_B, B)
        conv_q = np.zeros((CB, nq)); conv_it = np.zeros(CB, dtype=np.int32); conv_ok = np.zeros(CB, dtype=np.int32)
        for b in range(CB):
            cfg.update(q[b])
            for k, t in enumerate(frame_tasks):
                t.set_target(mink.SE3(wxyz_xyz=inp["frame_targets"][b, k]))
            if com is not None:
                com.set_target(inp["com_target"][b])
            for i in range(CONV_MAX_ITERS):
                v = mink.solve_ik(cfg, tasks, dt, "quadprog", damping, limits=limits)
                cfg.integrate_inplace(v, dt)
                ok = True
                for t in frame_tasks:
                    err = t.compute_error(cfg)
                    ok = ok and bool(np.linalg.norm(err[:3]) <= CONV_POS) and bool(np.linalg.norm(err[3:]) <= CONV_ORI)
                conv_it[b] = i + 1
                if ok:
                    conv_ok[b] = 1
                    break
            conv_q[b] = cfg.q
        rec.update(conv_q=conv_q, conv_iters=conv_it, conv_ok=conv_ok, conv_params=np.array([CONV_MAX_ITERS, CONV_POS, CONV_ORI]))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(f"{name}: B={B} nv={nv} F={F} rows={0 if not Gs else Gs[0].shape[0]} "
              f"active mean={nact.mean():.1f} max={nact.max()} |dq|max={np.abs(dq).max():.4f} conv iters={conv_it.tolist()} ok={conv_ok.tolist()} "
              f"cond(H) median={np.median([np.linalg.cond(h) for h in H]):.2e}")


if __name__ == "__main__":
    main()
