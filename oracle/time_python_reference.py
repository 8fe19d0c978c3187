"""ORACLE / TEST INFRASTRUCTURE -- time the UNMODIFIED reference (Python, /root/reference/mink) on the numpy shims of
mujoco / qpsolvers (SURVEY.md 8d "CPU baseline timing", rows (i) and (ii)): solve_ik + integrate per instance, one process
on one core and one worker per core.  Runs only where /root/reference exists (the build container, not the GPU box); the
result is committed as profiles/r2_python_reference_cpu.json and quoted by bench.py as a labelled, separately-measured row.

    python oracle/time_python_reference.py [workload ...]
"""

import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("MINK_REFERENCE", "/root/reference")
sys.path[:0] = [HERE, REF, REPO]


def _run(args):
    name, n, seed = args
    import numpy as np
    import mujoco  # noqa: F401  (shim)
    import mink

    from gen_golden import build_reference_problem
    from mink_b200.flatten import flatten
    from mink_b200.workloads import WORKLOADS, make_inputs

    wl = WORKLOADS[name]
    scene = os.path.join(REPO, wl["scene"][1:]) if wl["scene"].startswith("@") else os.path.join(REF, "examples", wl["scene"])
    model = mujoco.MjModel.from_xml_path(scene)
    fm = flatten(model)
    tasks, frame_tasks, posture, com, limits = build_reference_problem(model, wl)
    cfg = mink.Configuration(model)

    def fk(qb):
        poses = np.zeros((qb.shape[0], len(frame_tasks), 7)); coms = np.zeros((qb.shape[0], 3))
        for b in range(qb.shape[0]):
            cfg.update(qb[b])
            for k, f in enumerate(wl["frames"]):
                poses[b, k] = cfg.get_transform_frame_to_world(f["name"], f["type"]).wxyz_xyz
            coms[b] = cfg.data.subtree_com[1]
        return poses, coms

    inp = make_inputs(fm, wl, n, fk, seed=seed)
    if posture is not None and not isinstance(posture, mink.DampingTask):
        posture.set_target(inp["posture_target"])
    t0 = time.perf_counter()
    for b in range(n):
        cfg.update(inp["q"][b])
        for k, t in enumerate(frame_tasks):
            t.set_target(mink.SE3(wxyz_xyz=inp["frame_targets"][b, k]))
        if com is not None:
            com.set_target(inp["com_target"][b])
        v = mink.solve_ik(cfg, tasks, wl["dt"], "quadprog", wl["damping"], limits=limits)
        cfg.integrate_inplace(v, wl["dt"])
    return n / (time.perf_counter() - t0)


def main():
    names = sys.argv[1:] or ["g1", "ur5e_dls", "shadow", "spot"]
    ncpu = len(os.sched_getaffinity(0))
    out = {"where": "build container (no GPU); numpy shims of mujoco/qpsolvers, not the native wheels", "cores": ncpu, "rows": {}}
    for name in names:
        n1 = 60
        try:
            one = _run((name, n1, 0))
        except Exception as exc:   # a shim limitation must not lose the other rows
            out["rows"][name] = {"error": f"{type(exc).__name__}: {exc}"}
            print(name, out["rows"][name], flush=True)
            continue
        with mp.Pool(ncpu) as pool:
            t0 = time.perf_counter()
            rates = pool.map(_run, [(name, n1, 100 + i) for i in range(ncpu)])
            wall = time.perf_counter() - t0
        out["rows"][name] = {"steps_per_s_one_core": one, "steps_per_s_all_cores_sum_of_workers": float(sum(rates)),
                             "instances_per_worker": n1, "workers": ncpu, "pool_wall_s": wall}
        print(name, out["rows"][name], flush=True)
    path = os.path.join(REPO, "profiles", "r2_python_reference_cpu.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
