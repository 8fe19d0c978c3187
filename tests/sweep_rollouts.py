"""TEST INFRASTRUCTURE (not collected by pytest): robustness sweep of the small-group K2 on the host emulation.

    python tests/sweep_rollouts.py g1 32768 12 5 [sigma]      # workload, instances, steps, seed, target distance (rad)

Rolls a batch out with the kernel headers compiled for the host (tests/host_emu, the rollout state of bik_step carried
between steps) and checks every step of every instance against the exact fp64 oracle (oracle/ik_oracle.c: Goldfarb-Idnani,
the algorithm quadprog implements).  Prints per-step iteration statistics, the number of flagged / mismatching instances
and the worst |dq - dq_oracle|.  DESIGN.md section 3 (K2) quotes the figures of the sweeps run at the end of round 1.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mink_b200._abi import spec_from_workload  # noqa: E402
from mink_b200.workloads import WORKLOADS, make_inputs  # noqa: E402
from oracle.ikoracle import Oracle  # noqa: E402
from tests.emu_lib import Emu  # noqa: E402
from tests.helpers import load_flat, task_frames  # noqa: E402


def main():
    name, B, T, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    emu = Emu(fm.to_blob(), spec, fm.nq, fm.nv)
    frames = task_frames(wl, fm)
    orc = Oracle(fm.to_blob(), spec, fm.nq, fm.nv)
    inp = make_inputs(fm, wl, B, lambda qq: orc.fk(qq, frames), seed=seed, sigma=sigma)
    q = inp["q"].astype(np.float32)
    warm = np.zeros((B, emu.header()["nu"]), np.int8)
    dq = np.zeros((B, fm.nv), np.float32)
    t0, worst, bad_total, itmax = time.time(), 0.0, 0, 0
    for step in range(T):
        J, e, ep, Gc, hc = emu.fk_jac(q, inp["frame_targets"], inp["posture_target"], inp.get("com_target"), dt=wl["dt"])
        dq, st, it = emu.solve_warm(q, J, e, ep, wl["dt"], wl["damping"], dq, warm)
        dq_ref, _, st_ref, nact = orc.step(q.astype(np.float64), inp["frame_targets"], inp["posture_target"], inp.get("com_target"),
                                           dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=False)
        err = np.abs(dq - dq_ref).max(axis=1)
        bad = int(((st != 0) | (err > 1e-4)).sum())
        bad_total += bad; itmax = max(itmax, int(it.max())); worst = max(worst, float(err.max()))
        print(f"{name} step {step}: iterations mean {it.mean():.3f} max {it.max()}  flagged {int((st != 0).sum())} (oracle {int((st_ref != 0).sum())})"
              f"  max|dq-dq_oracle| {err.max():.2e}  active bounds {nact.mean():.1f}", flush=True)
        q = emu.integrate(q, dq)
    print(f"SUMMARY {name} B={B} T={T} seed={seed} sigma={sigma}: flagged or > 1e-4 off: {bad_total}, worst {worst:.2e}, max iterations {itmax}, {time.time() - t0:.0f} s")
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
