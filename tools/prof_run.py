"""Profiling driver: a few solve_ik steps of a BASELINE workload, nothing else (for ncu)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mink_b200._abi import spec_from_workload
from mink_b200.engine import DeviceModel, Problem
from mink_b200.workloads import WORKLOADS, make_inputs
from mink_b200.workloads import load_flat, task_frames

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="g1")
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
wl = WORKLOADS[args.workload]
fm = load_flat(wl["robot"])
spec = spec_from_workload(fm, wl)
model = DeviceModel(fm, 0)
prob = Problem(model, spec)
frames = task_frames(wl, fm)


def fk(qq):
    p, c = model.fk(qq, frames, want_com=spec.ncom > 0)
    return p.cpu().numpy().astype(np.float64), (None if c is None else c.cpu().numpy().astype(np.float64))


inp = make_inputs(fm, wl, args.batch, fk, seed=1000)
f32 = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0")
q0, ft, pt, ct = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"]), f32(inp.get("com_target"))
q = q0.clone()
for s in range(args.warmup + args.steps):
    q.copy_(q0)
    prob.step(q, ft, pt, ct, dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=True)
torch.cuda.synchronize()
J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=wl["dt"])
dq, st, it = prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"], return_iters=True)
it = it.cpu().numpy()
print("iters mean %.2f max %d hist %s" % (it.mean(), it.max(), np.bincount(it)[:20].tolist()))
print("done")
