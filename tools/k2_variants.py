"""A/B harness: build libbik variants (compile-time knobs) and time K1 / K2 alone for each, with
launch knobs from the environment.  Usage on the GPU box:  python tools/k2_variants.py"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "base": [],
    "k1old": ["-DBIK_K1_NORMALIZE_ALL", "-DBIK_K1_MINBLOCKS=4"],
}

CHILD = r'''
import os, sys, json, statistics
sys.path.insert(0, os.environ["BIK_REPO"])
import numpy as np, torch
from mink_b200 import _lib
_lib._LIB_PATH = os.environ["BIK_LIB"]
from mink_b200._abi import spec_from_workload
from mink_b200.engine import DeviceModel, Problem
from mink_b200.workloads import WORKLOADS, make_inputs
from tests.helpers import load_flat, task_frames
WL = os.environ.get("BIK_WL", "g1")
wl = WORKLOADS[WL]; fm = load_flat(wl["robot"]); spec = spec_from_workload(fm, wl)
model = DeviceModel(fm, 0); prob = Problem(model, spec); frames = task_frames(wl, fm)
def fk(qq):
    p, c = model.fk(qq, frames); return p.cpu().numpy().astype(np.float64), None
B = int(os.environ.get("BIK_B", str(wl["batch"])))
inp = make_inputs(fm, wl, B, fk, seed=1000)
f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0")
q0, ft, pt = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"])
flush = torch.empty(64 * 1024 * 1024, device="cuda:0")
J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, None, dt=wl["dt"])
k1, k2 = [], []
for s in range(8):
    flush.fill_(1.0)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda._sleep(2000000); a.record(); J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, None, dt=wl["dt"]); b.record()
    dq, st = prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"]); c.record()
    torch.cuda.synchronize()
    k1.append(a.elapsed_time(b)); k2.append(b.elapsed_time(c))
dq2, st2, it = prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"], return_iters=True)
print(json.dumps({"wl": WL, "B": B, "k1_ms": statistics.median(k1[2:]), "k2_ms": statistics.median(k2[2:]), "status": int(st.max()), "iters_mean": float(it.float().mean()), "iters_max": int(it.max())}))
'''


def main():
    out = os.path.join(REPO, "mink_b200", "lib", "variants")
    os.makedirs(out, exist_ok=True)
    build_only = "--build" in sys.argv
    for name, flags in VARIANTS.items():
        so = os.path.join(out, f"libbik_{name}.so")
        if not os.path.exists(so) or build_only:
            cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                   "-shared"] + flags + ["-o", so, os.path.join(REPO, "mink_b200", "csrc", "bik.cu")]
            subprocess.check_call(cmd)
    if build_only:
        return
    ENVS = {"base": ({}, {"BIK_WL": "shadow"}, {"BIK_WL": "ur5e_dls"}, {"BIK_WL": "g1_rel"}), "k1old": ({}, {"BIK_WL": "shadow"})}
    for name in VARIANTS:
        for env in ENVS.get(name, ({},)):
            e = dict(os.environ, BIK_REPO=REPO, BIK_LIB=os.path.join(out, f"libbik_{name}.so"), **env)
            r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(name, env, line, flush=True)

if __name__ == "__main__":
    main()
