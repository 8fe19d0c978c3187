"""A/B harness (GPU box): time the pieces of one solve_ik step for library variants and environment knobs.

  python tools/ab.py [--build-only] [spec ...]

A spec is  name[:flag,flag...][@ENV=V,ENV=V...][#workload[,B]]  e.g.  base  k1m6:-DBIK_K1_MINBLOCKS=6  base@BIK_K2_GROUP=4#shadow
Variants are built here (CPU box, `--build-only`) into mink_b200/lib/variants/<name>/libbik.so and travel with gpurun.
Per spec it prints one JSON line: step (event-timed, L2 flushed, as bench.py), K1 / K2 of the step in situ (BIK_STEP_PHASE),
K1 / K2 through the dense API (bik_fk_jac / bik_solve), iterations.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CHILD = r'''
import os, sys, json, statistics
sys.path.insert(0, os.environ["BIK_REPO"])
import numpy as np, torch
from mink_b200 import _lib
_lib._LIB_PATH = os.environ["BIK_LIB"]
from mink_b200._abi import spec_from_workload
from mink_b200.engine import DeviceModel, Problem
from mink_b200.workloads import WORKLOADS, make_inputs, load_flat, task_frames
WL = os.environ.get("BIK_WL", "g1")
wl = WORKLOADS[WL]; fm = load_flat(wl["robot"]); spec = spec_from_workload(fm, wl)
model = DeviceModel(fm, 0); prob = Problem(model, spec); frames = task_frames(wl, fm)
def fk(qq):
    p, c = model.fk(qq, frames, want_com=spec.ncom > 0); return p.cpu().numpy().astype(np.float64), (None if c is None else c.cpu().numpy().astype(np.float64))
B = int(os.environ.get("BIK_B", str(wl["batch"])))
inp = make_inputs(fm, wl, B, fk, seed=1000)
f32 = lambda a: None if a is None else torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0")
q0, ft, pt, ct = f32(inp["q"]), f32(inp["frame_targets"]), f32(inp["posture_target"]), f32(inp.get("com_target"))
flush = torch.empty(64 * 1024 * 1024, device="cuda:0")
q = q0.clone(); dq = torch.empty((B, fm.nv), device="cuda:0"); st = torch.empty(B, device="cuda:0", dtype=torch.int32)
def timed(fn, n=10, pre=None):
    out = []
    for s in range(n + 2):
        if pre: pre()
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(1000000); a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out[2:])
step = lambda: prob.step(q, ft, pt, ct, dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=True, dq=dq, status=st)
res = {"wl": WL, "B": B}
res["step_ms"] = timed(step, pre=lambda: q.copy_(q0))
res["status"] = int(st.max())
os.environ["BIK_STEP_PHASE"] = "1"
res["k1_step_ms"] = timed(step, pre=lambda: q.copy_(q0))
os.environ["BIK_STEP_PHASE"] = "2"
res["k2_step_ms"] = timed(step, pre=lambda: q.copy_(q0))
os.environ["BIK_STEP_PHASE"] = "0"
if True:
    J, e, ep, Gc, hc = prob.fk_jac(q0, ft, pt, ct, dt=wl["dt"])
    res["k1_dense_ms"] = timed(lambda: prob.fk_jac(q0, ft, pt, ct, dt=wl["dt"]))
    res["k2_dense_ms"] = timed(lambda: prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"]))
    _, _, it = prob.solve(q0, J, e, ep, Gc, hc, wl["dt"], wl["damping"], return_iters=True)
    res["iters_mean"] = float(it.float().mean()); res["iters_max"] = int(it.max())
    res["describe"] = prob.describe(wl["damping"])
print(json.dumps(res))
'''


def parse(spec):
    wl = None
    if "#" in spec:
        spec, wl = spec.split("#", 1)
    env = {}
    if "@" in spec:
        spec, e = spec.split("@", 1)
        env = dict(kv.split("=", 1) for kv in e.split(",") if kv)
    flags = []
    if ":" in spec:
        spec, f = spec.split(":", 1)
        flags = [x for x in f.split(",") if x]
    if wl:
        parts = wl.split(",")
        env["BIK_WL"] = parts[0]
        if len(parts) > 1:
            env["BIK_B"] = parts[1]
    return spec, flags, env


def build_variant(name, flags):
    """Compile the four units with extra flags into mink_b200/lib/variants/<name>/ (objects cached by flag set)."""
    from mink_b200 import build as b

    out = os.path.join(REPO, "mink_b200", "lib", "variants", name)
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libbik.so")
    stamp = os.path.join(out, "flags.txt")
    newest = max(os.path.getmtime(s) for s in b.sources())
    if os.path.exists(so) and os.environ.get("BIK_AB_NOBUILD"):   # a variant built from another revision of the sources
        return so
    if os.path.exists(so) and os.path.getmtime(so) > newest and os.path.exists(stamp) and open(stamp).read() == " ".join(flags):
        return so
    nvcc = b._nvcc()

    def one(u):
        o = os.path.join(out, u[:-3] + ".o")
        subprocess.check_call([nvcc] + b.NVCC_FLAGS + flags + ["-c", "-o", o, os.path.join(b.CSRC, u)])
        return o

    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, b.UNITS))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs)
    open(stamp, "w").write(" ".join(flags))
    return so


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    specs = [parse(a) for a in (args or ["base"])]
    libs = {}
    for name, flags, _ in specs:
        if name not in libs:
            libs[name] = os.path.join(REPO, "mink_b200", "lib", "libbik.so") if (name == "base" and not flags) else build_variant(name, flags)
    if "--build-only" in sys.argv:
        print({k: v for k, v in libs.items()})
        return
    for name, flags, env in specs:
        e = dict(os.environ, BIK_REPO=REPO, BIK_LIB=libs[name], **env)
        r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-400:]
        print(name, flags, env, line, flush=True)


if __name__ == "__main__":
    main()
