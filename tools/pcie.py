"""PCIe ceiling on this box (pinned host memory): H2D alone, D2H alone, both at once; then bik_step_host for several chunk plans."""
import os, sys, time, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mink_b200._abi import spec_from_workload
from mink_b200.engine import DeviceModel, Problem
from mink_b200.workloads import WORKLOADS, make_inputs, load_flat, task_frames

dev = torch.device("cuda:0")
n = 32 * 1024 * 1024
h1 = torch.empty(n // 4, dtype=torch.float32, pin_memory=True); h2 = torch.empty(n // 4, dtype=torch.float32, pin_memory=True)
d1 = torch.empty(n // 4, dtype=torch.float32, device=dev); d2 = torch.empty(n // 4, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def up():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): up(); down()
res = {"h2d_gbs": n / t(up) / 1e9, "d2h_gbs": n / t(down) / 1e9, "both_gbs_each": n / t(both) / 1e9}
wl = WORKLOADS["g1"]; fm = load_flat("g1"); spec = spec_from_workload(fm, wl)
model = DeviceModel(fm, 0); prob = Problem(model, spec); frames = task_frames(wl, fm)
def fk(qq):
    p, c = model.fk(qq, frames); return p.cpu().numpy().astype(np.float64), None
B = 65536
inp = make_inputs(fm, wl, B, fk, seed=1000)
def pinned(a, dtype=torch.float32):
    x = torch.empty(a.shape, dtype=dtype, pin_memory=True); x.copy_(torch.as_tensor(np.ascontiguousarray(a)).to(dtype)); return x.numpy()
hq0 = pinned(inp["q"]); hq = pinned(inp["q"]); hft = pinned(inp["frame_targets"]); hpt = pinned(inp["posture_target"])
hdq = pinned(np.zeros((B, fm.nv), np.float32)); hst = pinned(np.zeros(B, np.int32), torch.int32)
for chunks in ("0", "1", "2", "4", "8", "16"):
    os.environ["BIK_HOST_CHUNKS"] = chunks
    ts = []
    for i in range(8):
        hq[:] = hq0
        t0 = time.perf_counter()
        prob.step_host(hq, hft, hpt, None, dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=True, out_dq=hdq, out_status=hst)
        ts.append(time.perf_counter() - t0)
    res[f"step_host_ms_chunks_{chunks}"] = 1e3 * statistics.median(ts[2:])
print(json.dumps(res))
os.environ["BIK_HOST_CHUNKS"] = "0"; os.environ["BIK_HOST_TRACE"] = "1"
for i in range(2):
    hq[:] = hq0
    t0 = time.perf_counter()
    prob.step_host(hq, hft, hpt, None, dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=True, out_dq=hdq, out_status=hst)
    print("wall ms", 1e3 * (time.perf_counter() - t0))
os.environ["BIK_HOST_CHUNKS"] = "1"
hq[:] = hq0
t0 = time.perf_counter()
prob.step_host(hq, hft, hpt, None, dt=wl["dt"], damping=wl["damping"], nsteps=1, integrate=True, out_dq=hdq, out_status=hst)
print("wall ms (1 chunk)", 1e3 * (time.perf_counter() - t0))
