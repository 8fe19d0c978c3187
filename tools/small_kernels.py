import os, sys, json, statistics
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mink_b200._abi import spec_from_workload
from mink_b200.engine import DeviceModel, Problem
from mink_b200.workloads import WORKLOADS
from tests.helpers import load_flat
wl = WORKLOADS["g1"]; fm = load_flat("g1"); spec = spec_from_workload(fm, wl)
model = DeviceModel(fm, 0)
B = 65536
q = torch.tensor(np.tile(fm.key(wl["key"]), (B, 1)), dtype=torch.float32, device="cuda:0")
dq = torch.full((B, fm.nv), 1e-3, device="cuda:0")
flush = torch.empty(64 * 1024 * 1024, device="cuda:0")
ti, tc = [], []
for s in range(8):
    flush.fill_(1.0)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); model.integrate(q, dq); b.record(); model.check_limits(q); c.record(); torch.cuda.synchronize()
    ti.append(a.elapsed_time(b)); tc.append(b.elapsed_time(c))
print("integrate_ms", statistics.median(ti[2:]), "check_limits_ms", statistics.median(tc[2:]))
