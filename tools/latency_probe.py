"""Single-instance latency of a few problems (bench.py's latency_row): device time per step inside one bik_step call and host
wall time of one-step calls.  GPU box:  python tools/latency_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

flush = torch.empty(16 * 1024 * 1024, device="cuda:0")
cache = {}
for name in ("ur5e", "g1", "ur5e_wall", "spot"):
    print(name, json.dumps(bench.latency_row(torch, cache, name, flush)))
