import sys, json; sys.path.insert(0, '/root/repo')
import torch, bench
flush = torch.empty(16*1024*1024, device='cuda:0')
cache = {}
for name in ("ur5e", "g1", "ur5e_wall", "spot"):
    print(name, json.dumps(bench.latency_row(torch, cache, name, flush)))
