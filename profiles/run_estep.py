#!/usr/bin/env python
"""Run a few E-step launches on a synthetic batch (for ncu): python profiles/run_estep.py K N L"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine
K, N, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev); t = m._host_tables()
obs = bench.device_reads(N, L, K, False, 1, dev)
ws = engine.estep_workspace(t, dev, N, L)
for _ in range(4):
    s = engine.estep(t, obs, None, workspace=ws)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    s = engine.estep(t, obs, None, workspace=ws)
b.record(); torch.cuda.synchronize()
print(f"estep K={K} N={N} L={L}: {a.elapsed_time(b)/5:.3f} ms/iter, {N*L/(a.elapsed_time(b)/5)/1e6:.1f} Grp/s")
