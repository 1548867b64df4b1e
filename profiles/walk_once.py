#!/usr/bin/env python
"""One posterior call per (K, L) for ncu launch lists: python profiles/walk_once.py K L [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine

K, L = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else max(1024, int(4e8 // L))
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev)
t = m._host_tables()
obs = bench.device_reads(N, L, K, False, 1, dev)
g = torch.empty((N, L, K), dtype=torch.float32, device=dev)
s = torch.empty((N, L), dtype=torch.int8, device=dev)
for _ in range(2):
    engine.posterior(t, obs, None, out_gamma=g, out_states=s)
torch.cuda.synchronize()
