#!/usr/bin/env python
"""Posterior-kernel timing sweep over read length / state count (uses the current SMF_FB2_* env)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from smftools_b200 import engine

dev = torch.device("cuda", 0)
for K in [int(k) for k in (sys.argv[1] if len(sys.argv) > 1 else "2").split(",")]:
    m = bench.make_model(K, dev)
    t = m._host_tables()
    for L in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "170,500,1000,2000,5000,10000,20000").split(",")]:
        N = max(1024, int(4e8 // L))
        obs = bench.device_reads(N, L, K, False, 1, dev)
        g = torch.empty((N, L, K), dtype=torch.float32, device=dev)
        s = torch.empty((N, L), dtype=torch.int8, device=dev)
        for _ in range(3):
            engine.posterior(t, obs, None, out_gamma=g, out_states=s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            engine.posterior(t, obs, None, out_gamma=g, out_states=s)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        bytes_ = (4 + 4 * K + 1) * N * L
        print(f"K={K} L={L:6d} N={N:8d} {ms:8.3f} ms  {N*L/ms/1e6:8.1f} Grp/s  {bytes_/ms/1e6/6544.3:6.3f} of HBM peak", flush=True)
        del obs, g, s
