#!/usr/bin/env python
"""One generic-path launch for ncu: arch(multi|binned) K [post|estep]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from helpers import model_from_params
from oracle import hmm_oracle as O
from smftools_b200 import engine
arch = {"multi": "multi", "binned": "single_distance_binned"}[sys.argv[1]]
K = int(sys.argv[2]); op = sys.argv[3] if len(sys.argv) > 3 else "post"
N, L, C = 100_000, 5_000, (2 if arch == "multi" else 1)
dev = torch.device("cuda", 0)
kw = {"distance_bins": [1, 5, 10, 25, 50, 100]} if arch != "multi" else {}
p = O.synth_params(K, arch=arch, n_channels=C, **kw)
m = model_from_params(p, arch=arch); tables = m._host_tables()
coords = np.cumsum(np.random.default_rng(3).integers(1, 40, size=L)).astype(np.int64)
bins = m._bins_for(coords, L, dev)
base = bench.device_reads(N, L, K, False, 1, dev)
obs = base if C == 1 else torch.stack([base, torch.roll(base, 7, dims=1)], dim=2).contiguous()
if op == "post":
    gam = torch.empty((N, L, K), dtype=torch.float32, device=dev); st = torch.empty((N, L), dtype=torch.int8, device=dev)
    for _ in range(2): engine.posterior(tables, obs, bins, out_gamma=gam, out_states=st)
else:
    ws = engine.estep_workspace(tables, dev, N, L); out = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
    for _ in range(2): engine.estep(tables, obs, bins, out_stats=out, workspace=ws)
torch.cuda.synchronize(); print("ok")
