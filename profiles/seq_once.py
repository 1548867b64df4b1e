#!/usr/bin/env python
"""One E-step launch sequence for ncu: K L N [u8|f32] [seq mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine
K, L, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
u8 = (sys.argv[4] if len(sys.argv) > 4 else "u8") == "u8"
engine.set_option("seq", int(sys.argv[5]) if len(sys.argv) > 5 else -1)
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev)
tables = m._host_tables()
obs = bench.device_reads(N, L, K, False, 1, dev, coded_u8=u8)
ws = engine.estep_workspace(tables, dev, N, L)
st = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
for _ in range(2):
    engine.estep(tables, obs, None, out_stats=st, workspace=ws)
torch.cuda.synchronize()
print(st.cpu().numpy()[-1])
