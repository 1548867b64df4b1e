#!/usr/bin/env python
"""One posterior launch for ncu: K L N [u8|f32] [prob]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine
K, L, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
u8 = (sys.argv[4] if len(sys.argv) > 4 else "f32") == "u8"
prob = len(sys.argv) > 5 and sys.argv[5] == "prob"
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev)
tables = m._host_tables()
obs = bench.device_reads(N, L, K, prob, 1, dev, coded_u8=u8)
gam = torch.empty((N, L, K), dtype=torch.float32, device=dev)
st = torch.empty((N, L), dtype=torch.int8, device=dev)
for _ in range(2):
    engine.posterior(tables, obs, out_gamma=gam, out_states=st)
torch.cuda.synchronize()
print(float(gam[0, :8].sum()))
