#!/usr/bin/env python
"""One Viterbi launch for ncu: K L N [u8|f32] [prob] [vit mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine
K, L, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
u8 = sys.argv[4] == "u8"; prob = sys.argv[5] == "prob"
engine.set_option("vit", int(sys.argv[6]))
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev); tables = m._host_tables()
obs = bench.device_reads(N, L, K, prob, 1, dev, coded_u8=u8)
st = torch.empty((N, L), dtype=torch.int8, device=dev); ws = torch.empty((N * L,), dtype=torch.uint8, device=dev)
for _ in range(2): engine.viterbi(tables, obs, None, out_states=st, workspace=ws)
torch.cuda.synchronize(); print(int(st[0, :8].sum()))
