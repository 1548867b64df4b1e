"""Viterbi kernels: staged (viterbi_kernel) vs direct (viterbi_direct_kernel).  python profiles/vit_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine
dev = torch.device("cuda", 0)
PEAK = 6544.3
print(f"{'K':>2} {'L':>6} {'N':>8} {'obs':>4} {'staged ms':>10} {'frac':>5} {'direct ms':>10} {'frac':>5} {'Grp/s':>7}")
for K, L, N, prob in [(3, 10000, 262144, True), (3, 10000, 65536, True), (2, 5008, 100000, False), (4, 8000, 131072, False),
                      (2, 10000, 262144, False), (3, 1008, 1000000, True), (3, 304, 2000000, True)]:
    m = bench.make_model(K, dev); tables = m._host_tables()
    for u8 in ((False, True) if not prob and L % 16 == 0 else (False,)):
        obs = bench.device_reads(N, L, K, prob, 1, dev, coded_u8=u8)
        st = torch.empty((N, L), dtype=torch.int8, device=dev); ws = torch.empty((N * L,), dtype=torch.uint8, device=dev)
        res = {}
        for mode in (0, 1):
            engine.set_option("vit", mode)
            f = lambda: engine.viterbi(tables, obs, None, out_states=st, workspace=ws)
            f(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): f()
            e1.record(); torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / 3
        by = N * L * ((1 if u8 else 4) + 1) / 1e9
        print(f"{K:>2} {L:>6} {N:>8} {'u8' if u8 else 'f32':>4} {res[0]:10.3f} {by / res[0] * 1e3 / PEAK:5.2f} {res[1]:10.3f} "
              f"{by / res[1] * 1e3 / PEAK:5.2f} {N * L / res[1] / 1e6:7.1f}", flush=True)
        del obs, st, ws
engine.set_option("vit", -1)
