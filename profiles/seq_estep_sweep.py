#!/usr/bin/env python
"""E-step throughput: sequential (thread-per-read) kernels vs the chunk-parallel paths, by K / L / N / dtype."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine

dev = torch.device("cuda", 0)
shapes = [(4, 8000, n) for n in (16384, 32768, 65536, 131072, 262144)] + \
         [(3, 8000, n) for n in (32768, 65536, 262144)] + [(2, 8000, n) for n in (32768, 65536, 131072, 262144)] + \
         [(4, 1008, 262144), (4, 1008, 1048576), (2, 5008, 100000), (2, 1008, 1048576), (3, 10000, 262144),
          (4, 304, 1048576), (2, 304, 1048576), (2, 176, 2000000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in s.split(",")) for s in sys.argv[1:]]
print(f"{'K':>2} {'L':>6} {'N':>8} {'obs':>4} {'default ms':>11} {'seq ms':>9} {'seq Grp/s':>10} {'speedup':>8}")
for K, L, N in shapes:
    m = bench.make_model(K, dev)
    tables = m._host_tables()
    for u8 in (True, False):
        obs = bench.device_reads(N, L, K, False, 1, dev, coded_u8=u8)
        res = {}
        for mode in (0, 1):
            engine.set_option("seq", mode)
            ws = engine.estep_workspace(tables, dev, N, L)
            st = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
            for _ in range(2):
                engine.estep(tables, obs, None, out_stats=st, workspace=ws)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                engine.estep(tables, obs, None, out_stats=st, workspace=ws)
            b.record()
            torch.cuda.synchronize()
            res[mode] = (a.elapsed_time(b) / 3, st.cpu().numpy().copy())
            del ws
        engine.set_option("seq", -1)
        rel = abs(res[0][1] - res[1][1]).max() / max(1.0, abs(res[0][1]).max())
        print(f"{K:>2} {L:>6} {N:>8} {'u8' if u8 else 'f32':>4} {res[0][0]:>11.3f} {res[1][0]:>9.3f} "
              f"{N * L / res[1][0] / 1e6:>10.1f} {res[0][0] / res[1][0]:>8.2f}  (max rel diff {rel:.1e})")
        del obs
        torch.cuda.empty_cache()
