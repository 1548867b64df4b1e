#!/usr/bin/env python
"""Second audit: multi-channel / distance-binned models (fast-path variants vs the generic kernel) and the two
Viterbi kernels, default vs forced, over small to large shapes.
    python profiles/policy_audit2.py > profiles/r02_policy_audit2.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from helpers import model_from_params
from oracle import hmm_oracle as O
from smftools_b200 import engine

dev = torch.device("cuda", 0)


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("# forward-backward of multi-channel / distance-binned models: default (fast-path variant where eligible) vs generic kernel, ms")
print(f"{'arch':>8} {'op':>6} {'K':>2} {'C':>2} {'N':>7} {'L':>6} | {'default':>9} {'generic':>9}")
flag = 0
for arch, K, C in [("multi", 2, 2), ("multi", 3, 3), ("binned", 2, 1), ("binned", 3, 1)]:
    a = "multi" if arch == "multi" else "single_distance_binned"
    kw = {"distance_bins": [1, 5, 10, 25, 50, 100]} if arch == "binned" else {}
    m = model_from_params(O.synth_params(K, arch=a, n_channels=C, **kw), arch=a); tables = m._host_tables()
    for N, L in [(100, 50), (100, 1000), (1000, 170), (1000, 3000), (10000, 500), (10000, 8000), (200000, 170), (100000, 2000)]:
        bins = m._bins_for(np.cumsum(np.random.default_rng(3).integers(1, 40, size=L)).astype(np.int64), L, dev)
        base = bench.device_reads(N, L, K, False, 1, dev)
        obs = base if C == 1 else torch.stack([base] + [torch.roll(base, 7 * c, dims=1) for c in range(1, C)], dim=2).contiguous()
        gam = torch.empty((N, L, K), dtype=torch.float32, device=dev); st = torch.empty((N, L), dtype=torch.int8, device=dev)
        out = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
        for op, f in (("post", lambda: engine.posterior(tables, obs, bins, out_gamma=gam, out_states=st)),
                      ("estep", lambda: engine.estep(tables, obs, bins, out_stats=out))):
            engine.set_option("generic", 0); t0 = timeit(f)
            engine.set_option("generic", 1); t1 = timeit(f)
            engine.set_option("generic", 0)
            mark = "  <-- generic faster" if t1 < 0.87 * t0 and t0 > 0.05 else ""
            flag += bool(mark)
            print(f"{arch:>8} {op:>6} {K:>2} {C:>2} {N:>7} {L:>6} | {t0:9.3f} {t1:9.3f}{mark}", flush=True)
        del obs, base, gam, st
print(f"# shapes where the generic kernel beats the default by > 15 %: {flag}")
print("# Viterbi: default policy vs staged (vit = 0) vs direct (vit = 1), ms")
print(f"{'K':>2} {'N':>7} {'L':>6} {'obs':>4} | {'default':>9} {'staged':>9} {'direct':>9}")
flag = 0
for K in (2, 3, 4):
    m = bench.make_model(K, dev); tables = m._host_tables()
    for N, L in [(1000, 176), (1000, 4000), (20000, 176), (20000, 2000), (20000, 10000), (300000, 176), (200000, 2000)]:
        for u8, prob in ((False, True), (True, False)):
            obs = bench.device_reads(N, L, K, prob, 1, dev, coded_u8=u8)
            st = torch.empty((N, L), dtype=torch.int8, device=dev); ws = torch.empty((N * L,), dtype=torch.uint8, device=dev)
            f = lambda: engine.viterbi(tables, obs, None, out_states=st, workspace=ws)
            res = []
            for mode in (-1, 0, 1):
                engine.set_option("vit", mode); res.append(timeit(f))
            engine.set_option("vit", -1)
            best = min(res)
            mark = "  <-- default %.2fx slower" % (res[0] / best) if res[0] > 1.15 * best and res[0] > 0.05 else ""
            flag += bool(mark)
            print(f"{K:>2} {N:>7} {L:>6} {'u8' if u8 else 'f32':>4} | {res[0]:9.3f} {res[1]:9.3f} {res[2]:9.3f}{mark}", flush=True)
            del obs, st, ws
print(f"# Viterbi shapes where the default is > 15 % slower than the better kernel: {flag}")
