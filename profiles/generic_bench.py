"""Generic forward-backward kernel (fb_kernel: multi-channel / distance-binned models) at the C2 shape.
    python profiles/generic_bench.py  ->  profiles/r02_generic.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import bench
from helpers import model_from_params
from oracle import hmm_oracle as O
from smftools_b200 import engine

dev = torch.device("cuda", 0)
PEAK = 6544.3
N, L = 100_000, 5_000


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"generic path at {N} reads x {L} positions (float32 observations, 30 % missing)")
print(f"{'arch':>24} {'K':>2} {'C':>2} {'op':>10} {'ms':>8} {'Grp/s':>7} {'algo B/rp':>9} {'frac HBM':>8}")
for arch, K, C in [("single", 2, 1), ("multi", 2, 2), ("multi", 3, 2), ("multi", 2, 3), ("single_distance_binned", 2, 1),
                   ("single_distance_binned", 3, 1)]:
    kw = {"distance_bins": [1, 5, 10, 25, 50, 100]} if arch == "single_distance_binned" else {}
    p = O.synth_params(K, arch=arch, n_channels=C, **kw)
    m = model_from_params(p, arch=arch)
    tables = m._host_tables()
    coords = np.cumsum(np.random.default_rng(3).integers(1, 40, size=L)).astype(np.int64)
    bins = m._bins_for(coords, L, dev)
    base = bench.device_reads(N, L, K, False, 1, dev)
    obs = base if C == 1 else torch.stack([base] + [torch.roll(base, 7 * c, dims=1) for c in range(1, C)], dim=2).contiguous()
    del base
    gam = torch.empty((N, L, K), dtype=torch.float32, device=dev)
    st = torch.empty((N, L), dtype=torch.int8, device=dev)
    ms = timeit(lambda: engine.posterior(tables, obs, bins, out_gamma=gam, out_states=st))
    by = 4 * C + 4 * K + 1
    print(f"{arch:>24} {K:>2} {C:>2} {'posterior':>10} {ms:8.3f} {N * L / ms / 1e6:7.1f} {by:>9} {N * L * by / ms / 1e6 / PEAK:8.2f}", flush=True)
    ws = engine.estep_workspace(tables, dev, N, L)
    out = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
    ms = timeit(lambda: engine.estep(tables, obs, bins, out_stats=out, workspace=ws))
    by = 4 * C
    print(f"{arch:>24} {K:>2} {C:>2} {'E-step':>10} {ms:8.3f} {N * L / ms / 1e6:7.1f} {by:>9} {N * L * by / ms / 1e6 / PEAK:8.2f}", flush=True)
    del obs, gam, st, ws
