#!/usr/bin/env python
"""Fused merge chain timing: python profiles/merge_chain_bench.py [N] [V]

Times ``smf_hmm_merge_chain`` (merge -> size classes -> read-span mask, packed output) on a synthetic
aggregate-feature layer resident in HBM, the three separate kernels it replaces, and the oracle's
restatement of the reference chain (per-read Python loops, as in the reference) on a bounded sample.
Algorithmic bytes per grid cell: 1 (layer u8) + 1 (code u8) + 4 (run length i32) = 6.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from oracle import hmm_oracle as O
from smftools_b200 import engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev)
gen.manual_seed(99)
# runs of 10-cell blocks: ~35 % membership, run lengths 10..100+, gaps 10..
blocks = (torch.rand((N, V // 10), device=dev, generator=gen) < 0.35).repeat_interleave(10, dim=1)
layer = blocks.to(torch.uint8).contiguous()
del blocks
coords_h = np.arange(V, dtype=np.int64) + 1000
coords = torch.from_numpy(coords_h).to(dev)
starts_h = 1000.0 + np.random.default_rng(1).integers(0, V // 10, size=N)
ends_h = 1000.0 + V - np.random.default_rng(2).integers(0, V // 10, size=N)
start = torch.from_numpy(starts_h).to(dev)
end = torch.from_numpy(ends_h).to(dev)
ranges = [(6.0, 40.0), (40.0, 100.0), (100.0, 200.0), (200.0, float("inf"))]


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def fused():
    return engine.merge_chain(layer, coords, True, 10, ranges, span_coords=coords, start=start, end=end)


def separate():
    merged, mlen, _, _ = engine.merge_runs(layer, coords, True, 10)
    _, _, cls, clen = engine.merge_runs(merged, coords, True, -(1 << 62), ranges)
    valid = engine.span_mask(N, V, dev, coords=coords, start=start, end=end)
    return merged, mlen, cls, clen, valid


ms_f = timed(fused)
ms_s = timed(separate)
peak = bench.measured_peak()[0]
cells = N * V
# CPU: oracle chain on a bounded sample of rows
n_cpu = 200
lay_h = layer[:n_cpu].cpu().numpy()
valid_h = O.span_valid(coords_h, starts_h[:n_cpu], ends_h[:n_cpu])
t0 = time.perf_counter()
O.merge_chain(lay_h, coords_h, 10, ranges, valid_h)
cpu_s = time.perf_counter() - t0
out = {
    "workload": f"merge chain on {N} reads x {V} grid cells (u8 layer), distance 10, 4 size classes, span mask",
    "fused_ms": ms_f, "fused_cells_per_s": cells / (ms_f * 1e-3), "fused_GBps": 6 * cells / (ms_f * 1e-3) / 1e9,
    "fused_frac_of_hbm_peak": 6 * cells / (ms_f * 1e-3) / 1e9 / float(peak),
    "separate_kernels_ms": ms_s,
    "cpu_port_cells_per_s": n_cpu * V / cpu_s, "cpu_sample": f"{n_cpu} rows, 1 core",
}
print(json.dumps(out))
