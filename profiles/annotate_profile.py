#!/usr/bin/env python
"""Host-side profile of annotate_adata + apply_merges on a duck-typed AnnData (N reads x V grid)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

import smftools_b200 as S
from oracle import hmm_oracle as O
from oracle.ref_loader import DuckAnnData

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
K = 2
X = O.synth_reads(N, L, K, 3, prob=False, nan_rate=0.3).astype(np.float64)
coords = O.synth_coords(L, 3, span_factor=12) + 1000
full = np.arange(coords[0] - 10, coords[-1] + 11)
V = len(full)
var_mask = np.isin(full, coords)
rng = np.random.default_rng(0)
obs = pd.DataFrame({"reference_start": full[0] + rng.integers(0, 40, size=N).astype(float),
                    "reference_end": full[-1] - rng.integers(0, 40, size=N).astype(float)},
                   index=[f"r{i}" for i in range(N)])
m = S.create_hmm({"hmm_n_states": K}, arch="single")
p = O.synth_params(K)
with torch.no_grad():
    m.start.data = torch.tensor(p.start); m.trans.data = torch.tensor(p.trans); m.emission.data = torch.tensor(p.emission)
fsets = S.normalize_hmm_feature_sets({
    "footprint": {"state": "Non-Modified", "features": {"small_bound_stretch": [6, 40], "medium_bound_stretch": [40, 100],
                                                        "putative_nucleosome": [100, 200], "large_bound_stretch": [200, "inf"]}},
    "accessible": {"state": "Modified", "features": {"small_accessible_patch": [3, 20], "mid_accessible_patch": [20, 40],
                                                     "large_accessible_patch": [40, 110], "nucleosome_depleted_region": [110, "inf"]}}})
cfg = {"hmm_merge_layer_features": [("all_footprint_features", 10), ("all_accessible_features", 60)]}


def run():
    ad = DuckAnnData(N, full, obs=obs)
    m.annotate_adata(ad, prefix="GpC", X=X, coords=coords, var_mask=var_mask, feature_sets=fsets)
    S.apply_merges(ad, m, "GpC", fsets, cfg)
    return ad


torch.cuda.synchronize()
t0 = time.perf_counter()
ad = run()  # cold: page-locked pool empty
dt = time.perf_counter() - t0
gb = sum(np.asarray(v).nbytes for v in ad.layers.values()) / 1e9
print(f"N={N} L={L} V={V}: [cold pool] annotate_adata + apply_merges {dt:.3f} s, {gb:.2f} GB of layers = {gb / dt:.1f} GB/s")
del ad
torch.cuda.synchronize()
t0 = time.perf_counter()
ad = run()
dt = time.perf_counter() - t0
print(f"N={N} L={L} V={V}: annotate_adata + apply_merges {dt:.3f} s, {len(ad.layers)} layers, "
      f"{sum(np.asarray(v).nbytes for v in ad.layers.values()) / 1e9:.2f} GB of layers = "
      f"{sum(np.asarray(v).nbytes for v in ad.layers.values()) / 1e9 / dt:.1f} GB/s [warm pool]")
del ad
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
