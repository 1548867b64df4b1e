#!/usr/bin/env python
"""Host-side cost of small decode / annotate calls (the reference's regime: ~1000 reads x ~170 sites per subset)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smftools_b200 as S
from oracle import hmm_oracle as O
X = O.synth_reads(1000, 170, 2, 3, prob=False, nan_rate=0.3).astype(np.float64)
m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
for _ in range(5): m.decode(X)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): st, g = m.decode(X)
dt = (time.perf_counter() - t0) / 100
print(f"decode 1000 x 170 (float64 in, int64 + float64 out): {dt*1e3:.3f} ms per call")
for _ in range(100): st, g = m.decode(X, decode="viterbi")
pr = cProfile.Profile(); pr.enable()
for _ in range(100): st, g = m.decode(X)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
