#!/usr/bin/env python
"""Per-iteration latency of fit() on a tiny (amplicon-sized) problem: launch + sync + host M-step overheads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smftools_b200 as S
from oracle import hmm_oracle as O
for N, L in ((500, 170), (1000, 170), (20000, 170), (2000, 2500)):
    X = O.synth_reads(N, L, 2, 3, prob=False, nan_rate=0.3).astype(np.float64)
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    m.fit(X, None, max_iter=3, tol=0.0)
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist = m.fit(X, None, max_iter=50, tol=0.0)
    dt = time.perf_counter() - t0
    print(f"fit N={N} L={L}: {len(hist)} iterations in {dt*1e3:.1f} ms = {dt/len(hist)*1e3:.3f} ms/iter")
    t0 = time.perf_counter(); st, g = m.decode(X); dt = time.perf_counter() - t0
    print(f"decode N={N} L={L}: {dt*1e3:.2f} ms")

# the same small fit with the host loop forced (E-step kernel + stats read-back + host M-step per iteration)
for N, L in ((1000, 170),):
    X = O.synth_reads(N, L, 2, 3, prob=False, nan_rate=0.3).astype(np.float64)
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    m._host_loop_only = True
    m.fit(X, None, max_iter=3, tol=0.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hist = m.fit(X, None, max_iter=50, tol=0.0)
    dt = time.perf_counter() - t0
    print(f"fit (host loop) N={N} L={L}: {len(hist)} iterations in {dt*1e3:.1f} ms = {dt/len(hist)*1e3:.3f} ms/iter")
    Xd = torch.from_numpy(X.astype(np.float32)).cuda()
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    m.fit(Xd, None, max_iter=3, tol=0.0)
    for iters in (50, 200):
        m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hist = m.fit(Xd, None, max_iter=iters, tol=0.0)
        dt = time.perf_counter() - t0
        print(f"fit (device loop, device-resident input) N={N} L={L}: {len(hist)} iterations in {dt*1e3:.2f} ms = {dt/len(hist)*1e6:.1f} us/iter")
