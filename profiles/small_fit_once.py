"""One small device-resident fit (1000 x 170, K = 2, 20 iterations) -- run under ncu's launch list to split an EM
iteration into its kernels:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_fit_launches.csv \
        python profiles/small_fit_once.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smftools_b200 as S  # noqa: E402

N, L, K = (int(v) for v in (sys.argv[1:4] + ["1000", "170", "2"][len(sys.argv) - 1:]))
g = torch.Generator(device="cuda").manual_seed(3)
x = (torch.rand((N, L), generator=g, device="cuda") < 0.35).to(torch.float32)
x[torch.rand((N, L), generator=g, device="cuda") < 0.3] = float("nan")
cfg = {"hmm_n_states": K, "hmm_init_emission_probs": list(np.linspace(0.8, 0.2, K))}
from smftools_b200 import engine  # noqa: E402

for loop in (0, 1):
    engine.set_option("fitloop", loop)  # 0: forward / backward / M-step kernels per iteration, 1: one cooperative launch
    for iters in (3, 20, 200):
        m = S.create_hmm(cfg, arch="single")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist = m.fit(x, None, max_iter=iters, tol=0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"fitloop={loop} fit N={N} L={L} K={K}: {len(hist)} iterations in {dt * 1e3:.3f} ms = "
              f"{dt / len(hist) * 1e6:.1f} us/iter, ll[-1]={hist[-1]:.6f}", flush=True)
