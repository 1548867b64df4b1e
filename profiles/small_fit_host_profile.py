"""Host-side cost of one small fit (1000 x 170, K = 2, 3 iterations: the kernels take < 0.1 ms): cProfile of 300 calls."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smftools_b200 as S  # noqa: E402

N, L, K = 1000, 170, 2
g = torch.Generator(device="cuda").manual_seed(3)
x = (torch.rand((N, L), generator=g, device="cuda") < 0.35).to(torch.float32)
x[torch.rand((N, L), generator=g, device="cuda") < 0.3] = float("nan")
xh = x.cpu().numpy().astype(np.float64)
cfg = {"hmm_n_states": K, "hmm_init_emission_probs": [0.8, 0.2]}


def once(inp, iters=3):
    m = S.create_hmm(cfg, arch="single")
    return m.fit(inp, None, max_iter=iters, tol=0.0)


for name, inp in (("device float32", x), ("host float64 numpy", xh)):
    for _ in range(20):
        once(inp)
    torch.cuda.synchronize()
    for iters in (3, 20):
        t0 = time.perf_counter()
        for _ in range(300):
            once(inp, iters)
        torch.cuda.synchronize()
        print(f"{name}: create_hmm + fit({iters} iterations) = {(time.perf_counter() - t0) / 300 * 1e6:.0f} us per call", flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        once(inp)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)
