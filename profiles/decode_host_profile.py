#!/usr/bin/env python
"""Reference-compatible decode (float64 in, int64 states + float64 gamma out) timing: N L"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dev = torch.device("cuda", 0)
m = bench.make_model(2, dev)
X = bench.device_reads(N, L, 2, False, 1, dev).cpu().numpy().astype(np.float64)
from smftools_b200 import hostmem
for compact in (False, True):
    m.decode(X[:1024], decode="marginal", compact=compact)
    for rep in ("cold pool", "warm pool", "warm pool"):
        t0 = time.perf_counter(); st, g = m.decode(X, decode="marginal", compact=compact); dt = time.perf_counter() - t0
        print(f"decode(compact={compact}) {N}x{L} [{rep}]: {dt:.3f} s = {N*L/dt/1e9:.2f} Grp/s, out {st.dtype},{g.dtype} "
              f"{(st.nbytes+g.nbytes)/1e9:.1f} GB = {(st.nbytes+g.nbytes)/dt/1e9:.1f} GB/s; pool {hostmem.pool().stats}")
        del st, g
for decode in ("viterbi",):
    for rep in ("warm pool", "warm pool"):
        t0 = time.perf_counter(); st, g = m.decode(X, decode=decode); dt = time.perf_counter() - t0
        print(f"decode({decode}) {N}x{L} [{rep}]: {dt:.3f} s = {N*L/dt/1e9:.2f} Grp/s")
        del st, g
