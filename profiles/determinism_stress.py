#!/usr/bin/env python
"""Run every kernel family repeatedly on the same inputs and require bit-identical outputs: data races
(e.g. an asynchronous copy overtaking shared-memory loads) show up as run-to-run differences."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from smftools_b200 import engine

dev = torch.device("cuda", 0)
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 25
bad = 0
cases = []
for K in (2, 3, 4):
    for L, N in ((34, 400000), (170, 150000), (300, 90000), (450, 60000), (700, 40000), (1200, 30000), (2500, 20000),
                 (5000, 12000), (10000, 9000), (16000, 8200), (24000, 8200)):
        cases.append((K, L, N))
for K, L, N in cases:
    m = bench.make_model(K, dev)
    t = m._host_tables()
    obs = bench.device_reads(N, L, K, K == 3, 7 + L, dev)
    ref = None
    for r in range(REPS):
        g, s, ll = engine.posterior(t, obs, None, want_loglik=True)
        st = engine.estep(t, obs, None)
        cur = (g, s, ll, st)
        if ref is None:
            ref = tuple(x.clone() for x in cur)
        else:
            same = all(torch.equal(a.view(torch.uint8), b.view(torch.uint8)) for a, b in zip(ref, cur))
            if not same:
                bad += 1
                print(f"NONDETERMINISTIC K={K} L={L} N={N} rep={r}", flush=True)
                break
    del obs, ref
    print(f"ok K={K} L={L} N={N}", flush=True)
# forced paths
for opt, val in (("walk", 1), ("walk", 2), ("short", 1)):
    engine.set_option(opt, val)
    for K in (2, 3, 4):
        for L, N in ((170, 60000), (450, 30000), (1200, 12000), (5000, 9000)):
            if opt == "walk" and val == 1 and K < 3:
                continue
            if opt == "short" and L > 900:
                continue
            m = bench.make_model(K, dev); t = m._host_tables()
            obs = bench.device_reads(N, L, K, False, 11 + L, dev)
            ref = None
            for r in range(REPS):
                g, s, ll = engine.posterior(t, obs, None, want_loglik=True)
                cur = (g, s, ll)
                if ref is None:
                    ref = tuple(x.clone() for x in cur)
                elif not all(torch.equal(a.view(torch.uint8), b.view(torch.uint8)) for a, b in zip(ref, cur)):
                    bad += 1
                    print(f"NONDETERMINISTIC {opt}={val} K={K} L={L} rep={r}", flush=True)
                    break
            print(f"ok {opt}={val} K={K} L={L}", flush=True)
    engine.set_option(opt, -1)
print("nondeterministic cases:", bad)
