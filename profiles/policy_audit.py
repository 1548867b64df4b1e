#!/usr/bin/env python
"""Policy audit: for a grid of (op, K, N, L, dtype) time the library's default choice against every forced
alternative (generic kernel, two-kernel path, sequential kernels, bundled short-read kernel on / off) and flag
the shapes where the default is more than 15 % slower than the best alternative.
    python profiles/policy_audit.py > profiles/r02_policy_audit.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from smftools_b200 import engine

dev = torch.device("cuda", 0)
VARIANTS = {"default": {}, "generic": {"generic": 1}, "walk": {"walk": 1}, "seq": {"seq": 1}, "short": {"short": 1},
            "noshort": {"short": 0}, "noseq": {"seq": 0}, "nowalk": {"walk": 0}}
RESET = {"generic": 0, "walk": -1, "seq": -1, "short": -1}


def timeit(f, reps=3):
    try:
        f(); torch.cuda.synchronize()
    except Exception as e:  # a forced path may reject the shape
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flagged = 0
print(f"{'op':>6} {'K':>2} {'N':>8} {'L':>6} {'obs':>4} | " + " ".join(f"{v:>9}" for v in VARIANTS) + " | best")
for op in ("post", "estep"):
    for K in (2, 3, 4):
        m = bench.make_model(K, dev); tables = m._host_tables()
        for N, L in [(1000, 100), (1000, 500), (1000, 2000), (1000, 8000), (1000, 12000), (1000, 20000),
                     (8000, 100), (8000, 500), (8000, 2000), (8000, 8000), (8000, 12000), (8000, 20000),
                     (65536, 100), (65536, 500), (65536, 2000), (65536, 8000), (30000, 12000), (20000, 20000),
                     (1000000, 100), (800000, 500), (200000, 2000)]:
            for u8 in (False, True):
                obs = bench.device_reads(N, L, K, False, 1, dev, coded_u8=u8)
                if op == "post":
                    gam = torch.empty((N, L, K), dtype=torch.float32, device=dev); st = torch.empty((N, L), dtype=torch.int8, device=dev)
                    f = lambda: engine.posterior(tables, obs, out_gamma=gam, out_states=st)
                else:
                    out = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
                    f = lambda: engine.estep(tables, obs, None, out_stats=out)
                res = {}
                for name, opts in VARIANTS.items():
                    for k, v in RESET.items(): engine.set_option(k, v)
                    for k, v in opts.items(): engine.set_option(k, v)
                    res[name] = timeit(f)
                for k, v in RESET.items(): engine.set_option(k, v)
                ok = {k: v for k, v in res.items() if v is not None}
                best = min(ok, key=ok.get)
                flag = " <-- default %.2fx slower" % (ok["default"] / ok[best]) if ok["default"] > 1.15 * ok[best] else ""
                flagged += bool(flag)
                print(f"{op:>6} {K:>2} {N:>8} {L:>6} {'u8' if u8 else 'f32':>4} | " +
                      " ".join(f"{res[v]:9.3f}" if res[v] is not None else f"{'-':>9}" for v in VARIANTS) + f" | {best}{flag}", flush=True)
                del obs
                if op == "post": del gam, st
print(f"# shapes where the default is > 15 % slower than the best forced alternative: {flagged}")
