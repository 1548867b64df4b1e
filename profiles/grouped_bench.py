"""Grouped launches vs per-group calls, and the sequential posterior kernels vs the chunk-parallel ones.

    python profiles/grouped_bench.py [post] [small] [c5]

post   one big group through smf_hmm_grouped_posterior (seq_forward + seq_backward<MODE_POST>) against
       smf_hmm_posterior's default path, K = 2, 3, 4, uint8-coded and float observations
small  the reference's regime: G groups of <= 1000 reads, 20 EM iterations each -- one grouped device-resident
       fit against a Python loop of model.fit calls
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import smftools_b200 as S  # noqa: E402
from smftools_b200 import engine, grouped  # noqa: E402

PEAK = 6544.3


def synth_dev(N, L, K, seed, prob, dev, u8):
    g = torch.Generator(device=dev).manual_seed(seed)
    r = torch.rand((N, L), generator=g, device=dev)
    x = (r < 0.35).to(torch.float32) if not prob else torch.rand((N, L), generator=g, device=dev)
    miss = torch.rand((N, L), generator=g, device=dev) < 0.3
    if u8:
        return torch.where(miss, torch.tensor(255, dtype=torch.uint8, device=dev), x.to(torch.uint8))
    return torch.where(miss, torch.tensor(float("nan"), device=dev), x)


def model(K):
    from oracle import hmm_oracle as O
    from helpers import model_from_params

    return model_from_params(O.synth_params(K))


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_post():
    dev = torch.device("cuda")
    print(f"{'K':>2} {'L':>6} {'N':>8} {'obs':>4} {'default ms':>11} {'frac':>5} {'seq ms':>8} {'frac':>5}  max|dgamma|")
    for K, L, N in [(2, 5008, 100000), (3, 5008, 100000), (4, 5008, 100000), (3, 10000, 65536), (4, 8000, 65536),
                    (2, 10000, 65536), (4, 1008, 400000), (3, 1008, 400000), (3, 304, 1000000), (4, 304, 1000000)]:
        m = model(K)
        tables = m._host_tables()
        for u8 in (True, False):
            obs = synth_dev(N, L, K, 1, False, dev, u8)
            es = 1 if u8 else 4
            gam = torch.empty((N, L, K), dtype=torch.float32, device=dev)
            st = torch.empty((N, L), dtype=torch.int8, device=dev)
            t_def = timeit(lambda: engine.posterior(tables, obs, out_gamma=gam, out_states=st))
            g_def = gam[:256].clone()
            X = S.CodedU8(obs) if u8 else obs
            res = [None]

            def run():
                res[0] = grouped.decode_groups([m], [X])

            t_seq = timeit(run)
            d = (res[0][0][0][:256] - g_def).abs().max().item()
            by = N * L * (es + 4 * K + 1) / 1e9
            print(f"{K:>2} {L:>6} {N:>8} {'u8' if u8 else 'f32':>4} {t_def:11.3f} {by / t_def * 1e3 / PEAK:5.2f} {t_seq:8.3f} "
                  f"{by / t_seq * 1e3 / PEAK:5.2f}  {d:.1e}", flush=True)
            del gam, st, obs, res


def bench_small():
    from oracle import hmm_oracle as O

    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    for K, G, iters in [(2, 64, 20), (2, 512, 20), (3, 512, 20), (4, 512, 20)]:
        obs, shapes = [], []
        for g in range(G):
            N = int(rng.integers(200, 1001))
            L = int(rng.integers(100, 3000))
            shapes.append((N, L))
            obs.append(S.CodedU8(synth_dev(N, L, K, 100 + g, False, dev, True)))
        pos = sum(n * l for n, l in shapes)
        from helpers import model_from_params

        def fresh():
            return [model_from_params(O.synth_params(K)) for _ in range(G)]

        ms = fresh()
        grouped.fit_groups(ms, obs, max_iter=iters, tol=0.0)
        torch.cuda.synchronize()
        ms = fresh()
        t0 = time.perf_counter()
        grouped.fit_groups(ms, obs, max_iter=iters, tol=0.0)
        torch.cuda.synchronize()
        t_g = time.perf_counter() - t0
        def fresh_host_loop():
            out = fresh()
            for m in out:
                m._host_loop_only = True  # the round-1 path: E-step kernel + statistics read-back + host M-step per iteration
            return out

        ml = fresh_host_loop()
        for m, x in list(zip(ml, obs))[:4]:
            m.fit(x, max_iter=2, tol=0.0)
        ml = fresh_host_loop()
        t0 = time.perf_counter()
        for m, x in zip(ml, obs):
            m.fit(x, max_iter=iters, tol=0.0)
        torch.cuda.synchronize()
        t_l = time.perf_counter() - t0
        dmax = max(float((a.emission - b.emission).abs().max()) for a, b in zip(ms, ml))
        print(f"K={K} G={G} groups x {iters} iterations, {pos / 1e6:.1f} M read-positions: grouped fit {t_g * 1e3:.1f} ms "
              f"({pos * iters / t_g / 1e9:.1f} Grp/s, {t_g / iters / G * 1e6:.1f} us per group-iteration) | per-group loop "
              f"{t_l * 1e3:.1f} ms ({t_l / iters / G * 1e6:.1f} us per group-iteration) | speed-up {t_l / t_g:.1f}x | "
              f"max |d emission| {dmax:.1e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["post", "small"]
    if "post" in what:
        bench_post()
    if "small" in what:
        bench_small()
