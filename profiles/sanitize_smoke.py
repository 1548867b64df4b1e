#!/usr/bin/env python
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import hmm_oracle as O
from smftools_b200 import engine
import smftools_b200 as S

dev = torch.device("cuda", 0)


def tables(K, arch="single", **kw):
    cfg = {"hmm_n_states": K, **kw}
    m = S.create_hmm(cfg, arch=arch)
    p = O.synth_params(K, arch=arch, **({"n_channels": kw["hmm_n_channels"]} if "hmm_n_channels" in kw else {}),
                       **({"distance_bins": kw["hmm_distance_bins"]} if "hmm_distance_bins" in kw else {}))
    with torch.no_grad():
        m.start.data = torch.tensor(p.start)
        m.trans.data = torch.tensor(p.trans)
        m.emission.data = torch.tensor(p.emission).reshape(m.emission.shape)
        if p.trans_by_bin is not None:
            m.trans_by_bin.data = torch.tensor(p.trans_by_bin)
    return m, m._host_tables()


def run(tag, fn):
    fn()
    torch.cuda.synchronize()
    print("ok", tag, flush=True)


for K in (2, 3, 4):
    m, t = tables(K)
    for L, N in ((1, 37), (45, 101), (171, 203), (333, 77), (600, 41), (1203, 19), (5000, 7), (30000, 3)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 5 + L, prob=True, nan_rate=0.3)).to(dev)
        run(f"posterior K{K} L{L}", lambda: engine.posterior(t, X, None, want_loglik=True))
        run(f"posterior-one K{K} L{L}", lambda: engine.posterior(t, X, None, gamma_state=1))
        run(f"estep K{K} L{L}", lambda: engine.estep(t, X, None))
        run(f"viterbi K{K} L{L}", lambda: engine.viterbi(t, X, None))
        X8 = torch.where(torch.isnan(X), torch.tensor(255.0, device=dev), (X > 0.5).float()).to(torch.uint8)
        run(f"posterior-u8 K{K} L{L}", lambda: engine.posterior(t, X8, None))
        run(f"estep-u8 K{K} L{L}", lambda: engine.estep(t, X8, None))
    # two-kernel path forced on a small batch
    engine.set_option("walk", 1)
    engine.set_option("short", 0)
    for L, N in ((45, 130), (171, 300), (1203, 140), (5000, 9)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 9 + L, prob=True, nan_rate=0.3)).to(dev)
        if K >= 3:
            run(f"walk posterior K{K} L{L}", lambda: engine.posterior(t, X, None, want_loglik=True))
            run(f"walk estep K{K} L{L}", lambda: engine.estep(t, X, None))
    engine.set_option("walk", -1)
    engine.set_option("short", 1)
    for L, N in ((45, 130), (400, 33), (900, 21)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 9 + L, prob=True, nan_rate=0.3)).to(dev)
        run(f"short-forced posterior K{K} L{L}", lambda: engine.posterior(t, X, None, want_loglik=True))
        run(f"short-forced estep K{K} L{L}", lambda: engine.estep(t, X, None))
    engine.set_option("short", -1)

# generic kernel: multi-channel and distance-binned
m, t = tables(3, arch="multi", hmm_n_channels=2)
X = torch.from_numpy(O.synth_reads(50, 300, 3, 3, prob=True, nan_rate=0.3, n_channels=2)).to(dev)
run("multi posterior", lambda: engine.posterior(t, X, None, want_loglik=True))
run("multi estep", lambda: engine.estep(t, X, None))
run("multi viterbi", lambda: engine.viterbi(t, X, None))
m, t = tables(2, arch="single_distance_binned", hmm_distance_bins=[1, 5, 10, 25, 50, 100])
coords = O.synth_coords(300, 4)
bins = m._bins_for(coords, 300, dev)
X = torch.from_numpy(O.synth_reads(50, 300, 2, 4, prob=False, nan_rate=0.3)).to(dev)
run("binned posterior", lambda: engine.posterior(t, X, bins, want_loglik=True))
run("binned estep", lambda: engine.estep(t, X, bins))
run("binned viterbi", lambda: engine.viterbi(t, X, bins))

# feature kernels
rng = np.random.default_rng(1)
for N, L in ((33, 700), (20, 1024), (7, 16)):
    states = torch.from_numpy(rng.integers(0, 3, size=(N, L)).astype(np.int8)).to(dev)
    coords_h = O.synth_coords(L, 2, span_factor=5) + 500
    full = np.arange(coords_h[0] - 7, coords_h[-1] + 9)
    left = torch.from_numpy(np.searchsorted(full, coords_h, side="left").astype(np.int32)).to(dev)
    right = torch.from_numpy(np.searchsorted(full, coords_h, side="right").astype(np.int32)).to(dev)
    ranges = [(6, 40), (40, 100), (100, 200), (200, np.inf)]
    kw = dict(states=states, post=None, target=1, threshold=0.0, coords=torch.from_numpy(coords_h).to(dev),
              grid_left=left, grid_right=right, n_grid=len(full), ranges=ranges, max_intervals=N * (L // 2 + 2))
    run(f"features dense L{L}", lambda: engine.call_features(**kw))
    run(f"features intervals L{L}", lambda: engine.call_features(want_dense=False, **kw))
    V = len(full)
    layer = torch.from_numpy((rng.random((N, V)) < 0.4).astype(np.uint8)).to(dev)
    fc = torch.from_numpy(full.astype(np.int64)).to(dev)
    run(f"merge_runs V{V}", lambda: engine.merge_runs(layer, fc, True, 10, ranges))
    st = torch.from_numpy(full[0] + rng.integers(0, 50, size=N).astype(np.float64)).to(dev)
    en = torch.from_numpy(full[-1] - rng.integers(0, 50, size=N).astype(np.float64)).to(dev)
    run(f"merge_chain V{V}", lambda: engine.merge_chain(layer, fc, True, 10, ranges, span_coords=fc, start=st, end=en))
    run(f"span_mask V{V}", lambda: engine.span_mask(N, V, dev, coords=fc, start=st, end=en))
# round 2: sequential kernels (forced on small batches), grouped launches, direct Viterbi, multi-channel / binned fast path
from smftools_b200 import grouped

engine.set_option("seq", 1)
for K in (2, 3, 4):
    m, t = tables(K)
    for L, N in ((16, 130), (170, 129), (1001, 70), (2048, 33)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 11 + L, prob=True, nan_rate=0.3)).to(dev)
        X8 = torch.where(torch.isnan(X), torch.tensor(255.0, device=dev), (X > 0.5).float()).to(torch.uint8)
        run(f"seq estep K{K} L{L}", lambda: engine.estep(t, X, None))
        run(f"seq estep-u8 K{K} L{L}", lambda: engine.estep(t, X8, None))
        if L % 4 == 0:
            run(f"seq posterior K{K} L{L}", lambda: engine.posterior(t, X, None, want_loglik=True))
            run(f"seq posterior-one K{K} L{L}", lambda: engine.posterior(t, X, None, gamma_state=1, want_states=False))
        if L % 16 == 0:
            run(f"seq posterior-u8 K{K} L{L}", lambda: engine.posterior(t, X8, None))
engine.set_option("seq", -1)
for K in (2, 3, 4):
    shapes = [(70, 16), (300, 170), (129, 1008), (1, 37), (40, 2003), (513, 333)]
    models = [tables(K)[0] for _ in shapes]
    obs = [S.CodedU8(torch.from_numpy(np.where(np.isnan(x), 255, x).astype(np.uint8)).to(dev))
           for x in (O.synth_reads(n, l, K, 50 + i, prob=False, nan_rate=0.3) for i, (n, l) in enumerate(shapes))]
    run(f"grouped fit K{K}", lambda: grouped.fit_groups(models, obs, max_iter=3, tol=0.0))
    run(f"grouped decode K{K}", lambda: grouped.decode_groups(models, obs, want_loglik=True))
    obsf = [torch.from_numpy(O.synth_reads(n, l, K, 80 + i, prob=True, nan_rate=0.3)).to(dev) for i, (n, l) in enumerate(shapes)]
    run(f"grouped fit f32 K{K}", lambda: grouped.fit_groups(models, obsf, max_iter=2, tol=0.0))
    run(f"grouped decode-one f32 K{K}", lambda: grouped.decode_groups(models, obsf, gamma_state=0, want_states=False))
engine.set_option("vit", 1)
for K in (2, 3, 4):
    m, t = tables(K)
    for L, N in ((4, 33), (64, 130), (1204, 70), (5000, 9)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 13 + L, prob=True, nan_rate=0.3)).to(dev)
        run(f"direct viterbi K{K} L{L}", lambda: engine.viterbi(t, X, None))
        if L % 16 == 0:
            X8 = torch.where(torch.isnan(X), torch.tensor(255.0, device=dev), (X > 0.5).float()).to(torch.uint8)
            run(f"direct viterbi-u8 K{K} L{L}", lambda: engine.viterbi(t, X8, None))
engine.set_option("vit", -1)
for K, C in ((2, 2), (3, 2), (2, 3), (3, 3)):
    m, t = tables(K, arch="multi", hmm_n_channels=C)
    for L, N in ((17, 40), (333, 50), (5000, 5), (9000, 3)):
        X = torch.from_numpy(O.synth_reads(N, L, K, 3 + L, prob=True, nan_rate=0.3, n_channels=C)).to(dev)
        run(f"multi fast posterior K{K} C{C} L{L}", lambda: engine.posterior(t, X, None, want_loglik=True))
        run(f"multi fast estep K{K} C{C} L{L}", lambda: engine.estep(t, X, None))
        X8 = torch.where(torch.isnan(X), torch.tensor(255.0, device=dev), (X > 0.5).float()).to(torch.uint8)
        run(f"multi fast posterior-u8 K{K} C{C} L{L}", lambda: engine.posterior(t, X8, None))
        run(f"multi fast estep-u8 K{K} C{C} L{L}", lambda: engine.estep(t, X8, None))
for K in (2, 3):
    m, t = tables(K, arch="single_distance_binned", hmm_distance_bins=[1, 5, 10, 25, 50, 100])
    for L, N in ((2, 9), (35, 100), (1000, 33), (9000, 3)):
        coords = O.synth_coords(L, 4)
        bins = m._bins_for(coords, L, dev)
        X = torch.from_numpy(O.synth_reads(N, L, K, 4 + L, prob=True, nan_rate=0.3)).to(dev)
        run(f"binned fast posterior K{K} L{L}", lambda: engine.posterior(t, X, bins, want_loglik=True))
        run(f"binned fast posterior-one K{K} L{L}", lambda: engine.posterior(t, X, bins, gamma_state=0))
print("all kernels ran")
