"""GPU parity at the BASELINE.json shapes, through the kernel the planner picks at that size.

The random-shape sweeps stay below the read counts / lengths at which the library switches to its
large-batch paths (two-kernel E-step from 8192 reads, vectorised interval calling on 16-byte-aligned
rows, Viterbi tiles at 10 kb), so these cases pin exactly those.  The oracle stays fast through exact
additivity: a batch of R replicas of a small read set has R times its sufficient statistics.
"""
import numpy as np
import pytest
import torch

from helpers import model_from_params, params_of_model
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu


def _tiled(base: np.ndarray, reps: int) -> np.ndarray:
    return np.ascontiguousarray(np.tile(base, (reps, 1)))


def test_two_kernel_estep_k4_8kb_8192_reads_vs_oracle_and_single_kernel():
    """BASELINE configs[3] kernel shape: K = 4, L = 8000, >= 8192 reads -> walk_kernel + estep_sweep_kernel
    (api.cu walk_eligible).  Checked against the oracle and against the single-kernel path."""
    from smftools_b200 import _lib, engine

    K, L, base_n, reps = 4, 8000, 64, 128
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    base = O.synth_reads(base_n, L, K, 404, prob=False, nan_rate=0.3)
    N = base_n * reps
    S = engine.stats_len(tables)
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_ESTEP, tables.ref, N, L))
    assert need > 148 * 8 * S * 8, "the policy did not pick the two-kernel path for the configs[3] shape"
    ref = O.stats_vector(O.estep(base.astype(np.float64), p, None)) * reps
    scale = np.maximum(np.abs(ref), 1.0)
    for dtype in ("f32", "u8"):
        if dtype == "f32":
            obs = torch.from_numpy(_tiled(base, reps)).to(dev)
        else:
            obs = torch.from_numpy(_tiled(np.where(np.isnan(base), 255, base).astype(np.uint8), reps)).to(dev)
        st = engine.estep(tables, obs, None).cpu().numpy()
        assert np.all(np.abs(st - ref) / scale <= 2e-6), (dtype, (np.abs(st - ref) / scale).max())
        assert np.array_equal(st, engine.estep(tables, obs, None).cpu().numpy()), "E-step is not deterministic"
        engine.set_option("walk", 0)
        try:
            st1 = engine.estep(tables, obs, None).cpu().numpy()
        finally:
            engine.set_option("walk", -1)
        assert np.all(np.abs(st1 - st) / scale <= 4e-6), (dtype, (np.abs(st1 - st) / scale).max())


def test_viterbi_k3_probabilistic_10kb_and_vector_interval_kernel():
    """BASELINE configs[2] shape: K = 3, probabilistic calls, L = 10 000 (a multiple of 16, so the
    interval records come from call_intervals_vec_kernel)."""
    from smftools_b200 import engine

    K, L, N = 3, 10_000, 48
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 505, prob=True, nan_rate=0.3)
    X32 = np.round(X32 * 255.0) / 255.0  # ML-tag resolution, like the bench generator
    obs = torch.from_numpy(X32.astype(np.float32)).to(dev)
    ref, margin = O.viterbi(X32.astype(np.float32).astype(np.float64), p, return_margins=True)
    sv = engine.viterbi(tables, obs, None)
    sv_h = sv.cpu().numpy().astype(np.int64)
    same = (sv_h == ref).all(axis=1)
    assert np.all(same | (margin < 1e-6)), f"{int((~same).sum())} reads differ outside the 1e-6 near-tie rule"
    assert same.sum() >= N - 4
    # footprint intervals from the decoded states: compacted records == oracle runs -> size class -> span fill
    coords = np.sort(np.random.default_rng(5).choice(np.arange(4 * L), size=L, replace=False)).astype(np.int64)
    full = np.arange(4 * L, dtype=np.int64)
    ranges = [(6, 40), (40, 100), (100, 200), (200, float("inf"))]
    left = np.searchsorted(full, coords, side="left").astype(np.int32)
    right = np.searchsorted(full, coords, side="right").astype(np.int32)
    assert sv.data_ptr() % 16 == 0 and L % 16 == 0
    for target in (1, 2):
        _, _, iv = engine.call_features(states=sv, post=None, target=target, threshold=0.0,
                                        coords=torch.from_numpy(coords).to(dev), grid_left=torch.from_numpy(left).to(dev),
                                        grid_right=torch.from_numpy(right).to(dev), n_grid=4 * L, ranges=ranges,
                                        want_dense=False, max_intervals=N * (L // 2 + 1))
        got = sorted(map(tuple, iv.cpu().numpy().tolist()))
        want = []
        for i in range(N):
            for a, b in O.runs_from_bool(sv_h[i] == target):
                c = O.pick_class(ranges, int(coords[b - 1]) - int(coords[a]) + 1)
                if c >= 0:
                    want.append((i, int(left[a]), int(right[b - 1]), c))
        assert got == sorted(want)
    # too small a record buffer is an error, not a silent truncation
    with pytest.raises(ValueError):
        engine.call_features(states=sv, post=None, target=1, threshold=0.0, coords=torch.from_numpy(coords).to(dev),
                             grid_left=torch.from_numpy(left).to(dev), grid_right=torch.from_numpy(right).to(dev),
                             n_grid=4 * L, ranges=ranges, want_dense=False, max_intervals=3)


@pytest.mark.parametrize("L,reps", [(2000, 128), (300, 16)])
def test_twenty_iteration_k4_fit_matches_oracle(L, reps):
    """'Fitted parameters within 1e-4 after the same iteration count' for configs[3]'s model: 20 Baum-Welch
    iterations, K = 4, through model.fit -- on 8192 reads (two-kernel E-step) and on 1024 reads (bundled
    short-read E-step), fp32-accumulated / fixed-order-reduced statistics in every iteration."""
    import smftools_b200 as S

    K, base_n, iters = 4, 64, 20
    base = O.synth_reads(base_n, L, K, 606, prob=False, nan_rate=0.3)
    X = _tiled(base, reps)
    init = [0.85, 0.08, 0.35, 0.60]
    m = S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": init}, arch="single")
    hist = m.fit(X, max_iter=iters, tol=0.0)
    # oracle on the distinct reads, statistics scaled by the replica count (exact additivity)
    p = O.init_params(K, init)
    hist_ref = []
    b64 = base.astype(np.float64)
    for _ in range(iters):
        st = O.estep(b64, p, None)
        st = {k: np.asarray(v) * reps for k, v in st.items()}
        hist_ref.append(float(O.stats_vector(st)[-1]))
        p = O.mstep(p, st)
    assert len(hist) == iters
    np.testing.assert_allclose(hist, hist_ref, rtol=1e-6)
    q = params_of_model(m)
    assert np.abs(q.start - p.start).max() <= 1e-4
    assert np.abs(q.trans - p.trans).max() <= 1e-4
    assert np.abs(q.emission.reshape(-1) - p.emission.reshape(-1)).max() <= 1e-4
