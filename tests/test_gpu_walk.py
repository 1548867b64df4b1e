"""GPU parity of the two-kernel forward-backward path (boundary walk + local sweeps, K >= 3).

The library picks this path for large batches when the caller provides the workspace
(smf_hmm_workspace_bytes(SMF_OP_POSTERIOR) > 0); the same batch without a workspace runs the
single-kernel path, so the two CUDA paths are also compared with each other on every read.
"""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

GAMMA_TOL = 1e-5
LL_RTOL = 1e-6
N_WALK = 8192  # kWalkMinReads in api.cu


@pytest.fixture(autouse=True)
def force_two_kernel_path():
    """The measured policy only picks the two-kernel path for some shapes; the parity tests force it."""
    from smftools_b200 import engine

    engine.set_option("walk", 1)
    yield
    engine.set_option("walk", -1)


def _reads(N, L, K, seed, prob):
    base = O.synth_reads(1024, L, K, seed, prob=prob, nan_rate=0.3)
    reps = (N + 1023) // 1024
    X = np.concatenate([np.roll(base, 7 * r, axis=0) if r % 2 == 0 else base[::-1] for r in range(reps)])[:N]
    # make the replicas differ: flip a sprinkling of observed sites
    rng = np.random.default_rng(seed + 1)
    flip = rng.random(X.shape) < 0.02
    if prob:
        X = np.where(flip & ~np.isnan(X), 1.0 - X, X)
    else:
        X = np.where(flip & ~np.isnan(X), 1.0 - X, X)
    return X.astype(np.float32)


@pytest.mark.parametrize("K,L,prob", [(3, 301, False), (3, 1000, True), (4, 257, True), (4, 1200, False), (3, 34, False)])
def test_walk_posterior_matches_oracle_and_single_kernel_path(K, L, prob):
    from smftools_b200 import _lib, engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    m = model_from_params(p)
    tables = m._host_tables()
    X32 = _reads(N_WALK, L, K, 40 + K, prob)
    obs = torch.from_numpy(X32).to(dev)
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, tables.ref, N_WALK, L))
    assert need > 0, "two-kernel path not offered for this batch"
    g_w, s_w, ll_w = engine.posterior(tables, obs, None, want_loglik=True)
    g_1, s_1, ll_1 = engine.posterior(tables, obs, None, want_loglik=True, use_workspace=False)
    assert torch.isfinite(g_w).all()
    assert (g_w - g_1).abs().max().item() <= GAMMA_TOL
    assert np.allclose(ll_w.cpu().numpy(), ll_1.cpu().numpy(), rtol=LL_RTOL)
    idx = np.r_[0:64, N_WALK - 64:N_WALK]
    sub = X32[idx].astype(np.float64)
    g_ref, ll_ref = O.posterior(sub, p)
    tidx = torch.from_numpy(idx).to(dev)
    assert np.abs(g_w[tidx].cpu().numpy() - g_ref).max() <= GAMMA_TOL
    assert np.allclose(ll_w[tidx].cpu().numpy(), ll_ref, rtol=LL_RTOL)
    assert_states_match_up_to_ties(s_w[tidx].cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    # single-state output + uint8 observations go through the same path
    if not prob:
        X8 = np.where(np.isnan(X32), 255, X32).astype(np.uint8)
        g_u8, s_u8, _ = engine.posterior(tables, torch.from_numpy(X8).to(dev), None, gamma_state=1)
        assert torch.equal(s_u8, s_w)
        assert (g_u8 - g_w[:, :, 1]).abs().max().item() <= 1e-6


@pytest.mark.parametrize("K,L", [(3, 257), (4, 300)])
def test_walk_estep_matches_oracle(K, L):
    from smftools_b200 import _lib, engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    m = model_from_params(p)
    tables = m._host_tables()
    X32 = _reads(N_WALK, L, K, 50 + K, True)
    obs = torch.from_numpy(X32).to(dev)
    S = engine.stats_len(tables)
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_ESTEP, tables.ref, N_WALK, L))
    assert need > 148 * 8 * S * 8, "E-step workspace does not include the boundary vectors"
    stats = engine.estep(tables, obs, None).cpu().numpy()
    ref = O.stats_vector(O.estep(X32.astype(np.float64), p, None))
    scale = np.maximum(np.abs(ref), 1.0)
    assert np.all(np.abs(stats - ref) / scale <= 2e-6), (np.abs(stats - ref) / scale).max()
    stats2 = engine.estep(tables, obs, None).cpu().numpy()
    assert np.array_equal(stats, stats2)


@pytest.mark.parametrize("K,L,N", [(2, 2, 50), (2, 45, 130), (3, 600, 77), (4, 1203, 41), (2, 5000, 19), (3, 5001, 13),
                                    # beyond the on-chip reach of every single-kernel variant
                                    (2, 30000, 12), (4, 12000, 20), (3, 23000, 10)])
def test_any_length_posterior_sweeps(K, L, N):
    """walk_kernel + post_sweep_kernel (option "walk" = 2): warp-independent slices, no read-length limit."""
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 70 + K + L % 97, prob=(L % 2 == 0), nan_rate=0.3)
    g_ref, ll_ref = O.posterior(X32.astype(np.float64), p)
    obs = torch.from_numpy(X32).to(dev)
    engine.set_option("walk", 2)
    engine.set_option("short", 0)
    try:
        g, s, ll = engine.posterior(tables, obs, None, want_loglik=True)
        assert np.abs(g.cpu().numpy() - g_ref).max() <= GAMMA_TOL
        assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=LL_RTOL, atol=2e-6)
        assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
        g1, s1, _ = engine.posterior(tables, obs, None, gamma_state=K - 1)
        assert torch.equal(s1, s) and (g1 - g[:, :, K - 1]).abs().max().item() <= 1e-6
        _, s2, _ = engine.posterior(tables, obs, None, want_gamma=False)
        assert torch.equal(s2, s)
        X8 = np.where(np.isnan(X32), 255, (X32 > 0.5)).astype(np.uint8)
        g8_ref, _ = O.posterior(np.where(X8 == 255, np.nan, X8).astype(np.float64), p)
        g8, _, _ = engine.posterior(tables, torch.from_numpy(X8).to(dev), None)
        assert np.abs(g8.cpu().numpy() - g8_ref).max() <= GAMMA_TOL
        g64, _, _ = engine.posterior(tables, torch.from_numpy(X32.astype(np.float64)).to(dev), None)
        assert (g64 - g).abs().max().item() <= 1e-6
    finally:
        engine.set_option("walk", -1)
        engine.set_option("short", -1)
