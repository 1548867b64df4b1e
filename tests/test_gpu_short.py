"""GPU parity of the short-read posterior kernel (several reads per warp, short_kernel.cuh) against the
oracle and against the one-read-per-warp path on the same inputs."""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

GAMMA_TOL = 1e-5
LL_RTOL = 1e-6


@pytest.fixture
def short_off():
    from smftools_b200 import engine

    def _set(v):
        engine.set_option("short", v)
    yield _set
    engine.set_option("short", -1)


@pytest.mark.parametrize("K", [2, 3, 4])
@pytest.mark.parametrize("L", [1, 2, 5, 17, 34, 64, 170, 257, 380])
def test_short_kernel_matches_oracle(K, L, short_off):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    N = 203  # not a multiple of any reads-per-warp count: the last bundle is partial
    p = O.synth_params(K)
    m = model_from_params(p)
    tables = m._host_tables()
    X32 = O.synth_reads(N, L, K, 60 + K + L, prob=(L % 2 == 0), nan_rate=0.3)
    X = X32.astype(np.float64)
    g_ref, ll_ref = O.posterior(X, p)
    obs = torch.from_numpy(X32).to(dev)
    short_off(1)
    g, s, ll = engine.posterior(tables, obs, None, want_loglik=True)
    assert np.abs(g.cpu().numpy() - g_ref).max() <= GAMMA_TOL
    assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=LL_RTOL, atol=2e-6)
    assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    # one-state output, float64 and (binary data) uint8 observations
    g1, s1, _ = engine.posterior(tables, torch.from_numpy(X).to(dev), None, gamma_state=K - 1)
    assert torch.equal(s1, s)
    assert (g1 - g[:, :, K - 1]).abs().max().item() <= 1e-6
    if L % 2 == 1:
        X8 = np.where(np.isnan(X32), 255, X32).astype(np.uint8)
        g8, s8, _ = engine.posterior(tables, torch.from_numpy(X8).to(dev), None)
        assert torch.equal(s8, s) and (g8 - g).abs().max().item() <= 1e-6
    # states only / posterior only
    _, s_only, _ = engine.posterior(tables, obs, None, want_gamma=False)
    assert_states_match_up_to_ties(s_only.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    g_only, _, _ = engine.posterior(tables, obs, None, want_states=False)
    assert (g_only - g).abs().max().item() <= 1e-6  # the planner may pick another chunk length
    # the one-read-per-warp path gives the same answer within fp32 rounding
    short_off(0)
    g0, s0, ll0 = engine.posterior(tables, obs, None, want_loglik=True)
    assert (g0 - g).abs().max().item() <= GAMMA_TOL
    assert np.allclose(ll0.cpu().numpy(), ll.cpu().numpy(), rtol=LL_RTOL, atol=2e-6)


def test_short_kernel_large_batch_unaligned_rows(short_off):
    """Many bundles per warp (prefetch pipeline) and rows whose byte offsets are not 16-byte multiples."""
    from smftools_b200 import engine

    dev = torch.device("cuda")
    K, L, N = 2, 171, 60_001
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    base = O.synth_reads(2048, L, K, 77, prob=False, nan_rate=0.3)
    reps = (N + 2047) // 2048
    X32 = np.concatenate([np.roll(base, 3 * r, axis=1) for r in range(reps)])[:N]
    obs = torch.from_numpy(X32).to(dev)
    short_off(1)
    g, s, ll = engine.posterior(tables, obs, None, want_loglik=True)
    short_off(0)
    g0, s0, ll0 = engine.posterior(tables, obs, None, want_loglik=True)
    assert (g - g0).abs().max().item() <= GAMMA_TOL
    assert np.allclose(ll.cpu().numpy(), ll0.cpu().numpy(), rtol=LL_RTOL)
    idx = np.r_[0:100, N - 100:N]
    g_ref, ll_ref = O.posterior(X32[idx].astype(np.float64), p)
    assert np.abs(g[torch.from_numpy(idx).to(dev)].cpu().numpy() - g_ref).max() <= GAMMA_TOL


@pytest.mark.parametrize("K", [2, 3])
@pytest.mark.parametrize("L", [1, 2, 5, 34, 100, 170, 257])
def test_short_estep_matches_oracle(K, L, short_off):
    """Baum-Welch statistics from the bundled short-read kernel vs the oracle and vs one read per warp."""
    from smftools_b200 import engine

    dev = torch.device("cuda")
    N = 1203
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 80 + K + L, prob=(L % 2 == 0), nan_rate=0.3)
    ref = O.stats_vector(O.estep(X32.astype(np.float64), p, None))
    scale = np.maximum(np.abs(ref), 1.0)
    obs = torch.from_numpy(X32).to(dev)
    short_off(1)
    st = engine.estep(tables, obs, None).cpu().numpy()
    assert np.all(np.abs(st - ref) / scale <= 2e-6), (np.abs(st - ref) / scale).max()
    assert np.array_equal(st, engine.estep(tables, obs, None).cpu().numpy())  # deterministic reduction
    if L % 2 == 1:
        X8 = np.where(np.isnan(X32), 255, X32).astype(np.uint8)
        st8 = engine.estep(tables, torch.from_numpy(X8).to(dev), None).cpu().numpy()
        assert np.all(np.abs(st8 - ref) / scale <= 2e-6)
    short_off(0)
    st0 = engine.estep(tables, obs, None).cpu().numpy()
    assert np.all(np.abs(st0 - st) / scale <= 4e-6)


def test_short_fit_matches_oracle(short_off):
    """A short-read fit (the reference's amplicon case) goes through the bundled E-step end to end."""
    K, L, N = 2, 170, 4000
    p0 = O.init_params(K, init_emission=[0.8, 0.2])
    X32 = O.synth_reads(N, L, K, 91, prob=False, nan_rate=0.3)
    p_ref, hist_ref = O.fit(X32.astype(np.float64), p0, None, max_iter=4, tol=0.0)
    m = model_from_params(p0)
    hist = m.fit(X32, None, max_iter=4, tol=0.0)
    assert np.allclose(hist, hist_ref, rtol=1e-6)
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    assert np.abs(sd["trans"] - p_ref.trans).max() <= 1e-4
    assert np.abs(sd["emission"].ravel() - p_ref.emission.ravel()).max() <= 1e-4
    assert np.abs(sd["start"] - p_ref.start).max() <= 1e-4
