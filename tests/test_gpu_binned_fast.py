"""GPU parity of the distance-binned variant of the fast-path kernel (fb2_kernel<..., BINNED>, K = 2, 3; posterior
decode): per-position transition matrices chosen by the distance bin of consecutive sites (HMM.py:2115-2188), against
the oracle, for float and uint8-coded input, ragged lengths, all output sets; K = 4 and float64 stay on the generic kernel."""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

BINS = [1, 5, 10, 25, 50, 100]


def _coords(L, seed):
    rng = np.random.default_rng(seed)
    gaps = rng.choice([1, 2, 4, 7, 12, 30, 60, 150, 400], size=L)
    return np.cumsum(gaps).astype(np.int64)


@pytest.mark.parametrize("K,L,N,prob", [(2, 64, 40, False), (2, 1000, 33, True), (3, 517, 50, True), (3, 2048, 9, False),
                                        (2, 5000, 12, False), (3, 17, 100, False), (2, 2, 5, True), (3, 8000, 4, True),
                                        (2, 10000, 6, True), (2, 35, 300, False)])
def test_binned_posterior_matches_oracle(K, L, N, prob):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K, arch="single_distance_binned", distance_bins=BINS)
    m = model_from_params(p, arch="single_distance_binned")
    tables = m._host_tables()
    coords = _coords(L, 5 + L)
    bins = m._bins_for(coords, L, dev)
    X = O.synth_reads(N, L, K, 7300 + K + L % 59, prob=prob, nan_rate=0.3)
    g_ref, ll_ref = O.posterior(X.astype(np.float64), p, coords)
    variants = [torch.from_numpy(X.astype(np.float32)).to(dev)]
    if not prob:
        variants.append(torch.from_numpy(np.where(np.isnan(X), 255, X).astype(np.uint8)).to(dev))
    for obs in variants:
        g, s, ll = engine.posterior(tables, obs, bins, want_loglik=True)
        assert np.abs(g.cpu().numpy() - g_ref).max() <= 1e-5
        assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
        np.testing.assert_allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)
        g1, s1, _ = engine.posterior(tables, obs, bins, gamma_state=0, want_states=False)
        assert s1 is None and np.abs(g1.cpu().numpy() - g_ref[:, :, 0]).max() <= 1e-5
        g2, s2, _ = engine.posterior(tables, obs, bins, want_gamma=False)
        assert g2 is None and torch.equal(s2, s)
        g3, _, _ = engine.posterior(tables, obs, bins)
        g4, _, _ = engine.posterior(tables, obs, bins)
        assert torch.equal(g3, g4), "not deterministic"
        assert (g3 - g).abs().max().item() <= 2e-6  # the log-likelihood variant rescales with exact reciprocals


@pytest.mark.parametrize("K,dt", [(4, np.float32), (2, np.float64)])
def test_binned_shapes_outside_the_fast_path(K, dt):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K, arch="single_distance_binned", distance_bins=BINS)
    m = model_from_params(p, arch="single_distance_binned")
    coords = _coords(300, 9)
    X = O.synth_reads(11, 300, K, 12, prob=True, nan_rate=0.3)
    g_ref, _ = O.posterior(X.astype(np.float64), p, coords)
    g, _, _ = engine.posterior(m._host_tables(), torch.from_numpy(X.astype(dt)).to(dev), m._bins_for(coords, 300, dev))
    assert np.abs(g.cpu().numpy() - g_ref).max() <= 1e-5


def test_binned_decode_and_fit_through_the_host_mirror():
    import smftools_b200 as S
    from helpers import params_of_model

    K = 2
    coords = _coords(400, 3)
    X = O.synth_reads(150, 400, K, 77, prob=False, nan_rate=0.3)
    m = S.create_hmm({"hmm_n_states": K, "hmm_distance_bins": BINS}, arch="single_distance_binned")
    p0 = params_of_model(m)
    hist = m.fit(X, coords, max_iter=4, tol=0.0)
    q, h_ref = O.fit(X.astype(np.float64), p0, coords, max_iter=4, tol=0.0)
    np.testing.assert_allclose(hist, h_ref, rtol=1e-6)
    st, g = m.decode(X, coords)
    g_ref, _ = O.posterior(X.astype(np.float64), params_of_model(m), coords)
    assert np.abs(g - g_ref).max() <= 1e-5


@pytest.mark.parametrize("L,N,prob", [(64, 40, False), (1000, 33, True), (5000, 12, False), (35, 300, False), (2, 5, True),
                                      (10000, 6, True), (517, 700, True)])
def test_binned_k2_estep_matches_oracle(L, N, prob):
    """fb2_kernel<2, ..., BINNED, STATS>: per-bin transition statistics kept per thread in shared memory."""
    from smftools_b200 import engine

    dev = torch.device("cuda")
    K = 2
    p = O.synth_params(K, arch="single_distance_binned", distance_bins=BINS)
    m = model_from_params(p, arch="single_distance_binned")
    tables = m._host_tables()
    coords = _coords(L, 11 + L)
    bins = m._bins_for(coords, L, dev)
    X = O.synth_reads(N, L, K, 8300 + L % 61, prob=prob, nan_rate=0.3)
    ref = O.stats_vector(O.estep(X.astype(np.float64), p, coords))
    scale = np.maximum(np.abs(ref), 1.0)
    variants = [torch.from_numpy(X.astype(np.float32)).to(dev)]
    if not prob:
        variants.append(torch.from_numpy(np.where(np.isnan(X), 255, X).astype(np.uint8)).to(dev))
    for obs in variants:
        st = engine.estep(tables, obs, bins).cpu().numpy()
        assert st.shape == ref.shape
        assert np.all(np.abs(st - ref) / scale <= 3e-6), (np.abs(st - ref) / scale).max()
        assert np.array_equal(st, engine.estep(tables, obs, bins).cpu().numpy()), "not deterministic"
        engine.set_option("generic", 1)
        try:
            st_g = engine.estep(tables, obs, bins).cpu().numpy()
        finally:
            engine.set_option("generic", 0)
        assert np.all(np.abs(st - st_g) / scale <= 5e-6)
