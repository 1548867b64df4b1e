"""GPU parity of the multi-channel variant of the fast-path kernel (fb2_kernel<..., C>, C = 2, 3; K = 2, 3):
posterior, states, per-read log-likelihood and E-step statistics against the oracle on [N, L, C] observations with
independent per-channel missingness (the union grid of build_multi_channel_union: a position often has a site of
only one channel), float and uint8-coded input, ragged / unaligned lengths, and the shapes it hands back to the
generic kernel (K = 4, C = 4, float64, reads beyond its reach)."""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, model_from_params, params_of_model
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu


def _multi_reads(N, L, K, C, seed, prob):
    X = O.synth_reads(N, L, K, seed, prob=prob, nan_rate=0.3, n_channels=C)
    rng = np.random.default_rng(seed + 1)
    # union-grid structure: each position carries a site of a random non-empty subset of the channels
    site = rng.random((L, C)) < 0.6
    site[np.arange(L), rng.integers(0, C, size=L)] = True
    X[:, ~site] = np.nan
    X[:, L // 3] = np.nan  # a position without any observed channel
    return X


CASES = [(2, 2, 64, 40, False), (2, 2, 1000, 33, True), (3, 2, 517, 50, True), (2, 3, 333, 70, False),
         (3, 3, 2048, 9, True), (2, 2, 5000, 12, False), (3, 2, 17, 100, False), (2, 3, 1, 5, True),
         (3, 3, 8000, 4, False), (2, 2, 10000, 6, True)]


@pytest.mark.parametrize("K,C,L,N,prob", CASES)
def test_multi_channel_posterior_and_estep_match_oracle(K, C, L, N, prob):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K, arch="multi", n_channels=C)
    tables = model_from_params(p, arch="multi")._host_tables()
    X = _multi_reads(N, L, K, C, 6100 + K * 10 + C + L % 53, prob)
    g_ref, ll_ref = O.posterior(X.astype(np.float64), p)
    ref_stats = O.stats_vector(O.estep(X.astype(np.float64), p, None))
    scale = np.maximum(np.abs(ref_stats), 1.0)
    variants = [("f32", torch.from_numpy(X.astype(np.float32)).to(dev))]
    if not prob:
        variants.append(("u8", torch.from_numpy(np.where(np.isnan(X), 255, X).astype(np.uint8)).to(dev)))
    for name, obs in variants:
        g, s, ll = engine.posterior(tables, obs, None, want_loglik=True)
        assert np.abs(g.cpu().numpy() - g_ref).max() <= 1e-5, name
        assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
        np.testing.assert_allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)
        g1, _, _ = engine.posterior(tables, obs, None, gamma_state=K - 1, want_states=False)
        assert np.abs(g1.cpu().numpy() - g_ref[:, :, K - 1]).max() <= 1e-5
        st = engine.estep(tables, obs, None).cpu().numpy()
        assert np.all(np.abs(st - ref_stats) / scale <= 3e-6), (name, (np.abs(st - ref_stats) / scale).max())
        assert np.array_equal(st, engine.estep(tables, obs, None).cpu().numpy()), "not deterministic"


@pytest.mark.parametrize("K,C,L,dt", [(4, 2, 300, np.float32), (2, 4, 300, np.float32), (2, 2, 300, np.float64),
                                      (2, 2, 12000, np.float32), (3, 3, 777, np.float64)])
def test_shapes_outside_the_multi_fast_path_use_the_generic_kernel(K, C, L, dt):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K, arch="multi", n_channels=C)
    tables = model_from_params(p, arch="multi")._host_tables()
    X = _multi_reads(7, L, K, C, 99 + K + C, True)
    g_ref, _ = O.posterior(X.astype(np.float64), p)
    g, _, _ = engine.posterior(tables, torch.from_numpy(X.astype(dt)).to(dev), None)
    assert np.abs(g.cpu().numpy() - g_ref).max() <= 1e-5
    st = engine.estep(tables, torch.from_numpy(X.astype(dt)).to(dev), None).cpu().numpy()
    ref = O.stats_vector(O.estep(X.astype(np.float64), p, None))
    assert np.all(np.abs(st - ref) / np.maximum(np.abs(ref), 1.0) <= 3e-6)


def test_multi_channel_fit_through_the_host_mirror():
    import smftools_b200 as S

    K, C = 2, 2
    X = _multi_reads(200, 400, K, C, 31, False)
    m = S.create_hmm({"hmm_n_states": K, "hmm_n_channels": C}, arch="multi")
    p0 = params_of_model(m)
    hist = m.fit(X, max_iter=5, tol=0.0)
    q, h_ref = O.fit(X.astype(np.float64), p0, max_iter=5, tol=0.0)
    np.testing.assert_allclose(hist, h_ref, rtol=1e-6)
    assert np.abs(params_of_model(m).emission - q.emission).max() <= 1e-4
