"""Randomised shape sweep: every posterior / E-step / Viterbi kernel family (bundled short reads,
one read per warp with all chunk lengths, two-kernel path, generic multi-channel / distance-binned,
segmented long reads) against the oracle on seeded random (K, L, N, dtype, output-set) draws."""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, assert_viterbi_matches, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

GAMMA_TOL = 1e-5


def _draws(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        K = int(rng.integers(2, 5))
        # lengths spread over the planner's regimes: bundled (<= ~400), one warp, several warps, long
        L = int(rng.choice([rng.integers(1, 60), rng.integers(60, 450), rng.integers(450, 1200),
                            rng.integers(1200, 4000), rng.integers(4000, 12000)]))
        N = int(rng.integers(1, 40)) if L > 4000 else int(rng.integers(1, 300))
        dtype = str(rng.choice(["f32", "f64", "u8"]))
        prob = bool(rng.integers(0, 2)) and dtype != "u8"
        out.append((i, K, L, N, dtype, prob, int(rng.integers(-1, K))))
    return out


@pytest.mark.parametrize("case", _draws(2024, 36), ids=lambda c: f"{c[0]}-K{c[1]}-L{c[2]}-N{c[3]}-{c[4]}")
def test_random_posterior_and_estep(case):
    from smftools_b200 import engine

    i, K, L, N, dtype, prob, gstate = case
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 500 + i, prob=prob, nan_rate=0.3)
    X = X32.astype(np.float64)
    if dtype == "f32":
        obs = torch.from_numpy(X32).to(dev)
    elif dtype == "f64":
        obs = torch.from_numpy(X).to(dev)
    else:
        obs = torch.from_numpy(np.where(np.isnan(X32), 255, X32).astype(np.uint8)).to(dev)
    g_ref, ll_ref = O.posterior(X, p)
    g, s, ll = engine.posterior(tables, obs, None, gamma_state=gstate, want_loglik=True)
    g = g.cpu().numpy()
    ref = g_ref if gstate < 0 else g_ref[:, :, gstate]
    assert g.shape == ref.shape
    assert np.abs(g - ref).max() <= GAMMA_TOL
    assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=2e-6)
    st = engine.estep(tables, obs, None).cpu().numpy()
    st_ref = O.stats_vector(O.estep(X, p, None))
    scale = np.maximum(np.abs(st_ref), 1.0)
    assert np.all(np.abs(st - st_ref) / scale <= 3e-6), (np.abs(st - st_ref) / scale).max()


@pytest.mark.parametrize("case", _draws(77, 10), ids=lambda c: f"{c[0]}-K{c[1]}-L{c[2]}-N{c[3]}-{c[4]}")
def test_random_viterbi(case):
    from smftools_b200 import engine

    i, K, L, N, dtype, prob, _ = case
    L = min(L, 3000)
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 900 + i, prob=prob, nan_rate=0.3)
    if dtype == "u8":
        obs = torch.from_numpy(np.where(np.isnan(X32), 255, X32).astype(np.uint8)).to(dev)
    elif dtype == "f64":
        obs = torch.from_numpy(X32.astype(np.float64)).to(dev)
    else:
        obs = torch.from_numpy(X32).to(dev)
    st = engine.viterbi(tables, obs, None).cpu().numpy().astype(np.int64)
    assert_viterbi_matches(X32.astype(np.float64), p, None, st)


@pytest.mark.parametrize("L,N", [(18500, 9), (21000, 5)])
def test_k2_reads_near_the_on_chip_limit(L, N):
    """K = 2 reads of 18-21 kb use the compact staging layout (one float per position, gamma expanded on
    the way out): posterior, states and the one-state / log-likelihood variants that use the regular one."""
    from smftools_b200 import engine

    dev = torch.device("cuda")
    K = 2
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 321 + L, prob=False, nan_rate=0.3)
    g_ref, ll_ref = O.posterior(X32.astype(np.float64), p)
    obs = torch.from_numpy(X32).to(dev)
    g, s, _ = engine.posterior(tables, obs, None)
    assert np.abs(g.cpu().numpy() - g_ref).max() <= GAMMA_TOL
    assert (g.sum(dim=2) - 1.0).abs().max().item() <= 1e-6
    assert_states_match_up_to_ties(s.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    g2, s2, ll = engine.posterior(tables, obs, None, want_loglik=True)
    assert np.abs(g2.cpu().numpy() - g_ref).max() <= GAMMA_TOL and torch.equal(s2, s)
    assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6)
    g1, _, _ = engine.posterior(tables, obs, None, gamma_state=1)
    assert np.abs(g1.cpu().numpy() - g_ref[:, :, 1]).max() <= GAMMA_TOL
