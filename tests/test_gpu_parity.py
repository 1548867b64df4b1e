"""GPU parity: CUDA kernels (through the host mirror / C-ABI) vs golden reference outputs and the oracle.

Tolerances (BASELINE.json north_star): posterior <= 1e-5 abs (fp32 kernels vs fp64 reference);
Viterbi paths bit-exact except reads with competing log-probabilities within 1e-6;
fitted parameters <= 1e-4 after the same number of iterations; LL history relative <= 1e-6.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, params_from_golden
from helpers import arch_of, assert_states_match_up_to_ties, assert_viterbi_matches, model_from_params, params_of_model
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

CORE = [n for n in golden_names() if not n.startswith(("annotate", "inputs_", "reduce_"))]
GAMMA_TOL = 1e-5
PARAM_TOL = 1e-4
LL_RTOL = 1e-6


@pytest.mark.parametrize("name", CORE)
def test_decode_marginal_against_reference_golden(name):
    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p, arch_of(g))
    states, gamma = m.decode(g["X"].astype(np.float64), g["coords"], decode="marginal")
    assert states.dtype == np.int64 and gamma.dtype == np.float64
    assert gamma.shape == g["gamma"].shape
    assert np.abs(gamma - g["gamma"]).max() <= GAMMA_TOL
    assert_states_match_up_to_ties(states, g["states_marginal"].astype(np.int64), g["gamma"])


@pytest.mark.parametrize("name", CORE)
def test_decode_viterbi_against_reference_golden(name):
    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p, arch_of(g))
    states, gamma = m.decode(g["X"].astype(np.float64), g["coords"], decode="viterbi")
    assert np.abs(gamma - g["gamma"]).max() <= GAMMA_TOL
    X = g["X"].astype(np.float64)
    ndiff = assert_viterbi_matches(X, p, g["coords"], states)
    # golden itself (reference output) must agree wherever the oracle does
    same_ref = (states == g["states_viterbi"]).all(axis=1)
    assert same_ref.sum() >= len(same_ref) - ndiff - 1


@pytest.mark.parametrize("name", CORE)
def test_loglik_against_reference_golden(name):
    from smftools_b200 import engine

    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p, arch_of(g))
    dev = torch.device("cuda")
    tables = m._host_tables()
    obs = torch.from_numpy(g["X"]).to(dev)
    bins = m._bins_for(g["coords"], obs.shape[1], dev)
    _, _, ll = engine.posterior(tables, obs, bins, want_gamma=False, want_states=False, want_loglik=True)
    ll = ll.cpu().numpy()
    assert np.allclose(ll, g["loglik"], rtol=LL_RTOL, atol=1e-5)


@pytest.mark.parametrize("name", CORE)
def test_fit_against_reference_golden(name):
    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p, arch_of(g))
    iters = len(g["fit_hist"])
    hist = m.fit(g["X"].astype(np.float64), g["coords"], max_iter=iters, tol=0.0)
    assert len(hist) == iters
    assert np.allclose(hist, g["fit_hist"], rtol=LL_RTOL, atol=1e-5)
    sd = m.state_dict()
    for key in ("start", "trans", "emission", "trans_by_bin"):
        if f"p1_{key}" in g:
            assert sd[key].dtype == torch.float64
            assert np.abs(sd[key].cpu().numpy() - g[f"p1_{key}"]).max() <= PARAM_TOL, key
    hist2 = m.adapt_emissions(g["X"].astype(np.float64), g["coords"], iters=len(g["adapt_hist"]))
    assert np.allclose(hist2, g["adapt_hist"], rtol=LL_RTOL, atol=1e-5)
    sd = m.state_dict()
    for key in ("start", "trans", "emission", "trans_by_bin"):
        if f"p2_{key}" in g:
            assert np.abs(sd[key].cpu().numpy() - g[f"p2_{key}"]).max() <= PARAM_TOL, key


CASES = [
    # arch, K, C, N, L, prob
    ("single", 2, 1, 300, 1000, False),
    ("single", 2, 1, 64, 5000, True),
    ("single", 3, 1, 64, 4097, True),
    ("single", 4, 1, 48, 3000, False),
    ("multi", 2, 2, 96, 700, False),
    ("multi", 3, 4, 40, 500, True),
    ("single_distance_binned", 2, 1, 96, 900, False),
    ("single_distance_binned", 4, 1, 40, 800, True),
    ("single", 2, 1, 1000, 17, False),   # many short reads (several reads per CTA)
    ("single", 3, 1, 3, 9000, False),    # few long reads
]


@pytest.mark.parametrize("arch,K,C,N,L,prob", CASES)
def test_kernels_against_oracle(arch, K, C, N, L, prob):
    seed = 100 + K * 7 + C * 3 + L % 97
    X32 = O.synth_reads(N, L, K, seed, prob=prob, nan_rate=0.3, n_channels=C)
    X = X32.astype(np.float64)
    coords = O.synth_coords(L, seed) if arch == "single_distance_binned" else np.arange(L)
    p = O.synth_params(K, arch=arch, n_channels=C)
    m = model_from_params(p, arch)
    gamma_ref, ll_ref = O.posterior(X, p, coords)
    # float32 input path
    st, gamma = m.decode(X32, coords, decode="marginal", compact=True)
    assert gamma.dtype == np.float32 and st.dtype == np.int8
    assert np.abs(gamma.astype(np.float64) - gamma_ref).max() <= GAMMA_TOL
    assert_states_match_up_to_ties(st.astype(np.int64), gamma_ref.argmax(axis=2), gamma_ref)
    # float64 input path
    st64, gamma64 = m.decode(X, coords, decode="marginal", compact=True)
    # same values after conversion; the planner may pick another kernel variant / chunk length per input
    # dtype, so the two results agree to fp32 rounding rather than bit for bit
    assert np.abs(gamma64.astype(np.float64) - gamma_ref).max() <= GAMMA_TOL
    assert np.abs(gamma64 - gamma).max() <= 2e-6
    assert_states_match_up_to_ties(st64.astype(np.int64), gamma_ref.argmax(axis=2), gamma_ref)
    # Viterbi
    stv, _ = m.decode(X32, coords, decode="viterbi", compact=True)
    assert_viterbi_matches(X, p, coords, stv.astype(np.int64))
    # one E-step: statistics against the oracle
    from smftools_b200 import engine

    dev = torch.device("cuda")
    tables = m._host_tables()
    obs = torch.from_numpy(X32).to(dev)
    bins = m._bins_for(coords, L, dev)
    stats = engine.estep(tables, obs, bins).cpu().numpy()
    ref = O.stats_vector(O.estep(X, p, coords))
    scale = np.maximum(np.abs(ref), 1.0)
    assert np.all(np.abs(stats - ref) / scale <= 2e-6), np.abs(stats - ref).max()
    # determinism of the fixed-order reduction
    stats2 = engine.estep(tables, obs, bins).cpu().numpy()
    assert np.array_equal(stats, stats2)


def test_u8_binary_input_equals_float_input():
    N, L, K = 200, 1203, 2
    X32 = O.synth_reads(N, L, K, 5, prob=False, nan_rate=0.3)
    code = np.where(np.isnan(X32), 255, X32).astype(np.uint8)
    p = O.synth_params(K)
    m = model_from_params(p)
    st_f, g_f = m.decode(X32, decode="marginal", compact=True)
    import smftools_b200 as S

    st_u, g_u = m.decode(S.CodedU8(code), decode="marginal", compact=True)
    assert np.array_equal(g_f, g_u) and np.array_equal(st_f, st_u)
    sv_f, _ = m.decode(X32, decode="viterbi", compact=True)
    sv_u, _ = m.decode(S.CodedU8(code), decode="viterbi", compact=True)
    assert np.array_equal(sv_f, sv_u)
    hist_f = m.fit(X32, max_iter=2, tol=0.0)
    m2 = model_from_params(p)
    hist_u = m2.fit(S.CodedU8(code), max_iter=2, tol=0.0)
    assert np.allclose(hist_f, hist_u, rtol=1e-9)


def test_plain_uint8_input_keeps_its_numeric_meaning():
    """A plain uint8 matrix means what ``np.asarray(X, dtype=float)`` means in the reference (HMM.py:694):
    the byte value is the observation.  Only ``CodedU8`` opts into the 0 / 1 / missing code."""
    N, L, K = 64, 301, 2
    X32 = np.nan_to_num(O.synth_reads(N, L, K, 6, prob=False, nan_rate=0.0), nan=0.0)
    Xb = X32.astype(np.uint8)
    Xb[:, ::7] = 2  # a value that is neither 0 nor 1: numeric 2.0, not "missing"
    p = O.synth_params(K)
    m = model_from_params(p)
    st_u, g_u = m.decode(Xb, decode="marginal")
    g_ref, _ = O.posterior(Xb.astype(np.float64), p)
    assert np.abs(g_u - g_ref).max() <= GAMMA_TOL
    st_f, g_f = m.decode(Xb.astype(np.float32), decode="marginal")  # bytes are widened to fp32 on the device
    assert np.array_equal(g_u, g_f) and np.array_equal(st_u, st_f)


def test_reference_property_tests_hold_for_the_cuda_models():
    """The reference's own numeric tests (tests/unit/test_hmm_fit_em_log_likelihood.py:32-103),
    run against the CUDA classes registered under the same arch names."""
    import smftools_b200 as S

    def data(rng, n_reads=40, n_positions=24):
        X = np.zeros((n_reads, n_positions))
        hr, hp = n_reads // 2, n_positions // 2
        X[:hr, :hp] = 1
        X[hr:, hp:] = 1
        noise = rng.random(X.shape) < 0.05
        return np.logical_xor(X.astype(bool), noise).astype(float)

    cfg = {"hmm_n_states": 2, "hmm_distance_bins": [1, 5, 10, 25, 50, 100]}
    for arch in ("single", "single_distance_binned"):
        X = data(np.random.default_rng(0))
        m = S.create_hmm(cfg, arch=arch)
        hist = m.fit(X, np.arange(X.shape[1]), max_iter=25, tol=0.0)
        assert len(hist) == 25 and abs(hist[0]) > 1.0 and max(abs(v) for v in hist) > 1.0
    X = data(np.random.default_rng(1))
    hist = S.create_hmm(cfg, arch="single").fit(X, np.arange(X.shape[1]), max_iter=25, tol=0.0)
    assert (np.diff(np.asarray(hist)) >= -1e-6).all()
    X = data(np.random.default_rng(3))
    hist = S.create_hmm(cfg, arch="single").fit(X, np.arange(X.shape[1]), max_iter=200, tol=1e-5)
    assert len(hist) < 200
    xs = data(np.random.default_rng(2), 30, 16)
    X = np.stack([xs, 1.0 - xs], axis=-1)
    hist = S.create_hmm(cfg, arch="multi").fit(X, np.arange(X.shape[1]), max_iter=20, tol=0.0)
    assert len(hist) == 20 and abs(hist[0]) > 1.0


def test_full_size_properties_c2():
    """BASELINE config #2 shape (100k x 5 kb, K=2, 30 % NaN) -- size-independent properties:
    posteriors are distributions, states are their arg-max, a 512-read prefix matches the oracle,
    E-step mass balances (sum start = N, sum emit_den = #observed, sum xi = #observed pairs), and
    the E-step log-likelihood equals the sum of the per-read log-likelihoods."""
    from smftools_b200 import engine

    N, L, K = 100_000, 5_000, 2
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(2234)
    # cheap device generator: sticky two-state pattern by blocks + noise + NaNs
    blocks = (torch.rand((N, L // 50), device=dev, generator=gen) < 0.4).repeat_interleave(50, dim=1)
    pr = torch.where(blocks, torch.tensor(0.9, device=dev), torch.tensor(0.07, device=dev))
    obs = (torch.rand((N, L), device=dev, generator=gen) < pr).float()
    obs[torch.rand((N, L), device=dev, generator=gen) < 0.3] = float("nan")
    p = O.synth_params(K)
    m = model_from_params(p)
    tables = m._host_tables()
    gamma, states, ll = engine.posterior(tables, obs, None, want_loglik=True)
    assert torch.isfinite(gamma).all()
    assert (gamma.sum(dim=2) - 1.0).abs().max().item() < 1e-5
    assert torch.equal(states.long(), gamma.argmax(dim=2)) or (
        (states.long() != gamma.argmax(dim=2)).float().mean().item() < 1e-6)
    sub = obs[:512].cpu().numpy().astype(np.float64)
    g_ref, ll_ref = O.posterior(sub, p)
    assert np.abs(gamma[:512].cpu().numpy() - g_ref).max() <= GAMMA_TOL
    assert np.allclose(ll[:512].cpu().numpy(), ll_ref, rtol=LL_RTOL)
    stats = engine.estep(tables, obs, None).cpu().numpy()
    observed = ~torch.isnan(obs)
    n_obs = int(observed.sum().item())
    n_pairs = int((observed[:, 1:] & observed[:, :-1]).sum().item())
    assert abs(stats[:K].sum() - N) / N < 1e-6
    assert abs(stats[K:K + K * K].sum() - n_pairs) / n_pairs < 1e-6
    assert abs(stats[K + K * K + K:K + K * K + 2 * K].sum() - n_obs) / n_obs < 1e-6
    assert abs(stats[-1] - ll.sum().item()) / abs(stats[-1]) < 1e-9
    # Viterbi on the full matrix: a 512-read prefix is bit-exact against the oracle
    st_v = engine.viterbi(tables, obs, None)
    assert_viterbi_matches(sub, p, None, st_v[:512].cpu().numpy().astype(np.int64))
