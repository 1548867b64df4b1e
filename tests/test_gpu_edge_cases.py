"""Edge cases through the C-ABI / host mirror: empty and tiny inputs, all-missing reads, ragged
alignments (row starts that are not 16-byte aligned), single-state posterior output, generic
kernel fallback for long reads, loud errors for reads that do not fit."""
import numpy as np
import pytest
import torch

from helpers import assert_states_match_up_to_ties, assert_viterbi_matches, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_empty_inputs():
    m = model_from_params(O.synth_params(2))
    st, g = m.decode(np.zeros((0, 7)))
    assert st.shape == (0, 7) and g.shape == (0, 7, 2)
    st, g = m.decode(np.zeros((3, 0)))
    assert st.shape == (3, 0) and g.shape == (3, 0, 2)
    from smftools_b200 import engine

    t = m._host_tables()
    s = engine.estep(t, torch.zeros((0, 5), device=DEV)).cpu().numpy()
    assert s.shape == (11,) and not s.any()


@pytest.mark.parametrize("L", [1, 2, 16, 17, 18, 33, 34, 543, 544, 545])
def test_ragged_lengths_and_unaligned_rows(L):
    """Lengths around the chunk (17) and warp-slice (544) boundaries; N odd so that most rows start
    off a 16-byte boundary; inputs sliced at an odd element offset."""
    K = 3
    N = 37
    X32 = O.synth_reads(N + 1, L, K, 3000 + L, prob=True, nan_rate=0.25, truncate=False)
    p = O.synth_params(K)
    m = model_from_params(p)
    from smftools_b200 import engine

    t = m._host_tables()
    flat = torch.from_numpy(X32.reshape(-1)).to(DEV)
    obs = flat[L:].reshape(N, L)  # storage offset L elements: arbitrary alignment
    assert obs.is_contiguous()
    Xs = X32[1:].astype(np.float64)
    g_ref, ll_ref = O.posterior(Xs, p)
    big = torch.empty((N * L * K + 3,), dtype=torch.float32, device=DEV)
    gamma_out = big[3:].reshape(N, L, K) if K != 2 and K != 4 else big[:N * L * K].reshape(N, L, K)
    gamma, states, ll = engine.posterior(t, obs, None, want_loglik=True, out_gamma=gamma_out)
    assert np.abs(gamma.cpu().numpy() - g_ref).max() <= 1e-5
    assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-5)
    assert_states_match_up_to_ties(states.cpu().numpy().astype(np.int64), g_ref.argmax(2), g_ref)
    # single-state posterior output
    for k in range(K):
        g1, _, _ = engine.posterior(t, obs, None, gamma_state=k, want_states=False)
        assert g1.shape == (N, L)
        assert np.abs(g1.cpu().numpy() - g_ref[:, :, k]).max() <= 1e-5
    # annotate-mode specialisation (single state + states)
    g1, st1, _ = engine.posterior(t, obs, None, gamma_state=1, want_states=True)
    assert torch.equal(st1, states) and np.abs(g1.cpu().numpy() - g_ref[:, :, 1]).max() <= 1e-5
    sv = engine.viterbi(t, obs, None).cpu().numpy().astype(np.int64)
    assert_viterbi_matches(Xs, p, None, sv)
    stats = engine.estep(t, obs, None).cpu().numpy()
    ref = O.stats_vector(O.estep(Xs, p))
    assert np.all(np.abs(stats - ref) / np.maximum(np.abs(ref), 1.0) <= 2e-6)


def test_all_missing_and_fully_observed_reads():
    K = 2
    p = O.synth_params(K)
    m = model_from_params(p)
    X = O.synth_reads(6, 300, K, 1, nan_rate=0.0, truncate=False).astype(np.float64)
    X[0, :] = np.nan          # nothing observed: posterior = prior dynamics only
    X[1, :150] = np.nan       # leading block missing
    X[2, 150:] = np.nan       # trailing block missing
    X[3, ::2] = np.nan        # every other site
    st, g = m.decode(X)
    g_ref, _ = O.posterior(X, p)
    assert np.isfinite(g).all() and np.abs(g - g_ref).max() <= 1e-5
    sv, _ = m.decode(X, decode="viterbi")
    assert_viterbi_matches(X, p, None, sv)
    hist = m.fit(X, max_iter=2, tol=0.0)
    _, hist_ref = O.fit(X, p, max_iter=2, tol=0.0)
    assert np.allclose(hist, hist_ref, rtol=1e-6)


def test_extreme_parameters_stay_finite():
    """Near-deterministic transitions / emissions at the eps clamps must not under- or overflow."""
    K = 2
    p = O.Params([0.5, 0.5], [[1 - 1e-7, 1e-7], [1e-7, 1 - 1e-7]], [1.0, 0.0]).normalize()
    m = model_from_params(p)
    X = O.synth_reads(40, 700, K, 8, nan_rate=0.3).astype(np.float64)
    st, g = m.decode(X)
    g_ref, _ = O.posterior(X, p)
    assert np.isfinite(g).all()
    assert np.abs(g - g_ref).max() <= 1e-5


@pytest.mark.parametrize("arch,K,L", [("single", 2, 12000), ("single", 3, 11000), ("single", 4, 9000),
                                      ("single", 2, 20000), ("single", 3, 41003), ("single", 4, 30001),
                                      ("single_distance_binned", 2, 25000), ("multi", 2, 19000)])
def test_long_reads_use_the_generic_kernel(arch, K, L):
    """Beyond the fast path's thread budget the generic kernel takes over; beyond its on-chip
    staging it walks the read in segments (two passes).  Same results either way."""
    from smftools_b200 import engine

    N = 3
    C = 2 if arch == "multi" else 1
    X32 = O.synth_reads(N, L, K, 55 + K, prob=(K == 3), n_channels=C)
    coords = O.synth_coords(L, 9) if arch == "single_distance_binned" else np.arange(L)
    p = O.synth_params(K, arch=arch, n_channels=C)
    m = model_from_params(p, arch)
    X = X32.astype(np.float64)
    g_ref, ll_ref = O.posterior(X, p, coords)
    st, g = m.decode(X32, coords, compact=True)
    assert np.abs(g - g_ref).max() <= 1e-5
    assert_states_match_up_to_ties(st.astype(np.int64), g_ref.argmax(2), g_ref)
    t = m._host_tables()
    obs = torch.from_numpy(X32).to(DEV)
    bins = m._bins_for(coords, L, torch.device(DEV))
    _, _, ll = engine.posterior(t, obs, bins, want_gamma=False, want_states=False, want_loglik=True)
    assert np.allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6)
    g1, _, _ = engine.posterior(t, obs, bins, gamma_state=K - 1, want_states=False)
    assert np.abs(g1.cpu().numpy() - g_ref[:, :, K - 1]).max() <= 1e-5
    stats = engine.estep(t, obs, bins).cpu().numpy()
    ref = O.stats_vector(O.estep(X, p, coords))
    assert np.all(np.abs(stats - ref) / np.maximum(np.abs(ref), 1.0) <= 2e-6)


def test_reads_that_do_not_fit_raise_value_error():
    from smftools_b200 import engine

    m = model_from_params(O.synth_params(4))
    t = m._host_tables()
    lim = engine.max_positions(t)
    assert lim > 300_000  # 64 segments of the on-chip staging
    with pytest.raises(ValueError, match="too long"):
        m.decode(np.zeros((1, lim + 64), dtype=np.float32))
    # Viterbi has no such limit
    sv = engine.viterbi(t, torch.zeros((1, lim + 64), device=DEV), None)
    assert sv.shape == (1, lim + 64)


def test_generic_and_fast_kernels_agree(monkeypatch):
    """SMF_HMM_FORCE_GENERIC is read at library load; compare through a subprocess-free route:
    multi-channel data with one channel replicated exercises the generic kernel against the
    single-channel fast path."""
    K = 3
    X32 = O.synth_reads(64, 900, K, 77, prob=True)
    p1 = O.synth_params(K)
    m1 = model_from_params(p1)
    _, g_fast = m1.decode(X32, compact=True)
    # two channels, the second entirely missing -> same posterior as the single-channel model
    X2 = np.stack([X32, np.full_like(X32, np.nan)], axis=2)
    p2 = O.Params(p1.start, p1.trans, np.stack([p1.emission, np.full(K, 0.5)], axis=1)).normalize()
    m2 = model_from_params(p2, "multi")
    _, g_gen = m2.decode(X2, compact=True)
    assert np.abs(g_fast - g_gen).max() <= 2e-6


def test_fit_stops_on_relative_tolerance_like_the_reference():
    K = 2
    X = O.synth_reads(200, 120, K, 5).astype(np.float64)
    p0 = O.init_params(K, [0.8, 0.2])
    m = model_from_params(p0)
    hist = m.fit(X, max_iter=60, tol=1e-5)
    _, hist_ref = O.fit(X, p0, max_iter=60, tol=1e-5)
    assert len(hist) == len(hist_ref) < 60
    assert np.allclose(hist, hist_ref, rtol=1e-6)
    # frozen parameters drift exactly like the reference (re-normalisation with +eps every iteration)
    m2 = model_from_params(p0)
    m2.fit(X, max_iter=3, tol=0.0, update_start=False, update_trans=False, update_emission=False)
    q, _ = O.fit(X, p0, max_iter=3, tol=0.0, update_start=False, update_trans=False, update_emission=False)
    np.testing.assert_allclose(m2.trans.cpu().numpy(), q.trans, rtol=0, atol=1e-15)
    np.testing.assert_allclose(m2.start.cpu().numpy(), q.start, rtol=0, atol=1e-15)
