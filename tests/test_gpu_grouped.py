"""GPU parity of the grouped launches (smf_hmm_grouped_fit / smf_hmm_grouped_posterior): many ragged
(model, N_g, L_g) groups in one launch sequence, the Baum-Welch loop entirely on the device.

Every group is checked against the oracle run on that group alone with that group's parameters, and against
the per-group path of the host mirror (model.fit / model.decode)."""
import numpy as np
import pytest
import torch

import smftools_b200 as S
from helpers import assert_states_match_up_to_ties, model_from_params, params_of_model
from oracle import hmm_oracle as O
from smftools_b200 import grouped

pytestmark = pytest.mark.gpu


def _code(X32):
    return np.where(np.isnan(X32), 255, X32).astype(np.uint8)


def _perturbed(K, g):
    """Per-group parameters: the synthetic ones with a group-dependent tilt (every group its own model)."""
    p = O.synth_params(K)
    em = np.clip(p.emission + 0.04 * ((g % 5) - 2) * np.linspace(1.0, -1.0, K), 0.02, 0.98)
    tr = p.trans.copy()
    tr *= 1.0 + 0.1 * (g % 3)
    np.fill_diagonal(tr, 0.0)
    np.fill_diagonal(tr, 1.0 - tr.sum(axis=1))
    return O.Params(p.start, tr, em, eps=p.eps).normalize()


SHAPES = [(70, 16), (300, 170), (129, 1008), (1, 37), (257, 500), (40, 2003), (128, 64), (513, 333)]


def _groups(K, prob, seed0):
    out = []
    for g, (N, L) in enumerate(SHAPES):
        X = O.synth_reads(N, L, K, seed0 + g, prob=prob, nan_rate=0.3)
        out.append((X, _perturbed(K, g)))
    return out


@pytest.mark.parametrize("K,prob", [(2, False), (3, False), (4, False), (3, True), (2, True)])
def test_grouped_posterior_matches_oracle_per_group(K, prob):
    gs = _groups(K, prob, 1000 + 10 * K)
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) if not prob else X.astype(np.float32) for X, _ in gs]
    res = grouped.decode_groups(models, obs, want_loglik=True)
    torch.cuda.synchronize()
    for (X, p), (gam, st, ll) in zip(gs, res):
        g_ref, ll_ref = O.posterior(X.astype(np.float64), p)
        assert tuple(gam.shape) == g_ref.shape
        gam_h = gam.cpu().numpy()
        assert np.abs(gam_h - g_ref).max() <= 1e-5
        assert_states_match_up_to_ties(st.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
        np.testing.assert_allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)


def test_grouped_posterior_single_state_column_and_states_only():
    K = 3
    gs = _groups(K, False, 77)
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    res1 = grouped.decode_groups(models, obs, gamma_state=1, want_states=False)
    res2 = grouped.decode_groups(models, obs, want_gamma=False)
    torch.cuda.synchronize()
    for (X, p), (gam, st, _), (gam2, st2, _) in zip(gs, res1, res2):
        g_ref, _ = O.posterior(X.astype(np.float64), p)
        assert st is None and gam2 is None
        assert np.abs(gam.cpu().numpy() - g_ref[:, :, 1]).max() <= 1e-5
        assert_states_match_up_to_ties(st2.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)


@pytest.mark.parametrize("K,prob", [(2, False), (3, False), (4, False), (3, True)])
def test_grouped_fit_matches_oracle_per_group(K, prob):
    gs = _groups(K, prob, 2000 + 10 * K)
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) if not prob else X.astype(np.float32) for X, _ in gs]
    hists, status = grouped.fit_groups(models, obs, max_iter=6, tol=0.0, return_status=True)
    assert status == [grouped.FIT_RAN_ALL] * len(gs)
    for (X, p), m, h in zip(gs, models, hists):
        q, h_ref = O.fit(X.astype(np.float64), p, max_iter=6, tol=0.0)
        np.testing.assert_allclose(h, h_ref, rtol=1e-6)
        got = params_of_model(m)
        assert np.abs(got.start - q.start).max() <= 1e-4
        assert np.abs(got.trans - q.trans).max() <= 1e-4
        assert np.abs(got.emission.reshape(-1) - q.emission.reshape(-1)).max() <= 1e-4


def test_grouped_fit_equals_per_group_fit_and_is_deterministic():
    K = 2
    gs = _groups(K, False, 31)
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    a = [model_from_params(p) for _, p in gs]
    b = [model_from_params(p) for _, p in gs]
    c = [model_from_params(p) for _, p in gs]
    ha = grouped.fit_groups(a, obs, max_iter=5, tol=0.0)
    hb = grouped.fit_groups(b, obs, max_iter=5, tol=0.0)
    assert ha == hb, "grouped fit is not run-to-run deterministic"
    for ma, mb in zip(a, b):
        for k in ("start", "trans", "emission"):
            assert torch.equal(ma.state_dict()[k], mb.state_dict()[k])
    for (X, _), ma, mc, h in zip(gs, a, c, ha):
        hc = mc.fit(S.CodedU8(_code(X)), max_iter=5, tol=0.0)
        np.testing.assert_allclose(h, hc, rtol=1e-6)  # different kernels (chunk-parallel vs sequential), fp32 recursions
        for k in ("start", "trans", "emission"):
            assert (ma.state_dict()[k] - mc.state_dict()[k]).abs().max().item() <= 1e-5


def test_grouped_fit_stop_rule_is_per_group():
    """With a loose tolerance groups stop at different iterations; each must stop exactly where the oracle stops,
    with the parameters of that iteration (HMM.py:1616-1628), while the others keep iterating."""
    K = 2
    gs = _groups(K, False, 555)
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    tol = 2e-4
    hists, status = grouped.fit_groups(models, obs, max_iter=12, tol=tol, return_status=True)
    n_iters = set()
    for (X, p), m, h, st in zip(gs, models, hists, status):
        q, h_ref = O.fit(X.astype(np.float64), p, max_iter=12, tol=tol)
        assert len(h) == len(h_ref), (len(h), len(h_ref))
        assert st == (grouped.FIT_CONVERGED if len(h_ref) < 12 else st)
        np.testing.assert_allclose(h, h_ref, rtol=1e-6, atol=1e-5)  # a one-read group's log-likelihood goes to ~0
        assert np.abs(params_of_model(m).emission.reshape(-1) - q.emission.reshape(-1)).max() <= 1e-4
        n_iters.add(len(h))
    assert len(n_iters) > 1, "test needs groups that stop at different iterations"


def test_grouped_fit_update_flags():
    K = 3
    gs = _groups(K, False, 808)[:4]
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    grouped.fit_groups(models, obs, max_iter=3, tol=0.0, update_start=False, update_trans=False)
    for (X, p), m in zip(gs, models):
        q, _ = O.fit(X.astype(np.float64), p, max_iter=3, tol=0.0, update_start=False, update_trans=False)
        got = params_of_model(m)
        assert np.abs(got.start - q.start).max() <= 1e-12
        assert np.abs(got.trans - q.trans).max() <= 1e-12
        assert np.abs(got.emission.reshape(-1) - q.emission.reshape(-1)).max() <= 1e-4


def test_grouped_fit_hands_extreme_groups_back():
    """A group whose emissions sit at the eps clamp needs a finer rescale cadence than the grouped kernels use:
    it is reported as handed back and finished through the model's own fit -- same result as fitting it alone."""
    K = 2
    gs = _groups(K, False, 99)[:3]
    p_ext = O.Params(np.array([0.5, 0.5]), np.array([[0.999999, 1e-6], [1e-6, 0.999999]]), np.array([1e-8, 1.0 - 1e-8]),
                     eps=1e-8).normalize()
    Xe = O.synth_reads(50, 200, K, 5, prob=False, nan_rate=0.3)
    gs.insert(1, (Xe, p_ext))
    models = [model_from_params(p) for _, p in gs]
    alone = model_from_params(p_ext)
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    hists, status = grouped.fit_groups(models, obs, max_iter=3, tol=0.0, return_status=True)
    assert status[1] == grouped.FIT_HANDED_BACK and status[0] == status[2] == grouped.FIT_RAN_ALL
    h_alone = alone.fit(S.CodedU8(_code(Xe)), max_iter=3, tol=0.0)
    np.testing.assert_allclose(hists[1], h_alone, rtol=1e-12)
    for (X, p), m, h in zip(gs, models, hists):
        _, h_ref = O.fit(X.astype(np.float64), p, max_iter=3, tol=0.0)
        np.testing.assert_allclose(h, h_ref, rtol=1e-6)


def test_grouped_accepts_padded_views_and_rejects_mixed_input():
    K = 2
    X = O.synth_reads(64, 100, K, 3, prob=False, nan_rate=0.3)
    p = O.synth_params(K)
    dev = torch.device("cuda")
    buf = torch.full((64, 112), 255, dtype=torch.uint8, device=dev)
    buf[:, :100] = torch.from_numpy(_code(X)).to(dev)
    view = S.CodedU8(buf[:, :100])
    (gam, st, _), = grouped.decode_groups([model_from_params(p)], [view])
    g_ref, _ = O.posterior(X.astype(np.float64), p)
    assert np.abs(gam.cpu().numpy() - g_ref).max() <= 1e-5
    with pytest.raises(ValueError):
        grouped.decode_groups([model_from_params(p)] * 2, [view, X.astype(np.float32)])
    with pytest.raises(ValueError):
        grouped.decode_groups([model_from_params(p), model_from_params(O.synth_params(3))], [view, view])
    with pytest.raises(ValueError):
        grouped.fit_groups([S.create_hmm({"hmm_n_states": 2}, arch="multi")], [view])


def test_many_small_groups_one_launch():
    """The reference's regime: hundreds of <= 1000-read fits.  256 groups, all fitted in one enqueue."""
    K = 2
    rng = np.random.default_rng(4)
    gs = []
    for g in range(256):
        N = int(rng.integers(20, 400))
        L = int(rng.integers(30, 400))
        gs.append((O.synth_reads(N, L, K, 9000 + g, prob=False, nan_rate=0.3), _perturbed(K, g)))
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    hists = grouped.fit_groups(models, obs, max_iter=4, tol=0.0)
    for g in range(0, 256, 37):
        X, p = gs[g]
        q, h_ref = O.fit(X.astype(np.float64), p, max_iter=4, tol=0.0)
        np.testing.assert_allclose(hists[g], h_ref, rtol=1e-6)
        assert np.abs(params_of_model(models[g]).emission.reshape(-1) - q.emission.reshape(-1)).max() <= 1e-4


@pytest.mark.parametrize("K", [2, 4])
def test_stepwise_fit_of_read_sharded_groups_equals_the_whole_fit(K):
    """smf_hmm_grouped_fit_step: two 'ranks' (two workspaces in this process) each hold a shard of every group's reads
    -- one of them an empty shard of one group -- and the caller sums the two statistics matrices between the E-step
    and the M-step, which is what the NCCL all-reduce does across GPUs.  Both ranks must end with the parameters,
    histories and stop decisions of the unsharded grouped fit (and of the oracle)."""
    from smftools_b200 import _lib
    from smftools_b200.engine import _ptr, _stream_ptr

    dev = torch.device("cuda")
    gs = _groups(K, False, 4000 + K)
    G = len(gs)
    cuts = [max(0, min(X.shape[0], (X.shape[0] * (3 + g)) // 7)) for g, (X, _) in enumerate(gs)]
    cuts[3] = 0  # the one-read group lives entirely on rank 1: rank 0 holds an empty shard
    tol, max_iter = 1e-4, 9
    whole = [model_from_params(p) for _, p in gs]
    h_whole, st_whole = grouped.fit_groups(whole, [S.CodedU8(_code(X)) for X, _ in gs], max_iter=max_iter, tol=tol,
                                           return_status=True)
    lib = _lib.load()
    S_len = K + K * K + 2 * K + 1
    ranks = []
    for r in range(2):
        shards = [S.CodedU8(torch.from_numpy(_code(X[:c] if r == 0 else X[c:])).to(dev)) for (X, _), c in zip(gs, cuts)]
        tab = grouped._GroupTable(shards, dev)
        need = int(lib.smf_hmm_grouped_workspace_bytes(K, tab.arr, G, 1))
        assert need > 0
        st = dict(tab=tab, need=need, ws=torch.empty((need,), dtype=torch.uint8, device=dev),
                  params=torch.from_numpy(grouped.pack_parameters([model_from_params(p) for _, p in gs])).to(dev),
                  hist=torch.zeros((G, max_iter), dtype=torch.float64, device=dev),
                  n_iter=torch.zeros((G,), dtype=torch.int32, device=dev),
                  status=torch.zeros((G,), dtype=torch.int32, device=dev),
                  stats=torch.zeros((G, S_len), dtype=torch.float64, device=dev))
        ranks.append(st)

    def step(st, phase, it):
        rc = lib.smf_hmm_grouped_fit_step(phase, K, st["tab"].obs_dtype, st["tab"].arr, G, _ptr(st["params"]), it, max_iter, tol,
                                          1e-8, 7, _ptr(st["hist"]), _ptr(st["n_iter"]), _ptr(st["status"]), _ptr(st["stats"]),
                                          _ptr(st["ws"]), st["need"], _stream_ptr(dev))
        _lib.check(rc, "smf_hmm_grouped_fit_step")

    for st in ranks:
        step(st, _lib.FIT_PHASE_BEGIN, 0)
    for it in range(max_iter):
        for st in ranks:
            step(st, _lib.FIT_PHASE_ESTEP, it)
        total = ranks[0]["stats"] + ranks[1]["stats"]  # the all-reduce
        for st in ranks:
            st["stats"].copy_(total)
            step(st, _lib.FIT_PHASE_MSTEP, it)
    torch.cuda.synchronize()
    for k in ("params", "hist", "n_iter", "status"):
        assert torch.equal(ranks[0][k], ranks[1][k]), f"ranks disagree on {k}"  # deterministic replicated M-step
    n_it = ranks[0]["n_iter"].cpu().numpy()
    hist = ranks[0]["hist"].cpu().numpy()
    params = ranks[0]["params"].cpu().numpy()
    assert [int(v) for v in ranks[0]["status"].cpu().numpy()] == st_whole
    for g, ((X, p), m, h) in enumerate(zip(gs, whole, h_whole)):
        if st_whole[g] == grouped.FIT_HANDED_BACK:
            # fit_groups finished this group through model.fit; the raw entry point reports where it stopped
            n = int(n_it[g])
            assert 0 < n <= len(h)
            np.testing.assert_allclose(hist[g, :n], h[:n], rtol=1e-6, atol=1e-5)
            continue
        assert int(n_it[g]) == len(h)
        # the shards put different reads into a CTA, so the fp32 partial sums (flushed into fp64 every 8 chunks) differ
        np.testing.assert_allclose(hist[g, :len(h)], h, rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(params[g], grouped.pack_parameters([m])[0], rtol=0, atol=1e-5)
        q, h_ref = O.fit(X.astype(np.float64), p, max_iter=max_iter, tol=tol)
        assert len(h_ref) == len(h)
        np.testing.assert_allclose(h, h_ref, rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("K,prob,seed", [(2, False, 1), (3, True, 2), (4, False, 3), (4, True, 4)])
def test_random_ragged_groups_including_degenerate_lengths(K, prob, seed):
    """Seeded random group tables: 24 groups of 1-400 reads x 1-1500 positions (lengths 1, 2, 3, 15-17 and 31-33 forced
    in), fitted for 3 iterations and decoded in one launch sequence each; a sample is checked against the oracle."""
    rng = np.random.default_rng(seed)
    lengths = [1, 2, 3, 15, 16, 17, 31, 32, 33] + [int(x) for x in rng.integers(4, 1500, size=15)]
    reads = [int(x) for x in rng.integers(1, 400, size=len(lengths))]
    gs = []
    for g, (N, L) in enumerate(zip(reads, lengths)):
        gs.append((O.synth_reads(N, L, K, 9100 + 31 * seed + g, prob=prob, nan_rate=0.3), _perturbed(K, g)))
    models = [model_from_params(p) for _, p in gs]
    obs = [S.CodedU8(_code(X)) if not prob else X.astype(np.float32) for X, _ in gs]
    res = grouped.decode_groups(models, obs, want_loglik=True)
    hists = grouped.fit_groups(models, obs, max_iter=3, tol=0.0)
    torch.cuda.synchronize()
    for g in list(range(9)) + [int(x) for x in rng.choice(np.arange(9, len(gs)), size=4, replace=False)]:
        X, p = gs[g]
        g_ref, ll_ref = O.posterior(X.astype(np.float64), p)
        gam, st, ll = res[g]
        assert tuple(gam.shape) == g_ref.shape and tuple(st.shape) == g_ref.shape[:2]
        assert np.abs(gam.cpu().numpy() - g_ref).max() <= 1e-5, (g, X.shape)
        np.testing.assert_allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)
        q, h_ref = O.fit(X.astype(np.float64), p, max_iter=3, tol=0.0)
        np.testing.assert_allclose(hists[g], h_ref, rtol=1e-6, atol=1e-5)
        assert np.abs(params_of_model(models[g]).emission.reshape(-1) - q.emission.reshape(-1)).max() <= 1e-4


@pytest.mark.parametrize("K,prob,tol", [(2, False, 0.0), (2, False, 1e-4), (3, True, 1e-4), (4, False, 1e-5), (3, False, 0.0)])
def test_one_launch_fit_loop_is_bit_identical_to_the_kernel_sequence(K, prob, tol):
    """Small fits run the whole Baum-Welch loop in one cooperative launch (grid barriers between E-step and M-step);
    option "fitloop" = 0 enqueues forward / backward / M-step kernels per iteration instead.  Same arithmetic, same
    summation orders: parameters, log-likelihood histories, iteration counts and stop reasons are identical."""
    from smftools_b200 import engine

    gs = _groups(K, prob, 4100 + K)
    obs = [S.CodedU8(_code(X)) if not prob else X.astype(np.float32) for X, _ in gs]
    out = []
    try:
        for mode in (1, 0):
            engine.set_option("fitloop", mode)
            models = [model_from_params(p) for _, p in gs]
            hists, status = grouped.fit_groups(models, obs, max_iter=12, tol=tol, return_status=True)
            torch.cuda.synchronize()
            out.append((models, hists, status))
    finally:
        engine.set_option("fitloop", 1)
    (m1, h1, s1), (m0, h0, s0) = out
    assert list(s1) == list(s0)
    if tol > 0.0:
        assert len({len(h) for h in h1}) > 1  # groups stop at different iterations
    for a, b, ha, hb in zip(m1, m0, h1, h0):
        assert len(ha) == len(hb)
        np.testing.assert_array_equal(np.asarray(ha), np.asarray(hb))
        for name in ("start", "trans", "emission"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_one_launch_fit_loop_single_small_group_matches_oracle():
    K, N, L, iters = 2, 1000, 170, 25
    X = O.synth_reads(N, L, K, 5, prob=False, nan_rate=0.3)
    p = O.synth_params(K)
    m = model_from_params(p)
    hist = grouped.fit_groups([m], [S.CodedU8(_code(X))], max_iter=iters, tol=0.0)[0]
    ref, ref_hist = O.fit(X.astype(np.float64), p, max_iter=iters, tol=0.0)
    assert len(hist) == iters
    np.testing.assert_allclose(np.asarray(hist), np.asarray(ref_hist), rtol=1e-6)
    q = params_of_model(m)
    assert np.abs(q.emission - ref.emission).max() <= 1e-4
    assert np.abs(q.trans - ref.trans).max() <= 1e-4


def test_one_launch_fit_loop_with_several_resident_ctas_per_sm():
    """~330 work items: more CTAs than SMs, so the grid barrier runs with several resident CTAs per SM."""
    from smftools_b200 import engine

    K = 2
    shapes = [(30000, 64), (12000, 48), (257, 160)]
    gs = [(O.synth_reads(N, L, K, 900 + g, prob=False, nan_rate=0.3), _perturbed(K, g)) for g, (N, L) in enumerate(shapes)]
    obs = [S.CodedU8(_code(X)) for X, _ in gs]
    out = []
    try:
        for mode in (1, 0):
            engine.set_option("fitloop", mode)
            models = [model_from_params(p) for _, p in gs]
            hists = grouped.fit_groups(models, obs, max_iter=8, tol=0.0)
            torch.cuda.synchronize()
            out.append((models, hists))
    finally:
        engine.set_option("fitloop", 1)
    (m1, h1), (m0, h0) = out
    for a, b, ha, hb in zip(m1, m0, h1, h0):
        np.testing.assert_array_equal(np.asarray(ha), np.asarray(hb))
        for name in ("start", "trans", "emission"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
    # and against the oracle on the small group
    X, p = gs[2]
    ref, ref_hist = O.fit(X.astype(np.float64), p, max_iter=8, tol=0.0)
    np.testing.assert_allclose(np.asarray(h1[2]), np.asarray(ref_hist), rtol=1e-6)
