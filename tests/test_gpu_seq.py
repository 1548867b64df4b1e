"""GPU parity of the sequential (one thread per read) E-step: seq_forward_kernel + seq_backward_kernel.

The library picks this path for large batches (smf_hmm_set_option("seq", 1) forces it for any eligible
batch); it must agree with the oracle and with the chunk-parallel kernels on the same reads."""
import numpy as np
import pytest
import torch

from helpers import model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_seq():
    from smftools_b200 import engine

    engine.set_option("seq", 1)
    yield
    engine.set_option("seq", -1)


def _code(X32):
    return np.where(np.isnan(X32), 255, X32).astype(np.uint8)


@pytest.mark.parametrize("K,L,N,prob", [(2, 16, 70, False), (2, 512, 300, False), (3, 1008, 129, False), (4, 2000, 257, False),
                                        (4, 48, 1000, False), (3, 500, 200, True), (4, 1204, 131, True), (2, 4, 33, True),
                                        (4, 8000, 40, False), (2, 170, 129, False), (4, 37, 300, False), (3, 1001, 64, True),
                                        (4, 170, 500, False)])
def test_seq_estep_matches_oracle_and_chunk_parallel_path(K, L, N, prob, force_seq):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 300 + K + L % 89, prob=prob, nan_rate=0.3)
    ref = O.stats_vector(O.estep(X32.astype(np.float64), p, None))
    scale = np.maximum(np.abs(ref), 1.0)
    variants = [("f32", torch.from_numpy(X32.astype(np.float32)).to(dev))]
    if not prob:
        variants.append(("u8", torch.from_numpy(_code(X32)).to(dev)))  # rows off 16-byte boundaries are read element-wise
    for name, obs in variants:
        engine.set_option("seq", 1)
        st = engine.estep(tables, obs, None).cpu().numpy()
        assert np.all(np.isfinite(st)), name
        assert np.all(np.abs(st - ref) / scale <= 2e-6), (name, (np.abs(st - ref) / scale).max())
        assert np.array_equal(st, engine.estep(tables, obs, None).cpu().numpy()), "not deterministic"
        engine.set_option("seq", 0)
        st0 = engine.estep(tables, obs, None).cpu().numpy()
        assert np.all(np.abs(st0 - st) / scale <= 4e-6), (name, (np.abs(st0 - st) / scale).max())


def test_seq_path_is_the_policy_for_the_configs3_shape():
    """65,536 replicated reads x 8 kb, K = 4, uint8 code: the default policy takes the sequential kernels
    (workspace = checkpoints), and the statistics equal replica-count x the oracle's."""
    from smftools_b200 import _lib, engine

    K, L, base_n, reps = 4, 8000, 32, 2048
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    base = O.synth_reads(base_n, L, K, 909, prob=False, nan_rate=0.3)
    obs = torch.from_numpy(np.tile(_code(base), (reps, 1))).to(dev)
    N = base_n * reps
    S = engine.stats_len(tables)
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_ESTEP, tables.ref, N, L))
    assert need >= 148 * 8 * S * 8 + (L // 16) * N * 16
    st = engine.estep(tables, obs, None).cpu().numpy()
    ref = O.stats_vector(O.estep(base.astype(np.float64), p, None)) * reps
    scale = np.maximum(np.abs(ref), 1.0)
    assert np.all(np.abs(st - ref) / scale <= 2e-6), (np.abs(st - ref) / scale).max()
    engine.set_option("seq", 0)
    try:
        st0 = engine.estep(tables, obs, None).cpu().numpy()
    finally:
        engine.set_option("seq", -1)
    assert np.all(np.abs(st0 - st) / scale <= 4e-6)


@pytest.mark.parametrize("K,L,N,prob", [(3, 1008, 129, False), (4, 2000, 257, False), (2, 512, 300, False), (3, 500, 200, True),
                                        (4, 1204, 131, True), (4, 16, 1000, False), (3, 8000, 40, True)])
def test_seq_posterior_matches_oracle(K, L, N, prob, force_seq):
    """Posterior decode on the sequential kernels (seq_backward_kernel<MODE_POST>: posteriors staged in shared
    memory, written as whole sectors) for every output combination of smf_hmm_posterior."""
    from helpers import assert_states_match_up_to_ties
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 700 + K + L % 97, prob=prob, nan_rate=0.3)
    g_ref, ll_ref = O.posterior(X32.astype(np.float64), p)
    variants = [torch.from_numpy(X32.astype(np.float32)).to(dev)]
    if not prob and L % 16 == 0:
        variants.append(torch.from_numpy(_code(X32)).to(dev))
    for obs in variants:
        gam, st, ll = engine.posterior(tables, obs, want_loglik=True)
        assert np.abs(gam.cpu().numpy() - g_ref).max() <= 1e-5
        assert_states_match_up_to_ties(st.cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
        np.testing.assert_allclose(ll.cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)
        g1, s1, _ = engine.posterior(tables, obs, gamma_state=K - 1, want_states=False)
        assert s1 is None and np.abs(g1.cpu().numpy() - g_ref[:, :, K - 1]).max() <= 1e-5
        g2, s2, _ = engine.posterior(tables, obs, want_gamma=False)
        assert g2 is None and torch.equal(s2, st)
        # the same batch on the chunk-parallel kernels
        engine.set_option("seq", 0)
        gam0, st0, _ = engine.posterior(tables, obs)
        engine.set_option("seq", 1)
        assert (gam0 - gam).abs().max().item() <= 4e-6


@pytest.mark.parametrize("reps", [1024, 1400])
def test_seq_posterior_is_the_policy_for_large_k4_batches(reps):
    """49,152 / 67,200 replicated reads x 512 positions, K = 4: the default policy decodes on the sequential kernels
    (the library asks for the checkpoint workspace); every replica must get the oracle's posterior of its base read.
    (two batch sizes: 384 and 525 work items)"""
    from helpers import assert_states_match_up_to_ties
    from smftools_b200 import _lib, engine

    K, L, base_n = 4, 512, 48
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    base = O.synth_reads(base_n, L, K, 4242, prob=True, nan_rate=0.3)
    g_ref, ll_ref = O.posterior(base.astype(np.float64), p)
    obs = torch.from_numpy(np.tile(base.astype(np.float32), (reps, 1))).to(dev)
    N = base_n * reps
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, tables.ref, N, L))
    assert need >= (L // 16) * N * 16  # alpha checkpoints of the sequential kernels
    g, s, ll = engine.posterior(tables, obs, want_loglik=True)
    g = g.view(reps, base_n, L, K)
    assert torch.equal(g[0], g[reps - 1]) and torch.equal(g[0], g[reps // 2])  # same read, same result, any thread
    assert np.abs(g[3].cpu().numpy() - g_ref).max() <= 1e-5
    assert_states_match_up_to_ties(s.view(reps, base_n, L)[5].cpu().numpy().astype(np.int64), g_ref.argmax(axis=2), g_ref)
    np.testing.assert_allclose(ll.view(reps, base_n)[7].cpu().numpy(), ll_ref, rtol=1e-6, atol=1e-4)
    engine.set_option("seq", 0)
    try:
        g0, _, _ = engine.posterior(tables, obs)
    finally:
        engine.set_option("seq", -1)
    assert (g0.view(reps, base_n, L, K)[0] - g[0]).abs().max().item() <= 4e-6
