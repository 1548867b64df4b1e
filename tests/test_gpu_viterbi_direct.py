"""GPU parity of viterbi_direct_kernel (no shared-memory staging: per-thread 16-position chunks, packed
back-pointers written 16 bytes at a time) against the oracle and, bit for bit, against the staged
viterbi_kernel on the same reads."""
import numpy as np
import pytest
import torch

from helpers import assert_viterbi_matches, model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture
def vit_modes():
    from smftools_b200 import engine

    yield engine
    engine.set_option("vit", -1)


CASES = [(2, 16, 70, False), (3, 10000, 40, True), (4, 2000, 257, False), (3, 500, 333, True), (2, 4, 33, True),
         (4, 1204, 131, True), (3, 48, 1000, False), (2, 5000, 129, False), (3, 8, 1, True), (4, 20, 130, False)]


@pytest.mark.parametrize("K,L,N,prob", CASES)
def test_direct_viterbi_matches_oracle_and_staged_kernel(K, L, N, prob, vit_modes):
    engine = vit_modes
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 4100 + K + L % 91, prob=prob, nan_rate=0.3)
    variants = [torch.from_numpy(X32).to(dev)]
    if not prob and L % 16 == 0:
        variants.append(torch.from_numpy(np.where(np.isnan(X32), 255, X32).astype(np.uint8)).to(dev))
    for obs in variants:
        engine.set_option("vit", 1)
        st = engine.viterbi(tables, obs, None)
        engine.set_option("vit", 0)
        st0 = engine.viterbi(tables, obs, None)
        assert torch.equal(st, st0), "direct and staged Viterbi kernels disagree"
        assert_viterbi_matches(X32.astype(np.float64), p, None, st.cpu().numpy().astype(np.int64))


def test_direct_viterbi_exact_ties_take_the_first_index(vit_modes):
    """All-missing reads and symmetric models produce exact ties at every step: first maximum wins
    (torch.argmax / torch.max semantics of HMM.py:654-664)."""
    engine = vit_modes
    dev = torch.device("cuda")
    K = 3
    p = O.Params(np.full(K, 1.0 / K), np.full((K, K), 1.0 / K), np.full(K, 0.5), eps=1e-8).normalize()
    tables = model_from_params(p)._host_tables()
    X = np.full((5, 64), np.nan, dtype=np.float32)
    X[1, ::3] = 1.0
    X[2, ::2] = 0.0
    ref = O.viterbi(X.astype(np.float64), p)
    for v in (1, 0, -1):
        engine.set_option("vit", v)
        st = engine.viterbi(tables, torch.from_numpy(X).to(dev), None).cpu().numpy()
        assert np.array_equal(st, ref)


def test_direct_viterbi_unaligned_rows_fall_back(vit_modes):
    """L % 4 != 0 (or float64 input) is served by the staged kernel: same API, same results."""
    engine = vit_modes
    dev = torch.device("cuda")
    K = 2
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    for L, dt in [(37, np.float32), (64, np.float64), (1001, np.float32)]:
        X32 = O.synth_reads(50, L, K, 77 + L, prob=True, nan_rate=0.3)
        st = engine.viterbi(tables, torch.from_numpy(X32.astype(dt)).to(dev), None).cpu().numpy().astype(np.int64)
        assert_viterbi_matches(X32.astype(np.float64), p, None, st)
