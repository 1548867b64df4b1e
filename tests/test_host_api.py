"""CPU-side checks: registry / config / parameter layout of the host mirror, C-ABI exports,
and the loud failure when no CUDA device is present (no CPU fallback)."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import smftools_b200 as S
from smftools_b200 import _lib, hmm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_and_factory():
    assert set(hmm._HMM_REGISTRY) >= {"single", "multi", "single_distance_binned"}
    assert isinstance(S.create_hmm(None), S.SingleBernoulliHMM)                      # default arch
    assert isinstance(S.create_hmm({"hmm_arch": "multi"}), S.MultiBernoulliHMM)      # dict cfg
    assert isinstance(S.create_hmm(SimpleNamespace(hmm_arch="single_distance_binned")),
                      S.DistanceBinnedSingleBernoulliHMM)                            # object cfg
    assert isinstance(S.create_hmm({"hmm_arch": "multi"}, arch="single"), S.SingleBernoulliHMM)  # arch wins
    assert isinstance(S.HMM.from_config({}, arch="multi"), S.MultiBernoulliHMM)
    with pytest.raises(KeyError):
        S.create_hmm({}, arch="nope")
    assert S.SingleBernoulliHMM.hmm_name == "single"

    @S.register_hmm("custom_test_arch")
    class Custom(S.SingleBernoulliHMM):
        pass

    assert isinstance(S.create_hmm({}, arch="custom_test_arch"), Custom)
    hmm._HMM_REGISTRY.pop("custom_test_arch")


def test_initial_parameters_match_reference_layout():
    # SURVEY 8a "edge semantics": default 2x2 init flattens to [0.8, 0.2(, 0.2(, 0.8))]
    init = [[0.8, 0.2], [0.2, 0.8]]
    for K, want in ((2, [0.8, 0.2]), (3, [0.8, 0.2, 0.2]), (4, [0.8, 0.2, 0.2, 0.8])):
        m = S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": init}, arch="single")
        assert m.emission.dtype == torch.float64
        np.testing.assert_allclose(m.emission.numpy(), want)
        np.testing.assert_allclose(m.start.numpy(), np.full(K, 1.0 / K), atol=1e-12)
        np.testing.assert_allclose(m.trans.numpy().sum(1), 1.0, atol=1e-12)
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": init}, arch="multi")
    np.testing.assert_allclose(m.emission.numpy(), init)
    m = S.create_hmm({"hmm_n_states": 3, "hmm_init_emission_probs": init}, arch="multi")
    np.testing.assert_allclose(m.emission.numpy(), np.full((3, 2), 0.5))  # shape mismatch -> 0.5
    m = S.create_hmm({"hmm_n_states": 4}, arch="single_distance_binned")
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "start": (4,), "trans": (4, 4), "emission": (4,), "trans_by_bin": (7, 4, 4)}
    assert m.n_bins == 7 and list(m.distance_bins) == [1, 5, 10, 25, 50, 100]
    m32 = S.create_hmm({"hmm_dtype": "float32"}, arch="single")
    assert m32.start.dtype == torch.float32
    with pytest.raises(ValueError):
        S.SingleBernoulliHMM(n_states=1)


def test_parameters_match_oracle_initialisation():
    from oracle import hmm_oracle as O

    for arch in ("single", "multi", "single_distance_binned"):
        for K in (2, 3):
            m = S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": [0.85, 0.4, 0.05][:K]}, arch=arch)
            p = O.init_params(K, [0.85, 0.4, 0.05][:K], arch=arch)
            np.testing.assert_allclose(m.start.numpy(), p.start, rtol=0, atol=1e-15)
            np.testing.assert_allclose(m.trans.numpy(), p.trans, rtol=0, atol=1e-15)
            np.testing.assert_allclose(m.emission.numpy().reshape(p.emission.shape), p.emission, rtol=0, atol=1e-15)
            if arch == "single_distance_binned":
                np.testing.assert_allclose(m.trans_by_bin.numpy(), p.trans_by_bin, rtol=0, atol=1e-15)


def test_state_target_resolution():
    m = S.create_hmm({"hmm_n_states": 3, "hmm_init_emission_probs": [0.85, 0.4, 0.05]}, arch="single")
    assert m.resolve_target_state_index("Modified") == 0
    assert m.resolve_target_state_index("Non-Modified") == 2
    assert m.resolve_target_state_index("accessible") == 0
    assert m.resolve_target_state_index("closed") == 2
    assert m.resolve_target_state_index(1) == 1
    assert m.resolve_target_state_index(17) == 2 and m.resolve_target_state_index(-3) == 0
    assert m.resolve_target_state_index("something else") == 0
    mm = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [[0.1, 0.2], [0.9, 0.6]]}, arch="multi")
    assert mm.modified_state_index() == 1


def test_m_step_matches_oracle():
    """The host M-step (the only arithmetic left on the host) against the oracle's restatement of
    HMM.py:1598-1614, including the double eps normalisation and frozen-parameter drift."""
    from oracle import hmm_oracle as O

    rng = np.random.default_rng(0)
    for arch, C in (("single", 1), ("multi", 3), ("single_distance_binned", 1)):
        K = 3
        m = S.create_hmm({"hmm_n_states": K, "hmm_n_channels": C}, arch=arch)
        p = O.init_params(K, None, arch=arch, n_channels=C)
        nb = 7 if arch == "single_distance_binned" else 1
        st = dict(start=rng.random(K) * 50, trans=rng.random((nb, K, K)) * 1e4,
                  emit_num=rng.random((K, C)) * 1e4, emit_den=rng.random((K, C)) * 1e4 + 1e4, ll=-123.0)
        if arch != "multi":
            st["emit_num"] = st["emit_num"].reshape(K)
            st["emit_den"] = st["emit_den"].reshape(K)
        for flags in ((True, True, True), (False, False, True), (True, False, False)):
            mm = S.create_hmm({"hmm_n_states": K, "hmm_n_channels": C}, arch=arch)
            mm._m_step(torch.tensor(O.stats_vector(st)), update_start=flags[0], update_trans=flags[1],
                       update_emission=flags[2])
            q = O.mstep(p, st, *flags)
            np.testing.assert_allclose(mm.start.numpy(), q.start, rtol=0, atol=1e-15)
            np.testing.assert_allclose(mm.trans.numpy(), q.trans, rtol=0, atol=1e-15)
            np.testing.assert_allclose(mm.emission.numpy().reshape(q.emission.shape), q.emission, rtol=0, atol=1e-15)
            if q.trans_by_bin is not None:
                np.testing.assert_allclose(mm.trans_by_bin.numpy(), q.trans_by_bin, rtol=0, atol=1e-15)


def test_save_load_roundtrip(tmp_path):
    for arch in ("single", "multi", "single_distance_binned"):
        m = S.create_hmm({"hmm_n_states": 3, "hmm_init_emission_probs": [0.9, 0.5, 0.1]}, arch=arch)
        with torch.no_grad():
            m.trans.data = torch.tensor([[0.9, 0.05, 0.05], [0.1, 0.8, 0.1], [0.2, 0.2, 0.6]], dtype=torch.float64)
        path = tmp_path / f"{arch}.pt"
        m.save(path)
        for loader in (S.HMM.load, S.BaseHMM.load):
            r = loader(path)
            assert type(r) is type(m)
            for k, v in m.state_dict().items():
                np.testing.assert_allclose(r.state_dict()[k].numpy(), v.numpy(), rtol=0, atol=3e-8)
        # nn.Module protocol used by HMMTrainer
        import copy

        c = copy.deepcopy(m)
        c.load_state_dict(m.state_dict())
        assert next(m.parameters()).dtype == torch.float64
        m.eval()


def test_feature_set_normalisation():
    fs = S.normalize_hmm_feature_sets(
        '{"footprint": {"state": "Non-Modified", "features": {"small": [6, 40], "big": [200, "inf"]}}, '
        '"cpg": {"cpg_patch": null}, "x": {"features": {"a": 30}}}')
    assert fs["footprint"] == {"features": {"small": (6.0, 40.0), "big": (200.0, np.inf)}, "state": "Non-Modified"}
    assert fs["cpg"] == {"features": {}, "state": "Modified"}  # a dict without "features"/"ranges" has none
    assert fs["x"]["features"] == {"a": (0.0, 30.0)}
    fs2 = S.normalize_hmm_feature_sets({"cpg": {"features": {"cpg_patch": None}}, "plain": [1, 2]})
    assert fs2["cpg"]["features"] == {"cpg_patch": (0.0, np.inf)} and fs2["plain"]["features"] == {}
    assert S.normalize_hmm_feature_sets(None) == {} and S.normalize_hmm_feature_sets("not a dict") == {}


def test_interval_helpers():
    B = S.BaseHMM
    assert B._runs_from_bool(np.array([0, 1, 1, 0, 1], bool)) == [(1, 3), (4, 5)]
    assert B._runs_from_bool(np.zeros(4, bool)) == []
    assert B._interval_length(np.array([10, 12, 20]), 0, 2) == 3 and B._interval_length(np.arange(3), 2, 2) == 0
    np.testing.assert_array_equal(B._write_lengths_for_binary_layer(np.array([[1, 1, 0, 1], [0, 0, 0, 0]])),
                                  [[2, 2, 0, 1], [0, 0, 0, 0]])
    np.testing.assert_array_equal(B._write_lengths_for_state_layer(np.array([[0, 0, 1, -1, 1, 1]])),
                                  [[2, 2, 1, 0, 2, 2]])
    rng = np.random.default_rng(1)
    from oracle import hmm_oracle as O

    mat = rng.random((30, 101)) < 0.6
    np.testing.assert_array_equal(hmm._run_lengths_np(mat), O.lengths_for_binary(mat))


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "smf_hmm.h")).read()
    declared = set(re.findall(r"\b(smf_hmm_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    lib = _lib.load()
    assert lib.smf_hmm_abi_version() == _lib.ABI_VERSION
    assert b"too long" in lib.smf_hmm_strerror(_lib.SMF_ERR_TOO_LONG)
    # argument validation happens before any CUDA call
    m = S.create_hmm({"hmm_n_states": 4}, arch="single_distance_binned")
    t = m._host_tables()
    assert (t.K, t.C, t.n_bins) == (4, 1, 7)
    assert lib.smf_hmm_stats_len(t.ref) == 4 + 7 * 16 + 8 + 1
    assert lib.smf_hmm_workspace_bytes(_lib.OP_VITERBI, t.ref, 10, 100) == 1000
    assert lib.smf_hmm_max_positions(t.ref, 1) > 5000
    bad = _lib.HostTables(np.zeros(5), np.zeros((1, 5, 5)), np.zeros((5, 1)), np.zeros((5, 1)))
    assert lib.smf_hmm_stats_len(bad.ref) == -1
    assert lib.smf_hmm_posterior(bad.ref, None, 0, 1, 1, None, None, -1, None, None, None, 0, None) == _lib.SMF_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.SMF_ERR_BAD_ARG, "x")
    with pytest.raises(RuntimeError):
        _lib.check(_lib.SMF_ERR_CUDA, "x")


def test_tuning_options_and_workspace_policy():
    """smf_hmm_set_option / smf_hmm_workspace_bytes are pure host logic (no CUDA call)."""
    lib = _lib.load()
    assert lib.smf_hmm_set_option(b"walk", 3) == _lib.SMF_ERR_BAD_ARG
    assert lib.smf_hmm_set_option(b"short", -2) == _lib.SMF_ERR_BAD_ARG
    assert lib.smf_hmm_set_option(b"no_such_key", 0) == _lib.SMF_ERR_BAD_ARG
    assert lib.smf_hmm_set_option(None, 0) == _lib.SMF_ERR_BAD_ARG
    t2 = S.create_hmm({"hmm_n_states": 2}, arch="single")._host_tables()
    t4 = S.create_hmm({"hmm_n_states": 4}, arch="single")._host_tables()
    t4m = S.create_hmm({"hmm_n_states": 4, "hmm_n_channels": 2}, arch="multi")._host_tables()
    S4 = int(lib.smf_hmm_stats_len(t4.ref))
    partials = 148 * 8 * S4 * 8
    try:
        # measured policy: two-kernel path (and its boundary-vector workspace) for large K = 4 batches only
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t2.ref, 100_000, 500) == 0
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4.ref, 100, 500) == 0
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4m.ref, 100_000, 500) == 0
        need = lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4.ref, 100_000, 500)
        assert need == 2 * 30 * 100_000 * 4 * 4 + 100_000 * 8  # 2 x ceil(500/17) chunks x N x K floats + N doubles
        assert lib.smf_hmm_workspace_bytes(_lib.OP_ESTEP, t4.ref, 100_000, 500) == partials + need
        assert lib.smf_hmm_workspace_bytes(_lib.OP_ESTEP, t4.ref, 100, 500) == partials
        assert lib.smf_hmm_set_option(b"walk", 0) == _lib.SMF_OK
        # without the two-kernel path a large K = 4 batch still asks for the checkpoints of the sequential kernels:
        # ceil(500 / 16) chunks x N x 4 floats + N doubles (+ 64 bytes of alignment slack)
        seq_need = 32 * 100_000 * 16 + 100_000 * 8 + 64
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4.ref, 100_000, 500) == seq_need
        assert lib.smf_hmm_workspace_bytes(_lib.OP_ESTEP, t4.ref, 100_000, 500) == partials + seq_need
        assert lib.smf_hmm_set_option(b"seq", 0) == _lib.SMF_OK
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4.ref, 100_000, 500) == 0
        assert lib.smf_hmm_workspace_bytes(_lib.OP_ESTEP, t4.ref, 100_000, 500) == partials
        assert lib.smf_hmm_set_option(b"seq", -1) == _lib.SMF_OK
        assert lib.smf_hmm_set_option(b"walk", 1) == _lib.SMF_OK
        assert lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, t4.ref, 100, 500) > 0
    finally:
        assert lib.smf_hmm_set_option(b"walk", -1) == _lib.SMF_OK
        assert lib.smf_hmm_set_option(b"short", -1) == _lib.SMF_OK
        assert lib.smf_hmm_set_option(b"seq", -1) == _lib.SMF_OK


def test_every_option_key_is_documented_in_the_header():
    """Each key smf_hmm_set_option accepts (csrc/api.cu) is described in include/smf_hmm.h, and bad values are refused."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    api = open(os.path.join(root, "smftools_b200", "csrc", "api.cu")).read()
    body = api[api.index("int smf_hmm_set_option("):]
    body = body[:body.index("\n}\n")]
    keys = re.findall(r'strcmp\(key, "([a-z_]+)"\)', body)
    assert {"walk", "short", "seq", "vit", "generic", "fitloop"} <= set(keys)
    header = open(os.path.join(root, "include", "smf_hmm.h")).read()
    for k in keys:
        assert f'"{k}"' in header, f"option {k} is not documented in include/smf_hmm.h"
    lib = _lib.load()
    assert lib.smf_hmm_set_option(b"fitloop", 2) == _lib.SMF_ERR_BAD_ARG
    assert lib.smf_hmm_set_option(b"fitloop", 0) == _lib.SMF_OK
    assert lib.smf_hmm_set_option(b"fitloop", 1) == _lib.SMF_OK


def test_log_tables_are_the_reference_expressions():
    m = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.9, 0.07]}, arch="single")
    t = m._host_tables()
    eps = m.eps
    np.testing.assert_array_equal(t.log_start, torch.log(m.start + eps).numpy())
    np.testing.assert_array_equal(t.log_trans[0], torch.log(m.trans + eps).numpy())
    np.testing.assert_array_equal(t.log_emit1[:, 0], torch.log(m.emission + eps).numpy())
    np.testing.assert_array_equal(t.log_emit0[:, 0], torch.log1p(-m.emission + eps).numpy())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    m = S.create_hmm({}, arch="single")
    X = np.zeros((3, 10))
    for call in (lambda: m.decode(X), lambda: m.fit(X, max_iter=1), lambda: m.decode(X, device="cpu")):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    with pytest.raises(ValueError):  # shape errors come first in fit (HMM.py:760-762)
        m.fit(np.zeros((3,)))


def test_binned_bin_index_matches_numpy_digitize():
    m = S.create_hmm({}, arch="single_distance_binned")
    coords = np.array([0, 1, 3, 8, 18, 43, 93, 193, 400])
    assert m._bin_index(coords).tolist() == [0, 1, 1, 2, 3, 4, 5, 6]
