"""Container-only cross-checks against the live, unmodified reference (skipped on the GPU box,
where /root/reference does not exist): oracle == reference on fresh random inputs, and the host
mirror's config / feature-set handling == the reference's."""
import numpy as np
import pytest
import torch

from oracle import hmm_oracle as O
from oracle.ref_loader import load_reference_hmm, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def _set(model, p):
    with torch.no_grad():
        model.start.data = torch.tensor(p.start)
        model.trans.data = torch.tensor(p.trans)
        model.emission.data = torch.tensor(p.emission).reshape(model.emission.shape)
        if p.trans_by_bin is not None:
            model.trans_by_bin.data = torch.tensor(p.trans_by_bin)


@pytest.mark.parametrize("arch,K,C,prob", [("single", 2, 1, False), ("single", 4, 1, True), ("multi", 3, 2, True),
                                           ("single_distance_binned", 3, 1, False)])
def test_oracle_equals_reference_on_random_inputs(arch, K, C, prob):
    H = load_reference_hmm()
    X = O.synth_reads(20, 180, K, 900 + K, prob=prob, n_channels=C).astype(np.float64)
    coords = O.synth_coords(180, 5) if arch == "single_distance_binned" else np.arange(180)
    p = O.synth_params(K, arch=arch, n_channels=C)
    cfg = {"hmm_n_states": K, "hmm_n_channels": C}
    m = H.create_hmm(cfg, arch=arch, device="cpu")
    _set(m, p)
    st_ref, g_ref = m.decode(X, coords, decode="viterbi")
    st, g = O.decode(X, p, coords, "viterbi")
    assert np.abs(g - g_ref).max() < 1e-11
    _, margin = O.viterbi(X, p, coords, return_margins=True)
    assert np.all((st == st_ref).all(axis=1) | (margin < 1e-9))
    hist_ref = m.fit(X, coords, max_iter=3, tol=0.0, device="cpu")
    q, hist = O.fit(X, p, coords, max_iter=3, tol=0.0)
    assert np.allclose(hist, hist_ref, rtol=1e-11)
    assert np.abs(q.emission.reshape(-1) - m.emission.detach().numpy().reshape(-1)).max() < 1e-10


def test_host_mirror_config_handling_equals_reference():
    import smftools_b200 as S

    H = load_reference_hmm()
    cfgs = [None, {}, {"hmm_n_states": 3, "hmm_init_emission_probs": [[0.8, 0.2], [0.2, 0.8]], "hmm_eps": 1e-6},
            {"hmm_n_states": 2, "hmm_init_emission": [0.7, 0.1]}, {"hmm_dtype": "float32", "hmm_n_channels": 3}]
    for cfg in cfgs:
        for arch in ("single", "multi", "single_distance_binned"):
            a = S.create_hmm(cfg, arch=arch)
            b = H.create_hmm(cfg, arch=arch, device="cpu")
            assert a.eps == b.eps and a.n_states == b.n_states and a.dtype == b.dtype
            for k, v in b.state_dict().items():
                assert torch.equal(a.state_dict()[k], v), (cfg, arch, k)
            assert a.modified_state_index() == b.modified_state_index()
    raws = ['{"a": {"state": 1, "features": {"x": [1, "inf"], "y": null}}}', {"g": {"ranges": {"q": ["none", 5]}}},
            {"g": [3, 9]}, "junk", None]
    for raw in raws:
        assert S.normalize_hmm_feature_sets(raw) == H.normalize_hmm_feature_sets(raw)
