"""The reference caller's checkpoint round trip with the B200 classes swapped into its registry.

``HMMTrainer`` (cli/hmm_adata.py:488-524): ``_payload(model, hist)`` -> ``atomic_torch_save`` ->
``torch.load`` -> ``create_hmm(cfg, arch, override=..., device=...)`` -> dtype-coerced ``load_state_dict`` ->
``.to(device)`` -> ``.eval()`` -> ``annotate_adata``.

* CPU (container, reference present): the *real* ``HMMTrainer._payload / _save / _load`` run against
  ``install_into_reference`` -- no kernel is needed up to the loaded model.
* GPU (B200 box, no reference there): the same call sequence restated line by line, ending in
  ``annotate_adata``; the reloaded model must write exactly the layers the fitted model writes.
"""
import copy
import os
import types

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import hmm_oracle as O
from oracle.ref_loader import DuckAnnData, reference_available


def _fitted_like(S, arch, K=2, dtype="float32"):
    cfg = {"hmm_n_states": K, "hmm_dtype": dtype, "hmm_n_channels": 2, "hmm_distance_bins": [2, 6, 20]}
    m = S.create_hmm(cfg, arch=arch)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        m.start.data = torch.rand(m.start.shape, generator=g, dtype=torch.float64).to(m.start.dtype)
        m.trans.data = torch.rand(m.trans.shape, generator=g, dtype=torch.float64).to(m.trans.dtype)
        m.emission.data = torch.rand(m.emission.shape, generator=g, dtype=torch.float64).to(m.emission.dtype)
        if hasattr(m, "trans_by_bin"):
            m.trans_by_bin.data = torch.rand(m.trans_by_bin.shape, generator=g, dtype=torch.float64).to(m.trans_by_bin.dtype)
    m._normalize_all()
    return m, cfg


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("arch", ["single", "multi", "single_distance_binned"])
def test_reference_hmmtrainer_payload_save_load_with_swapped_registry(arch, tmp_path, monkeypatch):
    import pickle

    import smftools_b200 as S
    from oracle.ref_loader import load_reference_cli, load_reference_hmm

    ref = load_reference_hmm()
    saved_registry = dict(ref._HMM_REGISTRY)
    try:
        S.install_into_reference(ref)
        hmm_adata, _ = load_reference_cli()
        model, cfg_dict = _fitted_like(S, arch, K=3, dtype="float32")
        cfg_dict["hmm_dtype"] = "float64"  # the loading side asks for another dtype: _load coerces (:516-519)
        cfg = types.SimpleNamespace(**cfg_dict, force_redo_hmm_fit=False)
        trainer = hmm_adata.HMMTrainer(cfg=cfg, models_dir=tmp_path)
        assert trainer.choose_arch(multichannel=False) == "single"
        payload = trainer._payload(model, hist=[-10.0, -9.5])
        assert payload["hmm_arch"] == arch and payload["fit_history"] == [-10.0, -9.5]
        if arch == "multi":
            assert payload["override"] == {"hmm_n_channels": 2}
        if arch == "single_distance_binned":
            assert payload["override"] == {"hmm_distance_bins": [2, 6, 20]}
        path = tmp_path / "ckpt.pt"
        trainer._save(model, path, hist=[-10.0, -9.5])
        try:
            loaded = trainer._load(path, arch, torch.device("cpu"))
        except pickle.UnpicklingError:
            # torch >= 2.6 loads with weights_only=True and rejects the numpy integers that
            # `list(model.distance_bins)` puts into the override (cli/hmm_adata.py:493-496).  That is the
            # reference's own behaviour -- its class keeps distance_bins as the same numpy array
            # (HMM.py:2015-2017) -- so the payload is identical; finish the round trip the way older torch did.
            assert arch == "single_distance_binned"
            ref_model = saved_registry[arch].from_config(cfg_dict)
            assert [type(v) for v in trainer._payload(ref_model)["override"]["hmm_distance_bins"]] == \
                   [type(v) for v in payload["override"]["hmm_distance_bins"]]
            real_load = torch.load
            monkeypatch.setattr(torch, "load", lambda *a, **k: real_load(*a, **{**k, "weights_only": False}))
            loaded = trainer._load(path, arch, torch.device("cpu"))
        assert type(loaded) is type(model) and type(loaded).__module__.startswith("smftools_b200")
        assert next(loaded.parameters()).dtype == torch.float64
        sd0, sd1 = model.state_dict(), loaded.state_dict()
        assert list(sd0) == list(sd1)
        for k in sd0:
            np.testing.assert_array_equal(sd0[k].to(torch.float64).numpy(), sd1[k].numpy())
        assert copy.deepcopy(loaded).state_dict().keys() == sd1.keys()  # cli/hmm_adata.py:602, 705
    finally:
        ref._HMM_REGISTRY.clear()
        ref._HMM_REGISTRY.update(saved_registry)


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["single", "single_distance_binned"])
def test_trainer_call_sequence_ends_in_identical_annotation(arch, tmp_path):
    import smftools_b200 as S

    K, N, L = 2, 60, 120
    X = O.synth_reads(N, L, K, 11, prob=False, nan_rate=0.3).astype(np.float64)
    coords = O.synth_coords(L, 3, span_factor=6) + 500
    full = np.arange(coords[0] - 5, coords[-1] + 6)
    var_mask = np.isin(full, coords)
    obs = pd.DataFrame({"reference_start": np.full(N, float(full[3])), "reference_end": np.full(N, float(full[-4]))},
                       index=[f"r{i}" for i in range(N)])
    cfg = {"hmm_n_states": K, "hmm_init_emission_probs": [0.8, 0.2], "hmm_distance_bins": [1, 5, 10, 25, 50, 100]}
    fsets = S.normalize_hmm_feature_sets({
        "footprint": {"state": "Non-Modified", "features": {"small": [3, 40], "large": [40, "inf"]}},
        "accessible": {"state": "Modified", "features": {"patch": [1, "inf"]}}})
    model = S.create_hmm(cfg, arch=arch)
    hist = model.fit(X, coords, max_iter=3, tol=0.0)
    # HMMTrainer._payload (cli/hmm_adata.py:488-504)
    override = {}
    if getattr(model, "hmm_name", None) == "single_distance_binned":
        override["hmm_distance_bins"] = list(getattr(model, "distance_bins", [1, 5, 10, 25, 50, 100]))
    payload = {"state_dict": model.state_dict(), "hmm_arch": getattr(model, "hmm_name", None), "override": override,
               "fit_history": [float(v) for v in hist]}
    path = os.path.join(tmp_path, "PER_s_ref_GpC.pt")
    torch.save(payload, path)
    # HMMTrainer._load (:510-524)
    # (torch >= 2.6 defaults to weights_only=True, which rejects the numpy integers of the override list -- with
    # the reference's own class too, see the CPU test above; load the way the reference's torch did)
    payload = torch.load(path, map_location="cpu", weights_only=False)
    device = torch.device("cuda")
    m = S.create_hmm(cfg, arch=payload["hmm_arch"], override=payload.get("override", None), device=device)
    sd = payload["state_dict"]
    target_dtype = next(m.parameters()).dtype
    for k, v in list(sd.items()):
        if isinstance(v, torch.Tensor) and v.dtype != target_dtype:
            sd[k] = v.to(dtype=target_dtype)
    m.load_state_dict(sd)
    m.to(device)
    m.eval()
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), m.state_dict()[k].cpu())
    out = []
    for mod in (model, m):
        ad = DuckAnnData(N, full, obs=obs.copy())
        mod.annotate_adata(ad, prefix="GpC", X=X, coords=coords, var_mask=var_mask, feature_sets=fsets,
                           decode="viterbi", device=device)
        out.append(ad)
    assert list(out[0].layers) == list(out[1].layers) and out[0].uns["hmm_appended_layers"] == out[1].uns["hmm_appended_layers"]
    for nm in out[0].layers:
        np.testing.assert_array_equal(np.asarray(out[0].layers[nm]), np.asarray(out[1].layers[nm]))
