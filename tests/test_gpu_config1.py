"""BASELINE config #1 shape, end to end on the GPU against the oracle: default 2-state HMM
(`hmm_init_emission_probs=[[.8,.2],[.2,.8]]`, eps 1e-8, fp64 parameters, max_iter 50, tol 1e-5),
fit + annotate (marginal, threshold 0.5, default feature sets) + merges on an AnnData-like object:
~500 reads x ~2 kb window with GpC sites every ~12 bp (L ~ 170), per (sample, reference) group."""
import numpy as np
import pandas as pd
import pytest

import smftools_b200 as S
from helpers import params_of_model
from oracle import hmm_oracle as O
from oracle.ref_loader import DuckAnnData

pytestmark = pytest.mark.gpu

FEATURE_SETS = {
    "footprint": {"state": "Non-Modified", "features": {
        "small_bound_stretch": [6, 40], "medium_bound_stretch": [40, 100],
        "putative_nucleosome": [100, 200], "large_bound_stretch": [200, "inf"]}},
    "accessible": {"state": "Modified", "features": {
        "small_accessible_patch": [3, 20], "mid_accessible_patch": [20, 40],
        "large_accessible_patch": [40, 110], "nucleosome_depleted_region": [110, "inf"]}},
}
MERGES = (("all_accessible_features", 60), ("all_footprint_features", 10))


def _group(seed, N=500, V=2000):
    rng = np.random.default_rng(seed)
    sites = np.sort(rng.choice(np.arange(5, V - 5), size=V // 12, replace=False)) + 10_000
    full = np.arange(10_000, 10_000 + V)
    L = len(sites)
    X = O.synth_reads(N, L, 2, seed, prob=False, nan_rate=0.1).astype(np.float64)
    starts = full[0] + rng.integers(0, 200, N).astype(float)
    ends = full[-1] - rng.integers(0, 200, N).astype(float)
    obs = pd.DataFrame({"reference_start": starts, "reference_end": ends}, index=[f"r{i}" for i in range(N)])
    return X, sites, full, obs


@pytest.mark.parametrize("seed", [1, 2])
def test_fit_and_annotate_like_experiment_hmm(seed):
    X, sites, full, obs = _group(seed)
    cfg = {"hmm_n_states": 2, "hmm_init_emission_probs": [[0.8, 0.2], [0.2, 0.8]], "hmm_eps": 1e-8}
    m = S.create_hmm(cfg, arch="single")
    hist = m.fit(X, sites, max_iter=50, tol=1e-5)
    p0 = O.init_params(2, [[0.8, 0.2], [0.2, 0.8]])
    q, hist_ref = O.fit(X, p0, sites, max_iter=50, tol=1e-5)
    assert len(hist) == len(hist_ref) and len(hist) < 50
    assert np.allclose(hist, hist_ref, rtol=1e-6)
    sd = m.state_dict()
    assert np.abs(sd["emission"].numpy() - q.emission).max() <= 1e-4
    assert np.abs(sd["trans"].numpy() - q.trans).max() <= 1e-4
    assert np.abs(sd["start"].numpy() - q.start).max() <= 1e-4

    # annotate with the fitted model; the oracle uses the very same parameters
    p = params_of_model(m)
    ad = DuckAnnData(X.shape[0], full, obs=obs)
    var_mask = np.isin(full, sites)
    fsets = S.normalize_hmm_feature_sets(FEATURE_SETS)
    m.annotate_adata(ad, prefix="GpC", X=X, coords=sites, var_mask=var_mask, decode="marginal",
                     feature_sets=fsets, prob_threshold=0.5)
    st, gamma = O.decode(X, p, sites)
    valid = O.span_valid(full, obs["reference_start"], obs["reference_end"])
    cols = np.nonzero(var_mask)[0]
    N, V = X.shape[0], len(full)

    def masked(a):
        out = np.asarray(a).astype(np.float64)
        out[~valid] = np.nan
        return out

    states_layer = np.full((N, V), -1, dtype=np.int8)
    states_layer[:, cols] = st
    # arg-max ties aside (none expected at |gamma - .5| > 2e-5), every layer is identical
    near = np.abs(gamma[:, :, 0] - 0.5) < 2e-5
    ok = ~near.any(axis=1)  # reads without a posterior within 2e-5 of the 0.5 threshold / tie
    assert ok.mean() > 0.98
    real_eq = np.testing.assert_array_equal

    def assert_rows_equal(a, b):
        real_eq(np.asarray(a)[ok], np.asarray(b)[ok])

    np.testing.assert_array_equal = assert_rows_equal  # row-filtered comparison below
    try:
        _compare_layers(ad, m, p, st, gamma, fsets, sites, full, masked, states_layer, cols, N, V, ok)
    finally:
        np.testing.assert_array_equal = real_eq


def _compare_layers(ad, m, p, st, gamma, fsets, sites, full, masked, states_layer, cols, N, V, ok):
    np.testing.assert_array_equal(ad.layers["GpC_states"], masked(states_layer))
    post = np.zeros((N, V), dtype=np.float32)
    t_mod = int(np.argmax(p.emission))
    post[:, cols] = gamma[:, :, t_mod]
    np.testing.assert_allclose(ad.layers["GpC_posterior_modified"], masked(post).astype(np.float32), atol=1e-5,
                               equal_nan=True)
    assert len(ad.uns["hmm_appended_layers"]) == 22
    for group, fs in fsets.items():
        t = int(np.argmax(p.emission)) if fs["state"] == "Modified" else int(np.argmin(p.emission))
        member = gamma[:, :, t] >= 0.5
        ranges = list(fs["features"].values())
        all_l, all_len, cls_l, cls_len = O.call_features(member, sites, full, ranges)
        np.testing.assert_array_equal(ad.layers[f"GpC_all_{group}_features"], masked(all_l))
        np.testing.assert_array_equal(ad.layers[f"GpC_all_{group}_features_lengths"], masked(all_len))
        for ci, feat in enumerate(fs["features"]):
            np.testing.assert_array_equal(ad.layers[f"GpC_{feat}"], masked(cls_l[ci]))
            np.testing.assert_array_equal(ad.layers[f"GpC_{feat}_lengths"], masked(cls_len[ci]))
    # merges as tools/partitioned_hmm.py:_apply_merges does
    for base, dist in MERGES:
        layer = f"GpC_{base}"
        merged_name = m.merge_intervals_to_new_layer(ad, layer, distance_threshold=dist)
        grp = "accessible" if "accessible" in base else "footprint"
        created = m.write_size_class_layers_from_binary(
            ad, merged_name, out_prefix="GpC", feature_ranges=fsets[grp]["features"], suffix="_merged")
        mg, ml = O.merge_runs(ad.layers[layer], full, dist)
        np.testing.assert_array_equal(ad.layers[merged_name], mg)
        np.testing.assert_array_equal(ad.layers[merged_name + "_lengths"], ml)
        cl, cll = O.size_classes_from_binary(mg, full, list(fsets[grp]["features"].values()))
        for ci, feat in enumerate(fsets[grp]["features"]):
            np.testing.assert_array_equal(ad.layers[f"GpC_{feat}_merged"], cl[ci])
            np.testing.assert_array_equal(ad.layers[f"GpC_{feat}_merged_lengths"], cll[ci])
        assert len(created) == 2 * len(fsets[grp]["features"])
