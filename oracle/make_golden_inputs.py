"""Generate golden fixtures for the rows either side of the hot path from the UNMODIFIED
reference functions (container only; run ``python oracle/make_golden_inputs.py`` from the repo root):

* ``build_single_channel`` / ``build_multi_channel_union`` (cli/hmm_adata.py:251-319),
  ``_mask_uncovered_model_input`` / ``_prepare_model_input`` (tools/partitioned_hmm.py:140-188);
* ``_feature_run_lengths`` (tools/partitioned_hmm.py:572-585) and the count / size dictionaries its caller
  builds (:618-630); the fraction counts of ``_plot_feature_fractions`` (:519-529) and the column metric of
  ``call_hmm_peaks`` (hmm/call_hmm_peaks.py:151-160) are inline statements of plotting functions, restated
  here with the same numpy calls;
* ``extract_intervals_from_row`` (analysis/compute/hmm_features.py:15-79).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import DuckAnnData, load_reference_cli  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synthetic_adata(seed, N=24, V=180, with_cov=True, dtype=np.float32, prob=False):
    rng = np.random.default_rng(seed)
    names = np.arange(5000, 5000 + V)
    gpc = rng.random(V) < 0.25
    cpg = rng.random(V) < 0.15
    anyc = gpc | cpg | (rng.random(V) < 0.05)
    a_site = rng.random(V) < 0.3
    var = pd.DataFrame({"refA_GpC_site": gpc, "refA_CpG_site": cpg, "refA_any_C_site": anyc, "refA_A_site": a_site},
                       index=[str(x) for x in names])
    if prob:
        X = (rng.integers(0, 256, size=(N, V)) / 255.0).astype(dtype)
    else:
        X = rng.integers(0, 2, size=(N, V)).astype(dtype)
    X[rng.random((N, V)) < 0.25] = np.nan
    ad = DuckAnnData(N, names, var=var, X=X)
    ad.layers["binarized"] = np.where(np.isnan(X), np.nan, (np.nan_to_num(X) > 0.5)).astype(dtype)
    if with_cov:
        cov = np.ones((N, V), dtype=bool)
        for i in range(N):
            a, b = sorted(rng.integers(0, V, size=2))
            if b - a > 5 and i % 3 == 0:
                cov[i, a + 2:b - 2] = False  # paired-end insert gap
        ad.layers["covered_base_mask"] = cov
    return ad


def main():
    os.makedirs(OUT, exist_ok=True)
    ha, ph = load_reference_cli()
    feats = importlib.import_module("smftools.analysis.compute.hmm_features")
    cfg = types.SimpleNamespace(output_binary_layer_name="binarized", smf_modality="conversion")
    for tag, kw, modality in (("binary_cov", dict(seed=1), "conversion"), ("binary_nocov", dict(seed=2, with_cov=False), "conversion"),
                              ("direct_layer", dict(seed=3, prob=True), "direct"), ("f64_prob", dict(seed=4, dtype=np.float64, prob=True), "conversion")):
        ad = synthetic_adata(**kw)
        cfg.smf_modality = modality
        out = {"X": np.asarray(ad.X), "binarized": ad.layers["binarized"], "var_names": np.asarray(ad.var_names, dtype=str),
               "modality": np.array(modality)}
        for col in ad.var.columns:
            out[f"var_{col}"] = ad.var[col].to_numpy()
        if "covered_base_mask" in ad.layers:
            out["covered"] = ad.layers["covered_base_mask"]
        for mb in ("GpC", "CpG", "C", "A"):
            Xs, coords = ha.build_single_channel(ad, "refA", mb, modality, cfg)
            out[f"single_{mb}_X"] = Xs
            out[f"single_{mb}_coords"] = coords
            out[f"single_{mb}_masked"] = ph._mask_uncovered_model_input(ad, Xs, coords)
            spec = ph.HMMModelSpec(name="t", label=mb, signals=(mb,), feature_groups=("footprint",), architecture="single")
            v, c, pm = ph._prepare_model_input(ad, "refA", spec, cfg)
            assert np.array_equal(c, coords)
            out[f"prepared_{mb}_X"] = v
            out[f"prepared_{mb}_posmask"] = pm
        X3, c3, used = ha.build_multi_channel_union(ad, "refA", ["GpC", "CpG", "A"], modality, cfg)
        out["multi_X"] = X3
        out["multi_coords"] = c3
        out["multi_used"] = np.array(used)
        out["multi_masked"] = ph._mask_uncovered_model_input(ad, X3, c3)
        np.savez_compressed(os.path.join(OUT, f"inputs_{tag}.npz"), **out)

    # ---- reductions / interval extraction on feature-like layers ----
    rng = np.random.default_rng(9)
    N, V = 40, 300
    layer = np.zeros((N, V))
    for i in range(N):
        pos = 0
        while pos < V:
            gap = int(rng.integers(1, 30))
            run = int(rng.integers(1, 25))
            layer[i, pos + gap:pos + gap + run] = 1.0
            pos += gap + run
        a, b = int(rng.integers(0, 40)), int(V - rng.integers(0, 40))
        layer[i, :a] = np.nan
        layer[i, b:] = np.nan
        if i % 5 == 0:
            layer[i, 100:130] = np.nan  # a coverage gap cuts runs
    layer[3, :] = np.nan
    layer[4, :] = np.where(np.isnan(layer[4]), np.nan, 0.0)
    coords = np.sort(rng.choice(np.arange(-2000, 2000), size=V, replace=False)).astype(float)
    out = {"layer": layer, "coords": coords}
    out["modified"] = np.array(float(np.nansum(layer > 0)))  # partitioned_hmm.py:528
    out["valid"] = np.array(int(np.isfinite(layer).sum()))   # :520, 529
    count, size = {}, {}
    for row in layer:
        sizes = ph._feature_run_lengths(row)
        count[int(sizes.size)] = count.get(int(sizes.size), 0) + 1
        for s in sizes:
            size[int(s)] = size.get(int(s), 0) + 1
    out["count_keys"] = np.array(sorted(count))
    out["count_vals"] = np.array([count[k] for k in sorted(count)])
    out["size_keys"] = np.array(sorted(size))
    out["size_vals"] = np.array([size[k] for k in sorted(size)])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        means = np.nan_to_num(np.nanmean(layer, axis=0), nan=0.0)  # call_hmm_peaks.py:151-152
    out["means"] = means
    for w in (1, 4, 9):
        out[f"metric_w{w}"] = np.convolve(means, np.ones(w, dtype=float) / float(w), mode="same") if w > 1 else means
    # extract_intervals_from_row on windows of the rows
    win = slice(60, 220)
    sizes_all, dists_all, sizes_clip, dists_clip = [], [], [], []
    for row in layer:
        s, d = feats.extract_intervals_from_row(row, coords)
        sizes_all.append(s)
        dists_all.append(d)
        s2, d2 = feats.extract_intervals_from_row(row[win], coords[win], full_row=row, full_coords=coords,
                                                  max_mask_overlap_frac=0.5)
        sizes_clip.append(s2)
        dists_clip.append(d2)

    def ragged(lists, dtype):
        return (np.array([len(x) for x in lists]), np.array([v for x in lists for v in x], dtype=dtype))

    out["iv_sizes_n"], out["iv_sizes"] = ragged(sizes_all, np.int64)
    out["iv_dists_n"], out["iv_dists"] = ragged(dists_all, np.float64)
    out["ivc_sizes_n"], out["ivc_sizes"] = ragged(sizes_clip, np.int64)
    out["ivc_dists_n"], out["ivc_dists"] = ragged(dists_clip, np.float64)
    out["window"] = np.array([win.start, win.stop])
    np.savez_compressed(os.path.join(OUT, "reduce_layers.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith(("inputs_", "reduce_"))))


if __name__ == "__main__":
    main()
