"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container only).

    python oracle/make_golden.py            # needs /root/reference

Each file stores the inputs and what the reference's own ``smftools.hmm.HMM`` classes return
for them on CPU/float64: posteriors, marginal and Viterbi paths, per-iteration log-likelihoods
and fitted parameters, and the full ``annotate_adata`` / merge / size-class / read-span layer
sets on a duck-typed AnnData.  ``tests/test_oracle_golden.py`` pins ``oracle/hmm_oracle.py``
against them; the GPU parity tests use them as committed fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import hmm_oracle as O  # noqa: E402
from oracle.ref_loader import DuckAnnData, load_reference_hmm  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FEATURE_SETS = {
    "footprint": {"state": "Non-Modified", "features": {
        "small_bound_stretch": [6, 40], "medium_bound_stretch": [40, 100],
        "putative_nucleosome": [100, 200], "large_bound_stretch": [200, "inf"]}},
    "accessible": {"state": "Modified", "features": {
        "small_accessible_patch": [3, 20], "mid_accessible_patch": [20, 40],
        "large_accessible_patch": [40, 110], "nucleosome_depleted_region": [110, "inf"]}},
}


def set_params(model, p: O.Params):
    with torch.no_grad():
        model.start.data = torch.tensor(p.start, dtype=torch.float64)
        model.trans.data = torch.tensor(p.trans, dtype=torch.float64)
        model.emission.data = torch.tensor(p.emission, dtype=torch.float64).reshape(model.emission.shape)
        if p.trans_by_bin is not None:
            model.trans_by_bin.data = torch.tensor(p.trans_by_bin, dtype=torch.float64)


def get_params(model) -> dict:
    d = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    return d


def case_core(H, name, arch, K, N, L, seed, prob, cfg_extra=None, fitted=True, C=1, fit_iters=4):
    cfg = {"hmm_n_states": K}
    if arch == "multi":
        cfg["hmm_n_channels"] = C
    cfg.update(cfg_extra or {})
    X32 = O.synth_reads(N, L, K, seed, prob=prob, nan_rate=0.3, n_channels=C)
    X = X32.astype(np.float64)
    coords = O.synth_coords(L, seed) if arch == "single_distance_binned" else np.arange(L)
    model = H.create_hmm(cfg, arch=arch, device="cpu")
    if fitted:
        p = O.synth_params(K, arch=arch, n_channels=C)
        set_params(model, p)
    out = {"X": X32, "coords": coords, "arch": np.array(arch), "K": np.array(K)}
    for k, v in get_params(model).items():
        out[f"p0_{k}"] = v
    out["eps"] = np.array(model.eps)
    st_m, gamma = model.decode(X, coords, decode="marginal")
    st_v, _ = model.decode(X, coords, decode="viterbi")
    out["gamma"] = gamma
    out["states_marginal"] = st_m.astype(np.int8)
    out["states_viterbi"] = st_v.astype(np.int8)
    # per-read log-likelihood from the reference's forward pass
    obs = torch.tensor(np.nan_to_num(X, nan=0.0), dtype=torch.float64)
    mask = torch.tensor(~np.isnan(X))
    if arch == "single_distance_binned":
        alpha = model._alpha_beta_by_bin(obs, mask, coords=coords)[0]
    else:
        alpha = model._alpha_beta(obs, mask)[0]
    out["loglik"] = torch.logsumexp(alpha[:, -1, :], dim=1).numpy()
    hist = model.fit(X, coords, max_iter=fit_iters, tol=0.0, device="cpu")
    out["fit_hist"] = np.asarray(hist)
    for k, v in get_params(model).items():
        out[f"p1_{k}"] = v
    # frozen-transition adaptation (HMMTrainer.adapt_or_load path)
    hist2 = model.adapt_emissions(X, coords, iters=2, device="cpu")
    out["adapt_hist"] = np.asarray(hist2)
    for k, v in get_params(model).items():
        out[f"p2_{k}"] = v
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "gamma", gamma.shape, "hist", hist[:2], "...")


def case_annotate(H, name, K, decode, seed, N=10, L=70, prob=False, with_covered=False, state_ints=False):
    import pandas as pd

    X32 = O.synth_reads(N, L, K, seed, prob=prob, nan_rate=0.25)
    X = X32.astype(np.float64)
    site_coords = O.synth_coords(L, seed, span_factor=6) + 1000
    V = int(site_coords[-1] - site_coords[0] + 1 + 20)
    full = np.arange(site_coords[0] - 10, site_coords[0] - 10 + V)
    var_mask = np.isin(full, site_coords)
    rng = np.random.default_rng(seed)
    starts = full[0] + rng.integers(0, 40, size=N).astype(float)
    ends = full[-1] - rng.integers(0, 40, size=N).astype(float)
    starts[0] = np.nan  # non-finite -> row left unmasked
    obs_df = pd.DataFrame({"reference_start": starts, "reference_end": ends}, index=[f"r{i}" for i in range(N)])
    ad = DuckAnnData(N, full, obs=obs_df)
    if with_covered:
        cov = rng.random((N, V)) > 0.1
        ad.layers["covered_base_mask"] = cov
    model = H.create_hmm({"hmm_n_states": K}, arch="single", device="cpu")
    set_params(model, O.synth_params(K))
    fsets = H.normalize_hmm_feature_sets(FEATURE_SETS)
    if state_ints:
        fsets = H.normalize_hmm_feature_sets({
            "footprint": {"state": 1, "features": FEATURE_SETS["footprint"]["features"]},
            "tf": {"state": 2, "features": {"tf_small": [1, 30], "tf_big": [30, "inf"]}},
            "accessible": {"state": 0, "features": FEATURE_SETS["accessible"]["features"]},
        })
    model.annotate_adata(ad, prefix="GpC", X=X, coords=site_coords, var_mask=var_mask, decode=decode,
                         feature_sets=fsets, prob_threshold=0.5, device="cpu")
    out = {"X": X32, "coords": site_coords, "full": full, "var_mask": var_mask, "starts": starts, "ends": ends,
           "K": np.array(K), "decode": np.array(decode), "state_ints": np.array(state_ints)}
    if with_covered:
        out["covered"] = cov
    for k, v in get_params(model).items():
        out[f"p0_{k}"] = v
    appended = list(ad.uns["hmm_appended_layers"])
    out["appended"] = np.array(appended)
    for nm in appended:
        out[f"layer__{nm}"] = np.asarray(ad.layers[nm])
    # merges (tools/partitioned_hmm.py:78-108 chain): merge then size classes
    merged = model.merge_intervals_to_new_layer(ad, "GpC_all_footprint_features", distance_threshold=10)
    created = model.write_size_class_layers_from_binary(
        ad, merged, out_prefix="GpC", feature_ranges=fsets["footprint"]["features"], suffix="_merged")
    for nm in [merged, merged + "_lengths"] + created:
        out[f"merge__{nm}"] = np.asarray(ad.layers[nm])
    out["merge_names"] = np.array([merged, merged + "_lengths"] + created)
    # the complete _apply_merges chain (tools/partitioned_hmm.py:78-108) for the default merge pairs
    # (config/default.yaml hmm_merge_layer_features): merge -> size classes -> read-span NaN mask
    chain = []
    for core, dist in (("all_footprint_features", 10), ("all_accessible_features", 60)):
        base = f"GpC_{core}"
        if base not in ad.layers:
            continue
        mb = model.merge_intervals_to_new_layer(ad, base, distance_threshold=dist, suffix="_merged", overwrite=True)
        masked = [mb, f"{mb}_lengths"]
        ranges = {}
        for gname, fs in fsets.items():
            if core == f"all_{gname}_features":
                ranges = fs.get("features", {}) or {}
        if ranges:
            masked.extend(model.write_size_class_layers_from_binary(
                ad, mb, out_prefix="GpC", feature_ranges=ranges, suffix="_merged", overwrite=True))
        H.mask_layers_outside_read_span(ad, masked, use_original_var_names=True)
        chain.extend(masked)
    for nm in chain:
        out[f"chain__{nm}"] = np.asarray(ad.layers[nm])
    out["chain_names"] = np.array(chain)
    out["chain_appended"] = np.array(list(ad.uns["hmm_appended_layers"]))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, len(appended), "layers", {str(np.asarray(ad.layers[a]).dtype) for a in appended})


def main():
    H = load_reference_hmm()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    case_core(H, "single_k2_binary_default", "single", 2, 24, 64, 11, False, fitted=False)
    case_core(H, "single_k2_binary_fitted", "single", 2, 24, 257, 12, False)
    case_core(H, "single_k3_prob", "single", 3, 16, 300, 13, True,
              cfg_extra={"hmm_init_emission_probs": [0.85, 0.4, 0.05]})
    case_core(H, "single_k4_binary", "single", 4, 16, 200, 14, False)
    case_core(H, "multi_k2_c2", "multi", 2, 12, 90, 15, False, C=2)
    case_core(H, "multi_k3_c3_prob", "multi", 3, 10, 120, 16, True, C=3)
    case_core(H, "binned_k2", "single_distance_binned", 2, 16, 150, 17, False)
    case_core(H, "binned_k3_prob", "single_distance_binned", 3, 12, 130, 18, True)
    case_core(H, "single_k2_len1", "single", 2, 5, 1, 19, False)
    case_core(H, "single_k2_len33", "single", 2, 7, 33, 20, False)
    case_annotate(H, "annotate_k2_marginal", 2, "marginal", 31)
    case_annotate(H, "annotate_k2_viterbi_covered", 2, "viterbi", 32, with_covered=True)
    case_annotate(H, "annotate_k3_viterbi_ints", 3, "viterbi", 33, prob=True, state_ints=True, L=120)


if __name__ == "__main__":
    main()
