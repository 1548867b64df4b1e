"""GPU parity of the rows either side of the hot path: device input builders, layer
reductions from the packed results, interval records -> per-read interval statistics; against the
reference-generated fixtures and the CPU oracle."""
import types

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import hmm_oracle as O
from oracle import inputs_oracle as IO
from oracle.ref_loader import DuckAnnData

pytestmark = pytest.mark.gpu


def _adata_from_golden(g):
    names = [int(x) for x in g["var_names"]]
    var = pd.DataFrame({k[4:]: g[k] for k in g if k.startswith("var_")}, index=[str(x) for x in names])
    ad = DuckAnnData(g["X"].shape[0], names, var=var, X=g["X"])
    ad.layers["binarized"] = g["binarized"]
    if "covered" in g:
        ad.layers["covered_base_mask"] = g["covered"]
    return ad


def _to_float(values):
    """Builder output (CodedU8 / device tensor) -> float64 host matrix with NaN = missing."""
    import smftools_b200 as S

    if isinstance(values, S.CodedU8):
        code = values.data.cpu().numpy()
        out = code.astype(np.float64)
        out[code > 1] = np.nan
        return out
    return values.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("name", golden_names("inputs_"))
def test_device_input_builders_match_reference_fixtures(name):
    import smftools_b200 as S

    g = load_golden(name)
    ad = _adata_from_golden(g)
    modality = str(g["modality"])
    cfg = types.SimpleNamespace(output_binary_layer_name="binarized", smf_modality=modality)
    binary = np.all(np.isnan(g["single_GpC_X"]) | (g["single_GpC_X"] == 0) | (g["single_GpC_X"] == 1))
    for mb in ("GpC", "CpG", "C", "A"):
        X, coords = S.build_single_channel(ad, "refA", mb, modality, cfg)
        assert isinstance(X, S.CodedU8) == bool(binary), "binary calls must come back as the uint8 code"
        np.testing.assert_array_equal(coords, g[f"single_{mb}_coords"])
        np.testing.assert_array_equal(_to_float(X), g[f"single_{mb}_X"])
        np.testing.assert_array_equal(_to_float(S.mask_uncovered_model_input(ad, X, coords)), g[f"single_{mb}_masked"])
        spec = types.SimpleNamespace(signals=(mb,))
        v, c, pm = S.prepare_model_input(ad, "refA", spec, cfg)
        np.testing.assert_array_equal(_to_float(v), g[f"prepared_{mb}_X"])
        np.testing.assert_array_equal(c, coords)
        np.testing.assert_array_equal(pm, g[f"prepared_{mb}_posmask"])
    X3, c3, used = S.build_multi_channel_union(ad, "refA", ["GpC", "CpG", "A"], modality, cfg)
    np.testing.assert_array_equal(_to_float(X3), g["multi_X"])
    np.testing.assert_array_equal(c3, g["multi_coords"])
    assert used == [str(x) for x in g["multi_used"]]
    np.testing.assert_array_equal(_to_float(S.mask_uncovered_model_input(ad, X3, c3)), g["multi_masked"])
    v3, _, pm3 = S.prepare_model_input(ad, "refA", types.SimpleNamespace(signals=("GpC", "CpG", "A")), cfg)
    np.testing.assert_array_equal(_to_float(v3), g["multi_masked"])
    # the reference's error conventions
    with pytest.raises(ValueError):
        S.build_single_channel(ad, "refA", "nonexistent", modality, cfg)
    with pytest.raises(ValueError):
        S.build_multi_channel_union(ad, "refA", ["GpC", "nonexistent"], modality, cfg)
    if modality == "direct":
        with pytest.raises(KeyError):
            S.build_single_channel(ad, "refA", "GpC", modality, types.SimpleNamespace(output_binary_layer_name="nope"))


def test_built_input_feeds_fit_and_decode_like_the_float_matrix():
    """prepare_model_input -> model.fit / decode (device-resident uint8 code) == the reference-style float64
    host matrix through the same calls."""
    import smftools_b200 as S

    g = load_golden("inputs_binary_cov")
    ad = _adata_from_golden(g)
    cfg = types.SimpleNamespace(output_binary_layer_name="binarized", smf_modality="conversion")
    v, coords, _ = S.prepare_model_input(ad, "refA", types.SimpleNamespace(signals=("GpC",)), cfg)
    Xf = g["prepared_GpC_X"]
    m1 = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    m2 = S.create_hmm({"hmm_n_states": 2, "hmm_init_emission_probs": [0.8, 0.2]}, arch="single")
    h1 = m1.fit(v, coords, max_iter=4, tol=0.0)
    h2 = m2.fit(Xf, coords, max_iter=4, tol=0.0)
    np.testing.assert_allclose(h1, h2, rtol=1e-6)  # uint8 and float64 inputs run different instantiations
    s1, g1 = m1.decode(v, coords)
    s2, g2 = m2.decode(Xf, coords)
    assert np.abs(g1 - g2).max() <= 2e-6
    ref_g, _ = O.posterior(Xf, O.Params(m2.start.detach().cpu().numpy(), m2.trans.detach().cpu().numpy(),
                                        m2.emission.detach().cpu().numpy()))
    assert np.abs(g1 - ref_g).max() <= 1e-5


def _packed_from_layer(layer, rng, n_classes=3):
    """class_out + valid reproducing a float layer: runs get random classes (0 = unclassified is not used)."""
    valid = np.isfinite(layer)
    member = np.nan_to_num(layer, nan=0.0) > 0
    cls = np.zeros(layer.shape, dtype=np.uint8)
    for i, row in enumerate(member):
        b = np.concatenate(([False], row, [False]))
        for s, e in zip(np.flatnonzero(~b[:-1] & b[1:]), np.flatnonzero(b[:-1] & ~b[1:])):
            cls[i, s:e] = rng.integers(1, n_classes + 1)
    return cls, valid.astype(np.uint8)


def test_layer_reductions_match_reference_fixtures():
    from smftools_b200 import reductions as R

    g = load_golden("reduce_layers")
    layer = g["layer"]
    rng = np.random.default_rng(2)
    cls_h, valid_h = _packed_from_layer(layer, rng)
    dev = torch.device("cuda")
    cls = torch.from_numpy(cls_h).to(dev)
    valid = torch.from_numpy(valid_h).to(dev)
    names = ["GpC_all_footprint_features", "c0", "c1", "c2"]
    fr = R.feature_fraction_counts(cls, names, valid=valid)
    assert fr[names[0]]["modified"] == float(g["modified"]) and fr[names[0]]["valid"] == int(g["valid"])
    for c in range(3):
        mod, val = IO.layer_fraction_counts(np.where(valid_h > 0, (cls_h == c + 1).astype(float), np.nan))
        assert fr[names[c + 1]]["modified"] == mod and fr[names[c + 1]]["valid"] == val
    count, size = R.run_histograms(cls, valid=valid)
    assert sorted(count) == g["count_keys"].tolist() and [count[k] for k in sorted(count)] == g["count_vals"].tolist()
    assert sorted(size) == g["size_keys"].tolist() and [size[k] for k in sorted(size)] == g["size_vals"].tolist()
    for c in range(3):
        cc, ss = R.run_histograms(cls, select_class=c, valid=valid)
        oc, os_ = IO.run_histograms(np.where(valid_h > 0, (cls_h == c + 1).astype(float), np.nan))
        assert cc == oc and ss == os_
    colsum, colvalid = R.layer_column_sums(cls, n_classes=3, valid=valid)
    for w in (1, 4, 9):
        means, metric = R.column_mean_metric(colsum[0], colvalid, w)
        np.testing.assert_allclose(means.cpu().numpy(), g["means"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(metric.cpu().numpy(), g[f"metric_w{w}"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("N,V", [(1, 1), (3, 31), (5, 32), (7, 33), (64, 1000), (300, 4097)])
def test_reductions_random_shapes_and_packed_codes(N, V):
    from smftools_b200 import reductions as R

    rng = np.random.default_rng(N * 1000 + V)
    member = rng.random((N, V)) < 0.55
    member[:, ::7] &= rng.random((N, (V + 6) // 7)) < 0.3
    cls_of = rng.integers(0, 4, size=(N, V)).astype(np.uint8)
    valid = rng.random((N, V)) < 0.9
    if N > 2:
        member[1] = True
        valid[1] = True
        member[2] = False
    code = (np.where(member, 0x40, 0) | np.where(member, cls_of, 0) | np.where(valid, 0x80, 0)).astype(np.uint8)
    dev = torch.device("cuda")
    code_d = torch.from_numpy(code).to(dev)
    layer = np.where(valid, member.astype(float), np.nan)
    fr = R.feature_fraction_counts(code_d, ["merged", "a", "b", "c"], packed=True)
    mod, val = IO.layer_fraction_counts(layer)
    assert fr["merged"]["modified"] == mod and fr["merged"]["valid"] == val
    cc, ss = R.run_histograms(code_d, packed=True)
    oc, os_ = IO.run_histograms(layer)
    assert cc == oc and ss == os_
    for c in range(3):
        lay_c = np.where(valid, (member & (cls_of == c + 1)).astype(float), np.nan)
        cc, ss = R.run_histograms(code_d, select_class=c, packed=True)
        oc, os_ = IO.run_histograms(lay_c)
        assert cc == oc and ss == os_
        assert fr["abc"[c]]["modified"] == IO.layer_fraction_counts(lay_c)[0]
    colsum, colvalid = R.layer_column_sums(code_d, n_classes=3, packed=True)
    _, metric = R.column_mean_metric(colsum[0], colvalid, 5)
    np.testing.assert_allclose(metric.cpu().numpy(), IO.column_mean_metric(layer, 5), rtol=0, atol=1e-12)


def test_interval_records_feed_the_interval_consumer():
    """call_features interval records (sorted) -> intervals_by_read == the reference's per-row extraction on
    the dense aggregate layer the same kernel writes."""
    from helpers import model_from_params
    from smftools_b200 import engine
    from smftools_b200 import reductions as R

    K, N, L = 2, 37, 400
    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X32 = O.synth_reads(N, L, K, 12, prob=False, nan_rate=0.3)
    sv = engine.viterbi(tables, torch.from_numpy(X32).to(dev), None)
    coords = O.synth_coords(L, 4, span_factor=5) + 100
    full = np.arange(coords[0] - 3, coords[-1] + 4)
    left = np.searchsorted(full, coords, side="left").astype(np.int32)
    right = np.searchsorted(full, coords, side="right").astype(np.int32)
    ranges = [(1, 30), (30, 90), (90, float("inf"))]
    args = dict(states=sv, post=None, target=1, threshold=0.0, coords=torch.from_numpy(coords.astype(np.int64)).to(dev),
                grid_left=torch.from_numpy(left).to(dev), grid_right=torch.from_numpy(right).to(dev), n_grid=len(full),
                ranges=ranges)
    cls, rl, iv = engine.call_features(**args, max_intervals=N * L)
    rec = R.sort_interval_records(iv)
    dense = (cls.cpu().numpy() > 0).astype(float)
    sizes, dists = R.intervals_by_read(rec, full.astype(float), N)
    win = (40, len(full) - 55)
    sizes_w, dists_w = R.intervals_by_read(rec, full.astype(float), N, window=win)
    for i in range(N):
        s_ref, d_ref = IO.extract_intervals_from_row(dense[i], full.astype(float))
        assert sizes[i] == s_ref and dists[i] == d_ref
        s_ref, d_ref = IO.extract_intervals_from_row(dense[i, win[0]:win[1]], full[win[0]:win[1]].astype(float),
                                                     full_row=dense[i], full_coords=full.astype(float))
        assert sizes_w[i] == s_ref and dists_w[i] == d_ref
    # read-span clipping of the records == NaN-masking the dense rows
    first = np.random.default_rng(1).integers(0, 60, size=N)
    last = len(full) - 1 - np.random.default_rng(2).integers(0, 60, size=N)
    rec_c = R.clip_records_to_span(rec, first, last)
    sizes_c, _ = R.intervals_by_read(rec_c, full.astype(float), N)
    for i in range(N):
        row = dense[i].copy()
        row[:first[i]] = np.nan
        row[last[i] + 1:] = np.nan
        assert sizes_c[i] == IO.extract_intervals_from_row(row, full.astype(float))[0]


def test_fit_recodes_binary_float_observations_once():
    """fit() of 0 / 1 / NaN floats (the reference's dtype) re-codes them as the uint8 code for all iterations;
    probabilistic calls stay floats; the fitted model equals the oracle's either way."""
    import smftools_b200 as S
    from helpers import params_of_model
    from oracle import hmm_oracle as O

    dev = torch.device("cuda")
    for C, arch in ((1, "single"), (2, "multi")):
        Xb = O.synth_reads(120, 1500, 3, 41, prob=False, nan_rate=0.3, n_channels=C)
        Xp = O.synth_reads(120, 1500, 3, 42, prob=True, nan_rate=0.3, n_channels=C)
        res_b = S.BaseHMM._code_binary_resident([torch.from_numpy(Xb.astype(np.float64)).to(dev)])
        res_p = S.BaseHMM._code_binary_resident([torch.from_numpy(Xp.astype(np.float32)).to(dev)])
        assert res_b[0].dtype == torch.uint8 and tuple(res_b[0].shape) == Xb.shape
        assert np.array_equal(res_b[0].cpu().numpy(), np.where(np.isnan(Xb), 255, Xb).astype(np.uint8))
        assert res_p[0].dtype == torch.float32
        cfg = {"hmm_n_states": 3}
        if C > 1:
            cfg["hmm_n_channels"] = C
        for X in (Xb, Xp):
            m = S.create_hmm(cfg, arch=arch)
            p0 = params_of_model(m)
            hist = m.fit(X.astype(np.float64), max_iter=4, tol=0.0)
            q, h_ref = O.fit(X.astype(np.float64), p0, max_iter=4, tol=0.0)
            np.testing.assert_allclose(hist, h_ref, rtol=1e-6)
            assert np.abs(params_of_model(m).emission - q.emission).max() <= 1e-4
