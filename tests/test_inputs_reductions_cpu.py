"""CPU checks of the rows either side of the hot path: the numpy restatement
(oracle/inputs_oracle.py) against fixtures generated from the unmodified reference functions
(oracle/make_golden_inputs.py), live against the reference when its tree is present, and the host-side
interval consumer (smftools_b200.reductions.intervals_by_read) against the reference's
``extract_intervals_from_row`` results."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import inputs_oracle as IO
from oracle.ref_loader import reference_available

METHBASES = ("GpC", "CpG", "C", "A")


def _matrix(g):
    return g["binarized"] if str(g["modality"]) == "direct" else g["X"]


def _mask(g, mb):
    col = IO.pos_mask_column([k[4:] for k in g if k.startswith("var_")], "refA", mb)
    return g[f"var_{col}"].astype(bool)


@pytest.mark.parametrize("name", golden_names("inputs_"))
def test_input_builder_oracle_matches_reference_fixtures(name):
    g = load_golden(name)
    mat, names = _matrix(g), g["var_names"]
    cov = g["covered"] if "covered" in g else None
    for mb in METHBASES:
        X, coords = IO.build_single_channel(mat, names, _mask(g, mb))
        np.testing.assert_array_equal(X, g[f"single_{mb}_X"])
        np.testing.assert_array_equal(coords, g[f"single_{mb}_coords"])
        np.testing.assert_array_equal(IO.mask_uncovered(X, coords, names, cov), g[f"single_{mb}_masked"])
        np.testing.assert_array_equal(g[f"prepared_{mb}_X"], g[f"single_{mb}_masked"])
        np.testing.assert_array_equal(g[f"prepared_{mb}_posmask"], _mask(g, mb))
    X3, c3, kept = IO.build_multi_channel_union(mat, names, [_mask(g, mb) for mb in ("GpC", "CpG", "A")])
    np.testing.assert_array_equal(X3, g["multi_X"])
    np.testing.assert_array_equal(c3, g["multi_coords"])
    assert [("GpC", "CpG", "A")[i] for i in kept] == [str(x) for x in g["multi_used"]]
    np.testing.assert_array_equal(IO.mask_uncovered(X3, c3, names, cov), g["multi_masked"])


def test_reduction_oracle_matches_reference_fixtures():
    g = load_golden("reduce_layers")
    layer = g["layer"]
    mod, valid = IO.layer_fraction_counts(layer)
    assert mod == float(g["modified"]) and valid == int(g["valid"])
    count, size = IO.run_histograms(layer)
    assert sorted(count) == g["count_keys"].tolist() and [count[k] for k in sorted(count)] == g["count_vals"].tolist()
    assert sorted(size) == g["size_keys"].tolist() and [size[k] for k in sorted(size)] == g["size_vals"].tolist()
    for w in (1, 4, 9):
        np.testing.assert_allclose(IO.column_mean_metric(layer, w), g[f"metric_w{w}"], rtol=0, atol=1e-15)


def _ragged(n, flat):
    out, pos = [], 0
    for k in n.tolist():
        out.append(flat[pos:pos + k].tolist())
        pos += k
    return out


def _records_of(layer):
    rec = []
    for i, row in enumerate(layer):
        b = np.concatenate(([False], np.nan_to_num(row, nan=0.0) > 0, [False]))
        starts = np.flatnonzero(~b[:-1] & b[1:])
        ends = np.flatnonzero(b[:-1] & ~b[1:])
        rec.extend((i, int(s), int(e), 0) for s, e in zip(starts, ends))
    return np.asarray(rec, dtype=np.int32).reshape(-1, 4)


def test_interval_consumer_matches_reference_extract_intervals():
    """intervals_by_read on (read, start, end, class) records == extract_intervals_from_row on the dense rows
    (analysis/compute/hmm_features.py:15-79), whole rows and window-clipped with the overlap rule."""
    from smftools_b200.reductions import intervals_by_read

    g = load_golden("reduce_layers")
    layer, coords = g["layer"], g["coords"]
    rec = _records_of(layer)
    sizes, dists = intervals_by_read(rec, coords, layer.shape[0])
    assert sizes == _ragged(g["iv_sizes_n"], g["iv_sizes"])
    got_d, want_d = dists, _ragged(g["iv_dists_n"], g["iv_dists"])
    assert [len(x) for x in got_d] == [len(x) for x in want_d]
    np.testing.assert_allclose(np.concatenate([np.asarray(x, float) for x in got_d]),
                               np.concatenate([np.asarray(x, float) for x in want_d]), rtol=0, atol=0)
    lo, hi = g["window"].tolist()
    sizes, dists = intervals_by_read(rec, coords, layer.shape[0], window=(lo, hi), max_mask_overlap_frac=0.5)
    assert sizes == _ragged(g["ivc_sizes_n"], g["ivc_sizes"])
    want_d = _ragged(g["ivc_dists_n"], g["ivc_dists"])
    assert [len(x) for x in dists] == [len(x) for x in want_d]
    np.testing.assert_allclose(np.concatenate([np.asarray(x, float) for x in dists] + [np.zeros(0)]),
                               np.concatenate([np.asarray(x, float) for x in want_d] + [np.zeros(0)]), rtol=0, atol=0)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_oracle_equals_live_reference_builders_on_fresh_input():
    import types

    import pandas as pd

    from oracle.make_golden_inputs import synthetic_adata
    from oracle.ref_loader import load_reference_cli

    ha, ph = load_reference_cli()
    cfg = types.SimpleNamespace(output_binary_layer_name="binarized", smf_modality="conversion")
    ad = synthetic_adata(seed=77, N=17, V=133)
    names = np.asarray(ad.var_names, dtype=str)
    for mb in METHBASES:
        pm = ha._resolve_pos_mask_for_methbase(ad, "refA", mb)
        col = IO.pos_mask_column(list(ad.var.columns), "refA", mb)
        np.testing.assert_array_equal(pm, ad.var[col].to_numpy(dtype=bool))
        Xr, cr = ha.build_single_channel(ad, "refA", mb, "conversion", cfg)
        Xo, co = IO.build_single_channel(np.asarray(ad.X), names, pm)
        np.testing.assert_array_equal(Xr, Xo)
        np.testing.assert_array_equal(cr, co)
        np.testing.assert_array_equal(ph._mask_uncovered_model_input(ad, Xr, cr),
                                      IO.mask_uncovered(Xo, co, names, ad.layers["covered_base_mask"]))
    with pytest.raises(ValueError):
        ha.build_single_channel(ad, "refA", "nonexistent", "conversion", cfg)
    with pytest.raises(ValueError):
        IO.build_multi_channel_union(np.asarray(ad.X), names, [ha._resolve_pos_mask_for_methbase(ad, "refA", "GpC"), None])
    row = np.array([np.nan, 1, 1, 0, 1, np.nan, 1, 1, 1, 0])
    np.testing.assert_array_equal(ph._feature_run_lengths(row), IO.feature_run_lengths(row))
