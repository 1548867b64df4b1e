"""GPU parity of the AnnData annotation path (feature calling, merge, size classes, span mask)
against the layer sets written by the unmodified reference (tests/golden/annotate_*.npz) and
against the oracle on larger random inputs."""
import numpy as np
import pandas as pd
import pytest
import torch

from conftest import golden_names, load_golden, params_from_golden
from helpers import model_from_params
from oracle import hmm_oracle as O
from oracle.ref_loader import DuckAnnData
from test_oracle_golden import annotate_groups

pytestmark = pytest.mark.gpu

ANNOT = golden_names("annotate")


def _feature_sets(g):
    return {grp: {"state": st, "features": {k: tuple(v) for k, v in feats.items()}}
            for grp, (st, feats) in annotate_groups(g).items()}


def _adata_from_golden(g):
    N = len(g["starts"])
    obs = pd.DataFrame({"reference_start": g["starts"], "reference_end": g["ends"]},
                       index=[f"r{i}" for i in range(N)])
    ad = DuckAnnData(N, g["full"], obs=obs)
    if "covered" in g:
        ad.layers["covered_base_mask"] = g["covered"]
    return ad


@pytest.mark.parametrize("name", ANNOT)
def test_annotate_adata_layers_match_reference(name):
    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p)
    ad = _adata_from_golden(g)
    m.annotate_adata(ad, prefix="GpC", X=g["X"].astype(np.float64), coords=g["coords"], var_mask=g["var_mask"],
                     decode=str(g["decode"]), feature_sets=_feature_sets(g), prob_threshold=0.5)
    appended = [str(a) for a in g["appended"]]
    assert list(ad.uns["hmm_appended_layers"]) == appended  # same layers, same creation order
    assert ad.uns["hmm_annotated"] is True
    for nm in appended:
        ours = np.asarray(ad.layers[nm])
        ref = g[f"layer__{nm}"]
        assert ours.dtype == ref.dtype, (nm, ours.dtype, ref.dtype)
        if "posterior" in nm:
            np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-5, equal_nan=True)
        else:
            np.testing.assert_array_equal(ours, ref, err_msg=nm)
    merged = m.merge_intervals_to_new_layer(ad, "GpC_all_footprint_features", distance_threshold=10)
    created = m.write_size_class_layers_from_binary(
        ad, merged, out_prefix="GpC", feature_ranges=_feature_sets(g)["footprint"]["features"], suffix="_merged")
    names = [str(a) for a in g["merge_names"]]
    assert [merged, merged + "_lengths"] + created == names
    for nm in names:
        ours = np.asarray(ad.layers[nm])
        ref = g[f"merge__{nm}"]
        assert ours.dtype == ref.dtype, nm
        np.testing.assert_array_equal(ours, ref, err_msg=nm)
    # a second call is a no-op unless force_redo (HMM.py:1219-1220)
    before = {k: np.array(v) for k, v in ad.layers.items()}
    m.annotate_adata(ad, prefix="GpC", X=g["X"].astype(np.float64), coords=g["coords"], var_mask=g["var_mask"],
                     decode=str(g["decode"]), feature_sets=_feature_sets(g))
    for k, v in before.items():
        np.testing.assert_array_equal(np.asarray(ad.layers[k]), v)


@pytest.mark.parametrize("name", ANNOT)
def test_annotate_in_small_read_chunks_gives_identical_layers(name):
    """Force the chunked device pipeline (a few reads per chunk) and the multi-chunk EM path."""
    g = load_golden(name)
    p = params_from_golden(g)
    m = model_from_params(p)
    m.device_chunk_bytes = 40_000  # ~3 reads per chunk
    ad = _adata_from_golden(g)
    m.annotate_adata(ad, prefix="GpC", X=g["X"].astype(np.float64), coords=g["coords"], var_mask=g["var_mask"],
                     decode=str(g["decode"]), feature_sets=_feature_sets(g), prob_threshold=0.5)
    for nm in [str(a) for a in g["appended"]]:
        if "posterior" in nm:
            np.testing.assert_allclose(ad.layers[nm], g[f"layer__{nm}"], rtol=0, atol=1e-5, equal_nan=True)
        else:
            np.testing.assert_array_equal(ad.layers[nm], g[f"layer__{nm}"], err_msg=nm)
    # fit with the observations split over several resident chunks == one chunk
    X = g["X"].astype(np.float64)
    m1 = model_from_params(p)
    m2 = model_from_params(p)
    m2.device_chunk_bytes = 3 * X.shape[1] * 8
    h1 = m1.fit(X, g["coords"], max_iter=3, tol=0.0)
    h2 = m2.fit(X, g["coords"], max_iter=3, tol=0.0)
    assert np.allclose(h1, h2, rtol=1e-12)
    assert np.abs(m1.emission.detach().numpy() - m2.emission.detach().numpy()).max() < 1e-12
    # feature sets absent -> only states / posterior, unmasked (HMM.py:1280-1283)
    ad2 = _adata_from_golden(g)
    m.annotate_adata(ad2, prefix="GpC", X=X, coords=g["coords"], var_mask=g["var_mask"], decode=str(g["decode"]),
                     feature_sets={})
    assert list(ad2.uns["hmm_appended_layers"]) == ["GpC_states", "GpC_posterior_modified"]
    assert ad2.layers["GpC_states"].dtype == np.int8 and not np.isnan(ad2.layers["GpC_posterior_modified"]).any()


@pytest.mark.parametrize("N,L,seed", [(257, 700, 1), (64, 3001, 2), (1, 40, 3), (300, 5, 4),
                                      # rows on 16-byte boundaries: vectorised interval kernel (1024-position super-blocks)
                                      (130, 2048, 5), (33, 1024, 6), (70, 4112, 7), (9, 16, 8), (40, 1040, 9)])
def test_feature_kernels_against_oracle(N, L, seed):
    from smftools_b200 import engine

    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")
    K = 3
    states = O.synth_reads(N, L, K, seed, nan_rate=0.0, truncate=False)  # reuse the chain as fake calls
    states = (rng.integers(0, K, size=(N, L)) if L < 20 else (states * 2).astype(np.int64) % K).astype(np.int8)
    coords = O.synth_coords(L, seed, span_factor=5) + 500
    full = np.arange(coords[0] - 7, coords[-1] + 9)
    V = len(full)
    ranges = [(6, 40), (40, 100), (100, 200), (200, np.inf)]
    left = np.searchsorted(full, coords, side="left").astype(np.int32)
    right = np.searchsorted(full, coords, side="right").astype(np.int32)
    for target in (0, 2):
        member = states == target
        all_l, all_len, cls_l, cls_len = O.call_features(member, coords, full, ranges)
        cls_d, len_d, iv = engine.call_features(
            states=torch.from_numpy(states).to(dev), post=None, target=target, threshold=0.0,
            coords=torch.from_numpy(coords).to(dev), grid_left=torch.from_numpy(left).to(dev),
            grid_right=torch.from_numpy(right).to(dev), n_grid=V, ranges=ranges, max_intervals=N * (L // 2 + 2))
        cls_h, len_h = cls_d.cpu().numpy(), len_d.cpu().numpy()
        np.testing.assert_array_equal((cls_h > 0).astype(np.uint8), all_l)
        np.testing.assert_array_equal(np.where(cls_h > 0, len_h, 0), all_len)
        for c in range(len(ranges)):
            np.testing.assert_array_equal((cls_h == c + 1).astype(np.uint8), cls_l[c])
            np.testing.assert_array_equal(np.where(cls_h == c + 1, len_h, 0), cls_len[c])
        # compacted intervals describe exactly the dense layers
        iv = iv.cpu().numpy()
        dense = np.zeros((N, V), dtype=np.uint8)
        for r, s, e, c in iv:
            assert (dense[r, s:e] == 0).all()
            dense[r, s:e] = c + 1
        np.testing.assert_array_equal(dense, cls_h)
        # interval-only mode returns the same records without writing dense layers
        _, _, iv2 = engine.call_features(
            states=torch.from_numpy(states).to(dev), post=None, target=target, threshold=0.0,
            coords=torch.from_numpy(coords).to(dev), grid_left=torch.from_numpy(left).to(dev),
            grid_right=torch.from_numpy(right).to(dev), n_grid=V, ranges=ranges, want_dense=False,
            max_intervals=N * (L // 2 + 2))
        key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]  # noqa: E731
        np.testing.assert_array_equal(key(iv2.cpu().numpy()), key(iv))
        # posterior-threshold membership gives the same result as the equivalent state test
        post = np.where(member, 0.75, 0.25).astype(np.float32)
        cls_p, len_p, _ = engine.call_features(
            states=None, post=torch.from_numpy(post).to(dev), target=target, threshold=0.5,
            coords=torch.from_numpy(coords).to(dev), grid_left=torch.from_numpy(left).to(dev),
            grid_right=torch.from_numpy(right).to(dev), n_grid=V, ranges=ranges)
        assert torch.equal(cls_p, cls_d) and torch.equal(len_p, len_d)
        _, _, iv3 = engine.call_features(
            states=None, post=torch.from_numpy(post).to(dev), target=target, threshold=0.5,
            coords=torch.from_numpy(coords).to(dev), grid_left=torch.from_numpy(left).to(dev),
            grid_right=torch.from_numpy(right).to(dev), n_grid=V, ranges=ranges, want_dense=False,
            max_intervals=N * (L // 2 + 2))
        np.testing.assert_array_equal(key(iv3.cpu().numpy()), key(iv))
        # merge + size classes
        for dt in (0, 10, 60):
            mg, ml = O.merge_runs(all_l, full, dt)
            cl, cll = O.size_classes_from_binary(mg, full, ranges)
            mg_d, ml_d, c_d, cl_d = engine.merge_runs(torch.from_numpy(all_l).to(dev), torch.from_numpy(full).to(dev),
                                                      True, dt, ranges)
            np.testing.assert_array_equal(mg_d.cpu().numpy(), mg)
            np.testing.assert_array_equal(ml_d.cpu().numpy(), ml)
            ch, clh = c_d.cpu().numpy(), cl_d.cpu().numpy()
            for c in range(len(ranges)):
                np.testing.assert_array_equal((ch == c + 1).astype(np.uint8), cl[c])
                np.testing.assert_array_equal(np.where(ch == c + 1, clh, 0), cll[c])
        # index-gap variant (non-integer var names)
        mg, ml = O.merge_runs(all_l, full, 3, coords_are_ints=False)
        mg_d, ml_d, _, _ = engine.merge_runs(torch.from_numpy(all_l).to(dev), None, False, 3)
        np.testing.assert_array_equal(mg_d.cpu().numpy(), mg)
        np.testing.assert_array_equal(ml_d.cpu().numpy(), ml)


def test_span_mask_kernel_against_oracle_and_reference_kats():
    from smftools_b200 import engine

    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    N, V = 333, 1201
    coords = np.sort(rng.choice(np.arange(5000), size=V, replace=False)).astype(np.int64)
    starts = rng.integers(0, 2500, size=N).astype(float) + rng.random(N)
    ends = starts + rng.integers(0, 2500, size=N)
    starts[3] = np.nan
    ends[7] = np.inf
    ref = O.span_valid(coords, starts, ends)
    got = engine.span_mask(N, V, dev, coords=torch.from_numpy(coords).to(dev), start=torch.from_numpy(starts).to(dev),
                           end=torch.from_numpy(ends).to(dev)).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, ref)
    cov = rng.random((N, V)) > 0.2
    got = engine.span_mask(N, V, dev, covered=torch.from_numpy(cov.view(np.uint8)).to(dev)).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, cov)
    # reference KATs (tests/unit/hmm/test_mask_read_span.py:10-58) through the public function
    import smftools_b200 as S

    ad = DuckAnnData(2, ["1", "2", "3", "4"], obs=pd.DataFrame({"reference_start": [2, 1], "reference_end": [3, 4]}))
    ad.layers["hmm_test"] = np.array([[1, 2, 3, 4], [5, 6, 7, 8]], dtype=int)
    S.mask_layers_outside_read_span(ad, ["hmm_test"])
    out = ad.layers["hmm_test"]
    assert out.dtype == np.float64
    assert np.isnan(out[0, 0]) and np.isnan(out[0, 3]) and out[0, 1] == 2 and out[0, 2] == 3
    assert np.all(~np.isnan(out[1]))
    ad = DuckAnnData(1, ["4", "3", "2", "1"], obs=pd.DataFrame({"reference_start": [2], "reference_end": [3]}),
                     var=pd.DataFrame({"Original_var_names": ["1", "2", "3", "4"]}, index=["4", "3", "2", "1"]))
    ad.layers["hmm_test"] = np.array([[10, 11, 12, 13]], dtype=int)
    S.mask_layers_outside_read_span(ad, ["hmm_test"])
    out = ad.layers["hmm_test"]
    assert np.isnan(out[0, 0]) and np.isnan(out[0, 3]) and out[0, 1] == 11 and out[0, 2] == 12
    ad = DuckAnnData(1, ["0", "1", "2", "3", "4"], obs=pd.DataFrame({"reference_start": [0], "reference_end": [4]}))
    ad.layers["covered_base_mask"] = np.array([[1, 1, 0, 1, 1]], dtype=np.int8)
    ad.layers["hmm_test"] = np.ones((1, 5), dtype=float)
    S.mask_layers_outside_read_span(ad, ["hmm_test"])
    np.testing.assert_array_equal(ad.layers["hmm_test"], [[1.0, 1.0, np.nan, 1.0, 1.0]])


def test_footprint_merge_known_answer():
    """tests/unit/test_hmm_partitioned_cli.py:217-248 through the public methods."""
    import smftools_b200 as S

    values = np.zeros((1, 14), dtype=np.uint8)
    values[0, :2] = 1
    values[0, 12:] = 1
    ad = DuckAnnData(1, [str(i) for i in range(14)],
                     obs=pd.DataFrame({"reference_start": [0], "reference_end": [13]}, index=["read1"]))
    ad.layers["C_all_footprint_features"] = values
    ad.uns["hmm_appended_layers"] = np.asarray(["C_all_footprint_features"])
    m = S.SingleBernoulliHMM()
    merged = m.merge_intervals_to_new_layer(ad, "C_all_footprint_features", distance_threshold=10)
    m.write_size_class_layers_from_binary(ad, merged, out_prefix="C", feature_ranges={"small_bound_stretch": [6, 30]},
                                          suffix="_merged")
    assert np.all(np.asarray(ad.layers["C_all_footprint_features_merged"]) == 1)
    assert np.all(np.asarray(ad.layers["C_all_footprint_features_merged_lengths"]) == 14)
    assert "C_small_bound_stretch_merged" in ad.layers
    assert "C_all_footprint_features_merged" in list(ad.uns["hmm_appended_layers"])
