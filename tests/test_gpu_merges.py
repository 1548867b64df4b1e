"""GPU parity of the fused merge chain: ``apply_merges`` / ``smf_hmm_merge_chain``
against the layers the unmodified reference leaves after merge -> size classes -> read-span mask
(tests/golden/annotate_*.npz, keys chain__*) and against the oracle on larger random layers."""
import types

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, params_from_golden
from helpers import model_from_params
from oracle import hmm_oracle as O
from test_gpu_annotate import _adata_from_golden, _feature_sets

pytestmark = pytest.mark.gpu

ANNOT = golden_names("annotate")
PAIRS = [("all_footprint_features", 10), ("all_accessible_features", 60)]


@pytest.mark.parametrize("name", ANNOT)
@pytest.mark.parametrize("cfg_kind", ["namespace", "dict"])
def test_apply_merges_matches_reference_chain(name, cfg_kind):
    import smftools_b200 as S

    g = load_golden(name)
    m = model_from_params(params_from_golden(g))
    ad = _adata_from_golden(g)
    fsets = S.normalize_hmm_feature_sets(_feature_sets(g))
    m.annotate_adata(ad, prefix="GpC", X=g["X"].astype(np.float64), coords=g["coords"], var_mask=g["var_mask"],
                     decode=str(g["decode"]), feature_sets=fsets, prob_threshold=0.5)
    # the goldens ran the un-masked merge of the footprint layer first (same layers, overwritten by the chain)
    merged = m.merge_intervals_to_new_layer(ad, "GpC_all_footprint_features", distance_threshold=10)
    m.write_size_class_layers_from_binary(ad, merged, out_prefix="GpC",
                                          feature_ranges=fsets["footprint"]["features"], suffix="_merged")
    cfg = {"hmm_merge_layer_features": PAIRS + [("all_missing_features", 5)], "hmm_merged_suffix": "_merged"}
    if cfg_kind == "namespace":
        cfg = types.SimpleNamespace(**cfg)
    S.apply_merges(ad, m, "GpC", fsets, cfg)
    names = [str(n) for n in g["chain_names"]]
    for nm in names:
        ours = np.asarray(ad.layers[nm])
        ref = g[f"chain__{nm}"]
        assert ours.dtype == ref.dtype == np.float64, nm
        np.testing.assert_array_equal(ours, ref, err_msg=nm)
    assert list(ad.uns["hmm_appended_layers"]) == [str(a) for a in g["chain_appended"]]


def _random_layer(rng, N, V, density):
    # runs with geometric lengths separated by geometric gaps, some rows empty / full
    lay = np.zeros((N, V), dtype=np.uint8)
    for i in range(N):
        if i % 17 == 0:
            continue
        if i % 19 == 0:
            lay[i] = 1
            continue
        v = int(rng.integers(0, 20))
        while v < V:
            ln = int(rng.geometric(0.08))
            lay[i, v:v + ln] = 1
            v += ln + int(rng.geometric(density))
    return lay


@pytest.mark.parametrize("mode", ["span", "covered", "index_coords", "no_classes"])
def test_merge_chain_packed_matches_oracle(mode):
    from smftools_b200 import engine

    rng = np.random.default_rng(7)
    N, V = 150, 2500
    dev = torch.device("cuda")
    lay = _random_layer(rng, N, V, 0.05)
    coords = np.cumsum(rng.integers(1, 6, size=V)).astype(np.int64) + 500
    ranges = [(6.0, 40.0), (40.0, 100.0), (100.0, 200.0), (200.0, float("inf"))]
    if mode == "no_classes":
        ranges = []
    are_ints = mode != "index_coords"
    starts = coords[0] + rng.integers(-5, 300, size=N).astype(float)
    ends = coords[-1] - rng.integers(-5, 300, size=N).astype(float)
    starts[3] = np.nan
    ends[5] = np.inf
    covered = (rng.random((N, V)) > 0.2) if mode == "covered" else None
    valid = O.span_valid(coords, starts, ends, covered=covered)
    ref = O.merge_chain(lay, coords if are_ints else np.arange(V), 12, ranges, valid, coords_are_ints=are_ints)
    code, rl = engine.merge_chain(
        torch.from_numpy(lay).to(dev), torch.from_numpy(coords).to(dev) if are_ints else None, are_ints, 12, ranges,
        covered=torch.from_numpy(covered.view(np.uint8)).to(dev) if covered is not None else None,
        span_coords=torch.from_numpy(coords).to(dev), start=torch.from_numpy(starts).to(dev),
        end=torch.from_numpy(ends).to(dev))
    code = code.cpu().numpy()
    rl = rl.cpu().numpy()
    drop = (code & 0x80) == 0
    np.testing.assert_array_equal(~drop, valid)

    def masked(a):
        f = a.astype(np.float64)
        f[drop] = np.nan
        return f

    np.testing.assert_array_equal(masked((code & 0x40) != 0), ref[0])
    np.testing.assert_array_equal(masked(rl), ref[1])
    for c in range(len(ranges)):
        hit = (code & 15) == c + 1
        np.testing.assert_array_equal(masked(hit), ref[2 + 2 * c])
        np.testing.assert_array_equal(masked(np.where(hit, rl, 0)), ref[3 + 2 * c])
    # cells of merged runs that match no class carry class 0; class bits never appear outside runs
    assert not np.any(((code & 15) != 0) & ((code & 0x40) == 0))


def test_merge_chain_argument_errors():
    from smftools_b200 import engine

    dev = torch.device("cuda")
    lay = torch.zeros((4, 10), dtype=torch.uint8, device=dev)
    coords = torch.arange(10, dtype=torch.int64, device=dev)
    with pytest.raises(ValueError):
        engine.merge_chain(lay, coords, True, 3, [])  # neither covered nor span inputs
    with pytest.raises(ValueError):
        engine.merge_chain(lay, coords, True, 3, [(0.0, 1.0)] * 16, covered=torch.ones_like(lay))
    code, rl = engine.merge_chain(lay, coords, True, 3, [], covered=torch.ones_like(lay))
    assert int(code.min()) == 0x80 and int(code.max()) == 0x80 and int(rl.abs().max()) == 0
