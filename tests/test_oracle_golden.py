"""Pin oracle/hmm_oracle.py against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, params_from_golden
from oracle import hmm_oracle as O

CORE = [n for n in golden_names() if not n.startswith(("annotate", "inputs_", "reduce_"))]
ANNOT = golden_names("annotate")


def _coords(g):
    return g["coords"] if "trans_by_bin" in "".join(g.keys()) else None


@pytest.mark.parametrize("name", CORE)
def test_posterior_and_loglik(name):
    g = load_golden(name)
    p = params_from_golden(g)
    gamma, ll = O.posterior(g["X"], p, g["coords"])
    assert np.abs(gamma - g["gamma"]).max() < 1e-11
    assert np.allclose(ll, g["loglik"], rtol=1e-12, atol=1e-10)
    assert np.array_equal(gamma.argmax(axis=2).astype(np.int8), g["states_marginal"])


@pytest.mark.parametrize("name", CORE)
def test_viterbi_bit_exact(name):
    g = load_golden(name)
    p = params_from_golden(g)
    st, margin = O.viterbi(g["X"], p, g["coords"], return_margins=True)
    same = (st == g["states_viterbi"]).all(axis=1)
    # the oracle may only differ from the reference on reads that contain a near-tie
    assert np.all(same | (margin < 1e-9)), name


@pytest.mark.parametrize("name", CORE)
def test_fit_history_and_parameters(name):
    g = load_golden(name)
    p = params_from_golden(g)
    iters = len(g["fit_hist"])
    q, hist = O.fit(g["X"], p, g["coords"], max_iter=iters, tol=0.0)
    assert np.allclose(hist, g["fit_hist"], rtol=1e-11, atol=1e-9)
    for key in ("start", "trans", "emission"):
        assert np.abs(getattr(q, key).reshape(-1) - g[f"p1_{key}"].reshape(-1)).max() < 1e-10, key
    if "p1_trans_by_bin" in g:
        assert np.abs(q.trans_by_bin - g["p1_trans_by_bin"]).max() < 1e-10
    # emission-only adaptation with frozen start/transitions (still re-normalised every iteration)
    r, hist2 = O.fit(g["X"], q, g["coords"], max_iter=len(g["adapt_hist"]), tol=0.0,
                     update_start=False, update_trans=False)
    assert np.allclose(hist2, g["adapt_hist"], rtol=1e-11, atol=1e-9)
    for key in ("start", "trans", "emission"):
        assert np.abs(getattr(r, key).reshape(-1) - g[f"p2_{key}"].reshape(-1)).max() < 1e-10, key


def _target_index(p, state):
    if isinstance(state, (int, np.integer)):
        return max(0, min(int(state), p.K - 1))
    score = p.emission if p.emission.ndim == 1 else p.emission.mean(axis=1)
    return int(np.argmin(score)) if str(state).lower().startswith("non") else int(np.argmax(score))


def annotate_groups(g):
    feats_fp = {"small_bound_stretch": (6, 40), "medium_bound_stretch": (40, 100),
                "putative_nucleosome": (100, 200), "large_bound_stretch": (200, np.inf)}
    feats_ac = {"small_accessible_patch": (3, 20), "mid_accessible_patch": (20, 40),
                "large_accessible_patch": (40, 110), "nucleosome_depleted_region": (110, np.inf)}
    if bool(g["state_ints"]):
        return {"footprint": (1, feats_fp), "tf": (2, {"tf_small": (1, 30), "tf_big": (30, np.inf)}),
                "accessible": (0, feats_ac)}
    return {"footprint": ("Non-Modified", feats_fp), "accessible": ("Modified", feats_ac)}


@pytest.mark.parametrize("name", ANNOT)
def test_annotate_layers(name):
    g = load_golden(name)
    p = params_from_golden(g)
    decode = str(g["decode"])
    st, gamma = O.decode(g["X"], p, g["coords"], decode=decode)
    full = g["full"]
    valid = O.span_valid(full, g["starts"], g["ends"], g.get("covered"))
    N, V = len(g["starts"]), len(full)
    cols = np.nonzero(g["var_mask"])[0]

    def masked(arr):
        out = np.asarray(arr).astype(np.float64) if not np.issubdtype(np.asarray(arr).dtype, np.floating) else np.array(arr)
        out[~valid] = np.nan
        return out

    states_layer = np.full((N, V), -1, dtype=np.int8)
    states_layer[:, cols] = st
    np.testing.assert_array_equal(masked(states_layer), g["layer__GpC_states"])
    t_mod = _target_index(p, "Modified")
    post = np.zeros((N, V), dtype=np.float32)
    post[:, cols] = gamma[:, :, t_mod].astype(np.float32)
    np.testing.assert_allclose(masked(post), g["layer__GpC_posterior_modified"], rtol=0, atol=1e-7)
    assert g["layer__GpC_posterior_modified"].dtype == np.float32

    for group, (state, feats) in annotate_groups(g).items():
        t = _target_index(p, state)
        member = (st == t) if decode == "viterbi" else (gamma[:, :, t] >= 0.5)
        all_l, all_len, cls_l, cls_len = O.call_features(member, g["coords"], full, list(feats.values()))
        np.testing.assert_array_equal(masked(all_l), g[f"layer__GpC_all_{group}_features"])
        np.testing.assert_array_equal(masked(all_len), g[f"layer__GpC_all_{group}_features_lengths"])
        for ci, feat in enumerate(feats):
            np.testing.assert_array_equal(masked(cls_l[ci]), g[f"layer__GpC_{feat}"])
            np.testing.assert_array_equal(masked(cls_len[ci]), g[f"layer__GpC_{feat}_lengths"])

    # merge + size classes on the (masked, float) footprint layer
    base = g["layer__GpC_all_footprint_features"]
    merged, mlen = O.merge_runs(base, full, 10)
    np.testing.assert_array_equal(merged, g["merge__GpC_all_footprint_features_merged"])
    np.testing.assert_array_equal(mlen, g["merge__GpC_all_footprint_features_merged_lengths"])
    feats = annotate_groups(g)["footprint"][1]
    cl, cll = O.size_classes_from_binary(merged, full, list(feats.values()))
    for ci, feat in enumerate(feats):
        np.testing.assert_array_equal(cl[ci], g[f"merge__GpC_{feat}_merged"])
        np.testing.assert_array_equal(cll[ci], g[f"merge__GpC_{feat}_merged_lengths"])


@pytest.mark.parametrize("name", golden_names("annotate"))
def test_merge_chain_matches_reference_apply_merges(name):
    """oracle merge_chain == merge -> size classes -> read-span mask as the reference runs them
    (tools/partitioned_hmm.py:78-108) for both default merge pairs."""
    g = load_golden(name)
    full = g["full"]
    valid = O.span_valid(full, g["starts"], g["ends"], covered=g["covered"] if "covered" in g else None)
    groups = annotate_groups(g)
    names = [str(n) for n in g["chain_names"]]
    got = {}
    for core, dist in (("all_footprint_features", 10), ("all_accessible_features", 60)):
        base = g[f"layer__GpC_{core}"]
        grp = core[len("all_"):-len("_features")]
        feats = groups[grp][1]
        arrs = O.merge_chain(base, full, dist, list(feats.values()), valid)
        order = [f"GpC_{core}_merged", f"GpC_{core}_merged_lengths"]
        for feat in feats:
            order += [f"GpC_{feat}_merged", f"GpC_{feat}_merged_lengths"]
        got.update(dict(zip(order, arrs)))
    assert list(got) == names
    for nm in names:
        ref = g[f"chain__{nm}"]
        assert ref.dtype == np.float64
        np.testing.assert_array_equal(got[nm], ref, err_msg=nm)


def test_known_answer_merge_from_reference_test():
    # tests/unit/test_hmm_partitioned_cli.py:217-248: two runs 10 bp apart on a 14-wide row merge
    row = np.zeros((1, 14), dtype=np.uint8)
    row[0, :2] = 1
    row[0, 12:] = 1
    merged, lens = O.merge_runs(row, np.arange(14), 10)
    assert merged.tolist() == [[1] * 14]
    assert lens.tolist() == [[14] * 14]


def test_known_answer_read_span_from_reference_test():
    # tests/unit/hmm/test_mask_read_span.py:10-26 (var_names 1..4, spans [2,3] and [1,4])
    valid = O.span_valid(np.array([1, 2, 3, 4]), [2, 1], [3, 4])
    assert valid.tolist() == [[False, True, True, False], [True, True, True, True]]
    # :29-45 (coordinates come from Original_var_names, order 1..4, span [2,3])
    assert O.span_valid(np.array([1, 2, 3, 4]), [2], [3]).tolist() == [[False, True, True, False]]
    # :48-58 covered_base_mask takes precedence over the span
    cov = np.array([[1, 1, 0, 1, 1]], dtype=np.int8)
    assert O.span_valid(np.arange(5), [0], [4], covered=cov).tolist() == [[True, True, False, True, True]]
