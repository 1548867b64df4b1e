"""CPU restatement of the reference's model-input builders and downstream layer reductions.

TEST INFRASTRUCTURE -- never imported by the product package.  numpy only; every function cites the
reference lines it follows.  Pinned against the unmodified reference functions by
``oracle/make_golden_inputs.py`` (fixtures ``tests/golden/inputs_*.npz``, ``reduce_*.npz``) and, when the
reference tree is present, live in ``tests/test_inputs_reductions_cpu.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# N2: input builders
# --------------------------------------------------------------------------------------
def pos_mask_column(var_columns: Sequence[str], ref: str, methbase: str) -> Optional[str]:
    """Which ``adata.var`` column holds the site mask of a methylation channel
    (cli/hmm_adata.py:204-248 ``_resolve_pos_mask_for_methbase``); None when absent."""
    key = str(methbase).strip().lower()
    cols = set(var_columns)
    if key in ("a",):
        cands = [f"{ref}_A_site"]
    elif key in ("c", "any_c", "anyc", "any-c"):
        cands = [f"{ref}_any_C_site", f"{ref}_C_site"]
    elif key in ("gpc", "gpc_site", "gpc-site"):
        cands = [f"{ref}_GpC_site"]
    elif key in ("cpg", "cpg_site", "cpg-site"):
        cands = [f"{ref}_CpG_site"]
    else:
        cands = [f"{ref}_{methbase}_site"]
    for c in cands:
        if c in cols:
            return c
    return None


def build_single_channel(matrix: np.ndarray, var_names: Sequence, pos_mask: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """cli/hmm_adata.py:251-278: X = dense float of the masked columns, coords = their integer names."""
    pm = np.asarray(pos_mask, dtype=bool)
    if int(pm.sum()) == 0:
        raise ValueError("No columns for methbase")
    X = np.asarray(matrix)[:, pm].astype(float)
    coords = np.asarray([int(v) for v in np.asarray(var_names)[pm]], dtype=int)
    return X, coords


def build_multi_channel_union(matrix: np.ndarray, var_names: Sequence, pos_masks: Sequence[np.ndarray]):
    """cli/hmm_adata.py:281-319: (N, L_union, C) on the union coordinate grid, NaN where a channel has no
    site; channels with an empty mask are dropped, fewer than two left is an error.  Returns
    (X3, coords, kept channel indices)."""
    per = []
    names = np.asarray(var_names)
    for ci, pm in enumerate(pos_masks):
        if pm is None:
            continue
        pm = np.asarray(pm, dtype=bool)
        if int(pm.sum()) == 0:
            continue
        per.append((ci, np.asarray(matrix)[:, pm].astype(float), np.asarray([int(v) for v in names[pm]], dtype=int)))
    if len(per) < 2:
        raise ValueError("Need >=2 methbases with columns for union multi-channel")
    coords = np.unique(np.concatenate([c for _, _, c in per])).astype(int)
    index = {int(v): i for i, v in enumerate(coords.tolist())}
    N = per[0][1].shape[0]
    X3 = np.full((N, coords.shape[0], len(per)), np.nan, dtype=float)
    for k, (_, Xc, cc) in enumerate(per):
        X3[:, np.array([index[int(v)] for v in cc.tolist()], dtype=int), k] = Xc
    return X3, coords, [ci for ci, _, _ in per]


def mask_uncovered(values: np.ndarray, coordinates: np.ndarray, var_names: Sequence, covered: Optional[np.ndarray]):
    """tools/partitioned_hmm.py:170-188: NaN where ``covered_base_mask`` is False at the coordinate's column."""
    if covered is None:
        return values
    positions = np.asarray([int(v) for v in var_names], dtype=np.int64)
    col_of = {int(p): i for i, p in enumerate(positions)}
    try:
        columns = np.asarray([col_of[int(p)] for p in coordinates])
    except KeyError as exc:
        raise ValueError(f"HMM coordinate {exc.args[0]} is absent from materialized coverage") from exc
    cov = np.asarray(covered, dtype=bool)[:, columns]
    out = np.asarray(values, dtype=float).copy()
    if out.ndim == 2:
        out[~cov] = np.nan
    else:
        out[~cov, :] = np.nan
    return out


def code_u8(X: np.ndarray) -> np.ndarray:
    """The kernels' compact code of a binary float matrix: 0, 1, 255 = missing (NaN)."""
    X = np.asarray(X, dtype=float)
    out = np.full(X.shape, 255, dtype=np.uint8)
    out[X == 0.0] = 0
    out[X == 1.0] = 1
    return out


# --------------------------------------------------------------------------------------
# N3: layer reductions
# --------------------------------------------------------------------------------------
def layer_fraction_counts(values: np.ndarray) -> Tuple[float, int]:
    """tools/partitioned_hmm.py:519-529: (#cells > 0 ignoring NaN, #finite cells) of one layer."""
    v = np.asarray(values, dtype=float)
    return float(np.nansum(v > 0)), int(np.isfinite(v).sum())


def feature_run_lengths(row: np.ndarray) -> np.ndarray:
    """tools/partitioned_hmm.py:572-585: column-count lengths of the positive runs of one row (NaN = 0)."""
    valid = np.nan_to_num(np.asarray(row, dtype=float), nan=0.0) > 0.5
    idx = np.flatnonzero(valid)
    if idx.size == 0:
        return np.array([], dtype=int)
    breaks = np.where(np.diff(idx) > 1)[0]
    starts = np.r_[idx[0], idx[breaks + 1]]
    ends = np.r_[idx[breaks] + 1, idx[-1] + 1]
    return ends - starts


def run_histograms(layer: np.ndarray) -> Tuple[Dict[int, int], Dict[int, int]]:
    """tools/partitioned_hmm.py:618-630: {runs per read: reads}, {run size: runs} over the rows of a layer."""
    count: Dict[int, int] = {}
    size: Dict[int, int] = {}
    for row in np.asarray(layer, dtype=float):
        sizes = feature_run_lengths(row)
        count[int(sizes.size)] = count.get(int(sizes.size), 0) + 1
        for s in sizes:
            size[int(s)] = size.get(int(s), 0) + 1
    return count, size


def column_mean_metric(matrix: np.ndarray, rolling_window: int = 1) -> np.ndarray:
    """hmm/call_hmm_peaks.py:151-160: nan-mean over reads per column (all-NaN columns -> 0), then the
    'same'-mode rolling mean."""
    with np.errstate(invalid="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            means = np.nanmean(np.asarray(matrix, dtype=float), axis=0)
    means = np.nan_to_num(means, nan=0.0)
    w = max(1, int(rolling_window))
    if w > 1:
        return np.convolve(means, np.ones(w, dtype=float) / float(w), mode="same")
    return means


# --------------------------------------------------------------------------------------
# N4: intervals of one row
# --------------------------------------------------------------------------------------
def extract_intervals_from_row(row, coords, full_row=None, full_coords=None, max_mask_overlap_frac=0.5):
    """analysis/compute/hmm_features.py:15-79: (sizes in bp, centre-to-centre distances of neighbours)."""
    binary = np.asarray(np.asarray(row, dtype=float) > 0, dtype=bool)
    if binary.size == 0 or not binary.any():
        return [], []
    pad = np.concatenate(([False], binary, [False]))
    starts = np.flatnonzero(~pad[:-1] & pad[1:])
    ends = np.flatnonzero(pad[:-1] & ~pad[1:]) - 1
    sizes: List[int] = []
    centers: List[float] = []
    for s, e in zip(starts, ends):
        sb, eb = float(coords[s]), float(coords[e])
        if full_row is not None and full_coords is not None:
            sm = np.flatnonzero(np.asarray(full_coords) == sb)
            em = np.flatnonzero(np.asarray(full_coords) == eb)
            if sm.size and em.size:
                fb = np.asarray(np.asarray(full_row, dtype=float) > 0, dtype=bool)
                fs, fe = int(sm[0]), int(em[-1])
                while fs > 0 and fb[fs - 1]:
                    fs -= 1
                while fe < len(fb) - 1 and fb[fe + 1]:
                    fe += 1
                clipped = eb - sb + 1.0
                full_len = float(full_coords[fe] - full_coords[fs] + 1.0)
                outside = 1.0 - (clipped / full_len) if full_len > 0 else 0.0
                if outside > max_mask_overlap_frac:
                    continue
        sizes.append(int(round(eb - sb + 1.0)))
        centers.append((sb + eb) / 2.0)
    dists = np.diff(np.asarray(centers, dtype=float)).tolist() if len(centers) > 1 else []
    return sizes, dists
