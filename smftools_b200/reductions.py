"""Downstream reductions and interval consumers fed from the packed device results.

The reference's consumers re-read the dense float64 layers on the host:

* ``_plot_feature_fractions`` (tools/partitioned_hmm.py:509-531): per layer ``nansum(values > 0)`` and
  ``isfinite(values).sum()``;
* ``_plot_feature_count_size_histograms`` (:572-630): runs per read and run sizes of every aggregate layer;
* ``call_hmm_peaks`` (hmm/call_hmm_peaks.py:151-160): per-column nan-mean over reads + rolling mean;
* ``extract_intervals_from_row`` (analysis/compute/hmm_features.py:15-79): interval sizes and neighbour
  centre distances of one binary row, optionally clipped to an analysis window.

Here the first three come from ``class_out`` / merge-chain code bytes + read-span validity on the device
(``smf_hmm_layer_colsums`` / ``smf_hmm_column_metric`` / ``smf_hmm_run_histograms``), and the last from the
compacted ``(read, grid_start, grid_end, class)`` interval records of ``smf_hmm_call_features`` -- 100-1000x
smaller than a dense layer -- instead of dense rows.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine


def _check_packed(cls: torch.Tensor, valid: Optional[torch.Tensor]):
    engine._require_cuda(cls, "class/code")
    if cls.dtype != torch.uint8 or cls.dim() != 2 or not cls.is_contiguous():
        raise ValueError("class/code must be a contiguous uint8 [N, V] tensor")
    if valid is not None and (valid.dtype != torch.uint8 or tuple(valid.shape) != tuple(cls.shape) or not valid.is_contiguous()):
        raise ValueError("valid must be a contiguous uint8 [N, V] tensor")
    return int(cls.shape[0]), int(cls.shape[1])


def layer_column_sums(cls: torch.Tensor, *, n_classes: int, packed: bool = False, valid: Optional[torch.Tensor] = None):
    """``(colsum int64 [1 + n_classes, V], colvalid int64 [V])`` on the device: row 0 = the aggregate
    (``all_*_features`` / ``*_merged``) layer, row ``c + 1`` = size class ``c``."""
    N, V = _check_packed(cls, valid)
    dev = cls.device
    colsum = torch.empty((n_classes + 1, V), dtype=torch.int64, device=dev)
    colvalid = torch.empty((V,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().smf_hmm_layer_colsums(engine._ptr(cls), 1 if packed else 0, engine._ptr(valid), N, V,
                                               int(n_classes), engine._ptr(colsum), engine._ptr(colvalid),
                                               engine._stream_ptr(dev))
    _lib.check(rc, "smf_hmm_layer_colsums")
    return colsum, colvalid


def feature_fraction_counts(cls: torch.Tensor, layer_names: Sequence[str], *, packed: bool = False,
                            valid: Optional[torch.Tensor] = None) -> Dict[str, Dict[str, float]]:
    """``{layer: {"modified", "valid", "fraction"}}`` as ``_plot_feature_fractions`` computes them
    (tools/partitioned_hmm.py:519-541); ``layer_names[0]`` is the aggregate layer, then one per size class."""
    colsum, colvalid = layer_column_sums(cls, n_classes=len(layer_names) - 1, packed=packed, valid=valid)
    tot = colsum.sum(dim=1).cpu().numpy()
    nvalid = int(colvalid.sum().item())
    out = {}
    for i, nm in enumerate(layer_names):
        out[str(nm)] = {"modified": float(tot[i]), "valid": nvalid,
                        "fraction": (float(tot[i]) / nvalid) if nvalid else float("nan")}
    return out


def column_mean_metric(colsum: torch.Tensor, colvalid: torch.Tensor, rolling_window: int = 1):
    """``(means, peak_metric)`` float64 [V] of ``call_hmm_peaks`` (hmm/call_hmm_peaks.py:151-158) from one row
    of ``layer_column_sums``."""
    if colsum.dtype != torch.int64 or colvalid.dtype != torch.int64 or colsum.shape != colvalid.shape or colsum.dim() != 1:
        raise ValueError("colsum / colvalid must be int64 [V]")
    V = int(colsum.shape[0])
    dev = colsum.device
    colsum = colsum.contiguous()
    means = torch.empty((V,), dtype=torch.float64, device=dev)
    w = max(1, int(rolling_window))
    metric = torch.empty((max(V, w) if w > 1 else V,), dtype=torch.float64, device=dev)  # np.convolve "same" length
    with torch.cuda.device(dev):
        rc = _lib.load().smf_hmm_column_metric(engine._ptr(colsum), engine._ptr(colvalid), V, max(1, int(rolling_window)),
                                               engine._ptr(means), engine._ptr(metric), engine._stream_ptr(dev))
    _lib.check(rc, "smf_hmm_column_metric")
    return means, metric


def run_histograms(cls: torch.Tensor, *, select_class: int = -1, packed: bool = False,
                   valid: Optional[torch.Tensor] = None) -> Tuple[Dict[int, int], Dict[int, int]]:
    """``({runs per read: reads}, {run size in grid columns: runs})`` of one binary layer, the dictionaries
    ``_plot_feature_count_size_histograms`` accumulates (tools/partitioned_hmm.py:618-630)."""
    N, V = _check_packed(cls, valid)
    dev = cls.device
    count_hist = torch.empty((V // 2 + 2,), dtype=torch.int64, device=dev)
    size_hist = torch.empty((V + 1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().smf_hmm_run_histograms(engine._ptr(cls), 1 if packed else 0, int(select_class), engine._ptr(valid),
                                                N, V, engine._ptr(count_hist), engine._ptr(size_hist),
                                                engine._stream_ptr(dev))
    _lib.check(rc, "smf_hmm_run_histograms")
    ch = count_hist.cpu().numpy()
    sh = size_hist.cpu().numpy()
    return ({int(k): int(ch[k]) for k in np.flatnonzero(ch)}, {int(s): int(sh[s]) for s in np.flatnonzero(sh)})


# --------------------------------------------------------------------------------------
# N4: interval records instead of dense rows
# --------------------------------------------------------------------------------------
def sort_interval_records(records: torch.Tensor) -> np.ndarray:
    """Interval records ``(read, grid_start, grid_end, class)`` ordered by read, then start (the kernel
    appends them in completion order); returned as a host int32 [n, 4] array."""
    if records.numel() == 0:
        return np.zeros((0, 4), dtype=np.int32)
    key = records[:, 0].to(torch.int64) * (1 << 32) + records[:, 1].to(torch.int64)
    order = torch.argsort(key)
    return records.index_select(0, order).cpu().numpy()


def clip_records_to_span(records: np.ndarray, first_col: np.ndarray, last_col: np.ndarray) -> np.ndarray:
    """Clip interval records to each read's valid column range ``[first_col[n], last_col[n]]`` (inclusive) --
    the effect the read-span NaN mask has on the dense rows (HMM.py:148-231) when the span is one contiguous
    range; records left empty are dropped."""
    rec = np.array(records, dtype=np.int64, copy=True)
    if rec.shape[0] == 0:
        return rec.astype(np.int32)
    lo = np.asarray(first_col, dtype=np.int64)[rec[:, 0]]
    hi = np.asarray(last_col, dtype=np.int64)[rec[:, 0]] + 1
    rec[:, 1] = np.maximum(rec[:, 1], lo)
    rec[:, 2] = np.minimum(rec[:, 2], hi)
    return rec[rec[:, 2] > rec[:, 1]].astype(np.int32)


def intervals_by_read(records: np.ndarray, coords: np.ndarray, n_reads: int, *, window: Optional[Tuple[int, int]] = None,
                      max_mask_overlap_frac: float = 0.5) -> Tuple[List[List[int]], List[List[float]]]:
    """Per read ``(sizes_bp, neighbour centre distances)`` -- what ``extract_intervals_from_row`` returns for
    the read's binary row (analysis/compute/hmm_features.py:15-79) -- from sorted interval records.

    ``coords`` are the (bp) coordinates of the grid columns.  With ``window = (lo, hi)`` (grid columns) the
    intervals are clipped to the window and those lying more than ``max_mask_overlap_frac`` outside it are
    dropped, exactly as the reference does with ``full_row`` / ``full_coords``."""
    coords = np.asarray(coords, dtype=float)
    sizes: List[List[int]] = [[] for _ in range(n_reads)]
    dists: List[List[float]] = [[] for _ in range(n_reads)]
    rec = np.asarray(records)
    if rec.shape[0] == 0:
        return sizes, dists
    r, s, e = rec[:, 0].astype(np.int64), rec[:, 1].astype(np.int64), rec[:, 2].astype(np.int64) - 1  # inclusive end
    fs, fe = s.copy(), e.copy()
    keep = np.ones(r.shape[0], dtype=bool)
    if window is not None:
        lo, hi = int(window[0]), int(window[1])
        keep = (e >= lo) & (s < hi)
        s = np.maximum(s, lo)
        e = np.minimum(e, hi - 1)
    sb, eb = coords[np.clip(s, 0, len(coords) - 1)], coords[np.clip(e, 0, len(coords) - 1)]
    if window is not None:
        clipped = eb - sb + 1.0
        full_len = coords[fe] - coords[fs] + 1.0
        with np.errstate(divide="ignore", invalid="ignore"):
            outside = np.where(full_len > 0, 1.0 - clipped / full_len, 0.0)
        keep &= ~(outside > max_mask_overlap_frac)
    size_bp = np.rint(eb - sb + 1.0).astype(np.int64)
    centers = (sb + eb) / 2.0
    idx = np.flatnonzero(keep)
    rr = r[idx]
    bounds = np.flatnonzero(np.diff(rr)) + 1
    for block in np.split(idx, bounds):
        if block.size == 0:
            continue
        read = int(r[block[0]])
        sizes[read] = size_bp[block].tolist()
        dists[read] = np.diff(centers[block]).tolist() if block.size > 1 else []
    return sizes, dists
