"""Model-input builders on the device: the step immediately before the hot path.

Host mirror of the reference's ``build_single_channel`` / ``build_multi_channel_union``
(cli/hmm_adata.py:251-319), ``_mask_uncovered_model_input`` and ``_prepare_model_input``
(tools/partitioned_hmm.py:140-188): same names, arguments, return order and exceptions.  What differs is
where the matrix is formed and what it is made of: the reference gathers the site columns on the host and
hands a float64 matrix (8 bytes per position) to ``fit`` / ``annotate_adata``; here the per-position matrix is
streamed to the device once, ``smf_hmm_build_obs`` gathers the site columns, applies the coverage mask and
codes binary calls as uint8 (0 / 1 / 255 = missing), and the result STAYS on the device as a ``CodedU8``
(or a float32 tensor when the calls are probabilistic) that ``fit`` / ``decode`` / ``annotate_adata`` take
directly -- no float64 host matrix, no second host-to-device copy.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine
from .hmm import COVERED_BASE_MASK, CodedU8, _default_cuda_device, _safe_int_coords, _to_dense_np


def resolve_pos_mask_for_methbase(subset, ref: str, methbase: str) -> Optional[np.ndarray]:
    """Boolean site mask of a methylation channel from ``subset.var`` (cli/hmm_adata.py:204-248)."""
    key = str(methbase).strip().lower()

    def _mask(column: str) -> np.ndarray:
        return subset.var[column].astype("boolean").fillna(False).to_numpy(dtype=bool)

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
    for col in cands:
        if col in subset.var:
            return _mask(col)
    return None


def _source_matrix(subset, smf_modality: Optional[str], cfg) -> np.ndarray:
    """The (n_obs, n_vars) matrix the reference trains on (cli/hmm_adata.py:173-201 without the column
    slice): ``cfg.output_binary_layer_name`` for the direct modality, ``.X`` otherwise."""
    if smf_modality == "direct":
        hmm_layer = getattr(cfg, "output_binary_layer_name", None)
        if hmm_layer is None or hmm_layer not in subset.layers:
            raise KeyError(f"Missing HMM training layer '{hmm_layer}' in subset.")
        mat = subset.layers[hmm_layer]
    else:
        mat = subset.X
    mat = _to_dense_np(mat)
    if mat.ndim != 2:
        raise ValueError(f"Expected 2D training matrix; got {mat.shape}")
    return mat


def _device_layer_dtype(mat: np.ndarray) -> np.ndarray:
    if mat.dtype in (np.float32, np.float64):
        return mat
    return mat.astype(np.float32)  # bool / small-integer calls: 0 / 1 are exact


def _build_on_device(mat: np.ndarray, cols: np.ndarray, n_channels: int, device: torch.device,
                     covered: Optional[np.ndarray], covered_cols: Optional[np.ndarray], chunk_bytes: int = 256 << 20):
    """Stream the rows of ``mat`` through the device and gather / mask / code them; returns a ``CodedU8``
    device tensor, or a float32 (float64 when float32 cannot carry the values) device tensor."""
    from .stream import ChunkUploader

    mat = _device_layer_dtype(np.asarray(mat))
    N, V = int(mat.shape[0]), int(mat.shape[1])
    C = int(n_channels)
    L = int(cols.shape[0])
    shape = (N, L) if C == 1 else (N, L, C)
    cols_d = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32).reshape(L, C)).to(device)
    ccols_d = None
    if covered is not None:
        covered = np.ascontiguousarray(covered).view(np.uint8)
        ccols_d = torch.from_numpy(np.ascontiguousarray(covered_cols, dtype=np.int32)).to(device)
    rows = max(1, min(N, int(chunk_bytes // max(V * mat.dtype.itemsize, 1)))) if N else 1
    for out_u8 in (True, False):
        out = torch.empty(shape, dtype=torch.uint8 if out_u8 else torch.float32, device=device)
        flag = torch.zeros((1,), dtype=torch.int32, device=device)
        if N and L:
            cur = torch.cuda.current_stream(device)
            up = ChunkUploader(mat, rows, device)
            cov_up = ChunkUploader(covered, rows, device) if covered is not None else None
            entry = torch.cuda.Event()
            entry.record(cur)
            up.wait_entry(entry)
            if cov_up is not None:
                cov_up.wait_entry(entry)
            ev = [None, None]
            try:
                for i, lo in enumerate(range(0, N, rows)):
                    hi = min(N, lo + rows)
                    b = i % 2
                    layer_d = up.upload(i, lo, hi, ev[b])
                    cov_d = cov_up.upload(i, lo, hi, ev[b]) if cov_up is not None else None
                    engine.build_obs(layer_d, cols_d, C, covered=cov_d, covered_cols=ccols_d, out_u8=out_u8,
                                     out=out[lo:hi], flag=flag)
                    ev[b] = torch.cuda.Event()
                    ev[b].record(cur)
            finally:
                up.finish()
                if cov_up is not None:
                    cov_up.finish()
                cur.synchronize()
        if int(flag.item()) == 0:
            return CodedU8(out) if out_u8 else out
    # float64 values that float32 does not represent: keep the reference's host gather, upload float64
    X = np.full(shape, np.nan, dtype=np.float64)
    flat_cols = np.asarray(cols).reshape(L, C)
    for c in range(C):
        ok = flat_cols[:, c] >= 0
        block = np.asarray(mat)[:, flat_cols[ok, c]].astype(np.float64)
        if C == 1:
            X[:, ok] = block
        else:
            X[:, ok, c] = block
    if covered is not None:
        cov = covered.view(np.bool_)[:, np.asarray(covered_cols)]
        X[~cov] = np.nan
    return torch.from_numpy(X).to(device)


def _coverage(adata, coordinates: np.ndarray):
    """(covered uint8 [N, V], column of every coordinate) or (None, None) (tools/partitioned_hmm.py:172-181)."""
    if COVERED_BASE_MASK not in adata.layers:
        return None, None
    positions = np.asarray(adata.var_names, dtype=np.int64)
    column_by_position = {int(position): index for index, position in enumerate(positions)}
    try:
        columns = np.asarray([column_by_position[int(position)] for position in coordinates], dtype=np.int32)
    except KeyError as exc:
        raise ValueError(f"HMM coordinate {exc.args[0]} is absent from materialized coverage") from exc
    covered = np.asarray(adata.layers[COVERED_BASE_MASK], dtype=bool)
    return covered, columns


def build_single_channel(subset, ref: str, methbase: str, smf_modality: Optional[str], cfg, *, device=None,
                         _coverage_of=None) -> Tuple[Any, np.ndarray]:
    """(X, coords) for one methylation channel (cli/hmm_adata.py:251-278).  ``X`` is a device-resident
    ``CodedU8`` (N, L) for binary calls, a float32 / float64 device tensor (NaN = missing) otherwise."""
    pm = resolve_pos_mask_for_methbase(subset, ref, methbase)
    if pm is None or int(np.sum(pm)) == 0:
        raise ValueError(f"No columns for methbase={methbase} on ref={ref}")
    mat = _source_matrix(subset, smf_modality, cfg)
    coords, _ = _safe_int_coords(subset.var_names[pm])
    dev = torch.device(device) if device is not None else _default_cuda_device()
    cols = np.flatnonzero(pm).astype(np.int32)
    covered = ccols = None
    if _coverage_of is not None:
        covered, ccols = _coverage(_coverage_of, coords)
    return _build_on_device(mat, cols.reshape(-1, 1), 1, dev, covered, ccols), coords


def build_multi_channel_union(subset, ref: str, methbases: Sequence[str], smf_modality: Optional[str], cfg, *,
                              device=None, _coverage_of=None) -> Tuple[Any, np.ndarray, List[str]]:
    """(X3, coords, used methbases) on the union coordinate grid (cli/hmm_adata.py:281-319): (N, L_union, C),
    missing where a channel has no site."""
    per = []
    for mb in methbases:
        pm = resolve_pos_mask_for_methbase(subset, ref, mb)
        if pm is None or int(np.sum(pm)) == 0:
            continue
        cmb, _ = _safe_int_coords(subset.var_names[pm])
        per.append((mb, np.flatnonzero(pm).astype(np.int64), cmb.astype(int)))
    if len(per) < 2:
        raise ValueError(f"Need >=2 methbases with columns for union multi-channel on ref={ref}")
    mat = _source_matrix(subset, smf_modality, cfg)
    coords = np.unique(np.concatenate([c for _, _, c in per], axis=0)).astype(int)
    C = len(per)
    cols = np.full((coords.shape[0], C), -1, dtype=np.int32)
    for ci, (_, grid_cols, cmb) in enumerate(per):
        cols[np.searchsorted(coords, cmb), ci] = grid_cols
    dev = torch.device(device) if device is not None else _default_cuda_device()
    covered = ccols = None
    if _coverage_of is not None:
        covered, ccols = _coverage(_coverage_of, coords)
    return _build_on_device(mat, cols, C, dev, covered, ccols), coords, [mb for mb, _, _ in per]


def mask_uncovered_model_input(adata, values, coordinates: np.ndarray):
    """Missing where ``covered_base_mask`` is False at the coordinate's column (tools/partitioned_hmm.py:170-188).
    ``values`` may be what the builders above return (stays on the device) or a host array."""
    covered, columns = _coverage(adata, np.asarray(coordinates))
    if covered is None:
        return values
    coded = isinstance(values, CodedU8)
    data = values.data if coded else values
    if isinstance(data, torch.Tensor):
        t = data
        dev = t.device if t.is_cuda else _default_cuda_device()
        t = t.to(dev)
    else:
        arr = np.asarray(data)
        dev = _default_cuda_device()
        if not coded and arr.dtype not in (np.float32, np.float64):
            arr = np.asarray(arr, dtype=float)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    N, L = int(t.shape[0]), int(t.shape[1])
    C = int(t.shape[2]) if t.dim() == 3 else 1
    if not coded and t.dtype == torch.uint8:
        t = t.to(torch.float32)
    cols = torch.arange(L * C, dtype=torch.int32, device=dev)
    cov_d = torch.from_numpy(np.ascontiguousarray(covered).view(np.uint8)).to(dev)
    ccols_d = torch.from_numpy(columns).to(dev)
    flat = t.contiguous().reshape(N, L * C)
    if coded:
        out, _ = engine.build_obs(flat, cols, C, coded_layer=True, covered=cov_d, covered_cols=ccols_d, out_u8=True)
        return CodedU8(out.reshape(t.shape))
    if t.dtype == torch.float32:
        out, _ = engine.build_obs(flat, cols, C, covered=cov_d, covered_cols=ccols_d, out_u8=False)
        return out.reshape(t.shape)
    # float64 values keep their precision: mask on the device with the gathered coverage columns
    keep = cov_d.index_select(1, ccols_d.to(torch.int64)) != 0
    out = t.clone()
    if out.dim() == 2:
        out[~keep] = float("nan")
    else:
        out[~keep, :] = float("nan")
    return out


def prepare_model_input(adata, reference: str, spec, cfg, *, device=None):
    """(values, coordinates, position_mask) of one configured HMM signal (tools/partitioned_hmm.py:140-167):
    builder + coverage mask in one device pass.  ``spec.signals`` lists the methylation channels."""
    smf_modality = str(getattr(cfg, "smf_modality", "conversion"))
    signals = tuple(spec.signals)
    if len(signals) == 1:
        signal = signals[0]
        values, coordinates = build_single_channel(adata, ref=reference, methbase=signal, smf_modality=smf_modality,
                                                   cfg=cfg, device=device, _coverage_of=adata)
        return values, coordinates, resolve_pos_mask_for_methbase(adata, reference, signal)
    values, coordinates, used = build_multi_channel_union(adata, ref=reference, methbases=signals,
                                                          smf_modality=smf_modality, cfg=cfg, device=device,
                                                          _coverage_of=adata)
    position_mask = None
    for signal in used:
        signal_mask = resolve_pos_mask_for_methbase(adata, reference, signal)
        position_mask = signal_mask if position_mask is None else position_mask | signal_mask
    return values, coordinates, position_mask
