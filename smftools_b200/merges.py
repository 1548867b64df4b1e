"""Fused merge chain: host mirror of the step that follows ``annotate_adata`` in the reference.

``apply_merges`` has the signature and the end state of ``_apply_merges``
(/root/reference/src/smftools/tools/partitioned_hmm.py:78-108): for every configured
``(core_layer, distance)`` pair it merges neighbouring runs of ``{prefix}_{core_layer}``
(``merge_intervals_to_new_layer``, HMM.py:1006-1067), re-bins the merged runs into the size classes
of the owning feature group (``write_size_class_layers_from_binary``, HMM.py:1069-1123;
``_feature_ranges_for_merged_layer``, cli/hmm_adata.py:774-779) and NaN-masks every layer of the
chain outside each read's span (``mask_layers_outside_read_span``, HMM.py:148-231).

The reference walks the layer three times in per-read Python loops and materialises 2 + 2·classes
dense layers in between.  Here the layer is uploaded once, ``smf_hmm_merge_chain`` produces one code
byte + one int32 run length per cell in a single pass (``merge_chain_packed``), and the AnnData layers
are expanded from that packed form (``smf_hmm_expand_layers``, one pass per block of reads, straight into
page-locked host layers) only because the AnnData contract asks for them; callers that can consume the
packed form (interval writers, reductions) skip the expansion.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine, hostmem
from . import stream as _stream_mod
from .hmm import BaseHMM, _alloc_host_layers, _default_cuda_device, _safe_int_coords, _span_inputs

CODE_CLASS_MASK = 0x0F
CODE_MEMBER = 0x40
CODE_VALID = 0x80


def feature_ranges_for_merged_layer(core_layer: str, feature_sets: Mapping[str, Mapping]) -> Dict[str, Tuple[float, float]]:
    """Size classes belonging to an aggregate feature layer (cli/hmm_adata.py:774-779)."""
    for group_name, feature_set in feature_sets.items():
        if str(core_layer) == f"all_{group_name}_features":
            return feature_set.get("features", {}) or {}
    return {}


def _member_rows(layer: np.ndarray, lo: int, hi: int, out: np.ndarray) -> np.ndarray:
    """``(layer[lo:hi] > 0)`` as uint8 into the (page-locked) staging array ``out``, split over the host
    threads (NaN counts as 0, HMM.py:1033)."""
    n = hi - lo
    dst = out[:n].view(np.bool_)
    pool, nthr = hostmem.workers()
    if nthr < 2 or layer[lo:hi].nbytes < (8 << 20):
        np.greater(layer[lo:hi], 0, out=dst)
        return out[:n]
    step = (n + nthr - 1) // nthr
    list(pool.map(lambda a: np.greater(layer[lo + a:min(hi, lo + a + step)], 0, out=dst[a:a + step]),
                  range(0, n, step)))
    return out[:n]


def _chain_rows(n_obs: int, n_vars: int, n_layers: int, budget: int) -> int:
    """Reads per device block: member 1 + covered 1 + code 1 + run_len 4 bytes per cell plus two staging
    slots of ``n_layers`` float64 layers."""
    per_row = n_vars * (7 + 2 * 8 * max(n_layers, 1)) + 64
    return max(1, min(n_obs, int(budget // per_row)))


def merge_chain_packed(
    adata,
    base_layer: str,
    *,
    distance_threshold: int,
    feature_ranges: Optional[Mapping[str, Tuple[float, float]]] = None,
    start_key: str = "reference_start",
    end_key: str = "reference_end",
    use_original_var_names: bool = True,
    device=None,
    rows: Optional[slice] = None,
):
    """One pass of ``smf_hmm_merge_chain`` over ``adata.layers[base_layer]`` (or the read block ``rows``).

    Returns ``(code, run_len)`` as device tensors (uint8 / int32, ``[n_rows, n_vars]``):
    ``code & 15`` = index of the size class of the merged run + 1 (order of ``feature_ranges``),
    ``code & 0x40`` = member of a merged run, ``code & 0x80`` = inside the read span.
    Whole-matrix calls put 7 bytes per cell on the device; ``apply_merges`` walks large inputs in blocks.
    """
    if base_layer not in adata.layers:
        raise KeyError(f"Layer '{base_layer}' not found.")
    if start_key not in adata.obs or end_key not in adata.obs:
        raise KeyError(f"Missing {start_key!r} or {end_key!r} in adata.obs.")
    dev = torch.device(device) if device is not None else _default_cuda_device()
    coords, are_ints = _safe_int_coords(adata.var_names)
    lo, hi = (0, int(adata.n_obs)) if rows is None else (int(rows.start or 0), int(rows.stop))
    layer = np.asarray(adata.layers[base_layer])
    member = np.empty((hi - lo, int(adata.n_vars)), dtype=np.uint8)
    _member_rows(layer, lo, hi, member)
    span = _span_inputs(adata, start_key=start_key, end_key=end_key,
                        use_original_var_names=use_original_var_names, device=dev)
    ranges = [tuple(feature_ranges[f]) for f in feature_ranges] if feature_ranges else []
    return _chain_block(member, coords, are_ints, int(distance_threshold), ranges, span, lo, hi, dev)


def _chain_block(member: np.ndarray, coords, are_ints, distance, ranges, span, lo, hi, dev, coords_d=None):
    cov = span.get("covered")
    return engine.merge_chain(
        torch.from_numpy(member).to(dev),
        coords_d if coords_d is not None else torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(dev),
        are_ints, distance, ranges,
        covered=None if cov is None else torch.from_numpy(cov[lo:hi]).to(dev),
        span_coords=span.get("coords"),
        start=None if span.get("start") is None else span["start"][lo:hi].contiguous(),
        end=None if span.get("end") is None else span["end"][lo:hi].contiguous(),
    )


def _chain_layer_plan(base_layer: str, out_prefix: str, feats: Sequence[str], suffix: str):
    merged_name = f"{base_layer}{suffix}"
    plan = [(merged_name, _lib.LAYER_ANY, 0), (f"{merged_name}_lengths", _lib.LAYER_ANY_LEN, 0)]
    for ci, feat in enumerate(feats):
        plan.append((f"{out_prefix}_{feat}{suffix}", _lib.LAYER_CLASS, ci))
        plan.append((f"{out_prefix}_{feat}{suffix}_lengths", _lib.LAYER_CLASS_LEN, ci))
    return plan


def expand_merge_chain(
    adata,
    base_layer: str,
    code: np.ndarray,
    run_len: np.ndarray,
    *,
    out_prefix: str,
    feature_ranges: Optional[Mapping[str, Tuple[float, float]]] = None,
    suffix: str = "_merged",
) -> List[str]:
    """Write the reference's dense layers from the packed form on the HOST (numpy; for callers that hold
    the packed arrays in host memory); returns their names in the reference's creation order (merged,
    merged_lengths, then each class layer and its lengths).  All are float64 with NaN outside the read
    span, as ``mask_layers_outside_read_span`` leaves them."""
    code = np.asarray(code)
    run_len = np.asarray(run_len)
    drop = (code & CODE_VALID) == 0
    member = (code & CODE_MEMBER) != 0
    cls = code & CODE_CLASS_MASK

    def _masked(values: np.ndarray) -> np.ndarray:
        out = values.astype(np.float64)
        out[drop] = np.nan
        return out

    merged_name = f"{base_layer}{suffix}"
    names = [merged_name, f"{merged_name}_lengths"]
    adata.layers[merged_name] = _masked(member)
    adata.layers[f"{merged_name}_lengths"] = _masked(run_len)
    for ci, feat in enumerate(feature_ranges or {}):
        nm = f"{out_prefix}_{feat}{suffix}"
        hit = cls == (ci + 1)
        adata.layers[nm] = _masked(hit)
        adata.layers[f"{nm}_lengths"] = _masked(np.where(hit, run_len, 0))
        names.extend([nm, f"{nm}_lengths"])
    BaseHMM._register_appended_layers(adata, names)
    return names


def expand_merge_chain_device(
    adata,
    base_layer: str,
    code: torch.Tensor,
    run_len: torch.Tensor,
    *,
    out_prefix: str,
    feature_ranges: Optional[Mapping[str, Tuple[float, float]]] = None,
    suffix: str = "_merged",
    rows_per_step: int = 0,
) -> List[str]:
    """``expand_merge_chain`` for packed DEVICE tensors: ``smf_hmm_expand_layers`` writes every final layer
    of a block of reads (float64, NaN outside the read span) in one pass; the blocks land in page-locked
    host layers through the copy engine while the next block is expanded."""
    n_obs, n_vars = int(code.shape[0]), int(code.shape[1])
    dev = code.device
    feats = list(feature_ranges or {})
    plan = _chain_layer_plan(base_layer, out_prefix, feats, suffix)
    names = [p[0] for p in plan]
    for nm, arr in _alloc_host_layers((n_obs, n_vars), {nm: np.float64 for nm in names}).items():
        adata.layers[nm] = arr
    rows = rows_per_step or max(1, min(n_obs, int((256 << 20) // max(8 * len(plan) * n_vars, 1))))
    cur = torch.cuda.current_stream(dev)
    s_out = _stream_mod.side_stream(dev, "d2h")
    entry = torch.cuda.Event()
    entry.record(cur)
    s_out.wait_event(entry)
    stage = [torch.empty((len(plan), rows, n_vars), dtype=torch.float64, device=dev) for _ in range(2)]
    ev_out = [None, None]
    try:
        for i, lo in enumerate(range(0, n_obs, rows)):
            hi = min(n_obs, lo + rows)
            b = i % 2
            if ev_out[b] is not None:
                cur.wait_event(ev_out[b])
            _expand_block(code[lo:hi], run_len[lo:hi], plan, stage[b], adata, lo, hi, cur, s_out)
            ev_out[b] = torch.cuda.Event()
            ev_out[b].record(s_out)
    finally:
        s_out.synchronize()
        cur.synchronize()
    BaseHMM._register_appended_layers(adata, names)
    return names


def _expand_block(code_blk, len_blk, plan, stage_slot, adata, lo, hi, cur, s_out):
    n = hi - lo
    n_vars = int(code_blk.shape[1])
    outs = [stage_slot[i, :n] for i in range(len(plan))]
    engine.expand_layers(n, n_vars, 0, states=None, post=None, post_stride=1, site_of_col=None, valid=None,
                         groups=[(code_blk.contiguous(), len_blk.contiguous(), True)],
                         layers=[(kind, 0, ci, out) for (nm, kind, ci), out in zip(plan, outs)])
    ev = torch.cuda.Event()
    ev.record(cur)
    s_out.wait_event(ev)
    for (nm, _k, _c), out in zip(plan, outs):
        engine.copy_to_host_async(adata.layers[nm][lo:hi], out, s_out)


def apply_merges(adata, model, prefix: str, feature_sets: dict, cfg) -> None:
    """Drop-in for ``_apply_merges(adata, model, prefix, feature_sets, cfg)``
    (tools/partitioned_hmm.py:78-108); ``model`` is accepted for signature parity only.

    The chain runs a block of reads at a time (sized from ``BaseHMM.device_chunk_bytes``; reads are
    independent): host threads reduce the block of the source layer to its membership bytes, the fused
    kernel produces the packed code, the expand kernel the final float64 / NaN layers, and the copy engine
    writes them into page-locked host layers while the next block is processed."""
    merged_suffix = str(getattr(cfg, "hmm_merged_suffix", "_merged")) if not isinstance(cfg, dict) else str(
        cfg.get("hmm_merged_suffix", "_merged"))
    pairs = (getattr(cfg, "hmm_merge_layer_features", []) if not isinstance(cfg, dict)
             else cfg.get("hmm_merge_layer_features", [])) or []
    dev = None
    for core_layer, distance in pairs:
        base_layer = f"{prefix}_{core_layer}"
        if base_layer not in adata.layers:
            continue
        if "reference_start" not in adata.obs or "reference_end" not in adata.obs:
            raise KeyError("Missing 'reference_start' or 'reference_end' in adata.obs.")
        if dev is None:
            dev = _default_cuda_device()
        feature_ranges = feature_ranges_for_merged_layer(core_layer, feature_sets)
        feats = list(feature_ranges or {})
        ranges = [tuple(feature_ranges[f]) for f in feats]
        if len(ranges) > 15:
            raise ValueError("at most 15 size classes fit the packed code")
        n_obs, n_vars = int(adata.n_obs), int(adata.n_vars)
        plan = _chain_layer_plan(base_layer, prefix, feats, merged_suffix)
        names = [p[0] for p in plan]
        layer = np.asarray(adata.layers[base_layer])
        coords, are_ints = _safe_int_coords(adata.var_names)
        coords_d = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(dev)
        span = _span_inputs(adata, start_key="reference_start", end_key="reference_end",
                            use_original_var_names=True, device=dev)
        out_layers = _alloc_host_layers((n_obs, n_vars), {nm: np.float64 for nm in names})
        if n_obs == 0 or n_vars == 0:
            adata.layers.update(out_layers)
            BaseHMM._register_appended_layers(adata, names)
            continue
        rows = _chain_rows(n_obs, n_vars, len(plan), min(BaseHMM.device_chunk_bytes, 1 << 30))
        cur = torch.cuda.current_stream(dev)
        s_out = _stream_mod.side_stream(dev, "d2h")
        entry = torch.cuda.Event()
        entry.record(cur)
        s_out.wait_event(entry)
        stage = [torch.empty((len(plan), rows, n_vars), dtype=torch.float64, device=dev) for _ in range(2)]
        member_stage = [hostmem.pool().empty((rows, n_vars), np.uint8) for _ in range(2)]
        ev_up = [None, None]
        ev_out = [None, None]

        class _Out:  # _expand_block writes into the new layers, not into adata (the base layer may be replaced)
            layers = out_layers

        try:
            for i, lo in enumerate(range(0, n_obs, rows)):
                hi = min(n_obs, lo + rows)
                b = i % 2
                if ev_up[b] is not None:
                    ev_up[b].synchronize()
                member = _member_rows(layer, lo, hi, member_stage[b])
                code, run_len = _chain_block(member, coords, are_ints, int(distance), ranges, span, lo, hi, dev,
                                             coords_d=coords_d)
                ev_up[b] = torch.cuda.Event()
                ev_up[b].record(cur)
                if ev_out[b] is not None:
                    cur.wait_event(ev_out[b])
                _expand_block(code, run_len, plan, stage[b], _Out, lo, hi, cur, s_out)
                ev_out[b] = torch.cuda.Event()
                ev_out[b].record(s_out)
                del code, run_len
        finally:
            s_out.synchronize()
            cur.synchronize()
        for nm in names:
            adata.layers[nm] = out_layers[nm]
        BaseHMM._register_appended_layers(adata, names)
