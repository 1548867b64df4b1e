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
are expanded from that packed form on the host only because the AnnData contract asks for them;
callers that can consume the packed form (interval writers, reductions) skip the expansion.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine
from .hmm import (BaseHMM, _alloc_host_layers, _default_cuda_device, _device_to_host_layer, _safe_int_coords,
                  _span_inputs)

CODE_CLASS_MASK = 0x0F
CODE_MEMBER = 0x40
CODE_VALID = 0x80


def feature_ranges_for_merged_layer(core_layer: str, feature_sets: Mapping[str, Mapping]) -> Dict[str, Tuple[float, float]]:
    """Size classes belonging to an aggregate feature layer (cli/hmm_adata.py:774-779)."""
    for group_name, feature_set in feature_sets.items():
        if str(core_layer) == f"all_{group_name}_features":
            return feature_set.get("features", {}) or {}
    return {}


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
):
    """One pass of ``smf_hmm_merge_chain`` over ``adata.layers[base_layer]``.

    Returns ``(code, run_len)`` as device tensors (uint8 / int32, ``[n_obs, n_vars]``):
    ``code & 15`` = index of the size class of the merged run + 1 (order of ``feature_ranges``),
    ``code & 0x40`` = member of a merged run, ``code & 0x80`` = inside the read span.
    """
    if base_layer not in adata.layers:
        raise KeyError(f"Layer '{base_layer}' not found.")
    if start_key not in adata.obs or end_key not in adata.obs:
        raise KeyError(f"Missing {start_key!r} or {end_key!r} in adata.obs.")
    dev = torch.device(device) if device is not None else _default_cuda_device()
    coords, are_ints = _safe_int_coords(adata.var_names)
    member = (np.asarray(adata.layers[base_layer]) > 0).astype(np.uint8)  # NaN counts as 0 (HMM.py:1033)
    span = _span_inputs(adata, start_key=start_key, end_key=end_key,
                        use_original_var_names=use_original_var_names, device=dev)
    ranges = [tuple(feature_ranges[f]) for f in feature_ranges] if feature_ranges else []
    return engine.merge_chain(
        torch.from_numpy(np.ascontiguousarray(member)).to(dev),
        torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(dev),
        are_ints, int(distance_threshold), ranges,
        covered=span.get("covered"), span_coords=span.get("coords"), start=span.get("start"), end=span.get("end"),
    )


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
    """Write the reference's dense layers from the packed form; returns their names in the
    reference's creation order (merged, merged_lengths, then each class layer and its lengths).
    All are float64 with NaN outside the read span, as ``mask_layers_outside_read_span`` leaves them."""
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
    """``expand_merge_chain`` with the float64 / NaN expansion done on the device, a block of reads at a
    time, so every dense layer is written to host memory exactly once (the numpy version converts and
    masks each layer on one CPU core)."""
    n_obs, n_vars = int(code.shape[0]), int(code.shape[1])
    dev = code.device
    merged_name = f"{base_layer}{suffix}"
    feats = list(feature_ranges or {})
    names = [merged_name, f"{merged_name}_lengths"]
    for feat in feats:
        names.extend([f"{out_prefix}_{feat}{suffix}", f"{out_prefix}_{feat}{suffix}_lengths"])
    for nm, arr in _alloc_host_layers((n_obs, n_vars), {nm: np.float64 for nm in names}).items():
        adata.layers[nm] = arr
    nan = torch.tensor(float("nan"), dtype=torch.float64, device=dev)
    zero = torch.zeros((), dtype=run_len.dtype, device=dev)
    rows = rows_per_step or max(1, int((1 << 28) // max(n_vars, 1)))  # ~2 GB of float64 per block
    for lo in range(0, n_obs, rows):
        hi = min(n_obs, lo + rows)
        c = code[lo:hi]
        rl = run_len[lo:hi]
        valid = (c & CODE_VALID) != 0
        cls = c & CODE_CLASS_MASK

        def put(name, values):
            _device_to_host_layer(adata.layers[name][lo:hi], torch.where(valid, values.to(torch.float64), nan))

        put(merged_name, (c & CODE_MEMBER) != 0)
        put(f"{merged_name}_lengths", rl)
        for ci, feat in enumerate(feats):
            hit = cls == (ci + 1)
            put(f"{out_prefix}_{feat}{suffix}", hit)
            put(f"{out_prefix}_{feat}{suffix}_lengths", torch.where(hit, rl, zero))
    BaseHMM._register_appended_layers(adata, names)
    return names


def apply_merges(adata, model, prefix: str, feature_sets: dict, cfg) -> None:
    """Drop-in for ``_apply_merges(adata, model, prefix, feature_sets, cfg)``
    (tools/partitioned_hmm.py:78-108); ``model`` is accepted for signature parity only."""
    merged_suffix = str(getattr(cfg, "hmm_merged_suffix", "_merged")) if not isinstance(cfg, dict) else str(
        cfg.get("hmm_merged_suffix", "_merged"))
    pairs = (getattr(cfg, "hmm_merge_layer_features", []) if not isinstance(cfg, dict)
             else cfg.get("hmm_merge_layer_features", [])) or []
    for core_layer, distance in pairs:
        base_layer = f"{prefix}_{core_layer}"
        if base_layer not in adata.layers:
            continue
        feature_ranges = feature_ranges_for_merged_layer(core_layer, feature_sets)
        code, run_len = merge_chain_packed(adata, base_layer, distance_threshold=int(distance),
                                           feature_ranges=feature_ranges)
        expand_merge_chain_device(adata, base_layer, code, run_len, out_prefix=prefix,
                                  feature_ranges=feature_ranges, suffix=merged_suffix)
