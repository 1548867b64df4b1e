"""Host-side mirror of the reference's HMM operator interface, backed by sm_100a kernels.

Drop-in for ``smftools.hmm.HMM`` (reference ``src/smftools/hmm/HMM.py``): the same registry
(``register_hmm`` / ``create_hmm`` / ``HMM`` facade, HMM.py:27-54, 2398-2432), the same three
architectures (``single`` :1411, ``multi`` :1673, ``single_distance_binned`` :1984), the same
``fit`` / ``adapt_emissions`` / ``decode`` / ``annotate_adata`` / ``merge_intervals_to_new_layer``
/ ``write_size_class_layers_from_binary`` / ``save`` / ``load`` signatures, ``state_dict`` keys
(``start``, ``trans``, ``emission``[, ``trans_by_bin``]) and AnnData layer outputs.

What differs is where the arithmetic runs: forward-backward, Viterbi, the Baum-Welch E-step and
run-length feature calling execute in ``libsmfhmm.so`` on a B200.  There is NO CPU fallback: a
compute call without a CUDA device raises.  Only the O(K^2) M-step and the log tables are
formed on the host, with the same torch ops (and therefore the same rounding) as the reference.
"""
from __future__ import annotations

import os
import time

import ast
import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _lib, engine, hostmem
from . import stream as _stream_mod

COVERED_BASE_MASK = "covered_base_mask"  # reference constants.py:141

# ----------------------------------------------------------------------------------------
# registry / factory  (HMM.py:27-54)
# ----------------------------------------------------------------------------------------
_HMM_REGISTRY: Dict[str, type] = {}


def register_hmm(name: str):
    """Class decorator: make ``create_hmm(cfg, arch=name)`` build this class."""

    def _wrap(cls):
        cls.hmm_name = name
        _HMM_REGISTRY[name] = cls
        return cls

    return _wrap


def create_hmm(cfg: Union[dict, Any, None], arch: Optional[str] = None, **kwargs):
    """Resolve ``arch`` (argument, then ``cfg.hmm_arch``, then "single") and build the model."""
    key = arch
    if not key:
        key = getattr(cfg, "hmm_arch", None)
    if not key and isinstance(cfg, dict):
        key = cfg.get("hmm_arch")
    if not key:
        key = "single"
    try:
        klass = _HMM_REGISTRY[key]
    except KeyError:
        raise KeyError(f"Unknown hmm_arch={key!r}. Known: {sorted(_HMM_REGISTRY.keys())}") from None
    return klass.from_config(cfg, **kwargs)


def install_into_reference(ref_module, names: Sequence[str] = ("single", "multi", "single_distance_binned")):
    """Overwrite the reference registry entries so its callers (HMMTrainer, partitioned_hmm)
    build B200 models unchanged.  ``ref_module`` is the imported ``smftools.hmm.HMM``."""
    for nm in names:
        ref_module._HMM_REGISTRY[nm] = _HMM_REGISTRY[nm]
    return ref_module


# ----------------------------------------------------------------------------------------
# small utilities shared with the reference's callers (HMM.py:60-146, 276-306)
# ----------------------------------------------------------------------------------------
def _parse_maybe_text(x: Any) -> Any:
    if not isinstance(x, str):
        return x
    s = x.strip()
    if not s:
        return None
    for parser in (json.loads, ast.literal_eval):
        try:
            return parser(s)
        except Exception:
            continue
    return x


def _resolve_dtype(entry: Any) -> torch.dtype:
    if entry is None:
        return torch.float64
    if isinstance(entry, torch.dtype):
        return entry
    s = str(entry).lower()
    if "16" in s:
        return torch.float16
    if "32" in s:
        return torch.float32
    return torch.float64


def _safe_int_coords(var_names) -> Tuple[np.ndarray, bool]:
    """Integer coordinates from var names, else 0..n-1 and ``False``."""
    try:
        return np.asarray(var_names, dtype=int), True
    except Exception:
        return np.arange(len(var_names), dtype=int), False


def _to_dense_np(x):
    if x is None:
        return None
    if hasattr(x, "toarray") and not isinstance(x, np.ndarray):
        return x.toarray()
    return np.asarray(x)


def _ensure_2d_np(x):
    x = _to_dense_np(x)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if x.ndim != 2:
        raise ValueError(f"Expected 2D array; got shape {x.shape}")
    return x


def normalize_hmm_feature_sets(raw: Any) -> Dict[str, Dict[str, Any]]:
    """Canonical ``{group: {"state": ..., "features": {name: (lo, hi)}}}`` (HMM.py:314-390)."""
    parsed = _parse_maybe_text(raw)
    if not isinstance(parsed, dict):
        return {}

    def bound(v):
        if v is None:
            return None
        if isinstance(v, (int, float)):
            return float(v)
        s = str(v).strip().lower()
        if s in ("inf", "infty", "np.inf", "infinite"):
            return np.inf
        if s in ("none", ""):
            return None
        try:
            return float(v)
        except Exception:
            return None

    def ranges(feats):
        out = {}
        if not isinstance(feats, dict):
            return out
        for name, rng in feats.items():
            if rng is None:
                out[name] = (0.0, np.inf)
            elif isinstance(rng, (list, tuple)) and len(rng) >= 2:
                lo, hi = bound(rng[0]), bound(rng[1])
                out[name] = (0.0 if lo is None else float(lo), np.inf if hi is None else float(hi))
            else:
                hi = bound(rng)
                out[name] = (0.0, np.inf if hi is None else float(hi))
        return out

    result: Dict[str, Dict[str, Any]] = {}
    for group, info in parsed.items():
        if isinstance(info, dict):
            feats = ranges(info.get("features", info.get("ranges", {})))
            state = info.get("state", info.get("label", "Modified"))
        else:
            feats = ranges(info)
            state = "Modified"
        result[group] = {"features": feats, "state": state}
    return result


def _default_cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "smftools_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback. "
            "Use the reference smftools.hmm for CPU execution."
        )
    return torch.device("cuda", torch.cuda.current_device())


def _span_inputs(adata, *, start_key: str, end_key: str, use_original_var_names: bool, device):
    """Inputs of the read-span test (HMM.py:178-226): ``{"covered": u8 [N,V] (HOST numpy)}`` when the
    coverage mask layer exists, else device tensors ``{"coords": i64 [V], "start": f64 [N], "end": f64 [N]}``."""
    coord_source = adata.var_names
    if use_original_var_names and "Original_var_names" in adata.var:
        orig = np.asarray(adata.var["Original_var_names"])
        if orig.size == adata.n_vars:
            try:
                numeric = np.asarray(orig, dtype=float)
            except (TypeError, ValueError):
                numeric = None
            if numeric is not None and np.isfinite(numeric).any():
                coord_source = orig
    coords, _ = _safe_int_coords(coord_source)
    if coords.shape[0] != adata.n_vars:
        raise ValueError("Coordinate source length does not match adata.n_vars.")
    if COVERED_BASE_MASK in adata.layers:
        covered = np.asarray(adata.layers[COVERED_BASE_MASK], dtype=bool)
        if covered.shape != tuple(adata.shape):
            raise ValueError("covered_base_mask must match adata shape")
        # host array: callers upload it a block of reads at a time (it is as large as a layer)
        return {"covered": np.ascontiguousarray(covered).view(np.uint8)}
    try:
        starts = np.asarray(adata.obs[start_key], dtype=float)
        ends = np.asarray(adata.obs[end_key], dtype=float)
    except (TypeError, ValueError) as exc:
        raise ValueError("Start/end positions must be numeric.") from exc
    return {
        "coords": torch.from_numpy(np.array(coords, dtype=np.int64, copy=True)).to(device),
        "start": torch.from_numpy(np.array(starts, dtype=np.float64, copy=True)).to(device),
        "end": torch.from_numpy(np.array(ends, dtype=np.float64, copy=True)).to(device),
    }


def _span_validity(adata, *, start_key: str, end_key: str, use_original_var_names: bool) -> np.ndarray:
    """bool [n_obs, n_vars]: True where a layer value is kept (HMM.py:178-226)."""
    dev = _default_cuda_device()
    inp = _span_inputs(adata, start_key=start_key, end_key=end_key,
                       use_original_var_names=use_original_var_names, device=dev)
    n_obs, n_vars = int(adata.n_obs), int(adata.n_vars)
    out = np.empty((n_obs, n_vars), dtype=bool)
    rows = max(1, int((1 << 30) // max(n_vars, 1)))
    for lo in range(0, n_obs, rows):
        hi = min(n_obs, lo + rows)
        if "covered" in inp:
            valid = engine.span_mask(hi - lo, n_vars, dev, covered=torch.from_numpy(inp["covered"][lo:hi]).to(dev))
        else:
            valid = engine.span_mask(hi - lo, n_vars, dev, coords=inp["coords"], start=inp["start"][lo:hi].contiguous(),
                                     end=inp["end"][lo:hi].contiguous())
        out[lo:hi] = valid.cpu().numpy().astype(bool)
    return out


def mask_layers_outside_read_span(
    adata,
    layers: Sequence[str],
    *,
    start_key: str = "reference_start",
    end_key: str = "reference_end",
    use_original_var_names: bool = True,
) -> List[str]:
    """NaN-mask layer cells outside each read's reference span (HMM.py:148-231).

    Non-float layers become float64.  The validity bitmap comes from ``smf_hmm_span_mask``.
    """
    if not layers:
        return []
    if start_key not in adata.obs or end_key not in adata.obs:
        raise KeyError(f"Missing {start_key!r} or {end_key!r} in adata.obs.")
    for layer in layers:
        if layer not in adata.layers:
            raise KeyError(f"Layer {layer!r} not found in adata.layers.")
    valid = _span_validity(adata, start_key=start_key, end_key=end_key,
                           use_original_var_names=use_original_var_names)
    drop = ~valid
    any_drop = bool(drop.any())
    done = []
    for layer in layers:
        arr = np.asarray(adata.layers[layer])
        if not np.issubdtype(arr.dtype, np.floating):
            arr = arr.astype(float, copy=True)
        if any_drop:
            arr[drop] = np.nan
        adata.layers[layer] = arr
        done.append(layer)
    return done


def _alloc_host_arrays(specs: Dict[str, Tuple[Tuple[int, ...], Any]]) -> Dict[str, np.ndarray]:
    """Result arrays ``name -> (shape, dtype)`` carved from the page-locked host pool (``hostmem``): their
    pages are first touched by several threads and registered with the driver once, device-to-host copies
    land in them at link speed, and the memory is reused by later calls once the caller drops the arrays."""
    return hostmem.pool().empty_many(specs)


def _alloc_host_layers(shape, dtypes: Dict[str, Any]) -> Dict[str, np.ndarray]:
    """Dense (n_obs, n_vars) layers with pre-touched pages (see ``_alloc_host_arrays``)."""
    return _alloc_host_arrays({nm: (tuple(shape), dt) for nm, dt in dtypes.items()})


def _device_to_host_layer(dst: np.ndarray, src: torch.Tensor) -> None:
    """Copy a device tensor straight into (a row block of) a host layer -- no intermediate host tensor."""
    torch.from_numpy(dst).copy_(src)


class CodedU8:
    """Opt-in marker for observations already in the kernels' compact code: uint8 with 0 = unmodified,
    1 = modified, anything else = missing (1 byte per position over the host link instead of 8).

    A *plain* uint8 array keeps the reference's meaning, ``np.asarray(X, dtype=float)`` (HMM.py:694): the
    byte value itself is the observation.  ``smftools_b200.build_obs`` / the input builders return this
    wrapper; ``CodedU8(array)`` wraps an array the caller coded itself."""

    __slots__ = ("data",)

    def __init__(self, data):
        dt = data.dtype
        if dt not in (np.uint8, torch.uint8):
            raise ValueError("CodedU8 wraps uint8 data")
        self.data = data

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.data.shape)

    def __len__(self):
        return int(self.data.shape[0])

    def __getitem__(self, idx):
        """Row / column selection keeps the marker (``CodedU8(x)[lo:hi]`` is the coded rows ``lo:hi``)."""
        return CodedU8(self.data[idx])


def _run_lengths_np(bin_mat: np.ndarray) -> np.ndarray:
    """Per-cell length of the True-run it belongs to (int32); vectorised HMM.py:947-959."""
    b = np.asarray(bin_mat).astype(bool)
    n, L = b.shape
    out = np.zeros((n, L), dtype=np.int32)
    if L == 0 or n == 0:
        return out
    pad = np.zeros((n, L + 2), dtype=np.int8)
    pad[:, 1:-1] = b
    d = np.diff(pad, axis=1)
    rs, cs = np.nonzero(d == 1)
    re_, ce = np.nonzero(d == -1)
    lens = (ce - cs).astype(np.int32)
    # expand: for each run write its length over [cs, ce)
    if lens.size:
        flat = out.reshape(-1)
        starts = rs.astype(np.int64) * L + cs
        idx = np.repeat(starts, lens) + (np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens))
        flat[idx] = np.repeat(lens, lens)
    return out


# ----------------------------------------------------------------------------------------
# BaseHMM
# ----------------------------------------------------------------------------------------
class BaseHMM(nn.Module):
    """Shared parameters (``start``, ``trans``), decoding, EM driver and AnnData annotation."""

    #: reads per device chunk are sized so that inputs + outputs stay under this many bytes
    device_chunk_bytes = 6 << 30

    def __init__(self, n_states: int = 2, eps: float = 1e-8, dtype: torch.dtype = torch.float64):
        super().__init__()
        if n_states < 2:
            raise ValueError("n_states must be >= 2")
        if n_states > _lib.MAX_STATES:
            raise ValueError(f"smftools_b200 supports at most {_lib.MAX_STATES} states")
        self.n_states = int(n_states)
        self.eps = float(eps)
        self.dtype = dtype
        K = self.n_states
        self.start = nn.Parameter(torch.full((K,), 1.0 / K, dtype=dtype), requires_grad=False)
        self.trans = nn.Parameter(torch.full((K, K), 1.0 / K, dtype=dtype), requires_grad=False)
        self._normalize_params()
        #: optional torch.distributed process group: E-step statistics are all-reduced over it
        self.process_group = None
        self.distributed = False

    # ---- config -------------------------------------------------------------------------
    @staticmethod
    def _cfg_to_dict(cfg: Union[dict, Any, None]) -> dict:
        if cfg is None:
            return {}
        if isinstance(cfg, dict):
            return dict(cfg)
        to_dict = getattr(cfg, "to_dict", None)
        if callable(to_dict):
            return dict(to_dict())
        found = {}
        for key in dir(cfg):
            if key.startswith("hmm_") or key in ("smf_modality", "cpg"):
                try:
                    found[key] = getattr(cfg, key)
                except Exception:
                    pass
        return found

    @classmethod
    def _merged_cfg(cls, cfg, override):
        merged = cls._cfg_to_dict(cfg)
        if override:
            merged.update(override)
        return merged

    @classmethod
    def from_config(cls, cfg: Union[dict, Any, None], *, override: Optional[dict] = None, device=None):
        merged = cls._merged_cfg(cfg, override)
        model = cls(
            n_states=int(merged.get("hmm_n_states", merged.get("n_states", 2))),
            eps=float(merged.get("hmm_eps", merged.get("eps", 1e-8))),
            dtype=_resolve_dtype(merged.get("hmm_dtype", merged.get("dtype", None))),
        )
        return cls._finish_from_config(model, merged, device)

    @staticmethod
    def _finish_from_config(model, merged, device):
        if device is not None:
            model.to(torch.device(device) if isinstance(device, str) else device)
        model._persisted_cfg = merged
        return model

    # ---- parameters ----------------------------------------------------------------------
    def _normalize_params(self):
        """``+eps`` then renormalise start and every transition row (HMM.py:492-511)."""
        with torch.no_grad():
            K = self.n_states
            s = self.start.data.reshape(-1)
            if s.numel() != K:
                s = torch.full((K,), 1.0 / K, dtype=self.dtype, device=self.start.device)
            s = s + self.eps
            self.start.data = s / s.sum()
            t = self.trans.data.reshape(K, K) + self.eps
            rows = t.sum(dim=1, keepdim=True)
            rows[rows == 0.0] = 1.0
            self.trans.data = t / rows

    def _ensure_device_dtype(self, device=None) -> torch.device:
        """Resolve the CUDA compute device and coerce parameter dtype.

        Parameters are tiny; they stay where ``.to()`` put them.  Returns the CUDA device the
        kernels run on; raises if that is not a CUDA device (no CPU fallback).
        """
        if device is None:
            pdev = next(self.parameters()).device
            device = pdev if pdev.type == "cuda" else _default_cuda_device()
        device = torch.device(device) if isinstance(device, str) else device
        if device.type != "cuda":
            raise RuntimeError(
                f"smftools_b200 HMM kernels run only on CUDA devices (got device={device!s}); "
                "there is no CPU fallback"
            )
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if not torch.cuda.is_available():
            _default_cuda_device()
        for p in self._param_list():
            p.data = p.data.to(dtype=self.dtype)
        return device

    def _param_list(self):
        return [self.start, self.trans]

    # ---- state labelling (HMM.py:533-558) --------------------------------------------------
    def _state_modified_score(self) -> torch.Tensor:
        raise NotImplementedError

    def modified_state_index(self) -> int:
        return int(torch.argmax(self._state_modified_score()).item())

    def resolve_target_state_index(self, state_target: Any) -> int:
        if isinstance(state_target, (int, np.integer)):
            return max(0, min(int(state_target), self.n_states - 1))
        s = str(state_target).strip().lower()
        if s in ("non-modified", "closed", "inaccessible", "0", "neg", "negative"):
            return int(torch.argmin(self._state_modified_score()).item())
        return self.modified_state_index()

    # ---- tables for the kernels ------------------------------------------------------------
    def _emission_matrix(self) -> torch.Tensor:
        """(K, C) emission probabilities on the CPU in ``self.dtype``."""
        raise NotImplementedError

    def _trans_stack(self) -> torch.Tensor:
        """(n_bins, K, K) transition matrices on the CPU."""
        return self.trans.detach().to("cpu", self.dtype).reshape(1, self.n_states, self.n_states)

    def _host_tables(self) -> _lib.HostTables:
        """Log tables formed with the reference's torch ops on the CPU (HMM.py:593-595,
        1494-1496, 1773-1775, 2149-2150), widened to fp64 for the C-ABI."""
        eps = float(self.eps)
        start = self.start.detach().to("cpu", self.dtype)
        trans = self._trans_stack()
        p = self._emission_matrix()
        log_start = torch.log(start + eps)
        log_trans = torch.log(trans + eps)
        log_e1 = torch.log(p + eps)
        log_e0 = torch.log1p(-p + eps)
        f64 = lambda t: t.to(torch.float64).numpy()  # noqa: E731
        return _lib.HostTables(f64(log_start), f64(log_trans), f64(log_e1), f64(log_e0))

    def _bins_for(self, coords: Optional[np.ndarray], L: int, device) -> Optional[torch.Tensor]:
        return None

    # ---- input staging ---------------------------------------------------------------------
    def _expected_ndim(self) -> Optional[int]:
        return None

    @staticmethod
    def _as_obs_source(X):
        """Accept numpy / torch input; returns ``(array, numeric_u8)``.  Float64 (the reference's host
        dtype) and float32 go to the device unchanged; ``CodedU8`` input is the kernels' 0 / 1 / missing
        code; plain uint8 keeps its numeric meaning like ``np.asarray(X, dtype=float)`` (HMM.py:694) -- it
        crosses the link as bytes and is widened on the device (``numeric_u8``); anything else is widened
        to float64 on the host."""
        if isinstance(X, CodedU8):
            return X.data, False
        if isinstance(X, torch.Tensor):
            if X.dtype == torch.uint8:
                return X, True
            if X.dtype not in (torch.float32, torch.float64):
                X = X.to(torch.float64)
            return X, False
        X = np.asarray(X)
        if X.dtype == np.uint8:
            return X, True
        if X.dtype not in (np.float32, np.float64):
            X = np.asarray(X, dtype=float)
        return X, False

    def _chunks(self, N: int, L: int, C: int, in_item: int, out_bytes_per_pos: int):
        per_read = L * (C * in_item + out_bytes_per_pos) + 64
        rows = max(1, int(self.device_chunk_bytes // per_read))
        for lo in range(0, N, rows):
            yield lo, min(N, lo + rows)

    @staticmethod
    def _to_device(chunk, device, numeric_u8: bool = False) -> torch.Tensor:
        if isinstance(chunk, torch.Tensor):
            t = chunk.to(device).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(chunk)).to(device)
        if numeric_u8 and t.dtype == torch.uint8:
            t = t.to(torch.float32)  # byte value = observation (exact in fp32)
        return t

    # ---- decoding (HMM.py:673-720) -----------------------------------------------------------
    def decode(
        self,
        X,
        coords: Optional[np.ndarray] = None,
        *,
        decode: str = "marginal",
        device: Optional[Union[str, torch.device]] = None,
        compact: bool = False,
        out: Optional[Tuple[np.ndarray, np.ndarray]] = None,
    ) -> Tuple[np.ndarray, np.ndarray]:
        """States (N,L) and posteriors (N,L,K).

        Like the reference, the posterior is always computed; ``decode="viterbi"`` replaces the
        arg-max states with the Viterbi path.  Returns ``(int64, model dtype)`` host arrays, or the
        kernels' native ``(int8, float32)`` when ``compact=True``.  ``out=(states, gamma)`` lets the
        caller supply result buffers (page-locked in place for the call when they are not already).
        Otherwise the results are fresh arrays from the page-locked host pool (``hostmem``).  Host
        inputs are streamed through the device in read chunks: staging + H2D copy, kernels (incl. the
        widening to int64 / float64) and D2H copy of consecutive chunks overlap (``stream.py``).
        """
        device = self._ensure_device_dtype(device)
        src, numeric_u8 = self._as_obs_source(X)
        if src.ndim not in (2, 3):
            raise ValueError(f"X must be 2D or 3D; got shape {tuple(src.shape)}")
        N, L = int(src.shape[0]), int(src.shape[1])
        if coords is None:
            coords = np.arange(L, dtype=int)
        coords = np.asarray(coords, dtype=int)
        tables = self._host_tables()
        bins = self._bins_for(coords, L, device)
        K = self.n_states
        use_viterbi = str(decode).lower() == "viterbi"

        st_dtype = np.int8 if compact else np.int64
        g_dtype = np.float32 if compact else _TORCH_TO_NP[self.dtype]
        if out is not None:
            states_out, gamma_out = out
            if states_out.shape != (N, L) or gamma_out.shape != (N, L, K):
                raise ValueError("out buffers have the wrong shape")
            if states_out.dtype != st_dtype or gamma_out.dtype != g_dtype:
                raise ValueError(f"out buffers must be ({np.dtype(st_dtype)}, {np.dtype(g_dtype)})")
        else:
            bufs = _alloc_host_arrays({"states": ((N, L), st_dtype), "gamma": ((N, L, K), g_dtype)})
            states_out, gamma_out = bufs["states"], bufs["gamma"]
        if N == 0 or L == 0:
            return states_out, gamma_out
        from .stream import decode_stream

        decode_stream(self, src, numeric_u8, tables, bins, use_viterbi, states_out, gamma_out, device)
        return states_out, gamma_out

    # ---- EM fit (HMM.py:724-821, 1518-1630) ----------------------------------------------------
    def fit(
        self,
        X,
        coords: Optional[np.ndarray] = None,
        *,
        max_iter: int = 50,
        tol: float = 1e-5,
        device: Optional[Union[str, torch.device]] = None,
        update_start: bool = True,
        update_trans: bool = True,
        update_emission: bool = True,
        verbose: bool = False,
        **kwargs,
    ) -> List[float]:
        src, numeric_u8 = self._as_obs_source(X)
        if src.ndim not in (2, 3):
            raise ValueError(f"X must be 2D or 3D; got {tuple(src.shape)}")
        L = int(src.shape[1])
        if coords is None:
            coords = np.arange(L, dtype=int)
        coords = np.asarray(coords, dtype=int)
        device = self._ensure_device_dtype(device)
        if src.dtype in (np.uint8, torch.uint8) and not numeric_u8:
            src = CodedU8(src)  # keep the opt-in visible to fit_em
        return self.fit_em(
            src, coords, device=device, max_iter=max_iter, tol=tol, update_start=update_start,
            update_trans=update_trans, update_emission=update_emission, verbose=verbose, **kwargs,
        )

    def adapt_emissions(
        self,
        X,
        coords: Optional[np.ndarray] = None,
        *,
        device: Optional[Union[str, torch.device]] = None,
        iters: Optional[int] = None,
        max_iter: Optional[int] = None,
        verbose: bool = False,
        freeze_start: bool = True,
        freeze_trans: bool = True,
        **kwargs,
    ) -> List[float]:
        """Emission-only EM with ``tol=0`` (HMM.py:783-821 and the subclass aliases :1632-1665)."""
        if iters is None:
            iters = int(max_iter) if max_iter is not None else int(kwargs.pop("iters", 5))
        return self.fit(
            X, coords, max_iter=int(iters), tol=0.0, device=device, update_start=not freeze_start,
            update_trans=not freeze_trans, update_emission=True, verbose=verbose,
        )

    def _check_fit_input(self, src):
        want = self._expected_ndim()
        if want is not None and src.ndim != want:
            shape = "(N,L)" if want == 2 else "(N,L,C)"
            raise ValueError(f"{type(self).__name__} expects X shape {shape}.")

    def _prepare_fit(self, src, device):
        """Hook run once before the EM loop (multi: channel-count growth)."""

    def fit_em(
        self,
        X,
        coords: np.ndarray,
        *,
        device: torch.device,
        max_iter: int,
        tol: float,
        update_start: bool,
        update_trans: bool,
        update_emission: bool,
        verbose: bool,
        **kwargs,
    ) -> List[float]:
        """Baum-Welch: device E-step (sufficient statistics in fp64), host M-step.

        The observations are uploaded once and stay resident for all iterations.  With
        ``self.distributed`` (or an explicit ``self.process_group``) the statistics vector is
        sum-all-reduced over NCCL before the (replicated, deterministic) M-step, so every rank
        holds bit-identical parameters.
        """
        src, numeric_u8 = self._as_obs_source(X)
        self._check_fit_input(src)
        device = self._ensure_device_dtype(device)
        N, L = int(src.shape[0]), int(src.shape[1])
        if self._device_loop_eligible(N, L, int(max_iter), verbose):
            # small fit (the reference's regime: hmm_max_fit_reads = 1000): the whole Baum-Welch loop -- E-step,
            # M-step, stop rule -- runs on the device without a host round trip per iteration (grouped.py, one group)
            from .grouped import fit_groups

            return fit_groups([self], [X], max_iter=int(max_iter), tol=float(tol), update_start=update_start,
                              update_trans=update_trans, update_emission=update_emission, device=device)[0]
        self._prepare_fit(src, device)
        bins = self._bins_for(np.asarray(coords, dtype=int), L, device)
        resident = self._resident_chunks(src, device, numeric_u8)
        if int(max_iter) >= 2:
            resident = self._code_binary_resident(resident)
        hist: List[float] = []
        workspace = None
        # optional per-iteration timing split (bench.py): set ``model.fit_profile = []`` before calling fit
        prof = getattr(self, "fit_profile", None)
        marks = []
        for it in range(1, int(max_iter) + 1):
            tables = self._host_tables()
            if workspace is None:
                # sized for the largest resident chunk (includes the two-kernel path's boundary vectors)
                workspace = engine.estep_workspace(tables, device, getattr(resident, "max_rows", None) or max(
                    (int(o.shape[0]) for o in resident), default=1), L)
            if prof is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            total = self._estep_local(tables, resident, bins, workspace, device)
            if prof is not None:
                ev[1].record()
            total = self._allreduce_stats(total)
            if prof is not None:
                ev[2].record()
            stats = total.cpu()
            t_host = time.perf_counter()
            hist.append(float(stats[-1].item()))
            self._m_step(stats, update_start=update_start, update_trans=update_trans,
                         update_emission=update_emission)
            if prof is not None:
                marks.append((ev, time.perf_counter() - t_host))
            if verbose:
                print(f"[{type(self).__name__}.fit] iter={it} ll_proxy={hist[-1]:.6f}")
            if len(hist) > 1 and abs(hist[-1] - hist[-2]) < float(tol) * max(abs(hist[-1]), 1.0):
                break
        if prof is not None:
            for ev, host_s in marks:
                prof.append({"estep_ms": ev[0].elapsed_time(ev[1]), "allreduce_ms": ev[1].elapsed_time(ev[2]),
                             "host_mstep_ms": 1e3 * host_s})
        return hist

    # read-positions up to which ``fit`` keeps the EM loop on the device (beyond, one E-step takes long enough that
    # the ~0.25 ms host M-step per iteration does not matter and the chunk-parallel kernels are the faster E-step)
    device_loop_max_positions = 4_000_000
    # ... and the longest read: the device loop walks a read sequentially (one thread per read, ~80 ns per position
    # and iteration), which only beats the ~0.2 ms host round trip for amplicon-sized reads
    device_loop_max_length = 1024

    def _device_loop_eligible(self, N: int, L: int, max_iter: int, verbose: bool) -> bool:
        """Only the plain single-channel model has the device-side M-step (``SingleBernoulliHMM`` overrides)."""
        return False

    def _estep_local(self, tables, resident, bins, workspace, device) -> torch.Tensor:
        """This rank's sufficient statistics [start | trans | emit_num | emit_den | ll] (fp64, on
        the device) from ``smf_hmm_estep``."""
        total = None
        for obs in resident:
            st = engine.estep(tables, obs, bins, workspace=workspace)
            total = st if total is None else total + st
        if total is None:
            total = torch.zeros((engine.stats_len(tables),), dtype=torch.float64, device=device)
        return total

    def _resident_chunks(self, src, device, numeric_u8: bool = False):
        """Observations for the EM loop: device-resident read chunks (uploaded once, reused by every
        iteration), or -- when the input does not fit in the free device memory -- a generator that
        streams the chunks from the host in every iteration."""
        N = int(src.shape[0])
        row_bytes = int(np.prod(src.shape[1:])) * (src.element_size() if isinstance(src, torch.Tensor)
                                                   else src.dtype.itemsize)
        rows = max(1, int(self.device_chunk_bytes // max(row_bytes, 1)))
        bounds = [(lo, min(N, lo + rows)) for lo in range(0, N, rows)]
        if isinstance(src, torch.Tensor) and src.is_cuda:
            # already resident: row-block views (no copy) keep the per-launch workspace bounded
            src = src.contiguous()
            return [self._to_device(src[lo:hi], device, numeric_u8) for lo, hi in bounds]
        free, _total = torch.cuda.mem_get_info(device)
        if N * row_bytes * (4 if numeric_u8 else 1) <= 0.8 * free:
            return [self._to_device(src[lo:hi], device, numeric_u8) for lo, hi in bounds]

        class _Streamed:
            max_rows = max(hi - lo for lo, hi in bounds)

            def __iter__(inner):
                for lo, hi in bounds:
                    yield self._to_device(src[lo:hi], device, numeric_u8)

        return _Streamed()

    @staticmethod
    def _code_binary_resident(resident):
        """Binary calls that arrived as floats (0 / 1 / NaN -- the reference's host dtype) are re-coded once as the
        kernels' uint8 code (``smf_hmm_build_obs`` with the identity column map; its flag reports any other value):
        every EM iteration then reads 1 byte instead of 4-8 per position and takes the table-lookup emission path
        (K = 4, 262,144 x 8 kb: 13.4 -> 10.3 ms per iteration).  Probabilistic calls are left as they are."""
        if not isinstance(resident, list) or not resident:
            return resident
        if any((not isinstance(t, torch.Tensor)) or (not t.is_cuda) or t.dtype not in (torch.float32, torch.float64)
               or t.numel() == 0 for t in resident):
            return resident
        dev = resident[0].device
        L = int(resident[0].shape[1])
        C = int(resident[0].shape[2]) if resident[0].dim() == 3 else 1
        cols = torch.arange(L * C, dtype=torch.int32, device=dev).reshape(L, C)
        flag = torch.zeros((1,), dtype=torch.int32, device=dev)
        coded = []
        for t in resident:
            n = int(t.shape[0])
            out, _ = engine.build_obs(t.reshape(n, L * C), cols, C, out_u8=True, flag=flag)
            coded.append(out)
        if int(flag.item()) != 0:
            return resident  # some call is not 0 / 1 / NaN: keep the float observations
        return coded

    def _allreduce_stats(self, stats: torch.Tensor) -> torch.Tensor:
        if self.process_group is None and not self.distributed:
            return stats
        import torch.distributed as dist

        if not dist.is_available() or not dist.is_initialized():
            return stats
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.process_group)
        return stats

    def _split_stats(self, stats: torch.Tensor):
        K = self.n_states
        nb = int(self._trans_stack().shape[0])
        C = int(self._emission_matrix().shape[1])
        o = 0
        start_acc = stats[o:o + K]; o += K
        trans_acc = stats[o:o + nb * K * K].reshape(nb, K, K); o += nb * K * K
        emit_num = stats[o:o + K * C].reshape(K, C); o += K * C
        emit_den = stats[o:o + K * C].reshape(K, C); o += K * C
        return start_acc, trans_acc, emit_num, emit_den

    def _m_step(self, stats: torch.Tensor, *, update_start: bool, update_trans: bool, update_emission: bool):
        """Closed-form updates with the reference's eps smoothing, followed by its unconditional
        re-normalisation of every parameter (HMM.py:1598-1614)."""
        eps = float(self.eps)
        start_acc, trans_acc, emit_num, emit_den = self._split_stats(stats.to(self.dtype))
        with torch.no_grad():
            if update_start:
                s = start_acc + eps
                self.start.data = (s / s.sum()).to(self.start.device)
            if update_trans:
                self._assign_trans(trans_acc, eps)
            if update_emission:
                em = ((emit_num + eps) / (emit_den + 2.0 * eps)).clamp(min=eps, max=1.0 - eps)
                self._assign_emission(em)
        self._normalize_all()

    def _assign_trans(self, trans_acc: torch.Tensor, eps: float):
        t = trans_acc[0] + eps
        rows = t.sum(dim=1, keepdim=True)
        rows[rows == 0.0] = 1.0
        self.trans.data = (t / rows).to(self.trans.device)

    def _assign_emission(self, em: torch.Tensor):
        raise NotImplementedError

    def _normalize_all(self):
        self._normalize_params()

    # ---- save / load (HMM.py:858-923) ------------------------------------------------------------
    def _extra_save_payload(self) -> dict:
        return {}

    def _load_extra_payload(self, payload: dict, *, device: torch.device):
        return

    def save(self, path: Union[str, Path]) -> None:
        payload = {
            "hmm_name": getattr(self, "hmm_name", type(self).__name__),
            "class": type(self).__name__,
            "n_states": int(self.n_states),
            "eps": float(self.eps),
            "dtype": str(self.dtype),
            "start": self.start.detach().cpu(),
            "trans": self.trans.detach().cpu(),
        }
        payload.update(self._extra_save_payload())
        torch.save(payload, str(path))

    @classmethod
    def load(cls, path: Union[str, Path], device: Optional[Union[str, torch.device]] = None):
        payload = torch.load(str(path), map_location="cpu")
        klass = _HMM_REGISTRY.get(payload.get("hmm_name", None), cls)
        dtype = getattr(torch, str(payload.get("dtype", "torch.float64")).split(".")[-1], torch.float64)
        model = klass(n_states=int(payload["n_states"]), eps=float(payload.get("eps", 1e-8)), dtype=dtype)
        dev = torch.device(device) if isinstance(device, str) else (device or torch.device("cpu"))
        model.to(dev)
        with torch.no_grad():
            model.start.data = payload["start"].to(device=dev, dtype=model.dtype)
            model.trans.data = payload["trans"].to(device=dev, dtype=model.dtype)
        model._load_extra_payload(payload, device=dev)
        model._normalize_params()
        return model

    # ---- interval helpers kept for API parity (HMM.py:927-1002) ------------------------------------
    @staticmethod
    def _runs_from_bool(mask_1d: np.ndarray) -> List[Tuple[int, int]]:
        m = np.asarray(mask_1d).astype(bool)
        if m.size == 0:
            return []
        edges = np.diff(np.concatenate(([0], m.view(np.int8), [0])))
        return list(zip(np.nonzero(edges == 1)[0], np.nonzero(edges == -1)[0]))

    @staticmethod
    def _interval_length(coords: np.ndarray, s: int, e: int) -> int:
        return 0 if e <= s else int(coords[e - 1]) - int(coords[s]) + 1

    @staticmethod
    def _write_lengths_for_binary_layer(bin_mat: np.ndarray) -> np.ndarray:
        return _run_lengths_np(bin_mat)

    @staticmethod
    def _write_lengths_for_state_layer(states: np.ndarray) -> np.ndarray:
        st = np.asarray(states)
        out = np.zeros(st.shape, dtype=np.int32)
        for v in np.unique(st[st >= 0]):
            out += _run_lengths_np(st == v)
        return out

    @staticmethod
    def _register_appended_layers(adata, names: Sequence[str]) -> None:
        current = adata.uns.get("hmm_appended_layers")
        if current is None:
            registered = []
        elif isinstance(current, str):
            registered = [current]
        else:
            registered = list(current)
        for nm in names:
            if nm not in registered:
                registered.append(nm)
        adata.uns["hmm_appended_layers"] = registered

    # ---- merging (HMM.py:1006-1123) ----------------------------------------------------------------
    def merge_intervals_to_new_layer(
        self,
        adata,
        base_layer: str,
        *,
        distance_threshold: int,
        suffix: str = "_merged",
        overwrite: bool = True,
    ) -> str:
        """Merge neighbouring 1-runs whose coordinate gap is <= ``distance_threshold``; writes
        ``{base_layer}{suffix}`` (uint8) and ``..._lengths`` (int32, index units)."""
        if base_layer not in adata.layers:
            raise KeyError(f"Layer '{base_layer}' not found.")
        merged_name = f"{base_layer}{suffix}"
        len_name = f"{merged_name}_lengths"
        if (merged_name in adata.layers or len_name in adata.layers) and not overwrite:
            raise KeyError(f"Merged outputs exist (use overwrite=True): {merged_name}")
        coords, are_ints = _safe_int_coords(adata.var_names)
        member = (np.asarray(adata.layers[base_layer]) > 0).astype(np.uint8)
        dev = _default_cuda_device()
        merged, mlen, _, _ = engine.merge_runs(
            torch.from_numpy(np.ascontiguousarray(member)).to(dev),
            torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(dev),
            are_ints, int(distance_threshold),
        )
        adata.layers[merged_name] = merged.cpu().numpy()
        adata.layers[len_name] = mlen.cpu().numpy()
        self._register_appended_layers(adata, (merged_name, len_name))
        return merged_name

    def write_size_class_layers_from_binary(
        self,
        adata,
        base_layer: str,
        *,
        out_prefix: str,
        feature_ranges: Dict[str, Tuple[float, float]],
        suffix: str = "",
        overwrite: bool = True,
    ) -> List[str]:
        """Bin the runs of a binary layer into size classes by genomic length; writes
        ``{out_prefix}_{feature}{suffix}`` (uint8) and ``..._lengths`` (int32)."""
        if base_layer not in adata.layers:
            raise KeyError(f"Layer '{base_layer}' not found.")
        coords, are_ints = _safe_int_coords(adata.var_names)
        member = (np.asarray(adata.layers[base_layer]) > 0).astype(np.uint8)
        names = list(feature_ranges.keys())
        created: List[str] = []
        writable = []
        for feat in names:
            nm = f"{out_prefix}_{feat}{suffix}"
            ln = f"{nm}_lengths"
            if (nm in adata.layers or ln in adata.layers) and not overwrite:
                writable.append(False)
                continue
            writable.append(True)
            created.extend([nm, ln])
        dev = _default_cuda_device()
        # distance_threshold = -1: nothing is merged (a gap is never negative), classes are assigned
        # to the layer's own runs
        _, _, cls, clen = engine.merge_runs(
            torch.from_numpy(np.ascontiguousarray(member)).to(dev),
            torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64)).to(dev),
            are_ints, -(1 << 62), [feature_ranges[f] for f in names] or [(0.0, 0.0)],
        )
        cls_h = cls.cpu().numpy()
        clen_h = clen.cpu().numpy()
        for ci, feat in enumerate(names):
            nm = f"{out_prefix}_{feat}{suffix}"
            hit = cls_h == (ci + 1)
            if writable[ci]:
                adata.layers[nm] = hit.astype(np.uint8)
                adata.layers[f"{nm}_lengths"] = np.where(hit, clen_h, 0).astype(np.int32)
            else:
                # existing layers that may not be recreated still receive the runs (HMM.py:1103-1119)
                lay = np.asarray(adata.layers[nm])
                lay[hit] = 1
                adata.layers[nm] = lay
                adata.layers[f"{nm}_lengths"] = _run_lengths_np(lay)
        self._register_appended_layers(adata, created)
        return created

    # ---- AnnData annotation (HMM.py:1169-1381) --------------------------------------------------------
    def annotate_adata(
        self,
        adata,
        *,
        prefix: str,
        X,
        coords: np.ndarray,
        var_mask: np.ndarray,
        span_fill: bool = True,
        config=None,
        decode: str = "marginal",
        write_posterior: bool = True,
        posterior_state: str = "Modified",
        feature_sets: Optional[Dict[str, Dict[str, Any]]] = None,
        prob_threshold: float = 0.5,
        uns_key: str = "hmm_appended_layers",
        uns_flag: str = "hmm_annotated",
        force_redo: bool = False,
        mask_to_read_span: bool = True,
        mask_use_original_var_names: bool = True,
        device: Optional[Union[str, torch.device]] = None,
        **kwargs,
    ):
        """Decode ``X`` and write the reference's layer set into ``adata`` (in place).

        Device work: posterior (+ Viterbi), per-group run calling / size classes / span-fill /
        run lengths, read-span validity.  Host work: placing the packed device results into
        (n_obs, n_vars) numpy layers with the reference's names, order and dtypes.
        """
        if bool(adata.uns.get(uns_flag, False)) and not force_redo:
            return None
        if adata.uns.get(uns_key) is None:
            adata.uns[uns_key] = []
        appended = list(adata.uns.get(uns_key, []))

        src, numeric_u8 = self._as_obs_source(X)
        coords = np.asarray(coords, dtype=int)
        var_mask = np.asarray(var_mask, dtype=bool)
        if var_mask.shape[0] != adata.n_vars:
            raise ValueError(f"var_mask length {var_mask.shape[0]} != adata.n_vars {adata.n_vars}")
        device = self._ensure_device_dtype(device)
        if src.ndim not in (2, 3):
            raise ValueError(f"X must be 2D or 3D; got shape {tuple(src.shape)}")
        N, L = int(src.shape[0]), int(src.shape[1])
        if N != adata.n_obs:
            raise ValueError(f"X has N={N} rows but adata.n_obs={adata.n_obs}")
        n_obs, n_vars = adata.n_obs, adata.n_vars
        use_viterbi = str(decode).lower() == "viterbi"

        full_coords, full_int = _safe_int_coords(adata.var_names)
        masked_idx = np.nonzero(var_mask)[0]
        masked_coords, _ = _safe_int_coords(adata.var_names[var_mask])
        pos_of_coord = {int(c): i for i, c in enumerate(coords.tolist())}
        take = np.array([pos_of_coord.get(int(c), -1) for c in masked_coords.tolist()], dtype=int)
        good = take >= 0
        masked_idx = masked_idx[good]
        take = take[good]
        masked_coords_good = masked_coords[good]

        def note(name):
            if name not in appended:
                appended.append(name)

        # ---- allocate / register the output layers in the reference's order (HMM.py:1256-1315) ----
        states_name = f"{prefix}_states"
        fresh_base = states_name not in adata.layers
        if states_name not in adata.layers:
            adata.layers[states_name] = np.full((n_obs, n_vars), -1, dtype=np.int8)
        note(states_name)
        post_name = None
        t_post = None
        if write_posterior:
            t_post = self.resolve_target_state_index(posterior_state)
            tag = str(posterior_state).strip().lower().replace(" ", "_").replace("-", "_")
            post_name = f"{prefix}_posterior_{tag}"
            if post_name not in adata.layers:
                adata.layers[post_name] = np.zeros((n_obs, n_vars), dtype=np.float32)
            else:
                fresh_base = False
            note(post_name)

        if feature_sets is None:
            feature_sets = normalize_hmm_feature_sets(self._cfg_to_dict(config).get("hmm_feature_sets", None))
        groups = []
        for group, fs in (feature_sets or {}).items():
            fmap = fs.get("features", {}) or {}
            if not fmap:
                continue
            all_layer = f"{prefix}_all_{group}_features"
            pre_existing = set()
            for nm, dt in ((all_layer, np.uint8), (f"{all_layer}_lengths", np.int32)):
                if nm in adata.layers:
                    pre_existing.add(nm)
                else:
                    adata.layers[nm] = np.zeros((n_obs, n_vars), dtype=dt)
                note(nm)
            for feat in fmap.keys():
                nm = f"{prefix}_{feat}"
                if nm in adata.layers:
                    pre_existing.add(nm)
                else:
                    adata.layers[nm] = np.zeros((n_obs, n_vars),
                                                dtype=np.int32 if nm.endswith("_lengths") else np.uint8)
                if f"{nm}_lengths" not in adata.layers:
                    adata.layers[f"{nm}_lengths"] = np.zeros((n_obs, n_vars), dtype=np.int32)
                note(nm)
                note(f"{nm}_lengths")
            feats = list(fmap.keys())
            groups.append(dict(name=group, all_layer=all_layer, feats=feats, ranges=[fmap[f] for f in feats],
                               target=self.resolve_target_state_index(fs.get("state", "Modified")),
                               clean=not pre_existing))

        # can the device span-fill kernel be used?  (sites strictly increasing and on the grid)
        kernel_ok = bool(groups and span_fill and full_int and L > 0)
        if kernel_ok:
            left = np.searchsorted(full_coords, coords, side="left")
            right = np.searchsorted(full_coords, coords, side="right")
            kernel_ok = bool(
                np.all(np.diff(coords) > 0)
                and np.all(np.diff(full_coords) >= 0)
                and np.all(left < n_vars)
                and np.all(full_coords[np.minimum(left, n_vars - 1)] == coords)
            )
        if kernel_ok:
            coords_d = torch.from_numpy(np.array(coords, dtype=np.int64, copy=True)).to(device)
            left_d = torch.from_numpy(np.ascontiguousarray(left, dtype=np.int32)).to(device)
            right_d = torch.from_numpy(np.ascontiguousarray(right, dtype=np.int32)).to(device)

        # Final-form fast path: when every layer of this call is new and the read-span mask follows anyway,
        # the layers are expanded to their final dtype (float64 / float32 posterior) and NaN-masked ON THE
        # DEVICE, chunk by chunk, and land in host memory once -- instead of uint8 / int32 host layers that
        # mask_layers_outside_read_span re-reads, converts and masks on one CPU core (HMM.py:1372-1377).
        fast_final = bool(kernel_ok and mask_to_read_span and feature_sets and fresh_base
                          and all(gr["clean"] for gr in groups) and len(groups) <= 8)
        layer_plan = []  # (name, kind, group index, class index, is_float32)
        if fast_final:
            if "reference_start" not in adata.obs or "reference_end" not in adata.obs:
                raise KeyError("Missing 'reference_start' or 'reference_end' in adata.obs.")
            span_in = _span_inputs(adata, start_key="reference_start", end_key="reference_end",
                                   use_original_var_names=mask_use_original_var_names, device=device)
            layer_plan.append((states_name, _lib.LAYER_STATES, 0, 0, False))
            if post_name is not None:
                layer_plan.append((post_name, _lib.LAYER_POSTERIOR, 0, 0, True))
            for gi, gr in enumerate(groups):
                layer_plan.append((gr["all_layer"], _lib.LAYER_ANY, gi, 0, False))
                layer_plan.append((f"{gr['all_layer']}_lengths", _lib.LAYER_ANY_LEN, gi, 0, False))
                for ci, feat in enumerate(gr["feats"]):
                    layer_plan.append((f"{prefix}_{feat}", _lib.LAYER_CLASS, gi, ci, False))
                    layer_plan.append((f"{prefix}_{feat}_lengths", _lib.LAYER_CLASS_LEN, gi, ci, False))
            fast_final = len(layer_plan) <= 64
        if fast_final:
            fresh_names = [p[0] for p in layer_plan]
            for nm, arr in _alloc_host_layers(
                    (n_obs, n_vars), {p[0]: (np.float32 if p[4] else np.float64) for p in layer_plan}).items():
                adata.layers[nm] = arr
            site_of_col = np.full((n_vars,), -1, dtype=np.int32)
            site_of_col[masked_idx] = take.astype(np.int32)
            site_of_col_d = torch.from_numpy(site_of_col).to(device)

        # ---- decode + feature calling on the device, one chunk of reads at a time ----
        from .stream import ChunkUploader

        tables = self._host_tables()
        bins = self._bins_for(coords, L, device)
        K = self.n_states
        C = int(src.shape[2]) if src.ndim == 3 else 1
        item = src.element_size() if isinstance(src, torch.Tensor) else src.dtype.itemsize
        if numeric_u8:
            item += 4
        n_f64 = sum(1 for p in layer_plan if not p[4])
        n_f32 = sum(1 for p in layer_plan if p[4])
        stage_row = n_vars * (8 * n_f64 + 4 * n_f32)
        per_read = L * (2 * C * item + 4 * K + 2) + n_vars * (5 * max(1, len(groups)) + 2) + 2 * stage_row + 64
        rows = max(1, int(self.device_chunk_bytes // per_read))
        if fast_final:
            rows = max(1, min(rows, int((384 << 20) // max(stage_row, 1))))
        rows = min(rows, max(N, 1))
        cur = torch.cuda.current_stream(device)
        uploader = ChunkUploader(src, rows, device)
        entry = torch.cuda.Event()
        entry.record(cur)
        uploader.wait_entry(entry)
        ev_cmp = [None, None]
        ev_out = [None, None]
        if fast_final:
            s_out = _stream_mod.side_stream(device, "d2h")
            s_out.wait_event(entry)
            stage64 = [torch.empty((n_f64, rows, n_vars), dtype=torch.float64, device=device) for _ in range(2)]
            stage32 = [torch.empty((max(n_f32, 1), rows, n_vars), dtype=torch.float32, device=device) for _ in range(2)]
        try:
            for ck, lo in enumerate(range(0, N, rows)):
                hi = min(N, lo + rows)
                b = ck % 2
                obs = uploader.upload(ck, lo, hi, ev_cmp[b])
                if numeric_u8 and obs.dtype == torch.uint8:
                    obs = obs.to(torch.float32)
                gamma_d, states_d, _ = engine.posterior(tables, obs, bins, want_gamma=True,
                                                        want_states=not use_viterbi)
                if use_viterbi:
                    states_d = engine.viterbi(tables, obs, bins)
                del obs
                if not fast_final:
                    states_h = states_d.cpu().numpy()
                    adata.layers[states_name][lo:hi, masked_idx] = states_h[:, take]
                    if post_name is not None:
                        post_h = gamma_d[:, :, t_post].contiguous().cpu().numpy()
                        adata.layers[post_name][lo:hi, masked_idx] = post_h[:, take]
                packed = []
                for gr in groups:
                    all_layer, feats, ranges, target_idx = gr["all_layer"], gr["feats"], gr["ranges"], gr["target"]
                    if kernel_ok:
                        if use_viterbi:
                            cls_d, len_d, _ = engine.call_features(
                                states=states_d, post=None, target=target_idx, threshold=0.0, coords=coords_d,
                                grid_left=left_d, grid_right=right_d, n_grid=n_vars, ranges=ranges)
                        else:
                            post_d = gamma_d[:, :, target_idx].contiguous()
                            cls_d, len_d, _ = engine.call_features(
                                states=None, post=post_d, target=target_idx, threshold=float(prob_threshold),
                                coords=coords_d, grid_left=left_d, grid_right=right_d, n_grid=n_vars, ranges=ranges)
                        if fast_final:
                            packed.append((cls_d, len_d, False))
                            continue
                        cls_h = cls_d.cpu().numpy()
                        len_h = len_d.cpu().numpy()
                        any_hit = cls_h > 0
                        if gr["clean"]:
                            adata.layers[all_layer][lo:hi] = any_hit.astype(np.uint8)
                            adata.layers[f"{all_layer}_lengths"][lo:hi] = np.where(any_hit, len_h, 0).astype(np.int32)
                            for ci, feat in enumerate(feats):
                                hit = cls_h == (ci + 1)
                                adata.layers[f"{prefix}_{feat}"][lo:hi] = hit.astype(np.uint8)
                                adata.layers[f"{prefix}_{feat}_lengths"][lo:hi] = np.where(hit, len_h, 0).astype(np.int32)
                        else:
                            # layers carried over from an earlier annotation: OR the new runs in and
                            # recompute run lengths on the union (HMM.py:1348-1370)
                            lay = np.asarray(adata.layers[all_layer])
                            lay[lo:hi][any_hit] = 1
                            adata.layers[all_layer] = lay
                            adata.layers[f"{all_layer}_lengths"][lo:hi] = _run_lengths_np(lay[lo:hi])
                            for ci, feat in enumerate(feats):
                                lay = np.asarray(adata.layers[f"{prefix}_{feat}"])
                                lay[lo:hi][cls_h == (ci + 1)] = 1
                                adata.layers[f"{prefix}_{feat}"] = lay
                                adata.layers[f"{prefix}_{feat}_lengths"][lo:hi] = _run_lengths_np(lay[lo:hi])
                    else:
                        membership = (
                            (states_h == target_idx) if use_viterbi
                            else (gamma_d[:, :, target_idx].contiguous().cpu().numpy().astype(np.float64)
                                  >= float(prob_threshold))
                        )
                        self._fill_features_host(adata, prefix, gr["name"], feats, ranges, membership, coords,
                                                 full_coords, bool(span_fill and full_int), masked_idx,
                                                 masked_coords_good, lo)
                if fast_final:
                    # ONE kernel writes every final layer of the chunk (final dtype, NaN outside the read span)
                    # into the device staging slot; the copy engine moves each layer block to its host array
                    n_chunk = hi - lo
                    if "covered" in span_in:
                        cov_d = torch.from_numpy(np.ascontiguousarray(span_in["covered"][lo:hi])).to(device)
                        valid_u8 = engine.span_mask(n_chunk, n_vars, device, covered=cov_d)
                    else:
                        valid_u8 = engine.span_mask(n_chunk, n_vars, device, coords=span_in["coords"],
                                                    start=span_in["start"][lo:hi].contiguous(),
                                                    end=span_in["end"][lo:hi].contiguous())
                    if ev_out[b] is not None:
                        cur.wait_event(ev_out[b])
                    outs = []
                    i64 = i32 = 0
                    for (nm, kind, gi, ci, is32) in layer_plan:
                        if is32:
                            outs.append(stage32[b][i32, :n_chunk])
                            i32 += 1
                        else:
                            outs.append(stage64[b][i64, :n_chunk])
                            i64 += 1
                    engine.expand_layers(
                        n_chunk, n_vars, L, states=states_d,
                        post=gamma_d[:, :, t_post] if post_name is not None else None, post_stride=K,
                        site_of_col=site_of_col_d, valid=valid_u8, groups=packed,
                        layers=[(kind, gi, ci, out) for (nm, kind, gi, ci, is32), out in zip(layer_plan, outs)])
                    ev_cmp[b] = torch.cuda.Event()
                    ev_cmp[b].record(cur)
                    s_out.wait_event(ev_cmp[b])
                    for (nm, *_rest), out in zip(layer_plan, outs):
                        engine.copy_to_host_async(adata.layers[nm][lo:hi], out, s_out)
                    ev_out[b] = torch.cuda.Event()
                    ev_out[b].record(s_out)
                    del packed, valid_u8
                else:
                    ev_cmp[b] = torch.cuda.Event()
                    ev_cmp[b].record(cur)
                del gamma_d, states_d
        finally:
            uploader.finish()
            if fast_final:
                s_out.synchronize()
            cur.synchronize()

        if not feature_sets:
            # no feature sets: only states / posterior are written and, like the reference, no
            # read-span masking is applied (HMM.py:1280-1283)
            adata.uns[uns_key] = appended
            adata.uns[uns_flag] = True
            return None

        if mask_to_read_span and appended:
            if fast_final:
                # this call's layers are final already; layers appended by earlier calls get the (idempotent) pass
                others = [nm for nm in appended if nm not in set(fresh_names)]
                if others:
                    mask_layers_outside_read_span(adata, others, use_original_var_names=mask_use_original_var_names)
            else:
                mask_layers_outside_read_span(adata, appended, use_original_var_names=mask_use_original_var_names)

        adata.uns[uns_key] = appended
        adata.uns[uns_flag] = True
        return None

    @staticmethod
    def _fill_features_host(adata, prefix, group, feats, ranges, membership, coords, full_coords, span_ok,
                            masked_idx, masked_coords, row0=0):
        """Generic (slow) span-fill used when the site grid does not satisfy the kernel's
        preconditions (unsorted / off-grid coordinates, ``span_fill=False``, non-integer var names)."""
        all_layer = f"{prefix}_all_{group}_features"
        lo = np.array([float(r[0]) for r in ranges])
        hi = np.array([float(r[1]) for r in ranges])
        for r in range(membership.shape[0]):
            i = row0 + r
            for s, e in BaseHMM._runs_from_bool(membership[r]):
                glen = int(coords[e - 1]) - int(coords[s]) + 1 if e > s else 0
                if glen <= 0:
                    continue
                ok = np.nonzero((lo <= float(glen)) & (float(glen) < hi))[0]
                if ok.size == 0:
                    continue
                chosen = feats[int(ok[0])]
                if span_ok:
                    a = int(np.searchsorted(full_coords, int(coords[s]), side="left"))
                    b = int(np.searchsorted(full_coords, int(coords[e - 1]), side="right"))
                    if a >= b:
                        continue
                    adata.layers[f"{prefix}_{chosen}"][i, a:b] = 1
                    adata.layers[all_layer][i, a:b] = 1
                else:
                    cols = masked_idx[(masked_coords >= coords[s]) & (masked_coords <= coords[e - 1])]
                    if cols.size == 0:
                        continue
                    adata.layers[f"{prefix}_{chosen}"][i, cols] = 1
                    adata.layers[all_layer][i, cols] = 1
        # run lengths of the rows this chunk touched only (reads are independent)
        r0, r1 = row0, row0 + membership.shape[0]
        adata.layers[f"{all_layer}_lengths"][r0:r1] = _run_lengths_np(np.asarray(adata.layers[all_layer])[r0:r1])
        for feat in feats:
            nm = f"{prefix}_{feat}"
            adata.layers[f"{nm}_lengths"][r0:r1] = _run_lengths_np(np.asarray(adata.layers[nm])[r0:r1])

    def _ensure_final_layer_and_assign(self, final_adata, layer_name: str, subset_idx_mask: np.ndarray, sub_data):
        """Row-copy helper used by the legacy workflow (HMM.py:1385-1403)."""
        rows = np.nonzero(np.asarray(subset_idx_mask).astype(bool))[0]
        sub = np.asarray(sub_data)
        if layer_name not in final_adata.layers:
            final_adata.layers[layer_name] = np.zeros(tuple(final_adata.shape), dtype=sub.dtype)
        if sub.shape[0] != rows.size:
            raise ValueError(f"Sub rows {sub.shape[0]} != mask sum {rows.size}")
        dst = np.asarray(final_adata.layers[layer_name])
        dst[rows, :] = sub
        final_adata.layers[layer_name] = dst


_TORCH_TO_NP = {torch.float64: np.float64, torch.float32: np.float32, torch.float16: np.float16}


# ----------------------------------------------------------------------------------------
# single-channel Bernoulli  (HMM.py:1411-1665)
# ----------------------------------------------------------------------------------------
@register_hmm("single")
class SingleBernoulliHMM(BaseHMM):
    """``emission[k] = P(obs == 1 | state k)`` over one methylation channel."""

    def __init__(self, n_states: int = 2, init_emission: Optional[Sequence[float]] = None, eps: float = 1e-8,
                 dtype: torch.dtype = torch.float64):
        super().__init__(n_states=n_states, eps=eps, dtype=dtype)
        K = self.n_states
        em = None
        if init_emission is not None:
            flat = np.asarray(init_emission, dtype=float).reshape(-1)[:K]
            if flat.size == K:
                em = flat
        if em is None:
            em = np.full((K,), 0.5, dtype=float)
        self.emission = nn.Parameter(torch.tensor(em, dtype=self.dtype), requires_grad=False)
        self._normalize_emission()

    @classmethod
    def from_config(cls, cfg, *, override=None, device=None):
        merged = cls._merged_cfg(cfg, override)
        model = cls(
            n_states=int(merged.get("hmm_n_states", 2)),
            init_emission=merged.get("hmm_init_emission_probs", merged.get("hmm_init_emission", None)),
            eps=float(merged.get("hmm_eps", 1e-8)),
            dtype=_resolve_dtype(merged.get("hmm_dtype", None)),
        )
        return cls._finish_from_config(model, merged, device)

    def _normalize_emission(self):
        with torch.no_grad():
            e = self.emission.data.reshape(-1)
            if e.numel() != self.n_states:
                e = torch.full((self.n_states,), 0.5, dtype=self.dtype, device=self.emission.device)
            self.emission.data = e.clamp(min=self.eps, max=1.0 - self.eps)

    def _param_list(self):
        return super()._param_list() + [self.emission]

    def _state_modified_score(self) -> torch.Tensor:
        return self.emission.detach()

    def _emission_matrix(self) -> torch.Tensor:
        return self.emission.detach().to("cpu", self.dtype).reshape(self.n_states, 1)

    def _expected_ndim(self):
        return 2

    def _device_loop_eligible(self, N: int, L: int, max_iter: int, verbose: bool) -> bool:
        return (type(self) is SingleBernoulliHMM and not verbose and max_iter >= 1 and N >= 1 and L >= 1
                and N * L <= self.device_loop_max_positions and L <= self.device_loop_max_length and not self.distributed and self.process_group is None
                and getattr(self, "fit_profile", None) is None and not getattr(self, "_host_loop_only", False))

    def _assign_emission(self, em: torch.Tensor):
        self.emission.data = em.reshape(-1).to(self.emission.device)

    def _normalize_all(self):
        self._normalize_params()
        self._normalize_emission()

    def _extra_save_payload(self) -> dict:
        return {"emission": self.emission.detach().cpu()}

    def _load_extra_payload(self, payload: dict, *, device: torch.device):
        with torch.no_grad():
            self.emission.data = payload["emission"].to(device=device, dtype=self.dtype)
        self._normalize_emission()


# ----------------------------------------------------------------------------------------
# multi-channel Bernoulli on a union grid  (HMM.py:1673-1976)
# ----------------------------------------------------------------------------------------
@register_hmm("multi")
class MultiBernoulliHMM(BaseHMM):
    """Independent Bernoulli channels: ``emission[k, c] = P(obs_c == 1 | state k)``."""

    def __init__(self, n_states: int = 2, n_channels: int = 2, init_emission: Optional[Any] = None,
                 eps: float = 1e-8, dtype: torch.dtype = torch.float64):
        super().__init__(n_states=n_states, eps=eps, dtype=dtype)
        self.n_channels = int(n_channels)
        if self.n_channels < 1:
            raise ValueError("n_channels must be >=1")
        K, C = self.n_states, self.n_channels
        em = None
        if init_emission is not None:
            arr = np.asarray(init_emission, dtype=float)
            if arr.ndim == 1:
                cand = np.repeat(arr.reshape(-1, 1)[:K, :], C, axis=1)
            else:
                cand = arr[:K, :C]
            if cand.shape == (K, C):
                em = cand
        if em is None:
            em = np.full((K, C), 0.5, dtype=float)
        self.emission = nn.Parameter(torch.tensor(em, dtype=self.dtype), requires_grad=False)
        self._normalize_emission()

    @classmethod
    def from_config(cls, cfg, *, override=None, device=None):
        merged = cls._merged_cfg(cfg, override)
        model = cls(
            n_states=int(merged.get("hmm_n_states", 2)),
            n_channels=int(merged.get("hmm_n_channels", merged.get("n_channels", 2))),
            init_emission=merged.get("hmm_init_emission_probs", None),
            eps=float(merged.get("hmm_eps", 1e-8)),
            dtype=_resolve_dtype(merged.get("hmm_dtype", None)),
        )
        return cls._finish_from_config(model, merged, device)

    def _normalize_emission(self):
        with torch.no_grad():
            e = self.emission.data.reshape(self.n_states, self.n_channels)
            self.emission.data = e.clamp(min=self.eps, max=1.0 - self.eps)

    def _param_list(self):
        return super()._param_list() + [self.emission]

    def _state_modified_score(self) -> torch.Tensor:
        return self.emission.detach().mean(dim=1)

    def _emission_matrix(self) -> torch.Tensor:
        return self.emission.detach().to("cpu", self.dtype).reshape(self.n_states, self.n_channels)

    def _expected_ndim(self):
        return 3

    def _assign_emission(self, em: torch.Tensor):
        self.emission.data = em.reshape(self.n_states, self.n_channels).to(self.emission.device)

    def _normalize_all(self):
        self._normalize_params()
        self._normalize_emission()

    def _prepare_fit(self, src, device):
        self._ensure_n_channels(int(src.shape[2]), device)

    def _ensure_n_channels(self, C: int, device):
        """Grow / shrink the emission table when the data has a different channel count: kept
        columns are copied, new columns take the per-state mean (HMM.py:1953-1976)."""
        C = int(C)
        if C == self.n_channels:
            return
        if C > _lib.MAX_CHANNELS:
            raise ValueError(f"smftools_b200 supports at most {_lib.MAX_CHANNELS} channels")
        with torch.no_grad():
            old = self.emission.detach().cpu().numpy()
            grown = np.full((old.shape[0], C), 0.5, dtype=float)
            keep = min(old.shape[1], C)
            grown[:, :keep] = old[:, :keep]
            if C > old.shape[1]:
                grown[:, keep:] = old.mean(axis=1, keepdims=True)
            self.n_channels = C
            self.emission = nn.Parameter(
                torch.tensor(grown, dtype=self.dtype, device=self.emission.device), requires_grad=False)
            self._normalize_emission()

    def _extra_save_payload(self) -> dict:
        return {"n_channels": int(self.n_channels), "emission": self.emission.detach().cpu()}

    def _load_extra_payload(self, payload: dict, *, device: torch.device):
        self.n_channels = int(payload.get("n_channels", self.n_channels))
        with torch.no_grad():
            self.emission.data = payload["emission"].to(device=device, dtype=self.dtype)
        self._normalize_emission()


# ----------------------------------------------------------------------------------------
# distance-binned transitions  (HMM.py:1984-2390)
# ----------------------------------------------------------------------------------------
@register_hmm("single_distance_binned")
class DistanceBinnedSingleBernoulliHMM(SingleBernoulliHMM):
    """One transition matrix per bin of the bp distance between consecutive sites."""

    def __init__(self, n_states: int = 2, init_emission: Optional[Sequence[float]] = None,
                 distance_bins: Optional[Sequence[int]] = None, init_trans_by_bin: Optional[Any] = None,
                 eps: float = 1e-8, dtype: torch.dtype = torch.float64):
        super().__init__(n_states=n_states, init_emission=init_emission, eps=eps, dtype=dtype)
        edges = [1, 5, 10, 25, 50, 100] if distance_bins is None else distance_bins
        self.distance_bins = np.asarray(edges, dtype=int)
        self.n_bins = int(len(self.distance_bins) + 1)
        if self.n_bins > _lib.MAX_BINS:
            raise ValueError(f"smftools_b200 supports at most {_lib.MAX_BINS} distance bins")
        K = self.n_states
        tb = None
        if init_trans_by_bin is not None:
            cand = np.asarray(init_trans_by_bin, dtype=float)
            if cand.shape == (self.n_bins, K, K):
                tb = cand
        if tb is None:
            tb = np.repeat(self.trans.detach().cpu().numpy()[None, :, :], self.n_bins, axis=0)
        self.trans_by_bin = nn.Parameter(torch.tensor(tb, dtype=self.dtype), requires_grad=False)
        self._normalize_trans_by_bin()

    @classmethod
    def from_config(cls, cfg, *, override=None, device=None):
        merged = cls._merged_cfg(cfg, override)
        model = cls(
            n_states=int(merged.get("hmm_n_states", 2)),
            init_emission=merged.get("hmm_init_emission_probs", None),
            distance_bins=merged.get("hmm_distance_bins", [1, 5, 10, 25, 50, 100]),
            init_trans_by_bin=merged.get("hmm_init_transitions_by_bin", None),
            eps=float(merged.get("hmm_eps", 1e-8)),
            dtype=_resolve_dtype(merged.get("hmm_dtype", None)),
        )
        return cls._finish_from_config(model, merged, device)

    def _normalize_trans_by_bin(self):
        with torch.no_grad():
            tb = self.trans_by_bin.data.reshape(self.n_bins, self.n_states, self.n_states) + self.eps
            rows = tb.sum(dim=2, keepdim=True)
            rows[rows == 0.0] = 1.0
            self.trans_by_bin.data = tb / rows

    def _param_list(self):
        return super()._param_list() + [self.trans_by_bin]

    def _trans_stack(self) -> torch.Tensor:
        return self.trans_by_bin.detach().to("cpu", self.dtype).reshape(self.n_bins, self.n_states, self.n_states)

    def _bin_index(self, coords: np.ndarray) -> np.ndarray:
        """Bin of every step from its bp distance (``np.digitize(..., right=True)``)."""
        return np.digitize(np.diff(np.asarray(coords, dtype=int)), self.distance_bins, right=True)

    def _bins_for(self, coords, L, device):
        if coords is None:
            raise ValueError("Distance-binned HMM requires coords.")
        if L <= 1:
            return None
        if len(coords) != L:
            raise ValueError(f"coords has length {len(coords)} but X has {L} positions")
        b = self._bin_index(coords).astype(np.int8)
        return torch.from_numpy(np.ascontiguousarray(b)).to(device)

    def _assign_trans(self, trans_acc: torch.Tensor, eps: float):
        tb = trans_acc + eps
        rows = tb.sum(dim=2, keepdim=True)
        rows[rows == 0.0] = 1.0
        self.trans_by_bin.data = (tb / rows).to(self.trans_by_bin.device)

    def _normalize_all(self):
        self._normalize_params()
        self._normalize_emission()
        self._normalize_trans_by_bin()

    def _extra_save_payload(self) -> dict:
        payload = super()._extra_save_payload()
        payload["distance_bins"] = torch.tensor(self.distance_bins, dtype=torch.long)
        payload["trans_by_bin"] = self.trans_by_bin.detach().cpu()
        return payload

    def _load_extra_payload(self, payload: dict, *, device: torch.device):
        super()._load_extra_payload(payload, device=device)
        bins = payload.get("distance_bins", torch.tensor([1, 5, 10, 25, 50, 100]))
        self.distance_bins = bins.cpu().numpy().astype(int)
        self.n_bins = int(len(self.distance_bins) + 1)
        with torch.no_grad():
            self.trans_by_bin.data = payload["trans_by_bin"].to(device=device, dtype=self.dtype)
        self._normalize_trans_by_bin()


# ----------------------------------------------------------------------------------------
# facade  (HMM.py:2398-2432)
# ----------------------------------------------------------------------------------------
class HMM:
    """``HMM.from_config(cfg, arch=...)`` / ``HMM.load(path)`` as the workflow imports it."""

    @staticmethod
    def from_config(cfg, arch: Optional[str] = None, **kwargs) -> BaseHMM:
        return create_hmm(cfg, arch=arch, **kwargs)

    @staticmethod
    def load(path: Union[str, Path], device: Optional[Union[str, torch.device]] = None) -> BaseHMM:
        return BaseHMM.load(path, device=device)
