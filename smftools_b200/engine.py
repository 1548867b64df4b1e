"""Device-side entry points: torch CUDA tensors in, torch CUDA tensors out, through the C-ABI.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every arithmetic
step runs in ``libsmfhmm.so``.  All functions raise if no CUDA device / library is present.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_OBS_DTYPES = {torch.float32: _lib.OBS_F32, torch.float64: _lib.OBS_F64, torch.uint8: _lib.OBS_U8}


class IntervalOverflow(ValueError):
    """``call_features`` found more runs than ``max_intervals``; ``needed`` is the size to retry with."""

    def __init__(self, needed: int, capacity: int):
        super().__init__(f"smf_hmm_call_features found {needed} intervals but the record buffer holds {capacity}; "
                         f"call again with max_intervals >= {needed}")
        self.needed = int(needed)
        self.capacity = int(capacity)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: smftools_b200 runs only on a CUDA device (no CPU fallback)"
        )


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else int(t.data_ptr())


def _obs_args(obs: torch.Tensor, tables: _lib.HostTables) -> Tuple[int, int, int]:
    _require_cuda(obs, "obs")
    if obs.dtype not in _OBS_DTYPES:
        raise ValueError(f"unsupported observation dtype {obs.dtype}; use float32, float64 or uint8")
    if obs.dim() == 2:
        N, L = obs.shape
        C = 1
    elif obs.dim() == 3:
        N, L, C = obs.shape
    else:
        raise ValueError(f"X must be 2D or 3D; got shape {tuple(obs.shape)}")
    if C != tables.C:
        raise ValueError(f"observations have {C} channels but the model has {tables.C}")
    if not obs.is_contiguous():
        raise ValueError("obs must be contiguous")
    return int(N), int(L), _OBS_DTYPES[obs.dtype]


def _bins_arg(bins: Optional[torch.Tensor], tables: _lib.HostTables, L: int) -> Optional[int]:
    if tables.n_bins <= 1:
        return None
    if L <= 1:
        return None
    if bins is None:
        raise ValueError("Distance-binned HMM requires coords.")
    _require_cuda(bins, "bins")
    if bins.dtype != torch.int8 or bins.numel() != L - 1 or not bins.is_contiguous():
        raise ValueError("bins must be a contiguous int8 tensor of length L-1")
    return int(bins.data_ptr())


def _narrow_f64(obs: torch.Tensor, tables: _lib.HostTables) -> torch.Tensor:
    """Multi-channel / distance-binned models: float64 observations are narrowed to float32 on the device first.
    Every forward-backward kernel converts an observation to float32 when it reads it (the posterior tolerance is an
    fp32 one), so the results are the same -- but only the float32 / uint8 rows have the fast-path instantiations of
    these two model families (a float64 row would run on the generic kernel, 3-4x slower)."""
    if obs.dtype == torch.float64 and (tables.C > 1 or tables.n_bins > 1):
        return obs.to(torch.float32)
    return obs


def set_option(key: str, value: int) -> None:
    """Process-wide tuning knob of the library (``smf_hmm_set_option``), e.g. ``set_option("walk", 1)``."""
    _lib.check(_lib.load().smf_hmm_set_option(key.encode(), int(value)), "smf_hmm_set_option")


def max_positions(tables: _lib.HostTables) -> int:
    return int(_lib.load().smf_hmm_max_positions(tables.ref, 1))


def posterior(
    tables: _lib.HostTables,
    obs: torch.Tensor,
    bins: Optional[torch.Tensor] = None,
    *,
    want_gamma: bool = True,
    gamma_state: int = -1,
    want_states: bool = True,
    want_loglik: bool = False,
    out_gamma: Optional[torch.Tensor] = None,
    out_states: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    use_workspace: bool = True,
):
    """Posterior decode (``smf_hmm_posterior``): returns (gamma, states, loglik), any may be None.

    ``workspace`` (uint8 device tensor) is only used by the two-kernel path the library picks for
    large K >= 3 batches; it is allocated here when the library asks for one and none is passed
    (``use_workspace=False`` withholds it, which selects the single-kernel path).
    """
    lib = _lib.load()
    obs = _narrow_f64(obs, tables)
    N, L, dt = _obs_args(obs, tables)
    dev = obs.device
    K = tables.K
    gamma = states = ll = None
    if want_gamma:
        shape = (N, L, K) if gamma_state < 0 else (N, L)
        gamma = out_gamma if out_gamma is not None else torch.empty(shape, dtype=torch.float32, device=dev)
        if tuple(gamma.shape) != shape or gamma.dtype != torch.float32 or not gamma.is_contiguous():
            raise ValueError("out_gamma has the wrong shape / dtype")
    if want_states:
        states = out_states if out_states is not None else torch.empty((N, L), dtype=torch.int8, device=dev)
        if tuple(states.shape) != (N, L) or states.dtype != torch.int8 or not states.is_contiguous():
            raise ValueError("out_states has the wrong shape / dtype")
    if want_loglik:
        ll = torch.empty((N,), dtype=torch.float64, device=dev)
    need = int(lib.smf_hmm_workspace_bytes(_lib.OP_POSTERIOR, tables.ref, N, L)) if use_workspace else 0
    if need > 0 and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
    ws_bytes = workspace.numel() * workspace.element_size() if (need > 0 and workspace is not None) else 0
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_posterior(tables.ref, _ptr(obs), dt, N, L, _bins_arg(bins, tables, L), _ptr(gamma),
                                   int(gamma_state), _ptr(states), _ptr(ll),
                                   _ptr(workspace) if ws_bytes else None, ws_bytes, _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_posterior")
    return gamma, states, ll


def viterbi(tables: _lib.HostTables, obs: torch.Tensor, bins: Optional[torch.Tensor] = None,
            *, out_states: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Viterbi decode (``smf_hmm_viterbi``): int8 states [N, L]."""
    lib = _lib.load()
    N, L, dt = _obs_args(obs, tables)
    dev = obs.device
    states = out_states if out_states is not None else torch.empty((N, L), dtype=torch.int8, device=dev)
    need = int(lib.smf_hmm_workspace_bytes(_lib.OP_VITERBI, tables.ref, N, L))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_viterbi(tables.ref, _ptr(obs), dt, N, L, _bins_arg(bins, tables, L), _ptr(states),
                                 _ptr(workspace), workspace.numel() * workspace.element_size(), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_viterbi")
    return states


def stats_len(tables: _lib.HostTables) -> int:
    return int(_lib.load().smf_hmm_stats_len(tables.ref))


def estep(tables: _lib.HostTables, obs: torch.Tensor, bins: Optional[torch.Tensor] = None,
          *, out_stats: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Baum-Welch E-step (``smf_hmm_estep``): fp64 statistics vector on the device."""
    lib = _lib.load()
    obs = _narrow_f64(obs, tables)
    N, L, dt = _obs_args(obs, tables)
    dev = obs.device
    S = stats_len(tables)
    stats = out_stats if out_stats is not None else torch.empty((S,), dtype=torch.float64, device=dev)
    need = int(lib.smf_hmm_workspace_bytes(_lib.OP_ESTEP, tables.ref, N, L))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need // 8, 1),), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_estep(tables.ref, _ptr(obs), dt, N, L, _bins_arg(bins, tables, L), _ptr(stats),
                               _ptr(workspace), workspace.numel() * workspace.element_size(), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_estep")
    return stats


def estep_workspace(tables: _lib.HostTables, device, n_reads: int = 1, n_pos: int = 1) -> torch.Tensor:
    """E-step scratch for batches of up to ``n_reads`` x ``n_pos`` (per-CTA partial statistics, plus the
    boundary vectors of the two-kernel path when the library would pick it for that batch shape)."""
    need = int(_lib.load().smf_hmm_workspace_bytes(_lib.OP_ESTEP, tables.ref, int(n_reads), int(n_pos)))
    return torch.empty((max(need // 8, 1),), dtype=torch.float64, device=device)


def _class_arrays(ranges: Sequence[Tuple[float, float]]):
    lo = np.ascontiguousarray([float(r[0]) for r in ranges], dtype=np.float64)
    hi = np.ascontiguousarray([float(r[1]) for r in ranges], dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    return lo, hi, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp)


def call_features(
    *,
    states: Optional[torch.Tensor],
    post: Optional[torch.Tensor],
    target: int,
    threshold: float,
    coords: torch.Tensor,
    grid_left: torch.Tensor,
    grid_right: torch.Tensor,
    n_grid: int,
    ranges: Sequence[Tuple[float, float]],
    want_len: bool = True,
    want_dense: bool = True,
    max_intervals: int = 0,
):
    """Feature calling (``smf_hmm_call_features``).

    Returns (class_out uint8 [N,V], run_len int32 [N,V] | None, intervals int32 [n,4] | None).
    Raises ``IntervalOverflow`` (a ``ValueError`` carrying the needed capacity) when more runs were found
    than ``max_intervals`` records fit.
    """
    lib = _lib.load()
    src = states if states is not None else post
    _require_cuda(src, "states/post")
    N, L = src.shape
    dev = src.device
    if states is not None and (states.dtype != torch.int8 or not states.is_contiguous()):
        raise ValueError("states must be contiguous int8")
    if post is not None and (post.dtype != torch.float32 or not post.is_contiguous()):
        raise ValueError("post must be contiguous float32")
    for t, dt_, nm in ((coords, torch.int64, "coords"), (grid_left, torch.int32, "grid_left"),
                       (grid_right, torch.int32, "grid_right")):
        _require_cuda(t, nm)
        if t.dtype != dt_ or t.numel() != L or not t.is_contiguous():
            raise ValueError(f"{nm} must be contiguous {dt_} of length L")
    V = int(n_grid)
    if not want_dense and max_intervals <= 0:
        raise ValueError("interval-only calling needs max_intervals > 0")
    cls = torch.empty((N, V), dtype=torch.uint8, device=dev) if want_dense else None
    rlen = torch.empty((N, V), dtype=torch.int32, device=dev) if (want_len and want_dense) else None
    ivals = cnt = None
    if max_intervals > 0:
        ivals = torch.empty((max_intervals, 4), dtype=torch.int32, device=dev)
        cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    lo, hi, plo, phi = _class_arrays(ranges)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_call_features(_ptr(states), _ptr(post), int(target), float(threshold), N, L,
                                       _ptr(coords), _ptr(grid_left), _ptr(grid_right), V, plo, phi, len(lo),
                                       _ptr(cls), _ptr(rlen), _ptr(ivals), int(max_intervals), _ptr(cnt),
                                       _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_call_features")
    if ivals is not None:
        found = int(cnt.item())
        if found > max_intervals:
            # the kernel counted every run but could only store max_intervals of them: never hand back a
            # silently truncated (and, being atomically appended, arbitrary) subset
            raise IntervalOverflow(found, max_intervals)
        ivals = ivals[:found]
    return cls, rlen, ivals


def merge_runs(layer: torch.Tensor, coords: Optional[torch.Tensor], coords_are_ints: bool,
               distance_threshold: int, ranges: Sequence[Tuple[float, float]] = ()):
    """Run merging + size classes (``smf_hmm_merge_runs``).

    Returns (merged uint8, merged_len int32, class_out uint8 | None, len_cls int32 | None), all [N, V].
    """
    lib = _lib.load()
    _require_cuda(layer, "layer")
    if layer.dtype != torch.uint8 or not layer.is_contiguous() or layer.dim() != 2:
        raise ValueError("layer must be a contiguous uint8 [N, V] tensor")
    N, V = layer.shape
    dev = layer.device
    if coords is not None:
        _require_cuda(coords, "coords")
        if coords.dtype != torch.int64 or coords.numel() != V or not coords.is_contiguous():
            raise ValueError("coords must be contiguous int64 of length V")
    merged = torch.empty((N, V), dtype=torch.uint8, device=dev)
    mlen = torch.empty((N, V), dtype=torch.int32, device=dev)
    cls = clen = None
    lo, hi, plo, phi = _class_arrays(ranges)
    if len(lo) > 0:
        cls = torch.empty((N, V), dtype=torch.uint8, device=dev)
        clen = torch.empty((N, V), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_merge_runs(_ptr(layer), N, V, _ptr(coords), int(bool(coords_are_ints)),
                                    int(distance_threshold), _ptr(merged), _ptr(mlen), plo, phi, len(lo),
                                    _ptr(cls), _ptr(clen), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_merge_runs")
    return merged, mlen, cls, clen


def merge_chain(layer: torch.Tensor, coords: Optional[torch.Tensor], coords_are_ints: bool,
                distance_threshold: int, ranges: Sequence[Tuple[float, float]] = (), *,
                covered: Optional[torch.Tensor] = None, span_coords: Optional[torch.Tensor] = None,
                start: Optional[torch.Tensor] = None, end: Optional[torch.Tensor] = None):
    """Fused merge -> size classes -> read-span mask (``smf_hmm_merge_chain``).

    Returns (code uint8, run_len int32), both [N, V]; see include/smf_hmm.h for the code bits.
    """
    lib = _lib.load()
    _require_cuda(layer, "layer")
    if layer.dtype != torch.uint8 or not layer.is_contiguous() or layer.dim() != 2:
        raise ValueError("layer must be a contiguous uint8 [N, V] tensor")
    N, V = layer.shape
    dev = layer.device
    if coords is not None:
        _require_cuda(coords, "coords")
        if coords.dtype != torch.int64 or coords.numel() != V or not coords.is_contiguous():
            raise ValueError("coords must be contiguous int64 of length V")
    if covered is not None:
        _require_cuda(covered, "covered")
        if covered.dtype != torch.uint8 or tuple(covered.shape) != (N, V) or not covered.is_contiguous():
            raise ValueError("covered must be a contiguous uint8 [N, V] tensor")
    else:
        if span_coords is None or start is None or end is None:
            raise ValueError("either covered or (span_coords, start, end) is required")
        for t, nm, dt, n in ((span_coords, "span_coords", torch.int64, V), (start, "start", torch.float64, N),
                             (end, "end", torch.float64, N)):
            _require_cuda(t, nm)
            if t.dtype != dt or t.numel() != n or not t.is_contiguous():
                raise ValueError(f"{nm} has the wrong dtype / length")
    lo, hi, plo, phi = _class_arrays(ranges)
    if len(lo) > 15:
        raise ValueError("at most 15 size classes fit the packed code")
    code = torch.empty((N, V), dtype=torch.uint8, device=dev)
    run_len = torch.empty((N, V), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_merge_chain(_ptr(layer), N, V, _ptr(coords), int(bool(coords_are_ints)),
                                     int(distance_threshold), plo, phi, len(lo), _ptr(covered), _ptr(span_coords),
                                     _ptr(start), _ptr(end), _ptr(code), _ptr(run_len), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_merge_chain")
    return code, run_len


def span_mask(n_reads: int, n_grid: int, device, *, covered: Optional[torch.Tensor] = None,
              coords: Optional[torch.Tensor] = None, start: Optional[torch.Tensor] = None,
              end: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Read-span validity bitmap (``smf_hmm_span_mask``): uint8 [N, V], 1 = keep, 0 = NaN."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("span_mask needs a CUDA device (no CPU fallback)")
    valid = torch.empty((n_reads, n_grid), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_span_mask(_ptr(covered), _ptr(coords), _ptr(start), _ptr(end), int(n_reads),
                                   int(n_grid), _ptr(valid), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_span_mask")
    return valid


# ---------------------------------------------------------------------------------------------
# boundary helpers: widening, layer expansion, explicit async copies (csrc/host_io.cu)
# ---------------------------------------------------------------------------------------------
def widen_decode(states8: Optional[torch.Tensor], states64: Optional[torch.Tensor],
                 gamma32: Optional[torch.Tensor], gamma64: Optional[torch.Tensor]) -> None:
    """int8 states -> int64 and float posteriors -> double in one launch (``smf_hmm_widen_decode``)."""
    lib = _lib.load()
    ref = states8 if states8 is not None else gamma32
    if ref is None:
        return
    _require_cuda(ref, "states/gamma")
    for t, dt, nm in ((states8, torch.int8, "states8"), (states64, torch.int64, "states64"),
                      (gamma32, torch.float32, "gamma32"), (gamma64, torch.float64, "gamma64")):
        if t is not None and (t.dtype != dt or not t.is_contiguous()):
            raise ValueError(f"{nm} must be contiguous {dt}")
    if (states8 is None) != (states64 is None) or (gamma32 is None) != (gamma64 is None):
        raise ValueError("widen_decode needs source and destination of each pair")
    ns = states8.numel() if states8 is not None else 0
    ng = gamma32.numel() if gamma32 is not None else 0
    if (states64 is not None and states64.numel() != ns) or (gamma64 is not None and gamma64.numel() != ng):
        raise ValueError("widen_decode: size mismatch")
    dev = ref.device
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_widen_decode(_ptr(states8), _ptr(states64), ns, _ptr(gamma32), _ptr(gamma64), ng,
                                      _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_widen_decode")


def expand_layers(n_reads: int, n_grid: int, n_pos: int, *, states: Optional[torch.Tensor],
                  post: Optional[torch.Tensor], post_stride: int, site_of_col: Optional[torch.Tensor],
                  valid: Optional[torch.Tensor], groups: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor], bool]],
                  layers: Sequence[Tuple[int, int, int, torch.Tensor]]) -> None:
    """Write final dense layers from packed results (``smf_hmm_expand_layers``).

    ``groups``: ``(class_or_code uint8 [N,V], run_len int32 [N,V] | None, packed)``;
    ``layers``: ``(kind, group, cls, out)`` with ``out`` a float64 / float32 CUDA tensor of N*V elements.
    ``post`` is the tensor whose element ``(n, t)`` sits at ``post.data_ptr() + 4 * (n*L + t) * post_stride``.
    """
    lib = _lib.load()
    if not layers:
        return
    dev = layers[0][3].device
    G = (_lib.LayerGroup * max(1, len(groups)))()
    for i, (cls, rl, packed) in enumerate(groups):
        _require_cuda(cls, "group class/code")
        if cls.dtype != torch.uint8 or cls.numel() != n_reads * n_grid or not cls.is_contiguous():
            raise ValueError("group class/code must be contiguous uint8 [N, V]")
        if rl is not None and (rl.dtype != torch.int32 or rl.numel() != n_reads * n_grid or not rl.is_contiguous()):
            raise ValueError("group run_len must be contiguous int32 [N, V]")
        G[i] = _lib.LayerGroup(_ptr(cls), _ptr(rl), 1 if packed else 0, 0)
    D = (_lib.LayerDesc * len(layers))()
    for i, (kind, group, cls_idx, out) in enumerate(layers):
        _require_cuda(out, "layer output")
        if out.dtype not in (torch.float64, torch.float32) or out.numel() != n_reads * n_grid or not out.is_contiguous():
            raise ValueError("layer outputs must be contiguous float64 / float32 tensors of N*V elements")
        D[i] = _lib.LayerDesc(_ptr(out), int(kind), int(group), int(cls_idx), 1 if out.dtype == torch.float32 else 0)
    if site_of_col is not None and (site_of_col.dtype != torch.int32 or site_of_col.numel() != n_grid):
        raise ValueError("site_of_col must be int32 of length V")
    if valid is not None and (valid.dtype != torch.uint8 or valid.numel() != n_reads * n_grid or not valid.is_contiguous()):
        raise ValueError("valid must be contiguous uint8 [N, V]")
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_expand_layers(int(n_reads), int(n_grid), int(n_pos), _ptr(states), _ptr(post), int(post_stride),
                                       _ptr(site_of_col), _ptr(valid), G, len(groups), D, len(layers), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_expand_layers")


def copy_to_host_async(dst: np.ndarray, src: torch.Tensor, stream: torch.cuda.Stream) -> None:
    """cudaMemcpyAsync device -> host on ``stream`` (link speed when ``dst`` is page-locked)."""
    if not (dst.flags.c_contiguous and src.is_contiguous()):
        raise ValueError("copy_to_host_async needs contiguous source and destination")
    nbytes = src.numel() * src.element_size()
    if dst.nbytes != nbytes:
        raise ValueError(f"copy_to_host_async: {dst.nbytes} destination bytes for {nbytes} source bytes")
    with torch.cuda.device(src.device):
        rc = _lib.load().smf_hmm_copy_async(ctypes.c_void_p(dst.ctypes.data), _ptr(src), nbytes, _lib.COPY_D2H,
                                            int(stream.cuda_stream))
    _lib.check(rc, "smf_hmm_copy_async")


def copy_to_device_async(dst: torch.Tensor, src: np.ndarray, stream: torch.cuda.Stream) -> None:
    """cudaMemcpyAsync host -> device on ``stream``."""
    if not (src.flags.c_contiguous and dst.is_contiguous()):
        raise ValueError("copy_to_device_async needs contiguous source and destination")
    nbytes = dst.numel() * dst.element_size()
    if src.nbytes != nbytes:
        raise ValueError(f"copy_to_device_async: {src.nbytes} source bytes for {nbytes} destination bytes")
    with torch.cuda.device(dst.device):
        rc = _lib.load().smf_hmm_copy_async(_ptr(dst), ctypes.c_void_p(src.ctypes.data), nbytes, _lib.COPY_H2D,
                                            int(stream.cuda_stream))
    _lib.check(rc, "smf_hmm_copy_async")


def build_obs(layer: torch.Tensor, cols: torch.Tensor, n_channels: int, *, coded_layer: bool = False,
              covered: Optional[torch.Tensor] = None, covered_cols: Optional[torch.Tensor] = None,
              out_u8: bool = True, out: Optional[torch.Tensor] = None, flag: Optional[torch.Tensor] = None):
    """Model-input builder (``smf_hmm_build_obs``): gathers the site columns ``cols`` (int32 [L, C], -1 = no
    site) of ``layer`` [N, V] (float32 / float64 with NaN = missing, or -- ``coded_layer`` -- the uint8 code)
    into ``[N, L(, C)]`` observations, missing where ``covered`` (uint8 [N, VC]) is 0 at ``covered_cols[l]``.
    Returns ``(obs, flag)``; ``flag`` (device int32) becomes 1 when the output type cannot carry a value exactly."""
    lib = _lib.load()
    _require_cuda(layer, "layer")
    if layer.dim() != 2 or not layer.is_contiguous():
        raise ValueError("layer must be a contiguous [N, V] tensor")
    if coded_layer:
        if layer.dtype != torch.uint8:
            raise ValueError("a coded layer must be uint8")
        ldt = _lib.OBS_U8
    elif layer.dtype in (torch.float32, torch.float64):
        ldt = _lib.OBS_F32 if layer.dtype == torch.float32 else _lib.OBS_F64
    else:
        raise ValueError(f"unsupported layer dtype {layer.dtype}")
    N, V = int(layer.shape[0]), int(layer.shape[1])
    C = int(n_channels)
    if cols.dtype != torch.int32 or not cols.is_contiguous() or cols.numel() % C:
        raise ValueError("cols must be a contiguous int32 [L, C] tensor")
    L = cols.numel() // C
    dev = layer.device
    vc = 0
    if covered is not None:
        if covered.dtype != torch.uint8 or covered.dim() != 2 or int(covered.shape[0]) != N or not covered.is_contiguous():
            raise ValueError("covered must be a contiguous uint8 [N, VC] tensor")
        if covered_cols is None or covered_cols.dtype != torch.int32 or covered_cols.numel() != L:
            raise ValueError("covered_cols must be int32 of length L")
        vc = int(covered.shape[1])
    shape = (N, L) if C == 1 else (N, L, C)
    odt = torch.uint8 if out_u8 else torch.float32
    if out is None:
        out = torch.empty(shape, dtype=odt, device=dev)
    elif out.dtype != odt or out.numel() != N * L * C or not out.is_contiguous():
        raise ValueError("out has the wrong dtype / size")
    if flag is None:
        flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.smf_hmm_build_obs(_ptr(layer), ldt, N, V, _ptr(cols), L, C, _ptr(covered), vc, _ptr(covered_cols),
                                   _ptr(out), _lib.OBS_U8 if out_u8 else _lib.OBS_F32, _ptr(flag), _stream_ptr(dev))
    _lib.check(rc, "smf_hmm_build_obs")
    return out, flag
