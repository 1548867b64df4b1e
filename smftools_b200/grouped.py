"""Grouped launches: many (model, N_g, L_g) groups per call (``smf_hmm_grouped_fit`` / ``_posterior``).

The reference fits and applies one HMM per (reference x methylation type [x sample / barcode]) task, one after
the other: ``HMMTrainer.fit_or_load`` + ``annotate_adata`` per task (cli/hmm_adata.py:488-524), at most
``hmm_max_fit_reads`` = 1000 reads per fit (hmm/fit_plan.py).  A group that small cannot fill a GPU and every EM
iteration of every group would cost a host round trip.  Here a whole list of groups -- each with its own model
and its own ragged ``(N_g, L_g)`` observation matrix -- is fitted / decoded in ONE launch sequence; the
Baum-Welch loop (E-step, per-group reduction, M-step, stop rule) stays on the device.

Multi-GPU (SURVEY.md 8e "Multi-group"): groups are dealt WHOLE to the ranks, largest first (``deal_groups``); a
rank fits its own groups with no collective and the fitted parameters (a few doubles per group) are exchanged
once at the end (``exchange_group_parameters``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream_ptr

FIT_RAN_ALL, FIT_CONVERGED, FIT_HANDED_BACK = 0, 1, 2


# --------------------------------------------------------------------------------------------------------
# host-side planning
# --------------------------------------------------------------------------------------------------------
def deal_groups(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Assign whole groups to ranks, largest first onto the least-loaded rank (LPT).  ``costs[g]`` is the
    work of group g (read-positions).  Returns, per rank, the group indices it owns in ascending order.
    Deterministic: ties go to the lower group index / lower rank."""
    world_size = int(world_size)
    load = [0.0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for g in sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        owned[r].append(g)
        load[r] += float(costs[g])
    return [sorted(o) for o in owned]


def partition_reads(n_reads: Sequence[int], n_pos: Sequence[int], world_size: int) -> List[List[Tuple[int, int, int]]]:
    """Balanced alternative to ``deal_groups`` for groups that are large next to a rank's share: lay the groups end to
    end on a cost axis (cost of a read = its length), cut the axis into ``world_size`` equal pieces and give rank r the
    reads inside piece r.  Returns, per rank, ``(group, lo, hi)`` read ranges in group order; a group that crosses a
    cut is read-sharded over the ranks on either side (its statistics are then all-reduced: ``fit_groups(...,
    process_group=...)``).  Every read is owned by exactly one rank; loads differ by at most one read per cut."""
    world_size = int(world_size)
    costs = [int(n) * int(l) for n, l in zip(n_reads, n_pos)]
    total = sum(costs)
    out: List[List[Tuple[int, int, int]]] = [[] for _ in range(world_size)]
    if total == 0:
        return out
    cuts = [(total * r) // world_size for r in range(world_size + 1)]
    c0 = 0
    for g, (n, l) in enumerate(zip(n_reads, n_pos)):
        n, l = int(n), int(l)
        c1 = c0 + n * l
        for r in range(world_size):
            lo_c, hi_c = max(cuts[r], c0), min(cuts[r + 1], c1)
            if lo_c >= hi_c:
                continue
            # read i of the group covers [c0 + i*l, c0 + (i+1)*l): it belongs to the piece that holds its first position
            lo = -(-(lo_c - c0) // l)
            hi = -(-(hi_c - c0) // l)
            if hi > lo:
                out[r].append((g, lo, hi))
        c0 = c1
    return out


def pack_parameters(models) -> np.ndarray:
    """[G, K + K*K + K] fp64: start | trans | emission of every model (the layout of the C-ABI)."""
    rows = []
    for m in models:
        rows.append(np.concatenate([
            m.start.detach().to("cpu", torch.float64).numpy().reshape(-1),
            m.trans.detach().to("cpu", torch.float64).numpy().reshape(-1),
            m.emission.detach().to("cpu", torch.float64).numpy().reshape(-1)]))
    return np.ascontiguousarray(np.stack(rows), dtype=np.float64)


def unpack_parameters(models, params: np.ndarray) -> None:
    """Write rows of ``pack_parameters`` layout back into the models (their dtype / device are kept)."""
    for m, row in zip(models, np.asarray(params, dtype=np.float64)):
        K = m.n_states
        with torch.no_grad():
            m.start.data = torch.as_tensor(row[:K].copy(), dtype=m.start.dtype).to(m.start.device)
            m.trans.data = torch.as_tensor(row[K:K + K * K].reshape(K, K).copy(), dtype=m.trans.dtype).to(m.trans.device)
            m.emission.data = torch.as_tensor(row[K + K * K:].copy(), dtype=m.emission.dtype).to(m.emission.device)


def exchange_group_parameters(models, owned: Sequence[Sequence[int]], process_group=None) -> None:
    """After every rank fitted the groups ``owned[rank]``, give every rank all fitted parameters: one
    all-reduce of a [G, P] fp64 matrix in which each rank filled only its own rows (works with gloo and nccl)."""
    import torch.distributed as dist

    rank = dist.get_rank(process_group)
    packed = pack_parameters(models)
    mine = np.zeros_like(packed)
    for g in owned[rank]:
        mine[g] = packed[g]
    t = torch.from_numpy(mine)
    if dist.get_backend(process_group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    unpack_parameters(models, t.cpu().numpy())


# --------------------------------------------------------------------------------------------------------
# device-side calls
# --------------------------------------------------------------------------------------------------------
def _check_models(models) -> Tuple[int, float]:
    from .hmm import DistanceBinnedSingleBernoulliHMM, SingleBernoulliHMM

    if len(models) == 0:
        raise ValueError("no groups")
    K = int(models[0].n_states)
    eps = float(models[0].eps)
    for m in models:
        if not isinstance(m, SingleBernoulliHMM) or isinstance(m, DistanceBinnedSingleBernoulliHMM):
            raise ValueError("grouped launches support arch 'single' (SingleBernoulliHMM) only")
        if int(m.n_states) != K:
            raise ValueError("all groups of one call must have the same number of states")
        if float(m.eps) != eps:
            raise ValueError("all groups of one call must share eps")
    return K, eps


def _group_obs(X, device) -> Tuple[torch.Tensor, int, int]:
    """Device observation matrix of one group in a layout the kernels accept: uint8 0 / 1 / other = missing code
    (``CodedU8`` or an already-coded device tensor) or float32 with NaN = missing; rows padded to whole 16-byte
    groups.  Returns (tensor [N, stride], L, obs_dtype)."""
    from .hmm import BaseHMM

    src, numeric_u8 = BaseHMM._as_obs_source(X)
    if src.ndim != 2:
        raise ValueError(f"grouped launches expect X of shape (N, L); got {tuple(src.shape)}")
    if isinstance(src, np.ndarray):
        src = torch.from_numpy(np.ascontiguousarray(src))
    t = src.to(device)
    if t.dtype == torch.uint8 and numeric_u8:  # plain uint8: the byte value is the observation (HMM.py:694)
        t = t.to(torch.float32)
    elif t.dtype == torch.float64:
        t = t.to(torch.float32)
    N, L = int(t.shape[0]), int(t.shape[1])
    unit = 16 if t.dtype == torch.uint8 else 4
    Lp = (L + unit - 1) // unit * unit
    es = t.element_size()
    # usable as is: unit-stride columns, rows on 16-byte boundaries, and every row's pad (up to Lp) addressable
    ok = (N > 0 and t.stride(1) == 1 and t.stride(0) >= Lp and t.stride(0) % unit == 0 and t.data_ptr() % 16 == 0
          and t.untyped_storage().nbytes() >= (t.storage_offset() + (N - 1) * t.stride(0) + Lp) * es)
    if not ok:
        buf = torch.empty((N, Lp), dtype=t.dtype, device=device)
        buf[:, :L] = t
        if Lp > L:
            buf[:, L:] = 255 if t.dtype == torch.uint8 else float("nan")
        t = buf[:, :L]
    return t, L, (_lib.OBS_U8 if t.dtype == torch.uint8 else _lib.OBS_F32)


class _GroupTable:
    """Host array of ``smf_hmm_group`` + the tensors it points at (kept alive)."""

    def __init__(self, obs_list, device):
        self.device = device
        self.obs = []
        dts = set()
        for X in obs_list:
            t, L, dt = _group_obs(X, device)
            self.obs.append((t, L))
            dts.add(dt)
        if len(dts) != 1:
            raise ValueError("all groups of one call must use the same observation type (coded uint8 or float32)")
        self.obs_dtype = dts.pop()
        self.G = len(self.obs)
        self.arr = (_lib.Group * self.G)()
        for g, (t, L) in enumerate(self.obs):
            q = self.arr[g]
            q.obs = int(t.data_ptr())
            q.n_reads = int(t.shape[0])
            q.n_pos = L
            q.row_stride = int(t.stride(0))
            q.gamma = None
            q.states = None
            q.loglik = None
            q.gamma_state = -1
            q.out_stride = 0

    def positions(self) -> List[int]:
        return [int(t.shape[0]) * L for t, L in self.obs]


def fit_groups(models, obs_list, *, max_iter: int = 50, tol: float = 1e-5, update_start: bool = True,
               update_trans: bool = True, update_emission: bool = True, device=None,
               return_status: bool = False, process_group=None, sharded: bool = False):
    """Baum-Welch fit of every ``models[g]`` on ``obs_list[g]`` (``BaseHMM.fit`` semantics per group:
    HMM.py:724-781, 1518-1630) in one device-resident loop.  The models are updated in place; returns the
    per-group log-likelihood histories (and, with ``return_status``, the per-group SMF_FIT_* codes).

    ``obs_list[g]``: (N_g, L_g) ``CodedU8`` / coded uint8 device tensor, or float matrix with NaN = missing.
    A group whose parameters leave the dynamic range of the scaled fp32 kernels is finished through the
    model's own ``fit`` (same result as fitting it alone).

    ``sharded=True`` (or an explicit ``process_group``): the reads of the groups are sharded over the ranks of the
    process group (``partition_reads``).  Every rank passes ALL groups in the same order, each with its own shard of
    that group's reads (possibly zero rows); per iteration the ``[G, S]`` statistics matrix is sum-all-reduced (NCCL,
    stream-ordered: still no host round trip) between the device E-step and the device M-step, so every rank ends
    with bit-identical parameters for every group."""
    K, eps = _check_models(models)
    if len(models) != len(obs_list):
        raise ValueError("models and obs_list differ in length")
    device = models[0]._ensure_device_dtype(device)
    for m in models[1:]:
        m._ensure_device_dtype(device)
    lib = _lib.load()
    max_iter = int(max_iter)
    G = len(models)
    if max_iter <= 0:
        return ([[] for _ in models], [FIT_RAN_ALL] * G) if return_status else [[] for _ in models]
    tab = _GroupTable(obs_list, device)
    init = pack_parameters(models)
    with torch.cuda.device(device):
        # parameters | log-likelihood histories | iteration counts | stop reasons share one buffer: one fill, one read-back
        P = init.shape[1]
        o_hist, o_n, o_st, o_end = 8 * G * P, 8 * G * (P + max_iter), 8 * G * (P + max_iter) + 4 * G, 8 * G * (P + max_iter + 1)
        buf = torch.zeros((o_end,), dtype=torch.uint8, device=device)
        params = buf[:o_hist].view(torch.float64).view(G, P)
        params.copy_(torch.from_numpy(init))
        hist = buf[o_hist:o_n].view(torch.float64).view(G, max_iter)
        n_iter = buf[o_n:o_st].view(torch.int32)
        status = buf[o_st:o_end].view(torch.int32)
        need = int(lib.smf_hmm_grouped_workspace_bytes(K, tab.arr, G, 1))
        if need == 0:
            raise ValueError("smf_hmm_grouped_workspace_bytes rejected the group table")
        ws = torch.empty((need,), dtype=torch.uint8, device=device)
        mask = (1 if update_start else 0) | (2 if update_trans else 0) | (4 if update_emission else 0)
        if sharded or process_group is not None:
            import torch.distributed as dist

            S_len = K + K * K + 2 * K + 1
            stats = torch.zeros((G, S_len), dtype=torch.float64, device=device)

            def step(phase, it):
                rc = lib.smf_hmm_grouped_fit_step(phase, K, tab.obs_dtype, tab.arr, G, _ptr(params), it, max_iter,
                                                  float(tol), eps, mask, _ptr(hist), _ptr(n_iter), _ptr(status),
                                                  _ptr(stats), _ptr(ws), need, _stream_ptr(device))
                _lib.check(rc, "smf_hmm_grouped_fit_step")

            step(_lib.FIT_PHASE_BEGIN, 0)
            for it in range(max_iter):
                step(_lib.FIT_PHASE_ESTEP, it)
                if dist.is_available() and dist.is_initialized():
                    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=process_group)
                step(_lib.FIT_PHASE_MSTEP, it)
        else:
            rc = lib.smf_hmm_grouped_fit(K, tab.obs_dtype, tab.arr, G, _ptr(params), max_iter, float(tol), eps, mask,
                                         _ptr(hist), _ptr(n_iter), _ptr(status), _ptr(ws), need, _stream_ptr(device))
            _lib.check(rc, "smf_hmm_grouped_fit")
        # one read-back for the whole fit
        host = buf.cpu().numpy()
        params_h = host[:o_hist].view(np.float64).reshape(G, P)
        hist_h = host[o_hist:o_n].view(np.float64).reshape(G, max_iter)
        n_h = host[o_n:o_st].view(np.int32)
        st_h = host[o_st:o_end].view(np.int32)
    out: List[List[float]] = []
    unpack_parameters(models, params_h)
    for g, m in enumerate(models):
        if int(st_h[g]) == FIT_HANDED_BACK:
            unpack_parameters([m], init[g:g + 1])
            t, L = tab.obs[g]
            from .hmm import CodedU8

            X = CodedU8(t) if t.dtype == torch.uint8 else t
            m._host_loop_only = True  # the general path: host M-step, kernels with a finer rescale cadence
            was = (m.distributed, m.process_group)
            if sharded or process_group is not None:
                # every rank reaches this point for the same groups (the M-step is replicated): the host loop
                # all-reduces the statistics of the shards like the device loop did
                m.distributed, m.process_group = True, process_group
            try:
                out.append(m.fit(X, max_iter=max_iter, tol=tol, device=device, update_start=update_start,
                                 update_trans=update_trans, update_emission=update_emission))
            finally:
                m._host_loop_only = False
                m.distributed, m.process_group = was
        else:
            out.append([float(v) for v in hist_h[g, :int(n_h[g])]])
    if return_status:
        return out, [int(v) for v in st_h]
    return out


def decode_groups(models, obs_list, *, want_gamma: bool = True, gamma_state: int = -1, want_states: bool = True,
                  want_loglik: bool = False, device=None):
    """Posterior decode of every group with its own model in one launch pair.  Returns a list of
    ``(gamma, states, loglik)`` device tensors per group (float32 [N_g, L_g, K] or [N_g, L_g]; int8 [N_g, L_g];
    float64 [N_g]); entries not asked for are None.  gamma / states are views of row-padded buffers."""
    K, _eps = _check_models(models)
    if len(models) != len(obs_list):
        raise ValueError("models and obs_list differ in length")
    device = models[0]._ensure_device_dtype(device)
    lib = _lib.load()
    tab = _GroupTable(obs_list, device)
    G = tab.G
    host_tabs = [m._host_tables() for m in models]
    tarr = (_lib.Tables * G)(*[h.struct for h in host_tabs])
    results = []
    keep = []
    with torch.cuda.device(device):
        for g, (t, L) in enumerate(tab.obs):
            N = int(t.shape[0])
            Ls = (L + 15) // 16 * 16  # rows on 16-byte boundaries: one 128-bit store per thread and chunk of states
            gam = st = ll = None
            q = tab.arr[g]
            q.out_stride = Ls
            q.gamma_state = int(gamma_state)
            if want_gamma:
                full = torch.empty((N, Ls, K) if gamma_state < 0 else (N, Ls), dtype=torch.float32, device=device)
                q.gamma = int(full.data_ptr())
                gam = full[:, :L]
                keep.append(full)
            if want_states:
                full = torch.empty((N, Ls), dtype=torch.int8, device=device)
                q.states = int(full.data_ptr())
                st = full[:, :L]
                keep.append(full)
            if want_loglik:
                ll = torch.empty((N,), dtype=torch.float64, device=device)
                q.loglik = int(ll.data_ptr())
            results.append((gam, st, ll))
        need = int(lib.smf_hmm_grouped_workspace_bytes(K, tab.arr, G, 0))
        if need == 0:
            raise ValueError("smf_hmm_grouped_workspace_bytes rejected the group table")
        ws = torch.empty((need,), dtype=torch.uint8, device=device)
        rc = lib.smf_hmm_grouped_posterior(tab.obs_dtype, tab.arr, G, tarr, _ptr(ws), need, _stream_ptr(device))
        _lib.check(rc, "smf_hmm_grouped_posterior")
        # the workspace and the padded buffers must outlive the launches: tie them to the stream
        ws.record_stream(torch.cuda.current_stream(device))
    return results
