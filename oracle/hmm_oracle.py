"""CPU oracle: numpy fp64 restatement of the reference's HMM footprinting arithmetic.

TEST INFRASTRUCTURE ONLY.  Nothing under ``smftools_b200/`` imports this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker / the timed CPU port, never as the shipped path.

Parity pinning: the reference holds no golden vectors for this path, so the
oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the
unmodified ``/root/reference/src/smftools/hmm/HMM.py`` in the build container and commits its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function here
against those files, and ``tests/test_oracle_vs_reference.py`` re-runs the comparison live
whenever the reference tree is present.

Every function cites the reference lines it restates (paths relative to
``/root/reference/src/smftools/hmm/``).  Like the reference it is "batched over reads,
sequential over positions": python ``for t in range(L)`` loops around small numpy ops.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_BINS = (1, 5, 10, 25, 50, 100)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
class Params:
    """start (K,), trans (K,K), emission (K,) or (K,C), optional trans_by_bin (B,K,K)."""

    def __init__(self, start, trans, emission, eps=1e-8, trans_by_bin=None, distance_bins=None):
        self.start = np.array(start, dtype=np.float64)
        self.trans = np.array(trans, dtype=np.float64)
        self.emission = np.array(emission, dtype=np.float64)
        self.eps = float(eps)
        self.trans_by_bin = None if trans_by_bin is None else np.array(trans_by_bin, dtype=np.float64)
        self.distance_bins = np.asarray(DEFAULT_BINS if distance_bins is None else distance_bins, dtype=int)

    @property
    def K(self):
        return self.start.shape[0]

    def copy(self):
        return Params(self.start, self.trans, self.emission, self.eps, self.trans_by_bin, self.distance_bins)

    def normalize(self):
        """HMM.py:492-511 (start/trans), :1470-1478 / :1749-1753 (emission clamp),
        :2076-2083 (trans_by_bin).  Runs after construction and after every M-step."""
        eps = self.eps
        s = self.start + eps
        self.start = s / s.sum()
        t = self.trans + eps
        rs = t.sum(axis=1, keepdims=True)
        rs[rs == 0.0] = 1.0
        self.trans = t / rs
        self.emission = np.clip(self.emission, eps, 1.0 - eps)
        if self.trans_by_bin is not None:
            tb = self.trans_by_bin + eps
            rs = tb.sum(axis=2, keepdims=True)
            rs[rs == 0.0] = 1.0
            self.trans_by_bin = tb / rs
        return self


def init_params(K=2, init_emission=None, eps=1e-8, arch="single", n_channels=2, distance_bins=None) -> Params:
    """Initial parameters: uniform start/trans (HMM.py:427-432); emission from the flattened
    ``hmm_init_emission_probs`` or 0.5 (single :1434-1442; multi :1705-1718); per-bin copies of the
    normalised ``trans`` for the distance-binned model (:2020-2030)."""
    start = np.full((K,), 1.0 / K)
    trans = np.full((K, K), 1.0 / K)
    if arch == "multi":
        C = int(n_channels)
        em = None
        if init_emission is not None:
            arr = np.asarray(init_emission, dtype=float)
            cand = np.repeat(arr.reshape(-1, 1)[:K, :], C, axis=1) if arr.ndim == 1 else arr[:K, :C]
            if cand.shape == (K, C):
                em = cand
        if em is None:
            em = np.full((K, C), 0.5)
    else:
        em = None
        if init_emission is not None:
            flat = np.asarray(init_emission, dtype=float).reshape(-1)[:K]
            if flat.size == K:
                em = flat
        if em is None:
            em = np.full((K,), 0.5)
    p = Params(start, trans, em, eps=eps, distance_bins=distance_bins)
    p.normalize()
    if arch == "single_distance_binned":
        nb = len(p.distance_bins) + 1
        p.trans_by_bin = np.repeat(p.trans[None, :, :], nb, axis=0)
        tb = p.trans_by_bin + eps  # _normalize_trans_by_bin at construction (:2030)
        p.trans_by_bin = tb / tb.sum(axis=2, keepdims=True)
    return p


def bin_index(coords, distance_bins) -> np.ndarray:
    """HMM.py:2115-2125: bin of each step from the bp distance, ``digitize(right=True)``."""
    return np.digitize(np.diff(np.asarray(coords, dtype=int)), np.asarray(distance_bins, dtype=int), right=True)


def log_tables(p: Params):
    """log(start+eps), log(trans[_by_bin]+eps), log(p+eps), log1p(-p+eps)
    (HMM.py:594-595, 1495-1496, 1774-1775, 2149-2150)."""
    eps = p.eps
    log_start = np.log(p.start + eps)
    stack = p.trans_by_bin if p.trans_by_bin is not None else p.trans[None, :, :]
    log_trans = np.log(stack + eps)
    log_e1 = np.log(p.emission + eps)
    log_e0 = np.log1p(-p.emission + eps)
    return log_start, log_trans, log_e1, log_e0


def _lse(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(x - m), axis=axis))


# --------------------------------------------------------------------------------------
# emissions / forward-backward / viterbi
# --------------------------------------------------------------------------------------
def split_obs(X):
    """HMM.py:707-708: obs = nan_to_num(X), mask = ~isnan(X)."""
    X = np.asarray(X, dtype=np.float64)
    return np.nan_to_num(X, nan=0.0), ~np.isnan(X)


def log_emission(obs, mask, log_e1, log_e0):
    """Single (HMM.py:1490-1501): logB[n,t,k] = o*log(p_k+eps) + (1-o)*log1p(-p_k+eps), 0 where
    masked.  Multi (:1766-1782): the same per channel, summed over channels."""
    if obs.ndim == 2:
        o = obs[:, :, None]
        lb = o * log_e1.reshape(1, 1, -1) + (1.0 - o) * log_e0.reshape(1, 1, -1)
        return np.where(mask[:, :, None], lb, 0.0)
    o = obs[:, :, None, :]
    lb = o * log_e1[None, None, :, :] + (1.0 - o) * log_e0[None, None, :, :]
    lb = np.where(mask[:, :, None, :], lb, 0.0)
    return lb.sum(axis=3)


def alpha_beta(logB, log_start, log_trans, bins=None):
    """HMM.py:572-612 (homogeneous) and :2127-2170 (transition matrix chosen by ``bins[t-1]``)."""
    N, L, K = logB.shape
    alpha = np.empty((N, L, K))
    alpha[:, 0, :] = log_start[None, :] + logB[:, 0, :]
    for t in range(1, L):
        logA = log_trans[bins[t - 1] if bins is not None else 0]
        alpha[:, t, :] = _lse(alpha[:, t - 1, :, None] + logA[None, :, :], axis=1) + logB[:, t, :]
    beta = np.empty((N, L, K))
    beta[:, L - 1, :] = 0.0
    for t in range(L - 2, -1, -1):
        logA = log_trans[bins[t] if bins is not None else 0]
        beta[:, t, :] = _lse(logA[None, :, :] + (logB[:, t + 1, :] + beta[:, t + 1, :])[:, None, :], axis=2)
    return alpha, beta


def posterior(X, p: Params, coords=None):
    """gamma = softmax_k(alpha+beta) (HMM.py:614-629); returns (gamma, per-read logZ)."""
    obs, mask = split_obs(X)
    log_start, log_trans, le1, le0 = log_tables(p)
    bins = bin_index(coords, p.distance_bins) if p.trans_by_bin is not None else None
    logB = log_emission(obs, mask, le1, le0)
    alpha, beta = alpha_beta(logB, log_start, log_trans, bins)
    lg = alpha + beta
    gamma = np.exp(lg - _lse(lg, axis=2)[:, :, None])
    return gamma, _lse(alpha[:, -1, :], axis=1)


def viterbi(X, p: Params, coords=None, return_margins=False):
    """HMM.py:631-671 (binned :2190-2236): max-plus forward sweep with first-index arg-max,
    back-trace through psi.  With ``return_margins`` also returns, per read, the smallest gap
    between the best and second-best candidate over every decision taken (used by the tests to
    recognise the 1e-6 near-tie exemption)."""
    obs, mask = split_obs(X)
    log_start, log_trans, le1, le0 = log_tables(p)
    bins = bin_index(coords, p.distance_bins) if p.trans_by_bin is not None else None
    logB = log_emission(obs, mask, le1, le0)
    N, L, K = logB.shape
    delta = log_start[None, :] + logB[:, 0, :]
    psi = np.zeros((N, L, K), dtype=np.int64)
    margin = np.full((N,), np.inf)
    for t in range(1, L):
        logA = log_trans[bins[t - 1] if bins is not None else 0]
        cand = delta[:, :, None] + logA[None, :, :]  # (N, i, j)
        best = cand.argmax(axis=1)
        psi[:, t, :] = best
        if return_margins:
            srt = np.sort(cand, axis=1)
            margin = np.minimum(margin, (srt[:, -1, :] - srt[:, -2, :]).min(axis=1))
        delta = cand.max(axis=1) + logB[:, t, :]
    states = np.empty((N, L), dtype=np.int64)
    states[:, L - 1] = delta.argmax(axis=1)
    if return_margins:
        srt = np.sort(delta, axis=1)
        margin = np.minimum(margin, srt[:, -1] - srt[:, -2])
    rows = np.arange(N)
    for t in range(L - 2, -1, -1):
        states[:, t] = psi[rows, t + 1, states[:, t + 1]]
    if return_margins:
        return states, margin
    return states


def path_log_prob(X, states, p: Params, coords=None):
    """Joint log-probability of a state path (for tie-tolerant Viterbi checks)."""
    obs, mask = split_obs(X)
    log_start, log_trans, le1, le0 = log_tables(p)
    bins = bin_index(coords, p.distance_bins) if p.trans_by_bin is not None else None
    logB = log_emission(obs, mask, le1, le0)
    N, L, K = logB.shape
    rows = np.arange(N)
    lp = log_start[states[:, 0]] + logB[rows, 0, states[:, 0]]
    for t in range(1, L):
        logA = log_trans[bins[t - 1] if bins is not None else 0]
        lp = lp + logA[states[:, t - 1], states[:, t]] + logB[rows, t, states[:, t]]
    return lp


def decode(X, p: Params, coords=None, decode="marginal"):
    """HMM.py:673-720: always the posterior; states = Viterbi path or argmax gamma."""
    gamma, _ = posterior(X, p, coords)
    if str(decode).lower() == "viterbi":
        st = viterbi(X, p, coords)
    else:
        st = gamma.argmax(axis=2)
    return st, gamma


# --------------------------------------------------------------------------------------
# Baum-Welch
# --------------------------------------------------------------------------------------
def estep(X, p: Params, coords=None) -> Dict[str, np.ndarray]:
    """Sufficient statistics of one EM iteration (HMM.py:1566-1596; multi :1849-1882; binned
    :2289-2320): start_acc, trans_acc (per bin), emit_num, emit_den, log-likelihood."""
    obs, mask = split_obs(X)
    log_start, log_trans, le1, le0 = log_tables(p)
    binned = p.trans_by_bin is not None
    bins = bin_index(coords, p.distance_bins) if binned else None
    logB = log_emission(obs, mask, le1, le0)
    alpha, beta = alpha_beta(logB, log_start, log_trans, bins)
    N, L, K = logB.shape
    lg = alpha + beta
    gamma = np.exp(lg - _lse(lg, axis=2)[:, :, None])
    start_acc = gamma[:, 0, :].sum(axis=0)
    ll = float(_lse(alpha[:, L - 1, :], axis=1).sum())
    nb = log_trans.shape[0]
    trans_acc = np.zeros((nb, K, K))
    valid_pos = mask if obs.ndim == 2 else mask.any(axis=2)
    for t in range(L - 1):
        b = bins[t] if binned else 0
        valid = (valid_pos[:, t] & valid_pos[:, t + 1]).astype(np.float64)[:, None, None]
        log_xi = alpha[:, t, :, None] + log_trans[b][None, :, :] + (logB[:, t + 1, :] + beta[:, t + 1, :])[:, None, :]
        norm = _lse(log_xi.reshape(N, -1), axis=1)[:, None, None]
        trans_acc[b] += (np.exp(log_xi - norm) * valid).sum(axis=0)
    if obs.ndim == 2:
        mf = mask.astype(np.float64)[:, :, None]
        emit_num = (gamma * obs[:, :, None] * mf).sum(axis=(0, 1))
        emit_den = (gamma * mf).sum(axis=(0, 1))
    else:
        mf = mask.astype(np.float64)[:, :, None, :]
        emit_num = (gamma[:, :, :, None] * obs[:, :, None, :] * mf).sum(axis=(0, 1))
        emit_den = (gamma[:, :, :, None] * mf).sum(axis=(0, 1))
    return dict(start=start_acc, trans=trans_acc, emit_num=emit_num, emit_den=emit_den, ll=ll)


def stats_vector(st: Dict[str, np.ndarray]) -> np.ndarray:
    """Flatten to the C-ABI layout [start | trans | emit_num | emit_den | ll]."""
    return np.concatenate([st["start"].ravel(), st["trans"].ravel(), st["emit_num"].ravel(),
                           st["emit_den"].ravel(), [st["ll"]]])


def mstep(p: Params, st, update_start=True, update_trans=True, update_emission=True) -> Params:
    """HMM.py:1598-1614 (binned :2322-2339): eps-smoothed closed forms, then the unconditional
    ``+eps`` re-normalisation of every parameter."""
    eps = p.eps
    q = p.copy()
    if update_start:
        s = st["start"] + eps
        q.start = s / s.sum()
    if update_trans:
        if q.trans_by_bin is not None:
            tb = st["trans"] + eps
            rs = tb.sum(axis=2, keepdims=True)
            rs[rs == 0.0] = 1.0
            q.trans_by_bin = tb / rs
        else:
            t = st["trans"][0] + eps
            rs = t.sum(axis=1, keepdims=True)
            rs[rs == 0.0] = 1.0
            q.trans = t / rs
    if update_emission:
        em = (st["emit_num"] + eps) / (st["emit_den"] + 2.0 * eps)
        q.emission = np.clip(em.reshape(q.emission.shape), eps, 1.0 - eps)
    return q.normalize()


def fit(X, p: Params, coords=None, max_iter=50, tol=1e-5, update_start=True, update_trans=True,
        update_emission=True) -> Tuple[Params, List[float]]:
    """EM driver with the relative-tolerance stop (HMM.py:1560-1630)."""
    hist: List[float] = []
    for _ in range(int(max_iter)):
        st = estep(X, p, coords)
        hist.append(st["ll"])
        p = mstep(p, st, update_start, update_trans, update_emission)
        if len(hist) > 1 and abs(hist[-1] - hist[-2]) < float(tol) * max(abs(hist[-1]), 1.0):
            break
    return p, hist


# --------------------------------------------------------------------------------------
# feature calling
# --------------------------------------------------------------------------------------
def runs_from_bool(row) -> List[Tuple[int, int]]:
    """HMM.py:927-938: maximal True runs as (start, end_exclusive)."""
    idx = np.nonzero(np.asarray(row).astype(bool))[0]
    if idx.size == 0:
        return []
    brk = np.nonzero(np.diff(idx) > 1)[0]
    starts = np.r_[idx[0], idx[brk + 1]]
    ends = np.r_[idx[brk] + 1, idx[-1] + 1]
    return [(int(s), int(e)) for s, e in zip(starts, ends)]


def lengths_for_binary(mat) -> np.ndarray:
    """HMM.py:947-959: every cell of a True run gets the run length (index units), int32."""
    mat = np.asarray(mat)
    out = np.zeros(mat.shape, dtype=np.int32)
    for i in range(mat.shape[0]):
        for s, e in runs_from_bool(mat[i]):
            out[i, s:e] = e - s
    return out


def pick_class(ranges: Sequence[Tuple[float, float]], length) -> int:
    """First class with lo <= length < hi (HMM.py:1335-1338); -1 if none."""
    for c, (lo, hi) in enumerate(ranges):
        if float(lo) <= float(length) < float(hi):
            return c
    return -1


def call_features(membership, coords, full_coords, ranges):
    """HMM.py:1325-1370 with span_fill and integer grids: per run -> genomic length -> size class
    -> fill [searchsorted_left(full, coords[s]), searchsorted_right(full, coords[e-1])) in the class
    layer and the all-layer; then run lengths of each binary layer.

    Returns (all_layer u8, all_len i32, class_layers u8 [n_cls,N,V], class_len i32 [n_cls,N,V])."""
    membership = np.asarray(membership).astype(bool)
    coords = np.asarray(coords, dtype=int)
    full_coords = np.asarray(full_coords, dtype=int)
    N, V = membership.shape[0], full_coords.shape[0]
    ncls = len(ranges)
    all_layer = np.zeros((N, V), dtype=np.uint8)
    cls_layers = np.zeros((ncls, N, V), dtype=np.uint8)
    for i in range(N):
        for s, e in runs_from_bool(membership[i]):
            glen = int(coords[e - 1]) - int(coords[s]) + 1 if e > s else 0
            if glen <= 0:
                continue
            c = pick_class(ranges, glen)
            if c < 0:
                continue
            left = int(np.searchsorted(full_coords, int(coords[s]), side="left"))
            right = int(np.searchsorted(full_coords, int(coords[e - 1]), side="right"))
            if left >= right:
                continue
            cls_layers[c, i, left:right] = 1
            all_layer[i, left:right] = 1
    all_len = lengths_for_binary(all_layer)
    cls_len = np.stack([lengths_for_binary(cls_layers[c]) for c in range(ncls)]) if ncls else np.zeros((0, N, V), np.int32)
    return all_layer, all_len, cls_layers, cls_len


def merge_runs(layer, coords, distance_threshold, coords_are_ints=True):
    """HMM.py:1040-1063: chain-merge 1-runs whose gap ``coords[s]-coords[prev_end-1]-1`` (or index
    gap) is <= threshold; returns (merged u8, lengths i32)."""
    arr = (np.asarray(layer) > 0).astype(np.uint8)
    coords = np.asarray(coords)
    out = np.zeros_like(arr)
    dt = int(distance_threshold)
    for i in range(arr.shape[0]):
        runs = runs_from_bool(arr[i] != 0)
        if not runs:
            continue
        ms, me = runs[0]
        merged = []
        for s, e in runs[1:]:
            gap = int(coords[s]) - int(coords[me - 1]) - 1 if coords_are_ints else s - me
            if gap <= dt:
                me = e
            else:
                merged.append((ms, me))
                ms, me = s, e
        merged.append((ms, me))
        for s, e in merged:
            out[i, s:e] = 1
    return out, lengths_for_binary(out)


def size_classes_from_binary(layer, coords, ranges, coords_are_ints=True):
    """HMM.py:1103-1119: runs of a binary layer -> first matching size class by bp length;
    returns (class_layers u8 [n_cls,N,V], class_len i32 [n_cls,N,V])."""
    arr = (np.asarray(layer) > 0).astype(np.uint8)
    coords = np.asarray(coords)
    n, V = arr.shape
    ncls = len(ranges)
    layers = np.zeros((ncls, n, V), dtype=np.uint8)
    for i in range(n):
        for s, e in runs_from_bool(arr[i] != 0):
            length = (int(coords[e - 1]) - int(coords[s]) + 1) if coords_are_ints else (e - s)
            c = pick_class(ranges, length)
            if c >= 0:
                layers[c, i, s:e] = 1
    lens = np.stack([lengths_for_binary(layers[c]) for c in range(ncls)]) if ncls else np.zeros((0, n, V), np.int32)
    return layers, lens


def span_valid(coords, starts, ends, covered=None):
    """HMM.py:193-226: True where a layer cell is kept (not NaN-masked)."""
    coords = np.asarray(coords)
    if covered is not None:
        return np.asarray(covered).astype(bool)
    starts = np.asarray(starts, dtype=float)
    ends = np.asarray(ends, dtype=float)
    valid = np.ones((starts.shape[0], coords.shape[0]), dtype=bool)
    for i in range(starts.shape[0]):
        if not np.isfinite(starts[i]) or not np.isfinite(ends[i]):
            continue
        valid[i] = ~((coords < int(starts[i])) | (coords > int(ends[i])))
    return valid


def merge_chain(layer, coords, distance_threshold, ranges, valid, coords_are_ints=True):
    """The per-layer body of ``_apply_merges`` (tools/partitioned_hmm.py:78-108): merge_runs ->
    size_classes_from_binary on the merged layer -> NaN outside ``valid`` (span_valid) on every layer.

    Returns float64 arrays in the reference's creation order:
    [merged, merged_lengths, class_0, class_0_lengths, class_1, ...]."""
    merged, mlen = merge_runs(layer, coords, distance_threshold, coords_are_ints)
    out = [merged, mlen]
    if len(ranges):
        cl, cll = size_classes_from_binary(merged, coords, ranges, coords_are_ints)
        for c in range(len(ranges)):
            out.extend([cl[c], cll[c]])
    res = []
    drop = ~np.asarray(valid, dtype=bool)
    for a in out:
        f = np.asarray(a).astype(np.float64)
        f[drop] = np.nan
        res.append(f)
    return res


# --------------------------------------------------------------------------------------
# synthetic SMF-like data -- shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------
STATE_EMISSION = (0.85, 0.08, 0.35, 0.60)
STATE_DWELL = (60.0, 150.0, 25.0, 80.0)


def synth_params(K: int, arch="single", n_channels=2, eps=1e-8, distance_bins=None) -> Params:
    """'Fitted-looking' parameters: sticky transitions from the dwell times, per-state emission."""
    trans = np.zeros((K, K))
    for i in range(K):
        leave = 1.0 / STATE_DWELL[i]
        trans[i] = leave / (K - 1)
        trans[i, i] = 1.0 - leave
    start = np.linspace(1.0, 2.0, K)
    start /= start.sum()
    em = np.array(STATE_EMISSION[:K])
    if arch == "multi":
        em = np.stack([np.clip(em * (1.0 - 0.15 * c) + 0.03 * c, 0.02, 0.98) for c in range(n_channels)], axis=1)
    p = Params(start, trans, em, eps=eps, distance_bins=distance_bins)
    if arch == "single_distance_binned":
        nb = len(p.distance_bins) + 1
        tb = []
        for b in range(nb):
            f = 1.0 + 0.6 * b  # farther sites -> more switching
            t = np.zeros((K, K))
            for i in range(K):
                leave = min(0.5, f / STATE_DWELL[i])
                t[i] = leave / (K - 1)
                t[i, i] = 1.0 - leave
            tb.append(t)
        p.trans_by_bin = np.stack(tb)
    return p.normalize()


def synth_reads(N: int, L: int, K: int, seed: int, prob=False, nan_rate=0.3, n_channels=1, truncate=True):
    """Reads from a K-state chain with the dwell times above; binary Bernoulli calls or
    Beta-distributed probabilities quantised to k/255; iid NaN plus read-span truncation.
    Returns float32 (N,L) or (N,L,C) with NaN = missing."""
    rng = np.random.default_rng(seed)
    leave = 1.0 / np.array(STATE_DWELL[:K])
    s = np.empty((N, L), dtype=np.int64)
    s[:, 0] = rng.integers(0, K, size=N)
    u = rng.random((N, L))
    jump = rng.integers(1, K, size=(N, L))
    for t in range(1, L):
        sw = u[:, t] < leave[s[:, t - 1]]
        s[:, t] = np.where(sw, (s[:, t - 1] + jump[:, t]) % K, s[:, t - 1])
    em = np.array(STATE_EMISSION[:K])
    chans = []
    for c in range(n_channels):
        pc = np.clip(em * (1.0 - 0.15 * c) + 0.03 * c, 0.02, 0.98)[s]
        if prob:
            a = 6.0
            x = rng.beta(pc * a + 0.3, (1.0 - pc) * a + 0.3)
            x = np.round(x * 255.0) / 255.0
        else:
            x = (rng.random((N, L)) < pc).astype(np.float64)
        x = x.astype(np.float32)
        x[rng.random((N, L)) < nan_rate] = np.nan
        chans.append(x)
    X = chans[0] if n_channels == 1 else np.stack(chans, axis=2)
    if truncate and L >= 8:
        lo = rng.integers(0, max(1, L // 10), size=N)
        hi = L - rng.integers(0, max(1, L // 10), size=N)
        pos = np.arange(L)[None, :]
        outside = (pos < lo[:, None]) | (pos >= hi[:, None])
        X[outside] = np.nan
    return X


def synth_coords(L: int, seed: int, span_factor: int = 4) -> np.ndarray:
    """Sorted random sample of L sites from a span of ~span_factor*L bp."""
    rng = np.random.default_rng(seed + 7919)
    return np.sort(rng.choice(np.arange(span_factor * L), size=L, replace=False)).astype(np.int64)
