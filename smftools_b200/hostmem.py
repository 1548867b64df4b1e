"""Page-locked host memory for results and staging: a small caching pool + parallel host copies.

The reference's contract is host numpy arrays in, fresh host numpy arrays out (``decode`` HMM.py:673-720,
``annotate_adata`` layers HMM.py:1256-1377).  A copy into *pageable* memory runs through the driver's
bounce buffer at 11-17 GB/s on the B200 boxes, a copy into page-locked memory at 53-55 GB/s
(``profiles/host_link_probe.py``).  So result arrays are carved from page-locked blocks:

* a block is ordinary page-aligned numpy memory whose pages are first touched by a few threads
  (50 GB/s with 16 threads vs 5.7 GB/s with one) and then registered with ``smf_hmm_host_register``
  (cudaHostRegister, 49 GB/s on touched pages; cudaHostAlloc would pin at 2.5 GB/s);
* the array handed to the caller is a plain ``numpy.ndarray`` view of the block; when its last view is
  garbage-collected the block returns to the pool, still registered, and the next call of the same size
  reuses it at no cost (bounded by ``SMF_PINNED_POOL_BYTES``, default 25 % of host memory, at most 64 GiB;
  blocks beyond the bound are unregistered and freed).

Pageable *inputs* are copied into a page-locked staging ring by the same thread pool
(``numpy.copyto`` releases the GIL; 77 GB/s with 8 threads) while the previous chunk is in flight.
If page-locking is refused the arrays are still valid, copies just run at pageable speed.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

_PAGE = 4096
_GRAIN = 2 << 20  # block capacities are multiples of 2 MiB


def _host_threads() -> int:
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    env = os.environ.get("SMF_HOST_THREADS")
    if env:
        return max(1, int(env))
    return max(1, min(16, cpus // local_world))


_workers: Optional[ThreadPoolExecutor] = None
_workers_n = 0
_workers_lock = threading.Lock()


def workers() -> Tuple[ThreadPoolExecutor, int]:
    """Process-wide pool of host copy threads."""
    global _workers, _workers_n
    with _workers_lock:
        if _workers is None:
            _workers_n = _host_threads()
            _workers = ThreadPoolExecutor(max_workers=_workers_n, thread_name_prefix="smf-host")
        return _workers, _workers_n


def _split(n: int, parts: int, grain: int = 1) -> List[Tuple[int, int]]:
    parts = max(1, min(parts, (n + grain - 1) // max(grain, 1)))
    step = (n + parts - 1) // parts
    return [(lo, min(n, lo + step)) for lo in range(0, n, step)]


def parallel_copy(dst: np.ndarray, src: np.ndarray) -> None:
    """``dst[...] = src`` (same shape; dtype conversion allowed) split over the host threads along axis 0."""
    if dst.shape != src.shape:
        raise ValueError(f"parallel_copy: shape mismatch {dst.shape} vs {src.shape}")
    n = dst.shape[0] if dst.ndim else 0
    if dst.nbytes < (8 << 20) or n < 2:
        np.copyto(dst, src, casting="unsafe")
        return
    pool, nthr = workers()
    if nthr < 2:
        np.copyto(dst, src, casting="unsafe")
        return
    jobs = _split(n, nthr)
    list(pool.map(lambda b: np.copyto(dst[b[0]:b[1]], src[b[0]:b[1]], casting="unsafe"), jobs))


def parallel_touch(raw: np.ndarray) -> None:
    """First-touch the pages of a fresh uint8 buffer from several threads."""
    n = raw.shape[0]
    pool, nthr = workers()
    if n < (64 << 20) or nthr < 2:
        raw.fill(0)
        return
    list(pool.map(lambda b: raw[b[0]:b[1]].fill(0), _split(n, nthr, _PAGE)))


def is_pinned(arr: np.ndarray) -> bool:
    if arr.nbytes == 0:
        return False
    return bool(_lib.load().smf_hmm_host_is_pinned(ctypes.c_void_p(arr.ctypes.data)))


class _Block:
    __slots__ = ("raw", "addr", "cap", "registered")

    def __init__(self, raw: np.ndarray, addr: int, cap: int, registered: bool):
        self.raw, self.addr, self.cap, self.registered = raw, addr, cap, registered


class HostPool:
    """Caching pool of page-locked host blocks (see the module docstring)."""

    def __init__(self, max_cached_bytes: Optional[int] = None):
        if max_cached_bytes is None:
            env = os.environ.get("SMF_PINNED_POOL_BYTES")
            if env is not None:
                max_cached_bytes = int(env)
            else:
                try:
                    total = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
                except (ValueError, OSError):
                    total = 16 << 30
                max_cached_bytes = min(64 << 30, total // 4)
        self.max_cached = int(max_cached_bytes)
        self._free: List[_Block] = []
        self._cached = 0
        self._lock = threading.Lock()
        self.stats = {"allocs": 0, "hits": 0, "evictions": 0, "register_failed": 0}

    # -- block management -------------------------------------------------------------------------
    def _new_block(self, cap: int) -> _Block:
        raw = np.empty(cap + _PAGE, dtype=np.uint8)
        off = (-raw.ctypes.data) % _PAGE
        view = raw[off:off + cap]
        parallel_touch(view)
        addr = view.ctypes.data
        rc = _lib.load().smf_hmm_host_register(ctypes.c_void_p(addr), ctypes.c_size_t(cap))
        if rc != _lib.SMF_OK:
            self.stats["register_failed"] += 1
        self.stats["allocs"] += 1
        return _Block(raw, addr, cap, rc == _lib.SMF_OK)

    def _drop(self, blk: _Block) -> None:
        if blk.registered:
            try:
                _lib.load().smf_hmm_host_unregister(ctypes.c_void_p(blk.addr))
            except Exception:  # interpreter shutdown
                pass
            blk.registered = False
        blk.raw = None

    def _take(self, nbytes: int) -> _Block:
        cap = max(_GRAIN, (nbytes + _GRAIN - 1) // _GRAIN * _GRAIN)
        with self._lock:
            best = None
            for i, blk in enumerate(self._free):
                if blk.cap >= cap and blk.cap <= cap + cap // 4 + _GRAIN and (best is None or blk.cap < self._free[best].cap):
                    best = i
            if best is not None:
                blk = self._free.pop(best)
                self._cached -= blk.cap
                self.stats["hits"] += 1
                return blk
        return self._new_block(cap)

    def _give_back(self, blk: _Block) -> None:
        evict: List[_Block] = []
        with self._lock:
            if blk.registered and blk.cap <= self.max_cached:
                self._free.append(blk)
                self._cached += blk.cap
                while self._cached > self.max_cached and self._free:
                    old = self._free.pop(0)
                    self._cached -= old.cap
                    evict.append(old)
            else:
                evict.append(blk)
        for old in evict:
            self.stats["evictions"] += 1
            self._drop(old)

    def trim(self) -> None:
        """Unregister and free every cached block."""
        with self._lock:
            blocks, self._free, self._cached = self._free, [], 0
        for blk in blocks:
            self._drop(blk)

    @property
    def cached_bytes(self) -> int:
        return self._cached

    # -- arrays -----------------------------------------------------------------------------------
    def empty(self, shape: Sequence[int], dtype) -> np.ndarray:
        """Uninitialised page-locked array; its block returns to the pool when the last view dies."""
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype=dtype)
        blk = self._take(nbytes)
        buf = (ctypes.c_ubyte * nbytes).from_address(blk.addr)
        weakref.finalize(buf, self._give_back, blk).atexit = False
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def empty_many(self, specs: Dict[str, Tuple[Sequence[int], object]]) -> Dict[str, np.ndarray]:
        return {nm: self.empty(shape, dt) for nm, (shape, dt) in specs.items()}


_pool: Optional[HostPool] = None
_pool_lock = threading.Lock()


def pool() -> HostPool:
    global _pool
    with _pool_lock:
        if _pool is None:
            _pool = HostPool()
        return _pool


class registered:
    """Context manager: page-lock a caller-owned array in place for the duration of a call (no-op for
    small arrays and for memory that is already page-locked)."""

    def __init__(self, *arrays: np.ndarray, min_bytes: int = 32 << 20):
        self.arrays = [a for a in arrays if a is not None]
        self.min_bytes = min_bytes
        self.done: List[int] = []

    def __enter__(self):
        lib = _lib.load()
        for a in self.arrays:
            if a.nbytes < self.min_bytes or not a.flags.c_contiguous or is_pinned(a):
                continue
            if lib.smf_hmm_host_register(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes)) == _lib.SMF_OK:
                self.done.append(a.ctypes.data)
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        for addr in self.done:
            lib.smf_hmm_host_unregister(ctypes.c_void_p(addr))
        self.done = []
        return False
