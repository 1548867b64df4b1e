"""Chunked host <-> device pipelines behind the reference-signature calls.

``decode`` (HMM.py:673-720) takes a host matrix and returns host arrays; at B200 kernel rates the call
is bound by the host link, so the work is cut into read chunks and three things overlap:

* host threads copy chunk i+1 of a *pageable* input into a page-locked staging ring (skipped when the
  caller's array is already page-locked or lives on the device), the copy engine uploads it;
* the kernels of chunk i run (posterior, optional Viterbi, widening to the caller's dtypes);
* the copy engine downloads chunk i-1 straight into the page-locked result arrays.

All copies are explicit ``cudaMemcpyAsync`` calls through the C-ABI (``smf_hmm_copy_async``).
"""
from __future__ import annotations

import threading
from typing import List, Optional

import numpy as np
import torch

from . import engine, hostmem


_side_streams: dict = {}
_side_lock = threading.Lock()


def side_stream(device: torch.device, role: str) -> torch.cuda.Stream:
    """Per-(device, role) copy stream, created once per process: a decode call of a small subset (the reference's
    regime: ~1000 reads x ~170 sites) should not pay two cudaStreamCreate per call.  Ordering between calls is kept by
    the entry / completion events every call records anyway."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), role)
    with _side_lock:
        st = _side_streams.get(key)
        if st is None:
            st = torch.cuda.Stream(device)
            _side_streams[key] = st
        return st


class ChunkUploader:
    """Uploads row blocks ``src[lo:hi]`` of a host / device matrix into device memory, one chunk ahead.

    ``src`` may be a numpy array, a CPU tensor or a CUDA tensor (then nothing is copied).  Pageable host
    rows go through a two-slot page-locked staging ring filled by the host thread pool.
    """

    def __init__(self, src, rows: int, device: torch.device, nbuf: int = 2):
        self.device = device
        self.rows = int(rows)
        self.nbuf = nbuf
        if isinstance(src, torch.Tensor) and src.is_cuda:
            self.dev_src = src
            self.host = None
        else:
            self.dev_src = None
            self.host = src.detach().numpy() if isinstance(src, torch.Tensor) else np.asarray(src)
        self.tail = tuple(int(s) for s in (src.shape[1:]))
        self.stream = side_stream(device, "h2d")
        self.ev_in: List[Optional[torch.cuda.Event]] = [None] * nbuf
        self.d_buf: List[Optional[torch.Tensor]] = [None] * nbuf
        self.stage: List[Optional[np.ndarray]] = [None] * nbuf
        self.direct = False
        if self.host is not None:
            tdt = torch.from_numpy(np.empty(0, dtype=self.host.dtype)).dtype
            self.d_buf = [torch.empty((self.rows,) + self.tail, dtype=tdt, device=device) for _ in range(nbuf)]
            row_contig = self.host.ndim >= 1 and self.host[:1].flags.c_contiguous and (
                self.host.shape[0] <= 1 or self.host.strides[0] == int(np.prod(self.tail, dtype=np.int64)) * self.host.itemsize)
            self.direct = bool(row_contig and hostmem.is_pinned(self.host))
            if not self.direct:
                self.stage = [hostmem.pool().empty((self.rows,) + self.tail, self.host.dtype) for _ in range(nbuf)]

    def wait_entry(self, event: torch.cuda.Event) -> None:
        self.stream.wait_event(event)

    def upload(self, i: int, lo: int, hi: int, reuse_after: Optional[torch.cuda.Event]) -> torch.Tensor:
        """Device view of rows [lo, hi).  ``reuse_after``: event after which slot ``i % nbuf`` of the device
        ring is no longer read by earlier kernels."""
        n = hi - lo
        if self.dev_src is not None:
            return self.dev_src[lo:hi].contiguous()
        b = i % self.nbuf
        if self.direct:
            src_rows = self.host[lo:hi]
        else:
            if self.ev_in[b] is not None:
                self.ev_in[b].synchronize()  # the upload that last read this staging slot has finished
            hostmem.parallel_copy(self.stage[b][:n], self.host[lo:hi])
            src_rows = self.stage[b][:n]
        if reuse_after is not None:
            self.stream.wait_event(reuse_after)
        engine.copy_to_device_async(self.d_buf[b][:n], src_rows, self.stream)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.ev_in[b] = ev
        torch.cuda.current_stream(self.device).wait_event(ev)
        return self.d_buf[b][:n]

    def finish(self) -> None:
        self.stream.synchronize()


def decode_stream(model, src, numeric_u8: bool, tables, bins, use_viterbi: bool, states_out: np.ndarray,
                  gamma_out: np.ndarray, device: torch.device, chunk_out_bytes: int = 96 << 20) -> None:
    """Fill ``states_out`` (N, L) and ``gamma_out`` (N, L, K) -- any of int8 / int64 and float32 / float64
    (other posterior dtypes are converted with a torch cast) -- from observations ``src``."""
    N, L = int(src.shape[0]), int(src.shape[1])
    K = model.n_states
    if not (states_out.flags.c_contiguous and gamma_out.flags.c_contiguous):
        raise ValueError("out buffers must be C-contiguous")
    g_np = gamma_out.dtype
    s_np = states_out.dtype
    per_read_out = L * (K * g_np.itemsize + s_np.itemsize)
    rows = max(1, int(chunk_out_bytes // max(per_read_out, 1)))
    rows = min(rows, N)
    cur = torch.cuda.current_stream(device)
    s_out = side_stream(device, "d2h")
    entry = torch.cuda.Event()
    entry.record(cur)
    s_out.wait_event(entry)
    up = ChunkUploader(src, rows, device)
    up.wait_entry(entry)
    nbuf = 2
    d_g = [torch.empty((rows, L, K), dtype=torch.float32, device=device) for _ in range(nbuf)]
    d_s = [torch.empty((rows, L), dtype=torch.int8, device=device) for _ in range(nbuf)]
    g_t = torch.from_numpy(np.empty(0, dtype=g_np)).dtype
    s_t = torch.from_numpy(np.empty(0, dtype=s_np)).dtype
    d_gw = [torch.empty((rows, L, K), dtype=g_t, device=device) for _ in range(nbuf)] if g_t != torch.float32 else None
    d_sw = [torch.empty((rows, L), dtype=s_t, device=device) for _ in range(nbuf)] if s_t != torch.int8 else None
    ws = torch.empty((rows * L,), dtype=torch.uint8, device=device) if use_viterbi else None
    ev_cmp: List[Optional[torch.cuda.Event]] = [None] * nbuf
    ev_out: List[Optional[torch.cuda.Event]] = [None] * nbuf
    kernel_widen = (g_t in (torch.float32, torch.float64)) and (s_t in (torch.int8, torch.int64))
    with hostmem.registered(states_out, gamma_out):
        try:
            for i, lo in enumerate(range(0, N, rows)):
                hi = min(N, lo + rows)
                n = hi - lo
                b = i % nbuf
                obs = up.upload(i, lo, hi, ev_cmp[b])
                if numeric_u8 and obs.dtype == torch.uint8:
                    obs = obs.to(torch.float32)  # numeric uint8 input: the value itself is the observation
                if ev_out[b] is not None:
                    cur.wait_event(ev_out[b])
                engine.posterior(tables, obs, bins, want_gamma=True, want_states=not use_viterbi,
                                 out_gamma=d_g[b][:n], out_states=None if use_viterbi else d_s[b][:n])
                if use_viterbi:
                    engine.viterbi(tables, obs, bins, out_states=d_s[b][:n], workspace=ws)
                g_src, s_src = d_g[b][:n], d_s[b][:n]
                if kernel_widen:
                    if d_gw is not None or d_sw is not None:
                        engine.widen_decode(s_src if d_sw is not None else None, d_sw[b][:n] if d_sw is not None else None,
                                            g_src if d_gw is not None else None, d_gw[b][:n] if d_gw is not None else None)
                else:
                    if d_gw is not None:
                        d_gw[b][:n].copy_(g_src)
                    if d_sw is not None:
                        d_sw[b][:n].copy_(s_src)
                if d_gw is not None:
                    g_src = d_gw[b][:n]
                if d_sw is not None:
                    s_src = d_sw[b][:n]
                ev_cmp[b] = torch.cuda.Event()
                ev_cmp[b].record(cur)
                s_out.wait_event(ev_cmp[b])
                engine.copy_to_host_async(gamma_out[lo:hi], g_src, s_out)
                engine.copy_to_host_async(states_out[lo:hi], s_src, s_out)
                ev_out[b] = torch.cuda.Event()
                ev_out[b].record(s_out)
        finally:
            # also on errors: nothing may still be in flight when the staging buffers go back to their pools
            up.finish()
            s_out.synchronize()
            cur.synchronize()
