"""Host <-> device link probe: what bounds the end-to-end (host buffers in, host buffers out) numbers.

    python profiles/host_link_probe.py [--gb 2] [--seconds 3]

Prints one JSON object: pinned / pageable H2D and D2H rates, concurrent bidirectional rate, the cost of
cudaHostAlloc / cudaHostRegister, first-touch and memcpy rates of the host cores.  Under torchrun every
rank runs the pinned bidirectional probe at the same time (the per-box host-link ceiling for N GPUs).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


def _rate(nbytes, secs):
    return round(nbytes / secs / 1e9, 2)


def timed(fn, sync=True):
    if sync:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    if sync:
        torch.cuda.synchronize()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--seconds", type=float, default=3.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    n = int(args.gb * (1 << 30))
    out = {"rank": rank, "world": world, "bytes": n, "cpus": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
    rt = ctypes.CDLL("libcudart.so.12")

    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    if rank == 0:
        # --- allocation costs ---
        t = timed(lambda: torch.empty(n, dtype=torch.uint8, pin_memory=True))
        out["pinned_alloc_cold_gbs"] = _rate(n, t)
        t0 = time.perf_counter()
        pg = np.empty(n, dtype=np.uint8)
        pg.fill(1)
        out["pageable_first_touch_1thr_gbs"] = _rate(n, time.perf_counter() - t0)
        for thr in (8, 16, 32):
            arr = np.empty(n, dtype=np.uint8)
            step = (n + thr - 1) // thr
            t0 = time.perf_counter()
            with ThreadPoolExecutor(thr) as pool:
                list(pool.map(lambda lo: arr[lo:lo + step].fill(0), range(0, n, step)))
            out[f"pageable_first_touch_{thr}thr_gbs"] = _rate(n, time.perf_counter() - t0)
            del arr
        # --- cudaHostRegister on touched pageable memory ---
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(ctypes.c_void_p(pg.ctypes.data), ctypes.c_size_t(n), ctypes.c_uint(0))
        out["host_register_rc"] = rc
        out["host_register_gbs"] = _rate(n, time.perf_counter() - t0)
        if rc == 0:
            t_pg = torch.from_numpy(pg)
            t = min(timed(lambda: d_a.copy_(t_pg, non_blocking=True)) for _ in range(3))
            out["h2d_registered_gbs"] = _rate(n, t)
            t = min(timed(lambda: t_pg.copy_(d_a, non_blocking=True)) for _ in range(3))
            out["d2h_registered_gbs"] = _rate(n, t)
            t0 = time.perf_counter()
            rt.cudaHostUnregister(ctypes.c_void_p(pg.ctypes.data))
            out["host_unregister_gbs"] = _rate(n, time.perf_counter() - t0)
        # register untouched memory (fresh np.empty)
        fresh = np.empty(n, dtype=np.uint8)
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(ctypes.c_void_p(fresh.ctypes.data), ctypes.c_size_t(n), ctypes.c_uint(0))
        out["host_register_fresh_gbs"] = _rate(n, time.perf_counter() - t0)
        if rc == 0:
            rt.cudaHostUnregister(ctypes.c_void_p(fresh.ctypes.data))
        del fresh
        # --- pageable copies through torch / the driver's bounce buffer ---
        t_pg = torch.from_numpy(pg)
        t = min(timed(lambda: d_a.copy_(t_pg)) for _ in range(2))
        out["h2d_pageable_gbs"] = _rate(n, t)
        t = min(timed(lambda: t_pg.copy_(d_a)) for _ in range(2))
        out["d2h_pageable_gbs"] = _rate(n, t)
        # --- host memcpy rates (pinned -> touched pageable), GIL released inside numpy copies ---
        pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        pin_np = pin.numpy()
        for thr in (1, 4, 8, 16, 32):
            step = (n + thr - 1) // thr
            t0 = time.perf_counter()
            with ThreadPoolExecutor(thr) as pool:
                list(pool.map(lambda lo: np.copyto(pg[lo:lo + step], pin_np[lo:lo + step]), range(0, n, step)))
            out[f"host_memcpy_{thr}thr_gbs"] = _rate(n, time.perf_counter() - t0)
        # widening on the host: int8 -> int64, f32 -> f64 (what a reference-signature decode needs)
        src8 = pin_np[: n // 8].view(np.int8)
        dst64 = np.empty(n // 8, dtype=np.int64)
        dst64.fill(0)
        for thr in (1, 8, 32):
            m = src8.shape[0]
            step = (m + thr - 1) // thr
            t0 = time.perf_counter()
            with ThreadPoolExecutor(thr) as pool:
                list(pool.map(lambda lo: np.copyto(dst64[lo:lo + step], src8[lo:lo + step], casting="unsafe"),
                              range(0, m, step)))
            out[f"host_widen_i8_i64_{thr}thr_out_gbs"] = _rate(m * 8, time.perf_counter() - t0)
        del pg, t_pg, dst64, src8, pin_np, pin

    # --- pinned link rates, every rank at once ---
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_a.fill_(1)
    h_b.fill_(2)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def loop(kind):
        barrier()
        reps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < args.seconds:
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
            reps += 1
        dt = time.perf_counter() - t0
        per_dir = n * reps / dt / 1e9
        t = torch.tensor([per_dir], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return round(float(t.item()), 2)

    out["pinned_h2d_gbs_all_ranks"] = loop("h2d")
    out["pinned_d2h_gbs_all_ranks"] = loop("d2h")
    out["pinned_bidir_gbs_per_direction_all_ranks"] = loop("both")
    if rank == 0:
        try:
            import subprocess

            out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True,
                                         timeout=20).stdout[-3000:]
            out["numa_nodes"] = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
        except Exception as exc:  # noqa: BLE001
            out["topo"] = f"unavailable: {exc}"
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
