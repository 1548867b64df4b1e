"""Multi-GPU plumbing: reads shard across ranks, only the Baum-Welch statistics are exchanged.

Decode / Viterbi / feature calling are independent per read: every rank processes a contiguous
block of reads and no collective is needed.  Fitting keeps each rank's block resident on its GPU;
per EM iteration the fp64 statistics vector (K + n_bins*K*K + 2*K*C + 1 doubles) is
sum-all-reduced with NCCL and every rank runs the same deterministic M-step, so all ranks hold
bit-identical parameters.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch


def shard_bounds(n_reads: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of reads owned by ``rank``; sizes differ by at most one."""
    base, extra = divmod(int(n_reads), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def enable_distributed_fit(model, process_group=None):
    """Make ``model.fit`` all-reduce its E-step statistics over ``process_group`` (default group
    when None).  Every rank must call ``fit`` with its own shard and the same hyper-parameters."""
    model.distributed = True
    model.process_group = process_group
    return model


def broadcast_parameters(model, src: int = 0, process_group=None):
    """Optional safety net: make every rank start from rank ``src``'s parameters."""
    import torch.distributed as dist

    for p in model.parameters():
        t = p.data
        buf = t if t.is_cuda or dist.get_backend(process_group) != "nccl" else t.cuda()
        dist.broadcast(buf, src=src, group=process_group)
        if buf is not t:
            p.data = buf.to(t.device)
    return model


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def gpu_cpu_affinity(device_index: int) -> Optional[set]:
    """CPUs the driver reports as local to the GPU (NVML ``nvmlDeviceGetCpuAffinity``, the same source as
    ``nvidia-smi topo -m``); None when NVML is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            uuid = torch.cuda.get_device_properties(device_index).uuid
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{uuid}".encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            return cpus or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[dict]:
    """Pin this process to the CPUs local to its GPU, so that host buffers allocated afterwards are
    first-touched on that NUMA node and the host<->device copies of several ranks do not all cross the
    same socket link.  Sources, in order: the PCI device's ``numa_node`` in sysfs, then NVML's CPU affinity
    of the GPU.  Returns ``{"source", "node", "cpus", "changed"}`` (``changed`` False when the affinity
    already equals the allowed set, e.g. on a single-node VM), or None when no topology is exposed."""
    allowed_now = set(os.sched_getaffinity(0))
    info = None
    try:
        bdf = str(torch.cuda.get_device_properties(device_index).pci_bus_id).lower()
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
                info = {"source": "sysfs", "node": node, "cpus": _parse_cpulist(fh.read())}
    except (OSError, ValueError, AttributeError, RuntimeError):
        info = None
    if info is None:
        cpus = gpu_cpu_affinity(device_index)
        if cpus:
            info = {"source": "nvml", "node": None, "cpus": cpus}
    if info is None:
        return None
    target = info["cpus"] & allowed_now
    if not target:
        return None
    changed = target != allowed_now
    if changed:
        try:
            os.sched_setaffinity(0, target)
        except OSError:
            return None
    return {"source": info["source"], "node": info["node"], "cpus": len(target), "changed": changed}
