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


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (Linux sysfs), so that pinned /
    pageable host buffers allocated afterwards are first-touched on that node and host<->device copies
    of several ranks do not all cross the same socket link.  Returns the node, or None when the
    topology cannot be read (then nothing is changed)."""
    try:
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id  # e.g. "0000:1B:00.0"
    except Exception:
        try:
            import ctypes

            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL("libcudart.so")
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device_index)) != 0:
                return None
            bdf = buf.value.decode()
        except Exception:
            return None
    try:
        with open(f"/sys/bus/pci/devices/{str(bdf).lower()}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None
