"""Multi-GPU plumbing: reads shard across ranks, only the Baum-Welch statistics are exchanged.

Decode / Viterbi / feature calling are independent per read: every rank processes a contiguous
block of reads and no collective is needed.  Fitting keeps each rank's block resident on its GPU;
per EM iteration the fp64 statistics vector (K + n_bins*K*K + 2*K*C + 1 doubles) is
sum-all-reduced with NCCL and every rank runs the same deterministic M-step, so all ranks hold
bit-identical parameters.
"""
from __future__ import annotations

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
