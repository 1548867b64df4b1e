"""CPU tests of the host logic of the grouped launches: dealing groups to ranks (largest first), the packed
parameter layout, the world_size-2 gloo parameter exchange, and argument checks that need no device."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deal_groups_is_lpt_and_covers_every_group_once():
    from smftools_b200.grouped import deal_groups

    costs = [5, 9, 1, 7, 3, 3, 8, 2]
    owned = deal_groups(costs, 3)
    assert sorted(g for o in owned for g in o) == list(range(len(costs)))
    loads = [sum(costs[g] for g in o) for o in owned]
    # LPT: 9 | 8 | 7, then 5 -> rank 2 (12), 3 -> rank 1 (11), 3 -> rank 0 (12), 2 -> rank 1 (13), 1 -> rank 0 (tie: lower rank)
    assert owned == [[1, 2, 5], [4, 6, 7], [0, 3]] and loads == [13, 13, 12]
    assert max(loads) <= sum(costs) / 3 + max(costs)
    assert deal_groups(costs, 1) == [list(range(len(costs)))]
    assert deal_groups([], 2) == [[], []]
    assert deal_groups([4, 4], 4) == [[0], [1], [], []]
    # deterministic under ties
    assert deal_groups([2, 2, 2, 2], 2) == [[0, 2], [1, 3]]


def test_partition_reads_covers_every_read_once_and_balances():
    from smftools_b200.grouped import partition_reads

    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        n_reads = [int(x) for x in rng.integers(1, 5000, size=16)]
        n_pos = [int(x) for x in rng.integers(2000, 20000, size=16)]
        parts = partition_reads(n_reads, n_pos, world)
        assert len(parts) == world
        seen = {g: [] for g in range(16)}
        loads = []
        for r in parts:
            loads.append(sum((hi - lo) * n_pos[g] for g, lo, hi in r))
            assert [g for g, _, _ in r] == sorted(g for g, _, _ in r)  # group order, contiguous on the cost axis
            for g, lo, hi in r:
                assert 0 <= lo < hi <= n_reads[g]
                seen[g].append((lo, hi))
        for g, ranges in seen.items():
            ranges.sort()
            assert ranges[0][0] == 0 and ranges[-1][1] == n_reads[g]
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))  # disjoint, no gaps
        total = sum(n * l for n, l in zip(n_reads, n_pos))
        assert sum(loads) == total
        assert max(loads) - min(loads) <= 2 * max(n_pos)  # within one read per cut
    assert partition_reads([10, 4, 7, 1], [100, 50, 200, 1000], 3) == [[(0, 0, 10), (1, 0, 4)], [(2, 0, 6)],
                                                                       [(2, 6, 7), (3, 0, 1)]]
    assert partition_reads([3], [7], 4) == [[(0, 0, 1)], [(0, 1, 2)], [(0, 2, 3)], []]


def test_pack_unpack_round_trip_keeps_dtype():
    import smftools_b200 as S
    from smftools_b200.grouped import pack_parameters, unpack_parameters

    models = [S.create_hmm({"hmm_n_states": 3, "hmm_init_emission_probs": [0.7, 0.4, 0.1 + 0.01 * g]}, arch="single")
              for g in range(3)]
    P = pack_parameters(models)
    assert P.shape == (3, 3 + 9 + 3) and P.dtype == np.float64
    np.testing.assert_array_equal(P[2, 12:], models[2].emission.detach().numpy().reshape(-1))
    Q = P.copy()
    Q[:, 12:] = 0.25
    unpack_parameters(models, Q)
    for m in models:
        assert m.emission.dtype == torch.float64 and tuple(m.emission.shape) == (3,)
        assert torch.all(m.emission == 0.25)
        np.testing.assert_array_equal(m.trans.detach().numpy(), P[0, 3:12].reshape(3, 3))


def test_grouped_workspace_bytes_validates_the_group_table():
    from smftools_b200 import _lib

    lib = _lib.load()
    arr = (_lib.Group * 2)()
    for g, (n, L, stride) in enumerate([(1000, 170, 176), (50, 2000, 2000)]):
        arr[g].obs = 4096
        arr[g].n_reads = n
        arr[g].n_pos = L
        arr[g].row_stride = stride
    fit = lib.smf_hmm_grouped_workspace_bytes(4, arr, 2, 1)
    post = lib.smf_hmm_grouped_workspace_bytes(4, arr, 2, 0)
    # checkpoints: ceil(L / 16) x N x 16 B per group (K = 4) + per-read log-likelihoods; the fit adds per-item partials
    ck = (11 * 1000 + 125 * 50) * 16
    assert post >= ck + 1050 * 8 and fit > post
    assert lib.smf_hmm_grouped_workspace_bytes(2, arr, 2, 1) < fit  # K = 2 keeps 2 floats per checkpoint
    arr[1].row_stride = 1999  # shorter than the read
    assert lib.smf_hmm_grouped_workspace_bytes(4, arr, 2, 1) == 0
    arr[1].row_stride = 2000
    arr[0].n_reads = -1
    assert lib.smf_hmm_grouped_workspace_bytes(4, arr, 2, 1) == 0
    arr[0].n_reads = 0  # an empty shard of a read-sharded group is legal
    assert 0 < lib.smf_hmm_grouped_workspace_bytes(4, arr, 2, 1) < fit
    assert lib.smf_hmm_grouped_workspace_bytes(5, arr, 2, 1) == 0
    assert lib.smf_hmm_grouped_workspace_bytes(4, None, 0, 1) == 0
    # null outputs / workspace are rejected before anything touches the device
    arr[0].n_reads = 10
    assert lib.smf_hmm_grouped_fit(4, _lib.OBS_U8, arr, 2, None, 5, 0.0, 1e-8, 7, None, None, None, None, 0, None) == _lib.SMF_ERR_BAD_ARG
    assert ctypes.sizeof(_lib.Group) == 56


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import smftools_b200 as S
    from smftools_b200.grouped import deal_groups, exchange_group_parameters, pack_parameters

    costs = [300, 1200, 50, 700, 700]
    owned = deal_groups(costs, world)
    models = [S.create_hmm({"hmm_n_states": 2}, arch="single") for _ in costs]
    # stand-in for the device fit of this rank's groups: a rank-independent function of the group index
    for g in owned[rank]:
        with torch.no_grad():
            models[g].emission.data = torch.tensor([0.9 - 0.01 * g, 0.1 + 0.01 * g], dtype=torch.float64)
    exchange_group_parameters(models, owned)
    q.put((rank, owned, pack_parameters(models)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_group_parameter_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, owned0, P0), (_, owned1, P1) = res
    assert owned0 == owned1 == [[0, 1], [2, 3, 4]]  # 1200 + 300 | 700 + 700 + 50
    np.testing.assert_array_equal(P0, P1)  # every rank ends with every group's fitted parameters
    for g in range(5):
        np.testing.assert_allclose(P0[g, -2:], [0.9 - 0.01 * g, 0.1 + 0.01 * g], rtol=0, atol=1e-15)
