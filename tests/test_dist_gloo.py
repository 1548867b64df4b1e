"""world_size-2 gloo test (CPU) of the multi-GPU fit path: reads are sharded, each rank's
E-step statistics are all-reduced, the replicated M-step gives every rank the parameters a
single process obtains on the full data.  The CUDA E-step kernel cannot run here, so the ranks'
local statistics come from the oracle (test infrastructure); everything else -- log tables,
all-reduce, M-step, convergence test, history -- is the product code."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_estep_patch(model, X_local, coords):
    from oracle import hmm_oracle as O
    from helpers import params_of_model

    def _estep_local(tables, resident, bins, workspace, device):
        p = params_of_model(model)
        return torch.tensor(O.stats_vector(O.estep(X_local, p, coords)), dtype=torch.float64)

    model._estep_local = _estep_local
    model._ensure_device_dtype = lambda device=None: torch.device("cpu")
    model._bins_for = lambda coords, L, device: None
    model._resident_chunks = lambda src, device, numeric_u8=False: [src]
    import smftools_b200.hmm as H

    H.engine.estep_workspace = lambda tables, device, n_reads=1, n_pos=1: torch.zeros(1, dtype=torch.float64)


def _worker(rank, world, port, arch, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import smftools_b200 as S
    from smftools_b200.dist import enable_distributed_fit, shard_bounds
    from oracle import hmm_oracle as O

    K = 3
    X = O.synth_reads(41, 60, K, 77, prob=True).astype(np.float64)
    coords = O.synth_coords(60, 3) if arch == "single_distance_binned" else np.arange(60)
    lo, hi = shard_bounds(X.shape[0], rank, world)
    m = S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": [0.8, 0.5, 0.1]}, arch=arch)
    enable_distributed_fit(m)
    _oracle_estep_patch(m, X[lo:hi], coords)
    hist = m.fit(X[lo:hi], coords, max_iter=4, tol=0.0)
    q.put((rank, hist, {k: v.numpy().copy() for k, v in m.state_dict().items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("arch", ["single", "single_distance_binned"])
def test_two_rank_fit_equals_single_process(arch):
    from oracle import hmm_oracle as O
    from smftools_b200.dist import shard_bounds

    assert [shard_bounds(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, arch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer on the full data
    K = 3
    X = O.synth_reads(41, 60, K, 77, prob=True).astype(np.float64)
    coords = O.synth_coords(60, 3) if arch == "single_distance_binned" else np.arange(60)
    p0 = O.init_params(K, [0.8, 0.5, 0.1], arch=arch)
    q_ref, hist_ref = O.fit(X, p0, coords, max_iter=4, tol=0.0)
    (_, h0, sd0), (_, h1, sd1) = res
    assert h0 == h1  # both ranks saw the same global log-likelihood
    np.testing.assert_allclose(h0, hist_ref, rtol=1e-12)
    for k in sd0:
        np.testing.assert_array_equal(sd0[k], sd1[k])  # bit-identical parameters on every rank
    np.testing.assert_allclose(sd0["emission"], q_ref.emission, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sd0["start"], q_ref.start, rtol=0, atol=1e-12)
    if arch == "single_distance_binned":
        np.testing.assert_allclose(sd0["trans_by_bin"], q_ref.trans_by_bin, rtol=0, atol=1e-12)
    else:
        np.testing.assert_allclose(sd0["trans"], q_ref.trans, rtol=0, atol=1e-12)
