"""Out-of-bounds guard: every output buffer of the round-2 kernels (sequential posterior with its rounded 16-byte
copy-out pieces, direct Viterbi with 4- / 16-byte stores, multi-channel and binned fast path, grouped decode) sits
between sentinel zones that must come back untouched (compute-sanitizer is not available on the GPU pool)."""
import numpy as np
import pytest
import torch

from helpers import model_from_params
from oracle import hmm_oracle as O

pytestmark = pytest.mark.gpu

GUARD = 4096


def _guarded(shape, dtype, dev, fill):
    n = int(np.prod(shape))
    es = torch.empty((), dtype=dtype).element_size()
    gelems = GUARD // es
    buf = torch.full((n + 2 * gelems,), fill, dtype=dtype, device=dev)
    view = buf[gelems:gelems + n].view(*shape)
    return buf, view, gelems


def _intact(buf, gelems, fill):
    head, tail = buf[:gelems], buf[-gelems:]
    if buf.dtype.is_floating_point:
        return bool((head == fill).all() and (tail == fill).all())
    return bool((head == fill).all() and (tail == fill).all())


@pytest.mark.parametrize("K,L,N,seq,vit", [(3, 500, 77, 1, 1), (4, 1204, 31, 1, 1), (2, 20, 130, 1, 1), (3, 16, 200, 1, 0),
                                           (4, 2000, 9, -1, -1), (2, 5000, 5, 1, 1), (3, 37, 50, -1, -1)])
def test_posterior_viterbi_outputs_stay_inside_their_buffers(K, L, N, seq, vit):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    p = O.synth_params(K)
    tables = model_from_params(p)._host_tables()
    X = O.synth_reads(N, L, K, 8800 + L, prob=True, nan_rate=0.3).astype(np.float32)
    obs = torch.from_numpy(X).to(dev)
    engine.set_option("seq", seq)
    engine.set_option("vit", vit)
    try:
        for gstate in (-1, K - 1):
            gb, gv, gg = _guarded((N, L, K) if gstate < 0 else (N, L), torch.float32, dev, -7.0)
            sb, sv, sg = _guarded((N, L), torch.int8, dev, 99)
            engine.posterior(tables, obs, None, gamma_state=gstate, out_gamma=gv, out_states=sv)
            torch.cuda.synchronize()
            assert _intact(gb, gg, -7.0) and _intact(sb, sg, 99), ("posterior", gstate)
            assert bool((gv >= 0).all()) and bool((sv >= 0).all() and (sv < K).all())
        vb, vv, vg = _guarded((N, L), torch.int8, dev, 99)
        wb, wv, wg = _guarded((N * L,), torch.uint8, dev, 77)
        engine.viterbi(tables, obs, None, out_states=vv, workspace=wv)
        torch.cuda.synchronize()
        assert _intact(vb, vg, 99) and _intact(wb, wg, 77), "viterbi"
        assert bool((vv >= 0).all() and (vv < K).all())
    finally:
        engine.set_option("seq", -1)
        engine.set_option("vit", -1)


@pytest.mark.parametrize("arch,K,C,L", [("multi", 2, 2, 333), ("multi", 3, 3, 1000), ("single_distance_binned", 2, 1, 35),
                                        ("single_distance_binned", 3, 1, 2048)])
def test_multi_and_binned_fast_path_outputs_stay_inside_their_buffers(arch, K, C, L):
    from smftools_b200 import engine

    dev = torch.device("cuda")
    kw = {"distance_bins": [1, 5, 10, 25, 50, 100]} if arch != "multi" else {}
    p = O.synth_params(K, arch=arch, n_channels=C, **kw)
    m = model_from_params(p, arch=arch)
    tables = m._host_tables()
    N = 23
    X = O.synth_reads(N, L, K, 31 + L, prob=True, nan_rate=0.3, n_channels=C).astype(np.float32)
    obs = torch.from_numpy(X).to(dev)
    bins = m._bins_for(O.synth_coords(L, 4), L, dev)
    gb, gv, gg = _guarded((N, L, K), torch.float32, dev, -7.0)
    sb, sv, sg = _guarded((N, L), torch.int8, dev, 99)
    engine.posterior(tables, obs, bins, out_gamma=gv, out_states=sv)
    torch.cuda.synchronize()
    assert _intact(gb, gg, -7.0) and _intact(sb, sg, 99)
    assert bool((gv >= 0).all()) and bool((sv >= 0).all() and (sv < K).all())
