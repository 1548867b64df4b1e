"""Shared helpers for the GPU parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import hmm_oracle as O


def arch_of(g):
    return str(g["arch"]) if "arch" in g else "single"


def model_from_params(p: O.Params, arch="single", device=None):
    """smftools_b200 model carrying exactly the oracle parameters ``p``."""
    import smftools_b200 as S

    K = p.K
    cfg = {"hmm_n_states": K}
    if arch == "multi":
        cfg["hmm_n_channels"] = int(p.emission.shape[1])
    if arch == "single_distance_binned":
        cfg["hmm_distance_bins"] = [int(b) for b in p.distance_bins]
    m = S.create_hmm(cfg, arch=arch, device=device)
    with torch.no_grad():
        m.start.data = torch.tensor(p.start, dtype=torch.float64)
        m.trans.data = torch.tensor(p.trans, dtype=torch.float64)
        m.emission.data = torch.tensor(p.emission, dtype=torch.float64).reshape(m.emission.shape)
        if p.trans_by_bin is not None:
            m.trans_by_bin.data = torch.tensor(p.trans_by_bin, dtype=torch.float64)
    return m


def params_of_model(m) -> O.Params:
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    return O.Params(sd["start"], sd["trans"], sd["emission"], eps=m.eps, trans_by_bin=sd.get("trans_by_bin"),
                    distance_bins=getattr(m, "distance_bins", None))


def assert_states_match_up_to_ties(states, ref_states, gamma_ref, tol=2e-5):
    """Marginal states must equal argmax of the reference posterior except where the top two
    posteriors are within ``tol`` (fp32 kernel vs fp64 reference)."""
    srt = np.sort(gamma_ref, axis=2)
    near = (srt[:, :, -1] - srt[:, :, -2]) < tol
    bad = (states != ref_states) & ~near
    assert not bad.any(), f"{bad.sum()} state mismatches away from ties"


def assert_viterbi_matches(X, p, coords, states, tol=1e-6):
    """Bit-exact Viterbi paths, except on reads containing a decision whose competing
    log-probabilities are within ``tol``; those must still be (near-)optimal paths."""
    ref, margin = O.viterbi(X, p, coords, return_margins=True)
    same = (states == ref).all(axis=1)
    tie = margin < tol
    assert np.all(same | tie), f"{(~(same | tie)).sum()} reads differ without a near-tie"
    if (~same).any():
        idx = np.nonzero(~same)[0]
        lp_ref = O.path_log_prob(X[idx], ref[idx], p, coords)
        lp_our = O.path_log_prob(X[idx], states[idx].astype(np.int64), p, coords)
        assert np.all(np.abs(lp_ref - lp_our) <= tol * np.maximum(1.0, np.abs(lp_ref)) * 10)
    return int((~same).sum())
