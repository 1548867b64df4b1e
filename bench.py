#!/usr/bin/env python
"""bench.py -- HMM footprinting hot path throughput on B200 (read-positions / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3p|...]

One "step" = one pass of the hot path over one batch of synthetic reads.  At N=1 the default
workload is BASELINE.json configs[1]: 100k reads x 5 kb binary GpC calls (30 % NaN), 2-state
posterior decode.  With N>1 (torchrun) every rank decodes its own batch of the same shape (weak
scaling, no data-path collective) and additionally times Baum-Welch iterations whose sufficient
statistics are NCCL-all-reduced; rank 0 prints ONE JSON line.

`--impl reference` times the CPU port of the reference algorithm (oracle/hmm_oracle.py: the
reference is pure Python and /root/reference does not travel to the GPU box) on the host cores,
one single-threaded worker process per core like the reference's own process pool.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {  # algorithmic bytes per read-position: obs (es bytes) in + gamma f32*K + state i8
    "posterior": lambda K, C, es=4: es * C + 4 * K + 1,
    "viterbi": lambda K, C, es=4: es * C + 1,
    "estep": lambda K, C, es=4: es * C,
}

WORKLOADS = {
    # name: (N, L, K, prob, op)
    "c2": dict(N=100_000, L=5_000, K=2, prob=False, op="posterior",
               desc="BASELINE configs[1]: 100k reads x 5 kb binary GpC (30% NaN), 2-state posterior decode"),
    "c3p": dict(N=65_536, L=10_000, K=2, prob=False, op="posterior",
                desc="headline shape: 1M x 10 kb posterior decode, one 65,536-read device-resident chunk"),
    "c3p_k3": dict(N=65_536, L=10_000, K=3, prob=True, op="posterior",
                   desc="1M x 10 kb m6A probabilistic, 3-state posterior decode, one 65,536-read chunk"),
    "c3v": dict(N=262_144, L=10_000, K=3, prob=True, op="viterbi",
                desc="BASELINE configs[2]: 1M x 10 kb m6A probabilistic 3-state Viterbi + footprint interval calling, "
                     "one 262,144-read chunk"),
    "c4": dict(N=262_144, L=8_000, K=4, prob=False, op="estep",
               desc="BASELINE configs[3]: 4-state Baum-Welch iteration, 262,144 reads x 8 kb per GPU"),
    "c5": dict(N=500_000, L=20_000, K=2, prob=False, op="groups",
               desc="BASELINE configs[4]: 16 reference groups x 500,000 reads, ragged lengths 2-20 kb, "
                    "per-group 2-state HMM: 5 Baum-Welch iterations (grouped device-resident fit) + posterior decode"),
}


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def device_reads(N, L, K, prob, seed, device, coded_u8=False):
    """Synthetic SMF-like reads generated ON the device (fp32, NaN = missing): K-state block
    structure with the survey's per-state emission means, Bernoulli calls or k/255-quantised
    probabilities, 30 % iid NaN plus read-span truncation.  ``coded_u8``: binary calls as the kernels'
    compact uint8 code (0 / 1 / 255 = missing) instead of fp32."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    em = torch.tensor([0.85, 0.08, 0.35, 0.60][:K], device=device)
    blk = 40
    nb = (L + blk - 1) // blk
    out = torch.empty((N, L), dtype=torch.uint8 if coded_u8 else torch.float32, device=device)
    step = max(1, min(N, (1 << 28) // max(L, 1)))
    for lo in range(0, N, step):
        hi = min(N, lo + step)
        n = hi - lo
        st = torch.randint(0, K, (n, nb), device=device, generator=gen)
        keep = torch.rand((n, nb), device=device, generator=gen) < 0.6  # sticky: often keep previous block's state
        keep[:, 0] = False
        src = torch.where(keep, torch.zeros_like(st), torch.arange(nb, device=device)[None, :].expand(n, nb))
        st = st.gather(1, torch.cummax(src, dim=1).values)  # state of the last block that drew a fresh state
        p = em[st].repeat_interleave(blk, dim=1)[:, :L]
        u = torch.rand((n, L), device=device, generator=gen)
        if prob:
            noise = torch.rand((n, L), device=device, generator=gen)
            x = (p + (noise - 0.5) * 0.5).clamp_(0.0, 1.0)
            x = torch.where(u < 0.15, 1.0 - x, x)
            x = torch.round(x * 255.0) / 255.0
        else:
            x = (u < p).float()
        x[torch.rand((n, L), device=device, generator=gen) < 0.3] = float("nan")
        lo_t = torch.randint(0, max(1, L // 10), (n, 1), device=device, generator=gen)
        hi_t = L - torch.randint(0, max(1, L // 10), (n, 1), device=device, generator=gen)
        pos = torch.arange(L, device=device)[None, :]
        x[(pos < lo_t) | (pos >= hi_t)] = float("nan")
        out[lo:hi] = torch.where(torch.isnan(x), torch.full_like(x, 255.0), x).to(torch.uint8) if coded_u8 else x
        del st, keep, p, u, x
    return out


def bench_params(K):
    """'Fitted-looking' parameters: dwell 60/150/25/80 positions, emission
    .85/.08/.35/.60 -- the same table the CPU port uses (oracle.synth_params)."""
    dwell = [60.0, 150.0, 25.0, 80.0][:K]
    trans = np.zeros((K, K))
    for i in range(K):
        leave = 1.0 / dwell[i]
        trans[i] = leave / (K - 1)
        trans[i, i] = 1.0 - leave
    start = np.linspace(1.0, 2.0, K)
    start /= start.sum()
    eps = 1e-8
    start = (start + eps) / (start + eps).sum()
    trans = (trans + eps) / (trans + eps).sum(axis=1, keepdims=True)
    em = np.clip(np.array([0.85, 0.08, 0.35, 0.60][:K]), eps, 1 - eps)
    return start, trans, em


def make_model(K, device):
    import torch

    import smftools_b200 as S

    start, trans, em = bench_params(K)
    m = S.create_hmm({"hmm_n_states": K}, arch="single", device=None)
    with torch.no_grad():
        m.start.data = torch.tensor(start, dtype=torch.float64)
        m.trans.data = torch.tensor(trans, dtype=torch.float64)
        m.emission.data = torch.tensor(em, dtype=torch.float64)
    return m


# --------------------------------------------------------------------------------------
# CPU baseline (oracle port on the host cores)
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    n, L, K, prob, op, seed, reps = args
    from oracle import hmm_oracle as O

    X = O.synth_reads(n, L, K, seed, prob=prob, nan_rate=0.3, truncate=False).astype(np.float64)
    p = O.synth_params(K)
    t0 = time.perf_counter()
    for _ in range(reps):  # same batch again: bounded memory, more CPU seconds
        if op == "posterior":
            O.decode(X, p, None, "marginal")
        elif op == "viterbi":
            O.decode(X, p, None, "viterbi")
        else:
            O.estep(X, p, None)
    return time.perf_counter() - t0


def cpu_baseline(wl, cores=None, reads_per_worker=None, reps=4):
    """Time the CPU port on a bounded sample: P single-threaded worker processes (the reference's
    own way of using all cores, parallel_utils.py:25-96), wall = slowest worker."""
    import multiprocessing as mp

    if cores is None:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    L, K = wl["L"], wl["K"]
    if reads_per_worker is None:
        reads_per_worker = max(64, int(8e6 // L))  # ~8e6 read-positions per worker (~10-20 s)
    ctx = mp.get_context("spawn")
    args = [(reads_per_worker, L, K, wl["prob"], wl["op"], 9000 + i, reps) for i in range(workers)]
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        times = pool.map(_cpu_worker, args)
    wall_all = time.perf_counter() - t0
    rp = workers * reads_per_worker * L * reps
    return {
        "value": rp / max(times), "unit": "read-positions/s", "cores": workers, "kind": "port",
        "sample": f"{workers} workers x {reps} passes x {reads_per_worker} reads x {L} positions, K={K}, op={wl['op']}, "
                  f"numpy fp64 oracle port, compute time of slowest worker {max(times):.1f}s (pool wall {wall_all:.1f}s)",
    }


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    total = args.steps + args.warmup
    # bounded: the whole run must end within a few minutes
    per = max(32, int(2.5e6 // wl["L"]))
    for i in range(total):
        last = cpu_baseline(wl, reads_per_worker=per, reps=2)
        if i >= args.warmup:
            vals.append(last["value"])
    v = statistics.median(vals)
    last["value"] = v
    rp_step = last["cores"] * per * wl["L"] * 2
    line = {
        "impl": "reference", "metric": "HMM read-positions/sec", "value": v, "unit": "read-positions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rp_step / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "op": wl["op"], "K": wl["K"], "L": wl["L"],
                   "note": "reference algorithm (CPU port, oracle/hmm_oracle.py) on a bounded sample per step"},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "read-positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
PARITY_TOL = {"gamma_abs": 1e-5, "viterbi_margin": 1e-6, "stats_rel": 2e-6}


def _dist_ctx():
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    return world, rank, local, torch.device("cuda", local)


def _max_over_ranks(ms, device, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def _feature_setup(L, device):
    """Sites every ~4 bp on a grid of 4*L columns; default footprint size classes."""
    import torch

    coords_h = np.sort(np.random.default_rng(5).choice(np.arange(4 * L), size=L, replace=False)).astype(np.int64)
    full_h = np.arange(4 * L, dtype=np.int64)
    return dict(
        coords_h=coords_h, full_h=full_h,
        coords=torch.from_numpy(coords_h).to(device),
        left=torch.from_numpy(np.searchsorted(full_h, coords_h, side="left").astype(np.int32)).to(device),
        right=torch.from_numpy(np.searchsorted(full_h, coords_h, side="right").astype(np.int32)).to(device),
        V=4 * L, ranges=[(6, 40), (40, 100), (100, 200), (200, float("inf"))], target=1)


def parity_record(op, K, obs_prefix, tables, feat=None):
    """Outside every timed region: the kernels' results on a read prefix against the CPU oracle
    (test infrastructure; the only place bench.py's GPU arm touches oracle/)."""
    import torch

    from oracle import hmm_oracle as O
    from smftools_b200 import engine

    t0 = time.perf_counter()
    p = O.synth_params(K)
    n, L = int(obs_prefix.shape[0]), int(obs_prefix.shape[1])
    if obs_prefix.dtype == torch.uint8:
        xh = obs_prefix.cpu().numpy().astype(np.float64)
        xh[xh > 1] = np.nan
    else:
        xh = obs_prefix.cpu().numpy().astype(np.float64)
    rec = {"reads": n, "positions": L, "against": "oracle/hmm_oracle.py (numpy fp64 port pinned to the reference)"}
    if op == "posterior":
        g, s, _ = engine.posterior(tables, obs_prefix, None)
        g_ref, _ = O.posterior(xh, p)
        err = float(np.abs(g.cpu().numpy() - g_ref).max())
        top2 = np.sort(g_ref, axis=2)
        close = (top2[:, :, -1] - top2[:, :, -2]) < 2e-5
        bad = int(((s.cpu().numpy() != g_ref.argmax(axis=2)) & ~close).sum())
        rec.update({"max_abs_dgamma": err, "tol": PARITY_TOL["gamma_abs"], "state_mismatches_outside_ties": bad,
                    "ok": bool(err <= PARITY_TOL["gamma_abs"] and bad == 0)})
    elif op == "viterbi":
        sv = engine.viterbi(tables, obs_prefix, None).cpu().numpy().astype(np.int64)
        ref, margin = O.viterbi(xh, p, return_margins=True)
        same = (sv == ref).all(axis=1)
        viol = int((~same & ~(margin < PARITY_TOL["viterbi_margin"])).sum())
        rec.update({"reads_identical": int(same.sum()), "reads_differing_within_1e-6_margin": int((~same).sum()) - viol,
                    "violations": viol, "ok": viol == 0})
        if feat is not None:
            ivals = engine.call_features(states=torch.from_numpy(sv.astype(np.int8)).to(obs_prefix.device), post=None,
                                         target=feat["target"], threshold=0.0, coords=feat["coords"],
                                         grid_left=feat["left"], grid_right=feat["right"], n_grid=feat["V"],
                                         ranges=feat["ranges"], want_dense=False, max_intervals=n * (L // 2 + 1))[2]
            got = sorted(map(tuple, ivals.cpu().numpy().tolist()))
            want = []
            for i in range(n):
                for (a, b) in O.runs_from_bool(sv[i] == feat["target"]):
                    glen = int(feat["coords_h"][b - 1]) - int(feat["coords_h"][a]) + 1
                    c = O.pick_class(feat["ranges"], glen)
                    if c >= 0:
                        want.append((i, int(np.searchsorted(feat["full_h"], feat["coords_h"][a], side="left")),
                                     int(np.searchsorted(feat["full_h"], feat["coords_h"][b - 1], side="right")), c))
            rec["intervals"] = len(got)
            rec["intervals_identical"] = bool(got == sorted(want))
            rec["ok"] = bool(rec["ok"] and rec["intervals_identical"])
    else:
        st = engine.estep(tables, obs_prefix, None).cpu().numpy()
        ref = O.stats_vector(O.estep(xh, p))
        rel = float((np.abs(st - ref) / np.maximum(np.abs(ref), 1.0)).max())
        rec.update({"max_rel_dstats": rel, "tol": PARITY_TOL["stats_rel"], "ok": bool(rel <= PARITY_TOL["stats_rel"])})
    rec["seconds"] = round(time.perf_counter() - t0, 2)
    return rec


def roofline_of(op, K, es, positions, kernel_ms, workload=None):
    peak, peak_src = measured_peak()
    algo = ALGO_BYTES[op](K, 1, es)
    achieved = algo * positions / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(workload) if workload else None, "peak_source": peak_src,
            "algorithmic_bytes_per_read_position": algo, "kernel_ms": kernel_ms}


def stream_workload(name, ctx, steps_cap=None):
    """A BASELINE shape at its FULL size, streamed through the device in resident chunks dealt to the
    ranks; generation is outside the timed region (CUDA events around each chunk's kernels)."""
    import torch

    from smftools_b200 import engine

    world, rank, local, device = ctx
    spec = FULL_WORKLOADS[name]
    K, L, op, prob, es = spec["K"], spec["L"], spec["op"], spec["prob"], spec.get("es", 4)
    chunk, n_chunks = spec["chunk"], spec["chunks"]
    model = make_model(K, device)
    tables = model._host_tables()
    mine = [c for c in range(n_chunks) if c % world == rank]
    gamma = torch.empty((chunk, L, K), dtype=torch.float32, device=device) if op == "posterior" else None
    states = torch.empty((chunk, L), dtype=torch.int8, device=device) if op in ("posterior", "viterbi") else None
    ws = stats = feat = None
    if op == "viterbi":
        ws = torch.empty((chunk * L,), dtype=torch.uint8, device=device)
        feat = _feature_setup(L, device)
    if op == "estep":
        ws = engine.estep_workspace(tables, device, chunk, L)
        stats = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=device)
    launches = 0
    n_intervals = 0

    def run(obs):
        nonlocal launches, n_intervals
        if op == "posterior":
            engine.posterior(tables, obs, None, out_gamma=gamma, out_states=states)
            launches += 1
        elif op == "viterbi":
            engine.viterbi(tables, obs, None, out_states=states, workspace=ws)
            iv = engine.call_features(states=states, post=None, target=feat["target"], threshold=0.0,
                                      coords=feat["coords"], grid_left=feat["left"], grid_right=feat["right"],
                                      n_grid=feat["V"], ranges=feat["ranges"], want_dense=False,
                                      max_intervals=int(chunk) * 64)[2]
            n_intervals += int(iv.shape[0])
            launches += 2
        else:
            engine.estep(tables, obs, None, out_stats=stats, workspace=ws)
            launches += 3

    total_ms = 0.0
    first = None
    for ci, c in enumerate(mine):
        obs = device_reads(chunk, L, K, prob, 7000 + 100 * c + spec.get("seed", 0), device, coded_u8=(es == 1))
        if first is None:
            first = obs[:spec.get("parity_reads", 128)].clone()
            for _ in range(3):  # warm-up on the first chunk
                run(obs)
            launches = 0
            n_intervals = 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(obs)
        b.record()
        torch.cuda.synchronize()
        total_ms += a.elapsed_time(b)
        del obs
    ms = _max_over_ranks(total_ms, device, world)
    positions = float(chunk) * L * n_chunks
    rec = None
    if rank == 0:
        per_rank_positions = float(chunk) * L * len(mine)
        rec = {"name": name, "workload": spec["desc"], "op": op, "K": K, "positions": L, "reads_total": chunk * n_chunks,
               "chunks": n_chunks, "reads_per_chunk": chunk, "obs": "u8 coded" if es == 1 else "f32 NaN-coded",
               "value": positions / (ms * 1e-3), "unit": "read-positions/s", "ms_total": ms,
               "gpu_launches_rank0": launches,
               "roofline": roofline_of(op, K, es, per_rank_positions, total_ms)}
        rec["roofline"]["traffic"] = ncu_traffic(name)  # DRAM bytes per chunk launch from the committed ncu pass, or None
        if name == "c4_u8":
            # the E-step moves 1 algorithmic byte per read-position: it is bound by instruction issue, not HBM.  Thread
            # instructions per read-position from the committed ncu captures (profiles/r02_walk_ncu.txt: 30.5 forward +
            # 106.0 backward) against the chip's issue rate (SMs x 4 schedulers x 32 lanes x SM clock)
            instr = 136.5
            peak_issue = 148 * 4 * 32 * 1.965e9
            rec["roofline"]["issue"] = {"thread_instr_per_read_position": instr, "source": "profiles/r02_walk_ncu.txt",
                                        "achieved_thread_instr_per_s": instr * per_rank_positions / (total_ms * 1e-3),
                                        "peak_thread_instr_per_s": peak_issue,
                                        "frac": instr * per_rank_positions / (total_ms * 1e-3) / peak_issue}
        rec["roofline"]["algorithmic_bytes_per_launch"] = ALGO_BYTES[op](K, 1, es) * float(chunk) * L
        if op == "viterbi":
            rec["intervals_rank0"] = n_intervals
        rec["parity"] = parity_record(op, K, first, tables, feat)
    del gamma, states, ws
    torch.cuda.empty_cache()
    return rec


FULL_WORKLOADS = {
    "c3p": dict(K=2, L=10_000, op="posterior", prob=False, chunk=65_536, chunks=16,
                desc="headline target: posterior decode of 1M (16 x 65,536) reads x 10 kb, K=2 binary, f32 obs"),
    "c3p_u8": dict(K=2, L=10_000, op="posterior", prob=False, chunk=65_536, chunks=16, es=1,
                   desc="posterior decode of 1M (16 x 65,536) reads x 10 kb, K=2 binary, uint8-coded obs (what the device "
                        "input builders emit: 0 / 1 / 255 = missing)"),
    "c3p_k3": dict(K=3, L=10_000, op="posterior", prob=True, chunk=65_536, chunks=16,
                   desc="posterior decode of 1M (16 x 65,536) reads x 10 kb, K=3 m6A probabilistic, f32 obs"),
    "c3v": dict(K=3, L=10_000, op="viterbi", prob=True, chunk=262_144, chunks=4, parity_reads=96,
                desc="BASELINE configs[2]: 1M (4 x 262,144) reads x 10 kb m6A probabilistic, 3-state Viterbi + "
                     "footprint interval calling"),
    "c4": dict(K=4, L=8_000, op="estep", prob=False, chunk=262_144, chunks=2, parity_reads=96,
               desc="BASELINE configs[3] kernel: 4-state Baum-Welch E-step, 2 x 262,144 reads x 8 kb, f32 obs"),
    "c4_u8": dict(K=4, L=8_000, op="estep", prob=False, chunk=262_144, chunks=2, es=1, parity_reads=96,
                  desc="BASELINE configs[3] kernel: 4-state Baum-Welch E-step, 2 x 262,144 reads x 8 kb, u8 coded obs"),
}


def fit_strong_scaling(args, ctx):
    """BASELINE configs[3] as north_star states it: 4-state HMM, 8 kb reads, 20 Baum-Welch iterations over
    ``--bw-reads`` (4 M) reads IN TOTAL, sharded N ways (strong scaling), through ``model.fit`` with the
    NCCL all-reduce of the 29-double statistics vector, the host M-step and the convergence test."""
    import torch

    import smftools_b200 as S
    from smftools_b200.dist import enable_distributed_fit, shard_bounds

    world, rank, local, device = ctx
    K, L, total = 4, 8_000, int(args.bw_reads)
    lo, hi = shard_bounds(total, rank, world)
    n_local = hi - lo
    obs = device_reads(n_local, L, K, False, 4000 + rank, device, coded_u8=True)
    torch.cuda.synchronize()

    def fresh_model():
        m = S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": [0.85, 0.08, 0.35, 0.60]}, arch="single")
        if world > 1:
            enable_distributed_fit(m)
        return m

    m = fresh_model()
    m.fit(S.CodedU8(obs), max_iter=2, tol=0.0)  # warm-up (allocations, NCCL channels)
    m = fresh_model()
    m.fit_profile = []
    iters = int(args.bw_iters)
    _barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    hist = m.fit(S.CodedU8(obs), max_iter=iters, tol=0.0)
    b.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = _max_over_ranks(a.elapsed_time(b), device, world)
    wall = _max_over_ranks(wall * 1e3, device, world)
    rec = None
    if rank == 0:
        prof = m.fit_profile
        mean = lambda k: float(statistics.mean(p[k] for p in prof)) if prof else None  # noqa: E731
        rec = {
            "config": f"BASELINE configs[3]: K=4, {total} reads x {L} positions in total, sharded over {world} GPU(s) "
                      f"({n_local} on rank 0), u8 coded obs, {iters} iterations through model.fit (tol=0)",
            "scaling": "strong", "value": float(total) * L * len(hist) / (ms * 1e-3),
            "unit": "read-positions/s (all iterations)", "iterations": len(hist), "ms_total": ms, "wall_ms_total": wall,
            "ms_per_iter": ms / max(len(hist), 1),
            "split_ms_per_iter": {"estep_kernels": mean("estep_ms"), "allreduce": mean("allreduce_ms"),
                                  "host_mstep_tables": mean("host_mstep_ms")},
            "collective": (f"nccl all_reduce(sum) of {K + K * K + 2 * K + 1} fp64 per iteration" if world > 1 else "none"),
            "loglik_first_last": [hist[0], hist[-1]], "loglik_monotone": bool(all(b2 >= a2 - 1e-6 * abs(a2) for a2, b2 in zip(hist, hist[1:]))),
            "reads_total": total, "reads_rank0": n_local,
        }
    del obs
    torch.cuda.empty_cache()
    return rec


def host_link_probe(ctx, seconds=2.0, nbytes=1 << 30):
    """Pinned H2D, D2H and concurrent both-way copies on every rank at once: the host-link ceiling the
    end-to-end numbers are bounded by (GB/s summed over ranks)."""
    import torch
    import torch.distributed as dist

    world, rank, local, device = ctx
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    out = {}
    for kind in ("h2d", "d2h", "both"):
        _barrier(world)
        reps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
            reps += 1
        rate = nbytes * reps / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([rate], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t)
        out[f"{kind}_gbs" if kind != "both" else "bidir_gbs_per_direction"] = float(t.item())
    out["note"] = f"pinned {nbytes >> 20} MiB copies for {seconds:.0f} s per pattern, all {world} rank(s) at once, summed"
    return out


def e2e_legs(args, ctx, model, obs, N, L, K):
    """End to end through the public API with HOST buffers, copies inside the timed region.
    ``e2e`` is the reference signature itself: ``decode(X float64 pageable numpy)`` returning fresh
    ``(int64 states, float64 posteriors)`` host arrays (HMM.py:673-720).  ``e2e_compact`` is the opt-in
    extension (float32 in, int8 + float32 out into caller-provided pinned buffers)."""
    import torch

    world, rank, local, device = ctx
    steps = max(1, min(args.steps, 3))
    n_e2e = min(N, 32_768)
    X64 = obs[:n_e2e].cpu().numpy().astype(np.float64)  # ordinary pageable numpy, like the reference's callers hold
    res = model.decode(X64, decode="marginal")  # warm-up: fills the page-locked result pool
    del res
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        st, g = model.decode(X64, decode="marginal")
        del st, g
    torch.cuda.synchronize()
    dt = _max_over_ranks(time.perf_counter() - t0, device, world)
    e2e = {
        "value": world * n_e2e * L * steps / dt, "unit": "read-positions/s",
        "h2d_bytes_per_step": int(n_e2e * L * 8), "d2h_bytes_per_step": int(n_e2e * L * (8 * K + 8)),
        "steps": steps, "reads_per_step": n_e2e,
        "api": "SingleBernoulliHMM.decode(X_host_f64_pageable, decode='marginal') -> fresh (int64 states, float64 gamma) "
               "numpy arrays: the reference signature, nothing pre-allocated by the caller",
        "d2h_gbs_per_gpu": n_e2e * L * (8 * K + 8) * steps / dt / 1e9,
    }
    del X64
    host_in = torch.empty((n_e2e, L), dtype=torch.float32, pin_memory=True)
    host_in.copy_(obs[:n_e2e])
    out_states = torch.empty((n_e2e, L), dtype=torch.int8, pin_memory=True)
    out_gamma = torch.empty((n_e2e, L, K), dtype=torch.float32, pin_memory=True)
    X_host = host_in.numpy()
    outs = (out_states.numpy(), out_gamma.numpy())
    model.decode(X_host, decode="marginal", compact=True, out=outs)
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.decode(X_host, decode="marginal", compact=True, out=outs)
    torch.cuda.synchronize()
    dt = _max_over_ranks(time.perf_counter() - t0, device, world)
    compact = {
        "value": world * n_e2e * L * steps / dt, "unit": "read-positions/s",
        "h2d_bytes_per_step": int(n_e2e * L * 4), "d2h_bytes_per_step": int(n_e2e * L * (4 * K + 1)),
        "steps": steps, "reads_per_step": n_e2e,
        "api": "extension: decode(X_host_f32_pinned, compact=True, out=pinned (int8, float32) buffers)",
    }
    del host_in, out_states, out_gamma
    return e2e, compact


def groups_workload(ctx, reads_per_group=500_000, n_groups=16, iters=5, K=2):
    """BASELINE configs[4]: per-reference HMMs on 16 references x 500k reads x 2-20 kb ragged lengths.
    With one GPU all groups are fitted in ONE grouped, device-resident Baum-Welch (smf_hmm_grouped_fit: `iters`
    iterations, M-step and stop rule on the device, no host round trip).  With N GPUs a group (up to 1e10 read-positions)
    is larger than a rank's share, so whole-group dealing cannot balance: the groups are laid end to end and the reads
    cut into N equal-cost pieces (grouped.partition_reads) -- a group that crosses a cut is read-sharded -- every rank
    runs the same grouped fit on its shards, and per iteration the [16, S] statistics matrix is NCCL-all-reduced between
    the device E-step and the device M-step (stream-ordered, still no host round trip), which leaves identical
    parameters on every rank.  Then every rank posterior-decodes its own reads with each group's model."""
    import torch
    import torch.distributed as dist

    import smftools_b200 as S
    from smftools_b200 import engine, grouped

    world, rank, local, device = ctx
    rng = np.random.default_rng(17)
    lengths = (np.exp(rng.uniform(np.log(2000), np.log(20000), size=n_groups)).astype(int) // 16 * 16).tolist()
    N = int(reads_per_group)
    costs = [float(N) * L for L in lengths]
    parts = grouped.partition_reads([N] * n_groups, lengths, world)
    local = {g: (lo, hi) for g, lo, hi in parts[rank]}
    mine = sorted(local)

    def fresh_models():
        return [S.create_hmm({"hmm_n_states": K, "hmm_init_emission_probs": [0.8 - 0.01 * g, 0.2 + 0.005 * g][:K]},
                             arch="single") for g in range(n_groups)]

    data = {g: device_reads(local[g][1] - local[g][0], lengths[g], K, False, 500 + g + 1000 * rank, device, coded_u8=True)
            for g in mine}
    torch.cuda.synchronize()
    # warm-up on read prefixes (module load, allocator) + in-run parity of the same calls against the oracle
    parity = None
    models = fresh_models()
    if mine:
        pre = [S.CodedU8(data[g][:96]) for g in mine[:3]]
        pm = [models[g] for g in mine[:3]]
        init = grouped.pack_parameters(pm)
        hist = grouped.fit_groups(pm, pre, max_iter=iters, tol=0.0)
        dec = grouped.decode_groups(pm, pre)
        torch.cuda.synchronize()
        if rank == 0:
            from oracle import hmm_oracle as O

            t0 = time.perf_counter()
            worst_ll = worst_p = worst_g = 0.0
            for i, g in enumerate(mine[:3]):
                xh = data[g][:96].cpu().numpy().astype(np.float64)
                xh[xh > 1] = np.nan
                p0 = O.Params(init[i, :K], init[i, K:K + K * K].reshape(K, K), init[i, K + K * K:], eps=1e-8)
                q, h_ref = O.fit(xh, p0, max_iter=iters, tol=0.0)
                worst_ll = max(worst_ll, float(np.max(np.abs(np.array(hist[i]) - np.array(h_ref)) / np.maximum(np.abs(h_ref), 1.0))))
                got = grouped.pack_parameters([pm[i]])[0]
                want = np.concatenate([q.start, q.trans.reshape(-1), q.emission.reshape(-1)])
                worst_p = max(worst_p, float(np.abs(got - want).max()))
                # the decode ran with the device-fitted parameters: the oracle decodes with exactly those
                pg = O.Params(got[:K], got[K:K + K * K].reshape(K, K), got[K + K * K:], eps=1e-8)
                g_ref, _ = O.posterior(xh, pg)
                worst_g = max(worst_g, float(np.abs(dec[i][0].cpu().numpy() - g_ref).max()))
            parity = {"groups": len(mine[:3]), "reads_per_group": 96, "iterations": iters,
                      "against": "oracle/hmm_oracle.py fit + posterior per group",
                      "max_rel_dloglik": worst_ll, "max_abs_dparam": worst_p, "max_abs_dgamma": worst_g,
                      "tol": {"loglik_rel": 1e-6, "param_abs": 1e-4, "gamma_abs": 1e-5},
                      "ok": bool(worst_ll <= 1e-6 and worst_p <= 1e-4 and worst_g <= 1e-5),
                      "seconds": round(time.perf_counter() - t0, 2)}
        del pre, dec
    models = fresh_models()
    if world > 1:
        # every rank passes every group, with its own shard of that group's reads (possibly none)
        fit_models = models
        obs_list = [S.CodedU8(data[g] if g in local else torch.empty((0, lengths[g]), dtype=torch.uint8, device=device))
                    for g in range(n_groups)]
    else:
        fit_models = [models[g] for g in mine]
        obs_list = [S.CodedU8(data[g]) for g in mine]
    # reusable decode output buffers (one 65,536-read chunk of the longest group): allocated outside the timed region
    chunk = 65_536
    Lmax = max([lengths[g] for g in mine] or [1])
    gam = torch.empty((chunk * Lmax * K,), dtype=torch.float32, device=device)
    stt = torch.empty((chunk * Lmax,), dtype=torch.int8, device=device)
    _barrier(world)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    hists = grouped.fit_groups(fit_models, obs_list, max_iter=iters, tol=0.0, sharded=world > 1)
    ev[1].record()
    launches = 0
    for g in mine:
        tables = models[g]._host_tables()
        L = lengths[g]
        n_loc = int(data[g].shape[0])
        for lo in range(0, n_loc, chunk):
            n = min(chunk, n_loc - lo)
            engine.posterior(tables, data[g][lo:lo + n], None, out_gamma=gam[:n * L * K].view(n, L, K),
                             out_states=stt[:n * L].view(n, L))
            launches += 1
    ev[2].record()
    torch.cuda.synchronize()
    fit_ms = _max_over_ranks(ev[0].elapsed_time(ev[1]), device, world)
    dec_ms = _max_over_ranks(ev[1].elapsed_time(ev[2]), device, world)
    total_ms = _max_over_ranks(ev[0].elapsed_time(ev[2]), device, world)
    rec = None
    if rank == 0:
        pos = float(sum(costs))
        my_pos = float(sum((local[g][1] - local[g][0]) * lengths[g] for g in mine))
        rec = {"name": "c5", "workload": f"BASELINE configs[4]: {n_groups} reference groups x {N} reads, ragged lengths "
                                         f"{min(lengths)}-{max(lengths)}, per-group {K}-state HMM: {iters} Baum-Welch "
                                         "iterations (grouped device-resident fit) + posterior decode, u8 coded obs",
               "op": "grouped fit + decode", "K": K, "groups": n_groups, "reads_per_group": N, "lengths": lengths,
               "partition": ("all groups on the one GPU" if world == 1 else
                             "reads cut into equal-cost contiguous pieces (grouped.partition_reads): (group, first read, "
                             "last read + 1) per rank; per-iteration NCCL all-reduce of the [%d, %d] statistics matrix"
                             % (n_groups, K + K * K + 2 * K + 1)),
               "reads_of_rank": parts, "value": pos * (iters + 1) / (total_ms * 1e-3), "unit": "read-positions/s",
               "ms_total": total_ms, "fit_ms": fit_ms, "decode_ms": dec_ms,
               "fit_read_positions_per_s": pos * iters / (fit_ms * 1e-3), "decode_read_positions_per_s": pos / (dec_ms * 1e-3),
               "gpu_launches_rank0": launches + 1 + 3 * iters, "host_round_trips_in_fit": 1,
               "roofline": {"fit": roofline_of("estep", K, 1, my_pos * iters, ev[0].elapsed_time(ev[1])),
                            "decode": roofline_of("posterior", K, 1, my_pos, ev[1].elapsed_time(ev[2]))},
               "loglik_monotone": bool(all(all(b2 >= a2 - 1e-6 * abs(a2) for a2, b2 in zip(h, h[1:])) for h in hists)),
               "parity": parity}
    del data, gam, stt, obs_list
    torch.cuda.empty_cache()
    return rec


def run_groups(args, wl):
    """`--workload c5`: BASELINE configs[4] on its own (the default run carries it in `workloads`)."""
    import torch.distributed as dist

    ctx = _dist_ctx()
    world, rank, local, device = ctx
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rec = groups_workload(ctx, reads_per_group=args.group_reads)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        emit({
            "metric": "HMM read-positions/sec", "value": rec["value"], "unit": "read-positions/s", "n_gpus": world,
            "steps": 1, "warmup": 1, "ms_per_step": rec["ms_total"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": rec["workload"], "op": rec["op"], "groups": rec["groups"],
                       "reads_per_group": rec["reads_per_group"], "lengths": rec["lengths"], "K": rec["K"]},
            "clocks": clocks, "gpu_launches": rec["gpu_launches_rank0"], "roofline": rec["roofline"]["decode"],
            "parity": rec["parity"], "workloads": [rec],
        })
    if world > 1:
        dist.destroy_process_group()


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from smftools_b200 import engine

    if wl["op"] == "groups":
        return run_groups(args, wl)

    ctx = _dist_ctx()
    world, rank, local, device = ctx
    numa = None
    if world > 1:
        # several ranks stream through host memory at once (e2e leg): keep each rank's host buffers and
        # copy threads on the CPUs local to its GPU
        from smftools_b200.dist import bind_to_gpu_numa_node

        numa = bind_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=device)
    N, L, K, op = wl["N"], wl["L"], wl["K"], wl["op"]
    C = 1
    model = make_model(K, device)
    tables = model._host_tables()
    obs = device_reads(N, L, K, wl["prob"], 1234 + 1000 * rank, device)
    torch.cuda.synchronize()

    gamma = torch.empty((N, L, K), dtype=torch.float32, device=device) if op == "posterior" else None
    states = torch.empty((N, L), dtype=torch.int8, device=device) if op in ("posterior", "viterbi") else None
    ws = None
    if op == "viterbi":
        ws = torch.empty((N * L,), dtype=torch.uint8, device=device)
    if op == "estep":
        ws = engine.estep_workspace(tables, device, N, L)
        stats = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=device)
    feat = _feature_setup(L, device) if op == "viterbi" else None

    launches = {"n": 0}
    # E-step launches per call: two-kernel path (walk + sweeps + reduce) for K = 4 and long K = 3 batches
    # (api.cu walk_eligible), otherwise one forward-backward kernel + reduce
    two_kernel = N >= 8192 and (K == 4 or (K == 3 and L > 8704))
    seq_estep = N >= (24576 if L >= 8192 else (49152 if K > 2 else (65536 if L >= 2000 else 98304))) and L % 4 == 0  # seq_min_reads()
    estep_launches = 3 if (two_kernel or seq_estep) else 2

    def step():
        if op == "posterior":
            engine.posterior(tables, obs, None, out_gamma=gamma, out_states=states)
            launches["n"] += 1
        elif op == "viterbi":
            engine.viterbi(tables, obs, None, out_states=states, workspace=ws)
            engine.call_features(states=states, post=None, target=feat["target"], threshold=0.0, coords=feat["coords"],
                                 grid_left=feat["left"], grid_right=feat["right"], n_grid=feat["V"],
                                 ranges=feat["ranges"], want_dense=False, max_intervals=int(N) * 64)
            launches["n"] += 2
        else:
            engine.estep(tables, obs, None, out_stats=stats, workspace=ws)
            launches["n"] += estep_launches
            if world > 1:
                dist.all_reduce(stats)

    for _ in range(max(3, args.warmup)):
        step()
    _barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    stream = torch.cuda.current_stream(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    _barrier(world)
    t_start.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
    t_end.record(stream)
    _barrier(world)
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop() if rank == 0 else None
    n_launch = launches["n"]
    total_ms = _max_over_ranks(total_ms, device, world)
    ms_per_step = total_ms / args.steps
    value = world * N * L / (ms_per_step * 1e-3)

    # ---- parity of this very process's kernels on a read prefix (outside the timed region) ----
    parity = parity_record(op, K, obs[:256 if L <= 5000 else 128].clone(), tables, feat) if rank == 0 else None

    # ---- end to end through the public API with HOST buffers ----
    e2e = e2e_compact = None
    if op == "posterior" and not args.no_e2e:
        e2e, e2e_compact = e2e_legs(args, ctx, model, obs, N, L, K)
    del gamma, states, ws
    obs_keep = obs if (world == 1 and not args.no_cpu) else None
    del obs
    torch.cuda.empty_cache()

    link = None
    if not args.no_e2e:
        link = host_link_probe(ctx)
        if e2e is not None and rank == 0:
            e2e["fraction_of_host_link_d2h"] = (e2e["d2h_gbs_per_gpu"] * world) / max(link["bidir_gbs_per_direction"], 1e-9)
            e2e["limiter"] = ("device->host link: every read-position returns 8*K + 8 = %d bytes (int64 state + float64 "
                              "posteriors) against %.1f GB/s of measured pinned D2H under two-way traffic" %
                              (8 * K + 8, link["bidir_gbs_per_direction"]))
            e2e["numa_binding_rank0"] = numa

    # ---- Baum-Welch: BASELINE configs[3], strong scaling ----
    bw = fit_strong_scaling(args, ctx) if not args.no_bw else None

    # ---- the other BASELINE shapes at full size ----
    extra = []
    if args.workload == "c2" and not args.no_extra:
        for name in ("c3p", "c3p_u8", "c3p_k3", "c3v", "c4", "c4_u8"):
            rec = stream_workload(name, ctx)
            if rec is not None:
                extra.append(rec)
        rec = groups_workload(ctx, reads_per_group=args.group_reads)
        if rec is not None:
            extra.append(rec)

    if rank == 0:
        k_ms = statistics.mean(kernel_ms)
        roof = roofline_of(op, K, 4, float(N) * L, k_ms, args.workload)
        roof["kernel"] = {"posterior": "smf::fb2_kernel<K, CH, STATS=false, ...> (one launch per step)",
                          "viterbi": "smf::viterbi_kernel (+ smf::call_intervals_kernel)",
                          "estep": ("smf::seq_forward_kernel + smf::seq_backward_kernel<MODE_STATS> (+ reduce_partials_kernel)"
                                    if seq_estep else
                                    "smf::walk_kernel + smf::estep_sweep_kernel (+ reduce_partials_kernel)"
                                    if two_kernel else
                                    "smf::fb2_kernel<K, CH, STATS=true, ...> (+ reduce_partials_kernel)")}[op]
        line = {
            "metric": "HMM read-positions/sec", "value": value, "unit": "read-positions/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "op": op, "reads_per_gpu": N, "positions": L, "K": K, "obs": "f32 NaN-coded",
                       "l2": f"inputs {N * L * 4 / 1e9:.2f} GB + outputs per step, far larger than the 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_compact": e2e_compact,
            "host_link": link,
            "gpu_launches": n_launch,
            "roofline": roof,
            "parity": parity,
            "baum_welch": bw,
            "workloads": extra,
        }
        if obs_keep is not None:
            del obs_keep
            line["cpu_baseline"] = cpu_baseline(wl)
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """Route everything that writes to fd 1 (NCCL's version banner, library chatter) to stderr so
    that the ONE JSON line is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-bw", action="store_true", help="skip the Baum-Welch leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end legs and the host-link probe")
    ap.add_argument("--no-extra", action="store_true", help="skip the full-size runs of the other BASELINE shapes")
    ap.add_argument("--bw-reads", type=int, default=4_000_000, help="total reads of the configs[3] fit (sharded over the GPUs)")
    ap.add_argument("--bw-iters", type=int, default=20, help="Baum-Welch iterations of the configs[3] fit")
    ap.add_argument("--group-reads", type=int, default=500_000, help="reads per reference group of the configs[4] workload")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
