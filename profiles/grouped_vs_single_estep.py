"""Single-model E-step (smf_hmm_estep -> non-grouped sequential kernels) against the grouped kernels on the same batch:
    python profiles/grouped_vs_single_estep.py K L N [u8|f32]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import smftools_b200 as S  # noqa: E402
from smftools_b200 import engine, grouped  # noqa: E402

K, L, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
u8 = (sys.argv[4] if len(sys.argv) > 4 else "u8") == "u8"
dev = torch.device("cuda", 0)
m = bench.make_model(K, dev)
tables = m._host_tables()
obs = bench.device_reads(N, L, K, False, 1, dev, coded_u8=u8)
ws = engine.estep_workspace(tables, dev, N, L)
st = torch.empty((engine.stats_len(tables),), dtype=torch.float64, device=dev)
for _ in range(2):
    engine.estep(tables, obs, None, out_stats=st, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    engine.estep(tables, obs, None, out_stats=st, workspace=ws)
e1.record()
torch.cuda.synchronize()
print(f"single-model E-step  K={K} {N} x {L} {'u8' if u8 else 'f32'}: {e0.elapsed_time(e1) / 5:.3f} ms per iteration", flush=True)
X = S.CodedU8(obs) if u8 else obs
iters = 5
grouped.fit_groups([bench.make_model(K, dev)], [X], max_iter=2, tol=0.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
grouped.fit_groups([bench.make_model(K, dev)], [X], max_iter=iters, tol=0.0)
torch.cuda.synchronize()
print(f"grouped fit (1 group): {(time.perf_counter() - t0) * 1e3 / iters:.3f} ms per iteration (E-step + device M-step)", flush=True)
