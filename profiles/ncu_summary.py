#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall mix and per-opcode thread-instructions per read-position.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep <read_positions_per_launch>
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rp = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:62s} {vals[i]:>20s} {units[i]}")
    rows = page(rep, "source")
    hdr = rows[1]
    data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ssum = Counter()
    ops = Counter()
    tot = 0.0
    for r in data:
        for s in stalls:
            try:
                ssum[s] += float(r[ix[s]] or 0)
            except ValueError:
                pass
        src = r[ix["Source"]].split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        n = float(r[ix["Instructions Executed"]] or 0)
        ops[op.split(".")[0]] += n
        tot += n
    st = sum(ssum.values()) or 1.0
    print("stall mix (% of samples):", {k[6:]: round(100 * v / st, 1) for k, v in ssum.most_common(9)})
    print(f"warp instructions executed: {tot:.4g}" + (f"  = {tot * 32 / rp:.1f} thread-instr / read-position" if rp else ""))
    if rp:
        print("thread-instr / read-position by opcode:",
              {k: round(v * 32 / rp, 2) for k, v in ops.most_common(24)})


if __name__ == "__main__":
    main()
