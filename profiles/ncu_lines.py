#!/usr/bin/env python
"""Top source lines of an .ncu-rep by stall samples / executed instructions.
    python profiles/ncu_lines.py prof.ncu-rep <read_positions_per_launch> [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; rp = float(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; acc = []
for r in rows:
    if len(r) < 8:
        if r and r[0] == "File Path": cur = r[1].split("/")[-1]
        continue
    if r[0] not in ("", "Line No") and r[2] == "-":
        try: acc.append((cur, int(r[0]), r[1].strip()[:86], float(r[4] or 0), float(r[7] or 0)))
        except ValueError: pass
ts = sum(a[3] for a in acc) or 1; ti = sum(a[4] for a in acc)
print(f"total warp inst {ti:.4g} = {ti*32/rp:.1f} thread-instr/rp")
for a in sorted(acc, key=lambda a: -a[4])[:n]:
    print(f"{a[0]:20s} L{a[1]:4d} inst {a[4]*32/rp:6.2f}/rp samp {100*a[3]/ts:5.1f}%  {a[2]}")
