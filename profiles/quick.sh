#!/bin/bash
# quick kernel-only timings for a list of workloads (no CPU baseline / e2e / Baum-Welch legs)
for w in "$@"; do
  timeout 180 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-bw 2>gpurun_out/quick_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$w', 'ms/step %.3f' % d['ms_per_step'], 'rp/s %.3e' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'GB/s %.0f' % d['roofline']['achieved'], d['clocks'])
"
done
