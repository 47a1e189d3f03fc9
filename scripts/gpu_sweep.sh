#!/bin/bash
# kernel-time sweep over shapes (experiments; prints the per-kernel ms from bench.py's roofline leg)
for shp in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape $shp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$shp', 'ms/step %.3f' % d['ms_per_step'], {k: round(v,4) for k,v in r['kernel_ms_per_step'].items()}, 'bwd TF/s alg %.0f' % r['achieved'], 'fwd %.0f' % r['fwd_kernel']['achieved'])"
done
