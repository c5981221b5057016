#!/bin/bash
set -u
o=gpurun_out
q() { python bench.py --steps 2000 --warmup 50 --quick "$@" 2>$o/r2_last.err | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print('%.2f us  iso %.2f  frac %.3f' % (1e3*d['ms_per_step'], 1e3*(d['iso'] or 0), d['frac']))
except Exception as e:
    print('FAILED', t[:200]); print(open('$o/r2_last.err').read()[-1500:])"; }
{
for g in 1036 1000 970 956 940 900 850 800 765 700; do echo -n "c2 grid=$g : "; VN_BULK_GRID=$g q --workload c2 --gather bulk; done
for g in 1480 1400 1300 1275 1250 1200 1100 1000 956 900; do echo -n "rgb grid=$g : "; VN_BULK_GRID=$g q --workload rgb --gather bulk; done
for n in 3108 4144 5180; do echo -n "c2 envs=$n : "; q --workload c2 --gather bulk --envs-per-gpu $n; done
} 2>&1 | tee $o/r2_grid_sweep.txt
