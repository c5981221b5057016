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
for g in 1 2 3 4; do for wl in c2 rgb; do echo -n "$wl groups=$g : "; VN_BULK_GROUPS=$g q --workload $wl --gather bulk; done; done
for g in 1 2; do echo -n "c4 groups=$g : "; VN_BULK_GROUPS=$g q --workload c4 --steps 300; done
} 2>&1 | tee $o/r2_groups.txt
