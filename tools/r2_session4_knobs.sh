#!/bin/bash
set -u
o=gpurun_out
q() { python bench.py --steps 2000 --warmup 50 --quick "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.2f us  iso %.2f  frac %.3f' % (1e3*d['ms_per_step'], 1e3*(d['iso'] or 0), d['frac']))"; }
{
for wl in c2 rgb; do
for ps in 7 6 5 4 3; do echo -n "$wl bulk PER_SM=$ps : "; VN_BULK_PER_SM=$ps q --workload $wl --gather bulk; done
for sp in 2 3; do for ps in 20 14 10 8; do echo -n "$wl bulk SPLIT=$sp PER_SM=$ps : "; VN_BULK_SPLIT=$sp VN_BULK_PER_SM=$ps q --workload $wl --gather bulk; done; done
echo -n "$wl ldg : "; q --workload $wl --gather ldg
echo -n "$wl bulk static tickets : "; VN_BULK_DYNAMIC=0 q --workload $wl --gather bulk
echo -n "$wl bulk hints 0 : "; VN_BULK_L2_HINTS=0 q --workload $wl --gather bulk
echo -n "$wl bulk no PDL : "; VN_NO_PDL=1 q --workload $wl --gather bulk
done
} 2>&1 | tee $o/r2_knobs.txt
