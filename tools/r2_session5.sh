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
python -m pytest tests/test_gpu_parity.py -x -q -k "golden or full_size or skipping or buffers or randomised" 2>&1 | tail -3
for wl in c2 rgb; do
echo -n "$wl bulk dynamic+prefetch : "; q --workload $wl --gather bulk
echo -n "$wl bulk static+prefetch : "; VN_BULK_DYNAMIC=0 q --workload $wl --gather bulk
echo -n "$wl bulk SPLIT=2 dynamic+prefetch : "; VN_BULK_SPLIT=2 q --workload $wl --gather bulk
echo -n "$wl bulk SPLIT=2 static+prefetch : "; VN_BULK_SPLIT=2 VN_BULK_DYNAMIC=0 q --workload $wl --gather bulk
echo -n "$wl persistent : "; q --workload $wl --gather persistent
done
echo -n "c3 dynamic : "; q --workload c3 --steps 300
echo -n "c3 static : "; VN_BULK_DYNAMIC=0 q --workload c3 --steps 300
echo -n "c4 dynamic : "; q --workload c4 --steps 300
echo -n "c4 static : "; VN_BULK_DYNAMIC=0 q --workload c4 --steps 300
echo -n "c2 h0.01 dynamic : "; q --workload c2 --hardness 0.01
echo -n "c2 h0.01 static : "; VN_BULK_DYNAMIC=0 q --workload c2 --hardness 0.01
} 2>&1 | tee $o/r2_prefetch.txt
