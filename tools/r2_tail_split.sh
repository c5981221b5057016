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
python -m pytest tests/test_gpu_parity.py -x -q -k "full_size_properties_c2 or skipping or buffers" 2>&1 | tail -2
for ts in 1 2 3 4; do for w in 16 8; do echo -n "c2 tail=$ts waves16=$w : "; VN_BULK_TAIL_SPLIT=$ts VN_BULK_TAIL_WAVES16=$w q --workload c2 --gather bulk; done; done
for ts in 1 2; do echo -n "rgb tail=$ts : "; VN_BULK_TAIL_SPLIT=$ts q --workload rgb --gather bulk; done
for ts in 1 2; do echo -n "c4 tail=$ts : "; VN_BULK_TAIL_SPLIT=$ts q --workload c4 --gather bulk --steps 300; done
} 2>&1 | tee $o/r2_tail_split.txt
