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
python -m pytest tests/test_gpu_parity.py -x -q -k "golden or full_size or skipping or buffers or randomised or graph or pipelined or third or multiple_of_16 or trainer" 2>&1 | tail -3
for w in 0 1; do for wl in c2 aux5 c4; do echo -n "$wl VN_NO_WHOLE_RECORD=$w : "; VN_NO_WHOLE_RECORD=$w q --workload $wl --gather bulk --steps 1000; done; done
for w in 0 1; do echo -n "c2 512 envs persistent VN_NO_WHOLE_RECORD=$w : "; VN_NO_WHOLE_RECORD=$w q --envs-per-gpu 512 --gather persistent; done
} 2>&1 | tee $o/r2_whole_record.txt
