#!/bin/bash
set -u
o=gpurun_out
q() { python bench.py --steps 2000 --warmup 50 --quick "$@" 2>$o/r2_last.err | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print('%.2f us  frac %.3f' % (1e3*d['ms_per_step'], d['frac']))
except Exception as e:
    print('FAILED', t[:200]); print(open('$o/r2_last.err').read()[-1500:])"; }
{
for n in 1024 2048 4096 8192 16384; do for g in bulk persistent; do
echo -n "c2 envs=$n $g serial : "; q --envs-per-gpu $n --gather $g --serial
done; done
for g in bulk persistent; do echo -n "rgb 4096 $g serial : "; q --workload rgb --gather $g --serial; done
} 2>&1 | tee $o/r2_serial_modes.txt
