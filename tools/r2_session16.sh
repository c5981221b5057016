#!/bin/bash
set -u
o=gpurun_out
q() { python bench.py --steps 2000 --warmup 50 --quick "$@" 2>$o/r2_last.err | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); print('%.2f us  frac %.3f  launches/step %s' % (1e3*d['ms_per_step'], d['frac'], d['launches_per_step']))
except Exception as e:
    print('FAILED', t[:200]); print(open('$o/r2_last.err').read()[-1500:])"; }
python -m pytest tests -m gpu -x -q > $o/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2i_pytest.log; tail -3 $o/r2i_pytest.log
{
for n in 512 1024 4096 8192; do echo -n "c2 envs=$n auto pipelined : "; q --envs-per-gpu $n; echo -n "c2 envs=$n auto serial : "; q --envs-per-gpu $n --serial; done
echo -n "rgb auto pipelined : "; q --workload rgb; echo -n "rgb auto serial : "; q --workload rgb --serial
} 2>&1 | tee $o/r2i_auto_modes.txt
python tools/soak.py 100000 2>&1 | tee $o/r2i_soak.txt
