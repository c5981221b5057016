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
python -m pytest tests -m gpu -x -q > $o/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2c_pytest.log
tail -4 $o/r2c_pytest.log
{
for n in 300 444 512 768 1024; do for g in bulk persistent; do echo -n "envs=$n $g : "; q --envs-per-gpu $n --gather $g; done; done
} 2>&1 | tee $o/r2c_small.txt
for wl in c2 rgb; do python tools/host_path_breakdown.py $wl auto; done > $o/r2c_host_path.txt 2>&1
cat $o/r2c_host_path.txt
python bench.py > $o/r2c_bench_c2.json 2> $o/r2c_bench_c2.err; echo "bench rc=$?"; tail -c 400 $o/r2c_bench_c2.err
python tools/a2c_pass.py 30 > $o/r2_a2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file $o/r2_a2c_launches.csv \
    python tools/a2c_pass.py 30 > $o/r2_ncu_a2c.log 2>&1; echo "a2c rc=$?"
cat $o/r2_a2c_plain.log
