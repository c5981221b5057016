#!/bin/bash
set -u
o=gpurun_out
python -m pytest tests/test_gpu_rollout.py -x -q > $o/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2e_pytest.log
tail -4 $o/r2e_pytest.log
python tools/a2c_pass.py 60
A2C_SERIAL=1 python tools/a2c_pass.py 60
python bench.py --workload c5 > $o/r2e_bench_c5.json 2> $o/r2e_bench_c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('$o/r2e_bench_c5.json')); print(d['value'], d['run_stats'], d['roofline']['frac'])"
