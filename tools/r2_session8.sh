#!/bin/bash
set -u
o=gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity.py -x -q > $o/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2d_pytest.log
tail -4 $o/r2d_pytest.log
python tools/a2c_pass.py 60
python bench.py --workload c5 > $o/r2d_bench_c5.json 2> $o/r2d_bench_c5.err; echo "c5 rc=$?"; cat $o/r2d_bench_c5.json | cut -c1-900
