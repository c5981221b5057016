#!/bin/bash
set -u
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2b_pytest.log
tail -5 $o/r2b_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $o/r2b_bench_c2_k20.json 2> $o/r2b_bench_c2_k20.err; echo "bench rc=$?"
tail -c 1500 $o/r2b_bench_c2_k20.err
for wl in c2 rgb; do python tools/host_path_breakdown.py $wl auto; done > $o/r2b_host_path.txt 2>&1
cat $o/r2b_host_path.txt
