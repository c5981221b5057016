#!/bin/bash
# Round-2 GPU session 1: tests, the driver's bench command, launch-mode A/B, host path breakdown, steady-state DRAM counters.
set -u
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2a_pytest.log
tail -5 $o/r2a_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $o/r2a_bench_c2_k20.json 2> $o/r2a_bench_c2_k20.err; echo "bench rc=$?"
tail -c 600 $o/r2a_bench_c2_k20.err
for wl in c2 rgb; do for g in auto persistent; do
  echo -n "$wl $g : "; python bench.py --workload $wl --gather $g --steps 2000 --warmup 50 --quick 2>/dev/null | tail -1 | cut -c1-260
done; done 2>&1 | tee $o/r2a_ab.txt
for sp in 2 4; do echo -n "c2 auto VN_BULK_SPLIT=$sp : "; VN_BULK_SPLIT=$sp python bench.py --gather auto --steps 2000 --warmup 50 --quick 2>/dev/null | tail -1 | cut -c1-200; done 2>&1 | tee -a $o/r2a_ab.txt
for n in 512 1024 2048 16384; do for g in auto persistent; do
  echo -n "envs=$n $g : "; python bench.py --envs-per-gpu $n --gather $g --steps 2000 --warmup 50 --quick 2>/dev/null | tail -1 | cut -c1-200
done; done 2>&1 | tee -a $o/r2a_ab.txt
for wl in c2 rgb; do for g in auto persistent; do python tools/host_path_breakdown.py $wl $g; done; done > $o/r2a_host_path.txt 2>&1
cat $o/r2a_host_path.txt
