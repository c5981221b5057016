#!/bin/bash
# Steady-state DRAM counters (no cache flush between launches), one --set full capture, and the A2C-pass launch list.
set -u
o=gpurun_out
mkdir -p $o
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
run() {  # name, skip, count, kernel regex, bench args...
  local name=$1 skip=$2 cnt=$3 k=$4; shift 4
  python bench.py "$@" > $o/r2_plain_$name.log 2>&1 &&
  ncu --cache-control none --clock-control none --metrics $M -k regex:$k -s $skip -c $cnt --csv \
      --log-file $o/r2_traffic_$name.csv python bench.py "$@" > $o/r2_ncu_$name.log 2>&1
  echo "$name rc=$? lines=$(wc -l < $o/r2_traffic_$name.csv)"
}
run c2 400 60 vn_gather_bulk --workload c2 --quick --steps 200 --warmup 20 --mix 300
run rgb 400 60 vn_gather_bulk --workload rgb --quick --steps 200 --warmup 20 --mix 300
run c3 100 60 vn_gather_bulk --workload c3 --quick --steps 100 --warmup 10 --mix 100
run c4 100 60 vn_gather_bulk --workload c4 --quick --steps 100 --warmup 10 --mix 100
run c2_step 400 60 vn_step_kernel --workload c2 --quick --steps 200 --warmup 20 --mix 300
# the gather with full sections (3 launches, cache control left at ncu's default = flush: cold-cache view)
ncu --set full --clock-control none --import-source on -k regex:vn_gather_bulk -s 400 -c 3 -o $o/r2_gather_full \
    python bench.py --workload c2 --quick --steps 200 --warmup 20 --mix 300 > $o/r2_ncu_full.log 2>&1; echo "full rc=$?"
# launch list of the A2C data pass
python tools/a2c_pass.py 30 > $o/r2_a2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 700 --csv --log-file $o/r2_a2c_launches.csv \
    python tools/a2c_pass.py 30 > $o/r2_ncu_a2c.log 2>&1; echo "a2c rc=$?"
cat $o/r2_a2c_plain.log
