#!/bin/bash
set -u
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -q > $o/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2f_pytest.log; tail -3 $o/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/bench_all.sh r2f
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
run() { local name=$1 skip=$2 cnt=$3 k=$4; shift 4
  python bench.py "$@" > $o/r2f_plain_$name.log 2>&1 &&
  ncu --cache-control none --clock-control none --metrics $M -k regex:$k -s $skip -c $cnt --csv \
      --log-file $o/r2f_traffic_$name.csv python bench.py "$@" > $o/r2f_ncu_$name.log 2>&1
  echo "$name rc=$? lines=$(wc -l < $o/r2f_traffic_$name.csv)"; }
run c2 400 60 vn_gather_bulk --workload c2 --quick --steps 200 --warmup 20 --mix 300
run rgb 400 60 vn_gather_bulk --workload rgb --quick --steps 200 --warmup 20 --mix 300
run c3 100 60 vn_gather_bulk --workload c3 --quick --steps 100 --warmup 10 --mix 100
run c4 100 60 vn_gather_bulk --workload c4 --quick --steps 100 --warmup 10 --mix 100
python bench.py --workload c2 --quick --steps 200 --warmup 20 --mix 300 > $o/r2f_plain_ll.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $o/r2f_launches_bench_steady.csv \
    python bench.py --workload c2 --quick --steps 200 --warmup 20 --mix 300 > $o/r2f_ncu_ll.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vn_gather_bulk -s 400 -c 3 -o $o/r2f_gather_full \
    python bench.py --workload c2 --quick --steps 200 --warmup 20 --mix 300 > $o/r2f_ncu_full.log 2>&1; echo "full rc=$?"
python tools/a2c_pass.py 30 > $o/r2f_a2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file $o/r2f_a2c_launches.csv \
    python tools/a2c_pass.py 30 > $o/r2f_ncu_a2c.log 2>&1; echo "a2c rc=$?"
