#!/bin/bash
set -u
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2g_pytest.log; tail -3 $o/r2g_pytest.log
for wl in c2 rgb; do python tools/host_path_breakdown.py $wl auto; done 2>&1 | tee $o/r2g_host_path.txt
python bench.py --workload rgb --no-cpu-baseline > $o/r2g_bench_rgb.json 2>$o/r2g.err; python -c "
import json; d=json.load(open('$o/r2g_bench_rgb.json')); print('rgb value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6))"
python bench.py --no-cpu-baseline > $o/r2g_bench_c2.json 2>>$o/r2g.err; python -c "
import json; d=json.load(open('$o/r2g_bench_c2.json')); print('c2 value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v['value']/1e6,2) for k,v in d['e2e']['variants'].items()}, d['secondary']['reference_run_config']['us_per_vector_step'])"
