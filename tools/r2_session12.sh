#!/bin/bash
set -u
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $o/r2h_pytest.log; tail -3 $o/r2h_pytest.log
python bench.py --workload rgb --no-cpu-baseline > $o/r2h_bench_rgb.json 2>$o/r2h.err; python -c "
import json; d=json.load(open('$o/r2h_bench_rgb.json')); print('rgb value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6))"
VN_NO_PARAM_ACTIONS=1 python bench.py --workload rgb --no-cpu-baseline > $o/r2h_bench_rgb_nopa.json 2>>$o/r2h.err; python -c "
import json; d=json.load(open('$o/r2h_bench_rgb_nopa.json')); print('rgb (actions read from pinned memory) value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6))"
python bench.py --no-cpu-baseline > $o/r2h_bench_c2.json 2>>$o/r2h.err; python -c "
import json; d=json.load(open('$o/r2h_bench_c2.json')); print('c2 value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v['value']/1e6,2) for k,v in d['e2e']['variants'].items()}, d['secondary']['reference_run_config']['us_per_vector_step'])"
python bench.py --workload c1 --no-cpu-baseline > $o/r2h_bench_c1.json 2>>$o/r2h.err; python -c "
import json; d=json.load(open('$o/r2h_bench_c1.json')); print('c1 value %.2f M e2e %.2f M' % (d['value']/1e6, d['e2e']['value']/1e6))"
