#!/bin/bash
# Runs every bench line of DESIGN.md section 5 on one B200 and leaves the JSON lines in gpurun_out/.
set -u
o=gpurun_out
python bench.py > $o/r1b_bench_c2.json 2> $o/r1b_bench_c2.err
python bench.py --hardness 0.01 --no-cpu-baseline > $o/r1b_bench_c2_h001.json 2>> $o/r1b_bench_c2.err
python bench.py --workload rgb --no-cpu-baseline > $o/r1b_bench_rgb.json 2>> $o/r1b_bench_c2.err
python bench.py --workload c1 --no-cpu-baseline > $o/r1b_bench_c1.json 2>> $o/r1b_bench_c2.err
python bench.py --workload c1 --cuda-graph --no-cpu-baseline > $o/r1b_bench_c1_cuda_graph.json 2>> $o/r1b_bench_c2.err
python bench.py --workload c3 --steps 3000 --no-cpu-baseline > $o/r1b_bench_c3.json 2>> $o/r1b_bench_c2.err
python bench.py --workload c4 --steps 2000 --no-cpu-baseline > $o/r1b_bench_c4.json 2>> $o/r1b_bench_c2.err
python bench.py --workload c5 > $o/r1b_bench_c5.json 2>> $o/r1b_bench_c2.err
python bench.py --impl reference > $o/r1b_bench_reference.json 2>> $o/r1b_bench_c2.err
tail -c 400 $o/r1b_bench_c2.err
for f in $o/r1b_bench_*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value=%.4g" % d["value"], "ms=%.4g" % d["ms_per_step"], "e2e=%.4g" % (d.get("e2e") or {}).get("value", 0),
      "frac=%.3f" % r.get("frac", 0), "iso=%.3f" % (r.get("isolated") or {}).get("frac", 0), (d.get("cpu_baseline") or {}).get("value"))
PY
done
