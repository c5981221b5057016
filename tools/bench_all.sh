#!/bin/bash
# Runs every bench line of DESIGN.md section 5 on one B200 and leaves the JSON lines in gpurun_out/ (tag = $1, default r2).
set -u
t=${1:-r2}
o=gpurun_out
mkdir -p $o
python bench.py > $o/${t}_bench_c2.json 2> $o/${t}_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $o/${t}_bench_c2_k20.json 2>> $o/${t}_bench.err
python bench.py --hardness 0.01 --no-cpu-baseline > $o/${t}_bench_c2_h001.json 2>> $o/${t}_bench.err
python bench.py --workload rgb --no-cpu-baseline > $o/${t}_bench_rgb.json 2>> $o/${t}_bench.err
python bench.py --workload c1 --no-cpu-baseline > $o/${t}_bench_c1.json 2>> $o/${t}_bench.err
python bench.py --workload c1 --cuda-graph --no-cpu-baseline > $o/${t}_bench_c1_cuda_graph.json 2>> $o/${t}_bench.err
python bench.py --workload c3 --steps 3000 --no-cpu-baseline > $o/${t}_bench_c3.json 2>> $o/${t}_bench.err
python bench.py --workload c4 --steps 2000 --no-cpu-baseline > $o/${t}_bench_c4.json 2>> $o/${t}_bench.err
python bench.py --workload c5 > $o/${t}_bench_c5.json 2>> $o/${t}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $o/${t}_bench_reference.json 2>> $o/${t}_bench.err
tail -c 400 $o/${t}_bench.err
for f in $o/${t}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[1], "UNREADABLE", e); sys.exit(0)
r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value=%.4g" % d["value"], "ms=%.4g" % d["ms_per_step"], "e2e=%.4g" % (d.get("e2e") or {}).get("value", 0),
      "frac=%.3f" % (r.get("frac") or 0), "dram_frac=%s" % r.get("dram_frac"), "iso=%.3f" % ((r.get("isolated") or {}).get("frac") or 0),
      (d.get("cpu_baseline") or {}).get("value"))
PY
done
