#!/bin/bash
# 1 -> 8 GPU weak scaling on one box, launched exactly as the driver does (torchrun, --steps 20 --warmup 5).
set -u
o=gpurun_out
mkdir -p $o
tr() { local n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $o/r2s_n1_k20.json 2> $o/r2s.err
for n in 2 4 8; do tr $n --steps 20 --warmup 5 > $o/r2s_n${n}_k20.json 2>> $o/r2s.err; done
tr 8 > $o/r2s_n8.json 2>> $o/r2s.err
tr 8 --workload c4 --steps 2000 > $o/r2s_n8_c4.json 2>> $o/r2s.err
tr 8 --workload c3 --steps 2000 > $o/r2s_n8_c3.json 2>> $o/r2s.err
tail -c 300 $o/r2s.err
for f in $o/r2s_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[1], "UNREADABLE", e); sys.exit(0)
s = d.get("secondary") or {}
print(sys.argv[1].split("/")[-1], "N=%d" % d["n_gpus"], "value=%.4g" % d["value"], "ms=%.5f" % d["ms_per_step"], "e2e=%.4g" % d["e2e"]["value"],
      "rgb=%.4g" % (s.get("rgb_only") or {}).get("value", 0), "a2c=%.4g" % (s.get("a2c_pass") or {}).get("value", 0),
      "blocks=%s" % d["run_stats"]["blocks"], "coll_ms=%s" % d["run_stats"]["collective_ms"], d["clocks"])
PY
done
