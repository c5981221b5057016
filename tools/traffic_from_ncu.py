#!/usr/bin/env python
"""gpurun_out/r2_traffic_<workload>.csv (ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum over >= 50 consecutive steady-state launches, tools/r2_session3_ncu.sh) -> profiles/roofline_traffic.json:
median DRAM bytes read / written per launch of the dominant kernel of every workload, what bench.py reports as
`roofline.traffic` (and `roofline.dram_frac` = traffic / in-situ kernel time / measured peak)."""
import collections
import csv
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1,
        "msecond": 1e3}


def parse(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)
    kernel = None
    for row in csv.DictReader(lines):
        per[row["ID"]][row["Metric Name"]] = float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
        kernel = row["Kernel Name"].split("(")[0].replace("void ", "").replace("vn::", "")
    rd = np.array([m["dram__bytes_read.sum"] for m in per.values()])
    wr = np.array([m["dram__bytes_write.sum"] for m in per.values()])
    us = np.array([m["gpu__time_duration.sum"] for m in per.values()])
    return kernel, rd, wr, us


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
    out = {"how": "ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                  "gpu__time_duration.sum -k <kernel> -s <warm> -c 60 over bench.py --workload <w> --quick: 60 CONSECUTIVE "
                  "steady-state launches, no cache flush between them (tools/r2_session3_ncu.sh); medians.  Under ncu the "
                  "launches are serialised (no overlap of the scalar kernel with the previous gather), so `kernel_us_under_ncu` "
                  "is longer than the in-situ period bench.py times; the BYTES are what carries over.",
           "workloads": {}}
    for w in ("c2", "rgb", "c3", "c4"):
        path = os.path.join(src, "%s_traffic_%s.csv" % (tag, w))
        if not os.path.exists(path):
            continue
        kernel, rd, wr, us = parse(path)
        out["workloads"][w] = {
            "kernel": kernel, "launches": int(len(rd)),
            "dram_bytes_read_per_launch": float(np.median(rd)), "dram_bytes_write_per_launch": float(np.median(wr)),
            "dram_bytes_read_min_max": [float(rd.min()), float(rd.max())],
            "dram_bytes_write_min_max": [float(wr.min()), float(wr.max())],
            "kernel_us_under_ncu": float(np.median(us)),
            "dram_gbs_under_ncu": float(np.median(rd + wr) / np.median(us) / 1e3),
            "source": "profiles/%s_traffic_%s.csv" % (tag, w)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    for w, d in out["workloads"].items():
        print("%-4s %-24s read %.1f MB  write %.1f MB  %.2f us under ncu  %.0f GB/s" %
              (w, d["kernel"], d["dram_bytes_read_per_launch"] / 1e6, d["dram_bytes_write_per_launch"] / 1e6,
               d["kernel_us_under_ncu"], d["dram_gbs_under_ncu"]))


if __name__ == "__main__":
    main()
