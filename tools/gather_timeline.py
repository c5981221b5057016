#!/usr/bin/env python
"""Per-CTA timeline of the bulk gather in the steady-state device-resident loop (development aid; needs a B200):
when do its CTAs become resident, when does the predecessor complete, how long are the first / last units, when does the
last CTA drain.  Usage: python tools/gather_timeline.py [workload] """
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    _, world, n, layout = bench.make_workload(vn, wl)
    env = vn.GraphVecEnv(world, n, device="cuda:0", seed=1, max_episode_steps=900, obs_layout=layout, host_outputs=False,
                         gather="bulk")
    env.reset()
    acts = torch.randint(0, 4, (512, n), device="cuda", dtype=torch.int32)
    for i in range(1500):
        env.step_enqueue(acts[i % 512], actions_ready=True)
    torch.cuda.synchronize()
    K, G = 40, 148 * 32
    traces = torch.zeros((K, G, 16), dtype=torch.int64, device="cuda")
    lib = env.lib
    for k in range(K):
        lib.vn_debug_gather_trace(traces[k].data_ptr())
        env.step_enqueue(acts[k % 512], actions_ready=True)
    lib.vn_debug_gather_trace(None)
    torch.cuda.synchronize()
    t = traces.cpu().numpy()
    print("workload %s: %d envs; per-launch medians over %d steady-state launches (us, relative to the first CTA's residency)" % (wl, n, K - 5))
    rows = []
    for k in range(5, K):
        a = t[k]
        a = a[a[:, 0] > 0]
        t0 = a[:, 0].min()
        prev_end = t[k - 1][t[k - 1][:, 0] > 0][:, 4].max()
        rows.append(dict(grid=len(a), resident_last=(a[:, 0].max() - t0) / 1e3, pred_done_first=(a[:, 1].min() - t0) / 1e3,
                         pred_done_last=(a[:, 1].max() - t0) / 1e3, first_unit_median=np.median(a[:, 2] - a[:, 1]) / 1e3,
                         last_issue_median=(np.median(a[:, 3]) - t0) / 1e3, drain_first=(a[:, 4].min() - t0) / 1e3,
                         drain_median=(np.median(a[:, 4]) - t0) / 1e3, drain_last=(a[:, 4].max() - t0) / 1e3,
                         units_mean=a[:, 5].mean(), units_max=a[:, 5].max(),
                         gap_prev_end_to_start=(a[:, 1].min() - prev_end) / 1e3,
                         period=(a[:, 4].max() - prev_end) / 1e3))
    for key in rows[0]:
        v = np.array([r[key] for r in rows], float)
        print("  %-24s median %8.2f   min %8.2f   max %8.2f" % (key, np.median(v), v.min(), v.max()))
    # one launch in detail: unit durations by position, by SM, and who finishes last
    a = t[K - 3]
    a = a[a[:, 0] > 0]
    start = a[:, 1].min()
    MASK = (1 << 63) - 1
    ends = (a[:, 8:16].astype(np.uint64) & np.uint64(MASK)).astype(np.int64)
    empty = (a[:, 8:16].astype(np.uint64) >> np.uint64(63)).astype(bool)
    nun = np.minimum(a[:, 5], 8)
    prev = np.concatenate([a[:, 1:2], ends[:, :-1]], 1)
    dur = (ends - prev) / 1e3
    print("  one launch: unit duration (us) by position in the CTA's sequence, real copies only")
    for j in range(6):
        m = (nun > j) & ~empty[:, j]
        if m.sum():
            print("    unit %d: n=%4d  median %6.2f  p10 %6.2f  p90 %6.2f  max %6.2f   issued at median %6.2f us" %
                  (j, m.sum(), np.median(dur[m, j]), np.percentile(dur[m, j], 10), np.percentile(dur[m, j], 90), dur[m, j].max(),
                   np.median(prev[m, j] - start) / 1e3))
    fin = (a[:, 4] - start) / 1e3
    real = np.array([(~empty[i, :nun[i]]).sum() for i in range(len(a))])
    print("  finish time (us after start) by number of real units: " +
          "  ".join("%d units: n=%d median %.1f max %.1f" % (u, (real == u).sum(), np.median(fin[real == u]), fin[real == u].max())
                    for u in sorted(set(real))))
    sm = a[:, 6]
    per_sm = np.array([fin[sm == k].max() for k in sorted(set(sm))])
    print("  per-SM last finish: min %.1f  median %.1f  max %.1f  (SMs: %d); CTAs per SM: %s" %
          (per_sm.min(), np.median(per_sm), per_sm.max(), len(per_sm), sorted(set(np.bincount(sm.astype(int))))))
    order = np.argsort(fin)[-8:]
    print("  the 8 last CTAs: " + "  ".join("sm%d fin %.1f units %d real %d" % (sm[i], fin[i], a[i, 5], real[i]) for i in order))


if __name__ == "__main__":
    main()
