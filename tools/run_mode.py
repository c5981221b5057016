#!/usr/bin/env python
"""Runs a few hundred device-resident steps of one configuration (profiling target for ncu -k regex:<kernel>).
Usage: python tools/run_mode.py [workload] [serial|pipelined] [float] [steps]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    serial = "serial" in sys.argv[2:]
    flt = "float" in sys.argv[2:]
    steps = next((int(a) for a in sys.argv[2:] if a.isdigit()), 400)
    _, world, n, layout = bench.make_workload(vn, wl)
    env = vn.GraphVecEnv(world, n, device="cuda:0", seed=1, max_episode_steps=900, obs_layout=layout, host_outputs=False,
                         scaled_float=flt)
    env.reset()
    acts = torch.randint(0, 4, (256, n), device="cuda", dtype=torch.int32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(steps):
        if i == steps // 2:
            e0.record()
        env.step_enqueue(acts[i % 256], actions_ready=not serial)
    e1.record()
    torch.cuda.synchronize()
    print("%s %s%s: %.2f us per step, mode of the last step %d" % (wl, "serial" if serial else "pipelined", " float" if flt else "",
                                                                   1e3 * e0.elapsed_time(e1) / (steps - steps // 2), env._prev_mode))


if __name__ == "__main__":
    main()
