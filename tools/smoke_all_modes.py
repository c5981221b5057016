#!/usr/bin/env python
"""Every launch mode of the step path on a small batch with frequent resets (memory-checker target:
compute-sanitizer --tool memcheck python tools/smoke_all_modes.py)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    T = vn.tables
    scene = vn.scenes.make_thor_scene(120, (14, 18), seed=5, n_goals=3, planes=("rgb", "depth", "segmentation"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    dw = vn.DeviceWorld(world)
    rng = np.random.RandomState(0)
    for n, gather, flt, host in ((23, "fused", False, True), (23, "persistent", False, True), (700, "bulk", False, True),
                                 (700, "ldg", False, False), (700, "persistent", False, False), (300, "auto", True, True),
                                 (2000, "auto", False, False)):
        env = vn.GraphVecEnv(world, n, seed=3, max_episode_steps=5, obs_layout="aux5", gather=gather, scaled_float=flt,
                             host_outputs=host, device_world=dw)
        env.reset()
        for t in range(12):
            a = rng.randint(0, 4, n).astype(np.int32)
            if host:
                env.step(a)
            else:
                env.step_enqueue(torch.from_numpy(a).cuda(), actions_ready=(t % 2 == 1))
        torch.cuda.synchronize()
        assert env.episode_stats()["resets"] > n
        rb = vn.rollout.RolloutBuffer(dw, n, 4)
        rb.start(env)
        for t in range(4):
            rb.step(env, torch.from_numpy(rng.randint(0, 4, n).astype(np.int32)).cuda())
        rb.targets(torch.zeros(n, device="cuda"), 0.99, torch.zeros(n, 400, device="cuda"), 0.9, 4, (20, 20))
        rb.auxiliary_targets(4, (20, 20))
        torch.cuda.synchronize()
        print("ok", n, gather, flt, host)


if __name__ == "__main__":
    main()
