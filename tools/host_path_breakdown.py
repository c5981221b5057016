#!/usr/bin/env python
"""Where the host-facing VecEnv.step() spends its time (development aid; needs a B200).
Prints microseconds per call of the pieces of GraphVecEnv.step on the bench workload."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    workload, world, n, layout = bench.make_workload(vn, sys.argv[1] if len(sys.argv) > 1 else "c2")
    gather = sys.argv[2] if len(sys.argv) > 2 else "auto"
    env = vn.GraphVecEnv(world, n, device="cuda:0", seed=2, max_episode_steps=900, obs_layout=layout, host_outputs=True,
                         gather=gather)
    print("workload %s, gather %s, %d envs, %d host sequence words" % (sys.argv[1] if len(sys.argv) > 1 else "c2", gather,
                                                                          n, env._seq_words))
    env.reset()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 4, (256, n)).astype(np.int32)
    for i in range(200):
        env.step(acts[i % 256])
    torch.cuda.synchronize()
    K = 3000
    t_async = t_wait = 0.0
    t0 = time.perf_counter()
    for i in range(K):
        a = time.perf_counter()
        env.step_async(acts[i % 256])
        b = time.perf_counter()
        env.step_wait()
        c = time.perf_counter()
        t_async += b - a
        t_wait += c - b
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print("step(): %.2f us total per call; step_async %.2f us, step_wait %.2f us" %
          (1e6 * total / K, 1e6 * t_async / K, 1e6 * t_wait / K))
    # pieces of step_wait with the GPU idle (event already complete)
    torch.cuda.synchronize()
    import ctypes as C
    L = vn.lib
    t0 = time.perf_counter()
    for _ in range(K):
        L.check(env.lib.vn_host_wait_seq(env._seq_host.data_ptr(), env._seq_words, env._seq, env._stream(), 1000000))
    print("vn_host_wait_seq on a published word: %.2f us" % (1e6 * (time.perf_counter() - t0) / K))
    t0 = time.perf_counter()
    for _ in range(K):
        (env._pack_np[:4 * n].copy(), env._pack_np[16 * n:17 * n].copy())
    print("reward + done copies: %.2f us" % (1e6 * (time.perf_counter() - t0) / K))
    t0 = time.perf_counter()
    for _ in range(K):
        env._obs()
    print("_obs(): %.2f us" % (1e6 * (time.perf_counter() - t0) / K))
    t0 = time.perf_counter()
    for _ in range(K):
        env._stream()
    print("_stream(): %.2f us" % (1e6 * (time.perf_counter() - t0) / K))
    t0 = time.perf_counter()
    for i in range(K):
        env.step_async(acts[i % 256])
        env._pending = False
    torch.cuda.synchronize()
    print("step_async back to back (no wait, includes GPU back-pressure): %.2f us" % (1e6 * (time.perf_counter() - t0) / K))


if __name__ == "__main__":
    main()
