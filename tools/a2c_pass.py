#!/usr/bin/env python
"""One A2C / UNREAL data pass in a loop (profiling aid): 20 env steps that write their own rollout rows, n-step returns,
pixel-control returns (rewards + back-up in one pass), reward-prediction labels - the work of bench.py's
`secondary.a2c_pass`.  Usage: python tools/a2c_pass.py [passes]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    _, world, n, layout = bench.make_workload(vn, "c2")
    dev = torch.device("cuda", 0)
    env = vn.GraphVecEnv(world, n, device=dev, seed=1, max_episode_steps=900, obs_layout=layout, host_outputs=False)
    env.reset()
    T = 20
    acts = torch.randint(0, 4, (256, n), device=dev, dtype=torch.int32)
    rb = vn.rollout.RolloutBuffer(env.dw, n, T)
    v, q = torch.randn(n, device=dev), torch.rand(n, 400, device=dev)
    vn.rollout.target_tables(env.dw, 4, (20, 20))
    for i in range(200):
        env.step_enqueue(acts[i % 256], actions_ready=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for p in range(passes):
        if p == 5:
            e0.record()
        rb.start(env)
        for t in range(T):
            rb.step(env, acts[(p * T + t) % 256], actions_ready=True)
        if os.environ.get("A2C_SERIAL"):
            rb.returns(v, 0.99)
            rb.pixel_control_returns(q, 0.9, 4, (20, 20))
            rb.reward_prediction()
        else:
            rb.targets(v, 0.99, q, 0.9, 4, (20, 20), overlap=bool(os.environ.get("A2C_OVERLAP")))
    e1.record()
    torch.cuda.synchronize()
    print("a2c pass: %.1f us (%d passes)" % (1e3 * e0.elapsed_time(e1) / max(1, passes - 5), passes))


if __name__ == "__main__":
    main()
