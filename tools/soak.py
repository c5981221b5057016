#!/usr/bin/env python
"""Long device-resident runs with invariant checks (development aid; needs a B200): after every block of steps the
observation / goal batches must equal the store rows of the env states, in uint8 and float mode, pipelined, with row
skipping, across CUDA-graph replays.  Prints one line per configuration."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(steps=200000, block=1000):
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    scene = vn.scenes.make_thor_scene(400, (30, 30), seed=3, n_goals=4, planes=("rgb", "depth", "segmentation"))
    world = vn.compile_world([scene], vn.GYM_GRAPH)
    dw = vn.DeviceWorld(world)
    for name, n, kw in (("uint8 aux5 pipelined", 2048, dict(obs_layout="aux5")),
                        ("float rgbd_goal", 1024, dict(obs_layout="rgbd_goal", scaled_float=True)),
                        ("fused 200 envs", 200, dict(obs_layout="aux5")),
                        ("persistent launch 600 envs", 600, dict(obs_layout="rgbd_goal")),
                        ("persistent launch 3000 envs", 3000, dict(obs_layout="rgbd_goal", gather="persistent")),
                        ("auto, random serial / pipelined", 3000, dict(obs_layout="rgbd_goal")),
                        ("hardness 0.01 (reset-heavy)", 4096, dict(obs_layout="rgbd_goal"))):
        env = vn.GraphVecEnv(world, n, seed=11, max_episode_steps=37, host_outputs=False, device_world=dw, **kw)
        if "hardness" in name:
            env.set_hardness(0.01)
        env.reset()
        acts = torch.randint(0, 4, (1024, n), device="cuda", dtype=torch.int32)
        t0 = time.perf_counter()
        for b0 in range(0, steps, block):
            mixed = name.startswith("auto, random")
            flips = torch.rand(block).tolist() if mixed else None
            for i in range(block):
                env.step_enqueue(acts[(b0 + i) % 1024], actions_ready=(flips[i] < 0.5) if mixed else True)
            torch.cuda.synchronize()
            s, g = env.state.long(), env.goal.long()
            if env.scaled_float:
                want = vn.rollout.policy_input(dw, env.state, "rgb")
                assert torch.equal(env.float_buf["rgb"], want)
                assert torch.equal(env.float_buf["goal_rgb"], vn.rollout.policy_input(dw, env.goal, "rgb"))
            else:
                for p, buf in env.obs_buf.items():
                    assert torch.equal(buf, dw.plane_view(p)[s]), (name, p, b0)
                for p, buf in env.goal_buf.items():
                    assert torch.equal(buf, dw.plane_view(p)[g]), (name, p, b0)
            assert torch.equal(env.goal, dw.task_goal[env.task.long()])
        st = env.episode_stats()
        assert st["steps"] == steps * n and st["episodes"] == st["resets"] - n
        print("%-28s %d envs x %d steps ok: %.0f episodes, %.1f %% rows skipped, %.1f M env-steps/s incl. checks"
              % (name, n, steps, st["episodes"], 100 * st["rows_skipped"] / st["steps"],
                 steps * n / (time.perf_counter() - t0) / 1e6))


def host_twin(steps=30000):
    """The host-facing step() (numpy in / numpy out, mapped pinned memory, sequence words, alternating host packs, lazy
    infos) against a device-resident twin, every step."""
    import numpy as np
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    scene = vn.scenes.make_thor_scene(400, (30, 30), seed=3, n_goals=4, planes=("rgb", "depth"))
    world = vn.compile_world([scene], vn.GYM_GRAPH)
    n = 4096
    host = vn.GraphVecEnv(world, n, seed=11, max_episode_steps=37, obs_layout="rgbd_goal")
    dev = vn.GraphVecEnv(world, n, seed=11, max_episode_steps=37, obs_layout="rgbd_goal", host_outputs=False, device_world=host.dw)
    host.reset()
    dev.reset()
    acts = np.random.RandomState(0).randint(0, 4, (256, n)).astype(np.int32)
    dacts = torch.from_numpy(acts).cuda()
    t0 = time.perf_counter()
    kept = []
    for t in range(steps):
        _, r, d, infos = host.step(acts[t % 256])
        dev.step_enqueue(dacts[t % 256], actions_ready=True)
        if t % 50 == 0:
            torch.cuda.synchronize()
            assert np.array_equal(r.view(np.uint32), dev.reward.cpu().numpy().view(np.uint32)), t
            assert np.array_equal(d, dev.done.cpu().numpy().astype(bool)) and torch.equal(host.state, dev.state), t
            kept.append((infos, [dict(x) for x in infos[:64]]))
            kept = kept[-8:]
    for infos, want in kept:                      # lazy infos kept across many later steps are still right
        assert [dict(x) for x in infos[:64]] == want
    assert host.episode_stats() == dev.episode_stats()
    print("host-facing step vs device twin  %d envs x %d steps ok, %.1f M env-steps/s incl. the twin and checks"
          % (n, steps, steps * n / (time.perf_counter() - t0) / 1e6))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000)
    host_twin()
