#!/usr/bin/env python
"""A2C (n = 5) on BASELINE.json configs[0] - the 10 x 10 grid maze with MazeGraph frames - written against the
public API only, as a trainer in the style of deep_rl's A2C loop would use it:

    env.reset() / env.step(actions)          GraphVecEnv (device-resident: host_outputs=False)
    RolloutBuffer.start / step / returns      the step writes the rollout row; n-step returns on the device
    rollout.policy_input                      gather + TransposeImage + ScaledFloatFrame in one kernel
    env.set_hardness                          the reference's curriculum (thor_cached_auxiliary.py:68-70)
    env.episode_stats                         RewardCollector-style statistics, accumulated on the device

The policy is a deliberately small conv net (the reference's models/ are out of scope); the point is the data path:
no frame ever visits the host.  Needs a B200.  `python examples/a2c_maze.py --updates 400`
"""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class SmallPolicy(nn.Module):
    def __init__(self, n_actions=4):
        super().__init__()
        self.c1 = nn.Conv2d(3, 16, 8, stride=4)
        self.c2 = nn.Conv2d(16, 32, 4, stride=2)
        self.fc = nn.Linear(32 * 9 * 9, 128)
        self.pi = nn.Linear(128, n_actions)
        self.v = nn.Linear(128, 1)

    def forward(self, x):
        x = F.relu(self.c2(F.relu(self.c1(x))))
        x = F.relu(self.fc(x.flatten(1)))
        return self.pi(x), self.v(x).squeeze(1)


def build_env(vn, num_envs, seed, max_episode_steps=100):
    S, T = vn.scenes, vn.tables
    maze = S.random_maze((10, 10), 0.25, 0)
    goal = tuple(int(v) for v in np.argwhere(maze)[0])                     # graph/dungeon_graph.py:20 convention
    plain = S.GridScene(maze, [goal], False, (84, 84), ("rgb",))
    frames = S.render_maze_frames(plain, goal, (84, 84))                  # MazeGraph.render + GraphResize, hoisted
    scene = S.GridScene(maze, [goal], False, (84, 84), ("rgb",), explicit={"rgb": frames})
    world = T.compile_world([scene], T.SIMPLE_GRAPH)
    return vn.GraphVecEnv(world, num_envs, seed=seed, max_episode_steps=max_episode_steps, obs_layout="frame",
                          unreal_wrapper=False, host_outputs=False)


def train(updates=400, num_envs=16, n_step=5, gamma=0.99, lr=7e-4, seed=0, hardness=1.0, log_every=50, quiet=False):
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    torch.manual_seed(seed)
    env = build_env(vn, num_envs, seed)
    env.set_hardness(hardness)
    dev = env.device
    model = SmallPolicy().to(dev)
    opt = torch.optim.RMSprop(model.parameters(), lr=lr, alpha=0.99, eps=1e-5)   # thor_cached_auxiliary.py:31-37
    buf = vn.rollout.RolloutBuffer(env.dw, num_envs, n_step)
    env.reset()
    history = []
    t0 = time.perf_counter()
    for it in range(updates):
        buf.start(env)
        with torch.no_grad():
            for _ in range(n_step):
                logits, _ = model(vn.rollout.policy_input(env.dw, env.obs_state))
                a = torch.distributions.Categorical(logits=logits).sample().to(torch.int32)
                buf.step(env, a)                # env step + rollout row in the same launch
            _, last_v = model(vn.rollout.policy_input(env.dw, env.obs_state))
        returns = buf.returns(last_v, gamma)                                # [B, T]
        x = vn.rollout.policy_input(env.dw, buf.states[:-1].t().contiguous())      # [B, T, 3, 84, 84]
        logits, values = model(x.flatten(0, 1))
        dist = torch.distributions.Categorical(logits=logits)
        act = buf.actions.t().reshape(-1).long()
        adv = returns.reshape(-1) - values
        loss = -(dist.log_prob(act) * adv.detach()).mean() + 0.5 * adv.pow(2).mean() - 0.01 * dist.entropy().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        if (it + 1) % log_every == 0:
            st = env.episode_stats(reset=True)
            ep = max(1.0, st["episodes"])
            rec = dict(update=it + 1, episodes=int(st["episodes"]), success=st["successes"] / ep,
                       episode_length=st["length_sum"] / ep, reward=st["return_sum"] / ep,
                       fps=(it + 1) * n_step * num_envs / (time.perf_counter() - t0))
            history.append(rec)
            if not quiet:
                print(rec)
    env.close()
    return history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--updates", type=int, default=400)
    ap.add_argument("--envs", type=int, default=16)
    ap.add_argument("--hardness", type=float, default=1.0)
    args = ap.parse_args()
    train(args.updates, args.envs, hardness=args.hardness)
