"""Evaluation services: what ``trainer.test()`` (test-train.py:26) needs from the
environment side - evaluation roll-outs of a policy with success rate, episode length and SPL - and the hardness
schedule the reference sketches (``scene_complexity = LinearSchedule(0.3, 1.0, 200000)``,
experiments/thor_cached_auxiliary.py:45).  Everything stays on the device; only the final numbers are read back.

SPL (success weighted by path length): mean over episodes of ``success * l / max(p, l)`` with ``l`` the
fewest actions from the episode's start state to its goal on the STATE graph (tables.optimal_policy_table) and
``p`` the number of actions the agent took.
"""
import torch


class LinearSchedule:
    """deep_rl.common.schedules.LinearSchedule as used by the experiments (thor_cached_auxiliary.py:38,45):
    linear from ``initial`` to ``final`` over ``steps`` time steps, then constant."""

    def __init__(self, initial, final, steps):
        self.initial, self.final, self.steps = float(initial), float(final), float(steps)

    def __call__(self, time_step):
        f = min(max(time_step / self.steps, 0.0), 1.0) if self.steps > 0 else 1.0
        return self.initial + (self.final - self.initial) * f


def apply_hardness_schedule(env, schedule, time_step):
    """``env.set_hardness(schedule(time_step))``; returns the value set."""
    h = schedule(time_step)
    env.set_hardness(h)
    return h


@torch.no_grad()
def evaluate(env, policy, episodes, max_steps=None):
    """Runs ``policy(obs) -> int32 CUDA actions [N]`` on ``env`` (a GraphVecEnv with ``host_outputs=False``) until
    ``episodes`` episodes have finished (the first ``episodes`` to finish are counted, like a validation run over a
    vectorised env).  Returns dict(episodes, success_rate, episode_length, reward, spl, truncated_rate, steps)."""
    if env.host_outputs:
        raise ValueError("evaluate() drives the device-resident interface: build the env with host_outputs=False")
    dev = env.device
    n = env.num_envs
    obs = env.reset()
    _, start = env.optimal_actions()
    start = start.clone().float()
    acc = torch.zeros(6, dtype=torch.float64, device=dev)      # episodes, successes, length, return, spl, truncated
    budget = torch.tensor(float(episodes), dtype=torch.float64, device=dev)
    steps = 0
    limit = max_steps if max_steps is not None else 1 << 62
    while steps < limit:
        obs, reward, done, _ = env.step(policy(obs))
        steps += 1
        # count finished episodes in env order until the budget is used up
        take = done & (torch.cumsum(done.to(torch.float64), 0) <= (budget - acc[0]))
        if bool(take.any()) or steps % 64 == 0:
            t = take.to(torch.float64)
            win = env.win.to(torch.float64) * t
            length = env.episode_length.to(torch.float64)
            acc[0] += t.sum()
            acc[1] += win.sum()
            acc[2] += (length * t).sum()
            acc[3] += (env.episode_return.to(torch.float64) * t).sum()
            acc[4] += (win * start.double() / torch.maximum(length, start.double()).clamp_min(1.0)).sum()
            acc[5] += ((env.truncated == 1).to(torch.float64) * t).sum()
            if float(acc[0]) >= episodes:
                break
        _, dist = env.optimal_actions()                         # envs that reset start a new episode here
        start = torch.where(done, dist.float(), start)
    e = max(float(acc[0]), 1.0)
    vals = acc.cpu().tolist()
    return dict(episodes=int(vals[0]), success_rate=vals[1] / e, episode_length=vals[2] / e, reward=vals[3] / e,
                spl=vals[4] / e, truncated_rate=vals[5] / e, steps=steps)
