"""ORACLE (test infrastructure).  The vectorised-env conventions that live in UN-VENDORED
third-party code: gym 0.15.7 ``TimeLimit``, baselines-0.1.6-style ``SubprocVecEnv`` auto-reset,
and deep_rl 0.2.9's ``RewardCollector`` / ``TransposeImage`` / ``ScaledFloatFrame`` /
``UnrealEnvBaseWrapper`` (call sites: /root/reference/experiments/thor_cached_auxiliary.py:58-71,
environments/gym_graph/__init__.py:12,21,27).

PARITY UNPINNED for this file: none of that source is under /root/reference and it is not
installed offline; the behaviour below is restated from the published versions pinned in
not_explicit_list.txt:11,29,45.  The env classes it drives ARE pinned.
"""
import numpy as np

from . import graph_util as gu
from . import philox


class TimeLimit:
    """gym 0.15.7 wrappers/time_limit.py: after ``max_episode_steps`` steps
    ``info['TimeLimit.truncated'] = not done; done = True``; counter cleared by reset()."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max = max_episode_steps
        self._elapsed = 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self._elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed += 1
        if self._elapsed >= self._max:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info


class RewardCollector:
    """deep_rl.common.env.RewardCollector [recalled]: accumulates the episode return and length and
    on done sets ``info['episode'] = {'r': return, 'l': length}``."""

    def __init__(self, env):
        self.env = env
        self.ret, self.len = 0.0, 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self.ret, self.len = 0.0, 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.ret = float(np.float32(np.float32(self.ret) + np.float32(reward)))
        self.len += 1
        if done:
            info["episode"] = dict(r=self.ret, l=self.len)
        return obs, reward, done, info


def last_action_reward(action, reward, n_actions):
    """deep_rl.a2c_unreal.util.UnrealEnvBaseWrapper [recalled]: one_hot(action) ++ [clip(r,-1,1)],
    float32 [n_actions + 1]; all zeros after reset."""
    v = np.zeros(n_actions + 1, np.float32)
    if action is not None and 0 <= action < n_actions:
        v[action] = 1.0
    v[-1] = np.clip(reward, -1.0, 1.0)
    return v


def transpose_scale(frame_u8):
    """TransposeImage then ScaledFloatFrame [recalled]: HWC uint8 -> CHW float32 / 255."""
    return np.transpose(frame_u8, (2, 0, 1)).astype(np.float32) / np.float32(255.0)


class VecEnv:
    """baselines-style vectorised env [recalled, D1]: per worker
    ``ob, r, done, info = env.step(a); if done: ob = env.reset()`` - the returned observation is the
    FIRST observation of the next episode; leaves are stacked on axis 0; ``lar`` is the
    UnrealEnvBaseWrapper vector (zeros for an env that just reset)."""

    def __init__(self, envs, n_actions=4):
        self.envs = envs
        self.n_actions = n_actions

    @staticmethod
    def _stack(obs_list):
        first = obs_list[0]
        if isinstance(first, tuple):
            return tuple(np.stack([o[i] for o in obs_list]) for i in range(len(first)))
        if isinstance(first, dict):
            return {k: np.stack([o[k] for o in obs_list]) for k in first}
        return np.stack(obs_list)

    def reset(self):
        obs = [e.reset() for e in self.envs]
        lar = np.zeros((len(self.envs), self.n_actions + 1), np.float32)
        return self._stack(obs), lar

    def step(self, actions):
        obs, rews, dones, infos, lars = [], [], [], [], []
        for e, a in zip(self.envs, actions):
            a = int(a)
            ob, r, d, info = e.step(None if a < 0 else a)
            if d:
                ob = e.reset()
                lars.append(np.zeros(self.n_actions + 1, np.float32))
            else:
                lars.append(last_action_reward(a, r, self.n_actions))
            obs.append(ob)
            rews.append(r)
            dones.append(d)
            infos.append(info)
        return (self._stack(obs), np.stack(lars)), np.array(rews, np.float32), np.array(dones, bool), infos


# --------------------------------------------------------------------------- Philox reset source
class PhiloxResetSource:
    """CPU restatement of the DEVICE path's own reset sampling (csrc/vn_kernels.cu: reset_env),
    built from the oracle's candidate lists (reference order, graph/util.py:119-143 / :88-117),
    not from the product's compiled tables:

      draws = philox4x32_10(key = seed, ctr = (global_env_id, epoch, 0, 0)); epoch += 1
      task  = mulhi(draws[0], n_tasks_of_env)
      candidates of the task sorted (stable) by curriculum distance;
      oriented:      k = #(dist <= optimal_distance) (all if no curriculum); idx = mulhi(draws[2], k)
      un-oriented:   P = #(dist <= od), Q = rest; bucket = P if (Q == 0 or draws[1] < 0.9*2^32) else Q;
                     idx = mulhi(draws[2], |bucket|) (+ P for the far bucket)

    This is the reference's distribution (uniform over the eligible set; 0.9/0.1 two-level for
    sample_initial_position) with a different, reproducible source of randomness.
    """

    def __init__(self, seed, env_id, tasks, optimal_distance_fn, two_level):
        """tasks: list of (potentials, dists) in reference order, one per task of this env."""
        self.seed, self.env_id, self.epoch = seed, env_id, 0
        self.sorted = []
        for pots, dists in tasks:
            order = np.argsort(np.asarray(dists), kind="stable")
            self.sorted.append(([pots[i] for i in order], np.asarray(dists)[order]))
        self.od = optimal_distance_fn
        self.two_level = two_level

    def __call__(self):
        d = philox.reset_draws(self.seed, self.env_id, self.epoch)
        self.epoch += 1
        t = int(philox.mulhi(d[0], len(self.sorted)))
        pots, dists = self.sorted[t]
        od = self.od(t)
        n = len(pots)
        p = n if od is None else int(np.searchsorted(dists, od, side="right"))
        if self.two_level and od is not None:
            q = n - p
            if q == 0 or int(d[1]) < philox.BUCKET_THRESHOLD:
                idx = int(philox.mulhi(d[2], p))
            else:
                idx = p + int(philox.mulhi(d[2], q))
        else:
            idx = int(philox.mulhi(d[2], p))
        return t, pots[idx]


class ReferenceStyleResetSource:
    """Reset sampling with the reference's COST profile, for the CPU baseline only: like
    sample_initial_state (graph/util.py:119-143) it re-enumerates every free cell, recomputes the
    rotation steps and the weights on EVERY reset, then draws with numpy's ``choice(p=weights)`` from a
    seeded RandomState (the reference uses the unseeded global stream)."""

    def __init__(self, scene, goals, optimal_distance_fn, seed):
        self.scene, self.goals, self.od = scene, goals, optimal_distance_fn
        self.rng = np.random.RandomState(seed)

    def __call__(self):
        t = int(self.rng.randint(len(self.goals)))
        goal = self.goals[t]
        pots, dists = gu.initial_state_candidates(self.scene.maze, self.scene.graph, self.scene.optimal_actions, goal)
        w = gu.initial_state_weights(dists, self.od(t))
        x = self.rng.choice(np.arange(len(pots)), p=w)
        return t, pots[x]


# --------------------------------------------------------------------------- the experiment's wrapper stack as classes
# experiments/thor_cached_auxiliary.py:58-64 builds every env as
#     UnrealEnvBaseWrapper(ScaledFloatFrame(TransposeImage(RewardCollector(env))))
# and hands the env functions to SubprocVecEnv.  The four wrappers and the VecEnv live in deep_rl (UN-VENDORED,
# PARITY UNPINNED): the classes below restate them [recalled] as gym-style wrappers - observation_space included,
# because Trainer.create_model reads it (thor_cached_auxiliary.py:55) - so that the reference's own create_envs /
# Trainer can be executed unmodified against them (oracle/ref_harness.ref_experiment).
class _Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class _Tuple:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)


def _map_space(space, fn):
    if hasattr(space, "spaces"):
        return _Tuple(tuple(_map_space(s, fn) for s in space.spaces))
    return fn(space)


def _map_obs(obs, fn):
    if isinstance(obs, tuple):
        return tuple(_map_obs(o, fn) for o in obs)
    return fn(obs)


class _Wrapper:
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    @property
    def unwrapped(self):
        return self.env.unwrapped if hasattr(self.env, "unwrapped") else self.env

    def reset(self):
        return self.observation(self.env.reset())

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    def observation(self, obs):
        return obs


class TimeLimitWrapper(_Wrapper):
    """gym 0.15.7 TimeLimit as gym.make applies it for ids registered with max_episode_steps
    (environments/gym_graph/__init__.py:24-28: 900 for AuxiliaryGraph-v0)."""

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max, self._elapsed = max_episode_steps, 0

    def reset(self):
        self._elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed += 1
        if self._elapsed >= self._max:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info


class RewardCollectorWrapper(_Wrapper):
    """deep_rl.common.env.RewardCollector [recalled]."""

    def __init__(self, env):
        super().__init__(env)
        self.ret, self.len = 0.0, 0

    def reset(self):
        self.ret, self.len = 0.0, 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.ret = float(np.float32(np.float32(self.ret) + np.float32(reward)))
        self.len += 1
        if done:
            info["episode"] = dict(r=self.ret, l=self.len)
        return obs, reward, done, info


class TransposeImageWrapper(_Wrapper):
    """deep_rl.common.env.TransposeImage [recalled]: every image leaf HWC -> CHW, spaces included."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = _map_space(env.observation_space,
                                            lambda b: _Box(b.low, b.high, (b.shape[2], b.shape[0], b.shape[1]), b.dtype))

    def observation(self, obs):
        return _map_obs(obs, lambda x: np.transpose(x, (2, 0, 1)))


class ScaledFloatFrameWrapper(_Wrapper):
    """deep_rl.common.env.ScaledFloatFrame [recalled]: float32(x) / 255.0, spaces become Box(0, 1, float32)."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = _map_space(env.observation_space, lambda b: _Box(0.0, 1.0, b.shape, np.float32))

    def observation(self, obs):
        return _map_obs(obs, lambda x: np.asarray(x).astype(np.float32) / 255.0)


class UnrealEnvBaseWrapper(_Wrapper):
    """deep_rl.a2c_unreal.util.UnrealEnvBaseWrapper [recalled]: obs -> (obs, last_action_reward)."""

    def __init__(self, env):
        super().__init__(env)
        self.n_actions = env.action_space.n
        self.observation_space = _Tuple((env.observation_space, _Box(0.0, 1.0, (self.n_actions + 1,), np.float32)))

    def reset(self):
        return self.env.reset(), np.zeros(self.n_actions + 1, np.float32)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return (obs, last_action_reward(int(action), reward, self.n_actions)), reward, done, info


class InProcessVecEnv:
    """deep_rl.common.vec_env.SubprocVecEnv / DummyVecEnv [recalled], all workers in this process: the worker loop
    ``ob, r, done, info = env.step(a); if done: ob = env.reset()``, leaves stacked on axis 0, ``call_unwrapped``."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.observation_space = self.envs[0].observation_space
        self.action_space = self.envs[0].action_space
        self.unwrapped_calls = []

    @staticmethod
    def _stack(obs_list):
        first = obs_list[0]
        if isinstance(first, tuple):
            return tuple(InProcessVecEnv._stack([o[i] for o in obs_list]) for i in range(len(first)))
        return np.stack(obs_list)

    def reset(self):
        return self._stack([e.reset() for e in self.envs])

    def step(self, actions):
        obs, rews, dones, infos = [], [], [], []
        for e, a in zip(self.envs, actions):
            ob, r, d, info = e.step(int(a))
            if d:
                ob = e.reset()
            obs.append(ob)
            rews.append(r)
            dones.append(d)
            infos.append(info)
        return self._stack(obs), np.array(rews, np.float32), np.array(dones, bool), infos

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        return self.step(self._pending)

    def call_unwrapped(self, name, *args, **kwargs):
        self.unwrapped_calls.append((name, args))
        return [getattr(e.unwrapped, name)(*args, **kwargs) for e in self.envs]

    def close(self):
        pass
