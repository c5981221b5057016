"""Shared test plumbing: load a golden fixture, rebuild its synthetic scene, and replay it through
the ORACLE (oracle/envs.py + oracle/vec.py) with the recorded action / reset streams."""
import importlib
import os
import zlib

import numpy as np

from oracle import envs as oenvs
from oracle import vec as ovec

vn = importlib.import_module("a2cat-vn-pytorch_b200")
scenes = vn.scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


class StreamSource:
    """reset_source that replays recorded (choice, start) pairs of one env."""

    def __init__(self, choices, starts, count):
        self.c, self.s, self.n, self.k = choices, starts, int(count), 0

    def __call__(self):
        assert self.k < self.n, "reset stream exhausted"
        c, s = int(self.c[self.k]), self.s[self.k]
        self.k += 1
        return c, (tuple(int(v) for v in np.atleast_1d(s)) if np.ndim(s) else int(s))


def schedule_of(g):
    return {int(t): (None if c < 0 else float(c)) for t, c in zip(g["sched_t"], g["sched_c"])}


def scene_from_golden(g, oriented, planes, suffix=""):
    if ("goals" + suffix) in g:
        goals = [tuple(int(v) for v in x) for x in np.atleast_2d(g["goals" + suffix])]
    elif ("goal" + suffix) in g:
        goals = [tuple(int(v) for v in g["goal" + suffix])]
    else:
        goals = []
    return scenes.GridScene(g["maze" + suffix], goals, oriented, (84, 84), planes,
                            frame_seed=int(g["frame_seed" + suffix]),
                            scene_id=int(g["scene_id"]) if "scene_id" in g else int(suffix or 0))


def replay_oracle(g, make_env, n_leaves, state_of=lambda e: e.state, width=3):
    """Runs the golden protocol (TimeLimit + auto-reset) on oracle envs.  Returns the same record
    layout as tests/golden/make_golden.py: drive()."""
    actions = g["actions"]
    T, N = actions.shape
    sched = schedule_of(g) if "sched_t" in g else {}
    key_c = "reset_choice" if "reset_choice" in g else "reset_goal"
    envs = []
    for i in range(N):
        e = make_env(i)
        e.reset_source = StreamSource(g[key_c][i], g["reset_start"][i], g["reset_count"][i])
        envs.append(ovec.TimeLimit(e, int(g["max_episode_steps"])))
    if 0 in sched:
        for e in envs:
            e.set_complexity(sched[0])

    def leafs(ob):
        if isinstance(ob, tuple):
            return [crc(x) for x in ob]
        return [crc(ob)]

    rec = dict(rewards=np.zeros((T, N)), dones=np.zeros((T, N), bool), env_dones=np.zeros((T, N), bool),
               truncated=np.zeros((T, N), bool), wins=np.zeros((T, N), bool),
               obs_crc=np.zeros((T, N, n_leaves), np.uint32),
               states=np.zeros((T, N, width), np.int32), post_states=np.zeros((T, N, width), np.int32))
    rec["reset_obs_crc"] = np.array([leafs(e.reset()) for e in envs], np.uint32)
    rec["reset_states"] = np.array([np.atleast_1d(state_of(e)) for e in envs], np.int32)
    for t in range(T):
        if t in sched and t != 0:
            for e in envs:
                e.set_complexity(sched[t])
        for i, e in enumerate(envs):
            a = int(actions[t, i])
            ob, r, d, info = e.step(None if a < 0 else a)
            rec["states"][t, i] = np.atleast_1d(state_of(e))
            rec["wins"][t, i] = bool(info.get("win", False))
            rec["truncated"][t, i] = bool(info.get("TimeLimit.truncated", False))
            rec["env_dones"][t, i] = d and not rec["truncated"][t, i]
            if d:
                ob = e.reset()
            rec["post_states"][t, i] = np.atleast_1d(state_of(e))
            rec["rewards"][t, i] = r
            rec["dones"][t, i] = d
            rec["obs_crc"][t, i] = leafs(ob)
    return rec


def assert_record_equal(rec, g, obs=True):
    for k in ("states", "post_states", "dones", "env_dones", "truncated", "wins", "reset_states"):
        a, b = np.asarray(rec[k]), np.asarray(g[k])
        assert np.array_equal(a.reshape(b.shape), b), k
    # rewards: bit-exact as float64 including the sign of zero (cached.py returns -0.0)
    assert np.array_equal(np.asarray(rec["rewards"], np.float64).view(np.uint64),
                          np.asarray(g["rewards"], np.float64).view(np.uint64)), "rewards"
    if obs:
        assert np.array_equal(rec["obs_crc"], g["obs_crc"]), "obs_crc"
        assert np.array_equal(rec["reset_obs_crc"], g["reset_obs_crc"]), "reset_obs_crc"


def write_reference_style_pickle(scene, path, frame=12):
    """Pickles ``scene`` the way the reference stores its scenes: an object of class
    graph.multi_graph_no_tp.ThorGridWorld whose state is its __dict__ (dense [X,Y,4,H,W,C] arrays).  The class is
    faked in a throw-away module, so this works on boxes without the reference tree."""
    import pickle
    import sys
    import types
    X, Y = scene.maze.shape
    h, w = scene.frame_hw
    pkg, mod = types.ModuleType("graph"), types.ModuleType("graph.multi_graph_no_tp")

    class ThorGridWorld:
        pass

    ThorGridWorld.__module__, ThorGridWorld.__qualname__ = "graph.multi_graph_no_tp", "ThorGridWorld"
    mod.ThorGridWorld = ThorGridWorld
    saved = {k: sys.modules.get(k) for k in ("graph", "graph.multi_graph_no_tp")}
    sys.modules["graph"], sys.modules["graph.multi_graph_no_tp"] = pkg, mod
    try:
        wld = ThorGridWorld()
        wld._maze = scene.maze.copy()
        for attr, plane, c in (("_observations", "rgb", 3), ("_depths", "depth", 1), ("_segmentations", "segmentation", 3)):
            a = np.zeros((X, Y, 4, h, w, c), np.uint8)
            a[scene.cells[:, 0], scene.cells[:, 1]] = scene.plane_frames(plane).reshape(scene.n_cells, 4, h, w, c)
            setattr(wld, attr, a)
        wld.goals = list(scene.goals)
        wld.graph = "stale tables are ignored"
        with open(path, "wb") as f:
            pickle.dump(wld, f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def trainer_contract_scene(g):
    """The synthetic 174 x 174 scene of tests/golden/trainer_contract.npz (make_golden.trainer_contract_scene)."""
    return scenes.GridScene(g["maze"], [tuple(int(v) for v in g["goal"])], True, (174, 174),
                            ("rgb", "depth", "segmentation"), frame_seed=int(g["frame_seed"]), scene_id=int(g["scene_id"]))


def check_trainer_contract_run(g, env, crc_of_leaf, complexity_setter):
    """Replays the golden of the reference's own Trainer / create_envs (experiments/thor_cached_auxiliary.py) on `env`
    - anything with the VecEnv surface - and compares everything that crosses the seam: leaf order / shape / dtype,
    every float32 CHW leaf (CRC), last_action_reward, rewards, dones, episode infos."""
    T, N = g["actions"].shape
    obs = env.reset()
    leaves, lar = obs
    assert len(leaves) == 5 and [tuple(x.shape) for x in leaves] == [tuple(s) for s in g["obs_shapes"]]
    assert [str(x.dtype).replace("torch.", "") for x in leaves] == list(g["obs_dtypes"][:5])
    assert np.array_equal(np.array([[crc_of_leaf(x, i) for x in leaves] for i in range(N)], np.uint32), g["reset_obs_crc"])
    assert np.array_equal(np.asarray(lar.cpu() if hasattr(lar, "cpu") else lar), g["reset_lar"])
    sched = dict(zip(g["hardness_t"].tolist(), g["hardness_c"].tolist()))
    for t in range(T):
        if t in sched:
            complexity_setter(sched[t])
        (leaves, lar), r, d, infos = env.step(g["actions"][t])
        assert np.array_equal(np.asarray(d), g["dones"][t]), t
        assert np.array_equal(np.asarray(r, np.float32).view(np.uint32), g["rewards"][t].view(np.uint32)), t
        assert np.array_equal(np.asarray(lar.cpu() if hasattr(lar, "cpu") else lar, np.float32).view(np.uint32),
                              g["lar"][t].view(np.uint32)), t
        got = np.array([[crc_of_leaf(x, i) for x in leaves] for i in range(N)], np.uint32)
        assert np.array_equal(got, g["obs_crc"][t]), t
        for i in range(N):
            if g["dones"][t, i]:
                ep = infos[i]["episode"]
                assert np.float32(ep["r"]) == g["ep_r"][t, i] and ep["l"] == g["ep_l"][t, i], (t, i)
            else:
                assert "episode" not in infos[i]
    assert g["dones"].sum() > 100


def thor_cached_task_scenes(g):
    """The two synthetic scenes of tests/golden/thor_cached_tasks.npz (make_golden.thor_cached_task_scenes)."""
    return [scenes.GridScene(g["maze%d" % k], [], True, (84, 84), ("rgb",), frame_seed=int(g["frame_seed%d" % k]),
                             scene_id=k) for k in range(2)]
