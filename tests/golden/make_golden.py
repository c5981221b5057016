"""Generates tests/golden/*.npz by running the UNMODIFIED reference classes from /root/reference
under injected action / reset streams (oracle/ref_harness.py).  Run in the build container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; the committed .npz files are what travels.  Every fixture
stores the inputs (maze, goals, frame seed, actions, reset draws) and what the reference returned
(states, rewards, dones, CRC32 of every observation leaf), so both the oracle (CPU tests) and the
CUDA path (-m gpu tests) can be replayed against it.

Protocol used to drive the single-env reference classes as a vectorised env (the reference gets
this from un-vendored gym TimeLimit + baselines SubprocVecEnv):
    ob, r, done, info = env.step(a); elapsed += 1
    if elapsed >= max_episode_steps: truncated = not done; done = True
    if done: ob = env.reset(); elapsed = 0
"""
import io
import os
import pickle
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import importlib  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

vn = importlib.import_module("a2cat-vn-pytorch_b200")
scenes = vn.scenes


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def dense_world(ref, scene):
    """GridScene -> reference ThorGridWorld with dense [X,Y,4,H,W,C] arrays
    (graph/multi_graph_no_tp.py:141-152)."""
    X, Y = scene.maze.shape
    h, w = scene.frame_hw
    arrs = {}
    for plane, c in (("rgb", 3), ("depth", 1), ("segmentation", 3)):
        a = np.zeros((X, Y, 4, h, w, c), np.uint8)
        fr = scene.plane_frames(plane)
        a[scene.cells[:, 0], scene.cells[:, 1]] = fr.reshape(scene.n_cells, 4, h, w, c)
        arrs[plane] = a
    return ref.thor_world.ThorGridWorld(scene.maze.copy(), arrs["rgb"], arrs["depth"], arrs["segmentation"])


def graph_file(world):
    return io.BytesIO(pickle.dumps(world))


def drive(envs, actions, max_episode_steps, leaves, complexity_schedule=None, get_state=None, on_reset=None):
    """Runs the protocol in the module docstring.  Returns dict of recorded arrays."""
    T, N = actions.shape
    get_state = get_state or (lambda e: e.state)
    rec = dict(rewards=np.zeros((T, N), np.float64), dones=np.zeros((T, N), bool),
               env_dones=np.zeros((T, N), bool), truncated=np.zeros((T, N), bool),
               wins=np.zeros((T, N), bool), obs_crc=np.zeros((T, N, leaves), np.uint32),
               states=[[None] * N for _ in range(T)], post_states=[[None] * N for _ in range(T)])

    def leaf_crcs(ob):
        if isinstance(ob, tuple):
            return [crc(x) for x in ob]
        return [crc(ob)]

    obs0 = []
    for i, e in enumerate(envs):
        ob = e.reset()
        if on_reset:
            on_reset(i, e)
        obs0.append(leaf_crcs(ob))
    rec["reset_obs_crc"] = np.array(obs0, np.uint32)
    rec["reset_states"] = [get_state(e) for e in envs]
    elapsed = [0] * N
    for t in range(T):
        if complexity_schedule and t in complexity_schedule:
            for e in envs:
                e.set_complexity(complexity_schedule[t])
        for i, e in enumerate(envs):
            a = int(actions[t, i])
            ob, r, d, info = e.step(None if a < 0 else a)
            rec["states"][t][i] = get_state(e)
            rec["env_dones"][t, i] = d
            rec["wins"][t, i] = bool(info.get("win", False))
            elapsed[i] += 1
            if elapsed[i] >= max_episode_steps:
                rec["truncated"][t, i] = not d
                d = True
            if d:
                ob = e.reset()
                if on_reset:
                    on_reset(i, e)
                elapsed[i] = 0
            rec["post_states"][t][i] = get_state(e)
            rec["rewards"][t, i] = r
            rec["dones"][t, i] = d
            rec["obs_crc"][t, i] = leaf_crcs(ob)
    rec["states"] = np.array(rec["states"], np.int32)
    rec["post_states"] = np.array(rec["post_states"], np.int32)
    rec["reset_states"] = np.array(rec["reset_states"], np.int32)
    return rec


class ResetLog:
    """Collects per-env (goal_choice, start_state) in the order resets happened."""

    def __init__(self, n):
        self.choice = [[] for _ in range(n)]
        self.start = [[] for _ in range(n)]

    def pack(self, width):
        R = max(len(x) for x in self.start)
        n = len(self.start)
        c = np.zeros((n, R), np.int32)
        s = np.zeros((n, R, width), np.int32)
        cnt = np.zeros(n, np.int32)
        for i in range(n):
            cnt[i] = len(self.start[i])
            c[i, :cnt[i]] = self.choice[i]
            s[i, :cnt[i]] = np.array(self.start[i], np.int32).reshape(-1, width)
        return c, s, cnt


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


# --------------------------------------------------------------------------------------------
def gen_graph_util(ref):
    """graph/util.py: compute_shortest_path_data (recursive DFS, unmodified), candidate lists and
    weights of sample_initial_state / sample_initial_position (captured at np.random.choice)."""
    out = {}
    specs = [((10, 10), 0.25, 0), ((9, 7), 0.3, 3), ((6, 12), 0.15, 5)]
    out["n_mazes"] = len(specs)
    for k, (shape, wp, seed) in enumerate(specs):
        maze = scenes.random_maze(shape, wp, seed)
        dist, act = ref.util.compute_shortest_path_data(maze)
        out["maze%d" % k] = maze
        out["dist%d" % k] = dist
        out["act%d" % k] = act
        g = type("G", (), {})()
        g.maze, g.graph, g.optimal_actions = maze, dist, act
        cells = np.argwhere(maze)
        rng = np.random.RandomState(100 + k)
        goals = [tuple(int(v) for v in cells[i]) + (int(rng.randint(4)),) for i in rng.choice(len(cells), 3, False)]
        out["goals%d" % k] = np.array(goals, np.int32)
        maxd = int(dist.max())
        for gi, goal in enumerate(goals):
            for ci, c in enumerate([None, 0.01, 0.3, 0.7, 1.0]):
                # oriented (gym_graph/graph.py:49-51 offset +3) and un-oriented (graph/env.py:105 offset -1)
                for kind in ("state", "position"):
                    inj = rh.InjectedChoice(np.arange(4, dtype=np.uint64) * 977 + 5)
                    orig = np.random.choice
                    np.random.choice = inj
                    try:
                        if kind == "state":
                            od = None if c is None else c * (maxd + 4 - 1) + 1
                            ref.util.sample_initial_state(g, goal, optimal_distance=od)
                        else:
                            od = None if c is None else c * (maxd - 1) + 1
                            ref.util.sample_initial_position(g, goal[:2], optimal_distance=od)
                    finally:
                        np.random.choice = orig
                    n, p, _ = inj.calls[0]
                    w = np.full(n, 1.0 / n) if p is None else p
                    out["w_%s_%d_%d_%d" % (kind, k, gi, ci)] = w
    out["complexities"] = np.array([-1, 0.01, 0.3, 0.7, 1.0])   # -1 stands for None
    save("graph_util", **out)


def _gym_graph_run(ref, name, aux, goals, rewards, seed, n_envs, T, max_steps, schedule, shape=(10, 10)):
    scene = scenes.make_maze_scene(shape, 0.25, seed, n_goals=1)
    world = dense_world(ref, scene)
    rng = np.random.RandomState(seed + 1)
    actions = rng.randint(0, 4, size=(T, n_envs)).astype(np.int32)
    # bias towards forward moves so that goals are actually reached
    actions[rng.rand(T, n_envs) < 0.35] = 0
    cls = ref.gym_graph.GoalGymGraphAuxiliaryEnv if aux else ref.gym_graph.OrientedGraphEnv
    envs = []
    for i in range(n_envs):
        e = cls(graph_file=graph_file(world), goals=goals, screen_size=(84, 84), rewards=list(rewards))
        if aux:   # the aux env does not forward screen_size to GraphResize
            e.graph = ref.core.GraphResize(e.graph._graph, (84, 84))
        envs.append(e)
    log = ResetLog(n_envs)
    goal_rng = np.random.RandomState(seed + 2)
    inj_random = rh.InjectedRandom(goal_rng.randint(0, 1 << 30, size=100000))
    inj_choice = rh.InjectedChoice(np.random.RandomState(seed + 3).randint(0, 1 << 32, size=100000, dtype=np.uint64))
    ref.gym_graph.random = inj_random
    orig = np.random.choice
    np.random.choice = inj_choice
    last_choice = {}

    class _R:  # record which goal index random.choice returned
        @staticmethod
        def choice(seq):
            v = inj_random._next()
            last_choice["v"] = v % len(seq)
            return seq[v % len(seq)]

    ref.gym_graph.random = _R

    def on_reset(i, e):
        log.choice[i].append(last_choice.get("v", 0) if isinstance(goals, list) else 0)
        log.start[i].append(e.state)

    try:
        if schedule and 0 in schedule:
            for e in envs:
                e.set_complexity(schedule[0])
        rec = drive(envs, actions, max_steps, 5 if aux else 1, schedule, on_reset=on_reset)
    finally:
        np.random.choice = orig
        import random as _random
        ref.gym_graph.random = _random
    c, s, cnt = log.pack(3)
    sched_t = np.array(sorted(schedule.keys()), np.int32) if schedule else np.zeros(0, np.int32)
    sched_c = np.array([-1.0 if schedule[t] is None else schedule[t] for t in sorted(schedule.keys())]) if schedule \
        else np.zeros(0)
    save(name, maze=scene.maze, frame_seed=scene.frame_seed, scene_id=scene.scene_id,
         goals=np.array(goals if isinstance(goals, list) else [goals], np.int32), goals_is_list=isinstance(goals, list),
         rewards_cfg=np.array(rewards, np.float64), actions=actions, max_episode_steps=max_steps,
         reset_choice=c, reset_start=s, reset_count=cnt, sched_t=sched_t, sched_c=sched_c,
         largest_distance=envs[0].largest_distance,
         **{k: v for k, v in rec.items()})


def gen_gym_graph(ref):
    scene = scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=1)
    cells = scene.cells
    g0 = (int(cells[5][0]), int(cells[5][1]), 1)
    g1 = (int(cells[40][0]), int(cells[40][1]), 3)
    g2 = (int(cells[-1][0]), int(cells[-1][1]), 0)
    # C1: GoalGymGraphAuxiliaryEnv, 16 envs, goal list, curriculum on then off, TimeLimit truncations
    _gym_graph_run(ref, "gym_graph_aux", True, [g0, g1, g2], (1.0, 0.0, 0.0), 0, 16, 400, 60,
                   {0: 0.3, 150: 0.05, 250: None, 330: 1.0})
    # OrientedGraphEnv, single tuple goal, non-default rewards, no curriculum
    _gym_graph_run(ref, "gym_graph_oriented", False, g1, (1.0, -0.01, -0.1), 0, 8, 500, 900, None)


def gen_graph_env(ref):
    """graph/env.py SimpleGraphEnv / MultipleGraphEnv / OrientedGraphEnv over cached-frame scenes."""

    class CachedCellScene(ref.core.GridWorldScene):
        """An un-oriented scene whose render(state) returns the cached uint8 frame of the cell
        (dtype uint8 -> the envs apply /255, graph/env.py:110-115)."""

        def __init__(self, gs):
            super().__init__()
            self._gs = gs
            self._maze = gs.maze
            self.graph, self.optimal_actions = ref.util.compute_shortest_path_data(gs.maze)
            self.goal = tuple(gs.goals[0])
            self._frames = gs.plane_frames("rgb")

        @property
        def maze(self):
            return self._maze

        @property
        def observation_shape(self):
            return self._gs.frame_hw + (3,)

        def render(self, state):
            return self._frames[self._gs.cell_rank[state[0], state[1]]]

    def run(name, env_factory, n_envs, T, max_steps, schedule, seed, width, get_choice=None, noop=True):
        rng = np.random.RandomState(seed + 1)
        actions = rng.randint(0, 4, size=(T, n_envs)).astype(np.int32)
        if noop:
            actions[rng.rand(T, n_envs) < 0.03] = -1
        envs = [env_factory() for _ in range(n_envs)]
        log = ResetLog(n_envs)
        inj_random = rh.InjectedRandom(np.random.RandomState(seed + 2).randint(0, 1 << 30, size=100000))
        inj_choice = rh.InjectedChoice(
            np.random.RandomState(seed + 3).randint(0, 1 << 32, size=100000, dtype=np.uint64))
        ref.graph_env.random = inj_random
        orig = np.random.choice
        np.random.choice = inj_choice

        def on_reset(i, e):
            log.choice[i].append(get_choice(e) if get_choice else 0)
            log.start[i].append(e.state)

        try:
            if schedule and 0 in schedule:
                for e in envs:
                    e.set_complexity(schedule[0])
            rec = drive(envs, actions, max_steps, 1, schedule, on_reset=on_reset)
        finally:
            np.random.choice = orig
            import random as _random
            ref.graph_env.random = _random
        c, s, cnt = log.pack(width)
        sched_t = np.array(sorted(schedule.keys()), np.int32) if schedule else np.zeros(0, np.int32)
        sched_c = np.array([-1.0 if schedule[t] is None else schedule[t] for t in sorted(schedule.keys())]) \
            if schedule else np.zeros(0)
        return dict(actions=actions, max_episode_steps=max_steps, reset_choice=c, reset_start=s, reset_count=cnt,
                    sched_t=sched_t, sched_c=sched_c, **rec)

    # SimpleGraphEnv on one un-oriented maze
    gs = scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=1, oriented=False, planes=("rgb",))
    sc = CachedCellScene(gs)
    rec = run("simple", lambda: ref.graph_env.SimpleGraphEnv(sc, rewards=[1.0, -0.02, -0.2]), 8, 400, 80,
              {0: 0.2, 200: None, 300: 0.9}, 11, 2)
    save("graph_env_simple", maze=gs.maze, frame_seed=gs.frame_seed, scene_id=gs.scene_id,
         goals=np.array(gs.goals, np.int32), rewards_cfg=np.array([1.0, -0.02, -0.2]), **rec)

    # MultipleGraphEnv on three mazes
    gss = [scenes.make_maze_scene((8, 9), 0.2, 20 + k, n_goals=1, oriented=False, planes=("rgb",), scene_id=k)
           for k in range(3)]
    scs = [CachedCellScene(g) for g in gss]
    rec = run("multi", lambda: ref.graph_env.MultipleGraphEnv(scs), 8, 400, 50, {0: 0.5, 250: None}, 13, 2,
              get_choice=lambda e: e.graph_number)
    extra = {}
    for k, g in enumerate(gss):
        extra["maze%d" % k] = g.maze
        extra["goal%d" % k] = np.array(g.goals[0], np.int32)
        extra["frame_seed%d" % k] = g.frame_seed
    save("graph_env_multiple", n_graphs=3, rewards_cfg=np.array([1.0, 0.0, 0.0]), **extra, **rec)

    # graph/env.py OrientedGraphEnv: never reaches its goal -> only time-limit dones
    so = scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=1)
    world = dense_world(ref, so)
    world.graph, world.optimal_actions = ref.util.compute_shortest_path_data(world.maze)
    goal = (int(so.cells[7][0]), int(so.cells[7][1]), 2)
    rec = run("oriented", lambda: ref.graph_env.OrientedGraphEnv(world, goal), 4, 300, 40, None, 17, 3, noop=False)
    save("graph_env_oriented", maze=so.maze, frame_seed=so.frame_seed, scene_id=so.scene_id,
         goals=np.array([goal], np.int32), rewards_cfg=np.array([1.0, 0.0, 0.0]), **rec)


def gen_maze_render(ref):
    """MazeGraph.render (graph/maze_graph.py:20-24) through GraphResize((84,84)) (graph/core.py:43-57)."""
    maze = scenes.random_maze((10, 10), 0.25, 0)
    cells = np.argwhere(maze)
    goal = tuple(int(v) for v in cells[0])
    # maze as float array like DungeonGraph produces (tiles 0/1)
    mg = ref.maze_graph.MazeGraph.__new__(ref.maze_graph.MazeGraph)
    mg._maze, mg.goal = maze.astype(np.float32), goal
    rz = ref.core.GraphResize(mg, (84, 84))
    frames = np.stack([rz.render(tuple(int(v) for v in c)) for c in cells])        # float32 [n,84,84,3]
    raw = np.stack([mg.render(tuple(int(v) for v in c)) for c in cells])           # float32 [n,10,10,3]
    save("maze_render", maze=maze, goal=np.array(goal, np.int32), raw_crc=np.array([crc(x) for x in raw], np.uint32),
         resized_u8_crc=np.array([crc(np.clip(np.rint(x * 255.0), 0, 255).astype(np.uint8)) for x in frames],
                                 np.uint32), sample_resized=frames[3].astype(np.float32))


def gen_thor_cached(ref):
    """environments/gym_ai2thor/envs/cached.py THORDiscreteCachedEnv on the flat schema."""
    from oracle import graph_util as gu
    gs = scenes.make_maze_scene((9, 8), 0.2, 31, n_goals=1, planes=("rgb",))
    dist, _ = ref.util.compute_shortest_path_data(gs.maze)
    locs, graph, spd = gu.h5_tables(gs.maze, dist)
    obs = gs.plane_frames("rgb")
    rh.FakeH5File.registry["mem://scene31"] = dict(observation=obs, location=np.zeros((len(locs) * 4, 2)),
                                                   graph=graph, shortest_path_distance=spd)
    n_envs, T, max_steps = 8, 500, 70
    rng = np.random.RandomState(41)
    actions = rng.randint(0, 4, size=(T, n_envs)).astype(np.int32)
    actions[rng.rand(T, n_envs) < 0.3] = 0
    inj = rh.InjectedRandom(np.random.RandomState(42).randint(0, 1 << 30, size=200000))
    ref.cached.random = inj
    log = ResetLog(n_envs)
    try:
        envs = []
        for i in range(n_envs):
            e = ref.cached.THORDiscreteCachedEnv(h5_file_path="mem://scene31", image_size=(84, 84))
            envs.append(e)   # __init__ already called reset() once (cached.py:36); drive() resets again

        def on_reset(i, e):
            log.choice[i].append(int(e._current_goal_idx))
            log.start[i].append(int(e._current_state_idx))

        def leafs(e):
            return e._current_state_idx

        # drive() expects tuple observations -> crc of (obs, goal) float64 leaves
        rec = drive(envs, actions, max_steps, 2, None, get_state=leafs, on_reset=on_reset)
    finally:
        import random as _random
        ref.cached.random = _random
    c, s, cnt = log.pack(1)
    # the reference returns float64 frames / 255 (skimage resize); store the CRC of the equivalent
    # uint8 frames as well so byte-parity of the gather can be asserted
    save("thor_cached", maze=gs.maze, frame_seed=gs.frame_seed, scene_id=gs.scene_id, actions=actions,
         max_episode_steps=max_steps, reset_goal=c, reset_start=s[:, :, 0], reset_count=cnt, graph=graph,
         spd_crc=crc(spd), **rec)


def gen_aux_target(ref_trainer):
    import torch
    rng = np.random.RandomState(7)
    x = (rng.randint(0, 256, size=(2, 3, 3, 84, 84)).astype(np.float32) / np.float32(255.0))
    y20 = ref_trainer.compute_auxiliary_target(torch.from_numpy(x), 4, (20, 20)).numpy()
    y21 = ref_trainer.compute_auxiliary_target(torch.from_numpy(x), 4, None).numpy()
    d = (rng.randint(0, 256, size=(2, 3, 1, 84, 84)).astype(np.float32) / np.float32(255.0))
    tup = ref_trainer.compute_auxiliary_targets(((x, x, d, x, x), None), 4, (20, 20)) \
        if False else None
    yd = ref_trainer.compute_auxiliary_target(torch.from_numpy(d), 4, (20, 20)).numpy()
    save("aux_target", x_seed=7, y20=y20, y21=y21, yd=yd)


def thor_cached_task_scenes():
    return [scenes.make_maze_scene((7, 8), 0.2, 51, n_goals=1, planes=("rgb",), scene_id=0),
            scenes.make_maze_scene((6, 9), 0.25, 52, n_goals=1, planes=("rgb",), scene_id=1)]


def gen_thor_cached_tasks(ref):
    """environments/gym_thor_cached.py THORCachedEnv (the unfinished multi-scene class) run AS WRITTEN on
    two synthetic scenes and a task list, through oracle/ref_harness.A8Driver (which only supplies the names the class
    never defines).  Pins: task choice from the (scene, goal) list (:47), start sampling by rejection on
    shortest_path_distance > 0 (:37-43), the raw uint8 (obs, goal) pair of observe() (:52-53), and process() (:72-95):
    reward / terminal rules and the {'image', 'goal'} dict of float32 / 255 frames, previous dict on a terminal step."""
    from oracle import graph_util as gu
    mod = rh.ref_thor_cached_tasks()
    scs = thor_cached_task_scenes()
    names = ["sceneA", "sceneB"]
    for name, gs in zip(names, scs):
        dist, _ = ref.util.compute_shortest_path_data(gs.maze)
        locs, graph, spd = gu.h5_tables(gs.maze, dist)
        rh.FakeH5File.registry["mem:/%s.h5" % name] = dict(observation=gs.plane_frames("rgb"),
                                                           location=np.zeros((len(locs) * 4, 2)), graph=graph,
                                                           shortest_path_distance=spd)
    tasks = [("sceneA", 7), ("sceneB", 30), ("sceneA", 61), ("sceneB", 2)]
    n_envs, T, max_steps = 6, 400, 25
    rng = np.random.RandomState(61)
    actions = rng.randint(0, 4, size=(T, n_envs)).astype(np.int32)
    actions[rng.rand(T, n_envs) < 0.3] = 0
    log = ResetLog(n_envs)
    drivers = []
    for i in range(n_envs):
        e = mod.THORCachedEnv(tasks, image_size=(84, 84))       # __init__ resets once with its own unseeded Random
        e._random = rh.InjectedRandom(np.random.RandomState(70 + i).randint(0, 1 << 30, size=100000))
        drivers.append(rh.A8Driver(e))

    def on_reset(i, d):
        e = d.e
        scene_name = [n for n in names if e.scenes.get(n) is e.current_scene][0]
        log.choice[i].append(tasks.index((scene_name, e.goal)))
        log.start[i].append(int(e.state))

    class _Leaves:                     # drive() records tuple observations leaf by leaf
        def __init__(self, d):
            self.d, self.e = d, d.e

        def reset(self):
            return self.d.reset()

        def step(self, a):
            st, r, term, info = self.d.step(a)
            return (st["image"], st["goal"]), r, term, info

    rec = drive([_Leaves(d) for d in drivers], actions, max_steps, 2, None, get_state=lambda w: int(w.e.state),
                on_reset=lambda i, w: on_reset(i, w.d))
    c, s, cnt = log.pack(1)
    save("thor_cached_tasks", actions=actions, max_episode_steps=max_steps,
         task_scene=np.array([names.index(n) for n, _ in tasks], np.int32), task_goal=np.array([g for _, g in tasks], np.int32),
         reset_choice=c, reset_start=s[:, :, 0], reset_count=cnt,
         **{"maze%d" % k: gs.maze for k, gs in enumerate(scs)},
         **{"frame_seed%d" % k: gs.frame_seed for k, gs in enumerate(scs)},
         **rec)


def trainer_contract_scene():
    """The synthetic stand-in for 'thor-cached-212-174' (download.py:21-29): native 174 x 174 frames."""
    return scenes.make_maze_scene((6, 6), 0.2, 11, n_goals=1, frame_hw=(174, 174))


def gen_trainer_contract(ref):
    """experiments/thor_cached_auxiliary.py run UNMODIFIED (rh.ref_experiment): ``Trainer()`` is constructed,
    ``Trainer.create_env`` (:50-52) calls the reference's own ``create_envs`` (:58-71) with ``default_args()`` - only
    the scene name / goal are ours - and ``Trainer.create_model`` (:54-56) reads the observation space.  The VecEnv is
    then driven with an injected action stream; what crosses the seam is recorded: spaces, Model arguments, the
    set_hardness calls, and per step every float32 CHW leaf (CRC), last_action_reward, rewards, dones, episode infos."""
    scene = trainer_contract_scene()
    world = dense_world(ref, scene)
    world.graph, world.optimal_actions = ref.util.compute_shortest_path_data(scene.maze)
    calls = []
    exp = rh.ref_experiment(lambda name: world, calls)
    args = exp.default_args()
    goal = scene.goals[0]
    env_kwargs = dict(args["env_kwargs"], tasks=[("synthetic-174", [goal])] * 4)
    N, T = 4, 240
    inj_choice = rh.InjectedChoice(np.random.RandomState(5).randint(0, 1 << 32, size=100000, dtype=np.uint64))
    orig = np.random.choice
    np.random.choice = inj_choice
    log = ResetLog(N)
    try:
        trainer = exp.Trainer()
        assert (trainer.num_processes, trainer.num_steps, trainer.gamma) == (4, 20, .99)
        env = trainer.env = trainer.create_env(env_kwargs)
        trainer.create_model()
        for i, e in enumerate(env.envs):             # record every start state the reference samples from here on
            base = e.unwrapped

            def rec_reset(orig_reset=base.reset, i=i, base=base):
                ob = orig_reset()
                log.choice[i].append(0)
                log.start[i].append(base.state)
                return ob
            base.reset = rec_reset
        rng = np.random.RandomState(6)
        actions = rng.randint(0, 4, size=(T, N)).astype(np.int32)
        greedy = rng.rand(T, N) < 0.6

        def toward_goal(state):
            """A policy that actually finishes episodes: the action whose successor is closest to the goal (grid
            distance, then rotation mismatch) - only a way to pick the recorded action stream."""
            best, best_a = None, 0
            for a in range(4):
                n = ref.util.step(state, a)
                if not ref.util.is_valid_state(world.maze, n):
                    continue
                key = (int(world.graph[n[0], n[1], goal[0], goal[1]]), min((n[2] - goal[2]) % 4, (goal[2] - n[2]) % 4))
                if best is None or key < best:
                    best, best_a = key, a
            return best_a

        def leaf_crcs(obs):
            leaves, lar = obs
            return np.array([[crc(x[i]) for x in leaves] for i in range(N)], np.uint32), np.array(lar, np.float32)

        obs = env.reset()
        reset_crc, reset_lar = leaf_crcs(obs)
        shapes = np.array([x.shape for x in obs[0]], np.int32)
        dtypes = np.array([str(x.dtype) for x in obs[0]] + [str(obs[1].dtype)])
        rec = dict(rewards=np.zeros((T, N), np.float32), dones=np.zeros((T, N), bool), obs_crc=np.zeros((T, N, 5), np.uint32),
                   lar=np.zeros((T, N, 5), np.float32), ep_r=np.zeros((T, N), np.float32), ep_l=np.zeros((T, N), np.int32),
                   states=np.zeros((T, N, 3), np.int32), hardness_t=np.array([120], np.int32), hardness_c=np.array([0.3]))
        for t in range(T):
            if t == 120:
                env.set_hardness(0.3)                                   # the schedule the trainer applies, :45 / :68
            for i, e in enumerate(env.envs):
                if greedy[t, i]:
                    actions[t, i] = toward_goal(e.unwrapped.state)
            obs, r, d, infos = env.step(actions[t])
            rec["rewards"][t], rec["dones"][t] = r, d
            rec["obs_crc"][t], rec["lar"][t] = leaf_crcs(obs)
            rec["states"][t] = [e.unwrapped.state for e in env.envs]
            for i, info in enumerate(infos):
                if d[i]:
                    rec["ep_r"][t, i], rec["ep_l"][t, i] = info["episode"]["r"], info["episode"]["l"]
    finally:
        np.random.choice = orig
    sp = env.observation_space
    c, s, cnt = log.pack(3)
    save("trainer_contract", maze=scene.maze, frame_seed=scene.frame_seed, scene_id=scene.scene_id,
         goal=np.array(goal, np.int32), actions=actions, max_episode_steps=900,
         reset_choice=c, reset_start=s, reset_count=cnt,
         # what Trainer.create_model passed to Model (:55) and what it read it from
         model_args=np.array(calls[0][0], np.int32),
         space_leaf_shapes=np.array([b.shape for b in sp.spaces[0].spaces], np.int32),
         space_lar_shape=np.array(sp.spaces[1].shape, np.int32), action_n=env.action_space.n,
         unwrapped_calls=np.array([[a[0] for _, a in env.unwrapped_calls]]),       # set_complexity(0.01), set_complexity(0.3)
         unwrapped_names=np.array([n for n, _ in env.unwrapped_calls]),
         obs_shapes=shapes, obs_dtypes=dtypes, reset_obs_crc=reset_crc, reset_lar=reset_lar,
         reset_states=np.array([log.start[i][0] for i in range(N)], np.int32), **rec)


if __name__ == "__main__":
    ref = rh.ref_modules()
    gen_graph_util(ref)
    gen_gym_graph(ref)
    gen_graph_env(ref)
    gen_maze_render(ref)
    gen_thor_cached(ref)
    gen_aux_target(rh.ref_aux_trainer())
    gen_trainer_contract(ref)
    gen_thor_cached_tasks(ref)
