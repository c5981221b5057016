"""Parity of the CUDA path (through the C ABI of libvn_b200.so) with
  (a) the golden trajectories recorded from the UNMODIFIED reference classes, and
  (b) the oracle restatement driven with the same streams (incl. the Philox reset path).
Bit-exact for states / rewards / dones / observation bytes."""
import importlib

import numpy as np
import pytest

import helpers as H
from oracle import envs as oenvs
from oracle import graph_util as gu
from oracle import vec as ovec

pytestmark = pytest.mark.gpu

vn = importlib.import_module("a2cat-vn-pytorch_b200")
T = vn.tables


def f32bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def replay_device(g, world, env_kwargs, inject_task, inject_start, n_leaves, tuple_width, noop_ok=False):
    """Drives GraphVecEnv with the golden's actions + injected resets; returns a record like drive()."""
    import torch
    actions = g["actions"]
    Tn, N = actions.shape
    sched = H.schedule_of(g) if "sched_t" in g else {}
    env = vn.GraphVecEnv(world, N, max_episode_steps=int(g["max_episode_steps"]), unreal_wrapper=False,
                         inject=(inject_task, inject_start), **env_kwargs)
    if 0 in sched:
        env.set_complexity(sched[0])

    def leaf_crcs(obs):
        leaves = obs if isinstance(obs, tuple) else (obs,)
        host = [x.cpu().numpy() for x in leaves]
        return [[H.crc(x[i]) for x in host] for i in range(N)]

    rec = dict(rewards=np.zeros((Tn, N), np.float32), dones=np.zeros((Tn, N), bool), truncated=np.zeros((Tn, N), bool),
               wins=np.zeros((Tn, N), bool), env_dones=np.zeros((Tn, N), bool),
               obs_crc=np.zeros((Tn, N, n_leaves), np.uint32), states=np.zeros((Tn, N, tuple_width), np.int32),
               post_states=np.zeros((Tn, N, tuple_width), np.int32))
    obs = env.reset()
    rec["reset_obs_crc"] = np.array(leaf_crcs(obs), np.uint32)
    tup = (lambda s: np.atleast_1d(world.state_tuple(int(s)))) if tuple_width > 1 else (lambda s: np.array([int(s)]))
    rec["reset_states"] = np.array([tup(s) for s in env.state.cpu().numpy()], np.int32)
    for t in range(Tn):
        if t in sched and t != 0:
            env.set_complexity(sched[t])
        obs, rew, done, infos = env.step(actions[t])
        h = infos._host()
        rec["rewards"][t] = rew
        rec["dones"][t] = done
        rec["truncated"][t] = h["truncated"] == 1
        rec["wins"][t] = h["win"].astype(bool)
        rec["env_dones"][t] = done & ~(h["truncated"] == 1)
        rec["states"][t] = np.array([tup(s) for s in h["info_state"]], np.int32)
        rec["post_states"][t] = np.array([tup(s) for s in env.state.cpu().numpy()], np.int32)
        rec["obs_crc"][t] = leaf_crcs(obs)
        # infos dictionaries carry the reference keys
        if t % 97 == 0:
            for i in range(N):
                info = infos[i]
                if done[i]:
                    assert "episode" in info
    return rec, env


def assert_equal(rec, g, obs=True, wins=True):
    for k in ("states", "post_states", "dones", "env_dones", "truncated", "reset_states") + (("wins",) if wins else ()):
        a, b = np.asarray(rec[k]), np.asarray(g[k])
        assert np.array_equal(a.reshape(b.shape), b), k
    assert np.array_equal(f32bits(rec["rewards"]), f32bits(g["rewards"])), "rewards (bitwise, incl. -0.0)"
    if obs:
        assert np.array_equal(rec["obs_crc"], g["obs_crc"]), "observation bytes"
        assert np.array_equal(rec["reset_obs_crc"], g["reset_obs_crc"]), "reset observation bytes"


def starts_to_index(scene, starts, counts):
    out = np.zeros(starts.shape[:2], np.int32)
    for i in range(starts.shape[0]):
        for k in range(int(counts[i])):
            out[i, k] = scene.state_index(tuple(int(v) for v in starts[i, k]))
    return out


@pytest.mark.parametrize("gather,skip", [("ldg", True), ("bulk", True), ("bulk", False), ("fused", True),
                                         ("fused", False), ("auto", True), ("persistent", True), ("persistent", False)])
def test_golden_gym_graph_auxiliary(gather, skip):
    g = H.load("gym_graph_aux")
    scene = H.scene_from_golden(g, True, ("rgb", "depth", "segmentation"))
    goals = [tuple(int(v) for v in x) for x in g["goals"]]
    world = T.compile_world([scene], T.GYM_GRAPH, tasks=[(0, gl) for gl in goals])
    N = g["actions"].shape[1]
    env_tasks = np.tile(np.array([[0, len(goals)]], np.int32), (N, 1))
    rec, env = replay_device(g, world, dict(obs_layout="aux5", rewards=tuple(g["rewards_cfg"]), env_tasks=env_tasks,
                                            gather=gather, skip_unchanged=skip),
                             g["reset_choice"], starts_to_index(scene, g["reset_start"], g["reset_count"]), 5, 3)
    assert_equal(rec, g)
    st = env.episode_stats()
    assert st["episodes"] == g["dones"].sum() and st["successes"] == g["wins"].sum()
    assert st["steps"] == g["actions"].size and st["truncations"] == g["truncated"].sum()
    # rows are skipped exactly when the env's record did not change: a collision, or a reset onto the same state
    same = (g["post_states"] == np.concatenate([g["reset_states"][None], g["post_states"][:-1]])).all(-1)
    assert st["rows_skipped"] == (same.sum() if skip else 0)
    # one launch per step when fused, two otherwise (+ the reset)
    per_step = 1 if gather in ("fused", "auto", "persistent") else 2
    assert env.kernel_launches == per_step * (g["actions"].shape[0] + 1)


def test_golden_gym_graph_oriented():
    g = H.load("gym_graph_oriented")
    scene = H.scene_from_golden(g, True, ("rgb", "depth", "segmentation"))
    goal = tuple(int(v) for v in g["goals"][0])
    world = T.compile_world([scene], T.GYM_GRAPH, tasks=[(0, goal)])
    rec, _ = replay_device(g, world, dict(obs_layout="frame", rewards=tuple(g["rewards_cfg"])),
                           g["reset_choice"], starts_to_index(scene, g["reset_start"], g["reset_count"]), 1, 3)
    assert_equal(rec, g)


def test_golden_graph_env_simple():
    g = H.load("graph_env_simple")
    scene = H.scene_from_golden(g, False, ("rgb",))
    world = T.compile_world([scene], T.SIMPLE_GRAPH)
    rec, _ = replay_device(g, world, dict(obs_layout="frame", rewards=tuple(g["rewards_cfg"])),
                           g["reset_choice"], starts_to_index(scene, g["reset_start"], g["reset_count"]), 1, 2)
    # the reference env returns float32 frame / 255 (graph/env.py:110-115): compare those bytes
    assert_equal(rec, g, obs=False)
    import torch
    env = vn.GraphVecEnv(world, 1, unreal_wrapper=False, obs_layout="frame",
                         inject=(g["reset_choice"][:1], starts_to_index(scene, g["reset_start"][:1], g["reset_count"][:1])))
    ob = env.reset()
    f = (ob[0].cpu().numpy().astype(np.float32) / 255.0)
    assert H.crc(f) == int(g["reset_obs_crc"][0, 0])
    f2 = vn.rollout.policy_input(env.dw, env.state)[0].cpu().numpy()       # fused gather + /255 + CHW
    assert np.array_equal(np.transpose(f2, (1, 2, 0)), f)


def test_golden_graph_env_multiple():
    g = H.load("graph_env_multiple")
    K = int(g["n_graphs"])
    scs = [H.scenes.GridScene(g["maze%d" % k], [tuple(int(v) for v in g["goal%d" % k])], False, (84, 84), ("rgb",),
                              frame_seed=int(g["frame_seed%d" % k]), scene_id=k) for k in range(K)]
    world = T.compile_world(scs, T.SIMPLE_GRAPH)
    N = g["actions"].shape[1]
    starts = np.zeros(g["reset_start"].shape[:2], np.int32)
    for i in range(N):
        for k in range(int(g["reset_count"][i])):
            gi = int(g["reset_choice"][i, k])
            starts[i, k] = world.scene_base[gi] + scs[gi].state_index(tuple(int(v) for v in g["reset_start"][i, k]))
    env_tasks = np.tile(np.array([[0, K]], np.int32), (N, 1))
    rec, _ = replay_device(g, world, dict(obs_layout="frame", rewards=tuple(g["rewards_cfg"]), env_tasks=env_tasks),
                           g["reset_choice"], starts, 1, 2)
    assert_equal(rec, g, obs=False)


def test_golden_graph_env_oriented_never_terminates():
    g = H.load("graph_env_oriented")
    scene = H.scene_from_golden(g, True, ("rgb", "depth", "segmentation"))
    goal = tuple(int(v) for v in g["goals"][0])
    world = T.compile_world([scene], T.GRAPH_ENV_ORIENTED, tasks=[(0, goal)])
    rec, _ = replay_device(g, world, dict(obs_layout="frame"), g["reset_choice"],
                           starts_to_index(scene, g["reset_start"], g["reset_count"]), 1, 3)
    assert_equal(rec, g, obs=False)


def test_golden_thor_cached():
    g = H.load("thor_cached")
    scene = H.scene_from_golden(g, True, ("rgb",))
    # the reference draws the goal from all states (cached.py:39): one task per distinct recorded goal
    goals = sorted(set(int(v) for i in range(g["reset_goal"].shape[0]) for v in g["reset_goal"][i, :g["reset_count"][i]]))
    world = T.compile_world([scene], T.THOR_CACHED, tasks=[(0, gl) for gl in goals])
    assert np.array_equal(world.adj, g["graph"])          # the h5 'graph' dataset, action order fwd/back/rot+/rot-
    lut = {gl: k for k, gl in enumerate(goals)}
    N = g["actions"].shape[1]
    task = np.zeros_like(g["reset_goal"])
    for i in range(N):
        for k in range(int(g["reset_count"][i])):
            task[i, k] = lut[int(g["reset_goal"][i, k])]
    env_tasks = np.tile(np.array([[0, len(goals)]], np.int32), (N, 1))
    rec, _ = replay_device(g, world, dict(obs_layout="pair", env_tasks=env_tasks), task, g["reset_start"], 2, 1)
    assert_equal(rec, g, obs=False, wins=False)
    assert np.signbit(rec["rewards"][rec["rewards"] == 0]).any()        # -0.0 of cached.py:84 reproduced


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("family,oriented", [("gym_graph", True), ("simple_graph", False)])
def test_philox_reset_path_matches_oracle(family, oriented):
    """No injection: the device's Philox sampling + curriculum prefix vs the oracle restatement that
    is built from the oracle's own candidate lists."""
    import torch
    fam = T.FAMILIES[family]
    planes = ("rgb", "depth", "segmentation") if oriented else ("rgb",)
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 3, n_goals=3, oriented=oriented, planes=planes)
    osc = oenvs.OracleScene(scene)
    world = T.compile_world([scene], fam)
    N, Tn, seed, max_steps = 12, 300, 1234, 25
    env_tasks = np.tile(np.array([[0, 3]], np.int32), (N, 1)) if oriented else None
    env = vn.GraphVecEnv(world, N, seed=seed, max_episode_steps=max_steps, unreal_wrapper=True,
                         obs_layout="aux5" if oriented else "frame", env_tasks=env_tasks)
    oes = []
    for i in range(N):
        if oriented:
            e = oenvs.GymGraphAuxiliaryEnv(osc, goals=list(scene.goals))
            cands = [gu.initial_state_candidates(scene.maze, osc.graph, osc.optimal_actions, gl) for gl in scene.goals]
            e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t, e=e: e.optimal_distance(), False)
        else:
            # default env_tasks deal tasks round-robin: env i owns task i % 3
            gl = scene.goals[i % 3]
            e = oenvs.SimpleGraphEnv(osc, goal=gl)
            cands = [gu.initial_position_candidates(scene.maze, osc.graph, gl)]
            e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t, e=e: e.optimal_distance(), True)
        oes.append(ovec.RewardCollector(ovec.TimeLimit(e, max_steps)))
    ov = ovec.VecEnv(oes)
    rng = np.random.RandomState(5)
    for c in (0.25, None):
        env.set_complexity(c)
        for e in oes:
            e.set_complexity(c)
        (obs, lar) = env.reset()
        oobs, olar = ov.reset()
        for t in range(Tn):
            a = rng.randint(-1 if not oriented else 0, 4, size=N)
            (obs, lar), rew, done, infos = env.step(a)
            (oobs, olar), orew, odone, oinfos = ov.step(a)
            assert np.array_equal(done, odone), t
            assert np.array_equal(f32bits(rew), f32bits(orew)), t
            assert np.array_equal(f32bits(lar.cpu().numpy()), f32bits(olar)), t
            assert [oe.state for oe in oes] == env.states(), t
            leaves = obs if isinstance(obs, tuple) else (obs,)
            oleaves = oobs if isinstance(oobs, tuple) else (oobs,)
            for x, y in zip(leaves, oleaves):
                y = y if y.dtype == np.uint8 else np.rint(y * 255).astype(np.uint8)
                assert np.array_equal(x.cpu().numpy(), y), t
            for i in range(N):
                info, oinfo = infos[i], oinfos[i]
                assert info.get("episode") == oinfo.get("episode"), (t, i)
                assert info.get("TimeLimit.truncated") == oinfo.get("TimeLimit.truncated"), (t, i)
                assert info.get("state") == oinfo.get("state") and info.get("win") == oinfo.get("win"), (t, i)
        assert done.sum() >= 0


def test_store_fill_matches_host_hash():
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 9, n_goals=1, planes=("rgb", "depth", "segmentation"), scene_id=4)
    world = T.compile_world([scene], T.GYM_GRAPH)
    dw = vn.DeviceWorld(world)
    for p in ("rgb", "depth", "segmentation"):
        dev = dw.plane_view(p).cpu().numpy()
        assert np.array_equal(dev, scene.plane_frames(p)), p
    lay = world.layout
    pad = np.ones(lay.state_pitch, bool)
    for o, b in zip(lay.plane_off, lay.plane_bytes):
        pad[o:o + b] = False
    assert not dw.frames.cpu().numpy()[:, pad].any()


_ORACLE_SCENES = {}


def _oracle_scene(key, scene):
    """All-pairs tables of a full-size scene take ~15 s in the scalar oracle: built once per test session."""
    from oracle import envs as oenvs
    if key not in _ORACLE_SCENES:
        _ORACLE_SCENES[key] = oenvs.OracleScene(scene)
    return _ORACLE_SCENES[key]


def _oracle_subsample(world, scene, osc, env_ids, seed, limit, complexity, make_env):
    """Scalar oracle envs for the GLOBAL env ids `env_ids` of a default-task-dealt batch (env i owns task i % n_tasks),
    driven by the device path's own Philox reset stream."""
    from oracle import graph_util as gu, vec as ovec
    cands = {}
    oes = []
    for i in env_ids:
        task = world.tasks[int(i) % len(world.tasks)]
        goal = tuple(task.goal) if not isinstance(task.goal, (int, np.integer)) else world.state_tuple(int(task.goal))
        if goal not in cands:
            cands[goal] = gu.initial_state_candidates(scene.maze, osc.graph, osc.optimal_actions, goal)
        oe = make_env(osc, goal)
        oe.set_complexity(complexity)
        oe.reset_source = ovec.PhiloxResetSource(seed, int(i), [cands[goal]], lambda t, oe=oe: oe.optimal_distance(), False)
        oes.append(ovec.RewardCollector(ovec.TimeLimit(oe, limit)))
    return ovec.VecEnv(oes), oes


def test_full_size_properties_c2():
    """BASELINE.json configs[1] at full size (1,500 cells x 4 rotations, 4,096 envs): every launch mode is bit-identical
    to the others, size-independent properties hold for ALL envs, and a sub-sample of 64 envs (every 64th) is driven
    through the scalar oracle with the same actions and the same Philox reset stream: states, rewards, dones and
    observation bytes of those envs equal the oracle's bit for bit at every step."""
    import torch
    from oracle import envs as oenvs, graph_util as gu, vec as ovec
    scene = H.scenes.make_thor_scene(1500, (50, 60), seed=0, n_goals=4, planes=("rgb", "depth"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    assert world.n_states == 6000
    N, seed, limit = 4096, 7, 50
    envs = {v: vn.GraphVecEnv(world, N, seed=seed, max_episode_steps=limit, obs_layout="rgbd_goal", gather=v,
                              host_outputs=False) for v in ("ldg", "bulk", "persistent")}
    dw = envs["ldg"].dw
    adj = dw.adj.view(-1, 4)
    rgb, depth = dw.plane_view("rgb"), dw.plane_view("depth")
    gen = torch.Generator(device="cuda").manual_seed(0)
    for e in envs.values():
        e.set_complexity(0.1)
        e.reset()
    e0, e1, e2 = envs["ldg"], envs["bulk"], envs["persistent"]
    # the oracle sub-sample: env i owns task i % 4 (default env_tasks: tasks dealt round-robin)
    sub = np.arange(0, N, 64)
    osc = _oracle_scene("thor1500-s0", scene)
    ov, oes = _oracle_subsample(world, scene, osc, sub, seed, limit, 0.1,
                                lambda o, goal: oenvs.GymGraphRgbdGoalEnv(o, goals=goal))
    (oobs, _) = ov.reset()
    assert [oe.state for oe in oes] == [world.state_tuple(int(v)) for v in e0.state[sub].cpu().numpy()]
    for t in range(120):
        a = torch.randint(0, 4, (N,), device="cuda", generator=gen, dtype=torch.int32)
        prev = e0.state.clone()
        ((o_rgb, o_goal, o_depth), lar), rew, done, _ = e0.step(a)
        for other in (e1, e2):
            ((b_rgb, b_goal, b_depth), blar), brew, bdone, _ = other.step(a)
            # the launch modes are bit-identical
            assert torch.equal(o_rgb, b_rgb) and torch.equal(o_depth, b_depth) and torch.equal(o_goal, b_goal)
            assert torch.equal(rew, brew) and torch.equal(done, bdone) and torch.equal(e0.state, other.state)
            assert torch.equal(lar, blar)
        # the oracle sub-sample
        (oobs, olar), orew, odone, _ = ov.step(a[sub].cpu().numpy())
        assert np.array_equal(done[sub].cpu().numpy(), odone) and np.array_equal(f32bits(rew[sub].cpu().numpy()), f32bits(orew))
        assert [oe.state for oe in oes] == [world.state_tuple(int(v)) for v in e0.state[sub].cpu().numpy()], t
        for leaf, oleaf in zip((o_rgb, o_goal, o_depth), oobs):
            assert np.array_equal(leaf[sub].cpu().numpy(), oleaf), t
        assert np.array_equal(f32bits(lar[sub].cpu().numpy()), f32bits(olar)), t
        # transition table property: next = adj[prev, a], unchanged on collision
        nxt = adj[prev.long(), a.long()]
        moved = torch.where(nxt >= 0, nxt, prev)
        assert torch.equal(e0.info_state, moved)
        # gather property: every observation row equals the store frame of the env's current state
        s = e0.state.long()
        assert torch.equal(o_rgb, rgb[s]) and torch.equal(o_depth, depth[s])
        assert torch.equal(o_goal, rgb[e0.goal.long()])
        # resets land on curriculum-eligible candidates and zero the wrapper vector
        r = e0.did_reset.bool()
        assert torch.equal(r, done)
        assert (lar[r] == 0).all()
    st = e0.episode_stats()
    assert st["steps"] == 120 * N and st["episodes"] == st["resets"] - N > 0


def test_checkpoint_roundtrip():
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    a = vn.GraphVecEnv(world, 32, seed=3, max_episode_steps=20)
    a.reset()
    rng = np.random.RandomState(0)
    for _ in range(30):
        a.step(rng.randint(0, 4, 32))
    sd = a.state_dict()
    b = vn.GraphVecEnv(world, 32, seed=3, max_episode_steps=20, device_world=a.dw)
    b.load_state_dict(sd)
    for _ in range(40):
        act = rng.randint(0, 4, 32)
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert np.array_equal(ra, rb) and np.array_equal(da, db) and torch.equal(a.state, b.state)
        assert all(torch.equal(x, y) for x, y in zip(oa[0], ob[0]))


def test_sharding_is_invisible():
    """RNG keyed by GLOBAL env id: 2 shards of 24 envs == one process with 48 envs."""
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    whole = vn.GraphVecEnv(world, 48, seed=11, max_episode_steps=15)
    parts = [vn.GraphVecEnv(world, 48, seed=11, max_episode_steps=15, rank=r, world_size=2, device_world=whole.dw)
             for r in range(2)]
    whole.reset()
    [p.reset() for p in parts]
    rng = np.random.RandomState(1)
    for _ in range(60):
        a = rng.randint(0, 4, 48)
        _, r, d, _ = whole.step(a)
        outs = [p.step(a[p.env_lo:p.env_lo + p.num_envs]) for p in parts]
        assert np.array_equal(r, np.concatenate([o[1] for o in outs]))
        assert np.array_equal(d, np.concatenate([o[2] for o in outs]))
        assert torch.equal(whole.state, torch.cat([p.state for p in parts]))


def test_c_abi_rejects_bad_arguments():
    import ctypes as C
    L = vn.lib
    lib = L.load()
    st = L.Store()
    rc = lib.vn_gather_plane(C.byref(st), 0, None, 0, None, 0, None)
    assert rc == -1 and b"store" in lib.vn_last_error()
    with pytest.raises(ValueError):
        scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=1, planes=("rgb",))
        vn.GraphVecEnv(T.compile_world([scene], T.GYM_GRAPH), 4, obs_layout="aux5")


def _property_run(env, world, steps, obs_names, oracle=None):
    """Size-independent properties of a vectorised run (used for the full-size C3 / C4 configurations).
    oracle = (local env indices, oracle VecEnv, oracle envs, local scene base): those envs are also checked, bit for
    bit, against the scalar oracle driven with the same actions and the same Philox reset stream."""
    import torch
    dw = env.dw
    adj = dw.adj.view(-1, 4)
    planes = {p: dw.plane_view(p) for p in obs_names}
    base = torch.from_numpy(world.scene_base).to(dw.device)
    task_scene = torch.tensor([t.scene for t in world.tasks], device=dw.device)
    gen = torch.Generator(device="cuda").manual_seed(0)
    env.reset()
    N = env.num_envs
    if oracle is not None:
        sub, ov, oes = oracle
        ov.reset()
        assert [oe.state for oe in oes] == [world.state_tuple(int(v)) for v in env.state[sub].cpu().numpy()]
    for t in range(steps):
        a = torch.randint(0, 4, (N,), device="cuda", generator=gen, dtype=torch.int32)
        prev = env.state.clone()
        obs, rew, done, _ = env.step(a)
        if oracle is not None:
            (oobs, olar), orew, odone, _ = ov.step(a[sub].cpu().numpy())
            assert np.array_equal(done[sub].cpu().numpy(), odone), t
            assert np.array_equal(f32bits(rew[sub].cpu().numpy()), f32bits(orew)), t
            assert [oe.state for oe in oes] == [world.state_tuple(int(v)) for v in env.state[sub].cpu().numpy()], t
            oleaves = oobs if isinstance(oobs, tuple) else (oobs,)
            leaves = obs[0] if env.unreal_wrapper else obs
            leaves = leaves if isinstance(leaves, tuple) else (leaves,)
            for leaf, oleaf in zip(leaves, oleaves):          # the oracle class may emit fewer leaves (frame only)
                assert np.array_equal(leaf[sub].cpu().numpy(), oleaf), t
        nxt = adj[prev.long(), a.long()]
        assert torch.equal(env.info_state, torch.where(nxt >= 0, nxt, prev))
        s = env.state.long()
        for name, buf in env.obs_buf.items():
            assert torch.equal(buf, planes[name][s]), name
        for name, buf in env.goal_buf.items():
            assert torch.equal(buf, planes[name][env.goal.long()]), name
        # an env never leaves the scene of its current task, and its goal is that task's goal
        sc = torch.bucketize(s, base, right=True) - 1
        assert torch.equal(sc, task_scene[env.task.long()])
        assert torch.equal(env.goal, dw.task_goal[env.task.long()])
        assert torch.equal(env.did_reset.bool(), done)
        assert torch.equal(rew == 1.0, env.win.bool())
    st = env.episode_stats()
    assert st["steps"] == steps * N and st["episodes"] == st["resets"] - N


def test_full_size_properties_c3_dungeon():
    """BASELINE.json configs[2]: dungeon 64x64 multi-room, 65,536 envs, goal-image conditioning, auto-reset."""
    sc = H.scenes.make_dungeon_scene((64, 64), 0, oriented=True, planes=("rgb",))
    world = T.compile_world([sc], T.GYM_GRAPH)
    from oracle import envs as oenvs
    for gather in ("auto", "persistent"):
        env = vn.GraphVecEnv(world, 65536, seed=3, max_episode_steps=40, obs_layout="pair", host_outputs=False,
                             gather=gather)
        env.set_complexity(0.1)
        # 64 of the 65,536 envs also run through the scalar oracle (OrientedGraphEnv: the frame leaf; the goal leaf is
        # checked against the store for every env by the property run)
        sub = np.arange(0, 65536, 1024)
        osc = _oracle_scene("dungeon64-s0", sc)
        ov, oes = _oracle_subsample(world, sc, osc, sub, 3, 40, 0.1, lambda o, goal: oenvs.GymGraphEnv(o, goals=goal))
        _property_run(env, world, 25, ("rgb",), oracle=(sub, ov, oes))
        del env


def test_full_size_properties_c4_multi_scene():
    """BASELINE.json configs[3] on one GPU's shard: 30 scenes resident (180,000 states, 5.1 GB store),
    32,768 envs (= 262,144 / 8), 120 (scene, goal) tasks dealt round-robin."""
    scs = [H.scenes.make_thor_scene(1500, (50, 60), seed=k, n_goals=4, planes=("rgb", "depth"), scene_id=k)
           for k in range(30)]
    world = T.compile_world(scs, T.GYM_GRAPH)
    assert world.n_states == 180000 and len(world.tasks) == 120
    env = vn.GraphVecEnv(world, 262144, seed=5, max_episode_steps=30, obs_layout="rgbd_goal", host_outputs=False,
                         rank=3, world_size=8)
    assert env.num_envs == 32768 and env.env_lo == 3 * 32768
    env.set_complexity(0.05)
    # oracle sub-sample: the first 64 envs of this shard whose task lives in scene 0 (tasks are dealt round-robin by
    # GLOBAL env id: task = id % 120, scene 0 owns tasks 0..3) - scene 0 is the C2 scene, its oracle tables are shared
    from oracle import envs as oenvs
    ids = np.arange(env.env_lo, env.env_lo + env.num_envs)
    sub_global = ids[(ids % 120) < 4][:64]
    osc = _oracle_scene("thor1500-s0", scs[0])
    ov, oes = _oracle_subsample(world, scs[0], osc, sub_global, 5, 30, 0.05,
                                lambda o, goal: oenvs.GymGraphRgbdGoalEnv(o, goals=goal))
    _property_run(env, world, 12, ("rgb", "depth"), oracle=(sub_global - env.env_lo, ov, oes))
    # the hash-filled store matches the host hash for a sample of (scene, state) pairs
    import torch
    rgb = env.dw.plane_view("rgb")
    for si, ls in ((0, 0), (7, 123), (29, 5999)):
        g = int(world.scene_base[si]) + ls
        assert np.array_equal(rgb[g].cpu().numpy(), scs[si].plane_frames("rgb", [ls])[0])


def test_single_env_surface_without_auto_reset():
    """GraphEnv = the bare reference class: no auto-reset, no TimeLimit; compared with the oracle env
    (pinned to the reference by the golden tests) step by step, including what happens AFTER done."""
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    osc = oenvs.OracleScene(scene)
    seed = 77
    env = vn.GraphEnv(world, seed=seed, obs_layout="aux5")
    o = oenvs.GymGraphAuxiliaryEnv(osc, goals=list(scene.goals))
    cands = [gu.initial_state_candidates(scene.maze, osc.graph, osc.optimal_actions, g) for g in scene.goals]
    o.reset_source = ovec.PhiloxResetSource(seed, 0, cands, lambda t: o.optimal_distance(), False)
    rng = np.random.RandomState(0)
    for c in (0.1, None):
        env.set_complexity(c)
        o.set_complexity(c)
        for ep in range(6):
            ob, oob = env.reset(), o.reset()
            assert env.state == o.state and env.goal == o.goal
            for t in range(60):
                a = int(rng.randint(0, 4))
                (ob, r, d, info), (oob, orr, od, oinfo) = env.step(a), o.step(a)
                assert (r, d) == (orr, od) and env.state == o.state
                assert info.get("state") == oinfo.get("state") and info.get("win") == oinfo.get("win")
                assert all(np.array_equal(x, y) for x, y in zip(ob, oob))
                if d:
                    break
    assert np.array_equal(env.render("rgbarray"), ob[0])
    with pytest.raises(Exception):
        env.render("human")


def test_thor_cached_single_env_returns_previous_obs_on_terminal():
    """cached.py:90-97: on the terminal step the env returns the PREVIOUS observation pair."""
    g = H.load("thor_cached")
    scene = H.scene_from_golden(g, True, ("rgb",))
    dist, _ = gu.compute_shortest_path_data(scene.maze)
    _, graph, spd = gu.h5_tables(scene.maze, dist)
    frames = scene.plane_frames("rgb")
    goal = 14
    world = T.compile_world([scene], T.THOR_CACHED, tasks=[(0, goal)])
    # start next to the goal: find (s, a) with graph[s][a] == goal
    s0, a0 = [(s, a) for s in range(graph.shape[0]) for a in range(4) if graph[s, a] == goal and s != goal][0]
    env = vn.GraphEnv(world, obs_layout="pair", inject=(np.zeros((1, 4), np.int32), np.full((1, 4), s0, np.int32)))
    o = oenvs.ThorCachedEnv(graph, frames, spd, tasks=[goal])
    o.reset_source = lambda: (0, s0)
    ob, oob = env.reset(), o.reset()
    assert all(np.array_equal(x, y) for x, y in zip(ob, oob))
    (ob, r, d, info), (oob, orr, od, _) = env.step(a0), o.step(a0)
    assert d and od and r == orr == 1.0 and info == {}
    assert all(np.array_equal(x, y) for x, y in zip(ob, oob))
    assert np.array_equal(ob[0], frames[s0])             # previous observation, not the goal frame


@pytest.mark.parametrize("N", [8, 300])
def test_scaled_float_observation_mode(N):
    """scaled_float=True: leaves are what TransposeImage + ScaledFloatFrame produce (float32 CHW / 255), kept in
    persistent batches that are filled straight from the store (rows that did not change are skipped, goal leaves
    are converted only for envs that reset; no uint8 batch exists in this mode)."""
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    a = vn.GraphVecEnv(world, N, seed=1, max_episode_steps=9)
    b = vn.GraphVecEnv(world, N, seed=1, max_episode_steps=9, scaled_float=True, device_world=a.dw)
    assert b.observation_space.spaces[0].spaces[0].shape == (3, 84, 84)       # thor_cached_auxiliary.py:55
    assert not b.obs_buf and not b.goal_buf and len(b.float_buf) == 5
    (oa, _), (ob, _) = a.reset(), b.reset()
    rng = np.random.RandomState(0)
    for _ in range(25):
        act = rng.randint(0, 4, N)
        (oa, _), _, _, _ = a.step(act)
        (ob, _), _, _, _ = b.step(act)
        for x, y in zip(oa, ob):
            # numpy true division like the reference's ScaledFloatFrame (torch's CUDA `x / 255.0` multiplies
            # by a rounded reciprocal and differs in the last bit for 126 of the 256 byte values)
            want = np.stack([ovec.transpose_scale(f) for f in x.cpu().numpy()])
            assert y.dtype == torch.float32 and np.array_equal(y.cpu().numpy(), want)
    assert b.episode_stats() == a.episode_stats() and b.episode_stats()["rows_skipped"] > 0
    # a checkpoint round trip refreshes the float batches
    sd = b.state_dict()
    for buf in b.float_buf.values():
        buf.zero_()
    b.load_state_dict(sd)
    for x, y in zip(oa, b._obs()[0]):
        assert np.array_equal(y.cpu().numpy(), np.stack([ovec.transpose_scale(f) for f in x.cpu().numpy()]))


def test_edge_cases_empty_ragged_and_errors():
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    # ragged sharding: 3 envs over 4 ranks -> one rank owns nothing; every call is a no-op there
    envs = [vn.GraphVecEnv(world, 3, seed=4, rank=r, world_size=4, max_episode_steps=5) for r in range(4)]
    assert [e.num_envs for e in envs] == [1, 1, 1, 0]
    empty = envs[3]
    obs = empty.reset()
    _, r, d, infos = empty.step(np.zeros(0, np.int32))
    assert r.shape == (0,) and d.shape == (0,) and len(infos) == 0 and obs[0][0].shape == (0, 84, 84, 3)
    # the three 1-env shards together equal one 3-env env
    whole = vn.GraphVecEnv(world, 3, seed=4, max_episode_steps=5, device_world=envs[0].dw)
    whole.reset()
    [e.reset() for e in envs[:3]]
    rng = np.random.RandomState(0)
    for _ in range(30):
        a = rng.randint(0, 4, 3)
        _, rw, dw_, _ = whole.step(a)
        parts = [e.step(a[i:i + 1]) for i, e in enumerate(envs[:3])]
        assert np.array_equal(rw, np.concatenate([p[1] for p in parts]))
        assert np.array_equal(dw_, np.concatenate([p[2] for p in parts]))
    # action containers: list, int64 CUDA tensor, int32 CPU tensor
    e = vn.GraphVecEnv(world, 4, seed=1, max_episode_steps=50, device_world=envs[0].dw)
    e.reset()
    e.step([0, 1, 2, 3])
    e.step(torch.tensor([0, 1, 2, 3], device="cuda"))
    e.step(torch.tensor([0, 1, 2, 3], dtype=torch.int32))
    with pytest.raises(ValueError):
        e.step([0, 1, 2])
    with pytest.raises(RuntimeError):
        e.step_wait()
    # an action outside [0, 4) has no transition in the reference (graph.util.step returns None): collision here
    before = e.state.clone()
    _, r, d, _ = e.step([7, -3, 4, 99])
    assert torch.equal(e.state, before) and (r == 0).all() and not d.any()
    e.close()
    with pytest.raises(RuntimeError):
        e.reset()
    # bad env_tasks
    with pytest.raises(ValueError):
        vn.GraphVecEnv(world, 2, env_tasks=[[0, 3], [0, 1]], device_world=envs[0].dw)


def test_rollout_edge_cases():
    import torch
    from oracle import rollout as orl
    R = vn.rollout
    # T = 1
    r = torch.tensor([[1.0], [0.5]], device="cuda")
    d = torch.tensor([[1], [0]], dtype=torch.uint8, device="cuda")
    v = torch.tensor([3.0, 4.0], device="cuda")
    got = R.nstep_returns(r, d, v, 0.9).cpu().numpy()
    assert np.array_equal(got, orl.nstep_returns(r.cpu().numpy(), d.cpu().numpy(), v.cpu().numpy(), 0.9))
    # no rewards at all / all non-zero
    for arr in (np.zeros(100, np.float32), np.ones(100, np.float32) * -1):
        lab, z, nz = R.reward_prediction_labels(torch.from_numpy(arr).cuda())
        assert len(z) + len(nz) == 100 and (len(nz) == 0 or len(z) == 0)
        assert np.array_equal(lab.cpu().numpy(), orl.rp_labels(arr))
    lab, z, nz = R.reward_prediction_labels(torch.zeros(0, device="cuda"))
    assert lab.numel() == 0 and z.numel() == 0 and nz.numel() == 0


def test_cuda_graph_capture_is_bit_identical():
    """The C ABI only enqueues work on the caller's stream, so whole multi-step rollouts capture into one
    CUDA graph; replays equal eager stepping bit for bit."""
    import torch
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    N, S = 16, 24
    a = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=7, host_outputs=False)
    b = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=7, host_outputs=False, device_world=a.dw)
    a.reset()
    b.reset()
    acts = torch.randint(0, 4, (S, N), device="cuda", dtype=torch.int32)
    rb = torch.zeros(S, N, device="cuda")
    db = torch.zeros(S, N, dtype=torch.uint8, device="cuda")

    def record(t):
        rb[t].copy_(b.reward)
        db[t].copy_(b.done)

    graph = b.capture_steps(acts, after_step=record)
    assert torch.equal(a.state, b.state)                         # capture left the env untouched
    for rep in range(3):
        ra, da = [], []
        for t in range(S):
            a.step_enqueue(acts[t])
            ra.append(a.reward.clone())
            da.append(a.done.clone())
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(torch.stack(ra), rb) and torch.equal(torch.stack(da), db)
        assert torch.equal(a.state, b.state) and torch.equal(a.epoch, b.epoch)
        assert all(torch.equal(x, y) for x, y in zip(a.obs_buf.values(), b.obs_buf.values()))
        assert all(torch.equal(x, y) for x, y in zip(a.goal_buf.values(), b.goal_buf.values()))
        acts.copy_(torch.randint(0, 4, (S, N), device="cuda", dtype=torch.int32))   # new content, same graph
    assert db.sum() > 0


def test_third_person_planes_six_tuple():
    """graph/thor_graph.py: scenes with third-person planes; render returns (rgb, depth, seg, tp_rgb, tp_depth,
    tp_seg).  Six planes in one store record, gathered by both kernel variants."""
    import torch
    planes = ("rgb", "depth", "segmentation", "tp_rgb", "tp_depth", "tp_segmentation")
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 2, n_goals=1, planes=planes)
    world = T.compile_world([scene], T.GYM_GRAPH)
    assert world.layout.state_pitch == 2 * (21248 + 7168 + 21248)
    for gather in ("bulk", "ldg", "fused"):
        env = vn.GraphVecEnv(world, 40, seed=2, max_episode_steps=11, obs_layout="thor6", unreal_wrapper=False,
                             gather=gather)
        obs = env.reset()
        rng = np.random.RandomState(1)
        for _ in range(20):
            obs, _, _, _ = env.step(rng.randint(0, 4, 40))
        s = env.state.cpu().numpy()
        for leaf, name in zip(obs, planes):
            assert np.array_equal(leaf.cpu().numpy(), scene.plane_frames(name, s)), (gather, name)


def test_optimal_policy_reaches_every_goal_in_the_shortest_number_of_actions():
    """End-to-end property of adjacency + goal test + reward + auto-reset: an agent that follows the
    state-graph shortest paths wins every episode in exactly dist(start) steps (SPL = 1)."""
    import torch
    scene = H.scenes.make_thor_scene(300, (30, 30), seed=4, n_goals=3, planes=("rgb",))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N = 512
    env = vn.GraphVecEnv(world, N, seed=8, max_episode_steps=900, obs_layout="frame", host_outputs=False)
    env.reset()
    a, d0 = env.optimal_actions()
    assert (d0 > 0).all()
    remaining = d0.clone()
    wins = torch.zeros(N, dtype=torch.int64, device="cuda")
    for t in range(200):
        a, d = env.optimal_actions()
        assert torch.equal(d, remaining)
        _, rew, done, _ = env.step(a)
        remaining = remaining - 1
        reached = remaining == 0
        assert torch.equal(done, reached) and torch.equal(rew == 1.0, reached)
        wins += reached
        # auto-reset: new episode, new distance
        _, dn = env.optimal_actions()
        remaining = torch.where(reached, dn, remaining)
    st = env.episode_stats()
    assert st["episodes"] == st["successes"] == int(wins.sum()) and st["collisions"] == 0 and st["truncations"] == 0
    assert st["length_sum"] > 0
    # oriented shortest paths are never shorter than the grid distance the curriculum uses minus nothing odd:
    dist, act = T.optimal_policy_table(world, 0)
    assert dist[world.tasks[0].goal_state] == 0 and (dist >= 0).all()


def test_c1_maze_graph_frames_a2c_n5_end_to_end():
    """BASELINE.json configs[0]: 10x10 grid maze, 16 envs, 84x84 RGB cached observation, A2C n = 5.
    The frames are MazeGraph.render (graph/maze_graph.py:20-24) hoisted through GraphResize((84, 84)) at store
    build time (scenes.render_maze_frames, pinned by tests/golden/maze_render.npz); the env is the
    SimpleGraphEnv family; 5-step rollouts feed the n-step return builder.  Everything vs the oracle."""
    import torch
    from oracle import rollout as orl
    maze = H.scenes.random_maze((10, 10), 0.25, 0)
    goal = tuple(int(v) for v in np.argwhere(maze)[0])               # DungeonGraph convention: first free cell
    plain = H.scenes.GridScene(maze, [goal], False, (84, 84), ("rgb",))
    frames = H.scenes.render_maze_frames(plain, goal, (84, 84))
    scene = H.scenes.GridScene(maze, [goal], False, (84, 84), ("rgb",), explicit={"rgb": frames})
    world = T.compile_world([scene], T.SIMPLE_GRAPH)
    N, n_step, seed, limit = 16, 5, 31, 100
    env = vn.GraphVecEnv(world, N, seed=seed, max_episode_steps=limit, obs_layout="frame", unreal_wrapper=False)
    osc = oenvs.OracleScene(scene)
    oes = []
    for i in range(N):
        e = oenvs.SimpleGraphEnv(osc, goal=goal)
        cands = [gu.initial_position_candidates(maze, osc.graph, goal)]
        e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t, e=e: e.optimal_distance(), True)
        oes.append(ovec.TimeLimit(e, limit))
    ov = ovec.VecEnv(oes)
    env.set_complexity(0.5)
    [e.set_complexity(0.5) for e in oes]
    obs, (oobs, _) = env.reset(), ov.reset()
    buf = vn.rollout.RolloutBuffer(env.dw, N, n_step)
    rng = np.random.RandomState(2)
    total_done = 0
    for it in range(40):                                            # 40 A2C iterations of n = 5 steps
        buf.start(env)
        R, D = np.zeros((N, n_step), np.float32), np.zeros((N, n_step), bool)
        for t in range(n_step):
            a = rng.randint(0, 4, N)
            obs, r, d, _ = env.step(a)
            (oobs, _), orr, od, _ = ov.step(a)
            buf.insert(env, torch.from_numpy(a))
            assert np.array_equal(d, od) and np.array_equal(f32bits(r), f32bits(orr))
            assert np.array_equal(obs.cpu().numpy(), np.rint(oobs * 255).astype(np.uint8))
            R[:, t], D[:, t] = orr, od
        v = rng.randn(N).astype(np.float32)
        got = buf.returns(torch.from_numpy(v).cuda(), 0.99).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), orl.nstep_returns(R, D, v, 0.99).view(np.uint32))
        total_done += int(D.sum())
    assert total_done > 5
    # the rendered frames themselves: free = white, walls = black, agent red, goal green (before the resize)
    assert frames.shape == (scene.n_cells, 84, 84, 3) and frames.max() == 255


def test_pipelined_mode_equals_serial_mode():
    """VN_STEP_ACTIONS_READY: the scalar kernel of step t+1 overlaps the gather of step t (double-buffered
    gather descriptors).  Results must not change - checked with frequent resets so that goal rows are rewritten."""
    import torch
    scene = H.scenes.make_thor_scene(200, (20, 25), seed=2, n_goals=4, planes=("rgb", "depth", "segmentation"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N, S = 2048, 150
    a = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=4, host_outputs=False)
    b = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=4, host_outputs=False, device_world=a.dw)
    a.set_complexity(0.3)
    a.reset()
    b.reset()
    acts = torch.randint(0, 4, (S, N), device="cuda", dtype=torch.int32)
    torch.cuda.synchronize()
    for t in range(S):
        a.step_enqueue(acts[t])
        b.step_enqueue(acts[t], actions_ready=True)
        if t % 7 == 0 or t == S - 1:
            torch.cuda.synchronize()
            assert torch.equal(a._pack, b._pack) and torch.equal(a.state, b.state) and torch.equal(a.lar, b.lar)
            for x, y in zip(list(a.obs_buf.values()) + list(a.goal_buf.values()),
                            list(b.obs_buf.values()) + list(b.goal_buf.values())):
                assert torch.equal(x, y)
            s = b.state.long()
            assert torch.equal(b.obs_buf["rgb"], b.dw.plane_view("rgb")[s])
            assert torch.equal(b.goal_buf["segmentation"], b.dw.plane_view("segmentation")[b.goal.long()])
    assert a.episode_stats() == b.episode_stats() and a.episode_stats()["episodes"] > N


@pytest.mark.parametrize("family", ["gym_graph", "thor_cached", "simple"])
def test_skipping_unchanged_rows_changes_nothing(family):
    """VN_STEP_SKIP_UNCHANGED: rows of envs whose record did not change (collisions, no-ops, the previous
    observation of a terminal THOR step) are not copied again.  The batch must stay byte-identical to the
    always-copy run, in every launch mode, and the skipped rows are counted."""
    import torch
    if family == "simple":
        scene = H.scenes.make_maze_scene((12, 12), 0.3, 3, n_goals=1, oriented=False, planes=("rgb",))
        world, layout = T.compile_world([scene], T.SIMPLE_GRAPH), "frame"
    else:
        scene = H.scenes.make_thor_scene(150, (16, 20), seed=4, n_goals=3, planes=("rgb", "depth"))
        if family == "gym_graph":
            world, layout = T.compile_world([scene], T.GYM_GRAPH), "rgbd_goal"
        else:       # flat-index goals, cached.py:39
            world, layout = T.compile_world([scene], T.THOR_CACHED, tasks=[(0, 5), (0, 77), (0, 301)]), "pair"
    for N, modes in ((1500, ("bulk", "ldg", "pipelined", "persistent")), (24, ("auto", "fused", "persistent"))):
        S = 60
        ref = vn.GraphVecEnv(world, N, seed=9, max_episode_steps=13, host_outputs=False, obs_layout=layout,
                             skip_unchanged=False, gather="bulk")
        envs = {m: vn.GraphVecEnv(world, N, seed=9, max_episode_steps=13, host_outputs=False, obs_layout=layout,
                                  device_world=ref.dw, gather="bulk" if m == "pipelined" else m) for m in modes}
        ref.reset()
        [e.reset() for e in envs.values()]
        lo = -1 if family == "simple" else 0          # SimpleGraphEnv: action -1 is a no-op (graph/env.py:118-120)
        acts = torch.randint(lo, 4, (S, N), device="cuda", dtype=torch.int32)
        torch.cuda.synchronize()
        for t in range(S):
            ref.step_enqueue(acts[t])
            for m, e in envs.items():
                e.step_enqueue(acts[t], actions_ready=(m == "pipelined"))
            if t % 5 == 0 or t == S - 1:
                torch.cuda.synchronize()
                for m, e in envs.items():
                    assert torch.equal(ref._pack, e._pack) and torch.equal(ref.state, e.state), (m, t)
                    for x, y in zip(list(ref.obs_buf.values()) + list(ref.goal_buf.values()),
                                    list(e.obs_buf.values()) + list(e.goal_buf.values())):
                        assert torch.equal(x, y), (m, t)
        base = ref.episode_stats()
        assert base["rows_skipped"] == 0 and base["collisions"] > 0
        for m, e in envs.items():
            st = e.episode_stats()
            assert st["rows_skipped"] > 0, m
            assert {k: v for k, v in st.items() if k != "rows_skipped"} == \
                   {k: v for k, v in base.items() if k != "rows_skipped"}, m
            assert st["rows_skipped"] == envs[modes[0]].episode_stats()["rows_skipped"]
        if family == "gym_graph":
            # gym_graph: the record is unchanged exactly on a collision that does not end at the time limit
            assert st["rows_skipped"] <= base["collisions"] and st["rows_skipped"] >= base["collisions"] - base["resets"]


def test_outputs_stay_inside_their_buffers():
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are looked for with canaries: every
    output and state array the kernels write is re-pointed into the middle of a guarded allocation, all launch
    modes run with frequent resets, and the guard bytes on both sides must be untouched."""
    import ctypes as C
    import torch
    G = 4096
    scene = H.scenes.make_thor_scene(120, (14, 18), seed=5, n_goals=3, planes=("rgb", "depth", "segmentation"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    lay = world.layout
    for N, gather in ((23, "fused"), (23, "bulk"), (777, "bulk"), (777, "ldg"), (777, "persistent"), (5, "persistent")):
        env = vn.GraphVecEnv(world, N, seed=3, max_episode_steps=5, host_outputs=False, obs_layout="aux5", gather=gather)
        guards = []

        def guarded(t):
            raw = torch.full((t.numel() * t.element_size() + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
            inner = raw[G:G + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            inner.copy_(t)
            guards.append(raw)
            return inner

        for i, p in enumerate(lay.planes):
            if p in env.obs_buf:
                env.obs_buf[p] = guarded(env.obs_buf[p])
                env._c_out.obs[i] = env.obs_buf[p].data_ptr()
            if p in env.goal_buf:
                env.goal_buf[p] = guarded(env.goal_buf[p])
                env._c_out.goal_obs[i] = env.goal_buf[p].data_ptr()
        env.lar = guarded(env.lar)
        env._c_out.last_action_reward = env.lar.data_ptr()
        env.obs_state = guarded(env.obs_state)
        env._c_out.obs_state = env.obs_state.data_ptr()
        env._gather_desc = guarded(env._gather_desc)
        env._c_out.gather_desc = env._gather_desc.data_ptr()
        for name in ("state", "goal", "task", "elapsed", "epoch", "ep_return", "ep_length"):
            setattr(env, name, guarded(getattr(env, name)))
            setattr(env._c_envs, name, getattr(env, name).data_ptr())
        env.reset()
        acts = torch.randint(0, 4, (40, N), device="cuda", dtype=torch.int32)
        for t in range(40):
            env.step_enqueue(acts[t], actions_ready=(t % 2 == 1 and gather == "bulk"))
        torch.cuda.synchronize()
        for raw in guards:
            assert bool((raw[:G] == 0xA5).all()) and bool((raw[-G:] == 0xA5).all()), (N, gather)
        s = env.state.long()
        assert torch.equal(env.obs_buf["segmentation"], env.dw.plane_view("segmentation")[s])
        assert torch.equal(env.goal_buf["rgb"], env.dw.plane_view("rgb")[env.goal.long()])
        assert env.episode_stats()["resets"] > N


@pytest.mark.parametrize("hw,cell,out", [((174, 174), 4, (42, 42)), ((18, 22), 2, (8, 10))])
def test_frames_that_are_not_a_multiple_of_16_bytes(hw, cell, out):
    """The reference's native frame size is 174 x 174 (GraphResize default, graph/core.py:43-49; 90,828-byte
    rgb frames, PC / deconv targets 42 x 42 - models/goal.py:151-156): records and batch rows are padded to 16
    bytes, every gather variant and every target builder must still return the exact frames."""
    import torch
    from oracle import rollout as orl
    planes = ("rgb", "depth", "segmentation")
    scene = H.scenes.make_maze_scene((7, 7), 0.2, 3, n_goals=2, planes=planes, frame_hw=hw)
    world = T.compile_world([scene], T.GYM_GRAPH)
    dw = vn.DeviceWorld(world)
    for p in planes:
        assert np.array_equal(dw.plane_view(p).cpu().numpy(), scene.plane_frames(p)), p
    for N, gather in ((12, "auto"), (300, "bulk"), (300, "ldg")):     # 212 KB records: sliced, never one persistent launch
        env = vn.GraphVecEnv(world, N, seed=2, max_episode_steps=6, obs_layout="aux5", gather=gather, device_world=dw)
        env.set_complexity(0.4)
        (obs, lar) = env.reset()
        rng = np.random.RandomState(4)
        for t in range(25):
            (obs, lar), r, d, infos = env.step(rng.randint(0, 4, N))
        s, g = env.state.cpu().numpy(), env.goal.cpu().numpy()
        want = (scene.plane_frames("rgb", s), scene.plane_frames("rgb", g), scene.plane_frames("depth", s),
                scene.plane_frames("segmentation", s), scene.plane_frames("segmentation", g))
        for leaf, w in zip(obs, want):
            assert leaf.shape == w.shape and np.array_equal(leaf.cpu().numpy(), w), (N, gather)
        assert env.episode_stats()["resets"] > N
    # target builders and the float policy input on padded records
    st = torch.randint(0, world.n_states, (3, 5), dtype=torch.int32)
    fr = scene.plane_frames("rgb", st.numpy().reshape(-1)).reshape(3, 5, hw[0], hw[1], 3)
    pc = vn.rollout.pixel_control_reward(dw, st, cell, out)
    np.testing.assert_allclose(pc.cpu().numpy(), orl.pixel_control_reward(orl.u8_to_policy_input(fr), cell, out),
                               rtol=1e-5, atol=1e-7)
    x = vn.rollout.policy_input(dw, st, "rgb").cpu().numpy()
    assert np.array_equal(x, orl.u8_to_policy_input(fr))
    dep = scene.plane_frames("depth", st.numpy().reshape(-1)).reshape(3, 5, hw[0], hw[1], 1)
    aux = vn.rollout.auxiliary_target(dw, st, "depth", cell, out)
    np.testing.assert_allclose(aux.cpu().numpy(), orl.aux_target(orl.u8_to_policy_input(dep), cell, out),
                               rtol=1e-5, atol=1e-7)


def test_philox_reset_path_thor_cached_task_list():
    """THORCachedEnv's intent (gym_thor_cached.py:45-53: a task list, start with dist[start][goal] > 0, dict
    observation) on the device's own Philox reset path, against the oracle env driven by the oracle restatement
    of that sampling built from the h5 'shortest_path_distance' table - not from the product's compiled tables."""
    seed, N, Tn, limit = 77, 10, 250, 9
    scene = H.scenes.make_maze_scene((8, 8), 0.2, 6, n_goals=1, planes=("rgb",))
    osc = oenvs.OracleScene(scene)
    _, graph, spd = gu.h5_tables(scene.maze, osc.graph)
    goals = [3, 41, 90, 17]
    world = T.compile_world([scene], T.THOR_CACHED, tasks=[(0, g) for g in goals])
    assert np.array_equal(world.adj, graph)
    env_tasks = np.tile(np.array([[0, len(goals)]], np.int32), (N, 1))
    env = vn.GraphVecEnv(world, N, seed=seed, max_episode_steps=limit, obs_layout="dict", unreal_wrapper=False,
                         env_tasks=env_tasks)
    frames = scene.plane_frames("rgb")
    oes = []
    for i in range(N):
        e = oenvs.ThorCachedEnv(graph, frames, spd, tasks=goals, dict_obs=True)
        cands = [(list(np.nonzero(spd[:, g] > 0)[0]), spd[np.nonzero(spd[:, g] > 0)[0], g]) for g in goals]
        e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t: None, False)
        oes.append(ovec.TimeLimit(e, limit))
    obs = env.reset()
    oobs = [e.reset() for e in oes]
    rng = np.random.RandomState(8)
    n_done = 0
    for t in range(Tn):
        a = rng.randint(0, 4, N)
        obs, rew, done, infos = env.step(a)
        img, goal = obs["image"].cpu().numpy(), obs["goal"].cpu().numpy()
        for i, e in enumerate(oes):
            o, r, d, info = e.step(int(a[i]))
            if d:
                o = e.reset()                       # the VecEnv worker's auto-reset
            assert bool(done[i]) == bool(d) and f32bits(rew[i]) == f32bits(np.float32(r)), (t, i)
            assert int(env.state[i]) == e.env._current_state_idx and int(env.goal[i]) == e.env._current_goal_idx, (t, i)
            assert np.array_equal(img[i], o["image"]) and np.array_equal(goal[i], o["goal"]), (t, i)
            assert infos[i].get("TimeLimit.truncated") == info.get("TimeLimit.truncated"), (t, i)
        n_done += int(done.sum())
    assert n_done > N


def test_evaluation_service_success_rate_and_spl():
    """trainer.test()-style evaluation roll-outs: the shortest-path policy scores success 1 and SPL 1; a random
    policy under a short time limit scores less on both; the hardness schedule is the reference's LinearSchedule."""
    import torch
    E = vn.evaluation
    scene = H.scenes.make_thor_scene(200, (20, 25), seed=3, n_goals=3, planes=("rgb",))
    world = T.compile_world([scene], T.GYM_GRAPH)
    env = vn.GraphVecEnv(world, 256, seed=5, max_episode_steps=60, obs_layout="frame", host_outputs=False)
    res = E.evaluate(env, lambda obs: env.optimal_actions()[0], episodes=1000)
    assert res["episodes"] == 1000 and res["success_rate"] == 1.0 and abs(res["spl"] - 1.0) < 1e-12
    assert res["truncated_rate"] == 0.0 and res["reward"] == 1.0
    gen = torch.Generator(device="cuda").manual_seed(0)
    rnd = E.evaluate(env, lambda obs: torch.randint(0, 4, (256,), device="cuda", generator=gen, dtype=torch.int32),
                     episodes=1000)
    assert rnd["episodes"] == 1000 and rnd["success_rate"] < 0.9 and rnd["spl"] < rnd["success_rate"] + 1e-12
    assert rnd["truncated_rate"] > 0 and abs(rnd["success_rate"] + rnd["truncated_rate"] - 1.0) < 1e-12
    sched = E.LinearSchedule(0.3, 1.0, 200000)                       # thor_cached_auxiliary.py:45
    assert sched(0) == 0.3 and abs(sched(100000) - 0.65) < 1e-12 and sched(10 ** 7) == 1.0
    assert E.apply_hardness_schedule(env, sched, 50000) == sched(50000) and env.dw.complexity == sched(50000)


@pytest.mark.parametrize("case", range(10))
def test_randomised_configurations_against_oracle(case):
    """Differential fuzz: random maze, family, goals, batch size, time limit, reward triple, curriculum schedule,
    launch mode and row skipping - every step compared with the oracle envs driven by the same Philox reset
    stream (states, float bits of rewards and last_action_reward, dones, observation bytes, info keys)."""
    rng = np.random.RandomState(1000 + case)
    oriented = bool(rng.randint(2))
    shape = (int(rng.randint(4, 12)), int(rng.randint(4, 12)))
    n_goals = int(rng.randint(1, 4))
    planes = ("rgb", "depth", "segmentation") if oriented else ("rgb",)
    scene = H.scenes.make_maze_scene(shape, float(rng.uniform(0.05, 0.35)), int(rng.randint(1 << 20)), n_goals=n_goals,
                                     oriented=oriented, planes=planes)
    if scene.n_cells < 3:
        pytest.skip("degenerate maze")
    fam = T.GYM_GRAPH if oriented else T.SIMPLE_GRAPH
    world = T.compile_world([scene], fam)
    osc = oenvs.OracleScene(scene)
    N = int(rng.choice([1, 2, 7, 33, 150, 260]))
    limit = int(rng.randint(2, 40))
    rewards = tuple(float(x) for x in rng.choice([1.0, 0.0, -0.01, 0.5, -1.0, 2.0], 3))
    seed = int(rng.randint(1 << 30))
    gather = str(rng.choice(["auto", "bulk", "ldg", "fused", "persistent"] if N <= 148 else
                            ["auto", "bulk", "ldg", "persistent"]))
    skip = bool(rng.randint(2))
    goals = list(scene.goals)
    env_tasks = np.tile(np.array([[0, len(goals)]], np.int32), (N, 1)) if oriented else None
    env = vn.GraphVecEnv(world, N, seed=seed, max_episode_steps=limit, rewards=rewards, gather=gather,
                         skip_unchanged=skip, obs_layout="aux5" if oriented else "frame", env_tasks=env_tasks)
    oes = []
    for i in range(N):
        if oriented:
            e = oenvs.GymGraphAuxiliaryEnv(osc, goals=goals, rewards=rewards)
            cands = [gu.initial_state_candidates(scene.maze, osc.graph, osc.optimal_actions, gl) for gl in goals]
            e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t, e=e: e.optimal_distance(), False)
        else:
            gl = goals[i % len(goals)]                    # default env_tasks: one task per env, round-robin
            e = oenvs.SimpleGraphEnv(osc, goal=gl, rewards=rewards)
            cands = [gu.initial_position_candidates(scene.maze, osc.graph, gl)]
            e.reset_source = ovec.PhiloxResetSource(seed, i, cands, lambda t, e=e: e.optimal_distance(), True)
        oes.append(ovec.RewardCollector(ovec.TimeLimit(e, limit)))
    ov = ovec.VecEnv(oes)
    c0 = None if rng.randint(3) == 0 else float(rng.uniform(0.0, 1.0))
    env.set_complexity(c0)
    [e.set_complexity(c0) for e in oes]
    (obs, lar), (oobs, olar) = env.reset(), ov.reset()
    for t in range(120):
        if t and t % 37 == 0:                             # the curriculum moves while episodes are running
            c = None if rng.randint(4) == 0 else float(rng.uniform(0.0, 1.0))
            env.set_complexity(c)
            [e.set_complexity(c) for e in oes]
        a = rng.randint(0 if oriented else -1, 4, size=N)
        (obs, lar), rew, done, infos = env.step(a)
        (oobs, olar), orew, odone, oinfos = ov.step(a)
        assert np.array_equal(done, odone), (case, t)
        assert np.array_equal(f32bits(rew), f32bits(orew)), (case, t)
        assert np.array_equal(f32bits(lar.cpu().numpy()), f32bits(olar)), (case, t)
        assert [oe.state for oe in oes] == env.states(), (case, t)
        leaves = obs if isinstance(obs, tuple) else (obs,)
        oleaves = oobs if isinstance(oobs, tuple) else (oobs,)
        for x, y in zip(leaves, oleaves):
            y = y if y.dtype == np.uint8 else np.rint(y * 255).astype(np.uint8)
            assert np.array_equal(x.cpu().numpy(), y), (case, t)
        for i in (0, N // 2, N - 1):
            info, oinfo = infos[i], oinfos[i]
            assert info.get("episode") == oinfo.get("episode"), (case, t, i)
            assert info.get("TimeLimit.truncated") == oinfo.get("TimeLimit.truncated"), (case, t, i)
            assert info.get("state") == oinfo.get("state") and info.get("win") == oinfo.get("win"), (case, t, i)


def test_host_facing_c_abi_failure_paths():
    """vn_env_step_host / vn_host_wait_seq fail loudly instead of hanging or faulting: pageable host buffers are
    refused, a sequence word that never arrives ends the wait once the stream has drained, and the explicit FUSED
    variant refuses records that do not fit in shared memory."""
    import ctypes as C
    import torch
    L = vn.lib
    lib = L.load()
    scene = H.scenes.make_maze_scene((6, 6), 0.1, 1, n_goals=1, planes=("rgb",))
    world = T.compile_world([scene], T.GYM_GRAPH)
    env = vn.GraphVecEnv(world, 5, seed=1, obs_layout="frame")
    env.reset()
    stream = torch.cuda.current_stream().cuda_stream
    pageable = np.zeros(5, np.int32)
    rc = lib.vn_env_step_host(C.byref(env.dw.store), C.byref(env.dw.tables), C.byref(env._c_envs), C.byref(env._c_rules),
                              None, pageable.ctypes.data, None, C.byref(env._c_out_host), None, env.gather, stream)
    assert rc == -1 and b"pinned" in lib.vn_last_error()
    # nothing was enqueued: the env still steps normally afterwards
    _, r, d, _ = env.step(np.zeros(5, np.int32))
    assert r.shape == (5,)
    # a word nobody will ever publish: the wait notices that the stream is idle and returns an error
    torch.cuda.synchronize()
    never = (env._seq + 12345) & 0x7FFFFFFF
    rc = lib.vn_host_wait_seq(env._seq_host.data_ptr(), env._seq_words, never, stream, 5_000_000)
    assert rc == -2 and b"drained" in lib.vn_last_error()
    # 174 x 174 x (rgb + depth + segmentation) = 212 KB per record: too large for the fused launch when asked explicitly
    big = H.scenes.make_maze_scene((5, 5), 0.0, 1, n_goals=1, frame_hw=(174, 174))
    wbig = T.compile_world([big], T.GYM_GRAPH)
    with pytest.raises(L.VnError, match="fused"):
        vn.GraphVecEnv(wbig, 4, obs_layout="aux5", gather="fused")
    with pytest.raises(L.VnError, match="persistent"):               # ... and for whole-record CTAs of the persistent launch
        vn.GraphVecEnv(wbig, 4, obs_layout="aux5", gather="persistent")
    ok = vn.GraphVecEnv(wbig, 4, obs_layout="aux5", gather="auto")       # AUTO falls back to two launches
    ok.reset()
    ok.step(np.zeros(4, np.int32))
    assert ok.kernel_launches == 4


def test_cuda_graph_replays_with_an_odd_number_of_steps():
    """Back-to-back replays of a graph with an ODD number of steps, and eager pipelined steps right after a replay:
    the descriptor parity of the last captured step then equals that of the next first step, so that step must not
    overlap the previous gather (it runs in serial mode).  Large batch so that a gather is long enough to race."""
    import torch
    scene = H.scenes.make_thor_scene(150, (16, 20), seed=1, n_goals=3, planes=("rgb", "depth"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N, S = 4096, 5
    a = vn.GraphVecEnv(world, N, seed=2, max_episode_steps=6, host_outputs=False, obs_layout="rgbd_goal")
    b = vn.GraphVecEnv(world, N, seed=2, max_episode_steps=6, host_outputs=False, obs_layout="rgbd_goal",
                       device_world=a.dw)
    a.reset()
    b.reset()
    acts = torch.randint(0, 4, (S, N), device="cuda", dtype=torch.int32)
    extra = torch.randint(0, 4, (3, N), device="cuda", dtype=torch.int32)
    cap = b.capture_steps(acts)
    for rep in range(12):
        for t in range(S):
            a.step_enqueue(acts[t])
        cap.replay()
        if rep % 3 == 2:                      # eager pipelined steps straight after a replay
            for t in range(3):
                a.step_enqueue(extra[t])
                b.step_enqueue(extra[t], actions_ready=True)
        if rep % 4 == 3:
            torch.cuda.synchronize()
            assert torch.equal(a.state, b.state) and torch.equal(a._pack, b._pack)
            for x, y in zip(list(a.obs_buf.values()) + list(a.goal_buf.values()),
                            list(b.obs_buf.values()) + list(b.goal_buf.values())):
                assert torch.equal(x, y)
            s = b.state.long()
            assert torch.equal(b.obs_buf["rgb"], b.dw.plane_view("rgb")[s])
            assert torch.equal(b.goal_buf["rgb"], b.dw.plane_view("rgb")[b.goal.long()])
    assert a.episode_stats() == b.episode_stats()


def test_numpy_observation_mode_for_host_trainers():
    """numpy_obs: reset() / step() return numpy arrays in the reference's layout (what deep_rl's trainer gets from
    SubprocVecEnv), equal to the CUDA batches of a twin env, in uint8 and in float mode.  True hands out views of two
    alternating pinned staging sets (valid until the step after next), "copy" private arrays."""
    scene = H.scenes.make_maze_scene((10, 10), 0.25, 0, n_goals=2)
    world = T.compile_world([scene], T.GYM_GRAPH)
    for flt in (False, True):
        for mode in (True, "copy"):
            a = vn.GraphVecEnv(world, 6, seed=1, max_episode_steps=9, scaled_float=flt)
            b = vn.GraphVecEnv(world, 6, seed=1, max_episode_steps=9, scaled_float=flt, numpy_obs=mode, device_world=a.dw)
            (oa, la), (ob, lb) = a.reset(), b.reset()
            rng = np.random.RandomState(0)
            kept, prev = None, None
            for t in range(12):
                act = rng.randint(0, 4, 6)
                (oa, la), ra, da, _ = a.step(act)
                if prev is not None:
                    prev = (prev[0], prev[0].copy())
                (ob, lb), rb, db, _ = b.step(act)
                if prev is not None:
                    assert np.array_equal(*prev)      # the previous step's arrays survive one more step in both modes
                prev = (ob[0], None)
                assert all(isinstance(x, np.ndarray) for x in ob) and isinstance(lb, np.ndarray)
                assert all(np.array_equal(x.cpu().numpy(), y) for x, y in zip(oa, ob)) and np.array_equal(la.cpu().numpy(), lb)
                assert np.array_equal(ra, rb) and np.array_equal(da, db)
                if t == 3:
                    kept = (ob[0], ob[0].copy())
            if mode == "copy":
                assert np.array_equal(*kept)          # private arrays are never overwritten by later steps
            assert ob[0].dtype == (np.float32 if flt else np.uint8)


def test_make_vec_from_scene_pickles(tmp_path):
    """loaders.make_vec: the experiments' create_envs in one call - gym id, [(scene, goals)] task list, scene pickles
    in the reference's format (read without the reference package) -> a stepping GraphVecEnv."""
    scs = [H.scenes.make_maze_scene((7, 6), 0.2, 40 + k, n_goals=2, frame_hw=(84, 84), scene_id=k) for k in range(2)]
    files = {}
    for k, sc in enumerate(scs):
        files["scene-%d" % k] = str(tmp_path / ("scene-%d.pkl" % k))
        H.write_reference_style_pickle(sc, files["scene-%d" % k])
    tasks = [("scene-%d" % k, list(sc.goals)) for k, sc in enumerate(scs)]
    env = vn.loaders.make_vec("AuxiliaryGraph-v0", tasks, graph_files=files, scaled_float=True, seed=3)
    assert env.num_envs == 4 and env.max_episode_steps == 900                 # one env per (scene, goal), :66
    assert env.observation_space.spaces[0].spaces[0].shape == (3, 84, 84)
    env.set_hardness(0.01)
    (obs, lar) = env.reset()
    rng = np.random.RandomState(0)
    for _ in range(30):
        (obs, lar), r, d, infos = env.step(rng.randint(0, 4, 4))
    st = env.state.cpu().numpy()
    for i in range(4):
        k = i // 2                                                             # envs 0,1 -> scene 0; 2,3 -> scene 1
        local = int(st[i] - env.world.scene_base[k])
        want = ovec.transpose_scale(scs[k].plane_frames("rgb", [local])[0])
        assert np.array_equal(obs[0][i].cpu().numpy(), want)
    small = vn.loaders.make_vec("OrientedGraph-v0", tasks[:1], graph_files=files, screen_size=(42, 42), num_envs=8)
    assert small.reset()[0].shape == (8, 42, 42, 3) and small.max_episode_steps == 900


@pytest.mark.parametrize("oriented", [True, False])
def test_reset_sampling_distribution_matches_the_reference_weights(oriented):
    """The reference draws start states with np.random.choice(p=weights) (graph/util.py:132-142 oriented: uniform over
    the eligible set; :103-116 un-oriented: 0.9 / 0.1 over the near / far buckets) and goals with random.choice.  The
    device's Philox draws must follow the same distribution: chi-square of ~260k resets against the oracle's weights,
    and never a state outside the candidate set."""
    import torch
    from scipy import stats
    scene = H.scenes.make_maze_scene((9, 9), 0.2, 12, n_goals=2, oriented=oriented,
                                     planes=("rgb",))
    osc = oenvs.OracleScene(scene)
    fam = T.GYM_GRAPH if oriented else T.SIMPLE_GRAPH
    world = T.compile_world([scene], fam)
    N, rounds = 8192, 32
    goals = list(scene.goals)
    env_tasks = np.tile(np.array([[0, len(goals)]], np.int32), (N, 1))
    env = vn.GraphVecEnv(world, N, seed=2024, max_episode_steps=1, obs_layout="frame", env_tasks=env_tasks,
                         host_outputs=False)
    c = 0.35
    env.set_complexity(c)
    env.reset()
    counts = torch.zeros((len(goals), world.n_states), dtype=torch.int64, device="cuda")
    zeros = torch.zeros(N, dtype=torch.int32, device="cuda")
    for _ in range(rounds):                                   # TimeLimit 1: every step ends the episode and resets
        env.step_enqueue(zeros)
        counts.index_put_((env.task.long(), env.state.long()), torch.ones(N, dtype=torch.int64, device="cuda"),
                          accumulate=True)
    counts = counts.cpu().numpy()
    total = N * rounds
    # goals: uniform (random.choice, gym_graph/graph.py:47)
    per_goal = counts.sum(1)
    assert stats.chisquare(per_goal).pvalue > 1e-4, per_goal
    for t, goal in enumerate(goals):
        if oriented:
            pots, dists = gu.initial_state_candidates(scene.maze, osc.graph, osc.optimal_actions, goal)
            od = c * (int(np.max(osc.graph)) + 4 - 1) + 1    # gym_graph/graph.py:49-51
            w = gu.initial_state_weights(dists, od)
        else:
            pots, dists = gu.initial_position_candidates(scene.maze, osc.graph, goal)
            od = c * (int(np.max(osc.graph)) - 1) + 1        # graph/env.py:103-105
            w = gu.initial_position_weights(dists, od)
        idx = np.array([scene.state_index(p) for p in pots])
        assert counts[t].sum() == counts[t][idx].sum()        # nothing outside the candidate set
        live = w > 0
        assert counts[t][idx[~live]].sum() == 0               # nothing the reference gives zero weight
        obs, exp = counts[t][idx[live]], w[live] / w[live].sum() * per_goal[t]
        assert stats.chisquare(obs, exp).pvalue > 1e-4, (t, obs[:8], exp[:8])
    assert total == counts.sum()


def test_host_received_scalars_are_never_stale():
    """The host-facing step publishes its scalars through mapped pinned memory and one sequence word per block (one
    system-scope fence per block).  3,000 back-to-back steps of 4,096 envs at full speed: what the host received each
    step must equal what the same kernel logged on the device (vn_step_out_t.rec_*), element for element."""
    import torch
    scene = H.scenes.make_thor_scene(300, (30, 30), seed=6, n_goals=4, planes=("rgb", "depth"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N, Tn = 4096, 3000
    env = vn.GraphVecEnv(world, N, seed=12, max_episode_steps=9, obs_layout="rgbd_goal")
    env.set_complexity(0.2)
    env.reset()
    log_r = torch.zeros((Tn, N), dtype=torch.float32, device="cuda")
    log_d = torch.zeros((Tn, N), dtype=torch.uint8, device="cuda")
    log_s = torch.zeros((Tn, N), dtype=torch.int32, device="cuda")
    rng = np.random.RandomState(3)
    acts = rng.randint(0, 4, (64, N)).astype(np.int32)
    got_r, got_d, got_len = [], [], []
    for t in range(Tn):
        o = env._c_out_host
        o.rec_reward, o.rec_done, o.rec_state = log_r[t].data_ptr(), log_d[t].data_ptr(), log_s[t].data_ptr()
        obs, r, d, infos = env.step(acts[t % 64])
        got_r.append(r)
        got_d.append(d)
        if t % 500 == 499:
            got_len.append((t, infos._host()["episode_length"].copy(), d.copy()))
    torch.cuda.synchronize()
    assert np.array_equal(np.stack(got_r).view(np.uint32), log_r.cpu().numpy().view(np.uint32))
    assert np.array_equal(np.stack(got_d), log_d.cpu().numpy().astype(bool))
    assert log_d.sum() > Tn and (log_r != 0).sum() > 0
    for t, ln, d in got_len:                         # episode lengths are reported where done, and plausible
        assert ((ln[d] >= 1) & (ln[d] <= 9)).all()
    assert torch.equal(log_s[-1], env.obs_state)


@pytest.mark.parametrize("N", [7, 700])
def test_lazy_infos_stay_valid_after_later_steps(N):
    """The host-facing step copies rewards / dones out of the pinned host pack and lets `infos` read the rest in place
    from one of two alternating packs.  An `infos` object that is still alive when its pack is about to be reused gets
    a private copy first: infos kept from every step, evaluated only at the end, equal the ones evaluated right away on
    a twin env - in both the one-call step() and the step_async / step_wait forms."""
    scene = H.scenes.make_thor_scene(120, (14, 18), seed=5, n_goals=3, planes=("rgb", "depth"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    a = vn.GraphVecEnv(world, N, seed=4, max_episode_steps=7, obs_layout="rgbd_goal")
    b = vn.GraphVecEnv(world, N, seed=4, max_episode_steps=7, obs_layout="rgbd_goal", device_world=a.dw)
    a.reset()
    b.reset()
    rng = np.random.RandomState(8)
    kept, eager, kept_rd = [], [], []
    for t in range(40):
        act = rng.randint(0, 4, N).astype(np.int32)
        if t % 3 == 2:
            a.step_async(act)
            _, r, d, infos = a.step_wait()
        else:
            _, r, d, infos = a.step(act)
        _, rb, db, infos_b = b.step(act)
        kept.append(infos)                      # NOT evaluated yet; the pack it reads is reused two steps later
        kept_rd.append((r, d))
        eager.append(([dict(x) for x in infos_b], rb.copy(), db.copy()))
    for t, (infos, (r, d), (want, rb, db)) in enumerate(zip(kept, kept_rd, eager)):
        assert np.array_equal(r, rb) and np.array_equal(d, db), t      # rewards / dones are private copies
        assert [dict(x) for x in infos] == want, t
    assert sum("episode" in x for w, _, _ in eager for x in w) > N


def test_drop_in_under_the_reference_trainer_contract():
    """INTEGRATION.md section 1's replacement of create_envs, replayed against what the reference's OWN
    experiments/thor_cached_auxiliary.py produced (tests/golden/trainer_contract.npz: Trainer.create_env -> create_envs
    -> wrap() run unmodified, deep_rl's wrappers restated): native 174 x 174 frames, 4 envs, 5-tuple observation as
    float32 CHW / 255 + last_action_reward, TimeLimit 900, set_hardness(0.01) then 0.3.  Every leaf the trainer would
    receive is bit-identical; Trainer.create_model's reads of the spaces (:55) give the same Model arguments."""
    g = H.load("trainer_contract")
    scene = H.trainer_contract_scene(g)
    goal = tuple(int(v) for v in g["goal"])
    N = g["actions"].shape[1]
    world = T.compile_world([scene], T.GYM_GRAPH, tasks=[(0, goal)])
    starts = np.zeros(g["reset_start"].shape[:2], np.int32)
    for i in range(N):
        for k in range(int(g["reset_count"][i])):
            starts[i, k] = scene.state_index(tuple(int(v) for v in g["reset_start"][i, k]))
    env = vn.GraphVecEnv(world, num_envs=N, max_episode_steps=int(g["max_episode_steps"]), obs_layout="aux5",
                         unreal_wrapper=True, scaled_float=True, inject=(g["reset_choice"], starts))
    env.set_hardness(0.01)                                                  # thor_cached_auxiliary.py:68-70
    assert env.call_unwrapped("set_complexity", 0.01) == [None] * N
    # Trainer.create_model (:55): Model(observation_space.spaces[0].spaces[0].shape[0], action_space.n)
    sp = env.observation_space
    assert [sp.spaces[0].spaces[0].shape[0], env.action_space.n] == g["model_args"].tolist()
    assert tuple(sp.spaces[0].spaces[0].shape) == tuple(g["space_leaf_shapes"][0])
    assert tuple(sp.spaces[1].shape) == tuple(g["space_lar_shape"]) and env.num_envs == N
    # the reference declares leaves 1.. with its unused screen_size (172); the frames it emits are 174 - ours says 174
    assert [tuple(b.shape) for b in sp.spaces[0].spaces] == [tuple(s[1:]) for s in g["obs_shapes"]]
    H.check_trainer_contract_run(g, env, lambda x, i: H.crc(x[i].cpu().numpy()), env.set_hardness)
    assert env.episode_stats()["episodes"] == g["dones"].sum()


def test_host_route_equals_device_route_for_any_action_value():
    """The host-facing step (actions read by the scalar kernel from mapped pinned memory) and the device-resident step
    give the same states / rewards / dones for every int32 action value - out-of-range ones have collision semantics -
    and both record the action actually sent (rollout record, device copy of the host's actions)."""
    import torch
    scene = H.scenes.make_thor_scene(120, (14, 18), seed=5, n_goals=3, planes=("rgb",))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N = 700
    host = vn.GraphVecEnv(world, N, seed=4, max_episode_steps=7, obs_layout="frame")
    dev = vn.GraphVecEnv(world, N, seed=4, max_episode_steps=7, obs_layout="frame", host_outputs=False, device_world=host.dw)
    host.reset()
    dev.reset()
    rng = np.random.RandomState(2)
    rec = torch.zeros((2, N), dtype=torch.int32, device="cuda")
    for t in range(30):
        a = rng.randint(-2, 6, N).astype(np.int32)              # includes actions outside [0, 4)
        if t % 3 == 1:
            a[rng.randint(N)] = 1000
        if t % 3 == 2:
            a[rng.randint(N)] = -129
        host._c_out_host.rec_action = rec[0].data_ptr()
        _, r, d, _ = host.step(a)
        dev._c_out.rec_action = rec[1].data_ptr()
        _, rd, dd, _ = dev.step(torch.from_numpy(a).cuda())
        assert np.array_equal(r.view(np.uint32), rd.cpu().numpy().view(np.uint32)) and np.array_equal(d, dd.cpu().numpy()), t
        assert torch.equal(host.state, dev.state), t
        assert torch.equal(rec[0], rec[1]) and np.array_equal(rec[0].cpu().numpy(), a), t
        assert torch.equal(host.actions_dev, torch.from_numpy(a).cuda()), t      # the device copy of the host's actions
    assert host.episode_stats() == dev.episode_stats()


def test_auto_mode_switches_between_launch_modes_safely():
    """Under VN_GATHER_AUTO the launch mode of a step depends on whether it is pipelined (VN_STEP_ACTIONS_READY) - serial
    steps of a few thousand envs run as the persistent single launch, pipelined ones as scalar kernel + gather - so one
    env batch may switch modes from step to step.  A pipelined scalar kernel may only overlap its own batch's previous
    GATHER, never a one-launch step that is still writing the env state it reads (VN_STEP_NO_OVERLAP, set by the
    wrapper from the previous call's mode).  Random mode switching, resets and graph replays in between: every step
    equals an always-serial two-kernel twin."""
    import torch
    scene = H.scenes.make_thor_scene(150, (16, 20), seed=1, n_goals=3, planes=("rgb", "depth"))
    world = T.compile_world([scene], T.GYM_GRAPH)
    N, S = 3000, 300
    ref = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=11, host_outputs=False, obs_layout="rgbd_goal", gather="bulk")
    auto = vn.GraphVecEnv(world, N, seed=6, max_episode_steps=11, host_outputs=False, obs_layout="rgbd_goal",
                          device_world=ref.dw)
    ref.reset()
    auto.reset()
    rng = np.random.RandomState(0)
    acts = torch.randint(0, 4, (S, N), device="cuda", dtype=torch.int32)
    graph = auto.capture_steps(acts[:3])
    modes = set()
    t = 0
    while t < S - 3:
        if rng.rand() < 0.05:                                   # a 3-step graph replay in the middle
            acts[:3].copy_(acts[t:t + 3])
            graph.replay()
            for k in range(3):
                ref.step_enqueue(acts[t + k])
            t += 3
        else:
            ready = bool(rng.rand() < 0.5)
            auto.step_enqueue(acts[t], actions_ready=ready)
            modes.add((ready, auto._prev_mode))
            ref.step_enqueue(acts[t])
            t += 1
        if rng.rand() < 0.02:
            mask = torch.from_numpy(rng.rand(N) < 0.3).cuda()
            auto.reset(mask)
            ref.reset(mask)
        if t % 7 == 0:
            torch.cuda.synchronize()
            assert torch.equal(ref._pack, auto._pack) and torch.equal(ref.state, auto.state), t
            for x, y in zip(list(ref.obs_buf.values()) + list(ref.goal_buf.values()),
                            list(auto.obs_buf.values()) + list(auto.goal_buf.values())):
                assert torch.equal(x, y), t
    # both modes were exercised: serial -> persistent launch, pipelined -> two kernels
    assert (False, vn.lib.MODE_PERSISTENT) in modes and (True, vn.lib.MODE_SPLIT) in modes
    assert ref.episode_stats()["episodes"] > N


def test_golden_thor_cached_task_list_as_written():
    """the reference's unfinished multi-scene THORCachedEnv (environments/gym_thor_cached.py), run AS
    WRITTEN through the harness (tests/golden/thor_cached_tasks.npz: two scenes, four (scene, goal) tasks).  The device
    path - two scenes resident in one store, `obs_layout="dict"`, THOR_CACHED rules - reproduces task choices, states,
    rewards (the -0.0 of :80 included), terminals, truncations and the observation bytes: the uint8 frames equal the
    raw pair of observe() (:52-53) and, divided by 255 in float32, the {'image', 'goal'} dict of process() (:89-92)."""
    g = H.load("thor_cached_tasks")
    scs = H.thor_cached_task_scenes(g)
    tasks = [(int(s), int(gl)) for s, gl in zip(g["task_scene"], g["task_goal"])]
    world = T.compile_world(scs, T.THOR_CACHED, tasks=tasks)
    actions = g["actions"]
    Tn, N = actions.shape
    base = world.scene_base
    starts = np.zeros_like(g["reset_start"])
    for i in range(N):
        for k in range(int(g["reset_count"][i])):
            starts[i, k] = int(base[tasks[int(g["reset_choice"][i, k])][0]]) + int(g["reset_start"][i, k])
    env = vn.GraphVecEnv(world, N, max_episode_steps=int(g["max_episode_steps"]), obs_layout="dict", unreal_wrapper=False,
                         env_tasks=np.tile(np.array([[0, len(tasks)]], np.int32), (N, 1)), inject=(g["reset_choice"], starts))
    local = lambda e: [int(s) - int(base[world.scene_of_state(int(s))]) for s in e.state.cpu().numpy()]
    u8 = lambda ob, i: [H.crc(ob["image"][i].cpu().numpy()), H.crc(ob["goal"][i].cpu().numpy())]
    f32 = lambda ob, i: [H.crc(ob["image"][i].cpu().numpy().astype(np.float32) / 255.0),
                         H.crc(ob["goal"][i].cpu().numpy().astype(np.float32) / 255.0)]
    ob = env.reset()
    assert local(env) == g["reset_states"].tolist()
    assert [u8(ob, i) for i in range(N)] == g["reset_obs_crc"].tolist()
    for t in range(Tn):
        ob, r, d, infos = env.step(actions[t])
        h = infos._host()
        assert np.array_equal(d, g["dones"][t]) and np.array_equal(h["truncated"] == 1, g["truncated"][t]), t
        assert np.array_equal(d & ~(h["truncated"] == 1), g["env_dones"][t]), t
        assert np.array_equal(f32bits(r), f32bits(g["rewards"][t])), t            # bitwise: -0.0 on plain steps
        assert [int(s) - int(base[world.scene_of_state(int(s))]) for s in h["info_state"]] == g["states"][t].tolist(), t
        assert local(env) == g["post_states"][t].tolist(), t
        # after a reset the golden holds the raw uint8 pair of observe(), otherwise process()'s float32 / 255 dict
        want = g["obs_crc"][t].tolist()
        assert [(u8 if d[i] else f32)(ob, i) for i in range(N)] == want, t
    assert env.episode_stats()["episodes"] == g["dones"].sum() and g["env_dones"].sum() >= 3
