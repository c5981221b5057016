"""Pins the ORACLE to the golden vectors produced by the unmodified reference classes
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import helpers as H
from oracle import envs as oenvs
from oracle import graph_util as gu
from oracle import philox
from oracle import rollout as orl


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    f = lambda c, k: [int(x) for x in philox.philox4x32(np.array(c, np.uint32), k)]
    assert f([0, 0, 0, 0], (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f([0xffffffff] * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_shortest_paths_and_candidates_match_reference():
    g = H.load("graph_util")
    cs = g["complexities"]
    for k in range(int(g["n_mazes"])):
        maze = g["maze%d" % k]
        dist, act = gu.compute_shortest_path_data(maze)
        assert np.array_equal(dist, g["dist%d" % k])
        assert np.array_equal(act, g["act%d" % k])
        maxd = int(dist.max())
        for gi, goal in enumerate(g["goals%d" % k]):
            goal = tuple(int(v) for v in goal)
            pots, d = gu.initial_state_candidates(maze, dist, act, goal)
            pots2, d2 = gu.initial_position_candidates(maze, dist, goal[:2])
            for ci, c in enumerate(cs):
                od = None if c < 0 else c * (maxd + 4 - 1) + 1
                w = gu.initial_state_weights(d, od)
                assert np.array_equal(w, g["w_state_%d_%d_%d" % (k, gi, ci)])
                od = None if c < 0 else c * (maxd - 1) + 1
                w = gu.initial_position_weights(d2, od)
                assert np.array_equal(w, g["w_position_%d_%d_%d" % (k, gi, ci)])


def _gym_graph(name, aux):
    g = H.load(name)
    scene = H.scene_from_golden(g, True, ("rgb", "depth", "segmentation"))
    osc = oenvs.OracleScene(scene)
    goals = [tuple(int(v) for v in x) for x in g["goals"]]
    goals = goals if bool(g["goals_is_list"]) else goals[0]
    cls = oenvs.GymGraphAuxiliaryEnv if aux else oenvs.GymGraphEnv
    rec = H.replay_oracle(g, lambda i: cls(osc, goals=goals, rewards=tuple(g["rewards_cfg"])), 5 if aux else 1)
    assert int(g["largest_distance"]) == int(np.max(osc.graph))
    H.assert_record_equal(rec, g)


def test_gym_graph_auxiliary_env():
    _gym_graph("gym_graph_aux", True)


def test_gym_graph_oriented_env():
    _gym_graph("gym_graph_oriented", False)


def test_graph_env_simple():
    g = H.load("graph_env_simple")
    scene = H.scene_from_golden(g, False, ("rgb",))
    osc = oenvs.OracleScene(scene)
    rec = H.replay_oracle(g, lambda i: oenvs.SimpleGraphEnv(osc, rewards=tuple(g["rewards_cfg"])), 1, width=2)
    H.assert_record_equal(rec, g)


def test_graph_env_multiple():
    g = H.load("graph_env_multiple")
    oscs = []
    for k in range(int(g["n_graphs"])):
        sc = H.scenes.GridScene(g["maze%d" % k], [tuple(int(v) for v in g["goal%d" % k])], False, (84, 84), ("rgb",),
                                frame_seed=int(g["frame_seed%d" % k]), scene_id=k)
        oscs.append(oenvs.OracleScene(sc))
    rec = H.replay_oracle(g, lambda i: oenvs.MultipleGraphEnv(oscs), 1, width=2)
    H.assert_record_equal(rec, g)


def test_graph_env_oriented_never_terminates():
    g = H.load("graph_env_oriented")
    scene = H.scene_from_golden(g, True, ("rgb", "depth", "segmentation"))
    osc = oenvs.OracleScene(scene)
    goal = tuple(int(v) for v in g["goals"][0])
    rec = H.replay_oracle(g, lambda i: oenvs.GraphEnvOriented(osc, goal), 1)
    H.assert_record_equal(rec, g)
    assert not g["env_dones"].any() and g["truncated"].any()


def test_thor_cached_env():
    g = H.load("thor_cached")
    scene = H.scene_from_golden(g, True, ("rgb",))
    dist, _ = gu.compute_shortest_path_data(scene.maze)
    locs, graph, spd = gu.h5_tables(scene.maze, dist)
    assert np.array_equal(graph, g["graph"]) and H.crc(spd) == int(g["spd_crc"])
    obs = scene.plane_frames("rgb")
    rec = H.replay_oracle(g, lambda i: oenvs.ThorCachedEnv(graph, obs, spd), 2,
                          state_of=lambda e: e._current_state_idx, width=1)
    # the reference returns skimage-resized float64 frames (== uint8 / 255 at equal size); the oracle
    # returns the raw uint8 frames, so compare CRCs of uint8/255 for the observation leaves
    H.assert_record_equal(rec, g, obs=False)
    # recompute the crc on float64/255 for a sample of steps
    e = oenvs.ThorCachedEnv(graph, obs, spd)
    e.reset_source = H.StreamSource(g["reset_goal"][0], g["reset_start"][0], g["reset_count"][0])
    ob = e.reset()
    assert [H.crc(x.astype(np.float64) / 255.0) for x in ob] == [int(v) for v in g["reset_obs_crc"][0]]
    assert np.signbit(g["rewards"][g["rewards"] == 0]).any()      # the -0.0 of cached.py:84 is in the fixture


def test_maze_render_hoist():
    g = H.load("maze_render")
    sc = H.scenes.GridScene(g["maze"], [tuple(int(v) for v in g["goal"])], False, (84, 84), ("rgb",))
    fr = H.scenes.render_maze_frames(sc, tuple(int(v) for v in g["goal"]), (84, 84))
    assert [H.crc(x) for x in fr] == [int(v) for v in g["resized_u8_crc"]]


def test_aux_target_matches_reference_function():
    g = H.load("aux_target")
    rng = np.random.RandomState(int(g["x_seed"]))
    x = rng.randint(0, 256, size=(2, 3, 3, 84, 84)).astype(np.float32) / np.float32(255.0)
    np.testing.assert_allclose(orl.aux_target(x, 4, (20, 20)), g["y20"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orl.aux_target(x, 4, None), g["y21"], rtol=1e-6, atol=1e-7)
    d = rng.randint(0, 256, size=(2, 3, 1, 84, 84)).astype(np.float32) / np.float32(255.0)
    np.testing.assert_allclose(orl.aux_target(d, 4, (20, 20)), g["yd"], rtol=1e-6, atol=1e-7)


def test_rollout_oracle_agrees_with_the_torch_formulation():
    """The D5 / A15 restatements (numpy, explicit summation order) against the same computation written with the
    torch ops deep_rl uses (autocrop -> abs diff -> F.avg_pool2d -> channel mean; F.avg_pool2d for the aux targets).
    torch's CPU pooling sums in its own order, hence the 1e-6 tolerance - the north-star bound is 1e-5."""
    import torch
    import torch.nn.functional as F
    from oracle import rollout as orl
    rng = np.random.RandomState(0)
    obs = (rng.randint(0, 256, size=(3, 6, 3, 84, 84)).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    for cell, out in ((4, (20, 20)), (4, None), (3, (27, 26))):
        x = torch.from_numpy(obs)
        h, w = x.shape[3:]
        nh, nw = ((h // cell) * cell, (w // cell) * cell) if out is None else (out[0] * cell, out[1] * cell)
        top, left = (h - nh) // 2, (w - nw) // 2
        xc = x[:, :, :, top:top + nh, left:left + nw]
        d = (xc[:, 1:] - xc[:, :-1]).abs()
        p = F.avg_pool2d(d.reshape(-1, *d.shape[2:]), cell, stride=cell).mean(1, keepdim=True)
        want = p.view(d.shape[0], d.shape[1], 1, nh // cell, nw // cell).numpy()
        np.testing.assert_allclose(orl.pixel_control_reward(obs, cell, out), want, rtol=1e-6, atol=1e-7)
        a = F.avg_pool2d(xc.reshape(-1, *xc.shape[2:]), cell, stride=cell).view(3, 6, 3, nh // cell, nw // cell).numpy()
        np.testing.assert_allclose(orl.aux_target(obs, cell, out), a, rtol=1e-6, atol=1e-7)
    # n-step returns: the recurrence written with torch tensors step by step
    r = torch.from_numpy(rng.rand(5, 9).astype(np.float32))
    dn = torch.from_numpy((rng.rand(5, 9) < 0.2))
    v = torch.from_numpy(rng.randn(5).astype(np.float32))
    ret = torch.zeros(5, 10)
    ret[:, -1] = v * (1.0 - dn[:, -1].float())
    for t in reversed(range(9)):
        ret[:, t] = r[:, t] + 0.99 * ret[:, t + 1] * (1.0 - dn[:, t].float())
    got = orl.nstep_returns(r.numpy(), dn.numpy(), v.numpy(), 0.99)
    np.testing.assert_allclose(got, ret[:, :-1].numpy(), rtol=1e-6, atol=1e-7)


def test_trainer_contract_oracle_wrapper_stack():
    """tests/golden/trainer_contract.npz was recorded by running the reference's OWN experiments/thor_cached_auxiliary.py
    (Trainer.create_env -> create_envs -> wrap(), :50-71) unmodified.  The oracle env under the oracle's restated
    wrapper stack reproduces it: leaf order, shapes, float32 CHW bytes, last_action_reward, rewards, dones, episodes -
    and the observation space Trainer.create_model reads (:55), including the 172-vs-174 screen_size quirk (A10)."""
    from oracle import vec as ovec
    g = H.load("trainer_contract")
    scene = H.trainer_contract_scene(g)
    osc = oenvs.OracleScene(scene)
    goal = tuple(int(v) for v in g["goal"])
    N = g["actions"].shape[1]

    def make(i):
        e = oenvs.GymGraphAuxiliaryEnv(osc, goals=goal, screen_size=(172, 172))      # default_args(), :83
        e.reset_source = H.StreamSource(g["reset_choice"][i], g["reset_start"][i], g["reset_count"][i])
        e = ovec.TimeLimitWrapper(e, int(g["max_episode_steps"]))
        return ovec.UnrealEnvBaseWrapper(ovec.ScaledFloatFrameWrapper(ovec.TransposeImageWrapper(
            ovec.RewardCollectorWrapper(e))))                                        # wrap(), :59-64

    env = ovec.InProcessVecEnv([(lambda i=i: make(i)) for i in range(N)])
    env.call_unwrapped("set_complexity", 0.01)                                       # :68-70
    sp = env.observation_space
    assert np.array_equal(np.array([b.shape for b in sp.spaces[0].spaces]), g["space_leaf_shapes"])
    assert tuple(sp.spaces[1].shape) == tuple(g["space_lar_shape"]) and env.action_space.n == int(g["action_n"])
    assert [sp.spaces[0].spaces[0].shape[0], env.action_space.n] == g["model_args"].tolist()
    H.check_trainer_contract_run(g, env, lambda x, i: H.crc(x[i]), lambda c: env.call_unwrapped("set_complexity", c))


def test_thor_cached_task_list_env_as_written():
    """tests/golden/thor_cached_tasks.npz: the reference's unfinished THORCachedEnv (gym_thor_cached.py) run as written on
    two scenes and four (scene, goal) tasks.  The oracle restatement reproduces task choices, start states, rewards
    (-0.0 included), terminals, time-limit truncations and the bytes of both observation forms: the raw uint8 pair of
    observe() after a reset and the float32 / 255 dict of process()."""
    g = H.load("thor_cached_tasks")
    scs = H.thor_cached_task_scenes(g)
    tables = {}
    for k, sc in enumerate(scs):
        dist, _ = gu.compute_shortest_path_data(sc.maze)
        _, graph, spd = gu.h5_tables(sc.maze, dist)
        tables[k] = dict(transition_graph=graph, observations=sc.plane_frames("rgb"), shortest_path_distances=spd)
    tasks = [(int(s), int(gl)) for s, gl in zip(g["task_scene"], g["task_goal"])]

    class Leaves:
        def __init__(self, e):
            self.e = e

        def __getattr__(self, name):
            return getattr(self.e, name)

        def reset(self):
            return self.e.reset()

        def step(self, a):
            st, r, term, info = self.e.step(a)
            return (st["image"], st["goal"]), r, term, info

    def make(i):
        e = Leaves(oenvs.ThorCachedTasksEnv(tables, tasks))
        return e

    actions = g["actions"]
    T_, N = actions.shape
    envs = []
    for i in range(N):
        e = make(i)
        e.e.reset_source = H.StreamSource(g["reset_choice"][i], g["reset_start"][i], g["reset_count"][i])
        envs.append(e)
    leafs = lambda ob: [H.crc(x) for x in ob]
    assert np.array_equal(np.array([leafs(e.reset()) for e in envs], np.uint32), g["reset_obs_crc"])
    assert [e.e.state for e in envs] == g["reset_states"].tolist()
    elapsed = [0] * N
    for t in range(T_):
        for i, e in enumerate(envs):
            ob, r, d, info = e.step(int(actions[t, i]))
            assert e.e.state == g["states"][t, i] and bool(d) == bool(g["env_dones"][t, i]), (t, i)
            elapsed[i] += 1
            trunc = False
            if elapsed[i] >= int(g["max_episode_steps"]):
                trunc, d = not d, True
            if d:
                ob = e.reset()
                elapsed[i] = 0
            assert trunc == bool(g["truncated"][t, i]) and bool(d) == bool(g["dones"][t, i]), (t, i)
            assert np.float64(r).view(np.uint64) == g["rewards"][t, i].view(np.uint64), (t, i)
            assert leafs(ob) == g["obs_crc"][t, i].tolist() and e.e.state == g["post_states"][t, i], (t, i)
    assert g["env_dones"].sum() >= 3 and np.signbit(g["rewards"][g["rewards"] == 0]).any()
