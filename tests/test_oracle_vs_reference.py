"""Live check against the UNMODIFIED reference tree (this container only: /root/reference is not on the
GPU box, where these tests skip).  Regenerates every golden fixture from the reference classes and
requires the committed fixtures to be reproduced array for array - so the fixtures the GPU tests replay
are provably what the reference computes, not a stale or hand-edited copy."""
import importlib.util
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.needs_reference


def _load_generator(tmpdir):
    path = os.path.join(H.GOLDEN, "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden_live", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.HERE = str(tmpdir)          # write the regenerated fixtures somewhere else
    return mod


def test_committed_goldens_reproduce_from_live_reference(tmp_path, capsys):
    from oracle import ref_harness as rh
    mg = _load_generator(tmp_path)
    ref = rh.ref_modules()
    mg.gen_graph_util(ref)
    mg.gen_gym_graph(ref)
    mg.gen_graph_env(ref)
    mg.gen_maze_render(ref)
    mg.gen_thor_cached(ref)
    mg.gen_aux_target(rh.ref_aux_trainer())
    mg.gen_trainer_contract(ref)
    mg.gen_thor_cached_tasks(ref)
    names = sorted(f for f in os.listdir(H.GOLDEN) if f.endswith(".npz"))
    assert names == sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz")) and len(names) == 11
    for f in names:
        a, b = np.load(os.path.join(H.GOLDEN, f)), np.load(os.path.join(tmp_path, f))
        assert sorted(a.files) == sorted(b.files), f
        for k in a.files:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (f, k)


def test_reference_reset_cost_is_what_the_cpu_baseline_models():
    """The reference enumerates every free cell on EVERY reset (graph/util.py:119-143); the CPU baseline's
    ReferenceStyleResetSource must do the same work (same candidate list, same weights)."""
    from oracle import ref_harness as rh, graph_util as gu
    ref = rh.ref_modules()
    maze = H.scenes.random_maze((10, 10), 0.25, 0)
    dist, act = ref.util.compute_shortest_path_data(maze)
    g = type("G", (), {})()
    g.maze, g.graph, g.optimal_actions = maze, dist, act
    cells = np.argwhere(maze)
    goal = (int(cells[9][0]), int(cells[9][1]), 2)
    inj = rh.InjectedChoice([12345])
    orig = np.random.choice
    np.random.choice = inj
    try:
        s = ref.util.sample_initial_state(g, goal, optimal_distance=7.5)
    finally:
        np.random.choice = orig
    pots, d = gu.initial_state_candidates(maze, dist, act, goal)
    n, p, idx = inj.calls[0]
    assert n == len(pots) and np.array_equal(p, gu.initial_state_weights(d, 7.5)) and pots[idx] == s


def test_loader_consumes_a_reference_pickle(tmp_path):
    """A scene pickled the way the reference stores them (ThorGridWorld, graph/multi_graph_no_tp.py) and read
    back with the reference's own load_graph (graph/util.py:69-79) goes through loaders.scene_from_thor_grid_world
    into the same tables / frames as the synthetic scene it was exported from."""
    import importlib
    import pickle
    from oracle import ref_harness as rh
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    T = vn.tables
    ref = rh.ref_modules()
    scene = H.scenes.make_maze_scene((9, 8), 0.2, 31, n_goals=2, planes=("rgb", "depth", "segmentation"))
    X, Y = scene.maze.shape
    arrs = {}
    for plane, c in (("rgb", 3), ("depth", 1), ("segmentation", 3)):
        a = np.zeros((X, Y, 4, 84, 84, c), np.uint8)
        a[scene.cells[:, 0], scene.cells[:, 1]] = scene.plane_frames(plane).reshape(scene.n_cells, 4, 84, 84, c)
        arrs[plane] = a
    world = ref.thor_world.ThorGridWorld(scene.maze.copy(), arrs["rgb"], arrs["depth"], arrs["segmentation"])
    world.goals = list(scene.goals)
    path = tmp_path / "scene.pkl"
    with open(path, "wb") as f:
        pickle.dump(world, f)
    graph = ref.util.load_graph(str(path))                       # adds graph.graph / graph.optimal_actions
    loaded = vn.loaders.scene_from_thor_grid_world(graph, graph.goals)
    w0, w1 = T.compile_world([scene], T.GYM_GRAPH), T.compile_world([loaded], T.GYM_GRAPH)
    assert np.array_equal(w0.adj, w1.adj) and np.array_equal(w0.cand_state, w1.cand_state)
    assert all(t.max_dist == int(np.max(graph.graph)) for t in w1.tasks)          # largest_distance, graph.py:35
    for p in ("rgb", "depth", "segmentation"):
        assert np.array_equal(loaded.plane_frames(p), scene.plane_frames(p))
    # GraphResize hoisted: what the reference would render per step at another screen size
    rz = ref.core.GraphResize(graph, (44, 44))
    small = vn.loaders.scene_from_thor_grid_world(graph, graph.goals, screen_size=(44, 44))
    for s in (0, 17, scene.n_states - 1):
        x, y, r = scene.state_tuple(s)
        rgb, depth, seg = rz.render((x, y), r, modes=["rgb", "depth", "segmentation"])
        assert np.array_equal(small.plane_frames("rgb", [s])[0], rgb)
        assert np.array_equal(small.plane_frames("depth", [s])[0], depth)
        assert np.array_equal(small.plane_frames("segmentation", [s])[0], seg)


def test_reference_trainer_accepts_the_drop_in_spaces():
    """experiments/thor_cached_auxiliary.py imported unmodified: ``Trainer.create_model`` (:54-56) builds the model from
    ``self.env.observation_space.spaces[0].spaces[0].shape[0]`` and ``self.env.action_space.n``.  With ``self.env`` = the
    reference's own ``create_envs`` result and with the spaces of the INTEGRATION.md drop-in (GraphVecEnv's
    ``build_spaces``: aux5, scaled_float, unreal_wrapper) the Model receives the same arguments; ``set_hardness`` is the
    attribute ``create_envs`` installs (:68-70)."""
    import importlib
    from oracle import ref_harness as rh
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    vec_env = importlib.import_module("a2cat-vn-pytorch_b200.vec_env")
    g = H.load("trainer_contract")
    scene = H.trainer_contract_scene(g)
    ref = rh.ref_modules()
    import io, pickle  # noqa: E401
    X, Y = scene.maze.shape
    arrs = {}
    for plane, c in (("rgb", 3), ("depth", 1), ("segmentation", 3)):
        a = np.zeros((X, Y, 4, 174, 174, c), np.uint8)
        a[scene.cells[:, 0], scene.cells[:, 1]] = scene.plane_frames(plane).reshape(scene.n_cells, 4, 174, 174, c)
        arrs[plane] = a
    world = ref.thor_world.ThorGridWorld(scene.maze.copy(), arrs["rgb"], arrs["depth"], arrs["segmentation"])
    world.graph, world.optimal_actions = ref.util.compute_shortest_path_data(scene.maze)
    calls = []
    exp = rh.ref_experiment(lambda name: world, calls)
    trainer = exp.Trainer()
    goal = tuple(int(v) for v in g["goal"])
    trainer.env = trainer.create_env(dict(exp.default_args()["env_kwargs"], tasks=[("synthetic-174", [goal])] * 4))
    trainer.create_model()
    assert trainer.env.unwrapped_calls == [("set_complexity", (0.01,))]            # create_envs: env.set_hardness(0.01)
    ref_sp = trainer.env.observation_space
    # the drop-in's spaces, built without a device
    w = vn.compile_world([scene], vn.GYM_GRAPH, tasks=[(0, goal)])
    obs_space, act_space = vec_env.build_spaces(w.layout, "aux5", scaled_float=True, unreal_wrapper=True)
    trainer.env = type("Env", (), {"observation_space": obs_space, "action_space": act_space})()
    trainer.create_model()
    assert calls[0] == calls[1] == ((3, 4), {})
    assert calls[0][0] == tuple(g["model_args"].tolist())
    # same nesting, same first leaf (the one the model and _get_input_for_pixel_control use, :47-48,55), same lar Box
    assert len(obs_space.spaces) == len(ref_sp.spaces) == 2 and len(obs_space.spaces[0].spaces) == len(ref_sp.spaces[0].spaces) == 5
    assert tuple(obs_space.spaces[0].spaces[0].shape) == tuple(ref_sp.spaces[0].spaces[0].shape) == (3, 174, 174)
    assert tuple(obs_space.spaces[1].shape) == tuple(ref_sp.spaces[1].shape) == (5,)
    assert obs_space.spaces[0].spaces[0].dtype == ref_sp.spaces[0].spaces[0].dtype == np.float32
    # _get_input_for_pixel_control (:47-48) picks inputs[0][0]: the rgb leaf in both layouts
    assert vec_env.OBS_LAYOUTS["aux5"][0] == "rgb"
