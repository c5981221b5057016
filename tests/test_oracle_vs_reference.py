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
    names = sorted(f for f in os.listdir(H.GOLDEN) if f.endswith(".npz"))
    assert names == sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz")) and len(names) == 9
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
