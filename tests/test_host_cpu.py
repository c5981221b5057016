"""CPU tests of the host-side logic: the flat-table compiler against the oracle (which follows the
reference's tuple arithmetic), the C-ABI library's exported symbols, sharding and the gloo
statistics reduction (world_size 2)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from oracle import graph_util as gu
from oracle import philox

vn = importlib.import_module("a2cat-vn-pytorch_b200")
T = vn.tables
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,shape", [(0, (10, 10)), (3, (9, 7)), (5, (6, 12))])
def test_adjacency_matches_reference_step(seed, shape):
    scene = H.scenes.make_maze_scene(shape, 0.25, seed, n_goals=1)
    adj = T.build_adjacency(scene, "graph")
    adj_h5 = T.build_adjacency(scene, "h5")
    for s in range(scene.n_states):
        st = scene.state_tuple(s)
        for a in range(4):
            n = gu.step(st, a)                                   # graph/util.py:15-25
            want = scene.state_index(n) if gu.is_valid_state(scene.maze, n) else -1
            assert adj[s, a] == want
        assert list(adj_h5[s]) == [adj[s, 0], adj[s, 2], adj[s, 1], adj[s, 3]]
    dist, _ = gu.compute_shortest_path_data(scene.maze)
    _, graph, _ = gu.h5_tables(scene.maze, dist)                 # graph/util.py:212-233
    assert np.array_equal(adj_h5, graph)
    un = H.scenes.GridScene(scene.maze, [tuple(scene.cells[0])], False, (84, 84), ("rgb",))
    adj_c = T.build_adjacency(un, "compass")
    for s in range(un.n_states):
        x, y = un.state_tuple(s)
        for a in range(4):
            c = gu.direction_to_change(a)
            n = (x + c[0], y + c[1])
            assert adj_c[s, a] == (un.state_index(n) if gu.is_valid_state(un.maze, n) else -1)


def test_bfs_tables_match_reference_golden():
    g = H.load("graph_util")
    for k in range(int(g["n_mazes"])):
        d, a = T.all_pairs(g["maze%d" % k])
        assert np.array_equal(d, g["dist%d" % k]) and np.array_equal(a, g["act%d" % k])


@pytest.mark.parametrize("family", ["gym_graph", "simple_graph"])
def test_candidate_tables_equal_reference_eligible_sets(family):
    """For every golden (maze, goal, complexity): the device's curriculum prefix == the set of
    candidates the reference gives positive weight (graph/util.py:135-142 / :103-116)."""
    g = H.load("graph_util")
    fam = T.FAMILIES[family]
    for k in range(int(g["n_mazes"])):
        maze = g["maze%d" % k]
        dist, act = gu.compute_shortest_path_data(maze)
        for gi, goal in enumerate(g["goals%d" % k]):
            goal = tuple(int(v) for v in goal)
            goal = goal if fam.oriented else goal[:2]
            scene = H.scenes.GridScene(maze, [goal], fam.oriented, (84, 84), ("rgb",))
            task = T.build_task(scene, 0, goal, fam)
            assert task.max_dist == int(dist.max())
            if fam.oriented:
                pots, d = gu.initial_state_candidates(maze, dist, act, goal)
            else:
                pots, d = gu.initial_position_candidates(maze, dist, goal)
            by_state = {scene.state_index(p): dd for p, dd in zip(pots, d)}
            assert sorted(by_state) == sorted(task.cand_state.tolist())
            assert [by_state[s] for s in task.cand_state.tolist()] == task.cand_dist.tolist()
            assert (np.diff(task.cand_dist) >= 0).all()
            for ci, c in enumerate(g["complexities"]):
                c = None if c < 0 else float(c)
                w = g["w_%s_%d_%d_%d" % ("state" if fam.oriented else "position", k, gi, ci)]
                pre = T.curriculum_prefix(task, fam, c)
                levels = np.unique(w[w > 0])
                hi = {scene.state_index(p) for p, ww in zip(pots, w) if ww == levels.max()}
                if c is None or len(levels) == 1:
                    want = {scene.state_index(p) for p, ww in zip(pots, w) if ww > 0}
                else:     # two-level rule: the prefix is the 0.9 bucket
                    mass = {lv: w[w == lv].sum() for lv in levels}
                    big = max(mass, key=mass.get)
                    want = {scene.state_index(p) for p, ww in zip(pots, w) if ww == big}
                assert set(task.cand_state[:pre].tolist()) == want, (k, gi, c)


def test_thor_cached_candidates_follow_h5_distance():
    scene = H.scenes.make_maze_scene((9, 8), 0.2, 31, n_goals=1, planes=("rgb",))
    dist, _ = gu.compute_shortest_path_data(scene.maze)
    _, _, spd = gu.h5_tables(scene.maze, dist)
    for goal in (0, 17, scene.n_states - 1):
        task = T.build_task(scene, 0, goal, T.THOR_CACHED)
        assert sorted(task.cand_state.tolist()) == np.nonzero(spd[:, goal] > 0)[0].tolist()   # cached.py:41-44


def test_world_concatenation_and_layout():
    scs = [H.scenes.make_maze_scene((8, 9), 0.2, 20 + k, n_goals=2, scene_id=k) for k in range(3)]
    w = T.compile_world(scs, T.GYM_GRAPH)
    assert w.n_states == sum(s.n_states for s in scs) and len(w.tasks) == 6
    for si, s in enumerate(scs):
        b = int(w.scene_base[si])
        local = T.build_adjacency(s, "graph")
        blk = w.adj[b:b + s.n_states]
        assert np.array_equal(np.where(blk >= 0, blk - b, -1), local)
        assert w.state_tuple(b + 5) == s.state_tuple(5)
    lay = w.layout
    assert lay.plane_bytes == (21168, 7056, 21168) and lay.state_pitch % 128 == 0
    assert all(o % 128 == 0 for o in lay.plane_off)
    # frames that are not a whole number of 16-byte units are padded (records and batch rows alike): the
    # reference's native 174 x 174 (graph/core.py:43-49)
    big = T.StoreLayout.make(("rgb", "depth", "segmentation"), (174, 174))
    assert big.frame_bytes == (90828, 30276, 90828) and big.plane_bytes == (90832, 30288, 90832)
    assert big.channels == (3, 1, 3) and all(o % 128 == 0 for o in big.plane_off) and big.state_pitch % 128 == 0
    small = T.StoreLayout.make(("rgb",), (10, 10))
    assert small.frame_bytes == (300,) and small.plane_bytes == (304,)
    assert H.scenes.frame_bytes(1, 2, np.arange(3), 0, 300).shape == (3, 300)
    assert np.array_equal(H.scenes.frame_bytes(1, 2, np.arange(3), 0, 300),
                          H.scenes.frame_bytes(1, 2, np.arange(3), 0, 304)[:, :300])


def test_frame_hash_is_stable():
    # known-answer bytes: guards the host hash that the device fill kernel must reproduce
    b = H.scenes.frame_bytes(0, 0, 0, 0, 16)
    assert b.dtype == np.uint8 and b.shape == (16,)
    b2 = H.scenes.frame_bytes(0, 0, np.array([0, 1]), 0, 16)
    assert np.array_equal(b2[0], b) and not np.array_equal(b2[1], b)
    assert H.crc(H.scenes.frame_bytes(7, 3, 11, 2, 21168)) == H.crc(H.scenes.frame_bytes(7, 3, 11, 2, 21168))


def test_generators():
    m = H.scenes.grown_scene_maze(1500, (50, 60), 0)
    assert m.sum() == 1500
    assert H.scenes._component_of(m, tuple(np.argwhere(m)[0])).sum() == 1500
    d = H.scenes.dungeon_maze((64, 64), 0)
    assert d.shape == (64, 64) and 0.1 < d.mean() < 0.7
    assert H.scenes._component_of(d, tuple(np.argwhere(d)[0])).sum() == d.sum()
    sc = H.scenes.make_dungeon_scene((64, 64), 0)
    assert sc.goals[0] == tuple(int(v) for v in np.argwhere(sc.maze)[0])      # dungeon_graph.py:20


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/vn_b200.h declares."""
    L = vn.lib
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "vn_b200.h")).read()
    declared = set(re.findall(r"^(?:int32_t|int64_t|const char \*)\s*\*?(vn_[a-z0-9_]+)\(", header, re.M))
    assert declared == set(L.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vn_abi_version() == vn.lib.ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", L.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vn_[a-z0-9_]+)", out))
    assert declared <= exported
    # struct sizes agree with the header (guards the ctypes mirrors)
    assert ctypes.sizeof(L.Store) == 8 + 8 + 4 + 4 + 24 + 24
    assert ctypes.sizeof(L.Rules) == 40 and ctypes.sizeof(L.Inject) == 24
    # 12 plane pointers, 12 output pointers + (parity, flags), sched / host_pack / host_seq + (seq, reserved), 5 rollout
    # record pointers, float_leaves pointer + 4 int32
    assert ctypes.sizeof(L.StepOut) == 8 * (6 + 6 + 15 + 2 + 5) + 8 + 16
    assert ctypes.sizeof(L.Replay) == 7 * 8 + 16 and ctypes.sizeof(L.FloatLeaf) == 24


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "a2cat-vn-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_shard_range_partitions():
    from importlib import import_module
    shard_range = import_module("a2cat-vn-pytorch_b200.vec_env").shard_range
    for n, w in ((262144, 8), (10, 3), (7, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ve = importlib.import_module("a2cat-vn-pytorch_b200.vec_env")
    lo, hi = ve.shard_range(101, rank, world)
    # each rank contributes statistics of its own env shard; the reduced vector is the job total
    local = np.array([hi - lo, 0.5 * (hi - lo), 3 * (hi - lo), rank, 0, 10 * (hi - lo), 0, 1], np.float64)
    total = ve.reduce_stats(local)
    # per-env Philox draws depend on the GLOBAL env id only
    draws = philox.reset_draws(9, np.arange(lo, hi), 0)
    q.put((rank, total.tolist(), lo, hi, draws[:, 0].tolist()))
    dist.destroy_process_group()


def test_multi_rank_stats_reduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, t0, lo0, hi0, d0), (r1, t1, lo1, hi1, d1) = res
    assert t0 == t1 and t0[0] == 101 and t0[5] == 1010 and t0[3] == 1 and t0[7] == 2
    assert (lo0, hi1) == (0, 101) and hi0 == lo1
    assert d0 + d1 == philox.reset_draws(9, np.arange(101), 0)[:, 0].tolist()


def test_loaders_thor_grid_world_and_h5():
    """Reference on-disk formats -> compiled world (loaders.py): dense ThorGridWorld arrays and the flat
    h5 schema reproduce the tables / frames of the scene they were exported from."""
    L = vn.loaders
    scene = H.scenes.make_maze_scene((9, 8), 0.2, 31, n_goals=2, planes=("rgb", "depth", "segmentation"))
    X, Y = scene.maze.shape

    class TGW:      # duck-typed graph/multi_graph_no_tp.py:6-11
        pass
    g = TGW()
    g._maze = scene.maze
    for attr, plane, c in (("_observations", "rgb", 3), ("_depths", "depth", 1), ("_segmentations", "segmentation", 3)):
        a = np.zeros((X, Y, 4, 84, 84, c), np.uint8)
        a[scene.cells[:, 0], scene.cells[:, 1]] = scene.plane_frames(plane).reshape(scene.n_cells, 4, 84, 84, c)
        setattr(g, attr, a)
    loaded = L.scene_from_thor_grid_world(g, scene.goals)
    for p in ("rgb", "depth", "segmentation"):
        assert np.array_equal(loaded.plane_frames(p), scene.plane_frames(p))
    w0, w1 = T.compile_world([scene], T.GYM_GRAPH), T.compile_world([loaded], T.GYM_GRAPH)
    assert np.array_equal(w0.adj, w1.adj) and np.array_equal(w0.cand_state, w1.cand_state)
    small = L.scene_from_thor_grid_world(g, scene.goals, screen_size=(44, 44), planes=("rgb",))
    assert small.plane_frames("rgb").shape == (scene.n_states, 44, 44, 3)       # resize hoisted to load time

    dist, _ = gu.compute_shortest_path_data(scene.maze)
    _, graph, spd = gu.h5_tables(scene.maze, dist)
    h5 = L.scene_from_h5_arrays(graph, scene.plane_frames("rgb"), spd)
    wa = T.compile_world([H.scenes.GridScene(scene.maze, [], True, (84, 84), ("rgb",), frame_seed=scene.frame_seed)],
                         T.THOR_CACHED, tasks=[(0, 5), (0, 40)])
    wb = T.compile_world([h5], T.THOR_CACHED, tasks=[(0, 5), (0, 40)])
    assert np.array_equal(wa.adj, wb.adj) and np.array_equal(wa.cand_state, wb.cand_state)
    assert np.array_equal(wa.task_cand_off, wb.task_cand_off)


def test_loader_third_person_planes():
    L = vn.loaders
    scene = H.scenes.make_maze_scene((6, 6), 0.2, 4, n_goals=1,
                                     planes=("rgb", "depth", "segmentation", "tp_rgb", "tp_depth", "tp_segmentation"))
    X, Y = scene.maze.shape

    class TGW:      # duck-typed graph/thor_graph.py:6-13
        pass
    g = TGW()
    g._maze = scene.maze
    for attr, plane, c in (("_observations", "rgb", 3), ("_depths", "depth", 1), ("_segmentations", "segmentation", 3),
                           ("_tp_observations", "tp_rgb", 3), ("_tp_depths", "tp_depth", 1),
                           ("_tp_segmentations", "tp_segmentation", 3)):
        a = np.zeros((X, Y, 4, 84, 84, c), np.uint8)
        a[scene.cells[:, 0], scene.cells[:, 1]] = scene.plane_frames(plane).reshape(scene.n_cells, 4, 84, 84, c)
        setattr(g, attr, a)
    loaded = L.scene_from_thor_grid_world(g, scene.goals, planes=scene.planes)
    for p in scene.planes:
        assert np.array_equal(loaded.plane_frames(p), scene.plane_frames(p))
    assert not np.array_equal(scene.plane_frames("rgb"), scene.plane_frames("tp_rgb"))


def test_scene_pickle_loads_without_the_reference_package(tmp_path):
    """loaders.load_scene_pickle: a ThorGridWorld pickle (class path graph.multi_graph_no_tp.ThorGridWorld) unpickles on
    a box that has no `graph` package, and compiles into the same world as the scene it was exported from."""
    import pickle
    scene = H.scenes.make_maze_scene((6, 7), 0.2, 5, n_goals=2, planes=("rgb", "depth", "segmentation"), frame_hw=(12, 12))
    path = tmp_path / "scene.pkl"
    H.write_reference_style_pickle(scene, path)
    with pytest.raises(ModuleNotFoundError):
        pickle.load(open(path, "rb"))                # the stock unpickler needs the reference package
    g = vn.loaders.load_scene_pickle(str(path))
    assert g.goals == list(scene.goals) and np.array_equal(g.maze, scene.maze)
    loaded = vn.loaders.scene_from_thor_grid_world(g, g.goals)
    w0, w1 = T.compile_world([scene], T.GYM_GRAPH), T.compile_world([loaded], T.GYM_GRAPH)
    assert np.array_equal(w0.adj, w1.adj) and np.array_equal(w0.cand_state, w1.cand_state)
    for p in ("rgb", "depth", "segmentation"):
        assert np.array_equal(loaded.plane_frames(p), scene.plane_frames(p))
    assert vn.loaders.THOR_CACHED_GOALS["thor-cached-225"][0] == (3, 17, 2)
    with pytest.raises(ValueError):
        vn.loaders.make_vec("NoSuchEnv-v0", [])


def test_h5_scene_file_reader_with_optional_h5py(monkeypatch, tmp_path):
    """loaders.load_scene_h5 reads the datasets THORDiscreteCachedEnv loads (cached.py:26-32) through h5py when it is
    installed; offline it is not, so a stand-in module serves a registry of arrays (the same one the reference harness
    uses) - and without any h5py the loader raises ImportError pointing at scene_from_h5_arrays."""
    import sys
    import types
    loaders = importlib.import_module("a2cat-vn-pytorch_b200.loaders")
    T = vn.tables
    rng = np.random.RandomState(3)
    scene = H.scenes.make_maze_scene((5, 6), 0.2, 2, n_goals=1, planes=("rgb",))
    w = T.compile_world([scene], T.GYM_GRAPH)
    graph = T.build_adjacency(scene, "h5")
    obs = scene.plane_frames("rgb")
    spd = rng.randint(0, 9, (scene.n_states, scene.n_states)).astype(np.int64)
    files = {"scene.h5": {"graph": graph, "observation": obs, "shortest_path_distance": spd}}

    class _DS:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, k):
            assert k == ()
            return self.a

    class _File:
        def __init__(self, path, mode="r"):
            self.d = files[os.path.basename(str(path))]

        def __getitem__(self, k):
            return _DS(self.d[k])

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    monkeypatch.setitem(sys.modules, "h5py", None)               # import h5py -> ImportError
    with pytest.raises(ImportError, match="scene_from_h5_arrays"):
        loaders.load_scene_h5(tmp_path / "scene.h5")
    fake = types.ModuleType("h5py")
    fake.File = _File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    sc = loaders.load_scene_h5(tmp_path / "scene.h5")
    ref = loaders.scene_from_h5_arrays(graph, obs, spd)
    assert np.array_equal(sc.h5_graph, ref.h5_graph) and np.array_equal(sc.h5_spd, spd)
    assert np.array_equal(sc.plane_frames("rgb"), obs) and sc.frame_hw == (84, 84)
    wc = T.compile_world([sc], T.THOR_CACHED, tasks=[(0, 5)])
    assert np.array_equal(wc.adj, graph) and w.n_states == wc.n_states
    small = loaders.load_scene_h5(tmp_path / "scene.h5", screen_size=(42, 42))
    assert small.frame_hw == (42, 42) and small.plane_frames("rgb").shape == (scene.n_states, 42, 42, 3)


def test_skimage_resize_restatement_properties():
    """skimage.transform.resize(anti_aliasing=True) (cached.py:62-64) restated - UNPINNED (scikit-image is absent).  What
    can be checked without it: same size is a plain / 255; constants stay constant; up-scaling (no pre-filter) is the
    half-pixel-centred bilinear interpolation cv2 implements too; an integer down-scale equals Gaussian(sigma =
    (s - 1) / 2, mirror) followed by sampling the filtered image between the two centre taps."""
    import cv2
    from scipy import ndimage
    loaders = importlib.import_module("a2cat-vn-pytorch_b200.loaders")
    rz = loaders.skimage_resize_restated
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (30, 40, 3)).astype(np.uint8)
    assert np.array_equal(rz(img, (30, 40)), img.astype(np.float64) / 255.0)
    assert np.abs(rz(np.full((20, 24, 3), 77, np.uint8), (7, 9)) - 77 / 255.0).max() < 1e-15
    up = rz(img, (75, 100))
    want = cv2.resize(img.astype(np.float64) / 255.0, (100, 75), interpolation=cv2.INTER_LINEAR)
    assert np.abs(up - want).max() < 1e-12
    down = rz(img, (15, 20))                                   # factor 2 on both axes: sigma 0.5
    f = ndimage.gaussian_filter(img.astype(np.float64) / 255.0, (0.5, 0.5, 0), mode="mirror")
    want = 0.25 * (f[0::2, 0::2] + f[1::2, 0::2] + f[0::2, 1::2] + f[1::2, 1::2])      # source coordinate 2 i + 0.5
    assert np.abs(down - want).max() < 1e-12
    q = loaders._resize_frames_skimage(img[None], (15, 20))
    assert q.dtype == np.uint8 and np.abs(q[0] / 255.0 - down).max() <= 0.5 / 255 + 1e-12
