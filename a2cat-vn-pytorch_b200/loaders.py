"""On-disk / in-memory scene formats of the reference -> GridScene with explicit planes.

* pickled ``ThorGridWorld`` (graph/multi_graph_no_tp.py:6-25; attribute names as printed by
  testPickles.py:23-37): ``_maze [X,Y]``, ``_observations/_segmentations [X,Y,4,H,W,3]``,
  ``_depths [X,Y,4,H,W,1]`` dense over the bounding box.
* the flat h5 schema written by ``save_graph_as_h5`` (graph/util.py:222-227): ``graph [S,4]``,
  ``observation [S,H,W,C]``, ``shortest_path_distance [S,S]``.

The per-step ``cv2.resize`` of ``GraphResize`` (graph/core.py:30-57) is applied ONCE here.
"""
import numpy as np

from .scenes import GridScene


def _resize_frames(frames, hw):
    """[n, H, W, C] uint8 -> [n, h, w, C] with cv2.resize (bilinear, the GraphResize default)."""
    if tuple(frames.shape[1:3]) == tuple(hw):
        return np.ascontiguousarray(frames)
    import cv2
    out = np.empty((frames.shape[0], hw[0], hw[1], frames.shape[3]), frames.dtype)
    for i, f in enumerate(frames):
        r = cv2.resize(f, (hw[1], hw[0]))            # cv2 takes (width, height); the reference passes a square size
        out[i] = r if r.ndim == 3 else r[:, :, None]  # graph/core.py:38-39 re-adds the channel axis
    return out


def scene_from_thor_grid_world(graph, goals, screen_size=None, planes=("rgb", "depth", "segmentation"), scene_id=0,
                               name="thor"):
    """``graph``: any object with the ThorGridWorld attributes.  Frames are compacted to free cells
    (state = free_cell_rank * 4 + rotation) and resized once to ``screen_size`` (default: stored size).
    Pass ``planes`` including ``tp_rgb`` / ``tp_depth`` / ``tp_segmentation`` for the third-person variant
    (graph/thor_graph.py), whose ``render`` returns the 6-tuple of ``obs_layout="thor6"``."""
    maze = np.asarray(graph._maze).astype(bool)
    xs, ys = np.nonzero(maze)
    src = {"rgb": graph._observations, "depth": graph._depths, "segmentation": graph._segmentations}
    for name, attr in (("tp_rgb", "_tp_observations"), ("tp_depth", "_tp_depths"), ("tp_segmentation", "_tp_segmentations")):
        if hasattr(graph, attr):                    # graph/thor_graph.py:6-13 (third-person camera of the 2nd agent)
            src[name] = getattr(graph, attr)
    hw = tuple(screen_size) if screen_size is not None else tuple(src["rgb"].shape[3:5])
    explicit = {}
    for p in planes:
        a = np.asarray(src[p])[xs, ys]                              # [cells, 4, H, W, C]
        a = a.reshape((-1,) + a.shape[2:])                          # [cells * 4, H, W, C]
        explicit[p] = _resize_frames(a, hw)
    return GridScene(maze, [tuple(g) for g in goals], True, hw, tuple(planes), scene_id=scene_id, name=name,
                     explicit=explicit)


def scene_from_h5_arrays(graph, observation, shortest_path_distance=None, location=None, goals=(), scene_id=0,
                         name="thor-h5"):
    """Flat h5 schema -> GridScene.  The grid geometry is recovered from the adjacency: states are
    grouped in fours (one free cell each, graph/util.py:229-237) and laid out on a single row of a
    [1, cells] maze when no ``location`` is given - the env only ever uses the flat ``graph`` table."""
    graph = np.asarray(graph)
    S = graph.shape[0]
    assert S % 4 == 0
    cells = S // 4
    obs = np.ascontiguousarray(observation)
    sc = GridScene(np.ones((1, cells), bool), list(goals), True, tuple(obs.shape[1:3]), ("rgb",), scene_id=scene_id,
                   name=name, explicit={"rgb": obs})
    sc.h5_graph = graph.astype(np.int32)         # used verbatim by tables.compile_world for THOR_CACHED
    # start-state candidates of a goal g are {s : shortest_path_distance[s][g] > 0} (cached.py:41-44)
    sc.h5_spd = None if shortest_path_distance is None else np.asarray(shortest_path_distance)
    return sc
