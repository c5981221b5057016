"""Synthetic scenes (no AI2-THOR / MINOS data offline) and the counter-based frame hash.

A scene is what the reference keeps in a pickled ``ThorGridWorld`` / ``MazeGraph``:
a boolean ``maze [X, Y]`` of free cells (graph/multi_graph_no_tp.py:134-152) plus, for every
free cell and each of 4 rotations, the cached planes ``rgb [H,W,3]``, ``depth [H,W,1]``,
``segmentation [H,W,3]`` (uint8).  Here the planes are never stored on disk: byte ``k`` of plane
``p`` of state ``s`` in scene ``c`` is a pure function ``frame_hash(seed, c, s, p, k)`` so the host
(numpy, this file), the oracle and the device fill kernel (csrc/vn_kernels.cu: vn_fill_store)
all produce identical, incompressible bytes.

State numbering (shared with graph/util.py:208-210,229-237 ``save_graph_as_h5``):
free cells are ranked row-major (x outer, y inner, graph/util.py:27-31) and
``state = cell_rank * 4 + rotation`` for oriented scenes, ``state = cell_rank`` otherwise.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

#: first-person planes (graph/multi_graph_no_tp.py:6-25) and the third-person ones the two-agent scenes add
#: (graph/thor_graph.py:6-36: ``_tp_observations`` / ``_tp_depths`` / ``_tp_segmentations``)
PLANE_NAMES = ("rgb", "depth", "segmentation", "tp_rgb", "tp_depth", "tp_segmentation")
PLANE_CHANNELS = {"rgb": 3, "depth": 1, "segmentation": 3, "tp_rgb": 3, "tp_depth": 1, "tp_segmentation": 3}
PLANE_ID = {"rgb": 0, "depth": 1, "segmentation": 2, "tp_rgb": 3, "tp_depth": 4, "tp_segmentation": 5}

_M64 = (1 << 64) - 1


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def frame_key(seed: int, scene: int, state, plane: int):
    """64-bit key of one (scene, state, plane) frame; ``state`` may be an array."""
    with np.errstate(over="ignore"):
        k = splitmix64(np.uint64(seed & _M64) ^ (np.uint64(scene) * np.uint64(0xD1B54A32D192ED03)))
        k = splitmix64(k ^ (np.asarray(state, dtype=np.uint64) * np.uint64(0x8CB92BA72F3D8DD7)))
        k = splitmix64(k ^ np.uint64(plane + 1))
    return k


def frame_bytes(seed: int, scene: int, state, plane: int, nbytes: int) -> np.ndarray:
    """uint8 ``[len(state), nbytes]`` (or ``[nbytes]`` for a scalar state): 8 bytes per hash word,
    word ``w`` of the frame = splitmix64(key + w), little-endian (a frame that is not a whole number of
    words ends inside its last word)."""
    scalar = np.ndim(state) == 0
    key = np.atleast_1d(frame_key(seed, scene, state, plane))
    nwords = -(-nbytes // 8)
    words = np.arange(nwords, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(key[:, None] + words[None, :])
    out = np.ascontiguousarray(h.astype("<u8").view(np.uint8).reshape(len(key), nwords * 8)[:, :nbytes])
    return out[0] if scalar else out


@dataclass
class GridScene:
    """One synthetic scene.  ``goals`` are ``(x, y, r)`` tuples for oriented scenes
    (environments/gym_graph/download.py:21-29) or ``(x, y)`` for un-oriented ones
    (graph/maze_graph.py:9, graph/dungeon_graph.py:20)."""
    maze: np.ndarray                       # [X, Y] bool
    goals: List[tuple]
    oriented: bool = True
    frame_hw: Tuple[int, int] = (84, 84)
    planes: Tuple[str, ...] = ("rgb", "depth", "segmentation")
    frame_seed: int = 0
    scene_id: int = 0
    name: str = "synthetic"
    # optional explicit planes {name: uint8 [n_states, H, W, C]}; when None the hash is used
    explicit: Optional[dict] = None
    cells: np.ndarray = field(init=False)   # [n_cells, 2] int32, row-major rank -> (x, y)
    cell_rank: np.ndarray = field(init=False)  # [X, Y] int32, -1 on walls

    def __post_init__(self):
        self.maze = np.asarray(self.maze).astype(bool)
        xs, ys = np.nonzero(self.maze)       # np.nonzero is row-major: x outer, y inner
        self.cells = np.stack([xs, ys], 1).astype(np.int32)
        self.cell_rank = np.full(self.maze.shape, -1, np.int32)
        self.cell_rank[xs, ys] = np.arange(len(xs), dtype=np.int32)

    @property
    def n_cells(self):
        return len(self.cells)

    @property
    def n_states(self):
        return self.n_cells * (4 if self.oriented else 1)

    def plane_nbytes(self, plane):
        h, w = self.frame_hw
        return h * w * PLANE_CHANNELS[plane]

    def state_index(self, state) -> int:
        """(x, y[, r]) -> flat state index."""
        c = int(self.cell_rank[state[0], state[1]])
        if c < 0:
            raise ValueError("state %r is not a free cell" % (state,))
        return c * 4 + int(state[2]) if self.oriented else c

    def state_tuple(self, idx: int) -> tuple:
        if self.oriented:
            x, y = self.cells[idx >> 2]
            return (int(x), int(y), int(idx & 3))
        x, y = self.cells[idx]
        return (int(x), int(y))

    def plane_frames(self, plane: str, states=None) -> np.ndarray:
        """uint8 [n, H, W, C] frames of ``plane`` for the given flat states (default all)."""
        if states is None:
            states = np.arange(self.n_states)
        states = np.asarray(states)
        h, w = self.frame_hw
        c = PLANE_CHANNELS[plane]
        if self.explicit is not None:
            return self.explicit[plane][states]
        return frame_bytes(self.frame_seed, self.scene_id, states, PLANE_ID[plane], h * w * c).reshape(
            len(states), h, w, c)


# --------------------------------------------------------------------------- generators
def random_maze(shape=(10, 10), wall_prob=0.25, seed=0) -> np.ndarray:
    """C1: i.i.d. walls, restricted to the largest 4-connected component so that every free cell
    can reach every goal (ties broken by the component met first in row-major order)."""
    rng = np.random.RandomState(seed)
    maze = rng.rand(*shape) >= wall_prob
    best, seen = None, np.zeros(shape, bool)
    for x, y in np.argwhere(maze):
        if not seen[x, y]:
            comp = _component_of(maze, (int(x), int(y)))
            seen |= comp
            if best is None or comp.sum() > best.sum():
                best = comp
    return best


def _component_of(maze, start):
    X, Y = maze.shape
    seen = np.zeros_like(maze, dtype=bool)
    stack = [start]
    seen[start] = True
    while stack:
        x, y = stack.pop()
        for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
            nx, ny = x + dx, y + dy
            if 0 <= nx < X and 0 <= ny < Y and maze[nx, ny] and not seen[nx, ny]:
                seen[nx, ny] = True
                stack.append((nx, ny))
    return seen


def grown_scene_maze(n_cells=1500, shape=(50, 60), seed=0) -> np.ndarray:
    """C2/C4: a connected free region of exactly ``n_cells`` cells grown by a lazy random walk
    with restarts from visited cells (apartment-like blobs with corridors)."""
    rng = np.random.RandomState(seed)
    X, Y = shape
    assert n_cells <= X * Y
    maze = np.zeros(shape, bool)
    pos = (X // 2, Y // 2)
    maze[pos] = True
    visited = [pos]
    count = 1
    dirs = ((1, 0), (0, 1), (-1, 0), (0, -1))
    while count < n_cells:
        if rng.rand() < 0.02:
            pos = visited[rng.randint(len(visited))]
        dx, dy = dirs[rng.randint(4)]
        nx, ny = pos[0] + dx, pos[1] + dy
        if 0 <= nx < X and 0 <= ny < Y:
            pos = (nx, ny)
            if not maze[pos]:
                maze[pos] = True
                visited.append(pos)
                count += 1
    return maze


def dungeon_maze(shape=(64, 64), seed=0, max_rooms=24, room_min=4, room_max=12) -> np.ndarray:
    """C3: rooms + L-shaped corridors, floor = True (contract of graph/dungeon_graph.py:7-15:
    a 0/1 tile array [H, W] with floor = 1; the generator module ``environment.util.dungeon``
    is missing from the reference tree, so this is a re-specification, not a restatement)."""
    rng = np.random.RandomState(seed)
    X, Y = shape
    maze = np.zeros(shape, bool)
    rooms = []
    for _ in range(max_rooms * 4):
        if len(rooms) >= max_rooms:
            break
        w = rng.randint(room_min, room_max + 1)
        h = rng.randint(room_min, room_max + 1)
        x = rng.randint(1, X - w - 1)
        y = rng.randint(1, Y - h - 1)
        if any(x < rx + rw + 1 and rx < x + w + 1 and y < ry + rh + 1 and ry < y + h + 1 for rx, ry, rw, rh in rooms):
            continue
        maze[x:x + w, y:y + h] = True
        if rooms:
            px, py, pw, ph = rooms[-1]
            cx, cy = x + w // 2, y + h // 2
            qx, qy = px + pw // 2, py + ph // 2
            if rng.rand() < 0.5:
                maze[min(cx, qx):max(cx, qx) + 1, cy] = True
                maze[qx, min(cy, qy):max(cy, qy) + 1] = True
            else:
                maze[cx, min(cy, qy):max(cy, qy) + 1] = True
                maze[min(cx, qx):max(cx, qx) + 1, qy] = True
        rooms.append((x, y, w, h))
    first = tuple(np.argwhere(maze)[0])
    return _component_of(maze, first)


def pick_goals(maze, n_goals, oriented=True, seed=0) -> List[tuple]:
    rng = np.random.RandomState(seed + 7919)
    cells = np.argwhere(maze)
    sel = rng.choice(len(cells), size=n_goals, replace=False)
    goals = []
    for i in sel:
        x, y = cells[i]
        goals.append((int(x), int(y), int(rng.randint(4))) if oriented else (int(x), int(y)))
    return goals


def make_maze_scene(shape=(10, 10), wall_prob=0.25, seed=0, n_goals=1, oriented=True, frame_hw=(84, 84),
                    planes=("rgb", "depth", "segmentation"), scene_id=0) -> GridScene:
    maze = random_maze(shape, wall_prob, seed)
    return GridScene(maze, pick_goals(maze, n_goals, oriented, seed), oriented, frame_hw, tuple(planes),
                     frame_seed=seed, scene_id=scene_id, name="maze%dx%d-s%d" % (shape[0], shape[1], seed))


def make_thor_scene(n_cells=1500, shape=(50, 60), seed=0, n_goals=4, frame_hw=(84, 84),
                    planes=("rgb", "depth"), scene_id=0) -> GridScene:
    maze = grown_scene_maze(n_cells, shape, seed)
    return GridScene(maze, pick_goals(maze, n_goals, True, seed), True, frame_hw, tuple(planes),
                     frame_seed=seed, scene_id=scene_id, name="thor-synth-%d-s%d" % (n_cells, seed))


def make_dungeon_scene(shape=(64, 64), seed=0, oriented=False, frame_hw=(84, 84), planes=("rgb",),
                       scene_id=0) -> GridScene:
    maze = dungeon_maze(shape, seed)
    first = tuple(int(v) for v in np.argwhere(maze)[0])   # goal = first free cell, dungeon_graph.py:20
    goal = first + (0,) if oriented else first
    return GridScene(maze, [goal], oriented, frame_hw, tuple(planes), frame_seed=seed, scene_id=scene_id,
                     name="dungeon%dx%d-s%d" % (shape[0], shape[1], seed))


def render_maze_frames(scene: GridScene, goal_xy, screen_hw=(84, 84)) -> np.ndarray:
    """MazeGraph.render (graph/maze_graph.py:20-24) for every free cell, then GraphResize
    (graph/core.py:30-41, cv2 bilinear) and x255 -> uint8: the store-build-time hoist of the
    per-step procedural render.  Returns uint8 [n_cells, H, W, 3]."""
    import cv2
    base = np.tile(np.expand_dims(scene.maze, 2), [1, 1, 3]).astype(np.float32)
    out = np.zeros((scene.n_cells,) + tuple(screen_hw) + (3,), np.uint8)
    for i, (x, y) in enumerate(scene.cells):
        r = base.copy()
        r[x, y] = (1.0, 0.0, 0.0)
        r[goal_xy[0], goal_xy[1]] = (0.0, 1.0, 0.0)
        if r.shape[:2] != tuple(screen_hw):
            r = cv2.resize(r, tuple(screen_hw))
        out[i] = np.clip(np.rint(r * 255.0), 0, 255).astype(np.uint8)
    return out
