"""Host-side compiler: scenes -> the flat tables the device kernels index.

Everything the reference recomputes per step / per reset in Python is hoisted here, once:

* ``adj [S, 4] int32`` - next-state table with -1 for a collision.  Replaces
  ``graph.util.step`` + ``is_valid_state`` (graph/util.py:15-25,36-37) and is the ``graph``
  dataset of the h5 schema (graph/util.py:212-218,229-233).
* per task ``(scene, goal)`` a list of candidate start states sorted by curriculum distance.
  Replaces the O(cells x 4) Python loop of ``sample_initial_state`` (graph/util.py:119-143) /
  ``sample_initial_position`` (graph/util.py:88-117) that the reference runs on EVERY reset;
  the curriculum becomes a prefix length of the sorted list.
* grid shortest paths by BFS - same tables as the recursive relaxation
  ``compute_shortest_path_data`` (graph/util.py:146-176).

Three rule families (``Family``) capture the semantic differences between the reference envs;
see DESIGN.md "Rule families".
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .scenes import GridScene

DIRS = np.array([(1, 0), (0, 1), (-1, 0), (0, -1)], np.int32)   # graph/util.py:4-13


# --------------------------------------------------------------------------- rule families
@dataclass(frozen=True)
class Family:
    """How one reference env class turns (state, action) into (state', reward, done)."""
    name: str
    oriented: bool
    action_order: str          # 'graph' = fwd,rot+,back,rot-   'h5' = fwd,back,rot+,rot-   'compass'
    goal_compare: int          # 0 = full state, 1 = position only (state>>2), 2 = never
    collision_skips_goal: bool  # graph envs return before the goal test on collision
    neg_step_reward: bool       # h5 env: reward = -cfg[1]
    collision_overrides: bool   # h5 env: collided reward overrides terminal reward
    term_prev_obs: bool         # h5 env returns the previous observation on terminal
    two_level_sampling: bool    # un-oriented curriculum: 0.9 / 0.1 buckets
    noop_action: bool           # SimpleGraphEnv: action None / -1 is a no-op with reward 0.0
    curriculum_offset: int      # optimal_distance = c * (max_dist + offset) + 1


#: environments/gym_graph/graph.py:9-93 (OrientedGraphEnv, GoalGymGraphAuxiliaryEnv)
GYM_GRAPH = Family("gym_graph", True, "graph", 0, True, False, False, False, False, False, 3)
#: graph/env.py:8-70 - goal test ``state[:2] == goal`` compares a 2-tuple with the 3-tuple goal
#: that reset() needs, so it is never true; reproduced literally.
GRAPH_ENV_ORIENTED = Family("graph_env_oriented", True, "graph", 2, True, False, False, False, False, False, 3)
#: the evident intent of graph/env.py:61 (position-only goal), offered explicitly
GRAPH_ENV_ORIENTED_POSITION = Family("graph_env_oriented_position", True, "graph", 1, True, False, False,
                                     False, False, False, 3)
#: graph/env.py:73-143,150-222 (SimpleGraphEnv / MultipleGraphEnv)
SIMPLE_GRAPH = Family("simple_graph", False, "compass", 0, True, False, False, False, True, True, -1)
#: environments/gym_ai2thor/envs/cached.py:10-103, environments/gym_thor_cached.py:7-95
THOR_CACHED = Family("thor_cached", True, "h5", 0, False, True, True, True, False, False, 0)

FAMILIES = {f.name: f for f in (GYM_GRAPH, GRAPH_ENV_ORIENTED, GRAPH_ENV_ORIENTED_POSITION, SIMPLE_GRAPH,
                                THOR_CACHED)}


# --------------------------------------------------------------------------- shortest paths
def bfs_distances(maze: np.ndarray, goal_xy) -> np.ndarray:
    """int32 [X, Y] grid distance of every cell TO ``goal_xy`` (4-connected), -1 for walls and
    unreachable cells - one column ``distances[:, :, gx, gy]`` of graph/util.py:146-176."""
    maze = np.asarray(maze).astype(bool)
    X, Y = maze.shape
    dist = np.full((X, Y), -1, np.int32)
    if not maze[goal_xy[0], goal_xy[1]]:
        return dist
    frontier = np.zeros((X, Y), bool)
    frontier[goal_xy[0], goal_xy[1]] = True
    d = 0
    while frontier.any():
        dist[frontier] = d
        nxt = np.zeros((X, Y), bool)
        nxt[1:, :] |= frontier[:-1, :]
        nxt[:-1, :] |= frontier[1:, :]
        nxt[:, 1:] |= frontier[:, :-1]
        nxt[:, :-1] |= frontier[:, 1:]
        frontier = nxt & maze & (dist < 0)
        d += 1
    return dist


def optimal_actions_from(dist: np.ndarray) -> np.ndarray:
    """bool [X, Y, 4]: ``out[x, y, d]`` is True iff moving in direction ``d`` from (x, y) lands on a
    cell one step closer to the goal - the ``actions[..., gx, gy, :]`` column of
    graph/util.py:146-176 (all optimal first moves; all False at the goal itself)."""
    X, Y = dist.shape
    out = np.zeros((X, Y, 4), bool)
    for d, (dx, dy) in enumerate(DIRS):
        nb = np.full((X, Y), -1, np.int32)
        xs = slice(max(0, -dx), X - max(0, dx))
        xd = slice(max(0, dx), X - max(0, -dx))
        ys = slice(max(0, -dy), Y - max(0, dy))
        yd = slice(max(0, dy), Y - max(0, -dy))
        nb[xs, ys] = dist[xd, yd]
        out[:, :, d] = (dist > 0) & (nb >= 0) & (nb == dist - 1)
    return out


def all_pairs(maze: np.ndarray):
    """(distances [X,Y,X,Y] int32, actions [X,Y,X,Y,4] bool) == compute_shortest_path_data."""
    maze = np.asarray(maze).astype(bool)
    X, Y = maze.shape
    distances = np.full((X, Y, X, Y), -1, np.int32)
    actions = np.zeros((X, Y, X, Y, 4), bool)
    for gx, gy in np.argwhere(maze):
        d = bfs_distances(maze, (gx, gy))
        distances[:, :, gx, gy] = d
        actions[:, :, gx, gy, :] = optimal_actions_from(d)
    return distances, actions


def rotation_steps(opt_actions_xy: np.ndarray, goal_r: int) -> np.ndarray:
    """int32 [X, Y, 4]: compute_rotation_steps (graph/util.py:82-86) for every state.
    ``min over optimal directions x of ((r - (goal_r + x)) % 4, with 3 -> 1)``; cells with no
    optimal direction get a large sentinel (the reference would raise on an empty min)."""
    X, Y, _ = opt_actions_xy.shape
    r = np.arange(4, dtype=np.int32)[None, None, :, None]
    x = np.arange(4, dtype=np.int32)[None, None, None, :]
    steps = (r - (goal_r + x)) % 4
    steps = np.where(steps == 3, 1, steps)
    steps = np.where(opt_actions_xy[:, :, None, :], steps, 1 << 20)
    return steps.min(-1).astype(np.int32)


# --------------------------------------------------------------------------- adjacency
def build_adjacency(scene: GridScene, action_order: str) -> np.ndarray:
    """int32 [S, 4] next-state table, -1 = collision (local state indices)."""
    rank = scene.cell_rank
    X, Y = rank.shape
    cells = scene.cells

    def neighbour(d):
        nx = cells[:, 0] + DIRS[d, 0]
        ny = cells[:, 1] + DIRS[d, 1]
        ok = (nx >= 0) & (ny >= 0) & (nx < X) & (ny < Y)
        out = np.full(len(cells), -1, np.int32)
        out[ok] = rank[nx[ok], ny[ok]]
        return out

    nb = np.stack([neighbour(d) for d in range(4)], 1)      # [cells, 4] by compass direction
    if action_order == "compass":                            # graph/env.py:122
        assert not scene.oriented
        return nb.astype(np.int32)
    assert scene.oriented
    C = len(cells)
    adj = np.empty((C, 4, 4), np.int32)
    for r in range(4):
        fwd = np.where(nb[:, r] >= 0, nb[:, r] * 4 + r, -1)
        back = np.where(nb[:, (r + 2) % 4] >= 0, nb[:, (r + 2) % 4] * 4 + r, -1)
        rot_p = np.arange(C) * 4 + (r + 1) % 4
        rot_m = np.arange(C) * 4 + (r + 3) % 4
        if action_order == "graph":      # graph/util.py:15-25: 0 fwd, 1 rot+, 2 back, 3 rot-
            adj[:, r, :] = np.stack([fwd, rot_p, back, rot_m], 1)
        elif action_order == "h5":       # graph/util.py:217-218,231-232: fwd, back, rot+, rot-
            adj[:, r, :] = np.stack([fwd, back, rot_p, rot_m], 1)
        else:
            raise ValueError(action_order)
    return adj.reshape(C * 4, 4)


# --------------------------------------------------------------------------- candidates
@dataclass
class Task:
    scene: int                 # index into the scene list
    goal: tuple                # (x, y[, r]) or flat index for THOR_CACHED
    goal_state: int            # local flat state index of the goal
    cand_state: np.ndarray     # int32 [K] local flat states sorted by cand_dist (stable)
    cand_dist: np.ndarray      # int32 [K]
    max_dist: int              # np.max(graph.graph) of the scene (largest_distance)


def scene_max_dist(scene: GridScene) -> int:
    """``np.max(self.graph.graph)`` (gym_graph/graph.py:35, graph/env.py:26): the largest grid
    distance over all pairs of free cells.  All-pairs BFS on the cell graph (scipy's C implementation
    when available, else one numpy BFS per cell)."""
    try:
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import shortest_path
    except Exception:
        m = 0
        for x, y in scene.cells:
            m = max(m, int(bfs_distances(scene.maze, (x, y)).max()))
        return m
    rank = scene.cell_rank
    X, Y = rank.shape
    rows, cols = [], []
    for dx, dy in ((1, 0), (0, 1)):
        a = rank[:X - dx, :Y - dy]
        b = rank[dx:, dy:]
        ok = (a >= 0) & (b >= 0)
        rows.append(a[ok])
        cols.append(b[ok])
    r, c = np.concatenate(rows), np.concatenate(cols)
    n = scene.n_cells
    g = csr_matrix((np.ones(len(r) * 2, np.int8), (np.concatenate([r, c]), np.concatenate([c, r]))), shape=(n, n))
    m = 0
    for lo in range(0, n, 512):      # bounded memory: 512 sources at a time
        d = shortest_path(g, method="D", unweighted=True, indices=np.arange(lo, min(n, lo + 512)))
        d = d[np.isfinite(d)]
        m = max(m, int(d.max()) if d.size else 0)
    return m


def build_task(scene: GridScene, scene_idx: int, goal, family: Family, max_dist: Optional[int] = None) -> Task:
    if max_dist is None:
        max_dist = 0 if family.name == "thor_cached" else scene_max_dist(scene)   # cached.py has no curriculum
    rank = scene.cell_rank
    if family.name == "thor_cached":
        # goal is a flat state index (cached.py:39); candidates = {s : dist[s][goal] > 0} with
        # dist = grid_dist + rot_diff, rot_diff 3 -> 1 (graph/util.py:240-247).  Literal, including
        # the quirk that an unreachable cell (-1) at rot_diff 2 has "distance" 1 > 0.
        g = int(goal)
        spd = getattr(scene, "h5_spd", None)
        if spd is not None:     # the file's own 'shortest_path_distance' dataset (loaders.scene_from_h5_arrays)
            col = np.asarray(spd)[:, g]
            states = np.nonzero(col > 0)[0].astype(np.int32)
            dist = col[states].astype(np.int32)
            order = np.argsort(dist, kind="stable")
            return Task(scene_idx, g, g, states[order], dist[order], max_dist)
        gx, gy = scene.cells[g >> 2]
        gr = g & 3
        d = bfs_distances(scene.maze, (gx, gy))[scene.cells[:, 0], scene.cells[:, 1]]   # [C]
        rd = np.abs(np.arange(4)[None, :] - gr)
        rd = np.where(rd == 3, 1, rd)
        full = d[:, None] + rd                                                          # [C, 4]
        states = np.nonzero(full.reshape(-1) > 0)[0].astype(np.int32)
        dist = full.reshape(-1)[states].astype(np.int32)
        order = np.argsort(dist, kind="stable")
        return Task(scene_idx, g, g, states[order], dist[order], max_dist)
    d_xy = bfs_distances(scene.maze, goal[:2])
    d = d_xy[scene.cells[:, 0], scene.cells[:, 1]]                                       # [C]
    if family.oriented:
        opt = optimal_actions_from(d_xy)
        rot = rotation_steps(opt, int(goal[2]))[scene.cells[:, 0], scene.cells[:, 1]]    # [C, 4]
        ok = d > 0                                                                       # util.py:123-124
        cd = d[:, None] + rot                                                            # util.py:128
        states = (np.nonzero(ok)[0][:, None] * 4 + np.arange(4)[None, :]).reshape(-1).astype(np.int32)
        dist = cd[ok].reshape(-1).astype(np.int32)
        goal_state = scene.state_index(goal)
    else:
        ok = d > 0                                                                       # util.py:93-95
        states = np.nonzero(ok)[0].astype(np.int32)
        dist = d[ok].astype(np.int32)
        goal_state = scene.state_index(goal)
    order = np.argsort(dist, kind="stable")
    return Task(scene_idx, tuple(goal), goal_state, states[order], dist[order], max_dist)


def curriculum_prefix(task: Task, family: Family, complexity: Optional[float]) -> int:
    """Number of leading candidates eligible under ``set_complexity(c)``.
    optimal_distance = c * (largest_distance + 4 - 1) + 1 (gym_graph/graph.py:49-51) or
    c * (largest_distance - 1) + 1 (graph/env.py:105,184); eligibility ``dist <= optimal_distance``
    (graph/util.py:135, :104).  ``None`` = no curriculum = all candidates."""
    if complexity is None or family.name == "thor_cached":
        return len(task.cand_state)
    optimal_distance = complexity * (task.max_dist + family.curriculum_offset) + 1
    return int(np.searchsorted(task.cand_dist, optimal_distance, side="right"))


# --------------------------------------------------------------------------- compiled world
STORE_ALIGN = 128   # plane offsets and the per-state pitch are multiples of one cache line


@dataclass
class StoreLayout:
    """``plane_bytes`` is what the kernels copy per plane: the frame (``frame_bytes`` = H * W * C) rounded up to
    a whole number of 16-byte units.  84 x 84 frames need no rounding; the reference's native 174 x 174 frames
    (GraphResize default, graph/core.py:43-49: 90,828 B rgb, 30,276 B depth) get 4 / 12 bytes of zero padding,
    in the store records and in the rows of the observation batch alike."""
    planes: Tuple[str, ...]
    plane_bytes: Tuple[int, ...]
    plane_off: Tuple[int, ...]
    state_pitch: int
    frame_hw: Tuple[int, int]
    frame_bytes: Tuple[int, ...] = ()
    channels: Tuple[int, ...] = ()

    @staticmethod
    def make(planes, frame_hw):
        from .scenes import PLANE_CHANNELS
        off, offs, sizes, exact, chans = 0, [], [], [], []
        for p in planes:
            nb = frame_hw[0] * frame_hw[1] * PLANE_CHANNELS[p]
            exact.append(nb)
            chans.append(PLANE_CHANNELS[p])
            nb = -(-nb // 16) * 16
            offs.append(off)
            sizes.append(nb)
            off += -(-nb // STORE_ALIGN) * STORE_ALIGN
        return StoreLayout(tuple(planes), tuple(sizes), tuple(offs), off, tuple(frame_hw), tuple(exact), tuple(chans))

    def batch(self, plane, n, device):
        """uint8 ``[n, H, W, C]`` observation batch of ``plane`` whose rows are ``plane_bytes`` apart (what the
        gather kernels write): contiguous when the frame is a whole number of 16-byte units, else a strided view
        of a padded allocation."""
        import torch
        i = self.planes.index(plane)
        h, w = self.frame_hw
        c, pitch = self.channels[i], self.plane_bytes[i]
        raw = torch.zeros((max(n, 1), pitch), dtype=torch.uint8, device=device)
        return raw.as_strided((n, h, w, c), (pitch, w * c, c, 1))


@dataclass
class World:
    """All scenes concatenated: global state index = scene_base[scene] + local state."""
    family: Family
    scenes: List[GridScene]
    scene_base: np.ndarray        # int64 [n_scenes + 1]
    adj: np.ndarray               # int32 [S_total, 4] GLOBAL indices, -1 collision
    tasks: List[Task]
    task_goal: np.ndarray         # int32 [T] global goal state
    task_cand_off: np.ndarray     # int32 [T + 1]
    cand_state: np.ndarray        # int32 [sum K] global candidate states
    layout: StoreLayout

    @property
    def n_states(self):
        return int(self.scene_base[-1])

    def prefixes(self, complexity) -> np.ndarray:
        return np.array([curriculum_prefix(t, self.family, complexity) for t in self.tasks], np.int32)

    def scene_of_state(self, gstate):
        return np.searchsorted(self.scene_base, gstate, side="right") - 1

    def state_tuple(self, gstate: int):
        s = int(self.scene_of_state(gstate))
        return self.scenes[s].state_tuple(int(gstate - self.scene_base[s]))


def compile_world(scenes: Sequence[GridScene], family: Family, tasks: Optional[Sequence[Tuple[int, tuple]]] = None,
                  planes: Optional[Sequence[str]] = None) -> World:
    """``tasks`` = [(scene_index, goal)], default: every goal of every scene (the reference builds one
    env per (scene, goal), experiments/thor_cached_auxiliary.py:66)."""
    scenes = list(scenes)
    for s in scenes:
        if s.oriented != family.oriented:
            raise ValueError("scene %s orientation does not match family %s" % (s.name, family.name))
        if s.frame_hw != scenes[0].frame_hw:
            raise ValueError("all scenes must share one frame size (resize is hoisted to store build)")
    planes = tuple(planes) if planes is not None else scenes[0].planes
    base = np.zeros(len(scenes) + 1, np.int64)
    adjs = []
    for i, s in enumerate(scenes):
        if getattr(s, "h5_graph", None) is not None and family.action_order == "h5":
            a = np.asarray(s.h5_graph, np.int64)      # the file's own 'graph' dataset (loaders.scene_from_h5_arrays)
        else:
            a = build_adjacency(s, family.action_order).astype(np.int64)
        a = np.where(a >= 0, a + base[i], -1)
        adjs.append(a.astype(np.int32))
        base[i + 1] = base[i] + s.n_states
    if tasks is None:
        tasks = [(i, g) for i, s in enumerate(scenes) for g in s.goals]
    maxd = {}
    built = []
    for si, goal in tasks:
        if si not in maxd:
            maxd[si] = 0 if family.name == "thor_cached" else scene_max_dist(scenes[si])
        built.append(build_task(scenes[si], si, goal, family, maxd[si]))
    off = np.zeros(len(built) + 1, np.int64)
    for i, t in enumerate(built):
        if len(t.cand_state) == 0:
            raise ValueError("task %d has no candidate start states" % i)
        off[i + 1] = off[i] + len(t.cand_state)
    if off[-1] >= 2 ** 31:
        raise ValueError("candidate table too large")
    cand = np.concatenate([t.cand_state.astype(np.int64) + base[t.scene] for t in built]).astype(np.int32)
    goal = np.array([t.goal_state + base[t.scene] for t in built], np.int32)
    return World(family, scenes, base, np.concatenate(adjs, 0), built, goal, off.astype(np.int32), cand,
                 StoreLayout.make(planes, scenes[0].frame_hw))


# --------------------------------------------------------------------------- evaluation services
def optimal_policy_table(world: World, task_index: int):
    """Shortest action sequences on the STATE graph of one task (the
    reference only keeps grid-level ``optimal_actions``, graph/util.py:146-176, which ignore rotations).

    Returns ``(dist, action)`` over the task's scene states (local indices): ``dist[s]`` = fewest actions
    from ``s`` to the goal state (-1 if unreachable), ``action[s]`` = the lowest-numbered action that starts
    such a sequence (-1 at the goal / unreachable).  Reverse BFS over ``adj`` from the goal."""
    from collections import deque
    t = world.tasks[task_index]
    base = int(world.scene_base[t.scene])
    n = world.scenes[t.scene].n_states
    adj = world.adj[base:base + n]
    local = np.where(adj >= 0, adj - base, -1)
    rev = [[] for _ in range(n)]
    for s in range(n):
        for a in range(4):
            d = int(local[s, a])
            if d >= 0 and d != s:
                rev[d].append(s)
    goal_compare = world.family.goal_compare
    goals = [t.goal_state] if goal_compare == 0 else \
        ([(t.goal_state >> 2) * 4 + r for r in range(4)] if goal_compare == 1 else [])
    dist = np.full(n, -1, np.int32)
    q = deque()
    for g in goals:
        dist[g] = 0
        q.append(g)
    while q:
        d = q.popleft()
        for s in rev[d]:
            if dist[s] < 0:
                dist[s] = dist[d] + 1
                q.append(s)
    action = np.full(n, -1, np.int32)
    for s in range(n):
        if dist[s] > 0:
            for a in range(4):
                d = int(local[s, a])
                if d >= 0 and dist[d] == dist[s] - 1:
                    action[s] = a
                    break
    return dist, action
