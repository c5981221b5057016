"""ORACLE (test infrastructure, never imported by the product package).

Scalar, one-object-per-env CPU restatement of the reference's environment classes on the hot
path.  Each class follows the reference line by line on Python tuples (citations per method) and
is pinned to the UNMODIFIED reference classes by the golden trajectories in tests/golden/
(make_golden.py drives both with the same injected action / reset streams) and by the live
comparison in tests/test_oracle_vs_reference.py when /root/reference is present.

Randomness: the reference draws from global Mersenne-Twister streams; here every ``reset`` takes
its draw from ``self.reset_source`` - a callable returning ``(goal_choice, start_state)`` - so the
same injected stream can drive the reference, this oracle and the device path.
"""
import numpy as np

from . import graph_util as gu


class OracleScene:
    """What the reference's pickled scene objects expose: ``maze``, ``graph`` (all-pairs distances),
    ``optimal_actions``, ``render``.  Built from a synthetic GridScene (input data only)."""

    def __init__(self, grid_scene, with_all_pairs=True, cache_frames=False):
        self.src = grid_scene
        self.maze = grid_scene.maze
        self.goals = list(grid_scene.goals)
        if with_all_pairs:
            self.graph, self.optimal_actions = gu.compute_shortest_path_data(self.maze)
        self.dtype = np.uint8
        h, w = grid_scene.frame_hw
        self.observation_shape = (h, w, 3)
        # the reference keeps every frame in RAM (multi_graph_no_tp.py:141-144); the CPU baseline
        # does the same so that render() is an index, not a hash evaluation
        self._cache = {p: grid_scene.plane_frames(p) for p in grid_scene.planes} if cache_frames else None

    def _frame(self, plane, s):
        if self._cache is not None:
            return self._cache[plane][s]
        return self.src.plane_frames(plane, [s])[0]

    # graph/multi_graph_no_tp.py:12-25 ThorGridWorld.render
    def render(self, position, direction, modes=("rgb",)):
        s = self.src.state_index(tuple(position) + (direction,))
        ret = tuple()
        for m in ("rgb", "depth", "segmentation"):
            if m in modes:
                ret = ret + (self._frame(m, s),)
        return ret[0] if len(ret) == 1 else ret

    # un-oriented scenes (graph/maze_graph.py:20-24 after the store-build-time hoist): frame by cell
    def render_cell(self, position):
        return self._frame("rgb", self.src.state_index(tuple(position)))


class GymGraphEnv:
    """environments/gym_graph/graph.py:9-93 OrientedGraphEnv."""

    def __init__(self, scene: OracleScene, goals=None, rewards=(1.0, 0.0, 0.0)):
        self.graph = scene
        self.goals = scene.goals if goals is None else goals           # :21-24
        from .vec import _Box
        self.observation_space = _Box(0, 255, scene.observation_shape, np.uint8)    # :26-31 (uint8 scenes)
        self.action_space = type("Discrete", (), {"n": 4})()                        # :33
        self.state = None
        self.largest_distance = int(np.max(scene.graph))               # :35
        self.complexity = None
        self.rewards = list(rewards)
        self.reset_source = None

    def set_complexity(self, complexity=None):                          # :43-44
        self.complexity = complexity

    def optimal_distance(self):                                         # :49-51
        if self.complexity is None:
            return None
        return self.complexity * (self.largest_distance + 4 - 1) + 1

    def reset(self):                                                    # :46-54
        choice, start = self.reset_source()
        self.goal = self.goals[choice % len(self.goals)] if isinstance(self.goals, list) else self.goals
        self.state = tuple(start)
        return self.observe(self.state)

    def observe(self, state):                                           # :56-58
        return self.graph.render(state[:2], state[2])

    def is_goal(self, state):                                           # :60-61
        return max(abs(a - b) for a, b in zip(state[:2], self.goal[:2])) == 0 and state[2] == self.goal[2]

    def step(self, action):                                             # :67-79
        nstate = gu.step(self.state, action)
        if not gu.is_valid_state(self.graph.maze, nstate):
            return self.observe(self.state), self.rewards[2], False, dict(state=self.state)
        self.state = nstate
        if self.is_goal(self.state):
            return self.observe(self.state), self.rewards[0], True, dict(state=self.state, win=True)
        return self.observe(self.state), self.rewards[1], False, dict(state=self.state)


class GymGraphAuxiliaryEnv(GymGraphEnv):
    """environments/gym_graph/graph.py:96-120 GoalGymGraphAuxiliaryEnv: 5-tuple observation
    ``(rgb, goal_rgb, depth, segmentation, goal_segmentation)``, goal planes memoised per goal."""

    def __init__(self, *a, screen_size=None, **k):
        super().__init__(*a, **k)
        from .vec import _Box, _Tuple
        # :98-104 - the declared Boxes of leaves 1.. use the env's own `screen_size` argument, which it does NOT forward
        # to the scene's GraphResize: the frames keep the scene's size whatever is declared here
        ss = tuple(screen_size) if screen_size is not None else tuple(self.graph.observation_shape[:2])
        self.observation_space = _Tuple((self.observation_space, _Box(0, 255, ss + (3,), np.uint8),
                                         _Box(0, 255, ss + (1,), np.uint8), _Box(0, 255, ss + (3,), np.uint8),
                                         _Box(0, 255, ss + (3,), np.uint8)))
        self._cached_goal = (None, None)

    def render_goal(self):                                              # :110-115
        cached, value = self._cached_goal
        if cached is None or cached != self.goal:
            value = self.graph.render(self.goal[:2], self.goal[2], modes=["rgb", "segmentation"])
            self._cached_goal = (self.goal, value)
        return value

    def observe(self, state):                                           # :117-120
        goal_rgb, goal_seg = self.render_goal()
        rgb, depth, seg = self.graph.render(state[:2], state[2], modes=["rgb", "depth", "segmentation"])
        return (rgb, goal_rgb, depth, seg, goal_seg)


class GymGraphRgbdGoalEnv(GymGraphAuxiliaryEnv):
    """BASELINE.json configs[1] observation "RGB + depth + goal": GoalGymGraphAuxiliaryEnv restricted to
    the leaves (rgb, goal_rgb, depth) - same render calls with modes without 'segmentation'."""

    def render_goal(self):
        cached, value = self._cached_goal
        if cached is None or cached != self.goal:
            value = self.graph.render(self.goal[:2], self.goal[2], modes=["rgb"])
            self._cached_goal = (self.goal, value)
        return value

    def observe(self, state):
        goal_rgb = self.render_goal()
        rgb, depth = self.graph.render(state[:2], state[2], modes=["rgb", "depth"])
        return (rgb, goal_rgb, depth)


class GraphEnvOriented(GymGraphEnv):
    """graph/env.py:8-70: float32 observation (/255 for uint8 scenes, :45-50) and the goal test
    ``self.state[:2] == self.goal`` (:61).  With the 3-tuple goal that reset() requires this
    compares a 2-tuple with a 3-tuple and is never true; restated literally."""

    def __init__(self, scene, goal, rewards=(1.0, 0.0, 0.0)):
        super().__init__(scene, goals=goal, rewards=rewards)
        self.goal = goal

    def reset(self):                                                    # :36-43
        _, start = self.reset_source()
        self.state = tuple(start)
        return self.observe(self.state)

    def observe(self, state):                                           # :45-50
        return self.graph.render(state[:2], state[2]).astype(np.float32) / 255.0

    def step(self, action):                                             # :52-65
        nstate = gu.step(self.state, action)
        if not gu.is_valid_state(self.graph.maze, nstate):
            return self.observe(self.state), self.rewards[2], False, dict(state=self.state)
        self.state = nstate
        if self.state[:2] == self.goal:
            return self.observe(self.state), self.rewards[0], True, dict(state=self.state, win=True)
        return self.observe(self.state), self.rewards[1], False, dict(state=self.state)


class SimpleGraphEnv:
    """graph/env.py:73-143 SimpleGraphEnv (un-oriented; 4 compass actions; -1 / None = no-op)."""

    def __init__(self, scene: OracleScene, goal=None, rewards=(1.0, 0.0, 0.0)):
        self.graph = scene
        self.goal = tuple(scene.goals[0]) if goal is None else tuple(goal)   # :81 graph.goal
        self.state = None
        self.largest_distance = int(np.max(scene.graph))               # :91
        self.complexity = None
        self._rewards = list(rewards)
        self.reset_source = None

    def set_complexity(self, complexity=None):
        self.complexity = complexity

    def optimal_distance(self):                                         # :103-105
        if self.complexity is None:
            return None
        return self.complexity * (self.largest_distance - 1) + 1

    def reset(self):                                                    # :102-108
        _, start = self.reset_source()
        self.state = tuple(start)
        return self.observe(self.state)

    def observe(self, state):                                           # :110-115
        return self.graph.render_cell(state).astype(np.float32) / 255.0

    def step(self, action):                                             # :117-133
        if action is None or action == -1:
            return self.observe(self.state), 0.0, False, dict()
        c = gu.direction_to_change(action)
        nstate = (self.state[0] + c[0], self.state[1] + c[1])
        if not gu.is_valid_state(self.graph.maze, nstate):
            return self.observe(self.state), self._rewards[2], False, dict(state=self.state)
        self.state = nstate
        if self.state[:2] == self.goal:
            return self.observe(self.state), self._rewards[0], True, dict(state=self.state, win=True)
        return self.observe(self.state), self._rewards[1], False, dict(state=self.state)


class MultipleGraphEnv:
    """graph/env.py:150-222: on reset pick ``graph_number = random.randrange(len(graphs))`` (:181),
    then behave like SimpleGraphEnv on that graph with its own goal (:185,209)."""

    def __init__(self, scenes, rewards=(1.0, 0.0, 0.0)):
        self.graphs = list(scenes)
        self.largest_distances = [int(np.max(s.graph)) for s in self.graphs]    # :166
        self.graph_number = None
        self.complexity = None
        self._rewards = list(rewards)
        self.state = None
        self.reset_source = None

    def set_complexity(self, complexity=None):
        self.complexity = complexity

    def optimal_distance(self, graph_number):                           # :182-184
        if self.complexity is None:
            return None
        return self.complexity * (self.largest_distances[graph_number] - 1) + 1

    @property
    def goal(self):
        return tuple(self.graphs[self.graph_number].goals[0])

    def reset(self):                                                    # :180-188
        choice, start = self.reset_source()
        self.graph_number = choice % len(self.graphs)
        self.state = tuple(start)
        return self.observe(self.state)

    def observe(self, state):                                           # :190-195
        return self.graphs[self.graph_number].render_cell(state).astype(np.float32) / 255.0

    def step(self, action):                                             # :197-213
        if action is None or action == -1:
            return self.observe(self.state), 0.0, False, dict()
        g = self.graphs[self.graph_number]
        c = gu.direction_to_change(action)
        nstate = (self.state[0] + c[0], self.state[1] + c[1])
        if not gu.is_valid_state(g.maze, nstate):
            return self.observe(self.state), self._rewards[2], False, dict(state=self.state)
        self.state = nstate
        if self.state[:2] == self.goal:
            return self.observe(self.state), self._rewards[0], True, dict(state=self.state, win=True)
        return self.observe(self.state), self._rewards[1], False, dict(state=self.state)


class ThorCachedEnv:
    """environments/gym_ai2thor/envs/cached.py:10-103 THORDiscreteCachedEnv on the flat h5 schema
    (graph [S,4], observation [S,H,W,3], shortest_path_distance [S,S]); with ``tasks`` and
    ``dict_obs`` it is the intended behaviour of the unfinished multi-scene rewrite
    environments/gym_thor_cached.py:7-95 (task list :47, raw uint8 pair :52-53, dict obs :89-92).
    Observations are returned as raw uint8 frames: the skimage resize of cached.py:62-64 is a
    float64 /255 of the same bytes when sizes match and is hoisted to store-build time otherwise."""
    reward_configuration = (1.0, 0.0, 0.0)                              # cached.py:70-72

    def __init__(self, graph, observations, shortest_path_distance, tasks=None, dict_obs=False):
        self._transition_graph = graph
        self._observations = observations
        self._shortest_path_distances = shortest_path_distance
        self._n_locations = graph.shape[0]                              # :30
        self.tasks = tasks
        self.dict_obs = dict_obs
        self.reset_source = None
        self.last_state = None

    def _pack(self, obs, goal):
        return {"image": obs, "goal": goal} if self.dict_obs else (obs, goal)

    def reset(self):                                                    # :47-57, :38-45
        goal_choice, start = self.reset_source()
        if self.tasks is None:
            self._current_goal_idx = goal_choice % self._n_locations    # randrange(n_locations)
        else:
            self._current_goal_idx = self.tasks[goal_choice % len(self.tasks)]
        assert self._shortest_path_distances[start][self._current_goal_idx] > 0   # rejection loop :41-44
        self._current_state_idx = int(start)
        self.last_state = self._pack(self._observations[self._current_state_idx],
                                     self._observations[self._current_goal_idx])
        return self.last_state

    def step(self, action):                                             # :74-99
        collided = False
        if self._transition_graph[self._current_state_idx][action] != -1:
            self._current_state_idx = int(self._transition_graph[self._current_state_idx][action])
        else:
            collided = True
        obs = self._observations[self._current_state_idx]
        goal = self._observations[self._current_goal_idx]
        terminal = self._current_goal_idx == self._current_state_idx
        reward = -self.reward_configuration[1]
        if terminal:
            reward = self.reward_configuration[0]
        if collided:
            reward = self.reward_configuration[2]
        state = self._pack(obs, goal) if not terminal else self.last_state
        self.last_state = state
        return state, reward, terminal, dict()


class ThorCachedTasksEnv:
    """environments/gym_thor_cached.py:7-95 THORCachedEnv - the UNFINISHED multi-scene rewrite of cached.py - restated
    as written: ``tasks = [(scene_name, goal_state)]``; ``reset`` (:45-50) draws a task, loads its scene and samples a
    start by rejection on ``shortest_path_distance[start][goal] > 0`` (:37-43); ``observe`` (:52-53) returns the RAW uint8
    ``(obs, goal)`` pair; ``process`` (:72-95) is cached.py's step with a ``{'image', 'goal'}`` dict of float32 / 255
    frames and the previous dict on a terminal step.  The class never sets the attributes ``process`` reads; they are
    kept in step with ``state`` / ``goal`` here exactly as oracle/ref_harness.A8Driver does for the reference, which is
    how tests/golden/thor_cached_tasks.npz was recorded.  ``scenes``: name -> dict(transition_graph, observations,
    shortest_path_distances)."""
    reward_configuration = (1.0, 0.0, 0.0)                                           # :68-70

    def __init__(self, scenes, tasks):
        self.scenes, self.tasks = scenes, tasks
        self.reset_source = None
        self.last_state = None

    @staticmethod
    def _preprocess_frame(image):                                                    # :57-61 (same-size resize)
        return image.astype(np.float32) / 255.0

    def reset(self):                                                                 # :45-50
        choice, start = self.reset_source()
        scene_id, self.goal = self.tasks[choice % len(self.tasks)]
        self.current_scene = self.scenes[scene_id]
        assert self.current_scene["shortest_path_distances"][start][self.goal] > 0  # _sample_start :37-43
        self.state = int(start)
        ob = self.observe()
        self.last_state = {"image": self._preprocess_frame(ob[0]), "goal": self._preprocess_frame(ob[1])}
        return ob

    def observe(self):                                                               # :52-53
        o = self.current_scene["observations"]
        return o[self.state], o[self.goal]

    def step(self, action):                                                          # process, :72-95
        graph = self.current_scene["transition_graph"]
        collided = False
        if graph[self.state][action] != -1:
            self.state = int(graph[self.state][action])
        else:
            collided = True
        o = self.current_scene["observations"]
        obs, goal = o[self.state], o[self.goal]
        terminal = self.goal == self.state
        reward = -self.reward_configuration[1]
        if terminal:
            reward = self.reward_configuration[0]
        if collided:
            reward = self.reward_configuration[2]
        if not terminal:
            state = {"image": self._preprocess_frame(obs), "goal": self._preprocess_frame(goal)}
        else:
            state = self.last_state
        self.last_state = state
        return state, reward, terminal, dict()
