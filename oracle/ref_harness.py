"""Harness that imports the UNMODIFIED reference env classes from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It is used
 (a) by ``tests/golden/make_golden.py`` to generate the committed golden vectors, and
 (b) by the CPU tests marked ``needs_reference`` that re-check the oracle restatement
     against the live reference when ``/root/reference`` is present (this container only;
     the GPU box has no reference tree, so those tests skip there).

The reference needs ``gym`` (0.15.7), ``h5py`` and ``skimage`` which are not installed
offline.  We put minimal in-memory stand-ins into ``sys.modules``:
  * ``gym``: ``Env``, ``spaces.{Box,Discrete,Tuple}``, ``register``/``make`` - just enough for
    the class bodies to execute (reference use: environments/gym_graph/graph.py:1-37,
    graph/env.py:1-30, environments/gym_ai2thor/envs/cached.py:1-36).
  * ``h5py.File``: reads from an in-memory dict registry (cached.py:26-32 does
    ``file['observation'][()]``).
  * ``skimage.transform.resize``: identity on same-size frames returning float64/255
    (what skimage does for a same-size uint8 image, cached.py:62-64).
The package ``environments`` is registered as a *namespace stub* so that importing
``environments.gym_graph.graph`` does not execute ``environments/__init__.py`` (which pulls
in the live simulators).
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("VN_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graph"))


# --------------------------------------------------------------------------- gym stand-in
def _make_gym():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:
        metadata = {}

        @property
        def unwrapped(self):
            return self

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    class Discrete(Space):
        def __init__(self, n):
            self.n = n

    class Tuple(Space):
        def __init__(self, spaces_):
            self.spaces = tuple(spaces_)

    registry = {}

    def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
        registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs or {})

    spaces.Space, spaces.Box, spaces.Discrete, spaces.Tuple = Space, Box, Discrete, Tuple
    gym.Env, gym.spaces, gym.register, gym.registry = Env, spaces, register, registry
    return gym, spaces


class _H5Dataset:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, key):
        return self._arr[key]


class FakeH5File:
    """``h5py.File(path, 'r')`` stand-in backed by ``FakeH5File.registry[path]`` (dict of arrays)."""
    registry = {}

    def __init__(self, path, mode="r"):
        self._d = FakeH5File.registry[path]

    def __getitem__(self, key):
        return _H5Dataset(self._d[key])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _fake_skimage_resize(image, size, anti_aliasing=True):
    # skimage.transform.resize on a same-size uint8 image returns float64 in [0,1]
    assert tuple(image.shape[:2]) == tuple(size), "harness only supports same-size frames"
    return image.astype(np.float64) / 255.0


_installed = False


def install():
    """Put the stand-ins and the reference tree on sys.path / sys.modules (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    gym, spaces = _make_gym()
    sys.modules.setdefault("gym", gym)
    sys.modules.setdefault("gym.spaces", spaces)
    h5 = types.ModuleType("h5py")
    h5.File = FakeH5File
    sys.modules.setdefault("h5py", h5)
    sk = types.ModuleType("skimage")
    skio = types.ModuleType("skimage.io")
    skt = types.ModuleType("skimage.transform")
    skt.resize = _fake_skimage_resize
    sk.io, sk.transform = skio, skt
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.io", skio)
    sys.modules.setdefault("skimage.transform", skt)
    # namespace stubs: do not execute environments/__init__.py (imports live simulators)
    for name, rel in (("environments", "environments"),
                      ("environments.gym_graph", "environments/gym_graph"),
                      ("environments.gym_ai2thor", "environments/gym_ai2thor"),
                      ("environments.gym_ai2thor.envs", "environments/gym_ai2thor/envs")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules.setdefault(name, m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def ref_modules():
    """Returns the reference modules on the hot path, imported unmodified."""
    install()
    m = types.SimpleNamespace()
    m.util = importlib.import_module("graph.util")
    m.core = importlib.import_module("graph.core")
    m.maze_graph = importlib.import_module("graph.maze_graph")
    m.thor_world = importlib.import_module("graph.multi_graph_no_tp")
    m.graph_env = importlib.import_module("graph.env")
    m.gym_graph = importlib.import_module("environments.gym_graph.graph")
    m.cached = importlib.import_module("environments.gym_ai2thor.envs.cached")
    return m


def ref_thor_cached_tasks():
    """environments/gym_thor_cached.py (THORCachedEnv, the UNFINISHED multi-scene rewrite of cached.py) imported
    unmodified.  As written the module cannot run: ``np`` and ``resize`` are used without being imported (:59-61), and
    ``process`` (:72-95) reads ``_transition_graph`` / ``_current_state_idx`` / ``_current_goal_idx`` / ``last_state``,
    which nothing in the class ever sets.  The harness supplies exactly those missing names and nothing else:

      * module globals ``np`` (numpy) and ``resize`` (the same-size stand-in of skimage.transform.resize used for
        cached.py: identity on a float image of the requested size);
      * ``h5py.File`` reads the in-memory registry (``THOR_DATASET_PATH=mem:`` -> ``mem:/<scene>.h5``);
      * ``A8Driver`` below keeps the four attributes in step with ``state`` / ``goal`` / ``current_scene`` around every
        ``process`` call, the way the finished predecessor keeps its own (cached.py:49-57, 76-97).

    ``reset`` / ``_sample_start`` / ``observe`` / ``process`` themselves execute as they are in the reference."""
    install()
    os.environ["THOR_DATASET_PATH"] = "mem:"
    mod = importlib.import_module("environments.gym_thor_cached")
    mod.np = np
    mod.resize = lambda image, size, anti_aliasing=True: _same_size(image, size)
    return mod


def _same_size(image, size):
    assert tuple(image.shape[:2]) == tuple(size), "harness only supports same-size frames"
    return image


class A8Driver:
    """Single-env ``reset() / step(a)`` surface over an unmodified THORCachedEnv instance (see ref_thor_cached_tasks)."""

    def __init__(self, env):
        self.e = env

    def _sync(self):
        e = self.e
        e._transition_graph = e.current_scene["transition_graph"]
        e._current_state_idx, e._current_goal_idx = e.state, e.goal

    def reset(self):
        e = self.e
        ob = e.reset()                                   # :45-50 + observe() :52-53: raw uint8 (obs, goal)
        self._sync()
        e.last_state = {"image": e._preprocess_frame(ob[0]), "goal": e._preprocess_frame(ob[1])}
        return ob

    def step(self, action):
        e = self.e
        self._sync()
        state, reward, terminal, info = e.process(action)        # :72-95, unmodified
        e.state = int(e._current_state_idx)
        e.last_state = state
        return state, reward, terminal, info


def ref_aux_trainer():
    """experiments/ai2_auxiliary/trainer.py with ``deep_rl`` stubbed; ``autocrop_observations``
    is OUR restatement (oracle.rollout.autocrop_observations) because deep_rl is absent -
    so only the avg_pool half of compute_auxiliary_target is pinned by the reference."""
    install()
    from oracle import rollout as _r
    import torch

    def _stub(name, **attrs):
        mod = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        return mod

    class UnrealTrainer:
        def __init__(self, *a, **k):
            pass

    def _autocrop(x, cell, output_size=None):
        return torch.from_numpy(_r.autocrop_observations(x.numpy(), cell, output_size))

    _stub("deep_rl")
    _stub("deep_rl.a2c_unreal", UnrealTrainer=UnrealTrainer)
    _stub("deep_rl.a2c_unreal.unreal", without_last_item=lambda x: x)
    _stub("deep_rl.a2c_unreal.util", autocrop_observations=_autocrop)
    _stub("deep_rl.common")
    _stub("deep_rl.common.pytorch", to_tensor=lambda x, d=None: x)
    for name, rel in (("experiments", "experiments"), ("experiments.ai2_auxiliary", "experiments/ai2_auxiliary")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules.setdefault(name, mod)
    return importlib.import_module("experiments.ai2_auxiliary.trainer")


_EXPERIMENT = {"get_graph": None, "model_calls": []}      # what the (import-once) experiment module currently sees


def ref_experiment(get_graph, model_calls=None):
    """experiments/thor_cached_auxiliary.py imported UNMODIFIED: ``Trainer`` (:26-56), ``create_envs`` (:58-71) and
    ``default_args`` (:73-84).  What the module needs from outside the tree is supplied as follows:

      * ``deep_rl`` (un-vendored): the wrapper stack and the VecEnv are the restatements of oracle/vec.py
        (RewardCollector, TransposeImage, ScaledFloatFrame, UnrealEnvBaseWrapper, SubprocVecEnv = in-process);
        ``register_trainer`` / schedules / tester classes are inert stand-ins;
      * ``environments.make(id=..., **kwargs)``: gym.make for the ids registered in environments/gym_graph/__init__.py
        (entry point + TimeLimit(max_episode_steps)), building the reference's own env class;
      * ``get_graph(name)`` (environments/gym_graph/download.py:37-75 reads ~/.visual_navigation/scenes): the callable
        given here, returning a ThorGridWorld built from a synthetic scene;
      * ``models.AuxiliaryBigGoalHouseModel``: a recorder - ``model_calls`` receives the constructor arguments.
    """
    install()
    from oracle import vec as ovec
    import torch  # noqa: F401  (the module imports it)

    def _stub(name, **attrs):
        mod = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        return mod

    ref_aux_trainer()          # deep_rl.a2c_unreal.* stubs + experiments namespace packages
    _EXPERIMENT["get_graph"] = get_graph
    _EXPERIMENT["model_calls"] = model_calls if model_calls is not None else []
    if "experiments.thor_cached_auxiliary" in sys.modules:      # imported once per process; only the state changes
        return sys.modules["experiments.thor_cached_auxiliary"]

    class _Model:
        def __init__(self, *args, **kwargs):
            _EXPERIMENT["model_calls"].append((args, kwargs))

    class _Schedule:
        def __init__(self, *a, **k):
            self.args = a

    def register_trainer(**kwargs):
        def deco(cls):
            cls.registered_kwargs = kwargs
            for k, v in kwargs.items():      # deep_rl exposes the registration kwargs as trainer attributes
                setattr(cls, k, v)           # (Trainer.__init__ reads self.max_time_steps, :36)
            return cls
        return deco

    class _UnrealTrainer:
        def __init__(self, *a, **k):
            self.max_time_steps = k.get("max_time_steps", 2e6)

    sys.modules["deep_rl.a2c_unreal"].UnrealTrainer = _UnrealTrainer
    _stub("deep_rl", register_trainer=register_trainer)
    _stub("deep_rl.common.env", RewardCollector=ovec.RewardCollectorWrapper, TransposeImage=ovec.TransposeImageWrapper,
          ScaledFloatFrame=ovec.ScaledFloatFrameWrapper)
    _stub("deep_rl.common.vec_env", DummyVecEnv=ovec.InProcessVecEnv, SubprocVecEnv=ovec.InProcessVecEnv)
    sys.modules["deep_rl.a2c_unreal.util"].UnrealEnvBaseWrapper = ovec.UnrealEnvBaseWrapper
    _stub("deep_rl.configuration", configuration=types.SimpleNamespace())
    _stub("deep_rl.common.schedules", LinearSchedule=_Schedule, MultistepSchedule=_Schedule)
    _stub("deep_rl.model", TimeDistributed=object, Flatten=object, MaskedRNN=object)
    _stub("deep_rl.common.tester", TestingEnv=type("TestingEnv", (), {}), TestingVecEnv=type("TestingVecEnv", (), {}))
    _stub("models", AuxiliaryBigGoalHouseModel=_Model)
    _stub("environments.gym_house")
    _stub("environments.gym_house.multi", create_multiscene=lambda *a, **k: None)

    gym_graph = importlib.import_module("environments.gym_graph.graph")
    gym_graph.get_graph = lambda name: _EXPERIMENT["get_graph"](name)
    registry = {   # environments/gym_graph/__init__.py:18-28
        "OrientedGraph-v0": (gym_graph.OrientedGraphEnv, 900),
        "AuxiliaryGraph-v0": (gym_graph.GoalGymGraphAuxiliaryEnv, 900),
    }

    def make(id, **kwargs):
        cls, limit = registry[id]
        return ovec.TimeLimitWrapper(cls(**kwargs), limit)

    sys.modules["environments"].make = make
    return importlib.import_module("experiments.thor_cached_auxiliary")


# --------------------------------------------------------------------------- stream injection
class InjectedRandom:
    """Replaces the module-level ``random`` seen by a reference env module so that
    ``random.choice`` / ``random.randrange`` consume a supplied integer stream
    (reference: environments/gym_graph/graph.py:47, graph/env.py:181, cached.py:39-42)."""

    def __init__(self, stream):
        self.stream = list(stream)
        self.pos = 0

    def _next(self):
        v = self.stream[self.pos]
        self.pos += 1
        return int(v)

    def choice(self, seq):
        return seq[self._next() % len(seq)]

    def randrange(self, n):
        return self._next() % n

    def Random(self, x=None):  # cached.py:24 builds random.Random(x=seed)
        return self


class InjectedChoice:
    """Replacement for ``np.random.choice`` as called by graph/util.py:102,116,132,142.
    Consumes a stream of uint32 and maps r -> the floor(r * k / 2^32)-th candidate with positive
    weight (k = number of positive-weight candidates).  Every call is recorded as
    ``(n, p, idx)`` so the generator can pin the reference's candidate set AND its weights;
    the chosen start state itself is then injected into the device path, so the r -> idx rule
    used here does not have to match any device RNG."""

    def __init__(self, stream):
        self.stream = list(stream)
        self.pos = 0
        self.calls = []

    def __call__(self, a, p=None):
        r = int(self.stream[self.pos]) & 0xFFFFFFFF
        self.pos += 1
        n = len(a)
        if p is None:
            idx = (r * n) >> 32
        else:
            p = np.asarray(p, dtype=np.float64)
            pos = np.nonzero(p > 0)[0]
            idx = int(pos[(r * len(pos)) >> 32])
        self.calls.append((n, None if p is None else p.copy(), int(idx)))
        return a[idx]
