"""Minimal stand-ins for the ``gym.spaces`` classes the reference builds its observation / action
spaces from (environments/gym_graph/graph.py:26-33,101-106).  gym is not installable offline; when the
real package is importable its classes are used instead so that isinstance checks in a trainer hold."""
import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Box, Discrete, Tuple, Dict  # type: ignore  # noqa: F401
except Exception:
    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Tuple(Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)
