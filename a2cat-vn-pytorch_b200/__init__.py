"""B200-native batched cached-graph navigation environments (hot path of a2cat-vn-pytorch).

The directory name carries a hyphen (it mirrors the reference repo name), so import it with
``importlib.import_module("a2cat-vn-pytorch_b200")`` or through the ``vn_b200`` alias module at
the repository root.
"""
from . import scenes, tables  # noqa: F401

__version__ = "0.1.0"
