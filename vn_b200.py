"""Import alias: ``import vn_b200 as vn`` == importlib.import_module("a2cat-vn-pytorch_b200")
(the package directory mirrors the reference repo name and therefore contains hyphens)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("a2cat-vn-pytorch_b200")
