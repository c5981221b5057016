"""The example trainer loop (examples/a2c_maze.py) runs against the public API: a smoke test of the whole data path
env -> policy input -> rollout buffer -> n-step returns -> update, with no frame leaving the device."""
import importlib.util
import math
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_a2c_maze_example_runs():
    spec = importlib.util.spec_from_file_location("a2c_maze", os.path.join(ROOT, "examples", "a2c_maze.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist = mod.train(updates=40, num_envs=16, log_every=20, hardness=0.3, quiet=True)
    assert len(hist) == 2
    for rec in hist:
        assert rec["episodes"] > 0 and 0.0 <= rec["success"] <= 1.0 and rec["fps"] > 0
        assert all(math.isfinite(float(v)) for v in rec.values())
