#!/usr/bin/env python
"""Device time and HBM roofline fraction of the secondary kernels (development aid; needs a B200):
policy_input (gather + TransposeImage + ScaledFloatFrame), direct auxiliary targets, direct pixel control."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, k=30, w=5):
    import torch
    for _ in range(w):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    R = vn.rollout
    scene = vn.scenes.make_thor_scene(1500, (50, 60), seed=0, n_goals=4, planes=("rgb", "depth", "segmentation"))
    dw = vn.DeviceWorld(vn.compile_world([scene], vn.GYM_GRAPH))
    peak, _ = bench.measured_peak()
    n = 4096
    gen = torch.Generator(device="cuda").manual_seed(0)
    st = torch.randint(0, 6000, (n,), device="cuda", generator=gen, dtype=torch.int32)
    out = {}
    for plane, fb in (("rgb", 21168), ("depth", 7056)):
        ms = timed(lambda: R.policy_input(dw, st, plane))
        out["policy_input_" + plane] = dict(ms=ms, gbs=n * fb * 5 / ms / 1e6, frac=n * fb * 5 / ms / 1e6 / peak)
    ms = timed(lambda: R._aux_direct(dw, st, "segmentation", 4, (20, 20)))
    b = n * (21168 + 3 * 400 * 4)
    out["aux_direct_segmentation"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
    st2 = torch.randint(0, 6000, (256, 17), device="cuda", generator=gen, dtype=torch.int32)
    ms = timed(lambda: R._pixel_control_direct(dw, st2, 4, (20, 20), "rgb"))
    b = 256 * (17 * 21168 + 16 * 1600)
    out["pixel_control_direct"] = dict(ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak, ns_per_transition=ms * 1e6 / (256 * 16))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
