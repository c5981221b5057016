#!/usr/bin/env python
"""bench.py - env-steps/s of the cached-graph navigation hot path (step + reset + frame gather).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, one process per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]       # reference arm: CPU port on all host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...           # N > 1 (the driver launches this)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): one synthetic
thor-cached scene, 1,500 positions x 4 rotations = 6,000 states, 84x84 RGB + depth + goal frame,
4,096 envs per GPU (weak scaling: every GPU steps its own 4,096 envs, store replicated), uniform
random actions, TimeLimit 900, no curriculum by default (start states uniform over the scene, so the
4,096 envs read ~4,096 different frames per step and the gather is HBM-bound; with the reference's
initial hardness 0.01 - `--hardness 0.01` - the envs cluster around the goals and most reads hit L2).
One "step" = one vectorised step of all envs.  Rank 0 prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ENVS_PER_GPU = 4096
_OVERRIDE = {}
N_CELLS, GRID = 1500, (50, 60)
MAX_EPISODE_STEPS = 900
HARDNESS = None
SCENE_SEED = 0
F_RGB, F_DEPTH = 84 * 84 * 3, 84 * 84
WORKLOAD = "C2 synthetic thor-cached scene: 1,500 positions x 4 rotations, 84x84 RGB+depth+goal, 4,096 envs per GPU"


def make_scene(vn):
    return vn.scenes.make_thor_scene(N_CELLS, GRID, seed=SCENE_SEED, n_goals=4, planes=("rgb", "depth"))


F_SEG = F_RGB = 84 * 84 * 3
PLANE_BYTES = {"rgb": 84 * 84 * 3, "depth": 84 * 84, "segmentation": 84 * 84 * 3}


def make_workload(vn, name):
    """BASELINE.json configs -> (description, world, envs per GPU, obs_layout).  configs[1] (c2) is the one
    the headline metric is quoted on; the others are secondary lines (`--workload`)."""
    S, T = vn.scenes, vn.tables
    if name == "c2":
        return WORKLOAD, vn.compile_world([make_scene(vn)], vn.GYM_GRAPH), ENVS_PER_GPU, "rgbd_goal"
    if name == "c1":
        sc = S.make_maze_scene((10, 10), 0.25, 0, n_goals=1)
        return ("C1 10x10 grid maze, 16 envs, 84x84 aux observation (rgb, goal, depth, seg, goal seg)",
                vn.compile_world([sc], vn.GYM_GRAPH), 16, "aux5")
    if name == "c3":
        sc = S.make_dungeon_scene((64, 64), 0, oriented=True, planes=("rgb",))
        return ("C3 dungeon 64x64 multi-room (oriented), 65,536 envs per GPU, RGB + goal image, auto-reset",
                vn.compile_world([sc], vn.GYM_GRAPH), 65536, "pair")
    if name == "c4":
        scs = [S.make_thor_scene(N_CELLS, GRID, seed=k, n_goals=4, planes=("rgb", "depth"), scene_id=k) for k in range(30)]
        return ("C4 30 synthetic thor-cached scenes resident (180,000 states, 5.1 GB), 32,768 envs per GPU, RGB+depth+goal",
                vn.compile_world(scs, vn.GYM_GRAPH), 32768, "rgbd_goal")
    if name == "rgb":
        return ("84x84 RGB-only cached-graph nav (north_star target line), C2 scene, 4,096 envs per GPU",
                vn.compile_world([make_scene(vn)], vn.GYM_GRAPH), ENVS_PER_GPU, "frame")
    raise ValueError(name)


def layout_bytes(vn, layout):
    """(observation bytes per env-step, goal bytes per reset) of an obs layout."""
    leaves = importlib.import_module("a2cat-vn-pytorch_b200.vec_env").resolve_layout(layout)
    names = list(leaves.values()) if isinstance(leaves, dict) else (list(leaves) if isinstance(leaves, tuple) else [leaves])
    obs = sum(PLANE_BYTES[x] for x in dict.fromkeys(n for n in names if not n.startswith("goal_")))
    goal = sum(PLANE_BYTES[x[5:]] for x in dict.fromkeys(n for n in names if n.startswith("goal_")))
    return obs, goal


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic():
    """DRAM bytes per gather launch from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("gather_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ CPU port
_CPU = {}


def _cpu_world(n_envs, seed):
    """Oracle envs of the bench workload with the reference's cost profile (frames in RAM, per-reset
    candidate enumeration).  Built once per process (workers inherit it by fork)."""
    from oracle import envs as oenvs, vec as ovec
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    if "scene" not in _CPU:
        scene = make_scene(vn)
        _CPU["scene"] = scene
        _CPU["osc"] = oenvs.OracleScene(scene, with_all_pairs=True, cache_frames=True)
    scene, osc = _CPU["scene"], _CPU["osc"]
    envs = []
    for i in range(n_envs):
        goal = scene.goals[i % len(scene.goals)]          # one env per (scene, goal) task, dealt round-robin
        e = oenvs.GymGraphRgbdGoalEnv(osc, goals=goal)
        e.set_complexity(_CPU.get("hardness", HARDNESS))
        e.reset_source = ovec.ReferenceStyleResetSource(osc, [goal], lambda t, e=e: e.optimal_distance(), seed + i)
        tl = ovec.TimeLimit(e, MAX_EPISODE_STEPS)
        envs.append(ovec.RewardCollector(tl))
    ve = ovec.VecEnv(envs)
    ve._time_limits = [w.env for w in envs]
    return ve


def _randomise_phases(ve, seed):
    """After reset(): put every env at a random phase of its 900-step episode so that a bounded sample sees
    time-limit resets at the steady-state rate (1 / 900 per env-step) instead of none at all."""
    rng = np.random.RandomState(seed + 4242)
    for tl in ve._time_limits:
        tl._elapsed = int(rng.randint(0, MAX_EPISODE_STEPS))


def _cpu_worker(conn, n_envs, seed):
    """baselines SubprocVecEnv worker protocol: step -> auto-reset -> send obs over a pipe."""
    ve = _cpu_world(n_envs, seed)
    while True:
        cmd, data = conn.recv()
        if cmd == "step":
            conn.send(ve.step(data)[:3])
        elif cmd == "reset":
            ve.reset()
            _randomise_phases(ve, seed)
            conn.send(None)
        else:
            conn.close()
            return


def cpu_run(n_envs, steps, warmup, workers, seed=0):
    """Returns (env-steps/s, resets) of the CPU port on `workers` processes."""
    import multiprocessing as mp
    rng = np.random.RandomState(seed)
    if workers <= 1:
        ve = _cpu_world(n_envs, seed)
        ve.reset()
        _randomise_phases(ve, seed)
        for _ in range(warmup):
            ve.step(rng.randint(0, 4, n_envs))
        t0 = time.perf_counter()
        nres = 0
        for _ in range(steps):
            _, _, d, _ = ve.step(rng.randint(0, 4, n_envs))
            nres += int(d.sum())
        dt = time.perf_counter() - t0
        return n_envs * steps / dt, nres, dt
    _cpu_world(0, seed)                                   # build the scene before forking
    ctx = mp.get_context("fork")
    per = [n_envs // workers + (1 if r < n_envs % workers else 0) for r in range(workers)]
    pipes, procs = [], []
    for r, n in enumerate(per):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=(b, n, seed + 100000 * r), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)

    def vec(cmd, acts=None):
        off = 0
        for n, c in zip(per, pipes):
            c.send((cmd, None if acts is None else acts[off:off + n]))
            off += n
        outs = [c.recv() for c in pipes]
        if cmd == "reset":
            return None
        obs = tuple(np.concatenate([o[0][0][k] for o in outs]) for k in range(3))      # parent-side stack
        return obs, np.concatenate([o[1] for o in outs]), np.concatenate([o[2] for o in outs])

    vec("reset")
    for _ in range(warmup):
        vec("step", rng.randint(0, 4, n_envs))
    t0 = time.perf_counter()
    nres = 0
    for _ in range(steps):
        _, _, d = vec("step", rng.randint(0, 4, n_envs))
        nres += int(d.sum())
    dt = time.perf_counter() - t0
    for c in pipes:
        c.send(("close", None))
    for p in procs:
        p.join(5)
    return n_envs * steps / dt, nres, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _CPU["hardness"] = None if args.hardness in (None, "none") else float(args.hardness)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    steps = args.steps if args.steps else 300
    # bounded sample of the workload: 16 envs per worker, fewer when K is large (<= ~1.5 M env-steps in all)
    n_envs = max(workers, min(16 * workers, int(1.5e6 / max(1, steps))))
    warm = args.warmup if args.warmup is not None else 5
    value, nres, dt = cpu_run(n_envs, steps, warm, workers)
    line = {
        "metric": "env-steps/s (obs gather+step+reset)", "value": value, "unit": "env-steps/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d envs x %d vector steps" % (n_envs, steps)},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": workers, "kind": "port",
                         "sample": "%d envs x %d vector steps over %d worker processes (pipes), oracle port of "
                                   "GoalGymGraphAuxiliaryEnv incl. per-reset candidate enumeration, episode phases "
                                   "randomised to the steady state; p_reset=%.5f"
                                   % (n_envs, steps, workers, nres / max(1, n_envs * steps))},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    L = vn.lib
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    K = args.steps if args.steps else 20000
    W = args.warmup if args.warmup is not None else 200
    W = max(W, 3)

    workload, world, envs_per_gpu, layout = make_workload(vn, args.workload)
    F_OBS, F_GOAL = layout_bytes(vn, layout)
    n_total = (args.envs_per_gpu or envs_per_gpu) * world_size
    env = vn.GraphVecEnv(world, n_total, device=dev, seed=1, max_episode_steps=MAX_EPISODE_STEPS,
                         obs_layout=layout, unreal_wrapper=True, rank=rank, world_size=world_size,
                         gather=args.gather, host_outputs=False)
    hardness = None if args.hardness in (None, "none") else float(args.hardness)
    _CPU["hardness"] = hardness
    env.set_complexity(hardness)
    N = env.num_envs
    # action stream resident in HBM: cyclic buffer of uniform random actions (Philox via torch generator)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_rows = min(K + W, 2048)
    actions = torch.randint(0, 4, (n_rows, N), device=dev, generator=gen, dtype=torch.int32)
    env.reset()

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    graph_len = 0
    l0 = env.kernel_launches
    env.step_enqueue(actions[0])
    launches_per_step = env.kernel_launches - l0        # 1 (fused launch, small batches) or 2 (scalar + gather)
    if args.cuda_graph:
        # launch-bound batches (C1): replay CUDA graphs of `graph_len` steps instead of 2 launches per step
        graph_len = min(64, n_rows)
        graph = env.capture_steps(actions[:graph_len])

    def device_loop(k0, k):
        if graph_len:
            for _ in range((k + graph_len - 1) // graph_len):
                graph.replay()
            return
        for i in range(k0, k0 + k):
            # the action stream is pre-computed and resident: VN_STEP_ACTIONS_READY lets the scalar kernel of
            # step i + 1 run while the gather of step i is still copying
            env.step_enqueue(actions[i % n_rows], actions_ready=True)

    device_loop(0, args.mix)      # un-timed: lets the state distribution settle (random-walk mixing)
    device_loop(0, W)
    env.stats.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = env.kernel_launches
    if graph_len:
        K = ((K + graph_len - 1) // graph_len) * graph_len      # whole graphs
    ev0.record()
    device_loop(W, K)
    if world_size > 1:
        # the only collective near this path: episode statistics, once per logging interval
        stats_vec = env.stats.to(torch.float64)
        dist.all_reduce(stats_vec)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = launches_per_step * K if graph_len else env.kernel_launches - launches0
    if world_size > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stats = env.episode_stats(reduce=True)
    value = n_total * K / (ms * 1e-3)
    p_reset = stats["resets"] / max(1.0, stats["steps"])
    coll = stats["collisions"] / max(1.0, stats["steps"])
    p_skip = stats["rows_skipped"] / max(1.0, stats["steps"])   # rows whose record did not change: not copied again

    # ---- roofline of the dominant kernel (the frame gather): instrumented pass, events around K2 only
    Kr = min(K, 500)
    es = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kr)]
    import ctypes as C
    stream = torch.cuda.current_stream(dev).cuda_stream
    env.stats.zero_()
    torch.cuda.synchronize(dev)
    for i in range(Kr):
        a = actions[(W + K + i) % n_rows]
        env._tick(env._c_out, env._step_flags)      # serial mode: next descriptor half, no overlap with the previous gather
        L.check(env.lib.vn_env_step_scalar(C.byref(env.dw.tables), C.byref(env._c_envs), C.byref(env._c_rules), None,
                                           a.data_ptr(), C.byref(env._c_out), stream))
        es[i][0].record()
        L.check(env.lib.vn_env_gather(C.byref(env.dw.store), C.byref(env._c_envs), C.byref(env._c_out), env.gather,
                                      stream))
        es[i][1].record()
    torch.cuda.synchronize(dev)
    gather_ms = float(np.mean([a.elapsed_time(b) for a, b in es]))
    rs = env.episode_stats()
    p_reset_r = rs["resets"] / max(1.0, rs["steps"])
    p_skip_r = rs["rows_skipped"] / max(1.0, rs["steps"])
    # algorithmic bytes of one launch: one read + one write of every observation row that
    # CHANGED (rows of envs that collided hold the right frames already and are skipped, like the goal rows of envs
    # that did not reset), plus the goal rows of the envs that reset
    iso_bytes = N * ((1 - p_skip_r) * 2 * F_OBS + p_reset_r * 2 * F_GOAL)
    peak, peak_src = measured_peak()
    iso_achieved = iso_bytes / (gather_ms * 1e-3) / 1e9
    # in the timed region the scalar kernel of step k+1 overlaps the gather of step k (pipelined mode), so the
    # gather's launch-to-launch period there is the step time: that is the kernel's duration in situ (an upper
    # bound of it - everything else the step does is inside)
    alg_bytes = N * ((1 - p_skip) * 2 * F_OBS + p_reset * 2 * F_GOAL)
    situ_ms = ms / K
    achieved = alg_bytes / (situ_ms * 1e-3) / 1e9
    # calibration: plain contiguous device copies of the SAME number of bytes (torch copy_, the operation the
    # measured peak was taken with, but at this kernel's size instead of 2 GiB), back to back over 4 distinct
    # (src, dst) pairs so that, like the gather in steady state, every copy starts with L2 full of the previous
    # copy's dirty lines - what a ~40 us transfer can sustain on this GPU
    nb = N * F_OBS
    pairs = [(torch.empty(nb, dtype=torch.uint8, device=dev).random_(0, 255),
              torch.empty(nb, dtype=torch.uint8, device=dev)) for _ in range(4)]
    for i in range(8):
        pairs[i % 4][1].copy_(pairs[i % 4][0])
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    c0.record()
    for i in range(200):
        pairs[i % 4][1].copy_(pairs[i % 4][0])
    c1.record()
    torch.cuda.synchronize(dev)
    copy_ms = c0.elapsed_time(c1) / 200
    del pairs

    kname = "vn_step_fused_kernel" if launches_per_step == 1 else \
        "vn_gather_%s_kernel" % ("ldg" if args.gather == "ldg" else "bulk")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # the committed ncu capture is of the C2 bulk gather; other workloads / variants have none
                "traffic": profiled_traffic() if (args.workload == "c2" and launches_per_step == 2 and
                                                  args.gather in ("auto", "bulk") and hardness is None) else None,
                "kernel": kname,
                "kernel_ms": situ_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "how": "in situ: CUDA events around the K timed steps / K launches (the scalar kernel overlaps the "
                       "previous gather, so this is the gather's launch-to-launch period on this rank)",
                "rows_skipped_rate": p_skip,
                "isolated": {"kernel_ms": gather_ms, "achieved": iso_achieved, "frac": iso_achieved / peak,
                             "algorithmic_bytes_per_launch": iso_bytes,
                             "how": "CUDA events around each gather launch alone (serialised, includes launch latency)"},
                "same_size_copy_ms": copy_ms, "same_size_copy_gbs": 2 * nb / (copy_ms * 1e-3) / 1e9}

    # ---- secondary line, same store: RGB-only observation (the north-star "84x84 cached-graph nav" target of
    # >= 1e9 env-steps/s on 8 GPUs refers to this 42,336 B/step variant)
    secondary = {}
    if args.workload == "c2" and not args.quick:
        env_r = vn.GraphVecEnv(world, n_total, device=dev, seed=3, max_episode_steps=MAX_EPISODE_STEPS,
                               obs_layout="frame", unreal_wrapper=True, rank=rank, world_size=world_size,
                               gather=args.gather, host_outputs=False, device_world=env.dw)
        env_r.reset()
        Kr2 = min(K, 5000)
        for i in range(300 + W):
            env_r.step_enqueue(actions[i % n_rows], actions_ready=True)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for i in range(Kr2):
            env_r.step_enqueue(actions[i % n_rows], actions_ready=True)
        r1.record()
        barrier()
        ms_r = r0.elapsed_time(r1)
        if world_size > 1:
            t = torch.tensor([ms_r], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_r = float(t.item())
        secondary["rgb_only"] = {"value": n_total * Kr2 / (ms_r * 1e-3), "unit": "env-steps/s", "steps": Kr2,
                                 "ms_per_step": ms_r / Kr2, "algorithmic_bytes_per_env_step": 2 * F_RGB,
                                 "frac_if_every_row_were_copied": N * 2 * F_RGB / (ms_r * 1e-3 / Kr2) / 1e9 / peak}
        del env_r
        # secondary line: float32 CHW observations in [0, 1] - what the reference's wrappers hand to its model
        # (TransposeImage + ScaledFloatFrame, thor_cached_auxiliary.py:61-62) - converted straight from the store
        # into persistent batches (no uint8 batch in this mode): 4 bytes written per byte read
        env_f = vn.GraphVecEnv(world, n_total, device=dev, seed=4, max_episode_steps=MAX_EPISODE_STEPS,
                               obs_layout=layout, unreal_wrapper=True, rank=rank, world_size=world_size,
                               gather=args.gather, host_outputs=False, device_world=env.dw, scaled_float=True)
        env_f.reset()
        Kf = min(K, 2000)
        for i in range(300 + W):
            env_f.step_enqueue(actions[i % n_rows], actions_ready=True)
        env_f.stats.zero_()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(Kf):
            env_f.step_enqueue(actions[i % n_rows], actions_ready=True)
        f1.record()
        barrier()
        ms_f = f0.elapsed_time(f1)
        if world_size > 1:
            t = torch.tensor([ms_f], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_f = float(t.item())
        fs = env_f.episode_stats()
        skip_f, reset_f = fs["rows_skipped"] / max(1.0, fs["steps"]), fs["resets"] / max(1.0, fs["steps"])
        bytes_f = N * 5 * ((1 - skip_f) * F_OBS + reset_f * F_GOAL)
        secondary["float_chw"] = {"value": n_total * Kf / (ms_f * 1e-3), "unit": "env-steps/s", "steps": Kf,
                                  "ms_per_step": ms_f / Kf, "algorithmic_bytes_per_step": bytes_f,
                                  "frac": bytes_f / (ms_f * 1e-3 / Kf) / 1e9 / peak,
                                  "note": "float32 CHW leaves (rgb, goal, depth) / 255 in persistent batches"}
        del env_f
        # secondary line: one whole A2C / UNREAL data pass as the trainer sees it (thor_cached_auxiliary.py:30-42:
        # num_steps = 20): 20 vectorised steps that write their rollout rows themselves, then n-step returns,
        # pixel-control rewards + their back-up and reward-prediction labels, all on the device
        R = vn.rollout
        T_roll = 20
        rb = R.RolloutBuffer(env.dw, N, T_roll)
        v_last = torch.randn(N, device=dev)
        q_last = torch.rand(N, 400, device=dev)
        R.target_tables(env.dw, 4, (20, 20))                # per-world pixel-control table, built once

        def a2c_pass(i0):
            rb.start(env)
            for t in range(T_roll):
                rb.step(env, actions[(i0 + t) % n_rows], actions_ready=True)
            ret = rb.returns(v_last, 0.99)
            pc = rb.pixel_control(4, (20, 20))
            pcr = R.discounted_backup(pc.view(N, T_roll, 400), rb.dones.t().contiguous(), q_last, 0.9)
            return ret, pcr, rb.reward_prediction()

        for i in range(5):
            a2c_pass(i * T_roll)
        Kp = 100
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(Kp):
            a2c_pass(i * T_roll)
        p1.record()
        barrier()
        ms_p = p0.elapsed_time(p1)
        if world_size > 1:
            t = torch.tensor([ms_p], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_p = float(t.item())
        secondary["a2c_pass"] = {"value": n_total * T_roll * Kp / (ms_p * 1e-3), "unit": "env-steps/s",
                                 "ms_per_pass": ms_p / Kp, "num_steps": T_roll,
                                 "note": "20 env steps (rollout rows written by the step kernel) + n-step returns + "
                                         "pixel-control rewards and back-up + RP labels, per pass"}

    if args.quick:
        if rank == 0:
            emit({"value": value, "ms_per_step": ms / K, "frac": roofline["frac"],
                  "iso_kernel_ms": gather_ms, "iso_frac": roofline["isolated"]["frac"],
                  "p_reset": p_reset, "p_skip": p_skip, "launches_per_step": launches_per_step})
        if world_size > 1:
            dist.destroy_process_group()
        return

    # ---- e2e through the public VecEnv API: host actions in, host rewards/dones out, every step
    env_e = vn.GraphVecEnv(world, n_total, device=dev, seed=2, max_episode_steps=MAX_EPISODE_STEPS,
                           obs_layout=layout, unreal_wrapper=True, rank=rank, world_size=world_size,
                           gather=args.gather, host_outputs=True, device_world=env.dw)
    env_e.reset()
    Ke = min(K, 3000)
    host_actions = actions[:min(n_rows, 512)].cpu().numpy()
    for i in range(20):
        env_e.step(host_actions[i % len(host_actions)])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        obs, rew, done, infos = env_e.step(host_actions[i % len(host_actions)])
    barrier()
    dt = time.perf_counter() - t0
    if world_size > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": n_total * Ke / dt, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * N, "d2h_bytes_per_step": 20 * N,
           "steps": Ke, "note": "VecEnv.step(numpy actions) -> (CUDA uint8 obs, numpy rewards, numpy dones, infos); "
                                "observations stay in HBM for the policy"}
    # secondary: also bring the observation batch to pinned host memory every step (what a CPU policy would need)
    obs_leaves = list(env_e.obs_buf.values())
    pin = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in obs_leaves]
    Kh = min(Ke, 200)
    barrier()
    t0 = time.perf_counter()
    for i in range(Kh):
        obs, rew, done, infos = env_e.step(host_actions[i % len(host_actions)])
        for dst, src in zip(pin, obs_leaves):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    dth = time.perf_counter() - t0
    e2e["host_obs_value"] = n_total * Kh / dth if world_size == 1 else None
    e2e["host_obs_d2h_bytes_per_step"] = N * F_OBS + 20 * N

    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline and args.workload == "c2":
        # bounded sample (10-30 s of CPU work): 32 envs x 8,000 vector steps cross the 900-step TimeLimit about nine
        # times per env, so the reference's per-reset candidate enumeration (~0.1 s each here) is included
        v, nres, cdt = cpu_run(32, 8000, 3, 1)
        cpu_baseline = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": "32 envs x 8000 vector steps, one process, sequential + np.stack (DummyVecEnv "
                                  "equivalent), episode phases randomised, %.1f s, p_reset=%.5f"
                                  % (cdt, nres / (32 * 8000.0))}

    if rank == 0:
        line = {
            "metric": "env-steps/s (obs gather+step+reset)", "value": value, "unit": "env-steps/s",
            "n_gpus": world_size, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "envs_total": n_total, "states": world.n_states,
                       "store_bytes": env.dw.nbytes(), "batch_bytes_per_step": N * F_OBS,
                       "l2": "inputs larger than L2: %.0f MB store + %.0f MB batch written per step vs 126 MB L2"
                             % (env.dw.nbytes() / 1e6, N * F_OBS / 1e6),
                       "gather": args.gather, "p_reset": p_reset, "collision_rate": coll, "rows_skipped_rate": p_skip,
                       "launches_per_step": launches_per_step,
                       "max_episode_steps": MAX_EPISODE_STEPS, "hardness": hardness, "mix_steps": args.mix,
                       "cuda_graph_steps": graph_len,
                       "parallelism": "env-sharded x%d, no data-path collective" % world_size},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world_size, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "secondary": secondary,
        }
        emit(line)
    if world_size > 1:
        dist.destroy_process_group()


def run_rollout(args):
    """BASELINE.json configs[4]: the A2C rollout builder - 128-step n-step returns + pixel-control rewards
    and their discounted back-up + reward-prediction labels / index lists (+ the 3 auxiliary targets), on a
    2^20 (env x step) batch, straight from state indices and the HBM store."""
    import torch
    vn = importlib.import_module("a2cat-vn-pytorch_b200")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    N, T = 8192, 128
    K = args.steps if args.steps else 20
    W = max(args.warmup if args.warmup is not None else 3, 3)
    scene = vn.scenes.make_thor_scene(N_CELLS, GRID, seed=SCENE_SEED, n_goals=4, planes=("rgb", "depth", "segmentation"))
    world = vn.compile_world([scene], vn.GYM_GRAPH)
    env = vn.GraphVecEnv(world, N, device=dev, seed=1, max_episode_steps=50, obs_layout="frame", host_outputs=False)
    env.set_complexity(0.05)
    env.reset()
    buf = vn.rollout.RolloutBuffer(env.dw, N, T)
    buf.start(env)
    gen = torch.Generator(device=dev).manual_seed(3)
    for _ in range(T):
        a = torch.randint(0, 4, (N,), device=dev, generator=gen, dtype=torch.int32)
        env.step_enqueue(a)
        buf.insert(env, a)
    states = buf.states.t().contiguous()          # [N, T+1]
    goals = buf.goals[:-1].t().contiguous()
    v_last = torch.randn(N, device=dev)
    q_last = torch.rand(N, 20, 20, device=dev)
    done_bt = buf.dones.t().contiguous()
    R = vn.rollout

    def builder(with_aux):
        ret = R.nstep_returns(buf.rewards, buf.dones, v_last, 0.99, time_major=True)
        pc = R.pixel_control_reward(env.dw, states, 4, (20, 20))
        pcr = R.discounted_backup(pc.view(N, T, 400), done_bt, q_last.view(N, 400), 0.9)
        lab = R.reward_prediction_labels(buf.rewards, with_lists=True)
        aux = R.auxiliary_targets(env.dw, states[:, :-1], goals, 4, (20, 20)) if with_aux else None
        return ret, pc, pcr, lab, aux

    def timed(fn, k):
        for _ in range(W):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / k

    ms_core = timed(lambda: builder(False), K)
    ms_all = timed(lambda: builder(True), max(2, K // 4))
    ms_pc = timed(lambda: R.pixel_control_reward(env.dw, states, 4, (20, 20)), K)
    ms_ret = timed(lambda: R.nstep_returns(buf.rewards, buf.dones, v_last, 0.99, time_major=True), K)
    ms_rp = timed(lambda: R.reward_prediction_labels(buf.rewards, with_lists=True), K)
    peak, peak_src = measured_peak()
    # table path: every transition reads one 1,600-byte table row and writes one 1,600-byte output row; the
    # ~2 % reset transitions read two 21,168-byte frames instead of the table row
    miss = float(buf.dones.float().mean())
    pc_bytes = N * T * (1600 + (1 - miss) * 1600 + miss * 2 * F_RGB)
    line = {
        "metric": "rollout-builder env-steps/s (n-step returns + pixel-control + back-up + RP)", "value": N * T / (ms_core * 1e-3),
        "unit": "env-steps/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_core, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 A2C rollout builder: T=128, N=8,192 (2^20 env-steps), gamma .99, PC cell 4 -> 20x20, gamma_pc .9",
                   "ms": {"returns": ms_ret, "pixel_control": ms_pc, "rp_labels_and_lists": ms_rp, "core_total": ms_core,
                          "with_aux_targets_total": ms_all}, "done_rate": float(buf.dones.float().mean())},
        "roofline": {"bound": "hbm", "achieved": pc_bytes / (ms_pc * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": pc_bytes / (ms_pc * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "vn_transition_rows + vn_gather_rows + vn_pixel_control_list (table-driven pixel control)",
                     "kernel_ms": ms_pc, "algorithmic_bytes_per_launch": pc_bytes, "peak_source": peak_src},
        "gpu_launches": 7,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else any library prints while the
    benchmark runs (NCCL's version banner, warnings) was diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gather", default="auto", choices=["auto", "ldg", "bulk", "fused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true", help="replay 64-step CUDA graphs in the device-resident loop")
    ap.add_argument("--quick", action="store_true", help="development: device-resident number + roofline only")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "rgb"],
                    help="BASELINE.json config; c2 is the headline, the others are secondary lines")
    ap.add_argument("--hardness", default="none", help="curriculum hardness (set_complexity); 'none' = uniform starts")
    ap.add_argument("--mix", type=int, default=1000, help="un-timed steps before warm-up")
    ap.add_argument("--envs-per-gpu", type=int, default=None, help="development: override the 4,096 envs per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_rollout(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
