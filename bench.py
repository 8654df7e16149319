#!/usr/bin/env python
"""Benchmark of the environment-step hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

Workload (per GPU, weak scaling): ``MultiAgentInvManagement`` 4-stage chain in preset MA_6 mode
(hyperparams.py:482-484), 65 536 environments, replayed Poisson(5) demand, host-random uniform
actions pre-staged on the device.  One bench *step* = one 30-period episode of the whole batch
(reset + 30 step() launches, replayed as a CUDA graph) = 65 536 x 4 x 30 agent-steps.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: ``roofline`` (step kernel:
algorithmic bytes / CUDA-event time vs. the measured HBM copy peak), ``roofline_large_n`` (same
kernel at 4 Mi envs, working set >> L2), ``replay_fused`` (the same 30 periods as ONE imx_step_many launch),
``cpu_baseline`` (the oracle port on the host cores; ``cpu_baseline_c``: the plain-C OpenMP restatement),
``e2e`` (same metric through the host-buffer ABI with H2D/D2H inside the timed region; ``e2e_f32_obs``: the same
with float32 observations).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

T_PERIODS = 30
M_STAGES = 4
KERNEL_VARIANTS = {0: "imx::step_kernel (ahead-of-time, direct global accesses)",
                   1: "imx::step_kernel_tma (ahead-of-time, TMA-staged tiles)",
                   2: "imx::step_kernel_tma<4,3,1,1,false> (NVRTC-specialised, TMA-staged tiles)"}
ENVS_PER_GPU = 65536
NCU_TRAFFIC_BYTES_65536 = 8402176        # dram__bytes_read.sum + dram__bytes_write.sum, one launch (cold L2), profiles/r1_ncu_step_kernel_final.txt
NCU_TRAFFIC_BYTES_4MI = 537099264 + 1402789000   # same counters at 4 Mi envs, profiles/r1_ncu_step_kernel_tma_specialised_4Mi_envs.txt
WORKLOAD = ("MAIM_env 4-stage serial, MA_6 obs mode (td=T,pd=T,pa=F,P=1, shared reward), step() on "
            "65536 envs per GPU, 30-period episodes, replayed Poisson(5) demand, uniform(-1,1) actions pre-staged")


# ----------------------------------------------------------------------------------------
# CPU baseline: the oracle port (oracle/im_oracle.py) on the host cores, one env per worker
# ----------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, episodes = args
    from marl_for_im_b200 import presets
    from oracle import im_oracle
    cfg = presets.serial4()
    env = im_oracle.OracleEnv("MAIM", cfg)
    rng = np.random.default_rng(seed)
    demand = rng.poisson(5, size=(episodes, T_PERIODS))
    actions = rng.uniform(-1, 1, size=(episodes, T_PERIODS, M_STAGES))
    t0 = time.perf_counter()
    for e in range(episodes):
        env.reset(demand[e])
        for t in range(T_PERIODS):
            env.step(actions[e, t])
    return time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(episodes_per_worker=6000, pool=None):
    """agent-steps/s of the oracle port with one process per host core (bounded sample)."""
    import multiprocessing as mp
    cores = host_cores()
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("fork").Pool(cores)
    try:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(1000 + i, episodes_per_worker) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        if own_pool:
            pool.close()
            pool.join()
    agent_steps = cores * episodes_per_worker * T_PERIODS * M_STAGES
    return {"value": agent_steps / wall, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"{cores} processes x {episodes_per_worker} episodes x {T_PERIODS} periods of the same 4-stage "
                      f"MA_6 workload through oracle/im_oracle.py (one env per process, like the reference); {wall:.2f} s wall"}


def cpu_baseline_c():
    """The plain-C restatement (oracle/imx_oracle.c) with all OpenMP threads on the full 65536-env
    episode: the strongest CPU implementation of the same path we can offer, for context."""
    from marl_for_im_b200 import presets
    from oracle import c_oracle
    N, T, m = ENVS_PER_GPU, T_PERIODS, M_STAGES
    demand = np.random.default_rng(420).poisson(5, size=(N, T)).astype(np.int32)
    actions = np.random.default_rng(0).uniform(-1, 1, size=(T, N, m))
    co = c_oracle.COracle("MAIM", presets.serial4())
    co.run(demand[:1024], actions[:, :1024])               # warm-up
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        co.run(demand, actions)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": N * T * m / best, "unit": "agent-steps/s", "cores": c_oracle.max_threads(), "kind": "port (plain C, OpenMP)",
            "sample": f"full config-2 episode batch ({N} envs x {T} periods), best of 3, {best * 1e3:.1f} ms"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU design (per-env Python/numpy objects), restated by the
    oracle port because the reference tree cannot travel to the GPU box; all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    pool = mp.get_context("fork").Pool(cores)
    per_worker = 400                                       # ~0.3 s of wall clock per step on 16 cores
    try:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, [(i, 4) for i in range(cores)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_worker, [(s * cores + i, per_worker) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    agent_steps = args.steps * cores * per_worker * T_PERIODS * M_STAGES
    value = agent_steps / wall
    sample = (f"each step = {cores} processes x {per_worker} episodes x {T_PERIODS} periods (a bounded sample of the "
              f"65536-env batch), oracle/im_oracle.py port of the reference's per-env numpy path")
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def algorithmic_bytes_per_env_step(env):
    """B = 2*4*S + 4*R + 8*m*(2 + O)."""
    S, R, m, O = env.state_words, len(env._retailers), env.num_nodes, env.obs_len
    return 2 * 4 * S + 4 * R + 8 * m * (2 + O)


class RawEpisode:
    """reset + T step launches through the C ABI on preallocated device buffers (graph-capturable)."""

    def __init__(self, env, demand_dev, actions_dev, obs_bufs, rew_bufs, obs0):
        from marl_for_im_b200 import _lib
        self.env, self.lib, self._lib = env, env._lib, _lib
        self.demand, self.actions, self.obs, self.rew, self.obs0 = demand_dev, actions_dev, obs_bufs, rew_bufs, obs0

    def reset(self, stream):
        self._lib.check(self.lib.imx_reset(self.env._handle, C.c_void_p(self.demand.data_ptr()), None, 0, 1,
                                           C.c_void_p(self.obs0.data_ptr()), C.c_void_p(stream)))

    def steps_many(self, stream, periods):
        """the same periods as ONE imx_step_many call on the stored plan; needs the per-period observation / reward
        buffers to be consecutive slices of one tensor"""
        self._lib.check(self.lib.imx_step_many(self.env._handle, C.c_void_p(self.actions.data_ptr()), periods,
                                               C.c_void_p(self.obs[0].data_ptr()), C.c_void_p(self.rew[0].data_ptr()), None, C.c_void_p(stream)))

    def steps(self, stream, periods):
        h = self.env._handle
        for t in range(periods):
            self._lib.check(self.lib.imx_step(h, C.c_void_p(self.actions[t].data_ptr()),
                                              C.c_void_p(self.obs[t % len(self.obs)].data_ptr()),
                                              C.c_void_p(self.rew[t % len(self.rew)].data_ptr()), None, C.c_void_p(stream)))


def capture(fn, torch):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(s.cuda_stream)            # warm-up outside capture (module load, attribute setup)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().wait_stream(s)
    return g


def timed_replays(graph, reps, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    cpu = cpu_c = cpu_1 = None
    if world == 1 and not args.skip_cpu:
        cpu = cpu_baseline()                       # before CUDA is initialised (fork-safe)
        t1 = _cpu_worker((999, 400))               # the same port on ONE core
        cpu_1 = {"value": 400 * T_PERIODS * M_STAGES / t1, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                 "sample": f"400 episodes x {T_PERIODS} periods in this process; {t1:.2f} s"}
        try:
            cpu_c = cpu_baseline_c()
        except Exception as exc:
            cpu_c = {"error": str(exc)[:200]}

    import torch
    import torch.distributed as dist
    from marl_for_im_b200 import _lib as _lib_mod
    from marl_for_im_b200 import presets
    from marl_for_im_b200.envs import MultiAgentInvManagement

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from marl_for_im_b200.dist import bind_host_to_gpu      # host buffers of the e2e leg next to this rank's GPU
    numa_cores = bind_host_to_gpu(local_rank)                # (2 GPUs: 0.99 -> 1.37 G agent-steps/s end to end)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"

    N, T, m = args.envs, T_PERIODS, M_STAGES
    cfg = presets.serial4()
    env = MultiAgentInvManagement(dict(cfg, num_envs=N, device=str(dev), env_offset=rank * N))
    O = env.obs_len
    rng = np.random.default_rng(420 + rank)
    demand_h = rng.poisson(5, size=(N, 1, T)).astype(np.int32)
    actions_h = np.random.default_rng(rank).uniform(-1, 1, size=(T, N, m))
    demand = torch.as_tensor(demand_h, device=dev)
    actions = torch.as_tensor(actions_h, device=dev)
    obs_all = torch.empty((T, N, m, O), dtype=torch.float64, device=dev)                  # one buffer per period (trajectory storage)
    obs = [obs_all[t] for t in range(T)]
    rew = torch.empty((T, N, m), dtype=torch.float64, device=dev)
    obs0 = torch.empty((N, m, O), dtype=torch.float64, device=dev)
    ep = RawEpisode(env, demand, actions, obs, [rew[t] for t in range(T)], obs0)

    # Episode statistics [n, sum, sum of squares, per-agent ...] accumulate on the device; like the reference's
    # evaluation loops (np.mean / np.std over all test episodes, MA_inv_management.py:591-595) they are
    # reduced ONCE per evaluation batch = the K timed episodes: a single NCCL all-reduce of 3 + 2m doubles.
    stats_acc = torch.zeros(3 + 2 * m, dtype=torch.float64, device=dev)

    def episode(stream):
        ep.reset(stream)
        ep.steps(stream, T)                                  # 30 step() launches: the per-step API an RL loop calls
        _lib_mod.check(env._lib.imx_episode_stats(env._handle, C.c_void_p(rew.data_ptr()), T, None, C.c_void_p(stats_acc.data_ptr()),
                                                  1, C.c_void_p(stream)))

    launches0 = env.launch_count()
    episode(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    launches_per_episode = env.launch_count() - launches0
    g_episode = capture(episode, torch)

    def steps_only(stream):
        env._lib.imx_set_period(env._handle, 0)
        ep.steps(stream, T)

    g_steps = capture(steps_only, torch)

    def steps_only_many(stream):
        env._lib.imx_set_period(env._handle, 0)
        ep.steps_many(stream, T)

    g_steps_many = capture(steps_only_many, torch)
    def bench_step():
        g_episode.replay()                                   # reset + 30 step launches + episode statistics, one graph

    def reduce_batch():
        if world > 1:
            dist.all_reduce(stats_acc)                       # the single collective of the data path
        return stats_acc

    for _ in range(max(args.warmup, 3)):
        bench_step()
    reduce_batch()                                           # also initialises the NCCL communicator outside the timed region
    stats_acc.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                      # before the barrier: spawning nvidia-smi must not skew the ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        bench_step()
    final_stats = reduce_batch()                             # inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    if os.environ.get("IMX_BENCH_DEBUG"):
        print(f"[rank {rank}] elapsed {elapsed * 1e3:.3f} ms for {args.steps} steps", file=sys.stderr, flush=True)
    # keep the GPU busy a little longer so that the 100 ms clock sampler sees the kernel under load
    t_end = time.perf_counter() + 0.6
    while rank == 0 and time.perf_counter() < t_end:
        g_steps.replay()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed = float(tmax.item())
    agent_steps_per_bench_step = world * N * m * T
    value = agent_steps_per_bench_step * args.steps / elapsed

    # ---- roofline of the step kernel (live, CUDA events, steps-only graph: 30 dependent launches) ----
    B = algorithmic_bytes_per_env_step(env)
    reps = max(10, args.steps)
    timed_replays(g_steps, 3, torch)
    dt = timed_replays(g_steps, reps, torch) / (reps * T)
    achieved = B * N / dt / 1e9
    # the same 30 periods as ONE imx_step_many launch (stored action plan; the tiles' state stays in shared memory):
    # state moves once per launch instead of once per period, so its algorithmic bytes per env-step are lower
    timed_replays(g_steps_many, 3, torch)
    dt_many = timed_replays(g_steps_many, reps, torch) / reps
    S_words, R_rows = env.state_words, len(env._retailers)
    B_many = 4 * R_rows + 8 * m * (2 + O) + 2 * 4 * S_words / T
    replay_fused = {"agent_steps_per_sec": world * N * m * T / dt_many, "ms_per_30_periods": dt_many * 1e3,
                    "algorithmic_bytes_per_env_step": B_many, "achieved": B_many * N * T / dt_many / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": B_many * N * T / dt_many / 1e9 / peak,
                    "api": "imx_step_many(K=30): one launch per episode, state resident in shared memory, inputs prefetched two periods "
                           "ahead, outputs streamed behind the compute; bit-identical to 30 imx_step calls (tests/test_gpu_step_many.py)"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000_GBs": achieved / 8000.0,
                "traffic": NCU_TRAFFIC_BYTES_65536 if N == ENVS_PER_GPU else None,
                "traffic_source": "profiles/r1_ncu_step_kernel_final.txt (ncu --set full: dram__bytes_read+write per launch; "
                                  "the 31 MB working set of one launch is L2-resident, hence traffic << algorithmic bytes)", "kernel": KERNEL_VARIANTS[env._lib.imx_kernel_variant(env._handle)], "us_per_launch": dt * 1e6,
                "algorithmic_bytes_per_env_step": B, "envs_per_launch": N, "peak_source": peak_src,
                "note": "per-launch time = steps-only graph of 30 dependent launches / 30 (includes inter-kernel gaps)"}

    # ---- same kernel at large N (working set >> L2), rank 0 only -------------------------------------
    roof_large = None
    if rank == 0 and not args.skip_large:
        try:
            roof_large = large_n_roofline(torch, dev, peak)
        except Exception as exc:                             # never lose the headline line to the side measurement
            roof_large = {"error": str(exc)[:200]}

    # ---- end-to-end through the host-buffer ABI (H2D + kernel + D2H + sync per step) ----------------
    e2e = e2e_measure(env, demand_h, actions_h, torch, dev, world, episodes=max(2, min(args.steps, 5)))
    # same call with float32 observations (the cast RLlib's preprocessor applies to every observation anyway):
    # 40 % fewer bytes over PCIe; reported beside the float64 drop-in number, never instead of it
    env32 = MultiAgentInvManagement(dict(presets.serial4(), num_envs=N, device=str(dev), env_offset=rank * N, obs_dtype="float32"))
    e2e_f32 = e2e_measure(env32, demand_h, actions_h, torch, dev, world, episodes=max(2, min(args.steps, 5)))
    e2e_f32["api"] += "; obs_f32 = 1"
    del env32

    if rank == 0:
        line = {
            "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 state / f64 obs+reward",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": N, "periods_per_step": T, "agents": m,
                       "agent_steps_per_step": agent_steps_per_bench_step,
                       "l2": (f"inputs larger than L2: each step streams {T} distinct action/obs/reward buffers "
                              f"({(T * N * m * (2 + O) * 8) / 1e6:.0f} MB > 126 MB L2); the {env.state_words * 4 * N / 1e6:.1f} MB state stays cached by design"),
                       "timing": "CUDA events around K CUDA-graph replays (reset + 30 step launches) + per-episode return statistics"
                                 + (" + 1 NCCL all-reduce of the batch statistics" if world > 1 else "") + ", max over ranks"},
            "roofline": roofline, "roofline_large_n": roof_large, "replay_fused": replay_fused, "cpu_baseline": cpu, "cpu_baseline_1core": cpu_1, "cpu_baseline_c": cpu_c, "e2e": e2e, "e2e_f32_obs": e2e_f32,
            "gpu_launches": int(launches_per_episode * args.steps), "host_cores_bound_to_gpu_numa_node": numa_cores,
            "clocks": clocks,
            "episode_stats": {"n": float(final_stats[0].item()), "mean_return": float((final_stats[1] / final_stats[0]).item())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def large_n_roofline(torch, dev, peak, N=4 * 1024 * 1024, periods=8):
    from marl_for_im_b200 import presets
    from marl_for_im_b200.envs import MultiAgentInvManagement
    m, T = M_STAGES, T_PERIODS
    env = MultiAgentInvManagement(dict(presets.serial4(), num_envs=N, device=str(dev)))
    O = env.obs_len
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    demand = torch.randint(0, 11, (N, 1, T), dtype=torch.int32, device=dev, generator=g)
    actions = torch.rand((periods, N, m), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    obs = [torch.empty((N, m, O), dtype=torch.float64, device=dev) for _ in range(2)]
    rew = [torch.empty((N, m), dtype=torch.float64, device=dev) for _ in range(2)]
    obs0 = obs[0]
    ep = RawEpisode(env, demand, actions, obs, rew, obs0)
    s = torch.cuda.current_stream().cuda_stream
    ep.reset(s)
    ep.steps(s, periods)                                   # warm-up
    torch.cuda.synchronize()
    env._lib.imx_set_period(env._handle, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ep.steps(s, periods)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / periods
    B = algorithmic_bytes_per_env_step(env)
    achieved = B * N / dt / 1e9
    del env
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "envs_per_launch": N, "us_per_launch": dt * 1e6, "working_set_mb": B * N / 1e6,
            "agent_steps_per_sec": N * m / dt, "algorithmic_bytes_per_launch": B * N,
            "traffic": NCU_TRAFFIC_BYTES_4MI if N == 4 * 1024 * 1024 else None,
            "traffic_source": "profiles/r1_ncu_step_kernel_tma_specialised_4Mi_envs.txt (dram read 537 MB + write 1403 MB per launch = 0.97 x algorithmic)"}


def e2e_measure(env, demand_h, actions_h, torch, dev, world, episodes):
    """Host buffers in, host buffers out, every step: imx_reset_host + 30 x imx_step_host."""
    import torch.distributed as dist
    from marl_for_im_b200 import _lib
    N, m, T, O = env.num_envs, env.num_nodes, T_PERIODS, env.obs_len
    pin = lambda a: torch.as_tensor(a).pin_memory()   # noqa: E731
    dem_p, act_p = pin(demand_h), pin(actions_h)
    obs_bytes = 4 if env.obs_dtype == torch.float32 else 8
    obs_p = torch.empty((N, m, O), dtype=env.obs_dtype).pin_memory()
    rew_p = torch.empty((N, m), dtype=torch.float64).pin_memory()
    lib, h = env._lib, env._handle

    def one_episode():
        _lib.check(lib.imx_reset_host(h, C.c_void_p(dem_p.data_ptr()), None, 0, 7, C.c_void_p(obs_p.data_ptr())))
        for t in range(T):
            _lib.check(lib.imx_step_host(h, C.c_void_p(act_p[t].data_ptr()), C.c_void_p(obs_p.data_ptr()), C.c_void_p(rew_p.data_ptr())))
        return float(rew_p[0, 0])                      # device→host read of the step's result

    one_episode()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(episodes):
        one_episode()
    wall = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        wall = float(tmax.item())
    h2d = T * N * m * 8 + N * T * 4
    d2h = T * (N * m * O * obs_bytes + N * m * 8) + N * m * O * obs_bytes
    return {"value": world * N * m * T * episodes / wall, "unit": "agent-steps/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "episodes_timed": episodes,
            "api": "imx_reset_host + 30 x imx_step_host on pinned host buffers (wall clock, each call synchronises)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--large-only", action="store_true", help="only the 4 Mi-env roofline point (used for the ncu capture)")
    args = ap.parse_args()
    if args.large_only:
        import torch
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        print(json.dumps(large_n_roofline(torch, torch.device("cuda:0"), peak)), flush=True)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
