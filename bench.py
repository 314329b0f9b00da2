#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): Connector env-steps/s with the random agent at
10x10 / 5 agents, 65 536 envs per GPU (configs[1]), plus solved ParallelRandomWalk boards/s
as secondary lines.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle, all host threads)

A "step" is one VmapAutoResetWrapper(Connector).step over the whole batch with actions from the
random policy: random-action sampling + Connector.step + observation + auto-reset (regeneration
of finished envs with ParallelRandomWalkGenerator), i.e. one iteration of the reference's
`agent=random` benchmark loop (benchmark_on_random_agent.py:59-102).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers.  libgomp reads it when it is first loaded, and a
# later num_threads(n) clause then runs the CPU arm 8x SLOWER than a single thread (measured), so the
# value is replaced by the core count before anything that links OpenMP is imported.
if os.environ.get("OMP_NUM_THREADS") == "1" and "RANK" in os.environ:
    try:
        os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0))))
    except AttributeError:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

G, N = 10, 5
ENVS_PER_GPU = 65536
TIME_LIMIT = 50
CELLS = G * G
STATE_BYTES = 4 * CELLS + 4 + 28 * N + 8            # 552
TS_BYTES = 4 * N * CELLS + 5 * N + 4 + 8 * N + 1 + 12  # obs + mask + step_count + reward/discount + step_type + extras = 2082
STEP_BYTES = STATE_BYTES + 4 * N + STATE_BYTES + TS_BYTES  # 3206 B per env-step
PRW_BOARD_BYTES = {(10, 5): 488, (20, 10): 1768, (32, 16): 4360}
METRIC = "connector_env_steps_per_sec"
UNIT = "env-steps/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build_library()
    if world > 1:
        dist.barrier()
    import routing_board_generation_b200 as rbg
    from routing_board_generation_b200 import engine, sharding

    lib = rbg._lib.load()
    dev = torch.device("cuda", local)
    B = args.envs
    total = B * world  # weak scaling: 65 536 envs per GPU
    keys = sharding.shard_keys(rbg.PRNGKey(0), total, rank, world)
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(G, N), time_limit=TIME_LIMIT))
    state, ts = env.reset(keys)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # The timed loop drives the engine the way the reference's training / benchmark loop does: a
    # scan of `n_steps` = 20 env steps per call (configs/env/connector.yaml:27), here one
    # rbg_connector_rollout_random call per chunk writing stacked [20, B, ...] TimeSteps.
    chunk = args.chunk
    ts = engine.alloc_timestep(B, G, N, chunk)
    act = torch.empty((chunk, B, N), dtype=torch.int32, device=dev)

    def run_steps(k):
        nonlocal state
        while k > 0:
            n = min(k, chunk)
            engine.connector_rollout_random(state, n, TIME_LIMIT, -0.03, 0.1, autoreset_kind="parallel_random_walk", out=ts, actions=act)
            k -= n

    # Burn-in to the stationary regime: right after reset() every env is at step 0, so the
    # first episodes end in lock-step (all survivors hit time_limit together); a few episode
    # lengths later terminations are spread evenly (about 4 % of the envs per step).
    run_steps(args.burnin)
    run_steps(max(args.warmup, 3))
    sync_all()

    # ---- timed region: exactly K steps, device-timed, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rbg.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    run_steps(args.steps)
    e1.record()
    sync_all()
    launches = rbg.launch_count()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = total * args.steps / (ms_max / 1e3)

    # ---- the dominant kernel alone: event pairs recorded inside the library around every launch,
    # over a second pass of the same K steps
    rbg._lib.kernel_timing(True)
    for k in ("env", "prw", "rollout"):
        rbg._lib.kernel_time(k)
    run_steps(args.steps)
    torch.cuda.synchronize()
    n_env, ms_env = rbg._lib.kernel_time("env")
    n_prw, ms_prw = rbg._lib.kernel_time("prw")
    n_ro, ms_ro = rbg._lib.kernel_time("rollout")
    rbg._lib.kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed region and this identical second pass
    peak, peak_src = _peaks()
    if n_ro:  # fused path: rollout_warp_kernel covers `lib_chunk` steps per launch
        # (with kernel timing on the library launches the whole batch per kernel, one kernel at a time;
        # the timed region above runs two half-batch slices on two streams so that they overlap)
        steps_per_launch = args.steps / n_ro
        # State read once + written once per launch; per step only the emitted TimeStep and the actions
        alg = B * (2 * STATE_BYTES + steps_per_launch * (TS_BYTES + 4 * N))
        k_ms = ms_ro / n_ro
        achieved = alg / (k_ms / 1e3) / 1e9
        step_alg = B * (TS_BYTES + 4 * N + 2 * STATE_BYTES / steps_per_launch)  # algorithmic bytes of one whole step
        step_gbs = step_alg / (ms_max / args.steps / 1e3) / 1e9
        roofline = {"kernel": "rollout_warp_kernel", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": _traffic_from_profile("rollout_warp_kernel"), "algorithmic_bytes_per_launch": int(alg), "steps_per_launch": steps_per_launch,
                    "avg_launch_ms": round(k_ms, 5), "peak_source": peak_src, "kernel_share_of_step": round(ms_ro / max(ms_ro + ms_prw + ms_env, 1e-9), 4),
                    "refill_kernel_avg_ms": round(ms_prw / max(n_prw, 1), 5),
                    "whole_step": {"achieved": round(step_gbs, 1), "frac": round(step_gbs / peak, 4),
                                   "note": "algorithmic bytes of a step / the timed region's ms_per_step: rollout and cache-refill kernels of two half-batch slices overlapping on two streams"}}
    else:
        env_ms = ms_env / max(n_env, 1)
        achieved = STEP_BYTES * B / (env_ms / 1e3) / 1e9
        roofline = {"kernel": "env_warp_kernel", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": _traffic_from_profile("env_kernel"),
                    "algorithmic_bytes_per_launch": STEP_BYTES * B, "avg_launch_ms": round(env_ms, 5), "peak_source": peak_src,
                    "kernel_share_of_step": round(ms_env / max(ms_env + ms_prw, 1e-9), 4), "prw_reset_kernel_avg_ms": round(ms_prw / max(n_prw, 1), 5)}

    # ---- e2e: the env-step call with HOST buffers through the C-ABI (rbg_connector_step_host_io):
    # actions H2D from pinned memory, step, whole TimeStep D2H, every step.
    e2e = None
    secondary = []
    cpu_baseline = None
    if not args.skip_e2e:
        e2e = _e2e_host(args, rbg, lib, state, B, world, rank, dev)
    if rank == 0 and world == 1 and not args.skip_secondary:
        secondary = _secondary_prw(args, rbg, peak)
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_baseline = _cpu_baseline(budget_s=args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_max / args.steps, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic (keys = split(PRNGKey(0), B), random-policy actions)",
            "config": {"workload": "connector_step_random_agent_autoreset_prw", "grid": G, "agents": N, "envs_per_gpu": B, "envs_total": total, "time_limit": TIME_LIMIT,
                       "generator": "ParallelRandomWalkGenerator", "parallelism": f"env-sharded x{world}", "burnin_steps": args.burnin,
                       "api": f"rbg_connector_rollout_random, {chunk} steps per call (the reference's n_steps scan), State in place, stacked TimeSteps and actions written every step",
                       "l2": f"no flush: every step writes {(TS_BYTES + 4 * N) * B / 1e6:.0f} MB per GPU into a fresh slice of the stacked [{chunk}, B, ...] outputs ({(TS_BYTES + 4 * N) * B * chunk / 1e9:.1f} GB per call), far beyond the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _traffic_from_profile(kernel: str):
    """dram bytes per launch from the committed ncu --set full capture, if one exists (profiles/*.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def _e2e_host(args, rbg, lib, state, B, world, rank, dev):
    """The env step as a host-side consumer sees it (rbg_connector_step_host_io): every step the
    actions are copied H2D from pinned memory and the whole TimeStep (observation, mask, reward,
    discount, step_type, extras) is copied back D2H; the State is the env's own device-resident
    state, exactly as in the reference's loop."""
    import numpy as np
    import torch
    import torch.distributed as dist

    L = rbg._lib

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()

    hts = dict(obs=pinned((B, N, G, G), torch.int32), mask=pinned((B, N, 5), torch.uint8), sc=pinned((B,), torch.int32), reward=pinned((B, N), torch.float32), discount=pinned((B, N), torch.float32),
               step_type=pinned((B,), torch.int8), nc=pinned((B,), torch.int32), rc=pinned((B,), torch.float32), tpl=pinned((B,), torch.int32))
    act = pinned((B, N), torch.int32)
    rng = np.random.default_rng(rank)
    act.copy_(torch.from_numpy(rng.integers(0, 5, size=(B, N)).astype(np.int32)))
    a = state.agents
    s = L.rbg_state(state.grid.data_ptr(), state.step_count.data_ptr(), a.id.data_ptr(), a.start.data_ptr(), a.target.data_ptr(), a.position.data_ptr(), state.key.data_ptr())
    t = L.rbg_timestep(*(hts[k].data_ptr() for k in ("obs", "mask", "sc", "reward", "discount", "step_type", "nc", "rc", "tpl")))
    params = L.rbg_env_params(TIME_LIMIT, -0.03, 0.1, 0)
    h2d = act.numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in hts.values())

    def step():
        L.check(lib.rbg_connector_step_host_io(C.byref(s), act.data_ptr(), B, G, N, C.byref(params), C.byref(t), -1))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k = max(3, min(args.steps, 30))
    t0 = time.perf_counter()
    for _ in range(k):
        step()  # synchronous: returns after the D2H copies have landed
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    # context: what the bus gives a plain pinned D2H copy of the observation buffer alone
    dobs = torch.empty((B, N, G, G), dtype=torch.int32, device=dev)
    hts["obs"].copy_(dobs, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        hts["obs"].copy_(dobs, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    bus = dobs.numel() * 4 * 5 / (c0.elapsed_time(c1) / 1e3) / 1e9
    mine = d2h * k / dt / 1e9
    return {"value": round(B * world * k / dt, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": k,
            "api": "rbg_connector_step_host_io (pinned host actions in, full TimeStep out to pinned host memory, State device-resident, auto-reset on)",
            "timer": "host wall clock around synchronous calls, max over ranks",
            "d2h_gbs": round(mine, 1), "pinned_d2h_copy_gbs": round(bus, 1), "frac_of_bus": round(mine / bus, 3)}


def _secondary_prw(args, rbg, peak):
    """Solved boards/s of ParallelRandomWalkBoard.generate_board (the other half of BASELINE's metric)."""
    import torch

    out = []
    for (g, n, b) in ((10, 5, 65536), (20, 10, 131072), (32, 16, 32768)):
        keys = rbg.split(rbg.PRNGKey(0), b)
        board = rbg.ParallelRandomWalkBoard(g, g, n)
        for _ in range(3):
            board.generate_board(keys)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            board.generate_board(keys)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_ = PRW_BOARD_BYTES[(g, n)] * b
        out.append({"metric": "prw_solved_boards_per_sec", "workload": f"ParallelRandomWalkBoard.generate_board {g}x{g}/{n} B={b}", "value": round(b / (ms / 1e3), 1), "unit": "boards/s", "ms_per_batch": round(ms, 4),
                    "output_gbs": round(bytes_ / (ms / 1e3) / 1e9, 2), "hbm_frac": round(bytes_ / (ms / 1e3) / 1e9 / peak, 5), "bound": "integer issue (threefry2x32), see DESIGN.md"})
    # the same headline workload through the per-step API (one rbg_connector_step_random call per env step:
    # env_warp_kernel + reset kernel + side-stream cache refill), for callers that cannot use the fused rollout
    g, n, b = G, N, 65536
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(g, n), time_limit=TIME_LIMIT))
    st, _ = env.reset(rbg.split(rbg.PRNGKey(0), b))
    ts1 = rbg.engine.alloc_timestep(b, g, n)
    for _ in range(160):
        st, _, _ = rbg.engine.connector_step(st, None, TIME_LIMIT, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        st, _, _ = rbg.engine.connector_step(st, None, TIME_LIMIT, -0.03, 0.1, autoreset_kind="parallel_random_walk", inplace=True, random_policy=True, out=ts1, owner=env)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.append({"metric": "connector_env_steps_per_sec", "workload": f"per-step API (one Python call = one rbg_connector_step_random per env step) {g}x{g}/{n} B={b}", "value": round(b / (ms / 1e3), 1), "unit": "env-steps/s",
                "ms_per_step": round(ms, 4), "output_gbs": round(STEP_BYTES * b / (ms / 1e3) / 1e9, 1), "hbm_frac": round(STEP_BYTES * b / (ms / 1e3) / 1e9 / peak, 4)})
    # fused generate + reset + rollout at the A2C rollout shape (BASELINE configs[4]): 32x32 / 16 agents, 20 steps
    g, n, b, T = 32, 16, 8192, 20
    env = rbg.VmapAutoResetWrapper(rbg.Connector(generator=rbg.ParallelRandomWalkGenerator(g, n), time_limit=TIME_LIMIT))
    st, _ = env.reset(rbg.split(rbg.PRNGKey(0), b))
    ts_big = rbg.engine.alloc_timestep(b, g, n, T)
    for _ in range(6):  # past the first episodes
        st, _, _ = env.rollout_random(st, T, out=ts_big)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        st, _, _ = env.rollout_random(st, T, out=ts_big)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    step_bytes = 4 * n * g * g + 5 * n + 4 + 8 * n + 1 + 12 + 4 * n  # what a fused step must emit (65 825 B)
    out.append({"metric": "connector_env_steps_per_sec", "workload": f"fused generate+reset+rollout {g}x{g}/{n} B={b}, {T} steps per call", "value": round(b * T / (ms / 1e3), 1), "unit": "env-steps/s",
                "ms_per_step": round(ms / T, 4), "output_gbs": round(step_bytes * b * T / (ms / 1e3) / 1e9, 1), "hbm_frac": round(step_bytes * b * T / (ms / 1e3) / 1e9 / peak, 4)})
    del ts_big, st
    torch.cuda.empty_cache()
    # SeedExtension 14x14/7 (BASELINE configs[3]): generation + on-device validity of every board
    g, n, b = 14, 7, 65536
    keys = rbg.split(rbg.PRNGKey(0), b)
    board = rbg.SeedExtensionBoard(g, g, n)
    for _ in range(2):
        solved = board.return_solved_board(keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        solved = board.return_solved_board(keys)
        flags = rbg.engine.validate(solved, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.append({"metric": "seedext_solved_boards_per_sec", "workload": f"SeedExtensionBoard.return_solved_board {g}x{g}/{n} B={b} + rbg_validate", "value": round(b / (ms / 1e3), 1), "unit": "boards/s",
                "ms_per_batch": round(ms, 4), "invalid_boards": int((flags != 0).sum()), "bound": "sequential threefry chain per board (one lane per board), see DESIGN.md"})
    return out


# ------------------------------------------------------------------ CPU arm
def _cpu_workload(orc, B, nthreads):
    """reset B envs with the oracle and return a stepping closure (random policy + auto-reset step)."""
    kref = orc.split(orc.PRNGKey(0), B)
    st, _ = orc.connector_reset_batch("parallel_random_walk", kref, G, N, nthreads=nthreads)
    box = {"st": st, "ts": None}

    def step():
        act = orc.random_actions_batch(box["st"], nthreads=nthreads)
        # state updated in place, timestep buffers reused: no allocation inside the timed loop
        box["st"], box["ts"] = orc.connector_step_batch(box["st"], act, time_limit=TIME_LIMIT, autoreset_kind="parallel_random_walk", nthreads=nthreads, inplace=True, out=box["ts"])
        return box["ts"]

    return step


def _host_threads() -> int:
    """Every core this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    to its workers, which would quietly turn the CPU arm into a single-thread run at --gpus > 1."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        n = max(1, os.cpu_count() or 1)
    return n


def _cpu_baseline(budget_s: float = 15.0):
    cores = _host_threads()
    from oracle import oracle as orc

    orc.build()
    B = 16384
    step = _cpu_workload(orc, B, cores)
    for _ in range(2):
        step()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": round(B * n / dt, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} auto-reset random-agent steps over {B} envs 10x10/5 (same workload, smaller batch), OpenMP over envs, {dt:.1f} s"}


def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is pure Python on JAX and
    neither jax nor jumanji exist in this image, so this arm times the C oracle port
    with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _host_threads()
    from oracle import oracle as orc

    orc.build()
    B = args.envs
    step = _cpu_workload(orc, B, cores)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic (keys = split(PRNGKey(0), B), random-policy actions)",
        "config": {"workload": "connector_step_random_agent_autoreset_prw", "grid": G, "agents": N, "envs_per_gpu": B, "envs_total": B, "time_limit": TIME_LIMIT, "generator": "ParallelRandomWalkGenerator",
                   "parallelism": f"OpenMP x{cores} host threads"},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{args.steps} steps over {B} envs (rank 0 only; the CPU arm does not scale with --gpus)"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--burnin", type=int, default=160, help="untimed steps after reset() so that episode ends are desynchronised")
    ap.add_argument("--chunk", type=int, default=20, help="env steps per rollout call (the reference's n_steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
