#!/usr/bin/env python
"""bench.py — env-steps/s of the fused B200 env step (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload taxi] [--log2-envs 22]

A "step" is one pass of the hot path (one fused kernel launch: transition + reward + done + autoreset +
obs) over one batch of B envs per GPU with synthetic uniform-random actions.  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gym-po-taxi_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC, UNIT = "env_steps_per_sec", "env-steps/s"
# action / rollout-storage slots cycled through = upper bound of the steps fused into one launch (footprint >> L2).  20 = the
# driver's K: its 20 timed steps are ONE rollout launch (state read and written once per 20 steps).  Round 2 measured
# T = 10 -> 20 -> 40 at 482 -> 523 -> 543 G env-steps/s with the same fraction of the HBM peak on the restated bytes.
MAX_SLOTS = 20
# configurations also timed (briefly) after the headline so that the driver-run line carries them
EXTRA_WORKLOADS = ("rooms_hansen8", "rooms_grid5", "rooms_grid9", "crooms", "tag", "msrooms", "rooms_grid5_l32", "rooms_hansen8_l32")

# algorithmic bytes per env-step (DESIGN.md "Algorithmic bytes")
WORKLOADS = {
    "taxi": dict(alg_bytes=29, state_bytes=18, n_act=5, dtype="int32", cpu_family="taxi",
                 desc="Taxi POMDP 5x5 (4 locations, time_limit 200), fused step+obs+autoreset, Philox RNG, uniform random actions"),
    "taxi_hansen": dict(alg_bytes=29, state_bytes=18, n_act=5, dtype="int32", cpu_family="taxi",
                        desc="Hansen-obs Taxi 5x5, fused step+obs+autoreset, Philox RNG"),
    "rooms_hansen8": dict(alg_bytes=23, state_bytes=12, n_act=8, dtype="int32", cpu_family="rooms_hansen8",
                          desc="FourRooms '4' discrete, hansen8 obs, 0.2 action-slip, fixed goal, Philox RNG"),
    "rooms_grid5": dict(alg_bytes=19 + 25, state_bytes=12, n_act=8, dtype="u8", cpu_family="rooms_grid5",
                        desc="FourRooms '4', 5x5 egocentric window obs, 0.2 action-slip, fixed goal"),
    "crooms": dict(alg_bytes=46, n_act=0, dtype="f32", cpu_family="crooms",
                   desc="continuous ROOMS '4' (float32 fast mode, Gaussian action noise 0.2, wall rejection), vector_mdp obs, yx float32 actions"),
    "tag": dict(alg_bytes=62, n_act=0, dtype="f32", cpu_family="tag",
                desc="point-mass Tag (AntTag pursuit rules, float32 fast mode), yx float32 actions"),
    "crooms_f64": dict(alg_bytes=70, n_act=0, dtype="f64", cpu_family="crooms",
                       desc="continuous ROOMS '4' (float64 parity mode), vector_mdp obs, yx float32 actions"),
    "tag_f64": dict(alg_bytes=102, n_act=0, dtype="f64", cpu_family="tag",
                    desc="point-mass Tag (float64 parity mode), yx float32 actions"),
    "car": dict(alg_bytes=44, n_act=0, act_cols=1, dtype="f32", cpu_family="car",
                desc="car-flag (heaven/hell with priest), float32 forces; obs is the live float32 state row"),
    "msrooms": dict(alg_bytes=23, state_bytes=12, n_act=4, dtype="int32", cpu_family="msrooms",
                    desc="multistory FourRooms (3 floors, stairs), mdp obs, 1/3 action-slip, cardinal actions, fixed goal, Philox RNG"),
    "rooms_grid9": dict(alg_bytes=19 + 81, state_bytes=12, n_act=8, dtype="u8", cpu_family="rooms_grid9",
                        desc="FourRooms '4', 9x9 egocentric window obs, 0.2 action-slip, fixed goal"),
    "rooms_grid5_l32": dict(alg_bytes=19 + 25, state_bytes=12, n_act=8, dtype="u8", cpu_family="rooms_grid5",
                            desc="ROOMS layout '32' (25x49 map, the large-map case), 5x5 egocentric window obs, 0.2 action-slip, fixed goal"),
    "rooms_hansen8_l32": dict(alg_bytes=23, state_bytes=12, n_act=8, dtype="int32", cpu_family="rooms_hansen8",
                              desc="ROOMS layout '32' (25x49 map), hansen8 obs, 0.2 action-slip, fixed goal"),
}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the GPU is under load."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
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
        self.smi = None
        if not self.ok:   # fall back to the nvidia-smi query of the profiling recipe
            try:
                self.smi = subprocess.Popen(
                    ["nvidia-smi", f"--id={index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.smi = None

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.smi.stdout:
            if self._stop_evt.is_set():
                break
            if not self.active.is_set():
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(int(float(f[0])))
                self.max_mhz = int(float(f[1]))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
                self.ok = True
            except Exception:
                pass

    def run(self):
        if not self.ok:
            if self.smi is not None:
                self._run_smi()
            return
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.smi is not None:
            try:
                self.smi.terminate()
            except Exception:
                pass

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def cpu_reference(family, total_envs, *, steps=None, seconds_target=12.0, warmup=3, max_seconds=150.0):
    """Times the reference's CPU vectorized step on every available host core: one process per core (numpy
    elementwise ops are single-threaded), each with total_envs/cores envs, started together.  The UNMODIFIED
    reference (baseline/_ref, kind "reference") when it is installed and implements the family, else the numpy
    oracle port (kind "port").  `steps` given: exactly that many steps of the whole batch (capped so that the run
    stays under max_seconds); else as many as fit seconds_target."""
    sys.path.insert(0, ROOT)
    from oracle import cpu_bench
    kind = "reference" if (cpu_bench.reference_installed() and family in cpu_bench.REFERENCE_FAMILIES) else "port"
    cores = len(os.sched_getaffinity(0))
    per = max(1, -(-total_envs // cores))
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_bench", "--impl", kind, "--family", family, "--envs", str(per),
                               "--warmup", str(max(1, warmup)), "--seed", str(i), "--sync-stdin"],
                              stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
             for i in range(cores)]
    per_step = []
    for p in procs:
        line = p.stdout.readline()
        if not line.startswith("READY"):
            raise RuntimeError("oracle.cpu_bench worker failed: " + p.stderr.read()[-2000:])
        per_step.append(float(line.split()[1]))
    est = max(per_step)   # s per step of the whole batch with all cores busy (the warm-ups ran concurrently)
    if steps is None:
        n = max(3, int(seconds_target / max(est, 1e-6)))
    else:
        n = max(1, min(steps, int(max_seconds / max(est, 1e-6))))
    start_at = time.monotonic() + 0.3
    for p in procs:
        p.stdin.write(f"GO {start_at!r} {n}\n")
        p.stdin.flush()
    res = []
    for p in procs:
        so, se = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("oracle.cpu_bench worker failed: " + se[-2000:])
        res.append(json.loads(so.strip().splitlines()[-1]))
    wall = max(x["t1"] for x in res) - min(x["t0"] for x in res)
    total = sum(x["envs"] * x["steps"] for x in res)
    what = ("the unmodified reference (baseline/_ref/gym_po via oracle.ref_loader)" if kind == "reference"
            else "the numpy oracle port (oracle/)")
    return {"value": total / wall, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{cores} processes x {per} envs (= {per * cores} envs, the whole batch) x {n} steps of {what}, "
                      f"episode phases de-synchronised",
            "envs": per * cores, "steps": n, "ms_per_step": wall / n * 1e3, "wall_s": wall}


# ------------------------------------------------------------------------------------------------
def make_env(workload, b, rank, seed=0):
    from gym_po.envs import TaxiVecEnv
    if workload == "taxi":
        return TaxiVecEnv(b, seed=seed, env_offset=rank * b)
    if workload == "taxi_hansen":
        return TaxiVecEnv(b, hansen_obs=True, seed=seed, env_offset=rank * b)
    from gym_po.envs import RoomsEnv
    if workload == "rooms_hansen8":
        return RoomsEnv(b, "4", obs_type="hansen8", seed=seed, env_offset=rank * b)
    if workload == "rooms_grid5":
        return RoomsEnv(b, "4", obs_type="grid", obs_n=5, seed=seed, env_offset=rank * b)
    if workload == "rooms_grid9":
        return RoomsEnv(b, "4", obs_type="grid", obs_n=9, seed=seed, env_offset=rank * b)
    if workload == "rooms_grid5_l32":
        return RoomsEnv(b, "32", obs_type="grid", obs_n=5, seed=seed, env_offset=rank * b)
    if workload == "rooms_hansen8_l32":
        return RoomsEnv(b, "32", obs_type="hansen8", seed=seed, env_offset=rank * b)
    if workload in ("crooms", "crooms_f64"):
        from gym_po.envs import CRoomsEnv
        return CRoomsEnv(b, "4", obs_type="vector_mdp", seed=seed, env_offset=rank * b,
                         precision="float64" if workload.endswith("f64") else "float32")
    if workload == "msrooms":
        from gym_po.envs import MultistoryFourRoomsEnv
        return MultistoryFourRoomsEnv(b, grid_z=3, seed=seed, env_offset=rank * b)
    if workload == "car":
        from gym_po.envs import CarVecEnv
        return CarVecEnv(b, seed=seed, env_offset=rank * b)
    if workload in ("tag", "tag_f64"):
        from gym_po.envs import TagVecEnv
        return TagVecEnv(b, seed=seed, env_offset=rank * b, precision="float64" if workload.endswith("f64") else "float32")
    raise KeyError(workload)


def steps_per_launch_for(k):
    """Steps fused into one launch: the largest divisor of K in [4, MAX_SLOTS], so that the K timed steps are whole
    launches of equal length (no ragged 8+8+4); K without such a divisor falls back to launches of up to 8."""
    for t in range(min(k, MAX_SLOTS), 3, -1):
        if k % t == 0:
            return t
    return min(8, k)


class Workload:
    """One env family at B envs per GPU with its synthetic action slots and rollout storage."""

    def __init__(self, name, b, rank, dev, k):
        import torch
        self.name, self.b, self.wl = name, b, WORKLOADS[name]
        self.env = make_env(name, b, rank)
        self.cap = self.env.capacity
        self.slots = steps_per_launch_for(k)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        wl = self.wl
        if wl["n_act"]:
            self.actions = torch.randint(0, wl["n_act"], (self.slots, self.cap), dtype=torch.int8, device=dev, generator=gen)
        else:
            cols = wl.get("act_cols", 2)
            self.actions = torch.rand((self.slots, self.cap, cols) if cols > 1 else (self.slots, self.cap), device=dev, generator=gen) * 2 - 1
        # rollout storage: outputs of step t go to slot t % slots (like an RL rollout buffer); with the action
        # slots this makes the per-step footprint rotate through > L2-size memory
        self.out = {}
        for nm in ("obs", "reward", "terminated", "truncated"):
            a = self.env._arrays[nm]
            self.out[nm] = torch.zeros((self.slots,) + tuple(a.shape), dtype=a.dtype, device=dev)
        self.env.reset(seed=0)
        # de-synchronise episode phases (after a synchronised reset every env would truncate on the same step)
        self.env._arrays["elapsed"][:b] = torch.randint(0, self.env.time_limit + 1, (b,), device=dev, generator=gen, dtype=torch.int32)
        self.footprint = sum(t.numel() * t.element_size() for t in self.out.values()) + self.actions.numel() * self.actions.element_size()

    def run_steps(self, k):
        done = 0
        while done < k:
            n = min(self.slots, k - done)
            self.env.step_many(self.actions[:n], self.out)
            done += n

    def close(self):
        self.env.close()
        self.out = self.actions = self.env = None


def timed_blocks(w, k, world, dev, min_total_ms, min_repeats=7, max_repeats=400):
    """R >= 7 back-to-back repeats of the K-step block, each between its own pair of CUDA events on the launching
    stream (torch's current stream); barrier + synchronize on both sides of the timed region; one UNTIMED block
    between the barrier and the first event (the first launches after an idle wait run at ramping clocks).  Returns
    the per-repeat times in ms (max over ranks) and the launches of one block."""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    w.run_steps(k)
    e1.record()
    torch.cuda.synchronize()
    est = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)   # every rank must run the same number of repeats
    reps = int(min(max_repeats, max(min_repeats, -(-min_total_ms // max(float(est.item()), 1e-3)))))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w.run_steps(k)                     # untimed
    l0 = w.env.launch_count
    evs[0].record()
    for r in range(reps):
        w.run_steps(k)                 # exactly K steps
        evs[r + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = (w.env.launch_count - l0) // reps
    ms = torch.tensor([evs[r].elapsed_time(evs[r + 1]) for r in range(reps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return [float(x) for x in ms.tolist()], launches


def pcie_ceiling(env, host_actions, b, world, dev, reps):
    """Copy-only lower bound of one end-to-end step (see the call site).  Max over ranks."""
    import torch
    import torch.distributed as dist
    hn = env._ensure_host()
    host = env._host                                    # pinned torch tensors behind step_host's numpy views
    names = ("obs", "reward", "terminated", "truncated")
    d_out = [env._arrays[n][:b] for n in names]
    d_act = torch.empty_like(host["actions"], device=dev)
    h_act = torch.from_numpy(host_actions[0])           # pinned (pinned_actions)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def step():
        with torch.cuda.stream(s1):
            d_act.copy_(h_act, non_blocking=True)
        with torch.cuda.stream(s2):
            for n, d in zip(names, d_out):
                host[n].copy_(d, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    def timed(fn, n):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_s = timed(step, reps)
    nbytes = sum(host[n].numel() * host[n].element_size() for n in names)
    big_h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    big_d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d2h_s = timed(lambda: (big_h.copy_(big_d, non_blocking=True), torch.cuda.current_stream().synchronize()), max(3, reps // 2))
    h2d_s = timed(lambda: (big_d.copy_(big_h, non_blocking=True), torch.cuda.current_stream().synchronize()), max(3, reps // 2))
    return {"step_s": step_s, "raw_d2h_gbs": nbytes / d2h_s / 1e9, "raw_h2d_gbs": nbytes / h2d_s / 1e9}


def device_numbers(w, k, ms_list, launches, world, peak):
    """env-steps/s and the roofline figures of one timed leg (median repeat)."""
    wl = w.wl
    ms = statistics.median(ms_list)
    spl = k / max(launches, 1)
    # Fused multi-step launches keep the state in registers for T steps: the state bytes move once per LAUNCH, so
    # the algorithmic bytes are restated downward — never count bytes that are not moved.
    alg = wl["alg_bytes"]
    if spl > 1.0:
        alg = wl["alg_bytes"] - wl.get("state_bytes", 0) + wl.get("state_bytes", 0) / spl
    per_launch_s = ms * 1e-3 / max(launches, 1)
    achieved = alg * spl * w.cap / per_launch_s / 1e9
    return {"value": w.b * world * k / (ms * 1e-3), "ms_per_step": ms / k, "ms_per_step_best": min(ms_list) / k,
            "ms_per_step_worst": max(ms_list) / k, "repeats": len(ms_list), "launches": launches, "steps_per_launch": spl,
            "alg_bytes_per_env_step": alg, "kernel_us": per_launch_s * 1e6, "achieved": achieved, "frac": achieved / peak}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout must carry exactly one JSON line, but libraries print there too (NCCL's version banner at init):
    # point fd 1 at stderr while the run is in progress and restore it for the final line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gym_po.sharding import bind_to_gpu_numa_node
    # multi-GPU runs: pin each rank next to its GPU before any pinned host buffer is allocated (single-GPU runs keep
    # the whole cpuset: the CPU baseline of the same run uses every host core)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "not bound (single process)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    b = 1 << args.log2_envs
    k = args.steps
    peak, peak_src = measured_peak_gbs()
    w = Workload(args.workload, b, rank, dev, k)
    env, cap = w.env, w.cap

    sampler = ClockSampler(local_rank)
    sampler.start()
    # warm-up: W steps, then a fixed ~1 s of load so clocks are sampled under the same kernel
    w.run_steps(max(3, args.warmup))
    torch.cuda.synchronize()
    sampler.active.set()
    t_end = time.monotonic() + (0.0 if args.quick else 1.0)
    while time.monotonic() < t_end:
        w.run_steps(w.slots * 16)
        torch.cuda.synchronize()

    # ---- timed region (see timed_blocks): repeats of exactly K steps; the median repeat is the reported one
    ms_list, launches = timed_blocks(w, k, world, dev, args.min_timed_ms)
    head = device_numbers(w, k, ms_list, launches, world, peak)

    # ---- when step_many ran as fused launches, also time one launch per step (what env.step() costs)
    single = None
    if launches < k:
        env.set_fused_steps(False)
        w.run_steps(w.slots * 4)
        ms1, l1 = timed_blocks(w, k, world, dev, args.min_timed_ms)
        single = device_numbers(w, k, ms1, l1, world, peak)
        env.set_fused_steps(True)

    # ---- end-to-end: public API with HOST buffers (pinned), H2D + step + D2H inside the timed region
    host_actions = env.pinned_actions(w.slots)          # the steps' inputs live in pinned host memory
    hrng = np.random.default_rng(99 + rank)
    if wl["n_act"]:
        host_actions[:] = hrng.integers(0, wl["n_act"], size=(w.slots, b)).astype(np.int8)
    else:
        host_actions[:] = hrng.uniform(-1, 1, size=host_actions.shape).astype(np.float32)
    e2e_k = args.e2e_steps if args.e2e_steps is not None else min(k, 100)   # a host-path step takes ~9 ms at 2^22 envs
    e2e_reps = 1 if args.quick else 7
    for i in range(3):
        env.step_host(host_actions[i % w.slots])
    e2e_times = []
    for r in range(e2e_reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_k):
            obs, rew, term, trunc, _ = env.step_host(host_actions[i % w.slots])   # H2D actions, step, D2H results
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    te = torch.tensor(e2e_times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = statistics.median(te.tolist())
    h2d, d2h = env.host_bytes_per_step()
    # ---- what the box's PCIe / host memory gives for the SAME bytes with no kernel in between: all ranks at once, the
    # step's H2D (actions) and D2H (obs, reward, terminated, truncated) copies on two streams, synchronised per step
    # like step_host; and one large contiguous copy each way (raw link rate)
    pcie = pcie_ceiling(env, host_actions, b, world, dev, max(5, min(e2e_k, 50)))
    clocks_head = sampler.summary()
    slots, footprint = w.slots, w.footprint
    # episode statistics all-reduce (the only collective on this path; logging cadence, outside the timed region)
    if world > 1:
        st = env.stats_tensor()
        dist.all_reduce(st)
    w.close()
    del w, env
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, driver-visible (short: repeats until >= 10 ms each)
    extra = {}
    if not args.quick and not args.no_workloads:
        if world == 1:
            todo = [(n, args.log2_envs) for n in EXTRA_WORKLOADS if n != args.workload] + [("rooms_hansen8", args.log2_envs - 1)]
        else:   # configs[2]: FourRooms hansen8 sharded over the GPUs, 2^21 envs per GPU (2^24 in total on 8)
            todo = [("rooms_hansen8", 21)]
        for name, lg in todo:
            i0 = len(sampler.samples)
            x = Workload(name, 1 << lg, rank, dev, k)
            x.run_steps(x.slots * 8)
            torch.cuda.synchronize()
            t_end = time.monotonic() + 0.25
            while time.monotonic() < t_end:
                x.run_steps(x.slots * 8)
                torch.cuda.synchronize()
            msx, lx = timed_blocks(x, k, world, dev, 10.0)
            d = device_numbers(x, k, msx, lx, world, peak)
            smp = sampler.samples[i0:]
            extra[f"{name}_2p{lg}"] = {
                "value": d["value"], "unit": UNIT, "envs_per_gpu": 1 << lg, "envs_total": (1 << lg) * world,
                "kernel_us": d["kernel_us"], "steps_per_launch": d["steps_per_launch"],
                "alg_bytes_per_env_step": d["alg_bytes_per_env_step"], "achieved": d["achieved"], "frac": d["frac"],
                "repeats": d["repeats"], "ms_per_step": d["ms_per_step"], "dtype": x.wl["dtype"], "desc": x.wl["desc"],
                "clocks": {"sm_mhz": statistics.median(smp) if smp else None, "samples": len(smp)}}
            x.close()
            del x
            torch.cuda.empty_cache()
    sampler.active.clear()
    sampler.stop()
    clocks_all = sampler.summary()

    if rank == 0:
        total_envs = b * world
        spl = head["steps_per_launch"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": k, "warmup": max(3, args.warmup),
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "envs_per_gpu": b, "envs_total": total_envs,
                       "rng": "philox4x32-7",
                       "action_stream": "open loop: pre-generated synthetic action slots resident in HBM (north_star: 'synthetic action streams')"
                                        + (f"; gpt_step_many fuses {spl:g} consecutive steps per launch — the closed-loop "
                                           "one-launch-per-step rate of env.step() is reported as single_step_launches" if spl > 1 else ""),
                       "l2": f"inputs rotate over {slots} action slots and outputs over {slots} rollout slots "
                             f"(footprint {footprint / 1e6:.0f} MB per GPU > 126 MB L2); "
                             + ("the state arrays are re-read every step" if spl <= 1.0 else
                                f"state read and written once per launch of {spl:g} steps, actions read and outputs written every step"),
                       "timing": f"median of {head['repeats']} back-to-back repeats of the {k}-step block, each between its own CUDA "
                                 "events (max over ranks per repeat); barrier + synchronize on both sides; one untimed block after the barrier",
                       "episode_phases": "de-synchronised (elapsed ~ U[0,time_limit]) before warm-up"},
            "roofline": {"bound": "hbm", "achieved": head["achieved"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                         "traffic": None, "peak_source": peak_src, "alg_bytes_per_env_step": head["alg_bytes_per_env_step"],
                         "steps_per_launch": spl, "kernel_us": head["kernel_us"], "frac_of_nominal_8TBs": head["achieved"] / 8000.0,
                         "frac_best_repeat": head["frac"] * head["ms_per_step"] / head["ms_per_step_best"],
                         "frac_worst_repeat": head["frac"] * head["ms_per_step"] / head["ms_per_step_worst"]},
            "repeats": head["repeats"], "ms_per_step_best": head["ms_per_step_best"], "ms_per_step_worst": head["ms_per_step_worst"],
            "e2e": {"value": total_envs * e2e_k / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": e2e_k, "repeats": e2e_reps,
                    "api": "env.step_host(numpy) -> gpt_step_host", "host_numa": numa,
                    "pcie_gbs_achieved": (h2d + d2h) * world * e2e_k / e2e_s / 1e9,
                    "pcie_bound": total_envs / pcie["step_s"], "frac_of_pcie": (total_envs * e2e_k / e2e_s) / (total_envs / pcie["step_s"]),
                    "pcie_bound_note": "env-steps/s if a step were ONLY its host<->device copies (same pinned buffers, same sizes, all ranks at "
                                       "once, two streams, synchronised per step); raw = one contiguous copy of the same total bytes each way",
                    "pcie_copy_gbs": (h2d + d2h) * world / pcie["step_s"] / 1e9, "pcie_raw_h2d_gbs": pcie["raw_h2d_gbs"] * world,
                    "pcie_raw_d2h_gbs": pcie["raw_d2h_gbs"] * world},
            "gpu_launches": launches * world,
            "clocks": clocks_head,
        }
        if single is not None:
            line["single_step_launches"] = {
                "value": single["value"], "unit": UNIT, "steps": k, "repeats": single["repeats"], "kernel_us": single["kernel_us"],
                "alg_bytes_per_env_step": single["alg_bytes_per_env_step"], "achieved": single["achieved"], "frac": single["frac"],
                "note": "same workload with one launch per step (gpt_step, what a closed-loop env.step() costs): state read and written every step"}
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture — only when the capture was
        # taken at this run's steps per launch (the bytes are per launch)
        traffic_file = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
        if os.path.exists(traffic_file):
            try:
                tr = json.load(open(traffic_file))
                if abs(float(tr.get("steps_per_launch", -1)) - spl) < 1e-9 and int(tr.get("envs", cap)) == cap:
                    line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
                    line["roofline"]["traffic_source"] = tr.get("source")
                else:
                    line["roofline"]["traffic_note"] = (f"committed capture is for {tr.get('steps_per_launch')} steps per launch, "
                                                        f"this run used {spl:g}: not comparable, left null")
            except Exception:
                pass
        if extra:
            line["workloads"] = extra
            line["clocks_all_workloads"] = clocks_all
        if world == 1 and not args.no_cpu and not args.quick:
            line["cpu_baseline"] = cpu_reference(wl["cpu_family"], total_envs, seconds_target=args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path — the unmodified package installed in
    baseline/_ref (kind "reference"; the numpy oracle port only if that directory is missing) — on all host cores:
    K steps of the same workload (the whole batch of envs_total envs, split over one process per core).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    b = 1 << args.log2_envs
    cb = cpu_reference(wl["cpu_family"], b * world, steps=args.steps, warmup=max(3, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": cb["steps"],
        "warmup": max(3, args.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "envs_per_gpu": b, "envs_total": b * world,
                   "note": "CPU arm: the same batch (envs_total envs) stepped by the reference's numpy code, one process per host core"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    global MAX_SLOTS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="taxi", choices=sorted(WORKLOADS))
    ap.add_argument("--log2-envs", type=int, default=22, help="log2 of envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--min-timed-ms", type=float, default=25.0, help="repeat the K-step block until at least this much GPU time (and >= 7 repeats)")
    ap.add_argument("--max-steps-per-launch", type=int, default=MAX_SLOTS, help="upper bound of the steps fused into one gpt_step_many launch (= rollout slots)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no load phase, one e2e repeat, no CPU baseline, no other workloads")
    args = ap.parse_args()
    MAX_SLOTS = max(1, args.max_steps_per_launch)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
