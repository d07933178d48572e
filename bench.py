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
SLOTS = 8  # distinct action vectors / rollout-storage slots cycled through (footprint > L2)

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
def cpu_reference(family, seconds_target=12.0, envs=1 << 16):
    """Times the oracle port (numpy restatement of the reference step) on every available host core."""
    cores = len(os.sched_getaffinity(0))
    # calibrate: single short run to size the sample
    t = time.monotonic()
    out = subprocess.run([sys.executable, "-m", "oracle.cpu_bench", "--family", family, "--envs", str(envs), "--steps", "10",
                          "--warmup", "3"], capture_output=True, text=True, cwd=ROOT)
    if out.returncode != 0:
        raise RuntimeError("oracle.cpu_bench failed: " + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    per_step = (r["t1"] - r["t0"]) / r["steps"]
    startup = time.monotonic() - t
    steps = max(20, int(seconds_target / max(per_step, 1e-6) / 1.3))  # 1.3: all-core runs are slower than single
    start_at = time.monotonic() + startup + 3.0
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.cpu_bench", "--family", family, "--envs", str(envs), "--steps",
                               str(steps), "--warmup", "10", "--seed", str(i), "--start-at", repr(start_at)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT) for i in range(cores)]
    res = []
    for p in procs:
        so, se = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("oracle.cpu_bench worker failed: " + se[-2000:])
        res.append(json.loads(so.strip().splitlines()[-1]))
    wall = max(x["t1"] for x in res) - min(x["t0"] for x in res)
    total = sum(x["envs"] * x["steps"] for x in res)
    single = envs / per_step
    return {"value": total / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} processes x {envs} envs x {steps} steps of the numpy oracle port (oracle/), "
                      f"episode phases de-synchronised; single-core rate {single:.3e}",
            "single_core_value": single, "wall_s": wall}


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
    env = make_env(args.workload, b, rank)
    cap = env.capacity
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    if wl["n_act"]:
        actions = torch.randint(0, wl["n_act"], (SLOTS, cap), dtype=torch.int8, device=dev, generator=gen)
    else:
        cols = wl.get("act_cols", 2)
        actions = torch.rand((SLOTS, cap, cols) if cols > 1 else (SLOTS, cap), device=dev, generator=gen) * 2 - 1
    # rollout storage: outputs of step t go to slot t % SLOTS (like an RL rollout buffer); with the
    # action slots this makes the per-step footprint rotate through > L2-size memory
    out = {}
    for name in ("obs", "reward", "terminated", "truncated"):
        a = env._arrays[name]
        out[name] = torch.zeros((SLOTS,) + tuple(a.shape), dtype=a.dtype, device=dev)
    env.reset(seed=0)
    # de-synchronise episode phases (after a synchronised reset every env would truncate on the same step)
    env._arrays["elapsed"][:b] = torch.randint(0, env.time_limit + 1, (b,), device=dev, generator=gen, dtype=torch.int32)

    def run_steps(k):
        done = 0
        while done < k:
            n = min(SLOTS, k - done)
            env.step_many(actions[:n], out)
            done += n

    sampler = ClockSampler(local_rank)
    sampler.start()
    # warm-up: W steps, then a fixed ~1 s of load so clocks are sampled under the same kernel
    run_steps(max(3, args.warmup))
    torch.cuda.synchronize()
    sampler.active.set()
    t_end = time.monotonic() + (0.0 if args.quick else 1.0)
    while time.monotonic() < t_end:
        run_steps(SLOTS * 16)
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, barrier + sync on both sides
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = env.launch_count
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- when step_many ran as fused launches, also time one launch per step (what env.step() costs)
    single = None
    if launches < args.steps:
        env.set_fused_steps(False)
        run_steps(SLOTS * 4)
        torch.cuda.synchronize()
        k1 = min(args.steps, 1000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(k1)
        e1.record()
        torch.cuda.synchronize()
        ts = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        single = (k1, float(ts.item()))
        env.set_fused_steps(True)

    # ---- end-to-end: public API with HOST buffers (pinned), H2D + step + D2H inside the timed region
    if args.e2e_steps is not None:
        e2e_steps = max(1, args.e2e_steps)
    else:
        e2e_steps = 1 if args.quick else max(3, min(args.steps, 100))
    host_actions = env.pinned_actions(SLOTS)          # the steps' inputs live in pinned host memory
    hrng = np.random.default_rng(99 + rank)
    if wl["n_act"]:
        host_actions[:] = hrng.integers(0, wl["n_act"], size=(SLOTS, b)).astype(np.int8)
    else:
        host_actions[:] = hrng.uniform(-1, 1, size=host_actions.shape).astype(np.float32)
    for i in range(3):
        env.step_host(host_actions[i % SLOTS])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        obs, rew, term, trunc, _ = env.step_host(host_actions[i % SLOTS])   # H2D actions, step, D2H results
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d, d2h = env.host_bytes_per_step()
    sampler.active.clear()
    sampler.stop()

    # episode statistics all-reduce (the only collective on this path; logging cadence, outside the timed region)
    if world > 1:
        st = env.stats_tensor()
        dist.all_reduce(st)

    if rank == 0:
        total_envs = b * world
        value = total_envs * args.steps / (ms_max * 1e-3)
        peak, peak_src = measured_peak_gbs()
        # Fused multi-step launches (gpt_step_many keeps the state in registers for T steps): the state bytes move
        # once per LAUNCH, so the algorithmic bytes are restated downward — never count bytes that are not moved.
        steps_per_launch = args.steps / max(launches, 1)
        alg_bytes = wl["alg_bytes"]
        if steps_per_launch > 1.0:
            alg_bytes = wl["alg_bytes"] - wl.get("state_bytes", 0) + wl.get("state_bytes", 0) / steps_per_launch
        per_launch_s = ms_max * 1e-3 / max(launches, 1)
        achieved = alg_bytes * steps_per_launch * cap / per_launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "envs_per_gpu": b, "envs_total": total_envs,
                       "rng": "philox4x32-10", "l2": f"inputs rotate over {SLOTS} action slots and outputs over {SLOTS} "
                       "rollout slots (footprint > 126 MB L2); " + ("the state arrays are re-read every step" if steps_per_launch <= 1.0 else
                       f"fused launches of {steps_per_launch:g} steps: state read and written once per launch, actions read and outputs written every step"),
                       "episode_phases": "de-synchronised (elapsed ~ U[0,time_limit]) before warm-up"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "alg_bytes_per_env_step": alg_bytes,
                         "steps_per_launch": steps_per_launch, "kernel_us": per_launch_s * 1e6, "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": total_envs * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": e2e_steps, "api": "env.step_host(numpy) -> gpt_step_host",
                    "host_numa": numa},
            "gpu_launches": launches * world,
            "clocks": sampler.summary(),
        }
        if single is not None:
            k1, ms1 = single
            ach1 = wl["alg_bytes"] * cap / (ms1 * 1e-3 / k1) / 1e9
            line["single_step_launches"] = {"value": total_envs * k1 / (ms1 * 1e-3), "unit": UNIT, "steps": k1, "kernel_us": ms1 * 1e3 / k1,
                                            "alg_bytes_per_env_step": wl["alg_bytes"], "achieved": ach1, "frac": ach1 / peak,
                                            "note": "same workload with one launch per step (gpt_step): state read and written every step"}
        traffic_file = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
        if os.path.exists(traffic_file):
            try:
                line["roofline"]["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu and not args.quick:
            line["cpu_baseline"] = cpu_reference(wl["cpu_family"], seconds_target=args.cpu_seconds)
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure Python
    and cannot travel to the GPU box, so this times the oracle port (oracle/, a numpy restatement pinned
    bit-exactly to the reference by tests/golden) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    # K "steps" each a bounded sample; the run as a whole is bounded to args.cpu_seconds
    cb = cpu_reference(wl["cpu_family"], seconds_target=args.cpu_seconds)
    b = 1 << args.log2_envs
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "envs_per_gpu": b, "envs_total": b * world,
                   "note": "CPU arm: bounded sample of the same workload on all host cores"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="taxi", choices=sorted(WORKLOADS))
    ap.add_argument("--log2-envs", type=int, default=22, help="log2 of envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: no load phase, no e2e, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
