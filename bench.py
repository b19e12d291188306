#!/usr/bin/env python
"""bench.py -- car-ticks/sec of the batched Car::step hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one process per GPU (torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation on host cores

Workload: N=1 -> BASELINE.json configs[1]: 4096 demo-car envs (ks_toyota_ae86_drift on driftplayground) on one
B200, randomised synthetic controls resampled every 33 ticks, env-style auto-reset.  N>1 -> configs[2]:
65536 envs per GPU, sharded by global env id, replicated track BVH, no per-tick collective (weak scaling).
A "step" is 33 physics ticks (1/333 s each) of every env = one control period.

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = through the public
API with host buffers (pinned H2D of actions and D2H of obs/reward/done every tick, as a user of the env does),
roofline for the dominant kernel (k_tick), cpu_baseline = the oracle timed on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 1.0 / 333.0
TICKS_PER_STEP = 33
METRIC = "car_ticks_per_sec"
UNIT = "car-ticks/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                parts = [x.strip() for x in out.split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0])); self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set(); self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (reference Car/Sim/Core sources + restated ODE) on host cores
# ----------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one host core: `sims` reference simulators stepped `ticks` ticks with the bench workload.  The whole loop
    (controls, Simulator::step, the env's reset rule) runs inside oracle/_ref/libpdref.so (pdref_bench_loop): the figure carries
    no Python / ctypes time."""
    wid, sims, ticks, seed, preroll = args
    import ctypes
    pdref = _oracle()
    L = pdref.lib()
    L.pdref_bench_loop.restype = ctypes.c_double
    L.pdref_bench_loop.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_void_p]
    S = [pdref.RefSim() for _ in range(sims)]
    for k, s in enumerate(S):
        s.set_collision_response(False)          # the env terminates on hit (terminate_on_hit): detection only, as our arm runs it
        s.teleport_spline(((wid * 7919 + k * 104729) % 1000) / 1000.0)
    arr = (ctypes.c_void_p * sims)(*[s.h for s in S])
    resets = ctypes.c_longlong(0)
    secs = L.pdref_bench_loop(arr, sims, ticks, preroll, seed + wid, ctypes.byref(resets))
    return sims * ticks, secs


def _physical_cores():
    try:
        sibs = set()
        for c in os.sched_getaffinity(0):
            with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c) as f:
                sibs.add(f.read().strip())
        return len(sibs) or None
    except Exception:
        return None


def _oracle():
    """ctypes wrapper of oracle/_ref/libpdref.so -- the ONLY oracle use of this file: the reference arm / cpu_baseline."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pdref", os.path.join(ROOT, "oracle", "pdref.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def run_reference_cpu(total_sims, ticks, cores, preroll=666):
    import multiprocessing as mp
    per = max(1, total_sims // cores)
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(w, per, ticks, 1234, preroll) for w in range(cores)])
        wall = time.perf_counter() - t0
    units = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return units, slowest, wall


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pdref = _oracle()
    if not pdref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)      # one process per usable hardware thread
    physical = _physical_cores()
    sims_per_core = 2
    # warm-up + K steps, each step = 33 ticks of (cores * sims_per_core) reference simulators
    times = []
    units = 0
    for s in range(args.warmup + args.steps):
        u, slow, wall = run_reference_cpu(cores * sims_per_core, TICKS_PER_STEP * 4, cores)
        if s >= args.warmup:
            times.append(slow); units += u
    total = sum(times)
    value = units / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] sample: demo car on driftplayground, random controls resampled every 33 ticks, env auto-reset; "
                               "each step = %d ticks of %d reference simulators (one process per host core) after 666 untimed pre-roll ticks" % (TICKS_PER_STEP * 4, cores * sims_per_core)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "physical_cores": physical,
                         "sample": "%d sims x %d ticks per step (after 666 untimed pre-roll ticks), %d steps, one process per hardware thread (%d threads on %s physical cores), the whole loop inside the C++ library (no Python per tick); reference Car/Sim/Core sources (g++ -O2) + restated ODE 0.16.3 back-end, collision detection without response (the env terminates on hit)" % (cores * sims_per_core, TICKS_PER_STEP * 4, args.steps, cores, physical)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    from projectd_core_b200 import Batch
    from projectd_core_b200.assets import default_base
    from projectd_core_b200.env import configure_like_env

    from projectd_core_b200 import dist as pdist
    rank, local, world = pdist.init_from_env("nccl")
    dist = None
    if world > 1:
        import torch.distributed as dist
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local if world > 1 else 0)
    n_envs = args.envs if args.envs else (4096 if world == 1 else 65536)
    K, W = args.steps, args.warmup

    b = configure_like_env(Batch(default_base(), n_envs=n_envs, device=dev.index, synthetic_tris=args.synthetic_tris, car=args.car))
    env_offset, _ = pdist.shard_range(world * n_envs, rank, world)
    b.set_seed(1234, env_offset)             # RNG keyed by global env id: results do not depend on the sharding
    b.teleport_mode(2)                        # random start positions u ~ U[0,1)
    b.set_autoreset(1)                        # finished envs reset inside the next step's tick launch (gymnasium >= 1.0 VectorEnv convention)
    stream = torch.cuda.ExternalStream(b.stream(), device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(99 + rank)
    # synthetic controls for every step, resident in HBM before the timed region
    acts = (torch.rand((W + K, n_envs, 2), device=dev, generator=gen) * 2 - 1).contiguous()
    rew = torch.zeros(n_envs, device=dev); done = torch.zeros(n_envs, device=dev, dtype=torch.int32)
    obs = b.obs_tensor()
    torch.cuda.synchronize()

    def run_steps(lo, hi):
        for s in range(lo, hi):
            a = acts[s]
            for _ in range(TICKS_PER_STEP):
                b.env_step(a, DT, None, rew, done)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pre-roll (untimed set-up): the cars start at rest on the grid; the workload the metric is quoted on is the
    #      steady state of the rollout (cars at speed, episodes ending and resetting all the time, mean episode length
    #      ~1300 ticks), so the simulation is advanced past that transient before anything is timed ----
    pre = torch.rand((args.preroll // TICKS_PER_STEP + 1, n_envs, 2), device=dev, generator=gen) * 2 - 1
    for t in range(args.preroll):
        b.env_step(pre[t // TICKS_PER_STEP], DT, None, rew, done)
    b.env_stats(reset=True)
    del pre

    # ---- device-resident throughput ----
    run_steps(0, W)
    barrier()
    l0 = b.launch_count()
    sampler = ClockSampler(dev.index); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        run_steps(W, W + K)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = b.launch_count() - l0
    ms = pdist.max_over_ranks(ms, dev)
    ticks = K * TICKS_PER_STEP
    value = world * n_envs * ticks / (ms * 1e-3)

    # ---- dominant kernel (the tick kernel): CUDA events on the batch's stream between consecutive launches of the
    #      same env-step workload (one launch per env step), continuing the rollout ----
    NK = 2 * TICKS_PER_STEP
    with torch.cuda.stream(stream):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(NK + 1)]
        a = acts[W + K - 1]
        evs[0].record(stream)
        for i in range(NK):
            b.env_step(a, DT, None, rew, done)
            evs[i + 1].record(stream)
    torch.cuda.synchronize()
    tick_ms = sum(evs[i].elapsed_time(evs[i + 1]) for i in range(NK)) / NK
    words = b.words
    alg_bytes = n_envs * (2 * words * 4 + 8 + 96 + 8)      # state read + written once, controls in, obs + reward/flags out
    peak, peak_kind = _peaks()
    achieved = alg_bytes / (tick_ms * 1e-3) / 1e9
    # issue-side view (not tensor work): audited ~50 kflop per car-tick against 74 TFLOP/s fp32
    flop_frac = (n_envs / (tick_ms * 1e-3)) * 50e3 / 74e12
    kernel = b.tick_kernel()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("%s@%d" % (kernel, n_envs))      # dram bytes of one launch of this kernel at this batch size, from the committed ncu capture
    except Exception:
        pass

    # ---- end to end through the public API with host buffers ----
    h_act = torch.empty((n_envs, 2), dtype=torch.float32).pin_memory()
    h_obs = torch.empty((n_envs, 24), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n_envs, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n_envs, dtype=torch.int32).pin_memory()
    d_act = torch.empty((n_envs, 2), device=dev)
    acts_host = acts[W:W + min(K, 8)].cpu()
    e2e_steps = min(K, 8)

    def e2e_tick(a_host):
        # the call a host-side user makes: actions in (pinned) host memory in, obs / reward / done in host memory out
        h_act.copy_(a_host)
        b.env_step_host(h_act, DT, h_obs, h_rew, h_done)
        return float(h_rew[0])

    for _ in range(3):
        e2e_tick(acts_host[0])
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        for _ in range(TICKS_PER_STEP):
            e2e_tick(acts_host[s])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()                   # sampled across the three timed regions (device-resident steps, per-launch events, end to end)
    e2e_s = pdist.max_over_ranks(e2e_s, dev)
    e2e_value = world * n_envs * e2e_steps * TICKS_PER_STEP / e2e_s

    # ---- episode statistics: the only collective of the path (once per rollout, never per tick) ----
    stats = pdist.reduce_stats(b.env_stats(reset=False), dev).tolist()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpdref.so")):
        # in a fresh interpreter (this process holds a CUDA context and pinned buffers; workers forked from it ran ~3x slower)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "12", "--warmup", "1"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600).stdout.strip().splitlines()
            ref = json.loads(out[-1])
            cpu = ref.get("cpu_baseline")
        except Exception as ex:  # the baseline is a reported figure, not part of the measurement
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference", "sample": "failed: %s" % ex}

    # ---- like-for-like anchor of the weak-scaling curve (N=1 default run only): the per-GPU load of configs[2], 65536 envs,
    #      on this one GPU, same pre-roll / warm-up / event timing as above, in this process ----
    anchor = None
    if world == 1 and n_envs == 4096 and not args.no_anchor and not args.synthetic_tris:
        anchor = _anchor_65536(args, dev, gen)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("configs[1]: 4096 demo-car envs on 1 B200" if (world == 1 and n_envs == 4096) else "%d demo-car envs per GPU (configs[2] uses 65536)" % n_envs)
                       + (", ks_toyota_ae86_drift on driftplayground" if not args.synthetic_tris else ", ks_toyota_ae86_drift on a generated 20.8 km circuit of ~%d triangles (configs[3])" % args.synthetic_tris) + ", random controls resampled every 33 ticks, env auto-reset (next-step convention: a finished env resets inside the next step's launch); "
                         "measured in the rollout's steady state after %d untimed pre-roll ticks" % args.preroll,
                       "envs_per_gpu": n_envs, "ticks_per_step": TICKS_PER_STEP, "dt": DT,
                       "l2": ("no flush: the working set is the env state (%.1f MB), read and written every tick; " % (n_envs * words * 4 / 1e6))
                             + ("it is larger than L2 (126 MB)" if n_envs * words * 4 > 126e6 else "it fits L2 and staying L2-resident between consecutive ticks IS the workload (a simulation steps the same state), see DESIGN.md"),
                       "parallelism": "env-sharded x%d, no per-tick collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": TICKS_PER_STEP * n_envs * 8, "d2h_bytes_per_step": TICKS_PER_STEP * n_envs * (96 + 4 + 4),
                    "note": "per tick one pd_env_step_host call with pinned host buffers: the tick kernel reads the actions from and writes obs / reward / done to host memory directly (zero-copy over PCIe), then one stream sync"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_kind, "algorithmic_bytes_per_car_tick": alg_bytes / n_envs, "kernel_ms": tick_ms,
                         "fp32_issue_frac_of_74TFLOPs_at_50kflop_per_car_tick": flop_frac},
            "cpu_baseline": cpu,
            "extras": {"weak_scaling_anchor": anchor},
            "episode_stats": {"episodes": stats[0], "mean_return": (stats[1] / stats[0]) if stats[0] else None, "mean_length": (stats[2] / stats[0]) if stats[0] else None,
                              "collisions": stats[3], "offtrack": stats[4], "stuck": stats[5], "lowreward": stats[6], "nan": stats[7]},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    # orderly teardown (no os._exit: the driver's exit hook records which native libraries this process loaded): drop every
    # torch handle that aliases the batch's buffers, release the batch, then let the interpreter exit normally
    del obs, evs, e0, e1, stream
    b.sync(); b.close()
    torch.cuda.synchronize()


def _anchor_65536(args, dev, gen, n_envs=65536):
    """Device-resident car-ticks/s of ONE GPU at the per-GPU load the N>1 runs use (65536 envs): the like-for-like first
    point of the weak-scaling curve, measured by the same rules as the headline (pre-roll, >= 3 warm-up steps, CUDA events
    on the batch's stream; the state, 175 MB, is larger than L2)."""
    import torch
    from projectd_core_b200 import Batch
    from projectd_core_b200.assets import default_base
    from projectd_core_b200.env import configure_like_env
    K, W = min(args.steps, 10), 3
    b = configure_like_env(Batch(default_base(), n_envs=n_envs, device=dev.index, car=args.car))
    b.set_seed(1234, 0); b.teleport_mode(2); b.set_autoreset(1)
    stream = torch.cuda.ExternalStream(b.stream(), device=dev)
    acts = (torch.rand((args.preroll // TICKS_PER_STEP + 1 + W + K, n_envs, 2), device=dev, generator=gen) * 2 - 1).contiguous()
    rew = torch.zeros(n_envs, device=dev); done = torch.zeros(n_envs, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    for t in range(args.preroll):
        b.env_step(acts[t // TICKS_PER_STEP], DT, None, rew, done)
    base = args.preroll // TICKS_PER_STEP + 1
    for s in range(W):
        for _ in range(TICKS_PER_STEP):
            b.env_step(acts[base + s], DT, None, rew, done)
    torch.cuda.synchronize()
    l0 = b.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for s in range(W, W + K):
            for _ in range(TICKS_PER_STEP):
                b.env_step(acts[base + s], DT, None, rew, done)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"envs_per_gpu": n_envs, "value": n_envs * K * TICKS_PER_STEP / (ms * 1e-3), "unit": UNIT, "steps": K, "warmup": W,
           "ms_per_step": ms / K, "gpu_launches": b.launch_count() - l0, "kernel": b.tick_kernel(),
           "note": "one GPU at the per-GPU load of configs[2]; divide the N-GPU value by N times this for the like-for-like weak-scaling efficiency"}
    del e0, e1, stream
    b.sync(); b.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: 4096 at N=1, 65536 at N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-anchor", action="store_true", help="skip the 65536-env single-GPU anchor the default N=1 run adds under extras")
    ap.add_argument("--car", default="ks_toyota_ae86_drift", help="car model (BASELINE's configs all use the demo car; the other four bundled cars run the double-wishbone kernel instances)")
    ap.add_argument("--synthetic-tris", type=int, default=0, help="BASELINE configs[3]: generated 20.8 km circuit of about this many triangles instead of driftplayground")
    ap.add_argument("--preroll", type=int, default=1998, help="untimed ticks before the warm-up (brings the rollout to its steady state)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
