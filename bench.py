#!/usr/bin/env python
"""bench.py -- PerformanceTest-style throughput of the B200 rigid-body step (steps/s and body-steps/s).

  python bench.py --gpus N --steps K --warmup W [--workload pile|pyramid|convex_vs_mesh|max_bodies] [--bodies B]
  python bench.py --impl reference ...     # the reference's own CPU implementation (oracle/_ref) on the same config

A "step" is one PhysicsSystem::Update(1/60, 1) of the workload world, timed as PerformanceTest.cpp:380-391 does (Update only).
`value` = body-steps/s with the world resident in HBM (CUDA events on the library's stream); `e2e` = the same metric through
the reference-facing facade with HOST buffers every step (forces in, positions out). One world per GPU: under torchrun every
rank steps its own replica (a single world does not shard) and the values are summed (weak scaling).
"""
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

DT = 1.0 / 60.0
# SURVEY 8(d) byte constants
S_V = 24


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=120)
    p.add_argument("--warmup", type=int, default=120)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default=os.environ.get("B2J_BENCH_WORKLOAD", "pile"))
    p.add_argument("--bodies", type=int, default=int(os.environ.get("B2J_BENCH_BODIES", "1000000")))
    p.add_argument("--ref-bodies", type=int, default=100000, help="bodies of the bounded sample the reference arm / cpu_baseline steps")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget of the cpu_baseline leg")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def scene_params(args, bodies=None):
    b = bodies if bodies is not None else args.bodies
    if args.workload == "pile":
        return "pile", b, 15
    if args.workload == "max_bodies":
        return "max_bodies", b, 0
    if args.workload == "pyramid":
        return "pyramid", 15, 0
    if args.workload == "convex_vs_mesh":
        return "convex_vs_mesh", 10, 0
    raise SystemExit("unknown workload " + args.workload)


def config_of(args, num_dynamic, extra=None):
    cfg = {"workload": {"pile": "Pile (SURVEY 8d config 4: mixed sphere/box/capsule/12-point hull in a static box container)",
                        "pyramid": "PerformanceTest -s=Pyramid", "convex_vs_mesh": "PerformanceTest -s=ConvexVsMesh",
                        "max_bodies": "PerformanceTest -s=MaxBodies (N bodies)"}[args.workload],
           "bodies": num_dynamic, "dt": DT, "collision_steps": 1,
           "timing": "every step moves the whole world state through HBM (working set >> L2 for >= 1e5 bodies); no L2 flush between steps"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 6:
                    try:
                        self.samples.append((float(f[0]), float(f[1])))
                    except ValueError:
                        continue
                    for n, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples), "reasons": sorted(self.reasons)}


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref, FMA build, all host threads) on a bounded sample of the config."""
    import refharness as R
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not R.have_ref("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjoltref_fast.so missing"}))
        return
    bodies = min(args.bodies, args.ref_bodies)
    scene, p0, p1 = scene_params(args, bodies)
    ref = R.RefWorld(scene, p0, p1, variant="fast")
    ref.set_recording(False)
    threads = ref.L.jref_hardware_threads()
    nd = ref.num_dynamic
    # bound the run: time the first warm-up step, then scale steps so that the whole run stays within ~4 minutes
    t_first = ref.time_steps(1, DT, threads)
    budget = 200.0
    warm = max(3, min(args.warmup, int(0.4 * budget / max(t_first, 1e-4))))
    steps = max(1, min(args.steps, int(0.6 * budget / max(t_first, 1e-4))))
    ref.time_steps(warm - 1, DT, threads)
    t = ref.time_steps(steps, DT, threads)
    value = steps * nd / t
    line = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1000.0 * t / steps, "steps_per_sec": steps / t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, nd, {"sample": f"{nd} of {args.bodies} bodies" if nd != args.bodies else "full"}),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "reference",
                         "sample": f"{scene} with {nd} dynamic bodies, steps {warm}..{warm + steps} of the run, JobSystemThreadPool with {threads} threads, FMA build"},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline(args):
    """oracle/_ref timed on the host cores on a bounded sample (about args.cpu_seconds of CPU work)."""
    import refharness as R
    if not R.have_ref("fast"):
        return None
    bodies = min(args.bodies, args.ref_bodies)
    scene, p0, p1 = scene_params(args, bodies)
    ref = R.RefWorld(scene, p0, p1, variant="fast")
    ref.set_recording(False)
    threads = ref.L.jref_hardware_threads()
    nd = ref.num_dynamic
    t0 = time.time()
    steps, total = 0, 0.0
    warm = 0
    # skip the first steps (no contacts yet) up to a third of the budget, then time the rest
    while time.time() - t0 < args.cpu_seconds / 3 and warm < args.warmup:
        ref.time_steps(1, DT, threads)
        warm += 1
    while time.time() - t0 < args.cpu_seconds and steps < args.steps:
        total += ref.time_steps(1, DT, threads)
        steps += 1
    if steps == 0:
        return None
    return {"value": steps * nd / total, "unit": "body-steps/s", "cores": threads, "kind": "reference",
            "sample": f"{scene} with {nd} dynamic bodies (of {args.bodies}), steps {warm}..{warm + steps}, JobSystemThreadPool {threads} threads, FMA build, {total:.1f} s"}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import joltphysics_b200
    from joltphysics_b200 import _capi
    import facade as F

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libjolt_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    api = joltphysics_b200.load()
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
    os.environ["B2J_DEVICE"] = str(local_rank)
    scene, p0, p1 = scene_params(args)
    fs = F.FacadeScene(flib, scene, p0, p1)
    world = fs.world
    nd = fs.num_dynamic

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed); the per-step times are kept to compare with the CPU sample on the same early steps
    warm_ms = []
    for _ in range(args.warmup):
        _, st = world.step(DT)
        warm_ms.append(st.gpu_ms)

    # ---- device resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    gpu_ms, launches = 0.0, 0
    agg = {}
    for _ in range(args.steps):
        _, st = world.step(DT)
        gpu_ms += st.gpu_ms
        launches += st.kernel_launches
        for k in ("num_body_pairs", "num_pairs_from_cache", "num_manifolds", "num_contact_points", "num_constraints", "num_phases", "velocity_iterations", "position_iterations", "num_active_bodies"):
            agg[k] = agg.get(k, 0) + getattr(st, k)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.finish()

    # ---- per kernel device time (CUDA events around every launch on the library's stream) over extra steps that continue the
    # same run; kept out of the timed region above because the event records perturb back-to-back launches
    prof_steps = max(1, min(args.steps, 20))
    api.b2j_world_set_profiling(world.h, 1)
    prof_gpu_ms = 0.0
    pagg = {}
    for _ in range(prof_steps):
        _, st = world.step(DT)
        prof_gpu_ms += st.gpu_ms
        for k in ("num_contact_points", "num_constraints", "velocity_iterations"):
            pagg[k] = pagg.get(k, 0) + getattr(st, k)
    prof = world.profile()
    api.b2j_world_set_profiling(world.h, 0)

    # ---- end to end through the facade with host buffers (pinned): forces in, positions out, every step
    forces = torch.zeros((nd, 3), dtype=torch.float32).pin_memory().numpy()
    positions = torch.zeros((nd, 3), dtype=torch.float32).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 30))
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        fs.step_e2e(DT, forces, positions)
    barrier()
    e2e_wall = time.perf_counter() - t1
    nb = fs.num_bodies
    h2d = nd * (12 + 4)
    d2h = nb * (12 + 16 + 12 + 12 + 4)

    # ---- max over ranks, sum of work
    t_dev = gpu_ms / 1000.0
    vals = torch.tensor([t_dev, wall, e2e_wall, float(nd), float(launches)], dtype=torch.float64, device="cuda")
    if world_size > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_dev, wall, e2e_wall = mx[0].item(), mx[1].item(), mx[2].item()
        total_bodies, launches = sm[3].item(), int(sm[4].item())
    else:
        total_bodies = float(nd)
    if rank != 0:
        return

    K = args.steps
    value = K * total_bodies / t_dev
    # roofline of the dominant kernel (velocity solve), SURVEY 8(d) row (5): per constraint and iteration
    # C(c) + 4*S_v + 4*(3+c) algorithmic bytes, C(c) = 220 + 64 c
    M = pagg["num_constraints"] / prof_steps
    cbar = pagg["num_contact_points"] / max(pagg["num_constraints"], 1)
    V = pagg["velocity_iterations"] / prof_steps
    bytes_per_constraint_iter = (220 + 64 * cbar) + 4 * S_V + 4 * (3 + cbar)
    solve = prof.get("KSolveVelocity", {"ms": 0.0, "launches": 0})
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    roofline = None
    if solve["ms"] > 0:
        total_bytes = V * M * bytes_per_constraint_iter * prof_steps
        achieved = total_bytes / (solve["ms"] / 1000.0) / 1e9
        roofline = {"bound": "hbm", "kernel": "KSolveVelocity", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                    "bytes_per_launch": total_bytes / max(solve["launches"], 1), "avg_launch_us": 1000.0 * solve["ms"] / max(solve["launches"], 1),
                    "share_of_step": solve["ms"] / max(prof_gpu_ms, 1e-9), "measured_over": f"{prof_steps} profiled steps after the timed region"}
    line = {
        "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_dev / K, "steps_per_sec": K / t_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args, nd, {"parallelism": f"replicas x{world_size}" if world_size > 1 else "single world"}),
        "clocks": clocks,
        "e2e": {"value": e2e_steps * total_bodies / e2e_wall, "unit": "body-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": launches,
        "wall_ms_per_step": 1000.0 * wall / K,
        "roofline": roofline,
        "step_counters_mean": {k: v / K for k, v in agg.items()},
        "kernel_ms_per_step": {k: v["ms"] / prof_steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]},
        "profiled_ms_per_step": prof_gpu_ms / prof_steps,
    }
    if world_size == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args)
        if cb is not None:
            # GPU on the same early steps as the CPU sample (per step times recorded during warm-up)
            line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
