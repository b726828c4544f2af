#!/usr/bin/env python
"""bench.py -- PerformanceTest-style throughput of the B200 rigid-body step (steps/s and body-steps/s).

  python bench.py --gpus N --steps K --warmup W [--workload batch|pile|pyramid|convex_vs_mesh|max_bodies]
  python bench.py --impl reference ...     # the reference's own CPU implementation (oracle/_ref) on the same config

Default workload = BASELINE.json configs[4], the one the "1/2/4/8 B200" metric is quoted on: 4096 independent Pyramid worlds
(PerformanceTest -s=Pyramid, 1240 boxes each) batched RL-env style, sharded world_id -> rank over the N GPUs with no inter-GPU
traffic on the data path (only the final statistics are reduced over NCCL). A "step" advances EVERY world of the job by one
PhysicsSystem::Update(1/60, 1), timed as PerformanceTest.cpp:380-391 does (Update only); body-steps/s = steps/s x dynamic bodies.

Every leg measures the SAME simulation window, steps [W, W + K) counted from the creation state of the scene:
  value      device resident: CUDA events on the library's stream around b2j_step / b2j_batch_step
  e2e        through the C ABI / facade with HOST buffers every step (per body forces in, positions out); the worlds are reset
             to their creation state (b2j_batch_reset_worlds / a new scene) and warmed up again first
  roofline   per kernel CUDA-event timing (b2j_world_set_profiling) of the dominant kernel over the same window after another reset
  reference  (--impl reference, and the cpu_baseline leg) the unmodified reference on the host cores, same window

At N = 1 the line also carries `pile` (configs[3], one 1M body world) and `extra` (configs[0..2]: Pyramid, ConvexVsMesh, MaxBodies),
each with the reference timed beside it on the SAME body count and step window. Single world workloads (--workload pile ...) run one
replica per GPU ("replicas only").
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
S_V = 24  # bytes of (v, w) of one body
PYRAMID_BODIES = 1240           # PerformanceTest/PyramidScene.h:23-47, height 15
CONVEX_VS_MESH_BODIES = 1764    # PerformanceTest/ConvexVsMeshScene.h:28-117
MAX_BODIES_FULL = 8388608       # PerformanceTest/MaxBodiesScene.h:44-80 (BASELINE configs[2])
BATCH_PAIRS_PER_WORLD, BATCH_CONSTRAINTS_PER_WORLD = 16384, 12288  # SURVEY 8(d) config 5 limits
T_START = time.time()


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=60)
    p.add_argument("--warmup", type=int, default=60)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default=os.environ.get("B2J_BENCH_WORKLOAD", "batch"))
    p.add_argument("--worlds", type=int, default=int(os.environ.get("B2J_BENCH_WORLDS", "4096")), help="total worlds of the batch workload (sharded over the GPUs)")
    p.add_argument("--bodies", type=int, default=int(os.environ.get("B2J_BENCH_BODIES", "0")), help="bodies of the pile / max_bodies workloads (default: 1 000 000 / 8 388 608)")
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU work budget of the cpu_baseline leg of the main workload")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-pile", action="store_true", help="skip the single-world Pile-1M measurement at N=1")
    p.add_argument("--no-extras", action="store_true", help="skip the configs[0..2] measurements at N=1")
    p.add_argument("--budget-seconds", type=float, default=float(os.environ.get("B2J_BENCH_BUDGET", "480")), help="wall clock after which the remaining secondary measurements are skipped")
    a = p.parse_args()
    if a.bodies <= 0:
        a.bodies = MAX_BODIES_FULL if a.workload == "max_bodies" else 1000000
    return a


def scene_params(workload, bodies):
    if workload == "pile":
        return "pile", bodies, 15
    if workload == "max_bodies":
        return "max_bodies", bodies, 0
    if workload in ("pyramid", "batch"):
        return "pyramid", 15, 0
    if workload == "convex_vs_mesh":
        return "convex_vs_mesh", 10, 0
    raise SystemExit("unknown workload " + workload)


WORKLOAD_NAMES = {
    "batch": "4096 independent Pyramid worlds batched (BASELINE.json configs[4])",
    "pile": "Pile (BASELINE.json configs[3] / SURVEY 8d config 4: mixed sphere/box/capsule/12-point hull in a static box container, one world)",
    "pyramid": "PerformanceTest -s=Pyramid (configs[0])", "convex_vs_mesh": "PerformanceTest -s=ConvexVsMesh (configs[1])",
    "max_bodies": "PerformanceTest -s=MaxBodies (configs[2], N bodies)",
}


def job_bodies(args):
    """Dynamic bodies of ONE replica of the workload (the whole job for the batch)."""
    return {"batch": args.worlds * PYRAMID_BODIES, "pyramid": PYRAMID_BODIES, "convex_vs_mesh": CONVEX_VS_MESH_BODIES}.get(args.workload, args.bodies)


def job_config(args):
    """The `config` object: identical on the b200 arm and on the reference arm (it names the job, not how an arm runs it)."""
    batch = args.workload == "batch"
    return {"workload": WORKLOAD_NAMES[args.workload], "worlds": args.worlds if batch else 1, "bodies": int(job_bodies(args)), "dt": DT, "collision_steps": 1,
            "parallelism": (f"worlds sharded over {args.gpus} GPUs (world_id -> rank blocks), no data-path collective" if batch else f"one world per GPU x{args.gpus} (replicas only)"),
            "window": f"simulation steps [{args.warmup}, {args.warmup + args.steps}) from the creation state",
            "timing": "every step streams the whole job state through HBM (working set >> 126 MB L2); no L2 flush between steps"}


def over_budget(args):
    return time.time() - T_START > args.budget_seconds


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML in-process every 200 ms.
    (A `nvidia-smi -lms 200` child process, the recipe's form, was measured to slow the launch-heavy 1M body pile step from 30 ms
    to 96 ms on this box; the same NVML queries from a thread do not. nvidia-smi remains the fallback when pynvml is missing.)"""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.proc = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx))
                bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS:
                    if bits & bit:
                        self.reasons.add(name)
                time.sleep(0.2)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "1000"], stdout=subprocess.PIPE, text=True)
            names = [n for n, _ in self.REASONS]
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
        elif self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[1] for s in self.samples), "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle/_ref = the unmodified reference compiled with plain g++, FMA + LTO build)
# ---------------------------------------------------------------------------------------------------------------------

CPU_BUILD = "FMA + LTO build of the unmodified reference (oracle/_ref)"


def cpu_run(workload, bodies, worlds, warmup, steps, budget_s=None):
    """Times the reference on the host cores over simulation steps [warmup, warmup + steps) of the workload, same body count as the
    GPU arm. With a budget (cpu_baseline leg inside the GPU arm) the window is shortened -- never the body count -- and the sample
    string says which steps were timed. Returns (value, steps, warm, description, threads)."""
    import refharness as R
    L = R.ref_lib("fast")
    threads = L.jref_hardware_threads()
    if workload == "batch":
        # many small independent worlds: every host thread owns one world and steps it with a single threaded job system
        # (no cross thread synchronisation -- the best case for the CPU); throughput scales with the number of worlds in flight
        L.jref_time_worlds_parallel.restype = C.c_double
        L.jref_time_worlds_parallel.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        warm, n = warmup, steps
        if budget_s is not None:
            probe = L.jref_time_worlds_parallel(b"pyramid", 15, 0, threads, 1, 2, DT) / 2.0  # seconds per step with all threads busy
            total = max(2, int(budget_s / max(probe, 1e-4)))
            if warm + n > total:
                warm = min(warm, max(1, total // 3))
                n = max(1, min(n, total - warm))
        wall = L.jref_time_worlds_parallel(b"pyramid", 15, 0, threads, warm, n, DT)
        value = threads * n * PYRAMID_BODIES / wall
        return value, n, warm, f"{threads} concurrent Pyramid worlds (one per host thread, single threaded job system each) of the {worlds}, steps [{warm}, {warm + n})", threads
    scene, p0, p1 = scene_params(workload, bodies)
    ref = R.RefWorld(scene, p0, p1, variant="fast")
    ref.set_recording(False)
    nd = ref.num_dynamic
    warm, n = warmup, steps
    done = 0
    if budget_s is not None:
        t_first = ref.time_steps(1, DT, threads)
        done = 1
        total = max(2, int(budget_s / max(t_first, 1e-4)))
        if warm + n > total:
            warm = min(warm, max(1, total // 3))
            n = max(1, min(n, total - warm))
    if warm > done:
        ref.time_steps(warm - done, DT, threads)
    t = ref.time_steps(n, DT, threads)
    ref.close()
    return n * nd / t, n, warm, f"{scene} with {nd} dynamic bodies, steps [{warm}, {warm + n}), JobSystemThreadPool with {threads} threads", threads


def cpu_baseline_object(workload, bodies, worlds, warmup, steps, budget_s):
    try:
        value, n, warm, sample, threads = cpu_run(workload, bodies, worlds, warmup, steps, budget_s)
        return {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "reference", "steps": n, "warmup": warm, "sample": sample + "; " + CPU_BUILD}
    except Exception as e:  # the CPU leg must never take the GPU line down
        return {"error": str(e)}


def run_reference(args):
    import refharness as R
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not R.have_ref("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjoltref_fast.so missing"}))
        return
    t0 = time.time()
    value, steps, warm, sample, threads = cpu_run(args.workload, args.bodies, args.worlds, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": None, "higher_is_better": True, "scaling": "strong" if args.workload == "batch" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": job_config(args),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "reference", "sample": sample + "; " + CPU_BUILD},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------

def shard_worlds(total_worlds, rank, world_size):
    """world_id -> rank: contiguous blocks, the first (total % world_size) ranks hold one world more. Returns (first, count)."""
    base, rem = divmod(total_worlds, world_size)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def reduce_job(dist, world_size, times, work, device):
    """The job's only collective: MAX over ranks of the times, SUM of the work counters (tensors live on `device`)."""
    import torch
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()


class Workload:
    """One rank's share of a workload behind a uniform step / e2e / profile / reset interface."""

    def __init__(self, workload, bodies, worlds, api, flib, rank, world_size):
        import numpy as np
        from joltphysics_b200 import _capi
        self.api, self.flib, self.np, self._capi = api, flib, np, _capi
        self.kind, self.bodies = workload, bodies
        self.scene, self.batch = None, None
        self.first_world, self.n_worlds = shard_worlds(worlds, rank, world_size) if workload == "batch" else (0, 1)
        self._create()
        self.stats = _capi.StepStats()
        self._e2e_ready = False

    def _create(self):
        import facade as F
        scene, p0, p1 = scene_params(self.kind, self.bodies)
        self.scene = F.FacadeScene(self.flib, scene, p0, p1)
        if self.kind == "batch":
            self.batch = self.api.b2j_batch_create(self.scene.world.h, self.n_worlds, BATCH_PAIRS_PER_WORLD, BATCH_CONSTRAINTS_PER_WORLD)
            if not self.batch:
                raise SystemExit("b2j_batch_create failed: " + self.api.last_error())
            self.num_dynamic = self.n_worlds * self.scene.num_dynamic
            self.num_slots = self.n_worlds * self.scene.num_bodies
        else:
            self.num_dynamic = self.scene.num_dynamic
            self.num_slots = self.scene.num_bodies

    def close(self):
        """Frees the device memory of the workload (the batch holds > 100 GB at 4096 worlds)."""
        if self.batch:
            self.api.b2j_batch_destroy(self.batch)
            self.batch = None
        if self.scene is not None:
            self.scene.close()
            self.scene = None

    def reset(self):
        """Back to the creation state: the batch resets its worlds on the device (b2j_batch_reset_worlds), a single world is rebuilt."""
        if self.batch:
            ids = self.np.arange(self.n_worlds, dtype=self.np.uint32)
            if self.api.b2j_batch_reset_worlds(self.batch, ids.ctypes.data_as(C.POINTER(C.c_uint32)), self.n_worlds) != 0:
                raise SystemExit("b2j_batch_reset_worlds failed: " + self.api.last_error())
        else:
            self.close()
            self._create()
            self._e2e_ready = False

    def step(self):
        if self.batch:
            r = self.api.b2j_batch_step(self.batch, DT, 1, C.byref(self.stats))
        else:
            r = self.api.b2j_step(self.scene.world.h, DT, 1, C.byref(self.stats))
        if r < 0:
            raise SystemExit("step failed: " + self.api.last_error())
        return self.stats

    def set_profiling(self, on):
        if self.batch:
            self.api.b2j_batch_set_profiling(self.batch, on)
        else:
            self.api.b2j_world_set_profiling(self.scene.world.h, on)

    def profile(self):
        cap, stride = 128, 64
        names = C.create_string_buffer(cap * stride)
        ms, launches = (C.c_float * cap)(), (C.c_uint32 * cap)()
        if self.batch:
            n = self.api.b2j_batch_get_profile(self.batch, names, stride, ms, launches, cap)
        else:
            n = self.api.b2j_world_get_profile(self.scene.world.h, names, stride, ms, launches, cap)
        return {names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(): {"ms": float(ms[i]), "launches": int(launches[i])} for i in range(min(n, cap))}

    def e2e_buffers(self, torch):
        n = self.num_slots if self.batch else self.num_dynamic
        if not hasattr(self, "forces") or len(self.forces) != n:
            self.forces = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
            self.positions = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
        fp = C.POINTER(C.c_float)
        self._f = self.forces.ctypes.data_as(fp)
        self._st = self._capi.BodyState(self.positions.ctypes.data, None, None, None, None, None, None)
        self._e2e_ready = True
        if self.batch:
            return n * 12, self.num_slots * 12
        return self.num_dynamic * (12 + 4), self.scene.num_bodies * (12 + 16 + 12 + 12 + 4)

    def step_e2e(self, split=None):
        """One step through the reference-facing boundary with host buffers: forces in (H2D), step, positions out (D2H).
        split (a 3 element list) accumulates the wall time of the three calls."""
        if self.batch:
            t0 = time.perf_counter()
            self.api.b2j_batch_add_force_torque(self.batch, self.num_slots, self._f, None)
            t1 = time.perf_counter()
            self.api.b2j_batch_step(self.batch, DT, 1, C.byref(self.stats))
            t2 = time.perf_counter()
            self.api.b2j_batch_get_state(self.batch, 0xffffffff, self.num_slots, C.byref(self._st))
            if split is not None:
                t3 = time.perf_counter()
                split[0] += t1 - t0; split[1] += t2 - t1; split[2] += t3 - t2
        else:
            self.scene.step_e2e(DT, self.forces, self.positions)  # facade: BodyInterface::AddForce..., PhysicsSystem::Update, GetPosition


COUNTERS = ("num_body_pairs", "num_pairs_from_cache", "num_manifolds", "num_contact_points", "num_constraints", "num_phases", "velocity_iterations", "position_iterations", "num_active_bodies")


def measure(wl, torch, dist, world_size, local_rank, steps, warmup, with_e2e=True, with_profile=True, sample_clocks=True):
    """value leg, then (after a reset + the same warm-up each) the e2e leg and the per kernel profile leg: all over steps [W, W + K)."""
    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        wl.step()
    sampler = ClockSampler(local_rank)
    if sample_clocks and os.environ.get("B2J_BENCH_NO_SAMPLER") != "1":  # (diagnostics: does the sampler perturb the run?)
        sampler.start()
    barrier()
    cuprofile = os.environ.get("B2J_BENCH_CUPROFILE") == "1"  # `ncu --profile-from-start off`: capture the timed steps only
    if cuprofile:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    gpu_ms, launches, agg, series = 0.0, 0, {}, []
    for _ in range(steps):
        st = wl.step()
        gpu_ms += st.gpu_ms
        series.append(round(st.gpu_ms, 3))
        launches += st.kernel_launches
        for k in COUNTERS:
            agg[k] = agg.get(k, 0) + getattr(st, k)
    barrier()
    wall = time.perf_counter() - t0
    if cuprofile:
        torch.cuda.profiler.stop()
    clocks = sampler.finish()
    out = {"gpu_ms": gpu_ms, "wall": wall, "launches": launches, "agg": agg, "clocks": clocks, "series": series, "e2e": None, "prof": None}

    if with_e2e:
        # same window through the host-buffer boundary: reset, warm up again (device resident, untimed), then K timed e2e steps
        h2d, d2h = wl.e2e_buffers(torch)
        wl.step_e2e()  # untimed: the first call through the boundary sizes the library's staging buffers (pinned allocations)
        wl.reset()
        if not wl._e2e_ready:
            h2d, d2h = wl.e2e_buffers(torch)
        for _ in range(warmup):
            wl.step()
        barrier()
        t1 = time.perf_counter()
        split = [0.0, 0.0, 0.0]
        for _ in range(steps):
            wl.step_e2e(split)
        barrier()
        out["e2e"] = {"wall": time.perf_counter() - t1, "steps": steps, "h2d": h2d, "d2h": d2h, "split_ms": [round(1000.0 * x / steps, 3) for x in split]}

    if with_profile:
        # per kernel device time over the same window (event records perturb back-to-back launches, so this is its own leg)
        wl.reset()
        for _ in range(warmup):
            wl.step()
        prof_steps = max(1, min(steps, 10))
        wl.set_profiling(1)
        prof_gpu_ms, pagg = 0.0, {}
        for _ in range(prof_steps):
            st = wl.step()
            prof_gpu_ms += st.gpu_ms
            for k in ("num_contact_points", "num_constraints", "velocity_iterations"):
                pagg[k] = pagg.get(k, 0) + getattr(st, k)
        out.update({"prof": wl.profile(), "prof_steps": prof_steps, "prof_gpu_ms": prof_gpu_ms, "pagg": pagg})
        wl.set_profiling(0)
    return out


def ncu_traffic_ratio():
    """DRAM bytes / algorithmic bytes of the velocity solve kernel from the committed ncu --set full capture (profiles/*.json written
    by tools/ncu_traffic.py from the .ncu-rep of the round). None when there is no capture."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith("_solve_traffic.json"):
            try:
                best = (json.load(open(os.path.join(pdir, name))), name)
            except Exception:
                pass
    return best


SOLVE_KERNELS = ("KSolveVelocityWorlds", "KSolveVelocityAll", "KSolveVelocityAllPlain", "KSolveVelocity", "KSolveSmallVelocity")


def roofline_of(m):
    """Roofline of the dominant kernel class (velocity solve), SURVEY 8(d) row (5): per constraint and iteration
    C(c) + 4*S_v + 4*(3+c) algorithmic bytes with C(c) = 220 + 64 c; kernels that also run the warm start pass (iteration 0 of the
    survey's V + 1) count it."""
    prof, ps, pagg = m["prof"], m["prof_steps"], m["pagg"]
    kernel, solve = None, None
    for k in SOLVE_KERNELS:
        if prof.get(k) and prof[k]["ms"] > 0:
            kernel, solve = k, prof[k]
            break
    if solve is None:
        return None
    M = pagg["num_constraints"] / ps
    cbar = pagg["num_contact_points"] / max(pagg["num_constraints"], 1)
    V = pagg["velocity_iterations"] / ps
    passes = V if kernel == "KSolveVelocity" else V + 1   # the one launch kernels include the warm start pass
    total_bytes = passes * M * ((220 + 64 * cbar) + 4 * S_V + 4 * (3 + cbar)) * ps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = total_bytes / (solve["ms"] / 1000.0) / 1e9
    traffic, traffic_source = None, None
    t = ncu_traffic_ratio()
    if t is not None and t[0].get("kernel") == kernel and t[0].get("dram_bytes_per_algorithmic_byte"):
        traffic = t[0]["dram_bytes_per_algorithmic_byte"] * total_bytes / max(solve["launches"], 1)
        traffic_source = f"profiles/{t[1]}: (dram__bytes_read.sum + dram__bytes_write.sum) / algorithmic bytes of the captured launch = {t[0]['dram_bytes_per_algorithmic_byte']:.3f}, scaled to this run's bytes per launch"
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_source,
            "peak_source": "MEASURED_PEAKS.json (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
            "bytes_per_launch": total_bytes / max(solve["launches"], 1), "avg_launch_us": 1000.0 * solve["ms"] / max(solve["launches"], 1),
            "launches": solve["launches"], "passes_per_step": passes,
            "share_of_step": solve["ms"] / max(m["prof_gpu_ms"], 1e-9), "measured_over": f"{ps} profiled steps of the timed window (after a reset and the same warm-up)",
            "constraints_per_step": M, "points_per_constraint": cbar}


def secondary(api, flib, torch, dist, local_rank, workload, bodies, warmup, steps, cpu_budget, with_cpu=True, with_e2e=True):
    """One single-world measurement next to the main line: GPU value (+ e2e + roofline) and the reference on the SAME body count and
    step window."""
    wl = Workload(workload, bodies, 1, api, flib, 0, 1)
    m = measure(wl, torch, dist, 1, local_rank, steps, warmup, with_e2e=with_e2e, with_profile=True, sample_clocks=False)
    nd = wl.num_dynamic
    out = {"config": WORKLOAD_NAMES[workload], "bodies": nd, "steps": steps, "warmup": warmup, "window": f"simulation steps [{warmup}, {warmup + steps})",
           "value": steps * nd / (m["gpu_ms"] / 1000.0), "unit": "body-steps/s", "ms_per_step": m["gpu_ms"] / steps, "steps_per_sec": steps / (m["gpu_ms"] / 1000.0),
           "gpu_launches": m["launches"], "roofline": roofline_of(m), "step_counters_mean": {k: v / steps for k, v in m["agg"].items()},
           "kernel_ms_per_step": {k: round(v["ms"] / m["prof_steps"], 4) for k, v in sorted(m["prof"].items(), key=lambda kv: -kv[1]["ms"])[:10]}}
    if m["e2e"] is not None:
        out["e2e"] = {"value": m["e2e"]["steps"] * nd / m["e2e"]["wall"], "unit": "body-steps/s", "h2d_bytes_per_step": m["e2e"]["h2d"], "d2h_bytes_per_step": m["e2e"]["d2h"],
                      "path": "facade: BodyInterface::AddForcesAndTorques, PhysicsSystem::Update, BodyInterface::GetCenterOfMassPosition"}
    wl.close()
    del wl
    torch.cuda.synchronize()
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline_object(workload, bodies, 1, warmup, steps, cpu_budget)
        if "value" in out["cpu_baseline"] and out["cpu_baseline"]["steps"] == steps and out["cpu_baseline"]["warmup"] == warmup:
            out["speedup_same_window"] = out["value"] / out["cpu_baseline"]["value"]
    return out


def constraints_batch_extra(api, flib, torch, worlds, warmup, steps, with_cpu):
    """SURVEY 8 f4 measured (not a BASELINE config): `worlds` copies of the joints feature scene (feature_scenes.inl variant 11: point /
    distance / hinge / fixed constraints -- a pinned chain, a rope, a 10 x 10 cloth that is one large island, hinged planks and doors,
    welded cantilevers -- 216 constraints and 337 bodies per world) stepped as one batch, the reference beside it with one world per host
    thread. The RL pattern with articulated worlds."""
    import facade as F
    from joltphysics_b200 import _capi
    variant = 11
    scene = F.FacadeScene(flib, "feature", variant, 0)
    nd, nc = scene.num_dynamic, api.b2j_num_constraints(scene.world.h)
    batch = api.b2j_batch_create(scene.world.h, worlds, 0, 0)
    if not batch:
        scene.close()
        raise RuntimeError("b2j_batch_create failed: " + api.last_error())
    st = _capi.StepStats()
    try:
        for _ in range(warmup):
            if api.b2j_batch_step(batch, DT, 1, C.byref(st)) < 0:
                raise RuntimeError(api.last_error())
        torch.cuda.synchronize()
        gpu_ms, launches, contacts = 0.0, 0, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            if api.b2j_batch_step(batch, DT, 1, C.byref(st)) < 0:
                raise RuntimeError(api.last_error())
            gpu_ms += st.gpu_ms
            launches += st.kernel_launches
            contacts += st.num_constraints
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    finally:
        api.b2j_batch_destroy(batch)
        scene.close()
    out = {"config": f"{worlds} worlds of the joints feature scene (point / distance / hinge / fixed constraints), batched", "worlds": worlds, "bodies": worlds * nd,
           "constraints": worlds * nc, "contact_constraints_per_step": contacts / steps, "steps": steps, "warmup": warmup, "window": f"simulation steps [{warmup}, {warmup + steps})",
           "value": steps * worlds * nd / (gpu_ms / 1000.0), "unit": "body-steps/s", "ms_per_step": gpu_ms / steps, "wall_ms_per_step": 1000.0 * wall / steps, "gpu_launches": launches}
    if with_cpu:
        try:
            import refharness as R
            L = R.ref_lib("fast")
            threads = L.jref_hardware_threads()
            L.jref_time_worlds_parallel.restype = C.c_double
            L.jref_time_worlds_parallel.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
            cpu_wall = L.jref_time_worlds_parallel(b"feature", variant, 0, threads, warmup, steps, DT)
            out["cpu_baseline"] = {"value": threads * steps * nd / cpu_wall, "unit": "body-steps/s", "cores": threads, "kind": "reference", "steps": steps, "warmup": warmup,
                                   "sample": f"{threads} concurrent worlds of the same scene (one per host thread, single threaded job system each), steps [{warmup}, {warmup + steps}); " + CPU_BUILD}
            out["speedup_same_window"] = out["value"] / out["cpu_baseline"]["value"]
        except Exception as e:
            out["cpu_baseline"] = {"error": str(e)}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import joltphysics_b200
    import facade as F

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libjolt_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["B2J_DEVICE"] = str(local_rank)
    api = joltphysics_b200.load()
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)

    wl = Workload(args.workload, args.bodies, args.worlds, api, flib, rank, world_size)
    m = measure(wl, torch, dist, world_size, local_rank, args.steps, args.warmup)

    # max over ranks of the times, sum of the work (the only collective: a reduction of the statistics)
    (t_dev, wall, e2e_wall), (total_bodies, launches, total_worlds, h2d, d2h) = reduce_job(
        dist, world_size, (m["gpu_ms"] / 1000.0, m["wall"], m["e2e"]["wall"]), (wl.num_dynamic, m["launches"], wl.n_worlds, m["e2e"]["h2d"], m["e2e"]["d2h"]), "cuda")
    launches, total_worlds, h2d, d2h = int(launches), int(total_worlds), int(h2d), int(d2h)
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    batch = args.workload == "batch"
    config = job_config(args)
    line = {
        "metric": "body_steps_per_sec", "value": K * total_bodies / t_dev, "unit": "body-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_dev / K, "steps_per_sec": K / t_dev, "world_steps_per_sec": K * total_worlds / t_dev, "higher_is_better": True,
        "scaling": "strong" if batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "clocks": m["clocks"],
        "e2e": {"value": m["e2e"]["steps"] * total_bodies / e2e_wall, "unit": "body-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": m["e2e"]["steps"],
                "window": config["window"] + " (worlds reset to the creation state and warmed up again)",
                "ms_per_step": 1000.0 * e2e_wall / m["e2e"]["steps"], "forces_in__step__positions_out_ms": m["e2e"]["split_ms"],
                "path": "b2j_batch_add_force_torque + b2j_batch_step + b2j_batch_get_state (C ABI, pinned host buffers)" if batch else "facade: BodyInterface::AddForcesAndTorques, PhysicsSystem::Update, BodyInterface::GetCenterOfMassPosition"},
        "gpu_launches": launches,
        "wall_ms_per_step": 1000.0 * wall / K,
        "roofline": roofline_of(m),
        "step_counters_mean": {k: v / K for k, v in m["agg"].items()},
        "kernel_ms_per_step": {k: round(v["ms"] / m["prof_steps"], 4) for k, v in sorted(m["prof"].items(), key=lambda kv: -kv[1]["ms"])[:14]},
        "profiled_ms_per_step": m["prof_gpu_ms"] / m["prof_steps"],
        "ms_per_step_series": m["series"],
    }
    if world_size == 1:
        wl.close()
        del wl
        torch.cuda.synchronize()
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_object(args.workload, args.bodies, args.worlds, args.warmup, args.steps, args.cpu_seconds)
        if batch and not args.no_pile and not over_budget(args):
            # configs[3], one 1M body world on one B200. Two windows: the one the reference can reach on this host inside the bench
            # budget (same body count, same steps on both arms: the like-for-like ratio) and the formed pile (steps [120, 150), GPU only)
            try:
                line["pile"] = secondary(api, flib, torch, dist, local_rank, "pile", 1000000, 5, 10, 90.0, with_cpu=not args.no_cpu_baseline, with_e2e=True)
                if not over_budget(args):
                    formed = secondary(api, flib, torch, dist, local_rank, "pile", 1000000, 120, 30, 0.0, with_cpu=False, with_e2e=False)
                    line["pile"]["formed"] = {k: formed[k] for k in ("window", "steps", "warmup", "value", "unit", "ms_per_step", "roofline", "step_counters_mean", "kernel_ms_per_step")}
            except Exception as e:
                line["pile"] = {"error": str(e)}
        if batch and not args.no_extras:
            # configs[0..2] as single worlds, the reference beside them on the same body count and window
            line["extra"] = {}
            for name, bodies, w, k, cpu in (("pyramid", PYRAMID_BODIES, 100, 400, True), ("convex_vs_mesh", CONVEX_VS_MESH_BODIES, 100, 400, True),
                                            ("max_bodies", MAX_BODIES_FULL, 3, 6, False), ("max_bodies_1m", 1048576, 3, 6, True)):
                if over_budget(args):
                    line["extra"][name] = {"skipped": "bench wall clock budget reached"}
                    continue
                try:
                    line["extra"][name] = secondary(api, flib, torch, dist, local_rank, name.split("_1m")[0], bodies, w, k, 60.0, with_cpu=cpu and not args.no_cpu_baseline, with_e2e=name != "max_bodies")
                except Exception as e:
                    line["extra"][name] = {"error": str(e)}
            if over_budget(args):
                line["extra"]["constraints_batch"] = {"skipped": "bench wall clock budget reached"}
            else:
                try:
                    # two batch sizes: articulated toy worlds are the reference's best case (a world step is ~40 us of single threaded
                    # work), the device needs thousands of them to draw level
                    line["extra"]["constraints_batch"] = constraints_batch_extra(api, flib, torch, 8192, 20, 60, not args.no_cpu_baseline)
                    if not over_budget(args):
                        small = constraints_batch_extra(api, flib, torch, 2048, 20, 60, False)
                        line["extra"]["constraints_batch"]["worlds_2048"] = {k: small[k] for k in ("worlds", "bodies", "constraints", "value", "unit", "ms_per_step", "gpu_launches")}
                except Exception as e:
                    line["extra"]["constraints_batch"] = {"error": str(e)}
            line["extra"]["note"] = ("max_bodies runs the reference scene at its full 8 388 608 bodies on the GPU only (the reference needs ~30 s per step at that size on "
                                     "this host); max_bodies_1m is the same scene at 1 048 576 bodies on BOTH arms")
    line["bench_wall_s"] = time.time() - T_START
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
