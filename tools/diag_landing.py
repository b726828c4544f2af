"""Where do the steps go in which the boxes of the batched Pyramid worlds land (steps 19..25 from the creation state cost 1.5x the steps
before)? Per step: wall clock, counters, and the per kernel device times of that step (run under gpurun):
python tools/diag_landing.py [worlds=512] [first=15] [last=26]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: F401  (device context like bench.py)
import bench, joltphysics_b200, facade as F
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 512
first = int(sys.argv[2]) if len(sys.argv) > 2 else 15
last = int(sys.argv[3]) if len(sys.argv) > 3 else 26
wl = bench.Workload("batch", 0, worlds, api, flib, 0, 1)
for phase in ("plain", "profiled"):
    wl.reset()
    for s in range(last):
        if phase == "profiled" and s >= first:
            wl.set_profiling(1)
        t0 = time.perf_counter()
        st = wl.step()
        t1 = time.perf_counter()
        if s >= first:
            line = f"[{phase}] step {s}: wall {1e3 * (t1 - t0):.2f} ms gpu {st.gpu_ms:.2f} ms pairs {st.num_body_pairs} cached {st.num_pairs_from_cache} manifolds {st.num_manifolds} constraints {st.num_constraints} phases {st.num_phases} launches {st.kernel_launches}"
            if phase == "profiled":
                prof = wl.profile()
                wl.set_profiling(0)
                top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]
                line += "\n    " + "  ".join(f"{k} {v['ms']:.2f}/{v['launches']}" for k, v in top)
            print(line, flush=True)
wl.close()
