import sys, os, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import argparse, bench, joltphysics_b200, facade as F, torch
api = joltphysics_b200.load()
flib = F.FacadeLib("/root/repo/joltphysics_b200/libjolt_b200_facade.so", api)
args = argparse.Namespace(workload=sys.argv[1], worlds=int(sys.argv[2]), bodies=int(sys.argv[2]))
wl = bench.Workload(args, api, flib, 0, 1)
win = (int(sys.argv[3]), int(sys.argv[4]))
# optional: steps to bracket with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)
capture = set(int(x) for x in sys.argv[6].split(",")) if len(sys.argv) > 6 else set()
for i in range(int(sys.argv[5])):
    if i in capture: torch.cuda.synchronize(); torch.cuda.profiler.start()
    if i == win[0]: wl.set_profiling(1)
    st = wl.step()
    if i in capture: torch.cuda.synchronize(); torch.cuda.profiler.stop()
    if i == win[1]:
        prof = wl.profile(); wl.set_profiling(0)
        n = win[1]-win[0]+1
        print("PROFILE steps", win, {k: round(v["ms"]/n,2) for k,v in sorted(prof.items(), key=lambda kv:-kv[1]["ms"])[:16]})
    if i % 10 == 0 or i in win:
        print(i, "ms %.2f"%st.gpu_ms, "pairs", st.num_body_pairs, "cached", st.num_pairs_from_cache, "man", st.num_manifolds, "cons", st.num_constraints, "phases", st.num_phases, "launch", st.kernel_launches, "active", st.num_active_bodies, flush=True)
