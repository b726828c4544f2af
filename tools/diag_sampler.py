"""Does the nvidia-smi clock sampler of bench.py perturb a launch-heavy single world step? usage: python tools/diag_sampler.py"""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import argparse, bench, joltphysics_b200, facade as F, torch
api = joltphysics_b200.load()
flib = F.FacadeLib("/root/repo/joltphysics_b200/libjolt_b200_facade.so", api)
wl = bench.Workload(argparse.Namespace(workload="pile", worlds=1, bodies=1000000), api, flib, 0, 1)
for _ in range(120): wl.step()
def run(n, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g = 0.0
    for _ in range(n): g += wl.step().gpu_ms
    torch.cuda.synchronize(); w = (time.perf_counter() - t0) * 1000 / n
    print(f"{tag}: gpu {g/n:.2f} ms/step, wall {w:.2f} ms/step", flush=True)
run(10, "no sampler")
s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
run(10, "sampler on")
print(s.finish())
run(10, "sampler off again")
