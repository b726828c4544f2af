"""Experiment: total W worlds stepped as K independent batches from K host threads (own stream each) vs one batch.
usage: python tools/diag_streams.py W K steps"""
import sys, os, time, threading, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import argparse, bench, joltphysics_b200, facade as F, torch
api = joltphysics_b200.load()
flib = F.FacadeLib("/root/repo/joltphysics_b200/libjolt_b200_facade.so", api)
W, K, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
wls = [bench.Workload(argparse.Namespace(workload="batch", worlds=W // K, bodies=0), api, flib, 0, 1) for _ in range(K)]
def run(wl, n):
    for _ in range(n): wl.step()
def step_all(n):
    ts = [threading.Thread(target=run, args=(wl, n)) for wl in wls]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1000
done = 0
for chunk in (40, 20, 40, 20, 20):
    if done >= steps: break
    ms = step_all(chunk); done += chunk
    print(f"W={W} K={K} steps {done-chunk}..{done}: {ms:.2f} ms/step  {W*1240/ms/1e3:.2f} M body-steps/s", flush=True)
