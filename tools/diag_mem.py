"""Device memory of the default bench workloads. usage: python tools/diag_mem.py"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import argparse, bench, joltphysics_b200, facade as F, torch
api = joltphysics_b200.load()
flib = F.FacadeLib("/root/repo/joltphysics_b200/libjolt_b200_facade.so", api)
free0, total = torch.cuda.mem_get_info()
for name, ns in (("batch 4096", argparse.Namespace(workload="batch", worlds=4096, bodies=0)), ("pile 1M", argparse.Namespace(workload="pile", worlds=1, bodies=1000000))):
    wl = bench.Workload(ns, api, flib, 0, 1)
    for _ in range(3): wl.step()
    free1, _ = torch.cuda.mem_get_info()
    print(name, "uses %.1f GB of %.1f GB" % ((free0 - free1) / 2**30, total / 2**30), flush=True)
    wl.close(); del wl
    torch.cuda.synchronize()
