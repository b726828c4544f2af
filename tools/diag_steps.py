"""Per step device time of a workload (no profiling). usage: python tools/diag_steps.py WORKLOAD SIZE FIRST LAST"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import argparse, bench, joltphysics_b200, facade as F
api = joltphysics_b200.load()
flib = F.FacadeLib("/root/repo/joltphysics_b200/libjolt_b200_facade.so", api)
wl = bench.Workload(argparse.Namespace(workload=sys.argv[1], worlds=int(sys.argv[2]), bodies=int(sys.argv[2])), api, flib, 0, 1)
first, last = int(sys.argv[3]), int(sys.argv[4])
ms = []
for i in range(last):
    st = wl.step()
    if i >= first: ms.append(round(st.gpu_ms, 1))
print(ms)
