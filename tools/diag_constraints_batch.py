"""The constraints_batch extra of bench.py on its own (run under gpurun): python tools/diag_constraints_batch.py [worlds ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench, joltphysics_b200, facade as F
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
for worlds in [int(x) for x in sys.argv[1:]] or [256, 2048]:
    print(json.dumps(bench.constraints_batch_extra(api, flib, torch, worlds, 20, 60, worlds >= 2048)), flush=True)
