"""The constraints_batch extra of bench.py on its own (run under gpurun): python tools/diag_constraints_batch.py [worlds ...]
With B2J_DIAG_PROFILE=1 the per kernel device times of 10 steps of the last batch size are printed as well."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench, joltphysics_b200, facade as F
from joltphysics_b200 import _capi
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
sizes = [int(x) for x in sys.argv[1:]] or [256, 2048]
for worlds in sizes:
    print(json.dumps(bench.constraints_batch_extra(api, flib, torch, worlds, 20, 60, worlds >= 2048)), flush=True)
if os.environ.get("B2J_DIAG_PROFILE") == "1":
    scene = F.FacadeScene(flib, "feature", 11, 0)
    batch = api.b2j_batch_create(scene.world.h, sizes[-1], 0, 0)
    st = _capi.StepStats()
    for _ in range(20):
        api.b2j_batch_step(batch, 1 / 60, 1, C.byref(st))
    api.b2j_batch_set_profiling(batch, 1)
    for _ in range(10):
        api.b2j_batch_step(batch, 1 / 60, 1, C.byref(st))
    cap, stride = 128, 64
    names = C.create_string_buffer(cap * stride)
    ms, launches = (C.c_float * cap)(), (C.c_uint32 * cap)()
    n = api.b2j_batch_get_profile(batch, names, stride, ms, launches, cap)
    rows = sorted(((names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(), ms[i] / 10, launches[i] // 10) for i in range(min(n, cap))), key=lambda r: -r[1])
    print(f"per kernel ms per step ({sizes[-1]} worlds, one group at a time):", [(a, round(b, 3), c) for a, b, c in rows[:14]])
    api.b2j_batch_destroy(batch)
    scene.close()
