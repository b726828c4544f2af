"""Per stage wall clock of single small worlds (run under gpurun): python tools/diag_small.py [scene ...]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import joltphysics_b200, facade as F
from joltphysics_b200 import _capi
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
for scene, p0 in (("convex_vs_mesh", 10), ("pyramid", 15), ("feature", 11)):  # feature 11 = the joints scene (constraints: one cooperative launch per solve)
    s = F.FacadeScene(flib, scene, p0, 0)
    st = _capi.StepStats()
    for _ in range(150):
        api.b2j_step(s.world.h, 1 / 60, 1, C.byref(st))
    t0 = time.perf_counter(); g = 0.0
    for _ in range(100):
        api.b2j_step(s.world.h, 1 / 60, 1, C.byref(st)); g += st.gpu_ms
    t1 = time.perf_counter()
    print(f"{scene}: wall {10 * (t1 - t0):.3f} ms/step, gpu {g / 100:.3f} ms/step, launches {st.kernel_launches}, constraints {st.num_constraints}, phases {st.num_phases}", flush=True)
    os.environ["B2J_TRACE_STEP"] = "1"
    for _ in range(4):
        api.b2j_step(s.world.h, 1 / 60, 1, C.byref(st))
    del os.environ["B2J_TRACE_STEP"]
    s.close()
