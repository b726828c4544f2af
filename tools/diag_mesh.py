"""Steps the ConvexVsMesh scene (484 mixed convex bodies on a 100x100 quad terrain) through the impact phase: the workload of
KCollideMesh (run under ncu by tools/r2_run10.sh)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import joltphysics_b200, facade as F
from joltphysics_b200 import _capi
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
s = F.FacadeScene(flib, "convex_vs_mesh", 10, 0)
st = _capi.StepStats()
for _ in range(260):
    api.b2j_step(s.world.h, 1 / 60, 1, C.byref(st))
print("constraints", st.num_constraints)
s.close()
