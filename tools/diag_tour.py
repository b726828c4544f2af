"""API tour phase by phase against the reference on the GPU: which body diverges first (run under gpurun; needs oracle/_ref)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import joltphysics_b200, refharness as R, facade as F
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
ref = R.RefWorld("api_tour", 0); fs = F.FacadeScene(flib, "api_tour", 0, 0)
for phase, steps in ((0, 10), (1, 25), (2, 40), (3, 30), (4, 60), (5, 50)):
    if phase: ref.mutate(phase); fs.mutate(phase)
    for step in range(steps):
        ref.step(); fs.update()
        fs.world.n = 18
        rs, gs = ref.state(18), fs.world.state()
        bad = [i for i in range(18) if not (np.array_equal(rs.pos[i], gs.pos[i]) and np.array_equal(rs.rot[i], gs.rot[i]) and np.array_equal(rs.lin[i], gs.lin[i]) and np.array_equal(rs.ang[i], gs.ang[i]))]
        if bad:
            print(f"phase {phase} step {step}: bodies {bad}")
            for i in bad[:4]:
                print("  ", i, "pos", rs.pos[i], gs.pos[i], "lin", rs.lin[i], gs.lin[i], "ang", rs.ang[i], gs.ang[i], "active", rs.active_index[i], gs.active_index[i])
            if phase == 5 and step > 3: sys.exit(0)
print("no divergence")
