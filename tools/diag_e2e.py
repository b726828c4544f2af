"""Where does the host-buffer boundary spend its time? (run under gpurun) forces in / positions out of a 1024 world batch with page
locked and with pageable caller buffers."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import joltphysics_b200, facade as F
from joltphysics_b200 import _capi
api = joltphysics_b200.load()
flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), api)
proto = F.FacadeScene(flib, "pyramid", 15, 0)
nw = int(os.environ.get("WORLDS", "1024"))
batch = api.b2j_batch_create(proto.world.h, nw, 16384, 12288)
n = nw * proto.num_bodies
stats = _capi.StepStats()
for _ in range(3):
    api.b2j_batch_step(batch, 1 / 60, 1, C.byref(stats))
fp = C.POINTER(C.c_float)
for kind in ("pinned", "pageable"):
    if kind == "pinned":
        f = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy(); p = torch.zeros((n, 3), dtype=torch.float32).pin_memory().numpy()
    else:
        f = np.zeros((n, 3), np.float32); p = np.zeros((n, 3), np.float32)
    st = _capi.BodyState(p.ctypes.data, None, None, None, None, None, None)
    for rep in range(3):
        t0 = time.perf_counter(); api.b2j_batch_add_force_torque(batch, n, f.ctypes.data_as(fp), None)
        t1 = time.perf_counter(); api.b2j_batch_get_state(batch, 0xffffffff, n, C.byref(st))
        t2 = time.perf_counter()
        print(f"{kind} rep {rep}: {n * 12 / 1e6:.1f} MB each way; forces in {1e3 * (t1 - t0):.2f} ms, positions out {1e3 * (t2 - t1):.2f} ms", flush=True)
t = torch.zeros((n, 3), dtype=torch.float32).pin_memory(); d = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(t, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter(); t.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"torch pinned copies of the same size: H2D {1e3 * (t1 - t0):.2f} ms, D2H {1e3 * (t2 - t1):.2f} ms")
