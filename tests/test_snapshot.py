"""Device resident SaveState / RestoreState (PhysicsSystem.cpp:2899-2964, ContactConstraintManager.cpp:467-548): a snapshot
taken at step k and restored after 30 more steps must put the world back EXACTLY where the reference is at step k (bodies, active
list, contact cache) and stepping on from there must retrace the first run bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
import refharness as R
import facade as F
from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("pos", "rot", "lin", "ang", "bounds", "active_index", "sleep_timer")


def _equal_states(a, b, what):
    for name in FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), f"{what}: {name} differs"


def _cache_rows(world):
    return R.cache_rows(R.cache_summary(*world.cache()))


def _check_save_restore(api, scene, p0, p1, k, more):
    ref = R.RefWorld(scene, p0, p1)
    for _ in range(k):
        ref.step()
    world = ref.export(api)
    world.step(); ref.step()                      # one step so that the snapshot holds a cache this library wrote itself
    at_save, cache_at_save, active_at_save = world.state(), _cache_rows(world), world.active_bodies()
    ref_at_save = ref.state()
    snap = api.b2j_world_save_state(world.h)
    assert snap, api.last_error()
    assert api.b2j_snapshot_size(snap) > 0
    first_run = []
    for _ in range(more):
        world.step()
        first_run.append(world.state())
    assert not np.array_equal(first_run[-1].pos, at_save.pos), "the scene must move between save and restore"
    # mutate the set of bodies too: a snapshot restores it (RestoreState re-adds / removes bodies)
    victim = np.array([int(at_save_id) for at_save_id in ref.state().ids[1:3]], dtype=np.uint32)
    assert api.b2j_bodies_remove(world.h, victim.ctypes.data_as(C.POINTER(C.c_uint32)), 2) == 0, api.last_error()
    n_removed = api.b2j_num_bodies(world.h)
    assert api.b2j_world_restore_state(world.h, snap) == 0, api.last_error()
    assert api.b2j_num_bodies(world.h) == n_removed + 2
    _equal_states(at_save, world.state(), "right after the restore")
    assert np.array_equal(cache_at_save, _cache_rows(world)), "contact cache after the restore"
    assert np.array_equal(active_at_save, world.active_bodies()), "active list order after the restore"
    # against the oracle: the restored world is where the reference was when the snapshot was taken
    worst = R.compare_states(ref_at_save, world.state())
    assert max(worst[x] for x in ("pos", "rot", "lin", "ang")) <= 1.0, worst
    for i in range(more):
        world.step()
        _equal_states(first_run[i], world.state(), f"step {i + 1} after the restore")
    # restore twice from the same snapshot
    assert api.b2j_world_restore_state(world.h, snap) == 0, api.last_error()
    _equal_states(at_save, world.state(), "second restore")
    # ... and the reference stepped the same number of steps agrees with the replay
    for _ in range(more):
        ref.step(); world.step()
    worst = R.compare_states(ref.state(), world.state())
    assert max(worst[x] for x in ("pos", "rot", "lin", "ang")) <= 1.0, worst
    api.b2j_snapshot_destroy(snap)
    world.close()
    ref.close()


CASES = [("small_stack", 4, 0, 30, 30), ("pyramid", 5, 0, 20, 30), ("feature", parity.FEATURES.index("zoo"), 0, 60, 30),
         ("compound", 1, 0, 100, 40)]  # compounds, decorated parts and cylinders on a mesh


@pytest.mark.parametrize("scene,p0,p1,k,more", CASES)
def test_save_restore_hostsim(hostsim_api, scene, p0, p1, k, more):
    _check_save_restore(hostsim_api, scene, p0, p1, k, more)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1,k,more", CASES + [("pyramid", 15, 0, 60, 30), ("pile", 3000, 15, 100, 30), ("convex_vs_mesh", 6, 0, 60, 30)])
def test_save_restore_gpu(gpu_api, ref_available, scene, p0, p1, k, more):
    _check_save_restore(gpu_api, scene, p0, p1, k, more)


def _batch_state(api, batch, n):
    s = R.State(n)
    st = _capi.BodyState(s.pos.ctypes.data, s.rot.ctypes.data, s.lin.ctypes.data, s.ang.ctypes.data, s.bounds.ctypes.data, s.active_index.ctypes.data, s.sleep_timer.ctypes.data)
    assert api.b2j_batch_get_state(batch, 0xffffffff, n, C.byref(st)) == 0, api.last_error()
    return s


def _check_batch_save_restore(api, flib, n_worlds, k, more):
    proto = F.FacadeScene(flib, "pyramid", 5, 0)
    n = proto.num_bodies * n_worlds
    batch = api.b2j_batch_create(proto.world.h, n_worlds, 0, 0)
    assert batch, api.last_error()
    stats = _capi.StepStats()
    for _ in range(k):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0
    at_save = _batch_state(api, batch, n)
    snap = api.b2j_batch_save_state(batch)
    assert snap, api.last_error()
    first = []
    for _ in range(more):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0
        first.append(_batch_state(api, batch, n))
    assert api.b2j_world_restore_state(proto.world.h, snap) == -1, "a batch snapshot is not a world snapshot"
    assert api.b2j_batch_restore_state(batch, snap) == 0, api.last_error()
    _equal_states(at_save, _batch_state(api, batch, n), "batch right after the restore")
    for i in range(more):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0
        _equal_states(first[i], _batch_state(api, batch, n), f"batch step {i + 1} after the restore")
    api.b2j_snapshot_destroy(snap)
    api.b2j_batch_destroy(batch)
    proto.close()


def test_batch_save_restore_hostsim(hostsim_api):
    flib = F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api)
    _check_batch_save_restore(hostsim_api, flib, 3, 25, 20)


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [1, 3])
def test_batch_save_restore_gpu(gpu_api, monkeypatch, groups):
    monkeypatch.setenv("B2J_BATCH_GROUPS", str(groups))
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check_batch_save_restore(gpu_api, flib, 7, 30, 25)
