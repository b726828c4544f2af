"""Batched independent worlds (config 5): n clones of one prototype world live in ONE device world and must evolve exactly like
the prototype stepped alone (the reference's TestMultiplePhysicsSystems pattern, PhysicsTests.cpp:1548)."""
import ctypes as C
import os

import numpy as np
import pytest

import refharness as R
import facade as F
from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch_state(api, batch, world, n):
    s = R.State(n)
    st = _capi.BodyState(s.pos.ctypes.data, s.rot.ctypes.data, s.lin.ctypes.data, s.ang.ctypes.data, s.bounds.ctypes.data, s.active_index.ctypes.data, s.sleep_timer.ctypes.data)
    assert api.b2j_batch_get_state(batch, world, n, C.byref(st)) == 0, api.last_error()
    return s


def _check_batch(api, flib, scene, p0, p1, n_worlds, steps):
    proto = F.FacadeScene(flib, scene, p0, p1)
    n = proto.num_bodies
    batch = api.b2j_batch_create(proto.world.h, n_worlds, 0, 0)
    assert batch, api.last_error()
    assert api.b2j_batch_size(batch) == n_worlds
    stats = _capi.StepStats()
    single = None
    for _ in range(steps):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, api.last_error()
        _, single = proto.world.step()
    assert stats.num_constraints == n_worlds * single.num_constraints
    assert stats.num_body_pairs == n_worlds * single.num_body_pairs
    want = proto.world.state()
    for w in sorted({0, n_worlds // 2, n_worlds - 1}):
        got = _batch_state(api, batch, w, n)
        for name in ("pos", "rot", "lin", "ang", "bounds"):
            assert np.array_equal(getattr(want, name), getattr(got, name)), f"world {w}: {name} differs from the prototype stepped alone"
        # active indices are per batch (world major): same active flags
        assert np.array_equal(want.active_index != 0xffffffff, got.active_index != 0xffffffff)
    api.b2j_batch_destroy(batch)
    proto.close()


@pytest.fixture(scope="session")
def hostsim_facade(hostsim_api):
    return F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api)


@pytest.mark.parametrize("scene,p0,p1,n_worlds,steps", [("pyramid", 4, 0, 5, 40), ("pile", 300, 15, 3, 60), ("convex_vs_mesh", 1, 0, 4, 80), ("compound", 0, 0, 3, 100),
                                                         ("feature", 11, 0, 3, 90)])  # feature 11 = the joints scene: every world carries the constraints
def test_batch_matches_single_world_hostsim(hostsim_api, hostsim_facade, scene, p0, p1, n_worlds, steps):
    _check_batch(hostsim_api, hostsim_facade, scene, p0, p1, n_worlds, steps)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1,n_worlds,steps", [("pyramid", 15, 0, 16, 60), ("pile", 1000, 15, 8, 90), ("convex_vs_mesh", 3, 0, 6, 120), ("compound", 0, 0, 12, 150),
                                                         ("feature", 11, 0, 9, 200)])
def test_batch_matches_single_world_gpu(gpu_api, scene, p0, p1, n_worlds, steps):
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check_batch(gpu_api, flib, scene, p0, p1, n_worlds, steps)


def _check_batch_vs_oracle(api, scene, p0, p1, n_worlds, steps, check_every):
    """n clones of a reference-created world against the REFERENCE itself stepped alone (not against this library's single world)."""
    ref = R.RefWorld(scene, p0, p1)
    proto = ref.export(api)          # step 0: empty contact cache, as b2j_batch_create requires
    n = ref.num_slots()
    batch = api.b2j_batch_create(proto.h, n_worlds, 0, 0)
    assert batch, api.last_error()
    stats = _capi.StepStats()
    worst_all = {}
    for step in range(1, steps + 1):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, api.last_error()
        ref.step(1.0 / 60.0, 1, 0)
        if step % check_every == 0 or step == steps:
            want = ref.state()
            for w in sorted({0, 1, n_worlds // 2, n_worlds - 1}):
                got = _batch_state(api, batch, w, n)
                got.ids = want.ids
                worst = R.compare_states(want, got)
                for k in ("pos", "rot", "lin", "ang"):
                    assert worst[k] <= 1.0, f"step {step} world {w}: {k} out of tolerance vs the reference: {worst}"
                    worst_all[k] = max(worst_all.get(k, 0.0), worst[k])
                assert np.array_equal(want.active_index != 0xffffffff, got.active_index != 0xffffffff), f"step {step} world {w}: active flags"
    api.b2j_batch_destroy(batch)
    proto.close()
    ref.close()
    return worst_all


def test_batch_vs_oracle_hostsim(hostsim_api):
    _check_batch_vs_oracle(hostsim_api, "pyramid", 4, 0, 3, 40, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1,n_worlds,steps", [("pyramid", 15, 0, 64, 60), ("feature", 8, 0, 64, 120)], ids=["pyramid-64worlds", "feature_zoo-64worlds"])
def test_batch_vs_oracle_gpu(gpu_api, ref_available, scene, p0, p1, n_worlds, steps):
    """64 world batch against the reference stepped alone (the batch had only been compared with this library's own
    single world)."""
    _check_batch_vs_oracle(gpu_api, scene, p0, p1, n_worlds, steps, 20)


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [2, 3])
def test_batch_groups_gpu(gpu_api, groups, monkeypatch):
    """Large batches are split into groups with their own stream and host thread (B2J_BATCH_GROUPS forces it for a small batch):
    every world must still evolve exactly like the prototype stepped alone, through the all-worlds and the per-world getters."""
    monkeypatch.setenv("B2J_BATCH_GROUPS", str(groups))
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check_batch(gpu_api, flib, "pyramid", 6, 0, 7, 50)
    # all-worlds getter / force setter across the group boundaries
    proto = F.FacadeScene(flib, "pyramid", 5, 0)
    n, n_worlds = proto.num_bodies, 7
    batch = gpu_api.b2j_batch_create(proto.world.h, n_worlds, 0, 0)
    assert batch, gpu_api.last_error()
    force = np.zeros((n_worlds * n, 3), dtype=np.float32)
    force[:, 0] = 4.0e5  # boxes of 8000 kg
    fp = C.POINTER(C.c_float)
    stats = _capi.StepStats()
    for _ in range(20):
        assert gpu_api.b2j_batch_add_force_torque(batch, n_worlds * n, force.ctypes.data_as(fp), None) == 0, gpu_api.last_error()
        assert gpu_api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, gpu_api.last_error()
    everything = _batch_state(gpu_api, batch, 0xffffffff, n_worlds * n)
    for w in range(n_worlds):
        one = _batch_state(gpu_api, batch, w, n)
        assert np.array_equal(everything.pos[w * n:(w + 1) * n], one.pos)
        assert np.array_equal(everything.pos[:n], one.pos), "all worlds get the same forces: identical trajectories"
    assert everything.pos[1:n, 0].mean() > 0.1 + proto.world.state().pos[1:n, 0].mean(), "the forces must have pushed the dynamic bodies along +x"
    gpu_api.b2j_batch_destroy(batch)
    proto.close()


def _check_reset(api, flib, scene, p0, p1, n_worlds, steps_before, steps_after, reset):
    """b2j_batch_reset_worlds: reset worlds evolve exactly like a newly created world, the others are not disturbed."""
    old = F.FacadeScene(flib, scene, p0, p1)     # stepped along with the untouched worlds
    n = old.num_bodies
    batch = api.b2j_batch_create(old.world.h, n_worlds, 0, 0)
    assert batch, api.last_error()
    stats = _capi.StepStats()
    for _ in range(steps_before):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, api.last_error()
        old.world.step()
    ids = np.array(reset, dtype=np.uint32)
    assert api.b2j_batch_reset_worlds(batch, ids.ctypes.data_as(C.POINTER(C.c_uint32)), len(ids)) == 0, api.last_error()
    new = F.FacadeScene(flib, scene, p0, p1)     # what a reset world must look like
    for w in reset:
        got, want = _batch_state(api, batch, w, n), new.world.state()
        for name in ("pos", "rot", "lin", "ang", "bounds", "sleep_timer"):
            assert np.array_equal(getattr(want, name), getattr(got, name)), f"world {w} right after the reset: {name}"
        assert np.array_equal(want.active_index != 0xffffffff, got.active_index != 0xffffffff)
    for _ in range(steps_after):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, api.last_error()
        old.world.step()
        new.world.step()
    for w in range(n_worlds):
        want = (new if w in reset else old).world.state()
        got = _batch_state(api, batch, w, n)
        for name in ("pos", "rot", "lin", "ang", "bounds"):
            assert np.array_equal(getattr(want, name), getattr(got, name)), f"world {w} ({'reset' if w in reset else 'untouched'}): {name} differs"
        assert np.array_equal(want.active_index != 0xffffffff, got.active_index != 0xffffffff), f"world {w}: active flags"
    api.b2j_batch_destroy(batch)
    old.close()
    new.close()


# pile: bodies fall asleep before the reset (they must be re-activated); small_stack would do as well but has no facade scene
@pytest.mark.parametrize("scene,p0,p1,before,after", [("pyramid", 4, 0, 25, 30), ("pile", 200, 15, 260, 40), ("feature", 11, 0, 50, 40)])
def test_batch_reset_worlds_hostsim(hostsim_api, hostsim_facade, scene, p0, p1, before, after):
    _check_reset(hostsim_api, hostsim_facade, scene, p0, p1, 5, before, after, [1, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1,before,after,groups", [("pyramid", 8, 0, 40, 40, 1), ("pile", 500, 15, 300, 60, 1), ("pyramid", 6, 0, 30, 30, 3), ("feature", 11, 0, 60, 60, 2)])
def test_batch_reset_worlds_gpu(gpu_api, monkeypatch, scene, p0, p1, before, after, groups):
    monkeypatch.setenv("B2J_BATCH_GROUPS", str(groups))
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check_reset(gpu_api, flib, scene, p0, p1, 7, before, after, [0, 2, 6])


@pytest.mark.gpu
def test_batch_on_two_devices_gpu(gpu_api):
    """Library level multi device batch (b2j_batch_create_on_devices): worlds split over two GPUs of one process, every world still
    evolves exactly like the prototype stepped alone; reset and snapshots work across the devices."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    proto = F.FacadeScene(flib, "pyramid", 6, 0)
    n, n_worlds = proto.num_bodies, 7
    devices = (C.c_int32 * 2)(0, 1)
    batch = gpu_api.b2j_batch_create_on_devices(proto.world.h, n_worlds, devices, 2, 0, 0)
    assert batch, gpu_api.last_error()
    stats = _capi.StepStats()
    for _ in range(40):
        assert gpu_api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0, gpu_api.last_error()
        _, single = proto.world.step()
    assert stats.num_constraints == n_worlds * single.num_constraints
    want = proto.world.state()
    for w in range(n_worlds):
        got = _batch_state(gpu_api, batch, w, n)
        for name in ("pos", "rot", "lin", "ang", "bounds"):
            assert np.array_equal(getattr(want, name), getattr(got, name)), f"world {w}: {name} differs from the prototype stepped alone"
    snap = gpu_api.b2j_batch_save_state(batch)
    assert snap, gpu_api.last_error()
    for _ in range(10):
        assert gpu_api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0
    assert gpu_api.b2j_batch_restore_state(batch, snap) == 0, gpu_api.last_error()
    for w in (0, n_worlds - 1):
        got = _batch_state(gpu_api, batch, w, n)
        assert np.array_equal(want.pos, got.pos), f"world {w} after the restore"
    ids = np.array([1, n_worlds - 1], dtype=np.uint32)
    assert gpu_api.b2j_batch_reset_worlds(batch, ids.ctypes.data_as(C.POINTER(C.c_uint32)), 2) == 0, gpu_api.last_error()
    fresh = F.FacadeScene(flib, "pyramid", 6, 0)
    for w in (1, n_worlds - 1):
        assert np.array_equal(fresh.world.state().pos, _batch_state(gpu_api, batch, w, n).pos), f"world {w} after the reset"
    gpu_api.b2j_snapshot_destroy(snap)
    gpu_api.b2j_batch_destroy(batch)
    fresh.close()
    proto.close()
