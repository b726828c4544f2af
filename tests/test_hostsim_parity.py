"""Kernel logic vs the reference on the CPU: the kernel bodies of joltphysics_b200/csrc compiled for the host (tests/hostsim,
a debug aid -- NOT the product and not a fallback) are stepped against the reference oracle. The same protocol runs on the
GPU through libjolt_b200.so in test_gpu_parity.py."""
import pytest

import parity

CASES = [
    ("pyramid", 4, 0, 0), ("pyramid", 4, 0, 1), ("pyramid", 4, 0, 40),
    ("pyramid", 9, 0, 60),          # 330 boxes: one large island, exercises the split colouring
    ("small_stack", 0, 0, 30), ("small_stack", 1, 0, 30), ("small_stack", 2, 0, 30), ("small_stack", 3, 0, 30),
    ("small_stack", 4, 0, 5), ("small_stack", 4, 0, 45), ("small_stack", 4, 0, 200),
    # ScaledShape / RotatedTranslatedShape decorated convex bodies landing on the mesh (p1 = 1)
    ("convex_vs_mesh", 1, 1, 90), ("convex_vs_mesh", 1, 1, 150),
    # ... and with the mesh itself scaled + rotated (p1 bit 1)
    ("convex_vs_mesh", 1, 2, 300), ("convex_vs_mesh", 1, 3, 200),
    # StaticCompoundShape bodies (dumbbells, L shapes, tables, a 9 part cross with decorated parts) against a floor / each other / a
    # static compound staircase (p0 = 0) and on a terrain mesh (p0 = 1)
    ("compound", 0, 0, 0), ("compound", 0, 0, 60), ("compound", 0, 0, 100), ("compound", 0, 0, 300),
    ("compound", 1, 0, 80), ("compound", 1, 0, 120), ("compound", 1, 0, 300),
]


@pytest.mark.parametrize("scene,p0,p1,warm", CASES)
def test_single_step_parity(hostsim_api, scene, p0, p1, warm):
    parity.single_step_parity(hostsim_api, scene, p0, p1, warm)


FEATURE_CASES = [(name, warm) for name in parity.FEATURES for warm in ((0, 30, 120) if name != "zoo" else (10, 60, 250))]


@pytest.mark.parametrize("feature,warm", FEATURE_CASES, ids=[f"{n}-step{w}" for n, w in FEATURE_CASES])
def test_feature_single_step_parity(hostsim_api, feature, warm):
    """Motion types, body flags, DOF locks, step overrides, broadphase layers (see parity.FEATURES)."""
    parity.single_step_parity(hostsim_api, "feature", parity.FEATURES.index(feature), 0, warm)


@pytest.mark.parametrize("slot", [1, 5, 24])
def test_body_recreated_in_the_same_slot_is_not_served_from_the_cache(hostsim_api, slot):
    """destroy a resting box, create a sphere with the next sequence number in the same slot at the same pose. The
    contact cache still holds the dead body's pairs; the reference keys them by the full BodyID, so the new body must collide afresh
    (ContactAdded with the new id, ContactRemoved for the old manifolds)."""
    out = parity.single_step_parity(hostsim_api, "small_stack", 1, 0, 60, before_export=lambda ref: ref.replace_body(slot))
    assert out["manifolds"] > 0


def test_two_collision_steps(hostsim_api):
    parity.single_step_parity(hostsim_api, "pyramid", 4, 0, 20, collision_steps=2)


def test_set_params_through_the_c_abi(hostsim_api):
    """b2j_bodies_set_params (BodyInterface::SetGravityFactor / SetMaxLinearVelocity ...): only the given members of the given bodies
    change. A free falling body with gravity factor 0 stays where it is, one with a velocity cap falls at the cap."""
    import ctypes as C
    import numpy as np
    import refharness as R
    from joltphysics_b200 import _capi
    ref = R.RefWorld("small_stack", 0)
    world = ref.export(hostsim_api)
    before = world.state()
    ids = before.ids if hasattr(before, "ids") else None
    state_ids = ref.state().ids
    top, second = int(state_ids[24]), int(state_ids[23])   # the two highest bodies of the stack: nothing rests on them
    arr = np.array([top, second], dtype=np.uint32)
    gf = np.array([0.0, 1.0], dtype=np.float32)
    cap = np.array([500.0, 0.25], dtype=np.float32)
    p = _capi.BodyParams(None, None, gf.ctypes.data, None, None, cap.ctypes.data, None)
    assert hostsim_api.b2j_bodies_set_params(world.h, arr.ctypes.data_as(C.POINTER(C.c_uint32)), 2, C.byref(p)) == 0, hostsim_api.last_error()
    for _ in range(10):
        world.step()
    after = world.state()
    assert abs(after.pos[24, 1] - before.pos[24, 1]) < 1e-6 and np.allclose(after.lin[24], 0.0), "gravity factor 0: the body must not move"
    assert np.linalg.norm(after.lin[23]) <= 0.25 + 1e-6, "max linear velocity must clamp the falling body"
    assert after.pos[22, 1] < before.pos[22, 1] - 0.05, "untouched bodies fall freely"
    world.close()
    ref.close()


def test_bulk_mutations_and_id_validation_through_the_c_abi(hostsim_api):
    """ABI hygiene (VERDICT r1 #10 / ADVICE r1): bulk remove / deactivate / activate in one call each, stale and out of range ids are
    rejected (never touch the body that lives in the slot now), WereBodiesInContact through the device pair table, event recording
    follows the listener switch."""
    import ctypes as C
    import numpy as np
    import refharness as R
    api = hostsim_api
    u32p = C.POINTER(C.c_uint32)
    ref = R.RefWorld("small_stack", 1)
    for _ in range(40):
        ref.step()
    world = ref.export(api)
    world.step()
    ids = ref.state().ids
    floor, a, b, c = int(ids[0]), int(ids[1]), int(ids[2]), int(ids[4])
    # contact query: bodies 1..3 rest on the floor (slot 0)
    assert api.b2j_were_bodies_in_contact(world.h, floor, a) == 1
    assert api.b2j_were_bodies_in_contact(world.h, a, floor) == 1
    assert api.b2j_were_bodies_in_contact(world.h, a, int(ids[24])) == 0
    assert api.b2j_were_bodies_in_contact(world.h, a, 0x7fffff) == 0
    # stale / invalid ids
    stale = np.array([a ^ (1 << 23)], dtype=np.uint32)   # same slot, other sequence number
    assert api.b2j_bodies_activate(world.h, stale.ctypes.data_as(u32p), 1) == -1 and "not a body" in api.last_error()
    assert api.b2j_bodies_remove(world.h, stale.ctypes.data_as(u32p), 1) == -1
    big = np.array([0x7ffff0], dtype=np.uint32)
    assert api.b2j_bodies_remove(world.h, big.ctypes.data_as(u32p), 1) == -1
    n_before, active_before = api.b2j_num_bodies(world.h), api.b2j_num_active_bodies(world.h)
    assert api.b2j_bodies_deactivate(world.h, stale.ctypes.data_as(u32p), 1) == 0
    assert api.b2j_num_active_bodies(world.h) == active_before, "a stale id must not deactivate the body living in that slot"
    # bulk deactivate (with a duplicate), bulk activate (with duplicates and already active bodies)
    lst = np.array([a, b, a, c], dtype=np.uint32)
    assert api.b2j_bodies_deactivate(world.h, lst.ctypes.data_as(u32p), 4) == 0, api.last_error()
    assert api.b2j_num_active_bodies(world.h) == active_before - 3
    st = world.state()
    assert np.all(st.active_index[[1, 2, 4]] == 0xffffffff) and np.all(st.lin[[1, 2, 4]] == 0.0)
    order = world.active_bodies()
    assert len(set(order.tolist())) == len(order) == active_before - 3
    lst2 = np.array([c, a, int(ids[10]), c, floor], dtype=np.uint32)   # ids[10] is already active, the floor is static
    assert api.b2j_bodies_activate(world.h, lst2.ctypes.data_as(u32p), 5) == 0, api.last_error()
    assert api.b2j_num_active_bodies(world.h) == active_before - 1
    assert world.active_bodies()[-2:].tolist() == [c, a], "activation appends in argument order, duplicates once"
    # bulk remove
    rm = np.array([b, int(ids[7]), int(ids[20])], dtype=np.uint32)
    assert api.b2j_bodies_remove(world.h, rm.ctypes.data_as(u32p), 3) == 0, api.last_error()
    assert api.b2j_num_bodies(world.h) == n_before - 3
    assert api.b2j_bodies_remove(world.h, rm.ctypes.data_as(u32p), 3) == -1, "removing twice: the ids are stale now"
    err, stats = world.step()
    assert err == 0 and stats.num_bodies == n_before - 3
    # events: on by default, none with recording off
    assert len(world.contact_events()) > 0
    assert api.b2j_world_set_event_recording(world.h, 0, 0) == 0
    world.step()
    assert world.contact_events() == [] and world.activation_events() == []
    assert api.b2j_world_set_event_recording(world.h, 1, 1) == 0
    world.step()
    assert len(world.contact_events()) > 0
    world.close()
    ref.close()
