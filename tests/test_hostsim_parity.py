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
]


@pytest.mark.parametrize("scene,p0,p1,warm", CASES)
def test_single_step_parity(hostsim_api, scene, p0, p1, warm):
    parity.single_step_parity(hostsim_api, scene, p0, p1, warm)


FEATURE_CASES = [(name, warm) for name in parity.FEATURES for warm in ((0, 30, 120) if name != "zoo" else (10, 60, 250))]


@pytest.mark.parametrize("feature,warm", FEATURE_CASES, ids=[f"{n}-step{w}" for n, w in FEATURE_CASES])
def test_feature_single_step_parity(hostsim_api, feature, warm):
    """Motion types, body flags, DOF locks, step overrides, broadphase layers (see parity.FEATURES)."""
    parity.single_step_parity(hostsim_api, "feature", parity.FEATURES.index(feature), 0, warm)


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
