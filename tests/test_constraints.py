"""Non contact constraints on the step: PointConstraint / DistanceConstraint through the C ABI (b2j_constraints_*) against
the reference's ConstraintManager -- the joints feature scene (chain, rope with limits, a cloth that is one large island, kinematic tow,
a constraint that wakes a sleeping body, priorities, solver step overrides, a disabled constraint). Single step parity of the scene at
several snapshots runs with the other feature scenes (tests/test_hostsim_parity.py, tests/test_gpu_parity.py); here: long evolution,
the accumulated impulses, removal (the last constraint takes the freed index), enabling / disabling, snapshots, the error paths."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
import refharness as R
import facade as F
from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JOINTS = parity.FEATURES.index("joints")


def _same_state(ref, world, what):
    worst = R.compare_states(ref.state(), world.state(), parity.REL_TOL, parity.ABS_TOL)
    for k in ("pos", "rot", "lin", "ang"):
        assert worst[k] <= 1.0, (what, k, worst)
    assert np.array_equal(ref.state().active_index != 0xffffffff, world.state().active_index != 0xffffffff), f"{what}: active flags"
    rs, gs = ref.constraint_states(), world.constraint_states()
    assert len(rs) == len(gs), what
    assert np.allclose(rs["total_lambda"], gs["total_lambda"], rtol=1e-4, atol=1e-5), f"{what}: accumulated impulses differ by {np.abs(rs['total_lambda'] - gs['total_lambda']).max()}"
    assert np.allclose(rs["world_space_normal"], gs["world_space_normal"], rtol=1e-4, atol=1e-6), f"{what}: distance constraint normals"
    for f in ("total_lambda_rotation", "total_lambda_limits", "total_lambda_motor"):
        assert np.allclose(rs[f], gs[f], rtol=1e-4, atol=1e-5), f"{what}: hinge {f} differs by {np.abs(rs[f] - gs[f]).max()}"
    return worst


def _check_evolution_and_mutations(api, steps):
    ref = R.RefWorld("feature", JOINTS, 0)
    for _ in range(3):
        ref.step()
    world = ref.export(api)
    n = api.b2j_num_constraints(world.h)
    assert n == len(ref.constraint_states()) and n > 128, "the scene holds a cloth of > 128 constraints"
    for step in range(steps):
        ref.step(); world.step()
        if step % 25 == 24:
            _same_state(ref, world, f"step {step}")
    assert np.abs(ref.constraint_states()["total_lambda"]).max() > 0.0
    # remove constraints: a link of the chain (the last constraint takes its index), one from the middle of the cloth, the last one
    for index in (3, 60, -1):
        if index < 0:
            index = api.b2j_num_constraints(world.h) - 1
        ref.remove_constraint(index); world.remove_constraint(index)
        for _ in range(10):
            ref.step(); world.step()
        _same_state(ref, world, f"after removing constraint {index}")
    assert api.b2j_num_constraints(world.h) == n - 3
    # disable a cloth constraint and the tow bar, run, enable again
    for index in (20, 100):
        ref.set_constraint_enabled(index, False); world.set_constraint_enabled(index, False)
    for _ in range(15):
        ref.step(); world.step()
    _same_state(ref, world, "with two constraints disabled")
    for index in (20, 100):
        ref.set_constraint_enabled(index, True); world.set_constraint_enabled(index, True)
    for _ in range(15):
        ref.step(); world.step()
    _same_state(ref, world, "after enabling them again")
    world.close(); ref.close()


def _check_snapshot(api):
    ref = R.RefWorld("feature", JOINTS, 0)
    for _ in range(40):
        ref.step()
    world = ref.export(api)
    snap = api.b2j_world_save_state(world.h)
    assert snap, api.last_error()
    first = []
    for _ in range(30):
        world.step()
        first.append(world.state().pos.copy())
    lam = world.constraint_states().copy()
    world.remove_constraint(5)  # the snapshot holds the list: restoring brings the constraint back
    assert api.b2j_world_restore_state(world.h, snap) == 0, api.last_error()
    for i in range(30):
        world.step()
        assert np.array_equal(first[i], world.state().pos), f"step {i} after the restore differs"
    assert np.array_equal(lam["total_lambda"], world.constraint_states()["total_lambda"])
    api.b2j_snapshot_destroy(snap)
    world.close(); ref.close()


def _check_errors(api, flib):
    ref = R.RefWorld("feature", JOINTS, 0)
    world = ref.export(api)
    desc = np.zeros(256, np.uint8)  # a b2j_constraint_desc with body ids 0 / 0: not bodies of this world (sequence numbers are set)
    desc.view(np.uint32)[1] = 0x7fffffff
    assert api.b2j_constraints_add(world.h, desc.ctypes.data, 1) != 0 and "not a body" in api.last_error()
    bad = np.array([10 ** 6], np.uint32)
    assert api.b2j_constraints_remove(world.h, bad.ctypes.data, 1) != 0
    # a body with constraints attached cannot be removed (the reference asks for the constraints to go first); without them it can
    ids = ref.state().ids
    chain_link = np.array([ids[3]], np.uint32)  # slot 0 = floor, 1 = anchor, 2.. = the chain
    assert api.b2j_bodies_remove(world.h, chain_link.ctypes.data_as(C.POINTER(C.c_uint32)), 1) != 0 and "constraints attached" in api.last_error()
    for index in (2, 1):  # the two point constraints of the second chain link (removal moves the last constraint into the gap: higher index first)
        world.remove_constraint(index)
    assert api.b2j_bodies_remove(world.h, chain_link.ctypes.data_as(C.POINTER(C.c_uint32)), 1) == 0, api.last_error()
    err, _ = world.step()
    assert err == 0
    world.close(); ref.close()


def test_constraints_hostsim(hostsim_api):
    _check_evolution_and_mutations(hostsim_api, 150)


def test_constraints_snapshot_hostsim(hostsim_api):
    _check_snapshot(hostsim_api)


def test_constraints_errors_hostsim(hostsim_api):
    _check_errors(hostsim_api, None)


@pytest.mark.gpu
def test_constraints_gpu(gpu_api, ref_available):
    _check_evolution_and_mutations(gpu_api, 300)


@pytest.mark.gpu
def test_constraints_snapshot_gpu(gpu_api, ref_available):
    _check_snapshot(gpu_api)


@pytest.mark.gpu
def test_constraints_errors_gpu(gpu_api, ref_available):
    _check_errors(gpu_api, None)
