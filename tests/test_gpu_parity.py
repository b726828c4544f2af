"""Parity tests proper: the CUDA path through the C ABI of libjolt_b200.so against the reference oracle (oracle/_ref)."""
import pytest

import parity

pytestmark = pytest.mark.gpu

CASES = [
    ("pyramid", 4, 0, 0), ("pyramid", 4, 0, 40),
    ("pyramid", 15, 0, 0), ("pyramid", 15, 0, 1), ("pyramid", 15, 0, 30), ("pyramid", 15, 0, 120), ("pyramid", 15, 0, 300),
    ("small_stack", 0, 0, 30), ("small_stack", 1, 0, 30), ("small_stack", 2, 0, 30), ("small_stack", 3, 0, 30),
    ("small_stack", 4, 0, 5), ("small_stack", 4, 0, 45), ("small_stack", 4, 0, 200),
    ("convex_vs_mesh", 10, 0, 60), ("pile", 2000, 15, 100),
    # ScaledShape / RotatedTranslatedShape decorated convex bodies landing on the mesh (p1 = 1)
    ("convex_vs_mesh", 4, 1, 90), ("convex_vs_mesh", 4, 1, 120), ("convex_vs_mesh", 4, 1, 200),
    # ... and with the mesh itself scaled + rotated (p1 bit 1)
    ("convex_vs_mesh", 4, 2, 200), ("convex_vs_mesh", 4, 3, 150), ("convex_vs_mesh", 4, 3, 300),
    # StaticCompoundShape bodies against a floor / each other / a static compound staircase (p0 = 0) and on a terrain mesh (p0 = 1)
    ("compound", 0, 0, 0), ("compound", 0, 0, 60), ("compound", 0, 0, 90), ("compound", 0, 0, 120), ("compound", 0, 0, 300),
    ("compound", 1, 0, 80), ("compound", 1, 0, 120), ("compound", 1, 0, 200), ("compound", 1, 0, 450),
    # worlds with more than 4096 bodies: the wavefront schedule runs as one cooperative launch (sched_grid_kernel)
    ("pile", 6000, 15, 80), ("max_bodies", 6000, 0, 10),
]


@pytest.mark.parametrize("scene,p0,p1,warm", CASES)
def test_single_step_parity(gpu_api, ref_available, scene, p0, p1, warm):
    out = parity.single_step_parity(gpu_api, scene, p0, p1, warm)
    assert out["stats"]["kernel_launches"] > 0


FEATURE_CASES = [(name, warm) for name in parity.FEATURES for warm in ((0, 10, 30, 60, 120, 250) if name == "zoo" else (1, 30, 60, 120))]


@pytest.mark.parametrize("feature,warm", FEATURE_CASES, ids=[f"{n}-step{w}" for n, w in FEATURE_CASES])
def test_feature_single_step_parity(gpu_api, ref_available, feature, warm):
    """Motion types (kinematic movers / platforms), B2J_BODY_SENSOR (static, kinematic and dynamic sensors, the kinematic vs sensor
    pair rule), allowed DOFs (Plane2D, translation only, rotation only), B2J_BODY_GYROSCOPIC, per body velocity / position step
    overrides, B2J_BODY_USE_MANIFOLD_REDUCTION off, two moving broadphase layers, B2J_BODY_KIN_VS_NONDYN -- and all in one world."""
    out = parity.single_step_parity(gpu_api, "feature", parity.FEATURES.index(feature), 0, warm)
    assert out["stats"]["kernel_launches"] > 0


@pytest.mark.parametrize("slot", [1, 5, 24])
def test_body_recreated_in_the_same_slot_is_not_served_from_the_cache(gpu_api, ref_available, slot):
    """a sphere created in the slot of a destroyed resting box (next sequence number, same pose) collides afresh."""
    out = parity.single_step_parity(gpu_api, "small_stack", 1, 0, 60, before_export=lambda ref: ref.replace_body(slot))
    assert out["manifolds"] > 0


def test_feature_zoo_multi_step(gpu_api, ref_available):
    history = parity.multi_step_drift(gpu_api, "feature", parity.FEATURES.index("zoo"), 0, 0, steps=150)
    for k in ("pos", "rot", "lin", "ang"):
        assert max(h[k] for h in history) <= 1.0, (k, [h[k] for h in history])


def test_two_collision_steps(gpu_api, ref_available):
    parity.single_step_parity(gpu_api, "pyramid", 4, 0, 20, collision_steps=2)
    parity.single_step_parity(gpu_api, "feature", parity.FEATURES.index("zoo"), 0, 30, collision_steps=3)


@pytest.mark.parametrize("bodies,warm", [(100000, 40), (100000, 90)], ids=["pile100k-step40", "pile100k-step90"])
def test_pile_100k_single_step_parity(gpu_api, ref_available, bodies, warm):
    """One snapshot step of a 100 000 body pile (the size the bench's CPU leg steps): the cooperative grid scheduler, the ordered
    convex pair queue and every counter / queue at a size where they actually fill up."""
    out = parity.single_step_parity(gpu_api, "pile", bodies, 15, warm, warm_threads=0, check_events=False)
    assert out["pairs"] > 100000 and out["stats"]["num_constraints"] > 10000, out["stats"]


def test_pyramid_long_run_energy_and_heights(gpu_api, ref_available):
    # long runs: stacking chaos prevents trajectory identity -> compare resting heights and kinetic energy
    import numpy as np
    import refharness as R
    ref = R.RefWorld("pyramid", 8)
    world = ref.export(gpu_api)
    for _ in range(240):
        ref.step(); world.step()
    rs, gs = ref.state(), world.state()
    assert np.allclose(np.sort(rs.pos[1:, 1]), np.sort(gs.pos[1:, 1]), atol=2e-2)
    assert abs(ref.kinetic_energy()) < 1.0
    ke = 0.5 * np.sum(gs.lin[1:] ** 2) * 8000.0  # box mass 2x2x2 * 1000
    assert ke < 1.0


@pytest.mark.parametrize("scene,p0,p1,warm", [("pile", 2000, 15, 60), ("pile", 2000, 15, 140), ("convex_vs_mesh", 6, 0, 50)])
def test_single_world_ordered_pair_queue(gpu_api, ref_available, monkeypatch, scene, p0, p1, warm):
    """Big single worlds process the convex pair queue grouped by shape type pair (1M body pile); the threshold is lowered here so
    that the small parity scenes take that path too."""
    monkeypatch.setenv("B2J_COLLIDE_ORDER_MIN", "8")
    # (the ordering buffers exist from the second step of a world on, so the check steps both sides for a while)
    history = parity.multi_step_drift(gpu_api, scene, p0, p1, warm, steps=40)
    for k in ("pos", "rot", "lin", "ang"):
        assert max(h[k] for h in history) <= 1.0, (k, [h[k] for h in history])
