"""Shared single-step parity protocol: from one Jolt-created world at step k -> (i) candidate pair set equality,
(ii) per pair manifold / contact point count equality, (iii) one step on both -> positions, rotations, velocities within
max(1e-5, 1e-4 * |x|), (iv) active flags equal. `api` is any binding of the jolt_b200.h ABI."""
import numpy as np

import refharness as R

REL_TOL = 1e-4   # north star: 1e-4 relative
ABS_TOL = 1e-5   # or 1e-5 m absolute


FEATURES = ["kinematic", "sensor", "dof_plane2d", "gyroscopic", "step_overrides", "no_manifold_reduction", "two_moving_layers", "kinematic_vs_nondynamic", "zoo", "decorated", "cylinder", "joints"]
"""Variants of the `feature` scene of oracle/ref_harness.cpp (sSceneFeature): motion types, body flags (B2J_BODY_SENSOR,
B2J_BODY_GYROSCOPIC, B2J_BODY_KIN_VS_NONDYN, B2J_BODY_USE_MANIFOLD_REDUCTION off, B2J_BODY_ALLOW_SLEEPING), allowed DOFs, per body
solver step overrides, two moving broadphase layers; `zoo` = all of them in one world; `decorated` = ScaledShape / RotatedTranslatedShape
around every convex leaf type (SURVEY 8 f4); `cylinder` = CylinderShape plain / scaled / rotated against the other convex shapes; `joints` = PointConstraint / DistanceConstraint
(chain, rope with limits, a cloth that forms one large island, kinematic tow, constraint that wakes a sleeping body, priorities,
solver step overrides, a disabled constraint)."""


def single_step_parity(api, scene, p0=0, p1=0, warm=0, dt=1.0 / 60.0, collision_steps=1, check_events=True, warm_threads=1, before_export=None):
    ref = R.RefWorld(scene, p0, p1)
    for _ in range(warm):
        ref.step(dt, 1, warm_threads)  # (the deterministic build gives the same state for any thread count)
    if before_export is not None:
        before_export(ref)
    world = ref.export(api)
    out = {}
    # (i) broadphase candidate pairs of the snapshot
    rp, gp = ref.find_pairs(), world.find_pairs()
    assert np.array_equal(rp, gp), f"broadphase pair sets differ: ref {len(rp)} got {len(gp)}"
    out["pairs"] = len(rp)
    err, stats = world.step(dt, collision_steps)
    ref_err = ref.step(dt, collision_steps, warm_threads)
    assert err == ref_err, (err, ref_err)
    # (ii) body pairs processed (every candidate pair gets a cache entry) and manifolds / contact point counts
    rc, gc = R.cache_summary(*ref.cache()), R.cache_summary(*world.cache())
    assert set(rc) == set(gc), f"cached body pair sets differ: ref {len(rc)} got {len(gc)}"
    bad = [k for k in rc if sorted(rc[k]) != sorted(gc[k])]
    assert not bad, f"manifold / contact point counts differ for {len(bad)} pairs, e.g. {bad[:3]}: {[(rc[k], gc[k]) for k in bad[:3]]}"
    out["manifolds"] = sum(len(v) for v in rc.values())
    # (iii) state after one step
    rs, gs = ref.state(), world.state()
    worst = R.compare_states(rs, gs, REL_TOL, ABS_TOL)
    for k in ("pos", "rot", "lin", "ang"):
        assert worst[k] <= 1.0, f"{k} out of tolerance: {worst}"
    # bounds are what the next broadphase sees
    assert np.allclose(rs.bounds[rs.ids != 0xffffffff], gs.bounds[rs.ids != 0xffffffff], rtol=REL_TOL, atol=ABS_TOL)
    # (iv) active flags
    assert np.array_equal(rs.active_index != 0xffffffff, gs.active_index != 0xffffffff), "active flags differ"
    if check_events:
        re_, ge = ref.contact_events(), world.contact_events()
        key = lambda e: (e.kind, e.body1, e.body2, e.sub_shape1, e.sub_shape2, e.num_points)
        assert sorted(map(key, re_)) == sorted(map(key, ge)), f"contact events differ: ref {len(re_)} got {len(ge)}"
        assert sorted(ref.activation_events()) == sorted(world.activation_events()), "activation events differ"
    # the solve schedule of the step: every phase touches disjoint dynamic bodies
    assert api.b2j_debug_check_schedule(world.h) == 0, "a dynamic body is touched by two constraints of the same phase"
    out["worst"] = worst
    out["stats"] = stats.as_dict()
    world.close()
    ref.close()
    return out


def multi_step_drift(api, scene, p0=0, p1=0, warm=0, steps=60, dt=1.0 / 60.0):
    """Both sides step independently from one snapshot; returns the worst tolerance ratio seen per step."""
    ref = R.RefWorld(scene, p0, p1)
    for _ in range(warm):
        ref.step(dt)
    world = ref.export(api)
    history = []
    for _ in range(steps):
        world.step(dt)
        ref.step(dt)
        history.append(R.compare_states(ref.state(), world.state(), REL_TOL, ABS_TOL))
    world.close()
    ref.close()
    return history
