"""Device queries: batched closest hit ray casts (NarrowPhaseQuery::CastRay, NarrowPhaseQuery.h:31) and AABox broadphase
queries (BroadPhaseQuery::CollideAABox, BroadPhaseQuery.h:38) against the reference on the benchmark and feature scenes."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
import refharness as R
import facade as F
from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rays_for(state, n, seed, span):
    """Rays through the scene: from a shell around the bodies towards random bodies / points (most hit something), plus rays that
    start inside bodies, axis parallel rays and zero length rays."""
    rng = np.random.default_rng(seed)
    valid = state.ids != 0xffffffff
    if not valid.any():
        valid[:] = True  # (a state read through the C ABI carries no ids: every slot of the facade scenes is in use)
    centre = state.pos[valid].mean(axis=0)
    targets = state.pos[valid][rng.integers(0, valid.sum(), n)] + rng.normal(0, 0.3, (n, 3))
    origins = centre + rng.normal(0, 1, (n, 3)) * span
    rays = np.zeros((n, 6), np.float32)
    rays[:, :3] = origins
    rays[:, 3:] = (targets - origins) * rng.uniform(0.5, 2.0, (n, 1))
    k = n // 8
    rays[:k, :3] = state.pos[valid][rng.integers(0, valid.sum(), k)]          # start inside a body
    rays[k:2 * k, 3:] *= np.array([0.0, 1.0, 0.0], np.float32)                # parallel to two axes (straight down / up)
    rays[2 * k:2 * k + 4, 3:] = 0.0                                           # zero length
    return rays


def _check_rays(ref, world, rays, layer=0xffffffff):
    want, got = ref.cast_rays(rays, layer), world.cast_rays(rays, layer)
    hit_w, hit_g = want["body"] != 0xffffffff, got["body"] != 0xffffffff
    # the same rays hit; fractions within 1e-4 relative / 1e-5 absolute (a ray grazing a surface may differ in the last ulps: allow a
    # handful of hit / miss flips only where the other side's fraction is at the very end of the ray)
    flips = np.flatnonzero(hit_w != hit_g)
    assert len(flips) <= max(2, len(rays) // 500), f"hit / miss differs for {len(flips)} of {len(rays)} rays: {flips[:8]}"
    both = hit_w & hit_g
    assert both.sum() > len(rays) // 4, "the test rays should mostly hit"
    df = np.abs(want["fraction"][both] - got["fraction"][both])
    assert np.all(df <= np.maximum(1e-5, 1e-4 * np.abs(want["fraction"][both]))), f"fractions differ by up to {df.max()}"
    same_body = want["body"][both] == got["body"][both]
    # a different body only where two bodies are hit at the same fraction within the tolerance above (bodies resting on each other:
    # a ray that grazes the contact region sees both surfaces at once)
    assert same_body.mean() > 0.95, f"closest body differs for {(~same_body).sum()} of {both.sum()} hits"
    sub_equal = want["sub_shape"][both][same_body] == got["sub_shape"][both][same_body]
    assert sub_equal.mean() > 0.98, "sub shape ids (mesh triangles) differ"
    return int(both.sum())


def _check_boxes(ref, world, state, n, seed, layer=0xffffffff):
    rng = np.random.default_rng(seed)
    valid = state.ids != 0xffffffff
    c = state.pos[valid][rng.integers(0, valid.sum(), n)] + rng.normal(0, 0.5, (n, 3))
    half = rng.uniform(0.05, 2.0, (n, 3))
    boxes = np.concatenate([c - half, c + half], axis=1).astype(np.float32)
    (wc, wi), (gc, gi) = ref.collide_aabox(boxes, layer), world.collide_aabox(boxes, layer)
    assert np.array_equal(wc, gc), f"hit counts differ for {np.flatnonzero(wc != gc)[:8]}"
    for i in range(n):
        k = min(int(wc[i]), wi.shape[1])
        if wc[i] <= wi.shape[1]:
            assert np.array_equal(np.sort(gi[i, :k]), wi[i, :k]), f"box {i}: body sets differ"
    return int(wc.sum())


QUERY_SHAPES = [(0, [0.6]), (1, [0.5, 0.3, 0.7, 0.05]), (2, [0.4, 0.3]), (3, [0.5, 0.4, 0.05]), (1, [0.4, 0.4, 0.4, 0.0])]


def _check_collide_shape(ref, world, state, n, seed, layer=0xffffffff):
    """NarrowPhaseQuery::CollideShape, all hits: the same (body, sub shape) hit sets, depths / points / axes within 1e-4 (bit equal in
    practice: the pair code is the step's)."""
    rng = np.random.default_rng(seed)
    valid = state.ids != 0xffffffff
    if not valid.any():
        valid[:] = True
    total = 0
    for qi, (kind, params) in enumerate(QUERY_SHAPES):
        q = np.zeros(n, R.SHAPE_QUERY_DTYPE)
        q["position"] = state.pos[valid][rng.integers(0, valid.sum(), n)] + rng.normal(0, 0.4, (n, 3))
        rot = rng.normal(0, 1, (n, 4))
        q["rotation"] = rot / np.linalg.norm(rot, axis=1, keepdims=True)
        q["base_offset"] = q["position"] if qi % 2 == 0 else 0.0
        sep = [0.0, 0.05, 0.3][qi % 3]
        (wc, wh), (gc, gh) = ref.collide_shape(kind, params, q, sep, layer, 96), world.collide_shape(kind, params, q, sep, layer, 96)
        assert np.array_equal(wc, gc), f"shape {qi}: hit counts differ for queries {np.flatnonzero(wc != gc)[:8]}: {wc[wc != gc][:8]} vs {gc[wc != gc][:8]}"
        for i in range(n):
            k = min(int(wc[i]), 96)
            if k == 0 or wc[i] > 96:
                continue
            a, b = wh[i, :k], gh[i, :k]
            a = a[np.lexsort((a["sub_shape2"], a["sub_shape1"], a["body"]))]
            b = b[np.lexsort((b["sub_shape2"], b["sub_shape1"], b["body"]))]
            for f in ("body", "sub_shape1", "sub_shape2"):
                assert np.array_equal(a[f], b[f]), f"shape {qi} query {i}: {f} differs"
            for f in ("penetration_depth", "point1", "point2", "axis"):
                assert np.allclose(a[f], b[f], rtol=1e-4, atol=1e-5), f"shape {qi} query {i}: {f} differs by {np.abs(a[f] - b[f]).max()}"
            total += k
    return total


def _check_volumes(ref, world, state, n, seed, layer=0xffffffff):
    rng = np.random.default_rng(seed)
    valid = state.ids != 0xffffffff
    if not valid.any():
        valid[:] = True
    c = state.pos[valid][rng.integers(0, valid.sum(), n)] + rng.normal(0, 0.5, (n, 3))
    spheres = np.concatenate([c, rng.uniform(0.0, 2.0, (n, 1))], axis=1).astype(np.float32)
    total = 0
    for mode, data in ((1, spheres), (2, c.astype(np.float32))):
        (wc, wi), (gc, gi) = ref.collide_volume(mode, data, layer), world.collide_volume(mode, data, layer)
        assert np.array_equal(wc, gc), f"mode {mode}: hit counts differ for {np.flatnonzero(wc != gc)[:8]}"
        for i in range(n):
            if wc[i] <= wi.shape[1]:
                assert np.array_equal(np.sort(gi[i, :wc[i]]), wi[i, :wc[i]]), f"mode {mode} query {i}: body sets differ"
        total += int(wc.sum())
    return total


SCENES = [("small_stack", 4, 0, 40, 6.0), ("pyramid", 6, 0, 30, 20.0), ("convex_vs_mesh", 2, 0, 60, 30.0), ("pile", 500, 15, 80, 12.0),
          ("feature", parity.FEATURES.index("zoo"), 0, 50, 60.0), ("feature", parity.FEATURES.index("decorated"), 0, 70, 10.0),
          ("convex_vs_mesh", 1, 3, 150, 30.0),  # decorated bodies on a scaled + rotated mesh
          ("feature", parity.FEATURES.index("cylinder"), 0, 70, 12.0),
          ("compound", 0, 0, 120, 12.0), ("compound", 1, 0, 150, 20.0)]


def _run(api, scene, p0, p1, warm, span, n_rays, n_boxes):
    ref = R.RefWorld(scene, p0, p1)
    for _ in range(warm):
        ref.step()
    world = ref.export(api)
    state = ref.state()
    rays = _rays_for(state, n_rays, 11, span)
    assert _check_rays(ref, world, rays) > 0
    assert _check_rays(ref, world, rays, layer=1) > 0          # cast as a MOVING body: everything collides
    assert _check_boxes(ref, world, state, n_boxes, 12) > 0
    _check_boxes(ref, world, state, n_boxes, 13, layer=0)      # as NON_MOVING: only moving bodies
    assert _check_volumes(ref, world, state, n_boxes, 14) > 0
    _check_volumes(ref, world, state, n_boxes, 15, layer=0)
    assert _check_collide_shape(ref, world, state, max(20, n_boxes // 5), 16) > 0
    _check_collide_shape(ref, world, state, max(20, n_boxes // 5), 17, layer=0)
    # after a step the trees are rebuilt for the new bounds
    world.step(); ref.step()
    assert _check_rays(ref, world, rays) > 0
    world.close()
    ref.close()


@pytest.mark.parametrize("scene,p0,p1,warm,span", SCENES)
def test_queries_hostsim(hostsim_api, scene, p0, p1, warm, span):
    _run(hostsim_api, scene, p0, p1, warm, span, 600, 100)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1,warm,span", SCENES + [("pyramid", 15, 0, 60, 40.0), ("convex_vs_mesh", 10, 0, 100, 120.0), ("pile", 20000, 15, 100, 40.0)])
def test_queries_gpu(gpu_api, ref_available, scene, p0, p1, warm, span):
    _run(gpu_api, scene, p0, p1, warm, span, 20000, 2000)


def _check_batch_rays(api, flib, n_worlds):
    proto = F.FacadeScene(flib, "pyramid", 5, 0)
    batch = api.b2j_batch_create(proto.world.h, n_worlds, 0, 0)
    assert batch, api.last_error()
    stats = _capi.StepStats()
    for _ in range(20):
        assert api.b2j_batch_step(batch, 1.0 / 60.0, 1, C.byref(stats)) == 0
        proto.world.step()
    rays = _rays_for(proto.world.state(), 500, 5, 12.0)
    want = proto.world.cast_rays(rays)
    worlds = (np.arange(len(rays)) % n_worlds).astype(np.uint32)
    got = np.zeros(len(rays), R.HIT_DTYPE)
    assert api.b2j_batch_query_cast_rays(batch, worlds.ctypes.data, rays.ctypes.data, len(rays), 0xffffffff, got.ctypes.data) == 0, api.last_error()
    assert np.array_equal(want["body"], got["body"]) and np.array_equal(want["fraction"], got["fraction"]), "every world of the batch is the prototype: same hits"
    api.b2j_batch_destroy(batch)
    proto.close()


def _check_facade_queries(flib):
    """PhysicsSystem::GetNarrowPhaseQuery().CastRay(s) / GetBroadPhaseQuery().CollideAABox through the C++ facade."""
    fs = F.FacadeScene(flib, "pile", 300, 15)
    for _ in range(60):
        fs.update()
    rays = _rays_for(fs.world.state(), 300, 3, 10.0)
    want, got = fs.world.cast_rays(rays), fs.cast_rays(rays)
    assert np.array_equal(want["body"], got["body"]) and np.array_equal(want["fraction"], got["fraction"]) and np.array_equal(want["sub_shape"], got["sub_shape"])
    box = np.array([-2, 0, -2, 2, 3, 2], np.float32)
    counts, ids = fs.world.collide_aabox(box[None, :], max_hits=256)
    assert counts[0] > 3 and np.array_equal(np.sort(ids[0, :counts[0]]), fs.collide_aabox(box))
    sphere = np.array([0.5, 1.0, -0.5, 1.5], np.float32)
    counts, ids = fs.world.collide_volume(1, sphere[None, :], max_hits=256)
    assert counts[0] > 3 and np.array_equal(np.sort(ids[0, :counts[0]]), fs.collide_sphere(sphere))
    state = fs.world.state()
    point = np.concatenate([state.pos[5], [-1.0]]).astype(np.float32)
    counts, ids = fs.world.collide_volume(2, point[None, :3], max_hits=256)
    assert counts[0] >= 1 and np.array_equal(np.sort(ids[0, :counts[0]]), fs.collide_sphere(point))
    # CollideShape: a scaled box at a body of the pile; the facade's shape (BoxShape half extent h, convex radius 0.05, scaled) is
    # the C ABI's b2j_shape_scaled(b2j_shape_box)
    he, scale = np.array([0.5, 0.4, 0.6], np.float32), np.array([1.5, 1.5, 1.5], np.float32)
    rot = np.array([0.1, 0.2, 0.3, 0.9], np.float32); rot /= np.linalg.norm(rot)
    body, sub, depth = fs.collide_shape_box(he, scale, rot, state.pos[7], 0.1)
    assert len(body) >= 1
    inner = fs.world.api.b2j_shape_box(fs.world.h, (C.c_float * 3)(*he), C.c_float(0.05))
    sid = fs.world.api.b2j_shape_scaled(fs.world.h, inner, (C.c_float * 3)(*scale))
    q = np.zeros(1, R.SHAPE_QUERY_DTYPE)
    q["shape"], q["position"], q["rotation"], q["base_offset"] = sid, state.pos[7], rot, state.pos[7]
    cnt = np.zeros(1, np.uint32); hits = np.zeros(256, R.SHAPE_HIT_DTYPE)
    assert fs.world.api.b2j_query_collide_shape(fs.world.h, q.ctypes.data, 1, 0.1, 0xffffffff, 256, cnt.ctypes.data, hits.ctypes.data) == 0
    assert cnt[0] == len(body) and np.array_equal(hits["body"][:cnt[0]], body) and np.array_equal(hits["penetration_depth"][:cnt[0]], depth)
    fs.close()


def test_facade_queries_hostsim(hostsim_api):
    _check_facade_queries(F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api))


@pytest.mark.gpu
def test_facade_queries_gpu(gpu_api):
    _check_facade_queries(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api))


def test_batch_rays_hostsim(hostsim_api):
    flib = F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api)
    _check_batch_rays(hostsim_api, flib, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("groups", [1, 3])
def test_batch_rays_gpu(gpu_api, monkeypatch, groups):
    monkeypatch.setenv("B2J_BATCH_GROUPS", str(groups))
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check_batch_rays(gpu_api, flib, 7)
