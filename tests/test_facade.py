"""The C++ facade (PhysicsSystem / BodyInterface mirror) builds the benchmark scenes itself: body ids, mass properties, bounds
and poses must be identical to what the reference creates, and stepping through the facade must track the reference."""
import os

import numpy as np
import pytest

import refharness as R
import facade as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = [("pyramid", 5, 0), ("convex_vs_mesh", 2, 0), ("pile", 600, 15), ("max_bodies", 300, 0)]


def _check(flib, scene, p0, p1, steps):
    ref = R.RefWorld(scene, p0, p1)
    fs = F.FacadeScene(flib, scene, p0, p1)
    assert fs.num_bodies == ref.num_bodies and fs.num_dynamic == ref.num_dynamic
    rs, gs = ref.state(), fs.world.state()
    assert np.array_equal(rs.pos, gs.pos) and np.array_equal(rs.rot, gs.rot)
    assert np.array_equal(rs.bounds, gs.bounds), "world bounds computed on the device differ from the reference's"
    # Same active SET. The order can differ for bulk adds: the reference activates in the order its quad tree partitioning
    # shuffled the ids into (BodyInterface::AddBodiesPrepare), which is internal to its broadphase; the facade keeps argument order.
    assert np.array_equal(np.sort(ref.active_bodies()), np.sort(fs.world.active_bodies()))
    for _ in range(steps):
        ref.step()
        err, _ = fs.update()
        assert err == 0
    worst = R.compare_states(ref.state(), fs.world.state())
    for k in ("pos", "rot", "lin", "ang"):
        assert worst[k] <= 1.0, (k, worst)
    fs.close()
    ref.close()


@pytest.fixture(scope="session")
def hostsim_facade(hostsim_api):
    return F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api)


@pytest.mark.parametrize("scene,p0,p1", SCENES)
def test_facade_scene_matches_reference_hostsim(hostsim_facade, scene, p0, p1):
    _check(hostsim_facade, scene, p0, p1, steps=25)


@pytest.mark.gpu
@pytest.mark.parametrize("scene,p0,p1", SCENES + [("pyramid", 15, 0), ("convex_vs_mesh", 10, 0)])
def test_facade_scene_matches_reference_gpu(gpu_api, ref_available, scene, p0, p1):
    flib = F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api)
    _check(flib, scene, p0, p1, steps=25)


def _check_feature(flib, variant, steps, scene="feature"):
    """feature_scenes.inl is ONE piece of user code compiled against the reference and against the facade: every motion type, body flag
    and per body override travels BodyCreationSettings -> facade -> C ABI -> device and must give the world the reference builds
    (creation state incl. DOF restricted mass properties) and the evolution the reference computes."""
    ref = R.RefWorld(scene, variant)
    fs = F.FacadeScene(flib, scene, variant, 0)
    assert fs.num_bodies == ref.num_bodies and fs.num_dynamic == ref.num_dynamic
    rs, gs = ref.state(), fs.world.state()
    assert np.array_equal(rs.pos, gs.pos) and np.array_equal(rs.rot, gs.rot) and np.array_equal(rs.lin, gs.lin) and np.array_equal(rs.ang, gs.ang)
    assert np.array_equal(rs.bounds, gs.bounds)
    assert np.array_equal(np.sort(ref.active_bodies()), np.sort(fs.world.active_bodies()))
    for step in range(steps):
        ref.step()
        err, _ = fs.update()
        assert err == 0
        if step % 20 == 19:
            worst = R.compare_states(ref.state(), fs.world.state())
            for k in ("pos", "rot", "lin", "ang"):
                assert worst[k] <= 1.0, (step, k, worst)
            assert np.array_equal(ref.state().active_index != 0xffffffff, fs.world.state().active_index != 0xffffffff), f"step {step}: active flags"
            # the contact caches hold the same manifolds: body pairs, sub shape ids of both sides, point counts
            assert np.array_equal(R.cache_rows(R.cache_summary(*ref.cache())), R.cache_rows(R.cache_summary(*fs.world.cache()))), f"step {step}: contact cache"
    fs.close()
    ref.close()


FEATURE_IDS = ["kinematic", "sensor", "dof_plane2d", "gyroscopic", "step_overrides", "no_manifold_reduction", "two_moving_layers", "kinematic_vs_nondynamic", "zoo", "decorated", "cylinder", "joints"]


@pytest.mark.parametrize("variant", range(12), ids=FEATURE_IDS)
def test_facade_feature_scene_matches_reference_hostsim(hostsim_facade, variant):
    _check_feature(hostsim_facade, variant, 120)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", range(12), ids=FEATURE_IDS)
def test_facade_feature_scene_matches_reference_gpu(gpu_api, ref_available, variant):
    _check_feature(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api), variant, 200)


def test_facade_compound_scene_matches_reference_hostsim(hostsim_facade):
    """StaticCompoundShapeSettings::Create in the facade builds what the reference builds: centre of mass, sub shape transforms, mass
    properties and the quad tree (same node order -> same manifolds), joltphysics_b200/host/compound_scene.inl on both sides."""
    _check_feature(hostsim_facade, 0, 160, scene="compound")


@pytest.mark.gpu
def test_facade_compound_scene_matches_reference_gpu(gpu_api, ref_available):
    _check_feature(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api), 0, 300, scene="compound")


def _check_api_tour(flib):
    """The BodyInterface / PhysicsSystem surface of SURVEY 8(b) beyond scene building (joltphysics_b200/host/api_tour.inl): the same
    user code runs against the reference and against the facade; states and query results must agree after every phase."""
    ref = R.RefWorld("api_tour")
    fs = F.FacadeScene(flib, "api_tour")

    def compare(tag, exact=False):
        fs.world.n = 18  # body slots 0..16 are used by the tour (13 at creation, 4 added later, one slot reused)
        rs, gs = ref.state(18), fs.world.state()
        worst = R.compare_states(rs, gs)
        for k in ("pos", "rot", "lin", "ang"):
            assert worst[k] <= 1.0, (tag, k, worst)
        if exact:
            # every step of the tour is bit identical, sleep timing included (EActivation::Activate on a body that is awake resets its
            # sleep timer, BodyInterface::ActivateBodyInternal): checked on the bodies that are in the world
            in_world = np.array([bool(x) for x in fs.world.state().active_index != 0xffffffff]) | (rs.active_index != 0xffffffff)
            for name in ("pos", "rot", "lin", "ang"):
                a, b = getattr(rs, name)[in_world], getattr(gs, name)[in_world]
                assert np.array_equal(a, b), (tag, name, np.abs(a - b).max())
            assert np.array_equal(rs.active_index != 0xffffffff, gs.active_index != 0xffffffff), (tag, "active flags")
        (ra, rn, rf), (ga, gn, gf) = ref.query(), fs.query()
        assert rn == gn, (tag, "GetBodies", rn, gn)
        assert np.array_equal(ra, ga), (tag, "GetActiveBodies", ra, ga)
        assert rf == gf, (tag, "IsAdded && IsActive", bin(rf), bin(gf))

    compare("created")
    for phase, steps in ((0, 10), (1, 25), (2, 40), (3, 30), (4, 60), (5, 50)):
        if phase:
            ref.mutate(phase)
            fs.mutate(phase)
            compare(f"after mutation {phase}")
        for step in range(steps):
            ref.step()
            err, _ = fs.update()
            assert err == 0
            compare(f"phase {phase} step {step}", exact=True)
        compare(f"after the steps of phase {phase}")
    # the BodyInterface getters (flat array fast path / Body mirror) agree with the device state for every tour body
    ref.step()
    n_dyn = fs.flib.lib.b2jf_scene_num_dynamic(fs.h)
    out = np.zeros((n_dyn, 3), np.float32)
    assert fs.step_e2e(1.0 / 60.0, None, out) == 0
    ids, rs = ref.state(18).ids, ref.state(18)
    fs.world.n = 18
    gs = fs.world.state()
    got = {tuple(np.round(p, 6)) for p in out}
    want = {tuple(np.round(gs.pos[i], 6)) for i in range(1, 18) if ids[i] != 0xffffffff}
    assert got == want, "GetCenterOfMassPosition through the facade differs from the device state"
    assert R.compare_states(rs, gs)["pos"] <= 1.0
    fs.close()
    ref.close()


def _check_incremental_download(flib):
    """once bodies sleep, Update mirrors only the bodies the step simulated; the getters still agree with the device for
    EVERY body (sleeping ones keep the rows of an earlier download)."""
    fs = F.FacadeScene(flib, "convex_vs_mesh", 2, 0)       # 100 bodies on the terrain mesh (slot 0): they settle and sleep one by one
    n_dyn = fs.flib.lib.b2jf_scene_num_dynamic(fs.h)
    out = np.zeros((n_dyn, 3), np.float32)
    counts = []
    for step in range(900):
        assert fs.step_e2e(1.0 / 60.0, None, out) == 0     # Update + GetCenterOfMassPosition of every dynamic body
        counts.append(fs.flib.lib.b2jf_scene_last_download_count(fs.h))
        if step % 50 == 49 or step > 890:
            gs = fs.world.state()
            assert np.array_equal(out, gs.pos[1:1 + n_dyn]), f"step {step}: getters differ from the device state"
    active = fs.flib.api.b2j_num_active_bodies(fs.world.h)
    assert counts[0] == fs.num_bodies, "the first refresh is a full download"
    assert active < n_dyn // 2, "most bodies should have gone to sleep"
    assert counts[-1] <= max(2 * active + 8, 8), f"only the simulated bodies are mirrored once the scene sleeps: {counts[-1]} rows for {active} active bodies"
    fs.close()


def test_facade_incremental_download_hostsim(hostsim_facade):
    _check_incremental_download(hostsim_facade)


@pytest.mark.gpu
def test_facade_incremental_download_gpu(gpu_api):
    _check_incremental_download(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api))


def test_facade_api_tour_hostsim(hostsim_facade):
    _check_api_tour(hostsim_facade)


@pytest.mark.gpu
def test_facade_api_tour_gpu(gpu_api, ref_available):
    _check_api_tour(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api))


def _check_overflow(flib, limit):
    """Error behaviour: with too small body pair / contact constraint limits Update must keep running and report the
    overflow through the reference's EPhysicsUpdateError bits (which contacts survive an overflow is unspecified)."""
    ref = R.RefWorld("pyramid_tight", 6, limit)
    fs = F.FacadeScene(flib, "pyramid_tight", 6, limit)
    seen_ref = seen_got = 0
    for _ in range(25):
        seen_ref |= ref.step()
        err, _ = fs.update()
        assert err >= 0, "the step itself must not fail"
        seen_got |= err
    assert seen_ref != 0, "the limits are meant to overflow"
    # The reference allocates body pairs and manifolds from ONE byte arena (ContactConstraintManager.cpp:299-306), so one
    # exhausted arena raises several bits there; the device has separate arrays and reports the limit that actually overflowed:
    # a non-empty subset of the reference's bits.
    assert seen_got != 0 and (seen_got & ~seen_ref) == 0, (bin(seen_got), bin(seen_ref))
    fs.close()
    ref.close()


@pytest.mark.parametrize("limit", [64, 200])
def test_facade_overflow_errors_hostsim(hostsim_facade, limit):
    _check_overflow(hostsim_facade, limit)


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [64, 200])
def test_facade_overflow_errors_gpu(gpu_api, ref_available, limit):
    _check_overflow(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api), limit)
