"""Golden fixtures (tests/golden/*.npz, made from the reference by tests/golden/make_golden.py): the benchmark scenes built through
the facade and stepped by the CUDA path must reproduce the reference's states, candidate pair sets and contact point counts
without the reference being present at run time (it does not exist on the GPU box); the fixtures themselves are pinned to the
oracle by test_golden_is_what_the_oracle_produces."""
import glob
import importlib.util
import os

import numpy as np
import pytest

import refharness as R
import facade as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _case(path):
    scene, p0, p1 = os.path.basename(path)[:-4].rsplit("_", 2)
    return scene, int(p0), int(p1), np.load(path)


def _snapshot(world):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.snapshot(world)


def _check_against_golden(step_fn, world, gold, exact):
    done = 0
    for k in gold["steps"]:
        while done < k:
            step_fn()
            done += 1
        got = _snapshot(world)
        want = {name: gold[f"s{k}_{name}"] for name in got}
        assert np.array_equal(want["pairs"], got["pairs"]), f"step {k}: candidate pair set differs"
        assert np.array_equal(want["cache"], got["cache"]), f"step {k}: manifolds per pair / points per manifold differ"
        assert np.array_equal(want["active"], got["active"]), f"step {k}: active flags differ"
        for name in ("pos", "rot", "lin", "ang"):
            if exact:
                assert np.array_equal(want[name], got[name]), f"step {k}: {name} is not bit identical"
            else:
                # north star tolerance: 1e-4 relative or 1e-5 absolute per body
                err = np.linalg.norm(want[name] - got[name], axis=1)
                tol = np.maximum(1e-5, 1e-4 * np.linalg.norm(want[name], axis=1))
                assert np.all(err <= tol), f"step {k}: {name} off by {float(np.max(err / tol)):.2f} tolerances"


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_is_what_the_oracle_produces(path, ref_available):
    scene, p0, p1, gold = _case(path)
    ref = R.RefWorld(scene, p0, p1)
    _check_against_golden(ref.step, ref, gold, exact=True)
    ref.close()


@pytest.fixture(scope="session")
def hostsim_facade(hostsim_api):
    return F.FacadeLib(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_facade_hostsim.so"), hostsim_api)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_hostsim(path, hostsim_facade):
    scene, p0, p1, gold = _case(path)
    fs = F.FacadeScene(hostsim_facade, scene, p0, p1)
    _check_against_golden(fs.update, fs.world, gold, exact=False)
    fs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_gpu(path, gpu_api):
    """No reference at run time: facade scene -> C ABI -> CUDA kernels vs the committed reference states."""
    scene, p0, p1, gold = _case(path)
    fs = F.FacadeScene(F.FacadeLib(os.path.join(ROOT, "joltphysics_b200", "libjolt_b200_facade.so"), gpu_api), scene, p0, p1)
    _check_against_golden(fs.update, fs.world, gold, exact=False)
    fs.close()
