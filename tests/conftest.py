import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_available():
    import refharness
    if not refharness.have_ref("det"):
        if os.path.isdir("/root/reference/Jolt"):
            subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "det"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return True


@pytest.fixture(scope="session")
def hostsim_api(ref_available):
    """Kernel bodies compiled for the host (debug aid, tests/hostsim) behind the same C ABI."""
    from joltphysics_b200 import _capi
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostsim")])
    return _capi.CApi(os.path.join(ROOT, "tests", "hostsim", "_build", "libb2j_hostsim.so"))


@pytest.fixture(scope="session")
def gpu_api():
    """The product: libjolt_b200.so (CUDA). No fallback: fails if the library is missing. (Tests that compare against the live
    reference also request `ref_available`; the golden / batch tests run without it.)"""
    import joltphysics_b200
    return joltphysics_b200.load()


@pytest.fixture(scope="session")
def port_lib():
    import ctypes
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.port"])
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "_port", "libb2j_port.so"))
