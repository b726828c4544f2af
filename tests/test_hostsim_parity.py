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


def test_two_collision_steps(hostsim_api):
    parity.single_step_parity(hostsim_api, "pyramid", 4, 0, 20, collision_steps=2)
