"""joltphysics_b200 -- B200-native rigid-body step behind Jolt's PhysicsSystem::Update surface.

The product is `libjolt_b200.so` (hand-written CUDA for sm_100a behind the C ABI of include/jolt_b200.h); this package
only binds it (ctypes). There is no CPU fallback: `load()` raises when the library is missing.
"""
from ._capi import CApi, load, library_path  # noqa: F401
