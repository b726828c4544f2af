"""Pins the plain-C oracle restatement (oracle/port.c) against outputs of the reference itself (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest

import refharness as R


def _port_pairs(port, ref):
    s = ref.state()
    n = len(s.ids)
    # harness scenes: slot 0..k static in layer 0 (NON_MOVING), rest dynamic in layer 1 (MOVING); recover from the state
    static = np.all(s.lin == 0, axis=1) & (s.active_index == 0xffffffff) & (np.arange(n) < 8)
    return s, n


@pytest.mark.parametrize("scene,p0,warm", [("pyramid", 6, 0), ("pyramid", 6, 25), ("small_stack", 4, 10), ("small_stack", 4, 90)])
def test_port_find_pairs_matches_reference(ref_available, port_lib, scene, p0, warm):
    ref = R.RefWorld(scene, p0)
    for _ in range(warm):
        ref.step()
    s = ref.state()
    n = len(s.ids)
    # layer / motion type: the harness puts static bodies in object layer 0 and everything else in layer 1 (dynamic)
    dynamic_count = ref.num_dynamic
    motion = np.where(np.arange(n) < n - dynamic_count, 0, 2).astype(np.uint8)
    layer = np.where(motion == 0, 0, 1).astype(np.uint16)
    flags = np.zeros(n, np.uint16)
    o2bp = np.array([0, 1], np.uint8)
    ovbp = np.array([0, 1, 1, 1], np.uint8)
    ovo = np.array([0, 1, 1, 1], np.uint8)
    cap = 1 << 16
    out = np.zeros((cap, 2), np.uint32)
    port_lib.port_find_pairs.restype = C.c_uint32
    cnt = port_lib.port_find_pairs(C.c_uint32(n), s.ids.ctypes.data_as(C.c_void_p), s.bounds.ctypes.data_as(C.c_void_p), motion.ctypes.data_as(C.c_void_p),
                                   layer.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p), s.active_index.ctypes.data_as(C.c_void_p),
                                   C.c_uint32(2), C.c_uint32(2), o2bp.ctypes.data_as(C.c_void_p), ovbp.ctypes.data_as(C.c_void_p), ovo.ctypes.data_as(C.c_void_p),
                                   C.c_float(0.02), out.ctypes.data_as(C.c_void_p), C.c_uint32(cap))
    got = out[:cnt]
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    want = ref.find_pairs()
    assert np.array_equal(got, want)


def test_port_free_fall_matches_reference(ref_available, port_lib):
    # bodies of the ConvexVsMesh scene fall freely for the first steps: compare one step of every dynamic body bit for bit
    ref = R.RefWorld("convex_vs_mesh", 2)
    before = ref.state()
    ref.step()
    after = ref.state()
    g = (C.c_float * 3)(0.0, -9.81, 0.0)
    zero = (C.c_float * 3)(0, 0, 0)
    for i in range(1, len(before.ids)):
        pos = (C.c_float * 3)(*before.pos[i]); rot = (C.c_float * 4)(*before.rot[i])
        lin = (C.c_float * 3)(*before.lin[i]); ang = (C.c_float * 3)(*before.ang[i])
        port_lib.port_free_body_step(pos, rot, lin, ang, g, C.c_float(1.0), C.c_float(1.0), zero, C.c_float(0.05), C.c_float(0.05),
                                     C.c_float(500.0), C.c_float(0.25 * np.pi * 60.0), C.c_float(1.0 / 60.0))
        assert np.array_equal(np.array(pos[:], np.float32), after.pos[i])
        assert np.array_equal(np.array(lin[:], np.float32), after.lin[i])
        assert np.array_equal(np.array(rot[:], np.float32), after.rot[i])


def test_port_hashes_known_answers(port_lib):
    port_lib.port_hash64.restype = C.c_uint64
    port_lib.port_hash64.argtypes = [C.c_uint64]
    port_lib.port_hash_sub_shape_id_pair.restype = C.c_uint64
    # FNV-1a 64 of 16 zero bytes and Thomas Wang's mix of 0 / 1 (computed from the published definitions)
    h = 0xcbf29ce484222325
    for _ in range(16):
        h = (h * 0x100000001b3) & 0xffffffffffffffff
    assert port_lib.port_hash_sub_shape_id_pair(0, 0, 0, 0) == h

    def wang(v):
        m = 0xffffffffffffffff
        v = (~v + (v << 21)) & m; v ^= v >> 24; v = (v + (v << 3) + (v << 8)) & m; v ^= v >> 14
        v = (v + (v << 2) + (v << 4)) & m; v ^= v >> 28; v = (v + (v << 31)) & m
        return v
    for x in (0, 1, 0x123456789abcdef):
        assert port_lib.port_hash64(x) == wang(x)


def test_reference_known_answer_tests_pass_on_the_oracle_build(ref_available):
    """Pins the oracle: the reference's OWN unit tests for the hot path -- GJK, EPA, closest point, CollideShape,
    ConvexVsTriangles, ActiveEdges, ContactListener, BroadPhase, Physics, determinism, math -- compiled against the same objects
    the parity oracle is linked from (oracle/Makefile `selftest`) must all pass."""
    import os
    import subprocess
    if not os.path.isdir("/root/reference/UnitTests"):
        pytest.skip("/root/reference absent")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["make", "-s", "-j8", "-C", os.path.join(root, "oracle"), "selftest"], capture_output=True, text=True, timeout=1800)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Status: SUCCESS!" in out.stdout
    line = [l for l in out.stdout.splitlines() if "test cases:" in l][0]
    assert " 0 failed" in line and int(line.split("|")[1].split()[0]) >= 300, line
