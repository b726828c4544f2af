"""Host-side checks that need no GPU: the product library loads and exports every symbol include/jolt_b200.h declares."""
import os
import re

import pytest

from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jolt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2j_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_capi.PROTOTYPES)


def test_product_library_exports_every_symbol():
    path = _capi.library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build_product()
    api = _capi.CApi(path)  # resolves every prototype, raises on a missing symbol
    for name in _declared_symbols():
        assert hasattr(api.lib, name)


def test_world_create_fails_loudly_without_cuda():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("CUDA present")
    except ImportError:
        pass
    import ctypes as C
    api = _capi.CApi(_capi.library_path())
    desc = _capi.WorldDesc()
    desc.max_bodies, desc.max_body_pairs, desc.max_contact_constraints = 16, 16, 16
    desc.num_object_layers = desc.num_broadphase_layers = 1
    t = (C.c_uint8 * 1)(0); t1 = (C.c_uint8 * 1)(1)
    desc.object_to_broadphase, desc.object_vs_broadphase, desc.object_vs_object = t, t1, t1
    api.b2j_settings_default(C.byref(desc.settings))
    assert not api.b2j_world_create(C.byref(desc))
    assert "no CPU fallback" in api.last_error()


def test_struct_sizes_match_header():
    # sizes the C side static_asserts (see jolt_b200.cu) -- a drifted ctypes layout would corrupt every call
    import ctypes as C
    assert C.sizeof(_capi.Settings) == 64
    assert C.sizeof(_capi.CachedBodyPair) == 40
    assert C.sizeof(_capi.CachedManifold) == 152
    assert C.sizeof(_capi.ContactEvent) == 148
    assert C.sizeof(_capi.BodyDesc) == 232


def test_query_and_constraint_struct_layouts_match_the_header(tmp_path):
    """The numpy views the tests use for b2j_shape_query / b2j_collide_shape_hit / b2j_constraint_state have the sizes the C header
    declares (compiled with the system gcc), and b2j_constraint_desc has the layout the tests' raw buffers assume."""
    import subprocess
    import refharness as R
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jolt_b200.h"\nint main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(b2j_shape_query), '
                   'sizeof(b2j_collide_shape_hit), sizeof(b2j_constraint_state), sizeof(b2j_constraint_desc), offsetof(b2j_constraint_desc, hinge_axis1)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes[0] == R.SHAPE_QUERY_DTYPE.itemsize and sizes[1] == R.SHAPE_HIT_DTYPE.itemsize and sizes[2] == R.CONSTRAINT_STATE_DTYPE.itemsize
    assert sizes[3] == 104 and sizes[4] == 52
