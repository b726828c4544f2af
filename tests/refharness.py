"""ctypes wrapper of oracle/_ref/libjoltref_{det,fast}.so (the UNMODIFIED reference + oracle/ref_harness.cpp).

Test infrastructure only. `RefWorld` builds a reference scene, steps it, and can re-create its current state inside a
b2j world through the C ABI (jref_export_to_b2j), which is how every parity test gets "an identical snapshot of a
Jolt-created world".
"""
import ctypes as C
import os

import numpy as np

from joltphysics_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def ref_lib_path(variant="det"):
    return os.path.join(REF_DIR, f"libjoltref_{variant}.so")


def have_ref(variant="det"):
    return os.path.exists(ref_lib_path(variant))


_libs = {}


def ref_lib(variant="det"):
    if variant not in _libs:
        L = C.CDLL(ref_lib_path(variant), mode=C.RTLD_LOCAL)
        vp, u32p, fp = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float)
        L.jref_last_error.restype = C.c_char_p
        L.jref_bind_b2j.argtypes = [C.c_char_p]
        L.jref_create_scene.restype = vp
        L.jref_create_scene.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.jref_destroy.argtypes = [vp]
        L.jref_set_recording.argtypes = [vp, C.c_int]
        L.jref_mutate.argtypes = [vp, C.c_int]
        L.jref_cast_rays.argtypes = [vp, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.jref_collide_aabox.argtypes = [vp, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.jref_collide_shape.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.jref_collide_volume.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.jref_num_constraints.restype = C.c_uint32
        L.jref_num_constraints.argtypes = [vp]
        L.jref_get_constraint_states.restype = C.c_uint32
        L.jref_get_constraint_states.argtypes = [vp, C.c_void_p, C.c_uint32]
        L.jref_remove_constraint.argtypes = [vp, C.c_uint32]
        L.jref_set_constraint_enabled.argtypes = [vp, C.c_uint32, C.c_int]
        L.jref_replace_body.restype = C.c_uint32
        L.jref_replace_body.argtypes = [vp, C.c_uint32, C.c_float]
        L.jref_query.argtypes = [vp, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.jref_step.argtypes = [vp, C.c_float, C.c_int, C.c_int]
        L.jref_time_steps.restype = C.c_double
        L.jref_time_steps.argtypes = [vp, C.c_float, C.c_int, C.c_int]
        for f in ("jref_num_bodies", "jref_num_dynamic", "jref_num_active", "jref_max_bodies"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [vp]
        L.jref_get_state.argtypes = [vp, C.c_uint32, u32p, fp, fp, fp, fp, fp, u32p, fp]
        L.jref_find_pairs.restype = C.c_uint32
        L.jref_find_pairs.argtypes = [vp, u32p, C.c_uint32]
        L.jref_get_cache.restype = C.c_uint32
        L.jref_get_cache.argtypes = [vp, C.POINTER(_capi.CachedBodyPair), C.c_uint32, C.POINTER(_capi.CachedManifold), C.c_uint32, u32p]
        L.jref_get_contact_events.restype = C.c_uint32
        L.jref_get_contact_events.argtypes = [vp, C.POINTER(_capi.ContactEvent), C.c_uint32]
        L.jref_get_activation_events.restype = C.c_uint32
        L.jref_get_activation_events.argtypes = [vp, C.POINTER(_capi.ActivationEvent), C.c_uint32]
        L.jref_get_active_bodies.restype = C.c_uint32
        L.jref_get_active_bodies.argtypes = [vp, u32p, C.c_uint32]
        L.jref_export_to_b2j.restype = vp
        L.jref_export_to_b2j.argtypes = [vp, C.c_int]
        L.jref_get_settings.argtypes = [vp, C.POINTER(_capi.Settings)]
        L.jref_kinetic_energy.restype = C.c_double
        L.jref_kinetic_energy.argtypes = [vp]
        _libs[variant] = L
    return _libs[variant]


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


HIT_DTYPE = np.dtype([("body", np.uint32), ("sub_shape", np.uint32), ("fraction", np.float32)])
# b2j_shape_query / b2j_collide_shape_hit (include/jolt_b200.h)
CONSTRAINT_STATE_DTYPE = np.dtype([("total_lambda", np.float32, 3), ("world_space_normal", np.float32, 3), ("total_lambda_rotation", np.float32, 3),
                                   ("total_lambda_limits", np.float32), ("total_lambda_motor", np.float32)])  # b2j_constraint_state
SHAPE_QUERY_DTYPE = np.dtype([("shape", np.int32), ("position", np.float32, 3), ("rotation", np.float32, 4), ("base_offset", np.float32, 3)])
SHAPE_HIT_DTYPE = np.dtype([("body", np.uint32), ("sub_shape1", np.uint32), ("sub_shape2", np.uint32), ("penetration_depth", np.float32),
                            ("point1", np.float32, 3), ("point2", np.float32, 3), ("axis", np.float32, 3)])


class State:
    """Body state by slot."""

    def __init__(self, n):
        self.ids = np.full(n, 0xffffffff, np.uint32)
        self.pos = np.zeros((n, 3), np.float32)
        self.rot = np.zeros((n, 4), np.float32)
        self.lin = np.zeros((n, 3), np.float32)
        self.ang = np.zeros((n, 3), np.float32)
        self.bounds = np.zeros((n, 6), np.float32)
        self.active_index = np.full(n, 0xffffffff, np.uint32)
        self.sleep_timer = np.zeros(n, np.float32)


class RefWorld:
    def __init__(self, scene, p0=0, p1=0, variant="det"):
        self.L = ref_lib(variant)
        self.h = self.L.jref_create_scene(scene.encode(), p0, p1)
        if not self.h:
            raise RuntimeError("jref_create_scene failed: " + self.L.jref_last_error().decode())

    def close(self):
        if self.h:
            self.L.jref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self, dt=1.0 / 60.0, collision_steps=1, threads=1):
        return self.L.jref_step(self.h, dt, collision_steps, threads)

    def time_steps(self, n, dt=1.0 / 60.0, threads=0):
        return self.L.jref_time_steps(self.h, dt, n, threads)

    def mutate(self, phase):
        self.L.jref_mutate(self.h, phase)

    def query(self):
        """api_tour.inl sApiTourQuery: (sorted active ids, number of bodies, active flags of the tour bodies)."""
        ids = np.zeros(4096, np.uint32)
        nb, flags = C.c_uint32(), C.c_uint32()
        n = self.L.jref_query(self.h, ids.ctypes.data, len(ids), C.addressof(nb), C.addressof(flags))
        return ids[:n].copy(), nb.value, flags.value

    def cast_rays(self, rays, object_layer=0xffffffff):
        """rays: float32 [n][6] (origin, direction) -> structured array of hits (body, sub_shape, fraction)."""
        rays = np.ascontiguousarray(rays, np.float32)
        hits = np.zeros(len(rays), HIT_DTYPE)
        self.L.jref_cast_rays(self.h, rays.ctypes.data, len(rays), object_layer, hits.ctypes.data)
        return hits

    def collide_aabox(self, boxes, object_layer=0xffffffff, max_hits=64):
        boxes = np.ascontiguousarray(boxes, np.float32)
        counts = np.zeros(len(boxes), np.uint32)
        ids = np.full((len(boxes), max_hits), 0xffffffff, np.uint32)
        self.L.jref_collide_aabox(self.h, boxes.ctypes.data, len(boxes), object_layer, max_hits, counts.ctypes.data, ids.ctypes.data)
        return counts, ids

    def constraint_states(self):
        n = self.L.jref_num_constraints(self.h)
        out = np.zeros(n, CONSTRAINT_STATE_DTYPE)
        self.L.jref_get_constraint_states(self.h, out.ctypes.data, n)
        return out

    def remove_constraint(self, index):
        self.L.jref_remove_constraint(self.h, index)

    def set_constraint_enabled(self, index, enabled):
        self.L.jref_set_constraint_enabled(self.h, index, int(enabled))

    def collide_shape(self, kind, params, queries, max_separation_distance=0.0, object_layer=0xffffffff, max_hits=64):
        """NarrowPhaseQuery::CollideShape (all hits) of one convex shape (kind 0 sphere / 1 box / 2 capsule / 3 cylinder) for every query."""
        queries = np.ascontiguousarray(queries, SHAPE_QUERY_DTYPE)
        params = np.ascontiguousarray(list(params) + [0.0] * (4 - len(params)), np.float32)
        counts = np.zeros(len(queries), np.uint32)
        hits = np.zeros((len(queries), max_hits), SHAPE_HIT_DTYPE)
        self.L.jref_collide_shape(self.h, kind, params.ctypes.data, queries.ctypes.data, len(queries), max_separation_distance, object_layer, max_hits,
                                  counts.ctypes.data, hits.ctypes.data)
        return counts, hits

    def collide_volume(self, mode, data, object_layer=0xffffffff, max_hits=64):
        """BroadPhaseQuery::CollideSphere (mode 1, [n][4]) / CollidePoint (mode 2, [n][3])."""
        data = np.ascontiguousarray(data, np.float32)
        counts = np.zeros(len(data), np.uint32)
        ids = np.full((len(data), max_hits), 0xffffffff, np.uint32)
        self.L.jref_collide_volume(self.h, mode, data.ctypes.data, len(data), object_layer, max_hits, counts.ctypes.data, ids.ctypes.data)
        return counts, ids

    def replace_body(self, index, radius=0.5):
        """Destroy the body in slot `index`, create a sphere there (same slot, next sequence number); returns the new id."""
        new_id = self.L.jref_replace_body(self.h, index, radius)
        assert new_id != 0xffffffff, "jref_replace_body: the freed slot was not reused"
        return new_id

    def set_recording(self, on):
        self.L.jref_set_recording(self.h, 1 if on else 0)

    @property
    def num_bodies(self):
        return self.L.jref_num_bodies(self.h)

    @property
    def num_dynamic(self):
        return self.L.jref_num_dynamic(self.h)

    @property
    def num_active(self):
        return self.L.jref_num_active(self.h)

    def num_slots(self):
        return self.num_bodies  # harness scenes never remove bodies: slots are dense

    def state(self, n=None):
        n = self.num_slots() if n is None else n
        s = State(n)
        self.L.jref_get_state(self.h, n, _u32p(s.ids), _fp(s.pos), _fp(s.rot), _fp(s.lin), _fp(s.ang), _fp(s.bounds), _u32p(s.active_index), _fp(s.sleep_timer))
        return s

    def find_pairs(self):
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 2), np.uint32)
            n = self.L.jref_find_pairs(self.h, _u32p(out), cap)
            if n <= cap:
                return out[:n].copy()
            cap = n

    def cache(self):
        """(pairs, manifolds) of the contact cache written by the last step, as ctypes arrays."""
        nm = C.c_uint32(0)
        n = self.L.jref_get_cache(self.h, None, 0, None, 0, C.byref(nm))
        pairs = (_capi.CachedBodyPair * max(n, 1))()
        mans = (_capi.CachedManifold * max(nm.value, 1))()
        self.L.jref_get_cache(self.h, pairs, n, mans, nm.value, C.byref(nm))
        return pairs, n, mans, nm.value

    def contact_events(self):
        n = self.L.jref_get_contact_events(self.h, None, 0)
        ev = (_capi.ContactEvent * max(n, 1))()
        self.L.jref_get_contact_events(self.h, ev, n)
        return [ev[i] for i in range(n)]

    def activation_events(self):
        n = self.L.jref_get_activation_events(self.h, None, 0)
        ev = (_capi.ActivationEvent * max(n, 1))()
        self.L.jref_get_activation_events(self.h, ev, n)
        return [(ev[i].kind, ev[i].body) for i in range(n)]

    def active_bodies(self):
        n = self.num_active
        out = np.zeros(max(n, 1), np.uint32)
        self.L.jref_get_active_bodies(self.h, _u32p(out), n)
        return out[:n]

    def kinetic_energy(self):
        return self.L.jref_kinetic_energy(self.h)

    def export(self, api, device=0):
        """Re-creates the current reference state inside a new b2j world of library `api`; returns a B2JWorld."""
        if self.L.jref_bind_b2j(api.path.encode()) != 0:
            raise RuntimeError("jref_bind_b2j failed: " + self.L.jref_last_error().decode())
        w = self.L.jref_export_to_b2j(self.h, device)
        if not w:
            raise RuntimeError("jref_export_to_b2j failed: " + self.L.jref_last_error().decode())
        return B2JWorld(api, w, self.num_slots())


class B2JWorld:
    """Thin convenience wrapper around a b2j_world handle."""

    def __init__(self, api, handle, num_slots):
        self.api = api
        self.h = handle
        self.n = num_slots
        self._query_shapes = {}

    def close(self):
        if self.h:
            self.api.b2j_world_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self, dt=1.0 / 60.0, collision_steps=1):
        stats = _capi.StepStats()
        r = self.api.b2j_step(self.h, dt, collision_steps, C.byref(stats))
        if r < 0:
            raise RuntimeError("b2j_step failed: " + self.api.last_error())
        return r, stats

    def state(self):
        s = State(self.n)
        st = _capi.BodyState(s.pos.ctypes.data, s.rot.ctypes.data, s.lin.ctypes.data, s.ang.ctypes.data, s.bounds.ctypes.data, s.active_index.ctypes.data, s.sleep_timer.ctypes.data)
        if self.api.b2j_bodies_get_state(self.h, None, self.n, C.byref(st)) != 0:
            raise RuntimeError("b2j_bodies_get_state failed: " + self.api.last_error())
        return s

    def find_pairs(self):
        if self.api.b2j_debug_find_pairs(self.h) != 0:
            raise RuntimeError("b2j_debug_find_pairs failed: " + self.api.last_error())
        return self.pairs()

    def pairs(self):
        n = self.api.b2j_debug_get_pairs(self.h, None, 0)
        out = np.zeros((max(n, 1), 2), np.uint32)
        self.api.b2j_debug_get_pairs(self.h, _u32p(out), n)
        return out[:n].copy()

    def cache(self):
        np_, nm = C.c_uint32(0), C.c_uint32(0)
        self.api.b2j_contact_cache_export(self.h, None, 0, C.byref(np_), None, 0, C.byref(nm))
        pairs = (_capi.CachedBodyPair * max(np_.value, 1))()
        mans = (_capi.CachedManifold * max(nm.value, 1))()
        if self.api.b2j_contact_cache_export(self.h, pairs, np_.value, C.byref(np_), mans, nm.value, C.byref(nm)) != 0:
            raise RuntimeError("b2j_contact_cache_export failed: " + self.api.last_error())
        return pairs, np_.value, mans, nm.value

    def contact_events(self):
        n = self.api.b2j_events_drain(self.h, None, 0)
        ev = (_capi.ContactEvent * max(n, 1))()
        self.api.b2j_events_drain(self.h, ev, n)
        return [ev[i] for i in range(n)]

    def activation_events(self):
        n = self.api.b2j_activation_events_drain(self.h, None, 0)
        ev = (_capi.ActivationEvent * max(n, 1))()
        self.api.b2j_activation_events_drain(self.h, ev, n)
        return [(ev[i].kind, ev[i].body) for i in range(n)]

    def cast_rays(self, rays, object_layer=0xffffffff):
        rays = np.ascontiguousarray(rays, np.float32)
        hits = np.zeros(len(rays), HIT_DTYPE)
        if self.api.b2j_query_cast_rays(self.h, rays.ctypes.data, len(rays), object_layer, hits.ctypes.data) != 0:
            raise RuntimeError("b2j_query_cast_rays failed: " + self.api.last_error())
        return hits

    def collide_aabox(self, boxes, object_layer=0xffffffff, max_hits=64):
        boxes = np.ascontiguousarray(boxes, np.float32)
        counts = np.zeros(len(boxes), np.uint32)
        ids = np.full((len(boxes), max_hits), 0xffffffff, np.uint32)
        if self.api.b2j_query_collide_aabox(self.h, boxes.ctypes.data, len(boxes), object_layer, max_hits, counts.ctypes.data, ids.ctypes.data) != 0:
            raise RuntimeError("b2j_query_collide_aabox failed: " + self.api.last_error())
        return counts, ids

    def constraint_states(self):
        n = self.api.b2j_num_constraints(self.h)
        out = np.zeros(n, CONSTRAINT_STATE_DTYPE)
        if n and self.api.b2j_constraints_get_state(self.h, 0, n, out.ctypes.data) != 0:
            raise RuntimeError("b2j_constraints_get_state failed: " + self.api.last_error())
        return out

    def remove_constraint(self, index):
        idx = np.array([index], np.uint32)
        if self.api.b2j_constraints_remove(self.h, idx.ctypes.data, 1) != 0:
            raise RuntimeError("b2j_constraints_remove failed: " + self.api.last_error())

    def set_constraint_enabled(self, index, enabled):
        idx, en = np.array([index], np.uint32), np.array([1 if enabled else 0], np.uint8)
        if self.api.b2j_constraints_set_enabled(self.h, idx.ctypes.data, 1, en.ctypes.data) != 0:
            raise RuntimeError("b2j_constraints_set_enabled failed: " + self.api.last_error())

    def collide_shape(self, kind, params, queries, max_separation_distance=0.0, object_layer=0xffffffff, max_hits=64):
        """b2j_query_collide_shape with a query shape made from the same parameters as RefWorld.collide_shape."""
        key = (kind, tuple(float(x) for x in params))
        if key not in self._query_shapes:
            p = [C.c_float(float(x)) for x in params]
            if kind == 0: sid = self.api.b2j_shape_sphere(self.h, p[0])
            elif kind == 1: sid = self.api.b2j_shape_box(self.h, (C.c_float * 3)(*[float(x) for x in params[:3]]), p[3])
            elif kind == 2: sid = self.api.b2j_shape_capsule(self.h, p[0], p[1])
            else: sid = self.api.b2j_shape_cylinder(self.h, p[0], p[1], p[2])
            assert sid >= 0, self.api.last_error()
            self._query_shapes[key] = sid
        queries = np.ascontiguousarray(queries, SHAPE_QUERY_DTYPE).copy()
        queries["shape"] = self._query_shapes[key]
        counts = np.zeros(len(queries), np.uint32)
        hits = np.zeros((len(queries), max_hits), SHAPE_HIT_DTYPE)
        if self.api.b2j_query_collide_shape(self.h, queries.ctypes.data, len(queries), max_separation_distance, object_layer, max_hits,
                                            counts.ctypes.data, hits.ctypes.data) != 0:
            raise RuntimeError("b2j_query_collide_shape failed: " + self.api.last_error())
        return counts, hits

    def collide_volume(self, mode, data, object_layer=0xffffffff, max_hits=64):
        data = np.ascontiguousarray(data, np.float32)
        counts = np.zeros(len(data), np.uint32)
        ids = np.full((len(data), max_hits), 0xffffffff, np.uint32)
        f = self.api.b2j_query_collide_sphere if mode == 1 else self.api.b2j_query_collide_point
        if f(self.h, data.ctypes.data, len(data), object_layer, max_hits, counts.ctypes.data, ids.ctypes.data) != 0:
            raise RuntimeError("b2j_query_collide_sphere / point failed: " + self.api.last_error())
        return counts, ids

    def profile(self):
        """{kernel name: {"ms": device time, "launches": count}} accumulated since b2j_world_set_profiling(1)."""
        cap, stride = 128, 64
        names = C.create_string_buffer(cap * stride)
        ms = (C.c_float * cap)()
        launches = (C.c_uint32 * cap)()
        n = self.api.b2j_world_get_profile(self.h, names, stride, ms, launches, cap)
        out = {}
        for i in range(min(n, cap)):
            name = names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()
            out[name] = {"ms": float(ms[i]), "launches": int(launches[i])}
        return out

    def active_bodies(self):
        n = self.api.b2j_num_active_bodies(self.h)
        out = np.zeros(max(n, 1), np.uint32)
        self.api.b2j_get_active_bodies(self.h, _u32p(out), n)
        return out[:n]


def cache_summary(pairs, n_pairs, mans, n_mans):
    """{(body1, body2): [(sub1, sub2, num_points), ...]} from a cache export."""
    out = {}
    for i in range(n_pairs):
        p = pairs[i]
        out[(p.body1, p.body2)] = [(mans[j].sub_shape1, mans[j].sub_shape2, mans[j].num_points) for j in range(p.first_manifold, p.first_manifold + p.num_manifolds)]
    return out


def sorted_rows(a):
    """Rows of an integer table in lexicographic order (set comparison of pair lists)."""
    a = np.asarray(a, dtype=np.int64)
    if a.size == 0:
        return a.reshape(0, a.shape[1] if a.ndim == 2 else 0)
    a = a.reshape(len(a), -1)
    return a[np.lexsort(a.T[::-1])]


def cache_rows(summary):
    """cache_summary as a sorted table of (body1, body2, sub1, sub2, num_points) rows."""
    rows = [(b1, b2, s1, s2, n) for (b1, b2), mans in summary.items() for (s1, s2, n) in mans]
    return sorted_rows(np.array(rows, dtype=np.int64).reshape(-1, 5))


def compare_states(ref, got, rel=1e-4, abs_pos=1e-5):
    """Max violation of |d| <= max(abs, rel * |ref|) over position / rotation / velocities (north star tolerance)."""
    worst = {}
    valid = ref.ids != 0xffffffff
    for name in ("pos", "rot", "lin", "ang"):
        a, b = getattr(ref, name)[valid], getattr(got, name)[valid]
        if name == "rot":  # q and -q are the same rotation
            sign = np.sign(np.sum(a * b, axis=1, keepdims=True))
            sign[sign == 0] = 1
            b = b * sign
        err = np.linalg.norm(a - b, axis=1)
        tol = np.maximum(abs_pos, rel * np.linalg.norm(a, axis=1))
        worst[name] = float(np.max(err / tol)) if len(err) else 0.0
        worst[name + "_abs"] = float(np.max(err)) if len(err) else 0.0
    return worst
