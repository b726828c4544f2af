"""ctypes wrapper of the facade C interface (joltphysics_b200/host/facade_capi.cpp): scenes built through the C++ facade."""
import ctypes as C
import os

import numpy as np

from joltphysics_b200 import _capi
import refharness as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "bench_assets")


class FacadeLib:
    def __init__(self, path, api):
        self.api = api
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        L.b2jf_last_error.restype = C.c_char_p
        L.b2jf_scene_create.restype = C.c_void_p
        L.b2jf_scene_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.b2jf_scene_destroy.argtypes = [C.c_void_p]
        L.b2jf_scene_world.restype = C.c_void_p
        L.b2jf_scene_world.argtypes = [C.c_void_p]
        L.b2jf_scene_num_dynamic.restype = C.c_uint32
        L.b2jf_scene_num_dynamic.argtypes = [C.c_void_p]
        L.b2jf_scene_num_bodies.restype = C.c_uint32
        L.b2jf_scene_num_bodies.argtypes = [C.c_void_p]
        L.b2jf_scene_flush.argtypes = [C.c_void_p]
        L.b2jf_scene_update.argtypes = [C.c_void_p, C.c_float, C.c_int, C.POINTER(_capi.StepStats)]
        L.b2jf_scene_step_e2e.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.b2jf_scene_mutate.argtypes = [C.c_void_p, C.c_int]
        L.b2jf_scene_last_download_count.restype = C.c_uint32
        L.b2jf_scene_last_download_count.argtypes = [C.c_void_p]
        L.b2jf_scene_cast_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2jf_scene_collide_aabox.restype = C.c_int
        L.b2jf_scene_collide_aabox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2jf_scene_collide_shape.restype = C.c_int
        L.b2jf_scene_collide_shape.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 3 + [C.c_int]
        L.b2jf_scene_collide_sphere.restype = C.c_int
        L.b2jf_scene_collide_sphere.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2jf_scene_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]


class FacadeScene:
    def __init__(self, flib, name, p0=0, p1=0):
        self.flib = flib
        self.h = flib.lib.b2jf_scene_create(name.encode(), p0, p1, ASSETS.encode())
        if not self.h:
            raise RuntimeError("b2jf_scene_create failed: " + flib.lib.b2jf_last_error().decode())
        flib.lib.b2jf_scene_flush(self.h)
        self.num_dynamic = flib.lib.b2jf_scene_num_dynamic(self.h)
        self.num_bodies = flib.lib.b2jf_scene_num_bodies(self.h)
        # a non-owning view of the underlying b2j world for state queries
        self.world = R.B2JWorld(flib.api, flib.lib.b2jf_scene_world(self.h), self.num_bodies)

    def close(self):
        if self.h:
            self.world.h = None  # owned by the scene
            self.flib.lib.b2jf_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def update(self, dt=1.0 / 60.0, collision_steps=1):
        stats = _capi.StepStats()
        r = self.flib.lib.b2jf_scene_update(self.h, dt, collision_steps, C.byref(stats))
        return r, stats

    def mutate(self, phase):
        self.flib.lib.b2jf_scene_mutate(self.h, phase)

    def query(self):
        """api_tour.inl sApiTourQuery: (sorted active ids, number of bodies, active flags of the tour bodies)."""
        import numpy as np
        ids = np.zeros(4096, np.uint32)
        nb, flags = C.c_uint32(), C.c_uint32()
        n = self.flib.lib.b2jf_scene_query(self.h, ids.ctypes.data, len(ids), C.addressof(nb), C.addressof(flags))
        return ids[:n].copy(), nb.value, flags.value

    def cast_rays(self, rays):
        """NarrowPhaseQuery::CastRays through the facade -> structured hits like B2JWorld.cast_rays."""
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        body, sub, frac = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self.flib.lib.b2jf_scene_cast_rays(self.h, rays.ctypes.data, n, body.ctypes.data, sub.ctypes.data, frac.ctypes.data)
        hits = np.zeros(n, R.HIT_DTYPE)
        hits["body"], hits["sub_shape"], hits["fraction"] = body, sub, frac
        return hits

    def collide_aabox(self, box, cap=256):
        box = np.ascontiguousarray(box, np.float32)
        ids = np.zeros(cap, np.uint32)
        n = self.flib.lib.b2jf_scene_collide_aabox(self.h, box.ctypes.data, ids.ctypes.data, cap)
        return np.sort(ids[:min(n, cap)])

    def collide_shape_box(self, half_extent, scale, rotation, position, max_separation=0.0, cap=256):
        """NarrowPhaseQuery::CollideShape of a box through the facade -> (body, sub_shape2, depth) arrays; -1 hits = the forms disagree."""
        he, sc, q, p = (np.ascontiguousarray(x, np.float32) for x in (half_extent, scale, rotation, position))
        body, sub, depth = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.float32)
        n = self.flib.lib.b2jf_scene_collide_shape(self.h, he.ctypes.data, sc.ctypes.data, q.ctypes.data, p.ctypes.data, max_separation,
                                                   body.ctypes.data, sub.ctypes.data, depth.ctypes.data, cap)
        assert n >= 0, "the quaternion, matrix and batched forms of CollideShape disagree"
        n = min(n, cap)
        return body[:n], sub[:n], depth[:n]

    def collide_sphere(self, sphere, cap=256):
        """BroadPhaseQuery::CollideSphere (radius >= 0) / CollidePoint (radius < 0) through the facade -> sorted ids."""
        sphere = np.ascontiguousarray(sphere, np.float32)
        ids = np.zeros(cap, np.uint32)
        n = self.flib.lib.b2jf_scene_collide_sphere(self.h, sphere.ctypes.data, ids.ctypes.data, cap)
        return np.sort(ids[:min(n, cap)])

    def step_e2e(self, dt, forces, out_positions):
        return self.flib.lib.b2jf_scene_step_e2e(self.h, dt, forces.ctypes.data if forces is not None else None, out_positions.ctypes.data if out_positions is not None else None)
