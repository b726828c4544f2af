"""ctypes binding of include/jolt_b200.h (the C ABI of libjolt_b200.so).

Plumbing only: struct layouts + prototypes. The product library is CUDA-only; `load()` raises if it is missing
(there is no CPU fallback). `CApi(path)` can bind any library that exports the same ABI -- tests use that to bind the
host-simulation debug build under tests/hostsim.
"""
import ctypes as C
import os

c_f3 = C.c_float * 3
c_f4 = C.c_float * 4


class Settings(C.Structure):
    _fields_ = [
        ("speculative_contact_distance", C.c_float), ("penetration_slop", C.c_float), ("baumgarte", C.c_float),
        ("max_penetration_distance", C.c_float), ("manifold_tolerance", C.c_float),
        ("body_pair_cache_max_delta_position_sq", C.c_float), ("body_pair_cache_cos_max_delta_rotation_div2", C.c_float),
        ("contact_normal_cos_max_delta_rotation", C.c_float), ("contact_point_preserve_lambda_max_dist_sq", C.c_float),
        ("min_velocity_for_restitution", C.c_float), ("time_before_sleep", C.c_float), ("point_velocity_sleep_threshold", C.c_float),
        ("num_velocity_steps", C.c_uint32), ("num_position_steps", C.c_uint32),
        ("deterministic_simulation", C.c_uint8), ("constraint_warm_start", C.c_uint8), ("use_body_pair_contact_cache", C.c_uint8),
        ("use_manifold_reduction", C.c_uint8), ("use_large_island_splitter", C.c_uint8), ("allow_sleeping", C.c_uint8),
        ("check_active_edges", C.c_uint8), ("_pad", C.c_uint8),
    ]


class WorldDesc(C.Structure):
    _fields_ = [
        ("max_bodies", C.c_uint32), ("max_body_pairs", C.c_uint32), ("max_contact_constraints", C.c_uint32),
        ("num_object_layers", C.c_uint32), ("num_broadphase_layers", C.c_uint32),
        ("object_to_broadphase", C.POINTER(C.c_uint8)), ("object_vs_broadphase", C.POINTER(C.c_uint8)), ("object_vs_object", C.POINTER(C.c_uint8)),
        ("settings", Settings), ("gravity", c_f3), ("device", C.c_int32),
    ]


class BodyDesc(C.Structure):
    _fields_ = [
        ("id", C.c_uint32), ("shape", C.c_int32), ("motion_type", C.c_uint8), ("allowed_dofs", C.c_uint8),
        ("num_velocity_steps_override", C.c_uint8), ("num_position_steps_override", C.c_uint8),
        ("object_layer", C.c_uint16), ("flags", C.c_uint16),
        ("position", c_f3), ("rotation", c_f4), ("linear_velocity", c_f3), ("angular_velocity", c_f3),
        ("force", c_f3), ("torque", c_f3), ("inv_mass", C.c_float), ("inv_inertia_diag", c_f3), ("inertia_rotation", c_f4),
        ("linear_damping", C.c_float), ("angular_damping", C.c_float), ("max_linear_velocity", C.c_float), ("max_angular_velocity", C.c_float),
        ("gravity_factor", C.c_float), ("friction", C.c_float), ("restitution", C.c_float),
        ("bounds_min", c_f3), ("bounds_max", c_f3), ("sleep_spheres", c_f4 * 3), ("sleep_timer", C.c_float),
        ("has_bounds", C.c_uint8), ("active", C.c_uint8), ("_pad", C.c_uint8 * 2),
    ]


class BodyState(C.Structure):
    _fields_ = [
        ("position", C.c_void_p), ("rotation", C.c_void_p), ("linear_velocity", C.c_void_p), ("angular_velocity", C.c_void_p),
        ("bounds", C.c_void_p), ("active_index", C.c_void_p), ("sleep_timer", C.c_void_p),
    ]


class BodyParams(C.Structure):
    """b2j_body_params: optional [n] float arrays (NULL = leave untouched)."""
    _fields_ = [(n, C.c_void_p) for n in ("friction", "restitution", "gravity_factor", "linear_damping", "angular_damping", "max_linear_velocity", "max_angular_velocity")]


class CachedBodyPair(C.Structure):
    _fields_ = [("body1", C.c_uint32), ("body2", C.c_uint32), ("delta_position", c_f3), ("delta_rotation", c_f3),
                ("first_manifold", C.c_uint32), ("num_manifolds", C.c_uint32)]


class CachedManifold(C.Structure):
    _fields_ = [("sub_shape1", C.c_uint32), ("sub_shape2", C.c_uint32), ("normal", c_f3), ("friction_lambda", C.c_float * 2),
                ("angular_friction_lambda", C.c_float), ("num_points", C.c_uint32), ("flags", C.c_uint32),
                ("position1", c_f3 * 4), ("position2", c_f3 * 4), ("non_penetration_lambda", c_f4)]


class StepStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "num_active_bodies", "num_bodies", "num_body_pairs", "num_pairs_from_cache", "num_manifolds", "num_contact_points",
        "num_constraints", "num_islands", "num_large_islands", "num_phases", "velocity_iterations", "position_iterations",
        "num_activated", "num_deactivated", "kernel_launches", "error_bits")] + [("gpu_ms", C.c_float), ("kinetic_energy", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ContactEvent(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("body1", C.c_uint32), ("body2", C.c_uint32), ("sub_shape1", C.c_uint32), ("sub_shape2", C.c_uint32),
                ("num_points", C.c_uint32), ("base_offset", c_f3), ("normal", c_f3), ("penetration_depth", C.c_float),
                ("points1", c_f3 * 4), ("points2", c_f3 * 4)]


class ActivationEvent(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("body", C.c_uint32)]


class DebugManifold(C.Structure):
    _fields_ = [("body1", C.c_uint32), ("body2", C.c_uint32), ("sub_shape1", C.c_uint32), ("sub_shape2", C.c_uint32),
                ("num_points", C.c_uint32), ("from_cache", C.c_uint32), ("normal", c_f3), ("penetration_depth", C.c_float)]


class HullDesc(C.Structure):
    _fields_ = [("num_points", C.c_uint32), ("points", C.POINTER(C.c_float)), ("point_num_faces", C.POINTER(C.c_int32)),
                ("point_faces", C.POINTER(C.c_int32)), ("num_faces", C.c_uint32), ("face_first_vertex", C.POINTER(C.c_uint16)),
                ("face_num_vertices", C.POINTER(C.c_uint16)), ("planes", C.POINTER(C.c_float)), ("num_vertex_idx", C.c_uint32),
                ("vertex_idx", C.POINTER(C.c_uint8)), ("convex_radius", C.c_float), ("center_of_mass", c_f3),
                ("local_bounds_min", c_f3), ("local_bounds_max", c_f3), ("inner_radius", C.c_float)]


class MeshDesc(C.Structure):
    _fields_ = [("tree", C.POINTER(C.c_uint8)), ("tree_size", C.c_uint32), ("local_bounds_min", c_f3), ("local_bounds_max", c_f3)]


class Ray(C.Structure):
    _fields_ = [("origin", c_f3), ("direction", c_f3)]


class RayHit(C.Structure):
    _fields_ = [("body", C.c_uint32), ("sub_shape", C.c_uint32), ("fraction", C.c_float)]


# every symbol include/jolt_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_U32P = C.POINTER(C.c_uint32)
PROTOTYPES = {
    "b2j_settings_default": (None, [C.POINTER(Settings)]),
    "b2j_world_create": (_VP, [C.POINTER(WorldDesc)]),
    "b2j_world_destroy": (None, [_VP]),
    "b2j_last_error": (C.c_char_p, []),
    "b2j_world_set_gravity": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "b2j_world_set_settings": (C.c_int, [_VP, C.POINTER(Settings)]),
    "b2j_world_get_settings": (C.c_int, [_VP, C.POINTER(Settings)]),
    "b2j_world_set_previous_delta_time": (C.c_int, [_VP, C.c_float]),
    "b2j_shape_sphere": (C.c_int32, [_VP, C.c_float]),
    "b2j_shape_box": (C.c_int32, [_VP, C.POINTER(C.c_float), C.c_float]),
    "b2j_shape_capsule": (C.c_int32, [_VP, C.c_float, C.c_float]),
    "b2j_shape_convex_hull": (C.c_int32, [_VP, C.POINTER(HullDesc)]),
    "b2j_shape_mesh": (C.c_int32, [_VP, C.POINTER(MeshDesc)]),
    "b2j_shape_cylinder": (C.c_int32, [_VP, C.c_float, C.c_float, C.c_float]),
    "b2j_shape_static_compound": (C.c_int32, [_VP, _VP]),
    "b2j_shape_scaled": (C.c_int32, [_VP, C.c_int32, C.POINTER(C.c_float)]),
    "b2j_shape_rotated_translated": (C.c_int32, [_VP, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b2j_bodies_add": (C.c_int, [_VP, C.POINTER(BodyDesc), C.c_uint32]),
    "b2j_bodies_remove": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_bodies_activate": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_bodies_activate_or_reset_sleep_timer": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_bodies_reset_sleep_timer": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_bodies_deactivate": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_set_active_list": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_bodies_get_state": (C.c_int, [_VP, _U32P, C.c_uint32, C.POINTER(BodyState)]),
    "b2j_bodies_set_state": (C.c_int, [_VP, _U32P, C.c_uint32, C.POINTER(BodyState)]),
    "b2j_host_buffer_register": (C.c_int, [_VP, C.c_size_t]),
    "b2j_host_buffer_unregister": (C.c_int, [_VP]),
    "b2j_bodies_get_stepped_state": (C.c_uint32, [_VP, C.c_uint32, _U32P, C.POINTER(BodyState)]),
    "b2j_bodies_set_params": (C.c_int, [_VP, _U32P, C.c_uint32, C.c_void_p]),
    "b2j_bodies_set_info": (C.c_int, [_VP, _U32P, C.c_uint32, C.c_void_p]),
    "b2j_bodies_add_force_torque": (C.c_int, [_VP, _U32P, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b2j_num_bodies": (C.c_uint32, [_VP]),
    "b2j_num_active_bodies": (C.c_uint32, [_VP]),
    "b2j_get_active_bodies": (C.c_uint32, [_VP, _U32P, C.c_uint32]),
    "b2j_contact_cache_import": (C.c_int, [_VP, C.POINTER(CachedBodyPair), C.c_uint32, C.POINTER(CachedManifold), C.c_uint32]),
    "b2j_contact_cache_export": (C.c_int, [_VP, C.POINTER(CachedBodyPair), C.c_uint32, _U32P, C.POINTER(CachedManifold), C.c_uint32, _U32P]),
    "b2j_were_bodies_in_contact": (C.c_int, [_VP, C.c_uint32, C.c_uint32]),
    "b2j_step": (C.c_int, [_VP, C.c_float, C.c_int, C.POINTER(StepStats)]),
    "b2j_events_drain": (C.c_uint32, [_VP, C.POINTER(ContactEvent), C.c_uint32]),
    "b2j_activation_events_drain": (C.c_uint32, [_VP, C.POINTER(ActivationEvent), C.c_uint32]),
    "b2j_debug_get_pairs": (C.c_uint32, [_VP, _U32P, C.c_uint32]),
    "b2j_debug_get_manifolds": (C.c_uint32, [_VP, C.POINTER(DebugManifold), C.c_uint32]),
    "b2j_debug_find_pairs": (C.c_int, [_VP]),
    "b2j_debug_check_schedule": (C.c_int, [_VP]),
    "b2j_world_set_profiling": (C.c_int, [_VP, C.c_int]),
    "b2j_world_set_event_recording": (C.c_int, [_VP, C.c_int, C.c_int]),
    "b2j_query_cast_rays": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, _VP]),
    "b2j_query_collide_aabox": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP]),
    "b2j_constraints_add": (C.c_int, [_VP, _VP, C.c_uint32]),
    "b2j_constraints_remove": (C.c_int, [_VP, _VP, C.c_uint32]),
    "b2j_num_constraints": (C.c_uint32, [_VP]),
    "b2j_constraints_set_enabled": (C.c_int, [_VP, _VP, C.c_uint32, _VP]),
    "b2j_constraints_get_state": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "b2j_constraints_set_state": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP]),
    "b2j_query_collide_sphere": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP]),
    "b2j_query_collide_point": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP]),
    "b2j_query_collide_shape": (C.c_int, [_VP, _VP, C.c_uint32, C.c_float, C.c_uint32, C.c_uint32, _VP, _VP]),
    "b2j_batch_query_cast_rays": (C.c_int, [_VP, _VP, _VP, C.c_uint32, C.c_uint32, _VP]),
    "b2j_world_save_state": (_VP, [_VP]),
    "b2j_world_restore_state": (C.c_int, [_VP, _VP]),
    "b2j_snapshot_destroy": (None, [_VP]),
    "b2j_snapshot_size": (C.c_uint64, [_VP]),
    "b2j_batch_save_state": (_VP, [_VP]),
    "b2j_batch_restore_state": (C.c_int, [_VP, _VP]),
    "b2j_world_get_profile": (C.c_uint32, [_VP, C.c_char_p, C.c_uint32, C.POINTER(C.c_float), _U32P, C.c_uint32]),
    "b2j_batch_create": (_VP, [_VP, C.c_uint32, C.c_uint32, C.c_uint32]),
    "b2j_batch_create_on_devices": (_VP, [_VP, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, C.c_uint32, C.c_uint32]),
    "b2j_batch_destroy": (None, [_VP]),
    "b2j_batch_reset_worlds": (C.c_int, [_VP, _U32P, C.c_uint32]),
    "b2j_batch_step": (C.c_int, [_VP, C.c_float, C.c_int, C.POINTER(StepStats)]),
    "b2j_batch_size": (C.c_uint32, [_VP]),
    "b2j_batch_get_state": (C.c_int, [_VP, C.c_uint32, C.c_uint32, C.POINTER(BodyState)]),
    "b2j_batch_add_force_torque": (C.c_int, [_VP, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b2j_batch_set_profiling": (C.c_int, [_VP, C.c_int]),
    "b2j_batch_get_profile": (C.c_uint32, [_VP, C.c_char_p, C.c_uint32, C.POINTER(C.c_float), _U32P, C.c_uint32]),
}


class CApi:
    """Binds a shared library exporting the jolt_b200.h ABI."""

    def __init__(self, path):
        self.path = os.path.abspath(path)
        self.lib = C.CDLL(self.path, mode=C.RTLD_LOCAL)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(self.lib, name)  # raises AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)

    def last_error(self):
        e = self.b2j_last_error()
        return e.decode() if e else ""


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjolt_b200.so")


_api = None


def load():
    """Returns the binding of the product library. Raises if it has not been built (no CPU fallback exists)."""
    global _api
    if _api is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError("libjolt_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` (CUDA only, no CPU fallback)")
        _api = CApi(path)
    return _api
