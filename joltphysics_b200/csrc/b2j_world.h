// b2j_world.h -- device-resident world state (structure of arrays in HBM) shared by all kernels.
//
// Layout rationale (B200): every per-body quantity that kernels gather by body slot is a 16-byte element (float4 or a packed
// 16/32-byte struct) so one random access costs one 32-byte sector; everything the solver streams is structure-of-arrays in
// solve order (see b2j_solver.h). Replaces the reference's pointer-chasing AoS Body (128 B, Body.h:445-471) +
// MotionProperties (192 B, MotionProperties.h:288-330).
#pragma once

#include "b2j_math.h"
#include "../../include/jolt_b200.h"

namespace b2j {

struct alignas(16) F4 { float x, y, z, w; };   // 16 byte aligned: one LDG.128 / STG.128
B2J_HD F4 f4(float x, float y, float z, float w) { F4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
B2J_HD F4 f4(V3 v, float w = 0.0f) { return f4(v.x, v.y, v.z, w); }
B2J_HD V3 to_v3(F4 f) { return v3(f.x, f.y, f.z); }
B2J_HD Q4 to_q4(F4 f) { return q4(f.x, f.y, f.z, f.w); }
B2J_HD F4 f4(Q4 q) { return f4(q.x, q.y, q.z, q.w); }

// Two per body quantities that are (almost) always read together live in ONE array as 32 byte pairs (element 2b and 2b + 1): a random
// access to a body then costs one DRAM sector instead of two half used ones. The views keep the kernels' `w.position[b]` spelling.
template <int K> struct F4PairView
{
	F4 *base;
	B2J_HD F4 &operator[](size_t i) const { return base[2 * i + K]; }
};

// Static per-body info (16 B)
struct alignas(16) BodyInfo
{
	uint32_t id;                 // full BodyID (index | sequence << 23), B2J_INVALID_ID for an empty slot
	int32_t  shape;              // index into shapes
	uint16_t object_layer;
	uint8_t  motion_type;        // B2J_MOTION_*
	uint8_t  bp_layer;
	uint16_t flags;              // B2J_BODY_*
	uint8_t  allowed_dofs;
	uint8_t  steps_override;     // velocity steps override | position steps override << 4 (both < 16 in practice, clamped)
};

// Per-body scalar parameters (32 B)
struct alignas(16) BodyParams
{
	float inv_mass, linear_damping, angular_damping, max_linear_velocity;
	float max_angular_velocity, gravity_factor, friction, restitution;
};

struct ShapeDesc
{
	uint32_t kind;               // B2J_SHAPE_*
	float radius;                // sphere / capsule radius
	float convex_radius;         // box / hull convex radius
	float half_height;           // capsule
	V3 half_extent;              // box
	float inner_radius;
	V3 local_min, local_max;     // Shape::GetLocalBounds
	V3 center_of_mass;
	uint32_t hull_point_offset, hull_num_points;   // hull_points / hull_shrunk (same indexing)
	uint32_t hull_face_offset, hull_num_faces;     // hull_planes / hull_faces
	uint32_t hull_vtx_offset;                      // hull_vtx
	uint32_t mesh_offset, mesh_size;               // mesh_bytes (mesh: the cooked tree; compound: the 64 byte quad tree nodes)
	uint32_t compound_sub_offset, compound_num_subs, compound_sub_bits; // compound_subs; CompoundShape::GetSubShapeIDBits
	// decorated convex shapes (ScaledShape / RotatedTranslatedShape around a convex leaf): the scale is baked into the
	// leaf parameters above by the host (box / sphere / capsule: exactly what the reference's scaled support functions compute; hull:
	// scaled points, shrunk points and planes), the rotation composes with the body's centre of mass transform where the shape is used
	uint32_t flags;                                // SHAPE_*
	uint32_t hull_orig_offset;                     // hull_points: the UNSCALED points (supporting face vertices = PreScaled(transform) * point)
	uint32_t base_leaf;                            // the undecorated leaf shape (ray casts scale the RAY and test the unscaled leaf, ScaledShape.cpp:114-119)
	V3 scale;                                      // accumulated scale (hull supporting face)
	M33 local_rot;                                 // Mat44::sRotation(RotatedTranslatedShape::mRotation)
	V3 outer_min, outer_max;                       // GetLocalBounds() of the outermost shape (Body::GetSleepTestPoints)
};
enum { SHAPE_LOCAL_ROTATION = 1, SHAPE_SCALED_HULL = 2 };

// CompoundShape::SubShape (CompoundShape.h:170-260): a convex shape placed in the compound, position relative to the compound's centre of mass
struct CompoundSub { uint32_t shape; uint32_t pad[3]; F4 position_com; F4 rotation; };

// Body pair cache entry (CachedBodyPair, ContactConstraintManager.h:335-355)
struct alignas(8) CachedPair
{
	uint32_t body1, body2;       // full ids, body1 < body2
	uint32_t slot1, slot2;       // body slots (differ from the id index in batched worlds)
	float dpos[3], drot[3];
	uint32_t first_manifold, num_manifolds;
};

// Cached manifold (CachedManifold + CachedContactPoint, ContactConstraintManager.h:265-327), fixed 4 point slots
struct alignas(16) CachedManifold
{
	uint32_t body1, body2, sub1, sub2;
	float normal[3];             // in body 2 space
	float friction_lambda[2], angular_lambda;
	uint16_t num_points, flags;  // flags: 1 = persisted (reused by the next step)
	float p1[4][3], p2[4][3], lambda[4];
};
enum { MANIFOLD_PERSISTED = 1, MANIFOLD_FROM_CACHE = 0x100 };

// One contact cache generation
struct ContactCache
{
	CachedPair *pairs;
	CachedManifold *manifolds;
	uint32_t *pair_table;        // open addressing hash: index into pairs or 0xffffffff
	uint32_t *num_pairs, *num_manifolds; // device counters
};

// Counters living in device memory (one allocation), zeroed at the start of every step
struct StepCounters
{
	uint32_t num_pairs;              // candidate body pairs (this broadphase round)
	uint32_t num_pairs_total;
	uint32_t num_collide_convex;     // work list sizes
	uint32_t num_collide_mesh;
	uint32_t num_cached;
	uint32_t num_epa;
	uint32_t num_new_manifold_ws;    // world space point blocks of non-cached manifolds
	uint32_t num_constraints;
	uint32_t num_contact_points;
	uint32_t num_pairs_from_cache;
	uint32_t num_woken;              // bodies activated by this step's contacts
	uint32_t num_events;
	uint32_t num_activation_events;
	uint32_t error_bits;
	uint32_t num_large_islands;
	uint32_t num_islands;
	uint32_t num_phases;
	uint32_t sched_remaining;
	uint32_t num_deactivated;
	uint32_t new_active_count;
	uint32_t max_velocity_steps, max_position_steps;
	uint32_t hash_tie;               // number of equal adjacent sort keys seen (documented deviation if != 0)
	uint32_t cache_pairs, cache_manifolds; // sizes of the write cache, published at the end of the step
	uint32_t num_active_joints;      // non contact constraints taking part in this step (b2j_joints.h)
	uint32_t pad[6];
};

// Everything a kernel needs, passed by value (pointers into HBM + scalars)
struct DWorld
{
	// capacities
	uint32_t max_bodies, max_body_pairs, max_constraints, pair_table_size;
	uint32_t num_object_layers, num_bp_layers;
	uint32_t world_stride;       // batched independent worlds: slots per world (world = slot / world_stride), 0 = a single world
	b2j_settings settings;
	V3 gravity;

	// layer tables
	const uint8_t *object_to_bp, *object_vs_bp, *object_vs_object;

	// bodies (SoA by slot)
	BodyInfo *info;
	BodyParams *params;
	F4PairView<0> position;      // centre of mass position         } pose pairs
	F4PairView<1> rotation;      //                                  }
	F4PairView<0> linear_velocity; F4PairView<1> angular_velocity;   // velocity pairs
	F4PairView<0> force; F4PairView<1> torque;
	F4PairView<0> inv_inertia_diag; F4PairView<1> inertia_rotation;  // xyz | quaternion
	F4PairView<0> bounds_min; F4PairView<1> bounds_max;
	F4 *sleep_spheres;           // [slot * 3 + i]
	float *sleep_timer;
	uint32_t *active_index;      // index in the active list or B2J_INACTIVE_INDEX
	uint32_t *active;            // active list: body slots
	uint32_t *num_active;        // device counter

	// shapes
	const ShapeDesc *shapes;
	const F4 *hull_points, *hull_shrunk, *hull_planes;
	const uint32_t *hull_faces;  // first vertex | num vertices << 16
	const uint8_t *hull_vtx;
	const uint8_t *mesh_bytes;
	const CompoundSub *compound_subs;

	// contact caches
	ContactCache read_cache, write_cache;

	StepCounters *counters;
};

B2J_HD uint32_t slot_of(uint32_t id) { return id & 0x7fffffu; }
B2J_HD uint32_t world_of(const DWorld &w, uint32_t slot) { return w.world_stride != 0? slot / w.world_stride : 0u; }

} // namespace b2j
