/* jolt_b200.h -- C ABI of the B200-native rigid-body step (libjolt_b200.so).
 *
 * Drop-in boundary for the hot path of jrouwe/JoltPhysics: PhysicsSystem::Update and the BodyInterface /
 * ContactListener / BodyActivationListener surface around it.  The reference has no whole-step plugin API
 * (PhysicsSystem is a concrete class, Jolt/Physics/PhysicsSystem.h:29-396, the broadphase is a #define,
 * Jolt/Physics/PhysicsSystem.cpp:46-47), so the boundary is source-level: a facade with the reference's
 * signatures forwards to the entry points below.  Each entry point cites the reference interface it replaces.
 *
 * Conventions: plain C structs, caller-allocated buffers, integer return codes (0 = OK, <0 = error, see
 * b2j_last_error), no callbacks across the ABI, no exceptions.  Thread-compatible (external synchronisation),
 * like the reference's PhysicsSystem::Update.  All floating point is fp32; ids are the reference's 32-bit BodyID
 * (23-bit index | 8-bit sequence number << 23, Jolt/Physics/Body/BodyID.h:18-21).
 *
 * There is NO CPU fallback: every function that touches a world needs a CUDA device (sm_100a).
 */
#ifndef JOLT_B200_H
#define JOLT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2J_VERSION 1

/* ---- enums (values identical to the reference) ------------------------------------------------------------- */

/* EMotionType, Jolt/Physics/Body/MotionType.h */
enum { B2J_MOTION_STATIC = 0, B2J_MOTION_KINEMATIC = 1, B2J_MOTION_DYNAMIC = 2 };

/* EPhysicsUpdateError bits, Jolt/Physics/EPhysicsUpdateError.h:12-16 */
enum {
	B2J_ERR_NONE = 0,
	B2J_ERR_MANIFOLD_CACHE_FULL = 1,
	B2J_ERR_BODY_PAIR_CACHE_FULL = 2,
	B2J_ERR_CONTACT_CONSTRAINTS_FULL = 4
};

/* Shape kinds on the path (EShapeSubType subset, Jolt/Physics/Collision/Shape/Shape.h) */
enum { B2J_SHAPE_SPHERE = 0, B2J_SHAPE_BOX = 1, B2J_SHAPE_CAPSULE = 2, B2J_SHAPE_CONVEX_HULL = 3, B2J_SHAPE_MESH = 4, B2J_SHAPE_CYLINDER = 5, B2J_SHAPE_COMPOUND = 6 };

/* Constraint kinds on the path (EConstraintSubType subset with its values, Jolt/Physics/Constraints/Constraint.h:33-54) */
enum { B2J_CONSTRAINT_FIXED = 0, B2J_CONSTRAINT_POINT = 1, B2J_CONSTRAINT_HINGE = 2, B2J_CONSTRAINT_DISTANCE = 4 };

/* Body flags */
enum {
	B2J_BODY_SENSOR = 1u << 0,                 /* Body::IsSensor */
	B2J_BODY_ALLOW_SLEEPING = 1u << 1,         /* MotionProperties::mAllowSleeping */
	B2J_BODY_USE_MANIFOLD_REDUCTION = 1u << 2, /* Body::EFlags::UseManifoldReduction */
	B2J_BODY_GYROSCOPIC = 1u << 3,             /* Body::EFlags::ApplyGyroscopicForce */
	B2J_BODY_KIN_VS_NONDYN = 1u << 4,          /* Body::EFlags::CollideKinematicVsNonDynamic */
	B2J_BODY_INVALIDATE_CACHE = 1u << 5        /* Body::EFlags::InvalidateContactCache */
};

#define B2J_INACTIVE_INDEX 0xffffffffu /* Body::cInactiveIndex */
#define B2J_INVALID_ID     0xffffffffu /* BodyID::cInvalidBodyID */

/* Contact event kinds (ContactListener, Jolt/Physics/Collision/ContactListener.h:96-140) */
enum { B2J_EVENT_CONTACT_ADDED = 0, B2J_EVENT_CONTACT_PERSISTED = 1, B2J_EVENT_CONTACT_REMOVED = 2 };
/* Activation event kinds (BodyActivationListener, Jolt/Physics/Body/BodyActivationListener.h:13-26) */
enum { B2J_EVENT_BODY_ACTIVATED = 0, B2J_EVENT_BODY_DEACTIVATED = 1 };

/* ---- settings ------------------------------------------------------------------------------------------------ */

/* The subset of PhysicsSettings (Jolt/Physics/PhysicsSettings.h:30-129) that the path reads; same defaults. */
typedef struct b2j_settings {
	float    speculative_contact_distance;     /* mSpeculativeContactDistance        0.02 */
	float    penetration_slop;                 /* mPenetrationSlop                   0.02 */
	float    baumgarte;                        /* mBaumgarte                         0.2  */
	float    max_penetration_distance;         /* mMaxPenetrationDistance            0.2  */
	float    manifold_tolerance;               /* mManifoldTolerance                 1e-3 */
	float    body_pair_cache_max_delta_position_sq;       /* 1e-6 (Square(0.001)) */
	float    body_pair_cache_cos_max_delta_rotation_div2; /* cos(2deg/2) = 0.99984769515639123915701155881391 */
	float    contact_normal_cos_max_delta_rotation;       /* cos(5deg)   = 0.99619469809174553229501040247389 */
	float    contact_point_preserve_lambda_max_dist_sq;   /* 1e-4 (Square(0.01)) */
	float    min_velocity_for_restitution;     /* mMinVelocityForRestitution         1.0  */
	float    time_before_sleep;                /* mTimeBeforeSleep                   0.5  */
	float    point_velocity_sleep_threshold;   /* mPointVelocitySleepThreshold       0.03 */
	uint32_t num_velocity_steps;               /* mNumVelocitySteps                  10   */
	uint32_t num_position_steps;               /* mNumPositionSteps                  2    */
	uint8_t  deterministic_simulation;         /* mDeterministicSimulation (always honoured; must be 1) */
	uint8_t  constraint_warm_start;            /* mConstraintWarmStart               1 */
	uint8_t  use_body_pair_contact_cache;      /* mUseBodyPairContactCache           1 */
	uint8_t  use_manifold_reduction;           /* mUseManifoldReduction              1 */
	uint8_t  use_large_island_splitter;        /* mUseLargeIslandSplitter            1 */
	uint8_t  allow_sleeping;                   /* mAllowSleeping                     1 */
	uint8_t  check_active_edges;               /* mCheckActiveEdges                  1 */
	uint8_t  _pad;
} b2j_settings;

/* Fills *s with the reference's defaults (PhysicsSettings.h). */
void b2j_settings_default(b2j_settings *s);

/* ---- world --------------------------------------------------------------------------------------------------- */

typedef struct b2j_world b2j_world; /* opaque */

/* Replaces PhysicsSystem::Init(maxBodies, numBodyMutexes, maxBodyPairs, maxContactConstraints, bpLayerInterface,
 * objectVsBpFilter, objectPairFilter) -- PhysicsSystem.h:59.  The three virtual filter objects are sampled into
 * tables by the caller (as ObjectVsBroadPhaseLayerFilterTable does, .../BroadPhase/ObjectVsBroadPhaseLayerFilterTable.h:38-50). */
typedef struct b2j_world_desc {
	uint32_t        max_bodies;
	uint32_t        max_body_pairs;
	uint32_t        max_contact_constraints;
	uint32_t        num_object_layers;       /* <= 64 */
	uint32_t        num_broadphase_layers;   /* <= 8  */
	const uint8_t  *object_to_broadphase;    /* [num_object_layers]                          BroadPhaseLayerInterface::GetBroadPhaseLayer */
	const uint8_t  *object_vs_broadphase;    /* [num_object_layers][num_broadphase_layers]   ObjectVsBroadPhaseLayerFilter::ShouldCollide */
	const uint8_t  *object_vs_object;        /* [num_object_layers][num_object_layers]       ObjectLayerPairFilter::ShouldCollide */
	b2j_settings    settings;
	float           gravity[3];              /* PhysicsSystem::SetGravity, default (0,-9.81,0) */
	int32_t         device;                  /* CUDA device ordinal */
} b2j_world_desc;

b2j_world  *b2j_world_create(const b2j_world_desc *desc);           /* NULL on failure, see b2j_last_error */
void        b2j_world_destroy(b2j_world *w);
const char *b2j_last_error(void);
int         b2j_world_set_gravity(b2j_world *w, const float g[3]);  /* PhysicsSystem::SetGravity  PhysicsSystem.h:194 */
int         b2j_world_set_settings(b2j_world *w, const b2j_settings *s); /* SetPhysicsSettings  PhysicsSystem.h:111 */
int         b2j_world_get_settings(const b2j_world *w, b2j_settings *s);
/* mPreviousStepDeltaTime (PhysicsSystem.h:395) drives the warm start ratio; part of a snapshot. */
int         b2j_world_set_previous_delta_time(b2j_world *w, float dt);

/* ---- shapes (immutable once uploaded; Shape is RefConst and immutable in the reference, Body.h:451) ------------ */

/* Cooked convex hull exactly as ConvexHullShape stores it (Jolt/Physics/Collision/Shape/ConvexHullShape.h:167-195);
 * cooking (ConvexHullBuilder) is host-side and out of scope -- pass the reference's output. */
typedef struct b2j_hull_desc {
	uint32_t        num_points;           /* <= 256 */
	const float    *points;               /* [num_points][3]  relative to the centre of mass */
	const int32_t  *point_num_faces;      /* [num_points]     Point::mNumFaces */
	const int32_t  *point_faces;          /* [num_points][3]  Point::mFaces */
	uint32_t        num_faces;
	const uint16_t *face_first_vertex;    /* [num_faces]      Face::mFirstVertex */
	const uint16_t *face_num_vertices;    /* [num_faces]      Face::mNumVertices */
	const float    *planes;               /* [num_faces][4]   (nx,ny,nz,c) */
	uint32_t        num_vertex_idx;
	const uint8_t  *vertex_idx;           /* [num_vertex_idx] */
	float           convex_radius;
	float           center_of_mass[3];
	float           local_bounds_min[3], local_bounds_max[3];
	float           inner_radius;
} b2j_hull_desc;

/* MeshShape's cooked byte buffer verbatim (NodeCodecQuadTreeHalfFloat + TriangleCodecIndexed8BitPackSOA4Flags,
 * Jolt/Physics/Collision/Shape/MeshShape.cpp:491-553); cooking is host-side and out of scope. */
typedef struct b2j_mesh_desc {
	const uint8_t  *tree;                 /* MeshShape::mTree */
	uint32_t        tree_size;
	float           local_bounds_min[3], local_bounds_max[3];
} b2j_mesh_desc;

/* All return a shape id >= 0, or < 0 on error. */
int32_t b2j_shape_sphere(b2j_world *w, float radius);                                        /* SphereShape */
int32_t b2j_shape_box(b2j_world *w, const float half_extent[3], float convex_radius);        /* BoxShape */
int32_t b2j_shape_capsule(b2j_world *w, float half_height_of_cylinder, float radius);        /* CapsuleShape */
int32_t b2j_shape_cylinder(b2j_world *w, float half_height, float radius, float convex_radius); /* CylinderShape (axis = y; CylinderShape.cpp) */
int32_t b2j_shape_convex_hull(b2j_world *w, const b2j_hull_desc *hull);                      /* ConvexHullShape */
int32_t b2j_shape_mesh(b2j_world *w, const b2j_mesh_desc *mesh);                             /* MeshShape (static bodies) */
/* StaticCompoundShape (Jolt/Physics/Collision/Shape/StaticCompoundShape.h, CompoundShape.h:170-260) of convex sub shapes (plain or
 * decorated). The quad tree over the sub shapes is passed as the reference built it (StaticCompoundShape::mNodes, 64 bytes per node:
 * building it is host side cooking like the convex hull and mesh builders); it decides the order in which sub shapes are tested and
 * with it the order of the contact manifolds. */
typedef struct b2j_compound_sub {
	int32_t shape;                        /* a convex shape id (sphere / box / capsule / cylinder / hull, optionally scaled / rotated) */
	float   position_com[3];              /* SubShape::GetPositionCOM: relative to the compound's centre of mass */
	float   rotation[4];                  /* SubShape::GetRotation (x,y,z,w) */
} b2j_compound_sub;
typedef struct b2j_compound_desc {
	uint32_t                num_subs;
	const b2j_compound_sub *subs;
	uint32_t                num_nodes;
	const uint8_t          *nodes;        /* [num_nodes][64] */
	float                   center_of_mass[3];
	float                   local_bounds_min[3], local_bounds_max[3];
	float                   inner_radius;
} b2j_compound_desc;
int32_t b2j_shape_static_compound(b2j_world *w, const b2j_compound_desc *compound);

/* Decorated shapes. `inner` = a sphere / box / capsule / convex hull / mesh or one of these two around one.
 * ScaledShape (Jolt/Physics/Collision/Shape/ScaledShape.cpp:190-204: the scale is handed down to the leaf's support function,
 * supporting face and bounds): positive scales; uniform for spheres and capsules (SphereShape::IsValidScale) and for an inner
 * RotatedTranslatedShape with a rotation (no RotateScale). */
int32_t b2j_shape_scaled(b2j_world *w, int32_t inner, const float scale[3]);
/* RotatedTranslatedShape (RotatedTranslatedShape.cpp:33-60,183-192): the inner shape rotated by `rotation` (x,y,z,w) about its centre of
 * mass; the translation only moves the centre of mass, which is where the body's position is anyway: center_of_mass = position +
 * rotation * inner centre of mass, as the reference computes it (returned by state getters that report the body origin). */
int32_t b2j_shape_rotated_translated(b2j_world *w, int32_t inner, const float rotation[4], const float center_of_mass[3]);

/* ---- bodies (BodyInterface, Jolt/Physics/Body/BodyInterface.h:39-313) ------------------------------------------ */

/* Everything one step depends on for a body: Body (Body.h:445-471) + MotionProperties
 * (MotionProperties.h:288-330).  position is the CENTRE OF MASS position (Body::mPosition). */
typedef struct b2j_body_desc {
	uint32_t id;                    /* BodyID (index | sequence << 23); the index selects the slot */
	int32_t  shape;                 /* shape id */
	uint8_t  motion_type;           /* B2J_MOTION_* */
	uint8_t  allowed_dofs;          /* EAllowedDOFs bit mask (0x3f = all) */
	uint8_t  num_velocity_steps_override;
	uint8_t  num_position_steps_override;
	uint16_t object_layer;
	uint16_t flags;                 /* B2J_BODY_* */
	float    position[3];
	float    rotation[4];           /* x y z w */
	float    linear_velocity[3];
	float    angular_velocity[3];
	float    force[3];              /* accumulated force  (MotionProperties::mForce) */
	float    torque[3];
	float    inv_mass;
	float    inv_inertia_diag[3];   /* mInvInertiaDiagonal */
	float    inertia_rotation[4];   /* mInertiaRotation */
	float    linear_damping, angular_damping;
	float    max_linear_velocity, max_angular_velocity;
	float    gravity_factor;
	float    friction, restitution;
	float    bounds_min[3], bounds_max[3];   /* cached world AABB (Body::mBounds); ignored unless has_bounds */
	float    sleep_spheres[3][4];            /* mSleepTestSpheres (centre xyz, radius) */
	float    sleep_timer;                    /* mSleepTestTimer */
	uint8_t  has_bounds;                     /* 1: take bounds_* / sleep_* as given (snapshot); 0: compute */
	uint8_t  active;                         /* add to the active list (BodyInterface::AddBody EActivation) */
	uint8_t  _pad[2];
} b2j_body_desc;

/* BodyInterface::AddBody / AddBodiesPrepare+Finalize (BodyInterface.h:89,124-133). Bodies with active=1 are appended
 * to the active list in array order. */
int b2j_bodies_add(b2j_world *w, const b2j_body_desc *bodies, uint32_t n);
/* BodyInterface::RemoveBody(s) (:99,:137) */
int b2j_bodies_remove(b2j_world *w, const uint32_t *ids, uint32_t n);
/* BodyInterface::ActivateBody / DeactivateBody (:142-145) */
int b2j_bodies_activate(b2j_world *w, const uint32_t *ids, uint32_t n);
int b2j_bodies_deactivate(b2j_world *w, const uint32_t *ids, uint32_t n);
/* BodyInterface::ActivateBodyInternal (BodyInterface.cpp:20-28), what every BodyInterface call with EActivation::Activate does: sleeping
 * bodies are activated, bodies that are active already get Body::ResetSleepTimer. And BodyInterface::ResetSleepTimer (:148) alone. */
int b2j_bodies_activate_or_reset_sleep_timer(b2j_world *w, const uint32_t *ids, uint32_t n);
int b2j_bodies_reset_sleep_timer(b2j_world *w, const uint32_t *ids, uint32_t n);
/* Snapshot helper: sets the active list to exactly ids[0..n) in this order (BodyManager::mActiveBodies). */
int b2j_set_active_list(b2j_world *w, const uint32_t *ids, uint32_t n);

/* SoA state block for get/set; any pointer may be NULL (skipped). All arrays have n entries. */
typedef struct b2j_body_state {
	float    *position;         /* [n][3] centre of mass position */
	float    *rotation;         /* [n][4] */
	float    *linear_velocity;  /* [n][3] */
	float    *angular_velocity; /* [n][3] */
	float    *bounds;           /* [n][6] min xyz, max xyz */
	uint32_t *active_index;     /* [n]    index in the active list or B2J_INACTIVE_INDEX */
	float    *sleep_timer;      /* [n] */
} b2j_body_state;

/* BodyInterface::GetPositionAndRotation / GetLinearAndAngularVelocity (:187-216); ids==NULL means slots 0..n-1. */
int b2j_bodies_get_state(b2j_world *w, const uint32_t *ids, uint32_t n, const b2j_body_state *out);
/* State of the bodies the LAST b2j_step simulated (every body that was active at some point of it: the active list before the
 * sleepers left it, bodies woken by contacts included), i.e. exactly the bodies whose state changed: the incremental download of
 * SURVEY 8f-1 (a world of mostly sleeping bodies mirrors only what moved). Copies up to cap ids + state rows (same order) and
 * returns the number of simulated bodies (may exceed cap). */
uint32_t b2j_bodies_get_stepped_state(b2j_world *w, uint32_t cap, uint32_t *ids, const b2j_body_state *out);
/* Page locks a caller owned host buffer (cudaHostRegister) so that b2j_bodies_get_state / b2j_batch_get_state / b2j_*_add_force_torque
 * copy between it and the device directly instead of through the library's pinned staging buffer. Buffers from cudaHostAlloc /
 * torch.Tensor.pin_memory need no registration. Unregister before freeing the buffer. Returns 0, or -1 (the buffer then simply
 * stays pageable: still correct, one host memcpy slower). */
int b2j_host_buffer_register(void *ptr, size_t bytes);
int b2j_host_buffer_unregister(void *ptr);
/* BodyInterface::SetPositionAndRotation / SetLinearAndAngularVelocity; NULL members are left untouched. */
int b2j_bodies_set_state(b2j_world *w, const uint32_t *ids, uint32_t n, const b2j_body_state *in);
/* BodyInterface::AddForce / AddTorque (:220-226): accumulate into mForce / mTorque (either may be NULL). One entry per body and
 * call: the entries are added in parallel (sum several forces of one body before the call, as the facade does). */
int b2j_bodies_add_force_torque(b2j_world *w, const uint32_t *ids, uint32_t n, const float *force, const float *torque);

/* BodyInterface::SetFriction / SetRestitution / SetGravityFactor / SetMaxLinearVelocity / SetMaxAngularVelocity and
 * MotionProperties::SetLinearDamping / SetAngularDamping (BodyInterface.h:241-281): per body scalars, [n] each, NULL members are
 * left untouched. */
typedef struct b2j_body_params {
	const float *friction, *restitution, *gravity_factor;
	const float *linear_damping, *angular_damping;
	const float *max_linear_velocity, *max_angular_velocity;
} b2j_body_params;
int b2j_bodies_set_params(b2j_world *w, const uint32_t *ids, uint32_t n, const b2j_body_params *in);

/* BodyInterface::SetMotionType (BodyInterface.h:241, Body::SetMotionType Body.cpp), SetObjectLayer (:181), SetShape (:169) and
 * InvalidateContactCache (:300) for bodies that are in the world. [n] arrays, NULL members are left untouched.
 *  motion_type: a body that becomes static leaves the active list and stops, static / kinematic bodies lose their accumulated force and
 *    torque; inv_mass (required with motion_type) = the inverse mass the body has as a DYNAMIC body (the device keeps 0 for the others).
 *  object_layer: the body moves to the broadphase tree of the new layer.
 *  shape: the centre of mass position follows the new shape's centre of mass, the bounds are recomputed, the body's cached contacts
 *    are not reused by the next step (BodyManager::InvalidateContactCacheForBody); mass properties change only if the three mass
 *    arrays are given (SetShape's inUpdateMassProperties).
 *  invalidate_contact_cache != 0: BodyInterface::InvalidateContactCache for all n bodies. */
typedef struct b2j_body_info_update {
	const uint8_t  *motion_type;
	const float    *inv_mass;
	const uint16_t *object_layer;
	const int32_t  *shape;
	const float    *inv_inertia_diag;   /* [n][3] with shape */
	const float    *inertia_rotation;   /* [n][4] with shape */
	uint32_t        invalidate_contact_cache;
	const uint16_t *flags_set;          /* [n] B2J_BODY_* bits to set: Body::SetIsSensor / SetUseManifoldReduction / SetAllowSleeping / */
	const uint16_t *flags_clear;        /* [n] ... and to clear         SetApplyGyroscopicForce / SetCollideKinematicVsNonDynamic       */
} b2j_body_info_update;
int b2j_bodies_set_info(b2j_world *w, const uint32_t *ids, uint32_t n, const b2j_body_info_update *in);

uint32_t b2j_num_bodies(const b2j_world *w);          /* PhysicsSystem::GetNumBodies        PhysicsSystem.h:219 */
uint32_t b2j_num_active_bodies(const b2j_world *w);   /* PhysicsSystem::GetNumActiveBodies  :222 */
/* PhysicsSystem::GetActiveBodies (:240): copies up to cap ids in active-list order, returns the count. */
uint32_t b2j_get_active_bodies(b2j_world *w, uint32_t *ids, uint32_t cap);

/* ---- non contact constraints between two bodies (PointConstraint, DistanceConstraint without limit springs,
 *      HingeConstraint with angle limits and friction, motor off, FixedConstraint).
 *      Replaces PhysicsSystem::AddConstraint(s) / RemoveConstraint(s) (PhysicsSystem.h:128-137 -> ConstraintManager::Add / Remove,
 *      Jolt/Physics/Constraints/ConstraintManager.cpp:17-62). A constraint is addressed by its position in the world's list, which is
 *      Constraint::mConstraintIndex: adding appends, removing moves the last constraint into the freed position. Active constraints take
 *      part in the step as the reference's do (islands, large island splits, warm start, velocity and position iterations) and count
 *      against max_contact_constraints. Remove a body's constraints before the body, as the reference asks. ------------------------ */

typedef struct b2j_constraint_desc {
	uint32_t type;                       /* B2J_CONSTRAINT_* */
	uint32_t body1, body2;               /* TwoBodyConstraint::mBody1 / mBody2 (ids; a static body plays Body::sFixedToWorld)             */
	float    point1[3], point2[3];       /* mLocalSpacePosition1 / 2: relative to the centre of mass of body 1 / 2                       */
	float    min_distance, max_distance; /* DistanceConstraint::mMinDistance / mMaxDistance (already resolved, >= 0)                      */
	uint32_t priority;                   /* Constraint::mConstraintPriority                                                              */
	uint8_t  num_velocity_steps_override, num_position_steps_override; /* Constraint::mNumVelocityStepsOverride / mNumPositionStepsOverride */
	uint8_t  enabled;                    /* Constraint::mEnabled                                                                          */
	uint8_t  reserved;
	/* HingeConstraint (HingeConstraint.h:118-160): mLocalSpaceHingeAxis1 / 2, mInvInitialOrientation, mLimitsMin / Max ([-pi, 0] / [0, pi]),
	 * mMaxFrictionTorque; the motor is off and the limits have no spring. FixedConstraint (FixedConstraint.h): point1 / point2 and
	 * inv_initial_orientation (mInvInitialOrientation) */
	float    hinge_axis1[3], hinge_axis2[3];
	float    inv_initial_orientation[4];
	float    limits_min, limits_max, max_friction_torque;
} b2j_constraint_desc;

/* What Constraint::SaveState writes plus what the distance constraint keeps between steps (DistanceConstraint.cpp:237-243): the
 * accumulated impulses the next step warm starts from and mWorldSpaceNormal. */
typedef struct b2j_constraint_state
{
	float total_lambda[3];           /* point / hinge: mPointConstraintPart (xyz); distance: mAxisConstraint (x)  */
	float world_space_normal[3];     /* distance: mWorldSpaceNormal                                                */
	float total_lambda_rotation[3];  /* hinge: mRotationConstraintPart (xy); fixed: mRotationConstraintPart (xyz)  */
	float total_lambda_limits;       /* hinge: mRotationLimitsConstraintPart                                       */
	float total_lambda_motor;        /* hinge: mMotorConstraintPart (the friction while the motor is off)          */
} b2j_constraint_state;

int      b2j_constraints_add(b2j_world *w, const b2j_constraint_desc *constraints, uint32_t n);
int      b2j_constraints_remove(b2j_world *w, const uint32_t *indices, uint32_t n);       /* each index as of the removals before it */
uint32_t b2j_num_constraints(const b2j_world *w);                                          /* ConstraintManager::GetNumConstraints */
int      b2j_constraints_set_enabled(b2j_world *w, const uint32_t *indices, uint32_t n, const uint8_t *enabled); /* Constraint::SetEnabled */
int      b2j_constraints_get_state(b2j_world *w, uint32_t first, uint32_t n, b2j_constraint_state *out);
int      b2j_constraints_set_state(b2j_world *w, uint32_t first, uint32_t n, const b2j_constraint_state *in);

/* ---- contact cache snapshot (ContactConstraintManager::ManifoldCache, SaveState stream sections
 *      Jolt/Physics/Constraints/ContactConstraintManager.cpp:467-548) -------------------------------------------- */

typedef struct b2j_cached_body_pair {
	uint32_t body1, body2;          /* body1 < body2 */
	float    delta_position[3];     /* CachedBodyPair::mDeltaPosition (in body-1 space) */
	float    delta_rotation[3];     /* CachedBodyPair::mDeltaRotation (xyz, w >= 0 reconstructed) */
	uint32_t first_manifold;        /* index into the manifold array */
	uint32_t num_manifolds;
} b2j_cached_body_pair;

typedef struct b2j_cached_manifold {
	uint32_t sub_shape1, sub_shape2;   /* SubShapeIDPair (body ids come from the owning pair) */
	float    normal[3];                /* CachedManifold::mContactNormal (body-2 space) */
	float    friction_lambda[2];
	float    angular_friction_lambda;
	uint32_t num_points;               /* 1..4 */
	uint32_t flags;                    /* CachedManifold::EFlags */
	float    position1[4][3];          /* CachedContactPoint::mPosition1 (body-1 space) */
	float    position2[4][3];
	float    non_penetration_lambda[4];
} b2j_cached_manifold;

/* Replaces RestoreState(Contacts): installs the READ cache (what the next step warm-starts from). */
int b2j_contact_cache_import(b2j_world *w, const b2j_cached_body_pair *pairs, uint32_t num_pairs,
                             const b2j_cached_manifold *manifolds, uint32_t num_manifolds);
/* Replaces SaveState(Contacts): pairs sorted by (body1, body2); returns counts through the out params. */
int b2j_contact_cache_export(b2j_world *w, b2j_cached_body_pair *pairs, uint32_t pairs_cap, uint32_t *num_pairs,
                             b2j_cached_manifold *manifolds, uint32_t manifolds_cap, uint32_t *num_manifolds);
/* PhysicsSystem::WereBodiesInContact (PhysicsSystem.h:251) */
int b2j_were_bodies_in_contact(b2j_world *w, uint32_t id1, uint32_t id2);

/* ---- queries on the device broadphase / shapes, batched: thousands per call, one thread per query ------------------ */

/* RayCast (Jolt/Physics/Collision/RayCast.h): origin + fraction * direction, fraction in [0, 1] */
typedef struct b2j_ray { float origin[3]; float direction[3]; } b2j_ray;
/* RayCastResult (Jolt/Physics/Collision/CastResult.h): body = B2J_INVALID_ID and fraction = 1 + FLT_EPSILON when nothing was hit */
typedef struct b2j_ray_hit { uint32_t body; uint32_t sub_shape; float fraction; } b2j_ray_hit;

/* NarrowPhaseQuery::CastRay, closest hit (NarrowPhaseQuery.h:31), for n rays at once. object_layer: the layer the rays collide as
 * (the world's ObjectVsBroadPhaseLayerFilter / ObjectLayerPairFilter tables play DefaultBroadPhaseLayerFilter / DefaultObjectLayerFilter,
 * Jolt/Physics/Collision/BroadPhase/BroadPhaseLayer.h:112, ObjectLayer.h:77); 0xffffffff = collide with every layer. */
int b2j_query_cast_rays(b2j_world *w, const b2j_ray *rays, uint32_t n, uint32_t object_layer, b2j_ray_hit *hits);
/* BroadPhaseQuery::CollideAABox (BroadPhaseQuery.h:38) for n boxes ([n][6] min xyz, max xyz): counts[i] = bodies whose world space
 * bounds overlap box i, ids[i * max_hits ...] = the first max_hits of them. Exact body bounds (the reference reports the possibly
 * widened bounds of its tree: a superset). */
int b2j_query_collide_aabox(b2j_world *w, const float *boxes, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids);
/* BroadPhaseQuery::CollideSphere (BroadPhaseQuery.h:41; spheres: [n][4] centre xyz, radius) and BroadPhaseQuery::CollidePoint (:44;
 * points: [n][3]): as b2j_query_collide_aabox with the sphere / point against the exact world space bounds of the bodies
 * (AABox4VsSphere / AABox4VsPoint, Jolt/Geometry/AABox4.h). */
int b2j_query_collide_sphere(b2j_world *w, const float *spheres, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids);
int b2j_query_collide_point(b2j_world *w, const float *points, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids);

/* One NarrowPhaseQuery::CollideShape call (NarrowPhaseQuery.h:53): a convex shape of the world's shape table (b2j_shape_sphere / box /
 * capsule / cylinder / convex_hull, optionally scaled / rotated + translated; inShapeScale is expressed with b2j_shape_scaled) at a
 * centre of mass transform (inCenterOfMassTransform = rotation, position), results relative to base_offset (inBaseOffset). */
typedef struct b2j_shape_query { int32_t shape; float position[3]; float rotation[4]; float base_offset[3]; } b2j_shape_query;
/* CollideShapeResult (Jolt/Physics/Collision/CollideShape.h:18-68) without the faces (ECollectFacesMode::NoFaces, the default of
 * CollideShapeSettings): contact points on shape 1 (the query shape) and shape 2 (the body) relative to base_offset, the penetration
 * axis (direction to move shape 2 out of collision, not normalised), the depth (negative: separated by less than
 * max_separation_distance), the sub shape ids of both sides and the body that was hit. */
typedef struct b2j_collide_shape_hit
{
	uint32_t body, sub_shape1, sub_shape2;
	float    penetration_depth;
	float    point1[3], point2[3], axis[3];
} b2j_collide_shape_hit;
/* NarrowPhaseQuery::CollideShape with an AllHitCollisionCollector for n queries at once: the bodies whose bounds overlap the query
 * shape's bounds (expanded by max_separation_distance) are collided with it exactly as TransformedShape::CollideShape does
 * (CollideShapeSettings defaults: mMaxSeparationDistance as given, IgnoreBackFaces, CollideOnlyWithActive edges with no movement
 * direction, NoFaces, tolerances 1e-4); bodies may be convex shapes, StaticCompoundShapes or meshes. counts[i] = hits of query i (can
 * exceed max_hits), hits[i * max_hits ...] = the first max_hits of them in the order they were found. */
int b2j_query_collide_shape(b2j_world *w, const b2j_shape_query *queries, uint32_t n, float max_separation_distance, uint32_t object_layer,
                            uint32_t max_hits, uint32_t *counts, b2j_collide_shape_hit *hits);

/* ---- state snapshots on the device (PhysicsSystem::SaveState / RestoreState, PhysicsSystem.cpp:2899-2964; what a snapshot holds:
 *      EStateRecorderState::Global | Bodies | Contacts, i.e. mPreviousStepDeltaTime + gravity, every body's state and the contact
 *      cache ContactConstraintManager::SaveState writes, .cpp:467-548 -- plus the order of the active list, which the reference
 *      leaves to the caller). The snapshot stays in HBM: rollback / environment reset without a host round trip of the state. -------- */

typedef struct b2j_snapshot b2j_snapshot; /* opaque; owned by the caller, tied to the world (or batch) it was taken from */

b2j_snapshot *b2j_world_save_state(b2j_world *w);                        /* NULL on failure */
/* Restores bodies (also the set of bodies: bodies added / removed since the save are undone), active list, contact cache and the
 * previous step's delta time. Shapes uploaded since the save stay. */
int           b2j_world_restore_state(b2j_world *w, const b2j_snapshot *s);
void          b2j_snapshot_destroy(b2j_snapshot *s);
uint64_t      b2j_snapshot_size(const b2j_snapshot *s);                  /* bytes of device memory the snapshot holds */

/* ---- step ---------------------------------------------------------------------------------------------------- */

/* Per-step counters the roofline is computed from; all are totals over the collision steps of the call. */
typedef struct b2j_step_stats {
	uint32_t num_active_bodies;       /* N_a at the end of the step */
	uint32_t num_bodies;              /* N_all */
	uint32_t num_body_pairs;          /* P   candidate pairs from the broadphase */
	uint32_t num_pairs_from_cache;    /* P_hit */
	uint32_t num_manifolds;           /* M   manifolds written to the cache */
	uint32_t num_contact_points;      /* sum of c */
	uint32_t num_constraints;         /* contact constraints solved */
	uint32_t num_islands;
	uint32_t num_large_islands;
	uint32_t num_phases;              /* dependent solver phases per iteration (launch/barrier count) */
	uint32_t velocity_iterations;     /* V executed (max over islands) */
	uint32_t position_iterations;     /* Pp */
	uint32_t num_activated, num_deactivated;
	uint32_t kernel_launches;         /* CUDA kernels launched by this call */
	uint32_t error_bits;
	float    gpu_ms;                  /* device time of the step (CUDA events on the world's stream) */
	float    kinetic_energy;          /* filled only when requested by b2j_step flags */
} b2j_step_stats;

/* Replaces PhysicsSystem::Update(deltaTime, collisionSteps, TempAllocator*, JobSystem*) -- PhysicsSystem.h:162.
 * Returns the EPhysicsUpdateError bit field (>= 0) or < 0 on a CUDA failure. stats may be NULL. */
int b2j_step(b2j_world *w, float delta_time, int collision_steps, b2j_step_stats *stats);

/* ---- events (replayed into ContactListener / BodyActivationListener by the facade after the step) ------------- */

typedef struct b2j_contact_event {
	uint32_t kind;                 /* B2J_EVENT_CONTACT_* */
	uint32_t body1, body2;         /* body1 < body2 (ContactListener.h:102,116) */
	uint32_t sub_shape1, sub_shape2;
	uint32_t num_points;           /* 0 for removed */
	float    base_offset[3];       /* ContactManifold::mBaseOffset */
	float    normal[3];            /* mWorldSpaceNormal */
	float    penetration_depth;
	float    points1[4][3];        /* mRelativeContactPointsOn1 */
	float    points2[4][3];
} b2j_contact_event;

typedef struct b2j_activation_event {
	uint32_t kind;                 /* B2J_EVENT_BODY_* */
	uint32_t body;
} b2j_activation_event;

/* Copies up to cap events of the last b2j_step (sorted by kind, body1, body2, sub shapes) and returns the total
 * number produced (may exceed cap). */
uint32_t b2j_events_drain(b2j_world *w, b2j_contact_event *out, uint32_t cap);
uint32_t b2j_activation_events_drain(b2j_world *w, b2j_activation_event *out, uint32_t cap);

/* PhysicsSystem::SetContactListener / SetBodyActivationListener (PhysicsSystem.h:62-68) as seen by the device: with a kind of
 * recording off the step writes no such events (no buffer traffic, the drains return 0). Default: both on for a world made by
 * b2j_world_create (buffers are allocated on first use), always off for batched worlds. */
int b2j_world_set_event_recording(b2j_world *w, int contact_events, int activation_events);

/* ---- parity hooks (debug; read the intermediate products of the LAST step) ------------------------------------ */

/* Candidate body pairs of the broadphase as (min id, max id), sorted; returns the total count. */
uint32_t b2j_debug_get_pairs(b2j_world *w, uint32_t *pairs /* [cap][2] */, uint32_t cap);

typedef struct b2j_debug_manifold {
	uint32_t body1, body2, sub_shape1, sub_shape2;
	uint32_t num_points;
	uint32_t from_cache;           /* 1 if copied by the body-pair cache */
	float    normal[3];            /* world space */
	float    penetration_depth;
} b2j_debug_manifold;
/* Manifolds written this step sorted by (body1, body2, sub1, sub2); returns the total count. */
uint32_t b2j_debug_get_manifolds(b2j_world *w, b2j_debug_manifold *out, uint32_t cap);

/* Checks the solve schedule of the LAST step: constraints of one phase run in parallel, so no dynamic body may be touched by two
 * constraints of the same phase (the reference asserts the same of its splits, LargeIslandSplitter.cpp). Returns the number of
 * (body, phase) conflicts -- 0 for a valid schedule -- or < 0 on a CUDA failure. */
int b2j_debug_check_schedule(b2j_world *w);

/* Only the broadphase of a step on the current state: fills the pair list for b2j_debug_get_pairs. */
int b2j_debug_find_pairs(b2j_world *w);

/* ---- per kernel timing (measurement hook for bench.py; CUDA events around every launch on the world's stream) ---- */
int      b2j_world_set_profiling(b2j_world *w, int on);   /* on != 0: start (resets the accumulators), 0: stop */
/* Accumulated device time / launch count per kernel since profiling was switched on. names: [cap][name_stride] chars.
 * Returns the number of kernel categories that ran. */
uint32_t b2j_world_get_profile(b2j_world *w, char *names, uint32_t name_stride, float *ms, uint32_t *launches, uint32_t cap);

/* ---- batched independent worlds (config 5) --------------------------------------------------------- */

typedef struct b2j_batch b2j_batch; /* opaque: n independent worlds with identical layout stepped together on one device */

/* Clones the bodies / shapes / active list of `proto` n_worlds times on its device (the contact cache is NOT cloned: use a
 * freshly populated prototype). All worlds live in one device world (slot = world * stride + body index) and never interact:
 * the broadphase tree is cut per world, body pair keys carry the world. Per world limits (0 = the prototype's) size the shared
 * caches. The prototype stays usable on its own. Contact / activation events are not recorded for batches.
 * Every world evolves bit-identically to the prototype stepped alone (TestMultiplePhysicsSystems pattern, PhysicsTests.cpp:1548). */
b2j_batch *b2j_batch_create(b2j_world *proto, uint32_t n_worlds, uint32_t max_body_pairs_per_world, uint32_t max_contact_constraints_per_world);
/* Library level multi device batch: the worlds are split in contiguous blocks over the given CUDA devices of THIS
 * process (device_ids[0] must be the prototype's device; the others need peer access to it), each block in groups as above. Stepping
 * moves no data between the devices: every group is stepped by its own host thread on its own device and stream, b2j_batch_step
 * sums the statistics on the host. (One process per GPU with the statistics reduced over NCCL is the other deployment: bench.py.) */
b2j_batch *b2j_batch_create_on_devices(b2j_world *proto, uint32_t n_worlds, const int32_t *device_ids, uint32_t n_devices,
                                       uint32_t max_body_pairs_per_world, uint32_t max_contact_constraints_per_world);
void       b2j_batch_destroy(b2j_batch *b);
/* Steps every world of the batch once; stats (may be NULL) receives the totals over all worlds. */
int        b2j_batch_step(b2j_batch *b, float delta_time, int collision_steps, b2j_step_stats *stats);
/* Resets the given worlds to the state the batch was created with (bodies + an empty contact cache), on the device: the RL
 * environment reset. A reset world evolves exactly like a newly created one; the other worlds are untouched. (the
 * reference does this with SaveState / RestoreState through host memory, PhysicsSystem.h:165-168.) */
int        b2j_batch_reset_worlds(b2j_batch *b, const uint32_t *world_indices, uint32_t n);
uint32_t   b2j_batch_size(const b2j_batch *b);
/* State of the bodies in slots [0, n) of one world (see b2j_bodies_get_state with ids == NULL). world_index = 0xffffffff:
 * the first n slots of the whole batch (world major, stride = body slots of the prototype), i.e. all worlds in one call. */
int        b2j_batch_get_state(b2j_batch *b, uint32_t world_index, uint32_t n, const b2j_body_state *out);
/* BodyInterface::AddForce / AddTorque for the first n slots of the whole batch (world major); arrays [n][3], either may be NULL. */
int        b2j_batch_add_force_torque(b2j_batch *b, uint32_t n, const float *force, const float *torque);
/* b2j_query_cast_rays for batched worlds: ray i is cast in world ray_world[i] (the RL observation pattern). */
int        b2j_batch_query_cast_rays(b2j_batch *b, const uint32_t *ray_world, const b2j_ray *rays, uint32_t n, uint32_t object_layer, b2j_ray_hit *hits);
/* SaveState / RestoreState of every world of the batch (device resident, see b2j_world_save_state). */
b2j_snapshot *b2j_batch_save_state(b2j_batch *b);
int        b2j_batch_restore_state(b2j_batch *b, const b2j_snapshot *s);
int        b2j_batch_set_profiling(b2j_batch *b, int on);
uint32_t   b2j_batch_get_profile(b2j_batch *b, char *names, uint32_t name_stride, float *ms, uint32_t *launches, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* JOLT_B200_H */
