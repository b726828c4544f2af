// jolt_b200.cu -- world management, the step (PhysicsSystem::Update replacement) and the C ABI of include/jolt_b200.h.
//
// One translation unit: all kernels are header-defined functors (b2j_*.h) instantiated through Runtime::launch*.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off (see __graft_entry__.build()).
#ifndef B2J_HOSTSIM
#include <cooperative_groups.h>
#endif
#include "b2j_runtime.h"
#include "b2j_world.h"
#include "b2j_shapes.h"
#include "b2j_broadphase.h"
#include "b2j_narrowphase.h"
#include "b2j_mesh.h"
#include "b2j_compound.h"
#include "b2j_solver.h"
#include "b2j_joints.h"
#include "b2j_query.h"

#ifndef B2J_HOSTSIM
#include <nvtx3/nvToolsExt.h>   // header only: ranges are no-ops unless a tool (nsys / ncu --nvtx) is attached
#endif
#include <chrono>
#include <thread>
#include <map>
#include <string>

using namespace b2j;

// Every C ABI entry that touches a world first selects the world's device (calls may come from any thread, and a process may hold
// worlds on several devices: the reference's PhysicsSystem has no such affinity, so the boundary must not have one either)
#ifndef B2J_HOSTSIM
#define B2J_DEVICE_GUARD(W) do { cudaSetDevice((W)->rt.device); } while (0)
#else
#define B2J_DEVICE_GUARD(W) do { } while (0)
#endif

// ---- small API kernels ---------------------------------------------------------------------------------------------
namespace b2j {

struct KAddBodies
{
	DWorld w; const b2j_body_desc *descs;
	B2J_D void operator()(uint32_t i) const
	{
		const b2j_body_desc &d = descs[i];
		uint32_t b = slot_of(d.id);
		BodyInfo info;
		info.id = d.id; info.shape = d.shape; info.object_layer = d.object_layer; info.motion_type = d.motion_type;
		info.bp_layer = w.object_to_bp[d.object_layer];
		info.flags = d.flags; info.allowed_dofs = d.allowed_dofs;
		uint32_t vs = d.num_velocity_steps_override > 15? 15 : d.num_velocity_steps_override;
		uint32_t ps = d.num_position_steps_override > 15? 15 : d.num_position_steps_override;
		info.steps_override = (uint8_t)(vs | (ps << 4));
		w.info[b] = info;
		BodyParams p;
		p.inv_mass = d.motion_type == B2J_MOTION_DYNAMIC? d.inv_mass : 0.0f;
		p.linear_damping = d.linear_damping; p.angular_damping = d.angular_damping;
		p.max_linear_velocity = d.max_linear_velocity; p.max_angular_velocity = d.max_angular_velocity;
		p.gravity_factor = d.gravity_factor; p.friction = d.friction; p.restitution = d.restitution;
		w.params[b] = p;
		V3 x = v3_load(d.position);
		Q4 q = q4_load(d.rotation);
		w.position[b] = f4(x);
		w.rotation[b] = f4(q);
		w.linear_velocity[b] = f4(v3_load(d.linear_velocity));
		w.angular_velocity[b] = f4(v3_load(d.angular_velocity));
		w.force[b] = f4(v3_load(d.force));
		w.torque[b] = f4(v3_load(d.torque));
		w.inv_inertia_diag[b] = f4(v3_load(d.inv_inertia_diag));
		w.inertia_rotation[b] = f4(q4_load(d.inertia_rotation));
		w.active_index[b] = B2J_INACTIVE_INDEX;
		const ShapeDesc &s = w.shapes[d.shape];
		if (d.has_bounds)
		{
			w.bounds_min[b] = f4(v3_load(d.bounds_min));
			w.bounds_max[b] = f4(v3_load(d.bounds_max));
			for (int k = 0; k < 3; ++k)
				w.sleep_spheres[b * 3 + k] = f4(d.sleep_spheres[k][0], d.sleep_spheres[k][1], d.sleep_spheres[k][2], d.sleep_spheres[k][3]);
			w.sleep_timer[b] = d.sleep_timer;
		}
		else
		{
			V3 mn, mx;
			world_bounds(w, s, x, q, mn, mx);
			w.bounds_min[b] = f4(mn);
			w.bounds_max[b] = f4(mx);
			V3 pts[3];
			sleep_test_points(s, x, q, pts);
			for (int k = 0; k < 3; ++k) w.sleep_spheres[b * 3 + k] = f4(pts[k], 0.0f);
			w.sleep_timer[b] = 0.0f;
		}
	}
};

struct KSetActive
{
	DWorld w; const uint32_t *ids; uint32_t base;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = slot_of(ids[i]);
		w.active[base + i] = b;
		w.active_index[b] = base + i;
	}
};

struct KClearActiveIndex
{
	DWorld w;
	B2J_D void operator()(uint32_t ai) const { w.active_index[w.active[ai]] = B2J_INACTIVE_INDEX; }
};

// b2j_bodies_deactivate / b2j_bodies_remove on the device: keep[ai] = 0 for the listed bodies that are active (ids are validated
// against the slot's current id: a stale id never touches the body that lives in the slot now)
struct KFillU32
{
	uint32_t *dst; uint32_t value;
	B2J_D void operator()(uint32_t i) const { dst[i] = value; }
};

struct KMarkDeactivate
{
	DWorld w; const uint32_t *ids; uint32_t *keep;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = slot_of(ids[i]);
		if (b >= w.max_bodies || w.info[b].id != ids[i]) return;
		uint32_t ai = w.active_index[b];
		if (ai != B2J_INACTIVE_INDEX) keep[ai] = 0;
	}
};

// BodyManager::DeactivateBodies (BodyManager.cpp:529-568): leaves the active list, velocities reset
struct KApiDeactivate
{
	DWorld w; const uint32_t *keep;
	B2J_D void operator()(uint32_t ai) const
	{
		if (keep[ai] != 0) return;
		uint32_t b = w.active[ai];
		w.linear_velocity[b] = f4(0, 0, 0, 0);
		w.angular_velocity[b] = f4(0, 0, 0, 0);
		w.active_index[b] = B2J_INACTIVE_INDEX;
	}
};

struct KClearBodies
{
	DWorld w; const uint32_t *ids;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = slot_of(ids[i]);
		if (b < w.max_bodies && w.info[b].id == ids[i]) w.info[b].id = B2J_INVALID_ID;
	}
};

// PhysicsSystem::WereBodiesInContact through the device pair table of the read cache
struct KWereInContact
{
	DWorld w; uint32_t id1, id2; uint32_t *out;
	B2J_D void operator()(uint32_t) const
	{
		uint32_t e = pair_table_find(w, w.read_cache, slot_of(id1), slot_of(id2));
		*out = (e != 0xffffffffu && w.read_cache.pairs[e].body1 == id1 && w.read_cache.pairs[e].body2 == id2 && w.read_cache.pairs[e].num_manifolds > 0)? 1u : 0u;
	}
};

struct KGetState
{
	DWorld w; const uint32_t *ids; float *pos, *rot, *lin, *ang, *bounds; uint32_t *active_index; float *sleep_timer; uint32_t first;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = ids != nullptr? slot_of(ids[i]) : first + i;
		if (pos) v3_store(to_v3(w.position[b]), pos + 3 * i);
		if (rot) q4_store(to_q4(w.rotation[b]), rot + 4 * i);
		if (lin) v3_store(to_v3(w.linear_velocity[b]), lin + 3 * i);
		if (ang) v3_store(to_v3(w.angular_velocity[b]), ang + 3 * i);
		if (bounds) { v3_store(to_v3(w.bounds_min[b]), bounds + 6 * i); v3_store(to_v3(w.bounds_max[b]), bounds + 6 * i + 3); }
		if (active_index) active_index[i] = w.active_index[b];
		if (sleep_timer) sleep_timer[i] = w.sleep_timer[b];
	}
};

// b2j_bodies_get_stepped_state: like KGetState over a list of body slots, the ids come back too
struct KGetStepped
{
	DWorld w; const uint32_t *slots; uint32_t *ids; float *pos, *rot, *lin, *ang, *bounds; uint32_t *active_index; float *sleep_timer;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = slots[i];
		ids[i] = w.info[b].id;
		if (pos) v3_store(to_v3(w.position[b]), pos + 3 * i);
		if (rot) q4_store(to_q4(w.rotation[b]), rot + 4 * i);
		if (lin) v3_store(to_v3(w.linear_velocity[b]), lin + 3 * i);
		if (ang) v3_store(to_v3(w.angular_velocity[b]), ang + 3 * i);
		if (bounds) { v3_store(to_v3(w.bounds_min[b]), bounds + 6 * i); v3_store(to_v3(w.bounds_max[b]), bounds + 6 * i + 3); }
		if (active_index) active_index[i] = w.active_index[b];
		if (sleep_timer) sleep_timer[i] = w.sleep_timer[b];
	}
};

struct KSetState
{
	DWorld w; const uint32_t *ids; const float *pos, *rot, *lin, *ang;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = ids != nullptr? slot_of(ids[i]) : i;
		if (pos) w.position[b] = f4(v3_load(pos + 3 * i));
		if (rot) w.rotation[b] = f4(q4_load(rot + 4 * i));
		if (lin) w.linear_velocity[b] = f4(v3_load(lin + 3 * i));
		if (ang) w.angular_velocity[b] = f4(v3_load(ang + 3 * i));
		if (pos || rot)
		{
			V3 mn, mx;
			world_bounds(w, w.shapes[w.info[b].shape], to_v3(w.position[b]), to_q4(w.rotation[b]), mn, mx);
			w.bounds_min[b] = f4(mn);
			w.bounds_max[b] = f4(mx);
		}
	}
};

struct KSetParams
{
	DWorld w; const uint32_t *ids; const float *friction, *restitution, *gravity_factor, *linear_damping, *angular_damping, *max_linear_velocity, *max_angular_velocity;
	B2J_D void operator()(uint32_t i) const
	{
		BodyParams &p = w.params[slot_of(ids[i])];
		if (friction) p.friction = friction[i];
		if (restitution) p.restitution = restitution[i];
		if (gravity_factor) p.gravity_factor = gravity_factor[i];
		if (linear_damping) p.linear_damping = linear_damping[i];
		if (angular_damping) p.angular_damping = angular_damping[i];
		if (max_linear_velocity) p.max_linear_velocity = max_linear_velocity[i];
		if (max_angular_velocity) p.max_angular_velocity = max_angular_velocity[i];
	}
};

// b2j_bodies_set_info: Body::SetMotionType / SetObjectLayerInternal / SetShapeInternal on the device
struct KSetInfo
{
	DWorld w; const uint32_t *ids; const uint8_t *motion_type; const float *inv_mass; const uint16_t *object_layer; const int32_t *shape;
	const float *inv_inertia_diag, *inertia_rotation; uint32_t invalidate; const uint16_t *flags_set, *flags_clear;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = slot_of(ids[i]);
		BodyInfo info = w.info[b];
		const uint16_t settable = B2J_BODY_SENSOR | B2J_BODY_ALLOW_SLEEPING | B2J_BODY_USE_MANIFOLD_REDUCTION | B2J_BODY_GYROSCOPIC | B2J_BODY_KIN_VS_NONDYN;
		if (flags_clear != nullptr) info.flags &= (uint16_t)~(flags_clear[i] & settable);
		if (flags_set != nullptr) info.flags |= (uint16_t)(flags_set[i] & settable);
		if (motion_type != nullptr && motion_type[i] != info.motion_type)
		{
			info.motion_type = motion_type[i];
			w.params[b].inv_mass = info.motion_type == B2J_MOTION_DYNAMIC? inv_mass[i] : 0.0f;
			if (info.motion_type == B2J_MOTION_STATIC) { w.linear_velocity[b] = f4(0, 0, 0, 0); w.angular_velocity[b] = f4(0, 0, 0, 0); }
			if (info.motion_type != B2J_MOTION_DYNAMIC) { w.force[b] = f4(0, 0, 0, 0); w.torque[b] = f4(0, 0, 0, 0); }
		}
		if (object_layer != nullptr)
		{
			info.object_layer = object_layer[i];
			info.bp_layer = w.object_to_bp[object_layer[i]];
		}
		if (shape != nullptr && shape[i] != info.shape)
		{
			// Body::SetShapeInternal -> UpdateCenterOfMassInternal: mPosition += mRotation * (new centre of mass - old centre of mass)
			V3 old_com = w.shapes[info.shape].center_of_mass, new_com = w.shapes[shape[i]].center_of_mass;
			info.shape = shape[i];
			Q4 q = to_q4(w.rotation[b]);
			V3 x = to_v3(w.position[b]) + rotate(q, new_com - old_com);
			w.position[b] = f4(x);
			if (inv_inertia_diag != nullptr)
			{
				if (info.motion_type == B2J_MOTION_DYNAMIC && inv_mass != nullptr) w.params[b].inv_mass = inv_mass[i];
				w.inv_inertia_diag[b] = f4(v3_load(inv_inertia_diag + 3 * i));
				w.inertia_rotation[b] = f4(q4_load(inertia_rotation + 4 * i));
			}
			V3 mn, mx;
			world_bounds(w, w.shapes[info.shape], x, q, mn, mx);
			w.bounds_min[b] = f4(mn);
			w.bounds_max[b] = f4(mx);
			info.flags |= B2J_BODY_INVALIDATE_CACHE;
		}
		if (invalidate) info.flags |= B2J_BODY_INVALIDATE_CACHE;
		w.info[b] = info;
	}
};

// BodyManager::ValidateContactCacheForAllBodies at the end of a step (PhysicsSystem.cpp JobContactRemovedCallbacks)
struct KValidateContactCache
{
	DWorld w; const uint32_t *slots;
	B2J_D void operator()(uint32_t i) const { w.info[slots[i]].flags &= (uint16_t)~B2J_BODY_INVALIDATE_CACHE; }
};

struct KAddForceTorque
{
	DWorld w; const uint32_t *ids; const float *force, *torque;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t b = ids != nullptr? slot_of(ids[i]) : i;
		if (force) w.force[b] = f4(to_v3(w.force[b]) + v3_load(force + 3 * i));
		if (torque) w.torque[b] = f4(to_v3(w.torque[b]) + v3_load(torque + 3 * i));
	}
};

struct KImportCache
{
	DWorld w; const b2j_cached_body_pair *pairs; const b2j_cached_manifold *manifolds;
	B2J_D void operator()(uint32_t i) const
	{
		const b2j_cached_body_pair &p = pairs[i];
		CachedPair &o = w.read_cache.pairs[i];
		o.body1 = p.body1; o.body2 = p.body2;
		o.slot1 = slot_of(p.body1); o.slot2 = slot_of(p.body2);
		for (int k = 0; k < 3; ++k) { o.dpos[k] = p.delta_position[k]; o.drot[k] = p.delta_rotation[k]; }
		o.first_manifold = p.first_manifold; o.num_manifolds = p.num_manifolds;
		for (uint32_t j = 0; j < p.num_manifolds; ++j)
		{
			const b2j_cached_manifold &m = manifolds[p.first_manifold + j];
			CachedManifold &cm = w.read_cache.manifolds[p.first_manifold + j];
			cm.body1 = p.body1; cm.body2 = p.body2; cm.sub1 = m.sub_shape1; cm.sub2 = m.sub_shape2;
			for (int k = 0; k < 3; ++k) cm.normal[k] = m.normal[k];
			cm.friction_lambda[0] = m.friction_lambda[0]; cm.friction_lambda[1] = m.friction_lambda[1];
			cm.angular_lambda = m.angular_friction_lambda;
			cm.num_points = (uint16_t)(m.num_points > 4? 4 : m.num_points);
			cm.flags = 0;
			for (int q = 0; q < 4; ++q)
			{
				for (int k = 0; k < 3; ++k) { cm.p1[q][k] = m.position1[q][k]; cm.p2[q][k] = m.position2[q][k]; }
				cm.lambda[q] = m.non_penetration_lambda[q];
			}
		}
		pair_table_insert(w, w.read_cache, slot_of(p.body1), slot_of(p.body2), i);
	}
};

struct KExportCache
{
	DWorld w; b2j_cached_body_pair *pairs; b2j_cached_manifold *manifolds;
	B2J_D void operator()(uint32_t i) const
	{
		const CachedPair &p = w.read_cache.pairs[i];
		b2j_cached_body_pair &o = pairs[i];
		o.body1 = p.body1; o.body2 = p.body2;
		for (int k = 0; k < 3; ++k) { o.delta_position[k] = p.dpos[k]; o.delta_rotation[k] = p.drot[k]; }
		o.first_manifold = p.first_manifold; o.num_manifolds = p.num_manifolds;
		for (uint32_t j = 0; j < p.num_manifolds; ++j)
		{
			const CachedManifold &cm = w.read_cache.manifolds[p.first_manifold + j];
			b2j_cached_manifold &m = manifolds[p.first_manifold + j];
			m.sub_shape1 = cm.sub1; m.sub_shape2 = cm.sub2;
			for (int k = 0; k < 3; ++k) m.normal[k] = cm.normal[k];
			m.friction_lambda[0] = cm.friction_lambda[0]; m.friction_lambda[1] = cm.friction_lambda[1];
			m.angular_friction_lambda = cm.angular_lambda;
			m.num_points = cm.num_points;
			m.flags = cm.flags;
			for (int q = 0; q < 4; ++q)
			{
				for (int k = 0; k < 3; ++k) { m.position1[q][k] = cm.p1[q][k]; m.position2[q][k] = cm.p2[q][k]; }
				m.non_penetration_lambda[q] = cm.lambda[q];
			}
		}
	}
};

// batched worlds: dst[world * dst_stride + j] = src[j] for j < n_src
template <class T> struct KReplicate
{
	T *dst; const T *src; uint32_t n_src, dst_stride;
	B2J_D void operator()(uint32_t i) const { uint32_t wi = i / n_src, j = i % n_src; dst[(size_t)wi * dst_stride + j] = src[j]; }
};

// batched worlds: the active list of the prototype repeated per world (world major keeps the order inside every world)
struct KReplicateActive
{
	DWorld w; const uint32_t *src_active; uint32_t na, stride;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t wi = i / na, k = i % na;
		uint32_t slot = src_active[k] + wi * stride;
		w.active[i] = slot;
		w.active_index[slot] = i;
	}
};

// ---- batch: reset worlds to the state they were created with (RL environment reset without a host round trip of the state) ----
struct BatchInitState
{
	F4 *pose, *velocity, *force_torque, *bounds, *sleep_spheres;   // pair arrays like DWorld's (2 elements per body)
	float *sleep_timer;
	uint32_t *was_active;        // per slot of ONE world
};

// bodies of the reset worlds that were active at creation but sleep now go to the woken list (activated before the copy)
struct KResetFindInactive
{
	DWorld w; NarrowCtx c; BatchInitState init; const uint32_t *worlds;
	B2J_D void operator()(uint32_t t) const
	{
		uint32_t j = t % w.world_stride, b = worlds[t / w.world_stride] * w.world_stride + j;
		if (init.was_active[j] != 0 && w.info[b].id != 0xffffffffu && w.active_index[b] == B2J_INACTIVE_INDEX)
			wake_body(w, c, b);
	}
};

struct KResetWorlds
{
	DWorld w; BatchInitState init; const uint32_t *worlds;
	B2J_D void operator()(uint32_t t) const
	{
		uint32_t j = t % w.world_stride, b = worlds[t / w.world_stride] * w.world_stride + j;
		w.position[b] = init.pose[2 * j]; w.rotation[b] = init.pose[2 * j + 1];
		w.linear_velocity[b] = init.velocity[2 * j]; w.angular_velocity[b] = init.velocity[2 * j + 1];
		w.force[b] = init.force_torque[2 * j]; w.torque[b] = init.force_torque[2 * j + 1];
		w.bounds_min[b] = init.bounds[2 * j]; w.bounds_max[b] = init.bounds[2 * j + 1];
		for (int i = 0; i < 3; ++i) w.sleep_spheres[b * 3 + i] = init.sleep_spheres[j * 3 + i];
		w.sleep_timer[b] = init.sleep_timer[j];
	}
};

// the non contact constraints of the reset worlds go back to their state at creation
struct KResetJoints
{
	JointState *state; const JointState *init; const uint32_t *worlds; uint32_t per_world;
	B2J_D void operator()(uint32_t t) const { uint32_t j = t % per_world; state[(size_t)worlds[t / per_world] * per_world + j] = init[j]; }
};

// bodies of the reset worlds that were NOT active at creation leave the active list (keep = 0), as KDeactivate does
struct KResetMarkKeep
{
	DWorld w; BatchInitState init; const uint32_t *reset_flag; uint32_t *keep;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		bool drop = reset_flag[b / w.world_stride] != 0 && init.was_active[b % w.world_stride] == 0;
		keep[ai] = drop? 0u : 1u;
		if (drop)
			w.active_index[b] = B2J_INACTIVE_INDEX;
	}
};

// the contact cache entries of the reset worlds must not be found by the next step (a new world has an empty cache)
struct KResetPurgeCache
{
	DWorld w; const uint32_t *reset_flag;
	B2J_D void operator()(uint32_t i) const
	{
		CachedPair &p = w.read_cache.pairs[i];
		if (p.slot1 != 0xffffffffu && reset_flag[p.slot1 / w.world_stride] != 0)
			p.slot1 = 0xffffffffu;
	}
};

// round bookkeeping on the device: pairs found so far become "processed", work lists restart
struct KNextRound
{
	DWorld w; uint32_t *round_begin;
	B2J_D void operator()(uint32_t) const
	{
		StepCounters &c = *w.counters;
		uint32_t n = c.num_pairs < w.max_body_pairs? c.num_pairs : w.max_body_pairs;
		if (c.num_pairs > w.max_body_pairs) c.error_bits |= B2J_ERR_BODY_PAIR_CACHE_FULL;
		*round_begin = n;
		*w.write_cache.num_pairs = n; // cache entries = processed pairs (KProcessPairs)
		c.num_collide_convex = 0; c.num_collide_mesh = 0; c.num_cached = 0; c.num_epa = 0; c.num_woken = 0;
	}
};

// RestoreState: the pair table of the restored read cache is rebuilt from its pairs (a snapshot does not store the table)
struct KRebuildPairTable
{
	DWorld w;
	B2J_D void operator()(uint32_t i) const
	{
		const CachedPair &p = w.read_cache.pairs[i];
		if (p.slot1 != 0xffffffffu) // (entries of reset worlds were made unfindable: b2j_batch_reset_worlds)
			pair_table_insert(w, w.read_cache, p.slot1, p.slot2, i);
	}
};

// the sizes of the write cache join the step counters (one readback at the end of the step)
struct KPublishCacheCounts
{
	DWorld w;
	B2J_D void operator()(uint32_t) const { w.counters->cache_pairs = *w.write_cache.num_pairs; w.counters->cache_manifolds = *w.write_cache.num_manifolds; }
};

struct KGatherSortKeys
{
	SolveCtx s; uint64_t *keys; uint32_t *vals;
	B2J_D void operator()(uint32_t i) const { keys[i] = s.src[i].sort_key; vals[i] = i; }
};

struct KCountTies
{
	DWorld w; const uint64_t *keys;
	B2J_D void operator()(uint32_t i) const { if (i > 0 && keys[i] == keys[i - 1]) atomic_add(&w.counters->hash_tie, 1u); }
};

struct KFinishCompact
{
	DWorld w; const uint32_t *keep, *keep_scan; uint32_t n;
	B2J_D void operator()(uint32_t) const { w.counters->new_active_count = n == 0? 0 : keep_scan[n - 1] + keep[n - 1]; }
};

struct KKineticEnergy
{
	DWorld w; float *out;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		if (w.info[b].motion_type != B2J_MOTION_DYNAMIC) return;
		BodyParams p = w.params[b];
		float e = 0.0f;
		if (p.inv_mass > 0.0f) e += 0.5f * length_sq(to_v3(w.linear_velocity[b])) / p.inv_mass;
		V3 wl = inverse_rotate(to_q4(w.rotation[b]) * to_q4(w.inertia_rotation[b]), to_v3(w.angular_velocity[b]));
		V3 d = to_v3(w.inv_inertia_diag[b]);
		if (d.x > 0.0f) e += 0.5f * wl.x * wl.x / d.x;
		if (d.y > 0.0f) e += 0.5f * wl.y * wl.y / d.y;
		if (d.z > 0.0f) e += 0.5f * wl.z * wl.z / d.z;
#ifndef B2J_HOSTSIM
		atomicAdd(out, e);
#else
		*out += e;
#endif
	}
};

} // namespace b2j

// ---- the world -------------------------------------------------------------------------------------------------------

struct b2j_world
{
	Runtime rt;
	DWorld d;
	b2j_world_desc desc;
	std::vector<uint8_t> t_o2bp, t_ovbp, t_ovo;
	uint8_t *d_o2bp = nullptr, *d_ovbp = nullptr, *d_ovo = nullptr;

	// host mirrors
	std::vector<uint32_t> h_ids;           // per slot: id or invalid
	std::vector<uint8_t> h_layer;          // per slot: broadphase layer
	std::vector<uint8_t> h_static;         // per slot: 1 = static body (never on the active list)
	std::vector<uint8_t> h_mark;           // per slot scratch marks of the bulk API calls (all zero between calls)
	std::vector<std::vector<uint32_t>> layer_bodies;
	std::vector<uint8_t> layer_list_dirty, layer_needs_build, layer_has_moving;
	uint32_t num_bodies = 0, num_active = 0, num_slots = 0;
	uint32_t num_worlds = 1;               // > 1: batched independent worlds (b2j_batch)
	uint32_t batch_groups = 1;             // groups of the batch this world belongs to (their cooperative constraint solvers share the SMs)
	uint32_t solve_grid_div = 1;           // the one launch solvers use 1 / solve_grid_div of the SMs (groups of a batch that solve concurrently)
	uint32_t get_state_first = 0;          // slot offset of b2j_bodies_get_state with ids == NULL (batch world selection)

	// shapes
	std::vector<ShapeDesc> h_shapes;
	// host side bookkeeping per shape for the decorators: the leaf a decorated shape was derived from, what has been accumulated on the way
	// down to it (ScaledShape: inScale * mScale, RotatedTranslatedShape: transform * rotation) and the hull connectivity the scaled shrink needs
	struct ShapeMeta { int32_t leaf = -1; V3 scale = { 1.0f, 1.0f, 1.0f }; bool scaled = false; Q4 rotation = { 0.0f, 0.0f, 0.0f, 1.0f }; bool rotated = false; std::vector<int32_t> point_num_faces, point_faces; };
	std::vector<ShapeMeta> h_shape_meta;
	std::vector<F4> h_hull_points, h_hull_shrunk, h_hull_planes;
	std::vector<uint32_t> h_hull_faces;
	std::vector<uint8_t> h_hull_vtx, h_mesh_bytes;
	std::vector<CompoundSub> h_compound_subs;
	CompoundSub *d_compound_subs = nullptr;
	bool shapes_dirty = false;
	ShapeDesc *d_shapes = nullptr; F4 *d_hull_points = nullptr, *d_hull_shrunk = nullptr, *d_hull_planes = nullptr;
	uint32_t *d_hull_faces = nullptr; uint8_t *d_hull_vtx = nullptr, *d_mesh_bytes = nullptr;

	// broadphase
	Tree trees[8];
	uint32_t tree_capacity[8] = { 0 };
	uint32_t *d_tree_vals_in = nullptr; uint32_t tree_vals_capacity = 0;

	// active list double buffer
	uint32_t *active_buf[2] = { nullptr, nullptr };
	int active_cur = 0;
	uint32_t *d_keep = nullptr, *d_keep_scan = nullptr;

	// contact caches
	ContactCache cache[2];
	int write_idx = 0;
	uint32_t cache_num_pairs[2] = { 0, 0 }, cache_num_manifolds[2] = { 0, 0 };

	// narrow phase + solver work memory
	NarrowCtx nc;
	SolveCtx sc;
	uint32_t *d_round_begin = nullptr;
	MeshScratch *d_mesh_scratch = nullptr;   // allocated when the first mesh shape is uploaded
	MeshScratch *d_query_scratch = nullptr;  // one block for b2j_query_collide_shape (the pair functions want a reference; queries never write it)
	uint64_t *d_sort_keys[2] = { nullptr, nullptr };
	uint32_t *d_sort_vals = nullptr;
	uint32_t *d_woken_sorted = nullptr;
	uint32_t *d_collide_keys[2] = { nullptr, nullptr }, *d_collide_vals[2] = { nullptr, nullptr }; // batch groups: ordering of the convex pair queue
	uint32_t *d_woken_keys = nullptr;
	b2j_activation_event *d_act_events = nullptr;
	uint32_t max_events = 0, max_act_events = 0;
	b2j_contact_event *events_buf = nullptr;            // owned buffers; nc.events / d_act_events point at them while recording is on
	b2j_activation_event *act_events_buf = nullptr;
	float *d_energy = nullptr;

	// non contact constraints (b2j_joints.h), by constraint index; device arrays grow with the list
	std::vector<b2j_constraint_desc> h_joints;
	JointCtx jc = { };
	uint32_t joint_capacity = 0;
	bool joints_dirty = false;               // definitions / order changed since the last upload
	JointState *d_joint_init = nullptr;      // batch group: the constraint state of one world at creation (b2j_batch_reset_worlds)

	float prev_dt = 0.0f;
	StepCounters h_counters;
	std::vector<uint32_t> h_phase_offsets;
	uint32_t last_num_events = 0, last_num_act_events = 0, last_num_pairs = 0;
	std::vector<uint32_t> h_cache_invalid;  // slots whose InvalidateContactCache flag is set (cleared after the next step, BodyManager::mBodiesCacheInvalid)
	uint32_t *d_cache_invalid = nullptr; uint32_t cache_invalid_capacity = 0;
	const uint32_t *stepped_list = nullptr; uint32_t stepped_count = 0; // body slots the last step simulated (b2j_bodies_get_stepped_state)
	uint32_t last_collide_convex = 0;      // longest convex pair queue of the previous steps, halved per step (sizes this step's queue ordering)
#ifndef B2J_HOSTSIM
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
#endif
};

// The worlds of a batch live in a few GROUPS; a group is ONE device world (slot = world in group * stride + body index) with its own
// stream, stepped by its own host thread: kernels and host round trips of different groups overlap (measured 1.3x at 1024 worlds).
struct b2j_batch
{
	BatchInitState init = BatchInitState();   // state of one world at creation (device, owned by group 0's runtime)
	bool has_init_inactive = false;           // some non-static body was NOT active at creation (a reset then also compacts the active list)
	std::vector<b2j_world *> groups;
	std::vector<uint32_t> first_world;   // first world of each group (+ n_worlds at the end)
	uint32_t n_worlds = 0, stride = 0, bodies_per_world = 0;
};

// Device resident snapshot of one world (PhysicsSystem::SaveState with Global | Bodies | Contacts + the active list order); a batch
// snapshot holds one per group.
struct WorldSnapshot
{
	b2j_world *owner = nullptr;
	uint32_t num_slots = 0, num_active = 0, num_bodies = 0, num_pairs = 0, num_manifolds = 0;
	float prev_dt = 0.0f;
	V3 gravity;
	BodyInfo *info = nullptr; BodyParams *params = nullptr;
	F4 *pose = nullptr, *velocity = nullptr, *force_torque = nullptr, *inertia = nullptr, *bounds = nullptr, *sleep_spheres = nullptr;
	float *sleep_timer = nullptr; uint32_t *active_index = nullptr, *active = nullptr;
	CachedPair *pairs = nullptr; CachedManifold *manifolds = nullptr;
	JointState *joint_state = nullptr; std::vector<b2j_constraint_desc> h_joints; // non contact constraints: the list and its state
	// host mirrors of the world (the set of bodies is part of the state)
	std::vector<uint32_t> h_ids; std::vector<uint8_t> h_layer, h_static, layer_has_moving;
	std::vector<std::vector<uint32_t>> layer_bodies;
	uint64_t bytes = 0;
};
struct b2j_snapshot
{
	b2j_batch *batch = nullptr;           // non null: a batch snapshot (one WorldSnapshot per group)
	std::vector<WorldSnapshot> worlds;
};

namespace {

template <class T> bool grow(Runtime &rt, T *&ptr, uint32_t &capacity, uint32_t needed, bool keep)
{
	if (needed <= capacity) return true;
	uint32_t ncap = capacity == 0? needed : capacity;
	while (ncap < needed) ncap *= 2;
	T *np = rt.alloc<T>(ncap);
	if (np == nullptr) return false;
	if (keep && ptr != nullptr) rt.copy(np, ptr, capacity);
	rt.sync();
	rt.free_(ptr);
	ptr = np;
	capacity = ncap;
	return true;
}

uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

void upload_shapes(b2j_world *W)
{
	if (!W->shapes_dirty) return;
	Runtime &rt = W->rt;
	rt.sync();
	rt.free_(W->d_shapes); rt.free_(W->d_hull_points); rt.free_(W->d_hull_shrunk); rt.free_(W->d_hull_planes);
	rt.free_(W->d_hull_faces); rt.free_(W->d_hull_vtx); rt.free_(W->d_mesh_bytes); rt.free_(W->d_compound_subs);
	W->d_compound_subs = rt.alloc<CompoundSub>(W->h_compound_subs.size()); rt.upload(W->d_compound_subs, W->h_compound_subs.data(), W->h_compound_subs.size());
	W->d_shapes = rt.alloc<ShapeDesc>(W->h_shapes.size()); rt.upload(W->d_shapes, W->h_shapes.data(), W->h_shapes.size());
	W->d_hull_points = rt.alloc<F4>(W->h_hull_points.size()); rt.upload(W->d_hull_points, W->h_hull_points.data(), W->h_hull_points.size());
	W->d_hull_shrunk = rt.alloc<F4>(W->h_hull_shrunk.size()); rt.upload(W->d_hull_shrunk, W->h_hull_shrunk.data(), W->h_hull_shrunk.size());
	W->d_hull_planes = rt.alloc<F4>(W->h_hull_planes.size()); rt.upload(W->d_hull_planes, W->h_hull_planes.data(), W->h_hull_planes.size());
	W->d_hull_faces = rt.alloc<uint32_t>(W->h_hull_faces.size()); rt.upload(W->d_hull_faces, W->h_hull_faces.data(), W->h_hull_faces.size());
	W->d_hull_vtx = rt.alloc<uint8_t>(W->h_hull_vtx.size() + 16); rt.upload(W->d_hull_vtx, W->h_hull_vtx.data(), W->h_hull_vtx.size());
	W->d_mesh_bytes = rt.alloc<uint8_t>(W->h_mesh_bytes.size() + 16); rt.upload(W->d_mesh_bytes, W->h_mesh_bytes.data(), W->h_mesh_bytes.size());
	if (!W->h_mesh_bytes.empty() && W->d_mesh_scratch == nullptr) W->d_mesh_scratch = rt.alloc<MeshScratch>(W->nc.num_scratch, false);
	W->d.shapes = W->d_shapes; W->d.hull_points = W->d_hull_points; W->d.hull_shrunk = W->d_hull_shrunk; W->d.hull_planes = W->d_hull_planes;
	W->d.hull_faces = W->d_hull_faces; W->d.hull_vtx = W->d_hull_vtx; W->d.mesh_bytes = W->d_mesh_bytes; W->d.compound_subs = W->d_compound_subs;
	W->shapes_dirty = false;
}

// (re)build the tree of one broadphase layer from the current cached body bounds
bool build_tree(b2j_world *W, uint32_t layer)
{
	Runtime &rt = W->rt;
	Tree &t = W->trees[layer];
	std::vector<uint32_t> &list = W->layer_bodies[layer];
	uint32_t n = (uint32_t)list.size();
	if (n > W->tree_capacity[layer])
	{
		uint32_t cap = next_pow2(n < 16? 16 : n);
		rt.sync();
		rt.free_(t.bodies); rt.free_(t.keys_in); rt.free_(t.keys_out); rt.free_(t.leaf_body); rt.free_(t.child_left); rt.free_(t.child_right);
		rt.free_(t.parent); rt.free_(t.node_min); rt.free_(t.node_max); rt.free_(t.visit);
		t.bodies = rt.alloc<uint32_t>(cap); t.keys_in = rt.alloc<uint64_t>(cap); t.keys_out = rt.alloc<uint64_t>(cap); t.leaf_body = rt.alloc<uint32_t>(cap);
		t.child_left = rt.alloc<int32_t>(cap); t.child_right = rt.alloc<int32_t>(cap); t.parent = rt.alloc<int32_t>(2 * cap);
		t.node_min = rt.alloc<F4>(2 * cap); t.node_max = rt.alloc<F4>(2 * cap); t.visit = rt.alloc<uint32_t>(cap);
		if (!t.visit) return false;
		W->tree_capacity[layer] = cap;
		W->layer_list_dirty[layer] = 1;
	}
	if (W->layer_list_dirty[layer])
	{
		rt.upload(t.bodies, list.data(), n);
		W->layer_list_dirty[layer] = 0;
	}
	t.n = n;
	if (n == 0) return true;
	KMorton km; km.w = W->d; km.t = t;
	rt.launch(km, n);
	if (t.world_root != nullptr) rt.memset_(t.world_root, 0xff, (size_t)W->num_worlds * 4);
	rt.sort_pairs<uint64_t>(t.keys_in, t.keys_out, t.bodies, t.leaf_body, n, W->num_worlds > 1? 64 : 30);
	if (n > 1)
	{
		KBuildHierarchy kb; kb.t = t;
		rt.launch(kb, n - 1);
	}
	KRefit kr; kr.w = W->d; kr.t = t;
	rt.launch(kr, n);
	return rt.check("build_tree");
}

void sync_dworld(b2j_world *W)
{
	W->d.active = W->active_buf[W->active_cur];
	W->d.write_cache = W->cache[W->write_idx];
	W->d.read_cache = W->cache[W->write_idx ^ 1];
}

void clear_cache(b2j_world *W, int idx)
{
	Runtime &rt = W->rt;
	rt.memset_(W->cache[idx].pair_table, 0xff, (size_t)W->d.pair_table_size * 4);
	rt.memset_(W->cache[idx].num_pairs, 0, 4);
	rt.memset_(W->cache[idx].num_manifolds, 0, 4);
	W->cache_num_pairs[idx] = 0;
	W->cache_num_manifolds[idx] = 0;
}

bool read_counters(b2j_world *W)
{
	W->rt.download(&W->h_counters, W->d.counters, 1);
	return W->rt.check("read_counters");
}

// The per phase velocity solve fetches the contact point parts where they are used, under a register budget (KSolveVelocityLate: 110
// registers, 16 resident warps per SM instead of 12 at 165). Measured with the driver's bench command: 4096 worlds 65.35 -> 62.94 ms
// per step, 512 worlds 9.57 -> 9.15 ms; 96 registers (B2J_SOLVE_LATE=2, spills) 62.82 / 9.40 ms. B2J_SOLVE_LATE=0: everything up front.
static int solve_late_mode()
{
	static const int mode = getenv("B2J_SOLVE_LATE") != nullptr? atoi(getenv("B2J_SOLVE_LATE")) : 1;
	return mode;
}

// B2J_SOLVE_PDL=0 turns programmatic dependent launch of the per phase solver kernels off (A/B measurements)
static bool solve_pdl_enabled()
{
#ifndef B2J_HOSTSIM
	static const bool on = getenv("B2J_SOLVE_PDL") == nullptr || atoi(getenv("B2J_SOLVE_PDL")) != 0;
	return on;
#else
	return false;
#endif
}

// Stages of a step as NVTX ranges (the reference marks the same stages with JPH_PROFILE scopes, PhysicsSystem.cpp), and with
// B2J_TRACE_STEP=1 their wall clock with the stream drained after every stage (diagnostics for small worlds, where the step is a chain
// of ~60 dependent launches: which stage pays how much launch / round trip latency). The trace prints to stderr.
struct StepTrace
{
	bool on; Runtime &rt; std::chrono::high_resolution_clock::time_point t0; std::string line;
	static void push(const char *name)
	{
#ifndef B2J_HOSTSIM
		nvtxRangePushA(name);
#else
		(void)name;
#endif
	}
	static void pop()
	{
#ifndef B2J_HOSTSIM
		nvtxRangePop();
#endif
	}
	StepTrace(Runtime &r, const char *first_stage) : on(getenv("B2J_TRACE_STEP") != nullptr), rt(r)
	{
		push("b2j collision step");
		push(first_stage);
		if (on) { rt.sync(); t0 = std::chrono::high_resolution_clock::now(); }
	}
	// the stage `finished` is over, `next` begins
	void mark(const char *finished, const char *next)
	{
		pop();
		push(next);
		if (!on) return;
		rt.sync();
		auto t1 = std::chrono::high_resolution_clock::now();
		char buf[64];
		snprintf(buf, sizeof(buf), " %s %.0f", finished, std::chrono::duration<double, std::micro>(t1 - t0).count());
		line += buf;
		t0 = t1;
	}
	~StepTrace()
	{
		pop(); pop();
		if (on) fprintf(stderr, "[b2j step us]%s launches %u\n", line.c_str(), rt.launches);
	}
};

// ---- non contact constraints: device storage of the world's list (b2j_joints.h)
bool joints_reserve(b2j_world *W, uint32_t needed)
{
	if (needed <= W->joint_capacity) return true;
	Runtime &rt = W->rt;
	uint32_t cap = W->joint_capacity == 0? 64 : W->joint_capacity;
	while (cap < needed) cap *= 2;
	JointCtx &j = W->jc;
	JointDef *defs = rt.alloc<JointDef>(cap); JointState *state = rt.alloc<JointState>(cap);
	uint32_t *steps = rt.alloc<uint32_t>(cap);
	if (defs == nullptr || state == nullptr || steps == nullptr) { last_error() = "out of device memory"; return false; }
	if (W->joint_capacity != 0)
	{
		rt.copy(state, j.state, W->joint_capacity);
		rt.sync();
		rt.free_(j.defs); rt.free_(j.state); rt.free_(j.active_flag); rt.free_(j.order_flag); rt.free_(j.order_scan); rt.free_(j.active_joints);
		uint32_t *old_order = const_cast<uint32_t *>(j.order); rt.free_(old_order);
		uint32_t *old_steps = const_cast<uint32_t *>(W->sc.joint_steps); rt.free_(old_steps);
	}
	j.defs = defs; j.state = state; W->sc.joint_steps = steps;
	j.active_flag = rt.alloc<uint32_t>(cap); j.order = rt.alloc<uint32_t>(cap); j.order_flag = rt.alloc<uint32_t>(cap + 1); j.order_scan = rt.alloc<uint32_t>(cap + 1);
	j.active_joints = rt.alloc<uint32_t>(cap);
	if (j.wake_key == nullptr)
	{
		j.wake_key = rt.alloc<uint32_t>(W->d.max_bodies, false);
		rt.memset_(j.wake_key, 0xff, (size_t)W->d.max_bodies * 4);
		W->sc.body_nj = rt.alloc<uint32_t>(W->d.max_bodies);
	}
	W->joint_capacity = cap;
	return j.active_joints != nullptr && j.wake_key != nullptr && W->sc.body_nj != nullptr;
}

// definitions, (priority, index) order and step overrides of the host list -> device (state stays where it is)
void joints_upload(b2j_world *W)
{
	if (!W->joints_dirty) return;
	W->joints_dirty = false;
	// (a batch group holds the list of ONE world: world w's copy of constraint i is entry w * n + i, its bodies in the slots of world w)
	uint32_t n = (uint32_t)W->h_joints.size(), nw = W->num_worlds, stride = W->d.world_stride;
	W->jc.num_joints = n * nw;
	if (n == 0) return;
	Runtime &rt = W->rt;
	std::vector<JointDef> defs((size_t)n * nw);
	std::vector<uint32_t> order((size_t)n * nw), steps((size_t)n * nw), sorted(n);
	for (uint32_t i = 0; i < n; ++i) sorted[i] = i;
	// ConstraintManager::sSortConstraints: priority, then constraint index
	std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) { return W->h_joints[a].priority < W->h_joints[b].priority; });
	for (uint32_t wi = 0; wi < nw; ++wi)
		for (uint32_t i = 0; i < n; ++i)
		{
			const b2j_constraint_desc &c = W->h_joints[i];
			JointDef &d = defs[(size_t)wi * n + i];
			memset(&d, 0, sizeof(d));
			d.type = c.type; d.b1 = slot_of(c.body1) + wi * stride; d.b2 = slot_of(c.body2) + wi * stride; d.flags = c.enabled? JOINT_ENABLED : 0u;
			d.priority = c.priority; d.steps_override = (uint32_t)c.num_velocity_steps_override | ((uint32_t)c.num_position_steps_override << 8);
			d.index = i;
			d.local1 = f4(v3_load(c.point1), c.min_distance); d.local2 = f4(v3_load(c.point2), c.max_distance);
			d.axis1 = f4(v3_load(c.hinge_axis1), c.limits_min); d.axis2 = f4(v3_load(c.hinge_axis2), c.limits_max);
			d.inv_initial_orientation = f4(c.inv_initial_orientation[0], c.inv_initial_orientation[1], c.inv_initial_orientation[2], c.inv_initial_orientation[3]);
			d.hinge = f4(c.max_friction_torque, 0.0f, 0.0f, 0.0f);
			order[(size_t)wi * n + i] = wi * n + sorted[i]; steps[(size_t)wi * n + i] = d.steps_override | (c.type << 16);
		}
	rt.upload(W->jc.defs, defs.data(), defs.size());
	rt.upload(const_cast<uint32_t *>(W->jc.order), order.data(), order.size());
	rt.upload(const_cast<uint32_t *>(W->sc.joint_steps), steps.data(), steps.size());
	rt.sync();
}

// One collision step. Returns false on a CUDA failure.
bool collision_step(b2j_world *W, float dt, float warm_start_ratio, bool is_last, b2j_step_stats *stats)
{
	Runtime &rt = W->rt;
	DWorld &d = W->d;
	sync_dworld(W);
	rt.memset_(d.counters, 0, sizeof(StepCounters));
	rt.memset_(W->d_round_begin, 0, 4);
	if (W->last_num_events != 0 || W->last_num_act_events != 0)
	{
		// events accumulate over the collision steps of one b2j_step
		rt.upload(&d.counters->num_events, &W->last_num_events, 1);
		rt.upload(&d.counters->num_activation_events, &W->last_num_act_events, 1);
	}

	StepTrace trace(rt, "gravity + broadphase trees");
	// (a2) gravity, forces, damping
	const uint32_t gravity_count = W->num_active;
	{ KApplyGravity k; k.w = d; k.dt = dt; rt.launch(k, W->num_active); }

	// (a4) broadphase maintenance: rebuild the trees whose bodies moved / changed
	for (uint32_t l = 0; l < d.num_bp_layers; ++l)
		if (W->layer_needs_build[l])
		{
			if (!build_tree(W, l)) return false;
			W->layer_needs_build[l] = 0;
		}

	// non contact constraints (b2j_joints.h): which take part in this step, the bodies they wake up (no gravity for those this step)
	uint32_t J = 0;
	const bool have_joints = !W->h_joints.empty();
	if (have_joints)
	{
		joints_upload(W);
		JointCtx &jc = W->jc;
		{ KJointActive k; k.w = d; k.c = W->nc; k.j = jc; rt.launch(k, jc.num_joints); }
		if (!read_counters(W)) return false;
		J = W->h_counters.num_active_joints;
		uint32_t woken = W->h_counters.num_woken;
		if (woken > 0)
		{
			// BodyManager::ActivateBodies in the order ConstraintManager::sBuildIslands calls it: by constraint, body 1 before body 2
			{ KJointWakeKeys k; k.c = W->nc; k.j = jc; k.keys = W->d_woken_keys; rt.launch(k, woken); }
			uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(W->d_sort_keys[0]);
			rt.sort_pairs<uint32_t>(W->d_woken_keys, keys_sorted, W->nc.woken_list, W->d_woken_sorted, woken, 32);
			{ KActivateWoken k; k.w = d; k.woken_sorted = W->d_woken_sorted; k.base = W->num_active; k.woken_flag = W->nc.woken_flag; k.events = W->d_act_events; k.max_events = W->max_act_events; rt.launch(k, woken); }
			{ KJointClearWoken k; k.w = d; rt.launch(k, 1); }
			W->num_active += woken;
		}
		if (J > 0)
		{
			{ KJointOrderFlags k; k.j = jc; rt.launch(k, jc.num_joints); }
			rt.exclusive_scan(jc.order_flag, jc.order_scan, jc.num_joints);
			{ KJointCompact k; k.j = jc; rt.launch(k, jc.num_joints); }
			{ KJointSetup k; k.w = d; k.j = jc; rt.launch(k, J); }
		}
	}
	const uint32_t woken_by_joints = W->num_active - gravity_count;

	trace.mark("gravity+trees", "find pairs + narrowphase");
	// (a3, a5..a9) find pairs + narrow phase; repeated for the bodies woken up by contacts until no new body wakes up
	uint32_t first_active = 0, n_query = W->num_active;
	uint32_t woken_total = woken_by_joints, longest_queue = 0;
	const int max_rounds = 64;
	for (int round = 0; ; ++round)
	{
		if (round == max_rounds) { last_error() = "bodies were still waking each other up after 64 broadphase rounds in one step"; return false; }
		{
			KFindPairs k; k.w = d; for (int l = 0; l < 8; ++l) k.trees[l] = W->trees[l]; k.pairs = W->nc.pairs; k.first = first_active; k.query_leaves = nullptr;
			if (round == 0)
			{
				// all active bodies: one launch per layer that holds moving bodies, in the leaf order of that layer's tree
				for (uint32_t l = 0; l < d.num_bp_layers; ++l)
					if (W->layer_has_moving[l] && W->trees[l].n > 0)
					{
						k.query_leaves = W->trees[l].leaf_body;
						rt.launch(k, W->trees[l].n);
					}
			}
			else
				rt.launch(k, n_query); // bodies woken by the previous round, active list order
		}
		{ KProcessPairs k; k.w = d; k.c = W->nc; k.first_ptr = W->d_round_begin; rt.launch_dev(k, &d.counters->num_pairs, W->d_round_begin, d.max_body_pairs, 16); } // 40 registers, a chain of ~6 dependent gathers: full occupancy
		{ KCopyCached k; k.w = d; k.c = W->nc; k.first_ptr = W->d_round_begin; rt.launch_dev(k, &d.counters->num_pairs, W->d_round_begin, d.max_body_pairs); }
		// convex pairs: GJK (thread per pair, lockstep) queues shallow hits as results and deep ones for EPA
		rt.memset_(W->nc.num_epa_overflow, 0, 4);
		rt.memset_(W->nc.num_epa_results, 0, 4);
		W->nc.collide_order = nullptr;
		W->nc.collide_order_n = 0;
		// (B2J_COLLIDE_ORDER_MIN: queue length from which a single world orders its queue; tests lower it to cover the path on small scenes)
		const char *order_env = getenv("B2J_COLLIDE_ORDER_MIN");
		const uint32_t order_min = order_env != nullptr? (uint32_t)atoi(order_env) : 65536u;
		if (d.world_stride == 0 && W->d_collide_keys[0] == nullptr && W->last_num_pairs >= 4 * order_min)
			for (int i = 0; i < 2; ++i) { W->d_collide_keys[i] = rt.alloc<uint32_t>(d.max_body_pairs, false); W->d_collide_vals[i] = rt.alloc<uint32_t>(d.max_body_pairs, false); }
		if (d.world_stride < 65536 && W->d_collide_keys[0] != nullptr)
		{
			// batch group: same pair of different worlds in neighbouring lanes (GJK / EPA / manifold code paths then coincide: 15 -> ~30
			// active lanes per instruction); one big world: pairs grouped by shape type pair. The queue length only exists on the device
			// and a radix sort is sized on the host: the sort covers a CAPACITY derived from the longest queue of the previous step
			// (entries past the real end carry the largest key and stay behind it; should the queue outgrow the capacity, the excess
			// is processed in queue order) -- no host round trip.
			uint32_t cap = 2 * W->last_collide_convex + 1024;
			if (cap > d.max_body_pairs) cap = d.max_body_pairs;
			if (W->last_collide_convex >= (d.world_stride != 0? 512u : order_min / 2))
			{
				uint32_t bits = 1;
				while ((1u << bits) < d.world_stride) ++bits;
				uint32_t key_bits = d.world_stride != 0? 2 * bits : 6;
				{ KCollideKeys k; k.w = d; k.c = W->nc; k.keys = W->d_collide_keys[0]; k.vals = W->d_collide_vals[0]; k.bits = bits; k.invalid_key = 1u << key_bits; rt.launch(k, cap); }
				rt.sort_pairs<uint32_t>(W->d_collide_keys[0], W->d_collide_keys[1], W->d_collide_vals[0], W->d_collide_vals[1], cap, (int)key_bits + 1);
				W->nc.collide_order = W->d_collide_vals[1];
				W->nc.collide_order_n = cap;
			}
		}
		{ KCollideConvex k; k.w = d; k.c = W->nc; rt.launch_dev_lockstep(k, &d.counters->num_collide_convex, nullptr, d.max_body_pairs); }
		// deep pairs: thread per pair, lanes in lockstep, EPA scratch in (lane interleaved, L1/L2 cached) local memory. Small tier first
		// (2 KB per lane, 16 warps per SM); the pairs that overflow it re-run on full size storage (21 KB per lane).
		{ KCollideEpaSmall k; k.w = d; k.c = W->nc; rt.launch_lane_local<KCollideEpaSmall, EpaStorageSmall>(k, &d.counters->num_epa, W->nc.max_epa, 4); }
		{ KCollideEpaFull k; k.w = d; k.c = W->nc; rt.launch_lane_local<KCollideEpaFull, EpaStorageFull>(k, W->nc.num_epa_overflow, W->nc.max_epa, 2); }
		{ KFinishPairs k; k.w = d; k.c = W->nc; rt.launch_dev(k, W->nc.num_epa_results, nullptr, W->nc.max_epa); }
		if (W->d_mesh_scratch != nullptr)
		{
			// all 32 lanes work on the triangles of one (convex, mesh) pair; the serial form is what tests/hostsim runs (and B2J_MESH_SERIAL=1)
			KCollideMesh k; k.w = d; k.c = W->nc; k.mesh_scratch = W->d_mesh_scratch;
#ifndef B2J_HOSTSIM
			static const bool serial = getenv("B2J_MESH_SERIAL") != nullptr && atoi(getenv("B2J_MESH_SERIAL")) != 0;
			if (serial)
				rt.launch_warp_smem<KCollideMesh, EpaStorageFull>(k, &d.counters->num_collide_mesh, d.max_body_pairs, W->nc.num_scratch);
			else
			{
				KCollideMeshWarp kw; kw.w = d; kw.c = W->nc; kw.mesh_scratch = W->d_mesh_scratch;
				rt.launch_warp_coop<KCollideMeshWarp, KCollideMesh, EpaStorageFull, MeshWarpShared>(kw, k, &d.counters->num_collide_mesh, d.max_body_pairs, W->nc.num_scratch);
			}
#else
			rt.launch_warp_smem<KCollideMesh, EpaStorageFull>(k, &d.counters->num_collide_mesh, d.max_body_pairs, W->nc.num_scratch);
#endif
		}
		if (!read_counters(W)) return false;
#ifndef B2J_HOSTSIM
		{
			// B2J_TRACE_EPA=1: how many pairs each EPA tier received this round (diagnostics, stderr)
			static const bool trace_epa = getenv("B2J_TRACE_EPA") != nullptr;
			if (trace_epa)
			{
				uint32_t n1 = 0, nr = 0, hist[130];
				rt.download(&n1, W->nc.num_epa_overflow, 1); rt.download(&nr, W->nc.num_epa_results, 1);
				rt.download(hist, W->nc.epa_hist, 130);
				rt.memset_(W->nc.epa_hist, 0, 130 * 4);
				uint32_t over[5] = { 0, 0, 0, 0, 0 };
				const uint32_t limits[5] = { 32, 48, 64, 96, 127 };
				for (uint32_t p = 0; p < 130; ++p) for (int j = 0; j < 5; ++j) if (p > limits[j]) over[j] += hist[p];
				fprintf(stderr, "[b2j epa] collide %u gjk->epa %u full tier %u (points > 32: %u, > 48: %u, > 64: %u, > 96: %u, 128: %u) results %u\n", W->h_counters.num_collide_convex, W->h_counters.num_epa, n1, over[0], over[1], over[2], over[3], over[4], nr);
			}
		}
#endif
		uint32_t woken = W->h_counters.num_woken;
		if (W->h_counters.num_collide_convex > longest_queue) longest_queue = W->h_counters.num_collide_convex;
		if (W->h_counters.num_epa > W->nc.max_epa) { last_error() = "EPA queue overflow"; return false; }
		{ KNextRound k; k.w = d; k.round_begin = W->d_round_begin; rt.launch(k, 1); }
		if (woken == 0)
			break;
		// activate the woken bodies in slot order (BodyManager::ActivateBodies), then query only them
		rt.sort_pairs<uint32_t>(W->nc.woken_list, W->d_woken_keys, W->nc.woken_list, W->d_woken_sorted, woken, 32);
		{ KActivateWoken k; k.w = d; k.woken_sorted = W->d_woken_sorted; k.base = W->num_active; k.woken_flag = W->nc.woken_flag; k.events = W->d_act_events; k.max_events = W->max_act_events; rt.launch(k, woken); }
		first_active = W->num_active;
		n_query = woken;
		W->num_active += woken;
		woken_total += woken;
	}
	// M contact constraints; the active non contact constraints follow them as constraint sources (items = J + M)
	const uint32_t num_contacts = W->h_counters.num_constraints < d.max_constraints? W->h_counters.num_constraints : d.max_constraints;
	if (J > d.max_constraints - num_contacts)
	{
		// (the constraint storage is shared: max_contact_constraints has to cover the active non contact constraints too)
		J = d.max_constraints - num_contacts;
		W->h_counters.error_bits |= B2J_ERR_CONTACT_CONSTRAINTS_FULL;
		if (stats != nullptr) stats->error_bits |= B2J_ERR_CONTACT_CONSTRAINTS_FULL;
	}
	if (J > 0) { KJointAppend k; k.w = d; k.j = W->jc; k.src = W->nc.con_src; k.first = num_contacts; k.count = J; rt.launch(k, J); }
	const uint32_t M = num_contacts + J;
	uint32_t num_pairs = W->h_counters.num_pairs < d.max_body_pairs? W->h_counters.num_pairs : d.max_body_pairs;
	W->last_num_pairs = num_pairs;
	// (decaying maximum: the queue of a scene at impact alternates between long and short from step to step, and a queue that outgrows the
	// capacity derived from this number is only partly ordered)
	if (longest_queue < W->last_collide_convex / 2) longest_queue = W->last_collide_convex / 2;
	W->last_collide_convex = longest_queue < d.max_body_pairs? longest_queue : d.max_body_pairs;
	uint32_t na = W->num_active;

	trace.mark("pairs+narrowphase", "islands + sort + adjacency");
	// (a12) islands
	SolveCtx &sc = W->sc;
	sc.num_slots = W->num_slots;
	{ KUfInit k; k.s = sc; rt.launch(k, W->num_slots); }
	sc.src = W->nc.con_src;
	{ KUfUnion k; k.w = d; k.s = sc; rt.launch(k, M); }
	{ KUfFlatten k; k.w = d; k.s = sc; rt.launch(k, na); }
	{ KIslandCount k; k.w = d; k.s = sc; rt.launch(k, M); }
	{ KIslandClassify k; k.w = d; k.s = sc; rt.launch(k, na); }

	uint32_t num_phases = 0, vsteps = 0, psteps = 0;
	bool block_solve = false, solved_by_phase_launches = false, diag_skipped = false;
	uint32_t joints_grid_position = 0; // grid of the cooperative constraint solvers (0: per phase launches)
	if (M > 0)
	{
		// (a14 SortContacts) order by sort key
		{ KGatherSortKeys k; k.s = sc; k.keys = W->d_sort_keys[0]; k.vals = W->d_sort_vals; rt.launch(k, num_contacts); }
		rt.sort_pairs<uint64_t>(W->d_sort_keys[0], W->d_sort_keys[1], W->d_sort_vals, sc.order + J, num_contacts);
		{ KCountTies k; k.w = d; k.keys = W->d_sort_keys[1]; rt.launch(k, num_contacts); }
		// the non contact constraints lead the order (an island solves its constraints before its contacts)
		if (J > 0) { KJointItemOrder k; k.s = sc; k.vals = W->d_sort_vals; k.num_contacts = num_contacts; k.num_joints = J; rt.launch(k, M); }

		// body -> constraints adjacency in sorted order
		rt.exclusive_scan(sc.body_deg, sc.body_off, W->num_slots);
		{ KAdjFill k; k.w = d; k.s = sc; rt.launch(k, M); }
		{ KAdjSort k; k.w = d; k.s = sc; rt.launch(k, na); }

		trace.mark("islands+sort+adjacency", "solve schedule");
		// (a13) schedule: wavefronts in sorted order; pass 0 = levels (small islands) / colours (large islands)
		const uint32_t rounds_per_check = 8;
#ifndef B2J_HOSTSIM
		// batch group: one block per world; small single world: one block. All rounds in one launch (see sched_block_kernel)
		const bool block_sched = d.world_stride != 0 || W->num_slots <= 4096;
		if (block_sched)
		{
			uint32_t blocks = d.world_stride != 0? W->num_worlds : 1, slots_per_block = d.world_stride != 0? d.world_stride : W->num_slots;
			++rt.launches;
			if (rt.profiling) rt.prof_begin(profile_category<KSchedBlock>());
			// as many threads per world as fit with every world of the group resident (a round costs one dependent chain of ~8 L2
			// round trips per body a thread owns: fewer bodies per thread = shorter rounds)
			uint32_t threads = 1024;
			while (threads > 128 && (uint64_t)blocks * threads > (uint64_t)rt.num_sms * 2048) threads /= 2;
			sched_block_kernel<<<blocks, threads, 0, rt.stream>>>(d, sc, slots_per_block);
			if (rt.profiling) rt.prof_end();
		}
#else
		const bool block_sched = false;
#endif
#ifndef B2J_HOSTSIM
		// big single world: all rounds in one cooperative launch (grid wide barriers)
		bool grid_sched = false;
		if (!block_sched)
		{
			int &blocks_per_sm = rt.func_blocks_per_sm[(const void *)sched_grid_kernel];
			if (blocks_per_sm == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sched_grid_kernel, 256, 0);
			if (blocks_per_sm > 0)
			{
				uint32_t arg_na = na, arg_m = M;
				void *args[] = { (void *)&d, (void *)&sc, (void *)&arg_na, (void *)&arg_m };
				++rt.launches;
				if (rt.profiling) rt.prof_begin(profile_category<KSchedGrid>());
				cudaError_t e = cudaLaunchCooperativeKernel((const void *)sched_grid_kernel, dim3((unsigned)(rt.num_sms * blocks_per_sm)), dim3(256), args, 0, rt.stream);
				if (rt.profiling) rt.prof_end();
				grid_sched = e == cudaSuccess;
				if (!grid_sched) cudaGetLastError(); // fall back to the per round launches
			}
		}
#else
		const bool grid_sched = false;
#endif
		for (uint32_t pass = 0; pass < 2 && !block_sched && !grid_sched; ++pass)
		{
			if (pass == 1)
			{
				if (W->h_counters.num_large_islands == 0)
				{
					read_counters(W);
					if (W->h_counters.num_large_islands == 0) break;
				}
				{ KSchedSerialize k; k.w = d; k.s = sc; rt.launch(k, M); }
				{ KSchedResetCursors k; k.w = d; k.s = sc; rt.launch(k, na); }
			}
			rt.memset_(sc.sched_flag, 0, 4096 * 4);
			{ KSchedAdvance k; k.w = d; k.s = sc; k.pass = pass; k.flag_index = 0; rt.launch(k, na); }
			uint32_t round = 0;
			for (;;)
			{
				uint32_t flag_index = 0;
				for (uint32_t r = 0; r < rounds_per_check; ++r, ++round)
				{
					flag_index = 1 + (round % 4095);
					{ KSchedDecide k; k.w = d; k.s = sc; k.round = round; k.pass = pass; rt.launch(k, na); }
					{ KSchedAdvance k; k.w = d; k.s = sc; k.pass = pass; k.flag_index = flag_index; rt.launch(k, na); }
				}
				uint32_t remaining = 0;
				rt.download(&remaining, sc.sched_flag + flag_index, 1);
				if (remaining == 0) break;
				if (round >= 4000) rt.memset_(sc.sched_flag, 0, 4096 * 4);
				if (round > 1000000) { last_error() = "schedule did not converge"; return false; }
			}
		}

		trace.mark("schedule", "placement + constraint setup");
		// B2J_SOLVE_MODE: how the velocity / position solve is launched.
		//   0 = one launch per phase per iteration (default: the groups of a batch overlap on their streams)
		//   1 = one persistent cooperative launch, phases separated by grid barriers, loads straight from HBM
		//   2 = the same with the constraint planes streamed through shared memory by TMA (solve_velocity_tma_kernel)
#ifndef B2J_HOSTSIM
		const char *solve_mode_env = getenv("B2J_SOLVE_MODE");
		int solve_mode = solve_mode_env != nullptr? atoi(solve_mode_env) : 0;
#endif
		// phases -> solve order
		{
			// the 64 bit key buffers of the constraint sort are free again: reuse them for the (phase, index) sort; d_sort_vals still holds 0..M-1
			uint32_t *sorted_phase = reinterpret_cast<uint32_t *>(W->d_sort_keys[0]), *sorted_idx = reinterpret_cast<uint32_t *>(W->d_sort_keys[1]);
			{ KPhaseClamp k; k.w = d; k.s = sc; rt.launch(k, M); }
			rt.sort_pairs<uint32_t>(sc.phase, sorted_phase, W->d_sort_vals, sorted_idx, M, 13);
			{ KPhasePlace k; k.s = sc; k.sorted_phase = sorted_phase; k.sorted_idx = sorted_idx; k.n = M; rt.launch(k, M); }
		}

		// (a11) constraint setup straight into solve order
		{ KSetupConstraints k; k.w = d; k.s = sc; k.dt = dt; rt.launch(k, M); }

		trace.mark("place+setup", "velocity solve");
#ifndef B2J_HOSTSIM
		// small single world: the whole velocity solve in one small cooperative launch (solve_small_kernel)
		block_solve = d.world_stride == 0 && W->num_slots <= 4096 && M <= 16384 && J == 0;
		// worlds with non contact constraints: one cooperative launch for the velocity solve, one for the position solve (b2j_joints.h);
		// B2J_JOINTS_COOP=0 keeps the per phase launches (A/B, and what the host simulation runs)
		uint32_t joints_grid = 0;
		if (J > 0)
		{
			solve_mode = 0; // (the other one launch solvers know contacts only)
			static const bool coop = getenv("B2J_JOINTS_COOP") == nullptr || atoi(getenv("B2J_JOINTS_COOP")) != 0;
			int &bv = rt.func_blocks_per_sm[(const void *)solve_velocity_joints_kernel], &bp = rt.func_blocks_per_sm[(const void *)solve_position_joints_kernel];
			if (bv == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bv, solve_velocity_joints_kernel, 128, 0) != cudaSuccess || bv < 1)) { cudaGetLastError(); bv = -1; }
			if (bp == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp, solve_position_joints_kernel, 128, 0) != cudaSuccess || bp < 1)) { cudaGetLastError(); bp = -1; }
			// (small single worlds only. Batches of the joints scene measured no better with it once a phase's constraints and contacts
			// share one launch: 256 worlds in one group 2.78 ms per step against 2.83, 2048 worlds in 8 groups 6.3 against 5.6, 8192 worlds
			// 14.6 against 11.2 -- the barrier'd passes of concurrent groups compete for the SMs, wide phases prefer their own launches)
			const bool small_world = d.world_stride == 0 && W->num_slots <= 4096 && M <= 16384;
			static const bool coop_always = getenv("B2J_JOINTS_COOP") != nullptr && atoi(getenv("B2J_JOINTS_COOP")) == 2;
			if (coop && bv > 0 && bp > 0 && (small_world || coop_always))
			{
				// a small single world: a small grid (the barrier is what a phase costs); else every SM this group may use
				const bool small = d.world_stride == 0 && W->num_slots <= 4096 && M <= 16384;
				uint32_t per_sm = (uint32_t)(bv < bp? bv : bp);
				// (the groups of a batch solve concurrently on their own share of the SMs: thin phases gain nothing from a wider grid, and
				// cooperative launches that each want every SM would take turns)
				uint32_t div = W->solve_grid_div > W->batch_groups? W->solve_grid_div : W->batch_groups;
				joints_grid = small? 16u : (uint32_t)rt.num_sms * per_sm / div;
				if (joints_grid < 1) joints_grid = 1;
				float ratio_arg = warm_start_ratio, dt_arg = dt;
				void *args[] = { (void *)&d, (void *)&sc, (void *)&W->jc, (void *)&ratio_arg, (void *)&dt_arg };
				++rt.launches;
				if (rt.profiling) rt.prof_begin(profile_category<KSolveVelocityJoints>());
				cudaError_t e = cudaLaunchCooperativeKernel((const void *)solve_velocity_joints_kernel, dim3(joints_grid), dim3(128), args, 0, rt.stream);
				if (rt.profiling) rt.prof_end();
				if (e != cudaSuccess) { cudaGetLastError(); joints_grid = 0; } // fall back to the per phase launches
			}
		}
		if (block_solve)
		{
			float ratio_arg = warm_start_ratio;
			void *args[] = { (void *)&d, (void *)&sc, (void *)&ratio_arg };
			++rt.launches;
			if (rt.profiling) rt.prof_begin(profile_category<KSolveSmallVelocity>());
			cudaError_t e = cudaLaunchCooperativeKernel((const void *)solve_small_kernel<false>, dim3(8), dim3(256), args, 0, rt.stream);
			if (rt.profiling) rt.prof_end();
			if (e != cudaSuccess) { cudaGetLastError(); block_solve = false; }
		}
		if (!block_solve && solve_mode != 0)
		{
			float ratio_arg = warm_start_ratio;
			void *args[] = { (void *)&d, (void *)&sc, (void *)&ratio_arg };
			cudaError_t e = cudaErrorUnknown;
			if (solve_mode == 2)
			{
				// warps per block of the TMA pipeline (B2J_SOLVE_TMA_WARPS: 12 (default), 10 or 8)
				const char *warps_env = getenv("B2J_SOLVE_TMA_WARPS");
				const int wsel = warps_env != nullptr? atoi(warps_env) : 12;
				const void *fn = wsel == 8? (const void *)solve_velocity_tma_kernel<8> : (wsel == 10? (const void *)solve_velocity_tma_kernel<10> : (const void *)solve_velocity_tma_kernel<12>);
				const int threads = (wsel == 8? 8 : (wsel == 10? 10 : 12)) * 32;
				const size_t smem = sv_smem_bytes(threads / 32);
				int &blocks_per_sm = rt.func_blocks_per_sm[fn];
				if (blocks_per_sm == 0)
				{
					cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
					if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, threads, smem) != cudaSuccess || blocks_per_sm < 1) { cudaGetLastError(); blocks_per_sm = -1; }
				}
				if (blocks_per_sm > 0)
				{
					++rt.launches;
					rt.memset_(sc.grid_barrier, 0, 4);
					if (rt.profiling) rt.prof_begin(profile_category<KSolveVelocityAll>());
					e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)(rt.num_sms * blocks_per_sm / W->solve_grid_div)), dim3(threads), args, smem, rt.stream);
					if (rt.profiling) rt.prof_end();
				}
			}
			else
			{
				int &blocks_per_sm = rt.func_blocks_per_sm[(const void *)solve_velocity_all_kernel];
				if (blocks_per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, solve_velocity_all_kernel, 128, 0) != cudaSuccess || blocks_per_sm < 1)) { cudaGetLastError(); blocks_per_sm = -1; }
				if (blocks_per_sm > 0)
				{
					++rt.launches;
					if (rt.profiling) rt.prof_begin(profile_category<KSolveVelocityAllPlain>());
					e = cudaLaunchCooperativeKernel((const void *)solve_velocity_all_kernel, dim3((unsigned)(rt.num_sms * blocks_per_sm / W->solve_grid_div)), dim3(128), args, 0, rt.stream);
					if (rt.profiling) rt.prof_end();
				}
			}
			if (e != cudaSuccess) { cudaGetLastError(); solve_mode = 0; } // fall back to the per phase launches
		}
		bool phase_launches = !block_solve && solve_mode == 0 && joints_grid == 0;
		joints_grid_position = joints_grid;
		// (diagnostics only, results are WRONG: B2J_DIAG_SKIP_SOLVE=1 skips the velocity / position solve to time the rest of the step)
		const bool diag_skip_solve = getenv("B2J_DIAG_SKIP_SOLVE") != nullptr;
		if (diag_skip_solve) phase_launches = false;
#else
		const bool phase_launches = true;
#endif
		if (phase_launches)
		{
			// per phase launches are sized on the host: phase count, iteration counts and phase offsets come back first
			if (!read_counters(W)) return false;
			num_phases = W->h_counters.num_phases;
			vsteps = W->h_counters.max_velocity_steps;
			W->h_phase_offsets.resize(num_phases + 1);
			rt.download(W->h_phase_offsets.data(), sc.phase_count, num_phases + 1);
			// (a14) warm start + velocity iterations, one launch per phase
			for (uint32_t p = 0; p < num_phases; ++p)
			{
				uint32_t begin = W->h_phase_offsets[p], n = W->h_phase_offsets[p + 1] - begin;
				KWarmStart k; k.w = d; k.c = sc.con; k.begin = begin; k.ratio = warm_start_ratio;
				if (J > 0) { KMixedWarmStart km; km.contacts = k; km.joints.w = d; km.joints.c = sc.con; km.joints.j = W->jc; km.joints.begin = begin; km.joints.ratio = warm_start_ratio; rt.launch(km, n); }
				else if (solve_pdl_enabled()) { k.pdl = 1; rt.launch_pdl(k, n); } else rt.launch(k, n);
			}
			for (uint32_t it = 0; it < vsteps; ++it)
				for (uint32_t p = 0; p < num_phases; ++p)
				{
					uint32_t begin = W->h_phase_offsets[p], n = W->h_phase_offsets[p + 1] - begin;
					KSolveVelocity k; k.w = d; k.c = sc.con; k.begin = begin; k.iteration = it;
					k.prefetch = 1;
					if (J > 0) { KMixedSolveVelocity km; km.contacts = k; km.joints.w = d; km.joints.c = sc.con; km.joints.j = W->jc; km.joints.begin = begin; km.joints.iteration = it; km.joints.dt = dt; rt.launch(km, n); }
					else if (solve_late_mode() != 0 && solve_pdl_enabled())
					{
						// contact point parts fetched late under a register budget (more resident warps), see solve_late_mode
						KSolveVelocityLate kl; kl.w = d; kl.c = sc.con; kl.begin = begin; kl.iteration = it; kl.prefetch = 1; kl.pdl = 1;
						if (solve_late_mode() == 1) rt.launch_pdl_cfg<KSolveVelocityLate, 128, 4>(kl, n);
						else rt.launch_pdl_cfg<KSolveVelocityLate, 128, 5>(kl, n);
					}
					else if (solve_pdl_enabled()) { k.pdl = 1; rt.launch_pdl(k, n); } else rt.launch(k, n);
				}
		}
		solved_by_phase_launches = phase_launches;
#ifndef B2J_HOSTSIM
		if (diag_skip_solve) diag_skipped = true;
#endif
		// the applied impulses are stored by the last velocity iteration of every constraint; islands without iterations only exist when
		// the default number of velocity steps is 0
		if (d.settings.num_velocity_steps == 0) { KStoreImpulses k; k.w = d; k.c = sc.con; rt.launch(k, M); }
	}

	trace.mark("velocity", "integrate + position solve");
	// (a15) integrate
	{ KIntegrate k; k.w = d; k.dt = dt; rt.launch(k, na); }

	// (a10) contact removed events for manifolds that were not persisted
	if (W->nc.events != nullptr)
	{
		uint32_t old_m = W->cache_num_manifolds[W->write_idx ^ 1];
		KRemovedEvents k; k.w = d; k.c = W->nc; rt.launch(k, old_m);
	}

	// (a16) position iterations
#ifndef B2J_HOSTSIM
	if (block_solve && M > 0)
	{
		float ratio_arg = 0.0f;
		void *args[] = { (void *)&d, (void *)&sc, (void *)&ratio_arg };
		++rt.launches;
		if (rt.profiling) rt.prof_begin(profile_category<KSolveSmallPosition>());
		cudaError_t e = cudaLaunchCooperativeKernel((const void *)solve_small_kernel<true>, dim3(8), dim3(256), args, 0, rt.stream);
		if (rt.profiling) rt.prof_end();
		if (e != cudaSuccess) { cudaGetLastError(); last_error() = "cooperative position solve launch failed"; return false; }
	}
	else if (M > 0 && J > 0 && !solved_by_phase_launches && !diag_skipped)
	{
		void *args[] = { (void *)&d, (void *)&sc, (void *)&W->jc };
		++rt.launches;
		if (rt.profiling) rt.prof_begin(profile_category<KSolvePositionJoints>());
		cudaError_t e = cudaLaunchCooperativeKernel((const void *)solve_position_joints_kernel, dim3(joints_grid_position), dim3(128), args, 0, rt.stream);
		if (rt.profiling) rt.prof_end();
		if (e != cudaSuccess) { cudaGetLastError(); solved_by_phase_launches = true; } // fall back to the per phase launches below
	}
	else if (M > 0 && !solved_by_phase_launches && !diag_skipped)
	{
		void *args[] = { (void *)&d, (void *)&sc };
		int &blocks_per_sm = rt.func_blocks_per_sm[(const void *)solve_position_all_kernel];
		if (blocks_per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, solve_position_all_kernel, 128, 0) != cudaSuccess || blocks_per_sm < 1)) { cudaGetLastError(); blocks_per_sm = -1; }
		cudaError_t e = cudaErrorUnknown;
		if (blocks_per_sm > 0)
		{
			++rt.launches;
			if (rt.profiling) rt.prof_begin(profile_category<KSolvePositionAll>());
			e = cudaLaunchCooperativeKernel((const void *)solve_position_all_kernel, dim3((unsigned)(rt.num_sms * blocks_per_sm / W->solve_grid_div)), dim3(128), args, 0, rt.stream);
			if (rt.profiling) rt.prof_end();
		}
		if (e != cudaSuccess) { cudaGetLastError(); solved_by_phase_launches = true; } // fall back to the per phase launches below
	}
#endif
	if (M > 0 && solved_by_phase_launches)
	{
		if (num_phases == 0)
		{
			if (!read_counters(W)) return false;
			num_phases = W->h_counters.num_phases;
			W->h_phase_offsets.resize(num_phases + 1);
			rt.download(W->h_phase_offsets.data(), sc.phase_count, num_phases + 1);
		}
		psteps = W->h_counters.max_position_steps;
		for (uint32_t it = 0; it < psteps; ++it)
			for (uint32_t p = 0; p < num_phases; ++p)
			{
				uint32_t begin = W->h_phase_offsets[p], n = W->h_phase_offsets[p + 1] - begin;
				KSolvePosition k; k.w = d; k.c = sc.con; k.begin = begin; k.iteration = it;
				if (J > 0) { KMixedSolvePosition km; km.contacts = k; km.joints.w = d; km.joints.c = sc.con; km.joints.j = W->jc; km.joints.begin = begin; km.joints.iteration = it; rt.launch(km, n); }
				else if (solve_pdl_enabled()) { k.pdl = 1; rt.launch_pdl(k, n); } else rt.launch(k, n);
			}
	}

	trace.mark("integrate+position", "bounds + sleep + compaction");
	// (a16, a17) bounds, sleeping, active list compaction
	{ KBoundsAndSleep k; k.w = d; k.s = sc; k.dt = dt; k.is_last = is_last? 1u : 0u; rt.launch(k, na); }
	uint32_t new_active = na;
	if (is_last)
	{
		{ KDeactivate k; k.w = d; k.s = sc; k.keep = W->d_keep; k.events = W->d_act_events; k.max_events = W->max_act_events; rt.launch(k, na); }
		rt.exclusive_scan(W->d_keep, W->d_keep_scan, na);
		uint32_t *new_list = W->active_buf[W->active_cur ^ 1];
		{ KCompactActive k; k.w = d; k.keep = W->d_keep; k.keep_scan = W->d_keep_scan; k.new_active = new_list; rt.launch(k, na); }
		{ KFinishCompact k; k.w = d; k.keep = W->d_keep; k.keep_scan = W->d_keep_scan; k.n = na; rt.launch(k, 1); }
		W->active_cur ^= 1;
	}
	// the bodies this step simulated: the active list as it was before the sleepers left it (the buffer that is no longer current)
	W->stepped_list = is_last? W->active_buf[W->active_cur ^ 1] : W->active_buf[W->active_cur];
	W->stepped_count = na;
	if (stats != nullptr && stats->kinetic_energy < 0.0f)
	{
		rt.memset_(W->d_energy, 0, 4);
		KKineticEnergy k; k.w = d; k.out = W->d_energy; rt.launch(k, na);
	}
	{ KPublishCacheCounts k; k.w = d; rt.launch(k, 1); }
	if (!read_counters(W)) return false;
	if (is_last) new_active = W->h_counters.new_active_count;
	if (M > 0)
	{
		// (the one launch solvers read these on the device; the host only reports them)
		num_phases = W->h_counters.num_phases;
		vsteps = W->h_counters.max_velocity_steps;
		psteps = W->h_counters.max_position_steps;
	}

	trace.mark("sleep+compact+readback", "cache swap");
	// swap the caches: this step's write cache is the next step's read cache
	uint32_t wi = W->write_idx;
	// (the sizes of the cache written by this step came back with the counters: one readback at the end of the step, not three)
	W->cache_num_manifolds[wi] = W->h_counters.cache_manifolds < d.max_constraints? W->h_counters.cache_manifolds : d.max_constraints;
	W->cache_num_pairs[wi] = W->h_counters.cache_pairs < d.max_body_pairs? W->h_counters.cache_pairs : d.max_body_pairs;
	W->write_idx ^= 1;
	clear_cache(W, W->write_idx);
	W->num_active = new_active;
	sync_dworld(W);

	// every layer that holds non-static bodies needs a rebuild next step
	for (uint32_t l = 0; l < d.num_bp_layers; ++l)
		if (W->layer_has_moving[l])
			W->layer_needs_build[l] = 1;

	if (stats != nullptr)
	{
		const StepCounters &c = W->h_counters;
		stats->num_body_pairs += num_pairs;
		stats->num_pairs_from_cache += c.num_pairs_from_cache;
		stats->num_manifolds += W->cache_num_manifolds[wi];
		stats->num_contact_points += c.num_contact_points;
		stats->num_constraints += num_contacts;
		stats->num_islands += c.num_islands;
		stats->num_large_islands += c.num_large_islands;
		stats->num_phases += num_phases;
		stats->velocity_iterations = vsteps > stats->velocity_iterations? vsteps : stats->velocity_iterations;
		stats->position_iterations = psteps > stats->position_iterations? psteps : stats->position_iterations;
		stats->num_activated += woken_total;
		stats->num_deactivated += c.num_deactivated;
		stats->error_bits |= c.error_bits & 7u;
		if (stats->kinetic_energy < 0.0f) { float e = 0.0f; rt.download(&e, W->d_energy, 1); stats->kinetic_energy = e; }
	}
	W->last_num_events = W->h_counters.num_events;
	W->last_num_act_events = W->h_counters.num_activation_events;
	return rt.check("collision_step");
}

// hull points shrunk by the convex radius (ConvexHullShape::GetSupportFunction, ExcludeConvexRadius, unscaled: ConvexHullShape.cpp:551-590)
void shrink_hull_points(const b2j_hull_desc *h, std::vector<F4> &out)
{
	float cr = h->convex_radius;
	for (uint32_t i = 0; i < h->num_points; ++i)
	{
		V3 pos = v3_load(h->points + 3 * i);
		int nf = h->point_num_faces[i];
		const int32_t *faces = h->point_faces + 3 * i;
		auto plane = [&](int f, V3 &n, float &c) { n = v3_load(h->planes + 4 * f); c = h->planes[4 * f + 3]; };
		V3 new_point;
		if (cr == 0.0f || nf <= 0)
			new_point = pos;
		else if (nf == 1)
		{
			V3 n; float c; plane(faces[0], n, c);
			new_point = pos - n * cr;
		}
		else
		{
			// planes offset inwards: Plane::Offset(-r) = (n, c - (-r))
			V3 n1, n2, n3; float c1, c2, c3;
			plane(faces[0], n1, c1); c1 = c1 - (-cr);
			plane(faces[1], n2, c2); c2 = c2 - (-cr);
			if (nf == 3) { plane(faces[2], n3, c3); c3 = c3 - (-cr); }
			else { n3 = cross(n1, n2); c3 = -dot(n3, pos); }
			// Plane::sIntersectPlanes (Plane.h:63-93)
			float denominator = dot(n1, cross(n2, n3));
			if (denominator == 0.0f)
				new_point = pos - n1 * cr;
			else
			{
				float ax = n1.x, ay = n1.y, az = n1.z, aw = c1, bx = n2.x, by = n2.y, bz = n2.z, bw = c2, cx = n3.x, cy = n3.y, cz = n3.z, cw = c3;
				V3 numerator = v3(
					aw * (bz * cy - by * cz) + ay * (bw * cz - bz * cw) + az * (by * cw - bw * cy),
					aw * (bx * cz - bz * cx) + ax * (bz * cw - bw * cz) + az * (bw * cx - bx * cw),
					aw * (by * cx - bx * cy) + ax * (bw * cy - by * cw) + ay * (bx * cw - bw * cx));
				new_point = numerator / denominator;
			}
		}
		out.push_back(f4(new_point));
	}
}

} // namespace

// ---- C ABI (the only symbols with default visibility; everything else is hidden: -fvisibility=hidden) -------------------

#pragma GCC visibility push(default)
extern "C" {

const char *b2j_last_error(void) { return last_error().c_str(); }

void b2j_settings_default(b2j_settings *s)
{
	memset(s, 0, sizeof(*s));
	s->speculative_contact_distance = 0.02f;
	s->penetration_slop = 0.02f;
	s->baumgarte = 0.2f;
	s->max_penetration_distance = 0.2f;
	s->manifold_tolerance = 1.0e-3f;
	s->body_pair_cache_max_delta_position_sq = 0.001f * 0.001f;
	s->body_pair_cache_cos_max_delta_rotation_div2 = 0.99984769515639123915701155881391f;
	s->contact_normal_cos_max_delta_rotation = 0.99619469809174553229501040247389f;
	s->contact_point_preserve_lambda_max_dist_sq = 0.01f * 0.01f;
	s->min_velocity_for_restitution = 1.0f;
	s->time_before_sleep = 0.5f;
	s->point_velocity_sleep_threshold = 0.03f;
	s->num_velocity_steps = 10;
	s->num_position_steps = 2;
	s->deterministic_simulation = 1;
	s->constraint_warm_start = 1;
	s->use_body_pair_contact_cache = 1;
	s->use_manifold_reduction = 1;
	s->use_large_island_splitter = 1;
	s->allow_sleeping = 1;
	s->check_active_edges = 1;
}

static bool g_create_without_events = false;

b2j_world *b2j_world_create(const b2j_world_desc *desc)
{
	if (desc == nullptr || desc->max_bodies == 0 || desc->num_object_layers == 0 || desc->num_object_layers > 64 || desc->num_broadphase_layers == 0 || desc->num_broadphase_layers > 8)
	{
		last_error() = "invalid world description";
		return nullptr;
	}
	b2j_world *W = new b2j_world;
	if (!W->rt.init(desc->device)) { delete W; return nullptr; }
	Runtime &rt = W->rt;
	W->desc = *desc;
	uint32_t no = desc->num_object_layers, nb = desc->num_broadphase_layers;
	W->t_o2bp.assign(desc->object_to_broadphase, desc->object_to_broadphase + no);
	W->t_ovbp.assign(desc->object_vs_broadphase, desc->object_vs_broadphase + no * nb);
	W->t_ovo.assign(desc->object_vs_object, desc->object_vs_object + no * no);
	W->d_o2bp = rt.alloc<uint8_t>(no); rt.upload(W->d_o2bp, W->t_o2bp.data(), no);
	W->d_ovbp = rt.alloc<uint8_t>(no * nb); rt.upload(W->d_ovbp, W->t_ovbp.data(), no * nb);
	W->d_ovo = rt.alloc<uint8_t>(no * no); rt.upload(W->d_ovo, W->t_ovo.data(), no * no);

	DWorld &d = W->d;
	memset(&d, 0, sizeof(d));
	uint32_t nbod = desc->max_bodies;
	d.max_bodies = nbod;
	d.max_body_pairs = desc->max_body_pairs < 4? 4 : desc->max_body_pairs;
	d.max_constraints = desc->max_contact_constraints < 4? 4 : desc->max_contact_constraints;
	d.pair_table_size = next_pow2(d.max_body_pairs * 2);
	d.num_object_layers = no; d.num_bp_layers = nb;
	d.settings = desc->settings;
	d.gravity = v3_load(desc->gravity);
	d.object_to_bp = W->d_o2bp; d.object_vs_bp = W->d_ovbp; d.object_vs_object = W->d_ovo;

	d.info = rt.alloc<BodyInfo>(nbod); d.params = rt.alloc<BodyParams>(nbod);
	// 32 byte pairs (see F4PairView): pose, velocity, force / torque, inertia, bounds
	d.position.base = d.rotation.base = rt.alloc<F4>(2 * (size_t)nbod);
	d.linear_velocity.base = d.angular_velocity.base = rt.alloc<F4>(2 * (size_t)nbod);
	d.force.base = d.torque.base = rt.alloc<F4>(2 * (size_t)nbod);
	d.inv_inertia_diag.base = d.inertia_rotation.base = rt.alloc<F4>(2 * (size_t)nbod);
	d.bounds_min.base = d.bounds_max.base = rt.alloc<F4>(2 * (size_t)nbod);
	d.sleep_spheres = rt.alloc<F4>((size_t)nbod * 3); d.sleep_timer = rt.alloc<float>(nbod);
	d.active_index = rt.alloc<uint32_t>(nbod);
	rt.memset_(d.active_index, 0xff, (size_t)nbod * 4);
	W->active_buf[0] = rt.alloc<uint32_t>(nbod); W->active_buf[1] = rt.alloc<uint32_t>(nbod);
	d.num_active = rt.alloc<uint32_t>(1);
	d.counters = rt.alloc<StepCounters>(1);
	W->d_keep = rt.alloc<uint32_t>(nbod); W->d_keep_scan = rt.alloc<uint32_t>(nbod);

	for (int i = 0; i < 2; ++i)
	{
		W->cache[i].pairs = rt.alloc<CachedPair>(d.max_body_pairs);
		W->cache[i].manifolds = rt.alloc<CachedManifold>(d.max_constraints);
		W->cache[i].pair_table = rt.alloc<uint32_t>(d.pair_table_size);
		W->cache[i].num_pairs = rt.alloc<uint32_t>(1);
		W->cache[i].num_manifolds = rt.alloc<uint32_t>(1);
		rt.memset_(W->cache[i].pair_table, 0xff, (size_t)d.pair_table_size * 4);
	}

	NarrowCtx &nc = W->nc;
	memset(&nc, 0, sizeof(nc));
	nc.pairs = rt.alloc<BodyPair>(d.max_body_pairs);
	nc.collide_convex = rt.alloc<CollideItem>(d.max_body_pairs);
	nc.collide_mesh = rt.alloc<CollideItem>(d.max_body_pairs);
	nc.cached = rt.alloc<CachedItem>(d.max_body_pairs);
	nc.max_epa = d.max_body_pairs;
	nc.epa = rt.alloc<EpaItem>(nc.max_epa, false);
	nc.epa_overflow = rt.alloc<EpaItem>(nc.max_epa, false);
	nc.num_epa_overflow = rt.alloc<uint32_t>(1);
	nc.epa_hist = getenv("B2J_TRACE_EPA") != nullptr? rt.alloc<uint32_t>(130) : nullptr;
	nc.epa_results = rt.alloc<EpaResult>(nc.max_epa, false);
	nc.num_epa_results = rt.alloc<uint32_t>(1);
#ifndef B2J_HOSTSIM
	nc.num_scratch = (uint32_t)rt.num_sms * 8; // 2 blocks of 4 warps per SM, each warp owns an EpaScratch in shared memory
#else
	nc.num_scratch = 1;
#endif
	nc.man_ws = rt.alloc<ManifoldWS>(d.max_constraints, false);
	nc.con_src = rt.alloc<ConstraintSrc>(d.max_constraints, false);
	nc.woken_flag = rt.alloc<uint32_t>(nbod);
	nc.woken_list = rt.alloc<uint32_t>(nbod);
	if (g_create_without_events)
	{
		// batched worlds: events are not recorded (no per world listener replay)
		W->max_events = 0; nc.events = nullptr; nc.max_events = 0; W->max_act_events = 0; W->d_act_events = nullptr;
	}
	else
	{
		W->max_events = 2 * d.max_constraints + 16;
		nc.events = W->events_buf = rt.alloc<b2j_contact_event>(W->max_events, false);
		nc.max_events = W->max_events;
		W->max_act_events = 2 * nbod;
		W->d_act_events = W->act_events_buf = rt.alloc<b2j_activation_event>(W->max_act_events, false);
	}
	W->d_woken_sorted = rt.alloc<uint32_t>(nbod); W->d_woken_keys = rt.alloc<uint32_t>(nbod);
	W->d_round_begin = rt.alloc<uint32_t>(1);
	W->d_energy = rt.alloc<float>(1);

	SolveCtx &sc = W->sc;
	memset(&sc, 0, sizeof(sc));
	uint32_t mc = d.max_constraints;
	sc.con.capacity = mc;
	sc.con.cp = rt.alloc<F4>((size_t)CP_NUM * mc, false);
	sc.con.hdr = rt.alloc<ConstraintHeader>(mc);
	sc.man_ws = nc.man_ws;
	sc.order = rt.alloc<uint32_t>(mc); sc.final_pos = rt.alloc<uint32_t>(mc); sc.solve_src = rt.alloc<uint32_t>(mc); sc.phase = rt.alloc<uint32_t>(mc);
	sc.max_phases = 8192;
	rt.reserve_temp(std::max(d.max_bodies, d.max_constraints), std::max(std::max(d.max_body_pairs, d.max_constraints), d.max_bodies), d.max_bodies);
	sc.phase_count = rt.alloc<uint32_t>(sc.max_phases + 2);
	sc.uf_parent = rt.alloc<uint32_t>(nbod); sc.root = rt.alloc<uint32_t>(nbod); sc.island_items = rt.alloc<uint32_t>(nbod);
	sc.island_large = rt.alloc<uint32_t>(nbod); sc.island_steps = rt.alloc<uint32_t>(nbod); sc.island_can_sleep = rt.alloc<uint32_t>(nbod);
	sc.large_color_count = rt.alloc<uint32_t>((size_t)(mc / 128 + 2) * 32);
	sc.body_deg = rt.alloc<uint32_t>(nbod + 1); sc.body_off = rt.alloc<uint32_t>(nbod + 1); sc.body_fill = rt.alloc<uint32_t>(nbod);
	sc.body_cur = rt.alloc<uint32_t>(nbod); sc.body_mask = rt.alloc<uint32_t>(nbod);
	sc.adj = rt.alloc<uint32_t>((size_t)2 * mc);
	sc.sched_flag = rt.alloc<uint32_t>(4096);
	sc.grid_barrier = rt.alloc<uint32_t>(1);
	W->d_sort_keys[0] = rt.alloc<uint64_t>(mc); W->d_sort_keys[1] = rt.alloc<uint64_t>(mc);
	W->d_sort_vals = rt.alloc<uint32_t>(mc);

	if (W->d_sort_vals == nullptr || sc.con.cp == nullptr )
	{
		last_error() = "out of device memory";
		b2j_world_destroy(W);
		return nullptr;
	}

	memset(W->trees, 0, sizeof(W->trees));
	for (uint32_t l = 0; l < nb; ++l)
	{
		W->trees[l].layer_bounds = rt.alloc<F4>(2);
		F4 init[2] = { f4(-1000.0f, -1000.0f, -1000.0f, 0.0f), f4(1000.0f, 1000.0f, 1000.0f, 0.0f) };
		rt.upload(W->trees[l].layer_bounds, init, 2);
	}
	W->layer_bodies.resize(nb);
	W->layer_list_dirty.assign(nb, 1);
	W->layer_needs_build.assign(nb, 1);
	W->layer_has_moving.assign(nb, 0);
	W->h_ids.assign(nbod, B2J_INVALID_ID);
	W->h_layer.assign(nbod, 0);
	W->h_static.assign(nbod, 0);
	W->shapes_dirty = true;
#ifndef B2J_HOSTSIM
	cudaEventCreate(&W->ev_begin);
	cudaEventCreate(&W->ev_end);
#endif
	sync_dworld(W);
	rt.sync();
	if (!rt.check("b2j_world_create")) { b2j_world_destroy(W); return nullptr; }
	return W;
}

void b2j_world_destroy(b2j_world *W)
{
	if (W == nullptr) return;
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	rt.sync();
	DWorld &d = W->d;
	rt.free_(W->d_o2bp); rt.free_(W->d_ovbp); rt.free_(W->d_ovo);
	rt.free_(d.info); rt.free_(d.params); rt.free_(d.position.base); rt.free_(d.linear_velocity.base);
	rt.free_(d.force.base); rt.free_(d.inv_inertia_diag.base); rt.free_(d.bounds_min.base);
	rt.free_(d.sleep_spheres); rt.free_(d.sleep_timer); rt.free_(d.active_index); rt.free_(W->active_buf[0]); rt.free_(W->active_buf[1]);
	rt.free_(d.num_active); rt.free_(d.counters); rt.free_(W->d_keep); rt.free_(W->d_keep_scan);
	for (int i = 0; i < 2; ++i)
	{
		rt.free_(W->cache[i].pairs); rt.free_(W->cache[i].manifolds); rt.free_(W->cache[i].pair_table); rt.free_(W->cache[i].num_pairs); rt.free_(W->cache[i].num_manifolds);
	}
	NarrowCtx &nc = W->nc;
	rt.free_(nc.pairs); rt.free_(nc.collide_convex); rt.free_(nc.collide_mesh); rt.free_(nc.cached); rt.free_(nc.epa); rt.free_(nc.epa_overflow); rt.free_(nc.num_epa_overflow); if (nc.epa_hist != nullptr) rt.free_(nc.epa_hist); rt.free_(nc.epa_results); rt.free_(nc.num_epa_results);
	rt.free_(nc.man_ws); rt.free_(nc.con_src); rt.free_(nc.woken_flag); rt.free_(nc.woken_list); rt.free_(W->events_buf);
	if (W->joint_capacity != 0)
	{
		JointCtx &j = W->jc;
		rt.free_(j.defs); rt.free_(j.state); rt.free_(j.active_flag); rt.free_(j.order_flag); rt.free_(j.order_scan); rt.free_(j.active_joints); rt.free_(j.wake_key);
		uint32_t *p1 = const_cast<uint32_t *>(j.order); rt.free_(p1);
		uint32_t *p2 = const_cast<uint32_t *>(W->sc.joint_steps); rt.free_(p2);
		rt.free_(W->sc.body_nj); rt.free_(W->d_joint_init);
	}
	rt.free_(W->d_mesh_scratch); rt.free_(W->d_query_scratch); rt.free_(W->d_cache_invalid);
	for (int i = 0; i < 2; ++i) { rt.free_(W->d_collide_keys[i]); rt.free_(W->d_collide_vals[i]); }
	rt.free_(W->act_events_buf); rt.free_(W->d_woken_sorted); rt.free_(W->d_woken_keys); rt.free_(W->d_round_begin); rt.free_(W->d_energy);
	SolveCtx &sc = W->sc;
	rt.free_(sc.con.cp); rt.free_(sc.con.hdr);
	rt.free_(sc.order); rt.free_(sc.final_pos); rt.free_(sc.solve_src); rt.free_(sc.phase); rt.free_(sc.phase_count);
	rt.free_(sc.uf_parent); rt.free_(sc.root); rt.free_(sc.island_items); rt.free_(sc.island_large); rt.free_(sc.island_steps); rt.free_(sc.island_can_sleep);
	rt.free_(sc.large_color_count); rt.free_(sc.body_deg); rt.free_(sc.body_off); rt.free_(sc.body_fill); rt.free_(sc.body_cur); rt.free_(sc.body_mask);
	rt.free_(sc.adj); rt.free_(sc.sched_flag); rt.free_(sc.grid_barrier);
	rt.free_(W->d_sort_keys[0]); rt.free_(W->d_sort_keys[1]); rt.free_(W->d_sort_vals);
	for (int l = 0; l < 8; ++l)
	{
		Tree &t = W->trees[l];
		rt.free_(t.bodies); rt.free_(t.keys_in); rt.free_(t.keys_out); rt.free_(t.leaf_body); rt.free_(t.child_left); rt.free_(t.child_right);
		rt.free_(t.parent); rt.free_(t.node_min); rt.free_(t.node_max); rt.free_(t.visit); rt.free_(t.layer_bounds); rt.free_(t.world_root);
	}
	rt.free_(W->d_shapes); rt.free_(W->d_hull_points); rt.free_(W->d_hull_shrunk); rt.free_(W->d_hull_planes);
	rt.free_(W->d_hull_faces); rt.free_(W->d_hull_vtx); rt.free_(W->d_mesh_bytes); rt.free_(W->d_compound_subs);
#ifndef B2J_HOSTSIM
	if (W->ev_begin) cudaEventDestroy(W->ev_begin);
	if (W->ev_end) cudaEventDestroy(W->ev_end);
#endif
	rt.shutdown();
	delete W;
}

int b2j_world_set_gravity(b2j_world *W, const float g[3]) { W->d.gravity = v3_load(g); return 0; }
int b2j_world_set_settings(b2j_world *W, const b2j_settings *s) { W->d.settings = *s; return 0; }
int b2j_world_get_settings(const b2j_world *W, b2j_settings *s) { *s = W->d.settings; return 0; }
int b2j_world_set_previous_delta_time(b2j_world *W, float dt) { W->prev_dt = dt; return 0; }

int b2j_world_set_event_recording(b2j_world *W, int contact_events, int activation_events)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	if (W->num_worlds > 1 && (contact_events || activation_events)) { last_error() = "events are not recorded for batched worlds"; return -1; }
	rt.sync();
	if (contact_events && W->events_buf == nullptr)
	{
		W->max_events = 2 * W->d.max_constraints + 16;
		W->events_buf = rt.alloc<b2j_contact_event>(W->max_events, false);
		if (W->events_buf == nullptr) return -1;
	}
	if (activation_events && W->act_events_buf == nullptr)
	{
		W->max_act_events = 2 * W->d.max_bodies;
		W->act_events_buf = rt.alloc<b2j_activation_event>(W->max_act_events, false);
		if (W->act_events_buf == nullptr) return -1;
	}
	// (the buffers are sized for the world limits, 2 x 148 bytes per contact constraint: released while nobody listens)
	if (!contact_events) rt.free_(W->events_buf);
	if (!activation_events) rt.free_(W->act_events_buf);
	W->nc.events = W->events_buf;
	W->nc.max_events = contact_events? W->max_events : 0;
	W->d_act_events = W->act_events_buf;
	if (!contact_events) W->last_num_events = 0;
	if (!activation_events) W->last_num_act_events = 0;
	return 0;
}

int b2j_world_set_profiling(b2j_world *W, int on)
{
	B2J_DEVICE_GUARD(W);
	W->rt.sync();
	W->rt.prof_reset();
	W->rt.profiling = on != 0;
	return 0;
}

uint32_t b2j_world_get_profile(b2j_world *W, char *names, uint32_t name_stride, float *ms, uint32_t *launches, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	W->rt.sync();
	W->rt.prof_collect();
	std::lock_guard<std::mutex> lock(profile_mutex());
	uint32_t n = 0;
	for (size_t i = 0; i < W->rt.prof_ms.size(); ++i)
	{
		if (W->rt.prof_launches[i] == 0) continue;
		if (n < cap)
		{
			if (names != nullptr && name_stride > 0) { strncpy(names + (size_t)n * name_stride, profile_names()[i].c_str(), name_stride - 1); names[(size_t)n * name_stride + name_stride - 1] = 0; }
			if (ms != nullptr) ms[n] = (float)W->rt.prof_ms[i];
			if (launches != nullptr) launches[n] = W->rt.prof_launches[i];
		}
		++n;
	}
	return n;
}

static int32_t add_shape(b2j_world *W, const ShapeDesc &desc, const b2j_world::ShapeMeta *meta = nullptr)
{
	ShapeDesc s = desc;
	if (meta == nullptr)
	{
		// a leaf: no decoration
		s.flags = 0; s.scale = v3_rep(1.0f); s.local_rot = m33_rotation(q4_identity());
		s.outer_min = s.local_min; s.outer_max = s.local_max;
		s.hull_orig_offset = s.hull_point_offset;
		s.base_leaf = (uint32_t)W->h_shapes.size();
	}
	W->h_shapes.push_back(s);
	W->h_shape_meta.resize(W->h_shapes.size());
	if (meta != nullptr) W->h_shape_meta.back() = *meta;
	else W->h_shape_meta.back().leaf = (int32_t)W->h_shapes.size() - 1;
	W->shapes_dirty = true;
	return (int32_t)W->h_shapes.size() - 1;
}

int32_t b2j_shape_sphere(b2j_world *W, float radius)
{
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_SPHERE; s.radius = radius; s.inner_radius = radius;
	s.local_min = -v3_rep(radius); s.local_max = v3_rep(radius);
	return add_shape(W, s);
}

int32_t b2j_shape_box(b2j_world *W, const float he[3], float convex_radius)
{
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_BOX; s.half_extent = v3_load(he); s.convex_radius = convex_radius;
	s.inner_radius = reduce_min(s.half_extent);
	s.local_min = -s.half_extent; s.local_max = s.half_extent;
	return add_shape(W, s);
}

int32_t b2j_shape_capsule(b2j_world *W, float half_height, float radius)
{
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_CAPSULE; s.half_height = half_height; s.radius = radius; s.inner_radius = radius;
	V3 extent = v3_rep(radius) + v3(0.0f, half_height, 0.0f);
	s.local_min = -extent; s.local_max = extent;
	return add_shape(W, s);
}

int32_t b2j_shape_cylinder(b2j_world *W, float half_height, float radius, float convex_radius)
{
	if (!(half_height >= 0.0f && radius >= 0.0f && convex_radius >= 0.0f)) { last_error() = "invalid cylinder"; return -1; }
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_CYLINDER; s.half_height = half_height; s.radius = radius;
	s.convex_radius = fmin_(convex_radius, fmin_(half_height, radius)); // CylinderShape::CylinderShape
	s.inner_radius = fmin_(half_height, radius);
	V3 extent = v3(radius, half_height, radius);
	s.local_min = -extent; s.local_max = extent;
	return add_shape(W, s);
}

int32_t b2j_shape_convex_hull(b2j_world *W, const b2j_hull_desc *h)
{
	if (h == nullptr || h->num_points == 0 || h->num_points > 256 || h->num_faces == 0) { last_error() = "invalid hull"; return -1; }
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_CONVEX_HULL; s.convex_radius = h->convex_radius; s.inner_radius = h->inner_radius;
	s.local_min = v3_load(h->local_bounds_min); s.local_max = v3_load(h->local_bounds_max);
	s.center_of_mass = v3_load(h->center_of_mass);
	s.hull_point_offset = (uint32_t)W->h_hull_points.size(); s.hull_num_points = h->num_points;
	for (uint32_t i = 0; i < h->num_points; ++i) W->h_hull_points.push_back(f4(v3_load(h->points + 3 * i)));
	shrink_hull_points(h, W->h_hull_shrunk);
	s.hull_face_offset = (uint32_t)W->h_hull_planes.size(); s.hull_num_faces = h->num_faces;
	for (uint32_t i = 0; i < h->num_faces; ++i)
	{
		W->h_hull_planes.push_back(f4(h->planes[4 * i], h->planes[4 * i + 1], h->planes[4 * i + 2], h->planes[4 * i + 3]));
		W->h_hull_faces.push_back((uint32_t)h->face_first_vertex[i] | ((uint32_t)h->face_num_vertices[i] << 16));
	}
	s.hull_vtx_offset = (uint32_t)W->h_hull_vtx.size();
	W->h_hull_vtx.insert(W->h_hull_vtx.end(), h->vertex_idx, h->vertex_idx + h->num_vertex_idx);
	int32_t id = add_shape(W, s);
	W->h_shape_meta[id].point_num_faces.assign(h->point_num_faces, h->point_num_faces + h->num_points);
	W->h_shape_meta[id].point_faces.assign(h->point_faces, h->point_faces + 3 * (size_t)h->num_points);
	return id;
}

int32_t b2j_shape_static_compound(b2j_world *W, const b2j_compound_desc *cd)
{
	if (cd == nullptr || cd->num_subs == 0 || cd->subs == nullptr || cd->num_nodes == 0 || cd->nodes == nullptr) { last_error() = "invalid compound"; return -1; }
	for (uint32_t i = 0; i < cd->num_subs; ++i)
	{
		int32_t sh = cd->subs[i].shape;
		if (sh < 0 || sh >= (int32_t)W->h_shapes.size()) { last_error() = "compound: invalid sub shape id"; return -1; }
		uint32_t kind = W->h_shapes[sh].kind;
		if (kind == B2J_SHAPE_MESH || kind == B2J_SHAPE_COMPOUND) { last_error() = "compound: sub shapes must be convex"; return -1; }
	}
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_COMPOUND; s.inner_radius = cd->inner_radius;
	s.local_min = v3_load(cd->local_bounds_min); s.local_max = v3_load(cd->local_bounds_max);
	s.center_of_mass = v3_load(cd->center_of_mass);
	while (W->h_mesh_bytes.size() % 16 != 0) W->h_mesh_bytes.push_back(0);
	s.mesh_offset = (uint32_t)W->h_mesh_bytes.size(); s.mesh_size = cd->num_nodes * 64;
	W->h_mesh_bytes.insert(W->h_mesh_bytes.end(), cd->nodes, cd->nodes + (size_t)cd->num_nodes * 64);
	s.compound_sub_offset = (uint32_t)W->h_compound_subs.size(); s.compound_num_subs = cd->num_subs;
	// CompoundShape::GetSubShapeIDBits: 32 - CountLeadingZeros(n - 1)
	uint32_t bits = 0;
	while (bits < 32 && (cd->num_subs - 1) >> bits != 0) ++bits;
	s.compound_sub_bits = bits;
	for (uint32_t i = 0; i < cd->num_subs; ++i)
	{
		CompoundSub sub; memset(&sub, 0, sizeof(sub));
		sub.shape = (uint32_t)cd->subs[i].shape;
		sub.position_com = f4(v3_load(cd->subs[i].position_com));
		sub.rotation = f4(cd->subs[i].rotation[0], cd->subs[i].rotation[1], cd->subs[i].rotation[2], cd->subs[i].rotation[3]);
		W->h_compound_subs.push_back(sub);
	}
	return add_shape(W, s);
}

// The device description of `leaf` under an accumulated scale and rotation (what the reference's dispatch hands to the leaf's collide /
// support / bounds functions after peeling the decorators, ScaledShape.cpp:190-204, RotatedTranslatedShape.cpp:183-192)
static int32_t add_decorated_shape(b2j_world *W, int32_t leaf, V3 scale, bool scaled, Q4 rotation, bool rotated, V3 center_of_mass)
{
	const ShapeDesc base = W->h_shapes[leaf];
	ShapeDesc s = base;
	s.flags = 0; s.scale = v3_rep(1.0f); s.local_rot = m33_rotation(q4_identity());
	s.hull_orig_offset = base.hull_point_offset;
	s.base_leaf = (uint32_t)leaf;
	s.center_of_mass = center_of_mass;
	if (scaled)
	{
		V3 abs_scale = v3_abs(scale);
		s.scale = scale;
		// GetLocalBounds().Scaled(inScale) (ConvexShape.cpp:52-56, AABox::Scaled): what the OBB pre-test and the default world bounds use
		s.local_min = v3_min(base.local_min * scale, base.local_max * scale); s.local_max = v3_max(base.local_min * scale, base.local_max * scale);
		switch (base.kind)
		{
		case B2J_SHAPE_SPHERE: // SphereShape::GetSupportFunction / GetWorldSpaceBounds: scaled_radius = abs(scale.x) * mRadius
			s.radius = abs_scale.x * base.radius; s.inner_radius = s.radius;
			break;
		case B2J_SHAPE_BOX: // BoxShape::GetSupportFunction: scaled half extent = |scale| * h, convex radius = ScaleHelpers::ScaleConvexRadius
			s.half_extent = abs_scale * base.half_extent;
			s.convex_radius = fmin_(base.convex_radius * reduce_min(abs_scale), 0.05f /* cDefaultConvexRadius */);
			s.inner_radius = reduce_min(s.half_extent);
			break;
		case B2J_SHAPE_CAPSULE: // CapsuleShape::GetSupportFunction: abs_scale = |scale.x| for both
			s.half_height = abs_scale.x * base.half_height; s.radius = abs_scale.x * base.radius; s.inner_radius = s.radius;
			break;
		case B2J_SHAPE_CYLINDER: // CylinderShape::GetSupportFunction: scale_xz = |scale.x|, scale_y = |scale.y| (IsValidScale: x == z)
			s.half_height = abs_scale.y * base.half_height; s.radius = abs_scale.x * base.radius;
			s.convex_radius = fmin_(base.convex_radius * reduce_min(abs_scale), 0.05f);
			s.inner_radius = fmin_(s.half_height, s.radius);
			break;
		case B2J_SHAPE_MESH: // the mesh kernels scale node bounds and vertices on the fly (s.scale)
			break;
		default: // ConvexHullShape::GetSupportFunction with a scale (ConvexHullShape.cpp:486-657)
			{
				const b2j_world::ShapeMeta &bm = W->h_shape_meta[leaf];
				s.flags |= SHAPE_SCALED_HULL;
				s.convex_radius = fmin_(base.convex_radius * reduce_min(abs_scale), 0.05f);
				s.hull_point_offset = (uint32_t)W->h_hull_points.size();
				V3 inv_scale = v3(1.0f / scale.x, 1.0f / scale.y, 1.0f / scale.z);
				// planes: inv_scale * normal (GetSupportingFace divides by its length itself), constant unchanged (nobody reads it on this path)
				s.hull_face_offset = (uint32_t)W->h_hull_planes.size();
				for (uint32_t f = 0; f < base.hull_num_faces; ++f)
				{
					F4 pl = W->h_hull_planes[base.hull_face_offset + f];
					V3 n = inv_scale * to_v3(pl);
					W->h_hull_planes.push_back(f4(n.x, n.y, n.z, pl.w));
					W->h_hull_faces.push_back(W->h_hull_faces[base.hull_face_offset + f]);
				}
				float cr = s.convex_radius;
				for (uint32_t i = 0; i < base.hull_num_points; ++i)
				{
					V3 pos = scale * to_v3(W->h_hull_points[base.hull_point_offset + i]);
					W->h_hull_points.push_back(f4(pos));
					V3 new_point = pos;
					int nf = bm.point_num_faces[i];
					const int32_t *faces = bm.point_faces.data() + 3 * (size_t)i;
					auto normal = [&](int f) { return normalized(inv_scale * to_v3(W->h_hull_planes[base.hull_face_offset + f])); };
					if (base.convex_radius != 0.0f && nf > 0)
					{
						V3 n1 = normal(faces[0]);
						if (nf == 1)
							new_point = pos - n1 * cr;
						else
						{
							V3 n2 = normal(faces[1]);
							// Plane::sFromPointAndNormal(pos, n).Offset(-cr): constant = -n.pos, then c - (-cr)
							float c1 = -dot(n1, pos) - (-cr), c2 = -dot(n2, pos) - (-cr), c3;
							V3 n3;
							if (nf == 3) { n3 = normal(faces[2]); c3 = -dot(n3, pos) - (-cr); }
							else { n3 = cross(n1, n2); c3 = -dot(n3, pos); }
							float denominator = dot(n1, cross(n2, n3));
							if (denominator == 0.0f)
								new_point = pos - n1 * cr;
							else
							{
								float ax = n1.x, ay = n1.y, az = n1.z, aw = c1, bx = n2.x, by = n2.y, bz = n2.z, bw = c2, cx = n3.x, cy = n3.y, cz = n3.z, cw = c3;
								V3 numerator = v3(
									aw * (bz * cy - by * cz) + ay * (bw * cz - bz * cw) + az * (by * cw - bw * cy),
									aw * (bx * cz - bz * cx) + ax * (bz * cw - bw * cz) + az * (bw * cx - bx * cw),
									aw * (by * cx - bx * cy) + ax * (bw * cy - by * cw) + ay * (bx * cw - bw * cx));
								new_point = numerator / denominator;
							}
						}
					}
					W->h_hull_shrunk.push_back(f4(new_point));
				}
			}
			break;
		}
	}
	s.outer_min = s.local_min; s.outer_max = s.local_max;
	if (rotated)
	{
		s.flags |= SHAPE_LOCAL_ROTATION;
		s.local_rot = m33_rotation(rotation);
		// RotatedTranslatedShape::GetLocalBounds: inner bounds .Transformed(Mat44::sRotation(mRotation)) (AABox.h:193-213)
		V3 new_min = v3_zero(), new_max = v3_zero();
		for (int c = 0; c < 3; ++c)
		{
			V3 col = m33_col(s.local_rot, c);
			V3 a = col * v3_get(s.local_min, c), b = col * v3_get(s.local_max, c);
			new_min += v3_min(a, b); new_max += v3_max(a, b);
		}
		s.outer_min = new_min; s.outer_max = new_max;
	}
	b2j_world::ShapeMeta meta;
	meta.leaf = leaf; meta.scale = scale; meta.scaled = scaled; meta.rotation = rotation; meta.rotated = rotated;
	return add_shape(W, s, &meta);
}

static bool is_uniform_scale(V3 s) { V3 d = v3(s.y, s.z, s.x) - s; return length_sq(d) <= 1.0e-8f; } // ScaleHelpers::IsUniformScale

int32_t b2j_shape_scaled(b2j_world *W, int32_t inner, const float scale_in[3])
{
	if (inner < 0 || inner >= (int32_t)W->h_shapes.size() || scale_in == nullptr) { last_error() = "invalid shape id"; return -1; }
	V3 scale = v3_load(scale_in);
	if (!(scale.x > 1.0e-6f && scale.y > 1.0e-6f && scale.z > 1.0e-6f)) { last_error() = "ScaledShape: only positive scales are supported"; return -1; }
	const b2j_world::ShapeMeta im = W->h_shape_meta[inner];
	const ShapeDesc &leaf = W->h_shapes[im.leaf];
	if (im.scaled) { last_error() = "ScaledShape: nested scales are not supported"; return -1; }
	if (leaf.kind == B2J_SHAPE_COMPOUND) { last_error() = "ScaledShape: a compound cannot be decorated"; return -1; }
	if ((leaf.kind == B2J_SHAPE_SPHERE || leaf.kind == B2J_SHAPE_CAPSULE || im.rotated) && !is_uniform_scale(scale))
	{ last_error() = "ScaledShape: this inner shape only takes a uniform scale"; return -1; }
	if (leaf.kind == B2J_SHAPE_CYLINDER && !(square(scale.z - scale.x) <= 1.0e-8f)) { last_error() = "ScaledShape: a cylinder takes the same scale in x and z"; return -1; }
	// ScaledShape::sCollideScaledVsShape: the inner shape sees inScale * mScale; ScaledShape::GetCenterOfMass = mScale * inner centre of mass
	const ShapeDesc inner_desc = W->h_shapes[inner];
	int32_t id = add_decorated_shape(W, im.leaf, scale, true, im.rotation, im.rotated, scale * inner_desc.center_of_mass);
	if (id >= 0 && im.rotated)
	{
		// ScaledShape::GetLocalBounds of a rotated inner shape: the inner shape's (rotated) bounds, scaled
		ShapeDesc &s = W->h_shapes[id];
		s.outer_min = v3_min(inner_desc.outer_min * scale, inner_desc.outer_max * scale); s.outer_max = v3_max(inner_desc.outer_min * scale, inner_desc.outer_max * scale);
	}
	return id;
}

int32_t b2j_shape_rotated_translated(b2j_world *W, int32_t inner, const float rotation[4], const float center_of_mass[3])
{
	if (inner < 0 || inner >= (int32_t)W->h_shapes.size() || rotation == nullptr || center_of_mass == nullptr) { last_error() = "invalid shape id"; return -1; }
	const b2j_world::ShapeMeta im = W->h_shape_meta[inner];
	if (im.rotated) { last_error() = "RotatedTranslatedShape: nested rotations are not supported"; return -1; }
	if (W->h_shapes[im.leaf].kind == B2J_SHAPE_COMPOUND) { last_error() = "RotatedTranslatedShape: a compound cannot be decorated"; return -1; }
	Q4 q; q.x = rotation[0]; q.y = rotation[1]; q.z = rotation[2]; q.w = rotation[3];
	return add_decorated_shape(W, im.leaf, im.scale, im.scaled, q, true, v3_load(center_of_mass));
}

int32_t b2j_shape_mesh(b2j_world *W, const b2j_mesh_desc *m)
{
	if (m == nullptr || m->tree == nullptr || m->tree_size < 32) { last_error() = "invalid mesh"; return -1; }
	ShapeDesc s; memset(&s, 0, sizeof(s));
	s.kind = B2J_SHAPE_MESH;
	s.local_min = v3_load(m->local_bounds_min); s.local_max = v3_load(m->local_bounds_max);
	while (W->h_mesh_bytes.size() % 16 != 0) W->h_mesh_bytes.push_back(0);
	s.mesh_offset = (uint32_t)W->h_mesh_bytes.size(); s.mesh_size = m->tree_size;
	W->h_mesh_bytes.insert(W->h_mesh_bytes.end(), m->tree, m->tree + m->tree_size);
	return add_shape(W, s);
}

int b2j_bodies_add(b2j_world *W, const b2j_body_desc *bodies, uint32_t n)
{
	B2J_DEVICE_GUARD(W);
	if (n == 0) return 0;
	Runtime &rt = W->rt;
	upload_shapes(W);
	std::vector<uint32_t> to_activate;
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(bodies[i].id);
		if (slot >= W->d.max_bodies) { last_error() = "body index out of range"; return -1; }
		if (W->h_ids[slot] != B2J_INVALID_ID) { last_error() = "body slot already in use"; return -1; }
		if (bodies[i].motion_type > B2J_MOTION_DYNAMIC) { last_error() = "invalid motion type"; return -1; }
		if (bodies[i].shape < 0 || bodies[i].shape >= (int32_t)W->h_shapes.size()) { last_error() = "invalid shape id"; return -1; }
		if (bodies[i].object_layer >= W->d.num_object_layers) { last_error() = "invalid object layer"; return -1; }
	}
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(bodies[i].id);
		W->h_ids[slot] = bodies[i].id;
		uint8_t layer = W->t_o2bp[bodies[i].object_layer];
		W->h_layer[slot] = layer;
		W->h_static[slot] = bodies[i].motion_type == B2J_MOTION_STATIC? 1 : 0;
		W->layer_bodies[layer].push_back(slot);
		W->layer_list_dirty[layer] = 1;
		W->layer_needs_build[layer] = 1;
		if (bodies[i].motion_type != B2J_MOTION_STATIC) W->layer_has_moving[layer] = 1;
		if (slot + 1 > W->num_slots) W->num_slots = slot + 1;
		if (bodies[i].active && bodies[i].motion_type != B2J_MOTION_STATIC) to_activate.push_back(bodies[i].id);
		if (bodies[i].flags & B2J_BODY_INVALIDATE_CACHE) W->h_cache_invalid.push_back(slot); // (a snapshot taken between InvalidateContactCache and the next update)
	}
	W->num_bodies += n;
	rt.stage_begin((size_t)n * sizeof(b2j_body_desc));
	b2j_body_desc *h_desc = nullptr;
	b2j_body_desc *tmp = rt.stage_alloc<b2j_body_desc>(n, &h_desc);
	memcpy(h_desc, bodies, (size_t)n * sizeof(b2j_body_desc));
	rt.stage_to_device(0, rt.stage_used);
	sync_dworld(W);
	KAddBodies k; k.w = W->d; k.descs = tmp;
	rt.launch(k, n);
	rt.sync();
	if (!to_activate.empty())
		return b2j_bodies_activate(W, to_activate.data(), (uint32_t)to_activate.size());
	return rt.check("b2j_bodies_add")? 0 : -1;
}

// ids must name live bodies of this world: a stale id (destroyed body, reused slot) or an index out of range is an error
static bool validate_ids(b2j_world *W, const uint32_t *ids, uint32_t n, const char *what)
{
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(ids[i]);
		if (slot >= W->d.max_bodies || W->h_ids[slot] != ids[i])
		{
			char buf[160];
			snprintf(buf, sizeof(buf), "%s: id 0x%08x (element %u) is not a body of this world", what, ids[i], i);
			last_error() = buf;
			return false;
		}
	}
	return true;
}

int b2j_bodies_remove(b2j_world *W, const uint32_t *ids, uint32_t n)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (ids == nullptr || !validate_ids(W, ids, n, "b2j_bodies_remove")) return -1;
	if (!W->h_joints.empty())
	{
		// the reference asks the caller to remove a body's constraints before the body (Constraint.h: "make sure the constraint is removed
		// before the bodies are"); here that is checked: a constraint left behind would solve against an empty slot
		if (W->h_mark.size() < W->d.max_bodies) W->h_mark.assign(W->d.max_bodies, 0);
		for (uint32_t i = 0; i < n; ++i) W->h_mark[slot_of(ids[i])] = 1;
		bool attached = false;
		for (const b2j_constraint_desc &c : W->h_joints)
			if (W->h_mark[slot_of(c.body1)] || W->h_mark[slot_of(c.body2)]) { attached = true; break; }
		for (uint32_t i = 0; i < n; ++i) W->h_mark[slot_of(ids[i])] = 0;
		if (attached) { last_error() = "b2j_bodies_remove: a body still has constraints attached (remove them first, b2j_constraints_remove)"; return -1; }
	}
	if (b2j_bodies_deactivate(W, ids, n) != 0) return -1;
	Runtime &rt = W->rt;
	// the slots are marked empty on the device too (id validation of the by-id kernels, device queries)
	rt.stage_begin((size_t)n * 4);
	uint32_t *h = nullptr;
	uint32_t *d_ids = rt.stage_alloc<uint32_t>(n, &h);
	memcpy(h, ids, (size_t)n * 4);
	rt.stage_to_device(0, rt.stage_used);
	sync_dworld(W);
	{ KClearBodies k; k.w = W->d; k.ids = d_ids; rt.launch(k, n); }
	rt.sync();
	uint32_t layers_touched = 0;
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(ids[i]);
		if (W->h_ids[slot] != ids[i]) continue; // listed twice
		W->h_ids[slot] = B2J_INVALID_ID;
		layers_touched |= 1u << W->h_layer[slot];
		W->num_bodies--;
	}
	// one pass per touched layer list (a per body erase is O(n * m))
	for (uint32_t l = 0; l < W->d.num_bp_layers; ++l)
		if (layers_touched & (1u << l))
		{
			std::vector<uint32_t> &list = W->layer_bodies[l];
			list.erase(std::remove_if(list.begin(), list.end(), [&](uint32_t slot) { return W->h_ids[slot] == B2J_INVALID_ID; }), list.end());
			W->layer_list_dirty[l] = 1;
			W->layer_needs_build[l] = 1;
		}
	return rt.check("b2j_bodies_remove")? 0 : -1;
}

int b2j_set_active_list(b2j_world *W, const uint32_t *ids, uint32_t n)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	sync_dworld(W);
	{ KClearActiveIndex k; k.w = W->d; rt.launch(k, W->num_active); }
	W->num_active = 0;
	if (n > 0)
	{
		if (ids == nullptr || !validate_ids(W, ids, n, "b2j_set_active_list")) return -1;
		rt.stage_begin((size_t)n * 4);
		uint32_t *h = nullptr;
		uint32_t *tmp = rt.stage_alloc<uint32_t>(n, &h);
		memcpy(h, ids, (size_t)n * 4);
		rt.stage_to_device(0, rt.stage_used);
		KSetActive k; k.w = W->d; k.ids = tmp; k.base = 0;
		rt.launch(k, n);
		rt.sync();
		W->num_active = n;
	}
	return rt.check("b2j_set_active_list")? 0 : -1;
}

// Body::ResetSleepTimer for bodies that are active already (BodyInterface::ActivateBodyInternal / ResetSleepTimer)
struct KResetSleepTimer
{
	DWorld w; const uint32_t *slots;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t b = slots[k];
		V3 points[3];
		sleep_test_points(w.shapes[w.info[b].shape], to_v3(w.position[b]), to_q4(w.rotation[b]), points);
		for (int i = 0; i < 3; ++i) w.sleep_spheres[b * 3 + i] = f4(points[i], 0.0f);
		w.sleep_timer[b] = 0.0f;
	}
};

static int bodies_activate(b2j_world *W, const uint32_t *ids, uint32_t n, bool activate, bool reset_active)
{
	// append the bodies that are not active yet, in argument order (BodyManager::ActivateBodies)
	Runtime &rt = W->rt;
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (ids == nullptr || !validate_ids(W, ids, n, "b2j_bodies_activate")) return -1;
	sync_dworld(W);
	// which of them sleep? (only the device knows: bodies fall asleep during steps) one staging round trip
	rt.stage_begin((size_t)n * 8);
	uint32_t *h_ids = nullptr, *h_idx = nullptr;
	uint32_t *d_ids = rt.stage_alloc<uint32_t>(n, &h_ids);
	memcpy(h_ids, ids, (size_t)n * 4);
	rt.stage_to_device(0, rt.stage_used);
	size_t out_begin = rt.stage_used;
	uint32_t *d_idx = rt.stage_alloc<uint32_t>(n, &h_idx);
	{ KGetState k; memset(&k, 0, sizeof(k)); k.w = W->d; k.ids = d_ids; k.active_index = d_idx; rt.launch(k, n); }
	rt.stage_to_host(out_begin, rt.stage_used);
	// first occurrence of every sleeping, non static body (a per slot mark instead of a search per id)
	if (W->h_mark.size() < W->d.max_bodies) W->h_mark.assign(W->d.max_bodies, 0);
	std::vector<uint32_t> add, reset;
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(ids[i]);
		if (W->h_mark[slot] || W->h_static[slot] != 0) continue;
		if (h_idx[i] == B2J_INACTIVE_INDEX) { if (activate) { W->h_mark[slot] = 1; add.push_back(ids[i]); } }
		else if (reset_active) { W->h_mark[slot] = 1; reset.push_back(slot); }
	}
	for (uint32_t id : add) W->h_mark[slot_of(id)] = 0;
	for (uint32_t slot : reset) W->h_mark[slot] = 0;
	if (!reset.empty())
	{
		uint32_t nr = (uint32_t)reset.size();
		rt.stage_begin((size_t)nr * 4);
		uint32_t *h_r = nullptr;
		uint32_t *d_r = rt.stage_alloc<uint32_t>(nr, &h_r);
		memcpy(h_r, reset.data(), (size_t)nr * 4);
		rt.stage_to_device(0, rt.stage_used);
		KResetSleepTimer k; k.w = W->d; k.slots = d_r;
		rt.launch(k, nr);
		rt.sync(); // the staging buffer is reused below
	}
	if (!add.empty())
	{
		uint32_t na = (uint32_t)add.size();
		rt.stage_begin((size_t)na * 8);
		uint32_t *h_a = nullptr, *h_s = nullptr;
		uint32_t *d_a = rt.stage_alloc<uint32_t>(na, &h_a), *d_s = rt.stage_alloc<uint32_t>(na, &h_s);
		for (uint32_t i = 0; i < na; ++i) { h_a[i] = add[i]; h_s[i] = slot_of(add[i]); }
		rt.stage_to_device(0, rt.stage_used);
		KSetActive k; k.w = W->d; k.ids = d_a; k.base = W->num_active;
		rt.launch(k, na);
		// Body::ResetSleepTimer
		KActivateWoken ka; ka.w = W->d; ka.woken_sorted = d_s; ka.base = W->num_active; ka.woken_flag = W->nc.woken_flag; ka.events = nullptr; ka.max_events = 0;
		rt.launch(ka, na);
		rt.sync(); // the staging buffer is reused by the next call
		W->num_active += na;
	}
	return rt.check("b2j_bodies_activate")? 0 : -1;
}

int b2j_bodies_activate(b2j_world *W, const uint32_t *ids, uint32_t n) { return bodies_activate(W, ids, n, true, false); }
int b2j_bodies_activate_or_reset_sleep_timer(b2j_world *W, const uint32_t *ids, uint32_t n) { return bodies_activate(W, ids, n, true, true); }
int b2j_bodies_reset_sleep_timer(b2j_world *W, const uint32_t *ids, uint32_t n) { return bodies_activate(W, ids, n, false, true); }

int b2j_bodies_deactivate(b2j_world *W, const uint32_t *ids, uint32_t n)
{
	// On the device: mark, reset velocities, stable compaction of the active list (the reference swaps the last body into the hole,
	// BodyManager.cpp:470-487; the order of the active list only decides which body of a pair queries the broadphase)
	Runtime &rt = W->rt;
	if (n == 0 || W->num_active == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (ids == nullptr) { last_error() = "b2j_bodies_deactivate: ids is null"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
		if (slot_of(ids[i]) >= W->d.max_bodies) { last_error() = "b2j_bodies_deactivate: body index out of range"; return -1; }
	sync_dworld(W);
	uint32_t na = W->num_active;
	rt.stage_begin((size_t)n * 4);
	uint32_t *h = nullptr;
	uint32_t *d_ids = rt.stage_alloc<uint32_t>(n, &h);
	memcpy(h, ids, (size_t)n * 4);
	rt.stage_to_device(0, rt.stage_used);
	{ KFillU32 k; k.dst = W->d_keep; k.value = 1; rt.launch(k, na); }
	{ KMarkDeactivate k; k.w = W->d; k.ids = d_ids; k.keep = W->d_keep; rt.launch(k, n); }
	{ KApiDeactivate k; k.w = W->d; k.keep = W->d_keep; rt.launch(k, na); }
	rt.exclusive_scan(W->d_keep, W->d_keep_scan, na);
	uint32_t *new_list = W->active_buf[W->active_cur ^ 1];
	W->stepped_list = nullptr; W->stepped_count = 0; // (that buffer held the list of the last step)
	{ KCompactActive k; k.w = W->d; k.keep = W->d_keep; k.keep_scan = W->d_keep_scan; k.new_active = new_list; rt.launch(k, na); }
	rt.memset_(W->d.counters, 0, sizeof(StepCounters));
	{ KFinishCompact k; k.w = W->d; k.keep = W->d_keep; k.keep_scan = W->d_keep_scan; k.n = na; rt.launch(k, 1); }
	W->active_cur ^= 1;
	sync_dworld(W);
	if (!read_counters(W)) return -1;
	W->num_active = W->h_counters.new_active_count;
	return rt.check("b2j_bodies_deactivate")? 0 : -1;
}

int b2j_bodies_get_state(b2j_world *W, const uint32_t *ids, uint32_t n, const b2j_body_state *out)
{
	B2J_DEVICE_GUARD(W);
	if (n == 0) return 0;
	Runtime &rt = W->rt;
	sync_dworld(W);
	if (ids == nullptr && (uint64_t)W->get_state_first + n > W->d.max_bodies) { last_error() = "n exceeds max_bodies"; return -1; }
	if (ids != nullptr && !validate_ids(W, ids, n, "b2j_bodies_get_state")) return -1;
	KGetState k; memset(&k, 0, sizeof(k)); k.w = W->d; k.first = W->get_state_first;
	// one persistent staging buffer (device + pinned mirror): ids up, one kernel, one copy down
	rt.stage_begin((size_t)n * (4 + 12 + 16 + 12 + 12 + 24 + 4 + 4));
	uint32_t *h_ids = nullptr; float *h_pos = nullptr, *h_rot = nullptr, *h_lin = nullptr, *h_ang = nullptr, *h_bounds = nullptr, *h_timer = nullptr; uint32_t *h_active = nullptr;
	if (ids != nullptr)
	{
		k.ids = rt.stage_alloc<uint32_t>(n, &h_ids);
		memcpy(h_ids, ids, (size_t)n * 4);
		rt.stage_to_device(0, rt.stage_used);
	}
	size_t out_begin = rt.stage_used;
	if (out->position) k.pos = rt.stage_alloc<float>((size_t)n * 3, &h_pos);
	if (out->rotation) k.rot = rt.stage_alloc<float>((size_t)n * 4, &h_rot);
	if (out->linear_velocity) k.lin = rt.stage_alloc<float>((size_t)n * 3, &h_lin);
	if (out->angular_velocity) k.ang = rt.stage_alloc<float>((size_t)n * 3, &h_ang);
	if (out->bounds) k.bounds = rt.stage_alloc<float>((size_t)n * 6, &h_bounds);
	if (out->active_index) k.active_index = rt.stage_alloc<uint32_t>(n, &h_active);
	if (out->sleep_timer) k.sleep_timer = rt.stage_alloc<float>(n, &h_timer);
	rt.launch(k, n);
	// page locked destination buffers receive the device arrays directly; pageable ones go through the pinned mirror
	struct Out { void *dst; const void *dev; const void *mirror; size_t bytes; };
	const Out outs[7] = { { out->position, k.pos, h_pos, (size_t)n * 12 }, { out->rotation, k.rot, h_rot, (size_t)n * 16 }, { out->linear_velocity, k.lin, h_lin, (size_t)n * 12 },
		{ out->angular_velocity, k.ang, h_ang, (size_t)n * 12 }, { out->bounds, k.bounds, h_bounds, (size_t)n * 24 }, { out->active_index, k.active_index, h_active, (size_t)n * 4 },
		{ out->sleep_timer, k.sleep_timer, h_timer, (size_t)n * 4 } };
	bool all_pinned = n >= 4096; // (small reads: the attribute queries cost more than the copy through the mirror)
	for (const Out &o : outs) if (o.dst != nullptr && all_pinned && !rt.is_pinned(o.dst)) all_pinned = false;
	if (all_pinned)
	{
		for (const Out &o : outs) if (o.dst != nullptr) rt.copy_to_host_async(o.dst, o.dev, o.bytes);
		rt.sync();
	}
	else
	{
		rt.stage_to_host(out_begin, rt.stage_used);
		for (const Out &o : outs) if (o.dst != nullptr) memcpy(o.dst, o.mirror, o.bytes);
	}
	return rt.check("b2j_bodies_get_state")? 0 : -1;
}

int b2j_host_buffer_register(void *ptr, size_t bytes)
{
#ifndef B2J_HOSTSIM
	if (ptr == nullptr || bytes == 0) return -1;
	cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) { cudaGetLastError(); last_error() = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return -1; }
	return 0;
#else
	(void)ptr; (void)bytes;
	return 0;
#endif
}

int b2j_host_buffer_unregister(void *ptr)
{
#ifndef B2J_HOSTSIM
	if (ptr == nullptr) return -1;
	cudaError_t e = cudaHostUnregister(ptr);
	if (e != cudaSuccess) { cudaGetLastError(); return -1; }
#else
	(void)ptr;
#endif
	return 0;
}

uint32_t b2j_bodies_get_stepped_state(b2j_world *W, uint32_t cap, uint32_t *ids, const b2j_body_state *out)
{
	B2J_DEVICE_GUARD(W);
	uint32_t total = W->stepped_list != nullptr? W->stepped_count : 0;
	uint32_t n = total < cap? total : cap;
	if (n == 0 || ids == nullptr || out == nullptr) return total;
	Runtime &rt = W->rt;
	sync_dworld(W);
	KGetStepped k; memset(&k, 0, sizeof(k)); k.w = W->d; k.slots = W->stepped_list;
	rt.stage_begin((size_t)n * (4 + 12 + 16 + 12 + 12 + 24 + 4 + 4));
	uint32_t *h_ids = nullptr, *h_active = nullptr; float *h_pos = nullptr, *h_rot = nullptr, *h_lin = nullptr, *h_ang = nullptr, *h_bounds = nullptr, *h_timer = nullptr;
	k.ids = rt.stage_alloc<uint32_t>(n, &h_ids);
	if (out->position) k.pos = rt.stage_alloc<float>((size_t)n * 3, &h_pos);
	if (out->rotation) k.rot = rt.stage_alloc<float>((size_t)n * 4, &h_rot);
	if (out->linear_velocity) k.lin = rt.stage_alloc<float>((size_t)n * 3, &h_lin);
	if (out->angular_velocity) k.ang = rt.stage_alloc<float>((size_t)n * 3, &h_ang);
	if (out->bounds) k.bounds = rt.stage_alloc<float>((size_t)n * 6, &h_bounds);
	if (out->active_index) k.active_index = rt.stage_alloc<uint32_t>(n, &h_active);
	if (out->sleep_timer) k.sleep_timer = rt.stage_alloc<float>(n, &h_timer);
	rt.launch(k, n);
	rt.stage_to_host(0, rt.stage_used);
	memcpy(ids, h_ids, (size_t)n * 4);
	if (h_pos) memcpy(out->position, h_pos, (size_t)n * 12);
	if (h_rot) memcpy(out->rotation, h_rot, (size_t)n * 16);
	if (h_lin) memcpy(out->linear_velocity, h_lin, (size_t)n * 12);
	if (h_ang) memcpy(out->angular_velocity, h_ang, (size_t)n * 12);
	if (h_bounds) memcpy(out->bounds, h_bounds, (size_t)n * 24);
	if (h_active) memcpy(out->active_index, h_active, (size_t)n * 4);
	if (h_timer) memcpy(out->sleep_timer, h_timer, (size_t)n * 4);
	if (!rt.check("b2j_bodies_get_stepped_state")) return 0;
	return total;
}

int b2j_bodies_set_state(b2j_world *W, const uint32_t *ids, uint32_t n, const b2j_body_state *in)
{
	B2J_DEVICE_GUARD(W);
	if (n == 0) return 0;
	Runtime &rt = W->rt;
	if (in == nullptr) { last_error() = "b2j_bodies_set_state: in is null"; return -1; }
	if (ids != nullptr? !validate_ids(W, ids, n, "b2j_bodies_set_state") : n > W->d.max_bodies) { if (ids == nullptr) last_error() = "n exceeds max_bodies"; return -1; }
	sync_dworld(W);
	upload_shapes(W);
	KSetState k; memset(&k, 0, sizeof(k)); k.w = W->d;
	rt.stage_begin((size_t)n * (4 + 12 + 16 + 12 + 12));
	uint32_t *h_ids = nullptr; float *h = nullptr;
	if (ids != nullptr) { k.ids = rt.stage_alloc<uint32_t>(n, &h_ids); memcpy(h_ids, ids, (size_t)n * 4); }
	if (in->position) { k.pos = rt.stage_alloc<float>((size_t)n * 3, &h); memcpy(h, in->position, (size_t)n * 12); }
	if (in->rotation) { k.rot = rt.stage_alloc<float>((size_t)n * 4, &h); memcpy(h, in->rotation, (size_t)n * 16); }
	if (in->linear_velocity) { k.lin = rt.stage_alloc<float>((size_t)n * 3, &h); memcpy(h, in->linear_velocity, (size_t)n * 12); }
	if (in->angular_velocity) { k.ang = rt.stage_alloc<float>((size_t)n * 3, &h); memcpy(h, in->angular_velocity, (size_t)n * 12); }
	rt.stage_to_device(0, rt.stage_used);
	rt.launch(k, n);
	rt.sync(); // the staging buffer is reused by the next call
	if (in->position || in->rotation)
		for (uint32_t l = 0; l < W->d.num_bp_layers; ++l) W->layer_needs_build[l] = 1;
	return rt.check("b2j_bodies_set_state")? 0 : -1;
}

int b2j_bodies_set_params(b2j_world *W, const uint32_t *ids, uint32_t n, const b2j_body_params *in)
{
	B2J_DEVICE_GUARD(W);
	if (n == 0) return 0;
	if (ids == nullptr || in == nullptr) { last_error() = "b2j_bodies_set_params: ids and in are required"; return -1; }
	if (!validate_ids(W, ids, n, "b2j_bodies_set_params")) return -1;
	Runtime &rt = W->rt;
	sync_dworld(W);
	KSetParams k; memset(&k, 0, sizeof(k)); k.w = W->d;
	rt.stage_begin((size_t)n * 4 * 8);
	uint32_t *h_ids = nullptr; float *h = nullptr;
	k.ids = rt.stage_alloc<uint32_t>(n, &h_ids); memcpy(h_ids, ids, (size_t)n * 4);
	const float *src[7] = { in->friction, in->restitution, in->gravity_factor, in->linear_damping, in->angular_damping, in->max_linear_velocity, in->max_angular_velocity };
	const float **dst[7] = { &k.friction, &k.restitution, &k.gravity_factor, &k.linear_damping, &k.angular_damping, &k.max_linear_velocity, &k.max_angular_velocity };
	for (int f = 0; f < 7; ++f)
		if (src[f] != nullptr) { *dst[f] = rt.stage_alloc<float>(n, &h); memcpy(h, src[f], (size_t)n * 4); }
	rt.stage_to_device(0, rt.stage_used);
	rt.launch(k, n);
	rt.sync(); // the staging buffer is reused by the next call
	return rt.check("b2j_bodies_set_params")? 0 : -1;
}

int b2j_bodies_add_force_torque(b2j_world *W, const uint32_t *ids, uint32_t n, const float *force, const float *torque)
{
	B2J_DEVICE_GUARD(W);
	if (n == 0) return 0;
	Runtime &rt = W->rt;
	if (ids != nullptr? !validate_ids(W, ids, n, "b2j_bodies_add_force_torque") : n > W->d.max_bodies) { if (ids == nullptr) last_error() = "n exceeds max_bodies"; return -1; }
	sync_dworld(W);
	KAddForceTorque k; k.w = W->d;
	rt.stage_begin((size_t)n * (4 + 12 + 12));
	uint32_t *h_ids = nullptr; float *h = nullptr;
	k.ids = nullptr;
	if (ids != nullptr) { k.ids = rt.stage_alloc<uint32_t>(n, &h_ids); memcpy(h_ids, ids, (size_t)n * 4); } // NULL: slots 0..n-1
	k.force = nullptr; k.torque = nullptr;
	// page locked source arrays are copied to the device directly; pageable ones through the pinned mirror
	bool direct = n >= 4096 && (force == nullptr || rt.is_pinned(force)) && (torque == nullptr || rt.is_pinned(torque));
	size_t mirror_end = rt.stage_used;
	float *d_force = nullptr, *d_torque = nullptr;
	if (force) { d_force = rt.stage_alloc<float>((size_t)n * 3, &h); if (!direct) { memcpy(h, force, (size_t)n * 12); mirror_end = rt.stage_used; } }
	if (torque) { d_torque = rt.stage_alloc<float>((size_t)n * 3, &h); if (!direct) { memcpy(h, torque, (size_t)n * 12); mirror_end = rt.stage_used; } }
	k.force = d_force; k.torque = d_torque;
	rt.stage_to_device(0, mirror_end);
	if (direct)
	{
		if (force) rt.copy_to_device_async(d_force, force, (size_t)n * 12);
		if (torque) rt.copy_to_device_async(d_torque, torque, (size_t)n * 12);
	}
	rt.launch(k, n);
	rt.sync();
	return rt.check("b2j_bodies_add_force_torque")? 0 : -1;
}

int b2j_bodies_set_info(b2j_world *W, const uint32_t *ids, uint32_t n, const b2j_body_info_update *in)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (ids == nullptr || in == nullptr) { last_error() = "b2j_bodies_set_info: ids and in are required"; return -1; }
	if (!validate_ids(W, ids, n, "b2j_bodies_set_info")) return -1;
	if (in->motion_type != nullptr && in->inv_mass == nullptr) { last_error() = "b2j_bodies_set_info: motion_type needs inv_mass"; return -1; }
	if ((in->inv_inertia_diag != nullptr) != (in->inertia_rotation != nullptr)) { last_error() = "b2j_bodies_set_info: inv_inertia_diag and inertia_rotation go together"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
	{
		if (in->motion_type != nullptr && in->motion_type[i] > B2J_MOTION_DYNAMIC) { last_error() = "b2j_bodies_set_info: invalid motion type"; return -1; }
		if (in->object_layer != nullptr && in->object_layer[i] >= W->d.num_object_layers) { last_error() = "b2j_bodies_set_info: invalid object layer"; return -1; }
		if (in->shape != nullptr && (in->shape[i] < 0 || in->shape[i] >= (int32_t)W->h_shapes.size())) { last_error() = "b2j_bodies_set_info: invalid shape id"; return -1; }
	}
	Runtime &rt = W->rt;
	upload_shapes(W);
	// bodies that become static leave the active list first (BodyInterface::SetMotionType)
	if (in->motion_type != nullptr)
	{
		std::vector<uint32_t> to_static;
		for (uint32_t i = 0; i < n; ++i)
			if (in->motion_type[i] == B2J_MOTION_STATIC && !W->h_static[slot_of(ids[i])]) to_static.push_back(ids[i]);
		if (!to_static.empty() && b2j_bodies_deactivate(W, to_static.data(), (uint32_t)to_static.size()) != 0) return -1;
	}
	sync_dworld(W);
	KSetInfo k; memset(&k, 0, sizeof(k)); k.w = W->d; k.invalidate = in->invalidate_contact_cache;
	rt.stage_begin((size_t)n * (4 + 1 + 4 + 2 + 4 + 12 + 16 + 2 + 2) + 1024);
	uint32_t *h_ids = nullptr;
	k.ids = rt.stage_alloc<uint32_t>(n, &h_ids); memcpy(h_ids, ids, (size_t)n * 4);
	if (in->motion_type) { uint8_t *h; k.motion_type = rt.stage_alloc<uint8_t>(n, &h); memcpy(h, in->motion_type, n); }
	if (in->inv_mass) { float *h; k.inv_mass = rt.stage_alloc<float>(n, &h); memcpy(h, in->inv_mass, (size_t)n * 4); }
	if (in->object_layer) { uint16_t *h; k.object_layer = rt.stage_alloc<uint16_t>(n, &h); memcpy(h, in->object_layer, (size_t)n * 2); }
	if (in->shape) { int32_t *h; k.shape = rt.stage_alloc<int32_t>(n, &h); memcpy(h, in->shape, (size_t)n * 4); }
	if (in->inv_inertia_diag) { float *h; k.inv_inertia_diag = rt.stage_alloc<float>((size_t)n * 3, &h); memcpy(h, in->inv_inertia_diag, (size_t)n * 12); }
	if (in->inertia_rotation) { float *h; k.inertia_rotation = rt.stage_alloc<float>((size_t)n * 4, &h); memcpy(h, in->inertia_rotation, (size_t)n * 16); }
	if (in->flags_set) { uint16_t *h; k.flags_set = rt.stage_alloc<uint16_t>(n, &h); memcpy(h, in->flags_set, (size_t)n * 2); }
	if (in->flags_clear) { uint16_t *h; k.flags_clear = rt.stage_alloc<uint16_t>(n, &h); memcpy(h, in->flags_clear, (size_t)n * 2); }
	rt.stage_to_device(0, rt.stage_used);
	rt.launch(k, n);
	rt.sync();
	// host mirrors: static flags, broadphase layer lists, the list of bodies whose contact cache flag is set
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t slot = slot_of(ids[i]);
		if (in->motion_type != nullptr)
		{
			W->h_static[slot] = in->motion_type[i] == B2J_MOTION_STATIC? 1 : 0;
			if (in->motion_type[i] != B2J_MOTION_STATIC) W->layer_has_moving[W->h_layer[slot]] = 1;
		}
		if (in->object_layer != nullptr)
		{
			uint8_t layer = W->t_o2bp[in->object_layer[i]];
			if (layer != W->h_layer[slot])
			{
				std::vector<uint32_t> &from = W->layer_bodies[W->h_layer[slot]];
				from.erase(std::find(from.begin(), from.end(), slot));
				W->layer_list_dirty[W->h_layer[slot]] = 1; W->layer_needs_build[W->h_layer[slot]] = 1;
				W->layer_bodies[layer].push_back(slot);
				W->layer_list_dirty[layer] = 1; W->layer_needs_build[layer] = 1;
				if (!W->h_static[slot]) W->layer_has_moving[layer] = 1;
				W->h_layer[slot] = layer;
			}
		}
		if (in->shape != nullptr || in->invalidate_contact_cache) W->h_cache_invalid.push_back(slot);
		if (in->shape != nullptr) W->layer_needs_build[W->h_layer[slot]] = 1;
	}
	return rt.check("b2j_bodies_set_info")? 0 : -1;
}

uint32_t b2j_num_bodies(const b2j_world *W) { return W->num_bodies; }
uint32_t b2j_num_active_bodies(const b2j_world *W) { return W->num_active; }

uint32_t b2j_get_active_bodies(b2j_world *W, uint32_t *ids, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	uint32_t n = W->num_active < cap? W->num_active : cap;
	if (n > 0)
	{
		std::vector<uint32_t> slots(n);
		sync_dworld(W);
		W->rt.download(slots.data(), W->d.active, n);
		for (uint32_t i = 0; i < n; ++i) ids[i] = W->h_ids[slots[i]];
	}
	return W->num_active;
}

// ---- non contact constraints (b2j_joints.h) ---------------------------------------------------------------------------------------
int b2j_constraints_add(b2j_world *W, const b2j_constraint_desc *constraints, uint32_t n)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (constraints == nullptr) { last_error() = "b2j_constraints_add: constraints is required"; return -1; }
	if (W->num_worlds != 1) { last_error() = "b2j_constraints_add: batched worlds do not take constraints"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
	{
		const b2j_constraint_desc &c = constraints[i];
		uint32_t ids[2] = { c.body1, c.body2 };
		if (!validate_ids(W, ids, 2, "b2j_constraints_add")) return -1;
		if (c.type != B2J_CONSTRAINT_POINT && c.type != B2J_CONSTRAINT_DISTANCE && c.type != B2J_CONSTRAINT_HINGE && c.type != B2J_CONSTRAINT_FIXED) { last_error() = "b2j_constraints_add: unknown constraint type"; return -1; }
		if (c.type == B2J_CONSTRAINT_HINGE && !(c.limits_min <= 0.0f && c.limits_max >= 0.0f && c.max_friction_torque >= 0.0f)) { last_error() = "b2j_constraints_add: hinge limits_min <= 0 <= limits_max and max_friction_torque >= 0 expected"; return -1; }
		if (c.body1 == c.body2) { last_error() = "b2j_constraints_add: a constraint connects two different bodies"; return -1; }
		if (c.type == B2J_CONSTRAINT_DISTANCE && !(c.min_distance >= 0.0f && c.max_distance >= c.min_distance)) { last_error() = "b2j_constraints_add: 0 <= min_distance <= max_distance expected"; return -1; }
	}
	uint32_t first = (uint32_t)W->h_joints.size();
	if (!joints_reserve(W, first + n)) return -1;
	Runtime &rt = W->rt;
	std::vector<JointState> init(n);
	memset(init.data(), 0, n * sizeof(JointState));
	for (uint32_t i = 0; i < n; ++i)
	{
		init[i].normal = f4(0.0f, 1.0f, 0.0f, 0.0f); // DistanceConstraint: mWorldSpaceNormal = Vec3::sAxisY()
		W->h_joints.push_back(constraints[i]);
	}
	rt.upload(W->jc.state + first, init.data(), n);
	rt.sync();
	W->joints_dirty = true;
	return rt.check("b2j_constraints_add")? 0 : -1;
}

int b2j_constraints_remove(b2j_world *W, const uint32_t *indices, uint32_t n)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (W->num_worlds != 1) { last_error() = "b2j_constraints_remove: the constraint list of batched worlds is fixed"; return -1; }
	Runtime &rt = W->rt;
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t last = (uint32_t)W->h_joints.size();
		if (indices == nullptr || indices[i] >= last) { last_error() = "b2j_constraints_remove: invalid constraint index"; return -1; }
		--last;
		if (indices[i] < last)
		{
			// ConstraintManager::Remove: the last constraint takes the freed index
			W->h_joints[indices[i]] = W->h_joints[last];
			rt.copy(W->jc.state + indices[i], W->jc.state + last, 1);
		}
		W->h_joints.pop_back();
	}
	rt.sync();
	W->joints_dirty = true;
	W->jc.num_joints = (uint32_t)W->h_joints.size();
	return rt.check("b2j_constraints_remove")? 0 : -1;
}

uint32_t b2j_num_constraints(const b2j_world *W) { return (uint32_t)W->h_joints.size(); }

int b2j_constraints_set_enabled(b2j_world *W, const uint32_t *indices, uint32_t n, const uint8_t *enabled)
{
	for (uint32_t i = 0; i < n; ++i)
	{
		if (indices[i] >= W->h_joints.size()) { last_error() = "b2j_constraints_set_enabled: invalid constraint index"; return -1; }
		W->h_joints[indices[i]].enabled = enabled[i];
	}
	W->joints_dirty = true;
	return 0;
}

int b2j_constraints_get_state(b2j_world *W, uint32_t first, uint32_t n, b2j_constraint_state *out)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (out == nullptr || (size_t)first + n > W->h_joints.size()) { last_error() = "b2j_constraints_get_state: invalid range"; return -1; }
	std::vector<JointState> st(n);
	W->rt.download(st.data(), W->jc.state + first, n);
	for (uint32_t i = 0; i < n; ++i)
	{
		out[i].total_lambda[0] = st[i].lambda.x; out[i].total_lambda[1] = st[i].lambda.y; out[i].total_lambda[2] = st[i].lambda.z;
		out[i].world_space_normal[0] = st[i].normal.x; out[i].world_space_normal[1] = st[i].normal.y; out[i].world_space_normal[2] = st[i].normal.z;
		bool fixed = W->h_joints[first + i].type == B2J_CONSTRAINT_FIXED;
		out[i].total_lambda_rotation[0] = st[i].lambda2.x; out[i].total_lambda_rotation[1] = st[i].lambda2.y; out[i].total_lambda_rotation[2] = fixed? st[i].lambda2.z : 0.0f;
		out[i].total_lambda_limits = fixed? 0.0f : st[i].lambda2.z; out[i].total_lambda_motor = st[i].lambda2.w;
	}
	return W->rt.check("b2j_constraints_get_state")? 0 : -1;
}

int b2j_constraints_set_state(b2j_world *W, uint32_t first, uint32_t n, const b2j_constraint_state *in)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (in == nullptr || (size_t)first + n > W->h_joints.size()) { last_error() = "b2j_constraints_set_state: invalid range"; return -1; }
	std::vector<JointState> st(n);
	W->rt.download(st.data(), W->jc.state + first, n);
	for (uint32_t i = 0; i < n; ++i)
	{
		st[i].lambda = f4(in[i].total_lambda[0], in[i].total_lambda[1], in[i].total_lambda[2], 0.0f);
		st[i].normal = f4(in[i].world_space_normal[0], in[i].world_space_normal[1], in[i].world_space_normal[2], 0.0f);
		bool fixed = W->h_joints[first + i].type == B2J_CONSTRAINT_FIXED;
		st[i].lambda2 = f4(in[i].total_lambda_rotation[0], in[i].total_lambda_rotation[1], fixed? in[i].total_lambda_rotation[2] : in[i].total_lambda_limits, in[i].total_lambda_motor);
	}
	W->rt.upload(W->jc.state + first, st.data(), n);
	W->rt.sync();
	return W->rt.check("b2j_constraints_set_state")? 0 : -1;
}

int b2j_contact_cache_import(b2j_world *W, const b2j_cached_body_pair *pairs, uint32_t num_pairs, const b2j_cached_manifold *manifolds, uint32_t num_manifolds)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	if (num_pairs > W->d.max_body_pairs || num_manifolds > W->d.max_constraints) { last_error() = "contact cache snapshot exceeds the world limits"; return -1; }
	for (uint32_t i = 0; i < num_pairs; ++i)
		if ((uint64_t)pairs[i].first_manifold + pairs[i].num_manifolds > num_manifolds || slot_of(pairs[i].body1) >= W->d.max_bodies || slot_of(pairs[i].body2) >= W->d.max_bodies)
		{
			last_error() = "contact cache snapshot is inconsistent (manifold range or body index out of bounds)";
			return -1;
		}
	int ri = W->write_idx ^ 1;
	clear_cache(W, ri);
	sync_dworld(W);
	if (num_pairs > 0)
	{
		b2j_cached_body_pair *dp = rt.alloc<b2j_cached_body_pair>(num_pairs, false);
		b2j_cached_manifold *dm = rt.alloc<b2j_cached_manifold>(num_manifolds, false);
		rt.upload(dp, pairs, num_pairs);
		rt.upload(dm, manifolds, num_manifolds);
		KImportCache k; k.w = W->d; k.pairs = dp; k.manifolds = dm;
		rt.launch(k, num_pairs);
		rt.upload(W->cache[ri].num_pairs, &num_pairs, 1);
		rt.upload(W->cache[ri].num_manifolds, &num_manifolds, 1);
		rt.sync();
		rt.free_(dp); rt.free_(dm);
	}
	W->cache_num_pairs[ri] = num_pairs;
	W->cache_num_manifolds[ri] = num_manifolds;
	return rt.check("b2j_contact_cache_import")? 0 : -1;
}

int b2j_contact_cache_export(b2j_world *W, b2j_cached_body_pair *pairs, uint32_t pairs_cap, uint32_t *num_pairs, b2j_cached_manifold *manifolds, uint32_t manifolds_cap, uint32_t *num_manifolds)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	int ri = W->write_idx ^ 1;
	uint32_t np = W->cache_num_pairs[ri], nm = W->cache_num_manifolds[ri];
	if (num_pairs) *num_pairs = np;
	if (num_manifolds) *num_manifolds = nm;
	if (np == 0 || pairs == nullptr || manifolds == nullptr) return 0;
	sync_dworld(W);
	b2j_cached_body_pair *dp = rt.alloc<b2j_cached_body_pair>(np);
	b2j_cached_manifold *dm = rt.alloc<b2j_cached_manifold>(nm);
	KExportCache k; k.w = W->d; k.pairs = dp; k.manifolds = dm;
	rt.launch(k, np);
	std::vector<b2j_cached_body_pair> hp(np);
	std::vector<b2j_cached_manifold> hm(nm);
	rt.download(hp.data(), dp, np);
	rt.download(hm.data(), dm, nm);
	rt.free_(dp); rt.free_(dm);
	// sorted by (body1, body2); manifolds of a pair sorted by (sub1, sub2), re-packed contiguously
	std::vector<uint32_t> order(np);
	std::iota(order.begin(), order.end(), 0u);
	std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hp[a].body1 != hp[b].body1? hp[a].body1 < hp[b].body1 : hp[a].body2 < hp[b].body2; });
	uint32_t mo = 0;
	for (uint32_t i = 0; i < np; ++i)
	{
		b2j_cached_body_pair p = hp[order[i]];
		std::vector<b2j_cached_manifold> ms(hm.begin() + p.first_manifold, hm.begin() + p.first_manifold + p.num_manifolds);
		std::sort(ms.begin(), ms.end(), [](const b2j_cached_manifold &a, const b2j_cached_manifold &b) { return a.sub_shape1 != b.sub_shape1? a.sub_shape1 < b.sub_shape1 : a.sub_shape2 < b.sub_shape2; });
		p.first_manifold = mo;
		for (const b2j_cached_manifold &m : ms) { if (mo < manifolds_cap) manifolds[mo] = m; ++mo; }
		if (i < pairs_cap) pairs[i] = p;
	}
	return rt.check("b2j_contact_cache_export")? 0 : -1;
}

int b2j_were_bodies_in_contact(b2j_world *W, uint32_t id1, uint32_t id2)
{
	// one probe of the device pair table of the read cache (the reference probes its lock free map, ContactConstraintManager.cpp:1531-1543)
	B2J_DEVICE_GUARD(W);
	if (slot_of(id1) >= W->d.max_bodies || slot_of(id2) >= W->d.max_bodies) return 0;
	Runtime &rt = W->rt;
	sync_dworld(W);
	uint32_t a = id1 < id2? id1 : id2, b = id1 < id2? id2 : id1, result = 0;
	uint32_t *out = reinterpret_cast<uint32_t *>(W->d_energy); // 4 byte scratch word
	{ KWereInContact k; k.w = W->d; k.id1 = a; k.id2 = b; k.out = out; rt.launch(k, 1); }
	rt.download(&result, out, 1);
	if (!rt.check("b2j_were_bodies_in_contact")) return -1;
	return (int)result;
}

// trees of the current body bounds (the step rebuilds them lazily: bodies may have moved / been added since)
static bool ensure_trees(b2j_world *W)
{
	upload_shapes(W);
	sync_dworld(W);
	for (uint32_t l = 0; l < W->d.num_bp_layers; ++l)
		if (W->layer_needs_build[l])
		{
			if (!build_tree(W, l)) return false;
			W->layer_needs_build[l] = 0;
		}
	return true;
}

static int cast_rays(b2j_world *W, const uint32_t *ray_world, uint32_t first_world, const b2j_ray *rays, uint32_t n, uint32_t object_layer, b2j_ray_hit *hits)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (rays == nullptr || hits == nullptr) { last_error() = "b2j_query_cast_rays: rays and hits are required"; return -1; }
	if (object_layer != 0xffffffffu && object_layer >= W->d.num_object_layers) { last_error() = "b2j_query_cast_rays: invalid object layer"; return -1; }
	Runtime &rt = W->rt;
	if (!ensure_trees(W)) return -1;
	rt.stage_begin((size_t)n * (sizeof(b2j_ray) + sizeof(b2j_ray_hit) + 4));
	b2j_ray *h_rays = nullptr; b2j_ray_hit *h_hits = nullptr; uint32_t *h_world = nullptr;
	KCastRays k; k.w = W->d; for (int l = 0; l < 8; ++l) k.trees[l] = W->trees[l];
	k.rays = rt.stage_alloc<b2j_ray>(n, &h_rays);
	memcpy(h_rays, rays, (size_t)n * sizeof(b2j_ray));
	k.ray_world = nullptr; k.first_world = first_world; k.num_worlds = W->num_worlds;
	if (ray_world != nullptr) { k.ray_world = rt.stage_alloc<uint32_t>(n, &h_world); memcpy(h_world, ray_world, (size_t)n * 4); }
	rt.stage_to_device(0, rt.stage_used);
	size_t out_begin = rt.stage_used;
	k.hits = rt.stage_alloc<b2j_ray_hit>(n, &h_hits);
	k.object_layer = object_layer;
	if (ray_world != nullptr) rt.memset_(k.hits, 0xff, (size_t)n * sizeof(b2j_ray_hit)); // rays of other groups' worlds stay "not mine" (body = 0xffffffff, fraction = NaN)
	rt.launch(k, n);
	rt.stage_to_host(out_begin, rt.stage_used);
	if (ray_world == nullptr)
		memcpy(hits, h_hits, (size_t)n * sizeof(b2j_ray_hit));
	else
		for (uint32_t i = 0; i < n; ++i)
			if (ray_world[i] >= first_world && ray_world[i] < first_world + W->num_worlds) hits[i] = h_hits[i];
	return rt.check("b2j_query_cast_rays")? 0 : -1;
}

int b2j_query_cast_rays(b2j_world *W, const b2j_ray *rays, uint32_t n, uint32_t object_layer, b2j_ray_hit *hits)
{
	return cast_rays(W, nullptr, 0, rays, n, object_layer, hits);
}

static int collide_volume(b2j_world *W, const char *what, int mode, uint32_t floats, const float *boxes, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (boxes == nullptr || counts == nullptr || (max_hits > 0 && ids == nullptr)) { last_error() = std::string(what) + ": the query volumes, counts and ids are required"; return -1; }
	if (object_layer != 0xffffffffu && object_layer >= W->d.num_object_layers) { last_error() = std::string(what) + ": invalid object layer"; return -1; }
	Runtime &rt = W->rt;
	if (!ensure_trees(W)) return -1;
	rt.stage_begin((size_t)n * (4 * floats + 4 + (size_t)max_hits * 4));
	float *h_boxes = nullptr; uint32_t *h_counts = nullptr, *h_ids = nullptr;
	KCollideAABox k; k.w = W->d; for (int l = 0; l < 8; ++l) k.trees[l] = W->trees[l];
	k.mode = mode;
	k.boxes = rt.stage_alloc<float>((size_t)n * floats, &h_boxes);
	memcpy(h_boxes, boxes, (size_t)n * floats * 4);
	rt.stage_to_device(0, rt.stage_used);
	size_t out_begin = rt.stage_used;
	k.counts = rt.stage_alloc<uint32_t>(n, &h_counts);
	k.ids = rt.stage_alloc<uint32_t>((size_t)n * max_hits, &h_ids);
	k.max_hits = max_hits; k.box_world = nullptr; k.first_world = 0; k.num_worlds = 1; k.object_layer = object_layer;
	rt.launch(k, n);
	rt.stage_to_host(out_begin, rt.stage_used);
	memcpy(counts, h_counts, (size_t)n * 4);
	if (max_hits > 0) memcpy(ids, h_ids, (size_t)n * max_hits * 4);
	return rt.check(what)? 0 : -1;
}

int b2j_query_collide_aabox(b2j_world *W, const float *boxes, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids)
{
	return collide_volume(W, "b2j_query_collide_aabox", 0, 6, boxes, n, object_layer, max_hits, counts, ids);
}

int b2j_query_collide_sphere(b2j_world *W, const float *spheres, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids)
{
	return collide_volume(W, "b2j_query_collide_sphere", 1, 4, spheres, n, object_layer, max_hits, counts, ids);
}

int b2j_query_collide_point(b2j_world *W, const float *points, uint32_t n, uint32_t object_layer, uint32_t max_hits, uint32_t *counts, uint32_t *ids)
{
	return collide_volume(W, "b2j_query_collide_point", 2, 3, points, n, object_layer, max_hits, counts, ids);
}

int b2j_query_collide_shape(b2j_world *W, const b2j_shape_query *queries, uint32_t n, float max_separation_distance, uint32_t object_layer,
	uint32_t max_hits, uint32_t *counts, b2j_collide_shape_hit *hits)
{
	if (n == 0) return 0;
	B2J_DEVICE_GUARD(W);
	if (queries == nullptr || counts == nullptr || (max_hits > 0 && hits == nullptr)) { last_error() = "b2j_query_collide_shape: queries, counts and hits are required"; return -1; }
	if (object_layer != 0xffffffffu && object_layer >= W->d.num_object_layers) { last_error() = "b2j_query_collide_shape: invalid object layer"; return -1; }
	if (W->num_worlds != 1) { last_error() = "b2j_query_collide_shape: single worlds only"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
	{
		int32_t sh = queries[i].shape;
		if (sh < 0 || (size_t)sh >= W->h_shapes.size()) { last_error() = "b2j_query_collide_shape: invalid shape id"; return -1; }
		uint32_t kind = W->h_shapes[sh].kind;
		if (kind == B2J_SHAPE_MESH || kind == B2J_SHAPE_COMPOUND) { last_error() = "b2j_query_collide_shape: the query shape must be a convex shape"; return -1; }
	}
	Runtime &rt = W->rt;
	if (!ensure_trees(W)) return -1;
	if (W->d_query_scratch == nullptr) W->d_query_scratch = rt.alloc<MeshScratch>(1, false);
	if (W->d_query_scratch == nullptr) { last_error() = "b2j_query_collide_shape: out of device memory"; return -1; }
	rt.stage_begin((size_t)n * (sizeof(b2j_shape_query) + 4 + (size_t)max_hits * sizeof(b2j_collide_shape_hit)) + 64);
	b2j_shape_query *h_queries = nullptr; uint32_t *h_counts = nullptr, *h_n = nullptr; b2j_collide_shape_hit *h_hits = nullptr;
	KCollideShape k; k.w = W->d; for (int l = 0; l < 8; ++l) k.trees[l] = W->trees[l];
	k.queries = rt.stage_alloc<b2j_shape_query>(n, &h_queries);
	memcpy(h_queries, queries, (size_t)n * sizeof(b2j_shape_query));
	const uint32_t *d_n = rt.stage_alloc<uint32_t>(4, &h_n);
	h_n[0] = n;
	rt.stage_to_device(0, rt.stage_used);
	size_t out_begin = rt.stage_used;
	k.counts = rt.stage_alloc<uint32_t>(n, &h_counts);
	k.hits = rt.stage_alloc<b2j_collide_shape_hit>((size_t)n * max_hits, &h_hits);
	k.max_hits = max_hits; k.max_separation_distance = max_separation_distance; k.object_layer = object_layer;
	k.mesh_scratch = W->d_query_scratch;
	rt.launch_warp_smem<KCollideShape, EpaStorageFull>(k, d_n, n, W->nc.num_scratch);
	rt.stage_to_host(out_begin, rt.stage_used);
	memcpy(counts, h_counts, (size_t)n * 4);
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t c = h_counts[i] < max_hits? h_counts[i] : max_hits;
		memcpy(hits + (size_t)i * max_hits, h_hits + (size_t)i * max_hits, (size_t)c * sizeof(b2j_collide_shape_hit));
	}
	return rt.check("b2j_query_collide_shape")? 0 : -1;
}

static void snapshot_free(WorldSnapshot &ws)
{
	if (ws.owner == nullptr) return;
	B2J_DEVICE_GUARD(ws.owner);
	Runtime &rt = ws.owner->rt;
	rt.sync();
	rt.free_(ws.info); rt.free_(ws.params); rt.free_(ws.pose); rt.free_(ws.velocity); rt.free_(ws.force_torque); rt.free_(ws.inertia); rt.free_(ws.bounds);
	rt.free_(ws.sleep_spheres); rt.free_(ws.sleep_timer); rt.free_(ws.active_index); rt.free_(ws.active); rt.free_(ws.pairs); rt.free_(ws.manifolds); rt.free_(ws.joint_state);
	ws.owner = nullptr;
}

static bool snapshot_take(b2j_world *W, WorldSnapshot &ws)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	sync_dworld(W);
	const DWorld &d = W->d;
	ws.owner = W;
	uint32_t n = W->num_slots, na = W->num_active;
	int ri = W->write_idx ^ 1;
	uint32_t np = W->cache_num_pairs[ri], nm = W->cache_num_manifolds[ri];
	ws.num_slots = n; ws.num_active = na; ws.num_bodies = W->num_bodies; ws.num_pairs = np; ws.num_manifolds = nm;
	ws.prev_dt = W->prev_dt; ws.gravity = d.gravity;
	bool ok = true;
	auto keep = [&](auto *&dst, const auto *src, size_t count) {
		typedef typename std::remove_reference<decltype(*dst)>::type T;
		dst = rt.alloc<T>(count, false);
		if (dst == nullptr) { ok = false; return; }
		rt.copy(dst, (const T *)src, count);
		ws.bytes += count * sizeof(T);
	};
	keep(ws.info, d.info, n); keep(ws.params, d.params, n);
	keep(ws.pose, d.position.base, 2 * (size_t)n); keep(ws.velocity, d.linear_velocity.base, 2 * (size_t)n); keep(ws.force_torque, d.force.base, 2 * (size_t)n);
	keep(ws.inertia, d.inv_inertia_diag.base, 2 * (size_t)n); keep(ws.bounds, d.bounds_min.base, 2 * (size_t)n);
	keep(ws.sleep_spheres, d.sleep_spheres, 3 * (size_t)n); keep(ws.sleep_timer, d.sleep_timer, n); keep(ws.active_index, d.active_index, n);
	keep(ws.active, d.active, na);
	ws.h_joints = W->h_joints;
	if (!W->h_joints.empty()) keep(ws.joint_state, W->jc.state, W->h_joints.size() * W->num_worlds);
	keep(ws.pairs, d.read_cache.pairs, np); keep(ws.manifolds, d.read_cache.manifolds, nm);
	ws.h_ids.assign(W->h_ids.begin(), W->h_ids.begin() + n); ws.h_layer.assign(W->h_layer.begin(), W->h_layer.begin() + n); ws.h_static.assign(W->h_static.begin(), W->h_static.begin() + n);
	ws.layer_bodies = W->layer_bodies; ws.layer_has_moving = W->layer_has_moving;
	rt.sync();
	if (!ok || !rt.check("b2j_world_save_state")) { snapshot_free(ws); last_error() = "b2j_world_save_state: out of device memory"; return false; }
	return true;
}

static bool snapshot_restore(b2j_world *W, const WorldSnapshot &ws)
{
	if (ws.owner != W) { last_error() = "b2j_world_restore_state: the snapshot was taken from another world"; return false; }
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	sync_dworld(W);
	DWorld &d = W->d;
	uint32_t n = ws.num_slots;
	// bodies added since the save live in slots the snapshot may not cover: empty them
	if (W->num_slots > n)
	{
		uint32_t extra = W->num_slots - n;
		{ KFillU32 k; k.dst = d.active_index + n; k.value = B2J_INACTIVE_INDEX; rt.launch(k, extra); }
		std::vector<BodyInfo> empty(extra);
		memset(empty.data(), 0, sizeof(BodyInfo) * extra);
		for (BodyInfo &i : empty) i.id = B2J_INVALID_ID;
		rt.upload(d.info + n, empty.data(), extra);
		for (uint32_t slot = n; slot < W->num_slots; ++slot) W->h_ids[slot] = B2J_INVALID_ID;
	}
	rt.copy(d.info, ws.info, n); rt.copy(d.params, ws.params, n);
	rt.copy(d.position.base, ws.pose, 2 * (size_t)n); rt.copy(d.linear_velocity.base, ws.velocity, 2 * (size_t)n); rt.copy(d.force.base, ws.force_torque, 2 * (size_t)n);
	rt.copy(d.inv_inertia_diag.base, ws.inertia, 2 * (size_t)n); rt.copy(d.bounds_min.base, ws.bounds, 2 * (size_t)n);
	rt.copy(d.sleep_spheres, ws.sleep_spheres, 3 * (size_t)n); rt.copy(d.sleep_timer, ws.sleep_timer, n); rt.copy(d.active_index, ws.active_index, n);
	rt.copy(d.active, ws.active, ws.num_active);
	if (!ws.h_joints.empty() && !joints_reserve(W, (uint32_t)ws.h_joints.size() * W->num_worlds)) return false;
	W->h_joints = ws.h_joints;
	W->joints_dirty = true;
	W->jc.num_joints = (uint32_t)ws.h_joints.size() * W->num_worlds;
	if (!ws.h_joints.empty()) rt.copy(W->jc.state, ws.joint_state, ws.h_joints.size() * W->num_worlds);
	// contact cache: the snapshot becomes the read cache, the write cache is empty between steps
	int ri = W->write_idx ^ 1;
	clear_cache(W, ri);
	rt.copy(W->cache[ri].pairs, ws.pairs, ws.num_pairs); rt.copy(W->cache[ri].manifolds, ws.manifolds, ws.num_manifolds);
	rt.upload(W->cache[ri].num_pairs, &ws.num_pairs, 1); rt.upload(W->cache[ri].num_manifolds, &ws.num_manifolds, 1);
	W->cache_num_pairs[ri] = ws.num_pairs; W->cache_num_manifolds[ri] = ws.num_manifolds;
	{ KRebuildPairTable k; k.w = d; rt.launch(k, ws.num_pairs); }
	rt.memset_(W->nc.woken_flag, 0, (size_t)d.max_bodies * 4);
	// host side of the world
	W->stepped_list = nullptr; W->stepped_count = 0;
	W->num_slots = n; W->num_active = ws.num_active; W->num_bodies = ws.num_bodies; W->prev_dt = ws.prev_dt; d.gravity = ws.gravity;
	std::copy(ws.h_ids.begin(), ws.h_ids.end(), W->h_ids.begin()); std::copy(ws.h_layer.begin(), ws.h_layer.end(), W->h_layer.begin()); std::copy(ws.h_static.begin(), ws.h_static.end(), W->h_static.begin());
	W->layer_bodies = ws.layer_bodies; W->layer_has_moving = ws.layer_has_moving;
	for (uint32_t l = 0; l < d.num_bp_layers; ++l) { W->layer_list_dirty[l] = 1; W->layer_needs_build[l] = 1; }
	W->last_num_events = 0; W->last_num_act_events = 0;
	rt.sync();
	return rt.check("b2j_world_restore_state");
}

b2j_snapshot *b2j_world_save_state(b2j_world *W)
{
	if (W == nullptr) { last_error() = "b2j_world_save_state: null world"; return nullptr; }
	b2j_snapshot *s = new b2j_snapshot;
	s->worlds.resize(1);
	if (!snapshot_take(W, s->worlds[0])) { delete s; return nullptr; }
	return s;
}

int b2j_world_restore_state(b2j_world *W, const b2j_snapshot *s)
{
	if (W == nullptr || s == nullptr || s->batch != nullptr || s->worlds.size() != 1) { last_error() = "b2j_world_restore_state: not a snapshot of a single world"; return -1; }
	return snapshot_restore(W, s->worlds[0])? 0 : -1;
}

void b2j_snapshot_destroy(b2j_snapshot *s)
{
	if (s == nullptr) return;
	for (WorldSnapshot &ws : s->worlds) snapshot_free(ws);
	delete s;
}

uint64_t b2j_snapshot_size(const b2j_snapshot *s)
{
	uint64_t total = 0;
	if (s != nullptr) for (const WorldSnapshot &ws : s->worlds) total += ws.bytes;
	return total;
}

int b2j_step(b2j_world *W, float delta_time, int collision_steps, b2j_step_stats *stats)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	bool want_energy = stats != nullptr && stats->kinetic_energy < 0.0f;
	if (stats != nullptr) { memset(stats, 0, sizeof(*stats)); if (want_energy) stats->kinetic_energy = -1.0f; }
	if (collision_steps < 1) collision_steps = 1;
	upload_shapes(W);
	sync_dworld(W);
	rt.launches = 0;
	W->last_num_events = 0;
	W->last_num_act_events = 0;
	W->stepped_list = nullptr; W->stepped_count = 0;
	if (W->num_active == 0 || delta_time <= 0.0f)
	{
		// PhysicsSystem.cpp:191-207: nothing to simulate; if time passes all cached contacts are reported as removed
		if (delta_time > 0.0f)
		{
			rt.memset_(W->d.counters, 0, sizeof(StepCounters));
			uint32_t old_m = W->cache_num_manifolds[W->write_idx ^ 1];
			if (W->nc.events != nullptr) { KRemovedEvents k; k.w = W->d; k.c = W->nc; rt.launch(k, old_m); }
			read_counters(W);
			W->last_num_events = W->h_counters.num_events;
			W->write_idx ^= 1;
			clear_cache(W, W->write_idx);
			sync_dworld(W);
		}
		if (stats != nullptr) { stats->num_bodies = W->num_bodies; stats->kernel_launches = rt.launches; }
		return rt.check("b2j_step")? 0 : -1;
	}
#ifndef B2J_HOSTSIM
	cudaEventRecord(W->ev_begin, rt.stream);
#else
	auto t0 = std::chrono::high_resolution_clock::now();
#endif
	float step_dt = delta_time / (float)collision_steps;
	float ratio = W->prev_dt > 0.0f? step_dt / W->prev_dt : 0.0f;
	W->prev_dt = step_dt;
	if (!W->d.settings.constraint_warm_start) ratio = 0.0f;
	uint32_t errors = 0;
	for (int s = 0; s < collision_steps; ++s)
	{
		float r = s == 0? ratio : (W->d.settings.constraint_warm_start? 1.0f : 0.0f);
		if (!collision_step(W, step_dt, r, s == collision_steps - 1, stats))
			return -1;
		errors |= W->h_counters.error_bits;
	}
	if (!W->h_cache_invalid.empty())
	{
		// BodyManager::ValidateContactCacheForAllBodies: the flags only live for one update
		uint32_t n = (uint32_t)W->h_cache_invalid.size();
		if (!grow(rt, W->d_cache_invalid, W->cache_invalid_capacity, n, false)) return -1;
		rt.upload(W->d_cache_invalid, W->h_cache_invalid.data(), n);
		sync_dworld(W);
		{ KValidateContactCache k; k.w = W->d; k.slots = W->d_cache_invalid; rt.launch(k, n); }
		W->h_cache_invalid.clear();
	}
	if (rt.profiling) { rt.sync(); rt.prof_collect(); }
	if (stats != nullptr)
	{
#ifndef B2J_HOSTSIM
		cudaEventRecord(W->ev_end, rt.stream);
		cudaEventSynchronize(W->ev_end);
		cudaEventElapsedTime(&stats->gpu_ms, W->ev_begin, W->ev_end);
#else
		stats->gpu_ms = (float)std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
#endif
		stats->num_active_bodies = W->num_active;
		stats->num_bodies = W->num_bodies;
		stats->kernel_launches = rt.launches;
		stats->error_bits = errors & 7u;
	}
	if (errors & 0x100u) { last_error() = "solver phase limit exceeded"; return -1; }
	return (int)(errors & 7u);
}

uint32_t b2j_events_drain(b2j_world *W, b2j_contact_event *out, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	if (W->nc.events == nullptr) return 0;
	uint32_t n = W->last_num_events < W->max_events? W->last_num_events : W->max_events;
	if (n > 0 && out != nullptr && cap > 0)
	{
		std::vector<b2j_contact_event> ev(n);
		W->rt.download(ev.data(), W->nc.events, n);
		// canonical order (kind, body1, body2, sub shapes): the 148 byte records stay put, a 4 byte index array is sorted
		std::vector<uint32_t> order(n);
		std::iota(order.begin(), order.end(), 0u);
		std::sort(order.begin(), order.end(), [&ev](uint32_t ia, uint32_t ib) {
			const b2j_contact_event &a = ev[ia], &b = ev[ib];
			if (a.kind != b.kind) return a.kind < b.kind;
			if (a.body1 != b.body1) return a.body1 < b.body1;
			if (a.body2 != b.body2) return a.body2 < b.body2;
			if (a.sub_shape1 != b.sub_shape1) return a.sub_shape1 < b.sub_shape1;
			return a.sub_shape2 < b.sub_shape2;
		});
		for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = ev[order[i]];
	}
	return n;
}

uint32_t b2j_activation_events_drain(b2j_world *W, b2j_activation_event *out, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	if (W->d_act_events == nullptr) return 0;
	uint32_t n = W->last_num_act_events < W->max_act_events? W->last_num_act_events : W->max_act_events;
	if (n > 0 && out != nullptr && cap > 0)
	{
		std::vector<b2j_activation_event> ev(n);
		W->rt.download(ev.data(), W->d_act_events, n);
		std::sort(ev.begin(), ev.end(), [](const b2j_activation_event &a, const b2j_activation_event &b) { return a.kind != b.kind? a.kind < b.kind : a.body < b.body; });
		for (uint32_t i = 0; i < n && i < cap; ++i) out[i] = ev[i];
	}
	return n;
}

uint32_t b2j_debug_get_pairs(b2j_world *W, uint32_t *pairs, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	uint32_t n = W->last_num_pairs;
	if (n > 0 && pairs != nullptr)
	{
		std::vector<BodyPair> bp(n);
		W->rt.download(bp.data(), W->nc.pairs, n);
		std::vector<std::pair<uint32_t, uint32_t>> sorted(n);
		for (uint32_t i = 0; i < n; ++i)
		{
			uint32_t a = W->h_ids[bp[i].a], b = W->h_ids[bp[i].b];
			sorted[i] = std::make_pair(a < b? a : b, a < b? b : a);
		}
		std::sort(sorted.begin(), sorted.end());
		for (uint32_t i = 0; i < n && i < cap; ++i) { pairs[2 * i] = sorted[i].first; pairs[2 * i + 1] = sorted[i].second; }
	}
	return n;
}

uint32_t b2j_debug_get_manifolds(b2j_world *W, b2j_debug_manifold *out, uint32_t cap)
{
	B2J_DEVICE_GUARD(W);
	// manifolds of the cache written by the last step (now the read cache)
	int ri = W->write_idx ^ 1;
	uint32_t nm = W->cache_num_manifolds[ri];
	if (nm > 0 && out != nullptr)
	{
		std::vector<CachedManifold> m(nm);
		W->rt.download(m.data(), W->cache[ri].manifolds, nm);
		std::vector<ManifoldWS> ws(nm);
		W->rt.download(ws.data(), W->nc.man_ws, nm);
		std::vector<b2j_debug_manifold> r(nm);
		for (uint32_t i = 0; i < nm; ++i)
		{
			r[i].body1 = m[i].body1; r[i].body2 = m[i].body2; r[i].sub_shape1 = m[i].sub1; r[i].sub_shape2 = m[i].sub2;
			r[i].num_points = m[i].num_points;
			r[i].from_cache = (m[i].flags & MANIFOLD_FROM_CACHE)? 1 : 0;
			for (int k = 0; k < 3; ++k) r[i].normal[k] = r[i].from_cache? m[i].normal[k] : ws[i].normal[k];
			r[i].penetration_depth = 0.0f;
		}
		std::sort(r.begin(), r.end(), [](const b2j_debug_manifold &a, const b2j_debug_manifold &b) {
			if (a.body1 != b.body1) return a.body1 < b.body1;
			if (a.body2 != b.body2) return a.body2 < b.body2;
			if (a.sub_shape1 != b.sub_shape1) return a.sub_shape1 < b.sub_shape1;
			return a.sub_shape2 < b.sub_shape2;
		});
		for (uint32_t i = 0; i < nm && i < cap; ++i) out[i] = r[i];
	}
	return nm;
}

int b2j_debug_check_schedule(b2j_world *W)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	if (W->stepped_list == nullptr || W->stepped_count == 0) return 0;
	sync_dworld(W);
	uint32_t *out = reinterpret_cast<uint32_t *>(W->d_energy);
	rt.memset_(out, 0, 4);
	// (the active list of the last step = the list before the sleepers left it)
	DWorld d = W->d;
	d.active = const_cast<uint32_t *>(W->stepped_list);
	{ KCheckSchedule k; k.w = d; k.s = W->sc; k.conflicts = out; rt.launch(k, W->stepped_count); }
	uint32_t n = 0;
	rt.download(&n, out, 1);
	if (!rt.check("b2j_debug_check_schedule")) return -1;
	return (int)n;
}

int b2j_debug_find_pairs(b2j_world *W)
{
	B2J_DEVICE_GUARD(W);
	Runtime &rt = W->rt;
	upload_shapes(W);
	sync_dworld(W);
	rt.memset_(W->d.counters, 0, sizeof(StepCounters));
	for (uint32_t l = 0; l < W->d.num_bp_layers; ++l)
		if (W->layer_needs_build[l])
		{
			if (!build_tree(W, l)) return -1;
			W->layer_needs_build[l] = 0;
		}
	KFindPairs k; k.w = W->d; for (int l = 0; l < 8; ++l) k.trees[l] = W->trees[l]; k.pairs = W->nc.pairs; k.first = 0; k.query_leaves = nullptr;
	rt.launch(k, W->num_active);
	if (!read_counters(W)) return -1;
	W->last_num_pairs = W->h_counters.num_pairs < W->d.max_body_pairs? W->h_counters.num_pairs : W->d.max_body_pairs;
	return 0;
}

} // extern "C" (paused for a template helper)

template <class T> static void replicate(Runtime &rt, T *dst, const T *src, uint32_t n_src, uint32_t stride, uint32_t n_worlds)
{
	KReplicate<T> k; k.dst = dst; k.src = src; k.n_src = n_src; k.dst_stride = stride;
	uint64_t total = (uint64_t)n_src * n_worlds;
	// launches are limited to 2^32 items: split by worlds
	uint32_t worlds_per_launch = n_src == 0? n_worlds : (uint32_t)std::max<uint64_t>(1, 0x7fffffffull / n_src);
	for (uint32_t w0 = 0; w0 < n_worlds; w0 += worlds_per_launch)
	{
		uint32_t nw = std::min(worlds_per_launch, n_worlds - w0);
		KReplicate<T> kk = k; kk.dst = dst + (size_t)w0 * stride;
		rt.launch(kk, n_src * nw);
	}
	(void)total;
}

extern "C" {

// How many groups (device worlds with their own stream and host thread) the worlds of one device are spread over: groups of at least
// 128 worlds, at most 8. Measured with Pyramid worlds: at 4096 worlds 149 ms per step with 4 groups, 130 with 8, 125 with 16 (but the
// 16 group launches are small enough to lose 20% of their own HBM efficiency); at 512 worlds (the share of one GPU of eight) 11.9 ms
// with 1 group, 10.7 with 2, 9.9 with 4, 10.6 with 8. B2J_BATCH_GROUPS overrides.
static uint32_t batch_default_groups(uint32_t n_worlds, uint32_t slots_per_world)
{
	// groups of >= 128 worlds AND >= ~150 k body slots, at most 8: a group must keep the SMs busy on its own stream, and every group
	// adds its own chain of launches (measured: Pyramid worlds of 1 241 bodies 512 worlds 11.9 / 10.7 / 9.9 / 10.6 ms per step with
	// 1 / 2 / 4 / 8 groups; worlds of the joints scene, 337 bodies: 2048 worlds 4.8 / 4.5 / 4.8 / 5.6 ms, 8192 worlds 12.8 / 11.8 / 11.2 / 11.3 ms)
	uint32_t K = n_worlds / 128;
	uint64_t by_slots = (uint64_t)n_worlds * slots_per_world / 150000u;
	if (K > by_slots) K = (uint32_t)by_slots;
	if (K > 8) K = 8;
	if (const char *e = getenv("B2J_BATCH_GROUPS")) K = (uint32_t)atoi(e);
	return K;
}

static b2j_world *batch_create_group(b2j_world *P, uint32_t n_worlds, uint32_t max_body_pairs_per_world, uint32_t max_contact_constraints_per_world, int device)
{
	B2J_DEVICE_GUARD(P);
	upload_shapes(P);
	sync_dworld(P);
	uint32_t stride = P->num_slots;
	if (stride == 0 || (uint64_t)stride * n_worlds > 0xfffffff0ull) { last_error() = "b2j_batch_create: too many bodies"; return nullptr; }
	b2j_world_desc desc = P->desc;
	desc.object_to_broadphase = P->t_o2bp.data(); desc.object_vs_broadphase = P->t_ovbp.data(); desc.object_vs_object = P->t_ovo.data();
	desc.settings = P->d.settings;
	v3_store(P->d.gravity, desc.gravity);
	desc.max_bodies = stride * n_worlds;
	uint64_t pairs = (uint64_t)(max_body_pairs_per_world != 0? max_body_pairs_per_world : P->d.max_body_pairs) * n_worlds;
	uint64_t cons = (uint64_t)(max_contact_constraints_per_world != 0? max_contact_constraints_per_world : P->d.max_constraints) * n_worlds;
	if (pairs > 0x7ffffff0ull || cons > 0x7ffffff0ull) { last_error() = "b2j_batch_create: limits exceed 2^31"; return nullptr; }
	desc.max_body_pairs = (uint32_t)pairs;
	desc.max_contact_constraints = (uint32_t)cons;
	desc.device = device;
#ifndef B2J_HOSTSIM
	if (device != P->rt.device)
	{
		// the group lives on another device of this process: its creation kernels read the prototype over NVLink (peer access), and the
		// reset kernels read the creation state kept on the first group's device
		int can = 0;
		cudaDeviceCanAccessPeer(&can, device, P->rt.device);
		if (!can) { last_error() = "b2j_batch_create_on_devices: no peer access between the devices"; return nullptr; }
		cudaSetDevice(device);
		cudaError_t pe = cudaDeviceEnablePeerAccess(P->rt.device, 0);
		if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { last_error() = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe); return nullptr; }
		cudaGetLastError();
	}
#endif
	g_create_without_events = true;
	b2j_world *B = b2j_world_create(&desc);
	g_create_without_events = false;
	if (B == nullptr) return nullptr;
	Runtime &rt = B->rt;
	B->num_worlds = n_worlds;
	for (int i = 0; i < 2; ++i) { B->d_collide_keys[i] = rt.alloc<uint32_t>(B->d.max_body_pairs, false); B->d_collide_vals[i] = rt.alloc<uint32_t>(B->d.max_body_pairs, false); }
	B->d.world_stride = stride;
	B->prev_dt = P->prev_dt;
	// shapes are shared by all worlds
	B->h_shapes = P->h_shapes; B->h_shape_meta = P->h_shape_meta; B->h_hull_points = P->h_hull_points; B->h_hull_shrunk = P->h_hull_shrunk; B->h_hull_planes = P->h_hull_planes;
	B->h_hull_faces = P->h_hull_faces; B->h_hull_vtx = P->h_hull_vtx; B->h_mesh_bytes = P->h_mesh_bytes; B->h_compound_subs = P->h_compound_subs;
	B->shapes_dirty = true;
	upload_shapes(B);
	// body state: world w occupies the slots [w * stride, (w + 1) * stride)
	P->rt.sync();
	const DWorld &s = P->d; DWorld &d = B->d;
	replicate(rt, d.info, s.info, stride, stride, n_worlds); replicate(rt, d.params, s.params, stride, stride, n_worlds);
	replicate(rt, d.position.base, s.position.base, 2 * stride, 2 * stride, n_worlds);               // (pair arrays: 2 elements per body)
	replicate(rt, d.linear_velocity.base, s.linear_velocity.base, 2 * stride, 2 * stride, n_worlds);
	replicate(rt, d.force.base, s.force.base, 2 * stride, 2 * stride, n_worlds);
	replicate(rt, d.inv_inertia_diag.base, s.inv_inertia_diag.base, 2 * stride, 2 * stride, n_worlds);
	replicate(rt, d.bounds_min.base, s.bounds_min.base, 2 * stride, 2 * stride, n_worlds);
	replicate(rt, d.sleep_spheres, s.sleep_spheres, stride * 3, stride * 3, n_worlds); replicate(rt, d.sleep_timer, s.sleep_timer, stride, stride, n_worlds);
	// host mirrors
	B->num_slots = stride * n_worlds;
	B->num_bodies = P->num_bodies * n_worlds;
	for (uint32_t w = 0; w < n_worlds; ++w)
		for (uint32_t i = 0; i < stride; ++i)
		{
			B->h_ids[(size_t)w * stride + i] = P->h_ids[i];
			B->h_layer[(size_t)w * stride + i] = P->h_layer[i];
			B->h_static[(size_t)w * stride + i] = P->h_static[i];
		}
	for (uint32_t l = 0; l < d.num_bp_layers; ++l)
	{
		B->layer_bodies[l].reserve(P->layer_bodies[l].size() * n_worlds);
		for (uint32_t w = 0; w < n_worlds; ++w)
			for (uint32_t slot : P->layer_bodies[l])
				B->layer_bodies[l].push_back(slot + w * stride);
		B->layer_list_dirty[l] = 1; B->layer_needs_build[l] = 1; B->layer_has_moving[l] = P->layer_has_moving[l];
		B->trees[l].world_root = rt.alloc<int32_t>(n_worlds);
	}
	// active list
	sync_dworld(B);
	if (P->num_active > 0)
	{
		KReplicateActive k; k.w = B->d; k.src_active = P->d.active; k.na = P->num_active; k.stride = stride;
		rt.launch(k, P->num_active * n_worlds);
	}
	B->num_active = P->num_active * n_worlds;
	// non contact constraints: every world gets the prototype's list (bodies offset to the world's slots) and its current state
	if (!P->h_joints.empty())
	{
		uint32_t nj = (uint32_t)P->h_joints.size();
		if ((uint64_t)nj * n_worlds > 0x7ffffff0ull || !joints_reserve(B, nj * n_worlds)) { b2j_world_destroy(B); return nullptr; }
		B->h_joints = P->h_joints;
		B->joints_dirty = true;
		B->d_joint_init = rt.alloc<JointState>(nj, false);
		if (B->d_joint_init == nullptr) { last_error() = "out of device memory"; b2j_world_destroy(B); return nullptr; }
		rt.copy(B->d_joint_init, P->jc.state, nj);
		replicate(rt, B->jc.state, P->jc.state, nj, nj, n_worlds);
	}
	rt.sync();
	if (!rt.check("b2j_batch_create")) { b2j_world_destroy(B); return nullptr; }
	return B;
}

} // extern "C"

// Runs fn(group index) for every group, one host thread per group (the caller runs group 0); returns false if any failed
template <class F> static bool batch_for_each_group(b2j_batch *b, const F &fn)
{
	size_t K = b->groups.size();
	std::vector<int> ok(K, 1);
	std::vector<std::string> err(K);
	auto work = [&](size_t g, bool worker)
	{
#ifndef B2J_HOSTSIM
		cudaSetDevice(b->groups[g]->rt.device); // (the calling thread too: an earlier call may have left another device of the batch current)
#endif
		(void)worker;
		if (!fn(g)) { ok[g] = 0; err[g] = last_error(); }
	};
	std::vector<std::thread> threads;
	for (size_t g = 1; g < K; ++g) threads.emplace_back(work, g, true);
	work(0, false);
	for (std::thread &t : threads) t.join();
	for (size_t g = 0; g < K; ++g)
		if (!ok[g]) { last_error() = err[g]; return false; }
	return true;
}

extern "C" {

// the creation state of one world (device resident, on the first group's device), for b2j_batch_reset_worlds
static bool batch_keep_init_state(b2j_batch *b, b2j_world *P)
{
	B2J_DEVICE_GUARD(b->groups[0]);
		Runtime &rt = b->groups[0]->rt;
		const DWorld &s = P->d;
		uint32_t stride = b->stride;
		BatchInitState &in = b->init;
		auto keep = [&](F4 *&dst, const F4 *src, uint32_t n) { dst = rt.alloc<F4>(n, false); rt.copy(dst, src, n); };
		keep(in.pose, s.position.base, 2 * stride); keep(in.velocity, s.linear_velocity.base, 2 * stride);
		keep(in.force_torque, s.force.base, 2 * stride); keep(in.bounds, s.bounds_min.base, 2 * stride);
		keep(in.sleep_spheres, s.sleep_spheres, stride * 3);
		in.sleep_timer = rt.alloc<float>(stride, false); rt.copy(in.sleep_timer, s.sleep_timer, stride);
		std::vector<uint32_t> active_index(stride), was_active(stride);
		std::vector<BodyInfo> info(stride);
		P->rt.download(active_index.data(), s.active_index, stride);
		P->rt.download(info.data(), s.info, stride);
		for (uint32_t j = 0; j < stride; ++j)
		{
			was_active[j] = active_index[j] != B2J_INACTIVE_INDEX? 1u : 0u;
			if (!was_active[j] && info[j].id != 0xffffffffu && info[j].motion_type != B2J_MOTION_STATIC) b->has_init_inactive = true;
		}
		in.was_active = rt.alloc<uint32_t>(stride, false);
		rt.upload(in.was_active, was_active.data(), stride);
		rt.sync();
	return rt.check("b2j_batch_create");
}

b2j_batch *b2j_batch_create(b2j_world *P, uint32_t n_worlds, uint32_t max_body_pairs_per_world, uint32_t max_contact_constraints_per_world)
{
	if (P == nullptr) { last_error() = "b2j_batch_create: invalid prototype"; return nullptr; }
	int32_t device = P->rt.device;
	return b2j_batch_create_on_devices(P, n_worlds, &device, 1, max_body_pairs_per_world, max_contact_constraints_per_world);
}

b2j_batch *b2j_batch_create_on_devices(b2j_world *P, uint32_t n_worlds, const int32_t *device_ids, uint32_t n_devices, uint32_t max_body_pairs_per_world, uint32_t max_contact_constraints_per_world)
{
	if (P == nullptr || n_worlds == 0 || P->num_worlds != 1 || device_ids == nullptr || n_devices == 0 || n_devices > n_worlds) { last_error() = "b2j_batch_create: invalid prototype or device list"; return nullptr; }
	if (device_ids[0] != P->rt.device) { last_error() = "b2j_batch_create_on_devices: the first device must be the prototype's"; return nullptr; }
	if (n_devices > 1)
	{
		// contiguous blocks of worlds per device, each block split into groups like a single device batch; the groups of all devices form
		// ONE batch (world index -> group -> device), stepped by one host thread per group with no inter device traffic
		b2j_batch *b = new b2j_batch;
		b->n_worlds = n_worlds; b->stride = P->num_slots; b->bodies_per_world = P->num_bodies;
		uint32_t first = 0;
		for (uint32_t dv = 0; dv < n_devices; ++dv)
		{
			uint32_t nd = n_worlds / n_devices + (dv < n_worlds % n_devices? 1 : 0);
			uint32_t K = batch_default_groups(nd, P->num_slots);
			if (K < 1) K = 1;
			if (K > nd) K = nd;
			for (uint32_t g = 0; g < K; ++g)
			{
				uint32_t n = nd / K + (g < nd % K? 1 : 0);
				b2j_world *G = batch_create_group(P, n, max_body_pairs_per_world, max_contact_constraints_per_world, device_ids[dv]);
				if (G == nullptr) { b2j_batch_destroy(b); return nullptr; }
				b->groups.push_back(G);
				b->first_world.push_back(first);
				first += n;
			}
		}
		b->first_world.push_back(first);
		if (!batch_keep_init_state(b, P)) { b2j_batch_destroy(b); return nullptr; }
#ifndef B2J_HOSTSIM
		// the reset kernels of the other devices read the creation state from the first group's device
		for (b2j_world *G : b->groups)
			if (G->rt.device != b->groups[0]->rt.device)
			{
				cudaSetDevice(G->rt.device);
				cudaDeviceEnablePeerAccess(b->groups[0]->rt.device, 0);
				cudaGetLastError();
			}
		cudaSetDevice(P->rt.device);
#endif
		return b;
	}
	uint32_t K = batch_default_groups(n_worlds, P->num_slots);
#ifdef B2J_HOSTSIM
	K = 1; // the host simulation is single threaded
#endif
	if (K < 1) K = 1;
	if (K > n_worlds) K = n_worlds;
	b2j_batch *b = new b2j_batch;
	b->n_worlds = n_worlds; b->stride = P->num_slots; b->bodies_per_world = P->num_bodies;
	uint32_t first = 0;
	for (uint32_t g = 0; g < K; ++g)
	{
		uint32_t n = n_worlds / K + (g < n_worlds % K? 1 : 0);
		b2j_world *G = batch_create_group(P, n, max_body_pairs_per_world, max_contact_constraints_per_world, P->rt.device);
		if (G == nullptr) { b2j_batch_destroy(b); return nullptr; }
		// (experiments: the one launch solvers of the groups share the SMs instead of taking turns on all of them)
		if (const char *e = getenv("B2J_SOLVE_GRID_DIV")) { int v = atoi(e); G->solve_grid_div = v < 1? 1u : (uint32_t)v; }
		G->batch_groups = K;
		b->groups.push_back(G);
		b->first_world.push_back(first);
		first += n;
	}
	b->first_world.push_back(first);
	if (!batch_keep_init_state(b, P)) { b2j_batch_destroy(b); return nullptr; }
	return b;
}

void b2j_batch_destroy(b2j_batch *b)
{
	if (b == nullptr) return;
	if (!b->groups.empty())
	{
		Runtime &rt = b->groups[0]->rt;
		BatchInitState &in = b->init;
		rt.free_(in.pose); rt.free_(in.velocity); rt.free_(in.force_torque); rt.free_(in.bounds); rt.free_(in.sleep_spheres); rt.free_(in.sleep_timer); rt.free_(in.was_active);
	}
	for (b2j_world *G : b->groups) b2j_world_destroy(G);
	delete b;
}

// Resets the given worlds to the state the batch was created with: bodies (pose, velocities, forces, bounds, sleep state, active
// flag) and an empty contact cache, exactly as if those worlds had just been created; the other worlds are untouched. No body
// state crosses the PCIe bus (the RL pattern: environments terminate at different steps).
int b2j_batch_reset_worlds(b2j_batch *b, const uint32_t *world_indices, uint32_t n)
{
	if (b == nullptr || (n > 0 && world_indices == nullptr)) { last_error() = "b2j_batch_reset_worlds: invalid arguments"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
		if (world_indices[i] >= b->n_worlds) { last_error() = "b2j_batch_reset_worlds: world index out of range"; return -1; }
	if (n == 0) return 0;
	bool ok = batch_for_each_group(b, [&](size_t g) {
		b2j_world *G = b->groups[g];
		Runtime &rt = G->rt;
		uint32_t first = b->first_world[g], end = b->first_world[g + 1], nw = end - first, stride = b->stride;
		std::vector<uint32_t> local, flags(nw, 0);
		for (uint32_t i = 0; i < n; ++i)
			if (world_indices[i] >= first && world_indices[i] < end && !flags[world_indices[i] - first]) { flags[world_indices[i] - first] = 1; local.push_back(world_indices[i] - first); }
		if (local.empty()) return true;
		sync_dworld(G);
		uint32_t *d_local = rt.alloc<uint32_t>(local.size(), false), *d_flags = rt.alloc<uint32_t>(nw, false);
		rt.upload(d_local, local.data(), local.size());
		rt.upload(d_flags, flags.data(), nw);
		uint32_t items = (uint32_t)local.size() * stride;
		// 1. bodies that were active at creation but sleep now: back onto the active list (sorted by slot, like the step does)
		rt.memset_(G->d.counters, 0, sizeof(StepCounters));
		{ KResetFindInactive k; k.w = G->d; k.c = G->nc; k.init = b->init; k.worlds = d_local; rt.launch(k, items); }
		if (!read_counters(G)) return false;
		uint32_t woken = G->h_counters.num_woken;
		if (woken > 0)
		{
			rt.sort_pairs<uint32_t>(G->nc.woken_list, G->d_woken_keys, G->nc.woken_list, G->d_woken_sorted, woken, 32);
			KActivateWoken k; k.w = G->d; k.woken_sorted = G->d_woken_sorted; k.base = G->num_active; k.woken_flag = G->nc.woken_flag; k.events = nullptr; k.max_events = 0;
			rt.launch(k, woken);
			G->num_active += woken;
		}
		// bodies that were NOT active at creation (rare) leave the active list: the stable compaction of the end of a step
		if (b->has_init_inactive && G->num_active > 0)
		{
			uint32_t na = G->num_active;
			{ KResetMarkKeep k; k.w = G->d; k.init = b->init; k.reset_flag = d_flags; k.keep = G->d_keep; rt.launch(k, na); }
			rt.exclusive_scan(G->d_keep, G->d_keep_scan, na);
			uint32_t *new_list = G->active_buf[G->active_cur ^ 1];
			{ KCompactActive k; k.w = G->d; k.keep = G->d_keep; k.keep_scan = G->d_keep_scan; k.new_active = new_list; rt.launch(k, na); }
			{ KFinishCompact k; k.w = G->d; k.keep = G->d_keep; k.keep_scan = G->d_keep_scan; k.n = na; rt.launch(k, 1); }
			G->active_cur ^= 1;
			if (!read_counters(G)) return false;
			G->num_active = G->h_counters.new_active_count;
		}
		// 2. state of the bodies, 3. forget the contacts of those worlds
		{ KResetWorlds k; k.w = G->d; k.init = b->init; k.worlds = d_local; rt.launch(k, items); }
		{ KResetPurgeCache k; k.w = G->d; k.reset_flag = d_flags; rt.launch(k, G->cache_num_pairs[G->write_idx ^ 1]); }
		if (!G->h_joints.empty())
		{
			// 4. the constraints of those worlds forget their accumulated impulses (back to the creation state)
			KResetJoints k; k.state = G->jc.state; k.init = G->d_joint_init; k.worlds = d_local; k.per_world = (uint32_t)G->h_joints.size();
			rt.launch(k, (uint32_t)(local.size() * G->h_joints.size()));
		}
		rt.sync();
		rt.free_(d_local); rt.free_(d_flags);
		for (uint32_t l = 0; l < G->d.num_bp_layers; ++l) G->layer_needs_build[l] = 1;
		return rt.check("b2j_batch_reset_worlds");
	});
	return ok? 0 : -1;
}

int b2j_batch_query_cast_rays(b2j_batch *b, const uint32_t *ray_world, const b2j_ray *rays, uint32_t n, uint32_t object_layer, b2j_ray_hit *hits)
{
	if (b == nullptr || (n > 0 && ray_world == nullptr)) { last_error() = "b2j_batch_query_cast_rays: batch and ray_world are required"; return -1; }
	for (uint32_t i = 0; i < n; ++i)
		if (ray_world[i] >= b->n_worlds) { last_error() = "b2j_batch_query_cast_rays: world index out of range"; return -1; }
	// every group casts the rays of its own worlds (each ray is answered by exactly one group)
	return batch_for_each_group(b, [&](size_t g) { return cast_rays(b->groups[g], ray_world, b->first_world[g], rays, n, object_layer, hits) == 0; })? 0 : -1;
}

b2j_snapshot *b2j_batch_save_state(b2j_batch *b)
{
	if (b == nullptr) { last_error() = "b2j_batch_save_state: null batch"; return nullptr; }
	b2j_snapshot *s = new b2j_snapshot;
	s->batch = b;
	s->worlds.resize(b->groups.size());
	bool ok = batch_for_each_group(b, [&](size_t g) { return snapshot_take(b->groups[g], s->worlds[g]); });
	if (!ok) { std::string e = last_error(); b2j_snapshot_destroy(s); last_error() = e; return nullptr; }
	return s;
}

int b2j_batch_restore_state(b2j_batch *b, const b2j_snapshot *s)
{
	if (b == nullptr || s == nullptr || s->batch != b || s->worlds.size() != b->groups.size()) { last_error() = "b2j_batch_restore_state: not a snapshot of this batch"; return -1; }
	return batch_for_each_group(b, [&](size_t g) { return snapshot_restore(b->groups[g], s->worlds[g]); })? 0 : -1;
}

int b2j_batch_step(b2j_batch *b, float dt, int collision_steps, b2j_step_stats *stats)
{
	size_t K = b->groups.size();
	if (K == 1) return b2j_step(b->groups[0], dt, collision_steps, stats);
	std::vector<b2j_step_stats> st(K);
	std::vector<int> rc(K, 0);
	if (stats != nullptr) for (size_t g = 0; g < K; ++g) st[g] = *stats; // carries the kinetic energy request
	bool ok = true;
	if (b->groups[0]->rt.profiling)
	{
		// per kernel timing: one group at a time, so that a kernel's events do not include kernels of the other streams
		for (size_t g = 0; g < K && ok; ++g) { rc[g] = b2j_step(b->groups[g], dt, collision_steps, stats != nullptr? &st[g] : nullptr); ok = rc[g] >= 0; }
	}
	else
		ok = batch_for_each_group(b, [&](size_t g) { rc[g] = b2j_step(b->groups[g], dt, collision_steps, stats != nullptr? &st[g] : nullptr); return rc[g] >= 0; });
	if (!ok) return -1;
	int r = 0;
	for (size_t g = 0; g < K; ++g) r |= rc[g];
	if (stats != nullptr)
	{
		b2j_step_stats a = st[0];
		for (size_t g = 1; g < K; ++g)
		{
			const b2j_step_stats &s = st[g];
			a.num_active_bodies += s.num_active_bodies; a.num_bodies += s.num_bodies; a.num_body_pairs += s.num_body_pairs; a.num_pairs_from_cache += s.num_pairs_from_cache;
			a.num_manifolds += s.num_manifolds; a.num_contact_points += s.num_contact_points; a.num_constraints += s.num_constraints; a.num_islands += s.num_islands;
			a.num_large_islands += s.num_large_islands; a.num_activated += s.num_activated; a.num_deactivated += s.num_deactivated; a.kernel_launches += s.kernel_launches;
			a.kinetic_energy += s.kinetic_energy; a.error_bits |= s.error_bits;
			a.num_phases = std::max(a.num_phases, s.num_phases); a.velocity_iterations = std::max(a.velocity_iterations, s.velocity_iterations);
			a.position_iterations = std::max(a.position_iterations, s.position_iterations);
			a.gpu_ms = b->groups[0]->rt.profiling? a.gpu_ms + s.gpu_ms : std::max(a.gpu_ms, s.gpu_ms); // profiling steps the groups one after the other
		}
		*stats = a;
	}
	return r;
}

uint32_t b2j_batch_size(const b2j_batch *b) { return b != nullptr? b->n_worlds : 0; }

static b2j_body_state offset_state(const b2j_body_state &s, size_t first)
{
	b2j_body_state o = s;
	if (o.position) o.position += 3 * first;
	if (o.rotation) o.rotation += 4 * first;
	if (o.linear_velocity) o.linear_velocity += 3 * first;
	if (o.angular_velocity) o.angular_velocity += 3 * first;
	if (o.bounds) o.bounds += 6 * first;
	if (o.active_index) o.active_index += first;
	if (o.sleep_timer) o.sleep_timer += first;
	return o;
}

int b2j_batch_get_state(b2j_batch *b, uint32_t world_index, uint32_t n, const b2j_body_state *out)
{
	if (b != nullptr && world_index == 0xffffffffu && (uint64_t)n <= (uint64_t)b->stride * b->n_worlds)
	{
		// all worlds: slots [0, n), every group copies its share
		bool ok = batch_for_each_group(b, [&](size_t g) {
			uint64_t first = (uint64_t)b->first_world[g] * b->stride, end = (uint64_t)b->first_world[g + 1] * b->stride;
			if (first >= n) return true;
			b2j_body_state o = offset_state(*out, (size_t)first);
			return b2j_bodies_get_state(b->groups[g], nullptr, (uint32_t)(std::min<uint64_t>(end, n) - first), &o) == 0;
		});
		return ok? 0 : -1;
	}
	if (b == nullptr || world_index >= b->n_worlds || n > b->stride) { last_error() = "b2j_batch_get_state: invalid arguments"; return -1; }
	size_t g = 0;
	while (world_index >= b->first_world[g + 1]) ++g;
	b2j_world *G = b->groups[g];
	G->get_state_first = (world_index - b->first_world[g]) * b->stride;
	int r = b2j_bodies_get_state(G, nullptr, n, out);
	G->get_state_first = 0;
	return r;
}

int b2j_batch_add_force_torque(b2j_batch *b, uint32_t n, const float *force, const float *torque)
{
	if (b == nullptr || (uint64_t)n > (uint64_t)b->stride * b->n_worlds) { last_error() = "b2j_batch_add_force_torque: invalid arguments"; return -1; }
	bool ok = batch_for_each_group(b, [&](size_t g) {
		uint64_t first = (uint64_t)b->first_world[g] * b->stride, end = (uint64_t)b->first_world[g + 1] * b->stride;
		if (first >= n) return true;
		return b2j_bodies_add_force_torque(b->groups[g], nullptr, (uint32_t)(std::min<uint64_t>(end, n) - first), force? force + 3 * first : nullptr, torque? torque + 3 * first : nullptr) == 0;
	});
	return ok? 0 : -1;
}

int b2j_batch_set_profiling(b2j_batch *b, int on)
{
	int r = 0;
	for (b2j_world *G : b->groups) r |= b2j_world_set_profiling(G, on);
	return r;
}

// kernel times summed over the groups (the groups run concurrently: the sum can exceed the wall time of the step)
uint32_t b2j_batch_get_profile(b2j_batch *b, char *names, uint32_t name_stride, float *ms, uint32_t *launches, uint32_t cap)
{
	if (b->groups.size() == 1) return b2j_world_get_profile(b->groups[0], names, name_stride, ms, launches, cap);
	std::vector<std::string> order;
	std::map<std::string, std::pair<float, uint32_t>> acc;
	const uint32_t kCap = 256, kStride = 96;
	std::vector<char> nm(kCap * kStride); std::vector<float> m(kCap); std::vector<uint32_t> l(kCap);
	for (b2j_world *G : b->groups)
	{
		uint32_t n = b2j_world_get_profile(G, nm.data(), kStride, m.data(), l.data(), kCap);
		for (uint32_t i = 0; i < n && i < kCap; ++i)
		{
			std::string name(&nm[i * kStride]);
			if (acc.find(name) == acc.end()) order.push_back(name);
			acc[name].first += m[i]; acc[name].second += l[i];
		}
	}
	uint32_t n = 0;
	for (const std::string &name : order)
	{
		if (n < cap)
		{
			if (names != nullptr && name_stride > 0) { strncpy(names + (size_t)n * name_stride, name.c_str(), name_stride - 1); names[(size_t)n * name_stride + name_stride - 1] = 0; }
			if (ms != nullptr) ms[n] = acc[name].first;
			if (launches != nullptr) launches[n] = acc[name].second;
		}
		++n;
	}
	return n;
}

} // extern "C"
#pragma GCC visibility pop
