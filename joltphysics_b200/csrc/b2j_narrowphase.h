// b2j_narrowphase.h -- body pair processing: contact cache test, convex vs convex collision, manifold creation, and the
// device resident contact cache (double buffered, warm start data).
//
// Restates:
//   PhysicsSystem::ProcessBodyPair                      PhysicsSystem.cpp:1049-1349 (pair orientation, reduction collector)
//   ContactConstraintManager::GetContactsFromCache      ContactConstraintManager.cpp:1001-1088, :843-999
//   ContactConstraintManager::AddBodyPair               :1090-1128
//   ContactConstraintManager::TemplatedAddContactConstraint / AddContactConstraint  :1130-1385
//   ConvexShape::sCollideConvexVsConvex                 Shape/ConvexShape.cpp:45-164
// One thread per body pair; pairs that need EPA are compacted into a second launch that owns a global memory scratch
// slot per thread (b2j_gjk.h EpaScratch). Constraints are NOT built here: every manifold that needs one gets a ConstraintSrc
// record; after island/schedule building the setup kernel (b2j_solver.h) writes constraints straight into solve order.
#pragma once

#include "b2j_world.h"
#include "b2j_shapes.h"
#include "b2j_gjk.h"
#include "b2j_manifold.h"
#include "b2j_broadphase.h"

namespace b2j {

struct CollideItem { uint32_t b1, b2; uint32_t pair_entry, old_pair; }; // b1 = body whose space the collision is done in
struct CachedItem { uint32_t pair_entry, old_pair; };
struct EpaResult { CollideItem c; float point1[3], point2[3], axis[3]; float max_separation_distance; }; // penetration found by GJK / EPA, finished by KFinishPairs
struct EpaItem { CollideItem c; }; // the EPA kernel re-runs the (deterministic) GJK step instead of carrying the simplex through HBM

// World space data of a manifold created this step (not from the cache), indexed like write_cache.manifolds
struct ManifoldWS { float normal[3]; float p1[4][3], p2[4][3]; };

// One future contact constraint
struct ConstraintSrc
{
	uint64_t sort_key;       // FNV-1a of SubShapeIDPair (mSortKey)
	uint32_t manifold;       // index in write_cache.manifolds
	uint32_t b1, b2;         // body slots, id(b1) < id(b2)
};

struct NarrowCtx
{
	BodyPair *pairs;
	CollideItem *collide_convex, *collide_mesh;
	CachedItem *cached;
	EpaItem *epa;
	const uint32_t *collide_order; // batch groups: collide_convex is processed in (pair inside the world, world) order so that the
	                             // lanes of a warp run the same pair of different worlds (near identical control flow); null = as queued
	uint32_t collide_order_n;    // entries of collide_order (queue positions past it are processed in queue order)
	EpaItem *epa_overflow;       // deep pairs that did not fit the small EPA tier (re-run on full size storage)
	uint32_t *num_epa_overflow;  // device counter
	uint32_t *epa_hist;          // B2J_TRACE_EPA only (else null): histogram of the support points the full tier's runs ended with
	EpaResult *epa_results;      // GJK / EPA output; supporting faces / clipping / manifold run in KFinishPairs
	uint32_t *num_epa_results;   // device counter
	uint32_t num_scratch;        // number of warps the scratch hungry kernels (EPA, mesh) may use
	ManifoldWS *man_ws;
	ConstraintSrc *con_src;
	uint32_t *woken_flag;        // per slot
	uint32_t *woken_list;
	b2j_contact_event *events;
	uint32_t max_events;
	uint32_t max_epa;
};

// ---- pair hash table ---------------------------------------------------------------------------------------------
// The table is keyed by the two body SLOTS (lower id first): unique across batched worlds, same as the ids in a single world
B2J_HD uint64_t pair_key(uint32_t slot1, uint32_t slot2) { return ((uint64_t)slot2 << 32) | slot1; }

B2J_D void pair_table_insert(const DWorld &w, const ContactCache &c, uint32_t slot1, uint32_t slot2, uint32_t entry)
{
	uint32_t mask = w.pair_table_size - 1;
	uint32_t h = (uint32_t)hash64(pair_key(slot1, slot2)) & mask;
	for (;;)
	{
		if (atomic_cas(&c.pair_table[h], 0xffffffffu, entry) == 0xffffffffu)
			return;
		h = (h + 1) & mask;
	}
}

B2J_D uint32_t pair_table_find(const DWorld &w, const ContactCache &c, uint32_t slot1, uint32_t slot2)
{
	uint32_t mask = w.pair_table_size - 1;
	uint32_t h = (uint32_t)hash64(pair_key(slot1, slot2)) & mask;
	for (;;)
	{
		uint32_t e = c.pair_table[h];
		if (e == 0xffffffffu)
			return 0xffffffffu;
		if (c.pairs[e].slot1 == slot1 && c.pairs[e].slot2 == slot2)
			return e;
		h = (h + 1) & mask;
	}
}

B2J_D void set_error(const DWorld &w, uint32_t bits) { atomic_or(&w.counters->error_bits, bits); }

// ---- KProcessPairs: orientation, cache test, work list routing ---------------------------------------------------
struct KProcessPairs
{
	DWorld w; NarrowCtx c;
	const uint32_t *first_ptr; // device: first pair of this round
	B2J_D void operator()(uint32_t k) const
	{
		BodyPair bp = c.pairs[*first_ptr + k];
		uint32_t b1 = bp.a, b2 = bp.b;
		BodyInfo i1 = w.info[b1], i2 = w.info[b2];
		// body1 = higher motion type, ties -> lower id (PhysicsSystem.cpp:1072-1077)
		if (i1.motion_type < i2.motion_type || (i1.motion_type == i2.motion_type && i2.id < i1.id))
		{
			uint32_t tb = b1; b1 = b2; b2 = tb;
			BodyInfo ti = i1; i1 = i2; i2 = ti;
		}
		// cache orientation: lower id first
		uint32_t cb1 = b1, cb2 = b2;
		uint32_t id1 = i1.id, id2 = i2.id;
		if (id2 < id1) { cb1 = b2; cb2 = b1; uint32_t t = id1; id1 = id2; id2 = t; }

		// one cache entry per candidate pair, in pair order (no allocation atomic; KNextRound publishes the count)
		uint32_t entry = *first_ptr + k;
		if (entry >= w.max_body_pairs)
		{
			set_error(w, B2J_ERR_BODY_PAIR_CACHE_FULL);
			return;
		}

		Q4 r1 = to_q4(w.rotation[cb1]), r2 = to_q4(w.rotation[cb2]);
		V3 x1 = to_v3(w.position[cb1]), x2 = to_v3(w.position[cb2]);
		Q4 inv_r1 = q4_conj(r1);
		V3 delta_position = rotate(inv_r1, x2 - x1);
		Q4 delta_rotation = inv_r1 * r2;

		CachedPair &out = w.write_cache.pairs[entry];
		out.body1 = id1; out.body2 = id2;
		out.slot1 = cb1; out.slot2 = cb2;
		out.first_manifold = 0; out.num_manifolds = 0;

		// the cached-pair work list is indexed like the cache entries (no allocation atomic): invalid unless the pair is served from the cache
		CachedItem cached_item; cached_item.pair_entry = 0xffffffffu; cached_item.old_pair = 0xffffffffu;
		uint32_t old = 0xffffffffu;
		bool handled = false;
		if (w.settings.use_body_pair_contact_cache)
			old = pair_table_find(w, w.read_cache, cb1, cb2);
		// The table is keyed by body SLOTS; the reference keys the cache by the full BodyID (index + sequence number, BodyPair hash):
		// an entry left behind by a destroyed body whose slot was reused is a miss, for the pair test and for the manifold lookup
		if (old != 0xffffffffu && (w.read_cache.pairs[old].body1 != id1 || w.read_cache.pairs[old].body2 != id2))
			old = 0xffffffffu;
		if (old != 0xffffffffu && !((i1.flags | i2.flags) & B2J_BODY_INVALIDATE_CACHE))
		{
			const CachedPair &in = w.read_cache.pairs[old];
			V3 old_dp = v3_load(in.dpos);
			if (!(length_sq(delta_position - old_dp) > w.settings.body_pair_cache_max_delta_position_sq))
			{
				Q4 old_dr = q4_from_xyz(v3_load(in.drot));
				if (!(fabs_(q4_dot(delta_rotation, old_dr)) < w.settings.body_pair_cache_cos_max_delta_rotation_div2))
				{
					handled = true;
					// memcpy of the old CachedBodyPair: the deltas are NOT refreshed (ContactConstraintManager.cpp:1059)
					for (int i = 0; i < 3; ++i) { out.dpos[i] = in.dpos[i]; out.drot[i] = in.drot[i]; }
					atomic_add(&w.counters->num_pairs_from_cache, 1u);
					if (in.num_manifolds != 0)
					{
						cached_item.pair_entry = entry; cached_item.old_pair = old;
					}
				}
			}
		}
		c.cached[entry] = cached_item;
		if (!handled)
		{
			v3_store(delta_position, out.dpos);
			v3_store(q4_xyz(q4_ensure_w_positive(delta_rotation)), out.drot);
			CollideItem item; item.b1 = b1; item.b2 = b2; item.pair_entry = entry; item.old_pair = old;
			uint32_t kind1 = w.shapes[i1.shape].kind, kind2 = w.shapes[i2.shape].kind;
			// pairs that collect several hits per pair (mesh triangles, compound sub shapes) share the collector kernel (b2j_mesh.h)
			if (kind1 == B2J_SHAPE_MESH || kind2 == B2J_SHAPE_MESH || kind1 == B2J_SHAPE_COMPOUND || kind2 == B2J_SHAPE_COMPOUND)
				c.collide_mesh[atomic_add(&w.counters->num_collide_mesh, 1u)] = item;
			else
				c.collide_convex[atomic_add(&w.counters->num_collide_convex, 1u)] = item;
		}
		pair_table_insert(w, w.write_cache, cb1, cb2, entry);
	}
};

// ---- helpers shared by the cached and the collide path ---------------------------------------------------------
B2J_D void wake_body(const DWorld &w, const NarrowCtx &c, uint32_t slot)
{
	if (atomic_exch(&c.woken_flag[slot], 1u) == 0u)
		c.woken_list[atomic_add(&w.counters->num_woken, 1u)] = slot;
}

B2J_D void emit_event(const DWorld &w, const NarrowCtx &c, uint32_t kind, uint32_t id1, uint32_t id2, uint32_t sub1, uint32_t sub2,
	V3 base_offset, V3 normal, float depth, const V3 *p1, const V3 *p2, int n)
{
	if (c.events == nullptr)
		return;
	uint32_t e = atomic_add(&w.counters->num_events, 1u);
	if (e >= c.max_events)
		return;
	b2j_contact_event &ev = c.events[e];
	ev.kind = kind; ev.body1 = id1; ev.body2 = id2; ev.sub_shape1 = sub1; ev.sub_shape2 = sub2; ev.num_points = (uint32_t)n;
	v3_store(base_offset, ev.base_offset);
	v3_store(normal, ev.normal);
	ev.penetration_depth = depth;
	for (int i = 0; i < 4; ++i)
	{
		v3_store(i < n? p1[i] : v3_zero(), ev.points1[i]);
		v3_store(i < n? p2[i] : v3_zero(), ev.points2[i]);
	}
}

// Registers a constraint for manifold m between cache-ordered bodies cb1 (id1) / cb2 (id2): wake up + ConstraintSrc.
// (CreateConstraint ContactConstraintManager.cpp:767-841 minus the constraint itself.)
B2J_D bool register_constraint(const DWorld &w, const NarrowCtx &c, uint32_t m, uint32_t cb1, uint32_t cb2, const BodyInfo &i1, const BodyInfo &i2, uint64_t key_hash, int num_points)
{
	bool sensor = ((i1.flags | i2.flags) & B2J_BODY_SENSOR) != 0;
	bool dyn1 = i1.motion_type == B2J_MOTION_DYNAMIC, dyn2 = i2.motion_type == B2J_MOTION_DYNAMIC;
	if (sensor || !(dyn1 || dyn2))
		return false;
	uint32_t ci = atomic_add(&w.counters->num_constraints, 1u);
	if (ci >= w.max_constraints)
	{
		set_error(w, B2J_ERR_CONTACT_CONSTRAINTS_FULL);
		return false;
	}
	if (dyn1 && w.active_index[cb1] == B2J_INACTIVE_INDEX) wake_body(w, c, cb1);
	if (dyn2 && w.active_index[cb2] == B2J_INACTIVE_INDEX) wake_body(w, c, cb2);
	ConstraintSrc s;
	s.sort_key = key_hash; s.manifold = m; s.b1 = cb1; s.b2 = cb2;
	c.con_src[ci] = s;
	atomic_add(&w.counters->num_contact_points, (uint32_t)num_points);
	return true;
}

// ---- KCopyCached: GetContactsFromCache for pairs whose relative pose did not change ----------------------------
struct KCopyCached
{
	DWorld w; NarrowCtx c;
	const uint32_t *first_ptr; // device: first pair of this round
	B2J_D void operator()(uint32_t k) const
	{
		CachedItem item = c.cached[*first_ptr + k];
		if (item.pair_entry == 0xffffffffu)
			return;
		const CachedPair &in = w.read_cache.pairs[item.old_pair];
		CachedPair &out = w.write_cache.pairs[item.pair_entry];
		uint32_t n = in.num_manifolds;
		uint32_t base = atomic_add(w.write_cache.num_manifolds, n);
		if (base + n > w.max_constraints)
		{
			set_error(w, B2J_ERR_MANIFOLD_CACHE_FULL);
			return;
		}
		uint32_t cb1 = in.slot1, cb2 = in.slot2;
		BodyInfo i1 = w.info[cb1], i2 = w.info[cb2];
		for (uint32_t j = 0; j < n; ++j)
		{
			CachedManifold &src = w.read_cache.manifolds[in.first_manifold + j];
			CachedManifold &dst = w.write_cache.manifolds[base + j];
			dst = src;
			dst.flags = MANIFOLD_FROM_CACHE;
			src.flags |= MANIFOLD_PERSISTED;
			uint64_t hash = hash_sub_shape_id_pair(in.body1, src.sub1, in.body2, src.sub2);
			if (c.events != nullptr)
			{
				// OnContactPersisted with the manifold reconstructed from the cache (ContactConstraintManager.cpp:890-913)
				Q4 q1 = to_q4(w.rotation[cb1]), q2 = to_q4(w.rotation[cb2]);
				V3 x1 = to_v3(w.position[cb1]), x2 = to_v3(w.position[cb2]);
				M33 r1 = m33_rotation(q1);
				Xf local2 = xf(m33_rotation(q2), x2 + (-x1));
				V3 wn = normalized(mul(local2.r, v3_load(src.normal)));
				V3 p1[4], p2[4];
				float depth = -FLT_MAX;
				for (int i = 0; i < src.num_points; ++i)
				{
					p1[i] = mul(r1, v3_load(src.p1[i]));
					p2[i] = mul(local2, v3_load(src.p2[i]));
					depth = fmax_(depth, dot(p1[i] - p2[i], wn));
				}
				emit_event(w, c, B2J_EVENT_CONTACT_PERSISTED, in.body1, in.body2, src.sub1, src.sub2, x1, wn, depth, p1, p2, src.num_points);
			}
			if (!register_constraint(w, c, base + j, cb1, cb2, i1, i2, hash, src.num_points))
			{
				// no constraint: the cached lambdas are kept as they are (memcpy semantics)
			}
		}
		out.first_manifold = base;
		out.num_manifolds = n;
	}
};

// A manifold in the collision space of body 1 (points relative to its centre of mass)
struct ManifoldOut
{
	V3 normal;           // world space, normalised
	float depth;
	uint32_t sub1, sub2;
	int n;
	V3 p1[4], p2[4];
};

// AddContactConstraint + TemplatedAddContactConstraint for all manifolds of one pair. b1/b2: collision order bodies,
// base_offset = centre of mass of b1.
B2J_D void add_manifolds(const DWorld &w, const NarrowCtx &c, const CollideItem &item, ManifoldOut *mans, int count)
{
	if (count == 0)
		return;
	uint32_t base = atomic_add(w.write_cache.num_manifolds, (uint32_t)count);
	if (base + (uint32_t)count > w.max_constraints)
	{
		set_error(w, B2J_ERR_MANIFOLD_CACHE_FULL);
		return;
	}
	uint32_t b1 = item.b1, b2 = item.b2;
	BodyInfo i1 = w.info[b1], i2 = w.info[b2];
	V3 base_offset = to_v3(w.position[b1]);
	bool swap = i2.id < i1.id;
	uint32_t cb1 = swap? b2 : b1, cb2 = swap? b1 : b2;
	BodyInfo ci1 = swap? i2 : i1, ci2 = swap? i1 : i2;
	Xf inv1 = xf_inverse_rotation_translation(to_q4(w.rotation[cb1]), to_v3(w.position[cb1]));
	Xf inv2 = xf_inverse_rotation_translation(to_q4(w.rotation[cb2]), to_v3(w.position[cb2]));
	const CachedPair *old_pair = item.old_pair != 0xffffffffu? &w.read_cache.pairs[item.old_pair] : nullptr;
	bool makes_constraint = !((ci1.flags | ci2.flags) & B2J_BODY_SENSOR) && (ci1.motion_type == B2J_MOTION_DYNAMIC || ci2.motion_type == B2J_MOTION_DYNAMIC);

	for (int mi = 0; mi < count; ++mi)
	{
		ManifoldOut &m = mans[mi];
		// SwapShapes when body 2 has the lower id
		V3 normal = swap? -m.normal : m.normal;
		uint32_t sub1 = swap? m.sub2 : m.sub1, sub2 = swap? m.sub1 : m.sub2;
		const V3 *rp1 = swap? m.p2 : m.p1;
		const V3 *rp2 = swap? m.p1 : m.p2;

		uint64_t key_hash = hash_sub_shape_id_pair(ci1.id, sub1, ci2.id, sub2);
		uint32_t mslot = base + (uint32_t)mi;
		CachedManifold &nm = w.write_cache.manifolds[mslot];
		nm.body1 = ci1.id; nm.body2 = ci2.id; nm.sub1 = sub1; nm.sub2 = sub2;
		nm.num_points = (uint16_t)m.n;
		nm.flags = 0;
		v3_store(normalized(mul(inv2.r, normal)), nm.normal);

		// old manifold with the same SubShapeIDPair (mReadCache->Find(key))
		CachedManifold *old_m = nullptr;
		if (old_pair != nullptr)
			for (uint32_t j = 0; j < old_pair->num_manifolds; ++j)
			{
				CachedManifold &cand = w.read_cache.manifolds[old_pair->first_manifold + j];
				if (cand.sub1 == sub1 && cand.sub2 == sub2) { old_m = &cand; break; }
			}
		if (old_m != nullptr)
			old_m->flags |= MANIFOLD_PERSISTED;
		emit_event(w, c, old_m != nullptr? B2J_EVENT_CONTACT_PERSISTED : B2J_EVENT_CONTACT_ADDED, ci1.id, ci2.id, sub1, sub2, base_offset, normal, m.depth, rp1, rp2, m.n);

		ManifoldWS &ws = c.man_ws[mslot];
		v3_store(normal, ws.normal);
		for (int i = 0; i < m.n; ++i)
		{
			V3 p1_ws = base_offset + rp1[i];
			V3 p2_ws = base_offset + rp2[i];
			v3_store(p1_ws, ws.p1[i]);
			v3_store(p2_ws, ws.p2[i]);
			V3 p1_ls = mul(inv1, p1_ws);
			V3 p2_ls = mul(inv2, p2_ws);
			v3_store(p1_ls, nm.p1[i]);
			v3_store(p2_ls, nm.p2[i]);
			float lambda = 0.0f;
			if (makes_constraint && old_m != nullptr)
				for (int j = 0; j < old_m->num_points; ++j)
					if (is_close(v3_load(old_m->p1[j]), p1_ls, w.settings.contact_point_preserve_lambda_max_dist_sq)
						&& is_close(v3_load(old_m->p2[j]), p2_ls, w.settings.contact_point_preserve_lambda_max_dist_sq))
					{
						lambda = old_m->lambda[j];
						break;
					}
			nm.lambda[i] = lambda;
		}
		for (int i = m.n; i < 4; ++i)
		{
			nm.lambda[i] = 0.0f;
			for (int k = 0; k < 3; ++k) { nm.p1[i][k] = 0.0f; nm.p2[i][k] = 0.0f; }
		}
		if (makes_constraint && old_m != nullptr)
		{
			nm.friction_lambda[0] = old_m->friction_lambda[0];
			nm.friction_lambda[1] = old_m->friction_lambda[1];
			nm.angular_lambda = old_m->angular_lambda;
		}
		else
		{
			nm.friction_lambda[0] = nm.friction_lambda[1] = 0.0f;
			nm.angular_lambda = 0.0f;
		}
		if (makes_constraint)
			register_constraint(w, c, mslot, cb1, cb2, ci1, ci2, key_hash, m.n);
	}
	CachedPair &out = w.write_cache.pairs[item.pair_entry];
	out.first_manifold = base;
	out.num_manifolds = (uint32_t)count;
}

// Result of a convex vs convex test in the space of body 1 (CollideShapeResult relative to body 1's centre of mass)
struct ConvexHit
{
	V3 point1, point2, axis;   // world orientation, relative to body 1 COM
	float depth;
};

// Everything of sCollideConvexVsConvex after the penetration depth is known: world space conversion + supporting faces +
// (ProcessBodyPair) manifold creation for a single hit. transform1 = R(q1), transform2 = T2 - x1 (PhysicsSystem.cpp:1110-1112).
B2J_D void finish_convex_pair(const DWorld &w, const NarrowCtx &c, const CollideItem &item, const Xf &transform1, const Xf &transform2, const Xf &transform_2_to_1,
	V3 point1, V3 point2, V3 penetration_axis, float max_separation_distance)
{
	float penetration_depth = length(point2 - point1) - max_separation_distance;
	// collector early out fraction is FLT_MAX: -depth >= FLT_MAX never true for finite depth
	if (-penetration_depth >= FLT_MAX)
		return;
	float penetration_axis_len = length(penetration_axis);
	if (penetration_axis_len > 0.0f)
		point1 -= penetration_axis * (max_separation_distance / penetration_axis_len);
	point1 = mul(transform1, point1);
	point2 = mul(transform1, point2);
	V3 axis_world = mul(transform1.r, penetration_axis);

	BodyInfo i1 = w.info[item.b1], i2 = w.info[item.b2];
	const ShapeDesc &s1 = w.shapes[i1.shape], &s2 = w.shapes[i2.shape];
	V3 face1[MAX_FACE_VERTS], face2[MAX_FACE_VERTS];
	int n1 = supporting_face(w, s1, -penetration_axis, transform1, face1);
	int n2 = supporting_face(w, s2, mul_transposed(transform_2_to_1.r, penetration_axis), transform2, face2);

	V3 pts1[MAX_MANIFOLD_POINTS], pts2[MAX_MANIFOLD_POINTS], scratch[3 * MAX_CLIP_VERTS];
	int num = 0;
	bool reduction = w.settings.use_manifold_reduction && (i1.flags & B2J_BODY_USE_MANIFOLD_REDUCTION) && (i2.flags & B2J_BODY_USE_MANIFOLD_REDUCTION);
	V3 normal = normalized(axis_world);
	manifold_between_two_faces(point1, point2, axis_world, w.settings.speculative_contact_distance + w.settings.manifold_tolerance, face1, n1, face2, n2, pts1, pts2, num, scratch);
	if (reduction)
	{
		// ReductionCollideShapeCollector: prune at > 32 against the first normal, then the summed normal is normalised again
		if (num > 32)
			prune_contact_points(normal, pts1, pts2, num, scratch);
		normal = normalized(normal);
	}
	if (num > 4)
		prune_contact_points(normal, pts1, pts2, num, scratch);

	ManifoldOut m;
	m.normal = normal; m.depth = penetration_depth; m.sub1 = 0xffffffffu; m.sub2 = 0xffffffffu; m.n = num;
	for (int i = 0; i < num; ++i) { m.p1[i] = pts1[i]; m.p2[i] = pts2[i]; }
	add_manifolds(w, c, item, &m, 1);
}

struct ConvexPairSetup
{
	Xf transform1, transform2, transform_2_to_1;
	float max_separation_distance;
};

B2J_D ConvexPairSetup convex_pair_setup(const DWorld &w, const CollideItem &item)
{
	ConvexPairSetup s;
	BodyInfo i1 = w.info[item.b1], i2 = w.info[item.b2];
	V3 x1 = to_v3(w.position[item.b1]), x2 = to_v3(w.position[item.b2]);
	Q4 q1 = to_q4(w.rotation[item.b1]), q2 = to_q4(w.rotation[item.b2]);
	s.transform1 = shape_transform(w.shapes[i1.shape], xf(m33_rotation(q1), v3_zero()));
	s.transform2 = shape_transform(w.shapes[i2.shape], xf(m33_rotation(q2), x2 + (-x1))); // GetCenterOfMassTransform().PostTranslated(-offset)
	// inverse_transform1 = transform1.InversedRotationTranslation(); transform_2_to_1 = inverse_transform1 * transform2
	M33 rt = transposed(s.transform1.r);
	Xf inv1 = xf(rt, -mul(rt, s.transform1.t));
	s.transform_2_to_1 = mul(inv1, s.transform2);
	s.max_separation_distance = ((i1.flags | i2.flags) & B2J_BODY_SENSOR)? 0.0f : w.settings.speculative_contact_distance;
	return s;
}

// batch groups: sort key of a queued convex pair = (body 1 inside its world, body 2 inside its world); the stable sort keeps the
// worlds of one pair next to each other
struct KCollideKeys
{
	DWorld w; NarrowCtx c; uint32_t *keys, *vals; uint32_t bits, invalid_key;
	B2J_D void operator()(uint32_t k) const
	{
		vals[k] = k;
		if (k >= w.counters->num_collide_convex) { keys[k] = invalid_key; return; } // past the end of the queue: sorts behind every real entry
		CollideItem item = c.collide_convex[k];
		if (w.world_stride != 0)
			keys[k] = ((item.b1 % w.world_stride) << bits) | (item.b2 % w.world_stride);
		else
			// one big world: group the pairs by the two shape types, so that a warp runs one combination of support functions
			keys[k] = (w.shapes[w.info[item.b1].shape].kind << 3) | w.shapes[w.info[item.b2].shape].kind;
	}
};

// ---- KCollideConvex: OBB pre-test + GJK; queues shallow hits for KFinishPairs and deep ones for EPA ------------------
// Thread per pair with the GJK loop in lockstep (all 32 lanes call run(), valid = lane has a pair).
struct KCollideConvex
{
	DWorld w; NarrowCtx c;
	B2J_D void run(uint32_t k, bool valid) const
	{
		bool alive = valid;
		CollideItem item = {};
		ConvexPairSetup s = {};
		ConvexSupport a_excl = {};
		TransformedSupport b_excl = {};
		if (alive)
		{
			item = c.collide_convex[k < c.collide_order_n? c.collide_order[k] : k];
			s = convex_pair_setup(w, item);
			const ShapeDesc &s1 = w.shapes[w.info[item.b1].shape], &s2 = w.shapes[w.info[item.b2].shape];
			V3 bb1_min = s1.local_min - v3_rep(s.max_separation_distance), bb1_max = s1.local_max + v3_rep(s.max_separation_distance);
			if (!obb_vs_aabb(s.transform_2_to_1, s2.local_min, s2.local_max, bb1_min, bb1_max))
				alive = false;
			else
			{
				a_excl = make_support(w, s1, SUPPORT_EXCLUDE_CONVEX_RADIUS);
				b_excl = make_transformed(s.transform_2_to_1, make_support(w, s2, SUPPORT_EXCLUDE_CONVEX_RADIUS));
			}
		}
		V3 penetration_axis = s.transform_2_to_1.t;
		if (is_near_zero(penetration_axis))
			penetration_axis = v3(1.0f, 0.0f, 0.0f);
		GjkSimplex simplex;
		V3 point1 = v3_zero(), point2 = v3_zero();
		int status = pen_depth_step_gjk<true>(simplex, a_excl, a_excl.convex_radius + s.max_separation_distance, b_excl, b_excl.s.convex_radius,
			1.0e-4f /* cDefaultCollisionTolerance */, penetration_axis, point1, point2, alive);
		if (!alive || status == PEN_NOT_COLLIDING)
			return;
		if (status == PEN_INDETERMINATE)
		{
			uint32_t e = atomic_add(&w.counters->num_epa, 1u);
			if (e < c.max_epa)
				c.epa[e].c = item;
			return;
		}
		// supporting faces / clipping / manifold: KFinishPairs (keeps this kernel small and its lanes converged)
		EpaResult &r = c.epa_results[atomic_add(c.num_epa_results, 1u)];
		r.c = item;
		v3_store(point1, r.point1); v3_store(point2, r.point2); v3_store(penetration_axis, r.axis);
		r.max_separation_distance = s.max_separation_distance;
	}
};

// ---- KCollideEpa: EPA for the deep pairs, in two capacity tiers ------------------------------------------------------------
// EPA's cost is decided by the pair: most finish within a dozen support points on a hull of < 24 triangles; box faces resting on box
// faces under convex radius rounding build hulls of > 64 triangles and often need > 48 points, and a few per thousand run to the
// reference's limit of 128 points (~124 iterations). A run that exceeds the PHYSICAL capacity of its tier is discarded and repeated
// from its GJK simplex by the next tier (results only ever come from a run that fitted):
//   tier 0: EpaStorageSmall (2 KB per lane),  thread per pair, lanes in lockstep, reads c.epa -> overflow to c.epa_overflow
//   tier 1: EpaStorageFull  (21 KB per lane), the same over c.epa_overflow; can never overflow
// Measured and NOT kept: a point capped middle tier followed by one pair per warp with the hull in SHARED memory (lane 0
// working, kLockstep = false). A 124 iteration pair takes 0.85 ms there instead of ~6 ms, but only 8 pairs per SM run at a time
// (3.5 M pairs/s against 33 M pairs/s for the lockstep form): -0.6 ms on the 1 M body pile, +0.14 ms on the single Pyramid, +5 ms
// per step on a batch at impact unless the route is chosen per step on the device.
template <class Storage, int kTier, bool kLockstep> struct KCollideEpa
{
	DWorld w; NarrowCtx c;
	// kLockstep: all 32 lanes of the warp call run() together (valid = lane has a pair), the GJK / EPA loops run in lockstep
	B2J_D void run(uint32_t k, bool valid, uint32_t slot, Storage &storage) const
	{
		(void)slot;
		EpaScratch scratch = storage.view();
		bool alive = valid;
		CollideItem item = {};
		ConvexPairSetup s = {};
		AddRadiusSupport a_incl = {};
		TransformedSupport b_incl = {};
		ConvexSupport a_excl = {};
		TransformedSupport b_excl = {};
		float max_separation_distance = 0.0f;
		if (alive)
		{
			const EpaItem &ei = kTier == 1? c.epa_overflow[k] : c.epa[k];
			item = ei.c;
			s = convex_pair_setup(w, item);
			const ShapeDesc &s1 = w.shapes[w.info[item.b1].shape], &s2 = w.shapes[w.info[item.b2].shape];
			max_separation_distance = fmin_(s.max_separation_distance, 1.0f);
			a_incl.s = make_support(w, s1, SUPPORT_INCLUDE_CONVEX_RADIUS);
			a_incl.radius = max_separation_distance;
			b_incl = make_transformed(s.transform_2_to_1, make_support(w, s2, SUPPORT_INCLUDE_CONVEX_RADIUS));
			a_excl = make_support(w, s1, SUPPORT_EXCLUDE_CONVEX_RADIUS);
			b_excl = make_transformed(s.transform_2_to_1, make_support(w, s2, SUPPORT_EXCLUDE_CONVEX_RADIUS));
		}
		// same GJK step as KCollideConvex (bit identical): yields the simplex and the initial axis EPA starts from
		V3 penetration_axis = s.transform_2_to_1.t, point1 = v3_zero(), point2 = v3_zero();
		if (is_near_zero(penetration_axis))
			penetration_axis = v3(1.0f, 0.0f, 0.0f);
		GjkSimplex simplex;
		if (pen_depth_step_gjk<kLockstep>(simplex, a_excl, a_excl.convex_radius + s.max_separation_distance, b_excl, b_excl.s.convex_radius, 1.0e-4f, penetration_axis, point1, point2, alive) != PEN_INDETERMINATE)
			alive = false;
		bool epa_ok = pen_depth_step_epa<kLockstep>(scratch, simplex, a_incl, b_incl, 1.0e-4f /* cDefaultPenetrationTolerance */, penetration_axis, point1, point2, alive);
		if (kTier == 1 && c.epa_hist != nullptr && alive)
			atomic_add(&c.epa_hist[scratch.num_points < 129? scratch.num_points : 129], 1u);
		if (!epa_ok)
		{
			if (kTier == 0 && alive && scratch.overflow)
				c.epa_overflow[atomic_add(c.num_epa_overflow, 1u)].c = item;
			return;
		}
		// the rest of the pair (supporting faces, clipping, manifold; ~5 KB of thread local arrays) runs in KFinishPairs
		EpaResult &r = c.epa_results[atomic_add(c.num_epa_results, 1u)];
		r.c = item;
		v3_store(point1, r.point1); v3_store(point2, r.point2); v3_store(penetration_axis, r.axis);
		r.max_separation_distance = max_separation_distance;
	}
};
using KCollideEpaSmall = KCollideEpa<EpaStorageSmall, 0, true>;
using KCollideEpaFull = KCollideEpa<EpaStorageFull, 1, true>;

struct KFinishPairs
{
	DWorld w; NarrowCtx c;
	B2J_D void operator()(uint32_t k) const
	{
		const EpaResult &r = c.epa_results[k];
		CollideItem item = r.c;
		ConvexPairSetup s = convex_pair_setup(w, item);
		finish_convex_pair(w, c, item, s.transform1, s.transform2, s.transform_2_to_1, v3_load(r.point1), v3_load(r.point2), v3_load(r.axis), r.max_separation_distance);
	}
};

} // namespace b2j
