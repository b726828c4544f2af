// b2j_solver.h -- islands, solve schedule, contact constraint setup, velocity / position solve, integration, sleeping.
//
// Restates the arithmetic of
//   ContactConstraintManager: CalculateNonPenetrationConstraintProperties (.cpp:39-134), CalculateFrictionConstraintProperties
//     (:140-189), TemplatedGetContactsFromCache (:843-999), TemplatedAddContactConstraint (:1130-1332), sWarmStartConstraint
//     (:1587-1622), sSolveVelocityConstraint (:1683-1767), sStoreAppliedImpulses (:1819-1835), sSolvePositionConstraint (:1880-1939)
//   ContactConstraintPart.h:85-252, AngularFrictionConstraintPart.h:40-138
//   MotionProperties.inl (ApplyForceTorqueAndDragInternal :127-149, gyroscopic :99-125, GetInverseInertiaForRotation :69-85)
//   PhysicsSystem.cpp JobApplyGravity :746-791, JobIntegrateVelocity :1583-1711, CheckSleepAndUpdateBounds :2511-2564
//   Body::AddRotationStep Body.inl:81-112, Body::UpdateSleepStateInternal Body.cpp:145-184
//   IslandBuilder::LinkBodies IslandBuilder.cpp:99-156 (union by lowest index), LargeIslandSplitter::AssignSplit/SplitIsland
//     LargeIslandSplitter.cpp:236-490 (greedy 32 bit masks in sorted order, splits < 32 items merged into the serial split)
//
// B200 design: the reference solves island by island, batch by batch (job graph + spinning). Here every constraint gets a
// PHASE such that running phases 0,1,2,... with all constraints of one phase in parallel is bit-identical to the reference's
// order: small islands (< 128 items) are serial in sort-key order -> phase = dependency depth in that order; large islands use
// the reference's own colouring -> phase = colour, the serial split -> 32 + dependency depth. Constraints are then built
// directly in (phase, sort order) into structure-of-arrays storage so the solver streams them fully coalesced.
#pragma once

#include "b2j_world.h"
#include "b2j_shapes.h"
#include "b2j_narrowphase.h"

namespace b2j {

// ---- constraint storage (solve order) ---------------------------------------------------------------------------
// float4 PLANES, plane major: cp[plane * capacity + i]. A warp reads 512 contiguous bytes of a plane with one LDG.128 (the first
// version used 133 scalar field arrays: 4x the load instructions / requests in flight for the same bytes, latency bound at 28% of HBM).
enum
{
	CP_NORMAL = 0,          // world space normal xyz, combined friction
	CP_MASS,                // inv_m1, inv_m2, dist0, dist1      (dist p = distance of contact point p to the friction point)
	CP_DIST,                // dist2, dist3, angular friction effective mass, angular friction bias
	CP_LAMBDA_PT,           // total lambda of the 4 contact points    (read + written by the solve kernels)
	CP_LAMBDA_FR,           // total lambda friction 1, friction 2, angular friction, unused
	CP_ANG_I1,              // angular friction: inverse inertia * normal of body 1 / 2 (xyz)
	CP_ANG_I2,
	CP_FR0,                 // 2 friction parts of 4 planes (CP_PART_*)
	CP_PT0 = CP_FR0 + 8,    // 4 contact point parts of 4 planes
	CP_LP0 = CP_PT0 + 16,   // per contact point: local point on body 1 xyz, local point on body 2 xyz (position solve)
	CP_NUM = CP_LP0 + 8
};
// planes of one ContactConstraintPart: r1 x axis + effective mass, invI1 (r1 x axis) + bias, r2 x axis, invI2 (r2 x axis)
enum { CP_PART_R1X_EFF = 0, CP_PART_I1_BIAS, CP_PART_R2X, CP_PART_I2 };

// meta bits: num points (3) | type1 (2) << 3 | type2 (2) << 5 | velocity steps << 8 | position steps << 16 | friction parts active << 24
//            | translation DOFs of body 1 (3) << 26 | translation DOFs of body 2 (3) << 29   (what store_vel_state masks with: no
//            gather of the BodyInfo per body and iteration)
//            bit 7: the item is a non contact constraint (b2j_joints.h): header only, `manifold` = constraint index; the contact kernels skip it
enum : uint32_t { META_JOINT = 1u << 7 };
enum : uint32_t { META_LINEAR_FRICTION = 1u << 24, META_ANGULAR_FRICTION = 1u << 25, META_DOFS1_SHIFT = 26, META_DOFS2_SHIFT = 29 };
struct alignas(16) ConstraintHeader { uint32_t b1, b2, manifold, meta; };
struct Constraints
{
	F4 *cp;
	ConstraintHeader *hdr;       // one 16 byte load gives the solve kernels everything their other loads depend on
	uint32_t capacity;
};

struct SolveCtx
{
	Constraints con;
	ConstraintSrc *src;          // unsorted constraint sources
	ManifoldWS *man_ws;
	uint32_t *order;             // [M] sorted position -> src index (sorted by sort key)
	uint32_t *final_pos;         // [M] sorted position -> solve position
	uint32_t *solve_src;         // [M] solve position -> src index
	uint32_t *phase;             // [M] by sorted position
	uint32_t *phase_count;       // [max_phases + 1] histogram / offsets
	uint32_t max_phases;
	// islands (indexed by body slot)
	uint32_t *uf_parent;
	uint32_t *root;              // flattened root per slot
	uint32_t *island_items;      // per root slot: number of constraints
	uint32_t *island_large;      // per root slot: compact large island index + 1, 0 = small
	uint32_t *island_steps;      // per root slot: max vel override | max pos override << 8 | apply default vel << 16 | apply default pos << 17
	uint32_t *island_can_sleep;  // per root slot
	uint32_t *large_color_count; // [num_large * 32]
	// per body adjacency (CSR by slot)
	uint32_t *body_deg, *body_off, *body_fill, *body_cur, *body_mask;
	uint32_t *adj;
	uint32_t num_slots;
	// non contact constraints (b2j_joints.h): null / 0 in worlds without any
	const uint32_t *joint_steps; // [constraint index] velocity steps override | position steps override << 8 | constraint type << 16
	uint32_t *body_nj;           // per body slot: how many of its items are joints (they lead its adjacency list)
	uint32_t *sched_flag;        // [2] remaining flags
	uint32_t *grid_barrier;      // arrival counter of solve_velocity_tma_kernel's grid barrier (zeroed before the launch)
};

B2J_HD F4 &cp_at(const Constraints &c, int plane, uint32_t i) { return c.cp[(size_t)plane * c.capacity + i]; }
// Read of a plane the running kernel never writes (everything but the lambdas in the solve kernels): goes through the non coherent
// path, which also tells the compiler that the lambda stores cannot alias it, so all loads of a constraint issue up front.
B2J_HD F4 cp_ro(const Constraints &c, int plane, uint32_t i)
{
#if defined(__CUDA_ARCH__)
	float4 v = __ldg(reinterpret_cast<const float4 *>(&c.cp[(size_t)plane * c.capacity + i]));
	return f4(v.x, v.y, v.z, v.w);
#else
	return c.cp[(size_t)plane * c.capacity + i];
#endif
}

// ---- body helpers ------------------------------------------------------------------------------------------------
B2J_HD V3 lock_translation(V3 v, uint32_t dofs) { return v3((dofs & 1)? v.x : 0.0f, (dofs & 2)? v.y : 0.0f, (dofs & 4)? v.z : 0.0f); }
B2J_HD V3 lock_angular(V3 v, uint32_t dofs) { return v3((dofs & 8)? v.x : 0.0f, (dofs & 16)? v.y : 0.0f, (dofs & 32)? v.z : 0.0f); }

// MotionProperties::GetInverseInertiaForRotation
B2J_HD M33 inverse_inertia_for_rotation(const M33 &body_rotation, Q4 inertia_rotation, V3 inv_inertia_diag, uint32_t dofs)
{
	M33 rotation = mul(body_rotation, m33_rotation(inertia_rotation));
	M33 rms = m33(inv_inertia_diag.x * rotation.c0, inv_inertia_diag.y * rotation.c1, inv_inertia_diag.z * rotation.c2);
	M33 inv = mul_right_transposed(rotation, rms);
	bool ax = (dofs & 8) != 0, ay = (dofs & 16) != 0, az = (dofs & 32) != 0;
	if (!(ax && ay && az))
	{
		// column j masked by (mask & splat(mask[j]))
		inv.c0 = v3(ax && ax? inv.c0.x : 0.0f, ay && ax? inv.c0.y : 0.0f, az && ax? inv.c0.z : 0.0f);
		inv.c1 = v3(ax && ay? inv.c1.x : 0.0f, ay && ay? inv.c1.y : 0.0f, az && ay? inv.c1.z : 0.0f);
		inv.c2 = v3(ax && az? inv.c2.x : 0.0f, ay && az? inv.c2.y : 0.0f, az && az? inv.c2.z : 0.0f);
	}
	return inv;
}

// MotionProperties::MultiplyWorldSpaceInverseInertiaByVector
B2J_HD V3 multiply_ws_inverse_inertia(Q4 body_rotation, Q4 inertia_rotation, V3 inv_inertia_diag, uint32_t dofs, V3 v_in)
{
	V3 v = lock_angular(v_in, dofs);
	M33 rotation = m33_rotation(body_rotation * inertia_rotation);
	V3 result = mul(rotation, inv_inertia_diag * mul_transposed(rotation, v));
	return lock_angular(result, dofs);
}

B2J_HD void clamp_velocity(V3 &v, float max_v)
{
	float len_sq = length_sq(v);
	if (len_sq > square(max_v))
		v *= max_v / sqrt_(len_sq);
}

// Body::AddRotationStep / SubRotationStep
B2J_HD Q4 add_rotation_step(Q4 rotation, V3 w_dt, bool sub)
{
	float len = length(w_dt);
	if (len > 1.0e-6f)
		return q4_normalized(q4_rotation(w_dt / len, sub? -len : len) * rotation);
	return rotation;
}

// ---- KApplyGravity (JobApplyGravity) -----------------------------------------------------------------------------
struct KApplyGravity
{
	DWorld w; float dt;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		BodyInfo info = w.info[b];
		if (info.motion_type != B2J_MOTION_DYNAMIC)
			return;
		BodyParams p = w.params[b];
		Q4 rot = to_q4(w.rotation[b]);
		V3 v = to_v3(w.linear_velocity[b]), av = to_v3(w.angular_velocity[b]);
		V3 diag = to_v3(w.inv_inertia_diag[b]);
		Q4 irot = to_q4(w.inertia_rotation[b]);
		if (info.flags & B2J_BODY_GYROSCOPIC)
		{
			// MotionProperties::ApplyGyroscopicForceInternal
			V3 denom = v3(diag.x == 0.0f? 1.0f : diag.x, diag.y == 0.0f? 1.0f : diag.y, diag.z == 0.0f? 1.0f : diag.z);
			V3 nom = v3(diag.x == 0.0f? 0.0f : 1.0f, diag.y == 0.0f? 0.0f : 1.0f, diag.z == 0.0f? 0.0f : 1.0f);
			V3 local_inertia = nom / denom;
			Q4 i2w = rot * irot;
			V3 local_av = inverse_rotate(i2w, av);
			V3 local_momentum = local_inertia * local_av;
			V3 new_local_momentum = local_momentum - dt * cross(local_av, local_momentum);
			float nl = length_sq(new_local_momentum);
			new_local_momentum = nl > 0.0f? new_local_momentum * sqrt_(length_sq(local_momentum) / nl) : v3_zero();
			av = rotate(i2w, diag * new_local_momentum);
		}
		// ApplyForceTorqueAndDragInternal
		V3 force = to_v3(w.force[b]), torque = to_v3(w.torque[b]);
		v = lock_translation(v + dt * (p.gravity_factor * w.gravity + p.inv_mass * force), info.allowed_dofs);
		av += dt * multiply_ws_inverse_inertia(rot, irot, diag, info.allowed_dofs, torque);
		v *= fmax_(0.0f, 1.0f - p.linear_damping * dt);
		av *= fmax_(0.0f, 1.0f - p.angular_damping * dt);
		clamp_velocity(v, p.max_linear_velocity);
		clamp_velocity(av, p.max_angular_velocity);
		w.linear_velocity[b] = f4(v);
		w.angular_velocity[b] = f4(av);
	}
};

// ---- islands: union find over body slots, root = lowest slot ----------------------------------------------------------
struct KUfInit
{
	SolveCtx s;
	B2J_D void operator()(uint32_t slot) const
	{
		s.uf_parent[slot] = slot;
		s.island_items[slot] = 0;
		s.island_large[slot] = 0;
		s.island_steps[slot] = 0;
		s.island_can_sleep[slot] = 1;
		s.body_deg[slot] = 0;
		s.body_fill[slot] = 0;
		s.body_cur[slot] = 0;
		s.body_mask[slot] = 0;
		if (s.body_nj != nullptr) s.body_nj[slot] = 0;
	}
};

// find with intermediate pointer jumping (path halving): parents only ever move to smaller slots of the same set, so redirecting
// a node to its grandparent is always valid, racing writers included
B2J_D uint32_t uf_find(uint32_t *parent, uint32_t x)
{
	uint32_t p = volatile_load(&parent[x]);
	while (p != x)
	{
		uint32_t gp = volatile_load(&parent[p]);
		if (gp != p) parent[x] = gp;
		x = p;
		p = gp;
	}
	return x;
}

struct KUfUnion
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t i) const
	{
		const ConstraintSrc &c = s.src[i];
		bool dyn1 = w.info[c.b1].motion_type == B2J_MOTION_DYNAMIC, dyn2 = w.info[c.b2].motion_type == B2J_MOTION_DYNAMIC;
		if (dyn1) atomic_add(&s.body_deg[c.b1], 1u);
		if (dyn2) atomic_add(&s.body_deg[c.b2], 1u);
		if (c.manifold & 0x80000000u)
		{
			// a non contact constraint (TwoBodyConstraint::BuildIslands links like a contact: both bodies dynamic)
			if (dyn1) atomic_add(&s.body_nj[c.b1], 1u);
			if (dyn2) atomic_add(&s.body_nj[c.b2], 1u);
		}
		if (!(dyn1 && dyn2))
			return;
		uint32_t a = c.b1, b = c.b2;
		for (;;)
		{
			a = uf_find(s.uf_parent, a);
			b = uf_find(s.uf_parent, b);
			if (a == b) break;
			if (a < b) { if (atomic_cas(&s.uf_parent[b], b, a) == b) break; }
			else { if (atomic_cas(&s.uf_parent[a], a, b) == a) break; }
		}
	}
};

struct KUfFlatten
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		s.root[b] = uf_find(s.uf_parent, b);
	}
};

// per constraint: island item count + solver step overrides (CalculateSolverSteps over the dynamic bodies of contacts)
struct KIslandCount
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t i) const
	{
		const ConstraintSrc &c = s.src[i];
		BodyInfo i1 = w.info[c.b1], i2 = w.info[c.b2];
		bool dyn1 = i1.motion_type == B2J_MOTION_DYNAMIC, dyn2 = i2.motion_type == B2J_MOTION_DYNAMIC;
		uint32_t r = s.root[dyn1? c.b1 : c.b2];
		atomic_add_matched(s.island_items, r, 1u); // (one giant island = one address for a million constraints: aggregate per warp)
		uint32_t vmax = 0, pmax = 0, flags = 0;
		if (c.manifold & 0x80000000u)
		{
			// CalculateSolverSteps::operator()(const Constraint *): the constraint's own overrides, not its bodies'
			uint32_t o = s.joint_steps[c.manifold & 0x7fffffffu];
			vmax = o & 0xff; pmax = (o >> 8) & 0xff;
			if (vmax == 0) flags |= 1u << 16;
			if (pmax == 0) flags |= 1u << 17;
			dyn1 = false; dyn2 = false;
		}
		if (dyn1) { uint32_t v = i1.steps_override & 15, p = i1.steps_override >> 4; vmax = v; pmax = p; if (v == 0) flags |= 1u << 16; if (p == 0) flags |= 1u << 17; }
		if (dyn2) { uint32_t v = i2.steps_override & 15, p = i2.steps_override >> 4; if (v > vmax) vmax = v; if (p > pmax) pmax = p; if (v == 0) flags |= 1u << 16; if (p == 0) flags |= 1u << 17; }
		// max of two packed bytes: do them separately to keep atomics simple
		uint32_t cur = volatile_load(&s.island_steps[r]);
		for (;;)
		{
			uint32_t cv = cur & 0xff, cp = (cur >> 8) & 0xff;
			uint32_t nv = cv > vmax? cv : vmax, np = cp > pmax? cp : pmax;
			uint32_t nw = nv | (np << 8) | (cur & 0x30000u) | flags;
			if (nw == cur) break;
			uint32_t prev = atomic_cas(&s.island_steps[r], cur, nw);
			if (prev == cur) break;
			cur = prev;
		}
	}
};

// compact index for large islands (>= 128 items, LargeIslandSplitter.h:34)
struct KIslandClassify
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		if (s.root[b] != b)
			return;
		atomic_add(&w.counters->num_islands, 1u);
		if (w.settings.use_large_island_splitter && s.island_items[b] >= 128)
		{
			uint32_t li = atomic_add(&w.counters->num_large_islands, 1u);
			s.island_large[b] = li + 1;
			for (int c = 0; c < 32; ++c) s.large_color_count[li * 32 + c] = 0;
		}
	}
};

// adjacency fill: i = sorted position
struct KAdjFill
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t i) const
	{
		const ConstraintSrc &c = s.src[s.order[i]];
		s.phase[i] = 0xffffffffu;
		if (w.info[c.b1].motion_type == B2J_MOTION_DYNAMIC)
			s.adj[s.body_off[c.b1] + atomic_add(&s.body_fill[c.b1], 1u)] = i;
		if (w.info[c.b2].motion_type == B2J_MOTION_DYNAMIC)
			s.adj[s.body_off[c.b2] + atomic_add(&s.body_fill[c.b2], 1u)] = i;
	}
};

// per active body: sort its adjacency list ascending (insertion sort, lists are short)
struct KAdjSort
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		uint32_t n = s.body_deg[b];
		uint32_t *a = s.adj + s.body_off[b];
		for (uint32_t i = 1; i < n; ++i)
		{
			uint32_t v = a[i];
			uint32_t j = i;
			while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; --j; }
			a[j] = v;
		}
	}
};

// ---- wavefront scheduling ------------------------------------------------------------------------------------------
enum { PHASE_UNSCHEDULED = 0xffffffffu, PHASE_PENDING_SERIAL = 0xfffffffeu };

// Item `cur` of body b's list in the order pass `pass` walks it. The list is stored in solve order (non contact constraints first,
// b2j_joints.h); the colouring of a large island (pass 0) takes the contacts first (LargeIslandSplitter::SplitIsland assigns the
// contacts their splits before the constraints), i.e. reads the list rotated by the body's joint count.
B2J_D bool sched_rotated(const DWorld &w, const SolveCtx &s, uint32_t b, uint32_t pass)
{
	(void)w;
	return s.body_nj != nullptr && pass == 0 && s.body_nj[b] != 0 && s.island_large[s.root[b]] != 0;
}
B2J_D uint32_t sched_item(const SolveCtx &s, uint32_t b, uint32_t cur, uint32_t deg, bool rotated)
{
	if (rotated)
	{
		uint32_t nj = s.body_nj[b], nc = deg - nj;
		cur = cur < nc? cur + nj : cur - nc;
	}
	return s.adj[s.body_off[b] + cur];
}

// decide step: one thread per active body; `pass` 0 = levels / colours, 1 = serial splits of large islands
struct KSchedDecide
{
	DWorld w; SolveCtx s; uint32_t round; uint32_t pass;
	B2J_D void operator()(uint32_t ai) const { decide(w.active[ai]); }
	B2J_D void decide(uint32_t b) const
	{
		uint32_t cur = s.body_cur[b];
		if (cur >= s.body_deg[b])
			return;
		uint32_t i = sched_item(s, b, cur, s.body_deg[b], sched_rotated(w, s, b, pass));
		const ConstraintSrc &c = s.src[s.order[i]];
		bool dyn1 = w.info[c.b1].motion_type == B2J_MOTION_DYNAMIC, dyn2 = w.info[c.b2].motion_type == B2J_MOTION_DYNAMIC;
		uint32_t owner = dyn1? c.b1 : c.b2;
		if (owner != b)
			return;
		uint32_t other = (dyn1 && dyn2)? c.b2 : 0xffffffffu;
		if (other != 0xffffffffu)
		{
			uint32_t oc = s.body_cur[other];
			if (oc >= s.body_deg[other] || sched_item(s, other, oc, s.body_deg[other], sched_rotated(w, s, other, pass)) != i)
				return;
		}
		uint32_t r = s.root[b];
		uint32_t large = s.island_large[r];
		if (pass == 0)
		{
			if (large != 0)
			{
				// LargeIslandSplitter::AssignSplit
				uint32_t m = s.body_mask[b] | (other != 0xffffffffu? s.body_mask[other] : 0u);
				int split = ctz32(~m);
				if (split > 31) split = 31;
				uint32_t bit = 1u << split;
				s.body_mask[b] |= bit;
				if (other != 0xffffffffu) s.body_mask[other] |= bit;
				s.phase[i] = (uint32_t)split;
				atomic_add(&s.large_color_count[(large - 1) * 32 + split], 1u);
			}
			else
				s.phase[i] = round;
		}
		else
			s.phase[i] = 32 + round;
	}
};

// advance step: move every body's cursor past scheduled constraints (pass 1: past everything that is not pending serial)
struct KSchedAdvance
{
	DWorld w; SolveCtx s; uint32_t pass; uint32_t flag_index;
	B2J_D void operator()(uint32_t ai) const
	{
		if (advance(w.active[ai]))
			s.sched_flag[flag_index] = 1;
	}
	// returns true while the body still has unscheduled constraints
	B2J_D bool advance(uint32_t b) const
	{
		uint32_t cur = s.body_cur[b], deg = s.body_deg[b];
		const uint32_t *a = s.adj + s.body_off[b];
		if (pass == 0)
		{
			if (sched_rotated(w, s, b, pass))
				while (cur < deg && s.phase[sched_item(s, b, cur, deg, true)] != PHASE_UNSCHEDULED) ++cur;
			else
				while (cur < deg && s.phase[a[cur]] != PHASE_UNSCHEDULED) ++cur;
		}
		else
			while (cur < deg && s.phase[a[cur]] != PHASE_PENDING_SERIAL) ++cur;
		s.body_cur[b] = cur;
		return cur < deg;
	}
};

// after pass 0: colours of large islands with < 32 items and colour 31 go to the serial split (SplitIsland :375-412)
struct KSchedSerialize
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t i) const { serialize(i); }
	// returns true if the constraint went to the serial split
	B2J_D bool serialize(uint32_t i) const
	{
		const ConstraintSrc &c = s.src[s.order[i]];
		bool dyn1 = w.info[c.b1].motion_type == B2J_MOTION_DYNAMIC;
		uint32_t large = s.island_large[s.root[dyn1? c.b1 : c.b2]];
		if (large == 0)
			return false;
		uint32_t color = s.phase[i];
		if (color == 31 || s.large_color_count[(large - 1) * 32 + color] < 32)
		{
			s.phase[i] = PHASE_PENDING_SERIAL;
			return true;
		}
		return false;
	}
};

#if !defined(B2J_HOSTSIM) && defined(__CUDACC__)
// Big single worlds: the same schedule as ONE cooperative launch, rounds separated by grid wide barriers (no host round trip every
// few rounds, ~200 launches less per step: the step time of the 1M body pile no longer depends on host scheduling jitter).
// sched_flag[r % 3] = "some body still has unscheduled constraints after round r".
struct KSchedGrid { }; // (profiling category)
__global__ void __launch_bounds__(256) sched_grid_kernel(const DWorld w, const SolveCtx s, uint32_t na, uint32_t num_constraints)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	volatile uint32_t *flag = s.sched_flag;
	for (uint32_t pass = 0; pass < 2; ++pass)
	{
		KSchedDecide decide; decide.w = w; decide.s = s; decide.pass = pass; decide.round = 0;
		KSchedAdvance advance; advance.w = w; advance.s = s; advance.pass = pass; advance.flag_index = 0;
		if (pass == 1)
		{
			KSchedSerialize serialize; serialize.w = w; serialize.s = s;
			if (tid == 0) { flag[0] = 0; flag[1] = 0; flag[2] = 0; }
			grid.sync();
			bool any = false;
			for (uint32_t i = tid; i < num_constraints; i += nt)
				if (serialize.serialize(i)) any = true;
			if (any) flag[0] = 1;
			grid.sync();
			if (flag[0] == 0)
				return; // (uniform)
			for (uint32_t ai = tid; ai < na; ai += nt)
				s.body_cur[w.active[ai]] = 0;
			grid.sync();
		}
		if (tid == 0) { flag[0] = 0; flag[1] = 0; flag[2] = 0; }
		grid.sync();
		bool remaining = false;
		for (uint32_t ai = tid; ai < na; ai += nt)
			if (advance.advance(w.active[ai])) remaining = true;
		if (remaining) flag[2] = 1;      // plays the role of "round -1"
		grid.sync();
		bool more = flag[2] != 0;
		for (uint32_t round = 0; more; ++round)
		{
			if (tid == 0) flag[(round + 1) % 3] = 0;
			decide.round = round;
			for (uint32_t ai = tid; ai < na; ai += nt)
				decide.decide(w.active[ai]);
			grid.sync();
			remaining = false;
			for (uint32_t ai = tid; ai < na; ai += nt)
				if (advance.advance(w.active[ai])) remaining = true;
			if (remaining) flag[round % 3] = 1;
			grid.sync();
			more = flag[round % 3] != 0;
		}
	}
}

// The whole wavefront schedule of ONE world (a batch group: block = world; a small single world: one block) in one launch: the
// rounds are separated by __syncthreads instead of kernel launches + host checks (54 rounds x 2 launches + 7 host round trips per
// step for a Pyramid). Same decide / advance / serialize bodies as the multi launch path, which remains for big single worlds.
struct KSchedBlock { }; // (profiling category)
__global__ void __launch_bounds__(1024) sched_block_kernel(const DWorld w, const SolveCtx s, uint32_t slots_per_block)
{
	uint32_t first = blockIdx.x * slots_per_block, end = first + slots_per_block;
	if (end > s.num_slots) end = s.num_slots;
	__shared__ uint32_t any_serial;
	if (threadIdx.x == 0) any_serial = 0;
	for (uint32_t pass = 0; pass < 2; ++pass)
	{
		KSchedDecide decide; decide.w = w; decide.s = s; decide.pass = pass; decide.round = 0;
		KSchedAdvance advance; advance.w = w; advance.s = s; advance.pass = pass; advance.flag_index = 0;
		if (pass == 1)
		{
			// colours with < 32 items and colour 31 -> serial split; every constraint is visited by its owner body
			KSchedSerialize serialize; serialize.w = w; serialize.s = s;
			for (uint32_t b = first + threadIdx.x; b < end; b += blockDim.x)
			{
				uint32_t deg = s.body_deg[b];
				const uint32_t *a = s.adj + s.body_off[b];
				for (uint32_t j = 0; j < deg; ++j)
				{
					uint32_t i = a[j];
					const ConstraintSrc &c = s.src[s.order[i]];
					uint32_t owner = w.info[c.b1].motion_type == B2J_MOTION_DYNAMIC? c.b1 : c.b2;
					if (owner == b && serialize.serialize(i))
						any_serial = 1;
				}
			}
			__syncthreads();
			if (any_serial == 0)
				return; // (uniform: shared flag read after the barrier)
			for (uint32_t b = first + threadIdx.x; b < end; b += blockDim.x)
				s.body_cur[b] = 0;
			__syncthreads();
		}
		int remaining = 0;
		for (uint32_t b = first + threadIdx.x; b < end; b += blockDim.x)
			if (advance.advance(b)) remaining = 1;
		remaining = __syncthreads_or(remaining);
		for (uint32_t round = 0; remaining != 0; ++round)
		{
			decide.round = round;
			for (uint32_t b = first + threadIdx.x; b < end; b += blockDim.x)
				decide.decide(b);
			__syncthreads();
			remaining = 0;
			for (uint32_t b = first + threadIdx.x; b < end; b += blockDim.x)
				if (advance.advance(b)) remaining = 1;
			remaining = __syncthreads_or(remaining);
			if (round > 1000000u) break; // cannot happen: every round schedules at least one constraint of an unfinished island
		}
	}
}
#endif

// Debug check of the solve schedule (the reference asserts split disjointness in LargeIslandSplitter): all constraints of one
// phase run in parallel, so no dynamic body may appear in two constraints of the same phase. One thread per active body over its
// adjacency list (sorted positions) of the last step; counts the (body, phase) conflicts.
struct KCheckSchedule
{
	DWorld w; SolveCtx s; uint32_t *conflicts;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		if (w.info[b].motion_type != B2J_MOTION_DYNAMIC) return;
		uint32_t deg = s.body_deg[b];
		const uint32_t *a = s.adj + s.body_off[b];
		for (uint32_t i = 1; i < deg; ++i)
			for (uint32_t j = 0; j < i; ++j)
				if (s.phase[a[i]] == s.phase[a[j]])
					atomic_add(conflicts, 1u);
	}
};

struct KSchedResetCursors
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t ai) const { s.body_cur[w.active[ai]] = 0; }
};

// Phases -> solve order. The constraints are radix sorted by phase (stable: sort key order inside a phase, deterministic layout,
// no histogram atomics); KPhaseClamp bounds the key and finds the phase count, KPhasePlace scatters and writes the phase offsets.
struct KPhaseClamp
{
	DWorld w; SolveCtx s;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t p = s.phase[i];
		if (p >= s.max_phases) { p = s.max_phases - 1; s.phase[i] = p; atomic_or(&w.counters->error_bits, 0x100u); }
		if (p + 1 > *(volatile const uint32_t *)&w.counters->num_phases) // almost always false: keeps the atomic off the hot path
			atomic_max(&w.counters->num_phases, p + 1);
	}
};

struct KPhasePlace
{
	SolveCtx s; const uint32_t *sorted_phase, *sorted_idx; uint32_t n;
	B2J_D void operator()(uint32_t pos) const
	{
		uint32_t i = sorted_idx[pos];
		s.final_pos[i] = pos;
		s.solve_src[pos] = s.order[i];
		// phase_count[q] = first position of phase q (empty phases included), phase_count[last + 1 ...] = n
		uint32_t p = sorted_phase[pos];
		uint32_t first = pos == 0? 0 : sorted_phase[pos - 1] + 1;
		for (uint32_t q = first; q <= p; ++q)
			s.phase_count[q] = pos;
		if (pos == n - 1)
			for (uint32_t q = p + 1; q <= s.max_phases; ++q)
				s.phase_count[q] = n;
	}
};

// ---- constraint setup in solve order ------------------------------------------------------------------------------

struct BodyKin
{
	V3 x, v, av;
	Q4 q;
	uint32_t type, dofs;
	float inv_mass, gravity_factor;
	V3 force;
	M33 inv_i;
};

B2J_D BodyKin load_body_kin(const DWorld &w, uint32_t b)
{
	BodyKin k;
	BodyInfo info = w.info[b];
	k.type = info.motion_type;
	k.dofs = info.allowed_dofs;
	k.x = to_v3(w.position[b]);
	k.q = to_q4(w.rotation[b]);
	if (k.type != B2J_MOTION_STATIC)
	{
		k.v = to_v3(w.linear_velocity[b]);
		k.av = to_v3(w.angular_velocity[b]);
	}
	else
	{
		k.v = v3_zero(); k.av = v3_zero();
	}
	BodyParams p = w.params[b];
	k.gravity_factor = p.gravity_factor;
	if (k.type == B2J_MOTION_DYNAMIC)
	{
		k.inv_mass = p.inv_mass;
		k.force = to_v3(w.force[b]);
		k.inv_i = inverse_inertia_for_rotation(m33_rotation(k.q), to_q4(w.inertia_rotation[b]), to_v3(w.inv_inertia_diag[b]), k.dofs);
	}
	else
	{
		k.inv_mass = 0.0f;
		k.force = v3_zero();
		k.inv_i = m33_zero();
	}
	return k;
}

// One ContactConstraintPart (Jolt/Physics/Constraints/ConstraintPart/AxisConstraintPart.h members) held in registers
struct PartRegs { V3 r1x, i1, r2x, i2; float eff, bias, lambda; };

// ContactConstraintPart::CalculateConstraintProperties; r.lambda holds the current total lambda on entry (Deactivate() clears it)
B2J_D void part_calculate(PartRegs &r, uint32_t type1, uint32_t type2, float inv_m1, const M33 &inv_i1, V3 r1,
	float inv_m2, const M33 &inv_i2, V3 r2, V3 axis, float bias)
{
	r.bias = bias;
	r.r1x = v3_zero(); r.i1 = v3_zero(); r.r2x = v3_zero(); r.i2 = v3_zero();
	float inv_effective_mass;
	if (type1 != B2J_MOTION_STATIC)
	{
		r.r1x = cross(r1, axis);
		if (type1 == B2J_MOTION_DYNAMIC)
		{
			r.i1 = mul(inv_i1, r.r1x);
			inv_effective_mass = inv_m1 + dot(r.i1, r.r1x);
		}
		else
			inv_effective_mass = 0.0f;
	}
	else
		inv_effective_mass = 0.0f;
	if (type2 != B2J_MOTION_STATIC)
	{
		r.r2x = cross(r2, axis);
		if (type2 == B2J_MOTION_DYNAMIC)
		{
			r.i2 = mul(inv_i2, r.r2x);
			inv_effective_mass += inv_m2 + dot(r.i2, r.r2x);
		}
	}
	if (inv_effective_mass == 0.0f)
	{
		// Deactivate(): effective mass AND total lambda are cleared
		r.eff = 0.0f;
		r.lambda = 0.0f;
	}
	else
		r.eff = 1.0f / inv_effective_mass;
}

// the planes the solve kernels read for these motion types (the rest is never read); the lambda goes to the lambda planes
B2J_D void part_store(const Constraints &c, int base, uint32_t i, uint32_t type1, uint32_t type2, const PartRegs &r)
{
	cp_at(c, base + CP_PART_R1X_EFF, i) = f4(r.r1x, r.eff);
	cp_at(c, base + CP_PART_I1_BIAS, i) = f4(r.i1, r.bias);
	if (type2 != B2J_MOTION_STATIC) cp_at(c, base + CP_PART_R2X, i) = f4(r.r2x);
	if (type2 == B2J_MOTION_DYNAMIC) cp_at(c, base + CP_PART_I2, i) = f4(r.i2);
}

// Where the solve kernels read the read-only planes of a constraint from: straight from HBM (per phase launches, small worlds) or
// from the shared memory stage a TMA bulk copy filled (solve_velocity_tma_kernel). Same arithmetic on both.
struct GlobalPlanes
{
	Constraints c; uint32_t i;
	B2J_D F4 ro(int plane) const { return cp_ro(c, plane, i); }
};

// shared memory stage of one warp tile: [slot][lane] float4 (a warp reads one slot with one conflict free LDS.128), slots = the planes
// the velocity solve reads minus the two lambda planes (read + written through registers)
enum { SV_SLOT_NORMAL = 0, SV_SLOT_MASS, SV_SLOT_DIST, SV_SLOT_ANG_I1, SV_SLOT_ANG_I2, SV_SLOT_FR0, SV_SLOT_PT0 = SV_SLOT_FR0 + 8, SV_NUM_SLOTS = SV_SLOT_PT0 + 16 };
B2J_HD int sv_slot_of_plane(int plane)
{
	return plane >= CP_FR0? SV_SLOT_FR0 + (plane - CP_FR0) : (plane == CP_ANG_I1? SV_SLOT_ANG_I1 : (plane == CP_ANG_I2? SV_SLOT_ANG_I2 : plane)); // NORMAL, MASS, DIST = 0, 1, 2
}
B2J_HD int sv_plane_of_slot(int slot)
{
	return slot >= SV_SLOT_FR0? CP_FR0 + (slot - SV_SLOT_FR0) : (slot == SV_SLOT_ANG_I1? CP_ANG_I1 : (slot == SV_SLOT_ANG_I2? CP_ANG_I2 : slot));
}
struct SmemPlanes
{
	const F4 *stage; uint32_t lane;
	B2J_D F4 ro(int plane) const { return stage[sv_slot_of_plane(plane) * 32 + lane]; }
};

// everything but the lambda
template <class Src> B2J_D PartRegs part_load(const Src &src, int base, uint32_t type1, uint32_t type2)
{
	PartRegs r;
	F4 a = src.ro(base + CP_PART_R1X_EFF), b = src.ro(base + CP_PART_I1_BIAS);
	r.r1x = to_v3(a); r.eff = a.w;
	r.i1 = to_v3(b); r.bias = b.w;
	r.r2x = type2 != B2J_MOTION_STATIC? to_v3(src.ro(base + CP_PART_R2X)) : v3_zero();
	r.i2 = type2 == B2J_MOTION_DYNAMIC? to_v3(src.ro(base + CP_PART_I2)) : v3_zero();
	r.lambda = 0.0f;
	(void)type1;
	return r;
}

struct KSetupConstraints
{
	DWorld w; SolveCtx s; float dt;
	// the velocity / position iteration counts of the item's island (CalculateSolverSteps::Finalize)
	B2J_D void island_steps(const ConstraintSrc &src, uint32_t type1, uint32_t &vsteps, uint32_t &psteps) const
	{
		uint32_t r = s.root[type1 == B2J_MOTION_DYNAMIC? src.b1 : src.b2];
		uint32_t steps = s.island_steps[r];
		vsteps = steps & 0xff; psteps = (steps >> 8) & 0xff;
		if (steps & (1u << 16)) vsteps = vsteps > w.settings.num_velocity_steps? vsteps : w.settings.num_velocity_steps;
		if (steps & (1u << 17)) psteps = psteps > w.settings.num_position_steps? psteps : w.settings.num_position_steps;
		// (almost always already at the maximum: keep the single address atomics off the hot path)
		if (vsteps > volatile_load(&w.counters->max_velocity_steps)) atomic_max(&w.counters->max_velocity_steps, vsteps);
		if (psteps > volatile_load(&w.counters->max_position_steps)) atomic_max(&w.counters->max_position_steps, psteps);
	}
	B2J_D void operator()(uint32_t i) const // i = solve position
	{
		const Constraints &c = s.con;
		const ConstraintSrc &src = s.src[s.solve_src[i]];
		uint32_t m = src.manifold;
		if (m & 0x80000000u)
		{
			// a non contact constraint (b2j_joints.h): header only -- bodies, constraint index, the island's iteration counts
			uint32_t jt1 = w.info[src.b1].motion_type, jt2 = w.info[src.b2].motion_type;
			uint32_t jv, jp;
			island_steps(src, jt1, jv, jp);
			// (bits 0-2, a contact's point count, carry the constraint type: b2j_joints.h reads it from here instead of its definition)
			uint32_t jtype = (s.joint_steps[m & 0x7fffffffu] >> 16) & 7u;
			ConstraintHeader jh; jh.b1 = src.b1; jh.b2 = src.b2; jh.manifold = m & 0x7fffffffu; jh.meta = META_JOINT | jtype | (jt1 << 3) | (jt2 << 5) | (jv << 8) | (jp << 16);
			c.hdr[i] = jh;
			return;
		}
		const CachedManifold &cm = w.write_cache.manifolds[m];
		BodyKin k1 = load_body_kin(w, src.b1), k2 = load_body_kin(w, src.b2);
		uint32_t type1 = k1.type, type2 = k2.type;
		int n = cm.num_points;

		uint32_t vsteps, psteps;
		island_steps(src, type1, vsteps, psteps);

		uint32_t meta = (uint32_t)n | (type1 << 3) | (type2 << 5) | (vsteps << 8) | (psteps << 16) | ((k1.dofs & 7u) << META_DOFS1_SHIFT) | ((k2.dofs & 7u) << META_DOFS2_SHIFT);

		BodyParams p1 = w.params[src.b1], p2 = w.params[src.b2];
		float combined_friction = sqrt_(p1.friction * p2.friction);
		float combined_restitution = fmax_(p1.restitution, p2.restitution);

		V3 normal;
		V3 p1_ws[4], p2_ws[4];
		if (cm.flags & MANIFOLD_FROM_CACHE)
		{
			Xf t1 = xf_rotation_translation(k1.q, k1.x), t2 = xf_rotation_translation(k2.q, k2.x);
			normal = normalized(mul(t2.r, v3_load(cm.normal)));
			for (int p = 0; p < n; ++p)
			{
				p1_ws[p] = mul(t1, v3_load(cm.p1[p]));
				p2_ws[p] = mul(t2, v3_load(cm.p2[p]));
			}
		}
		else
		{
			const ManifoldWS &ws = s.man_ws[m];
			normal = v3_load(ws.normal);
			for (int p = 0; p < n; ++p)
			{
				p1_ws[p] = v3_load(ws.p1[p]);
				p2_ws[p] = v3_load(ws.p2[p]);
			}
		}
		cp_at(c, CP_NORMAL, i) = f4(normal, combined_friction);

		V3 ws_contacts[4];
		float lambda_pt[4] = { 0.0f, 0.0f, 0.0f, 0.0f }, dist[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
		for (int p = 0; p < n; ++p)
		{
			int base = CP_PT0 + p * 4;
			ws_contacts[p] = 0.5f * (p1_ws[p] + p2_ws[p]);
			cp_at(c, CP_LP0 + p * 2, i) = f4(v3_load(cm.p1[p]));
			cp_at(c, CP_LP0 + p * 2 + 1, i) = f4(v3_load(cm.p2[p]));

			// CalculateNonPenetrationConstraintProperties
			V3 pm = 0.5f * (p1_ws[p] + p2_ws[p]);
			V3 r1 = pm - k1.x, r2 = pm - k2.x;
			V3 relative_velocity;
			if (type1 != B2J_MOTION_STATIC && type2 != B2J_MOTION_STATIC)
				relative_velocity = (k2.v + cross(k2.av, r2)) - (k1.v + cross(k1.av, r1));
			else if (type1 != B2J_MOTION_STATIC)
				relative_velocity = -(k1.v + cross(k1.av, r1));
			else
				relative_velocity = k2.v + cross(k2.av, r2);
			float normal_velocity = dot(relative_velocity, normal);
			float penetration = dot(p1_ws[p] - p2_ws[p], normal);
			float speculative_contact_velocity_bias = fmax_(0.0f, -penetration / dt);
			float normal_velocity_bias;
			if (combined_restitution > 0.0f && normal_velocity < -w.settings.min_velocity_for_restitution)
			{
				if (normal_velocity < -speculative_contact_velocity_bias)
				{
					V3 relative_acceleration;
					if (type1 != B2J_MOTION_STATIC && type2 != B2J_MOTION_STATIC)
						relative_acceleration = w.gravity * (k2.gravity_factor - k1.gravity_factor);
					else if (type1 != B2J_MOTION_STATIC)
						relative_acceleration = -w.gravity * k1.gravity_factor;
					else
						relative_acceleration = w.gravity * k2.gravity_factor;
					if (type1 == B2J_MOTION_DYNAMIC) relative_acceleration -= k1.force * k1.inv_mass;
					if (type2 == B2J_MOTION_DYNAMIC) relative_acceleration += k2.force * k2.inv_mass;
					float force_delta_velocity = fmin_(0.0f, dot(relative_acceleration, normal) * dt);
					normal_velocity_bias = combined_restitution * (normal_velocity - force_delta_velocity);
				}
				else
					normal_velocity_bias = speculative_contact_velocity_bias;
			}
			else
				normal_velocity_bias = speculative_contact_velocity_bias;
			PartRegs part;
			part.lambda = cm.lambda[p];
			part_calculate(part, type1, type2, k1.inv_mass, k1.inv_i, r1, k2.inv_mass, k2.inv_i, r2, normal, normal_velocity_bias);
			part_store(c, base, i, type1, type2, part);
			lambda_pt[p] = part.lambda;
		}
		cp_at(c, CP_LAMBDA_PT, i) = f4(lambda_pt[0], lambda_pt[1], lambda_pt[2], lambda_pt[3]);

		// friction (CalculateFrictionConstraintProperties)
		float lambda_f1 = 0.0f, lambda_f2 = 0.0f, angular_lambda = 0.0f, angular_eff = 0.0f;
		if (combined_friction > 0.0f)
		{
			V3 t1 = normalized_perpendicular(normal);
			V3 t2 = cross(normal, t1);
			V3 friction_point = v3_zero();
			for (int p = 0; p < n; ++p) friction_point += ws_contacts[p];
			friction_point = friction_point / (float)n;
			for (int p = 0; p < n; ++p)
			{
				V3 delta = ws_contacts[p] - friction_point;
				dist[p] = length(delta - dot(delta, normal) * normal);
			}
			V3 r1 = friction_point - k1.x, r2 = friction_point - k2.x;
			PartRegs f1, f2;
			f1.lambda = cm.friction_lambda[0];
			f2.lambda = cm.friction_lambda[1];
			part_calculate(f1, type1, type2, k1.inv_mass, k1.inv_i, r1, k2.inv_mass, k2.inv_i, r2, t1, 0.0f);
			part_calculate(f2, type1, type2, k1.inv_mass, k1.inv_i, r1, k2.inv_mass, k2.inv_i, r2, t2, 0.0f);
			part_store(c, CP_FR0, i, type1, type2, f1);
			part_store(c, CP_FR0 + 4, i, type1, type2, f2);
			lambda_f1 = f1.lambda; lambda_f2 = f2.lambda;
			if (f1.eff != 0.0f || f2.eff != 0.0f) meta |= META_LINEAR_FRICTION;
			if (n > 1)
			{
				// AngularFrictionConstraintPart::CalculateConstraintProperties (bias = 0)
				angular_lambda = cm.angular_lambda;
				V3 i1a = v3_zero(), i2a = v3_zero();
				if (type1 == B2J_MOTION_DYNAMIC) { i1a = mul(k1.inv_i, normal); cp_at(c, CP_ANG_I1, i) = f4(i1a); }
				if (type2 == B2J_MOTION_DYNAMIC) { i2a = mul(k2.inv_i, normal); cp_at(c, CP_ANG_I2, i) = f4(i2a); }
				float inv_effective_mass = 0.0f;
				if (type1 == B2J_MOTION_DYNAMIC && type2 == B2J_MOTION_DYNAMIC) inv_effective_mass = dot(normal, i1a + i2a);
				else if (type1 == B2J_MOTION_DYNAMIC) inv_effective_mass = dot(normal, i1a);
				else if (type2 == B2J_MOTION_DYNAMIC) inv_effective_mass = dot(normal, i2a);
				if (inv_effective_mass == 0.0f) angular_lambda = 0.0f;
				else angular_eff = 1.0f / inv_effective_mass;
			}
			if (angular_eff != 0.0f) meta |= META_ANGULAR_FRICTION;
		}
		// (no friction: Deactivate() x3 = the zero lambdas / effective masses below; the friction part planes are never read)
		cp_at(c, CP_MASS, i) = f4(k1.inv_mass, k2.inv_mass, dist[0], dist[1]);
		cp_at(c, CP_DIST, i) = f4(dist[2], dist[3], angular_eff, 0.0f);
		cp_at(c, CP_LAMBDA_FR, i) = f4(lambda_f1, lambda_f2, angular_lambda, 0.0f);
		ConstraintHeader hdr; hdr.b1 = src.b1; hdr.b2 = src.b2; hdr.manifold = m; hdr.meta = meta;
		c.hdr[i] = hdr;
	}
};

// ---- velocity solve --------------------------------------------------------------------------------------------------

struct VelState { V3 v1, w1, v2, w2; };

// ContactConstraintPart::ApplyVelocityStep
B2J_D bool part_apply_velocity_step(const PartRegs &r, uint32_t type1, uint32_t type2, VelState &s, float inv_m1, float inv_m2, V3 axis, float lambda)
{
	if (lambda != 0.0f)
	{
		if (type1 == B2J_MOTION_DYNAMIC)
		{
			s.v1 -= (lambda * inv_m1) * axis;
			s.w1 -= lambda * r.i1;
		}
		if (type2 == B2J_MOTION_DYNAMIC)
		{
			s.v2 += (lambda * inv_m2) * axis;
			s.w2 += lambda * r.i2;
		}
		return true;
	}
	return false;
}

// SolveVelocityConstraintGetTotalLambda
B2J_D float part_get_total_lambda(const PartRegs &r, uint32_t type1, uint32_t type2, const VelState &s, V3 axis)
{
	float jv;
	if (type1 != B2J_MOTION_STATIC && type2 != B2J_MOTION_STATIC)
		jv = dot(axis, s.v1 - s.v2);
	else if (type1 != B2J_MOTION_STATIC)
		jv = dot(axis, s.v1);
	else
		jv = dot(axis, -s.v2);
	if (type1 != B2J_MOTION_STATIC)
		jv += dot(r.r1x, s.w1);
	if (type2 != B2J_MOTION_STATIC)
		jv -= dot(r.r2x, s.w2);
	float lambda = r.eff * (jv - r.bias);
	return r.lambda + lambda;
}

// SolveVelocityConstraintApplyLambda
B2J_D bool part_apply_lambda(PartRegs &r, uint32_t type1, uint32_t type2, VelState &s, float inv_m1, float inv_m2, V3 axis, float total_lambda)
{
	float delta_lambda = total_lambda - r.lambda;
	r.lambda = total_lambda;
	return part_apply_velocity_step(r, type1, type2, s, inv_m1, inv_m2, axis, delta_lambda);
}

B2J_D void load_vel_state(const DWorld &w, uint32_t b1, uint32_t b2, uint32_t type1, uint32_t type2, VelState &s)
{
	if (type1 != B2J_MOTION_STATIC) { s.v1 = to_v3(w.linear_velocity[b1]); s.w1 = to_v3(w.angular_velocity[b1]); }
	else { s.v1 = v3_zero(); s.w1 = v3_zero(); }
	if (type2 != B2J_MOTION_STATIC) { s.v2 = to_v3(w.linear_velocity[b2]); s.w2 = to_v3(w.angular_velocity[b2]); }
	else { s.v2 = v3_zero(); s.w2 = v3_zero(); }
}

// (Body::SetLinearVelocityClamped... MotionProperties::LockTranslation: the translation DOF masks travel in the constraint's meta word)
B2J_D void store_vel_state(const DWorld &w, uint32_t b1, uint32_t b2, uint32_t meta, const VelState &s)
{
	uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
	if (type1 == B2J_MOTION_DYNAMIC)
	{
		w.linear_velocity[b1] = f4(lock_translation(s.v1, (meta >> META_DOFS1_SHIFT) & 7u));
		w.angular_velocity[b1] = f4(s.w1);
	}
	if (type2 == B2J_MOTION_DYNAMIC)
	{
		w.linear_velocity[b2] = f4(lock_translation(s.v2, (meta >> META_DOFS2_SHIFT) & 7u));
		w.angular_velocity[b2] = f4(s.w2);
	}
}

// sWarmStartConstraint (ContactConstraintManager.cpp:1587-1622) on registers: the lambdas (lpt = 4 points, lfr = friction 1, friction 2,
// angular) are scaled by the warm start ratio and applied; returns true if a velocity changed
template <class Src> B2J_D bool warm_start_core(const Src &src, uint32_t meta, float ratio, VelState &s, F4 &lpt, F4 &lfr)
{
	int n = (int)(meta & 7);
	uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
	V3 normal = to_v3(src.ro(CP_NORMAL));
	V3 t1 = normalized_perpendicular(normal);
	V3 t2 = cross(normal, t1);
	F4 mass = src.ro(CP_MASS);
	float inv_m1 = mass.x, inv_m2 = mass.y;
	bool any = false;
	if (meta & META_LINEAR_FRICTION)
	{
		PartRegs f1 = part_load(src, CP_FR0, type1, type2), f2 = part_load(src, CP_FR0 + 4, type1, type2);
		if (f1.eff != 0.0f)
		{
			lfr.x *= ratio;
			if (part_apply_velocity_step(f1, type1, type2, s, inv_m1, inv_m2, t1, lfr.x)) any = true;
		}
		if (f2.eff != 0.0f)
		{
			lfr.y *= ratio;
			if (part_apply_velocity_step(f2, type1, type2, s, inv_m1, inv_m2, t2, lfr.y)) any = true;
		}
	}
	if (meta & META_ANGULAR_FRICTION)
	{
		lfr.z *= ratio;
		float l = lfr.z;
		if (l != 0.0f)
		{
			if (type1 == B2J_MOTION_DYNAMIC) s.w1 -= l * to_v3(src.ro(CP_ANG_I1));
			if (type2 == B2J_MOTION_DYNAMIC) s.w2 += l * to_v3(src.ro(CP_ANG_I2));
			any = true;
		}
	}
	float lp[4] = { lpt.x, lpt.y, lpt.z, lpt.w };
#if defined(__CUDA_ARCH__)
	#pragma unroll
#endif
	for (int p = 0; p < 4; ++p)
		if (p < n)
		{
			PartRegs r = part_load(src, CP_PT0 + p * 4, type1, type2);
			lp[p] *= ratio;
			if (part_apply_velocity_step(r, type1, type2, s, inv_m1, inv_m2, normal, lp[p])) any = true;
		}
	lpt = f4(lp[0], lp[1], lp[2], lp[3]);
	return any;
}

// sSolveVelocityConstraint (ContactConstraintManager.cpp:1683-1767) on registers. Everything the constraint needs is fetched up front
// (one memory round trip per constraint instead of one per part); returns true if a velocity changed.
// kLate: the four contact point parts are fetched where they are used instead of up front (fewer live registers -> more resident
// warps, one more L2 round trip per point; an A/B form, see KSolveVelocityLate)
template <class Src, bool kLate = false> B2J_D bool solve_velocity_core(const Src &src, uint32_t meta, VelState &s, F4 &lpt, F4 &lfr)
{
	int n = (int)(meta & 7);
	uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
	bool linear_friction_active = (meta & META_LINEAR_FRICTION) != 0;
	bool angular_friction_active = (meta & META_ANGULAR_FRICTION) != 0;
	// ---- loads
	F4 nf = src.ro(CP_NORMAL), mass = src.ro(CP_MASS), dd = src.ro(CP_DIST);
	V3 normal = to_v3(nf);
	float mu = nf.w, inv_m1 = mass.x, inv_m2 = mass.y;
	float dist[4] = { mass.z, mass.w, dd.x, dd.y };
	float lp[4] = { lpt.x, lpt.y, lpt.z, lpt.w };
	PartRegs pt[4];
	if (!kLate)
	{
#if defined(__CUDA_ARCH__)
		#pragma unroll
#endif
		for (int p = 0; p < 4; ++p)
			if (p < n)
			{
				pt[p] = part_load(src, CP_PT0 + p * 4, type1, type2);
				pt[p].lambda = lp[p];
			}
	}
	PartRegs f1, f2;
	if (linear_friction_active)
	{
		f1 = part_load(src, CP_FR0, type1, type2);
		f2 = part_load(src, CP_FR0 + 4, type1, type2);
		f1.lambda = lfr.x; f2.lambda = lfr.y;
	}
	float ang_eff = dd.z, ang_bias = dd.w, ang_lambda = lfr.z;
	V3 ang_i1 = v3_zero(), ang_i2 = v3_zero();
	if (angular_friction_active)
	{
		if (type1 == B2J_MOTION_DYNAMIC) ang_i1 = to_v3(src.ro(CP_ANG_I1));
		if (type2 == B2J_MOTION_DYNAMIC) ang_i2 = to_v3(src.ro(CP_ANG_I2));
	}

	// ---- solve
	V3 t1 = normalized_perpendicular(normal);
	V3 t2 = cross(normal, t1);
	bool any = false;
	float max_linear_lambda = 0.0f, max_angular_lambda = 0.0f;
	if (linear_friction_active || angular_friction_active)
	{
#if defined(__CUDA_ARCH__)
		#pragma unroll
#endif
		for (int p = 0; p < 4; ++p)
			if (p < n)
			{
				float lambda = lp[p];
				max_linear_lambda += lambda;
				max_angular_lambda += dist[p] * lambda;
			}
		max_linear_lambda *= mu;
		max_angular_lambda *= mu;
	}
	if (linear_friction_active)
	{
		float lambda1 = part_get_total_lambda(f1, type1, type2, s, t1);
		float lambda2 = part_get_total_lambda(f2, type1, type2, s, t2);
		float total_lambda_sq = square(lambda1) + square(lambda2);
		if (total_lambda_sq > square(max_linear_lambda))
		{
			float scale = max_linear_lambda / sqrt_(total_lambda_sq);
			lambda1 *= scale;
			lambda2 *= scale;
		}
		if (part_apply_lambda(f1, type1, type2, s, inv_m1, inv_m2, t1, lambda1)) any = true;
		if (part_apply_lambda(f2, type1, type2, s, inv_m1, inv_m2, t2, lambda2)) any = true;
		lfr.x = f1.lambda; lfr.y = f2.lambda;
	}
	if (angular_friction_active)
	{
		// AngularFrictionConstraintPart::SolveVelocityConstraint
		float jv;
		if (type1 != B2J_MOTION_STATIC && type2 != B2J_MOTION_STATIC) jv = dot(normal, s.w1 - s.w2);
		else if (type1 != B2J_MOTION_STATIC) jv = dot(normal, s.w1);
		else jv = -dot(normal, s.w2);
		float total = ang_lambda;
		float lambda = ang_eff * (jv - ang_bias);
		float new_lambda = clamp_(total + lambda, -max_angular_lambda, max_angular_lambda);
		lambda = new_lambda - total;
		lfr.z = new_lambda;
		if (lambda != 0.0f)
		{
			if (type1 == B2J_MOTION_DYNAMIC) s.w1 -= lambda * ang_i1;
			if (type2 == B2J_MOTION_DYNAMIC) s.w2 += lambda * ang_i2;
			any = true;
		}
	}
#if defined(__CUDA_ARCH__)
	#pragma unroll
#endif
	for (int p = 0; p < 4; ++p)
		if (p < n)
		{
			if (kLate)
			{
				pt[p] = part_load(src, CP_PT0 + p * 4, type1, type2);
				pt[p].lambda = lp[p];
			}
			float total_lambda = part_get_total_lambda(pt[p], type1, type2, s, normal);
			total_lambda = fmax_(total_lambda, 0.0f);
			if (part_apply_lambda(pt[p], type1, type2, s, inv_m1, inv_m2, normal, total_lambda)) any = true;
			lp[p] = pt[p].lambda;
		}
	lpt = f4(lp[0], lp[1], lp[2], lp[3]);
	return any;
}

// sStoreAppliedImpulses (ContactConstraintManager.cpp:1819-1835) for one constraint
B2J_D void store_applied_impulses(const DWorld &w, uint32_t manifold, int n, F4 lpt, F4 lfr)
{
	CachedManifold &cm = w.write_cache.manifolds[manifold];
	float lp[4] = { lpt.x, lpt.y, lpt.z, lpt.w };
	for (int p = 0; p < n; ++p)
		cm.lambda[p] = lp[p];
	cm.friction_lambda[0] = lfr.x;
	cm.friction_lambda[1] = lfr.y;
	cm.angular_lambda = lfr.z;
}

// sWarmStartConstraint for solve positions [begin, begin + n)
// What a velocity solve thread may do BEFORE the previous phase's kernel has finished (programmatic dependent launch): the header is
// written by the setup kernel; prefetches only pull lines into L2 (the coherence point), they read no values.
B2J_D void solve_prologue_prefetch(const DWorld &w, const Constraints &c, const ConstraintHeader &hdr, uint32_t i, bool warm_start)
{
	uint32_t meta = hdr.meta;
	int n = (int)(meta & 7);
	for (int pl = CP_NORMAL; pl < CP_FR0; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
	if (meta & META_LINEAR_FRICTION) for (int pl = CP_FR0; pl < CP_PT0; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
	for (int pl = CP_PT0; pl < CP_PT0 + 4 * n; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
	(void)warm_start;
	if (((meta >> 3) & 3) != B2J_MOTION_STATIC) { prefetch_l2(&w.linear_velocity[hdr.b1]); prefetch_l2(&w.angular_velocity[hdr.b1]); }
	if (((meta >> 5) & 3) != B2J_MOTION_STATIC) { prefetch_l2(&w.linear_velocity[hdr.b2]); prefetch_l2(&w.angular_velocity[hdr.b2]); }
}

struct KWarmStart
{
	DWorld w; Constraints c; uint32_t begin; float ratio; uint32_t pdl = 0;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t i = begin + k;
		ConstraintHeader hdr = c.hdr[i];
		uint32_t meta = hdr.meta;
		uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
		if (meta & META_JOINT)
			return; // (worlds with non contact constraints are solved without programmatic dependent launch: KJointWarmStart owns this item)
		if (pdl)
		{
			solve_prologue_prefetch(w, c, hdr, i, true);
			grid_dependency_sync();
		}
		VelState s;
		load_vel_state(w, hdr.b1, hdr.b2, type1, type2, s);
		F4 lpt = cp_at(c, CP_LAMBDA_PT, i), lfr = cp_at(c, CP_LAMBDA_FR, i);
		GlobalPlanes src; src.c = c; src.i = i;
		bool any = warm_start_core(src, meta, ratio, s, lpt, lfr);
		cp_at(c, CP_LAMBDA_PT, i) = lpt;
		cp_at(c, CP_LAMBDA_FR, i) = lfr;
		if (any)
			store_vel_state(w, hdr.b1, hdr.b2, meta, s);
	}
};

// sSolveVelocityConstraint; iteration = 0 based velocity step index (constraints of islands with fewer steps skip).
template <bool kLate> struct KSolveVelocityT
{
	DWorld w; Constraints c; uint32_t begin; uint32_t iteration; uint32_t prefetch; uint32_t pdl = 0;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t i = begin + k;
		ConstraintHeader hdr = c.hdr[i];
		uint32_t meta = hdr.meta;
		const bool skip = iteration >= ((meta >> 8) & 0xff) || (meta & META_JOINT) != 0;
		if (pdl)
		{
			// (every thread synchronises before it leaves, also the ones whose island is done: the kernel after this one relies on it)
			if (!skip) solve_prologue_prefetch(w, c, hdr, i, false);
			grid_dependency_sync();
		}
		if (skip)
			return;
		int n = (int)(meta & 7);
		uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
		if (prefetch && !pdl)
		{
			// L2 prefetch of every plane of the constraint right after the header: -5 % on the per phase launches (measured); register
			// capped builds (more warps per SM) stay slower even with it (6.0 / 7.2 / 8.2 ms vs 5.0 ms at 128 / 96 / 80 registers)
			for (int pl = CP_NORMAL; pl < CP_FR0; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
			if (meta & META_LINEAR_FRICTION) for (int pl = CP_FR0; pl < CP_PT0; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
			for (int pl = CP_PT0; pl < CP_PT0 + 4 * n; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
		}
		VelState s;
		load_vel_state(w, hdr.b1, hdr.b2, type1, type2, s);
		F4 lpt = cp_at(c, CP_LAMBDA_PT, i), lfr = cp_at(c, CP_LAMBDA_FR, i);
		GlobalPlanes src; src.c = c; src.i = i;
		bool any = solve_velocity_core<GlobalPlanes, kLate>(src, meta, s, lpt, lfr);
		cp_at(c, CP_LAMBDA_PT, i) = lpt;
		if (meta & (META_LINEAR_FRICTION | META_ANGULAR_FRICTION))
			cp_at(c, CP_LAMBDA_FR, i) = lfr;
		if (any)
			store_vel_state(w, hdr.b1, hdr.b2, meta, s);
		// sStoreAppliedImpulses, fused into the last velocity iteration of the constraint's island (saves a pass over all constraints)
		if (iteration + 1 == ((meta >> 8) & 0xff))
			store_applied_impulses(w, hdr.manifold, n, lpt, lfr);
	}
};
typedef KSolveVelocityT<false> KSolveVelocity;
// the form the per phase launches use by default: contact point parts fetched late, launched with a register budget (jolt_b200.cu: solve_late_mode)
typedef KSolveVelocityT<true> KSolveVelocityLate;

// sStoreAppliedImpulses for constraints of islands that run NO velocity iteration (the others store in their last iteration)
struct KStoreImpulses
{
	DWorld w; Constraints c;
	B2J_D void operator()(uint32_t i) const
	{
		ConstraintHeader hdr = c.hdr[i];
		if (((hdr.meta >> 8) & 0xff) != 0 || (hdr.meta & META_JOINT) != 0)
			return;
		int n = (int)(hdr.meta & 7);
		CachedManifold &cm = w.write_cache.manifolds[hdr.manifold];
		F4 lpt = cp_at(c, CP_LAMBDA_PT, i), lfr = cp_at(c, CP_LAMBDA_FR, i);
		float lp[4] = { lpt.x, lpt.y, lpt.z, lpt.w };
		for (int p = 0; p < n; ++p)
			cm.lambda[p] = lp[p];
		cm.friction_lambda[0] = lfr.x;
		cm.friction_lambda[1] = lfr.y;
		cm.angular_lambda = lfr.z;
	}
};

// ---- integrate (JobIntegrateVelocity, discrete motion quality) -------------------------------------------------------
struct KIntegrate
{
	DWorld w; float dt;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		BodyInfo info = w.info[b];
		V3 v = to_v3(w.linear_velocity[b]), av = to_v3(w.angular_velocity[b]);
		if (info.motion_type == B2J_MOTION_DYNAMIC)
		{
			BodyParams p = w.params[b];
			clamp_velocity(v, p.max_linear_velocity);
			clamp_velocity(av, p.max_angular_velocity);
			w.linear_velocity[b] = f4(v);
			w.angular_velocity[b] = f4(av);
		}
		w.rotation[b] = f4(add_rotation_step(to_q4(w.rotation[b]), av * dt, false));
		V3 x = to_v3(w.position[b]);
		x += lock_translation(v * dt, info.allowed_dofs);
		w.position[b] = f4(x);
	}
};

// ---- position solve (sSolvePositionConstraint) ------------------------------------------------------------------------
struct KSolvePosition
{
	DWorld w; Constraints c; uint32_t begin; uint32_t iteration; uint32_t pdl = 0;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t i = begin + k;
		ConstraintHeader hdr = c.hdr[i];
		uint32_t meta = hdr.meta;
		const bool skip = iteration >= ((meta >> 16) & 0xff) || (meta & META_JOINT) != 0;
		if (pdl)
		{
			if (!skip)
			{
				int np = (int)(meta & 7);
				prefetch_l2(&c.cp[(size_t)CP_NORMAL * c.capacity + i]);
				prefetch_l2(&c.cp[(size_t)CP_MASS * c.capacity + i]);
				for (int pl = CP_LP0; pl < CP_LP0 + 2 * np; ++pl) prefetch_l2(&c.cp[(size_t)pl * c.capacity + i]);
				prefetch_l2(&w.position[hdr.b1]); prefetch_l2(&w.rotation[hdr.b1]);
				prefetch_l2(&w.position[hdr.b2]); prefetch_l2(&w.rotation[hdr.b2]);
			}
			grid_dependency_sync();
		}
		if (skip)
			return;
		int n = (int)(meta & 7);
		uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
		uint32_t b1 = hdr.b1, b2 = hdr.b2;
		V3 x1 = to_v3(w.position[b1]), x2 = to_v3(w.position[b2]);
		Q4 q1 = to_q4(w.rotation[b1]), q2 = to_q4(w.rotation[b2]);
		uint32_t dofs1 = w.info[b1].allowed_dofs, dofs2 = w.info[b2].allowed_dofs;
		// transforms are fetched once per constraint, inertia / positions are re-read per point (bodies move between points)
		Xf transform1 = xf_rotation_translation(q1, x1), transform2 = xf_rotation_translation(q2, x2);
		V3 normal = to_v3(cp_ro(c, CP_NORMAL, i));
		F4 mass = cp_ro(c, CP_MASS, i);
		float inv_m1 = mass.x, inv_m2 = mass.y;
		V3 diag1 = v3_zero(), diag2 = v3_zero();
		Q4 irot1 = q4_identity(), irot2 = q4_identity();
		if (type1 == B2J_MOTION_DYNAMIC) { diag1 = to_v3(w.inv_inertia_diag[b1]); irot1 = to_q4(w.inertia_rotation[b1]); }
		if (type2 == B2J_MOTION_DYNAMIC) { diag2 = to_v3(w.inv_inertia_diag[b2]); irot2 = to_q4(w.inertia_rotation[b2]); }
		bool any = false;
		for (int p = 0; p < n; ++p)
		{
			V3 p1 = mul(transform1, to_v3(cp_ro(c, CP_LP0 + p * 2, i)));
			V3 p2 = mul(transform2, to_v3(cp_ro(c, CP_LP0 + p * 2 + 1, i)));
			float separation = fmax_(dot(p2 - p1, normal) + w.settings.penetration_slop, -w.settings.max_penetration_distance);
			if (separation < 0.0f)
			{
				M33 inv_i1 = type1 == B2J_MOTION_DYNAMIC? inverse_inertia_for_rotation(m33_rotation(q1), irot1, diag1, dofs1) : m33_zero();
				M33 inv_i2 = type2 == B2J_MOTION_DYNAMIC? inverse_inertia_for_rotation(m33_rotation(q2), irot2, diag2, dofs2) : m33_zero();
				V3 pm = 0.5f * (p1 + p2);
				V3 r1 = pm - x1, r2 = pm - x2;
				// the part is recomputed in registers only: nothing reads the stored part after the velocity solve (the impulses
				// were already saved by KStoreImpulses)
				PartRegs part;
				part.lambda = 0.0f;
				part_calculate(part, type1, type2, inv_m1, inv_i1, r1, inv_m2, inv_i2, r2, normal, 0.0f);
				// ContactConstraintPart::SolvePositionConstraint
				if (separation != 0.0f)
				{
					float lambda = -part.eff * w.settings.baumgarte * separation;
					if (type1 == B2J_MOTION_DYNAMIC)
					{
						x1 -= lock_translation((lambda * inv_m1) * normal, dofs1);
						q1 = add_rotation_step(q1, lambda * part.i1, true);
					}
					if (type2 == B2J_MOTION_DYNAMIC)
					{
						x2 += lock_translation((lambda * inv_m2) * normal, dofs2);
						q2 = add_rotation_step(q2, lambda * part.i2, false);
					}
					any = true;
				}
			}
		}
		if (any)
		{
			if (type1 == B2J_MOTION_DYNAMIC) { w.position[b1] = f4(x1); w.rotation[b1] = f4(q1); }
			if (type2 == B2J_MOTION_DYNAMIC) { w.position[b2] = f4(x2); w.rotation[b2] = f4(q2); }
		}
	}
};

#if !defined(B2J_HOSTSIM) && defined(__CUDACC__)
// Small single worlds (a few thousand constraints): one launch per phase per iteration is ~340 dependent launches of a few
// microseconds each for a Pyramid. Here ONE small cooperative grid (8 blocks) runs warm start + all velocity iterations (or all
// position iterations), the phases separated by grid barriers; phase offsets and iteration counts are read on the device. Same per
// constraint functors as the per phase launches. (One block alone is too slow: 8 warps cannot hide the L2 latency; for batches of
// worlds per phase launches over all worlds win)
struct KSolveSmallVelocity { }; // (profiling categories)
struct KSolveSmallPosition { };
template <bool kPosition> __global__ void __launch_bounds__(256) solve_small_kernel(const DWorld w, const SolveCtx s, float warm_start_ratio)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = kPosition? w.counters->max_position_steps : w.counters->max_velocity_steps;
	const uint32_t *off = s.phase_count;
	if (!kPosition)
	{
		KWarmStart ws; ws.w = w; ws.c = s.con; ws.begin = 0; ws.ratio = warm_start_ratio;
		for (uint32_t p = 0; p < np; ++p)
		{
			for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) ws(k);
			grid.sync();
		}
		KSolveVelocity sv; sv.w = w; sv.c = s.con; sv.begin = 0; sv.prefetch = 0;
		for (uint32_t it = 0; it < steps; ++it)
		{
			sv.iteration = it;
			for (uint32_t p = 0; p < np; ++p)
			{
				for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) sv(k);
				grid.sync();
			}
		}
	}
	else
	{
		KSolvePosition sp; sp.w = w; sp.c = s.con; sp.begin = 0;
		for (uint32_t it = 0; it < steps; ++it)
		{
			sp.iteration = it;
			for (uint32_t p = 0; p < np; ++p)
			{
				for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) sp(k);
				grid.sync();
			}
		}
	}
}
#endif

#if !defined(B2J_HOSTSIM) && defined(__CUDACC__)
// ---- the whole velocity solve in ONE persistent launch, constraint planes streamed through shared memory by TMA -------------------
//
// One cooperative launch runs the warm start and every velocity iteration of every phase (a grid wide barrier between phases; phase
// offsets, phase and iteration counts are read on the device: no per phase launch, no host round trip). Inside a phase every WARP
// owns the tiles gw, gw + NW, ... of 32 consecutive constraints:
//   * the read-only planes of a tile are copied global -> shared memory by 1-D TMA bulk copies (cp.async.bulk, UBLKCP in SASS: one
//     512 byte copy per plane the tile needs) that complete on the warp's mbarrier; the copy of the next tile is issued as soon as the
//     warp is done with the stage, and while it is in flight the other warps of the SM compute (12 warps x 14.5 KB = 174 KB of shared
//     memory per SM in flight, independent of the register budget of the solve code),
//   * headers are register prefetched a tile ahead, the body velocities (L2 loads: other SMs wrote them in the previous phase) and
//     the two lambda planes (read + written, generic proxy) are loaded with the copy,
//   * the tile is solved out of shared memory (conflict free LDS.128) with the arithmetic of the per phase kernels.
// The grid barrier is split: a block ARRIVES (one atomic), then prepares its first tile of the next phase (header loads, plane copies:
// read-only data, safe before the barrier completes), then WAITS; only the velocities / lambdas are fetched after it.
// Measured (4096 Pyramid worlds in one group, steps 5..25): 28.7 ms per step = 0.55 of the HBM peak against 0.26 for the per phase
// launches -- the kernel is bound by the dependent instruction chain of a tile (~900 FP32 instructions without FMA), not by memory;
// twice the warps at 128 registers spill and lose.
enum { SV_STAGE_F4 = SV_NUM_SLOTS * 32 };
constexpr size_t sv_smem_bytes(int warps) { return (size_t)warps * SV_STAGE_F4 * sizeof(F4) + (size_t)warps * sizeof(uint64_t); }

B2J_D uint32_t sv_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
B2J_D void sv_mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sv_smem_addr(bar)), "r"(count) : "memory"); }
B2J_D void sv_mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sv_smem_addr(bar)), "r"(bytes) : "memory"); }
B2J_D void sv_mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(sv_smem_addr(bar)), "r"(parity) : "memory");
}
B2J_D void sv_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(sv_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(sv_smem_addr(bar)) : "memory");
}
// L2 load of a body quantity another SM may have written before the last grid barrier (the barrier does not invalidate L1)
B2J_D F4 sv_load_l2(const F4 *p) { float4 v = __ldcg(reinterpret_cast<const float4 *>(p)); return f4(v.x, v.y, v.z, v.w); }

// slots of the shared memory stage a constraint reads (bit = slot); kWarm: the warm start pass does not read the r2 x axis planes
template <bool kWarm> B2J_D uint32_t sv_slot_mask(uint32_t meta)
{
	uint32_t n = meta & 7, type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
	uint32_t part = (1u << CP_PART_R1X_EFF) | (1u << CP_PART_I1_BIAS);
	if (!kWarm && type2 != B2J_MOTION_STATIC) part |= 1u << CP_PART_R2X;
	if (type2 == B2J_MOTION_DYNAMIC) part |= 1u << CP_PART_I2;
	uint32_t mask = (1u << SV_SLOT_NORMAL) | (1u << SV_SLOT_MASS);
	if (!kWarm) mask |= 1u << SV_SLOT_DIST;
	if (meta & META_LINEAR_FRICTION) mask |= (part | (part << 4)) << SV_SLOT_FR0;
	if (meta & META_ANGULAR_FRICTION)
	{
		if (type1 == B2J_MOTION_DYNAMIC) mask |= 1u << SV_SLOT_ANG_I1;
		if (type2 == B2J_MOTION_DYNAMIC) mask |= 1u << SV_SLOT_ANG_I2;
	}
	for (uint32_t p = 0; p < n; ++p) mask |= part << (SV_SLOT_PT0 + 4 * p);
	return mask;
}

// what a lane fetches into registers with the copy of its tile: the velocities of its two bodies and its two lambda planes
struct SvPre { F4 v1, w1, v2, w2, lpt, lfr; };

struct KSolveVelocityAll { }; // (profiling category)
template <int SV_WARPS> __global__ void __launch_bounds__(SV_WARPS * 32, 1) solve_velocity_tma_kernel(const DWorld w, const SolveCtx s, float warm_start_ratio)
{
	extern __shared__ __align__(128) unsigned char sv_smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	F4 *stage = reinterpret_cast<F4 *>(sv_smem) + (size_t)warp * SV_STAGE_F4;
	uint64_t *bar = reinterpret_cast<uint64_t *>(sv_smem + (size_t)SV_WARPS * SV_STAGE_F4 * sizeof(F4)) + warp;
	if (lane == 0)
	{
		sv_mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	uint32_t parity = 0;

	const Constraints c = s.con;
	const uint32_t gw = blockIdx.x * SV_WARPS + warp, nw = gridDim.x * SV_WARPS;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = w.counters->max_velocity_steps;
	const uint32_t *off = s.phase_count;
	volatile uint32_t *grid_bar = s.grid_barrier;   // zeroed before the launch, counts block arrivals
	uint32_t barrier_target = 0;

	// the (pass, phase) steps of the solve in order, empty phases skipped: pass 0 = warm start, pass it + 1 = velocity iteration it
	uint32_t pass = 0, p = 0;
	auto skip_empty = [&]() { while (pass <= steps) { while (p < np && off[p] == off[p + 1]) ++p; if (p < np) return; p = 0; ++pass; } };
	skip_empty();

	uint32_t begin = 0, end = 0, ntiles = 0;
	bool warm = true; uint32_t iteration = 0;
	auto enter_phase = [&]() { begin = off[p]; end = off[p + 1]; ntiles = (end - begin + 31) >> 5; warm = pass == 0; iteration = pass - 1; };
	// header of the lane's constraint in tile t of the current phase (valid = false past the end of the phase / nothing to do this pass)
	auto load_hdr = [&](uint32_t t, bool &valid) -> ConstraintHeader {
		uint32_t i = begin + (t << 5) + lane;
		valid = t < ntiles && i < end;
		ConstraintHeader h; h.b1 = 0; h.b2 = 0; h.manifold = 0; h.meta = 0;
		if (valid) { uint4 v = __ldg(reinterpret_cast<const uint4 *>(&c.hdr[i])); h.b1 = v.x; h.b2 = v.y; h.manifold = v.z; h.meta = v.w; }
		// constraints of islands with fewer velocity steps are done: they neither load nor solve in this pass
		if (!warm && iteration >= ((h.meta >> 8) & 0xff)) valid = false;
		return h;
	};
	// TMA bulk copies of the planes tile t needs
	auto issue_planes = [&](uint32_t t, const ConstraintHeader &h, bool valid, uint32_t &tile_mask) {
		uint32_t first = begin + (t << 5);
		uint32_t count = end - first < 32u? end - first : 32u;
		uint32_t lane_mask = valid? (warm? sv_slot_mask<true>(h.meta) : sv_slot_mask<false>(h.meta)) : 0u;
		tile_mask = __reduce_or_sync(0xffffffffu, lane_mask);
		if (tile_mask != 0)
		{
			if (lane == 0) sv_mbar_expect_tx(bar, (uint32_t)__popc(tile_mask) * count * (uint32_t)sizeof(F4));
			__syncwarp();
			if (lane < SV_NUM_SLOTS && ((tile_mask >> lane) & 1u))
				sv_bulk_g2s(stage + lane * 32, &c.cp[(size_t)sv_plane_of_slot((int)lane) * c.capacity + first], count * (uint32_t)sizeof(F4), bar);
		}
	};
	// the lane's velocities and lambdas (only valid once the previous phase is complete on the whole grid)
	auto load_pre = [&](uint32_t t, const ConstraintHeader &h, bool valid, SvPre &pre) {
		if (valid)
		{
			uint32_t type1 = (h.meta >> 3) & 3, type2 = (h.meta >> 5) & 3;
			uint32_t i = begin + (t << 5) + lane;
			if (type1 != B2J_MOTION_STATIC) { pre.v1 = sv_load_l2(&w.linear_velocity[h.b1]); pre.w1 = sv_load_l2(&w.angular_velocity[h.b1]); }
			if (type2 != B2J_MOTION_STATIC) { pre.v2 = sv_load_l2(&w.linear_velocity[h.b2]); pre.w2 = sv_load_l2(&w.angular_velocity[h.b2]); }
			pre.lpt = cp_at(c, CP_LAMBDA_PT, i);
			pre.lfr = cp_at(c, CP_LAMBDA_FR, i);
		}
	};

	ConstraintHeader h0, h1;
	bool valid0 = false, valid1 = false;
	uint32_t mask0 = 0, mask1 = 0;
	SvPre pre0, pre1;
	if (pass <= steps)
	{
		enter_phase();
		h0 = load_hdr(gw, valid0);
		if (gw < ntiles) issue_planes(gw, h0, valid0, mask0);
	}
	while (pass <= steps)
	{
		// ---- the tiles of this warp in the current phase (h0 / mask0 were prepared before the barrier)
		uint32_t t = gw;
		if (t < ntiles) load_pre(t, h0, valid0, pre0);
		while (t < ntiles)
		{
			uint32_t tn = t + nw;
			h1 = load_hdr(tn, valid1);
			if (mask0 != 0)
			{
				sv_mbar_wait(bar, parity);
				parity ^= 1;
			}
			if (valid0)
			{
				uint32_t meta = h0.meta;
				uint32_t type1 = (meta >> 3) & 3, type2 = (meta >> 5) & 3;
				uint32_t i = begin + (t << 5) + lane;
				VelState vs;
				if (type1 != B2J_MOTION_STATIC) { vs.v1 = to_v3(pre0.v1); vs.w1 = to_v3(pre0.w1); } else { vs.v1 = v3_zero(); vs.w1 = v3_zero(); }
				if (type2 != B2J_MOTION_STATIC) { vs.v2 = to_v3(pre0.v2); vs.w2 = to_v3(pre0.w2); } else { vs.v2 = v3_zero(); vs.w2 = v3_zero(); }
				F4 lpt = pre0.lpt, lfr = pre0.lfr;
				SmemPlanes src; src.stage = stage; src.lane = lane;
				bool any;
				if (warm)
				{
					any = warm_start_core(src, meta, warm_start_ratio, vs, lpt, lfr);
					cp_at(c, CP_LAMBDA_PT, i) = lpt;
					cp_at(c, CP_LAMBDA_FR, i) = lfr;
				}
				else
				{
					any = solve_velocity_core(src, meta, vs, lpt, lfr);
					cp_at(c, CP_LAMBDA_PT, i) = lpt;
					if (meta & (META_LINEAR_FRICTION | META_ANGULAR_FRICTION))
						cp_at(c, CP_LAMBDA_FR, i) = lfr;
				}
				if (any)
					store_vel_state(w, h0.b1, h0.b2, meta, vs);
				if (!warm && iteration + 1 == ((meta >> 8) & 0xff))
					store_applied_impulses(w, h0.manifold, (int)(meta & 7), lpt, lfr);
			}
			__syncwarp(); // every lane is done reading the stage: the copy of the next tile may overwrite it
			mask1 = 0;
			if (tn < ntiles) { issue_planes(tn, h1, valid1, mask1); load_pre(tn, h1, valid1, pre1); }
			h0 = h1; valid0 = valid1; pre0 = pre1; mask0 = mask1;
			t = tn;
		}
		// ---- split grid barrier: arrive, prepare the first tile of the next phase while the other blocks finish, wait
		++p;
		skip_empty();
		__syncthreads();
		barrier_target += gridDim.x;
		if (threadIdx.x == 0) { __threadfence(); atomicAdd((uint32_t *)grid_bar, 1u); }
		mask0 = 0; valid0 = false;
		if (pass <= steps)
		{
			enter_phase();
			h0 = load_hdr(gw, valid0);
			if (gw < ntiles) issue_planes(gw, h0, valid0, mask0);
			if (threadIdx.x == 0) { while (*grid_bar < barrier_target) { } __threadfence(); }
			__syncthreads();
		}
	}
}

// The position iterations of every phase in one cooperative launch (same per constraint functor as the per phase launches)
struct KSolvePositionAll { };
__global__ void __launch_bounds__(128) solve_position_all_kernel(const DWorld w, const SolveCtx s)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = w.counters->max_position_steps;
	const uint32_t *off = s.phase_count;
	KSolvePosition sp; sp.w = w; sp.c = s.con; sp.begin = 0;
	for (uint32_t it = 0; it < steps; ++it)
	{
		sp.iteration = it;
		for (uint32_t p = 0; p < np; ++p)
		{
			if (off[p] == off[p + 1])
				continue;
			for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) sp(k);
			grid.sync();
		}
	}
}

// Same for the velocity solve without the shared memory pipeline (kept for A/B measurements: B2J_SOLVE_MODE=1)
struct KSolveVelocityAllPlain { };
__global__ void __launch_bounds__(128) solve_velocity_all_kernel(const DWorld w, const SolveCtx s, float warm_start_ratio)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = w.counters->max_velocity_steps;
	const uint32_t *off = s.phase_count;
	KWarmStart ws; ws.w = w; ws.c = s.con; ws.begin = 0; ws.ratio = warm_start_ratio;
	for (uint32_t p = 0; p < np; ++p)
	{
		if (off[p] == off[p + 1])
			continue;
		for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) ws(k);
		grid.sync();
	}
	KSolveVelocity sv; sv.w = w; sv.c = s.con; sv.begin = 0; sv.prefetch = 1;
	for (uint32_t it = 0; it < steps; ++it)
	{
		sv.iteration = it;
		for (uint32_t p = 0; p < np; ++p)
		{
			if (off[p] == off[p + 1])
				continue;
			for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) sv(k);
			grid.sync();
		}
	}
}
#endif

// ---- bounds + sleeping (CheckSleepAndUpdateBounds, last collision step) ------------------------------------------------
struct KBoundsAndSleep
{
	DWorld w; SolveCtx s; float dt; uint32_t is_last;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		BodyInfo info = w.info[b];
		const ShapeDesc &shape = w.shapes[info.shape];
		V3 x = to_v3(w.position[b]);
		Q4 q = to_q4(w.rotation[b]);
		V3 mn, mx;
		world_bounds(w, shape, x, q, mn, mx);
		w.bounds_min[b] = f4(mn);
		w.bounds_max[b] = f4(mx);
		if (!is_last)
			return;
		// Body::UpdateSleepStateInternal
		bool can_sleep = false;
		if ((info.flags & B2J_BODY_ALLOW_SLEEPING) && !(info.flags & B2J_BODY_SENSOR))
		{
			float max_movement = w.settings.point_velocity_sleep_threshold * w.settings.time_before_sleep;
			V3 points[3];
			sleep_test_points(shape, x, q, points);
			bool reset = false;
			for (int i = 0; i < 3 && !reset; ++i)
			{
				F4 sp = w.sleep_spheres[b * 3 + i];
				V3 center = to_v3(sp);
				float radius = sp.w;
				// Sphere::EncapsulatePoint
				V3 d_vec = points[i] - center;
				float d_sq = length_sq(d_vec);
				if (d_sq > square(radius))
				{
					float d = sqrt_(d_sq);
					float new_radius = 0.5f * (radius + d);
					center += (new_radius - radius) / d * d_vec;
					radius = new_radius;
					w.sleep_spheres[b * 3 + i] = f4(center, radius);
				}
				if (radius > max_movement)
					reset = true;
			}
			if (reset)
			{
				for (int i = 0; i < 3; ++i) w.sleep_spheres[b * 3 + i] = f4(points[i], 0.0f);
				w.sleep_timer[b] = 0.0f;
			}
			else
			{
				float t = w.sleep_timer[b] + dt;
				w.sleep_timer[b] = t;
				can_sleep = t >= w.settings.time_before_sleep;
			}
		}
		if (!(can_sleep && w.settings.allow_sleeping))
			s.island_can_sleep[s.root[b]] = 0;
		// reset force and torque
		w.force[b] = f4(0, 0, 0, 0);
		w.torque[b] = f4(0, 0, 0, 0);
	}
};

// marks sleeping bodies, zeroes their velocity, computes the keep flag for the compaction of the active list
struct KDeactivate
{
	DWorld w; SolveCtx s; uint32_t *keep; b2j_activation_event *events; uint32_t max_events;
	B2J_D void operator()(uint32_t ai) const
	{
		uint32_t b = w.active[ai];
		bool sleep = s.island_can_sleep[s.root[b]] != 0;
		keep[ai] = sleep? 0u : 1u;
		if (sleep)
		{
			w.linear_velocity[b] = f4(0, 0, 0, 0);
			w.angular_velocity[b] = f4(0, 0, 0, 0);
			w.active_index[b] = B2J_INACTIVE_INDEX;
			atomic_add(&w.counters->num_deactivated, 1u);
			if (events != nullptr)
			{
				uint32_t e = atomic_add(&w.counters->num_activation_events, 1u);
				if (e < max_events) { events[e].kind = B2J_EVENT_BODY_DEACTIVATED; events[e].body = w.info[b].id; }
			}
		}
	}
};

// stable compaction of the active list: keep_scan = exclusive prefix sum of keep
struct KCompactActive
{
	DWorld w; const uint32_t *keep, *keep_scan; uint32_t *new_active;
	B2J_D void operator()(uint32_t ai) const
	{
		if (keep[ai])
		{
			uint32_t b = w.active[ai];
			uint32_t ni = keep_scan[ai];
			new_active[ni] = b;
			w.active_index[b] = ni;
		}
	}
};

// appends the bodies woken by contacts (sorted by slot for determinism) to the active list (BodyManager::ActivateBodies)
struct KActivateWoken
{
	DWorld w; const uint32_t *woken_sorted; uint32_t base; uint32_t *woken_flag; b2j_activation_event *events; uint32_t max_events;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t b = woken_sorted[k];
		BodyInfo info = w.info[b];
		w.active[base + k] = b;
		w.active_index[b] = base + k;
		woken_flag[b] = 0;
		// Body::ResetSleepTimer
		V3 points[3];
		sleep_test_points(w.shapes[info.shape], to_v3(w.position[b]), to_q4(w.rotation[b]), points);
		for (int i = 0; i < 3; ++i) w.sleep_spheres[b * 3 + i] = f4(points[i], 0.0f);
		w.sleep_timer[b] = 0.0f;
		if (events != nullptr)
		{
			uint32_t e = atomic_add(&w.counters->num_activation_events, 1u);
			if (e < max_events) { events[e].kind = B2J_EVENT_BODY_ACTIVATED; events[e].body = info.id; }
		}
	}
};

// contact removed events: manifolds of the read cache that were not persisted (ManifoldCache::ContactPointRemovedCallbacks)
struct KRemovedEvents
{
	DWorld w; NarrowCtx c;
	B2J_D void operator()(uint32_t m) const
	{
		const CachedManifold &cm = w.read_cache.manifolds[m];
		if (cm.flags & MANIFOLD_PERSISTED)
			return;
		emit_event(w, c, B2J_EVENT_CONTACT_REMOVED, cm.body1, cm.body2, cm.sub1, cm.sub2, v3_zero(), v3_zero(), 0.0f, nullptr, nullptr, 0);
	}
};

} // namespace b2j
