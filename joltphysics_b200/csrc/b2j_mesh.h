// b2j_mesh.h -- convex / sphere vs static MeshShape: walks the reference's own cooked byte buffer on the device.
//
// Restates
//   MeshShape::sCollideConvexVsMesh / sCollideSphereVsMesh     Shape/MeshShape.cpp:1124-1211, WalkTreePerTriangle :491-553
//   NodeCodecQuadTreeHalfFloat::DecodingContext::WalkTree       AABBTree/NodeCodec/NodeCodecQuadTreeHalfFloat.h:245-310
//   TriangleCodecIndexed8BitPackSOA4Flags::DecodingContext      AABBTree/TriangleCodec/...Flags.h:338-425 (21/22/21 bit vertices)
//   CollideConvexVsTriangles::Collide                           CollideConvexVsTriangles.cpp:41-157
//   CollideSphereVsTriangles::Collide                           CollideSphereVsTriangles.cpp:48-123
//   ActiveEdges::FixNormal                                      ActiveEdges.h:42-111
//   ReductionCollideShapeCollector::AddHit                      PhysicsSystem.cpp:1139-1217 (<= 32 manifolds, merge within 5 degrees)
// Child visit order (overlapping children pushed in child order, popped last first) and triangle order inside a block are kept,
// so manifold reduction sees hits in the reference's order and SubShapeIDs (block id | 3 bit triangle index) are identical.
// One thread per (convex, mesh) pair; manifold accumulation and EPA scratch live in the thread's global memory slot.
#pragma once

#include "b2j_narrowphase.h"

namespace b2j {

enum { MESH_MAX_MANIFOLDS = 32 };

// Collector of a NarrowPhaseQuery::CollideShape query (b2j_query.h: AllHitCollisionCollector<CollideShapeCollector>): when a pair is
// collided for a query every leaf / triangle hit is appended here as a CollideShapeResult instead of joining the pair's manifolds.
struct QueryCollector
{
	b2j_collide_shape_hit *hits;
	uint32_t max_hits, count;
	uint32_t body;               // TransformedShape::mBodyID of the body being collided with (CollideShapeResult::mBodyID2)
};

B2J_D void query_add_hit(QueryCollector &q, V3 point1, V3 point2, V3 axis_world, float depth, uint32_t sub1, uint32_t sub2)
{
	if (q.count < q.max_hits)
	{
		b2j_collide_shape_hit &h = q.hits[q.count];
		h.body = q.body; h.sub_shape1 = sub1; h.sub_shape2 = sub2; h.penetration_depth = depth;
		h.point1[0] = point1.x; h.point1[1] = point1.y; h.point1[2] = point1.z;
		h.point2[0] = point2.x; h.point2[1] = point2.y; h.point2[2] = point2.z;
		h.axis[0] = axis_world.x; h.axis[1] = axis_world.y; h.axis[2] = axis_world.z;
	}
	++q.count;
}

struct MeshManifold
{
	V3 normal_sum, first_normal;
	float depth;
	uint32_t sub1, sub2;
	int n;
	V3 p1[MAX_MANIFOLD_POINTS], p2[MAX_MANIFOLD_POINTS];
};

// Per thread scratch for the mesh kernel (next to the EPA scratch of the same slot)
struct MeshScratch
{
	MeshManifold manifolds[MESH_MAX_MANIFOLDS];
	ManifoldOut out[MESH_MAX_MANIFOLDS];
	V3 clip[3 * MAX_CLIP_VERTS];
};

B2J_HD float half_to_float(uint16_t h)
{
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff;
	uint32_t bits;
	if (exp == 0)
	{
		if (man == 0) bits = sign;
		else
		{
			// denormal: renormalise
			int e = -1;
			do { ++e; man <<= 1; } while ((man & 0x400) == 0);
			bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
		}
	}
	else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
	else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

B2J_HD uint32_t load_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
B2J_HD uint16_t load_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
B2J_HD float load_f32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }

struct MeshCollideCtx
{
	// shared
	int sphere;                      // 1: CollideSphereVsTriangles
	Xf transform1, transform2;       // body 1 / 2 relative to the centre of mass of body 1
	float max_separation_distance;
	bool check_active_edges;
	V3 active_edge_movement_direction;
	uint32_t sub1;                   // sub shape id of shape 1 in its body (empty = 0xffffffff; a sub shape of a compound, b2j_compound.h)
	QueryCollector *query;           // not null: a CollideShape query (hits go to the collector, no faces: ECollectFacesMode::NoFaces)
	V3 scale2;                       // ScaledShape around the mesh: node bounds and vertices are scaled on the fly (MeshShape.cpp:1150-1152, CollideConvexVsTriangles.cpp:43-45)
	// convex
	Xf transform_2_to_1;
	V3 bounds1_min, bounds1_max;                 // mBoundsOf1 (expanded)
	V3 bounds1_in2_min, bounds1_in2_max;         // mBoundsOf1InSpaceOf2
	ConvexSupport s1_excl, s1_incl;
	// sphere
	V3 sphere_center_in2;
	float radius, radius_plus_max_sep_sq;
};

// One triangle hit of the (convex or sphere, mesh) pair before it is merged into the pair's manifolds: everything that does not depend
// on the hits found before it (the warp form computes these for 32 triangles at once, then merges them in visit order)
struct MeshHit
{
	V3 world_space_normal;       // normalised penetration axis
	float depth;
	uint32_t sub1, sub2;
	int n;                       // contact point pairs of the hit, in the order ManifoldBetweenTwoFaces emits them
};

// The manifold points of ONE hit on their own (ManifoldBetweenTwoFaces only appends: what it emits does not depend on the points
// already in the manifold); pts1 / pts2: MAX_MANIFOLD_POINTS entries, clip: 3 * MAX_CLIP_VERTS
B2J_D void mesh_hit_points(const DWorld &w, V3 point1, V3 point2, V3 axis_world, const V3 *face1, int n1, const V3 *face2, int n2, V3 *pts1, V3 *pts2, int &n, V3 *clip)
{
	n = 0;
	manifold_between_two_faces(point1, point2, axis_world, w.settings.speculative_contact_distance + w.settings.manifold_tolerance, face1, n1, face2, n2, pts1, pts2, n, clip);
}

// ReductionCollideShapeCollector::AddHit for a hit whose points were computed by mesh_hit_points: manifold selection by normal, append
// (the manifold holds at most MAX_MANIFOLD_POINTS), prune above 32. Equivalent to mesh_add_hit, call for call.
B2J_D void mesh_merge_hit(const DWorld &w, MeshScratch &ms, int &num_manifolds, const MeshHit &hit, const V3 *pts1, const V3 *pts2)
{
	int mi = -1;
	for (int i = 0; i < num_manifolds; ++i)
		if (dot(hit.world_space_normal, ms.manifolds[i].first_normal) >= w.settings.contact_normal_cos_max_delta_rotation)
		{
			ms.manifolds[i].normal_sum += hit.world_space_normal;
			ms.manifolds[i].depth = fmax_(ms.manifolds[i].depth, hit.depth);
			mi = i;
			break;
		}
	if (mi < 0)
	{
		if (num_manifolds == MESH_MAX_MANIFOLDS)
		{
			mi = 0;
			for (int i = 1; i < num_manifolds; ++i)
				if (ms.manifolds[i].depth < ms.manifolds[mi].depth) mi = i;
			if (hit.depth < ms.manifolds[mi].depth)
				return;
		}
		else
			mi = num_manifolds++;
		MeshManifold &m = ms.manifolds[mi];
		m.normal_sum = hit.world_space_normal; m.first_normal = hit.world_space_normal; m.depth = hit.depth;
		m.sub1 = hit.sub1; m.sub2 = hit.sub2; m.n = 0;
	}
	MeshManifold &m = ms.manifolds[mi];
	for (int i = 0; i < hit.n && m.n < MAX_MANIFOLD_POINTS; ++i) { m.p1[m.n] = pts1[i]; m.p2[m.n] = pts2[i]; ++m.n; }
	if (m.n > 32)
		prune_contact_points(m.first_normal, m.p1, m.p2, m.n, ms.clip);
}

// ReductionCollideShapeCollector::AddHit
B2J_D void mesh_add_hit(const DWorld &w, MeshScratch &ms, int &num_manifolds, V3 point1, V3 point2, V3 axis_world, float depth, uint32_t sub1, uint32_t sub2,
	const V3 *face1, int n1, const V3 *face2, int n2)
{
	V3 world_space_normal = normalized(axis_world);
	int mi = -1;
	for (int i = 0; i < num_manifolds; ++i)
		if (dot(world_space_normal, ms.manifolds[i].first_normal) >= w.settings.contact_normal_cos_max_delta_rotation)
		{
			ms.manifolds[i].normal_sum += world_space_normal;
			ms.manifolds[i].depth = fmax_(ms.manifolds[i].depth, depth);
			mi = i;
			break;
		}
	if (mi < 0)
	{
		if (num_manifolds == MESH_MAX_MANIFOLDS)
		{
			mi = 0;
			for (int i = 1; i < num_manifolds; ++i)
				if (ms.manifolds[i].depth < ms.manifolds[mi].depth) mi = i;
			if (depth < ms.manifolds[mi].depth)
				return;
		}
		else
			mi = num_manifolds++;
		MeshManifold &m = ms.manifolds[mi];
		m.normal_sum = world_space_normal; m.first_normal = world_space_normal; m.depth = depth;
		m.sub1 = sub1; m.sub2 = sub2; m.n = 0;
	}
	MeshManifold &m = ms.manifolds[mi];
	manifold_between_two_faces(point1, point2, axis_world, w.settings.speculative_contact_distance + w.settings.manifold_tolerance, face1, n1, face2, n2, m.p1, m.p2, m.n, ms.clip);
	if (m.n > 32)
		prune_contact_points(m.first_normal, m.p1, m.p2, m.n, ms.clip);
}

// ActiveEdges::FixNormal
B2J_D V3 active_edges_fix_normal(V3 v0, V3 v1, V3 v2, V3 triangle_normal, uint32_t active_edges, V3 point, V3 normal, V3 movement_direction)
{
	float normal_length = length(normal);
	float triangle_normal_length = length(triangle_normal);
	if (dot(movement_direction, normal) * triangle_normal_length < dot(movement_direction, triangle_normal) * normal_length)
		return normal;
	if (active_edges == 0)
		return triangle_normal;
	if (dot(triangle_normal, normal) > 0.999848f * normal_length * triangle_normal_length)
		return normal;
	const float cEpsilon = 1.0e-4f;
	const float cOneMinusEpsilon = 1.0f - cEpsilon;
	uint32_t colliding_edge;
	float u, v, wv;
	cp_barycentric_tri(v0 - point, v1 - point, v2 - point, u, v, wv);
	if (u > cOneMinusEpsilon) colliding_edge = 5;
	else if (v > cOneMinusEpsilon) colliding_edge = 3;
	else if (wv > cOneMinusEpsilon) colliding_edge = 6;
	else if (u < cEpsilon) colliding_edge = 2;
	else if (v < cEpsilon) colliding_edge = 4;
	else if (wv < cEpsilon) colliding_edge = 1;
	else return triangle_normal;
	return (active_edges & colliding_edge) != 0? normal : triangle_normal;
}

// CollideConvexVsTriangles::Collide
B2J_D void mesh_collide_convex_triangle(const DWorld &w, const ShapeDesc &s1, const MeshCollideCtx &c, EpaScratch &epa, MeshScratch &ms, int &num_manifolds,
	V3 in_v0, V3 in_v1, V3 in_v2, uint32_t active_edges, uint32_t sub2)
{
	V3 v0 = mul(c.transform_2_to_1, c.scale2 * in_v0), v1 = mul(c.transform_2_to_1, c.scale2 * in_v1), v2 = mul(c.transform_2_to_1, c.scale2 * in_v2);
	V3 triangle_normal = 1.0f * cross(v1 - v0, v2 - v0);
	bool back_facing = dot(triangle_normal, v0) > 0.0f;
	if (back_facing)
		return; // EBackFaceMode::IgnoreBackFaces (default of CollideShapeSettings)
	V3 tmin = v3_min(v3_min(v0, v1), v2), tmax = v3_max(v3_max(v0, v1), v2);
	if (!aabb_overlaps(tmin, tmax, c.bounds1_min, c.bounds1_max))
		return;
	TriangleSupport triangle; triangle.v1 = v0; triangle.v2 = v1; triangle.v3_ = v2;
	V3 penetration_axis = -triangle_normal, point1, point2;
	float max_separation_distance = c.max_separation_distance;
	GjkSimplex simplex;
	int status = pen_depth_step_gjk(simplex, c.s1_excl, c.s1_excl.convex_radius + max_separation_distance, triangle, 0.0f, 1.0e-4f, penetration_axis, point1, point2);
	if (status == PEN_NOT_COLLIDING)
		return;
	if (status == PEN_INDETERMINATE)
	{
		max_separation_distance = fmin_(max_separation_distance, 1.0f);
		AddRadiusSupport a_incl; a_incl.s = c.s1_incl; a_incl.radius = max_separation_distance;
		if (!pen_depth_step_epa(epa, simplex, a_incl, triangle, 1.0e-4f, penetration_axis, point1, point2))
			return;
	}
	float penetration_depth = length(point2 - point1) - max_separation_distance;
	if (-penetration_depth >= FLT_MAX)
		return;
	float penetration_axis_len = length(penetration_axis);
	if (penetration_axis_len > 0.0f)
		point1 -= penetration_axis * (max_separation_distance / penetration_axis_len);
	if (c.check_active_edges && active_edges != 7)
	{
		V3 dir = mul_transposed(c.transform1.r, c.active_edge_movement_direction);
		penetration_axis = active_edges_fix_normal(v0, v1, v2, back_facing? triangle_normal : -triangle_normal, active_edges, point2, penetration_axis, dir);
	}
	point1 = mul(c.transform1, point1);
	point2 = mul(c.transform1, point2);
	V3 axis_world = mul(c.transform1.r, penetration_axis);
	if (c.query != nullptr) { query_add_hit(*c.query, point1, point2, axis_world, penetration_depth, c.sub1, sub2); return; }
	V3 face1[MAX_FACE_VERTS], face2[3];
	int n1 = supporting_face(w, s1, -penetration_axis, c.transform1, face1);
	face2[0] = mul(c.transform1, v0); face2[1] = mul(c.transform1, v1); face2[2] = mul(c.transform1, v2);
	mesh_add_hit(w, ms, num_manifolds, point1, point2, axis_world, penetration_depth, c.sub1, sub2, face1, n1, face2, 3);
}

// CollideSphereVsTriangles::Collide
B2J_D void mesh_collide_sphere_triangle(const DWorld &w, const MeshCollideCtx &c, MeshScratch &ms, int &num_manifolds, V3 in_v0, V3 in_v1, V3 in_v2, uint32_t active_edges, uint32_t sub2)
{
	V3 v0 = c.scale2 * in_v0 - c.sphere_center_in2, v1 = c.scale2 * in_v1 - c.sphere_center_in2, v2 = c.scale2 * in_v2 - c.sphere_center_in2;
	V3 triangle_normal = 1.0f * cross(v1 - v0, v2 - v0);
	bool back_facing = dot(triangle_normal, v0) > 0.0f;
	if (back_facing)
		return;
	uint32_t closest_feature;
	V3 point2 = cp_on_triangle<false>(v0, v1, v2, closest_feature);
	float point2_len_sq = length_sq(point2);
	if (point2_len_sq > c.radius_plus_max_sep_sq)
		return;
	float penetration_depth = c.radius - sqrt_(point2_len_sq);
	if (-penetration_depth >= FLT_MAX)
		return;
	V3 penetration_axis = normalized_or(point2, v3(0.0f, 1.0f, 0.0f));
	V3 point1 = c.radius * penetration_axis;
	const uint32_t feature_to_edges[8] = { 0, 5, 3, 1, 6, 4, 2, 0 };
	if (c.check_active_edges && closest_feature != 7 && (active_edges & feature_to_edges[closest_feature & 7]) == 0)
	{
		V3 dir = mul_transposed(c.transform2.r, c.active_edge_movement_direction);
		V3 new_penetration_axis = back_facing? triangle_normal : -triangle_normal;
		if (dot(dir, penetration_axis) * length(new_penetration_axis) >= dot(dir, new_penetration_axis))
			penetration_axis = new_penetration_axis;
	}
	point1 = mul(c.transform2, c.sphere_center_in2 + point1);
	point2 = mul(c.transform2, c.sphere_center_in2 + point2);
	V3 axis_world = mul(c.transform2.r, penetration_axis);
	if (c.query != nullptr) { query_add_hit(*c.query, point1, point2, axis_world, penetration_depth, c.sub1, sub2); return; }
	V3 face2[3];
	face2[0] = mul(c.transform2, c.sphere_center_in2 + v0);
	face2[1] = mul(c.transform2, c.sphere_center_in2 + v1);
	face2[2] = mul(c.transform2, c.sphere_center_in2 + v2);
	mesh_add_hit(w, ms, num_manifolds, point1, point2, axis_world, penetration_depth, c.sub1, sub2, nullptr, 0, face2, 3);
}

// NodeCodecQuadTreeHalfFloat: does child ch of a 64 byte node overlap the query volume of the pair (AABox4VsSphere / AABox4VsBox)
B2J_D bool mesh_child_overlaps(const MeshCollideCtx &cc, const uint8_t *node, int ch)
{
	float mnx = half_to_float(load_u16(node + 0 + 2 * ch)), mny = half_to_float(load_u16(node + 8 + 2 * ch)), mnz = half_to_float(load_u16(node + 16 + 2 * ch));
	float mxx = half_to_float(load_u16(node + 24 + 2 * ch)), mxy = half_to_float(load_u16(node + 32 + 2 * ch)), mxz = half_to_float(load_u16(node + 40 + 2 * ch));
	// AABox4Scale (positive scales: minimum and maximum keep their roles; times one is exact)
	mnx = cc.scale2.x * mnx; mny = cc.scale2.y * mny; mnz = cc.scale2.z * mnz;
	mxx = cc.scale2.x * mxx; mxy = cc.scale2.y * mxy; mxz = cc.scale2.z * mxz;
	if (cc.sphere)
	{
		V3 p = cc.sphere_center_in2;
		float cx = fmin_(fmax_(p.x, mnx), mxx), cy = fmin_(fmax_(p.y, mny), mxy), cz = fmin_(fmax_(p.z, mnz), mxz);
		float d = square(cx - p.x) + square(cy - p.y) + square(cz - p.z);
		return d <= cc.radius_plus_max_sep_sq;
	}
	const V3 &bmn = cc.bounds1_in2_min, &bmx = cc.bounds1_in2_max;
	return !((bmn.x > mxx || mnx > bmx.x) || (bmn.y > mxy || mny > bmx.y) || (bmn.z > mxz || mnz > bmx.z));
}

// vertices, active edge flags and sub shape id of triangle t of a triangle block (TriangleCodecIndexed8BitPackSOA4Flags)
B2J_D void mesh_decode_triangle(const uint8_t *tree, uint32_t block_id, uint32_t t, uint32_t block_id_bits, V3 tri_offset, V3 tri_scale, V3 v[3], uint32_t &active_edges, uint32_t &sub2)
{
	const uint8_t *block_start = tree + ((size_t)block_id << 2);
	uint32_t header_flags = load_u32(block_start);
	const uint8_t *vertices = block_start + ((size_t)(header_flags & 0x1fffffffu) << 2);
	const uint8_t *blk = block_start + 4 + 16 * (t >> 2);
	uint32_t lane = t & 3;
	for (int vi = 0; vi < 3; ++vi)
	{
		uint32_t idx = blk[4 * vi + lane];
		uint32_t c1 = load_u32(vertices + 8 * idx), c2 = load_u32(vertices + 8 * idx + 4);
		uint32_t xc = c1 & 0x1fffffu, yc = (c1 >> 21) | ((c2 >> 21) << 11), zc = c2 & 0x1fffffu;
		v[vi] = v3((float)(int32_t)xc * tri_scale.x + tri_offset.x, (float)(int32_t)yc * tri_scale.y + tri_offset.y, (float)(int32_t)zc * tri_scale.z + tri_offset.z);
	}
	active_edges = ((uint32_t)blk[12 + lane] >> 5) & 7;
	uint32_t block_sub = block_id_bits >= 32? block_id : ((0xffffffffu & ~((1u << block_id_bits) - 1u)) | block_id);
	sub2 = (block_sub & ~(7u << block_id_bits)) | (t << block_id_bits);
}

// Everything of MeshShape::sCollideConvexVsMesh / sCollideSphereVsMesh that is computed once per (convex, mesh) pair;
// transform1 / transform2: the centre of mass transforms the dispatch hands to sCollideConvexVsMesh for the two shapes BEFORE their own
// decorators are peeled (a body's transform, or a compound's sub shape transform), relative to the centre of mass of body 1
B2J_D MeshCollideCtx mesh_collide_ctx_from(const DWorld &w, const ShapeDesc &s1, const Xf &transform1, const ShapeDesc &s2, const Xf &transform2,
	float max_separation_distance, V3 active_edge_movement_direction, uint32_t sub1)
{
	MeshCollideCtx cc;
	cc.transform1 = shape_transform(s1, transform1);
	cc.transform2 = shape_transform(s2, transform2);
	cc.scale2 = s2.scale;
	cc.query = nullptr;
	cc.sub1 = sub1;
	cc.max_separation_distance = max_separation_distance;
	cc.check_active_edges = w.settings.check_active_edges != 0;
	cc.active_edge_movement_direction = active_edge_movement_direction;
	cc.sphere = s1.kind == B2J_SHAPE_SPHERE;
	if (cc.sphere)
	{
		cc.sphere_center_in2 = mul_transposed(cc.transform2.r, cc.transform1.t - cc.transform2.t);
		cc.radius = 1.0f * s1.radius;
		cc.radius_plus_max_sep_sq = square(cc.radius + cc.max_separation_distance);
	}
	else
	{
		// inverse_transform2 = T2^-1; transform1_to_2 = inverse_transform2 * T1; mTransform2To1 = transform1_to_2^-1
		M33 r2t = transposed(cc.transform2.r);
		Xf inv2 = xf(r2t, -mul(r2t, cc.transform2.t));
		Xf t1_to_2 = mul(inv2, cc.transform1);
		M33 r12t = transposed(t1_to_2.r);
		cc.transform_2_to_1 = xf(r12t, -mul(r12t, t1_to_2.t));
		cc.bounds1_min = s1.local_min - v3_rep(cc.max_separation_distance);
		cc.bounds1_max = s1.local_max + v3_rep(cc.max_separation_distance);
		// AABox::Transformed(transform1_to_2)
		V3 nmin = t1_to_2.t, nmax = t1_to_2.t;
		for (int col = 0; col < 3; ++col)
		{
			V3 cv = m33_col(t1_to_2.r, col);
			V3 a = cv * v3_get(cc.bounds1_min, col), b = cv * v3_get(cc.bounds1_max, col);
			nmin += v3_min(a, b);
			nmax += v3_max(a, b);
		}
		cc.bounds1_in2_min = nmin; cc.bounds1_in2_max = nmax;
		cc.s1_excl = make_support(w, s1, SUPPORT_EXCLUDE_CONVEX_RADIUS);
		cc.s1_incl = make_support(w, s1, SUPPORT_INCLUDE_CONVEX_RADIUS);
	}

	return cc;
}

// PhysicsSystem::ProcessBodyPair: mActiveEdgeMovementDirection = velocity of body 1 relative to body 2 (PhysicsSystem.cpp:1106-1108)
B2J_D V3 pair_movement_direction(const DWorld &w, const CollideItem &item, const BodyInfo &i1, const BodyInfo &i2)
{
	V3 lv1 = i1.motion_type != B2J_MOTION_STATIC? to_v3(w.linear_velocity[item.b1]) : v3_zero();
	V3 lv2 = i2.motion_type != B2J_MOTION_STATIC? to_v3(w.linear_velocity[item.b2]) : v3_zero();
	return lv1 - lv2;
}

B2J_D MeshCollideCtx mesh_collide_ctx(const DWorld &w, const CollideItem &item, const BodyInfo &i1, const BodyInfo &i2, const ShapeDesc &s1)
{
	V3 x1 = to_v3(w.position[item.b1]), x2 = to_v3(w.position[item.b2]);
	float max_separation_distance = ((i1.flags | i2.flags) & B2J_BODY_SENSOR)? 0.0f : w.settings.speculative_contact_distance;
	return mesh_collide_ctx_from(w, s1, xf(m33_rotation(to_q4(w.rotation[item.b1])), v3_zero()), w.shapes[i2.shape], xf(m33_rotation(to_q4(w.rotation[item.b2])), x2 + (-x1)),
		max_separation_distance, pair_movement_direction(w, item, i1, i2), 0xffffffffu);
}

// MeshShape::WalkTreePerTriangle for one (convex, mesh) pair, serial: every hit joins ms / num_manifolds (mesh_add_hit)
B2J_D void mesh_walk_serial(const DWorld &w, const ShapeDesc &s1, const ShapeDesc &s2, const MeshCollideCtx &cc, EpaScratch &epa, MeshScratch &ms, int &num_manifolds)
{
	const uint8_t *tree = w.mesh_bytes + s2.mesh_offset;
	// NodeCodec header (32 B): root bounds min/max, root properties, block id bits; then TriangleHeader: offset, scale
	uint32_t root_properties = load_u32(tree + 24);
	uint32_t block_id_bits = tree[28];
	V3 tri_offset = v3(load_f32(tree + 32), load_f32(tree + 36), load_f32(tree + 40));
	V3 tri_scale = v3(load_f32(tree + 44), load_f32(tree + 48), load_f32(tree + 52));
	uint32_t stack[128];
	int top = 0;
	stack[0] = root_properties;
	do
	{
		uint32_t node_properties = stack[top];
		uint32_t tri_count = node_properties >> 28;
		if (tri_count == 0)
		{
			const uint8_t *node = tree + ((size_t)node_properties << 2);
			uint32_t props[4];
			int n = 0;
			for (int ch = 0; ch < 4; ++ch)
				if (mesh_child_overlaps(cc, node, ch)) props[n++] = load_u32(node + 48 + 4 * ch);
			for (int j = 0; j < n && top + j < 128; ++j) stack[top + j] = props[j];
			top += n;
		}
		else if (tri_count != 15)
		{
			uint32_t block_id = node_properties & 0x0fffffffu;
			for (uint32_t t = 0; t < tri_count; ++t)
			{
				V3 v[3];
				uint32_t active_edges, sub2;
				mesh_decode_triangle(tree, block_id, t, block_id_bits, tri_offset, tri_scale, v, active_edges, sub2);
				if (cc.sphere)
					mesh_collide_sphere_triangle(w, cc, ms, num_manifolds, v[0], v[1], v[2], active_edges, sub2);
				else
					mesh_collide_convex_triangle(w, s1, cc, epa, ms, num_manifolds, v[0], v[1], v[2], active_edges, sub2);
			}
		}
		--top;
	}
	while (top >= 0);
}

// ProcessBodyPair after the collector is full: normalise the summed normals, prune to 4, add the contacts
B2J_D void mesh_finish_pair(const DWorld &w, const NarrowCtx &c, const CollideItem &item, MeshScratch &ms, int num_manifolds)
{
	for (int i = 0; i < num_manifolds; ++i)
	{
		MeshManifold &m = ms.manifolds[i];
		V3 normal = normalized(m.normal_sum);
		if (m.n > 4)
			prune_contact_points(normal, m.p1, m.p2, m.n, ms.clip);
		ManifoldOut &o = ms.out[i];
		o.normal = normal; o.depth = m.depth; o.sub1 = m.sub1; o.sub2 = m.sub2; o.n = m.n;
		for (int p = 0; p < m.n; ++p) { o.p1[p] = m.p1[p]; o.p2[p] = m.p2[p]; }
	}
	add_manifolds(w, c, item, ms.out, num_manifolds);
}

B2J_D void compound_collide_pair(const DWorld &w, const NarrowCtx &c, const CollideItem &item, EpaScratch &epa, MeshScratch &ms); // b2j_compound.h

struct KCollideMesh
{
	DWorld w; NarrowCtx c; MeshScratch *mesh_scratch;
	B2J_D void run(uint32_t k, bool valid, uint32_t slot, EpaStorageFull &epa_storage) const
	{
		(void)valid;
		EpaScratch epa = epa_storage.view();
		CollideItem item = c.collide_mesh[k];
		BodyInfo i1 = w.info[item.b1], i2 = w.info[item.b2];
		const ShapeDesc &s1 = w.shapes[i1.shape], &s2 = w.shapes[i2.shape];
		MeshScratch &ms = mesh_scratch[slot];
		if (s1.kind == B2J_SHAPE_COMPOUND || s2.kind == B2J_SHAPE_COMPOUND)
		{
			compound_collide_pair(w, c, item, epa, ms);
			return;
		}
		if (s2.kind != B2J_SHAPE_MESH || s1.kind == B2J_SHAPE_MESH)
			return; // mesh as body 1 (sReversedCollideShape) is not on the path: meshes are static, body 1 has the higher motion type

		MeshCollideCtx cc = mesh_collide_ctx(w, item, i1, i2, s1);
		int num_manifolds = 0;
		mesh_walk_serial(w, s1, s2, cc, epa, ms, num_manifolds);
		mesh_finish_pair(w, c, item, ms, num_manifolds);
	}
};

#if !defined(B2J_HOSTSIM) && defined(__CUDACC__)
// ---- the same pair with all 32 lanes of the warp (the form the GPU runs; KCollideMesh::run above stays the serial statement of the
// algorithm and what tests/hostsim executes) -----------------------------------------------------------------------------------
// Lane 0 walks the tree in the reference's order and hands out the triangles it meets 32 at a time; every lane then owns ONE triangle:
// decoding, culling, GJK with the warp kept in lockstep by votes (as in the convex pair kernels), the supporting face of the convex
// shape and the polygon clipping of the hit all run 32 wide. Only what depends on the hits found before -- choosing / creating the
// manifold a hit joins (first normal wins, deepest manifolds survive, ReductionCollideShapeCollector::AddHit) -- is done lane after
// lane in triangle order, and EPA (rare: a deep hit) uses the warp's one shared memory scratch lane after lane. Results are bit
// identical to the serial form: the per triangle arithmetic is the same code and does not depend on earlier hits.
struct MeshWarpShared
{
	uint32_t cand_block[32];     // triangle block of the candidates of this round
	uint32_t cand_tri[32];       // triangle index inside the block
	int count, done, num_manifolds;
};

struct KCollideMeshWarp
{
	DWorld w; NarrowCtx c; MeshScratch *mesh_scratch;
	B2J_D void run_warp(uint32_t k, uint32_t slot, EpaStorageFull &epa_storage, MeshWarpShared &sh) const
	{
		const uint32_t lane = threadIdx.x & 31;
		CollideItem item = c.collide_mesh[k];
		BodyInfo i1 = w.info[item.b1], i2 = w.info[item.b2];
		const ShapeDesc &s1 = w.shapes[i1.shape], &s2 = w.shapes[i2.shape];
		MeshScratch &ms = mesh_scratch[slot];
		if (s1.kind == B2J_SHAPE_COMPOUND || s2.kind == B2J_SHAPE_COMPOUND)
		{
			// a pair with a StaticCompoundShape: lane 0 runs the sub shape loops (b2j_compound.h)
			if (lane == 0) { EpaScratch epa = epa_storage.view(); compound_collide_pair(w, c, item, epa, ms); }
			__syncwarp();
			return;
		}
		if (s2.kind != B2J_SHAPE_MESH || s1.kind == B2J_SHAPE_MESH)
			return; // (uniform over the warp)
		const MeshCollideCtx cc = mesh_collide_ctx(w, item, i1, i2, s1);

		const uint8_t *tree = w.mesh_bytes + s2.mesh_offset;
		const uint32_t root_properties = load_u32(tree + 24);
		const uint32_t block_id_bits = tree[28];
		const V3 tri_offset = v3(load_f32(tree + 32), load_f32(tree + 36), load_f32(tree + 40));
		const V3 tri_scale = v3(load_f32(tree + 44), load_f32(tree + 48), load_f32(tree + 52));

		// lane 0: the tree walk, suspended whenever 32 candidate triangles are queued
		uint32_t stack[128];
		int top = -1;
		uint32_t leaf_block = 0, leaf_count = 0, leaf_next = 0;
		if (lane == 0)
		{
			stack[0] = root_properties; top = 0;
			sh.num_manifolds = 0; sh.done = 0;
		}
		__syncwarp();
		for (;;)
		{
			if (lane == 0)
			{
				int n = 0;
				while (n < 32)
				{
					if (leaf_next < leaf_count) { sh.cand_block[n] = leaf_block; sh.cand_tri[n] = leaf_next++; ++n; continue; }
					if (top < 0) break;
					uint32_t node_properties = stack[top];
					uint32_t tri_count = node_properties >> 28;
					if (tri_count == 0)
					{
						const uint8_t *node = tree + ((size_t)node_properties << 2);
						uint32_t props[4];
						int nh = 0;
						for (int ch = 0; ch < 4; ++ch)
							if (mesh_child_overlaps(cc, node, ch)) props[nh++] = load_u32(node + 48 + 4 * ch);
						for (int j = 0; j < nh && top + j < 128; ++j) stack[top + j] = props[j];
						top += nh;
						--top;
					}
					else
					{
						--top;
						if (tri_count != 15) { leaf_block = node_properties & 0x0fffffffu; leaf_count = tri_count; leaf_next = 0; }
					}
				}
				sh.count = n;
				if (n == 0) sh.done = 1;
			}
			__syncwarp();
			if (sh.done)
				break;
			const int count = sh.count;

			// ---- one triangle per lane
			bool alive = (int)lane < count;
			V3 v0 = v3_zero(), v1 = v3_zero(), v2 = v3_zero(), triangle_normal = v3_zero();
			uint32_t active_edges = 0, sub2 = 0;
			bool back_facing = false;
			if (alive)
			{
				V3 v[3];
				mesh_decode_triangle(tree, sh.cand_block[lane], sh.cand_tri[lane], block_id_bits, tri_offset, tri_scale, v, active_edges, sub2);
				if (cc.sphere) { v0 = cc.scale2 * v[0] - cc.sphere_center_in2; v1 = cc.scale2 * v[1] - cc.sphere_center_in2; v2 = cc.scale2 * v[2] - cc.sphere_center_in2; }
				else { v0 = mul(cc.transform_2_to_1, cc.scale2 * v[0]); v1 = mul(cc.transform_2_to_1, cc.scale2 * v[1]); v2 = mul(cc.transform_2_to_1, cc.scale2 * v[2]); }
				triangle_normal = 1.0f * cross(v1 - v0, v2 - v0);
				back_facing = dot(triangle_normal, v0) > 0.0f;
				if (back_facing)
					alive = false; // EBackFaceMode::IgnoreBackFaces
			}
			MeshHit hit; hit.n = 0; hit.depth = 0.0f; hit.sub1 = 0xffffffffu; hit.sub2 = sub2; hit.world_space_normal = v3_zero();
			V3 pts1[MAX_MANIFOLD_POINTS], pts2[MAX_MANIFOLD_POINTS];
			bool have_hit = false;
			if (cc.sphere)
			{
				// CollideSphereVsTriangles::Collide (no GJK: closest point on the triangle)
				if (alive)
				{
					uint32_t closest_feature;
					V3 point2 = cp_on_triangle<false>(v0, v1, v2, closest_feature);
					float point2_len_sq = length_sq(point2);
					float penetration_depth = cc.radius - sqrt_(point2_len_sq);
					if (!(point2_len_sq > cc.radius_plus_max_sep_sq) && !(-penetration_depth >= FLT_MAX))
					{
						V3 penetration_axis = normalized_or(point2, v3(0.0f, 1.0f, 0.0f));
						V3 point1 = cc.radius * penetration_axis;
						const uint32_t feature_to_edges[8] = { 0, 5, 3, 1, 6, 4, 2, 0 };
						if (cc.check_active_edges && closest_feature != 7 && (active_edges & feature_to_edges[closest_feature & 7]) == 0)
						{
							V3 dir = mul_transposed(cc.transform2.r, cc.active_edge_movement_direction);
							V3 new_penetration_axis = back_facing? triangle_normal : -triangle_normal;
							if (dot(dir, penetration_axis) * length(new_penetration_axis) >= dot(dir, new_penetration_axis))
								penetration_axis = new_penetration_axis;
						}
						point1 = mul(cc.transform2, cc.sphere_center_in2 + point1);
						point2 = mul(cc.transform2, cc.sphere_center_in2 + point2);
						V3 axis_world = mul(cc.transform2.r, penetration_axis);
						V3 face2[3];
						face2[0] = mul(cc.transform2, cc.sphere_center_in2 + v0);
						face2[1] = mul(cc.transform2, cc.sphere_center_in2 + v1);
						face2[2] = mul(cc.transform2, cc.sphere_center_in2 + v2);
						V3 clip[3 * MAX_CLIP_VERTS];
						hit.world_space_normal = normalized(axis_world); hit.depth = penetration_depth;
						mesh_hit_points(w, point1, point2, axis_world, nullptr, 0, face2, 3, pts1, pts2, hit.n, clip);
						have_hit = true;
					}
				}
			}
			else
			{
				// CollideConvexVsTriangles::Collide
				if (alive)
				{
					V3 tmin = v3_min(v3_min(v0, v1), v2), tmax = v3_max(v3_max(v0, v1), v2);
					if (!aabb_overlaps(tmin, tmax, cc.bounds1_min, cc.bounds1_max))
						alive = false;
				}
				TriangleSupport triangle; triangle.v1 = v0; triangle.v2 = v1; triangle.v3_ = v2;
				V3 penetration_axis = alive? -triangle_normal : v3(1.0f, 0.0f, 0.0f), point1 = v3_zero(), point2 = v3_zero();
				float max_separation_distance = cc.max_separation_distance;
				GjkSimplex simplex;
				int status = pen_depth_step_gjk<true>(simplex, cc.s1_excl, cc.s1_excl.convex_radius + max_separation_distance, triangle, 0.0f, 1.0e-4f, penetration_axis, point1, point2, alive);
				if (!alive) status = PEN_NOT_COLLIDING;
				bool collided = status == PEN_COLLIDING;
				// deep hits: EPA on the warp's shared memory scratch, lane after lane
				uint32_t epa_mask = __ballot_sync(0xffffffffu, status == PEN_INDETERMINATE);
				while (epa_mask != 0)
				{
					uint32_t l = (uint32_t)__ffs((int)epa_mask) - 1;
					if (lane == l)
					{
						max_separation_distance = fmin_(max_separation_distance, 1.0f);
						AddRadiusSupport a_incl; a_incl.s = cc.s1_incl; a_incl.radius = max_separation_distance;
						EpaScratch epa = epa_storage.view();
						collided = pen_depth_step_epa(epa, simplex, a_incl, triangle, 1.0e-4f, penetration_axis, point1, point2);
					}
					__syncwarp();
					epa_mask &= epa_mask - 1;
				}
				if (collided)
				{
					float penetration_depth = length(point2 - point1) - max_separation_distance;
					if (!(-penetration_depth >= FLT_MAX))
					{
						float penetration_axis_len = length(penetration_axis);
						if (penetration_axis_len > 0.0f)
							point1 -= penetration_axis * (max_separation_distance / penetration_axis_len);
						if (cc.check_active_edges && active_edges != 7)
						{
							V3 dir = mul_transposed(cc.transform1.r, cc.active_edge_movement_direction);
							penetration_axis = active_edges_fix_normal(v0, v1, v2, back_facing? triangle_normal : -triangle_normal, active_edges, point2, penetration_axis, dir);
						}
						point1 = mul(cc.transform1, point1);
						point2 = mul(cc.transform1, point2);
						V3 axis_world = mul(cc.transform1.r, penetration_axis);
						V3 face1[MAX_FACE_VERTS], face2[3];
						int n1 = supporting_face(w, s1, -penetration_axis, cc.transform1, face1);
						face2[0] = mul(cc.transform1, v0); face2[1] = mul(cc.transform1, v1); face2[2] = mul(cc.transform1, v2);
						V3 clip[3 * MAX_CLIP_VERTS];
						hit.world_space_normal = normalized(axis_world); hit.depth = penetration_depth;
						mesh_hit_points(w, point1, point2, axis_world, face1, n1, face2, 3, pts1, pts2, hit.n, clip);
						have_hit = true;
					}
				}
			}
			// ---- the hits join the pair's manifolds in triangle order
			uint32_t hit_mask = __ballot_sync(0xffffffffu, have_hit);
			while (hit_mask != 0)
			{
				uint32_t l = (uint32_t)__ffs((int)hit_mask) - 1;
				if (lane == l)
				{
					int nm = sh.num_manifolds;
					mesh_merge_hit(w, ms, nm, hit, pts1, pts2);
					sh.num_manifolds = nm;
				}
				__syncwarp();
				hit_mask &= hit_mask - 1;
			}
		}
		// ProcessBodyPair: normalise the summed normals, prune to 4, add the contacts
		if (lane == 0)
		{
			int num_manifolds = sh.num_manifolds;
			for (int i = 0; i < num_manifolds; ++i)
			{
				MeshManifold &m = ms.manifolds[i];
				V3 normal = normalized(m.normal_sum);
				if (m.n > 4)
					prune_contact_points(normal, m.p1, m.p2, m.n, ms.clip);
				ManifoldOut &o = ms.out[i];
				o.normal = normal; o.depth = m.depth; o.sub1 = m.sub1; o.sub2 = m.sub2; o.n = m.n;
				for (int p = 0; p < m.n; ++p) { o.p1[p] = m.p1[p]; o.p2[p] = m.p2[p]; }
			}
			add_manifolds(w, c, item, ms.out, num_manifolds);
		}
		__syncwarp();
	}
};
#endif

} // namespace b2j
