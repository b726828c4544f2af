// b2j_query.h -- batched queries on the device broadphase trees and shapes: closest hit ray casts and AABox overlap
// queries, thousands per call (the RL observation pattern), reusing the LBVH of every broadphase layer and the shape tables of the step.
//
// Restates
//   NarrowPhaseQuery::CastRay (closest hit)        Jolt/Physics/Collision/NarrowPhaseQuery.cpp:19-81
//   TransformedShape::CastRay, RayCast::Transformed Jolt/Physics/Collision/TransformedShape.cpp, RayCast.h:23-28
//   QuadTree::CastRay / CollideAABox               Jolt/Physics/Collision/BroadPhase/QuadTree.cpp:1079-1197 (closest child first, early out fraction)
//   RayAABox, RayInvDirection                      Jolt/Geometry/RayAABox.h
//   RaySphere / RayCylinder / RayCapsule, FindRoot Jolt/Geometry/RaySphere.h, RayCylinder.h, RayCapsule.h, Jolt/Math/FindRoot.h
//   BoxShape / SphereShape / CapsuleShape::CastRay Shape/BoxShape.cpp:177-188, SphereShape.cpp:222-232, CapsuleShape.cpp:287-298
//   ConvexHullShape::CastRayHelper                 Shape/ConvexHullShape.cpp:883-995
//   NarrowPhaseQuery::CollideShape                 Jolt/Physics/Collision/NarrowPhaseQuery.cpp:219-293, TransformedShape::CollideShape
//                                                  TransformedShape.cpp:70-86 (the pair itself: b2j_compound.h / b2j_mesh.h, as in the step)
//   QuadTree::CollideSphere / CollidePoint         BroadPhase/QuadTree.cpp:1199-1275, AABox4VsSphere / AABox4VsPoint Jolt/Geometry/AABox4.h
//   MeshShape::CastRay, RayTriangle                Shape/MeshShape.cpp:696-760, Jolt/Geometry/RayTriangle.h, TriangleCodec...Flags.h TestRay
// One thread per query. Results are defined by the true body state (pose + shape), not by the tree: a closest hit is the same
// whatever order the tree is walked in.
#pragma once

#include "b2j_mesh.h"
#include "b2j_compound.h"

namespace b2j {

struct RayInvDir { V3 inv; bool px, py, pz; };

// RayInvDirection::Set
B2J_HD RayInvDir ray_inv_direction(V3 d)
{
	RayInvDir r;
	r.px = fabs_(d.x) <= 1.0e-20f; r.py = fabs_(d.y) <= 1.0e-20f; r.pz = fabs_(d.z) <= 1.0e-20f;
	r.inv = v3(1.0f / (r.px? 1.0f : d.x), 1.0f / (r.py? 1.0f : d.y), 1.0f / (r.pz? 1.0f : d.z));
	return r;
}

// RayAABox: fraction at which the ray enters the box (can be negative when the origin is inside), FLT_MAX = no hit
B2J_HD float ray_aabox(V3 o, const RayInvDir &r, V3 mn, V3 mx)
{
	V3 t1 = (mn - o) * r.inv, t2 = (mx - o) * r.inv;
	float tminx = r.px? -FLT_MAX : fmin_(t1.x, t2.x), tminy = r.py? -FLT_MAX : fmin_(t1.y, t2.y), tminz = r.pz? -FLT_MAX : fmin_(t1.z, t2.z);
	float tmaxx = r.px? FLT_MAX : fmax_(t1.x, t2.x), tmaxy = r.py? FLT_MAX : fmax_(t1.y, t2.y), tmaxz = r.pz? FLT_MAX : fmax_(t1.z, t2.z);
	float t_min = fmax_(fmax_(tminx, tminy), tminz), t_max = fmin_(fmin_(tmaxx, tmaxy), tmaxz);
	bool no_intersection = t_min > t_max || t_max < 0.0f;
	if (r.px && (o.x < mn.x || o.x > mx.x)) no_intersection = true;
	if (r.py && (o.y < mn.y || o.y > mx.y)) no_intersection = true;
	if (r.pz && (o.z < mn.z || o.z > mx.z)) no_intersection = true;
	return no_intersection? FLT_MAX : t_min;
}

// FindRoot (a x^2 + b x + c = 0), returns the number of roots
B2J_HD int find_root(float a, float b, float c, float &x1, float &x2)
{
	if (a == 0.0f)
	{
		if (b == 0.0f)
		{
			if (c == 0.0f) { x1 = x2 = 0.0f; return 1; }
			return 0;
		}
		x1 = x2 = -c / b;
		return 1;
	}
	float det = b * b - (4.0f * a) * c; // DifferenceOfProducts without FMA
	if (det < 0.0f)
		return 0;
	float q = (b + (b < 0.0f? -1.0f : 1.0f) * sqrt_(det)) / -2.0f;
	x1 = q / a;
	if (q == 0.0f) { x2 = x1; return 1; }
	x2 = c / q;
	return 2;
}

// RaySphere (closest fraction, 0 when the origin is inside, FLT_MAX = no hit)
B2J_HD float ray_sphere(V3 o, V3 d, V3 center, float radius)
{
	V3 center_origin = o - center;
	float a = length_sq(d);
	float b = 2.0f * dot(d, center_origin);
	float c = length_sq(center_origin) - radius * radius;
	float f1, f2;
	if (find_root(a, b, c, f1, f2) == 0)
		return c <= 0.0f? 0.0f : FLT_MAX;
	if (f1 > f2) { float t = f1; f1 = f2; f2 = t; }
	if (f1 >= 0.0f) return f1;
	if (f2 >= 0.0f) return 0.0f;
	return FLT_MAX;
}

// RayCylinder (infinite cylinder along y)
B2J_HD float ray_cylinder(V3 o, V3 d, float radius)
{
	V3 origin_xz = v3(o.x, 0.0f, o.z);
	float origin_xz_len_sq = length_sq(origin_xz);
	float r_sq = square(radius);
	if (origin_xz_len_sq > r_sq)
	{
		V3 direction_xz = v3(d.x, 0.0f, d.z);
		float a = length_sq(direction_xz);
		float b = 2.0f * dot(origin_xz, direction_xz);
		float c = origin_xz_len_sq - r_sq;
		float f1, f2;
		if (find_root(a, b, c, f1, f2) == 0)
			return FLT_MAX;
		float f = fmin_(f1, f2);
		if (f >= 0.0f)
			return f;
		return FLT_MAX;
	}
	return 0.0f;
}

// RayCylinder (finite, RayCylinder.h:56-107): the infinite cylinder, then the cap the ray travels towards
B2J_HD float ray_cylinder_finite(V3 o, V3 d, float half_height, float radius)
{
	float fraction = ray_cylinder(o, d, radius);
	if (fraction == FLT_MAX)
		return FLT_MAX;
	if (fabs_(o.y + fraction * d.y) <= half_height)
		return fraction;
	if (d.y != 0.0f)
	{
		float plane_fraction = d.y < 0.0f? (half_height - o.y) / d.y : -(half_height + o.y) / d.y;
		if (plane_fraction >= 0.0f)
		{
			V3 point = o + plane_fraction * d;
			if (square(point.x) + square(point.z) <= square(radius))
				return plane_fraction;
		}
	}
	return FLT_MAX;
}

// RayCapsule
B2J_HD float ray_capsule(V3 o, V3 d, float half_height, float radius)
{
	float cylinder = ray_cylinder(o, d, radius);
	if (cylinder == FLT_MAX)
		return FLT_MAX;
	if (fabs_(o.y + cylinder * d.y) <= half_height)
		return cylinder;
	V3 sphere_center = v3(0.0f, half_height, 0.0f);
	float upper = ray_sphere(o, d, sphere_center, radius);
	float lower = ray_sphere(o, d, -sphere_center, radius);
	return fmin_(upper, lower);
}

// ConvexHullShape::CastRayHelper
B2J_D bool ray_hull(const DWorld &w, const ShapeDesc &s, V3 o, V3 d, float &out_min_fraction)
{
	if (s.hull_num_faces == 2)
	{
		// flat hull: plane + edge tests
		F4 p = w.hull_planes[s.hull_face_offset];
		V3 plane_normal = to_v3(p);
		float direction_projection = dot(d, plane_normal);
		if (fabs_(direction_projection) >= 1.0e-12f)
		{
			float distance_to_plane = dot(o, plane_normal) + p.w;
			float fraction = -distance_to_plane / direction_projection;
			if (fraction < 0.0f || fraction > 1.0f)
				return false;
			V3 intersection_point = o + fraction * d;
			uint32_t face = w.hull_faces[s.hull_face_offset];
			uint32_t first = face & 0xffffu, num = face >> 16;
			const uint8_t *vtx = w.hull_vtx + s.hull_vtx_offset + first;
			V3 p1 = to_v3(w.hull_points[s.hull_point_offset + vtx[num]]); // (the reference reads one past the face: the first vertex of the next face)
			for (uint32_t v = 0; v < num; ++v)
			{
				V3 p2 = to_v3(w.hull_points[s.hull_point_offset + vtx[v]]);
				if (dot(cross(p2 - p1, intersection_point - p1), plane_normal) < 0.0f)
					return false;
				p1 = p2;
			}
			out_min_fraction = fraction;
			return true;
		}
		return false;
	}
	int fractions_set = 0;
	bool all_inside = true;
	float min_fraction = 0.0f, max_fraction = 1.0f + FLT_EPSILON;
	for (uint32_t f = 0; f < s.hull_num_faces; ++f)
	{
		F4 p = w.hull_planes[s.hull_face_offset + f];
		V3 plane_normal = to_v3(p);
		float distance_to_plane = dot(o, plane_normal) + p.w;
		bool is_outside = distance_to_plane > 0.0f;
		all_inside = all_inside && !is_outside;
		float direction_projection = dot(d, plane_normal);
		if (fabs_(direction_projection) >= 1.0e-12f)
		{
			float fraction = -distance_to_plane / direction_projection;
			if (direction_projection < 0.0f) { min_fraction = fmax_(fraction, min_fraction); fractions_set |= 1; }
			else { max_fraction = fmin_(fraction, max_fraction); fractions_set |= 2; }
		}
		else if (is_outside)
			return false;
	}
	if (fractions_set == 3)
	{
		out_min_fraction = min_fraction;
		return min_fraction <= max_fraction && max_fraction >= 0.0f;
	}
	out_min_fraction = 0.0f;
	return all_inside;
}

// RayTriangle (Moeller-Trumbore as the reference writes it), FLT_MAX = no hit
B2J_HD float ray_triangle(V3 o, V3 d, V3 v0, V3 v1, V3 v2)
{
	V3 e1 = v1 - v0, e2 = v2 - v0;
	V3 p = cross(d, e2);
	float det = dot(e1, p);
	bool det_near_zero = fabs_(det) < 1.0e-12f;
	if (det_near_zero) det = 1.0f;
	V3 s = o - v0;
	float u = dot(s, p) / det;
	V3 q = cross(s, e1);
	float v = dot(d, q) / det;
	float t = dot(e2, q) / det;
	bool no_intersection = det_near_zero || u < 0.0f || v < 0.0f || u + v > 1.0f || t < 0.0f;
	return no_intersection? FLT_MAX : t;
}

// MeshShape::CastRay: closest triangle along the ray, children visited closest first with the best fraction as early out
B2J_D bool ray_mesh(const DWorld &w, const ShapeDesc &s, V3 o, V3 d, float &io_fraction, uint32_t &out_sub)
{
	const uint8_t *tree = w.mesh_bytes + s.mesh_offset;
	uint32_t root_properties = load_u32(tree + 24);
	uint32_t block_id_bits = tree[28];
	V3 tri_offset = v3(load_f32(tree + 32), load_f32(tree + 36), load_f32(tree + 40));
	V3 tri_scale = v3(load_f32(tree + 44), load_f32(tree + 48), load_f32(tree + 52));
	RayInvDir inv = ray_inv_direction(d);
	uint32_t stack[64];
	float dist[64];
	int top = 0;
	stack[0] = root_properties; dist[0] = -1.0f;
	bool hit = false;
	while (top >= 0 && io_fraction > 0.0f)
	{
		uint32_t node_properties = stack[top];
		float node_dist = dist[top];
		--top;
		if (!(node_dist < io_fraction))
			continue;
		uint32_t tri_count = node_properties >> 28;
		if (tri_count == 0)
		{
			const uint8_t *node = tree + ((size_t)node_properties << 2);
			uint32_t props[4]; float dd[4];
			int n = 0;
			for (int ch = 0; ch < 4; ++ch)
			{
				V3 mn = v3(half_to_float(load_u16(node + 0 + 2 * ch)), half_to_float(load_u16(node + 8 + 2 * ch)), half_to_float(load_u16(node + 16 + 2 * ch)));
				V3 mx = v3(half_to_float(load_u16(node + 24 + 2 * ch)), half_to_float(load_u16(node + 32 + 2 * ch)), half_to_float(load_u16(node + 40 + 2 * ch)));
				if (mn.x > mx.x || mn.y > mx.y || mn.z > mx.z) continue; // invalid (unused) child
				float t = ray_aabox(o, inv, mn, mx);
				if (t < io_fraction)
				{
					// insertion sort: farthest first (the closest child ends on top of the stack)
					int j = n++;
					while (j > 0 && dd[j - 1] < t) { dd[j] = dd[j - 1]; props[j] = props[j - 1]; --j; }
					dd[j] = t; props[j] = load_u32(node + 48 + 4 * ch);
				}
			}
			for (int j = 0; j < n && top < 62; ++j) { ++top; stack[top] = props[j]; dist[top] = dd[j]; }
		}
		else if (tri_count != 15)
		{
			uint32_t block_id = node_properties & 0x0fffffffu;
			const uint8_t *block_start = tree + ((size_t)block_id << 2);
			uint32_t header_flags = load_u32(block_start);
			const uint8_t *vertices = block_start + ((size_t)(header_flags & 0x1fffffffu) << 2);
			const uint8_t *blocks = block_start + 4;
			uint32_t block_sub = block_id_bits >= 32? block_id : ((0xffffffffu & ~((1u << block_id_bits) - 1u)) | block_id);
			for (uint32_t t = 0; t < tri_count; ++t)
			{
				const uint8_t *blk = blocks + 16 * (t >> 2);
				uint32_t lane = t & 3;
				V3 v[3];
				for (int vi = 0; vi < 3; ++vi)
				{
					uint32_t idx = blk[4 * vi + lane];
					uint32_t c1 = load_u32(vertices + 8 * idx), c2 = load_u32(vertices + 8 * idx + 4);
					uint32_t xc = c1 & 0x1fffffu, yc = (c1 >> 21) | ((c2 >> 21) << 11), zc = c2 & 0x1fffffu;
					v[vi] = v3((float)(int32_t)xc * tri_scale.x + tri_offset.x, (float)(int32_t)yc * tri_scale.y + tri_offset.y, (float)(int32_t)zc * tri_scale.z + tri_offset.z);
				}
				float f = ray_triangle(o, d, v[0], v[1], v[2]);
				if (f < io_fraction)
				{
					io_fraction = f;
					out_sub = (block_sub & ~(7u << block_id_bits)) | (t << block_id_bits);
					hit = true;
				}
			}
		}
	}
	return hit;
}

// Shape::CastRay in the centre of mass space of the shape: improves io_fraction / out_sub when the ray hits closer
B2J_D bool ray_shape(const DWorld &w, const ShapeDesc &decorated, V3 o, V3 d, float &io_fraction, uint32_t &out_sub);

// StaticCompoundShape::CastRay, closest hit: every sub shape in its own space (CastRayVisitor::VisitShape, CompoundShapeVisitors.h:
// Mat44::sInverseRotationTranslation(sub rotation, sub position)); the tree only prunes, the closest hit does not depend on the order
B2J_D bool ray_compound(const DWorld &w, const ShapeDesc &s, V3 o, V3 d, float &io_fraction, uint32_t &out_sub)
{
	bool hit = false;
	for (uint32_t i = 0; i < s.compound_num_subs; ++i)
	{
		const CompoundSub &sub = w.compound_subs[s.compound_sub_offset + i];
		Xf inv = xf_inverse_rotation_translation(to_q4(sub.rotation), to_v3(sub.position_com));
		V3 lo = mul(inv, o);
		V3 ld = mul(inv, o + d) - lo;
		uint32_t leaf_sub;
		if (ray_shape(w, w.shapes[sub.shape], lo, ld, io_fraction, leaf_sub))
		{
			hit = true;
			out_sub = s.compound_sub_bits == 0? 0xffffffffu : ((0xffffffffu & ~((1u << s.compound_sub_bits) - 1u)) | i);
		}
	}
	return hit;
}

B2J_D bool ray_shape(const DWorld &w, const ShapeDesc &decorated, V3 o, V3 d, float &io_fraction, uint32_t &out_sub)
{
	if (decorated.kind == B2J_SHAPE_COMPOUND)
		return ray_compound(w, decorated, o, d, io_fraction, out_sub);
	if (decorated.flags & SHAPE_LOCAL_ROTATION)
	{
		// RotatedTranslatedShape::CastRay: inRay.Transformed(Mat44::sRotation(mRotation.Conjugated())) (RotatedTranslatedShape.cpp:127-137)
		M33 inv = transposed(decorated.local_rot);
		V3 lo = mul(inv, o);
		d = mul(inv, o + d) - lo;
		o = lo;
	}
	const bool scaled = decorated.scale.x != 1.0f || decorated.scale.y != 1.0f || decorated.scale.z != 1.0f;
	if (scaled)
	{
		// ScaledShape::CastRay: the ray is scaled by 1 / scale and cast against the UNSCALED inner shape (ScaledShape.cpp:114-119)
		V3 inv_scale = v3(1.0f / decorated.scale.x, 1.0f / decorated.scale.y, 1.0f / decorated.scale.z);
		o = inv_scale * o;
		d = inv_scale * d;
	}
	const ShapeDesc &s = scaled? w.shapes[decorated.base_leaf] : decorated;
	float fraction = FLT_MAX;
	switch (s.kind)
	{
	case B2J_SHAPE_SPHERE: fraction = ray_sphere(o, d, v3_zero(), s.radius); break;
	case B2J_SHAPE_BOX: fraction = fmax_(ray_aabox(o, ray_inv_direction(d), -s.half_extent, s.half_extent), 0.0f); break;
	case B2J_SHAPE_CAPSULE: fraction = ray_capsule(o, d, s.half_height, s.radius); break;
	case B2J_SHAPE_CYLINDER: fraction = ray_cylinder_finite(o, d, s.half_height, s.radius); break;
	case B2J_SHAPE_CONVEX_HULL: { float f; if (ray_hull(w, s, o, d, f)) fraction = f; break; }
	case B2J_SHAPE_MESH: return ray_mesh(w, s, o, d, io_fraction, out_sub);
	default: break;
	}
	if (fraction < io_fraction)
	{
		io_fraction = fraction;
		out_sub = 0xffffffffu;
		return true;
	}
	return false;
}

// ---- NarrowPhaseQuery::CastRay, closest hit, one thread per ray ------------------------------------------------------------------
struct KCastRays
{
	DWorld w;
	Tree trees[8];
	const b2j_ray *rays; b2j_ray_hit *hits;
	const uint32_t *ray_world;   // batched worlds: world of every ray (null: a single world)
	uint32_t first_world, num_worlds; // batch group: the worlds this device world holds; rays of other worlds are left alone
	uint32_t object_layer;       // layer the ray collides as (ObjectLayerFilter / BroadPhaseLayerFilter from the world's tables), 0xffffffff = everything
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t world = 0;
		if (ray_world != nullptr)
		{
			world = ray_world[i];
			if (world < first_world || world >= first_world + num_worlds) return;
			world -= first_world;
		}
		b2j_ray ray = rays[i];
		V3 o = v3_load(ray.origin), d = v3_load(ray.direction);
		RayInvDir inv = ray_inv_direction(d);
		float best = 1.0f + FLT_EPSILON; // RayCastResult: no hit yet
		uint32_t best_body = B2J_INVALID_ID, best_sub = 0xffffffffu;
		for (uint32_t l = 0; l < w.num_bp_layers; ++l)
		{
			const Tree &t = trees[l];
			if (t.n == 0 || (object_layer != 0xffffffffu && !w.object_vs_bp[object_layer * w.num_bp_layers + l]))
				continue;
			int n = (int)t.n;
			int stack[64]; float dist[64];
			int top = 0;
			stack[0] = 0; dist[0] = -1.0f;
			if (t.world_root != nullptr)
			{
				int wr = t.world_root[world];
				if (wr < 0) continue;
				stack[0] = wr;
			}
			while (top >= 0 && best > 0.0f)
			{
				int node = stack[top]; float nd = dist[top];
				--top;
				if (!(nd < best))
					continue;
				if (node >= n - 1)
				{
					uint32_t b = t.leaf_body[node - (n - 1)];
					BodyInfo info = w.info[b];
					if (info.id == B2J_INVALID_ID || (object_layer != 0xffffffffu && !w.object_vs_object[object_layer * w.num_object_layers + info.object_layer]))
						continue;
					if (!(ray_aabox(o, inv, to_v3(w.bounds_min[b]), to_v3(w.bounds_max[b])) < best))
						continue;
					// TransformedShape::CastRay: the ray in the centre of mass space of the body (RayCast::Transformed)
					Xf inv_com = xf_inverse_rotation_translation(to_q4(w.rotation[b]), to_v3(w.position[b]));
					V3 lo = mul(inv_com, o);
					V3 ld = mul(inv_com, o + d) - lo;
					uint32_t sub;
					if (ray_shape(w, w.shapes[info.shape], lo, ld, best, sub)) { best_body = info.id; best_sub = sub; }
				}
				else
				{
					int l2 = t.child_left[node], r2 = t.child_right[node];
					float tl = ray_aabox(o, inv, to_v3(t.node_min[l2]), to_v3(t.node_max[l2]));
					float tr = ray_aabox(o, inv, to_v3(t.node_min[r2]), to_v3(t.node_max[r2]));
					// farther child first: the closer one is popped first
					bool left_first = tl <= tr;
					int far_node = left_first? r2 : l2, near_node = left_first? l2 : r2;
					float far_t = left_first? tr : tl, near_t = left_first? tl : tr;
					if (far_t < best && top < 62) { ++top; stack[top] = far_node; dist[top] = far_t; }
					if (near_t < best && top < 62) { ++top; stack[top] = near_node; dist[top] = near_t; }
				}
			}
		}
		b2j_ray_hit h;
		h.body = best <= 1.0f? best_body : B2J_INVALID_ID;
		h.sub_shape = best <= 1.0f? best_sub : 0xffffffffu;
		h.fraction = best <= 1.0f? best : 1.0f + FLT_EPSILON;
		hits[i] = h;
	}
};

// ---- BroadPhaseQuery::CollideAABox: bodies whose world space bounds overlap the box, one thread per box ---------------------------
struct KCollideAABox
{
	DWorld w;
	Tree trees[8];
	int mode;                    // 0: boxes, 1: spheres ([n][4] centre, radius: CollideSphere), 2: points ([n][3]: CollidePoint)
	const float *boxes;          // [n][6] min xyz, max xyz
	uint32_t *counts;            // [n] number of overlapping bodies (can exceed max_hits)
	uint32_t *ids;               // [n][max_hits] body ids (the first max_hits found)
	uint32_t max_hits;
	const uint32_t *box_world; uint32_t first_world, num_worlds;
	uint32_t object_layer;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t world = 0;
		if (box_world != nullptr)
		{
			world = box_world[i];
			if (world < first_world || world >= first_world + num_worlds) return;
			world -= first_world;
		}
		V3 mn, mx, centre = v3_zero();
		float radius_sq = 0.0f;
		if (mode == 0) { mn = v3_load(boxes + 6 * i); mx = v3_load(boxes + 6 * i + 3); }
		else if (mode == 1) { centre = v3_load(boxes + 4 * i); float r = boxes[4 * i + 3]; radius_sq = r * r; mn = centre - v3_rep(r); mx = centre + v3_rep(r); }
		else { centre = v3_load(boxes + 3 * i); mn = centre; mx = centre; }
		// AABox4VsSphere: the closest point of the box to the centre is within the radius (a point: radius 0 = AABox4VsPoint)
		auto overlaps = [&](V3 bmin, V3 bmax) {
			if (mode == 0) return aabb_overlaps(mn, mx, bmin, bmax);
			V3 closest = v3_min(v3_max(centre, bmin), bmax);
			return length_sq(closest - centre) <= radius_sq;
		};
		uint32_t count = 0;
		for (uint32_t l = 0; l < w.num_bp_layers; ++l)
		{
			const Tree &t = trees[l];
			if (t.n == 0 || (object_layer != 0xffffffffu && !w.object_vs_bp[object_layer * w.num_bp_layers + l]))
				continue;
			int n = (int)t.n;
			int stack[128];
			int top = 0;
			stack[0] = 0;
			if (t.world_root != nullptr)
			{
				int wr = t.world_root[world];
				if (wr < 0) continue;
				stack[0] = wr;
			}
			while (top >= 0)
			{
				int node = stack[top--];
				if (node >= n - 1)
				{
					uint32_t b = t.leaf_body[node - (n - 1)];
					BodyInfo info = w.info[b];
					if (info.id == B2J_INVALID_ID || (object_layer != 0xffffffffu && !w.object_vs_object[object_layer * w.num_object_layers + info.object_layer]))
						continue;
					if (overlaps(to_v3(w.bounds_min[b]), to_v3(w.bounds_max[b])))
					{
						if (count < max_hits) ids[(size_t)i * max_hits + count] = info.id;
						++count;
					}
				}
				else
				{
					int l2 = t.child_left[node], r2 = t.child_right[node];
					if (overlaps(to_v3(t.node_min[l2]), to_v3(t.node_max[l2])) && top < 126) stack[++top] = l2;
					if (overlaps(to_v3(t.node_min[r2]), to_v3(t.node_max[r2])) && top < 126) stack[++top] = r2;
				}
			}
		}
		counts[i] = count;
	}
};

// ---- NarrowPhaseQuery::CollideShape, all hits: one warp per query (lane 0 working, EPA hull in the warp's shared memory, as the
// compound pairs of the step) ---------------------------------------------------------------------------------------------------------
struct KCollideShape
{
	DWorld w;
	Tree trees[8];
	const b2j_shape_query *queries;
	b2j_collide_shape_hit *hits; uint32_t *counts; uint32_t max_hits;
	float max_separation_distance;
	uint32_t object_layer;
	MeshScratch *mesh_scratch;   // never written by a query (the collector takes the hits); the pair functions want a reference
	B2J_D void run(uint32_t i, bool valid, uint32_t slot, EpaStorageFull &epa_storage) const
	{
		(void)valid; (void)slot;
		EpaScratch epa = epa_storage.view();
		b2j_shape_query q = queries[i];
		const ShapeDesc &s1 = w.shapes[q.shape];
		V3 com = v3_load(q.position), base = v3_load(q.base_offset);
		Q4 rot = q4(q.rotation[0], q.rotation[1], q.rotation[2], q.rotation[3]);
		// bounds of the query shape, expanded by the max separation distance (NarrowPhaseQuery.cpp:287-288)
		V3 mn, mx;
		world_bounds(w, s1, com, rot, mn, mx);
		mn = mn - v3_rep(max_separation_distance); mx = mx + v3_rep(max_separation_distance);
		QueryCollector qc; qc.hits = hits + (size_t)i * max_hits; qc.max_hits = max_hits; qc.count = 0; qc.body = B2J_INVALID_ID;
		CompoundPairCtx p;
		p.max_separation_distance = max_separation_distance;
		p.movement_direction = v3_zero();
		p.epa = &epa; p.ms = mesh_scratch; p.num_manifolds = 0; p.query = &qc;
		// TransformedShape::CollideShape: both centre of mass transforms relative to the base offset
		CompoundSide a;
		a.shape = &s1; a.transform = xf(m33_rotation(rot), com + (-base)); a.sub = 0xffffffffu;
		for (uint32_t l = 0; l < w.num_bp_layers; ++l)
		{
			const Tree &t = trees[l];
			if (t.n == 0 || (object_layer != 0xffffffffu && !w.object_vs_bp[object_layer * w.num_bp_layers + l]))
				continue;
			int n = (int)t.n;
			int stack[128];
			int top = 0;
			stack[0] = 0;
			while (top >= 0)
			{
				int node = stack[top--];
				if (node >= n - 1)
				{
					uint32_t b = t.leaf_body[node - (n - 1)];
					BodyInfo info = w.info[b];
					if (info.id == B2J_INVALID_ID || (object_layer != 0xffffffffu && !w.object_vs_object[object_layer * w.num_object_layers + info.object_layer]))
						continue;
					if (!aabb_overlaps(mn, mx, to_v3(w.bounds_min[b]), to_v3(w.bounds_max[b])))
						continue;
					qc.body = info.id;
					CompoundSide side;
					side.shape = &w.shapes[info.shape];
					side.transform = xf(m33_rotation(to_q4(w.rotation[b])), to_v3(w.position[b]) + (-base));
					side.sub = 0xffffffffu;
					compound_dispatch(w, p, a, side);
				}
				else
				{
					int l2 = t.child_left[node], r2 = t.child_right[node];
					if (aabb_overlaps(mn, mx, to_v3(t.node_min[l2]), to_v3(t.node_max[l2])) && top < 126) stack[++top] = l2;
					if (aabb_overlaps(mn, mx, to_v3(t.node_min[r2]), to_v3(t.node_max[r2])) && top < 126) stack[++top] = r2;
				}
			}
		}
		counts[i] = qc.count;
	}
};

} // namespace b2j
