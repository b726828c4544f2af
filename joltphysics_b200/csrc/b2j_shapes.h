// b2j_shapes.h -- support functions, supporting faces and bounds of the convex shapes on the path.
//
// Restates (for unit scale, which is all the step ever passes: PhysicsSystem.cpp:1045 Vec3::sOne()):
//   SphereShape.cpp:67-139   BoxShape.cpp:78-146 + AABox.h:193-300   CapsuleShape.cpp:91-216,259-277
//   ConvexHullShape.cpp:395-449,531-590,674-719   ConvexSupport.h (TransformedConvexObject, AddConvexRadius, TriangleConvexSupport)
#pragma once

#include "b2j_world.h"

namespace b2j {

enum { SUPPORT_EXCLUDE_CONVEX_RADIUS = 0, SUPPORT_INCLUDE_CONVEX_RADIUS = 1 };

// A convex shape in its local space in one support mode (what ConvexShape::GetSupportFunction returns)
struct ConvexSupport
{
	uint32_t kind;
	float radius;            // sphere / capsule radius added by the support function itself (include mode) else 0
	float convex_radius;     // Support::GetConvexRadius()
	V3 ext;                  // box: (reduced) half extent, capsule: (0, half height, 0), cylinder: (radius, half height, 0) of this mode
	const F4 *points;        // hull points for this mode
	uint32_t num_points;

	B2J_HD V3 support(V3 dir) const
	{
		switch (kind)
		{
		case B2J_SHAPE_SPHERE:
			{
				if (radius == 0.0f) return v3_zero(); // SphereNoConvex
				float len = length(dir);
				return len > 0.0f? (radius / len) * dir : v3_zero();
			}
		case B2J_SHAPE_BOX: // AABox::GetSupport: select(max, min, dir < 0)
			return v3(dir.x < 0.0f? 0.0f - ext.x : ext.x, dir.y < 0.0f? 0.0f - ext.y : ext.y, dir.z < 0.0f? 0.0f - ext.z : ext.z);
		case B2J_SHAPE_CAPSULE:
			{
				if (radius == 0.0f) // CapsuleNoConvex
					return dir.y > 0.0f? ext : -ext;
				float len = length(dir);
				V3 r = len > 0.0f? dir * (radius / len) : v3_zero();
				return dir.y > 0.0f? r + ext : r - ext;
			}
		case B2J_SHAPE_CYLINDER: // CylinderShape::Cylinder::GetSupport (CylinderShape.cpp:143-152)
			{
				float o = sqrt_(square(dir.x) + square(dir.z));
				float y = (dir.y < 0.0f? -1.0f : 1.0f) * ext.y; // Sign(y) * mHalfHeight
				if (o > 0.0f)
					return v3((ext.x * dir.x) / o, y, (ext.x * dir.z) / o);
				return v3(0.0f, y, 0.0f);
			}
		default: // hull: first vertex with strictly greatest dot (ConvexHullShape.cpp:412-421)
			{
				float best_dot = -FLT_MAX;
				V3 best_point = v3_zero();
				for (uint32_t i = 0; i < num_points; ++i)
				{
					V3 p = to_v3(points[i]);
					float d = dot(p, dir);
					if (d > best_dot) { best_dot = d; best_point = p; }
				}
				return best_point;
			}
		}
	}
};

B2J_HD ConvexSupport make_support(const DWorld &w, const ShapeDesc &s, int mode)
{
	ConvexSupport c;
	c.kind = s.kind;
	c.radius = 0.0f;
	c.convex_radius = 0.0f;
	c.ext = v3_zero();
	c.points = nullptr;
	c.num_points = 0;
	switch (s.kind)
	{
	case B2J_SHAPE_SPHERE:
		if (mode == SUPPORT_INCLUDE_CONVEX_RADIUS) c.radius = s.radius; else c.convex_radius = s.radius;
		break;
	case B2J_SHAPE_BOX:
		if (mode == SUPPORT_INCLUDE_CONVEX_RADIUS) c.ext = s.half_extent;
		else { c.ext = s.half_extent - v3_rep(s.convex_radius); c.convex_radius = s.convex_radius; }
		break;
	case B2J_SHAPE_CAPSULE:
		c.ext = v3(0.0f, s.half_height, 0.0f);
		if (mode == SUPPORT_INCLUDE_CONVEX_RADIUS) c.radius = s.radius; else c.convex_radius = s.radius;
		break;
	case B2J_SHAPE_CYLINDER: // CylinderShape::GetSupportFunction (CylinderShape.cpp:171-196)
		if (mode == SUPPORT_INCLUDE_CONVEX_RADIUS) c.ext = v3(s.radius, s.half_height, 0.0f);
		else { c.ext = v3(s.radius - s.convex_radius, s.half_height - s.convex_radius, 0.0f); c.convex_radius = s.convex_radius; }
		break;
	default:
		c.num_points = s.hull_num_points;
		if (mode == SUPPORT_INCLUDE_CONVEX_RADIUS || s.convex_radius == 0.0f) c.points = w.hull_points + s.hull_point_offset;
		else { c.points = w.hull_shrunk + s.hull_point_offset; c.convex_radius = s.convex_radius; }
		break;
	}
	return c;
}

// Centre of mass transform of the convex leaf of a (possibly decorated) shape: RotatedTranslatedShape hands
// inCenterOfMassTransform * Mat44::sRotation(mRotation) down to its inner shape (RotatedTranslatedShape.cpp:73-77,183-192)
B2J_HD Xf shape_transform(const ShapeDesc &s, const Xf &body)
{
	return (s.flags & SHAPE_LOCAL_ROTATION)? mul(body, xf(s.local_rot, v3_zero())) : body;
}

// TransformedConvexObject (ConvexSupport.h)
struct TransformedSupport
{
	Xf xform;
	M33 rot_t;   // transposed rotation (Multiply3x3Transposed = Transposed3x3().Multiply3x3)
	ConvexSupport s;
	B2J_HD V3 support(V3 dir) const { return mul(xform, s.support(mul(rot_t, dir))); }
};
B2J_HD TransformedSupport make_transformed(const Xf &x, const ConvexSupport &s) { TransformedSupport t; t.xform = x; t.rot_t = transposed(x.r); t.s = s; return t; }

// AddConvexRadius (ConvexSupport.h)
struct AddRadiusSupport
{
	ConvexSupport s;
	float radius;
	B2J_HD V3 support(V3 dir) const
	{
		float len = length(dir);
		return len > 0.0f? s.support(dir) + (radius / len) * dir : s.support(dir);
	}
};

// TriangleConvexSupport (ConvexSupport.h)
struct TriangleSupport
{
	V3 v1, v2, v3_;
	B2J_HD V3 support(V3 dir) const
	{
		float d1 = dot(v1, dir), d2 = dot(v2, dir), d3 = dot(v3_, dir);
		if (d1 > d2) return d1 > d3? v1 : v3_;
		return d2 > d3? v2 : v3_;
	}
};

// Shape::GetSupportingFace for unit scale; vertices are transformed by xform. Returns the vertex count (<= 32).
B2J_HD int supporting_face(const DWorld &w, const ShapeDesc &s, V3 dir, const Xf &xform, V3 *out)
{
	switch (s.kind)
	{
	case B2J_SHAPE_SPHERE:
		return 0;
	case B2J_SHAPE_BOX:
		{
			V3 mn = -s.half_extent, mx = s.half_extent;
			int axis = highest_component_index(v3_abs(dir));
			if (v3_get(dir, axis) < 0.0f)
			{
				switch (axis)
				{
				case 0: out[0] = v3(mx.x, mn.y, mn.z); out[1] = v3(mx.x, mx.y, mn.z); out[2] = v3(mx.x, mx.y, mx.z); out[3] = v3(mx.x, mn.y, mx.z); break;
				case 1: out[0] = v3(mn.x, mx.y, mn.z); out[1] = v3(mn.x, mx.y, mx.z); out[2] = v3(mx.x, mx.y, mx.z); out[3] = v3(mx.x, mx.y, mn.z); break;
				default: out[0] = v3(mn.x, mn.y, mx.z); out[1] = v3(mx.x, mn.y, mx.z); out[2] = v3(mx.x, mx.y, mx.z); out[3] = v3(mn.x, mx.y, mx.z); break;
				}
			}
			else
			{
				switch (axis)
				{
				case 0: out[0] = v3(mn.x, mn.y, mn.z); out[1] = v3(mn.x, mn.y, mx.z); out[2] = v3(mn.x, mx.y, mx.z); out[3] = v3(mn.x, mx.y, mn.z); break;
				case 1: out[0] = v3(mn.x, mn.y, mn.z); out[1] = v3(mx.x, mn.y, mn.z); out[2] = v3(mx.x, mn.y, mx.z); out[3] = v3(mn.x, mn.y, mx.z); break;
				default: out[0] = v3(mn.x, mn.y, mn.z); out[1] = v3(mn.x, mx.y, mn.z); out[2] = v3(mx.x, mx.y, mn.z); out[3] = v3(mx.x, mn.y, mn.z); break;
				}
			}
			for (int i = 0; i < 4; ++i) out[i] = mul(xform, out[i]);
			return 4;
		}
	case B2J_SHAPE_CAPSULE:
		{
			V3 direction = v3(dir.x, 0.0f, dir.z);
			float len = length(direction);
			if (len == 0.0f)
				return 0;
			V3 hh = v3(0.0f, s.half_height, 0.0f);
			V3 support = (s.radius / len) * direction;
			V3 support_top = hh - support;
			V3 support_bottom = -hh - support;
			float proj_top = dot(support_top, dir), proj_bottom = dot(support_bottom, dir);
			if (fabs_(proj_top - proj_bottom) < 0.02f * length(dir)) // cCapsuleProjectionSlop (PhysicsSettings.h:19)
			{
				out[0] = mul(xform, support_top);
				out[1] = mul(xform, support_bottom);
				return 2;
			}
			return 0;
		}
	case B2J_SHAPE_CYLINDER: // CylinderShape::GetSupportingFace (CylinderShape.cpp:198-244)
		{
			float x = dir.x, y = dir.y, z = dir.z;
			float xz_sq = square(x) + square(z);
			float y_sq = square(y);
			if (xz_sq > y_sq)
			{
				// an edge of the side
				float f = (0.0f - s.radius) / sqrt_(xz_sq);
				float vx = x * f, vz = z * f;
				out[0] = mul(xform, v3(vx, s.half_height, vz));
				out[1] = mul(xform, v3(vx, 0.0f - s.half_height, vz));
				return 2;
			}
			// top or bottom: the 8 vertex approximation of the cap, rotated so that one vertex points along the direction
			Xf transform = xform;
			if (xz_sq > 0.00765427f * y_sq)
			{
				float inv = sqrt_(xz_sq);
				V3 base_x = v3(x / inv, 0.0f / inv, z / inv);
				V3 base_z = v3(base_x.z * -1.0f, base_x.y * 0.0f, base_x.x * 1.0f);
				transform = mul(transform, xf(m33(base_x, v3(0.0f, 1.0f, 0.0f), base_z), v3_zero()));
			}
			V3 multiplier = y < 0.0f? v3(s.radius, s.half_height, s.radius) : v3(0.0f - s.radius, 0.0f - s.half_height, s.radius);
			transform = xf(m33(multiplier.x * transform.r.c0, multiplier.y * transform.r.c1, multiplier.z * transform.r.c2), transform.t); // PreScaled
			const float h = 0.707106769f;
			const V3 top_face[8] = { v3(0.0f, 1.0f, 1.0f), v3(h, 1.0f, h), v3(1.0f, 1.0f, 0.0f), v3(h, 1.0f, -h), v3(-0.0f, 1.0f, -1.0f), v3(-h, 1.0f, -h), v3(-1.0f, 1.0f, 0.0f), v3(-h, 1.0f, h) };
			for (int i = 0; i < 8; ++i) out[i] = mul(transform, top_face[i]);
			return 8;
		}
	default:
		{
			const F4 *planes = w.hull_planes + s.hull_face_offset;
			V3 n0 = to_v3(planes[0]);
			float best_dot = dot(n0, dir) / length(n0);
			int best_face = 0;
			for (uint32_t i = 1; i < s.hull_num_faces; ++i)
			{
				V3 n = to_v3(planes[i]);
				float d = dot(n, dir) / length(n);
				if (d < best_dot) { best_dot = d; best_face = (int)i; }
			}
			uint32_t face = w.hull_faces[s.hull_face_offset + best_face];
			int first = (int)(face & 0xffff), num = (int)(face >> 16);
			const int max_vertices_to_return = 16; // SupportingFace capacity 32 / 2
			int delta = (num + max_vertices_to_return) / max_vertices_to_return;
			int n = 0;
			if (s.flags & SHAPE_SCALED_HULL)
			{
				// (the planes in the pool are already inv_scale * normal) transform = inCenterOfMassTransform.PreScaled(inScale), applied to the
				// unscaled points (ConvexHullShape.cpp:702-719; positive scales only: no winding flip)
				Xf scaled = xf(m33(s.scale.x * xform.r.c0, s.scale.y * xform.r.c1, s.scale.z * xform.r.c2), xform.t);
				for (int v = first; v < first + num; v += delta)
					out[n++] = mul(scaled, to_v3(w.hull_points[s.hull_orig_offset + w.hull_vtx[s.hull_vtx_offset + v]]));
				return n;
			}
			for (int v = first; v < first + num; v += delta)
				out[n++] = mul(xform, to_v3(w.hull_points[s.hull_point_offset + w.hull_vtx[s.hull_vtx_offset + v]]));
			return n;
		}
	}
}

// SubShape::GetLocalTransformNoScale(one): Mat44::sRotationTranslation(rotation, position relative to the compound's centre of mass)
B2J_HD Xf compound_sub_transform(const CompoundSub &sub) { return xf(m33_rotation(to_q4(sub.rotation)), to_v3(sub.position_com)); }

// Shape::GetWorldSpaceBounds(inCenterOfMassTransform, one) of a convex shape or mesh; x_in = the transform before the shape's own decorators
B2J_HD void world_bounds_leaf(const ShapeDesc &s, const Xf &x_in, V3 &out_min, V3 &out_max)
{
	Xf x = shape_transform(s, x_in);
	switch (s.kind)
	{
	case B2J_SHAPE_SPHERE: // SphereShape.cpp:67-74
		{
			V3 he = v3_rep(s.radius);
			out_min = -he + x.t;
			out_max = he + x.t;
		}
		break;
	case B2J_SHAPE_CAPSULE: // CapsuleShape.cpp:266-277
		{
			V3 extent = v3_rep(s.radius);
			V3 height = v3(0.0f, s.half_height, 0.0f);
			V3 p1 = mul(x, -height), p2 = mul(x, height);
			out_min = v3_min(p1, p2) - extent;
			out_max = v3_max(p1, p2) + extent;
		}
		break;
	default: // AABox::Transformed (AABox.h:193-213)
		{
			V3 new_min = x.t, new_max = x.t;
			for (int c = 0; c < 3; ++c)
			{
				V3 col = m33_col(x.r, c);
				V3 a = col * v3_get(s.local_min, c);
				V3 b = col * v3_get(s.local_max, c);
				new_min += v3_min(a, b);
				new_max += v3_max(a, b);
			}
			out_min = new_min;
			out_max = new_max;
		}
		break;
	}
}

// Body::CalculateWorldSpaceBoundsInternal -> Shape::GetWorldSpaceBounds(com transform, one)
B2J_HD void world_bounds(const DWorld &w, const ShapeDesc &s, V3 pos, Q4 rot, V3 &out_min, V3 &out_max)
{
	Xf x = xf_rotation_translation(rot, pos);
	if (s.kind == B2J_SHAPE_COMPOUND && s.compound_num_subs <= 10)
	{
		// CompoundShape::GetWorldSpaceBounds (CompoundShape.cpp:93-115): up to 10 sub shapes are bounded one by one
		V3 mn = v3_rep(FLT_MAX), mx = v3_rep(-FLT_MAX);
		for (uint32_t i = 0; i < s.compound_num_subs; ++i)
		{
			const CompoundSub &sub = w.compound_subs[s.compound_sub_offset + i];
			V3 a, b;
			world_bounds_leaf(w.shapes[sub.shape], mul(x, compound_sub_transform(sub)), a, b);
			mn = v3_min(mn, a); mx = v3_max(mx, b);
		}
		out_min = mn; out_max = mx;
		return;
	}
	world_bounds_leaf(s, x, out_min, out_max);
}

// Body::GetSleepTestPoints (Body.inl:156-188)
B2J_HD void sleep_test_points(const ShapeDesc &s, V3 pos, Q4 rot, V3 *out)
{
	out[0] = pos;
	V3 extent = 0.5f * (s.outer_max - s.outer_min); // (the bounds of the body's own shape, decorators included, in the body's frame)
	int lowest = lowest_component_index(extent);
	M33 r = m33_rotation(rot);
	switch (lowest)
	{
	case 0: out[1] = pos + extent.y * r.c1; out[2] = pos + extent.z * r.c2; break;
	case 1: out[1] = pos + extent.x * r.c0; out[2] = pos + extent.z * r.c2; break;
	default: out[1] = pos + extent.x * r.c0; out[2] = pos + extent.y * r.c1; break;
	}
}

// OrientedBox(transform, box).Overlaps(aabox) (OrientedBox.cpp:12-95, OrientedBox.h:26)
B2J_HD bool obb_vs_aabb(const Xf &orientation_in, V3 box2_min, V3 box2_max, V3 a_min, V3 a_max, float in_epsilon = 1.0e-6f)
{
	// OrientedBox(inOrientation.PreTranslated(inBox.GetCenter()), inBox.GetExtent())
	V3 c2 = 0.5f * (box2_min + box2_max);
	V3 he = 0.5f * (box2_max - box2_min);
	Xf orientation = xf(orientation_in.r, orientation_in.t + mul(orientation_in.r, c2));

	V3 a_center = 0.5f * (a_min + a_max);
	V3 a_he = 0.5f * (a_max - a_min);
	V3 t = orientation.t - a_center;
	const M33 &r = orientation.r;
	V3 eps = v3_rep(in_epsilon);
	V3 abs_r[3] = { v3_abs(r.c0) + eps, v3_abs(r.c1) + eps, v3_abs(r.c2) + eps };
	float hea[3] = { he.x, he.y, he.z };
	float ahe[3] = { a_he.x, a_he.y, a_he.z };
	float tt[3] = { t.x, t.y, t.z };
	// rot(row, col): col 0..2 = rotation columns, col 3 = translation
	#define B2J_ROT(row, col) ((col) == 0? v3_get(r.c0, row) : ((col) == 1? v3_get(r.c1, row) : v3_get(r.c2, row)))
	float ra, rb;
	for (int i = 0; i < 3; i++)
	{
		ra = ahe[i];
		rb = hea[0] * v3_get(abs_r[0], i) + hea[1] * v3_get(abs_r[1], i) + hea[2] * v3_get(abs_r[2], i);
		if (fabs_(tt[i]) > ra + rb) return false;
	}
	for (int i = 0; i < 3; i++)
	{
		ra = dot(a_he, abs_r[i]);
		rb = hea[i];
		if (fabs_(dot(t, m33_col(r, i))) > ra + rb) return false;
	}
	ra = ahe[1] * abs_r[0].z + ahe[2] * abs_r[0].y; rb = hea[1] * abs_r[2].x + hea[2] * abs_r[1].x;
	if (fabs_(tt[2] * B2J_ROT(1, 0) - tt[1] * B2J_ROT(2, 0)) > ra + rb) return false;
	ra = ahe[1] * abs_r[1].z + ahe[2] * abs_r[1].y; rb = hea[0] * abs_r[2].x + hea[2] * abs_r[0].x;
	if (fabs_(tt[2] * B2J_ROT(1, 1) - tt[1] * B2J_ROT(2, 1)) > ra + rb) return false;
	ra = ahe[1] * abs_r[2].z + ahe[2] * abs_r[2].y; rb = hea[0] * abs_r[1].x + hea[1] * abs_r[0].x;
	if (fabs_(tt[2] * B2J_ROT(1, 2) - tt[1] * B2J_ROT(2, 2)) > ra + rb) return false;
	ra = ahe[0] * abs_r[0].z + ahe[2] * abs_r[0].x; rb = hea[1] * abs_r[2].y + hea[2] * abs_r[1].y;
	if (fabs_(tt[0] * B2J_ROT(2, 0) - tt[2] * B2J_ROT(0, 0)) > ra + rb) return false;
	ra = ahe[0] * abs_r[1].z + ahe[2] * abs_r[1].x; rb = hea[0] * abs_r[2].y + hea[2] * abs_r[0].y;
	if (fabs_(tt[0] * B2J_ROT(2, 1) - tt[2] * B2J_ROT(0, 1)) > ra + rb) return false;
	ra = ahe[0] * abs_r[2].z + ahe[2] * abs_r[2].x; rb = hea[0] * abs_r[1].y + hea[1] * abs_r[0].y;
	if (fabs_(tt[0] * B2J_ROT(2, 2) - tt[2] * B2J_ROT(0, 2)) > ra + rb) return false;
	ra = ahe[0] * abs_r[0].y + ahe[1] * abs_r[0].x; rb = hea[1] * abs_r[2].z + hea[2] * abs_r[1].z;
	if (fabs_(tt[1] * B2J_ROT(0, 0) - tt[0] * B2J_ROT(1, 0)) > ra + rb) return false;
	ra = ahe[0] * abs_r[1].y + ahe[1] * abs_r[1].x; rb = hea[0] * abs_r[2].z + hea[2] * abs_r[0].z;
	if (fabs_(tt[1] * B2J_ROT(0, 1) - tt[0] * B2J_ROT(1, 1)) > ra + rb) return false;
	ra = ahe[0] * abs_r[2].y + ahe[1] * abs_r[2].x; rb = hea[0] * abs_r[1].z + hea[1] * abs_r[0].z;
	if (fabs_(tt[1] * B2J_ROT(0, 2) - tt[0] * B2J_ROT(1, 2)) > ra + rb) return false;
	#undef B2J_ROT
	return true;
}

} // namespace b2j
