// b2j_manifold.h -- contact manifold from two supporting faces: polygon clipping, plane projection, pruning to 4 points.
//
// Restates Jolt/Physics/Collision/ManifoldBetweenTwoFaces.cpp:16-269 (PruneContactPoints, ManifoldBetweenTwoFaces) and
// Jolt/Geometry/ClipPoly.h (ClipPolyVsPlane / VsPoly / VsEdge) with identical operation order and tie breaks.
#pragma once

#include "b2j_math.h"

namespace b2j {

enum { MAX_FACE_VERTS = 32, MAX_CLIP_VERTS = 64, MAX_MANIFOLD_POINTS = 64 };

// ClipPolyVsPlane; returns the output vertex count
B2J_HD int clip_poly_vs_plane(const V3 *poly, int n, V3 plane_origin, V3 plane_normal, V3 *out)
{
	int m = 0;
	V3 e1 = poly[n - 1];
	float prev_num = dot(plane_origin - e1, plane_normal);
	bool prev_inside = prev_num < 0.0f;
	for (int j = 0; j < n; ++j)
	{
		V3 e2 = poly[j];
		float num = dot(plane_origin - e2, plane_normal);
		bool cur_inside = num < 0.0f;
		if (cur_inside != prev_inside)
		{
			V3 e12 = e2 - e1;
			float denom = dot(e12, plane_normal);
			if (denom != 0.0f)
			{
				if (m < MAX_CLIP_VERTS) out[m++] = e1 + (prev_num / denom) * e12;
			}
			else
				cur_inside = prev_inside;
		}
		if (cur_inside)
		{
			if (m < MAX_CLIP_VERTS) out[m++] = e2;
		}
		prev_num = num;
		prev_inside = cur_inside;
		e1 = e2;
	}
	return m;
}

// ClipPolyVsPoly; tmp0/tmp1/out have MAX_CLIP_VERTS entries. Returns the output vertex count.
B2J_HD int clip_poly_vs_poly(const V3 *poly, int n, const V3 *clip, int nc, V3 clip_normal_in, V3 *tmp0, V3 *tmp1, V3 *out)
{
	V3 *tmp[2] = { tmp0, tmp1 };
	int tmp_n[2] = { 0, 0 };
	int tmp_idx = 0;
	int out_n = 0;
	for (int i = 0; i < nc; ++i)
	{
		V3 clip_e1 = clip[i];
		V3 clip_e2 = clip[(i + 1) % nc];
		V3 clip_normal = cross(clip_normal_in, clip_e2 - clip_e1);
		const V3 *src = i == 0? poly : tmp[tmp_idx];
		int src_n = i == 0? n : tmp_n[tmp_idx];
		tmp_idx ^= 1;
		bool last = i == nc - 1;
		V3 *tgt = last? out : tmp[tmp_idx];
		int tgt_n = clip_poly_vs_plane(src, src_n, clip_e1, clip_normal, tgt);
		if (last) out_n = tgt_n; else tmp_n[tmp_idx] = tgt_n;
		if (tgt_n < 3)
			return 0;
	}
	return out_n;
}

// ClipPolyVsEdge
B2J_HD int clip_poly_vs_edge(const V3 *poly, int n, V3 edge_v1, V3 edge_v2, V3 clipping_edge_normal, V3 *out)
{
	int m = 0;
	V3 edge = edge_v2 - edge_v1;
	V3 edge_normal = cross(clipping_edge_normal, edge);
	V3 polygon_normal = cross(poly[2] - poly[0], poly[1] - poly[0]);
	float polygon_normal_len_sq = length_sq(polygon_normal);
	V3 v1 = edge_v1 + (dot(polygon_normal, poly[0] - edge_v1) * polygon_normal) / polygon_normal_len_sq;
	V3 v2 = edge_v2 + (dot(polygon_normal, poly[0] - edge_v2) * polygon_normal) / polygon_normal_len_sq;
	V3 v12 = v2 - v1;
	float v12_len_sq = length_sq(v12);
	V3 e1 = poly[n - 1];
	float prev_num = dot(edge_v1 - e1, edge_normal);
	bool prev_inside = prev_num < 0.0f;
	for (int j = 0; j < n; ++j)
	{
		V3 e2 = poly[j];
		float num = dot(edge_v1 - e2, edge_normal);
		bool cur_inside = num < 0.0f;
		if (cur_inside != prev_inside)
		{
			V3 e12 = e2 - e1;
			float denom = dot(e12, edge_normal);
			V3 clipped_point = denom != 0.0f? e1 + (prev_num / denom) * e12 : e1;
			float projection = dot(clipped_point - v1, v12);
			V3 p = projection < 0.0f? v1 : (projection > v12_len_sq? v2 : clipped_point);
			if (m < MAX_CLIP_VERTS) out[m++] = p;
		}
		prev_num = num;
		prev_inside = cur_inside;
		e1 = e2;
	}
	return m;
}

// ManifoldBetweenTwoFaces: appends to points1/points2 (count in/out through num). Scratch: 3 * MAX_CLIP_VERTS vectors.
B2J_HD void manifold_between_two_faces(V3 contact_point1, V3 contact_point2, V3 penetration_axis_in, float max_contact_distance,
	const V3 *face1, int n1, const V3 *face2, int n2, V3 *points1, V3 *points2, int &num, V3 *scratch)
{
	int old_size = num;
	int mn = n1 < n2? n1 : n2, mx = n1 < n2? n2 : n1;
	if (mn >= 2 && mx >= 3)
	{
		const V3 *s1, *s2;
		int s1n, s2n;
		V3 *cp1, *cp2;
		V3 penetration_axis;
		if (n2 >= 3)
		{
			s1 = face1; s1n = n1; s2 = face2; s2n = n2; cp1 = points1; cp2 = points2; penetration_axis = penetration_axis_in;
		}
		else
		{
			s1 = face2; s1n = n2; s2 = face1; s2n = n1; cp1 = points2; cp2 = points1; penetration_axis = -penetration_axis_in;
		}
		V3 plane_origin = s1[0];
		V3 first_edge = s1[1] - plane_origin;
		V3 plane_normal;
		V3 *clipped = scratch;
		int nclipped;
		if (s1n >= 3)
		{
			nclipped = clip_poly_vs_poly(s2, s2n, s1, s1n, penetration_axis, scratch + MAX_CLIP_VERTS, scratch + 2 * MAX_CLIP_VERTS, clipped);
			plane_normal = cross(first_edge, s1[2] - plane_origin);
		}
		else
		{
			nclipped = clip_poly_vs_edge(s2, s2n, s1[0], s1[1], penetration_axis, clipped);
			plane_normal = cross(cross(first_edge, penetration_axis), first_edge);
		}
		float penetration_axis_dot_plane_normal = dot(penetration_axis, plane_normal);
		if (penetration_axis_dot_plane_normal != 0.0f)
		{
			float penetration_axis_len = length(penetration_axis);
			for (int i = 0; i < nclipped; ++i)
			{
				V3 p2 = clipped[i];
				float distance = dot(p2 - plane_origin, plane_normal) / penetration_axis_dot_plane_normal;
				if (distance * penetration_axis_len < max_contact_distance)
				{
					V3 p1 = p2 - distance * penetration_axis;
					if (num < MAX_MANIFOLD_POINTS)
					{
						cp1[num] = p1;
						cp2[num] = p2;
						++num;
					}
				}
			}
		}
	}
	if (num == old_size && num < MAX_MANIFOLD_POINTS)
	{
		points1[num] = contact_point1;
		points2[num] = contact_point2;
		++num;
	}
}

// PruneContactPoints: reduces num (> 4) points to at most 4
B2J_HD void prune_contact_points(V3 penetration_axis, V3 *points1, V3 *points2, int &num, V3 *projected /* MAX_MANIFOLD_POINTS */)
{
	const float cMinDistanceSq = 1.0e-6f;
	float penetration_depth_sq[MAX_MANIFOLD_POINTS];
	for (int i = 0; i < num; ++i)
	{
		V3 v1 = points1[i];
		projected[i] = v1 - dot(v1, penetration_axis) * penetration_axis;
		V3 v2 = points2[i];
		penetration_depth_sq[i] = fmax_(cMinDistanceSq, length_sq(v2 - v1));
	}
	int point1 = 0;
	float val = fmax_(cMinDistanceSq, length_sq(projected[0])) * penetration_depth_sq[0];
	for (int i = 0; i < num; ++i)
	{
		float v = fmax_(cMinDistanceSq, length_sq(projected[i])) * penetration_depth_sq[i];
		if (v > val) { val = v; point1 = i; }
	}
	V3 point1v = projected[point1];
	int point2 = -1;
	val = -FLT_MAX;
	for (int i = 0; i < num; ++i)
		if (i != point1)
		{
			float v = fmax_(cMinDistanceSq, length_sq(projected[i] - point1v)) * penetration_depth_sq[i];
			if (v > val) { val = v; point2 = i; }
		}
	V3 point2v = projected[point2];
	int point3 = -1, point4 = -1;
	float min_val = 0.0f, max_val = 0.0f;
	V3 perp = cross(point2v - point1v, penetration_axis);
	for (int i = 0; i < num; ++i)
		if (i != point1 && i != point2)
		{
			float v = dot(perp, projected[i] - point1v);
			if (v < min_val) { min_val = v; point3 = i; }
			else if (v > max_val) { max_val = v; point4 = i; }
		}
	V3 k1[4], k2[4];
	int k = 0;
	k1[k] = points1[point1]; k2[k] = points2[point1]; ++k;
	if (point3 != -1) { k1[k] = points1[point3]; k2[k] = points2[point3]; ++k; }
	k1[k] = points1[point2]; k2[k] = points2[point2]; ++k;
	if (point4 != -1) { k1[k] = points1[point4]; k2[k] = points2[point4]; ++k; }
	for (int i = 0; i < k; ++i) { points1[i] = k1[i]; points2[i] = k2[i]; }
	num = k;
}

} // namespace b2j
