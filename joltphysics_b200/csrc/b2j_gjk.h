// b2j_gjk.h -- closest point / GJK / EPA for one convex pair, single-thread device functions.
//
// Behavioural restatement (operation order, thresholds, tie-breaks, queue order) of:
//   Jolt/Geometry/ClosestPoint.h:15-493            closest point on line / triangle / tetrahedron
//   Jolt/Geometry/GJKClosestPoint.h:35-194,326-491 GetClosest, CalculatePointAAndB, GetClosestPoints
//   Jolt/Geometry/EPAPenetrationDepth.h:103-481    GetPenetrationDepthStepGJK / StepEPA
//   Jolt/Geometry/EPAConvexHullBuilder.h           hull builder (<= 256 triangles, <= 128 points), Jolt/Core/BinaryHeap.h
// so that contact counts are bit-exact against the reference's CROSS_PLATFORM_DETERMINISTIC build.
#pragma once

#include "b2j_math.h"

namespace b2j {

// a*b - c*d without FMA (Jolt/Math/Math.h DifferenceOfProducts, non-FMADD branch)
B2J_HD float diff_of_products(float a, float b, float c, float d) { return a * b - c * d; }

// ---- ClosestPoint ----------------------------------------------------------------------------------------------

B2J_HD bool cp_barycentric_line(V3 a, V3 b, float &u, float &v)
{
	V3 ab = b - a;
	float denominator = length_sq(ab);
	if (denominator < square(FLT_EPSILON))
	{
		if (length_sq(a) < length_sq(b)) { u = 1.0f; v = 0.0f; }
		else { u = 0.0f; v = 1.0f; }
		return false;
	}
	v = -dot(a, ab) / denominator;
	u = 1.0f - v;
	return true;
}

B2J_HD bool cp_barycentric_tri(V3 a, V3 b, V3 c, float &u, float &v, float &w)
{
	V3 v0 = b - a, v1 = c - a, v2 = c - b;
	float d00 = length_sq(v0), d11 = length_sq(v1), d22 = length_sq(v2);
	if (d00 <= d22)
	{
		float d01 = dot(v0, v1);
		float denominator = diff_of_products(d00, d11, d01, d01);
		if (denominator < 1.0e-12f)
		{
			if (d00 > d11) { cp_barycentric_line(a, b, u, v); w = 0.0f; }
			else { cp_barycentric_line(a, c, u, w); v = 0.0f; }
			return false;
		}
		float a0 = dot(a, v0), a1 = dot(a, v1);
		v = diff_of_products(d01, a1, d11, a0) / denominator;
		w = diff_of_products(d01, a0, d00, a1) / denominator;
		u = 1.0f - v - w;
	}
	else
	{
		float d12 = dot(v1, v2);
		float denominator = diff_of_products(d11, d22, d12, d12);
		if (denominator < 1.0e-12f)
		{
			if (d11 > d22) { cp_barycentric_line(a, c, u, w); v = 0.0f; }
			else { cp_barycentric_line(b, c, v, w); u = 0.0f; }
			return false;
		}
		float c1 = dot(c, v1), c2 = dot(c, v2);
		u = diff_of_products(d22, c1, d12, c2) / denominator;
		v = diff_of_products(d11, c2, d12, c1) / denominator;
		w = 1.0f - u - v;
	}
	return true;
}

B2J_HD V3 cp_on_line(V3 a, V3 b, uint32_t &set)
{
	float u, v;
	cp_barycentric_line(a, b, u, v);
	if (v <= 0.0f) { set = 1; return a; }
	if (u <= 0.0f) { set = 2; return b; }
	set = 3;
	return u * a + v * b;
}

template <bool MustIncludeC>
B2J_HD V3 cp_on_triangle(V3 inA, V3 inB, V3 inC, uint32_t &set)
{
	bool swap_ac;
	{
		V3 ba = inA - inB, bc = inC - inB;
		swap_ac = dot(bc, bc) < dot(ba, ba);
	}
	V3 a = swap_ac? inC : inA;
	V3 c = swap_ac? inA : inC;

	V3 ab = inB - a, ac = c - a;
	V3 n = cross(ab, ac);
	float n_len_sq = length_sq(n);

	if (n_len_sq < 1.0e-10f)
	{
		uint32_t closest_set = 4;
		V3 closest_point = inC;
		float best_dist_sq = length_sq(inC);
		if (!MustIncludeC)
		{
			float a_len_sq = length_sq(inA);
			if (a_len_sq < best_dist_sq) { closest_set = 1; closest_point = inA; best_dist_sq = a_len_sq; }
			float b_len_sq = length_sq(inB);
			if (b_len_sq < best_dist_sq) { closest_set = 2; closest_point = inB; best_dist_sq = b_len_sq; }
		}
		float ac_len_sq = length_sq(ac);
		if (ac_len_sq > square(FLT_EPSILON))
		{
			float v = clamp_(-dot(a, ac) / ac_len_sq, 0.0f, 1.0f);
			V3 q = a + v * ac;
			float dist_sq = length_sq(q);
			if (dist_sq < best_dist_sq) { closest_set = 5; closest_point = q; best_dist_sq = dist_sq; }
		}
		V3 bc = inC - inB;
		float bc_len_sq = length_sq(bc);
		if (bc_len_sq > square(FLT_EPSILON))
		{
			float v = clamp_(-dot(inB, bc) / bc_len_sq, 0.0f, 1.0f);
			V3 q = inB + v * bc;
			float dist_sq = length_sq(q);
			if (dist_sq < best_dist_sq) { closest_set = 6; closest_point = q; best_dist_sq = dist_sq; }
		}
		if (!MustIncludeC)
		{
			ab = inB - inA;
			float ab_len_sq = length_sq(ab);
			if (ab_len_sq > square(FLT_EPSILON))
			{
				float v = clamp_(-dot(inA, ab) / ab_len_sq, 0.0f, 1.0f);
				V3 q = inA + v * ab;
				float dist_sq = length_sq(q);
				if (dist_sq < best_dist_sq) { closest_set = 3; closest_point = q; best_dist_sq = dist_sq; }
			}
		}
		set = closest_set;
		return closest_point;
	}

	V3 ap = -a;
	float d1 = dot(ab, ap), d2 = dot(ac, ap);
	if (d1 <= 0.0f && d2 <= 0.0f) { set = swap_ac? 4 : 1; return a; }

	V3 bp = -inB;
	float d3 = dot(ab, bp), d4 = dot(ac, bp);
	if (d3 >= 0.0f && d4 <= d3) { set = 2; return inB; }

	if (d1 * d4 <= d3 * d2 && d1 >= 0.0f && d3 <= 0.0f)
	{
		float v = d1 / (d1 - d3);
		set = swap_ac? 6 : 3;
		return a + v * ab;
	}

	V3 cp = -c;
	float d5 = dot(ab, cp), d6 = dot(ac, cp);
	if (d6 >= 0.0f && d5 <= d6) { set = swap_ac? 1 : 4; return c; }

	if (d5 * d2 <= d1 * d6 && d2 >= 0.0f && d6 <= 0.0f)
	{
		float w = d2 / (d2 - d6);
		set = 5;
		return a + w * ac;
	}

	float d4_d3 = d4 - d3, d5_d6 = d5 - d6;
	if (d3 * d6 <= d5 * d4 && d4_d3 >= 0.0f && d5_d6 >= 0.0f)
	{
		float w = d4_d3 / (d4_d3 + d5_d6);
		set = swap_ac? 3 : 6;
		return inB + w * (c - inB);
	}

	set = 7;
	return (n * dot(a + inB + c, n)) / (3.0f * n_len_sq);
}

// OriginOutsideOfTetrahedronPlanes: bit i set = origin outside plane i (ABC, ACD, ADB, BDC)
B2J_HD uint32_t cp_origin_outside_tet_planes(V3 a, V3 b, V3 c, V3 d)
{
	V3 ab = b - a, ac = c - a, ad = d - a, bd = d - b, bc = c - b;
	V3 ab_cross_ac = cross(ab, ac), ac_cross_ad = cross(ac, ad), ad_cross_ab = cross(ad, ab), bd_cross_bc = cross(bd, bc);
	float signp[4] = { dot(a, ab_cross_ac), dot(a, ac_cross_ad), dot(a, ad_cross_ab), dot(b, bd_cross_bc) };
	float signd[4] = { dot(ad, ab_cross_ac), dot(ab, ac_cross_ad), dot(ac, ad_cross_ab), -dot(ab, bd_cross_bc) };
	int sign_bits = (signbit(signd[0])? 1 : 0) | (signbit(signd[1])? 2 : 0) | (signbit(signd[2])? 4 : 0) | (signbit(signd[3])? 8 : 0);
	uint32_t r = 0;
	if (sign_bits == 0)
	{
		for (int i = 0; i < 4; ++i) if (signp[i] >= -FLT_EPSILON) r |= 1u << i;
	}
	else if (sign_bits == 0xf)
	{
		for (int i = 0; i < 4; ++i) if (signp[i] <= FLT_EPSILON) r |= 1u << i;
	}
	else
		r = 0xf;
	return r;
}

template <bool MustIncludeD>
B2J_HD V3 cp_on_tetrahedron(V3 a, V3 b, V3 c, V3 d, uint32_t &out_set)
{
	uint32_t closest_set = 0xf;
	V3 closest_point = v3_zero();
	float best_dist_sq = FLT_MAX;
	uint32_t out_of_planes = cp_origin_outside_tet_planes(a, b, c, d);

	if (out_of_planes & 1)
	{
		if (MustIncludeD) { closest_set = 1; closest_point = a; }
		else closest_point = cp_on_triangle<false>(a, b, c, closest_set);
		best_dist_sq = length_sq(closest_point);
	}
	if (out_of_planes & 2)
	{
		uint32_t set;
		V3 q = cp_on_triangle<MustIncludeD>(a, c, d, set);
		float dist_sq = length_sq(q);
		if (dist_sq < best_dist_sq) { best_dist_sq = dist_sq; closest_point = q; closest_set = (set & 1) + ((set & 6) << 1); }
	}
	if (out_of_planes & 4)
	{
		uint32_t set;
		V3 q = cp_on_triangle<MustIncludeD>(a, b, d, set);
		float dist_sq = length_sq(q);
		if (dist_sq < best_dist_sq) { best_dist_sq = dist_sq; closest_point = q; closest_set = (set & 3) + ((set & 4) << 1); }
	}
	if (out_of_planes & 8)
	{
		uint32_t set;
		V3 q = cp_on_triangle<MustIncludeD>(b, c, d, set);
		float dist_sq = length_sq(q);
		if (dist_sq < best_dist_sq) { closest_point = q; closest_set = set << 1; }
	}
	out_set = closest_set;
	return closest_point;
}

// ---- GJK ---------------------------------------------------------------------------------------------------

struct GjkSimplex
{
	V3 y[4], p[4], q[4];
	int num_points;
};

B2J_HD bool gjk_get_closest(const GjkSimplex &s, float prev_v_len_sq, V3 &out_v, float &out_v_len_sq, uint32_t &out_set)
{
	uint32_t set;
	V3 v;
	switch (s.num_points)
	{
	case 1: set = 1; v = s.y[0]; break;
	case 2: v = cp_on_line(s.y[0], s.y[1], set); break;
	case 3: v = cp_on_triangle<true>(s.y[0], s.y[1], s.y[2], set); break;
	case 4: v = cp_on_tetrahedron<true>(s.y[0], s.y[1], s.y[2], s.y[3], set); break;
	default: return false;
	}
	float v_len_sq = length_sq(v);
	if (v_len_sq < prev_v_len_sq)
	{
		out_v = v; out_v_len_sq = v_len_sq; out_set = set;
		return true;
	}
	return false;
}

B2J_HD void gjk_calculate_point_a_and_b(const GjkSimplex &s, V3 &out_a, V3 &out_b)
{
	switch (s.num_points)
	{
	case 1: out_a = s.p[0]; out_b = s.q[0]; break;
	case 2:
		{
			float u, v;
			cp_barycentric_line(s.y[0], s.y[1], u, v);
			out_a = u * s.p[0] + v * s.p[1];
			out_b = u * s.q[0] + v * s.q[1];
		}
		break;
	case 3:
		{
			float u, v, w;
			cp_barycentric_tri(s.y[0], s.y[1], s.y[2], u, v, w);
			out_a = u * s.p[0] + v * s.p[1] + w * s.p[2];
			out_b = u * s.q[0] + v * s.q[1] + w * s.q[2];
		}
		break;
	default: break;
	}
}

// GJKClosestPoint::GetClosestPoints. A and B provide V3 support(V3 dir) const.
// One iteration of the GetClosestPoints loop; returns false when the loop ends (separated = the early out above max_dist_sq)
template <class A, class B>
B2J_HD bool gjk_closest_points_iteration(GjkSimplex &s, const A &a, const B &b, float tolerance_sq, float max_dist_sq, V3 &io_v, float &v_len_sq, float &prev_v_len_sq, bool &separated)
{
	V3 p = a.support(io_v);
	V3 q = b.support(-io_v);
	V3 w = p - q;
	float dt = dot(io_v, w);
	if (dt < 0.0f && dt * dt > v_len_sq * max_dist_sq)
	{
		separated = true;
		return false;
	}

	s.y[s.num_points] = w; s.p[s.num_points] = p; s.q[s.num_points] = q;
	++s.num_points;

	uint32_t set;
	if (!gjk_get_closest(s, prev_v_len_sq, io_v, v_len_sq, set))
	{
		--s.num_points;
		return false;
	}
	if (set == 0xf)
	{
		io_v = v3_zero();
		v_len_sq = 0.0f;
		return false;
	}
	// UpdatePointSetYPQ
	{
		int n = 0;
		for (int i = 0; i < s.num_points; ++i)
			if (set & (1u << i)) { s.y[n] = s.y[i]; s.p[n] = s.p[i]; s.q[n] = s.q[i]; ++n; }
		s.num_points = n;
	}
	if (v_len_sq <= tolerance_sq)
	{
		io_v = v3_zero();
		v_len_sq = 0.0f;
		return false;
	}
	float max_y = length_sq(s.y[0]);
	for (int i = 1; i < s.num_points; ++i) max_y = fmax_(max_y, length_sq(s.y[i]));
	if (v_len_sq <= FLT_EPSILON * max_y)
	{
		io_v = v3_zero();
		v_len_sq = 0.0f;
		return false;
	}
	io_v = -io_v;
	if (prev_v_len_sq - v_len_sq <= FLT_EPSILON * prev_v_len_sq)
		return false;
	prev_v_len_sq = v_len_sq;
	return true;
}

// GJKClosestPoint::GetClosestPoints. kLockstep (thread-per-pair kernels whose 32 lanes ALL call this, `alive` = lane has work): the
// loop trip count is made warp uniform with a vote so the lanes reconverge every iteration instead of drifting apart for good.
template <bool kLockstep = false, class A, class B>
B2J_HD float gjk_get_closest_points(GjkSimplex &s, const A &a, const B &b, float tolerance, float max_dist_sq, V3 &io_v, V3 &out_point_a, V3 &out_point_b, bool alive = true)
{
	float tolerance_sq = square(tolerance);
	s.num_points = 0;
	float v_len_sq = length_sq(io_v);
	float prev_v_len_sq = FLT_MAX;
	bool separated = false, run = alive;
	while (warp_any<kLockstep>(run))
		if (run)
			run = gjk_closest_points_iteration(s, a, b, tolerance_sq, max_dist_sq, io_v, v_len_sq, prev_v_len_sq, separated);
	if (separated || !alive)
		return FLT_MAX;
	gjk_calculate_point_a_and_b(s, out_point_a, out_point_b);
	return v_len_sq;
}

enum { PEN_NOT_COLLIDING = 0, PEN_COLLIDING = 1, PEN_INDETERMINATE = 2 };

// EPAPenetrationDepth::GetPenetrationDepthStepGJK
template <bool kLockstep = false, class AE, class BE>
B2J_HD int pen_depth_step_gjk(GjkSimplex &s, const AE &a_excl, float convex_radius_a, const BE &b_excl, float convex_radius_b, float tolerance, V3 &io_v, V3 &out_point_a, V3 &out_point_b, bool alive = true)
{
	float combined_radius = convex_radius_a + convex_radius_b;
	float combined_radius_sq = combined_radius * combined_radius;
	float closest_points_dist_sq = gjk_get_closest_points<kLockstep>(s, a_excl, b_excl, tolerance, combined_radius_sq, io_v, out_point_a, out_point_b, alive);
	if (closest_points_dist_sq > combined_radius_sq)
		return PEN_NOT_COLLIDING;
	if (closest_points_dist_sq > 0.0f)
	{
		float v_len = sqrt_(closest_points_dist_sq);
		out_point_a += io_v * (convex_radius_a / v_len);
		out_point_b -= io_v * (convex_radius_b / v_len);
		return PEN_COLLIDING;
	}
	return PEN_INDETERMINATE;
}

// ---- EPA ---------------------------------------------------------------------------------------------------

enum { EPA_MAX_TRIANGLES = 256, EPA_MAX_POINTS = 128, EPA_MAX_EDGE_LENGTH = 128, EPA_MAX_POINTS_TO_INCLUDE_ORIGIN = 32 };
#define B2J_EPA_NULL 0xffffu

struct EpaEdge { uint16_t neighbour_triangle; uint8_t neighbour_edge; uint8_t start_idx; };

struct EpaTriangle
{
	EpaEdge edge[3];
	V3 normal, centroid;
	float closest_len_sq;
	float lambda[2];
	uint8_t lambda_relative_to_0, closest_point_interior, removed, in_queue;
	uint16_t next_free;
};

struct EpaStackEntry { uint16_t tri; int8_t edge; int8_t iter; };

// Working set of one EPA run: a VIEW onto storage of some physical capacity. The reference's limits (256 triangles, 128 points,
// 128 edges) are semantic (hitting them makes EPA fail like the reference does); a smaller PHYSICAL capacity only raises
// `overflow`, in which case the caller discards the run and repeats it on full size storage (two tier scheme: the common shallow
// cases run out of ~4 KB of shared memory per warp, 5x more of them per SM than with the 21 KB worst case block).
struct EpaScratch
{
	EpaTriangle *tri;
	V3 *y, *p, *q;
	uint16_t *queue;
	EpaEdge *edges;
	uint16_t *new_triangles;
	EpaStackEntry *stack;
	int cap_tri, cap_pts, cap_edge;
	int overflow;
	int num_points, queue_size, next_free, high_watermark, num_new_triangles;
};

template <int TRI, int PTS, int EDGE> struct EpaStorage
{
	EpaTriangle tri[TRI];
	V3 y[PTS], p[PTS], q[PTS];
	uint16_t queue[TRI];
	EpaEdge edges[EDGE];
	uint16_t new_triangles[EDGE];
	EpaStackEntry stack[EDGE];
	B2J_HD EpaScratch view()
	{
		EpaScratch e;
		e.tri = tri; e.y = y; e.p = p; e.q = q; e.queue = queue; e.edges = edges; e.new_triangles = new_triangles; e.stack = stack;
		e.cap_tri = TRI; e.cap_pts = PTS; e.cap_edge = EDGE; e.overflow = 0;
		e.num_points = 0; e.queue_size = 0; e.next_free = 0; e.high_watermark = 0; e.num_new_triangles = 0;
		return e;
	}
};
using EpaStorageFull = EpaStorage<EPA_MAX_TRIANGLES, EPA_MAX_POINTS, EPA_MAX_EDGE_LENGTH>;   // 21 KB: can never overflow
using EpaStorageSmall = EpaStorage<24, 14, 12>;                                                  // 2 KB: one per THREAD, 3 warps per SM

B2J_HD bool epa_tri_is_facing(const EpaTriangle &t, V3 pos) { return dot(t.normal, pos - t.centroid) > 0.0f; }
B2J_HD bool epa_tri_is_facing_origin(const EpaTriangle &t) { return dot(t.normal, t.centroid) < 0.0f; }

// EPAConvexHullBuilder::Triangle::Triangle
B2J_HD void epa_tri_init(EpaTriangle &t, int idx0, int idx1, int idx2, const V3 *positions)
{
	t.edge[0].start_idx = (uint8_t)idx0; t.edge[1].start_idx = (uint8_t)idx1; t.edge[2].start_idx = (uint8_t)idx2;
	t.edge[0].neighbour_triangle = t.edge[1].neighbour_triangle = t.edge[2].neighbour_triangle = B2J_EPA_NULL;
	t.edge[0].neighbour_edge = t.edge[1].neighbour_edge = t.edge[2].neighbour_edge = 0;
	t.closest_len_sq = FLT_MAX;
	t.lambda[0] = t.lambda[1] = 0.0f;
	t.lambda_relative_to_0 = 0; t.closest_point_interior = 0; t.removed = 0; t.in_queue = 0;

	V3 y0 = positions[idx0], y1 = positions[idx1], y2 = positions[idx2];
	t.centroid = (y0 + y1 + y2) / 3.0f;
	V3 y10 = y1 - y0, y20 = y2 - y0, y21 = y2 - y1;
	float y20_dot_y20 = dot(y20, y20), y21_dot_y21 = dot(y21, y21);
	const float cMinTriangleArea = 1.0e-10f, cBarycentricEpsilon = 1.0e-3f;
	if (y20_dot_y20 < y21_dot_y21)
	{
		t.normal = cross(y10, y20);
		float normal_len_sq = length_sq(t.normal);
		if (normal_len_sq > cMinTriangleArea)
		{
			float c_dot_n = dot(t.centroid, t.normal);
			t.closest_len_sq = fabs_(c_dot_n) * c_dot_n / normal_len_sq;
			float y10_dot_y10 = length_sq(y10), y10_dot_y20 = dot(y10, y20);
			float determinant = diff_of_products(y10_dot_y10, y20_dot_y20, y10_dot_y20, y10_dot_y20);
			if (determinant > 0.0f)
			{
				float y0_dot_y10 = dot(y0, y10), y0_dot_y20 = dot(y0, y20);
				float l0 = diff_of_products(y10_dot_y20, y0_dot_y20, y20_dot_y20, y0_dot_y10) / determinant;
				float l1 = diff_of_products(y10_dot_y20, y0_dot_y10, y10_dot_y10, y0_dot_y20) / determinant;
				t.lambda[0] = l0; t.lambda[1] = l1; t.lambda_relative_to_0 = 1;
				if (l0 > -cBarycentricEpsilon && l1 > -cBarycentricEpsilon && l0 + l1 < 1.0f + cBarycentricEpsilon)
					t.closest_point_interior = 1;
			}
		}
	}
	else
	{
		t.normal = cross(y10, y21);
		float normal_len_sq = length_sq(t.normal);
		if (normal_len_sq > cMinTriangleArea)
		{
			float c_dot_n = dot(t.centroid, t.normal);
			t.closest_len_sq = fabs_(c_dot_n) * c_dot_n / normal_len_sq;
			float y10_dot_y10 = length_sq(y10), y10_dot_y21 = dot(y10, y21);
			float determinant = diff_of_products(y10_dot_y10, y21_dot_y21, y10_dot_y21, y10_dot_y21);
			if (determinant > 0.0f)
			{
				float y1_dot_y10 = dot(y1, y10), y1_dot_y21 = dot(y1, y21);
				float l0 = diff_of_products(y21_dot_y21, y1_dot_y10, y10_dot_y21, y1_dot_y21) / determinant;
				float l1 = diff_of_products(y10_dot_y21, y1_dot_y10, y10_dot_y10, y1_dot_y21) / determinant;
				t.lambda[0] = l0; t.lambda[1] = l1; t.lambda_relative_to_0 = 0;
				if (l0 > -cBarycentricEpsilon && l1 > -cBarycentricEpsilon && l0 + l1 < 1.0f + cBarycentricEpsilon)
					t.closest_point_interior = 1;
			}
		}
	}
}

// TriangleFactory: LIFO free list + high watermark
B2J_HD int epa_create_triangle(EpaScratch &e, int idx0, int idx1, int idx2)
{
	int t;
	if (e.next_free != (int)B2J_EPA_NULL)
	{
		t = e.next_free;
		e.next_free = e.tri[t].next_free;
	}
	else
	{
		if (e.high_watermark >= EPA_MAX_TRIANGLES)
			return -1;
		if (e.high_watermark >= e.cap_tri) { e.overflow = 1; return -1; }
		t = e.high_watermark++;
	}
	epa_tri_init(e.tri[t], idx0, idx1, idx2, e.y);
	return t;
}
B2J_HD void epa_free_triangle(EpaScratch &e, int t)
{
	e.tri[t].next_free = (uint16_t)e.next_free;
	e.next_free = t;
}

// TriangleQueue: binary heap ordered by closest_len_sq (Jolt/Core/BinaryHeap.h), pred(a, b) = a.closest > b.closest
B2J_HD bool epa_queue_pred(const EpaScratch &e, uint16_t t1, uint16_t t2) { return e.tri[t1].closest_len_sq > e.tri[t2].closest_len_sq; }
B2J_HD void epa_queue_push(EpaScratch &e, int t)
{
	e.queue[e.queue_size++] = (uint16_t)t;
	e.tri[t].in_queue = 1;
	int current = e.queue_size - 1;
	while (current > 0)
	{
		int parent = (current - 1) >> 1;
		if (epa_queue_pred(e, e.queue[parent], e.queue[current]))
		{
			uint16_t tmp = e.queue[parent]; e.queue[parent] = e.queue[current]; e.queue[current] = tmp;
			current = parent;
		}
		else
			break;
	}
}
B2J_HD int epa_queue_pop(EpaScratch &e)
{
	{ uint16_t tmp = e.queue[e.queue_size - 1]; e.queue[e.queue_size - 1] = e.queue[0]; e.queue[0] = tmp; }
	int count = e.queue_size - 1;
	int largest = 0;
	for (;;)
	{
		int child = (largest << 1) + 1;
		if (child >= count)
			break;
		int prev_largest = largest;
		if (epa_queue_pred(e, e.queue[largest], e.queue[child]))
			largest = child;
		++child;
		if (child < count && epa_queue_pred(e, e.queue[largest], e.queue[child]))
			largest = child;
		if (prev_largest == largest)
			break;
		uint16_t tmp = e.queue[prev_largest]; e.queue[prev_largest] = e.queue[largest]; e.queue[largest] = tmp;
	}
	int t = e.queue[e.queue_size - 1];
	--e.queue_size;
	return t;
}

B2J_HD void epa_link_triangle(EpaScratch &e, int t1, int edge1, int t2, int edge2)
{
	EpaEdge &e1 = e.tri[t1].edge[edge1];
	EpaEdge &e2 = e.tri[t2].edge[edge2];
	e1.neighbour_triangle = (uint16_t)t2; e1.neighbour_edge = (uint8_t)edge2;
	e2.neighbour_triangle = (uint16_t)t1; e2.neighbour_edge = (uint8_t)edge1;
}

B2J_HD void epa_unlink_triangle(EpaScratch &e, int t)
{
	for (int i = 0; i < 3; ++i)
	{
		EpaEdge &edge = e.tri[t].edge[i];
		if (edge.neighbour_triangle != B2J_EPA_NULL)
		{
			EpaEdge &neighbour_edge = e.tri[edge.neighbour_triangle].edge[edge.neighbour_edge];
			neighbour_edge.neighbour_triangle = B2J_EPA_NULL;
			edge.neighbour_triangle = B2J_EPA_NULL;
		}
	}
	if (!e.tri[t].in_queue)
		epa_free_triangle(e, t);
}

// EPAConvexHullBuilder::FindEdge; fills e.edges, returns the edge count or -1
B2J_HD int epa_find_edge(EpaScratch &e, int facing_triangle, V3 vertex)
{
	int num_edges = 0;
	e.tri[facing_triangle].removed = 1;
	int cur_stack_pos = 0;
	e.stack[0].tri = (uint16_t)facing_triangle; e.stack[0].edge = 0; e.stack[0].iter = -1;
	int next_expected_start_idx = -1;
	for (;;)
	{
		int ct = e.stack[cur_stack_pos].tri;
		if (++e.stack[cur_stack_pos].iter >= 3)
		{
			epa_unlink_triangle(e, ct);
			if (--cur_stack_pos < 0)
				break;
		}
		else
		{
			EpaEdge &ed = e.tri[ct].edge[(e.stack[cur_stack_pos].edge + e.stack[cur_stack_pos].iter) % 3];
			int n = ed.neighbour_triangle;
			if (n != (int)B2J_EPA_NULL && !e.tri[n].removed)
			{
				if (epa_tri_is_facing(e.tri[n], vertex))
				{
					e.tri[n].removed = 1;
					cur_stack_pos++;
					if (cur_stack_pos >= e.cap_edge) { e.overflow = e.cap_edge < EPA_MAX_EDGE_LENGTH; --cur_stack_pos; return -1; } // (the reference asserts at 128)
					e.stack[cur_stack_pos].tri = (uint16_t)n;
					e.stack[cur_stack_pos].edge = (int8_t)ed.neighbour_edge;
					e.stack[cur_stack_pos].iter = 0;
				}
				else
				{
					if ((int)ed.start_idx != next_expected_start_idx && next_expected_start_idx != -1)
						return -1;
					next_expected_start_idx = e.tri[n].edge[ed.neighbour_edge].start_idx;
					if (num_edges >= e.cap_edge) { e.overflow = e.cap_edge < EPA_MAX_EDGE_LENGTH; return -1; }
					e.edges[num_edges++] = ed;
				}
			}
		}
	}
	return num_edges;
}

// EPAConvexHullBuilder::AddPoint; new triangles end up in e.new_triangles[0..e.num_new_triangles)
B2J_HD bool epa_add_point(EpaScratch &e, int facing_triangle, int idx, float closest_dist_sq)
{
	e.num_new_triangles = 0;
	V3 pos = e.y[idx];
	int num_edges = epa_find_edge(e, facing_triangle, pos);
	if (num_edges < 0)
		return false;
	for (int i = 0; i < num_edges; ++i)
	{
		int nt = epa_create_triangle(e, e.edges[i].start_idx, e.edges[(i + 1) % num_edges].start_idx, idx);
		if (nt < 0)
			return false;
		e.new_triangles[e.num_new_triangles++] = (uint16_t)nt;
		const EpaTriangle &t = e.tri[nt];
		if ((t.closest_point_interior && t.closest_len_sq < closest_dist_sq) || t.closest_len_sq < 0.0f)
			epa_queue_push(e, nt);
	}
	for (int i = 0; i < num_edges; ++i)
	{
		epa_link_triangle(e, e.new_triangles[i], 0, e.edges[i].neighbour_triangle, e.edges[i].neighbour_edge);
		epa_link_triangle(e, e.new_triangles[i], 1, e.new_triangles[(i + 1) % num_edges], 2);
	}
	return true;
}

template <class A, class B>
B2J_HD V3 epa_add_support(EpaScratch &e, const A &a, const B &b, V3 direction, int &out_index)
{
	V3 p = a.support(direction);
	V3 q = b.support(-direction);
	V3 w = p - q;
	if (e.num_points >= e.cap_pts)
	{
		// physical capacity reached (semantic limits are checked by the callers before this can happen on full storage)
		e.overflow = 1;
		out_index = e.cap_pts - 1;
		return w;
	}
	out_index = e.num_points++;
	e.y[out_index] = w; e.p[out_index] = p; e.q[out_index] = q;
	return w;
}

// First half of EPAPenetrationDepth::GetPenetrationDepthStepEPA: complete the GJK simplex `s` to a hull (no long loops)
template <class AI, class BI>
B2J_HD bool epa_begin(EpaScratch &e, const GjkSimplex &s, const AI &a_incl, const BI &b_incl)
{
	e.overflow = 0;
	e.num_points = s.num_points;
	for (int i = 0; i < s.num_points; ++i) { e.y[i] = s.y[i]; e.p[i] = s.p[i]; e.q[i] = s.q[i]; }
	e.queue_size = 0;
	e.next_free = (int)B2J_EPA_NULL;
	e.high_watermark = 0;

	switch (e.num_points)
	{
	case 1:
		{
			e.num_points = 0;
			int i0;
			epa_add_support(e, a_incl, b_incl, v3(0, 1, 0), i0);
			epa_add_support(e, a_incl, b_incl, v3(-1, -1, -1), i0);
			epa_add_support(e, a_incl, b_incl, v3(1, -1, -1), i0);
			epa_add_support(e, a_incl, b_incl, v3(0, -1, 1), i0);
		}
		break;
	case 2:
		{
			V3 axis = normalized(e.y[1] - e.y[0]);
			M33 rotation = m33_rotation(q4_rotation(axis, 120.0f * (3.14159265358979323846f / 180.0f)));
			V3 dir1 = normalized_perpendicular(axis);
			Xf rot = xf(rotation, v3_zero());
			V3 dir2 = mul(rot, dir1);
			V3 dir3 = mul(rot, dir2);
			int i0;
			epa_add_support(e, a_incl, b_incl, dir1, i0);
			epa_add_support(e, a_incl, b_incl, dir2, i0);
			epa_add_support(e, a_incl, b_incl, dir3, i0);
		}
		break;
	default:
		break;
	}
	if (e.num_points < 3 || e.overflow)
		return false;

	// hull.Initialize(0, 1, 2)
	{
		int t1 = epa_create_triangle(e, 0, 1, 2);
		int t2 = epa_create_triangle(e, 0, 2, 1);
		epa_link_triangle(e, t1, 0, t2, 2);
		epa_link_triangle(e, t1, 1, t2, 1);
		epa_link_triangle(e, t1, 2, t2, 0);
		epa_queue_push(e, t1);
		epa_queue_push(e, t2);
	}

	int initial_points = e.num_points;
	for (int i = 3; i < initial_points; ++i)
	{
		// FindFacingTriangle
		int best = -1;
		float best_dist_sq = 0.0f;
		for (int qi = 0; qi < e.queue_size; ++qi)
		{
			const EpaTriangle &t = e.tri[e.queue[qi]];
			if (!t.removed)
			{
				float dt = dot(t.normal, e.y[i] - t.centroid);
				if (dt > 0.0f)
				{
					float dist_sq = dt * dt / length_sq(t.normal);
					if (dist_sq > best_dist_sq) { best = e.queue[qi]; best_dist_sq = dist_sq; }
				}
			}
		}
		if (best >= 0)
			if (!epa_add_point(e, best, i, FLT_MAX))
				return false;
	}
	return true;
}

// One iteration of the "loop until the origin is inside the hull"; returns false when the loop ends, failed = EPA gives up
template <class AI, class BI>
B2J_HD bool epa_include_origin_iteration(EpaScratch &e, const AI &a_incl, const BI &b_incl, bool &failed)
{
	int t = e.queue[0];
	if (e.tri[t].removed)
	{
		epa_queue_pop(e);
		if (e.queue_size == 0) { failed = true; return false; }
		epa_free_triangle(e, t);
		return true;
	}
	if (e.tri[t].closest_len_sq >= 0.0f)
		return false;
	epa_queue_pop(e);
	int new_index;
	V3 w = epa_add_support(e, a_incl, b_incl, e.tri[t].normal, new_index);
	if (e.overflow || !epa_tri_is_facing(e.tri[t], w) || !epa_add_point(e, t, new_index, FLT_MAX)) { failed = true; return false; }
	epa_free_triangle(e, t);
	if (e.queue_size == 0 || e.num_points >= EPA_MAX_POINTS_TO_INCLUDE_ORIGIN) { failed = true; return false; }
	return true;
}

struct EpaMainState { float closest_dist_sq; int last; bool flip_v_sign; };

// One iteration of the main EPA loop; returns false when the loop ends, failed = EPA gives up
template <class AI, class BI>
B2J_HD bool epa_main_iteration(EpaScratch &e, EpaMainState &m, const AI &a_incl, const BI &b_incl, float tolerance, bool &failed)
{
	int t = epa_queue_pop(e);
	if (e.tri[t].removed)
		epa_free_triangle(e, t);
	else
	{
		if (e.tri[t].closest_len_sq >= m.closest_dist_sq)
			return false;
		if (m.last >= 0)
			epa_free_triangle(e, m.last);
		m.last = t;

		int new_index;
		V3 tn = e.tri[t].normal;
		V3 w = epa_add_support(e, a_incl, b_incl, tn, new_index);
		if (e.overflow) { failed = true; return false; }
		float dt = dot(tn, w);
		if (dt < 0.0f) { failed = true; return false; }
		float dist_sq = square(dt) / length_sq(tn);
		if (dist_sq - e.tri[t].closest_len_sq < e.tri[t].closest_len_sq * tolerance)
			return false;
		m.closest_dist_sq = fmin_(m.closest_dist_sq, dist_sq);
		if (!epa_tri_is_facing(e.tri[t], w))
			return false;
		if (!epa_add_point(e, t, new_index, m.closest_dist_sq))
		{
			if (e.overflow) failed = true;
			return false;
		}
		bool has_defect = false;
		for (int i = 0; i < e.num_new_triangles; ++i)
			if (epa_tri_is_facing_origin(e.tri[e.new_triangles[i]])) { has_defect = true; break; }
		if (has_defect)
		{
			V3 w2 = a_incl.support(-tn) - b_incl.support(tn);
			float dot2 = -dot(tn, w2);
			if (dot2 < dt)
				m.flip_v_sign = true;
			return false;
		}
	}
	return e.queue_size > 0 && e.num_points < EPA_MAX_POINTS;
}

// EPAPenetrationDepth::GetPenetrationDepthStepEPA. The GJK simplex comes in through `s`. kLockstep / alive: see gjk_get_closest_points
// (the two long loops run with warp uniform trip counts so that the lanes of a thread-per-pair warp reconverge every iteration).
template <bool kLockstep = false, class AI, class BI>
B2J_HD bool pen_depth_step_epa(EpaScratch &e, const GjkSimplex &s, const AI &a_incl, const BI &b_incl, float tolerance, V3 &out_v, V3 &out_point_a, V3 &out_point_b, bool alive = true)
{
	if (alive)
		alive = epa_begin(e, s, a_incl, b_incl);

	// Loop until the origin is inside the hull
	bool failed = false, run = alive;
	while (warp_any<kLockstep>(run))
		if (run)
			run = epa_include_origin_iteration(e, a_incl, b_incl, failed);

	EpaMainState m;
	m.closest_dist_sq = FLT_MAX; m.last = -1; m.flip_v_sign = false;
	run = alive && !failed;
	while (warp_any<kLockstep>(run))
		if (run)
			run = epa_main_iteration(e, m, a_incl, b_incl, tolerance, failed);

	if (!alive || failed || m.last < 0 || e.overflow)
		return false;

	const EpaTriangle &lt = e.tri[m.last];
	out_v = (dot(lt.centroid, lt.normal) / length_sq(lt.normal)) * lt.normal;
	if (is_near_zero(out_v))
		return false;
	if (m.flip_v_sign)
		out_v = -out_v;

	V3 p0 = e.p[lt.edge[0].start_idx], p1 = e.p[lt.edge[1].start_idx], p2 = e.p[lt.edge[2].start_idx];
	V3 q0 = e.q[lt.edge[0].start_idx], q1 = e.q[lt.edge[1].start_idx], q2 = e.q[lt.edge[2].start_idx];
	if (lt.lambda_relative_to_0)
	{
		out_point_a = p0 + lt.lambda[0] * (p1 - p0) + lt.lambda[1] * (p2 - p0);
		out_point_b = q0 + lt.lambda[0] * (q1 - q0) + lt.lambda[1] * (q2 - q0);
	}
	else
	{
		out_point_a = p1 + lt.lambda[0] * (p0 - p1) + lt.lambda[1] * (p2 - p1);
		out_point_b = q1 + lt.lambda[0] * (q0 - q1) + lt.lambda[1] * (q2 - q1);
	}
	return true;
}

} // namespace b2j
