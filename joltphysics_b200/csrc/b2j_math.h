// b2j_math.h -- scalar fp32 vector math with the reference's operation ORDER (not its code).
//
// Parity with the reference's CROSS_PLATFORM_DETERMINISTIC build needs IEEE-identical results, so every expression here
// follows the order defined by the scalar fallbacks of Jolt/Math/*.inl: Dot = (x*x' + y*y') + (z*z' + 0)
// (Vec3.inl:898-934), unary minus = 0 - x (:669-700), M*v = (c0*x + c1*y) + c2*z (Mat44.inl Multiply3x3), quaternion
// product bracketed as in Quat.inl:7-88, Mat44::sRotation with x+x (Mat44.inl:85-140), Normalized = v / sqrt(dot).
// Compile with -fmad=false (device) / -ffp-contract=off (host): no fused multiply-add anywhere, denormals kept.
#pragma once

#include "b2j_platform.h"

namespace b2j {

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
struct M33 { V3 c0, c1, c2; }; // column major 3x3

B2J_HD float fmax_(float a, float b) { return a < b? b : a; }  // std::max
B2J_HD float fmin_(float a, float b) { return b < a? b : a; }  // std::min
B2J_HD float fabs_(float a) { return fabsf(a); }
B2J_HD float sqrt_(float a) { return sqrtf(a); }
B2J_HD float square(float a) { return a * a; }
B2J_HD float clamp_(float v, float lo, float hi) { return fmin_(fmax_(v, lo), hi); } // Clamp = min(max(v, lo), hi)  Math.h
B2J_HD float sign_(float v) { return v < 0.0f? -1.0f : 1.0f; }

B2J_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
B2J_HD V3 v3_zero() { return v3(0.0f, 0.0f, 0.0f); }
B2J_HD V3 v3_rep(float v) { return v3(v, v, v); }
B2J_HD V3 v3_load(const float *p) { return v3(p[0], p[1], p[2]); }
B2J_HD void v3_store(V3 v, float *p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
B2J_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2J_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2J_HD V3 operator-(V3 a) { return v3(0.0f - a.x, 0.0f - a.y, 0.0f - a.z); }
B2J_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
B2J_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
B2J_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
B2J_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
B2J_HD V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
B2J_HD V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
B2J_HD V3 &operator-=(V3 &a, V3 b) { a = a - b; return a; }
B2J_HD V3 &operator*=(V3 &a, float s) { a = a * s; return a; }
B2J_HD bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
B2J_HD float v3_get(V3 a, int i) { return i == 0? a.x : (i == 1? a.y : a.z); }
B2J_HD void v3_set(V3 &a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }
B2J_HD float reduce_sum(V3 a) { return (a.x + a.y) + (a.z + 0.0f); }
B2J_HD float dot(V3 a, V3 b) { return reduce_sum(a * b); }
B2J_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
B2J_HD float length_sq(V3 a) { return dot(a, a); }
B2J_HD float length(V3 a) { return sqrt_(length_sq(a)); }
B2J_HD V3 normalized(V3 a) { return a / length(a); }
B2J_HD V3 normalized_or(V3 a, V3 zero_value) { float l = length_sq(a); return l <= FLT_MIN? zero_value : a / sqrt_(l); }
B2J_HD V3 v3_min(V3 a, V3 b) { return v3(fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)); }
B2J_HD V3 v3_max(V3 a, V3 b) { return v3(fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)); }
B2J_HD V3 v3_abs(V3 a) { return v3(fabs_(a.x), fabs_(a.y), fabs_(a.z)); }
B2J_HD float reduce_min(V3 a) { return fmin_(fmin_(a.x, a.y), a.z); }
B2J_HD float reduce_max(V3 a) { return fmax_(fmax_(a.x, a.y), a.z); }
B2J_HD bool is_close(V3 a, V3 b, float max_dist_sq) { return length_sq(b - a) <= max_dist_sq; } // Vec3::IsClose
B2J_HD bool is_near_zero(V3 a, float max_dist_sq = 1.0e-12f) { return length_sq(a) <= max_dist_sq; }
B2J_HD bool v3_is_nan(V3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }
B2J_HD int lowest_component_index(V3 a) { return a.x < a.y? (a.z < a.x? 2 : 0) : (a.z < a.y? 2 : 1); }   // Vec3::GetLowestComponentIndex
B2J_HD int highest_component_index(V3 a) { return a.x > a.y? (a.z > a.x? 2 : 0) : (a.z > a.y? 2 : 1); }  // Vec3::GetHighestComponentIndex
// Vec3::GetNormalizedPerpendicular (Vec3.inl:1119-1158)
B2J_HD V3 normalized_perpendicular(V3 a)
{
	float xx = a.x * a.x, yy = a.y * a.y, zz = a.z * a.z;
	V3 perp = xx > yy? v3(a.z, 0.0f, 0.0f - a.x) : v3(0.0f, a.z, 0.0f - a.y);
	return perp / sqrt_(fmax_(xx, yy) + zz);
}
// Vec3::GetSign: 1 for >= +0, -1 for sign bit set
B2J_HD V3 v3_sign(V3 a) { return v3(signbit(a.x)? -1.0f : 1.0f, signbit(a.y)? -1.0f : 1.0f, signbit(a.z)? -1.0f : 1.0f); }

// ---- quaternion --------------------------------------------------------------------------------------------
B2J_HD Q4 q4(float x, float y, float z, float w) { Q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
B2J_HD Q4 q4_identity() { return q4(0.0f, 0.0f, 0.0f, 1.0f); }
B2J_HD Q4 q4_load(const float *p) { return q4(p[0], p[1], p[2], p[3]); }
B2J_HD void q4_store(Q4 q, float *p) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
B2J_HD V3 q4_xyz(Q4 q) { return v3(q.x, q.y, q.z); }
B2J_HD float q4_dot(Q4 a, Q4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); } // Vec4::ReduceSum order
B2J_HD Q4 q4_conj(Q4 q) { return q4(-q.x, -q.y, -q.z, q.w); } // FlipSign: pure sign bit flip
B2J_HD Q4 q4_normalized(Q4 q) { float l = sqrt_(q4_dot(q, q)); return q4(q.x / l, q.y / l, q.z / l, q.w / l); }
// Quat::operator* (Quat.inl:7-88)
B2J_HD Q4 operator*(Q4 l, Q4 r)
{
	float a = l.x, b = l.y, c = l.z, d = l.w, x = r.x, y = r.y, z = r.z, w = r.w;
	return q4((a * w + b * z) + (d * x - c * y),
			  (b * w + c * x) + (d * y - a * z),
			  (c * w + a * y) + (d * z - b * x),
			  -(a * x + b * y) + (d * w - c * z));
}
// Quat::operator*(Vec3) (Quat.inl:383-405)
B2J_HD V3 rotate(Q4 q, V3 p)
{
	V3 xyz = q4_xyz(q);
	// (p.yzx * xyz - yzx * p).yzx: component j of the un-swizzled = p[j+1]*q[j] - q[j+1]*p[j]; after .yzx x <- (p.z*q.y - q.z*p.y)
	V3 c = v3(p.z * xyz.y - xyz.z * p.y, p.x * xyz.z - xyz.x * p.z, p.y * xyz.x - xyz.y * p.x);
	V3 q_cross_q_cross_p = v3(c.z * xyz.y - xyz.z * c.y, c.x * xyz.z - xyz.x * c.z, c.y * xyz.x - xyz.y * c.x);
	V3 v = v3(q.w * c.x, q.w * c.y, q.w * c.z) + q_cross_q_cross_p;
	return p + (v + v);
}
// Quat::InverseRotate (Quat.inl:407-416)
B2J_HD V3 inverse_rotate(Q4 q, V3 p)
{
	V3 xyz = q4_xyz(q);
	// (yzx * p - p.yzx * xyz).yzx
	V3 c = v3(xyz.z * p.y - p.z * xyz.y, xyz.x * p.z - p.x * xyz.z, xyz.y * p.x - p.y * xyz.x);
	V3 cc = v3(xyz.z * c.y - c.z * xyz.y, xyz.x * c.z - c.x * xyz.z, xyz.y * c.x - c.y * xyz.x);
	V3 v = v3(q.w * c.x, q.w * c.y, q.w * c.z) + cc;
	return p + (v + v);
}
// Quat::EnsureWPositive: flips all sign bits if w has its sign bit set
B2J_HD Q4 q4_ensure_w_positive(Q4 q) { return signbit(q.w)? q4(-q.x, -q.y, -q.z, -q.w) : q; }
// Quat::sLoadFloat3Unsafe (Quat.inl:453-458)
B2J_HD Q4 q4_from_xyz(V3 v) { float w = sqrt_(fmax_(1.0f - length_sq(v), 0.0f)); return q4(v.x, v.y, v.z, w); }

// Vec4::SinCos for one lane (Vec4.inl:1171-1231): cephes style polynomial with 3 term Cody-Waite reduction
B2J_HD void sin_cos(float in, float &out_sin, float &out_cos)
{
	uint32_t bits;
	memcpy(&bits, &in, 4);
	uint32_t sin_sign = bits & 0x80000000u;
	uint32_t xb = bits ^ sin_sign;
	float x;
	memcpy(&x, &xb, 4);
	uint32_t quadrant = (uint32_t)(int32_t)(0.6366197723675814f * x + 0.5f); // ToInt truncates
	float fq = (float)(int32_t)quadrant;
	x = ((x - fq * 1.5703125f) - fq * 0.0004837512969970703125f) - fq * 7.549789948768648e-8f;
	float x2 = x * x;
	float taylor_cos = ((2.443315711809948e-5f * x2 - 1.388731625493765e-3f) * x2 + 4.166664568298827e-2f) * x2 * x2 - 0.5f * x2 + 1.0f;
	float taylor_sin = ((-1.9515295891e-4f * x2 + 8.3321608736e-3f) * x2 - 1.6666654611e-1f) * x2 * x + x;
	uint32_t bit1 = quadrant << 31;
	uint32_t bit2 = (quadrant << 30) & 0x80000000u;
	float s = bit1? taylor_cos : taylor_sin;
	float c = bit1? taylor_sin : taylor_cos;
	sin_sign ^= bit2;
	uint32_t cos_sign = bit1 ^ bit2;
	uint32_t sb, cb;
	memcpy(&sb, &s, 4);
	memcpy(&cb, &c, 4);
	sb ^= sin_sign;
	cb ^= cos_sign;
	memcpy(&out_sin, &sb, 4);
	memcpy(&out_cos, &cb, 4);
}
// Quat::sRotation(axis, angle) (Quat.inl:160-167)
B2J_HD Q4 q4_rotation(V3 axis, float angle)
{
	float s, c;
	sin_cos(0.5f * angle, s, c);
	return q4(axis.x * s, axis.y * s, axis.z * s, c);
}

// ---- 3x3 matrix -----------------------------------------------------------------------------------------------
B2J_HD M33 m33(V3 c0, V3 c1, V3 c2) { M33 m; m.c0 = c0; m.c1 = c1; m.c2 = c2; return m; }
B2J_HD M33 m33_zero() { return m33(v3_zero(), v3_zero(), v3_zero()); }
B2J_HD M33 m33_identity() { return m33(v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)); }
// Mat44::sRotation(Quat) (Mat44.inl:85-140)
B2J_HD M33 m33_rotation(Q4 q)
{
	float x = q.x, y = q.y, z = q.z, w = q.w;
	float tx = x + x, ty = y + y, tz = z + z;
	float xx = tx * x, yy = ty * y, zz = tz * z, xy = tx * y, xz = tx * z, xw = tx * w, yz = ty * z, yw = ty * w, zw = tz * w;
	return m33(v3((1.0f - yy) - zz, xy + zw, xz - yw),
			   v3(xy - zw, (1.0f - zz) - xx, yz + xw),
			   v3(xz + yw, yz - xw, (1.0f - xx) - yy));
}
// Mat44::Multiply3x3(Vec3): (c0*x + c1*y) + c2*z
B2J_HD V3 mul(const M33 &m, V3 v)
{
	return v3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z,
			  m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
			  m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
B2J_HD M33 transposed(const M33 &m) { return m33(v3(m.c0.x, m.c1.x, m.c2.x), v3(m.c0.y, m.c1.y, m.c2.y), v3(m.c0.z, m.c1.z, m.c2.z)); }
// Mat44::Multiply3x3Transposed(Vec3) = Transposed3x3().Multiply3x3(v)
B2J_HD V3 mul_transposed(const M33 &m, V3 v) { return mul(transposed(m), v); }
// Mat44::Multiply3x3(Mat44): col i = (c0*m.ci.x + c1*m.ci.y) + c2*m.ci.z
B2J_HD M33 mul(const M33 &a, const M33 &b) { return m33(mul(a, b.c0), mul(a, b.c1), mul(a, b.c2)); }
// Mat44::Multiply3x3RightTransposed: result.col[j] = c0*m.c0[j] + c1*m.c1[j] + c2*m.c2[j]
B2J_HD M33 mul_right_transposed(const M33 &a, const M33 &b)
{
	return m33(a.c0 * b.c0.x + a.c1 * b.c1.x + a.c2 * b.c2.x,
			   a.c0 * b.c0.y + a.c1 * b.c1.y + a.c2 * b.c2.y,
			   a.c0 * b.c0.z + a.c1 * b.c1.z + a.c2 * b.c2.z);
}
B2J_HD M33 operator*(float s, const M33 &m) { return m33(s * m.c0, s * m.c1, s * m.c2); } // operator*(float, Mat44): each column * s
B2J_HD V3 m33_col(const M33 &m, int i) { return i == 0? m.c0 : (i == 1? m.c1 : m.c2); }

// Rigid transform: rotation matrix + translation (Mat44 with last row 0 0 0 1)
struct Xf { M33 r; V3 t; };
B2J_HD Xf xf(const M33 &r, V3 t) { Xf x; x.r = r; x.t = t; return x; }
// Mat44 * Vec3: ((c0*x + c1*y) + c2*z) + c3
B2J_HD V3 mul(const Xf &m, V3 v)
{
	return v3(m.r.c0.x * v.x + m.r.c1.x * v.y + m.r.c2.x * v.z + m.t.x,
			  m.r.c0.y * v.x + m.r.c1.y * v.y + m.r.c2.y * v.z + m.t.y,
			  m.r.c0.z * v.x + m.r.c1.z * v.y + m.r.c2.z * v.z + m.t.z);
}
// Mat44::sRotationTranslation(q, t)
B2J_HD Xf xf_rotation_translation(Q4 q, V3 t) { return xf(m33_rotation(q), t); }
// Mat44::sInverseRotationTranslation(q, t): m = sRotation(q.Conjugated()); translation = -(m.Multiply3x3(t))
B2J_HD Xf xf_inverse_rotation_translation(Q4 q, V3 t)
{
	M33 m = m33_rotation(q4_conj(q));
	return xf(m, -mul(m, t));
}
// Mat44 * Mat44 for two rotation-translation matrices (Mat44.inl operator*): col i = a.c0*b.ci.x + a.c1*b.ci.y + a.c2*b.ci.z (+ a.c3 * 0),
// col 3 = a.c0*b.t.x + a.c1*b.t.y + a.c2*b.t.z + a.c3 * 1
B2J_HD Xf mul(const Xf &a, const Xf &b)
{
	Xf r;
	r.r.c0 = (a.r.c0 * b.r.c0.x + a.r.c1 * b.r.c0.y + a.r.c2 * b.r.c0.z) + a.t * 0.0f;
	r.r.c1 = (a.r.c0 * b.r.c1.x + a.r.c1 * b.r.c1.y + a.r.c2 * b.r.c1.z) + a.t * 0.0f;
	r.r.c2 = (a.r.c0 * b.r.c2.x + a.r.c1 * b.r.c2.y + a.r.c2 * b.r.c2.z) + a.t * 0.0f;
	r.t = (a.r.c0 * b.t.x + a.r.c1 * b.t.y + a.r.c2 * b.t.z) + a.t * 1.0f;
	return r;
}

// ---- hashes (integer, bit exact) --------------------------------------------------------------------------------
// HashBytes = FNV-1a 64 (Jolt/Core/HashCombine.h:15-24) over the 16 byte SubShapeIDPair {body1, sub1, body2, sub2}
B2J_HD uint64_t hash_sub_shape_id_pair(uint32_t body1, uint32_t sub1, uint32_t body2, uint32_t sub2)
{
	uint64_t hash = 0xcbf29ce484222325ull;
	uint32_t w[4] = { body1, sub1, body2, sub2 };
	for (int i = 0; i < 4; ++i)
		for (int b = 0; b < 4; ++b)
		{
			hash ^= (uint64_t)((w[i] >> (8 * b)) & 0xff);
			hash *= 0x100000001b3ull;
		}
	return hash;
}
// Hash64, Thomas Wang (HashCombine.h:43-55), used for BodyPair keys
B2J_HD uint64_t hash64(uint64_t v)
{
	uint64_t hash = v;
	hash = (~hash) + (hash << 21);
	hash = hash ^ (hash >> 24);
	hash = (hash + (hash << 3)) + (hash << 8);
	hash = hash ^ (hash >> 14);
	hash = (hash + (hash << 2)) + (hash << 4);
	hash = hash ^ (hash >> 28);
	hash = hash + (hash << 31);
	return hash;
}

} // namespace b2j
