/* port.c -- ORACLE (test infrastructure, never linked into the product): a plain-C restatement of the pieces of the
 * reference's step that can be stated independently of the device code. Whole-step parity is anchored on the REAL
 * reference (oracle/_ref/libjoltref_det.so, built from /root/reference by oracle/Makefile); this file covers:
 *
 *   port_find_pairs      brute force O(N^2) candidate pairs with the reference's predicate
 *                        (QuadTree::FindCollidingPairs QuadTree.cpp:1458-1482, Body::sFindCollidingPairsCanCollide Body.inl:30-79,
 *                         AABox::Overlaps AABox.h:164-167, layer tables as ObjectLayerPairFilter / ObjectVsBroadPhaseLayerFilter)
 *   port_free_body_step  one step of a body without contacts: MotionProperties::ApplyForceTorqueAndDragInternal
 *                        (MotionProperties.inl:127-149), JobIntegrateVelocity (PhysicsSystem.cpp:1583-1711), Body::AddRotationStep
 *                        (Body.inl:81-98) incl. the cephes style Vec4::SinCos (Vec4.inl:1171-1231)
 *   port_hash_sub_shape_id_pair / port_hash64   FNV-1a over SubShapeIDPair and Thomas Wang's Hash64 (HashCombine.h:15-24,43-55)
 *
 * Pinned by tests/test_oracle.py against outputs of the reference itself (jref_find_pairs, jref_step) -- parity is NOT unpinned.
 * Compile: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile.port).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define MOTION_STATIC 0
#define MOTION_KINEMATIC 1
#define MOTION_DYNAMIC 2
#define FLAG_SENSOR 1u
#define FLAG_KIN_VS_NONDYN 16u
#define INACTIVE 0xffffffffu

/* Candidate pairs as (min id, max id), unsorted. bounds = [n][6] (min xyz, max xyz). Returns the number of pairs (may exceed cap). */
uint32_t port_find_pairs(uint32_t n, const uint32_t *ids, const float *bounds, const uint8_t *motion_type, const uint16_t *object_layer,
	const uint16_t *flags, const uint32_t *active_index, uint32_t num_object_layers, uint32_t num_bp_layers, const uint8_t *object_to_bp,
	const uint8_t *object_vs_bp, const uint8_t *object_vs_object, float speculative_contact_distance, uint32_t *out_pairs, uint32_t cap)
{
	uint32_t count = 0;
	for (uint32_t a = 0; a < n; ++a)
	{
		if (ids[a] == 0xffffffffu || active_index[a] == INACTIVE)
			continue; /* only active bodies query */
		float min1[3], max1[3];
		for (int k = 0; k < 3; ++k) { min1[k] = bounds[6 * a + k] - speculative_contact_distance; max1[k] = bounds[6 * a + 3 + k] + speculative_contact_distance; }
		for (uint32_t b = 0; b < n; ++b)
		{
			if (b == a || ids[b] == 0xffffffffu)
				continue;
			/* BroadPhaseQuadTree.cpp:588: the tree of layer l is only visited when the object layer of the query collides with it */
			if (!object_vs_bp[object_layer[a] * num_bp_layers + object_to_bp[object_layer[b]]])
				continue;
			if (!object_vs_object[object_layer[a] * num_object_layers + object_layer[b]])
				continue;
			/* Body::sFindCollidingPairsCanCollide */
			int dyn1 = motion_type[a] == MOTION_DYNAMIC, dyn2 = motion_type[b] == MOTION_DYNAMIC;
			int kin1 = motion_type[a] == MOTION_KINEMATIC, kin2 = motion_type[b] == MOTION_KINEMATIC;
			if (!(flags[a] & FLAG_KIN_VS_NONDYN) && !(flags[b] & FLAG_KIN_VS_NONDYN) && (!dyn1 && !dyn2)
				&& !(kin1 && (flags[b] & FLAG_SENSOR)) && !(kin2 && (flags[a] & FLAG_SENSOR)))
				continue;
			if (active_index[a] >= active_index[b])
				continue;
			/* AABox::Overlaps with only the query box expanded */
			const float *min2 = bounds + 6 * b, *max2 = bounds + 6 * b + 3;
			if (min1[0] > max2[0] || min1[1] > max2[1] || min1[2] > max2[2] || max1[0] < min2[0] || max1[1] < min2[1] || max1[2] < min2[2])
				continue;
			if (count < cap)
			{
				out_pairs[2 * count] = ids[a] < ids[b]? ids[a] : ids[b];
				out_pairs[2 * count + 1] = ids[a] < ids[b]? ids[b] : ids[a];
			}
			++count;
		}
	}
	return count;
}

static float dot3(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + 0.0f); }

/* Vec4::SinCos, one lane */
void port_sin_cos(float in, float *out_sin, float *out_cos)
{
	uint32_t bits; memcpy(&bits, &in, 4);
	uint32_t sin_sign = bits & 0x80000000u;
	uint32_t xb = bits ^ sin_sign;
	float x; memcpy(&x, &xb, 4);
	uint32_t quadrant = (uint32_t)(int32_t)(0.6366197723675814f * x + 0.5f);
	float fq = (float)(int32_t)quadrant;
	x = ((x - fq * 1.5703125f) - fq * 0.0004837512969970703125f) - fq * 7.549789948768648e-8f;
	float x2 = x * x;
	float taylor_cos = ((2.443315711809948e-5f * x2 - 1.388731625493765e-3f) * x2 + 4.166664568298827e-2f) * x2 * x2 - 0.5f * x2 + 1.0f;
	float taylor_sin = ((-1.9515295891e-4f * x2 + 8.3321608736e-3f) * x2 - 1.6666654611e-1f) * x2 * x + x;
	uint32_t bit1 = quadrant << 31, bit2 = (quadrant << 30) & 0x80000000u;
	float s = bit1? taylor_cos : taylor_sin, c = bit1? taylor_sin : taylor_cos;
	uint32_t sb, cb; memcpy(&sb, &s, 4); memcpy(&cb, &c, 4);
	sb ^= sin_sign ^ bit2; cb ^= bit1 ^ bit2;
	memcpy(out_sin, &sb, 4); memcpy(out_cos, &cb, 4);
}

/* One step of a free dynamic body (no contacts, no torque, identity inertia handling not needed: torque must be zero).
 * pos[3], rot[4] (xyzw), lin[3], ang[3] are updated in place. */
void port_free_body_step(float *pos, float *rot, float *lin, float *ang, const float *gravity, float gravity_factor, float inv_mass, const float *force,
	float linear_damping, float angular_damping, float max_linear_velocity, float max_angular_velocity, float dt)
{
	/* ApplyForceTorqueAndDragInternal */
	for (int k = 0; k < 3; ++k) lin[k] = lin[k] + dt * (gravity_factor * gravity[k] + inv_mass * force[k]);
	float ld = 1.0f - linear_damping * dt; if (ld < 0.0f) ld = 0.0f;
	float ad = 1.0f - angular_damping * dt; if (ad < 0.0f) ad = 0.0f;
	for (int k = 0; k < 3; ++k) { lin[k] *= ld; ang[k] *= ad; }
	for (int pass = 0; pass < 2; ++pass) /* clamp in gravity job and again in integrate */
	{
		float l2 = dot3(lin, lin);
		if (l2 > max_linear_velocity * max_linear_velocity) { float s = max_linear_velocity / sqrtf(l2); for (int k = 0; k < 3; ++k) lin[k] *= s; }
		float a2 = dot3(ang, ang);
		if (a2 > max_angular_velocity * max_angular_velocity) { float s = max_angular_velocity / sqrtf(a2); for (int k = 0; k < 3; ++k) ang[k] *= s; }
	}
	/* Body::AddRotationStep(ang * dt) */
	float wdt[3] = { ang[0] * dt, ang[1] * dt, ang[2] * dt };
	float len = sqrtf(dot3(wdt, wdt));
	if (len > 1.0e-6f)
	{
		float s, c;
		port_sin_cos(0.5f * len, &s, &c);
		float a = wdt[0] / len * s, b = wdt[1] / len * s, cc = wdt[2] / len * s, d = c;
		float x = rot[0], y = rot[1], z = rot[2], w = rot[3];
		float q[4] = { (a * w + b * z) + (d * x - cc * y), (b * w + cc * x) + (d * y - a * z), (cc * w + a * y) + (d * z - b * x), -(a * x + b * y) + (d * w - cc * z) };
		float l = sqrtf((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
		for (int k = 0; k < 4; ++k) rot[k] = q[k] / l;
	}
	for (int k = 0; k < 3; ++k) pos[k] += lin[k] * dt;
}

uint64_t port_hash_sub_shape_id_pair(uint32_t body1, uint32_t sub1, uint32_t body2, uint32_t sub2)
{
	uint32_t w[4] = { body1, sub1, body2, sub2 };
	const uint8_t *data = (const uint8_t *)w; /* little endian, like the reference's HashBytes over the struct */
	uint64_t hash = 0xcbf29ce484222325ull;
	for (int i = 0; i < 16; ++i) { hash ^= (uint64_t)data[i]; hash *= 0x100000001b3ull; }
	return hash;
}

uint64_t port_hash64(uint64_t v)
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
