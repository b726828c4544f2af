// b2j_joints.h -- non contact constraints on the step ("the simplest joints slot into H12-H14 through
// ConstraintManager::sBuildIslands / sSolve*"): PointConstraint and DistanceConstraint between two bodies.
//
// Restates
//   ConstraintManager::GetActiveConstraints / sBuildIslands / sSortConstraints / sSetupVelocityConstraints / sWarmStart... / sSolve...
//                                                     Jolt/Physics/Constraints/ConstraintManager.cpp:72-182
//   TwoBodyConstraint::IsActive / BuildIslands / BuildIslandSplits   TwoBodyConstraint.h:38, TwoBodyConstraint.cpp:17-57
//   PointConstraint                                   PointConstraint.cpp:70-118, ConstraintPart/PointConstraintPart.h:50-237
//   DistanceConstraint                                DistanceConstraint.cpp:98-198, ConstraintPart/AxisConstraintPart.h:60-485 (no springs)
//   HingeConstraint (motor off, limits without spring) HingeConstraint.cpp:137-318, ConstraintPart/HingeRotationConstraintPart.h:44-190,
//                                                     ConstraintPart/AngleConstraintPart.h:38-215, Quat::GetRotationAngle Quat.h:197,
//                                                     Vec4::ATan Vec4.inl (cephes atanf), CenterAngleAroundZero Math.h:28-44
//   FixedConstraint                                   FixedConstraint.cpp:82-112, ConstraintPart/RotationEulerConstraintPart.h:139-230
//   where the step calls them                         PhysicsSystem.cpp:720-744 (active constraints), :795-828 (setup, islands: bodies a
//                                                     constraint wakes up join the active list but get no gravity this step, :746-791),
//                                                     :1415-1427, :1503-1540 (warm start, velocity), :2596-2603, :2661-2672 (position)
//   LargeIslandSplitter::SplitIsland                  LargeIslandSplitter.cpp:236-300: contacts are coloured first, then the constraints
//
// How they join the schedule (b2j_solver.h): an active joint is one more ITEM next to the contact constraints. The items of a step are
// [active joints in (priority, constraint index) order] ++ [contacts in sort key order] -- the order the reference solves an island in
// (constraints, then contacts). Small islands and the serial split of a large island get phase = dependency depth in that order; the
// colouring of a large island walks contacts first, then joints (a body's adjacency list is read rotated in that pass). A phase's
// items share no dynamic body, so its joints and its contacts run as two launches over the same range of solve positions: joint items
// carry META_JOINT in their header and the contact kernels leave them alone, and the other way round.
#pragma once

#include "b2j_solver.h"

namespace b2j {

enum { JOINT_POINT = B2J_CONSTRAINT_POINT, JOINT_DISTANCE = B2J_CONSTRAINT_DISTANCE, JOINT_HINGE = B2J_CONSTRAINT_HINGE, JOINT_FIXED = B2J_CONSTRAINT_FIXED };
enum : uint32_t { JOINT_ENABLED = 1u, SRC_JOINT = 0x80000000u };

// what the caller described (b2j_constraint_desc) with the bodies resolved to slots
struct alignas(16) JointDef
{
	uint32_t type, b1, b2, flags;          // b1 / b2: body slots
	uint32_t priority, steps_override;     // velocity steps override | position steps override << 8
	uint32_t index, pad;                   // Constraint::mConstraintIndex (position in the world's list)
	F4 local1, local2;                     // mLocalSpacePosition1 / 2 (relative to the centre of mass); local1.w = min distance, local2.w = max distance
	F4 axis1, axis2;                       // hinge: mLocalSpaceHingeAxis1 / 2; axis1.w = mLimitsMin, axis2.w = mLimitsMax
	F4 inv_initial_orientation;            // hinge: mInvInitialOrientation
	F4 hinge;                              // hinge: x = mMaxFrictionTorque
};

// the members of the reference's constraint objects that live across kernels (and, the first two, across steps)
struct alignas(16) JointState
{
	F4 lambda;                             // mTotalLambda: point xyz, distance x
	F4 normal;                             // distance: mWorldSpaceNormal (kept when the two points coincide)
	F4 wsp1, wsp2;                         // distance: mWorldSpacePosition1 / 2; wsp1.w = mMinLambda, wsp2.w = mMaxLambda
	F4 r1, r2;                             // point: mR1, mR2; distance: mR1PlusUxAxis, mR2xAxis, r1.w = mEffectiveMass
	F4 i1[3], i2[3];                       // point: columns of mInvI1_R1X / mInvI2_R2X; distance: [0] = mInvI1_R1PlusUxAxis / mInvI2_R2xAxis
	F4 eff[3];                             // point: columns of mEffectiveMass
	// hinge (its point part uses the members above)
	F4 lambda2;                            // x, y: mRotationConstraintPart.mTotalLambda, z: mRotationLimitsConstraintPart, w: mMotorConstraintPart
	F4 h_a1, h_b2, h_c2;                   // rotation part mA1, mB2, mC2; h_a1.w = mTheta, h_b2.w = limits effective mass, h_c2.w = motor effective mass
	F4 h_b2xa1, h_c2xa1;
	F4 h_inv1[3], h_inv2[3];               // rotation part mInvI1 / mInvI2
	F4 h_eff;                              // rotation part mEffectiveMass: (0,0), (0,1), (1,0), (1,1)
	F4 h_axis;                             // HingeConstraint::mA1 (world space hinge axis of body 1)
	F4 h_l1, h_l2, h_m1, h_m2;             // limits / motor part mInvI1_Axis, mInvI2_Axis
	// fixed: RotationEulerConstraintPart (mInvI1 / mInvI2 in h_inv1 / h_inv2, mTotalLambda in lambda2.xyz); its point part uses the members above
	F4 f_eff[3];                           // columns of its mEffectiveMass
	// what the velocity passes need of the two bodies (written by the setup: no gather of the body arrays per pass)
	F4 misc;                               // inverse mass of body 1 / 2, bits: motion type 1 | motion type 2 << 2 | allowed DOFs 1 << 4 | allowed DOFs 2 << 10
};

struct JointCtx
{
	JointDef *defs;
	JointState *state;
	uint32_t num_joints;
	uint32_t *active_flag;       // [num_joints] by constraint index: active this step
	const uint32_t *order;       // [num_joints] constraint indices sorted by (priority, index)
	uint32_t *order_flag, *order_scan; // [num_joints] flags / exclusive scan in that order
	uint32_t *active_joints;     // [J] the active joints in (priority, index) order
	uint32_t *wake_key;          // per body slot: first (constraint, body) that wakes it up (BodyManager::ActivateBodies call order)
};

// ---- Mat44 helpers the point constraint needs (Mat44.inl: sCrossProduct, GetDeterminant3x3, Adjointed3x3, SetInversed3x3) -----------
B2J_HD M33 m33_cross_product(V3 v) { return m33(v3(0.0f, v.z, 0.0f - v.y), v3(0.0f - v.z, 0.0f, v.x), v3(v.y, 0.0f - v.x, 0.0f)); }
B2J_HD M33 m33_add(const M33 &a, const M33 &b) { return m33(a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2); }
B2J_HD bool m33_inversed(const M33 &m, M33 &out)
{
	float det = dot(m.c0, cross(m.c1, m.c2));
	if (det == 0.0f)
		return false;
	// Adjointed3x3: JPH_EL(r, c) = column c, row r
	V3 a0 = v3(m.c1.y, m.c2.y, m.c0.y) * v3(m.c2.z, m.c0.z, m.c1.z) - v3(m.c2.y, m.c0.y, m.c1.y) * v3(m.c1.z, m.c2.z, m.c0.z);
	V3 a1 = v3(m.c2.x, m.c0.x, m.c1.x) * v3(m.c1.z, m.c2.z, m.c0.z) - v3(m.c1.x, m.c2.x, m.c0.x) * v3(m.c2.z, m.c0.z, m.c1.z);
	V3 a2 = v3(m.c1.x, m.c2.x, m.c0.x) * v3(m.c2.y, m.c0.y, m.c1.y) - v3(m.c2.x, m.c0.x, m.c1.x) * v3(m.c1.y, m.c2.y, m.c0.y);
	out = m33(a0 / det, a1 / det, a2 / det);
	return true;
}

struct JointBody
{
	uint32_t slot, type, dofs;
	V3 x; Q4 q;
	float inv_mass;
	V3 diag; Q4 irot;
};

B2J_D JointBody joint_body(const DWorld &w, uint32_t b)
{
	JointBody k;
	BodyInfo info = w.info[b];
	k.slot = b; k.type = info.motion_type; k.dofs = info.allowed_dofs;
	k.x = to_v3(w.position[b]); k.q = to_q4(w.rotation[b]);
	if (k.type == B2J_MOTION_DYNAMIC) { k.inv_mass = w.params[b].inv_mass; k.diag = to_v3(w.inv_inertia_diag[b]); k.irot = to_q4(w.inertia_rotation[b]); }
	else { k.inv_mass = 0.0f; k.diag = v3_zero(); k.irot = q4_identity(); }
	return k;
}

// the two bodies as the velocity passes see them: slot, motion type, DOFs and inverse mass from the constraint's own state
B2J_D uint32_t joint_misc_bits(const JointState &s) { float f = s.misc.z; uint32_t u; memcpy(&u, &f, 4); return u; }
B2J_D JointBody joint_body_light(const JointState &s, uint32_t slot, int which)
{
	JointBody k;
	uint32_t bits = joint_misc_bits(s);
	k.slot = slot;
	k.type = which == 0? (bits & 3u) : ((bits >> 2) & 3u);
	k.dofs = which == 0? ((bits >> 4) & 63u) : ((bits >> 10) & 63u);
	k.inv_mass = which == 0? s.misc.x : s.misc.y;
	k.x = v3_zero(); k.q = q4_identity(); k.diag = v3_zero(); k.irot = q4_identity();
	return k;
}
B2J_D void joint_store_misc(JointState &s, const JointBody &b1, const JointBody &b2)
{
	uint32_t bits = b1.type | (b2.type << 2) | (b1.dofs << 4) | (b2.dofs << 10);
	float f; memcpy(&f, &bits, 4);
	s.misc = f4(b1.inv_mass, b2.inv_mass, f, 0.0f);
}

B2J_D V3 joint_linear_velocity(const DWorld &w, const JointBody &b) { return b.type != B2J_MOTION_STATIC? to_v3(w.linear_velocity[b.slot]) : v3_zero(); }
B2J_D V3 joint_angular_velocity(const DWorld &w, const JointBody &b) { return b.type != B2J_MOTION_STATIC? to_v3(w.angular_velocity[b.slot]) : v3_zero(); }

// ---- PointConstraintPart ------------------------------------------------------------------------------------------------------------
B2J_D void point_calculate(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2)
{
	M33 rotation1 = m33_rotation(b1.q), rotation2 = m33_rotation(b2.q);
	V3 r1 = mul(rotation1, to_v3(d.local1)), r2 = mul(rotation2, to_v3(d.local2));
	s.r1 = f4(r1); s.r2 = f4(r2);
	float summed_inv_mass;
	M33 inv_effective_mass;
	if (b1.type == B2J_MOTION_DYNAMIC)
	{
		M33 inv_i1 = inverse_inertia_for_rotation(rotation1, b1.irot, b1.diag, b1.dofs);
		summed_inv_mass = b1.inv_mass;
		M33 r1x = m33_cross_product(r1);
		M33 i1 = mul(inv_i1, r1x);
		s.i1[0] = f4(i1.c0); s.i1[1] = f4(i1.c1); s.i1[2] = f4(i1.c2);
		inv_effective_mass = mul_right_transposed(mul(r1x, inv_i1), r1x);
	}
	else
	{
		summed_inv_mass = 0.0f;
		inv_effective_mass = m33_zero();
	}
	if (b2.type == B2J_MOTION_DYNAMIC)
	{
		M33 inv_i2 = inverse_inertia_for_rotation(rotation2, b2.irot, b2.diag, b2.dofs);
		summed_inv_mass += b2.inv_mass;
		M33 r2x = m33_cross_product(r2);
		M33 i2 = mul(inv_i2, r2x);
		s.i2[0] = f4(i2.c0); s.i2[1] = f4(i2.c1); s.i2[2] = f4(i2.c2);
		inv_effective_mass = m33_add(inv_effective_mass, mul_right_transposed(mul(r2x, inv_i2), r2x));
	}
	inv_effective_mass = m33_add(inv_effective_mass, m33(v3(summed_inv_mass, 0.0f, 0.0f), v3(0.0f, summed_inv_mass, 0.0f), v3(0.0f, 0.0f, summed_inv_mass)));
	M33 eff;
	if (!m33_inversed(inv_effective_mass, eff))
	{
		// Deactivate()
		eff = m33_zero();
		s.lambda = f4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	s.eff[0] = f4(eff.c0); s.eff[1] = f4(eff.c1); s.eff[2] = f4(eff.c2);
}

B2J_D bool point_apply_velocity_step(const DWorld &w, const JointState &s, const JointBody &b1, const JointBody &b2, V3 lambda)
{
	if (lambda == v3_zero())
		return false;
	if (b1.type == B2J_MOTION_DYNAMIC)
	{
		M33 i1 = m33(to_v3(s.i1[0]), to_v3(s.i1[1]), to_v3(s.i1[2]));
		w.linear_velocity[b1.slot] = f4(lock_translation(to_v3(w.linear_velocity[b1.slot]) - b1.inv_mass * lambda, b1.dofs));
		w.angular_velocity[b1.slot] = f4(to_v3(w.angular_velocity[b1.slot]) - mul(i1, lambda));
	}
	if (b2.type == B2J_MOTION_DYNAMIC)
	{
		M33 i2 = m33(to_v3(s.i2[0]), to_v3(s.i2[1]), to_v3(s.i2[2]));
		w.linear_velocity[b2.slot] = f4(lock_translation(to_v3(w.linear_velocity[b2.slot]) + b2.inv_mass * lambda, b2.dofs));
		w.angular_velocity[b2.slot] = f4(to_v3(w.angular_velocity[b2.slot]) + mul(i2, lambda));
	}
	return true;
}

// ---- AxisConstraintPart as the distance constraint uses it (bias 0, no spring) ------------------------------------------------------
B2J_D void distance_calculate(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2)
{
	V3 wsp1 = mul(xf_rotation_translation(b1.q, b1.x), to_v3(d.local1));
	V3 wsp2 = mul(xf_rotation_translation(b2.q, b2.x), to_v3(d.local2));
	V3 delta = wsp2 - wsp1;
	float delta_len = length(delta);
	V3 normal = to_v3(s.normal);
	if (delta_len > 0.0f)
		normal = delta / delta_len;
	s.normal = f4(normal);
	V3 r1_plus_u = wsp2 - b1.x, r2 = wsp2 - b2.x;
	float min_distance = d.local1.w, max_distance = d.local2.w;
	float min_lambda = s.wsp1.w, max_lambda = s.wsp2.w;
	bool calculate = true;
	if (min_distance == max_distance) { min_lambda = -FLT_MAX; max_lambda = FLT_MAX; }
	else if (delta_len <= min_distance) { min_lambda = 0.0f; max_lambda = FLT_MAX; }
	else if (delta_len >= max_distance) { min_lambda = -FLT_MAX; max_lambda = 0.0f; }
	else calculate = false;
	float eff = 0.0f;
	if (calculate)
	{
		// CalculateInverseEffectiveMass
		float inv_effective_mass;
		if (b1.type != B2J_MOTION_STATIC)
		{
			V3 r1x = cross(r1_plus_u, normal);
			s.r1 = f4(r1x);
			if (b1.type == B2J_MOTION_DYNAMIC)
			{
				V3 i1 = multiply_ws_inverse_inertia(b1.q, b1.irot, b1.diag, b1.dofs, r1x);
				s.i1[0] = f4(i1);
				inv_effective_mass = b1.inv_mass + dot(i1, r1x);
			}
			else
				inv_effective_mass = 0.0f;
		}
		else
			inv_effective_mass = 0.0f;
		if (b2.type != B2J_MOTION_STATIC)
		{
			V3 r2x = cross(r2, normal);
			s.r2 = f4(r2x);
			if (b2.type == B2J_MOTION_DYNAMIC)
			{
				V3 i2 = multiply_ws_inverse_inertia(b2.q, b2.irot, b2.diag, b2.dofs, r2x);
				s.i2[0] = f4(i2);
				inv_effective_mass += b2.inv_mass + dot(i2, r2x);
			}
		}
		if (inv_effective_mass != 0.0f)
			eff = 1.0f / inv_effective_mass;
	}
	if (eff == 0.0f)
		s.lambda.x = 0.0f; // Deactivate()
	s.r1.w = eff;
	s.wsp1 = f4(wsp1, min_lambda);
	s.wsp2 = f4(wsp2, max_lambda);
}

B2J_D bool axis_apply_velocity_step(const DWorld &w, const JointState &s, const JointBody &b1, const JointBody &b2, V3 axis, float lambda)
{
	if (lambda == 0.0f)
		return false;
	if (b1.type == B2J_MOTION_DYNAMIC)
	{
		w.linear_velocity[b1.slot] = f4(lock_translation(to_v3(w.linear_velocity[b1.slot]) - (lambda * b1.inv_mass) * axis, b1.dofs));
		w.angular_velocity[b1.slot] = f4(to_v3(w.angular_velocity[b1.slot]) - lambda * to_v3(s.i1[0]));
	}
	if (b2.type == B2J_MOTION_DYNAMIC)
	{
		w.linear_velocity[b2.slot] = f4(lock_translation(to_v3(w.linear_velocity[b2.slot]) + (lambda * b2.inv_mass) * axis, b2.dofs));
		w.angular_velocity[b2.slot] = f4(to_v3(w.angular_velocity[b2.slot]) + lambda * to_v3(s.i2[0]));
	}
	return true;
}

// ---- HingeConstraint ------------------------------------------------------------------------------------------------------------------
// Vec4::ATan (cephes atanf) for one lane
B2J_HD float jolt_atan(float in)
{
	uint32_t bits; memcpy(&bits, &in, 4);
	uint32_t sign = bits & 0x80000000u;
	bits ^= sign;
	float x; memcpy(&x, &bits, 4);
	float y = 0.0f;
	bool greater1 = x > 0.4142135623730950f, greater2 = x > 2.414213562373095f;
	float x1 = (x - 1.0f) / (x + 1.0f), x2 = -1.0f / x;
	if (greater1) { x = x1; y = 0.25f * 3.14159265358979323846f; }
	if (greater2) { x = x2; y = 0.5f * 3.14159265358979323846f; }
	float z = x * x;
	y += (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x;
	uint32_t ybits; memcpy(&ybits, &y, 4);
	ybits ^= sign;
	memcpy(&y, &ybits, 4);
	return y;
}

B2J_HD float center_angle_around_zero(float v)
{
	const float pi = 3.14159265358979323846f;
	if (v < -pi) { do v += 2.0f * pi; while (v < -pi); }
	else if (v > pi) { do v -= 2.0f * pi; while (v > pi); }
	return v;
}

B2J_D M33 joint_inverse_inertia(const JointBody &b, const M33 &rotation) { return b.type == B2J_MOTION_DYNAMIC? inverse_inertia_for_rotation(rotation, b.irot, b.diag, b.dofs) : m33_zero(); }

// HingeRotationConstraintPart::CalculateConstraintProperties
B2J_D void hinge_rotation_calculate(JointState &s, const JointBody &b1, const M33 &rotation1, V3 axis1, const JointBody &b2, const M33 &rotation2, V3 axis2)
{
	V3 a1 = axis1, a2 = axis2;
	float d = dot(a1, a2);
	if (d <= 1.0e-3f)
	{
		V3 perp = a2 - d * a1;
		if (length_sq(perp) < 1.0e-6f)
			perp = normalized_perpendicular(a1);
		a2 = normalized(0.99f * normalized(perp) + 0.01f * a1);
	}
	V3 b2v = normalized_perpendicular(a2);
	V3 c2v = cross(a2, b2v);
	M33 inv1 = joint_inverse_inertia(b1, rotation1), inv2 = joint_inverse_inertia(b2, rotation2);
	V3 b2xa1 = cross(b2v, a1), c2xa1 = cross(c2v, a1);
	M33 summed = m33_add(inv1, inv2);
	float m00 = dot(b2xa1, mul(summed, b2xa1)), m01 = dot(b2xa1, mul(summed, c2xa1));
	float m10 = dot(c2xa1, mul(summed, b2xa1)), m11 = dot(c2xa1, mul(summed, c2xa1));
	// Matrix<2, 2>::SetInversed
	float det = m00 * m11 - m01 * m10;
	if (det == 0.0f)
	{
		s.h_eff = f4(0.0f, 0.0f, 0.0f, 0.0f);
		s.lambda2.x = 0.0f; s.lambda2.y = 0.0f;
	}
	else
		s.h_eff = f4(m11 / det, -m01 / det, -m10 / det, m00 / det);
	s.h_a1 = f4(a1, s.h_a1.w); s.h_b2 = f4(b2v, s.h_b2.w); s.h_c2 = f4(c2v, s.h_c2.w);
	s.h_b2xa1 = f4(b2xa1); s.h_c2xa1 = f4(c2xa1);
	s.h_inv1[0] = f4(inv1.c0); s.h_inv1[1] = f4(inv1.c1); s.h_inv1[2] = f4(inv1.c2);
	s.h_inv2[0] = f4(inv2.c0); s.h_inv2[1] = f4(inv2.c1); s.h_inv2[2] = f4(inv2.c2);
}

B2J_D bool hinge_rotation_apply(const DWorld &w, const JointState &s, const JointBody &b1, const JointBody &b2, float l0, float l1)
{
	if (l0 == 0.0f && l1 == 0.0f)
		return false;
	V3 impulse = to_v3(s.h_b2xa1) * l0 + to_v3(s.h_c2xa1) * l1;
	if (b1.type == B2J_MOTION_DYNAMIC)
		w.angular_velocity[b1.slot] = f4(to_v3(w.angular_velocity[b1.slot]) - mul(m33(to_v3(s.h_inv1[0]), to_v3(s.h_inv1[1]), to_v3(s.h_inv1[2])), impulse));
	if (b2.type == B2J_MOTION_DYNAMIC)
		w.angular_velocity[b2.slot] = f4(to_v3(w.angular_velocity[b2.slot]) + mul(m33(to_v3(s.h_inv2[0]), to_v3(s.h_inv2[1]), to_v3(s.h_inv2[2])), impulse));
	return true;
}

// AngleConstraintPart::CalculateConstraintProperties (bias 0, no spring): returns the effective mass (0 = Deactivate)
B2J_D float angle_calculate(const JointBody &b1, const JointBody &b2, V3 axis, F4 &out_i1, F4 &out_i2)
{
	V3 i1 = b1.type == B2J_MOTION_DYNAMIC? multiply_ws_inverse_inertia(b1.q, b1.irot, b1.diag, b1.dofs, axis) : v3_zero();
	V3 i2 = b2.type == B2J_MOTION_DYNAMIC? multiply_ws_inverse_inertia(b2.q, b2.irot, b2.diag, b2.dofs, axis) : v3_zero();
	out_i1 = f4(i1); out_i2 = f4(i2);
	float inv_effective_mass = dot(axis, i1 + i2);
	return inv_effective_mass == 0.0f? 0.0f : 1.0f / inv_effective_mass;
}

B2J_D bool angle_apply(const DWorld &w, const JointBody &b1, const JointBody &b2, F4 i1, F4 i2, float lambda)
{
	if (lambda == 0.0f)
		return false;
	if (b1.type == B2J_MOTION_DYNAMIC) w.angular_velocity[b1.slot] = f4(to_v3(w.angular_velocity[b1.slot]) - lambda * to_v3(i1));
	if (b2.type == B2J_MOTION_DYNAMIC) w.angular_velocity[b2.slot] = f4(to_v3(w.angular_velocity[b2.slot]) + lambda * to_v3(i2));
	return true;
}

// AngleConstraintPart::SolveVelocityConstraint
B2J_D void angle_solve_velocity(const DWorld &w, const JointBody &b1, const JointBody &b2, V3 axis, float eff, F4 i1, F4 i2, float &total, float min_lambda, float max_lambda)
{
	float lambda = eff * (dot(axis, joint_angular_velocity(w, b1) - joint_angular_velocity(w, b2)) - 0.0f);
	float new_lambda = clamp_(total + lambda, min_lambda, max_lambda);
	lambda = new_lambda - total;
	total = new_lambda;
	angle_apply(w, b1, b2, i1, i2, lambda);
}

B2J_HD bool hinge_has_limits(const JointDef &d) { const float pi = 3.14159265358979323846f; return d.axis1.w > -pi || d.axis2.w < pi; }
B2J_HD float hinge_smallest_angle_to_limit(const JointDef &d, float theta)
{
	float dist_to_min = center_angle_around_zero(theta - d.axis1.w), dist_to_max = center_angle_around_zero(theta - d.axis2.w);
	return fabs_(dist_to_min) < fabs_(dist_to_max)? dist_to_min : dist_to_max;
}
B2J_HD bool hinge_is_min_limit_closest(const JointDef &d, float theta)
{
	float dist_to_min = center_angle_around_zero(theta - d.axis1.w), dist_to_max = center_angle_around_zero(theta - d.axis2.w);
	return fabs_(dist_to_min) < fabs_(dist_to_max);
}

// CalculateA1AndTheta + CalculateRotationLimitsConstraintProperties
B2J_D void hinge_limits_calculate(const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2)
{
	bool has_limits = hinge_has_limits(d);
	if (has_limits || d.hinge.x > 0.0f)
	{
		Q4 diff = (b2.q * to_q4(d.inv_initial_orientation)) * q4_conj(b1.q);
		V3 a1 = rotate(b1.q, to_v3(d.axis1));
		s.h_axis = f4(a1);
		// Quat::GetRotationAngle
		s.h_a1.w = diff.w == 0.0f? 3.14159265358979323846f : 2.0f * jolt_atan(dot(q4_xyz(diff), a1) / diff.w);
	}
	float theta = s.h_a1.w;
	if (has_limits && (theta <= d.axis1.w || theta >= d.axis2.w))
		s.h_b2.w = angle_calculate(b1, b2, to_v3(s.h_axis), s.h_l1, s.h_l2);
	else
		s.h_b2.w = 0.0f;
	if (s.h_b2.w == 0.0f)
		s.lambda2.z = 0.0f; // Deactivate()
}

B2J_D void hinge_setup(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2)
{
	M33 rotation1 = m33_rotation(b1.q), rotation2 = m33_rotation(b2.q);
	point_calculate(w, d, s, b1, b2);
	hinge_rotation_calculate(s, b1, rotation1, mul(rotation1, to_v3(d.axis1)), b2, rotation2, mul(rotation2, to_v3(d.axis2)));
	hinge_limits_calculate(d, s, b1, b2);
	// CalculateMotorConstraintProperties, EMotorState::Off: friction
	if (d.hinge.x > 0.0f)
		s.h_c2.w = angle_calculate(b1, b2, to_v3(s.h_axis), s.h_m1, s.h_m2);
	else
		s.h_c2.w = 0.0f;
	if (s.h_c2.w == 0.0f)
		s.lambda2.w = 0.0f;
}

B2J_D void hinge_warm_start(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2, float ratio)
{
	s.lambda2.w *= ratio;
	angle_apply(w, b1, b2, s.h_m1, s.h_m2, s.lambda2.w);
	V3 lambda = to_v3(s.lambda) * ratio;
	s.lambda = f4(lambda);
	point_apply_velocity_step(w, s, b1, b2, lambda);
	s.lambda2.x *= ratio; s.lambda2.y *= ratio;
	hinge_rotation_apply(w, s, b1, b2, s.lambda2.x, s.lambda2.y);
	s.lambda2.z *= ratio;
	angle_apply(w, b1, b2, s.h_l1, s.h_l2, s.lambda2.z);
}

B2J_D V3 point_velocity_lambda(const DWorld &w, const JointState &s, const JointBody &b1, const JointBody &b2)
{
	V3 v1 = joint_linear_velocity(w, b1), w1 = joint_angular_velocity(w, b1), v2 = joint_linear_velocity(w, b2), w2 = joint_angular_velocity(w, b2);
	M33 eff = m33(to_v3(s.eff[0]), to_v3(s.eff[1]), to_v3(s.eff[2]));
	return mul(eff, ((v1 - cross(to_v3(s.r1), w1)) - v2) + cross(to_v3(s.r2), w2));
}

B2J_D void hinge_solve_velocity(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2, float dt)
{
	V3 axis = to_v3(s.h_axis);
	if (s.h_c2.w != 0.0f)
	{
		float max_lambda = d.hinge.x * dt;
		angle_solve_velocity(w, b1, b2, axis, s.h_c2.w, s.h_m1, s.h_m2, s.lambda2.w, -max_lambda, max_lambda);
	}
	V3 lambda = point_velocity_lambda(w, s, b1, b2);
	s.lambda = f4(to_v3(s.lambda) + lambda);
	point_apply_velocity_step(w, s, b1, b2, lambda);
	// HingeRotationConstraintPart::SolveVelocityConstraint
	V3 delta_ang = joint_angular_velocity(w, b1) - joint_angular_velocity(w, b2);
	float jv0 = dot(to_v3(s.h_b2xa1), delta_ang), jv1 = dot(to_v3(s.h_c2xa1), delta_ang);
	float l0 = (0.0f + s.h_eff.x * jv0) + s.h_eff.y * jv1, l1 = (0.0f + s.h_eff.z * jv0) + s.h_eff.w * jv1;
	s.lambda2.x += l0; s.lambda2.y += l1;
	hinge_rotation_apply(w, s, b1, b2, l0, l1);
	if (s.h_b2.w != 0.0f)
	{
		float min_lambda, max_lambda;
		if (d.axis1.w == d.axis2.w) { min_lambda = -FLT_MAX; max_lambda = FLT_MAX; }
		else if (hinge_is_min_limit_closest(d, s.h_a1.w)) { min_lambda = 0.0f; max_lambda = FLT_MAX; }
		else { min_lambda = -FLT_MAX; max_lambda = 0.0f; }
		angle_solve_velocity(w, b1, b2, axis, s.h_b2.w, s.h_l1, s.h_l2, s.lambda2.z, min_lambda, max_lambda);
	}
}

// PointConstraintPart::SolvePositionConstraint after CalculateConstraintProperties (shared by the point and the hinge constraint)
B2J_D void point_solve_position(const DWorld &w, const JointDef &d, JointState &s, JointBody &b1, JointBody &b2, float baumgarte)
{
	point_calculate(w, d, s, b1, b2);
	V3 separation = ((b2.x - b1.x) + to_v3(s.r2)) - to_v3(s.r1);
	if (separation == v3_zero())
		return;
	M33 eff = m33(to_v3(s.eff[0]), to_v3(s.eff[1]), to_v3(s.eff[2]));
	// mEffectiveMass * -inBaumgarte * separation = (Mat44 * float) * Vec3
	float nb = -baumgarte;
	V3 lambda = mul(m33(eff.c0 * nb, eff.c1 * nb, eff.c2 * nb), separation);
	if (b1.type == B2J_MOTION_DYNAMIC)
	{
		M33 i1 = m33(to_v3(s.i1[0]), to_v3(s.i1[1]), to_v3(s.i1[2]));
		b1.x -= lock_translation(b1.inv_mass * lambda, b1.dofs);
		b1.q = add_rotation_step(b1.q, mul(i1, lambda), true);
		w.position[b1.slot] = f4(b1.x); w.rotation[b1.slot] = f4(b1.q);
	}
	if (b2.type == B2J_MOTION_DYNAMIC)
	{
		M33 i2 = m33(to_v3(s.i2[0]), to_v3(s.i2[1]), to_v3(s.i2[2]));
		b2.x += lock_translation(b2.inv_mass * lambda, b2.dofs);
		b2.q = add_rotation_step(b2.q, mul(i2, lambda), false);
		w.position[b2.slot] = f4(b2.x); w.rotation[b2.slot] = f4(b2.q);
	}
}

B2J_D void hinge_solve_position(const DWorld &w, const JointDef &d, JointState &s, JointBody &b1, JointBody &b2, float baumgarte)
{
	point_solve_position(w, d, s, b1, b2, baumgarte);
	// (b1 / b2 carry the poses the point part left behind)
	M33 rotation1 = m33_rotation(b1.q), rotation2 = m33_rotation(b2.q);
	hinge_rotation_calculate(s, b1, rotation1, mul(rotation1, to_v3(d.axis1)), b2, rotation2, mul(rotation2, to_v3(d.axis2)));
	// HingeRotationConstraintPart::SolvePositionConstraint
	float c0 = dot(to_v3(s.h_a1), to_v3(s.h_b2)), c1 = dot(to_v3(s.h_a1), to_v3(s.h_c2));
	if (!(c0 == 0.0f && c1 == 0.0f))
	{
		float e0 = (0.0f + s.h_eff.x * c0) + s.h_eff.y * c1, e1 = (0.0f + s.h_eff.z * c0) + s.h_eff.w * c1;
		float l0 = -baumgarte * e0, l1 = -baumgarte * e1;
		V3 impulse = to_v3(s.h_b2xa1) * l0 + to_v3(s.h_c2xa1) * l1;
		if (b1.type == B2J_MOTION_DYNAMIC)
		{
			b1.q = add_rotation_step(b1.q, mul(m33(to_v3(s.h_inv1[0]), to_v3(s.h_inv1[1]), to_v3(s.h_inv1[2])), impulse), true);
			w.rotation[b1.slot] = f4(b1.q);
		}
		if (b2.type == B2J_MOTION_DYNAMIC)
		{
			b2.q = add_rotation_step(b2.q, mul(m33(to_v3(s.h_inv2[0]), to_v3(s.h_inv2[1]), to_v3(s.h_inv2[2])), impulse), false);
			w.rotation[b2.slot] = f4(b2.q);
		}
	}
	if (hinge_has_limits(d))
	{
		hinge_limits_calculate(d, s, b1, b2);
		if (s.h_b2.w != 0.0f)
		{
			// AngleConstraintPart::SolvePositionConstraint
			float c = hinge_smallest_angle_to_limit(d, s.h_a1.w);
			if (c != 0.0f)
			{
				float lambda = -s.h_b2.w * baumgarte * c;
				if (b1.type == B2J_MOTION_DYNAMIC) { b1.q = add_rotation_step(b1.q, lambda * to_v3(s.h_l1), true); w.rotation[b1.slot] = f4(b1.q); }
				if (b2.type == B2J_MOTION_DYNAMIC) { b2.q = add_rotation_step(b2.q, lambda * to_v3(s.h_l2), false); w.rotation[b2.slot] = f4(b2.q); }
			}
		}
	}
}

// ---- FixedConstraint: RotationEulerConstraintPart + PointConstraintPart ---------------------------------------------------------------
B2J_D void euler_calculate(JointState &s, const JointBody &b1, const M33 &rotation1, const JointBody &b2, const M33 &rotation2)
{
	M33 inv1 = joint_inverse_inertia(b1, rotation1), inv2 = joint_inverse_inertia(b2, rotation2);
	s.h_inv1[0] = f4(inv1.c0); s.h_inv1[1] = f4(inv1.c1); s.h_inv1[2] = f4(inv1.c2);
	s.h_inv2[0] = f4(inv2.c0); s.h_inv2[1] = f4(inv2.c1); s.h_inv2[2] = f4(inv2.c2);
	M33 inertia_sum = m33_add(inv1, inv2), eff;
	if (!m33_inversed(inertia_sum, eff))
	{
		// a zero column is a locked axis: identity there (any impulse is multiplied by mInvI1 / mInvI2 afterwards)
		if (inertia_sum.c0 == v3_zero()) inertia_sum.c0 = v3(1.0f, 0.0f, 0.0f);
		if (inertia_sum.c1 == v3_zero()) inertia_sum.c1 = v3(0.0f, 1.0f, 0.0f);
		if (inertia_sum.c2 == v3_zero()) inertia_sum.c2 = v3(0.0f, 0.0f, 1.0f);
		if (!m33_inversed(inertia_sum, eff))
		{
			eff = m33_zero();
			s.lambda2.x = 0.0f; s.lambda2.y = 0.0f; s.lambda2.z = 0.0f;
		}
	}
	s.f_eff[0] = f4(eff.c0); s.f_eff[1] = f4(eff.c1); s.f_eff[2] = f4(eff.c2);
}

B2J_D bool euler_apply(const DWorld &w, const JointState &s, const JointBody &b1, const JointBody &b2, V3 lambda)
{
	if (lambda == v3_zero())
		return false;
	if (b1.type == B2J_MOTION_DYNAMIC)
		w.angular_velocity[b1.slot] = f4(to_v3(w.angular_velocity[b1.slot]) - mul(m33(to_v3(s.h_inv1[0]), to_v3(s.h_inv1[1]), to_v3(s.h_inv1[2])), lambda));
	if (b2.type == B2J_MOTION_DYNAMIC)
		w.angular_velocity[b2.slot] = f4(to_v3(w.angular_velocity[b2.slot]) + mul(m33(to_v3(s.h_inv2[0]), to_v3(s.h_inv2[1]), to_v3(s.h_inv2[2])), lambda));
	return true;
}

B2J_D void fixed_setup(const DWorld &w, const JointDef &d, JointState &s, const JointBody &b1, const JointBody &b2)
{
	euler_calculate(s, b1, m33_rotation(b1.q), b2, m33_rotation(b2.q));
	point_calculate(w, d, s, b1, b2);
}

B2J_D void fixed_warm_start(const DWorld &w, JointState &s, const JointBody &b1, const JointBody &b2, float ratio)
{
	V3 rot = v3(s.lambda2.x, s.lambda2.y, s.lambda2.z) * ratio;
	s.lambda2 = f4(rot, s.lambda2.w);
	euler_apply(w, s, b1, b2, rot);
	V3 lambda = to_v3(s.lambda) * ratio;
	s.lambda = f4(lambda);
	point_apply_velocity_step(w, s, b1, b2, lambda);
}

B2J_D void fixed_solve_velocity(const DWorld &w, JointState &s, const JointBody &b1, const JointBody &b2)
{
	M33 eff = m33(to_v3(s.f_eff[0]), to_v3(s.f_eff[1]), to_v3(s.f_eff[2]));
	V3 rot = mul(eff, joint_angular_velocity(w, b1) - joint_angular_velocity(w, b2));
	s.lambda2 = f4(v3(s.lambda2.x, s.lambda2.y, s.lambda2.z) + rot, s.lambda2.w);
	euler_apply(w, s, b1, b2, rot);
	V3 lambda = point_velocity_lambda(w, s, b1, b2);
	s.lambda = f4(to_v3(s.lambda) + lambda);
	point_apply_velocity_step(w, s, b1, b2, lambda);
}

B2J_D void fixed_solve_position(const DWorld &w, const JointDef &d, JointState &s, JointBody &b1, JointBody &b2, float baumgarte)
{
	euler_calculate(s, b1, m33_rotation(b1.q), b2, m33_rotation(b2.q));
	// RotationEulerConstraintPart::SolvePositionConstraint
	Q4 diff = (b2.q * to_q4(d.inv_initial_orientation)) * q4_conj(b1.q);
	V3 error = 2.0f * q4_xyz(q4_ensure_w_positive(diff));
	if (!(error == v3_zero()))
	{
		M33 eff = m33(to_v3(s.f_eff[0]), to_v3(s.f_eff[1]), to_v3(s.f_eff[2]));
		// -inBaumgarte * mEffectiveMass * error = ((-inBaumgarte) * Mat44) * Vec3 (Mat44 * Vec3 adds the translation column (0, 0, 0, 1): + 0)
		float nb = -baumgarte;
		V3 lambda = mul(m33(nb * eff.c0, nb * eff.c1, nb * eff.c2), error);
		if (b1.type == B2J_MOTION_DYNAMIC)
		{
			b1.q = add_rotation_step(b1.q, mul(m33(to_v3(s.h_inv1[0]), to_v3(s.h_inv1[1]), to_v3(s.h_inv1[2])), lambda), true);
			w.rotation[b1.slot] = f4(b1.q);
		}
		if (b2.type == B2J_MOTION_DYNAMIC)
		{
			b2.q = add_rotation_step(b2.q, mul(m33(to_v3(s.h_inv2[0]), to_v3(s.h_inv2[1]), to_v3(s.h_inv2[2])), lambda), false);
			w.rotation[b2.slot] = f4(b2.q);
		}
	}
	point_solve_position(w, d, s, b1, b2, baumgarte);
}

// ---- the four solver entry points of a constraint -----------------------------------------------------------------------------------
B2J_D void joint_setup_velocity(const DWorld &w, const JointDef &d, JointState &s)
{
	JointBody b1 = joint_body(w, d.b1), b2 = joint_body(w, d.b2);
	joint_store_misc(s, b1, b2);
	if (d.type == JOINT_POINT) point_calculate(w, d, s, b1, b2);
	else if (d.type == JOINT_DISTANCE) distance_calculate(w, d, s, b1, b2);
	else if (d.type == JOINT_HINGE) hinge_setup(w, d, s, b1, b2);
	else fixed_setup(w, d, s, b1, b2);
}

// (velocity passes: `type` and the body slots come from the item header, the bodies' motion types / DOFs / inverse masses from the state)
B2J_D void joint_warm_start(const DWorld &w, const JointDef &d, JointState &s, uint32_t type, uint32_t slot1, uint32_t slot2, float ratio)
{
	JointBody b1 = joint_body_light(s, slot1, 0), b2 = joint_body_light(s, slot2, 1);
	if (type == JOINT_POINT)
	{
		V3 lambda = to_v3(s.lambda) * ratio;
		s.lambda = f4(lambda);
		point_apply_velocity_step(w, s, b1, b2, lambda);
	}
	else if (type == JOINT_DISTANCE)
	{
		s.lambda.x *= ratio;
		axis_apply_velocity_step(w, s, b1, b2, to_v3(s.normal), s.lambda.x);
	}
	else if (type == JOINT_HINGE)
		hinge_warm_start(w, d, s, b1, b2, ratio);
	else
		fixed_warm_start(w, s, b1, b2, ratio);
}

B2J_D void joint_solve_velocity(const DWorld &w, const JointDef &d, JointState &s, uint32_t type, uint32_t slot1, uint32_t slot2, float dt)
{
	JointBody b1 = joint_body_light(s, slot1, 0), b2 = joint_body_light(s, slot2, 1);
	if (type == JOINT_HINGE)
	{
		hinge_solve_velocity(w, d, s, b1, b2, dt);
		return;
	}
	if (type == JOINT_FIXED)
	{
		fixed_solve_velocity(w, s, b1, b2);
		return;
	}
	V3 v1 = joint_linear_velocity(w, b1), w1 = joint_angular_velocity(w, b1), v2 = joint_linear_velocity(w, b2), w2 = joint_angular_velocity(w, b2);
	if (type == JOINT_POINT)
	{
		V3 lambda = point_velocity_lambda(w, s, b1, b2);
		s.lambda = f4(to_v3(s.lambda) + lambda);
		point_apply_velocity_step(w, s, b1, b2, lambda);
	}
	else
	{
		float eff = s.r1.w;
		if (eff == 0.0f)
			return; // !mAxisConstraint.IsActive()
		V3 axis = to_v3(s.normal);
		float jv;
		if (b1.type != B2J_MOTION_STATIC)
			jv = b2.type != B2J_MOTION_STATIC? dot(axis, v1 - v2) : dot(axis, v1);
		else
			jv = dot(axis, -v2);
		if (b1.type != B2J_MOTION_STATIC) jv += dot(to_v3(s.r1), w1);
		if (b2.type != B2J_MOTION_STATIC) jv -= dot(to_v3(s.r2), w2);
		float total = s.lambda.x;
		float lambda = eff * (jv - 0.0f);
		float new_lambda = clamp_(total + lambda, s.wsp1.w, s.wsp2.w);
		lambda = new_lambda - total;
		s.lambda.x = new_lambda;
		axis_apply_velocity_step(w, s, b1, b2, axis, lambda);
	}
}

B2J_D void joint_store_pose(const DWorld &w, const JointBody &b) { w.position[b.slot] = f4(b.x); w.rotation[b.slot] = f4(b.q); }

B2J_D void joint_solve_position(const DWorld &w, const JointDef &d, JointState &s, float baumgarte)
{
	JointBody b1 = joint_body(w, d.b1), b2 = joint_body(w, d.b2);
	if (d.type == JOINT_POINT)
		point_solve_position(w, d, s, b1, b2, baumgarte);
	else if (d.type == JOINT_HINGE)
		hinge_solve_position(w, d, s, b1, b2, baumgarte);
	else if (d.type == JOINT_FIXED)
		fixed_solve_position(w, d, s, b1, b2, baumgarte);
	else
	{
		// (the distance of the points as the LAST CalculateConstraintProperties saw them)
		float distance = dot(to_v3(s.wsp2) - to_v3(s.wsp1), to_v3(s.normal));
		float min_distance = d.local1.w, max_distance = d.local2.w;
		float position_error = 0.0f;
		if (distance < min_distance) position_error = distance - min_distance;
		else if (distance > max_distance) position_error = distance - max_distance;
		if (position_error == 0.0f)
			return;
		distance_calculate(w, d, s, b1, b2);
		// AxisConstraintPart::SolvePositionConstraint
		float eff = s.r1.w;
		float lambda = -eff * baumgarte * position_error;
		V3 axis = to_v3(s.normal);
		if (b1.type == B2J_MOTION_DYNAMIC)
		{
			b1.x -= lock_translation((lambda * b1.inv_mass) * axis, b1.dofs);
			b1.q = add_rotation_step(b1.q, lambda * to_v3(s.i1[0]), true);
			joint_store_pose(w, b1);
		}
		if (b2.type == B2J_MOTION_DYNAMIC)
		{
			b2.x += lock_translation((lambda * b2.inv_mass) * axis, b2.dofs);
			b2.q = add_rotation_step(b2.q, lambda * to_v3(s.i2[0]), false);
			joint_store_pose(w, b2);
		}
	}
}

// ---- kernels ------------------------------------------------------------------------------------------------------------------------

// JobDetermineActiveConstraints + the activations of ConstraintManager::sBuildIslands: one thread per constraint (by constraint index)
struct KJointActive
{
	DWorld w; NarrowCtx c; JointCtx j;
	B2J_D void operator()(uint32_t i) const
	{
		const JointDef &d = j.defs[i];
		uint32_t type1 = w.info[d.b1].motion_type, type2 = w.info[d.b2].motion_type;
		bool active1 = w.active_index[d.b1] != B2J_INACTIVE_INDEX, active2 = w.active_index[d.b2] != B2J_INACTIVE_INDEX;
		bool active = (d.flags & JOINT_ENABLED) != 0 && (active1 || active2) && (type1 == B2J_MOTION_DYNAMIC || type2 == B2J_MOTION_DYNAMIC);
		j.active_flag[i] = active? 1u : 0u;
		if (!active)
			return;
		atomic_add(&w.counters->num_active_joints, 1u);
		if (type1 == B2J_MOTION_DYNAMIC && !active1) { atomic_min(&j.wake_key[d.b1], 2 * i); wake_body(w, c, d.b1); }
		if (type2 == B2J_MOTION_DYNAMIC && !active2) { atomic_min(&j.wake_key[d.b2], 2 * i + 1); wake_body(w, c, d.b2); }
	}
};

struct KJointWakeKeys
{
	NarrowCtx c; JointCtx j; uint32_t *keys;
	B2J_D void operator()(uint32_t k) const { uint32_t b = c.woken_list[k]; keys[k] = j.wake_key[b]; j.wake_key[b] = 0xffffffffu; }
};

struct KJointClearWoken
{
	DWorld w;
	B2J_D void operator()(uint32_t) const { w.counters->num_woken = 0; }
};

// the active constraints in (priority, constraint index) order: flags in that order -> scan -> compaction
struct KJointOrderFlags
{
	JointCtx j;
	B2J_D void operator()(uint32_t k) const { j.order_flag[k] = j.active_flag[j.order[k]]; }
};

struct KJointCompact
{
	JointCtx j;
	B2J_D void operator()(uint32_t k) const { if (j.order_flag[k]) j.active_joints[j.order_scan[k]] = j.order[k]; }
};

// Constraint::SetupVelocityConstraint for the active constraints
struct KJointSetup
{
	DWorld w; JointCtx j;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t i = j.active_joints[k];
		joint_setup_velocity(w, j.defs[i], j.state[i]);
	}
};

// the active joints as constraint sources behind the step's contact constraints: src[first + k]
struct KJointAppend
{
	DWorld w; JointCtx j; ConstraintSrc *src; uint32_t first, count;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t i = j.active_joints[k];
		const JointDef &d = j.defs[i];
		ConstraintSrc s;
		s.sort_key = 0; s.manifold = SRC_JOINT | i; s.b1 = d.b1; s.b2 = d.b2;
		src[first + k] = s;
	}
};

// sorted position -> source for the joints (they precede the contacts), and the identity values the phase sort starts from
struct KJointItemOrder
{
	SolveCtx s; uint32_t *vals; uint32_t num_contacts, num_joints;
	B2J_D void operator()(uint32_t i) const
	{
		if (i < num_joints) s.order[i] = num_contacts + i;
		vals[i] = i;
	}
};

struct KJointWarmStart
{
	DWorld w; Constraints c; JointCtx j; uint32_t begin; float ratio;
	B2J_D void operator()(uint32_t k) const
	{
		ConstraintHeader hdr = c.hdr[begin + k];
		if (!(hdr.meta & META_JOINT))
			return;
		// (the state is read and written in place: only the members the constraint's type touches move)
		joint_warm_start(w, j.defs[hdr.manifold], j.state[hdr.manifold], hdr.meta & 7u, hdr.b1, hdr.b2, ratio);
	}
};

struct KJointSolveVelocity
{
	DWorld w; Constraints c; JointCtx j; uint32_t begin, iteration; float dt;
	B2J_D void operator()(uint32_t k) const
	{
		ConstraintHeader hdr = c.hdr[begin + k];
		if (!(hdr.meta & META_JOINT) || iteration >= ((hdr.meta >> 8) & 0xff))
			return;
		joint_solve_velocity(w, j.defs[hdr.manifold], j.state[hdr.manifold], hdr.meta & 7u, hdr.b1, hdr.b2, dt);
	}
};

struct KJointSolvePosition
{
	DWorld w; Constraints c; JointCtx j; uint32_t begin, iteration;
	B2J_D void operator()(uint32_t k) const
	{
		ConstraintHeader hdr = c.hdr[begin + k];
		if (!(hdr.meta & META_JOINT) || iteration >= ((hdr.meta >> 16) & 0xff))
			return;
		joint_solve_position(w, j.defs[hdr.manifold], j.state[hdr.manifold], w.settings.baumgarte);
	}
};

// Both kinds of items of a phase in ONE launch (they share no dynamic body, so the order inside the phase is free): the per phase form of
// worlds with non contact constraints. Two launches per phase and pass cost twice the launch latency such thin phases are bound by.
struct KMixedWarmStart
{
	KJointWarmStart joints; KWarmStart contacts;
	B2J_D void operator()(uint32_t k) const { joints(k); contacts(k); }
};
struct KMixedSolveVelocity
{
	KJointSolveVelocity joints; KSolveVelocity contacts;
	B2J_D void operator()(uint32_t k) const { joints(k); contacts(k); }
};
struct KMixedSolvePosition
{
	KJointSolvePosition joints; KSolvePosition contacts;
	B2J_D void operator()(uint32_t k) const { joints(k); contacts(k); }
};

#if !defined(B2J_HOSTSIM) && defined(__CUDACC__)
// Worlds with non contact constraints: the whole velocity solve (warm start + all iterations) and the whole position solve as ONE
// cooperative launch each, phases separated by grid barriers, both kinds of items in the same pass over a phase (its items share no
// dynamic body). Phase offsets and iteration counts are read on the device: no host round trip before the solve. Articulated worlds
// have many thin phases (dependency depth of chains, the colours of small large-islands): per phase launches -- two per phase and
// pass -- leave them launch latency bound (measured: 949 launches = 3.9 ms per step for 256 worlds of the joints scene).
struct KSolveVelocityJoints { }; // (profiling categories)
struct KSolvePositionJoints { };
__global__ void __launch_bounds__(128) solve_velocity_joints_kernel(const DWorld w, const SolveCtx s, const JointCtx j, float warm_start_ratio, float dt)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = w.counters->max_velocity_steps;
	const uint32_t *off = s.phase_count;
	KWarmStart ws; ws.w = w; ws.c = s.con; ws.begin = 0; ws.ratio = warm_start_ratio;
	KJointWarmStart jw; jw.w = w; jw.c = s.con; jw.j = j; jw.begin = 0; jw.ratio = warm_start_ratio;
	for (uint32_t p = 0; p < np; ++p)
	{
		if (off[p] == off[p + 1])
			continue;
		for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) { jw(k); ws(k); }
		grid.sync();
	}
	KSolveVelocity sv; sv.w = w; sv.c = s.con; sv.begin = 0; sv.prefetch = 0;
	KJointSolveVelocity js; js.w = w; js.c = s.con; js.j = j; js.begin = 0; js.dt = dt;
	for (uint32_t it = 0; it < steps; ++it)
	{
		sv.iteration = it; js.iteration = it;
		for (uint32_t p = 0; p < np; ++p)
		{
			if (off[p] == off[p + 1])
				continue;
			for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) { js(k); sv(k); }
			grid.sync();
		}
	}
}

__global__ void __launch_bounds__(128) solve_position_joints_kernel(const DWorld w, const SolveCtx s, const JointCtx j)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
	const uint32_t np = w.counters->num_phases;
	const uint32_t steps = w.counters->max_position_steps;
	const uint32_t *off = s.phase_count;
	KSolvePosition sp; sp.w = w; sp.c = s.con; sp.begin = 0;
	KJointSolvePosition jp; jp.w = w; jp.c = s.con; jp.j = j; jp.begin = 0;
	for (uint32_t it = 0; it < steps; ++it)
	{
		sp.iteration = it; jp.iteration = it;
		for (uint32_t p = 0; p < np; ++p)
		{
			if (off[p] == off[p + 1])
				continue;
			for (uint32_t k = off[p] + tid; k < off[p + 1]; k += nt) { jp(k); sp(k); }
			grid.sync();
		}
	}
}
#endif

} // namespace b2j
