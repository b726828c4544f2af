// jolt_b200_facade.h -- C++ host side above the C ABI: the reference's public surface for the hot path.
//
// Mirrors (names, argument meaning, error behaviour) of
//   PhysicsSystem      Jolt/Physics/PhysicsSystem.h:29-396      Init / Update / SetGravity / SetPhysicsSettings / listeners / GetBodyInterface
//   BodyInterface      Jolt/Physics/Body/BodyInterface.h:39-313  CreateBody / AddBody / CreateAndAddBody / RemoveBody / Activate / getters / setters / AddForce
//   ContactListener    Jolt/Physics/Collision/ContactListener.h:78-141 (observer semantics: callbacks are replayed after the step)
//   BodyActivationListener  Jolt/Physics/Body/BodyActivationListener.h:13-26
//   BodyCreationSettings    Jolt/Physics/Body/BodyCreationSettings.h (same defaults), MassProperties / MotionProperties::SetMassProperties
// in namespace JPH_B200 so that `namespace JPH = JPH_B200;` makes reference-style code compile against it.
// Header only; links against libjolt_b200.so (include/jolt_b200.h). No exceptions, no RTTI needed.
//
// Documented deviations: listeners cannot alter the step they are reported in (OnContactValidate is not called,
// ContactSettings edits are ignored); virtual layer filters are sampled into tables at Init; JobSystem / TempAllocator
// arguments of Update are accepted and ignored (the GPU owns the step).
#pragma once

#include "jolt_b200.h"

#include <cmath>
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <memory>
#include <vector>

namespace JPH_B200 {

using uint = unsigned int;
using uint8 = uint8_t;
using uint16 = uint16_t;
using uint32 = uint32_t;
using uint64 = uint64_t;

static constexpr float JPH_PI = 3.14159265358979323846f;

struct Vec3
{
	float x = 0, y = 0, z = 0;
	Vec3() = default;
	Vec3(float inX, float inY, float inZ) : x(inX), y(inY), z(inZ) { }
	static Vec3 sZero() { return Vec3(0, 0, 0); }
	static Vec3 sReplicate(float v) { return Vec3(v, v, v); }
	float GetX() const { return x; } float GetY() const { return y; } float GetZ() const { return z; }
	float operator[](int i) const { return i == 0? x : (i == 1? y : z); }
	Vec3 operator+(const Vec3 &o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
	Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
	Vec3 operator-() const { return Vec3(0.0f - x, 0.0f - y, 0.0f - z); }
	Vec3 operator*(float s) const { return Vec3(x * s, y * s, z * s); }
	Vec3 operator*(const Vec3 &o) const { return Vec3(x * o.x, y * o.y, z * o.z); }
	Vec3 Cross(const Vec3 &b) const { return Vec3(y * b.z - z * b.y, z * b.x - x * b.z, x * b.y - y * b.x); }
	float Dot(const Vec3 &b) const { return (x * b.x + y * b.y) + (z * b.z + 0.0f); }
	float LengthSq() const { return Dot(*this); }
	float Length() const { return std::sqrt(LengthSq()); }
	Vec3 Normalized() const { float l = Length(); return Vec3(x / l, y / l, z / l); } // *this / Length()
	bool operator==(const Vec3 &o) const { return x == o.x && y == o.y && z == o.z; }
	static Vec3 sAxisX() { return Vec3(1, 0, 0); } static Vec3 sAxisY() { return Vec3(0, 1, 0); } static Vec3 sAxisZ() { return Vec3(0, 0, 1); }
};
using RVec3 = Vec3;
inline Vec3 operator*(float s, const Vec3 &v) { return Vec3(s * v.x, s * v.y, s * v.z); }

struct Quat
{
	float x = 0, y = 0, z = 0, w = 1;
	Quat() = default;
	Quat(float inX, float inY, float inZ, float inW) : x(inX), y(inY), z(inZ), w(inW) { }
	static Quat sIdentity() { return Quat(0, 0, 0, 1); }
	float GetX() const { return x; } float GetY() const { return y; } float GetZ() const { return z; } float GetW() const { return w; }
	Quat Normalized() const { float l = std::sqrt((x * x + y * y) + (z * z + w * w)); return Quat(x / l, y / l, z / l, w / l); }
	// Quat::operator*(Quat) (Jolt/Math/Quat.inl, scalar path), same operation order as the device code (b2j_math.h)
	Quat operator*(const Quat &r) const
	{
		float a = x, b = y, c = z, d = w;
		return Quat((a * r.w + b * r.z) + (d * r.x - c * r.y), (b * r.w + c * r.x) + (d * r.y - a * r.z), (c * r.w + a * r.y) + (d * r.z - b * r.x), -(a * r.x + b * r.y) + (d * r.w - c * r.z));
	}
	// Quat::operator*(Vec3) (Jolt/Math/Quat.inl:383-405)
	Vec3 operator*(const Vec3 &p) const
	{
		Vec3 c(p.z * y - z * p.y, p.x * z - x * p.z, p.y * x - y * p.x);
		Vec3 cc(c.z * y - z * c.y, c.x * z - x * c.z, c.y * x - y * c.x);
		Vec3 v = Vec3(w * c.x, w * c.y, w * c.z) + cc;
		return p + (v + v);
	}
};

using ObjectLayer = uint16;
struct BroadPhaseLayer
{
	using Type = uint8;
	constexpr BroadPhaseLayer() = default;
	explicit constexpr BroadPhaseLayer(Type inValue) : mValue(inValue) { }
	explicit constexpr operator Type() const { return mValue; }
	constexpr bool operator==(const BroadPhaseLayer &o) const { return mValue == o.mValue; }
	Type mValue = 0xff;
};

enum class EMotionType : uint8 { Static = 0, Kinematic = 1, Dynamic = 2 };
enum class EMotionQuality : uint8 { Discrete = 0, LinearCast = 1 };
enum class EActivation { Activate, DontActivate };
enum class EAllowedDOFs : uint8 { None = 0, All = 0x3f, TranslationX = 1, TranslationY = 2, TranslationZ = 4, RotationX = 8, RotationY = 16, RotationZ = 32, Plane2D = 1 | 2 | 32 };
inline constexpr EAllowedDOFs operator|(EAllowedDOFs a, EAllowedDOFs b) { return EAllowedDOFs(uint8(a) | uint8(b)); }
inline constexpr EAllowedDOFs operator&(EAllowedDOFs a, EAllowedDOFs b) { return EAllowedDOFs(uint8(a) & uint8(b)); }
enum class EPhysicsUpdateError : uint32 { None = 0, ManifoldCacheFull = 1, BodyPairCacheFull = 2, ContactConstraintsFull = 4 };
inline EPhysicsUpdateError operator|(EPhysicsUpdateError a, EPhysicsUpdateError b) { return EPhysicsUpdateError(uint32(a) | uint32(b)); }
enum class EOverrideMassProperties : uint8 { CalculateMassAndInertia, CalculateInertia, MassAndInertiaProvided };
enum class EShapeSubType : uint8 { Sphere = B2J_SHAPE_SPHERE, Box = B2J_SHAPE_BOX, Capsule = B2J_SHAPE_CAPSULE, ConvexHull = B2J_SHAPE_CONVEX_HULL, Mesh = B2J_SHAPE_MESH, Cylinder = B2J_SHAPE_CYLINDER, StaticCompound = B2J_SHAPE_COMPOUND, RotatedTranslated = 64, Scaled = 65 };

class BroadPhaseLayerInterface { public: virtual ~BroadPhaseLayerInterface() = default; virtual uint GetNumBroadPhaseLayers() const = 0; virtual BroadPhaseLayer GetBroadPhaseLayer(ObjectLayer inLayer) const = 0; };
class ObjectVsBroadPhaseLayerFilter { public: virtual ~ObjectVsBroadPhaseLayerFilter() = default; virtual bool ShouldCollide(ObjectLayer, BroadPhaseLayer) const { return true; } };
class ObjectLayerPairFilter { public: virtual ~ObjectLayerPairFilter() = default; virtual bool ShouldCollide(ObjectLayer, ObjectLayer) const { return true; } };

// Accepted by Update for signature compatibility; the step runs on the GPU
class TempAllocator { public: virtual ~TempAllocator() = default; };
class TempAllocatorImpl : public TempAllocator { public: explicit TempAllocatorImpl(size_t) { } };
class JobSystem { public: virtual ~JobSystem() = default; };
class JobSystemThreadPool : public JobSystem { public: JobSystemThreadPool(uint = 0, uint = 0, int = -1) { } };

struct BodyID
{
	static constexpr uint32 cInvalidBodyID = 0xffffffff;
	BodyID() = default;
	explicit BodyID(uint32 inID) : mID(inID) { }
	BodyID(uint32 inIndex, uint8 inSequence) : mID((uint32(inSequence) << 23) | inIndex) { }
	uint32 GetIndex() const { return mID & 0x7fffff; }
	uint8 GetSequenceNumber() const { return uint8(mID >> 23); }
	uint32 GetIndexAndSequenceNumber() const { return mID; }
	bool IsInvalid() const { return mID == cInvalidBodyID; }
	bool operator==(const BodyID &o) const { return mID == o.mID; }
	bool operator!=(const BodyID &o) const { return mID != o.mID; }
	bool operator<(const BodyID &o) const { return mID < o.mID; }
	uint32 mID = cInvalidBodyID;
};

struct SubShapeID { uint32 mValue = 0xffffffff; uint32 GetValue() const { return mValue; } };
struct SubShapeIDPair
{
	BodyID mBody1ID; SubShapeID mSubShapeID1; BodyID mBody2ID; SubShapeID mSubShapeID2;
	const BodyID &GetBody1ID() const { return mBody1ID; } const BodyID &GetBody2ID() const { return mBody2ID; }
	const SubShapeID &GetSubShapeID1() const { return mSubShapeID1; } const SubShapeID &GetSubShapeID2() const { return mSubShapeID2; }
};

// A rotation + translation with Mat44's arithmetic (column major 3x3 + translation; Mat44.inl operator*, sRotationTranslation)
struct Mat44RT
{
	Vec3 c0 = Vec3(1, 0, 0), c1 = Vec3(0, 1, 0), c2 = Vec3(0, 0, 1), t = Vec3(0, 0, 0);
	static Mat44RT sRotationTranslation(const Quat &q, const Vec3 &inT)
	{
		Mat44RT m;
		float x = q.x, y = q.y, z = q.z, w = q.w;
		float tx = x + x, ty = y + y, tz = z + z;
		float xx = tx * x, yy = ty * y, zz = tz * z, xy = tx * y, xz = tx * z, xw = tx * w, yz = ty * z, yw = ty * w, zw = tz * w;
		m.c0 = Vec3((1.0f - yy) - zz, xy + zw, xz - yw); m.c1 = Vec3(xy - zw, (1.0f - zz) - xx, yz + xw); m.c2 = Vec3(xz + yw, yz - xw, (1.0f - xx) - yy);
		m.t = inT;
		return m;
	}
	static Mat44RT sRotation(const Quat &q) { return sRotationTranslation(q, Vec3::sZero()); }
	// Mat44::sInverseRotationTranslation (Mat44.inl:206-211): rotation by the conjugate, translation = -(R^-1 t)
	static Mat44RT sInverseRotationTranslation(const Quat &q, const Vec3 &inT)
	{
		Mat44RT m = sRotation(Quat(-q.x, -q.y, -q.z, q.w));
		m.t = -((m.c0 * inT.x + m.c1 * inT.y) + m.c2 * inT.z);
		return m;
	}
	Vec3 operator*(const Vec3 &v) const { return ((c0 * v.x + c1 * v.y) + c2 * v.z) + t; } // Mat44 * Vec3 (Mat44.inl:386-391)
	Vec3 Multiply3x3(const Vec3 &v) const { return (c0 * v.x + c1 * v.y) + c2 * v.z; }
	inline Quat GetQuaternion() const;
	Vec3 GetTranslation() const { return t; }
	Vec3 GetColumn3(int i) const { return i == 0? c0 : (i == 1? c1 : c2); }
	Vec3 GetAxisX() const { return c0; } Vec3 GetAxisY() const { return c1; } Vec3 GetAxisZ() const { return c2; }
	Mat44RT operator*(const Mat44RT &b) const
	{
		Mat44RT r;
		r.c0 = ((c0 * b.c0.x + c1 * b.c0.y) + c2 * b.c0.z) + t * 0.0f;
		r.c1 = ((c0 * b.c1.x + c1 * b.c1.y) + c2 * b.c1.z) + t * 0.0f;
		r.c2 = ((c0 * b.c2.x + c1 * b.c2.y) + c2 * b.c2.z) + t * 0.0f;
		r.t = ((c0 * b.t.x + c1 * b.t.y) + c2 * b.t.z) + t * 1.0f;
		return r;
	}
};

// Mat44::GetQuaternion (Mat44.inl:998-1053)
inline Quat Mat44RT::GetQuaternion() const
{
	float tr = c0.x + c1.y + c2.z;
	if (tr >= 0.0f)
	{
		float s = std::sqrt(tr + 1.0f), is = 0.5f / s;
		return Quat((c1.z - c2.y) * is, (c2.x - c0.z) * is, (c0.y - c1.x) * is, 0.5f * s);
	}
	int i = 0;
	if (c1.y > c0.x) i = 1;
	float diag[3] = { c0.x, c1.y, c2.z };
	if (c2.z > diag[i]) i = 2;
	if (i == 0)
	{
		float s = std::sqrt(c0.x - (c1.y + c2.z) + 1.0f), is = 0.5f / s;
		return Quat(0.5f * s, (c1.x + c0.y) * is, (c0.z + c2.x) * is, (c1.z - c2.y) * is);
	}
	if (i == 1)
	{
		float s = std::sqrt(c1.y - (c2.z + c0.x) + 1.0f), is = 0.5f / s;
		return Quat((c1.x + c0.y) * is, 0.5f * s, (c2.y + c1.z) * is, (c2.x - c0.z) * is);
	}
	float s = std::sqrt(c2.z - (c0.x + c1.y) + 1.0f), is = 0.5f / s;
	return Quat((c0.z + c2.x) * is, (c2.y + c1.z) * is, 0.5f * s, (c0.y - c1.x) * is);
}

using Mat44 = Mat44RT;
using RMat44 = Mat44RT;

// AABox (Jolt/Geometry/AABox.h): the operations shape bounds need
struct AABox
{
	Vec3 mMin = Vec3(FLT_MAX, FLT_MAX, FLT_MAX), mMax = Vec3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
	AABox() = default;
	AABox(const Vec3 &inMin, const Vec3 &inMax) : mMin(inMin), mMax(inMax) { }
	static Vec3 sMin(const Vec3 &a, const Vec3 &b) { return Vec3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
	static Vec3 sMax(const Vec3 &a, const Vec3 &b) { return Vec3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }
	void Encapsulate(const AABox &o) { mMin = sMin(mMin, o.mMin); mMax = sMax(mMax, o.mMax); }
	Vec3 GetCenter() const { return 0.5f * (mMin + mMax); }
	AABox Scaled(const Vec3 &inScale) const { Vec3 a = mMin * inScale, b = mMax * inScale; return AABox(sMin(a, b), sMax(a, b)); }
	AABox Transformed(const Mat44RT &m) const // AABox.h:193-213
	{
		Vec3 new_min = m.t, new_max = m.t;
		const Vec3 *cols[3] = { &m.c0, &m.c1, &m.c2 };
		for (int c = 0; c < 3; ++c)
		{
			Vec3 a = *cols[c] * mMin[c], b = *cols[c] * mMax[c];
			new_min = new_min + sMin(a, b);
			new_max = new_max + sMax(a, b);
		}
		return AABox(new_min, new_max);
	}
};

// MassProperties (Jolt/Physics/Body/MassProperties.h): mass + 3x3 inertia (column major)
struct MassProperties
{
	float mMass = 0.0f;
	float mInertia[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } }; // [column][row]

	void SetDiagonal(float a, float b, float c) { memset(mInertia, 0, sizeof(mInertia)); mInertia[0][0] = a; mInertia[1][1] = b; mInertia[2][2] = c; }
	// MassProperties::SetMassAndInertiaOfSolidBox (MassProperties.cpp:66-75)
	void SetMassAndInertiaOfSolidBox(const Vec3 &inBoxSize, float inDensity)
	{
		mMass = inBoxSize.x * inBoxSize.y * inBoxSize.z * inDensity;
		Vec3 size_sq = inBoxSize * inBoxSize;
		float s = mMass / 12.0f;
		SetDiagonal((size_sq.y + size_sq.z) * s, (size_sq.x + size_sq.z) * s, (size_sq.x + size_sq.y) * s);
	}
	// MassProperties::Scale (MassProperties.cpp:103-155): what ScaledShape::GetMassProperties applies to the inner shape's properties
	void Scale(const Vec3 &inScale)
	{
		Vec3 diagonal(mInertia[0][0], mInertia[1][1], mInertia[2][2]);
		Vec3 xyz_sq = Vec3::sReplicate(Vec3::sReplicate(0.5f).Dot(diagonal)) - diagonal;
		Vec3 xyz_scaled_sq = inScale * inScale * xyz_sq;
		float i_xx = xyz_scaled_sq.y + xyz_scaled_sq.z, i_yy = xyz_scaled_sq.x + xyz_scaled_sq.z, i_zz = xyz_scaled_sq.x + xyz_scaled_sq.y;
		float i_xy = inScale.x * inScale.y * mInertia[1][0], i_xz = inScale.x * inScale.z * mInertia[2][0], i_yz = inScale.y * inScale.z * mInertia[2][1]; // mInertia(r, c) = [c][r]
		mInertia[0][0] = i_xx; mInertia[1][0] = i_xy; mInertia[0][1] = i_xy; mInertia[1][1] = i_yy;
		mInertia[2][0] = i_xz; mInertia[0][2] = i_xz; mInertia[2][1] = i_yz; mInertia[1][2] = i_yz; mInertia[2][2] = i_zz;
		float mass_scale = std::fabs(inScale.x * inScale.y * inScale.z);
		mMass *= mass_scale;
		for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mInertia[c][r] *= mass_scale;
	}
	// MassProperties::Rotate with Mat44::sRotation(inRotation) (MassProperties.cpp:157-160): R.Multiply3x3(I).Multiply3x3RightTransposed(R)
	void Rotate(const Quat &inRotation)
	{
		// Mat44::sRotation(Quat) (Mat44.inl), columns
		float x = inRotation.x, y = inRotation.y, z = inRotation.z, w = inRotation.w;
		float tx = x + x, ty = y + y, tz = z + z;
		float xx = tx * x, yy = ty * y, zz = tz * z, xy = tx * y, xz = tx * z, xw = tx * w, yz = ty * z, yw = ty * w, zw = tz * w;
		Vec3 rc[3] = { Vec3((1.0f - yy) - zz, xy + zw, xz - yw), Vec3(xy - zw, (1.0f - zz) - xx, yz + xw), Vec3(xz + yw, yz - xw, (1.0f - xx) - yy) };
		Vec3 ic[3], tc[3];
		for (int c = 0; c < 3; ++c) ic[c] = Vec3(mInertia[c][0], mInertia[c][1], mInertia[c][2]);
		// Multiply3x3: column i = (R.c0 * I[i].x + R.c1 * I[i].y) + R.c2 * I[i].z
		for (int i = 0; i < 3; ++i) tc[i] = (rc[0] * ic[i].x + rc[1] * ic[i].y) + rc[2] * ic[i].z;
		// Multiply3x3RightTransposed: column j = (T.c0 * R.c0[j] + T.c1 * R.c1[j]) + T.c2 * R.c2[j]
		for (int j = 0; j < 3; ++j)
		{
			Vec3 col = (tc[0] * rc[0][j] + tc[1] * rc[1][j]) + tc[2] * rc[2][j];
			mInertia[j][0] = col.x; mInertia[j][1] = col.y; mInertia[j][2] = col.z;
		}
	}
	// MassProperties::Translate (MassProperties.cpp:162-171): parallel axis theorem, I += m * (|t|^2 E - t t^T)
	void Translate(const Vec3 &inTranslation)
	{
		float d = inTranslation.Dot(inTranslation);
		for (int c = 0; c < 3; ++c)
			for (int r = 0; r < 3; ++r)
			{
				float scale_term = c == r? d : 0.0f;               // Mat44::sScale(t.t)
				float outer = inTranslation[r] * inTranslation[c];   // Mat44::sOuterProduct(t, t)(r, c) = t[r] * t[c]
				mInertia[c][r] += mMass * (scale_term - outer);
			}
	}
	// MassProperties::ScaleToMass
	void ScaleToMass(float inMass)
	{
		if (mMass > 0.0f)
		{
			float mass_scale = inMass / mMass;
			mMass = inMass;
			for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mInertia[c][r] *= mass_scale;
		}
		else
			mMass = inMass;
	}

	// EigenValueSymmetric (Jolt/Math/EigenValueSymmetric.h) + sort + handedness (MassProperties.cpp:23-57) + Mat44::GetQuaternion
	bool DecomposePrincipalMomentsOfInertia(Quat &outRotation, Vec3 &outDiagonal) const
	{
		const int n = 3;
		float a[3][3]; // a(row, col) = a[row][col]
		for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = mInertia[c][r];
		float vec[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } }; // vec(row, col)
		float val[3], b[3], z[3];
		for (int ip = 0; ip < n; ++ip) { b[ip] = a[ip][ip]; val[ip] = a[ip][ip]; z[ip] = 0.0f; }
		bool converged = false;
		for (int sweep = 0; sweep < 50 && !converged; ++sweep)
		{
			float sm = 0.0f;
			for (int ip = 0; ip < n - 1; ++ip) for (int iq = ip + 1; iq < n; ++iq) sm += std::fabs(a[ip][iq]);
			float avg_sm = sm / float(n * n);
			if (avg_sm < FLT_MIN) { converged = true; break; }
			float thresh = sweep < 4? 0.2f * avg_sm : FLT_MIN;
			for (int ip = 0; ip < n - 1; ++ip)
				for (int iq = ip + 1; iq < n; ++iq)
				{
					float &a_pq = a[ip][iq];
					float &eigval_p = val[ip];
					float &eigval_q = val[iq];
					float abs_a_pq = std::fabs(a_pq);
					float g = 100.0f * abs_a_pq;
					if (sweep > 4 && std::fabs(eigval_p) + g == std::fabs(eigval_p) && std::fabs(eigval_q) + g == std::fabs(eigval_q))
						a_pq = 0.0f;
					else if (abs_a_pq > thresh)
					{
						float h = eigval_q - eigval_p;
						float abs_h = std::fabs(h);
						float t;
						if (abs_h + g == abs_h)
							t = a_pq / h;
						else
						{
							float theta = 0.5f * h / a_pq;
							t = 1.0f / (std::fabs(theta) + std::sqrt(1.0f + theta * theta));
							if (theta < 0.0f) t = -t;
						}
						float c = 1.0f / std::sqrt(1.0f + t * t);
						float s = t * c;
						float tau = s / (1.0f + c);
						h = t * a_pq;
						a_pq = 0.0f;
						z[ip] -= h; z[iq] += h;
						eigval_p -= h; eigval_q += h;
						auto rotate = [&](float (&m)[3][3], int i, int j, int k, int l) { float gg = m[i][j], hh = m[k][l]; m[i][j] = gg - s * (hh + gg * tau); m[k][l] = hh + s * (gg - hh * tau); };
						int j;
						for (j = 0; j < ip; ++j) rotate(a, j, ip, j, iq);
						for (j = ip + 1; j < iq; ++j) rotate(a, ip, j, j, iq);
						for (j = iq + 1; j < n; ++j) rotate(a, ip, j, iq, j);
						for (j = 0; j < n; ++j) rotate(vec, j, ip, j, iq);
					}
				}
			for (int ip = 0; ip < n; ++ip) { b[ip] += z[ip]; val[ip] = b[ip]; z[ip] = 0.0f; }
		}
		if (!converged)
			return false;
		// insertion sort, biggest first
		int indices[3] = { 0, 1, 2 };
		for (int i = 1; i < 3; ++i)
		{
			int x = indices[i], j = i;
			while (j > 0 && val[x] > val[indices[j - 1]]) { indices[j] = indices[j - 1]; --j; }
			indices[j] = x;
		}
		Vec3 col[3];
		float diag[3];
		for (int i = 0; i < 3; ++i)
		{
			col[i] = Vec3(vec[0][indices[i]], vec[1][indices[i]], vec[2][indices[i]]);
			diag[i] = val[indices[i]];
		}
		if (col[0].Cross(col[1]).Dot(col[2]) < 0.0f)
			col[2] = -col[2];
		outDiagonal = Vec3(diag[0], diag[1], diag[2]);
		// Mat44::GetQuaternion (Mat44.inl:998-1050); m(row, col) = col[col][row]
		float m00 = col[0].x, m11 = col[1].y, m22 = col[2].z;
		float tr = m00 + m11 + m22;
		if (tr >= 0.0f)
		{
			float s = std::sqrt(tr + 1.0f), is = 0.5f / s;
			outRotation = Quat((col[1].z - col[2].y) * is, (col[2].x - col[0].z) * is, (col[0].y - col[1].x) * is, 0.5f * s);
		}
		else
		{
			int i = 0;
			if (m11 > m00) i = 1;
			if (m22 > (i == 0? m00 : m11)) i = 2;
			if (i == 0)
			{
				float s = std::sqrt(m00 - (m11 + m22) + 1), is = 0.5f / s;
				outRotation = Quat(0.5f * s, (col[1].x + col[0].y) * is, (col[0].z + col[2].x) * is, (col[1].z - col[2].y) * is);
			}
			else if (i == 1)
			{
				float s = std::sqrt(m11 - (m22 + m00) + 1), is = 0.5f / s;
				outRotation = Quat((col[1].x + col[0].y) * is, 0.5f * s, (col[2].y + col[1].z) * is, (col[2].x - col[0].z) * is);
			}
			else
			{
				float s = std::sqrt(m22 - (m00 + m11) + 1), is = 0.5f / s;
				outRotation = Quat((col[0].z + col[2].x) * is, (col[2].y + col[1].z) * is, 0.5f * s, (col[0].y - col[1].x) * is);
			}
		}
		return true;
	}
};

// ---- shapes (immutable, shared) -----------------------------------------------------------------------------------
class Shape
{
public:
	virtual ~Shape() = default;
	virtual EShapeSubType GetSubType() const = 0;
	virtual MassProperties GetMassProperties() const = 0;
	virtual Vec3 GetCenterOfMass() const { return Vec3::sZero(); }
	virtual int32_t Upload(b2j_world *inWorld) const = 0;
	// (what a compound needs from its sub shapes: Shape::GetLocalBounds / GetWorldSpaceBounds / GetInnerRadius)
	virtual AABox GetLocalBounds() const = 0;
	virtual float GetInnerRadius() const = 0;
	virtual AABox GetWorldSpaceBounds(const Mat44RT &inCenterOfMassTransform, const Vec3 &inScale) const { return GetLocalBounds().Scaled(inScale).Transformed(inCenterOfMassTransform); } // Shape.h:221
};
using ShapeRef = std::shared_ptr<const Shape>;

class ConvexShape : public Shape { public: void SetDensity(float d) { mDensity = d; } float GetDensity() const { return mDensity; } protected: float mDensity = 1000.0f; };

class SphereShape final : public ConvexShape
{
public:
	explicit SphereShape(float inRadius) : mRadius(inRadius) { }
	float GetRadius() const { return mRadius; }
	EShapeSubType GetSubType() const override { return EShapeSubType::Sphere; }
	MassProperties GetMassProperties() const override // SphereShape.cpp:143-156
	{
		MassProperties p;
		float r2 = mRadius * mRadius;
		p.mMass = (4.0f / 3.0f * JPH_PI) * mRadius * r2 * GetDensity();
		float inertia = (2.0f / 5.0f) * p.mMass * r2;
		p.SetDiagonal(inertia, inertia, inertia);
		return p;
	}
	int32_t Upload(b2j_world *w) const override { return b2j_shape_sphere(w, mRadius); }
	AABox GetLocalBounds() const override { return AABox(Vec3::sReplicate(-mRadius), Vec3::sReplicate(mRadius)); }
	float GetInnerRadius() const override { return mRadius; }
	AABox GetWorldSpaceBounds(const Mat44RT &m, const Vec3 &inScale) const override // SphereShape.cpp:67-74
	{
		Vec3 half_extent = Vec3::sReplicate(std::fabs(inScale.x) * mRadius);
		return AABox(-half_extent + m.t, half_extent + m.t);
	}
private:
	float mRadius;
};

class BoxShape final : public ConvexShape
{
public:
	explicit BoxShape(const Vec3 &inHalfExtent, float inConvexRadius = 0.05f) : mHalfExtent(inHalfExtent), mConvexRadius(inConvexRadius) { }
	Vec3 GetHalfExtent() const { return mHalfExtent; }
	float GetConvexRadius() const { return mConvexRadius; }
	EShapeSubType GetSubType() const override { return EShapeSubType::Box; }
	MassProperties GetMassProperties() const override { MassProperties p; p.SetMassAndInertiaOfSolidBox(2.0f * mHalfExtent, GetDensity()); return p; } // BoxShape.cpp:149-154
	int32_t Upload(b2j_world *w) const override { float he[3] = { mHalfExtent.x, mHalfExtent.y, mHalfExtent.z }; return b2j_shape_box(w, he, mConvexRadius); }
	AABox GetLocalBounds() const override { return AABox(-mHalfExtent, mHalfExtent); }
	float GetInnerRadius() const override { return std::min(mHalfExtent.x, std::min(mHalfExtent.y, mHalfExtent.z)); }
private:
	Vec3 mHalfExtent;
	float mConvexRadius;
};

class CapsuleShape final : public ConvexShape
{
public:
	CapsuleShape(float inHalfHeightOfCylinder, float inRadius) : mHalfHeightOfCylinder(inHalfHeightOfCylinder), mRadius(inRadius) { }
	EShapeSubType GetSubType() const override { return EShapeSubType::Capsule; }
	MassProperties GetMassProperties() const override // CapsuleShape.cpp:218-248
	{
		MassProperties p;
		float density = GetDensity();
		float radius_sq = mRadius * mRadius;
		float height = 2.0f * mHalfHeightOfCylinder;
		float cylinder_mass = JPH_PI * height * radius_sq * density;
		float hemisphere_mass = (2.0f * JPH_PI / 3.0f) * radius_sq * mRadius * density;
		float height_sq = height * height;
		float inertia_y = radius_sq * cylinder_mass * 0.5f;
		float inertia_xz = inertia_y * 0.5f + cylinder_mass * height_sq / 12.0f;
		float temp = hemisphere_mass * 4.0f * radius_sq / 5.0f;
		inertia_y += temp;
		inertia_xz += temp + hemisphere_mass * (0.5f * height_sq + (3.0f / 4.0f) * height * mRadius);
		p.mMass = cylinder_mass + hemisphere_mass * 2.0f;
		p.SetDiagonal(inertia_xz, inertia_y, inertia_xz);
		return p;
	}
	int32_t Upload(b2j_world *w) const override { return b2j_shape_capsule(w, mHalfHeightOfCylinder, mRadius); }
	AABox GetLocalBounds() const override { Vec3 extent = Vec3::sReplicate(mRadius) + Vec3(0, mHalfHeightOfCylinder, 0); return AABox(-extent, extent); } // CapsuleShape.cpp:259-264
	float GetInnerRadius() const override { return mRadius; }
	AABox GetWorldSpaceBounds(const Mat44RT &m, const Vec3 &inScale) const override // CapsuleShape.cpp:266-277
	{
		float scale = std::fabs(inScale.x);
		Vec3 extent = Vec3::sReplicate(scale * mRadius), height = Vec3(0, scale * mHalfHeightOfCylinder, 0);
		Vec3 p1 = m * -height, p2 = m * height;
		return AABox(AABox::sMin(p1, p2) - extent, AABox::sMax(p1, p2) + extent);
	}
private:
	float mHalfHeightOfCylinder, mRadius;
};

class CylinderShape final : public ConvexShape
{
public:
	CylinderShape(float inHalfHeight, float inRadius, float inConvexRadius = 0.05f) : mHalfHeight(inHalfHeight), mRadius(inRadius), mConvexRadius(std::min(inConvexRadius, std::min(inHalfHeight, inRadius))) { }
	float GetHalfHeight() const { return mHalfHeight; } float GetRadius() const { return mRadius; } float GetConvexRadius() const { return mConvexRadius; }
	EShapeSubType GetSubType() const override { return EShapeSubType::Cylinder; }
	MassProperties GetMassProperties() const override // CylinderShape.cpp:246-263
	{
		MassProperties p;
		float radius_sq = mRadius * mRadius;
		float height = 2.0f * mHalfHeight;
		p.mMass = JPH_PI * radius_sq * height * GetDensity();
		float inertia_y = radius_sq * p.mMass * 0.5f;
		float inertia_x = inertia_y * 0.5f + p.mMass * height * height / 12.0f;
		p.SetDiagonal(inertia_x, inertia_y, inertia_x);
		return p;
	}
	int32_t Upload(b2j_world *w) const override { return b2j_shape_cylinder(w, mHalfHeight, mRadius, mConvexRadius); }
	AABox GetLocalBounds() const override { Vec3 extent(mRadius, mHalfHeight, mRadius); return AABox(-extent, extent); }
	float GetInnerRadius() const override { return std::min(mHalfHeight, mRadius); }
private:
	float mHalfHeight, mRadius, mConvexRadius;
};

// A convex hull cooked by the reference's ConvexHullBuilder (host-side cooking is out of scope, SURVEY 2a Jolt/Geometry)
class ConvexHullShape final : public ConvexShape
{
public:
	std::vector<float> mPoints, mPlanes;          // [n][3], [f][4]
	std::vector<int32_t> mPointNumFaces, mPointFaces;
	std::vector<uint16_t> mFaceFirstVertex, mFaceNumVertices;
	std::vector<uint8_t> mVertexIdx;
	float mConvexRadius = 0.0f, mInnerRadius = 0.0f, mVolume = 0.0f;
	float mCenterOfMass[3] = { 0, 0, 0 }, mBoundsMin[3] = { 0, 0, 0 }, mBoundsMax[3] = { 0, 0, 0 };
	float mInertia[3][3] = { { 0 } };           // density 1, [column][row]
	EShapeSubType GetSubType() const override { return EShapeSubType::ConvexHull; }
	Vec3 GetCenterOfMass() const override { return Vec3(mCenterOfMass[0], mCenterOfMass[1], mCenterOfMass[2]); }
	AABox GetLocalBounds() const override { return AABox(Vec3(mBoundsMin[0], mBoundsMin[1], mBoundsMin[2]), Vec3(mBoundsMax[0], mBoundsMax[1], mBoundsMax[2])); }
	float GetInnerRadius() const override { return mInnerRadius; }
	MassProperties GetMassProperties() const override // ConvexHullShape.cpp:356-370
	{
		MassProperties p;
		float density = GetDensity();
		p.mMass = density * mVolume;
		for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) p.mInertia[c][r] = density * mInertia[c][r];
		return p;
	}
	int32_t Upload(b2j_world *w) const override
	{
		b2j_hull_desc d;
		memset(&d, 0, sizeof(d));
		d.num_points = (uint32_t)mPointNumFaces.size(); d.points = mPoints.data(); d.point_num_faces = mPointNumFaces.data(); d.point_faces = mPointFaces.data();
		d.num_faces = (uint32_t)mFaceFirstVertex.size(); d.face_first_vertex = mFaceFirstVertex.data(); d.face_num_vertices = mFaceNumVertices.data(); d.planes = mPlanes.data();
		d.num_vertex_idx = (uint32_t)mVertexIdx.size(); d.vertex_idx = mVertexIdx.data(); d.convex_radius = mConvexRadius; d.inner_radius = mInnerRadius;
		memcpy(d.center_of_mass, mCenterOfMass, 12); memcpy(d.local_bounds_min, mBoundsMin, 12); memcpy(d.local_bounds_max, mBoundsMax, 12);
		return b2j_shape_convex_hull(w, &d);
	}
};

// A mesh cooked by the reference's MeshShapeSettings (tree byte buffer verbatim)
class MeshShape final : public Shape
{
public:
	std::vector<uint8_t> mTree;
	float mBoundsMin[3] = { 0, 0, 0 }, mBoundsMax[3] = { 0, 0, 0 };
	EShapeSubType GetSubType() const override { return EShapeSubType::Mesh; }
	MassProperties GetMassProperties() const override { return MassProperties(); } // static only
	AABox GetLocalBounds() const override { return AABox(Vec3(mBoundsMin[0], mBoundsMin[1], mBoundsMin[2]), Vec3(mBoundsMax[0], mBoundsMax[1], mBoundsMax[2])); }
	float GetInnerRadius() const override { return 0.0f; }
	int32_t Upload(b2j_world *w) const override
	{
		b2j_mesh_desc d;
		memset(&d, 0, sizeof(d));
		d.tree = mTree.data(); d.tree_size = (uint32_t)mTree.size();
		memcpy(d.local_bounds_min, mBoundsMin, 12); memcpy(d.local_bounds_max, mBoundsMax, 12);
		return b2j_shape_mesh(w, &d);
	}
};

// ---- decorated shapes around convex shapes (DecoratedShape.h, ScaledShape.h, RotatedTranslatedShape.h) ------------
class DecoratedShape : public Shape
{
public:
	explicit DecoratedShape(ShapeRef inInnerShape) : mInnerShape(std::move(inInnerShape)) { }
	const Shape *GetInnerShape() const { return mInnerShape.get(); }
protected:
	ShapeRef mInnerShape;
};

// ScaledShape(shape, scale): positive scales; spheres / capsules / rotated inner shapes take uniform scales (b2j_shape_scaled)
class ScaledShape final : public DecoratedShape
{
public:
	ScaledShape(ShapeRef inShape, const Vec3 &inScale) : DecoratedShape(std::move(inShape)), mScale(inScale) { }
	Vec3 GetScale() const { return mScale; }
	EShapeSubType GetSubType() const override { return EShapeSubType::Scaled; }
	Vec3 GetCenterOfMass() const override { return mScale * mInnerShape->GetCenterOfMass(); }                                   // ScaledShape.h:56
	MassProperties GetMassProperties() const override { MassProperties p = mInnerShape->GetMassProperties(); p.Scale(mScale); return p; } // ScaledShape.cpp:49-54
	AABox GetLocalBounds() const override { return mInnerShape->GetLocalBounds().Scaled(mScale); }
	float GetInnerRadius() const override { return std::min(mScale.x, std::min(mScale.y, mScale.z)) * mInnerShape->GetInnerRadius(); }
	AABox GetWorldSpaceBounds(const Mat44RT &m, const Vec3 &inScale) const override { return mInnerShape->GetWorldSpaceBounds(m, inScale * mScale); }
	int32_t Upload(b2j_world *w) const override
	{
		int32_t inner = mInnerShape->Upload(w);
		if (inner < 0) return inner;
		float scale[3] = { mScale.x, mScale.y, mScale.z };
		return b2j_shape_scaled(w, inner, scale);
	}
private:
	Vec3 mScale;
};

// RotatedTranslatedShape(position, rotation, shape): the inner shape rotated and moved relative to the body (RotatedTranslatedShape.cpp:50-67)
class RotatedTranslatedShape final : public DecoratedShape
{
public:
	RotatedTranslatedShape(const Vec3 &inPosition, const Quat &inRotation, ShapeRef inShape) : DecoratedShape(std::move(inShape)), mRotation(inRotation)
	{
		mCenterOfMass = inPosition + inRotation * mInnerShape->GetCenterOfMass();
	}
	Quat GetRotation() const { return mRotation; }
	Vec3 GetPosition() const { return mCenterOfMass - mRotation * mInnerShape->GetCenterOfMass(); }
	EShapeSubType GetSubType() const override { return EShapeSubType::RotatedTranslated; }
	Vec3 GetCenterOfMass() const override { return mCenterOfMass; }
	MassProperties GetMassProperties() const override { MassProperties p = mInnerShape->GetMassProperties(); p.Rotate(mRotation); return p; }
	AABox GetLocalBounds() const override { return mInnerShape->GetLocalBounds().Transformed(Mat44RT::sRotation(mRotation)); }
	float GetInnerRadius() const override { return mInnerShape->GetInnerRadius(); }
	AABox GetWorldSpaceBounds(const Mat44RT &m, const Vec3 &inScale) const override { return mInnerShape->GetWorldSpaceBounds(m * Mat44RT::sRotation(mRotation), inScale); } // (uniform scales only: TransformScale is the identity)
	int32_t Upload(b2j_world *w) const override
	{
		int32_t inner = mInnerShape->Upload(w);
		if (inner < 0) return inner;
		float rotation[4] = { mRotation.x, mRotation.y, mRotation.z, mRotation.w }, com[3] = { mCenterOfMass.x, mCenterOfMass.y, mCenterOfMass.z };
		return b2j_shape_rotated_translated(w, inner, rotation, com);
	}
private:
	Quat mRotation;
	Vec3 mCenterOfMass;
};

// ---- StaticCompoundShape (StaticCompoundShape.h / CompoundShape.h) of convex sub shapes ------------------------------------------
// Built the way the reference builds it, because the build decides what the simulation computes: the centre of mass (mass weighted),
// the sub shape positions relative to it, the compressed rotations, and the quad tree over the sub shape bounds whose layout fixes
// the order in which sub shapes are collided (StaticCompoundShape.cpp:110-357).
class StaticCompoundShape final : public Shape
{
public:
	struct SubShape { ShapeRef mShape; Vec3 mPositionCOM; Quat mRotation; bool mIsRotationIdentity = true; };

	EShapeSubType GetSubType() const override { return EShapeSubType::StaticCompound; }
	Vec3 GetCenterOfMass() const override { return mCenterOfMass; }
	AABox GetLocalBounds() const override { return mLocalBounds; }
	float GetInnerRadius() const override { return mInnerRadius; }
	uint GetNumSubShapes() const { return (uint)mSubShapes.size(); }
	const SubShape &GetSubShape(uint inIdx) const { return mSubShapes[inIdx]; }
	MassProperties GetMassProperties() const override // CompoundShape.cpp:68-91
	{
		MassProperties p;
		for (const SubShape &shape : mSubShapes)
		{
			MassProperties child = shape.mShape->GetMassProperties();
			child.Rotate(shape.mRotation);
			child.Translate(shape.mPositionCOM);
			p.mMass += child.mMass;
			for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) p.mInertia[c][r] += child.mInertia[c][r];
		}
		return p;
	}
	AABox GetWorldSpaceBounds(const Mat44RT &m, const Vec3 &inScale) const override // CompoundShape.cpp:93-115
	{
		if (mSubShapes.size() > 10)
			return Shape::GetWorldSpaceBounds(m, inScale);
		AABox bounds;
		for (const SubShape &shape : mSubShapes)
			bounds.Encapsulate(shape.mShape->GetWorldSpaceBounds(m * Mat44RT::sRotationTranslation(shape.mRotation, inScale * shape.mPositionCOM), inScale));
		return bounds;
	}
	int32_t Upload(b2j_world *w) const override
	{
		std::vector<b2j_compound_sub> subs(mSubShapes.size());
		for (size_t i = 0; i < mSubShapes.size(); ++i)
		{
			const SubShape &in = mSubShapes[i];
			subs[i].shape = in.mShape->Upload(w);
			if (subs[i].shape < 0) return subs[i].shape;
			subs[i].position_com[0] = in.mPositionCOM.x; subs[i].position_com[1] = in.mPositionCOM.y; subs[i].position_com[2] = in.mPositionCOM.z;
			subs[i].rotation[0] = in.mRotation.x; subs[i].rotation[1] = in.mRotation.y; subs[i].rotation[2] = in.mRotation.z; subs[i].rotation[3] = in.mRotation.w;
		}
		b2j_compound_desc d;
		memset(&d, 0, sizeof(d));
		d.num_subs = (uint32_t)subs.size(); d.subs = subs.data();
		d.num_nodes = (uint32_t)mNodes.size(); d.nodes = reinterpret_cast<const uint8_t *>(mNodes.data());
		d.center_of_mass[0] = mCenterOfMass.x; d.center_of_mass[1] = mCenterOfMass.y; d.center_of_mass[2] = mCenterOfMass.z;
		d.local_bounds_min[0] = mLocalBounds.mMin.x; d.local_bounds_min[1] = mLocalBounds.mMin.y; d.local_bounds_min[2] = mLocalBounds.mMin.z;
		d.local_bounds_max[0] = mLocalBounds.mMax.x; d.local_bounds_max[1] = mLocalBounds.mMax.y; d.local_bounds_max[2] = mLocalBounds.mMax.z;
		d.inner_radius = mInnerRadius;
		return b2j_shape_static_compound(w, &d);
	}

	// 4 child bounding boxes as half floats + 4 child properties (StaticCompoundShape::Node, 64 bytes)
	struct Node { uint16_t mBoundsMinX[4], mBoundsMinY[4], mBoundsMinZ[4], mBoundsMaxX[4], mBoundsMaxY[4], mBoundsMaxZ[4]; uint32_t mNodeProperties[4]; };
	static_assert(sizeof(Node) == 64, "Node should be 64 bytes");
	enum : uint32_t { IS_SUBSHAPE = 0x80000000u, INVALID_NODE = 0x7fffffffu };

private:
	friend class StaticCompoundShapeSettings;
	// float -> half rounding towards -inf (inUp = false) or +inf (HalfFloatConversion::FromFloat<ROUND_TO_NEG_INF / ROUND_TO_POS_INF>)
	static uint16_t sToHalf(float inV, bool inUp)
	{
		uint32_t value; memcpy(&value, &inV, 4);
		uint32_t exponent = (value >> 23) & 0xffu, mantissa = value & 0x7fffffu;
		uint16_t sign = uint16_t(value >> 16) & 0x8000u;
		bool away = (sign == 0) == inUp; // rounding in this direction grows the magnitude
		if (exponent == 0xffu) return sign | (mantissa == 0? 0x7c00u : 0x7e00u);
		int e = int(exponent) - 127 + 15;
		if (e >= 31) return sign | (away? 0x7c00u : 0x7bffu);
		if (e < -10) return sign | ((away && (value & 0x7fffffffu) != 0)? 1u : 0u);
		uint16_t hf_exponent; int shift;
		if (e <= 0) { hf_exponent = 0; mantissa |= 1u << 23; shift = 23 - 10 + 1 - e; }
		else { hf_exponent = uint16_t(e << 10); shift = 13; }
		uint16_t hf = sign | hf_exponent | uint16_t(mantissa >> shift);
		if (away && (mantissa & ((1u << shift) - 1u)) != 0) ++hf;
		return hf;
	}
	static void sSetChildBounds(Node &n, uint i, const AABox &b)
	{
		n.mBoundsMinX[i] = sToHalf(b.mMin.x, false); n.mBoundsMinY[i] = sToHalf(b.mMin.y, false); n.mBoundsMinZ[i] = sToHalf(b.mMin.z, false);
		n.mBoundsMaxX[i] = sToHalf(b.mMax.x, true); n.mBoundsMaxY[i] = sToHalf(b.mMax.y, true); n.mBoundsMaxZ[i] = sToHalf(b.mMax.z, true);
	}
	static void sSetChildInvalid(Node &n, uint i)
	{
		n.mNodeProperties[i] = INVALID_NODE;
		n.mBoundsMinX[i] = n.mBoundsMinY[i] = n.mBoundsMinZ[i] = n.mBoundsMaxX[i] = n.mBoundsMaxY[i] = n.mBoundsMaxZ[i] = 0x7bffu; // HALF_FLT_MAX
	}
	// StaticCompoundShape::sPartition: split a range of sub shapes at the middle of the widest axis of their bounds' centres
	static void sPartition(uint *ioIdx, AABox *ioBounds, int inNumber, int &outMidPoint)
	{
		if (inNumber <= 4) { outMidPoint = inNumber / 2; return; }
		Vec3 center_min = Vec3::sReplicate(FLT_MAX), center_max = Vec3::sReplicate(-FLT_MAX);
		for (int i = 0; i < inNumber; ++i) { Vec3 c = ioBounds[i].GetCenter(); center_min = AABox::sMin(center_min, c); center_max = AABox::sMax(center_max, c); }
		Vec3 extent = center_max - center_min;
		int dimension = extent.x > extent.y? (extent.z > extent.x? 2 : 0) : (extent.z > extent.y? 2 : 1); // Vec3::GetHighestComponentIndex
		float split = (0.5f * (center_min + center_max))[dimension];
		int start = 0, end = inNumber;
		while (start < end)
		{
			while (start < end && ioBounds[start].GetCenter()[dimension] < split) ++start;
			while (start < end && ioBounds[end - 1].GetCenter()[dimension] >= split) --end;
			if (start < end) { std::swap(ioIdx[start], ioIdx[end - 1]); std::swap(ioBounds[start], ioBounds[end - 1]); ++start; --end; }
		}
		outMidPoint = (start > 0 && start < inNumber)? start : inNumber / 2;
	}
	static void sPartition4(uint *ioIdx, AABox *ioBounds, int inBegin, int inEnd, int *outSplit)
	{
		uint *idx = ioIdx + inBegin; AABox *bounds = ioBounds + inBegin; int number = inEnd - inBegin;
		sPartition(idx, bounds, number, outSplit[2]);
		sPartition(idx, bounds, outSplit[2], outSplit[1]);
		sPartition(idx + outSplit[2], bounds + outSplit[2], number - outSplit[2], outSplit[3]);
		outSplit[0] = inBegin; outSplit[1] += inBegin; outSplit[2] += inBegin; outSplit[3] += outSplit[2]; outSplit[4] = inEnd;
	}

	Vec3 mCenterOfMass = Vec3::sZero();
	AABox mLocalBounds = AABox(Vec3::sZero(), Vec3::sZero());
	float mInnerRadius = FLT_MAX;
	std::vector<SubShape> mSubShapes;
	std::vector<Node> mNodes;
};

// StaticCompoundShapeSettings: AddShape(position, rotation, shape) ... Create() (CompoundShape.h:30-70, StaticCompoundShape.cpp:24-66)
class StaticCompoundShapeSettings
{
public:
	void AddShape(const Vec3 &inPosition, const Quat &inRotation, ShapeRef inShape) { mParts.push_back({ inPosition, inRotation, std::move(inShape) }); }
	// One sub shape: the shape itself or a RotatedTranslatedShape, like the reference; none: null
	ShapeRef Create() const
	{
		if (mParts.empty()) return nullptr;
		if (mParts.size() == 1)
		{
			const Part &s = mParts[0];
			if (s.mPosition.x == 0.0f && s.mPosition.y == 0.0f && s.mPosition.z == 0.0f && s.mRotation.x == 0.0f && s.mRotation.y == 0.0f && s.mRotation.z == 0.0f && s.mRotation.w == 1.0f)
				return s.mShape;
			return std::make_shared<RotatedTranslatedShape>(s.mPosition, s.mRotation, s.mShape);
		}
		auto out = std::make_shared<StaticCompoundShape>();
		StaticCompoundShape &c = *out;
		uint n = (uint)mParts.size();
		c.mSubShapes.resize(n);
		float mass = 0.0f;
		for (uint i = 0; i < n; ++i)
		{
			// SubShape::FromSettings -> SetTransform(position, rotation, zero): position of the sub shape's centre of mass, rotation compressed
			// to x, y, z with w >= 0 (Quat::StoreFloat3 / sLoadFloat3Unsafe)
			StaticCompoundShape::SubShape &sub = c.mSubShapes[i];
			sub.mShape = mParts[i].mShape;
			const Quat &q = mParts[i].mRotation;
			sub.mPositionCOM = (mParts[i].mPosition - Vec3::sZero()) + q * sub.mShape->GetCenterOfMass();
			auto is_close = [](const Quat &a, float sign) { float dx = a.x, dy = a.y, dz = a.z, dw = a.w - sign; return (dx * dx + dy * dy) + (dz * dz + dw * dw) <= 1.0e-12f; }; // Quat::IsClose
			sub.mIsRotationIdentity = is_close(q, 1.0f) || is_close(q, -1.0f);
			if (sub.mIsRotationIdentity)
				sub.mRotation = Quat::sIdentity();
			else
			{
				float sign = q.w < 0.0f? -1.0f : 1.0f; // EnsureWPositive flips the sign bits
				Vec3 v(sign < 0.0f? -q.x : q.x, sign < 0.0f? -q.y : q.y, sign < 0.0f? -q.z : q.z);
				sub.mRotation = Quat(v.x, v.y, v.z, std::sqrt(std::max(1.0f - v.LengthSq(), 0.0f)));
			}
			MassProperties child = sub.mShape->GetMassProperties();
			mass += child.mMass;
			c.mCenterOfMass = c.mCenterOfMass + sub.mPositionCOM * child.mMass;
		}
		if (mass > 0.0f)
			c.mCenterOfMass = Vec3(c.mCenterOfMass.x / mass, c.mCenterOfMass.y / mass, c.mCenterOfMass.z / mass);
		c.mInnerRadius = FLT_MAX;
		for (const StaticCompoundShape::SubShape &sub : c.mSubShapes) c.mInnerRadius = std::min(c.mInnerRadius, sub.mShape->GetInnerRadius());

		std::vector<AABox> bounds(n);
		std::vector<uint> idx(n);
		for (uint i = 0; i < n; ++i)
		{
			StaticCompoundShape::SubShape &sub = c.mSubShapes[i];
			sub.mPositionCOM = sub.mPositionCOM - c.mCenterOfMass;
			bounds[i] = sub.mShape->GetWorldSpaceBounds(Mat44RT::sRotationTranslation(sub.mRotation, sub.mPositionCOM), Vec3::sReplicate(1.0f));
			idx[i] = i;
			c.mLocalBounds.Encapsulate(bounds[i]);
		}

		// the quad tree, built depth first with an explicit stack like the reference
		struct StackEntry { uint32_t mNodeIdx; int mChildIdx; int mSplit[5]; AABox mBounds; };
		std::vector<StackEntry> stack(n);
		c.mNodes.assign(n + (n + 2) / 3, StaticCompoundShape::Node());
		uint32_t next_node_idx = 0;
		int top = 0;
		stack[0].mNodeIdx = next_node_idx++; stack[0].mChildIdx = -1; stack[0].mBounds = AABox();
		StaticCompoundShape::sPartition4(idx.data(), bounds.data(), 0, (int)n, stack[0].mSplit);
		for (;;)
		{
			StackEntry &cur = stack[top];
			cur.mChildIdx++;
			if (cur.mChildIdx >= 4)
			{
				if (top <= 0) break;
				StackEntry &prev = stack[top - 1];
				prev.mBounds.Encapsulate(cur.mBounds);
				StaticCompoundShape::Node &parent = c.mNodes[prev.mNodeIdx];
				parent.mNodeProperties[prev.mChildIdx] = cur.mNodeIdx;
				StaticCompoundShape::sSetChildBounds(parent, (uint)prev.mChildIdx, cur.mBounds);
				--top;
			}
			else
			{
				int low = cur.mSplit[cur.mChildIdx], high = cur.mSplit[cur.mChildIdx + 1];
				int num = high - low;
				StaticCompoundShape::Node &node = c.mNodes[cur.mNodeIdx];
				if (num == 0)
					StaticCompoundShape::sSetChildInvalid(node, (uint)cur.mChildIdx);
				else if (num == 1)
				{
					node.mNodeProperties[cur.mChildIdx] = idx[low] | StaticCompoundShape::IS_SUBSHAPE;
					StaticCompoundShape::sSetChildBounds(node, (uint)cur.mChildIdx, bounds[low]);
					cur.mBounds.Encapsulate(bounds[low]);
				}
				else
				{
					StackEntry &next = stack[++top];
					next.mNodeIdx = next_node_idx++; next.mChildIdx = -1; next.mBounds = AABox();
					StaticCompoundShape::sPartition4(idx.data(), bounds.data(), low, high, next.mSplit);
				}
			}
		}
		c.mNodes.resize(next_node_idx);
		return out;
	}
private:
	struct Part { Vec3 mPosition; Quat mRotation; ShapeRef mShape; };
	std::vector<Part> mParts;
};

// ---- BodyCreationSettings (same defaults as the reference) ------------------------------------------------------
class BodyCreationSettings
{
public:
	BodyCreationSettings() = default;
	BodyCreationSettings(ShapeRef inShape, const RVec3 &inPosition, const Quat &inRotation, EMotionType inMotionType, ObjectLayer inObjectLayer) :
		mPosition(inPosition), mRotation(inRotation), mObjectLayer(inObjectLayer), mMotionType(inMotionType), mShape(std::move(inShape)) { }
	void SetShape(ShapeRef inShape) { mShape = std::move(inShape); }
	const ShapeRef &GetShape() const { return mShape; }
	bool HasMassProperties() const { return mAllowDynamicOrKinematic || mMotionType != EMotionType::Static; }
	MassProperties GetMassProperties() const // BodyCreationSettings.cpp:196-218
	{
		MassProperties mp;
		switch (mOverrideMassProperties)
		{
		case EOverrideMassProperties::CalculateMassAndInertia:
			mp = mShape->GetMassProperties();
			for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mp.mInertia[c][r] *= mInertiaMultiplier;
			break;
		case EOverrideMassProperties::CalculateInertia:
			mp = mShape->GetMassProperties();
			mp.ScaleToMass(mMassPropertiesOverride.mMass);
			for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mp.mInertia[c][r] *= mInertiaMultiplier;
			break;
		case EOverrideMassProperties::MassAndInertiaProvided:
			mp = mMassPropertiesOverride;
			break;
		}
		return mp;
	}

	RVec3 mPosition = RVec3::sZero();
	Quat mRotation = Quat::sIdentity();
	Vec3 mLinearVelocity = Vec3::sZero();
	Vec3 mAngularVelocity = Vec3::sZero();
	uint64 mUserData = 0;
	ObjectLayer mObjectLayer = 0;
	EMotionType mMotionType = EMotionType::Dynamic;
	EAllowedDOFs mAllowedDOFs = EAllowedDOFs::All;
	bool mAllowDynamicOrKinematic = false;
	bool mIsSensor = false;
	bool mCollideKinematicVsNonDynamic = false;
	bool mUseManifoldReduction = true;
	bool mApplyGyroscopicForce = false;
	EMotionQuality mMotionQuality = EMotionQuality::Discrete;
	bool mAllowSleeping = true;
	float mFriction = 0.2f;
	float mRestitution = 0.0f;
	float mLinearDamping = 0.05f;
	float mAngularDamping = 0.05f;
	float mMaxLinearVelocity = 500.0f;
	float mMaxAngularVelocity = 0.25f * JPH_PI * 60.0f;
	float mGravityFactor = 1.0f;
	uint mNumVelocityStepsOverride = 0;
	uint mNumPositionStepsOverride = 0;
	EOverrideMassProperties mOverrideMassProperties = EOverrideMassProperties::CalculateMassAndInertia;
	float mInertiaMultiplier = 1.0f;
	MassProperties mMassPropertiesOverride;
private:
	ShapeRef mShape;
};

struct PhysicsSettings
{
	float mBaumgarte = 0.2f;
	float mSpeculativeContactDistance = 0.02f;
	float mPenetrationSlop = 0.02f;
	float mMaxPenetrationDistance = 0.2f;
	float mManifoldTolerance = 1.0e-3f;
	float mMinVelocityForRestitution = 1.0f;
	float mTimeBeforeSleep = 0.5f;
	float mPointVelocitySleepThreshold = 0.03f;
	uint mNumVelocitySteps = 10;
	uint mNumPositionSteps = 2;
	bool mDeterministicSimulation = true;
	bool mConstraintWarmStart = true;
	bool mUseBodyPairContactCache = true;
	bool mUseManifoldReduction = true;
	bool mUseLargeIslandSplitter = true;
	bool mAllowSleeping = true;
	bool mCheckActiveEdges = true;
};

class PhysicsSystem;
// ---- queries (NarrowPhaseQuery / BroadPhaseQuery, Jolt/Physics/Collision/NarrowPhaseQuery.h:31, BroadPhase/BroadPhaseQuery.h:38) ----
struct RRayCast { RVec3 mOrigin; Vec3 mDirection; RRayCast() = default; RRayCast(const RVec3 &o, const Vec3 &d) : mOrigin(o), mDirection(d) { } };
using RayCast = RRayCast;
struct RayCastResult { BodyID mBodyID; float mFraction = 1.0f + FLT_EPSILON; SubShapeID mSubShapeID2; };
// CollideShapeSettings / CollideShapeResult (Jolt/Physics/Collision/CollideShape.h): the members the device path honours (back faces are
// ignored, only active edges collide, no faces are collected: the defaults of the reference)
struct CollideShapeSettings { float mMaxSeparationDistance = 0.0f; };
struct CollideShapeResult
{
	Vec3 mContactPointOn1, mContactPointOn2, mPenetrationAxis;
	float mPenetrationDepth = 0.0f;
	SubShapeID mSubShapeID1, mSubShapeID2;
	BodyID mBodyID2;
};

// Closest hit ray casts on the device broadphase + shapes. The reference casts one ray per call; the device wants thousands, so next
// to the reference's signature there is a batched form (the RL observation pattern: all rays of a step in one call).
class NarrowPhaseQuery
{
public:
	// NarrowPhaseQuery::CastRay(inRay, ioHit): true if the ray hits closer than ioHit.mFraction
	bool CastRay(const RRayCast &inRay, RayCastResult &ioHit) const
	{
		RayCastResult hit;
		CastRays(&inRay, 1, &hit);
		if (hit.mBodyID.IsInvalid() || !(hit.mFraction < ioHit.mFraction)) return false;
		ioHit = hit;
		return true;
	}
	// n rays at once; inObjectLayer: the layer the rays collide as (DefaultBroadPhaseLayerFilter / DefaultObjectLayerFilter of the
	// tables given to Init), cNoLayer = collide with everything (the reference's default filters)
	static constexpr uint32 cNoLayer = 0xffffffffu;
	inline void CastRays(const RRayCast *inRays, int inNumber, RayCastResult *outHits, uint32 inObjectLayer = cNoLayer) const;
	// NarrowPhaseQuery::CollideShape (NarrowPhaseQuery.h:53) with an all hits collector: inShape (a convex shape, kept alive by the
	// caller as in the reference) scaled by inShapeScale at the centre of mass transform (inRotation, inPosition); results relative to
	// inBaseOffset. The batched form takes n transforms of the same shape (one device call).
	inline void CollideShape(const Shape *inShape, const Vec3 &inShapeScale, const Quat &inRotation, const RVec3 &inPosition, const CollideShapeSettings &inSettings,
		const RVec3 &inBaseOffset, std::vector<CollideShapeResult> &outHits, uint32 inObjectLayer = cNoLayer) const;
	// (the reference's signature: the rotation of the matrix is handed to the device as a quaternion, Mat44::GetQuaternion)
	inline void CollideShape(const Shape *inShape, const Vec3 &inShapeScale, const RMat44 &inCenterOfMassTransform, const CollideShapeSettings &inSettings,
		const RVec3 &inBaseOffset, std::vector<CollideShapeResult> &outHits, uint32 inObjectLayer = cNoLayer) const;
	inline void CollideShapes(const Shape *inShape, const Vec3 &inShapeScale, const Quat *inRotations, const RVec3 *inPositions, int inNumber, const CollideShapeSettings &inSettings,
		std::vector<std::vector<CollideShapeResult>> &outHits, uint32 inObjectLayer = cNoLayer) const;
private:
	friend class PhysicsSystem;
	PhysicsSystem *mSystem = nullptr;
};

class BroadPhaseQuery
{
public:
	// BroadPhaseQuery::CollideAABox with an all hits collector: ids of the bodies whose world space bounds overlap the box
	inline void CollideAABox(const AABox &inBox, std::vector<BodyID> &outBodies, uint32 inObjectLayer = NarrowPhaseQuery::cNoLayer) const;
	// BroadPhaseQuery::CollideSphere / CollidePoint (BroadPhaseQuery.h:41,44) with an all hits collector
	inline void CollideSphere(const Vec3 &inCenter, float inRadius, std::vector<BodyID> &outBodies, uint32 inObjectLayer = NarrowPhaseQuery::cNoLayer) const;
	inline void CollidePoint(const Vec3 &inPoint, std::vector<BodyID> &outBodies, uint32 inObjectLayer = NarrowPhaseQuery::cNoLayer) const;
private:
	inline void CollideVolume(int inMode, const float *inData, std::vector<BodyID> &outBodies, uint32 inObjectLayer) const;
	friend class PhysicsSystem;
	PhysicsSystem *mSystem = nullptr;
};

// ---- listeners ------------------------------------------------------------------------------------------------------
struct ContactManifold
{
	RVec3 mBaseOffset;
	Vec3 mWorldSpaceNormal;
	float mPenetrationDepth = 0.0f;
	SubShapeID mSubShapeID1, mSubShapeID2;
	std::vector<Vec3> mRelativeContactPointsOn1, mRelativeContactPointsOn2;
};
struct ContactSettings { float mCombinedFriction = 0, mCombinedRestitution = 0; bool mIsSensor = false; };

// Host mirror of a body (what the reference hands to listeners)
class PhysicsSystem;
// StateRecorder (Jolt/Physics/StateRecorder.h) as the facade sees it: the recorded state is a snapshot that STAYS ON THE DEVICE
// (b2j_world_save_state); the recorder owns it. SaveState replaces what the recorder held.
enum class EStateRecorderState : uint8 { None = 0, Global = 1, Bodies = 2, Contacts = 4, Constraints = 8, All = 15 };
class StateRecorder
{
public:
	StateRecorder() = default;
	StateRecorder(const StateRecorder &) = delete;
	~StateRecorder() { Clear(); }
	void Clear() { if (mSnapshot != nullptr) b2j_snapshot_destroy(mSnapshot); mSnapshot = nullptr; }
	void Rewind() { }
	bool IsFailed() const { return false; }
	uint64 GetDataSize() const { return mSnapshot != nullptr? b2j_snapshot_size(mSnapshot) : 0; }
	b2j_snapshot *mSnapshot = nullptr;
};
using StateRecorderImpl = StateRecorder;
enum class EBodyType : uint8 { RigidBody, SoftBody };   // Jolt/Physics/Body/BodyType.h (soft bodies are out of scope)
using BodyIDVector = std::vector<BodyID>;

class Body
{
public:
	const BodyID &GetID() const { return mID; }
	RVec3 GetCenterOfMassPosition() const { Sync(); return mPosition; }
	Quat GetRotation() const { Sync(); return mRotation; }
	Vec3 GetLinearVelocity() const { Sync(); return mLinearVelocity; }
	Vec3 GetAngularVelocity() const { Sync(); return mAngularVelocity; }
	RVec3 GetPosition() const { Sync(); return mPosition - mRotation * mShape->GetCenterOfMass(); }
	bool IsActive() const { Sync(); return mActive; }
	bool IsStatic() const { return mMotionType == EMotionType::Static; }
	bool IsDynamic() const { return mMotionType == EMotionType::Dynamic; }
	EMotionType GetMotionType() const { return mMotionType; }
	ObjectLayer GetObjectLayer() const { return mObjectLayer; }
	uint64 GetUserData() const { return mUserData; }
	const Shape *GetShape() const { return mShape.get(); }

	// PhysicsSystem::Update downloads the state of all bodies into flat arrays with one copy; the per body mirror below is refreshed
	// from them on first access after a step (no per body work in Update: it matters at 1M bodies).
	inline void Sync() const;

	BodyID mID;
	mutable RVec3 mPosition;      // centre of mass position
	mutable Quat mRotation;
	mutable Vec3 mLinearVelocity, mAngularVelocity;
	mutable uint32 mSyncGeneration = 0;
	const PhysicsSystem *mSystem = nullptr;
	ShapeRef mShape;
	EMotionType mMotionType = EMotionType::Static;
	ObjectLayer mObjectLayer = 0;
	uint64 mUserData = 0;
	mutable bool mActive = false;
	bool mInWorld = false, mDestroyed = false;
	b2j_body_desc mDesc;          // creation time descriptor (uploaded by AddBody)
	float mDynamicInvMass = 0.0f; // MotionProperties::mInvMass (the device holds 0 while the body is not dynamic)
	bool mHasMotionProperties = false;
};

// ---- non contact constraints: PointConstraint / DistanceConstraint with the reference's settings classes ---------------
// (Jolt/Physics/Constraints/Constraint.h, TwoBodyConstraint.h, PointConstraint.h, DistanceConstraint.h). A constraint is created from
// its settings and two bodies, handed to PhysicsSystem::AddConstraint (which owns it from then on, like the reference's Ref<Constraint>)
// and lives on the device as one entry of the world's constraint list (b2j_constraints_add).
enum class EConstraintSpace { LocalToBodyCOM, WorldSpace };
enum class EConstraintSubType { Point = B2J_CONSTRAINT_POINT, Distance = B2J_CONSTRAINT_DISTANCE, Hinge = B2J_CONSTRAINT_HINGE, Fixed = B2J_CONSTRAINT_FIXED };

class Constraint
{
public:
	virtual ~Constraint() = default;
	virtual EConstraintSubType GetSubType() const = 0;
	void SetConstraintPriority(uint32 inPriority) { mDesc.priority = inPriority; Changed(); }
	uint32 GetConstraintPriority() const { return mDesc.priority; }
	void SetNumVelocityStepsOverride(uint inN) { mDesc.num_velocity_steps_override = (uint8)inN; Changed(); }
	uint GetNumVelocityStepsOverride() const { return mDesc.num_velocity_steps_override; }
	void SetNumPositionStepsOverride(uint inN) { mDesc.num_position_steps_override = (uint8)inN; Changed(); }
	uint GetNumPositionStepsOverride() const { return mDesc.num_position_steps_override; }
	inline void SetEnabled(bool inEnabled);
	bool GetEnabled() const { return mDesc.enabled != 0; }
	static constexpr uint32 cInvalidConstraintIndex = 0xffffffffu;
protected:
	friend class PhysicsSystem;
	Constraint() { memset(&mDesc, 0, sizeof(mDesc)); mDesc.enabled = 1; }
	// (priority / step overrides of a constraint that is already in a system are fixed: set them before AddConstraint, as the scenes do)
	void Changed() { }
	b2j_constraint_desc mDesc;
	uint32 mConstraintIndex = cInvalidConstraintIndex;
	PhysicsSystem *mSystem = nullptr;
};

class TwoBodyConstraint : public Constraint
{
public:
	Body *GetBody1() const { return mBody1; }
	Body *GetBody2() const { return mBody2; }
protected:
	TwoBodyConstraint(Body &inBody1, Body &inBody2) : mBody1(&inBody1), mBody2(&inBody2) { mDesc.body1 = inBody1.GetID().mID; mDesc.body2 = inBody2.GetID().mID; }
	// the constructors of PointConstraint / DistanceConstraint: world space points are taken to the space of the bodies' centres of mass
	void SetPoints(EConstraintSpace inSpace, const RVec3 &inPoint1, const RVec3 &inPoint2, RVec3 &outWorld1, RVec3 &outWorld2)
	{
		Vec3 l1, l2;
		if (inSpace == EConstraintSpace::WorldSpace)
		{
			l1 = Mat44RT::sInverseRotationTranslation(mBody1->GetRotation(), mBody1->GetCenterOfMassPosition()) * inPoint1;
			l2 = Mat44RT::sInverseRotationTranslation(mBody2->GetRotation(), mBody2->GetCenterOfMassPosition()) * inPoint2;
			outWorld1 = inPoint1; outWorld2 = inPoint2;
		}
		else
		{
			l1 = inPoint1; l2 = inPoint2;
			outWorld1 = Mat44RT::sRotationTranslation(mBody1->GetRotation(), mBody1->GetCenterOfMassPosition()) * inPoint1;
			outWorld2 = Mat44RT::sRotationTranslation(mBody2->GetRotation(), mBody2->GetCenterOfMassPosition()) * inPoint2;
		}
		mDesc.point1[0] = l1.x; mDesc.point1[1] = l1.y; mDesc.point1[2] = l1.z;
		mDesc.point2[0] = l2.x; mDesc.point2[1] = l2.y; mDesc.point2[2] = l2.z;
	}
	Body *mBody1, *mBody2;
};

class TwoBodyConstraintSettings
{
public:
	virtual ~TwoBodyConstraintSettings() = default;
	virtual TwoBodyConstraint *Create(Body &inBody1, Body &inBody2) const = 0;
	uint32 mConstraintPriority = 0;
	uint mNumVelocityStepsOverride = 0, mNumPositionStepsOverride = 0;
	bool mEnabled = true;
};

class PointConstraintSettings;
class PointConstraint final : public TwoBodyConstraint
{
public:
	inline PointConstraint(Body &inBody1, Body &inBody2, const PointConstraintSettings &inSettings);
	EConstraintSubType GetSubType() const override { return EConstraintSubType::Point; }
	Vec3 GetLocalSpacePoint1() const { return Vec3(mDesc.point1[0], mDesc.point1[1], mDesc.point1[2]); }
	Vec3 GetLocalSpacePoint2() const { return Vec3(mDesc.point2[0], mDesc.point2[1], mDesc.point2[2]); }
	inline Vec3 GetTotalLambdaPosition() const;
};
class PointConstraintSettings final : public TwoBodyConstraintSettings
{
public:
	TwoBodyConstraint *Create(Body &inBody1, Body &inBody2) const override { return new PointConstraint(inBody1, inBody2, *this); }
	EConstraintSpace mSpace = EConstraintSpace::WorldSpace;
	RVec3 mPoint1 = RVec3::sZero(), mPoint2 = RVec3::sZero();
};

class DistanceConstraintSettings;
class DistanceConstraint final : public TwoBodyConstraint
{
public:
	inline DistanceConstraint(Body &inBody1, Body &inBody2, const DistanceConstraintSettings &inSettings);
	EConstraintSubType GetSubType() const override { return EConstraintSubType::Distance; }
	float GetMinDistance() const { return mDesc.min_distance; }
	float GetMaxDistance() const { return mDesc.max_distance; }
	inline float GetTotalLambdaPosition() const;
};
class DistanceConstraintSettings final : public TwoBodyConstraintSettings
{
public:
	TwoBodyConstraint *Create(Body &inBody1, Body &inBody2) const override { return new DistanceConstraint(inBody1, inBody2, *this); }
	EConstraintSpace mSpace = EConstraintSpace::WorldSpace;
	RVec3 mPoint1 = RVec3::sZero(), mPoint2 = RVec3::sZero();
	float mMinDistance = -1.0f, mMaxDistance = -1.0f; // < 0: the distance at creation (DistanceConstraint.cpp:73-84)
};

inline PointConstraint::PointConstraint(Body &inBody1, Body &inBody2, const PointConstraintSettings &inSettings) : TwoBodyConstraint(inBody1, inBody2)
{
	mDesc.type = B2J_CONSTRAINT_POINT;
	mDesc.priority = inSettings.mConstraintPriority; mDesc.enabled = inSettings.mEnabled;
	mDesc.num_velocity_steps_override = (uint8)inSettings.mNumVelocityStepsOverride; mDesc.num_position_steps_override = (uint8)inSettings.mNumPositionStepsOverride;
	RVec3 w1, w2;
	SetPoints(inSettings.mSpace, inSettings.mPoint1, inSettings.mPoint2, w1, w2);
}

inline DistanceConstraint::DistanceConstraint(Body &inBody1, Body &inBody2, const DistanceConstraintSettings &inSettings) : TwoBodyConstraint(inBody1, inBody2)
{
	mDesc.type = B2J_CONSTRAINT_DISTANCE;
	mDesc.priority = inSettings.mConstraintPriority; mDesc.enabled = inSettings.mEnabled;
	mDesc.num_velocity_steps_override = (uint8)inSettings.mNumVelocityStepsOverride; mDesc.num_position_steps_override = (uint8)inSettings.mNumPositionStepsOverride;
	RVec3 w1, w2;
	SetPoints(inSettings.mSpace, inSettings.mPoint1, inSettings.mPoint2, w1, w2);
	// DistanceConstraint.cpp:73-84
	float distance = (w2 - w1).Length();
	float mn = inSettings.mMinDistance, mx = inSettings.mMaxDistance;
	if (mn < 0.0f && mx < 0.0f) { mDesc.min_distance = distance; mDesc.max_distance = distance; }
	else
	{
		mDesc.min_distance = mn < 0.0f? std::min(distance, mx) : mn;
		mDesc.max_distance = mx < 0.0f? std::max(distance, mn) : mx;
	}
}

// HingeConstraint (HingeConstraint.h): rotation about one axis with optional angle limits and friction; the motor stays off and the
// limits have no spring on this path
class HingeConstraintSettings;
class HingeConstraint final : public TwoBodyConstraint
{
public:
	inline HingeConstraint(Body &inBody1, Body &inBody2, const HingeConstraintSettings &inSettings);
	EConstraintSubType GetSubType() const override { return EConstraintSubType::Hinge; }
	float GetLimitsMin() const { return mDesc.limits_min; }
	float GetLimitsMax() const { return mDesc.limits_max; }
	float GetMaxFrictionTorque() const { return mDesc.max_friction_torque; }
	Vec3 GetLocalSpaceHingeAxis1() const { return Vec3(mDesc.hinge_axis1[0], mDesc.hinge_axis1[1], mDesc.hinge_axis1[2]); }
	Vec3 GetLocalSpaceHingeAxis2() const { return Vec3(mDesc.hinge_axis2[0], mDesc.hinge_axis2[1], mDesc.hinge_axis2[2]); }
};
class HingeConstraintSettings final : public TwoBodyConstraintSettings
{
public:
	TwoBodyConstraint *Create(Body &inBody1, Body &inBody2) const override { return new HingeConstraint(inBody1, inBody2, *this); }
	EConstraintSpace mSpace = EConstraintSpace::WorldSpace;
	RVec3 mPoint1 = RVec3::sZero(), mPoint2 = RVec3::sZero();
	Vec3 mHingeAxis1 = Vec3::sAxisY(), mNormalAxis1 = Vec3::sAxisX(), mHingeAxis2 = Vec3::sAxisY(), mNormalAxis2 = Vec3::sAxisX();
	float mLimitsMin = -3.14159265358979323846f, mLimitsMax = 3.14159265358979323846f;
	float mMaxFrictionTorque = 0.0f;
};

inline HingeConstraint::HingeConstraint(Body &inBody1, Body &inBody2, const HingeConstraintSettings &inSettings) : TwoBodyConstraint(inBody1, inBody2)
{
	mDesc.type = B2J_CONSTRAINT_HINGE;
	mDesc.priority = inSettings.mConstraintPriority; mDesc.enabled = inSettings.mEnabled;
	mDesc.num_velocity_steps_override = (uint8)inSettings.mNumVelocityStepsOverride; mDesc.num_position_steps_override = (uint8)inSettings.mNumPositionStepsOverride;
	mDesc.limits_min = inSettings.mLimitsMin; mDesc.limits_max = inSettings.mLimitsMax; mDesc.max_friction_torque = inSettings.mMaxFrictionTorque;
	// RotationEulerConstraintPart::sGetInvInitialOrientationXZ(normal 1, hinge 1, normal 2, hinge 2)
	Quat inv_initial = Quat::sIdentity();
	if (!(inSettings.mNormalAxis1 == inSettings.mNormalAxis2 && inSettings.mHingeAxis1 == inSettings.mHingeAxis2))
	{
		Mat44RT constraint1, constraint2;
		constraint1.c0 = inSettings.mNormalAxis1; constraint1.c1 = inSettings.mHingeAxis1.Cross(inSettings.mNormalAxis1); constraint1.c2 = inSettings.mHingeAxis1;
		constraint2.c0 = inSettings.mNormalAxis2; constraint2.c1 = inSettings.mHingeAxis2.Cross(inSettings.mNormalAxis2); constraint2.c2 = inSettings.mHingeAxis2;
		Quat q1 = constraint1.GetQuaternion();
		inv_initial = constraint2.GetQuaternion() * Quat(-q1.x, -q1.y, -q1.z, q1.w);
	}
	RVec3 w1, w2;
	SetPoints(inSettings.mSpace, inSettings.mPoint1, inSettings.mPoint2, w1, w2);
	Vec3 a1 = inSettings.mHingeAxis1, a2 = inSettings.mHingeAxis2;
	if (inSettings.mSpace == EConstraintSpace::WorldSpace)
	{
		Quat r1 = inBody1.GetRotation(), r2 = inBody2.GetRotation();
		a1 = Mat44RT::sInverseRotationTranslation(r1, inBody1.GetCenterOfMassPosition()).Multiply3x3(inSettings.mHingeAxis1).Normalized();
		a2 = Mat44RT::sInverseRotationTranslation(r2, inBody2.GetCenterOfMassPosition()).Multiply3x3(inSettings.mHingeAxis2).Normalized();
		inv_initial = (Quat(-r2.x, -r2.y, -r2.z, r2.w) * inv_initial) * r1;
	}
	mDesc.hinge_axis1[0] = a1.x; mDesc.hinge_axis1[1] = a1.y; mDesc.hinge_axis1[2] = a1.z;
	mDesc.hinge_axis2[0] = a2.x; mDesc.hinge_axis2[1] = a2.y; mDesc.hinge_axis2[2] = a2.z;
	mDesc.inv_initial_orientation[0] = inv_initial.x; mDesc.inv_initial_orientation[1] = inv_initial.y; mDesc.inv_initial_orientation[2] = inv_initial.z; mDesc.inv_initial_orientation[3] = inv_initial.w;
}

// FixedConstraint (FixedConstraint.h): two bodies welded at a point, no relative rotation
class FixedConstraintSettings;
class FixedConstraint final : public TwoBodyConstraint
{
public:
	inline FixedConstraint(Body &inBody1, Body &inBody2, const FixedConstraintSettings &inSettings);
	EConstraintSubType GetSubType() const override { return EConstraintSubType::Fixed; }
};
class FixedConstraintSettings final : public TwoBodyConstraintSettings
{
public:
	TwoBodyConstraint *Create(Body &inBody1, Body &inBody2) const override { return new FixedConstraint(inBody1, inBody2, *this); }
	EConstraintSpace mSpace = EConstraintSpace::WorldSpace;
	bool mAutoDetectPoint = false;
	RVec3 mPoint1 = RVec3::sZero(), mPoint2 = RVec3::sZero();
	Vec3 mAxisX1 = Vec3::sAxisX(), mAxisY1 = Vec3::sAxisY(), mAxisX2 = Vec3::sAxisX(), mAxisY2 = Vec3::sAxisY();
};

inline FixedConstraint::FixedConstraint(Body &inBody1, Body &inBody2, const FixedConstraintSettings &inSettings) : TwoBodyConstraint(inBody1, inBody2)
{
	mDesc.type = B2J_CONSTRAINT_FIXED;
	mDesc.priority = inSettings.mConstraintPriority; mDesc.enabled = inSettings.mEnabled;
	mDesc.num_velocity_steps_override = (uint8)inSettings.mNumVelocityStepsOverride; mDesc.num_position_steps_override = (uint8)inSettings.mNumPositionStepsOverride;
	// RotationEulerConstraintPart::sGetInvInitialOrientationXY
	Quat inv_initial = Quat::sIdentity();
	if (!(inSettings.mAxisX1 == inSettings.mAxisX2 && inSettings.mAxisY1 == inSettings.mAxisY2))
	{
		Mat44RT constraint1, constraint2;
		constraint1.c0 = inSettings.mAxisX1; constraint1.c1 = inSettings.mAxisY1; constraint1.c2 = inSettings.mAxisX1.Cross(inSettings.mAxisY1);
		constraint2.c0 = inSettings.mAxisX2; constraint2.c1 = inSettings.mAxisY2; constraint2.c2 = inSettings.mAxisX2.Cross(inSettings.mAxisY2);
		Quat q1 = constraint1.GetQuaternion();
		inv_initial = constraint2.GetQuaternion() * Quat(-q1.x, -q1.y, -q1.z, q1.w);
	}
	RVec3 p1 = inSettings.mPoint1, p2 = inSettings.mPoint2, w1, w2;
	if (inSettings.mSpace == EConstraintSpace::WorldSpace)
	{
		if (inSettings.mAutoDetectPoint)
		{
			// the anchor: the centre of mass of the body that can move, or the mass weighted average (FixedConstraint.cpp:44-66)
			RVec3 anchor;
			if (!inBody1.mHasMotionProperties) anchor = inBody2.GetCenterOfMassPosition();
			else if (!inBody2.mHasMotionProperties) anchor = inBody1.GetCenterOfMassPosition();
			else
			{
				float inv_m1 = inBody1.mDynamicInvMass, inv_m2 = inBody2.mDynamicInvMass; // MotionProperties::GetInverseMassUnchecked
				float total_inv_mass = inv_m1 + inv_m2;
				if (total_inv_mass != 0.0f)
				{
					RVec3 sum = inv_m1 * inBody1.GetCenterOfMassPosition() + inv_m2 * inBody2.GetCenterOfMassPosition();
					float div = inv_m1 + inv_m2;
					anchor = RVec3(sum.x / div, sum.y / div, sum.z / div);
				}
				else
					anchor = inBody1.GetCenterOfMassPosition();
			}
			p1 = anchor; p2 = anchor;
		}
		Quat r1 = inBody1.GetRotation(), r2 = inBody2.GetRotation();
		inv_initial = (Quat(-r2.x, -r2.y, -r2.z, r2.w) * inv_initial) * r1;
	}
	SetPoints(inSettings.mSpace, p1, p2, w1, w2);
	mDesc.inv_initial_orientation[0] = inv_initial.x; mDesc.inv_initial_orientation[1] = inv_initial.y; mDesc.inv_initial_orientation[2] = inv_initial.z; mDesc.inv_initial_orientation[3] = inv_initial.w;
}

class ContactListener
{
public:
	virtual ~ContactListener() = default;
	virtual void OnContactAdded(const Body &, const Body &, const ContactManifold &, ContactSettings &) { }
	virtual void OnContactPersisted(const Body &, const Body &, const ContactManifold &, ContactSettings &) { }
	virtual void OnContactRemoved(const SubShapeIDPair &) { }
};

class BodyActivationListener
{
public:
	virtual ~BodyActivationListener() = default;
	virtual void OnBodyActivated(const BodyID &, uint64) = 0;
	virtual void OnBodyDeactivated(const BodyID &, uint64) = 0;
};

// ---- BodyInterface ----------------------------------------------------------------------------------------------------
class BodyInterface
{
public:
	// BodyInterface::CreateBody (BodyInterface.cpp:30-39): nullptr when out of bodies
	Body *CreateBody(const BodyCreationSettings &inSettings);
	void AddBody(const BodyID &inBodyID, EActivation inActivationMode);
	BodyID CreateAndAddBody(const BodyCreationSettings &inSettings, EActivation inActivationMode)
	{
		Body *b = CreateBody(inSettings);
		if (b == nullptr) return BodyID();
		AddBody(b->GetID(), inActivationMode);
		return b->GetID();
	}
	// Bulk add (AddBodiesPrepare / Finalize / Abort, BodyInterface.h:124-133): one upload for all bodies. The reference prepares a
	// broadphase sub tree in Prepare; here the state handle only remembers the batch and Finalize queues it.
	using AddState = void *;
	AddState AddBodiesPrepare(BodyID *ioBodies, int inNumber) { (void)ioBodies; return inNumber > 0? (AddState)this : nullptr; }
	void AddBodiesFinalize(BodyID *ioBodies, int inNumber, AddState inAddState, EActivation inActivationMode) { if (inAddState != nullptr) AddBodies(ioBodies, inNumber, inActivationMode); }
	void AddBodiesAbort(BodyID *ioBodies, int inNumber, AddState inAddState) { (void)ioBodies; (void)inNumber; (void)inAddState; }
	void AddBodies(const BodyID *inBodies, int inNumber, EActivation inActivationMode);
	void RemoveBody(const BodyID &inBodyID);
	void RemoveBodies(BodyID *ioBodies, int inNumber);
	void DestroyBody(const BodyID &inBodyID);
	void ActivateBody(const BodyID &inBodyID) { SetActive(inBodyID, true); }
	void DeactivateBody(const BodyID &inBodyID) { SetActive(inBodyID, false); }
	bool IsActive(const BodyID &inBodyID) const { const Body *b = TryGet(inBodyID); return b != nullptr && b->IsActive(); }
	bool IsAdded(const BodyID &inBodyID) const { const Body *b = TryGet(inBodyID); return b != nullptr && b->mInWorld; }

	RVec3 GetPosition(const BodyID &id) const { const Body *b = TryGet(id); return b? b->GetPosition() : RVec3::sZero(); }
	// The state getters first look in the flat arrays PhysicsSystem::Update downloaded (no Body object is touched: it matters when a
	// caller reads a million bodies per step); bodies changed through the interface since then take the Body mirror path.
	inline RVec3 GetCenterOfMassPosition(const BodyID &id) const;
	inline Quat GetRotation(const BodyID &id) const;
	inline Vec3 GetLinearVelocity(const BodyID &id) const;
	inline Vec3 GetAngularVelocity(const BodyID &id) const;
	void GetPositionAndRotation(const BodyID &id, RVec3 &outPosition, Quat &outRotation) const { outPosition = GetPosition(id); outRotation = GetRotation(id); }
	void SetPositionAndRotation(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, EActivation inActivationMode);
	void SetLinearAndAngularVelocity(const BodyID &id, const Vec3 &inLinearVelocity, const Vec3 &inAngularVelocity);
	void SetLinearVelocity(const BodyID &id, const Vec3 &v) { SetLinearAndAngularVelocity(id, v, GetAngularVelocity(id)); }
	void SetAngularVelocity(const BodyID &id, const Vec3 &v) { SetLinearAndAngularVelocity(id, GetLinearVelocity(id), v); }
	// the rest of the pose / velocity surface (BodyInterface.h:187-216), host side arithmetic of Body / MotionProperties on the mirror
	void SetPosition(const BodyID &id, const RVec3 &inPosition, EActivation inActivationMode) { SetPositionAndRotation(id, inPosition, GetRotation(id), inActivationMode); }
	void SetRotation(const BodyID &id, const Quat &inRotation, EActivation inActivationMode) { SetPositionAndRotation(id, GetPosition(id), inRotation, inActivationMode); }
	void SetPositionAndRotationWhenChanged(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, EActivation inActivationMode);
	void SetPositionRotationAndVelocity(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, const Vec3 &inLinearVelocity, const Vec3 &inAngularVelocity);
	void GetLinearAndAngularVelocity(const BodyID &id, Vec3 &outLinearVelocity, Vec3 &outAngularVelocity) const { outLinearVelocity = GetLinearVelocity(id); outAngularVelocity = GetAngularVelocity(id); }
	void AddLinearVelocity(const BodyID &id, const Vec3 &inLinearVelocity) { SetLinearAndAngularVelocity(id, GetLinearVelocity(id) + inLinearVelocity, GetAngularVelocity(id)); }
	void AddLinearAndAngularVelocity(const BodyID &id, const Vec3 &inLinearVelocity, const Vec3 &inAngularVelocity) { SetLinearAndAngularVelocity(id, GetLinearVelocity(id) + inLinearVelocity, GetAngularVelocity(id) + inAngularVelocity); }
	Vec3 GetPointVelocity(const BodyID &id, const RVec3 &inPoint) const; // Body::GetPointVelocity: v + w x (point - centre of mass)
	void MoveKinematic(const BodyID &id, const RVec3 &inTargetPosition, const Quat &inTargetRotation, float inDeltaTime);
	Mat44RT GetWorldTransform(const BodyID &id) const { return Mat44RT::sRotationTranslation(GetRotation(id), GetPosition(id)); }
	Mat44RT GetCenterOfMassTransform(const BodyID &id) const { return Mat44RT::sRotationTranslation(GetRotation(id), GetCenterOfMassPosition(id)); }
	Mat44RT GetInverseInertia(const BodyID &id) const; // Body::GetInverseInertia (world space, locked DOFs masked; translation part unused)
	void AddForce(const BodyID &id, const Vec3 &inForce, const RVec3 &inPoint, EActivation inActivationMode = EActivation::Activate);
	void AddForceAndTorque(const BodyID &id, const Vec3 &inForce, const Vec3 &inTorque, EActivation inActivationMode = EActivation::Activate) { AddForce(id, inForce, inActivationMode); AddTorque(id, inTorque, inActivationMode); }
	void ActivateBodies(const BodyID *inBodyIDs, int inNumber) { for (int i = 0; i < inNumber; ++i) ActivateNoReset(inBodyIDs[i]); } // BodyManager::ActivateBodies: sleeping bodies only
	void ResetSleepTimer(const BodyID &id) { const Body *b = TryGet(id); if (b == nullptr || !b->mInWorld || b->mMotionType == EMotionType::Static) return; Flush(); uint32 bid = id.mID; b2j_bodies_reset_sleep_timer(World(), &bid, 1); }
	void DeactivateBodies(const BodyID *inBodyIDs, int inNumber) { for (int i = 0; i < inNumber; ++i) DeactivateBody(inBodyIDs[i]); }
	void DestroyBodies(const BodyID *inBodyIDs, int inNumber) { for (int i = 0; i < inNumber; ++i) DestroyBody(inBodyIDs[i]); }
	EMotionQuality GetMotionQuality(const BodyID &) const { return EMotionQuality::Discrete; } // (LinearCast is not on the path: SURVEY 8b)
	// Body flags on bodies that exist (BodyInterface.h:286-293, Body::SetIsSensor / SetUseManifoldReduction / SetAllowSleeping ...)
	void SetIsSensor(const BodyID &id, bool inIsSensor) { SetFlag(id, B2J_BODY_SENSOR, inIsSensor); }
	bool IsSensor(const BodyID &id) const { const Body *b = TryGet(id); return b != nullptr && (b->mDesc.flags & B2J_BODY_SENSOR) != 0; }
	void SetUseManifoldReduction(const BodyID &id, bool inUseReduction) { SetFlag(id, B2J_BODY_USE_MANIFOLD_REDUCTION, inUseReduction); }
	bool GetUseManifoldReduction(const BodyID &id) const { const Body *b = TryGet(id); return b != nullptr && (b->mDesc.flags & B2J_BODY_USE_MANIFOLD_REDUCTION) != 0; }
	void SetFlag(const BodyID &id, uint16_t inFlag, bool inValue)
	{
		Body *b = const_cast<Body *>(TryGet(id));
		if (b == nullptr) return;
		if (inValue) b->mDesc.flags |= inFlag; else b->mDesc.flags &= (uint16_t)~inFlag;
		if (!b->mInWorld) return;
		Flush();
		uint32 bid = id.mID;
		b2j_body_info_update u;
		memset(&u, 0, sizeof(u));
		uint16_t bits = inFlag;
		if (inValue) u.flags_set = &bits; else u.flags_clear = &bits;
		b2j_bodies_set_info(World(), &bid, 1, &u);
	}
	void SetUserData(const BodyID &id, uint64 inUserData) const { Body *b = const_cast<Body *>(TryGet(id)); if (b != nullptr) b->mUserData = inUserData; }
	void AddForce(const BodyID &id, const Vec3 &inForce, EActivation inActivationMode = EActivation::Activate);
	void AddTorque(const BodyID &id, const Vec3 &inTorque, EActivation inActivationMode = EActivation::Activate);
	// BodyInterface::AddImpulse / AddAngularImpulse (BodyInterface.h:227-230, Body.inl AddImpulse): dynamic bodies only, activates the body
	void AddImpulse(const BodyID &id, const Vec3 &inImpulse);
	void AddImpulse(const BodyID &id, const Vec3 &inImpulse, const RVec3 &inPoint);
	void AddAngularImpulse(const BodyID &id, const Vec3 &inAngularImpulse);
	// per body material / motion parameters (BodyInterface.h:241-281)
	void SetFriction(const BodyID &id, float v) { SetParam(id, &b2j_body_desc::friction, &b2j_body_params::friction, v); }
	float GetFriction(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mDesc.friction : 0.0f; }
	void SetRestitution(const BodyID &id, float v) { SetParam(id, &b2j_body_desc::restitution, &b2j_body_params::restitution, v); }
	float GetRestitution(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mDesc.restitution : 0.0f; }
	void SetGravityFactor(const BodyID &id, float v) { SetParam(id, &b2j_body_desc::gravity_factor, &b2j_body_params::gravity_factor, v); }
	float GetGravityFactor(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mDesc.gravity_factor : 1.0f; }
	void SetMaxLinearVelocity(const BodyID &id, float v) { SetParam(id, &b2j_body_desc::max_linear_velocity, &b2j_body_params::max_linear_velocity, v); }
	float GetMaxLinearVelocity(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mDesc.max_linear_velocity : 0.0f; }
	void SetMaxAngularVelocity(const BodyID &id, float v) { SetParam(id, &b2j_body_desc::max_angular_velocity, &b2j_body_params::max_angular_velocity, v); }
	float GetMaxAngularVelocity(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mDesc.max_angular_velocity : 0.0f; }
	EMotionType GetMotionType(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mMotionType : EMotionType::Static; }
	ObjectLayer GetObjectLayer(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mObjectLayer : 0; }
	// BodyInterface::SetMotionType / SetObjectLayer / SetShape / InvalidateContactCache (BodyInterface.h:241,181,169,300)
	void SetMotionType(const BodyID &id, EMotionType inMotionType, EActivation inActivationMode);
	void SetObjectLayer(const BodyID &id, ObjectLayer inLayer);
	void SetShape(const BodyID &id, const ShapeRef &inShape, bool inUpdateMassProperties, EActivation inActivationMode);
	void InvalidateContactCache(const BodyID &id);
	// Bulk force application from host arrays (n bodies, force/torque [n][3], either may be null): the RL pattern
	void AddForcesAndTorques(const BodyID *inBodies, int inNumber, const float *inForces, const float *inTorques);
	uint64 GetUserData(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mUserData : 0; }
	const Shape *GetShape(const BodyID &id) const { const Body *b = TryGet(id); return b? b->mShape.get() : nullptr; }
	const Body *TryGet(const BodyID &id) const;

private:
	friend class PhysicsSystem;
	friend class NarrowPhaseQuery;
	friend class BroadPhaseQuery;
	b2j_world *World() const;
	void Flush();
	void SetActive(const BodyID &id, bool inActive);
	void ActivateNoReset(const BodyID &id) { const Body *b = TryGet(id); if (b != nullptr && !b->IsActive()) SetActive(id, true); } // (!IsActive -> BodyManager::ActivateBodies)
	void SetParam(const BodyID &id, float b2j_body_desc::*inDescMember, const float *b2j_body_params::*inParamMember, float inValue);
	PhysicsSystem *mSystem = nullptr;
};

// ---- PhysicsSystem ----------------------------------------------------------------------------------------------------
class PhysicsSystem
{
public:
	PhysicsSystem() { mBodyInterface.mSystem = this; mNarrowPhaseQuery.mSystem = this; mBroadPhaseQuery.mSystem = this; }
	// PhysicsSystem::GetNarrowPhaseQuery / GetBroadPhaseQuery (PhysicsSystem.h:121-130)
	const NarrowPhaseQuery &GetNarrowPhaseQuery() const { return mNarrowPhaseQuery; }
	const NarrowPhaseQuery &GetNarrowPhaseQueryNoLock() const { return mNarrowPhaseQuery; }
	const BroadPhaseQuery &GetBroadPhaseQuery() const { return mBroadPhaseQuery; }
	~PhysicsSystem()
	{
		if (mStateRegistered) { b2j_host_buffer_unregister(mPos.data()); b2j_host_buffer_unregister(mRot.data()); b2j_host_buffer_unregister(mLin.data()); b2j_host_buffer_unregister(mAng.data()); b2j_host_buffer_unregister(mActiveIndex.data()); }
		for (Constraint *c : mConstraints) delete c;
		if (mWorld) b2j_world_destroy(mWorld);
	}
	PhysicsSystem(const PhysicsSystem &) = delete;

	// PhysicsSystem::Init (PhysicsSystem.h:59). inNumBodyMutexes is ignored (the GPU owns the bodies during a step).
	// inNumObjectLayers tells how many object layers to sample the virtual filters for (the reference has no such query).
	bool Init(uint inMaxBodies, uint inNumBodyMutexes, uint inMaxBodyPairs, uint inMaxContactConstraints, const BroadPhaseLayerInterface &inBroadPhaseLayerInterface,
		const ObjectVsBroadPhaseLayerFilter &inObjectVsBroadPhaseLayerFilter, const ObjectLayerPairFilter &inObjectLayerPairFilter, uint inNumObjectLayers = 2, int inDevice = 0)
	{
		(void)inNumBodyMutexes;
		uint nb = inBroadPhaseLayerInterface.GetNumBroadPhaseLayers();
		std::vector<uint8_t> o2bp(inNumObjectLayers), ovbp(inNumObjectLayers * nb), ovo(inNumObjectLayers * inNumObjectLayers);
		for (uint o = 0; o < inNumObjectLayers; ++o)
		{
			o2bp[o] = (uint8_t)(BroadPhaseLayer::Type)inBroadPhaseLayerInterface.GetBroadPhaseLayer((ObjectLayer)o);
			for (uint b = 0; b < nb; ++b) ovbp[o * nb + b] = inObjectVsBroadPhaseLayerFilter.ShouldCollide((ObjectLayer)o, BroadPhaseLayer((BroadPhaseLayer::Type)b));
			for (uint o2 = 0; o2 < inNumObjectLayers; ++o2) ovo[o * inNumObjectLayers + o2] = inObjectLayerPairFilter.ShouldCollide((ObjectLayer)o, (ObjectLayer)o2);
		}
		b2j_world_desc desc;
		memset(&desc, 0, sizeof(desc));
		desc.max_bodies = inMaxBodies; desc.max_body_pairs = inMaxBodyPairs; desc.max_contact_constraints = inMaxContactConstraints;
		desc.num_object_layers = inNumObjectLayers; desc.num_broadphase_layers = nb;
		desc.object_to_broadphase = o2bp.data(); desc.object_vs_broadphase = ovbp.data(); desc.object_vs_object = ovo.data();
		FillSettings(desc.settings);
		desc.gravity[0] = mGravity.x; desc.gravity[1] = mGravity.y; desc.gravity[2] = mGravity.z;
		desc.device = inDevice;
		mWorld = b2j_world_create(&desc);
		SyncEventRecording();
		mBodies.clear();
		mBodies.reserve(1024);
		mMaxBodies = inMaxBodies;
		return mWorld != nullptr;
	}

	void SetGravity(const Vec3 &g) { mGravity = g; if (mWorld) { float v[3] = { g.x, g.y, g.z }; b2j_world_set_gravity(mWorld, v); } }
	Vec3 GetGravity() const { return mGravity; }
	void SetPhysicsSettings(const PhysicsSettings &s) { mSettings = s; if (mWorld) { b2j_settings bs; FillSettings(bs); b2j_world_set_settings(mWorld, &bs); } }
	const PhysicsSettings &GetPhysicsSettings() const { return mSettings; }
	void SetContactListener(ContactListener *l) { mContactListener = l; SyncEventRecording(); }
	ContactListener *GetContactListener() const { return mContactListener; }
	void SetBodyActivationListener(BodyActivationListener *l) { mActivationListener = l; SyncEventRecording(); }
	BodyActivationListener *GetBodyActivationListener() const { return mActivationListener; }
	BodyInterface &GetBodyInterface() { return mBodyInterface; }
	BodyInterface &GetBodyInterfaceNoLock() { return mBodyInterface; }
	void OptimizeBroadPhase() { } // the device broadphase is rebuilt from scratch every step
	uint GetNumBodies() const { return mWorld? b2j_num_bodies(mWorld) : 0; }
	uint GetNumActiveBodies() const { return mWorld? b2j_num_active_bodies(mWorld) : 0; }
	uint GetMaxBodies() const { return mMaxBodies; }
	// PhysicsSystem::GetBodies / GetActiveBodies (PhysicsSystem.h:226-240); active bodies come back in the device's active list order
	void GetBodies(BodyIDVector &outBodyIDs) const
	{
		outBodyIDs.clear();
		for (const std::unique_ptr<Body> &b : mBodies) if (b && !b->mDestroyed) outBodyIDs.push_back(b->mID); // BodyManager::GetBodyIDs: every created body, added or not
	}
	void GetActiveBodies(EBodyType inType, BodyIDVector &outBodyIDs) const
	{
		outBodyIDs.clear();
		if (inType != EBodyType::RigidBody || mWorld == nullptr) return;
		const_cast<PhysicsSystem *>(this)->mBodyInterface.Flush();
		std::vector<uint32> ids(b2j_num_active_bodies(mWorld));
		uint32 n = ids.empty()? 0 : b2j_get_active_bodies(mWorld, ids.data(), (uint32)ids.size());
		for (uint32 i = 0; i < n; ++i) { BodyID id; id.mID = ids[i]; outBodyIDs.push_back(id); }
	}
	bool WereBodiesInContact(const BodyID &a, const BodyID &b) const { return mWorld && b2j_were_bodies_in_contact(mWorld, a.mID, b.mID) == 1; }
	const b2j_step_stats &GetLastStepStats() const { return mStats; }
	// (diagnostics / tests: how many body rows the last refresh of the host mirror fetched -- see EnsureState)
	uint32 GetLastDownloadCount() const { EnsureState(); return mLastDownloadCount; }
	// the C ABI handle (everything added through the interface so far is on the device when this returns)
	b2j_world *GetWorld() const { PhysicsSystem *self = const_cast<PhysicsSystem *>(this); self->mBodyInterface.Flush(); self->FlushConstraints(); return mWorld; }
	const char *GetLastError() const { return b2j_last_error(); }

	// PhysicsSystem::SaveState / RestoreState (PhysicsSystem.h:165-168): always the whole state (Global | Bodies | Contacts); filters
	// are not supported. As in the reference the same Body objects must exist at restore time.
	void SaveState(StateRecorder &inStream, EStateRecorderState = EStateRecorderState::All) const
	{
		const_cast<PhysicsSystem *>(this)->mBodyInterface.Flush();
		inStream.Clear();
		inStream.mSnapshot = b2j_world_save_state(mWorld);
	}
	bool RestoreState(StateRecorder &inStream)
	{
		if (inStream.mSnapshot == nullptr) return false;
		mBodyInterface.Flush();
		if (b2j_world_restore_state(mWorld, inStream.mSnapshot) != 0) return false;
		DownloadState(); // the Body mirrors follow the restored device state (full download)
		return true;
	}

	// PhysicsSystem::AddConstraint(s) / RemoveConstraint(s) / GetConstraints (PhysicsSystem.h:128-140). The system owns a constraint from
	// AddConstraint to RemoveConstraint (which deletes it: the facade has no reference counting).
	void AddConstraint(Constraint *inConstraint) { AddConstraints(&inConstraint, 1); }
	void AddConstraints(Constraint **inConstraints, int inNumber)
	{
		for (int i = 0; i < inNumber; ++i)
		{
			Constraint *c = inConstraints[i];
			c->mSystem = this;
			c->mConstraintIndex = (uint32)mConstraints.size(); // ConstraintManager::Add
			mConstraints.push_back(c);
		}
	}
	void RemoveConstraint(Constraint *inConstraint) { RemoveConstraints(&inConstraint, 1); }
	void RemoveConstraints(Constraint **inConstraints, int inNumber)
	{
		for (int i = 0; i < inNumber; ++i)
		{
			Constraint *c = inConstraints[i];
			uint32 index = c->mConstraintIndex, last = (uint32)mConstraints.size() - 1;
			if (index < mNumUploadedConstraints)
			{
				// on the device already: flush the additions first so that both lists hold the same entries, then remove there too
				FlushConstraints();
				b2j_constraints_remove(mWorld, &index, 1);
				--mNumUploadedConstraints;
			}
			// ConstraintManager::Remove: the last constraint takes the freed index
			if (index < last) { mConstraints[index] = mConstraints[last]; mConstraints[index]->mConstraintIndex = index; }
			mConstraints.pop_back();
			delete c;
		}
	}
	const std::vector<Constraint *> &GetConstraints() const { return mConstraints; }
	// constraints added since the last flush go to the device after their bodies (BodyInterface::Flush)
	void FlushConstraints()
	{
		if (mNumUploadedConstraints == mConstraints.size()) return;
		mBodyInterface.Flush();
		std::vector<b2j_constraint_desc> descs;
		for (size_t i = mNumUploadedConstraints; i < mConstraints.size(); ++i) descs.push_back(mConstraints[i]->mDesc);
		if (b2j_constraints_add(mWorld, descs.data(), (uint32)descs.size()) == 0) mNumUploadedConstraints = (uint32)mConstraints.size();
	}

	// PhysicsSystem::Update (PhysicsSystem.h:162): uploads pending API mutations, runs the step on the GPU, mirrors the body state
	// back to the host and replays contact / activation events into the listeners on the calling thread.
	EPhysicsUpdateError Update(float inDeltaTime, int inCollisionSteps, TempAllocator *, JobSystem *)
	{
		mBodyInterface.Flush();
		FlushConstraints();
		int r = b2j_step(mWorld, inDeltaTime, inCollisionSteps, &mStats);
		if (r < 0)
			return EPhysicsUpdateError(0x80000000u);
		// the host mirror of the body state is refreshed on first use after the step (EnsureState), and only for the bodies the step
		// simulated unless something else changed the device state since the last download
		++mStateGeneration; // Body::Sync picks the new state up on first access
		for (uint8 &f : mSlotFlags) f &= 1;
		MarkStateStale();
		ReplayEvents();
		return EPhysicsUpdateError(r);
	}

private:
	friend class BodyInterface;
	friend class Body;
	friend class NarrowPhaseQuery;
	friend class BroadPhaseQuery;

	// the device records contact / activation events only while a listener is attached (no event traffic otherwise)
	void SyncEventRecording() { if (mWorld) b2j_world_set_event_recording(mWorld, mContactListener != nullptr, mActivationListener != nullptr); }

	void FillSettings(b2j_settings &s) const
	{
		b2j_settings_default(&s);
		s.baumgarte = mSettings.mBaumgarte; s.speculative_contact_distance = mSettings.mSpeculativeContactDistance; s.penetration_slop = mSettings.mPenetrationSlop;
		s.max_penetration_distance = mSettings.mMaxPenetrationDistance; s.manifold_tolerance = mSettings.mManifoldTolerance;
		s.min_velocity_for_restitution = mSettings.mMinVelocityForRestitution; s.time_before_sleep = mSettings.mTimeBeforeSleep;
		s.point_velocity_sleep_threshold = mSettings.mPointVelocitySleepThreshold; s.num_velocity_steps = mSettings.mNumVelocitySteps; s.num_position_steps = mSettings.mNumPositionSteps;
		s.constraint_warm_start = mSettings.mConstraintWarmStart; s.use_body_pair_contact_cache = mSettings.mUseBodyPairContactCache; s.use_manifold_reduction = mSettings.mUseManifoldReduction;
		s.use_large_island_splitter = mSettings.mUseLargeIslandSplitter; s.allow_sleeping = mSettings.mAllowSleeping; s.check_active_edges = mSettings.mCheckActiveEdges;
	}

	int32_t ShapeID(const ShapeRef &inShape)
	{
		for (size_t i = 0; i < mShapes.size(); ++i) if (mShapes[i].get() == inShape.get()) return mShapeIDs[i];
		int32_t id = inShape->Upload(mWorld);
		mShapes.push_back(inShape); mShapeIDs.push_back(id);
		return id;
	}

	std::vector<Constraint *> mConstraints;   // by Constraint::mConstraintIndex
	uint32 mNumUploadedConstraints = 0;
	friend class Constraint;
	friend class PointConstraint;
	friend class DistanceConstraint;

	// query shapes are kept alive by the caller (raw pointers, as the reference takes them); a scale other than one uploads a ScaledShape
	struct QueryShape { const Shape *shape; Vec3 scale; int32_t id; };
	std::vector<QueryShape> mQueryShapes;
	int32_t QueryShapeID(const Shape *inShape, const Vec3 &inScale)
	{
		for (const QueryShape &q : mQueryShapes) if (q.shape == inShape && q.scale.x == inScale.x && q.scale.y == inScale.y && q.scale.z == inScale.z) return q.id;
		int32_t id = -1;
		for (size_t i = 0; i < mShapes.size(); ++i) if (mShapes[i].get() == inShape) id = mShapeIDs[i];
		if (id < 0) id = inShape->Upload(mWorld);
		if (id >= 0 && !(inScale.x == 1.0f && inScale.y == 1.0f && inScale.z == 1.0f))
		{
			float scale[3] = { inScale.x, inScale.y, inScale.z };
			id = b2j_shape_scaled(mWorld, id, scale);
		}
		mQueryShapes.push_back({ inShape, inScale, id });
		return id;
	}

	// ---- host mirror of the body state ------------------------------------------------------------------------------------
	// Flat arrays by body index, refreshed LAZILY on first use after a step and only as far as needed:
	//  * nothing but the step touched the device state since the arrays were filled -> only the rows of the bodies the step simulated
	//    are fetched and scattered (b2j_bodies_get_stepped_state; SURVEY 8f-1 incremental download),
	//  * otherwise (first use, API mutations, RestoreState, most bodies moving) a bulk copy of only the ARRAYS the caller reads: a
	//    loop that reads a million positions per step moves 12 bytes per body, not the 56 of the whole state.
	// The arrays are page locked (b2j_host_buffer_register) so the bulk copies are direct DMA.
	enum : uint32 { cStatePos = 1, cStateRot = 2, cStateLin = 4, cStateAng = 8, cStateActive = 16, cStateAll = 31 };

	void ResizeStateArrays(uint32 n)
	{
		if (mActiveIndex.size() == n) return;
		void *old[5] = { mPos.data(), mRot.data(), mLin.data(), mAng.data(), mActiveIndex.data() };
		if (mStateRegistered) for (void *p : old) if (p != nullptr) b2j_host_buffer_unregister(p);
		mPos.resize(3 * (size_t)n); mRot.resize(4 * (size_t)n); mLin.resize(3 * (size_t)n); mAng.resize(3 * (size_t)n); mActiveIndex.resize(n);
		mStateRegistered = n >= 4096; // (small worlds: the staging path is as fast and registration is not free)
		if (mStateRegistered)
		{
			b2j_host_buffer_register(mPos.data(), mPos.size() * 4); b2j_host_buffer_register(mRot.data(), mRot.size() * 4);
			b2j_host_buffer_register(mLin.data(), mLin.size() * 4); b2j_host_buffer_register(mAng.data(), mAng.size() * 4);
			b2j_host_buffer_register(mActiveIndex.data(), mActiveIndex.size() * 4);
		}
	}

	// bulk copy of the given arrays for every body slot
	void DownloadArrays(uint32 inMask)
	{
		uint32 n = (uint32)mBodies.size();
		if (n == 0 || inMask == 0) return;
		ResizeStateArrays(n);
		b2j_body_state st;
		memset(&st, 0, sizeof(st));
		if (inMask & cStatePos) st.position = mPos.data();
		if (inMask & cStateRot) st.rotation = mRot.data();
		if (inMask & cStateLin) st.linear_velocity = mLin.data();
		if (inMask & cStateAng) st.angular_velocity = mAng.data();
		if (inMask & cStateActive) st.active_index = mActiveIndex.data();
		b2j_bodies_get_state(mWorld, nullptr, n, &st);
		mLastDownloadCount = n;
	}

	// Eager full refresh (RestoreState): every array, new generation for the Body mirrors
	void DownloadState()
	{
		DownloadArrays(cStateAll);
		mStaleMask = 0; mBehindMask = 0;
		++mStateGeneration; // Body::Sync picks the new state up on first access
		for (uint8 &f : mSlotFlags) f &= 1; // the arrays are current for every body in the world
	}

	// The state getters call this first with the arrays they read.
	// mStaleMask: arrays that do not reflect the last step yet. mBehindMask: arrays that missed more than the last step (never filled,
	// not read for a while, or the device state was changed through the API): only a bulk copy makes them current. An array that is
	// exactly one step behind can be brought up to date from the rows of the bodies that step simulated.
	void EnsureState(uint32 inMask = cStateAll) const { if (mStaleMask & inMask) const_cast<PhysicsSystem *>(this)->RefreshState(inMask); }
	void RefreshState(uint32 inMask)
	{
		uint32 n = (uint32)mBodies.size();
		if (mActiveIndex.size() != n) mBehindMask = cStateAll;
		uint32 need = inMask & mStaleMask;
		uint32 bulk = need & mBehindMask;
		uint32 one_behind = mStaleMask & ~mBehindMask;   // every array the incremental rows can update, asked for or not
		if (need & ~mBehindMask)
		{
			uint32 count = b2j_bodies_get_stepped_state(mWorld, 0, nullptr, nullptr);
			if (count <= n / 2)
			{
				mLastDownloadCount = count;
				if (count > 0)
				{
					mSteppedIDs.resize(count); mSteppedPos.resize(3 * (size_t)count); mSteppedRot.resize(4 * (size_t)count); mSteppedLin.resize(3 * (size_t)count); mSteppedAng.resize(3 * (size_t)count); mSteppedActive.resize(count);
					b2j_body_state st;
					memset(&st, 0, sizeof(st));
					if (one_behind & cStatePos) st.position = mSteppedPos.data();
					if (one_behind & cStateRot) st.rotation = mSteppedRot.data();
					if (one_behind & cStateLin) st.linear_velocity = mSteppedLin.data();
					if (one_behind & cStateAng) st.angular_velocity = mSteppedAng.data();
					if (one_behind & cStateActive) st.active_index = mSteppedActive.data();
					b2j_bodies_get_stepped_state(mWorld, count, mSteppedIDs.data(), &st);
					for (uint32 k = 0; k < count; ++k)
					{
						size_t i = mSteppedIDs[k] & 0x7fffffu;
						if (i >= n) continue;
						if (one_behind & cStatePos) memcpy(&mPos[3 * i], &mSteppedPos[3 * (size_t)k], 12);
						if (one_behind & cStateRot) memcpy(&mRot[4 * i], &mSteppedRot[4 * (size_t)k], 16);
						if (one_behind & cStateLin) memcpy(&mLin[3 * i], &mSteppedLin[3 * (size_t)k], 12);
						if (one_behind & cStateAng) memcpy(&mAng[3 * i], &mSteppedAng[3 * (size_t)k], 12);
						if (one_behind & cStateActive) mActiveIndex[i] = mSteppedActive[k];
					}
				}
				mStaleMask &= ~one_behind;
			}
			else
				bulk = need; // most bodies moved: bulk copies (of the arrays that get read) are cheaper than a scatter
		}
		if (bulk != 0)
		{
			DownloadArrays(bulk);
			mStaleMask &= ~bulk;
			mBehindMask &= ~bulk;
		}
	}
	// after a step: everything is stale until read; what was still stale has now missed more than one step
	void MarkStateStale()
	{
		ResizeStateArrays((uint32)mBodies.size());
		mBehindMask |= mStaleMask;
		mStaleMask = cStateAll;
	}

	// current device state of a few bodies straight into their Body mirrors (bodies whose state changed through the interface since
	// the last Update, e.g. woken / pushed and then removed before the next step)
	void RefreshBodies(const std::vector<uint32> &inIDs)
	{
		uint32 n = (uint32)inIDs.size();
		std::vector<float> pos(3 * n), rot(4 * n), lin(3 * n), ang(3 * n);
		std::vector<uint32> active(n);
		b2j_body_state st;
		memset(&st, 0, sizeof(st));
		st.position = pos.data(); st.rotation = rot.data(); st.linear_velocity = lin.data(); st.angular_velocity = ang.data(); st.active_index = active.data();
		if (b2j_bodies_get_state(mWorld, inIDs.data(), n, &st) != 0) return;
		for (uint32 i = 0; i < n; ++i)
		{
			Body *b = mBodies[inIDs[i] & 0x7fffffu].get();
			b->mPosition = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			b->mRotation = Quat(rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3]);
			b->mLinearVelocity = Vec3(lin[3 * i], lin[3 * i + 1], lin[3 * i + 2]);
			b->mAngularVelocity = Vec3(ang[3 * i], ang[3 * i + 1], ang[3 * i + 2]);
			b->mActive = active[i] != B2J_INACTIVE_INDEX;
			b->mSyncGeneration = mStateGeneration;
		}
	}

	void ReplayEvents()
	{
		if (mActivationListener != nullptr)
		{
			uint32 n = b2j_activation_events_drain(mWorld, nullptr, 0);
			mActEvents.resize(n);
			if (n > 0) b2j_activation_events_drain(mWorld, mActEvents.data(), n);
			for (const b2j_activation_event &e : mActEvents)
			{
				BodyID id(e.body);
				const Body *b = mBodyInterface.TryGet(id);
				if (e.kind == B2J_EVENT_BODY_ACTIVATED) mActivationListener->OnBodyActivated(id, b? b->mUserData : 0);
				else mActivationListener->OnBodyDeactivated(id, b? b->mUserData : 0);
			}
		}
		if (mContactListener != nullptr)
		{
			uint32 n = b2j_events_drain(mWorld, nullptr, 0);
			mContactEvents.resize(n);
			if (n > 0) b2j_events_drain(mWorld, mContactEvents.data(), n);
			for (const b2j_contact_event &e : mContactEvents)
			{
				if (e.kind == B2J_EVENT_CONTACT_REMOVED)
				{
					SubShapeIDPair p;
					p.mBody1ID = BodyID(e.body1); p.mBody2ID = BodyID(e.body2); p.mSubShapeID1.mValue = e.sub_shape1; p.mSubShapeID2.mValue = e.sub_shape2;
					mContactListener->OnContactRemoved(p);
					continue;
				}
				const Body *b1 = mBodyInterface.TryGet(BodyID(e.body1)), *b2 = mBodyInterface.TryGet(BodyID(e.body2));
				if (b1 == nullptr || b2 == nullptr) continue;
				ContactManifold m;
				m.mBaseOffset = Vec3(e.base_offset[0], e.base_offset[1], e.base_offset[2]);
				m.mWorldSpaceNormal = Vec3(e.normal[0], e.normal[1], e.normal[2]);
				m.mPenetrationDepth = e.penetration_depth;
				m.mSubShapeID1.mValue = e.sub_shape1; m.mSubShapeID2.mValue = e.sub_shape2;
				for (uint32 i = 0; i < e.num_points; ++i)
				{
					m.mRelativeContactPointsOn1.push_back(Vec3(e.points1[i][0], e.points1[i][1], e.points1[i][2]));
					m.mRelativeContactPointsOn2.push_back(Vec3(e.points2[i][0], e.points2[i][1], e.points2[i][2]));
				}
				ContactSettings s;
				if (e.kind == B2J_EVENT_CONTACT_ADDED) mContactListener->OnContactAdded(*b1, *b2, m, s);
				else mContactListener->OnContactPersisted(*b1, *b2, m, s);
			}
		}
	}

	b2j_world *mWorld = nullptr;
	uint mMaxBodies = 0;
	Vec3 mGravity = Vec3(0.0f, -9.81f, 0.0f);
	PhysicsSettings mSettings;
	ContactListener *mContactListener = nullptr;
	BodyActivationListener *mActivationListener = nullptr;
	BodyInterface mBodyInterface;
	NarrowPhaseQuery mNarrowPhaseQuery;
	BroadPhaseQuery mBroadPhaseQuery;
	std::vector<std::unique_ptr<Body>> mBodies;   // by body index
	std::vector<uint32> mFreeIndices;
	std::vector<ShapeRef> mShapes;
	std::vector<int32_t> mShapeIDs;
	std::vector<b2j_body_desc> mPendingAdd;       // bodies added since the last flush
	std::vector<uint32> mPendingActivate;
	std::vector<uint32> mForceIDs;                // accumulated AddForce / AddTorque calls
	std::vector<float> mForces, mTorques;
	b2j_step_stats mStats = b2j_step_stats();
	std::vector<float> mPos, mRot, mLin, mAng;    // state of all body slots after the last Update (see Body::Sync)
	std::vector<uint32> mActiveIndex;
	std::vector<uint32> mSteppedIDs, mSteppedActive;   // scratch of the incremental download
	std::vector<float> mSteppedPos, mSteppedRot, mSteppedLin, mSteppedAng;
	uint32 mStaleMask = 0;                        // arrays (cState*) a step ran over since they were refreshed
	uint32 mBehindMask = 31;                      // arrays only a bulk copy can make current (see RefreshState)
	bool mStateRegistered = false;                // the arrays are page locked
	uint32 mLastDownloadCount = 0;
	uint32 mStateGeneration = 1;
	// per body index: full id (or invalid) and flags for the getters' fast path: bit 0 = in the world, bit 1 = the Body mirror is
	// newer than the downloaded arrays (added / changed through the interface since the last Update)
	std::vector<uint32> mSlotID;
	std::vector<uint8> mSlotFlags;
	bool FastSlot(const BodyID &id, size_t &outIndex, uint32 inArrays) const
	{
		EnsureState(inArrays);
		outIndex = id.GetIndex();
		return outIndex < mSlotID.size() && mSlotID[outIndex] == id.mID && mSlotFlags[outIndex] == 1 && outIndex < mActiveIndex.size();
	}
	void MarkMirrorNewer(const BodyID &id) { size_t i = id.GetIndex(); if (i < mSlotFlags.size()) mSlotFlags[i] |= 2; mBehindMask = cStateAll; }
	std::vector<b2j_contact_event> mContactEvents;
	std::vector<b2j_activation_event> mActEvents;
};

inline void Constraint::SetEnabled(bool inEnabled)
{
	mDesc.enabled = inEnabled;
	if (mSystem != nullptr && mConstraintIndex < mSystem->mNumUploadedConstraints)
	{
		uint8_t e = inEnabled;
		b2j_constraints_set_enabled(mSystem->mWorld, &mConstraintIndex, 1, &e);
	}
}

inline Vec3 PointConstraint::GetTotalLambdaPosition() const
{
	b2j_constraint_state st;
	memset(&st, 0, sizeof(st));
	if (mSystem != nullptr && mConstraintIndex < mSystem->mNumUploadedConstraints) b2j_constraints_get_state(mSystem->mWorld, mConstraintIndex, 1, &st);
	return Vec3(st.total_lambda[0], st.total_lambda[1], st.total_lambda[2]);
}

inline float DistanceConstraint::GetTotalLambdaPosition() const
{
	b2j_constraint_state st;
	memset(&st, 0, sizeof(st));
	if (mSystem != nullptr && mConstraintIndex < mSystem->mNumUploadedConstraints) b2j_constraints_get_state(mSystem->mWorld, mConstraintIndex, 1, &st);
	return st.total_lambda[0];
}

inline void NarrowPhaseQuery::CastRays(const RRayCast *inRays, int inNumber, RayCastResult *outHits, uint32 inObjectLayer) const
{
	if (inNumber <= 0) return;
	mSystem->mBodyInterface.Flush();
	static_assert(sizeof(RRayCast) == sizeof(b2j_ray), "RRayCast must be origin + direction as 6 floats");
	std::vector<b2j_ray_hit> hits((size_t)inNumber);
	if (b2j_query_cast_rays(mSystem->mWorld, reinterpret_cast<const b2j_ray *>(inRays), (uint32)inNumber, inObjectLayer, hits.data()) != 0) return;
	for (int i = 0; i < inNumber; ++i)
	{
		outHits[i].mBodyID = BodyID(hits[i].body);
		outHits[i].mFraction = hits[i].fraction;
		outHits[i].mSubShapeID2.mValue = hits[i].sub_shape;
	}
}

inline void NarrowPhaseQuery::CollideShapes(const Shape *inShape, const Vec3 &inShapeScale, const Quat *inRotations, const RVec3 *inPositions, int inNumber,
	const CollideShapeSettings &inSettings, std::vector<std::vector<CollideShapeResult>> &outHits, uint32 inObjectLayer) const
{
	outHits.assign((size_t)std::max(inNumber, 0), std::vector<CollideShapeResult>());
	if (inNumber <= 0) return;
	mSystem->mBodyInterface.Flush();
	int32_t shape = mSystem->QueryShapeID(inShape, inShapeScale);
	if (shape < 0) return;
	std::vector<b2j_shape_query> queries((size_t)inNumber);
	for (int i = 0; i < inNumber; ++i)
	{
		b2j_shape_query &q = queries[i];
		q.shape = shape;
		q.position[0] = inPositions[i].x; q.position[1] = inPositions[i].y; q.position[2] = inPositions[i].z;
		q.rotation[0] = inRotations[i].x; q.rotation[1] = inRotations[i].y; q.rotation[2] = inRotations[i].z; q.rotation[3] = inRotations[i].w;
		q.base_offset[0] = q.base_offset[1] = q.base_offset[2] = 0.0f;
	}
	uint32 cap = 16;
	std::vector<uint32> counts((size_t)inNumber);
	std::vector<b2j_collide_shape_hit> hits;
	for (;;)
	{
		hits.resize((size_t)inNumber * cap);
		if (b2j_query_collide_shape(mSystem->mWorld, queries.data(), (uint32)inNumber, inSettings.mMaxSeparationDistance, inObjectLayer, cap, counts.data(), hits.data()) != 0) return;
		uint32 most = 0;
		for (uint32 c : counts) most = std::max(most, c);
		if (most <= cap) break;
		cap = most;
	}
	for (int i = 0; i < inNumber; ++i)
		for (uint32 j = 0; j < counts[i]; ++j)
		{
			const b2j_collide_shape_hit &h = hits[(size_t)i * cap + j];
			CollideShapeResult r;
			r.mContactPointOn1 = Vec3(h.point1[0], h.point1[1], h.point1[2]); r.mContactPointOn2 = Vec3(h.point2[0], h.point2[1], h.point2[2]);
			r.mPenetrationAxis = Vec3(h.axis[0], h.axis[1], h.axis[2]); r.mPenetrationDepth = h.penetration_depth;
			r.mSubShapeID1.mValue = h.sub_shape1; r.mSubShapeID2.mValue = h.sub_shape2; r.mBodyID2 = BodyID(h.body);
			outHits[i].push_back(r);
		}
}

inline void NarrowPhaseQuery::CollideShape(const Shape *inShape, const Vec3 &inShapeScale, const Quat &inRotation, const RVec3 &inPosition, const CollideShapeSettings &inSettings,
	const RVec3 &inBaseOffset, std::vector<CollideShapeResult> &outHits, uint32 inObjectLayer) const
{
	// (one query: the base offset is applied on the host side of the call -- the device collides relative to the query position, which is
	// what the reference's callers pass as base offset, and the results are shifted to the requested one)
	outHits.clear();
	mSystem->mBodyInterface.Flush();
	int32_t shape = mSystem->QueryShapeID(inShape, inShapeScale);
	if (shape < 0) return;
	b2j_shape_query q;
	q.shape = shape;
	q.position[0] = inPosition.x; q.position[1] = inPosition.y; q.position[2] = inPosition.z;
	q.rotation[0] = inRotation.x; q.rotation[1] = inRotation.y; q.rotation[2] = inRotation.z; q.rotation[3] = inRotation.w;
	q.base_offset[0] = inBaseOffset.x; q.base_offset[1] = inBaseOffset.y; q.base_offset[2] = inBaseOffset.z;
	uint32 cap = 32, count = 0;
	std::vector<b2j_collide_shape_hit> hits;
	for (;;)
	{
		hits.resize(cap);
		if (b2j_query_collide_shape(mSystem->mWorld, &q, 1, inSettings.mMaxSeparationDistance, inObjectLayer, cap, &count, hits.data()) != 0) return;
		if (count <= cap) break;
		cap = count;
	}
	for (uint32 j = 0; j < count; ++j)
	{
		const b2j_collide_shape_hit &h = hits[j];
		CollideShapeResult r;
		r.mContactPointOn1 = Vec3(h.point1[0], h.point1[1], h.point1[2]); r.mContactPointOn2 = Vec3(h.point2[0], h.point2[1], h.point2[2]);
		r.mPenetrationAxis = Vec3(h.axis[0], h.axis[1], h.axis[2]); r.mPenetrationDepth = h.penetration_depth;
		r.mSubShapeID1.mValue = h.sub_shape1; r.mSubShapeID2.mValue = h.sub_shape2; r.mBodyID2 = BodyID(h.body);
		outHits.push_back(r);
	}
}

inline void NarrowPhaseQuery::CollideShape(const Shape *inShape, const Vec3 &inShapeScale, const RMat44 &inCenterOfMassTransform, const CollideShapeSettings &inSettings,
	const RVec3 &inBaseOffset, std::vector<CollideShapeResult> &outHits, uint32 inObjectLayer) const
{
	Quat q = inCenterOfMassTransform.GetQuaternion();
	CollideShape(inShape, inShapeScale, q, inCenterOfMassTransform.t, inSettings, inBaseOffset, outHits, inObjectLayer);
}

inline void BroadPhaseQuery::CollideVolume(int inMode, const float *inData, std::vector<BodyID> &outBodies, uint32 inObjectLayer) const
{
	mSystem->mBodyInterface.Flush();
	outBodies.clear();
	uint32 count = 0, cap = 64;
	std::vector<uint32> ids;
	for (;;)
	{
		ids.resize(cap);
		int r = inMode == 0? b2j_query_collide_aabox(mSystem->mWorld, inData, 1, inObjectLayer, cap, &count, ids.data())
			: (inMode == 1? b2j_query_collide_sphere(mSystem->mWorld, inData, 1, inObjectLayer, cap, &count, ids.data())
			: b2j_query_collide_point(mSystem->mWorld, inData, 1, inObjectLayer, cap, &count, ids.data()));
		if (r != 0) return;
		if (count <= cap) break;
		cap = count;
	}
	for (uint32 i = 0; i < count; ++i) outBodies.push_back(BodyID(ids[i]));
}

inline void BroadPhaseQuery::CollideAABox(const AABox &inBox, std::vector<BodyID> &outBodies, uint32 inObjectLayer) const
{
	float box[6] = { inBox.mMin.x, inBox.mMin.y, inBox.mMin.z, inBox.mMax.x, inBox.mMax.y, inBox.mMax.z };
	CollideVolume(0, box, outBodies, inObjectLayer);
}

inline void BroadPhaseQuery::CollideSphere(const Vec3 &inCenter, float inRadius, std::vector<BodyID> &outBodies, uint32 inObjectLayer) const
{
	float sphere[4] = { inCenter.x, inCenter.y, inCenter.z, inRadius };
	CollideVolume(1, sphere, outBodies, inObjectLayer);
}

inline void BroadPhaseQuery::CollidePoint(const Vec3 &inPoint, std::vector<BodyID> &outBodies, uint32 inObjectLayer) const
{
	float point[3] = { inPoint.x, inPoint.y, inPoint.z };
	CollideVolume(2, point, outBodies, inObjectLayer);
}

inline void Body::Sync() const
{
	if (mSystem == nullptr || !mInWorld || mSyncGeneration == mSystem->mStateGeneration)
		return;
	mSystem->EnsureState();
	mSyncGeneration = mSystem->mStateGeneration;
	size_t i = mID.GetIndex();
	if (i >= mSystem->mActiveIndex.size())
		return; // added after the last Update
	const std::vector<float> &p = mSystem->mPos, &r = mSystem->mRot, &l = mSystem->mLin, &a = mSystem->mAng;
	mPosition = Vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
	mRotation = Quat(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
	mLinearVelocity = Vec3(l[3 * i], l[3 * i + 1], l[3 * i + 2]);
	mAngularVelocity = Vec3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
	mActive = mSystem->mActiveIndex[i] != B2J_INACTIVE_INDEX;
}

inline RVec3 BodyInterface::GetCenterOfMassPosition(const BodyID &id) const
{
	size_t i;
	if (mSystem->FastSlot(id, i, PhysicsSystem::cStatePos)) return Vec3(mSystem->mPos[3 * i], mSystem->mPos[3 * i + 1], mSystem->mPos[3 * i + 2]);
	const Body *b = TryGet(id);
	return b? b->GetCenterOfMassPosition() : RVec3::sZero();
}

inline Quat BodyInterface::GetRotation(const BodyID &id) const
{
	size_t i;
	if (mSystem->FastSlot(id, i, PhysicsSystem::cStateRot)) return Quat(mSystem->mRot[4 * i], mSystem->mRot[4 * i + 1], mSystem->mRot[4 * i + 2], mSystem->mRot[4 * i + 3]);
	const Body *b = TryGet(id);
	return b? b->GetRotation() : Quat::sIdentity();
}

inline Vec3 BodyInterface::GetLinearVelocity(const BodyID &id) const
{
	size_t i;
	if (mSystem->FastSlot(id, i, PhysicsSystem::cStateLin)) return Vec3(mSystem->mLin[3 * i], mSystem->mLin[3 * i + 1], mSystem->mLin[3 * i + 2]);
	const Body *b = TryGet(id);
	return b? b->GetLinearVelocity() : Vec3::sZero();
}

inline Vec3 BodyInterface::GetAngularVelocity(const BodyID &id) const
{
	size_t i;
	if (mSystem->FastSlot(id, i, PhysicsSystem::cStateAng)) return Vec3(mSystem->mAng[3 * i], mSystem->mAng[3 * i + 1], mSystem->mAng[3 * i + 2]);
	const Body *b = TryGet(id);
	return b? b->GetAngularVelocity() : Vec3::sZero();
}

// ---- BodyInterface implementation -----------------------------------------------------------------------------------
inline b2j_world *BodyInterface::World() const { return mSystem->mWorld; }

inline const Body *BodyInterface::TryGet(const BodyID &id) const
{
	uint32 idx = id.GetIndex();
	if (id.IsInvalid() || idx >= mSystem->mBodies.size() || !mSystem->mBodies[idx] || mSystem->mBodies[idx]->mID != id || mSystem->mBodies[idx]->mDestroyed) return nullptr;
	return mSystem->mBodies[idx].get();
}

inline Body *BodyInterface::CreateBody(const BodyCreationSettings &s)
{
	PhysicsSystem &sys = *mSystem;
	uint32 index;
	uint8 sequence = 1; // BodyManager::AddBody: sequence numbers start at 1 (BodyManager.cpp GetNextSequenceNumber)
	if (!sys.mFreeIndices.empty()) { index = sys.mFreeIndices.back(); sys.mFreeIndices.pop_back(); sequence = uint8(sys.mBodies[index]? sys.mBodies[index]->mID.GetSequenceNumber() + 1 : 1); }
	else
	{
		if (sys.mBodies.size() >= sys.mMaxBodies) return nullptr; // out of bodies
		index = (uint32)sys.mBodies.size();
		sys.mBodies.emplace_back();
	}
	std::unique_ptr<Body> body(new Body);
	body->mID = BodyID(index, sequence);
	body->mSystem = &sys;
	body->mShape = s.GetShape();
	body->mMotionType = s.mMotionType;
	body->mObjectLayer = s.mObjectLayer;
	body->mUserData = s.mUserData;
	body->mRotation = s.mRotation;
	// Body::SetPositionAndRotationInternal: mPosition = inPosition + inRotation * shape centre of mass
	body->mPosition = s.mPosition + s.mRotation * body->mShape->GetCenterOfMass();
	body->mLinearVelocity = s.mLinearVelocity;
	body->mAngularVelocity = s.mAngularVelocity;

	b2j_body_desc &d = body->mDesc;
	memset(&d, 0, sizeof(d));
	d.id = body->mID.mID;
	d.shape = sys.ShapeID(body->mShape);
	d.motion_type = (uint8_t)s.mMotionType;
	d.allowed_dofs = (uint8_t)s.mAllowedDOFs;
	d.num_velocity_steps_override = (uint8_t)s.mNumVelocityStepsOverride;
	d.num_position_steps_override = (uint8_t)s.mNumPositionStepsOverride;
	d.object_layer = s.mObjectLayer;
	d.flags = (uint16_t)((s.mIsSensor? B2J_BODY_SENSOR : 0) | (s.mAllowSleeping? B2J_BODY_ALLOW_SLEEPING : 0) | (s.mUseManifoldReduction? B2J_BODY_USE_MANIFOLD_REDUCTION : 0)
		| (s.mApplyGyroscopicForce? B2J_BODY_GYROSCOPIC : 0) | (s.mCollideKinematicVsNonDynamic? B2J_BODY_KIN_VS_NONDYN : 0));
	d.position[0] = body->mPosition.x; d.position[1] = body->mPosition.y; d.position[2] = body->mPosition.z;
	d.rotation[0] = s.mRotation.x; d.rotation[1] = s.mRotation.y; d.rotation[2] = s.mRotation.z; d.rotation[3] = s.mRotation.w;
	d.linear_velocity[0] = s.mLinearVelocity.x; d.linear_velocity[1] = s.mLinearVelocity.y; d.linear_velocity[2] = s.mLinearVelocity.z;
	d.angular_velocity[0] = s.mAngularVelocity.x; d.angular_velocity[1] = s.mAngularVelocity.y; d.angular_velocity[2] = s.mAngularVelocity.z;
	d.inertia_rotation[3] = 1.0f;
	d.linear_damping = s.mLinearDamping; d.angular_damping = s.mAngularDamping;
	d.max_linear_velocity = s.mMaxLinearVelocity; d.max_angular_velocity = s.mMaxAngularVelocity;
	d.gravity_factor = s.mGravityFactor; d.friction = s.mFriction; d.restitution = s.mRestitution;
	body->mHasMotionProperties = s.HasMassProperties();
	if (s.HasMassProperties())
	{
		// MotionProperties::SetMassProperties (MotionProperties.cpp:12-61)
		MassProperties mp = s.GetMassProperties();
		uint dofs = (uint)s.mAllowedDOFs;
		d.inv_mass = (dofs & 7) == 0? 0.0f : 1.0f / mp.mMass;
		if (((dofs >> 3) & 7) != 0)
		{
			Quat rot; Vec3 diag;
			if (mp.DecomposePrincipalMomentsOfInertia(rot, diag) && !(diag.LengthSq() <= 1.0e-12f))
			{
				d.inv_inertia_diag[0] = 1.0f / diag.x; d.inv_inertia_diag[1] = 1.0f / diag.y; d.inv_inertia_diag[2] = 1.0f / diag.z;
				d.inertia_rotation[0] = rot.x; d.inertia_rotation[1] = rot.y; d.inertia_rotation[2] = rot.z; d.inertia_rotation[3] = rot.w;
			}
			else
				d.inv_inertia_diag[0] = d.inv_inertia_diag[1] = d.inv_inertia_diag[2] = 2.5f * d.inv_mass;
		}
		body->mDynamicInvMass = d.inv_mass;
		if (s.mMotionType != EMotionType::Dynamic) d.inv_mass = 0.0f;
	}
	d.has_bounds = 0; // bounds and sleep test spheres are computed on the device from shape + pose
	Body *result = body.get();
	sys.mBodies[index] = std::move(body);
	if (sys.mSlotID.size() <= index) { sys.mSlotID.resize(index + 1, BodyID::cInvalidBodyID); sys.mSlotFlags.resize(index + 1, 0); }
	sys.mSlotID[index] = result->mID.mID;
	sys.mSlotFlags[index] = 0;
	return result;
}

inline void BodyInterface::AddBodies(const BodyID *inBodies, int inNumber, EActivation inActivationMode)
{
	PhysicsSystem &sys = *mSystem;
	for (int i = 0; i < inNumber; ++i)
	{
		Body *b = const_cast<Body *>(TryGet(inBodies[i]));
		if (b == nullptr || b->mInWorld) continue;
		b->mInWorld = true;
		b->mSyncGeneration = sys.mStateGeneration; // the creation state is current until the next Update
		sys.mSlotFlags[b->mID.GetIndex()] = 1 | 2;
		// the body enters the world with the state it HAS (creation state, or the state it left the world with / was given since:
		// BodyManager keeps the Body object between RemoveBody and AddBody), not with its creation time descriptor
		b2j_body_desc &d = b->mDesc;
		d.position[0] = b->mPosition.x; d.position[1] = b->mPosition.y; d.position[2] = b->mPosition.z;
		d.rotation[0] = b->mRotation.x; d.rotation[1] = b->mRotation.y; d.rotation[2] = b->mRotation.z; d.rotation[3] = b->mRotation.w;
		d.linear_velocity[0] = b->mLinearVelocity.x; d.linear_velocity[1] = b->mLinearVelocity.y; d.linear_velocity[2] = b->mLinearVelocity.z;
		d.angular_velocity[0] = b->mAngularVelocity.x; d.angular_velocity[1] = b->mAngularVelocity.y; d.angular_velocity[2] = b->mAngularVelocity.z;
		sys.mPendingAdd.push_back(d);
		sys.mBehindMask = PhysicsSystem::cStateAll; // (a body added asleep is not among the bodies the next step simulates)
		if (inActivationMode == EActivation::Activate && b->mMotionType != EMotionType::Static)
		{
			sys.mPendingActivate.push_back(b->mID.mID);
			b->mActive = true;
		}
	}
}

inline void BodyInterface::AddBody(const BodyID &inBodyID, EActivation inActivationMode) { AddBodies(&inBodyID, 1, inActivationMode); }

inline void BodyInterface::Flush()
{
	PhysicsSystem &sys = *mSystem;
	if (!sys.mPendingAdd.empty())
	{
		b2j_bodies_add(sys.mWorld, sys.mPendingAdd.data(), (uint32)sys.mPendingAdd.size());
		sys.mPendingAdd.clear();
	}
	if (!sys.mPendingActivate.empty())
	{
		b2j_bodies_activate_or_reset_sleep_timer(sys.mWorld, sys.mPendingActivate.data(), (uint32)sys.mPendingActivate.size()); // BodyInterface::ActivateBodyInternal
		sys.mPendingActivate.clear();
	}
	if (!sys.mForceIDs.empty())
	{
		// several calls for one body (AddForce at a point = a force and a torque entry): summed here in call order, like the reference adds
		// them to Body::mForce / mTorque one call after the other; the device call wants every id once
		size_t n = sys.mForceIDs.size(), m = 0;
		std::vector<uint32> &ids = sys.mForceIDs;
		for (size_t i = 0; i < n; ++i)
		{
			size_t j = 0;
			while (j < m && ids[j] != ids[i]) ++j; // (a handful of entries per flush)
			if (j == m) { ids[m] = ids[i]; for (int k = 0; k < 3; ++k) { sys.mForces[3 * m + k] = sys.mForces[3 * i + k]; sys.mTorques[3 * m + k] = sys.mTorques[3 * i + k]; } ++m; }
			else for (int k = 0; k < 3; ++k) { sys.mForces[3 * j + k] += sys.mForces[3 * i + k]; sys.mTorques[3 * j + k] += sys.mTorques[3 * i + k]; }
		}
		ids.resize(m);
		b2j_bodies_add_force_torque(sys.mWorld, sys.mForceIDs.data(), (uint32)sys.mForceIDs.size(), sys.mForces.data(), sys.mTorques.data());
		sys.mForceIDs.clear(); sys.mForces.clear(); sys.mTorques.clear();
	}
}

inline void BodyInterface::RemoveBody(const BodyID &inBodyID) { BodyID id = inBodyID; RemoveBodies(&id, 1); }

// BodyInterface::RemoveBodies (BodyInterface.cpp:259-281): one b2j_bodies_remove call for the whole batch
inline void BodyInterface::RemoveBodies(BodyID *ioBodies, int inNumber)
{
	PhysicsSystem &sys = *mSystem;
	std::vector<uint32> ids;
	std::vector<Body *> bodies;
	for (int i = 0; i < inNumber; ++i)
	{
		Body *b = const_cast<Body *>(TryGet(ioBodies[i]));
		if (b == nullptr || !b->mInWorld) continue;
		bodies.push_back(b);
		ids.push_back(b->mID.mID);
	}
	if (ids.empty()) return;
	Flush();
	// bodies changed through the interface since the last Update have a mirror that is newer than the downloaded arrays
	std::vector<uint32> stale;
	for (Body *b : bodies) if (b->mSyncGeneration != sys.mStateGeneration || (sys.mSlotFlags[b->mID.GetIndex()] & 2)) stale.push_back(b->mID.mID);
	if (!stale.empty()) sys.RefreshBodies(stale);
	b2j_bodies_remove(World(), ids.data(), (uint32)ids.size());
	sys.mBehindMask = PhysicsSystem::cStateAll;
	for (Body *b : bodies)
	{
		b->Sync(); // the Body keeps the pose it left the world with
		if (b->mActive) { b->mLinearVelocity = Vec3::sZero(); b->mAngularVelocity = Vec3::sZero(); } // BodyManager::DeactivateBodies (BodyManager.cpp:529-568)
		b->mInWorld = false; b->mActive = false;
		sys.mSlotFlags[b->mID.GetIndex()] = 0;
	}
}

inline void BodyInterface::DestroyBody(const BodyID &inBodyID)
{
	Body *b = const_cast<Body *>(TryGet(inBodyID));
	if (b == nullptr) return;
	if (b->mInWorld) RemoveBody(inBodyID);
	mSystem->mFreeIndices.push_back(inBodyID.GetIndex());
	mSystem->mSlotID[inBodyID.GetIndex()] = BodyID::cInvalidBodyID;
	b->mShape.reset();
	b->mDestroyed = true; // the slot keeps the Body for its sequence number; it can no longer be looked up
}

inline void BodyInterface::SetPositionAndRotation(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, EActivation inActivationMode)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr) return;
	b->Sync();
	mSystem->MarkMirrorNewer(id);
	b->mRotation = inRotation;
	b->mPosition = inPosition + inRotation * b->mShape->GetCenterOfMass();
	if (!b->mInWorld) { b->mDesc.position[0] = b->mPosition.x; b->mDesc.position[1] = b->mPosition.y; b->mDesc.position[2] = b->mPosition.z; b->mDesc.rotation[0] = inRotation.x; b->mDesc.rotation[1] = inRotation.y; b->mDesc.rotation[2] = inRotation.z; b->mDesc.rotation[3] = inRotation.w; return; }
	Flush();
	uint32 bid = id.mID;
	float pos[3] = { b->mPosition.x, b->mPosition.y, b->mPosition.z }, rot[4] = { inRotation.x, inRotation.y, inRotation.z, inRotation.w };
	b2j_body_state st;
	memset(&st, 0, sizeof(st));
	st.position = pos; st.rotation = rot;
	b2j_bodies_set_state(World(), &bid, 1, &st);
	if (inActivationMode == EActivation::Activate && b->mMotionType != EMotionType::Static) { b2j_bodies_activate_or_reset_sleep_timer(World(), &bid, 1); b->mActive = true; }
}

// MotionProperties::LockTranslation / LockAngular (inMask3 = the three DOF bits of the vector), ClampLinear / AngularVelocity
inline Vec3 sLockDOFs(const Vec3 &v, uint inMask3) { return Vec3((inMask3 & 1)? v.x : 0.0f, (inMask3 & 2)? v.y : 0.0f, (inMask3 & 4)? v.z : 0.0f); }
inline Vec3 sClampVelocity(const Vec3 &v, float inMax)
{
	float len_sq = v.LengthSq();
	return len_sq > inMax * inMax? v * (inMax / std::sqrt(len_sq)) : v;
}

inline void BodyInterface::SetLinearAndAngularVelocity(const BodyID &id, const Vec3 &inLV, const Vec3 &inAV)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr || b->mMotionType == EMotionType::Static) return;
	b->Sync();
	mSystem->MarkMirrorNewer(id);
	// Body::SetLinearVelocityClamped / SetAngularVelocityClamped: locked DOFs zeroed, clamped to the maximum velocities
	Vec3 lv = sClampVelocity(sLockDOFs(inLV, b->mDesc.allowed_dofs), b->mDesc.max_linear_velocity), av = sClampVelocity(sLockDOFs(inAV, b->mDesc.allowed_dofs >> 3), b->mDesc.max_angular_velocity);
	b->mLinearVelocity = lv; b->mAngularVelocity = av;
	if (!b->mInWorld) { memcpy(b->mDesc.linear_velocity, &lv, 12); memcpy(b->mDesc.angular_velocity, &av, 12); return; }
	Flush();
	uint32 bid = id.mID;
	float l[3] = { lv.x, lv.y, lv.z }, a[3] = { av.x, av.y, av.z };
	b2j_body_state st;
	memset(&st, 0, sizeof(st));
	st.linear_velocity = l; st.angular_velocity = a;
	b2j_bodies_set_state(World(), &bid, 1, &st);
	// BodyInterface::SetLinearAndAngularVelocity activates the body when the velocity is non zero
	if (!b->mActive && (inLV.LengthSq() > 1.0e-12f || inAV.LengthSq() > 1.0e-12f)) { b2j_bodies_activate(World(), &bid, 1); b->mActive = true; } // (!IsNearZero)
}

inline void BodyInterface::AddForce(const BodyID &id, const Vec3 &f, EActivation inActivationMode)
{
	PhysicsSystem &sys = *mSystem;
	// BodyInterface::AddForce (BodyInterface.cpp): dynamic bodies only; applied when the body is active or gets activated
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic || !(inActivationMode == EActivation::Activate || b->IsActive())) return;
	if (inActivationMode == EActivation::Activate && b->mInWorld) { sys.mPendingActivate.push_back(id.mID); b->mActive = true; } // (the device follows at the next flush)
	sys.mForceIDs.push_back(id.mID);
	sys.mForces.push_back(f.x); sys.mForces.push_back(f.y); sys.mForces.push_back(f.z);
	sys.mTorques.push_back(0); sys.mTorques.push_back(0); sys.mTorques.push_back(0);
}

inline void BodyInterface::AddTorque(const BodyID &id, const Vec3 &t, EActivation inActivationMode)
{
	PhysicsSystem &sys = *mSystem;
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic || !(inActivationMode == EActivation::Activate || b->IsActive())) return;
	if (inActivationMode == EActivation::Activate && b->mInWorld) { sys.mPendingActivate.push_back(id.mID); b->mActive = true; } // (the device follows at the next flush)
	sys.mForceIDs.push_back(id.mID);
	sys.mForces.push_back(0); sys.mForces.push_back(0); sys.mForces.push_back(0);
	sys.mTorques.push_back(t.x); sys.mTorques.push_back(t.y); sys.mTorques.push_back(t.z);
}

inline void BodyInterface::SetActive(const BodyID &id, bool inActive)
{
	const Body *b = TryGet(id);
	if (b == nullptr || !b->mInWorld || b->mMotionType == EMotionType::Static) return;
	uint32 bid = id.mID;
	Flush();
	// (ActivateBody = BodyInterface::ActivateBodyInternal: a body that is active already gets its sleep timer reset)
	if (inActive) b2j_bodies_activate_or_reset_sleep_timer(World(), &bid, 1); else b2j_bodies_deactivate(World(), &bid, 1);
	mSystem->mBehindMask = PhysicsSystem::cStateAll;
	b->Sync();
	mSystem->MarkMirrorNewer(id);
	b->mActive = inActive;
	if (!inActive) { b->mLinearVelocity = Vec3::sZero(); b->mAngularVelocity = Vec3::sZero(); } // BodyManager::DeactivateBodies resets the velocities
}

inline void BodyInterface::SetMotionType(const BodyID &id, EMotionType inMotionType, EActivation inActivationMode)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr || b->mMotionType == inMotionType) { if (b != nullptr && b->mInWorld && inMotionType != EMotionType::Static && inActivationMode == EActivation::Activate) ActivateBody(id); return; }
	if (inMotionType != EMotionType::Static && !b->mHasMotionProperties) return; // Body::SetMotionType asserts: created without mAllowDynamicOrKinematic
	b->Sync();
	b->mMotionType = inMotionType;
	b->mDesc.motion_type = (uint8_t)inMotionType;
	b->mDesc.inv_mass = inMotionType == EMotionType::Dynamic? b->mDynamicInvMass : 0.0f;
	if (inMotionType == EMotionType::Static) { b->mLinearVelocity = Vec3::sZero(); b->mAngularVelocity = Vec3::sZero(); b->mActive = false; }
	if (inMotionType != EMotionType::Dynamic) { memset(b->mDesc.force, 0, sizeof(b->mDesc.force)); memset(b->mDesc.torque, 0, sizeof(b->mDesc.torque)); }
	if (!b->mInWorld) return;
	Flush();
	uint32 bid = id.mID; uint8_t mt = (uint8_t)inMotionType;
	b2j_body_info_update u;
	memset(&u, 0, sizeof(u));
	u.motion_type = &mt; u.inv_mass = &b->mDynamicInvMass;
	b2j_bodies_set_info(World(), &bid, 1, &u);
	mSystem->MarkMirrorNewer(id);
	if (inMotionType != EMotionType::Static && inActivationMode == EActivation::Activate) ActivateBody(id);
}

inline void BodyInterface::SetObjectLayer(const BodyID &id, ObjectLayer inLayer)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr || b->mObjectLayer == inLayer) return;
	b->mObjectLayer = inLayer;
	b->mDesc.object_layer = inLayer;
	if (!b->mInWorld) return;
	Flush();
	uint32 bid = id.mID; uint16_t layer = inLayer;
	b2j_body_info_update u;
	memset(&u, 0, sizeof(u));
	u.object_layer = &layer;
	b2j_bodies_set_info(World(), &bid, 1, &u);
}

inline void BodyInterface::SetShape(const BodyID &id, const ShapeRef &inShape, bool inUpdateMassProperties, EActivation inActivationMode)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr || b->mShape.get() == inShape.get()) return;
	b->Sync();
	// Body::SetShapeInternal: the centre of mass position follows the new shape
	Vec3 old_com = b->mShape->GetCenterOfMass();
	b->mShape = inShape;
	b->mPosition = b->mPosition + b->mRotation * (inShape->GetCenterOfMass() - old_com);
	b2j_body_desc &d = b->mDesc;
	d.shape = mSystem->ShapeID(inShape);
	d.position[0] = b->mPosition.x; d.position[1] = b->mPosition.y; d.position[2] = b->mPosition.z;
	if (inUpdateMassProperties && b->mHasMotionProperties)
	{
		// Body::UpdateCenterOfMassInternal -> MotionProperties::SetMassProperties(allowed DOFs, shape mass properties)
		MassProperties mp = inShape->GetMassProperties();
		uint dofs = d.allowed_dofs;
		b->mDynamicInvMass = (dofs & 7) == 0? 0.0f : 1.0f / mp.mMass;
		d.inv_inertia_diag[0] = d.inv_inertia_diag[1] = d.inv_inertia_diag[2] = 0.0f;
		d.inertia_rotation[0] = d.inertia_rotation[1] = d.inertia_rotation[2] = 0.0f; d.inertia_rotation[3] = 1.0f;
		if (((dofs >> 3) & 7) != 0)
		{
			Quat rot; Vec3 diag;
			if (mp.DecomposePrincipalMomentsOfInertia(rot, diag) && !(diag.LengthSq() <= 1.0e-12f))
			{
				d.inv_inertia_diag[0] = 1.0f / diag.x; d.inv_inertia_diag[1] = 1.0f / diag.y; d.inv_inertia_diag[2] = 1.0f / diag.z;
				d.inertia_rotation[0] = rot.x; d.inertia_rotation[1] = rot.y; d.inertia_rotation[2] = rot.z; d.inertia_rotation[3] = rot.w;
			}
			else
				d.inv_inertia_diag[0] = d.inv_inertia_diag[1] = d.inv_inertia_diag[2] = 2.5f * b->mDynamicInvMass;
		}
		d.inv_mass = b->mMotionType == EMotionType::Dynamic? b->mDynamicInvMass : 0.0f;
	}
	if (!b->mInWorld) return;
	Flush();
	uint32 bid = id.mID; int32_t shape = d.shape;
	b2j_body_info_update u;
	memset(&u, 0, sizeof(u));
	u.shape = &shape;
	if (inUpdateMassProperties && b->mHasMotionProperties) { u.inv_mass = &b->mDynamicInvMass; u.inv_inertia_diag = d.inv_inertia_diag; u.inertia_rotation = d.inertia_rotation; }
	b2j_bodies_set_info(World(), &bid, 1, &u);
	mSystem->MarkMirrorNewer(id);
	if (inActivationMode == EActivation::Activate && b->mMotionType != EMotionType::Static) ActivateBody(id);
}

inline void BodyInterface::InvalidateContactCache(const BodyID &id)
{
	const Body *b = TryGet(id);
	if (b == nullptr || !b->mInWorld) return;
	Flush();
	uint32 bid = id.mID;
	b2j_body_info_update u;
	memset(&u, 0, sizeof(u));
	u.invalidate_contact_cache = 1;
	b2j_bodies_set_info(World(), &bid, 1, &u);
}

inline void BodyInterface::SetParam(const BodyID &id, float b2j_body_desc::*inDescMember, const float *b2j_body_params::*inParamMember, float inValue)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr) return;
	b->mDesc.*inDescMember = inValue;
	if (!b->mInWorld) return; // uploaded with the body
	Flush();
	b2j_body_params p;
	memset(&p, 0, sizeof(p));
	p.*inParamMember = &inValue;
	uint32 bid = id.mID;
	b2j_bodies_set_params(World(), &bid, 1, &p);
}

inline void BodyInterface::AddImpulse(const BodyID &id, const Vec3 &inImpulse)
{
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic) return;
	// Body::AddImpulse: v += invM * impulse
	SetLinearVelocity(id, b->GetLinearVelocity() + b->mDesc.inv_mass * inImpulse);
	if (b->mInWorld) ActivateNoReset(id); // (BodyInterface::AddImpulse: if (!body.IsActive()) ActivateBodies)
}

inline void BodyInterface::AddAngularImpulse(const BodyID &id, const Vec3 &inAngularImpulse)
{
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic) return;
	// MotionProperties::MultiplyWorldSpaceInverseInertiaByVector (MotionProperties.inl:77-92): R D R^T v with R = body rotation * inertia rotation
	const b2j_body_desc &d = b->mDesc;
	Quat q = b->GetRotation() * Quat(d.inertia_rotation[0], d.inertia_rotation[1], d.inertia_rotation[2], d.inertia_rotation[3]);
	// Mat44::sRotation(q) columns (Mat44.inl), then rotation.Multiply3x3(invI * rotation.Multiply3x3Transposed(v))
	float tx = q.x + q.x, ty = q.y + q.y, tz = q.z + q.z;
	float xx = tx * q.x, yy = ty * q.y, zz = tz * q.z, xy = tx * q.y, xz = tx * q.z, xw = tx * q.w, yz = ty * q.z, yw = ty * q.w, zw = tz * q.w;
	Vec3 c0((1.0f - yy) - zz, xy + zw, xz - yw), c1(xy - zw, (1.0f - zz) - xx, yz + xw), c2(xz + yw, yz - xw, (1.0f - xx) - yy);
	const Vec3 &v = inAngularImpulse;
	Vec3 local(c0.x * v.x + c0.y * v.y + c0.z * v.z, c1.x * v.x + c1.y * v.y + c1.z * v.z, c2.x * v.x + c2.y * v.y + c2.z * v.z);
	local = Vec3(d.inv_inertia_diag[0] * local.x, d.inv_inertia_diag[1] * local.y, d.inv_inertia_diag[2] * local.z);
	Vec3 delta(c0.x * local.x + c1.x * local.y + c2.x * local.z, c0.y * local.x + c1.y * local.y + c2.y * local.z, c0.z * local.x + c1.z * local.y + c2.z * local.z);
	SetAngularVelocity(id, b->GetAngularVelocity() + delta);
	if (b->mInWorld) ActivateNoReset(id); // (BodyInterface::AddImpulse: if (!body.IsActive()) ActivateBodies)
}

inline void BodyInterface::AddImpulse(const BodyID &id, const Vec3 &inImpulse, const RVec3 &inPoint)
{
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic) return;
	// Body::AddImpulse(impulse, point): linear part + angular part (point - centre of mass) x impulse
	Vec3 r = inPoint - b->GetCenterOfMassPosition();
	AddImpulse(id, inImpulse);
	AddAngularImpulse(id, r.Cross(inImpulse));
}

inline void BodyInterface::SetPositionAndRotationWhenChanged(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, EActivation inActivationMode)
{
	// BodyInterface.cpp: only when the body is not already (close to) there: !IsClose(position) || !IsClose(rotation), default tolerances 1e-12
	const Body *b = TryGet(id);
	if (b == nullptr) return;
	Vec3 dp = b->GetPosition() - inPosition;
	Quat r = b->GetRotation();
	float dr = ((r.x - inRotation.x) * (r.x - inRotation.x) + (r.y - inRotation.y) * (r.y - inRotation.y)) + ((r.z - inRotation.z) * (r.z - inRotation.z) + (r.w - inRotation.w) * (r.w - inRotation.w));
	if (dp.LengthSq() > 1.0e-12f || dr > 1.0e-12f)
		SetPositionAndRotation(id, inPosition, inRotation, inActivationMode);
}

inline void BodyInterface::SetPositionRotationAndVelocity(const BodyID &id, const RVec3 &inPosition, const Quat &inRotation, const Vec3 &inLinearVelocity, const Vec3 &inAngularVelocity)
{
	const Body *b = TryGet(id);
	if (b == nullptr) return;
	SetPositionAndRotation(id, inPosition, inRotation, EActivation::DontActivate);
	if (b->mMotionType == EMotionType::Static) return;
	bool was_active = b->IsActive();
	SetLinearAndAngularVelocity(id, inLinearVelocity, inAngularVelocity); // (activates when a velocity is not near zero, like the reference)
	(void)was_active;
}

inline Vec3 BodyInterface::GetPointVelocity(const BodyID &id, const RVec3 &inPoint) const
{
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType == EMotionType::Static) return Vec3::sZero();
	// MotionProperties::GetPointVelocityCOM: mLinearVelocity + mAngularVelocity.Cross(inPointRelativeToCOM)
	return b->GetLinearVelocity() + b->GetAngularVelocity().Cross(inPoint - b->GetCenterOfMassPosition());
}

// Jolt's ACos (Vec4::ASin / ACos, Vec4.inl:1266-1305: the cephes asinf polynomial), scalar
inline float sJoltACos(float inX)
{
	float sign = inX < 0.0f? -1.0f : 1.0f;
	float a = std::min(inX < 0.0f? -inX : inX, 1.0f);
	bool greater = a > 0.5f;
	float z = greater? 0.5f * (1.0f - a) : a * a;
	float x = greater? std::sqrt(z) : a;
	z = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
	if (greater) z = 0.5f * JPH_PI - (z + z);
	float asin = sign < 0.0f? -z : z;
	return 0.5f * JPH_PI - asin;
}

inline void BodyInterface::MoveKinematic(const BodyID &id, const RVec3 &inTargetPosition, const Quat &inTargetRotation, float inDeltaTime)
{
	Body *b = const_cast<Body *>(TryGet(id));
	if (b == nullptr || b->mMotionType == EMotionType::Static) return;
	// Body::MoveKinematic (Body.cpp:81-95) + MotionProperties::MoveKinematic (MotionProperties.inl:9-21): the velocities that take the body
	// to the target in inDeltaTime, not clamped
	Vec3 new_com = inTargetPosition + inTargetRotation * b->mShape->GetCenterOfMass();
	Vec3 delta_pos = new_com - b->GetCenterOfMassPosition();
	Quat r = b->GetRotation();
	Quat delta_rotation = inTargetRotation * Quat(-r.x, -r.y, -r.z, r.w);
	Vec3 lv = sLockDOFs(Vec3(delta_pos.x / inDeltaTime, delta_pos.y / inDeltaTime, delta_pos.z / inDeltaTime), b->mDesc.allowed_dofs);
	// Quat::GetAngularVelocity (Quat.inl:186-204)
	bool flip = delta_rotation.w < 0.0f;
	Vec3 xyz(flip? -delta_rotation.x : delta_rotation.x, flip? -delta_rotation.y : delta_rotation.y, flip? -delta_rotation.z : delta_rotation.z);
	float w = flip? -delta_rotation.w : delta_rotation.w;
	float xyz_len_sq = xyz.LengthSq();
	Vec3 av;
	if (xyz_len_sq < 4.0e-4f)
		av = (2.0f / inDeltaTime) * xyz;
	else
	{
		float angle = 2.0f * sJoltACos(w);
		float d = std::sqrt(xyz_len_sq) * inDeltaTime;
		av = Vec3(xyz.x / d, xyz.y / d, xyz.z / d) * angle;
	}
	av = sLockDOFs(av, b->mDesc.allowed_dofs >> 3);
	b->Sync();
	mSystem->MarkMirrorNewer(id);
	b->mLinearVelocity = lv; b->mAngularVelocity = av;
	if (!b->mInWorld) { memcpy(b->mDesc.linear_velocity, &lv, 12); memcpy(b->mDesc.angular_velocity, &av, 12); return; }
	Flush();
	uint32 bid = id.mID;
	float l[3] = { lv.x, lv.y, lv.z }, a[3] = { av.x, av.y, av.z };
	b2j_body_state st;
	memset(&st, 0, sizeof(st));
	st.linear_velocity = l; st.angular_velocity = a;
	b2j_bodies_set_state(World(), &bid, 1, &st);
	auto near_zero = [](const Vec3 &v) { return v.LengthSq() <= 1.0e-12f; };
	if (!b->mActive && (!near_zero(lv) || !near_zero(av))) { b2j_bodies_activate(World(), &bid, 1); b->mActive = true; }
}

inline Mat44RT BodyInterface::GetInverseInertia(const BodyID &id) const
{
	Mat44RT out;
	out.c0 = out.c1 = out.c2 = Vec3::sZero();
	const Body *b = TryGet(id);
	if (b == nullptr || b->mMotionType != EMotionType::Dynamic) return out;
	// MotionProperties::GetInverseInertiaForRotation (MotionProperties.inl:66-81)
	const b2j_body_desc &d = b->mDesc;
	Mat44RT rot = Mat44RT::sRotation(b->GetRotation()), irot = Mat44RT::sRotation(Quat(d.inertia_rotation[0], d.inertia_rotation[1], d.inertia_rotation[2], d.inertia_rotation[3]));
	Vec3 r0 = (rot.c0 * irot.c0.x + rot.c1 * irot.c0.y) + rot.c2 * irot.c0.z, r1 = (rot.c0 * irot.c1.x + rot.c1 * irot.c1.y) + rot.c2 * irot.c1.z, r2 = (rot.c0 * irot.c2.x + rot.c1 * irot.c2.y) + rot.c2 * irot.c2.z; // Multiply3x3
	Vec3 s0 = d.inv_inertia_diag[0] * r0, s1 = d.inv_inertia_diag[1] * r1, s2 = d.inv_inertia_diag[2] * r2;
	// rotation.Multiply3x3RightTransposed(scaled): column j = (r0 * s0[j] + r1 * s1[j]) + r2 * s2[j]
	Vec3 cols[3];
	for (int j = 0; j < 3; ++j) cols[j] = (r0 * s0[j] + r1 * s1[j]) + r2 * s2[j];
	uint mask = uint(d.allowed_dofs) >> 3;
	for (int j = 0; j < 3; ++j) cols[j] = (mask & (1u << j))? sLockDOFs(cols[j], mask) : Vec3::sZero();
	out.c0 = cols[0]; out.c1 = cols[1]; out.c2 = cols[2];
	return out;
}

inline void BodyInterface::AddForce(const BodyID &id, const Vec3 &inForce, const RVec3 &inPoint, EActivation inActivationMode)
{
	// Body::AddForce(force, position): AddForce(force); AddTorque((position - centre of mass) x force)
	const Body *b = TryGet(id);
	if (b == nullptr) return;
	AddForce(id, inForce, inActivationMode);
	AddTorque(id, (inPoint - b->GetCenterOfMassPosition()).Cross(inForce), inActivationMode);
}

inline void BodyInterface::AddForcesAndTorques(const BodyID *inBodies, int inNumber, const float *inForces, const float *inTorques)
{
	Flush();
	static_assert(sizeof(BodyID) == sizeof(uint32_t), "BodyID must be a plain 32 bit id");
	b2j_bodies_add_force_torque(World(), reinterpret_cast<const uint32_t *>(inBodies), (uint32)inNumber, inForces, inTorques);
}

} // namespace JPH_B200
