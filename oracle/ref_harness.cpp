// ref_harness.cpp -- TEST INFRASTRUCTURE. A C interface around the UNMODIFIED reference (jrouwe/JoltPhysics, compiled by
// oracle/Makefile from /root/reference into oracle/_ref/libjoltref_{det,fast}.so). It builds the benchmark scenes with the
// reference's own API, steps them with JobSystemThreadPool, records what the parity tests compare against (candidate body
// pairs, contact events, manifolds, post-step body state) and can re-create a live reference world inside a b2j_world through
// the C ABI (jolt_adapter.h). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Scenes follow the specifications in PerformanceTest/PyramidScene.h:23-47, PerformanceTest/ConvexVsMeshScene.h:28-117,
// PerformanceTest/MaxBodiesScene.h:44-80 and SURVEY.md 8(d) config 4 (Pile); they are re-stated here, not copied.

#include "jolt_adapter.h"

#include <Jolt/RegisterTypes.h>
#include <Jolt/Core/Factory.h>
#include <Jolt/Core/TempAllocator.h>
#include <Jolt/Core/JobSystemThreadPool.h>
#include <Jolt/Physics/PhysicsSettings.h>
#include <Jolt/Physics/Collision/Shape/ConvexHullShape.h>
#include <Jolt/Physics/Collision/Shape/MeshShape.h>
#include <Jolt/Physics/Collision/BroadPhase/BroadPhase.h>
#include <Jolt/Physics/Collision/ContactListener.h>
#include <Jolt/Physics/Body/BodyActivationListener.h>
#include <Jolt/Physics/Collision/RayCast.h>
#include <Jolt/Physics/Collision/CastResult.h>
#include <Jolt/Physics/Collision/NarrowPhaseQuery.h>
#include <Jolt/Physics/Collision/CollideShape.h>
#include <Jolt/Physics/Constraints/PointConstraint.h>
#include <Jolt/Physics/Constraints/DistanceConstraint.h>
#include <Jolt/Physics/Constraints/HingeConstraint.h>
#include <Jolt/Physics/Constraints/FixedConstraint.h>
#include <Jolt/Physics/Collision/Shape/CylinderShape.h>
#include <Jolt/Physics/Collision/Shape/CapsuleShape.h>
#include <Jolt/Physics/Collision/Shape/SphereShape.h>
#include <Jolt/Physics/Collision/Shape/BoxShape.h>
#include <Jolt/Physics/Collision/CollisionCollectorImpl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <random>
#include <tuple>
#include <initializer_list>
#include <thread>

using namespace JPH;

namespace {

// ---- layers: the 2 object layer / 2 broadphase layer configuration used by PerformanceTest/Layers.h and UnitTests/Layers.h, plus a
// third moving layer (DEBRIS: collides with NON_MOVING and MOVING but not with itself) in its own broadphase tree for the scenes
// that need two moving broadphase layers (TestPhysicsBroadPhaseLayers pattern, UnitTests/Physics/PhysicsTests.cpp:1398)
namespace Layers { constexpr ObjectLayer NON_MOVING = 0, MOVING = 1, DEBRIS = 2, NUM_LAYERS = 3; }
namespace BPLayers { constexpr BroadPhaseLayer NON_MOVING(0), MOVING(1), DEBRIS(2); constexpr uint NUM_LAYERS = 3; }

class OLPairFilter final : public ObjectLayerPairFilter
{
public:
	bool ShouldCollide(ObjectLayer in1, ObjectLayer in2) const override
	{
		if (in1 == Layers::MOVING || in2 == Layers::MOVING) return true;
		return in1 != in2; // NON_MOVING vs DEBRIS
	}
};

class BPLInterface final : public BroadPhaseLayerInterface
{
public:
	uint GetNumBroadPhaseLayers() const override { return BPLayers::NUM_LAYERS; }
	BroadPhaseLayer GetBroadPhaseLayer(ObjectLayer inLayer) const override { return inLayer == Layers::NON_MOVING? BPLayers::NON_MOVING : (inLayer == Layers::MOVING? BPLayers::MOVING : BPLayers::DEBRIS); }
#if defined(JPH_EXTERNAL_PROFILE) || defined(JPH_PROFILE_ENABLED)
	const char *GetBroadPhaseLayerName(BroadPhaseLayer) const override { return "layer"; }
#endif
};

class OVBPFilter final : public ObjectVsBroadPhaseLayerFilter
{
public:
	bool ShouldCollide(ObjectLayer in1, BroadPhaseLayer in2) const override
	{
		if (in1 == Layers::MOVING || in2 == BPLayers::MOVING) return true;
		return (in1 == Layers::NON_MOVING) != (in2 == BPLayers::NON_MOVING); // NON_MOVING vs DEBRIS
	}
};

// ---- recorded contact / activation events
struct ContactRecord
{
	uint32 kind;        // b2j event kind
	uint32 body1, body2, sub1, sub2, num_points;
	float base_offset[3], normal[3], depth;
	float p1[4][3], p2[4][3];
};

class RecordingContactListener final : public ContactListener
{
public:
	void Record(uint32 inKind, const Body &inBody1, const Body &inBody2, const ContactManifold &inManifold)
	{
		ContactRecord r;
		memset(&r, 0, sizeof(r));
		r.kind = inKind;
		r.body1 = inBody1.GetID().GetIndexAndSequenceNumber();
		r.body2 = inBody2.GetID().GetIndexAndSequenceNumber();
		r.sub1 = inManifold.mSubShapeID1.GetValue();
		r.sub2 = inManifold.mSubShapeID2.GetValue();
		r.num_points = (uint32)inManifold.mRelativeContactPointsOn1.size();
		b2j_adapter::sStore(Vec3(inManifold.mBaseOffset), r.base_offset);
		b2j_adapter::sStore(inManifold.mWorldSpaceNormal, r.normal);
		r.depth = inManifold.mPenetrationDepth;
		for (uint32 i = 0; i < r.num_points && i < 4; ++i)
		{
			b2j_adapter::sStore(inManifold.mRelativeContactPointsOn1[i], r.p1[i]);
			b2j_adapter::sStore(inManifold.mRelativeContactPointsOn2[i], r.p2[i]);
		}
		std::lock_guard<std::mutex> lock(mMutex);
		mRecords.push_back(r);
	}

	void OnContactAdded(const Body &inBody1, const Body &inBody2, const ContactManifold &inManifold, ContactSettings &) override { Record(B2J_EVENT_CONTACT_ADDED, inBody1, inBody2, inManifold); }
	void OnContactPersisted(const Body &inBody1, const Body &inBody2, const ContactManifold &inManifold, ContactSettings &) override { Record(B2J_EVENT_CONTACT_PERSISTED, inBody1, inBody2, inManifold); }
	void OnContactRemoved(const SubShapeIDPair &inPair) override
	{
		ContactRecord r;
		memset(&r, 0, sizeof(r));
		r.kind = B2J_EVENT_CONTACT_REMOVED;
		r.body1 = inPair.GetBody1ID().GetIndexAndSequenceNumber();
		r.body2 = inPair.GetBody2ID().GetIndexAndSequenceNumber();
		r.sub1 = inPair.GetSubShapeID1().GetValue();
		r.sub2 = inPair.GetSubShapeID2().GetValue();
		std::lock_guard<std::mutex> lock(mMutex);
		mRecords.push_back(r);
	}

	std::mutex mMutex;
	std::vector<ContactRecord> mRecords;
};

class RecordingActivationListener final : public BodyActivationListener
{
public:
	void OnBodyActivated(const BodyID &inID, uint64) override { std::lock_guard<std::mutex> lock(mMutex); mRecords.push_back({ B2J_EVENT_BODY_ACTIVATED, inID.GetIndexAndSequenceNumber() }); }
	void OnBodyDeactivated(const BodyID &inID, uint64) override { std::lock_guard<std::mutex> lock(mMutex); mRecords.push_back({ B2J_EVENT_BODY_DEACTIVATED, inID.GetIndexAndSequenceNumber() }); }
	std::mutex mMutex;
	std::vector<b2j_activation_event> mRecords;
};

struct World
{
	BPLInterface bpl;
	OVBPFilter ovbp;
	OLPairFilter olp;
	PhysicsSystem system;
	TempAllocator *temp = nullptr;
	JobSystemThreadPool *jobs = nullptr;
	int jobs_threads = -1;
	RecordingContactListener contacts;
	RecordingActivationListener activations;
	uint max_body_pairs = 0, max_contact_constraints = 0;
	uint num_dynamic = 0;
	std::vector<BodyID> tour_bodies; // api_tour scene

	~World() { delete jobs; delete temp; }
};

std::once_flag sInitFlag;
b2j_adapter::Api sApi;
String sLastError;

void sInit()
{
	std::call_once(sInitFlag, []() {
		RegisterDefaultAllocator();
		Factory::sInstance = new Factory();
		RegisterTypes();
	});
}

World *sNewWorld(uint inMaxBodies, uint inMaxBodyPairs, uint inMaxContactConstraints, uint inTempMB)
{
	sInit();
	World *w = new World;
	w->max_body_pairs = inMaxBodyPairs;
	w->max_contact_constraints = inMaxContactConstraints;
	w->system.Init(inMaxBodies, 0, inMaxBodyPairs, inMaxContactConstraints, w->bpl, w->ovbp, w->olp);
	if (inTempMB == 0)
		w->temp = new TempAllocatorMalloc();
	else
		w->temp = new TempAllocatorImpl(size_t(inTempMB) * 1024 * 1024);
	return w;
}

void sEnsureJobs(World *w, int inNumThreads)
{
	if (inNumThreads <= 0)
		inNumThreads = (int)std::thread::hardware_concurrency();
	if (w->jobs == nullptr || w->jobs_threads != inNumThreads)
	{
		delete w->jobs;
		// N threads = N-1 workers + the calling thread (PerformanceTest/PerformanceTest.cpp:314)
		w->jobs = new JobSystemThreadPool(cMaxPhysicsJobs, cMaxPhysicsBarriers, inNumThreads - 1);
		w->jobs_threads = inNumThreads;
	}
}

// ---- scenes -------------------------------------------------------------------------------------------------

World *sScenePyramid(int inHeight, int inTightLimit = 0)
{
	// PerformanceTest/PyramidScene.h:23-47 (height 15 -> 1240 boxes); limits as PerformanceTest.cpp (10240 bodies, 65536 pairs, 20480 constraints)
	// inTightLimit > 0: max body pairs = max contact constraints = that value (error path tests: the step must report the overflow)
	World *w = inTightLimit > 0? sNewWorld(10240, (uint)inTightLimit, (uint)inTightLimit, 32) : sNewWorld(10240, 65536, 20480, 32);
	BodyInterface &bi = w->system.GetBodyInterface();
	bi.CreateAndAddBody(BodyCreationSettings(new BoxShape(Vec3(50.0f, 1.0f, 50.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING), EActivation::DontActivate);
	const float box_size = 2.0f, separation = 0.5f, half = 1.0f;
	RefConst<Shape> box = new BoxShape(Vec3::sReplicate(half), 0.0f);
	for (int i = 0; i < inHeight; ++i)
		for (int j = i / 2; j < inHeight - (i + 1) / 2; ++j)
			for (int k = i / 2; k < inHeight - (i + 1) / 2; ++k)
			{
				RVec3 pos(float(-inHeight) + box_size * j + ((i & 1)? half : 0.0f), 1.0f + (box_size + separation) * i, float(-inHeight) + box_size * k + ((i & 1)? half : 0.0f));
				BodyCreationSettings s(box, pos, Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
				s.mAllowSleeping = false;
				bi.CreateAndAddBody(s, EActivation::Activate);
				w->num_dynamic++;
			}
	return w;
}

Ref<Shape> sTerrainMesh(int n, float cell_size, float max_height)
{
	VertexList vertices;
	vertices.resize((n + 1) * (n + 1));
	for (int x = 0; x <= n; ++x)
		for (int z = 0; z <= n; ++z)
		{
			float height = Sin(float(x) * 50.0f / n) * Cos(float(z) * 50.0f / n);
			vertices[z * (n + 1) + x] = Float3(cell_size * x, max_height * height, cell_size * z);
		}
	IndexedTriangleList indices;
	indices.resize(n * n * 2);
	IndexedTriangle *next = indices.data();
	for (int x = 0; x < n; ++x)
		for (int z = 0; z < n; ++z)
		{
			int start = (n + 1) * z + x;
			next->mIdx[0] = start; next->mIdx[1] = start + n + 1; next->mIdx[2] = start + 1; next++;
			next->mIdx[0] = start + 1; next->mIdx[1] = start + n + 1; next->mIdx[2] = start + n + 2; next++;
		}
	Ref<MeshShapeSettings> settings = new MeshShapeSettings(vertices, indices);
	settings->mMaxTrianglesPerLeaf = 4;
	return settings->Create().Get();
}

World *sSceneConvexVsMesh(int inHalfGrid, int inDecorated = 0)
{
	// PerformanceTest/ConvexVsMeshScene.h:28-117 (half grid 10 -> 21*4*21 = 1764 bodies on a 20000 triangle mesh)
	World *w = sNewWorld(10240, 65536, 20480, 32);
	PhysicsSettings settings = w->system.GetPhysicsSettings();
	settings.mNumVelocitySteps = 4;
	settings.mNumPositionSteps = 1;
	w->system.SetPhysicsSettings(settings);

	const int n = 100;
	const float cell_size = 3.0f, max_height = 5.0f, center = n * cell_size / 2;
	RefConst<Shape> terrain = sTerrainMesh(n, cell_size, max_height);
	if ((inDecorated & 2) != 0) // the mesh itself scaled (non uniform) and rotated a little
		terrain = new RotatedTranslatedShape(Vec3(1.0f, 0.5f, -2.0f), Quat(0.04361939f, 0.0f, 0.0f, 0.99904822f), new ScaledShape(terrain, Vec3(1.2f, 0.7f, 0.9f)));
	BodyCreationSettings mesh(terrain, RVec3(-center, max_height, -center), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
	mesh.mFriction = 0.5f;
	mesh.mRestitution = 0.6f;
	BodyInterface &bi = w->system.GetBodyInterface();
	bi.CreateAndAddBody(mesh, EActivation::DontActivate);

	Array<Ref<Shape>> shapes = {
		new BoxShape(Vec3(0.5f, 0.75f, 1.0f)),
		new SphereShape(0.5f),
		new CapsuleShape(0.75f, 0.5f),
		ConvexHullShapeSettings({ Vec3(0, 1, 0), Vec3(1, 0, 0), Vec3(-1, 0, 0), Vec3(0, 0, 1), Vec3(0, 0, -1) }).Create().Get(),
	};
	if ((inDecorated & 1) != 0)
	{
		// the same bodies behind ScaledShape / RotatedTranslatedShape decorators (decorated convex shapes against the mesh)
		Quat tilt = Quat(0.0f, 0.38268343f, 0.0f, 0.92387953f), roll = Quat(0.25881905f, 0.0f, 0.0f, 0.96592583f);
		shapes[0] = new ScaledShape(shapes[0], Vec3(1.3f, 0.6f, 0.9f));
		shapes[1] = new RotatedTranslatedShape(Vec3(0.2f, 0.1f, 0.0f), roll, new ScaledShape(shapes[1], Vec3::sReplicate(1.3f)));
		shapes[2] = new RotatedTranslatedShape(Vec3(0.0f, 0.3f, 0.1f), tilt * roll, shapes[2]);
		shapes[3] = new ScaledShape(new RotatedTranslatedShape(Vec3(0.1f, 0.0f, -0.2f), tilt, shapes[3]), Vec3::sReplicate(0.8f));
	}
	for (int x = -inHalfGrid; x <= inHalfGrid; ++x)
		for (int y = 0; y < (int)shapes.size(); ++y)
			for (int z = -inHalfGrid; z <= inHalfGrid; ++z)
			{
				BodyCreationSettings s;
				s.mMotionType = EMotionType::Dynamic;
				s.mObjectLayer = Layers::MOVING;
				s.mPosition = RVec3(7.5f * x, 15.0f + 2.0f * y, 7.5f * z);
				s.mFriction = 0.5f;
				s.mRestitution = 0.6f;
				s.SetShape(shapes[y]);
				bi.CreateAndAddBody(s, EActivation::Activate);
				w->num_dynamic++;
			}
	return w;
}

World *sSceneMaxBodies(int inNumBodies)
{
	// PerformanceTest/MaxBodiesScene.h:44-80 restated for N bodies: unit boxes of mass 1000 on a cubic grid, x neighbours touching.
	uint n = (uint)inNumBodies;
	World *w = sNewWorld(n, std::max(65536u, n), std::max(20480u, n), 0);
	PhysicsSettings settings = w->system.GetPhysicsSettings();
	settings.mNumVelocitySteps = 4;
	settings.mNumPositionSteps = 1;
	w->system.SetPhysicsSettings(settings);
	BodyInterface &bi = w->system.GetBodyInterface();
	uint side = (uint)ceil(cbrt(double(n)));
	BodyCreationSettings s;
	s.SetShape(new BoxShape(Vec3::sReplicate(0.5f)));
	s.mMotionType = EMotionType::Dynamic;
	s.mObjectLayer = Layers::MOVING;
	s.mOverrideMassProperties = EOverrideMassProperties::CalculateInertia;
	s.mMassPropertiesOverride.mMass = 1000.0f;
	std::vector<BodyID> ids;
	ids.reserve(n);
	uint count = 0;
	for (uint x = 0; x < side && count < n; ++x)
		for (uint y = 0; y < side && count < n; ++y)
			for (uint z = 0; z < side && count < n; ++z, ++count)
			{
				s.mPosition = RVec3(1.0f * x, 3.0f * y, 3.0f * z);
				ids.push_back(bi.CreateBody(s)->GetID());
			}
	BodyInterface::AddState state = bi.AddBodiesPrepare(ids.data(), (int)ids.size());
	bi.AddBodiesFinalize(ids.data(), (int)ids.size(), state, EActivation::Activate);
	w->num_dynamic = n;
	return w;
}

Ref<Shape> sRandomHull(std::mt19937 &ioRandom, float inExtent)
{
	std::uniform_real_distribution<float> d(-inExtent, inExtent);
	Array<Vec3> points;
	for (int i = 0; i < 12; ++i)
	{
		float x = d(ioRandom), y = d(ioRandom), z = d(ioRandom);
		points.push_back(Vec3(x, y, z));
	}
	return ConvexHullShapeSettings(points).Create().Get();
}

Quat sRandomQuat(std::mt19937 &ioRandom)
{
	std::normal_distribution<float> n(0.0f, 1.0f);
	float x = n(ioRandom), y = n(ioRandom), z = n(ioRandom), q = n(ioRandom);
	return Quat(x, y, z, q).Normalized();
}

World *sScenePile(int inNumBodies, int inShapeMask)
{
	// seed 12345; container of 5 static boxes; bodies i mod 4 -> sphere / box / capsule / 12 point hull on a
	// jittered cubic grid with spacing 1.15; friction 0.5, restitution 0.1. inShapeMask selects which of the 4 kinds are used (bit per kind).
	uint n = (uint)inNumBodies;
	uint side = (uint)ceil(cbrt(double(n)));
	float spacing = 1.15f;
	float half_width = 0.5f * side * spacing + 2.0f;
	World *w = sNewWorld(n + 128, std::max(65536u, 16 * n), std::max(20480u, 8 * n), 0);
	BodyInterface &bi = w->system.GetBodyInterface();

	auto add_static = [&](Vec3 inHalfExtent, Vec3 inPos) {
		BodyCreationSettings s(new BoxShape(inHalfExtent), RVec3(inPos), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		s.mFriction = 0.5f; s.mRestitution = 0.1f;
		bi.CreateAndAddBody(s, EActivation::DontActivate);
	};
	float wall_h = 0.5f * side * spacing + 2.0f;
	add_static(Vec3(half_width + 2.0f, 1.0f, half_width + 2.0f), Vec3(0, -1.0f, 0));
	add_static(Vec3(1.0f, wall_h, half_width + 2.0f), Vec3(-half_width - 1.0f, wall_h, 0));
	add_static(Vec3(1.0f, wall_h, half_width + 2.0f), Vec3(half_width + 1.0f, wall_h, 0));
	add_static(Vec3(half_width + 2.0f, wall_h, 1.0f), Vec3(0, wall_h, -half_width - 1.0f));
	add_static(Vec3(half_width + 2.0f, wall_h, 1.0f), Vec3(0, wall_h, half_width + 1.0f));

	// palette of 256 random 12 point hulls (own generator so that the body loop below only draws jitter + rotation)
	std::mt19937 hull_random(4242);
	Array<Ref<Shape>> hulls;
	if (inShapeMask <= 0 || (inShapeMask & 8))
		for (int i = 0; i < 256; ++i)
			hulls.push_back(sRandomHull(hull_random, 0.5f));
	uint hull_idx = 0;

	std::mt19937 random(12345);
	std::uniform_real_distribution<float> jitter(-0.05f, 0.05f);
	Ref<Shape> sphere = new SphereShape(0.5f);
	Ref<Shape> box = new BoxShape(Vec3(0.5f, 0.4f, 0.3f), 0.05f);
	Ref<Shape> capsule = new CapsuleShape(0.4f, 0.3f);
	std::vector<int> kinds;
	for (int k = 0; k < 4; ++k)
		if (inShapeMask & (1 << k))
			kinds.push_back(k);
	if (kinds.empty())
		kinds = { 0, 1, 2, 3 };

	std::vector<BodyID> ids;
	ids.reserve(n);
	uint count = 0;
	for (uint y = 0; y < side && count < n; ++y)
		for (uint x = 0; x < side && count < n; ++x)
			for (uint z = 0; z < side && count < n; ++z, ++count)
			{
				BodyCreationSettings s;
				switch (kinds[count % kinds.size()])
				{
				case 0: s.SetShape(sphere); break;
				case 1: s.SetShape(box); break;
				case 2: s.SetShape(capsule); break;
				default: s.SetShape(hulls[hull_idx++ % hulls.size()]); break;
				}
				s.mMotionType = EMotionType::Dynamic;
				s.mObjectLayer = Layers::MOVING;
				float jx = jitter(random), jy = jitter(random), jz = jitter(random);
				s.mPosition = RVec3((float(x) - 0.5f * (side - 1)) * spacing + jx, 0.8f + float(y) * spacing + jy, (float(z) - 0.5f * (side - 1)) * spacing + jz);
				s.mRotation = sRandomQuat(random);
				s.mFriction = 0.5f;
				s.mRestitution = 0.1f;
				ids.push_back(bi.CreateBody(s)->GetID());
			}
	BodyInterface::AddState state = bi.AddBodiesPrepare(ids.data(), (int)ids.size());
	bi.AddBodiesFinalize(ids.data(), (int)ids.size(), state, EActivation::Activate);
	w->num_dynamic = n;
	return w;
}

World *sSceneSmallStack(int inVariant)
{
	// Small scenes for fast unit-level parity: a floor and a handful of bodies (variant picks the shapes), allows sleeping.
	World *w = sNewWorld(1024, 4096, 1024, 4);
	BodyInterface &bi = w->system.GetBodyInterface();
	bi.CreateAndAddBody(BodyCreationSettings(new BoxShape(Vec3(100.0f, 1.0f, 100.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING), EActivation::DontActivate);
	std::mt19937 random(777 + inVariant);
	Ref<Shape> shapes[4] = { new SphereShape(0.5f), new BoxShape(Vec3(0.5f, 0.5f, 0.5f)), new CapsuleShape(0.5f, 0.3f), sRandomHull(random, 0.6f) };
	for (int i = 0; i < 24; ++i)
	{
		int kind = inVariant < 4? inVariant : i % 4;
		BodyCreationSettings s(shapes[kind], RVec3(float(i % 3) * 0.9f - 0.9f, 0.6f + 1.1f * float(i / 3), float(i % 2) * 0.4f), inVariant < 4 && kind == 1? Quat::sIdentity() : sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
		s.mFriction = 0.5f;
		s.mRestitution = (i % 5 == 0)? 0.5f : 0.0f;
		bi.CreateAndAddBody(s, EActivation::Activate);
		w->num_dynamic++;
	}
	return w;
}

// Feature scenes: joltphysics_b200/host/feature_scenes.inl, the same user code the facade compiles (every motion type / body flag /
// per body override the step has a branch for)
const char *sFeatureNames[] = { "kinematic", "sensor", "dof_plane2d", "gyroscopic", "step_overrides", "no_manifold_reduction", "two_moving_layers", "kinematic_vs_nondynamic", "zoo" };
#define B2J_SHAPE_REF RefConst<Shape>
#define B2J_NEW_SHAPE(Type, ...) new Type(__VA_ARGS__)
#include "../joltphysics_b200/host/feature_scenes.inl"

World *sSceneFeature(int inVariant)
{
	World *w = sNewWorld(1024, 8192, 4096, 8);
	std::mt19937 hull_random(4242); // the first hull of the pile's palette (bench_assets/pile_hulls.b2js holds it cooked for the facade)
	RefConst<Shape> hull = sRandomHull(hull_random, 0.5f);
	uint32_t num_dynamic = 0;
	sFeatureCreate(w->system, inVariant, hull, num_dynamic);
	w->num_dynamic = num_dynamic;
	return w;
}

#define B2J_CREATE_COMPOUND(settings) RefConst<Shape>((settings).Create().Get())
#include "../joltphysics_b200/host/compound_scene.inl"

World *sSceneCompound(int inVariant)
{
	// joltphysics_b200/host/compound_scene.inl (user code shared with the facade build); variant 1 puts the bodies on a terrain mesh
	World *w = sNewWorld(1024, 16384, 8192, 0);
	BodyInterface &bi = w->system.GetBodyInterface();
	if (inVariant == 1)
	{
		const int n = 20;
		const float cell_size = 3.0f, max_height = 2.0f, center = n * cell_size / 2;
		BodyCreationSettings mesh(sTerrainMesh(n, cell_size, max_height), RVec3(-center, max_height, -center), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		mesh.mFriction = 0.5f;
		bi.CreateAndAddBody(mesh, EActivation::DontActivate);
	}
	std::mt19937 hull_random(4242);
	uint32_t num_dynamic = 0;
	sCompoundCreate(w->system, inVariant, sRandomHull(hull_random, 0.5f), num_dynamic);
	w->num_dynamic += (int)num_dynamic;
	return w;
}


// the facade's API tour (same user code on both sides)
#include "../joltphysics_b200/host/api_tour.inl"

World *sSceneApiTour()
{
	World *w = sNewWorld(1024, 4096, 1024, 4);
	sApiTourCreate(w->system, w->tour_bodies);
	w->num_dynamic = (uint)w->tour_bodies.size();
	return w;
}

} // namespace

// ---- C interface ------------------------------------------------------------------------------------------------

extern "C" {

const char *jref_last_error() { return sLastError.c_str(); }

// 0 on success. Resolves the b2j C ABI from a shared library (the product libjolt_b200.so on a GPU box).
int jref_bind_b2j(const char *inLibPath)
{
	sInit();
	return sApi.Load(inLibPath, sLastError)? 0 : -1;
}

void *jref_create_scene(const char *inName, int inParam0, int inParam1)
{
	String name(inName);
	World *w = nullptr;
	if (name == "pyramid") w = sScenePyramid(inParam0 > 0? inParam0 : 15);
	else if (name == "pyramid_tight") w = sScenePyramid(inParam0 > 0? inParam0 : 6, inParam1 > 0? inParam1 : 64);
	else if (name == "convex_vs_mesh") w = sSceneConvexVsMesh(inParam0 > 0? inParam0 : 10, inParam1);
	else if (name == "max_bodies") w = sSceneMaxBodies(inParam0 > 0? inParam0 : 10000);
	else if (name == "pile") w = sScenePile(inParam0 > 0? inParam0 : 1000, inParam1 > 0? inParam1 : 15);
	else if (name == "small_stack") w = sSceneSmallStack(inParam0);
	else if (name == "api_tour") w = sSceneApiTour();
	else if (name == "feature") w = sSceneFeature(inParam0);
	else if (name == "compound") w = sSceneCompound(inParam0);
	else { sLastError = "unknown scene"; return nullptr; }
	w->system.SetContactListener(&w->contacts);
	w->system.SetBodyActivationListener(&w->activations);
	w->system.OptimizeBroadPhase();
	return w;
}

void jref_destroy(void *h) { delete (World *)h; }

// api_tour scene: the mutation phases and the queries of api_tour.inl
void jref_mutate(void *h, int inPhase) { World *w = (World *)h; sApiTourMutate(w->system, w->tour_bodies, inPhase); }
int jref_query(void *h, uint32_t *outIDs, int inCapacity, uint32_t *outNumBodies, uint32_t *outFlags) { World *w = (World *)h; return sApiTourQuery(w->system, w->tour_bodies, outIDs, inCapacity, outNumBodies, outFlags); }

// Destroys the body in slot inIndex and creates a SPHERE in its place: same slot (the freed index is reused first), same centre of mass
// position, next sequence number. The contact cache still holds the pairs of the destroyed body (a cache keyed by body
// slots alone would serve the new body from them). Returns the new id or 0xffffffff.
uint32_t jref_replace_body(void *h, uint32_t inIndex, float inRadius)
{
	World *w = (World *)h;
	const BodyVector &bodies = w->system.mBodyManager.GetBodies();
	if (inIndex >= bodies.size() || !BodyManager::sIsValidBodyPointer(bodies[inIndex])) return 0xffffffffu;
	const Body *old_body = bodies[inIndex];
	BodyCreationSettings s(new SphereShape(inRadius), old_body->GetCenterOfMassPosition(), old_body->GetRotation(), old_body->GetMotionType(), old_body->GetObjectLayer());
	s.mFriction = old_body->GetFriction();
	bool active = old_body->IsActive();
	BodyID old_id = old_body->GetID();
	BodyInterface &bi = w->system.GetBodyInterface();
	bi.RemoveBody(old_id);
	bi.DestroyBody(old_id);
	BodyID id = bi.CreateAndAddBody(s, active? EActivation::Activate : EActivation::DontActivate);
	return id.GetIndex() == inIndex? id.GetIndexAndSequenceNumber() : 0xffffffffu;
}

// NarrowPhaseQuery::CastRay (closest hit) for n rays; inObjectLayer = the layer the rays collide as (0xffffffff: no layer filtering)
void jref_cast_rays(void *h, const b2j_ray *inRays, uint32_t inNum, uint32_t inObjectLayer, b2j_ray_hit *outHits)
{
	World *w = (World *)h;
	const NarrowPhaseQuery &query = w->system.GetNarrowPhaseQueryNoLock();
	for (uint32_t i = 0; i < inNum; ++i)
	{
		RRayCast ray(RVec3(inRays[i].origin[0], inRays[i].origin[1], inRays[i].origin[2]), Vec3(inRays[i].direction[0], inRays[i].direction[1], inRays[i].direction[2]));
		RayCastResult hit;
		bool had_hit;
		if (inObjectLayer == 0xffffffffu)
			had_hit = query.CastRay(ray, hit);
		else
			had_hit = query.CastRay(ray, hit, DefaultBroadPhaseLayerFilter(w->ovbp, (ObjectLayer)inObjectLayer), DefaultObjectLayerFilter(w->olp, (ObjectLayer)inObjectLayer));
		outHits[i].body = had_hit? hit.mBodyID.GetIndexAndSequenceNumber() : 0xffffffffu;
		outHits[i].sub_shape = had_hit? hit.mSubShapeID2.GetValue() : 0xffffffffu;
		outHits[i].fraction = had_hit? hit.mFraction : 1.0f + FLT_EPSILON;
	}
}

// Non contact constraints of the world by Constraint::mConstraintIndex: count, the state Constraint::SaveState writes (accumulated
// impulses, the distance constraint's normal), removal (ConstraintManager::Remove: the last constraint takes the index), enabling
uint32_t jref_num_constraints(void *h) { return (uint32_t)((World *)h)->system.GetConstraints().size(); }
uint32_t jref_get_constraint_states(void *h, b2j_constraint_state *outStates, uint32_t inCapacity)
{
	Constraints constraints = ((World *)h)->system.GetConstraints();
	for (uint32_t i = 0; i < constraints.size() && i < inCapacity; ++i)
	{
		b2j_constraint_state &o = outStates[i];
		memset(&o, 0, sizeof(o));
		const Constraint *c = constraints[i];
		if (c->GetSubType() == EConstraintSubType::Point)
			static_cast<const PointConstraint *>(c)->GetTotalLambdaPosition().StoreFloat3((Float3 *)o.total_lambda);
		else if (c->GetSubType() == EConstraintSubType::Distance)
		{
			o.total_lambda[0] = static_cast<const DistanceConstraint *>(c)->GetTotalLambdaPosition();
			static_cast<const DistanceConstraint *>(c)->mWorldSpaceNormal.StoreFloat3((Float3 *)o.world_space_normal);
		}
		else if (c->GetSubType() == EConstraintSubType::Hinge)
		{
			const HingeConstraint *hc = static_cast<const HingeConstraint *>(c);
			hc->GetTotalLambdaPosition().StoreFloat3((Float3 *)o.total_lambda);
			o.total_lambda_rotation[0] = hc->GetTotalLambdaRotation()[0]; o.total_lambda_rotation[1] = hc->GetTotalLambdaRotation()[1];
			o.total_lambda_limits = hc->GetTotalLambdaRotationLimits(); o.total_lambda_motor = hc->GetTotalLambdaMotor();
		}
		else if (c->GetSubType() == EConstraintSubType::Fixed)
		{
			static_cast<const FixedConstraint *>(c)->GetTotalLambdaPosition().StoreFloat3((Float3 *)o.total_lambda);
			static_cast<const FixedConstraint *>(c)->GetTotalLambdaRotation().StoreFloat3((Float3 *)o.total_lambda_rotation);
		}
	}
	return (uint32_t)constraints.size();
}
void jref_remove_constraint(void *h, uint32_t inIndex)
{
	World *w = (World *)h;
	Constraints constraints = w->system.GetConstraints();
	if (inIndex < constraints.size()) w->system.RemoveConstraint(constraints[inIndex]);
}
void jref_set_constraint_enabled(void *h, uint32_t inIndex, int inEnabled)
{
	Constraints constraints = ((World *)h)->system.GetConstraints();
	if (inIndex < constraints.size()) constraints[inIndex]->SetEnabled(inEnabled != 0);
}

// NarrowPhaseQuery::CollideShape with an AllHitCollisionCollector for n queries of ONE convex query shape (kind: 0 sphere (p0 = radius),
// 1 box (p0..2 = half extent, p3 = convex radius), 2 capsule (p0 = half height, p1 = radius), 3 cylinder (p0 = half height, p1 = radius,
// p2 = convex radius)); queries as b2j_shape_query (the shape member is ignored). Hits in the order the collector received them.
void jref_collide_shape(void *h, uint32_t inKind, const float *inParams, const b2j_shape_query *inQueries, uint32_t inNum, float inMaxSeparationDistance,
	uint32_t inObjectLayer, uint32_t inMaxHits, uint32_t *outCounts, b2j_collide_shape_hit *outHits)
{
	World *w = (World *)h;
	RefConst<Shape> shape;
	switch (inKind)
	{
	case 0: shape = new SphereShape(inParams[0]); break;
	case 1: shape = new BoxShape(Vec3(inParams[0], inParams[1], inParams[2]), inParams[3]); break;
	case 2: shape = new CapsuleShape(inParams[0], inParams[1]); break;
	default: shape = new CylinderShape(inParams[0], inParams[1], inParams[2]); break;
	}
	const NarrowPhaseQuery &query = w->system.GetNarrowPhaseQueryNoLock();
	CollideShapeSettings settings;
	settings.mMaxSeparationDistance = inMaxSeparationDistance;
	for (uint32_t i = 0; i < inNum; ++i)
	{
		const b2j_shape_query &q = inQueries[i];
		RMat44 com = RMat44::sRotationTranslation(Quat(q.rotation[0], q.rotation[1], q.rotation[2], q.rotation[3]), RVec3(q.position[0], q.position[1], q.position[2]));
		AllHitCollisionCollector<CollideShapeCollector> collector;
		RVec3 base(q.base_offset[0], q.base_offset[1], q.base_offset[2]);
		if (inObjectLayer == 0xffffffffu)
			query.CollideShape(shape, Vec3::sOne(), com, settings, base, collector);
		else
			query.CollideShape(shape, Vec3::sOne(), com, settings, base, collector, DefaultBroadPhaseLayerFilter(w->ovbp, (ObjectLayer)inObjectLayer), DefaultObjectLayerFilter(w->olp, (ObjectLayer)inObjectLayer));
		outCounts[i] = (uint32_t)collector.mHits.size();
		for (uint32_t j = 0; j < collector.mHits.size() && j < inMaxHits; ++j)
		{
			const CollideShapeResult &r = collector.mHits[j];
			b2j_collide_shape_hit &o = outHits[(size_t)i * inMaxHits + j];
			o.body = r.mBodyID2.GetIndexAndSequenceNumber();
			o.sub_shape1 = r.mSubShapeID1.GetValue(); o.sub_shape2 = r.mSubShapeID2.GetValue();
			o.penetration_depth = r.mPenetrationDepth;
			r.mContactPointOn1.StoreFloat3((Float3 *)o.point1); r.mContactPointOn2.StoreFloat3((Float3 *)o.point2); r.mPenetrationAxis.StoreFloat3((Float3 *)o.axis);
		}
	}
}

// BroadPhaseQuery::CollideSphere (inMode 1: [n][4] centre, radius) / CollidePoint (inMode 2: [n][3]), narrowed to the true bounds like
// jref_collide_aabox
void jref_collide_volume(void *h, uint32_t inMode, const float *inData, uint32_t inNum, uint32_t inObjectLayer, uint32_t inMaxHits, uint32_t *outCounts, uint32_t *outIDs)
{
	World *w = (World *)h;
	const BodyLockInterfaceNoLock &li = w->system.GetBodyLockInterfaceNoLock();
	for (uint32_t i = 0; i < inNum; ++i)
	{
		AllHitCollisionCollector<CollideShapeBodyCollector> collector;
		const float *d = inData + (inMode == 1? 4 : 3) * (size_t)i;
		Vec3 centre(d[0], d[1], d[2]);
		float radius = inMode == 1? d[3] : 0.0f;
		DefaultBroadPhaseLayerFilter bpf(w->ovbp, (ObjectLayer)(inObjectLayer == 0xffffffffu? 0 : inObjectLayer));
		DefaultObjectLayerFilter olf(w->olp, (ObjectLayer)(inObjectLayer == 0xffffffffu? 0 : inObjectLayer));
		BroadPhaseLayerFilter all_bp; ObjectLayerFilter all_ol;
		const BroadPhaseLayerFilter &f1 = inObjectLayer == 0xffffffffu? all_bp : (const BroadPhaseLayerFilter &)bpf;
		const ObjectLayerFilter &f2 = inObjectLayer == 0xffffffffu? all_ol : (const ObjectLayerFilter &)olf;
		if (inMode == 1)
			w->system.GetBroadPhaseQuery().CollideSphere(centre, radius, collector, f1, f2);
		else
			w->system.GetBroadPhaseQuery().CollidePoint(centre, collector, f1, f2);
		std::vector<uint32_t> ids;
		for (const BodyID &id : collector.mHits)
		{
			const Body *b = li.TryGetBody(id);
			if (b == nullptr) continue;
			const AABox &bounds = b->GetWorldSpaceBounds();
			if (bounds.GetSqDistanceTo(centre) <= radius * radius) ids.push_back(id.GetIndexAndSequenceNumber());
		}
		std::sort(ids.begin(), ids.end());
		outCounts[i] = (uint32_t)ids.size();
		for (uint32_t j = 0; j < ids.size() && j < inMaxHits; ++j) outIDs[(size_t)i * inMaxHits + j] = ids[j];
	}
}

// BroadPhaseQuery::CollideAABox for n boxes, narrowed to the bodies whose TRUE world space bounds overlap (the tree keeps widened
// bounds of moving bodies, so the raw result is a superset). outIDs: [n][inMaxHits] sorted ascending.
void jref_collide_aabox(void *h, const float *inBoxes, uint32_t inNum, uint32_t inObjectLayer, uint32_t inMaxHits, uint32_t *outCounts, uint32_t *outIDs)
{
	World *w = (World *)h;
	const BodyLockInterfaceNoLock &li = w->system.GetBodyLockInterfaceNoLock();
	for (uint32_t i = 0; i < inNum; ++i)
	{
		AABox box(Vec3(inBoxes[6 * i], inBoxes[6 * i + 1], inBoxes[6 * i + 2]), Vec3(inBoxes[6 * i + 3], inBoxes[6 * i + 4], inBoxes[6 * i + 5]));
		AllHitCollisionCollector<CollideShapeBodyCollector> collector;
		if (inObjectLayer == 0xffffffffu)
			w->system.GetBroadPhaseQuery().CollideAABox(box, collector);
		else
			w->system.GetBroadPhaseQuery().CollideAABox(box, collector, DefaultBroadPhaseLayerFilter(w->ovbp, (ObjectLayer)inObjectLayer), DefaultObjectLayerFilter(w->olp, (ObjectLayer)inObjectLayer));
		std::vector<uint32_t> ids;
		for (const BodyID &id : collector.mHits)
		{
			const Body *b = li.TryGetBody(id);
			if (b != nullptr && b->GetWorldSpaceBounds().Overlaps(box)) ids.push_back(id.GetIndexAndSequenceNumber());
		}
		std::sort(ids.begin(), ids.end());
		outCounts[i] = (uint32_t)ids.size();
		for (uint32_t j = 0; j < ids.size() && j < inMaxHits; ++j) outIDs[(size_t)i * inMaxHits + j] = ids[j];
	}
}

// Turn the recording listeners off (timing runs) or on.
void jref_set_recording(void *h, int inOn)
{
	World *w = (World *)h;
	w->system.SetContactListener(inOn? &w->contacts : nullptr);
	w->system.SetBodyActivationListener(inOn? &w->activations : nullptr);
}

// PhysicsSystem::Update; clears the event records first. Returns EPhysicsUpdateError bits.
int jref_step(void *h, float inDeltaTime, int inCollisionSteps, int inNumThreads)
{
	World *w = (World *)h;
	sEnsureJobs(w, inNumThreads);
	w->contacts.mRecords.clear();
	w->activations.mRecords.clear();
	return (int)w->system.Update(inDeltaTime, inCollisionSteps, w->temp, w->jobs);
}

// Times inNumSteps calls of Update(dt, 1) exactly as PerformanceTest.cpp:380-391 does (chrono around Update only). Returns seconds.
double jref_time_steps(void *h, float inDeltaTime, int inNumSteps, int inNumThreads)
{
	World *w = (World *)h;
	sEnsureJobs(w, inNumThreads);
	double total = 0.0;
	for (int i = 0; i < inNumSteps; ++i)
	{
		w->contacts.mRecords.clear();
		w->activations.mRecords.clear();
		auto t0 = std::chrono::high_resolution_clock::now();
		w->system.Update(inDeltaTime, 1, w->temp, w->jobs);
		auto t1 = std::chrono::high_resolution_clock::now();
		total += std::chrono::duration<double>(t1 - t0).count();
	}
	return total;
}

int jref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// CPU side of the batched-worlds config: inNumThreads host threads each own one independent world of the
// scene and step it inNumSteps times with a single threaded job system (no cross thread synchronisation: the best case for the
// CPU when there are many small worlds, TestMultiplePhysicsSystems pattern). inWarmup untimed steps first.
// Returns the wall time in seconds of the timed part; total work = inNumThreads * inNumSteps world steps.
double jref_time_worlds_parallel(const char *inName, int inParam0, int inParam1, int inNumThreads, int inWarmup, int inNumSteps, float inDeltaTime)
{
	sInit();
	if (inNumThreads <= 0) inNumThreads = (int)std::thread::hardware_concurrency();
	std::vector<World *> worlds(inNumThreads, nullptr);
	std::vector<std::thread> threads;
	std::atomic<int> ready { 0 };
	std::atomic<bool> go { false };
	std::vector<double> seconds(inNumThreads, 0.0);
	String name(inName);
	for (int t = 0; t < inNumThreads; ++t)
		threads.emplace_back([&, t]() {
			World *w = nullptr;
			if (name == "pyramid") w = sScenePyramid(inParam0 > 0? inParam0 : 15);
			else if (name == "convex_vs_mesh") w = sSceneConvexVsMesh(inParam0 > 0? inParam0 : 10);
			else w = sScenePile(inParam0 > 0? inParam0 : 1000, inParam1 > 0? inParam1 : 15);
			w->system.OptimizeBroadPhase();
			sEnsureJobs(w, 1);
			for (int i = 0; i < inWarmup; ++i) w->system.Update(inDeltaTime, 1, w->temp, w->jobs);
			ready++;
			while (!go.load()) std::this_thread::yield();
			auto t0 = std::chrono::high_resolution_clock::now();
			for (int i = 0; i < inNumSteps; ++i) w->system.Update(inDeltaTime, 1, w->temp, w->jobs);
			seconds[t] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
			worlds[t] = w;
		});
	while (ready.load() < inNumThreads) std::this_thread::yield();
	auto t0 = std::chrono::high_resolution_clock::now();
	go = true;
	for (std::thread &t : threads) t.join();
	double wall = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
	for (World *w : worlds) delete w;
	return wall;
}
uint32_t jref_num_bodies(void *h) { return ((World *)h)->system.GetNumBodies(); }
uint32_t jref_num_dynamic(void *h) { return ((World *)h)->num_dynamic; }
uint32_t jref_num_active(void *h) { return ((World *)h)->system.GetNumActiveBodies(EBodyType::RigidBody); }
uint32_t jref_max_bodies(void *h) { return ((World *)h)->system.GetMaxBodies(); }

// Body state by slot (body index) for slots [0, n): arrays may be null. ids[i] = 0xffffffff for empty slots.
void jref_get_state(void *h, uint32_t n, uint32_t *ids, float *pos, float *rot, float *lin, float *ang, float *bounds, uint32_t *active_index, float *sleep_timer)
{
	World *w = (World *)h;
	const BodyVector &bodies = w->system.mBodyManager.GetBodies();
	for (uint32_t i = 0; i < n; ++i)
	{
		const Body *b = i < bodies.size() && BodyManager::sIsValidBodyPointer(bodies[i])? bodies[i] : nullptr;
		if (ids) ids[i] = b? b->GetID().GetIndexAndSequenceNumber() : 0xffffffffu;
		if (b == nullptr) continue;
		if (pos) b2j_adapter::sStore(Vec3(b->GetCenterOfMassPosition()), pos + 3 * i);
		if (rot) b2j_adapter::sStore(b->GetRotation(), rot + 4 * i);
		if (lin) b2j_adapter::sStore(b->IsStatic()? Vec3::sZero() : b->GetLinearVelocity(), lin + 3 * i);
		if (ang) b2j_adapter::sStore(b->IsStatic()? Vec3::sZero() : b->GetAngularVelocity(), ang + 3 * i);
		if (bounds) { b2j_adapter::sStore(b->GetWorldSpaceBounds().mMin, bounds + 6 * i); b2j_adapter::sStore(b->GetWorldSpaceBounds().mMax, bounds + 6 * i + 3); }
		if (active_index) active_index[i] = b->IsStatic()? 0xffffffffu : b->GetMotionPropertiesUnchecked()->GetIndexInActiveBodiesInternal();
		if (sleep_timer) sleep_timer[i] = b->IsStatic()? 0.0f : b->GetMotionPropertiesUnchecked()->mSleepTestTimer;
	}
}

// Candidate pairs of the CURRENT state through the public virtual BroadPhase::FindCollidingPairs (BroadPhase.h:93), as (min id, max id).
uint32_t jref_find_pairs(void *h, uint32_t *outPairs, uint32_t inCap)
{
	World *w = (World *)h;
	BodyIDVector active;
	w->system.GetActiveBodies(EBodyType::RigidBody, active);
	struct Collector : public BodyPairCollector
	{
		void AddHit(const BodyPair &inPair) override { pairs.push_back(inPair); }
		std::vector<BodyPair> pairs;
	} collector;
	if (!active.empty())
	{
		const BroadPhase &bp = static_cast<const BroadPhase &>(w->system.GetBroadPhaseQuery());
		bp.FindCollidingPairs(active.data(), (int)active.size(), w->system.GetPhysicsSettings().mSpeculativeContactDistance, w->ovbp, w->olp, collector);
	}
	std::vector<std::pair<uint32_t, uint32_t>> sorted;
	for (const BodyPair &p : collector.pairs)
	{
		uint32_t a = p.mBodyA.GetIndexAndSequenceNumber(), b = p.mBodyB.GetIndexAndSequenceNumber();
		sorted.emplace_back(std::min(a, b), std::max(a, b));
	}
	std::sort(sorted.begin(), sorted.end());
	for (uint32_t i = 0; i < sorted.size() && i < inCap; ++i) { outPairs[2 * i] = sorted[i].first; outPairs[2 * i + 1] = sorted[i].second; }
	return (uint32_t)sorted.size();
}

// Body pairs / manifolds of the contact cache written by the LAST step (every candidate pair gets an entry,
// ContactConstraintManager.cpp:1090-1128), sorted.
uint32_t jref_get_cache(void *h, b2j_cached_body_pair *outPairs, uint32_t inPairsCap, b2j_cached_manifold *outManifolds, uint32_t inManifoldsCap, uint32_t *outNumManifolds)
{
	World *w = (World *)h;
	std::vector<b2j_cached_body_pair> pairs;
	std::vector<b2j_cached_manifold> manifolds;
	b2j_adapter::sExportContactCache(w->system, pairs, manifolds);
	for (uint32_t i = 0; i < pairs.size() && i < inPairsCap; ++i) outPairs[i] = pairs[i];
	for (uint32_t i = 0; i < manifolds.size() && i < inManifoldsCap; ++i) outManifolds[i] = manifolds[i];
	if (outNumManifolds) *outNumManifolds = (uint32_t)manifolds.size();
	return (uint32_t)pairs.size();
}

// Contact events recorded during the last jref_step, sorted by (kind, body1, body2, sub1, sub2).
uint32_t jref_get_contact_events(void *h, b2j_contact_event *outEvents, uint32_t inCap)
{
	World *w = (World *)h;
	std::vector<ContactRecord> &r = w->contacts.mRecords;
	std::sort(r.begin(), r.end(), [](const ContactRecord &a, const ContactRecord &b) {
		if (a.kind != b.kind) return a.kind < b.kind;
		if (a.body1 != b.body1) return a.body1 < b.body1;
		if (a.body2 != b.body2) return a.body2 < b.body2;
		if (a.sub1 != b.sub1) return a.sub1 < b.sub1;
		return a.sub2 < b.sub2;
	});
	for (uint32_t i = 0; i < r.size() && i < inCap; ++i)
	{
		b2j_contact_event &e = outEvents[i];
		memset(&e, 0, sizeof(e));
		e.kind = r[i].kind; e.body1 = r[i].body1; e.body2 = r[i].body2; e.sub_shape1 = r[i].sub1; e.sub_shape2 = r[i].sub2;
		e.num_points = r[i].num_points;
		memcpy(e.base_offset, r[i].base_offset, sizeof(e.base_offset));
		memcpy(e.normal, r[i].normal, sizeof(e.normal));
		e.penetration_depth = r[i].depth;
		memcpy(e.points1, r[i].p1, sizeof(e.points1));
		memcpy(e.points2, r[i].p2, sizeof(e.points2));
	}
	return (uint32_t)r.size();
}

uint32_t jref_get_activation_events(void *h, b2j_activation_event *outEvents, uint32_t inCap)
{
	World *w = (World *)h;
	std::vector<b2j_activation_event> &r = w->activations.mRecords;
	std::sort(r.begin(), r.end(), [](const b2j_activation_event &a, const b2j_activation_event &b) { return a.kind != b.kind? a.kind < b.kind : a.body < b.body; });
	for (uint32_t i = 0; i < r.size() && i < inCap; ++i) outEvents[i] = r[i];
	return (uint32_t)r.size();
}

uint32_t jref_get_active_bodies(void *h, uint32_t *outIDs, uint32_t inCap)
{
	World *w = (World *)h;
	BodyIDVector active;
	w->system.GetActiveBodies(EBodyType::RigidBody, active);
	for (uint32_t i = 0; i < active.size() && i < inCap; ++i) outIDs[i] = active[i].GetIndexAndSequenceNumber();
	return (uint32_t)active.size();
}

// Re-creates the current state of the reference world inside a new b2j_world (needs jref_bind_b2j first). NULL on failure.
void *jref_export_to_b2j(void *h, int inDevice)
{
	World *w = (World *)h;
	if (sApi.handle == nullptr) { sLastError = "jref_bind_b2j not called"; return nullptr; }
	return b2j_adapter::sExportWorld(sApi, w->system, Layers::NUM_LAYERS, inDevice, w->max_body_pairs, w->max_contact_constraints, sLastError);
}

// Writes the cooked convex hulls and meshes of the world (unique shapes, in body order) to a 'B2JS' file: the cooked assets the
// product's facade loads (cooking is host-side and out of scope). Returns the number of shapes written or -1.
int jref_dump_cooked_shapes(void *h, const char *inPath)
{
	World *w = (World *)h;
	FILE *f = fopen(inPath, "wb");
	if (f == nullptr) return -1;
	std::vector<const Shape *> shapes;
	BodyIDVector ids;
	w->system.GetBodies(ids);
	const BodyLockInterfaceNoLock &li = w->system.GetBodyLockInterfaceNoLock();
	for (BodyID id : ids)
	{
		const Shape *s = li.TryGetBody(id)->GetShape();
		if ((s->GetSubType() == EShapeSubType::ConvexHull || s->GetSubType() == EShapeSubType::Mesh) && std::find(shapes.begin(), shapes.end(), s) == shapes.end())
			shapes.push_back(s);
	}
	auto put32 = [f](uint32_t v) { fwrite(&v, 4, 1, f); };
	auto putf = [f](float v) { fwrite(&v, 4, 1, f); };
	auto putv = [&](Vec3Arg v) { putf(v.GetX()); putf(v.GetY()); putf(v.GetZ()); };
	put32(0x534a3242u);
	put32((uint32_t)shapes.size());
	for (const Shape *s : shapes)
	{
		if (s->GetSubType() == EShapeSubType::ConvexHull)
		{
			const ConvexHullShape *hull = static_cast<const ConvexHullShape *>(s);
			put32(B2J_SHAPE_CONVEX_HULL);
			put32((uint32_t)hull->mPoints.size()); put32((uint32_t)hull->mFaces.size()); put32((uint32_t)hull->mVertexIdx.size());
			putf(hull->mConvexRadius); putv(hull->mCenterOfMass); putv(hull->mLocalBounds.mMin); putv(hull->mLocalBounds.mMax);
			putf(hull->mInnerRadius); putf(hull->mVolume);
			for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) putf(hull->mInertia(r, c));
			for (const ConvexHullShape::Point &p : hull->mPoints) putv(p.mPosition);
			for (const ConvexHullShape::Point &p : hull->mPoints) put32((uint32_t)p.mNumFaces);
			for (const ConvexHullShape::Point &p : hull->mPoints) for (int i = 0; i < 3; ++i) put32((uint32_t)p.mFaces[i]);
			for (const ConvexHullShape::Face &fc : hull->mFaces) fwrite(&fc.mFirstVertex, 2, 1, f);
			for (const ConvexHullShape::Face &fc : hull->mFaces) fwrite(&fc.mNumVertices, 2, 1, f);
			for (const Plane &p : hull->mPlanes) { putv(p.GetNormal()); putf(p.GetConstant()); }
			fwrite(hull->mVertexIdx.data(), 1, hull->mVertexIdx.size(), f);
		}
		else
		{
			const MeshShape *mesh = static_cast<const MeshShape *>(s);
			put32(B2J_SHAPE_MESH);
			put32((uint32_t)mesh->mTree.size());
			AABox b = mesh->GetLocalBounds();
			putv(b.mMin); putv(b.mMax);
			fwrite(mesh->mTree.data(), 1, mesh->mTree.size(), f);
		}
	}
	fclose(f);
	return (int)shapes.size();
}

void jref_get_settings(void *h, b2j_settings *outSettings) { b2j_adapter::sFillSettings(((World *)h)->system.GetPhysicsSettings(), *outSettings); }

// Total kinetic energy 0.5 m v^2 + 0.5 w^T I w over dynamic bodies (long-run comparison, SURVEY 8d parity protocol).
double jref_kinetic_energy(void *h)
{
	World *w = (World *)h;
	double e = 0.0;
	for (const Body *b : w->system.mBodyManager.GetBodies())
		if (BodyManager::sIsValidBodyPointer(b) && b->IsDynamic())
		{
			const MotionProperties *mp = b->GetMotionProperties();
			float inv_m = mp->GetInverseMass();
			if (inv_m > 0.0f)
				e += 0.5 * double(mp->GetLinearVelocity().LengthSq()) / double(inv_m);
			Vec3 wl = (b->GetRotation() * mp->GetInertiaRotation()).InverseRotate(mp->GetAngularVelocity());
			Vec3 d = mp->GetInverseInertiaDiagonal();
			for (int i = 0; i < 3; ++i)
				if (d[i] > 0.0f)
					e += 0.5 * double(wl[i]) * double(wl[i]) / double(d[i]);
		}
	return e;
}

} // extern "C"
