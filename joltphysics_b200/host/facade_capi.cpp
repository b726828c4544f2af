// facade_capi.cpp -- benchmark scenes built through the C++ facade (include/jolt_b200_facade.h) + a tiny C interface so that
// bench.py / tests can drive them. Scene definitions follow PerformanceTest/PyramidScene.h:23-47, ConvexVsMeshScene.h:28-117,
// MaxBodiesScene.h:44-80 and SURVEY.md 8(d) config 4 (Pile); cooked hulls / meshes come from bench_assets/*.b2js
// (reference-cooked, see bench_assets/make_assets.py). Everything goes through PhysicsSystem / BodyInterface exactly as a user
// of the reference would write it; only `namespace JPH = JPH_B200` differs.
#include "jolt_b200_facade.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

namespace JPH = JPH_B200;
using namespace JPH;

// CUDA device of the scenes (one process per GPU: bench.py sets B2J_DEVICE = LOCAL_RANK)


namespace {

// CUDA device of the scenes (one process per GPU: bench.py sets B2J_DEVICE = LOCAL_RANK)
int scene_device() { const char *e = getenv("B2J_DEVICE"); return e != nullptr? atoi(e) : 0; }

// the layer configuration of the reference harness (oracle/ref_harness.cpp): NON_MOVING, MOVING and DEBRIS (a second moving layer in
// its own broadphase tree that collides with the other two but not with itself)
namespace Layers { constexpr ObjectLayer NON_MOVING = 0, MOVING = 1, DEBRIS = 2, NUM_LAYERS = 3; }
namespace BPLayers { constexpr BroadPhaseLayer NON_MOVING(0), MOVING(1), DEBRIS(2); constexpr uint NUM_LAYERS = 3; }

class OLPairFilter final : public ObjectLayerPairFilter
{
public:
	bool ShouldCollide(ObjectLayer a, ObjectLayer b) const override { if (a == Layers::MOVING || b == Layers::MOVING) return true; return a != b; }
};
class BPLInterface final : public BroadPhaseLayerInterface
{
public:
	uint GetNumBroadPhaseLayers() const override { return BPLayers::NUM_LAYERS; }
	BroadPhaseLayer GetBroadPhaseLayer(ObjectLayer l) const override { return l == Layers::NON_MOVING? BPLayers::NON_MOVING : (l == Layers::MOVING? BPLayers::MOVING : BPLayers::DEBRIS); }
};
class OVBPFilter final : public ObjectVsBroadPhaseLayerFilter
{
public:
	bool ShouldCollide(ObjectLayer a, BroadPhaseLayer b) const override { if (a == Layers::MOVING || b == BPLayers::MOVING) return true; return (a == Layers::NON_MOVING) != (b == BPLayers::NON_MOVING); }
};

struct Scene
{
	BPLInterface bpl; OVBPFilter ovbp; OLPairFilter olp;
	PhysicsSystem system;
	std::vector<BodyID> dynamic_bodies;
	std::vector<float> positions, forces;
	std::string error;
};

// ---- cooked shape file (bench_assets/*.b2js) ----
struct Reader
{
	FILE *f;
	bool ok = true;
	template <class T> T get() { T v = T(); if (fread(&v, sizeof(T), 1, f) != 1) ok = false; return v; }
	template <class T> void get(std::vector<T> &v, size_t n) { v.resize(n); if (n > 0 && fread(v.data(), sizeof(T), n, f) != n) ok = false; }
};

bool load_cooked_shapes(const char *path, std::vector<ShapeRef> &out, std::string &error)
{
	FILE *f = fopen(path, "rb");
	if (f == nullptr) { error = std::string("cannot open ") + path; return false; }
	Reader r { f };
	if (r.get<uint32_t>() != 0x534a3242u) { error = "bad magic"; fclose(f); return false; } // 'B2JS'
	uint32_t count = r.get<uint32_t>();
	for (uint32_t i = 0; i < count && r.ok; ++i)
	{
		uint32_t kind = r.get<uint32_t>();
		if (kind == B2J_SHAPE_CONVEX_HULL)
		{
			auto h = std::make_shared<ConvexHullShape>();
			uint32_t np = r.get<uint32_t>(), nf = r.get<uint32_t>(), nv = r.get<uint32_t>();
			h->mConvexRadius = r.get<float>();
			for (int k = 0; k < 3; ++k) h->mCenterOfMass[k] = r.get<float>();
			for (int k = 0; k < 3; ++k) h->mBoundsMin[k] = r.get<float>();
			for (int k = 0; k < 3; ++k) h->mBoundsMax[k] = r.get<float>();
			h->mInnerRadius = r.get<float>();
			h->mVolume = r.get<float>();
			for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) h->mInertia[c][rr] = r.get<float>();
			r.get(h->mPoints, 3 * np); r.get(h->mPointNumFaces, np); r.get(h->mPointFaces, 3 * np);
			r.get(h->mFaceFirstVertex, nf); r.get(h->mFaceNumVertices, nf); r.get(h->mPlanes, 4 * nf); r.get(h->mVertexIdx, nv);
			out.push_back(h);
		}
		else if (kind == B2J_SHAPE_MESH)
		{
			auto m = std::make_shared<MeshShape>();
			uint32_t size = r.get<uint32_t>();
			for (int k = 0; k < 3; ++k) m->mBoundsMin[k] = r.get<float>();
			for (int k = 0; k < 3; ++k) m->mBoundsMax[k] = r.get<float>();
			r.get(m->mTree, size);
			out.push_back(m);
		}
		else { error = "unsupported cooked shape kind"; fclose(f); return false; }
	}
	fclose(f);
	if (!r.ok) error = "truncated cooked shape file";
	return r.ok;
}

Quat random_quat(std::mt19937 &rnd)
{
	std::normal_distribution<float> n(0.0f, 1.0f);
	float x = n(rnd), y = n(rnd), z = n(rnd), w = n(rnd);
	return Quat(x, y, z, w).Normalized();
}

bool scene_pyramid(Scene &s, int height, int tight_limit = 0)
{
	// tight_limit > 0: max body pairs = max contact constraints = that value (error path tests)
	if (!s.system.Init(10240, 0, tight_limit > 0? (uint)tight_limit : 65536, tight_limit > 0? (uint)tight_limit : 20480, s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	BodyInterface &bi = s.system.GetBodyInterface();
	bi.CreateAndAddBody(BodyCreationSettings(std::make_shared<BoxShape>(Vec3(50.0f, 1.0f, 50.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING), EActivation::DontActivate);
	const float box_size = 2.0f, separation = 0.5f, half = 1.0f;
	ShapeRef box = std::make_shared<BoxShape>(Vec3::sReplicate(half), 0.0f);
	for (int i = 0; i < height; ++i)
		for (int j = i / 2; j < height - (i + 1) / 2; ++j)
			for (int k = i / 2; k < height - (i + 1) / 2; ++k)
			{
				RVec3 pos(float(-height) + box_size * j + ((i & 1)? half : 0.0f), 1.0f + (box_size + separation) * i, float(-height) + box_size * k + ((i & 1)? half : 0.0f));
				BodyCreationSettings bs(box, pos, Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
				bs.mAllowSleeping = false;
				s.dynamic_bodies.push_back(bi.CreateAndAddBody(bs, EActivation::Activate));
			}
	return true;
}

bool scene_convex_vs_mesh(Scene &s, int half_grid, const char *assets_dir)
{
	std::vector<ShapeRef> cooked;
	if (!load_cooked_shapes((std::string(assets_dir) + "/convex_vs_mesh.b2js").c_str(), cooked, s.error) || cooked.size() < 2) return false;
	if (!s.system.Init(10240, 0, 65536, 20480, s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	PhysicsSettings settings = s.system.GetPhysicsSettings();
	settings.mNumVelocitySteps = 4;
	settings.mNumPositionSteps = 1;
	s.system.SetPhysicsSettings(settings);
	const int n = 100;
	const float cell_size = 3.0f, max_height = 5.0f, center = n * cell_size / 2;
	BodyInterface &bi = s.system.GetBodyInterface();
	BodyCreationSettings mesh(cooked[0], RVec3(-center, max_height, -center), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
	mesh.mFriction = 0.5f; mesh.mRestitution = 0.6f;
	bi.CreateAndAddBody(mesh, EActivation::DontActivate);
	ShapeRef shapes[4] = { std::make_shared<BoxShape>(Vec3(0.5f, 0.75f, 1.0f)), std::make_shared<SphereShape>(0.5f), std::make_shared<CapsuleShape>(0.75f, 0.5f), cooked[1] };
	for (int x = -half_grid; x <= half_grid; ++x)
		for (int y = 0; y < 4; ++y)
			for (int z = -half_grid; z <= half_grid; ++z)
			{
				BodyCreationSettings bs;
				bs.mMotionType = EMotionType::Dynamic;
				bs.mObjectLayer = Layers::MOVING;
				bs.mPosition = RVec3(7.5f * x, 15.0f + 2.0f * y, 7.5f * z);
				bs.mFriction = 0.5f; bs.mRestitution = 0.6f;
				bs.SetShape(shapes[y]);
				s.dynamic_bodies.push_back(bi.CreateAndAddBody(bs, EActivation::Activate));
			}
	return true;
}

bool scene_pile(Scene &s, int num_bodies, int shape_mask, const char *assets_dir)
{
	std::vector<ShapeRef> hulls;
	if ((shape_mask & 8) && !load_cooked_shapes((std::string(assets_dir) + "/pile_hulls.b2js").c_str(), hulls, s.error)) return false;
	uint n = (uint)num_bodies;
	uint side = (uint)std::ceil(std::cbrt(double(n)));
	float spacing = 1.15f;
	float half_width = 0.5f * side * spacing + 2.0f;
	if (!s.system.Init(n + 128, 0, std::max(65536u, 16 * n), std::max(20480u, 8 * n), s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	BodyInterface &bi = s.system.GetBodyInterface();
	auto add_static = [&](Vec3 he, Vec3 pos) {
		BodyCreationSettings bs(std::make_shared<BoxShape>(he), pos, Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		bs.mFriction = 0.5f; bs.mRestitution = 0.1f;
		bi.CreateAndAddBody(bs, EActivation::DontActivate);
	};
	float wall_h = 0.5f * side * spacing + 2.0f;
	add_static(Vec3(half_width + 2.0f, 1.0f, half_width + 2.0f), Vec3(0, -1.0f, 0));
	add_static(Vec3(1.0f, wall_h, half_width + 2.0f), Vec3(-half_width - 1.0f, wall_h, 0));
	add_static(Vec3(1.0f, wall_h, half_width + 2.0f), Vec3(half_width + 1.0f, wall_h, 0));
	add_static(Vec3(half_width + 2.0f, wall_h, 1.0f), Vec3(0, wall_h, -half_width - 1.0f));
	add_static(Vec3(half_width + 2.0f, wall_h, 1.0f), Vec3(0, wall_h, half_width + 1.0f));

	std::mt19937 rnd(12345);
	std::uniform_real_distribution<float> jitter(-0.05f, 0.05f);
	ShapeRef sphere = std::make_shared<SphereShape>(0.5f), box = std::make_shared<BoxShape>(Vec3(0.5f, 0.4f, 0.3f), 0.05f), capsule = std::make_shared<CapsuleShape>(0.4f, 0.3f);
	std::vector<int> kinds;
	for (int k = 0; k < 4; ++k) if (shape_mask & (1 << k)) kinds.push_back(k);
	if (kinds.empty()) kinds = { 0, 1, 2 };
	std::vector<BodyID> ids;
	ids.reserve(n);
	uint count = 0, hull_idx = 0;
	for (uint y = 0; y < side && count < n; ++y)
		for (uint x = 0; x < side && count < n; ++x)
			for (uint z = 0; z < side && count < n; ++z, ++count)
			{
				BodyCreationSettings bs;
				switch (kinds[count % kinds.size()])
				{
				case 0: bs.SetShape(sphere); break;
				case 1: bs.SetShape(box); break;
				case 2: bs.SetShape(capsule); break;
				default: bs.SetShape(hulls[hull_idx++ % hulls.size()]); break; // palette of reference-cooked 12 point hulls
				}
				bs.mMotionType = EMotionType::Dynamic;
				bs.mObjectLayer = Layers::MOVING;
				float jx = jitter(rnd), jy = jitter(rnd), jz = jitter(rnd);
				bs.mPosition = RVec3((float(x) - 0.5f * (side - 1)) * spacing + jx, 0.8f + float(y) * spacing + jy, (float(z) - 0.5f * (side - 1)) * spacing + jz);
				bs.mRotation = random_quat(rnd);
				bs.mFriction = 0.5f; bs.mRestitution = 0.1f;
				Body *b = bi.CreateBody(bs);
				if (b == nullptr) { s.error = "out of bodies"; return false; }
				ids.push_back(b->GetID());
			}
	bi.AddBodies(ids.data(), (int)ids.size(), EActivation::Activate);
	s.dynamic_bodies = ids;
	return true;
}

bool scene_max_bodies(Scene &s, int num_bodies)
{
	uint n = (uint)num_bodies;
	if (!s.system.Init(n, 0, std::max(65536u, n), std::max(20480u, n), s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	PhysicsSettings settings = s.system.GetPhysicsSettings();
	settings.mNumVelocitySteps = 4;
	settings.mNumPositionSteps = 1;
	s.system.SetPhysicsSettings(settings);
	BodyInterface &bi = s.system.GetBodyInterface();
	uint side = (uint)std::ceil(std::cbrt(double(n)));
	BodyCreationSettings bs;
	bs.SetShape(std::make_shared<BoxShape>(Vec3::sReplicate(0.5f)));
	bs.mMotionType = EMotionType::Dynamic;
	bs.mObjectLayer = Layers::MOVING;
	bs.mOverrideMassProperties = EOverrideMassProperties::CalculateInertia;
	bs.mMassPropertiesOverride.mMass = 1000.0f;
	std::vector<BodyID> ids;
	uint count = 0;
	for (uint x = 0; x < side && count < n; ++x)
		for (uint y = 0; y < side && count < n; ++y)
			for (uint z = 0; z < side && count < n; ++z, ++count)
			{
				bs.mPosition = RVec3(1.0f * x, 3.0f * y, 3.0f * z);
				ids.push_back(bi.CreateBody(bs)->GetID());
			}
	bi.AddBodies(ids.data(), (int)ids.size(), EActivation::Activate);
	s.dynamic_bodies = ids;
	return true;
}

#define B2J_SHAPE_REF ShapeRef
#define B2J_NEW_SHAPE(Type, ...) std::make_shared<Type>(__VA_ARGS__)
#include "api_tour.inl"

static Quat sRandomQuat(std::mt19937 &r) { return random_quat(r); }
#include "feature_scenes.inl"

// the feature scenes of feature_scenes.inl (same user code as the reference harness compiles); the hull is the first cooked hull of
// bench_assets/pile_hulls.b2js
bool scene_feature(Scene &s, int variant, const char *assets_dir)
{
	std::vector<ShapeRef> hulls;
	if (!load_cooked_shapes((std::string(assets_dir) + "/pile_hulls.b2js").c_str(), hulls, s.error) || hulls.empty()) return false;
	if (!s.system.Init(1024, 0, 8192, 4096, s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	uint32_t num_dynamic = 0;
	sFeatureCreate(s.system, variant, hulls[0], num_dynamic);
	// (dynamic_bodies = every non static body in creation order, for the e2e getters)
	BodyIDVector all;
	s.system.GetBodies(all);
	for (const BodyID &id : all) if (s.system.GetBodyInterface().GetMotionType(id) != EMotionType::Static) s.dynamic_bodies.push_back(id);
	return s.dynamic_bodies.size() == num_dynamic;
}

#define B2J_CREATE_COMPOUND(settings) (settings).Create()
#include "compound_scene.inl"

// the compound scene of compound_scene.inl on a box floor (variant 0; the terrain variant needs a mesh the reference cooks)
bool scene_compound(Scene &s, const char *assets_dir)
{
	std::vector<ShapeRef> hulls;
	if (!load_cooked_shapes((std::string(assets_dir) + "/pile_hulls.b2js").c_str(), hulls, s.error) || hulls.empty()) return false;
	if (!s.system.Init(1024, 0, 16384, 8192, s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	uint32_t num_dynamic = 0;
	sCompoundCreate(s.system, 0, hulls[0], num_dynamic);
	BodyIDVector all;
	s.system.GetBodies(all);
	for (const BodyID &id : all) if (s.system.GetBodyInterface().GetMotionType(id) != EMotionType::Static) s.dynamic_bodies.push_back(id);
	return s.dynamic_bodies.size() == num_dynamic;
}

bool scene_api_tour(Scene &s)
{
	if (!s.system.Init(1024, 0, 4096, 1024, s.bpl, s.ovbp, s.olp, Layers::NUM_LAYERS, scene_device())) return false;
	sApiTourCreate(s.system, s.dynamic_bodies);
	return true;
}

thread_local std::string g_error;

} // namespace

#define B2JF_API extern "C" __attribute__((visibility("default")))

B2JF_API const char *b2jf_last_error() { return g_error.c_str(); }

// Builds a scene through the facade. Returns a handle or NULL (see b2jf_last_error).
B2JF_API void *b2jf_scene_create(const char *name, int p0, int p1, const char *assets_dir)
{
	Scene *s = new Scene;
	std::string n(name);
	bool ok = false;
	if (n == "pyramid") ok = scene_pyramid(*s, p0 > 0? p0 : 15);
	else if (n == "pyramid_tight") ok = scene_pyramid(*s, p0 > 0? p0 : 6, p1 > 0? p1 : 64);
	else if (n == "convex_vs_mesh") ok = scene_convex_vs_mesh(*s, p0 > 0? p0 : 10, assets_dir);
	else if (n == "pile") ok = scene_pile(*s, p0 > 0? p0 : 1000, p1 > 0? p1 : 15, assets_dir);
	else if (n == "max_bodies") ok = scene_max_bodies(*s, p0 > 0? p0 : 10000);
	else if (n == "api_tour") ok = scene_api_tour(*s);
	else if (n == "feature") ok = scene_feature(*s, p0, assets_dir);
	else if (n == "compound") ok = scene_compound(*s, assets_dir);
	else s->error = "unknown scene";
	if (!ok)
	{
		g_error = s->error.empty()? std::string("scene creation failed: ") + s->system.GetLastError() : s->error;
		delete s;
		return nullptr;
	}
	return s;
}

B2JF_API void b2jf_scene_destroy(void *h) { delete (Scene *)h; }
B2JF_API void *b2jf_scene_world(void *h) { return ((Scene *)h)->system.GetWorld(); }
B2JF_API uint32_t b2jf_scene_num_dynamic(void *h) { return (uint32_t)((Scene *)h)->dynamic_bodies.size(); }
B2JF_API uint32_t b2jf_scene_num_bodies(void *h) { return ((Scene *)h)->system.GetNumBodies(); }
B2JF_API void b2jf_scene_flush(void *h) { ((Scene *)h)->system.GetBodyInterface().AddForcesAndTorques(nullptr, 0, nullptr, nullptr); }

// api_tour scene: the mutation phases and the queries of api_tour.inl
B2JF_API void b2jf_scene_mutate(void *h, int phase) { Scene *s = (Scene *)h; sApiTourMutate(s->system, s->dynamic_bodies, phase); }
B2JF_API int b2jf_scene_query(void *h, uint32_t *out_ids, int cap, uint32_t *out_num_bodies, uint32_t *out_flags) { Scene *s = (Scene *)h; return sApiTourQuery(s->system, s->dynamic_bodies, out_ids, cap, out_num_bodies, out_flags); }

// NarrowPhaseQuery::CastRays / BroadPhaseQuery::CollideAABox through the facade (rays: [n][6] floats; hits: body, sub shape, fraction)
B2JF_API void b2jf_scene_cast_rays(void *h, const float *rays, int n, uint32_t *out_body, uint32_t *out_sub, float *out_fraction)
{
	Scene *s = (Scene *)h;
	std::vector<RayCastResult> hits((size_t)n);
	s->system.GetNarrowPhaseQuery().CastRays(reinterpret_cast<const RRayCast *>(rays), n, hits.data());
	for (int i = 0; i < n; ++i) { out_body[i] = hits[i].mBodyID.GetIndexAndSequenceNumber(); out_sub[i] = hits[i].mSubShapeID2.GetValue(); out_fraction[i] = hits[i].mFraction; }
	// the single ray form agrees with the batch
	if (n > 0)
	{
		RayCastResult one;
		bool hit = s->system.GetNarrowPhaseQuery().CastRay(reinterpret_cast<const RRayCast *>(rays)[0], one);
		if (hit != !hits[0].mBodyID.IsInvalid() || (hit && one.mFraction != hits[0].mFraction)) out_body[0] = 0xdeadbeefu;
	}
}
B2JF_API int b2jf_scene_collide_aabox(void *h, const float *box, uint32_t *out_ids, int cap)
{
	Scene *s = (Scene *)h;
	std::vector<BodyID> ids;
	s->system.GetBroadPhaseQuery().CollideAABox(AABox(Vec3(box[0], box[1], box[2]), Vec3(box[3], box[4], box[5])), ids);
	for (size_t i = 0; i < ids.size() && (int)i < cap; ++i) out_ids[i] = ids[i].GetIndexAndSequenceNumber();
	return (int)ids.size();
}

// NarrowPhaseQuery::CollideShape (a box of the given half extent, scaled) and BroadPhaseQuery::CollideSphere / CollidePoint through the facade.
// out: per hit body, sub shape 2, depth; returns the number of hits. The matrix form and the batched form must agree with the single one.
B2JF_API int b2jf_scene_collide_shape(void *h, const float *half_extent, const float *scale, const float *rotation, const float *position, float max_separation,
	uint32_t *out_body, uint32_t *out_sub2, float *out_depth, int cap)
{
	Scene *s = (Scene *)h;
	static std::shared_ptr<const BoxShape> box; // (kept alive by the caller, as the reference asks)
	static float box_he[3];
	if (box == nullptr || box_he[0] != half_extent[0] || box_he[1] != half_extent[1] || box_he[2] != half_extent[2])
	{
		box = std::make_shared<BoxShape>(Vec3(half_extent[0], half_extent[1], half_extent[2]), 0.05f);
		box_he[0] = half_extent[0]; box_he[1] = half_extent[1]; box_he[2] = half_extent[2];
	}
	Quat q(rotation[0], rotation[1], rotation[2], rotation[3]);
	RVec3 p(position[0], position[1], position[2]);
	Vec3 sc(scale[0], scale[1], scale[2]);
	CollideShapeSettings settings; settings.mMaxSeparationDistance = max_separation;
	std::vector<CollideShapeResult> hits, hits_m;
	const NarrowPhaseQuery &query = s->system.GetNarrowPhaseQuery();
	query.CollideShape(box.get(), sc, q, p, settings, p, hits);
	query.CollideShape(box.get(), sc, RMat44::sRotationTranslation(q, p), settings, p, hits_m);
	std::vector<std::vector<CollideShapeResult>> batch;
	query.CollideShapes(box.get(), sc, &q, &p, 1, settings, batch);
	bool ok = hits_m.size() == hits.size() && batch.size() == 1 && batch[0].size() == hits.size();
	for (size_t i = 0; ok && i < hits.size(); ++i)
		ok = hits_m[i].mBodyID2 == hits[i].mBodyID2 && batch[0][i].mBodyID2 == hits[i].mBodyID2 && std::fabs(hits_m[i].mPenetrationDepth - hits[i].mPenetrationDepth) < 1.0e-4f;
	if (!ok) return -1;
	for (size_t i = 0; i < hits.size() && (int)i < cap; ++i)
	{
		out_body[i] = hits[i].mBodyID2.GetIndexAndSequenceNumber(); out_sub2[i] = hits[i].mSubShapeID2.GetValue(); out_depth[i] = hits[i].mPenetrationDepth;
	}
	return (int)hits.size();
}
B2JF_API int b2jf_scene_collide_sphere(void *h, const float *sphere, uint32_t *out_ids, int cap)
{
	Scene *s = (Scene *)h;
	std::vector<BodyID> ids;
	if (sphere[3] < 0.0f) s->system.GetBroadPhaseQuery().CollidePoint(Vec3(sphere[0], sphere[1], sphere[2]), ids);
	else s->system.GetBroadPhaseQuery().CollideSphere(Vec3(sphere[0], sphere[1], sphere[2]), sphere[3], ids);
	for (size_t i = 0; i < ids.size() && (int)i < cap; ++i) out_ids[i] = ids[i].GetIndexAndSequenceNumber();
	return (int)ids.size();
}

// rows of body state the last refresh of the host mirror fetched (the incremental download: only what the step simulated)
B2JF_API uint32_t b2jf_scene_last_download_count(void *h) { return ((Scene *)h)->system.GetLastDownloadCount(); }

// PhysicsSystem::Update through the facade (mirrors the state to the host, replays events). Returns the error bits.
B2JF_API int b2jf_scene_update(void *h, float dt, int collision_steps, b2j_step_stats *out_stats)
{
	Scene *s = (Scene *)h;
	int r = (int)s->system.Update(dt, collision_steps, nullptr, nullptr);
	if (out_stats) *out_stats = s->system.GetLastStepStats();
	return r;
}

// End to end step with HOST buffers (the RL pattern): forces [num_dynamic][3] are applied to the dynamic bodies (H2D), the step runs,
// the positions of the dynamic bodies are written to out_positions [num_dynamic][3] (D2H through the facade's host mirror).
B2JF_API int b2jf_scene_step_e2e(void *h, float dt, const float *forces, float *out_positions)
{
	Scene *s = (Scene *)h;
	BodyInterface &bi = s->system.GetBodyInterface();
	int n = (int)s->dynamic_bodies.size();
	if (forces != nullptr)
		bi.AddForcesAndTorques(s->dynamic_bodies.data(), n, forces, nullptr);
	int r = (int)s->system.Update(dt, 1, nullptr, nullptr);
	if (out_positions != nullptr)
		for (int i = 0; i < n; ++i)
		{
			RVec3 p = bi.GetCenterOfMassPosition(s->dynamic_bodies[i]);
			out_positions[3 * i] = p.x; out_positions[3 * i + 1] = p.y; out_positions[3 * i + 2] = p.z;
		}
	return r;
}
