// jolt_adapter.h -- TEST INFRASTRUCTURE (oracle side). Walks a live reference JPH::PhysicsSystem and re-creates it
// inside a b2j_world through the C ABI of include/jolt_b200.h. This is also the binding a reference maintainer would add
// (see INTEGRATION.md): everything the step depends on crosses the boundary as plain arrays.
//
// Compiled only into oracle/_ref/libjoltref_*.so with -fno-access-control (to read cooked hull / mesh / contact cache
// members that the reference keeps private). The product never includes this file.
#pragma once

#include <Jolt/Jolt.h>
#include <Jolt/Physics/PhysicsSystem.h>
#include <Jolt/Physics/Body/BodyCreationSettings.h>
#include <Jolt/Physics/Collision/Shape/SphereShape.h>
#include <Jolt/Physics/Collision/Shape/BoxShape.h>
#include <Jolt/Physics/Collision/Shape/CapsuleShape.h>
#include <Jolt/Physics/Collision/Shape/CylinderShape.h>
#include <Jolt/Physics/Collision/Shape/ConvexHullShape.h>
#include <Jolt/Physics/Collision/Shape/MeshShape.h>
#include <Jolt/Physics/Collision/Shape/StaticCompoundShape.h>
#include <Jolt/Physics/Collision/Shape/ScaledShape.h>
#include <Jolt/Physics/Collision/Shape/RotatedTranslatedShape.h>

#include <Jolt/Physics/Constraints/PointConstraint.h>
#include <Jolt/Physics/Constraints/DistanceConstraint.h>
#include <Jolt/Physics/Constraints/HingeConstraint.h>
#include <Jolt/Physics/Constraints/FixedConstraint.h>
#include <jolt_b200.h>

#include <unordered_map>
#include <vector>
#include <cstring>
#include <dlfcn.h>

namespace b2j_adapter {

using namespace JPH;

// Function table resolved from libjolt_b200.so (or any library exporting the same C ABI) at run time.
struct Api
{
	void *handle = nullptr;
#define B2J_FN(name) decltype(&::name) name = nullptr;
	B2J_FN(b2j_world_create) B2J_FN(b2j_world_destroy) B2J_FN(b2j_last_error) B2J_FN(b2j_settings_default)
	B2J_FN(b2j_world_set_previous_delta_time)
	B2J_FN(b2j_shape_sphere) B2J_FN(b2j_shape_box) B2J_FN(b2j_shape_capsule) B2J_FN(b2j_shape_cylinder) B2J_FN(b2j_shape_convex_hull) B2J_FN(b2j_shape_mesh) B2J_FN(b2j_shape_scaled) B2J_FN(b2j_shape_rotated_translated) B2J_FN(b2j_shape_static_compound)
	B2J_FN(b2j_bodies_add) B2J_FN(b2j_set_active_list) B2J_FN(b2j_contact_cache_import)
	B2J_FN(b2j_constraints_add) B2J_FN(b2j_constraints_set_state)
#undef B2J_FN

	bool Load(const char *inPath, String &outError)
	{
		handle = dlopen(inPath, RTLD_NOW | RTLD_LOCAL);
		if (handle == nullptr) { outError = dlerror(); return false; }
#define B2J_FN(name) name = (decltype(name))dlsym(handle, #name); if (name == nullptr) { outError = String("missing symbol ") + #name; return false; }
		B2J_FN(b2j_world_create) B2J_FN(b2j_world_destroy) B2J_FN(b2j_last_error) B2J_FN(b2j_settings_default)
		B2J_FN(b2j_world_set_previous_delta_time)
		B2J_FN(b2j_shape_sphere) B2J_FN(b2j_shape_box) B2J_FN(b2j_shape_capsule) B2J_FN(b2j_shape_cylinder) B2J_FN(b2j_shape_convex_hull) B2J_FN(b2j_shape_mesh) B2J_FN(b2j_shape_scaled) B2J_FN(b2j_shape_rotated_translated) B2J_FN(b2j_shape_static_compound)
		B2J_FN(b2j_bodies_add) B2J_FN(b2j_set_active_list) B2J_FN(b2j_contact_cache_import)
		B2J_FN(b2j_constraints_add) B2J_FN(b2j_constraints_set_state)
#undef B2J_FN
		return true;
	}
};

inline void sStore(Vec3Arg inV, float *outV) { outV[0] = inV.GetX(); outV[1] = inV.GetY(); outV[2] = inV.GetZ(); }
inline void sStore(QuatArg inQ, float *outV) { outV[0] = inQ.GetX(); outV[1] = inQ.GetY(); outV[2] = inQ.GetZ(); outV[3] = inQ.GetW(); }
inline void sStore(const Float3 &inV, float *outV) { outV[0] = inV.x; outV[1] = inV.y; outV[2] = inV.z; }

inline void sFillSettings(const PhysicsSettings &inS, b2j_settings &outS)
{
	memset(&outS, 0, sizeof(outS));
	outS.speculative_contact_distance = inS.mSpeculativeContactDistance;
	outS.penetration_slop = inS.mPenetrationSlop;
	outS.baumgarte = inS.mBaumgarte;
	outS.max_penetration_distance = inS.mMaxPenetrationDistance;
	outS.manifold_tolerance = inS.mManifoldTolerance;
	outS.body_pair_cache_max_delta_position_sq = inS.mBodyPairCacheMaxDeltaPositionSq;
	outS.body_pair_cache_cos_max_delta_rotation_div2 = inS.mBodyPairCacheCosMaxDeltaRotationDiv2;
	outS.contact_normal_cos_max_delta_rotation = inS.mContactNormalCosMaxDeltaRotation;
	outS.contact_point_preserve_lambda_max_dist_sq = inS.mContactPointPreserveLambdaMaxDistSq;
	outS.min_velocity_for_restitution = inS.mMinVelocityForRestitution;
	outS.time_before_sleep = inS.mTimeBeforeSleep;
	outS.point_velocity_sleep_threshold = inS.mPointVelocitySleepThreshold;
	outS.num_velocity_steps = inS.mNumVelocitySteps;
	outS.num_position_steps = inS.mNumPositionSteps;
	outS.deterministic_simulation = inS.mDeterministicSimulation;
	outS.constraint_warm_start = inS.mConstraintWarmStart;
	outS.use_body_pair_contact_cache = inS.mUseBodyPairContactCache;
	outS.use_manifold_reduction = inS.mUseManifoldReduction;
	outS.use_large_island_splitter = inS.mUseLargeIslandSplitter;
	outS.allow_sleeping = inS.mAllowSleeping;
	outS.check_active_edges = inS.mCheckActiveEdges;
}

// Upload one reference shape, returns the b2j shape id or < 0
inline int32_t sUploadShape(const Api &inApi, b2j_world *inWorld, const Shape *inShape, String &outError)
{
	switch (inShape->GetSubType())
	{
	case EShapeSubType::Sphere:
		return inApi.b2j_shape_sphere(inWorld, static_cast<const SphereShape *>(inShape)->GetRadius());

	case EShapeSubType::Box:
		{
			const BoxShape *box = static_cast<const BoxShape *>(inShape);
			float he[3];
			sStore(box->GetHalfExtent(), he);
			return inApi.b2j_shape_box(inWorld, he, box->GetConvexRadius());
		}

	case EShapeSubType::Capsule:
		{
			const CapsuleShape *capsule = static_cast<const CapsuleShape *>(inShape);
			return inApi.b2j_shape_capsule(inWorld, capsule->GetHalfHeightOfCylinder(), capsule->GetRadius());
		}

	case EShapeSubType::Cylinder:
		{
			const CylinderShape *cylinder = static_cast<const CylinderShape *>(inShape);
			return inApi.b2j_shape_cylinder(inWorld, cylinder->GetHalfHeight(), cylinder->GetRadius(), cylinder->GetConvexRadius());
		}

	case EShapeSubType::ConvexHull:
		{
			const ConvexHullShape *hull = static_cast<const ConvexHullShape *>(inShape);
			std::vector<float> points, planes;
			std::vector<int32_t> num_faces, faces;
			std::vector<uint16_t> first_vertex, num_vertices;
			for (const ConvexHullShape::Point &p : hull->mPoints)
			{
				points.push_back(p.mPosition.GetX()); points.push_back(p.mPosition.GetY()); points.push_back(p.mPosition.GetZ());
				num_faces.push_back(p.mNumFaces);
				for (int i = 0; i < 3; ++i) faces.push_back(p.mFaces[i]);
			}
			for (const ConvexHullShape::Face &f : hull->mFaces)
			{
				first_vertex.push_back(f.mFirstVertex);
				num_vertices.push_back(f.mNumVertices);
			}
			for (const Plane &p : hull->mPlanes)
			{
				planes.push_back(p.GetNormal().GetX()); planes.push_back(p.GetNormal().GetY()); planes.push_back(p.GetNormal().GetZ());
				planes.push_back(p.GetConstant());
			}
			b2j_hull_desc desc;
			memset(&desc, 0, sizeof(desc));
			desc.num_points = (uint32_t)hull->mPoints.size();
			desc.points = points.data();
			desc.point_num_faces = num_faces.data();
			desc.point_faces = faces.data();
			desc.num_faces = (uint32_t)hull->mFaces.size();
			desc.face_first_vertex = first_vertex.data();
			desc.face_num_vertices = num_vertices.data();
			desc.planes = planes.data();
			desc.num_vertex_idx = (uint32_t)hull->mVertexIdx.size();
			desc.vertex_idx = hull->mVertexIdx.data();
			desc.convex_radius = hull->mConvexRadius;
			sStore(hull->mCenterOfMass, desc.center_of_mass);
			sStore(hull->mLocalBounds.mMin, desc.local_bounds_min);
			sStore(hull->mLocalBounds.mMax, desc.local_bounds_max);
			desc.inner_radius = hull->mInnerRadius;
			return inApi.b2j_shape_convex_hull(inWorld, &desc);
		}

	case EShapeSubType::Mesh:
		{
			const MeshShape *mesh = static_cast<const MeshShape *>(inShape);
			b2j_mesh_desc desc;
			memset(&desc, 0, sizeof(desc));
			desc.tree = mesh->mTree.data();
			desc.tree_size = (uint32_t)mesh->mTree.size();
			AABox bounds = mesh->GetLocalBounds();
			sStore(bounds.mMin, desc.local_bounds_min);
			sStore(bounds.mMax, desc.local_bounds_max);
			return inApi.b2j_shape_mesh(inWorld, &desc);
		}

	case EShapeSubType::StaticCompound:
		{
			// sub shapes first, then the compound with the tree exactly as the reference built it (mNodes: private, the harness is a friend
			// by way of its access define)
			const StaticCompoundShape *compound = static_cast<const StaticCompoundShape *>(inShape);
			Array<b2j_compound_sub> subs;
			for (const CompoundShape::SubShape &sub : compound->GetSubShapes())
			{
				b2j_compound_sub out;
				out.shape = sUploadShape(inApi, inWorld, sub.mShape, outError);
				if (out.shape < 0) return -1;
				sStore(sub.GetPositionCOM(), out.position_com);
				sStore(sub.GetRotation(), out.rotation);
				subs.push_back(out);
			}
			b2j_compound_desc desc;
			memset(&desc, 0, sizeof(desc));
			desc.num_subs = (uint32_t)subs.size(); desc.subs = subs.data();
			desc.num_nodes = (uint32_t)compound->mNodes.size(); desc.nodes = reinterpret_cast<const uint8_t *>(compound->mNodes.data());
			sStore(compound->GetCenterOfMass(), desc.center_of_mass);
			AABox bounds = compound->GetLocalBounds();
			sStore(bounds.mMin, desc.local_bounds_min); sStore(bounds.mMax, desc.local_bounds_max);
			desc.inner_radius = compound->GetInnerRadius();
			int32_t id = inApi.b2j_shape_static_compound(inWorld, &desc);
			if (id < 0) outError = inApi.b2j_last_error();
			return id;
		}

	case EShapeSubType::Scaled:
		{
			// ScaledShape around a convex shape: the inner shape first, then the decoration
			const ScaledShape *scaled = static_cast<const ScaledShape *>(inShape);
			int32_t inner = sUploadShape(inApi, inWorld, scaled->GetInnerShape(), outError);
			if (inner < 0) return -1;
			float scale[3];
			sStore(scaled->GetScale(), scale);
			int32_t id = inApi.b2j_shape_scaled(inWorld, inner, scale);
			if (id < 0) outError = inApi.b2j_last_error();
			return id;
		}

	case EShapeSubType::RotatedTranslated:
		{
			const RotatedTranslatedShape *rt = static_cast<const RotatedTranslatedShape *>(inShape);
			int32_t inner = sUploadShape(inApi, inWorld, rt->GetInnerShape(), outError);
			if (inner < 0) return -1;
			float rotation[4], center_of_mass[3];
			sStore(rt->GetRotation(), rotation);
			sStore(rt->GetCenterOfMass(), center_of_mass);
			int32_t id = inApi.b2j_shape_rotated_translated(inWorld, inner, rotation, center_of_mass);
			if (id < 0) outError = inApi.b2j_last_error();
			return id;
		}

	default:
		outError = "unsupported shape sub type";
		return -1;
	}
}

// Fill one body descriptor from a reference Body (everything of SURVEY A.4)
inline void sFillBody(const Body &inBody, int32_t inShapeID, b2j_body_desc &outDesc)
{
	memset(&outDesc, 0, sizeof(outDesc));
	outDesc.id = inBody.GetID().GetIndexAndSequenceNumber();
	outDesc.shape = inShapeID;
	outDesc.motion_type = (uint8_t)inBody.GetMotionType();
	outDesc.object_layer = (uint16_t)inBody.GetObjectLayer();
	uint16_t flags = 0;
	if (inBody.IsSensor()) flags |= B2J_BODY_SENSOR;
	if (inBody.GetUseManifoldReduction()) flags |= B2J_BODY_USE_MANIFOLD_REDUCTION;
	if (inBody.GetApplyGyroscopicForce()) flags |= B2J_BODY_GYROSCOPIC;
	if (inBody.GetCollideKinematicVsNonDynamic()) flags |= B2J_BODY_KIN_VS_NONDYN;
	if (inBody.IsCollisionCacheInvalid()) flags |= B2J_BODY_INVALIDATE_CACHE;
	sStore(Vec3(inBody.GetCenterOfMassPosition()), outDesc.position);
	sStore(inBody.GetRotation(), outDesc.rotation);
	outDesc.friction = inBody.GetFriction();
	outDesc.restitution = inBody.GetRestitution();
	sStore(inBody.GetWorldSpaceBounds().mMin, outDesc.bounds_min);
	sStore(inBody.GetWorldSpaceBounds().mMax, outDesc.bounds_max);
	outDesc.has_bounds = 1;
	outDesc.allowed_dofs = 0x3f;
	outDesc.rotation[3] = inBody.GetRotation().GetW();
	outDesc.inertia_rotation[3] = 1.0f;
	outDesc.gravity_factor = 1.0f;
	if (!inBody.IsStatic())
	{
		const MotionProperties *mp = inBody.GetMotionPropertiesUnchecked();
		if (mp->mAllowSleeping) flags |= B2J_BODY_ALLOW_SLEEPING;
		outDesc.allowed_dofs = (uint8_t)mp->mAllowedDOFs;
		outDesc.num_velocity_steps_override = mp->mNumVelocityStepsOverride;
		outDesc.num_position_steps_override = mp->mNumPositionStepsOverride;
		sStore(mp->mLinearVelocity, outDesc.linear_velocity);
		sStore(mp->mAngularVelocity, outDesc.angular_velocity);
		sStore(mp->mForce, outDesc.force);
		sStore(mp->mTorque, outDesc.torque);
		outDesc.inv_mass = inBody.IsDynamic()? mp->mInvMass : 0.0f;
		sStore(mp->mInvInertiaDiagonal, outDesc.inv_inertia_diag);
		sStore(mp->mInertiaRotation, outDesc.inertia_rotation);
		outDesc.linear_damping = mp->mLinearDamping;
		outDesc.angular_damping = mp->mAngularDamping;
		outDesc.max_linear_velocity = mp->mMaxLinearVelocity;
		outDesc.max_angular_velocity = mp->mMaxAngularVelocity;
		outDesc.gravity_factor = mp->mGravityFactor;
		for (int s = 0; s < 3; ++s)
		{
			sStore(mp->mSleepTestSpheres[s].GetCenter(), outDesc.sleep_spheres[s]);
			outDesc.sleep_spheres[s][3] = mp->mSleepTestSpheres[s].GetRadius();
		}
		outDesc.sleep_timer = mp->mSleepTestTimer;
		outDesc.active = 0; // The active list is set explicitly (order matters)
	}
	outDesc.flags = flags;
}

// Export the READ contact cache (what the next step warm starts from)
inline void sExportContactCache(const PhysicsSystem &inSystem, std::vector<b2j_cached_body_pair> &outPairs, std::vector<b2j_cached_manifold> &outManifolds)
{
	const ContactConstraintManager &ccm = inSystem.mContactManager;
	const ContactConstraintManager::ManifoldCache &cache = ccm.mCache[ccm.mCacheWriteIdx ^ 1];

	Array<const ContactConstraintManager::BPKeyValue *> all_bp;
	cache.GetAllBodyPairsSorted(all_bp);
	for (const ContactConstraintManager::BPKeyValue *bp_kv : all_bp)
	{
		const BodyPair &key = bp_kv->GetKey();
		const ContactConstraintManager::CachedBodyPair &cbp = bp_kv->GetValue();
		b2j_cached_body_pair pair;
		memset(&pair, 0, sizeof(pair));
		pair.body1 = key.mBodyA.GetIndexAndSequenceNumber();
		pair.body2 = key.mBodyB.GetIndexAndSequenceNumber();
		sStore(cbp.mDeltaPosition, pair.delta_position);
		sStore(cbp.mDeltaRotation, pair.delta_rotation);
		pair.first_manifold = (uint32_t)outManifolds.size();

		Array<const ContactConstraintManager::MKeyValue *> all_m;
		cache.GetAllManifoldsSorted(cbp, all_m);
		for (const ContactConstraintManager::MKeyValue *m_kv : all_m)
		{
			const SubShapeIDPair &mkey = m_kv->GetKey();
			const ContactConstraintManager::CachedManifold &cm = m_kv->GetValue();
			b2j_cached_manifold m;
			memset(&m, 0, sizeof(m));
			m.sub_shape1 = mkey.GetSubShapeID1().GetValue();
			m.sub_shape2 = mkey.GetSubShapeID2().GetValue();
			sStore(cm.mContactNormal, m.normal);
			m.friction_lambda[0] = cm.mFrictionLambda[0];
			m.friction_lambda[1] = cm.mFrictionLambda[1];
			m.angular_friction_lambda = cm.mAngularFrictionLambda;
			m.num_points = cm.mNumContactPoints;
			m.flags = cm.mFlags;
			for (uint32 i = 0; i < cm.mNumContactPoints && i < 4; ++i)
			{
				sStore(cm.mContactPoints[i].mPosition1, m.position1[i]);
				sStore(cm.mContactPoints[i].mPosition2, m.position2[i]);
				m.non_penetration_lambda[i] = cm.mContactPoints[i].mNonPenetrationLambda;
			}
			outManifolds.push_back(m);
		}
		pair.num_manifolds = (uint32_t)outManifolds.size() - pair.first_manifold;
		outPairs.push_back(pair);
	}
}

// Re-create inSystem inside a new b2j_world. Filters are sampled into tables for inNumObjectLayers object layers.
inline b2j_world *sExportWorld(const Api &inApi, const PhysicsSystem &inSystem, uint inNumObjectLayers, int inDevice, uint inMaxBodyPairs, uint inMaxContactConstraints, String &outError)
{
	const BroadPhaseLayerInterface &bpli = *inSystem.mBodyManager.mBroadPhaseLayerInterface;
	uint num_bp_layers = bpli.GetNumBroadPhaseLayers();

	std::vector<uint8_t> o2bp(inNumObjectLayers), ovbp(inNumObjectLayers * num_bp_layers), ovo(inNumObjectLayers * inNumObjectLayers);
	for (uint o = 0; o < inNumObjectLayers; ++o)
	{
		o2bp[o] = (uint8_t)(BroadPhaseLayer::Type)bpli.GetBroadPhaseLayer((ObjectLayer)o);
		for (uint b = 0; b < num_bp_layers; ++b)
			ovbp[o * num_bp_layers + b] = inSystem.mObjectVsBroadPhaseLayerFilter->ShouldCollide((ObjectLayer)o, BroadPhaseLayer((BroadPhaseLayer::Type)b));
		for (uint o2 = 0; o2 < inNumObjectLayers; ++o2)
			ovo[o * inNumObjectLayers + o2] = inSystem.mObjectLayerPairFilter->ShouldCollide((ObjectLayer)o, (ObjectLayer)o2);
	}

	b2j_world_desc desc;
	memset(&desc, 0, sizeof(desc));
	desc.max_bodies = inSystem.GetMaxBodies();
	desc.max_body_pairs = inMaxBodyPairs;
	desc.max_contact_constraints = inMaxContactConstraints;
	desc.num_object_layers = inNumObjectLayers;
	desc.num_broadphase_layers = num_bp_layers;
	desc.object_to_broadphase = o2bp.data();
	desc.object_vs_broadphase = ovbp.data();
	desc.object_vs_object = ovo.data();
	sFillSettings(inSystem.GetPhysicsSettings(), desc.settings);
	sStore(inSystem.GetGravity(), desc.gravity);
	desc.device = inDevice;

	b2j_world *world = inApi.b2j_world_create(&desc);
	if (world == nullptr)
	{
		outError = String("b2j_world_create failed: ") + inApi.b2j_last_error();
		return nullptr;
	}
	inApi.b2j_world_set_previous_delta_time(world, inSystem.mPreviousStepDeltaTime);

	// Shapes (deduplicated by pointer) and bodies
	std::unordered_map<const Shape *, int32_t> shape_ids;
	BodyIDVector body_ids;
	inSystem.GetBodies(body_ids);
	std::vector<b2j_body_desc> bodies;
	bodies.reserve(body_ids.size());
	const BodyLockInterfaceNoLock &lock_interface = inSystem.GetBodyLockInterfaceNoLock();
	for (BodyID id : body_ids)
	{
		const Body *body = lock_interface.TryGetBody(id);
		if (body == nullptr || !body->IsInBroadPhase() || !body->IsRigidBody())
			continue;
		const Shape *shape = body->GetShape();
		auto it = shape_ids.find(shape);
		if (it == shape_ids.end())
		{
			int32_t sid = sUploadShape(inApi, world, shape, outError);
			if (sid < 0)
			{
				if (outError.empty()) outError = String("shape upload failed: ") + inApi.b2j_last_error();
				inApi.b2j_world_destroy(world);
				return nullptr;
			}
			it = shape_ids.emplace(shape, sid).first;
		}
		b2j_body_desc bd;
		sFillBody(*body, it->second, bd);
		bodies.push_back(bd);
	}
	if (!bodies.empty() && inApi.b2j_bodies_add(world, bodies.data(), (uint32_t)bodies.size()) != 0)
	{
		outError = String("b2j_bodies_add failed: ") + inApi.b2j_last_error();
		inApi.b2j_world_destroy(world);
		return nullptr;
	}

	// Active list in the reference's order (BodyManager::mActiveBodies; not recorded by SaveState)
	BodyIDVector active;
	inSystem.GetActiveBodies(EBodyType::RigidBody, active);
	std::vector<uint32_t> active_ids;
	for (BodyID id : active)
		active_ids.push_back(id.GetIndexAndSequenceNumber());
	if (inApi.b2j_set_active_list(world, active_ids.data(), (uint32_t)active_ids.size()) != 0)
	{
		outError = String("b2j_set_active_list failed: ") + inApi.b2j_last_error();
		inApi.b2j_world_destroy(world);
		return nullptr;
	}

	// Non contact constraints in ConstraintManager order (= Constraint::mConstraintIndex) with the state the next step starts from
	// (what Constraint::SaveState writes: accumulated impulses, the distance constraint's last normal)
	Constraints constraints = inSystem.GetConstraints();
	std::vector<b2j_constraint_desc> cdescs;
	std::vector<b2j_constraint_state> cstates;
	for (const Ref<Constraint> &c : constraints)
	{
		b2j_constraint_desc cd;
		b2j_constraint_state cs;
		memset(&cd, 0, sizeof(cd)); memset(&cs, 0, sizeof(cs));
		if (c->GetType() != EConstraintType::TwoBodyConstraint) { outError = "only two body constraints are on the path"; inApi.b2j_world_destroy(world); return nullptr; }
		const TwoBodyConstraint *tb = static_cast<const TwoBodyConstraint *>(c.GetPtr());
		cd.body1 = tb->GetBody1()->GetID().GetIndexAndSequenceNumber(); cd.body2 = tb->GetBody2()->GetID().GetIndexAndSequenceNumber();
		sStore(tb->GetConstraintToBody1Matrix().GetTranslation(), cd.point1); sStore(tb->GetConstraintToBody2Matrix().GetTranslation(), cd.point2);
		cd.priority = c->GetConstraintPriority();
		cd.num_velocity_steps_override = (uint8_t)c->GetNumVelocityStepsOverride(); cd.num_position_steps_override = (uint8_t)c->GetNumPositionStepsOverride();
		cd.enabled = c->GetEnabled();
		switch (c->GetSubType())
		{
		case EConstraintSubType::Point:
			cd.type = B2J_CONSTRAINT_POINT;
			sStore(static_cast<const PointConstraint *>(tb)->GetTotalLambdaPosition(), cs.total_lambda);
			break;
		case EConstraintSubType::Distance:
			{
				const DistanceConstraint *dc = static_cast<const DistanceConstraint *>(tb);
				if (dc->GetLimitsSpringSettings().HasStiffness()) { outError = "distance constraints with limit springs are not on the path"; inApi.b2j_world_destroy(world); return nullptr; }
				cd.type = B2J_CONSTRAINT_DISTANCE;
				cd.min_distance = dc->GetMinDistance(); cd.max_distance = dc->GetMaxDistance();
				cs.total_lambda[0] = dc->GetTotalLambdaPosition();
				sStore(dc->mWorldSpaceNormal, cs.world_space_normal); // (no getter: the member DistanceConstraint::SaveState writes)
			}
			break;
		case EConstraintSubType::Hinge:
			{
				const HingeConstraint *hc = static_cast<const HingeConstraint *>(tb);
				if (hc->GetMotorState() != EMotorState::Off || hc->GetLimitsSpringSettings().HasStiffness()) { outError = "hinge motors / limit springs are not on the path"; inApi.b2j_world_destroy(world); return nullptr; }
				cd.type = B2J_CONSTRAINT_HINGE;
				sStore(hc->GetLocalSpacePoint1(), cd.point1); sStore(hc->GetLocalSpacePoint2(), cd.point2);
				sStore(hc->GetLocalSpaceHingeAxis1(), cd.hinge_axis1); sStore(hc->GetLocalSpaceHingeAxis2(), cd.hinge_axis2);
				sStore(hc->mInvInitialOrientation, cd.inv_initial_orientation); // (no getter)
				cd.limits_min = hc->GetLimitsMin(); cd.limits_max = hc->GetLimitsMax(); cd.max_friction_torque = hc->GetMaxFrictionTorque();
				sStore(hc->GetTotalLambdaPosition(), cs.total_lambda);
				cs.total_lambda_rotation[0] = hc->GetTotalLambdaRotation()[0]; cs.total_lambda_rotation[1] = hc->GetTotalLambdaRotation()[1];
				cs.total_lambda_limits = hc->GetTotalLambdaRotationLimits(); cs.total_lambda_motor = hc->GetTotalLambdaMotor();
			}
			break;
		case EConstraintSubType::Fixed:
			{
				const FixedConstraint *fc = static_cast<const FixedConstraint *>(tb);
				cd.type = B2J_CONSTRAINT_FIXED;
				sStore(fc->mInvInitialOrientation, cd.inv_initial_orientation); // (no getter)
				sStore(fc->GetTotalLambdaPosition(), cs.total_lambda);
				sStore(fc->GetTotalLambdaRotation(), cs.total_lambda_rotation);
			}
			break;
		default:
			outError = "constraint type is not on the path (PointConstraint, DistanceConstraint, HingeConstraint, FixedConstraint)";
			inApi.b2j_world_destroy(world);
			return nullptr;
		}
		cdescs.push_back(cd); cstates.push_back(cs);
	}
	if (!cdescs.empty() && (inApi.b2j_constraints_add(world, cdescs.data(), (uint32_t)cdescs.size()) != 0
		|| inApi.b2j_constraints_set_state(world, 0, (uint32_t)cstates.size(), cstates.data()) != 0))
	{
		outError = String("b2j_constraints_add failed: ") + inApi.b2j_last_error();
		inApi.b2j_world_destroy(world);
		return nullptr;
	}

	// Contact cache
	std::vector<b2j_cached_body_pair> pairs;
	std::vector<b2j_cached_manifold> manifolds;
	sExportContactCache(inSystem, pairs, manifolds);
	if (inApi.b2j_contact_cache_import(world, pairs.data(), (uint32_t)pairs.size(), manifolds.data(), (uint32_t)manifolds.size()) != 0)
	{
		outError = String("b2j_contact_cache_import failed: ") + inApi.b2j_last_error();
		inApi.b2j_world_destroy(world);
		return nullptr;
	}
	return world;
}

} // namespace b2j_adapter
