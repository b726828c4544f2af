// feature_scenes.inl -- the same USER code compiled twice (like api_tour.inl): against the reference (oracle/ref_harness.cpp) and
// against the facade (facade_capi.cpp). Small worlds that exercise every motion type / body flag / per body override the step has a
// branch for; patterns follow UnitTests/Physics/SensorTests.cpp (sensor vs dynamic / kinematic / static), PhysicsTests.cpp
// (kinematic movers, allowed DOFs, gyroscopic force, solver step overrides, broadphase layers), ContactListenerTests.cpp (manifold
// reduction off).
// The including file provides: B2J_SHAPE_REF, B2J_NEW_SHAPE(Type, args...), Layers::{NON_MOVING, MOVING, DEBRIS}, sRandomQuat(std::mt19937 &).
// Variants: 0 kinematic, 1 sensor, 2 dof_plane2d, 3 gyroscopic, 4 step_overrides, 5 no_manifold_reduction, 6 two_moving_layers,
// 7 kinematic_vs_nondynamic, 8 zoo (0..7 in one world), 9 decorated (ScaledShape / RotatedTranslatedShape around convex shapes),
// 10 cylinder (CylinderShape plain, scaled and rotated against every other convex shape),
// 11 joints (PointConstraint / DistanceConstraint / HingeConstraint / FixedConstraint: chain, rope with limits, doors and flaps on hinges with limits and friction, a cloth that is one large island, kinematic tow, constraint
// that wakes a sleeping body, priorities, solver step overrides, a disabled constraint; patterns of UnitTests/Physics/DistanceConstraintTests.cpp
// and Samples/Tests/Constraints/{PointConstraintTest, DistanceConstraintTest}.cpp).
// inHull: a cooked convex hull (cooking is host side and out of scope).

static void sFeatureCreate(PhysicsSystem &inSystem, int inVariant, const B2J_SHAPE_REF &inHull, uint32_t &outNumDynamic)
{
	BodyInterface &bi = inSystem.GetBodyInterface();
	std::mt19937 random(4321 + inVariant);
	bool zoo = inVariant == 8;
	auto add = [&](BodyCreationSettings &s, EActivation a = EActivation::Activate) { BodyID id = bi.CreateAndAddBody(s, a); if (s.mMotionType != EMotionType::Static) outNumDynamic++; return id; };
	{
		BodyCreationSettings floor(B2J_NEW_SHAPE(BoxShape, Vec3(100.0f, 1.0f, 100.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		floor.mFriction = 0.6f;
		add(floor, EActivation::DontActivate);
	}
	B2J_SHAPE_REF box = B2J_NEW_SHAPE(BoxShape, Vec3::sReplicate(0.5f)), sphere = B2J_NEW_SHAPE(SphereShape, 0.5f), capsule = B2J_NEW_SHAPE(CapsuleShape, 0.5f, 0.3f), slab = B2J_NEW_SHAPE(BoxShape, Vec3(2.0f, 0.25f, 2.0f));
	B2J_SHAPE_REF hull = inHull;
	float x0 = 0.0f; // every feature gets its own strip of the floor along x when they share a world (zoo)

	if (inVariant == 0 || zoo)
	{
		// kinematic mover pushing a stack + a rotating kinematic platform carrying dynamic bodies + a kinematic that is put to sleep
		for (int i = 0; i < 9; ++i)
		{
			BodyCreationSettings s(i % 2? box : hull, RVec3(x0 + 2.0f + 1.05f * float(i % 3), 0.55f + 1.1f * float(i / 3), 0.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		BodyCreationSettings pusher(B2J_NEW_SHAPE(BoxShape, Vec3(0.5f, 1.5f, 2.0f)), RVec3(x0 - 0.5f, 1.6f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		pusher.mLinearVelocity = Vec3(2.5f, 0.0f, 0.0f);
		pusher.mAngularVelocity = Vec3(0.0f, 0.4f, 0.0f);
		add(pusher);
		BodyCreationSettings platform(slab, RVec3(x0 + 2.0f, 3.0f, 6.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		platform.mAngularVelocity = Vec3(0.0f, 1.0f, 0.3f);
		platform.mLinearVelocity = Vec3(0.0f, 0.5f, 0.0f);
		platform.mFriction = 0.9f;
		add(platform);
		for (int i = 0; i < 4; ++i)
		{
			BodyCreationSettings s(i % 2? sphere : capsule, RVec3(x0 + 1.0f + float(i), 3.9f, 6.0f + 0.5f * float(i % 2)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		BodyCreationSettings idle(box, RVec3(x0 + 8.0f, 0.5f, -4.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING); // zero velocity: goes to sleep
		add(idle);
		x0 += 20.0f;
	}
	if (inVariant == 1 || zoo)
	{
		// sensors: static sensor volume, kinematic sensor sweeping through dynamic / static / sleeping bodies, dynamic sensor (falls through
		// everything but still reports), kinematic non sensor vs static sensor (the kinematic vs sensor pair rule, Body.inl:41-44)
		BodyCreationSettings trigger(B2J_NEW_SHAPE(BoxShape, Vec3(3.0f, 1.0f, 3.0f)), RVec3(x0, 2.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::MOVING);
		trigger.mIsSensor = true;
		add(trigger, EActivation::DontActivate);
		for (int i = 0; i < 6; ++i)
		{
			BodyCreationSettings s(i % 3 == 0? sphere : (i % 3 == 1? box : hull), RVec3(x0 - 2.0f + 0.9f * float(i), 3.6f + 0.3f * float(i % 2), -1.0f + 0.7f * float(i % 3)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		BodyCreationSettings sweeper(B2J_NEW_SHAPE(BoxShape, Vec3(0.5f, 2.0f, 3.0f)), RVec3(x0 - 6.0f, 1.5f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		sweeper.mIsSensor = true;
		sweeper.mLinearVelocity = Vec3(3.0f, 0.0f, 0.0f);
		add(sweeper);
		BodyCreationSettings kin(box, RVec3(x0 + 6.0f, 2.0f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING); // non sensor kinematic entering the static sensor
		kin.mLinearVelocity = Vec3(-2.0f, 0.0f, 0.0f);
		add(kin);
		BodyCreationSettings ghost(sphere, RVec3(x0, 5.5f, 2.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING); // dynamic sensor
		ghost.mIsSensor = true;
		add(ghost);
		BodyCreationSettings sleeper(box, RVec3(x0 - 3.5f, 0.5f, 1.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING); // asleep: a sensor must not wake it
		add(sleeper, EActivation::DontActivate);
		x0 += 20.0f;
	}
	if (inVariant == 2 || zoo)
	{
		// EAllowedDOFs: a 2D stack (Plane2D), translation only and rotation only bodies hit by free bodies
		for (int i = 0; i < 8; ++i)
		{
			BodyCreationSettings s(i % 2? box : capsule, RVec3(x0 + 0.6f * float(i % 2), 0.6f + 1.15f * float(i), 0.0f), Quat(0.0f, 0.0f, 0.1f * float(i), 1.0f).Normalized(), EMotionType::Dynamic, Layers::MOVING);
			s.mAllowedDOFs = EAllowedDOFs::Plane2D;
			add(s);
		}
		BodyCreationSettings slider(box, RVec3(x0 + 4.0f, 0.5f, 0.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		slider.mAllowedDOFs = EAllowedDOFs::TranslationX | EAllowedDOFs::TranslationZ;
		slider.mLinearVelocity = Vec3(-1.0f, 0.0f, 0.5f);
		add(slider);
		BodyCreationSettings spinner(B2J_NEW_SHAPE(BoxShape, Vec3(1.5f, 0.2f, 0.3f)), RVec3(x0 + 4.0f, 2.5f, 3.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		spinner.mAllowedDOFs = EAllowedDOFs::RotationY | EAllowedDOFs::RotationX;
		spinner.mAngularVelocity = Vec3(0.0f, 3.0f, 0.0f);
		spinner.mGravityFactor = 0.0f;
		add(spinner);
		for (int i = 0; i < 3; ++i)
		{
			BodyCreationSettings s(sphere, RVec3(x0 + 3.0f + float(i), 4.0f + float(i), 3.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		x0 += 20.0f;
	}
	if (inVariant == 3 || zoo)
	{
		// gyroscopic force (MotionProperties::ApplyGyroscopicForceInternal): tumbling T-handle like boxes, some of them landing on the floor
		for (int i = 0; i < 6; ++i)
		{
			BodyCreationSettings s(B2J_NEW_SHAPE(BoxShape, Vec3(0.2f + 0.1f * float(i % 3), 0.5f, 1.0f)), RVec3(x0 + 2.0f * float(i), 1.2f + 0.8f * float(i % 3), 0.0f), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			s.mApplyGyroscopicForce = true;
			s.mAngularVelocity = Vec3(0.1f, 8.0f + float(i), 0.2f);
			s.mAngularDamping = 0.0f;
			s.mMaxAngularVelocity = 60.0f;
			s.mGravityFactor = i < 3? 0.0f : 1.0f;
			add(s);
		}
		x0 += 20.0f;
	}
	if (inVariant == 4 || zoo)
	{
		// per body solver step overrides (CalculateSolverSteps.h): islands with 0 / higher / lower iteration counts than the default
		for (int stack = 0; stack < 4; ++stack)
			for (int i = 0; i < 4; ++i)
			{
				BodyCreationSettings s(i % 2? box : sphere, RVec3(x0 + 3.0f * float(stack), 0.55f + 1.05f * float(i), 0.1f * float(i)), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
				if (i == 1)
				{
					s.mNumVelocityStepsOverride = stack == 0? 0 : (stack == 1? 14 : (stack == 2? 3 : 1));
					s.mNumPositionStepsOverride = stack == 0? 0 : (stack == 1? 5 : (stack == 2? 1 : 4));
				}
				if (stack == 2 && i != 1) { s.mNumVelocityStepsOverride = 2; s.mNumPositionStepsOverride = 1; } // every body overrides: the default does not apply
				if (stack == 3 && i == 2) s.mNumVelocityStepsOverride = 12;
				add(s);
			}
		x0 += 20.0f;
	}
	if (inVariant == 5 || zoo)
	{
		// manifold reduction switched off per body: hull / box / capsule resting on each other and on a reduced body
		for (int i = 0; i < 8; ++i)
		{
			BodyCreationSettings s(i % 4 == 0? hull : (i % 4 == 1? box : (i % 4 == 2? capsule : slab)), RVec3(x0 + 0.3f * float(i % 3), 0.7f + 1.2f * float(i), 0.2f * float(i % 2)), i % 4 == 3? Quat::sIdentity() : sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			s.mUseManifoldReduction = (i % 3) == 0;
			add(s);
		}
		x0 += 20.0f;
	}
	if (inVariant == 6 || zoo)
	{
		// two moving broadphase layers: DEBRIS lives in its own tree, collides with MOVING and NON_MOVING but not with itself
		for (int i = 0; i < 12; ++i)
		{
			BodyCreationSettings s(i % 2? box : sphere, RVec3(x0 + 0.8f * float(i % 4), 0.6f + 1.1f * float(i / 4), 0.7f * float(i % 2)), Quat::sIdentity(), EMotionType::Dynamic, (i % 3) == 0? Layers::MOVING : Layers::DEBRIS);
			add(s);
		}
		BodyCreationSettings kin(slab, RVec3(x0 + 1.0f, 5.0f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::DEBRIS);
		kin.mLinearVelocity = Vec3(0.0f, -1.0f, 0.0f);
		add(kin);
		x0 += 20.0f;
	}
	if (inVariant == 7 || zoo)
	{
		// CollideKinematicVsNonDynamic: a kinematic body with the flag reports contacts against static and kinematic bodies
		BodyCreationSettings probe(box, RVec3(x0, 0.45f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		probe.mCollideKinematicVsNonDynamic = true;
		probe.mLinearVelocity = Vec3(1.5f, 0.0f, 0.0f);
		add(probe);
		BodyCreationSettings other(box, RVec3(x0 + 1.5f, 0.5f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING); // no flag
		other.mLinearVelocity = Vec3(0.2f, 0.0f, 0.0f);
		add(other);
		BodyCreationSettings pillar(B2J_NEW_SHAPE(BoxShape, Vec3(0.5f, 2.0f, 0.5f)), RVec3(x0 + 4.0f, 2.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::MOVING);
		add(pillar, EActivation::DontActivate);
		BodyCreationSettings dyn(sphere, RVec3(x0 + 2.6f, 0.5f, 0.3f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		add(dyn);
		x0 += 20.0f;
	}
	if (inVariant == 9)
	{
		// decorated convex shapes: every leaf type scaled (non uniform where the leaf allows it), rotated + translated, and both nested
		// either way round; dynamic bodies dropped in a heap, a static rotated slab as a ramp, a kinematic scaled pusher
		Quat tilt = Quat(0.0f, 0.38268343f, 0.0f, 0.92387953f), roll = Quat(0.25881905f, 0.0f, 0.0f, 0.96592583f), yaw = Quat(0.0f, 0.0f, 0.70710678f, 0.70710678f);
		B2J_SHAPE_REF shapes[10] = {
			B2J_NEW_SHAPE(ScaledShape, box, Vec3(1.5f, 0.5f, 0.8f)),
			B2J_NEW_SHAPE(ScaledShape, hull, Vec3(0.7f, 1.3f, 1.1f)),
			B2J_NEW_SHAPE(ScaledShape, sphere, Vec3::sReplicate(1.4f)),
			B2J_NEW_SHAPE(ScaledShape, capsule, Vec3::sReplicate(0.8f)),
			B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.3f, 0.2f, -0.1f), tilt, box),
			B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(-0.2f, 0.4f, 0.0f), roll, hull),
			B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.0f, 0.5f, 0.0f), yaw, capsule),
			B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.1f, -0.3f, 0.2f), roll, B2J_SHAPE_REF(B2J_NEW_SHAPE(ScaledShape, box, Vec3(0.6f, 1.2f, 0.9f)))),
			B2J_NEW_SHAPE(ScaledShape, B2J_SHAPE_REF(B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.2f, 0.0f, 0.3f), tilt, hull)), Vec3::sReplicate(1.25f)),
			box };
		for (int i = 0; i < 30; ++i)
		{
			BodyCreationSettings s(shapes[i % 10], RVec3(-2.0f + 1.3f * float(i % 4), 1.2f + 1.1f * float(i / 4), -1.5f + 1.4f * float((i / 2) % 3)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			s.mFriction = 0.4f + 0.05f * float(i % 5);
			s.mRestitution = i % 6 == 0? 0.5f : 0.0f;
			add(s);
		}
		BodyCreationSettings ramp(B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.0f, 0.5f, 0.0f), roll, slab), RVec3(5.0f, 0.3f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		add(ramp, EActivation::DontActivate);
		for (int i = 0; i < 4; ++i)
		{
			BodyCreationSettings s(shapes[(3 * i + 1) % 10], RVec3(4.0f + 0.7f * float(i), 2.5f + 0.9f * float(i), -0.6f + 0.4f * float(i)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		BodyCreationSettings pusher(B2J_NEW_SHAPE(ScaledShape, box, Vec3(1.0f, 3.0f, 4.0f)), RVec3(-6.0f, 1.6f, 0.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		pusher.mLinearVelocity = Vec3(1.5f, 0.0f, 0.0f);
		add(pusher);
	}
	if (inVariant == 10)
	{
		// CylinderShape: standing and lying stacks (cap on cap, side on cap, side on side), cylinders rolling down a tilted slab, thin discs,
		// zero convex radius, scaled (non uniform: xz / y) and rotated cylinders, mixed with the other convex shapes
		B2J_SHAPE_REF cyl = B2J_NEW_SHAPE(CylinderShape, 0.5f, 0.4f), disc = B2J_NEW_SHAPE(CylinderShape, 0.08f, 0.7f), sharp = B2J_NEW_SHAPE(CylinderShape, 0.4f, 0.3f, 0.0f);
		Quat lying = Quat(0.0f, 0.0f, 0.70710678f, 0.70710678f), tilt = Quat(0.0f, 0.0f, 0.08715574f, 0.9961947f);
		B2J_SHAPE_REF shapes[8] = {
			cyl, disc, sharp,
			B2J_NEW_SHAPE(ScaledShape, cyl, Vec3(1.5f, 0.6f, 1.5f)),
			B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.0f, 0.2f, 0.1f), lying, cyl),
			box, hull, capsule };
		for (int i = 0; i < 5; ++i) // standing stack
		{
			BodyCreationSettings s(i % 2? cyl : sharp, RVec3(0.0f, 0.52f + 1.02f * float(i), 0.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		for (int i = 0; i < 4; ++i) // lying, side on side
		{
			BodyCreationSettings s(cyl, RVec3(3.0f + 0.85f * float(i % 2), 0.42f + 0.8f * float(i / 2), 0.0f), lying, EMotionType::Dynamic, Layers::MOVING);
			add(s);
		}
		BodyCreationSettings ramp(slab, RVec3(-5.0f, 1.0f, 0.0f), tilt, EMotionType::Static, Layers::NON_MOVING);
		add(ramp, EActivation::DontActivate);
		for (int i = 0; i < 3; ++i) // rolling down the ramp
		{
			BodyCreationSettings s(i == 1? disc : cyl, RVec3(-6.0f + 1.0f * float(i), 2.0f, -1.0f + 1.0f * float(i)), Quat(0.70710678f, 0.0f, 0.0f, 0.70710678f), EMotionType::Dynamic, Layers::MOVING);
			s.mFriction = 0.8f;
			add(s);
		}
		for (int i = 0; i < 24; ++i) // a heap of everything
		{
			BodyCreationSettings s(shapes[i % 8], RVec3(8.0f + 1.2f * float(i % 3), 1.0f + 1.0f * float(i / 3), -1.2f + 1.2f * float((i / 2) % 3)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			s.mRestitution = i % 5 == 0? 0.4f : 0.0f;
			add(s);
		}
	}
	if (inVariant == 11)
	{
		auto create = [&](BodyCreationSettings &bs, EActivation a = EActivation::Activate) { Body *b = bi.CreateBody(bs); bi.AddBody(b->GetID(), a); if (bs.mMotionType != EMotionType::Static) outNumDynamic++; return b; };
		auto point = [&](Body *b1, Body *b2, RVec3 p) { PointConstraintSettings ps; ps.mPoint1 = p; ps.mPoint2 = p; TwoBodyConstraint *c = ps.Create(*b1, *b2); inSystem.AddConstraint(c); return c; };
		auto distance = [&](Body *b1, Body *b2, RVec3 p1, RVec3 p2, float mn, float mx) { DistanceConstraintSettings ds; ds.mPoint1 = p1; ds.mPoint2 = p2; ds.mMinDistance = mn; ds.mMaxDistance = mx; TwoBodyConstraint *c = ds.Create(*b1, *b2); inSystem.AddConstraint(c); return c; };
		B2J_SHAPE_REF small_sphere = B2J_NEW_SHAPE(SphereShape, 0.2f), link = B2J_NEW_SHAPE(CapsuleShape, 0.35f, 0.12f);
		// (a) a chain of capsules pinned to a static anchor, starting horizontal: swings down, the links collide with a post
		BodyCreationSettings anchor_s(small_sphere, RVec3(0.0f, 8.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		Body *anchor = create(anchor_s, EActivation::DontActivate);
		Body *prev = anchor;
		Quat along_x = Quat(0.0f, 0.0f, 0.70710678f, 0.70710678f);
		for (int i = 0; i < 8; ++i)
		{
			BodyCreationSettings s(link, RVec3(0.5f + 1.0f * float(i), 8.0f, 0.0f), along_x, EMotionType::Dynamic, Layers::MOVING);
			Body *b = create(s);
			TwoBodyConstraint *c = point(prev, b, RVec3(1.0f * float(i), 8.0f, 0.0f));
			if (i == 3) c->SetNumVelocityStepsOverride(14);
			if (i == 5) c->SetNumPositionStepsOverride(4);
			prev = b;
		}
		BodyCreationSettings post(B2J_NEW_SHAPE(BoxShape, Vec3(0.3f, 3.0f, 1.0f)), RVec3(2.5f, 3.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		create(post, EActivation::DontActivate);
		// (b) a rope of spheres: distance constraints with a minimum and a maximum (slack at the start: inactive until stretched), the last
		// link a fixed distance (min = max = the distance at creation)
		BodyCreationSettings anchor2_s(small_sphere, RVec3(10.0f, 9.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		prev = create(anchor2_s, EActivation::DontActivate);
		for (int i = 0; i < 6; ++i)
		{
			BodyCreationSettings s(sphere, RVec3(10.6f + 0.7f * float(i), 9.0f - 0.2f * float(i), 0.3f * float(i % 2)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
			Body *b = create(s);
			RVec3 p1 = prev->GetCenterOfMassPosition(), p2 = b->GetCenterOfMassPosition();
			if (i < 5) distance(prev, b, p1, p2 + Vec3(0.0f, 0.3f, 0.0f), 0.4f, 1.2f);
			else distance(prev, b, p1, p2, -1.0f, -1.0f);
			prev = b;
		}
		// (c) a cloth of small spheres held together by fixed distances: one island of > 128 constraints (the large island splitter colours
		// contacts and constraints), two corners pinned to static anchors, falling over a box
		const int n = 10;
		Body *cloth[n][n];
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
			{
				BodyCreationSettings s(small_sphere, RVec3(20.0f + 0.5f * float(i), 4.0f, -2.25f + 0.5f * float(j)), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
				s.mFriction = 0.4f;
				cloth[i][j] = create(s);
			}
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
			{
				if (i + 1 < n) distance(cloth[i][j], cloth[i + 1][j], cloth[i][j]->GetCenterOfMassPosition(), cloth[i + 1][j]->GetCenterOfMassPosition(), -1.0f, -1.0f);
				if (j + 1 < n) distance(cloth[i][j], cloth[i][j + 1], cloth[i][j]->GetCenterOfMassPosition(), cloth[i][j + 1]->GetCenterOfMassPosition(), -1.0f, -1.0f);
			}
		for (int k = 0; k < 2; ++k)
		{
			Body *corner = cloth[0][k == 0? 0 : n - 1];
			BodyCreationSettings pin_s(small_sphere, corner->GetCenterOfMassPosition() + Vec3(-0.5f, 0.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
			Body *pin = create(pin_s, EActivation::DontActivate);
			point(pin, corner, pin->GetCenterOfMassPosition());
		}
		BodyCreationSettings table(box, RVec3(23.0f, 2.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		create(table, EActivation::DontActivate);
		// (d) a kinematic tractor towing a dynamic box over the floor with a point constraint, and a second box on a leash (max distance)
		BodyCreationSettings tractor_s(box, RVec3(-8.0f, 0.5f, 5.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
		tractor_s.mLinearVelocity = Vec3(1.5f, 0.0f, 0.0f);
		Body *tractor = create(tractor_s);
		BodyCreationSettings trailer_s(box, RVec3(-10.0f, 0.5f, 5.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *trailer = create(trailer_s);
		point(tractor, trailer, RVec3(-9.0f, 0.5f, 5.0f))->SetConstraintPriority(5);
		BodyCreationSettings dog_s(hull, RVec3(-12.0f, 0.6f, 5.5f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *dog = create(dog_s);
		distance(trailer, dog, trailer->GetCenterOfMassPosition(), dog->GetCenterOfMassPosition(), 0.0f, 2.5f)->SetConstraintPriority(2);
		// (e) a sleeping box tied to a box that falls: the constraint is active because one of its bodies is, and wakes the other up
		BodyCreationSettings sleeper_s(box, RVec3(-8.0f, 0.5f, -5.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *sleeper = create(sleeper_s, EActivation::DontActivate);
		BodyCreationSettings faller_s(sphere, RVec3(-8.0f, 4.0f, -5.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *faller = create(faller_s);
		distance(sleeper, faller, sleeper->GetCenterOfMassPosition(), faller->GetCenterOfMassPosition(), 0.0f, 5.0f);
		// ... and two sleeping boxes tied together (inactive constraint until something touches them), a disabled constraint
		BodyCreationSettings rest1_s(box, RVec3(-14.0f, 0.5f, -5.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING), rest2_s(box, RVec3(-12.5f, 0.5f, -5.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *rest1 = create(rest1_s, EActivation::DontActivate), *rest2 = create(rest2_s, EActivation::DontActivate);
		point(rest1, rest2, RVec3(-13.25f, 0.5f, -5.0f));
		BodyCreationSettings free1_s(sphere, RVec3(-14.0f, 3.0f, -8.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING), free2_s(sphere, RVec3(-12.0f, 3.0f, -8.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		free2_s.mLinearVelocity = Vec3(3.0f, 0.0f, 0.0f);
		Body *free1 = create(free1_s), *free2 = create(free2_s);
		distance(free1, free2, free1->GetCenterOfMassPosition(), free2->GetCenterOfMassPosition(), -1.0f, -1.0f)->SetEnabled(false);
		// (f) hinges (Samples/Tests/Constraints/HingeConstraintTest.cpp pattern): a chain of planks hinged about z with limits that
		// starts horizontal from a static post, a door about y with friction pushed by a ball, a free flap about x (no limits), a flap whose
		// two hinge axes start misaligned (the rotation part pulls them together)
		auto hinge = [&](Body *b1, Body *b2, RVec3 p, Vec3 axis, Vec3 normal, float mn, float mx, float friction) {
			HingeConstraintSettings hs; hs.mPoint1 = p; hs.mPoint2 = p; hs.mHingeAxis1 = axis; hs.mHingeAxis2 = axis; hs.mNormalAxis1 = normal; hs.mNormalAxis2 = normal;
			hs.mLimitsMin = mn; hs.mLimitsMax = mx; hs.mMaxFrictionTorque = friction;
			TwoBodyConstraint *c = hs.Create(*b1, *b2); inSystem.AddConstraint(c); return c; };
		const float pi = 3.14159265358979323846f;
		B2J_SHAPE_REF plank = B2J_NEW_SHAPE(BoxShape, Vec3(0.5f, 0.1f, 0.4f));
		BodyCreationSettings hpost_s(box, RVec3(0.0f, 6.0f, 10.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		prev = create(hpost_s, EActivation::DontActivate);
		for (int i = 0; i < 5; ++i)
		{
			BodyCreationSettings s(plank, RVec3(1.0f + 1.0f * float(i), 6.0f, 10.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			Body *b = create(s);
			hinge(prev, b, RVec3(0.5f + 1.0f * float(i), 6.0f, 10.0f), Vec3::sAxisZ(), Vec3::sAxisX(), i % 2? -0.3f * pi : -0.1f * pi, 0.25f * pi, 0.0f);
			prev = b;
		}
		BodyCreationSettings frame_s(B2J_NEW_SHAPE(BoxShape, Vec3(0.1f, 1.0f, 0.1f)), RVec3(8.0f, 1.0f, 10.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		Body *frame = create(frame_s, EActivation::DontActivate);
		BodyCreationSettings door_s(B2J_NEW_SHAPE(BoxShape, Vec3(0.6f, 0.9f, 0.05f)), RVec3(8.75f, 1.05f, 10.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *door = create(door_s);
		hinge(frame, door, RVec3(8.1f, 1.05f, 10.0f), Vec3::sAxisY(), Vec3::sAxisX(), -0.5f * pi, 0.5f * pi, 2.0f);
		BodyCreationSettings ball_s(sphere, RVec3(9.0f, 1.0f, 7.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		ball_s.mLinearVelocity = Vec3(0.0f, 1.0f, 6.0f);
		create(ball_s);
		BodyCreationSettings bar_s(B2J_NEW_SHAPE(BoxShape, Vec3(1.0f, 0.1f, 0.1f)), RVec3(14.0f, 4.0f, 10.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		Body *bar = create(bar_s, EActivation::DontActivate);
		BodyCreationSettings flap_s(B2J_NEW_SHAPE(BoxShape, Vec3(0.8f, 0.05f, 0.6f)), RVec3(14.0f, 4.0f, 10.8f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		Body *flap = create(flap_s);
		hinge(bar, flap, RVec3(14.0f, 4.0f, 10.1f), Vec3::sAxisX(), Vec3::sAxisY(), -pi, pi, 0.0f);
		BodyCreationSettings flap2_s(B2J_NEW_SHAPE(BoxShape, Vec3(0.8f, 0.05f, 0.6f)), RVec3(14.0f, 4.0f, 9.2f), Quat(0.0f, 0.05f, 0.0f, 0.99874922f), EMotionType::Dynamic, Layers::MOVING);
		Body *flap2 = create(flap2_s);
		{
			HingeConstraintSettings hs; hs.mPoint1 = RVec3(14.0f, 4.0f, 9.9f); hs.mPoint2 = RVec3(14.0f, 4.05f, 9.85f);
			hs.mHingeAxis1 = Vec3::sAxisX(); hs.mNormalAxis1 = Vec3::sAxisY();
			hs.mHingeAxis2 = Vec3(0.98006658f, 0.0f, 0.19866933f); hs.mNormalAxis2 = Vec3::sAxisY();
			hs.mLimitsMin = -0.2f * pi; hs.mLimitsMax = 0.6f * pi;
			inSystem.AddConstraint(hs.Create(*bar, *flap2));
		}
		// (g) fixed constraints (Samples/Tests/Constraints/FixedConstraintTest.cpp pattern): a cantilever of boxes welded to a static wall
		// (auto detected anchor points) that a falling box lands on, two bodies welded with rotated reference frames that tumble together,
		// a body welded to a kinematic carrier, a weld with a DOF locked body (singular summed inertia: locked axes get identity columns)
		BodyCreationSettings wall_s(box, RVec3(0.0f, 3.0f, -12.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		prev = create(wall_s, EActivation::DontActivate);
		for (int i = 0; i < 4; ++i)
		{
			BodyCreationSettings s(box, RVec3(1.0f + 1.0f * float(i), 3.0f, -12.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			Body *b = create(s);
			FixedConstraintSettings fs; fs.mAutoDetectPoint = true;
			inSystem.AddConstraint(fs.Create(*prev, *b));
			prev = b;
		}
		BodyCreationSettings load_s(hull, RVec3(3.8f, 5.5f, -12.0f), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
		create(load_s);
		{
			Quat q1 = sRandomQuat(random), q2 = sRandomQuat(random);
			BodyCreationSettings a_s(capsule, RVec3(8.0f, 5.0f, -12.0f), q1, EMotionType::Dynamic, Layers::MOVING), b_s(box, RVec3(8.9f, 5.3f, -12.2f), q2, EMotionType::Dynamic, Layers::MOVING);
			a_s.mAngularVelocity = Vec3(1.0f, 2.0f, -1.5f);
			Body *a = create(a_s), *b = create(b_s);
			FixedConstraintSettings fs;
			fs.mPoint1 = RVec3(8.4f, 5.1f, -12.1f); fs.mPoint2 = RVec3(8.45f, 5.15f, -12.1f);
			fs.mAxisX1 = Vec3::sAxisX(); fs.mAxisY1 = Vec3::sAxisY();
			fs.mAxisX2 = Vec3(0.0f, 0.0f, 1.0f); fs.mAxisY2 = Vec3::sAxisY();
			inSystem.AddConstraint(fs.Create(*a, *b));
		}
		{
			BodyCreationSettings carrier_s(slab, RVec3(14.0f, 2.0f, -12.0f), Quat::sIdentity(), EMotionType::Kinematic, Layers::MOVING);
			carrier_s.mAngularVelocity = Vec3(0.0f, 0.8f, 0.0f); carrier_s.mLinearVelocity = Vec3(0.0f, 0.0f, 0.3f);
			Body *carrier = create(carrier_s);
			BodyCreationSettings rider_s(sphere, RVec3(15.0f, 2.75f, -12.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			Body *rider = create(rider_s);
			FixedConstraintSettings fs; fs.mAutoDetectPoint = true;
			inSystem.AddConstraint(fs.Create(*carrier, *rider));
			BodyCreationSettings planar_s(box, RVec3(18.0f, 3.0f, -12.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING), mate_s(box, RVec3(19.0f, 3.0f, -12.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			planar_s.mAllowedDOFs = EAllowedDOFs::Plane2D; mate_s.mAllowedDOFs = EAllowedDOFs::Plane2D;
			Body *planar = create(planar_s), *mate = create(mate_s);
			FixedConstraintSettings fs2; fs2.mAutoDetectPoint = true;
			inSystem.AddConstraint(fs2.Create(*planar, *mate));
		}
	}
}
