// compound_scene.inl -- user code compiled twice (like feature_scenes.inl): against the reference (oracle/ref_harness.cpp) and against the
// facade (facade_capi.cpp). StaticCompoundShape bodies: dumbbells (spheres + capsule), L shapes (boxes), tables (box + 4
// cylinder legs: two tree levels), a 9 part cross with decorated sub shapes, dropped in a heap with plain convex bodies next to a static
// compound staircase with bodies tumbling down it. Variant 0 adds a box floor; variant 1 expects the caller to have added the ground (the
// reference harness adds a terrain mesh) and starts higher.
// The including file provides: B2J_SHAPE_REF, B2J_NEW_SHAPE(Type, args...), B2J_CREATE_COMPOUND(settings), Layers, sRandomQuat(std::mt19937 &).

static void sCompoundCreate(PhysicsSystem &inSystem, int inVariant, const B2J_SHAPE_REF &inHull, uint32_t &outNumDynamic)
{
	BodyInterface &bi = inSystem.GetBodyInterface();
	std::mt19937 random(777 + inVariant);
	Quat z90 = Quat(0.0f, 0.0f, 0.70710678f, 0.70710678f), y45 = Quat(0.0f, 0.38268343f, 0.0f, 0.92387953f);
	B2J_SHAPE_REF sphere = B2J_NEW_SHAPE(SphereShape, 0.4f), capsule = B2J_NEW_SHAPE(CapsuleShape, 0.6f, 0.15f), box = B2J_NEW_SHAPE(BoxShape, Vec3(0.4f, 0.4f, 0.4f)), leg = B2J_NEW_SHAPE(CylinderShape, 0.35f, 0.08f, 0.02f);
	B2J_SHAPE_REF hull = inHull;

	StaticCompoundShapeSettings dumbbell_settings;
	dumbbell_settings.AddShape(Vec3(-0.6f, 0.0f, 0.0f), Quat::sIdentity(), sphere);
	dumbbell_settings.AddShape(Vec3(0.6f, 0.0f, 0.0f), Quat::sIdentity(), sphere);
	dumbbell_settings.AddShape(Vec3::sZero(), z90, capsule);
	B2J_SHAPE_REF dumbbell = B2J_CREATE_COMPOUND(dumbbell_settings);

	StaticCompoundShapeSettings ell_settings;
	ell_settings.AddShape(Vec3::sZero(), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(BoxShape, Vec3(0.6f, 0.2f, 0.2f))));
	ell_settings.AddShape(Vec3(0.4f, 0.7f, 0.0f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(BoxShape, Vec3(0.2f, 0.5f, 0.2f))));
	B2J_SHAPE_REF ell = B2J_CREATE_COMPOUND(ell_settings);

	StaticCompoundShapeSettings table_settings;
	table_settings.AddShape(Vec3(0.0f, 0.7f, 0.0f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(BoxShape, Vec3(0.8f, 0.1f, 0.6f))));
	table_settings.AddShape(Vec3(-0.7f, 0.3f, -0.5f), Quat::sIdentity(), leg);
	table_settings.AddShape(Vec3(0.7f, 0.3f, -0.5f), Quat::sIdentity(), leg);
	table_settings.AddShape(Vec3(-0.7f, 0.3f, 0.5f), Quat::sIdentity(), leg);
	table_settings.AddShape(Vec3(0.7f, 0.3f, 0.5f), Quat::sIdentity(), leg);
	B2J_SHAPE_REF table = B2J_CREATE_COMPOUND(table_settings);

	StaticCompoundShapeSettings cross_settings;
	cross_settings.AddShape(Vec3::sZero(), y45, hull);
	cross_settings.AddShape(Vec3(0.9f, 0.0f, 0.0f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(ScaledShape, box, Vec3(0.5f, 0.3f, 0.3f))));
	cross_settings.AddShape(Vec3(-0.9f, 0.0f, 0.0f), y45, box);
	cross_settings.AddShape(Vec3(0.0f, 0.9f, 0.0f), Quat::sIdentity(), sphere);
	cross_settings.AddShape(Vec3(0.0f, -0.9f, 0.0f), z90, leg);
	cross_settings.AddShape(Vec3(0.0f, 0.0f, 0.9f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(RotatedTranslatedShape, Vec3(0.0f, 0.1f, 0.0f), z90, capsule)));
	cross_settings.AddShape(Vec3(0.0f, 0.0f, -0.9f), Quat::sIdentity(), hull);
	cross_settings.AddShape(Vec3(0.6f, 0.6f, 0.0f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(SphereShape, 0.2f)));
	cross_settings.AddShape(Vec3(-0.6f, -0.6f, 0.0f), y45, B2J_SHAPE_REF(B2J_NEW_SHAPE(BoxShape, Vec3(0.2f, 0.2f, 0.2f))));
	B2J_SHAPE_REF cross = B2J_CREATE_COMPOUND(cross_settings);

	if (inVariant != 1)
	{
		BodyCreationSettings floor(B2J_NEW_SHAPE(BoxShape, Vec3(60.0f, 1.0f, 60.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		bi.CreateAndAddBody(floor, EActivation::DontActivate);
	}
	{
		// a static staircase made of one compound (8 sub shapes: a two level tree)
		StaticCompoundShapeSettings stairs;
		for (int i = 0; i < 8; ++i)
			stairs.AddShape(Vec3(0.8f * float(i), 0.2f + 0.4f * float(i), 0.0f), Quat::sIdentity(), B2J_SHAPE_REF(B2J_NEW_SHAPE(BoxShape, Vec3(0.4f, 0.2f, 2.0f))));
		BodyCreationSettings s(B2J_CREATE_COMPOUND(stairs), RVec3(6.0f, inVariant == 1? 4.0f : 0.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING);
		bi.CreateAndAddBody(s, EActivation::DontActivate);
	}
	B2J_SHAPE_REF shapes[6] = { dumbbell, ell, table, cross, box, sphere };
	float y0 = inVariant == 1? 7.0f : 1.5f;
	for (int i = 0; i < 36; ++i)
	{
		BodyCreationSettings s(shapes[i % 6], RVec3(-3.0f + 2.2f * float(i % 4), y0 + 1.9f * float(i / 4), -2.0f + 2.1f * float((i / 2) % 3)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
		s.mFriction = 0.5f;
		bi.CreateAndAddBody(s, EActivation::Activate);
		outNumDynamic++;
	}
	for (int i = 0; i < 6; ++i) // bodies tumbling down the staircase
	{
		BodyCreationSettings s(shapes[(i + 1) % 6], RVec3(6.5f + 0.9f * float(i), (inVariant == 1? 4.0f : 0.0f) + 2.5f + 0.6f * float(i), -1.0f + 0.4f * float(i)), sRandomQuat(random), EMotionType::Dynamic, Layers::MOVING);
		bi.CreateAndAddBody(s, EActivation::Activate);
		outNumDynamic++;
	}
}
