// api_tour.inl -- the same USER code compiled twice: against the reference (oracle/ref_harness.cpp, namespace JPH) and against the
// facade (facade_capi.cpp, namespace JPH_B200). It walks the BodyInterface / PhysicsSystem surface of SURVEY 8(b) that the
// benchmark scenes do not touch (bulk add, remove / destroy / re-add, impulses, per body parameters, activation, queries) so that
// tests/test_facade.py can check that the facade behaves like the reference call for call.
// The including file provides: B2J_SHAPE_REF (shape reference type), B2J_NEW_SHAPE(Type, args...), Layers::MOVING / NON_MOVING.

static void sApiTourCreate(PhysicsSystem &inSystem, std::vector<BodyID> &outBodies)
{
	BodyInterface &bi = inSystem.GetBodyInterface();
	bi.CreateAndAddBody(BodyCreationSettings(B2J_NEW_SHAPE(BoxShape, Vec3(50.0f, 1.0f, 50.0f), 0.0f), RVec3(0.0f, -1.0f, 0.0f), Quat::sIdentity(), EMotionType::Static, Layers::NON_MOVING), EActivation::DontActivate);
	B2J_SHAPE_REF shapes[3] = { B2J_NEW_SHAPE(SphereShape, 0.5f), B2J_NEW_SHAPE(BoxShape, Vec3(0.5f, 0.4f, 0.3f)), B2J_NEW_SHAPE(CapsuleShape, 0.4f, 0.3f) };
	for (int i = 0; i < 12; ++i)
	{
		RVec3 position(-3.0f + 1.5f * float(i % 4), 0.6f + 1.2f * float(i / 4), 0.3f * float(i % 3));
		BodyCreationSettings settings(shapes[i % 3], position, Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		settings.mFriction = 0.4f;
		settings.mRestitution = 0.1f;
		outBodies.push_back(bi.CreateAndAddBody(settings, EActivation::Activate));
	}
}

static void sApiTourMutate(PhysicsSystem &inSystem, std::vector<BodyID> &ioBodies, int inPhase)
{
	BodyInterface &bi = inSystem.GetBodyInterface();
	if (inPhase == 1)
	{
		// parameters, impulses, velocities, forces
		bi.SetFriction(ioBodies[0], 0.05f);
		bi.SetRestitution(ioBodies[1], 0.8f);
		bi.SetGravityFactor(ioBodies[2], 0.25f);
		bi.AddImpulse(ioBodies[3], Vec3(800.0f, 0.0f, 0.0f));
		bi.SetMaxLinearVelocity(ioBodies[3], 2.0f);
		bi.AddAngularImpulse(ioBodies[4], Vec3(0.0f, 300.0f, 50.0f));
		bi.AddImpulse(ioBodies[5], Vec3(0.0f, 900.0f, 200.0f), bi.GetCenterOfMassPosition(ioBodies[5]) + Vec3(0.2f, 0.0f, 0.1f));
		bi.SetLinearVelocity(ioBodies[6], Vec3(0.0f, 5.0f, 0.0f));
		bi.SetAngularVelocity(ioBodies[7], Vec3(1.0f, 2.0f, 3.0f));
		bi.SetMaxAngularVelocity(ioBodies[7], 1.5f);
		bi.AddForce(ioBodies[8], Vec3(0.0f, 30000.0f, 0.0f));
		bi.AddTorque(ioBodies[9], Vec3(500.0f, 0.0f, 0.0f));
	}
	else if (inPhase == 2)
	{
		// remove two bodies, destroy one of them, bulk add three new ones, re-add the other, move / deactivate
		BodyID removed[2] = { ioBodies[0], ioBodies[5] };
		bi.RemoveBodies(removed, 2);
		bi.DestroyBody(ioBodies[0]);
		BodyID added[3];
		for (int i = 0; i < 3; ++i)
		{
			BodyCreationSettings settings(B2J_NEW_SHAPE(BoxShape, Vec3(0.3f, 0.3f, 0.3f)), RVec3(-8.0f + 8.0f * float(i), 3.0f, 6.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
			added[i] = bi.CreateBody(settings)->GetID();
		}
		BodyInterface::AddState state = bi.AddBodiesPrepare(added, 3);
		bi.AddBodiesFinalize(added, 3, state, EActivation::Activate);
		for (int i = 0; i < 3; ++i) ioBodies.push_back(added[i]);
		ioBodies[0] = added[0]; // the destroyed id must not be used again
		bi.SetPositionAndRotation(ioBodies[5], RVec3(4.0f, 5.0f, -2.0f), Quat::sIdentity(), EActivation::DontActivate);
		bi.AddBody(ioBodies[5], EActivation::Activate);
		bi.SetPositionAndRotation(ioBodies[9], RVec3(3.0f, 4.0f, 0.0f), Quat::sIdentity(), EActivation::Activate);
		bi.DeactivateBody(ioBodies[2]);
	}
	else if (inPhase == 3)
	{
		// remove bodies while they move and put them back WITHOUT touching the pose: they return where they left the world, at rest
		// (BodyManager::DeactivateBodies reset the velocities of the active ones); a body created with a velocity keeps it
		bi.SetLinearVelocity(ioBodies[3], Vec3(1.0f, 2.0f, 0.0f));
		bi.SetAngularVelocity(ioBodies[6], Vec3(0.0f, 4.0f, 0.0f));
		BodyID moving[3] = { ioBodies[3], ioBodies[6], ioBodies[8] };
		bi.RemoveBodies(moving, 3);
		bi.AddBody(moving[0], EActivation::Activate);
		bi.AddBody(moving[1], EActivation::DontActivate);
		BodyCreationSettings settings(B2J_NEW_SHAPE(BoxShape, Vec3(0.4f, 0.2f, 0.3f)), RVec3(0.0f, 6.0f, 3.0f), Quat::sIdentity(), EMotionType::Dynamic, Layers::MOVING);
		settings.mLinearVelocity = Vec3(0.5f, 3.0f, 1.0f);
		settings.mAngularVelocity = Vec3(1.0f, 0.0f, 0.5f);
		Body *flying = bi.CreateBody(settings);
		bi.AddBody(flying->GetID(), EActivation::Activate);
		ioBodies.push_back(flying->GetID());
		bi.RemoveBody(flying->GetID());          // leaves at once, comes back with the velocity reset
		bi.AddBody(flying->GetID(), EActivation::Activate);
	}
	else if (inPhase == 4)
	{
		// SetMotionType / SetObjectLayer / SetShape / InvalidateContactCache on bodies that are in the world
		bi.SetMotionType(ioBodies[1], EMotionType::Kinematic, EActivation::Activate);      // a resting dynamic body becomes a kinematic mover
		bi.SetLinearVelocity(ioBodies[1], Vec3(0.8f, 0.0f, 0.0f));
		bi.SetMotionType(ioBodies[2], EMotionType::Static, EActivation::DontActivate);     // ... a static obstacle
		bi.SetMotionType(ioBodies[4], EMotionType::Kinematic, EActivation::Activate);
		bi.SetMotionType(ioBodies[4], EMotionType::Dynamic, EActivation::Activate);        // and back: mass properties survive
		bi.SetObjectLayer(ioBodies[7], Layers::NON_MOVING);                                // no longer collides with the floor: falls through
		bi.SetShape(ioBodies[9], B2J_NEW_SHAPE(SphereShape, 0.45f), true, EActivation::Activate);
		bi.SetShape(ioBodies[10], B2J_NEW_SHAPE(BoxShape, Vec3(0.6f, 0.2f, 0.4f)), false, EActivation::Activate);
		bi.InvalidateContactCache(ioBodies[11]);
	}
	else if (inPhase == 5)
	{
		// the rest of the pose / velocity surface (BodyInterface.h:187-230)
		bi.MoveKinematic(ioBodies[1], RVec3(2.0f, 1.5f, 1.0f), Quat(0.0f, 0.38268343f, 0.0f, 0.92387953f), 0.5f); // the kinematic body of phase 4
		bi.SetPositionRotationAndVelocity(ioBodies[4], RVec3(-2.0f, 3.0f, 1.0f), Quat(0.25881905f, 0.0f, 0.0f, 0.96592583f), Vec3(1.0f, 0.0f, 0.0f), Vec3(0.0f, 2.0f, 0.0f));
		bi.AddLinearVelocity(ioBodies[6], Vec3(0.0f, 3.0f, 0.0f));
		bi.AddLinearAndAngularVelocity(ioBodies[7], Vec3(0.5f, 2.0f, 0.0f), Vec3(0.0f, 0.0f, 1.5f));
		bi.AddForce(ioBodies[9], Vec3(0.0f, 20000.0f, 0.0f), bi.GetCenterOfMassPosition(ioBodies[9]) + Vec3(0.1f, 0.0f, 0.2f));
		bi.AddForceAndTorque(ioBodies[10], Vec3(5000.0f, 15000.0f, 0.0f), Vec3(0.0f, 300.0f, 0.0f));
		bi.SetPosition(ioBodies[11], RVec3(5.0f, 2.0f, 2.0f), EActivation::Activate);
		bi.SetRotation(ioBodies[3], Quat(0.0f, 0.0f, 0.38268343f, 0.92387953f), EActivation::Activate);
		bi.SetLinearVelocity(ioBodies[3], Vec3(0.0f, 900.0f, 0.0f)); // above the maximum of 2 set in phase 1: clamped
		bi.SetPositionAndRotationWhenChanged(ioBodies[5], bi.GetPosition(ioBodies[5]), bi.GetRotation(ioBodies[5]), EActivation::Activate); // unchanged: no activation
		// getters feed mutations, so a wrong value shows in the state
		Vec3 point_velocity = bi.GetPointVelocity(ioBodies[7], bi.GetCenterOfMassPosition(ioBodies[7]) + Vec3(0.3f, 0.1f, 0.0f));
		Vec3 lv, av;
		bi.GetLinearAndAngularVelocity(ioBodies[6], lv, av);
		bi.AddLinearVelocity(ioBodies[12], 0.1f * point_velocity + 0.1f * lv);
		auto inverse_inertia = bi.GetInverseInertia(ioBodies[4]);
		bi.AddAngularImpulse(ioBodies[4], 20.0f * inverse_inertia.GetColumn3(1));
		auto world_transform = bi.GetWorldTransform(ioBodies[13]);
		bi.SetPosition(ioBodies[13], world_transform.GetTranslation() + Vec3(0.0f, 0.5f, 0.0f), EActivation::Activate);
		auto com_transform = bi.GetCenterOfMassTransform(ioBodies[14]);
		bi.SetPosition(ioBodies[14], com_transform * Vec3(0.0f, 0.3f, 0.0f), EActivation::Activate);
		bi.SetUserData(ioBodies[2], 1234);
		bi.SetIsSensor(ioBodies[12], true);               // a resting dynamic body becomes a sensor: it falls through the floor
		bi.SetUseManifoldReduction(ioBodies[13], false);
		BodyID two[2] = { ioBodies[6], ioBodies[11] };
		bi.DeactivateBodies(two, 2);
		bi.ActivateBodies(two + 1, 1);
	}
}

// Queries after the tour: number of bodies, active bodies (as a sorted id list written to outIDs), returns the active count
static int sApiTourQuery(PhysicsSystem &inSystem, const std::vector<BodyID> &inBodies, uint32_t *outIDs, int inCapacity, uint32_t *outNumBodies, uint32_t *outFlags)
{
	BodyIDVector active;
	inSystem.GetActiveBodies(EBodyType::RigidBody, active);
	BodyIDVector all;
	inSystem.GetBodies(all);
	*outNumBodies = (uint32_t)all.size();
	BodyInterface &bi = inSystem.GetBodyInterface();
	uint32_t flags = 0;
	for (size_t i = 0; i < inBodies.size() && i < 32; ++i)
		if (bi.IsAdded(inBodies[i]) && bi.IsActive(inBodies[i])) flags |= 1u << i;
	*outFlags = flags;
	int n = 0;
	for (const BodyID &id : active) if (n < inCapacity) outIDs[n++] = id.GetIndexAndSequenceNumber();
	std::sort(outIDs, outIDs + n);
	return (int)active.size();
}
