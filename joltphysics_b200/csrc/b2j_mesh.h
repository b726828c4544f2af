// b2j_mesh.h -- convex vs static MeshShape (placeholder until the tree walk lands; mesh pairs produce no contacts yet).
#pragma once

#include "b2j_narrowphase.h"

namespace b2j {

struct KCollideMesh
{
	DWorld w; NarrowCtx c;
	B2J_D void run(uint32_t, uint32_t) const { }
};

} // namespace b2j
