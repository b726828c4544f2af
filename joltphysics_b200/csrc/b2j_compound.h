// b2j_compound.h -- body pairs with a StaticCompoundShape.
//
// Restates StaticCompoundShape::sCollideCompoundVsShape / sCollideShapeVsCompound (StaticCompoundShape.cpp:590-644), WalkTree
// (:359-421), the visitors of CompoundShapeVisitors.h:283-430 and, for one (convex, convex) leaf pair, ConvexShape::sCollideConvexVsConvex
// (ConvexShape.cpp:45-164). Sub shapes are convex shapes (plain or decorated); the other body's shape may be convex, a mesh or another
// compound. Every leaf hit joins the pair's manifolds through the collector the mesh path already has (mesh_add_hit =
// ReductionCollideShapeCollector::AddHit), in the reference's visit order, with the sub shape ids of both sides.
// One lane per pair (the caller is lane 0 of a warp with an EPA hull in shared memory): compounds are a widening step, not yet a
// tuned kernel.
#pragma once

#include "b2j_mesh.h"

namespace b2j {

// AABox::Transformed (AABox.h:193-213)
B2J_D void aabox_transformed(const Xf &m, V3 mn, V3 mx, V3 &out_min, V3 &out_max)
{
	V3 new_min = m.t, new_max = m.t;
	for (int c = 0; c < 3; ++c)
	{
		V3 col = m33_col(m.r, c);
		V3 a = col * v3_get(mn, c), b = col * v3_get(mx, c);
		new_min += v3_min(a, b);
		new_max += v3_max(a, b);
	}
	out_min = new_min; out_max = new_max;
}

B2J_D Xf xf_inversed(const Xf &m) { M33 rt = transposed(m.r); return xf(rt, -mul(rt, m.t)); } // Mat44::InversedRotationTranslation

// SubShapeIDCreator::PushID on an id that has used `first_bit` bits so far
B2J_D uint32_t sub_shape_push(uint32_t id, uint32_t first_bit, uint32_t value, uint32_t bits)
{
	if (bits == 0) return id;
	uint32_t mask = bits >= 32? 0xffffffffu : ((1u << bits) - 1u);
	return (id & ~(mask << first_bit)) | (value << first_bit);
}

// One side of a leaf test: a convex shape description with the centre of mass transform the dispatch has reached for it
struct CompoundSide { const ShapeDesc *shape; Xf transform; uint32_t sub; };

struct CompoundPairCtx
{
	float max_separation_distance;
	V3 movement_direction;
	EpaScratch *epa;
	MeshScratch *ms;
	int num_manifolds;
	QueryCollector *query;       // not null: a CollideShape query (b2j_query.h) -- hits go to the collector instead of the manifolds
};

// ConvexShape::sCollideConvexVsConvex for one leaf pair; the hit (if any) joins the collector
B2J_D void compound_collide_leaf(const DWorld &w, CompoundPairCtx &p, const CompoundSide &a, const CompoundSide &b)
{
	const ShapeDesc &s1 = *a.shape, &s2 = *b.shape;
	Xf transform1 = shape_transform(s1, a.transform), transform2 = shape_transform(s2, b.transform);
	Xf transform_2_to_1 = mul(xf_inversed(transform1), transform2);
	float max_separation_distance = p.max_separation_distance;
	V3 bb1_min = s1.local_min - v3_rep(max_separation_distance), bb1_max = s1.local_max + v3_rep(max_separation_distance);
	if (!obb_vs_aabb(transform_2_to_1, s2.local_min, s2.local_max, bb1_min, bb1_max))
		return;
	ConvexSupport a_excl = make_support(w, s1, SUPPORT_EXCLUDE_CONVEX_RADIUS);
	TransformedSupport b_excl = make_transformed(transform_2_to_1, make_support(w, s2, SUPPORT_EXCLUDE_CONVEX_RADIUS));
	V3 penetration_axis = transform_2_to_1.t, point1 = v3_zero(), point2 = v3_zero();
	if (is_near_zero(penetration_axis))
		penetration_axis = v3(1.0f, 0.0f, 0.0f);
	GjkSimplex simplex;
	int status = pen_depth_step_gjk(simplex, a_excl, a_excl.convex_radius + max_separation_distance, b_excl, b_excl.s.convex_radius, 1.0e-4f, penetration_axis, point1, point2);
	if (status == PEN_NOT_COLLIDING)
		return;
	if (status == PEN_INDETERMINATE)
	{
		max_separation_distance = fmin_(max_separation_distance, 1.0f);
		AddRadiusSupport a_incl; a_incl.s = make_support(w, s1, SUPPORT_INCLUDE_CONVEX_RADIUS); a_incl.radius = max_separation_distance;
		TransformedSupport b_incl = make_transformed(transform_2_to_1, make_support(w, s2, SUPPORT_INCLUDE_CONVEX_RADIUS));
		if (!pen_depth_step_epa(*p.epa, simplex, a_incl, b_incl, 1.0e-4f, penetration_axis, point1, point2))
			return;
	}
	float penetration_depth = length(point2 - point1) - max_separation_distance;
	if (-penetration_depth >= FLT_MAX)
		return;
	float penetration_axis_len = length(penetration_axis);
	if (penetration_axis_len > 0.0f)
		point1 -= penetration_axis * (max_separation_distance / penetration_axis_len);
	point1 = mul(transform1, point1);
	point2 = mul(transform1, point2);
	V3 axis_world = mul(transform1.r, penetration_axis);
	if (p.query != nullptr) { query_add_hit(*p.query, point1, point2, axis_world, penetration_depth, a.sub, b.sub); return; }
	V3 face1[MAX_FACE_VERTS], face2[MAX_FACE_VERTS];
	int n1 = supporting_face(w, s1, -penetration_axis, transform1, face1);
	int n2 = supporting_face(w, s2, mul_transposed(transform_2_to_1.r, penetration_axis), transform2, face2);
	mesh_add_hit(w, *p.ms, p.num_manifolds, point1, point2, axis_world, penetration_depth, a.sub, b.sub, face1, n1, face2, n2);
}

// (convex shape a) against (shape b that is not a compound): the leaf test, or the mesh walk
B2J_D void compound_collide_simple(const DWorld &w, CompoundPairCtx &p, const CompoundSide &a, const CompoundSide &b)
{
	if (b.shape->kind == B2J_SHAPE_MESH)
	{
		MeshCollideCtx cc = mesh_collide_ctx_from(w, *a.shape, a.transform, *b.shape, b.transform, p.max_separation_distance, p.movement_direction, a.sub);
		if (p.query != nullptr) { cc.query = p.query; cc.check_active_edges = true; } // CollideShapeSettings: EActiveEdgeMode::CollideOnlyWithActive
		mesh_walk_serial(w, *a.shape, *b.shape, cc, *p.epa, *p.ms, p.num_manifolds);
	}
	else
		compound_collide_leaf(w, p, a, b);
}

// StaticCompoundShape::WalkTree with the box test both visitors use (AABox4Scale with unit scale, AABox4VsBox): calls
// visit(sub shape index) for every sub shape whose bounds overlap [bmin, bmax] (in the compound's space), in the reference's order
template <class Visit> B2J_D void compound_walk_tree(const DWorld &w, const ShapeDesc &compound, V3 bmin, V3 bmax, Visit visit)
{
	const uint8_t *nodes = w.mesh_bytes + compound.mesh_offset;
	uint32_t stack[128];
	int top = 0;
	stack[0] = 0;
	do
	{
		uint32_t node_properties = stack[top];
		if (node_properties != 0x7fffffffu) // INVALID_NODE
		{
			if ((node_properties & 0x80000000u) == 0)
			{
				const uint8_t *node = nodes + (size_t)node_properties * 64;
				uint32_t props[4];
				int n = 0;
				for (int ch = 0; ch < 4; ++ch)
				{
					float mnx = half_to_float(load_u16(node + 0 + 2 * ch)), mny = half_to_float(load_u16(node + 8 + 2 * ch)), mnz = half_to_float(load_u16(node + 16 + 2 * ch));
					float mxx = half_to_float(load_u16(node + 24 + 2 * ch)), mxy = half_to_float(load_u16(node + 32 + 2 * ch)), mxz = half_to_float(load_u16(node + 40 + 2 * ch));
					if (!((bmin.x > mxx || mnx > bmax.x) || (bmin.y > mxy || mny > bmax.y) || (bmin.z > mxz || mnz > bmax.z)))
						props[n++] = load_u32(node + 48 + 4 * ch);
				}
				for (int j = 0; j < n && top + j < 128; ++j) stack[top + j] = props[j];
				top += n;
			}
			else
				visit(node_properties ^ 0x80000000u);
		}
		--top;
	}
	while (top >= 0);
}

// sCollideCompoundVsShape: (compound a) against (shape b that is not a compound): the sub shapes of a whose bounds overlap b
B2J_D void compound_collide_compound_vs(const DWorld &w, CompoundPairCtx &p, const CompoundSide &a, const CompoundSide &b)
{
	// CollideCompoundVsShapeVisitor: bounds of shape 2 (its own GetLocalBounds, decorators included) in the space of the compound, expanded
	Xf transform2_to_1 = mul(xf_inversed(a.transform), b.transform);
	V3 bmin, bmax;
	aabox_transformed(transform2_to_1, b.shape->outer_min, b.shape->outer_max, bmin, bmax);
	bmin = bmin - v3_rep(p.max_separation_distance); bmax = bmax + v3_rep(p.max_separation_distance);
	const ShapeDesc &compound = *a.shape;
	compound_walk_tree(w, compound, bmin, bmax, [&](uint32_t index) {
		const CompoundSub &sub = w.compound_subs[compound.compound_sub_offset + index];
		CompoundSide side;
		side.shape = &w.shapes[sub.shape];
		side.transform = mul(a.transform, compound_sub_transform(sub));
		side.sub = sub_shape_push(a.sub, 0, index, compound.compound_sub_bits);
		compound_collide_simple(w, p, side, b);
	});
}

// CollisionDispatch::sCollideShapeVsShape for (a, b) where a compound may be on either side (see compound_collide_pair)
B2J_D void compound_dispatch(const DWorld &w, CompoundPairCtx &p, const CompoundSide &a, const CompoundSide &b)
{
	const ShapeDesc &s1 = *a.shape, &s2 = *b.shape;
	if (s2.kind == B2J_SHAPE_COMPOUND)
	{
		// CollideShapeVsCompoundVisitor: bounds of shape 1 in the space of the compound, expanded
		Xf transform1_to_2 = mul(xf_inversed(b.transform), a.transform);
		V3 bmin, bmax;
		aabox_transformed(transform1_to_2, s1.outer_min, s1.outer_max, bmin, bmax);
		bmin = bmin - v3_rep(p.max_separation_distance); bmax = bmax + v3_rep(p.max_separation_distance);
		compound_walk_tree(w, s2, bmin, bmax, [&](uint32_t index) {
			const CompoundSub &sub = w.compound_subs[s2.compound_sub_offset + index];
			CompoundSide side;
			side.shape = &w.shapes[sub.shape];
			side.transform = mul(b.transform, compound_sub_transform(sub));
			side.sub = sub_shape_push(b.sub, 0, index, s2.compound_sub_bits);
			if (s1.kind == B2J_SHAPE_COMPOUND)
				compound_collide_compound_vs(w, p, a, side);
			else
				compound_collide_simple(w, p, a, side);
		});
	}
	else if (s1.kind == B2J_SHAPE_COMPOUND)
		compound_collide_compound_vs(w, p, a, b); // (shape 1 is the compound; shape 2 convex or a mesh)
	else
		compound_collide_simple(w, p, a, b);
}

// PhysicsSystem::ProcessBodyPair for a pair in which at least one body's shape is a StaticCompoundShape. The dispatch table holds
// sCollideShapeVsCompound for (anything, compound) -- also for (compound, compound): StaticCompoundShape::sRegister writes that entry
// last -- and sCollideCompoundVsShape for (compound, anything else). So the OUTER loop runs over the sub shapes of shape 2 when that
// is a compound, and a compound shape 1 is then walked once per sub shape of shape 2.
B2J_D void compound_collide_pair(const DWorld &w, const NarrowCtx &c, const CollideItem &item, EpaScratch &epa, MeshScratch &ms)
{
	BodyInfo i1 = w.info[item.b1], i2 = w.info[item.b2];
	const ShapeDesc &s1 = w.shapes[i1.shape], &s2 = w.shapes[i2.shape];
	V3 x1 = to_v3(w.position[item.b1]), x2 = to_v3(w.position[item.b2]);
	CompoundPairCtx p;
	p.max_separation_distance = ((i1.flags | i2.flags) & B2J_BODY_SENSOR)? 0.0f : w.settings.speculative_contact_distance;
	p.movement_direction = pair_movement_direction(w, item, i1, i2);
	p.epa = &epa; p.ms = &ms; p.num_manifolds = 0; p.query = nullptr;
	CompoundSide a, b;
	a.shape = &s1; a.transform = xf(m33_rotation(to_q4(w.rotation[item.b1])), v3_zero()); a.sub = 0xffffffffu;
	b.shape = &s2; b.transform = xf(m33_rotation(to_q4(w.rotation[item.b2])), x2 + (-x1)); b.sub = 0xffffffffu;
	compound_dispatch(w, p, a, b);
	mesh_finish_pair(w, c, item, ms, p.num_manifolds);
}

} // namespace b2j
