// b2j_broadphase.h -- device broadphase: one linear BVH (Karras 2012) per broadphase layer over SoA body AABBs.
//
// Replaces BroadPhaseQuadTree / QuadTree (Jolt/Physics/Collision/BroadPhase/QuadTree.cpp): instead of widening a lock-free
// 4-ary tree on every move (NotifyBodiesAABBChanged :937-967) and rebuilding dirty sub-trees in a background job
// (UpdatePrepare :274-394), the tree of a layer whose bodies moved is rebuilt from scratch every step: Morton codes of the AABB
// centres -> radix sort -> parallel hierarchy -> bottom-up refit. The tree shape is NOT part of the state results depend on
// (SURVEY A.4): the pair predicate below is the reference's, evaluated on the true cached body AABBs, so candidate pair
// sets are identical:
//   QuadTree::FindCollidingPairs      QuadTree.cpp:1431-1537   (only the querying body's box is expanded by the speculative distance)
//   Body::sFindCollidingPairsCanCollide  Body.inl:30-79
//   BroadPhaseQuadTree::FindCollidingPairs  BroadPhaseQuadTree.cpp:563-600 (object vs broadphase layer filter per tree)
#pragma once

#include "b2j_world.h"

namespace b2j {

struct BodyPair { uint32_t a, b; }; // body slots: a = querying (active) body

// One LBVH. Node numbering: internal nodes [0, n-1), leaf i is node (n-1+i). n == 1: the root is leaf 0.
struct Tree
{
	uint32_t n;                  // number of bodies (leaves)
	uint32_t *bodies;            // [n] body slots in the layer (unsorted, maintained by the host)
	uint64_t *keys_in, *keys_out;   // [n] world index << 32 | 30 bit morton code
	int32_t *world_root;         // batched worlds: [num_worlds] node that spans exactly the bodies of one world (-1: none), else null
	uint32_t *leaf_body;         // [n] body slot per sorted leaf (sort output)
	int32_t *child_left, *child_right; // [n-1] node ids
	int32_t *parent;             // [2n-1]
	F4 *node_min, *node_max;     // [2n-1]
	uint32_t *visit;             // [n-1] refit arrival counters
	F4 *layer_bounds;            // [2] min, max of the whole layer as of the last refit (Morton normalisation for the next build)
};

B2J_HD uint32_t expand_bits_10(uint32_t v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

struct KMorton
{
	DWorld w; Tree t;
	B2J_D void operator()(uint32_t i) const
	{
		uint32_t body = t.bodies[i];
		V3 mn = to_v3(t.layer_bounds[0]), mx = to_v3(t.layer_bounds[1]);
		V3 c = 0.5f * (to_v3(w.bounds_min[body]) + to_v3(w.bounds_max[body]));
		V3 ext = mx - mn;
		float fx = ext.x > 0.0f? (c.x - mn.x) / ext.x : 0.0f;
		float fy = ext.y > 0.0f? (c.y - mn.y) / ext.y : 0.0f;
		float fz = ext.z > 0.0f? (c.z - mn.z) / ext.z : 0.0f;
		fx = fmin_(fmax_(fx * 1024.0f, 0.0f), 1023.0f);
		fy = fmin_(fmax_(fy * 1024.0f, 0.0f), 1023.0f);
		fz = fmin_(fmax_(fz * 1024.0f, 0.0f), 1023.0f);
		if (!(fx == fx)) fx = 0.0f;
		if (!(fy == fy)) fy = 0.0f;
		if (!(fz == fz)) fz = 0.0f;
		uint32_t morton = (expand_bits_10((uint32_t)fx) << 2) | (expand_bits_10((uint32_t)fy) << 1) | expand_bits_10((uint32_t)fz);
		t.keys_in[i] = ((uint64_t)world_of(w, body) << 32) | morton;
	}
};

// delta(i, j): common prefix length of the keys, ties broken by the index (keys are made unique), -1 out of range
B2J_D int lbvh_delta(const uint64_t *keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	uint64_t a = keys[i], b = keys[j];
	if (a != b) return clz64(a ^ b);
	return 64 + clz32((uint32_t)i ^ (uint32_t)j);
}

struct KBuildHierarchy
{
	Tree t;
	B2J_D void operator()(uint32_t idx) const
	{
		int n = (int)t.n, i = (int)idx;
		const uint64_t *keys = t.keys_out;
		int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0? 1 : -1;
		int delta_min = lbvh_delta(keys, n, i, i - d);
		int lmax = 2;
		while (lbvh_delta(keys, n, i, i + lmax * d) > delta_min) lmax *= 2;
		int l = 0;
		for (int t2 = lmax / 2; t2 >= 1; t2 /= 2)
			if (lbvh_delta(keys, n, i, i + (l + t2) * d) > delta_min) l += t2;
		int j = i + l * d;
		int delta_node = lbvh_delta(keys, n, i, j);
		int s = 0;
		int t2 = l;
		do
		{
			t2 = (t2 + 1) >> 1;
			if (lbvh_delta(keys, n, i, i + (s + t2) * d) > delta_node) s += t2;
		}
		while (t2 > 1);
		int gamma = i + s * d + (d < 0? -1 : 0);
		int lo = i < j? i : j, hi = i < j? j : i;
		int left = (lo == gamma)? (n - 1 + gamma) : gamma;
		int right = (hi == gamma + 1)? (n - 1 + gamma + 1) : (gamma + 1);
		t.child_left[i] = left;
		t.child_right[i] = right;
		t.parent[left] = i;
		t.parent[right] = i;
		if (i == 0) t.parent[0] = -1;
		t.visit[i] = 0;
		if (t.world_root != nullptr)
		{
			// the node whose key range is exactly one world becomes the traversal root of that world
			uint32_t wlo = (uint32_t)(keys[lo] >> 32), whi = (uint32_t)(keys[hi] >> 32);
			if (wlo == whi && (lo == 0 || (uint32_t)(keys[lo - 1] >> 32) != wlo) && (hi == n - 1 || (uint32_t)(keys[hi + 1] >> 32) != wlo))
				t.world_root[wlo] = i;
		}
	}
};

struct KRefit
{
	DWorld w; Tree t;
	B2J_D void operator()(uint32_t i) const
	{
		int n = (int)t.n;
		uint32_t body = t.leaf_body[i];
		int node = n - 1 + (int)i;
		F4 mn = w.bounds_min[body], mx = w.bounds_max[body];
		t.node_min[node] = mn;
		t.node_max[node] = mx;
		if (t.world_root != nullptr)
		{
			// a world with a single body in this layer: the leaf is its root
			uint32_t wi = (uint32_t)(t.keys_out[i] >> 32);
			if ((i == 0 || (uint32_t)(t.keys_out[i - 1] >> 32) != wi) && ((int)i == n - 1 || (uint32_t)(t.keys_out[i + 1] >> 32) != wi))
				t.world_root[wi] = node;
		}
		if (n == 1)
		{
			t.layer_bounds[0] = mn; t.layer_bounds[1] = mx;
			return;
		}
		mem_fence();
		int p = t.parent[node];
		while (p >= 0)
		{
			if (atomic_add(&t.visit[p], 1u) == 0)
				return; // first arrival: the sibling subtree is not done yet
			mem_fence();
			int l = t.child_left[p], r = t.child_right[p];
			F4 lmn = t.node_min[l], lmx = t.node_max[l], rmn = t.node_min[r], rmx = t.node_max[r];
			F4 nmn = f4(fmin_(lmn.x, rmn.x), fmin_(lmn.y, rmn.y), fmin_(lmn.z, rmn.z), 0.0f);
			F4 nmx = f4(fmax_(lmx.x, rmx.x), fmax_(lmx.y, rmx.y), fmax_(lmx.z, rmx.z), 0.0f);
			t.node_min[p] = nmn;
			t.node_max[p] = nmx;
			if (p == 0) { t.layer_bounds[0] = nmn; t.layer_bounds[1] = nmx; }
			mem_fence();
			p = t.parent[p];
		}
	}
};

// AABox::Overlaps (AABox.h:164-167): !(any(min1 > max2) || any(max1 < min2))
B2J_HD bool aabb_overlaps(V3 min1, V3 max1, V3 min2, V3 max2)
{
	return !(min1.x > max2.x || min1.y > max2.y || min1.z > max2.z || max1.x < min2.x || max1.y < min2.y || max1.z < min2.z);
}

// Body::sFindCollidingPairsCanCollide (Body.inl:30-79); group filters are not supported (default null filter -> always collide)
B2J_HD bool can_collide_pair(const BodyInfo &b1, uint32_t active_index1, const BodyInfo &b2, uint32_t active_index2)
{
	bool dyn1 = b1.motion_type == B2J_MOTION_DYNAMIC, dyn2 = b2.motion_type == B2J_MOTION_DYNAMIC;
	bool kin1 = b1.motion_type == B2J_MOTION_KINEMATIC, kin2 = b2.motion_type == B2J_MOTION_KINEMATIC;
	if (!(b1.flags & B2J_BODY_KIN_VS_NONDYN) && !(b2.flags & B2J_BODY_KIN_VS_NONDYN)
		&& (!dyn1 && !dyn2)
		&& !(kin1 && (b2.flags & B2J_BODY_SENSOR))
		&& !(kin2 && (b1.flags & B2J_BODY_SENSOR)))
		return false;
	if (active_index1 >= active_index2)
		return false;
	return true;
}

struct KFindPairs
{
	DWorld w;
	Tree trees[8];
	BodyPair *pairs;
	uint32_t first; // first active index to query (bodies woken mid-step are queried in a later round)
	// First round: the queries run in the (Morton sorted) LEAF order of one layer's tree instead of active list order, so that the
	// lanes of a warp query neighbouring boxes and walk nearly the same path (inactive leaves idle). Null = active list order.
	const uint32_t *query_leaves;
	B2J_D void operator()(uint32_t k) const
	{
		uint32_t ai, b1;
		if (query_leaves != nullptr)
		{
			b1 = query_leaves[k];
			ai = w.active_index[b1];
			if (ai == B2J_INACTIVE_INDEX)
				return;
		}
		else
		{
			ai = first + k;
			b1 = w.active[ai];
		}
		BodyInfo i1 = w.info[b1];
		float sd = w.settings.speculative_contact_distance;
		V3 min1 = to_v3(w.bounds_min[b1]) - v3_rep(sd);
		V3 max1 = to_v3(w.bounds_max[b1]) + v3_rep(sd);
		for (uint32_t l = 0; l < w.num_bp_layers; ++l)
		{
			const Tree &t = trees[l];
			if (t.n == 0 || !w.object_vs_bp[i1.object_layer * w.num_bp_layers + l])
				continue;
			int n = (int)t.n;
			int stack[128];
			int top = 0;
			stack[0] = 0; // the root (for n == 1 node 0 is the only leaf)
			if (t.world_root != nullptr)
			{
				int wr = t.world_root[world_of(w, b1)];
				if (wr < 0)
					continue; // this world has no body in the layer
				stack[0] = wr;
			}
			while (top >= 0)
			{
				int node = stack[top--];
				if (node >= n - 1)
				{
					uint32_t b2 = t.leaf_body[node - (n - 1)];
					if (b2 != b1)
					{
						BodyInfo i2 = w.info[b2];
						if (w.object_vs_object[i1.object_layer * w.num_object_layers + i2.object_layer]
							&& can_collide_pair(i1, ai, i2, w.active_index[b2])
							&& aabb_overlaps(min1, max1, to_v3(w.bounds_min[b2]), to_v3(w.bounds_max[b2])))
						{
							uint32_t idx = atomic_add(&w.counters->num_pairs, 1u);
							if (idx < w.max_body_pairs)
							{
								BodyPair p; p.a = b1; p.b = b2;
								pairs[idx] = p;
							}
						}
					}
				}
				else
				{
					int l2 = t.child_left[node], r2 = t.child_right[node];
					bool ol = aabb_overlaps(min1, max1, to_v3(t.node_min[l2]), to_v3(t.node_max[l2]));
					bool orr = aabb_overlaps(min1, max1, to_v3(t.node_min[r2]), to_v3(t.node_max[r2]));
					if (ol && top < 126) stack[++top] = l2;
					if (orr && top < 126) stack[++top] = r2;
				}
			}
		}
	}
};

} // namespace b2j
