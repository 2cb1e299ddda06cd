// traverse.cuh — closest-hit traversal, bit-exact against the reference.
//
// The reference walks its binary BVH with a 64-entry stack (reference Raytracing/BVH.cpp:122-191): slab-test BOTH
// children of an inner node against the current maxRayLength (Math::IntersectRayAABB, Bounds.cpp:582-604), descend
// into the nearer one (ties keep child1), push the farther one if it was hit, and in a leaf test every triangle in
// stored order (Moller-Trumbore, GLM variant with zero epsilon, Bounds.h:193-259 + Bounds.cpp:509-534), keeping a
// hit only when distance < maxRayLength (strict), which then shrinks.  Which triangle wins a bit-equal distance
// therefore depends on visit order; this kernel keeps the reference's topology AND visit order, so the winner, t, u
// and v are identical by construction (no tie post-pass needed).
//
// Layout for the device (built by bvh_pack): instead of the reference's 32-byte node whose two children live in two
// other cache lines, every INNER node carries both child boxes + both child references in one 64-byte record =
// four LDG.128, one 64-B sector pair per visit; leaves hold (v0, e1 = v1-v0, e2 = v2-v0, original index) as three
// LDG.128 in leaf order (the subtractions are the same IEEE operations the reference performs per test).
// Child reference: bit31 set -> leaf, low bits = first triangle slot (its count sits in that triangle's record);
// clear -> inner node index (pre-order, so the near-left subtree is adjacent in memory).
#pragma once
#include "backend.h"

namespace spt
{
	struct alignas(16) TNode { V4 q0, q1, q2; uint32_t left, right, pad0, pad1; };
	struct alignas(16) TTri { V4 a, b, c; };   // a=(v0, e1.x) b=(e1.y,e1.z,e2.x,e2.y) c=(e2.z, bits(triId), bits(leafCount), bits(lastInLeaf))
	static_assert(sizeof(TNode) == 64 && sizeof(TTri) == 48, "traversal layout");

	constexpr uint32_t kLeafBit = 0x80000000u;
	constexpr uint32_t kNoHit = 0xFFFFFFFFu;
	constexpr int kStackDepth = 64;               // BVH.cpp:126

	struct BvhView { const TNode* nodes; const TTri* tris; uint32_t rootRef; uint32_t numNodes, numTris; };
	struct Hit { float t, u, v; uint32_t tri; };
	// A closest-hit query of a wavefront level that hit something, as ClassifyKernel needs it: the hit, the ray it belongs to and the ray's
	// level-local index.  Written by the trace kernels at consecutive positions (one warp-aggregated counter), so classify streams 48-byte
	// records instead of chasing a list of indices into the ray queue and the hit array (two scattered sectors per entry).
	struct alignas(16) SlowRec { float t, u, v; uint32_t tri; float ox, oy, oz; uint32_t index; float dx, dy, dz; uint32_t pad; };

	// Math::IntersectRayAABB (Bounds.cpp:582-604). _mm_max_ps/_mm_min_ps return the SECOND operand on NaN.
	SPT_HD float SlabTest(V3 o, V3 rD, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float maxLen)
	{
		const float t1x = (bminx - o.x) * rD.x, t1y = (bminy - o.y) * rD.y, t1z = (bminz - o.z) * rD.z;
		const float t2x = (bmaxx - o.x) * rD.x, t2y = (bmaxy - o.y) * rD.y, t2z = (bmaxz - o.z) * rD.z;
		const float vmaxx = sse_max(t1x, t2x), vmaxy = sse_max(t1y, t2y), vmaxz = sse_max(t1z, t2z);
		const float vminx = sse_min(t1x, t2x), vminy = sse_min(t1y, t2y), vminz = sse_min(t1z, t2z);
		const float tmax = glm_min(vmaxx, glm_min(vmaxy, vmaxz));
		const float tmin = glm_max(vminx, glm_max(vminy, vminz));
		if (tmax >= tmin && tmin < maxLen && tmax > 0.0f) return tmin;
		return kFltMax;
	}

	// GLM-variant Moller-Trumbore with zero epsilon (Bounds.h:193-259) + the acceptance window of Bounds.cpp:521.
	SPT_HD bool TriTest(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float maxLen, float& outT, float& outU, float& outV)
	{
		const V3 p = cross(d, e2);
		const float det = dot(e1, p);
		float u, v;
		V3 perp;
		// The two signed branches of the GLM routine (Bounds.h:207-243) compute the same u, perp and v and differ only in the
		// comparisons, so they are evaluated once and the reference's comparisons are selected by the sign of det (NaNs
		// behave as in the reference: every comparison with a NaN is false).  A warp whose lanes see both signs issues the
		// arithmetic once: -7 % traversal time on the 1M-triangle scene (profiles/r01b_SUMMARY.md).
		{
			const V3 dist = o - v0;
			u = dot(dist, p);
			perp = cross(dist, e1);
			v = dot(d, perp);
			const float uv = u + v;
			const bool pos = (det > 0.0f) & !((u < 0.0f) | (u > det)) & !((v < 0.0f) | (uv > det));
			const bool neg = (det < 0.0f) & !((u > 0.0f) | (u < det)) & !((v > 0.0f) | (uv < det));
			if (!(pos | neg)) return false;
		}
		const float invDet = 1.0f / det;
		const float t = dot(e2, perp) * invDet;
		u *= invDet; v *= invDet;
		if (!(t < maxLen && t > -0.0000001f)) return false;
		// RaycastHit::HasIntersection (Bounds.h:63): the hit point must not be (+inf,+inf,+inf)
		const V3 pt = o + d * t;
		const float inf = u2f(0x7F800000u);
		if (!(pt.x != inf || pt.y != inf || pt.z != inf)) return false;
		outT = t; outU = u; outV = v;
		return true;
	}

	// BVH::IntersectBVH (BVH.cpp:122-191). `Stack` provides push/pop/empty (shared memory on the device).
	template<class Stack>
	SPT_HD bool TraceClosest(const BvhView& bvh, V3 o, V3 d, uint32_t ignoreTri, float maxLen, Stack& stack, Hit& hit,
		uint32_t* boxTests = nullptr, uint32_t* triTests = nullptr)
	{
		const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);      // Ray::SetDirection (Bounds.h:44-48)
		hit.t = u2f(0x7F800000u); hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
		uint32_t cur = bvh.rootRef;
		stack.clear();
		for (;;)
		{
			if (cur & kLeafBit)
			{
				const uint32_t first = cur & ~kLeafBit;
				const TTri* T = bvh.tris + first;
				const uint32_t count = f2u(ld4(&T[0].c).z);
				for (uint32_t i = 0; i < count; i++)
				{
					const V4 a = ld4(&T[i].a), b = ld4(&T[i].b), c = ld4(&T[i].c);
					const uint32_t triId = f2u(c.y);
					if (ignoreTri == triId) continue;
					if (triTests) (*triTests)++;
					float t, u, v;
					if (TriTest(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), maxLen, t, u, v))
					{
						hit.t = t; hit.u = u; hit.v = v; hit.tri = triId;
						maxLen = std_min(maxLen, t);
					}
				}
				if (stack.empty()) break;
				cur = stack.pop();
				continue;
			}
			const TNode* n = bvh.nodes + cur;
			const V4 q0 = ld4(&n->q0), q1 = ld4(&n->q1), q2 = ld4(&n->q2), q3 = ld4(reinterpret_cast<const V4*>(&n->left));
			uint32_t c1 = f2u(q3.x), c2 = f2u(q3.y);
			if (boxTests) (*boxTests) += 2;
			float d1 = SlabTest(o, rD, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, maxLen);
			float d2 = SlabTest(o, rD, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, maxLen);
			if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const uint32_t tc = c1; c1 = c2; c2 = tc; }
			if (d1 == kFltMax)
			{
				if (stack.empty()) break;
				cur = stack.pop();
			}
			else
			{
				cur = c1;
				if (d2 != kFltMax) stack.push(c2);
			}
		}
		return hit.tri != kNoHit;
	}

	struct LocalStack
	{
		uint32_t e[kStackDepth]; int n;
		SPT_HD void clear() { n = 0; }
		SPT_HD bool empty() const { return n == 0; }
		SPT_HD void push(uint32_t v) { e[n++] = v; }
		SPT_HD uint32_t pop() { return e[--n]; }
	};

	// ---- layout packing (device) ------------------------------------------------------------------------------
	struct PackNodesKernel   // one thread per build node
	{
		const uint32_t* left; const uint32_t* rank; const uint32_t* refIdx; const uint32_t* leafOffsetByRef; const float* aabb;
		TNode* out;
		SPT_HD uint32_t RefOf(uint32_t node) const { return left[node] ? rank[node] : (kLeafBit | leafOffsetByRef[refIdx[node]]); }
		SPT_KERNEL_BODY void operator()(uint32_t node) const
		{
			const uint32_t l = left[node];
			if (!l) return;
			const float* L = aabb + (size_t)l * 6; const float* R = aabb + (size_t)(l + 1) * 6;
			TNode n;
			n.q0 = v4(L[0], L[1], L[2], L[3]); n.q1 = v4(L[4], L[5], R[0], R[1]); n.q2 = v4(R[2], R[3], R[4], R[5]);
			n.left = RefOf(l); n.right = RefOf(l + 1); n.pad0 = n.pad1 = 0;
			out[rank[node]] = n;
		}
	};

	struct PackTrisKernel    // one thread per triangle slot in leaf order
	{
		const V4* vtx; const uint32_t* mapping; const uint32_t* leafCountAtSlot; TTri* out; uint32_t numTris;
		SPT_KERNEL_BODY void operator()(uint32_t slot) const
		{
			const uint32_t tri = mapping[slot];
			const V4 v0 = vtx[tri * 3], v1 = vtx[tri * 3 + 1], v2 = vtx[tri * 3 + 2];
			const V3 e1 = v3(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z), e2 = v3(v2.x - v0.x, v2.y - v0.y, v2.z - v0.z);
			TTri t;
			t.a = v4(v0.x, v0.y, v0.z, e1.x); t.b = v4(e1.y, e1.z, e2.x, e2.y);
			// c.z: size of the leaf (at its first slot, else 0); c.w: 1 on the last triangle of a leaf (the next slot starts a leaf)
			const uint32_t last = (slot + 1u == numTris || leafCountAtSlot[slot + 1u] != 0u) ? 1u : 0u;
			t.c = v4(e2.z, u2f(tri), u2f(leafCountAtSlot[slot]), u2f(last));
			out[slot] = t;
		}
	};

	struct LeafCountAtSlotKernel   // one thread per build node: write the leaf's size at its first slot
	{
		const uint32_t* left; const uint32_t* count; const uint32_t* refIdx; const uint32_t* leafOffsetByRef; uint32_t* leafCountAtSlot;
		SPT_KERNEL_BODY void operator()(uint32_t node) const { if (!left[node]) leafCountAtSlot[leafOffsetByRef[refIdx[node]]] = count[node]; }
	};

	// ---- ray generation (PathTracer.cpp:458-466) ------------------------------------------------------------
	struct CameraGpu { V3 pos, pixel00Dir, deltaU, deltaV; uint32_t width, height; };

	SPT_HD V3 PrimaryDir(const CameraGpu& c, uint32_t x, uint32_t y, float ox, float oy)
	{
		const V3 pixelDir = c.pixel00Dir + ((float)x + ox) * c.deltaU + ((float)y - oy) * c.deltaV;
		return normalize(pixelDir);
	}
}
