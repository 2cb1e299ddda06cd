// trace_fast.cuh — ORIGIN-LOCAL traversal of the reference's own binary tree for secondary rays, plus the exact replay pass.
//
// WHY.  A secondary ray starts ON a triangle (its `ignore` index) and, in the scenes this renderer is quoted on, usually ends
// within a few leaves of it.  The reference's top-down walk (BVH.cpp:122-191, traverse.cuh) nevertheless descends the whole
// chain of ~20 boxes that contain the origin before it tests the first nearby triangle: two thirds of its inner-node visits
// (profiles/r01g_SUMMARY.md).  Every query whose result does not depend on the visit order can start at the other end:
//
//   start     in the leaf that holds the ray's own triangle (triStart): its triangles are queued first, nearest geometry first;
//   UP        TWO levels up per step: the climb node of the subtree just finished (FastClimbKernel) holds its sibling's box, its parent's
//             sibling's box and the grandparent's link, laid out like any other node, so an UP step is a DOWN step on that record;
//   DOWN      the ordinary stack walk of the subtrees that were hit, then UP again until the root has been passed.
// Every subtree is visited at most once and none is skipped unless its box fails a slab test that is CONSERVATIVE with respect
// to the reference's own (see FastSlab), so the triangles tested are a superset of those the reference can reach (its slab
// distances are monotone under box nesting: a leaf whose box passes is reached whatever its ancestors were).  What changes is
// the ORDER, and the results are made independent of it:
//   * hit-or-miss queries (shadow rays, hemisphere rays, importance rays below the throughput cut) stop at the first accepted
//     triangle whose own leaf box passes the reference's slab test — "some reachable triangle is hit" is order independent;
//   * closest-hit queries keep the nearest candidate.  t, u, v come from the unchanged TriTest; only the WINNER among (nearly)
//     equal distances can depend on the order, so a second candidate within a relative band of 2^-16 of the best raises `tie`,
//     and a winner whose leaf box the reference would have culled at that ray length fails the winner check: both kinds of ray,
//     rays with a non-finite (or huge) reciprocal direction and walks that overflow the short stack are appended to the REPLAY
//     list and traced afterwards by the exact kernel (reference visit order) — SURVEY H1's recipe.
// The primary pass keeps the exact kernel (north_star: bit-exact primary hits; camera rays have no origin triangle anyway).
//
// WHAT BOUNDS IT ON A B200 (tools/micro/gather_bench.cu, profiles/r02_SUMMARY.md).  A warp whose 32 lanes fetch 32 different
// records pays the L1TEX pipe per (instruction, 128-byte line): 92 SM cycles for a 64-byte node as 4 x LDG.128, 71 as
// 2 x LDG.256, 40 for ONE 32-byte LDG.256 — and an SM issues 4 warp instructions per cycle, so every variant of the round-1
// kernel that read 64-byte nodes with 128-bit loads ran at 80 % L1TEX utilisation however its steps were scheduled.  Hence:
//   FNode   64 B = two 32-byte halves {child box (fp32, padded outwards), child reference, parent link}, read as 2 x LDG.256.  The climb
//           nodes are a second array of the same records (two per inner node).  Round 2 first climbed one level per step (one half of
//           the parent, 1 x LDG.256); two levels per step cost the same 64 bytes per two levels and half the steps: C3 traversal
//           81.8 -> 74.5 ms (profiles/r03b_two_level_climb_ab.txt).
//   FTri    64 B = {v0, e1, e2} in the first 32 + 8 bytes (LDG.256 + LDG.64 per test), the exact box of its leaf behind them
//           (read once, for the winner).
//   slab    6 FFMA (b * 1/d - o/d) + 3 FMNMX + 3 FMNMX + 2 FMNMX3-pairs per box instead of the reference's exact
//           subtract-multiply-compare chain: the boxes are padded by 2^-18 of the scene's largest coordinate, which covers the
//           difference between the two roundings (FastSlab), so nothing the reference reaches is missed.
//   lanes   every lane keeps NODE work (current node, stack, climb link) and TRIANGLE work (current leaf position, a short
//           queue of leaves) apart: leaves found by a node step are queued, and the warp runs a triangle step when enough lanes
//           have triangles waiting.  A lane takes part in a node step as long as it has any node left and in a triangle step as
//           long as it has any triangle left, so steps run with ~3/4 of the lanes instead of half of them.
#pragma once
#include "trace_kernels.cuh"
#include "wide_bvh.cuh"

namespace spt
{
	constexpr uint32_t kUpDone = 0xFFFFFFFFu;          // parent link of the root
	constexpr uint32_t kFastNone = 0xFFFFFFFFu;        // lane: no node work / no triangle work
	constexpr uint32_t kFastClimb = 0xFFFFFFFEu;       // lane: the subtree below is finished, go one level up (`up`)
	constexpr uint32_t kTriLastBit = 0x80000000u;      // FTri::id: last triangle of its leaf

	// ref: bit31 set -> leaf, low bits = first FTri slot; clear -> inner node index.  up: (parent inner node << 1 | side of this
	// node in its parent), kUpDone for the root; both halves of a node carry the node's own link.
	struct alignas(32) FHalf { float lox, loy, loz, hix, hiy, hiz; uint32_t ref, up; };
	struct alignas(64) FNode { FHalf h[2]; };
	struct alignas(64) FTri { float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z; uint32_t id; float blox, bloy, bloz, bhix, bhiy, bhiz; };
	// per original triangle id: the climb link of its leaf, the leaf's first slot, and how far the rest of the scene is: reach[k] = the
	// smallest distance between the (padded) box of this leaf and the box of any sibling subtree the climb has NOT tested yet after
	// kFastReach0/1/2 levels.  A closest-hit walk whose ray is already shorter than that stops climbing there.
	struct alignas(32) FStart { uint32_t up, slot; float reach[3]; uint32_t pad[3]; };
	struct FHeader { float originLimit; float pad; uint32_t r0, r1; };
	static_assert(sizeof(FNode) == 64 && sizeof(FTri) == 64 && sizeof(FStart) == 32, "fast traversal layout");

	struct FastView { const FNode* nodes; const FTri* tris; const FStart* start; const FHeader* header; const FNode* climb; uint32_t numNodes, numTris; };

	constexpr float kFastPad = 1.0f / 262144.0f;       // box padding, relative to the scene's largest |coordinate| (2^-18)
	constexpr float kFastOriginScale = 8.0f;           // rays that start farther out than this many scene extents are replayed
#ifndef SPT_FAST_REACH_L0
#define SPT_FAST_REACH_L0 4u       // even: a climb step takes two levels
#define SPT_FAST_REACH_L1 8u
#define SPT_FAST_REACH_L2 12u
#endif
	constexpr uint32_t kFastReachTab[3] = { SPT_FAST_REACH_L0, SPT_FAST_REACH_L1, SPT_FAST_REACH_L2 };
	constexpr uint32_t kFastReach0 = kFastReachTab[0], kFastReach1 = kFastReachTab[1], kFastReach2 = kFastReachTab[2];     // climb levels at which FStart::reach is sampled
	constexpr uint32_t kFastMaxRd = 0x6F800000u;       // |1/d| must stay below 2^96 (finite, and b * 1/d cannot overflow)

	// ---- build: one thread per node of the build numbering (the scratch of bvh_build.cuh is still in place) ----------------------
	struct FastHeaderKernel
	{
		const float* aabb; FHeader* header;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			float m = 0.0f;
			for (int k = 0; k < 6; k++) { const float a = fabsf(aabb[k]); if (a > m) m = a; }
			FHeader h; h.originLimit = m * kFastOriginScale; h.pad = m * kFastPad; h.r0 = h.r1 = 0u;
			*header = h;
		}
	};
	struct FastLinkKernel      // parent links of the inner nodes (traversal numbering)
	{
		const uint32_t* left; const uint32_t* rank; uint32_t* nodeUp;
		SPT_KERNEL_BODY void operator()(uint32_t node) const
		{
			const uint32_t l = left[node];
			if (!l) return;
			const uint32_t me = rank[node];
			if (node == 0u) nodeUp[me] = kUpDone;
			for (uint32_t side = 0; side < 2u; side++) if (left[l + side]) nodeUp[rank[l + side]] = (me << 1) | side;
		}
	};
	struct FastPackKernel      // node halves, triangle records of leaf children, start table
	{
		const uint32_t* left; const uint32_t* count; const uint32_t* rank; const uint32_t* refIdx; const uint32_t* leafOffsetByRef; const uint32_t* mapping;
		const float* aabb; const V4* vtx; const uint32_t* nodeUp; const FHeader* header;
		FNode* nodes; FTri* tris; FStart* start;
		SPT_KERNEL_BODY void operator()(uint32_t node) const
		{
			const uint32_t l = left[node];
			if (!l) return;
			const uint32_t me = rank[node];
			const uint32_t up = nodeUp[me];
			const float pad = header->pad;
			for (uint32_t side = 0; side < 2u; side++)
			{
				const uint32_t c = l + side;
				const float* bb = aabb + (size_t)c * 6;
				FHalf h;
				h.lox = bb[0] - pad; h.loy = bb[1] - pad; h.loz = bb[2] - pad; h.hix = bb[3] + pad; h.hiy = bb[4] + pad; h.hiz = bb[5] + pad;
				h.up = up;
				if (left[c]) h.ref = rank[c];
				else
				{
					const uint32_t first = leafOffsetByRef[refIdx[c]], n = count[c];
					h.ref = kLeafBit | first;
					for (uint32_t j = 0; j < n; j++)
					{
						const uint32_t tri = mapping[first + j];
						const V4 v0 = vtx[tri * 3], v1 = vtx[tri * 3 + 1], v2 = vtx[tri * 3 + 2];
						FTri t;
						t.v0x = v0.x; t.v0y = v0.y; t.v0z = v0.z;
						t.e1x = v1.x - v0.x; t.e1y = v1.y - v0.y; t.e1z = v1.z - v0.z;      // the reference's own subtractions (Bounds.h:199-200)
						t.e2x = v2.x - v0.x; t.e2y = v2.y - v0.y; t.e2z = v2.z - v0.z;
						t.id = tri | (j + 1u == n ? kTriLastBit : 0u);
						t.blox = bb[0]; t.bloy = bb[1]; t.bloz = bb[2]; t.bhix = bb[3]; t.bhiy = bb[4]; t.bhiz = bb[5];
						tris[first + j] = t;
						start[tri].up = (me << 1) | side; start[tri].slot = first;
					}
				}
				nodes[me].h[side] = h;
			}
		}
	};

	// Climb nodes: TWO levels of the way up per step.  For the child X = (P, side) of inner node P, climb[(P << 1) | side] is an ordinary
	// FNode whose two "children" are the subtrees a walk that has finished X must look at next: X's sibling (the other half of P) and P's
	// own sibling (the other half of P's parent G); its link is G's link, i.e. where the walk stands once both are done.  A climb step is
	// therefore a DOWN step on that record (both boxes tested, the nearer one first, the other one pushed) and the chain of ancestors is
	// walked in half as many steps.
	struct FastClimbKernel     // one thread per inner node (traversal numbering), after FastPackKernel
	{
		const FNode* nodes; const uint32_t* nodeUp; FNode* climb;
		SPT_KERNEL_BODY void operator()(uint32_t me) const
		{
			const uint32_t upMe = nodeUp[me];
			// no second level under the root: a point box in a far corner.  (An inverted box would not do: the slab test orders every interval
			// itself.  A ray aimed exactly at that corner "hits" it and walks the tree once more from the root: more work, same result.)
			FHalf uncle; uncle.lox = uncle.hix = kFltMax; uncle.loy = uncle.hiy = -kFltMax; uncle.loz = uncle.hiz = kFltMax; uncle.ref = 0u; uncle.up = kUpDone;
			uint32_t next = kUpDone;
			if (upMe != kUpDone) { uncle = nodes[upMe >> 1].h[(upMe & 1u) ^ 1u]; next = nodeUp[upMe >> 1]; }
			for (uint32_t side = 0; side < 2u; side++)
			{
				FNode c; c.h[0] = nodes[me].h[side ^ 1u]; c.h[1] = uncle;
				c.h[0].up = next; c.h[1].up = next;
				climb[(me << 1) | side] = c;
			}
		}
	};

	struct FastReachKernel     // one thread per triangle slot that starts a leaf: climb once, record the distances (after FastPackKernel)
	{
		const FNode* nodes; const FTri* tris; FStart* start; const FHeader* header; uint32_t numTris;
		SPT_KERNEL_BODY void operator()(uint32_t slot) const
		{
			if (slot != 0u && !(tris[slot - 1u].id & kTriLastBit)) return;          // not the first triangle of its leaf
			const FTri& t0 = tris[slot];
			const float pad = header->pad;
			const float lo[3] = { t0.blox - pad, t0.bloy - pad, t0.bloz - pad }, hi[3] = { t0.bhix + pad, t0.bhiy + pad, t0.bhiz + pad };
			float reach[3] = { kFltMax, kFltMax, kFltMax };
			uint32_t up = start[t0.id & ~kTriLastBit].up;
			for (uint32_t level = 1u; up != kUpDone; level++)
			{
				const FHalf& h = nodes[up >> 1].h[(up & 1u) ^ 1u];
				up = h.up;
				const float slo[3] = { h.lox, h.loy, h.loz }, shi[3] = { h.hix, h.hiy, h.hiz };
				float d2 = 0.0f;
				for (int k = 0; k < 3; k++)
				{
					const float g = slo[k] - hi[k] > lo[k] - shi[k] ? slo[k] - hi[k] : lo[k] - shi[k];
					if (g > 0.0f) d2 += g * g;
				}
				const float dist = sqrtf(d2) * 0.9999f;
				if (level > kFastReach0 && dist < reach[0]) reach[0] = dist;
				if (level > kFastReach1 && dist < reach[1]) reach[1] = dist;
				if (level > kFastReach2 && dist < reach[2]) reach[2] = dist;
			}
			for (uint32_t j = slot;; j++)
			{
				FStart& s = start[tris[j].id & ~kTriLastBit];
				s.reach[0] = reach[0]; s.reach[1] = reach[1]; s.reach[2] = reach[2]; s.pad[0] = s.pad[1] = s.pad[2] = 0u;
				if (tris[j].id & kTriLastBit) break;
			}
		}
	};

	// ---- the conservative slab test ------------------------------------------------------------------------------------------------
	// Reference (Bounds.cpp:582-604): t = fl(fl(b - o) * rD) per plane, hit iff tmax >= tmin && tmin < maxLen && tmax > 0.
	// Here: t' = fl(b' * rD - c) with c = fl(o * rD) and b' the plane moved outwards by pad = 2^-18 M (M: the scene's largest
	// |coordinate|).  With u = 2^-24 and T = (b - o) rD exactly: |t - T| <= 2u|T| and |fl(b rD - c) - T| <= u|o rD| + u|T|, so the
	// two differ by less than u |rD| (4|o| + 3|b|) <= 2^-21 |rD| max(|o|, |b|), while moving the plane by pad moves t' by
	// pad |rD| = 2^-18 |rD| M outwards.  For |o| <= 8 M (checked per ray; surface points satisfy |o| <= M) the padded interval
	// therefore contains the reference's, and  max(tmin', 0) <= min(tmax', limit)  holds whenever the reference's test passes
	// with maxLen <= limit.  Rays with a reciprocal direction that is not finite or above 2^96 never get here (replayed).
	struct FastRay { V3 rD, nc; };        // nc = -(o * rD)
	SPT_HD bool FastSlab(const FastRay& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float limit, float& tn)
	{
		const float ax = FmaF(lox, r.rD.x, r.nc.x), bx = FmaF(hix, r.rD.x, r.nc.x);
		const float ay = FmaF(loy, r.rD.y, r.nc.y), by = FmaF(hiy, r.rD.y, r.nc.y);
		const float az = FmaF(loz, r.rD.z, r.nc.z), bz = FmaF(hiz, r.rD.z, r.nc.z);
		tn = Max3(fminf(ax, bx), fminf(ay, by), fmaxf(fminf(az, bz), 0.0f));
		const float tf = Min3(fmaxf(ax, bx), fmaxf(ay, by), fminf(fmaxf(az, bz), limit));
		return tn <= tf;
	}
	// reach distances are Euclidean: into the ray's own parameter (t = distance / |d|), a little short
	SPT_HD float FastReachScale(V3 d) { return 0.9999f / sqrtf(d.x * d.x + d.y * d.y + d.z * d.z); }
	// distance from the ray origin to the box of the leaf it claims to start in (0 for a point of that leaf; the reach distances
	// were measured from the box, so a query that starts elsewhere simply has less reach, or none)
	SPT_HD float FastOutside(V3 o, float lox, float loy, float loz, float hix, float hiy, float hiz)
	{
		const float gx = fmaxf(fmaxf(lox - o.x, o.x - hix), 0.0f), gy = fmaxf(fmaxf(loy - o.y, o.y - hiy), 0.0f), gz = fmaxf(fmaxf(loz - o.z, o.z - hiz), 0.0f);
		return sqrtf(gx * gx + gy * gy + gz * gz) * 1.0001f;
	}
	SPT_HD bool FastReachStop(uint32_t level, float limit, float r4, float r7, float r11)
	{
		const float thr = level >= kFastReach2 ? r11 : (level >= kFastReach1 ? r7 : (level >= kFastReach0 ? r4 : -kFltMax));
		return limit < thr;
	}
	SPT_HD bool FastSafe(V3 o, V3 rD, float originLimit)
	{
		return (f2u(rD.x) & 0x7FFFFFFFu) < kFastMaxRd && (f2u(rD.y) & 0x7FFFFFFFu) < kFastMaxRd && (f2u(rD.z) & 0x7FFFFFFFu) < kFastMaxRd &&
			fabsf(o.x) <= originLimit && fabsf(o.y) <= originLimit && fabsf(o.z) <= originLimit;
	}
	// The winner must be a triangle the reference tests whatever it found before: its own leaf box passes the reference's slab
	// test at the winner's distance (closest hit) or at full length (hit-or-miss).  Ancestors follow (monotone under nesting).
	SPT_HD bool FastWinnerOk(const FTri* t, V3 o, V3 rD, float maxLen)
	{
		return SlabTest(o, rD, t->blox, t->bloy, t->bloz, t->bhix, t->bhiy, t->bhiz, maxLen) != kFltMax;
	}

#if defined(SPT_EMU) && defined(SPT_WIDE_STATS)
	static unsigned long long g_fastStats[4];      // host tuning aid: node steps, triangle tests, rays, winners rejected
#define SPT_FSTAT(i) g_fastStats[i]++
#else
#define SPT_FSTAT(i) do { } while (0)
#endif

	// Scalar form (host-compiled kernel bodies, and the specification of the warp loop below: the results do not depend on the order
	// of the steps, see WideBest).  Returns false when the ray must be replayed by the exact kernel.
	SPT_HD bool TraceFast(const FastView& w, V3 o, V3 d, uint32_t ignore, bool anyHit, Hit& hit)
	{
		SPT_FSTAT(2);
		const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
		hit.t = u2f(0x7F800000u); hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
		if (!FastSafe(o, rD, w.header->originLimit)) return false;
		FastRay r; r.rD = rD; r.nc = v3(-(o.x * rD.x), -(o.y * rD.y), -(o.z * rD.z));
		WideBest best; best.Reset();
		uint32_t stack[kStackDepth]; int sp = 0;
		uint32_t cur = 0u, up = kUpDone, level = 0u;
		float reach[3] = { 0.0f, 0.0f, 0.0f };
		if (ignore < w.numTris)
		{
			const FStart& s = w.start[ignore]; up = s.up; cur = kLeafBit | s.slot;
			const FTri* T = w.tris + s.slot;
			const float k = FastReachScale(d), away = FastOutside(o, T->blox, T->bloy, T->bloz, T->bhix, T->bhiy, T->bhiz);
			for (int c = 0; c < 3; c++) reach[c] = anyHit ? 0.0f : (s.reach[c] - away) * k;
		}
		for (;;)
		{
			if (cur & kLeafBit)
			{
				for (const FTri* T = w.tris + (cur & ~kLeafBit);; T++)
				{
					float t, u, v;
					SPT_FSTAT(1);
					if ((T->id & ~kTriLastBit) != ignore && TriTest(o, d, v3(T->v0x, T->v0y, T->v0z), v3(T->e1x, T->e1y, T->e1z), v3(T->e2x, T->e2y, T->e2z), kFltMax, t, u, v))
					{
						if (anyHit)
						{
							if (!FastWinnerOk(T, o, rD, kFltMax)) return false;
							hit.t = t; hit.u = u; hit.v = v; hit.tri = T->id & ~kTriLastBit;
							return true;
						}
						best.Offer(t, u, v, T->id & ~kTriLastBit, (uint32_t)(T - w.tris));
					}
					if (T->id & kTriLastBit) break;
				}
			}
			else
			{
				const FNode* n = w.nodes + cur;
				SPT_FSTAT(0);
				float t0, t1;
				const bool h0 = FastSlab(r, n->h[0].lox, n->h[0].loy, n->h[0].loz, n->h[0].hix, n->h[0].hiy, n->h[0].hiz, best.limit, t0);
				const bool h1 = FastSlab(r, n->h[1].lox, n->h[1].loy, n->h[1].loz, n->h[1].hix, n->h[1].hiy, n->h[1].hiz, best.limit, t1);
				if (h0 || h1)
				{
					const bool first0 = h0 && (!h1 || t0 <= t1);
					cur = first0 ? n->h[0].ref : n->h[1].ref;
					if (h0 && h1) { if (sp >= kStackDepth) return false; stack[sp++] = first0 ? n->h[1].ref : n->h[0].ref; }
					continue;
				}
			}
			if (sp > 0) { cur = stack[--sp]; continue; }
			bool found = false;
			while (up != kUpDone)            // the subtree is exhausted: climb (two levels per step, FastClimbKernel) until a box is hit
			{
				if (FastReachStop(level, best.limit, reach[0], reach[1], reach[2])) break;        // nothing above is within the ray's length
				level += 2u;
				const FNode* n = w.climb + up;
				up = n->h[0].up;
				SPT_FSTAT(0);
				float t0, t1;
				const bool h0 = FastSlab(r, n->h[0].lox, n->h[0].loy, n->h[0].loz, n->h[0].hix, n->h[0].hiy, n->h[0].hiz, best.limit, t0);
				const bool h1 = FastSlab(r, n->h[1].lox, n->h[1].loy, n->h[1].loz, n->h[1].hix, n->h[1].hiy, n->h[1].hiz, best.limit, t1);
				if (h0 || h1)
				{
					const bool first0 = h0 && (!h1 || t0 <= t1);
					cur = first0 ? n->h[0].ref : n->h[1].ref;
					if (h0 && h1) stack[sp++] = first0 ? n->h[1].ref : n->h[0].ref;          // the stack is empty here
					found = true; break;
				}
			}
			if (!found) break;
		}
		if (best.tie) return false;
		if (best.tri != kNoHit && !FastWinnerOk(w.tris + best.rec, o, rD, best.t)) { SPT_FSTAT(3); return false; }
		hit.t = best.t; hit.u = best.u; hit.v = best.v; hit.tri = best.tri;
		return true;
	}

	struct ReplayOut { uint32_t* list; uint32_t* count; };
	// Replay bookkeeping of one scene: [0] work counter of the fast walk, [1] replay count, [2] replay work counter, [3] total replayed (stats)
	struct ReplayBuffers { uint32_t* counters; uint32_t* replayList; uint32_t replayCap; };

#if !defined(SPT_EMU)
#ifndef SPT_FAST_BLOCK
#define SPT_FAST_BLOCK 128
#endif
#ifndef SPT_FAST_MIN_BLOCKS
#define SPT_FAST_MIN_BLOCKS 8      // 64 registers: 8 CTAs of 128 threads per SM (65 registers / 7 CTAs measured 6 % slower, profiles/r02_fast_variants.txt)
#endif
#ifndef SPT_FAST_NODE_STACK
#define SPT_FAST_NODE_STACK 12     // node entries per lane in shared memory; a deeper walk is replayed
#endif
#ifndef SPT_FAST_LEAF_QUEUE
#define SPT_FAST_LEAF_QUEUE 8      // queued leaves per lane (a power of two: ring buffer); a lane with fewer than two free entries sits out of node steps
#endif
#ifndef SPT_FAST_FETCH_MIN_IDLE
#define SPT_FAST_FETCH_MIN_IDLE 8
#endif
#ifndef SPT_FAST_TRI_VOTE
#define SPT_FAST_TRI_VOTE 18       // a triangle step runs when at least this many lanes have a triangle waiting
#endif
#ifndef SPT_FAST_NODE_REPS
#define SPT_FAST_NODE_REPS 4
#endif
#ifndef SPT_FAST_TRI_REPS
#define SPT_FAST_TRI_REPS 4
#endif
#ifndef SPT_FAST_NODE_UNROLL
#define SPT_FAST_NODE_UNROLL SPT_FAST_NODE_REPS     // the repetition loops are unrolled: rolled, the loop control and the re-evaluated lane
#endif                                              // predicates cost 2.6 % of the kernel (profiles/r03q_unroll_variants.txt: 73.9 -> 71.9 ms on C3)
#ifndef SPT_FAST_TRI_UNROLL
#define SPT_FAST_TRI_UNROLL SPT_FAST_TRI_REPS
#endif
#ifndef SPT_FAST_NODE_BIAS
#define SPT_FAST_NODE_BIAS 1       // SPT_FAST_IMMEDIATE: a node step runs when (lanes with a node) * bias >= lanes with a triangle
#endif
	constexpr int kFastBlock = SPT_FAST_BLOCK, kFastNodeStack = SPT_FAST_NODE_STACK, kFastLeafQueue = SPT_FAST_LEAF_QUEUE;
	constexpr int kFastSmemWords = (kFastNodeStack + kFastLeafQueue) * kFastBlock;

#if defined(SPT_FAST_LOOP_STATS)
	// tuning aid: 0 iterations, 1 idle lanes, 2 node steps, 3 lanes in node steps, 4 triangle steps, 5 lanes in triangle steps, 6 refills, 7 lanes refilled,
	// 8 rays retired, 9 triangle votes with no node work left
	__device__ unsigned long long g_fastLoopStats[64];
#define SPT_FL(i, v) do { if (lane == 0) fl_[i] += (v); } while (0)
#define SPT_FL_LANES(i, cond) do { const uint32_t m_ = __ballot_sync(0xffffffffu, (cond)); if (lane == 0) fl_[i] += __popc(m_); } while (0)
#define SPT_FL_MINE(v, cond) do { v += (cond) ? 1u : 0u; } while (0)
#else
#define SPT_FL(i, v) do { } while (0)
#define SPT_FL_LANES(i, cond) do { } while (0)
#define SPT_FL_MINE(v, cond) do { } while (0)
#endif

	__device__ __forceinline__ void ld256(const void* p, uint32_t (&r)[8])     // LDG.E.256 through the read-only path
	{
		asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
			: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
	}

	// TriTest (traverse.cuh) with every lane running the same instruction stream: the acceptance arithmetic (one IEEE division,
	// t, the hit-point test) is computed for all lanes instead of inside a branch that 90 % of the warp-level steps took with
	// three lanes (profiles/r02_SUMMARY.md).  Same operations in the same order, hence the same t, u, v.
	__device__ __forceinline__ bool FastTriTest(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float& outT, float& outU, float& outV)
	{
		const V3 p = cross(d, e2);
		const float det = dot(e1, p);
		const V3 dist = o - v0;
		float u = dot(dist, p);
		const V3 perp = cross(dist, e1);
		float v = dot(d, perp);
		const float uv = u + v;
		const bool pos = (det > 0.0f) & !((u < 0.0f) | (u > det)) & !((v < 0.0f) | (uv > det));
		const bool neg = (det < 0.0f) & !((u > 0.0f) | (u < det)) & !((v > 0.0f) | (uv < det));
		const float invDet = 1.0f / det;
		const float t = dot(e2, perp) * invDet;
		u *= invDet; v *= invDet;
		const V3 pt = o + d * t;
		const float inf = u2f(0x7F800000u);
		outT = t; outU = u; outV = v;
		return (pos | neg) & (t < kFltMax) & (t > -0.0000001f) & ((pt.x != inf) | (pt.y != inf) | (pt.z != inf));
	}

	// kMode 0: the queue's rays as they come.  kMode 1: hit-or-miss rays only, with everything a closest-hit walk needs compiled out (no
	// candidate record, no shrinking ray length, no reach distances).
#ifndef SPT_FAST_EXP_MODE
#define SPT_FAST_EXP_MODE 0
#endif
	template<int kMode, class Source, class Sink>
	__device__ __forceinline__ void TraceFastLoop(const FastView& w, uint32_t n, uint32_t* __restrict__ counter, uint32_t* stackMem, const ReplayOut& replay, Source& src, Sink& sink)
	{
		// [entry][thread], 4-byte entries: the 32 lanes of a warp always hit 32 different banks
		const uint32_t sNode = (uint32_t)__cvta_generic_to_shared(stackMem) + threadIdx.x * 4u;
		const uint32_t sLeaf = sNode + (uint32_t)kFastNodeStack * (kFastBlock * 4u);
		const uint32_t lane = threadIdx.x & 31;
		const float originLimit = w.header->originLimit;
		V3 o = v3(0.0f), d = v3(0.0f);
		FastRay r; r.rD = v3(0.0f); r.nc = v3(0.0f);
		WideBest best; best.Reset();
		uint32_t ignore = kNoHit, index = 0;
		bool anyHitVar = false, bad = false, active = false;
		constexpr bool kAnyOnly = kMode == 1;
#define anyHit (kAnyOnly ? true : anyHitVar)
		uint32_t cur = kFastNone, up = kUpDone, tcur = kFastNone;
		int sp = 0;
		uint32_t level = 0; float reach4 = 0.0f, reach7 = 0.0f, reach11 = 0.0f;          // climb levels done; scaled reach distances (FStart)
		uint32_t qw = 0, qr = 0;          // leaf queue: entries written / read (first in, first out: the leaves nearest to the origin are found first)
		bool exhausted = false;
#if defined(SPT_FAST_LOOP_STATS)
		unsigned long long fl_[16] = {}; uint32_t myNode = 0, myTri = 0;
#endif
		auto pushLeaf = [&](bool yes, uint32_t slot)
		{
			const bool direct = yes && tcur == kFastNone, queued = yes && tcur != kFastNone;
			if (queued) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sLeaf + (qw & (uint32_t)(kFastLeafQueue - 1)) * (kFastBlock * 4u)), "r"(slot) : "memory");
			qw += queued ? 1u : 0u;
			tcur = direct ? slot : tcur;
		};
		// a lane may take a node step when it has a node left and room for the (up to two) leaves the step can find.
		// SPT_FAST_IMMEDIATE: ... and no triangle waiting, i.e. leaves are tested as soon as they are found
#if defined(SPT_FAST_IMMEDIATE)
#define SPT_FAST_CAN_NODE (active && cur != kFastNone && tcur == kFastNone)
#else
#define SPT_FAST_CAN_NODE (active && cur != kFastNone && qw - qr <= (uint32_t)(kFastLeafQueue - 2))
#endif

		for (;;)
		{
			// ---- retire finished rays and refill, both in one go once enough lanes are free ----
			const bool done = active && cur == kFastNone && tcur == kFastNone;
			const uint32_t doneMask = __ballot_sync(0xffffffffu, done), idleMask = __ballot_sync(0xffffffffu, !active);
			const uint32_t freeMask = doneMask | idleMask;
			if (freeMask != 0u && ((!exhausted && __popc(freeMask) >= SPT_FAST_FETCH_MIN_IDLE) || freeMask == 0xffffffffu))
			{
				if (doneMask)
				{
					bool toReplay = done && (bad || (!anyHit && best.tie));
					if (done && !toReplay && best.tri != kNoHit)
					{
						const unsigned char* T = reinterpret_cast<const unsigned char*>(w.tris) + (size_t)best.rec * 64u;
						const float2 b0 = __ldg(reinterpret_cast<const float2*>(T + 40));
						const float4 b1 = __ldg(reinterpret_cast<const float4*>(T + 48));
						toReplay = SlabTestFast(o, r.rD, b0.x, b0.y, b1.x, b1.y, b1.z, b1.w, anyHit ? kFltMax : best.t) == kFltMax;
					}
					const uint32_t rm = __ballot_sync(0xffffffffu, toReplay);
					if (rm)
					{
						const int leader = __ffs(rm) - 1;
						uint32_t rb = 0;
						if ((int)lane == leader) rb = atomicAdd(replay.count, (uint32_t)__popc(rm));
						rb = __shfl_sync(0xffffffffu, rb, leader);
						if (toReplay) replay.list[rb + (uint32_t)__popc(rm & ((1u << lane) - 1u))] = index;
					}
					SPT_FL(8, __popc(doneMask)); SPT_FL_LANES(13, done && anyHit); SPT_FL_LANES(14, done && anyHit && best.tri != kNoHit); SPT_FL_LANES(15, done && !anyHit && best.tri != kNoHit);
#if defined(SPT_FAST_LOOP_STATS)
					if (done)
					{
						const int cls = (anyHit ? 0 : 2) + (best.tri != kNoHit ? 0 : 1);
						atomicAdd(&g_fastLoopStats[16 + cls], (unsigned long long)myNode); atomicAdd(&g_fastLoopStats[20 + cls], (unsigned long long)myTri); atomicAdd(&g_fastLoopStats[24 + cls], 1ull);
						if (cls == 0) atomicAdd(&g_fastLoopStats[28 + (myNode < 17u ? myNode : 17u)], 1ull);
						if (cls == 0) atomicAdd(&g_fastLoopStats[46 + (myTri < 17u ? myTri : 17u)], 1ull);
						myNode = 0; myTri = 0;
					}
#endif
					Hit h; h.t = best.t; h.u = best.u; h.v = best.v; h.tri = best.tri;
					sink.Retire(done && !toReplay, index, h, anyHit, o, d);
					if (done) active = false;
				}
				if (exhausted) { if (freeMask == 0xffffffffu) break; continue; }       // queue exhausted and every lane retired
				const uint32_t want = (uint32_t)__popc(freeMask);
				uint32_t base = 0;
				if (lane == 0) base = atomicAdd(counter, want);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base + want >= n) exhausted = true;
				SPT_FL(6, 1); SPT_FL(7, want);
				bool toReplay = false; uint32_t replayIndex = 0;
				if (!active)
				{
					const uint32_t i = base + (uint32_t)__popc(freeMask & ((1u << lane) - 1u));
					float maxLen;
					if (i < n && src.Load(i, o, d, ignore, maxLen, anyHitVar)
#if defined(SPT_FAST_EXP_SKIP_CLOSEST)
						&& anyHitVar
#endif
#if defined(SPT_FAST_EXP_SKIP_ANY)
						&& !anyHitVar
#endif
						)
					{
						r.rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);          // Ray::SetDirection (Bounds.h:44-48)
						if (!FastSafe(o, r.rD, originLimit)) { toReplay = true; replayIndex = i; }
						else
						{
							r.nc = v3(-(o.x * r.rD.x), -(o.y * r.rD.y), -(o.z * r.rD.z));
							index = i; active = true; bad = false; best.Reset(); sp = 0; qw = 0; qr = 0;
							level = 0u; reach4 = reach7 = reach11 = 0.0f;
							if (ignore < w.numTris)
							{
								uint32_t S[8];
								ld256(w.start + ignore, S);
								up = S[0]; tcur = S[1]; cur = kFastClimb;
								if (!anyHit)
								{
									const unsigned char* T = reinterpret_cast<const unsigned char*>(w.tris) + (size_t)S[1] * 64u;
									const float2 b0 = __ldg(reinterpret_cast<const float2*>(T + 40));
									const float4 b1 = __ldg(reinterpret_cast<const float4*>(T + 48));
									const float k = FastReachScale(d), away = FastOutside(o, b0.x, b0.y, b1.x, b1.y, b1.z, b1.w);
									reach4 = (__uint_as_float(S[2]) - away) * k; reach7 = (__uint_as_float(S[3]) - away) * k; reach11 = (__uint_as_float(S[4]) - away) * k;
								}
							}
							else { cur = 0u; up = kUpDone; tcur = kFastNone; }
						}
					}
				}
				const uint32_t rm = __ballot_sync(0xffffffffu, toReplay);    // rays this walk does not take go straight to the replay list
				if (rm)
				{
					const int leader = __ffs(rm) - 1;
					uint32_t rb = 0;
					if ((int)lane == leader) rb = atomicAdd(replay.count, (uint32_t)__popc(rm));
					rb = __shfl_sync(0xffffffffu, rb, leader);
					if (toReplay) replay.list[rb + (uint32_t)__popc(rm & ((1u << lane) - 1u))] = replayIndex;
				}
				continue;
			}
			// ---- vote ----
			const uint32_t nodeMask = __ballot_sync(0xffffffffu, SPT_FAST_CAN_NODE), triMask = __ballot_sync(0xffffffffu, active && tcur != kFastNone);
			SPT_FL(0, 1); SPT_FL(1, __popc(freeMask));
#if defined(SPT_FAST_IMMEDIATE)
			if (__popc(nodeMask) * SPT_FAST_NODE_BIAS >= __popc(triMask))
#else
			if (nodeMask != 0u && __popc(triMask) < SPT_FAST_TRI_VOTE)
#endif
			{
SPT_UNROLL(SPT_FAST_NODE_UNROLL)
				for (int rep = 0; rep < SPT_FAST_NODE_REPS; rep++)
				{
					const bool take = SPT_FAST_CAN_NODE;
					SPT_FL(2, 1); SPT_FL_LANES(3, take); SPT_FL_LANES(10, take && cur == kFastClimb); SPT_FL_LANES(11, take && anyHit); SPT_FL_MINE(myNode, take);
					if (take)
					{
						// DOWN: both halves of node `cur`.  UP: both halves of the climb node of the subtree just finished (its sibling and its
						// parent's sibling, FastClimbKernel) - the same step on another record, and the way up takes half as many of them.
						const bool climbing = cur == kFastClimb;
						const unsigned char* base = reinterpret_cast<const unsigned char*>(climbing ? w.climb : w.nodes) + (size_t)(climbing ? up : cur) * 64u;
						uint32_t A[8], B[8];
						ld256(base, A);
						ld256(base + 32, B);
						float tA, tB;
						const float limit = kAnyOnly ? kFltMax : best.limit;
						const bool hitA = FastSlab(r, __uint_as_float(A[0]), __uint_as_float(A[1]), __uint_as_float(A[2]), __uint_as_float(A[3]), __uint_as_float(A[4]), __uint_as_float(A[5]), limit, tA);
						const bool hitB = FastSlab(r, __uint_as_float(B[0]), __uint_as_float(B[1]), __uint_as_float(B[2]), __uint_as_float(B[3]), __uint_as_float(B[4]), __uint_as_float(B[5]), limit, tB);
						up = climbing ? A[7] : up;
						level += climbing ? 2u : 0u;
						const uint32_t refA = A[6], refB = B[6];
						const bool inA = hitA && !(refA & kLeafBit), inB = hitB && !(refB & kLeafBit);
						// leaves: the nearer one first
						const bool lfA = hitA && (refA & kLeafBit) != 0u, lfB = hitB && (refB & kLeafBit) != 0u;
						const bool bFirst = lfA && lfB && tB < tA;
						pushLeaf(lfA || lfB, (lfA && !bFirst ? refA : refB) & ~kLeafBit);
						pushLeaf(lfA && lfB, (bFirst ? refA : refB) & ~kLeafBit);
						const bool both = inA && inB;
						const bool aFirst = inA && (!inB || tA <= tB);
						if (both && sp >= kFastNodeStack) { bad = true; cur = kFastNone; tcur = kFastNone; sp = 0; qr = qw; }      // deeper than the short stack: the exact kernel takes the ray
						else
						{
							const bool pop = !(inA || inB) && sp > 0;
							const uint32_t sa = sNode + (uint32_t)(sp - (pop ? 1 : 0)) * (kFastBlock * 4u);
							uint32_t popped = 0u;
							if (both) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa), "r"(aFirst ? refB : refA) : "memory");
							if (pop) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(popped) : "r"(sa) : "memory");
							sp += both ? 1 : (pop ? -1 : 0);
							// nothing below: one level up, unless the root has been passed or nothing above is within the ray's length
							const bool climbOn = up != kUpDone && (kAnyOnly || !FastReachStop(level, best.limit, reach4, reach7, reach11));
							cur = (inA || inB) ? (aFirst ? refA : refB) : (pop ? popped : (climbOn ? kFastClimb : kFastNone));
						}
					}
				}
			}
			else
			{
SPT_UNROLL(SPT_FAST_TRI_UNROLL)
				for (int rep = 0; rep < SPT_FAST_TRI_REPS; rep++)
				{
					const bool take = active && tcur != kFastNone;
					SPT_FL(4, 1); SPT_FL_LANES(5, take); SPT_FL_LANES(12, take && anyHit); SPT_FL_MINE(myTri, take);
					if (take)
					{
						const unsigned char* T = reinterpret_cast<const unsigned char*>(w.tris) + (size_t)tcur * 64u;
						uint32_t P[8];
						ld256(T, P);
						const uint2 q = __ldg(reinterpret_cast<const uint2*>(T + 32));
						const uint32_t triId = q.y & ~kTriLastBit;
						float t, u, v;
						const bool ok = FastTriTest(o, d, v3(__uint_as_float(P[0]), __uint_as_float(P[1]), __uint_as_float(P[2])), v3(__uint_as_float(P[3]), __uint_as_float(P[4]), __uint_as_float(P[5])),
							v3(__uint_as_float(P[6]), __uint_as_float(P[7]), __uint_as_float(q.x)), t, u, v) && triId != ignore;      // BVH.cpp:136-139
						// WideBest::Offer, as selects
						if (kAnyOnly) { best.tri = ok ? triId : best.tri; best.rec = ok ? tcur : best.rec; }
						else
						{
							const bool none = best.tri == kNoHit;
							const bool better = ok && (none || t < best.t);
							const float band = __fmaf_rn(fabsf(t), kTieBand, t);
							best.tie = better ? (!none && best.t <= band) : (best.tie || (ok && t <= best.limit));
							best.limit = better ? fminf(band, kFltMax) : best.limit;
							best.t = better ? t : best.t; best.u = better ? u : best.u; best.v = better ? v : best.v;
							best.tri = better ? triId : best.tri; best.rec = better ? tcur : best.rec;
						}
						const bool stop = ok && anyHit;                                       // hit-or-miss query: the walk is over
						const bool last = (q.y & kTriLastBit) != 0u;
						const bool next = !stop && last && qr != qw;
						uint32_t queued = kFastNone;
						if (next) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(queued) : "r"(sLeaf + (qr & (uint32_t)(kFastLeafQueue - 1)) * (kFastBlock * 4u)) : "memory");
						qr += next ? 1u : 0u;
						tcur = stop ? kFastNone : (last ? queued : tcur + 1u);
						if (stop) { cur = kFastNone; sp = 0; qr = qw; }
					}
				}
			}
		}
#if defined(SPT_FAST_LOOP_STATS)
		if (lane == 0) for (int k = 0; k < 16; k++) atomicAdd(&g_fastLoopStats[k], fl_[k]);
#endif
#undef SPT_FAST_CAN_NODE
#undef anyHit
	}

	// exact replay: the rays on the replay list through the reference-visit-order warp loop
	// A short replay list is SPREAD: only every 2^shift-th work index carries a ray, so a warp that fetches 32 indices traces 32 >> shift
	// rays.  The rays on the list are the awkward ones (grazing ties, deep walks) and their walks have nothing in common: 32 of them in one
	// warp run one after the other, and a list of a few hundred rays kept a dozen warps busy for 0.3 ms while the machine idled.
	constexpr uint32_t kReplaySpreadShift = 3u, kReplaySpreadMax = 1u << 13;      // longer lists fill the replay grid (296 CTAs) as they are
	__device__ __forceinline__ uint32_t ReplaySpread(uint32_t& n) { const uint32_t shift = n <= kReplaySpreadMax ? kReplaySpreadShift : 0u; n <<= shift; return shift; }
	template<class Inner>
	struct ReplaySource
	{
		const uint32_t* list; Inner inner; uint32_t shift;
		__device__ __forceinline__ bool Load(uint32_t i, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) const
		{
			if (i & ((1u << shift) - 1u)) return false;
			return inner.Load(list[i >> shift], o, d, ignore, maxLen, anyHit);
		}
	};
	template<class Inner>
	struct ReplaySink
	{
		const uint32_t* list; Inner inner; uint32_t shift;
		__device__ __forceinline__ void Retire(bool finished, uint32_t i, const Hit& h, bool anyHit, V3 o, V3 d) const { inner.Retire(finished, finished ? list[i >> shift] : 0u, h, anyHit, o, d); }
	};

	__global__ void __launch_bounds__(kFastBlock, SPT_FAST_MIN_BLOCKS) k_trace_fast_rays(FastView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, ReplayOut replay)
	{
		__shared__ uint32_t stackMem[kFastSmemWords];
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; QueueSink sink{ hits };
		TraceFastLoop<0>(w, n, counter, stackMem, replay, src, sink);
	}
	__global__ void __launch_bounds__(kFastBlock, SPT_FAST_MIN_BLOCKS) k_trace_fast_level(FastView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, WavefrontOut out, ReplayOut replay)
	{
		__shared__ uint32_t stackMem[kFastSmemWords];
		{ const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; WavefrontSink sink{ out.status + *out.auxBase, out.slow, out.slowCount };
		TraceFastLoop<SPT_FAST_EXP_MODE>(w, n, counter, stackMem, replay, src, sink);
	}
	// The replay list is short (a few thousand rays of a 70 M ray level): a grid of one warp-sized CTA per 32 rays instead of the
	// machine-wide persistent grid keeps an (almost) empty replay at launch latency.
	__global__ void __launch_bounds__(kTraceBlock) k_replay_rays(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		const uint32_t* __restrict__ list, const uint32_t* __restrict__ nPtr, uint32_t cap, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		uint32_t n = *nPtr; if (n > cap) n = cap;
		if (!n) return;
		const uint32_t shift = ReplaySpread(n);
		ReplaySource<QueueSource> src{ list, QueueSource{ rays }, shift }; ReplaySink<QueueSink> sink{ list, QueueSink{ hits }, shift };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
	}
	__global__ void __launch_bounds__(kTraceBlock) k_replay_level(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		const uint32_t* __restrict__ list, const uint32_t* __restrict__ nPtr, uint32_t cap, uint32_t* __restrict__ counter, WavefrontOut out)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		uint32_t n = *nPtr; if (n > cap) n = cap;
		if (!n) return;
		const uint32_t shift = ReplaySpread(n);
		ReplaySource<QueueSource> src{ list, QueueSource{ rays }, shift };
		ReplaySink<WavefrontSink> sink{ list, WavefrontSink{ out.status + *out.auxBase, out.slow, out.slowCount }, shift };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
	}

	inline int FastGridSize()
	{
		static int grid = 0;
		if (!grid)
		{
			int dev = 0, sms = 148, perSm = 1;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace_fast_level, kFastBlock, 0);
			grid = sms * (perSm > 0 ? perSm : 1);
		}
		return grid;
	}
	inline int ReplayGridSize() { const int g = TraceGridSize(); return g < 296 ? g : 296; }

	struct AccumulateReplayKernel { uint32_t* c; SPT_KERNEL_BODY void operator()(uint32_t) const { c[3] += c[1]; } };

	// closest hits / hit-or-miss for a ray queue through the origin-local walk + exact replay (QueueSink: hits[i] for every ray)
	inline void LaunchTraceRaysFast(Ctx& ctx, const FastView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
	{
		if (!n || !ctx.ok) return;
		DevMemset(ctx, b.counters, 0, 3 * sizeof(uint32_t));
		k_trace_fast_rays<<<FastGridSize(), kFastBlock, 0, ctx.stream>>>(w, rays, hits, n, nPtr, b.counters, ReplayOut{ b.replayList, b.counters + 1 });
		k_replay_rays<<<ReplayGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, b.replayList, b.counters + 1, b.replayCap, b.counters + 2);
		launch_for(ctx, 1, AccumulateReplayKernel{ b.counters });
		ctx.kernelLaunches += 2;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
	inline void LaunchTraceLevelFast(Ctx& ctx, const FastView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
	{
		if (!cap || !ctx.ok) return;
		DevMemset(ctx, b.counters, 0, 3 * sizeof(uint32_t));
		k_trace_fast_level<<<FastGridSize(), kFastBlock, 0, ctx.stream>>>(w, rays, hits, cap, nPtr, b.counters, out, ReplayOut{ b.replayList, b.counters + 1 });
		k_replay_level<<<ReplayGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, b.replayList, b.counters + 1, b.replayCap, b.counters + 2, out);
		launch_for(ctx, 1, AccumulateReplayKernel{ b.counters });
		ctx.kernelLaunches += 2;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	inline void LaunchTraceRaysFast(Ctx& ctx, const FastView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
	{
		LocalStack st;
		if (nPtr && *nPtr < n) n = *nPtr;
		for (uint32_t i = 0; i < n; i++)
		{
			if (rays[i].tmax == -1.0f) continue;
			const V3 o = v3(rays[i].ox, rays[i].oy, rays[i].oz), d = v3(rays[i].dx, rays[i].dy, rays[i].dz);
			if (!TraceFast(w, o, d, rays[i].ignoreTri, rays[i].tmax < 0.0f, hits[i]))
			{
				TraceClosest(bvh, o, d, rays[i].ignoreTri, fabsf(rays[i].tmax), st, hits[i]);
				b.counters[3]++;
			}
		}
		ctx.kernelLaunches += 3;
	}
	inline void LaunchTraceLevelFast(Ctx& ctx, const FastView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
	{
		LocalStack st;
		const uint32_t n = *nPtr < cap ? *nPtr : cap;
		for (uint32_t i = 0; i < n; i++)
		{
			if (rays[i].tmax == -1.0f) continue;
			const V3 o = v3(rays[i].ox, rays[i].oy, rays[i].oz), d = v3(rays[i].dx, rays[i].dy, rays[i].dz);
			Hit h;
			if (!TraceFast(w, o, d, rays[i].ignoreTri, rays[i].tmax < 0.0f, h))
			{
				TraceClosest(bvh, o, d, rays[i].ignoreTri, fabsf(rays[i].tmax), st, h);
				b.counters[3]++;
			}
			out.status[*out.auxBase + i] = h.tri != kNoHit ? 1 : 0;
			if (h.tri != kNoHit && !(rays[i].tmax < 0.0f)) PushSlow(out, h, rays[i], i);
		}
		ctx.kernelLaunches += 3;
	}
#endif
}
