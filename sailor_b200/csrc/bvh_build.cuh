// bvh_build.cuh — on-device BVH construction that reproduces the reference's tree BIT FOR BIT.
//
// Reference algorithm (reference Raytracing/BVH.cpp): recursive top-down binned SAH, 8 bins x 3 axes
// (FindBestSplitPlane :15-88), leaf when triCount <= 4 or splitCost >= triCount*area (Subdivide :215-234),
// sequential in-place partition on centroid[axis] < splitPos (:236-251), children allocated depth-first (:262-263),
// then a leaf-contiguous re-layout with a per-leaf stable sort by Heron "square area", descending (:298-337).
// It is single-threaded ("TODO: Parallelize", :304).
//
// Device formulation: LEVEL-SYNCHRONOUS.  All nodes of one depth are processed by the same kernels, one thread
// per triangle slot, so every level streams the whole (L2-resident) triangle set once:
//   bounds  : vertex + centroid min/max per node, integer atomics on order-preserving float keys (min/max are
//             order independent, so the result equals the reference's sequential glm::min/max chain)
//   bin     : per node, 3 axes x 8 bins x {count, 6 bound keys} with the reference's bin index arithmetic
//   split   : one thread per node replays FindBestSplitPlane's plane sweep in the reference's float order
//   flags+scan, partition: the reference's swap loop (BVH.cpp:238-251) is sequential, but its final permutation has a
//             closed form (derived in DESIGN.md "partition"): left slots keep their "left" elements, each left hole
//             takes the next "left" element from the right end; right slots shift down by one, and each slot just
//             below a pulled element takes the next hole's element.  Computed with one prefix sum -> identical
//             m_triIdx order, hence identical leaf order after the stable area sort.
// A final pass renumbers nodes into the reference's depth-first allocation order and emits the reference's
// BVHNode array + m_triIdxMapping, plus the traversal layout (traverse.cuh).
#pragma once
#include "backend.h"
#include "../../include/sailor_pt.h"

namespace spt
{
	constexpr uint32_t kBins = 8;                     // BVH.cpp:21
	constexpr uint32_t kBinWords = 7;                 // count + 3 min keys + 3 max keys
	constexpr uint32_t kNodeBinWords = 3 * kBins * kBinWords;

	// Build-time node, breadth-first numbering. SoA in BuildState.
	struct BuildState
	{
		// per triangle slot
		const V4* vtx; const V4* centroid;
		uint32_t* idxA; uint32_t* idxB;               // m_triIdx, double buffered
		uint32_t* nodeOfA; uint32_t* nodeOfB;         // owning build node of each slot
		uint32_t* flags; uint32_t* scan;              // scan has N+1 entries
		uint32_t* holes; uint32_t* srcs;
		// per build node
		uint32_t* first; uint32_t* count; uint32_t* left;     // left = 0 -> leaf (root is never a child)
		uint32_t* keys;                               // 12 per node: vmin[3] vmax[3] cmin[3] cmax[3]
		float* aabb;                                  // 6 per node (decoded vmin, vmax)
		uint32_t* state;                              // per node: bit0 binning, bit1 partitioning, bit2 split
		float* splitPos; uint32_t* splitAxis; uint32_t* nL;
		float* binScale;                              // 6 per node: boundsMin[3], scale[3]; scale = 0 marks a skipped axis
		uint32_t* bins;                               // per BINNED node of the current level: kNodeBinWords
		uint32_t* binSlot; uint32_t* binCounter;      // node -> its slot in `bins` (handed out per level)
		uint32_t* splitFlag; uint32_t* splitScan;     // per node of the current level (+1)
		uint32_t n;
	};

	enum : uint32_t { kStBinning = 1u, kStPartition = 2u, kStSplit = 4u };

	SPT_HD float AabbArea(V3 mn, V3 mx)
	{
		// AABB::Area (Bounds.cpp:443-447) == BVHNode::CalculateCost's surfaceArea (BVH.h:28-33)
		const V3 e = mx - mn;
		return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
	}

	struct InitSlotsKernel
	{
		BuildState s;
		SPT_KERNEL_BODY void operator()(uint32_t p) const { s.idxA[p] = p; s.nodeOfA[p] = 0; }   // BVH.cpp:286-289
	};

	struct InitNodesKernel   // reset the key accumulators of the nodes of one level
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			uint32_t* k = s.keys + (size_t)(levelStart + i) * 12;
			const uint32_t kmin = float_key(1e30f), kmax = float_key(-1e30f);          // UpdateNodeBounds :199-200
			k[0] = k[1] = k[2] = kmin; k[3] = k[4] = k[5] = kmax;
			k[6] = k[7] = k[8] = float_key(kFltMax);                                    // boundsMin :26
			k[9] = k[10] = k[11] = float_key(-30000000.0f);                             // boundsMax :27
		}
	};

	struct BoundsKernel      // UpdateNodeBounds (:193-213) + the centroid min/max loop of FindBestSplitPlane (:29-34)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t p) const
		{
			const uint32_t node = s.nodeOfA[p];
			if (node < levelStart) return;
			const uint32_t tri = s.idxA[p];
			const V4 a = s.vtx[tri * 3], b = s.vtx[tri * 3 + 1], c = s.vtx[tri * 3 + 2], ce = s.centroid[tri];
			uint32_t* k = s.keys + (size_t)node * 12;
			const float mn[3] = { glm_min(glm_min(a.x, b.x), c.x), glm_min(glm_min(a.y, b.y), c.y), glm_min(glm_min(a.z, b.z), c.z) };
			const float mx[3] = { glm_max(glm_max(a.x, b.x), c.x), glm_max(glm_max(a.y, b.y), c.y), glm_max(glm_max(a.z, b.z), c.z) };
			const float cc[3] = { ce.x, ce.y, ce.z };
			for (int d = 0; d < 3; d++)
			{
				atomic_min_u32(k + d, float_key(mn[d]));
				atomic_max_u32(k + 3 + d, float_key(mx[d]));
				atomic_min_u32(k + 6 + d, float_key(cc[d]));
				atomic_max_u32(k + 9 + d, float_key(cc[d]));
			}
		}
	};

	struct PrepareKernel     // per node of the level: decode bounds, decide whether Subdivide goes on to bin (:223-231, :36-43)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			const uint32_t* k = s.keys + (size_t)node * 12;
			float* bb = s.aabb + (size_t)node * 6;
			for (int d = 0; d < 6; d++) bb[d] = key_float(k[d]);
			uint32_t st = 0;
			if (s.count[node] > 4)                                                      // :223
			{
				st = kStBinning;
				float* bs = s.binScale + (size_t)node * 6;
				for (int a = 0; a < 3; a++)
				{
					const float bmin = key_float(k[6 + a]), bmax = key_float(k[9 + a]);
					bs[a] = bmin;
					bs[3 + a] = (bmin == bmax) ? 0.0f : (float)kBins / (bmax - bmin);   // :36-43 (0 = axis skipped)
				}
				const uint32_t slot = atomic_add_u32(s.binCounter, 1u);                 // slot order is irrelevant: bins are per node
				s.binSlot[node] = slot;                                                 // the slot's words are reset by BinInitKernel
			}
			s.state[node] = st;
			s.left[node] = 0;
		}
	};

	struct BinInitKernel     // one thread per word of the bin slots handed out by PrepareKernel (a node's 168 words were 130 us of serial stores per level)
	{
		BuildState s; uint64_t base;      // base: first word of this launch (scenes beyond 2^32 bin words take several)
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint64_t w = base + i;
			if (w / kNodeBinWords >= *s.binCounter) return;
			const uint32_t r = (uint32_t)(w % kNodeBinWords) % kBinWords;
			s.bins[w] = r == 0 ? 0u : (r < 4 ? float_key(10e30f) : float_key(-10e30f));     // AABB defaults, Bounds.h:112-113
		}
	};

	struct BinKernel         // FindBestSplitPlane's binning loop (:44-54)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t p) const
		{
			const uint32_t node = s.nodeOfA[p];
			if (node < levelStart || !(s.state[node] & kStBinning)) return;
			const uint32_t tri = s.idxA[p];
			const V4 a = s.vtx[tri * 3], b = s.vtx[tri * 3 + 1], c = s.vtx[tri * 3 + 2], ce = s.centroid[tri];
			const float mn[3] = { glm_min(glm_min(a.x, b.x), c.x), glm_min(glm_min(a.y, b.y), c.y), glm_min(glm_min(a.z, b.z), c.z) };
			const float mx[3] = { glm_max(glm_max(a.x, b.x), c.x), glm_max(glm_max(a.y, b.y), c.y), glm_max(glm_max(a.z, b.z), c.z) };
			const float cc[3] = { ce.x, ce.y, ce.z };
			const float* bs = s.binScale + (size_t)node * 6;
			uint32_t* bins = s.bins + (size_t)s.binSlot[node] * kNodeBinWords;
			for (int ax = 0; ax < 3; ax++)
			{
				const float scale = bs[3 + ax];
				if (scale == 0.0f) continue;
				int32_t bi = (int32_t)((cc[ax] - bs[ax]) * scale);                      // :49-50
				bi = bi < (int32_t)kBins - 1 ? bi : (int32_t)kBins - 1;
				if (bi < 0) bi = 0;   // the reference would index out of bounds here; unreachable for finite inputs
				uint32_t* bn = bins + ((uint32_t)ax * kBins + (uint32_t)bi) * kBinWords;
				atomic_add_u32(bn, 1u);
				for (int d = 0; d < 3; d++) { atomic_min_u32(bn + 1 + d, float_key(mn[d])); atomic_max_u32(bn + 4 + d, float_key(mx[d])); }
			}
		}
	};

#if !defined(SPT_EMU)
	// ---- device forms of BoundsKernel / BinKernel --------------------------------------------------------------------
	// Triangle slots of one node are contiguous, so near the root a whole CTA works on ONE node and 256 threads would
	// hammer the same dozen addresses.  min/max/add are order independent, so they are first reduced on chip:
	//   bounds : warp reduction of the 12 integer keys (redux.sync), then shared memory across the CTA's warps -> 12
	//            global atomics per CTA instead of 3072;
	//   bins   : a 3 x 8 x 7-word histogram in shared memory per CTA (shared-memory atomics), flushed once.
	// CTAs that straddle several nodes (deep levels, small nodes, little contention) use the per-slot atomics.
	constexpr int kBuildBlock = 256;

	__global__ void __launch_bounds__(kBuildBlock) k_build_bounds(BuildState s, uint32_t levelStart)
	{
		__shared__ uint32_t sNode[2];
		__shared__ uint32_t sKeys[kBuildBlock / 32][12];
		const uint32_t p = blockIdx.x * kBuildBlock + threadIdx.x;
		const bool inRange = p < s.n;
		const uint32_t node = inRange ? s.nodeOfA[p] : 0xFFFFFFFFu;
		const uint32_t firstNode = s.nodeOfA[blockIdx.x * kBuildBlock];
		const uint32_t lastIdx = min((blockIdx.x + 1u) * kBuildBlock, s.n) - 1u;
		const bool uniform = firstNode == s.nodeOfA[lastIdx];       // slots of a node are contiguous
		const bool active = inRange && node >= levelStart;
		uint32_t k[12];
		if (active)
		{
			const uint32_t tri = s.idxA[p];
			const V4 a = s.vtx[tri * 3], b = s.vtx[tri * 3 + 1], c = s.vtx[tri * 3 + 2], ce = s.centroid[tri];
			k[0] = float_key(glm_min(glm_min(a.x, b.x), c.x)); k[1] = float_key(glm_min(glm_min(a.y, b.y), c.y)); k[2] = float_key(glm_min(glm_min(a.z, b.z), c.z));
			k[3] = float_key(glm_max(glm_max(a.x, b.x), c.x)); k[4] = float_key(glm_max(glm_max(a.y, b.y), c.y)); k[5] = float_key(glm_max(glm_max(a.z, b.z), c.z));
			k[6] = k[9] = float_key(ce.x); k[7] = k[10] = float_key(ce.y); k[8] = k[11] = float_key(ce.z);
		}
		else
		{
#pragma unroll
			for (int d = 0; d < 12; d++) k[d] = ((d % 6) < 3) ? 0xFFFFFFFFu : 0u;    // identities of min / max
		}
		if (!uniform)
		{
			// the CTA straddles nodes: lanes of one node (contiguous slots, so usually most of the warp) reduce among themselves first and
			// one of them touches memory - 12 atomics per (warp, node) instead of 12 per triangle.  Around level 12 of a 1 M-triangle build
			// (nodes of ~250 slots: every CTA straddles two) the per-triangle form spent 0.19 ms per level on the same few addresses.
			const uint32_t peers = __match_any_sync(0xffffffffu, active ? node : 0xFFFFFFFFu);
#pragma unroll
			for (int d = 0; d < 12; d++) k[d] = ((d % 6) < 3) ? __reduce_min_sync(peers, k[d]) : __reduce_max_sync(peers, k[d]);
			if (active && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1))
			{
				uint32_t* g = s.keys + (size_t)node * 12;
#pragma unroll
				for (int d = 0; d < 12; d++) { if ((d % 6) < 3) atomicMin(g + d, k[d]); else atomicMax(g + d, k[d]); }
			}
			return;
		}
		if (firstNode < levelStart) return;                           // the whole CTA belongs to a finished node
#pragma unroll
		for (int d = 0; d < 12; d++) k[d] = ((d % 6) < 3) ? __reduce_min_sync(0xffffffffu, k[d]) : __reduce_max_sync(0xffffffffu, k[d]);
		const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		if (lane == 0) { for (int d = 0; d < 12; d++) sKeys[warp][d] = k[d]; }
		__syncthreads();
		if (threadIdx.x < 12)
		{
			const int d = threadIdx.x;
			uint32_t v = sKeys[0][d];
			for (int w = 1; w < kBuildBlock / 32; w++) v = ((d % 6) < 3) ? min(v, sKeys[w][d]) : max(v, sKeys[w][d]);
			uint32_t* g = s.keys + (size_t)firstNode * 12;
			if ((d % 6) < 3) atomicMin(g + d, v); else atomicMax(g + d, v);
		}
		(void)sNode;
	}

	__global__ void __launch_bounds__(kBuildBlock) k_build_bin(BuildState s, uint32_t levelStart)
	{
		__shared__ uint32_t sBins[kNodeBinWords];
		const uint32_t p = blockIdx.x * kBuildBlock + threadIdx.x;
		const bool inRange = p < s.n;
		const uint32_t node = inRange ? s.nodeOfA[p] : 0xFFFFFFFFu;
		const uint32_t firstNode = s.nodeOfA[blockIdx.x * kBuildBlock];
		const uint32_t lastIdx = min((blockIdx.x + 1u) * kBuildBlock, s.n) - 1u;
		const bool uniform = firstNode == s.nodeOfA[lastIdx];
		if (uniform)
		{
			if (firstNode < levelStart || !(s.state[firstNode] & kStBinning)) return;
			const uint32_t kmin = float_key(10e30f), kmax = float_key(-10e30f);
			for (uint32_t w = threadIdx.x; w < kNodeBinWords; w += kBuildBlock) { const uint32_t r = w % kBinWords; sBins[w] = r == 0 ? 0u : (r < 4 ? kmin : kmax); }
			__syncthreads();
		}
		const bool active = inRange && node >= levelStart && (s.state[node] & kStBinning);
		float mnA[3] = { 0.0f, 0.0f, 0.0f }, mxA[3] = { 0.0f, 0.0f, 0.0f }, ccA[3] = { 0.0f, 0.0f, 0.0f };
		if (active)
		{
			const uint32_t tri = s.idxA[p];
			const V4 a = s.vtx[tri * 3], b = s.vtx[tri * 3 + 1], c = s.vtx[tri * 3 + 2], ce = s.centroid[tri];
			mnA[0] = glm_min(glm_min(a.x, b.x), c.x); mnA[1] = glm_min(glm_min(a.y, b.y), c.y); mnA[2] = glm_min(glm_min(a.z, b.z), c.z);
			mxA[0] = glm_max(glm_max(a.x, b.x), c.x); mxA[1] = glm_max(glm_max(a.y, b.y), c.y); mxA[2] = glm_max(glm_max(a.z, b.z), c.z);
			ccA[0] = ce.x; ccA[1] = ce.y; ccA[2] = ce.z;
		}
		// every lane takes part in the votes below (inactive lanes and skipped axes form their own group and write nothing)
		{
			const float* bs = active ? s.binScale + (size_t)node * 6 : nullptr;
			uint32_t* bins = active ? (uniform ? sBins : (s.bins + (size_t)s.binSlot[node] * kNodeBinWords)) : nullptr;
			uint32_t km[6];
#pragma unroll
			for (int d = 0; d < 3; d++) { km[d] = active ? float_key(mnA[d]) : 0xFFFFFFFFu; km[3 + d] = active ? float_key(mxA[d]) : 0u; }
#pragma unroll
			for (int ax = 0; ax < 3; ax++)
			{
				const float scale = active ? bs[3 + ax] : 0.0f;
				const bool on = active && scale != 0.0f;
				int32_t bi = on ? (int32_t)((ccA[ax] - bs[ax]) * scale) : 0;            // BVH.cpp:49-50
				bi = bi < (int32_t)kBins - 1 ? bi : (int32_t)kBins - 1;
				if (bi < 0) bi = 0;
				// lanes of the same (node, bin) reduce among themselves: one lane per group does the 7 atomics (count, 3 min keys, 3 max keys)
				const uint32_t peers = __match_any_sync(0xffffffffu, on ? ((node << 3) | (uint32_t)bi) : 0xFFFFFFFFu);
				uint32_t r[6];
#pragma unroll
				for (int d = 0; d < 3; d++) { r[d] = __reduce_min_sync(peers, km[d]); r[3 + d] = __reduce_max_sync(peers, km[3 + d]); }
				if (on && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1))
				{
					uint32_t* bn = bins + ((uint32_t)ax * kBins + (uint32_t)bi) * kBinWords;
					atomicAdd(bn, (uint32_t)__popc(peers));
#pragma unroll
					for (int d = 0; d < 3; d++) { atomicMin(bn + 1 + d, r[d]); atomicMax(bn + 4 + d, r[3 + d]); }
				}
			}
		}
		if (uniform)
		{
			__syncthreads();
			uint32_t* g = s.bins + (size_t)s.binSlot[firstNode] * kNodeBinWords;
			for (uint32_t w = threadIdx.x; w < kNodeBinWords; w += kBuildBlock)
			{
				const uint32_t r = w % kBinWords, v = sBins[w];
				if (r == 0) { if (v) atomicAdd(g + w, v); }
				else if (r < 4) atomicMin(g + w, v);
				else atomicMax(g + w, v);
			}
		}
	}

	inline void LaunchBuildBounds(Ctx& ctx, const BuildState& s, uint32_t levelStart)
	{
		if (!ctx.ok) return;
		k_build_bounds<<<(s.n + kBuildBlock - 1) / kBuildBlock, kBuildBlock, 0, ctx.stream>>>(s, levelStart);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
	inline void LaunchBuildBin(Ctx& ctx, const BuildState& s, uint32_t levelStart)
	{
		if (!ctx.ok) return;
		k_build_bin<<<(s.n + kBuildBlock - 1) / kBuildBlock, kBuildBlock, 0, ctx.stream>>>(s, levelStart);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	inline void LaunchBuildBounds(Ctx& ctx, const BuildState& s, uint32_t levelStart) { launch_for(ctx, s.n, BoundsKernel{ s, levelStart }); }
	inline void LaunchBuildBin(Ctx& ctx, const BuildState& s, uint32_t levelStart) { launch_for(ctx, s.n, BinKernel{ s, levelStart }); }
#endif

	struct SplitKernel       // the plane sweep of FindBestSplitPlane (:56-86) + Subdivide's cost test (:226-234)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			if (!(s.state[node] & kStBinning)) return;
			const float* bs = s.binScale + (size_t)node * 6;
			const uint32_t* bins = s.bins + (size_t)s.binSlot[node] * kNodeBinWords;
			float bestCost = kFltMax;
			int32_t axis = 0; float splitPos = 0.0f;                                     // Subdivide: int32_t axis{}; float splitPos{};
			for (uint32_t a = 0; a < 3; a++)
			{
				if (bs[3 + a] == 0.0f) continue;
				const float boundsMin = bs[a];
				const float boundsMax = key_float(s.keys[(size_t)node * 12 + 9 + a]);
				float leftArea[kBins - 1], rightArea[kBins - 1];
				int32_t leftCount[kBins - 1], rightCount[kBins - 1];
				V3 lmin = v3(10e30f), lmax = v3(-10e30f), rmin = v3(10e30f), rmax = v3(-10e30f);
				int32_t leftSum = 0, rightSum = 0;
				for (int32_t b = 0; b < (int32_t)kBins - 1; b++)
				{
					const uint32_t* L = bins + (a * kBins + (uint32_t)b) * kBinWords;
					leftSum += (int32_t)L[0];
					leftCount[b] = leftSum;
					// leftBox.Extend(bin.m_bounds): m_min = glm::min(inner.m_min, m_min) (Bounds.cpp:449-453)
					lmin = glm_min(v3(key_float(L[1]), key_float(L[2]), key_float(L[3])), lmin);
					lmax = glm_max(v3(key_float(L[4]), key_float(L[5]), key_float(L[6])), lmax);
					leftArea[b] = AabbArea(lmin, lmax);
					const uint32_t* R = bins + (a * kBins + (kBins - 1 - (uint32_t)b)) * kBinWords;
					rightSum += (int32_t)R[0];
					rightCount[kBins - 2 - b] = rightSum;
					rmin = glm_min(v3(key_float(R[1]), key_float(R[2]), key_float(R[3])), rmin);
					rmax = glm_max(v3(key_float(R[4]), key_float(R[5]), key_float(R[6])), rmax);
					rightArea[kBins - 2 - b] = AabbArea(rmin, rmax);
				}
				const float scale = (boundsMax - boundsMin) / (float)kBins;             // :76
				for (int32_t b = 0; b < (int32_t)kBins - 1; b++)
				{
					const float planeCost = (float)leftCount[b] * leftArea[b] + (float)rightCount[b] * rightArea[b];
					if (planeCost < bestCost)
					{
						axis = (int32_t)a;
						splitPos = boundsMin + scale * (float)(b + 1);
						bestCost = planeCost;
					}
				}
			}
			const float* bb = s.aabb + (size_t)node * 6;
			const float nosplitCost = (float)s.count[node] * AabbArea(v3(bb[0], bb[1], bb[2]), v3(bb[3], bb[4], bb[5]));
			uint32_t st = 0;
			if (!(bestCost >= nosplitCost))                                             // :231-234
			{
				st = kStPartition;
				s.splitAxis[node] = (uint32_t)axis;
				s.splitPos[node] = splitPos;
			}
			s.state[node] = st;
		}
	};

	struct FlagKernel        // predicate of the partition loop (:240)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t p) const
		{
			const uint32_t node = s.nodeOfA[p];
			uint32_t f = 0;
			if (node >= levelStart && (s.state[node] & kStPartition))
			{
				const V4 ce = s.centroid[s.idxA[p]];
				const uint32_t ax = s.splitAxis[node];
				const float c = ax == 0 ? ce.x : (ax == 1 ? ce.y : ce.z);
				f = (c < s.splitPos[node]) ? 1u : 0u;
			}
			s.flags[p] = f;
		}
	};

	struct CountKernel       // leftCount + "abort split if one of the sides is empty" (:253-259)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			uint32_t split = 0;
			if (s.state[node] & kStPartition)
			{
				const uint32_t f = s.first[node], c = s.count[node];
				const uint32_t nL = s.scan[f + c] - s.scan[f];
				s.nL[node] = nL;
				// NOTE: the reference has already permuted m_triIdx when it aborts; the order inside a leaf matters
				// (stable sort), so the permutation is applied for aborted splits too (state keeps kStPartition).
				if (nL != 0 && nL != c) split = 1;
			}
			if (split) s.state[node] |= kStSplit;
			s.splitFlag[i] = split;
		}
	};

	struct AllocKernel       // child creation (:261-271), breadth-first ids
	{
		BuildState s; uint32_t levelStart; uint32_t nextStart;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			if (!(s.state[node] & kStSplit)) return;
			const uint32_t l = nextStart + 2 * s.splitScan[i], r = l + 1;
			const uint32_t f = s.first[node], c = s.count[node], nL = s.nL[node];
			s.first[l] = f; s.count[l] = nL;
			s.first[r] = f + nL; s.count[r] = c - nL;
			s.left[node] = l;
		}
	};

	struct HoleKernel        // holes (left slots holding a "right" element) and sources ("left" elements in right slots)
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t p) const
		{
			const uint32_t node = s.nodeOfA[p];
			if (node < levelStart || !(s.state[node] & kStPartition)) return;
			const uint32_t f = s.first[node], nL = s.nL[node];
			const uint32_t rel = p - f;
			const uint32_t prefL = s.scan[p] - s.scan[f];
			const uint32_t isL = s.flags[p];
			if (rel < nL) { if (!isL) s.holes[f + (rel - prefL)] = p; }
			else { if (isL) s.srcs[f + (nL - prefL - 1)] = p; }
		}
	};

	struct ScatterKernel     // closed form of the swap loop (:238-251); also hands slots to the children
	{
		BuildState s; uint32_t levelStart;
		SPT_KERNEL_BODY void operator()(uint32_t p) const
		{
			const uint32_t node = s.nodeOfA[p];
			const uint32_t tri = s.idxA[p];
			if (node < levelStart || !(s.state[node] & kStPartition)) { s.idxB[p] = tri; s.nodeOfB[p] = node; return; }
			const uint32_t f = s.first[node], c = s.count[node], nL = s.nL[node];
			const uint32_t rel = p - f;
			const uint32_t prefL = s.scan[p] - s.scan[f];
			const uint32_t isL = s.flags[p];
			const uint32_t numHoles = nL - (s.scan[f + nL] - s.scan[f]);   // "right" elements among the first nL slots
			uint32_t dest;
			if (isL)
			{
				if (rel < nL) dest = p;                                     // stays
				else dest = s.holes[f + (nL - prefL - 1)];                  // k-th "left" from the right end fills the k-th hole
			}
			else
			{
				if (rel > nL) dest = p - 1;                                 // right region shifts down by one
				else
				{
					// a hole's element (rel < nL) or the pseudo-hole at slot nL: lands just below the previous source
					const uint32_t r = rel < nL ? (rel - prefL) : numHoles; // 0-based hole rank
					dest = r == 0 ? (f + c - 1) : (s.srcs[f + (r - 1)] - 1);
				}
			}
			s.idxB[dest] = tri;
			uint32_t owner = node;
			if (s.state[node] & kStSplit) owner = s.left[node] + ((dest - f) < nL ? 0u : 1u);
			s.nodeOfB[dest] = owner;
		}
	};

	// ---- renumbering into the reference's depth-first allocation order -----------------------------------------
	struct SubtreeKernel     // bottom-up: number of internal nodes in the subtree
	{
		BuildState s; uint32_t levelStart; uint32_t* internalCount;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			const uint32_t l = s.left[node];
			internalCount[node] = l ? (1u + internalCount[l] + internalCount[l + 1]) : 0u;
		}
	};

	struct RenumberKernel    // top-down: reference node index + pre-order rank among internal nodes
	{
		BuildState s; uint32_t levelStart; const uint32_t* internalCount; uint32_t* refIdx; uint32_t* rank;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t node = levelStart + i;
			const uint32_t l = s.left[node];
			if (!l) return;
			const uint32_t rk = rank[node];
			refIdx[l] = 1u + 2u * rk;          // m_nodesUsed++ twice, depth first (BVH.cpp:262-263, 276-277)
			refIdx[l + 1] = 2u + 2u * rk;
			rank[l] = rk + 1u;
			rank[l + 1] = rk + 1u + internalCount[l];
		}
	};

	struct LeafCountKernel   // leaf sizes in reference node order, for the re-layout offsets (:306-311)
	{
		BuildState s; const uint32_t* refIdx; uint32_t* leafCountByRef;
		SPT_KERNEL_BODY void operator()(uint32_t node) const { leafCountByRef[refIdx[node]] = s.left[node] ? 0u : s.count[node]; }
	};

	SPT_HD float SquareArea(V4 v0, V4 v1, V4 v2)
	{
		// Triangle::SquareArea (Bounds.cpp:494-502)
		const float a = length(v3(v0.x - v1.x, v0.y - v1.y, v0.z - v1.z));
		const float b = length(v3(v0.x - v2.x, v0.y - v2.y, v0.z - v2.z));
		const float c = length(v3(v2.x - v1.x, v2.y - v1.y, v2.z - v1.z));
		const float sp = (a + b + c) / 2;
		return sp * (sp - a) * (sp - b) * (sp - c);
	}

	struct EmitKernel        // reference BVHNode array + m_triIdxMapping (:298-337)
	{
		BuildState s; const uint32_t* refIdx; const uint32_t* leafOffsetByRef; const uint32_t* finalIdx;
		SailorPtBvhNode* refNodes; uint32_t* mapping; float* areaScratch;
		SPT_KERNEL_BODY void operator()(uint32_t node) const
		{
			const uint32_t ri = refIdx[node];
			const float* bb = s.aabb + (size_t)node * 6;
			SailorPtBvhNode out;
			out.aabbMin[0] = bb[0]; out.aabbMin[1] = bb[1]; out.aabbMin[2] = bb[2];
			out.aabbMax[0] = bb[3]; out.aabbMax[1] = bb[4]; out.aabbMax[2] = bb[5];
			const uint32_t l = s.left[node];
			if (l) { out.leftFirst = refIdx[l]; out.triCount = 0; refNodes[ri] = out; return; }
			const uint32_t f = s.first[node], c = s.count[node], off = leafOffsetByRef[ri];
			out.leftFirst = off; out.triCount = c;
			refNodes[ri] = out;
			// std::stable_sort by SquareArea descending == stable insertion sort with the same comparator
			for (uint32_t j = 0; j < c; j++)
			{
				const uint32_t tri = finalIdx[f + j];
				const float ar = SquareArea(s.vtx[tri * 3], s.vtx[tri * 3 + 1], s.vtx[tri * 3 + 2]);
				uint32_t k = j;
				while (k > 0 && ar > areaScratch[off + k - 1]) { areaScratch[off + k] = areaScratch[off + k - 1]; mapping[off + k] = mapping[off + k - 1]; k--; }
				areaScratch[off + k] = ar; mapping[off + k] = tri;
			}
		}
	};
}
