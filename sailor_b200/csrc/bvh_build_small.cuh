// bvh_build_small.cuh — the level-synchronous BVH build of bvh_build.cuh run by ONE thread block in ONE launch.
//
// A scene of a few thousand triangles (BASELINE configs C1/C2 are a 12-triangle cube) gives the multi-launch build nothing
// to parallelise: it spends its time in ~20 launches and one host read-back per tree level (99 launches + 6 syncs, 0.5 ms,
// for the cube — 6 % of a whole C2 frame, profiles/r01g_SUMMARY.md).  Here the same per-level functors (same code, hence
// the same bits: BoundsKernel, PrepareKernel, BinKernel, SplitKernel, FlagKernel, CountKernel, AllocKernel, HoleKernel,
// ScatterKernel, then Subtree / Renumber / LeafCount / Emit and the traversal-layout pack) are called from a single CTA
// with __syncthreads() between the phases; the level loop, the two prefix sums per level and the final renumbering run
// on the device and the host reads three words once.
#pragma once
#include "bvh_build.cuh"
#include "traverse.cuh"

namespace spt
{
	constexpr uint32_t kSmallBuildMax = 4096;       // triangles
	constexpr int kSmallBuildBlock = 1024;
	constexpr uint32_t kSmallBuildLevels = 128;     // deeper trees (degenerate inputs) fall back to the multi-launch build

	struct SmallBuildOut
	{
		uint32_t* internalCount; uint32_t* refIdx; uint32_t* rank; uint32_t* leafCountByRef; uint32_t* leafOffsetByRef; uint32_t* leafCountAtSlot;
		float* areaScratch; SailorPtBvhNode* refNodes; uint32_t* mapping; TNode* tnodes; TTri* ttris;
		uint32_t* result;                             // [0] nodesUsed [1] numInternal [2] numLevels [3] 1 = too deep, nothing usable was produced
		uint32_t maxNodes;
	};

#if !defined(SPT_EMU)
	template<class F>
	__device__ __forceinline__ void BlockFor(uint32_t n, const F& f)
	{
		for (uint32_t i = threadIdx.x; i < n; i += kSmallBuildBlock) f(i);
		__syncthreads();
	}
	__device__ __forceinline__ void BlockZero(uint32_t* p, uint32_t n)
	{
		for (uint32_t i = threadIdx.x; i < n; i += kSmallBuildBlock) p[i] = 0u;
		__syncthreads();
	}
	// out[i] = sum(in[0..i)), out[n] = total; in / out may not alias (the contract of ExclusiveScanU32)
	__device__ __forceinline__ void BlockScan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* sWarp)
	{
		const uint32_t per = (n + kSmallBuildBlock - 1u) / kSmallBuildBlock;
		const uint32_t b = threadIdx.x * per, e = b + per < n ? b + per : n;
		uint32_t sum = 0;
		for (uint32_t i = b; i < e; i++) sum += in[i];
		const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
		if (lane == 31u) sWarp[warp] = incl;
		__syncthreads();
		if (warp == 0)
		{
			uint32_t w = sWarp[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if ((int)lane >= o) w += t; }
			sWarp[lane] = w;
		}
		__syncthreads();
		uint32_t run = (warp ? sWarp[warp - 1u] : 0u) + incl - sum;
		for (uint32_t i = b; i < e; i++) { out[i] = run; run += in[i]; }
		if (threadIdx.x == kSmallBuildBlock - 1) out[n] = sWarp[31];
		__syncthreads();
	}

	__global__ void __launch_bounds__(kSmallBuildBlock) k_build_small(BuildState s, SmallBuildOut o)
	{
		__shared__ uint32_t sWarp[32];
		__shared__ uint32_t sLevelStart[kSmallBuildLevels], sLevelCount[kSmallBuildLevels];
		const uint32_t N = s.n;
		BlockFor(N, InitSlotsKernel{ s });
		if (threadIdx.x == 0) { s.first[0] = 0u; s.count[0] = N; }                      // BVH.cpp:291-293
		__syncthreads();

		uint32_t start = 0, cnt = 1, levels = 0;
		while (cnt)
		{
			if (levels == kSmallBuildLevels) { if (threadIdx.x == 0) o.result[3] = 1u; return; }
			if (threadIdx.x == 0) { sLevelStart[levels] = start; sLevelCount[levels] = cnt; *s.binCounter = 0u; }
			levels++;
			__syncthreads();
			BlockFor(cnt, InitNodesKernel{ s, start });
			BlockFor(N, BoundsKernel{ s, start });
			BlockFor(cnt, PrepareKernel{ s, start });
			BlockFor(*s.binCounter * kNodeBinWords, BinInitKernel{ s, 0ull });
			BlockFor(N, BinKernel{ s, start });
			BlockFor(cnt, SplitKernel{ s, start });
			BlockFor(N, FlagKernel{ s, start });
			BlockScan(s.flags, s.scan, N, sWarp);
			BlockFor(cnt, CountKernel{ s, start });
			BlockScan(s.splitFlag, s.splitScan, cnt, sWarp);
			BlockFor(cnt, AllocKernel{ s, start, start + cnt });
			BlockFor(N, HoleKernel{ s, start });
			BlockFor(N, ScatterKernel{ s, start });
			const uint32_t numSplit = s.splitScan[cnt];
			__syncthreads();                                                              // everyone has read numSplit before the next level rewrites splitScan
			{ uint32_t* t = s.idxA; s.idxA = s.idxB; s.idxB = t; t = s.nodeOfA; s.nodeOfA = s.nodeOfB; s.nodeOfB = t; }
			start += cnt; cnt = 2u * numSplit;
		}
		const uint32_t nodesUsed = start;

		// renumber into the reference's allocation order and emit both layouts (pipeline.cuh, BuildBvh)
		for (uint32_t i = threadIdx.x; i < o.maxNodes; i += kSmallBuildBlock) { SailorPtBvhNode z; memset(&z, 0, sizeof(z)); o.refNodes[i] = z; }
		BlockZero(o.refIdx, nodesUsed); BlockZero(o.rank, nodesUsed); BlockZero(o.leafCountAtSlot, N);
		for (uint32_t l = levels; l-- > 0;) BlockFor(sLevelCount[l], SubtreeKernel{ s, sLevelStart[l], o.internalCount });
		for (uint32_t l = 0; l < levels; l++) BlockFor(sLevelCount[l], RenumberKernel{ s, sLevelStart[l], o.internalCount, o.refIdx, o.rank });
		BlockFor(nodesUsed, LeafCountKernel{ s, o.refIdx, o.leafCountByRef });
		BlockScan(o.leafCountByRef, o.leafOffsetByRef, nodesUsed, sWarp);
		BlockFor(nodesUsed, EmitKernel{ s, o.refIdx, o.leafOffsetByRef, s.idxA, o.refNodes, o.mapping, o.areaScratch });
		BlockFor(nodesUsed, LeafCountAtSlotKernel{ s.left, s.count, o.refIdx, o.leafOffsetByRef, o.leafCountAtSlot });
		BlockFor(nodesUsed, PackNodesKernel{ s.left, o.rank, o.refIdx, o.leafOffsetByRef, s.aabb, o.tnodes });
		BlockFor(N, PackTrisKernel{ s.vtx, o.mapping, o.leafCountAtSlot, o.ttris, N });
		if (threadIdx.x == 0) { o.result[0] = nodesUsed; o.result[1] = o.internalCount[0]; o.result[2] = levels; o.result[3] = 0u; }
	}
#endif
}
