// trace_wide.cuh — traversal of the wide layout (wide_bvh.cuh) by persistent warps, plus the exact replay pass.
//
// Same skeleton as TraceWarpLoop (trace_kernels.cuh): the grid is sized to the machine, every lane owns one ray, idle lanes
// are refilled from the queue with one atomic per refill, and every iteration the warp votes for the step most of its busy
// lanes need.  The steps differ:
//   NODE step   the lane takes the next child of its current node group (hits of one node that are still to be visited, in
//               front-to-back order), fetches that 80-byte node with five LDG.128 and tests its eight quantised child boxes:
//               per child six PRMT (byte -> biased float), six FFMA and two FMNMX3 + two FMNMX.  The result is a new node
//               group (inner children hit) and a triangle group (one bit per triangle record of the leaf slots hit).
//   TRI step    the lane tests the next triangle of its triangle group (three LDG.128, the reference's own Moller-Trumbore).
// A lane that still has triangles when the warp votes NODE parks its triangle group on the stack and joins the node step
// (postponing, Ylitie et al. 2017), so node steps run with every lane that has a node left.
//
// The per-lane stack holds 8-byte groups, [entry][thread] in shared memory, deeper entries in local memory.
//
// Results: hit-or-miss queries stop at the first candidate that the reference could reach (see WideCandidate).  Closest-hit
// queries keep the nearest candidate and raise `tie` when a second candidate lies within the band; those rays, rays with a
// non-finite reciprocal direction and walks that exhaust the stack are appended to the REPLAY list and traced afterwards by the
// exact kernel (reference visit order), so every result handed on equals the reference's.
#pragma once
#include "trace_fast.cuh"

namespace spt
{
#if !defined(SPT_EMU)
#ifndef SPT_WIDE_BLOCK
#define SPT_WIDE_BLOCK 128
#endif
#ifndef SPT_WIDE_NODE_STACK
#define SPT_WIDE_NODE_STACK 10     // node-group entries per lane in shared memory (deeper ones spill to local memory)
#endif
#ifndef SPT_WIDE_TRI_STACK
#define SPT_WIDE_TRI_STACK 6       // parked triangle groups per lane (a lane whose stack is full forces a triangle step)
#endif
#ifndef SPT_WIDE_TRI_REPS
#define SPT_WIDE_TRI_REPS 2
#endif
#ifndef SPT_WIDE_TRI_VOTE
#define SPT_WIDE_TRI_VOTE 20       // a triangle step runs when at least this many lanes have a triangle waiting
#endif
#ifndef SPT_WIDE_MIN_BLOCKS
#define SPT_WIDE_MIN_BLOCKS 6
#endif
#ifndef SPT_WIDE_FETCH_MIN_IDLE
#define SPT_WIDE_FETCH_MIN_IDLE 12
#endif
	constexpr int kWideBlock = SPT_WIDE_BLOCK;
	constexpr int kWideNodeSmem = SPT_WIDE_NODE_STACK, kWideTriSmem = SPT_WIDE_TRI_STACK;
	constexpr int kWideSmemEntries = kWideNodeSmem + kWideTriSmem;

#if defined(SPT_WIDE_LOOP_STATS)
	// tuning aid (tools/wide_variants.py "stats"): sums over every vote of every warp
	// 0 votes, 1 idle lanes, 2 lanes with node work, 3 lanes with triangle work, 4 node steps, 5 lanes in node steps, 6 triangle reps, 7 lanes in triangle reps,
	// 8 refills, 9 lanes refilled, 10 forced triangle steps, 11 rays retired, 12 triangle groups parked, 13 node groups parked
	__device__ unsigned long long g_wideLoopStats[16];
#define SPT_WL(i, v) do { if (lane == 0) wl_[i] += (v); } while (0)
#define SPT_WL_LANES(i, cond) do { const uint32_t m_ = __ballot_sync(0xffffffffu, (cond)); if (lane == 0) wl_[i] += __popc(m_); } while (0)
#else
#define SPT_WL(i, v) do { } while (0)
#define SPT_WL_LANES(i, cond) do { } while (0)
#endif

	// Lane state: a node group and a triangle group "in hand" plus TWO stacks, one of parked node groups and one of parked
	// triangle groups.  Because the two kinds of work are kept apart, a lane can take part in a node step as long as it has any
	// node left and in a triangle step as long as it has any triangle left, whatever it produced last: the warp is not split by
	// what each lane happened to find in its last node.  Triangles go first when enough lanes have some (they end hit-or-miss
	// walks and shrink the ray of closest-hit walks); the order never changes a result (WideBest).
	template<class Source, class Sink>
	__device__ __forceinline__ void TraceWideLoop(const WideView& w, uint32_t n, uint32_t* __restrict__ counter, uint2* stackMem, const ReplayOut& replay, Source& src, Sink& sink)
	{
		const uint32_t sNode = (uint32_t)__cvta_generic_to_shared(stackMem) + threadIdx.x * 8u;
		const uint32_t sTri = sNode + (uint32_t)kWideNodeSmem * (kWideBlock * 8u);
		uint2 ovf[kWideStackDepth - kWideNodeSmem];
		const uint32_t lane = threadIdx.x & 31;
		V3 d = v3(0.0f), o0 = v3(0.0f);      // o0: the ray's origin as queued (the SlowRec of a closest hit carries it to ClassifyKernel)
		WideRay r; r.o = v3(0.0f); r.idir = v3(0.0f); r.signs = 0; r.octinv = 0;
		WideBest best; best.Reset();
		uint32_t ignore = kNoHit, index = 0;
		bool anyHit = false, active = false, bad = false;
		uint32_t gBase = 0, gBits = 0, tBase = 0, tBits = 0;
		int nsp = 0, tsp = 0;
		bool exhausted = false;
#if defined(SPT_WIDE_LOOP_STATS)
		unsigned long long wl_[14] = {};
#endif

		auto pushNode = [&](uint32_t x, uint32_t y)
		{
			if (nsp < kWideNodeSmem) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sNode + (uint32_t)nsp * (kWideBlock * 8u)), "r"(x), "r"(y) : "memory");
			else if (nsp < kWideStackDepth) ovf[nsp - kWideNodeSmem] = make_uint2(x, y);
			else bad = true;                                   // the walk is abandoned and replayed
			nsp++;
		};
		auto popNode = [&](uint32_t& x, uint32_t& y)
		{
			nsp--;
			if (nsp < kWideNodeSmem) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sNode + (uint32_t)nsp * (kWideBlock * 8u)) : "memory");
			else { x = ovf[nsp - kWideNodeSmem].x; y = ovf[nsp - kWideNodeSmem].y; }
		};

		for (;;)
		{
			// ---- refill idle lanes ----
			const uint32_t idleMask = __ballot_sync(0xffffffffu, !active);
			if (idleMask)
			{
				if (!exhausted && (__popc(idleMask) >= SPT_WIDE_FETCH_MIN_IDLE))
				{
					const uint32_t want = (uint32_t)__popc(idleMask);
					uint32_t base = 0;
					if (lane == 0) base = atomicAdd(counter, want);
					base = __shfl_sync(0xffffffffu, base, 0);
					if (base + want >= n) exhausted = true;
					SPT_WL(8, 1); SPT_WL(9, want);
					bool toReplay = false; uint32_t replayIndex = 0;
					if (!active)
					{
						const uint32_t i = base + (uint32_t)__popc(idleMask & ((1u << lane) - 1u));
						float maxLen; V3 o;
						if (i < n && src.Load(i, o, d, ignore, maxLen, anyHit))
						{
							const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
							if (!WideSafe(o, rD)) { toReplay = true; replayIndex = i; }
							else
							{
								index = i; active = true; bad = false; o0 = o;
								r.Set(o, d, rD, !anyHit);
								best.Reset();
								nsp = 0; tsp = 0; tBits = 0;
								// a fresh walk is a node group that holds only the root: base 0, one hit at the top position, an imask with
								// that one slot set (so the child index is base + 0)
								gBase = 0; gBits = 0x80u | ((1u << (7u ^ r.octinv)) << 8);
							}
						}
					}
					// rays the wide walk does not take go straight to the replay list (warp-aggregated append)
					const uint32_t rm = __ballot_sync(0xffffffffu, toReplay);
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
				if (idleMask == 0xffffffffu)                    // queue exhausted and every lane retired
				{
#if defined(SPT_WIDE_LOOP_STATS)
					if (lane == 0) for (int k = 0; k < 14; k++) atomicAdd(&g_wideLoopStats[k], wl_[k]);
#endif
					break;
				}
			}
			// ---- vote ----
			const bool triFull = tsp >= kWideTriSmem;                                  // cannot park another triangle group
			const bool hasNode = active && ((gBits & 0xFFu) != 0u || nsp > 0) && !(triFull && tBits != 0u);
			const bool hasTri = active && (tBits != 0u || tsp > 0);
			const uint32_t nodeMask = __ballot_sync(0xffffffffu, hasNode), triMask = __ballot_sync(0xffffffffu, hasTri);
			const bool force = __any_sync(0xffffffffu, active && triFull);
			bool finished = false;
			SPT_WL(0, 1); SPT_WL(1, __popc(idleMask)); SPT_WL(2, __popc(nodeMask)); SPT_WL(3, __popc(triMask)); SPT_WL(10, force ? 1 : 0);
			if (!(force || nodeMask == 0u || __popc(triMask) >= SPT_WIDE_TRI_VOTE))
			{
				SPT_WL(4, 1); SPT_WL(5, __popc(nodeMask));
				if (hasNode)
				{
					if ((gBits & 0xFFu) == 0u) popNode(gBase, gBits);
					const uint32_t pos = 31u - (uint32_t)__clz((int)(gBits & 0xFFu));
					gBits ^= 1u << pos;
					const uint32_t slot = pos ^ r.octinv;
					const uint32_t child = gBase + (uint32_t)__popc((gBits >> 8) & ((1u << slot) - 1u));
					if (gBits & 0xFFu) pushNode(gBase, gBits);
					uint32_t nb, nt;
					WideNodeTest(w.nodes + child, r, best.limit, gBase, gBits, nb, nt);
					if (nt)
					{
						if (tBits)
						{
							asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sTri + (uint32_t)tsp * (kWideBlock * 8u)), "r"(tBase), "r"(tBits) : "memory");
							tsp++;
						}
						tBase = nb; tBits = nt;
					}
				}
			}
			else
			{
#pragma unroll 1
				for (int rep = 0; rep < SPT_WIDE_TRI_REPS; rep++)
				{
					SPT_WL(6, 1); SPT_WL_LANES(7, active && (tBits != 0u || tsp > 0));
					if (active && (tBits != 0u || tsp > 0))
					{
						if (tBits == 0u)
						{
							tsp--;
							asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tBase), "=r"(tBits) : "r"(sTri + (uint32_t)tsp * (kWideBlock * 8u)) : "memory");
						}
						const uint32_t i = (uint32_t)__ffs((int)tBits) - 1u; tBits &= tBits - 1u;
						const TTri* T = w.tris + (tBase + i);
						const V4 a = ld4(&T->a), b = ld4(&T->b), c = ld4(&T->c);
						const uint32_t triId = f2u(c.y);
						float t, u, v;
						if (triId != ignore && TriTest(r.o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), kFltMax, t, u, v))
						{
							best.Offer(t, u, v, triId, tBase + i);
							if (anyHit) { nsp = 0; tsp = 0; tBits = 0; gBits = 0; }           // hit-or-miss query: done
						}
					}
				}
			}
			// ---- the walk is over when nothing is in hand and both stacks are empty ----
			bool toReplay = false;
			if (active && (bad || (tBits == 0u && tsp == 0 && (gBits & 0xFFu) == 0u && nsp == 0)))
			{
				finished = true;
				toReplay = bad || (!anyHit && best.tie);
				// the winner of a closest-hit walk must be reachable through its own leaf box in the reference's tree (wide_bvh.cuh)
				if (!toReplay && !anyHit && best.tri != kNoHit) toReplay = !WideWinnerReachable(w, best.rec, r.o, r.idir);
			}
			// ---- retire ----
			{
				const uint32_t rm = __ballot_sync(0xffffffffu, toReplay);
				if (rm)
				{
					const int leader = __ffs(rm) - 1;
					uint32_t rb = 0;
					if ((int)lane == leader) rb = atomicAdd(replay.count, (uint32_t)__popc(rm));
					rb = __shfl_sync(0xffffffffu, rb, leader);
					if (toReplay) replay.list[rb + (uint32_t)__popc(rm & ((1u << lane) - 1u))] = index;
				}
			}
			SPT_WL_LANES(11, finished);
			Hit h; h.t = best.t; h.u = best.u; h.v = best.v; h.tri = best.tri;
			sink.Retire(finished && !toReplay, index, h, anyHit, o0, d);
			if (finished) { active = false; gBits = 0; tBits = 0; nsp = 0; tsp = 0; }
		}
	}

	__global__ void __launch_bounds__(kWideBlock, SPT_WIDE_MIN_BLOCKS) k_trace_wide_rays(WideView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, ReplayOut replay)
	{
		__shared__ uint2 stackMem[kWideSmemEntries * kWideBlock];
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; QueueSink sink{ hits };
		TraceWideLoop(w, n, counter, stackMem, replay, src, sink);
	}
	__global__ void __launch_bounds__(kWideBlock, SPT_WIDE_MIN_BLOCKS) k_trace_wide_level(WideView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, WavefrontOut out, ReplayOut replay)
	{
		__shared__ uint2 stackMem[kWideSmemEntries * kWideBlock];
		{ const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; WavefrontSink sink{ out.status + *out.auxBase, out.slow, out.slowCount };
		TraceWideLoop(w, n, counter, stackMem, replay, src, sink);
	}
	inline int WideGridSize()
	{
		static int grid = 0;
		if (!grid)
		{
			int dev = 0, sms = 148, perSm = 1;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace_wide_level, kWideBlock, 0);
			grid = sms * (perSm > 0 ? perSm : 1);
		}
		return grid;
	}
#endif

#if !defined(SPT_EMU)
	// closest hits / hit-or-miss for a ray queue through the wide layout + exact replay (QueueSink: hits[i] for every ray)
	inline void LaunchTraceRaysWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
	{
		if (!n || !ctx.ok) return;
		DevMemset(ctx, b.counters, 0, 3 * sizeof(uint32_t));
		k_trace_wide_rays<<<WideGridSize(), kWideBlock, 0, ctx.stream>>>(w, rays, hits, n, nPtr, b.counters, ReplayOut{ b.replayList, b.counters + 1 });
		k_replay_rays<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, b.replayList, b.counters + 1, b.replayCap, b.counters + 2);
		launch_for(ctx, 1, AccumulateReplayKernel{ b.counters });
		ctx.kernelLaunches += 2;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
	inline void LaunchTraceLevelWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
	{
		if (!cap || !ctx.ok) return;
		DevMemset(ctx, b.counters, 0, 3 * sizeof(uint32_t));
		k_trace_wide_level<<<WideGridSize(), kWideBlock, 0, ctx.stream>>>(w, rays, hits, cap, nPtr, b.counters, out, ReplayOut{ b.replayList, b.counters + 1 });
		k_replay_level<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, b.replayList, b.counters + 1, b.replayCap, b.counters + 2, out);
		launch_for(ctx, 1, AccumulateReplayKernel{ b.counters });
		ctx.kernelLaunches += 2;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	inline void LaunchTraceRaysWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
	{
		LocalStack st;
		if (nPtr && *nPtr < n) n = *nPtr;
		for (uint32_t i = 0; i < n; i++)
		{
			if (rays[i].tmax == -1.0f) continue;
			const V3 o = v3(rays[i].ox, rays[i].oy, rays[i].oz), d = v3(rays[i].dx, rays[i].dy, rays[i].dz);
			if (!TraceWide(w, o, d, rays[i].ignoreTri, rays[i].tmax < 0.0f, hits[i]))
			{
				TraceClosest(bvh, o, d, rays[i].ignoreTri, fabsf(rays[i].tmax), st, hits[i]);
				b.counters[3]++;
			}
		}
		ctx.kernelLaunches += 3;
	}
	inline void LaunchTraceLevelWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const ReplayBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
	{
		LocalStack st;
		const uint32_t n = *nPtr < cap ? *nPtr : cap;
		for (uint32_t i = 0; i < n; i++)
		{
			if (rays[i].tmax == -1.0f) continue;
			const V3 o = v3(rays[i].ox, rays[i].oy, rays[i].oz), d = v3(rays[i].dx, rays[i].dy, rays[i].dz);
			Hit h;
			if (!TraceWide(w, o, d, rays[i].ignoreTri, rays[i].tmax < 0.0f, h))
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
