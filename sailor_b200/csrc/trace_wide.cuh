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
#include "trace_kernels.cuh"
#include "wide_bvh.cuh"

namespace spt
{
#if !defined(SPT_EMU)
#ifndef SPT_WIDE_BLOCK
#define SPT_WIDE_BLOCK 128
#endif
#ifndef SPT_WIDE_SMEM_STACK
#define SPT_WIDE_SMEM_STACK 12
#endif
#ifndef SPT_WIDE_TRI_REPS
#define SPT_WIDE_TRI_REPS 2
#endif
#ifndef SPT_WIDE_MIN_BLOCKS
#define SPT_WIDE_MIN_BLOCKS 5
#endif
	constexpr int kWideBlock = SPT_WIDE_BLOCK;
	constexpr int kWideSmemStack = SPT_WIDE_SMEM_STACK;

	struct ReplayOut { uint32_t* list; uint32_t* count; };

	template<class Source, class Sink>
	__device__ __forceinline__ void TraceWideLoop(const WideView& w, uint32_t n, uint32_t* __restrict__ counter, uint2* stackMem, const ReplayOut& replay, Source& src, Sink& sink)
	{
		const uint32_t sAddr = (uint32_t)__cvta_generic_to_shared(stackMem) + threadIdx.x * 8u;
		uint2 ovf[kWideStackDepth - kWideSmemStack];
		const uint32_t lane = threadIdx.x & 31;
		V3 o = v3(0.0f), d = v3(0.0f);
		WideRay r; r.o = v3(0.0f); r.idir = v3(0.0f); r.octinv = 0;
		WideBest best; best.Reset();
		uint32_t ignore = kNoHit, index = 0;
		bool anyHit = false, active = false, bad = false;
		uint32_t gBase = 0, gBits = 0, tBase = 0, tBits = 0;
		int sp = 0;
		bool exhausted = false;

		auto push = [&](uint32_t x, uint32_t y)
		{
			if (sp < kWideSmemStack) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sAddr + (uint32_t)sp * (kWideBlock * 8u)), "r"(x), "r"(y) : "memory");
			else if (sp < kWideStackDepth) ovf[sp - kWideSmemStack] = make_uint2(x, y);
			else bad = true;                                   // the walk is abandoned below
			sp++;
		};

		for (;;)
		{
			// ---- refill idle lanes ----
			const uint32_t idleMask = __ballot_sync(0xffffffffu, !active);
			if (idleMask)
			{
				if (!exhausted && (__popc(idleMask) >= (int)kFetchMinIdle))
				{
					const uint32_t want = (uint32_t)__popc(idleMask);
					uint32_t base = 0;
					if (lane == 0) base = atomicAdd(counter, want);
					base = __shfl_sync(0xffffffffu, base, 0);
					if (base + want >= n) exhausted = true;
					bool toReplay = false; uint32_t replayIndex = 0;
					if (!active)
					{
						const uint32_t i = base + (uint32_t)__popc(idleMask & ((1u << lane) - 1u));
						float maxLen;
						if (i < n && src.Load(i, o, d, ignore, maxLen, anyHit))
						{
							const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
							if (!WideSafe(o, rD)) { toReplay = true; replayIndex = i; }
							else
							{
								index = i; active = true; bad = false;
								r.o = o; r.idir = rD;
								r.octinv = 7u - ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
								best.Reset();
								sp = 0; tBits = 0;
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
				if (idleMask == 0xffffffffu) break;             // queue exhausted and every lane retired
			}
			// ---- vote ----
			const bool hasNode = active && (gBits & 0xFFu) != 0u;
			const bool hasTri = active && tBits != 0u;
			const int nNode = __popc(__ballot_sync(0xffffffffu, hasNode)), nTri = __popc(__ballot_sync(0xffffffffu, hasTri));
			bool finished = false;
			if (nNode >= nTri)
			{
				if (hasNode)
				{
					if (tBits) { push(tBase | kTriGroupTag, tBits); tBits = 0; }        // postpone the triangles
					const uint32_t pos = 31u - (uint32_t)__clz((int)(gBits & 0xFFu));
					gBits ^= 1u << pos;
					const uint32_t slot = pos ^ r.octinv;
					const uint32_t child = gBase + (uint32_t)__popc((gBits >> 8) & ((1u << slot) - 1u));
					if (gBits & 0xFFu) push(gBase, gBits);
					WideNodeTest(w.nodes + child, r, best.limit, gBase, gBits, tBase, tBits);
				}
			}
			else
			{
#pragma unroll 1
				for (int rep = 0; rep < SPT_WIDE_TRI_REPS; rep++)
				{
					if (active && tBits)
					{
						const uint32_t i = (uint32_t)__ffs((int)tBits) - 1u; tBits &= tBits - 1u;
						const TTri* T = w.tris + (tBase + i);
						const V4 a = ld4(&T->a), b = ld4(&T->b), c = ld4(&T->c);
						const uint32_t triId = f2u(c.y);
						float t, u, v;
						if (triId != ignore && TriTest(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), kFltMax, t, u, v) && t <= best.limit)
						{
							// rare path: the reference's slab test of the triangle's own leaf (WideCandidate)
							const V4 b0 = ld4(w.leafBox + (size_t)(tBase + i) * 2), b1 = ld4(w.leafBox + (size_t)(tBase + i) * 2 + 1);
							if (SlabTest(o, r.idir, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, kFltMax) != kFltMax)
							{
								best.Offer(t, u, v, triId);
								if (anyHit) { sp = 0; tBits = 0; gBits = 0; }                 // hit-or-miss query: done
							}
						}
					}
				}
			}
			// ---- nothing left in hand: next group from the stack, or the walk is over ----
			if (active && tBits == 0u && (gBits & 0xFFu) == 0u)
			{
				if (sp == 0 || bad) finished = true;
				else
				{
					sp--;
					uint32_t x, y;
					if (sp < kWideSmemStack) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(sAddr + (uint32_t)sp * (kWideBlock * 8u)) : "memory");
					else { x = ovf[sp - kWideSmemStack].x; y = ovf[sp - kWideSmemStack].y; }
					if (x & kTriGroupTag) { tBase = x & ~kTriGroupTag; tBits = y; }
					else { gBase = x; gBits = y; }
				}
			}
			else if (active && bad) { finished = true; }
			// ---- retire ----
			const bool toReplay = finished && (bad || (!anyHit && best.tie));
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
			Hit h; h.t = best.t; h.u = best.u; h.v = best.v; h.tri = best.tri;
			sink.Retire(finished && !toReplay, index, h, anyHit);
			if (finished) { active = false; gBits = 0; tBits = 0; sp = 0; }
		}
	}

	// exact replay: the rays on the replay list through the reference-visit-order warp loop
	template<class Inner>
	struct ReplaySource
	{
		const uint32_t* list; Inner inner;
		__device__ __forceinline__ bool Load(uint32_t i, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) const { return inner.Load(list[i], o, d, ignore, maxLen, anyHit); }
	};
	template<class Inner>
	struct ReplaySink
	{
		const uint32_t* list; Inner inner;
		__device__ __forceinline__ void Retire(bool finished, uint32_t i, const Hit& h, bool anyHit) const { inner.Retire(finished, finished ? list[i] : 0u, h, anyHit); }
	};

	__global__ void __launch_bounds__(kWideBlock, SPT_WIDE_MIN_BLOCKS) k_trace_wide_rays(WideView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, ReplayOut replay)
	{
		__shared__ uint2 stackMem[kWideSmemStack * kWideBlock];
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; QueueSink sink{ hits };
		TraceWideLoop(w, n, counter, stackMem, replay, src, sink);
	}
	__global__ void __launch_bounds__(kWideBlock, SPT_WIDE_MIN_BLOCKS) k_trace_wide_level(WideView w, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, WavefrontOut out, ReplayOut replay)
	{
		__shared__ uint2 stackMem[kWideSmemStack * kWideBlock];
		{ const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; WavefrontSink sink{ hits, out.status + *out.auxBase, out.slowList, out.slowCount };
		TraceWideLoop(w, n, counter, stackMem, replay, src, sink);
	}
	__global__ void __launch_bounds__(kTraceBlock) k_replay_rays(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		const uint32_t* __restrict__ list, const uint32_t* __restrict__ nPtr, uint32_t cap, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		uint32_t n = *nPtr; if (n > cap) n = cap;
		if (!n) return;
		ReplaySource<QueueSource> src{ list, QueueSource{ rays } }; ReplaySink<QueueSink> sink{ list, QueueSink{ hits } };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
	}
	__global__ void __launch_bounds__(kTraceBlock) k_replay_level(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		const uint32_t* __restrict__ list, const uint32_t* __restrict__ nPtr, uint32_t cap, uint32_t* __restrict__ counter, WavefrontOut out)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		uint32_t n = *nPtr; if (n > cap) n = cap;
		if (!n) return;
		ReplaySource<QueueSource> src{ list, QueueSource{ rays } };
		ReplaySink<WavefrontSink> sink{ list, WavefrontSink{ hits, out.status + *out.auxBase, out.slowList, out.slowCount } };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
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

	// Replay bookkeeping of one scene: [0] wide work counter, [1] replay count, [2] replay work counter, [3] total replayed (stats)
	struct WideTraceBuffers { uint32_t* counters; uint32_t* replayList; uint32_t replayCap; };

#if !defined(SPT_EMU)
	struct AccumulateReplayKernel { uint32_t* c; SPT_KERNEL_BODY void operator()(uint32_t) const { c[3] += c[1]; } };

	// closest hits / hit-or-miss for a ray queue through the wide layout + exact replay (QueueSink: hits[i] for every ray)
	inline void LaunchTraceRaysWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const WideTraceBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
	{
		if (!n || !ctx.ok) return;
		DevMemset(ctx, b.counters, 0, 3 * sizeof(uint32_t));
		k_trace_wide_rays<<<WideGridSize(), kWideBlock, 0, ctx.stream>>>(w, rays, hits, n, nPtr, b.counters, ReplayOut{ b.replayList, b.counters + 1 });
		k_replay_rays<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, b.replayList, b.counters + 1, b.replayCap, b.counters + 2);
		launch_for(ctx, 1, AccumulateReplayKernel{ b.counters });
		ctx.kernelLaunches += 2;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
	inline void LaunchTraceLevelWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const WideTraceBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
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
	inline void LaunchTraceRaysWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const WideTraceBuffers& b, const RayRec* rays, Hit* hits, uint32_t n, const uint32_t* nPtr = nullptr)
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
	inline void LaunchTraceLevelWide(Ctx& ctx, const WideView& w, const BvhView& bvh, const WideTraceBuffers& b, const RayRec* rays, Hit* hits, uint32_t cap, const uint32_t* nPtr, const WavefrontOut& out)
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
			if (h.tri != kNoHit && !(rays[i].tmax < 0.0f)) { hits[i] = h; out.slowList[(*out.slowCount)++] = i; }
		}
		ctx.kernelLaunches += 3;
	}
#endif
}
