// trace_kernels.cuh — launchers of the traversal stage.
//
// Device form: PERSISTENT WARPS WITH DYNAMIC RAY FETCH.  The grid is sized to the machine (SMs x resident CTAs).
// Every lane owns one ray and is in one of three modes: idle, at an INNER node, or inside a LEAF (one triangle per
// step).  Each iteration the warp votes (ballot) and executes the step the majority of its busy lanes needs, so an
// instruction is never issued for a handful of lanes while the rest wait in the other branch; lanes that finished
// are refilled from the global queue as soon as half of the warp is idle (one atomicAdd per refill), so the
// warp does not drain to its slowest ray.  Per lane the order of box and triangle tests is exactly the
// reference's (BVH.cpp:122-191) whatever the warp does, so hit ids and barycentrics stay bit-identical.
//
// The per-lane traversal stack keeps its first kSmemStack entries in shared memory, laid out [entry][thread] so the
// 32 lanes of a warp hit 32 different banks, and spills deeper entries to local memory (the reference allows 64;
// a 1M-triangle tree needs ~22).  Keeping the shared part small leaves the SM's L1 for nodes and triangles and lets
// ~40 warps per SM hide the L2 latency of the dependent node fetches.  Node and triangle records are fetched as
// 128-bit loads through the read-only path (see traverse.cuh for the layouts).
//
// Slab test: CUDA's FMNMX drops NaNs while the reference's SSE/glm min/max propagate them in operand order
// (SURVEY H2).  NaNs can only arise from 0 * inf, i.e. when a component of 1/dir is not finite; rays whose
// reciprocal direction and origin are finite take the FMNMX form (bit-identical results on non-NaN inputs),
// all others take the reference's compare-select form.
#pragma once
#include "traverse.cuh"

namespace spt
{
	// ray queue record: 32 bytes in, 16 bytes out ("48 B per ray")
	struct alignas(16) RayRec { float ox, oy, oz; uint32_t ignoreTri; float dx, dy, dz; float tmax; };
	static_assert(sizeof(RayRec) == 32 && sizeof(Hit) == 16, "ray queue layout");

#if !defined(SPT_EMU)
	// tuning knobs (tools/trace_variants.py builds variants with -D to measure them on the GPU)
#ifndef SPT_TRACE_BLOCK
#define SPT_TRACE_BLOCK 128
#endif
#ifndef SPT_SMEM_STACK
#define SPT_SMEM_STACK 24
#endif
#ifndef SPT_FETCH_MIN_IDLE
#define SPT_FETCH_MIN_IDLE 16
#endif
#ifndef SPT_VOTE_LEAF_BIAS
#define SPT_VOTE_LEAF_BIAS 1
#endif
#ifndef SPT_INNER_REPS
#define SPT_INNER_REPS 4
#endif
#ifndef SPT_LEAF_REPS
#define SPT_LEAF_REPS 2
#endif
	constexpr int kTraceBlock = SPT_TRACE_BLOCK;
	constexpr int kSmemStack = SPT_SMEM_STACK;
	constexpr uint32_t kFetchMinIdle = SPT_FETCH_MIN_IDLE;

	struct SmemStack     // used by the one-ray-per-thread helpers (TraceClosest); whole stack in shared memory
	{
		uint32_t* base; int n;     // base already offset by threadIdx.x; stride = blockDim.x
		__device__ __forceinline__ void clear() { n = 0; }
		__device__ __forceinline__ bool empty() const { return n == 0; }
		__device__ __forceinline__ void push(uint32_t v) { base[n * kTraceBlock] = v; n++; }
		__device__ __forceinline__ uint32_t pop() { n--; return base[n * kTraceBlock]; }
	};

	// Slab test on non-NaN operands: same values as SlabTest, min/max as single FMNMX instructions.
	__device__ __forceinline__ float SlabTestFast(V3 o, V3 rD, float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float maxLen)
	{
		const float t1x = (bminx - o.x) * rD.x, t1y = (bminy - o.y) * rD.y, t1z = (bminz - o.z) * rD.z;
		const float t2x = (bmaxx - o.x) * rD.x, t2y = (bmaxy - o.y) * rD.y, t2z = (bmaxz - o.z) * rD.z;
		const float tmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
		const float tmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
		return (tmax >= tmin && tmin < maxLen && tmax > 0.0f) ? tmin : kFltMax;
	}

	__device__ __forceinline__ bool FiniteF(float f) { return (__float_as_uint(f) & 0x7F800000u) != 0x7F800000u; }

	// The warp loop.  Source: bool Load(uint32_t index, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) (false = nothing to
	// trace at this index).  Sink: void Retire(bool finished, uint32_t index, const Hit&, bool anyHit, V3 o, V3 d) called by ALL lanes each iteration.
	// Lane state is one word: kLaneIdle, an inner node index, or kLeafBit | triangle slot (the next triangle to test).
#define SPT_PRAGMA_(x) _Pragma(#x)
#define SPT_UNROLL(n) SPT_PRAGMA_(unroll n)
#ifndef SPT_EXACT_UNROLL
#define SPT_EXACT_UNROLL 4         // the repetition loops of the warp loops are unrolled (exact kernel on C3: 90.0 -> 88.1 ms, profiles/r03s_exact_unroll_c3.txt)
#endif
#ifndef SPT_VOTE_INNER_BIAS
#define SPT_VOTE_INNER_BIAS 1      // inner step when nInner * bias >= nLeaf
#endif
	constexpr uint32_t kLaneIdle = 0xFFFFFFFFu;

#if defined(SPT_TRACE_STATS)
	// tuning aid (tools/trace_variants.py "stats" variant): lane-state sums over every vote of every warp
	// 0 votes, 1 idle lanes, 2 leaf lanes, 3 inner lanes, 4 inner reps, 5 lanes active in inner reps, 6 leaf reps, 7 lanes active in leaf reps, 8 refills, 9 lanes refilled
	__device__ unsigned long long g_traceStats[16];
#define SPT_STAT(i, v) do { if (lane == 0) st_[i] += (v); } while (0)
#else
#define SPT_STAT(i, v) do { } while (0)
#endif

	template<class Source, class Sink>
	__device__ __forceinline__ void TraceWarpLoop(const BvhView& bvh, uint32_t n, uint32_t* __restrict__ counter, uint32_t* stackMem, Source& src, Sink& sink)
	{
		// 32-bit shared-window address of this lane's stack column: [entry][thread], 4-byte entries
		const uint32_t sAddr = (uint32_t)__cvta_generic_to_shared(stackMem) + threadIdx.x * 4u;
		uint32_t ovf[kStackDepth - kSmemStack];
		const uint32_t lane = threadIdx.x & 31;
		V3 o = v3(0.0f), d = v3(0.0f), rD = v3(0.0f);
		float maxLen = 0.0f; uint32_t ignore = kNoHit, index = 0;
		bool safe = false, anyHit = false;
		Hit hit; hit.t = 0.0f; hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
		uint32_t cur = kLaneIdle;
		int sp = 0;
		bool exhausted = false;
#if defined(SPT_TRACE_STATS)
		unsigned long long st_[10] = {};
#endif

		for (;;)
		{
			// ---- refill idle lanes -------------------------------------------------------------------------
			const uint32_t idleMask = __ballot_sync(0xffffffffu, cur == kLaneIdle);
			if (idleMask)
			{
				if (!exhausted && (__popc(idleMask) >= (int)kFetchMinIdle))
				{
					const uint32_t want = (uint32_t)__popc(idleMask);
					uint32_t base = 0;
					if (lane == 0) base = atomicAdd(counter, want);
					base = __shfl_sync(0xffffffffu, base, 0);
					if (base + want >= n) exhausted = true;
					SPT_STAT(8, 1); SPT_STAT(9, want);
					if (cur == kLaneIdle)
					{
						const uint32_t i = base + (uint32_t)__popc(idleMask & ((1u << lane) - 1u));
						if (i < n && src.Load(i, o, d, ignore, maxLen, anyHit))
						{
							index = i;
							rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);          // Ray::SetDirection (Bounds.h:44-48)
							safe = FiniteF(rD.x) && FiniteF(rD.y) && FiniteF(rD.z) && FiniteF(o.x) && FiniteF(o.y) && FiniteF(o.z);
							hit.t = u2f(0x7F800000u); hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
							sp = 0;
							cur = bvh.rootRef;
						}
					}
					continue;
				}
				if (idleMask == 0xffffffffu)                  // queue exhausted and every lane retired
				{
#if defined(SPT_TRACE_STATS)
					if (lane == 0) for (int k = 0; k < 10; k++) atomicAdd(&g_traceStats[k], st_[k]);
#endif
					break;
				}
			}
			// ---- vote: the step most busy lanes need --------------------------------------------------------
			const bool isLeaf = (cur & kLeafBit) && cur != kLaneIdle;
			const uint32_t leafMask = __ballot_sync(0xffffffffu, isLeaf);
			const int nLeaf = __popc(leafMask), nInner = 32 - __popc(idleMask) - nLeaf;
			// The chosen step is repeated a few times per vote (lanes that left the mode sit out): the vote, the refill check
			// and the retire cost ~45 warp-wide instructions, a step ~80.
			bool finished = false;
			SPT_STAT(0, 1); SPT_STAT(1, __popc(idleMask)); SPT_STAT(2, nLeaf); SPT_STAT(3, nInner);
			if (nInner * SPT_VOTE_INNER_BIAS >= nLeaf * SPT_VOTE_LEAF_BIAS)
			{
SPT_UNROLL(SPT_EXACT_UNROLL)
				for (int rep = 0; rep < SPT_INNER_REPS; rep++)
				{
#if defined(SPT_TRACE_STATS)
					{ const uint32_t am = __ballot_sync(0xffffffffu, !(cur & kLeafBit)); SPT_STAT(4, 1); SPT_STAT(5, __popc(am)); }
#endif
					if (!(cur & kLeafBit))
					{
						const TNode* nd = bvh.nodes + cur;
						const V4 q0 = ld4(&nd->q0), q1 = ld4(&nd->q1), q2 = ld4(&nd->q2);
						const V4 q3 = ld4(reinterpret_cast<const V4*>(&nd->left));
						uint32_t c1 = f2u(q3.x), c2 = f2u(q3.y);
						float d1, d2;
						if (safe)
						{
							d1 = SlabTestFast(o, rD, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, maxLen);
							d2 = SlabTestFast(o, rD, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, maxLen);
						}
						else
						{
							d1 = SlabTest(o, rD, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, maxLen);
							d2 = SlabTest(o, rD, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, maxLen);
						}
						if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const uint32_t tc = c1; c1 = c2; c2 = tc; }   // BVH.cpp:163-167
						if (d1 == kFltMax)                                          // both children missed: pop
						{
							if (sp == 0) { finished = true; cur = kLaneIdle; }
							else
							{
								sp--;
								if (sp < kSmemStack) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(sAddr + (uint32_t)sp * (kTraceBlock * 4u)) : "memory");
								else cur = ovf[sp - kSmemStack];
							}
						}
						else
						{
							cur = c1;
							if (d2 != kFltMax)
							{
								if (sp < kSmemStack) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sAddr + (uint32_t)sp * (kTraceBlock * 4u)), "r"(c2) : "memory");
								else ovf[sp - kSmemStack] = c2;
								sp++;
							}
						}
					}
#if defined(SPT_DYN_REPS)
					if (__popc(__ballot_sync(0xffffffffu, !(cur & kLeafBit))) * 2 < nInner) break;   // most lanes left the mode: vote again
#endif
				}
			}
			else
			{
SPT_UNROLL(SPT_EXACT_UNROLL)
				for (int rep = 0; rep < SPT_LEAF_REPS; rep++)
				{
#if defined(SPT_TRACE_STATS)
					{ const uint32_t am = __ballot_sync(0xffffffffu, (cur & kLeafBit) && cur != kLaneIdle); SPT_STAT(6, 1); SPT_STAT(7, __popc(am)); }
#endif
					if ((cur & kLeafBit) && cur != kLaneIdle)
					{
						const TTri* T = bvh.tris + (cur & ~kLeafBit);
						const V4 a = ld4(&T->a), b = ld4(&T->b), c = ld4(&T->c);
						const uint32_t triId = f2u(c.y);
						bool pop = f2u(c.w) != 0u;                                  // last triangle of its leaf
						if (ignore != triId)                                       // BVH.cpp:136-139
						{
							float t, u, v;
							if (TriTest(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), maxLen, t, u, v))
							{
								hit.t = t; hit.u = u; hit.v = v; hit.tri = triId;
								maxLen = std_min(maxLen, t);
								if (anyHit) { sp = 0; pop = true; }                 // hit-or-miss query: done
							}
						}
						if (!pop) cur++;
						else if (sp == 0) { finished = true; cur = kLaneIdle; }
						else
						{
							sp--;
							if (sp < kSmemStack) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cur) : "r"(sAddr + (uint32_t)sp * (kTraceBlock * 4u)) : "memory");
							else cur = ovf[sp - kSmemStack];
						}
					}
#if defined(SPT_DYN_REPS)
					if (__popc(__ballot_sync(0xffffffffu, (cur & kLeafBit) && cur != kLaneIdle)) * 2 < nLeaf) break;
#endif
				}
			}
			sink.Retire(finished, index, hit, anyHit, o, d);
		}
	}

	// ---- sources and sinks ---------------------------------------------------------------------------------------
	struct QueueSource
	{
		const RayRec* rays;
		__device__ __forceinline__ bool Load(uint32_t i, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) const
		{
			const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays + i));
			const float4 r1 = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
			if (r1.w == -1.0f) return false;   // inactive queue entry
			o = v3(r0.x, r0.y, r0.z); ignore = __float_as_uint(r0.w); d = v3(r1.x, r1.y, r1.z);
			anyHit = r1.w < 0.0f; maxLen = fabsf(r1.w);
			return true;
		}
	};
	struct QueueSink
	{
		Hit* hits;
		__device__ __forceinline__ void Retire(bool finished, uint32_t i, const Hit& h, bool, V3, V3) const
		{
			if (finished) *reinterpret_cast<float4*>(hits + i) = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
		}
	};
	// Wavefront levels: every ray leaves ONE status byte (hit or miss).  Only closest-hit queries that hit something (a child
	// activation or a TraceSky walk may start there) also leave a SlowRec (hit + ray + index) for ClassifyKernel.
	struct WavefrontOut { uint8_t* status; const uint32_t* auxBase; SlowRec* slow; uint32_t* slowCount; };
	struct WavefrontSink
	{
		uint8_t* status; SlowRec* slowRecs; uint32_t* slowCount;     // status already offset by the level's auxBase
		__device__ __forceinline__ void Retire(bool finished, uint32_t i, const Hit& h, bool anyHit, V3 o, V3 d) const
		{
			const bool hitSomething = h.tri != kNoHit;
			if (finished) status[i] = hitSomething ? 1 : 0;
			const bool slow = finished && hitSomething && !anyHit;
			const uint32_t m = __ballot_sync(0xffffffffu, slow);
			if (m)
			{
				const uint32_t lane = threadIdx.x & 31;
				const int leader = __ffs(m) - 1;
				uint32_t base = 0;
				if ((int)lane == leader) base = atomicAdd(slowCount, (uint32_t)__popc(m));
				base = __shfl_sync(0xffffffffu, base, leader);
				if (slow)
				{
					float4* r = reinterpret_cast<float4*>(slowRecs + (base + (uint32_t)__popc(m & ((1u << lane) - 1u))));
					r[0] = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
					r[1] = make_float4(o.x, o.y, o.z, __uint_as_float(i));
					r[2] = make_float4(d.x, d.y, d.z, 0.0f);
				}
			}
		}
	};

	// nPtr (optional): the queue length lives in device memory (wavefront levels); n is then its capacity
	__global__ void __launch_bounds__(kTraceBlock) k_trace_rays(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; QueueSink sink{ hits };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
	}

	__global__ void __launch_bounds__(kTraceBlock) k_trace_level(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter, WavefrontOut out)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		{ const uint32_t m = *nPtr; if (m < n) n = m; }
		QueueSource src{ rays }; WavefrontSink sink{ out.status + *out.auxBase, out.slow, out.slowCount };
		TraceWarpLoop(bvh, n, counter, stackMem, src, sink);
	}

	// ---- small scenes: the whole traversal layout lives in shared memory --------------------------------------------
	// A scene of a few hundred triangles (BASELINE config C1/C2 is a 12-triangle cube) has no memory problem to solve: the
	// node and triangle records are copied into shared memory once per CTA and every thread walks its own ray with the
	// reference's loop (TraceClosest's order), grid-stride over the queue.  Nothing is voted or refilled because a ray
	// lives for a handful of steps.
	constexpr uint32_t kSmallSceneBytes = 40u * 1024u;
	constexpr int kSmallBlock = 256;

	template<bool kWavefront>
	__global__ void __launch_bounds__(kSmallBlock) k_trace_rays_small(const TNode* __restrict__ gNodes, uint32_t numNodes, const TTri* __restrict__ gTris, uint32_t numTris,
		uint32_t rootRef, const RayRec* __restrict__ rays, Hit* __restrict__ hits, uint32_t n, const uint32_t* __restrict__ nPtr, WavefrontOut out)
	{
		extern __shared__ __align__(16) unsigned char smallMem[];
		V4* sNodes = reinterpret_cast<V4*>(smallMem);                         // 4 x V4 per node
		V4* sTris = sNodes + (size_t)numNodes * 4;                            // 3 x V4 per triangle
		{
			const V4* gn = reinterpret_cast<const V4*>(gNodes); const V4* gt = reinterpret_cast<const V4*>(gTris);
			for (uint32_t i = threadIdx.x; i < numNodes * 4u; i += kSmallBlock) sNodes[i] = ld4(gn + i);
			for (uint32_t i = threadIdx.x; i < numTris * 3u; i += kSmallBlock) sTris[i] = ld4(gt + i);
		}
		__syncthreads();
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		uint32_t stack[kStackDepth];
		for (uint32_t i = blockIdx.x * kSmallBlock + threadIdx.x; i < n; i += gridDim.x * kSmallBlock)
		{
			const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays + i));
			const float4 r1 = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
			if (r1.w == -1.0f) continue;                                       // inactive queue entry
			const V3 o = v3(r0.x, r0.y, r0.z), d = v3(r1.x, r1.y, r1.z);
			const uint32_t ignore = __float_as_uint(r0.w);
			const bool anyHit = r1.w < 0.0f;
			float maxLen = fabsf(r1.w);
			const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);            // Ray::SetDirection (Bounds.h:44-48)
			const bool safe = FiniteF(rD.x) && FiniteF(rD.y) && FiniteF(rD.z) && FiniteF(o.x) && FiniteF(o.y) && FiniteF(o.z);
			Hit hit; hit.t = u2f(0x7F800000u); hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
			uint32_t cur = rootRef; int sp = 0;
			for (;;)                                                           // BVH.cpp:128-189
			{
				if (cur & kLeafBit)
				{
					const V4* T = sTris + (size_t)(cur & ~kLeafBit) * 3;
					bool done = false;
					for (;;)
					{
						const V4 a = T[0], b = T[1], c = T[2];
						const uint32_t triId = f2u(c.y);
						if (ignore != triId)
						{
							float t, u, v;
							if (TriTest(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), maxLen, t, u, v))
							{
								hit.t = t; hit.u = u; hit.v = v; hit.tri = triId;
								maxLen = std_min(maxLen, t);
								if (anyHit) { done = true; break; }
							}
						}
						if (f2u(c.w)) break;
						T += 3;
					}
					if (done || sp == 0) break;
					cur = stack[--sp];
					continue;
				}
				const V4* nd = sNodes + (size_t)cur * 4;
				const V4 q0 = nd[0], q1 = nd[1], q2 = nd[2], q3 = nd[3];
				uint32_t c1 = f2u(q3.x), c2 = f2u(q3.y);
				float d1, d2;
				if (safe)
				{
					d1 = SlabTestFast(o, rD, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, maxLen);
					d2 = SlabTestFast(o, rD, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, maxLen);
				}
				else
				{
					d1 = SlabTest(o, rD, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, maxLen);
					d2 = SlabTest(o, rD, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, maxLen);
				}
				if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const uint32_t tc = c1; c1 = c2; c2 = tc; }
				if (d1 == kFltMax)
				{
					if (sp == 0) break;
					cur = stack[--sp];
				}
				else
				{
					cur = c1;
					if (d2 != kFltMax) stack[sp++] = c2;
				}
			}
			if (!kWavefront) *reinterpret_cast<float4*>(hits + i) = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.tri));
			else
			{
				const bool hitSomething = hit.tri != kNoHit;
				out.status[*out.auxBase + i] = hitSomething ? 1 : 0;
				if (hitSomething && !anyHit)
				{
					float4* r = reinterpret_cast<float4*>(out.slow + atomicAdd(out.slowCount, 1u));
					r[0] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.tri));
					r[1] = make_float4(o.x, o.y, o.z, __uint_as_float(i));
					r[2] = make_float4(d.x, d.y, d.z, 0.0f);
				}
			}
		}
	}

	// Two other forms of the small-scene kernel were measured on C2 and rejected (profiles/r01g_SUMMARY.md): the persistent
	// vote / refill warp loop reading shared memory (3.54 ms against 2.52 ms: with walks of a few steps the bookkeeping costs
	// more than the idle lanes it removes) and a two-phase form that retires root-miss rays first and compacts the survivors
	// (2.67 ms: 70 % of a first-hit level's rays point INTO a convex object and cross it, so there is little to compact).

	// primary rays of sample 0 generated in-kernel (no ray queue traffic): work index -> 8x4 pixel tile + lane
	struct PrimarySource
	{
		CameraGpu cam; uint32_t tilesX;
		__device__ __forceinline__ void Pixel(uint32_t i, uint32_t& x, uint32_t& y) const
		{
			const uint32_t tile = i >> 5, l = i & 31u;
			x = (tile % tilesX) * 8u + (l & 7u); y = (tile / tilesX) * 4u + (l >> 3);
		}
		__device__ __forceinline__ bool Load(uint32_t i, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) const
		{
			uint32_t x, y; Pixel(i, x, y);
			if (x >= cam.width || y >= cam.height) return false;
			o = cam.pos; d = PrimaryDir(cam, x, y, 0.5f, 0.5f); ignore = kNoHit; maxLen = kFltMax; anyHit = false;
			return true;
		}
	};
	struct PrimarySink
	{
		Hit* hits; PrimarySource src;
		__device__ __forceinline__ void Retire(bool finished, uint32_t i, const Hit& h, bool, V3, V3) const
		{
			if (!finished) return;
			uint32_t x, y; src.Pixel(i, x, y);
			*reinterpret_cast<float4*>(hits + (size_t)y * src.cam.width + x) = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
		}
	};

	__global__ void __launch_bounds__(kTraceBlock) k_trace_primary(BvhView bvh, CameraGpu cam, Hit* __restrict__ hits, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		// 8x4 pixel tiles per warp keep the 32 rays of a fetch spatially coherent
		const uint32_t tilesX = (cam.width + 7) / 8, tilesY = (cam.height + 3) / 4;
		PrimarySource src{ cam, tilesX }; PrimarySink sink{ hits, src };
		TraceWarpLoop(bvh, tilesX * tilesY * 32u, counter, stackMem, src, sink);
	}

	inline int TraceGridSize()
	{
		static int grid = 0;
		if (!grid)
		{
			int dev = 0, sms = 148, perSm = 1;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace_rays, kTraceBlock, 0);
			grid = sms * (perSm > 0 ? perSm : 1);
		}
		return grid;
	}

	inline bool SmallScene(const BvhView& bvh, size_t& bytes)
	{
		bytes = (size_t)bvh.numNodes * sizeof(TNode) + (size_t)bvh.numTris * sizeof(TTri);
		if (bytes > kSmallSceneBytes) return false;
		static bool attr = false;
		if (!attr)
		{
			cudaFuncSetAttribute(k_trace_rays_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSceneBytes);
			cudaFuncSetAttribute(k_trace_rays_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSceneBytes);
			attr = true;
		}
		return true;
	}
	inline uint32_t SmallGrid(uint32_t n)
	{
		uint32_t blocks = (n + kSmallBlock - 1) / kSmallBlock;
		const uint32_t lim = (uint32_t)RangeGridBlocks();
		return blocks > lim ? lim : blocks;
	}

	inline void LaunchTraceRays(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t n, uint32_t* counter, const uint32_t* nPtr = nullptr)
	{
		if (!n || !ctx.ok) return;
		size_t smallBytes;
		if (SmallScene(bvh, smallBytes))
		{
			k_trace_rays_small<false><<<SmallGrid(n), kSmallBlock, smallBytes, ctx.stream>>>(bvh.nodes, bvh.numNodes, bvh.tris, bvh.numTris, bvh.rootRef, rays, hits, n, nPtr, WavefrontOut{});
			ctx.kernelLaunches++;
			SPT_CUDA_CHECK(ctx, cudaGetLastError());
			return;
		}
		DevMemset(ctx, counter, 0, sizeof(uint32_t));
		k_trace_rays<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, n, nPtr, counter);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}

	// one level of the wavefront: queue length in device memory, status bytes + slow list out
	inline void LaunchTraceLevel(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t cap, uint32_t* counter, const uint32_t* nPtr, const WavefrontOut& out)
	{
		if (!cap || !ctx.ok) return;
		size_t smallBytes;
		if (SmallScene(bvh, smallBytes))
		{
			k_trace_rays_small<true><<<SmallGrid(cap), kSmallBlock, smallBytes, ctx.stream>>>(bvh.nodes, bvh.numNodes, bvh.tris, bvh.numTris, bvh.rootRef, rays, hits, cap, nPtr, out);
		}
		else
		{
			DevMemset(ctx, counter, 0, sizeof(uint32_t));
			k_trace_level<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, cap, nPtr, counter, out);
		}
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}

	inline void LaunchTracePrimary(Ctx& ctx, const BvhView& bvh, const CameraGpu& cam, Hit* hits, uint32_t* counter)
	{
		if (!ctx.ok) return;
		DevMemset(ctx, counter, 0, sizeof(uint32_t));
		k_trace_primary<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, cam, hits, counter);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	inline bool SmallScene(const BvhView& bvh, size_t& bytes)
	{
		bytes = (size_t)bvh.numNodes * sizeof(TNode) + (size_t)bvh.numTris * sizeof(TTri);
		return bytes <= 40u * 1024u;
	}
	inline void LaunchTraceRays(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t n, uint32_t*, const uint32_t* nPtr = nullptr)
	{
		LocalStack st;
		if (nPtr && *nPtr < n) n = *nPtr;
		for (uint32_t i = 0; i < n; i++)
			if (rays[i].tmax != -1.0f) TraceClosest(bvh, v3(rays[i].ox, rays[i].oy, rays[i].oz), v3(rays[i].dx, rays[i].dy, rays[i].dz), rays[i].ignoreTri, fabsf(rays[i].tmax), st, hits[i]);
		ctx.kernelLaunches++;
	}
	struct WavefrontOut { uint8_t* status; const uint32_t* auxBase; SlowRec* slow; uint32_t* slowCount; };
	inline void PushSlow(const WavefrontOut& out, const Hit& h, const RayRec& r, uint32_t i)
	{
		SlowRec s; s.t = h.t; s.u = h.u; s.v = h.v; s.tri = h.tri; s.ox = r.ox; s.oy = r.oy; s.oz = r.oz; s.index = i; s.dx = r.dx; s.dy = r.dy; s.dz = r.dz; s.pad = 0;
		out.slow[(*out.slowCount)++] = s;
	}
	inline void LaunchTraceLevel(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t cap, uint32_t*, const uint32_t* nPtr, const WavefrontOut& out)
	{
		LocalStack st;
		const uint32_t n = *nPtr < cap ? *nPtr : cap;
		for (uint32_t i = 0; i < n; i++)
		{
			if (rays[i].tmax == -1.0f) continue;
			Hit h;
			TraceClosest(bvh, v3(rays[i].ox, rays[i].oy, rays[i].oz), v3(rays[i].dx, rays[i].dy, rays[i].dz), rays[i].ignoreTri, fabsf(rays[i].tmax), st, h);
			out.status[*out.auxBase + i] = h.tri != kNoHit ? 1 : 0;
			if (h.tri != kNoHit && !(rays[i].tmax < 0.0f)) PushSlow(out, h, rays[i], i);
		}
		ctx.kernelLaunches++;
	}
	inline void LaunchTracePrimary(Ctx& ctx, const BvhView& bvh, const CameraGpu& cam, Hit* hits, uint32_t*)
	{
		LocalStack st;
		for (uint32_t y = 0; y < cam.height; y++)
			for (uint32_t x = 0; x < cam.width; x++)
				TraceClosest(bvh, cam.pos, PrimaryDir(cam, x, y, 0.5f, 0.5f), kNoHit, kFltMax, st, hits[(size_t)y * cam.width + x]);
		ctx.kernelLaunches++;
	}
#endif
}
