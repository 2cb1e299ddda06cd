// render.cuh — host driver of one frame (the tile loop of PathTracer.cpp:418-487 becomes kernel launches):
//
//   primary pass   one fused kernel: camera ray generation + closest hit for EVERY (pixel, primary sample) of the
//                  shard.  Misses are final (Raytrace returns the ambient colour, :873-876) and are written straight
//                  to the per-sample buffer; hits are appended to a compact first-hit queue (warp-aggregated atomics).
//   wavefront      the first hits are processed in batches; a batch evaluates the reference's recursion tree level by
//                  level (integrator.cuh): [expand (level 0: from the first-hit queue), fan-out, trace, classify, (sky), next-level] per level, then gather bottom-up.
//                  Level sizes live in device memory, so a whole batch is queued without host synchronisation; the
//                  host reads one small word (overflow flag + ray count) per batch.
//   resolve        accumulator = sum over the pixel's primary samples in index order / msaa, row flip (:449,468-469).
//
// All working buffers live with the scene and are reused across frames (no allocation in the steady state).
#pragma once
#include "integrator.cuh"
#include <mutex>
#include <unordered_map>

#ifndef SPT_EXPAND_MIN_BLOCKS
#define SPT_EXPAND_MIN_BLOCKS 2
#endif
#ifndef SPT_FAN_MIN_BLOCKS
#define SPT_FAN_MIN_BLOCKS 4
#endif
#ifndef SPT_CLASSIFY_MIN_BLOCKS
#define SPT_CLASSIFY_MIN_BLOCKS 1
#endif
#ifndef SPT_GATHER_MIN_BLOCKS
#define SPT_GATHER_MIN_BLOCKS 1
#endif

namespace spt
{
	struct PrimaryArgs
	{
		BvhView bvh; CameraGpu cam;
		uint32_t rowBegin, rowEnd, msBegin, msEnd, msaa; uint64_t seed; V3 ambient;
		float* sampleBuf; PrimaryHitRec* queue; uint32_t* hitCount; uint32_t total;
	};

	// decode work index g -> (x, y, sample); 8x4 pixel tiles, the 32 lanes of a warp share the sample index
	SPT_HD bool DecodePrimary(const PrimaryArgs& a, uint32_t g, uint32_t& x, uint32_t& y, uint32_t& sample)
	{
		const uint32_t rows = a.rowEnd - a.rowBegin, ns = a.msEnd - a.msBegin;
		const uint32_t tilesX = (a.cam.width + 7u) / 8u;
		const uint32_t lane = g & 31u, rest = g >> 5;
		const uint32_t s = rest % ns, tile = rest / ns;
		x = (tile % tilesX) * 8u + (lane & 7u);
		const uint32_t yb = (tile / tilesX) * 4u + (lane >> 3);
		y = a.rowBegin + yb; sample = a.msBegin + s;
		return x < a.cam.width && yb < rows;
	}

	template<class Stack>
	SPT_KERNEL_BODY bool PrimarySample(const PrimaryArgs& a, uint32_t x, uint32_t y, uint32_t sample, Stack& stack, PrimaryHitRec& rec)
	{
		const uint32_t pixel = y * a.cam.width + x;
		Rng rng; rng.key = PrimaryRngKey(a.seed, pixel, a.msaa, sample); rng.counter = 0;
		float ox = 0.5f, oy = 0.5f;                                               // PathTracer.cpp:460
		if (sample != 0) { ox = rng.Float01(); oy = rng.Float01(); }
		Hit h;
		const bool hit = TraceClosest(a.bvh, a.cam.pos, PrimaryDir(a.cam, x, y, ox, oy), kNoHit, kFltMax, stack, h);
		if (!hit)
		{
			const size_t idx = ((size_t)(y - a.rowBegin) * a.cam.width + x) * (a.msEnd - a.msBegin) + (sample - a.msBegin);
			a.sampleBuf[idx * 3] = a.ambient.x; a.sampleBuf[idx * 3 + 1] = a.ambient.y; a.sampleBuf[idx * 3 + 2] = a.ambient.z;   // :873-876
			return false;
		}
		rec.pixel = pixel; rec.sample = sample; rec.pad0 = rec.pad1 = 0; rec.t = h.t; rec.u = h.u; rec.v = h.v; rec.tri = h.tri;
		return true;
	}

#if !defined(SPT_EMU)
	struct PrimaryPassSource
	{
		PrimaryArgs a;
		__device__ __forceinline__ bool Load(uint32_t g, V3& o, V3& d, uint32_t& ignore, float& maxLen, bool& anyHit) const
		{
			uint32_t x, y, sample;
			if (!DecodePrimary(a, g, x, y, sample)) return false;
			const uint32_t pixel = y * a.cam.width + x;
			Rng rng; rng.key = PrimaryRngKey(a.seed, pixel, a.msaa, sample); rng.counter = 0;
			float ox = 0.5f, oy = 0.5f;                                               // PathTracer.cpp:460
			if (sample != 0) { ox = rng.Float01(); oy = rng.Float01(); }
			o = a.cam.pos; d = PrimaryDir(a.cam, x, y, ox, oy); ignore = kNoHit; maxLen = kFltMax; anyHit = false;
			return true;
		}
	};
	struct PrimaryPassSink
	{
		PrimaryArgs a;
		__device__ __forceinline__ void Retire(bool finished, uint32_t g, const Hit& h, bool, V3, V3) const
		{
			const bool hit = finished && h.tri != kNoHit;
			uint32_t x = 0, y = 0, sample = 0;
			if (finished) DecodePrimary(a, g, x, y, sample);
			if (finished && !hit)
			{
				const size_t idx = ((size_t)(y - a.rowBegin) * a.cam.width + x) * (a.msEnd - a.msBegin) + (sample - a.msBegin);
				a.sampleBuf[idx * 3] = a.ambient.x; a.sampleBuf[idx * 3 + 1] = a.ambient.y; a.sampleBuf[idx * 3 + 2] = a.ambient.z;   // :873-876
			}
			// warp-aggregated append: one atomic per warp and iteration, the warp's hits stay contiguous in the queue
			const uint32_t m = __ballot_sync(0xffffffffu, hit);
			if (m)
			{
				const uint32_t lane = threadIdx.x & 31;
				const int leader = __ffs(m) - 1;
				uint32_t qbase = 0;
				if ((int)lane == leader) qbase = atomicAdd(a.hitCount, (uint32_t)__popc(m));
				qbase = __shfl_sync(0xffffffffu, qbase, leader);
				if (hit)
				{
					const uint32_t slot = qbase + (uint32_t)__popc(m & ((1u << lane) - 1u));
					float4* q = reinterpret_cast<float4*>(a.queue + slot);
					q[0] = make_float4(__uint_as_float(y * a.cam.width + x), __uint_as_float(sample), 0.0f, 0.0f);
					q[1] = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
				}
			}
		}
	};

	__global__ void __launch_bounds__(kTraceBlock) k_primary_pass(PrimaryArgs a, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kSmemStack * kTraceBlock];
		PrimaryPassSource src{ a }; PrimaryPassSink sink{ a };
		TraceWarpLoop(a.bvh, a.total, counter, stackMem, src, sink);
	}

	inline void LaunchPrimaryPass(Ctx& ctx, const PrimaryArgs& a, uint32_t* counter)
	{
		if (!ctx.ok || !a.total) return;
		DevMemset(ctx, counter, 0, sizeof(uint32_t));
		k_primary_pass<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(a, counter);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	inline void LaunchPrimaryPass(Ctx& ctx, const PrimaryArgs& a, uint32_t*)
	{
		LocalStack st;
		for (uint32_t g = 0; g < a.total; g++)
		{
			uint32_t x, y, sample; PrimaryHitRec rec;
			if (DecodePrimary(a, g, x, y, sample) && PrimarySample(a, x, y, sample, st, rec)) a.queue[atomic_add_u32(a.hitCount, 1u)] = rec;
		}
		ctx.kernelLaunches++;
	}
#endif

	template<class T> inline T* EnsureBytes(Ctx& ctx, PlainBuf& b, size_t count)
	{
		b.Ensure(ctx, count * sizeof(T) + 16);
		return reinterpret_cast<T*>(b.p);
	}

	// Arena sizes of one batch of first hits.  Per first hit: level 0 needs D+A+S+1 rays; every importance sample can spawn
	// one child activation with D+2 (+1 alpha) rays; deeper levels shrink (misses, the 0.01 throughput cut), which is
	// budgeted as `kDepthFactor` levels' worth of the level-1 worst case.  Running out is detected on the device
	// (BatchCounters::overflow) and the batch is redone at half the size, so the factors only affect speed.
	struct BatchPlan { uint32_t firstHits, rayCap, auxCap, recCap, skyCap; };

	// Working-set budget of one batch of first hits: 40 GiB of the B200's 180 GB (a whole C2 frame is then ONE batch: every batch
	// boundary costs a host round trip, and a descheduled host thread shows up as an idle GPU), never more than a third of what the device
	// can still give (`held` = bytes the scene's arenas already own, which the batch reuses).  Decided once per frame.
	inline std::mutex& BudgetCacheMutex() { static std::mutex m; return m; }
	inline double* BudgetCacheTimes() { static double t[64] = {}; return t; }          // 0 = ask the driver again
	inline uint64_t BatchBudget(uint64_t held)
	{
		uint64_t budget = 40960ull << 20;
		// cudaMemGetInfo goes through the driver's resource-manager lock: tens of milliseconds on a busy box (profiles/r01g_SUMMARY.md), and
		// 0.2-0.4 s when two host threads of a multi-device frame ask while both devices are busy (profiles/r03_SUMMARY.md: every frame that
		// fell on the former 5-second refresh).  Free + held memory of a device only changes when somebody else allocates on it, so it is asked
		// once, again after five minutes, and after SailorPt_TrimMemory (one cache entry per device).
		const int dev = DevCurrent() & 63;
		const double now = HostNow();
		uint64_t cachedFree;
		{
			std::lock_guard<std::mutex> lock(BudgetCacheMutex());
			double* cachedAtOf = BudgetCacheTimes(); static uint64_t cachedFreeOf[64] = {};
			if (cachedAtOf[dev] == 0.0 || now - cachedAtOf[dev] > 300.0) { cachedFreeOf[dev] = (uint64_t)DevMemAvailable() + held; cachedAtOf[dev] = now; }
			cachedFree = cachedFreeOf[dev];
		}
		const uint64_t avail = cachedFree / 3u;
		if (avail && budget > avail) budget = avail;
		if (budget < (256ull << 20)) budget = 256ull << 20;
		if (const char* e = getenv("SAILOR_PT_BATCH_MB")) { const long v = atol(e); if (v > 0) budget = (uint64_t)v << 20; }
		return budget;
	}

	inline BatchPlan PlanBatch(uint64_t budget, uint64_t hitCount, uint32_t D, uint32_t A, uint32_t S, uint32_t maxBounces, bool ambientOn, uint32_t shrink)
	{
		const uint64_t kDepthFactor = maxBounces < 3u ? maxBounces : 3u;
		if (!ambientOn) { A = 0; S = 0; }
		const uint64_t lvl0 = (uint64_t)D + A + S + 1u, lvl1 = (uint64_t)S * (D + 3u) + 1u;
		const uint64_t raysPer = lvl0 > lvl1 ? lvl0 : lvl1;
		const uint64_t auxPer = lvl0 + kDepthFactor * lvl1;
		const uint64_t recPer = 2u + kDepthFactor * S;
		const uint64_t bytesPer = raysPer * (sizeof(RayRec) + sizeof(SlowRec)) + auxPer * (sizeof(RayAux) + 1u) + recPer * sizeof(NodeRec);
		uint64_t B = budget / bytesPer;
		B >>= shrink;
		if (B < 1024u) B = 1024u;
		if (B > hitCount) B = hitCount;
		const uint64_t cap32 = 0xFFFF0000ull;
		while (B > 1u && (B * auxPer >= cap32 || B * recPer >= cap32)) B >>= 1;
		BatchPlan p;
		p.firstHits = (uint32_t)B;
		p.rayCap = (uint32_t)(B * raysPer + 1024u);
		p.auxCap = (uint32_t)(B * auxPer + 1024u);
		p.recCap = (uint32_t)(B * recPer + 1024u);
		p.skyCap = (uint32_t)((B * raysPer) / 4u + 65536u);
		return p;
	}

	// Wavefront working set, shared by every scene object of the process on one device and kept for the life of the process (it
	// only ever grows): 0 activation records, 1 RayAux arena, 2 rays, 3 (unused), 4 per-sample results, 5 primary-hit queue, 6 batch
	// counters, 7 blue-noise table, 8 fan-out contexts, 9 sky states, 10 sky rays, 11 sky hits, 12 ray status bytes, 13 SlowRec stream (closest hits for classify),
	// 14 fan-out slot tables, 15 replay list of the wide traversal.  A host that creates one scene object per frame (the reference's PathTracer object per Run) would
	// otherwise allocate and free tens of GiB of arenas every frame (~1.5 ms per C2 frame even from the stream-ordered pool); plain
	// cudaMalloc memory, because a stream-ordered allocation must be freed on a stream that may no longer exist.  One
	// frame renders at a time per device (`frame` mutex); a frame saturates the GPU anyway.
	struct SharedArenas { PlainBuf mem[16]; std::mutex frame; };
	inline SharedArenas& ArenasOfCurrentDevice()
	{
		static std::mutex m;
		static std::unordered_map<int, SharedArenas*>* all = new std::unordered_map<int, SharedArenas*>();   // leaked on purpose: never freed after CUDA teardown
		std::lock_guard<std::mutex> lock(m);
		SharedArenas*& p = (*all)[DevCurrent()];
		if (!p) p = new SharedArenas();
		return *p;
	}

	// Give the shared working set of the current device back (SailorPt_TrimMemory): the next frame allocates it again.
	inline void ReleaseSharedArenas(Ctx& ctx)
	{
		SharedArenas& arenas = ArenasOfCurrentDevice();
		std::lock_guard<std::mutex> frameLock(arenas.frame);
		ctx.Sync();
		for (PlainBuf& b : arenas.mem) { if (b.p) DevFreePlain(b.p); b.p = nullptr; b.n = 0; }
		{ std::lock_guard<std::mutex> lock(BudgetCacheMutex()); BudgetCacheTimes()[DevCurrent() & 63] = 0.0; }
	}

	inline bool SceneHasThickTransmission(const HostScene& h)
	{
		for (const auto& m : h.materials) if (m.transmission > 0.0f && m.thickness > 0.0f) return true;
		return false;
	}

	// Progressive pass (optional): `running` carries the un-normalised sum of the primary samples rendered so far, `runningValid`
	// says whether to continue from it, and the image is normalised by `norm` (samples accumulated so far) instead of p.msaa.
	struct ProgressiveArgs { float* running = nullptr; bool runningValid = false; uint32_t norm = 0; };

	inline int RenderFrame(SceneDevice& D, const CameraGpu& cam, const SailorPtParams& p, float* dImage, RenderStats& rs, const ProgressiveArgs& prog = ProgressiveArgs())
	{
		Ctx& ctx = D.ctx;
		const uint32_t rowBegin = p.rowEnd ? p.rowBegin : 0u, rowEnd = p.rowEnd ? (p.rowEnd < cam.height ? p.rowEnd : cam.height) : cam.height;
		const uint32_t msBegin = p.msaaEnd ? p.msaaBegin : 0u, msEnd = p.msaaEnd ? (p.msaaEnd < p.msaa ? p.msaaEnd : p.msaa) : p.msaa;
		if (rowBegin >= rowEnd || msBegin >= msEnd) { ctx.error = "empty shard"; return SAILOR_PT_ERR_ARG; }
		if (p.maxBounces > 64u) { ctx.error = "maxBounces > 64 is not supported"; return SAILOR_PT_ERR_LIMIT; }
		if (p.numSamples > 65535u || p.numAmbientSamples > 65535u) { ctx.error = "more than 65535 samples per first hit are not supported"; return SAILOR_PT_ERR_LIMIT; }
		const uint32_t rows = rowEnd - rowBegin, ns = msEnd - msBegin;
		const uint64_t tiles = (uint64_t)((cam.width + 7u) / 8u) * ((rows + 3u) / 4u);
		const uint64_t total = tiles * ns * 32ull;
		const uint64_t realSamples = (uint64_t)rows * cam.width * ns;
		if (total >= 0xFFFFFF00ull) { ctx.error = "shard too large for one launch: split rows or samples"; return SAILOR_PT_ERR_LIMIT; }

		SharedArenas& arenas = ArenasOfCurrentDevice();
		std::lock_guard<std::mutex> frameLock(arenas.frame);
		PlainBuf* renderMem = arenas.mem;
		float* sampleBuf = EnsureBytes<float>(ctx, renderMem[4], (size_t)realSamples * 3);
		PrimaryHitRec* queue = EnsureBytes<PrimaryHitRec>(ctx, renderMem[5], (size_t)realSamples);
		BatchCounters* counters = EnsureBytes<BatchCounters>(ctx, renderMem[6], 1);
		if (renderMem[7].n == 0) { uint16_t* b = EnsureBytes<uint16_t>(ctx, renderMem[7], kBlueNoiseCount); DevUpload(ctx, b, kBlueNoiseK, sizeof(kBlueNoiseK)); }
		const uint16_t* blue = reinterpret_cast<const uint16_t*>(renderMem[7].p);
		uint32_t* hitCounter = D.counter.p + 2;
		if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
		DevMemset(ctx, hitCounter, 0, sizeof(uint32_t));

		rs = RenderStats{};
		const BvhView view = D.View();
		// secondary rays take the origin-local walk (or, on request, the wide layout) unless the caller asks for the reference visit
		// order everywhere or the scene is small enough for the shared-memory kernel; ambiguous rays are replayed exactly either way
		const bool exactOnly = (p.flags & SAILOR_PT_FLAG_EXACT_TRAVERSAL) != 0u;
		if (!exactOnly && (p.flags & SAILOR_PT_FLAG_WIDE_TRAVERSAL)) { const int rcw = D.EnsureWide(); if (rcw != SAILOR_PT_OK) return rcw; }
		const bool useWide = !exactOnly && (p.flags & SAILOR_PT_FLAG_WIDE_TRAVERSAL) && D.hasWide;
		bool useFast = !exactOnly && !useWide && D.hasFast;
		// The origin-local walk wins where occluders are near the ray origin (a 1M-triangle terrain: x1.1 over the top-down kernel) and loses
		// where rays cross the scene (a gallery of boxes under far emitters: x0.7).  Which it is depends on the scene, not on its size, so the
		// first frame of a scene times both kernels on the first rays of its first level and the scene keeps the winner (same bits either way).
		// SAILOR_PT_TRAVERSAL=local|exact skips the probe.
		bool probeTraversal = false;
		if (useFast)
		{
			if (D.traceChoiceTris != D.numTris) { D.traceChoice = D.Host().traversalChoice; D.traceChoiceTris = D.numTris; }
			if (const char* e = getenv("SAILOR_PT_TRAVERSAL")) D.traceChoice = e[0] == 'e' ? 2u : 1u;
			probeTraversal = D.traceChoice == 0u;
			if (D.traceChoice == 2u) useFast = false;
		}
		const WideView wide = D.Wide();
		const FastView fast = D.Fast();
		DevMemset(ctx, D.counter.p + 15, 0, sizeof(uint32_t));
		SpanTimer& tt = D.traceTimer;
		SpanTimer* st = D.stageTimer;

		const bool hostTrace = getenv("SAILOR_PT_TRACE_HOST") != nullptr;
		const double tFrame0 = HostNow();
		ctx.Mark(0);
		// ---- primary pass ----
		PrimaryArgs pa;
		pa.bvh = view; pa.cam = cam; pa.rowBegin = rowBegin; pa.rowEnd = rowEnd; pa.msBegin = msBegin; pa.msEnd = msEnd; pa.msaa = p.msaa; pa.seed = p.seed;
		pa.ambient = v3(p.ambient[0], p.ambient[1], p.ambient[2]); pa.sampleBuf = sampleBuf; pa.queue = queue; pa.hitCount = hitCounter; pa.total = (uint32_t)total;
		tt.Begin(ctx);
		LaunchPrimaryPass(ctx, pa, D.counter.p);
		tt.End(ctx);
		uint32_t hitCount = 0;
		bool resolved = false;
		DevDownload(ctx, &hitCount, hitCounter, 4);
		rs.rays = realSamples; rs.primarySamples = realSamples;

		if (hitCount && ctx.ok)
		{
			const uint32_t numLights = (uint32_t)D.Host().lights.size();
			const bool ambientOn = pa.ambient.x + pa.ambient.y + pa.ambient.z > 0.0f;
			const bool hasSky = ambientOn && SceneHasThickTransmission(D.Host());
			const uint32_t levels = p.maxBounces + 1u;
			uint32_t shrink = 0;
			uint32_t done = 0;
			uint64_t held = 0;
			for (const int k : { 0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15 }) held += renderMem[k].n;
			const uint64_t budget = BatchBudget(held);
			if (hostTrace) fprintf(stderr, "[sailor_pt] t=%.3f ms primary pass done: %u first hits\n", (HostNow() - tFrame0) * 1e3, hitCount);
			while (done < hitCount && ctx.ok)
			{
				BatchPlan plan = PlanBatch(budget, hitCount - done, numLights, p.numAmbientSamples, p.numSamples, p.maxBounces, ambientOn, shrink);
				// test hook (tests/test_*: the overflow-and-retry path): the first attempt of every batch gets arenas an eighth of the plan, overflows, and is redone
				if (shrink == 0u && plan.firstHits > 1024u && getenv("SAILOR_PT_TEST_OVERFLOW")) { plan.auxCap = plan.auxCap / 8u + 64u; plan.recCap = plan.recCap / 8u + 64u; }
				IntegratorArgs a;
				a.shade = D.shade.p; a.centroid = D.centroid.p; a.materials = D.materials.p; a.tex.texels = D.texels.p; a.tex.textures = D.textures.p; a.tex.srgbLut = D.srgbLut.p; a.tex.texelsF = D.texelsF.p;
				a.lights = D.lights.p; a.numLights = numLights; a.blueNoise = blue;
				a.cam = cam; a.rowBegin = rowBegin; a.rowEnd = rowEnd; a.msBegin = msBegin; a.msEnd = msEnd; a.msaa = p.msaa;
				a.maxBounces = p.maxBounces; a.numSamples = p.numSamples; a.numAmbientSamples = p.numAmbientSamples;
				a.ambient = pa.ambient; a.seed = p.seed;
				a.hitQueue = queue; a.queueBegin = done; a.queueCount = plan.firstHits;
				a.recs = EnsureBytes<NodeRec>(ctx, renderMem[0], plan.recCap); a.recCap = plan.recCap;
				a.aux = EnsureBytes<RayAux>(ctx, renderMem[1], plan.auxCap); a.auxCap = plan.auxCap; a.hasSky = hasSky ? 1u : 0u;
				a.rays = EnsureBytes<RayRec>(ctx, renderMem[2], plan.rayCap); a.slow = EnsureBytes<SlowRec>(ctx, renderMem[13], plan.rayCap); a.rayCap = plan.rayCap;
				a.hits = reinterpret_cast<Hit*>(a.slow);          // plain hit records: only the traversal probe below writes them, before the level's own trace
				a.skyCap = hasSky ? plan.skyCap : 16u;
				a.sky[0] = EnsureBytes<SkyState>(ctx, renderMem[9], (size_t)a.skyCap * 2); a.sky[1] = a.sky[0] + a.skyCap;
				a.skyRays = EnsureBytes<RayRec>(ctx, renderMem[10], (size_t)a.skyCap * 2); a.skyHits = EnsureBytes<Hit>(ctx, renderMem[11], a.skyCap);
				a.status = EnsureBytes<uint8_t>(ctx, renderMem[12], plan.auxCap);
				// the fan-out slot words keep the ShadeCtx index in 24 bits (integrator.cuh): activations beyond 2^24 per batch take the inline path
				a.fanCap = plan.firstHits + 1024u < (1u << 24) ? plan.firstHits + 1024u : (1u << 24);
				a.fan = EnsureBytes<ShadeCtx>(ctx, renderMem[8], a.fanCap);
				a.fanSlots[0] = EnsureBytes<uint32_t>(ctx, renderMem[14], (size_t)a.fanCap * 8u); a.fanSlots[1] = a.fanSlots[0] + (size_t)a.fanCap * 4u;
				const uint32_t replayCap = plan.rayCap > a.skyCap ? plan.rayCap : a.skyCap;
				uint32_t* replayList = (useWide || useFast) ? EnsureBytes<uint32_t>(ctx, renderMem[15], replayCap) : nullptr;
				const ReplayBuffers wb = D.Replay(replayList, replayCap);
				a.c = counters; a.sampleBuf = sampleBuf;
				if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
				if (hostTrace) fprintf(stderr, "[sailor_pt] t=%.3f ms batch: first hits %u of %u (done %u), budget %.1f GiB, rayCap %u auxCap %u recCap %u shrink %u\n",
					(HostNow() - tFrame0) * 1e3, plan.firstHits, hitCount, done, (double)budget / (1 << 30), plan.rayCap, plan.auxCap, plan.recCap, shrink);

				launch_for(ctx, 1, BeginBatchKernel{ counters, plan.firstHits });
				uint32_t usedLevels = levels;
				uint32_t* pin = ctx.Pinned();
				constexpr int kMarkLevel = 10;                                  // markers 10 / 11: level read-backs (alternating)
				if (!ctx.ok || !pin) return SAILOR_PT_ERR_CUDA;
				for (uint32_t level = 0; level < levels; level++)
				{
					const LevelInfo* L = &counters->level[level];
					// upper bounds for the grid: level 0 is exact, deeper levels are bounded by the arenas
					const uint32_t maxRecs = level == 0 ? plan.firstHits : plan.recCap;
					st[0].Begin(ctx);
					launch_for_range<SPT_EXPAND_MIN_BLOCKS>(ctx, &L->recBegin, &L->recEnd, plan.recCap, maxRecs, ExpandKernel{ a, level });
					st[0].End(ctx);
					st[1].Begin(ctx);
					launch_for_range<SPT_FAN_MIN_BLOCKS>(ctx, &counters->zero, &counters->fanThreads[0], a.fanCap * 32u, a.fanCap * 32u, FanOutKernel{ a, level, 0u });
					launch_for_range<SPT_FAN_MIN_BLOCKS>(ctx, &counters->zero, &counters->fanThreads[1], a.fanCap * 32u, a.fanCap * 32u, FanOutKernel{ a, level, 1u });
					st[1].End(ctx);
					launch_for(ctx, 1, OverflowGuardKernel{ counters, level });
					if (probeTraversal && level == 0u)
					{
						// hit-only launches (QueueSink) over the level's rays (at most 64 M: a short prefix is not representative, it holds the rays of
						// the first image rows only): no status bytes, no slow list; the level launch below traces them again.  One round: tens of
						// milliseconds, once per scene file.
						probeTraversal = false;
						const uint32_t probeN = plan.rayCap < (64u << 20) ? plan.rayCap : (64u << 20);
						double best[2] = { 1e30, 1e30 };
						for (int round = 0; round < 1 && ctx.ok; round++)
						{
							ctx.TimerStart(); LaunchTraceRaysFast(ctx, fast, view, wb, a.rays, a.hits, probeN, &L->rayCount); const double tf = ctx.TimerStop();
							ctx.TimerStart(); LaunchTraceRays(ctx, view, a.rays, a.hits, probeN, D.counter.p, &L->rayCount); const double te = ctx.TimerStop();
							if (tf < best[0]) best[0] = tf;
							if (te < best[1]) best[1] = te;
						}
						DevMemset(ctx, D.counter.p + 15, 0, sizeof(uint32_t));          // the probe's replays are not the frame's
						D.probeMs[0] = (float)(best[0] * 1e3); D.probeMs[1] = (float)(best[1] * 1e3);
						D.traceChoice = best[1] < best[0] * 0.97 ? 2u : 1u;
						D.Host().traversalChoice = D.traceChoice;
						useFast = D.traceChoice == 1u;
						if (hostTrace) fprintf(stderr, "[sailor_pt] traversal probe: origin-local %.3f ms, exact %.3f ms -> %s\n", best[0] * 1e3, best[1] * 1e3, useFast ? "origin-local" : "exact");
					}
					tt.Begin(ctx);
					if (useFast) LaunchTraceLevelFast(ctx, fast, view, wb, a.rays, a.hits, plan.rayCap, &L->rayCount, WavefrontOut{ a.status, &L->auxBase, a.slow, &counters->slowCount });
					else if (useWide) LaunchTraceLevelWide(ctx, wide, view, wb, a.rays, a.hits, plan.rayCap, &L->rayCount, WavefrontOut{ a.status, &L->auxBase, a.slow, &counters->slowCount });
					else LaunchTraceLevel(ctx, view, a.rays, a.hits, plan.rayCap, D.counter.p, &L->rayCount, WavefrontOut{ a.status, &L->auxBase, a.slow, &counters->slowCount });
					tt.End(ctx);
					st[2].Begin(ctx);
					launch_for_range<SPT_CLASSIFY_MIN_BLOCKS>(ctx, &counters->zero, &counters->slowCount, plan.rayCap, plan.rayCap, ClassifyKernel{ a, level });
					st[2].End(ctx);
					if (hasSky)
					{
						uint32_t q = 0;
						for (uint32_t it = 0; it < p.maxBounces; it++, q ^= 1u)     // a TraceSky walk traces at most maxBounces rays (:581)
						{
							tt.Begin(ctx);
							if (useFast) LaunchTraceRaysFast(ctx, fast, view, wb, a.skyRays + (q ? a.skyCap : 0u), a.skyHits, a.skyCap, &counters->skyCount[q]);
							else if (useWide) LaunchTraceRaysWide(ctx, wide, view, wb, a.skyRays + (q ? a.skyCap : 0u), a.skyHits, a.skyCap, &counters->skyCount[q]);
							else LaunchTraceRays(ctx, view, a.skyRays + (q ? a.skyCap : 0u), a.skyHits, a.skyCap, D.counter.p, &counters->skyCount[q]);
							tt.End(ctx);
							st[2].Begin(ctx);
							launch_for_range(ctx, &counters->zero, &counters->skyCount[q], a.skyCap, a.skyCap, SkyKernel{ a, q });
							launch_for(ctx, 1, SkySwapKernel{ counters, q });
							st[2].End(ctx);
						}
					}
					launch_for(ctx, 1, NextLevelKernel{ counters, level, plan.recCap });
					// A level without activations ends the batch (paths that left the scene, the 0.01 throughput cut).  Its record range is
					// read back WITHOUT draining the queue: the copy for level L+1 is queued behind level L, and the host looks at it only
					// after it has queued level L+1 as well, so the GPU works on the next level while the host round trip happens.  The price
					// is one level of empty launches at the end of a batch (a few microseconds) instead of an idle GPU at every level.
					if (level + 1u < levels)
					{
						ctx.ReadAsync(pin + 2u * (level + 1u), &counters->level[level + 1u].recBegin, 2u * sizeof(uint32_t));
						ctx.Mark(kMarkLevel + (int)((level + 1u) & 1u));
					}
					if (level >= 1u)
					{
						ctx.WaitMark(kMarkLevel + (int)(level & 1u));
						if (!ctx.ok) break;
						const uint32_t nAct = pin[2u * level + 1u] > pin[2u * level] ? pin[2u * level + 1u] - pin[2u * level] : 0u;
						if (hostTrace) fprintf(stderr, "[sailor_pt] t=%.3f ms   level %u was queued with %u activations\n", (HostNow() - tFrame0) * 1e3, level, nAct);
						if (!nAct) { usedLevels = level; break; }              // this level's launches found nothing to do
					}
				}
				st[3].Begin(ctx);
				for (uint32_t level = usedLevels; level-- > 0;)
				{
					const LevelInfo* L = &counters->level[level];
					launch_for_range<SPT_GATHER_MIN_BLOCKS>(ctx, &L->recBegin, &L->recEnd, plan.recCap, level == 0 ? plan.firstHits : plan.recCap, GatherKernel{ a });
				}
				st[3].End(ctx);
				// The batch's counters (overflow flag, ray count) are read back with a host round trip.  For the LAST batch of the frame
				// the resolve kernel is queued first, so the GPU never idles on that read.
				const bool lastBatch = (uint64_t)done + plan.firstHits >= hitCount && !prog.running;   // a running sum must be updated exactly once: only after the batch is known to be valid
				if (lastBatch) launch_for(ctx, rows * cam.width, ResolveKernel{ sampleBuf, dImage, cam.width, cam.height, rowBegin, rowEnd, ns, prog.norm ? prog.norm : p.msaa, prog.running, prog.runningValid ? 1u : 0u });
				struct { uint32_t recAlloc, auxAlloc, overflow, sky0, sky1, zero, fanEntries, slowCount, fan0, fan1; unsigned long long rays, fanSamples; } head;
				DevDownload(ctx, &head, counters, sizeof(head));          // synchronises
				if (!ctx.ok) break;
				if (hostTrace) fprintf(stderr, "[sailor_pt] t=%.3f ms   batch done (%u levels)\n", (HostNow() - tFrame0) * 1e3, usedLevels);
				if (head.overflow)
				{
					if (plan.firstHits <= 1024u || shrink > 16u) { ctx.error = "wavefront arenas overflow even for the smallest batch"; return SAILOR_PT_ERR_LIMIT; }
					shrink++;
					continue;                                                // redo this batch smaller (results are keyed per activation, not per batch); the resolve is redone too
				}
				rs.rays += head.rays; rs.fanOutSamples += head.fanSamples; rs.batches++;
				done += plan.firstHits;
				resolved = lastBatch;
			}
		}
		if (!resolved) launch_for(ctx, rows * cam.width, ResolveKernel{ sampleBuf, dImage, cam.width, cam.height, rowBegin, rowEnd, ns, prog.norm ? prog.norm : p.msaa, prog.running, prog.runningValid ? 1u : 0u });
		ctx.Mark(1);
		ctx.Sync();
		rs.traverseLaunches = tt.Spans();
		{ uint32_t rep = 0; DevDownload(ctx, &rep, D.counter.p + 15, 4); rs.replayedRays = rep; }
		rs.secondsTraverse = tt.Collect(ctx);
		for (int k = 0; k < 4; k++) rs.secondsStage[k] = st[k].Collect(ctx);
		rs.secondsShade = ctx.Between(0, 1) - rs.secondsTraverse;      // everything of the frame that is not a trace launch
		return ctx.ok ? SAILOR_PT_OK : SAILOR_PT_ERR_CUDA;
	}
}
