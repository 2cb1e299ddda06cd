// render.cuh — host driver of one frame (the tile loop of PathTracer.cpp:418-487 becomes kernel launches):
//
//   primary pass   one fused kernel: camera ray generation + closest hit for EVERY (pixel, primary sample) of the
//                  shard.  Misses are final (Raytrace returns the ambient colour, :873-876) and are written straight
//                  to the per-sample buffer; hits are appended to a compact first-hit queue (warp-aggregated atomics).
//   wavefront      pool slots pull first hits from the queue and run the Raytrace state machine (integrator.cuh):
//                  [trace kernel, advance kernel] per iteration, launched in chunks without host synchronisation;
//                  the host only reads one 4-byte "slots still active" word per chunk.
//   resolve        accumulator = sum over the pixel's primary samples in index order / msaa, row flip (:449,468-469).
//
// All working buffers live with the scene and are reused across frames (no allocation in the steady state).
#pragma once
#include "integrator.cuh"

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
	__global__ void __launch_bounds__(kTraceBlock) k_primary_pass(PrimaryArgs a, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kStackDepth * kTraceBlock];
		SmemStack stack; stack.base = stackMem + threadIdx.x; stack.n = 0;
		const uint32_t lane = threadIdx.x & 31;
		for (;;)
		{
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(counter, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= a.total) break;
			uint32_t x, y, sample;
			PrimaryHitRec rec;
			bool hit = false;
			if (DecodePrimary(a, base + lane, x, y, sample)) hit = PrimarySample(a, x, y, sample, stack, rec);
			// warp-aggregated append: one atomic per warp, the warp's hits stay contiguous in the queue
			const uint32_t m = __ballot_sync(0xffffffffu, hit);
			if (m)
			{
				const int leader = __ffs(m) - 1;
				uint32_t qbase = 0;
				if ((int)lane == leader) qbase = atomicAdd(a.hitCount, (uint32_t)__popc(m));
				qbase = __shfl_sync(0xffffffffu, qbase, leader);
				if (hit)
				{
					const uint32_t slot = qbase + (uint32_t)__popc(m & ((1u << lane) - 1u));
					float4* q = reinterpret_cast<float4*>(a.queue + slot);
					q[0] = make_float4(__uint_as_float(rec.pixel), __uint_as_float(rec.sample), 0.0f, 0.0f);
					q[1] = make_float4(rec.t, rec.u, rec.v, __uint_as_float(rec.tri));
				}
			}
		}
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

	inline uint32_t PoolLimit()
	{
		long pool = 1l << 21;                                   // 2 Mi resident paths (~14k per SM): hides DRAM latency, bounds memory
		if (const char* e = getenv("SAILOR_PT_POOL")) { const long v = atol(e); if (v > 0) pool = v; }
		return (uint32_t)pool;
	}

	template<class T> inline T* EnsureBytes(Ctx& ctx, DevBuf<unsigned char>& b, size_t count)
	{
		b.Ensure(ctx, count * sizeof(T) + 16);
		return reinterpret_cast<T*>(b.p);
	}

	inline int RenderFrame(SceneDevice& D, const CameraGpu& cam, const SailorPtParams& p, float* dImage, RenderStats& rs)
	{
		Ctx& ctx = D.ctx;
		const uint32_t rowBegin = p.rowEnd ? p.rowBegin : 0u, rowEnd = p.rowEnd ? (p.rowEnd < cam.height ? p.rowEnd : cam.height) : cam.height;
		const uint32_t msBegin = p.msaaEnd ? p.msaaBegin : 0u, msEnd = p.msaaEnd ? (p.msaaEnd < p.msaa ? p.msaaEnd : p.msaa) : p.msaa;
		if (rowBegin >= rowEnd || msBegin >= msEnd) { ctx.error = "empty shard"; return SAILOR_PT_ERR_ARG; }
		const uint32_t rows = rowEnd - rowBegin, ns = msEnd - msBegin;
		const uint64_t tiles = (uint64_t)((cam.width + 7u) / 8u) * ((rows + 3u) / 4u);
		const uint64_t total = tiles * ns * 32ull;
		const uint64_t realSamples = (uint64_t)rows * cam.width * ns;
		if (total >= 0xFFFFFF00ull) { ctx.error = "shard too large for one launch: split rows or samples"; return SAILOR_PT_ERR_LIMIT; }
		const uint32_t maxDepth = p.maxBounces + 1u;

		float* sampleBuf = EnsureBytes<float>(ctx, D.renderMem[4], (size_t)realSamples * 3);
		PrimaryHitRec* queue = EnsureBytes<PrimaryHitRec>(ctx, D.renderMem[5], (size_t)realSamples);
		uint32_t* counters = EnsureBytes<uint32_t>(ctx, D.renderMem[6], 64);
		unsigned long long* counters64 = reinterpret_cast<unsigned long long*>(counters + 32);
		if (D.renderMem[7].n == 0) { uint16_t* b = EnsureBytes<uint16_t>(ctx, D.renderMem[7], kBlueNoiseCount); DevUpload(ctx, b, kBlueNoiseK, sizeof(kBlueNoiseK)); }
		const uint16_t* blue = reinterpret_cast<const uint16_t*>(D.renderMem[7].p);
		if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
		DevMemset(ctx, counters, 0, 64 * sizeof(uint32_t));

		rs = RenderStats{};
		const BvhView view = D.View();

		// ---- primary pass ----
		PrimaryArgs pa;
		pa.bvh = view; pa.cam = cam; pa.rowBegin = rowBegin; pa.rowEnd = rowEnd; pa.msBegin = msBegin; pa.msEnd = msEnd; pa.msaa = p.msaa; pa.seed = p.seed;
		pa.ambient = v3(p.ambient[0], p.ambient[1], p.ambient[2]); pa.sampleBuf = sampleBuf; pa.queue = queue; pa.hitCount = counters + 2; pa.total = (uint32_t)total;
		ctx.Mark(0);
		LaunchPrimaryPass(ctx, pa, D.counter.p);
		ctx.Mark(1);
		uint32_t hitCount = 0;
		DevDownload(ctx, &hitCount, counters + 2, 4);
		rs.secondsTraverse += ctx.Between(0, 1); rs.traverseLaunches++;
		rs.rays = realSamples; rs.primarySamples = realSamples;

		if (hitCount && ctx.ok)
		{
			uint32_t pool = PoolLimit();
			if (pool > hitCount) pool = hitCount;
			pool = (pool + 255u) & ~255u;
			PathHeader* headers = EnsureBytes<PathHeader>(ctx, D.renderMem[0], pool);
			Frame* frames = EnsureBytes<Frame>(ctx, D.renderMem[1], (size_t)pool * maxDepth);
			RayRec* rays = EnsureBytes<RayRec>(ctx, D.renderMem[2], pool);
			Hit* hits = EnsureBytes<Hit>(ctx, D.renderMem[3], pool);
			if (!ctx.ok) return SAILOR_PT_ERR_CUDA;

			IntegratorArgs a;
			a.shade = D.shade.p; a.centroid = D.centroid.p; a.materials = D.materials.p; a.tex.texels = D.texels.p; a.tex.textures = D.textures.p;
			a.lights = D.lights.p; a.numLights = (uint32_t)D.host.lights.size(); a.blueNoise = blue;
			a.cam = cam; a.rowBegin = rowBegin; a.rowEnd = rowEnd; a.msBegin = msBegin; a.msEnd = msEnd; a.msaa = p.msaa;
			a.maxBounces = p.maxBounces; a.numSamples = p.numSamples; a.numAmbientSamples = p.numAmbientSamples;
			a.ambient = pa.ambient; a.seed = p.seed;
			a.poolSize = pool; a.maxDepth = maxDepth;
			a.headers = headers; a.frames = frames; a.rays = rays; a.hits = hits; a.sampleBuf = sampleBuf;
			a.hitQueue = queue; a.queueCount = hitCount; a.nextSample = counters; a.rayCount = counters64;
			uint32_t* activeSlots = counters + 8;                        // one word per iteration of a chunk
			a.activeCount = activeSlots;

			uint32_t active = 0;
			launch_for(ctx, pool, AdvanceKernel{ a, 1u });                // fill the pool from the first-hit queue
			DevDownload(ctx, &active, activeSlots, 4);
			const int kChunk = 16;                                       // iterations launched between two host reads
			while (active && ctx.ok)
			{
				DevMemset(ctx, activeSlots, 0, kChunk * sizeof(uint32_t));
				for (int k = 0; k < kChunk; k++)
				{
					ctx.Mark(3 * k);
					LaunchTraceRays(ctx, view, rays, hits, pool, D.counter.p);
					ctx.Mark(3 * k + 1);
					a.activeCount = activeSlots + k;
					launch_for(ctx, pool, AdvanceKernel{ a, 0u });
					ctx.Mark(3 * k + 2);
				}
				uint32_t act[kChunk];
				DevDownload(ctx, act, activeSlots, sizeof(act));          // synchronises
				for (int k = 0; k < kChunk; k++)
				{
					rs.secondsTraverse += ctx.Between(3 * k, 3 * k + 1); rs.secondsShade += ctx.Between(3 * k + 1, 3 * k + 2);
					if (k == 0 || act[k - 1]) rs.traverseLaunches++;        // launches past the end of the frame are empty
				}
				active = act[kChunk - 1];
			}
			unsigned long long c64 = 0;
			DevDownload(ctx, &c64, counters64, sizeof(c64));
			rs.rays += c64;
		}
		launch_for(ctx, rows * cam.width, ResolveKernel{ sampleBuf, dImage, cam.width, cam.height, rowBegin, rowEnd, ns, p.msaa });
		return ctx.ok ? SAILOR_PT_OK : SAILOR_PT_ERR_CUDA;
	}
}
