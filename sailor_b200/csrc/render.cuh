// render.cuh — host driver of the wavefront loop (tile loop of PathTracer.cpp:418-487 becomes trace/advance launches).
#pragma once
#include "integrator.cuh"

namespace spt
{
	inline uint32_t PoolSizeFor(uint64_t totalSamples)
	{
		uint64_t pool = 1u << 21;                               // 2 Mi resident paths: ~14k per SM, enough to hide DRAM latency
		if (const char* e = getenv("SAILOR_PT_POOL")) { const long v = atol(e); if (v > 0) pool = (uint64_t)v; }
		if (pool > totalSamples) pool = totalSamples;
		pool = (pool + 255u) & ~255ull;
		return (uint32_t)pool;
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
		if (total >= 0xFFFFFF00ull) { ctx.error = "shard too large for one launch: split rows or samples"; return SAILOR_PT_ERR_LIMIT; }
		const uint32_t pool = PoolSizeFor(total);
		const uint32_t maxDepth = p.maxBounces + 1u;

		DevBuf<PathHeader> headers; DevBuf<Frame> frames; DevBuf<RayRec> rays; DevBuf<Hit> hits; DevBuf<float> sampleBuf;
		DevBuf<uint32_t> counters; DevBuf<unsigned long long> counters64; DevBuf<uint16_t> blue;
		headers.Alloc(ctx, pool); frames.Alloc(ctx, (size_t)pool * maxDepth); rays.Alloc(ctx, pool); hits.Alloc(ctx, pool);
		sampleBuf.Alloc(ctx, (size_t)rows * cam.width * ns * 3);
		counters.Alloc(ctx, 4); counters64.Alloc(ctx, 2);
		blue.Upload(ctx, kBlueNoiseK, kBlueNoiseCount);
		if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
		counters.Zero(ctx); counters64.Zero(ctx); headers.Zero(ctx);

		IntegratorArgs a;
		a.shade = D.shade.p; a.centroid = D.centroid.p; a.materials = D.materials.p; a.tex.texels = D.texels.p; a.tex.textures = D.textures.p;
		a.lights = D.lights.p; a.numLights = (uint32_t)D.host.lights.size(); a.blueNoise = blue.p;
		a.cam = cam; a.rowBegin = rowBegin; a.rowEnd = rowEnd; a.msBegin = msBegin; a.msEnd = msEnd; a.msaa = p.msaa;
		a.maxBounces = p.maxBounces; a.numSamples = p.numSamples; a.numAmbientSamples = p.numAmbientSamples;
		a.ambient = v3(p.ambient[0], p.ambient[1], p.ambient[2]); a.seed = p.seed;
		a.poolSize = pool; a.maxDepth = maxDepth;
		a.headers = headers.p; a.frames = frames.p; a.rays = rays.p; a.hits = hits.p; a.sampleBuf = sampleBuf.p;
		a.nextSample = counters.p; a.totalSamples = (uint32_t)total; a.activeCount = counters.p + 1;
		a.rayCount = counters64.p; a.sampleCount = counters64.p + 1;

		rs = RenderStats{};
		const BvhView view = D.View();
		uint32_t active = 0;
		launch_for(ctx, pool, AdvanceKernel{ a, 1u });            // fill the pool with primary rays
		DevDownload(ctx, &active, a.activeCount, 4);
		while (active && ctx.ok)
		{
			ctx.Mark(0);
			LaunchTraceRays(ctx, view, rays.p, hits.p, pool, D.counter.p);
			ctx.Mark(1);
			DevMemset(ctx, a.activeCount, 0, 4);
			launch_for(ctx, pool, AdvanceKernel{ a, 0u });
			ctx.Mark(2);
			DevDownload(ctx, &active, a.activeCount, 4);          // synchronises
			rs.secondsTraverse += ctx.Between(0, 1); rs.secondsShade += ctx.Between(1, 2); rs.traverseLaunches++;
		}
		launch_for(ctx, rows * cam.width, ResolveKernel{ sampleBuf.p, dImage, cam.width, cam.height, rowBegin, rowEnd, ns, p.msaa });
		unsigned long long c64[2] = { 0, 0 };
		DevDownload(ctx, c64, counters64.p, sizeof(c64));
		rs.rays = c64[0]; rs.primarySamples = c64[1];
		return ctx.ok ? SAILOR_PT_OK : SAILOR_PT_ERR_CUDA;
	}
}
