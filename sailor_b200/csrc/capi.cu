// capi.cu — the C-ABI of include/sailor_pt.h for the product library libsailor_pt_cuda.so.
//
// Thin layer: argument checks, host<->device staging of the caller's buffers, stage sequencing.  All arithmetic of
// the hot path runs in the kernels (flatten.cuh, bvh_build.cuh, traverse.cuh, integrator.cuh, output.cuh).
// There is no CPU fallback: without a CUDA device every computing entry point returns SAILOR_PT_ERR_NO_DEVICE.
#include "pipeline.cuh"
#include "textures.cuh"
#include "integrator.cuh"
#include "render.cuh"
#include "output.cuh"

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

using namespace spt;

// `dev` is the scene on the device it was loaded on; `replicas` are copies on further devices, made on demand by a multi-device
// render (SailorPtParams::deviceCount > 1) from the same parsed file and kept for the following frames.
struct SailorPtScene { SceneDevice dev; std::vector<std::unique_ptr<SceneDevice>> replicas; };

namespace
{
	thread_local std::string t_lastError;
	SailorPtStats g_stats{};

	int SetError(int code, const std::string& msg) { t_lastError = msg; return code; }

	// ---- parsed-scene cache --------------------------------------------------------------------------------------------------
	// A host that renders the same file frame after frame (the reference's PathTracer::Run loads its model on every call,
	// PathTracer.cpp:84-100) pays the glTF parse and the image decodes once: entries are keyed by (real path, mtime, size) of the
	// scene file, at most two are kept, SailorPt_TrimMemory drops them, SAILOR_PT_SCENE_CACHE=0 turns the cache off.  Only the HOST
	// side is cached: every SceneLoad still uploads and flattens on the device.  External .bin / image files are not part of the key.
	struct CachedScene { std::string key; std::shared_ptr<const HostScene> scene; };
	std::mutex g_sceneCacheMutex;
	std::vector<CachedScene> g_sceneCache;

	std::string SceneKey(const char* path)
	{
		struct stat st;
		if (stat(path, &st) != 0) return std::string();
		char real[4096];
		const char* rp = realpath(path, real);
		char buf[96];
		snprintf(buf, sizeof(buf), "|%lld.%09ld|%lld", (long long)st.st_mtim.tv_sec, (long)st.st_mtim.tv_nsec, (long long)st.st_size);
		return std::string(rp ? rp : path) + buf;
	}
	std::shared_ptr<const HostScene> AcquireHostScene(const char* path, int& rc, std::string& err, bool& fromCache)
	{
		const char* env = getenv("SAILOR_PT_SCENE_CACHE");
		const bool enabled = !(env && env[0] == '0');
		const std::string key = enabled ? SceneKey(path) : std::string();
		fromCache = false;
		if (!key.empty())
		{
			std::lock_guard<std::mutex> lock(g_sceneCacheMutex);
			for (const CachedScene& c : g_sceneCache) if (c.key == key) { rc = SAILOR_PT_OK; fromCache = true; return c.scene; }
		}
		std::shared_ptr<HostScene> fresh(new HostScene());
		rc = LoadGltf(path, *fresh, err);
		if (rc != SAILOR_PT_OK) return nullptr;
		if (!key.empty())
		{
			std::lock_guard<std::mutex> lock(g_sceneCacheMutex);
			if (g_sceneCache.size() >= 2u) g_sceneCache.erase(g_sceneCache.begin());
			g_sceneCache.push_back(CachedScene{ key, fresh });
		}
		return fresh;
	}
	void DropSceneCache() { std::lock_guard<std::mutex> lock(g_sceneCacheMutex); g_sceneCache.clear(); }
	int FromCtx(SceneDevice& d, int rc)
	{
		if (!d.ctx.ok) { t_lastError = d.ctx.error; return rc == SAILOR_PT_OK ? SAILOR_PT_ERR_CUDA : rc; }
		if (rc != SAILOR_PT_OK && !d.ctx.error.empty()) t_lastError = d.ctx.error;
		return rc;
	}

	CameraGpu ToGpuCamera(const CameraSetup& c)
	{
		CameraGpu g;
		g.pos = v3(c.pos[0], c.pos[1], c.pos[2]); g.pixel00Dir = v3(c.pixel00Dir[0], c.pixel00Dir[1], c.pixel00Dir[2]);
		g.deltaU = v3(c.deltaU[0], c.deltaU[1], c.deltaU[2]); g.deltaV = v3(c.deltaV[0], c.deltaV[1], c.deltaV[2]);
		g.width = c.width; g.height = c.height;
		return g;
	}

	CameraSetup CameraOf(const SceneDevice& d, const SailorPtParams* p)
	{
		SailorPtParamsView v; v.camera = p->camera; v.height = p->height; v.widthOverride = p->widthOverride;
		return SetupCamera(d.Host(), v);
	}

	// standalone context for entry points that take no scene (OutputStage, EvalLighting)
	struct ScopedCtx
	{
		Ctx ctx; int rc;
		ScopedCtx() { rc = ctx.Init(); }
		~ScopedCtx() { ctx.Destroy(); }
	};
}

extern "C" {

const char* SailorPt_Backend(void)
{
#if defined(SPT_EMU)
	return "emu-host (tests only)";
#else
	return "cuda sm_100a";
#endif
}
const char* SailorPt_LastError(void) { return t_lastError.c_str(); }
int32_t SailorPt_SetDevice(int32_t device)
{
#if defined(SPT_EMU)
	return device == 0 ? SAILOR_PT_OK : SAILOR_PT_ERR_ARG;
#else
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return SetError(SAILOR_PT_ERR_NO_DEVICE, "no CUDA device: the sailor_b200 product library has no CPU path"); }
	if (device < 0 || device >= count) return SetError(SAILOR_PT_ERR_ARG, "device index out of range");
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SetError(SAILOR_PT_ERR_CUDA, "cudaSetDevice failed"); }
	return SAILOR_PT_OK;
#endif
}
int32_t SailorPt_TrimMemory(void)
{
	ScopedCtx sc;
	if (sc.rc != SAILOR_PT_OK) return SetError(sc.rc, sc.ctx.error);
	ReleaseSharedArenas(sc.ctx);
	TrimDevicePool();
	DropSceneCache();
	return sc.ctx.ok ? SAILOR_PT_OK : SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
}
int32_t SailorPt_PinHostBuffer(void* hostBuffer, uint64_t bytes)
{
	if (!hostBuffer || !bytes) return SetError(SAILOR_PT_ERR_ARG, "null buffer");
	return spt::HostPin(hostBuffer, (size_t)bytes) == 0 ? SAILOR_PT_OK : SetError(SAILOR_PT_ERR_CUDA, "cudaHostRegister failed (or the buffer is already pinned with a smaller size)");
}
int32_t SailorPt_UnpinHostBuffer(void* hostBuffer)
{
	if (!hostBuffer) return SetError(SAILOR_PT_ERR_ARG, "null buffer");
	return spt::HostUnpin(hostBuffer) == 0 ? SAILOR_PT_OK : SetError(SAILOR_PT_ERR_ARG, "buffer was not pinned");
}
#if defined(SPT_TRACE_STATS) && !defined(SPT_EMU)
// tuning builds only (tools/trace_variants.py): read and clear the lane-state counters of the traversal warp loop
extern "C" SAILOR_PT_API int32_t SailorPt_DebugTraceStats(unsigned long long* out)
{
	if (cudaMemcpyFromSymbol(out, spt::g_traceStats, sizeof(unsigned long long) * 16) != cudaSuccess) return SAILOR_PT_ERR_CUDA;
	static const unsigned long long zero[16] = {};
	cudaMemcpyToSymbol(spt::g_traceStats, zero, sizeof(zero));
	return SAILOR_PT_OK;
}
#endif
#if defined(SPT_WIDE_LOOP_STATS) && !defined(SPT_EMU)
// tuning builds only (tools/wide_variants.py): read and clear the lane-state counters of the wide traversal loop
extern "C" SAILOR_PT_API int32_t SailorPt_DebugWideStats(unsigned long long* out)
{
	if (cudaMemcpyFromSymbol(out, spt::g_wideLoopStats, sizeof(unsigned long long) * 16) != cudaSuccess) return SAILOR_PT_ERR_CUDA;
	static const unsigned long long zero[16] = {};
	cudaMemcpyToSymbol(spt::g_wideLoopStats, zero, sizeof(zero));
	return SAILOR_PT_OK;
}
#endif
#if defined(SPT_FAST_LOOP_STATS) && !defined(SPT_EMU)
// tuning builds only (tools/fast_variants.py): read and clear the lane-state counters of the origin-local traversal loop
extern "C" SAILOR_PT_API int32_t SailorPt_DebugFastStats(unsigned long long* out)
{
	if (cudaMemcpyFromSymbol(out, spt::g_fastLoopStats, sizeof(unsigned long long) * 64) != cudaSuccess) return SAILOR_PT_ERR_CUDA;
	static const unsigned long long zero[64] = {};
	cudaMemcpyToSymbol(spt::g_fastLoopStats, zero, sizeof(zero));
	return SAILOR_PT_OK;
}
#endif
int32_t SailorPt_GetStats(SailorPtStats* s) { if (!s) return SAILOR_PT_ERR_ARG; *s = g_stats; return SAILOR_PT_OK; }

int32_t SailorPt_ParseCommandLineArgs(SailorPtParams* res, const char** args, int32_t num)
{
	// PathTracer::ParseCommandLineArgs (PathTracer.cpp:30-73) + Utils::GetArgValue (Core/Utils.cpp:466-487)
	if (!res || (!args && num > 0)) return SAILOR_PT_ERR_ARG;
	static thread_local std::string sIn, sOut, sCam;
	auto argValue = [&](int32_t& i) -> std::string
		{
			if (i + 1 >= num) return "";
			std::string v = args[++i];
			if (!v.empty() && v[0] == '\"')
			{
				while (i < num && v[v.length() - 1] != '\"') { ++i; v += " " + std::string(args[i]); }
				v = v.substr(1, v.length() - 2);
			}
			return v;
		};
	for (int32_t i = 1; i < num; i++)
	{
		const std::string arg = args[i];
		if (arg == "--in") { sIn = argValue(i); res->pathToModel = sIn.c_str(); }
		else if (arg == "--out") { sOut = argValue(i); res->output = sOut.c_str(); }
		else if (arg == "--height") res->height = (uint32_t)atoi(argValue(i).c_str());
		else if (arg == "--samples")
		{
			const uint32_t samples = (uint32_t)atoi(argValue(i).c_str());
			res->msaa = samples <= 32 ? (samples < 4u ? samples : 4u) : 8u;
			const long r = lroundf((float)samples / (float)res->msaa);
			res->numSamples = r > 1 ? (uint32_t)r : 1u;
		}
		else if (arg == "--bounces") res->maxBounces = (uint32_t)atoi(argValue(i).c_str());
		else if (arg == "--camera") { sCam = argValue(i); res->camera = sCam.c_str(); }
		else if (arg == "--ambient")
		{
			const std::string hex = argValue(i);
			if (hex.size() < 6) return SAILOR_PT_ERR_ARG;   // the reference would throw from std::stoi here
			const int r = (int)strtol(hex.substr(0, 2).c_str(), nullptr, 16), g = (int)strtol(hex.substr(2, 2).c_str(), nullptr, 16), b = (int)strtol(hex.substr(4, 2).c_str(), nullptr, 16);
			res->ambient[0] = r / 255.0f; res->ambient[1] = g / 255.0f; res->ambient[2] = b / 255.0f;
		}
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneLoad(const char* path, SailorPtScene** outScene)
{
	if (!path || !outScene) return SAILOR_PT_ERR_ARG;
	*outScene = nullptr;
	std::unique_ptr<SailorPtScene> s(new SailorPtScene());
	int rc = s->dev.ctx.Init();
	if (rc != SAILOR_PT_OK) return SetError(rc, s->dev.ctx.error);
	std::string err;
	bool cached = false;
	s->dev.device = DevCurrent();
	s->dev.hostPtr = AcquireHostScene(path, rc, err, cached);
	if (rc != SAILOR_PT_OK) { s->dev.ctx.Destroy(); return SetError(rc, err); }
	s->dev.ctx.kernelLaunches = 0;
	rc = FromCtx(s->dev, s->dev.Upload());
	g_stats = s->dev.stats; g_stats.kernelLaunches = s->dev.ctx.kernelLaunches; g_stats.h2dBytes = s->dev.ctx.h2dBytes;
	if (rc != SAILOR_PT_OK) { s->dev.ctx.Destroy(); return rc; }
	*outScene = s.release();
	return SAILOR_PT_OK;
}

void SailorPt_SceneFree(SailorPtScene* s)
{
	if (!s) return;
	const int home = DevCurrent();
	for (auto& r : s->replicas)
	{
		if (!r) continue;
		DevSetCurrent(r->device);             // a replica's buffers, stream and events belong to its own device
		r->ctx.Sync();
		r->traceTimer.Destroy();
		for (SpanTimer& t : r->stageTimer) t.Destroy();
		Ctx keepR = r->ctx;
		r.reset();
		keepR.Destroy();
	}
	DevSetCurrent(s->dev.device);
	s->dev.ctx.Sync();
	s->dev.traceTimer.Destroy();
	for (SpanTimer& t : s->dev.stageTimer) t.Destroy();
	Ctx keep = s->dev.ctx;
	delete s;
	keep.Destroy();
	DevSetCurrent(home);
}

int32_t SailorPt_SceneCounts(const SailorPtScene* s, uint32_t c[6])
{
	if (!s || !c) return SAILOR_PT_ERR_ARG;
	c[0] = s->dev.numTris; c[1] = (uint32_t)s->dev.Host().materials.size(); c[2] = (uint32_t)s->dev.Host().textures.size();
	c[3] = (uint32_t)s->dev.Host().lights.size(); c[4] = (uint32_t)s->dev.Host().cameras.size(); c[5] = s->dev.built ? s->dev.nodesUsed : 0;
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetMaterials(const SailorPtScene* s, uint32_t* words)
{
	if (!s || !words) return SAILOR_PT_ERR_ARG;
	for (size_t i = 0; i < s->dev.Host().materials.size(); i++)
	{
		const MaterialGpu& m = s->dev.Host().materials[i];
		uint32_t* w = words + i * SAILOR_PT_MATERIAL_WORDS;
		float f[26];
		for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) f[c * 3 + r] = m.uvTransform[c * 4 + r];
		for (int k = 0; k < 4; k++) f[9 + k] = m.baseColor[k];
		for (int k = 0; k < 3; k++) { f[13 + k] = m.emissive[k]; f[16 + k] = m.attenuationColor[k]; }
		f[19] = m.metallic; f[20] = m.roughness; f[21] = m.ior; f[22] = m.transmission; f[23] = m.alphaCutoff; f[24] = m.thickness; f[25] = m.attenuationDistance;
		memcpy(w, f, sizeof(f));
		w[26] = m.blendMode; w[27] = m.texBase; w[28] = m.texNormal; w[29] = m.texMetallicRoughness; w[30] = m.texEmissive; w[31] = m.texTransmission;
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetLights(const SailorPtScene* s, float* out)
{
	if (!s || !out) return SAILOR_PT_ERR_ARG;
	for (size_t i = 0; i < s->dev.Host().lights.size(); i++)
		for (int k = 0; k < 3; k++) { out[i * 6 + k] = s->dev.Host().lights[i].direction[k]; out[i * 6 + 3 + k] = s->dev.Host().lights[i].intensity[k]; }
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetTriangles(const SailorPtScene* cs, float* tris, uint8_t* mat)
{
	if (!cs) return SAILOR_PT_ERR_ARG;
	SailorPtScene* s = const_cast<SailorPtScene*>(cs);
	const uint32_t n = s->dev.numTris;
	if (!n) return SAILOR_PT_OK;
	std::vector<V4> vtx((size_t)n * 3), cen(n), sh((size_t)n * 9); std::vector<V2> uv2((size_t)n * 3);
	s->dev.vtx.Download(s->dev.ctx, vtx.data(), vtx.size()); s->dev.centroid.Download(s->dev.ctx, cen.data(), cen.size());
	s->dev.shade.Download(s->dev.ctx, sh.data(), sh.size()); s->dev.uv2.Download(s->dev.ctx, uv2.data(), uv2.size());
	if (!s->dev.ctx.ok) return FromCtx(s->dev, SAILOR_PT_ERR_CUDA);
	for (uint32_t i = 0; i < n; i++)
	{
		if (tris)
		{
			float* o = tris + (size_t)i * SAILOR_PT_TRI_FLOATS;
			const V4* S = sh.data() + (size_t)i * 9;
			o[0] = cen[i].x; o[1] = cen[i].y; o[2] = cen[i].z;
			for (int k = 0; k < 3; k++) { const V4& v = vtx[(size_t)i * 3 + k]; o[3 + k * 3] = v.x; o[4 + k * 3] = v.y; o[5 + k * 3] = v.z; }
			for (int k = 0; k < 3; k++) { o[12 + k * 3] = S[k].x; o[13 + k * 3] = S[k].y; o[14 + k * 3] = S[k].z; }
			for (int k = 0; k < 3; k++) { o[21 + k * 3] = S[3 + k].x; o[22 + k * 3] = S[3 + k].y; o[23 + k * 3] = S[3 + k].z; }
			for (int k = 0; k < 3; k++) { o[30 + k * 3] = S[6 + k].x; o[31 + k * 3] = S[6 + k].y; o[32 + k * 3] = S[6 + k].z; }
			o[39] = S[0].w; o[40] = S[1].w; o[41] = S[2].w; o[42] = S[3].w; o[43] = S[4].w; o[44] = S[5].w;
			for (int k = 0; k < 3; k++) { o[45 + k * 2] = uv2[(size_t)i * 3 + k].x; o[46 + k * 2] = uv2[(size_t)i * 3 + k].y; }
		}
		if (mat) mat[i] = (uint8_t)f2u(cen[i].w);
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_BuildBVH(SailorPtScene* s)
{
	if (!s) return SAILOR_PT_ERR_ARG;
	const bool was = s->dev.built;
	s->dev.ctx.kernelLaunches = 0;
	const int rc = FromCtx(s->dev, s->dev.BuildBvh());
	if (!was) { g_stats = s->dev.stats; g_stats.kernelLaunches = s->dev.ctx.kernelLaunches; }
	return rc;
}

int32_t SailorPt_GetBVH(const SailorPtScene* cs, SailorPtBvhNode* nodes, uint32_t* mapping)
{
	if (!cs || !cs->dev.built) return SAILOR_PT_ERR_ARG;
	SailorPtScene* s = const_cast<SailorPtScene*>(cs);
	if (nodes) s->dev.refNodes.Download(s->dev.ctx, nodes, (size_t)2 * s->dev.numTris - 1);
	if (mapping) s->dev.mapping.Download(s->dev.ctx, mapping, s->dev.numTris);
	return FromCtx(s->dev, SAILOR_PT_OK);
}

int32_t SailorPt_GetCamera(const SailorPtScene* s, const SailorPtParams* p, uint32_t* w, uint32_t* h, float cam[12])
{
	if (!s || !p || !p->height) return SAILOR_PT_ERR_ARG;
	const CameraSetup c = CameraOf(s->dev, p);
	if (w) *w = c.width;
	if (h) *h = c.height;
	if (cam) { memcpy(cam, c.pos, 12); memcpy(cam + 3, c.pixel00Dir, 12); memcpy(cam + 6, c.deltaU, 12); memcpy(cam + 9, c.deltaV, 12); }
	return SAILOR_PT_OK;
}

int32_t SailorPt_IntersectRays(SailorPtScene* s, uint32_t count, const float* o, const float* d, const uint32_t* ignore, SailorPtHit* hits)
{
	return SailorPt_IntersectRaysEx(s, count, o, d, ignore, 0u, hits);
}

int32_t SailorPt_IntersectRaysEx(SailorPtScene* s, uint32_t count, const float* o, const float* d, const uint32_t* ignore, uint32_t flags, SailorPtHit* hits)
{
	if (!s || !o || !d || !hits) return SAILOR_PT_ERR_ARG;
	int rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	if (!count) return SAILOR_PT_OK;
	SceneDevice& D = s->dev;
	const double t0 = HostNow();
	D.ctx.kernelLaunches = 0;
	std::vector<RayRec> rays(count);
	for (uint32_t i = 0; i < count; i++)
	{
		RayRec& r = rays[i];
		r.ox = o[3 * i]; r.oy = o[3 * i + 1]; r.oz = o[3 * i + 2]; r.ignoreTri = ignore ? ignore[i] : kNoHit;
		r.dx = d[3 * i]; r.dy = d[3 * i + 1]; r.dz = d[3 * i + 2]; r.tmax = (flags & SAILOR_PT_RAYS_ANY_HIT) ? -kFltMax : kFltMax;
	}
	if (flags & SAILOR_PT_RAYS_WIDE) { rc = FromCtx(D, D.EnsureWide()); if (rc != SAILOR_PT_OK) return rc; }
	const bool useWide = (flags & SAILOR_PT_RAYS_WIDE) && D.hasWide;
	const bool useFast = !useWide && (flags & SAILOR_PT_RAYS_LOCAL) && D.hasFast;
	DevBuf<RayRec> dRays; DevBuf<Hit> dHits;
	dRays.Upload(D.ctx, rays); dHits.Alloc(D.ctx, count);
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	if (useWide || useFast) { D.replayList.Ensure(D.ctx, count); DevMemset(D.ctx, D.counter.p + 15, 0, sizeof(uint32_t)); }
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	D.ctx.TimerStart();
	if (useWide) LaunchTraceRaysWide(D.ctx, D.Wide(), D.View(), D.Replay(D.replayList.p, count), dRays.p, dHits.p, count);
	else if (useFast) LaunchTraceRaysFast(D.ctx, D.Fast(), D.View(), D.Replay(D.replayList.p, count), dRays.p, dHits.p, count);
	else LaunchTraceRays(D.ctx, D.View(), dRays.p, dHits.p, count, D.counter.p);
	const double tk = D.ctx.TimerStop();
	dHits.Download(D.ctx, reinterpret_cast<Hit*>(hits), count);
	uint32_t replayed = 0;
	if (useWide || useFast) DevDownload(D.ctx, &replayed, D.counter.p + 15, 4);
	g_stats = SailorPtStats{};
	g_stats.replayedRays = replayed;
#if defined(SPT_EMU) && defined(SPT_WIDE_STATS)
	// host tuning aid: nodes visited / triangles tested by the wide or the origin-local walk
	g_stats.boxTests = g_wideStats[0] + g_fastStats[0]; g_stats.triTests = g_wideStats[1] + g_fastStats[1]; g_stats.threads = (uint32_t)(g_wideStats[3] + g_fastStats[3]);
	g_wideStats[0] = g_wideStats[1] = g_wideStats[3] = 0; g_fastStats[0] = g_fastStats[1] = g_fastStats[3] = 0;
#endif
	g_stats.rays = count; g_stats.secondsTraverse = tk; g_stats.traverseLaunches = 1; g_stats.kernelLaunches = D.ctx.kernelLaunches;
	g_stats.secondsTotal = HostNow() - t0;
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_PrimaryHits(SailorPtScene* s, const SailorPtParams* p, SailorPtHit* hits)
{
	if (!s || !p || !hits || !p->height) return SAILOR_PT_ERR_ARG;
	int rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	SceneDevice& D = s->dev;
	const double t0 = HostNow();
	D.ctx.kernelLaunches = 0;
	const CameraSetup c = CameraOf(D, p);
	const size_t n = (size_t)c.width * c.height;
	if (!n) return SAILOR_PT_ERR_ARG;
	DevBuf<Hit> dHits;
	dHits.Alloc(D.ctx, n);
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	D.ctx.TimerStart();
	LaunchTracePrimary(D.ctx, D.View(), ToGpuCamera(c), dHits.p, D.counter.p);
	const double tk = D.ctx.TimerStop();
	dHits.Download(D.ctx, reinterpret_cast<Hit*>(hits), n);
	g_stats = SailorPtStats{};
	g_stats.rays = n; g_stats.secondsTraverse = tk; g_stats.traverseLaunches = 1; g_stats.kernelLaunches = D.ctx.kernelLaunches;
	g_stats.secondsTotal = HostNow() - t0;
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_OutputStage(uint32_t width, uint32_t height, const float* linearRGB, uint8_t* srgb8)
{
	if (!linearRGB || !srgb8 || !width || !height) return SAILOR_PT_ERR_ARG;
	ScopedCtx sc;
	if (sc.rc != SAILOR_PT_OK) return SetError(sc.rc, sc.ctx.error);
	const size_t n = (size_t)width * height;
	DevBuf<float> dLin; DevBuf<uint8_t> dOut;
	dLin.Upload(sc.ctx, linearRGB, n * 3); dOut.Alloc(sc.ctx, n * 3);
	RunOutputStage(sc.ctx, width, height, dLin.p, dOut.p);
	dOut.Download(sc.ctx, srgb8, n * 3);
	if (!sc.ctx.ok) return SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
	return SAILOR_PT_OK;
}

} // extern "C"

namespace
{
	// ---- one frame over several devices (SailorPtParams::deviceCount > 1) ------------------------------------------------------
	// Replaces the reference's tile loop (PathTracer.cpp:418-487: tasks of 32x32 pixels handed to the worker threads) one level up:
	// the rows of the frame are cut into 4 bands per device, one host thread per device takes bands from a shared counter (so a
	// device that draws cheap rows simply takes more of them), renders each band with the ordinary single-device frame and copies it
	// into the frame of the scene's own device over NVLink (peer copy).  The scene is replicated: every device keeps its own copy of
	// the flattened triangles and builds its own BVH.  RNG streams are keyed by (pixel, primary sample) and a pixel's samples are
	// summed in index order by whichever device owns its row, so the frame has the bit pattern of the single-device render whatever
	// the band assignment was (tests/test_multi_device.py).  The output stage runs afterwards on the scene's own device (its
	// aberration taps cross band borders).  SAILOR_PT_MULTI_SAME_DEVICE=1 puts every replica on the scene's own device (tests on a
	// one-GPU box and the host-compiled build).
	constexpr int kNotMulti = 1;
	constexpr uint32_t kBandsPerDevice = 4;
	struct MultiWorker
	{
		SceneDevice* D = nullptr; int device = 0; int rc = SAILOR_PT_OK; std::string error;
		RenderStats rs{}; double tBuild = 0.0; uint32_t launches = 0, bands = 0; uint64_t h2d = 0;
	};

	int RenderMulti(SailorPtScene* s, const SailorPtParams* p, uint32_t flags)
	{
		SceneDevice& P = s->dev;
		const bool same = getenv("SAILOR_PT_MULTI_SAME_DEVICE") != nullptr;
		const int avail = same ? p->deviceCount : DevCount();
		const CameraSetup c = CameraOf(P, p);
		const size_t n = (size_t)c.width * c.height;
		if (!n) return SAILOR_PT_ERR_ARG;
		const uint32_t R0 = p->rowEnd ? p->rowBegin : 0u, R1 = p->rowEnd ? (p->rowEnd < c.height ? p->rowEnd : c.height) : c.height;
		if (R0 >= R1) return SetError(SAILOR_PT_ERR_ARG, "empty shard");
		const uint32_t rows = R1 - R0;
		int N = p->deviceCount < avail ? p->deviceCount : avail;
		if ((uint32_t)N > rows) N = (int)rows;
		if (N <= 1) return kNotMulti;
		const double t0 = HostNow();
		const int home = DevCurrent();
		if (!DevSetCurrent(P.device)) return SetError(SAILOR_PT_ERR_CUDA, "cannot select the scene's device");
		if (s->replicas.size() < (size_t)(N - 1)) s->replicas.resize((size_t)(N - 1));
		std::vector<MultiWorker> W((size_t)N);
		W[0].D = &P; W[0].device = P.device;
		for (int i = 1; i < N; i++) W[(size_t)i].device = same ? P.device : (P.device + i) % avail;

		P.residentLin.Ensure(P.ctx, n * 3);
		P.residentW = c.width; P.residentH = c.height;
		if (R0 > 0 || R1 < c.height) P.residentLin.Zero(P.ctx, n * 3);          // rows outside the shard stay 0
		P.ctx.Sync();                                                            // the frame exists before any peer writes into it
		if (!P.ctx.ok) { DevSetCurrent(home); return FromCtx(P, SAILOR_PT_ERR_CUDA); }

		uint32_t perDevice = kBandsPerDevice;
		if (const char* e = getenv("SAILOR_PT_BANDS")) { const int v = atoi(e); if (v >= 1 && v <= 64) perDevice = (uint32_t)v; }      // tuning aid
		const uint32_t nBands = rows < (uint32_t)N * perDevice ? rows : (uint32_t)N * perDevice;
		const CameraGpu cam = ToGpuCamera(c);
		std::atomic<uint32_t> next{ 0u };
		const bool wantWide = (p->flags & SAILOR_PT_FLAG_WIDE_TRAVERSAL) != 0u && !(p->flags & SAILOR_PT_FLAG_EXACT_TRAVERSAL);
		auto work = [&](int wi)
		{
			MultiWorker& w = W[(size_t)wi];
			if (!DevSetCurrent(w.device)) { w.rc = SAILOR_PT_ERR_CUDA; w.error = "cannot select device " + std::to_string(w.device); return; }
			if (wi > 0)
			{
				std::unique_ptr<SceneDevice>& rep = s->replicas[(size_t)(wi - 1)];
				if (!rep)
				{
					rep.reset(new SceneDevice());
					rep->device = w.device; rep->hostPtr = P.hostPtr;
					w.rc = rep->ctx.Init();
					if (w.rc == SAILOR_PT_OK) w.rc = rep->Upload();
					if (w.rc == SAILOR_PT_OK && !rep->ctx.ok) w.rc = SAILOR_PT_ERR_CUDA;
					w.h2d += rep->ctx.h2dBytes;
					if (w.rc != SAILOR_PT_OK) { w.error = rep->ctx.error; return; }
				}
				w.D = rep.get();
			}
			SceneDevice& D = *w.D;
			D.ctx.kernelLaunches = 0; D.ctx.h2dBytes = D.ctx.d2hBytes = 0;
			if (flags & 1u) D.built = false;
			D.wantWide = wantWide;
			if (!D.built)
			{
				w.rc = D.BuildBvh();
				if (w.rc == SAILOR_PT_OK && !D.ctx.ok) w.rc = SAILOR_PT_ERR_CUDA;
				if (w.rc != SAILOR_PT_OK) { w.error = D.ctx.error; return; }
				w.tBuild = D.stats.secondsBvhBuild;
			}
			if (wi > 0) D.residentLin.Ensure(D.ctx, n * 3);
			for (;;)
			{
				const uint32_t b = next.fetch_add(1u);
				if (b >= nBands) break;
				SailorPtParams q = *p;
				q.deviceCount = 0;
				q.rowBegin = R0 + (uint32_t)((uint64_t)rows * b / nBands); q.rowEnd = R0 + (uint32_t)((uint64_t)rows * (b + 1u) / nBands);
				RenderStats rs{};
				w.rc = RenderFrame(D, cam, q, D.residentLin.p, rs);
				if (w.rc == SAILOR_PT_OK && !D.ctx.ok) w.rc = SAILOR_PT_ERR_CUDA;
				if (w.rc != SAILOR_PT_OK) { w.error = D.ctx.error; return; }
				if (wi > 0)
				{
					// task row y lands in image row height-1-y (PathTracer.cpp:449): the band is the image rows [height-rowEnd, height-rowBegin)
					const size_t off = (size_t)(c.height - q.rowEnd) * c.width * 3, cnt = (size_t)(q.rowEnd - q.rowBegin) * c.width * 3;
					DevCopyPeer(D.ctx, P.residentLin.p + off, P.device, D.residentLin.p + off, w.device, cnt * sizeof(float));
				}
				w.rs.rays += rs.rays; w.rs.primarySamples += rs.primarySamples; w.rs.secondsTraverse += rs.secondsTraverse; w.rs.secondsShade += rs.secondsShade;
				for (int k = 0; k < 4; k++) w.rs.secondsStage[k] += rs.secondsStage[k];
				w.rs.fanOutSamples += rs.fanOutSamples; w.rs.replayedRays += rs.replayedRays; w.rs.traverseLaunches += rs.traverseLaunches; w.rs.batches += rs.batches;
				w.bands++;
			}
			D.ctx.Sync();
			if (!D.ctx.ok) { w.rc = SAILOR_PT_ERR_CUDA; w.error = D.ctx.error; }
			w.launches = D.ctx.kernelLaunches; w.h2d += D.ctx.h2dBytes;
		};
		std::vector<std::thread> threads;
		for (int i = 1; i < N; i++) threads.emplace_back(work, i);
		work(0);
		for (std::thread& t : threads) t.join();
		DevSetCurrent(P.device);
		for (const MultiWorker& w : W) if (w.rc != SAILOR_PT_OK) { DevSetCurrent(home); return SetError(w.rc, "device " + std::to_string(w.device) + ": " + w.error); }
		double tOut = 0.0;
		if (flags & 2u)
		{
			P.residentSrgb.Ensure(P.ctx, n * 3);
			P.ctx.TimerStart();
			RunOutputStage(P.ctx, c.width, c.height, P.residentLin.p, P.residentSrgb.p);
			tOut = P.ctx.TimerStop();
		}
		P.ctx.Sync();
		g_stats = SailorPtStats{};
		for (const MultiWorker& w : W)
		{
			g_stats.rays += w.rs.rays; g_stats.primarySamples += w.rs.primarySamples; g_stats.fanOutSamples += w.rs.fanOutSamples; g_stats.replayedRays += w.rs.replayedRays;
			g_stats.traverseLaunches += w.rs.traverseLaunches; g_stats.batches += w.rs.batches; g_stats.kernelLaunches += w.launches; g_stats.h2dBytes += w.h2d;
			// per-stage device time: the busiest device's
			if (w.rs.secondsTraverse > g_stats.secondsTraverse) g_stats.secondsTraverse = w.rs.secondsTraverse;
			if (w.rs.secondsShade > g_stats.secondsShade) g_stats.secondsShade = w.rs.secondsShade;
			if (w.rs.secondsStage[0] > g_stats.secondsExpand) g_stats.secondsExpand = w.rs.secondsStage[0];
			if (w.rs.secondsStage[1] > g_stats.secondsFanOut) g_stats.secondsFanOut = w.rs.secondsStage[1];
			if (w.rs.secondsStage[2] > g_stats.secondsClassify) g_stats.secondsClassify = w.rs.secondsStage[2];
			if (w.rs.secondsStage[3] > g_stats.secondsGather) g_stats.secondsGather = w.rs.secondsStage[3];
			if (w.tBuild > g_stats.secondsBvhBuild) g_stats.secondsBvhBuild = w.tBuild;
		}
		g_stats.secondsOutput = tOut;
		g_stats.devicesUsed = (uint32_t)N;
		g_stats.secondaryTraversal = (P.hasFast && !(p->flags & (SAILOR_PT_FLAG_EXACT_TRAVERSAL | SAILOR_PT_FLAG_WIDE_TRAVERSAL))) ? P.traceChoice : 0u;
		g_stats.secondsCall = g_stats.secondsTotal = HostNow() - t0;      // several devices: no common CUDA clock, host clock around the whole call
		DevSetCurrent(home);
		return FromCtx(P, SAILOR_PT_OK);
	}
}

extern "C" {

int32_t SailorPt_RenderResident(SailorPtScene* s, const SailorPtParams* p, uint32_t flags)
{
	if (!s || !p || !p->height || !p->msaa) return SAILOR_PT_ERR_ARG;
	if (p->deviceCount > 1) { const int rcm = RenderMulti(s, p, flags); if (rcm != kNotMulti) return rcm; }
	SceneDevice& D = s->dev;
	const double t0 = HostNow();
	D.ctx.kernelLaunches = 0; D.ctx.h2dBytes = D.ctx.d2hBytes = 0;
	double tBuild = 0.0;
	D.ctx.Mark(Ctx::kMarkCall0);
	if (flags & 1u) D.built = false;                      // BVH build is part of this pass
	D.wantWide = (p->flags & SAILOR_PT_FLAG_WIDE_TRAVERSAL) != 0u && !(p->flags & SAILOR_PT_FLAG_EXACT_TRAVERSAL);
	if (!D.built)
	{
		const int rcb = FromCtx(D, D.BuildBvh());
		if (rcb != SAILOR_PT_OK) return rcb;
		tBuild = D.stats.secondsBvhBuild;
	}
	const CameraSetup c = CameraOf(D, p);
	const size_t n = (size_t)c.width * c.height;
	if (!n) return SAILOR_PT_ERR_ARG;
	D.residentLin.Ensure(D.ctx, n * 3);
	D.residentW = c.width; D.residentH = c.height;
	if ((p->rowEnd && (p->rowBegin > 0 || p->rowEnd < c.height))) D.residentLin.Zero(D.ctx, n * 3);   // rows outside the shard stay 0
	RenderStats rs{};
	const int rc = RenderFrame(D, ToGpuCamera(c), *p, D.residentLin.p, rs);
	if (rc != SAILOR_PT_OK) return FromCtx(D, rc);
	double tOut = 0.0;
	if (flags & 2u)
	{
		D.residentSrgb.Ensure(D.ctx, n * 3);
		D.ctx.TimerStart();
		RunOutputStage(D.ctx, c.width, c.height, D.residentLin.p, D.residentSrgb.p);
		tOut = D.ctx.TimerStop();
	}
	D.ctx.Mark(Ctx::kMarkCall1);
	D.ctx.Sync();
	g_stats = SailorPtStats{};
	g_stats.secondsCall = D.ctx.Between(Ctx::kMarkCall0, Ctx::kMarkCall1);            // whole call on the launch stream (CUDA events)
	g_stats.rays = rs.rays; g_stats.primarySamples = rs.primarySamples; g_stats.secondsTraverse = rs.secondsTraverse;
	g_stats.secondsShade = rs.secondsShade; g_stats.secondsExpand = rs.secondsStage[0]; g_stats.secondsFanOut = rs.secondsStage[1]; g_stats.secondsClassify = rs.secondsStage[2];
	g_stats.secondsGather = rs.secondsStage[3]; g_stats.fanOutSamples = rs.fanOutSamples; g_stats.replayedRays = rs.replayedRays; g_stats.secondsOutput = tOut; g_stats.secondsBvhBuild = tBuild; g_stats.traverseLaunches = rs.traverseLaunches; g_stats.batches = rs.batches;
	g_stats.kernelLaunches = D.ctx.kernelLaunches; g_stats.h2dBytes = D.ctx.h2dBytes; g_stats.d2hBytes = D.ctx.d2hBytes;
	g_stats.devicesUsed = 1u;
	g_stats.secondaryTraversal = (D.hasFast && !(p->flags & (SAILOR_PT_FLAG_EXACT_TRAVERSAL | SAILOR_PT_FLAG_WIDE_TRAVERSAL))) ? D.traceChoice : 0u;
	g_stats.secondsTotal = HostNow() - t0;
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_ReadResident(SailorPtScene* s, float* linearRGB, uint8_t* srgb8)
{
	if (!s || !s->dev.residentW) return SAILOR_PT_ERR_ARG;
	SceneDevice& D = s->dev;
	const size_t n = (size_t)D.residentW * D.residentH;
	if (linearRGB) D.residentLin.Download(D.ctx, linearRGB, n * 3);
	if (srgb8) { if (D.residentSrgb.n < n * 3) return SAILOR_PT_ERR_ARG; D.residentSrgb.Download(D.ctx, srgb8, n * 3); }
	g_stats.d2hBytes = D.ctx.d2hBytes;
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_CopyResidentToDevice(SailorPtScene* s, void* dstDevice, uint64_t bytes)
{
	if (!s || !dstDevice || !s->dev.residentW) return SAILOR_PT_ERR_ARG;
	SceneDevice& D = s->dev;
	if (bytes != (uint64_t)D.residentW * D.residentH * 3 * sizeof(float)) return SAILOR_PT_ERR_ARG;
	DevCopy(D.ctx, dstDevice, D.residentLin.p, (size_t)bytes);
	D.ctx.Sync();
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_OutputStageResident(SailorPtScene* s, const void* srcDevice, uint64_t bytes)
{
	if (!s || !s->dev.residentW) return SAILOR_PT_ERR_ARG;
	SceneDevice& D = s->dev;
	const size_t n = (size_t)D.residentW * D.residentH;
	if (srcDevice)
	{
		if (bytes != (uint64_t)n * 3 * sizeof(float)) return SAILOR_PT_ERR_ARG;
		DevCopy(D.ctx, D.residentLin.p, srcDevice, (size_t)bytes);
	}
	D.residentSrgb.Ensure(D.ctx, n * 3);
	D.ctx.kernelLaunches = 0;
	D.ctx.TimerStart();
	RunOutputStage(D.ctx, D.residentW, D.residentH, D.residentLin.p, D.residentSrgb.p);
	g_stats.secondsOutput = D.ctx.TimerStop();
	g_stats.kernelLaunches = D.ctx.kernelLaunches;
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_Render(SailorPtScene* s, const SailorPtParams* p, float* linearRGB, uint8_t* srgb8)
{
	if (!s || !p || !linearRGB || !p->height || !p->msaa) return SAILOR_PT_ERR_ARG;
	const double t0 = HostNow();
	int rc = SailorPt_RenderResident(s, p, srgb8 ? 2u : 0u);
	if (rc != SAILOR_PT_OK) return rc;
	rc = SailorPt_ReadResident(s, linearRGB, srgb8);
	g_stats.secondsTotal = HostNow() - t0;
	return rc;
}

int32_t SailorPt_WriteImage(const char* path, uint32_t width, uint32_t height, const float* linearRGB)
{
	if (!path || !path[0] || !linearRGB || !width || !height) return SAILOR_PT_ERR_ARG;
	std::string err;
	int rc;
	switch (ImageFormatOf(path))
	{
	case ImageFormat::Pfm: rc = WritePfm(path, width, height, linearRGB, err); break;
	case ImageFormat::Hdr: rc = WriteHdr(path, width, height, linearRGB, err); break;
	default:
	{
		std::vector<uint8_t> srgb((size_t)width * height * 3);
		rc = SailorPt_OutputStage(width, height, linearRGB, srgb.data());                  // PathTracer.cpp:535-565
		if (rc == SAILOR_PT_OK) rc = EncodePngRgb8(path, width, height, srgb.data(), err);  // stbi_write_png (:560-564)
	}
	}
	if (rc != SAILOR_PT_OK && !err.empty()) t_lastError = err;
	return rc;
}

int32_t SailorPt_CompareImages(uint32_t width, uint32_t height, const float* a, const float* b, double metrics[4])
{
	if (!a || !b || !metrics || !width || !height) return SAILOR_PT_ERR_ARG;
	CompareImages((size_t)width * height * 3, a, b, metrics);
	return SAILOR_PT_OK;
}

int32_t SailorPt_RenderProgressive(SailorPtScene* s, const SailorPtParams* p, uint32_t msaaPerPass, uint32_t maxPasses, const char* checkpointPath,
	uint32_t flags, float* linearRGB, uint8_t* srgb8, uint32_t* msaaDoneOut)
{
	if (!s || !p || !p->height || !p->msaa || !msaaPerPass) return SAILOR_PT_ERR_ARG;
	if (p->msaaEnd) return SetError(SAILOR_PT_ERR_ARG, "a progressive render walks the primary-sample range itself: leave msaaBegin/msaaEnd 0");
	int rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	SceneDevice& D = s->dev;
	const double t0 = HostNow();
	const CameraSetup c = CameraOf(D, p);
	const size_t n = (size_t)c.width * c.height;
	if (!n) return SAILOR_PT_ERR_ARG;
	const uint32_t rowBegin = p->rowEnd ? p->rowBegin : 0u, rowEnd = p->rowEnd ? (p->rowEnd < c.height ? p->rowEnd : c.height) : c.height;
	if (rowBegin >= rowEnd) return SetError(SAILOR_PT_ERR_ARG, "empty shard");
	const size_t bandFloats = (size_t)(rowEnd - rowBegin) * c.width * 3;

	CheckpointHeader hd; memset(&hd, 0, sizeof(hd));
	hd.version = 1; hd.width = c.width; hd.height = c.height; hd.rowBegin = rowBegin; hd.rowEnd = rowEnd; hd.msaaTotal = p->msaa; hd.msaaDone = 0;
	hd.numSamples = p->numSamples; hd.numAmbientSamples = p->numAmbientSamples; hd.maxBounces = p->maxBounces; hd.numTriangles = D.numTris; hd.seed = p->seed;
	memcpy(hd.ambient, p->ambient, sizeof(hd.ambient));
	memcpy(hd.camera, c.pos, 3 * sizeof(float)); memcpy(hd.camera + 3, c.pixel00Dir, 3 * sizeof(float)); memcpy(hd.camera + 6, c.deltaU, 3 * sizeof(float)); memcpy(hd.camera + 9, c.deltaV, 3 * sizeof(float));

	DevBuf<float> running;
	running.Alloc(D.ctx, bandFloats);
	D.residentLin.Ensure(D.ctx, n * 3);
	D.residentW = c.width; D.residentH = c.height;
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	D.residentLin.Zero(D.ctx, n * 3);
	uint32_t done = 0;
	if ((flags & 1u) && checkpointPath && checkpointPath[0])
	{
		CheckpointHeader got; std::vector<float> sum; std::string err;
		const int rr = ReadCheckpoint(checkpointPath, got, sum, err);
		if (rr == SAILOR_PT_OK)
		{
			CheckpointHeader want = hd; want.msaaDone = got.msaaDone;
			if (memcmp(&want, &got, sizeof(want)) != 0) return SetError(SAILOR_PT_ERR_ARG, std::string("checkpoint was written for another scene, camera or parameter set: ") + checkpointPath);
			running.Upload(D.ctx, sum.data(), bandFloats);
			done = got.msaaDone;
		}
		else if (rr != SAILOR_PT_ERR_IO) return SetError(rr, err);          // a missing file means "start from scratch"; a damaged one is an error
	}
	uint64_t rays = 0, samples = 0; uint32_t launches = 0;
	double tTrav = 0.0;
	uint32_t passes = 0;
	SailorPtParams q = *p;
	while (done < p->msaa && (!maxPasses || passes < maxPasses))
	{
		q.msaaBegin = done; q.msaaEnd = done + msaaPerPass < p->msaa ? done + msaaPerPass : p->msaa;
		D.ctx.kernelLaunches = 0;
		RenderStats rs{};
		ProgressiveArgs prog; prog.running = running.p; prog.runningValid = done > 0; prog.norm = q.msaaEnd;
		rc = RenderFrame(D, ToGpuCamera(c), q, D.residentLin.p, rs, prog);
		if (rc != SAILOR_PT_OK) return FromCtx(D, rc);
		rays += rs.rays; samples += rs.primarySamples; launches += D.ctx.kernelLaunches; tTrav += rs.secondsTraverse;
		done = q.msaaEnd; passes++;
		const bool finalPass = done >= p->msaa || (maxPasses && passes >= maxPasses);
		if (checkpointPath && checkpointPath[0] && ((flags & 2u) || finalPass))
		{
			std::vector<float> sum(bandFloats); std::string err;
			running.Download(D.ctx, sum.data(), bandFloats);
			if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
			hd.msaaDone = done;
			rc = WriteCheckpoint(checkpointPath, hd, sum.data(), err);
			if (rc != SAILOR_PT_OK) return SetError(rc, err);
		}
		if ((flags & 4u) && p->output && p->output[0] && !finalPass)         // preview of the estimate so far
		{
			std::vector<float> lin(n * 3);
			D.residentLin.Download(D.ctx, lin.data(), n * 3);
			if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
			rc = SailorPt_WriteImage(p->output, c.width, c.height, lin.data());
			if (rc != SAILOR_PT_OK) return rc;
		}
	}
	if (srgb8)
	{
		D.residentSrgb.Ensure(D.ctx, n * 3);
		RunOutputStage(D.ctx, c.width, c.height, D.residentLin.p, D.residentSrgb.p);
	}
	if (passes == 0 && done > 0)
	{
		// nothing left to render (the checkpoint was complete): rebuild the image from the running sum
		launch_for(D.ctx, (rowEnd - rowBegin) * c.width, ResolveKernel{ running.p, D.residentLin.p, c.width, c.height, rowBegin, rowEnd, 0u, done, running.p, 1u });
		if (srgb8) RunOutputStage(D.ctx, c.width, c.height, D.residentLin.p, D.residentSrgb.p);
	}
	if (linearRGB) D.residentLin.Download(D.ctx, linearRGB, n * 3);
	if (srgb8) D.residentSrgb.Download(D.ctx, srgb8, n * 3);
	if (msaaDoneOut) *msaaDoneOut = done;
	g_stats = SailorPtStats{};
	g_stats.rays = rays; g_stats.primarySamples = samples; g_stats.kernelLaunches = launches; g_stats.secondsTraverse = tTrav; g_stats.secondsTotal = HostNow() - t0;
	if (rc == SAILOR_PT_OK && done >= p->msaa && p->output && p->output[0])
	{
		std::vector<float> own;
		if (!linearRGB) { own.resize(n * 3); D.residentLin.Download(D.ctx, own.data(), n * 3); if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA); }
		rc = SailorPt_WriteImage(p->output, c.width, c.height, linearRGB ? linearRGB : own.data());
	}
	return FromCtx(D, rc);
}

int32_t SailorPt_Run(const SailorPtParams* p)
{
	// PathTracer::Run (PathTracer.cpp:75-575)
	if (!p || !p->pathToModel) return SAILOR_PT_ERR_ARG;
	SailorPtScene* s = nullptr;
	int rc = SailorPt_SceneLoad(p->pathToModel, &s);
	if (rc != SAILOR_PT_OK) return rc;
	uint32_t w = 0, h = 0;
	rc = SailorPt_GetCamera(s, p, &w, &h, nullptr);
	if (rc == SAILOR_PT_OK)
	{
		std::vector<float> lin((size_t)w * h * 3);
		std::vector<uint8_t> srgb((size_t)w * h * 3);
		rc = SailorPt_Render(s, p, lin.data(), srgb.data());
		if (rc == SAILOR_PT_OK && p->output && p->output[0])
		{
			std::string err;
			switch (ImageFormatOf(p->output))
			{
			case ImageFormat::Pfm: rc = WritePfm(p->output, w, h, lin.data(), err); break;     // extension: the linear accumulator, exact bits
			case ImageFormat::Hdr: rc = WriteHdr(p->output, w, h, lin.data(), err); break;
			default: rc = EncodePngRgb8(p->output, w, h, srgb.data(), err);                    // stbi_write_png (PathTracer.cpp:560-564)
			}
			if (rc != SAILOR_PT_OK) t_lastError = err;
		}
	}
	SailorPt_SceneFree(s);
	return rc;
}

int32_t SailorPt_SampleTexture(SailorPtScene* s, uint32_t textureIndex, uint32_t count, const float* uv, float* out)
{
	if (!s || !uv || !out || textureIndex >= s->dev.hostTextures.size()) return SAILOR_PT_ERR_ARG;
	if (!count) return SAILOR_PT_OK;
	SceneDevice& D = s->dev;
	DevBuf<V2> dUv; DevBuf<V4> dOut;
	dUv.Upload(D.ctx, reinterpret_cast<const V2*>(uv), count); dOut.Alloc(D.ctx, count);
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	SampleTextureKernel k; k.ts.texels = D.texels.p; k.ts.textures = D.textures.p; k.ts.srgbLut = D.srgbLut.p; k.ts.texelsF = D.texelsF.p; k.index = textureIndex; k.uv = dUv.p; k.out = dOut.p;
	launch_for(D.ctx, count, k);
	std::vector<V4> host(count);
	dOut.Download(D.ctx, host.data(), count);
	for (uint32_t i = 0; i < count; i++) { out[4 * i] = host[i].x; out[4 * i + 1] = host[i].y; out[4 * i + 2] = host[i].z; out[4 * i + 3] = host[i].w; }
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_DecodeImage(const uint8_t* data, uint64_t size, uint32_t* width, uint32_t* height, uint8_t* rgba8, uint64_t capacity)
{
	if (!data || !size || !width || !height) return SAILOR_PT_ERR_ARG;
	int32_t w = 0, h = 0; std::vector<uint8_t> rgba; std::string err;
	const int rc = DecodeImageRgba8(data, (size_t)size, w, h, rgba, err);
	if (rc != SAILOR_PT_OK) return SetError(rc, err);
	*width = (uint32_t)w; *height = (uint32_t)h;
	if (rgba8)
	{
		if (capacity < rgba.size()) return SetError(SAILOR_PT_ERR_ARG, "image buffer too small");
		memcpy(rgba8, rgba.data(), rgba.size());
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_ShadeHits(SailorPtScene* s, uint32_t count, const uint32_t* triIds, const float* baryUV, const float* rayDirs, uint32_t numSamples, uint32_t numAmbient, float* out)
{
	if (!s || !triIds || !baryUV || !rayDirs || !out) return SAILOR_PT_ERR_ARG;
	if (!count) return SAILOR_PT_OK;
	SceneDevice& D = s->dev;
	DevBuf<uint32_t> dTri; DevBuf<float> dUv, dDir, dOut;
	dTri.Upload(D.ctx, triIds, count); dUv.Upload(D.ctx, baryUV, (size_t)count * 2); dDir.Upload(D.ctx, rayDirs, (size_t)count * 3); dOut.Alloc(D.ctx, (size_t)count * SAILOR_PT_SHADE_FLOATS);
	if (!D.ctx.ok) return FromCtx(D, SAILOR_PT_ERR_CUDA);
	ShadeHitsKernel k;
	k.shade = D.shade.p; k.materials = D.materials.p; k.tex.texels = D.texels.p; k.tex.textures = D.textures.p; k.tex.srgbLut = D.srgbLut.p; k.tex.texelsF = D.texelsF.p;
	k.tri = dTri.p; k.uv = dUv.p; k.dir = dDir.p; k.out = dOut.p; k.numTris = D.numTris; k.numSamples = numSamples; k.numAmbient = numAmbient;
	launch_for(D.ctx, count, k);
	dOut.Download(D.ctx, out, (size_t)count * SAILOR_PT_SHADE_FLOATS);
	return FromCtx(D, SAILOR_PT_OK);
}

int32_t SailorPt_SampleGenerators(uint64_t streamKey, uint32_t kind, uint32_t count, float* out)
{
	if (!out || kind > 2u) return SAILOR_PT_ERR_ARG;
	if (!count) return SAILOR_PT_OK;
	ScopedCtx sc;
	if (sc.rc != SAILOR_PT_OK) return SetError(sc.rc, sc.ctx.error);
	const size_t n = (size_t)count * (kind == 2u ? 4u : 1u);
	DevBuf<float> dOut; DevBuf<uint16_t> dBlue;
	dOut.Alloc(sc.ctx, n); dBlue.Upload(sc.ctx, kBlueNoiseK, (size_t)kBlueNoiseCount);
	if (!sc.ctx.ok) return SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
	launch_for(sc.ctx, 1, SampleGeneratorsKernel{ streamKey, kind, count, dBlue.p, dOut.p });
	dOut.Download(sc.ctx, out, n);
	if (!sc.ctx.ok) return SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
	return SAILOR_PT_OK;
}

int32_t SailorPt_EvalLighting(uint32_t count, const float* in, float* out)
{
	if (!in || !out) return SAILOR_PT_ERR_ARG;
	if (!count) return SAILOR_PT_OK;
	ScopedCtx sc;
	if (sc.rc != SAILOR_PT_OK) return SetError(sc.rc, sc.ctx.error);
	DevBuf<float> dIn, dOut;
	dIn.Upload(sc.ctx, in, (size_t)count * 24); dOut.Alloc(sc.ctx, (size_t)count * 28);
	if (!sc.ctx.ok) return SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
	launch_for(sc.ctx, count, EvalLightingKernel{ dIn.p, dOut.p });
	dOut.Download(sc.ctx, out, (size_t)count * 28);
	if (!sc.ctx.ok) return SetError(SAILOR_PT_ERR_CUDA, sc.ctx.error);
	return SAILOR_PT_OK;
}

} // extern "C"
