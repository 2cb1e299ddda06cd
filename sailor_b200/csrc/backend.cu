// backend.cu — implementation of backend.h (CUDA product build; SPT_EMU host build for tests/emu only).
#include "backend.h"
#include "../../include/sailor_pt.h"
#include <chrono>
#include <stdio.h>
#include <array>
#include <mutex>
#include <unordered_map>
#include <thread>
#include <condition_variable>
#include <atomic>
#include <vector>

namespace spt
{
	bool Ctx::Fail(const char* what, int code)
	{
		if (ok)
		{
			ok = false;
			char buf[512];
#if !defined(SPT_EMU)
			snprintf(buf, sizeof(buf), "%s -> %s (%d)", what, cudaGetErrorString((cudaError_t)code), code);
#else
			snprintf(buf, sizeof(buf), "%s -> %d", what, code);
#endif
			error = buf;
		}
		return false;
	}

#if !defined(SPT_EMU)
	// ------------------------------------------------------------------------------------------------ CUDA
	namespace
	{
		std::mutex g_pinnedPoolMutex;
		std::vector<uint32_t*> g_pinnedPool;
	}
	int Ctx::Init()
	{
		int count = 0;
		if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
		{
			cudaGetLastError();
			ok = false;
			error = "no CUDA device: the sailor_b200 product library has no CPU path";
			return SAILOR_PT_ERR_NO_DEVICE;
		}
		if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess ||
			cudaEventCreate(&evA) != cudaSuccess || cudaEventCreate(&evB) != cudaSuccess)
		{
			Fail("cudaStreamCreate/cudaEventCreate", (int)cudaGetLastError());
			return SAILOR_PT_ERR_CUDA;
		}
		// marker events are created on first use (Mark): a scene object per frame, like the reference's PathTracer object per Run,
		// should not pay for 64 events it never records
		ok = true;
		return SAILOR_PT_OK;
	}
	void Ctx::Destroy()
	{
		if (pinned)
		{
			std::lock_guard<std::mutex> lock(g_pinnedPoolMutex);
			if (g_pinnedPool.size() < 64) g_pinnedPool.push_back(pinned); else cudaFreeHost(pinned);
			pinned = nullptr;
		}
		if (evA) cudaEventDestroy(evA);
		if (evB) cudaEventDestroy(evB);
		for (int i = 0; i < kMarkers; i++) { if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; }
		if (stream) cudaStreamDestroy(stream);
		evA = evB = nullptr; stream = nullptr;
	}
	void Ctx::Sync() { if (ok) SPT_CUDA_CHECK(*this, cudaStreamSynchronize(stream)); }
	void Ctx::TimerStart() { if (ok) SPT_CUDA_CHECK(*this, cudaEventRecord(evA, stream)); }
	double Ctx::TimerStop()
	{
		if (!ok) return 0.0;
		SPT_CUDA_CHECK(*this, cudaEventRecord(evB, stream));
		SPT_CUDA_CHECK(*this, cudaEventSynchronize(evB));
		float ms = 0.0f;
		if (ok) SPT_CUDA_CHECK(*this, cudaEventElapsedTime(&ms, evA, evB));
		return (double)ms * 1e-3;
	}

	void Ctx::Mark(int i)
	{
		if (!ok) return;
		if (!ev[i]) SPT_CUDA_CHECK(*this, cudaEventCreate(&ev[i]));
		if (ok) SPT_CUDA_CHECK(*this, cudaEventRecord(ev[i], stream));
	}
	void Ctx::WaitMark(int i) { if (ok) SPT_CUDA_CHECK(*this, cudaEventSynchronize(ev[i])); }
	// The pinned scratch block is recycled through a process-wide free list: cudaHostAlloc / cudaFreeHost cost about a millisecond
	// each (page locking + an implicit device synchronisation), which a scene object per frame would pay every frame.
	uint32_t* Ctx::Pinned()
	{
		if (!pinned && ok)
		{
			{
				std::lock_guard<std::mutex> lock(g_pinnedPoolMutex);
				if (!g_pinnedPool.empty()) { pinned = g_pinnedPool.back(); g_pinnedPool.pop_back(); }
			}
			if (!pinned) { void* p = nullptr; SPT_CUDA_CHECK(*this, cudaHostAlloc(&p, 256 * sizeof(uint32_t), cudaHostAllocPortable)); pinned = (uint32_t*)p; }
		}
		return pinned;
	}
	void Ctx::ReadAsync(uint32_t* pinnedDst, const void* src, size_t bytes)
	{
		d2hBytes += bytes;
		if (ok) SPT_CUDA_CHECK(*this, cudaMemcpyAsync(pinnedDst, src, bytes, cudaMemcpyDeviceToHost, stream));
	}
	double Ctx::Between(int i, int j)
	{
		float ms = 0.0f;
		if (ok) SPT_CUDA_CHECK(*this, cudaEventElapsedTime(&ms, ev[i], ev[j]));
		return (double)ms * 1e-3;
	}

	// Device memory comes from CUDA's stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so a
	// host that renders frame after frame (one scene object per call, like the reference's PathTracer object per Run)
	// reuses the same HBM instead of paying cudaMalloc/cudaFree of multi-GB arenas every call.  A buffer is freed on
	// the stream it was allocated on (kept in a small registry), which orders the free after the work that used it.
	namespace
	{
		std::mutex g_allocMutex;
		std::unordered_map<void*, cudaStream_t> g_allocStream;
		uint64_t g_poolReady = 0;            // bit d: the release threshold of device d's default pool has been lifted

		void EnsurePool()
		{
			int dev = 0; cudaMemPool_t pool = nullptr;
			if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
			if (dev < 64 && (g_poolReady >> dev & 1u)) return;
			if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
			{
				uint64_t threshold = UINT64_MAX;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
			}
			cudaGetLastError();
			if (dev < 64) g_poolReady |= 1ull << dev;
		}
	}

	void* DevAllocBytes(Ctx& ctx, size_t bytes)
	{
		void* p = nullptr;
		if (!ctx.ok) return nullptr;
		std::lock_guard<std::mutex> lock(g_allocMutex);
		EnsurePool();
		SPT_CUDA_CHECK(ctx, cudaMallocAsync(&p, bytes, ctx.stream));
		if (!ctx.ok) return nullptr;
		g_allocStream[p] = ctx.stream;
		return p;
	}
	int DevCurrent() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d; }
	int DevCount() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; } return n; }
	bool DevSetCurrent(int device) { if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return false; } return true; }
	void DevCopyPeer(Ctx& ctx, void* dst, int dstDevice, const void* src, int srcDevice, size_t bytes)
	{
		if (!ctx.ok || !bytes) return;
		if (dstDevice == srcDevice) SPT_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx.stream));
		else SPT_CUDA_CHECK(ctx, cudaMemcpyPeerAsync(dst, dstDevice, src, srcDevice, bytes, ctx.stream));
	}
	void TrimDevicePool()
	{
		int dev = 0; cudaMemPool_t pool = nullptr;
		if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
		cudaGetLastError();
	}
	void* DevAllocPlain(Ctx& ctx, size_t bytes)
	{
		void* p = nullptr;
		if (ctx.ok) SPT_CUDA_CHECK(ctx, cudaMalloc(&p, bytes));
		return ctx.ok ? p : nullptr;
	}
	void DevFreePlain(void* p) { if (p && cudaFree(p) != cudaSuccess) cudaGetLastError(); }
	size_t DevMemAvailable()
	{
		size_t freeB = 0, totalB = 0;
		if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) { cudaGetLastError(); return 0; }
		// memory this process freed earlier sits in the stream-ordered pool (release threshold lifted) and is reusable
		int dev = 0; cudaMemPool_t pool = nullptr; uint64_t reserved = 0, used = 0;
		if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
			cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
			cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
			freeB += (size_t)(reserved - used);
		else cudaGetLastError();
		return freeB;
	}
	void DevFreeBytes(void* p)
	{
		if (!p) return;
		cudaStream_t st = nullptr; bool known = false;
		{
			std::lock_guard<std::mutex> lock(g_allocMutex);
			auto it = g_allocStream.find(p);
			if (it != g_allocStream.end()) { st = it->second; known = true; g_allocStream.erase(it); }
		}
		if (known) { if (cudaFreeAsync(p, st) != cudaSuccess) { cudaGetLastError(); cudaFree(p); } }
		else cudaFree(p);
	}
	void DevUpload(Ctx& ctx, void* dst, const void* src, size_t bytes) { ctx.h2dBytes += bytes; if (ctx.ok) SPT_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx.stream)); }
	// Large device->host reads go through two pinned staging chunks: the DMA of chunk k overlaps the host copy of chunk
	// k-1 into the caller's (pageable) buffer, and that host copy is split over a few persistent helper threads (one
	// core moves ~10 GB/s, the PCIe link ~50 GB/s).  Small reads use the plain path.
	namespace
	{
		constexpr size_t kStageChunk = 8u << 20;
		std::mutex g_stageMutex;
		unsigned char* g_stage[2] = { nullptr, nullptr };
		std::unordered_map<int, std::array<cudaEvent_t, 2>> g_stageEvByDevice;     // an event belongs to the device it was created on
		cudaEvent_t* g_stageEv = nullptr;          // the current device's pair (valid under g_stageMutex after EnsureStage)

		bool EnsureStage()
		{
			if (!g_stage[0])
				for (int k = 0; k < 2; k++)
					if (cudaHostAlloc((void**)&g_stage[k], kStageChunk, cudaHostAllocPortable) != cudaSuccess)
					{
						cudaGetLastError();
						for (int j = 0; j < 2; j++) { if (g_stage[j]) cudaFreeHost(g_stage[j]); g_stage[j] = nullptr; }
						return false;
					}
			const int dev = DevCurrent();
			auto it = g_stageEvByDevice.find(dev);
			if (it == g_stageEvByDevice.end())
			{
				std::array<cudaEvent_t, 2> ev{ nullptr, nullptr };
				for (int k = 0; k < 2; k++) if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
				it = g_stageEvByDevice.emplace(dev, ev).first;
			}
			g_stageEv = it->second.data();
			return true;
		}

		// persistent helpers for the host side of staged copies (created on first use, detached, idle on a condvar)
		struct CopyPool
		{
			static constexpr int kWorkers = 3;
			std::mutex m; std::condition_variable cvWork, cvDone;
			unsigned char* dst = nullptr; const unsigned char* src = nullptr; size_t bytes = 0;
			uint64_t generation = 0; int pending = 0; bool started = false;

			void Worker(int id)
			{
				uint64_t seen = 0;
				for (;;)
				{
					unsigned char* d; const unsigned char* s; size_t n;
					{
						std::unique_lock<std::mutex> lk(m);
						cvWork.wait(lk, [&] { return generation != seen; });
						seen = generation; d = dst; s = src; n = bytes;
					}
					const size_t part = (n / (kWorkers + 1) + 63) & ~size_t(63);
					const size_t b = (size_t)(id + 1) * part, e = b + part < n ? b + part : n;
					if (b < n) memcpy(d + b, s + b, (id == kWorkers - 1 ? n : e) - b);
					{
						std::lock_guard<std::mutex> lk(m);
						if (--pending == 0) cvDone.notify_one();
					}
				}
			}
			void Copy(void* d, const void* s, size_t n)
			{
				if (n < (1u << 20)) { memcpy(d, s, n); return; }
				{
					std::lock_guard<std::mutex> lk(m);
					if (!started) { for (int i = 0; i < kWorkers; i++) std::thread(&CopyPool::Worker, this, i).detach(); started = true; }
					dst = (unsigned char*)d; src = (const unsigned char*)s; bytes = n; pending = kWorkers; generation++;
				}
				cvWork.notify_all();
				const size_t part = (n / (kWorkers + 1) + 63) & ~size_t(63);
				memcpy(d, s, part < n ? part : n);                        // the caller copies the first share
				std::unique_lock<std::mutex> lk(m);
				cvDone.wait(lk, [&] { return pending == 0; });
			}
		};
		CopyPool* g_copyPool = new CopyPool();      // leaked on purpose: helper threads may outlive static destruction
	}

	namespace
	{
		std::mutex g_pinMutex;
		std::unordered_map<void*, size_t> g_pinned;
		bool IsPinned(const void* p, size_t bytes)
		{
			std::lock_guard<std::mutex> lock(g_pinMutex);
			for (const auto& kv : g_pinned)
			{
				const unsigned char* b = (const unsigned char*)kv.first;
				if ((const unsigned char*)p >= b && (const unsigned char*)p + bytes <= b + kv.second) return true;
			}
			return false;
		}
	}
	int HostPin(void* p, size_t bytes)
	{
		if (!p || !bytes) return -1;
		std::lock_guard<std::mutex> lock(g_pinMutex);
		if (g_pinned.count(p)) return g_pinned[p] >= bytes ? 0 : -1;
		if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
		g_pinned[p] = bytes;
		return 0;
	}
	int HostUnpin(void* p)
	{
		std::lock_guard<std::mutex> lock(g_pinMutex);
		auto it = g_pinned.find(p);
		if (it == g_pinned.end()) return -1;
		g_pinned.erase(it);
		if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return -1; }
		return 0;
	}

	void DevDownload(Ctx& ctx, void* dst, const void* src, size_t bytes)
	{
		if (!ctx.ok) return;
		ctx.d2hBytes += bytes;
		if (bytes >= (1u << 20) && !IsPinned(dst, bytes))
		{
			std::lock_guard<std::mutex> lock(g_stageMutex);
			if (EnsureStage())
			{
				size_t prevOff = 0, prevN = 0; int prevK = -1;
				int k = 0;
				for (size_t off = 0; off < bytes && ctx.ok; off += kStageChunk, k ^= 1)
				{
					const size_t n = bytes - off < kStageChunk ? bytes - off : kStageChunk;
					SPT_CUDA_CHECK(ctx, cudaMemcpyAsync(g_stage[k], (const unsigned char*)src + off, n, cudaMemcpyDeviceToHost, ctx.stream));
					SPT_CUDA_CHECK(ctx, cudaEventRecord(g_stageEv[k], ctx.stream));
					if (prevK >= 0)
					{
						SPT_CUDA_CHECK(ctx, cudaEventSynchronize(g_stageEv[prevK]));
						g_copyPool->Copy((unsigned char*)dst + prevOff, g_stage[prevK], prevN);
					}
					prevOff = off; prevN = n; prevK = k;
				}
				if (prevK >= 0 && ctx.ok)
				{
					SPT_CUDA_CHECK(ctx, cudaEventSynchronize(g_stageEv[prevK]));
					g_copyPool->Copy((unsigned char*)dst + prevOff, g_stage[prevK], prevN);
				}
				SPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx.stream));
				return;
			}
		}
		SPT_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx.stream));
		SPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx.stream));
	}
	void DevMemset(Ctx& ctx, void* dst, int byte, size_t bytes) { if (ctx.ok) SPT_CUDA_CHECK(ctx, cudaMemsetAsync(dst, byte, bytes, ctx.stream)); }
	void DevCopy(Ctx& ctx, void* dst, const void* src, size_t bytes) { if (ctx.ok) SPT_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx.stream)); }

	int RangeGridBlocks()
	{
		// one value per process: the devices of one box are identical (initialised once, thread-safe)
		static const int blocks = [] { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); return sms * 8; }();
		return blocks;
	}

	// Timing events are recycled through a process-wide free list: a host that creates one scene object per frame (the reference's
	// PathTracer object per Run) would otherwise create and destroy ~40 events per frame.
	namespace
	{
		std::mutex g_eventPoolMutex;
		std::unordered_map<int, std::vector<cudaEvent_t>> g_eventPool;     // per device: an event belongs to the device it was created on
		int CurrentDevice() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d; }
	}
	void SpanTimer::Begin(Ctx& ctx)
	{
		if (!ctx.ok) return;
		while (ev.size() < used + 2)
		{
			cudaEvent_t e = nullptr;
			{
				std::lock_guard<std::mutex> lock(g_eventPoolMutex);
				if (device < 0) device = CurrentDevice();
				std::vector<cudaEvent_t>& pool = g_eventPool[device];
				if (!pool.empty()) { e = pool.back(); pool.pop_back(); }
			}
			if (!e && cudaEventCreate(&e) != cudaSuccess) { ctx.Fail("cudaEventCreate", (int)cudaGetLastError()); return; }
			ev.push_back(e);
		}
		SPT_CUDA_CHECK(ctx, cudaEventRecord(ev[used], ctx.stream));
	}
	void SpanTimer::End(Ctx& ctx)
	{
		if (!ctx.ok || ev.size() < used + 2) return;
		SPT_CUDA_CHECK(ctx, cudaEventRecord(ev[used + 1], ctx.stream));
		used += 2;
	}
	double SpanTimer::Collect(Ctx& ctx)
	{
		double s = 0.0;
		for (size_t i = 0; i + 1 < used && ctx.ok; i += 2) { float ms = 0.0f; SPT_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms, ev[i], ev[i + 1])); s += (double)ms * 1e-3; }
		used = 0;
		return s;
	}
	void SpanTimer::Destroy()
	{
		std::lock_guard<std::mutex> lock(g_eventPoolMutex);
		std::vector<cudaEvent_t>& pool = g_eventPool[device < 0 ? CurrentDevice() : device];
		for (cudaEvent_t e : ev) { if (pool.size() < 4096) pool.push_back(e); else cudaEventDestroy(e); }
		ev.clear(); used = 0;
	}

	// ---- exclusive scan: reduce-then-scan, 2048 items per CTA, coalesced, warp shuffles -----------------
	namespace
	{
		constexpr uint32_t kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

		__device__ __forceinline__ uint32_t BlockExclusiveScan(uint32_t v, uint32_t& total)
		{
			__shared__ uint32_t warpSums[kScanThreads / 32];
			const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			uint32_t incl = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
			if (lane == 31) warpSums[warp] = incl;
			__syncthreads();
			if (warp == 0)
			{
				uint32_t w = lane < kScanThreads / 32 ? warpSums[lane] : 0;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
				if (lane < kScanThreads / 32) warpSums[lane] = w;
			}
			__syncthreads();
			total = warpSums[kScanThreads / 32 - 1];
			const uint32_t base = warp ? warpSums[warp - 1] : 0;
			__syncthreads();
			return base + incl - v;
		}

		__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tileSums)
		{
			const uint32_t base = blockIdx.x * kScanTile;
			uint32_t s = 0;
#pragma unroll
			for (uint32_t k = 0; k < kScanItems; k++) { const uint32_t i = base + k * kScanThreads + threadIdx.x; if (i < n) s += in[i]; }
			uint32_t total;
			BlockExclusiveScan(s, total);
			if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
		}

		__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(uint32_t* tileSums, uint32_t numTiles)
		{
			uint32_t carry = 0;
			for (uint32_t base = 0; base < numTiles; base += kScanThreads)
			{
				const uint32_t i = base + threadIdx.x;
				const uint32_t v = i < numTiles ? tileSums[i] : 0;
				uint32_t total;
				const uint32_t ex = BlockExclusiveScan(v, total);
				if (i < numTiles) tileSums[i] = carry + ex;
				carry += total;
			}
			if (threadIdx.x == 0) tileSums[numTiles] = carry;
		}

		__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tileOffsets, uint32_t numTiles)
		{
			// each thread owns kScanItems CONSECUTIVE items so the scan is a plain running sum per thread
			const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
			uint32_t v[kScanItems];
			uint32_t s = 0;
#pragma unroll
			for (uint32_t k = 0; k < kScanItems; k++) { v[k] = (base + k) < n ? in[base + k] : 0; s += v[k]; }
			uint32_t total;
			uint32_t run = tileOffsets[blockIdx.x] + BlockExclusiveScan(s, total);
#pragma unroll
			for (uint32_t k = 0; k < kScanItems; k++) { if ((base + k) < n) out[base + k] = run; run += v[k]; }
			if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tileOffsets[numTiles];
		}
	}

	void ExclusiveScanU32(Ctx& ctx, const uint32_t* in, uint32_t* out, uint32_t n, DevBuf<uint32_t>& scratch)
	{
		if (!ctx.ok) return;
		if (n == 0) { DevMemset(ctx, out, 0, sizeof(uint32_t)); return; }
		const uint32_t numTiles = (n + kScanTile - 1) / kScanTile;
		scratch.Ensure(ctx, numTiles + 1);
		if (!ctx.ok) return;
		k_scan_reduce<<<numTiles, kScanThreads, 0, ctx.stream>>>(in, n, scratch.p);
		k_scan_tiles<<<1, kScanThreads, 0, ctx.stream>>>(scratch.p, numTiles);
		k_scan_apply<<<numTiles, kScanThreads, 0, ctx.stream>>>(in, out, n, scratch.p, numTiles);
		ctx.kernelLaunches += 3;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}

#else
	// ------------------------------------------------------------------------------------------------ EMU (tests only)
	static std::chrono::steady_clock::time_point g_t0, g_marks[64];
	static double NowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	void SpanTimer::Begin(Ctx&) { t0 = NowS(); }
	void SpanTimer::End(Ctx&) { acc += NowS() - t0; used += 2; }
	double SpanTimer::Collect(Ctx&) { const double s = acc; acc = 0.0; used = 0; return s; }
	void SpanTimer::Destroy() {}
	void Ctx::Mark(int i) { g_marks[i] = std::chrono::steady_clock::now(); }
	void Ctx::WaitMark(int) {}
	uint32_t* Ctx::Pinned() { if (!pinned) pinned = (uint32_t*)calloc(256, sizeof(uint32_t)); return pinned; }
	void Ctx::ReadAsync(uint32_t* pinnedDst, const void* src, size_t bytes) { d2hBytes += bytes; memcpy(pinnedDst, src, bytes); }
	double Ctx::Between(int i, int j) { return std::chrono::duration<double>(g_marks[j] - g_marks[i]).count(); }
	int Ctx::Init() { ok = true; return SAILOR_PT_OK; }
	void Ctx::Destroy() { free(pinned); pinned = nullptr; }
	void Ctx::Sync() {}
	void Ctx::TimerStart() { g_t0 = std::chrono::steady_clock::now(); }
	double Ctx::TimerStop() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count(); }
	void* DevAllocBytes(Ctx&, size_t bytes) { return malloc(bytes ? bytes : 1); }
	void DevFreeBytes(void* p) { free(p); }
	size_t DevMemAvailable() { return (size_t)4 << 30; }
	int DevCurrent() { return 0; }
	int DevCount() { return 1; }
	bool DevSetCurrent(int device) { return device == 0; }
	void DevCopyPeer(Ctx&, void* dst, int, const void* src, int, size_t bytes) { memcpy(dst, src, bytes); }
	void TrimDevicePool() {}
	void* DevAllocPlain(Ctx&, size_t bytes) { return malloc(bytes ? bytes : 1); }
	void DevFreePlain(void* p) { free(p); }
	void DevUpload(Ctx& ctx, void* dst, const void* src, size_t bytes) { ctx.h2dBytes += bytes; memcpy(dst, src, bytes); }
	void DevDownload(Ctx& ctx, void* dst, const void* src, size_t bytes) { ctx.d2hBytes += bytes; memcpy(dst, src, bytes); }
	int HostPin(void* p, size_t bytes) { return (p && bytes) ? 0 : -1; }
	int HostUnpin(void* p) { return p ? 0 : -1; }
	void DevMemset(Ctx&, void* dst, int byte, size_t bytes) { memset(dst, byte, bytes); }
	void DevCopy(Ctx&, void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
	void ExclusiveScanU32(Ctx& ctx, const uint32_t* in, uint32_t* out, uint32_t n, DevBuf<uint32_t>&)
	{
		uint32_t s = 0;
		for (uint32_t i = 0; i < n; i++) { out[i] = s; s += in[i]; }
		out[n] = s;
		ctx.kernelLaunches += 3;
	}
#endif
}
