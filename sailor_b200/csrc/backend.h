// backend.h — device memory, stream, launch and atomic primitives used by the pipeline.
//
// PRODUCT BUILD (nvcc, sm_100a): CUDA stream + cudaMalloc + real kernels.  There is no CPU execution path in the
// product library: Ctx::Init() fails with SAILOR_PT_ERR_NO_DEVICE when no usable CUDA device exists.
//
// SPT_EMU (g++, tests/emu only): the same kernel bodies are compiled for the host and each "launch" is a serial
// loop.  This exists because the development container has no GPU; it lets tests/ check the host orchestration and
// the kernel logic against the oracle on CPU.  It is built only by tests/emu/build_emu.py into tests/emu/_build/,
// is never loaded by the sailor_b200 package, and is not a fallback.
#pragma once
#include "hd.h"
#include <stddef.h>
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>

#if !defined(SPT_EMU)
#include <cuda_runtime.h>
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include <cooperative_groups/reduce.h>
#endif
#endif

namespace spt
{
	struct Ctx
	{
#if !defined(SPT_EMU)
		cudaStream_t stream = nullptr;
		cudaEvent_t evA = nullptr, evB = nullptr;
		static constexpr int kMarkers = 64;
		cudaEvent_t ev[kMarkers] = {};      // stage markers: 0..59 wavefront iterations (3 per iteration), 60-61 whole call
#endif
		uint32_t kernelLaunches = 0;
		uint64_t h2dBytes = 0, d2hBytes = 0;
		static constexpr int kMarkCall0 = 60, kMarkCall1 = 61;
		std::string error;
		bool ok = true;

		int Init();       // SAILOR_PT_OK or SAILOR_PT_ERR_NO_DEVICE
		void Destroy();
		void Sync();
		bool Fail(const char* what, int code);
		// GPU time of a bracketed region on the launch stream (seconds); emu: host clock
		void TimerStart();
		double TimerStop();
		void Mark(int i);                  // record marker i on the launch stream
		double Between(int i, int j);      // seconds between two recorded markers (after a sync)
		void WaitMark(int i);              // block the host until marker i has been reached
		// small read-backs that must not drain the queue: copy into the context's pinned scratch (256 words) behind the work
		// queued so far; the data is valid after WaitMark() of a marker recorded after the copy
		uint32_t* pinned = nullptr;
		uint32_t* Pinned();
		void ReadAsync(uint32_t* pinnedDst, const void* src, size_t bytes);
	};

#if !defined(SPT_EMU)
#define SPT_CUDA_CHECK(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (ctx).Fail(#call, (int)e_); } } while (0)
#endif

	void* DevAllocBytes(Ctx& ctx, size_t bytes);
	int DevCurrent();                  // index of the current CUDA device (emu: 0)
	int DevCount();                    // CUDA devices visible to the process (emu: 1)
	bool DevSetCurrent(int device);    // make `device` current for the calling host thread
	// copy between two devices on ctx.stream (falls back to a staged copy inside the driver when the pair has no peer access)
	void DevCopyPeer(Ctx& ctx, void* dst, int dstDevice, const void* src, int srcDevice, size_t bytes);
	void TrimDevicePool();             // hand the stream-ordered pool's cached memory back to the driver (emu: no-op)
	// allocations that outlive the context (stream) they were made from: plain cudaMalloc / cudaFree, not stream-ordered
	void* DevAllocPlain(Ctx& ctx, size_t bytes);
	void DevFreePlain(void* p);
	size_t DevMemAvailable();          // bytes a new allocation can still get (free device memory + what the pool holds back); emu: 4 GiB
	void DevFreeBytes(void* p);
	void DevUpload(Ctx& ctx, void* dst, const void* src, size_t bytes);
	void DevDownload(Ctx& ctx, void* dst, const void* src, size_t bytes);   // synchronises
	void DevMemset(Ctx& ctx, void* dst, int byte, size_t bytes);
	// caller-owned host buffers page-locked through the C-ABI: DevDownload copies into them without staging
	int HostPin(void* p, size_t bytes);       // 0 ok, -1 failed
	int HostUnpin(void* p);
	void DevCopy(Ctx& ctx, void* dst, const void* src, size_t bytes);

	template<class T>
	struct DevBuf
	{
		T* p = nullptr;
		size_t n = 0;
		DevBuf() = default;
		DevBuf(const DevBuf&) = delete;
		DevBuf& operator=(const DevBuf&) = delete;
		~DevBuf() { Free(); }
		void Free() { if (p) DevFreeBytes(p); p = nullptr; n = 0; }
		void Alloc(Ctx& ctx, size_t count) { Free(); if (count) { p = (T*)DevAllocBytes(ctx, count * sizeof(T)); n = p ? count : 0; } }
		void Ensure(Ctx& ctx, size_t count) { if (count > n) Alloc(ctx, count); }
		void Upload(Ctx& ctx, const T* src, size_t count) { Ensure(ctx, count); if (count) DevUpload(ctx, p, src, count * sizeof(T)); }
		void Upload(Ctx& ctx, const std::vector<T>& v) { Upload(ctx, v.data(), v.size()); }
		void Download(Ctx& ctx, T* dst, size_t count) const { if (count) DevDownload(ctx, dst, p, count * sizeof(T)); }
		void Zero(Ctx& ctx) { if (n) DevMemset(ctx, p, 0, n * sizeof(T)); }
		void Zero(Ctx& ctx, size_t count) { if (count) DevMemset(ctx, p, 0, count * sizeof(T)); }
	};

	// Buffer shared by every context of the process on one device (the wavefront arenas): not tied to any stream.
	struct PlainBuf
	{
		unsigned char* p = nullptr; size_t n = 0;
		void Ensure(Ctx& ctx, size_t bytes)
		{
			if (bytes <= n) return;
			if (p) { ctx.Sync(); DevFreePlain(p); p = nullptr; n = 0; }
			p = (unsigned char*)DevAllocPlain(ctx, bytes); n = p ? bytes : 0;
		}
	};

	// ---- atomics usable from kernel bodies -------------------------------------------------------------
#if !defined(SPT_EMU)
	__device__ __forceinline__ void atomic_min_u32(uint32_t* p, uint32_t v) { atomicMin(p, v); }
	__device__ __forceinline__ void atomic_max_u32(uint32_t* p, uint32_t v) { atomicMax(p, v); }
	__device__ __forceinline__ uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
	__device__ __forceinline__ unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
#define SPT_KERNEL_BODY __device__ __forceinline__
#if defined(__CUDACC__)
	// Warp-aggregated forms: the lanes that are active at the call each reserve v (or 1) consecutive units of *p, which must
	// be the SAME address in every lane; one atomic per warp instead of one per lane, ranges handed out in lane order.
	// The wavefront counters (ray queue length, record arena, fan-out tables) are bumped once per activation: per-lane
	// atomics on four addresses were what bounded ExpandKernel (profiles/r01g_SUMMARY.md).
	__device__ __forceinline__ uint32_t atomic_inc_u32_agg(uint32_t* p)
	{
		const uint32_t mask = __activemask(), lane = threadIdx.x & 31u;
		const int leader = __ffs(mask) - 1;
		uint32_t base = 0;
		if ((int)lane == leader) base = atomicAdd(p, (uint32_t)__popc(mask));
		base = __shfl_sync(mask, base, leader);
		return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
	}
	__device__ __forceinline__ uint32_t atomic_add_u32_agg(uint32_t* p, uint32_t v)
	{
		namespace cg = cooperative_groups;
		const cg::coalesced_group g = cg::coalesced_threads();
		const uint32_t excl = cg::exclusive_scan(g, v);
		const uint32_t last = g.size() - 1u;
		uint32_t base = 0;
		if (g.thread_rank() == last) base = atomicAdd(p, excl + v);
		base = g.shfl(base, last);
		return base + excl;
	}
	__device__ __forceinline__ void atomic_add_u64_agg(unsigned long long* p, unsigned long long v)
	{
		namespace cg = cooperative_groups;
		const cg::coalesced_group g = cg::coalesced_threads();
		const unsigned long long total = cg::reduce(g, v, cg::plus<unsigned long long>());
		if (g.thread_rank() == 0) atomicAdd(p, total);
	}
#endif
#else
	inline void atomic_min_u32(uint32_t* p, uint32_t v) { if (v < *p) *p = v; }
	inline void atomic_max_u32(uint32_t* p, uint32_t v) { if (v > *p) *p = v; }
	inline uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
	inline unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
	inline uint32_t atomic_inc_u32_agg(uint32_t* p) { return atomic_add_u32(p, 1u); }
	inline uint32_t atomic_add_u32_agg(uint32_t* p, uint32_t v) { return atomic_add_u32(p, v); }
	inline void atomic_add_u64_agg(unsigned long long* p, unsigned long long v) { *p += v; }
#define SPT_KERNEL_BODY inline
#endif

	// ---- launch_for: functor f(i) for i in [0,n) ---------------------------------------------------------
#if !defined(SPT_EMU)
	template<class F>
	__global__ void __launch_bounds__(256) k_for(uint32_t n, F f)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < n) f(i);
	}

	template<class F>
	inline void launch_for(Ctx& ctx, uint32_t n, const F& f)
	{
		if (!n || !ctx.ok) return;
		k_for<F><<<(n + 255u) / 256u, 256, 0, ctx.stream>>>(n, f);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	template<class F>
	inline void launch_for(Ctx& ctx, uint32_t n, const F& f)
	{
		if (!ctx.ok) return;
		for (uint32_t i = 0; i < n; i++) f(i);
		ctx.kernelLaunches++;
	}
#endif

	// ---- launch_for_range: functor f(i) for i in [*begin, min(*end, cap)), bounds read from DEVICE memory ----
	// The wavefront levels are sized by device-side counters; reading them here would cost a host round trip per level,
	// so the grid is fixed (a few CTAs per SM, grid-stride) and the kernel fetches its own range.
#if !defined(SPT_EMU)
	// MinBlocks: occupancy hint (caps registers per thread) for the register-hungry shading functors
	template<class F, int MinBlocks>
	__global__ void __launch_bounds__(256, MinBlocks) k_for_range(const uint32_t* __restrict__ begin, const uint32_t* __restrict__ end, uint32_t cap, F f)
	{
		const uint32_t b = *begin;
		uint32_t e = *end; if (e > cap) e = cap;
		for (uint32_t i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) f(i);
	}

	int RangeGridBlocks();    // SMs x 8

	template<int MinBlocks = 1, class F>
	inline void launch_for_range(Ctx& ctx, const uint32_t* dBegin, const uint32_t* dEnd, uint32_t cap, uint32_t maxCount, const F& f)
	{
		if (!maxCount || !ctx.ok) return;
		uint32_t blocks = (maxCount + 255u) / 256u;
		const uint32_t lim = (uint32_t)RangeGridBlocks();
		if (blocks > lim) blocks = lim;
		k_for_range<F, MinBlocks><<<blocks, 256, 0, ctx.stream>>>(dBegin, dEnd, cap, f);
		ctx.kernelLaunches++;
		SPT_CUDA_CHECK(ctx, cudaGetLastError());
	}
#else
	template<int MinBlocks = 1, class F>
	inline void launch_for_range(Ctx& ctx, const uint32_t* dBegin, const uint32_t* dEnd, uint32_t cap, uint32_t, const F& f)
	{
		if (!ctx.ok) return;
		uint32_t e = *dEnd; if (e > cap) e = cap;
		for (uint32_t i = *dBegin; i < e; i++) f(i);
		ctx.kernelLaunches++;
	}
#endif

	// Sum of the device time of bracketed spans on the launch stream (CUDA events, read after a sync); emu: host clock.
	struct SpanTimer
	{
#if !defined(SPT_EMU)
		std::vector<cudaEvent_t> ev;
#endif
		size_t used = 0; double acc = 0.0; double t0 = 0.0;
		int device = -1;            // device the events were created on (they go back to that device's free list)
		void Begin(Ctx& ctx);
		void End(Ctx& ctx);
		double Collect(Ctx& ctx);     // after ctx.Sync(): seconds of all spans since the last Collect
		uint32_t Spans() const { return (uint32_t)(used / 2); }
		void Destroy();
	};

	// out[i] = sum(in[0..i)), out has n+1 entries (out[n] = total).  in/out may not alias.
	void ExclusiveScanU32(Ctx& ctx, const uint32_t* in, uint32_t* out, uint32_t n, DevBuf<uint32_t>& scratch);
}
