// trace_kernels.cuh — launchers of the traversal stage.
//
// Device form: PERSISTENT THREADS.  The grid is sized to the machine (SMs x resident CTAs), each warp pulls batches
// of 32 rays from a global counter until the queue is empty, so long and short rays balance across the chip without
// a tail of half-empty CTAs.  The per-lane traversal stack lives in shared memory, laid out [entry][thread] so the
// 32 lanes of a warp always hit 32 different banks.  Node and triangle records are fetched as 128-bit loads through
// the read-only path (see traverse.cuh for the layouts).
#pragma once
#include "traverse.cuh"

namespace spt
{
	// ray queue record: 32 bytes in, 16 bytes out ("48 B per ray")
	struct alignas(16) RayRec { float ox, oy, oz; uint32_t ignoreTri; float dx, dy, dz; float tmax; };
	static_assert(sizeof(RayRec) == 32 && sizeof(Hit) == 16, "ray queue layout");

#if !defined(SPT_EMU)
	constexpr int kTraceBlock = 128;

	struct SmemStack
	{
		uint32_t* base; int n;     // base already offset by threadIdx.x; stride = blockDim.x
		__device__ __forceinline__ void clear() { n = 0; }
		__device__ __forceinline__ bool empty() const { return n == 0; }
		__device__ __forceinline__ void push(uint32_t v) { base[n * kTraceBlock] = v; n++; }
		__device__ __forceinline__ uint32_t pop() { n--; return base[n * kTraceBlock]; }
	};

	__device__ __forceinline__ V4 ldg4(const V4* p) { const float4 f = __ldg(reinterpret_cast<const float4*>(p)); return v4(f.x, f.y, f.z, f.w); }

	// nPtr (optional): the queue length lives in device memory (wavefront levels); n is then its capacity
	__global__ void __launch_bounds__(kTraceBlock) k_trace_rays(BvhView bvh, const RayRec* __restrict__ rays, Hit* __restrict__ hits,
		uint32_t n, const uint32_t* __restrict__ nPtr, uint32_t* __restrict__ counter)
	{
		if (nPtr) { const uint32_t m = *nPtr; if (m < n) n = m; }
		__shared__ uint32_t stackMem[kStackDepth * kTraceBlock];
		SmemStack stack; stack.base = stackMem + threadIdx.x; stack.n = 0;
		const uint32_t lane = threadIdx.x & 31;
		for (;;)
		{
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(counter, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= n) break;
			const uint32_t i = base + lane;
			if (i < n)
			{
				const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays + i));
				const float4 r1 = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
				if (r1.w < 0.0f) continue;     // idle pool slot
				Hit h;
				TraceClosest(bvh, v3(r0.x, r0.y, r0.z), v3(r1.x, r1.y, r1.z), __float_as_uint(r0.w), r1.w, stack, h);
				*reinterpret_cast<float4*>(hits + i) = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
			}
		}
	}

	// primary rays of sample 0 generated in-kernel (no ray queue traffic): pixel index = y*width + x, task order
	__global__ void __launch_bounds__(kTraceBlock) k_trace_primary(BvhView bvh, CameraGpu cam, Hit* __restrict__ hits, uint32_t* __restrict__ counter)
	{
		__shared__ uint32_t stackMem[kStackDepth * kTraceBlock];
		SmemStack stack; stack.base = stackMem + threadIdx.x; stack.n = 0;
		const uint32_t lane = threadIdx.x & 31;
		// 8x4 pixel tiles per warp keep the 32 rays of a batch spatially coherent
		const uint32_t tilesX = (cam.width + 7) / 8, tilesY = (cam.height + 3) / 4;
		const uint32_t n = tilesX * tilesY * 32u;
		for (;;)
		{
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(counter, 32u);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= n) break;
			{
				const uint32_t tile = base / 32, tx = tile % tilesX, ty = tile / tilesX;
				const uint32_t x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
				if (x < cam.width && y < cam.height)
				{
					Hit h;
					TraceClosest(bvh, cam.pos, PrimaryDir(cam, x, y, 0.5f, 0.5f), kNoHit, kFltMax, stack, h);
					*reinterpret_cast<float4*>(hits + (size_t)y * cam.width + x) = make_float4(h.t, h.u, h.v, __uint_as_float(h.tri));
				}
			}
		}
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

	inline void LaunchTraceRays(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t n, uint32_t* counter, const uint32_t* nPtr = nullptr)
	{
		if (!n || !ctx.ok) return;
		DevMemset(ctx, counter, 0, sizeof(uint32_t));
		k_trace_rays<<<TraceGridSize(), kTraceBlock, 0, ctx.stream>>>(bvh, rays, hits, n, nPtr, counter);
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
	inline void LaunchTraceRays(Ctx& ctx, const BvhView& bvh, const RayRec* rays, Hit* hits, uint32_t n, uint32_t*, const uint32_t* nPtr = nullptr)
	{
		LocalStack st;
		if (nPtr && *nPtr < n) n = *nPtr;
		for (uint32_t i = 0; i < n; i++)
			if (!(rays[i].tmax < 0.0f)) TraceClosest(bvh, v3(rays[i].ox, rays[i].oy, rays[i].oz), v3(rays[i].dx, rays[i].dy, rays[i].dz), rays[i].ignoreTri, rays[i].tmax, st, hits[i]);
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
