// gather_bench.cu — what one divergent record fetch costs on the L1TEX path of a B200 SM (tools/micro, not product code).
// Every lane fetches a random record of a table (64 MB: L2-resident like the C3 node array; 4 MB; 128 KB: L1-resident) with
//   A: 4 x LDG.128 (64 B)   B: 2 x LDG.256 (64 B)   C: 1 x LDG.256 (32 B)   D: 1 x LDG.128 (16 B)   E: 3 x LDG.128 (48 B)
// and with 1 / 2 / 4 / 8 / 32 lanes sharing a record.  Output: SM cycles per warp-level record fetch (all SMs busy, 32 warps per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(32) R8 { uint32_t v[8]; };
__device__ __forceinline__ void ld256(const void* p, uint32_t* r)
{
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ uint4 ld128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
template<int MODE>
__global__ void __launch_bounds__(256) k_gather(const unsigned char* __restrict__ tab, uint32_t recMask, uint32_t share, int iters, uint32_t* out)
{
	uint32_t lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t s = (gw * 977u + (lane / share) * 7919u + 12345u) * 2654435761u;
	uint32_t acc = 0;
	for (int i = 0; i < iters; i++)
	{
		s = s * 1664525u + 1013904223u;
		const unsigned char* p = tab + (size_t)((s >> 8) & recMask) * 64u;
		if (MODE == 0) { uint4 a = ld128(p), b = ld128(p + 16), c = ld128(p + 32), d = ld128(p + 48); acc += a.x ^ b.y ^ c.z ^ d.w; }
		if (MODE == 1) { uint32_t a[8], b[8]; ld256(p, a); ld256(p + 32, b); acc += a[0] ^ a[7] ^ b[1] ^ b[6]; }
		if (MODE == 2) { uint32_t a[8]; ld256(p, a); acc += a[0] ^ a[7]; }
		if (MODE == 3) { uint4 a = ld128(p); acc += a.x ^ a.w; }
		if (MODE == 4) { uint4 a = ld128(p), b = ld128(p + 16), c = ld128(p + 32); acc += a.x ^ b.y ^ c.z; }
		if (MODE == 5) { acc += s >> 9; }     // loop overhead only
	}
	if (acc == 0x12345u) out[0] = acc;
}
int main()
{
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const size_t maxBytes = 64u << 20;
	unsigned char* tab; cudaMalloc(&tab, maxBytes); cudaMemset(tab, 1, maxBytes);
	uint32_t* out; cudaMalloc(&out, 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const char* names[6] = { "4xLDG.128(64B)", "2xLDG.256(64B)", "1xLDG.256(32B)", "1xLDG.128(16B)", "3xLDG.128(48B)", "loop only" };
	const int iters = 2000, blocks = sms * 4;       // 4 x 256 threads = 32 warps per SM
	printf("SMs %d clock %d kHz\n", sms, clk);
	for (size_t bytes : { (size_t)64 << 20, (size_t)4 << 20, (size_t)128 << 10 })
		for (uint32_t share : { 1u, 2u, 4u, 8u, 32u })
			for (int mode = 0; mode < 6; mode++)
			{
				const uint32_t recMask = (uint32_t)(bytes / 64) - 1;
				float best = 1e30f;
				for (int rep = 0; rep < 3; rep++)
				{
					cudaEventRecord(e0);
					switch (mode)
					{
					case 0: k_gather<0><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					case 1: k_gather<1><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					case 2: k_gather<2><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					case 3: k_gather<3><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					case 4: k_gather<4><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					default: k_gather<5><<<blocks, 256>>>(tab, recMask, share, iters, out); break;
					}
					cudaEventRecord(e1); cudaEventSynchronize(e1);
					float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
				}
				const double warpFetchesPerSm = 32.0 * iters;
				const double cyc = best * 1e-3 * 1.965e9 / warpFetchesPerSm;   // at the 1965 MHz the pool's B200s run at under load
				printf("table %6zu KB  share %2u  %-16s %8.3f ms  %7.1f SM-cycles per warp fetch\n", bytes >> 10, share, names[mode], best, cyc);
			}
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return 0;
}
