// glibc_powf.h — powf as the reference's host computes it, bit for bit.
//
// The reference's output stage and texel conversion call glm::pow -> std::pow(float, float) = glibc powf
// (Core/Utils.cpp:48-64).  glibc's powf is NOT correctly rounded (error < 0.52 ulp): about 6e-4 of its results differ from a
// correctly rounded power in the last bit, so neither CUDA's powf nor a double-precision pow rounded once reproduces it.  This
// header restates the published algorithm of glibc 2.39 (sysdeps/ieee754/flt-32/e_powf.c, e_powf_log2_data.c,
// e_exp2f_data.c — Szabolcs Nagy's implementation from ARM's optimized-routines): log2(x) from a 16-entry table + a degree-5
// polynomial in double, y*log2(x), 2^x from a 32-entry table + a degree-3 polynomial in double, one final rounding to float.
// On x86-64 with FMA (every host this runs beside) glibc dispatches to its -mfma build, in which every a*b+c of that code is one
// fused operation: the fma() calls below are exactly those.  The constants are glibc's (third-party dependency of the reference,
// not vendored under /root/reference).  Checked against the host's powf on 76 M arguments (exponents 1/2.4 and 2.4): 0 mismatches
// (tests/test_host_logic.py::test_powf_restatement_equals_the_hosts_powf).
// Only the main path is restated (x positive and normal, result in range); callers route anything else to the library powf.
#pragma once
#include "hd.h"

namespace spt
{
#define SPT_POWF_LOG2_TABLE { \
		0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2, \
		0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2, \
		0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3, \
		0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4, \
		0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1.0000000000000p+0, 0x0.0p+0, \
		0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, 0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3, \
		0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, \
		0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 }
#define SPT_POWF_LOG2_POLY { 0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0 }
#define SPT_EXP2F_TABLE { \
		0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
		0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
		0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
		0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
		0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
		0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
		0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
		0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL }
#define SPT_EXP2F_POLY { 0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1 }
#if defined(__CUDACC__)
	__device__ const double kPowfLog2TabDev[32] = SPT_POWF_LOG2_TABLE;
	__device__ const double kPowfLog2PolyDev[5] = SPT_POWF_LOG2_POLY;
	__device__ const unsigned long long kExp2fTabDev[32] = SPT_EXP2F_TABLE;
	__device__ const double kExp2fPolyDev[3] = SPT_EXP2F_POLY;
#endif
	static const double kPowfLog2TabHost[32] = SPT_POWF_LOG2_TABLE;
	static const double kPowfLog2PolyHost[5] = SPT_POWF_LOG2_POLY;
	static const unsigned long long kExp2fTabHost[32] = SPT_EXP2F_TABLE;
	static const double kExp2fPolyHost[3] = SPT_EXP2F_POLY;

	SPT_HD bool GlibcPowfMainPath(float x, float y)
	{
		const uint32_t ix = f2u(x);
		return ix - 0x00800000u < 0x7f800000u - 0x00800000u && y == y && fabsf(y) < 64.0f;      // positive normal x, modest finite y
	}

	// powf(x, y) for arguments on the main path (GlibcPowfMainPath) whose result is a normal float
	SPT_HD float GlibcPowf(float x, float y)
	{
#if defined(__CUDA_ARCH__)
		const double* T = kPowfLog2TabDev; const double* A = kPowfLog2PolyDev; const unsigned long long* E = kExp2fTabDev; const double* C = kExp2fPolyDev;
#define SPT_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
		const double* T = kPowfLog2TabHost; const double* A = kPowfLog2PolyHost; const unsigned long long* E = kExp2fTabHost; const double* C = kExp2fPolyHost;
#define SPT_FMA(a, b, c) fma((a), (b), (c))
#endif
		// log2_inline (e_powf.c): x = 2^k z, z in [0x1.66p-1, 0x1.66p0), log2(x) = log1p(z/c - 1)/ln2 + log2(c) + k
		const uint32_t ix = f2u(x);
		const uint32_t tmp = ix - 0x3f330000u;
		const uint32_t i = (tmp >> (23 - 4)) % 16u;
		const uint32_t top = tmp & 0xff800000u;
		const uint32_t iz = ix - top;
		const int32_t k = (int32_t)top >> 23;
		const double invc = T[2 * i], logc = T[2 * i + 1];
		const double z = (double)u2f(iz);
		const double r = SPT_FMA(z, invc, -1.0);
		const double y0 = logc + (double)k;
		const double r2 = r * r;
		double yy = SPT_FMA(A[0], r, A[1]);
		const double p = SPT_FMA(A[2], r, A[3]);
		const double r4 = r2 * r2;
		double q = SPT_FMA(A[4], r, y0);
		q = SPT_FMA(p, r2, q);
		yy = SPT_FMA(yy, r4, q);
		const double xd = (double)y * yy;
		// exp2_inline: xd = k/32 + r, 2^xd = 2^(k/32) (C0 r^3 + C1 r^2 + C2 r + 1)
		const double shift = 211106232532992.0;           // 0x1.8p+52 / 32
		double kd = xd + shift;
		union { double d; unsigned long long u; } cv; cv.d = kd;
		const unsigned long long ki = cv.u;
		kd -= shift;
		const double rr = xd - kd;
		cv.u = E[ki % 32u] + (ki << (52 - 5));
		const double s = cv.d;
		const double zz = SPT_FMA(C[0], rr, C[1]);
		const double rr2 = rr * rr;
		double res = SPT_FMA(C[2], rr, 1.0);
		res = SPT_FMA(zz, rr2, res);
		return (float)(res * s);
#undef SPT_FMA
	}
}
