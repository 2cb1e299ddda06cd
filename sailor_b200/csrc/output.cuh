// output.cuh — the output stage: chromatic-aberration taps + sRGB OETF + 8-bit quantisation.
//
// Restates the (commented-out) tail of PathTracer::Run (reference Raytracing/PathTracer.cpp:535-565): for every
// pixel three bilinear taps of the float accumulator through CombinedSampler2D::Sample with CLAMP addressing
// (MaterialUtils.h:75-124, so u maps to u*(w-1)), G at uv+(a,0), B at uv+(a,a), R at uv-(a,a), a = 0.5/width; then
// Utils::LinearToSRGB (Core/Utils.cpp:48-57), *255, clamp, truncate to u8.  One thread per pixel, coalesced RGB8
// stores; the 12 accumulator texels a pixel touches are neighbours, so the stage streams the image once from L2.
// powf is glibc's own algorithm restated (glibc_powf.h), so the bytes equal the reference's.
#pragma once
#include "backend.h"
#include "glibc_powf.h"

namespace spt
{
	SPT_HD float OutputTap(const float* img, uint32_t W, uint32_t H, float u, float v, int channel)
	{
		const float wu = std_clamp(u, 0.0f, 1.0f), wv = std_clamp(v, 0.0f, 1.0f);
		const int32_t w = (int32_t)W, h = (int32_t)H;
		const float fx = wu * (float)(w - 1), fy = wv * (float)(h - 1);
		const int32_t x0 = (int32_t)fx, y0 = (int32_t)fy;
		const int32_t x1 = (x0 + 1) < (w - 1) ? (x0 + 1) : (w - 1), y1 = (y0 + 1) < (h - 1) ? (y0 + 1) : (h - 1);
		const float fracX = fx - (float)x0, fracY = fy - (float)y0;
		const float tl = img[((size_t)y0 * W + x0) * 3 + channel], tr = img[((size_t)y0 * W + x1) * 3 + channel];
		const float bl = img[((size_t)y1 * W + x0) * 3 + channel], br = img[((size_t)y1 * W + x1) * 3 + channel];
		const float top = tl + fracX * (tr - tl), bot = bl + fracX * (br - bl);
		return top + fracY * (bot - top);
	}

	SPT_HD float LinearToSrgb(float c)                 // Core/Utils.cpp:48-57
	{
		if (c < 0.0031308f) return c * 12.92f;
		// glm::pow -> glibc powf: restated bit for bit (glibc_powf.h); inf / NaN accumulators take the library function
		const float e = 1.f / 2.4f;
		const float p = (GlibcPowfMainPath(c, e) && c < 1e30f) ? GlibcPowf(c, e) : powf(c, e);
		return 1.055f * p - 0.055f;
	}

	SPT_HD uint8_t Quantise(float s)                   // glm::clamp(x*255, 0, 255) -> u8 (PathTracer.cpp:557)
	{
		const float q = glm_clamp(s * 255.0f, 0.0f, 255.0f);
		return (uint8_t)q;
	}

	struct OutputKernel
	{
		const float* img; uint8_t* out; uint32_t W, H;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t x = i % W, y = i / W;
			const float ab = 0.5f / (float)W;
			const float u = (float)x / (float)W, v = (float)y / (float)H;
			const float g = OutputTap(img, W, H, u + ab, v + 0.0f, 1);
			const float b = OutputTap(img, W, H, u + ab, v + ab, 2);
			const float r = OutputTap(img, W, H, u + -ab, v + -ab, 0);
			out[(size_t)i * 3] = Quantise(LinearToSrgb(r));
			out[(size_t)i * 3 + 1] = Quantise(LinearToSrgb(g));
			out[(size_t)i * 3 + 2] = Quantise(LinearToSrgb(b));
		}
	};

	inline void RunOutputStage(Ctx& ctx, uint32_t W, uint32_t H, const float* dImg, uint8_t* dOut)
	{
		launch_for(ctx, W * H, OutputKernel{ dImg, dOut, W, H });
	}
}
