// textures.cuh — texel conversion and the software sampler.
//
// CombinedSampler2D::Initialize (reference MaterialUtils.h:42-65) turns every RGBA8 texel into float texels at load:
//   normal maps  texel * (1/127.5) - 1          colour  SRGBToLinear(texel * (1/255))  (alpha: texel * (1/255))
//   data maps    texel * (1/255)
// An 8-bit input has 256 possible values per channel, so the conversion is a 256-entry table per mode; the tables
// are evaluated on the host with the same libm powf the reference build uses (Core/Utils.cpp:59-63), which makes
// the float texels bit-identical to the reference's, and the device kernel is a pure gather (RGBA8 in, float4 out).
//
// CombinedSampler2D::Sample (MaterialUtils.h:75-124): wrap (Clamp: clamp(uv,0,1); Repeat: uv - floor(uv)), scale by
// (w-1),(h-1) (texel-corner convention, NOT the half-texel convention of hardware filtering), clamped +1 neighbour
// (no wrap across the seam), three lerps  a + f*(b-a)  in x, x, y order.  Hardware bilinear filtering (9-bit
// weights, u*w-0.5) cannot reproduce that, so texels are point-fetched as one LDG.128 each and blended in fp32.
#pragma once
#include "pipeline.cuh"

namespace spt
{
	struct TexConvertKernel
	{
		const uint32_t* rgba; V4* out; const float* lutR; const float* lutA; uint32_t channels;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t p = rgba[i];
			out[i] = v4(lutR[p & 255u], lutR[(p >> 8) & 255u], lutR[(p >> 16) & 255u], channels == 4 ? lutA[p >> 24] : 0.0f);
		}
	};

	inline void BuildTexelLuts(std::vector<float>& srgb, std::vector<float>& linear, std::vector<float>& normal)
	{
		srgb.resize(256); linear.resize(256); normal.resize(256);
		for (int i = 0; i < 256; i++)
		{
			const float s = (float)i * (1.0f / 255.0f);
			linear[i] = s;
			normal[i] = ((float)i * (1.0f / 127.5f)) - 1.0f;
			// Utils::SRGBToLinear (Core/Utils.cpp:59-63): mix(s/12.92, pow((s+0.055)/1.055, 2.4), step(0.04045, s)), arithmetic mix
			const float a = s < 0.04045f ? 0.0f : 1.0f;
			const float lo = s / 12.92f;
			const float hi = std::pow((s + 0.055f) / 1.055f, 2.4f);
			srgb[i] = lo * (1.0f - a) + hi * a;
		}
	}

	inline int SceneDevice::UploadTextures()
	{
		hostTextures.clear();
		uint64_t total = 0;
		for (const auto& t : Host().textures)
		{
			DeviceTexture d; d.width = (uint32_t)t.width; d.height = (uint32_t)t.height; d.channels = t.channels; d.clamping = t.clamping; d.offset = total;
			total += (uint64_t)t.width * t.height;
			hostTextures.push_back(d);
		}
		if (hostTextures.empty()) return SAILOR_PT_OK;
		std::vector<float> srgb, linear, normal;
		BuildTexelLuts(srgb, linear, normal);
		DevBuf<float> dSrgb, dLinear, dNormal; DevBuf<uint32_t> staging;
		dSrgb.Upload(ctx, srgb); dLinear.Upload(ctx, linear); dNormal.Upload(ctx, normal);
		texels.Alloc(ctx, total);
		textures.Upload(ctx, hostTextures);
		if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
		for (size_t i = 0; i < Host().textures.size(); i++)
		{
			const HostTexture& t = Host().textures[i];
			const uint32_t n = (uint32_t)t.width * (uint32_t)t.height;
			if (!n) continue;
			staging.Ensure(ctx, n);
			DevUpload(ctx, staging.p, t.rgba.data(), (size_t)n * 4);
			TexConvertKernel k;
			k.rgba = staging.p; k.out = texels.p + hostTextures[i].offset; k.channels = t.channels;
			k.lutR = t.normalMap ? dNormal.p : (t.srgb ? dSrgb.p : dLinear.p);
			k.lutA = dLinear.p;
			launch_for(ctx, n, k);
			ctx.Sync();   // staging is reused
		}
		return ctx.ok ? SAILOR_PT_OK : SAILOR_PT_ERR_CUDA;
	}

	struct TextureSet { const V4* texels; const DeviceTexture* textures; };

	// CombinedSampler2D::Sample<T> (MaterialUtils.h:75-124); vec3 textures carry w = 0
	SPT_HD V4 SampleTexture(const TextureSet& ts, uint32_t index, float u, float v)
	{
		const DeviceTexture t = ts.textures[index];
		float wu, wv;
		if (t.clamping == kClamp) { wu = std_clamp(u, 0.0f, 1.0f); wv = std_clamp(v, 0.0f, 1.0f); }
		else { wu = u - floorf(u); wv = v - floorf(v); }
		const int32_t W = (int32_t)t.width, H = (int32_t)t.height;
		const float fx = wu * (float)(W - 1), fy = wv * (float)(H - 1);
		const int32_t x0 = (int32_t)fx, y0 = (int32_t)fy;
		const int32_t x1 = (x0 + 1) < (W - 1) ? (x0 + 1) : (W - 1);    // std::min(tX0 + 1, m_width - 1)
		const int32_t y1 = (y0 + 1) < (H - 1) ? (y0 + 1) : (H - 1);
		const float fracX = fx - (float)x0, fracY = fy - (float)y0;
		const V4* base = ts.texels + t.offset;
		const V4 tl = ld4(base + x0 + (int64_t)y0 * W), tr = ld4(base + x1 + (int64_t)y0 * W);
		const V4 bl = ld4(base + x0 + (int64_t)y1 * W), br = ld4(base + x1 + (int64_t)y1 * W);
		V4 r;
		{
			const float top = tl.x + fracX * (tr.x - tl.x), bot = bl.x + fracX * (br.x - bl.x); r.x = top + fracY * (bot - top);
		}
		{
			const float top = tl.y + fracX * (tr.y - tl.y), bot = bl.y + fracX * (br.y - bl.y); r.y = top + fracY * (bot - top);
		}
		{
			const float top = tl.z + fracX * (tr.z - tl.z), bot = bl.z + fracX * (br.z - bl.z); r.z = top + fracY * (bot - top);
		}
		{
			const float top = tl.w + fracX * (tr.w - tl.w), bot = bl.w + fracX * (br.w - bl.w); r.w = top + fracY * (bot - top);
		}
		return r;
	}

	struct SampleTextureKernel
	{
		TextureSet ts; uint32_t index; const V2* uv; V4* out;
		SPT_KERNEL_BODY void operator()(uint32_t i) const { out[i] = SampleTexture(ts, index, uv[i].x, uv[i].y); }
	};
}
