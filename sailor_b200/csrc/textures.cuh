// textures.cuh — texel conversion and the software sampler.
//
// CombinedSampler2D::Initialize (reference MaterialUtils.h:42-65) turns every RGBA8 texel into float texels at load:
//   normal maps  texel * (1/127.5) - 1          colour  SRGBToLinear(texel * (1/255))  (alpha: texel * (1/255))
//   data maps    texel * (1/255)
// The reference keeps those float texels (16 bytes each); here the pool keeps the file's RGBA8 texels (4 bytes each: C4's 224
// 1024x1024 textures are 0.94 GB instead of 3.76 GB, and an import is one H2D copy per texture) and the conversion runs at fetch time.
// It is the same single-precision arithmetic on a byte, so the fetched values are bit-identical to the reference's: the two linear
// modes are one or two exact float operations, and the sRGB curve (which goes through libm's powf, Core/Utils.cpp:59-63) is a
// 256-entry table evaluated on the host with the powf the reference build uses.
//
// CombinedSampler2D::Sample (MaterialUtils.h:75-124): wrap (Clamp: clamp(uv,0,1); Repeat: uv - floor(uv)), scale by
// (w-1),(h-1) (texel-corner convention, NOT the half-texel convention of hardware filtering), clamped +1 neighbour
// (no wrap across the seam), three lerps  a + f*(b-a)  in x, x, y order.  Hardware bilinear filtering (9-bit
// weights, u*w-0.5) cannot reproduce that, so texels are point-fetched as one LDG.128 each and blended in fp32.
#pragma once
#include "pipeline.cuh"

namespace spt
{
	// Utils::SRGBToLinear (Core/Utils.cpp:59-63) of byte / 255: mix(s/12.92, pow((s+0.055)/1.055, 2.4), step(0.04045, s)), arithmetic mix,
	// evaluated on the host with the libm powf the reference build uses
	inline void BuildSrgbLut(std::vector<float>& srgb)
	{
		srgb.resize(256);
		for (int i = 0; i < 256; i++)
		{
			const float s = (float)i * (1.0f / 255.0f);
			const float a = s < 0.04045f ? 0.0f : 1.0f;
			const float lo = s / 12.92f;
			const float hi = std::pow((s + 0.055f) / 1.055f, 2.4f);
			srgb[i] = lo * (1.0f - a) + hi * a;
		}
	}

	inline int SceneDevice::UploadTextures()
	{
		hostTextures.clear();
		uint64_t total = 0, totalF = 0;
		for (const auto& t : Host().textures)
		{
			const bool isFloat = !t.rgbaF.empty();
			DeviceTexture d; d.width = (uint32_t)t.width; d.height = (uint32_t)t.height; d.channels = t.channels; d.clamping = t.clamping; d.offset = isFloat ? totalF : total;
			d.mode = isFloat ? kTexelFloat : (t.normalMap ? kTexelNormal : (t.srgb ? kTexelSrgb : kTexelLinear)); d.pad = 0;
			(isFloat ? totalF : total) += (uint64_t)t.width * t.height;
			hostTextures.push_back(d);
		}
		if (hostTextures.empty()) return SAILOR_PT_OK;
		std::vector<float> srgb;
		BuildSrgbLut(srgb);
		srgbLut.Upload(ctx, srgb);
		texels.Alloc(ctx, total ? total : 1);
		if (totalF) texelsF.Alloc(ctx, totalF);
		textures.Upload(ctx, hostTextures);
		if (!ctx.ok) return SAILOR_PT_ERR_CUDA;
		// the file's RGBA8 texels go to the pool as they are: one transfer per texture, no staging, no conversion pass
		for (size_t i = 0; i < Host().textures.size(); i++)
		{
			const HostTexture& t = Host().textures[i];
			const size_t n = (size_t)t.width * (size_t)t.height;
			if (!n) continue;
			if (!t.rgbaF.empty())
			{
				// CombinedSampler2D::Initialize<T, vec4> (MaterialUtils.h:42-65) on the float pixels of a .hdr image, with the host's powf
				std::vector<V4> conv(n);
				for (size_t k = 0; k < n; k++)
				{
					const float* src = t.rgbaF.data() + k * 4; float o[4];
					for (int c = 0; c < 4; c++)
					{
						if (t.normalMap) o[c] = (src[c] * (1.0f / 127.5f)) - 1.0f;
						else
						{
							const float s = src[c] * (1.0f / 255.0f);
							if (t.srgb && c < 3) { const float a = s < 0.04045f ? 0.0f : 1.0f; o[c] = (s / 12.92f) * (1.0f - a) + std::pow((s + 0.055f) / 1.055f, 2.4f) * a; }
							else o[c] = s;
						}
					}
					conv[k] = v4(o[0], o[1], o[2], t.channels == 4 ? o[3] : 0.0f);
				}
				DevUpload(ctx, texelsF.p + hostTextures[i].offset, conv.data(), n * sizeof(V4));
				ctx.Sync();          // `conv` dies here
				continue;
			}
			DevUpload(ctx, texels.p + hostTextures[i].offset, t.rgba.data(), n * 4);
		}
		return ctx.ok ? SAILOR_PT_OK : SAILOR_PT_ERR_CUDA;
	}

	struct TextureSet { const uint32_t* texels; const DeviceTexture* textures; const float* srgbLut; const V4* texelsF; };

	// CombinedSampler2D::Initialize (MaterialUtils.h:42-65) for one texel: the same single-precision expressions on the byte (normal maps
	// byte * (1/127.5) - 1, data maps and every alpha byte * (1/255)); sRGB colour through the 256-entry table.  vec3 textures carry w = 0.
	SPT_HD V4 TexelToFloat(uint32_t p, uint32_t mode, bool hasAlpha, const float* srgbLut)
	{
		const uint32_t r = p & 255u, g = (p >> 8) & 255u, b = (p >> 16) & 255u, a = p >> 24;
		V4 o;
		if (mode == kTexelSrgb) { o.x = ldf(srgbLut + r); o.y = ldf(srgbLut + g); o.z = ldf(srgbLut + b); }
		else if (mode == kTexelNormal) { o.x = ((float)r * (1.0f / 127.5f)) - 1.0f; o.y = ((float)g * (1.0f / 127.5f)) - 1.0f; o.z = ((float)b * (1.0f / 127.5f)) - 1.0f; }
		else { o.x = (float)r * (1.0f / 255.0f); o.y = (float)g * (1.0f / 255.0f); o.z = (float)b * (1.0f / 255.0f); }
		o.w = hasAlpha ? (float)a * (1.0f / 255.0f) : 0.0f;
		return o;
	}

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
		V4 tl, tr, bl, br;
		if (t.mode == kTexelFloat)
		{
			const V4* base = ts.texelsF + t.offset;
			tl = ld4(base + x0 + (int64_t)y0 * W); tr = ld4(base + x1 + (int64_t)y0 * W); bl = ld4(base + x0 + (int64_t)y1 * W); br = ld4(base + x1 + (int64_t)y1 * W);
		}
		else
		{
			const uint32_t* base = ts.texels + t.offset;
			const bool alpha = t.channels == 4;
			const uint32_t ptl = ldu(base + x0 + (int64_t)y0 * W), ptr = ldu(base + x1 + (int64_t)y0 * W), pbl = ldu(base + x0 + (int64_t)y1 * W), pbr = ldu(base + x1 + (int64_t)y1 * W);
			tl = TexelToFloat(ptl, t.mode, alpha, ts.srgbLut); tr = TexelToFloat(ptr, t.mode, alpha, ts.srgbLut);
			bl = TexelToFloat(pbl, t.mode, alpha, ts.srgbLut); br = TexelToFloat(pbr, t.mode, alpha, ts.srgbLut);
		}
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
