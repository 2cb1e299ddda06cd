// png_codec.cpp — PNG decode (textures in) and encode (image out), host C++ over zlib.
//
// The reference decodes textures with stbi_load(..., STBI_rgb_alpha) (reference MaterialUtils.h:226-249) and writes
// the result with stbi_write_png (PathTracer.cpp:560-564).  This decoder produces the same RGBA8 texels stb does:
// 16-bit samples keep the high byte, 1/2/4-bit greys are scaled to 0..255, palettes and tRNS keys are expanded.
// Only the pixel values are part of the parity contract, not the file bytes.
#include "host_scene.h"
#include "../../include/sailor_pt.h"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace spt
{
	namespace
	{
		uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

		int Paeth(int a, int b, int c)
		{
			const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
			if (pa <= pb && pa <= pc) return a;
			return pb <= pc ? b : c;
		}

		bool Unfilter(std::vector<uint8_t>& raw, size_t offset, uint32_t rowBytes, uint32_t rows, uint32_t bpp, uint8_t* out)
		{
			const uint8_t* src = raw.data() + offset;
			for (uint32_t y = 0; y < rows; y++)
			{
				const uint8_t filter = *src++;
				uint8_t* cur = out + (size_t)y * rowBytes;
				const uint8_t* up = y ? cur - rowBytes : nullptr;
				for (uint32_t x = 0; x < rowBytes; x++)
				{
					const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
					int v = src[x];
					switch (filter)
					{
					case 0: break;
					case 1: v += a; break;
					case 2: v += b; break;
					case 3: v += (a + b) >> 1; break;
					case 4: v += Paeth(a, b, c); break;
					default: return false;
					}
					cur[x] = (uint8_t)v;
				}
				src += rowBytes;
			}
			return true;
		}
	}

	int DecodePngRgba8(const uint8_t* data, size_t size, int32_t& w, int32_t& h, std::vector<uint8_t>& rgba, std::string& err)
	{
		static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
		if (size < 8 || memcmp(data, sig, 8)) { err = "not a PNG"; return SAILOR_PT_ERR_FORMAT; }
		uint32_t width = 0, height = 0; int depth = 0, color = 0, interlace = 0;
		std::vector<uint8_t> idat, palette, trns;
		size_t off = 8; bool haveHdr = false, done = false;
		// chunk walk with stb_image's tolerance: CRCs are never checked (the last one may even be missing), the file must reach an IEND
		// chunk header, a chunk's body must be complete, and an unknown CRITICAL chunk (upper-case first letter) is an error
		while (!done)
		{
			if (off + 8 > size) { err = "PNG ends before IEND"; return SAILOR_PT_ERR_FORMAT; }
			const uint32_t len = be32(data + off);
			const uint8_t* type = data + off + 4;
			const uint8_t* body = data + off + 8;
			if (!memcmp(type, "IEND", 4)) { done = true; break; }
			if (off + 8 + (size_t)len > size) { err = "truncated PNG"; return SAILOR_PT_ERR_FORMAT; }
			if (!memcmp(type, "IHDR", 4))
			{
				if (len != 13 || haveHdr) { err = "bad PNG header"; return SAILOR_PT_ERR_FORMAT; }
				width = be32(body); height = be32(body + 4); depth = body[8]; color = body[9]; interlace = body[12]; haveHdr = true;
				if (body[10] || body[11]) { err = "bad PNG compression / filter method"; return SAILOR_PT_ERR_FORMAT; }
			}
			else if (!haveHdr) { err = "PNG: first chunk is not IHDR"; return SAILOR_PT_ERR_FORMAT; }
			else if (!memcmp(type, "PLTE", 4)) { if (len > 256 * 3 || len % 3) { err = "bad PNG palette"; return SAILOR_PT_ERR_FORMAT; } palette.assign(body, body + len); }
			else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
			else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
			else if (!(type[0] & 32)) { err = "unknown critical PNG chunk"; return SAILOR_PT_ERR_FORMAT; }
			off += 12 + (size_t)len;
		}
		if (!haveHdr || !width || !height || width > 65536 || height > 65536) { err = "bad PNG header"; return SAILOR_PT_ERR_FORMAT; }
		if (interlace > 1) { err = "bad PNG interlace method"; return SAILOR_PT_ERR_FORMAT; }
		int channels;
		switch (color) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break; default: err = "bad PNG colour type"; return SAILOR_PT_ERR_FORMAT; }
		if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "bad PNG depth"; return SAILOR_PT_ERR_FORMAT; }
		const uint32_t bitsPerPixel = (uint32_t)channels * depth;
		const uint32_t bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;

		// Adam7 (interlace method 1): seven reduced images, each filtered on its own; pass p holds the pixels (yOrig + j * ySpc, xOrig + i * xSpc)
		static const uint32_t xOrig[7] = { 0, 4, 0, 2, 0, 1, 0 }, yOrig[7] = { 0, 0, 4, 0, 2, 0, 1 }, xSpc[7] = { 8, 8, 4, 4, 2, 2, 1 }, ySpc[7] = { 8, 8, 8, 4, 4, 2, 2 };
		const int passes = interlace ? 7 : 1;
		size_t rawSize = 0;
		for (int ps = 0; ps < passes; ps++)
		{
			const uint32_t pw = interlace ? (width > xOrig[ps] ? (width - xOrig[ps] + xSpc[ps] - 1) / xSpc[ps] : 0u) : width, ph = interlace ? (height > yOrig[ps] ? (height - yOrig[ps] + ySpc[ps] - 1) / ySpc[ps] : 0u) : height;
			if (pw && ph) rawSize += (size_t)((pw * bitsPerPixel + 7) / 8 + 1) * ph;
		}
		std::vector<uint8_t> raw(rawSize);
		uLongf rawLen = (uLongf)raw.size();
		// zlib stream: the 2-byte header is checked the way stb_image does (method 8, FCHECK, no preset dictionary), the deflate data is
		// inflated raw and the Adler-32 trailer is NOT verified (stb_image never looks at it: a file with a damaged checksum still loads)
		if (idat.size() < 2 || (idat[0] & 15) != 8 || ((idat[0] << 8) | idat[1]) % 31 != 0 || (idat[1] & 32)) { err = "bad zlib header"; return SAILOR_PT_ERR_FORMAT; }
		{
			z_stream zs; memset(&zs, 0, sizeof(zs));
			if (inflateInit2(&zs, -15) != Z_OK) { err = "PNG inflate failed"; return SAILOR_PT_ERR_FORMAT; }
			zs.next_in = idat.data() + 2; zs.avail_in = (uInt)(idat.size() - 2);
			zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
			const int zr = inflate(&zs, Z_FINISH);
			rawLen = (uLongf)zs.total_out;
			inflateEnd(&zs);
			// stb_image accepts more decoded bytes than the image needs and rejects fewer ("not enough pixels")
			if ((zr != Z_STREAM_END && zr != Z_BUF_ERROR && zr != Z_OK) || rawLen != raw.size()) { err = "PNG inflate failed"; return SAILOR_PT_ERR_FORMAT; }
		}

		w = (int32_t)width; h = (int32_t)height;
		rgba.resize((size_t)width * height * 4);
		static const uint8_t depthScale[9] = { 0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01 };
		auto pixel = [&](const uint8_t* row, uint32_t x, uint8_t* o)
			{
				uint32_t s[4] = { 0, 0, 0, 0 };  // raw samples (full depth)
				for (int c = 0; c < channels; c++)
				{
					if (depth == 8) s[c] = row[x * channels + c];
					else if (depth == 16) s[c] = ((uint32_t)row[(x * channels + c) * 2] << 8) | row[(x * channels + c) * 2 + 1];
					else { const uint32_t bit = x * depth; s[c] = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1); }
				}
				auto to8 = [&](uint32_t v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : (depth < 8 ? (uint8_t)(v * depthScale[depth]) : (uint8_t)v); };
				switch (color)
				{
				case 0:
				{
					o[0] = o[1] = o[2] = to8(s[0]); o[3] = 255;
					if (trns.size() >= 2 && s[0] == (((uint32_t)trns[0] << 8) | trns[1])) o[3] = 0;
					break;
				}
				case 2:
				{
					o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]); o[3] = 255;
					if (trns.size() >= 6 && s[0] == (((uint32_t)trns[0] << 8) | trns[1]) && s[1] == (((uint32_t)trns[2] << 8) | trns[3]) && s[2] == (((uint32_t)trns[4] << 8) | trns[5])) o[3] = 0;
					break;
				}
				case 3:
				{
					const uint32_t i = s[0];
					if ((i + 1) * 3 <= palette.size()) { o[0] = palette[i * 3]; o[1] = palette[i * 3 + 1]; o[2] = palette[i * 3 + 2]; }
					else { o[0] = o[1] = o[2] = 0; }
					o[3] = i < trns.size() ? trns[i] : 255;
					break;
				}
				case 4: o[0] = o[1] = o[2] = to8(s[0]); o[3] = to8(s[1]); break;
				default: o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]); o[3] = to8(s[3]); break;
				}
			};
		std::vector<uint8_t> img;
		size_t rawOff = 0;
		for (int ps = 0; ps < passes; ps++)
		{
			const uint32_t x0 = interlace ? xOrig[ps] : 0, y0 = interlace ? yOrig[ps] : 0, dx = interlace ? xSpc[ps] : 1, dy = interlace ? ySpc[ps] : 1;
			const uint32_t pw = width > x0 ? (width - x0 + dx - 1) / dx : 0u, ph = height > y0 ? (height - y0 + dy - 1) / dy : 0u;
			if (!pw || !ph) continue;
			const uint32_t passRowBytes = (pw * bitsPerPixel + 7) / 8;
			img.assign((size_t)passRowBytes * ph, 0);
			if (!Unfilter(raw, rawOff, passRowBytes, ph, bpp, img.data())) { err = "bad PNG filter"; return SAILOR_PT_ERR_FORMAT; }
			rawOff += (size_t)(passRowBytes + 1) * ph;
			for (uint32_t j = 0; j < ph; j++)
			{
				const uint8_t* row = img.data() + (size_t)j * passRowBytes;
				for (uint32_t i = 0; i < pw; i++) pixel(row, i, rgba.data() + ((size_t)(y0 + j * dy) * width + (x0 + i * dx)) * 4);
			}
		}
		return SAILOR_PT_OK;
	}

	int EncodePngRgb8(const char* path, uint32_t w, uint32_t h, const uint8_t* rgb, std::string& err)
	{
		std::vector<uint8_t> raw((size_t)(w * 3 + 1) * h);
		for (uint32_t y = 0; y < h; y++)
		{
			raw[(size_t)y * (w * 3 + 1)] = 0;
			memcpy(raw.data() + (size_t)y * (w * 3 + 1) + 1, rgb + (size_t)y * w * 3, (size_t)w * 3);
		}
		uLongf zlen = compressBound((uLong)raw.size());
		std::vector<uint8_t> z(zlen);
		if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "PNG deflate failed"; return SAILOR_PT_ERR_IO; }
		FILE* f = fopen(path, "wb");
		if (!f) { err = std::string("cannot write ") + path; return SAILOR_PT_ERR_IO; }
		auto chunk = [&](const char* type, const uint8_t* body, uint32_t len)
			{
				uint8_t hdr[8] = { (uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3] };
				fwrite(hdr, 1, 8, f);
				if (len) fwrite(body, 1, len, f);
				uLong crc = crc32(0L, hdr + 4, 4);
				if (len) crc = crc32(crc, body, len);
				const uint8_t c[4] = { (uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc };
				fwrite(c, 1, 4, f);
			};
		static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
		fwrite(sig, 1, 8, f);
		const uint8_t ihdr[13] = { (uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0 };
		chunk("IHDR", ihdr, 13);
		chunk("IDAT", z.data(), (uint32_t)zlen);
		chunk("IEND", nullptr, 0);
		const bool ok = !ferror(f);
		fclose(f);
		if (!ok) { err = "write error"; return SAILOR_PT_ERR_IO; }
		return SAILOR_PT_OK;
	}
}
