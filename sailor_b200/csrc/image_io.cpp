// image_io.cpp — linear image dumps, image comparison and the progressive-render checkpoint file.
//
// The reference only writes the tonemapped sRGB8 PNG (stbi_write_png, reference PathTracer.cpp:560-564).  A host that
// renders long frames also wants the LINEAR accumulator on disk — to resume, to compare against another render, to
// tonemap later — so next to the PNG the product writes two lossless-enough float formats chosen by file extension:
//   .pfm  Portable Float Map: "PF\n<w> <h>\n-1.0\n" + fp32 RGB rows, bottom row first (exact bits of the accumulator)
//   .hdr  Radiance RGBE, flat (non-RLE) scanlines, top row first (8-bit mantissa shared exponent; for viewers)
// and a checkpoint file that holds the un-normalised running sum of a progressive render (see RenderProgressive in
// capi.cu): header (geometry + every parameter the estimate depends on) + fp32 payload + CRC-32 of both.
#include "host_scene.h"
#include "../../include/sailor_pt.h"

#include <zlib.h>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace spt
{
	namespace
	{
		bool EndsWith(const std::string& s, const char* suffix)
		{
			const size_t n = strlen(suffix);
			if (s.size() < n) return false;
			for (size_t i = 0; i < n; i++) { char c = s[s.size() - n + i]; if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a'); if (c != suffix[i]) return false; }
			return true;
		}
	}

	ImageFormat ImageFormatOf(const char* path)
	{
		const std::string s = path ? path : "";
		if (EndsWith(s, ".pfm")) return ImageFormat::Pfm;
		if (EndsWith(s, ".hdr")) return ImageFormat::Hdr;
		return ImageFormat::Png;
	}

	int WritePfm(const char* path, uint32_t w, uint32_t h, const float* rgb, std::string& err)
	{
		FILE* f = fopen(path, "wb");
		if (!f) { err = std::string("cannot open for writing: ") + path; return SAILOR_PT_ERR_IO; }
		fprintf(f, "PF\n%u %u\n-1.0\n", w, h);                                  // negative scale = little endian
		bool ok = true;
		for (uint32_t y = h; y-- > 0 && ok;) ok = fwrite(rgb + (size_t)y * w * 3, sizeof(float), (size_t)w * 3, f) == (size_t)w * 3;
		ok = (fclose(f) == 0) && ok;
		if (!ok) { err = std::string("short write: ") + path; return SAILOR_PT_ERR_IO; }
		return SAILOR_PT_OK;
	}

	int WriteHdr(const char* path, uint32_t w, uint32_t h, const float* rgb, std::string& err)
	{
		FILE* f = fopen(path, "wb");
		if (!f) { err = std::string("cannot open for writing: ") + path; return SAILOR_PT_ERR_IO; }
		fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", h, w);
		std::vector<uint8_t> row((size_t)w * 4);
		bool ok = true;
		for (uint32_t y = 0; y < h && ok; y++)
		{
			for (uint32_t x = 0; x < w; x++)
			{
				const float* p = rgb + ((size_t)y * w + x) * 3;
				float r = p[0] > 0.0f ? p[0] : 0.0f, g = p[1] > 0.0f ? p[1] : 0.0f, b = p[2] > 0.0f ? p[2] : 0.0f;   // RGBE has no sign; NaN fails the compare -> 0
				const float m = r > g ? (r > b ? r : b) : (g > b ? g : b);
				uint8_t* o = row.data() + (size_t)x * 4;
				if (!(m > 1e-32f) || !std::isfinite(m)) { o[0] = o[1] = o[2] = o[3] = 0; if (std::isinf(m)) { o[0] = o[1] = o[2] = 255; o[3] = 255; } continue; }
				int e = 0;
				const float scale = std::frexp(m, &e) * 256.0f / m;
				o[0] = (uint8_t)(r * scale); o[1] = (uint8_t)(g * scale); o[2] = (uint8_t)(b * scale); o[3] = (uint8_t)(e + 128);
			}
			ok = fwrite(row.data(), 1, row.size(), f) == row.size();
		}
		ok = (fclose(f) == 0) && ok;
		if (!ok) { err = std::string("short write: ") + path; return SAILOR_PT_ERR_IO; }
		return SAILOR_PT_OK;
	}

	// [0] mean relative error = mean|a-b| / mean|b| (the tolerance metric of the converged-image tests), [1] RMSE, [2] max |a-b|,
	// [3] PSNR in dB against peak 1.0 (inf -> 1e30 when the images are identical)
	void CompareImages(size_t count, const float* a, const float* b, double out[4])
	{
		double sumAbs = 0.0, sumRef = 0.0, sumSq = 0.0, mx = 0.0;
		for (size_t i = 0; i < count; i++)
		{
			const double d = std::fabs((double)a[i] - (double)b[i]);
			sumAbs += d; sumRef += std::fabs((double)b[i]); sumSq += d * d; if (d > mx) mx = d;
		}
		const double n = count ? (double)count : 1.0;
		out[0] = sumRef > 0.0 ? sumAbs / sumRef : (sumAbs > 0.0 ? 1e30 : 0.0);
		out[1] = std::sqrt(sumSq / n);
		out[2] = mx;
		out[3] = sumSq > 0.0 ? 10.0 * std::log10(1.0 / (sumSq / n)) : 1e30;
	}

	// ---- checkpoint -------------------------------------------------------------------------------------------------
	static const char kCkptMagic[8] = { 'S', 'P', 'T', 'C', 'K', 'P', 'T', '1' };

	int WriteCheckpoint(const char* path, const CheckpointHeader& hd, const float* runningSum, std::string& err)
	{
		const std::string tmp = std::string(path) + ".tmp";                      // write-then-rename: an interrupted write never leaves a half checkpoint behind
		FILE* f = fopen(tmp.c_str(), "wb");
		if (!f) { err = "cannot open for writing: " + tmp; return SAILOR_PT_ERR_IO; }
		const size_t count = (size_t)hd.width * (hd.rowEnd - hd.rowBegin) * 3;
		uLong crc = crc32(0L, Z_NULL, 0);
		crc = crc32(crc, reinterpret_cast<const Bytef*>(&hd), (uInt)sizeof(hd));
		const unsigned char* bytes = reinterpret_cast<const unsigned char*>(runningSum);
		for (size_t off = 0, total = count * sizeof(float); off < total;) { const size_t n = total - off < (1u << 30) ? total - off : (1u << 30); crc = crc32(crc, bytes + off, (uInt)n); off += n; }
		const uint32_t crc32v = (uint32_t)crc;
		bool ok = fwrite(kCkptMagic, 1, 8, f) == 8 && fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(runningSum, sizeof(float), count, f) == count && fwrite(&crc32v, 4, 1, f) == 1;
		ok = (fclose(f) == 0) && ok;
		if (!ok) { remove(tmp.c_str()); err = "short write: " + tmp; return SAILOR_PT_ERR_IO; }
		if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); err = std::string("cannot rename checkpoint into place: ") + path; return SAILOR_PT_ERR_IO; }
		return SAILOR_PT_OK;
	}

	int ReadCheckpoint(const char* path, CheckpointHeader& hd, std::vector<float>& runningSum, std::string& err)
	{
		FILE* f = fopen(path, "rb");
		if (!f) { err = std::string("no checkpoint at ") + path; return SAILOR_PT_ERR_IO; }
		char magic[8];
		bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kCkptMagic, 8) == 0 && fread(&hd, sizeof(hd), 1, f) == 1;
		if (ok) ok = hd.version == 1u && hd.width && hd.rowEnd > hd.rowBegin && hd.rowEnd <= hd.height && hd.msaaDone <= hd.msaaTotal && (uint64_t)hd.width * (hd.rowEnd - hd.rowBegin) < (1ull << 32);
		if (!ok) { fclose(f); err = std::string("not a sailor_pt checkpoint: ") + path; return SAILOR_PT_ERR_FORMAT; }
		const size_t count = (size_t)hd.width * (hd.rowEnd - hd.rowBegin) * 3;
		runningSum.resize(count);
		uint32_t stored = 0;
		ok = fread(runningSum.data(), sizeof(float), count, f) == count && fread(&stored, 4, 1, f) == 1;
		fclose(f);
		if (!ok) { err = std::string("truncated checkpoint: ") + path; return SAILOR_PT_ERR_FORMAT; }
		uLong crc = crc32(0L, Z_NULL, 0);
		crc = crc32(crc, reinterpret_cast<const Bytef*>(&hd), (uInt)sizeof(hd));
		const unsigned char* bytes = reinterpret_cast<const unsigned char*>(runningSum.data());
		for (size_t off = 0, total = count * sizeof(float); off < total;) { const size_t n = total - off < (1u << 30) ? total - off : (1u << 30); crc = crc32(crc, bytes + off, (uInt)n); off += n; }
		if ((uint32_t)crc != stored) { err = std::string("checkpoint CRC mismatch: ") + path; return SAILOR_PT_ERR_FORMAT; }
		return SAILOR_PT_OK;
	}
}
