// jpeg_codec.cpp — baseline and progressive JPEG to RGBA8, byte for byte what the reference's image loader produces (host code).
//
// The reference decodes every texture with stbi_load(..., STBI_rgb_alpha) (reference Runtime/Raytracing/MaterialUtils.h:226-249,
// External/stb/stb_image.h v2.27), and a texel is an INPUT of the hot path: a decoder that rounds differently changes base colours
// by an LSB and the parity of every textured hit with it.  Huffman decoding and the progressive refinement passes are fixed by
// ITU-T T.81; what is NOT fixed by the standard, and is therefore restated here from stb_image's published algorithm, is
//   * the integer inverse DCT (jidctint-style, 12-bit constants, +2 extra bits after the column pass, one rounding at >> 17),
//   * chroma upsampling: 3:1 "triangle" filters for 2x1, 1x2 and 2x2 (the 2x2 one filters vertically first, in 16ths), nearest for
//     any other factor,
//   * YCbCr -> RGB in 20-bit fixed point with the constants rounded to 12 bits and the Cb term of green masked to its high half
//     (stb does that so that its scalar and SSE2 paths agree),
//   * CMYK / YCCK (Adobe APP14) through the rounded 8x8 multiply.
// tests/test_image_codecs.py compares this decoder with stbi_load_from_memory itself (linked into the oracle) on baseline /
// progressive / 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 / grey / CMYK / restart-interval / odd-size files.
#include "host_scene.h"
#include "../../include/sailor_pt.h"

#include <math.h>
#include <memory>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace spt
{
	namespace
	{
		const uint8_t kZigzag[64 + 15] = {
			0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
			35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
			63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };      // a run that overshoots lands on the last coefficient

		struct Huffman
		{
			bool present = false;
			uint8_t values[256]; int count[17];
			int minCode[18], maxCode[18], firstIndex[18];       // per code length (T.81 F.2.2.3)
			bool Build(const int* counts, const uint8_t* vals, int total)
			{
				memcpy(values, vals, (size_t)total);
				int code = 0, k = 0;
				for (int len = 1; len <= 16; len++)
				{
					count[len] = counts[len - 1];
					firstIndex[len] = k; minCode[len] = code;
					code += count[len]; k += count[len];
					if (count[len] && code > (1 << len)) return false;
					maxCode[len] = code;                         // exclusive
					code <<= 1;
				}
				present = true;
				return true;
			}
		};

		struct Component
		{
			int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dcPred = 0;
			int x = 0, y = 0, w2 = 0, h2 = 0, coeffW = 0, coeffH = 0;
			std::vector<uint8_t> data; std::vector<int16_t> coeff;
		};

		struct Decoder
		{
			const uint8_t* p; const uint8_t* end;
			uint32_t bitBuf = 0; int bitCount = 0; int marker = 0xFF; bool noMore = false;      // marker 0xFF: none pending
			Huffman dc[4], ac[4]; uint16_t dequant[4][64];
			Component comp[4]; int numComp = 0;
			int width = 0, height = 0, hMax = 1, vMax = 1, mcuW = 0, mcuH = 0, mcuX = 0, mcuY = 0;
			bool progressive = false, jfif = false; int adobeTransform = -1, rgbIds = 0;
			int specStart = 0, specEnd = 63, succHigh = 0, succLow = 0, eobRun = 0;
			int scanN = 0, order[4]; int restartInterval = 0, todo = 0;
			std::string err;

			int Get8() { return p < end ? *p++ : 0; }
			int Get16() { const int a = Get8(); return (a << 8) | Get8(); }
			bool Fail(const char* what) { if (err.empty()) err = what; return false; }

			// entropy-coded segment: 0xFF00 is a data byte 0xFF, any other 0xFFxx ends the segment (the rest reads as zero bits)
			void Fill()
			{
				do
				{
					const unsigned b = noMore ? 0u : (unsigned)Get8();
					if (b == 0xFFu)
					{
						int c = Get8();
						while (c == 0xFF) c = Get8();
						if (c != 0) { marker = c; noMore = true; return; }
					}
					bitBuf |= b << (24 - bitCount);
					bitCount += 8;
				} while (bitCount <= 24);
			}
			int GetBits(int n)
			{
				if (n == 0) return 0;
				if (bitCount < n) Fill();
				const int v = (int)(bitBuf >> (32 - n));
				bitBuf <<= n; bitCount -= n;
				return v;
			}
			int GetBit() { return GetBits(1); }
			// returns the symbol or -1
			int DecodeSymbol(const Huffman& h)
			{
				if (bitCount < 16) Fill();
				int code = 0;
				for (int len = 1; len <= 16; len++)
				{
					code = (int)(bitBuf >> (32 - len));
					if (code < h.maxCode[len] && code >= h.minCode[len] && h.count[len])
					{
						if (len > bitCount) return -1;
						bitBuf <<= len; bitCount -= len;
						return h.values[h.firstIndex[len] + code - h.minCode[len]];
					}
				}
				return -1;
			}
			// T.81 F.2.2.1 EXTEND
			int Receive(int n)
			{
				if (n == 0) return 0;
				const int v = GetBits(n);
				return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
			}
			void ResetEntropy()
			{
				bitBuf = 0; bitCount = 0; noMore = false; marker = 0xFF; eobRun = 0;
				for (int i = 0; i < 4; i++) comp[i].dcPred = 0;
				todo = restartInterval ? restartInterval : 0x7fffffff;
			}

			// ---- blocks ----
			bool BlockBaseline(int16_t* data, Component& c)
			{
				const Huffman& hdc = dc[c.hd]; const Huffman& hac = ac[c.ha]; const uint16_t* dq = dequant[c.tq];
				const int t = DecodeSymbol(hdc);
				if (t < 0 || t > 15) return Fail("bad huffman code");
				memset(data, 0, 64 * sizeof(int16_t));
				const int diff = t ? Receive(t) : 0;
				const int dcv = c.dcPred + diff;
				c.dcPred = dcv;
				data[0] = (int16_t)(dcv * dq[0]);
				int k = 1;
				do
				{
					const int rs = DecodeSymbol(hac);
					if (rs < 0) return Fail("bad huffman code");
					const int s = rs & 15, r = rs >> 4;
					if (s == 0)
					{
						if (rs != 0xF0) break;                    // end of block
						k += 16;
					}
					else
					{
						k += r;
						const int zig = kZigzag[k++];
						data[zig] = (int16_t)(Receive(s) * dq[zig]);
					}
				} while (k < 64);
				return true;
			}
			bool BlockProgDc(int16_t* data, Component& c)
			{
				if (specEnd != 0) return Fail("can't merge dc and ac");
				if (succHigh == 0)
				{
					memset(data, 0, 64 * sizeof(int16_t));
					const int t = DecodeSymbol(dc[c.hd]);
					if (t < 0 || t > 15) return Fail("can't merge dc and ac");
					const int diff = t ? Receive(t) : 0;
					const int dcv = c.dcPred + diff;
					c.dcPred = dcv;
					data[0] = (int16_t)(dcv * (1 << succLow));
				}
				else if (GetBit()) data[0] = (int16_t)(data[0] + (1 << succLow));
				return true;
			}
			bool BlockProgAc(int16_t* data, const Huffman& hac)
			{
				if (specStart == 0) return Fail("can't merge dc and ac");
				if (succHigh == 0)
				{
					const int shift = succLow;
					if (eobRun) { --eobRun; return true; }
					int k = specStart;
					do
					{
						const int rs = DecodeSymbol(hac);
						if (rs < 0) return Fail("bad huffman code");
						const int s = rs & 15, r = rs >> 4;
						if (s == 0)
						{
							if (r < 15)
							{
								eobRun = 1 << r;
								if (r) eobRun += GetBits(r);
								--eobRun;
								break;
							}
							k += 16;
						}
						else
						{
							k += r;
							const int zig = kZigzag[k++];
							data[zig] = (int16_t)(Receive(s) * (1 << shift));
						}
					} while (k <= specEnd);
				}
				else
				{
					const int16_t bit = (int16_t)(1 << succLow);
					auto refine = [&](int16_t* q) { if (GetBit() && (*q & bit) == 0) *q = (int16_t)(*q > 0 ? *q + bit : *q - bit); };
					if (eobRun)
					{
						--eobRun;
						for (int k = specStart; k <= specEnd; k++) { int16_t* q = &data[kZigzag[k]]; if (*q != 0) refine(q); }
					}
					else
					{
						int k = specStart;
						do
						{
							const int rs = DecodeSymbol(hac);
							if (rs < 0) return Fail("bad huffman code");
							int s = rs & 15, r = rs >> 4;
							if (s == 0)
							{
								if (r < 15)
								{
									eobRun = (1 << r) - 1;
									if (r) eobRun += GetBits(r);
									r = 64;                       // refine the rest of the block, place nothing
								}
							}
							else
							{
								if (s != 1) return Fail("bad huffman code");
								s = GetBit() ? bit : -bit;
							}
							while (k <= specEnd)
							{
								int16_t* q = &data[kZigzag[k++]];
								if (*q != 0) refine(q);
								else
								{
									if (r == 0) { *q = (int16_t)s; break; }
									--r;
								}
							}
						} while (k <= specEnd);
					}
				}
				return true;
			}

			// ---- inverse DCT (see the header) ----
			static uint8_t Clamp(int x) { return (unsigned)x > 255u ? (x < 0 ? 0 : 255) : (uint8_t)x; }
			static int F2F(double x) { return (int)(x * 4096 + 0.5); }
			struct Idct1D { int x0, x1, x2, x3, t0, t1, t2, t3; };
			static Idct1D Pass(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7)
			{
				static const int c0541 = F2F(0.5411961f), cm1847 = F2F(-1.847759065f), c0765 = F2F(0.765366865f), c1175 = F2F(1.175875602f),
					c0298 = F2F(0.298631336f), c2053 = F2F(2.053119869f), c3072 = F2F(3.072711026f), c1501 = F2F(1.501321110f),
					cm0899 = F2F(-0.899976223f), cm2562 = F2F(-2.562915447f), cm1961 = F2F(-1.961570560f), cm0390 = F2F(-0.390180644f);
				Idct1D o;
				int p2 = s2, p3 = s6;
				int p1 = (p2 + p3) * c0541;
				int t2 = p1 + p3 * cm1847, t3 = p1 + p2 * c0765;
				p2 = s0; p3 = s4;
				int t0 = (p2 + p3) * 4096, t1 = (p2 - p3) * 4096;
				o.x0 = t0 + t3; o.x3 = t0 - t3; o.x1 = t1 + t2; o.x2 = t1 - t2;
				t0 = s7; t1 = s5; t2 = s3; t3 = s1;
				p3 = t0 + t2; int p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;
				const int p5 = (p3 + p4) * c1175;
				t0 = t0 * c0298; t1 = t1 * c2053; t2 = t2 * c3072; t3 = t3 * c1501;
				p1 = p5 + p1 * cm0899; p2 = p5 + p2 * cm2562; p3 = p3 * cm1961; p4 = p4 * cm0390;
				o.t3 = t3 + p1 + p4; o.t2 = t2 + p2 + p3; o.t1 = t1 + p2 + p4; o.t0 = t0 + p1 + p3;
				return o;
			}
			static void IdctBlock(uint8_t* out, int stride, const int16_t* d)
			{
				int val[64];
				for (int i = 0; i < 8; i++)
				{
					const int16_t* c = d + i; int* v = val + i;
					if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0)
					{
						const int dcterm = c[0] * 4;
						v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
					}
					else
					{
						Idct1D r = Pass(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
						r.x0 += 512; r.x1 += 512; r.x2 += 512; r.x3 += 512;
						v[0] = (r.x0 + r.t3) >> 10; v[56] = (r.x0 - r.t3) >> 10; v[8] = (r.x1 + r.t2) >> 10; v[48] = (r.x1 - r.t2) >> 10;
						v[16] = (r.x2 + r.t1) >> 10; v[40] = (r.x2 - r.t1) >> 10; v[24] = (r.x3 + r.t0) >> 10; v[32] = (r.x3 - r.t0) >> 10;
					}
				}
				for (int i = 0; i < 8; i++)
				{
					const int* v = val + i * 8; uint8_t* o = out + (size_t)i * stride;
					Idct1D r = Pass(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
					const int bias = 65536 + (128 << 17);
					r.x0 += bias; r.x1 += bias; r.x2 += bias; r.x3 += bias;
					o[0] = Clamp((r.x0 + r.t3) >> 17); o[7] = Clamp((r.x0 - r.t3) >> 17); o[1] = Clamp((r.x1 + r.t2) >> 17); o[6] = Clamp((r.x1 - r.t2) >> 17);
					o[2] = Clamp((r.x2 + r.t1) >> 17); o[5] = Clamp((r.x2 - r.t1) >> 17); o[3] = Clamp((r.x3 + r.t0) >> 17); o[4] = Clamp((r.x3 - r.t0) >> 17);
				}
			}

			// ---- markers ----
			bool ProcessMarker(int m)
			{
				switch (m)
				{
				case 0xFF: return Fail("expected marker");
				case 0xDD:
					if (Get16() != 4) return Fail("bad DRI len");
					restartInterval = Get16();
					return true;
				case 0xDB:
				{
					int L = Get16() - 2;
					while (L > 0)
					{
						const int q = Get8(), prec = q >> 4, t = q & 15;
						if (prec != 0 && prec != 1) return Fail("bad DQT type");
						if (t > 3) return Fail("bad DQT table");
						for (int i = 0; i < 64; i++) dequant[t][kZigzag[i]] = (uint16_t)(prec ? Get16() : Get8());
						L -= prec ? 129 : 65;
					}
					return L == 0;
				}
				case 0xC4:
				{
					int L = Get16() - 2;
					while (L > 0)
					{
						int sizes[16], n = 0; uint8_t vals[256];
						const int q = Get8(), tc = q >> 4, th = q & 15;
						if (tc > 1 || th > 3) return Fail("bad DHT header");
						for (int i = 0; i < 16; i++) { sizes[i] = Get8(); n += sizes[i]; }
						if (n > 256) return Fail("bad DHT header");
						L -= 17;
						for (int i = 0; i < n; i++) vals[i] = (uint8_t)Get8();
						if (!(tc == 0 ? dc[th] : ac[th]).Build(sizes, vals, n)) return Fail("bad code lengths");
						L -= n;
					}
					return L == 0;
				}
				default: break;
				}
				if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE)
				{
					int L = Get16();
					if (L < 2) return Fail(m == 0xFE ? "bad COM len" : "bad APP len");
					L -= 2;
					if (m == 0xE0 && L >= 5)
					{
						static const uint8_t tag[5] = { 'J', 'F', 'I', 'F', 0 };
						bool ok = true;
						for (int i = 0; i < 5; i++) if (Get8() != tag[i]) ok = false;
						L -= 5;
						if (ok) jfif = true;
					}
					else if (m == 0xEE && L >= 12)
					{
						static const uint8_t tag[6] = { 'A', 'd', 'o', 'b', 'e', 0 };
						bool ok = true;
						for (int i = 0; i < 6; i++) if (Get8() != tag[i]) ok = false;
						L -= 6;
						if (ok) { Get8(); Get16(); Get16(); adobeTransform = Get8(); L -= 6; }
					}
					p = (end - p) < L ? end : p + L;
					return true;
				}
				return Fail("unknown marker");
			}
			int NextMarker()
			{
				if (marker != 0xFF) { const int m = marker; marker = 0xFF; return m; }
				int x = Get8();
				if (x != 0xFF) return 0xFF;
				while (x == 0xFF) x = Get8();
				return x;
			}
			bool FrameHeader()
			{
				const int Lf = Get16();
				if (Lf < 11) return Fail("bad SOF len");
				if (Get8() != 8) return Fail("only 8-bit");
				height = Get16(); width = Get16();
				if (height == 0) return Fail("no header height");
				if (width == 0) return Fail("0 width");
				if (width > (1 << 24) || height > (1 << 24)) return Fail("too large");
				numComp = Get8();
				if (numComp != 3 && numComp != 1 && numComp != 4) return Fail("bad component count");
				if (Lf != 8 + 3 * numComp) return Fail("bad SOF len");
				rgbIds = 0;
				static const uint8_t rgb[3] = { 'R', 'G', 'B' };
				for (int i = 0; i < numComp; i++)
				{
					Component& c = comp[i];
					c.id = Get8();
					if (numComp == 3 && c.id == rgb[i]) rgbIds++;
					const int q = Get8();
					c.h = q >> 4; c.v = q & 15;
					if (!c.h || c.h > 4) return Fail("bad H");
					if (!c.v || c.v > 4) return Fail("bad V");
					c.tq = Get8();
					if (c.tq > 3) return Fail("bad TQ");
				}
				hMax = vMax = 1;
				for (int i = 0; i < numComp; i++) { if (comp[i].h > hMax) hMax = comp[i].h; if (comp[i].v > vMax) vMax = comp[i].v; }
				for (int i = 0; i < numComp; i++) { if (hMax % comp[i].h != 0) return Fail("bad H"); if (vMax % comp[i].v != 0) return Fail("bad V"); }
				if ((uint64_t)width * height > (1ull << 28)) return Fail("too large");
				mcuW = hMax * 8; mcuH = vMax * 8;
				mcuX = (width + mcuW - 1) / mcuW; mcuY = (height + mcuH - 1) / mcuH;
				for (int i = 0; i < numComp; i++)
				{
					Component& c = comp[i];
					c.x = (width * c.h + hMax - 1) / hMax; c.y = (height * c.v + vMax - 1) / vMax;
					c.w2 = mcuX * c.h * 8; c.h2 = mcuY * c.v * 8;
					c.data.assign((size_t)c.w2 * c.h2, 0);
					if (progressive) { c.coeffW = c.w2 / 8; c.coeffH = c.h2 / 8; c.coeff.assign((size_t)c.w2 * c.h2, 0); }
				}
				return true;
			}
			bool ScanHeader()
			{
				const int Ls = Get16();
				scanN = Get8();
				if (scanN < 1 || scanN > 4 || scanN > numComp) return Fail("bad SOS component count");
				if (Ls != 6 + 2 * scanN) return Fail("bad SOS len");
				for (int i = 0; i < scanN; i++)
				{
					const int id = Get8(), q = Get8();
					int which = 0;
					for (; which < numComp; which++) if (comp[which].id == id) break;
					if (which == numComp) return false;
					comp[which].hd = q >> 4; if (comp[which].hd > 3) return Fail("bad DC huff");
					comp[which].ha = q & 15; if (comp[which].ha > 3) return Fail("bad AC huff");
					order[i] = which;
				}
				specStart = Get8(); specEnd = Get8();
				const int aa = Get8();
				succHigh = aa >> 4; succLow = aa & 15;
				if (progressive)
				{
					if (specStart > 63 || specEnd > 63 || specStart > specEnd || succHigh > 13 || succLow > 13) return Fail("bad SOS");
				}
				else
				{
					if (specStart != 0) return Fail("bad SOS");
					if (succHigh != 0 || succLow != 0) return Fail("bad SOS");
					specEnd = 63;
				}
				return true;
			}
			// after every restart interval: the next marker must be RSTn; anything else ends the scan
			bool RestartOrStop(bool& stop)
			{
				stop = false;
				if (--todo <= 0)
				{
					if (bitCount < 24) Fill();
					if (!(marker >= 0xD0 && marker <= 0xD7)) { stop = true; return true; }
					ResetEntropy();
				}
				return true;
			}
			bool ParseScan()
			{
				ResetEntropy();
				bool stop = false;
				if (!progressive)
				{
					if (scanN == 1)
					{
						Component& c = comp[order[0]];
						const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
						int16_t data[64];
						for (int j = 0; j < h; j++) for (int i = 0; i < w; i++)
						{
							if (!dc[c.hd].present || !ac[c.ha].present) return Fail("missing huffman table");
							if (!BlockBaseline(data, c)) return false;
							IdctBlock(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, data);
							RestartOrStop(stop); if (stop) return true;
						}
						return true;
					}
					int16_t data[64];
					for (int j = 0; j < mcuY; j++) for (int i = 0; i < mcuX; i++)
					{
						for (int k = 0; k < scanN; k++)
						{
							Component& c = comp[order[k]];
							if (!dc[c.hd].present || !ac[c.ha].present) return Fail("missing huffman table");
							for (int y = 0; y < c.v; y++) for (int x = 0; x < c.h; x++)
							{
								const int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
								if (!BlockBaseline(data, c)) return false;
								IdctBlock(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, data);
							}
						}
						RestartOrStop(stop); if (stop) return true;
					}
					return true;
				}
				if (scanN == 1)
				{
					Component& c = comp[order[0]];
					const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
					for (int j = 0; j < h; j++) for (int i = 0; i < w; i++)
					{
						int16_t* data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeffW);
						if (specStart == 0) { if (!dc[c.hd].present) return Fail("missing huffman table"); if (!BlockProgDc(data, c)) return false; }
						else { if (!ac[c.ha].present) return Fail("missing huffman table"); if (!BlockProgAc(data, ac[c.ha])) return false; }
						RestartOrStop(stop); if (stop) return true;
					}
					return true;
				}
				for (int j = 0; j < mcuY; j++) for (int i = 0; i < mcuX; i++)
				{
					for (int k = 0; k < scanN; k++)
					{
						Component& c = comp[order[k]];
						if (!dc[c.hd].present) return Fail("missing huffman table");
						for (int y = 0; y < c.v; y++) for (int x = 0; x < c.h; x++)
						{
							const int x2 = i * c.h + x, y2 = j * c.v + y;
							if (!BlockProgDc(c.coeff.data() + 64 * ((size_t)x2 + (size_t)y2 * c.coeffW), c)) return false;
						}
					}
					RestartOrStop(stop); if (stop) return true;
				}
				return true;
			}
			void FinishProgressive()
			{
				for (int n = 0; n < numComp; n++)
				{
					Component& c = comp[n];
					const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
					for (int j = 0; j < h; j++) for (int i = 0; i < w; i++)
					{
						int16_t* data = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeffW);
						for (int k = 0; k < 64; k++) data[k] = (int16_t)(data[k] * dequant[c.tq][k]);
						IdctBlock(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, data);
					}
				}
			}
			bool DecodeImage()
			{
				int m = NextMarker();
				if (m != 0xD8) return Fail("no SOI");
				m = NextMarker();
				while (!(m == 0xC0 || m == 0xC1 || m == 0xC2))
				{
					if (!ProcessMarker(m)) return false;
					m = NextMarker();
					while (m == 0xFF) { if (p >= end) return Fail("no SOF"); m = NextMarker(); }
				}
				progressive = m == 0xC2;
				if (!FrameHeader()) return false;
				m = NextMarker();
				while (m != 0xD9)
				{
					if (m == 0xDA)
					{
						if (!ScanHeader()) return false;
						if (!ParseScan()) return false;
						if (marker == 0xFF)
						{
							// bytes after the entropy-coded data that are no marker (stb_image v2.27: "handle 0s at the end of image data"): the byte
							// after the next 0xFF is taken as the marker, whatever it is
							while (p < end) { const int x = Get8(); if (x == 255) { marker = Get8(); break; } }
						}
					}
					else if (m == 0xDC) { const int Ld = Get16(); const int NL = Get16(); if (Ld != 4) return Fail("bad DNL len"); if (NL != height) return Fail("bad DNL height"); }
					else if (!ProcessMarker(m)) return false;
					m = NextMarker();
				}
				if (progressive) FinishProgressive();
				return true;
			}

			// ---- upsampling ----
			static uint8_t Div4(int x) { return (uint8_t)(x >> 2); }
			static uint8_t Div16(int x) { return (uint8_t)(x >> 4); }
			static const uint8_t* Resample(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs)
			{
				if (hs == 1 && vs == 1) return nearRow;
				if (hs == 1 && vs == 2) { for (int i = 0; i < w; i++) out[i] = Div4(3 * nearRow[i] + farRow[i] + 2); return out; }
				if (hs == 2 && vs == 1)
				{
					const uint8_t* in = nearRow;
					if (w == 1) { out[0] = out[1] = in[0]; return out; }
					out[0] = in[0]; out[1] = Div4(in[0] * 3 + in[1] + 2);
					int i;
					for (i = 1; i < w - 1; i++) { const int n = 3 * in[i] + 2; out[i * 2] = Div4(n + in[i - 1]); out[i * 2 + 1] = Div4(n + in[i + 1]); }
					out[i * 2] = Div4(in[w - 2] * 3 + in[w - 1] + 2); out[i * 2 + 1] = in[w - 1];
					return out;
				}
				if (hs == 2 && vs == 2)
				{
					if (w == 1) { out[0] = out[1] = Div4(3 * nearRow[0] + farRow[0] + 2); return out; }
					int t1 = 3 * nearRow[0] + farRow[0];
					out[0] = Div4(t1 + 2);
					for (int i = 1; i < w; i++)
					{
						const int t0 = t1;
						t1 = 3 * nearRow[i] + farRow[i];
						out[i * 2 - 1] = Div16(3 * t0 + t1 + 8); out[i * 2] = Div16(3 * t1 + t0 + 8);
					}
					out[w * 2 - 1] = Div4(t1 + 2);
					return out;
				}
				for (int i = 0; i < w; i++) for (int j = 0; j < hs; j++) out[i * hs + j] = nearRow[i];
				return out;
			}
			static int Fixed(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
			static void YCbCrRow(uint8_t* out, const uint8_t* y, const uint8_t* pcb, const uint8_t* pcr, int count)
			{
				static const int kR = Fixed(1.40200f), kGr = Fixed(0.71414f), kGb = Fixed(0.34414f), kB = Fixed(1.77200f);
				for (int i = 0; i < count; i++, out += 4)
				{
					const int yf = (y[i] << 20) + (1 << 19);
					const int cr = pcr[i] - 128, cb = pcb[i] - 128;
					int r = yf + cr * kR;
					int g = yf + (cr * -kGr) + (int)((unsigned)(cb * -kGb) & 0xffff0000u);
					int b = yf + cb * kB;
					r >>= 20; g >>= 20; b >>= 20;
					out[0] = Clamp(r); out[1] = Clamp(g); out[2] = Clamp(b); out[3] = 255;
				}
			}
			static uint8_t Blinn(uint8_t x, uint8_t y) { const unsigned t = (unsigned)x * y + 128u; return (uint8_t)((t + (t >> 8)) >> 8); }

			bool ToRgba(std::vector<uint8_t>& rgba)
			{
				const bool isRgb = numComp == 3 && (rgbIds == 3 || (adobeTransform == 0 && !jfif));
				struct Res { int hs, vs, ystep, wLores, ypos; const uint8_t* line0; const uint8_t* line1; std::vector<uint8_t> buf; } res[4];
				for (int k = 0; k < numComp; k++)
				{
					Res& r = res[k];
					r.hs = hMax / comp[k].h; r.vs = vMax / comp[k].v; r.ystep = r.vs >> 1; r.wLores = (width + r.hs - 1) / r.hs; r.ypos = 0;
					r.line0 = r.line1 = comp[k].data.data(); r.buf.assign((size_t)width + 3 + 8, 0);
				}
				rgba.resize((size_t)width * height * 4);
				const uint8_t* co[4] = { nullptr, nullptr, nullptr, nullptr };
				for (int j = 0; j < height; j++)
				{
					uint8_t* out = rgba.data() + (size_t)4 * width * j;
					for (int k = 0; k < numComp; k++)
					{
						Res& r = res[k];
						const bool bot = r.ystep >= (r.vs >> 1);
						co[k] = Resample(r.buf.data(), bot ? r.line1 : r.line0, bot ? r.line0 : r.line1, r.wLores, r.hs, r.vs);
						if (++r.ystep >= r.vs)
						{
							r.ystep = 0; r.line0 = r.line1;
							if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
						}
					}
					if (numComp == 3)
					{
						if (isRgb) for (int i = 0; i < width; i++, out += 4) { out[0] = co[0][i]; out[1] = co[1][i]; out[2] = co[2][i]; out[3] = 255; }
						else YCbCrRow(out, co[0], co[1], co[2], width);
					}
					else if (numComp == 4)
					{
						if (adobeTransform == 0) for (int i = 0; i < width; i++, out += 4) { const uint8_t m = co[3][i]; out[0] = Blinn(co[0][i], m); out[1] = Blinn(co[1][i], m); out[2] = Blinn(co[2][i], m); out[3] = 255; }
						else if (adobeTransform == 2)
						{
							YCbCrRow(out, co[0], co[1], co[2], width);
							for (int i = 0; i < width; i++, out += 4) { const uint8_t m = co[3][i]; out[0] = Blinn((uint8_t)(255 - out[0]), m); out[1] = Blinn((uint8_t)(255 - out[1]), m); out[2] = Blinn((uint8_t)(255 - out[2]), m); }
						}
						else YCbCrRow(out, co[0], co[1], co[2], width);
					}
					else for (int i = 0; i < width; i++, out += 4) { out[0] = out[1] = out[2] = co[0][i]; out[3] = 255; }
				}
				return true;
			}
		};
	}

	int DecodeImageRgba8(const uint8_t* data, size_t size, int32_t& w, int32_t& h, std::vector<uint8_t>& rgba, std::string& err)
	{
		static const uint8_t pngSig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
		if (size >= 8 && !memcmp(data, pngSig, 8)) return DecodePngRgba8(data, size, w, h, rgba, err);
		if (IsJpeg(data, size)) return DecodeJpegRgba8(data, size, w, h, rgba, err);
		err = "unsupported image format (PNG and JPEG are decoded)";
		return SAILOR_PT_ERR_UNSUPPORTED;
	}

	// ---- Radiance RGBE ------------------------------------------------------------------------------------------------------------
	// stb_image's stbi__hdr_load restated: "#?RADIANCE" / "#?RGBE" header, FORMAT=32-bit_rle_rgbe, "-Y h +X w", new-style RLE scanlines (flat
	// data for widths < 8 or >= 32768), value = mantissa * 2^(e - 136) (one float multiply per channel), alpha 1.
	bool IsHdr(const uint8_t* data, size_t size)
	{
		return (size >= 11 && !memcmp(data, "#?RADIANCE\n", 11)) || (size >= 7 && !memcmp(data, "#?RGBE\n", 7));
	}
	int DecodeHdrRgba32F(const uint8_t* data, size_t size, int32_t& w, int32_t& h, std::vector<float>& out, std::string& err)
	{
		const uint8_t* p = data; const uint8_t* end = data + size;
		auto get8 = [&]() -> int { return p < end ? *p++ : 0; };
		auto token = [&]() -> std::string { std::string t; while (p < end) { const char c = (char)*p++; if (c == '\n') break; if (t.size() < 1023) t.push_back(c); } return t; };
		const std::string magic = token();
		if (magic != "#?RADIANCE" && magic != "#?RGBE") { err = "not HDR"; return SAILOR_PT_ERR_FORMAT; }
		bool valid = false;
		for (;;) { const std::string t = token(); if (t.empty()) break; if (t == "FORMAT=32-bit_rle_rgbe") valid = true; if (p >= end) break; }
		if (!valid) { err = "unsupported HDR format"; return SAILOR_PT_ERR_FORMAT; }
		const std::string dims = token();
		if (dims.compare(0, 3, "-Y ") != 0) { err = "unsupported HDR data layout"; return SAILOR_PT_ERR_FORMAT; }
		char* rest = nullptr;
		const long height = strtol(dims.c_str() + 3, &rest, 10);
		while (*rest == ' ') ++rest;
		if (strncmp(rest, "+X ", 3)) { err = "unsupported HDR data layout"; return SAILOR_PT_ERR_FORMAT; }
		const long width = strtol(rest + 3, nullptr, 10);
		if (width <= 0 || height <= 0 || width > (1 << 24) || height > (1 << 24) || (uint64_t)width * (uint64_t)height > (1ull << 28)) { err = "HDR image too large"; return SAILOR_PT_ERR_FORMAT; }
		out.assign((size_t)width * height * 4, 0.0f);
		auto convert = [](float* o, const uint8_t* in)
			{
				if (in[3] != 0)
				{
					const float f1 = (float)ldexp(1.0f, (int)in[3] - (int)(128 + 8));
					o[0] = in[0] * f1; o[1] = in[1] * f1; o[2] = in[2] * f1; o[3] = 1.0f;
				}
				else { o[0] = o[1] = o[2] = 0.0f; o[3] = 1.0f; }
			};
		auto flat = [&](long j0, long i0)
			{
				for (long j = j0; j < height; j++) for (long i = (j == j0 ? i0 : 0); i < width; i++)
				{
					uint8_t rgbe[4]; for (int k = 0; k < 4; k++) rgbe[k] = (uint8_t)get8();
					convert(out.data() + ((size_t)j * width + i) * 4, rgbe);
				}
			};
		if (width < 8 || width >= 32768) flat(0, 0);
		else
		{
			std::vector<uint8_t> scan((size_t)width * 4);
			for (long j = 0; j < height; j++)
			{
				const int c1 = get8(), c2 = get8(); int len = get8();
				if (c1 != 2 || c2 != 2 || (len & 0x80))
				{
					// not run-length encoded: these four bytes are the first pixel of a flat file
					uint8_t rgbe[4] = { (uint8_t)c1, (uint8_t)c2, (uint8_t)len, (uint8_t)get8() };
					convert(out.data(), rgbe);
					flat(0, 1);
					break;
				}
				len = (len << 8) | get8();
				if (len != width) { err = "corrupt HDR: bad scanline length"; return SAILOR_PT_ERR_FORMAT; }
				for (int k = 0; k < 4; k++)
				{
					long i = 0, nleft;
					while ((nleft = width - i) > 0)
					{
						int count = get8();
						if (count > 128)
						{
							const uint8_t value = (uint8_t)get8();
							count -= 128;
							if (count > nleft) { err = "corrupt HDR: bad RLE data"; return SAILOR_PT_ERR_FORMAT; }
							for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = value;
						}
						else
						{
							if (count > nleft || count == 0) { err = "corrupt HDR: bad RLE data"; return SAILOR_PT_ERR_FORMAT; }
							for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = (uint8_t)get8();
						}
					}
				}
				for (long i = 0; i < width; i++) convert(out.data() + ((size_t)j * width + i) * 4, scan.data() + (size_t)i * 4);
			}
		}
		w = (int32_t)width; h = (int32_t)height;
		return SAILOR_PT_OK;
	}

	bool IsJpeg(const uint8_t* data, size_t size) { return size >= 3 && data[0] == 0xFF && data[1] == 0xD8 && data[2] == 0xFF; }

	int DecodeJpegRgba8(const uint8_t* data, size_t size, int32_t& w, int32_t& h, std::vector<uint8_t>& rgba, std::string& err)
	{
		std::unique_ptr<Decoder> d(new Decoder());
		d->p = data; d->end = data + size;
		memset(d->dequant, 0, sizeof(d->dequant));
		if (!d->DecodeImage()) { err = "JPEG: " + (d->err.empty() ? std::string("corrupt") : d->err); return SAILOR_PT_ERR_FORMAT; }
		if (!d->ToRgba(rgba)) { err = "JPEG: conversion failed"; return SAILOR_PT_ERR_FORMAT; }
		w = d->width; h = d->height;
		return SAILOR_PT_OK;
	}
}
