// integrator.cuh — the path integrator as a WAVEFRONT of resumable paths.
//
// The reference's Raytrace (reference Raytracing/PathTracer.cpp:622-879) is a recursive estimator, not a flat path
// sum: per-level clamps to [0,10] (:733,:788,:816), MIS on per-hit AVERAGES (:838-847), S-way branching at the
// first hit only (:716-717), a deterministic cut-off on the accumulated throughput (:810-811), an alpha-blend
// continuation that is a second recursive call (:858-871), a tail call when leaving a thick volume (:673-688) and
// the TraceSky transmission loop (:577-620).  To keep every one of those non-linearities, each pool slot runs ONE
// primary sample as an explicit state machine over a small stack of call frames (depth <= maxBounces+1, because
// every recursive call decrements bounceLimit and is only made when it is > 0).  A slot has at most one ray in
// flight; one wavefront iteration is
//      trace kernel   : closest hit for every slot's pending ray        (traverse.cuh, persistent threads)
//      advance kernel : resume every slot with its hit, run shading / sampling / accumulation until the slot needs
//                       its next ray (or finishes and pulls the next primary sample from the global counter)
// so all rays of an iteration are traced together and the scene data is only touched by the two kernels.
//
// Deliberate, documented differences from the reference (none changes the estimator):
//  * the reference re-traces the importance ray when it recurses (:786 then :632 with the same ray and ignore
//    index); the child frame here starts from the hit the parent already has;
//  * random numbers: same distributions as glm::linearRand on rand()%255 bytes and the same blue-noise
//    table walk (:934-1077), but drawn from a counter-based generator keyed by (seed, pixel, primary-sample index),
//    so the image does not depend on thread scheduling or on how the frame is split across GPUs;
//  * the unbounded rejection loop (:761-767) is capped at 4096 tries; an exhausted sample is skipped.
#pragma once
#include "pipeline.cuh"
#include "textures.cuh"
#include "lighting.cuh"
#include "blue_noise_table.h"

namespace spt
{
	enum Phase : uint32_t
	{
		kPhMain = 0,        // waiting for the closest hit of the frame's own ray (:632)
		kPhLight,           // waiting for a directional-light shadow ray (:699)
		kPhSkyHemi,         // inside TraceSky for a hemisphere sample (:729)
		kPhSample,          // waiting for the importance-sampled ray (:786)
		kPhSkySample,       // inside TraceSky behind a transmissive hit (:825)
		kPhChildSample,     // child Raytrace of an importance sample is running (:813)
		kPhChildAlpha,      // child Raytrace of the alpha-blend continuation is running (:869)
	};

	enum : uint32_t { kFlOpposite = 1u, kFlThick = 2u, kFlAlpha = 4u, kFlHasTransRay = 8u, kFlTransRay = 16u, kFlFirst = 32u, kFlLightBlocked = 64u };

	// One Raytrace() activation record.
	struct alignas(16) Frame
	{
		// call arguments (:622)
		V3 rayO; uint32_t ignoreTri;
		V3 rayD; uint32_t bounceLimit;
		float inAcc, envIor; uint32_t pMaxBounces, pNumSamples;
		uint32_t pNumAmbient, seedX, seedY, phase;
		// shading context of the hit (:636-671)
		V3 hitPoint; uint32_t hitTri;
		V3 N; uint32_t matIdx;
		V3 V; uint32_t flags;
		V3 offset; uint32_t loopI;
		V4 baseColor;
		V3 orm; float ior;
		V3 emissive; float transmission;
		uint32_t S, A, nA, nS;
		// accumulators (:690, :712, :742-747)
		V3 res; float avgPdf;
		V3 amb1; float cnt;
		V3 amb2; float pdf;
		V3 indirect; float newIor;
		// pending importance sample (:755-781)
		V3 term; uint32_t h2tri;
		V3 att; float toIor;
		V3 r2o; float thickness;
		V3 r2d; float pad0;
		V3 value; float pad1;
		// TraceSky state (:577-620)
		V3 skyAtt; uint32_t skyIgnore;
		V3 skyPrev; uint32_t skyJ;
		V3 skyStart; float skyIor;
		V3 skyDir; float pad2;
		V3 skyToL; float pad3;
	};

	struct alignas(16) PathHeader
	{
		uint32_t depth;       // index of the running frame
		uint32_t active;      // 0 = slot idle (no more primary samples)
		uint32_t pixel;       // y*width + x, task coordinates
		uint32_t sample;      // primary-sample (msaa) index
		uint64_t rngKey;
		uint32_t rngCounter;
		uint32_t pad;
	};

	// One first hit of the primary pass: which (pixel, primary sample) and the closest hit of its camera ray.
	struct alignas(16) PrimaryHitRec { uint32_t pixel, sample, pad0, pad1; float t, u, v; uint32_t tri; };
	static_assert(sizeof(PrimaryHitRec) == 32, "PrimaryHitRec layout");

	struct RenderStats { uint64_t rays, primarySamples; double secondsTraverse, secondsShade; uint32_t traverseLaunches; };

	struct IntegratorArgs
	{
		// scene
		const V4* shade; const V4* centroid; const MaterialGpu* materials; TextureSet tex; const V4* lights; uint32_t numLights;
		const uint16_t* blueNoise;
		// camera / params
		CameraGpu cam; uint32_t rowBegin, rowEnd, msBegin, msEnd, msaa;
		uint32_t maxBounces, numSamples, numAmbientSamples; V3 ambient; uint64_t seed;
		// pool
		uint32_t poolSize; uint32_t maxDepth;
		PathHeader* headers; Frame* frames; RayRec* rays; const Hit* hits;
		float* sampleBuf;                       // 3 floats per (pixel in band, sample in range)
		const PrimaryHitRec* hitQueue; uint32_t queueCount;   // first hits found by the primary pass
		uint32_t* nextSample;                                   // work counter over hitQueue
		uint32_t* activeCount; unsigned long long* rayCount;
	};

	// ---- random numbers --------------------------------------------------------------
	SPT_HD uint64_t Mix64(uint64_t z)
	{
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
	}
	SPT_HD uint64_t PrimaryRngKey(uint64_t seed, uint32_t pixel, uint32_t msaa, uint32_t sample)
	{
		return Mix64(seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)pixel * msaa + sample) + 0x632BE59BD9B4E019ULL);
	}
	struct Rng
	{
		uint64_t key; uint32_t counter;
		// glm compute_rand<uint32> (gtc/random.inl:19-27,66-85): four bytes, each rand() % 255 (never 255)
		SPT_KERNEL_BODY uint32_t U32()
		{
			const uint64_t h = Mix64(key + 0x9E3779B97F4A7C15ULL * (uint64_t)(++counter));
			const uint32_t b0 = (uint32_t)(h & 0xFFFF) % 255u, b1 = (uint32_t)((h >> 16) & 0xFFFF) % 255u;
			const uint32_t b2 = (uint32_t)((h >> 32) & 0xFFFF) % 255u, b3 = (uint32_t)((h >> 48) & 0xFFFF) % 255u;
			return (b3 << 24) | (b2 << 16) | (b1 << 8) | b0;
		}
		// glm::linearRand(0.f, 1.f) (random.inl:176-183): float(u32) / float(UINT32_MAX) * (Max - Min) + Min
		SPT_KERNEL_BODY float Float01() { return (float)U32() / 4294967296.0f * (1.0f - 0.0f) + 0.0f; }
		// glm::linearRand(0, 680) (random.inl:140-147)
		SPT_KERNEL_BODY uint32_t Seed681() { return U32() % 681u; }
	};

	// ---- the state machine --------------------------------------------------------------------------------------
	struct PathMachine
	{
		const IntegratorArgs& a;
		uint32_t slot;
		PathHeader hd;
		Rng rng;
		Frame* f;
		bool rayEmitted;

		SPT_KERNEL_BODY PathMachine(const IntegratorArgs& args, uint32_t s) : a(args), slot(s), f(nullptr), rayEmitted(false) {}

		SPT_KERNEL_BODY Frame* FrameAt(uint32_t depth) const { return a.frames + (size_t)depth * a.poolSize + slot; }

		SPT_KERNEL_BODY void Emit(V3 o, V3 d, uint32_t ignore)
		{
			RayRec r; r.ox = o.x; r.oy = o.y; r.oz = o.z; r.ignoreTri = ignore; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = kFltMax;
			a.rays[slot] = r;
			rayEmitted = true;
		}

		SPT_KERNEL_BODY V3 HitNormal(uint32_t tri, float u, float v) const   // Bounds.cpp:525-527
		{
			const V4* S = a.shade + (size_t)tri * 9;
			const V4 n0 = ld4(S), n1 = ld4(S + 1), n2 = ld4(S + 2);
			const float w = 1.0f - u - v;
			return w * v3(n0.x, n0.y, n0.z) + u * v3(n1.x, n1.y, n1.z) + v * v3(n2.x, n2.y, n2.z);
		}
		SPT_KERNEL_BODY uint32_t MaterialOf(uint32_t tri) const { return f2u(ld4(a.centroid + tri).w); }

		SPT_KERNEL_BODY V2 BlueNoise()                                        // PathTracer.cpp:934-1077
		{
			if (f->seedX >= 688u) f->seedX = rng.Seed681();
			if (f->seedY >= 688u) f->seedY = rng.Seed681();
			const float x = (float)a.blueNoise[f->seedX++] * (1.0f / 1024.0f), y = (float)a.blueNoise[f->seedY++] * (1.0f / 1024.0f);
			return v2(x, y);
		}

		// Entry of Raytrace (:622-632): draw the two table seeds; the ray is traced unless the caller already has the hit.
		SPT_KERNEL_BODY void BeginCall()
		{
			f->seedX = rng.Seed681(); f->seedY = rng.Seed681();
			f->phase = kPhMain;
		}

		SPT_KERNEL_BODY void StartSky(V3 start, V3 toL, float ior, uint32_t ignore)   // TraceSky prologue (:579-580)
		{
			f->skyAtt = v3(1.0f); f->skyPrev = start; f->skyStart = start; f->skyDir = toL; f->skyIor = ior; f->skyIgnore = ignore; f->skyJ = 0;
		}

		// Start the next first-hit record of the primary pass (:444-466): frame 0 is set up exactly as Raytrace would
		// see it and resumed with the hit the primary pass already found.  Returns false when the queue is exhausted.
		SPT_KERNEL_BODY bool NextFromQueue()
		{
			const uint32_t g = atomic_add_u32(a.nextSample, 1u);
			if (g >= a.queueCount) return false;
			const PrimaryHitRec rec = a.hitQueue[g];
			const uint32_t x = rec.pixel % a.cam.width, y = rec.pixel / a.cam.width;
			hd.active = 1; hd.depth = 0; hd.pixel = rec.pixel; hd.sample = rec.sample;
			hd.rngKey = PrimaryRngKey(a.seed, rec.pixel, a.msaa, rec.sample);
			rng.key = hd.rngKey; rng.counter = 0;
			float ox = 0.5f, oy = 0.5f;                                           // :460
			if (rec.sample != 0) { ox = rng.Float01(); oy = rng.Float01(); }
			f = FrameAt(0);
			f->rayO = a.cam.pos; f->rayD = PrimaryDir(a.cam, x, y, ox, oy);
			f->ignoreTri = kNoHit; f->bounceLimit = a.maxBounces; f->inAcc = 1.0f; f->envIor = 1.0f;
			f->pMaxBounces = a.maxBounces; f->pNumSamples = a.numSamples; f->pNumAmbient = a.numAmbientSamples;
			BeginCall();
			Hit h; h.t = rec.t; h.u = rec.u; h.v = rec.v; h.tri = rec.tri;
			Advance(h);
			return true;
		}

		SPT_KERNEL_BODY SampledData Sampled() const
		{
			SampledData s; s.baseColor = f->baseColor; s.orm = f->orm; s.emissive = f->emissive; s.normal = v3(0.0f, 0.0f, 1.0f);
			s.ior = f->ior; s.thickness = f->thickness; s.transmission = f->transmission; s.opaque = true;
			return s;
		}

		// Runs until a ray has been emitted or the slot has no more work.  `hit` is the result of the pending ray.
		SPT_KERNEL_BODY void Advance(Hit hit)
		{
			enum Act { OnResult, Shade, LightsNext, AmbientBegin, HemiNext, SkyStep, SkyDone, SamplesBegin, SampleNext, AfterChildSample, AmbientEnd, Finish, AfterChildAlpha, Return, Done };
			Act act = OnResult;
			V3 retVal = v3(0.0f);      // value being returned by a finished call
			V3 skyResult = v3(0.0f);
			for (;;)
			{
				switch (act)
				{
				case OnResult:
				{
					switch (f->phase)
					{
					case kPhMain:
						if (hit.tri == kNoHit) { retVal = a.ambient; act = Return; }   // :873-876
						else act = Shade;
						break;
					case kPhLight:
					{
						if (hit.tri == kNoHit)                                          // :699-703
						{
							const V4 ld = a.lights[f->loopI * 2], li = a.lights[f->loopI * 2 + 1];
							const V3 toL = -v3(ld.x, ld.y, ld.z);
							const float angle = glm_max(0.0f, dot(toL, f->N));
							f->res = f->res + CalculateBRDF(f->V, f->N, toL, Sampled()) * v3(li.x, li.y, li.z) * angle;
						}
						else f->flags |= kFlLightBlocked;
						f->loopI++;
						act = LightsNext;
						break;
					}
					case kPhSkyHemi:
					case kPhSkySample:
					{
						// TraceSky loop body (:585-616)
						if (hit.tri == kNoHit) { skyResult = f->skyAtt; act = SkyDone; break; }
						const MaterialGpu& m = a.materials[MaterialOf(hit.tri)];
						const V3 hn = HitNormal(hit.tri, hit.u, hit.v);
						const bool hitOpp = dot(f->skyDir, hn) < 0.0f;
						if (!(m.transmission > 0.0f && m.thickness > 0.0f)) { skyResult = v3(0.0f); act = SkyDone; break; }
						const V3 hp = f->skyStart + f->skyDir * hit.t;
						const float distance = length(hp - f->skyPrev);
						f->skyPrev = hp;
						const V3 wn = hitOpp ? hn : -hn;
						const float toIor = hitOpp ? m.ior : 1.0f;
						f->skyDir = CalculateRefraction(f->skyDir, wn, f->skyIor, toIor);
						f->skyIor = toIor;
						if (!hitOpp)
						{
							const V3 c = v3(-logf(m.attenuationColor[0]), -logf(m.attenuationColor[1]), -logf(m.attenuationColor[2])) / m.attenuationDistance;
							const V3 e = -c * distance;
							f->skyAtt = f->skyAtt * v3(expf(e.x), expf(e.y), expf(e.z));
						}
						f->skyStart = hp; f->skyIgnore = hit.tri; f->skyJ++;
						act = SkyStep;
						break;
					}
					case kPhSample:
					{
						const V3 term = f->term;
						if (hit.tri == kNoHit)                                          // :786-796
						{
							const V3 value = glm_clamp(term * a.ambient, 0.0f, 10.0f);
							f->amb2 = f->amb2 + value; f->avgPdf += f->pdf; f->indirect = f->indirect + value;
							f->cnt += 1.0f; f->loopI++;
							act = SampleNext;
						}
						else if (f->bounceLimit > 0)                                    // :797-833
						{
							V3 att = v3(1.0f);
							const V3 h2p = f->r2o + f->r2d * hit.t;
							if ((f->flags & kFlOpposite) && (f->flags & kFlTransRay) && (f->flags & kFlThick))
							{
								const MaterialGpu& m = a.materials[f->matIdx];
								const float distance = length(h2p - f->hitPoint);
								const V3 c = v3(-logf(m.attenuationColor[0]), -logf(m.attenuationColor[1]), -logf(m.attenuationColor[2])) / m.attenuationDistance;
								const V3 e = -c * distance;
								att = v3(expf(e.x), expf(e.y), expf(e.z));
							}
							f->att = att; f->h2tri = hit.tri;
							const float newAcc = f->inAcc * length(term * att) * f->baseColor.w;
							if (newAcc > 0.01f)                                         // :810-814
							{
								Frame* parent = f;
								parent->phase = kPhChildSample;
								hd.depth++;
								f = FrameAt(hd.depth);
								f->rayO = parent->r2o; f->rayD = parent->r2d; f->ignoreTri = parent->hitTri; f->bounceLimit = parent->bounceLimit - 1;
								f->inAcc = newAcc; f->envIor = parent->newIor;
								f->pMaxBounces = parent->pMaxBounces; f->pNumSamples = parent->pNumSamples; f->pNumAmbient = parent->pNumAmbient;
								BeginCall();
								act = Shade;   // the child's own IntersectBVH (:632) would return exactly `hit`
							}
							else { retVal = v3(0.0f); act = AfterChildSample; }
						}
						else { f->cnt += 1.0f; f->loopI++; act = SampleNext; }        // hit, but no bounces left (:835)
						break;
					}
					default: act = Done; break;   // unreachable
					}
					break;
				}
				case Shade:
				{
					// :636-688
					const uint32_t tri = hit.tri;
					const V4* S = a.shade + (size_t)tri * 9;
					const V4 s0 = ld4(S), s1 = ld4(S + 1), s2 = ld4(S + 2), s3 = ld4(S + 3), s4 = ld4(S + 4), s5 = ld4(S + 5), s6 = ld4(S + 6), s7 = ld4(S + 7), s8 = ld4(S + 8);
					const float bu = hit.u, bv = hit.v, bw = 1.0f - bu - bv;
					V3 faceNormal = bw * v3(s0.x, s0.y, s0.z) + bu * v3(s1.x, s1.y, s1.z) + bv * v3(s2.x, s2.y, s2.z);
					const V3 tangent = bw * v3(s3.x, s3.y, s3.z) + bu * v3(s4.x, s4.y, s4.z) + bv * v3(s5.x, s5.y, s5.z);
					const V3 bitangent = bw * v3(s6.x, s6.y, s6.z) + bu * v3(s7.x, s7.y, s7.z) + bv * v3(s8.x, s8.y, s8.z);
					const bool opposite = dot(faceNormal, f->rayD) < 0.0f;
					if (!opposite) faceNormal = faceNormal * -1.0f;
					const V2 uv = bw * v2(s0.w, s1.w) + bu * v2(s2.w, s3.w) + bv * v2(s4.w, s5.w);
					const uint32_t matIdx = f2u(s6.w);
					const MaterialGpu& m = a.materials[matIdx];
					const float* T = m.uvTransform;
					const float tu = T[0] * uv.x + T[4] * uv.y + T[8] * 1.0f, tv = T[1] * uv.x + T[5] * uv.y + T[9] * 1.0f;
					// GetMaterialData (:881-927)
					V4 baseColor = v4(m.baseColor[0], m.baseColor[1], m.baseColor[2], m.baseColor[3]);
					V3 nrm = v3(0.0f, 0.0f, 1.0f);
					V3 orm = v3(0.0f, m.roughness, m.metallic);
					V3 emissive = v3(m.emissive[0], m.emissive[1], m.emissive[2]);
					float transmission = m.transmission;
					if (m.texBase != kNoTexture) { const V4 t = SampleTexture(a.tex, m.texBase, tu, tv); baseColor = v4(baseColor.x * t.x, baseColor.y * t.y, baseColor.z * t.z, baseColor.w * t.w); }
					if (m.texEmissive != kNoTexture) { const V4 t = SampleTexture(a.tex, m.texEmissive, tu, tv); emissive = emissive * v3(t.x, t.y, t.z); }
					if (m.texMetallicRoughness != kNoTexture) { const V4 t = SampleTexture(a.tex, m.texMetallicRoughness, tu, tv); orm = v3(t.x, orm.y * t.y, orm.z * t.z); }
					if (m.texNormal != kNoTexture) { const V4 t = SampleTexture(a.tex, m.texNormal, tu, tv); nrm = v3(t.x, t.y, t.z); }
					if (m.texTransmission != kNoTexture) { const V4 t = SampleTexture(a.tex, m.texTransmission, tu, tv); transmission *= t.x; }
					if (m.blendMode == kMask) baseColor.w = (baseColor.w > m.alphaCutoff) ? 1.0f : 0.0f;
					const bool opaque = m.blendMode == kOpaque;

					const V3 V = -normalize(f->rayD);
					// tbn * normal, glm mat3 * vec3 order (type_mat3x3.inl:468-474)
					const V3 N = normalize(v3(tangent.x * nrm.x + bitangent.x * nrm.y + faceNormal.x * nrm.z,
						tangent.y * nrm.x + bitangent.y * nrm.y + faceNormal.y * nrm.z,
						tangent.z * nrm.x + bitangent.z * nrm.y + faceNormal.z * nrm.z));
					const bool alphaBlend = !opaque && baseColor.w < 1.0f;
					const uint32_t sRound = (uint32_t)roundf(baseColor.w * (float)f->pNumSamples), aRound = (uint32_t)roundf(baseColor.w * (float)f->pNumAmbient);
					const uint32_t numSamples = alphaBlend ? (sRound > 1u ? sRound : 1u) : f->pNumSamples;
					const uint32_t numAmbient = alphaBlend ? (aRound > 1u ? aRound : 1u) : f->pNumAmbient;
					const V3 offset = 0.000001f * faceNormal;
					const bool fullMetal = orm.z == 1.0f;
					const bool hasTrans = !fullMetal && transmission > 0.0f;
					const bool thick = hasTrans && m.thickness > 0.0f;
					const V3 hitPoint = f->rayO + f->rayD * hit.t;

					if (!opposite && thick)                                             // :673-688 (tail call)
					{
						const V3 nd = CalculateRefraction(f->rayD, N, f->envIor, 1.0f);
						if (eq0(nd) || f->bounceLimit == 0) { retVal = v3(0.0f); act = Return; break; }
						f->rayO = hitPoint; f->rayD = nd - offset; f->bounceLimit -= 1; f->ignoreTri = tri; f->envIor = 1.0f;
						BeginCall();
						Emit(f->rayO, f->rayD, f->ignoreTri);
						act = Done;
						break;
					}
					const bool first = f->bounceLimit == f->pMaxBounces;                // :636
					f->hitPoint = hitPoint; f->hitTri = tri; f->N = N; f->matIdx = matIdx; f->V = V; f->offset = offset;
					f->flags = (opposite ? kFlOpposite : 0u) | (thick ? kFlThick : 0u) | (alphaBlend ? kFlAlpha : 0u) | (first ? kFlFirst : 0u);
					f->baseColor = baseColor; f->orm = orm; f->ior = m.ior; f->emissive = emissive; f->transmission = transmission; f->thickness = m.thickness;
					f->S = numSamples; f->A = numAmbient;
					f->res = v3(0.0f); f->loopI = 0;
					act = LightsNext;
					break;
				}
				case LightsNext:                                                        // :691-705
				{
					// QUIRK kept: `RaycastHit hitLight{}` is declared outside the light loop (:694) and IntersectBVH only
					// writes it on a hit, so after the first shadowed light every later light reads the stale hit and is
					// treated as shadowed too.  Their rays cannot change the result and are not traced.
					if (f->loopI < a.numLights && !(f->flags & kFlLightBlocked))
					{
						const V4 ld = a.lights[f->loopI * 2];
						f->phase = kPhLight;
						Emit(f->hitPoint + f->offset, -v3(ld.x, ld.y, ld.z), f->hitTri);
						act = Done;
					}
					else act = AmbientBegin;
					break;
				}
				case AmbientBegin:                                                      // :708-717
				{
					if (!(a.ambient.x + a.ambient.y + a.ambient.z > 0.0f)) { act = Finish; break; }
					const bool first = (f->flags & kFlFirst) != 0;
					f->nA = first ? f->A : 1u; f->nS = first ? f->S : 1u;
					f->amb1 = v3(0.0f); f->loopI = 0;
					act = HemiNext;
					break;
				}
				case HemiNext:                                                          // :720-739
				{
					if (!(f->flags & kFlThick) && f->loopI < f->nA)
					{
						const float r0 = rng.Float01(), r1 = rng.Float01();               // NextVec2_Linear (:929-932)
						const V3 H = ImportanceSampleHemisphere(v2(r0, r1), f->N);
						const V3 toL = 2.0f * dot(f->V, H) * H - f->V;
						f->skyToL = toL;
						StartSky(f->hitPoint + f->offset, toL, f->envIor, f->hitTri);
						f->phase = kPhSkyHemi;
						act = SkyStep;
					}
					else { f->amb1 = f->amb1 / (float)f->nA; act = SamplesBegin; }
					break;
				}
				case SkyStep:                                                           // for (j < maxBounces) (:581-590)
				{
					if (f->skyJ < f->pMaxBounces) { Emit(f->skyStart, f->skyDir, f->skyIgnore); act = Done; }
					else { skyResult = v3(0.0f); act = SkyDone; }
					break;
				}
				case SkyDone:
				{
					const V3 att = skyResult;
					const bool nonZero = att.x != 0.0f || att.y != 0.0f || att.z != 0.0f;
					if (f->phase == kPhSkyHemi)                                         // :730-735
					{
						if (nonZero)
						{
							const float pdfHemisphere = 1.0f / (kPiSailor * 2.0f);
							const float angle = glm_max(0.0f, dot(f->skyToL, f->N));
							const V3 at = att * CalculateBRDF(f->V, f->N, f->skyToL, Sampled());
							f->amb1 = f->amb1 + glm_clamp((at * a.ambient * angle) / pdfHemisphere, 0.0f, 10.0f);
						}
						f->loopI++;
						act = HemiNext;
					}
					else                                                                // :827-831
					{
						if (nonZero) { f->amb2 = f->amb2 + f->value * att; f->avgPdf += f->pdf; }
						f->cnt += 1.0f; f->loopI++;
						act = SampleNext;
					}
					break;
				}
				case SamplesBegin:                                                      // :741-750
				{
					f->amb2 = v3(0.0f); f->avgPdf = 0.0f; f->indirect = v3(0.0f); f->cnt = 0.0f;
					f->toIor = (f->flags & kFlThick) ? ((f->flags & kFlOpposite) ? f->ior : 1.0f) : f->envIor;
					f->loopI = 0;
					act = SampleNext;
					break;
				}
				case SampleNext:                                                        // :753-784
				{
					if (f->loopI >= f->nS) { act = AmbientEnd; break; }
					const SampledData s = Sampled();
					const bool thick = (f->flags & kFlThick) != 0;
					const bool fullMetal = s.orm.z == 1.0f, mirror = fullMetal && s.orm.y <= 0.001f, hasTrans = !fullMetal && s.transmission > 0.0f;
					V3 term = v3(0.0f), direction = v3(0.0f);
					float pdf = 0.0f; bool transRay = false, ok = false, hasTransRay = (f->flags & kFlHasTransRay) != 0;
					int tries = 0;
					while ((!ok || (thick && !hasTransRay && f->loopI == (f->nS - 1))) && tries < 4096)
					{
						direction = v3(0.0f);
						const V2 Xi = BlueNoise();
						const float rs = mirror ? 1.0f : rng.Float01();
						const float rt = hasTrans ? rng.Float01() : 0.0f;
						ok = SampleBsdf(s, f->N, f->V, f->envIor, f->toIor, term, pdf, transRay, direction, Xi, rs, rt);
						hasTransRay = hasTransRay || transRay;
						tries++;
					}
					if (hasTransRay) f->flags |= kFlHasTransRay;
					if (!ok) { f->loopI++; break; }                                     // rejection budget exhausted: skip the sample
					float newIor = f->envIor;                                           // :769-778
					const bool opposite = (f->flags & kFlOpposite) != 0;
					if (opposite && transRay && thick) newIor = s.ior;
					else if (!opposite && transRay && thick) newIor = 1.0f;
					f->term = term; f->pdf = pdf; f->newIor = newIor;
					f->flags = (f->flags & ~kFlTransRay) | (transRay ? kFlTransRay : 0u);
					f->r2o = f->hitPoint + (transRay ? -f->offset : f->offset); f->r2d = direction;
					f->phase = kPhSample;
					Emit(f->r2o, f->r2d, f->hitTri);
					act = Done;
					break;
				}
				case AfterChildSample:                                                  // :816-833
				{
					const V3 value = glm_clamp(f->term * f->att * retVal, 0.0f, 10.0f);
					f->indirect = f->indirect + value;
					const MaterialGpu& hm = a.materials[MaterialOf(f->h2tri)];
					if (!(f->flags & kFlThick) && hm.transmission > 0.0f && hm.thickness > 0.0f)
					{
						f->value = value;
						StartSky(f->r2o, f->r2d, f->envIor, f->h2tri);
						f->phase = kPhSkySample;
						act = SkyStep;
					}
					else { f->cnt += 1.0f; f->loopI++; act = SampleNext; }
					break;
				}
				case AmbientEnd:                                                        // :838-852
				{
					const float pdfHemisphere = 1.0f / (kPiSailor * 2.0f);
					f->amb2 = f->amb2 / f->cnt; f->avgPdf = f->avgPdf / f->cnt;
					const V3 ambient = f->amb1 + f->amb2;
					if (ambient.x + ambient.y + ambient.z > 0.0f)
					{
						const V3 combined = f->amb1 * PowerHeuristic((int32_t)f->nA, pdfHemisphere, (int32_t)f->cnt, f->avgPdf) +
							f->amb2 * PowerHeuristic((int32_t)f->cnt, f->avgPdf, (int32_t)f->nA, pdfHemisphere);
						f->res = f->res + combined;
					}
					if (f->cnt > 0.0f) f->res = f->res + (f->indirect / f->cnt);
					act = Finish;
					break;
				}
				case Finish:                                                            // :855-871
				{
					f->res = f->res + f->emissive;
					if (f->bounceLimit > 0 && (f->flags & kFlAlpha))
					{
						Frame* parent = f;
						parent->phase = kPhChildAlpha;
						hd.depth++;
						f = FrameAt(hd.depth);
						f->rayD = parent->rayD; f->rayO = parent->hitPoint + parent->rayD * 0.0001f;
						f->ignoreTri = parent->hitTri; f->bounceLimit = parent->bounceLimit - 1;
						f->inAcc = parent->inAcc * (1.0f - parent->baseColor.w); f->envIor = parent->envIor;
						const uint32_t mb = parent->pMaxBounces - 1u, nsm = parent->pNumSamples - parent->S, nam = parent->pNumAmbient - parent->A;
						f->pMaxBounces = mb;                       // std::max(0u, x) is x
						f->pNumSamples = nsm > 1u ? nsm : 1u; f->pNumAmbient = nam > 1u ? nam : 1u;
						BeginCall();
						Emit(f->rayO, f->rayD, f->ignoreTri);
						act = Done;
					}
					else { retVal = f->res; act = Return; }
					break;
				}
				case AfterChildAlpha:                                                   // :868-870
				{
					const float al = f->baseColor.w;
					f->res = f->res * al + retVal * (1.0f - al);
					retVal = f->res;
					act = Return;
					break;
				}
				case Return:
				{
					if (hd.depth == 0)
					{
						// accumulator += Raytrace(...) (:466): one slot per (pixel, sample); summed in order by ResolveKernel
						const uint32_t x = hd.pixel % a.cam.width, y = hd.pixel / a.cam.width;
						const size_t idx = ((size_t)(y - a.rowBegin) * a.cam.width + x) * (a.msEnd - a.msBegin) + (hd.sample - a.msBegin);
						a.sampleBuf[idx * 3] = retVal.x; a.sampleBuf[idx * 3 + 1] = retVal.y; a.sampleBuf[idx * 3 + 2] = retVal.z;
						hd.active = 0;
						act = Done;
					}
					else
					{
						hd.depth--;
						f = FrameAt(hd.depth);
						act = f->phase == kPhChildAlpha ? AfterChildAlpha : AfterChildSample;
					}
					break;
				}
				case Done:
					return;
				}
				if (act == Done) return;
			}
		}
	};

	// One thread per pool slot: resume with the hit of the pending ray, refill finished slots from the hit queue.
	struct AdvanceKernel
	{
		IntegratorArgs a; uint32_t firstIteration;
		SPT_KERNEL_BODY void operator()(uint32_t slot) const
		{
			PathMachine pm(a, slot);
			pm.hd = a.headers[slot];
			if (firstIteration) { pm.hd.active = 0; pm.hd.depth = 0; }
			else if (pm.hd.active == 0) return;                               // idle for good: queue was exhausted when it finished
			if (pm.hd.active == 1)
			{
				pm.rng.key = pm.hd.rngKey; pm.rng.counter = pm.hd.rngCounter;
				pm.f = pm.FrameAt(pm.hd.depth);
				const Hit hit = a.hits[slot];
				pm.Advance(hit);
			}
			// refill: a finished (or never started) slot pulls primary samples until one produces a ray
			while (!pm.rayEmitted)
			{
				if (!pm.NextFromQueue()) { pm.hd.active = 0; break; }
			}
			if (pm.rayEmitted)
			{
				atomic_add_u32(a.activeCount, 1u);
				atomic_add_u64(a.rayCount, 1ull);
			}
			else
			{
				RayRec r; r.ox = r.oy = r.oz = 0.0f; r.ignoreTri = kNoHit; r.dx = r.dy = r.dz = 0.0f; r.tmax = -1.0f;   // idle marker
				a.rays[slot] = r;
			}
			pm.hd.rngCounter = pm.rng.counter;
			a.headers[slot] = pm.hd;
		}
	};

	// accumulator / msaa with the row flip of PathTracer.cpp:449,468-469
	struct ResolveKernel
	{
		const float* sampleBuf; float* image; uint32_t width, height, rowBegin, rowEnd, numSamples, msaa;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t x = i % width, yb = i / width, y = rowBegin + yb;
			V3 acc = v3(0.0f);
			const float* s = sampleBuf + (size_t)i * numSamples * 3;
			for (uint32_t k = 0; k < numSamples; k++) acc = acc + v3(s[k * 3], s[k * 3 + 1], s[k * 3 + 2]);
			const V3 res = acc / (float)msaa;
			const size_t o = ((size_t)(height - y - 1) * width + x) * 3;
			image[o] = res.x; image[o + 1] = res.y; image[o + 2] = res.z;
		}
	};

}
