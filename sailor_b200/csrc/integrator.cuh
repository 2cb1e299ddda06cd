// integrator.cuh — the path integrator as a BREADTH-FIRST evaluation of the reference's recursion tree
// (SURVEY §8 rows a13/a15/a17, Appendix A).
//
// The reference's Raytrace (reference Raytracing/PathTracer.cpp:622-879) is a recursive estimator, not a flat path
// sum: per-level clamps to [0,10] (:733,:788,:816), MIS on per-hit AVERAGES (:838-847), S-way branching at the
// first hit only (:716-717), a deterministic cut-off on the accumulated throughput (:810-811), an alpha-blend
// continuation that is a second recursive call (:858-871), a tail call when leaving a thick volume (:673-688) and
// the TraceSky transmission loop (:577-620).  Every one of those non-linearities is kept by evaluating the SAME
// call tree, but level by level instead of depth first:
//
//   record   = one Raytrace() activation whose closest hit is known (NodeRec, 128 B)
//   expand   : one thread per record of level d: shade the hit (:636-671), then emit EVERY ray the activation needs
//              at once — D shadow rays (:691-705), nA hemisphere rays (:720-737), nS importance rays (:753-784), the
//              alpha continuation / thick-volume tail ray — into the level's ray queue, with a 32-byte RayAux per
//              ray that carries what the ray's result will be weighted with (BRDF, term, pdf, ...)
//   trace    : one launch for the whole queue (trace_kernels.cuh / trace_fast.cuh, persistent threads): a status byte per ray (hit or
//              miss is all a shadow test, a hemisphere ray or an importance ray below the throughput cut needs), and a SlowRec (hit +
//              ray + index) per closest-hit query that hit something
//   classify : one thread per SlowRec: an importance ray that hit something and passes the throughput cut (:810-811) spawns a child
//              record of level d+1 that starts from this very hit (the reference re-traces the same ray at :632)
//   sky      : TraceSky continuations through thick transmissive surfaces (:591-616), a small side queue
//   gather   : bottom-up, one thread per record: replays the accumulation of :690-871 over the record's RayAux
//              entries IN INDEX ORDER (lights, hemisphere, importance samples) and hands  clamp(term*att*L)  to the
//              parent's RayAux (:816).
//
// Every ray of one recursion level of a whole batch of first hits is therefore traced by ONE launch, the number of
// launches per batch is bounded by maxBounces, and no thread ever waits on another path's tail.
//
// Deliberate, documented differences from the reference (none changes the estimator):
//  * the reference re-traces the importance ray when it recurses (:786 then :632 with the same ray and ignore
//    index); the child record starts from the hit the parent's ray already produced;
//  * all D shadow rays are traced even after one was blocked; the stale-`hitLight` quirk (:694: every light after
//    the first blocked one counts as blocked) is applied when the lights are summed;
//  * random numbers: same distributions as glm::linearRand on rand()%255 bytes and the same blue-noise
//    table walk (:934-1077), but drawn from a counter-based generator keyed per activation (seed, pixel, primary
//    sample, path in the call tree), so the image does not depend on scheduling or on how the frame is sharded;
//  * the unbounded rejection loop (:761-767) is capped at 4096 tries; an exhausted sample is skipped.
#pragma once
#include "pipeline.cuh"
#include "textures.cuh"
#include "lighting.cuh"
#include "blue_noise_table.h"

namespace spt
{
	// One first hit of the primary pass: which (pixel, primary sample) and the closest hit of its camera ray.
	struct alignas(16) PrimaryHitRec { uint32_t pixel, sample, pad0, pad1; float t, u, v; uint32_t tri; };
	static_assert(sizeof(PrimaryHitRec) == 32, "PrimaryHitRec layout");

	struct RenderStats { uint64_t rays, primarySamples, fanOutSamples, replayedRays; double secondsTraverse, secondsShade, secondsStage[4]; uint32_t traverseLaunches, batches; };

	constexpr uint32_t kNone = 0xFFFFFFFFu;
#ifndef SPT_FAN_OUT_MIN
#define SPT_FAN_OUT_MIN 8
#endif
	constexpr uint32_t kFanOutMinSamples = SPT_FAN_OUT_MIN;   // activations with at least this many hemisphere + importance samples get a warp

	// record flags
	enum : uint32_t
	{
		kNfOpposite = 1u, kNfThick = 2u, kNfAlpha = 4u, kNfFirst = 8u, kNfAmbientOn = 16u,
		kNfDone = 32u,          // result is final (own ray missed / tail call refused): expand and gather leave it alone
		kNfTail = 64u,          // thick-volume tail call (:673-688): result = the child's result
		kNfLevel0 = 128u,       // parent = pixel, parentAux = primary-sample index
	};

	// One Raytrace() activation (:622), four 32-byte sectors: [0] the ray and its closest hit (expand; classify only for thick volumes),
	// [1] what classify reads of an OWNER when one of its importance rays hit (rows 4-5), [2]-[3] everything gather reads and everything
	// expand / gather write (rows 5-8) - a classify pass and a gather pass each move half a record.
	struct alignas(16) NodeRec
	{
		V3 rayO; float envIor;
		V3 rayD; uint32_t pad0;
		float t, u, v; uint32_t tri;         // closest hit of (rayO, rayD)
		float inAcc; uint32_t pNum; uint64_t rngKey;             // pNum: params.m_numSamples | params.m_numAmbientSamples << 16 of this activation
		uint32_t bounces, flags, auxBase; float alpha;           // bounces = bounceLimit | params.m_maxBounces << 16; auxBase: first RayAux of this record; alpha: sample.m_baseColor.a
		uint32_t parent, parentAux, nA, nS;                      // parent record (or pixel for level 0); global RayAux index of the importance sample this activation hangs on (kNone for alpha / tail children); loop counts of :716-717
		V3 emissive; uint32_t child;         // sample.m_emissive; alpha / tail child record
		V3 result; uint32_t pad1;
	};
	SPT_HD uint32_t PackNum(uint32_t samples, uint32_t ambient) { return (samples & 0xFFFFu) | (ambient << 16); }      // both <= 65535 (checked by RenderFrame)
	static_assert(sizeof(NodeRec) == 128, "NodeRec layout");

	// ray kinds (RayAux::tag low 3 bits) and state bits.  Whether a ray hit anything is NOT in the tag: the trace kernel writes
	// one status byte per ray (IntegratorArgs::status), so rays that only need hit-or-miss are never touched again until gather.
	enum : uint32_t
	{
		kRkInactive = 0, kRkLight = 1, kRkHemi = 2, kRkSample = 3, kRkOwn = 4, kRkMask = 7u,
		kRsTransRay = 16u,      // sample: bTransmissionRay
		kRsHit = 64u,           // sample: hit with bounceLimit > 0 -> a = clamp(term*att*L) (written by the child's gather, or by classify when no child runs)
		kRsSky2 = 256u,         // sample: TraceSky behind the hit returned non-zero -> b = att (:825-831); hemi: a = final contribution of a sky walk
		kRsSkipped = 512u,      // sample: rejection budget exhausted
	};

	// What a ray's result is weighted with; rewritten in place by classify / sky / the child's gather.
	struct alignas(16) RayAux
	{
		float a0, a1, a2, a3;    // light: BRDF*intensity*angle | hemi: BRDF, angle (-> contribution after a sky walk) | sample: term, pdf -> value, pdf
		float b0, b1, b2;        // b0: sample's new environment IOR; b1: bits(owner record; child record for kRkOwn); sample after a sky walk: b = TraceSky attenuation
		uint32_t tag;            // kind | state
	};
	static_assert(sizeof(RayAux) == 32, "RayAux layout");

	// TraceSky (:577-620) in flight through a thick transmissive surface
	struct alignas(16) SkyState
	{
		V3 att; uint32_t targetAux;
		V3 prev; uint32_t ignore;
		V3 start; float ior;
		V3 dir; uint32_t jAndMax;      // j | maxBounces << 16
	};
	static_assert(sizeof(SkyState) == 64, "SkyState layout");

	struct ShadeCtx;
	struct LevelInfo { uint32_t recBegin, recEnd, rayCount, auxBase; };

	// device-side allocation state of one batch
	struct BatchCounters
	{
		uint32_t recAlloc;         // records allocated so far (all levels)
		uint32_t auxAlloc;         // RayAux entries handed to finished levels
		uint32_t overflow;         // any arena ran out: the batch is invalid, the host retries with a smaller one
		uint32_t skyCount[2];      // ping-pong sky queues
		uint32_t zero;             // always 0 (range begin)
		uint32_t fanEntries;       // activations handed to FanOutKernel at the current level
		uint32_t slowCount;        // rays of the current level whose closest hit needs ClassifyKernel
		uint32_t fanThreads[2];    // 8 x slots of the hemisphere / importance fan-out passes
		unsigned long long rays;   // closest-hit queries of the batch
		unsigned long long fanSamples;   // rays emitted by FanOutKernel
		LevelInfo level[66];
	};

	struct IntegratorArgs
	{
		// scene
		const V4* shade; const V4* centroid; const MaterialGpu* materials; TextureSet tex; const V4* lights; uint32_t numLights;
		const uint16_t* blueNoise;
		// camera / params
		CameraGpu cam; uint32_t rowBegin, rowEnd, msBegin, msEnd, msaa;
		uint32_t maxBounces, numSamples, numAmbientSamples; V3 ambient; uint64_t seed;
		// batch
		const PrimaryHitRec* hitQueue; uint32_t queueBegin, queueCount;     // first hits [queueBegin, queueBegin+queueCount) are level 0
		NodeRec* recs; uint32_t recCap;
		RayAux* aux; uint32_t auxCap; uint32_t hasSky;                       // hasSky: the scene has thick transmissive materials (TraceSky can continue)
		RayRec* rays; Hit* hits; uint32_t rayCap;                             // ray queue of the current level (hits: scratch of the traversal probe)
		uint8_t* status; SlowRec* slow;                                       // per RayAux: 1 = the ray hit something (written by the trace kernel); closest-hit queries that hit, for ClassifyKernel
		SkyState* sky[2]; RayRec* skyRays; Hit* skyHits; uint32_t skyCap;
		ShadeCtx* fan; uint32_t fanCap; uint32_t* fanSlots[2];                // activations whose samples are produced by FanOutKernel; 8-lane slot tables (4 x fanCap each)
		BatchCounters* c;
		float* sampleBuf;                       // 3 floats per (pixel in band, sample in range)
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
	SPT_HD uint64_t ChildRngKey(uint64_t key, uint32_t slot) { return Mix64(key + 0xD1B54A32D192ED03ULL * (uint64_t)(slot + 1u)); }
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

	// ---- helpers shared by the kernels ---------------------------------------------------------------------------
	SPT_KERNEL_BODY uint32_t MaterialOfTri(const IntegratorArgs& a, uint32_t tri) { return f2u(ld4(a.centroid + tri).w); }

	SPT_KERNEL_BODY V3 HitNormalOf(const IntegratorArgs& a, uint32_t tri, float u, float v)   // Bounds.cpp:525-527
	{
		const V4* S = a.shade + (size_t)tri * 9;
		const V4 n0 = ld4(S), n1 = ld4(S + 1), n2 = ld4(S + 2);
		const float w = 1.0f - u - v;
		return w * v3(n0.x, n0.y, n0.z) + u * v3(n1.x, n1.y, n1.z) + v * v3(n2.x, n2.y, n2.z);
	}

	// tmax encodes the query: +FLT_MAX closest hit; -FLT_MAX "any hit" (only hit-or-miss is consumed: the traversal may stop
	// at the first accepted triangle, which the reference's closest-hit walk would also have accepted); -1 inactive entry.
	SPT_KERNEL_BODY void WriteRay(RayRec* q, uint32_t i, V3 o, V3 d, uint32_t ignore, bool active, bool anyHit = false)
	{
		RayRec r; r.ox = o.x; r.oy = o.y; r.oz = o.z; r.ignoreTri = ignore; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.tmax = active ? (anyHit ? -kFltMax : kFltMax) : -1.0f;
		q[i] = r;
	}

	SPT_KERNEL_BODY void WriteAux(const IntegratorArgs& a, uint32_t g, float a0, float a1, float a2, float a3, float b0, uint32_t tag, uint32_t owner)
	{
		RayAux x; x.a0 = a0; x.a1 = a1; x.a2 = a2; x.a3 = a3; x.b0 = b0; x.b1 = u2f(owner); x.b2 = 0.0f; x.tag = tag;
		a.aux[g] = x;
	}

	SPT_KERNEL_BODY uint32_t AllocRecord(const IntegratorArgs& a)
	{
		const uint32_t i = atomic_inc_u32_agg(&a.c->recAlloc);
		if (i >= a.recCap) { a.c->overflow = 1u; return kNone; }
		return i;
	}

	// One TraceSky loop iteration (:585-616) on the result of the state's pending ray.
	// Returns true when the walk goes on (state updated, next ray = (start, dir, ignore)); false when done (att = result).
	SPT_KERNEL_BODY bool SkyAdvance(const IntegratorArgs& a, SkyState& s, const Hit& hit, V3& result)
	{
		if (hit.tri == kNoHit) { result = s.att; return false; }                      // :587-590
		const MaterialGpu& m = a.materials[MaterialOfTri(a, hit.tri)];
		const V3 hn = HitNormalOf(a, hit.tri, hit.u, hit.v);
		const bool hitOpp = dot(s.dir, hn) < 0.0f;
		if (!(m.transmission > 0.0f && m.thickness > 0.0f)) { result = v3(0.0f); return false; }   // :597-600
		const V3 hp = s.start + s.dir * hit.t;
		const float distance = length(hp - s.prev);
		s.prev = hp;
		const V3 wn = hitOpp ? hn : -hn;
		const float toIor = hitOpp ? m.ior : 1.0f;
		s.dir = CalculateRefraction(s.dir, wn, s.ior, toIor);
		s.ior = toIor;
		if (!hitOpp)
		{
			const V3 c = v3(-logf(m.attenuationColor[0]), -logf(m.attenuationColor[1]), -logf(m.attenuationColor[2])) / m.attenuationDistance;
			const V3 e = -c * distance;
			s.att = s.att * v3(expf(e.x), expf(e.y), expf(e.z));
		}
		s.start = hp; s.ignore = hit.tri;
		const uint32_t j = (s.jAndMax & 0xFFFFu) + 1u, mx = s.jAndMax >> 16;
		s.jAndMax = j | (mx << 16);
		if (j < mx) return true;
		result = v3(0.0f);                                                             // loop ran out (:619)
		return false;
	}

	// Final value of a TraceSky walk lands in the RayAux it was started for.
	SPT_KERNEL_BODY void SkyFinish(const IntegratorArgs& a, uint32_t g, V3 att)
	{
		RayAux x = a.aux[g];
		const bool nonZero = att.x != 0.0f || att.y != 0.0f || att.z != 0.0f;
		if ((x.tag & kRkMask) == kRkHemi)                                               // :730-735
		{
			V3 c = v3(0.0f);
			if (nonZero)
			{
				const float pdfHemisphere = 1.0f / (kPiSailor * 2.0f);
				const V3 at = att * v3(x.a0, x.a1, x.a2);
				c = glm_clamp((at * a.ambient * x.a3) / pdfHemisphere, 0.0f, 10.0f);
			}
			x.a0 = c.x; x.a1 = c.y; x.a2 = c.z; x.tag |= kRsSky2;
		}
		else                                                                            // :825-831
		{
			x.b0 = att.x; x.b1 = att.y; x.b2 = att.z;
			if (nonZero) x.tag |= kRsSky2;
		}
		a.aux[g] = x;
	}

	SPT_KERNEL_BODY void SkyPush(const IntegratorArgs& a, uint32_t q, const SkyState& s)
	{
		const uint32_t i = atomic_inc_u32_agg(&a.c->skyCount[q]);
		if (i >= a.skyCap) { a.c->overflow = 1u; return; }
		a.sky[q][i] = s;
		WriteRay(a.skyRays + (q ? a.skyCap : 0u), i, s.start, s.dir, s.ignore, true);      // each queue owns one half of skyRays
	}

	// ---- per-sample generation ---------------------------------------------------------------------------------------
	// The hemisphere and importance samples of one activation are independent of each other except for two walks the
	// reference does sequentially: the blue-noise table positions (:934-1077) and the "at least one transmission ray"
	// rule of thick volumes (:761).  Both are re-derived per sample from counters, so sample i can be produced by any
	// thread: the thread-per-activation loop of ExpandKernel and the warp-per-activation FanOutKernel (first hits
	// with many samples) emit bit-identical rays.
	struct alignas(16) ShadeCtx        // shading context of one activation whose samples are fanned out over a warp
	{
		V3 N; float rough;
		V3 V; float metal;
		V3 hitPoint; float ior;
		V3 offset; float thickness;
		V4 baseColor;
		float transmission, envIor, inAcc; uint32_t flags;       // flags: kNf* of the activation
		uint64_t rngKey; uint32_t seedX, seedY;
		uint32_t rec, tri, rayBase, nHemi;                       // rayBase: level-local index of the first hemisphere ray
		uint32_t nS, bounceLimit, pMaxBounces, pad;
	};
	static_assert(sizeof(ShadeCtx) == 144, "ShadeCtx layout");

	SPT_KERNEL_BODY SampledData SampledOf(const ShadeCtx& c)
	{
		SampledData s; s.baseColor = c.baseColor; s.orm = v3(0.0f, c.rough, c.metal); s.emissive = v3(0.0f); s.normal = v3(0.0f, 0.0f, 1.0f);
		s.ior = c.ior; s.thickness = c.thickness; s.transmission = c.transmission; s.opaque = true;
		return s;
	}

	// Table position of the m-th draw of a walk that starts at seed0: indices run up to 687, then the walk reseeds with
	// linearRand(0, 680) (:1068-1076).  Reseed values come from a counter-based stream so the walk has random access.
	SPT_KERNEL_BODY uint32_t BlueNoiseIndex(uint64_t recKey, uint32_t axis, uint32_t seed0, uint32_t m)
	{
		uint32_t s = seed0, rem = m;
		Rng r; r.key = ChildRngKey(recKey, 0x30000000u + axis); r.counter = 0;
		while (rem >= 688u - s) { rem -= 688u - s; s = r.Seed681(); }
		return s + rem;
	}

	// Hemisphere sample k of an activation (:722-727, 732-733): direction, BRDF and cosine.
	SPT_KERNEL_BODY void DrawHemisphere(const ShadeCtx& c, const SampledData& s, uint32_t k, V3& toL, V3& brdf, float& angle)
	{
		Rng r; r.key = ChildRngKey(c.rngKey, 0x10000000u + k); r.counter = 0;
		const float r0 = r.Float01(), r1 = r.Float01();                               // NextVec2_Linear (:929-932)
		const V3 H = ImportanceSampleHemisphere(v2(r0, r1), c.N);
		toL = 2.0f * dot(c.V, H) * H - c.V;
		brdf = c.pMaxBounces > 0 ? CalculateBRDF(c.V, c.N, toL, s) : v3(0.0f);       // TraceSky's loop runs maxBounces times (:581)
		angle = glm_max(0.0f, dot(toL, c.N));
	}

	// Importance sample i (:755-767).  anyTrans carries bHasTransmissionRay in and out.  Returns false when the (capped)
	// rejection loop gives up.
	SPT_KERNEL_BODY bool DrawImportance(const IntegratorArgs& a, const ShadeCtx& c, const SampledData& s, uint32_t i, bool requireTrans, bool& anyTrans,
		V3& term, float& pdf, bool& transRay, V3& direction)
	{
		const bool thick = (c.flags & kNfThick) != 0, opposite = (c.flags & kNfOpposite) != 0;
		const bool fullMetal = s.orm.z == 1.0f, mirror = fullMetal && s.orm.y <= 0.001f, hasTrans = !fullMetal && s.transmission > 0.0f;
		const float toIor = thick ? (opposite ? s.ior : 1.0f) : c.envIor;             // :749
		Rng r; r.key = ChildRngKey(c.rngKey, 0x20000000u + i); r.counter = 0;
		for (uint32_t tries = 0; tries < 4096u; tries++)
		{
			const uint32_t m = i + tries * c.nS;                                       // every retry takes a later position of the walk
			const V2 Xi = v2((float)a.blueNoise[BlueNoiseIndex(c.rngKey, 0, c.seedX, m)] * (1.0f / 1024.0f),
				(float)a.blueNoise[BlueNoiseIndex(c.rngKey, 1, c.seedY, m)] * (1.0f / 1024.0f));
			const float rs = mirror ? 1.0f : r.Float01();
			const float rt = hasTrans ? r.Float01() : 0.0f;
			direction = v3(0.0f);
			const bool ok = SampleBsdf(s, c.N, c.V, c.envIor, toIor, term, pdf, transRay, direction, Xi, rs, rt);
			anyTrans = anyTrans || transRay;
			if (ok && !(requireTrans && !anyTrans)) return true;
		}
		return false;
	}

	// Emit hemisphere ray k / importance ray i of an activation into the level's queue (ray r, RayAux g).
	SPT_KERNEL_BODY void EmitHemisphere(const IntegratorArgs& a, const ShadeCtx& c, const SampledData& s, uint32_t k, uint32_t r, uint32_t g)
	{
		V3 toL, brdf; float angle;
		DrawHemisphere(c, s, k, toL, brdf, angle);
		const bool live = c.pMaxBounces > 0;
		WriteRay(a.rays, r, c.hitPoint + c.offset, toL, c.tri, live, !a.hasSky);       // without thick transmissive materials TraceSky is a boolean
		// a dead hemisphere ray contributes 0: gather skips entries that are neither kRsMiss nor kRsSky2
		WriteAux(a, g, brdf.x, brdf.y, brdf.z, angle, 0.0f, live ? kRkHemi : kRkInactive, c.rec);
	}
	SPT_KERNEL_BODY void EmitImportance(const IntegratorArgs& a, const ShadeCtx& c, const SampledData& s, uint32_t i, bool requireTrans, bool& anyTrans, uint32_t r, uint32_t g)
	{
		V3 term = v3(0.0f), direction = v3(0.0f); float pdf = 0.0f; bool transRay = false;
		if (!DrawImportance(a, c, s, i, requireTrans, anyTrans, term, pdf, transRay, direction))
		{
			WriteRay(a.rays, r, c.hitPoint, v3(0.0f), c.tri, false);
			WriteAux(a, g, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, kRkInactive | kRsSkipped, c.rec);
			return;
		}
		const bool thick = (c.flags & kNfThick) != 0, opposite = (c.flags & kNfOpposite) != 0;
		float newIor = c.envIor;                                                       // :769-778
		if (opposite && transRay && thick) newIor = s.ior;
		else if (!opposite && transRay && thick) newIor = 1.0f;
		// the hit itself is consumed only when a child activation can start from it (:797-814) or TraceSky can continue (:822-832)
		const bool boolOnly = c.bounceLimit == 0 || (!thick && !a.hasSky && !(c.inAcc * length(term) * s.baseColor.w > 0.01f));
		WriteRay(a.rays, r, c.hitPoint + (transRay ? -c.offset : c.offset), direction, c.tri, true, boolOnly);
		WriteAux(a, g, term.x, term.y, term.z, pdf, newIor, kRkSample | (transRay ? kRsTransRay : 0u), c.rec);
	}

	// ---- level 0: first hits of the primary pass become records (inside ExpandKernel: a separate pass wrote 128 bytes per first hit that
	// expand read back at once) ------------------------------------------------------------------------------------
	struct SeedKernel
	{
		IntegratorArgs a;
		SPT_KERNEL_BODY NodeRec Make(uint32_t i) const
		{
			const PrimaryHitRec rec = a.hitQueue[a.queueBegin + i];
			const uint32_t x = rec.pixel % a.cam.width, y = rec.pixel / a.cam.width;
			Rng rng; rng.key = PrimaryRngKey(a.seed, rec.pixel, a.msaa, rec.sample); rng.counter = 0;
			float ox = 0.5f, oy = 0.5f;                                           // :460
			if (rec.sample != 0) { ox = rng.Float01(); oy = rng.Float01(); }
			NodeRec n;
			n.rayO = a.cam.pos; n.parent = rec.pixel;
			n.rayD = PrimaryDir(a.cam, x, y, ox, oy); n.parentAux = rec.sample;
			n.t = rec.t; n.u = rec.u; n.v = rec.v; n.tri = rec.tri;
			n.inAcc = 1.0f; n.envIor = 1.0f; n.bounces = a.maxBounces | (a.maxBounces << 16); n.flags = kNfLevel0;
			n.rngKey = ChildRngKey(rng.key, 0x7FFFFFFFu); n.auxBase = 0; n.child = kNone;
			n.pNum = PackNum(a.numSamples, a.numAmbientSamples); n.pad0 = 0; n.nA = 0; n.nS = 0;
			n.emissive = v3(0.0f); n.alpha = 1.0f; n.result = v3(0.0f); n.pad1 = 0;
			return n;
		}
	};

	// ---- expand: shade one activation and emit all of its rays ----------------------------------------------------
	// The shading context of a hit (PathTracer.cpp:636-661) and GetMaterialData (:881-927): interpolated frame, the face normal turned
	// against the ray, transformed uv, the material's factors times its textures.  One definition for the integrator (ExpandKernel) and for the
	// parity hook SailorPt_ShadeHits, which compares it value for value with the reference's own GetMaterialData.
	struct HitShading { SampledData s; V3 faceNormal, tangent, bitangent; V2 uvT; uint32_t matIdx; bool opposite; };
	SPT_HD void ShadeHit(const V4* shade, const MaterialGpu* materials, const TextureSet& tex, uint32_t tri, float bu, float bv, V3 rayD, HitShading& h)
	{
		const V4* S = shade + (size_t)tri * 9;
		const V4 s0 = ld4(S), s1 = ld4(S + 1), s2 = ld4(S + 2), s3 = ld4(S + 3), s4 = ld4(S + 4), s5 = ld4(S + 5), s6 = ld4(S + 6), s7 = ld4(S + 7), s8 = ld4(S + 8);
		const float bw = 1.0f - bu - bv;
		V3 faceNormal = bw * v3(s0.x, s0.y, s0.z) + bu * v3(s1.x, s1.y, s1.z) + bv * v3(s2.x, s2.y, s2.z);
		h.tangent = bw * v3(s3.x, s3.y, s3.z) + bu * v3(s4.x, s4.y, s4.z) + bv * v3(s5.x, s5.y, s5.z);
		h.bitangent = bw * v3(s6.x, s6.y, s6.z) + bu * v3(s7.x, s7.y, s7.z) + bv * v3(s8.x, s8.y, s8.z);
		h.opposite = dot(faceNormal, rayD) < 0.0f;
		if (!h.opposite) faceNormal = faceNormal * -1.0f;
		h.faceNormal = faceNormal;
		const V2 uv = bw * v2(s0.w, s1.w) + bu * v2(s2.w, s3.w) + bv * v2(s4.w, s5.w);
		h.matIdx = f2u(s6.w);
		const MaterialGpu& m = materials[h.matIdx];
		const float* T = m.uvTransform;
		const float tu = T[0] * uv.x + T[4] * uv.y + T[8] * 1.0f, tv = T[1] * uv.x + T[5] * uv.y + T[9] * 1.0f;
		h.uvT = v2(tu, tv);
		// GetMaterialData (:881-927)
		SampledData& s = h.s;
		s.baseColor = v4(m.baseColor[0], m.baseColor[1], m.baseColor[2], m.baseColor[3]);
		V3 nrm = v3(0.0f, 0.0f, 1.0f);
		s.orm = v3(0.0f, m.roughness, m.metallic);
		s.emissive = v3(m.emissive[0], m.emissive[1], m.emissive[2]);
		s.transmission = m.transmission;
		if (m.texBase != kNoTexture) { const V4 t = SampleTexture(tex, m.texBase, tu, tv); s.baseColor = v4(s.baseColor.x * t.x, s.baseColor.y * t.y, s.baseColor.z * t.z, s.baseColor.w * t.w); }
		if (m.texEmissive != kNoTexture) { const V4 t = SampleTexture(tex, m.texEmissive, tu, tv); s.emissive = s.emissive * v3(t.x, t.y, t.z); }
		if (m.texMetallicRoughness != kNoTexture) { const V4 t = SampleTexture(tex, m.texMetallicRoughness, tu, tv); s.orm = v3(t.x, s.orm.y * t.y, s.orm.z * t.z); }
		if (m.texNormal != kNoTexture) { const V4 t = SampleTexture(tex, m.texNormal, tu, tv); nrm = v3(t.x, t.y, t.z); }
		if (m.texTransmission != kNoTexture) { const V4 t = SampleTexture(tex, m.texTransmission, tu, tv); s.transmission *= t.x; }
		if (m.blendMode == kMask) s.baseColor.w = (s.baseColor.w > m.alphaCutoff) ? 1.0f : 0.0f;
		s.opaque = m.blendMode == kOpaque;
		s.normal = nrm; s.ior = m.ior; s.thickness = m.thickness;
	}
	// normalize(tbn * normal), glm mat3 * vec3 order (type_mat3x3.inl:468-474), PathTracer.cpp:655
	SPT_HD V3 WorldNormal(const HitShading& h)
	{
		const V3 nrm = h.s.normal;
		return normalize(v3(h.tangent.x * nrm.x + h.bitangent.x * nrm.y + h.faceNormal.x * nrm.z,
			h.tangent.y * nrm.x + h.bitangent.y * nrm.y + h.faceNormal.y * nrm.z,
			h.tangent.z * nrm.x + h.bitangent.z * nrm.y + h.faceNormal.z * nrm.z));
	}
	// :657-659: alpha-blended surfaces scale the sample counts
	SPT_HD uint32_t AlphaScaledCount(const SampledData& s, uint32_t count)
	{
		const bool alphaBlend = !s.opaque && s.baseColor.w < 1.0f;
		const uint32_t r = (uint32_t)roundf(s.baseColor.w * (float)count);
		return alphaBlend ? (r > 1u ? r : 1u) : count;
	}

	// SailorPt_SampleGenerators (include/sailor_pt.h): the generators exactly as the kernels call them
	struct SampleGeneratorsKernel
	{
		uint64_t key; uint32_t kind, count; const uint16_t* blueNoise; float* out;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			Rng r; r.key = key; r.counter = 0;
			if (kind == 0u) for (uint32_t i = 0; i < count; i++) out[i] = u2f(r.U32());
			else if (kind == 1u) for (uint32_t i = 0; i < count; i++) out[i] = r.Float01();
			else
			{
				const uint32_t seedX = r.Seed681(), seedY = r.Seed681();                   // PathTracer.cpp:626-627
				for (uint32_t m = 0; m < count; m++)
				{
					const uint32_t ix = BlueNoiseIndex(key, 0, seedX, m), iy = BlueNoiseIndex(key, 1, seedY, m);
					out[4 * m] = (float)blueNoise[ix] * (1.0f / 1024.0f); out[4 * m + 1] = (float)blueNoise[iy] * (1.0f / 1024.0f);
					out[4 * m + 2] = (float)ix; out[4 * m + 3] = (float)iy;
				}
			}
		}
	};

	// SailorPt_ShadeHits: 28 floats per hit (include/sailor_pt.h)
	struct ShadeHitsKernel
	{
		const V4* shade; const MaterialGpu* materials; TextureSet tex; const uint32_t* tri; const float* uv; const float* dir; float* out; uint32_t numTris, numSamples, numAmbient;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			float* o = out + (size_t)i * 28;
			if (tri[i] >= numTris) { for (int k = 0; k < 28; k++) o[k] = 0.0f; return; }
			HitShading h;
			ShadeHit(shade, materials, tex, tri[i], uv[2 * i], uv[2 * i + 1], v3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), h);
			const V3 N = WorldNormal(h);
			o[0] = h.s.baseColor.x; o[1] = h.s.baseColor.y; o[2] = h.s.baseColor.z; o[3] = h.s.baseColor.w;
			o[4] = h.s.orm.x; o[5] = h.s.orm.y; o[6] = h.s.orm.z; o[7] = h.s.emissive.x; o[8] = h.s.emissive.y; o[9] = h.s.emissive.z;
			o[10] = h.s.normal.x; o[11] = h.s.normal.y; o[12] = h.s.normal.z; o[13] = h.s.transmission; o[14] = h.s.ior; o[15] = h.s.thickness; o[16] = h.s.opaque ? 1.0f : 0.0f;
			o[17] = N.x; o[18] = N.y; o[19] = N.z; o[20] = h.faceNormal.x; o[21] = h.faceNormal.y; o[22] = h.faceNormal.z; o[23] = h.uvT.x; o[24] = h.uvT.y;
			o[25] = h.opposite ? 1.0f : 0.0f; o[26] = (float)AlphaScaledCount(h.s, numSamples); o[27] = (float)AlphaScaledCount(h.s, numAmbient);
		}
	};

	struct ExpandKernel
	{
		IntegratorArgs a; uint32_t level;

		struct Ctx2
		{
			Rng rng; uint32_t seedX, seedY;
		};

		// child activation that still needs its own closest hit (alpha continuation / thick-volume tail call)
		SPT_KERNEL_BODY uint32_t SpawnOwn(const NodeRec& n, uint32_t self, V3 o, V3 d, uint32_t bounceLimit, uint32_t pMaxBounces, uint32_t pNumSamples,
			uint32_t pNumAmbient, float inAcc, float envIor, uint32_t slot) const
		{
			const uint32_t ci = AllocRecord(a);
			if (ci == kNone) return kNone;
			NodeRec c;
			c.rayO = o; c.parent = self; c.rayD = d; c.parentAux = kNone;
			c.t = 0.0f; c.u = 0.0f; c.v = 0.0f; c.tri = kNoHit;
			c.inAcc = inAcc; c.envIor = envIor; c.bounces = bounceLimit | (pMaxBounces << 16); c.flags = kNfDone;   // a miss is final (:873-876); a hit clears kNfDone (ClassifyKernel)
			c.rngKey = ChildRngKey(n.rngKey, slot); c.auxBase = 0; c.child = kNone;
			c.pNum = PackNum(pNumSamples, pNumAmbient); c.pad0 = 0; c.nA = 0; c.nS = 0;
			c.emissive = v3(0.0f); c.alpha = 1.0f; c.result = a.ambient; c.pad1 = 0;
			a.recs[ci] = c;
			return ci;
		}

		SPT_KERNEL_BODY void operator()(uint32_t ri) const
		{
			NodeRec n;
			if (level == 0u) { n = SeedKernel{ a }.Make(ri); a.recs[ri] = n; }          // level 0 = the batch's first hits, records [0, queueCount)
			else n = a.recs[ri];
			if (n.flags & kNfDone) return;
			LevelInfo* L = &a.c->level[level];
			const uint32_t bounceLimit = n.bounces & 0xFFFFu, pMaxBounces = n.bounces >> 16;
			const uint32_t pNumSamples = n.pNum & 0xFFFFu, pNumAmbient = n.pNum >> 16;
			Ctx2 cx; cx.rng.key = n.rngKey; cx.rng.counter = 0;
			cx.seedX = cx.rng.Seed681(); cx.seedY = cx.rng.Seed681();                   // :626-627

			// ---- :636-671 shading context
			const uint32_t tri = n.tri;
			HitShading hs;
			ShadeHit(a.shade, a.materials, a.tex, tri, n.u, n.v, n.rayD, hs);
			const MaterialGpu& m = a.materials[hs.matIdx];
			const SampledData& s = hs.s;
			const V3 faceNormal = hs.faceNormal, nrm = hs.s.normal;
			const V3 tangent = hs.tangent, bitangent = hs.bitangent;
			const bool opposite = hs.opposite;

			const V3 V = -normalize(n.rayD);
			// tbn * normal, glm mat3 * vec3 order (type_mat3x3.inl:468-474)
			const V3 N = normalize(v3(tangent.x * nrm.x + bitangent.x * nrm.y + faceNormal.x * nrm.z,
				tangent.y * nrm.x + bitangent.y * nrm.y + faceNormal.y * nrm.z,
				tangent.z * nrm.x + bitangent.z * nrm.y + faceNormal.z * nrm.z));
			const bool alphaBlend = !s.opaque && s.baseColor.w < 1.0f;
			const uint32_t sRound = (uint32_t)roundf(s.baseColor.w * (float)pNumSamples), aRound = (uint32_t)roundf(s.baseColor.w * (float)pNumAmbient);
			const uint32_t numSamples = alphaBlend ? (sRound > 1u ? sRound : 1u) : pNumSamples;
			const uint32_t numAmbient = alphaBlend ? (aRound > 1u ? aRound : 1u) : pNumAmbient;
			const V3 offset = 0.000001f * faceNormal;
			const bool fullMetal = s.orm.z == 1.0f;
			const bool hasTrans = !fullMetal && s.transmission > 0.0f;
			const bool thick = hasTrans && m.thickness > 0.0f;
			const V3 hitPoint = n.rayO + n.rayD * n.t;

			if (!opposite && thick)                                                      // :673-688 (tail call)
			{
				const V3 nd = CalculateRefraction(n.rayD, N, n.envIor, 1.0f);
				NodeRec* out = a.recs + ri;
				if (eq0(nd) || bounceLimit == 0) { out->result = v3(0.0f); out->flags = n.flags | kNfDone; return; }
				const uint32_t base = atomic_inc_u32_agg(&L->rayCount);
				if (base >= a.rayCap || L->auxBase + base >= a.auxCap) { a.c->overflow = 1u; out->result = v3(0.0f); out->flags = n.flags | kNfDone; return; }
				const V3 d = nd - offset;
				const uint32_t ci = SpawnOwn(n, ri, hitPoint, d, bounceLimit - 1u, pMaxBounces, pNumSamples, pNumAmbient, n.inAcc, 1.0f, 0x40000000u);
				WriteRay(a.rays, base, hitPoint, d, tri, ci != kNone);
				WriteAux(a, L->auxBase + base, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, ci != kNone ? kRkOwn : kRkInactive, ci);
				out->auxBase = L->auxBase + base; out->child = ci; out->flags = n.flags | kNfTail;
				out->nA = 0; out->nS = 0;
				return;
			}

			const bool first = bounceLimit == pMaxBounces;                               // :636
			const bool ambientOn = a.ambient.x + a.ambient.y + a.ambient.z > 0.0f;       // :708
			const uint32_t nA = ambientOn ? (first ? numAmbient : 1u) : 0u;              // :716
			const uint32_t nS = ambientOn ? (first ? numSamples : 1u) : 0u;              // :717
			const uint32_t nHemi = thick ? 0u : nA;                                      // :720
			const bool alphaChild = bounceLimit > 0 && alphaBlend;                       // :858
			const uint32_t nRays = a.numLights + nHemi + nS + (alphaChild ? 1u : 0u);
			NodeRec* out = a.recs + ri;
			uint32_t base = 0;
			if (nRays)
			{
				base = atomic_add_u32_agg(&L->rayCount, nRays);
				if (base + nRays > a.rayCap || L->auxBase + base + nRays > a.auxCap)
				{
					a.c->overflow = 1u; out->result = v3(0.0f); out->flags = n.flags | kNfDone; return;
				}
			}
			const uint32_t g0 = L->auxBase + base;
			uint32_t r = base;

			// ---- :691-705 directional lights
			for (uint32_t j = 0; j < a.numLights; j++, r++)
			{
				const V4 ld = a.lights[j * 2], li = a.lights[j * 2 + 1];
				const V3 toL = -v3(ld.x, ld.y, ld.z);
				const float angle = glm_max(0.0f, dot(toL, N));
				const V3 w = CalculateBRDF(V, N, toL, s) * v3(li.x, li.y, li.z) * angle;
				WriteRay(a.rays, r, hitPoint + offset, toL, tri, true, true);            // shadow test: boolean (:699)
				WriteAux(a, L->auxBase + r, w.x, w.y, w.z, 0.0f, 0.0f, kRkLight, ri);
			}
			// ---- :720-784 hemisphere + importance samples
			{
				ShadeCtx c;
				c.N = N; c.rough = s.orm.y; c.V = V; c.metal = s.orm.z; c.hitPoint = hitPoint; c.ior = s.ior; c.offset = offset; c.thickness = s.thickness;
				c.baseColor = s.baseColor; c.transmission = s.transmission; c.envIor = n.envIor; c.inAcc = n.inAcc;
				c.flags = (opposite ? kNfOpposite : 0u) | (thick ? kNfThick : 0u);
				c.rngKey = n.rngKey; c.seedX = cx.seedX; c.seedY = cx.seedY;
				c.rec = ri; c.tri = tri; c.rayBase = r; c.nHemi = nHemi; c.nS = nS; c.bounceLimit = bounceLimit; c.pMaxBounces = pMaxBounces; c.pad = 0;
				bool fanned = false;
				if (nHemi + nS >= kFanOutMinSamples)
				{
					// many samples (first hits): hand them to FanOutKernel in slots of 8 lanes; this thread only reserves the rays
					const uint32_t e = atomic_inc_u32_agg(&a.c->fanEntries);
					if (e < a.fanCap)
					{
						a.fan[e] = c; fanned = true;
						atomic_add_u64_agg(&a.c->fanSamples, (unsigned long long)(nHemi + nS));
						const uint32_t cnt[2] = { nHemi, nS };
						for (uint32_t pass = 0; pass < 2; pass++)
						{
							uint32_t slots = (cnt[pass] + 7u) / 8u; if (slots > 4u) slots = 4u;
							if (!slots) continue;
							const uint32_t base = atomic_add_u32_agg(&a.c->fanThreads[pass], 8u * slots) / 8u;
							for (uint32_t k = 0; k < slots; k++) a.fanSlots[pass][base + k] = e | (k << 24) | ((slots - 1u) << 27);
						}
					}
				}
				if (!fanned)
				{
					const SampledData sc = SampledOf(c);
					for (uint32_t k = 0; k < nHemi; k++) EmitHemisphere(a, c, sc, k, r + k, L->auxBase + r + k);
					bool anyTrans = false;
					for (uint32_t i = 0; i < nS; i++) EmitImportance(a, c, sc, i, thick && i == nS - 1u, anyTrans, r + nHemi + i, L->auxBase + r + nHemi + i);
				}
				r += nHemi + nS;
			}
			// ---- :858-871 alpha-blend continuation: an independent activation, started now
			uint32_t child = kNone;
			if (alphaChild)
			{
				const uint32_t mb = pMaxBounces - 1u, nsm = pNumSamples - numSamples, nam = pNumAmbient - numAmbient;   // std::max(0u, x) is x
				const V3 o = hitPoint + n.rayD * 0.0001f;
				child = SpawnOwn(n, ri, o, n.rayD, bounceLimit - 1u, mb, (pNumSamples > numSamples && nsm > 1u) ? nsm : 1u,
					(pNumAmbient > numAmbient && nam > 1u) ? nam : 1u, n.inAcc * (1.0f - s.baseColor.w), n.envIor, 0x40000001u);
				WriteRay(a.rays, r, o, n.rayD, tri, child != kNone);
				WriteAux(a, L->auxBase + r, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, child != kNone ? kRkOwn : kRkInactive, child);
				r++;
			}
			out->auxBase = g0; out->child = child;
			out->flags = n.flags | (opposite ? kNfOpposite : 0u) | (thick ? kNfThick : 0u) | (alphaBlend ? kNfAlpha : 0u) | (first ? kNfFirst : 0u) | (ambientOn ? kNfAmbientOn : 0u);
			out->nA = nA; out->nS = nS;
			out->emissive = s.emissive; out->alpha = s.baseColor.w;
		}
	};

	// ---- fan-out: activations with many samples get 1-4 slots of 8 lanes per pass; lane l produces samples l, l+stride, ... ----
	// Two passes (hemisphere rays, importance rays) so that the lanes of a warp run the same code; rays of one activation are
	// written by consecutive lanes (coalesced).
	struct FanOutKernel
	{
		IntegratorArgs a; uint32_t level; uint32_t pass;
		SPT_KERNEL_BODY void operator()(uint32_t w) const                             // w = slot * 8 + lane
		{
			const uint32_t packed = a.fanSlots[pass][w >> 3];
			const uint32_t e = packed & 0xFFFFFFu, slot = (packed >> 24) & 7u, stride = (((packed >> 27) & 7u) + 1u) * 8u;
			const uint32_t lane = slot * 8u + (w & 7u);
			const ShadeCtx c = a.fan[e];
			const SampledData s = SampledOf(c);
			const uint32_t auxBase = a.c->level[level].auxBase;
			if (pass == 0)
			{
				for (uint32_t k = lane; k < c.nHemi; k += stride) EmitHemisphere(a, c, s, k, c.rayBase + k, auxBase + c.rayBase + k);
				return;
			}
			const bool thick = (c.flags & kNfThick) != 0;
			for (uint32_t i = lane; i < c.nS; i += stride)
			{
				bool anyTrans = false;
				const bool last = thick && i == c.nS - 1u;
				if (last)
				{
					// bHasTransmissionRay (:761) looks back over every attempt of the earlier samples: replay them (thick volumes only)
					for (uint32_t j = 0; j + 1u < c.nS && !anyTrans; j++)
					{
						V3 term, direction; float pdf; bool transRay;
						DrawImportance(a, c, s, j, false, anyTrans, term, pdf, transRay, direction);
					}
				}
				EmitImportance(a, c, s, i, last, anyTrans, c.rayBase + c.nHemi + i, auxBase + c.rayBase + c.nHemi + i);
			}
		}
	};

	// ---- classify: fold one ray's closest hit into its RayAux, spawn the child activation of an importance hit --------
	struct ClassifyKernel            // one thread per entry of the level's slow list
	{
		IntegratorArgs a; uint32_t level;
		SPT_KERNEL_BODY void operator()(uint32_t k) const
		{
			const SlowRec sr = a.slow[k];
			const uint32_t i = sr.index;
			const uint32_t g = a.c->level[level].auxBase + i;
			RayAux x = a.aux[g];
			const uint32_t kind = x.tag & kRkMask;
			Hit h; h.t = sr.t; h.u = sr.u; h.v = sr.v; h.tri = sr.tri;
			if (kind == kRkInactive || kind == kRkLight || h.tri == kNoHit) return;      // cannot be on the list
			const uint32_t owner = f2u(x.b1);
			if (kind == kRkOwn)
			{
				NodeRec* c = a.recs + owner;                                              // the child activation has its closest hit now
				c->t = h.t; c->u = h.u; c->v = h.v; c->tri = h.tri; c->flags &= ~kNfDone;
				return;
			}
			const V3 ro = v3(sr.ox, sr.oy, sr.oz), rd = v3(sr.dx, sr.dy, sr.dz);
			if (kind == kRkHemi)
			{
				const NodeRec* o = a.recs + owner;
				SkyState s; s.att = v3(1.0f); s.targetAux = g; s.prev = ro; s.ignore = a.rays[i].ignoreTri; s.start = ro; s.ior = o->envIor; s.dir = rd;
				s.jAndMax = 0u | ((o->bounces >> 16) << 16);
				V3 att;
				if (SkyAdvance(a, s, h, att)) SkyPush(a, 0, s);
				else SkyFinish(a, g, att);
				return;
			}
			// kRkSample (:786-833)
			const V3 term = v3(x.a0, x.a1, x.a2);
			const NodeRec* o = a.recs + owner;
			const uint32_t bounceLimit = o->bounces & 0xFFFFu, pMaxBounces = o->bounces >> 16;
			if (bounceLimit == 0) return;                                                // :797, :835 (such rays are boolean and never get here)
			const uint32_t oflags = o->flags;
			V3 att = v3(1.0f);
			if ((oflags & kNfOpposite) && (x.tag & kRsTransRay) && (oflags & kNfThick))  // :801-807
			{
				const MaterialGpu& m = a.materials[MaterialOfTri(a, o->tri)];
				const V3 hitPoint = o->rayO + o->rayD * o->t;
				const V3 h2p = ro + rd * h.t;
				const float distance = length(h2p - hitPoint);
				const V3 c = v3(-logf(m.attenuationColor[0]), -logf(m.attenuationColor[1]), -logf(m.attenuationColor[2])) / m.attenuationDistance;
				const V3 e = -c * distance;
				att = v3(expf(e.x), expf(e.y), expf(e.z));
			}
			const V3 T = term * att;
			const float newAcc = o->inAcc * length(T) * o->alpha;                        // :810
			bool spawned = false;
			if (newAcc > 0.01f)                                                           // :811-814
			{
				const uint32_t ci = AllocRecord(a);
				if (ci != kNone)
				{
					NodeRec c;
					c.rayO = ro; c.parent = owner; c.rayD = rd; c.parentAux = g;
					c.t = h.t; c.u = h.u; c.v = h.v; c.tri = h.tri;
					c.inAcc = newAcc; c.envIor = x.b0; c.bounces = (bounceLimit - 1u) | (pMaxBounces << 16); c.flags = 0;
					c.rngKey = ChildRngKey(o->rngKey, g - o->auxBase); c.auxBase = 0; c.child = kNone;
					c.pNum = o->pNum; c.pad0 = 0; c.nA = 0; c.nS = 0;
					c.emissive = v3(0.0f); c.alpha = 1.0f; c.result = v3(0.0f); c.pad1 = 0;
					a.recs[ci] = c;
					spawned = true;
				}
			}
			// a = term*att until the child's gather replaces it by clamp(term*att*L); without a child L = 0 (:809,:816)
			V3 val = T;
			if (!spawned) val = glm_clamp(T * v3(0.0f), 0.0f, 10.0f);
			x.a0 = val.x; x.a1 = val.y; x.a2 = val.z; x.b0 = 0.0f; x.b1 = 0.0f; x.b2 = 0.0f; x.tag |= kRsHit;
			a.aux[g] = x;
			// :822-832 sky behind a thick transmissive surface, independent of the child.  Without such a material in the scene (hasSky)
			// the test cannot pass, and the two dependent scattered loads it needs (triangle -> material) are skipped.
			if (a.hasSky && !(oflags & kNfThick))
			{
				const MaterialGpu& hm = a.materials[MaterialOfTri(a, h.tri)];
				if (hm.transmission > 0.0f && hm.thickness > 0.0f && pMaxBounces > 0)
				{
					SkyState s; s.att = v3(1.0f); s.targetAux = g; s.prev = ro; s.ignore = h.tri; s.start = ro; s.ior = o->envIor; s.dir = rd;
					s.jAndMax = 0u | (pMaxBounces << 16);
					SkyPush(a, 0, s);
				}
			}
		}
	};

	// ---- sky: one TraceSky iteration for every walk in flight ------------------------------------------------------
	struct SkyKernel
	{
		IntegratorArgs a; uint32_t q;            // states in sky[q], survivors go to sky[q^1]
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			SkyState s = a.sky[q][i];
			const Hit h = a.skyHits[i];
			V3 att;
			if (SkyAdvance(a, s, h, att)) SkyPush(a, q ^ 1u, s);
			else SkyFinish(a, s.targetAux, att);
		}
	};

#ifndef SPT_GATHER_BLOCK
#define SPT_GATHER_BLOCK 4
#endif
	constexpr uint32_t kGatherBlock = SPT_GATHER_BLOCK;      // RayAux entries fetched per round of GatherKernel's loops

	// ---- gather: replay :690-871 over one activation's rays, bottom-up ----------------------------------------------
	struct GatherKernel
	{
		IntegratorArgs a;
		SPT_KERNEL_BODY void operator()(uint32_t ri) const
		{
			NodeRec* rec = a.recs + ri;
			const uint32_t flags = rec->flags;
			V3 res;
			if (flags & kNfDone) res = rec->result;
			else if (flags & kNfTail) res = rec->child != kNone ? a.recs[rec->child].result : v3(0.0f);
			else
			{
				res = v3(0.0f);
				uint32_t g = rec->auxBase;
				const uint32_t nLights = a.numLights, nA = rec->nA, nS = rec->nS;
				// The loops below read the record's entries one dependent load after the other (a light's weight only if its status byte
				// says unblocked, ...): six DRAM round trips for the three rays of a deep activation.  Everything they will touch is known
				// now, so it is requested now: the first entries, their status bytes and the parent's entry this result goes to.
				{
					const uint32_t total = nLights + ((flags & kNfThick) ? 0u : nA) + nS;
					prefetch(a.status + g);
					for (uint32_t j = 0; j < total && j < 8u; j += 4u) prefetch(a.aux + g + j);          // 4 entries per 128-byte line
					if (total) prefetch(a.aux + g + (total < 8u ? total : 8u) - 1u);
					if (!(flags & kNfLevel0) && rec->parentAux != kNone) prefetch(a.aux + rec->parentAux);
				}      // an expanded activation has one shadow ray per light (:691-705)
				bool blocked = false;                                                     // stale hitLight (:694)
				for (uint32_t j = 0; j < nLights; j++, g++)
				{
					if (a.status[g]) blocked = true;                                       // :699
					if (!blocked) { const RayAux x = a.aux[g]; res = res + v3(x.a0, x.a1, x.a2); }
				}
				if (flags & kNfAmbientOn)
				{
					const float pdfHemisphere = 1.0f / (kPiSailor * 2.0f);
					V3 amb1 = v3(0.0f);
					if (!(flags & kNfThick))
					{
						// entries are fetched kGatherBlock at a time (records AND status bytes, unconditionally) before any of them is looked at:
						// the loop is bound by memory latency, and a status load that waits for its record's tag doubles the chain
						for (uint32_t k0 = 0; k0 < nA; k0 += kGatherBlock)
						{
							RayAux xs[kGatherBlock]; uint8_t hitSomething[kGatherBlock];
#pragma unroll
							for (uint32_t j = 0; j < kGatherBlock; j++) if (k0 + j < nA) { xs[j] = a.aux[g + j]; hitSomething[j] = a.status[g + j]; }
#pragma unroll
							for (uint32_t j = 0; j < kGatherBlock; j++)
							{
								if (k0 + j >= nA) break;
								const RayAux& x = xs[j];
								if ((x.tag & kRkMask) != kRkHemi) continue;                                                   // dead ray (maxBounces == 0)
								if (x.tag & kRsSky2) amb1 = amb1 + v3(x.a0, x.a1, x.a2);                                      // a sky walk ended (:730-735)
								else if (!hitSomething[j]) amb1 = amb1 + glm_clamp((v3(x.a0, x.a1, x.a2) * a.ambient * x.a3) / pdfHemisphere, 0.0f, 10.0f);   // miss: att = 1 (:587-590)
							}
							g += (nA - k0 < kGatherBlock) ? nA - k0 : kGatherBlock;
						}
					}
					amb1 = amb1 / (float)nA;                                              // :739
					V3 amb2 = v3(0.0f), indirect = v3(0.0f); float avgPdf = 0.0f, cnt = 0.0f;
					const bool bouncesLeft = (rec->bounces & 0xFFFFu) > 0;
					for (uint32_t i0 = 0; i0 < nS; i0 += kGatherBlock)
					{
						RayAux xs[kGatherBlock]; uint8_t hitSomething[kGatherBlock];
#pragma unroll
						for (uint32_t j = 0; j < kGatherBlock; j++) if (i0 + j < nS) { xs[j] = a.aux[g + j]; hitSomething[j] = a.status[g + j]; }
#pragma unroll
						for (uint32_t j = 0; j < kGatherBlock; j++)
						{
							if (i0 + j >= nS) break;
							const RayAux& x = xs[j];
							if (x.tag & kRsSkipped) continue;
							V3 value = v3(x.a0, x.a1, x.a2);
							if (!hitSomething[j])                                                                         // miss (:786-796)
							{
								value = glm_clamp(value * a.ambient, 0.0f, 10.0f);
								amb2 = amb2 + value; avgPdf += x.a3; indirect = indirect + value;
							}
							else if (x.tag & kRsHit)                                                                      // :816-832, value = clamp(term*att*L)
							{
								indirect = indirect + value;
								if (x.tag & kRsSky2) { amb2 = amb2 + value * v3(x.b0, x.b1, x.b2); avgPdf += x.a3; }
							}
							else if (bouncesLeft)                                                                         // boolean ray below the 0.01 cut: raytraced = 0 (:809,:816)
								indirect = indirect + glm_clamp(value * v3(0.0f), 0.0f, 10.0f);
							cnt += 1.0f;                                                                                  // :835
						}
						g += (nS - i0 < kGatherBlock) ? nS - i0 : kGatherBlock;
					}
					amb2 = amb2 / cnt; avgPdf = avgPdf / cnt;                             // :838-839
					const V3 ambient = amb1 + amb2;
					if (ambient.x + ambient.y + ambient.z > 0.0f)
					{
						res = res + (amb1 * PowerHeuristic((int32_t)nA, pdfHemisphere, (int32_t)cnt, avgPdf) +
							amb2 * PowerHeuristic((int32_t)cnt, avgPdf, (int32_t)nA, pdfHemisphere));
					}
					if (cnt > 0.0f) res = res + (indirect / cnt);
				}
				res = res + rec->emissive;                                                // :855
				if ((rec->bounces & 0xFFFFu) > 0 && (flags & kNfAlpha))                   // :858-871
				{
					const float al = rec->alpha;
					const V3 cr = rec->child != kNone ? a.recs[rec->child].result : v3(0.0f);
					res = res * al + cr * (1.0f - al);
				}
			}
			rec->result = res;
			if (flags & kNfLevel0)
			{
				// accumulator += Raytrace(...) (:466): one slot per (pixel, sample); summed in order by ResolveKernel
				const uint32_t pixel = rec->parent, sample = rec->parentAux;
				const uint32_t x = pixel % a.cam.width, y = pixel / a.cam.width;
				const size_t idx = ((size_t)(y - a.rowBegin) * a.cam.width + x) * (a.msEnd - a.msBegin) + (sample - a.msBegin);
				a.sampleBuf[idx * 3] = res.x; a.sampleBuf[idx * 3 + 1] = res.y; a.sampleBuf[idx * 3 + 2] = res.z;
			}
			else if (rec->parentAux != kNone)
			{
				RayAux* px = a.aux + rec->parentAux;                                      // :816  value = clamp(term*att*raytraced)
				const V3 value = glm_clamp(v3(px->a0, px->a1, px->a2) * res, 0.0f, 10.0f);
				px->a0 = value.x; px->a1 = value.y; px->a2 = value.z;
			}
		}
	};

	// ---- level bookkeeping (one thread) -----------------------------------------------------------------------------
	struct BeginBatchKernel      // level 0 = the seeded first hits
	{
		BatchCounters* c; uint32_t count;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			c->recAlloc = count; c->auxAlloc = 0; c->overflow = 0; c->skyCount[0] = c->skyCount[1] = 0; c->zero = 0; c->fanEntries = 0; c->fanThreads[0] = c->fanThreads[1] = 0; c->slowCount = 0; c->rays = 0; c->fanSamples = 0;
			c->level[0].recBegin = 0; c->level[0].recEnd = count; c->level[0].rayCount = 0; c->level[0].auxBase = 0;
		}
	};
	struct NextLevelKernel       // after classify(level): the records allocated since are level+1
	{
		BatchCounters* c; uint32_t level; uint32_t recCap;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			const LevelInfo cur = c->level[level];
			LevelInfo nx; nx.recBegin = cur.recEnd; nx.recEnd = c->recAlloc < recCap ? c->recAlloc : recCap; nx.rayCount = 0; nx.auxBase = cur.auxBase + cur.rayCount;
			c->level[level + 1] = nx;
			c->auxAlloc = nx.auxBase;
			c->rays += cur.rayCount;
			c->fanEntries = 0; c->fanThreads[0] = c->fanThreads[1] = 0; c->slowCount = 0;
		}
	};
	// An arena overflowed while this level was expanded: some activations left their reserved queue entries unwritten, so nothing of the
	// level may be traced or classified (stale entries would send ClassifyKernel after record indices of another batch).  The batch is
	// redone smaller by the host (render.cuh); until it notices, the remaining launches of the batch find empty ranges.
	struct OverflowGuardKernel
	{
		BatchCounters* c; uint32_t level;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			if (!c->overflow) return;
			c->level[level].rayCount = 0; c->slowCount = 0; c->skyCount[0] = c->skyCount[1] = 0;
		}
	};
	struct SkySwapKernel         // after a sky iteration: queue q is consumed
	{
		BatchCounters* c; uint32_t q;
		SPT_KERNEL_BODY void operator()(uint32_t) const { c->rays += c->skyCount[q]; c->skyCount[q] = 0; }
	};

	// accumulator / msaa with the row flip of PathTracer.cpp:449,468-469
	// Progressive renders (SailorPt_RenderProgressive) keep the UN-normalised sum in `running` between passes: a pass continues the
	// same left-to-right chain of additions the one-shot render performs, so the final image has the same bits.
	struct ResolveKernel
	{
		const float* sampleBuf; float* image; uint32_t width, height, rowBegin, rowEnd, numSamples, msaa;
		float* running; uint32_t runningValid;      // running: optional [rows*width*3] sum carried across passes; valid = continue from it
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const uint32_t x = i % width, yb = i / width, y = rowBegin + yb;
			V3 acc = v3(0.0f);
			if (running && runningValid) acc = v3(running[(size_t)i * 3], running[(size_t)i * 3 + 1], running[(size_t)i * 3 + 2]);
			const float* s = sampleBuf + (size_t)i * numSamples * 3;
			for (uint32_t k = 0; k < numSamples; k++) acc = acc + v3(s[k * 3], s[k * 3 + 1], s[k * 3 + 2]);
			if (running) { running[(size_t)i * 3] = acc.x; running[(size_t)i * 3 + 1] = acc.y; running[(size_t)i * 3 + 2] = acc.z; }
			const V3 res = acc / (float)msaa;
			const size_t o = ((size_t)(height - y - 1) * width + x) * 3;
			image[o] = res.x; image[o + 1] = res.y; image[o + 2] = res.z;
		}
	};
}
