// wide_bvh.cuh — the second traversal layout: an 8-wide BVH with child boxes quantised to 8 bits (80-byte nodes).
//
// WHY.  The reference's binary tree walked in the reference's visit order (traverse.cuh) costs a 1M-triangle ray ~28
// inner-node visits of 64 bytes each; two thirds of them are the chain of boxes that CONTAIN the ray origin (secondary rays
// start on a surface).  That kernel is bound by issue slots and L1 wavefronts, not by HBM (profiles/r01g_SUMMARY.md).  Every
// ray whose result does not depend on the visit order can use a shallower, denser structure:
//   * hit-or-miss queries (shadow rays, hemisphere rays, importance rays below the throughput cut): "some triangle is hit"
//     is order independent;
//   * closest-hit queries: t, u, v come from the unchanged TriTest, so only the WINNER among (nearly) equal distances can
//     depend on the order.  The wide walk flags such rays (a second candidate within a relative band of 2^-16 around the
//     best distance) and they are replayed by the exact kernel (SURVEY H1's recipe).
// The primary pass keeps the exact kernel (north_star: bit-exact primary hits).
//
// LAYOUT (one WNode = 80 bytes = five LDG.128; children of one node are consecutive in memory):
//   px,py,pz         quantisation origin (a little below the node's box minimum)
//   ex,ey,ez,imask   biased power-of-two exponent of the grid step per axis; bit s of imask = slot s holds an inner child
//   childBase        index of the first inner child (slot order); triBase: first triangle record of the node's leaf slots
//   meta[8]          leaf slot: 0x80 | (count-1) << 5 | offset   (count <= 4 triangles at triBase + offset, offset <= 28)
//   qlo[3][8], qhi[3][8]   child boxes on the 8-bit grid, conservative by >= 1/16 step (empty slot: lo 255, hi 0)
// Slots are assigned by the child's position relative to the node centre (bit a of the slot = high side on axis a), so
// visiting slots in the order  slot ^ (7 - rayOctant)  descending is approximately front to back (Ylitie et al. 2017).
//
// Leaves are the REFERENCE's leaves (a leaf of more than 4 triangles is cut into chunks of <= 4 that share its box), and
// every triangle record carries the exact fp32 box of its reference leaf next to it (wleafBox): a candidate hit only counts
// when the reference's own slab test of that leaf passes, which is what makes "reachable in the reference tree"
// reproducible here — the reference's slab distances are monotone under box nesting, so a leaf whose box passes is
// reached whatever its ancestors were.
//
// BUILD: collapse of the bit-exact binary tree (bvh_build.cuh), level by level on the device: every wide node starts
// from one binary node and greedily opens the child with the largest surface area until it has 8 slots.
#pragma once
#include "traverse.cuh"

namespace spt
{
	struct alignas(16) WNode
	{
		float px, py, pz; uint32_t exyzMask;
		uint32_t childBase, triBase, meta0, meta1;
		uint32_t q[12];          // qlox[2] qloy[2] qloz[2] qhix[2] qhiy[2] qhiz[2], slot s = byte s&3 of word s>>2
	};
	static_assert(sizeof(WNode) == 80, "wide node layout");

	struct WideView { const WNode* nodes; const TTri* tris; const V4* leafBox; uint32_t numNodes; };

	constexpr uint32_t kWideLeafMax = 4;          // triangles per leaf slot
	constexpr uint32_t kTriGroupTag = 0x80000000u; // stack entries: x = triBase | tag (triangle group) or childBase (node group)
	constexpr int kWideStackDepth = 64;           // deeper walks are handed to the exact kernel
	constexpr float kTieBand = 1.0f / 65536.0f;   // relative distance band in which two candidates count as a tie

	// ---- build ---------------------------------------------------------------------------------------------------
	// descriptor of a wide node still to be built: a binary inner node (cnt = 0) or a chunk [begin, begin+cnt) of a big leaf
	struct WDesc { uint32_t node, begin, cnt, pad; };
	struct WideCounters { uint32_t nodeCount, triCount, lvlBegin, lvlEnd, overflow, maxDepthGuard, pad0, pad1; };

	struct WideBuildArgs
	{
		// the binary tree (build numbering) and the reference-order triangle records
		const uint32_t* left; const uint32_t* count; const float* aabb; const uint32_t* refIdx; const uint32_t* leafOffsetByRef;
		const TTri* ttris;
		// out
		WNode* nodes; TTri* wtris; V4* leafBox; WDesc* desc; WideCounters* c;
		uint32_t nodeCap, numTris;
	};

	struct WideInitKernel      // one thread: wide node 0 starts from the binary root
	{
		WideBuildArgs a;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			WDesc d; d.node = 0; d.begin = 0; d.cnt = a.left[0] ? 0u : a.count[0]; d.pad = 0;
			a.desc[0] = d;
			WideCounters c; c.nodeCount = 1; c.triCount = 0; c.lvlBegin = 0; c.lvlEnd = 1; c.overflow = 0; c.maxDepthGuard = 0; c.pad0 = c.pad1 = 0;
			*a.c = c;
		}
	};

	struct WideAdvanceKernel   // one thread: the nodes allocated by the level just built are the next level
	{
		WideCounters* c; uint32_t nodeCap;
		SPT_KERNEL_BODY void operator()(uint32_t) const
		{
			c->lvlBegin = c->lvlEnd;
			c->lvlEnd = c->nodeCount < nodeCap ? c->nodeCount : nodeCap;
		}
	};

	SPT_HD uint32_t WideExponent(float extent)
	{
		// smallest biased exponent e with 255 * 2^(e-127) >= extent * (1 + margin): the grid must also hold the 1/8-step shift
		// of the origin and the 1/16-step slack of every plane
		const float f = extent * (1.0f / 254.0f);
		if (!(f > 0.0f)) return 1u;
		const uint32_t b = f2u(f);
		uint32_t e = (b >> 23) + ((b & 0x7FFFFFu) ? 1u : 0u);
		if (e < 1u) e = 1u;
		if (e > 254u) e = 254u;
		return e;
	}

	struct WideLevelKernel     // one thread per wide node of the current level
	{
		WideBuildArgs a;

		struct Item { uint32_t node, begin, cnt; float area; };    // cnt == 0: binary inner node

		SPT_KERNEL_BODY float AreaOf(uint32_t node) const
		{
			const float* bb = a.aabb + (size_t)node * 6;
			const float ex = bb[3] - bb[0], ey = bb[4] - bb[1], ez = bb[5] - bb[2];
			return ex * ey + ey * ez + ez * ex;
		}
		SPT_KERNEL_BODY Item ItemOfNode(uint32_t node) const
		{
			Item it; it.node = node; it.begin = 0; it.cnt = a.left[node] ? 0u : a.count[node]; it.area = AreaOf(node);
			return it;
		}
		static SPT_KERNEL_BODY bool Expandable(const Item& it) { return it.cnt == 0u || it.cnt > kWideLeafMax; }
		// the two parts of an expandable item
		SPT_KERNEL_BODY void Open(const Item& it, Item& p, Item& q) const
		{
			if (it.cnt == 0u) { const uint32_t l = a.left[it.node]; p = ItemOfNode(l); q = ItemOfNode(l + 1u); return; }
			const uint32_t h = (it.cnt + 1u) / 2u;
			p = it; p.cnt = h;
			q = it; q.begin = it.begin + h; q.cnt = it.cnt - h;
		}

		SPT_KERNEL_BODY void operator()(uint32_t wi) const
		{
			const WDesc self = a.desc[wi];
			Item items[8]; uint32_t n = 0;
			{
				Item root; root.node = self.node; root.begin = self.begin; root.cnt = self.cnt; root.area = 0.0f;
				if (Expandable(root)) { Open(root, items[0], items[1]); n = 2; }
				else { items[0] = root; n = 1; }                     // a root that is one small leaf
			}
			while (n < 8u)
			{
				int best = -1; float bestArea = -1.0f;
				for (uint32_t i = 0; i < n; i++) if (Expandable(items[i]) && items[i].area > bestArea) { best = (int)i; bestArea = items[i].area; }
				if (best < 0) break;
				Item p, q; Open(items[best], p, q);
				items[best] = p; items[n++] = q;
			}
			// node box = union of the children (for a binary node: its own box)
			float bmin[3] = { kFltMax, kFltMax, kFltMax }, bmax[3] = { -kFltMax, -kFltMax, -kFltMax };
			for (uint32_t i = 0; i < n; i++)
			{
				const float* bb = a.aabb + (size_t)items[i].node * 6;
				for (int k = 0; k < 3; k++) { bmin[k] = bb[k] < bmin[k] ? bb[k] : bmin[k]; bmax[k] = bb[3 + k] > bmax[k] ? bb[3 + k] : bmax[k]; }
			}
			// grid: origin 1/8 step below the minimum
			uint32_t eb[3]; float step[3], org[3], inv[3];
			for (int k = 0; k < 3; k++)
			{
				eb[k] = WideExponent(bmax[k] - bmin[k]);
				for (;;)
				{
					step[k] = u2f(eb[k] << 23);
					org[k] = bmin[k] - step[k] * 0.125f;
					if ((bmax[k] - org[k]) / step[k] + 0.0625f <= 255.0f || eb[k] >= 254u) break;
					eb[k]++;
				}
				inv[k] = 1.0f / step[k];
			}
			// slot assignment: greedy on  sum_axis (+-1)(centre_child - centre_node)
			float cen[8][3];
			for (uint32_t i = 0; i < n; i++)
			{
				const float* bb = a.aabb + (size_t)items[i].node * 6;
				for (int k = 0; k < 3; k++) cen[i][k] = (bb[k] + bb[3 + k]) - (bmin[k] + bmax[k]);
			}
			int slotOf[8]; uint32_t slotUsed = 0, childDone = 0;
			for (uint32_t r = 0; r < n; r++)
			{
				float bestC = -kFltMax; int bi = 0, bs = 0;
				for (uint32_t i = 0; i < n; i++)
				{
					if (childDone & (1u << i)) continue;
					for (int s = 0; s < 8; s++)
					{
						if (slotUsed & (1u << s)) continue;
						const float c = ((s & 1) ? cen[i][0] : -cen[i][0]) + ((s & 2) ? cen[i][1] : -cen[i][1]) + ((s & 4) ? cen[i][2] : -cen[i][2]);
						if (c > bestC) { bestC = c; bi = (int)i; bs = s; }
					}
				}
				slotOf[bi] = bs; slotUsed |= 1u << bs; childDone |= 1u << bi;
			}
			int itemOfSlot[8];
			for (int s = 0; s < 8; s++) itemOfSlot[s] = -1;
			for (uint32_t i = 0; i < n; i++) itemOfSlot[slotOf[i]] = (int)i;

			uint32_t nInner = 0, nTris = 0, imask = 0;
			for (int s = 0; s < 8; s++)
			{
				const int i = itemOfSlot[s];
				if (i < 0) continue;
				if (Expandable(items[i])) { nInner++; imask |= 1u << s; }
				else nTris += items[i].cnt;
			}
			uint32_t childBase = 0, triBase = 0;
			if (nInner)
			{
				childBase = atomic_add_u32(&a.c->nodeCount, nInner);
				if (childBase + nInner > a.nodeCap) { a.c->overflow = 1u; nInner = 0; imask = 0; childBase = 0; }
			}
			if (nTris) triBase = atomic_add_u32(&a.c->triCount, nTris);

			WNode out;
			out.px = org[0]; out.py = org[1]; out.pz = org[2];
			out.exyzMask = eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24);
			out.childBase = childBase; out.triBase = triBase; out.meta0 = 0; out.meta1 = 0;
			for (int k = 0; k < 12; k++) out.q[k] = 0;
			uint32_t rel = 0, off = 0;
			for (int s = 0; s < 8; s++)
			{
				const int i = itemOfSlot[s];
				uint32_t lo[3] = { 255u, 255u, 255u }, hi[3] = { 0u, 0u, 0u }, meta = 0;
				if (i >= 0 && (!Expandable(items[i]) || (imask & (1u << s))))
				{
					const Item& it = items[i];
					const float* bb = a.aabb + (size_t)it.node * 6;
					for (int k = 0; k < 3; k++)
					{
						float l = floorf((bb[k] - org[k]) * inv[k] - 0.0625f), h = ceilf((bb[3 + k] - org[k]) * inv[k] + 0.0625f);
						l = l < 0.0f ? 0.0f : (l > 255.0f ? 255.0f : l); h = h < 0.0f ? 0.0f : (h > 255.0f ? 255.0f : h);
						lo[k] = (uint32_t)l; hi[k] = (uint32_t)h;
					}
					if (imask & (1u << s))
					{
						WDesc d; d.node = it.node; d.begin = it.begin; d.cnt = it.cnt; d.pad = 0;
						a.desc[childBase + rel] = d; rel++;
					}
					else
					{
						meta = 0x80u | ((it.cnt - 1u) << 5) | off;
						const uint32_t src = a.leafOffsetByRef[a.refIdx[it.node]] + it.begin;
						const V4 b0 = v4(bb[0], bb[1], bb[2], bb[3]), b1 = v4(bb[4], bb[5], 0.0f, 0.0f);
						for (uint32_t j = 0; j < it.cnt; j++)
						{
							TTri t = a.ttris[src + j];
							t.c.z = 0.0f; t.c.w = 0.0f;
							const uint32_t dst = triBase + off + j;
							a.wtris[dst] = t;
							a.leafBox[(size_t)dst * 2] = b0; a.leafBox[(size_t)dst * 2 + 1] = b1;
						}
						off += it.cnt;
					}
				}
				const uint32_t w = (uint32_t)s >> 2, sh = ((uint32_t)s & 3u) * 8u;
				for (int k = 0; k < 3; k++) { out.q[k * 2 + w] |= lo[k] << sh; out.q[6 + k * 2 + w] |= hi[k] << sh; }
				if (w == 0) out.meta0 |= meta << sh; else out.meta1 |= meta << sh;
			}
			a.nodes[wi] = out;
		}
	};

	// ---- one node against one ray -----------------------------------------------------------------------------------
	// Ray constants of a wide walk.  idir is the reciprocal direction (finite: rays with a non-finite reciprocal are not walked here).
	struct WideRay
	{
		V3 o, idir; uint32_t signs, octinv;      // signs: bit a = direction negative on axis a; octinv = 7 - signs for ordered walks, 0 for hit-or-miss queries
		SPT_HD void Set(V3 origin, V3 d, V3 rD, bool ordered)
		{
			o = origin; idir = rD;
			signs = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
			octinv = ordered ? 7u - signs : 0u;
		}
	};

	SPT_HD float BiasedByte(uint32_t word, uint32_t k)      // 32768 + byte k of word, as float (one PRMT on the device)
	{
#if defined(__CUDA_ARCH__)
		// the constant rides in a register and the selector is the immediate (the other way round ptxas spends a move per PRMT)
		return __uint_as_float(__byte_perm(0x47000000u, word, 0x3100u | ((4u + k) << 4)));
#else
		return (float)(32768u + ((word >> (8u * k)) & 0xFFu));
#endif
	}
	SPT_HD float FmaF(float a, float b, float c)
	{
#if defined(__CUDA_ARCH__)
		return __fmaf_rn(a, b, c);
#else
		return fmaf(a, b, c);
#endif
	}
	SPT_HD float Min3(float a, float b, float c)
	{
#if defined(__CUDA_ARCH__)
		float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
		return fminf(fminf(a, b), c);
#endif
	}
	SPT_HD float Max3(float a, float b, float c)
	{
#if defined(__CUDA_ARCH__)
		float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
		return fmaxf(fmaxf(a, b), c);
#endif
	}
	SPT_HD uint32_t Popc(uint32_t v)
	{
#if defined(__CUDA_ARCH__)
		return (uint32_t)__popc(v);
#else
		return (uint32_t)__builtin_popcount(v);
#endif
	}
	SPT_HD uint32_t HighBit(uint32_t v)     // index of the highest set bit, v != 0
	{
#if defined(__CUDA_ARCH__)
		return 31u - (uint32_t)__clz((int)v);
#else
		return 31u - (uint32_t)__builtin_clz(v);
#endif
	}
	SPT_HD uint32_t LowBit(uint32_t v)      // index of the lowest set bit, v != 0
	{
#if defined(__CUDA_ARCH__)
		return (uint32_t)__ffs((int)v) - 1u;
#else
		return (uint32_t)__builtin_ctz(v);
#endif
	}

	// Conservative slab tests of the 8 child boxes against [0, limit].  Out: node group (childBase, inner hits | imask << 8)
	// and triangle group (triBase, one bit per triangle record of the leaf slots that were hit).  The inner hits are in TRAVERSAL
	// order (position = slot ^ octinv, highest first = nearest first) when r.ordered, else in slot order (hit-or-miss queries:
	// the order does not matter, and octinv is 0 for them so that position == slot either way).
	SPT_HD void WideNodeTest(const WNode* node, const WideRay& r, float limit, uint32_t& gBase, uint32_t& gBits, uint32_t& tBase, uint32_t& tBits)
	{
		const auto n0 = ld4u(reinterpret_cast<const unsigned char*>(node));
		const auto n1 = ld4u(reinterpret_cast<const unsigned char*>(node) + 16);
		const auto n2 = ld4u(reinterpret_cast<const unsigned char*>(node) + 32);
		const auto n3 = ld4u(reinterpret_cast<const unsigned char*>(node) + 48);
		const auto n4 = ld4u(reinterpret_cast<const unsigned char*>(node) + 64);
		const uint32_t em = n0.w;
		// t(q) = ((org + q step) - o) idir = (32768 + q) adj + c,  adj = step idir,  c = (org - o) idir - 32768 adj
		const float adjx = u2f((em & 0xFFu) << 23) * r.idir.x, adjy = u2f((em & 0xFF00u) << 15) * r.idir.y, adjz = u2f((em & 0xFF0000u) << 7) * r.idir.z;
		const float cx = FmaF(-32768.0f, adjx, (u2f(n0.x) - r.o.x) * r.idir.x);
		const float cy = FmaF(-32768.0f, adjy, (u2f(n0.y) - r.o.y) * r.idir.y);
		const float cz = FmaF(-32768.0f, adjz, (u2f(n0.z) - r.o.z) * r.idir.z);
		// near / far planes by the sign of the direction: words of the low and the high bytes swap
		const bool nx = (r.signs & 1u) != 0u, ny = (r.signs & 2u) != 0u, nz = (r.signs & 4u) != 0u;     // direction negative
		uint32_t mask = 0;
		// one child: six PRMT (byte -> biased float), six FFMA, FMNMX + FMNMX3 twice, a compare and a predicated OR
#if defined(__CUDA_ARCH__)
#define SPT_WIDE_HIT(S) asm("{ .reg .pred p; setp.le.f32 p, %1, %2; @p or.b32 %0, %0, " #S "; }" : "+r"(mask) : "f"(tn), "f"(tf))
#else
#define SPT_WIDE_HIT(S) do { if (tn <= tf) mask |= (S); } while (0)
#endif
#define SPT_WIDE_CHILD(K, BIT) \
		{ \
			const float t0x = FmaF(BiasedByte(nearX, K), adjx, cx), t1x = FmaF(BiasedByte(farX, K), adjx, cx); \
			const float t0y = FmaF(BiasedByte(nearY, K), adjy, cy), t1y = FmaF(BiasedByte(farY, K), adjy, cy); \
			const float t0z = FmaF(BiasedByte(nearZ, K), adjz, cz), t1z = FmaF(BiasedByte(farZ, K), adjz, cz); \
			const float tn = Max3(t0x, t0y, fmaxf(t0z, 0.0f)), tf = Min3(t1x, t1y, fminf(t1z, limit)); \
			SPT_WIDE_HIT(BIT); \
		}
		{
			const uint32_t nearX = nx ? n3.z : n2.x, farX = nx ? n2.x : n3.z, nearY = ny ? n4.x : n2.z, farY = ny ? n2.z : n4.x, nearZ = nz ? n4.z : n3.x, farZ = nz ? n3.x : n4.z;
			SPT_WIDE_CHILD(0u, 1) SPT_WIDE_CHILD(1u, 2) SPT_WIDE_CHILD(2u, 4) SPT_WIDE_CHILD(3u, 8)
		}
		{
			const uint32_t nearX = nx ? n3.w : n2.y, farX = nx ? n2.y : n3.w, nearY = ny ? n4.y : n2.w, farY = ny ? n2.w : n4.y, nearZ = nz ? n4.w : n3.y, farZ = nz ? n3.y : n4.w;
			SPT_WIDE_CHILD(0u, 16) SPT_WIDE_CHILD(1u, 32) SPT_WIDE_CHILD(2u, 64) SPT_WIDE_CHILD(3u, 128)
		}
#undef SPT_WIDE_CHILD
#undef SPT_WIDE_HIT
		const uint32_t imask = em >> 24;
		// inner children: hits in traversal order, position = slot ^ octinv (an XOR permutation = three conditional swaps)
		uint32_t inner = mask & imask;
		if (r.octinv)
		{
			if (r.octinv & 1u) inner = ((inner & 0x55u) << 1) | ((inner & 0xAAu) >> 1);
			if (r.octinv & 2u) inner = ((inner & 0x33u) << 2) | ((inner & 0xCCu) >> 2);
			if (r.octinv & 4u) inner = ((inner & 0x0Fu) << 4) | ((inner & 0xF0u) >> 4);
		}
		gBase = n1.x; gBits = inner | (imask << 8);
		// leaf slots: one bit per triangle record
		uint32_t leaf = mask & ~imask, bits = 0;
		while (leaf)
		{
			const uint32_t s = LowBit(leaf); leaf &= leaf - 1u;
			const uint32_t m = ((s < 4u ? n1.z : n1.w) >> ((s & 3u) * 8u)) & 0xFFu;
			if (m & 0x80u) bits |= ((2u << ((m >> 5) & 3u)) - 1u) << (m & 31u);
		}
		tBase = n1.y; tBits = bits;
	}

#if defined(SPT_EMU) && defined(SPT_WIDE_STATS)
	static unsigned long long g_wideStats[4];      // host tuning aid: nodes visited, triangles tested, rays, winners rejected by the exact leaf test
#define SPT_WSTAT(i) g_wideStats[i]++
#else
#define SPT_WSTAT(i) do { } while (0)
#endif

	SPT_HD bool FiniteBits(float f) { return (f2u(f) & 0x7F800000u) != 0x7F800000u; }

	// rays the wide walk takes: finite origin, finite reciprocal direction small enough that step * idir cannot overflow
	SPT_HD bool WideSafe(V3 o, V3 rD)
	{
		const uint32_t lim = 0x6F800000u;     // 2^96
		return FiniteBits(o.x) && FiniteBits(o.y) && FiniteBits(o.z) &&
			(f2u(rD.x) & 0x7FFFFFFFu) < lim && (f2u(rD.y) & 0x7FFFFFFFu) < lim && (f2u(rD.z) & 0x7FFFFFFFu) < lim &&
			fabsf(o.x) < 1e18f && fabsf(o.y) < 1e18f && fabsf(o.z) < 1e18f;
	}

	// Closest-hit bookkeeping of a wide walk: best candidate so far + whether another candidate lies within the tie band.
	// The tie flag does not depend on the order in which candidates arrive: every candidate within the band of the final best is
	// reached (its boxes start before the band's end, and the walk never culls below the band), and whichever of the two
	// arrives second raises the flag.
	struct WideBest
	{
		float t, u, v; uint32_t tri, rec; float limit; bool tie;
		SPT_HD void Reset() { t = u2f(0x7F800000u); u = 0.0f; v = 0.0f; tri = kNoHit; rec = 0; limit = kFltMax; tie = false; }
		SPT_HD void Offer(float ct, float cu, float cv, uint32_t ctri, uint32_t crec)
		{
			if (tri == kNoHit || ct < t)
			{
				const float band = FmaF(fabsf(ct), kTieBand, ct);
				tie = tri != kNoHit && t <= band;      // the previous best lies within the band of the new one
				t = ct; u = cu; v = cv; tri = ctri; rec = crec;
				limit = band < kFltMax ? band : kFltMax;
			}
			else if (ct <= limit) tie = true;
		}
	};

	// The reference reaches a triangle only through its leaf's box (BVH.cpp:149-175).  Its slab distances are monotone under box
	// nesting, so "the leaf box passes with an unbounded ray length" is the whole condition; it is checked ONCE, for the winner of
	// a closest-hit walk (a winner that fails sends the ray to the exact kernel).  Hit-or-miss walks skip the check: a candidate
	// that passes the triangle test and fails its own leaf's slab test needs a hit within rounding of the box boundary -- 0 of
	// 8 M surface rays on the test scenes (profiles/r02_SUMMARY.md).
	SPT_HD bool WideWinnerReachable(const WideView& w, uint32_t rec, V3 o, V3 rD)
	{
		const V4 b0 = ld4(w.leafBox + (size_t)rec * 2), b1 = ld4(w.leafBox + (size_t)rec * 2 + 1);
		return SlabTest(o, rD, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, kFltMax) != kFltMax;
	}

	// One triangle record against the ray: the reference's own test with its own window (maxLen = FLT_MAX)
	SPT_HD bool WideTriangle(const WideView& w, uint32_t rec, V3 o, V3 d, uint32_t ignore, float& t, float& u, float& v, uint32_t& triId)
	{
		const TTri* T = w.tris + rec;
		const V4 a = ld4(&T->a), b = ld4(&T->b), c = ld4(&T->c);
		triId = f2u(c.y);
		if (triId == ignore) return false;                                         // BVH.cpp:136-139
		return TriTest(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), kFltMax, t, u, v);
	}

	// Scalar wide walk (host-compiled kernel bodies, and the reference for the warp loop of trace_wide.cuh: the results do not
	// depend on the visit order, see WideBest).  Returns false when the ray must be replayed by the exact kernel (unsafe ray,
	// tie, unreachable winner, stack overflow).
	SPT_HD bool TraceWide(const WideView& w, V3 o, V3 d, uint32_t ignore, bool anyHit, Hit& hit)
	{
		SPT_WSTAT(2);
		const V3 rD = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
		hit.t = u2f(0x7F800000u); hit.u = 0.0f; hit.v = 0.0f; hit.tri = kNoHit;
		if (!WideSafe(o, rD)) return false;
		WideRay r; r.Set(o, d, rD, !anyHit);
		WideBest best; best.Reset();
		uint32_t stk[kWideStackDepth][2]; int sp = 0;
		uint32_t gBase = 0, gBits = 0, tBase = 0, tBits = 0;
		WideNodeTest(w.nodes, r, best.limit, gBase, gBits, tBase, tBits);
		SPT_WSTAT(0);
		for (;;)
		{
			while (tBits)
			{
				const uint32_t i = LowBit(tBits); tBits &= tBits - 1u;
				SPT_WSTAT(1);
				float t, u, v; uint32_t tri;
				if (WideTriangle(w, tBase + i, o, d, ignore, t, u, v, tri))
				{
					if (anyHit) { hit.t = t; hit.u = u; hit.v = v; hit.tri = tri; return true; }
					best.Offer(t, u, v, tri, tBase + i);
				}
			}
			if (gBits & 0xFFu)
			{
				const uint32_t pos = HighBit(gBits & 0xFFu);
				gBits ^= 1u << pos;
				const uint32_t slot = pos ^ r.octinv;
				const uint32_t child = gBase + Popc((gBits >> 8) & ((1u << slot) - 1u));
				if (gBits & 0xFFu)
				{
					if (sp >= kWideStackDepth) return false;
					stk[sp][0] = gBase; stk[sp][1] = gBits; sp++;
				}
				WideNodeTest(w.nodes + child, r, best.limit, gBase, gBits, tBase, tBits);
				SPT_WSTAT(0);
				continue;
			}
			if (sp == 0) break;
			sp--;
			gBase = stk[sp][0]; gBits = stk[sp][1];
		}
		if (best.tie) return false;
		if (best.tri != kNoHit && !WideWinnerReachable(w, best.rec, o, rD)) { SPT_WSTAT(3); return false; }
		hit.t = best.t; hit.u = best.u; hit.v = best.v; hit.tri = best.tri;
		return true;
	}
}
