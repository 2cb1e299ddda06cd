// lighting.cuh — the BSDF of the reference path tracer as device functions.
//
// Restates LightingModel (reference Raytracing/LightingModel.cpp) expression by expression, in fp32, with the
// reference's own constants (two different pi values are in play: glm::pi<float>() and Math::Pi = 3.1415926f),
// clamps (max(r*r,1e-3), max(NdotH,1e-4) ...) and early-outs.  The integer powers the reference writes as pow(x, 5.0f),
// pow(x, 2.0f), pow(x, 4.0f) are evaluated by multiplication (within 2 ulp of glibc's correctly rounded pow, closer to
// it than CUDA's powf and an order of magnitude cheaper); the vec3 / float that ends CalculateBRDF / CalculateBTDF multiplies by
// one reciprocal (<= 1.5 ulp, see ScaleByReciprocal); expf/logf/sinf/cosf are the CUDA libm ones, which differ from
// glibc in the last ulp or two.  All of that is inside the converged-image tolerance and is checked per function against
// the reference's own LightingModel by SailorPt_EvalLighting (tests: check_lighting, rtol 2e-4).
#pragma once
#include "backend.h"

namespace spt
{
	// LightingModel::SampledData (LightingModel.h:20-30)
	struct SampledData
	{
		V4 baseColor; V3 orm; V3 emissive; V3 normal;
		float ior, thickness, transmission; bool opaque;
	};

	// vec3 / float at the end of CalculateBRDF / CalculateBTDF (:119,:157-158): the reference divides each component; here the
	// reciprocal is taken once (one IEEE division) and multiplied in, <= 1.5 ulp from the component-wise quotients.  Three IEEE
	// divisions per vector were 22 % of the hemisphere pass of FanOutKernel, a third of it in the divider's slow path, which a ZERO
	// numerator (a base colour with zero channels) always takes; a multiplication has no slow path (profiles/r01g_SUMMARY.md).
	SPT_HD V3 ScaleByReciprocal(V3 a, float s)
	{
#if defined(SPT_BSDF_EXACT_DIV)
		return a / s;
#else
		return a * (1.0f / s);
#endif
	}

	SPT_HD float DistributionGGX(V3 N, V3 H, float roughness)                 // LightingModel.cpp:28-40
	{
		const float a = std_max(roughness * roughness, 0.001f);
		const float a2 = a * a;
		const float NdotH = std_max(dot(N, H), 0.0001f);
		const float NdotH2 = NdotH * NdotH;
		float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
		denom = kPiGlm * denom * denom;
		return a2 / denom;
	}

	SPT_HD V3 FresnelSchlick(float cosTheta, V3 F0)                            // :42-45
	{
		const float x = 1.0f - cosTheta, x2 = x * x;
		return F0 + (1.0f - F0) * (x2 * x2 * x);                                   // pow(1 - cosTheta, 5.0f)
	}

	SPT_HD float GeometrySchlickGGX(float NdotV, float roughness)             // :47-52
	{
		const float k = (roughness * roughness) / 2.0f;
		const float denom = NdotV * (1.0f - k) + k;
		return NdotV / denom;
	}

	SPT_HD V3 CalculateBTDF(V3 V, V3 N, V3 lDir, const SampledData& s)        // :81-121
	{
		const V3 L = lDir + 2.0f * N * dot(-lDir, N);
		const float roughness = s.orm.y, metallic = s.orm.z, transmission = s.transmission;
		if (transmission <= 0.0f) return v3(0.0f);
		const float nDotL = fabsf(dot(N, L)), nDotV = fabsf(dot(N, V));
		const V3 F0 = v3(0.04f);
		const V3 H = normalize(V + L);
		const float NDF = DistributionGGX(N, H, roughness);
		const V3 F = FresnelSchlick(std_max(fabsf(dot(H, V)), 0.0f), F0);
		const float G = GeometrySchlickGGX(nDotL, roughness) * GeometrySchlickGGX(nDotV, roughness);
		const V3 base = v3(s.baseColor.x, s.baseColor.y, s.baseColor.z);
		const V3 kT = (1.0f - F) * transmission * (1.0f - metallic) * base;
		if (nDotL < 0.0f || nDotV < 0.0f) return v3(0.0f);
		const float denominator = (4.0f * std_max(nDotV, 0.0f) * std_max(nDotL, 0.0f)) + 0.001f;
		return ScaleByReciprocal(kT * NDF * G, denominator);
	}

	SPT_HD V3 CalculateBRDF(V3 V, V3 N, V3 L, const SampledData& s)           // :123-160
	{
		const float roughness = s.orm.y, metallic = s.orm.z;
		const float nDotL = dot(N, L), nDotV = dot(N, V);
		const V3 base = v3(s.baseColor.x, s.baseColor.y, s.baseColor.z);
		const V3 F0 = glm_mix(v3(0.04f), base, metallic);
		const V3 H = normalize(V + L);
		const float NDF = DistributionGGX(N, H, roughness);
		const V3 F = FresnelSchlick(std_max(dot(H, V), 0.0f), F0);
		const float G = GeometrySchlickGGX(nDotL, roughness) * GeometrySchlickGGX(nDotV, roughness);
		V3 kD = v3(1.0f) - F;
		kD = kD * (1.0f - metallic);
		kD = kD * (1.0f - s.transmission);
		if (nDotL < 0.0f || nDotV < 0.0f) return v3(0.0f);
		const float denominator = (4.0f * std_max(nDotV, 0.0f) * std_max(nDotL, 0.0f)) + 0.001f;
		const V3 specular = ScaleByReciprocal(F * NDF * G, denominator);
		const V3 diffuse = ScaleByReciprocal(kD * base, kPiGlm);
		return diffuse + specular;
	}

	// shared tail of the four ImportanceSample* functions (:162-249)
	SPT_HD V3 ToWorld(float sinTheta, float cosTheta, float phi, V3 n)
	{
		float sp, cp;
		sincosf(phi, &sp, &cp);
		const V3 h = v3(sinTheta * cp, sinTheta * sp, cosTheta);
		const V3 up = fabsf(n.z) < 0.999f ? v3(0.0f, 0.0f, 1.0f) : v3(1.0f, 0.0f, 0.0f);
		const V3 tangent = normalize(cross(up, n));
		const V3 bitangent = cross(n, tangent);
		return normalize(tangent * h.x + bitangent * h.y + n * h.z);
	}

	SPT_HD V3 ImportanceSampleBeckmann(V2 Xi, float roughness, V3 n)          // :162-182
	{
		const float alpha = std_max(roughness * roughness, 0.001f);
		const float phi = 2.0f * kPiGlm * Xi.x;
		const float tanTheta2 = -alpha * alpha * logf(1.0f - Xi.y);
		const float cosTheta = 1.0f / sqrtf(1.0f + tanTheta2);
		const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
		return ToWorld(sinTheta, cosTheta, phi, n);
	}

	SPT_HD V3 ImportanceSampleGGX(V2 Xi, float roughness, V3 n)               // :184-204
	{
		const float a = std_max(roughness * roughness, 0.001f);
		const float phi = 2.0f * kPiGlm * Xi.x;
		const float cosTheta = sqrtf((1.0f - Xi.y) / (1.0f + (a * a - 1.0f) * Xi.y));
		const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
		return ToWorld(sinTheta, cosTheta, phi, n);
	}

	SPT_HD float PowerHeuristic(int32_t nf, float fPdf, int32_t ng, float gPdf)  // :206-211
	{
		const float f = (float)nf * fPdf, g = (float)ng * gPdf;
		return (f * f) / (f * f + g * g);
	}

	SPT_HD V3 ImportanceSampleLambert(V2 Xi, V3 n)                            // :213-230
	{
		const float phi = 2.0f * kPiGlm * Xi.x;
		const float cosTheta = sqrtf(1.0f - Xi.y);
		const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
		return ToWorld(sinTheta, cosTheta, phi, n);
	}

	SPT_HD V3 ImportanceSampleHemisphere(V2 Xi, V3 n)                         // :232-249
	{
		const float phi = 2.0f * kPiGlm * Xi.x;
		const float cosTheta = 1.0f - Xi.y;
		const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
		return ToWorld(sinTheta, cosTheta, phi, n);
	}

	SPT_HD float GGX_PDF(V3 N, V3 H, V3 V, float roughness)                   // :251-263
	{
		const float a = std_max(roughness * roughness, 0.001f);
		const float NdotH = std_max(dot(N, H), 0.001f);
		const float VdotH = std_max(dot(V, H), 0.001f);
		const float dd = NdotH * NdotH * (a * a - 1.0f) + 1.0f;
		const float D = (a * a) / (kPiGlm * (dd * dd));                              // pow(., 2.0f)
		return D * NdotH / (4.0f * VdotH);
	}

	SPT_HD float Beckmann_PDF(V3 N, V3 H, V3 V, float roughness)              // :265-283
	{
		const float alpha = std_max(roughness * roughness, 0.001f);
		const float NdotH = std_max(dot(N, H), 0.001f);
		const float VdotH = std_max(dot(V, H), 0.001f);
		const float tanThetaH = sqrtf(1.0f - NdotH * NdotH) / NdotH;
		const float tanThetaHSquared = tanThetaH * tanThetaH;
		const float alphaSquared = alpha * alpha;
		const float n2 = NdotH * NdotH;
		const float D = expf(-tanThetaHSquared / alphaSquared) / (kPiGlm * alphaSquared * (n2 * n2));   // pow(NdotH, 4.0f)
		return D * NdotH / (4.0f * VdotH);
	}

	SPT_HD V3 CalculateRefraction(V3 rayDirection, V3 N, float fromIor, float toIor)   // :285-300
	{
		const float eta = fromIor / toIor;
		const float cosi = -dot(N, rayDirection);
		const float k = 1 - eta * eta * (1 - cosi * cosi);
		if (k < 0) return v3(0.0f);
		return normalize(eta * rayDirection + (eta * cosi - sqrtf(k)) * N);
	}

	// LightingModel::Sample (:302-386).  randSpecular / randTransmission replace the two glm::linearRand(0,1) draws
	// of :310-311 (drawn by the caller in that order; the second draw happens only when bHasTransmission, like the
	// short-circuit && in the reference).
	SPT_HD bool SampleBsdf(const SampledData& s, V3 N, V3 V, float fromIor, float toIor, V3& outTerm, float& outPdf,
		bool& outTransmissionRay, V3& inOutDirection, V2 Xi, float randSpecular, float randTransmission)
	{
		const bool bFullMetallic = s.orm.z == 1.0f;
		const bool bMirror = bFullMetallic && s.orm.y <= 0.001f;
		const bool bHasTransmission = !bFullMetallic && s.transmission > 0.0f;
		const bool bIsThickVolume = bHasTransmission && s.thickness > 0.0f;
		const bool bSpecular = bMirror || randSpecular > 0.5f;
		outTransmissionRay = bHasTransmission && (randTransmission > 0.5f);
		const float importanceRoughness = bSpecular ? s.orm.y : 1.0f;
		const bool bBeckmann = importanceRoughness < 0.2f;
		// The three lobes (ImportanceSampleBeckmann / GGX / Lambert, :162-230) differ only in cos(theta); phi, sin(theta) and the change
		// of basis are the same expressions.  The 50/50 lobe pick splits a warp in two, so only cos(theta) is computed inside the
		// branch and the expensive common tail (sincos, two normalisations) runs once with all lanes: same values, half the issue slots.
		float cosTheta;
		if (bSpecular)
		{
			if (bBeckmann)
			{
				const float alpha = std_max(s.orm.y * s.orm.y, 0.001f);
				const float tanTheta2 = -alpha * alpha * logf(1.0f - Xi.y);
				cosTheta = 1.0f / sqrtf(1.0f + tanTheta2);
			}
			else
			{
				const float a = std_max(s.orm.y * s.orm.y, 0.001f);
				cosTheta = sqrtf((1.0f - Xi.y) / (1.0f + (a * a - 1.0f) * Xi.y));
			}
		}
		else cosTheta = sqrtf(1.0f - Xi.y);
		const V3 H = ToWorld(sqrtf(1.0f - cosTheta * cosTheta), cosTheta, 2.0f * kPiGlm * Xi.x, N);

		if (eq0(inOutDirection))
		{
			inOutDirection = 2.0f * dot(V, H) * H - V;
			if (outTransmissionRay)
			{
				inOutDirection = inOutDirection + 2.0f * N * dot(-inOutDirection, N);
				if (bIsThickVolume)
				{
					inOutDirection = CalculateRefraction(inOutDirection, N, fromIor, toIor);
					if (eq0(inOutDirection)) return false;
					outTerm = v3(1.0f);
					return true;
				}
			}
		}
		const float pdfSpec = bBeckmann ? Beckmann_PDF(N, H, V, s.orm.y) : GGX_PDF(N, H, V, s.orm.y);
		const float pdfLambert = fabsf(dot(inOutDirection, N)) / kPiSailor;
		outPdf = bMirror ? pdfSpec : ((pdfSpec + pdfLambert) * 0.5f);
		if (bHasTransmission) outPdf *= 0.5f;
		if (!(outPdf != outPdf) && outPdf > 0.0001f)
		{
			const float angle = fabsf(dot(inOutDirection, N));
			const V3 term = outTransmissionRay ? CalculateBTDF(V, N, inOutDirection, s) : CalculateBRDF(V, N, inOutDirection, s);
			const float weight = 1.0f / outPdf;
			outTerm = weight * term * angle;
			return true;
		}
		return false;
	}

	// SailorPt_EvalLighting: 24 floats in, 28 floats out per record (layout in tests/test_lighting.py)
	struct EvalLightingKernel
	{
		const float* in; float* out;
		SPT_KERNEL_BODY void operator()(uint32_t i) const
		{
			const float* r = in + (size_t)i * 24;
			float* o = out + (size_t)i * 28;
			SampledData sd;
			sd.baseColor = v4(r[0], r[1], r[2], r[3]); sd.orm = v3(r[4], r[5], r[6]); sd.emissive = v3(r[7], r[8], r[9]);
			const V3 N = v3(r[10], r[11], r[12]), V = v3(r[13], r[14], r[15]), L = v3(r[16], r[17], r[18]);
			sd.ior = r[19]; sd.thickness = r[20]; sd.transmission = r[21]; sd.opaque = true; sd.normal = v3(0.0f, 0.0f, 1.0f);
			const V2 Xi = v2(r[22], r[23]);
			const float rough = sd.orm.y;
			const V3 H = normalize(V + L);
			const V3 brdf = CalculateBRDF(V, N, L, sd), btdf = CalculateBTDF(V, N, L, sd);
			o[0] = brdf.x; o[1] = brdf.y; o[2] = brdf.z; o[3] = btdf.x; o[4] = btdf.y; o[5] = btdf.z;
			o[6] = DistributionGGX(N, H, rough);
			o[7] = GeometrySchlickGGX(dot(N, L), rough);
			o[8] = GGX_PDF(N, H, V, rough);
			o[9] = Beckmann_PDF(N, H, V, rough);
			const V3 s0 = ImportanceSampleGGX(Xi, rough, N), s1 = ImportanceSampleBeckmann(Xi, rough, N), s2 = ImportanceSampleLambert(Xi, N), s3 = ImportanceSampleHemisphere(Xi, N);
			o[10] = s0.x; o[11] = s0.y; o[12] = s0.z; o[13] = s1.x; o[14] = s1.y; o[15] = s1.z;
			o[16] = s2.x; o[17] = s2.y; o[18] = s2.z; o[19] = s3.x; o[20] = s3.y; o[21] = s3.z;
			o[22] = PowerHeuristic(3, o[8], 2, o[9]);
			const V3 rf = CalculateRefraction(-V, N, 1.0f, sd.ior);
			o[23] = rf.x; o[24] = rf.y; o[25] = rf.z;
			o[26] = FresnelSchlick(std_max(dot(H, V), 0.0f), v3(0.04f)).x;
			o[27] = 0.0f;
		}
	};
}
