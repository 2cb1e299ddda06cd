// ORACLE (test infrastructure; never linked into or called by the product library).
//
// C-ABI (include/sailor_pt.h) over the REFERENCE's own compiled code: BVH::BuildBVH / BVH::IntersectBVH,
// Math::IntersectRay*, PathTracer::Raytrace / TraceSky / GetMaterialData, LightingModel::*, CombinedSampler2D.
// Only the parts of PathTracer::Run that are commented out in the reference (PathTracer.cpp:78-574) and the Assimp
// front end that is not vendored are restated here, each block citing the lines it follows:
//   camera            PathTracer.cpp:102-153      flattening   MaterialUtils.cpp:64-198 (commented)
//   materials         PathTracer.cpp:164-360      textures     MaterialUtils.h:189-269 (commented)
//   lights            PathTracer.cpp:362-381      viewport     PathTracer.cpp:390-403
//   tile loop         PathTracer.cpp:418-487 (Tasks -> std::thread)      output   PathTracer.cpp:535-565
// glTF semantics that Assimp would have defined are documented in DESIGN.md ("loader contract").
#include "Raytracing/PathTracer.h"
#include "Raytracing/BVH.h"
#include "Raytracing/LightingModel.h"
#include "Raytracing/MaterialUtils.h"
#include "Math/Bounds.h"
#include "Core/Utils.h"
#include "glm/glm/gtc/type_ptr.hpp"
#include "glm/glm/gtc/random.hpp"

#define TINYGLTF_NO_STB_IMAGE_WRITE
#include "tinygltf/tiny_gltf.h"

#include "stb/stb_image_write.h"
#include "sailor_pt.h"

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include <string>
#include <cstring>
#include <cstdio>

using namespace Sailor;
using namespace Sailor::Math;
using namespace Sailor::Raytracing;

extern "C" unsigned char* stbi_load_from_memory(unsigned char const* buffer, int len, int* x, int* y, int* comp, int req_comp);
extern "C" void stbi_image_free(void* p);
extern "C" int stbi_is_hdr_from_memory(unsigned char const* buffer, int len);
extern "C" float* stbi_loadf_from_memory(unsigned char const* buffer, int len, int* x, int* y, int* comp, int req_comp);

// ---------------------------------------------------------------------------------------------------------
// rand(): glm::linearRand draws bytes with std::rand() % 255 (glm/gtc/random.inl:19-27).  glibc's rand() takes a
// lock and is unseeded in the reference.  This definition binds inside the library (-Bsymbolic) and
// gives every (pixel, primary sample) its own stream, so oracle renders are reproducible and threads do not
// serialise.  Distribution of rand() % 255 is unchanged (uniform 31-bit values).
// ---------------------------------------------------------------------------------------------------------
static thread_local uint64_t t_rng = 0x853c49e6748fea9bULL;
extern "C" int rand(void)
{
	t_rng ^= t_rng >> 12; t_rng ^= t_rng << 25; t_rng ^= t_rng >> 27;
	return (int)((t_rng * 0x2545F4914F6CDD1DULL) >> 33);
}
static void SeedStream(uint64_t seed, uint64_t index)
{
	uint64_t z = seed * 0x9E3779B97F4A7C15ULL + index + 0x632BE59BD9B4E019ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
	t_rng = z ? z : 0x9E3779B97F4A7C15ULL;
}

thread_local unsigned long long g_oracleBox = 0, g_oracleTri = 0, g_oracleRay = 0;

static thread_local std::string t_lastError;
static SailorPtStats g_stats{};

namespace
{
	struct RefBVH : public BVH
	{
		RefBVH(uint32_t n) : BVH(n) {}
		uint32_t NodesUsed() const { return m_nodesUsed; }
		void CopyOut(SailorPtBvhNode* nodes, uint32_t* mapping) const
		{
			if (nodes)
			{
				for (size_t i = 0; i < m_nodes.Num(); i++)
				{
					const auto& n = m_nodes[i];
					nodes[i].aabbMin[0] = n.m_aabbMin.x; nodes[i].aabbMin[1] = n.m_aabbMin.y; nodes[i].aabbMin[2] = n.m_aabbMin.z;
					nodes[i].aabbMax[0] = n.m_aabbMax.x; nodes[i].aabbMax[1] = n.m_aabbMax.y; nodes[i].aabbMax[2] = n.m_aabbMax.z;
					nodes[i].leftFirst = n.m_leftFirst; nodes[i].triCount = n.m_triCount;
				}
			}
			if (mapping)
			{
				for (size_t i = 0; i < m_triIdxMapping.Num(); i++) mapping[i] = m_triIdxMapping[i];
			}
		}
	};

	struct CameraDesc
	{
		std::string name;
		mat4 matrix{ 1 };   // reference convention: glm matrix whose memory is the row-major world matrix
		float aspect = 0.0f;
		float hFov = 0.0f;
	};

	struct RefTracer : public PathTracer
	{
		using PathTracer::m_triangles;
		using PathTracer::m_materials;
		using PathTracer::m_textures;
		using PathTracer::m_directionalLights;
		using PathTracer::m_textureMapping;
		using PathTracer::GetMaterialData;
		vec3 DoRaytrace(const Ray& r, const BVH& bvh, uint32_t bounces, const Params& p) const
		{
			return Raytrace(r, bvh, bounces, (uint32_t)(-1), p, 1.0f, 1.0f);
		}
	};
}

struct SailorPtScene
{
	RefTracer tracer;
	RefBVH* bvh = nullptr;
	std::vector<CameraDesc> cameras;
	std::vector<float> residentLin; std::vector<uint8_t> residentSrgb; uint32_t residentW = 0, residentH = 0;
	~SailorPtScene() { delete bvh; }
};

// ---------------------------------------------------------------------------------------------------------
// glTF front end (tinygltf instead of Assimp; loader contract in DESIGN.md)
// ---------------------------------------------------------------------------------------------------------
namespace
{
	// Node local matrix in the reference's convention (MaterialUtils.cpp:186-191: memcpy of Assimp's row-major
	// matrix into a glm mat4, i.e. the transpose of the mathematical matrix).
	mat4 NodeLocalMatrix(const tinygltf::Node& n)
	{
		float m[16]; // mathematical matrix, column-major as glTF stores it
		if (n.matrix.size() == 16)
		{
			for (int i = 0; i < 16; i++) m[i] = (float)n.matrix[i];
		}
		else
		{
			float t[3] = { 0, 0, 0 }, s[3] = { 1, 1, 1 }, q[4] = { 0, 0, 0, 1 };
			if (n.translation.size() == 3) for (int i = 0; i < 3; i++) t[i] = (float)n.translation[i];
			if (n.scale.size() == 3) for (int i = 0; i < 3; i++) s[i] = (float)n.scale[i];
			if (n.rotation.size() == 4) for (int i = 0; i < 4; i++) q[i] = (float)n.rotation[i];
			const float x = q[0], y = q[1], z = q[2], w = q[3];
			// R[r][c], loader contract: exactly these expressions
			const float R[3][3] = {
				{ 1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y - w * z), 2.0f * (x * z + w * y) },
				{ 2.0f * (x * y + w * z), 1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z - w * x) },
				{ 2.0f * (x * z - w * y), 2.0f * (y * z + w * x), 1.0f - 2.0f * (x * x + y * y) } };
			for (int c = 0; c < 3; c++)
			{
				for (int r = 0; r < 3; r++) m[c * 4 + r] = R[r][c] * s[c];
				m[c * 4 + 3] = 0.0f;
			}
			m[12] = t[0]; m[13] = t[1]; m[14] = t[2]; m[15] = 1.0f;
		}
		return glm::transpose(glm::make_mat4(m));
	}

	struct Accessor
	{
		const unsigned char* base = nullptr;
		size_t stride = 0;
		size_t count = 0;
		int componentType = 0;
		int comps = 0;
		bool normalized = false;
		bool valid = false;
	};

	Accessor MakeAccessor(const tinygltf::Model& model, int index)
	{
		Accessor a;
		if (index < 0 || index >= (int)model.accessors.size()) return a;
		const auto& acc = model.accessors[index];
		if (acc.bufferView < 0 || acc.bufferView >= (int)model.bufferViews.size()) return a;
		const auto& view = model.bufferViews[acc.bufferView];
		const auto& buf = model.buffers[view.buffer];
		a.comps = tinygltf::GetNumComponentsInType(acc.type);
		a.componentType = acc.componentType;
		const size_t elem = (size_t)tinygltf::GetComponentSizeInBytes(acc.componentType) * a.comps;
		a.stride = view.byteStride ? view.byteStride : elem;
		a.base = buf.data.data() + view.byteOffset + acc.byteOffset;
		a.count = acc.count;
		a.normalized = acc.normalized;
		a.valid = true;
		return a;
	}

	float ReadFloat(const Accessor& a, size_t i, int c)
	{
		const unsigned char* p = a.base + a.stride * i;
		switch (a.componentType)
		{
		case TINYGLTF_COMPONENT_TYPE_FLOAT: { float v; memcpy(&v, p + 4 * c, 4); return v; }
		case TINYGLTF_COMPONENT_TYPE_UNSIGNED_BYTE: { const float v = (float)p[c]; return a.normalized ? v / 255.0f : v; }
		case TINYGLTF_COMPONENT_TYPE_BYTE: { const float v = (float)((const signed char*)p)[c]; return a.normalized ? std::max(v / 127.0f, -1.0f) : v; }
		case TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT: { uint16_t u; memcpy(&u, p + 2 * c, 2); const float v = (float)u; return a.normalized ? v / 65535.0f : v; }
		case TINYGLTF_COMPONENT_TYPE_SHORT: { int16_t u; memcpy(&u, p + 2 * c, 2); const float v = (float)u; return a.normalized ? std::max(v / 32767.0f, -1.0f) : v; }
		default: return 0.0f;
		}
	}

	uint32_t ReadIndex(const Accessor& a, size_t i)
	{
		const unsigned char* p = a.base + a.stride * i;
		switch (a.componentType)
		{
		case TINYGLTF_COMPONENT_TYPE_UNSIGNED_BYTE: return p[0];
		case TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT: { uint16_t u; memcpy(&u, p, 2); return u; }
		case TINYGLTF_COMPONENT_TYPE_UNSIGNED_INT: { uint32_t u; memcpy(&u, p, 4); return u; }
		default: return 0;
		}
	}

	vec3 ReadVec3(const Accessor& a, size_t i) { return vec3(ReadFloat(a, i, 0), ReadFloat(a, i, 1), ReadFloat(a, i, 2)); }
	vec2 ReadVec2(const Accessor& a, size_t i) { return vec2(ReadFloat(a, i, 0), ReadFloat(a, i, 1)); }

	static thread_local bool t_needDefaultMaterial = false;     // set while flattening: some primitive uses the default material

	// ProcessMesh_Assimp (MaterialUtils.cpp:64-163) for one glTF primitive.
	void FlattenPrimitive(const tinygltf::Model& model, const tinygltf::Primitive& prim, TVector<Triangle>& out, const mat4& matrix)
	{
		if (prim.mode != TINYGLTF_MODE_TRIANGLES && prim.mode != -1) return;
		auto find = [&](const char* name) { auto it = prim.attributes.find(name); return it == prim.attributes.end() ? -1 : it->second; };
		const Accessor pos = MakeAccessor(model, find("POSITION"));
		if (!pos.valid) return;
		// an attribute stream whose element count differs from POSITION's is malformed: it is ignored (loader contract)
		auto attr = [&](const char* name) { Accessor a = MakeAccessor(model, find(name)); if (a.valid && a.count != pos.count) a.valid = false; return a; };
		const Accessor nrm = attr("NORMAL");
		const Accessor tan = attr("TANGENT");
		const Accessor uv0 = attr("TEXCOORD_0");
		const Accessor uv1 = attr("TEXCOORD_1");
		const Accessor idx = MakeAccessor(model, prim.indices);
		const size_t numIdx = idx.valid ? idx.count : pos.count;
		const size_t numFaces = numIdx / 3;

		const size_t start = out.Num();
		out.AddDefault(numFaces);
		for (size_t f = 0; f < numFaces; f++)
		{
			uint32_t vi[3];
			for (int k = 0; k < 3; k++) vi[k] = idx.valid ? ReadIndex(idx, f * 3 + k) : (uint32_t)(f * 3 + k);
			Triangle& tri = out[start + f];
			vec3 lp[3];
			for (int k = 0; k < 3; k++) lp[k] = ReadVec3(pos, vi[k]);

			// aiProcess_GenNormals (PathTracer.cpp:88-92): flat face normal in mesh space when NORMAL is absent
			vec3 ln[3];
			if (nrm.valid) { for (int k = 0; k < 3; k++) ln[k] = ReadVec3(nrm, vi[k]); }
			else { const vec3 fn = glm::normalize(glm::cross(lp[1] - lp[0], lp[2] - lp[0])); ln[0] = ln[1] = ln[2] = fn; }

			for (int k = 0; k < 3; k++)
			{
				tri.m_normals[k] = vec3(vec4(ln[k].x, ln[k].y, ln[k].z, 0.0f) * matrix);            // :90-92
				const vec4 temp = vec4(lp[k].x, lp[k].y, lp[k].z, 1.0f) * matrix;                  // :96-103
				tri.m_vertices[k] = vec3(temp.xyz) / temp.w;
			}
			tri.m_centroid = (tri.m_vertices[0] + tri.m_vertices[1] + tri.m_vertices[2]) * 0.333f; // :105

			// Assimp's glTF2 importer stores v' = 1-v and aiProcess_FlipUVs (PathTracer.cpp:87) flips it back: net identity
			if (uv0.valid) for (int k = 0; k < 3; k++) tri.m_uvs[k] = ReadVec2(uv0, vi[k]);         // :107-112
			if (uv1.valid) for (int k = 0; k < 3; k++) tri.m_uvs2[k] = ReadVec2(uv1, vi[k]);        // :114-119

			if (tan.valid && tan.comps == 4)                                                        // :121-140
			{
				for (int k = 0; k < 3; k++)
				{
					const vec3 t = ReadVec3(tan, vi[k]);
					const float w = ReadFloat(tan, vi[k], 3);
					const vec3 b = glm::cross(ln[k], t) * w; // Assimp glTF2 importer: bitangent = (normal x tangent) * w
					tri.m_tangent[k] = vec3(vec4(t.x, t.y, t.z, 0.0f) * matrix);
					tri.m_bitangent[k] = vec3(vec4(b.x, b.y, b.z, 0.0f) * matrix);
				}
			}
			else                                                                                    // :141-157
			{
				vec3 t = vec3(0.0f), b = vec3(0.0f);
				Raytracing::GenerateTangentBitangent(t, b, &tri.m_vertices[0], &tri.m_uvs[0]);
				for (int k = 0; k < 3; k++) { tri.m_tangent[k] = t; tri.m_bitangent[k] = b; }
			}
			// :159.  A primitive without a material (or with an index outside the array) gets the DEFAULT material, which Assimp's
			// glTF2 importer appends after the file's own materials (index = their count)
			const int nm = (int)model.materials.size();
			const bool own = prim.material >= 0 && prim.material < nm;
			if (!own) t_needDefaultMaterial = true;
			tri.m_materialIndex = (u8)(own ? prim.material : nm);
		}
	}

	struct Loader
	{
		const tinygltf::Model& model;
		SailorPtScene& scene;
		std::vector<mat4> lightMatrix;   // per KHR light index: world matrix of the first node that uses it
		std::vector<int> lightSeen;

		// ProcessNode_Assimp (MaterialUtils.cpp:186-198): world = node * parent in the reference's transposed convention
		void Node(int nodeIndex, const mat4& parentMatrix)
		{
			const auto& node = model.nodes[nodeIndex];
			const mat4 world = NodeLocalMatrix(node) * parentMatrix;
			if (node.mesh >= 0 && node.mesh < (int)model.meshes.size())
			{
				for (const auto& prim : model.meshes[node.mesh].primitives) FlattenPrimitive(model, prim, scene.tracer.m_triangles, world);
			}
			if (node.camera >= 0 && node.camera < (int)scene.cameras.size() && scene.cameras[node.camera].name == "\x01unset")
			{
				// DEVIATION (documented): GetWorldTransformMatrix (MaterialUtils.cpp:168-184) multiplies leaf->root, which
				// is the world matrix only when the ancestors commute; the hierarchical matrix is used instead.
				scene.cameras[node.camera].matrix = world;
				scene.cameras[node.camera].name = node.name.empty() ? model.cameras[node.camera].name : node.name;
			}
			auto ext = node.extensions.find("KHR_lights_punctual");
			if (ext != node.extensions.end() && ext->second.Has("light"))
			{
				const int li = ext->second.Get("light").GetNumberAsInt();
				if (li >= 0 && li < (int)lightMatrix.size() && !lightSeen[li]) { lightMatrix[li] = world; lightSeen[li] = 1; }
			}
			for (int child : node.children) Node(child, world);
		}
	};

	float ExtNumber(const tinygltf::Value& v, const char* key, float def)
	{
		return (v.IsObject() && v.Has(key) && v.Get(key).IsNumber()) ? (float)v.Get(key).GetNumberAsDouble() : def;
	}

	SamplerClamping ClampingOf(const tinygltf::Model& model, int textureIndex)
	{
		// Assimp maps REPEAT -> aiTextureMapMode_Wrap; everything else is not Wrap (PathTracer.cpp:226,243,...)
		const auto& tex = model.textures[textureIndex];
		if (tex.sampler < 0 || tex.sampler >= (int)model.samplers.size()) return SamplerClamping::Repeat;
		return model.samplers[tex.sampler].wrapS == TINYGLTF_TEXTURE_WRAP_REPEAT ? SamplerClamping::Repeat : SamplerClamping::Clamp;
	}

	// LoadTexture_Task (MaterialUtils.h:189-269): decode to RGBA8 with stb, then CombinedSampler2D::Initialize
	template<typename T>
	bool LoadTexture(SailorPtScene& scene, const std::vector<std::vector<unsigned char>>& imageBytes, uint32_t slot, int imageIndex,
		SamplerClamping clamping, bool bConvertToLinear, bool bNormalMap)
	{
		auto ptr = scene.tracer.m_textures[slot] = TSharedPtr<CombinedSampler2D>::Make();
		ptr->m_clamping = clamping;
		ptr->m_channels = sizeof(T) == sizeof(vec4) ? 4 : 3;
		if (imageIndex < 0 || imageIndex >= (int)imageBytes.size() || imageBytes[imageIndex].empty()) return false;
		int ch = 0;
		if (stbi_is_hdr_from_memory(imageBytes[imageIndex].data(), (int)imageBytes[imageIndex].size()))          // MaterialUtils.h:224-229, 250-253
		{
			float* pf = stbi_loadf_from_memory(imageBytes[imageIndex].data(), (int)imageBytes[imageIndex].size(), &ptr->m_width, &ptr->m_height, &ch, 4);
			if (!pf) return false;
			ptr->template Initialize<T, vec4>((vec4*)pf, bConvertToLinear, bNormalMap);
			stbi_image_free(pf);
			return true;
		}
		unsigned char* px = stbi_load_from_memory(imageBytes[imageIndex].data(), (int)imageBytes[imageIndex].size(), &ptr->m_width, &ptr->m_height, &ch, 4);
		if (!px) return false;
		ptr->template Initialize<T, u8vec4>((u8vec4*)px, bConvertToLinear, bNormalMap);
		stbi_image_free(px);
		return true;
	}

	int32_t LoadScene(const char* path, SailorPtScene& scene)
	{
		tinygltf::Model model;
		tinygltf::TinyGLTF ctx;
		std::vector<std::vector<unsigned char>> imageBytes;
		// keep the encoded bytes: the reference decodes with stbi_load(..., STBI_rgb_alpha) itself (MaterialUtils.h:226-249)
		ctx.SetImageLoader([&](tinygltf::Image*, const int idx, std::string*, std::string*, int, int, const unsigned char* bytes, int size, void*) -> bool
			{
				if ((int)imageBytes.size() <= idx) imageBytes.resize(idx + 1);
				imageBytes[idx].assign(bytes, bytes + size);
				return true;
			}, nullptr);
		std::string err, warn;
		const std::string p = path;
		const bool isGlb = p.size() >= 4 && (p.substr(p.size() - 4) == ".glb" || p.substr(p.size() - 4) == ".GLB");
		FILE* probe = fopen(path, "rb");
		if (!probe) { t_lastError = "cannot open " + p; return SAILOR_PT_ERR_IO; }
		fclose(probe);
		const bool ok = isGlb ? ctx.LoadBinaryFromFile(&model, &err, &warn, p) : ctx.LoadASCIIFromFile(&model, &err, &warn, p);
		if (!ok) { t_lastError = "tinygltf: " + err; return SAILOR_PT_ERR_FORMAT; }
		imageBytes.resize(model.images.size());

		if (model.materials.size() > 256) { t_lastError = "more than 256 materials"; return SAILOR_PT_ERR_LIMIT; }

		// cameras (PathTracer.cpp:111-148); Assimp: mHorizontalFOV = 2 atan(tan(yfov/2) * aspect)
		scene.cameras.resize(model.cameras.size());
		for (size_t i = 0; i < model.cameras.size(); i++)
		{
			auto& c = scene.cameras[i];
			c.name = "\x01unset";
			if (model.cameras[i].type == "perspective")
			{
				c.aspect = (float)model.cameras[i].perspective.aspectRatio;
				const float yfov = (float)model.cameras[i].perspective.yfov;
				c.hFov = 2.0f * std::atan(std::tan(yfov * 0.5f) * ((c.aspect == 0.0f) ? 1.0f : c.aspect));
			}
		}

		size_t numLights = 0;
		const tinygltf::Value* lightsArr = nullptr;
		{
			auto it = model.extensions.find("KHR_lights_punctual");
			if (it != model.extensions.end() && it->second.Has("lights") && it->second.Get("lights").IsArray())
			{
				lightsArr = &it->second.Get("lights");
				numLights = lightsArr->ArrayLen();
			}
		}

		Loader loader{ model, scene };
		loader.lightMatrix.assign(numLights, mat4(1));
		loader.lightSeen.assign(numLights, 0);
		const int sceneIndex = model.defaultScene >= 0 ? model.defaultScene : 0;
		t_needDefaultMaterial = false;
		if (sceneIndex < (int)model.scenes.size())
		{
			for (int root : model.scenes[sceneIndex].nodes) loader.Node(root, mat4(1.0f));   // PathTracer.cpp:161
		}
		for (auto& c : scene.cameras) if (c.name == "\x01unset") c.name = "";

		// materials + textures (PathTracer.cpp:164-360)
		auto& T = scene.tracer;
		// the glTF default material (baseColor 1, metallic 1, roughness 1, opaque) is appended when a primitive needs it
		std::vector<tinygltf::Material> mats = model.materials;
		if (t_needDefaultMaterial) mats.push_back(tinygltf::Material());
		if (mats.size() > 256) { t_lastError = "more than 256 materials (with the default material)"; return SAILOR_PT_ERR_LIMIT; }
		T.m_materials.Resize(mats.size());
		T.m_textures.Resize(mats.size() * 5);
		uint32_t textureIndex = 0;
		bool limit = false;
		for (size_t i = 0; i < mats.size(); i++)
		{
			auto& material = T.m_materials[i];
			const auto& gm = mats[i];

			material.m_blendMode = BlendMode::Opaque;                                        // :186-203
			if (gm.alphaMode == "BLEND") material.m_blendMode = BlendMode::Blend;
			else if (gm.alphaMode == "MASK") material.m_blendMode = BlendMode::Mask;
			material.m_alphaCutoff = (float)gm.alphaCutoff;                                  // :204-208
			{
				auto it = gm.extensions.find("KHR_materials_ior");                           // :209-213
				if (it != gm.extensions.end()) material.m_indexOfRefraction = ExtNumber(it->second, "ior", 1.5f);
			}

			auto bind = [&](int gltfTexture, uint8_t channels, u8& slotOut, bool linear, bool normalMap, bool checkKey)
				{
					if (gltfTexture < 0 || gltfTexture >= (int)model.textures.size()) return;
					const int image = model.textures[gltfTexture].source;
					const std::string file = "*" + std::to_string(image);                    // Assimp names embedded textures "*N"
					const SamplerClamping clamping = ClampingOf(model, gltfTexture);
					const bool known = T.m_textureMapping.ContainsKey(file);
					if (known && (!checkKey || (T.m_textures[T.m_textureMapping[file]]->m_clamping == clamping &&
						T.m_textures[T.m_textureMapping[file]]->m_channels == channels)))
					{
						slotOut = (u8)T.m_textureMapping[file];
						return;
					}
					if (textureIndex >= 255) { limit = true; return; }
					if (channels == 4) LoadTexture<vec4>(scene, imageBytes, textureIndex, image, clamping, linear, normalMap);
					else LoadTexture<vec3>(scene, imageBytes, textureIndex, image, clamping, linear, normalMap);
					T.m_textureMapping[file] = textureIndex;
					slotOut = (u8)textureIndex++;
				};

			bind(gm.pbrMetallicRoughness.baseColorTexture.index, 4, material.m_baseColorIndex, true, false, true);          // :217-235
			bind(gm.normalTexture.index, 3, material.m_normalIndex, false, true, true);                                     // :239-255
			bind(gm.pbrMetallicRoughness.metallicRoughnessTexture.index, 3, material.m_metallicRoughnessIndex, false, false, true); // :259-276
			bind(gm.emissiveTexture.index, 3, material.m_emissiveIndex, true, false, true);                                 // :280-296
			float transmission = 0.0f;
			{
				auto it = gm.extensions.find("KHR_materials_transmission");
				if (it != gm.extensions.end())
				{
					transmission = ExtNumber(it->second, "transmissionFactor", 0.0f);
					if (it->second.Has("transmissionTexture"))
					{
						const int ti = it->second.Get("transmissionTexture").Get("index").GetNumberAsInt();
						bind(ti, 3, material.m_transmissionIndex, false, false, false);                                       // :300-313
					}
				}
			}

			material.m_emissiveFactor = vec3((float)gm.emissiveFactor[0], (float)gm.emissiveFactor[1], (float)gm.emissiveFactor[2]); // :319-322
			{
				auto it = gm.extensions.find("KHR_materials_emissive_strength");
				material.m_emissiveFactor *= it != gm.extensions.end() ? ExtNumber(it->second, "emissiveStrength", 1.0f) : 1.0f;    // :324
			}
			material.m_transmissionFactor = transmission;                                                                         // :325
			const auto& bc = gm.pbrMetallicRoughness.baseColorFactor;
			material.m_baseColorFactor = vec4((float)bc[0], (float)bc[1], (float)bc[2], (float)bc[3]);                            // :326
			material.m_roughnessFactor = (float)gm.pbrMetallicRoughness.roughnessFactor;                                          // :327
			material.m_metallicFactor = (float)gm.pbrMetallicRoughness.metallicFactor;                                            // :328
			{
				auto it = gm.extensions.find("KHR_materials_volume");                                                             // :330-332
				if (it != gm.extensions.end())
				{
					material.m_thicknessFactor = ExtNumber(it->second, "thicknessFactor", 0.0f);
					material.m_attenuationDistance = ExtNumber(it->second, "attenuationDistance", std::numeric_limits<float>().max());
					if (it->second.Has("attenuationColor") && it->second.Get("attenuationColor").ArrayLen() == 3)
					{
						const auto& c = it->second.Get("attenuationColor");
						material.m_attenuationColor = vec3((float)c.Get(0).GetNumberAsDouble(), (float)c.Get(1).GetNumberAsDouble(), (float)c.Get(2).GetNumberAsDouble());
					}
				}
			}
			{
				auto it = gm.pbrMetallicRoughness.baseColorTexture.extensions.find("KHR_texture_transform");                      // :334-357
				if (it != gm.pbrMetallicRoughness.baseColorTexture.extensions.end())
				{
					float offX = 0, offY = 0, scX = 1, scY = 1, rot = ExtNumber(it->second, "rotation", 0.0f);
					if (it->second.Has("offset") && it->second.Get("offset").ArrayLen() == 2) { offX = (float)it->second.Get("offset").Get(0).GetNumberAsDouble(); offY = (float)it->second.Get("offset").Get(1).GetNumberAsDouble(); }
					if (it->second.Has("scale") && it->second.Get("scale").ArrayLen() == 2) { scX = (float)it->second.Get("scale").Get(0).GetNumberAsDouble(); scY = (float)it->second.Get("scale").Get(1).GetNumberAsDouble(); }
					const glm::mat3 scale = mat3(scX, 0, 0, 0, scY, 0, 0, 0, 1);
					const glm::mat3 translation = mat3(1, 0, 0, 0, 1, 0, offX, offY, 1);
					const glm::mat3 rotation = mat3(cos(rot), -sin(rot), 0, sin(rot), cos(rot), 0, 0, 0, 1);
					material.m_uvTransform = translation * rotation * scale;
				}
			}
		}
		if (limit) { t_lastError = "more than 255 textures"; return SAILOR_PT_ERR_LIMIT; }

		// directional lights (PathTracer.cpp:362-381); loader contract: intensity vector = color * intensity / 683
		for (size_t i = 0; i < numLights; i++)
		{
			const auto& L = lightsArr->Get((int)i);
			if (!L.Has("type") || L.Get("type").Get<std::string>() != "directional") continue;
			DirectionalLight dl;
			vec3 color(1.0f);
			if (L.Has("color") && L.Get("color").ArrayLen() == 3) color = vec3((float)L.Get("color").Get(0).GetNumberAsDouble(), (float)L.Get("color").Get(1).GetNumberAsDouble(), (float)L.Get("color").Get(2).GetNumberAsDouble());
			const float intensity = ExtNumber(L, "intensity", 1.0f);
			dl.m_direction = glm::normalize(glm::vec3(glm::vec4(0.0f, 0.0f, -1.0f, 0.0f) * loader.lightMatrix[i]));
			dl.m_intensity = color * intensity;
			dl.m_intensity /= 683.0f;
			T.m_directionalLights.Add(dl);
		}
		return SAILOR_PT_OK;
	}

	struct CameraSetup
	{
		uint32_t width = 0, height = 0;
		vec3 pos, pixel00Dir, deltaU, deltaV;
	};

	// PathTracer.cpp:102-153 + 390-403
	CameraSetup SetupCamera(const SailorPtScene& scene, const SailorPtParams& params)
	{
		auto cameraPos = vec3(0, 0.75f, 5.0f);
		auto cameraUp = normalize(vec3(0, 1, 0));
		auto cameraForward = normalize(-cameraPos);
		auto axis = normalize(cross(cameraForward, cameraUp));
		cameraUp = normalize(cross(axis, cameraForward));

		const bool hasCameras = !scene.cameras.empty();
		int32_t cameraIndex = 0;
		if (hasCameras)
		{
			const char* want = params.camera ? params.camera : "";
			for (uint32_t i = 0; i < scene.cameras.size(); i++)
			{
				if (std::strcmp(want, scene.cameras[i].name.c_str()) == 0) { cameraIndex = i; break; }
			}
			const mat4& matrix = scene.cameras[cameraIndex].matrix;
			// Assimp glTF camera: position 0, lookAt -Z, up +Y
			cameraUp = vec3(0, 1, 0); cameraForward = vec3(0, 0, -1);
			const vec4 translation = glm::vec4(0.0f, 0.0f, 0.0f, 1.0f) * matrix;
			cameraPos = vec3(translation.xyz) / translation.w;
			cameraUp = glm::normalize(glm::vec3(glm::vec4(cameraUp, 0.0f) * matrix));
			cameraForward = glm::normalize(glm::vec3(glm::vec4(cameraForward, 0.0f) * matrix));
		}

		float aspectRatio = (hasCameras && scene.cameras[cameraIndex].aspect > 0.0f) ? scene.cameras[cameraIndex].aspect : (4.0f / 3.0f);
		const uint32_t height = params.height;
		uint32_t width = static_cast<uint32_t>(height * aspectRatio);
		if (params.widthOverride)
		{
			// extension: explicit width; the aspect follows the image
			width = params.widthOverride;
			aspectRatio = (float)width / (float)height;
		}
		const float hFov = (hasCameras && scene.cameras[cameraIndex].hFov > 0.0f) ? scene.cameras[cameraIndex].hFov : glm::radians(60.0f);
		const float vFov = 2.0f * atan(tan(hFov * 0.5f) * (1.0f / aspectRatio));

		float h = tan(vFov / 2);
		const float ViewportHeight = 2.0f * h;
		const float ViewportWidth = aspectRatio * ViewportHeight;
		vec3 _u = normalize(cross(cameraUp, -cameraForward));
		vec3 _v = cross(-cameraForward, _u);
		const vec3 ViewportU = ViewportWidth * _u;
		const vec3 ViewportV = ViewportHeight * _v;
		const vec3 ViewportPivot = cameraPos - (ViewportU + ViewportV) * 0.5f + cameraForward;

		CameraSetup c;
		c.width = width; c.height = height; c.pos = cameraPos;
		c.deltaU = ViewportU / (float)width;
		c.deltaV = ViewportV / (float)height;
		c.pixel00Dir = ViewportPivot + 0.5f * (c.deltaU + c.deltaV) - cameraPos;
		return c;
	}

	PathTracer::Params ToRefParams(const SailorPtParams& p)
	{
		PathTracer::Params r;
		r.m_height = p.height; r.m_numSamples = p.numSamples; r.m_numAmbientSamples = p.numAmbientSamples;
		r.m_maxBounces = p.maxBounces; r.m_msaa = p.msaa;
		r.m_ambient = vec3(p.ambient[0], p.ambient[1], p.ambient[2]);
		return r;
	}

	uint32_t WorkerCount()
	{
		if (const char* e = getenv("SAILOR_PT_REF_THREADS")) { const int n = atoi(e); if (n > 0) return (uint32_t)n; }
		const uint32_t hc = std::thread::hardware_concurrency();
		return hc ? hc : 1;
	}

	template<typename F>
	void ParallelFor(uint32_t count, F&& body)
	{
		const uint32_t nThreads = std::max(1u, std::min(WorkerCount(), count));
		g_stats.threads = nThreads;
		std::atomic<uint32_t> next{ 0 };
		std::atomic<unsigned long long> box{ 0 }, tri{ 0 }, ray{ 0 };
		auto worker = [&]()
			{
				g_oracleBox = g_oracleTri = g_oracleRay = 0;
				for (;;) { const uint32_t i = next.fetch_add(1); if (i >= count) break; body(i); }
				box += g_oracleBox; tri += g_oracleTri; ray += g_oracleRay;
			};
		std::vector<std::thread> th;
		for (uint32_t t = 1; t < nThreads; t++) th.emplace_back(worker);
		worker();
		for (auto& t : th) t.join();
		g_stats.boxTests = box; g_stats.triTests = tri; g_stats.rays = ray;
	}

	double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

	void WriteHit(SailorPtHit& o, bool bHit, const RaycastHit& h)
	{
		if (bHit) { o.t = h.m_rayLenght; o.baryU = h.m_barycentricCoordinate.y; o.baryV = h.m_barycentricCoordinate.z; o.triId = h.m_triangleIndex; }
		else { o.t = std::numeric_limits<float>::infinity(); o.baryU = 0; o.baryV = 0; o.triId = 0xFFFFFFFFu; }
	}
}

// ---------------------------------------------------------------------------------------------------------
extern "C" {

const char* SailorPt_Backend(void) { return "reference-cpu"; }
int32_t SailorPt_OutputStageResident(SailorPtScene*, const void*, uint64_t) { return SAILOR_PT_ERR_UNSUPPORTED; }
int32_t SailorPt_SetDevice(int32_t device) { return device == 0 ? SAILOR_PT_OK : SAILOR_PT_ERR_ARG; }
int32_t SailorPt_WriteImage(const char*, uint32_t, uint32_t, const float*) { return SAILOR_PT_ERR_UNSUPPORTED; }
int32_t SailorPt_CompareImages(uint32_t, uint32_t, const float*, const float*, double*) { return SAILOR_PT_ERR_UNSUPPORTED; }
int32_t SailorPt_RenderProgressive(SailorPtScene*, const SailorPtParams*, uint32_t, uint32_t, const char*, uint32_t, float*, uint8_t*, uint32_t*) { return SAILOR_PT_ERR_UNSUPPORTED; }
int32_t SailorPt_TrimMemory(void) { return SAILOR_PT_OK; }
int32_t SailorPt_PinHostBuffer(void* hostBuffer, uint64_t bytes) { return (hostBuffer && bytes) ? SAILOR_PT_OK : SAILOR_PT_ERR_ARG; }   // host memory is the oracle's own memory
int32_t SailorPt_UnpinHostBuffer(void* hostBuffer) { return hostBuffer ? SAILOR_PT_OK : SAILOR_PT_ERR_ARG; }
const char* SailorPt_LastError(void) { return t_lastError.c_str(); }
int32_t SailorPt_GetStats(SailorPtStats* s) { if (!s) return SAILOR_PT_ERR_ARG; *s = g_stats; return SAILOR_PT_OK; }

int32_t SailorPt_ParseCommandLineArgs(SailorPtParams* out, const char** args, int32_t num)
{
	if (!out || (!args && num > 0)) return SAILOR_PT_ERR_ARG;
	static thread_local PathTracer::Params p;
	static thread_local std::string sIn, sOut;
	p = PathTracer::Params();
	p.m_height = out->height; p.m_numSamples = out->numSamples; p.m_numAmbientSamples = out->numAmbientSamples;
	p.m_maxBounces = out->maxBounces; p.m_msaa = out->msaa; p.m_ambient = vec3(out->ambient[0], out->ambient[1], out->ambient[2]);
	if (out->pathToModel) p.m_pathToModel = out->pathToModel;
	if (out->output) p.m_output = out->output;
	if (out->camera) p.m_camera = out->camera;
	PathTracer::ParseCommandLineArgs(p, args, num);       // the reference's own parser
	sIn = p.m_pathToModel.string(); sOut = p.m_output.string();
	out->pathToModel = sIn.c_str(); out->output = sOut.c_str(); out->camera = p.m_camera.c_str();
	out->height = p.m_height; out->numSamples = p.m_numSamples; out->numAmbientSamples = p.m_numAmbientSamples;
	out->maxBounces = p.m_maxBounces; out->msaa = p.m_msaa;
	out->ambient[0] = p.m_ambient.x; out->ambient[1] = p.m_ambient.y; out->ambient[2] = p.m_ambient.z;
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneLoad(const char* path, SailorPtScene** outScene)
{
	if (!path || !outScene) return SAILOR_PT_ERR_ARG;
	static_assert(sizeof(Math::Triangle) == 208 && sizeof(Math::Ray) == 48 && sizeof(Math::RaycastHit) == 60, "layout drifted from the reference");
	auto* s = new SailorPtScene();
	const int32_t rc = LoadScene(path, *s);
	if (rc != SAILOR_PT_OK) { delete s; *outScene = nullptr; return rc; }
	*outScene = s;
	return SAILOR_PT_OK;
}

void SailorPt_SceneFree(SailorPtScene* s) { delete s; }

int32_t SailorPt_SceneCounts(const SailorPtScene* s, uint32_t c[6])
{
	if (!s || !c) return SAILOR_PT_ERR_ARG;
	uint32_t nTex = 0;
	for (size_t i = 0; i < s->tracer.m_textures.Num(); i++) nTex += s->tracer.m_textures[i].IsValid() ? 1 : 0;
	c[0] = (uint32_t)s->tracer.m_triangles.Num(); c[1] = (uint32_t)s->tracer.m_materials.Num(); c[2] = nTex;
	c[3] = (uint32_t)s->tracer.m_directionalLights.Num(); c[4] = (uint32_t)s->cameras.size(); c[5] = s->bvh ? s->bvh->NodesUsed() : 0;
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetMaterials(const SailorPtScene* s, uint32_t* words)
{
	if (!s || !words) return SAILOR_PT_ERR_ARG;
	for (size_t i = 0; i < s->tracer.m_materials.Num(); i++)
	{
		const Material& m = s->tracer.m_materials[i];          // MaterialUtils.h:138-177
		uint32_t* w = words + i * SAILOR_PT_MATERIAL_WORDS;
		float f[26];
		for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) f[c * 3 + r] = m.m_uvTransform[c][r];
		for (int k = 0; k < 4; k++) f[9 + k] = m.m_baseColorFactor[k];
		for (int k = 0; k < 3; k++) { f[13 + k] = m.m_emissiveFactor[k]; f[16 + k] = m.m_attenuationColor[k]; }
		f[19] = m.m_metallicFactor; f[20] = m.m_roughnessFactor; f[21] = m.m_indexOfRefraction; f[22] = m.m_transmissionFactor;
		f[23] = m.m_alphaCutoff; f[24] = m.m_thicknessFactor; f[25] = m.m_attenuationDistance;
		memcpy(w, f, sizeof(f));
		w[26] = (uint32_t)m.m_blendMode; w[27] = m.m_baseColorIndex; w[28] = m.m_normalIndex; w[29] = m.m_metallicRoughnessIndex;
		w[30] = m.m_emissiveIndex; w[31] = m.m_transmissionIndex;
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetLights(const SailorPtScene* s, float* out)
{
	if (!s || !out) return SAILOR_PT_ERR_ARG;
	for (size_t i = 0; i < s->tracer.m_directionalLights.Num(); i++)
	{
		const DirectionalLight& l = s->tracer.m_directionalLights[i];   // LightingModel.h:10-14
		for (int k = 0; k < 3; k++) { out[i * 6 + k] = l.m_direction[k]; out[i * 6 + 3 + k] = l.m_intensity[k]; }
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_SceneGetTriangles(const SailorPtScene* s, float* tris, uint8_t* mat)
{
	if (!s) return SAILOR_PT_ERR_ARG;
	for (size_t i = 0; i < s->tracer.m_triangles.Num(); i++)
	{
		const Triangle& t = s->tracer.m_triangles[i];
		if (tris) memcpy(tris + i * SAILOR_PT_TRI_FLOATS, &t.m_centroid, sizeof(float) * SAILOR_PT_TRI_FLOATS);
		if (mat) mat[i] = t.m_materialIndex;
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_BuildBVH(SailorPtScene* s)
{
	if (!s) return SAILOR_PT_ERR_ARG;
	if (s->bvh) return SAILOR_PT_OK;
	if (s->tracer.m_triangles.Num() == 0) { t_lastError = "scene has no triangles"; return SAILOR_PT_ERR_FORMAT; }
	const double t0 = Now();
	s->bvh = new RefBVH((uint32_t)s->tracer.m_triangles.Num());     // PathTracer.cpp:384-385
	s->bvh->BuildBVH(s->tracer.m_triangles);
	g_stats.secondsBvhBuild = Now() - t0;
	return SAILOR_PT_OK;
}

int32_t SailorPt_GetBVH(const SailorPtScene* s, SailorPtBvhNode* nodes, uint32_t* mapping)
{
	if (!s || !s->bvh) return SAILOR_PT_ERR_ARG;
	s->bvh->CopyOut(nodes, mapping);
	return SAILOR_PT_OK;
}

int32_t SailorPt_GetCamera(const SailorPtScene* s, const SailorPtParams* p, uint32_t* w, uint32_t* h, float cam[12])
{
	if (!s || !p) return SAILOR_PT_ERR_ARG;
	const CameraSetup c = SetupCamera(*s, *p);
	if (w) *w = c.width;
	if (h) *h = c.height;
	if (cam)
	{
		memcpy(cam + 0, &c.pos, 12); memcpy(cam + 3, &c.pixel00Dir, 12); memcpy(cam + 6, &c.deltaU, 12); memcpy(cam + 9, &c.deltaV, 12);
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_IntersectRays(SailorPtScene* s, uint32_t count, const float* o, const float* d, const uint32_t* ignore, SailorPtHit* hits)
{
	if (!s || !o || !d || !hits) return SAILOR_PT_ERR_ARG;
	int32_t rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	const double t0 = Now();
	const uint32_t chunk = 4096, nChunks = (count + chunk - 1) / chunk;
	ParallelFor(nChunks, [&](uint32_t c)
		{
			const uint32_t end = std::min(count, (c + 1) * chunk);
			for (uint32_t i = c * chunk; i < end; i++)
			{
				const Ray ray(vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
				RaycastHit hit;
				const bool b = s->bvh->IntersectBVH(ray, hit, 0, std::numeric_limits<float>().max(), ignore ? ignore[i] : (uint32_t)(-1));
				WriteHit(hits[i], b, hit);
			}
		});
	g_stats.secondsTraverse = g_stats.secondsTotal = Now() - t0;
	if (!g_stats.rays) g_stats.rays = count;
	return SAILOR_PT_OK;
}

int32_t SailorPt_IntersectRaysEx(SailorPtScene* s, uint32_t count, const float* o, const float* d, const uint32_t* ignore, uint32_t, SailorPtHit* hits)
{
	// the reference has one traversal (closest hit, BVH.cpp:122-191): a hit-or-miss query is "did it return true"
	return SailorPt_IntersectRays(s, count, o, d, ignore, hits);
}

int32_t SailorPt_PrimaryHits(SailorPtScene* s, const SailorPtParams* p, SailorPtHit* hits)
{
	if (!s || !p || !hits) return SAILOR_PT_ERR_ARG;
	int32_t rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	const CameraSetup c = SetupCamera(*s, *p);
	const double t0 = Now();
	ParallelFor(c.height, [&](uint32_t y)
		{
			Ray ray;
			ray.SetOrigin(c.pos);
			for (uint32_t x = 0; x < c.width; x++)
			{
				const vec2 offset = vec2(0.5f, 0.5f);                                                        // PathTracer.cpp:460
				const vec3 pixelDir = c.pixel00Dir + ((float)(x) + offset.x) * c.deltaU + ((float)(y) - offset.y) * c.deltaV;
				ray.SetDirection(glm::normalize(pixelDir));
				RaycastHit hit;
				const bool b = s->bvh->IntersectBVH(ray, hit, 0, std::numeric_limits<float>().max(), (uint32_t)(-1));
				WriteHit(hits[(size_t)y * c.width + x], b, hit);
			}
		});
	g_stats.secondsTraverse = g_stats.secondsTotal = Now() - t0;
	if (!g_stats.rays) g_stats.rays = (uint64_t)c.width * c.height;
	return SAILOR_PT_OK;
}

int32_t SailorPt_OutputStage(uint32_t width, uint32_t height, const float* linearRGB, uint8_t* srgb8)
{
	if (!linearRGB || !srgb8 || !width || !height) return SAILOR_PT_ERR_ARG;
	CombinedSampler2D outputTex;                                   // PathTracer.cpp:387-388
	outputTex.Initialize<vec3>(width, height);
	memcpy(outputTex.m_data.GetData(), linearRGB, (size_t)width * height * sizeof(vec3));
	const float aberrationAmount = (0.5f / width);                 // :539
	for (uint32_t y = 0; y < height; y++)
	{
		for (uint32_t x = 0; x < width; x++)
		{
			vec2 uv = vec2((float)x / width, (float)y / height);
			vec3 greenColor = outputTex.Sample<vec3>(uv + vec2(aberrationAmount, 0));
			vec3 blueColor = outputTex.Sample<vec3>(uv + vec2(aberrationAmount, aberrationAmount));
			vec3 redColor = outputTex.Sample<vec3>(uv + vec2(-aberrationAmount, -aberrationAmount));
			vec3 chromaAberratedColor = vec3(redColor.r, greenColor.g, blueColor.b);
			const u8vec3 px = glm::clamp(Utils::LinearToSRGB(chromaAberratedColor) * 255.0f, 0.0f, 255.0f);   // :557
			uint8_t* o = srgb8 + ((size_t)x + (size_t)y * width) * 3;
			o[0] = px.r; o[1] = px.g; o[2] = px.b;
		}
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_Render(SailorPtScene* s, const SailorPtParams* p, float* linearRGB, uint8_t* srgb8)
{
	if (!s || !p || !linearRGB) return SAILOR_PT_ERR_ARG;
	int32_t rc = SailorPt_BuildBVH(s);
	if (rc != SAILOR_PT_OK) return rc;
	const CameraSetup c = SetupCamera(*s, *p);
	const PathTracer::Params params = ToRefParams(*p);
	const uint32_t width = c.width, height = c.height;
	const uint32_t GroupSize = 32;                                 // PathTracer.cpp:82
	const uint32_t rowBegin = p->rowEnd ? p->rowBegin : 0, rowEnd = p->rowEnd ? std::min(p->rowEnd, height) : height;
	const uint32_t msBegin = p->msaaEnd ? p->msaaBegin : 0, msEnd = p->msaaEnd ? std::min(p->msaaEnd, params.m_msaa) : params.m_msaa;
	const uint32_t tilesX = (width + GroupSize - 1) / GroupSize, tilesY = (height + GroupSize - 1) / GroupSize;
	memset(linearRGB, 0, (size_t)width * height * 3 * sizeof(float));
	const double t0 = Now();
	std::atomic<uint64_t> samples{ 0 };
	ParallelFor(tilesX * tilesY, [&](uint32_t tile)
		{
			const uint32_t x = (tile % tilesX) * GroupSize, y = (tile / tilesX) * GroupSize;
			Ray ray;
			ray.SetOrigin(c.pos);
			uint64_t local = 0;
			for (uint32_t v = 0; (v < GroupSize) && (y + v) < height; v++)             // :444-471
			{
				if ((y + v) < rowBegin || (y + v) >= rowEnd) continue;
				for (uint32_t u = 0; u < GroupSize && (u + x) < width; u++)
				{
					const uint32_t index = (height - (y + v) - 1) * width + (x + u);
					vec3 accumulator = vec3(0);
					for (uint32_t sample = msBegin; sample < msEnd; sample++)
					{
						SeedStream(p->seed, ((uint64_t)(y + v) * width + (x + u)) * params.m_msaa + sample);
						const vec2 offset = sample == 0 ? vec2(0.5f, 0.5f) : glm::linearRand(vec2(0, 0), vec2(1.0f, 1.0f));
						const vec3 pixelDir = c.pixel00Dir + ((float)(u + x) + offset.x) * c.deltaU + ((float)(y + v) - offset.y) * c.deltaV;
						ray.SetDirection(glm::normalize(pixelDir));
						accumulator += s->tracer.DoRaytrace(ray, *s->bvh, params.m_maxBounces, params);
						local++;
					}
					vec3 res = accumulator / (float)params.m_msaa;
					memcpy(linearRGB + (size_t)index * 3, &res, sizeof(vec3));          // outputTex.SetPixel(x+u, height-(y+v)-1, res)
				}
			}
			samples += local;
		});
	g_stats.secondsTotal = Now() - t0;
	g_stats.secondsShade = g_stats.secondsTotal;
	g_stats.primarySamples = samples;
	if (srgb8)
	{
		const double t1 = Now();
		SailorPt_OutputStage(width, height, linearRGB, srgb8);
		g_stats.secondsOutput = Now() - t1;
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_RenderResident(SailorPtScene* s, const SailorPtParams* p, uint32_t flags)
{
	if (!s || !p) return SAILOR_PT_ERR_ARG;
	if ((flags & 1u) && s->bvh) { delete s->bvh; s->bvh = nullptr; }
	const CameraSetup c = SetupCamera(*s, *p);
	s->residentW = c.width; s->residentH = c.height;
	s->residentLin.resize((size_t)c.width * c.height * 3);
	if (flags & 2u) s->residentSrgb.resize((size_t)c.width * c.height * 3);
	return SailorPt_Render(s, p, s->residentLin.data(), (flags & 2u) ? s->residentSrgb.data() : nullptr);
}

int32_t SailorPt_ReadResident(SailorPtScene* s, float* linearRGB, uint8_t* srgb8)
{
	if (!s || !s->residentW) return SAILOR_PT_ERR_ARG;
	if (linearRGB) memcpy(linearRGB, s->residentLin.data(), s->residentLin.size() * sizeof(float));
	if (srgb8) { if (s->residentSrgb.empty()) return SAILOR_PT_ERR_ARG; memcpy(srgb8, s->residentSrgb.data(), s->residentSrgb.size()); }
	return SAILOR_PT_OK;
}

int32_t SailorPt_CopyResidentToDevice(SailorPtScene*, void*, uint64_t) { t_lastError = "the CPU oracle has no device memory"; return SAILOR_PT_ERR_UNSUPPORTED; }

int32_t SailorPt_Run(const SailorPtParams* p)
{
	if (!p || !p->pathToModel) return SAILOR_PT_ERR_ARG;
	SailorPtScene* s = nullptr;
	int32_t rc = SailorPt_SceneLoad(p->pathToModel, &s);
	if (rc != SAILOR_PT_OK) return rc;
	uint32_t w = 0, h = 0;
	SailorPt_GetCamera(s, p, &w, &h, nullptr);
	std::vector<float> lin((size_t)w * h * 3);
	std::vector<uint8_t> srgb((size_t)w * h * 3);
	rc = SailorPt_Render(s, p, lin.data(), srgb.data());
	if (rc == SAILOR_PT_OK && p->output && p->output[0])
	{
		if (!stbi_write_png(p->output, w, h, 3, srgb.data(), w * 3)) { t_lastError = "Raytracing WriteImage error"; rc = SAILOR_PT_ERR_IO; }   // :560-564
	}
	SailorPt_SceneFree(s);
	return rc;
}

int32_t SailorPt_SampleTexture(SailorPtScene* s, uint32_t textureIndex, uint32_t count, const float* uv, float* out)
{
	if (!s || !uv || !out || textureIndex >= s->tracer.m_textures.Num() || !s->tracer.m_textures[textureIndex].IsValid()) return SAILOR_PT_ERR_ARG;
	const auto& tex = s->tracer.m_textures[textureIndex];
	for (uint32_t i = 0; i < count; i++)
	{
		const vec2 c(uv[2 * i], uv[2 * i + 1]);
		if (tex->m_channels == 4) { const vec4 v = tex->Sample<vec4>(c); memcpy(out + 4 * i, &v, 16); }
		else { const vec3 v = tex->Sample<vec3>(c); memcpy(out + 4 * i, &v, 12); out[4 * i + 3] = 0.0f; }
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_DecodeImage(const uint8_t* data, uint64_t size, uint32_t* width, uint32_t* height, uint8_t* rgba8, uint64_t capacity)
{
	// the reference's own decoder call (MaterialUtils.h:226-249): stbi_load_from_memory(..., STBI_rgb_alpha)
	if (!data || !size || !width || !height) return SAILOR_PT_ERR_ARG;
	int w = 0, h = 0, ch = 0;
	unsigned char* px = stbi_load_from_memory(data, (int)size, &w, &h, &ch, 4);
	if (!px) { t_lastError = "stbi_load_from_memory failed"; return SAILOR_PT_ERR_FORMAT; }
	*width = (uint32_t)w; *height = (uint32_t)h;
	int32_t rc = SAILOR_PT_OK;
	if (rgba8) { if (capacity < (uint64_t)w * h * 4) rc = SAILOR_PT_ERR_ARG; else memcpy(rgba8, px, (size_t)w * h * 4); }
	stbi_image_free(px);
	return rc;
}

int32_t SailorPt_ShadeHits(SailorPtScene* s, uint32_t count, const uint32_t* triIds, const float* baryUV, const float* rayDirs, uint32_t numSamples, uint32_t numAmbient, float* out)
{
	// The head of PathTracer::Raytrace after a hit (PathTracer.cpp:636-661), statement for statement, around the reference's OWN
	// GetMaterialData (:881-927, called here through the compiled reference source)
	if (!s || !triIds || !baryUV || !rayDirs || !out) return SAILOR_PT_ERR_ARG;
	const RefTracer& T = s->tracer;
	for (uint32_t i = 0; i < count; i++)
	{
		float* o = out + (size_t)i * SAILOR_PT_SHADE_FLOATS;
		if (triIds[i] >= T.m_triangles.Num()) { for (int k = 0; k < SAILOR_PT_SHADE_FLOATS; k++) o[k] = 0.0f; continue; }
		const Math::Triangle& tri = T.m_triangles[triIds[i]];
		const float bu = baryUV[2 * i], bv = baryUV[2 * i + 1];
		const vec3 bc(1.0f - bu - bv, bu, bv);                                      // RaycastHit::m_barycentricCoordinate (Bounds.cpp:521-523)
		const vec3 dir(rayDirs[3 * i], rayDirs[3 * i + 1], rayDirs[3 * i + 2]);
		vec3 faceNormal = vec3(bc.x * tri.m_normals[0] + bc.y * tri.m_normals[1] + bc.z * tri.m_normals[2]);
		const vec3 tangent = vec3(bc.x * tri.m_tangent[0] + bc.y * tri.m_tangent[1] + bc.z * tri.m_tangent[2]);
		const vec3 bitangent = vec3(bc.x * tri.m_bitangent[0] + bc.y * tri.m_bitangent[1] + bc.z * tri.m_bitangent[2]);
		const bool bIsOppositeRay = dot(faceNormal, dir) < 0.0f;
		if (!bIsOppositeRay) faceNormal *= -1.0f;
		const mat3 tbn(tangent, bitangent, faceNormal);
		const vec2 uv = bc.x * tri.m_uvs[0] + bc.y * tri.m_uvs[1] + bc.z * tri.m_uvs[2];
		const auto material = T.m_materials[tri.m_materialIndex];
		const vec2 uvTransformed = (material.m_uvTransform * vec3(uv, 1));
		const LightingModel::SampledData sample = T.GetMaterialData(tri.m_materialIndex, uvTransformed);
		const vec3 worldNormal = normalize(tbn * sample.m_normal);
		const bool bHasAlphaBlending = !sample.m_bIsOpaque && sample.m_baseColor.a < 1.0f;
		const uint32_t nS = bHasAlphaBlending ? std::max(1u, (uint32_t)round(sample.m_baseColor.a * (float)numSamples)) : numSamples;
		const uint32_t nA = bHasAlphaBlending ? std::max(1u, (uint32_t)round(sample.m_baseColor.a * (float)numAmbient)) : numAmbient;
		o[0] = sample.m_baseColor.x; o[1] = sample.m_baseColor.y; o[2] = sample.m_baseColor.z; o[3] = sample.m_baseColor.w;
		o[4] = sample.m_orm.x; o[5] = sample.m_orm.y; o[6] = sample.m_orm.z; o[7] = sample.m_emissive.x; o[8] = sample.m_emissive.y; o[9] = sample.m_emissive.z;
		o[10] = sample.m_normal.x; o[11] = sample.m_normal.y; o[12] = sample.m_normal.z; o[13] = sample.m_transmission; o[14] = sample.m_ior; o[15] = sample.m_thicknessFactor;
		o[16] = sample.m_bIsOpaque ? 1.0f : 0.0f;
		o[17] = worldNormal.x; o[18] = worldNormal.y; o[19] = worldNormal.z; o[20] = faceNormal.x; o[21] = faceNormal.y; o[22] = faceNormal.z;
		o[23] = uvTransformed.x; o[24] = uvTransformed.y; o[25] = bIsOppositeRay ? 1.0f : 0.0f; o[26] = (float)nS; o[27] = (float)nA;
	}
	return SAILOR_PT_OK;
}

int32_t SailorPt_SampleGenerators(uint64_t, uint32_t, uint32_t, float*) { return SAILOR_PT_ERR_UNSUPPORTED; }      // product only: the reference draws from unseeded rand()

int32_t SailorPt_EvalLighting(uint32_t count, const float* in, float* out)
{
	if (!in || !out) return SAILOR_PT_ERR_ARG;
	for (uint32_t i = 0; i < count; i++)
	{
		const float* r = in + 24 * i;
		float* o = out + 28 * i;
		LightingModel::SampledData sd;
		sd.m_baseColor = vec4(r[0], r[1], r[2], r[3]); sd.m_orm = vec3(r[4], r[5], r[6]); sd.m_emissive = vec3(r[7], r[8], r[9]);
		const vec3 N(r[10], r[11], r[12]), V(r[13], r[14], r[15]), L(r[16], r[17], r[18]);
		sd.m_ior = r[19]; sd.m_thicknessFactor = r[20]; sd.m_transmission = r[21];
		const vec2 Xi(r[22], r[23]);
		const float rough = sd.m_orm.y;
		const vec3 H = normalize(V + L);
		const vec3 brdf = LightingModel::CalculateBRDF(V, N, L, sd), btdf = LightingModel::CalculateBTDF(V, N, L, sd);
		memcpy(o + 0, &brdf, 12); memcpy(o + 3, &btdf, 12);
		o[6] = LightingModel::DistributionGGX(N, H, rough);
		o[7] = LightingModel::GeometrySchlickGGX(dot(N, L), rough);
		o[8] = LightingModel::GGX_PDF(N, H, V, rough);
		o[9] = LightingModel::Beckmann_PDF(N, H, V, rough);
		const vec3 s0 = LightingModel::ImportanceSampleGGX(Xi, rough, N), s1 = LightingModel::ImportanceSampleBeckmann(Xi, rough, N),
			s2 = LightingModel::ImportanceSampleLambert(Xi, N), s3 = LightingModel::ImportanceSampleHemisphere(Xi, N);
		memcpy(o + 10, &s0, 12); memcpy(o + 13, &s1, 12); memcpy(o + 16, &s2, 12); memcpy(o + 19, &s3, 12);
		o[22] = LightingModel::PowerHeuristic(3, o[8], 2, o[9]);
		const vec3 rf = LightingModel::CalculateRefraction(-V, N, 1.0f, sd.m_ior);
		memcpy(o + 23, &rf, 12);
		o[26] = LightingModel::FresnelSchlick(std::max(dot(H, V), 0.0f), vec3(0.04f)).x;
		o[27] = 0.0f;
	}
	return SAILOR_PT_OK;
}

} // extern "C"
