// Shim for the oracle build only. Declares the three helpers the path tracer uses from
// Runtime/Core/Utils.h (the real header drags in Sailor.h -> Platform/Win32/Window.h).
#pragma once
#include "Core/Defines.h" // same GLM_FORCE_* configuration in every TU (vec3 ABI depends on GLM_FORCE_SWIZZLE)
#include <string>
#include <cstdint>
#include <glm/glm/glm.hpp>
namespace Sailor { namespace Utils {
	glm::vec3 LinearToSRGB(const glm::vec3& linearRGB);
	glm::vec3 SRGBToLinear(const glm::vec3& srgbIn);
	glm::vec4 LinearToSRGB(const glm::vec4& linearRGB);
	glm::vec4 SRGBToLinear(const glm::vec4& srgbIn);
	glm::vec4 LinearToSRGB(const glm::u8vec4& linearRGB);
	glm::vec4 SRGBToLinear(const glm::u8vec4& srgbIn);
	std::string GetArgValue(const char** args, int32_t& i, int32_t num);
}}
