// ORACLE (test infrastructure) — restatement of the three Runtime/Core/Utils.cpp helpers the reference path
// tracer calls; the real Utils.cpp cannot be compiled on Linux (it includes Sailor.h -> Win32 window code).
//   LinearToSRGB  : reference Core/Utils.cpp:48-57   (vec3), :38-41 (vec4 keeps alpha)
//   SRGBToLinear  : reference Core/Utils.cpp:59-64   (vec3), :43-46 (vec4 keeps alpha)
//   GetArgValue   : reference Core/Utils.cpp:466-487
// Written against the reference's vendored glm so per-component arithmetic (std::pow, glm::mix) is the reference's.
#include "Core/Utils.h"

namespace Sailor { namespace Utils {

glm::vec3 LinearToSRGB(const glm::vec3& c)
{
	// mix(higher, lower, bvec): component-wise select, no arithmetic blend (glm/detail/func_common.inl, bool mix)
	const glm::vec3 hi = glm::vec3(1.055f) * glm::pow(c, glm::vec3(1.f / 2.4f)) - glm::vec3(0.055f);
	const glm::vec3 lo = c * glm::vec3(12.92f);
	return glm::mix(hi, lo, glm::lessThan(c, glm::vec3(0.0031308f)));
}

glm::vec3 SRGBToLinear(const glm::vec3& s)
{
	// here the selector is a FLOAT step(), so mix() is the arithmetic x*(1-a)+y*a — kept as the reference has it
	const glm::vec3 a = glm::step(glm::vec3(0.04045f), s);
	return glm::mix(s / glm::vec3(12.92f), glm::pow((s + glm::vec3(0.055f)) / glm::vec3(1.055f), glm::vec3(2.4f)), a);
}

glm::vec4 LinearToSRGB(const glm::vec4& c) { return glm::vec4(LinearToSRGB(glm::vec3(c)), c.a); }
glm::vec4 SRGBToLinear(const glm::vec4& s) { return glm::vec4(SRGBToLinear(glm::vec3(s)), s.a); }
glm::vec4 LinearToSRGB(const glm::u8vec4& c) { return LinearToSRGB(glm::vec4(c)); }
glm::vec4 SRGBToLinear(const glm::u8vec4& s) { return SRGBToLinear(glm::vec4(s)); }

std::string GetArgValue(const char** args, int32_t& i, int32_t num)
{
	if (i + 1 >= num) return "";
	std::string v = args[++i];
	if (!v.empty() && v[0] == '\"')
	{
		while (i < num && v[v.length() - 1] != '\"')
		{
			++i;
			v += " " + std::string(args[i]);
		}
		v = v.substr(1, v.length() - 2);
	}
	return v;
}

}}
