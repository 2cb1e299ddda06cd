// ORACLE (test infrastructure) — the one translation unit that instantiates the reference's vendored tinygltf
// (External/tinygltf/tiny_gltf.h v2.9) and stb_image.  stb_image_write is instantiated by the reference's own
// PathTracer.cpp:14-18, so it is disabled here.  The decoder instantiated is the one the reference's path tracer includes
// (External/stb/stb_image.h v2.27, MaterialUtils.h:9), NOT tinygltf's own newer copy (v2.28): the two agree on valid files and differ
// on corrupt ones.
#define TINYGLTF_IMPLEMENTATION
#define TINYGLTF_NO_STB_IMAGE_WRITE
#define TINYGLTF_NO_INCLUDE_STB_IMAGE
#define STB_IMAGE_IMPLEMENTATION
#include "stb/stb_image.h"
#include "tinygltf/tiny_gltf.h"
