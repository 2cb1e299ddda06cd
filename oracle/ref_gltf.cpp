// ORACLE (test infrastructure) — the one translation unit that instantiates the reference's vendored tinygltf
// (External/tinygltf/tiny_gltf.h v2.9) and stb_image.  stb_image_write is instantiated by the reference's own
// PathTracer.cpp:14-18, so it is disabled here.
#define TINYGLTF_IMPLEMENTATION
#define TINYGLTF_NO_STB_IMAGE_WRITE
#define STB_IMAGE_IMPLEMENTATION
#include "tinygltf/tiny_gltf.h"
