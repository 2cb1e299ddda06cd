#!/usr/bin/env python3
"""Build the CPU oracle `oracle/_ref/libsailor_pt_ref.so` from the UNMODIFIED live path-tracer sources of the
reference checkout (BVH.cpp, Bounds.cpp, LightingModel.cpp, MaterialUtils.cpp, PathTracer.cpp) plus this repo's restated driver.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product library
(sailor_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

Why a recipe and not the reference's CMake: the reference targets MSVC/Windows.  g++ rejects three
anonymous-struct unions that hold glm::vec3 (Math/Bounds.h:52-54,99-103; Raytracing/BVH.h:15-24), and quoted
includes resolve next to the including file first, so the ten sources are staged into a scratch directory under
the git-ignored oracle/_ref/, the three unions are rewritten there (layout unchanged: sizeof(Ray)=48,
sizeof(BVHNode)=32, checked by the driver at load), the objects are built, and the staged copies are deleted.
Reference sources never enter the repository history.

Usage: python oracle/build_ref.py [--reference /root/reference] [--keep-stage]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

STAGED = [
    "Runtime/Raytracing/BVH.h", "Runtime/Raytracing/BVH.cpp",
    "Runtime/Raytracing/LightingModel.h", "Runtime/Raytracing/LightingModel.cpp",
    "Runtime/Raytracing/MaterialUtils.h", "Runtime/Raytracing/MaterialUtils.cpp",
    "Runtime/Raytracing/PathTracer.h", "Runtime/Raytracing/PathTracer.cpp",
    "Runtime/Math/Bounds.h", "Runtime/Math/Bounds.cpp",
]

FLAGS = [
    "-std=c++20", "-O2", "-mavx2", "-ffp-contract=off", "-fPIC", "-w", "-fpermissive", "-DNDEBUG",
    "-D_MSC_EXTENSIONS", "-D__forceinline=", "-D__declspec(x)=",
    "-DSAILOR_PROFILE_ALLOC(a,b)=", "-DSAILOR_PROFILE_FREE(a)=", "-Dsprintf_s=snprintf",
    "-include", "immintrin.h", "-include", "cfloat", "-include", "cstring", "-include", "cmath",
]


def patch_bounds_h(src: str) -> str:
    """Math/Bounds.h: Ray's three `union { struct { vec3 m_x; float dummyN; }; __m128 X4; };` -> plain members."""
    n_total = 0
    for member, dummy, simd in (("m_origin", "dummy1", "O4"), ("m_direction", "dummy2", "D4"),
                                ("m_rDirection", "dummy3", "rD4")):
        pat = re.compile(r"union\s*\{\s*struct\s*\{\s*vec3\s+%s;\s*float\s+%s;\s*\};\s*__m128\s+%s;\s*\};"
                         % (member, dummy, simd))
        src, n = pat.subn("vec3 %s; float %s;" % (member, dummy), src)
        n_total += n
    assert n_total == 3, "Bounds.h Ray unions not found (%d)" % n_total
    src, n = re.subn(r"Ray\(\)\s*\{\s*O4\s*=\s*D4\s*=\s*rD4\s*=\s*_mm_set1_ps\(1\);\s*\}",
                     "Ray() { m_origin = m_direction = m_rDirection = vec3(1); dummy1 = dummy2 = dummy3 = 1; }", src)
    assert n == 1
    for getter, member in (("GetOrigin4", "m_origin"), ("GetDirection4", "m_direction"),
                           ("GetReciprocalDirection4", "m_rDirection")):
        src, n = re.subn(r"const\s+__m128&\s+%s\(\)\s*const\s*\{\s*return\s+\w+;\s*\}" % getter,
                         "__m128 %s() const { return _mm_loadu_ps(&%s.x); }" % (getter, member), src)
        assert n == 1, getter
    # Sphere
    src, n = re.subn(r"union\s*\{\s*struct\s*\{\s*glm::vec3\s+m_center;\s*float\s+m_radius;\s*\};\s*glm::vec4\s+m_vec4;\s*\};",
                     "glm::vec3 m_center; float m_radius;", src)
    assert n == 1
    src, n = re.subn(r"const\s+vec4&\s+GetVec4\(\)\s*const\s*\{\s*return\s+m_vec4;\s*\}",
                     "vec4 GetVec4() const { return vec4(m_center, m_radius); }", src)
    assert n == 1
    return src


def patch_bvh_h(src: str) -> str:
    """Raytracing/BVH.h: drop the unused __m128 arms of BVHNode."""
    for member, tail, simd in (("m_aabbMin", "m_leftFirst", "m_aabbMin4"), ("m_aabbMax", "m_triCount", "m_aabbMax4")):
        pat = re.compile(r"union\s*\{\s*struct\s*\{\s*vec3\s+%s;\s*uint\s+%s;\s*\};\s*__m128\s+%s;\s*\};"
                         % (member, tail, simd))
        src, n = pat.subn("vec3 %s; uint %s;" % (member, tail), src)
        assert n == 1, member
    return src


def patch_counters(path: str, text: str) -> str:
    """count build: instrument the staged copies with the n_box / n_tri / n_ray counters SURVEY.md §8(d) defines."""
    if path.endswith("Math/Bounds.cpp"):
        text = text.replace("float Math::IntersectRayAABB(const Ray& ray, const glm::vec3& bmin, const glm::vec3& bmax, float maxRayLength)\n{",
                            "float Math::IntersectRayAABB(const Ray& ray, const glm::vec3& bmin, const glm::vec3& bmax, float maxRayLength)\n{ g_oracleBox++;")
        text = "extern thread_local unsigned long long g_oracleBox, g_oracleTri, g_oracleRay;\n" + text
        text = text.replace("bool Math::IntersectRayTriangle(const Ray& ray, const Triangle& tri, RaycastHit& outRaycastHit, float maxRayLength)\n{",
                            "bool Math::IntersectRayTriangle(const Ray& ray, const Triangle& tri, RaycastHit& outRaycastHit, float maxRayLength)\n{ g_oracleTri++;")
        assert "g_oracleBox++" in text and "g_oracleTri++" in text
    if path.endswith("Raytracing/BVH.cpp"):
        text = text.replace("bool BVH::IntersectBVH(const Math::Ray& ray, Math::RaycastHit& outResult, const uint nodeIdx, float maxRayLength, uint32_t ignoreTriangle) const\n{",
                            "extern thread_local unsigned long long g_oracleBox, g_oracleTri, g_oracleRay;\n"
                            "bool BVH::IntersectBVH(const Math::Ray& ray, Math::RaycastHit& outResult, const uint nodeIdx, float maxRayLength, uint32_t ignoreTriangle) const\n{ g_oracleRay++;")
        assert "g_oracleRay++" in text
    return text


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("oracle build failed")


def build(ref: str, count: bool, keep_stage: bool = False) -> str:
    """count=False -> libsailor_pt_ref.so (timed as the CPU baseline); count=True -> libsailor_pt_ref_count.so
    (same sources + the n_box/n_tri/n_ray counters; used to derive algorithmic bytes per ray)."""
    if not os.path.isdir(os.path.join(ref, "Runtime", "Raytracing")):
        raise SystemExit("reference checkout not found at %s (the GPU box uses the prebuilt oracle/_ref/*.so)" % ref)

    stage = os.path.join(OUT, "stage_count" if count else "stage")
    shutil.rmtree(stage, ignore_errors=True)
    os.makedirs(stage)
    for rel in STAGED:
        sub = rel[len("Runtime/"):]
        dst = os.path.join(stage, sub)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        text = open(os.path.join(ref, rel), encoding="utf-8-sig").read().replace("\r\n", "\n")
        if rel.endswith("Math/Bounds.h"):
            text = patch_bounds_h(text)
        elif rel.endswith("Raytracing/BVH.h"):
            text = patch_bvh_h(text)
        if count:
            text = patch_counters(rel, text)
        open(dst, "w", encoding="utf-8").write(text)

    inc = ["-I", os.path.join(HERE, "shim"), "-I", stage, "-I", os.path.join(ref, "Runtime"),
           "-I", os.path.join(ref, "External"), "-I", os.path.join(ref, "External", "tinygltf"),
           "-I", os.path.join(ref, "External", "nlohmann_json", "include"),
           "-I", os.path.join(HERE, "..", "include")]
    units = [
        (os.path.join(stage, "Raytracing", "PathTracer.cpp"), "PathTracer.o", FLAGS),
        (os.path.join(stage, "Raytracing", "BVH.cpp"), "BVH.o", FLAGS),
        (os.path.join(stage, "Raytracing", "LightingModel.cpp"), "LightingModel.o", FLAGS),
        (os.path.join(stage, "Math", "Bounds.cpp"), "Bounds.o", FLAGS),
        (os.path.join(stage, "Raytracing", "MaterialUtils.cpp"), "MaterialUtils.o", FLAGS),
        (os.path.join(HERE, "ref_driver.cpp"), "ref_driver.o", FLAGS),
        (os.path.join(HERE, "ref_utils.cpp"), "ref_utils.o", FLAGS),
        (os.path.join(HERE, "ref_gltf.cpp"), "ref_gltf.o", FLAGS),
    ]
    objs = [os.path.join(stage, o) for _, o, _ in units]
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda u: run(["g++"] + u[2] + inc + ["-c", u[0], "-o", os.path.join(stage, u[1])]), units))
    lib = os.path.join(OUT, "libsailor_pt_ref_count.so" if count else "libsailor_pt_ref.so")
    # -Bsymbolic: the driver's thread-local rand() must bind inside this library (glm::linearRand -> std::rand).
    run(["g++", "-shared", "-o", lib] + objs + ["-Wl,-Bsymbolic", "-lpthread"])
    if not keep_stage:
        shutil.rmtree(stage, ignore_errors=True)
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--keep-stage", action="store_true")
    args = ap.parse_args()
    with ThreadPoolExecutor(max_workers=2) as ex:
        for lib in ex.map(lambda c: build(args.reference, c, args.keep_stage), (False, True)):
            print("built", lib)


if __name__ == "__main__":
    main()
