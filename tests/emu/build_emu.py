#!/usr/bin/env python3
"""Build tests/emu/_build/libsailor_pt_emu.so: the product's kernel BODIES and host orchestration compiled for the
host with g++ (-DSPT_EMU), every launch a serial loop.

TEST TOOL ONLY.  The development container has no GPU, so this is how `-m "not gpu"` tests check the host logic and
the kernel arithmetic against the oracle before GPU time is spent.  It is not part of the product: the sailor_b200
package never loads it, `build()` does not ship it, and the product library has no CPU path (backend.cu, Ctx::Init).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SRC = os.path.join(ROOT, "sailor_b200", "csrc")
OUT = os.path.join(HERE, "_build")


def build():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libsailor_pt_emu.so")
    srcs = [os.path.join(SRC, f) for f in ("capi.cu", "backend.cu", "gltf_loader.cpp", "png_codec.cpp", "jpeg_codec.cpp", "image_io.cpp")]
    newest = max(os.path.getmtime(os.path.join(SRC, f)) for f in os.listdir(SRC))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "include", "sailor_pt.h")))
    if os.path.exists(lib) and os.path.getmtime(lib) > newest:
        return lib
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DSPT_EMU", "-fvisibility=hidden",
           "-I", os.path.join(ROOT, "include"), "-o", lib]
    cmd += ["-D" + d for d in os.environ.get("SPT_EMU_DEFINES", "").split()]      # tuning variants are checked on the host first
    for s in srcs:
        cmd += (["-x", "c++", s] if s.endswith(".cu") else ["-x", "c++", s])
    cmd += ["-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit("emu build failed")
    return lib


if __name__ == "__main__":
    print(build())
