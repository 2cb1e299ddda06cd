"""Build sailor_b200/libsailor_pt_cuda.so (the product) with nvcc for sm_100a, in-tree.

    python -m sailor_b200.build [--force] [--verbose]

Flags that matter for parity (DESIGN.md "float contract"): -fmad=false (no FMA contraction, like the reference's
g++ -ffp-contract=off build), IEEE division and square root, no flush-to-zero, no fast-math.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsailor_pt_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return [os.path.join(SRC, f) for f in ("capi.cu", "backend.cu", "gltf_loader.cpp", "png_codec.cpp", "jpeg_codec.cpp", "image_io.cpp")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    newest = max(os.path.getmtime(os.path.join(SRC, f)) for f in os.listdir(SRC))
    newest = max(newest, os.path.getmtime(os.path.join(HERE, "..", "include", "sailor_pt.h")), os.path.getmtime(__file__))
    return os.path.getmtime(LIB) < newest


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (tools/trace_variants.py) — extra -D flags, written to another file."""
    if out is None and not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in sources():
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(HERE, "..", "include"), "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    link = [_nvcc(), "-shared", "-o", out or LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if out is None:
        build_cli()
    return out or LIB


CLI = os.path.join(HERE, "sailor_pt")


def build_cli():
    """sailor_b200/sailor_pt: the C++ host of csrc/cli_main.cpp, linked against the C-ABI library only (rpath $ORIGIN)."""
    cmd = ["g++", "-std=c++17", "-O2", "-o", CLI, os.path.join(SRC, "cli_main.cpp"), "-I", os.path.join(HERE, "..", "include"),
           "-L", HERE, "-lsailor_pt_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the sailor_pt host failed")
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
