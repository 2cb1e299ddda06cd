"""sailor_b200 — B200-native replacement for the CPU path tracer of aantropov/Sailor (Runtime/Raytracing).

The package is a thin host-side mirror of the reference interface (reference Runtime/Raytracing/PathTracer.h:17-36)
over the C-ABI shared library `libsailor_pt_cuda.so` (include/sailor_pt.h), which holds the CUDA kernels for sm_100a.
There is NO CPU execution path: if the library has not been built, or no CUDA device is present, calls fail loudly.

    from sailor_b200 import PathTracer, Params
    p = Params()
    PathTracer.ParseCommandLineArgs(p, ["sailor", "--in", "scene.glb", "--out", "out.png", "--height", "512",
                                        "--samples", "16", "--bounces", "4", "--ambient", "ffffff"])
    PathTracer().Run(p)
"""
import os

from .capi import (ERR_ARG, ERR_CUDA, ERR_FORMAT, ERR_IO, ERR_LIMIT, ERR_NO_DEVICE, ERR_UNSUPPORTED, OK, Library, Params,
                   SailorPtError, Scene)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsailor_pt_cuda.so")
_lib = None


def library() -> Library:
    """The product library. Raises if it is missing — build it with `python -m sailor_b200.build`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("sailor_b200: %s is missing (run `python -m sailor_b200.build`); there is no CPU fallback" % LIB_PATH)
        _lib = Library(LIB_PATH)
        if not _lib.backend().startswith("cuda"):
            raise RuntimeError("sailor_b200: %s is not the CUDA build (%s)" % (LIB_PATH, _lib.backend()))
    return _lib


class PathTracer:
    """Mirror of Sailor::Raytracing::PathTracer (reference PathTracer.h:17-36): same two entry points."""

    Params = Params

    @staticmethod
    def ParseCommandLineArgs(params: Params, args):
        return library().parse_command_line_args(params, list(args))

    def Run(self, params: Params):
        rc = library().run(params)
        if rc != OK:
            # the reference logs and returns (PathTracer.cpp:94-98); surface the reason without raising for scene errors
            import sys
            sys.stderr.write("PathTracer::Run failed (%d): %s\n" % (rc, library().lib.SailorPt_LastError().decode(errors="replace")))
        return rc


def load_scene(path) -> Scene:
    return library().load_scene(path)
