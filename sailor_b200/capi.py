"""ctypes binding of include/sailor_pt.h.

The same C-ABI is exported by the product (`sailor_b200/libsailor_pt_cuda.so`) and by the test oracle
(`oracle/_ref/libsailor_pt_ref.so`); `Library(path)` binds whichever file it is given.  The product package only
ever binds its own CUDA library (see `sailor_b200/__init__.py`); the oracle is bound by tests/ and bench.py only.

Host-side mirror of the reference interface (reference Runtime/Raytracing/PathTracer.h:17-36):
`Params` keeps the reference's field names, `PathTracer.parse_command_line_args` / `PathTracer.run` keep the
reference's method names and error behaviour (return value, no exception for a bad scene: the reference logs and
returns, PathTracer.cpp:94-98).
"""
import ctypes as C
import os

import numpy as np

MATERIAL_WORDS = 32
TRI_FLOATS = 51

OK = 0
FLAG_EXACT_TRAVERSAL, FLAG_WIDE_TRAVERSAL = 1, 2
RAYS_WIDE, RAYS_ANY_HIT, RAYS_LOCAL = 1, 2, 4
ERR_ARG, ERR_IO, ERR_FORMAT, ERR_NO_DEVICE, ERR_CUDA, ERR_LIMIT, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7


class SailorPtParams(C.Structure):
    _fields_ = [
        ("pathToModel", C.c_char_p), ("output", C.c_char_p), ("camera", C.c_char_p),
        ("height", C.c_uint32), ("numSamples", C.c_uint32), ("numAmbientSamples", C.c_uint32),
        ("maxBounces", C.c_uint32), ("msaa", C.c_uint32), ("ambient", C.c_float * 3),
        ("widthOverride", C.c_uint32), ("seed", C.c_uint64),
        ("rowBegin", C.c_uint32), ("rowEnd", C.c_uint32), ("msaaBegin", C.c_uint32), ("msaaEnd", C.c_uint32),
        ("deviceCount", C.c_int32), ("flags", C.c_uint32),
    ]


class SailorPtStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("primarySamples", C.c_uint64), ("boxTests", C.c_uint64), ("triTests", C.c_uint64),
        ("secondsTotal", C.c_double), ("secondsFlatten", C.c_double), ("secondsBvhBuild", C.c_double),
        ("secondsTraverse", C.c_double), ("secondsShade", C.c_double), ("secondsOutput", C.c_double),
        ("traverseLaunches", C.c_uint32), ("kernelLaunches", C.c_uint32), ("threads", C.c_uint32),
        ("batches", C.c_uint32), ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64),
        ("secondsExpand", C.c_double), ("secondsFanOut", C.c_double), ("secondsClassify", C.c_double), ("secondsGather", C.c_double),
        ("fanOutSamples", C.c_uint64), ("secondsCall", C.c_double), ("replayedRays", C.c_uint64),
        ("devicesUsed", C.c_uint32), ("secondaryTraversal", C.c_uint32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


BVH_NODE_DTYPE = np.dtype([("aabbMin", "<f4", 3), ("leftFirst", "<u4"), ("aabbMax", "<f4", 3), ("triCount", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("baryU", "<f4"), ("baryV", "<f4"), ("triId", "<u4")])

# every symbol include/sailor_pt.h declares (tests check the built libraries export all of them)
SYMBOLS = [
    "SailorPt_ParseCommandLineArgs", "SailorPt_Run", "SailorPt_SceneLoad", "SailorPt_SceneFree",
    "SailorPt_SceneCounts", "SailorPt_SceneGetTriangles", "SailorPt_SceneGetMaterials", "SailorPt_SceneGetLights", "SailorPt_BuildBVH", "SailorPt_GetBVH",
    "SailorPt_GetCamera", "SailorPt_IntersectRays", "SailorPt_PrimaryHits", "SailorPt_Render",
    "SailorPt_OutputStage", "SailorPt_SampleTexture", "SailorPt_EvalLighting", "SailorPt_GetStats",
    "SailorPt_LastError", "SailorPt_Backend", "SailorPt_RenderResident", "SailorPt_ReadResident", "SailorPt_CopyResidentToDevice", "SailorPt_SetDevice", "SailorPt_OutputStageResident",
    "SailorPt_PinHostBuffer", "SailorPt_UnpinHostBuffer", "SailorPt_WriteImage", "SailorPt_CompareImages", "SailorPt_RenderProgressive",
    "SailorPt_TrimMemory", "SailorPt_IntersectRaysEx", "SailorPt_ShadeHits", "SailorPt_SampleGenerators", "SailorPt_DecodeImage",
]


class SailorPtError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__("%s failed with code %d%s" % (what, code, (": " + detail) if detail else ""))
        self.code = code


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class Params:
    """PathTracer::Params (reference PathTracer.h:21-32) + the documented extensions."""

    def __init__(self, path_to_model="", output="", camera="", height=512, num_samples=1, num_ambient_samples=1,
                 max_bounces=4, msaa=1, ambient=(0.0, 0.0, 0.0), width_override=0, seed=0,
                 rows=(0, 0), msaa_range=(0, 0), device_count=0, flags=0):
        self.m_pathToModel = path_to_model
        self.m_output = output
        self.m_camera = camera
        self.m_height = height
        self.m_numSamples = num_samples
        self.m_numAmbientSamples = num_ambient_samples
        self.m_maxBounces = max_bounces
        self.m_msaa = msaa
        self.m_ambient = tuple(ambient)
        self.width_override = width_override
        self.seed = seed
        self.rows = tuple(rows)
        self.msaa_range = tuple(msaa_range)
        self.device_count = device_count
        self.flags = flags

    def to_c(self):
        p = SailorPtParams()
        self._keep = [os.fsencode(self.m_pathToModel), os.fsencode(self.m_output), self.m_camera.encode()]
        p.pathToModel, p.output, p.camera = self._keep
        p.height, p.numSamples, p.numAmbientSamples = self.m_height, self.m_numSamples, self.m_numAmbientSamples
        p.maxBounces, p.msaa = self.m_maxBounces, self.m_msaa
        p.ambient[0], p.ambient[1], p.ambient[2] = self.m_ambient
        p.widthOverride, p.seed = self.width_override, self.seed
        p.rowBegin, p.rowEnd = self.rows
        p.msaaBegin, p.msaaEnd = self.msaa_range
        p.deviceCount, p.flags = self.device_count, self.flags
        return p

    @staticmethod
    def from_samples(samples, **kw):
        """`--samples n` decoding of PathTracer.cpp:48-54 (msaa = n<=32 ? min(4,n) : 8; S = max(1, lround(n/msaa))).
        numAmbientSamples := numSamples (the reference never sets it)."""
        msaa = min(4, samples) if samples <= 32 else 8
        s = max(1, int(np.float32(samples) / np.float32(msaa) + np.float32(0.5)))
        return Params(num_samples=s, num_ambient_samples=s, msaa=msaa, **kw)


class Library:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = lib = C.CDLL(path)
        P = C.POINTER
        lib.SailorPt_ParseCommandLineArgs.argtypes = [P(SailorPtParams), P(C.c_char_p), C.c_int32]
        lib.SailorPt_Run.argtypes = [P(SailorPtParams)]
        lib.SailorPt_SceneLoad.argtypes = [C.c_char_p, P(C.c_void_p)]
        lib.SailorPt_SceneFree.argtypes = [C.c_void_p]
        lib.SailorPt_SceneFree.restype = None
        lib.SailorPt_SceneCounts.argtypes = [C.c_void_p, P(C.c_uint32)]
        lib.SailorPt_SceneGetTriangles.argtypes = [C.c_void_p, P(C.c_float), P(C.c_uint8)]
        lib.SailorPt_SceneGetMaterials.argtypes = [C.c_void_p, P(C.c_uint32)]
        lib.SailorPt_SceneGetLights.argtypes = [C.c_void_p, P(C.c_float)]
        lib.SailorPt_BuildBVH.argtypes = [C.c_void_p]
        lib.SailorPt_GetBVH.argtypes = [C.c_void_p, C.c_void_p, P(C.c_uint32)]
        lib.SailorPt_GetCamera.argtypes = [C.c_void_p, P(SailorPtParams), P(C.c_uint32), P(C.c_uint32), P(C.c_float)]
        lib.SailorPt_IntersectRays.argtypes = [C.c_void_p, C.c_uint32, P(C.c_float), P(C.c_float), P(C.c_uint32), C.c_void_p]
        lib.SailorPt_IntersectRaysEx.argtypes = [C.c_void_p, C.c_uint32, P(C.c_float), P(C.c_float), P(C.c_uint32), C.c_uint32, C.c_void_p]
        lib.SailorPt_PrimaryHits.argtypes = [C.c_void_p, P(SailorPtParams), C.c_void_p]
        lib.SailorPt_Render.argtypes = [C.c_void_p, P(SailorPtParams), P(C.c_float), P(C.c_uint8)]
        lib.SailorPt_RenderResident.argtypes = [C.c_void_p, P(SailorPtParams), C.c_uint32]
        lib.SailorPt_ReadResident.argtypes = [C.c_void_p, P(C.c_float), P(C.c_uint8)]
        lib.SailorPt_CopyResidentToDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        lib.SailorPt_OutputStage.argtypes = [C.c_uint32, C.c_uint32, P(C.c_float), P(C.c_uint8)]
        lib.SailorPt_SampleTexture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, P(C.c_float), P(C.c_float)]
        lib.SailorPt_EvalLighting.argtypes = [C.c_uint32, P(C.c_float), P(C.c_float)]
        lib.SailorPt_DecodeImage.argtypes = [C.c_char_p, C.c_uint64, P(C.c_uint32), P(C.c_uint32), P(C.c_uint8), C.c_uint64]
        lib.SailorPt_ShadeHits.argtypes = [C.c_void_p, C.c_uint32, P(C.c_uint32), P(C.c_float), P(C.c_float), C.c_uint32, C.c_uint32, P(C.c_float)]
        lib.SailorPt_SampleGenerators.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, P(C.c_float)]
        lib.SailorPt_GetStats.argtypes = [P(SailorPtStats)]
        lib.SailorPt_SetDevice.argtypes = [C.c_int32]
        lib.SailorPt_OutputStageResident.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        lib.SailorPt_WriteImage.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, P(C.c_float)]
        lib.SailorPt_CompareImages.argtypes = [C.c_uint32, C.c_uint32, P(C.c_float), P(C.c_float), P(C.c_double)]
        lib.SailorPt_RenderProgressive.argtypes = [C.c_void_p, P(SailorPtParams), C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, P(C.c_float), P(C.c_uint8), P(C.c_uint32)]
        lib.SailorPt_PinHostBuffer.argtypes = [C.c_void_p, C.c_uint64]
        lib.SailorPt_UnpinHostBuffer.argtypes = [C.c_void_p]
        lib.SailorPt_LastError.restype = C.c_char_p
        lib.SailorPt_Backend.restype = C.c_char_p
        for s in SYMBOLS:
            f = getattr(lib, s)
            if s not in ("SailorPt_SceneFree", "SailorPt_LastError", "SailorPt_Backend"):
                f.restype = C.c_int32

    # -- helpers -------------------------------------------------------------------------------------------
    def check(self, rc, what):
        if rc != OK:
            raise SailorPtError(rc, what, (self.lib.SailorPt_LastError() or b"").decode(errors="replace"))

    def backend(self):
        return self.lib.SailorPt_Backend().decode()

    def set_device(self, index):
        self.check(self.lib.SailorPt_SetDevice(index), "SailorPt_SetDevice")

    def write_image(self, path, linear):
        """SailorPt_WriteImage: .pfm / .hdr linear dumps, anything else = output stage + PNG."""
        lin = np.ascontiguousarray(linear, dtype=np.float32)
        self.check(self.lib.SailorPt_WriteImage(str(path).encode(), lin.shape[1], lin.shape[0], _ptr(lin, C.c_float)), "SailorPt_WriteImage")

    def compare_images(self, a, b):
        """SailorPt_CompareImages -> dict(mean_rel_error, rmse, max_abs, psnr_db)."""
        a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
        assert a.shape == b.shape and a.ndim == 3 and a.shape[2] == 3
        m = (C.c_double * 4)()
        self.check(self.lib.SailorPt_CompareImages(a.shape[1], a.shape[0], _ptr(a, C.c_float), _ptr(b, C.c_float), m), "SailorPt_CompareImages")
        return dict(mean_rel_error=m[0], rmse=m[1], max_abs=m[2], psnr_db=m[3])

    def trim_memory(self):
        """SailorPt_TrimMemory: hand the working memory kept between frames back to the device."""
        self.check(self.lib.SailorPt_TrimMemory(), "SailorPt_TrimMemory")

    def pin_host_buffer(self, array):
        """Page-lock a numpy array the caller reuses for results (SailorPt_PinHostBuffer); unpin before dropping it."""
        self.check(self.lib.SailorPt_PinHostBuffer(array.ctypes.data, array.nbytes), "SailorPt_PinHostBuffer")

    def unpin_host_buffer(self, array):
        self.check(self.lib.SailorPt_UnpinHostBuffer(array.ctypes.data), "SailorPt_UnpinHostBuffer")

    def stats(self):
        s = SailorPtStats()
        self.check(self.lib.SailorPt_GetStats(C.byref(s)), "SailorPt_GetStats")
        return s.as_dict()

    # -- the reference entry points ------------------------------------------------------------------------
    def parse_command_line_args(self, params: Params, args):
        """PathTracer::ParseCommandLineArgs (PathTracer.cpp:30-73); args[0] is skipped like argv[0]."""
        cp = params.to_c()
        arr = (C.c_char_p * len(args))(*[a.encode() for a in args])
        self.check(self.lib.SailorPt_ParseCommandLineArgs(C.byref(cp), arr, len(args)), "SailorPt_ParseCommandLineArgs")
        params.m_pathToModel = (cp.pathToModel or b"").decode()
        params.m_output = (cp.output or b"").decode()
        params.m_camera = (cp.camera or b"").decode()
        params.m_height, params.m_numSamples, params.m_numAmbientSamples = cp.height, cp.numSamples, cp.numAmbientSamples
        params.m_maxBounces, params.m_msaa = cp.maxBounces, cp.msaa
        params.m_ambient = (cp.ambient[0], cp.ambient[1], cp.ambient[2])
        return params

    def run(self, params: Params):
        """PathTracer::Run (PathTracer.cpp:75-575). Returns the C-ABI code (0 = ok) like the reference returns quietly."""
        cp = params.to_c()
        return self.lib.SailorPt_Run(C.byref(cp))

    def load_scene(self, path):
        return Scene(self, path)

    def output_stage(self, linear):
        linear = np.ascontiguousarray(linear, dtype=np.float32)
        h, w, _ = linear.shape
        out = np.empty((h, w, 3), np.uint8)
        self.check(self.lib.SailorPt_OutputStage(w, h, _ptr(linear, C.c_float), _ptr(out, C.c_uint8)), "SailorPt_OutputStage")
        return out

    def decode_image(self, data: bytes):
        """SailorPt_DecodeImage: an image file in memory -> uint8[h, w, 4] (stbi_load_from_memory(..., 4) convention)."""
        w, h = C.c_uint32(0), C.c_uint32(0)
        self.check(self.lib.SailorPt_DecodeImage(data, len(data), C.byref(w), C.byref(h), None, 0), "SailorPt_DecodeImage")
        out = np.empty((h.value, w.value, 4), np.uint8)
        self.check(self.lib.SailorPt_DecodeImage(data, len(data), C.byref(w), C.byref(h), _ptr(out, C.c_uint8), out.nbytes), "SailorPt_DecodeImage")
        return out

    def sample_generators(self, key, kind, count):
        """SailorPt_SampleGenerators: kind 0 -> uint32[count], 1 -> float32[count], 2 -> float32[count, 4] (x, y, index of x, index of y)."""
        out = np.empty((count, 4) if kind == 2 else (count,), np.float32)
        self.check(self.lib.SailorPt_SampleGenerators(C.c_uint64(key), kind, count, _ptr(out, C.c_float)), "SailorPt_SampleGenerators")
        return out.view(np.uint32) if kind == 0 else out

    def eval_lighting(self, records):
        records = np.ascontiguousarray(records, dtype=np.float32)
        assert records.ndim == 2 and records.shape[1] == 24
        out = np.empty((records.shape[0], 28), np.float32)
        self.check(self.lib.SailorPt_EvalLighting(records.shape[0], _ptr(records, C.c_float), _ptr(out, C.c_float)), "SailorPt_EvalLighting")
        return out


class Scene:
    def __init__(self, library: Library, path):
        self.L = library
        self.h = C.c_void_p()
        library.check(library.lib.SailorPt_SceneLoad(os.fsencode(path), C.byref(self.h)), "SailorPt_SceneLoad")

    def close(self):
        if self.h:
            self.L.lib.SailorPt_SceneFree(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def counts(self):
        c = (C.c_uint32 * 6)()
        self.L.check(self.L.lib.SailorPt_SceneCounts(self.h, c), "SailorPt_SceneCounts")
        return dict(zip(("triangles", "materials", "textures", "lights", "cameras", "nodes"), list(c)))

    def triangles(self):
        n = self.counts()["triangles"]
        tris = np.empty((n, TRI_FLOATS), np.float32)
        mat = np.empty(n, np.uint8)
        self.L.check(self.L.lib.SailorPt_SceneGetTriangles(self.h, _ptr(tris, C.c_float), _ptr(mat, C.c_uint8)), "SailorPt_SceneGetTriangles")
        return tris, mat

    def materials(self):
        """(n, 32) uint32 words per material: 26 float fields then blend mode and texture slots (include/sailor_pt.h)."""
        words = np.zeros((self.counts()["materials"], MATERIAL_WORDS), np.uint32)
        if len(words):
            self.L.check(self.L.lib.SailorPt_SceneGetMaterials(self.h, _ptr(words, C.c_uint32)), "SailorPt_SceneGetMaterials")
        return words

    def lights(self):
        """(n, 6) float32: direction, intensity of every directional light."""
        out = np.zeros((self.counts()["lights"], 6), np.float32)
        if len(out):
            self.L.check(self.L.lib.SailorPt_SceneGetLights(self.h, _ptr(out, C.c_float)), "SailorPt_SceneGetLights")
        return out

    def build_bvh(self):
        self.L.check(self.L.lib.SailorPt_BuildBVH(self.h), "SailorPt_BuildBVH")

    def bvh(self):
        self.build_bvh()
        c = self.counts()
        n = c["triangles"]
        nodes = np.zeros(2 * n - 1, BVH_NODE_DTYPE)
        mapping = np.empty(n, np.uint32)
        self.L.check(self.L.lib.SailorPt_GetBVH(self.h, nodes.ctypes.data, _ptr(mapping, C.c_uint32)), "SailorPt_GetBVH")
        return nodes[:c["nodes"]], mapping

    def camera(self, params: Params):
        cp = params.to_c()
        w, h = C.c_uint32(), C.c_uint32()
        cam = (C.c_float * 12)()
        self.L.check(self.L.lib.SailorPt_GetCamera(self.h, C.byref(cp), C.byref(w), C.byref(h), cam), "SailorPt_GetCamera")
        return w.value, h.value, np.array(list(cam), np.float32)

    def intersect_rays(self, origins, directions, ignore=None, wide=False, any_hit=False, local=False):
        """BVH::IntersectBVH for a batch of rays.  wide / any_hit: SailorPt_IntersectRaysEx (the traversal variants the integrator
        uses for secondary rays)."""
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(directions, dtype=np.float32)
        n = o.shape[0]
        ig = np.ascontiguousarray(ignore, dtype=np.uint32) if ignore is not None else None
        hits = np.empty(n, HIT_DTYPE)
        if wide or any_hit or local:
            self.L.check(self.L.lib.SailorPt_IntersectRaysEx(self.h, n, _ptr(o, C.c_float), _ptr(d, C.c_float), _ptr(ig, C.c_uint32),
                                                            (RAYS_WIDE if wide else 0) | (RAYS_ANY_HIT if any_hit else 0) | (RAYS_LOCAL if local else 0), hits.ctypes.data), "SailorPt_IntersectRaysEx")
        else:
            self.L.check(self.L.lib.SailorPt_IntersectRays(self.h, n, _ptr(o, C.c_float), _ptr(d, C.c_float),
                                                          _ptr(ig, C.c_uint32), hits.ctypes.data), "SailorPt_IntersectRays")
        return hits

    def primary_hits(self, params: Params):
        w, h, _ = self.camera(params)
        cp = params.to_c()
        hits = np.empty(w * h, HIT_DTYPE)
        self.L.check(self.L.lib.SailorPt_PrimaryHits(self.h, C.byref(cp), hits.ctypes.data), "SailorPt_PrimaryHits")
        return hits.reshape(h, w)

    def render(self, params: Params, want_srgb=True, out=None):
        """out=(linear float32 [h,w,3], srgb uint8 [h,w,3] or None): caller-owned host buffers to fill (a host that renders
        frame after frame reuses them); by default fresh arrays are allocated."""
        w, h, _ = self.camera(params)
        cp = params.to_c()
        if out is not None:
            lin, srgb = out
            assert lin.shape == (h, w, 3) and lin.dtype == np.float32 and lin.flags.c_contiguous
            assert srgb is None or (srgb.shape == (h, w, 3) and srgb.dtype == np.uint8 and srgb.flags.c_contiguous)
        else:
            lin = np.empty((h, w, 3), np.float32)
            srgb = np.empty((h, w, 3), np.uint8) if want_srgb else None
        self.L.check(self.L.lib.SailorPt_Render(self.h, C.byref(cp), _ptr(lin, C.c_float), _ptr(srgb, C.c_uint8)), "SailorPt_Render")
        return lin, srgb

    def render_progressive(self, params: Params, msaa_per_pass, max_passes=0, checkpoint=None, resume=False, checkpoint_every_pass=False,
                           preview=False, want_srgb=True):
        """SailorPt_RenderProgressive -> (linear, srgb or None, primary-sample indices accumulated so far)."""
        w, h, _ = self.camera(params)
        cp = params.to_c()
        lin = np.empty((h, w, 3), np.float32)
        srgb = np.empty((h, w, 3), np.uint8) if want_srgb else None
        done = C.c_uint32(0)
        flags = (1 if resume else 0) | (2 if checkpoint_every_pass else 0) | (4 if preview else 0)
        self.L.check(self.L.lib.SailorPt_RenderProgressive(self.h, C.byref(cp), msaa_per_pass, max_passes, str(checkpoint).encode() if checkpoint else None, flags,
                                                          _ptr(lin, C.c_float), _ptr(srgb, C.c_uint8) if srgb is not None else None, C.byref(done)), "SailorPt_RenderProgressive")
        return lin, srgb, done.value

    def render_resident(self, params: Params, rebuild_bvh=False, output_stage=True):
        cp = params.to_c()
        self.L.check(self.L.lib.SailorPt_RenderResident(self.h, C.byref(cp), (1 if rebuild_bvh else 0) | (2 if output_stage else 0)), "SailorPt_RenderResident")

    def read_resident(self, params: Params, want_srgb=True):
        w, h, _ = self.camera(params)
        lin = np.empty((h, w, 3), np.float32)
        srgb = np.empty((h, w, 3), np.uint8) if want_srgb else None
        self.L.check(self.L.lib.SailorPt_ReadResident(self.h, _ptr(lin, C.c_float), _ptr(srgb, C.c_uint8)), "SailorPt_ReadResident")
        return lin, srgb

    def read_resident_into(self, lin, srgb=None):
        """SailorPt_ReadResident into caller-owned (possibly page-locked) host buffers."""
        self.L.check(self.L.lib.SailorPt_ReadResident(self.h, _ptr(lin, C.c_float), _ptr(srgb, C.c_uint8)), "SailorPt_ReadResident")

    def copy_resident_to_device(self, device_ptr, nbytes):
        self.L.check(self.L.lib.SailorPt_CopyResidentToDevice(self.h, C.c_void_p(device_ptr), nbytes), "SailorPt_CopyResidentToDevice")

    def output_stage_resident(self, device_ptr=None, nbytes=0):
        self.L.check(self.L.lib.SailorPt_OutputStageResident(self.h, C.c_void_p(device_ptr) if device_ptr else None, nbytes), "SailorPt_OutputStageResident")

    def shade_hits(self, tri_ids, bary_uv, ray_dirs, num_samples=4, num_ambient_samples=4):
        """SailorPt_ShadeHits: float32[count, 28] (layout in include/sailor_pt.h)."""
        tri = np.ascontiguousarray(tri_ids, dtype=np.uint32); uv = np.ascontiguousarray(bary_uv, dtype=np.float32); d = np.ascontiguousarray(ray_dirs, dtype=np.float32)
        assert uv.shape == (tri.shape[0], 2) and d.shape == (tri.shape[0], 3)
        out = np.empty((tri.shape[0], 28), np.float32)
        self.L.check(self.L.lib.SailorPt_ShadeHits(self.h, tri.shape[0], _ptr(tri, C.c_uint32), _ptr(uv, C.c_float), _ptr(d, C.c_float), num_samples, num_ambient_samples, _ptr(out, C.c_float)), "SailorPt_ShadeHits")
        return out

    def sample_texture(self, index, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float32)
        out = np.empty((uv.shape[0], 4), np.float32)
        self.L.check(self.L.lib.SailorPt_SampleTexture(self.h, index, uv.shape[0], _ptr(uv, C.c_float), _ptr(out, C.c_float)), "SailorPt_SampleTexture")
        return out
