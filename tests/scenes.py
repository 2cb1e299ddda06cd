"""Deterministic synthetic glTF scenes for tests and bench (no network, no reference checkout needed at run time).

* cube()              — the geometry of the reference's smallest bundled model (Content/Models/Box/Box.gltf: 24 vertices,
                        12 triangles, one red dielectric material, a Y-up root rotation, no camera, no lights), generated
                        from the cube's definition so BASELINE configs C1/C2 can run on the GPU box.
                        tests/test_scenes.py checks here (where the reference exists) that it flattens bit-identically.
* heightfield(n)      — SURVEY §8(d) config C3: (n x n) grid over [-0.5,0.5]^2, heights 0.05*U[0,1) from the LCG
                        s = s*1664525 + 1013904223 (seed 12345), one diffuse material, one directional light
                        (0.3,-1,0.2)/|.|, camera at (0,0.6,0.9) looking along (0,-0.55,-0.9).  n=707 -> 999,698 triangles.
* pbr_scene()         — small textured scene that reaches every integrator branch: base-colour / normal /
                        metallic-roughness / emissive textures (procedural PNGs), KHR_texture_transform, an alpha-BLEND
                        quad, a MASK quad, a thick transmissive (volume) box, a mirror, emissive quads, two directional
                        lights, a named camera.
* gallery()           — SURVEY §8(d) config C4: an 8 x 8 floor of PBR tiles, every tile with its OWN material (baseColor +
                        normal + occlusion/roughness/metallic textures, procedural, seed 1) and a box of that material on
                        it, 32 emissive quads (each with its own emissive texture) hovering above, four directional
                        lights, one camera: 96 materials, 224 textures (the reference's u8 slots allow 255).
* instanced_heightfield() — config C5: ONE heightfield mesh (LCG seed 2) instanced by `instances` nodes with their own
                        translation / rotation / scale (node-level instancing is what glTF offers); 10 instances of
                        n=707 -> 9,996,980 triangles.
All files are GLB with embedded buffers, written into a caller-supplied directory.
"""
import json
import math
import os
import struct
import zlib

import numpy as np


# ----------------------------------------------------------------------------------------------- GLB writer
class GlbBuilder:
    def __init__(self):
        self.bin = bytearray()
        self.j = {"asset": {"version": "2.0", "generator": "sailor_b200 tests/scenes.py"}, "scene": 0, "scenes": [{"nodes": []}],
                  "nodes": [], "meshes": [], "accessors": [], "bufferViews": [], "buffers": [], "materials": []}

    def _view(self, data: bytes, target=None):
        while len(self.bin) % 4:
            self.bin.append(0)
        v = {"buffer": 0, "byteOffset": len(self.bin), "byteLength": len(data)}
        if target:
            v["target"] = target
        self.bin += data
        self.j["bufferViews"].append(v)
        return len(self.j["bufferViews"]) - 1

    def accessor(self, arr: np.ndarray, typ: str, target=None, minmax=False):
        arr = np.ascontiguousarray(arr)
        ct = {np.dtype("float32"): 5126, np.dtype("uint32"): 5125, np.dtype("uint16"): 5123, np.dtype("uint8"): 5121}[arr.dtype]
        a = {"bufferView": self._view(arr.tobytes(), target), "componentType": ct, "count": int(arr.shape[0]), "type": typ}
        if minmax:
            a["min"] = [float(x) for x in arr.reshape(arr.shape[0], -1).min(axis=0)]
            a["max"] = [float(x) for x in arr.reshape(arr.shape[0], -1).max(axis=0)]
        self.j["accessors"].append(a)
        return len(self.j["accessors"]) - 1

    def mesh(self, pos, idx, material, nrm=None, uv=None, tan=None, uv1=None):
        attrs = {"POSITION": self.accessor(np.asarray(pos, np.float32), "VEC3", 34962, True)}
        if nrm is not None:
            attrs["NORMAL"] = self.accessor(np.asarray(nrm, np.float32), "VEC3", 34962)
        if uv is not None:
            attrs["TEXCOORD_0"] = self.accessor(np.asarray(uv, np.float32), "VEC2", 34962)
        if uv1 is not None:
            attrs["TEXCOORD_1"] = self.accessor(np.asarray(uv1, np.float32), "VEC2", 34962)
        if tan is not None:
            attrs["TANGENT"] = self.accessor(np.asarray(tan, np.float32), "VEC4", 34962)
        prim = {"attributes": attrs, "mode": 4, "material": material}
        if idx is not None:
            idx = np.asarray(idx)
            idx = idx.astype(np.uint16) if idx.max() < 65536 else idx.astype(np.uint32)
            prim["indices"] = self.accessor(idx.reshape(-1), "SCALAR", 34963)
        self.j["meshes"].append({"primitives": [prim]})
        return len(self.j["meshes"]) - 1

    def node(self, root=True, **kw):
        self.j["nodes"].append(kw)
        i = len(self.j["nodes"]) - 1
        if root:
            self.j["scenes"][0]["nodes"].append(i)
        return i

    def material(self, **kw):
        self.j["materials"].append(kw)
        return len(self.j["materials"]) - 1

    def texture(self, png: bytes, wrap=10497):
        mime = "image/jpeg" if png[:3] == b"\xff\xd8\xff" else ("image/vnd.radiance" if png[:2] == b"#?" else "image/png")
        self.j.setdefault("images", []).append({"bufferView": self._view(png), "mimeType": mime})
        self.j.setdefault("samplers", []).append({"magFilter": 9729, "minFilter": 9729, "wrapS": wrap, "wrapT": wrap})
        self.j.setdefault("textures", []).append({"sampler": len(self.j["samplers"]) - 1, "source": len(self.j["images"]) - 1})
        return len(self.j["textures"]) - 1

    def write(self, path):
        self.j["buffers"] = [{"byteLength": len(self.bin)}]
        js = json.dumps(self.j, separators=(",", ":")).encode()
        js += b" " * ((4 - len(js) % 4) % 4)
        b = bytes(self.bin) + b"\0" * ((4 - len(self.bin) % 4) % 4)
        with open(path, "wb") as f:
            f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(b)))
            f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
            f.write(struct.pack("<II", len(b), 0x004E4942) + b)
        return path


def png_bytes(img: np.ndarray, level=6) -> bytes:
    """8-bit RGB / RGBA PNG (filter 0)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    rows = np.zeros((h, 1 + w * c), np.uint8)
    rows[:, 1:] = img.reshape(h, w * c)
    raw = rows.tobytes()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, level)) + chunk(b"IEND", b""))


def quat_from_to_neg_z(direction):
    """Unit quaternion (x,y,z,w) rotating (0,0,-1) onto `direction`."""
    d = np.asarray(direction, np.float64)
    d = d / np.linalg.norm(d)
    f = np.array([0.0, 0.0, -1.0])
    c = float(np.dot(f, d))
    if c < -0.999999:
        return [0.0, 1.0, 0.0, 0.0]
    ax = np.cross(f, d)
    q = np.array([ax[0], ax[1], ax[2], 1.0 + c])
    q /= np.linalg.norm(q)
    return [float(v) for v in q]


def look_at_quat(forward, up=(0, 1, 0)):
    """Quaternion whose rotation maps -Z to `forward` and +Y close to `up` (glTF camera convention)."""
    f = np.asarray(forward, np.float64); f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64)); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.array([r, u, -f]).T          # columns: +X, +Y, +Z images
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    return [float(v) for v in q]


# ----------------------------------------------------------------------------------------------- scenes
def _cube_arrays(half=0.5):
    # per face: normal, then the four corners as signs along the two other axes, in Box.gltf's corner order
    faces = [
        ((0, 0, 1), (0, 1), [(-1, -1), (1, -1), (-1, 1), (1, 1)]),
        ((0, -1, 0), (0, 2), [(1, 1), (-1, 1), (1, -1), (-1, -1)]),
        ((1, 0, 0), (1, 2), [(1, 1), (-1, 1), (1, -1), (-1, -1)]),
        ((0, 1, 0), (0, 2), [(-1, 1), (1, 1), (-1, -1), (1, -1)]),
        ((-1, 0, 0), (1, 2), [(-1, 1), (1, 1), (-1, -1), (1, -1)]),
        ((0, 0, -1), (0, 1), [(-1, -1), (-1, 1), (1, -1), (1, 1)]),
    ]
    pos, nrm, idx = [], [], []
    for f, (n, axes, corners) in enumerate(faces):
        for sa, sb in corners:
            p = [c * half for c in n]
            p[axes[0]] = sa * half
            p[axes[1]] = sb * half
            pos.append(p)
            nrm.append(list(n))
        idx += [4 * f + k for k in (0, 1, 2, 3, 2, 1)]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(idx, np.uint16)


def cube(path):
    g = GlbBuilder()
    pos, nrm, idx = _cube_arrays()
    mat = g.material(pbrMetallicRoughness={"baseColorFactor": [0.800000011920929, 0.0, 0.0, 1.0], "metallicFactor": 0.0}, name="Red")
    mesh = g.mesh(pos, idx, mat, nrm=nrm)
    child = g.node(root=False, mesh=mesh)
    g.node(children=[child], matrix=[1.0, 0.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])
    return g.write(path)


def zero_normal_cube(path):
    """The cube with a NORMAL accessor of zeros on a floor quad: every shading normal on the cube is normalize(0) = NaN, so
    LightingModel::Sample can never return a valid direction there and the reference's rejection loop (PathTracer.cpp:761-767) would
    spin forever.  The product caps the loop at 4096 tries; tests check that such a frame finishes and stays finite."""
    g = GlbBuilder()
    pos, nrm, idx = _cube_arrays(0.3)
    mat = g.material(pbrMetallicRoughness={"baseColorFactor": [0.8, 0.8, 0.8, 1.0], "metallicFactor": 0.0, "roughnessFactor": 0.5})
    g.node(mesh=g.mesh(pos, idx, mat, nrm=np.zeros_like(nrm)), translation=[0.0, 0.3, 0.0])
    floor = np.array([[-2, 0, -2], [2, 0, -2], [-2, 0, 2], [2, 0, 2]], np.float32)
    g.node(mesh=g.mesh(floor, np.array([0, 2, 1, 1, 2, 3], np.uint16), mat, nrm=np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))))
    g.j["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.7, "aspectRatio": 1.5, "znear": 0.01, "zfar": 100.0}}]
    g.node(camera=0, name="main", translation=[0.0, 1.0, 2.2], rotation=look_at_quat((0.0, -0.35, -1.0)))
    return g.write(path)


def default_material_scene(path, with_materials=False):
    """Three cubes that exercise the DEFAULT material and malformed attribute streams (the reference's Assimp front end appends
    a default material; a primitive without `material`, or with an index outside the array, uses it):
      with_materials=False  the file has NO `materials` array at all; one primitive also carries a NORMAL accessor whose count
                            differs from POSITION's (ignored: flat normals are generated)
      with_materials=True   one real material; the primitives use it, omit `material`, or point past the array."""
    g = GlbBuilder()
    pos, nrm, idx = _cube_arrays(0.3)
    mats = [None, None, None]
    if with_materials:
        mats = [g.material(pbrMetallicRoughness={"baseColorFactor": [0.1, 0.6, 0.2, 1.0], "metallicFactor": 0.0, "roughnessFactor": 0.7}), None, 7]
    else:
        del g.j["materials"]
    for k, m in enumerate(mats):
        mesh = g.mesh(pos, idx, 0, nrm=(nrm[:-2] if (k == 1 and not with_materials) else nrm))
        prim = g.j["meshes"][mesh]["primitives"][0]
        if m is None:
            del prim["material"]
        else:
            prim["material"] = m
        g.node(mesh=mesh, translation=[(k - 1) * 0.8, 0.0, 0.0])
    g.j["extensions"] = {"KHR_lights_punctual": {"lights": [{"type": "directional", "color": [1.0, 1.0, 1.0], "intensity": 1366.0}]}}
    g.j["extensionsUsed"] = ["KHR_lights_punctual"]
    g.node(name="sun", rotation=quat_from_to_neg_z((0.3, -1.0, -0.4)), extensions={"KHR_lights_punctual": {"light": 0}})
    return g.write(path)


def lcg_uniform(count, seed=12345):
    """U[0,1) floats from s = s*1664525 + 1013904223 (mod 2^32), value = s / 2^32."""
    out = np.empty(count, np.float64)
    s = seed & 0xFFFFFFFF
    for i in range(count):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = s / 4294967296.0
    return out


def lcg_uniform_fast(count, seed=12345):
    """Same sequence as lcg_uniform, vectorised by jumping: s_k = A^k s_0 + C (A^k - 1)/(A - 1) mod 2^32."""
    a, c = np.uint64(1664525), np.uint64(1013904223)
    mask = np.uint64(0xFFFFFFFF)
    # doubling tables
    k = np.arange(1, count + 1, dtype=np.uint64)
    mul = np.ones(count, np.uint64)
    add = np.zeros(count, np.uint64)
    pa, pc = a, c
    bit = np.uint64(1)
    kk = k.copy()
    while kk.any():
        sel = (kk & bit) != 0
        # compose (mul,add) with (pa,pc): x -> pa*(mul*x+add)+pc
        mul = np.where(sel, (mul * pa) & mask, mul)
        add = np.where(sel, (add * pa + pc) & mask, add)
        pc = (pa * pc + pc) & mask
        pa = (pa * pa) & mask
        kk = kk & ~bit
        bit = bit << np.uint64(1)
    s = (mul * np.uint64(seed) + add) & mask
    return s.astype(np.float64) / 4294967296.0


def heightfield(path, n=64, seed=12345, with_light=True):
    """(n x n) quads -> 2*n*n triangles. Vertex (i,j) at x = -0.5 + i/n, z = -0.5 + j/n, y = 0.05*U."""
    g = GlbBuilder()
    nv = n + 1
    u = lcg_uniform_fast(nv * nv, seed)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    pos = np.stack([-0.5 + ii / n, 0.05 * u.reshape(nv, nv), -0.5 + jj / n], axis=-1).reshape(-1, 3).astype(np.float32)
    q = (jj[:-1, :-1] * nv + ii[:-1, :-1]).reshape(-1)
    idx = np.stack([q, q + nv, q + 1, q + 1, q + nv, q + nv + 1], axis=-1).reshape(-1).astype(np.uint32)
    mat = g.material(pbrMetallicRoughness={"baseColorFactor": [0.7, 0.7, 0.7, 1.0], "metallicFactor": 0.0, "roughnessFactor": 1.0})
    g.node(mesh=g.mesh(pos, idx, mat))          # no NORMAL: flat face normals (aiProcess_GenNormals)
    g.j["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.8, "aspectRatio": 16.0 / 9.0, "znear": 0.01, "zfar": 100.0}}]
    g.node(camera=0, name="main", translation=[0.0, 0.6, 0.9], rotation=look_at_quat((0.0, -0.55, -0.9)))
    if with_light:
        g.j["extensionsUsed"] = ["KHR_lights_punctual"]
        g.j["extensions"] = {"KHR_lights_punctual": {"lights": [{"type": "directional", "color": [1.0, 1.0, 1.0], "intensity": 2049.0}]}}
        g.node(name="sun", rotation=quat_from_to_neg_z((0.3, -1.0, 0.2)), extensions={"KHR_lights_punctual": {"light": 0}})
    return g.write(path)


def _checker(size, seed, kind):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:size, 0:size]
    if kind == "base":
        a = ((x // 8 + y // 8) % 2).astype(np.float32)
        img = np.stack([60 + 160 * a, 200 - 120 * a, 90 + 100 * np.sin(x / 5.0) ** 2, np.full_like(a, 255)], -1)
        img[..., 3] = np.where((x + y) % 32 < 16, 255, 110)
    elif kind == "normal":
        nx = 0.25 * np.sin(x / 3.0); ny = 0.25 * np.cos(y / 4.0); nz = np.sqrt(1 - nx * nx - ny * ny)
        img = np.stack([(nx + 1) * 127.5, (ny + 1) * 127.5, (nz + 1) * 127.5], -1)
    elif kind == "orm":
        img = np.stack([np.full((size, size), 255.0), 40 + 200 * ((x // 16) % 2), 255 * ((y // 16) % 2)], -1)
    else:  # emissive
        img = np.stack([255 * ((x // 4 + y // 4) % 2), 180 * ((x // 4) % 2), 40 + 0 * x], -1).astype(np.float32)
    img = img + rng.randint(0, 3, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def _quad(center, ux, uy, uv_scale=1.0):
    c, ux, uy = (np.asarray(v, np.float32) for v in (center, ux, uy))
    pos = np.array([c - ux - uy, c + ux - uy, c - ux + uy, c + ux + uy], np.float32)
    n = np.cross(ux, uy); n = n / np.linalg.norm(n)
    nrm = np.tile(n, (4, 1)).astype(np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32) * uv_scale
    idx = np.array([0, 1, 2, 3, 2, 1], np.uint16)
    return pos, nrm, uv, idx


def jpeg_bytes(img: np.ndarray, **kw) -> bytes:
    """RGB8 -> JPEG file (Pillow; baseline 4:2:0 unless told otherwise)."""
    import io
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(img[..., :3])).save(b, "JPEG", **({"quality": 88} | kw))
    return b.getvalue()


def hdr_bytes(img: np.ndarray) -> bytes:
    """RGB8 -> Radiance .hdr file (new-style RLE scanlines, written by OpenCV) holding the values img / 64 (so some exceed 1)."""
    import cv2
    ok, enc = cv2.imencode(".hdr", (img[..., 2::-1].astype(np.float32) / np.float32(64.0)))
    assert ok
    return enc.tobytes()


def pbr_scene(path, tex_size=64, jpeg=False, hdr=False):
    """jpeg=True: the base-colour, ORM and emissive textures are JPEG files (baseline 4:2:0, progressive 4:4:4, baseline 4:2:2): the most
    common glTF texture format, decoded by csrc/jpeg_codec.cpp; the normal map stays PNG.
    hdr=True: the base-colour and the emissive textures are Radiance .hdr files (the reference keeps those as float texels,
    MaterialUtils.h:224-229, 250-253)."""
    g = GlbBuilder()
    if jpeg or hdr:
        enc = {1: dict(), 3: dict(progressive=True, subsampling=0), 4: dict(subsampling=1)} if jpeg else {1: None, 4: None}
        global png_bytes
        _png = png_bytes
        counter = [0]

        def as_file(img, level=6):
            counter[0] += 1
            if counter[0] in enc:
                return hdr_bytes(img) if hdr else jpeg_bytes(img, **enc[counter[0]])
            return _png(img, level)
        png_bytes = as_file
        try:
            return _pbr_scene_body(g, path, tex_size)
        finally:
            png_bytes = _png
    return _pbr_scene_body(g, path, tex_size)


def _pbr_scene_body(g, path, tex_size):
    g.j["extensionsUsed"] = ["KHR_lights_punctual", "KHR_materials_transmission", "KHR_materials_volume", "KHR_materials_ior",
                             "KHR_materials_emissive_strength", "KHR_texture_transform"]
    tb = g.texture(png_bytes(_checker(tex_size, 1, "base")))
    tn = g.texture(png_bytes(_checker(tex_size, 2, "normal")))
    to = g.texture(png_bytes(_checker(tex_size, 3, "orm")), wrap=33071)
    te = g.texture(png_bytes(_checker(tex_size, 4, "emissive")))
    m_floor = g.material(pbrMetallicRoughness={"baseColorTexture": {"index": tb, "extensions": {"KHR_texture_transform": {"offset": [0.1, 0.2], "scale": [2.0, 3.0], "rotation": 0.3}}},
                                               "metallicRoughnessTexture": {"index": to}, "metallicFactor": 0.9, "roughnessFactor": 0.8},
                         normalTexture={"index": tn})
    m_blend = g.material(pbrMetallicRoughness={"baseColorFactor": [0.2, 0.5, 0.9, 0.45], "metallicFactor": 0.0, "roughnessFactor": 0.6}, alphaMode="BLEND")
    m_mask = g.material(pbrMetallicRoughness={"baseColorTexture": {"index": tb}, "metallicFactor": 0.0}, alphaMode="MASK", alphaCutoff=0.6)
    m_glass = g.material(pbrMetallicRoughness={"baseColorFactor": [0.95, 0.98, 1.0, 1.0], "metallicFactor": 0.0, "roughnessFactor": 0.05},
                         extensions={"KHR_materials_transmission": {"transmissionFactor": 0.9}, "KHR_materials_ior": {"ior": 1.45},
                                     "KHR_materials_volume": {"thicknessFactor": 0.2, "attenuationColor": [0.6, 0.9, 0.7], "attenuationDistance": 0.5}})
    m_mirror = g.material(pbrMetallicRoughness={"baseColorFactor": [0.9, 0.9, 0.9, 1.0], "metallicFactor": 1.0, "roughnessFactor": 0.0})
    m_emit = g.material(pbrMetallicRoughness={"baseColorFactor": [0.1, 0.1, 0.1, 1.0], "metallicFactor": 0.0}, emissiveFactor=[1.0, 0.8, 0.5],
                        emissiveTexture={"index": te}, extensions={"KHR_materials_emissive_strength": {"emissiveStrength": 3.0}})
    m_rough = g.material(pbrMetallicRoughness={"baseColorFactor": [0.8, 0.3, 0.2, 1.0], "metallicFactor": 0.3, "roughnessFactor": 0.15})
    m_thin = g.material(pbrMetallicRoughness={"baseColorFactor": [0.9, 0.7, 0.3, 1.0], "metallicFactor": 0.0, "roughnessFactor": 0.4},
                        extensions={"KHR_materials_transmission": {"transmissionFactor": 0.6}})

    p, n, uv, i = _quad((0, 0, 0), (1.5, 0, 0), (0, 0, -1.5))
    g.node(mesh=g.mesh(p, i, m_floor, nrm=n, uv=uv))
    p, n, uv, i = _quad((0, 0.75, -1.5), (1.5, 0, 0), (0, 0.75, 0))
    g.node(mesh=g.mesh(p, i, m_mirror, nrm=n, uv=uv))
    p, n, uv, i = _quad((-0.6, 0.45, 0.3), (0.3, 0, 0), (0, 0.3, 0))
    g.node(mesh=g.mesh(p, i, m_blend, nrm=n, uv=uv))
    p, n, uv, i = _quad((0.7, 0.4, 0.2), (0.25, 0, 0.1), (0, 0.3, 0))
    g.node(mesh=g.mesh(p, i, m_mask, nrm=n, uv=uv))
    p, n, uv, i = _quad((-1.5, 0.6, -0.3), (0, 0, 0.5), (0, 0.4, 0))
    g.node(mesh=g.mesh(p, i, m_emit, nrm=n, uv=uv))
    p, n, uv, i = _quad((0.0, 0.35, 0.9), (0.2, 0, 0), (0, 0.2, 0.05))
    g.node(mesh=g.mesh(p, i, m_thin, nrm=n, uv=uv))
    cp, cn, ci = _cube_arrays(0.25)
    glass = g.mesh(cp, ci, m_glass, nrm=cn)
    g.node(mesh=glass, translation=[0.0, 0.27, -0.2], rotation=[0.0, 0.3826834, 0.0, 0.9238795])
    rough = g.mesh(cp, ci, m_rough, nrm=cn)
    parent = g.node(translation=[0.8, 0.0, -0.7], scale=[0.8, 1.4, 0.8], children=[])
    child = g.node(root=False, mesh=rough, translation=[0.0, 0.25, 0.0], rotation=[0.0, 0.2588190, 0.0, 0.9659258])
    g.j["nodes"][parent]["children"] = [child]
    g.j["cameras"] = [{"type": "perspective", "name": "wide", "perspective": {"yfov": 1.0, "znear": 0.01, "zfar": 100.0}},
                      {"type": "perspective", "perspective": {"yfov": 0.7, "aspectRatio": 1.25, "znear": 0.01, "zfar": 100.0}}]
    g.node(camera=0, name="wide_cam", translation=[0.0, 1.0, 3.0], rotation=look_at_quat((0.0, -0.2, -1.0)))
    g.node(camera=1, name="main_cam", translation=[0.3, 0.9, 2.4], rotation=look_at_quat((-0.1, -0.25, -1.0)))
    g.j["extensions"] = {"KHR_lights_punctual": {"lights": [
        {"type": "directional", "color": [1.0, 0.95, 0.9], "intensity": 1500.0},
        {"type": "point", "color": [1.0, 1.0, 1.0], "intensity": 10.0},
        {"type": "directional", "color": [0.4, 0.5, 1.0], "intensity": 700.0}]}}
    g.node(name="sun", rotation=quat_from_to_neg_z((-0.4, -1.0, -0.3)), extensions={"KHR_lights_punctual": {"light": 0}})
    g.node(name="bulb", translation=[0, 1, 0], extensions={"KHR_lights_punctual": {"light": 1}})
    g.node(name="fill", rotation=quat_from_to_neg_z((0.6, -0.5, -0.6)), extensions={"KHR_lights_punctual": {"light": 2}})
    return g.write(path)


def _gallery_texture(size, seed, kind):
    """Procedural tile textures, different for every seed (periods and phases drawn from RandomState(seed))."""
    r = np.random.RandomState(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / np.float32(size)
    f = r.randint(2, 9, 4).astype(np.float32); ph = r.uniform(0, 6.28, 4).astype(np.float32); col = r.uniform(0.15, 1.0, (2, 3)).astype(np.float32)
    a = 0.5 + 0.5 * np.sin(6.2831853 * f[0] * x + ph[0]) * np.sin(6.2831853 * f[1] * y + ph[1])
    b = ((np.floor(x * f[2] * 2) + np.floor(y * f[3] * 2)) % 2).astype(np.float32)
    if kind == "base":
        img = (col[0][None, None, :] * a[..., None] + col[1][None, None, :] * (1 - a[..., None])) * (0.6 + 0.4 * b[..., None]) * 255
    elif kind == "normal":
        nx = 0.3 * np.cos(6.2831853 * f[0] * x + ph[0]) * np.sin(6.2831853 * f[1] * y + ph[1]); ny = 0.3 * np.sin(6.2831853 * f[0] * x + ph[0]) * np.cos(6.2831853 * f[1] * y + ph[1])
        img = np.stack([(nx + 1) * 127.5, (ny + 1) * 127.5, (np.sqrt(1 - nx * nx - ny * ny) + 1) * 127.5], -1)
    elif kind == "orm":
        img = np.stack([np.full_like(a, 255.0), 30 + 220 * a, 255 * b * (r.uniform() < 0.5)], -1)
    else:  # emissive
        img = col[0][None, None, :] * (b * a)[..., None] * 255
    return np.clip(img, 0, 255).astype(np.uint8)


def gallery(path, tiles=8, emitters=32, tex_size=1024, seed=1):
    g = GlbBuilder()
    g.j["extensionsUsed"] = ["KHR_lights_punctual", "KHR_materials_emissive_strength", "KHR_texture_transform"]
    r = np.random.RandomState(seed)
    cp, cn, ci = _cube_arrays(0.5)
    cuv = np.tile(np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32), (6, 1))
    box = None
    span = 1.0 / tiles
    for ty in range(tiles):
        for tx in range(tiles):
            k = ty * tiles + tx
            tb = g.texture(png_bytes(_gallery_texture(tex_size, seed * 1000 + 3 * k, "base"), 1))
            tn = g.texture(png_bytes(_gallery_texture(tex_size, seed * 1000 + 3 * k + 1, "normal"), 1))
            to = g.texture(png_bytes(_gallery_texture(tex_size, seed * 1000 + 3 * k + 2, "orm"), 1), wrap=33071 if k % 2 else 10497)
            base = {"index": tb}
            if k % 4 == 1:
                base["extensions"] = {"KHR_texture_transform": {"offset": [0.25, 0.5], "scale": [2.0, 2.0], "rotation": 0.1 * k}}
            m = g.material(pbrMetallicRoughness={"baseColorTexture": base, "metallicRoughnessTexture": {"index": to},
                                                 "metallicFactor": float(r.uniform(0.0, 1.0)), "roughnessFactor": float(r.uniform(0.2, 1.0))},
                           normalTexture={"index": tn})
            c = (-0.5 + (tx + 0.5) * span, 0.0, -0.5 + (ty + 0.5) * span)
            p, n, uv, i = _quad(c, (span / 2, 0, 0), (0, 0, -span / 2))
            g.node(mesh=g.mesh(p, i, m, nrm=n, uv=uv))
            h = float(r.uniform(0.02, 0.09)); sx = float(r.uniform(0.25, 0.6)) * span
            q = quat_from_to_neg_z((math.sin(0.7 * k), 0.0, -math.cos(0.7 * k)))
            g.node(mesh=g.mesh(cp, ci, m, nrm=cn, uv=cuv), translation=[c[0], h / 2, c[2]], rotation=q, scale=[sx, h, sx])
    for e in range(emitters):
        te = g.texture(png_bytes(_gallery_texture(tex_size, seed * 1000 + 500 + e, "emissive"), 1))
        m = g.material(pbrMetallicRoughness={"baseColorFactor": [0.05, 0.05, 0.05, 1.0], "metallicFactor": 0.0},
                       emissiveFactor=[1.0, 1.0, 1.0], emissiveTexture={"index": te},
                       extensions={"KHR_materials_emissive_strength": {"emissiveStrength": float(r.uniform(2.0, 8.0))}})
        c = (float(r.uniform(-0.45, 0.45)), float(r.uniform(0.15, 0.3)), float(r.uniform(-0.45, 0.45)))
        p, n, uv, i = _quad(c, (0.03, 0, 0), (0, 0, 0.03))           # facing down
        g.node(mesh=g.mesh(p, i, m, nrm=n, uv=uv))
    g.j["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.75, "aspectRatio": 16.0 / 9.0, "znear": 0.01, "zfar": 100.0}}]
    g.node(camera=0, name="main", translation=[0.0, 0.55, 0.95], rotation=look_at_quat((0.0, -0.5, -0.9)))
    dirs = [(0.3, -1.0, 0.2), (-0.5, -0.8, -0.1), (0.1, -0.6, -0.7), (-0.2, -1.0, 0.6)]
    g.j["extensions"] = {"KHR_lights_punctual": {"lights": [
        {"type": "directional", "color": [1.0, 0.96 - 0.1 * k, 0.9 - 0.15 * k], "intensity": 683.0 * (0.8 - 0.15 * k)} for k in range(len(dirs))]}}
    for k, d in enumerate(dirs):
        g.node(name="sun%d" % k, rotation=quat_from_to_neg_z(d), extensions={"KHR_lights_punctual": {"light": k}})
    return g.write(path)


def instanced_heightfield(path, n=707, instances=10, seed=2):
    g = GlbBuilder()
    nv = n + 1
    u = lcg_uniform_fast(nv * nv, seed)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    pos = np.stack([-0.5 + ii / n, 0.05 * u.reshape(nv, nv), -0.5 + jj / n], axis=-1).reshape(-1, 3).astype(np.float32)
    q = (jj[:-1, :-1] * nv + ii[:-1, :-1]).reshape(-1)
    idx = np.stack([q, q + nv, q + 1, q + 1, q + nv, q + nv + 1], axis=-1).reshape(-1).astype(np.uint32)
    mats = [g.material(pbrMetallicRoughness={"baseColorFactor": [0.4 + 0.05 * k, 0.75 - 0.04 * k, 0.5, 1.0], "metallicFactor": 0.1 * (k % 3), "roughnessFactor": 0.5 + 0.05 * (k % 5)})
            for k in range(min(instances, 4))]
    # one set of accessors; one mesh object per material, all sharing them (instancing at the node level)
    first = g.mesh(pos, idx, mats[0])
    meshes = [first]
    for m in mats[1:]:
        prim = dict(g.j["meshes"][first]["primitives"][0]); prim["material"] = m
        g.j["meshes"].append({"primitives": [prim]}); meshes.append(len(g.j["meshes"]) - 1)
    cols = int(math.ceil(math.sqrt(instances)))
    for k in range(instances):
        cx, cz = k % cols, k // cols
        ang = 0.35 * k
        g.node(mesh=meshes[k % len(meshes)], translation=[(cx - (cols - 1) / 2) * 0.98, 0.01 * k, -(cz * 0.98)],
               rotation=[0.0, math.sin(ang / 2), 0.0, math.cos(ang / 2)], scale=[1.0, 1.0 + 0.2 * (k % 3), 1.0])
    g.j["cameras"] = [{"type": "perspective", "perspective": {"yfov": 0.8, "aspectRatio": 16.0 / 9.0, "znear": 0.01, "zfar": 100.0}}]
    g.node(camera=0, name="main", translation=[0.0, 1.1, 1.6], rotation=look_at_quat((0.0, -0.5, -0.9)))
    g.j["extensionsUsed"] = ["KHR_lights_punctual"]
    g.j["extensions"] = {"KHR_lights_punctual": {"lights": [{"type": "directional", "color": [1.0, 1.0, 1.0], "intensity": 2049.0}]}}
    g.node(name="sun", rotation=quat_from_to_neg_z((0.3, -1.0, 0.2)), extensions={"KHR_lights_punctual": {"light": 0}})
    return g.write(path)


def ensure(directory, name, **kw):
    """Create (once) and return the path of a named scene."""
    os.makedirs(directory, exist_ok=True)
    if name == "cube":
        p = os.path.join(directory, "cube.glb")
        return p if os.path.exists(p) else cube(p)
    if name == "pbr":
        p = os.path.join(directory, "pbr.glb")
        return p if os.path.exists(p) else pbr_scene(p)
    if name == "pbr_jpeg":
        p = os.path.join(directory, "pbr_jpeg.glb")
        return p if os.path.exists(p) else pbr_scene(p, jpeg=True)
    if name == "pbr_hdr":
        p = os.path.join(directory, "pbr_hdr.glb")
        return p if os.path.exists(p) else pbr_scene(p, hdr=True)
    if name == "heightfield":
        n = kw.get("n", 64)
        p = os.path.join(directory, "heightfield_%d.glb" % n)
        return p if os.path.exists(p) else heightfield(p, n=n)
    if name == "gallery":
        ts, tiles, em = kw.get("tex_size", 1024), kw.get("tiles", 8), kw.get("emitters", 32)
        p = os.path.join(directory, "gallery_%d_%d_%d.glb" % (tiles, em, ts))
        return p if os.path.exists(p) else gallery(p, tiles=tiles, emitters=em, tex_size=ts)
    if name == "instanced":
        n, k = kw.get("n", 707), kw.get("instances", 10)
        p = os.path.join(directory, "instanced_%d_x%d.glb" % (n, k))
        return p if os.path.exists(p) else instanced_heightfield(p, n=n, instances=k)
    if name == "zero_normals":
        p = os.path.join(directory, "zero_normals.glb")
        return p if os.path.exists(p) else zero_normal_cube(p)
    if name == "nomat":
        wm = bool(kw.get("with_materials", False))
        p = os.path.join(directory, "nomat_%d.glb" % int(wm))
        return p if os.path.exists(p) else default_material_scene(p, with_materials=wm)
    raise KeyError(name)
