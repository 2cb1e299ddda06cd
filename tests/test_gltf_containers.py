"""glTF front end: the same scene as a binary .glb, as a .gltf with an external .bin and external image files,
and as a .gltf with base64 data: URIs must import to the same triangles, materials, lights and texels — in the product (host-compiled
here, CUDA in the GPU run) and in the oracle's loader."""
import base64
import json
import os
import struct

import numpy as np
import pytest

import scenes


def _split_glb(path):
    blob = open(path, "rb").read()
    assert blob[:4] == b"glTF"
    off, js, bn = 12, None, b""
    while off + 8 <= len(blob):
        ln, ty = struct.unpack_from("<II", blob, off)
        body = blob[off + 8: off + 8 + ln]
        if ty == 0x4E4F534A:
            js = json.loads(body.decode())
        elif ty == 0x004E4942:
            bn = body
        off += 8 + ((ln + 3) & ~3)
    return js, bn


def _variants(glb, out_dir):
    js, bn = _split_glb(glb)
    # (a) external .bin + external images
    a = json.loads(json.dumps(js))
    a["buffers"][0]["uri"] = "scene data.bin"          # a space: the loader must percent-decode nothing and still find it
    open(os.path.join(out_dir, "scene data.bin"), "wb").write(bn)
    for i, im in enumerate(a.get("images", [])):
        if "bufferView" in im:
            v = a["bufferViews"][im.pop("bufferView")]
            o = v.get("byteOffset", 0)
            name = "tex_%d.png" % i
            open(os.path.join(out_dir, name), "wb").write(bn[o:o + v["byteLength"]])
            im["uri"] = name
            im.pop("mimeType", None)
    pa = os.path.join(out_dir, "external.gltf")
    json.dump(a, open(pa, "w"))
    # (b) data: URIs
    b = json.loads(json.dumps(js))
    b["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(bn).decode()
    pb = os.path.join(out_dir, "embedded.gltf")
    json.dump(b, open(pb, "w"))
    return pa, pb


def _snapshot(lib, path):
    with lib.load_scene(path) as s:
        tris, mat = s.triangles()
        c = s.counts()
        tex = [s.sample_texture(t, np.array([[0.25, 0.75], [0.6, 0.1]], np.float32)) for t in range(c["textures"])]
        return tris.view(np.uint32).copy(), mat.copy(), s.materials().copy(), s.lights().view(np.uint32).copy(), tex, c


@pytest.mark.parametrize("which", ["emu", "oracle"])
def test_glb_gltf_external_and_data_uri_import_identically(which, request, scene_dir, tmp_path):
    lib = request.getfixturevalue(which)
    glb = scenes.ensure(scene_dir, "pbr")
    ref = _snapshot(lib, glb)
    assert ref[5]["textures"] > 0 and ref[5]["materials"] > 1
    for variant in _variants(glb, str(tmp_path)):
        got = _snapshot(lib, variant)
        assert got[5] == ref[5]
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), variant
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]), variant
        for x, y in zip(got[4], ref[4]):
            assert np.array_equal(x, y), variant


@pytest.mark.gpu
def test_glb_gltf_external_and_data_uri_import_identically_gpu(gpu, scene_dir, tmp_path):
    glb = scenes.ensure(scene_dir, "pbr")
    ref = _snapshot(gpu, glb)
    for variant in _variants(glb, str(tmp_path)):
        got = _snapshot(gpu, variant)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
        for x, y in zip(got[4], ref[4]):
            assert np.array_equal(x, y)
