"""BASELINE configs C4 (textured PBR gallery: 96 materials, 224 textures, 32 emitters, 4 directional lights) and C5
(10 M-triangle instanced scene), against goldens minted from the oracle by tests/golden/make_golden_c4c5.py.

CPU part (kernel bodies compiled for the host): flatten, material / light import, BVH, texel fetch, reduced-size hits.
GPU part: the same through the product's C-ABI, plus the full-size cases (4K primary hits of both scenes, the 10 M
triangle BVH) and the converged-image tolerance on the gallery."""
import json
import os

import numpy as np
import pytest

import parity_checks as pc
import scenes
from golden import make_golden_c4c5 as MG
from sailor_b200.capi import Params

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G4():
    return np.load(os.path.join(HERE, "golden", "c4c5.npz"))


@pytest.fixture(scope="module")
def D4():
    return json.load(open(os.path.join(HERE, "golden", "digests_c4c5.json")))


def _check_bvh_digest(scene, d):
    nodes, mapping = scene.bvh()
    assert len(nodes) == d["nodes"]
    assert pc.sha(nodes["leftFirst"]) == d["leftFirst"] and pc.sha(nodes["triCount"]) == d["triCount"], "BVH topology differs"
    assert pc.sha(mapping) == d["mapping"], "m_triIdxMapping differs"
    assert pc.sha(nodes["aabbMin"] + np.float32(0)) == d["aabbMin"] and pc.sha(nodes["aabbMax"] + np.float32(0)) == d["aabbMax"], "node bounds differ"


def _check_hits_digest(hits, d):
    assert list(hits.shape) == d["shape"]
    assert int((hits["triId"] != pc.NOHIT).sum()) == d["nhit"]
    assert pc.sha(hits) == d["sha256"], "primary hits (t, u, v, triId) differ from the oracle's"


def _check_gallery_import(lib, path, G4, D4):
    pc.check_flatten(lib, path, G4["c4_tris"], G4["c4_mat"])
    pc.check_bvh(lib, path, G4["c4_nodes"], G4["c4_mapping"])
    with lib.load_scene(path) as s:
        c = s.counts()
        assert (c["materials"], c["textures"], c["lights"]) == (96, 224, 4)
        assert np.array_equal(s.materials(), G4["c4_materials"]), "material table differs (PathTracer.cpp:164-360)"
        assert np.array_equal(pc.bits(s.lights()), pc.bits(G4["c4_lights"])), "directional lights differ (PathTracer.cpp:362-381)"
        for t in range(c["textures"]):
            assert pc.sha(s.sample_texture(t, G4["c4_uv"])) == D4["c4_textures"]["sha256"][t], "texture %d: bilinear samples differ" % t


def test_c4_scene_import_cpu(emu, G4, D4, scene_dir):
    _check_gallery_import(emu, scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW), G4, D4)


def test_c5_instanced_scene_cpu(emu, D4, scene_dir):
    with emu.load_scene(scenes.ensure(scene_dir, "instanced", **MG.C5_SMALL_KW)) as s:
        tris, mat = s.triangles()
        assert len(tris) == D4["c5_small_tris"]["count"]
        assert pc.sha(tris) == D4["c5_small_tris"]["sha256"] and pc.sha(mat) == D4["c5_small_tris"]["mat"], "instanced flatten differs"
        _check_bvh_digest(s, D4["c5_small_bvh"])
        _check_hits_digest(s.primary_hits(Params(height=270, width_override=480)), D4["c5_small_hits"])


def test_material_and_light_tables_match_the_oracle(emu, oracle, scene_dir):
    """Every extension the importer reads (transmission, volume, ior, emissive_strength, texture_transform, alpha modes)."""
    for name, kw in (("pbr", {}), ("cube", {}), ("heightfield", {"n": 8})):
        path = scenes.ensure(scene_dir, name, **kw)
        with emu.load_scene(path) as a, oracle.load_scene(path) as b:
            assert np.array_equal(a.materials(), b.materials())
            assert np.array_equal(pc.bits(a.lights()), pc.bits(b.lights()))


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_c4_scene_import_gpu(gpu, G4, D4, scene_dir):
    _check_gallery_import(gpu, scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW), G4, D4)


@pytest.mark.gpu
def test_c4_primary_hits_4k(gpu, D4, scene_dir):
    with gpu.load_scene(scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW)) as s:
        _check_hits_digest(s.primary_hits(Params(height=2160, width_override=3840)), D4["c4_hits_4k"])


@pytest.mark.gpu
def test_c4_random_rays_match_the_oracle(gpu, oracle, scene_dir):
    pc.check_random_rays(gpu, oracle, scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW), n=200000)


@pytest.mark.gpu
def test_c4_converged_image_tolerance(gpu, G4, scene_dir):
    """Textured PBR + emitters + four lights: mean relative error against the oracle's 96-render average below 1.5 %
    (same statement as test_gpu_parity.test_converged_image_tolerance)."""
    img = pc.render_mean(gpu, scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW), Params(**MG.C4_CONVERGED), seeds=range(2000, 2256))
    err = pc.mean_rel_error(img, G4["c4_converged"])
    print("converged-image mean relative error c4: %.4f" % err)
    assert err < 0.015


@pytest.mark.gpu
def test_c4_render_is_partition_invariant_at_many_lights(gpu, scene_dir):
    """Row shards and primary-sample shards of the gallery reassemble the unsharded frame (the N-GPU split of C4)."""
    base = dict(height=54, width_override=96, num_samples=4, num_ambient_samples=4, max_bounces=3, msaa=4, ambient=(0.5, 0.5, 0.5), seed=4)
    with gpu.load_scene(scenes.ensure(scene_dir, "gallery", **MG.GALLERY_KW)) as s:
        full, _ = s.render(Params(**base))
        rows = [s.render(Params(rows=(a, b), **base))[0] for a, b in ((0, 13), (13, 40), (40, 54))]
        h = 54
        for (a, b), part in zip(((0, 13), (13, 40), (40, 54)), rows):
            assert np.array_equal(part[h - b:h - a], full[h - b:h - a])
        parts = [s.render(Params(msaa_range=(a, b), **base))[0].astype(np.float64) for a, b in ((0, 2), (2, 4))]
        assert np.allclose(sum(parts), full, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_c5_instanced_small_gpu(gpu, D4, scene_dir):
    with gpu.load_scene(scenes.ensure(scene_dir, "instanced", **MG.C5_SMALL_KW)) as s:
        tris, mat = s.triangles()
        assert pc.sha(tris) == D4["c5_small_tris"]["sha256"] and pc.sha(mat) == D4["c5_small_tris"]["mat"]
        _check_bvh_digest(s, D4["c5_small_bvh"])
        _check_hits_digest(s.primary_hits(Params(height=270, width_override=480)), D4["c5_small_hits"])


@pytest.mark.gpu
def test_c5_ten_million_triangles_bvh_and_4k_hits(gpu, D4, scene_dir):
    """C5 at full size: 9,996,980 triangles. Every node of the device-built BVH, the whole leaf order and every 4K
    primary hit equal the reference's (digests of the oracle's single-threaded build and trace)."""
    with gpu.load_scene(scenes.ensure(scene_dir, "instanced", **MG.C5_KW)) as s:
        assert s.counts()["triangles"] == D4["c5_counts"]["triangles"]
        _check_bvh_digest(s, D4["c5_bvh"])
        _check_hits_digest(s.primary_hits(Params(height=2160, width_override=3840)), D4["c5_hits_4k"])
        # a short 12-bounce render on the 10 M-triangle scene stays finite and deterministic
        p = Params(height=135, width_override=240, num_samples=4, num_ambient_samples=4, max_bounces=12, msaa=2, ambient=(1, 1, 1), seed=3)
        a, _ = s.render(p)
        b, _ = s.render(p)
        assert np.isfinite(a).all() and np.array_equal(a, b)
