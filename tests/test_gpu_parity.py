"""GPU parity tests: the product library (CUDA, sm_100a) through its C-ABI against the oracle and the goldens.
Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest

import parity_checks as pc
import scenes
from golden.make_golden import BIG_HIT_CASES, HIT_CASES
from sailor_b200.capi import Params

pytestmark = pytest.mark.gpu


def _scene(scene_dir, name, kw):
    return scenes.ensure(scene_dir, name, **kw)


def test_backend_is_cuda(gpu):
    assert gpu.backend() == "cuda sm_100a"


@pytest.mark.parametrize("key,name,kw", [("cube", "cube", {}), ("pbr", "pbr", {}), ("hf64", "heightfield", {"n": 64})])
def test_flatten_and_bvh_match_goldens(gpu, G, scene_dir, key, name, kw):
    pc.check_flatten(gpu, _scene(scene_dir, name, kw), G[key + "_tris"], G[key + "_mat"])
    pc.check_bvh(gpu, _scene(scene_dir, name, kw), G[key + "_nodes"], G[key + "_mapping"])


@pytest.mark.parametrize("case", HIT_CASES, ids=[c[0] for c in HIT_CASES])
def test_primary_hits_match_goldens(gpu, G, scene_dir, case):
    name, scene, kw, h, wo, cam = case
    pc.check_primary_hits(gpu, _scene(scene_dir, scene, kw), h, wo, cam, G[name + "_cam"], G[name + "_hits"])


@pytest.mark.parametrize("case", BIG_HIT_CASES, ids=[c[0] for c in BIG_HIT_CASES])
def test_full_size_primary_hits_match_digests(gpu, digests, scene_dir, case):
    """BASELINE configs C1, C2, C3 at full size: SHA-256 of (t, u, v, triId) for every pixel equals the oracle's."""
    name, scene, kw, h, wo, cam = case
    with gpu.load_scene(_scene(scene_dir, scene, kw)) as s:
        hits = s.primary_hits(Params(height=h, width_override=wo, camera=cam))
        assert list(hits.shape) == digests[name + "_hits"]["shape"]
        assert int((hits["triId"] != pc.NOHIT).sum()) == digests[name + "_hits"]["nhit"]
        assert pc.sha(hits) == digests[name + "_hits"]["sha256"]
        if name + "_bvh" in digests:
            nodes, mapping = s.bvh()
            d = digests[name + "_bvh"]
            assert len(nodes) == d["nodes"]
            assert pc.sha(nodes["leftFirst"]) == d["leftFirst"] and pc.sha(nodes["triCount"]) == d["triCount"]
            assert pc.sha(mapping) == d["mapping"]
            assert pc.sha(nodes["aabbMin"] + np.float32(0)) == d["aabbMin"] and pc.sha(nodes["aabbMax"] + np.float32(0)) == d["aabbMax"]


@pytest.mark.parametrize("name,kw,n", [("cube", {}, 200000), ("pbr", {}, 200000), ("heightfield", {"n": 64}, 200000), ("heightfield", {"n": 256}, 400000)])
def test_random_and_degenerate_rays_match_the_oracle(gpu, oracle, scene_dir, name, kw, n):
    pc.check_random_rays(gpu, oracle, _scene(scene_dir, name, kw), n=n)


@pytest.mark.parametrize("with_materials", [False, True])
def test_default_material_and_malformed_attribute_streams(gpu, oracle, scene_dir, with_materials):
    pc.check_default_material(gpu, oracle, scenes.ensure(scene_dir, "nomat", with_materials=with_materials), 1 if with_materials else 0)


@pytest.mark.parametrize("name,kw", [("cube", {}), ("pbr", {})])
def test_fast_walks_on_small_scenes(gpu, oracle, scene_dir, name, kw, monkeypatch):
    """Scenes the shared-memory kernel normally takes, forced through the origin-local walk and the wide layout (+ exact replay)."""
    monkeypatch.setenv("SAILOR_PT_FORCE_WIDE", "1")
    monkeypatch.setenv("SAILOR_PT_FORCE_LOCAL", "1")
    pc.check_random_rays(gpu, oracle, _scene(scene_dir, name, kw), n=50000)


def test_bvh_of_ragged_scenes(gpu, oracle, tmp_path):
    for tag, count, dup in (("one", 1, False), ("four", 4, False), ("five", 5, False), ("dups", 40, True), ("many", 3000, False)):
        g = scenes.GlbBuilder()
        r = np.random.RandomState(count)
        pos = r.uniform(-1, 1, (count * 3, 3)).astype(np.float32)
        if dup:
            pos = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (count, 1))
            pos[: count * 3 // 2, 2] = 0.5
        mat = g.material(pbrMetallicRoughness={"metallicFactor": 0.0})
        g.node(mesh=g.mesh(pos, None, mat))
        path = g.write(str(tmp_path / (tag + ".glb")))
        with oracle.load_scene(path) as o:
            nodes, mapping = o.bvh()
        pc.check_bvh(gpu, path, nodes, mapping)
        pc.check_random_rays(gpu, oracle, path, n=20000, seed=count)


def test_bvh_1m_triangles_matches_the_oracle(gpu, oracle, scene_dir):
    """C3-sized build: every node and the whole leaf order equal the reference's single-threaded build."""
    path = _scene(scene_dir, "heightfield", {"n": 500})
    with oracle.load_scene(path) as o:
        nodes, mapping = o.bvh()
    pc.check_bvh(gpu, path, nodes, mapping)


def test_texture_sampler_output_stage_and_bsdf_table(gpu, G, scene_dir):
    pc.check_textures(gpu, _scene(scene_dir, "pbr", {}), G["uv_grid"], [G["pbr_tex%d" % t] for t in range(4)])
    pc.check_output_stage(gpu, G["output_in"], G["output_srgb"])
    pc.check_lighting(gpu, G["lighting_in"], G["lighting_out"], rtol=2e-4)


def test_output_stage_full_size_round_trip(gpu, oracle):
    r = np.random.RandomState(0)
    acc = r.uniform(0, 1.1, (1080, 1920, 3)).astype(np.float32)
    pc.check_output_stage(gpu, acc, oracle.output_stage(acc))
    # idempotence-style property: a constant image is a fixed point of the aberration taps
    flat = np.full((270, 480, 3), 0.25, np.float32)
    out = gpu.output_stage(flat)
    assert (out == out[0, 0]).all()


def test_render_is_deterministic_and_partition_invariant(gpu, scene_dir):
    path = _scene(scene_dir, "pbr", {})
    base = dict(height=60, camera="main_cam", num_samples=2, num_ambient_samples=2, max_bounces=3, msaa=4, ambient=(1, 1, 1), seed=9)
    with gpu.load_scene(path) as s:
        full, _ = s.render(Params(**base))
        again, _ = s.render(Params(**base))
        assert np.array_equal(full, again)
        w, h, _ = s.camera(Params(**base))
        top, _ = s.render(Params(rows=(0, 21), **base))
        bottom, _ = s.render(Params(rows=(21, h), **base))
        assert np.array_equal(top[h - 21:], full[h - 21:]) and np.array_equal(bottom[:h - 21], full[:h - 21])
        parts = [s.render(Params(msaa_range=(a, b), **base))[0].astype(np.float64) for a, b in ((0, 1), (1, 3), (3, 4))]
        assert np.allclose(sum(parts), full, rtol=1e-6, atol=1e-7)


def test_render_into_pinned_host_buffers_is_identical(gpu, scene_dir):
    """SailorPt_PinHostBuffer only changes how the result travels (direct DMA instead of staged copies)."""
    p = Params(height=540, width_override=960, num_samples=2, num_ambient_samples=2, max_bounces=2, msaa=1, ambient=(1, 1, 1), seed=4)
    with gpu.load_scene(_scene(scene_dir, "cube", {})) as s:
        lin0, srgb0 = s.render(p)
        lin = np.zeros_like(lin0); srgb = np.zeros_like(srgb0)
        gpu.pin_host_buffer(lin); gpu.pin_host_buffer(srgb)
        try:
            s.render(p, out=(lin, srgb))
        finally:
            gpu.unpin_host_buffer(lin); gpu.unpin_host_buffer(srgb)
        assert lin0.nbytes > (1 << 20) and np.array_equal(lin, lin0) and np.array_equal(srgb, srgb0)
    with pytest.raises(Exception):
        gpu.unpin_host_buffer(lin)                      # not pinned any more


def test_gpu_render_equals_host_compiled_kernel_bodies(gpu, emu, scene_dir):
    """Same RNG streams, same state machine: the CUDA render differs from the host-compiled bodies only through
    libm transcendentals, far below Monte-Carlo noise."""
    p = Params(height=30, camera="main_cam", num_samples=4, num_ambient_samples=4, max_bounces=3, msaa=2, ambient=(1, 1, 1), seed=2)
    with gpu.load_scene(_scene(scene_dir, "pbr", {})) as a, emu.load_scene(_scene(scene_dir, "pbr", {})) as b:
        x, _ = a.render(p)
        y, _ = b.render(p)
    assert pc.mean_rel_error(x, y) < 2e-3


@pytest.mark.parametrize("key,name,kw,params", [
    ("pbr_converged", "pbr", {}, dict(height=24, camera="main_cam", num_samples=64, num_ambient_samples=64, max_bounces=4, msaa=8, ambient=(1.0, 1.0, 1.0))),
    ("hf64_converged", "heightfield", {"n": 64}, dict(height=18, num_samples=32, num_ambient_samples=32, max_bounces=3, msaa=8, ambient=(0.6, 0.7, 0.9))),
])
def test_converged_image_tolerance(gpu, G, scene_dir, key, name, kw, params):
    """north_star: converged images agree with the reference's own high-spp render within a stated mean relative
    error.  Stated tolerance: 1.5 % mean relative error (mean |a-b| / mean b over all pixels and channels) with 256
    GPU renders averaged against 96 oracle renders averaged; what remains is Monte-Carlo noise of the two averages
    (the oracle against its own average converges as 7.4 % / sqrt(renders))."""
    img = pc.render_mean(gpu, _scene(scene_dir, name, kw), Params(**params), seeds=range(1000, 1256))
    err = pc.mean_rel_error(img, G[key])
    print("converged-image mean relative error %s: %.4f" % (key, err))
    assert err < 0.015


def test_run_writes_a_png(gpu, scene_dir, tmp_path):
    import sailor_b200
    out = str(tmp_path / "cube.png")
    p = Params()
    sailor_b200.PathTracer.ParseCommandLineArgs(p, ["exe", "--in", _scene(scene_dir, "cube", {}), "--out", out, "--height", "64",
                                                    "--samples", "4", "--bounces", "2", "--ambient", "ffffff"])
    p.m_numAmbientSamples = p.m_numSamples
    assert sailor_b200.PathTracer().Run(p) == 0
    data = open(out, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and len(data) > 100


def test_progressive_render_checkpoint_resume_and_linear_image_files(gpu, scene_dir, tmp_path):
    pc.check_progressive_and_image_io(gpu, _scene(scene_dir, "pbr", {}), tmp_path)


def test_run_writes_linear_dumps_by_extension(gpu, scene_dir, tmp_path):
    import sailor_b200
    out = str(tmp_path / "cube.pfm")
    p = Params()
    sailor_b200.PathTracer.ParseCommandLineArgs(p, ["exe", "--in", _scene(scene_dir, "cube", {}), "--out", out, "--height", "48",
                                                    "--samples", "4", "--bounces", "2", "--ambient", "ffffff"])
    p.m_numAmbientSamples = p.m_numSamples
    assert sailor_b200.PathTracer().Run(p) == 0
    img = pc.read_pfm(out)
    assert img.shape[0] == 48 and np.isfinite(img).all() and img.max() > 0


def test_cpp_host_renders_like_the_library(gpu, scene_dir, tmp_path):
    """The C++ host (csrc/cli_main.cpp) over the C-ABI: one-shot PNG, and a progressive render whose .pfm has the bits of SailorPt_Render."""
    import subprocess
    from sailor_b200 import build as product_build
    exe = product_build.build_cli()
    scene = _scene(scene_dir, "cube", {})
    png, pfm, ck = str(tmp_path / "a.png"), str(tmp_path / "a.pfm"), str(tmp_path / "a.ckpt")
    common = ["--in", scene, "--height", "64", "--samples", "16", "--bounces", "2", "--ambient", "ffffff", "--seed", "5"]
    r = subprocess.run([exe] + common + ["--out", png], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(png, "rb").read(8) == b"\x89PNG\r\n\x1a\n"
    r = subprocess.run([exe] + common + ["--out", pfm, "--passes", "1", "--checkpoint", ck], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p = Params()
    gpu.parse_command_line_args(p, ["exe"] + common)
    p.m_numAmbientSamples = p.m_numSamples
    p.seed = 5
    with gpu.load_scene(scene) as s:
        lin, _ = s.render(p)
    assert np.array_equal(pc.bits(pc.read_pfm(pfm)), pc.bits(lin))
    # SailorPt_Run with deviceCount > 1 (--devices): the same bits from however many devices the box has
    pfm2 = str(tmp_path / "b.pfm")
    r = subprocess.run([exe] + common + ["--out", pfm2, "--devices", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "device(s)" in r.stderr
    assert np.array_equal(pc.bits(pc.read_pfm(pfm2)), pc.bits(lin))


def test_trim_memory_releases_the_working_set_and_rendering_goes_on(gpu, scene_dir):
    import torch
    p = Params(height=120, num_samples=4, num_ambient_samples=4, max_bounces=2, msaa=2, ambient=(1, 1, 1), seed=6)
    with gpu.load_scene(_scene(scene_dir, "cube", {})) as s:
        a, _ = s.render(p)
        free_before = torch.cuda.mem_get_info()[0]
        gpu.trim_memory()
        assert torch.cuda.mem_get_info()[0] >= free_before
        b, _ = s.render(p)                       # the scene is still valid; the arenas come back
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name,kw", [("pbr", {}), ("pbr_jpeg", {}), ("pbr_hdr", {}), ("gallery", {"tex_size": 64, "tiles": 4, "emitters": 8})])
def test_shading_context_is_bit_identical_to_get_material_data(gpu, oracle, scene_dir, name, kw):
    """a13 deterministically: SailorPt_ShadeHits (the function ExpandKernel calls) against the reference's own GetMaterialData."""
    pc.check_shade_hits(gpu, oracle, _scene(scene_dir, name, kw), n=20000)


def test_sample_generators_have_the_reference_distributions(gpu):
    pc.check_sample_generators(gpu)


@pytest.mark.parametrize("key,frames", [("c1_512", 512), ("c2_480x270", 64)])
def test_converged_image_tolerance_at_baseline_config_sizes(gpu, scene_dir, key, frames):
    """BASELINE configs[0] at FULL size (512x512, 16 spp, 4 bounces) and configs[1] at 480x270 (256 spp, 8 bounces): the mean of `frames`
    GPU frames against the mean of as many reference frames (tests/golden/make_golden_converged.py).  Stated tolerance: mean relative
    error < 1.5 % over all pixels and channels, and < 2 % for every channel on its own and for every quadrant of the image (a per-pixel
    bias that averages out over the frame would show there)."""
    import os
    from golden import make_golden_converged as M
    ref = np.load(os.path.join(os.path.dirname(M.__file__), "converged.npz"))[key].astype(np.float64)
    kw = M.C1 if key == "c1_512" else M.C2
    img = pc.render_mean(gpu, _scene(scene_dir, "cube", {}), Params(**kw), seeds=range(5000, 5000 + frames))
    assert img.shape == ref.shape
    err = pc.mean_rel_error(img, ref)
    per_channel = [pc.mean_rel_error(img[..., c], ref[..., c]) for c in range(3) if ref[..., c].mean() > 1e-3]
    h, w = ref.shape[:2]
    quads = [pc.mean_rel_error(img[y:y + h // 2, x:x + w // 2], ref[y:y + h // 2, x:x + w // 2]) for y in (0, h // 2) for x in (0, w // 2)]
    print("converged-image mean relative error %s: %.4f per channel %s per quadrant %s" % (key, err, ["%.4f" % e for e in per_channel], ["%.4f" % e for e in quads]))
    assert err < 0.015 and max(per_channel) < 0.02 and max(quads) < 0.02
    # the means themselves (bias, not noise): within 0.3 %
    assert abs(img.mean() - ref.mean()) / ref.mean() < 0.003


def test_rejection_loop_cap_keeps_the_frame_finite(gpu, scene_dir):
    """A surface whose shading normal is NaN (zero NORMAL accessor) never yields a valid BSDF sample: the reference's rejection loop
    (PathTracer.cpp:761-767) would spin forever, the product gives up after 4096 tries.  The frame finishes, is finite
    and deterministic, and the rest of the scene (the floor) is lit as usual."""
    p = Params(height=96, num_samples=4, num_ambient_samples=4, max_bounces=3, msaa=4, ambient=(1.0, 1.0, 1.0), seed=2)
    with gpu.load_scene(_scene(scene_dir, "zero_normals", {})) as s:
        a, _ = s.render(p)
        b, _ = s.render(p)
    assert np.isfinite(a).all() and np.array_equal(a, b)
    assert 0.0 <= a.min() and a.max() <= 64.0 and a[-8:].mean() > 0.05          # bottom rows: the floor in front of the cube


@pytest.mark.parametrize("name", ["pbr_jpeg", "pbr_hdr"])
def test_jpeg_and_hdr_textured_scenes_fetch_the_reference_texels(gpu, oracle, scene_dir, name):
    """ADVICE r1 / VERDICT r1 item 7: a glTF whose textures are JPEG files (baseline 4:2:0, progressive 4:4:4, baseline 4:2:2) imports, and every
    texture samples bit-identically to the reference's stb_image + CombinedSampler2D path."""
    path = _scene(scene_dir, name, {})          # pbr_hdr: two Radiance .hdr images, kept as float texels like the reference (MaterialUtils.h:224-229)
    uv = np.random.RandomState(7).uniform(-1.0, 2.0, (20000, 2)).astype(np.float32)
    with gpu.load_scene(path) as a, oracle.load_scene(path) as b:
        assert a.counts() == b.counts() and a.counts()["textures"] == 4
        for t in range(4):
            assert np.array_equal(pc.bits(a.sample_texture(t, uv)), pc.bits(b.sample_texture(t, uv)))


def test_frame_does_not_depend_on_batching_or_on_an_overflow_retry(gpu, scene_dir, monkeypatch):
    """VERDICT r1 (robustness): the overflow-and-retry planner, driven on the GPU."""
    pc.check_batching_and_overflow_retry(gpu, _scene(scene_dir, "pbr", {}), monkeypatch, camera="main_cam")
    pc.check_batching_and_overflow_retry(gpu, _scene(scene_dir, "heightfield", {"n": 64}), monkeypatch, height=90)
