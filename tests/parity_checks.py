"""Parity checks shared by the CPU-side tests (kernel bodies compiled for the host, tests/emu) and the GPU tests
(the product library through its C-ABI).  `lib` is the library under test, `oracle` the reference build, `G` the
golden vectors minted from the oracle by tests/golden/make_golden.py.

Bars: bit-exact for flattened triangles, BVH topology / leaf order, hit triangle ids, hit distance and barycentrics,
texel fetch + bilinear blend, and the 8-bit output (<= 1 LSB on <= 1e-4 of the bytes, because of powf); relative
1e-4 for the BSDF function table (libm transcendentals differ in the last ulps); a stated mean relative error for
converged images.
"""
import hashlib

import numpy as np

from sailor_b200.capi import Params

NOHIT = 0xFFFFFFFF


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_hits_equal(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a["triId"], b["triId"]), "hit triangle ids differ at %d rays" % int((a["triId"] != b["triId"]).sum())
    for k in ("t", "baryU", "baryV"):
        assert np.array_equal(bits(a[k]), bits(b[k])), "%s differs (bitwise) at %d rays" % (k, int((bits(a[k]) != bits(b[k])).sum()))


def check_flatten(lib, path, tris_ref, mat_ref):
    with lib.load_scene(path) as s:
        tris, mat = s.triangles()
    assert tris.shape == tris_ref.shape
    assert np.array_equal(bits(tris), bits(tris_ref)), "flattened triangles differ"
    assert np.array_equal(mat, mat_ref)


def check_bvh(lib, path, nodes_ref, mapping_ref):
    with lib.load_scene(path) as s:
        nodes, mapping = s.bvh()
    assert len(nodes) == len(nodes_ref), "node count %d vs %d" % (len(nodes), len(nodes_ref))
    assert np.array_equal(nodes["leftFirst"], nodes_ref["leftFirst"])
    assert np.array_equal(nodes["triCount"], nodes_ref["triCount"])
    assert np.array_equal(mapping, mapping_ref), "m_triIdxMapping differs"
    assert np.array_equal(nodes["aabbMin"], nodes_ref["aabbMin"]) and np.array_equal(nodes["aabbMax"], nodes_ref["aabbMax"])


def check_primary_hits(lib, path, height, width_override, camera, cam_ref, hits_ref):
    with lib.load_scene(path) as s:
        p = Params(height=height, width_override=width_override, camera=camera)
        w, h, cam = s.camera(p)
        hits = s.primary_hits(p)
    assert np.array_equal(bits(cam), bits(cam_ref)), "camera vectors differ"
    assert_hits_equal(hits, hits_ref)


def random_rays(n, seed, scale=1.5):
    r = np.random.RandomState(seed)
    o = r.uniform(-scale, scale, (n, 3)).astype(np.float32)
    d = r.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # edge cases the reference's slab test is sensitive to: axis-aligned directions (1/0 = inf, 0*inf = NaN), rays
    # starting exactly on box planes / vertices, un-normalised and zero directions
    k = n // 8
    d[:k] = 0; d[np.arange(k), r.randint(0, 3, k)] = r.choice([-1.0, 1.0], k)
    o[:k // 2] = np.round(o[:k // 2] * 2) / 2
    d[k:k + 8] *= 1e-3
    d[k + 8:k + 12] = 0.0
    d[k + 12:k + 16, 1] = -0.0
    return o, d


def check_random_rays(lib, oracle, path, n=20000, seed=1):
    o, d = random_rays(n, seed)
    with lib.load_scene(path) as a, oracle.load_scene(path) as b:
        nt = a.counts()["triangles"]
        ignore = np.random.RandomState(seed + 1).randint(0, nt, n).astype(np.uint32)
        assert_hits_equal(a.intersect_rays(o, d), b.intersect_rays(o, d))
        assert_hits_equal(a.intersect_rays(o, d, ignore), b.intersect_rays(o, d, ignore))


def check_textures(lib, path, uv, expected):
    with lib.load_scene(path) as s:
        for t, exp in enumerate(expected):
            got = s.sample_texture(t, uv)
            assert np.array_equal(bits(got), bits(exp)), "texture %d: %d samples differ" % (t, int((bits(got) != bits(exp)).any(axis=1).sum()))


def check_output_stage(lib, acc, expected):
    got = lib.output_stage(acc)
    diff = np.abs(got.astype(np.int32) - expected.astype(np.int32))
    assert diff.max() <= 1, "output stage differs by more than 1 LSB"
    assert (diff != 0).mean() <= 1e-4, "output stage: %.2e of the bytes differ" % (diff != 0).mean()


def check_lighting(lib, rec, expected, rtol=2e-4):
    got = lib.eval_lighting(rec)
    both_nan = np.isnan(got) & np.isnan(expected)
    scale = np.maximum(np.abs(expected), 1e-3)
    err = np.where(both_nan, 0.0, np.abs(got - expected) / scale)
    assert not np.isnan(err).any(), "NaN pattern differs"
    assert err.max() <= rtol, "BSDF table max relative error %.3g (column %d)" % (err.max(), int(np.argmax(err.max(axis=0))))


def mean_rel_error(img, ref):
    """Mean |a-b| over pixels / mean(b): the converged-image metric north_star names."""
    return float(np.abs(img.astype(np.float64) - ref).mean() / np.abs(ref).mean())


def render_mean(lib, path, params, seeds):
    acc = None
    with lib.load_scene(path) as s:
        for seed in seeds:
            params.seed = seed
            lin, _ = s.render(params, want_srgb=False)
            acc = lin.astype(np.float64) if acc is None else acc + lin
    return acc / len(seeds)
