#!/usr/bin/env python3
"""Mint the golden vectors from the oracle (the reference's own code, oracle/_ref/libsailor_pt_ref.so).

The reference ships no tests or known-answer vectors for this path, so the goldens are outputs of the
reference itself, generated HERE (where /root/reference exists) with:

    python oracle/build_ref.py && python tests/golden/make_golden.py

oracle build flags: g++ 13.3 -std=c++20 -O2 -mavx2 -ffp-contract=off -DNDEBUG (oracle/build_ref.py).
Small arrays are stored in full (golden.npz); large ones as SHA-256 digests of their raw bytes (digests.json).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from sailor_b200.capi import Library, Params  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lighting_inputs(n=2048, seed=3):
    r = np.random.RandomState(seed)

    def unit(k):
        v = r.normal(size=(k, 3)); return v / np.linalg.norm(v, axis=1, keepdims=True)
    rec = np.zeros((n, 24), np.float32)
    rec[:, 0:4] = r.uniform(0.02, 1, (n, 4))
    rec[:, 4] = r.uniform(0, 1, n)
    rec[:, 5] = np.where(r.uniform(size=n) < 0.2, r.uniform(0, 0.2, n), r.uniform(0.05, 1, n))
    rec[:, 6] = np.where(r.uniform(size=n) < 0.2, 1.0, r.uniform(0, 1, n))
    rec[:, 7:10] = r.uniform(0, 2, (n, 3))
    N = unit(n)
    rec[:, 10:13] = N
    for c in (13, 16):  # V, L: mostly in N's hemisphere
        v = unit(n); flip = (np.einsum("ij,ij->i", v, N) < 0) & (r.uniform(size=n) < 0.85); v[flip] *= -1
        rec[:, c:c + 3] = v
    rec[:, 19] = r.uniform(1.0, 2.0, n)
    rec[:, 20] = np.where(r.uniform(size=n) < 0.5, 0.0, r.uniform(0.05, 1, n))
    rec[:, 21] = np.where(r.uniform(size=n) < 0.5, 0.0, r.uniform(0.05, 1, n))
    rec[:, 22:24] = r.uniform(0, 0.996, (n, 2))
    return rec


def uv_grid(n=48):
    g = np.linspace(-0.75, 1.75, n, dtype=np.float32)
    return np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)


def synthetic_accumulator(w=96, h=64, seed=5):
    r = np.random.RandomState(seed)
    img = r.uniform(0, 1.2, (h, w, 3)).astype(np.float32)
    img[:8] *= 0.002            # the linear toe of the sRGB curve
    img[8:12] = 0.0
    img[12:14] = 10.0           # clamped highlights
    return img


HIT_CASES = [  # name, scene, scene kwargs, height, width override, camera
    ("cube_small", "cube", {}, 120, 0, ""),
    ("cube_square", "cube", {}, 96, 96, ""),
    ("hf64_small", "heightfield", {"n": 64}, 90, 0, ""),
    ("pbr_main", "pbr", {}, 96, 0, "main_cam"),
]
CONVERGED_SEEDS = 96   # oracle renders averaged into each converged image

BIG_HIT_CASES = [
    ("cube_c1", "cube", {}, 512, 0, ""),          # BASELINE config C1: 682x512 by the reference's aspect rule
    ("cube_c1_square", "cube", {}, 512, 512, ""),
    ("cube_c2", "cube", {}, 1080, 1920, ""),      # BASELINE config C2 resolution
    ("hf707_c3", "heightfield", {"n": 707}, 1080, 1920, ""),
]


def main():
    L = Library(os.path.join(ROOT, "oracle", "_ref", "libsailor_pt_ref.so"))
    out, dig = {}, {}
    with tempfile.TemporaryDirectory() as d:
        for name, kw in (("cube", {}), ("pbr", {}), ("heightfield", {"n": 64})):
            key = name if name != "heightfield" else "hf64"
            with L.load_scene(scenes.ensure(d, name, **kw)) as s:
                tris, mat = s.triangles()
                nodes, mapping = s.bvh()
                out[key + "_tris"] = tris; out[key + "_mat"] = mat
                out[key + "_nodes"] = nodes; out[key + "_mapping"] = mapping
        for name, scene, kw, h, wo, cam in HIT_CASES:
            with L.load_scene(scenes.ensure(d, scene, **kw)) as s:
                p = Params(height=h, width_override=wo, camera=cam)
                out[name + "_cam"] = s.camera(p)[2]
                out[name + "_hits"] = s.primary_hits(p)
        for name, scene, kw, h, wo, cam in BIG_HIT_CASES:
            with L.load_scene(scenes.ensure(d, scene, **kw)) as s:
                p = Params(height=h, width_override=wo, camera=cam)
                hits = s.primary_hits(p)
                dig[name + "_hits"] = {"sha256": sha(hits), "shape": list(hits.shape), "nhit": int((hits["triId"] != 0xFFFFFFFF).sum())}
                if scene == "heightfield":
                    nodes, mapping = s.bvh()
                    dig[name + "_bvh"] = {"nodes": int(len(nodes)), "leftFirst": sha(nodes["leftFirst"]), "triCount": sha(nodes["triCount"]),
                                          "mapping": sha(mapping),
                                          # + 0.0 folds -0.0 into +0.0: min/max chains may legitimately differ in the sign of a zero
                                          "aabbMin": sha(nodes["aabbMin"] + np.float32(0)), "aabbMax": sha(nodes["aabbMax"] + np.float32(0))}
        with L.load_scene(scenes.ensure(d, "pbr")) as s:
            uv = uv_grid()
            out["uv_grid"] = uv
            for t in range(s.counts()["textures"]):
                out["pbr_tex%d" % t] = s.sample_texture(t, uv)
            # converged images for the tolerance test: the reference's own high-spp render (S = A = 64 at the first hit, msaa 8)
            p = Params(height=24, camera="main_cam", num_samples=64, num_ambient_samples=64, max_bounces=4, msaa=8, ambient=(1.0, 1.0, 1.0), seed=11)
            acc = None
            for seed in range(CONVERGED_SEEDS):
                p.seed = 100 + seed
                lin, _ = s.render(p, want_srgb=False)
                acc = lin.astype(np.float64) if acc is None else acc + lin
            out["pbr_converged"] = (acc / CONVERGED_SEEDS).astype(np.float32)
        with L.load_scene(scenes.ensure(d, "heightfield", n=64)) as s:
            p = Params(height=18, num_samples=32, num_ambient_samples=32, max_bounces=3, msaa=8, ambient=(0.6, 0.7, 0.9), seed=1)
            acc = None
            for seed in range(CONVERGED_SEEDS):
                p.seed = 200 + seed
                lin, _ = s.render(p, want_srgb=False)
                acc = lin.astype(np.float64) if acc is None else acc + lin
            out["hf64_converged"] = (acc / CONVERGED_SEEDS).astype(np.float32)
    rec = lighting_inputs()
    out["lighting_in"] = rec
    out["lighting_out"] = L.eval_lighting(rec)
    acc = synthetic_accumulator()
    out["output_in"] = acc
    out["output_srgb"] = L.output_stage(acc)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    json.dump(dig, open(os.path.join(HERE, "digests.json"), "w"), indent=1, sort_keys=True)
    print("wrote golden.npz (%d arrays, %.1f KiB) and digests.json" % (len(out), os.path.getsize(os.path.join(HERE, "golden.npz")) / 1024))


if __name__ == "__main__":
    main()
