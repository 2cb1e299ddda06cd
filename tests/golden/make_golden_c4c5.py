#!/usr/bin/env python3
"""Golden vectors for BASELINE configs C4 (textured PBR gallery, many materials / emitters / lights, 4K) and
C5 (10 M-triangle instanced scene, 4K), minted from the oracle (the reference's own code) HERE:

    python oracle/build_ref.py && python tests/golden/make_golden_c4c5.py

Small arrays are stored in full (c4c5.npz); large ones as SHA-256 digests of their raw bytes (digests_c4c5.json).
The C4 goldens use 64x64 textures (the bench workload uses 1024x1024 of the same generator; texel decode and the
bilinear fetch do not depend on the size); the C5 goldens are at full size (9,996,980 triangles, 3840x2160).
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
from golden.make_golden import uv_grid  # noqa: E402
from sailor_b200.capi import Library, Params  # noqa: E402

GALLERY_KW = {"tex_size": 64}
C4_CONVERGED = dict(height=27, width_override=48, num_samples=32, num_ambient_samples=32, max_bounces=4, msaa=8, ambient=(0.5, 0.55, 0.6))
C4_CONVERGED_SEEDS = 96
C5_KW = {"n": 707, "instances": 10}
C5_SMALL_KW = {"n": 48, "instances": 10}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bvh_digest(nodes, mapping):
    # + 0.0 folds -0.0 into +0.0: min/max chains may legitimately differ in the sign of a zero
    return {"nodes": int(len(nodes)), "leftFirst": sha(nodes["leftFirst"]), "triCount": sha(nodes["triCount"]), "mapping": sha(mapping),
            "aabbMin": sha(nodes["aabbMin"] + np.float32(0)), "aabbMax": sha(nodes["aabbMax"] + np.float32(0))}


def hits_digest(hits):
    return {"sha256": sha(hits), "shape": list(hits.shape), "nhit": int((hits["triId"] != 0xFFFFFFFF).sum())}


def main():
    L = Library(os.path.join(ROOT, "oracle", "_ref", "libsailor_pt_ref.so"))
    out, dig = {}, {}
    with tempfile.TemporaryDirectory() as d:
        with L.load_scene(scenes.ensure(d, "gallery", **GALLERY_KW)) as s:
            tris, mat = s.triangles()
            nodes, mapping = s.bvh()
            out["c4_tris"] = tris; out["c4_mat"] = mat; out["c4_nodes"] = nodes; out["c4_mapping"] = mapping
            out["c4_materials"] = s.materials(); out["c4_lights"] = s.lights()
            uv = uv_grid(24)
            out["c4_uv"] = uv
            dig["c4_textures"] = {"count": s.counts()["textures"], "sha256": [sha(s.sample_texture(t, uv)) for t in range(s.counts()["textures"])]}
            dig["c4_hits_4k"] = hits_digest(s.primary_hits(Params(height=2160, width_override=3840)))
            p = Params(**C4_CONVERGED)
            acc = None
            for seed in range(C4_CONVERGED_SEEDS):
                p.seed = 300 + seed
                lin, _ = s.render(p, want_srgb=False)
                acc = lin.astype(np.float64) if acc is None else acc + lin
            out["c4_converged"] = (acc / C4_CONVERGED_SEEDS).astype(np.float32)
        with L.load_scene(scenes.ensure(d, "instanced", **C5_SMALL_KW)) as s:
            tris, mat = s.triangles()
            nodes, mapping = s.bvh()
            dig["c5_small_tris"] = {"sha256": sha(tris), "mat": sha(mat), "count": int(len(tris))}
            dig["c5_small_bvh"] = bvh_digest(nodes, mapping)
            dig["c5_small_hits"] = hits_digest(s.primary_hits(Params(height=270, width_override=480)))
        with L.load_scene(scenes.ensure(d, "instanced", **C5_KW)) as s:
            nodes, mapping = s.bvh()
            dig["c5_bvh"] = bvh_digest(nodes, mapping)
            dig["c5_hits_4k"] = hits_digest(s.primary_hits(Params(height=2160, width_override=3840)))
            dig["c5_counts"] = s.counts()
    np.savez_compressed(os.path.join(HERE, "c4c5.npz"), **out)
    json.dump(dig, open(os.path.join(HERE, "digests_c4c5.json"), "w"), indent=1, sort_keys=True)
    print("wrote c4c5.npz (%d arrays), digests_c4c5.json" % len(out))


if __name__ == "__main__":
    main()
