#!/usr/bin/env python3
"""Mints tests/golden/converged.npz: the reference's own converged renders at BASELINE config sizes (the converged-image
tests were postage stamps).  Needs the compiled reference (oracle/_ref, built from /root/reference by oracle/build_ref.py).

    c1_512      BASELINE configs[0] at FULL size: the cube, 512x512, --samples 16 (msaa 4 x S 4, A 4), 4 bounces, ambient ffffff;
                mean of C1_FRAMES oracle frames (unseeded rand(): every frame is another sample of the same estimator)
    c2_480x270  BASELINE configs[1] at 1/4 linear size: the cube, 480x270, --samples 256 (msaa 8 x S 32, A 32), 8 bounces; mean of C2_FRAMES frames
Stored as float16 (relative precision 1e-3, far below the 1.5 % tolerance of the tests) to keep the fixture small."""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from sailor_b200.capi import Library, Params  # noqa: E402

C1_FRAMES, C2_FRAMES = 512, 64
C1 = dict(height=512, width_override=512, num_samples=4, num_ambient_samples=4, max_bounces=4, msaa=4, ambient=(1.0, 1.0, 1.0))
C2 = dict(height=270, width_override=480, num_samples=32, num_ambient_samples=32, max_bounces=8, msaa=8, ambient=(1.0, 1.0, 1.0))


def mean_of(lib, path, kw, frames):
    acc = None
    with lib.load_scene(path) as s:
        s.build_bvh()
        for k in range(frames):
            lin, _ = s.render(Params(seed=k, **kw), want_srgb=False)
            acc = lin.astype(np.float64) if acc is None else acc + lin
    return (acc / frames)


if __name__ == "__main__":
    oracle = Library(os.path.join(ROOT, "oracle", "_ref", "libsailor_pt_ref.so"))
    cube = scenes.ensure(tempfile.mkdtemp(), "cube")
    out = {}
    for key, kw, frames in (("c1_512", C1, C1_FRAMES), ("c2_480x270", C2, C2_FRAMES)):
        t0 = time.time()
        out[key] = mean_of(oracle, cube, kw, frames).astype(np.float16)
        print(key, out[key].shape, "%.1f s" % (time.time() - t0), "mean %.4f" % float(out[key].astype(np.float64).mean()))
    np.savez_compressed(os.path.join(HERE, "converged.npz"), **out)
