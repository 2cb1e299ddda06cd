"""First-contact GPU run: parity spot checks + raw timings (not a bench; see bench.py)."""
import os, sys, time, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
import sailor_b200
from sailor_b200.capi import Library, Params

gpu = sailor_b200.library()
d = tempfile.mkdtemp()
for name, kw, h, w in (("cube", {}, 1080, 1920), ("heightfield", {"n": 707}, 1080, 1920)):
    path = scenes.ensure(d, name, **kw)
    t = time.time(); s = gpu.load_scene(path); t_load = time.time() - t; st_load = gpu.stats()
    t = time.time(); s.build_bvh(); t_bvh = time.time() - t; st_bvh = gpu.stats()
    p = Params(height=h, width_override=w)
    for rep in range(3):
        hits = s.primary_hits(p); st = gpu.stats()
        print(name, "primary", hits.shape, "kernel s", st["secondsTraverse"], "Mrays/s", st["rays"] / st["secondsTraverse"] / 1e6, flush=True)
    print(name, s.counts(), "load s", t_load, "flatten s", st_load["secondsFlatten"], "bvh wall", t_bvh, "bvh gpu s", st_bvh["secondsBvhBuild"], "launches", st_bvh["kernelLaunches"], flush=True)
    for spp, b in ((16, 4), (64, 4)):
        pr = Params.from_samples(spp, height=h // 2, width_override=w // 2, max_bounces=b, ambient=(1, 1, 1), seed=1)
        t = time.time(); lin, srgb = s.render(pr); dt = time.time() - t; st = gpu.stats()
        print(name, "render %dx%d spp=%d b=%d" % (w // 2, h // 2, spp, b), "wall", round(dt, 3), "rays", st["rays"], "samples", st["primarySamples"],
              "trav s", round(st["secondsTraverse"], 4), "shade s", round(st["secondsShade"], 4), "iters", st["traverseLaunches"],
              "Mrays/s(wall)", round(st["rays"] / dt / 1e6, 1), "Mrays/s(trav)", round(st["rays"] / max(st["secondsTraverse"], 1e-9) / 1e6, 1), "mean", lin.mean(), flush=True)
    s.close()
