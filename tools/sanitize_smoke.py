"""Small end-to-end run for compute-sanitizer: cube + pbr + heightfield(24), build, primary hits, a short render, progressive passes."""
import os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scenes, sailor_b200
from sailor_b200.capi import Params
gpu = sailor_b200.library()
d = tempfile.mkdtemp()
for name, kw, cam in (("cube", {}, ""), ("pbr", {}, "main_cam"), ("heightfield", {"n": 24}, ""), ("heightfield", {"n": 80}, "")):
    with gpu.load_scene(scenes.ensure(d, name, **kw)) as s:
        s.build_bvh()
        p = Params(height=48, camera=cam, num_samples=8, num_ambient_samples=8, max_bounces=3, msaa=2, ambient=(1, 1, 1), seed=1)
        h = s.primary_hits(p)
        a, _ = s.render(p)
        b, _, done = s.render_progressive(p, 1)
        assert np.array_equal(a, b) and done == 2
        o, dd = np.random.RandomState(1).uniform(-1, 1, (2000, 3)).astype(np.float32), np.random.RandomState(2).normal(size=(2000, 3)).astype(np.float32)
        s.intersect_rays(o, dd)
        print(name, kw, "ok", float(a.mean()), flush=True)
gpu.trim_memory()
print("done")
