"""Small end-to-end run for compute-sanitizer: cube + pbr (PNG / JPEG / HDR textures) + heightfields: build, primary hits, a short render through the
exact kernel, the origin-local walk and the wide layout, progressive passes, the shading and generator hooks, a multi-device frame with the
replicas on this device, a 1 MiB batch budget and a forced arena overflow."""
import os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scenes, sailor_b200
from sailor_b200.capi import Params
gpu = sailor_b200.library()
d = tempfile.mkdtemp()
from sailor_b200.capi import FLAG_EXACT_TRAVERSAL, FLAG_WIDE_TRAVERSAL
for name, kw, cam in (("cube", {}, ""), ("pbr", {}, "main_cam"), ("pbr_jpeg", {}, "main_cam"), ("pbr_hdr", {}, "main_cam"), ("heightfield", {"n": 24}, ""), ("heightfield", {"n": 80}, "")):
    with gpu.load_scene(scenes.ensure(d, name, **kw)) as s:
        s.build_bvh()
        p = Params(height=48, camera=cam, num_samples=8, num_ambient_samples=8, max_bounces=3, msaa=2, ambient=(1, 1, 1), seed=1)
        h = s.primary_hits(p)
        a, _ = s.render(p)
        b, _, done = s.render_progressive(p, 1)
        assert np.array_equal(a, b) and done == 2
        o, dd = np.random.RandomState(1).uniform(-1, 1, (2000, 3)).astype(np.float32), np.random.RandomState(2).normal(size=(2000, 3)).astype(np.float32)
        s.intersect_rays(o, dd)
        for mode in ("local", "wide"):
            s.intersect_rays(o, dd, **{mode: True}); s.intersect_rays(o, dd, any_hit=True, **{mode: True})
        for flags, env in ((FLAG_EXACT_TRAVERSAL, {}), (FLAG_WIDE_TRAVERSAL, {}), (0, {"SAILOR_PT_TRAVERSAL": "local"}), (0, {"SAILOR_PT_BATCH_MB": "1"}), (0, {"SAILOR_PT_TEST_OVERFLOW": "1"}), (0, {"SAILOR_PT_MULTI_SAME_DEVICE": "1"})):
            os.environ.update(env)
            q = Params(height=48, camera=cam, num_samples=8, num_ambient_samples=8, max_bounces=3, msaa=2, ambient=(1, 1, 1), seed=1, flags=flags,
                       device_count=3 if "SAILOR_PT_MULTI_SAME_DEVICE" in env else 0)
            c, _ = s.render(q)
            for k in env:
                del os.environ[k]
            assert np.array_equal(a, c), (name, flags, env)
        n = s.counts()["triangles"]
        r = np.random.RandomState(3)
        s.shade_hits(r.randint(0, n, 500), r.dirichlet((1, 1, 1), 500)[:, 1:], r.normal(size=(500, 3)))
        print(name, kw, "ok", float(a.mean()), flush=True)
gpu.sample_generators(5, 0, 1000); gpu.sample_generators(5, 2, 1000)
gpu.trim_memory()
print("done")
