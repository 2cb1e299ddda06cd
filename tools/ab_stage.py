"""A/B two builds of the library on the same box: python tools/ab_stage.py libA.so libB.so [workload] — mean stage times over alternating frames."""
import os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes, bench
from sailor_b200.capi import Library
libs = [Library(os.path.abspath(p)) for p in sys.argv[1:3]]
name = sys.argv[3] if len(sys.argv) > 3 else "c2"
w = bench.WORKLOADS[name]
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, seed=1)
sc = [L.load_scene(path) for L in libs]
keys = ("secondsCall", "secondsTraverse", "secondsExpand", "secondsFanOut", "secondsClassify", "secondsGather")
acc = [[] for _ in libs]
# two copies of the library = two sets of shared arenas (40 GiB each on a 180 GB device: both get the full budget)
for rep in range(3 + 12):
    for i, (L, s) in enumerate(zip(libs, sc)):
        s.render_resident(p, rebuild_bvh=True, output_stage=True)
        st = L.stats()
        if rep >= 3: acc[i].append([st[k] * 1e3 for k in keys])
for i, a in enumerate(acc):
    a = np.array(a)
    print(os.path.basename(sys.argv[1 + i]), " ".join("%s %.3f" % (k[7:], v) for k, v in zip(keys, np.median(a, axis=0))))
