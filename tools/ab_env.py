"""A/B one library with an environment switch read per frame: python tools/ab_env.py VAR [workload] — alternating frames with VAR unset / set."""
import os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes, bench, sailor_b200
var = sys.argv[1]; name = sys.argv[2] if len(sys.argv) > 2 else "c2"
gpu = sailor_b200.library()
w = bench.WORKLOADS[name]
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, seed=1)
keys = ("secondsCall", "secondsTraverse", "secondsExpand", "secondsFanOut", "secondsClassify", "secondsGather")
acc = {False: [], True: []}
with gpu.load_scene(path) as s:
    for rep in range(3 + 12):
        for on in (False, True):
            if on: os.environ[var] = "1"
            else: os.environ.pop(var, None)
            s.render_resident(p, rebuild_bvh=True, output_stage=True)
            st = gpu.stats()
            if rep >= 3: acc[on].append([st[k] * 1e3 for k in keys])
for on in (False, True):
    a = np.median(np.array(acc[on]), axis=0)
    print("%s %-6s" % (var, "set" if on else "unset"), " ".join("%s %.3f" % (k[7:], v) for k, v in zip(keys, a)))
