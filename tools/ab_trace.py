"""A/B on one box: the same C3 (or other) frames through the exact kernel, the origin-local walk (default) and the wide layout.
    python tools/ab_wide.py [workload] [frames] [modes, e.g. exact,fast]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, scenes, sailor_b200
from sailor_b200.capi import FLAG_EXACT_TRAVERSAL, FLAG_WIDE_TRAVERSAL

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = bench.WORKLOADS[name]
L = sailor_b200.library()
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
MODES = [m for m in (("exact", FLAG_EXACT_TRAVERSAL), ("fast", 0), ("wide", FLAG_WIDE_TRAVERSAL)) if len(sys.argv) <= 3 or m[0] in sys.argv[3].split(",")]
keys = ("secondsCall", "secondsTraverse", "secondsBvhBuild", "secondsExpand", "secondsFanOut", "secondsClassify", "secondsGather")
with L.load_scene(path) as s:
    imgs = {}
    for rnd in range(frames):
        for tag, flags in MODES:
            p = bench.make_params(w, seed=1); p.flags = flags
            s.render_resident(p, rebuild_bvh=True, output_stage=True)
            st = L.stats()
            print(tag, rnd, " ".join("%s %.3f" % (k[7:], st[k] * 1e3) for k in keys), "rays %.1fM" % (st["rays"] / 1e6), "replayed", st["replayedRays"],
                  "Grays/s %.3f" % (st["rays"] / st["secondsCall"] / 1e9), "trace-only Grays/s %.3f" % (st["rays"] / st["secondsTraverse"] / 1e9), flush=True)
            if rnd == 0:
                imgs[tag] = s.read_resident(p, want_srgb=False)[0]
    a = imgs[MODES[0][0]]
    for tag, _ in MODES[1:]:
        b = imgs[tag]
        print("%s vs %s image: identical bits" % (MODES[0][0], tag) if np.array_equal(a.view(np.uint32), b.view(np.uint32)) else
              "%s vs %s image: %d of %d floats differ, max abs %.3g, mean rel %.3g" % (MODES[0][0], tag, (a != b).sum(), a.size, np.abs(a - b).max(), np.abs(a - b).mean() / np.abs(a).mean()))
