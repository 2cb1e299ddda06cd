"""Where an end-to-end call (host buffers in, host buffers out) spends its time. Not a bench."""
import os, sys, time, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes, bench
import sailor_b200

gpu = sailor_b200.library()
for name in sys.argv[1:] or ["c2", "c3"]:
    w = bench.WORKLOADS[name]
    path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
    p = bench.make_params(w, seed=1)
    for rep in range(4):
        t0 = time.perf_counter(); s = gpu.load_scene(path); t1 = time.perf_counter(); st_load = gpu.stats()
        s.build_bvh(); t2 = time.perf_counter(); st_b = gpu.stats()
        s.render_resident(p, rebuild_bvh=False, output_stage=True); t3 = time.perf_counter(); st_r = gpu.stats()
        lin, srgb = s.read_resident(p); t4 = time.perf_counter()
        s.close(); t5 = time.perf_counter()
        print("%s rep%d load %.1f ms (flatten %.2f) | bvh %.1f ms (gpu %.1f, %d launches) | render %.1f ms (gpu %.1f trav %.1f, %d launches, %.1f Mrays) | read %.1f ms | free %.1f ms | total %.1f ms" % (
            name, rep, (t1 - t0) * 1e3, st_load["secondsFlatten"] * 1e3, (t2 - t1) * 1e3, st_b["secondsBvhBuild"] * 1e3, st_b["kernelLaunches"],
            (t3 - t2) * 1e3, st_r["secondsCall"] * 1e3, st_r["secondsTraverse"] * 1e3, st_r["kernelLaunches"], st_r["rays"] / 1e6,
            (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3), flush=True)
