"""End-to-end frame with the host's result buffers pinned (what bench.py's e2e leg does), per call. Not a bench."""
import os, sys, time, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes, bench, sailor_b200
gpu = sailor_b200.library()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[name]
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, seed=1)
with gpu.load_scene(path) as s0:
    wd, ht, _ = s0.camera(p)
lin = np.zeros((ht, wd, 3), np.float32); srgb = np.zeros((ht, wd, 3), np.uint8)
gpu.pin_host_buffer(lin); gpu.pin_host_buffer(srgb)
for rep in range(6):
    t0 = time.perf_counter(); s = gpu.load_scene(path); t1 = time.perf_counter()
    s.build_bvh(); t2 = time.perf_counter()
    s.render_resident(p, rebuild_bvh=False, output_stage=True); t3 = time.perf_counter(); st = gpu.stats()
    s.read_resident(p) if False else gpu.check(gpu.lib.SailorPt_ReadResident(s.h, lin.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)), srgb.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint8))), "read"); t4 = time.perf_counter()
    s.close(); t5 = time.perf_counter()
    print("%s rep%d load %.2f | bvh %.2f | render %.2f (gpu %.2f) | read %.2f | free %.2f | total %.2f ms" % (name, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["secondsCall"] * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t5 - t0) * 1e3), flush=True)
gpu.unpin_host_buffer(lin); gpu.unpin_host_buffer(srgb)
