"""SailorPtParams::deviceCount on a multi-GPU box: one process, one host thread per device inside the library.
    python tools/multi_device_bench.py [workload] [frames]
Prints, for deviceCount = 1, 2, 4, 8 (as far as the box goes): wall time of SailorPt_RenderResident (BVH build + render + output stage),
Mrays/s, and whether the frame has the bits of the single-device frame."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench, scenes, sailor_b200

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.WORKLOADS[name]
L = sailor_b200.library()
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
have = torch.cuda.device_count()
ref = None
with L.load_scene(path) as s:
    for n in [d for d in (1, 2, 4, 8) if d <= have]:
        p = bench.make_params(w, seed=1); p.device_count = n
        best = None
        for _ in range(frames + 1):                      # the first frame of a device count creates the replicas (upload + flatten)
            t0 = time.perf_counter()
            s.render_resident(p, rebuild_bvh=True, output_stage=True)
            dt = time.perf_counter() - t0
            st = L.stats()
            best = dt if best is None or dt < best else best
        lin, srgb = s.read_resident(p, want_srgb=True)
        if ref is None:
            ref = (lin.copy(), srgb.copy())
        same = bool(np.array_equal(lin.view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(srgb, ref[1]))
        print(json.dumps({"workload": name, "deviceCount": n, "devicesUsed": st["devicesUsed"], "ms_per_frame": round(best * 1e3, 2), "mrays_per_s": round(st["rays"] / best / 1e6, 1),
                          "rays": st["rays"], "bits_equal_single_device": same}), flush=True)
