"""One bench step of a workload bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py c3

Not a bench: numbers printed under a profiler are never reported.
"""
import ctypes
import os
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
import sailor_b200  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # divide width/height (full ncu captures replay each launch ~40x)
w = bench.WORKLOADS[name]
gpu = sailor_b200.library()
rt = ctypes.CDLL("libcuda.so.1")          # cuProfilerStart/Stop are process-wide (the product links cudart statically)
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, width=w["width"] // scale, height=w["height"] // scale, seed=1)
with gpu.load_scene(path) as s:
    s.render_resident(p, rebuild_bvh=True, output_stage=True)     # warm-up (allocations, caches)
    rt.cuProfilerStart()
    s.render_resident(p, rebuild_bvh=True, output_stage=True)
    rt.cuProfilerStop()
    print(gpu.stats())
