"""Host-side timeline of one frame (SAILOR_PT_TRACE_HOST=1): python tools/host_trace.py c3"""
import os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes, sailor_b200, bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[name]
gpu = sailor_b200.library()
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, seed=1)
with gpu.load_scene(path) as s:
    for i in range(3):
        if i == 2: os.environ["SAILOR_PT_TRACE_HOST"] = "1"
        s.render_resident(p, rebuild_bvh=True, output_stage=True)
        st = gpu.stats()
        print("frame", i, {k: round(st[k], 5) for k in ("secondsCall", "secondsTraverse", "secondsShade", "secondsBvhBuild")}, st["kernelLaunches"], flush=True)
