"""Per-step device time of the resident bench step, to see where run-to-run variance comes from. Not a bench."""
import os, sys, time, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes, sailor_b200, bench
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if len(sys.argv) > 3 and sys.argv[3] == "torch":
    import torch; torch.cuda.set_device(0); x = torch.empty(16, device="cuda")
w = bench.WORKLOADS[name]
gpu = sailor_b200.library()
path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
p = bench.make_params(w, seed=1)
with gpu.load_scene(path) as s:
    rows = []
    for i in range(n):
        if i == 2: os.environ["SAILOR_PT_TRACE_HOST"] = "1"
        t0 = time.perf_counter()
        s.render_resident(p, rebuild_bvh=True, output_stage=True)
        t1 = time.perf_counter()
        st = gpu.stats()
        rows.append((st["secondsCall"] * 1e3, (t1 - t0) * 1e3, st["secondsBvhBuild"] * 1e3, st["secondsTraverse"] * 1e3, st["secondsShade"] * 1e3))
    for i, r in enumerate(rows):
        print(i, " ".join("%.2f" % v for v in r))
