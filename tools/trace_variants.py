"""Build (here, no GPU needed) or time (on the GPU box) tuning variants of the traversal kernel.

    python tools/trace_variants.py build      # writes gpurun_out/variants/*.so (travel with gpurun? no: gpurun_out is not sent) -> uses sailor_b200/variants/
    python tools/trace_variants.py run        # on the GPU: times every variant found
"""
import os, sys, time, tempfile, json
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
VDIR = os.path.join(ROOT, "sailor_b200", "variants")

VARIANTS = {
    "base": [],
    "stats": ["SPT_TRACE_STATS"],
    "tri": ["SPT_TRI_UNIFIED"],
    "idle8": ["SPT_FETCH_MIN_IDLE=8"],
    "idle4": ["SPT_FETCH_MIN_IDLE=4"],
    "dyn": ["SPT_DYN_REPS", "SPT_INNER_REPS=8", "SPT_LEAF_REPS=4"],
    "dyn_tri_idle8": ["SPT_DYN_REPS", "SPT_INNER_REPS=8", "SPT_LEAF_REPS=4", "SPT_TRI_UNIFIED", "SPT_FETCH_MIN_IDLE=8"],
    "reps21": ["SPT_INNER_REPS=2", "SPT_LEAF_REPS=1"],
    "reps63": ["SPT_INNER_REPS=6", "SPT_LEAF_REPS=3"],
    "stack8": ["SPT_SMEM_STACK=8"],
    "stack12": ["SPT_SMEM_STACK=12"],
    "stack16": ["SPT_SMEM_STACK=16"],
    "block64": ["SPT_TRACE_BLOCK=64"],
    "block256": ["SPT_TRACE_BLOCK=256"],
    "ib2": ["SPT_VOTE_INNER_BIAS=2"],
    "lb2": ["SPT_VOTE_LEAF_BIAS=2"],
    "reps41": ["SPT_INNER_REPS=4", "SPT_LEAF_REPS=1"],
    "reps43": ["SPT_INNER_REPS=4", "SPT_LEAF_REPS=3"],
    "reps32": ["SPT_INNER_REPS=3", "SPT_LEAF_REPS=2"],
    "reps82": ["SPT_INNER_REPS=8", "SPT_LEAF_REPS=2"],
    "idle24": ["SPT_FETCH_MIN_IDLE=24"],
    "fan4": ["SPT_FAN_MIN_BLOCKS=4"],
    "fan2": ["SPT_FAN_MIN_BLOCKS=2"],
}

if sys.argv[1] == "build":
    from sailor_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    names = sys.argv[2:] or list(VARIANTS)
    with ThreadPoolExecutor(max_workers=3) as ex:
        list(ex.map(lambda n: B.build(force=True, defines=VARIANTS[n], out=os.path.join(VDIR, "libvar_%s.so" % n)), names))
    print("built", names)
else:
    import numpy as np
    import scenes, bench
    from sailor_b200.capi import Library, Params
    d = tempfile.mkdtemp()
    hf = scenes.ensure(d, "heightfield", n=707)
    cube = scenes.ensure(d, "cube")
    res = {}
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        L = Library(os.path.join(VDIR, f))
        row = {}
        for tag, path, wl in (("hf", hf, "c3"), ("cube", cube, "c2")):
            w = bench.WORKLOADS[wl]
            with L.load_scene(path) as s:
                s.build_bvh()
                p0 = Params(height=1080, width_override=1920)
                ts = []
                for _ in range(4):
                    s.primary_hits(p0); ts.append(L.stats()["secondsTraverse"])
                row[tag + "_primary_Grays"] = round(1920 * 1080 / min(ts) / 1e9, 3)
                p = bench.make_params(w, seed=1)
                tr = []
                for _ in range(3):
                    s.render_resident(p, rebuild_bvh=False, output_stage=False); st = L.stats(); tr.append((st["secondsTraverse"], st["secondsCall"], st["rays"]))
                best = min(tr)
                row[tag + "_trace_ms"] = round(best[0] * 1e3, 2); row[tag + "_step_ms"] = round(best[1] * 1e3, 2); row[tag + "_trace_Grays"] = round(best[2] / best[0] / 1e9, 3)
        if hasattr(L.lib, "SailorPt_DebugTraceStats"):
            import ctypes
            buf = (ctypes.c_ulonglong * 16)()
            L.lib.SailorPt_DebugTraceStats(buf)          # clear
            with L.load_scene(hf) as s:
                s.build_bvh(); s.render_resident(bench.make_params(bench.WORKLOADS["c3"], seed=1), rebuild_bvh=False, output_stage=False)
            L.lib.SailorPt_DebugTraceStats(buf)
            v = list(buf)
            row["stats"] = dict(votes=v[0], idle_per_vote=v[1] / v[0], leaf_per_vote=v[2] / v[0], inner_per_vote=v[3] / v[0], inner_reps=v[4], lanes_per_inner_rep=v[5] / max(v[4], 1),
                                leaf_reps=v[6], lanes_per_leaf_rep=v[7] / max(v[6], 1), refills=v[8], lanes_per_refill=v[9] / max(v[8], 1))
        L.trim_memory()          # every loaded copy of the library owns its own shared arenas: give them back before the next variant
        res[f] = row
        print(f, json.dumps(row), flush=True)
