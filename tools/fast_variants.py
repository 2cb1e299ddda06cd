"""Build (here, no GPU needed) or time (on the GPU box) tuning variants of the origin-local traversal kernel (trace_fast.cuh).

    python tools/fast_variants.py build [names...]     # -> sailor_b200/variants/libvar_<name>.so (travels with gpurun)
    python tools/fast_variants.py run [workload]       # on the GPU: two frames per variant, traversal time + loop statistics
"""
import os, sys, tempfile, json, ctypes
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
VDIR = os.path.join(ROOT, "sailor_b200", "variants")

VARIANTS = {
    "base": [],
    "stats": ["SPT_FAST_LOOP_STATS"],
    "tv14": ["SPT_FAST_TRI_VOTE=14"], "tv16": ["SPT_FAST_TRI_VOTE=16"], "tv20": ["SPT_FAST_TRI_VOTE=20"],
    "r32": ["SPT_FAST_NODE_REPS=3", "SPT_FAST_TRI_REPS=2"], "r53": ["SPT_FAST_NODE_REPS=5", "SPT_FAST_TRI_REPS=3"], "r44": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=4"], "r64": ["SPT_FAST_NODE_REPS=6", "SPT_FAST_TRI_REPS=4"],
    "idle4": ["SPT_FAST_FETCH_MIN_IDLE=4"], "idle6": ["SPT_FAST_FETCH_MIN_IDLE=6"], "idle10": ["SPT_FAST_FETCH_MIN_IDLE=10"],
    "q4": ["SPT_FAST_LEAF_QUEUE=4"], "mb8": ["SPT_FAST_MIN_BLOCKS=8"], "mb6": ["SPT_FAST_MIN_BLOCKS=6"],
    "b256": ["SPT_FAST_BLOCK=256", "SPT_FAST_MIN_BLOCKS=4"], "b64": ["SPT_FAST_BLOCK=64", "SPT_FAST_MIN_BLOCKS=16"],
    "imm": ["SPT_FAST_IMMEDIATE"],
    "xskip": ["SPT_FAST_EXP_SKIP_CLOSEST"], "xonlyclosest": ["SPT_FAST_EXP_SKIP_ANY"], "xonlyclosest_stats": ["SPT_FAST_EXP_SKIP_ANY", "SPT_FAST_LOOP_STATS"], "xskip_stats": ["SPT_FAST_EXP_SKIP_CLOSEST", "SPT_FAST_LOOP_STATS"], "xskipany": ["SPT_FAST_EXP_SKIP_CLOSEST", "SPT_FAST_EXP_MODE=1"], "xskipany10": ["SPT_FAST_EXP_SKIP_CLOSEST", "SPT_FAST_EXP_MODE=1", "SPT_FAST_MIN_BLOCKS=10"], "xskipany9": ["SPT_FAST_EXP_SKIP_CLOSEST", "SPT_FAST_EXP_MODE=1", "SPT_FAST_MIN_BLOCKS=9"],
    "rl4610": ["SPT_FAST_REACH_L0=4u", "SPT_FAST_REACH_L1=6u", "SPT_FAST_REACH_L2=10u"], "rl61014": ["SPT_FAST_REACH_L0=6u", "SPT_FAST_REACH_L1=10u", "SPT_FAST_REACH_L2=14u"], "rl2612": ["SPT_FAST_REACH_L0=2u", "SPT_FAST_REACH_L1=6u", "SPT_FAST_REACH_L2=12u"],
    "tv22": ["SPT_FAST_TRI_VOTE=22"], "tv24": ["SPT_FAST_TRI_VOTE=24"], "tv24q16": ["SPT_FAST_TRI_VOTE=24", "SPT_FAST_LEAF_QUEUE=16"],
    "r42": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=2"], "r43": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=3"], "r22": ["SPT_FAST_NODE_REPS=2", "SPT_FAST_TRI_REPS=2"], "r23": ["SPT_FAST_NODE_REPS=2", "SPT_FAST_TRI_REPS=3"],
    "r24": ["SPT_FAST_NODE_REPS=2", "SPT_FAST_TRI_REPS=4"], "r33": ["SPT_FAST_NODE_REPS=3", "SPT_FAST_TRI_REPS=3"], "r34": ["SPT_FAST_NODE_REPS=3", "SPT_FAST_TRI_REPS=4"], "r46": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=6"],
    "cl4": ["SPT_CLASSIFY_MIN_BLOCKS=4"], "cl5": ["SPT_CLASSIFY_MIN_BLOCKS=5"], "cl6": ["SPT_CLASSIFY_MIN_BLOCKS=6"], "g3": ["SPT_GATHER_MIN_BLOCKS=3"], "g4": ["SPT_GATHER_MIN_BLOCKS=4"], "g6": ["SPT_GATHER_MIN_BLOCKS=6"], "g8": ["SPT_GATHER_MIN_BLOCKS=8"],
    "e3": ["SPT_EXPAND_MIN_BLOCKS=3"], "e1": ["SPT_EXPAND_MIN_BLOCKS=1"], "f5": ["SPT_FAN_MIN_BLOCKS=5"], "f3": ["SPT_FAN_MIN_BLOCKS=3"],
    "lazy0": ["SPT_FAST_LAZY_ANY=0"], "lazy1": ["SPT_FAST_LAZY_ANY=1"], "lazy2": ["SPT_FAST_LAZY_ANY=2"], "lazy3": ["SPT_FAST_LAZY_ANY=3"], "lazy1tv12": ["SPT_FAST_LAZY_ANY=1", "SPT_FAST_TRI_VOTE=12"], "lazy2tv14": ["SPT_FAST_LAZY_ANY=2", "SPT_FAST_TRI_VOTE=14"], "lazy1r24": ["SPT_FAST_LAZY_ANY=1", "SPT_FAST_NODE_REPS=2", "SPT_FAST_TRI_REPS=4"],
    "un22": ["SPT_FAST_NODE_UNROLL=2", "SPT_FAST_TRI_UNROLL=2"], "un44": ["SPT_FAST_NODE_UNROLL=4", "SPT_FAST_TRI_UNROLL=4"], "un21": ["SPT_FAST_NODE_UNROLL=2"], "un12": ["SPT_FAST_TRI_UNROLL=2"],
    "un64": ["SPT_FAST_NODE_REPS=6", "SPT_FAST_TRI_REPS=4", "SPT_FAST_NODE_UNROLL=6", "SPT_FAST_TRI_UNROLL=4"], "un55": ["SPT_FAST_NODE_REPS=5", "SPT_FAST_TRI_REPS=5", "SPT_FAST_NODE_UNROLL=5", "SPT_FAST_TRI_UNROLL=5"],
    "un66": ["SPT_FAST_NODE_REPS=6", "SPT_FAST_TRI_REPS=6", "SPT_FAST_NODE_UNROLL=6", "SPT_FAST_TRI_UNROLL=6"], "un46": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=6", "SPT_FAST_NODE_UNROLL=4", "SPT_FAST_TRI_UNROLL=6"],
    "un33": ["SPT_FAST_NODE_REPS=3", "SPT_FAST_TRI_REPS=3", "SPT_FAST_NODE_UNROLL=3", "SPT_FAST_TRI_UNROLL=3"], "un43": ["SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=3", "SPT_FAST_NODE_UNROLL=4", "SPT_FAST_TRI_UNROLL=3"],
    "un84": ["SPT_FAST_NODE_REPS=8", "SPT_FAST_TRI_REPS=4", "SPT_FAST_NODE_UNROLL=8", "SPT_FAST_TRI_UNROLL=4"],
    "ex2": ["SPT_EXACT_UNROLL=2"], "ex4": ["SPT_EXACT_UNROLL=4"], "ex8": ["SPT_EXACT_UNROLL=8"],
    "mb9": ["SPT_FAST_MIN_BLOCKS=9"], "mb10": ["SPT_FAST_MIN_BLOCKS=10"], "mb12": ["SPT_FAST_MIN_BLOCKS=12"],
    "mb8q4": ["SPT_FAST_MIN_BLOCKS=8", "SPT_FAST_LEAF_QUEUE=4"], "mb10q4": ["SPT_FAST_MIN_BLOCKS=10", "SPT_FAST_LEAF_QUEUE=4", "SPT_FAST_NODE_STACK=10"],
    "mb8r44": ["SPT_FAST_MIN_BLOCKS=8", "SPT_FAST_NODE_REPS=4", "SPT_FAST_TRI_REPS=4"], "b256mb4": ["SPT_FAST_BLOCK=256", "SPT_FAST_MIN_BLOCKS=4"], "b64mb16": ["SPT_FAST_BLOCK=64", "SPT_FAST_MIN_BLOCKS=16"],
}

if sys.argv[1] == "build":
    from sailor_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    names = sys.argv[2:] or list(VARIANTS)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(lambda n: B.build(force=True, defines=VARIANTS[n], out=os.path.join(VDIR, "libvar_%s.so" % n)), names))
    print("built", names)
else:
    import scenes, bench
    from sailor_b200.capi import Library
    wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
    w = bench.WORKLOADS[wl]
    path = scenes.ensure(tempfile.mkdtemp(), w["scene"], **w["kw"])
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        L = Library(os.path.join(VDIR, f))
        with L.load_scene(path) as s:
            s.build_bvh()
            p = bench.make_params(w, seed=1)
            tr = []
            for _ in range(3):
                s.render_resident(p, rebuild_bvh=False, output_stage=False); st = L.stats(); tr.append((st["secondsTraverse"], st["secondsCall"], st["rays"], st["replayedRays"], st["secondsExpand"], st["secondsFanOut"], st["secondsClassify"], st["secondsGather"]))
            best = min(tr)
            row = dict(trace_ms=round(best[0] * 1e3, 2), step_ms=round(best[1] * 1e3, 2), trace_Grays=round(best[2] / best[0] / 1e9, 3), replayed=best[3], expand=round(best[4] * 1e3, 2), fan=round(best[5] * 1e3, 2), classify=round(best[6] * 1e3, 2), gather=round(best[7] * 1e3, 2))
            if hasattr(L.lib, "SailorPt_DebugFastStats"):
                buf = (ctypes.c_ulonglong * 64)()
                L.lib.SailorPt_DebugFastStats(buf)          # clear
                s.render_resident(p, rebuild_bvh=False, output_stage=False)
                rays = L.stats()["rays"]
                L.lib.SailorPt_DebugFastStats(buf)
                v = list(buf)
                row["stats"] = dict(iterations=v[0], idle_per_iter=round(v[1] / max(v[0], 1), 2), node_steps=v[2], lanes_per_node_step=round(v[3] / max(v[2], 1), 2),
                                    tri_steps=v[4], lanes_per_tri_step=round(v[5] / max(v[4], 1), 2), refills=v[6], lanes_per_refill=round(v[7] / max(v[6], 1), 2),
                                    retired=v[8], forced_tri_votes=v[9], node_visits_per_ray=round(v[3] / max(v[8], 1), 2), tri_tests_per_ray=round(v[5] / max(v[8], 1), 2),
                                    up_lane_steps=v[10], anyhit_node_lane_steps=v[11], anyhit_tri_lane_steps=v[12], retired_anyhit=v[13], retired_anyhit_hit=v[14], retired_closest_hit=v[15],
                                    per_class={n: dict(rays=v[24 + c], node_per_ray=round(v[16 + c] / max(v[24 + c], 1), 2), tri_per_ray=round(v[20 + c] / max(v[24 + c], 1), 2)) for c, n in enumerate(("anyhit_hit", "anyhit_miss", "closest_hit", "closest_miss"))},
                                    anyhit_hit_node_hist=v[28:46], anyhit_hit_tri_hist=v[46:64])
        L.trim_memory()          # every loaded copy of the library owns its own shared arenas: give them back before the next variant
        print(f, json.dumps(row), flush=True)
