"""Experiment (GPU): how much does the ORDER of the level-0 ray queue matter to the traversal kernel?

Builds the rays a first-hit level would hold on the C3 heightfield (16 hemisphere rays from every primary hit, 8 primary
samples per pixel) and times SailorPt_IntersectRays on the same ray SET in different orders:
  shuffled   activations interleaved in chunks of 16 (what the atomically appended first-hit queue looks like today)
  tile       activations in (tile, sample, lane) order, each activation's rays consecutive
  sortN_B    groups of N consecutive activations, rays of a group stably sorted into B direction bins
Not a bench: a design probe whose numbers go to profiles/ as evidence for the ray-queue layout.
"""
import os, sys, tempfile, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenes
import sailor_b200
from sailor_b200.capi import Params

W, H, ROWS, NS, K = 1920, 1080, 128, 8, 16
scene_name = sys.argv[1] if len(sys.argv) > 1 else "heightfield"
kw = {"n": 707} if scene_name == "heightfield" else {}
gpu = sailor_b200.library()
path = scenes.ensure(tempfile.mkdtemp(), scene_name, **kw)
rng = np.random.default_rng(1)
with gpu.load_scene(path) as s:
    s.build_bvh()
    p = Params(height=H, width_override=W)
    w, h, cam = s.camera(p)
    hits = s.primary_hits(p)
    y0 = (H - ROWS) // 2
    ys, xs = np.meshgrid(np.arange(y0, y0 + ROWS), np.arange(W), indexing="ij")
    # (tile, sample, lane) order: 8x4 tiles
    ty, tx = ys // 4, xs // 8
    lane = (ys % 4) * 8 + (xs % 8)
    tile = (ty - ty.min()) * (W // 8) + tx
    order = np.lexsort((lane.ravel(), tile.ravel()))
    px = xs.ravel()[order]; py = ys.ravel()[order]
    px = px.reshape(-1, 32); py = py.reshape(-1, 32)                 # [tile, lane]
    px = np.repeat(px[:, None, :], NS, axis=1).reshape(-1); py = np.repeat(py[:, None, :], NS, axis=1).reshape(-1)   # [tile, sample, lane]
    hh = hits[py, px]
    ok = hh["triId"] != 0xFFFFFFFF
    px, py, hh = px[ok], py[ok], hh[ok]
    pos, p00, du, dv = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    d = p00[None, :] + (px[:, None] + 0.5) * du[None, :] + (py[:, None] - 0.5) * dv[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = (pos[None, :] + d * hh["t"][:, None]).astype(np.float32)
    nA = org.shape[0]
    # cosine-weighted hemisphere about +y
    u1 = rng.random((nA, K)); u2 = rng.random((nA, K))
    r = np.sqrt(u1); phi = 2 * np.pi * u2
    dirs = np.stack([r * np.cos(phi), np.sqrt(1 - u1), r * np.sin(phi)], axis=2).astype(np.float32)   # [nA, K, 3]
    O = np.repeat(org[:, None, :], K, axis=1)
    IG = np.repeat(hh["triId"][:, None], K, axis=1).astype(np.uint32)
    print("activations", nA, "rays", nA * K, flush=True)

    def run(name, perm):
        o = O.reshape(-1, 3)[perm]; dd = dirs.reshape(-1, 3)[perm]; ig = IG.reshape(-1)[perm]
        ts = []
        for _ in range(3):
            res = s.intersect_rays(o, dd, ig); ts.append(gpu.stats()["secondsTraverse"])
        t = min(ts)
        chk = int(np.bitwise_xor.reduce(res["triId"][np.argsort(perm)]))
        print(json.dumps({"order": name, "ms": round(t * 1e3, 3), "Grays_s": round(len(perm) / t / 1e9, 3), "hit_frac": round(float((res["triId"] != 0xFFFFFFFF).mean()), 3), "xor": chk}), flush=True)

    n = nA * K
    ident = np.arange(n)
    # shuffled: chunks of 16 activations in random order
    nch = (nA + 15) // 16
    chp = rng.permutation(nch)
    act = (chp[:, None] * 16 + np.arange(16)[None, :]).reshape(-1); act = act[act < nA]
    run("shuffled16", (act[:, None] * K + np.arange(K)[None, :]).reshape(-1))
    run("tile", ident)

    def dir_bins(dd, B):
        # octahedral map of the direction -> B x B grid (B*B bins)
        a = np.abs(dd).sum(axis=1, keepdims=True)
        pxy = dd[:, [0, 2]] / a
        neg = dd[:, 1] < 0
        q = (1 - np.abs(pxy[:, ::-1])) * np.sign(pxy + 1e-30)
        pxy = np.where(neg[:, None], q, pxy)
        g = np.clip(((pxy * 0.5 + 0.5) * B).astype(np.int64), 0, B - 1)
        return g[:, 0] * B + g[:, 1]

    flat = dirs.reshape(-1, 3)
    for N in (256, 1024, 4096):
        for B in (2, 4, 8):
            grp = (ident // K) // N
            key = grp * (B * B) + dir_bins(flat, B)
            run("sort%d_%d" % (N, B * B), np.argsort(key, kind="stable"))
    # octant only
    octant = (flat[:, 0] < 0) * 1 + (flat[:, 2] < 0) * 2
    for N in (256, 1024):
        run("oct%d" % N, np.argsort(((ident // K) // N) * 4 + octant, kind="stable"))
