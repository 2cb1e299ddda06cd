"""World-size-2 test of the multi-GPU host logic on CPU (gloo): each rank renders its shard through the C-ABI (the
host-compiled kernel bodies of tests/emu stand in for the GPU), the accumulators are reduced with the same code
bench.py uses over NCCL, and rank 0 must hold exactly the unsharded frame."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
import scenes
from sailor_b200.capi import Library, Params
from sailor_b200.distributed import reduce_accumulators, shard_params, split_range

BASE = dict(height=20, camera="main_cam", num_samples=2, num_ambient_samples=2, max_bounces=3, msaa=4, ambient=(1, 1, 1), seed=9)


def test_split_range_covers_everything_once():
    for n in (1, 7, 8, 1080):
        for world in (1, 2, 3, 8):
            parts = [split_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, lib_path, scene_path, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = Library(lib_path)
        with lib.load_scene(scene_path) as s:
            full = Params(**BASE)
            w, h, _ = s.camera(full)
            p = shard_params(full, rank, world, mode=mode, height=h)
            lin, _ = s.render(p, want_srgb=False)
            if mode == "rows":                      # rows outside the band must be zero for the SUM to assemble the frame
                b, e = p.rows
                mask = np.ones(h, bool); mask[h - e:h - b] = False
                assert not lin[mask].any()
        acc = torch.from_numpy(lin.astype(np.float32))
        reduce_accumulators(acc, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, "reduced_%s.npy" % mode), acc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "rows"])
def test_two_ranks_reassemble_the_unsharded_frame(emu, scene_dir, tmp_path, mode):
    path = scenes.ensure(scene_dir, "pbr")
    with emu.load_scene(path) as s:
        full, _ = s.render(Params(**BASE), want_srgb=False)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, emu.path, path, mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(str(tmp_path / ("reduced_%s.npy" % mode)))
    if mode == "rows":
        assert np.array_equal(got, full)            # every pixel comes from exactly one rank
    else:
        assert np.allclose(got, full, rtol=1e-6, atol=1e-7)   # fp32 sum of two partial accumulators vs one running sum
