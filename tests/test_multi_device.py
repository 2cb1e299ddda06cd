"""SailorPtParams::deviceCount — the multi-device frame behind the entry API (replaces the reference's tile loop,
PathTracer.cpp:418-487).  CPU: the host-compiled build with every replica on the one (emulated) device, which exercises the band
scheduler, the per-device threads, the replica life cycle and the peer-copy path.  GPU: the same check on the real library, once with
the replicas forced onto one device (any box) and once over the devices the box has."""
import os

import numpy as np
import pytest

import parity_checks as pc
import scenes
from sailor_b200.capi import Params


@pytest.fixture()
def same_device(monkeypatch):
    monkeypatch.setenv("SAILOR_PT_MULTI_SAME_DEVICE", "1")


def test_multi_device_frame_equals_single_device_frame_cpu(emu, scene_dir, same_device):
    assert pc.check_multi_device_frame(emu, scenes.ensure(scene_dir, "pbr"), camera="main_cam") == 3


def test_device_count_is_clamped_to_the_box_cpu(emu, scene_dir):
    """Without the test switch the host-compiled build has ONE device: deviceCount = 8 renders the ordinary single-device frame."""
    with emu.load_scene(scenes.ensure(scene_dir, "cube")) as s:
        kw = dict(height=16, num_samples=2, num_ambient_samples=2, max_bounces=2, msaa=2, ambient=(1, 1, 1), seed=3)
        a, _ = s.render(Params(**kw)); b, _ = s.render(Params(device_count=8, **kw))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and emu.stats()["devicesUsed"] == 1


def test_scene_cache_serves_the_second_load_and_notices_a_changed_file(emu, tmp_path):
    """The parsed-scene cache (capi.cu) is keyed by path + mtime + size: the same file loads to the same triangles, a rewritten file is parsed again."""
    d = str(tmp_path)
    p = scenes.ensure(d, "cube")
    with emu.load_scene(p) as s:
        t0, m0 = s.triangles()
    with emu.load_scene(p) as s:
        t1, m1 = s.triangles()
    assert np.array_equal(t0, t1) and np.array_equal(m0, m1)
    os.remove(p)
    scenes.heightfield(p, n=3)                    # another scene under the same name
    os.utime(p, ns=(1, 1))
    with emu.load_scene(p) as s:
        assert s.counts()["triangles"] == 18
    emu.trim_memory()


@pytest.mark.gpu
def test_multi_device_frame_equals_single_device_frame_one_gpu(gpu, scene_dir, same_device):
    assert pc.check_multi_device_frame(gpu, scenes.ensure(scene_dir, "pbr"), camera="main_cam") == 3
    assert pc.check_multi_device_frame(gpu, scenes.ensure(scene_dir, "heightfield", n=64), height=40, devices=4) == 4


@pytest.mark.gpu
def test_multi_device_frame_over_the_devices_of_the_box(gpu, scene_dir):
    import torch
    n = torch.cuda.device_count()
    used = pc.check_multi_device_frame(gpu, scenes.ensure(scene_dir, "heightfield", n=64), height=64, devices=max(n, 2))
    assert used == max(n, 1) if n < 2 else used == n
