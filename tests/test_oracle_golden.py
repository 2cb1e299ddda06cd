"""Pins the oracle: the library built from the reference's sources must reproduce the committed golden vectors
(tests/golden/make_golden.py), and the restated front end must agree with the reference's bundled models."""
import os

import numpy as np
import pytest

import parity_checks as pc
import scenes
from conftest import REFERENCE
from sailor_b200.capi import Params
from golden.make_golden import BIG_HIT_CASES, HIT_CASES


def _scene(scene_dir, name, kw):
    return scenes.ensure(scene_dir, name, **kw)


def test_oracle_reproduces_flatten_and_bvh_goldens(oracle, G, scene_dir):
    for key, name, kw in (("cube", "cube", {}), ("pbr", "pbr", {}), ("hf64", "heightfield", {"n": 64})):
        pc.check_flatten(oracle, _scene(scene_dir, name, kw), G[key + "_tris"], G[key + "_mat"])
        pc.check_bvh(oracle, _scene(scene_dir, name, kw), G[key + "_nodes"], G[key + "_mapping"])


@pytest.mark.parametrize("case", HIT_CASES, ids=[c[0] for c in HIT_CASES])
def test_oracle_reproduces_primary_hit_goldens(oracle, G, scene_dir, case):
    name, scene, kw, h, wo, cam = case
    pc.check_primary_hits(oracle, _scene(scene_dir, scene, kw), h, wo, cam, G[name + "_cam"], G[name + "_hits"])


def test_oracle_reproduces_c1_digest(oracle, digests, scene_dir):
    name, scene, kw, h, wo, cam = BIG_HIT_CASES[0]
    with oracle.load_scene(_scene(scene_dir, scene, kw)) as s:
        hits = s.primary_hits(Params(height=h, width_override=wo, camera=cam))
    assert pc.sha(hits) == digests[name + "_hits"]["sha256"]
    assert int((hits["triId"] != pc.NOHIT).sum()) == 17288   # unit cube, default camera, 682x512


def test_oracle_reproduces_function_goldens(oracle, G, scene_dir):
    pc.check_textures(oracle, _scene(scene_dir, "pbr", {}), G["uv_grid"], [G["pbr_tex%d" % t] for t in range(4)])
    assert np.array_equal(oracle.output_stage(G["output_in"]), G["output_srgb"])
    got = oracle.eval_lighting(G["lighting_in"])
    assert np.array_equal(pc.bits(np.nan_to_num(got)), pc.bits(np.nan_to_num(G["lighting_out"])))


def test_generated_cube_is_the_bundled_box(oracle, scene_dir):
    box = os.path.join(REFERENCE, "Content", "Models", "Box", "Box.gltf")
    if not os.path.exists(box):
        pytest.skip("reference checkout absent")
    with oracle.load_scene(box) as a, oracle.load_scene(_scene(scene_dir, "cube", {})) as b:
        assert np.array_equal(pc.bits(a.triangles()[0]), pc.bits(b.triangles()[0]))
        assert a.counts() == b.counts()
        p = Params(height=64)
        pc.assert_hits_equal(a.primary_hits(p), b.primary_hits(p))


def test_oracle_render_is_reproducible_and_seeded(oracle, scene_dir):
    p = Params(height=16, num_samples=2, num_ambient_samples=2, max_bounces=2, msaa=2, ambient=(1, 1, 1), seed=5)
    with oracle.load_scene(_scene(scene_dir, "pbr", {})) as s:
        a, _ = s.render(p)
        b, _ = s.render(p)
        p.seed = 6
        c, _ = s.render(p)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
