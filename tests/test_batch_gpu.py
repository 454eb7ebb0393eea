"""Batches of small draws (include/ps3d.h, ps3d_debug_batch_counts): consecutive small draws into the same targets run as one pass
on the CUDA side. The frame must equal the reference arithmetic's (the oracle renders the draws one after the other) bit for bit,
counters included — with three programmes of 3, 1 and 0 varyings mixed in one batch, per-draw uniforms and vertex streams, tile
lists that outgrow the speculated capacity (retry with exact sizes) and the shared-memory sort (the batch's draws go down the first
path one by one), a big draw in between (runs alone, splits the batch), and repeated frames (warm speculation)."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu

CASES = {
    "mixed_programmes": lambda: scenes.scene_small_draws(320, 200, seed=31, draws=24, tris=40),
    "one_triangle_draws": lambda: scenes.scene_small_draws(200, 120, seed=32, draws=40, tris=1),
    "lists_outgrow_speculation": lambda: scenes.scene_small_draws(320, 200, seed=33, draws=12, tris=30, crowd=60),
    "lists_outgrow_the_sort": lambda: scenes.scene_small_draws(320, 200, seed=34, draws=12, tris=20, crowd=260),
    "big_draw_in_between": lambda: scenes.scene_small_draws(320, 200, seed=35, draws=9, tris=25, big_every=4),
    "odd_size": lambda: scenes.scene_small_draws(238, 131, seed=36, draws=17, tris=33),
}


def _same(got, want):
    assert np.array_equal(got["colour"].view(np.uint32), want["colour"].view(np.uint32))
    assert np.array_equal(got["depth"].view(np.uint32), want["depth"].view(np.uint32))
    assert np.array_equal(got["counts"], want["counts"])
    for key in ("draws", "triangles_submitted", "triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
        assert got["stats"][key] == want["stats"][key], key


@pytest.mark.parametrize("name", sorted(CASES))
def test_batched_small_draws_equal_the_draws_one_by_one(name, cuda_lib, oracle_lib):
    sc = CASES[name]()
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    pipe.debugCapture(sc.width, sc.height)
    up = scenes.upload(pipe, sc)
    for frame in range(3):                               # cold speculation, then warm
        pipe.resetStats()
        pipe.debugClearShadeCounts()
        scenes.replay(pipe, sc, up)
        got = dict(colour=pipe.readColour(), depth=pipe.readDepth(), stats=pipe.getStats(), counts=pipe.debugReadShadeCounts())
        _same(got, want)
    batches, draws = pipe.debugBatchCounts()
    assert batches >= 3 and draws > batches, (batches, draws)
    pipe.close()


@pytest.mark.parametrize("name", ["c3_demo2_desk", "demo1_planets", "demo2_objx_file"])
def test_demo_frames_run_in_batches(name, cuda_lib, oracle_lib):
    sc = SMALL[name]()
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    pipe.debugCapture(sc.width, sc.height)
    scenes.render(pipe, sc)
    got = dict(colour=pipe.readColour(), depth=pipe.readDepth(), stats=pipe.getStats(), counts=pipe.debugReadShadeCounts())
    batches, draws = pipe.debugBatchCounts()
    pipe.close()
    assert np.array_equal(got["depth"].view(np.uint32), want["depth"].view(np.uint32))
    assert np.array_equal(got["counts"], want["counts"])
    assert np.array_equal(got["colour"].view(np.uint32), want["colour"].view(np.uint32))
    if name != "demo1_planets":                          # (demo 1's draws alternate targets and blend: nothing to batch)
        assert batches >= 1 and draws >= 2 * batches, (batches, draws)


def test_captured_frame_of_small_draws(cuda_lib, oracle_lib):
    sc = CASES["mixed_programmes"]()
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    for _ in range(2):
        scenes.replay(pipe, sc, up)
    pipe.graphBegin()
    scenes.replay(pipe, sc, up, finish=False)
    g = pipe.graphEnd()
    for _ in range(3):
        pipe.graphLaunch(g)
    pipe.finish()
    assert np.array_equal(pipe.readColour().view(np.uint32), want["colour"].view(np.uint32))
    assert np.array_equal(pipe.readDepth().view(np.uint32), want["depth"].view(np.uint32))
    pipe.graphDestroy(g)
    pipe.close()


@pytest.mark.parametrize("name", ["c3_demo2_desk", "demo2_objx_file", "demo1_planets"])
def test_depth_only_passes_without_the_parity_hook(name, cuda_lib, oracle_lib):
    """Without the per-pixel count hook a pass whose fragment functor does nothing (FP_Null: the shadow maps) runs no shade kernel and
    keeps no survivor stream: the shadow map it leaves (seen through the lit pass), the frame and the counters must not change."""
    sc = SMALL[name]()
    want = render_all(oracle_lib, sc, capture=False)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    for frame in range(3):
        pipe.resetStats()
        scenes.replay(pipe, sc, up)
        st = pipe.getStats()
        assert np.array_equal(pipe.readDepth().view(np.uint32), want["depth"].view(np.uint32)), frame
        assert np.array_equal(pipe.readColour().view(np.uint32)[:-1], want["colour"].view(np.uint32)[:-1]), frame
        for key in ("draws", "triangles_submitted", "triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
            assert st[key] == want["stats"][key], (frame, key)
    pipe.close()
