"""Chain marks of long spans (device_types.cuh: SpanStreams). A pipe keeps marks once a draw has asked for them, so the SECOND
frame of a pipe is the first that uses them: every frame must stay bit-exact against the reference arithmetic (the oracle replays
every chain from the span start) — frames with marks, frames whose marks only partly fit (a scene with more long spans than the
one before it on the same pipe), batches of draws with marks, captured frames with marks."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _marks_for_small_scenes(monkeypatch):
    # a pipe keeps marks for draws that ask for 16384 or more (two extra launches per draw pay from there): read at ps3d_create
    monkeypatch.setenv("PS3D_MARKS_MIN", "1")

KEYS = ("draws", "triangles_submitted", "triangles_rasterised", "spans", "fragments_tested", "fragments_shaded")


def _frame(pipe, sc, up):
    pipe.resetStats()
    pipe.debugClearShadeCounts()
    scenes.replay(pipe, sc, up)
    return dict(colour=pipe.readColour(), depth=pipe.readDepth(), stats=pipe.getStats(), counts=pipe.debugReadShadeCounts())


def _same(got, want, what):
    assert np.array_equal(got["depth"].view(np.uint32), want["depth"].view(np.uint32)), what + ": depth"
    assert np.array_equal(got["counts"], want["counts"]), what + ": per-pixel shade counts"
    # (clear4 never touches the last buffer row, fbo.cpp:336: a pipe that rendered another scene before keeps that scene's last row)
    assert np.array_equal(got["colour"].view(np.uint32)[:-1], want["colour"].view(np.uint32)[:-1]), what + ": colour"
    for key in KEYS:
        assert got["stats"][key] == want["stats"][key], what + ": " + key


@pytest.mark.parametrize("name", ["c1_cube_640", "c1_cube_def03", "soup_def02", "soup_nocull", "soup_odd_size", "c3_demo2_desk",
                                  "demo2_objx_file", "crowded_tile", "ragged_streams", "small_draws_mixed", "c2_heightfield_small",
                                  "c4_blend_overdraw", "demo1_planets"])
def test_later_frames_of_a_pipe_equal_the_first(name, cuda_lib, oracle_lib):
    sc = SMALL[name]()
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    pipe.debugCapture(sc.width, sc.height)
    up = scenes.upload(pipe, sc)
    for frame in range(3):
        _same(_frame(pipe, sc, up), want, "frame %d" % frame)
    pipe.close()


def test_marks_that_only_partly_fit(cuda_lib, oracle_lib):
    few, many = SMALL["c1_cube_def01"](), SMALL["soup_def02"]()          # both 320 x 240
    assert (few.width, few.height) == (many.width, many.height)
    want_few, want_many = render_all(oracle_lib, few), render_all(oracle_lib, many)
    pipe = PuresoftPipeline(few.width, few.height, lib=cuda_lib)
    pipe.debugCapture(few.width, few.height)
    up_few, up_many = scenes.upload(pipe, few), scenes.upload(pipe, many)
    _same(_frame(pipe, few, up_few), want_few, "cube, no marks yet")
    _same(_frame(pipe, few, up_few), want_few, "cube, marks")
    _same(_frame(pipe, many, up_many), want_many, "soup, room for the cube's marks only")
    _same(_frame(pipe, many, up_many), want_many, "soup, marks")
    _same(_frame(pipe, few, up_few), want_few, "cube again")
    pipe.close()


def test_captured_frame_with_marks(cuda_lib, oracle_lib):
    sc = SMALL["c3_demo2_desk"]()
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    for _ in range(3):
        scenes.replay(pipe, sc, up)
    pipe.graphBegin()
    scenes.replay(pipe, sc, up, finish=False)
    g = pipe.graphEnd()
    for _ in range(2):
        pipe.graphLaunch(g)
    pipe.finish()
    assert np.array_equal(pipe.readDepth().view(np.uint32), want["depth"].view(np.uint32))
    assert np.array_equal(pipe.readColour().view(np.uint32), want["colour"].view(np.uint32))
    pipe.graphDestroy(g)
    pipe.close()


def test_very_large_rectangles_appended_by_the_grid(cuda_lib, oracle_lib):
    """A target of 64 x 48 tiles and triangles far larger than it: once the pipe keeps marks, rectangles of 2048 tiles or more are
    handed to tile_append_big_kernel instead of being appended by their geometry block."""
    sc = scenes.scene_crowded_tile(1024, 768, crowd=150)
    want = render_all(oracle_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    pipe.debugCapture(sc.width, sc.height)
    up = scenes.upload(pipe, sc)
    for frame in range(3):
        _same(_frame(pipe, sc, up), want, "frame %d" % frame)
    pipe.close()
