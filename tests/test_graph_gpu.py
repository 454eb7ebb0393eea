"""Captured frames (include/ps3d.h, ps3d_graph_*): the calls of a frame recorded once into a CUDA graph and replayed as one launch
must render the frame the same calls render one by one — bit for bit, counters included — for single-draw and multi-draw frames,
with vertex data changed between launches, and a capture without a sized warm-up frame must fail loudly, not silently."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c2_heightfield_small", "c3_demo2_desk", "c4_blend_overdraw", "demo1_planets", "soup_nocull"])
def test_captured_frame_equals_the_frame_call_by_call(name, cuda_lib):
    sc = SMALL[name]()
    want = render_all(cuda_lib, sc)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    pipe.debugCapture(sc.width, sc.height)
    up = scenes.upload(pipe, sc)
    for _ in range(2):                               # normally first: buffers are sized, speculation has settled
        scenes.replay(pipe, sc, up)
    pipe.graphBegin()
    scenes.replay(pipe, sc, up, finish=False)
    g = pipe.graphEnd()
    frames = 3
    pipe.resetStats()
    pipe.debugClearShadeCounts()
    for _ in range(frames):
        pipe.graphLaunch(g)
    pipe.finish()
    # clear4 never touches the last buffer row (fbo.cpp:336, replicated): a blending scene accumulates there from frame to frame
    assert np.array_equal(pipe.readColour().view(np.uint32)[:-1], want["colour"].view(np.uint32)[:-1])
    assert np.array_equal(pipe.readDepth().view(np.uint32), want["depth"].view(np.uint32))
    st = pipe.getStats()
    for key in ("draws", "triangles_submitted", "triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
        assert st[key] == frames * want["stats"][key], key
    assert np.array_equal(pipe.debugReadShadeCounts(), frames * want["counts"])
    pipe.graphDestroy(g)
    pipe.close()


def test_captured_frame_reads_the_vertex_data_of_the_moment(cuda_lib):
    """Vertex data may change between launches: the streams are read through the same pointers."""
    sc = SMALL["c2_heightfield_small"]()
    want = render_all(cuda_lib, sc)["colour"].view(np.uint32)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    for _ in range(2):
        scenes.replay(pipe, sc, up)
    pipe.graphBegin()
    scenes.replay(pipe, sc, up, finish=False)
    g = pipe.graphEnd()
    vbo, arr = up.vbos[0]
    vbo.updateContent(np.full(arr.shape, 3.0e9, dtype=arr.dtype))   # every triangle off screen
    pipe.graphLaunch(g)
    pipe.finish()
    assert not np.array_equal(pipe.readColour().view(np.uint32)[:-1], want[:-1])
    vbo.updateContent(np.ascontiguousarray(arr))
    pipe.graphLaunch(g)
    pipe.finish()
    assert np.array_equal(pipe.readColour().view(np.uint32)[:-1], want[:-1])
    pipe.close()


def test_capture_without_a_sized_frame_fails_loudly(cuda_lib):
    sc = SMALL["c2_heightfield_small"]()
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    pipe.graphBegin()
    with pytest.raises(Exception):
        # nothing has sized the draw's buffers: a captured frame cannot allocate. (A small draw is launched by the call that
        # ends its batch — include/ps3d.h — which here is the end of the capture.)
        scenes.replay(pipe, sc, up, finish=False)
        pipe.graphEnd()
    try:
        pipe.graphEnd()
    except Exception:  # noqa: BLE001 — whether the truncated capture still ends cleanly is not the point
        pass
    # ... and the pipe is usable afterwards
    scenes.replay(pipe, sc, up)
    assert pipe.getStats()["fragments_shaded"] > 0
    pipe.close()


def test_oracle_library_has_no_graphs(oracle_lib):
    sc = SMALL["c1_cube_def01"]()
    pipe = PuresoftPipeline(sc.width, sc.height, lib=oracle_lib)
    with pytest.raises(Exception):
        pipe.graphBegin()
    pipe.close()


def test_frame_that_needs_a_host_decision_is_refused_loudly(cuda_lib):
    """crowded_tile has a tile list too long for the shared-memory sort: its draw needs the radix path, which the host chooses after
    reading the draw's report — a captured frame has no such point. The replay must say so (at finish), not drop the draw."""
    sc = SMALL["crowded_tile"]()
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    up = scenes.upload(pipe, sc)
    for _ in range(2):
        scenes.replay(pipe, sc, up)
    pipe.graphBegin()
    scenes.replay(pipe, sc, up, finish=False)
    g = pipe.graphEnd()
    pipe.graphLaunch(g)
    with pytest.raises(ValueError):
        pipe.finish()
    # the same frame call by call still renders
    want = render_all(cuda_lib, sc)["colour"].view(np.uint32)
    scenes.replay(pipe, sc, up)
    assert np.array_equal(pipe.readColour().view(np.uint32)[:-1], want[:-1])
    pipe.close()


def test_launch_after_the_scratch_buffers_moved_is_refused(cuda_lib):
    """The recorded kernels hold the addresses of the pipe's scratch buffers: once a bigger draw, submitted normally, has made
    one of them grow (freed and allocated anew), replaying the captured frame would write through a dangling pointer."""
    small = SMALL["c1_cube_def01"]()
    big = scenes.scene_soup(small.width, small.height, seed=3, count=30000)
    pipe = PuresoftPipeline(small.width, small.height, lib=cuda_lib)
    up_small, up_big = scenes.upload(pipe, small), scenes.upload(pipe, big)
    for _ in range(2):
        scenes.replay(pipe, small, up_small)
    pipe.graphBegin()
    scenes.replay(pipe, small, up_small, finish=False)
    g = pipe.graphEnd()
    pipe.graphLaunch(g)
    pipe.finish()
    scenes.replay(pipe, big, up_big)                 # 30 000 triangles: headers, varyings, span records outgrow the cube's (megabytes: new address ranges)
    with pytest.raises(ValueError):
        pipe.graphLaunch(g)
    pipe.graphDestroy(g)
    scenes.replay(pipe, small, up_small)             # the pipe itself is fine
    pipe.close()
