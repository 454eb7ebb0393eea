"""Pipelined transfers (include/ps3d.h "Pipelined transfers"; SURVEY.md §8(f) presenter / read-back path): frames whose
vertex streams arrive by ps3d_vbo_update_async on the copy stream, from double-buffered VBO sets, and whose images leave
by ps3d_read_colour_async on the read-back stream, must be the frames the synchronous API renders — bit for bit, and
against the oracle — however the three streams interleave."""
import numpy as np
import pytest
import torch

from _compare import colour_stats, render_all
from _scenes_small import SMALL
from puresoft3d_b200 import scenes, sortfirst
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1_cube_def03", "c2_heightfield_small", "c4_blend_overdraw"])
def test_pipelined_frames_equal_synchronous_frames(name, cuda_lib, oracle_lib):
    sc = SMALL[name]()
    want = render_all(cuda_lib, sc)["colour"].view(np.uint32)          # the synchronous API
    frac, _ = colour_stats(want, render_all(oracle_lib, sc)["colour"].view(np.uint32))
    assert frac >= 0.999                                                # ... which is the oracle's frame (1/255 on >= 99.9 %)
    dev = torch.device("cuda", 0)
    pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
    ups = [scenes.upload(pipe, sc), scenes.upload(pipe, sc)]
    uploaders, images, poisons = [], [], []
    for u in ups:
        items, bad = [], []
        for vbo, arr in u.vbos:
            host = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).pin_memory()
            items.append((vbo, host))
            bad.append((vbo, np.full(arr.shape, 1.0e9, dtype=arr.dtype)))
        uploaders.append(sortfirst.ShardedUpload(pipe, items, 0, 1, dev))
        poisons.append(bad)
        images.append(torch.zeros((sc.height, sc.width), dtype=torch.int32).pin_memory())
    # garbage in both VBO sets first: the frames below can only be right if every asynchronous upload landed before its draw
    for bad in poisons:
        for vbo, junk in bad:
            vbo.updateContent(junk)
    for i in range(6):
        s = i & 1
        uploaders[s].step()
        scenes.replay(pipe, sc, ups[s], finish=False)
        pipe.readColourAsync(images[s].data_ptr(), sc.width * 4)
        pipe.swapBuffers()
    pipe.finish()
    # clear4 never touches the last buffer row (fbo.cpp:336, replicated), so a blending scene accumulates there from
    # frame to frame on a long-lived target, as it would in the reference: that row is not part of the comparison
    for s in range(2):
        assert np.array_equal(images[s].numpy().view(np.uint32)[:-1], want[:-1]), "set %d" % s
    # partial ranges: the first half synchronously wrong, then fixed by an asynchronous range update
    vbo, host = uploaders[0].items[0][0], uploaders[0].items[0][1]
    half = vbo.unitCount // 2
    junk = np.full(vbo.unitCount * vbo.unitBytes // 4, 3.0e8, dtype=np.float32)
    vbo.updateContent(junk)
    vbo.updateContentAsync(host.data_ptr(), 0, half)
    vbo.updateContentAsync(host.data_ptr() + half * vbo.unitBytes, half, vbo.unitCount - half)
    scenes.replay(pipe, sc, ups[0], finish=False)
    pipe.readColourAsync(images[0].data_ptr(), sc.width * 4)
    pipe.finish()
    assert np.array_equal(images[0].numpy().view(np.uint32)[:-1], want[:-1])
    with pytest.raises(Exception):
        vbo.updateContentAsync(host.data_ptr(), vbo.unitCount, 1)
    pipe.close()


def test_async_write_waits_for_a_draw_of_a_synchronously_filled_vbo(cuda_lib):
    """A VBO filled only by the synchronous updateContent and then drawn: the first asynchronous write must wait for that
    draw's geometry kernel (its lastRead event is recorded for EVERY attached VBO, not only for ones written asynchronously
    before). Sync update, draw without finish, async update with different data, draw: frame 1 must be the frame of the first
    data, frame 2 the frame of the second."""
    sc = SMALL["c2_heightfield_small"]()
    want_a = render_all(cuda_lib, sc)["colour"].view(np.uint32)
    images = [torch.zeros((sc.height, sc.width), dtype=torch.int32).pin_memory() for _ in range(2)]
    for rep in range(3):                    # several rounds on fresh pipes: a race does not lose every time
        pipe = PuresoftPipeline(sc.width, sc.height, lib=cuda_lib)
        up = scenes.upload(pipe, sc)        # synchronous updateContent on every VBO, nothing asynchronous before the draw
        junk = [torch.from_numpy(np.full(arr.size * arr.itemsize // 4, 2.0e9, dtype=np.float32)).pin_memory() for _, arr in up.vbos]
        scenes.replay(pipe, sc, up, finish=False)                       # draw, no finish
        pipe.readColourAsync(images[0].data_ptr(), sc.width * 4)
        pipe.swapBuffers()
        for (vbo, arr), j in zip(up.vbos, junk):                        # different data, asynchronously, right behind the draw
            vbo.updateContentAsync(j.data_ptr(), 0, vbo.unitCount)
        scenes.replay(pipe, sc, up, finish=False)
        pipe.readColourAsync(images[1].data_ptr(), sc.width * 4)
        pipe.swapBuffers()
        pipe.finish()
        assert np.array_equal(images[0].numpy().view(np.uint32)[:-1], want_a[:-1]), "the asynchronous write overtook the draw (round %d)" % rep
        assert not np.array_equal(images[1].numpy().view(np.uint32)[:-1], want_a[:-1]), "the second draw did not see the asynchronous write"
        pipe.close()
