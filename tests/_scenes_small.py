"""The seeded scene set shared by the oracle pin, the golden fixtures and the GPU parity tests (sizes the CPU
checkers finish in well under a second each)."""
import os
import tempfile

from puresoft3d_b200 import _capi as K
from puresoft3d_b200 import scenes


def _demo2_from_objx(post=False, size=(384, 240)):
    """Demo 2's frame driven by an OBJX FILE (native reader, loadScene's re-centring, programme strings, material
    uniforms): the file is written by the native writer first, in the shape of the reference's plane.objx."""
    path = os.path.join(tempfile.gettempdir(), "ps3d_demo2_%d.objx" % os.getpid())
    scenes.write_demo_objx(path, seed=11, clutter=6)
    try:
        return scenes.scene_desk_objx(path, size[0], size[1], shadow=256, tex_size=128, post=post)
    finally:
        os.unlink(path)


SMALL = {
    "c1_cube_def01": lambda: scenes.scene_cube(320, 240),
    "c1_cube_def03": lambda: scenes.scene_cube(320, 240, functor=K.FN_DEF03),
    "c1_cube_640": lambda: scenes.scene_cube(640, 480),
    "soup_def02": lambda: scenes.scene_soup(320, 240, seed=7),
    "soup_nocull": lambda: scenes.scene_soup(200, 150, seed=9, cull=False),
    # odd, non-tile-aligned size; not W%4==1, where pipeline.cpp:31 under-allocates the depth rows
    "soup_odd_size": lambda: scenes.scene_soup(238, 131, seed=11, count=300),
    "c2_heightfield_small": lambda: scenes.scene_heightfield(480, 270, grid=40, layers=2, tex_size=256),
    "c2_heightfield_tiny_tris": lambda: scenes.scene_heightfield(256, 144, grid=96, layers=3, tex_size=128),
    "c4_blend_overdraw": lambda: scenes.scene_blend_overdraw(480, 270, randoms=200),
    # the reference's two demos as headless scenes: shadow pass + projective shadow lookup + cube skybox + blending + discard
    "demo1_planets": lambda: scenes.scene_planets(400, 250, shadow=240, stacks=12, slices=24, tex_size=128),
    "c3_demo2_desk": lambda: scenes.scene_desk(384, 240, shadow=256, clutter=10, tex_size=128),
    "demo2_objx_file": _demo2_from_objx,
    # awkward inputs: ragged / empty / dangling vertex streams; a tile list too long for the shared-memory sort (device-side
    # refusal of the speculated tail + exact retry on the radix path), triangles far larger than the viewport, w < 0
    "ragged_streams": lambda: scenes.scene_ragged_streams(200, 120),
    "crowded_tile": lambda: scenes.scene_crowded_tile(320, 200, crowd=3000),
    # + the post-processing hook (PP_DepthofField through the reference's own postProcess)
    "demo2_post": lambda: _demo2_from_objx(post=True),
    # many small draws in a row, programmes and uniforms changing from draw to draw, overlapping inside the depth dead band
    # (the CUDA side runs them as batches: include/ps3d.h, ps3d_debug_batch_counts); with tile lists that outgrow the sort
    "small_draws_mixed": lambda: scenes.scene_small_draws(320, 200, seed=31, draws=24, tris=40, flatid=False),
    "small_draws_crowded": lambda: scenes.scene_small_draws(320, 200, seed=34, draws=12, tris=20, crowd=260, flatid=False),
}

# Extensions with no reference counterpart: checked CUDA-vs-oracle only, "parity unpinned".
EXTENSION = {
    "c1_cube_bilinear": lambda: scenes.scene_cube(320, 240, bilinear=True),
    "c1_cube_bilinear_def03": lambda: scenes.scene_cube(320, 240, functor=K.FN_DEF03, bilinear=True),
    # a 16x16 texture magnified ~10x so that every pixel is a real blend, WRAP addressing (modulo size-1, fbo.cpp:582-590)
    "c1_cube_bilinear_wrap_magnified": lambda: scenes.scene_cube(320, 240, bilinear=True, wrap=K.WRAP_WRAP, tex_size=16),
}

# PP_DepthofField on an ODD width: each row's last 8-byte step runs into the next row's first pixel, which another worker
# read-modify-writes at the same time in the reference (src/test2/testpost.cpp:27-40) — a data race there, so column 0 is
# +50 or +100 from run to run. The restatement and the CUDA functor implement the sequential result (+100); everything
# but column 0 is still pinned against the reference (tests/test_post_process.py).
RACY_IN_REFERENCE = {
    "demo2_post_odd_width": lambda: _demo2_from_objx(post=True, size=(251, 160)),
}
