"""Bilinear sampling (ps3d_texture_set_filter) — an EXTENSION: the reference's PuresoftSampler2D is nearest-only
(samplr2d.cpp:19-25), so there is nothing of the reference to pin this against ("parity unpinned").
What is checked: the CPU restatement against an independent numpy statement of the definition in include/ps3d.h, the
properties any bilinear filter has, that the reference build refuses the mode, and (on the GPU) CUDA == restatement."""
import numpy as np
import pytest

from _compare import colour_stats, render_all
from _scenes_small import EXTENSION
from puresoft3d_b200 import _capi as K
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline


def numpy_bilinear(tex, u, v, wrap):
    """include/ps3d.h: texel i at u = i/width; four taps through clampCoord; lerp(a,b,t) = a + (b-a)*t in fp32; +0.5, trunc."""
    h, w = tex.shape
    f = np.float32
    x, y = f(f(w) * f(u)), f(f(h) * f(v))
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = f(x - x0), f(y - y0)

    def clamp(r, c):
        if wrap == K.WRAP_CLAMP:
            return min(max(r, 0), h - 1), min(max(c, 0), w - 1)
        mr, mc = h - 1, w - 1
        r = int(np.fmod(r, mr)) if mr else 0
        c = int(np.fmod(c, mc)) if mc else 0
        return (r + mr if r < 0 else r), (c + mc if c < 0 else c)

    def tap(r, c):
        r, c = clamp(int(r), int(c))
        return tex[r, c]

    c00, c10, c01, c11 = tap(y0, x0), tap(y0, x0 + 1), tap(y0 + 1, x0), tap(y0 + 1, x0 + 1)
    out = 0
    for ch in range(4):
        a, b, c, d = (f((t >> (8 * ch)) & 0xFF) for t in (int(c00), int(c10), int(c01), int(c11)))
        lerp = lambda p, q, t: f(p + f(f(q - p) * t))
        r = lerp(lerp(a, b, fx), lerp(c, d, fx), fy)
        q = min(max(int(f(r + f(0.5))), 0), 255)
        out |= q << (8 * ch)
    return out


def test_numpy_definition_on_known_values():
    tex = np.array([[0x00000000, 0x000000FF], [0x0000FF00, 0x00FF0000]], dtype=np.uint32)
    assert numpy_bilinear(tex, 0.0, 0.0, K.WRAP_CLAMP) == 0                      # exactly texel (0,0)
    assert numpy_bilinear(tex, 0.5, 0.0, K.WRAP_CLAMP) == 0x000000FF              # x = 1.0: exactly texel (0,1)
    assert numpy_bilinear(tex, 0.25, 0.0, K.WRAP_CLAMP) == 0x00000080             # halfway: 127.5 + 0.5 -> 128
    assert numpy_bilinear(tex, 0.25, 0.25, K.WRAP_CLAMP) == 0x00404040            # centre of the four: 63.75 + 0.5 -> 64


def probe_expected(tex, width, height, wrap, uv_scale, uv_offset, bilinear):
    """The image of scenes.scene_texprobe with IDEAL uv = (x/W, y/H)*scale+offset (the renderer's uv differ from these by
    a few ulp of chain rounding). Memory row 0 of the top-down colour target is raster row H-1."""
    out = np.zeros((height, width), dtype=np.uint32)
    f = np.float32
    for y in range(height):
        for x in range(width):
            u = f(f(x) / f(width) * f(uv_scale) + f(uv_offset))
            v = f(f(y) / f(height) * f(uv_scale) + f(uv_offset))
            out[height - 1 - y, x] = numpy_bilinear(tex, u, v, wrap)
    return out


@pytest.mark.parametrize("wrap,scale,offset", [(K.WRAP_CLAMP, 1.0, 0.0), (K.WRAP_CLAMP, 1.5, -0.25), (K.WRAP_WRAP, 2.5, -0.75)])
def test_oracle_bilinear_matches_numpy_statement(wrap, scale, offset, oracle_lib):
    """The pin this extension can have: the CPU restatement's sampler, probed through the TEXPROBE functor, against an
    independent numpy statement of the definition. uv reach the sampler through the reference's interpolation, whose
    span ends sit on ROUNDED columns (interp.cpp:151-160): next to the quad's diagonal v is off by up to
    0.4 rows. Hence a ramp texture (<= 16 levels per texel, magnified 5-8x: < 1.5 levels of such drift) and the gate
    "every channel within 2/255 on >= 99 % of pixels"; a wrong tap, weight or rounding rule is off by 4-16 levels over
    whole regions (WRAP: also across the seam)."""
    r, c = np.mgrid[0:8, 0:8]
    tex = ((16 * c + 4 * r) | ((200 - 12 * r - 3 * c) << 8) | ((8 * c + 8 * r) << 16)).astype(np.uint32)
    if wrap == K.WRAP_WRAP:   # WRAP is modulo (size - 1) = 7 (fbo.cpp:582-590): a texture continuous across that seam
        s7, c7 = np.sin(2 * np.pi * c / 7.0), np.cos(2 * np.pi * r / 7.0)
        tex = ((120 + 12 * s7).astype(np.int64) | ((120 + 12 * c7).astype(np.int64) << 8) | ((120 + 12 * s7 * c7).astype(np.int64) << 16)).astype(np.uint32)
    sc = scenes.scene_texprobe(64, 48, tex=tex, bilinear=True, wrap=wrap, uv_scale=scale, uv_offset=offset)
    got = render_all(oracle_lib, sc, capture=False)["colour"]
    want = probe_expected(tex, 64, 48, wrap, scale, offset, True)
    # raster row 0 is skipped by clearColour (fbo.cpp:336) but drawn by the quad; the quad's top row H is off-screen
    ca = got.view(np.uint8).reshape(48, 64, 4).astype(np.int16)[:, :, :3]
    cb = want.view(np.uint8).reshape(48, 64, 4).astype(np.int16)[:, :, :3]
    d = np.abs(ca - cb).max(axis=-1)
    assert (d <= 2).mean() >= 0.99, ((d <= 2).mean(), int(d.max()))


@pytest.mark.parametrize("name", sorted(EXTENSION))
def test_oracle_bilinear_differs_from_nearest_only_in_colour(name, oracle_lib):
    sc = EXTENSION[name]()
    a = render_all(oracle_lib, sc)
    for t in sc.textures:
        t["bilinear"] = False
    b = render_all(oracle_lib, sc)
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert not np.array_equal(a["colour"], b["colour"])


def test_constant_texture_is_filter_invariant(oracle_lib):
    sc = scenes.scene_cube(160, 120, bilinear=True)
    sc.textures[0]["layers"][0] = np.full((256, 256), 0x00A0B0C0, dtype=np.uint32)
    a = render_all(oracle_lib, sc)
    sc.textures[0]["bilinear"] = False
    b = render_all(oracle_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])


def test_reference_build_refuses_bilinear(ref_lib):
    p = PuresoftPipeline(64, 64, lib=ref_lib)
    try:
        t = p.createTexture(4, 4, 4, pixels=np.zeros((4, 4), dtype=np.uint32))
        p.setTextureFilter(t, False)
        with pytest.raises(Exception):
            p.setTextureFilter(t, True)
    finally:
        p.close()


def _probe_scene():
    rng = np.random.default_rng(6)
    tex = scenes.tex_random_bgra(rng, 16, 16).view(np.uint32).reshape(16, 16)
    return scenes.scene_texprobe(200, 120, tex=tex, bilinear=True, wrap=K.WRAP_WRAP, uv_scale=3.0, uv_offset=-1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EXTENSION) + ["texprobe"])
def test_cuda_matches_oracle_bilinear(name, cuda_lib, oracle_lib):
    sc = _probe_scene() if name == "texprobe" else EXTENSION[name]()
    a, b = render_all(cuda_lib, sc), render_all(oracle_lib, sc)
    assert np.array_equal(a["counts"], b["counts"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    frac, worst = colour_stats(a["colour"], b["colour"])
    assert frac >= 0.999, (frac, worst)
    for key in ("triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key
