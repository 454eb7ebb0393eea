"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (where the CPU checkers take seconds to
minutes per frame): the three tile paths and the two binning paths are different parallel decompositions of the same
ordered computation, so they must agree word for word; sort-first bands must tile the frame; rendering is idempotent;
the counter of FragmentProcessor::process calls must equal the sum of the per-pixel capture."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL
from conftest import ROOT
from puresoft3d_b200 import scenes, sortfirst
from puresoft3d_b200.pipeline import PuresoftPipeline

pytestmark = pytest.mark.gpu


def run_hash(kind, **env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_render_hash.py"), kind], env=e, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("kind", ["c2_full", "c4_full", "soup_def02"])
def test_tile_and_binning_paths_agree(kind, built):
    base = run_hash(kind)
    assert base["stats"]["fragments_shaded"] == base["counts_sum"] > 0
    for env in ({"PS3D_TILE_PATH": "ordered"}, {"PS3D_TILE_PATH": "immediate"}, {"PS3D_BINNING": "radix"},
                {"PS3D_RASTER_PARTS": "1"}, {"PS3D_RASTER_PARTS": "2"}, {"PS3D_RASTER_PARTS": "4"},
                {"PS3D_SPECULATE": "0"}):
        other = run_hash(kind, **env)
        for key in ("depth", "counts", "colour"):
            assert other[key] == base[key], (env, key)
        for key in ("triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
            assert other["stats"][key] == base["stats"][key], (env, key)


def test_c2_full_size_is_idempotent_and_plausible(built):
    a, b = run_hash("c2_full"), run_hash("c2_full")
    assert a == b
    assert a["stats"]["triangles_submitted"] == 1002528
    assert a["covered"] > 0.9 * 1920 * 1080            # the mesh over-fills the viewport
    assert a["stats"]["fragments_shaded"] >= a["covered"]


@pytest.mark.parametrize("name,world", [("c2_heightfield_small", 2), ("c4_blend_overdraw", 3), ("soup_odd_size", 4), ("c2_heightfield_tiny_tris", 8),
                                        ("crowded_tile", 7)])
def test_row_bands_tile_the_frame(name, world, cuda_lib, oracle_lib):
    """Sort-first on one GPU: band by band into separate pipes; each band equals the oracle's band, their union the frame."""
    sc = SMALL[name]()
    whole = render_all(cuda_lib, sc)
    bands = sortfirst.row_bands(sc.height, world)
    composite = np.zeros_like(whole["colour"])
    shaded = 0
    for band in bands:
        outs = []
        for lib in (cuda_lib, oracle_lib):
            p = PuresoftPipeline(sc.width, sc.height, lib=lib)
            p.setRowBand(*band)
            scenes.render(p, sc)
            outs.append((p.readColour(), p.readDepth(), p.getStats()))
            p.close()
        (c0, d0, s0), (c1, d1, s1) = outs
        assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
        assert s0["fragments_shaded"] == s1["fragments_shaded"] and s0["fragments_tested"] == s1["fragments_tested"]
        m0, m1 = sortfirst.memory_rows(band, sc.height)
        composite[m0:m1] = c0[m0:m1]
        shaded += s0["fragments_shaded"]
    # clear4 leaves the last buffer row (memory row H-1 = raster row 0) untouched on every pipe (fbo.cpp:336): same in both
    assert np.array_equal(composite, whole["colour"])
    assert shaded == whole["stats"]["fragments_shaded"]
