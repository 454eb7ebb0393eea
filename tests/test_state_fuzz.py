"""Seeded frames of several draws with the pipeline's state changing between them (scenes.scene_state_fuzz): behaviour bits,
viewport sizes, the five default programmes' vertex layouts, CLAMP / WRAP textures, cube maps, triangles of every size with
perspective w. CPU: the restatement against the unmodified reference build, bit for bit (widens the oracle's pin beyond the
hand-written scenes). GPU: the CUDA library against the restatement through the C-ABI."""
import os
import sys

import numpy as np
import pytest

from _compare import colour_stats, render_all
from conftest import ROOT
from puresoft3d_b200 import scenes

sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from fuzz_hunt import family, render_band  # noqa: E402  (seeded variants of every scene family: demos, blend overdraw, crowded tile, height fields, soups)


@pytest.mark.parametrize("seed", range(120))
def test_oracle_equals_reference(seed, oracle_lib, ref_lib):
    sc = scenes.scene_state_fuzz(seed)
    a, b = render_all(oracle_lib, sc), render_all(ref_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    for key in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(48))
def test_cuda_equals_oracle(seed, cuda_lib, oracle_lib):
    sc = scenes.scene_state_fuzz(seed)
    a, b = render_all(cuda_lib, sc), render_all(oracle_lib, sc)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    frac, _ = colour_stats(a["colour"], b["colour"])
    assert frac >= 0.999
    for key in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key


@pytest.mark.parametrize("seed", range(70))
def test_scene_families_oracle_equals_reference(seed, oracle_lib, ref_lib):
    sc = family(seed)
    a, b = render_all(oracle_lib, sc), render_all(ref_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(42))
def test_scene_families_cuda_equals_oracle(seed, cuda_lib, oracle_lib):
    sc = family(seed)
    a, b = render_all(cuda_lib, sc), render_all(oracle_lib, sc)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    assert colour_stats(a["colour"], b["colour"])[0] >= 0.999
    for key in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(36))
def test_scene_families_with_a_sort_first_band(seed, cuda_lib, oracle_lib):
    """One rank's band of a random world size (2..9): small bands take the two-kernel geometry and the row-group split."""
    from puresoft3d_b200 import sortfirst
    rng = np.random.default_rng(9000 + seed)
    sc = family(seed)
    world = int(rng.integers(2, 10))
    band = sortfirst.row_bands(sc.height, world)[int(rng.integers(0, world))]
    a, b = render_band(cuda_lib, sc, band), render_band(oracle_lib, sc, band)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    m0, m1 = sortfirst.memory_rows(band, sc.height)
    if m1 > m0:
        assert colour_stats(a["colour"][m0:m1], b["colour"][m0:m1])[0] >= 0.999
    assert a["stats"]["fragments_tested"] == b["stats"]["fragments_tested"] and a["stats"]["fragments_shaded"] == b["stats"]["fragments_shaded"]
