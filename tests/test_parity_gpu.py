"""GPU parity: the CUDA library against the CPU oracle on identical seeded scenes, through the C-ABI.
Gates (BASELINE.json north_star): coverage mask and depth-test survivors bit-exact (per-pixel shade counts), depth
bit-exact (gate: <= 2 ulp), colour within 1/255 per channel on >= 99.9 % of pixels."""
import numpy as np
import pytest

from _compare import colour_stats, render_all, ulp_diff
from _scenes_small import SMALL

pytestmark = pytest.mark.gpu

COLOUR_FRACTION = 0.999  # of pixels with every 8-bit channel within 1 of the reference
DEPTH_ULP = 2


def check(a, b):
    assert np.array_equal(a["counts"], b["counts"]), "depth-test survivor set / coverage differs"
    assert int(ulp_diff(a["depth"], b["depth"]).max()) <= DEPTH_ULP, "depth beyond 2 ulp"
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)), "depth not bit-exact"
    frac, worst = colour_stats(a["colour"], b["colour"])
    assert frac >= COLOUR_FRACTION, "colour within 1/255 on only %.5f of pixels (max diff %d)" % (frac, worst)
    for key in ("triangles_submitted", "triangles_rasterised", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key


@pytest.mark.parametrize("name", sorted(SMALL))
def test_cuda_matches_oracle(name, cuda_lib, oracle_lib):
    sc = SMALL[name]()
    check(render_all(cuda_lib, sc), render_all(oracle_lib, sc))


@pytest.mark.parametrize("name", ["c1_cube_def01", "c2_heightfield_small", "c4_blend_overdraw"])
def test_cuda_matches_reference_build(name, cuda_lib, ref_lib):
    """Directly against the reference's own renderer (oracle/_ref) where the prebuilt library travelled."""
    sc = SMALL[name]()
    a, b = render_all(cuda_lib, sc), render_all(ref_lib, sc)
    assert np.array_equal(a["counts"], b["counts"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    frac, worst = colour_stats(a["colour"], b["colour"])
    assert frac >= COLOUR_FRACTION, (frac, worst)


def test_render_is_idempotent(cuda_lib):
    sc = SMALL["c2_heightfield_small"]()
    a, b = render_all(cuda_lib, sc), render_all(cuda_lib, sc)
    assert np.array_equal(a["colour"], b["colour"]) and np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
