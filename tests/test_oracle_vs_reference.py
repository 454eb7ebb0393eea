"""The parity PIN: oracle/ps3d_oracle.c against the reference's own renderer (oracle/_ref, built from the unmodified
sources through oracle/ref_shim). Same host => the x86 approximations are the same instructions on both sides, so
everything must be bit-exact: colour words, depth words, per-pixel FragmentProcessor::process counts, counters."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import SMALL


@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_matches_reference_bit_exact(name, oracle_lib, ref_lib):
    sc = SMALL[name]()
    a = render_all(ref_lib, sc)
    b = render_all(oracle_lib, sc)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)), "depth"
    assert np.array_equal(a["counts"], b["counts"]), "per-pixel shade counts (depth-test survivors)"
    assert np.array_equal(a["colour"], b["colour"]), "colour"
    for key in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key
    assert b["stats"]["fragments_shaded"] > 0


def test_reference_thread_count_does_not_change_the_image(oracle_lib, ref_lib):
    """Row interleave over N-1 workers (drawvao.cpp:78-85) must not change a pixel; the oracle is single-threaded."""
    sc = SMALL["soup_def02"]()
    a = render_all(ref_lib, sc)
    b = render_all(oracle_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])
