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


SMALL_DRAW_CASES = {
    "one_triangle_draws": dict(width=200, height=120, seed=32, draws=40, tris=1),
    "lists_outgrow_speculation": dict(width=320, height=200, seed=33, draws=12, tris=30, crowd=60),
    "big_draw_in_between": dict(width=320, height=200, seed=35, draws=9, tris=25, big_every=4),
    "odd_size": dict(width=238, height=131, seed=36, draws=17, tris=33),   # (W % 4 == 1 is the reference's own depth-pitch overrun: DESIGN.md, divergences)
}


@pytest.mark.parametrize("name", sorted(SMALL_DRAW_CASES))
def test_many_small_draws_oracle_matches_reference(name, oracle_lib, ref_lib):
    """The frames tests/test_batch_gpu.py renders in batches on the GPU (there with the parity-test functor FLATID in the mix, which
    the reference build does not have): the same geometry, uniforms and draw order with DEF02 in its place, oracle == reference."""
    from puresoft3d_b200 import scenes
    sc = scenes.scene_small_draws(flatid=False, **SMALL_DRAW_CASES[name])
    a = render_all(ref_lib, sc)
    b = render_all(oracle_lib, sc)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)), "depth"
    assert np.array_equal(a["counts"], b["counts"]), "per-pixel shade counts"
    assert np.array_equal(a["colour"], b["colour"]), "colour"
    for key in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded"):
        assert a["stats"][key] == b["stats"][key], key
