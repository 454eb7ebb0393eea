"""The post-processing hook: PuresoftPipeline::postProcess (post.cpp:3-19) with the reference's only
post-processor, PP_DepthofField (src/test2/testpost.cpp:9-43). Even widths are in the pinned scene set (demo2_post: oracle
== reference bit for bit, golden fixture, GPU parity). This file covers the odd-width case, where the reference races."""
import numpy as np
import pytest

from _compare import render_all
from _scenes_small import RACY_IN_REFERENCE, SMALL
from puresoft3d_b200 import _capi as K
from puresoft3d_b200.pipeline import PP_DepthofField, PuresoftPipeline


def test_post_processor_adds_50_to_every_byte_wrapping(oracle_lib):
    plain = render_all(oracle_lib, SMALL["demo2_objx_file"]())["colour"].view(np.uint8)
    post = render_all(oracle_lib, SMALL["demo2_post"]())["colour"].view(np.uint8)
    assert np.array_equal(post, plain + np.uint8(50))          # uint8 arithmetic wraps like paddb
    assert (plain.astype(np.int32) + 50 > 255).any()           # ... and some bytes really wrapped


def test_unknown_post_processor_is_refused(oracle_lib):
    p = PuresoftPipeline(32, 16, lib=oracle_lib)
    class Nope(PP_DepthofField):
        functor = 77
    with pytest.raises(Exception):
        p.postProcess(Nope())
    p.close()


@pytest.mark.parametrize("name", sorted(RACY_IN_REFERENCE))
def test_odd_width_matches_reference_outside_the_racy_column(name, oracle_lib, ref_lib):
    sc = RACY_IN_REFERENCE[name]()
    assert sc.width % 2 == 1
    a = render_all(oracle_lib, sc)
    for _ in range(3):
        b = render_all(ref_lib, sc)
        assert np.array_equal(a["colour"][:, 1:], b["colour"][:, 1:])
        assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
        # column 0: the sequential result is +100 (own row's step + the previous row's last step); the reference lands on
        # +100 or, when the two workers' read-modify-writes collide, +50
        seq = np.ascontiguousarray(a["colour"][:, 0]).view(np.uint8).reshape(-1, 4)
        got = np.ascontiguousarray(b["colour"][:, 0]).view(np.uint8).reshape(-1, 4)
        lost = seq - np.uint8(50)
        assert np.all((got == seq).all(axis=1) | (got == lost).all(axis=1))
    assert np.array_equal(a["colour"][0, 0], a["colour"][0, 1])   # memory row 0 has no row before it: +50 only
    assert a["colour"][1, 0] != a["colour"][1, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RACY_IN_REFERENCE))
def test_cuda_equals_oracle_on_odd_width(name, cuda_lib, oracle_lib):
    sc = RACY_IN_REFERENCE[name]()
    a, b = render_all(cuda_lib, sc), render_all(oracle_lib, sc)
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    d = np.abs(a["colour"].view(np.uint8).astype(np.int16) - b["colour"].view(np.uint8).astype(np.int16))
    d = np.minimum(d, 256 - d)                                  # the +50 wraps
    assert (d.reshape(a["colour"].shape + (4,)).max(-1) <= 1).mean() >= 0.999
