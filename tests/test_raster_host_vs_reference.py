"""puresoft3d_b200/csrc/raster.cuh — the product's triangle set-up, per-row spans and edge weights, host/device code — compiled
for the HOST and compared, triangle by triangle, with the reference's own PuresoftRasterizer::pushTriangle and
PuresoftInterpolater::lineSegmentlinearInterpolate (class from oracle/_ref/libps3d_ref.so, header read in place from the
reference tree): 200 000 random triangles of every kind on five target sizes, ~60 M RESULT_ROWs, bit for bit. This pins the
coverage arithmetic below the level of frames. Only where /root/reference exists."""
import os
import subprocess

import pytest

from conftest import REF_SO, ROOT

REF_INCLUDE = "/root/reference/src/puresoft3d"


@pytest.mark.skipif(not os.path.isdir(REF_INCLUDE), reason="the reference tree is only present in the build container")
def test_raster_cuh_equals_the_reference_rasterizer(tmp_path, built):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libps3d_ref.so not built")
    exe = str(tmp_path / "raster_vs_reference")
    refdir = os.path.dirname(REF_SO)
    subprocess.run(["g++", "-std=c++14", "-O1", "-msse4.1", "-mfpmath=sse", "-ffp-contract=off", "-x", "c++",
                    "-I", os.path.join(ROOT, "puresoft3d_b200", "csrc"), "-I", REF_INCLUDE, os.path.join(ROOT, "tests", "cpp", "raster_vs_reference.cpp"),
                    "-x", "none", "-L", refdir, "-lps3d_ref", "-Wl,-rpath," + refdir, "-o", exe], check=True)
    r = subprocess.run([exe, "200000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = dict((ln.split()[0], (int(ln.split()[1]), int(ln.split()[2]))) for ln in r.stdout.strip().splitlines())
    assert r.returncode == 0, r.stdout
    assert lines["return_code_rows_vertices"] == (200000, 0)
    for name in ("rows_written", "result_rows", "edge_weights"):
        assert lines[name][0] > 10 ** 7 and lines[name][1] == 0, (name, lines[name])
