"""puresoft3d_b200/csrc/shaders.cuh — every vertex, interpolation and fragment functor of the product, with the three samplers —
compiled for the HOST and compared call by call with the reference's own processor classes (oracle/_ref/libps3d_ref.so,
driven through the ps3d_ref_* test hooks of oracle/ref_shim/ref_capi.cpp): random vertices, spans and interpolated varyings,
the same uniforms and texture bytes (2-D BGRA with CLAMP and WRAP, float shadow maps, cube maps), this CPU's rcpps / rsqrtss
tables; and blend4 over every (source, destination, alpha) of a channel against PuresoftFBO::blend4. Positions, varyings, span start / step / corrected fragment data, colour words and the wrote / discarded / blendable
flags must be bit-identical. Pins the shader arithmetic below the level of frames, on CPU."""
import os
import subprocess

import pytest

from conftest import REF_SO, ROOT


def test_functors_equal_the_reference_classes(tmp_path, built):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libps3d_ref.so not built here")
    exe = str(tmp_path / "functors_vs_reference")
    csrc, refdir = os.path.join(ROOT, "puresoft3d_b200", "csrc"), os.path.dirname(REF_SO)
    subprocess.run(["g++", "-std=c++14", "-O1", "-msse4.1", "-mfpmath=sse", "-ffp-contract=off", "-x", "c++", "-I", csrc,
                    os.path.join(ROOT, "tests", "cpp", "functors_vs_reference.cpp"), os.path.join(csrc, "x86_approx.cpp"),
                    "-x", "none", "-L", refdir, "-lps3d_ref", "-Wl,-rpath," + refdir, "-o", exe], check=True)
    r = subprocess.run([exe, "150000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert r.returncode == 0, r.stdout
    assert lines[0][:2] == ["tables", "1"], "this CPU's rcpps / rsqrtss are not table machines of the expected shape"
    results = {ln[0]: (int(ln[1]), int(ln[2])) for ln in lines[1:]}
    assert len([k for k in results if k.startswith("V_")]) == 12 and len([k for k in results if k.startswith("I_")]) == 9 and len([k for k in results if k.startswith("F_")]) == 15
    assert results["blend4"][0] > 3000000              # every (source, destination / 5, alpha) of a channel; all 16.7 M with n >= 1 M
    for name, (checked, bad) in results.items():
        assert checked >= 15000 and bad == 0, (name, checked, bad)
