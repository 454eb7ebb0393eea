"""tests/cpp/demo1_objx.cpp — demo 1 of the reference (src/test/puresoft.cpp:113-206) as a headless C++ caller: sphere from
an OBJX file through the native reader, frame through the C++ mirror with the demo's own class names. CPU: linked against
the oracle library, its geometry must be the Python-driven frame's (same spans and tested fragments; the pictures differ,
so the cloud layer's discards and the colours do not have to agree). GPU: CUDA-linked run == oracle-linked run."""
import os
import subprocess

import numpy as np
import pytest

from _compare import render_all
from conftest import ORACLE_SO, PRODUCT_SO, ROOT
from puresoft3d_b200 import objx, scenes

SRC = os.path.join(ROOT, "tests", "cpp", "demo1_objx.cpp")
W, H, S = 400, 250, 240


def write_sphere(path):
    pos, tan, _bin, nrm, uv = scenes.sphere_mesh(12, 24, radius=0.5)
    p = pos.copy()
    p[:, 3] = 0.0                       # findOrCreateVao sets w = 1 (scenobj.cpp:114-117)
    objx.write_objx(path, {}, [{"name": "sphere", "vertices": p, "normals": nrm, "tangents": tan, "texcoords": uv}])
    return path


def build_and_run(tmp_path, so, tag, objx_path):
    exe = str(tmp_path / ("demo1_objx_" + tag))
    libs = []
    for lib in ([so] if so == PRODUCT_SO else [so, PRODUCT_SO]):      # ps3d_objx_* lives in the product library (host code)
        d, n = os.path.dirname(lib), os.path.basename(lib)[3:-3]
        libs += ["-L", d, "-l" + n, "-Wl,-rpath," + d]
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe] + libs, check=True)
    r = subprocess.run([exe, str(objx_path), str(W), str(H), str(S)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return dict(ln.split(" ", 1) for ln in r.stdout.strip().splitlines())


def test_cpp_driver_has_the_python_driven_frames_geometry(tmp_path, oracle_lib, built):
    out = build_and_run(tmp_path, ORACLE_SO, "oracle", write_sphere(tmp_path / "sphere.objx"))
    assert out["backend"] == "oracle-c"
    want = render_all(oracle_lib, scenes.scene_planets(W, H, shadow=S, stacks=12, slices=24, tex_size=128))["stats"]
    submitted, spans, tested, shaded, draws = [int(v) for v in out["stats"].split()]
    assert (submitted, spans, tested, draws) == (want["triangles_submitted"], want["spans"], want["fragments_tested"], want["draws"])
    assert 0 < shaded <= tested
    assert int(out["covered"].split()[0]) > 0.1 * W * H and int(out["covered"].split()[2]) > 0


@pytest.mark.gpu
def test_cpp_driver_cuda_equals_oracle(tmp_path, built):
    path = write_sphere(tmp_path / "sphere.objx")
    a = build_and_run(tmp_path, PRODUCT_SO, "cuda", path)
    b = build_and_run(tmp_path, ORACLE_SO, "oracle", path)
    assert a["backend"] == "cuda-sm100a" and b["backend"] == "oracle-c"
    for key in ("stats", "covered", "depth", "shadow", "colour"):
        assert a[key] == b[key], key
