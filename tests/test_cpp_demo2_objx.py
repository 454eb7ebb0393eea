"""tests/cpp/demo2_objx.cpp — demo 2 of the reference (src/test2/puresoft.cpp, loadscene.cpp) as a headless C++ caller: OBJX
file through the native reader (include/ps3d_objx.h), frame through the C++ mirror (include/puresoft3d_b200.hpp). On CPU it
is linked against the oracle library and must reproduce the counters of the Python-driven frame of the same file (two host
languages, one scene); on the GPU box it is linked against libps3d_b200.so and must print the oracle-linked run's hashes."""
import os
import subprocess

import numpy as np
import pytest

from _compare import render_all
from conftest import ORACLE_SO, PRODUCT_SO, ROOT
from puresoft3d_b200 import scenes

SRC = os.path.join(ROOT, "tests", "cpp", "demo2_objx.cpp")
W, H, S = 384, 240, 256


def build_and_run(tmp_path, so, tag, objx_path):
    exe = str(tmp_path / ("demo2_objx_" + tag))
    libs = []
    for lib in ([so] if so == PRODUCT_SO else [so, PRODUCT_SO]):      # ps3d_objx_* lives in the product library (host code, no GPU needed);
        d, n = os.path.dirname(lib), os.path.basename(lib)[3:-3]       # the pipeline entry points resolve to the first library named
        libs += ["-L", d, "-l" + n, "-Wl,-rpath," + d]
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe] + libs, check=True)
    r = subprocess.run([exe, str(objx_path), str(W), str(H), str(S)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return dict(ln.split(" ", 1) for ln in r.stdout.strip().splitlines())


def test_cpp_driver_reproduces_the_python_driven_frame(tmp_path, oracle_lib, built):
    path = scenes.write_demo_objx(tmp_path / "demo.objx", seed=11, clutter=6)
    out = build_and_run(tmp_path, ORACLE_SO, "oracle", path)
    assert out["backend"] == "oracle-c"
    want = render_all(oracle_lib, scenes.scene_desk_objx(path, W, H, shadow=S, tex_size=64))
    st = want["stats"]
    # the C++ driver renders the frame twice (the second from the first one's targets): counters are exactly double; its
    # matrices are built by its own float code, so depth words may differ in the last bit while coverage and counts do not
    assert [int(v) for v in out["stats"].split()] == [2 * st["triangles_submitted"], 2 * st["spans"], 2 * st["fragments_tested"], 2 * st["fragments_shaded"]]
    assert int(out["covered"]) == int((want["depth"] < 1.0).sum())
    comps, draws = int(out["components"].split()[0]), int(out["components"].split()[2])
    assert draws == 2 * st["draws"] and comps == len(scenes.scene_desk_objx(path, W, H, shadow=S, tex_size=64).meta["components"])


@pytest.mark.gpu
def test_cpp_driver_cuda_equals_oracle(tmp_path, built):
    path = scenes.write_demo_objx(tmp_path / "demo.objx", seed=11, clutter=6)
    a = build_and_run(tmp_path, PRODUCT_SO, "cuda", path)
    b = build_and_run(tmp_path, ORACLE_SO, "oracle", path)
    assert a["backend"] == "cuda-sm100a" and b["backend"] == "oracle-c"
    for key in ("components", "stats", "covered", "depth", "shadow", "colour"):
        assert a[key] == b[key], key
