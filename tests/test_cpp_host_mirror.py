"""include/puresoft3d_b200.hpp — the C++ host layer mirroring PuresoftPipeline / PuresoftVBO (pipeline.h:27-66,
vbo.h:16-21). tests/cpp/host_mirror_demo.cpp is written like the reference's demo code; it is linked against the oracle
library on CPU (host logic, exception mapping, ownership) and against libps3d_b200.so on the GPU box, where both runs
must print the same counters, colour hash and depth hash."""
import os
import subprocess

import pytest

from conftest import ORACLE_SO, PRODUCT_SO, ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp")


def build_and_run(tmp_path, so, tag):
    exe = str(tmp_path / ("host_mirror_" + tag))
    libdir, lib = os.path.dirname(so), os.path.basename(so)[3:-3]
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", libdir, "-l" + lib, "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    return dict(ln.split(" ", 1) for ln in r.stdout.strip().splitlines())


def test_host_mirror_against_oracle_library(tmp_path, built):
    out = build_and_run(tmp_path, ORACLE_SO, "oracle")
    assert out["backend"] == "oracle-c"
    assert out["errors"] == "ok" and out["ownership"] == "ok"
    assert int(out["stats"].split()[4]) > 1000
    assert out["post"] != out["colour"]


@pytest.mark.gpu
def test_host_mirror_cuda_equals_oracle(tmp_path, built):
    a = build_and_run(tmp_path, PRODUCT_SO, "cuda")
    b = build_and_run(tmp_path, ORACLE_SO, "oracle")
    assert a["backend"] == "cuda-sm100a"
    assert a["errors"] == "ok" and a["ownership"] == "ok"
    for key in ("stats", "depth", "colour", "post"):
        assert a[key] == b[key], key
