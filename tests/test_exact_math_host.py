"""The product's arithmetic contract compiled for the HOST (puresoft3d_b200/csrc/exact_math.cuh is host/device code) against the
x86 instructions it stands for, on this CPU: rcpps / rsqrtss table emulation with the tables x86_approx.cpp measures (40 M
random bit patterns + the special values), the cvttss2si conversions, and the operation order of hsum4 / m4v4 / opt_pow.
No GPU, no oracle: product sources only."""
import os
import subprocess

from conftest import ROOT


def test_emulation_equals_the_hardware_instructions(tmp_path):
    exe = str(tmp_path / "approx_host_test")
    csrc = os.path.join(ROOT, "puresoft3d_b200", "csrc")
    subprocess.run(["g++", "-std=c++14", "-O2", "-msse4.1", "-mfpmath=sse", "-ffp-contract=off", "-x", "c++", "-I", csrc,
                    os.path.join(ROOT, "tests", "cpp", "approx_host_test.cpp"), os.path.join(csrc, "x86_approx.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in r.stdout.strip().splitlines())
    assert r.returncode == 0, r.stdout
    assert lines["tables"][0] == "1", "this CPU's rcpps / rsqrtss are not table machines of the expected shape: " + r.stdout
    for name in ("x86_rcp", "x86_rsqrt", "cvtt", "cvtu", "hsum4_m4v4_optpow"):
        checked, bad = int(lines[name][0]), int(lines[name][1])
        assert checked > 1000000 and bad == 0, (name, checked, bad)
