#!/usr/bin/env python3
"""Do two builds of the library contain the same device code? Compares, kernel by kernel, the SASS instruction streams of
two .so files (addresses and encodings ignored). Used to show that a host-side change or a refactor (e.g. moving blend4 into
shaders.cuh so that it can be pinned on the host) left every kernel untouched, when no GPU is at hand to re-run the tests.

    python tools/sass_diff.py old/libps3d_b200.so puresoft3d_b200/libps3d_b200.so
"""
import re
import subprocess
import sys


def kernels(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, cur = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur:
            out[cur].append(m.group(1).strip())
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    only = sorted(set(a) ^ set(b))
    diff = sorted(k for k in a if k in b and a[k] != b[k])
    print("%d / %d kernels; only in one build: %d; different instruction streams: %d" % (len(a), len(b), len(only), len(diff)))
    for k in only[:20] + diff[:20]:
        print("  ", k[:120], len(a.get(k, [])), len(b.get(k, [])))
    return 1 if (only or diff) else 0


if __name__ == "__main__":
    sys.exit(main())
