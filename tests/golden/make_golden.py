#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libps3d_ref.so = the unmodified Puresoft3D sources
built through oracle/ref_shim; only possible where /root/reference exists). Run from the repo root:

    python tests/golden/make_golden.py

For every scene of tests/_scenes_small.py: SHA-256 of the depth words and of the per-pixel FragmentProcessor::process
counts (bit-exact gates: they never depend on the host's rcpps/rsqrtss), the counters, and the colour image itself
(zlib-compressed) for the tolerance gate. The reference ships no golden vectors of its own."""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from puresoft3d_b200 import _capi  # noqa: E402
from _compare import render_all  # noqa: E402
from _scenes_small import SMALL  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = _capi.bind(os.path.join(ROOT, "oracle", "_ref", "libps3d_ref.so"))
    assert ref.ps3d_backend_name() == b"reference-shim"
    index = {}
    for name in sorted(SMALL):
        sc = SMALL[name]()
        out = render_all(ref, sc)
        blob = zlib.compress(np.ascontiguousarray(out["colour"]).tobytes(), 9)
        with open(os.path.join(ROOT, "tests", "golden", name + ".colour.z"), "wb") as f:
            f.write(blob)
        index[name] = {
            "width": sc.width, "height": sc.height,
            "depth_sha256": sha(out["depth"].view(np.uint32)),
            "counts_sha256": sha(out["counts"]),
            "colour_sha256": sha(out["colour"]),
            "stats": {k: out["stats"][k] for k in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded")},
            "colour_file": name + ".colour.z", "colour_bytes": len(blob),
        }
        print(name, index[name]["stats"], len(blob))
    with open(os.path.join(ROOT, "tests", "golden", "index.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libps3d_ref.so (reference sources, shim build, this container's CPU)",
                   "scenes": index}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
