"""Render one scene on the CUDA library in a fresh process and print SHA-256 of colour / depth / shade counts + counters.
Used by test_properties_gpu.py to compare the tile paths and binning paths, which are selected by environment variables
read once per process (PS3D_TILE_PATH, PS3D_BINNING)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from puresoft3d_b200 import scenes
    from puresoft3d_b200.pipeline import PuresoftPipeline
    kind = sys.argv[1]
    if kind == "c2_full":
        sc = scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048)
    elif kind == "c2_mid":
        sc = scenes.scene_heightfield(1920, 1080, grid=120, layers=4, seed=2, tex_size=512)
    elif kind == "c4_full":
        sc = scenes.scene_blend_overdraw(1920, 1080)
    else:
        from _scenes_small import SMALL
        sc = SMALL[kind]()
    p = PuresoftPipeline(sc.width, sc.height)
    p.debugCapture(sc.width, sc.height)
    scenes.render(p, sc)
    colour, depth, counts, stats = p.readColour(), p.readDepth(), p.debugReadShadeCounts(), p.getStats()
    p.close()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    print(json.dumps({"colour": sha(colour), "depth": sha(depth.view(np.uint32)), "counts": sha(counts), "stats": stats,
                      "counts_sum": int(counts.astype(np.uint64).sum()), "covered": int((counts > 0).sum())}))


if __name__ == "__main__":
    main()
