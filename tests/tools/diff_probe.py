"""Where does the CUDA frame differ from the oracle's? (debug aid: prints differing pixels of a full-size config)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from puresoft3d_b200 import _capi, scenes
from _compare import render_all

which = sys.argv[1] if len(sys.argv) > 1 else "C4"
sc = {"C4": lambda: scenes.scene_blend_overdraw(1920, 1080), "C4q": lambda: scenes.scene_blend_overdraw(1920, 1080, randoms=0),
      "C4s": lambda: scenes.scene_blend_overdraw(960, 540)}[which]()
oracle = _capi.bind(os.path.join(ROOT, "oracle", "libps3d_oracle.so"))
a = render_all(_capi.load_product(), sc)
b = render_all(oracle, sc)
d = a["colour"] != b["colour"]
print(which, "differing pixels", int(d.sum()), "of", d.size, "depth eq", np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)),
      "counts eq", np.array_equal(a["counts"], b["counts"]))
ys, xs = np.nonzero(d)
if len(ys):
    cnt = a["counts"][sc.height - 1 - ys, xs]          # counts are in raster rows (bottom-up), colour in memory rows (top-down)
    print("survivor counts at differing pixels: histogram", np.bincount(cnt)[:24])
    print("all pixels histogram                          ", np.bincount(a["counts"].ravel())[:24])
    print("rows", np.unique(ys)[:20], "... cols", np.unique(xs)[:20], "tile x", np.unique(xs // 16)[:30], "tile y", np.unique((sc.height - 1 - ys) // 16)[:30])
    for k in range(min(12, len(ys))):
        print(ys[k], xs[k], "raster row", sc.height - 1 - ys[k], "cuda %08x oracle %08x count %d" % (a["colour"][ys[k], xs[k]], b["colour"][ys[k], xs[k]], cnt[k]))
