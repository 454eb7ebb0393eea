"""Kernel-class times of one full-size config (python tools/config_probe.py C3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline
sys.path.insert(0, os.path.join(ROOT, "tools"))
from measure_configs import CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
sc = CONFIGS[name]()
pipe = PuresoftPipeline(sc.width, sc.height, device=0)
up = scenes.upload(pipe, sc)
frame = scenes.compile_replay(pipe, sc, up)
for _ in range(3):
    frame()
pipe.finish()
pipe.resetStats()
pipe.profileEnable(True)
n = 10
for _ in range(n):
    frame()
pipe.finish()
pr = pipe.profileRead()
st = pipe.getStats()
print(name, sc.name, "draws/frame", st["draws"] / n, {k: round(pr[k] / n, 3) for k in ("geom_ms", "bin_ms", "tile_ms", "shade_ms")},
      "pairs/frame", pr["bin_pairs"] / n, "launches", {k: pr[k] / n for k in pr if k.endswith("launches")})
