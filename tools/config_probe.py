"""Kernel-class times of one full-size config (python tools/config_probe.py C3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline
CONFIGS = {   # BASELINE.json configs at full size (the same table as tests/tools/measure_configs.py)
    "C1": lambda: scenes.scene_cube(640, 480),
    "C2": lambda: scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048),
    "C3": lambda: scenes.scene_desk(1920, 1080, shadow=4096, clutter=24, tex_size=512),
    "C4": lambda: scenes.scene_blend_overdraw(1920, 1080),
    "C5-4k": lambda: scenes.scene_heightfield(3840, 2160, grid=1118, layers=4, seed=5, tex_size=2048),
}

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
