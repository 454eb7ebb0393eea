"""A few frames of one BASELINE.json config on one GPU: run under `ncu --metrics gpu__time_duration.sum` for per-kernel times.
    python tools/config_launches.py C3 [frames]"""
import sys
sys.path.insert(0, '.')
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline
name = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = {"C1": lambda: scenes.scene_cube(640, 480),
      "C2": lambda: scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048),
      "C3": lambda: scenes.scene_desk(1920, 1080, shadow=4096, clutter=24, tex_size=512),
      "C4": lambda: scenes.scene_blend_overdraw(1920, 1080)}[name]()
pipe = PuresoftPipeline(sc.width, sc.height, device=0)
up = scenes.upload(pipe, sc)
for _ in range(frames):
    scenes.replay(pipe, sc, up, finish=False)
pipe.finish()
