"""One sort-first band of C2 on one GPU, a few frames: run under `ncu --metrics gpu__time_duration.sum` for per-kernel times.
    python tools/band_launches.py <row0> <row1> [frames]"""
import sys
sys.path.insert(0, '.')
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline
r0, r1 = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sc = scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048)
pipe = PuresoftPipeline(sc.width, sc.height, device=0)
up = scenes.upload(pipe, sc)
pipe.setRowBand(r0, r1)
for _ in range(frames):
    scenes.replay(pipe, sc, up, finish=False)
pipe.finish()
