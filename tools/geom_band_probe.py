"""Kernel-class times of C2 against the height of the sort-first band (one GPU): what is fixed cost, what scales."""
import sys
sys.path.insert(0, '.')
from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline
sc = scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048)
pipe = PuresoftPipeline(sc.width, sc.height, device=0)
up = scenes.upload(pipe, sc)
for band in [(-1, -1), (0, 544), (272, 544), (0, 272), (0, 144), (0, 16)]:
    pipe.setRowBand(*band)
    for _ in range(5):
        scenes.replay(pipe, sc, up, finish=False)
    pipe.finish()
    pipe.profileEnable(True)
    for _ in range(50):
        scenes.replay(pipe, sc, up, finish=False)
    pipe.finish()
    pr = pipe.profileRead()
    pipe.profileEnable(False)
    print(band, {k: round(pr[k] / 50, 4) for k in ('geom_ms', 'bin_ms', 'tile_ms', 'shade_ms')})
