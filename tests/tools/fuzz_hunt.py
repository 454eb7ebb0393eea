"""Bug hunt beyond the test-suite's seeds: seeded variants of every scene family, CUDA against the restatement (on a GPU box),
or — `--cpu` — the restatement against the reference build (in the build container).

    python tests/tools/fuzz_hunt.py [--cpu] [--bands] [first_seed] [count]

--bands: every frame restricted to one rank's sort-first band of a random world size (ps3d_set_row_band) on both sides.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from puresoft3d_b200 import _capi, scenes
from _compare import colour_stats, render_all


def even(v):
    v = int(v)
    return v + 1 if v % 4 == 1 else v      # pipeline.cpp:31 under-allocates depth rows when W % 4 == 1


def family(seed):
    rng = np.random.default_rng(5000 + seed)
    w, h = even(rng.integers(48, 260)), int(rng.integers(40, 180))
    k = seed % 7
    if k == 0:
        return scenes.scene_state_fuzz(seed)
    if k == 1:
        return scenes.scene_desk(w, h, shadow=int(rng.choice([64, 96, 160])), seed=seed, clutter=int(rng.integers(1, 12)), tex_size=int(rng.choice([16, 64])), skybox=bool(rng.integers(0, 2)))
    if k == 2:
        return scenes.scene_planets(w, h, shadow=int(rng.choice([64, 120])), seed=seed, stacks=int(rng.integers(4, 12)), slices=int(rng.integers(6, 20)), tex_size=int(rng.choice([32, 64])))
    if k == 3:
        return scenes.scene_blend_overdraw(w, h, seed=seed, quads=int(rng.integers(1, 12)), randoms=int(rng.integers(0, 300)))
    if k == 4:
        return scenes.scene_crowded_tile(even(rng.integers(64, 200)), int(rng.integers(48, 150)), seed=seed, crowd=int(rng.choice([50, 700, 2300])))
    if k == 5:
        return scenes.scene_heightfield(w, h, grid=int(rng.integers(4, 60)), layers=int(rng.integers(1, 4)), seed=seed, tex_size=int(rng.choice([32, 128])),
                                        functor=int(rng.choice([_capi.FN_DEF01, _capi.FN_DEF03])))
    return scenes.scene_soup(w, h, seed=seed, count=int(rng.integers(10, 500)), cull=bool(rng.integers(0, 2)))


def render_band(lib, sc, band):
    from puresoft3d_b200.pipeline import PuresoftPipeline
    p = PuresoftPipeline(sc.width, sc.height, lib=lib)
    try:
        p.setRowBand(*band)
        scenes.render(p, sc)
        return dict(colour=p.readColour(), depth=p.readDepth(), stats=p.getStats())
    finally:
        p.close()


def main_bands(first, count):
    from puresoft3d_b200 import sortfirst
    oracle = _capi.bind(os.path.join(ROOT, "oracle", "libps3d_oracle.so"))
    cuda = _capi.load_product()
    bad = []
    for seed in range(first, first + count):
        rng = np.random.default_rng(9000 + seed)
        sc = family(seed)
        world = int(rng.integers(2, 10))
        bands = sortfirst.row_bands(sc.height, world)
        band = bands[int(rng.integers(0, world))]
        a, b = render_band(cuda, sc, band), render_band(oracle, sc, band)
        m0, m1 = sortfirst.memory_rows(band, sc.height)
        ok = (np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)),
              colour_stats(a["colour"][m0:m1], b["colour"][m0:m1])[0] >= 0.999 if m1 > m0 else True,
              all(a["stats"][k] == b["stats"][k] for k in ("fragments_tested", "fragments_shaded")))
        if not all(ok):
            bad.append(seed)
            print(seed, sc.name, "world", world, "band", band, "depth/colour/stats", ok, flush=True)
    print("cuda vs restatement with a sort-first band, seeds %d..%d: mismatching %s" % (first, first + count - 1, bad))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cpu = "--cpu" in sys.argv
    if "--bands" in sys.argv:
        return main_bands(int(args[0]) if args else 0, int(args[1]) if len(args) > 1 else 150)
    first, count = (int(args[0]) if args else 0), (int(args[1]) if len(args) > 1 else 200)
    oracle = _capi.bind(os.path.join(ROOT, "oracle", "libps3d_oracle.so"))
    other = _capi.bind(os.path.join(ROOT, "oracle", "_ref", "libps3d_ref.so")) if cpu else _capi.load_product()
    bad = []
    for seed in range(first, first + count):
        sc = family(seed)
        a, b = render_all(other, sc), render_all(oracle, sc)
        ok = (np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)), np.array_equal(a["counts"], b["counts"]),
              np.array_equal(a["colour"], b["colour"]) if cpu else colour_stats(a["colour"], b["colour"])[0] >= 0.999,
              all(a["stats"][k] == b["stats"][k] for k in ("triangles_submitted", "spans", "fragments_tested", "fragments_shaded")))
        if not all(ok):
            bad.append(seed)
            print(seed, sc.name, "depth/counts/colour/stats", ok, flush=True)
    print("%s vs restatement, seeds %d..%d: mismatching %s" % ("reference" if cpu else "cuda", first, first + count - 1, bad))


if __name__ == "__main__":
    main()
