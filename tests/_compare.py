"""Shared comparison helpers for the parity tests."""
import numpy as np

from puresoft3d_b200 import scenes
from puresoft3d_b200.pipeline import PuresoftPipeline


def render_all(lib, scene, capture=True):
    """Render `scene` on `lib`; returns dict(colour, depth, counts, stats)."""
    p = PuresoftPipeline(scene.width, scene.height, lib=lib)
    try:
        if capture:
            p.debugCapture(scene.width, scene.height)
        scenes.render(p, scene)
        out = dict(colour=p.readColour(), depth=p.readDepth(), stats=p.getStats())
        out["counts"] = p.debugReadShadeCounts() if capture else None
        return out
    finally:
        p.close()


def ulp_diff(a, b):
    """Elementwise distance in units in the last place between two float32 arrays (same-sign finite values)."""
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def colour_stats(a, b):
    """(fraction of pixels with every channel within 1/255, max abs channel difference)."""
    ca = a.view(np.uint8).reshape(a.shape + (4,)).astype(np.int16)
    cb = b.view(np.uint8).reshape(b.shape + (4,)).astype(np.int16)
    d = np.abs(ca - cb).max(axis=-1)
    return float((d <= 1).mean()), int(d.max())
