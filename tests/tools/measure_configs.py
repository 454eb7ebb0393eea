#!/usr/bin/env python3
"""every BASELINE.json config at its full size on one B200, beside the reference's own renderer on the
same box's host cores, with the parity gates evaluated on those very frames (the reference build is the checker here:
this is measurement + test tooling — it lives under tests/ because it loads the checkers under oracle/ — not the product path).

    python tests/tools/measure_configs.py [--configs C1,C2,C3,C4,C5-4k] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from puresoft3d_b200 import _capi, scenes  # noqa: E402
from puresoft3d_b200.pipeline import PuresoftPipeline  # noqa: E402
from _compare import colour_stats, ulp_diff  # noqa: E402

CONFIGS = {
    "C1": lambda: scenes.scene_cube(640, 480),
    "C2": lambda: scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048),
    "C3": lambda: scenes.scene_desk(1920, 1080, shadow=4096, clutter=24, tex_size=512),
    "C4": lambda: scenes.scene_blend_overdraw(1920, 1080),
    "C5-4k": lambda: scenes.scene_heightfield(3840, 2160, grid=1118, layers=4, seed=5, tex_size=2048),
    "C5-8k": lambda: scenes.scene_heightfield(7680, 4320, grid=1118, layers=4, seed=5, tex_size=2048),
}


def frame_bytes(sc, frags):
    vertex = sc.vertex_bytes_read()
    targets = 2 * sc.width * sc.height * 4 + sum(4 * t["width"] * t["height"] for t in sc.textures if t["layers"][0] is None)
    tex = sum(min(sum(a.nbytes for a in t["layers"]), 4 * 2 * frags) for t in sc.textures if t["layers"][0] is not None)
    return vertex + targets + tex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--ref-frames", type=int, default=2)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    ref = _capi.bind(os.path.join(ROOT, "oracle", "_ref", "libps3d_ref.so"))
    cores = os.cpu_count() or 1
    os.environ.setdefault("PS3D_REF_THREADS", str(cores))
    dev = torch.device("cuda", 0)
    for name in args.configs.split(","):
        sc = CONFIGS[name]()
        # ---- the CUDA library: frames per second with resident inputs, then one captured frame for the parity gates
        pipe = PuresoftPipeline(sc.width, sc.height, device=0)
        up = scenes.upload(pipe, sc)
        frame = scenes.compile_replay(pipe, sc, up)
        ext = torch.cuda.ExternalStream(pipe.deviceStream(), device=dev)
        for _ in range(5):
            frame()
        pipe.finish()
        if os.environ.get("PS3D_GRAPH", "1") != "0":
            # the frame as one launch (ps3d_graph_*): what 54 draws of 12-triangle boxes (C3) cost is mostly launches
            calls = frame
            pipe.graphBegin()
            calls()
            g = pipe.graphEnd()
            frame = (lambda pipe=pipe, g=g: pipe.graphLaunch(g))
            for _ in range(3):
                frame()
            pipe.finish()
        pipe.resetStats()
        steps = 50 if not name.startswith("C5") else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            frame()
        e1.record(ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        st = pipe.getStats()
        frags = st["fragments_shaded"] / steps
        # the frames that were timed (warm speculation, chain marks, batches, captured): the last one must be the frame a fresh pipe renders
        steady = dict(colour=pipe.readColour(), depth=pipe.readDepth())
        pipe.close()
        # the parity frame on a FRESH pipe, like the reference's below: clear4 never touches the last buffer row (fbo.cpp:336),
        # so on a long-lived target a blending scene accumulates there from frame to frame — on both sides
        pipe = PuresoftPipeline(sc.width, sc.height, device=0)
        pipe.debugCapture(sc.width, sc.height)
        scenes.render(pipe, sc)
        g = dict(colour=pipe.readColour(), depth=pipe.readDepth(), counts=pipe.debugReadShadeCounts())
        pipe.close()
        # ---- the reference's own renderer: counted frame (parity), then timed frames with the counting decorators off
        os.environ["PS3D_REF_COUNTING"] = "1"
        q = PuresoftPipeline(sc.width, sc.height, lib=ref)
        q.debugCapture(sc.width, sc.height)
        upq = scenes.upload(q, sc)
        scenes.replay(q, sc, upq)
        r = dict(colour=q.readColour(), depth=q.readDepth(), counts=q.debugReadShadeCounts(), stats=q.getStats())
        q.close()
        os.environ["PS3D_REF_COUNTING"] = "0"
        q = PuresoftPipeline(sc.width, sc.height, lib=ref)
        upq = scenes.upload(q, sc)
        scenes.replay(q, sc, upq)
        t0 = time.perf_counter()
        for _ in range(args.ref_frames):
            scenes.replay(q, sc, upq)
            q.swapBuffers()
        cpu_ms = (time.perf_counter() - t0) * 1000.0 / args.ref_frames
        q.close()
        frac, maxd = colour_stats(g["colour"], r["colour"])
        both = np.isfinite(g["depth"]) & np.isfinite(r["depth"])
        ulp = int(ulp_diff(g["depth"][both], r["depth"][both]).max()) if both.any() else 0
        balg = frame_bytes(sc, int(frags))
        line = {
            "config": name, "scene": sc.name, "triangles": sc.meta.get("triangles"), "resolution": [sc.width, sc.height],
            "fragments_per_frame": frags, "algorithmic_bytes": balg,
            "gpu_ms_per_frame": ms, "gpu_frames_per_s": 1000.0 / ms, "gpu_fragments_per_s": frags / (ms / 1e3),
            "roofline_frac": (balg / 1e9) / (ms / 1e3) / peak,
            "cpu_ms_per_frame": cpu_ms, "cpu_frames_per_s": 1000.0 / cpu_ms, "cpu_fragments_per_s": frags / (cpu_ms / 1e3), "cpu_threads": cores,
            "parity": {"survivor_counts_bit_exact": bool(np.array_equal(g["counts"], r["counts"])),
                       "coverage_bit_exact": bool(np.array_equal(g["counts"] > 0, r["counts"] > 0)),
                       "depth_bit_exact": bool(np.array_equal(g["depth"].view(np.uint32), r["depth"].view(np.uint32))), "depth_max_ulp": ulp,
                       "colour_within_1_of_255": frac, "colour_max_diff": maxd,
                       "fragments_shaded_equal": bool(int(frags) == int(r["stats"]["fragments_shaded"])),
                       # (clear4 leaves the last buffer row alone, fbo.cpp:336: a blending scene accumulates there on a long-lived pipe)
                       "timed_frames_equal_first_frame": bool(np.array_equal(steady["depth"].view(np.uint32), g["depth"].view(np.uint32))
                                                              and np.array_equal(steady["colour"].view(np.uint32)[:-1], g["colour"].view(np.uint32)[:-1]))},
        }
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
