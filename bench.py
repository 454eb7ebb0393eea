#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "C2"): BASELINE.json configs[1] — 1080p synthetic 1M-triangle mesh (4 stacked 354x354
height-field grids, shuffled so submission order matters), DEF03 = per-pixel Blinn-Phong + tangent-space normal map,
two 2048^2 BGRA textures, depth test + write, back-face culling. Synthetic, seed 2 (puresoft3d_b200/scenes.py).
A step = one frame of the hot path: clearDepth + clearColour + uniforms + drawVAO (+ the sort-first composite at N>1).

metric/value : shaded fragments/s, whole job. A shaded fragment is one FragmentProcessor::process invocation
               (fragthrd.cpp:231) — counted on the device; identical on every backend because parity is bit-exact.
value        : inputs resident in HBM when the timed region starts; CUDA events on the pipe's own stream, max over ranks.
e2e          : same metric through the public API with HOST buffers: every step uploads the five vertex streams from
               pinned host memory, renders, and reads the colour target back into pinned host memory — through the
               pipelined forms of those calls (ps3d_vbo_update_async / ps3d_read_colour_async, two VBO sets, two display
               targets), so step i+1's upload overlaps step i's frame; at N > 1 every rank uploads 1/N of each stream and
               the shards are all-gathered over NVLink (the inputs cross PCIe once in total).
roofline     : the dominant kernel class, live CUDA-event time on its launching stream (ps3d_profile_*), against
               MEASURED_PEAKS.json; frame_roofline is SURVEY.md §8(d)'s whole-frame formula.
cpu_baseline : the reference's own renderer (oracle/_ref, the unmodified sources through the shim) on this host's cores.

N > 1 (torchrun): sort-first — rank r renders raster rows [r*H/N, (r+1)*H/N) of the SAME frame (strong scaling); every
rank runs the position half of the geometry stage for all triangles; finished colour bands land in rank 0's target over
NCCL (NVLink), issued by the library itself on the pipe's stream (PS3D_SORTFIRST=torch: by torch.distributed).
--impl reference: the reference's CPU renderer alone (rank 0 only), same scene, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed regions, sampled in-process through NVML every 20 ms (the same
    fields as B200_PROFILING.md's nvidia-smi clocks line; nvidia-smi -lms is too coarse for a 50 ms timed region)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    power = None
                self.rows.append((sm, smax, reasons, power))
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        seen = set()
        for _, _, r, _ in self.rows:
            for bit, name in names.items():
                if r & bit:
                    seen.add(name)
        sm = [r[0] for r in self.rows]
        pw = [r[3] for r in self.rows if r[3] is not None]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(seen), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def build_scene(args):
    from puresoft3d_b200 import scenes
    if args.workload == "C2":
        return scenes.scene_heightfield(1920, 1080, grid=354, layers=4, seed=2, tex_size=2048)
    if args.workload == "C2-small":  # developer smoke of this script, not a bench line
        return scenes.scene_heightfield(1920, 1080, grid=100, layers=4, seed=2, tex_size=512)
    # BASELINE.json configs[4]: the C2 generator at 10 M triangles on a 4K / 8K target (the multi-GPU configuration;
    # not the default bench line, which stays configs[1] at every N so that the driver's scaling ratios compare like with like)
    if args.workload == "C5-4k":
        return scenes.scene_heightfield(3840, 2160, grid=1118, layers=4, seed=5, tex_size=2048)
    if args.workload == "C5-8k":
        return scenes.scene_heightfield(7680, 4320, grid=1118, layers=4, seed=5, tex_size=2048)
    raise SystemExit("unknown workload")


def bench_config(args, sc):
    """The workload, named identically by both arms (the driver compares the two dicts): static properties of the scene only.
    What differs between the arms (threads, ranks, transfers, counters of the run) lives in other keys of the line."""
    vertex_mb = sc.vertex_bytes_read() // 1000000
    tex_mb = sum(a.nbytes for t in sc.textures for a in t["layers"] if a is not None) // 1000000
    return {"workload": args.workload, "scene": sc.name, "triangles": sc.meta.get("triangles"), "resolution": [sc.width, sc.height],
            "shader": "DEF03 (Blinn-Phong + normal map)", "textures": "2 x %d^2 BGRA nearest" % sc.textures[0]["width"],
            "state": "depth test + write, back-face culling, no blending",
            "l2": "inputs larger than L2, no flush between steps: %d MB of vertex streams + %d MB of textures are read every step (126 MB L2)" % (vertex_mb, tex_mb)}


def frame_hashes(colour, depth):
    """sha256 of the frame's colour words (top-down BGRA8, W*4 per row) and depth words (bottom-up float32): equal hashes on
    1, 2, 4, 8 GPUs and on the reference arm are the check that sort-first sharding changes no result."""
    import hashlib
    return (hashlib.sha256(np.ascontiguousarray(colour).view(np.uint8).tobytes()).hexdigest(),
            hashlib.sha256(np.ascontiguousarray(depth).view(np.uint8).tobytes()).hexdigest())


def algorithmic_bytes(scene, tex_samples):
    """vertex bytes read once + every bound target written once + targets loaded because the clear is a
    separate pass here? No: the formula counts clears as fused (0 B) — kept as specified — + min(texture bytes, 4 B x samples)."""
    vertex = scene.vertex_bytes_read()
    targets = 2 * scene.width * scene.height * 4
    tex = 0
    for t in scene.textures:
        tex += min(sum(a.nbytes for a in t["layers"]), 4 * tex_samples)
    return vertex, targets, tex


def run_reference(args, rank):
    """The reference's own CPU renderer (unmodified sources, oracle/_ref) with every host thread it can use."""
    if rank != 0:
        return 0
    from puresoft3d_b200 import _capi, scenes
    from puresoft3d_b200.pipeline import PuresoftPipeline
    so = os.path.join(ROOT, "oracle", "_ref", "libps3d_ref.so")
    kind = "reference"
    if not os.path.exists(so):
        so = os.path.join(ROOT, "oracle", "libps3d_oracle.so")  # the restatement, single-threaded
        kind = "port"
    cores = os.cpu_count() or 1
    os.environ.setdefault("PS3D_REF_THREADS", str(cores))
    sc = build_scene(args)
    lib = _capi.bind(so)
    # pass 1 (untimed, counting decorators on): fragments per frame
    os.environ["PS3D_REF_COUNTING"] = "1"
    p = PuresoftPipeline(sc.width, sc.height, lib=lib)
    up = scenes.upload(p, sc)
    scenes.replay(p, sc, up)
    frags = p.getStats()["fragments_shaded"]
    colour_sha, depth_sha = frame_hashes(p.readColour(), p.readDepth())
    p.close()
    # pass 2 (timed, decorators off): the stock code path
    os.environ["PS3D_REF_COUNTING"] = "0"
    p = PuresoftPipeline(sc.width, sc.height, lib=lib)
    up = scenes.upload(p, sc)
    for _ in range(args.warmup):
        scenes.replay(p, sc, up)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scenes.replay(p, sc, up)
        p.swapBuffers()
    dt = time.perf_counter() - t0
    p.close()
    ms = dt * 1000.0 / args.steps
    value = frags * args.steps / dt
    threads = cores if kind == "reference" else 1
    line = {
        "impl": "reference", "metric": "shaded_fragments_per_s", "value": value, "unit": "fragments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "frames_per_s": 1000.0 / ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, sc), "parallelism": "%d host threads" % threads,
        "fragments_per_frame": frags, "colour_sha256": colour_sha, "depth_sha256": depth_sha,
        "cpu_baseline": {"value": value, "unit": "fragments/s", "cores": threads, "kind": kind,
                         "sample": "%d full frames of the workload, %d warm-up" % (args.steps, args.warmup)},
        "e2e": {"value": value, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        return run_reference(args, rank)
    args.steps = args.steps if args.steps is not None else 200
    args.warmup = args.warmup if args.warmup is not None else 10

    import torch
    import torch.distributed as dist
    from puresoft3d_b200 import scenes
    from puresoft3d_b200.pipeline import PuresoftPipeline
    from puresoft3d_b200 import sortfirst

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    sc = build_scene(args)
    pipe = PuresoftPipeline(sc.width, sc.height, device=local_rank)
    up = scenes.upload(pipe, sc)
    ext = torch.cuda.ExternalStream(pipe.deviceStream(), device=dev)
    native = sortfirst.init_native_comm(pipe, rank, world, dev) if world > 1 else False
    # the composite: peer stores into rank 0's target over NVLink (default) | the library's NCCL send/recv | torch.distributed
    peer_comp = sortfirst.init_peer_composite(pipe, rank, world, dev) if world > 1 else False
    native_comp = native and not peer_comp and os.environ.get("PS3D_SORTFIRST_COMPOSITE", "peer") != "torch"
    comp = sortfirst.Compositor(pipe, rank, world, dev, ext, native=native_comp, peer=peer_comp) if world > 1 else None
    if comp:
        pipe.setRowBand(*comp.band)

    replay_frame = scenes.compile_replay(pipe, sc, up)   # the frame's command list with its ctypes arguments prepared once

    def frame_calls():
        replay_frame()
        if comp:
            comp.gather_to_rank0()

    # The frame as ONE launch: its calls are captured once into a CUDA graph (ps3d_graph_*) and replayed. What a frame costs the
    # host matters when its kernels take tens of microseconds (8 ranks). A composite made of NCCL calls stays outside.
    use_graph = os.environ.get("PS3D_GRAPH", "1") != "0" and (comp is None or peer_comp)
    graphs = {}

    def capture(key, calls):
        pipe.graphBegin()
        calls()
        graphs[key] = pipe.graphEnd()

    def frame():
        if use_graph and "frame" in graphs:
            pipe.graphLaunch(graphs["frame"])
        else:
            frame_calls()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Frames in flight: a frame is a chain of kernels, several of them latency-bound (a tile's list, a block's phases), and a
    # sort-first rank's share of a frame does not fill the GPU: a second pipe (own stream, own scratch buffers and targets; on
    # ranks != 0 peer-mapped to rank 0's second pipe) renders frame i + 1 while frame i is still in flight. Same frames, same
    # images, counted once each. PS3D_FRAMES_IN_FLIGHT=1 turns it off.
    n_flight = max(1, int(os.environ.get("PS3D_FRAMES_IN_FLIGHT", "2" if world > 1 else "1")))   # (one GPU gains nothing: its kernels fill it)
    if comp is not None and not peer_comp:
        n_flight = 1
    flights = [dict(pipe=pipe, ext=ext, calls=frame_calls, launch=frame)]
    for _ in range(1, n_flight):
        q = PuresoftPipeline(sc.width, sc.height, device=local_rank)
        q_up = scenes.upload(q, sc)
        q_ext = torch.cuda.ExternalStream(q.deviceStream(), device=dev)
        q_comp = None
        if world > 1:
            if not sortfirst.init_peer_composite(q, rank, world, dev):
                raise SystemExit("the second pipe's peer composite could not be set up although the first one's was")
            q_comp = sortfirst.Compositor(q, rank, world, dev, q_ext, peer=True)
            q.setRowBand(*q_comp.band)
        q_replay = scenes.compile_replay(q, sc, q_up)

        def q_calls(q_replay=q_replay, q_comp=q_comp):
            q_replay()
            if q_comp:
                q_comp.gather_to_rank0()
        flights.append(dict(pipe=q, ext=q_ext, calls=q_calls, launch=q_calls))

    # ---- kernel-only: inputs resident in HBM ------------------------------------------------------------------
    for _ in range(args.warmup):
        for f in flights:
            f["calls"]()
    for f in flights:
        f["pipe"].finish()
    # per-kernel-class times (roofline, kernel_ms_per_frame): a pass of the same frames with an event pair around every kernel
    # class (ps3d_profile_*), before the timed region — a captured frame carries no such events, and they are not free
    pipe.profileEnable(True)
    for _ in range(args.steps):
        frame_calls()
    pipe.finish()
    prof = pipe.profileRead()
    pipe.profileEnable(False)
    if use_graph:
        capture("frame", frame_calls)
        for f in flights[1:]:
            f["pipe"].graphBegin()
            f["calls"]()
            g = f["pipe"].graphEnd()
            f["launch"] = (lambda q=f["pipe"], g=g: q.graphLaunch(g))
        for _ in range(3):
            for f in flights:
                f["launch"]()
        for f in flights:
            f["pipe"].finish()
    for f in flights:
        f["pipe"].resetStats()
    launches0 = sum(f["pipe"].deviceLaunchCount() for f in flights)
    ev0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in flights]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0.record(ext)                       # every stream is idle here (barrier + device synchronize)
    for i in range(args.steps):
        flights[i % len(flights)]["launch"]()
    for f, e in zip(flights, ends):
        e.record(f["ext"])
    for e in ends:
        e.synchronize()
    barrier()
    ms_total = max(ev0.elapsed_time(e) for e in ends)
    launches = sum(f["pipe"].deviceLaunchCount() for f in flights) - launches0
    stats = pipe.getStats()
    for f in flights[1:]:
        st = f["pipe"].getStats()
        for k in ("fragments_shaded", "fragments_tested", "draws"):
            stats[k] += st[k]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    frag_t = torch.tensor([float(stats["fragments_shaded"])], dtype=torch.float64, device=dev)
    tested_per_frame = stats["fragments_tested"] / args.steps
    per_rank_ms = [ms_total / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(v.item()) / args.steps for v in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(frag_t, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    frags_total = float(frag_t.item())
    ms_step = ms_total / args.steps
    value = frags_total / (ms_total / 1000.0)
    frags_per_frame = frags_total / args.steps

    # ---- end to end: host buffers in, colour image out, every step ----------------------------------------------
    # Through the public API with the transfers PIPELINED (include/ps3d.h "Pipelined transfers"): two sets of VBOs, so
    # that step i+1's vertex streams cross PCIe (copy stream) while step i renders (pipe stream) and step i's image goes
    # back (read-back stream, double-buffered display targets via swapBuffers). At N > 1 every rank uploads only its 1/N
    # of each stream from its own pinned buffer and the shards are all-gathered in place over NVLink
    # (sortfirst.ShardedUpload): the step's inputs cross PCIe once in total, not once per rank.
    ups = [up, scenes.upload(pipe, sc)]
    uploaders, host_colour = [], []
    host_streams = {}
    for u in ups:
        items = []
        for (vbo, arr) in u.vbos:
            key = arr.ctypes.data
            if key not in host_streams:
                host_streams[key] = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).pin_memory()
            items.append((vbo, host_streams[key]))
        uploaders.append(sortfirst.ShardedUpload(pipe, items, rank, world, dev, native=native))
        host_colour.append(torch.empty((sc.height, sc.width), dtype=torch.int32).pin_memory())
    h2d_t = torch.tensor([float(uploaders[0].h2d_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)
    h2d = int(h2d_t.item())          # all ranks together
    d2h = host_colour[0].numel() * 4

    replay_set = [replay_frame, scenes.compile_replay(pipe, sc, ups[1])]

    def frame_e2e(i):
        s = i & 1
        uploaders[s].step()
        replay_set[s]()
        if comp:
            comp.gather_to_rank0()
        if rank == 0:
            pipe.readColourAsync(host_colour[s].data_ptr(), sc.width * 4)
        pipe.swapBuffers()

    e2e_steps = max(4, min(args.steps, 40))
    for i in range(4):
        frame_e2e(i)
    pipe.finish()
    pipe.resetStats()
    barrier()
    ev0.record(ext)
    for i in range(e2e_steps):
        frame_e2e(i)
    pipe.deviceJoin()                # the last read-back belongs to the timed region
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record(ext)
    ev1.synchronize()
    pipe.finish()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    frag_t = torch.tensor([float(pipe.getStats()["fragments_shaded"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(frag_t, op=dist.ReduceOp.SUM)
    e2e_value = float(frag_t.item()) / (float(t.item()) / 1000.0)
    # the images the pipelined steps read back must be the frame the kernel-only leg renders (rank 0 holds the composite)
    frame_calls()
    e2e_check = None
    if rank == 0:
        want = pipe.readColour().view(np.uint32)
        e2e_check = bool(np.array_equal(want, host_colour[0].numpy().view(np.uint32)) and np.array_equal(want, host_colour[1].numpy().view(np.uint32)))
    clocks = sampler.stop()   # sampled across both timed regions
    # the frame's hashes (rank 0's colour target holds the composite; depth bands are collected here, outside any timed region)
    depth_full = pipe.readDepth()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (comp.band, depth_full[comp.band[0]:comp.band[1]].copy()))
        for (b0, b1), rows in parts:
            depth_full[b0:b1] = rows
    colour_sha, depth_sha = frame_hashes(pipe.readColour(), depth_full) if rank == 0 else (None, None)

    # ---- roofline ------------------------------------------------------------------------------------------------
    peak, peak_src = load_peaks()
    vertex_b, target_b, tex_b = algorithmic_bytes(sc, int(2 * frags_per_frame))
    classes = {
        # algorithmic bytes per launch of each kernel class (DESIGN.md "Kernels"): what an ideal implementation must move
        "geom_setup": (prof["geom_ms"], prof["geom_launches"], vertex_b),
        "binning": (prof["bin_ms"], prof["bin_launches"], 0),
        "tile_raster_depth": (prof["tile_ms"], prof["tile_launches"], target_b // 2),
        "shade": (prof["shade_ms"], prof["shade_launches"], target_b // 2 + tex_b),
    }
    dominant = max(classes, key=lambda k: classes[k][0])
    dom_ms, dom_launches, dom_bytes = classes[dominant]
    draws = max(1, stats["draws"])
    dom_ms_per_launch = dom_ms / draws if dom_ms > 0 else float("nan")
    achieved = (dom_bytes / 1e9) / (dom_ms_per_launch / 1e3) if dom_ms > 0 else 0.0
    frame_bytes = vertex_b + target_b + tex_b
    frame_achieved = (frame_bytes / 1e9) / (ms_step / 1e3)
    traffic, issue_frac, kernel_issue = None, None, {}
    try:  # per launch, from the committed `ncu --set full` capture of this build: DRAM bytes and warp instructions
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(args.workload, {})
        traffic = tj.get(dominant, {}).get("bytes") if world == 1 else None
        # The path is instruction-issue bound (exact IEEE operation order: no FMA, true divides), so next to the HBM
        # roofline: warp instructions per launch (ncu) / live launch time / (SMs x 4 schedulers x SM clock).
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 4 * sm_hz
        for k, (ms, _, _) in classes.items():
            wi = tj.get(k, {}).get("warp_inst")
            if wi and ms > 0 and world == 1:
                kernel_issue[k] = (wi / (ms / draws / 1e3)) / issue_peak
        issue_frac = kernel_issue.get(dominant)
    except Exception:
        pass
    if rank == 0:
        line = {
            "metric": "shaded_fragments_per_s", "value": value, "unit": "fragments/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "frames_per_s": 1000.0 / ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, sc),
            "parallelism": ("sort-first row bands x%d, composite to rank 0 %s" % (world, comp.how)) if world > 1 else "single GPU",
            "frame_launch": "one CUDA graph launch per frame (ps3d_graph_*)" if use_graph else "one launch per kernel",
            "frames_in_flight": len(flights), "ms_per_step_by_rank": per_rank_ms,
            "fragments_per_frame": frags_per_frame, "fragments_tested_per_frame": tested_per_frame,
            "approx": "x86 rcpps/rsqrtss tables bits=%s" % (pipe.hostApproxInfo(),),
            "colour_sha256": colour_sha, "depth_sha256": depth_sha,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "fragments/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t.item()) / e2e_steps, "steps": e2e_steps,
                    "transfers": "pipelined: double-buffered VBO sets + display targets, uploads on a copy stream" + (", 1/%d of every stream per rank + in-place NCCL all-gather" % world if world > 1 else ""),
                    "image_matches_kernel_only_leg": e2e_check},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "ms_per_launch": dom_ms_per_launch, "algorithmic_bytes": dom_bytes,
                         "issue_frac": issue_frac, "note": "issue-bound, not HBM-bound: issue_frac = warp instructions (ncu, profiles/traffic.json) / launch time / (SMs x 4 x SM clock)"},
            "frame_roofline": {"achieved": frame_achieved, "peak": peak, "unit": "GB/s", "frac": frame_achieved / peak,
                               "algorithmic_bytes": frame_bytes, "roofline_us": frame_bytes / (peak * 1e3)},
            "kernel_ms_per_frame": {k: v[0] / args.steps for k, v in classes.items()},
            "kernel_issue_frac": kernel_issue,
            "bin_pairs_per_frame": prof["bin_pairs"] / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, frags_per_frame)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    for f in flights[1:]:
        f["pipe"].close()
    pipe.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args, frags_per_frame):
    """The reference's renderer on this box's host cores, a bounded sample (about 10-30 s): whole frames of the same scene."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", args.workload], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    for ln in out.stdout.splitlines()[::-1]:
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"value": None, "unit": "fragments/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: " + out.stderr[-200:]}


if __name__ == "__main__":
    sys.exit(main())
