"""Worker of tests/test_sortfirst_nccl_gpu.py: one rank of a world_size-N sort-first frame on real GPUs, through the PRODUCT
path (CUDA library, its own composite step: ps3d_composite_bands / the peer-store composite). Rank 0 compares the composite
with the golden fixture of the scene (generated from the reference build) and with the oracle's whole frame; depth bands are
collected with torch.distributed and compared too. Writes 'ok' or the mismatch to the result file."""
import hashlib
import json
import os
import sys
import zlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from puresoft3d_b200 import _capi, scenes, sortfirst  # noqa: E402
from puresoft3d_b200.pipeline import PuresoftPipeline  # noqa: E402
from _scenes_small import SMALL  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    scene_name, out_path = sys.argv[1], sys.argv[2]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sc = SMALL[scene_name]()
    p = PuresoftPipeline(sc.width, sc.height, device=local)
    ext = torch.cuda.ExternalStream(p.deviceStream(), device=dev)
    mode = sys.argv[3] if len(sys.argv) > 3 else "peer"
    os.environ["PS3D_SORTFIRST_COMPOSITE"] = mode
    native = sortfirst.init_native_comm(p, rank, world, dev)
    peer = sortfirst.init_peer_composite(p, rank, world, dev)
    comp = sortfirst.Compositor(p, rank, world, dev, ext, native=native and not peer, peer=peer)
    p.setRowBand(*comp.band)
    up = scenes.upload(p, sc)
    msgs = []
    for rep in range(2):                       # several frames in flight back to back: the composite must not race the next frame's clear
        scenes.replay(p, sc, up, finish=False)
        comp.gather_to_rank0()
    # ... and once as a captured frame (ps3d_graph_*), where the composite can be captured (no NCCL call inside)
    if peer:
        p.graphBegin()
        scenes.replay(p, sc, up, finish=False)
        comp.gather_to_rank0()
        g = p.graphEnd()
        p.graphLaunch(g)
    else:
        scenes.replay(p, sc, up, finish=False)
        comp.gather_to_rank0()
    p.finish()
    dist.barrier()
    torch.cuda.synchronize()
    colour = p.readColour().view(np.uint32)
    depth = p.readDepth()
    shaded = torch.tensor([p.getStats()["fragments_shaded"]], dtype=torch.int64, device=dev)
    dist.all_reduce(shaded)
    parts = [None] * world
    dist.all_gather_object(parts, (comp.band, depth[comp.band[0]:comp.band[1]].copy()))
    for (b0, b1), rows in parts:
        depth[b0:b1] = rows
    if rank == 0:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "index.json")))["scenes"][scene_name]
        raw = zlib.decompress(open(os.path.join(ROOT, "tests", "golden", g["colour_file"]), "rb").read())
        gold = np.frombuffer(raw, dtype=np.uint32).reshape(g["height"], g["width"])
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
        if sha(depth.view(np.uint32)) != g["depth_sha256"]:
            msgs.append("depth words differ from the golden fixture")
        if int(shaded.item()) != 3 * g["stats"]["fragments_shaded"]:
            msgs.append("fragments shaded %d vs golden 3 x %d" % (int(shaded.item()), g["stats"]["fragments_shaded"]))
        ca = colour.view(np.uint8).reshape(colour.shape + (4,)).astype(np.int16)
        cb = gold.view(np.uint8).reshape(gold.shape + (4,)).astype(np.int16)
        frac = float((np.abs(ca - cb).max(axis=-1) <= 1).mean())
        if frac < 0.999:
            msgs.append("colour within 1/255 of the golden frame on only %.5f of pixels" % frac)
        # ... and bit for bit the frame one GPU renders alone
        q = PuresoftPipeline(sc.width, sc.height, device=local)
        scenes.render(q, sc)
        whole = q.readColour().view(np.uint32)
        q.close()
        if not np.array_equal(colour, whole):
            msgs.append("composite differs from the single-GPU frame in %d pixels" % int((colour != whole).sum()))
        with open(out_path, "w") as f:
            f.write("ok %s via %s" % (sha(colour), comp.how) if not msgs else "; ".join(msgs))
    dist.barrier()
    p.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
