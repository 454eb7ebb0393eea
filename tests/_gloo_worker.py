"""Worker of tests/test_sortfirst_gloo.py: one rank of a world_size-N sort-first frame on CPU (gloo).
Each rank renders its band of raster rows of the same scene with the ORACLE library (this is test code), the colour
bands are gathered onto rank 0 with the product's own compositor code (puresoft3d_b200.sortfirst), rank 0 compares the
composite with a single-rank render of the whole frame and writes 'ok' to the result file."""
import os
import sys

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
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    scene_name, out_path = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _capi.bind(os.path.join(ROOT, "oracle", "libps3d_oracle.so"))
    sc = SMALL[scene_name]()
    bands = sortfirst.row_bands(sc.height, world)
    p = PuresoftPipeline(sc.width, sc.height, lib=lib)
    p.setRowBand(*bands[rank])
    scenes.render(p, sc)
    frame = torch.from_numpy(p.readColour().view(np.int32).copy())
    shaded = torch.tensor([p.getStats()["fragments_shaded"]], dtype=torch.int64)
    p.close()
    sortfirst.gather_bands(frame, bands, rank, world, sc.height)
    # the sharded upload's exchange (sortfirst.ShardedUpload): every rank fills only its shard (+ the tail every rank
    # uploads itself) of each vertex stream, the in-place all-gather must rebuild the whole stream on every rank
    shards_ok = True
    for slots in sc.vaos:
        for _, (unit, arr) in sorted(slots.items()):
            whole_bytes = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy())
            units = whole_bytes.numel() // unit
            per, rem = sortfirst.shard_units(units, world)
            mine = torch.full_like(whole_bytes, 0xEE)
            mine[rank * per * unit:(rank + 1) * per * unit] = whole_bytes[rank * per * unit:(rank + 1) * per * unit]
            if rem:
                mine[world * per * unit:] = whole_bytes[world * per * unit:]
            sortfirst.all_gather_shards(mine, per * unit, rank, world)
            shards_ok = shards_ok and bool(torch.equal(mine, whole_bytes))
    flag = torch.tensor([1 if shards_ok else 0], dtype=torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.all_reduce(shaded)
    if rank == 0:
        q = PuresoftPipeline(sc.width, sc.height, lib=lib)
        scenes.render(q, sc)
        whole = q.readColour().view(np.int32)
        whole_shaded = q.getStats()["fragments_shaded"]
        q.close()
        ok = np.array_equal(frame.numpy(), whole) and int(shaded.item()) == whole_shaded and int(flag.item()) == 1
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch: %d differing pixels, shaded %d vs %d" % (int((frame.numpy() != whole).sum()), int(shaded.item()), whole_shaded))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
