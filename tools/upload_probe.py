"""torchrun probe (N ranks): where does the sharded upload's time go? H2D of this rank's shard alone, the in-place
all-gathers alone, and both pipelined, with and without NUMA-local pinned memory (nvmlDeviceSetCpuAffinity)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if os.environ.get("PS3D_NUMA", "0") == "1":
    import pynvml as nv
    nv.nvmlInit()
    nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(lr))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
from puresoft3d_b200 import sortfirst
from puresoft3d_b200.pipeline import PuresoftPipeline

pipe = PuresoftPipeline(64, 64, device=lr)
native = sortfirst.init_native_comm(pipe, rank, world, dev)
nverts = 3007584
vbos = []
for unit in (16, 16, 16, 16, 8):
    v = pipe.createVBO(unit, nverts)
    host = torch.from_numpy(np.random.default_rng(unit).integers(0, 255, nverts * unit, dtype=np.uint8)).pin_memory()
    vbos.append((v, host))
up = sortfirst.ShardedUpload(pipe, vbos, rank, world, dev, native=native)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    pipe.finish(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    pipe.finish(); torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3


def h2d_only():
    for vbo, host, per, rem, _, _ in up.items:
        vbo.updateContentAsync(host.data_ptr() + rank * per * vbo.unitBytes, rank * per, per)


def gather_only():
    for vbo, *_ in up.items:
        vbo.allGather()


a, b, c = timed(h2d_only), timed(gather_only), timed(up.step)
if rank == 0:
    print("N=%d numa=%s native=%s: h2d of 1/N %.3f ms, 5 all-gathers %.3f ms, both pipelined %.3f ms" % (world, os.environ.get("PS3D_NUMA", "0"), native, a, b, c), flush=True)
pipe.close()
dist.destroy_process_group()
