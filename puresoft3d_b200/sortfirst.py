"""Sort-first sharding across the GPUs of one box (new relative to the reference, which is single-process).

Rows are already independent units of work in Puresoft3D (each scanline is one worker's FIFO, drawvao.cpp:78-85), and
a pixel depends only on the ordered list of triangles covering it, so the screen is split into contiguous bands of
raster rows, one per rank (tile-row aligned). Every rank holds all vertex streams, uniforms and textures, runs the
geometry stage for every triangle, bins only the spans inside its band (ps3d_set_row_band) and shades them; then the
finished colour bands are gathered onto rank 0 — the path's one exchange step — with NCCL send/recv over NVLink.
No pixel is touched by two ranks, so the composite is a plain copy and sharding cannot change a result.
"""
import torch
import torch.distributed as dist

TILE = 16


def row_bands(height, world, tile=TILE):
    """Contiguous raster-row bands [r0, r1), one per rank, aligned to tile rows, covering [0, height) exactly."""
    tiles = (height + tile - 1) // tile
    bands = []
    for r in range(world):
        t0 = (tiles * r) // world
        t1 = (tiles * (r + 1)) // world
        bands.append((min(t0 * tile, height), min(t1 * tile, height)))
    return bands


def memory_rows(band, height):
    """The colour target is top-down (fbo.cpp:104-105): raster rows [r0, r1) live in memory rows [H - r1, H - r0)."""
    r0, r1 = band
    return height - r1, height - r0


def gather_bands(colour, bands, rank, world, height):
    """colour: (H, W) tensor holding this rank's finished band (other rows arbitrary). After the call rank 0 holds the
    whole frame. One batched group of point-to-point transfers (NCCL on CUDA tensors, gloo on CPU tensors)."""
    if world == 1:
        return
    ops = []
    if rank == 0:
        for r in range(1, world):
            m0, m1 = memory_rows(bands[r], height)
            if m1 > m0:
                ops.append(dist.P2POp(dist.irecv, colour[m0:m1], r))
    else:
        m0, m1 = memory_rows(bands[rank], height)
        if m1 > m0:
            ops.append(dist.P2POp(dist.isend, colour[m0:m1], 0))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _DevicePtr:
    def __init__(self, ptr, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class Compositor:
    """Binds a pipe's device colour target to torch and gathers the bands on the pipe's own stream."""

    def __init__(self, pipe, rank, world, device, ext_stream):
        self.pipe, self.rank, self.world, self.device, self.ext = pipe, rank, world, device, ext_stream
        self.bands = row_bands(pipe.height, world)
        self.band = self.bands[rank]
        self._views = {}

    def _colour(self):
        ptr, pitch = self.pipe.deviceColourPtr()
        if ptr not in self._views:
            assert pitch == self.pipe.width * 4
            self._views[ptr] = torch.as_tensor(_DevicePtr(ptr, (self.pipe.height, self.pipe.width)), device=self.device)
        return self._views[ptr]

    def gather_to_rank0(self):
        with torch.cuda.stream(self.ext):
            gather_bands(self._colour(), self.bands, self.rank, self.world, self.pipe.height)
