"""Sort-first sharding across the GPUs of one box (new relative to the reference, which is single-process).

Rows are already independent units of work in Puresoft3D (each scanline is one worker's FIFO, drawvao.cpp:78-85), and
a pixel depends only on the ordered list of triangles covering it, so the screen is split into contiguous bands of
raster rows, one per rank (tile-row aligned). Every rank holds all vertex streams, uniforms and textures, runs the
geometry stage for every triangle, bins only the spans inside its band (ps3d_set_row_band) and shades them; then the
finished colour bands are gathered onto rank 0 — the path's one exchange step — with NCCL send/recv over NVLink.
No pixel is touched by two ranks, so the composite is a plain copy and sharding cannot change a result.

Both exchange steps (the band composite, and the all-gather that completes vertex streams of which every rank uploaded
only its 1/N — ShardedUpload) are issued by the library itself on the pipe's streams once init_native_comm() has handed
it NCCL ids (ps3d_comm_init / ps3d_composite_bands / ps3d_vbo_all_gather); the torch.distributed forms below
(gather_bands, all_gather_shards) are the same steps on torch tensors: the CPU (gloo) tests run them, and
PS3D_SORTFIRST=torch selects them on GPUs for A/B runs.
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


def init_native_comm(pipe, rank, world, device):
    """The library's own NCCL communicators (ps3d_comm_init): rank 0 draws the unique ids, torch.distributed carries the
    256 bytes to every rank — the only thing the host language does for the exchange steps; the composite and the upload
    all-gathers are then issued from inside the library on the pipe's streams. Returns False (torch path stays) when
    PS3D_SORTFIRST=torch or libnccl cannot be bound."""
    import os
    if world == 1 or os.environ.get("PS3D_SORTFIRST", "native") == "torch":
        return False
    buf = torch.zeros(257, dtype=torch.uint8, device=device)
    if rank == 0:
        try:
            ids = pipe.commUniqueId()
            buf[:256] = torch.frombuffer(bytearray(ids), dtype=torch.uint8).to(device)
            buf[256] = 1
        except Exception:  # noqa: BLE001 — libnccl not bindable: every rank learns it from the flag below
            pass
    dist.broadcast(buf, 0)
    host = buf.cpu().numpy()
    if host[256] != 1:
        return False
    pipe.commInit(rank, world, bytes(host[:256]))
    return True


def init_peer_composite(pipe, rank, world, device):
    """The composite over NVLink peer memory (ps3d_peer_*): every rank exports CUDA IPC handles of its display targets and flag
    block, one all-gather hands them round, ranks != 0 map rank 0's targets and render their band straight into them. Returns
    False (another composite stays in charge) when PS3D_SORTFIRST_COMPOSITE names one or the handles cannot be had."""
    import os
    if world == 1 or os.environ.get("PS3D_SORTFIRST_COMPOSITE", "peer") != "peer":
        return False
    blob = torch.zeros(pipe.PEER_BLOB, dtype=torch.uint8, device=device)
    ok = torch.ones(1, dtype=torch.int32, device=device)
    try:
        blob.copy_(torch.frombuffer(bytearray(pipe.peerExport()), dtype=torch.uint8))
    except Exception:  # noqa: BLE001 — every rank learns it from the flag below
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        return False
    blobs = torch.empty(world * pipe.PEER_BLOB, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(blobs, blob)
    try:
        pipe.peerImport(rank, world, bytes(blobs.cpu().numpy()))
    except Exception:  # noqa: BLE001 — e.g. no peer access between two GPUs of the box
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) != 1:
        pipe.peerReset()
        return False
    return True


class _DevicePtr:
    def __init__(self, ptr, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class Compositor:
    """Binds a pipe's device colour target to torch and gathers the bands on the pipe's own stream."""

    def __init__(self, pipe, rank, world, device, ext_stream, native=False, peer=False):
        self.pipe, self.rank, self.world, self.device, self.ext = pipe, rank, world, device, ext_stream
        self.bands = row_bands(pipe.height, world)
        self.band = self.bands[rank]
        self.native = native          # ps3d_composite_bands (the library's own communicator) instead of torch.distributed
        self.peer = peer              # ps3d_composite_peer: the bands were rendered straight into rank 0's target over NVLink
        self.how = ("by peer stores of the shade kernel into rank 0's target over NVLink (no copy step; two counters per frame)" if peer else
                    "by NCCL send/recv issued by the library on the pipe's stream" if native else "by torch.distributed send/recv")
        self._views = {}

    def _colour(self):
        ptr, pitch = self.pipe.deviceColourPtr()
        if ptr not in self._views:
            assert pitch == self.pipe.width * 4
            self._views[ptr] = torch.as_tensor(_DevicePtr(ptr, (self.pipe.height, self.pipe.width)), device=self.device)
        return self._views[ptr]

    def gather_to_rank0(self):
        if self.peer:
            self.pipe.compositePeer()
            return
        if self.native:
            self.pipe.compositeBands(self.bands)
            return
        with torch.cuda.stream(self.ext):
            gather_bands(self._colour(), self.bands, self.rank, self.world, self.pipe.height)


def shard_units(units, world):
    """Equal shards of a vertex stream: rank r owns units [r*per, (r+1)*per); the `rem` units behind world*per (fewer
    than `world`) are uploaded by every rank itself, so the exchange is one in-place all-gather of equal pieces."""
    per = units // world
    return per, units - per * world


def all_gather_shards(flat, per_bytes, rank, world):
    """In place: flat[r*per_bytes:(r+1)*per_bytes] holds rank r's shard on rank r; afterwards on every rank."""
    if world == 1 or per_bytes == 0:
        return
    dist.all_gather_into_tensor(flat[:world * per_bytes], flat[rank * per_bytes:(rank + 1) * per_bytes])


class ShardedUpload:
    """A step's vertex streams cross PCIe once in total instead of once per rank: rank r copies only its 1/N of every
    stream from its pinned host buffer (ps3d_vbo_update_async, on the pipe's copy stream) and the shards are
    all-gathered in place in the VBOs' own device storage over NVLink — the upload's one exchange step, on a stream of
    its own behind a per-stream event, so that stream k is gathered while stream k+1 is still crossing PCIe and the
    next step's uploads never wait for a collective. Draws wait for the gathered streams through the VBOs' ready events
    (ps3d_vbo_device_written); nothing blocks the host."""

    def __init__(self, pipe, vbos, rank, world, device, native=False):
        """vbos: [(PuresoftVBO, pinned uint8 torch tensor holding the WHOLE stream's bytes)]. native: the all-gathers are
        issued by the library itself (ps3d_vbo_all_gather, its own communicator and gather stream)."""
        self.pipe, self.rank, self.world, self.native = pipe, rank, world, native
        self.copy_stream = pipe.deviceCopyStream()
        self.ext = torch.cuda.ExternalStream(self.copy_stream, device=device)
        self.gather = torch.cuda.Stream(device=device) if (world > 1 and not native) else None
        self.items = []
        for vbo, host in vbos:
            ptr, nbytes = vbo.devicePtr()
            per, rem = shard_units(vbo.unitCount, world)
            flat = torch.as_tensor(_DevicePtr(ptr, (nbytes,), "|u1"), device=device) if (world > 1 and not native) else None
            self.items.append((vbo, host, per, rem, flat, torch.cuda.Event() if (world > 1 and not native) else None))
        self.h2d_bytes = sum((per + rem) * vbo.unitBytes for vbo, _, per, rem, _, _ in self.items)

    def step(self):
        r, w = self.rank, self.world
        for vbo, host, per, rem, flat, ev in self.items:
            base = host.data_ptr()
            vbo.updateContentAsync(base + r * per * vbo.unitBytes, r * per, per)
            if rem:
                vbo.updateContentAsync(base + w * per * vbo.unitBytes, w * per, rem)
            if w > 1 and self.native:
                vbo.allGather()
            elif w > 1:
                # the copy stream has also waited for the last draw that read this VBO (ps3d_vbo_update_async), so the
                # gather behind this event cannot overwrite a stream a geometry kernel is still reading
                ev.record(self.ext)
                self.gather.wait_event(ev)
                with torch.cuda.stream(self.gather):
                    all_gather_shards(flat, per * vbo.unitBytes, r, w)
                vbo.deviceWritten(self.gather.cuda_stream)
