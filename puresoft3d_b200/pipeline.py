"""Host-side mirror of the reference's pipeline surface over the C-ABI (include/ps3d.h).

Same member names, argument meaning and error behaviour as `class PuresoftPipeline`
(/root/reference/src/puresoft3d/pipeline.h:24-66) and `PuresoftVBO` (vbo.h:16-21), so that a scene script
reads like the reference's demo code (src/test/puresoft.cpp:113-206). std::out_of_range -> IndexError,
std::invalid_argument -> ValueError, std::bad_alloc -> MemoryError. The C++ twin of this file is
include/puresoft3d_b200.hpp.

The class is library-agnostic: it drives whatever `lib` (a ctypes handle from _capi.bind) it is given. The
default is the CUDA product library; only the tests hand it an oracle library.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (BEHAVIOR_ALPHABLEND, BEHAVIOR_FACE_CULLING, BEHAVIOR_TEST_DEPTH, BEHAVIOR_UPDATE_DEPTH,  # noqa: F401
                    PROC_FRAGMENT, PROC_INTERPOLATION, PROC_VERTEX, WRAP_CLAMP, WRAP_WRAP)


class DeviceError(RuntimeError):
    pass


class UnsupportedError(RuntimeError):
    pass


_EXC = {
    _capi.ERR_OUT_OF_RANGE: IndexError,
    _capi.ERR_INVALID_ARGUMENT: ValueError,
    _capi.ERR_BAD_ALLOC: MemoryError,
    _capi.ERR_DEVICE: DeviceError,
    _capi.ERR_UNSUPPORTED: UnsupportedError,
}


class PuresoftProcessor:
    """A shader object. In the reference this is a C++ object with virtual methods (proc.h:8-71); here it names
    a device functor compiled into the library (`functor`) and which of the V/I/F roles it plays (`kind`)."""
    kind = None
    functor = None

    def __init__(self, functor=None):
        if functor is not None:
            self.functor = functor


class PuresoftPostProcessor:
    """proc.h:89-94. The reference calls process(threadIndex, threadCount, frame, depth) on every worker; here the object
    names a device functor run over the whole colour target by one kernel (include/ps3d.h, ps3d_post_process)."""
    functor = None


class PP_DepthofField(PuresoftPostProcessor):
    """src/test2/testpost.cpp:9-43 — the reference's only post-processor, an unfinished stub: +50 on every byte, wrapping."""
    functor = _capi.POST_DEPTHOFFIELD


def _proc(name, kind, functor):
    return type(name, (PuresoftProcessor,), {"kind": kind, "functor": functor})


# built-in shader families (src/puresoft3d/defproc.h) — class names as the reference spells them
VertexProcesserDEF01 = _proc("VertexProcesserDEF01", PROC_VERTEX, _capi.FN_DEF01)
InterpolationProcessorDEF01 = _proc("InterpolationProcessorDEF01", PROC_INTERPOLATION, _capi.FN_DEF01)
FragmentProcessorDEF01 = _proc("FragmentProcessorDEF01", PROC_FRAGMENT, _capi.FN_DEF01)
VertexProcesserDEF02 = _proc("VertexProcesserDEF02", PROC_VERTEX, _capi.FN_DEF02)
InterpolationProcessorDEF02 = _proc("InterpolationProcessorDEF02", PROC_INTERPOLATION, _capi.FN_DEF02)
FragmentProcessorDEF02 = _proc("FragmentProcessorDEF02", PROC_FRAGMENT, _capi.FN_DEF02)
VertexProcesserDEF03 = _proc("VertexProcesserDEF03", PROC_VERTEX, _capi.FN_DEF03)
InterpolationProcessorDEF03 = _proc("InterpolationProcessorDEF03", PROC_INTERPOLATION, _capi.FN_DEF03)
FragmentProcessorDEF03 = _proc("FragmentProcessorDEF03", PROC_FRAGMENT, _capi.FN_DEF03)
VertexProcesserDEF04 = _proc("VertexProcesserDEF04", PROC_VERTEX, _capi.FN_DEF04)
InterpolationProcessorDEF04 = _proc("InterpolationProcessorDEF04", PROC_INTERPOLATION, _capi.FN_DEF04)
FragmentProcessorDEF04 = _proc("FragmentProcessorDEF04", PROC_FRAGMENT, _capi.FN_DEF04)
VertexProcesserDEF05 = _proc("VertexProcesserDEF05", PROC_VERTEX, _capi.FN_DEF05)
InterpolationProcessorDEF05 = _proc("InterpolationProcessorDEF05", PROC_INTERPOLATION, _capi.FN_DEF05)
FragmentProcessorDEF05 = _proc("FragmentProcessorDEF05", PROC_FRAGMENT, _capi.FN_DEF05)


class PuresoftVBO:
    """`new PuresoftVBO(unitBytes, unitCount)` + `updateContent(src)` (vbo.h:16-18). Storage lives in the
    library (device memory for the CUDA build); the object is bound to the pipeline that created it."""

    def __init__(self, pipeline, unitBytes, unitCount):
        self._pipe = pipeline
        self.unitBytes = int(unitBytes)
        self.unitCount = int(unitCount)
        h = C.c_int(-1)
        pipeline._check(pipeline._lib.ps3d_vbo_create(pipeline._h, self.unitBytes, self.unitCount, C.byref(h)))
        self.handle = h.value

    def updateContent(self, src):
        a = np.ascontiguousarray(src)
        if a.nbytes != self.unitBytes * self.unitCount:
            raise ValueError("updateContent: expected %d bytes, got %d" % (self.unitBytes * self.unitCount, a.nbytes))
        self._pipe._check(self._pipe._lib.ps3d_vbo_update(self._pipe._h, self.handle, a.ctypes.data))

    def updateContentDevice(self, dev_ptr):
        """Source already resident in HBM (a CUDA device pointer)."""
        self._pipe._check(self._pipe._lib.ps3d_vbo_update_device(self._pipe._h, self.handle, C.c_void_p(dev_ptr)))

    # ---- pipelined transfers (CUDA library only; include/ps3d.h "Pipelined transfers") -----------------------
    def updateContentAsync(self, pinned_ptr, firstUnit=0, unitCount=None):
        """Units [firstUnit, firstUnit + unitCount) from PINNED host memory (an address) on the pipe's copy stream."""
        n = self.unitCount - int(firstUnit) if unitCount is None else int(unitCount)
        self._pipe._check(self._pipe._lib.ps3d_vbo_update_async(self._pipe._h, self.handle, int(firstUnit), n, C.c_void_p(pinned_ptr)))

    def devicePtr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._pipe._check(self._pipe._lib.ps3d_vbo_device_ptr(self._pipe._h, self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def allGather(self):
        """Sharded upload: completes the VBO from every rank's shard (ps3d_vbo_all_gather)."""
        self._pipe._check(self._pipe._lib.ps3d_vbo_all_gather(self._pipe._h, self.handle))

    def deviceWritten(self, cuda_stream):
        self._pipe._check(self._pipe._lib.ps3d_vbo_device_written(self._pipe._h, self.handle, C.c_void_p(cuda_stream)))


class PuresoftPipeline:
    def __init__(self, deviceWidth, deviceHeight, device=0, lib=None):
        self._lib = lib if lib is not None else _capi.load_product()
        self._h = C.c_void_p()
        rc = self._lib.ps3d_create(int(deviceWidth), int(deviceHeight), int(device), C.byref(self._h))
        if rc != _capi.OK:
            raise _EXC.get(rc, RuntimeError)("ps3d_create failed (%d)" % rc)
        self.width = int(deviceWidth)
        self.height = int(deviceHeight)
        self._vbos = {}

    # ---- plumbing ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != _capi.OK:
            msg = self._lib.ps3d_last_error(self._h)
            raise _EXC.get(rc, RuntimeError)((msg or b"").decode("utf-8", "replace") + " (ps3d error %d)" % rc)

    def close(self):
        if self._h:
            self._lib.ps3d_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def backend(self):
        return self._lib.ps3d_backend_name().decode()

    # ---- texture api (pipeline.h:31-33) --------------------------------------------------------------------
    def createTexture(self, width, height, elemLen=4, pixels=None, scanline=None, extraLayers=0, mode=WRAP_CLAMP):
        scanline = int(scanline if scanline is not None else width * elemLen)
        ptr = None
        if pixels is not None:
            pixels = np.ascontiguousarray(pixels)
            if pixels.nbytes != scanline * height:
                raise ValueError("createTexture: pixels must be scanline*height bytes")
            ptr = pixels.ctypes.data
        idx = C.c_int(-1)
        self._check(self._lib.ps3d_texture_create(self._h, int(width), scanline, int(height), int(elemLen), ptr,
                                                  int(extraLayers), int(mode), C.byref(idx)))
        return idx.value

    def uploadTexture(self, idx, pixels, layer=0):
        a = np.ascontiguousarray(pixels)
        self._check(self._lib.ps3d_texture_upload(self._h, int(idx), int(layer), a.ctypes.data))

    def getTexture(self, idx, shape, dtype, layer=0):
        out = np.empty(shape, dtype=dtype)
        self._check(self._lib.ps3d_texture_download(self._h, int(idx), int(layer), out.ctypes.data))
        return out

    def destroyTexture(self, idx):
        self._check(self._lib.ps3d_texture_destroy(self._h, int(idx)))

    def setTextureFilter(self, idx, bilinear):
        """Extension (include/ps3d.h): PuresoftSampler2D reads texture `idx` with bilinear filtering instead of the
        reference's nearest rule. The reference build refuses (PS3D_ERR_UNSUPPORTED)."""
        self._check(self._lib.ps3d_texture_set_filter(self._h, int(idx), 1 if bilinear else 0))

    # ---- processor api (pipeline.h:36-40) ------------------------------------------------------------------
    def addProcessor(self, proc):
        idx = C.c_int(-1)
        self._check(self._lib.ps3d_processor_add(self._h, int(proc.kind), int(proc.functor), C.byref(idx)))
        return idx.value

    def destroyProcessor(self, idx):
        self._check(self._lib.ps3d_processor_destroy(self._h, int(idx)))

    def createProgramme(self, vid, iid, fid):
        idx = C.c_int(-1)
        self._check(self._lib.ps3d_programme_create(self._h, int(vid), int(iid), int(fid), C.byref(idx)))
        return idx.value

    def destroyProgramme(self, idx):
        self._check(self._lib.ps3d_programme_destroy(self._h, int(idx)))

    def useProgramme(self, idx):
        self._check(self._lib.ps3d_programme_use(self._h, int(idx)))

    # ---- vao api (pipeline.h:43-47) ------------------------------------------------------------------------
    def createVBO(self, unitBytes, unitCount):
        v = PuresoftVBO(self, unitBytes, unitCount)
        self._vbos[v.handle] = v
        return v

    def createVAO(self):
        idx = C.c_int(-1)
        self._check(self._lib.ps3d_vao_create(self._h, C.byref(idx)))
        return idx.value

    def attachVBO(self, vao, idx, vbo):
        old = C.c_int(-1)
        self._check(self._lib.ps3d_vao_attach(self._h, int(vao), int(idx), vbo.handle, C.byref(old)))
        return self._vbos.get(old.value)

    def detachVBO(self, vao, idx):
        old = C.c_int(-1)
        self._check(self._lib.ps3d_vao_detach(self._h, int(vao), int(idx), C.byref(old)))
        return self._vbos.get(old.value)

    def getVBO(self, vao, idx):
        cur = C.c_int(-1)
        self._check(self._lib.ps3d_vao_get(self._h, int(vao), int(idx), C.byref(cur)))
        return self._vbos.get(cur.value)

    def destroyVAO(self, vao):
        self._check(self._lib.ps3d_vao_destroy(self._h, int(vao)))

    # ---- rendering api (pipeline.h:50-61) ------------------------------------------------------------------
    def setViewport(self, width, height):
        self._check(self._lib.ps3d_set_viewport(self._h, int(width), int(height)))

    def setDepth(self, idx=-1):
        self._check(self._lib.ps3d_set_depth(self._h, int(idx)))

    def setUniform(self, idx, data):
        if data is None:
            self._check(self._lib.ps3d_set_uniform(self._h, int(idx), None, 0))
            return
        a = np.ascontiguousarray(data)
        self._check(self._lib.ps3d_set_uniform(self._h, int(idx), a.ctypes.data, a.nbytes))

    def drawVAO(self, vao, callerThrdForFragProc=False):
        self._check(self._lib.ps3d_draw_vao(self._h, int(vao), 1 if callerThrdForFragProc else 0))

    def finish(self):
        self._check(self._lib.ps3d_finish(self._h))

    def postProcess(self, processor):
        """pipeline.h:56 / post.cpp:3-19"""
        self._check(self._lib.ps3d_post_process(self._h, int(processor.functor)))

    def swapBuffers(self):
        self._check(self._lib.ps3d_swap_buffers(self._h))

    def enable(self, behavior):
        self._check(self._lib.ps3d_enable(self._h, int(behavior)))

    def disable(self, behavior):
        self._check(self._lib.ps3d_disable(self._h, int(behavior)))

    def clearDepth(self, furthest=1.0):
        self._check(self._lib.ps3d_clear_depth(self._h, float(furthest)))

    def clearColour(self, bgra=0):
        self._check(self._lib.ps3d_clear_colour(self._h, int(bgra) & 0xFFFFFFFF))

    # ---- read-back / counters ------------------------------------------------------------------------------
    def readColour(self):
        """(H, W) uint32 BGRA words, memory order of the top-down display buffer (row 0 = top of the image)."""
        out = np.empty((self.height, self.width), dtype=np.uint32)
        self._check(self._lib.ps3d_read_colour(self._h, out.ctypes.data, self.width * 4))
        return out

    def readDepth(self):
        """(H, W) float32, bottom-up like the reference's depth FBO (row 0 = raster row 0)."""
        out = np.empty((self.height, self.width), dtype=np.float32)
        self._check(self._lib.ps3d_read_depth(self._h, out.ctypes.data, self.width * 4))
        return out

    def writeColour(self, img):
        a = np.ascontiguousarray(img, dtype=np.uint32)
        self._check(self._lib.ps3d_write_colour(self._h, a.ctypes.data, self.width * 4))

    def writeDepth(self, img):
        a = np.ascontiguousarray(img, dtype=np.float32)
        self._check(self._lib.ps3d_write_depth(self._h, a.ctypes.data, self.width * 4))

    def getStats(self):
        s = _capi.Stats()
        self._check(self._lib.ps3d_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def resetStats(self):
        self._check(self._lib.ps3d_reset_stats(self._h))

    def debugCapture(self, width, height):
        self._check(self._lib.ps3d_debug_capture(self._h, int(width), int(height)))
        self._cap = (int(height), int(width))

    def debugReadShadeCounts(self):
        out = np.empty(self._cap, dtype=np.uint32)
        self._check(self._lib.ps3d_debug_read_shade_counts(self._h, out.ctypes.data))
        return out

    def debugClearShadeCounts(self):
        self._check(self._lib.ps3d_debug_clear_shade_counts(self._h))

    def setRowBand(self, row0=-1, row1=-1):
        self._check(self._lib.ps3d_set_row_band(self._h, int(row0), int(row1)))

    # ---- device-resident access (CUDA library only) --------------------------------------------------------
    def deviceColourPtr(self):
        p, pitch = C.c_void_p(), C.c_size_t()
        self._check(self._lib.ps3d_device_colour_ptr(self._h, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    def deviceDepthPtr(self):
        p, pitch = C.c_void_p(), C.c_size_t()
        self._check(self._lib.ps3d_device_depth_ptr(self._h, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    def deviceStream(self):
        s = C.c_void_p()
        self._check(self._lib.ps3d_device_stream(self._h, C.byref(s)))
        return s.value or 0

    def deviceCopyStream(self):
        s = C.c_void_p()
        self._check(self._lib.ps3d_device_copy_stream(self._h, C.byref(s)))
        return s.value or 0

    def readColourAsync(self, pinned_ptr, pitchBytes=None):
        """Colour target into PINNED host memory behind the work enqueued so far; complete after finish()."""
        self._check(self._lib.ps3d_read_colour_async(self._h, C.c_void_p(pinned_ptr), int(pitchBytes or self.width * 4)))

    # ---- sort-first exchange steps inside the library (include/ps3d.h) ---------------------------------------
    def commUniqueId(self):
        """Rank 0: 256 opaque bytes every rank passes to commInit."""
        buf = (C.c_uint8 * 256)()
        self._check(self._lib.ps3d_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def commInit(self, rank, world, unique_id):
        buf = (C.c_uint8 * 256).from_buffer_copy(unique_id)
        self._check(self._lib.ps3d_comm_init(self._h, int(rank), int(world), C.cast(buf, C.c_void_p)))

    def commDestroy(self):
        self._lib.ps3d_comm_destroy(self._h)

    def compositeBands(self, bands):
        flat = (C.c_int * (2 * len(bands)))(*[int(v) for b in bands for v in b])
        self._check(self._lib.ps3d_composite_bands(self._h, flat))

    PEER_BLOB = 256

    def peerExport(self):
        """This rank's display targets and flag block as CUDA IPC handles (PEER_BLOB opaque bytes) for peerImport on every rank."""
        buf = (C.c_uint8 * self.PEER_BLOB)()
        self._check(self._lib.ps3d_peer_export(self._h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peerImport(self, rank, world, blobs):
        """blobs: world x PEER_BLOB bytes in rank order. From here on ranks != 0 render straight into rank 0's colour target."""
        buf = (C.c_uint8 * (self.PEER_BLOB * world)).from_buffer_copy(blobs)
        self._check(self._lib.ps3d_peer_import(self._h, int(rank), int(world), C.cast(buf, C.c_void_p)))

    def peerReset(self):
        """Undo peerImport (every rank does when one rank's import failed: fall back to another composite together)."""
        self._check(self._lib.ps3d_peer_import(self._h, 0, 0, None))

    def compositePeer(self):
        """Behind every frame on every rank: 'my band is written' / rank 0 waits for every rank (include/ps3d.h)."""
        self._check(self._lib.ps3d_composite_peer(self._h))

    # ---- captured frames (include/ps3d.h) --------------------------------------------------------------------------
    def graphBegin(self):
        self._check(self._lib.ps3d_graph_begin(self._h))

    def graphEnd(self):
        g = C.c_int(-1)
        self._check(self._lib.ps3d_graph_end(self._h, C.byref(g)))
        return g.value

    def graphLaunch(self, graph):
        self._check(self._lib.ps3d_graph_launch(self._h, int(graph)))

    def graphDestroy(self, graph):
        self._check(self._lib.ps3d_graph_destroy(self._h, int(graph)))

    def deviceJoin(self):
        """The pipe's stream waits for everything enqueued so far on the copy and read-back streams."""
        self._check(self._lib.ps3d_device_join(self._h))

    def profileEnable(self, on=True):
        self._check(self._lib.ps3d_profile_enable(self._h, 1 if on else 0))

    def profileRead(self):
        pr = _capi.Profile()
        self._check(self._lib.ps3d_profile_read(self._h, C.byref(pr)))
        return pr.as_dict()

    def hostApproxInfo(self):
        a, b = C.c_int(), C.c_int()
        self._lib.ps3d_host_approx_info(C.byref(a), C.byref(b))
        return a.value, b.value

    def debugBatchCounts(self):
        """(batches of small draws launched, draws that ran inside one) since creation — include/ps3d.h."""
        b, d = C.c_uint64(), C.c_uint64()
        self._check(self._lib.ps3d_debug_batch_counts(self._h, C.byref(b), C.byref(d)))
        return int(b.value), int(d.value)

    def deviceLaunchCount(self):
        n = C.c_uint64()
        self._check(self._lib.ps3d_device_launch_count(self._h, C.byref(n)))
        return int(n.value)
