/* ps3d.h — the C-ABI drop-in boundary of puresoft3d_b200.
 *
 * Puresoft3D has no FFI of its own: its boundary is the C++ class surface `PuresoftPipeline`
 * (src/puresoft3d/pipeline.h:27-66) plus `PuresoftVBO` (src/puresoft3d/vbo.h:16-21), linked statically by
 * the demos. Every entry point below replaces one of those member functions (cited beside it); a C++
 * mirror of the class that forwards to these symbols is in include/puresoft3d_b200.hpp, and
 * INTEGRATION.md shows how the reference's demos bind to it.
 *
 * The SAME header is implemented by three shared libraries so that the parity tests drive them with one
 * script:
 *   puresoft3d_b200/libps3d_b200.so   the product: hand-written sm_100a CUDA (no CPU fallback)
 *   oracle/libps3d_oracle.so          the CPU restatement (oracle/ps3d_oracle.c) — test infrastructure
 *   oracle/_ref/libps3d_ref.so        the unmodified reference built through oracle/ref_shim — test infrastructure
 *
 * Conventions: plain pointers and sizes only; every call returns 0 (PS3D_OK) or a negative PS3D_ERR_*.
 * Where the reference throws std::out_of_range / std::invalid_argument / std::bad_alloc the C-ABI returns
 * PS3D_ERR_OUT_OF_RANGE / PS3D_ERR_INVALID_ARGUMENT / PS3D_ERR_BAD_ALLOC and the C++ mirror re-throws the
 * same exception type. The API is single-threaded per pipe, like the reference's.
 */
#ifndef PS3D_H
#define PS3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS3D_OK                      0
#define PS3D_ERR_OUT_OF_RANGE       (-1)
#define PS3D_ERR_INVALID_ARGUMENT   (-2)
#define PS3D_ERR_BAD_ALLOC          (-3)
#define PS3D_ERR_DEVICE             (-4)   /* CUDA error; ps3d_last_error() has the text */
#define PS3D_ERR_UNSUPPORTED        (-5)   /* e.g. a processor triple with no device functor */

/* limits: src/puresoft3d/config.h:3-9 */
#define PS3D_MAX_VBOS      16
#define PS3D_MAX_UNIFORMS  1024
#define PS3D_MAX_TEXTURES  16

/* behaviour bits: src/puresoft3d/pipeline.h:17-20 */
#define PS3D_BEHAVIOR_UPDATE_DEPTH  0x1
#define PS3D_BEHAVIOR_TEST_DEPTH    0x2
#define PS3D_BEHAVIOR_FACE_CULLING  0x4
#define PS3D_BEHAVIOR_ALPHABLEND    0x8

/* PuresoftFBO::WRAPMODE, PuresoftFBO::LAYER: src/puresoft3d/fbo.h:18-19 */
#define PS3D_WRAP_CLAMP 0
#define PS3D_WRAP_WRAP  1

/* Sampler filter of a texture (extension: PuresoftSampler2D is nearest-only, samplr2d.cpp:19-25). */
#define PS3D_FILTER_NEAREST  0
#define PS3D_FILTER_BILINEAR 1
#define PS3D_LAYER_XPOS 0
#define PS3D_LAYER_XNEG 1
#define PS3D_LAYER_YPOS 2
#define PS3D_LAYER_YNEG 3
#define PS3D_LAYER_ZPOS 4
#define PS3D_LAYER_ZNEG 5

/* processor kinds (the three abstract classes of src/puresoft3d/proc.h:27-71) */
#define PS3D_PROC_VERTEX        0
#define PS3D_PROC_INTERPOLATION 1
#define PS3D_PROC_FRAGMENT      2

/* Device functor ids. In the reference a "processor" is a C++ object with virtual methods; here it is the id
 * of a device functor compiled into the library (puresoft3d_b200/csrc/shaders.cuh). One id names the
 * V/I/F triple of one reference shader file; ps3d_processor_add(kind, id) selects the member of the triple. */
#define PS3D_FN_DEF01  1   /* tex1light1.cpp      textured Blinn-Phong                      */
#define PS3D_FN_DEF02  2   /* colr1light1.cpp     vertex-colour Blinn-Phong                 */
#define PS3D_FN_DEF03  3   /* tex1bump1light1.cpp + tangent-space normal map                */
#define PS3D_FN_DEF04  4   /* skybox.cpp          cube-map skybox                           */
#define PS3D_FN_DEF05  5   /* shadow.cpp          depth-only, 0.9 shrink                    */
/* demo 1 (src/test/testproc.cpp): earth + cloud + moon with a shadow map */
#define PS3D_FN_PLANET       16  /* VP_Planet / IP_Planet / FP_Earth (the fragment id 16 is FP_Earth)            */
#define PS3D_FN_SATELLITE    17  /* FP_Satellite (fragment only; its V/I are PS3D_FN_PLANET)                      */
#define PS3D_FN_CLOUD        18  /* VP_Cloud / IP_Cloud / FP_Cloud (alpha from the texture, round-to-nearest pack)  */
#define PS3D_FN_CLOUDSHADOW  19  /* VP_CloudShadow / IP_CloudShadow / FP_CloudShadow (the one functor that discards) */
/* demo 2 (src/test2/testproc.cpp): desk scene with a spot light and a shadow map */
#define PS3D_FN_POSITIONONLY 32  /* VP_PositionOnly (vertex) / IP_Null (interpolation) / FP_SingleColourNoLighting   */
#define PS3D_FN_SINGLECOLOUR 33  /* VP_SingleColour / IP_SingleColour / FP_SingleColour                              */
#define PS3D_FN_DIFFUSEONLY  34  /* VP_DiffuseOnly / IP_DiffuseOnly / FP_DiffuseOnly                                 */
#define PS3D_FN_SHADOW2      35  /* VP_Shadow (vertex) / FP_Null (fragment); interpolation = PS3D_FN_POSITIONONLY    */
#define PS3D_FN_FLATID       64  /* parity-test functor (not in the reference): writes a per-triangle id */
#define PS3D_FN_TEXPROBE     65  /* parity-test functor (not in the reference): clip-space position from slot 0, uv from slot 4,
                                    writes PuresoftSampler2D::get4(texture of uniform 9, u, v) unlit */

typedef struct ps3d_pipe ps3d_pipe;

/* "cuda-sm100a", "oracle-c" or "reference-shim" */
const char* ps3d_backend_name(void);
/* text of the last error on this pipe (never NULL) */
const char* ps3d_last_error(const ps3d_pipe* p);

/* PuresoftPipeline::PuresoftPipeline(hwnd, W, H, rndr) / dtor — pipeline.cpp:24-116.
 * Creates the default float depth target (scanline round(W/4)*16 bytes, bottom-up) and the top-down BGRA8
 * display target W x H; behaviour = UPDATE_DEPTH|TEST_DEPTH|FACE_CULLING. `device` = CUDA ordinal (ignored by
 * the CPU libraries). */
int ps3d_create(int width, int height, int device, ps3d_pipe** out);
int ps3d_destroy(ps3d_pipe* p);

/* createTexture / getTexture / destroyTexture — pipeline.h:31-33, tex.cpp:4-58.
 * First-free-slot handle; elemLen must be 1 or 4 (fbo.cpp:21-24); copies scanline*height bytes when pixels
 * != NULL. getTexture() exposes writable storage in the reference; here upload/download copy one layer. */
int ps3d_texture_create(ps3d_pipe* p, unsigned width, unsigned scanline, unsigned height, unsigned elemLen,
                        const void* pixels, int extraLayers, int wrapMode, int* idx);
int ps3d_texture_upload(ps3d_pipe* p, int idx, int layer, const void* pixels);
int ps3d_texture_download(ps3d_pipe* p, int idx, int layer, void* pixels);
int ps3d_texture_destroy(ps3d_pipe* p, int idx);
/* EXTENSION (no reference counterpart; BASELINE.json's configs ask for bilinear sampling): how
 * PuresoftSampler2D::get4 reads this texture. NEAREST (the default) is the reference's sampler, bit for bit. BILINEAR:
 * texel i sits at u = i / width as in the reference's nearest rule; x = width*u, y = height*v; the four texels around
 * (floor(x), floor(y)) are fetched through clampCoord (so CLAMP / WRAP behave as for nearest) and every 8-bit channel is
 * lerp(lerp(c00, c10, fx), lerp(c01, c11, fx), fy) in fp32 with lerp(a, b, t) = a + (b - a) * t (separate mul and add),
 * then (int)(value + 0.5f). The CPU restatement in oracle/ implements the same definition ("parity unpinned": the
 * reference has nothing to compare with); the reference build (oracle/_ref) returns PS3D_ERR_UNSUPPORTED. */
int ps3d_texture_set_filter(ps3d_pipe* p, int idx, int filter);

/* new PuresoftVBO(unitBytes, unitCount) / updateContent / delete — vbo.h:16-18, vbo.cpp:8-31 */
int ps3d_vbo_create(ps3d_pipe* p, size_t unitBytes, size_t unitCount, int* vbo);
int ps3d_vbo_update(ps3d_pipe* p, int vbo, const void* src);
int ps3d_vbo_destroy(ps3d_pipe* p, int vbo);

/* createVAO / attachVBO / detachVBO / getVBO / destroyVAO — pipeline.h:43-47, pipeline.cpp:118-205.
 * attach/detach hand the displaced VBO back (-1 = none); destroyVAO destroys the VBOs still attached. */
int ps3d_vao_create(ps3d_pipe* p, int* vao);
int ps3d_vao_attach(ps3d_pipe* p, int vao, int slot, int vbo, int* displaced);
int ps3d_vao_detach(ps3d_pipe* p, int vao, int slot, int* displaced);
int ps3d_vao_get(ps3d_pipe* p, int vao, int slot, int* vbo);
int ps3d_vao_destroy(ps3d_pipe* p, int vao);

/* addProcessor / destroyProcessor / createProgramme / destroyProgramme / useProgramme — pipeline.h:36-40,
 * prog.cpp:3-131. */
int ps3d_processor_add(ps3d_pipe* p, int kind, int functor, int* idx);
int ps3d_processor_destroy(ps3d_pipe* p, int idx);
int ps3d_programme_create(ps3d_pipe* p, int vid, int iid, int fid, int* idx);
int ps3d_programme_destroy(ps3d_pipe* p, int idx);
int ps3d_programme_use(ps3d_pipe* p, int idx);

/* setViewport / setDepth / setUniform / enable / disable / clearDepth / clearColour — pipeline.h:50-61,
 * pipeline.cpp:207-342. setDepth(idx) requires elemLen==4 && scanline%4==0 (pipeline.cpp:232-235).
 * setUniform copies the bytes; data==NULL releases the slot; values are latched at draw time. */
int ps3d_set_viewport(ps3d_pipe* p, int width, int height);
int ps3d_set_depth(ps3d_pipe* p, int textureIdx /* -1 = default depth */);
int ps3d_set_uniform(ps3d_pipe* p, int idx, const void* data, size_t len);
int ps3d_enable(ps3d_pipe* p, int behaviorBits);
int ps3d_disable(ps3d_pipe* p, int behaviorBits);
int ps3d_clear_depth(ps3d_pipe* p, float furthest);
int ps3d_clear_colour(ps3d_pipe* p, uint32_t bgra);

/* drawVAO(vao, callerThrdForFragProc) — drawvao.cpp:3-133. Silently returns OK when no programme is in use
 * or vao is out of range (drawvao.cpp:12-15). The reference call is synchronous; the CUDA library enqueues
 * on the pipe's stream and ps3d_finish() waits. `callerThread` only matters to the reference build. */
int ps3d_draw_vao(ps3d_pipe* p, int vao, int callerThread);
int ps3d_finish(ps3d_pipe* p);

/* postProcess(PuresoftPostProcessor*) — pipeline.h:56, post.cpp:3-19. The reference hands (colour target, depth target)
 * to the processor once per worker thread with a row interleave (threadIndex, threadCount; proc.h:89-94); here the
 * processor is a device functor run over the whole current colour target by one tile-parallel kernel, after every draw
 * enqueued so far. Functors:
 *   PS3D_POST_DEPTHOFFIELD  PP_DepthofField, the reference's only (unfinished) post-processor, src/test2/testpost.cpp:9-43:
 *                           adds 50 to every BYTE of every pixel, wrapping (paddb), two pixels per step — with an odd
 *                           width each row's last step also touches the pixel behind the row's end (the next row's first
 *                           pixel; past the buffer on the last row, where it is dropped here). */
#define PS3D_POST_DEPTHOFFIELD 1
int ps3d_post_process(ps3d_pipe* p, int functor);

/* swapBuffers — pipeline.cpp:314-322 (headless: flips the two display buffers). */
int ps3d_swap_buffers(ps3d_pipe* p);

/* Read-back (replaces the debug dumps saveTexture(-2)/saveTexture(-1), dbg.cpp:3-43).
 * Colour: the display target's current back buffer exactly as stored — top-down, memory row 0 = raster row
 * H-1 — `pitchBytes` >= W*4 per row. Depth: the default depth target as stored — bottom-up floats. */
int ps3d_read_colour(ps3d_pipe* p, void* bgra, size_t pitchBytes);
int ps3d_read_depth(ps3d_pipe* p, float* depth, size_t pitchBytes);
/* Host -> device copies of the same two targets (lets a caller restore state; used by the sort-first composite). */
int ps3d_write_colour(ps3d_pipe* p, const void* bgra, size_t pitchBytes);
int ps3d_write_depth(ps3d_pipe* p, const float* depth, size_t pitchBytes);

/* Counters (replace getTaskQCounters, dbg.cpp:45-61, with quantities that exist on both sides). */
typedef struct ps3d_stats
{
	uint64_t triangles_submitted;   /* iterations of the per-triangle loop, drawvao.cpp:36 */
	uint64_t triangles_rasterised;  /* survived cull, z reject and pushTriangle (drawvao.cpp:46-63) */
	uint64_t spans;                 /* scanline tasks pushed, drawvao.cpp:66-110 */
	uint64_t fragments_tested;      /* interpolateNextStep calls, fragthrd.cpp:214 */
	uint64_t fragments_shaded;      /* FragmentProcessor::process calls, fragthrd.cpp:231 — THE metric's unit */
	uint64_t draws;
} ps3d_stats;
int ps3d_get_stats(ps3d_pipe* p, ps3d_stats* out);
int ps3d_reset_stats(ps3d_pipe* p);

/* Parity hook: when enabled, a width x height uint32 image counts FragmentProcessor::process invocations per
 * raster pixel (row 0 = raster row 0 = bottom). ps3d_debug_capture(p, 0, 0) turns it off. */
int ps3d_debug_capture(ps3d_pipe* p, int width, int height);
int ps3d_debug_read_shade_counts(ps3d_pipe* p, uint32_t* counts /* height*width */);
int ps3d_debug_clear_shade_counts(ps3d_pipe* p);

/* Sort-first sharding (new; the reference has no multi-device path): restrict rasterisation to raster rows
 * [row0, row1) of the viewport. The position half of the geometry stage still runs for every triangle; a triangle whose
 * rows all miss the band leaves nothing behind, spans outside the band are dropped before binning.
 * (-1,-1) = whole viewport. */
int ps3d_set_row_band(ps3d_pipe* p, int row0, int row1);

/* The exchange steps of sort-first rendering issued from inside the library (NCCL over NVLink on the pipe's own streams;
 * no host synchronisation, no per-frame work in the host language). libnccl.so.2 is bound at run time (the copy the
 * process already has, or can load); without it these return PS3D_ERR_UNSUPPORTED.
 *  ps3d_comm_unique_id   rank 0: 256 bytes (two NCCL unique ids) to hand to every rank by any host transport.
 *  ps3d_comm_init        collective over all ranks: one communicator for the frame composite (pipe stream) and one for the
 *                        upload all-gathers (a gather stream), so that frame i's composite and frame i+1's uploads overlap.
 *  ps3d_composite_bands  bands = 2 * world ints, the raster rows [row0, row1) each rank rendered (ps3d_set_row_band):
 *                        every rank's band lands in rank 0's colour target, one grouped send/recv behind the frame.
 *  ps3d_vbo_all_gather   sharded upload: rank r has uploaded units [r * per, (r + 1) * per), per = unitCount / world,
 *                        with ps3d_vbo_update_async; one in-place all-gather behind it completes the VBO on every rank
 *                        (units behind world * per are uploaded by every rank itself). Draws wait through the VBO's event. */
int ps3d_comm_unique_id(void* id256);
int ps3d_comm_init(ps3d_pipe* p, int rank, int world, const void* id256);
int ps3d_comm_destroy(ps3d_pipe* p);
int ps3d_composite_bands(ps3d_pipe* p, const int* bands);
int ps3d_vbo_all_gather(ps3d_pipe* p, int vbo);

/* The composite without a copy step: over NVLink peer memory every rank renders its band STRAIGHT INTO rank 0's colour target
 * (the shade kernel's stores are peer stores), so what is left of the exchange is a pair of counters per frame.
 *  ps3d_peer_export     PS3D_PEER_BLOB bytes describing this rank's display targets and flag block (CUDA IPC handles); hand
 *                       every rank's blob to every rank by any host transport (one all-gather at start-up).
 *  ps3d_peer_import     blobs = world x PS3D_PEER_BLOB bytes in rank order. Ranks != 0 map rank 0's targets; from here on their
 *                       draws write rank 0's colour target (their own row band of it, ps3d_set_row_band) and their
 *                       ps3d_clear_colour is rank 0's business. (rank = world = 0, blobs = NULL undoes an import: for a rank whose
 *                       peers could not map rank 0's targets and that falls back to another composite with them.)
 *  ps3d_composite_peer  behind every frame, on every rank: ranks != 0 publish "frame written" (release at system scope), rank 0
 *                       waits on its stream until every rank has. No rank overwrites a target rank 0 has not handed out again
 *                       (rank 0 hands it out with its next clear / draw into it, i.e. behind a read-back it enqueued).
 * Every rank must issue the same sequence of clear / draw / swap / composite calls. */
#define PS3D_PEER_BLOB 256
int ps3d_peer_export(ps3d_pipe* p, void* blob);
int ps3d_peer_import(ps3d_pipe* p, int rank, int world, const void* blobs);
int ps3d_composite_peer(ps3d_pipe* p);

/* Captured frames (CUDA library only). The calls between ps3d_graph_begin and ps3d_graph_end (clears, uniforms, draws, the
 * composite) are recorded once into a CUDA graph instead of being run; ps3d_graph_launch replays them as ONE launch — what a
 * frame costs the host when its kernels take tens of microseconds (C2 on 8 GPUs). Contract: run the same frame normally
 * first until its buffers have stopped growing (twice is enough: a captured frame cannot allocate), keep resources, state and uniforms of the frame
 * unchanged between launches (uniform values are latched at capture), and capture once per (VBO set, display target)
 * combination the frame is launched with. Vertex data may change between launches (ps3d_vbo_update*): the streams are read
 * through the same pointers. A replayed frame whose intermediates no longer fit the buffers it was captured with is reported
 * by the next ps3d_finish as PS3D_ERR_INVALID_ARGUMENT ("re-capture"); a launch after a draw submitted normally has outgrown
 * (reallocated) one of the pipe's scratch buffers is refused with the same code: the recorded kernels point into the old one. */
int ps3d_graph_begin(ps3d_pipe* p);
int ps3d_graph_end(ps3d_pipe* p, int* graph);
int ps3d_graph_launch(ps3d_pipe* p, int graph);
int ps3d_graph_destroy(ps3d_pipe* p, int graph);

/* Device-resident access for the benchmark and the multi-GPU composite (CUDA library only; the CPU
 * libraries return PS3D_ERR_UNSUPPORTED). Pointers are CUDA device pointers owned by the pipe. */
int ps3d_device_colour_ptr(ps3d_pipe* p, void** devPtr, size_t* pitchBytes);
int ps3d_device_depth_ptr(ps3d_pipe* p, void** devPtr, size_t* pitchBytes);
int ps3d_device_stream(ps3d_pipe* p, void** cudaStream);
/* Same as ps3d_vbo_update / ps3d_texture_upload but the source is a device pointer (HBM-resident inputs). */
int ps3d_vbo_update_device(ps3d_pipe* p, int vbo, const void* devSrc);
/* Pipelined transfers (CUDA library only; the reference's updateContent / getBuffer are synchronous, vbo.cpp:28-31 —
 * these are the asynchronous forms the double-buffered presenter of SURVEY.md §8(f) and the sharded upload use).
 *  ps3d_vbo_update_async   units [firstUnit, firstUnit + unitCount) from PINNED host memory on the pipe's copy stream;
 *                          returns at once. The copy waits for the last draw that read the VBO, later draws that read
 *                          it wait for the copy. The source must stay valid until ps3d_finish (or any synchronous call).
 *  ps3d_vbo_device_ptr     the VBO's storage in HBM (e.g. as the buffer of an all-gather between ranks).
 *  ps3d_vbo_device_written somebody else wrote that storage on `cudaStream` (a collective, a peer copy): later draws
 *                          wait for what is enqueued on that stream now.
 *  ps3d_device_copy_stream the copy stream (a cudaStream_t), so that such writers can be ordered behind the uploads.
 *  ps3d_read_colour_async  the colour target into PINNED host memory on a read-back stream, behind everything enqueued
 *                          on the pipe's stream so far; returns at once; complete after ps3d_finish. Later writes to the
 *                          same target wait for the read-back (use ps3d_swap_buffers to overlap them). */
int ps3d_vbo_update_async(ps3d_pipe* p, int vbo, size_t firstUnit, size_t unitCount, const void* pinnedSrc);
int ps3d_vbo_device_ptr(ps3d_pipe* p, int vbo, void** devPtr, size_t* bytes);
int ps3d_vbo_device_written(ps3d_pipe* p, int vbo, void* cudaStream);
int ps3d_device_copy_stream(ps3d_pipe* p, void** cudaStream);
int ps3d_read_colour_async(ps3d_pipe* p, void* pinnedBgra, size_t pitchBytes);
/* the pipe's stream waits for everything enqueued so far on the copy and read-back streams */
int ps3d_device_join(ps3d_pipe* p);
/* number of kernels this pipe has launched since creation (bench.py's gpu_launches) */
int ps3d_device_launch_count(ps3d_pipe* p, uint64_t* launches);
/* Batches of small draws. Consecutive ps3d_draw_vao calls of at most 16384 triangles each, into the same targets with the same
 * viewport and behaviour bits (any programmes that neither blend nor discard), are collected and run as ONE pass when the next
 * call arrives that is not such a draw: their triangles are numbered in submission order across the draws, so depth and colour
 * resolve exactly as if the draws had run one after the other (drawvao.cpp:78-96 ends before the next call starts), but a
 * tile is loaded, walked and stored once for all of them. Uniforms and texture bindings are latched per draw at the call, as
 * always. Errors of a collected draw's launch surface at the call that launches the batch. PS3D_BATCH=0 turns this off.
 * ps3d_debug_batch_counts: batches launched / draws that ran inside one, since creation (tests, bench). */
int ps3d_debug_batch_counts(ps3d_pipe* p, uint64_t* batches, uint64_t* draws);

/* Per-kernel-class device timing with CUDA events recorded on the pipe's own stream (bench.py's roofline figures).
 * While enabled every draw brackets its geometry kernel, its binning kernels, its raster/tile kernel and its shade kernel with event pairs;
 * ps3d_profile_read() waits for the stream, sums the elapsed times since the last read and resets. */
typedef struct ps3d_profile
{
	double geom_ms, bin_ms, tile_ms, shade_ms; /* tile_ms: the raster + depth kernel (or the one-kernel tile paths); shade_ms: the shade kernel */
	uint64_t geom_launches, bin_launches, tile_launches, shade_launches;
	uint64_t bin_pairs;            /* (tile, triangle) pairs binned */
	uint64_t survivors;            /* records the shade kernel consumed (split path) */
} ps3d_profile;
int ps3d_profile_enable(ps3d_pipe* p, int on);
int ps3d_profile_read(ps3d_pipe* p, ps3d_profile* out);
/* which rcpps / rsqrtss emulation the kernels use: table index bits measured on this host (0,0 = IEEE 1/x, 1/sqrt).
 * NOTE — colour depends on the HOST CPU: the reference's fragment shaders use the approximate x86 instructions rcpps / rsqrtss
 * (src/mcemath/vector.cpp:165,190,245), whose results differ between CPU models. At start-up this library measures the host's own
 * instructions over all 2^23 mantissas and hands the tables to the kernels, so the GPU renders the colours the reference would
 * render ON THIS MACHINE (bit for bit, which is what the parity tests observe); on a host with a different CPU the same scene can
 * differ in the low bits of a colour channel, exactly as the reference's own output would. Coverage, depth and 1/w never use
 * these instructions and do not depend on the host. PS3D_APPROX=ieee selects correctly rounded 1/x and 1/sqrt(x) instead
 * (host-independent; colour then within the 1/255 gate, not bit-equal). */
int ps3d_host_approx_info(int* rcpBits, int* rsqrtBits);

#ifdef __cplusplus
}
#endif
#endif /* PS3D_H */
