// device_types.cuh — data that crosses the host/device boundary of one draw.
//
// HBM layout (DESIGN.md "Data layout"):
//   attribute streams   one contiguous buffer per VBO, fixed stride, un-indexed (vbo.cpp:8-31): read once by geom_setup
//   TriHeader[ntris]    64 B per triangle (4 x 16-B quads): screen xy, 1/w, NDC z, row range, edge plan
//   vary[ntris][3][NV]  float4 varyings exactly as the vertex functor wrote them (PROCDATA_* order)
//   bins                (tile id, triangle id) pairs, stably sorted by tile => per-tile lists in submission order
//   targets             colour BGRA8 top-down, depth float bottom-up — the reference's own memory layout (fbo.cpp:98-110)
#pragma once
#include <stdint.h>
#include "exact_math.cuh"

#define PS_TILE 16            // screen tile edge in pixels; one warp owns one tile
#define PS_SEG 8              // a lane owns one row segment of PS_SEG pixels: 16 rows x 2 segments = 32 lanes
#define PS_MAX_VARY 6         // float4 varyings per vertex (PROCDATA_PLANET, src/test/testproc.h:7-16, has 6)
#define PS_UNIFORM_SLOTS 48   // slots latched per draw, first 64 bytes each (a mat4); demo 2 uses slots up to 43 (src/test2/testproc.h:4-37)
#define PS_MAX_BOUND_TEX 6

#define PS_BEHAVIOR_UPDATE_DEPTH 0x1
#define PS_BEHAVIOR_TEST_DEPTH 0x2
#define PS_BEHAVIOR_FACE_CULLING 0x4
#define PS_BEHAVIOR_ALPHABLEND 0x8

struct TexDesc
{
	const uint8_t* layer[6];
	int width, height, scanline, wrap, nLayers, elemLen;
	int filter;                 // PS3D_FILTER_*: 0 = nearest (the reference's sampler), 1 = bilinear (extension)
};

struct TargetDesc
{
	uint8_t* ptr;
	int width, height, scanline, topDown;
};

struct alignas(16) TriHeader
{
	float vx0, vy0, vx1, vy1;   // viewport-space vertices (rasterizer.cpp:73-77)
	float vx2, vy2, rw0, rw1;   // rw = 1/w per vertex ("correction factor 1", vertthrd.cpp:37-47)
	float rw2, z0, z1, z2;      // NDC z per vertex
	uint32_t rows;              // firstRow | lastRow << 16
	uint32_t half0, half1;      // y0 | y1 << 16 of the upper / lower half (flat triangles: half0 only)
	uint32_t plan;              // 2-bit vertex ids: eL0.i0,eL0.i1,eR0.i0,eR0.i1,eL1.i0,eL1.i1,eR1.i0,eR1.i1 ; bit 16: two halves
};

// Counters live in PS_STATS_COPIES replicas, one 128-byte line each: a block adds to replica (blockIdx.x mod copies), so
// no single address takes the whole grid's atomics (one shared line cost geom_setup 34 us of 170 on C2). The host sums
// the replicas; tile_scan_kernel folds and resets the per-draw part.
#define PS_STATS_COPIES 32
struct alignas(128) DeviceStats
{
	unsigned long long triangles_rasterised;
	unsigned long long spans;
	unsigned long long fragments_tested;
	unsigned long long fragments_shaded;
	// per draw (reset by tile_scan_kernel):
	unsigned long long fragBound;   // sum of clamped span lengths = upper bound of the draw's depth-test survivors
	// the range of tile indices the draw binned anything to, as two maxima so that all-zero means "none": ~lowest, highest + 1.
	// tile_scan_kernel scans only that range (a small object on a 4096^2 shadow map touches a few hundred of 65536 tiles)
	unsigned tileLoInv, tileHi1;
	// span path, per draw: what the geometry kernel counted; tile_plan_kernel adds them to the totals above only when the
	// draw's speculated capacities held (a draw that is run again must not be counted twice), and resets them
	unsigned long long dRasterised, dSpans, dTested;
};

// What the host learns about a draw after its geometry + tile scan, written by tile_scan_kernel into mapped pinned host
// memory (read after the event recorded behind that kernel; no copy is enqueued).
struct DrawReport
{
	unsigned int bad;               // a speculated capacity was too small (or a list too long): the guarded tail did nothing
	unsigned int pairs;             // (tile, triangle) pairs = sum of the per-tile counts
	unsigned int longest;           // longest tile list
	unsigned int spans;             // span path: span records the draw needs
	unsigned long long fragBound;
	unsigned int sticky;            // set with `bad`, cleared only by the host: a replayed captured frame (ps3d_graph_launch) is checked here
	unsigned int marks;             // span path: chain marks the draw's long spans ask for (SpanStreams::markZ)
};

// Survivors of the depth test, one record per FragmentProcessor::process call still to make (split path): structure of
// arrays in HBM, appended 32 at a time by the raster/depth kernel, consumed in any order by the shade kernel.
struct SurvivorStream
{
	uint32_t* tri;        // triangle id of the draw
	int* left;            // RESULT_ROW::left / right of the span (unclamped)
	int* right;
	float* inv;           // 1 / correctionFactor2 at the pixel (interp.cpp:85)
	uint32_t* misc;       // x | y << 13 | edge code << 26 (two 3-bit ordered vertex pairs)
	uint32_t* count;      // records appended so far
	uint32_t* winner;     // vpW x vpH: 1 + index of the LAST record of each pixel (only that one's colour lands); 0 = none
	uint32_t capacity;
};

// ---- span path (kernels_span.cuh): every RESULT_ROW of every surviving triangle is evaluated ONCE, by a dense
// (triangle, row) lane of the geometry kernel, together with the depth half of interpolateStartAndStep; the tile kernel
// reads these records instead of re-deriving spans from triangle headers.
// One record per raster row of a surviving triangle, 32 bytes = one sector, read by the tile kernel (all of it) and by the shade
// kernel (the first half).
struct alignas(32) SpanRec
{
	int left, right;          // RESULT_ROW::left / right (unclamped; an empty row is stored as 1, 0)
	float zmin;               // conservative lower bound of z over the clamped span; -inf: none; NaN: too long for a precomputed bound
	uint32_t triEdges;        // triangle id of the draw (24 bits) | the row's edge plan << 24 (2-bit vertex ids: l0, l1, r0, r1)
	float cf2, cf2Step, z0, zStep;   // correctionFactor2 / projected z chains at the CLAMPED start column (interp.cpp:26-80)
};
#define PS_SPAN_MAX_TRIS 0x1000000u   // triangle ids must fit 24 bits on the span path
struct alignas(8) TriSpan { uint32_t x, y; };   // (a plain struct: this header is also compiled for the host by the functor tests)
struct alignas(8) MarkZ { float cf2, z; };   // (a plain struct: this header is also compiled for the host by the functor tests)
struct SpanStreams
{
	SpanRec* rec;
	TriSpan* tri;         // per triangle: index of its first record, first row | last row << 16 of the records (inside band and targets)
	uint32_t* count;      // records allocated so far in this draw (one atomicAdd per geometry block)
	uint32_t capacity;
	// Chain marks of long spans (more than PS_SPAN_BOUND_MAX pixels): the k-th pixel's depth and varyings are k ROUNDED additions
	// from the span start (interp.cpp:88), which no closed form reproduces — a tile in the middle of a 4096-pixel span would replay
	// thousands of them, every tile of the row again. The chain of a long span is walked ONCE instead and its state kept every
	// PS_MARK_STEP pixels; tiles and fragments start from the nearest mark.
	uint32_t* markAt;     // per record (long spans only): index of the span's first mark, ~0 = the marks did not fit (replay from the start)
	MarkZ* markZ;         // (cf2, z) after PS_MARK_STEP * k steps
	F4* markV;            // the varyings after PS_MARK_STEP * k steps: [mark][NV]
	uint32_t* longList;   // record indices of the long spans, any order
	unsigned long long* longCount;   // long spans << 40 | marks asked for (one atomicAdd per long span)
	uint32_t* longLatched;           // [0] long spans, [1] marks asked for: what the plan kernel read before it reset the counter
	uint32_t markCap;     // marks that fit; 0: none are kept (only counted)
};
#define PS_MARK_STEP 16

// per-tile triangle lists of fixed capacity, appended to directly by the geometry kernel (any order), sorted by the list sort
struct TileLists
{
	uint32_t* fill;       // per tile: entries appended in the draw in flight (may exceed cap: the plan kernel then raises poison); zeroed by the plan kernel
	uint32_t* len;        // per tile: the finished count, written by the plan kernel
	uint32_t* ids;        // tile * cap + slot
	uint32_t cap;
	// triangles whose tile rectangle is PS_BIG_AREA tiles or more (a 4096^2 shadow map's ground quad: 65 536 each): the geometry
	// kernel only notes them (3 words: tx0 | tx1 << 16, ty0 | ty1 << 16, id), tile_append_big_kernel appends them with the whole
	// grid instead of one block. NULL: every block appends its own.
	uint32_t* bigList;
	uint32_t* bigCount;   // zeroed by the plan kernel
	uint32_t bigCap;
};
#define PS_BIG_AREA 2048u

// survivors of the depth test on the span path: 12 bytes each
struct SurvivorStream2
{
	uint32_t* span;       // index of the span record
	uint32_t* xy;         // x | y << 13 | PS_SV_WINNER
	float* inv;           // 1 / correctionFactor2 at the pixel (interp.cpp:85)
	uint32_t* count;
	uint32_t capacity;
	// draws that blend (blend4 is not commutative: a pixel's survivors must land in submission order). A warp of the raster kernel
	// appends its tile's survivors in groups of 32 consecutive stream slots, in submission order pixel by pixel; the groups of a
	// tile (of a row group of a tile) are chained: chain[2 u] = 1 + first slot of the first group of unit u (0: none),
	// chain[2 u + 1] = size of its last group, next[first slot of a group] = 1 + first slot of the next (0: last). The shade kernel
	// leaves every survivor's colour in `colour` instead of storing it, and shade_resolve_kernel walks each unit's groups in order.
	// NULL for every other draw.
	uint32_t* next;
	uint32_t* chain;
	uint32_t* colour;
};
#define PS_SV_WINNER 0x80000000u   // the LAST survivor of its pixel: the only one whose colour lands
#define PS_SV_WROTE 0x40000000u    // (blending draws) the fragment functor wrote a colour ...
#define PS_SV_BLEND 0x20000000u    // ... through write4 (blend4 under ALPHABLEND) rather than write

struct DrawParams
{
	const uint8_t* slot[16];
	uint32_t stride[16];
	uint32_t ntris;
	int vpW, vpH, halfW, halfH;
	int behavior;
	int band0, band1;           // raster rows [band0, band1) are rendered (sort-first sharding)
	int tilesX, tilesY;
	float u[PS_UNIFORM_SLOTS][16];
	TexDesc tex[PS_MAX_BOUND_TEX];
	TargetDesc colour, depth;
	ApproxTables approx;
	TriHeader* hdr;
	F4* vary;
	uint32_t* triCount;         // bin entries per triangle (0 = culled / rejected / empty)
	uint32_t* triRect;          // per triangle 3 words: tx0 | tx1 << 16, ty0 | ty1 << 16, mask of touched tiles (bit = (ty-ty0)*8 + tx-tx0;
	                            // ~0 when the rectangle exceeds 8 x 4 tiles and is used whole)
	uint32_t* tileCount;        // per tile: triangles binned to it (atomics in geom_setup)
	uint32_t* workList;         // sort-first, two-kernel geometry: indices of the triangles the position half kept, any order
	uint32_t* workCount;
	const uint32_t* tileOrder;  // the tiles by descending list length (tile_scan_kernel): the order the tile kernels take them in
	DeviceStats* stats;         // PS_STATS_COPIES replicas
	const uint32_t* poison;     // != 0: a speculated capacity of this draw was too small, every kernel behind the tile scan returns at once
	uint32_t* cap;              // per-pixel FragmentProcessor::process counts (parity hook) or NULL
	int capW, capH;
	SpanStreams sp;             // span path
	TileLists tl;
	uint32_t batchFirstBlock, batchTrisPerBlock;   // a draw inside a batch (kernels_span.cuh: BatchView): its first block of ids, triangles per block
};
