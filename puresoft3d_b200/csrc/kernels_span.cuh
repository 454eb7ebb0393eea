// kernels_span.cuh — the span path: every span of a draw is computed ONCE.
//
// The reference walks a triangle once (PuresoftRasterizer::pushTriangle, rasterizer.cpp:73-234), then once per row sets up
// the interpolation (PuresoftInterpolater::interpolateStartAndStep, interp.cpp:26-80) and hands the row to a worker
// (drawvao.cpp:66-111). The first CUDA path (kernels.cuh) re-derived every row from the 64-byte triangle header in each tile
// the triangle touched and walked rows per thread. Here:
//
//   geom_span<PROG>      block = 128 triangles.
//                        A  thread = triangle: vertex functor (positions staged by one TMA bulk copy), perspective divide,
//                           back-face, z reject, pushTriangle's set-up, band / target row range; survivors compacted.
//                        S  lane = (triangle, row), dense whatever the triangles' heights: RESULT_ROW (rasterizer.cpp:98-117)
//                           and the depth half of interpolateStartAndStep at the clamped start column, a conservative lower
//                           bound of the span's depth — one 36-byte record per row, stored coalesced.
//                        B  thread = surviving triangle, dense: 64-byte header + varyings (shade kernel), the tiles its
//                           spans reach, its id appended to those tiles' lists (fixed-capacity lists, any order).
//   tile_plan            one block: per-tile counts -> lengths, capacities judged (poison), tiles ordered by list length,
//                        the draw's counters folded into the totals.
//   tile_list_sort_cap   every list sorted by triangle id = submission order.
//   tile_raster_span     one warp per 16x16 tile (or row group of one), depth tile in shared memory (128-bit loads/stores):
//                        lane = (triangle, row) candidate -> record fetch, clip to the tile, whole-span depth reject, chain
//                        replay to the tile edge; lane = pixel, dense -> interpolateNextStep + the depth rule in submission
//                        order. Survivors go to a 12-byte stream; each pixel's last survivor is flagged in place.
//   shade_span<PROG>     flat loop over the survivor stream (varyings' half of interpolateStartAndStep, exact chain replay,
//                        fragment functor once per survivor); flagged survivors store their colour.
#pragma once
#include "kernels.cuh"

#define PS_SPAN_WINDOW 1024        // (triangle, row) lanes of a block handled per pass
#define PS_SPAN_BOUND_MAX 64       // spans longer than this get no precomputed depth bound

// One RESULT_ROW + the depth half of interpolateStartAndStep. h lives in shared memory.
struct SpanOut
{
	int left, right, edges;
	int x1, x2;               // clamped to viewport and depth target; x1 > x2: nothing to draw
	float cf2, cf2Step, z0, zStep, zmin;
	bool counted;             // left != right (drawvao.cpp:72): the reference dispatched this row
};

PS_D void evalSpan(const TriHeader& h, int iy, int vpW, int limitX, SpanOut& o)
{
	o.left = 1; o.right = 0; o.edges = 0; o.x1 = 1; o.x2 = 0; o.counted = false;
	o.cf2 = o.cf2Step = o.z0 = o.zStep = 0.0f; o.zmin = -INFINITY;
	RowSpan r;
	if(!rowOfXY(h, iy, r) || r.left == r.right) return;      // rows outside both halves were never written; drawvao.cpp:72
	o.counted = true;
	o.left = r.left; o.right = r.right; o.edges = r.edges;
	const int x1 = r.left < 0 ? 0 : r.left;                   // RESULT_ROW::leftClamped
	int x2 = r.right >= vpW ? vpW - 1 : r.right;              // RESULT_ROW::rightClamped
	if(x2 > limitX) x2 = limitX;                              // (columns beyond the depth target are dropped, DESIGN.md "divergences")
	o.x1 = x1; o.x2 = x2;
	if(x1 > x2) return;
	// interpolateStartAndStep, interp.cpp:26-80
	const float* xy = &h.vx0;
	const int e = r.edges;
	float cl[3], cr[3];
	edgeContribXY(xy, e & 3, (e >> 2) & 3, (float)r.left, (float)iy, cl);
	edgeContribXY(xy, (e >> 4) & 3, (e >> 6) & 3, (float)r.right, (float)iy, cr);
	cl[0] = fmul(cl[0], h.rw0); cl[1] = fmul(cl[1], h.rw1); cl[2] = fmul(cl[2], h.rw2); // mulvec_3_4 (:40-41); lane 3 is 0*0
	cr[0] = fmul(cr[0], h.rw0); cr[1] = fmul(cr[1], h.rw1); cr[2] = fmul(cr[2], h.rw2);
	const float rcpLen = fdiv(1.0f, (float)(r.right - r.left));                           // :47
	float z0 = hsum4(fmul(cl[0], h.z0), fmul(cl[1], h.z1), fmul(cl[2], h.z2), 0.0f);      // :49 dot_3_4
	float zStep = hsum4(fmul(cr[0], h.z0), fmul(cr[1], h.z1), fmul(cr[2], h.z2), 0.0f);   // :50
	zStep = fmul(fsub(zStep, z0), rcpLen);                                                // :51
	float cf2 = hsum4(cl[0], cl[1], cl[2], 0.0f);                                         // :55-68
	float cf2Step = hsum4(cr[0], cr[1], cr[2], 0.0f);
	cf2Step = fmul(fsub(cf2Step, cf2), rcpLen);                                           // :72
	const int skip = x1 - r.left;                                                         // drawvao.cpp:90
	if(skip > 0)                                                                          // interp.cpp:74-79
	{
		cf2 = fadd(cf2, fmul(cf2Step, (float)skip));
		z0 = fadd(z0, fmul(zStep, (float)skip));
	}
	o.cf2 = cf2; o.cf2Step = cf2Step; o.z0 = z0; o.zStep = zStep;
	// Conservative bound for the whole-span depth reject (tile kernel). Along the span z = z0_k / cf2_k is a ratio of two
	// linear functions of k, monotone while cf2 keeps its sign, so its minimum over the span is at one of the two ends; both
	// ends are ESTIMATED (approximate reciprocal, the end of the chain in closed form: error ~1e-6 over <= 64 steps) and the
	// margin covers the estimate. A fragment fails when z - cur >= -0.0001 (fragthrd.cpp:227), certainly when z >= cur.
	const int n = x2 - x1;
	if(n <= PS_SPAN_BOUND_MAX)
	{
		const float nf = (float)n;
		const float cf2e = cf2 + nf * cf2Step, z0e = z0 + nf * zStep;
		const float zs = __fdividef(z0, cf2), ze = __fdividef(z0e, cf2e);
		if(cf2 > 0.0f && cf2e > 0.0f && zs == zs && ze == ze) o.zmin = fminf(zs, ze) - PS_HIZ_MARGIN;
	}
	else o.zmin = __int_as_float(0x7fc00000);               // too long for the closed-form estimate: the tile kernel estimates per tile
}

// STAGED: 0 = every slot read from global memory, 1 = the position slot (slot 0 of every vertex functor) staged in shared memory by
// one TMA bulk copy per block, 2 = every slot the functor reads staged (all of a block's vertex bytes in flight at once, phase B
// reads no global memory; 28 KB per block for DEF03)
template<int STAGED> __host__ __device__ constexpr uint32_t stageMask(uint32_t slots) { return 2 == STAGED ? slots : (1 == STAGED ? (slots & 1u) : 0u); }

// Sort-first (a rank renders a band of rows): every rank sees every triangle, so what a rank spends on triangles of other
// bands bounds the scaling. geom_precull computes only the three viewport y (the vertex functor's y and w: the compiler drops
// the rest) and the row range, and appends the triangles with a row in the band to one list (any order: tile lists are sorted
// by triangle id later, span records are allocated block by block anyway); geom_span<LISTED> then runs over that list on
// dense blocks. (Tried and dropped: the pre-cull inside the geometry kernel, a block queueing the survivors of 8 chunks into
// dense batches — 90 registers and an eighth of the blocks: 66 us against 29 + 38 us for an eighth of C2's rows, 204 against
// 34 + 130 us for half of them; and the pre-cull in front of phase A of every block without the queue: 145 us for half.)
#define PS_PRECULL_THREADS 512
template<class PROG>
__global__ void __launch_bounds__(PS_PRECULL_THREADS) geom_precull_kernel(const __grid_constant__ DrawParams P)
{
	constexpr int NV = PROG::NV;
	const uint32_t tri = blockIdx.x * PS_PRECULL_THREADS + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ uint32_t sWarpKeep[PS_PRECULL_THREADS / 32];
	__shared__ uint32_t sBase;
	bool keep = false;
	if(tri < P.ntris)
	{
		float ndcY[3];
#pragma unroll
		for(int i = 0; i < 3; i++)
		{
			VertexProcessorInput in;
#pragma unroll
			for(int s = 0; s < 16; s++)
				in.data[s] = (PROG::V::SLOTS >> s) & 1 ? P.slot[s] + (size_t)(tri * 3 + i) * P.stride[s] : nullptr;
			VertexProcessorOutput<NV> vo;
			PROG::V::process(in, vo, P);
			ndcY[i] = fmul(vo.position.y, fdiv(1.0f, vo.position.w));   // vertthrd.cpp:37-38
		}
		int firstRow, lastRow;
		keep = rowRangeOnly(P.vpH, P.halfH, ndcY, firstRow, lastRow) && lastRow >= P.band0 && firstRow < P.band1;
	}
	// one append per BLOCK: the list's counter is a single address, and 31 000 returning atomics on it (one per warp) were
	// what this kernel took 28 us for on C2 — the arithmetic and the 48 MB of positions need a third of that
	const uint32_t keepBallot = __ballot_sync(PS_FULL, keep);
	if(0 == lane) sWarpKeep[warp] = (uint32_t)__popc(keepBallot);
	__syncthreads();
	if(0 == warp)
	{
		const uint32_t c = lane < PS_PRECULL_THREADS / 32 ? sWarpKeep[lane] : 0u;
		uint32_t incl = c;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		if(lane < PS_PRECULL_THREADS / 32) sWarpKeep[lane] = incl - c;
		if(31 == lane) sBase = incl ? atomicAdd(P.workCount, incl) : 0u;
	}
	__syncthreads();
	if(keep) P.workList[sBase + sWarpKeep[warp] + (uint32_t)__popc(keepBallot & ((1u << lane) - 1))] = tri;
}

struct GeomAcc { unsigned rasterised, spans; unsigned long long frags; unsigned loInv, hi1; };

// One batch of up to 128 triangles (one per thread; `candidate` false: none) through phases A, S and B. `orig` is the
// triangle's place in the staged copy of the block's vertex range (STAGED only). Ends with a block barrier, so batches can
// follow each other in one block. `tri` indexes the draw's vertex streams and varyings; tri + idBias is the id the triangle
// carries through records, lists and headers (a batch of draws numbers its triangles across the draws).
template<class PROG, int STAGED>
PS_D void geomBatch(const DrawParams& P, uint32_t tri, uint32_t orig, bool candidate, const uint8_t* stage, const uint32_t* stageOff, GeomAcc& acc, const uint32_t idBias = 0)
{
	constexpr int NV = PROG::NV;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ TriHeader sHdr[PS_GEOM_THREADS];          // survivors, compacted (submission order kept)
	__shared__ uint32_t sRowBase[PS_GEOM_THREADS + 1];   // first (triangle, row) lane of each survivor
	__shared__ int sRow0[PS_GEOM_THREADS];
	__shared__ uint8_t sOrig[PS_GEOM_THREADS];           // survivor -> its thread of phase A
	__shared__ uint8_t sOwner[PS_SPAN_WINDOW];           // (triangle, row) lane of the window -> survivor
	__shared__ uint32_t sInfo[PS_SPAN_WINDOW];           // tile columns reached by that row: lo | hi << 12 | valid << 31
	__shared__ uint32_t sWarpAlive[PS_GEOM_THREADS / 32], sWarpRows[PS_GEOM_THREADS / 32];
	__shared__ uint32_t sTri[PS_GEOM_THREADS];           // survivor -> its triangle id in the draw
	// survivors PS_TALL_ROWS rows high or more (a 4096-row shadow-map triangle): their extent in tiles is reduced by the whole
	// block instead of being walked row by row by their own thread (min tx, max tx, first ty, last ty)
	__shared__ uint32_t sTallExt[PS_GEOM_THREADS][4];
	__shared__ uint32_t sTallList[PS_GEOM_THREADS];
	__shared__ uint32_t sTallCount;
	__shared__ uint32_t sSpanBase;
	const bool useDepth = 0 != (P.behavior & (PS_BEHAVIOR_TEST_DEPTH | PS_BEHAVIOR_UPDATE_DEPTH));
	const int limitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	const int limitY = useDepth ? P.depth.height - 1 : 0x7fffffff;

	// ---- A, thread = triangle (LISTED: of this rank's band, dense) --------------------------------------------------------------
	unsigned rasterised = 0;
	bool alive = false;
	int row0 = 0, nrows = 0;
	TriHeader h;
	if(candidate)
	{
		float ndcX[3], ndcY[3], rw[3], pz[3];
		F4 pos[3];
#pragma unroll
		for(int i = 0; i < 3; i++)
		{
			// processVertices, vertthrd.cpp:14-51 : all attached slots advance in lock-step, un-indexed. Only the position is used
			// here (the compiler drops the varyings and the loads that feed nothing else); survivors run the functor again in B.
			VertexProcessorInput in;
#pragma unroll
			for(int s = 0; s < 16; s++)
				in.data[s] = (PROG::V::SLOTS >> s) & 1 ? ((0 != ((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1)) ? stage + stageOff[s] + (size_t)(orig * 3 + i) * P.stride[s]
				                                                  : P.slot[s] + (size_t)(tri * 3 + i) * P.stride[s]) : nullptr;
			VertexProcessorOutput<NV> vo;
			PROG::V::process(in, vo, P);
			const float reciprocalW = fdiv(1.0f, vo.position.w);          // vertthrd.cpp:37 (true divide)
			pos[i] = f4muls(vo.position, reciprocalW);                     // :38 all four lanes
			ndcX[i] = pos[i].x; ndcY[i] = pos[i].y; pz[i] = pos[i].z; rw[i] = reciprocalW;
		}
		alive = true;
		if((P.behavior & PS_BEHAVIOR_FACE_CULLING) && isBackFace(pos[0], pos[1], pos[2], P.approx)) alive = false; // drawvao.cpp:46
		// drawvao.cpp:51-56 : the only "clipping" — drop the whole triangle
		if(pz[0] < -1.0f || pz[0] > 1.0f || pz[1] < -1.0f || pz[1] > 1.0f || pz[2] < -1.0f || pz[2] > 1.0f) alive = false;
		if(alive)
		{
			float vx[3], vy[3];
			const int code = setupTriangle(P.vpW, P.vpH, P.halfW, P.halfH, ndcX, ndcY, h, vx, vy);
			rasterised = code != 0;
			alive = 1 == code;
			if(alive)
			{
				// rows of this rank's band that the targets hold (sort-first: the other rows are another rank's)
				row0 = max((int)(h.rows & 0xffff), P.band0);
				const int row1 = min(min((int)(h.rows >> 16), P.band1 - 1), limitY);
				nrows = row1 - row0 + 1;
				if(nrows <= 0) { alive = false; nrows = 0; }
			}
			h.rw0 = rw[0]; h.rw1 = rw[1]; h.rw2 = rw[2];
			h.z0 = pz[0]; h.z1 = pz[1]; h.z2 = pz[2];
		}
	}
	// what phase B will read of this triangle's other vertex slots is asked into L2 now (phase S lies in between): survivors only
	if(alive && NV > 0)
	{
#pragma unroll
		for(int s = 0; s < 16; s++)
			if(((PROG::V::SLOTS >> s) & 1) && 0 == ((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1))
			{
				const uint8_t* a0 = P.slot[s] + (size_t)tri * 3 * P.stride[s];
				asm volatile("prefetch.global.L2 [%0];" :: "l"(a0));
				asm volatile("prefetch.global.L2 [%0];" :: "l"(a0 + 3 * P.stride[s] - 1));
			}
	}
	// compaction + exclusive scan of the row counts, both in thread order (dead threads add nothing, so a survivor's prefix
	// is the same in either numbering)
	const uint32_t aliveBallot = __ballot_sync(PS_FULL, alive);
	uint32_t rowsIncl = (uint32_t)nrows;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1)
	{
		const uint32_t t = __shfl_up_sync(PS_FULL, rowsIncl, d);
		if(lane >= d) rowsIncl += t;
	}
	if(31 == lane) { sWarpAlive[warp] = (uint32_t)__popc(aliveBallot); sWarpRows[warp] = rowsIncl; }
	__syncthreads();
	uint32_t slotBase = 0, rowBase = 0, nAlive = 0, totalRows = 0;
#pragma unroll
	for(int w = 0; w < PS_GEOM_THREADS / 32; w++)
	{
		if(w < warp) { slotBase += sWarpAlive[w]; rowBase += sWarpRows[w]; }
		nAlive += sWarpAlive[w]; totalRows += sWarpRows[w];
	}
	if(alive)
	{
		const uint32_t slot = slotBase + (uint32_t)__popc(aliveBallot & ((1u << lane) - 1));
		uint4* dst = (uint4*)&sHdr[slot];
		const uint4* src = (const uint4*)&h;
		dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
		sRowBase[slot] = rowBase + rowsIncl - (uint32_t)nrows;
		sRow0[slot] = row0;
		sOrig[slot] = (uint8_t)orig;
		sTri[slot] = tri + idBias;
	}
	if(0 == threadIdx.x)
	{
		sRowBase[nAlive] = totalRows;
		sSpanBase = totalRows ? atomicAdd(P.sp.count, totalRows) : 0u;
	}
	if(0 == threadIdx.x) sTallCount = 0;
	__syncthreads();
	const uint32_t spanBase = sSpanBase;
	const bool fits = totalRows <= P.sp.capacity && spanBase <= P.sp.capacity - totalRows;   // else: the plan kernel raises poison, the draw runs again
	{
		const bool tallK = threadIdx.x < nAlive && sRowBase[threadIdx.x + 1] - sRowBase[threadIdx.x] >= PS_TALL_ROWS;
		if(tallK)
		{
			sTallList[atomicAdd(&sTallCount, 1u)] = threadIdx.x;
			sTallExt[threadIdx.x][0] = 0xffffffffu; sTallExt[threadIdx.x][1] = 0; sTallExt[threadIdx.x][2] = 0xffffffffu; sTallExt[threadIdx.x][3] = 0;
		}
	}
	__syncthreads();
	const uint32_t nTall = sTallCount;

	// ---- S, lane = (triangle, row): one record per row -------------------------------------------------------------------------
	const uint32_t k = threadIdx.x;                    // B: this thread's survivor
	const uint32_t myRows0 = k < nAlive ? sRowBase[k] : 0u, myRows1 = k < nAlive ? sRowBase[k + 1] : 0u;
	unsigned spans = 0;
	unsigned long long frags = 0;
	const int myRow0 = k < nAlive ? sRow0[k] : 0, tyBase = myRow0 / PS_TILE;
	const bool myTall = myRows1 - myRows0 >= PS_TALL_ROWS;
	uint32_t trLo[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu }, trHi[4] = { 0, 0, 0, 0 };
	int tyLast = -1;
	for(uint32_t w0 = 0; w0 < totalRows; w0 += PS_SPAN_WINDOW)
	{
		const uint32_t wEnd = min(totalRows, w0 + PS_SPAN_WINDOW);
		{
			const uint32_t a = max(myRows0, w0), b = min(myRows1, wEnd);
			for(uint32_t s = a; s < b; s++) sOwner[s - w0] = (uint8_t)k;
		}
		__syncthreads();
		for(uint32_t s = w0 + threadIdx.x; s < wEnd; s += PS_GEOM_THREADS)
		{
			const uint32_t o = sOwner[s - w0];
			const int iy = sRow0[o] + (int)(s - sRowBase[o]);
			SpanOut sp;
			evalSpan(sHdr[o], iy, P.vpW, limitX, sp);
			spans += sp.counted ? 1u : 0u;
			uint32_t info = 0;
			if(sp.x1 <= sp.x2)
			{
				frags += (unsigned)(sp.x2 - sp.x1 + 1);
				info = 0x80000000u | (uint32_t)(sp.x1 / PS_TILE) | ((uint32_t)(sp.x2 / PS_TILE) << 12);
			}
			sInfo[s - w0] = info;
			if(fits)
			{
				int4* dst = (int4*)(P.sp.rec + spanBase + s);
				dst[0] = make_int4(sp.left, sp.right, __float_as_int(sp.zmin), (int)(sTri[o] | ((uint32_t)sp.edges << 24)));
				dst[1] = make_int4(__float_as_int(sp.cf2), __float_as_int(sp.cf2Step), __float_as_int(sp.z0), __float_as_int(sp.zStep));
			}
			if(sp.x2 - sp.x1 > PS_SPAN_BOUND_MAX)
			{
				// a long span: room for its chain marks (SpanStreams), its record on the list the mark kernels walk
				const uint32_t need = (uint32_t)(sp.x2 - sp.x1) / PS_MARK_STEP + 1u;
				const unsigned long long old = atomicAdd(P.sp.longCount, (1ull << 40) | need);
				if(fits && P.sp.markCap)
				{
					const uint32_t j = (uint32_t)(old >> 40);
					const unsigned long long at = old & ((1ull << 40) - 1);
					P.sp.markAt[spanBase + s] = at + need <= P.sp.markCap ? (uint32_t)at : 0xffffffffu;
					if(j < P.sp.capacity) P.sp.longList[j] = spanBase + s;
				}
			}
		}
		__syncthreads();
		// B's view of its triangle, row by row from what the lanes left in shared memory: the tile columns its spans reach in the
		// four tile rows from the one its first row lies in (rows further down fold into the last slot). (Tried instead, all slower
		// on C2: shared-memory atomics by every lane 0.196 -> 0.214 ms, a segmented reduction by shuffles 0.214, MATCH + group REDUX 0.297.)
		if(!myTall)
		{
			const uint32_t a = max(myRows0, w0), b = min(myRows1, wEnd);
			for(uint32_t s = a; s < b; s++)
			{
				const uint32_t info = sInfo[s - w0];
				if(0 == (info & 0x80000000u)) continue;
				const uint32_t lo = info & 0xfff, hi = (info >> 12) & 0xfff;
				const int ty = (myRow0 + (int)(s - myRows0)) / PS_TILE;
				const int rel = min(ty - tyBase, 3);
#pragma unroll
				for(int r = 0; r < 4; r++)
					if(rel == r) { trLo[r] = min(trLo[r], lo); trHi[r] = max(trHi[r], hi); }
				tyLast = ty;
			}
		}
		// tall survivors: the block reduces the extent of each one's rows of this window
		for(uint32_t q = 0; q < nTall; q++)
		{
			const uint32_t kt = sTallList[q];
			const uint32_t a = max(sRowBase[kt], w0), b = min(sRowBase[kt + 1], wEnd);
			uint32_t lo = 0xffffffffu, hi = 0, t0 = 0xffffffffu, t1 = 0;
			for(uint32_t s = a + threadIdx.x; s < b; s += PS_GEOM_THREADS)
			{
				const uint32_t info = sInfo[s - w0];
				if(0 == (info & 0x80000000u)) continue;
				const uint32_t ty = (uint32_t)((sRow0[kt] + (int)(s - sRowBase[kt])) / PS_TILE);
				lo = min(lo, info & 0xfff); hi = max(hi, (info >> 12) & 0xfff);
				t0 = min(t0, ty); t1 = max(t1, ty);
			}
			lo = __reduce_min_sync(PS_FULL, lo); hi = __reduce_max_sync(PS_FULL, hi);
			t0 = __reduce_min_sync(PS_FULL, t0); t1 = __reduce_max_sync(PS_FULL, t1);
			if(0 == lane && lo != 0xffffffffu)
			{
				atomicMin(&sTallExt[kt][0], lo); atomicMax(&sTallExt[kt][1], hi);
				atomicMin(&sTallExt[kt][2], t0); atomicMax(&sTallExt[kt][3], t1);
			}
		}
		__syncthreads();                               // the next window overwrites sOwner / sInfo; B reads the tall extents
	}
	int minTx = 0x7fffffff, maxTx = -1, tyFirst = -1;
	if(myTall)
	{
		if(sTallExt[k][0] != 0xffffffffu)
		{
			minTx = (int)sTallExt[k][0]; maxTx = (int)sTallExt[k][1]; tyFirst = (int)sTallExt[k][2]; tyLast = (int)sTallExt[k][3];
		}
	}
	else
	{
#pragma unroll
		for(int r = 0; r < 4; r++)
			if(trLo[r] != 0xffffffffu)
			{
				if(tyFirst < 0) tyFirst = tyBase + r;
				minTx = min(minTx, (int)trLo[r]); maxTx = max(maxTx, (int)trHi[r]);
			}
	}

	// ---- B, thread = survivor (dense): records for the shade kernel, tile lists -----------------------------------------------
	uint32_t binned = 0, rect0 = 0, rect1 = 0;
	bool bigPending = false;
	uint32_t wtri = 0;
	if(k < nAlive && maxTx >= 0)
	{
		const uint32_t orig = sOrig[k];
		wtri = sTri[k];
		binned = 1;
		{
			// 64-byte record as four 16-byte stores
			uint4* dst = (uint4*)(P.hdr + wtri);
			const uint4* src = (const uint4*)&sHdr[k];
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
		}
		{ TriSpan tsv; tsv.x = spanBase + myRows0; tsv.y = (uint32_t)sRow0[k] | ((uint32_t)(sRow0[k] + (int)(myRows1 - myRows0) - 1) << 16); P.sp.tri[wtri] = tsv; }
		if(NV > 0)
		{
			const uint32_t ltri = wtri - idBias;                          // the triangle's place in its own draw
			float4* vd = (float4*)(P.vary + (size_t)ltri * 3 * NV);
#pragma unroll 1
			for(int i = 0; i < 3; i++)
			{
				VertexProcessorInput in;
#pragma unroll
				for(int s = 0; s < 16; s++)
					in.data[s] = (PROG::V::SLOTS >> s) & 1 ? ((0 != ((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1)) ? stage + stageOff[s] + (size_t)(orig * 3 + i) * P.stride[s]
					                                                  : P.slot[s] + (size_t)(ltri * 3 + i) * P.stride[s]) : nullptr;
				VertexProcessorOutput<NV> vo;
				PROG::V::process(in, vo, P);
#pragma unroll
				for(int q = 0; q < NV; q++)
					vd[i * NV + q] = make_float4(vo.user[q].x, vo.user[q].y, vo.user[q].z, vo.user[q].w);
			}
		}
		const int tx0 = minTx, tx1 = maxTx, ty0 = tyFirst, ty1 = tyLast;
		rect0 = (uint32_t)tx0 | ((uint32_t)tx1 << 16);
		rect1 = (uint32_t)ty0 | ((uint32_t)ty1 << 16);
		if(myTall || tx1 - tx0 >= 8 || ty1 - tyBase >= 4) bigPending = true;  // its whole rectangle, appended by the warp further down
		else
		{
			// small triangle (the common case): exactly the tiles some span of it reaches (bit = (ty - tyBase) * 8 + tx - tx0). Up to
			// four appends are in flight at once (the atomics return the slot).
			uint32_t mask = 0;
#pragma unroll
			for(int r = 0; r < 4; r++)
				if(trLo[r] != 0xffffffffu)
				{
					const int a = (int)trLo[r] - tx0, b = (int)trHi[r] - tx0;
					mask |= ((2u << b) - (1u << a)) << (r * 8);                       // 0 <= a <= b <= 7
				}
			while(mask)
			{
				uint32_t tile[4], at[4];
#pragma unroll
				for(int u = 0; u < 4; u++)
				{
					tile[u] = 0xffffffffu;
					if(mask)
					{
						const int bit = __ffs(mask) - 1;
						mask &= mask - 1;
						tile[u] = (uint32_t)((tyBase + (bit >> 3)) * P.tilesX + tx0 + (bit & 7));
					}
				}
#pragma unroll
				for(int u = 0; u < 4; u++) at[u] = tile[u] != 0xffffffffu ? atomicAdd(&P.tl.fill[tile[u]], 1u) : 0xffffffffu;
#pragma unroll
				for(int u = 0; u < 4; u++) if(at[u] < P.tl.cap) P.tl.ids[(size_t)tile[u] * P.tl.cap + at[u]] = wtri;
			}
		}
	}
	// whole-rectangle triangles: the BLOCK appends one's id to its tiles, every thread four tiles per step, so that hundreds of the
	// (returning) atomics are in flight instead of one lane's one (a 4096^2 shadow map's ground quad is 2 x 65 536 tiles: a
	// millisecond per triangle when one warp did it, an atomic's round trip per 32 tiles)
	if(bigPending && P.tl.bigList)
	{
		// a very large rectangle goes to the grid-wide append (TileLists::bigList), if the list has room
		const uint32_t area = ((rect0 >> 16) - (rect0 & 0xffff) + 1u) * ((rect1 >> 16) - (rect1 & 0xffff) + 1u);
		if(area >= PS_BIG_AREA)
		{
			const uint32_t j = atomicAdd(P.tl.bigCount, 1u);
			if(j < P.tl.bigCap)
			{
				P.tl.bigList[3 * j] = rect0; P.tl.bigList[3 * j + 1] = rect1; P.tl.bigList[3 * j + 2] = wtri;
				bigPending = false;
			}
		}
	}
	if(bigPending)
	{
		const uint32_t q = atomicAdd(&sTallCount, 0x10000u) >> 16;   // (the tall list's counter: low half tall survivors, high half big ones)
		sTallList[q] = k;                                             // (the tall list is dead by now: phase S is over)
		sTallExt[k][0] = rect0; sTallExt[k][1] = rect1; sTallExt[k][2] = wtri;
	}
	__syncthreads();
	{
		const uint32_t nBig = sTallCount >> 16;
		for(uint32_t q = 0; q < nBig; q++)
		{
			const uint32_t kb = sTallList[q];
			const uint32_t r0 = sTallExt[kb][0], r1 = sTallExt[kb][1], t = sTallExt[kb][2];
			const int tx0 = (int)(r0 & 0xffff), tx1 = (int)(r0 >> 16), ty0 = (int)(r1 & 0xffff), ty1 = (int)(r1 >> 16);
			const uint32_t w = (uint32_t)(tx1 - tx0 + 1), total = w * (uint32_t)(ty1 - ty0 + 1);
			for(uint32_t base = threadIdx.x; base < total; base += 4 * PS_GEOM_THREADS)
			{
				uint32_t tile[4], at[4];
#pragma unroll
				for(int u = 0; u < 4; u++)
				{
					const uint32_t i = base + (uint32_t)u * PS_GEOM_THREADS;
					tile[u] = i < total ? (uint32_t)((ty0 + (int)(i / w)) * P.tilesX + tx0 + (int)(i % w)) : 0xffffffffu;
				}
#pragma unroll
				for(int u = 0; u < 4; u++) at[u] = tile[u] != 0xffffffffu ? atomicAdd(&P.tl.fill[tile[u]], 1u) : 0xffffffffu;
#pragma unroll
				for(int u = 0; u < 4; u++) if(at[u] < P.tl.cap) P.tl.ids[(size_t)tile[u] * P.tl.cap + at[u]] = t;
			}
		}
	}
	acc.rasterised += rasterised; acc.spans += spans; acc.frags += frags;
	acc.loInv = max(acc.loInv, binned ? ~((rect1 & 0xffff) * (unsigned)P.tilesX + (rect0 & 0xffff)) : 0u);
	acc.hi1 = max(acc.hi1, binned ? (rect1 >> 16) * (unsigned)P.tilesX + (rect0 >> 16) + 1u : 0u);
	__syncthreads();
}

// counters: one set of atomics per block, on the block's replica (per-draw fields: the plan kernel folds them)
PS_D void geomFinish(const DrawParams& P, const GeomAcc& acc)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ unsigned long long blockSums[PS_GEOM_THREADS / 32][3];
	__shared__ unsigned blockRange[PS_GEOM_THREADS / 32][2];
	{
		const unsigned a = __reduce_max_sync(PS_FULL, acc.loInv), b = __reduce_max_sync(PS_FULL, acc.hi1);
		if(0 == lane) { blockRange[warp][0] = a; blockRange[warp][1] = b; }
	}
	const unsigned long long r = warpSumU64(acc.rasterised), s = warpSumU64(acc.spans);
	unsigned long long f = acc.frags;
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) f += __shfl_xor_sync(PS_FULL, f, d);
	if(0 == lane) { blockSums[warp][0] = r; blockSums[warp][1] = s; blockSums[warp][2] = f; }
	__syncthreads();
	if(threadIdx.x < 3)
	{
		unsigned long long v = 0;
#pragma unroll
		for(int w = 0; w < PS_GEOM_THREADS / 32; w++) v += blockSums[w][threadIdx.x];
		DeviceStats* st = P.stats + (blockIdx.x & (PS_STATS_COPIES - 1));
		if(v) atomicAdd(0 == threadIdx.x ? &st->dRasterised : (1 == threadIdx.x ? &st->dSpans : &st->dTested), v);
	}
	else if(threadIdx.x < 5)
	{
		unsigned v = 0;
#pragma unroll
		for(int w = 0; w < PS_GEOM_THREADS / 32; w++) v = max(v, blockRange[w][threadIdx.x - 3]);
		DeviceStats* st = P.stats + (blockIdx.x & (PS_STATS_COPIES - 1));
		if(v) atomicMax(3 == threadIdx.x ? &st->tileLoInv : &st->tileHi1, v);
	}
}

template<class PROG, int STAGED, bool LISTED>
__global__ void __launch_bounds__(PS_GEOM_THREADS) geom_span_kernel(const __grid_constant__ DrawParams P)
{
	static_assert(!(STAGED && LISTED), "the list-driven form gathers its vertices from global memory");
	constexpr int NV = PROG::NV;
	const uint32_t tri0 = blockIdx.x * PS_GEOM_THREADS;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	extern __shared__ __align__(128) uint8_t stage[];
	__shared__ uint64_t stageBar;
	uint32_t stageOff[16];
	if(STAGED)
	{
		const uint32_t nt = min((uint32_t)PS_GEOM_THREADS, P.ntris - tri0);
		uint32_t off = 0;
#pragma unroll
		for(int s = 0; s < 16; s++)
		{
			stageOff[s] = off;
			if((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1) off += (PS_GEOM_THREADS * 3 * P.stride[s] + 127u) & ~127u;
		}
		if(0 == threadIdx.x) mbarInit(&stageBar, 1);
		__syncthreads();
		if(0 == threadIdx.x)
		{
			uint32_t total = 0;
#pragma unroll
			for(int s = 0; s < 16; s++)
				if((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1) total += (nt * 3 * P.stride[s] + 15u) & ~15u;
			mbarExpectTx(&stageBar, total);
#pragma unroll
			for(int s = 0; s < 16; s++)
				if((stageMask<STAGED>(PROG::V::SLOTS) >> s) & 1)
					bulkCopyG2S(stage + stageOff[s], P.slot[s] + (size_t)tri0 * 3 * P.stride[s], (nt * 3 * P.stride[s] + 15u) & ~15u, &stageBar);
		}
		mbarWait(&stageBar, 0);
	}

	// the triangle this thread takes through phase A
	const uint32_t orig = threadIdx.x;
	uint32_t tri = tri0 + threadIdx.x;
	bool candidate = tri < P.ntris;
	if(!STAGED && !LISTED && P.batchTrisPerBlock)
	{
		// a small draw: fewer triangles per block, more blocks (a block's rows and whole-rectangle appends are its own threads' work)
		tri = blockIdx.x * P.batchTrisPerBlock + threadIdx.x;
		candidate = threadIdx.x < P.batchTrisPerBlock && tri < P.ntris;
	}
	if(LISTED)
	{
		const uint32_t nList = *P.workCount;
		if(tri0 >= nList) return;                       // (the grid is sized for every triangle)
		candidate = tri < nList;
		tri = candidate ? P.workList[tri] : 0u;
	}

	GeomAcc acc = { 0, 0, 0, 0, 0 };
	geomBatch<PROG, STAGED>(P, tri, orig, candidate, stage, stageOff, acc);
	geomFinish(P, acc);
}

// ---- batches of small draws (ps3d_cuda.cu: Batch) ---------------------------------------------------------------------------
// Consecutive small draws into the same targets with the same behaviour bits share ONE pass: their triangles are numbered in
// submission order across the draws (every draw starts a new block of PS_GEOM_THREADS ids, so a block of ids belongs to one
// draw), one geometry launch per programme present fills the shared span records and tile lists, one plan / sort / raster pass
// resolves depth in that order (which is the order the draws would have run in: drawvao.cpp:78-96 ends before the next call
// starts), one shade launch per programme takes its own survivors. What differs from draw to draw — vertex streams, latched
// uniforms, textures — sits in global memory, one DrawParams per draw (its stream pointers and triangle count shifted by the
// draw's first id, so that the functors index them with the batch-wide id).
struct BatchView
{
	const DrawParams* items;    // per draw
	const uint32_t* blockDraw;  // per block of ids: draw | programme group << 16
	const uint32_t* blockList;  // the blocks of this launch's programme group
	uint32_t group;
};

template<class PROG>
__global__ void __launch_bounds__(PS_GEOM_THREADS) geom_span_multi_kernel(const BatchView B)
{
	const uint32_t gb = B.blockList[blockIdx.x];
	const DrawParams& P = B.items[B.blockDraw[gb] & 0xffffu];
	// a small draw spreads its triangles over many blocks (a block's rows are walked by its own threads: a dozen triangles that
	// cover a 4096-row target are a dozen blocks' work, not one's): block gb takes batchTrisPerBlock of them
	const uint32_t first = (gb - P.batchFirstBlock) * P.batchTrisPerBlock;
	const uint32_t tri = first + threadIdx.x;
	uint32_t stageOff[16];
#pragma unroll
	for(int s = 0; s < 16; s++) stageOff[s] = 0;
	GeomAcc acc = { 0, 0, 0, 0, 0 };
	geomBatch<PROG, 0>(P, tri, threadIdx.x, threadIdx.x < P.batchTrisPerBlock && tri < P.ntris, nullptr, stageOff, acc, gb * PS_GEOM_THREADS - first);
	geomFinish(P, acc);
}

// the draws' DrawParams from the launch's parameter space into the batch's table (a kernel, not a copy: it can be recorded into a
// captured frame together with its source), and the draws' blocks into the block tables. Block = draw of the pack.
#define PS_BATCH_PACK 7
struct ItemPack
{
	DrawParams item[PS_BATCH_PACK];
	uint32_t nBlocks[PS_BATCH_PACK], listAt[PS_BATCH_PACK], tag[PS_BATCH_PACK];
	uint32_t first;             // place of item[0] in the table
};
static_assert(sizeof(ItemPack) <= 32000, "kernel parameters end at 32 764 bytes");
__global__ void __launch_bounds__(256) batch_items_kernel(const __grid_constant__ ItemPack K, DrawParams* table, uint32_t* blockDraw, uint32_t* blockList)
{
	static_assert(0 == sizeof(DrawParams) % 8, "copied as 8-byte words");
	const uint32_t i = blockIdx.x;
	const uint2* src = (const uint2*)&K.item[i];
	uint2* d = (uint2*)(table + K.first + i);
	for(uint32_t w = threadIdx.x; w < sizeof(DrawParams) / 8; w += blockDim.x) d[w] = src[w];
	const uint32_t firstBlock = K.item[i].batchFirstBlock;
	for(uint32_t j = threadIdx.x; j < K.nBlocks[i]; j += blockDim.x) { blockDraw[firstBlock + j] = K.tag[i]; blockList[K.listAt[i] + j] = firstBlock + j; }
}

// the noted very large rectangles (TileLists::bigList), appended by the whole grid: thread = tile of a rectangle
__global__ void __launch_bounds__(256) tile_append_big_kernel(TileLists tl, int tilesX)
{
	const uint32_t n = min(*tl.bigCount, tl.bigCap);
	const uint32_t threads = gridDim.x * blockDim.x, me = blockIdx.x * blockDim.x + threadIdx.x;
	for(uint32_t e = 0; e < n; e++)
	{
		const uint32_t r0 = tl.bigList[3 * e], r1 = tl.bigList[3 * e + 1], t = tl.bigList[3 * e + 2];
		const int tx0 = (int)(r0 & 0xffff), tx1 = (int)(r0 >> 16), ty0 = (int)(r1 & 0xffff), ty1 = (int)(r1 >> 16);
		const uint32_t w = (uint32_t)(tx1 - tx0 + 1), total = w * (uint32_t)(ty1 - ty0 + 1);
		for(uint32_t i = me; i < total; i += threads)
		{
			const uint32_t tile = (uint32_t)((ty0 + (int)(i / w)) * tilesX + tx0 + (int)(i % w));
			const uint32_t at = atomicAdd(&tl.fill[tile], 1u);
			if(at < tl.cap) tl.ids[(size_t)tile * tl.cap + at] = t;
		}
	}
}

// ======================================================================================================================
// plan: lengths, verdict on the speculated capacities, tiles by descending list length, the draw's counters
// ======================================================================================================================

__global__ void __launch_bounds__(1024) tile_plan_kernel(TileLists tl, uint32_t ntiles, DeviceStats* stats, uint32_t* spanCount, uint32_t spanCap,
                                                        unsigned long long survivorCap, uint32_t listLimit, uint32_t* poison, DrawReport* report,
                                                        uint32_t* __restrict__ tileOrder, unsigned long long* longCount, uint32_t* longLatched)
{
	__shared__ uint32_t hist[256];                     // tiles per length class, longest lists first
	// one histogram per warp: most tiles of a frame fall into a handful of classes, and ~8000 atomics of a whole block on five
	// shared-memory addresses were most of this kernel's time (MATCH-aggregated adds were slower still: 14 -> 21 us)
	__shared__ uint32_t whist[32][256];
	__shared__ uint32_t longestS, nonEmptyS, pairsS;
	__shared__ unsigned long long boundS, foldS[3];
	__shared__ uint32_t rangeLoS, rangeHiS;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if(0 == threadIdx.x) { longestS = 0; pairsS = 0; }
	if(threadIdx.x < 256) hist[threadIdx.x] = 0;
	for(int q = threadIdx.x; q < 32 * 256; q += 1024) (&whist[0][0])[q] = 0;
	if(0 == warp)
	{
		unsigned long long b = 0, a0 = 0, a1 = 0, a2 = 0;
		unsigned loInv = 0, hi1 = 0;
		if(lane < PS_STATS_COPIES)
		{
			DeviceStats& st = stats[lane];
			b = st.fragBound; a0 = st.dRasterised; a1 = st.dSpans; a2 = st.dTested; loInv = st.tileLoInv; hi1 = st.tileHi1;
			st.fragBound = 0; st.dRasterised = 0; st.dSpans = 0; st.dTested = 0; st.tileLoInv = 0; st.tileHi1 = 0;
		}
#pragma unroll
		for(int d = 16; d > 0; d >>= 1)
		{
			b += __shfl_xor_sync(PS_FULL, b, d); a0 += __shfl_xor_sync(PS_FULL, a0, d);
			a1 += __shfl_xor_sync(PS_FULL, a1, d); a2 += __shfl_xor_sync(PS_FULL, a2, d);
		}
		loInv = __reduce_max_sync(PS_FULL, loInv); hi1 = __reduce_max_sync(PS_FULL, hi1);
		if(0 == lane)
		{
			boundS = a2;                               // every clamped fragment is tested: the survivors' upper bound
			foldS[0] = a0; foldS[1] = a1; foldS[2] = a2;
			rangeLoS = hi1 ? min(~loInv, ntiles) : ntiles; rangeHiS = min(hi1, ntiles);
			(void)b;
		}
	}
	__syncthreads();
	const uint32_t lo = rangeLoS, hi = max(rangeHiS, rangeLoS);
	uint32_t longest = 0, pairs = 0;
	// 8 tiles per thread per pass, all eight loads in flight at once (one dependent load per tile made this one-block kernel
	// 19 us of pure latency on C2's 8160 tiles); a pass covers 8192 tiles, the common case is a single pass whose counts stay
	// in registers for the ordering step below
	const bool onePass = hi - lo <= 8192;
	uint32_t keep[8];
	for(uint32_t b0 = lo; b0 < hi; b0 += 8192)
	{
		uint32_t v[8];
#pragma unroll
		for(int q = 0; q < 8; q++) { const uint32_t i = b0 + q * 1024 + threadIdx.x; v[q] = i < hi ? tl.fill[i] : 0u; }
#pragma unroll
		for(int q = 0; q < 8; q++)
		{
			const uint32_t i = b0 + q * 1024 + threadIdx.x;
			keep[q] = v[q];
			if(i < hi)
			{
				tl.len[i] = v[q];
				if(v[q])
				{
					tl.fill[i] = 0;
					longest = max(longest, v[q]); pairs += v[q];
				}
			}
			if(i < hi && v[q]) atomicAdd(&whist[warp][255u - min(v[q] >> 2, 255u)], 1u);
		}
	}
	longest = __reduce_max_sync(PS_FULL, longest);
	pairs = __reduce_add_sync(PS_FULL, pairs);
	if(0 == lane) { if(longest) atomicMax(&longestS, longest); if(pairs) atomicAdd(&pairsS, pairs); }
	__syncthreads();
	if(threadIdx.x < 256)
	{
		// class totals; whist[w][c] becomes warp w's first slot inside class c
		uint32_t run = 0;
		for(int w = 0; w < 32; w++) { const uint32_t c = whist[w][threadIdx.x]; whist[w][threadIdx.x] = run; run += c; }
		hist[threadIdx.x] = run;
	}
	__syncthreads();
	if(0 == warp)
	{
		uint32_t hh[8], sum = 0;
#pragma unroll
		for(int q = 0; q < 8; q++) { hh[q] = hist[lane * 8 + q]; sum += hh[q]; }
		uint32_t incl = sum;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		uint32_t run = incl - sum;
#pragma unroll
		for(int q = 0; q < 8; q++) { hist[lane * 8 + q] = run; run += hh[q]; }
		if(31 == lane) nonEmptyS = run;
	}
	__syncthreads();
	if(onePass)
	{
#pragma unroll
		for(int q = 0; q < 8; q++)
		{
			const uint32_t i = lo + q * 1024 + threadIdx.x;
			if(i < hi && keep[q])
			{
				const uint32_t cls = 255u - min(keep[q] >> 2, 255u);
				tileOrder[hist[cls] + atomicAdd(&whist[warp][cls], 1u)] = i;
			}
		}
	}
	else
	{
		for(uint32_t i = lo + threadIdx.x; i < hi; i += 1024)
		{
			const uint32_t v = tl.len[i];
			if(v) { const uint32_t cls = 255u - min(v >> 2, 255u); tileOrder[hist[cls] + atomicAdd(&whist[warp][cls], 1u)] = i; }
		}
	}
	if(0 == threadIdx.x)
	{
		tileOrder[ntiles] = nonEmptyS;
		const uint32_t lng = longestS, nspans = *spanCount;
		*spanCount = 0;
		// long spans and the chain marks they asked for (no verdict hangs on them: marks that did not fit are not kept, their spans
		// replay from the start; the host sizes the next draw's buffer by the report)
		const unsigned long long lc = *longCount;
		*longCount = 0;
		if(tl.bigCount) *tl.bigCount = 0;
		const unsigned long long marks = lc & ((1ull << 40) - 1);
		longLatched[0] = (uint32_t)min(lc >> 40, (unsigned long long)spanCap);
		longLatched[1] = (uint32_t)min(marks, 0xffffffffull);
		report->marks = (unsigned int)min(marks, 0xffffffffull);
		const unsigned long long bound = boundS;
		const uint32_t bad = (lng > tl.cap || lng > listLimit || nspans > spanCap || bound > survivorCap) ? 1u : 0u;
		if(!bad)
		{
			// the draw stands: its counters join the totals (replica 0)
			atomicAdd(&stats[0].triangles_rasterised, foldS[0]);
			atomicAdd(&stats[0].spans, foldS[1]);
			atomicAdd(&stats[0].fragments_tested, foldS[2]);
		}
		report->pairs = pairsS; report->longest = lng; report->spans = nspans; report->fragBound = bound; report->bad = bad;
		if(bad) report->sticky = 1;
		__threadfence_system();
		*poison = bad;
	}
}

// ======================================================================================================================
// sort-first composite over peer memory (NVLink): every rank renders its band straight into rank 0's colour target; what
// is left of the exchange step is two counters per frame. All counts live on the device (a captured frame replays them).
// ======================================================================================================================

struct PeerFlags
{
	// in RANK 0's memory, written by the ranks over NVLink
	unsigned int done[64];          // done[r]: frames rank r has finished writing
	unsigned int released[2];       // per display target: how often rank 0 has handed it out for a new frame
};
struct PeerCounters                 // in each rank's own memory
{
	unsigned int done;              // frames this rank has finished (rank 0: frames whose composite it has waited for)
	unsigned int released[2];       // how often this rank has taken (rank 0: handed out) each display target
};

PS_D unsigned int ldAcquireSys(const unsigned int* p)
{
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
PS_D void stReleaseSys(unsigned int* p, unsigned int v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// rank != 0, behind its frame: everything this rank stored into rank 0's target is ordered before the count (release, system scope)
__global__ void peer_signal_done_kernel(PeerCounters* mine, PeerFlags* rank0, int rank)
{
	const unsigned int v = ++mine->done;
	__threadfence_system();
	stReleaseSys(&rank0->done[rank], v);
}
// rank 0, behind its frame: the composite is complete when every rank has counted this frame
__global__ void peer_wait_done_kernel(PeerCounters* mine, PeerFlags* flags, int world)
{
	__shared__ unsigned int want;
	if(0 == threadIdx.x) want = ++mine->done;
	__syncthreads();
	const unsigned int v = want;
	if((int)threadIdx.x > 0 && (int)threadIdx.x < world)
		while((int)(ldAcquireSys(&flags->done[threadIdx.x]) - v) < 0) __nanosleep(100);
	__syncthreads();
}
// rank 0, in front of the first write of a frame into display target b: hands the target out
__global__ void peer_release_kernel(PeerCounters* mine, PeerFlags* flags, int b)
{
	const unsigned int v = ++mine->released[b];
	__threadfence_system();
	stReleaseSys(&flags->released[b], v);
}
// rank != 0, in front of its first write of a frame into rank 0's display target b: waits until rank 0 has handed it out
// (rank 0 does so behind the read-back of the frame before, so no rank overwrites an image still being read)
__global__ void peer_take_kernel(PeerCounters* mine, const PeerFlags* rank0, int b)
{
	const unsigned int v = ++mine->released[b];
	while((int)(ldAcquireSys(&rank0->released[b]) - v) < 0) __nanosleep(200);
}

// one warp per tile: bitonic sort of its list (<= PS_SORT_LIMIT ids) in shared memory; lists at tile * cap. (A block per tile with a
// block barrier per stage: 13.3 -> 10.6 us for a sort-first eighth, but 32 -> 44 us for the whole frame: dropped.)
// A warp's bitonic sort of 32 * E keys held in registers, E consecutive keys per lane (key i = lane * E + r): exchanges at a
// distance below E stay inside the lane, the others are one shuffle per key — no shared-memory round trip per stage (the
// shared-memory network below took 32 us on C2's 8160 lists of ~145 ids; this one is used for lists up to 512).
template<int E>
PS_D void warpBitonicSort(uint32_t (&v)[E], int lane)
{
#pragma unroll
	for(int kk = 2; kk <= 32 * E; kk <<= 1)
	{
#pragma unroll
		for(int j = kk >> 1; j > 0; j >>= 1)
		{
			if(j >= E)
			{
#pragma unroll
				for(int r = 0; r < E; r++)
				{
					const uint32_t o = __shfl_xor_sync(PS_FULL, v[r], j / E);
					const int i = lane * E + r;
					const bool up = 0 == (i & kk), lower = 0 == (i & j);
					v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
				}
			}
			else
			{
#pragma unroll
				for(int r = 0; r < E; r++)
					if(0 == (r & j))
					{
						const bool up = 0 == ((lane * E + r) & kk);
						const uint32_t x = v[r], y = v[r | j];
						if((x > y) == up) { v[r] = y; v[r | j] = x; }
					}
			}
		}
	}
}
template<int E>
PS_D void sortListInRegisters(uint32_t* list, uint32_t n, int lane)
{
	uint32_t v[E];
#pragma unroll
	for(int r = 0; r < E; r++) { const uint32_t i = (uint32_t)(lane * E + r); v[r] = i < n ? list[i] : 0xffffffffu; }
	warpBitonicSort<E>(v, lane);
#pragma unroll
	for(int r = 0; r < E; r++) { const uint32_t i = (uint32_t)(lane * E + r); if(i < n) list[i] = v[r]; }
}

__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK) tile_list_sort_cap_kernel(TileLists tl, uint32_t ntiles, const uint32_t* __restrict__ poison,
                                                                                    const uint32_t* __restrict__ tileOrder)
{
	__shared__ uint32_t buf[PS_WARPS_PER_BLOCK][PS_SORT_LIMIT];
	if(*poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t slot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(slot >= tileOrder[ntiles]) return;              // tiles with a list
	const uint32_t tile = tileOrder[slot];
	const uint32_t n = tl.len[tile];
	uint32_t* list = tl.ids + (size_t)tile * tl.cap;
	if(n < 2 || n > PS_SORT_LIMIT) return;
	uint32_t* a = buf[w];
	uint32_t P2 = 2;
	while(P2 < n) P2 <<= 1;
	if(P2 <= 512)
	{
		// (a list already in order — a tile reached by one geometry block only — is left alone)
		bool sorted = true;
		for(uint32_t i = lane + 1; i < n; i += 32) if(list[i - 1] > list[i]) sorted = false;
		if(__all_sync(PS_FULL, sorted)) return;
		if(P2 <= 32) sortListInRegisters<1>(list, n, lane);
		else if(P2 <= 64) sortListInRegisters<2>(list, n, lane);
		else if(P2 <= 128) sortListInRegisters<4>(list, n, lane);
		else if(P2 <= 256) sortListInRegisters<8>(list, n, lane);
		else sortListInRegisters<16>(list, n, lane);
		return;
	}
	bool sorted = true;
	for(uint32_t i = lane; i < P2; i += 32)
	{
		const uint32_t v = i < n ? list[i] : 0xffffffffu;
		a[i] = v;
		if(i > 0 && i < n && list[i - 1] > v) sorted = false;
	}
	__syncwarp();
	if(__all_sync(PS_FULL, sorted)) return;
	for(uint32_t kk = 2; kk <= P2; kk <<= 1)
		for(uint32_t j = kk >> 1; j > 0; j >>= 1)
		{
			for(uint32_t t = lane; t < (P2 >> 1); t += 32)
			{
				const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				const uint32_t x = a[i], y = a[i | j];
				const bool up = 0 == (i & kk);
				if((x > y) == up) { a[i] = y; a[i | j] = x; }
			}
			__syncwarp();
		}
	for(uint32_t i = lane; i < n; i += 32) list[i] = a[i];
}

// ======================================================================================================================
// tile raster + depth over span records
// ======================================================================================================================

#define PS_SLOTS 64            // ring of spans waiting for the pixel phase
#define PS_CAND_MAX 512        // candidates of one chunk: 32 triangles x 16 rows

struct RasterSmem2
{
	float depth[PS_TILE * PS_TILE];
	uint32_t lastIdx[PS_TILE * PS_TILE];   // 1 + stream index of the last survivor of each pixel
	float segMax[32];                      // upper bound of the depth of every 8-pixel row segment (index = row * 2 + segment)
	uint32_t triBase[32];                  // candidate s of the chunk: record index = triBase[owner] + s
	int rowOff[32];                        //                           raster row   = rowOff[owner] + s
	uint8_t owner[PS_CAND_MAX];
	// spans that reached the tile and survived the whole-span reject, in list order (order = submission order on every row)
	float sCf2[PS_SLOTS], sCf2Step[PS_SLOTS], sZ[PS_SLOTS], sZStep[PS_SLOTS];   // chains at the span's first pixel inside the tile
	uint32_t sSpan[PS_SLOTS];
	uint32_t sMisc[PS_SLOTS];              // xs | row << 4 | (len - 1) << 8 (tile-relative)
	uint32_t pixBase[32];                  // pass: first pixel index of each slot
	// staging queue of survivors
	uint32_t qSpan[PS_RQCAP], qXY[PS_RQCAP];   // px | row << 4 while queued
	float qInv[PS_RQCAP];
	uint32_t chainUnit, chainLast;             // blending draws: this warp's unit, 1 + first slot of the last group it appended
	uint32_t chainPad[2];                      // (the array of these is walked with 16-byte accesses)
};
static_assert(0 == sizeof(RasterSmem2) % 16, "one per warp, 16-byte accesses");

struct RasterCtx2
{
	int lane, tx0, ty0;
	bool testDepth, updateDepth;
	uint32_t ltMask;
	uint32_t qCount;           // warp-uniform
	unsigned survived;
	bool depthWrote;           // per lane
	uint32_t reserved;         // lane 0: first stream index of the group in flight (the atomic's answer, not waited for until flushEnd)
	bool pending;              // warp-uniform: a group sits in registers, its slots being reserved
	uint32_t pSpan, pXY;       // per lane: that group's record
	float pInv;
};
#define PS_SV_HOLE 0xffffffffu     // survivor stream: a reserved slot that no survivor took (the last, partial group of a tile)

// A full group of 32 survivors leaves the queue in two steps: flushBegin takes the records into registers and issues the atomic
// that reserves their stream slots; flushEnd — at the next flush or the end of the tile — stores them. The atomic's ~1 us round
// trip (7 % of this kernel's stall samples when waited for on the spot) passes while the next pixels are tested, and the stream
// stays dense. (Tried: slots reserved 32 at a time ahead of need, the last partial group padded with holes the shade kernel
// skips — same gain here, but 4 % more warps for the shade kernel.)
// a blending draw (SurvivorStream2::next): one lane hangs the group of `n` survivors at stream slot `base` onto its unit's chain
PS_D void chainGroup(const SurvivorStream2& Q, RasterSmem2& S, uint32_t base, uint32_t n)
{
	if(base >= Q.capacity) return;                     // (the draw is poisoned: nothing reads the chain)
	Q.next[base] = 0;
	if(S.chainLast) Q.next[S.chainLast - 1] = base + 1;
	else Q.chain[2 * S.chainUnit] = base + 1;
	Q.chain[2 * S.chainUnit + 1] = n;
	S.chainLast = base + 1;
}

PS_D void flushEnd(const SurvivorStream2& Q, RasterSmem2& S, RasterCtx2& C)
{
	if(!C.pending) return;                             // (warp-uniform)
	C.pending = false;
	const uint32_t i = __shfl_sync(PS_FULL, C.reserved, 0) + (uint32_t)C.lane;
	if(i < Q.capacity)
	{
		Q.span[i] = C.pSpan; Q.inv[i] = C.pInv;
		Q.xy[i] = (uint32_t)(C.tx0 + (int)(C.pXY & 15)) | ((uint32_t)(C.ty0 + (int)((C.pXY >> 4) & 15)) << 13);
	}
	// queue order = submission order inside a pixel and a warp's reservations grow with time: the latest record has the highest index
	atomicMax(&S.lastIdx[C.pXY & 0xff], i + 1);
	if(Q.next && 0 == C.lane) chainGroup(Q, S, i, 32u);      // (lane 0's slot is the group's first)
}
PS_D void flushBegin(const SurvivorStream2& Q, RasterSmem2& S, RasterCtx2& C)
{
	if(nullptr == Q.span) return;                      // a draw whose fragment functor does nothing keeps no survivors (they are counted)
	flushEnd(Q, S, C);
	C.pSpan = S.qSpan[C.lane]; C.pXY = S.qXY[C.lane]; C.pInv = S.qInv[C.lane];
	if(0 == C.lane) C.reserved = atomicAdd(Q.count, 32u);
	C.pending = true;
}
// the last, partial group of a tile: n < 32 records, reserved and stored on the spot
PS_D void flushRest(const SurvivorStream2& Q, RasterSmem2& S, RasterCtx2& C, uint32_t n)
{
	if(nullptr == Q.span) return;
	flushEnd(Q, S, C);
	const int lane = C.lane;
	uint32_t base = 0;
	if(0 == lane) base = atomicAdd(Q.count, n);
	base = __shfl_sync(PS_FULL, base, 0);
	if((uint32_t)lane < n)
	{
		const uint32_t i = base + lane;
		const uint32_t m = S.qXY[lane];
		if(i < Q.capacity)
		{
			Q.span[i] = S.qSpan[lane]; Q.inv[i] = S.qInv[lane];
			Q.xy[i] = (uint32_t)(C.tx0 + (int)(m & 15)) | ((uint32_t)(C.ty0 + (int)((m >> 4) & 15)) << 13);
		}
		atomicMax(&S.lastIdx[m & 0xff], i + 1);
	}
	if(Q.next && 0 == lane && n) chainGroup(Q, S, base, n);
	__syncwarp();
}

// The pixel phase over ring slots [head, head + n), n <= 32: lane = pixel of a span, dense; spans in slot order, pixels left
// to right. interpolateNextStep (interp.cpp:82-92) + the depth rule (fragthrd.cpp:217-237).
PS_D void pixelPass(const SurvivorStream2& Q, RasterSmem2& S, RasterCtx2& C, uint32_t head, uint32_t n)
{
	const int lane = C.lane;
	int len = 0;
	if((uint32_t)lane < n) len = (int)((S.sMisc[(head + lane) & (PS_SLOTS - 1)] >> 8) & 15) + 1;
	uint32_t pincl = (uint32_t)len;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1)
	{
		const uint32_t t = __shfl_up_sync(PS_FULL, pincl, d);
		if(lane >= d) pincl += t;
	}
	const int myBase = (int)(pincl - (uint32_t)len);
	const uint32_t totalPix = __shfl_sync(PS_FULL, pincl, 31);
	S.pixBase[lane] = (uint32_t)myBase;
	__syncwarp();
	int startedBefore = 0;                             // slots whose first pixel lies before this batch (warp-uniform)
	for(uint32_t p0 = 0; p0 < totalPix; p0 += 32)
	{
		// slot of pixel p = (number of slots starting at or before p) - 1 : one bit per slot start inside the batch
		const int dStart = myBase - (int)p0;
		const uint32_t starts = __reduce_or_sync(PS_FULL, (len > 0 && dStart >= 0 && dStart < 32) ? 1u << dStart : 0u);
		const uint32_t p = p0 + lane;
		const bool act = p < totalPix;
		const int slot = startedBefore + __popc(starts & (C.ltMask | (1u << lane))) - 1;
		startedBefore += __popc(starts);
		uint32_t pix = 0x1000u + lane, span = 0, pos = 0;
		int kk = 0;
		if(act)
		{
			pos = (head + (uint32_t)slot) & (PS_SLOTS - 1);
			kk = (int)(p - S.pixBase[slot]);
			const uint32_t smisc = S.sMisc[pos];
			span = S.sSpan[pos];
			pix = (((smisc >> 4) & 15) << 4) | ((smisc & 15) + (uint32_t)kk);
		}
		// fragments of one pixel are tested in lane order = submission order (§9.7). (The match is issued before the chain replay
		// and the divide, which do not depend on it: its latency was 9 % of this kernel's stall samples.)
		const uint32_t peers = __match_any_sync(PS_FULL, pix);
		float z = 0, inv = 0;
		if(act)
		{
			float c2 = S.sCf2[pos], zz = S.sZ[pos];
			const float c2Step = S.sCf2Step[pos], zzStep = S.sZStep[pos];
#pragma unroll 1
			for(int j = 0; j < kk; j++)
			{
				c2 = fadd(c2, c2Step);
				zz = fadd(zz, zzStep);
			}
			// interpolateNextStep, interp.cpp:82-92
			inv = fdiv(1.0f, c2);
			z = fmul(zz, inv);
		}
		const int rank = __popc(peers & C.ltMask);
		const int maxRank = __reduce_max_sync(PS_FULL, act ? rank : 0);
		bool pass = false;
#pragma unroll 1
		for(int r = 0; r <= maxRank; r++)
		{
			if(act && rank == r)
			{
				const float cur = C.testDepth ? S.depth[pix] : 1.0f;             // fragthrd.cpp:217-225
				if(-1.0f < z && fsub(z, cur) < -0.0001f)                         // fragthrd.cpp:227
				{
					pass = true;
					// no functor on this path discards, so the depth write does not wait for the shading (fragthrd.cpp:234-237)
					if(C.updateDepth) { S.depth[pix] = z; C.depthWrote = true; }
				}
			}
			__syncwarp();
		}
		const uint32_t b = __ballot_sync(PS_FULL, pass);
		if(pass)
		{
			const uint32_t q = C.qCount + __popc(b & C.ltMask);
			S.qSpan[q] = span; S.qXY[q] = pix; S.qInv[q] = inv;
		}
		C.qCount += __popc(b);
		__syncwarp();
		if(C.qCount >= 32)
		{
			flushBegin(Q, S, C);
			C.survived += 32;
			const uint32_t rem = C.qCount - 32;
			uint32_t a = 0, d = 0; float f = 0;
			if((uint32_t)lane < rem) { a = S.qSpan[32 + lane]; d = S.qXY[32 + lane]; f = S.qInv[32 + lane]; }
			__syncwarp();
			if((uint32_t)lane < rem) { S.qSpan[lane] = a; S.qXY[lane] = d; S.qInv[lane] = f; }
			C.qCount = rem;
			__syncwarp();
		}
	}
}

// MARKS: long spans start their chains from the nearest chain mark (SpanStreams::markZ, span_mark_depth_kernel)
template<int MINB, bool MARKS = false>
__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK, MINB) tile_raster_span_kernel(const __grid_constant__ DrawParams P, const SurvivorStream2 Q, int parts)
{
	__shared__ __align__(16) RasterSmem2 smem[PS_WARPS_PER_BLOCK];
	if(*P.poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	// A warp walks its tile's list as one dependent chain: rows are independent in this rasteriser, so when there are fewer
	// tiles than the GPU has warp slots a tile is cut into `parts` (1, 2 or 4) groups of PS_TILE / parts rows, one warp each.
	const int warpSlot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(warpSlot >= (int)P.tileOrder[P.tilesX * P.tilesY] * parts) return;   // tiles with a list
	const int tile = (int)P.tileOrder[warpSlot / parts];   // longest lists first (tile_plan_kernel)
	const int rowsPer = PS_TILE / parts, partRow0 = (warpSlot % parts) * rowsPer;
	const uint32_t listLen = P.tl.len[tile];
	if(0 == listLen) return;
	const uint32_t* list = P.tl.ids + (size_t)tile * P.tl.cap;
	RasterSmem2& S = smem[w];

	const int tx0 = (tile % P.tilesX) * PS_TILE, ty0 = (tile / P.tilesX) * PS_TILE;
	const int rr = lane >> 1, seg = lane & 1;          // staging / write-back ownership: row rr, pixels [sx0, sx0+7]
	const int y = ty0 + rr, sx0 = tx0 + seg * PS_SEG;
	const bool testDepth = 0 != (P.behavior & PS_BEHAVIOR_TEST_DEPTH);
	const bool updateDepth = 0 != (P.behavior & PS_BEHAVIOR_UPDATE_DEPTH);
	const bool useDepth = testDepth || updateDepth;
	const bool mine = rr >= partRow0 && rr < partRow0 + rowsPer;   // rows of the tile this warp owns (staging, write-back)
	const bool depthRowOk = y < P.depth.height;
	uint8_t* depthRow = P.depth.ptr + (size_t)(P.depth.topDown ? P.depth.height - 1 - y : y) * P.depth.scanline;
	// fbo.cpp:98-110: a lane's 8-pixel segment is 32 contiguous bytes of one depth row: two 128-bit loads when the target allows
	const bool vec = 0 == (((uintptr_t)P.depth.ptr | (uintptr_t)P.depth.scanline) & 15) && sx0 + PS_SEG <= P.depth.width;
	{
		float4 lo = make_float4(1.0f, 1.0f, 1.0f, 1.0f), hi = lo;
		if(mine && useDepth && depthRowOk)
		{
			if(vec)
			{
				const float4* src = (const float4*)(depthRow + (size_t)sx0 * 4);
				lo = src[0]; hi = src[1];
			}
			else
			{
				float v[PS_SEG];
#pragma unroll
				for(int i = 0; i < PS_SEG; i++) v[i] = sx0 + i < P.depth.width ? *(const float*)(depthRow + (size_t)(sx0 + i) * 4) : 1.0f;
				lo = make_float4(v[0], v[1], v[2], v[3]); hi = make_float4(v[4], v[5], v[6], v[7]);
			}
		}
		float4* d = (float4*)&S.depth[rr * PS_TILE + seg * PS_SEG];
		d[0] = lo; d[1] = hi;
		uint4* li = (uint4*)&S.lastIdx[rr * PS_TILE + seg * PS_SEG];
		li[0] = make_uint4(0, 0, 0, 0); li[1] = make_uint4(0, 0, 0, 0);
	}
	__syncwarp();

	if(Q.next && 0 == lane) { S.chainUnit = (uint32_t)warpSlot; S.chainLast = 0; Q.chain[2 * warpSlot] = 0; }
	RasterCtx2 C;
	C.lane = lane; C.tx0 = tx0; C.ty0 = ty0;
	C.testDepth = testDepth; C.updateDepth = updateDepth;
	C.ltMask = (1u << lane) - 1;
	C.qCount = 0; C.survived = 0; C.depthWrote = false;
	C.reserved = 0; C.pending = false; C.pSpan = 0; C.pXY = 0; C.pInv = 0.0f;
	const int tileX1 = tx0 + PS_TILE - 1;
	const int depthLimitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	uint32_t head = 0, waiting = 0;                    // ring of spans waiting for the pixel phase (warp-uniform)

	// The list walk is a chain of dependent gathers (list -> per-triangle record index -> span records) whose targets were
	// written by another kernel and mostly left L2 since: every level is fetched one step ahead of its use — the ids two
	// chunks ahead, the per-triangle words one chunk ahead, the span records one candidate pass ahead.
	uint32_t idAhead = 0;
	uint2 tsAhead = make_uint2(0, 0);
	if((uint32_t)lane < listLen) tsAhead = __ldg((const uint2*)&P.sp.tri[list[lane]]);
	if(32u + lane < listLen) idAhead = list[32u + lane];

	for(uint32_t chunk = 0; chunk < listLen; chunk += 32)
	{
		// ---- lane = triangle of the chunk: where its records are, which of its rows lie in this warp's rows ----
		const uint32_t li = chunk + lane;
		const uint2 ts = tsAhead;
		if(li + 32 < listLen) tsAhead = __ldg((const uint2*)&P.sp.tri[idAhead]);
		if(li + 64 < listLen) idAhead = list[li + 64];
		int nrows = 0, ra = 0;
		uint32_t recBase = 0;
		if(li < listLen)
		{
			const int r0 = (int)(ts.y & 0xffff), r1 = (int)(ts.y >> 16);
			ra = max(r0, ty0 + partRow0);
			const int rb = min(r1, ty0 + partRow0 + rowsPer - 1);
			nrows = rb >= ra ? rb - ra + 1 : 0;
			recBase = ts.x + (uint32_t)(ra - r0);
		}
		uint32_t incl = (uint32_t)nrows;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		const uint32_t base = incl - (uint32_t)nrows;
		const uint32_t total = __shfl_sync(PS_FULL, incl, 31);
		S.triBase[lane] = recBase - base;
		S.rowOff[lane] = ra - (int)base;
		for(int q = 0; q < nrows; q++) S.owner[base + q] = (uint8_t)lane;
		// depths only decrease while a draw tests them, so a bound refreshed once per chunk stays an upper bound
		if(testDepth)
		{
			const float4 a = *(const float4*)&S.depth[lane * PS_SEG], b = *(const float4*)&S.depth[lane * PS_SEG + 4];
			S.segMax[lane] = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
		}
		__syncwarp();

		// every record of the chunk is asked into L2 now (no register holds a prefetch); the loads below, one pass ahead of their
		// use, then meet L2 latency instead of DRAM's
		for(uint32_t s = (uint32_t)lane; s < total; s += 32)
		{
			const SpanRec* r = P.sp.rec + (S.triBase[S.owner[s]] + s);
			asm volatile("prefetch.global.L2 [%0];" :: "l"(r));
		}
		// the first pass's records
		int4 recA = make_int4(1, 0, 0, 0), recD = make_int4(0, 0, 0, 0);
		int rowNext = 0;
		uint32_t idxNext = 0;
		if((uint32_t)lane < total)
		{
			const uint32_t o = S.owner[lane];
			idxNext = S.triBase[o] + (uint32_t)lane;
			rowNext = S.rowOff[o] + lane - ty0;
			const int4* src = (const int4*)(P.sp.rec + idxNext);
			recA = __ldg(src); recD = __ldg(src + 1);
		}
		for(uint32_t s0 = 0; s0 < total; s0 += 32)
		{
			// ---- lane = candidate (triangle, row): its record, clipped to the tile; whole-span reject; chains to the tile edge ----
			const uint32_t s = s0 + lane;
			const int4 A = recA, D = recD;
			const int row = rowNext;
			const uint32_t idx = idxNext;
			recA = make_int4(1, 0, 0, 0);
			if(s + 32 < total)
			{
				const uint32_t o = S.owner[s + 32];
				idxNext = S.triBase[o] + s + 32;
				rowNext = S.rowOff[o] + (int)(s + 32) - ty0;
				const int4* src = (const int4*)(P.sp.rec + idxNext);
				recA = __ldg(src); recD = __ldg(src + 1);
			}
			bool ok = false;
			float cf2 = 0, cf2Step = 0, z0 = 0, zStep = 0;
			uint32_t misc = 0;
			{
				const int x1 = A.x < 0 ? 0 : A.x;                           // RESULT_ROW::leftClamped
				const int x2 = min(A.y >= P.vpW ? P.vpW - 1 : A.y, depthLimitX);   // RESULT_ROW::rightClamped (+ the depth target's width)
				const int xs = max(x1, tx0), xe = min(x2, tileX1);
				ok = x1 <= x2 && xs <= xe;                                  // (lanes beyond the chunk's candidates hold the empty record 1, 0)
				const float zmin = __int_as_float(A.z);
				float bound = 0.0f;
				if(ok && testDepth)
				{
					// conservative whole-span reject against the segments the span's pixels in this tile lie in (NaN: no bound)
					bound = fmaxf(S.segMax[row * 2 + ((xs - tx0) >> 3)], S.segMax[row * 2 + ((xe - tx0) >> 3)]);
					if(zmin >= bound) ok = false;
				}
				if(ok)
				{
					cf2 = __int_as_float(D.x); cf2Step = __int_as_float(D.y); z0 = __int_as_float(D.z); zStep = __int_as_float(D.w);
					// the k-th pixel's value is k rounded additions from the span start (§9.6): replay them up to the tile
					int xr = x1;
					if(MARKS && zmin != zmin && xs - x1 >= PS_MARK_STEP)
					{
						const uint32_t at = __ldg(P.sp.markAt + idx);
						if(at != 0xffffffffu)
						{
							const int k = (xs - x1) / PS_MARK_STEP;
							const float2 m = __ldg((const float2*)(P.sp.markZ + at + k));
							cf2 = m.x; z0 = m.y;
							xr = x1 + k * PS_MARK_STEP;
						}
					}
#pragma unroll 1
					for(int x = xr; x < xs; x++)
					{
						cf2 = fadd(cf2, cf2Step);
						z0 = fadd(z0, zStep);
					}
					if(testDepth && zmin != zmin)
					{
						// long spans carry no bound: estimate both ends of the part inside the tile (as evalSpan does for short ones)
						const float nf = (float)(xe - xs);
						const float cf2e = cf2 + nf * cf2Step, z0e = z0 + nf * zStep;
						const float zs = __fdividef(z0, cf2), ze = __fdividef(z0e, cf2e);
						if(cf2 > 0.0f && cf2e > 0.0f && zs - PS_HIZ_MARGIN >= bound && ze - PS_HIZ_MARGIN >= bound) ok = false;
					}
					misc = (uint32_t)(xs - tx0) | ((uint32_t)row << 4) | ((uint32_t)(xe - xs) << 8);
				}
			}
			const uint32_t okb = __ballot_sync(PS_FULL, ok);
			if(ok)
			{
				const uint32_t pos = (head + waiting + __popc(okb & C.ltMask)) & (PS_SLOTS - 1);
				S.sCf2[pos] = cf2; S.sCf2Step[pos] = cf2Step; S.sZ[pos] = z0; S.sZStep[pos] = zStep;
				S.sSpan[pos] = idx; S.sMisc[pos] = misc;
			}
			waiting += __popc(okb);
			__syncwarp();
			if(waiting >= 32)
			{
				pixelPass(Q, S, C, head, 32);
				head = (head + 32) & (PS_SLOTS - 1);
				waiting -= 32;
			}
		}
		__syncwarp();
	}
	if(waiting) pixelPass(Q, S, C, head, waiting);
	flushRest(Q, S, C, C.qCount);
	C.survived += C.qCount;

	// ---- write back: depth tile (128-bit stores), and the flag of each pixel's last survivor
	if(__any_sync(PS_FULL, C.depthWrote) && depthRowOk && mine)
	{
		const float4* d = (const float4*)&S.depth[rr * PS_TILE + seg * PS_SEG];
		if(vec)
		{
			float4* dst = (float4*)(depthRow + (size_t)sx0 * 4);
			dst[0] = d[0]; dst[1] = d[1];
		}
		else
		{
#pragma unroll
			for(int i = 0; i < PS_SEG; i++)
				if(sx0 + i < P.depth.width) *(float*)(depthRow + (size_t)(sx0 + i) * 4) = S.depth[rr * PS_TILE + seg * PS_SEG + i];
		}
	}
	if(C.survived && mine)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
		{
			const uint32_t last = S.lastIdx[rr * PS_TILE + seg * PS_SEG + i];
			if(last && last - 1 < Q.capacity) Q.xy[last - 1] = (uint32_t)(sx0 + i) | ((uint32_t)y << 13) | PS_SV_WINNER;
		}
	}
	if(0 == lane && C.survived) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_shaded, (unsigned long long)C.survived);   // every survivor is shaded exactly once by shade_span_kernel
}

#define PS_SHADE_THREADS 128

// lane = survivor record, any order: varyings (interp.cpp:26-92), fragment functor (fragthrd.cpp:231), the pixel's last survivor
// stores its colour. A survivor's inputs sit behind three dependent gathers (stream -> span record -> triangle header and 3 x NV
// varyings): the stream words are read two iterations ahead and the span record one ahead, so only the last level's latency
// is exposed. (Tried and dropped: that level staged through shared memory by 16-byte asynchronous copies, all issued together —
// LDGSTS neither merges the lanes that name the same triangle nor uses L1, the kernel went 0.205 -> 0.355 ms.)
// The varyings' half of interpolateStartAndStep (interp.cpp:26-80) for one span: the chain's first value (at the clamped start x1,
// which is returned) and its step.
template<class PROG>
PS_D int varyingChainStart(const uint4& q0, const uint4& q1, const uint4& q2, const F4* v, int left, int right, int e, int y,
                           F4* vStart, F4* vStep)
{
	constexpr int NV = PROG::NV;
	const float vx[3] = { __uint_as_float(q0.x), __uint_as_float(q0.z), __uint_as_float(q1.x) };
	const float vy[3] = { __uint_as_float(q0.y), __uint_as_float(q0.w), __uint_as_float(q1.y) };
	const float rw0 = __uint_as_float(q1.z), rw1 = __uint_as_float(q1.w), rw2 = __uint_as_float(q2.x);
	float cl[3], cr[3];
	edgeContrib(vx, vy, e & 3, (e >> 2) & 3, (float)left, (float)y, cl);
	edgeContrib(vx, vy, (e >> 4) & 3, (e >> 6) & 3, (float)right, (float)y, cr);
	cl[0] = fmul(cl[0], rw0); cl[1] = fmul(cl[1], rw1); cl[2] = fmul(cl[2], rw2);
	cr[0] = fmul(cr[0], rw0); cr[1] = fmul(cr[1], rw1); cr[2] = fmul(cr[2], rw2);
	const int stepCount = right - left;
	const int x1 = left < 0 ? 0 : left;
	const int skip = x1 - left;
	// every varying is an independent float4 (the IP's methods are per-field loops, tex1light1.cpp:60-135)
	typedef InterpolationProcessorVec4<1> IP1;
	typedef typename PROG::I IP;
	const float rStep = IP1::reciprocalStepCount(stepCount);        // one divide per span, as in calcStep (tex1light1.cpp:93-107)
#pragma unroll
	for(int q = 0; q < NV; q++)
	{
		const float4 a = __ldg((const float4*)(v + q)), b = __ldg((const float4*)(v + NV + q)), c = __ldg((const float4*)(v + 2 * NV + q));
		const F4 v0 = f4(a.x, a.y, a.z, a.w), v1 = f4(b.x, b.y, b.z, b.w), v2 = f4(c.x, c.y, c.z, c.w);
		F4 vEnd;
		IP1::interpolateByContributes(&vStart[q], &v0, &v1, &v2, cl[0], cl[1], cl[2]);
		IP1::interpolateByContributes(&vEnd, &v0, &v1, &v2, cr[0], cr[1], cr[2]);
		IP1::calcStepR(&vStep[q], &vStart[q], &vEnd, rStep);
	}
	if(skip > 0) IP::stepForward(vStart, vStep, skip);               // interp.cpp:74-79
	return x1;
}

// MULTI: a batch of draws (BatchView above): this launch shades the survivors of its programme group's draws, with that draw's
// uniforms, textures and varyings. MARKS: fragments of long spans start their chains from the nearest chain mark
// (SpanStreams::markV, span_mark_vary_kernel).
// ORDERED: a draw that blends — the colour is left in the stream for shade_resolve_kernel (SurvivorStream2::colour).
template<class PROG, int MINB, bool MULTI = false, bool MARKS = false, bool ORDERED = false>
__global__ void __launch_bounds__(PS_SHADE_THREADS, MINB) shade_span_kernel(const __grid_constant__ DrawParams P, const SurvivorStream2 Q, const BatchView B)
{
	constexpr int NV = PROG::NV;
	if(*P.poison) return;
	const uint32_t n = min(*Q.count, Q.capacity);
	const uint32_t stride = gridDim.x * blockDim.x;
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	// pipeline registers: item i (xy, inv, record), item i + stride (stream words); item i + 2 * stride is read inside the loop.
	// (Tried and dropped: the record two items ahead and the next item's header + varyings asked into L1 by prefetch
	// instructions while this one is shaded: 0.186 -> 0.196 ms.)
	uint32_t xy0 = PS_SV_HOLE, sp1 = 0, xy1 = PS_SV_HOLE, sp0 = 0;
	float inv0 = 0, inv1 = 0;
	int4 rec0 = make_int4(1, 0, 0, 0);
	if(NV > 0)
	{
		if(i < n) { xy0 = Q.xy[i]; inv0 = Q.inv[i]; sp0 = Q.span[i]; rec0 = __ldg((const int4*)(P.sp.rec + sp0)); }
		if(i + stride < n && i + stride >= i) { xy1 = Q.xy[i + stride]; inv1 = Q.inv[i + stride]; sp1 = Q.span[i + stride]; }
	}
	for(; i < n; i += stride)
	{
		uint32_t xy;
		F4 frag[NV > 0 ? NV : 1];
		const DrawParams* D = &P;                                          // MULTI: the survivor's draw
		if(NV > 0)
		{
			xy = xy0;
			const float inv = inv0;
			const int4 A = rec0;
			const uint32_t spCur = sp0;                                    // (MARKS only: the survivor's record index)
			// the pipeline moves on
			{
				const uint32_t i1 = i + stride, i2 = i + 2 * stride;
				xy0 = xy1; inv0 = inv1; sp0 = sp1;
				if(i1 < n && i1 >= i) rec0 = __ldg((const int4*)(P.sp.rec + sp1));
				if(i2 < n && i2 >= i1 && i1 >= i) { xy1 = Q.xy[i2]; inv1 = Q.inv[i2]; sp1 = Q.span[i2]; }
			}
			if(PS_SV_HOLE == xy) continue;                                 // a reserved slot no survivor took
			const uint32_t tri = (uint32_t)A.w & 0xffffffu;
			if(MULTI)
			{
				const uint32_t bd = __ldg(B.blockDraw + tri / PS_GEOM_THREADS);
				if((bd >> 16) != B.group) continue;                          // another programme's survivor
				D = B.items + (bd & 0xffffu);
			}
			const uint4* src = (const uint4*)(P.hdr + tri);
			const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
			// (a batch's draw keeps its varyings densely, by the triangle's place in the draw)
			const uint32_t ltri = MULTI ? (tri % PS_GEOM_THREADS) + (tri / PS_GEOM_THREADS - D->batchFirstBlock) * D->batchTrisPerBlock : tri;
			const F4* v = D->vary + (size_t)ltri * 3 * NV;
			const int x = (int)(xy & 0x1fff), y = (int)((xy >> 13) & 0x1fff);
			const int left = A.x, right = A.y, e = (int)((uint32_t)A.w >> 24);
			// interpolateStartAndStep, interp.cpp:26-80 (the varyings' half; the depth half ran in the geometry kernel)
			F4 vStart[NV > 0 ? NV : 1], vStep[NV > 0 ? NV : 1];
			typedef typename PROG::I IP;
			const int x1 = varyingChainStart<PROG>(q0, q1, q2, v, left, right, e, y, vStart, vStep);
			int xr = x1;
			if(MARKS && P.sp.markCap && ((uint32_t)A.z & 0x7fffffffu) > 0x7f800000u && x - x1 >= PS_MARK_STEP)
			{
				// a long span (its depth bound is the NaN evalSpan left): from the nearest mark of its chain
				const uint32_t at = __ldg(P.sp.markAt + spCur);
				if(at != 0xffffffffu)
				{
					const int k = (x - x1) / PS_MARK_STEP;
					const float4* m = (const float4*)(P.sp.markV + ((size_t)at + k) * NV);
#pragma unroll
					for(int q = 0; q < NV; q++) { const float4 t = __ldg(m + q); vStart[q] = f4(t.x, t.y, t.z, t.w); }
					xr = x1 + k * PS_MARK_STEP;
				}
			}
#pragma unroll 1
			for(int q = xr; q < x; q++) IP::stepForward(vStart, vStep, 1);    // interp.cpp:88, one rounded add per pixel
			IP::correctInterpolation(frag, vStart, inv);
		}
		else
		{
			xy = Q.xy[i];
			if(PS_SV_HOLE == xy) continue;
			if(MULTI)
			{
				const uint32_t tri = P.sp.rec[Q.span[i]].triEdges & 0xffffffu;
				const uint32_t bd = __ldg(B.blockDraw + tri / PS_GEOM_THREADS);
				if((bd >> 16) != B.group) continue;
				D = B.items + (bd & 0xffffu);
			}
		}
		const int x = (int)(xy & 0x1fff), y = (int)((xy >> 13) & 0x1fff);
		FragmentProcessorOutput out;
		out.discarded = false; out.wrote = false; out.blendable = false; out.bgra = 0;
		PROG::F::process(frag, out, *D);                                     // fragthrd.cpp:231
		if(P.cap && x < P.capW && y < P.capH) atomicAdd(&P.cap[(size_t)y * P.capW + x], 1u);
		if(ORDERED)
		{
			Q.colour[i] = out.bgra;
			if(out.wrote) Q.xy[i] = xy | PS_SV_WROTE | (out.blendable ? PS_SV_BLEND : 0u);
		}
		else if(out.wrote && (xy & PS_SV_WINNER) && y < P.colour.height && x < P.colour.width)
		{
			// FBOBridge::write / write4 without ALPHABLEND: a plain store (fragthrd.cpp:54-82); later survivors of the pixel overwrite
			uint8_t* row = P.colour.ptr + (size_t)(P.colour.topDown ? P.colour.height - 1 - y : y) * P.colour.scanline;
			*(uint32_t*)(row + (size_t)x * 4) = out.bgra;
		}
	}
}

// ======================================================================================================================
// chain marks of long spans (SpanStreams): each long span's chain walked once, its state kept every PS_MARK_STEP pixels
// ======================================================================================================================

#define PS_MARK_THREADS 128

// lane = long span: the depth half's chain (cf2, z) — the same rounded additions, in the same order, as the replay in the
// raster kernel they stand in for
__global__ void __launch_bounds__(PS_MARK_THREADS) span_mark_depth_kernel(const __grid_constant__ DrawParams P)
{
	if(*P.poison) return;
	const uint32_t n = min(P.sp.longLatched[0], P.sp.capacity);
	const bool useDepth = 0 != (P.behavior & (PS_BEHAVIOR_TEST_DEPTH | PS_BEHAVIOR_UPDATE_DEPTH));
	const int limitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
	{
		const uint32_t idx = P.sp.longList[j];
		const uint32_t at = P.sp.markAt[idx];
		if(0xffffffffu == at) continue;
		const int4* rec = (const int4*)(P.sp.rec + idx);
		const int4 A = __ldg(rec), D = __ldg(rec + 1);
		const int x1 = A.x < 0 ? 0 : A.x;
		const int x2 = min(A.y >= P.vpW ? P.vpW - 1 : A.y, limitX);
		const int steps = x2 - x1;
		float cf2 = __int_as_float(D.x), z = __int_as_float(D.z);
		const float cf2Step = __int_as_float(D.y), zStep = __int_as_float(D.w);
		float2* out = (float2*)(P.sp.markZ + at);
		for(int k = 0; ; k++)
		{
			out[k] = make_float2(cf2, z);
			if((k + 1) * PS_MARK_STEP > steps) break;
#pragma unroll
			for(int i = 0; i < PS_MARK_STEP; i++)
			{
				cf2 = fadd(cf2, cf2Step);
				z = fadd(z, zStep);
			}
		}
	}
}

// lane = long span: the varyings' chain (shade_span_kernel's replay). MULTI: the spans of this launch's programme group.
template<class PROG, bool MULTI>
__global__ void __launch_bounds__(PS_MARK_THREADS) span_mark_vary_kernel(const __grid_constant__ DrawParams P, const BatchView B)
{
	constexpr int NV = PROG::NV > 0 ? PROG::NV : 1;
	if(0 == PROG::NV || *P.poison) return;
	const uint32_t n = min(P.sp.longLatched[0], P.sp.capacity);
	const bool useDepth = 0 != (P.behavior & (PS_BEHAVIOR_TEST_DEPTH | PS_BEHAVIOR_UPDATE_DEPTH));
	const int limitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	typedef typename PROG::I IP;
	for(uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
	{
		const uint32_t idx = P.sp.longList[j];
		const uint32_t at = P.sp.markAt[idx];
		if(0xffffffffu == at) continue;
		const int4 A = __ldg((const int4*)(P.sp.rec + idx));
		const uint32_t tri = (uint32_t)A.w & 0xffffffu;
		const DrawParams* D = &P;
		if(MULTI)
		{
			const uint32_t bd = __ldg(B.blockDraw + tri / PS_GEOM_THREADS);
			if((bd >> 16) != B.group) continue;
			D = B.items + (bd & 0xffffu);
		}
		const TriSpan ts = P.sp.tri[tri];
		const int y = (int)(ts.y & 0xffff) + (int)(idx - ts.x);         // records of a triangle are consecutive rows from its first
		const uint4* src = (const uint4*)(P.hdr + tri);
		const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
		const uint32_t ltri = MULTI ? (tri % PS_GEOM_THREADS) + (tri / PS_GEOM_THREADS - D->batchFirstBlock) * D->batchTrisPerBlock : tri;
		const F4* v = D->vary + (size_t)ltri * 3 * PROG::NV;
		F4 vStart[NV], vStep[NV];
		const int x1 = varyingChainStart<PROG>(q0, q1, q2, v, A.x, A.y, (int)((uint32_t)A.w >> 24), y, vStart, vStep);
		const int x2 = min(A.y >= P.vpW ? P.vpW - 1 : A.y, limitX);
		const int steps = x2 - x1;
		float4* out = (float4*)(P.sp.markV + (size_t)at * PROG::NV);
		for(int k = 0; ; k++)
		{
#pragma unroll
			for(int q = 0; q < PROG::NV; q++) out[(size_t)k * PROG::NV + q] = make_float4(vStart[q].x, vStart[q].y, vStart[q].z, vStart[q].w);
			if((k + 1) * PS_MARK_STEP > steps) break;
#pragma unroll 4
			for(int i = 0; i < PS_MARK_STEP; i++) IP::stepForward(vStart, vStep, 1);
		}
	}
}

// ======================================================================================================================
// draws that blend: a pixel's colours applied first to last (FBOBridge::write4 -> blend4 under ALPHABLEND, fragthrd.cpp:54-82)
// ======================================================================================================================

// warp = unit of the raster kernel (a tile, or a row group of one): its groups of survivors first to last; lane = survivor of the
// group; survivors of one pixel inside a group land in lane order.
__global__ void __launch_bounds__(128) shade_resolve_kernel(const __grid_constant__ DrawParams P, const SurvivorStream2 Q, int parts)
{
	if(*P.poison) return;
	const int lane = threadIdx.x & 31;
	const uint32_t u = blockIdx.x * 4 + (threadIdx.x >> 5);
	if(u >= P.tileOrder[P.tilesX * P.tilesY] * (uint32_t)parts) return;       // the units the raster kernel ran (same test)
	if(0 == P.tl.len[P.tileOrder[u / (uint32_t)parts]]) return;
	const bool alphaBlend = 0 != (P.behavior & PS_BEHAVIOR_ALPHABLEND);
	const uint32_t ltMask = (1u << lane) - 1;
	const uint32_t lastN = Q.chain[2 * u + 1];
	uint32_t at = Q.chain[2 * u];
	while(at)
	{
		const uint32_t base = at - 1;
		const uint32_t nextAt = Q.next[base];
		const uint32_t n = nextAt ? 32u : lastN;
		uint32_t xy = 0, c = 0;
		bool act = false;
		if((uint32_t)lane < n)
		{
			xy = Q.xy[base + lane];
			c = Q.colour[base + lane];
			const int x = (int)(xy & 0x1fff), y = (int)((xy >> 13) & 0x1fff);
			act = (xy & PS_SV_WROTE) && y < P.colour.height && x < P.colour.width;
		}
		const uint32_t peers = __match_any_sync(PS_FULL, act ? (xy & 0x3ffffffu) : 0x80000000u + (uint32_t)lane);
		const int rank = __popc(peers & ltMask);
		const int maxRank = __reduce_max_sync(PS_FULL, act ? rank : 0);
		uint32_t* dst = nullptr;
		if(act)
		{
			const int x = (int)(xy & 0x1fff), y = (int)((xy >> 13) & 0x1fff);
			uint8_t* row = P.colour.ptr + (size_t)(P.colour.topDown ? P.colour.height - 1 - y : y) * P.colour.scanline;
			dst = (uint32_t*)(row + (size_t)x * 4);
		}
#pragma unroll 1
		for(int r = 0; r <= maxRank; r++)
		{
			// FBOBridge::write4 -> blend4 under ALPHABLEND, FBOBridge::write -> plain store (fragthrd.cpp:54-82)
			if(act && rank == r) *(volatile uint32_t*)dst = ((xy & PS_SV_BLEND) && alphaBlend) ? blend4(c, *(volatile uint32_t*)dst) : c;
			__syncwarp();
		}
		at = nextAt;
	}
}
