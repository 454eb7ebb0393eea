// kernels.cuh — the hot path: geom_setup -> ordered binning -> tile raster + depth -> shade.
//
// Reference call stack replaced: PuresoftPipeline::drawVAO (drawvao.cpp:3-133) with
// processVertices / isBackFace (vertthrd.cpp), PuresoftRasterizer::pushTriangle (rasterizer.cpp),
// PuresoftInterpolater (interp.cpp), the per-scanline ring queues (rinque.h) and fragmentThread (fragthrd.cpp).
// One draw = these kernels enqueued back to back on the pipe's stream, no host synchronisation:
//
//   geom_setup<PROG, STAGED, MODE>
//                        thread = triangle. Position half: vertex functor x3 (positions staged by one TMA bulk copy per
//                        block), perspective divide, back-face, whole-triangle z reject, pushTriangle's set-up, sort-first
//                        band reject. Other half, for survivors: 64-byte TriHeader + varyings, the reference's per-row spans
//                        to find the tiles really touched, per-tile counts. Tall triangles are walked by a whole warp, whole-
//                        rectangle ones counted by a warp. Fused for whole frames; with a small band the two halves are two
//                        kernels around one global survivor list (PS_GEOM_APPEND / PS_GEOM_LIST).
//   tile_scan            one block: exclusive scan of the per-tile counts over the draw's tile range, the verdict on the
//                        capacities the host speculated with (poison), the tiles ordered by list length.
//   bin_fill, tile_list_sort
//                        lists filled in arrival order (atomics; whole-rectangle triangles by the whole block), then every
//                        tile's list sorted by triangle id = SUBMISSION ORDER, which the depth dead band and blend4 require
//                        (SURVEY.md §9.7). Lists too long for shared memory: emit_pairs + stable LSD radix sort.
//   tile_raster_depth    (split path, the default) one warp per 16x16 tile or per group of 8 / 4 of its rows, longest lists
//                        first, depth tile in shared memory for the whole draw; per chunk of 32 triangles: A1 lane = triangle,
//                        X lane = (triangle, row), Y lane = span, B lane = pixel — exact serial chains (§9.6), the depth rule
//                        in submission order; every survivor appended to a stream, the last one of each pixel remembered.
//   shade<PROG>          flat loop over the survivor stream: varyings, fragment functor once per survivor (fragthrd.cpp:231),
//                        the pixel's last survivor stores its colour.
//   tile_raster_shade_ordered<PROG>
//                        one-kernel tile path with colours committed in submission order: draws that blend (blend4 is not
//                        commutative).
//   tile_raster_shade_immediate<PROG>
//                        the first version (fragment functor inside the pixel loop): functors that may discard() while depth
//                        writes are on — there the depth write depends on the functor's result (fragthrd.cpp:234-237).
//   clear_depth / clear_colour / post_process<POST>
//                        fbo.cpp:332-371 (clear4 skips the last buffer row, replicated); post.cpp:3-19.
#pragma once
#include "shaders.cuh"
#include "raster.cuh"

#define PS_WARPS_PER_BLOCK 4
#define PS_FULL 0xffffffffu

PS_D unsigned long long warpSumU64(unsigned v)
{
	v += __shfl_xor_sync(PS_FULL, v, 16);
	v += __shfl_xor_sync(PS_FULL, v, 8);
	v += __shfl_xor_sync(PS_FULL, v, 4);
	v += __shfl_xor_sync(PS_FULL, v, 2);
	v += __shfl_xor_sync(PS_FULL, v, 1);
	return v;
}

// ======================================================================================================================
// geometry
// ======================================================================================================================

// vertthrd.cpp:56-65 isBackFace — sign of the NDC cross product's z after mcemaths_norm_3_4; NaN (zero area) => keep
PS_D bool isBackFace(F4 v0, F4 v1, F4 v2, const ApproxTables& ap)
{
	const F4 a = f4sub(v1, v0), b = f4sub(v2, v1);
	F4 c; // mcemaths_cross_3, vector.cpp:114-144
	c.x = fsub(fmul(a.y, b.z), fmul(a.z, b.y));
	c.y = fsub(fmul(a.z, b.x), fmul(a.x, b.z));
	c.z = fsub(fmul(a.x, b.y), fmul(a.y, b.x));
	c.w = fsub(fmul(a.w, b.w), fmul(a.w, b.w));
	c = f4norm(c, ap);
	return f4dot(f4(0, 0, 1.0f, 0), c) < 0;
}

// ---- bulk asynchronous copy global -> shared (TMA, 1-D) completing on an mbarrier -----------------------------------
PS_D uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PS_D void mbarInit(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PS_D void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
PS_D void bulkCopyG2S(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
PS_D void mbarWait(uint64_t* bar, uint32_t parity)
{
	uint32_t done = 0;
	while(!done)
		asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
		             : "=r"(done) : "r"(smemAddr(bar)), "r"(parity) : "memory");
}

// One tall triangle (held by lane `src`) walked by the whole warp, lane = every 32nd row: its spans and clamped fragments are
// added to the calling lanes' counters, its extent comes back reduced. Out of line: its registers stay out of the main path.
struct TallExtent { int minX, maxX, minY, maxY; unsigned spans, frags; };
__device__ __noinline__ TallExtent tallWalk(const DrawParams& P, uint32_t tri, int lane)
{
	unsigned spans = 0, frags = 0;
	// the triangle's record was stored by its owner lane before the __syncwarp() in front of this call: read it back (L2)
	const uint4* rec = (const uint4*)(P.hdr + tri);
	const uint4 q0 = __ldcg(rec), q1 = __ldcg(rec + 1), q3 = __ldcg(rec + 3);
	float vx[3], vy[3];
	vx[0] = __uint_as_float(q0.x); vy[0] = __uint_as_float(q0.y);
	vx[1] = __uint_as_float(q0.z); vy[1] = __uint_as_float(q0.w);
	vx[2] = __uint_as_float(q1.x); vy[2] = __uint_as_float(q1.y);
	const uint32_t rows = q3.x, half0 = q3.y, half1 = q3.z, plan = q3.w;
	const int firstRow = (int)(rows & 0xffff), lastRow = (int)(rows >> 16);
	const int l0 = (int)(half1 & 0xffff), l1 = (int)(half1 >> 16);
	const int u0 = (int)(half0 & 0xffff), u1 = (int)(half0 >> 16);
	const int selL = (int)((plan >> 8) & 0xff), selU = (int)(plan & 0xff);
	const Edge LU = makeEdge(vx, vy, selU & 3, (selU >> 2) & 3), RU = makeEdge(vx, vy, (selU >> 4) & 3, (selU >> 6) & 3);
	Edge LL = LU, RL = RU;
	if(l0 <= l1) { LL = makeEdge(vx, vy, selL & 3, (selL >> 2) & 3); RL = makeEdge(vx, vy, (selL >> 4) & 3, (selL >> 6) & 3); }
	int minX = 0x7fffffff, maxX = -1, minY = 0x7fffffff, maxY = -1;
#pragma unroll 1
	for(int iy = firstRow + lane; iy <= lastRow; iy += 32)
	{
		const bool lower = iy >= l0 && iy <= l1;
		if(!lower && !(iy >= u0 && iy <= u1)) continue;
		const float y = (float)iy;
		const Edge& L = lower ? LL : LU;
		const Edge& R = lower ? RL : RU;
		const int left = cvtt(fadd(edgeAt(L, y), 0.5f)), right = cvtt(fadd(edgeAt(R, y), 0.5f));   // rasterizer.cpp:98-117
		if(left == right) continue;                            // drawvao.cpp:72
		spans++;
		if(iy < P.band0 || iy >= P.band1) continue;
		const int x1 = left < 0 ? 0 : left;                     // RESULT_ROW::leftClamped
		const int x2 = right >= P.vpW ? P.vpW - 1 : right;      // RESULT_ROW::rightClamped
		if(x1 > x2) continue;
		frags += (unsigned)(x2 - x1 + 1);
		minX = min(minX, x1); maxX = max(maxX, x2);
		minY = min(minY, iy); maxY = max(maxY, iy);
	}
	minX = __reduce_min_sync(PS_FULL, minX); maxX = __reduce_max_sync(PS_FULL, maxX);
	minY = __reduce_min_sync(PS_FULL, minY); maxY = __reduce_max_sync(PS_FULL, maxY);
	TallExtent e;
	e.minX = minX; e.maxX = maxX; e.minY = minY; e.maxY = maxY; e.spans = spans; e.frags = frags;
	return e;
}

#define PS_GEOM_THREADS 128
#define PS_TALL_ROWS 64        // triangles this many rows high are walked by a whole warp

// STAGED: the block's vertex range (PS_GEOM_THREADS x 3 consecutive elements, contiguous in the un-indexed stream) of the
// POSITION slot — slot 0 in every vertex functor of the reference — is brought into shared memory by one bulk copy; the
// position pass, which every triangle runs, then reads shared memory with no load latency on its arithmetic path. The
// other slots are read straight from global memory by the varyings pass, i.e. only for triangles that survive culling
// and (sort-first) lie in this rank's band: at N ranks most triangles never touch them.
#define PS_GEOM_STAGE_SLOTS 1u
// MODE: PS_GEOM_FUSED  one kernel does everything (whole frames).
//       PS_GEOM_APPEND the position half only: survivors of the sort-first band get their header stored and their index
//                      appended to one global list (warp-aggregated).
//       PS_GEOM_LIST   the varyings / row-walk half as a dense kernel over that list. With a band a rank keeps ~1/N of a
//                      shuffled stream; compaction inside a block does not shorten this half in proportion (the block pays
//                      its latency whatever the lane count), a grid over the list does.
#define PS_GEOM_FUSED 0
#define PS_GEOM_APPEND 1
#define PS_GEOM_LIST 2
template<class PROG, bool STAGED, int MODE>
__global__ void __launch_bounds__(PS_GEOM_THREADS) geom_setup_kernel(const __grid_constant__ DrawParams P)
{
	constexpr int NV = PROG::NV;
	static_assert(!(STAGED && PS_GEOM_LIST == MODE), "the list-driven half reads global memory");
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
	if(PS_GEOM_LIST == MODE && blockIdx.x * blockDim.x >= *P.workCount) return;   // (the grid is sized for every triangle)
	extern __shared__ __align__(128) uint8_t stage[];
	__shared__ uint64_t stageBar;
	uint32_t stageOff[16];
	if(STAGED)
	{
		const uint32_t tri0 = blockIdx.x * PS_GEOM_THREADS;
		const uint32_t nt = min((uint32_t)PS_GEOM_THREADS, P.ntris - tri0);
		uint32_t off = 0;
#pragma unroll
		for(int s = 0; s < 16; s++)
		{
			stageOff[s] = off;
			if(((PROG::V::SLOTS & PS_GEOM_STAGE_SLOTS) >> s) & 1) off += (PS_GEOM_THREADS * 3 * P.stride[s] + 127u) & ~127u;
		}
		if(0 == threadIdx.x) mbarInit(&stageBar, 1);
		__syncthreads();
		if(0 == threadIdx.x)
		{
			uint32_t total = 0;
#pragma unroll
			for(int s = 0; s < 16; s++)
				if(((PROG::V::SLOTS & PS_GEOM_STAGE_SLOTS) >> s) & 1) total += (nt * 3 * P.stride[s] + 15u) & ~15u;
			mbarExpectTx(&stageBar, total);
#pragma unroll
			for(int s = 0; s < 16; s++)
				if(((PROG::V::SLOTS & PS_GEOM_STAGE_SLOTS) >> s) & 1)
					bulkCopyG2S(stage + stageOff[s], P.slot[s] + (size_t)tri0 * 3 * P.stride[s], (nt * 3 * P.stride[s] + 15u) & ~15u, &stageBar);
		}
		mbarWait(&stageBar, 0);
	}
	unsigned rasterised = 0, spans = 0, frags = 0;
	// ---- part A, thread = triangle: positions, perspective divide, back-face, z reject, pushTriangle's set-up, band reject.
	// Survivors are COMPACTED through shared memory (order kept), so that part B — varyings, records, row walk — runs on
	// dense warps: with culling, and above all with sort-first bands (a rank keeps ~1/N of a shuffled stream), the
	// survivors are scattered over the block's lanes and every warp would otherwise execute part B for a few of them.
	__shared__ TriHeader workHdr[PS_GEOM_THREADS];
	__shared__ uint32_t workTri[PS_GEOM_THREADS];
	__shared__ uint32_t warpAlive[PS_GEOM_THREADS / 32];
	bool alive = false;
	TriHeader h;
	if(PS_GEOM_LIST != MODE && tri < P.ntris)
	{
		float ndcX[3], ndcY[3], rw[3], pz[3];
		F4 pos[3];
		// The vertex functor runs twice per vertex: here only its position is used (the compiler drops the varyings and the
		// loads that feed nothing else), and again in part B — for surviving triangles only — for the varyings, one vertex
		// at a time. Culled triangles never touch their normal / tangent / uv streams.
#pragma unroll
		for(int i = 0; i < 3; i++)
		{
			// processVertices, vertthrd.cpp:14-51 : all attached slots advance in lock-step, un-indexed
			VertexProcessorInput in;
#pragma unroll
			for(int s = 0; s < 16; s++)
				in.data[s] = (PROG::V::SLOTS >> s) & 1 ? ((STAGED && ((PS_GEOM_STAGE_SLOTS >> s) & 1)) ? stage + stageOff[s] + (size_t)(threadIdx.x * 3 + i) * P.stride[s]
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
			// sort-first: a triangle whose rows all lie outside this rank's band leaves nothing behind here (no header, no
			// varyings, no row walk; its spans are counted by the rank that owns them)
			if(alive && ((int)(h.rows >> 16) < P.band0 || (int)(h.rows & 0xffff) >= P.band1)) alive = false;
			h.rw0 = rw[0]; h.rw1 = rw[1]; h.rw2 = rw[2];
			h.z0 = pz[0]; h.z1 = pz[1]; h.z2 = pz[2];
		}
		if(!alive)
		{
			P.triCount[tri] = 0;
			P.triRect[3 * tri] = 0; P.triRect[3 * tri + 1] = 0; P.triRect[3 * tri + 2] = 0;
		}
	}
	// Compaction pays when survivors are sparse in the block — a sort-first band on a shuffled stream; on a whole frame of
	// front-facing triangles it only adds two barriers (C2, one GPU: 0.127 -> 0.141 ms), so it is taken with a band only.
	const bool compact = PS_GEOM_FUSED == MODE && (P.band0 > 0 || P.band1 < P.vpH);
	bool work = alive;
	uint32_t wtri = tri;
	if(PS_GEOM_APPEND == MODE)
	{
		// survivors: header to its final place, index to the list (one atomic per warp; the list's order does not matter)
		const uint32_t aliveBallot = __ballot_sync(PS_FULL, alive);
		uint32_t base = 0;
		if(0 == (threadIdx.x & 31) && aliveBallot) base = atomicAdd(P.workCount, (uint32_t)__popc(aliveBallot));
		base = __shfl_sync(PS_FULL, base, 0);
		if(alive)
		{
			P.workList[base + (uint32_t)__popc(aliveBallot & ((1u << (threadIdx.x & 31)) - 1))] = tri;
			uint4* dst = (uint4*)(P.hdr + tri);
			const uint4* src = (const uint4*)&h;
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
		}
		work = false;
	}
	if(PS_GEOM_LIST == MODE)
	{
		work = tri < *P.workCount;
		if(work)
		{
			wtri = P.workList[tri];
			const uint4* src = (const uint4*)(P.hdr + wtri);
			uint4* dst = (uint4*)&h;
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
		}
	}
	if(compact)
	{
		const uint32_t aliveBallot = __ballot_sync(PS_FULL, alive);
		if(0 == (threadIdx.x & 31)) warpAlive[threadIdx.x >> 5] = (uint32_t)__popc(aliveBallot);
		__syncthreads();
		uint32_t slotBase = 0, nWork = 0;
#pragma unroll
		for(int w = 0; w < PS_GEOM_THREADS / 32; w++)
		{
			if(w < (int)(threadIdx.x >> 5)) slotBase += warpAlive[w];
			nWork += warpAlive[w];
		}
		if(alive)
		{
			const uint32_t slot = slotBase + (uint32_t)__popc(aliveBallot & ((1u << (threadIdx.x & 31)) - 1));
			uint4* dst = (uint4*)&workHdr[slot];
			const uint4* src = (const uint4*)&h;
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
			workTri[slot] = tri;
		}
		__syncthreads();
		work = threadIdx.x < nWork;
		if(work)
		{
			wtri = workTri[threadIdx.x];
			const uint4* src = (const uint4*)&workHdr[threadIdx.x];
			uint4* dst = (uint4*)&h;
			dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
		}
	}

	// ---- part B, thread = surviving triangle (dense when compacted): records, varyings, the row walk that finds the tiles really touched
	uint32_t wCount = 0, wRect0 = 0, wRect1 = 0, wMask = 0;
	bool tall = false, bigPending = false;
	if(work)
	{
		const float vx[3] = { h.vx0, h.vx1, h.vx2 }, vy[3] = { h.vy0, h.vy1, h.vy2 };
		uint32_t count = 0, rect0 = 0, rect1 = 0, mask = 0;
		{
			{
				// The records go out BEFORE the row walk: the 3 x NV varyings would otherwise stay live in registers across it
				// and halve the occupancy. (A triangle that turns out to cover no pixel wrote its record for nothing.)
				{
					// 64-byte record as four 16-byte stores
					uint4* dst = (uint4*)(P.hdr + wtri);
					const uint4* src = (const uint4*)&h;
					dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
					if(NV > 0)
					{
						float4* vd = (float4*)(P.vary + (size_t)wtri * 3 * NV);
						const uint32_t stageIdx = wtri - blockIdx.x * PS_GEOM_THREADS;   // the triangle's place in the block's staged range
#pragma unroll 1
						for(int i = 0; i < 3; i++)
						{
							VertexProcessorInput in;
#pragma unroll
							for(int s = 0; s < 16; s++)
								in.data[s] = (PROG::V::SLOTS >> s) & 1 ? ((STAGED && ((PS_GEOM_STAGE_SLOTS >> s) & 1)) ? stage + stageOff[s] + (size_t)(stageIdx * 3 + i) * P.stride[s]
								                                                  : P.slot[s] + (size_t)(wtri * 3 + i) * P.stride[s]) : nullptr;
							VertexProcessorOutput<NV> vo;
							PROG::V::process(in, vo, P);
#pragma unroll
							for(int k = 0; k < NV; k++)
								vd[i * NV + k] = make_float4(vo.user[k].x, vo.user[k].y, vo.user[k].z, vo.user[k].w);
						}
					}
				}
				const int firstRow = (int)(h.rows & 0xffff), lastRow = (int)(h.rows >> 16);
				// a triangle PS_TALL_ROWS rows high or more is walked by the whole warp further down (a 4096-row shadow-map
				// triangle would keep this one thread busy for a millisecond)
				tall = lastRow - firstRow >= PS_TALL_ROWS;
				if(!tall)
				{
				int minX = 0x7fffffff, maxX = -1, minY = 0x7fffffff, maxY = -1;
				// for the first four tile rows the triangle touches: the tile columns its spans reach (lo | hi << 16)
				uint32_t tr0 = 0, tr1 = 0, tr2 = 0, tr3 = 0;
				int tyFirst = -1, kCur = -1, curLo = 0, curHi = 0;
				uint32_t trValid = 0;                                  // tile rows (relative to the first) that hold a span
				// The rows of RESULT (drawvao.cpp:66-75) in one flat loop (a lane's trip count is its row count, whichever half the
				// rows belong to). The edges of both halves are set up once; the lower half owns the shared row
				// (rasterizer.cpp:128-139), rows outside both halves were never written by the reference and are skipped.
				const int l0 = (int)(h.half1 & 0xffff), l1 = (int)(h.half1 >> 16);
				const int u0 = (int)(h.half0 & 0xffff), u1 = (int)(h.half0 >> 16);
				const int selL = (int)((h.plan >> 8) & 0xff), selU = (int)(h.plan & 0xff);
				const Edge LU = makeEdge(vx, vy, selU & 3, (selU >> 2) & 3), RU = makeEdge(vx, vy, (selU >> 4) & 3, (selU >> 6) & 3);
				Edge LL = LU, RL = RU;
				if(l0 <= l1) { LL = makeEdge(vx, vy, selL & 3, (selL >> 2) & 3); RL = makeEdge(vx, vy, (selL >> 4) & 3, (selL >> 6) & 3); }
#pragma unroll 1
				for(int iy = firstRow; iy <= lastRow; iy++)
				{
					const bool lower = iy >= l0 && iy <= l1;
					if(!lower && !(iy >= u0 && iy <= u1)) continue;
					const float y = (float)iy;
					Edge L, R;
					L.dx = lower ? LL.dx : LU.dx; L.dy = lower ? LL.dy : LU.dy; L.x0 = lower ? LL.x0 : LU.x0; L.y0 = lower ? LL.y0 : LU.y0;
					R.dx = lower ? RL.dx : RU.dx; R.dy = lower ? RL.dy : RU.dy; R.x0 = lower ? RL.x0 : RU.x0; R.y0 = lower ? RL.y0 : RU.y0;
					const int left = cvtt(fadd(edgeAt(L, y), 0.5f)), right = cvtt(fadd(edgeAt(R, y), 0.5f));   // rasterizer.cpp:98-117
					if(left == right) continue;                            // drawvao.cpp:72
					spans++;
					if(iy < P.band0 || iy >= P.band1) continue;
					const int x1 = left < 0 ? 0 : left;                     // RESULT_ROW::leftClamped
					const int x2 = right >= P.vpW ? P.vpW - 1 : right;      // RESULT_ROW::rightClamped
					if(x1 > x2) continue;
					frags += (unsigned)(x2 - x1 + 1);
					minX = min(minX, x1); maxX = max(maxX, x2);
					minY = min(minY, iy); maxY = max(maxY, iy);
					const int ty = iy / PS_TILE;
					if(tyFirst < 0) tyFirst = ty;
					const int k = ty - tyFirst;
					trValid |= 1u << min(k, 31);
					if(k != kCur)
					{
						const uint32_t packed = (uint32_t)curLo | ((uint32_t)curHi << 16);
						if(0 == kCur) tr0 = packed; else if(1 == kCur) tr1 = packed; else if(2 == kCur) tr2 = packed; else if(3 == kCur) tr3 = packed;
						kCur = k; curLo = x1 / PS_TILE; curHi = x2 / PS_TILE;
					}
					else { curLo = min(curLo, x1 / PS_TILE); curHi = max(curHi, x2 / PS_TILE); }
				}
				{
					const uint32_t packed = (uint32_t)curLo | ((uint32_t)curHi << 16);
					if(0 == kCur) tr0 = packed; else if(1 == kCur) tr1 = packed; else if(2 == kCur) tr2 = packed; else if(3 == kCur) tr3 = packed;
				}
				if(maxX >= 0)
				{
					const int tx0 = minX / PS_TILE, tx1 = maxX / PS_TILE, ty0 = minY / PS_TILE, ty1 = maxY / PS_TILE;
					rect0 = (uint32_t)tx0 | ((uint32_t)tx1 << 16);
					rect1 = (uint32_t)ty0 | ((uint32_t)ty1 << 16);
					if(tx1 - tx0 >= 8 || ty1 - ty0 >= 4)
					{
						// large triangle: the whole rectangle of tiles (the tile kernels drop rows that miss a tile)
						mask = 0xffffffffu;
						count = (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1);
						bigPending = true;                              // its tiles are counted by the whole warp further down
					}
					else
					{
						// small triangle (the common case): exactly the tiles some span of it reaches
#pragma unroll
						for(int k = 0; k < 4; k++)
						{
							const uint32_t cur = 0 == k ? tr0 : (1 == k ? tr1 : (2 == k ? tr2 : tr3));
							const int a = (int)(cur & 0xffff) - tx0, b = (int)(cur >> 16) - tx0;
							if((trValid >> k) & 1) mask |= ((2u << b) - (1u << a)) << (k * 8);   // 0 <= a <= b <= 7
						}
						count = (uint32_t)__popc(mask);
						for(uint32_t m = mask; m; m &= m - 1)
						{
							const int bit = __ffs(m) - 1;
							atomicAdd(&P.tileCount[(ty0 + (bit >> 3)) * P.tilesX + tx0 + (bit & 7)], 1u);
						}
					}
				}
				}
			}
		}
		wCount = count; wRect0 = rect0; wRect1 = rect1; wMask = mask;
	}
	// ---- tall triangles: the warp walks one at a time, lane = every 32nd row. Only the extent is needed: they are binned
	// by their whole rectangle (a rectangle of fewer than 8 x 4 tiles included: a superset of the tiles touched is harmless,
	// the tile kernels drop the rows that miss a tile).
	__syncwarp();
	{
		const int lane = threadIdx.x & 31;
		uint32_t todo = __ballot_sync(PS_FULL, work && tall);
		while(todo)
		{
			const int src = __ffs(todo) - 1;
			todo &= todo - 1;
			const TallExtent te = tallWalk(P, __shfl_sync(PS_FULL, wtri, src), lane);
			const int minX = te.minX, maxX = te.maxX, minY = te.minY, maxY = te.maxY;
			spans += te.spans; frags += te.frags;
			if(lane == src && maxX >= 0)
			{
				const int tx0 = minX / PS_TILE, tx1 = maxX / PS_TILE, ty0 = minY / PS_TILE, ty1 = maxY / PS_TILE;
				wRect0 = (uint32_t)tx0 | ((uint32_t)tx1 << 16);
				wRect1 = (uint32_t)ty0 | ((uint32_t)ty1 << 16);
				wMask = 0xffffffffu;
				wCount = (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1);
				bigPending = true;
			}
		}
		if(work)
		{
			P.triCount[wtri] = wCount;
			P.triRect[3 * wtri] = wRect0;
			P.triRect[3 * wtri + 1] = wRect1;
			P.triRect[3 * wtri + 2] = wMask;
		}
		// ---- whole-rectangle triangles: the warp counts one's tiles at a time, lane = every 32nd tile
		todo = __ballot_sync(PS_FULL, bigPending);
		while(todo)
		{
			const int src = __ffs(todo) - 1;
			todo &= todo - 1;
			const uint32_t r0 = __shfl_sync(PS_FULL, wRect0, src), r1 = __shfl_sync(PS_FULL, wRect1, src);
			const int tx0 = (int)(r0 & 0xffff), tx1 = (int)(r0 >> 16), ty0 = (int)(r1 & 0xffff), ty1 = (int)(r1 >> 16);
			const uint32_t w = (uint32_t)(tx1 - tx0 + 1), total = w * (uint32_t)(ty1 - ty0 + 1);
			for(uint32_t i = (uint32_t)lane; i < total; i += 32)
				atomicAdd(&P.tileCount[(ty0 + (int)(i / w)) * P.tilesX + tx0 + (int)(i % w)], 1u);
		}
	}
	// counters: one set of atomics per block, on the block's replica
	__shared__ unsigned long long blockSums[PS_GEOM_THREADS / 32][3];
	__shared__ unsigned blockRange[PS_GEOM_THREADS / 32][2];
	{
		// the draw's range of tile indices: corners of every binned triangle's rectangle (bounds of the tiles it touches)
		const unsigned loInv = wCount ? ~((wRect1 & 0xffff) * (unsigned)P.tilesX + (wRect0 & 0xffff)) : 0u;
		const unsigned hi1 = wCount ? (wRect1 >> 16) * (unsigned)P.tilesX + (wRect0 >> 16) + 1u : 0u;
		const unsigned a = __reduce_max_sync(PS_FULL, loInv), b = __reduce_max_sync(PS_FULL, hi1);
		if(0 == (threadIdx.x & 31)) { blockRange[threadIdx.x >> 5][0] = a; blockRange[threadIdx.x >> 5][1] = b; }
	}
	const unsigned long long r = warpSumU64(rasterised), s = warpSumU64(spans);
	unsigned long long f = frags;
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) f += __shfl_xor_sync(PS_FULL, f, d);
	if(0 == (threadIdx.x & 31)) { blockSums[threadIdx.x >> 5][0] = r; blockSums[threadIdx.x >> 5][1] = s; blockSums[threadIdx.x >> 5][2] = f; }
	__syncthreads();
	if(threadIdx.x < 3)
	{
		unsigned long long v = 0;
#pragma unroll
		for(int w = 0; w < PS_GEOM_THREADS / 32; w++) v += blockSums[w][threadIdx.x];
		DeviceStats* st = P.stats + (blockIdx.x & (PS_STATS_COPIES - 1));
		if(v) atomicAdd(0 == threadIdx.x ? &st->triangles_rasterised : (1 == threadIdx.x ? &st->spans : &st->fragBound), v);
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

// ======================================================================================================================
// exclusive scan (uint32), three small kernels
// ======================================================================================================================

#define PS_SCAN_THREADS 256
#define PS_SCAN_ITEMS 8
#define PS_SCAN_BLOCK (PS_SCAN_THREADS * PS_SCAN_ITEMS)

__global__ void __launch_bounds__(PS_SCAN_THREADS) scan_local_kernel(const uint32_t* in, uint32_t* out,
                                                                    uint32_t* __restrict__ blockSums, uint32_t n)
{
	__shared__ uint32_t warpTotals[PS_SCAN_THREADS / 32];
	const uint32_t base = blockIdx.x * PS_SCAN_BLOCK + threadIdx.x * PS_SCAN_ITEMS;
	uint32_t v[PS_SCAN_ITEMS], sum = 0;
#pragma unroll
	for(int i = 0; i < PS_SCAN_ITEMS; i++)
	{
		v[i] = (base + i < n) ? in[base + i] : 0;
		sum += v[i];
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = sum;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1)
	{
		uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
		if(lane >= d) incl += t;
	}
	if(31 == lane) warpTotals[warp] = incl;
	__syncthreads();
	uint32_t warpBase = 0;
	for(int w = 0; w < warp; w++) warpBase += warpTotals[w];
	uint32_t run = warpBase + incl - sum;
#pragma unroll
	for(int i = 0; i < PS_SCAN_ITEMS; i++)
	{
		if(base + i < n) out[base + i] = run;
		run += v[i];
	}
	if(PS_SCAN_THREADS - 1 == threadIdx.x) blockSums[blockIdx.x] = run;
}

// one block: exclusive scan of blockSums[nb] in place; total -> *total
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* __restrict__ blockSums, uint32_t nb, uint32_t* __restrict__ total)
{
	__shared__ uint32_t warpTotals[32];
	__shared__ uint32_t carryS;
	if(0 == threadIdx.x) carryS = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for(uint32_t base = 0; base < nb; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < nb ? blockSums[i] : 0;
		uint32_t incl = v;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		if(31 == lane) warpTotals[warp] = incl;
		__syncthreads();
		uint32_t warpBase = 0;
		for(int w = 0; w < warp; w++) warpBase += warpTotals[w];
		const uint32_t carry = carryS;
		if(i < nb) blockSums[i] = carry + warpBase + incl - v;
		__syncthreads();
		if(1023 == threadIdx.x) carryS = carry + warpBase + incl;
		__syncthreads();
	}
	if(0 == threadIdx.x) *total = carryS;
}

__global__ void __launch_bounds__(PS_SCAN_THREADS) scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ blockSums, uint32_t n)
{
	const uint32_t base = blockIdx.x * PS_SCAN_BLOCK + threadIdx.x * PS_SCAN_ITEMS;
	const uint32_t add = blockSums[blockIdx.x];
#pragma unroll
	for(int i = 0; i < PS_SCAN_ITEMS; i++)
		if(base + i < n) out[base + i] += add;
}

// ======================================================================================================================
// binning: emit (tile, triangle) pairs in triangle order, then a stable LSD radix sort on the tile id
// ======================================================================================================================

__global__ void __launch_bounds__(128) emit_pairs_kernel(const uint32_t* __restrict__ triCount, const uint32_t* __restrict__ triOffset,
                                                        const uint32_t* __restrict__ triRect, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ vals, uint32_t ntris, int tilesX)
{
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
	if(tri >= ntris) return;
	if(0 == triCount[tri]) return;
	uint32_t o = triOffset[tri];
	const uint32_t r0 = triRect[3 * tri], r1 = triRect[3 * tri + 1], mask = triRect[3 * tri + 2];
	const int tx0 = (int)(r0 & 0xffff), tx1 = (int)(r0 >> 16), ty0 = (int)(r1 & 0xffff), ty1 = (int)(r1 >> 16);
	const bool whole = 0xffffffffu == mask;
	for(int ty = ty0; ty <= ty1; ty++)
		for(int tx = tx0; tx <= tx1; tx++)
		{
			if(!whole && 0 == ((mask >> ((ty - ty0) * 8 + tx - tx0)) & 1)) continue;
			keys[o] = (uint32_t)(ty * tilesX + tx);
			vals[o] = tri;
			o++;
		}
}

// ---- the default binning: per-tile counts came from geom_setup; scan them, fill the lists with atomics (any order), then
// sort every tile's list by triangle id = submission order (§9.7). Lists longer than PS_SORT_LIMIT use the radix path.

#define PS_SORT_LIMIT 2048

// one block: tileStart = exclusive scan of tileCount (ntiles + 1 entries); tileCount and tileFill are left zeroed for the
// next draw; the draw's fragment bound is folded out of the counter replicas. Then the verdict on the capacities the
// host speculated with: if the pairs, the survivor bound or the longest list do not fit, *poison is raised and every
// kernel enqueued behind this one returns at once, leaving the targets untouched for the host's exact retry.
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tileStart,
                                                        uint32_t* __restrict__ tileFill, uint32_t ntiles, DeviceStats* stats,
                                                        uint32_t pairCap, unsigned long long survivorCap, uint32_t listLimit,
                                                        uint32_t* poison, DrawReport* report, uint32_t* __restrict__ tileOrder)
{
	__shared__ uint32_t warpTotals[32];
	__shared__ uint32_t carryS, longestS, passTotalS, nonEmptyS;
	__shared__ unsigned long long boundS;
	__shared__ uint32_t rangeLoS, rangeHiS;            // the draw binned to tiles [lo, hi) only (geom_setup's replicas, folded below)
	__shared__ uint32_t hist[256];                     // tiles per length class, longest lists first
	if(0 == threadIdx.x) { carryS = 0; longestS = 0; }
	if(threadIdx.x < 256) hist[threadIdx.x] = 0;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if(0 == warp)
	{
		unsigned long long b = lane < PS_STATS_COPIES ? stats[lane].fragBound : 0ull;
		if(lane < PS_STATS_COPIES) stats[lane].fragBound = 0;
#pragma unroll
		for(int d = 16; d > 0; d >>= 1) b += __shfl_xor_sync(PS_FULL, b, d);
		if(0 == lane) boundS = b;
		unsigned loInv = lane < PS_STATS_COPIES ? stats[lane].tileLoInv : 0u, hi1 = lane < PS_STATS_COPIES ? stats[lane].tileHi1 : 0u;
		if(lane < PS_STATS_COPIES) { stats[lane].tileLoInv = 0; stats[lane].tileHi1 = 0; }
		loInv = __reduce_max_sync(PS_FULL, loInv); hi1 = __reduce_max_sync(PS_FULL, hi1);
		if(0 == lane) { rangeLoS = hi1 ? min(~loInv, ntiles) : ntiles; rangeHiS = min(hi1, ntiles); }
	}
	__syncthreads();
	// everything outside [lo, hi) is empty: its tileStart is never read (only tiles with a list are ordered and visited)
	const uint32_t lo = rangeLoS, hi = max(rangeHiS, rangeLoS);
	uint32_t longest = 0;
	if(ntiles <= 16384)
	{
	// one tile per thread, 1024 per pass (C2's 8160 tiles: measured faster than the 8-per-thread form below)
	for(uint32_t base = lo; base < hi + 1; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < hi ? tileCount[i] : 0;
		longest = max(longest, v);
		if(i < hi && v) atomicAdd(&hist[255u - min(v >> 2, 255u)], 1u);
		uint32_t incl = v;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		if(31 == lane) warpTotals[warp] = incl;
		__syncthreads();
		uint32_t warpBase = 0;
		for(int w = 0; w < warp; w++) warpBase += warpTotals[w];
		const uint32_t carry = carryS;
		if(i < hi + 1) { tileStart[i] = carry + warpBase + incl - v; if(i < hi && v) { tileFill[i] = 0; tileCount[i] = 0; } }
		__syncthreads();
		if(1023 == threadIdx.x) carryS = carry + warpBase + incl;
		__syncthreads();
	}
	}
	else
	{
	// 8 consecutive tiles per thread, 8192 per pass: thread-local prefix, one block-wide scan of the thread sums, carry
	// between passes. (No faster than 1024-tile passes on C2's 8160 tiles, but a 4096^2 shadow map has 65536: 8 passes, not 64.)
	for(uint32_t base = lo; base < hi + 1; base += 8192)
	{
		const uint32_t i0 = base + threadIdx.x * 8;
		uint32_t v[8], sum = 0;
#pragma unroll
		for(int k = 0; k < 8; k++)
		{
			v[k] = i0 + k < hi ? tileCount[i0 + k] : 0;
			longest = max(longest, v[k]);
			if(i0 + k < hi && v[k]) atomicAdd(&hist[255u - min(v[k] >> 2, 255u)], 1u);
			sum += v[k];
		}
		uint32_t incl = sum;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		if(31 == lane) warpTotals[warp] = incl;
		__syncthreads();
		if(0 == warp)
		{
			const uint32_t w = warpTotals[lane];
			uint32_t wi = w;
#pragma unroll
			for(int d = 1; d < 32; d <<= 1)
			{
				const uint32_t t = __shfl_up_sync(PS_FULL, wi, d);
				if(lane >= d) wi += t;
			}
			warpTotals[lane] = wi - w;                     // exclusive
			if(31 == lane) passTotalS = wi;
		}
		__syncthreads();
		const uint32_t carry = carryS;
		uint32_t run = carry + warpTotals[warp] + incl - sum;
#pragma unroll
		for(int k = 0; k < 8; k++)
		{
			if(i0 + k < hi + 1) tileStart[i0 + k] = run;
			if(i0 + k < hi && v[k]) { tileFill[i0 + k] = 0; tileCount[i0 + k] = 0; }
			run += v[k];
		}
		__syncthreads();
		if(0 == threadIdx.x) carryS = carry + passTotalS;
		__syncthreads();
	}
	}
	longest = __reduce_max_sync(PS_FULL, longest);
	if(0 == lane && longest) atomicMax(&longestS, longest);
	__syncthreads();
	// tileOrder: the tiles by descending list length (counting sort over 256 length classes; any order inside a class).
	// The tile kernels take their tiles in this order, so the longest lists start first and the grid's tail is made of
	// short ones; the four warps of a block get lists of similar length.
	if(0 == warp)
	{
		uint32_t h[8], sum = 0;
#pragma unroll
		for(int k = 0; k < 8; k++) { h[k] = hist[lane * 8 + k]; sum += h[k]; }
		uint32_t incl = sum;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		uint32_t run = incl - sum;
#pragma unroll
		for(int k = 0; k < 8; k++) { hist[lane * 8 + k] = run; run += h[k]; }
		if(31 == lane) nonEmptyS = run;
	}
	__syncthreads();
	// only tiles with a list are ordered; tileOrder[ntiles] = how many there are (the tile kernels stop there)
	for(uint32_t i = lo + threadIdx.x; i < hi; i += 1024)
	{
		const uint32_t v = tileStart[i + 1] - tileStart[i];
		if(v) tileOrder[atomicAdd(&hist[255u - min(v >> 2, 255u)], 1u)] = i;
	}
	if(0 == threadIdx.x) tileOrder[ntiles] = nonEmptyS;
	if(0 == threadIdx.x)
	{
		const uint32_t total = carryS, lng = longestS;
		const unsigned long long bound = boundS;
		const uint32_t bad = (total > pairCap || bound > survivorCap || lng > listLimit) ? 1u : 0u;
		report->pairs = total; report->longest = lng; report->fragBound = bound; report->bad = bad;
		if(bad) report->sticky = 1;
		__threadfence_system();
		*poison = bad;
	}
}

__global__ void __launch_bounds__(128) bin_fill_kernel(const uint32_t* __restrict__ triCount, const uint32_t* __restrict__ triRect,
                                                      const uint32_t* __restrict__ tileStart, uint32_t* __restrict__ tileFill,
                                                      uint32_t* __restrict__ lists, uint32_t ntris, int tilesX, const uint32_t* __restrict__ poison)
{
	// triangles binned by their whole rectangle (more than 8 x 4 tiles — up to every tile of the target): the BLOCK takes
	// them one at a time, every thread four tiles per step, so that hundreds of the (returning) atomics are in flight
	// instead of one thread's one (a 4096^2 shadow map's ground quad is 2 x 23 000 tiles: 12 ms per triangle serially)
	__shared__ uint32_t bigTri[128], bigR0[128], bigR1[128];
	__shared__ uint32_t nBig;
	if(*poison) return;
	if(0 == threadIdx.x) nBig = 0;
	__syncthreads();
	const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
	if(tri < ntris && 0 != triCount[tri])
	{
		const uint32_t r0 = triRect[3 * tri], r1 = triRect[3 * tri + 1], mask = triRect[3 * tri + 2];
		if(0xffffffffu == mask)
		{
			const uint32_t k = atomicAdd(&nBig, 1u);
			bigTri[k] = tri; bigR0[k] = r0; bigR1[k] = r1;
		}
		else
		{
			const int tx0 = (int)(r0 & 0xffff), ty0 = (int)(r1 & 0xffff);
			for(uint32_t m = mask; m; m &= m - 1)
			{
				const int bit = __ffs(m) - 1;
				const uint32_t tile = (uint32_t)((ty0 + (bit >> 3)) * tilesX + tx0 + (bit & 7));
				lists[tileStart[tile] + atomicAdd(&tileFill[tile], 1u)] = tri;
			}
		}
	}
	__syncthreads();
	const uint32_t n = nBig;
	for(uint32_t k = 0; k < n; k++)
	{
		const uint32_t t = bigTri[k], r0 = bigR0[k], r1 = bigR1[k];
		const int tx0 = (int)(r0 & 0xffff), tx1 = (int)(r0 >> 16), ty0 = (int)(r1 & 0xffff), ty1 = (int)(r1 >> 16);
		const uint32_t w = (uint32_t)(tx1 - tx0 + 1), total = w * (uint32_t)(ty1 - ty0 + 1);
		for(uint32_t base = threadIdx.x; base < total; base += 4 * 128)
		{
			uint32_t tile[4], at[4];
#pragma unroll
			for(int u = 0; u < 4; u++)
			{
				const uint32_t i = base + (uint32_t)u * 128;
				tile[u] = i < total ? (uint32_t)((ty0 + (int)(i / w)) * tilesX + tx0 + (int)(i % w)) : 0xffffffffu;
			}
#pragma unroll
			for(int u = 0; u < 4; u++) at[u] = tile[u] != 0xffffffffu ? tileStart[tile[u]] + atomicAdd(&tileFill[tile[u]], 1u) : 0;
#pragma unroll
			for(int u = 0; u < 4; u++) if(tile[u] != 0xffffffffu) lists[at[u]] = t;
		}
	}
}

// one warp per tile: bitonic sort of its list (<= PS_SORT_LIMIT ids) in shared memory
__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK) tile_list_sort_kernel(const uint32_t* __restrict__ tileStart, uint32_t* __restrict__ lists, uint32_t ntiles,
                                                                                const uint32_t* __restrict__ poison, const uint32_t* __restrict__ tileOrder)
{
	__shared__ uint32_t buf[PS_WARPS_PER_BLOCK][PS_SORT_LIMIT];
	if(*poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t slot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(slot >= tileOrder[ntiles]) return;              // tiles with a list
	const uint32_t tile = tileOrder[slot];
	const uint32_t begin = tileStart[tile], n = tileStart[tile + 1] - begin;
	if(n < 2 || n > PS_SORT_LIMIT) return;
	uint32_t* a = buf[w];
	uint32_t P2 = 2;
	while(P2 < n) P2 <<= 1;
	bool sorted = true;
	for(uint32_t i = lane; i < P2; i += 32)
	{
		const uint32_t v = i < n ? lists[begin + i] : 0xffffffffu;
		a[i] = v;
		if(i > 0 && i < n && lists[begin + i - 1] > v) sorted = false;
	}
	__syncwarp();
	if(__all_sync(PS_FULL, sorted)) return;
	for(uint32_t k = 2; k <= P2; k <<= 1)
		for(uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for(uint32_t t = lane; t < (P2 >> 1); t += 32)
			{
				// the t-th compare-exchange pair of this stage: i has bit j clear
				const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				const uint32_t x = a[i], y = a[i | j];
				const bool up = 0 == (i & k);
				if((x > y) == up) { a[i] = y; a[i | j] = x; }
			}
			__syncwarp();
		}
	for(uint32_t i = lane; i < n; i += 32) lists[begin + i] = a[i];
}

#define PS_SORT_ITEMS_PER_WARP 1024

// counts[digit * nwarps + warp]
__global__ void __launch_bounds__(128) sort_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ counts, uint32_t nwarps)
{
	__shared__ uint32_t hist[PS_WARPS_PER_BLOCK][256];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t wg = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	for(int d = lane; d < 256; d += 32) hist[w][d] = 0;
	__syncwarp();
	if(wg < nwarps)
	{
		const uint32_t base = wg * PS_SORT_ITEMS_PER_WARP;
		for(int s = 0; s < PS_SORT_ITEMS_PER_WARP / 32; s++)
		{
			const uint32_t i = base + s * 32 + lane;
			if(i < n) atomicAdd(&hist[w][(keys[i] >> shift) & 0xff], 1u);
		}
		__syncwarp();
		for(int d = lane; d < 256; d += 32) counts[(size_t)d * nwarps + wg] = hist[w][d];
	}
}

__global__ void __launch_bounds__(128) sort_scatter_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                          uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n, int shift,
                                                          const uint32_t* __restrict__ offsets, uint32_t nwarps)
{
	__shared__ uint32_t off[PS_WARPS_PER_BLOCK][256];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t wg = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(wg >= nwarps) return;
	for(int d = lane; d < 256; d += 32) off[w][d] = offsets[(size_t)d * nwarps + wg];
	__syncwarp();
	const uint32_t base = wg * PS_SORT_ITEMS_PER_WARP;
	const uint32_t ltMask = (1u << lane) - 1;
	for(int s = 0; s < PS_SORT_ITEMS_PER_WARP / 32; s++)
	{
		const uint32_t i = base + s * 32 + lane;
		const bool valid = i < n;
		uint32_t k = 0, v = 0;
		if(valid) { k = keysIn[i]; v = valsIn[i]; }
		const uint32_t d = valid ? ((k >> shift) & 0xff) : 0x100u; // invalid lanes form their own group
		const uint32_t peers = __match_any_sync(PS_FULL, d);
		const uint32_t rank = __popc(peers & ltMask);
		uint32_t pos = 0;
		if(valid) pos = off[w][d] + rank; // lanes ascend with item index => stable
		__syncwarp();
		if(valid && 0 == rank) off[w][d] += __popc(peers);
		__syncwarp();
		if(valid) { keysOut[pos] = k; valsOut[pos] = v; }
	}
}

// ======================================================================================================================
// clears (pipeline.cpp:334-342 -> fbo.cpp:332-371)
// ======================================================================================================================

// clear16: fills m_bytes/16 whole quads with the value
__global__ void clear_depth_kernel(float4* __restrict__ buf, size_t quads, float v)
{
	const float4 q = make_float4(v, v, v, v);
	for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) buf[i] = q;
}
// clear4: every pixel of every buffer row EXCEPT the last one (fbo.cpp:336: `y < m_height - 1`)
__global__ void clear_colour_kernel(uint8_t* __restrict__ buf, int width, int rows, int scanline, uint32_t v)
{
	const size_t total = (size_t)width * rows;
	if(scanline == width * 4 && 0 == ((uintptr_t)buf & 15))
	{
		// the rows are contiguous (the display targets, fbo.cpp:104-105): one flat fill, 128-bit stores
		const size_t quads = total >> 2;
		const uint4 q = make_uint4(v, v, v, v);
		for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) ((uint4*)buf)[i] = q;
		if(0 == blockIdx.x && threadIdx.x < (total & 3)) ((uint32_t*)buf)[(quads << 2) + threadIdx.x] = v;
		return;
	}
	for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
	{
		const size_t y = i / width, x = i - y * width;
		*(uint32_t*)(buf + y * scanline + x * 4) = v;
	}
}

// ======================================================================================================================
// post-processing (post.cpp:3-19): device functors over the finished colour target
// ======================================================================================================================

// PP_DepthofField (src/test2/testpost.cpp:9-43), the reference's stub: paddb 50 on 8 bytes = two pixels per step, x += 2.
// A pixel is touched once by its own row, and — odd widths only — pixel 0 of a row once more by the last step of the row
// before it in memory (scanline = width * 4, so "the pixel behind the end" is the next row's first).
struct PostDepthofField
{
	PS_D static uint32_t apply(uint32_t bgra, float /*depth*/, int x, int y, int width, int /*height*/)
	{
		const int times = 1 + ((width & 1) && 0 == x && y > 0 ? 1 : 0);
		const uint32_t add = 50u * (uint32_t)times;
		uint32_t out = 0;
#pragma unroll
		for(int ch = 0; ch < 4; ch++) out |= ((((bgra >> (8 * ch)) & 0xff) + add) & 0xff) << (8 * ch);
		return out;
	}
};

// y = memory row of the colour target (top-down, fbo.cpp:104-105); depth is bottom-up
template<class POST>
__global__ void __launch_bounds__(256) post_process_kernel(TargetDesc colour, TargetDesc depth)
{
	const size_t total = (size_t)colour.width * colour.height;
	for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
	{
		const int y = (int)(i / colour.width), x = (int)(i - (size_t)y * colour.width);
		uint32_t* px = (uint32_t*)(colour.ptr + (size_t)y * colour.scanline) + x;
		const int dy = colour.height - 1 - y;
		const float d = (depth.ptr && dy >= 0 && dy < depth.height && x < depth.width) ? *((const float*)(depth.ptr + (size_t)dy * depth.scanline) + x) : 1.0f;
		*px = POST::apply(*px, d, x, y, colour.width, colour.height);
	}
}

// ======================================================================================================================
// tile raster + shade
// ======================================================================================================================

// (blend4 — fbo.cpp:208-229 — lives in shaders.cuh next to FragmentProcessorOutput: host/device code, pinned on the host too)

struct TileSmem
{
	float depth[PS_TILE][PS_TILE];
	uint32_t colour[PS_TILE][PS_TILE];
	TriHeader hdr[32];
	int spanL[PS_TILE][32];
	int spanR[PS_TILE][32];
	uint8_t spanE[PS_TILE][32];
	uint32_t rowMask[PS_TILE];
	uint32_t triId[32];
};

template<class PROG>
__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK) tile_raster_shade_immediate_kernel(const __grid_constant__ DrawParams P,
                                                                                  const uint32_t* __restrict__ tileStart,
                                                                                  const uint32_t* __restrict__ sortedTris)
{
	constexpr int NV = PROG::NV;
	typedef typename PROG::I IP;
	__shared__ TileSmem smem[PS_WARPS_PER_BLOCK];
	if(*P.poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int tileSlot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(tileSlot >= (int)P.tileOrder[P.tilesX * P.tilesY]) return;   // tiles with a list
	const int tile = (int)P.tileOrder[tileSlot];         // longest lists first (tile_scan_kernel)
	const uint32_t listBegin = tileStart[tile], listEnd = tileStart[tile + 1];
	if(listBegin == listEnd) return;
	TileSmem& S = smem[w];

	const int tx0 = (tile % P.tilesX) * PS_TILE, ty0 = (tile / P.tilesX) * PS_TILE;
	const int rr = lane >> 1, seg = lane & 1;          // phase-B ownership: row rr, pixels [sx0, sx0+7]
	const int y = ty0 + rr, sx0 = tx0 + seg * PS_SEG;
	const bool testDepth = 0 != (P.behavior & PS_BEHAVIOR_TEST_DEPTH);
	const bool updateDepth = 0 != (P.behavior & PS_BEHAVIOR_UPDATE_DEPTH);
	const bool useDepth = testDepth || updateDepth;
	const bool alphaBlend = 0 != (P.behavior & PS_BEHAVIOR_ALPHABLEND);

	// ---- stage the tile: each lane loads the 8-pixel segment it owns (fbo.cpp:98-110: colour top-down, depth bottom-up)
	const bool depthRowOk = y < P.depth.height, colourRowOk = y < P.colour.height;
	uint8_t* depthRow = P.depth.ptr + (size_t)(P.depth.topDown ? P.depth.height - 1 - y : y) * P.depth.scanline;
	uint8_t* colourRow = P.colour.ptr + (size_t)(P.colour.topDown ? P.colour.height - 1 - y : y) * P.colour.scanline;
#pragma unroll
	for(int i = 0; i < PS_SEG; i++)
	{
		const int x = sx0 + i;
		float d = 1.0f;
		uint32_t c = 0;
		if(useDepth && depthRowOk && x < P.depth.width) d = *(const float*)(depthRow + (size_t)x * 4);
		if(colourRowOk && x < P.colour.width) c = *(const uint32_t*)(colourRow + (size_t)x * 4);
		S.depth[rr][seg * PS_SEG + i] = d;
		S.colour[rr][seg * PS_SEG + i] = c;
	}
	if(lane < PS_TILE) S.rowMask[lane] = 0;
	__syncwarp();

	unsigned tested = 0, shaded = 0;
	bool depthDirty = false, colourDirty = false;
	// the reference reads a clamped column/row when the viewport exceeds the depth target (fbo.cpp:101,150); that
	// behaviour is not reproducible tile-locally, such fragments are dropped (DESIGN.md "divergences")
	const int depthLimitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	const bool rowDrawable = !useDepth || depthRowOk;

	for(uint32_t chunk = listBegin; chunk < listEnd; chunk += 32)
	{
		// ---- phase A: lane = triangle of the chunk (submission order). Evaluate its rows inside this tile. ----
		const uint32_t li = chunk + lane;
		if(li < listEnd)
		{
			const uint32_t tri = sortedTris[li];
			S.triId[lane] = tri;
			const uint4* src = (const uint4*)(P.hdr + tri);
			uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
			uint4* dst = (uint4*)&S.hdr[lane];
			dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3;
			TriHeader h;
			*(uint4*)&h = q0; *((uint4*)&h + 1) = q1; *((uint4*)&h + 2) = q2; *((uint4*)&h + 3) = q3;
			const float vx[3] = { h.vx0, h.vx1, h.vx2 }, vy[3] = { h.vy0, h.vy1, h.vy2 };
			int r0 = (int)(h.rows & 0xffff), r1 = (int)(h.rows >> 16);
			r0 = max(max(r0, ty0), P.band0);
			r1 = min(min(r1, ty0 + PS_TILE - 1), P.band1 - 1);
			for(int iy = r0; iy <= r1; iy++)
			{
				RowSpan r;
				if(!rowOf(h, vx, vy, iy, r)) continue;
				if(r.left == r.right) continue;                           // drawvao.cpp:72
				const int x1 = r.left < 0 ? 0 : r.left;
				const int x2 = r.right >= P.vpW ? P.vpW - 1 : r.right;
				if(x1 > x2 || x2 < tx0 || x1 > tx0 + PS_TILE - 1) continue;
				S.spanL[iy - ty0][lane] = r.left;
				S.spanR[iy - ty0][lane] = r.right;
				S.spanE[iy - ty0][lane] = (uint8_t)r.edges;
				atomicOr(&S.rowMask[iy - ty0], 1u << lane);
			}
		}
		__syncwarp();

		// ---- phase B: lane = (row, 8-pixel segment). Spans of my row in bit order = submission order. ----
		uint32_t mask = rowDrawable ? S.rowMask[rr] : 0;
		while(mask)
		{
			const int t = __ffs(mask) - 1;
			mask &= mask - 1;
			const int left = S.spanL[rr][t], right = S.spanR[rr][t];
			const int x1 = left < 0 ? 0 : left;
			const int x2 = right >= P.vpW ? P.vpW - 1 : right;
			const int xs = max(x1, sx0), xe = min(min(x2, sx0 + PS_SEG - 1), depthLimitX);
			if(xs > xe) continue;
			const int e = S.spanE[rr][t];
			const TriHeader& h = S.hdr[t];
			const float vx[3] = { h.vx0, h.vx1, h.vx2 }, vy[3] = { h.vy0, h.vy1, h.vy2 };

			// interpolateStartAndStep, interp.cpp:26-80
			float cl[3], cr[3];
			edgeContrib(vx, vy, e & 3, (e >> 2) & 3, (float)left, (float)y, cl);
			edgeContrib(vx, vy, (e >> 4) & 3, (e >> 6) & 3, (float)right, (float)y, cr);
			cl[0] = fmul(cl[0], h.rw0); cl[1] = fmul(cl[1], h.rw1); cl[2] = fmul(cl[2], h.rw2); // mulvec_3_4 (:40-41); lane 3 is 0*0
			cr[0] = fmul(cr[0], h.rw0); cr[1] = fmul(cr[1], h.rw1); cr[2] = fmul(cr[2], h.rw2);
			const int stepCount = right - left;
			const float rcpLen = fdiv(1.0f, (float)stepCount);                                   // :47
			float zStart = hsum4(fmul(cl[0], h.z0), fmul(cl[1], h.z1), fmul(cl[2], h.z2), 0.0f);  // :49 dot_3_4
			float zStep = hsum4(fmul(cr[0], h.z0), fmul(cr[1], h.z1), fmul(cr[2], h.z2), 0.0f);   // :50
			zStep = fmul(fsub(zStep, zStart), rcpLen);                                           // :51
			float cf2Start = hsum4(cl[0], cl[1], cl[2], 0.0f);                                   // :55-68
			float cf2Step = hsum4(cr[0], cr[1], cr[2], 0.0f);
			cf2Step = fmul(fsub(cf2Step, cf2Start), rcpLen);                                     // :72
			const int skip = x1 - left;                                                          // drawvao.cpp:90
			if(skip > 0)                                                                         // interp.cpp:74-79
			{
				cf2Start = fadd(cf2Start, fmul(cf2Step, (float)skip));
				zStart = fadd(zStart, fmul(zStep, (float)skip));
			}
			// the k-th pixel's value is k rounded additions from the span start (§9.6): replay them up to my segment
			for(int x = x1; x < xs; x++)
			{
				cf2Start = fadd(cf2Start, cf2Step);
				zStart = fadd(zStart, zStep);
			}

			F4 vStart[NV > 0 ? NV : 1], vStep[NV > 0 ? NV : 1];
			bool varyReady = false;
			for(int x = xs; x <= xe; x++)
			{
				// interpolateNextStep, interp.cpp:82-92
				const float inv = fdiv(1.0f, cf2Start);
				cf2Start = fadd(cf2Start, cf2Step);
				const float z = fmul(zStart, inv);
				zStart = fadd(zStart, zStep);
				tested++;
				const int px = x - tx0;
				const float cur = testDepth ? S.depth[rr][px] : 1.0f;            // fragthrd.cpp:217-225
				if(-1.0f < z && fsub(z, cur) < -0.0001f)                         // fragthrd.cpp:227
				{
					if(NV > 0 && !varyReady)
					{
						// IP::interpolateByContributes x2, calcStep, and the left-clip skip, then catch the chain up to x
						const F4* v = P.vary + (size_t)S.triId[t] * 3 * NV;
						F4 v0[NV > 0 ? NV : 1], v1[NV > 0 ? NV : 1], v2[NV > 0 ? NV : 1], vEnd[NV > 0 ? NV : 1];
#pragma unroll
						for(int k = 0; k < NV; k++)
						{
							const float4 a = __ldg((const float4*)(v + k)), b = __ldg((const float4*)(v + NV + k)), c = __ldg((const float4*)(v + 2 * NV + k));
							v0[k] = f4(a.x, a.y, a.z, a.w); v1[k] = f4(b.x, b.y, b.z, b.w); v2[k] = f4(c.x, c.y, c.z, c.w);
						}
						IP::interpolateByContributes(vStart, v0, v1, v2, cl[0], cl[1], cl[2]);
						IP::interpolateByContributes(vEnd, v0, v1, v2, cr[0], cr[1], cr[2]);
						IP::calcStep(vStep, vStart, vEnd, stepCount);
						if(skip > 0) IP::stepForward(vStart, vStep, skip);
						for(int k = x1; k < x; k++) IP::stepForward(vStart, vStep, 1);
						varyReady = true;
					}
					F4 frag[NV > 0 ? NV : 1];
					IP::correctInterpolation(frag, vStart, inv);
					FragmentProcessorOutput out;
					out.discarded = false; out.wrote = false; out.blendable = false; out.bgra = 0;
					PROG::F::process(frag, out, P);                              // fragthrd.cpp:231
					shaded++;
					if(P.cap && x < P.capW && y < P.capH) atomicAdd(&P.cap[(size_t)y * P.capW + x], 1u);
					if(out.wrote && colourRowOk && x < P.colour.width)
					{
						// FBOBridge::write4 -> blend4 under ALPHABLEND, FBOBridge::write -> plain store (fragthrd.cpp:54-82)
						S.colour[rr][px] = (out.blendable && alphaBlend) ? blend4(out.bgra, S.colour[rr][px]) : out.bgra;
						colourDirty = true;
					}
					if(!out.discarded && updateDepth)                            // fragthrd.cpp:234-237
					{
						S.depth[rr][px] = z;
						depthDirty = true;
					}
				}
				if(NV > 0 && varyReady) IP::stepForward(vStart, vStep, 1);       // interp.cpp:88
			}
		}
		__syncwarp();
		if(lane < PS_TILE) S.rowMask[lane] = 0;
		__syncwarp();
	}

	// ---- write back the segments this lane dirtied
	if(depthDirty)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.depth.width) *(float*)(depthRow + (size_t)(sx0 + i) * 4) = S.depth[rr][seg * PS_SEG + i];
	}
	if(colourDirty)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.colour.width) *(uint32_t*)(colourRow + (size_t)(sx0 + i) * 4) = S.colour[rr][seg * PS_SEG + i];
	}
	const unsigned long long t = warpSumU64(tested), s = warpSumU64(shaded);
	if(0 == lane)
	{
		if(t) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_tested, t);
		if(s) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_shaded, s);
	}
}

// ======================================================================================================================
// tile raster + shade, ordered (one kernel): lane = span for set-up AND depth test, lane = survivor for shading, colour
// committed in submission order in shared memory. Used when the draw blends (blend4 is not commutative).
// ======================================================================================================================

#define PS_QCAP 192   // survivor queue entries per warp; a flush leaves < 32 behind, a round admits what fits

struct TileSmem2
{
	float depth[PS_TILE][PS_TILE];
	uint32_t colour[PS_TILE][PS_TILE];
	TriHeader hdr[32];          // the chunk's triangles, submission order
	uint32_t triId[32];
	uint32_t spanBase[33];      // exclusive scan of "rows of triangle t inside this tile"
	int triRow0[32];
	uint32_t rExtent[32];       // spans of the current pass: xs | xe << 4 (tile-relative, inclusive); slot order = submission order
	uint32_t rowMask[PS_TILE];  // slots of the current pass per tile row
	// survivors of the depth test waiting to be shaded; per pixel the queue order is submission order
	uint32_t qTri[PS_QCAP];
	int qLeft[PS_QCAP], qRight[PS_QCAP];   // RESULT_ROW::left / right (unclamped)
	float qInv[PS_QCAP];                   // 1 / correctionFactor2 at the pixel (interp.cpp:85)
	uint32_t qMisc[PS_QCAP];               // px | row << 4 | edges << 8
	uint32_t qCount;
};

// Phase C: shade queue entries [0, n) in batches of 32 (only full batches unless `final`), keep the remainder.
// One copy of the fragment functor in the kernel (the instruction footprint matters: see profiles/).
template<class PROG>
__device__ __noinline__ void shadeSurvivors(const DrawParams& P, TileSmem2& S, int tx0, int ty0, bool final, unsigned& shadedOut, bool& colourDirtyOut)
{
	constexpr int NV = PROG::NV;
	typedef typename PROG::I IP;
	const int lane = threadIdx.x & 31;
	const bool alphaBlend = 0 != (P.behavior & PS_BEHAVIOR_ALPHABLEND);
	const uint32_t n = S.qCount;
	uint32_t head = 0;
	const uint32_t ltMask = (1u << lane) - 1;
	unsigned shaded = 0;
	bool colourDirty = false;
	while(head + 32 <= n || (final && head < n))
	{
		const uint32_t i = head + lane;
		const bool active = i < n;
		bool wants = false;
		uint32_t bgra = 0, pix = 0x1000u + lane;
		bool blend = false;
		if(active)
		{
			const uint32_t tri = S.qTri[i];
			const int left = S.qLeft[i], right = S.qRight[i];
			const float inv = S.qInv[i];
			const uint32_t misc = S.qMisc[i];
			const int px = (int)(misc & 15), row = (int)((misc >> 4) & 15), e = (int)((misc >> 8) & 0xff);
			const int x = tx0 + px, y = ty0 + row;
			const uint4* src = (const uint4*)(P.hdr + tri);
			const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
			const float vx[3] = { __uint_as_float(q0.x), __uint_as_float(q0.z), __uint_as_float(q1.x) };
			const float vy[3] = { __uint_as_float(q0.y), __uint_as_float(q0.w), __uint_as_float(q1.y) };
			const float rw0 = __uint_as_float(q1.z), rw1 = __uint_as_float(q1.w), rw2 = __uint_as_float(q2.x);
			F4 frag[NV > 0 ? NV : 1];
			if(NV > 0)
			{
				// interpolateStartAndStep, interp.cpp:26-80 (the varyings' half; the depth half ran with the span set-up)
				float cl[3], cr[3];
				edgeContrib(vx, vy, e & 3, (e >> 2) & 3, (float)left, (float)y, cl);
				edgeContrib(vx, vy, (e >> 4) & 3, (e >> 6) & 3, (float)right, (float)y, cr);
				cl[0] = fmul(cl[0], rw0); cl[1] = fmul(cl[1], rw1); cl[2] = fmul(cl[2], rw2);
				cr[0] = fmul(cr[0], rw0); cr[1] = fmul(cr[1], rw1); cr[2] = fmul(cr[2], rw2);
				const int stepCount = right - left;
				const int x1 = left < 0 ? 0 : left;
				const int skip = x1 - left;
				const F4* v = P.vary + (size_t)tri * 3 * NV;
				F4 vStart[NV > 0 ? NV : 1], vStep[NV > 0 ? NV : 1];
				// every varying is an independent float4 (the IP's methods are per-field loops, tex1light1.cpp:60-135): one at a
				// time keeps the three vertex values of only one field live
				typedef InterpolationProcessorVec4<1> IP1;
#pragma unroll
				for(int k = 0; k < NV; k++)
				{
					const float4 a = __ldg((const float4*)(v + k)), b = __ldg((const float4*)(v + NV + k)), c = __ldg((const float4*)(v + 2 * NV + k));
					const F4 v0 = f4(a.x, a.y, a.z, a.w), v1 = f4(b.x, b.y, b.z, b.w), v2 = f4(c.x, c.y, c.z, c.w);
					F4 vEnd;
					IP1::interpolateByContributes(&vStart[k], &v0, &v1, &v2, cl[0], cl[1], cl[2]);
					IP1::interpolateByContributes(&vEnd, &v0, &v1, &v2, cr[0], cr[1], cr[2]);
					IP1::calcStep(&vStep[k], &vStart[k], &vEnd, stepCount);
				}
				if(skip > 0) IP::stepForward(vStart, vStep, skip);               // interp.cpp:74-79
				for(int k = x1; k < x; k++) IP::stepForward(vStart, vStep, 1);    // interp.cpp:88, one rounded add per pixel
				IP::correctInterpolation(frag, vStart, inv);
			}
			FragmentProcessorOutput out;
			out.discarded = false; out.wrote = false; out.blendable = false; out.bgra = 0;
			PROG::F::process(frag, out, P);                                      // fragthrd.cpp:231
			shaded++;
			if(P.cap && x < P.capW && y < P.capH) atomicAdd(&P.cap[(size_t)y * P.capW + x], 1u);
			if(out.wrote && y < P.colour.height && x < P.colour.width)
			{
				wants = true; bgra = out.bgra; pix = (uint32_t)(row * PS_TILE + px);
				blend = out.blendable && alphaBlend;
			}
		}
		// ordered commit: entries of one pixel are applied in queue order (= submission order)
		const uint32_t peers = __match_any_sync(PS_FULL, pix);
		const int rank = __popc(peers & ltMask);
		const int maxRank = __reduce_max_sync(PS_FULL, wants ? rank : 0);
		for(int r = 0; r <= maxRank; r++)
		{
			if(wants && rank == r)
			{
				uint32_t* dst = &S.colour[0][0] + pix;
				// FBOBridge::write4 -> blend4 under ALPHABLEND, FBOBridge::write -> plain store (fragthrd.cpp:54-82)
				*dst = blend ? blend4(bgra, *dst) : bgra;
			}
			__syncwarp();
		}
		if(__any_sync(PS_FULL, wants)) colourDirty = true;
		head += 32;
	}
	if(head > 0)
	{
		// keep the remainder (< 32 entries) at the front of the queue
		const uint32_t rem = n > head ? n - head : 0;
		uint32_t a = 0, d = 0; int b = 0, c = 0; float f = 0;
		if((uint32_t)lane < rem) { a = S.qTri[head + lane]; b = S.qLeft[head + lane]; c = S.qRight[head + lane]; f = S.qInv[head + lane]; d = S.qMisc[head + lane]; }
		__syncwarp();
		if((uint32_t)lane < rem) { S.qTri[lane] = a; S.qLeft[lane] = b; S.qRight[lane] = c; S.qInv[lane] = f; S.qMisc[lane] = d; }
		if(0 == lane) S.qCount = rem;
		__syncwarp();
	}
	shadedOut += shaded;
	colourDirtyOut = colourDirtyOut || colourDirty;
}

template<class PROG>
__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK) tile_raster_shade_ordered_kernel(const __grid_constant__ DrawParams P,
                                                                                  const uint32_t* __restrict__ tileStart,
                                                                                  const uint32_t* __restrict__ sortedTris)
{
	__shared__ TileSmem2 smem[PS_WARPS_PER_BLOCK];
	if(*P.poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int tileSlot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(tileSlot >= (int)P.tileOrder[P.tilesX * P.tilesY]) return;   // tiles with a list
	const int tile = (int)P.tileOrder[tileSlot];         // longest lists first (tile_scan_kernel)
	const uint32_t listBegin = tileStart[tile], listEnd = tileStart[tile + 1];
	if(listBegin == listEnd) return;
	TileSmem2& S = smem[w];

	const int tx0 = (tile % P.tilesX) * PS_TILE, ty0 = (tile / P.tilesX) * PS_TILE;
	const int rr = lane >> 1, seg = lane & 1;          // staging / write-back ownership: row rr, pixels [sx0, sx0+7]
	const int y = ty0 + rr, sx0 = tx0 + seg * PS_SEG;
	const bool testDepth = 0 != (P.behavior & PS_BEHAVIOR_TEST_DEPTH);
	const bool updateDepth = 0 != (P.behavior & PS_BEHAVIOR_UPDATE_DEPTH);
	const bool useDepth = testDepth || updateDepth;

	// ---- stage the tile: each lane loads one 8-pixel segment (fbo.cpp:98-110: colour top-down, depth bottom-up)
	const bool depthRowOk = y < P.depth.height, colourRowOk = y < P.colour.height;
	uint8_t* depthRow = P.depth.ptr + (size_t)(P.depth.topDown ? P.depth.height - 1 - y : y) * P.depth.scanline;
	uint8_t* colourRow = P.colour.ptr + (size_t)(P.colour.topDown ? P.colour.height - 1 - y : y) * P.colour.scanline;
#pragma unroll
	for(int i = 0; i < PS_SEG; i++)
	{
		const int x = sx0 + i;
		float d = 1.0f;
		uint32_t c = 0;
		if(useDepth && depthRowOk && x < P.depth.width) d = *(const float*)(depthRow + (size_t)x * 4);
		if(colourRowOk && x < P.colour.width) c = *(const uint32_t*)(colourRow + (size_t)x * 4);
		S.depth[rr][seg * PS_SEG + i] = d;
		S.colour[rr][seg * PS_SEG + i] = c;
	}
	if(lane < PS_TILE) S.rowMask[lane] = 0;
	if(0 == lane) S.qCount = 0;
	__syncwarp();

	unsigned tested = 0, shaded = 0;
	bool depthWrote = false, colourDirty = false;
	// the reference reads a clamped column/row when the viewport exceeds the depth target (fbo.cpp:101,150); that
	// behaviour is not reproducible tile-locally, such fragments are dropped (DESIGN.md "divergences")
	const int depthLimitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	const int depthLimitY = useDepth ? P.depth.height - 1 : 0x7fffffff;
	const int tileX1 = tx0 + PS_TILE - 1;
	const uint32_t ltMask = (1u << lane) - 1;

	for(uint32_t chunk = listBegin; chunk < listEnd; chunk += 32)
	{
		// ---- A1: lane = triangle of the chunk. Header to shared memory; rows of the triangle inside tile and band. ----
		const uint32_t li = chunk + lane;
		int nrows = 0, r0 = 0;
		if(li < listEnd)
		{
			const uint32_t tri = sortedTris[li];
			S.triId[lane] = tri;
			const uint4* src = (const uint4*)(P.hdr + tri);
			const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
			uint4* dst = (uint4*)&S.hdr[lane];
			dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3;
			r0 = max(max((int)(q3.x & 0xffff), ty0), P.band0);
			const int r1 = min(min(min((int)(q3.x >> 16), ty0 + PS_TILE - 1), P.band1 - 1), depthLimitY);
			nrows = r1 >= r0 ? r1 - r0 + 1 : 0;
		}
		uint32_t incl = (uint32_t)nrows;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		S.spanBase[lane] = incl - (uint32_t)nrows;
		S.triRow0[lane] = r0;
		const uint32_t total = __shfl_sync(PS_FULL, incl, 31);
		__syncwarp();

		for(uint32_t s0 = 0; s0 < total; s0 += 32)
		{
			// ---- A2: lane = span (triangle, row), dense. RESULT_ROW + interpolateStartAndStep's depth half. ----
			const uint32_t s = s0 + lane;
			bool valid = false;
			int row = 0, xs = 0, xe = -1, left = 0, right = 0, edges = 0;
			uint32_t tri = 0;
			float cf2 = 0, cf2Step = 0, z0 = 0, zStep = 0;
			if(s < total)
			{
				int t = 0; // the last triangle whose base <= s
#pragma unroll
				for(int b = 16; b > 0; b >>= 1)
					if(t + b < 32 && S.spanBase[t + b] <= s) t += b;
				const int iy = S.triRow0[t] + (int)(s - S.spanBase[t]);
				const TriHeader& h = S.hdr[t];
				const float vx[3] = { h.vx0, h.vx1, h.vx2 }, vy[3] = { h.vy0, h.vy1, h.vy2 };
				RowSpan r;
				if(rowOf(h, vx, vy, iy, r) && r.left != r.right)                  // drawvao.cpp:72
				{
					const int x1 = r.left < 0 ? 0 : r.left;                       // RESULT_ROW::leftClamped
					const int x2 = r.right >= P.vpW ? P.vpW - 1 : r.right;        // RESULT_ROW::rightClamped
					const int xsA = max(x1, tx0), xeA = min(min(x2, tileX1), depthLimitX);
					if(x1 <= x2 && xsA <= xeA)
					{
						const int e = r.edges;
						// interpolateStartAndStep, interp.cpp:26-80
						float cl[3], cr[3];
						edgeContrib(vx, vy, e & 3, (e >> 2) & 3, (float)r.left, (float)iy, cl);
						edgeContrib(vx, vy, (e >> 4) & 3, (e >> 6) & 3, (float)r.right, (float)iy, cr);
						cl[0] = fmul(cl[0], h.rw0); cl[1] = fmul(cl[1], h.rw1); cl[2] = fmul(cl[2], h.rw2); // mulvec_3_4 (:40-41); lane 3 is 0*0
						cr[0] = fmul(cr[0], h.rw0); cr[1] = fmul(cr[1], h.rw1); cr[2] = fmul(cr[2], h.rw2);
						const float rcpLen = fdiv(1.0f, (float)(r.right - r.left));                          // :47
						z0 = hsum4(fmul(cl[0], h.z0), fmul(cl[1], h.z1), fmul(cl[2], h.z2), 0.0f);            // :49 dot_3_4
						zStep = hsum4(fmul(cr[0], h.z0), fmul(cr[1], h.z1), fmul(cr[2], h.z2), 0.0f);         // :50
						zStep = fmul(fsub(zStep, z0), rcpLen);                                               // :51
						cf2 = hsum4(cl[0], cl[1], cl[2], 0.0f);                                              // :55-68
						cf2Step = hsum4(cr[0], cr[1], cr[2], 0.0f);
						cf2Step = fmul(fsub(cf2Step, cf2), rcpLen);                                          // :72
						const int skip = x1 - r.left;                                                        // drawvao.cpp:90
						if(skip > 0)                                                                         // interp.cpp:74-79
						{
							cf2 = fadd(cf2, fmul(cf2Step, (float)skip));
							z0 = fadd(z0, fmul(zStep, (float)skip));
						}
						// the k-th pixel's value is k rounded additions from the span start (§9.6): replay them up to the tile
						for(int x = x1; x < xsA; x++)
						{
							cf2 = fadd(cf2, cf2Step);
							z0 = fadd(z0, zStep);
						}
						valid = true;
						row = iy - ty0; xs = xsA - tx0; xe = xeA - tx0;
						left = r.left; right = r.right; edges = e; tri = S.triId[t];
						S.rExtent[lane] = (uint32_t)xs | ((uint32_t)xe << 4);
						atomicOr(&S.rowMask[row], 1u << lane);
					}
				}
			}
			__syncwarp();
			// earlier spans of this pass that touch a pixel of mine must be tested before me (§9.7)
			uint32_t deps = 0;
			if(valid)
			{
				uint32_t m = S.rowMask[row] & ltMask;
				while(m)
				{
					const int j = __ffs(m) - 1;
					m &= m - 1;
					const uint32_t o = S.rExtent[j];
					if((int)(o & 15) <= xe && xs <= (int)(o >> 4)) deps |= 1u << j;
				}
			}
			__syncwarp();
			if(lane < PS_TILE) S.rowMask[lane] = 0;
			__syncwarp();

			// ---- B: lane = span still. Rounds: a span runs when every earlier overlapping span has run and the queue has room.
			uint32_t pending = __ballot_sync(PS_FULL, valid);
			while(pending)
			{
				const bool ready = valid && 0 == (deps & pending);
				uint32_t room = (uint32_t)(ready ? xe - xs + 1 : 0);
#pragma unroll
				for(int d = 1; d < 32; d <<= 1)
				{
					const uint32_t t = __shfl_up_sync(PS_FULL, room, d);
					if(lane >= d) room += t;
				}
				const bool admit = ready && room <= PS_QCAP - S.qCount;
				__syncwarp();
				if(admit)
				{
					for(int px = xs; px <= xe; px++)
					{
						// interpolateNextStep, interp.cpp:82-92
						const float inv = fdiv(1.0f, cf2);
						cf2 = fadd(cf2, cf2Step);
						const float z = fmul(z0, inv);
						z0 = fadd(z0, zStep);
						tested++;
						const float cur = testDepth ? S.depth[row][px] : 1.0f;           // fragthrd.cpp:217-225
						if(-1.0f < z && fsub(z, cur) < -0.0001f)                         // fragthrd.cpp:227
						{
							// no functor on this path discards, so the depth write does not wait for the shading (fragthrd.cpp:234-237)
							if(updateDepth) { S.depth[row][px] = z; depthWrote = true; }
							const uint32_t q = atomicAdd(&S.qCount, 1u);
							S.qTri[q] = tri;
							S.qLeft[q] = left; S.qRight[q] = right;
							S.qInv[q] = inv;
							S.qMisc[q] = (uint32_t)px | ((uint32_t)row << 4) | ((uint32_t)edges << 8);
						}
					}
					valid = false;
				}
				__syncwarp();
				pending &= ~__ballot_sync(PS_FULL, admit);
				if(S.qCount >= 32) shadeSurvivors<PROG>(P, S, tx0, ty0, false, shaded, colourDirty);
			}
		}
	}
	shadeSurvivors<PROG>(P, S, tx0, ty0, true, shaded, colourDirty);

	// ---- write back
	if(__any_sync(PS_FULL, depthWrote) && depthRowOk)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.depth.width) *(float*)(depthRow + (size_t)(sx0 + i) * 4) = S.depth[rr][seg * PS_SEG + i];
	}
	if(colourDirty && colourRowOk)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.colour.width) *(uint32_t*)(colourRow + (size_t)(sx0 + i) * 4) = S.colour[rr][seg * PS_SEG + i];
	}
	const unsigned long long t = warpSumU64(tested), sh = warpSumU64(shaded);
	if(0 == lane)
	{
		if(t) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_tested, t);
		if(sh) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_shaded, sh);
	}
}


// ======================================================================================================================
// the split path (default): raster + depth kernel -> survivor stream in HBM -> shade kernel
//
// Without blending only the LAST survivor of a pixel decides its colour, so shading needs no order at all: the raster
// kernel resolves the order-dependent part (the depth test with its dead band, §9.7) tile by tile and appends every
// survivor to a stream; when a tile is finished it publishes, per pixel, which record was the last one. The shade kernel
// is then a flat loop over the stream with every lane busy, each record running the fragment functor exactly once
// (fragthrd.cpp:231) and only the winner storing its colour. Two kernels also keep each instruction footprint inside the
// SM's instruction cache, which the one-kernel version did not (profiles/r01c).
// ======================================================================================================================

PS_D uint32_t edgeCode3(int v0, int v1) { return (uint32_t)(v0 * 2 + (v1 > v0 ? v1 - 1 : v1)); }   // ordered pair of distinct vertex ids -> 0..5
PS_D void edgeDecode3(uint32_t c, int& v0, int& v1) { v0 = (int)(c >> 1); const int t = (int)(c & 1); v1 = t + (t >= v0 ? 1 : 0); }

#define PS_RQCAP 64   // raster kernel's staging queue: < 32 left by a flush + at most 32 pushed by one pixel batch
#define PS_HIZ_MARGIN 0.00002f   // slack of the conservative span reject (chain estimate + approximate reciprocal err << this)

struct RasterSmem
{
	float depth[PS_TILE * PS_TILE];
	uint32_t lastIdx[PS_TILE * PS_TILE];   // stream index of the last survivor of each pixel
	float segMax[32];                      // upper bound of the depth of every 8-pixel row segment (index = row * 2 + segment)
	TriHeader hdr[32];                     // the chunk's triangles, submission order
	uint32_t triId[32];
	uint32_t spanBase[33];
	int triRow0[32];
	// spans that reach this tile, waiting for their set-up (ring of 64; order = submission order on every row)
	uint32_t lInfo[64];                    // chunk lane of the triangle | row << 5 | RowSpan::edges << 9
	int lLeft[64], lRight[64];
	// slots of the current pass: the non-empty spans, compacted, in list order
	float rCf2[32], rCf2Step[32], rZ[32], rZStep[32];
	int rLeft[32], rRight[32];
	uint32_t rMisc[32];                    // xs | row << 4 | edge code << 8 (xs, row tile-relative)
	uint32_t rTri[32];
	uint32_t pixBase[32];                  // first pixel index of each slot in the pass
	// staging queue of survivors
	uint32_t qTri[PS_RQCAP];
	int qLeft[PS_RQCAP], qRight[PS_RQCAP];
	float qInv[PS_RQCAP];
	uint32_t qMisc[PS_RQCAP];              // px | row << 4 | edge code << 8
};

// append queue entries [0, n) (n <= 32) to the stream; remember the last record of every pixel
PS_D void flushSurvivors(const SurvivorStream& Q, RasterSmem& S, int lane, uint32_t n, int tx0, int ty0)
{
	uint32_t base = 0;
	if(0 == lane) base = atomicAdd(Q.count, n);
	base = __shfl_sync(PS_FULL, base, 0);
	if((uint32_t)lane < n)
	{
		const uint32_t i = base + lane;
		const uint32_t m = S.qMisc[lane];
		if(i < Q.capacity)
		{
			Q.tri[i] = S.qTri[lane]; Q.left[i] = S.qLeft[lane]; Q.right[i] = S.qRight[lane]; Q.inv[i] = S.qInv[lane];
			Q.misc[i] = (uint32_t)(tx0 + (int)(m & 15)) | ((uint32_t)(ty0 + (int)((m >> 4) & 15)) << 13) | ((m >> 8) << 26);
		}
		// queue order = submission order inside a pixel and bases grow with time: the latest record has the highest index
		atomicMax(&S.lastIdx[m & 0xff], i + 1);
	}
	__syncwarp();
}

struct RasterCtx
{
	int lane, tx0, ty0, tileX1, depthLimitX, vpW;
	bool testDepth, updateDepth;
	uint32_t ltMask;
	// warp-uniform running state
	uint32_t qCount;
	unsigned survived;
	// per lane
	unsigned tested;
	bool depthWrote;
};

// One pass over n <= 32 waiting spans (ring positions head .. head + n - 1):
//   Y  lane = span: the depth half of interpolateStartAndStep (interp.cpp:26-80), chains advanced to the tile's edge,
//      then a conservative whole-span depth reject against the segment bounds
//   B  lane = pixel of a surviving span, dense: interpolateNextStep (interp.cpp:82-92) + the depth rule (fragthrd.cpp:217-237)
PS_D void rasterPass(const DrawParams& P, const SurvivorStream& Q, RasterSmem& S, RasterCtx& C, uint32_t head, uint32_t n)
{
	const int lane = C.lane;
	int len = 0;
	float cf2 = 0, cf2Step = 0, z0 = 0, zStep = 0;
	int left = 0, right = 0;
	uint32_t misc = 0, triId = 0;
	if((uint32_t)lane < n)
	{
		const uint32_t pos = (head + lane) & 63;
		const uint32_t info = S.lInfo[pos];
		left = S.lLeft[pos]; right = S.lRight[pos];
		const int t = (int)(info & 31), row = (int)((info >> 5) & 15), e = (int)(info >> 9);
		const int iy = C.ty0 + row;
		const TriHeader& h = S.hdr[t];
		const float* xy = &h.vx0;
		const int x1 = left < 0 ? 0 : left;                           // RESULT_ROW::leftClamped
		const int x2 = right >= C.vpW ? C.vpW - 1 : right;            // RESULT_ROW::rightClamped
		const int xsA = max(x1, C.tx0), xeA = min(min(x2, C.tileX1), C.depthLimitX);
		// interpolateStartAndStep, interp.cpp:26-80
		float cl[3], cr[3];
		edgeContribXY(xy, e & 3, (e >> 2) & 3, (float)left, (float)iy, cl);
		edgeContribXY(xy, (e >> 4) & 3, (e >> 6) & 3, (float)right, (float)iy, cr);
		cl[0] = fmul(cl[0], h.rw0); cl[1] = fmul(cl[1], h.rw1); cl[2] = fmul(cl[2], h.rw2); // mulvec_3_4 (:40-41); lane 3 is 0*0
		cr[0] = fmul(cr[0], h.rw0); cr[1] = fmul(cr[1], h.rw1); cr[2] = fmul(cr[2], h.rw2);
		const float rcpLen = fdiv(1.0f, (float)(right - left));                               // :47
		z0 = hsum4(fmul(cl[0], h.z0), fmul(cl[1], h.z1), fmul(cl[2], h.z2), 0.0f);            // :49 dot_3_4
		zStep = hsum4(fmul(cr[0], h.z0), fmul(cr[1], h.z1), fmul(cr[2], h.z2), 0.0f);         // :50
		zStep = fmul(fsub(zStep, z0), rcpLen);                                                // :51
		cf2 = hsum4(cl[0], cl[1], cl[2], 0.0f);                                               // :55-68
		cf2Step = hsum4(cr[0], cr[1], cr[2], 0.0f);
		cf2Step = fmul(fsub(cf2Step, cf2), rcpLen);                                           // :72
		const int skip = x1 - left;                                                           // drawvao.cpp:90
		if(skip > 0)                                                                          // interp.cpp:74-79
		{
			cf2 = fadd(cf2, fmul(cf2Step, (float)skip));
			z0 = fadd(z0, fmul(zStep, (float)skip));
		}
		// the k-th pixel's value is k rounded additions from the span start (§9.6): replay them up to the tile
#pragma unroll 1
		for(int x = x1; x < xsA; x++)
		{
			cf2 = fadd(cf2, cf2Step);
			z0 = fadd(z0, zStep);
		}
		len = xeA - xsA + 1;
		misc = (uint32_t)(xsA - C.tx0) | ((uint32_t)row << 4)
		     | ((edgeCode3(e & 3, (e >> 2) & 3) | (edgeCode3((e >> 4) & 3, (e >> 6) & 3) << 3)) << 8);
		triId = S.triId[t];
		if(C.testDepth)
		{
			// Conservative whole-span reject. A fragment fails when z - cur >= -0.0001 (fragthrd.cpp:227), certainly when
			// z >= cur. Along the span z = z0_k / cf2_k is a ratio of two linear functions of k, monotone while cf2 keeps its
			// sign, so its minimum over the tile's pixels is at one of the two ends; both ends are ESTIMATED here (approximate
			// reciprocal, end of the chain in closed form; error ~1e-6) and compared with a margin against an upper bound of
			// the depths the span could meet. Rejected spans still count as tested; everything else takes the exact path.
			const float nf = (float)(len - 1);
			const float cf2e = cf2 + nf * cf2Step, z0e = z0 + nf * zStep;
			const float zs = __fdividef(z0, cf2), ze = __fdividef(z0e, cf2e);
			const int sA = (xsA - C.tx0) >> 3, sB = (xeA - C.tx0) >> 3;
			const float bound = fmaxf(S.segMax[row * 2 + sA], S.segMax[row * 2 + sB]);
			if(cf2 > 0.0f && cf2e > 0.0f && zs - PS_HIZ_MARGIN >= bound && ze - PS_HIZ_MARGIN >= bound)
			{
				C.tested += (unsigned)len;
				len = 0;
			}
		}
	}
	// compact the non-empty spans into slots; pixel index space of the pass
	const uint32_t nz = __ballot_sync(PS_FULL, len > 0);
	uint32_t pincl = (uint32_t)len;
#pragma unroll
	for(int d = 1; d < 32; d <<= 1)
	{
		const uint32_t t = __shfl_up_sync(PS_FULL, pincl, d);
		if(lane >= d) pincl += t;
	}
	const int myBase = (int)(pincl - (uint32_t)len);
	const uint32_t totalPix = __shfl_sync(PS_FULL, pincl, 31);
	if(0 == totalPix) return;
	if(len > 0)
	{
		const int slot = __popc(nz & C.ltMask);
		S.rCf2[slot] = cf2; S.rCf2Step[slot] = cf2Step; S.rZ[slot] = z0; S.rZStep[slot] = zStep;
		S.rLeft[slot] = left; S.rRight[slot] = right; S.rTri[slot] = triId; S.rMisc[slot] = misc;
		S.pixBase[slot] = (uint32_t)myBase;
	}
	__syncwarp();

	// ---- B: lane = pixel of a span, dense; spans in slot order, pixels left to right. ----
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
		uint32_t pix = 0x1000u + lane, smisc = 0;
		float z = 0, inv = 0;
		if(act)
		{
			const int k = (int)(p - S.pixBase[slot]);
			smisc = S.rMisc[slot];
			float c2 = S.rCf2[slot], zz = S.rZ[slot];
			const float c2Step = S.rCf2Step[slot], zzStep = S.rZStep[slot];
#pragma unroll 1
			for(int j = 0; j < k; j++)
			{
				c2 = fadd(c2, c2Step);
				zz = fadd(zz, zzStep);
			}
			// interpolateNextStep, interp.cpp:82-92
			inv = fdiv(1.0f, c2);
			z = fmul(zz, inv);
			pix = (((smisc >> 4) & 15) << 4) | ((smisc & 15) + (uint32_t)k);
			C.tested++;
		}
		// fragments of one pixel are tested in lane order = submission order (§9.7)
		const uint32_t peers = __match_any_sync(PS_FULL, pix);
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
			S.qTri[q] = S.rTri[slot]; S.qLeft[q] = S.rLeft[slot]; S.qRight[q] = S.rRight[slot]; S.qInv[q] = inv;
			S.qMisc[q] = pix | ((smisc >> 8) << 8);
		}
		C.qCount += __popc(b);
		__syncwarp();
		if(C.qCount >= 32)
		{
			flushSurvivors(Q, S, lane, 32, C.tx0, C.ty0);
			C.survived += 32;
			const uint32_t rem = C.qCount - 32;
			uint32_t a = 0, d = 0; int bb = 0, c = 0; float f = 0;
			if((uint32_t)lane < rem) { a = S.qTri[32 + lane]; bb = S.qLeft[32 + lane]; c = S.qRight[32 + lane]; f = S.qInv[32 + lane]; d = S.qMisc[32 + lane]; }
			__syncwarp();
			if((uint32_t)lane < rem) { S.qTri[lane] = a; S.qLeft[lane] = bb; S.qRight[lane] = c; S.qInv[lane] = f; S.qMisc[lane] = d; }
			C.qCount = rem;
			__syncwarp();
		}
	}
}

__global__ void __launch_bounds__(32 * PS_WARPS_PER_BLOCK) tile_raster_depth_kernel(const __grid_constant__ DrawParams P, const SurvivorStream Q,
                                                                                  const uint32_t* __restrict__ tileStart,
                                                                                  const uint32_t* __restrict__ sortedTris, int parts)
{
	__shared__ RasterSmem smem[PS_WARPS_PER_BLOCK];
	if(*P.poison) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	// A warp walks its tile's list as one dependent chain (~100 us on C2 whatever the load): rows are independent in this
	// rasteriser, so when there are fewer tiles than the GPU has warp slots (a sort-first band, a small target) a tile is
	// cut into `parts` (1, 2 or 4) groups of PS_TILE / parts rows, one warp each, all reading the same list.
	const int warpSlot = blockIdx.x * PS_WARPS_PER_BLOCK + w;
	if(warpSlot >= (int)P.tileOrder[P.tilesX * P.tilesY] * parts) return;   // tiles with a list
	const int tile = (int)P.tileOrder[warpSlot / parts];   // longest lists first (tile_scan_kernel)
	const int rowsPer = PS_TILE / parts, partRow0 = (warpSlot % parts) * rowsPer;
	const uint32_t listBegin = tileStart[tile], listEnd = tileStart[tile + 1];
	if(listBegin == listEnd) return;
	RasterSmem& S = smem[w];

	const int tx0 = (tile % P.tilesX) * PS_TILE, ty0 = (tile / P.tilesX) * PS_TILE;
	const int rr = lane >> 1, seg = lane & 1;          // staging / write-back ownership: row rr, pixels [sx0, sx0+7]
	const int y = ty0 + rr, sx0 = tx0 + seg * PS_SEG;
	const bool testDepth = 0 != (P.behavior & PS_BEHAVIOR_TEST_DEPTH);
	const bool updateDepth = 0 != (P.behavior & PS_BEHAVIOR_UPDATE_DEPTH);
	const bool useDepth = testDepth || updateDepth;

	const bool mine = rr >= partRow0 && rr < partRow0 + rowsPer;   // rows of the tile this warp owns (staging, write-back)
	const bool depthRowOk = y < P.depth.height;
	uint8_t* depthRow = P.depth.ptr + (size_t)(P.depth.topDown ? P.depth.height - 1 - y : y) * P.depth.scanline;
#pragma unroll
	for(int i = 0; i < PS_SEG; i++)
	{
		const int x = sx0 + i;
		float d = 1.0f;
		if(mine && useDepth && depthRowOk && x < P.depth.width) d = *(const float*)(depthRow + (size_t)x * 4);
		S.depth[rr * PS_TILE + seg * PS_SEG + i] = d;
		S.lastIdx[rr * PS_TILE + seg * PS_SEG + i] = 0;
	}
	__syncwarp();

	RasterCtx C;
	C.lane = lane; C.tx0 = tx0; C.ty0 = ty0; C.tileX1 = tx0 + PS_TILE - 1; C.vpW = P.vpW;
	// the reference reads a clamped column/row when the viewport exceeds the depth target (fbo.cpp:101,150); that
	// behaviour is not reproducible tile-locally, such fragments are dropped (DESIGN.md "divergences")
	C.depthLimitX = useDepth ? P.depth.width - 1 : 0x7fffffff;
	const int depthLimitY = useDepth ? P.depth.height - 1 : 0x7fffffff;
	C.testDepth = testDepth; C.updateDepth = updateDepth;
	C.ltMask = (1u << lane) - 1;
	C.qCount = 0; C.survived = 0; C.tested = 0; C.depthWrote = false;
	uint32_t lHead = 0, lCount = 0;                    // ring of waiting spans (warp-uniform)

	for(uint32_t chunk = listBegin; chunk < listEnd; chunk += 32)
	{
		// ---- A1: lane = triangle of the chunk. Header to shared memory; rows of the triangle inside tile and band. ----
		const uint32_t li = chunk + lane;
		int nrows = 0, r0 = 0;
		if(li < listEnd)
		{
			const uint32_t tri = sortedTris[li];
			S.triId[lane] = tri;
			const uint4* src = (const uint4*)(P.hdr + tri);
			const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2), q3 = __ldg(src + 3);
			uint4* dst = (uint4*)&S.hdr[lane];
			dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3;
			r0 = max(max((int)(q3.x & 0xffff), ty0 + partRow0), P.band0);
			const int r1 = min(min(min((int)(q3.x >> 16), ty0 + partRow0 + rowsPer - 1), P.band1 - 1), depthLimitY);
			nrows = r1 >= r0 ? r1 - r0 + 1 : 0;
		}
		uint32_t incl = (uint32_t)nrows;
#pragma unroll
		for(int d = 1; d < 32; d <<= 1)
		{
			const uint32_t t = __shfl_up_sync(PS_FULL, incl, d);
			if(lane >= d) incl += t;
		}
		S.spanBase[lane] = incl - (uint32_t)nrows;
		S.triRow0[lane] = r0;
		const uint32_t total = __shfl_sync(PS_FULL, incl, 31);
		// depths only decrease while a draw runs, so a bound refreshed once per chunk stays an upper bound
		if(testDepth)
		{
			float m = S.depth[lane * PS_SEG];
#pragma unroll
			for(int i = 1; i < PS_SEG; i++) m = fmaxf(m, S.depth[lane * PS_SEG + i]);
			S.segMax[lane] = m;
		}
		__syncwarp();

		for(uint32_t s0 = 0; s0 < total; s0 += 32)
		{
			// ---- X: lane = candidate (triangle, row), dense. RESULT_ROW; only the spans that reach the tile go on. ----
			const uint32_t s = s0 + lane;
			bool ok = false;
			uint32_t info = 0;
			RowSpan r;
			r.left = r.right = r.edges = 0;
			if(s < total)
			{
				int t = 0; // the last triangle whose base <= s
#pragma unroll
				for(int b = 16; b > 0; b >>= 1)
					if(S.spanBase[t + b] <= s) t += b;
				const int iy = S.triRow0[t] + (int)(s - S.spanBase[t]);
				if(rowOfXY(S.hdr[t], iy, r) && r.left != r.right)                 // drawvao.cpp:72
				{
					const int x1 = r.left < 0 ? 0 : r.left;
					const int x2 = r.right >= P.vpW ? P.vpW - 1 : r.right;
					const int xsA = max(x1, tx0), xeA = min(min(x2, C.tileX1), C.depthLimitX);
					ok = x1 <= x2 && xsA <= xeA;
					info = (uint32_t)t | ((uint32_t)(iy - ty0) << 5) | ((uint32_t)r.edges << 9);
				}
			}
			const uint32_t okb = __ballot_sync(PS_FULL, ok);
			if(ok)
			{
				const uint32_t pos = (lHead + lCount + __popc(okb & C.ltMask)) & 63;
				S.lInfo[pos] = info; S.lLeft[pos] = r.left; S.lRight[pos] = r.right;
			}
			lCount += __popc(okb);
			__syncwarp();
			if(lCount >= 32)
			{
				rasterPass(P, Q, S, C, lHead, 32);
				lHead = (lHead + 32) & 63;
				lCount -= 32;
			}
		}
		// the waiting spans refer to this chunk's headers: drain before the next chunk overwrites them
		if(lCount)
		{
			rasterPass(P, Q, S, C, lHead, lCount);
			lHead = 0; lCount = 0;
		}
		__syncwarp();
	}
	if(C.qCount) { flushSurvivors(Q, S, lane, C.qCount, tx0, ty0); C.survived += C.qCount; }

	// ---- write back: depth tile, and which record won each pixel
	if(__any_sync(PS_FULL, C.depthWrote) && depthRowOk && mine)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.depth.width) *(float*)(depthRow + (size_t)(sx0 + i) * 4) = S.depth[rr * PS_TILE + seg * PS_SEG + i];
	}
	if(C.survived && y < P.vpH && mine)
	{
#pragma unroll
		for(int i = 0; i < PS_SEG; i++)
			if(sx0 + i < P.vpW) Q.winner[(size_t)y * P.vpW + sx0 + i] = S.lastIdx[rr * PS_TILE + seg * PS_SEG + i];
	}
	const unsigned long long t = warpSumU64(C.tested);
	if(0 == lane)
	{
		if(t) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_tested, t);
		if(C.survived) atomicAdd(&P.stats[blockIdx.x & (PS_STATS_COPIES - 1)].fragments_shaded, (unsigned long long)C.survived);   // every survivor is shaded exactly once by shade_kernel
	}
}


// lane = survivor record, any order: varyings (interp.cpp:26-92), fragment functor (fragthrd.cpp:231), winner stores its colour
template<class PROG>
__global__ void __launch_bounds__(128) shade_kernel(const __grid_constant__ DrawParams P, const SurvivorStream Q)
{
	constexpr int NV = PROG::NV;
	if(*P.poison) return;
	const uint32_t n = min(*Q.count, Q.capacity);
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const uint32_t tri = Q.tri[i];
		const int left = Q.left[i], right = Q.right[i];
		const float inv = Q.inv[i];
		const uint32_t misc = Q.misc[i];
		const int x = (int)(misc & 0x1fff), y = (int)((misc >> 13) & 0x1fff);
		F4 frag[NV > 0 ? NV : 1];
		if(NV > 0)
		{
			const uint4* src = (const uint4*)(P.hdr + tri);
			const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
			const float vx[3] = { __uint_as_float(q0.x), __uint_as_float(q0.z), __uint_as_float(q1.x) };
			const float vy[3] = { __uint_as_float(q0.y), __uint_as_float(q0.w), __uint_as_float(q1.y) };
			const float rw0 = __uint_as_float(q1.z), rw1 = __uint_as_float(q1.w), rw2 = __uint_as_float(q2.x);
			int l0, l1, r0, r1;
			edgeDecode3((misc >> 26) & 7, l0, l1);
			edgeDecode3((misc >> 29) & 7, r0, r1);
			// interpolateStartAndStep, interp.cpp:26-80 (the varyings' half; the depth half ran in the raster kernel)
			float cl[3], cr[3];
			edgeContrib(vx, vy, l0, l1, (float)left, (float)y, cl);
			edgeContrib(vx, vy, r0, r1, (float)right, (float)y, cr);
			cl[0] = fmul(cl[0], rw0); cl[1] = fmul(cl[1], rw1); cl[2] = fmul(cl[2], rw2);
			cr[0] = fmul(cr[0], rw0); cr[1] = fmul(cr[1], rw1); cr[2] = fmul(cr[2], rw2);
			const int stepCount = right - left;
			const int x1 = left < 0 ? 0 : left;
			const int skip = x1 - left;
			const F4* v = P.vary + (size_t)tri * 3 * NV;
			F4 vStart[NV > 0 ? NV : 1], vStep[NV > 0 ? NV : 1];
			// every varying is an independent float4 (the IP's methods are per-field loops, tex1light1.cpp:60-135)
			typedef InterpolationProcessorVec4<1> IP1;
			typedef typename PROG::I IP;
			const float rStep = IP1::reciprocalStepCount(stepCount);        // one divide per span, as in calcStep (tex1light1.cpp:93-107)
#pragma unroll
			for(int k = 0; k < NV; k++)
			{
				const float4 a = __ldg((const float4*)(v + k)), b = __ldg((const float4*)(v + NV + k)), c = __ldg((const float4*)(v + 2 * NV + k));
				const F4 v0 = f4(a.x, a.y, a.z, a.w), v1 = f4(b.x, b.y, b.z, b.w), v2 = f4(c.x, c.y, c.z, c.w);
				F4 vEnd;
				IP1::interpolateByContributes(&vStart[k], &v0, &v1, &v2, cl[0], cl[1], cl[2]);
				IP1::interpolateByContributes(&vEnd, &v0, &v1, &v2, cr[0], cr[1], cr[2]);
				IP1::calcStepR(&vStep[k], &vStart[k], &vEnd, rStep);
			}
			if(skip > 0) IP::stepForward(vStart, vStep, skip);               // interp.cpp:74-79
#pragma unroll 1
			for(int k = x1; k < x; k++) IP::stepForward(vStart, vStep, 1);    // interp.cpp:88, one rounded add per pixel
			IP::correctInterpolation(frag, vStart, inv);
		}
		FragmentProcessorOutput out;
		out.discarded = false; out.wrote = false; out.blendable = false; out.bgra = 0;
		PROG::F::process(frag, out, P);                                      // fragthrd.cpp:231
		if(P.cap && x < P.capW && y < P.capH) atomicAdd(&P.cap[(size_t)y * P.capW + x], 1u);
		if(out.wrote && y < P.colour.height && x < P.colour.width && Q.winner[(size_t)y * P.vpW + x] == i + 1)
		{
			// FBOBridge::write / write4 without ALPHABLEND: a plain store (fragthrd.cpp:54-82); later survivors of the pixel overwrite
			uint8_t* row = P.colour.ptr + (size_t)(P.colour.topDown ? P.colour.height - 1 - y : y) * P.colour.scanline;
			*(uint32_t*)(row + (size_t)x * 4) = out.bgra;
		}
	}
}
