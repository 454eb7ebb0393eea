// ps3d_cuda.cu — include/ps3d.h implemented on one B200 (sm_100a). The product library.
//
// Host side = the reference's resource tables (pipeline.cpp, tex.cpp, prog.cpp, vao.cpp, vbo.cpp) with storage in
// device memory, plus the per-draw enqueue of the kernels in kernels.cuh on the pipe's CUDA stream (the reference's
// ring queues and worker threads, rinque.h / fragthrd.cpp, have no counterpart: CUDA streams order the draws).
// There is NO CPU rendering path in this file: every ps3d_draw_vao either launches the kernels or returns an error.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <dlfcn.h>
#include "ps3d.h"
#include "kernels.cuh"
#include "kernels_span.cuh"
#include "x86_approx.h"

// the few NCCL types the run-time binding below needs (stable NCCL 2 ABI: a 128-byte unique id by value, an opaque communicator)
typedef struct { char internal[128]; } Ps3dNcclUniqueId;
typedef struct ncclComm* Ps3dNcclComm;

namespace
{

struct Texture
{
	int width, height, scanline, elemLen, wrap, nLayers, filter;
	uint8_t* layer[6];
};
// ready    : recorded on the copy stream behind the last asynchronous write (ps3d_vbo_update_async / ps3d_vbo_device_written);
//            every draw that reads the VBO makes the pipe's stream wait for it
// lastRead : recorded on the pipe's stream behind the last geometry kernel that read the VBO; an asynchronous write waits for it
struct Vbo { size_t unitBytes, unitCount; uint8_t* data; bool alive; cudaEvent_t ready, lastRead; bool readyValid, readValid; };
struct Vao { bool alive; int vbo[PS3D_MAX_VBOS]; };
struct Proc { bool alive; int kind, functor; };
struct Prog { int vp, ip, fp; };

typedef void (*LaunchGeom)(const DrawParams&, cudaStream_t);
typedef void (*LaunchTile)(const DrawParams&, const uint32_t*, const uint32_t*, cudaStream_t);
typedef void (*LaunchShade)(const DrawParams&, const SurvivorStream&, cudaStream_t);
typedef void (*LaunchShade2)(const DrawParams&, const SurvivorStream2&, bool, cudaStream_t);
typedef void (*LaunchGeomMulti)(const BatchView&, unsigned, cudaStream_t);
typedef void (*LaunchShadeMulti)(const DrawParams&, const SurvivorStream2&, const BatchView&, bool, cudaStream_t);
typedef void (*LaunchMarkVary)(const DrawParams&, const BatchView*, unsigned, cudaStream_t);

struct ProgEntry
{
	int fnV, fnI, fnF;
	int nv;
	uint32_t slots;
	uint64_t uniforms;
	int ntex;
	int texSlot[PS_MAX_BOUND_TEX];
	bool mayDiscard, usesWrite4;
	bool noop;                  // the fragment functor does nothing (FP_Null: depth-only passes)
	LaunchGeom geom;
	LaunchTile tileImmediate, tileOrdered;
	LaunchShade shade;
	LaunchGeom geomSpan;        // span path (kernels_span.cuh)
	LaunchShade2 shadeSpan;
	LaunchGeomMulti geomSpanMulti;   // batches of small draws
	LaunchShadeMulti shadeSpanMulti;
	LaunchMarkVary markVary;         // chain marks of long spans (the varyings' half)
};

// PS3D_TILE_PATH=immediate|ordered|split forces one tile path for every draw (A/B checks); default: chosen per draw
int tilePathForced()
{
	static int v = -2;
	if(v == -2)
	{
		const char* e = getenv("PS3D_TILE_PATH");
		v = !e ? -1 : (!strcmp(e, "immediate") ? 0 : (!strcmp(e, "ordered") ? 1 : (!strcmp(e, "split") ? 2 : -1)));
	}
	return v;
}

// PS3D_BINNING=radix forces the stable radix-sort binning (the fallback for very long tile lists) for every draw
bool radixBinningForced() { static int on = -1; if(on < 0) { const char* e = getenv("PS3D_BINNING"); on = (e && !strcmp(e, "radix")) ? 1 : 0; } return on == 1; }

// PS3D_GEOM_STAGE=0 turns the shared-memory staging of the vertex streams off (A/B checks)
bool geomStagingOn() { static int on = -1; if(on < 0) { const char* e = getenv("PS3D_GEOM_STAGE"); on = (e && e[0] == '0') ? 0 : 1; } return on == 1; }

// function-local launch state (attributes set, occupancy) is kept per device ordinal
#define PS_MAX_DEVICES 64
int currentDevice() { int d = 0; cudaGetDevice(&d); return d >= 0 && d < PS_MAX_DEVICES ? d : 0; }

template<class PROG> void launchGeom(const DrawParams& P, cudaStream_t s)
{
	const unsigned blocks = (P.ntris + PS_GEOM_THREADS - 1) / PS_GEOM_THREADS;
	// bytes of the block's vertex range of the staged (position) slot (bulk copies need 16-byte aligned sources: the
	// block's first element sits at a multiple of 384 * stride bytes from the 256-byte aligned VBO base)
	size_t bytes = 0;
	bool aligned = true;
	for(int i = 0; i < 16; i++)
		if(((PROG::V::SLOTS & PS_GEOM_STAGE_SLOTS) >> i) & 1)
		{
			bytes += ((size_t)PS_GEOM_THREADS * 3 * P.stride[i] + 127) & ~(size_t)127;
			aligned = aligned && 0 == ((uintptr_t)P.slot[i] & 15);
		}
	const bool staged = geomStagingOn() && aligned && bytes > 0 && bytes <= 96 * 1024;
	// with a sort-first band: the position half appends the band's survivors to a list, the other half runs over the list
	const bool split = P.workList && (P.band0 > 0 || P.band1 < P.vpH);
	if(staged)
	{
		// cudaFuncSetAttribute is per device: one flag per device ordinal (a process may own pipes on several GPUs)
		static bool attrSet[PS_MAX_DEVICES] = { false };
		const int dev = currentDevice();
		if(!attrSet[dev])
		{
			cudaFuncSetAttribute(geom_setup_kernel<PROG, true, PS_GEOM_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024));
			cudaFuncSetAttribute(geom_setup_kernel<PROG, true, PS_GEOM_APPEND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024));
			attrSet[dev] = true;
		}
		if(split) geom_setup_kernel<PROG, true, PS_GEOM_APPEND><<<blocks, PS_GEOM_THREADS, bytes, s>>>(P);
		else geom_setup_kernel<PROG, true, PS_GEOM_FUSED><<<blocks, PS_GEOM_THREADS, bytes, s>>>(P);
	}
	else
	{
		if(split) geom_setup_kernel<PROG, false, PS_GEOM_APPEND><<<blocks, PS_GEOM_THREADS, 0, s>>>(P);
		else geom_setup_kernel<PROG, false, PS_GEOM_FUSED><<<blocks, PS_GEOM_THREADS, 0, s>>>(P);
	}
	if(split) geom_setup_kernel<PROG, false, PS_GEOM_LIST><<<blocks, PS_GEOM_THREADS, 0, s>>>(P);
}
bool geomSplitWanted(int bandRows, int vpH)
{
	static int mode = -2;   // -1 = by band size, 0 = never, 1 = always
	if(-2 == mode) { const char* e = getenv("PS3D_GEOM_SPLIT"); mode = !e ? -1 : (e[0] == '0' ? 0 : 1); }
	return mode < 0 ? (long long)bandRows * 6 <= vpH : 1 == mode;
}
// PS3D_RASTER_PARTS=1|2|4 forces the number of row groups a tile is cut into (A/B checks, tests)
int rasterPartsForced() { static int v = -2; if(-2 == v) { const char* e = getenv("PS3D_RASTER_PARTS"); v = e ? atoi(e) : 0; if(v != 1 && v != 2 && v != 4 && v != 8) v = 0; } return v; }
unsigned tileBlocks(const DrawParams& P) { return ((unsigned)(P.tilesX * P.tilesY) + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK; }
template<class PROG> void launchTileImmediate(const DrawParams& P, const uint32_t* tileStart, const uint32_t* sortedTris, cudaStream_t s)
{
	tile_raster_shade_immediate_kernel<PROG><<<tileBlocks(P), 32 * PS_WARPS_PER_BLOCK, 0, s>>>(P, tileStart, sortedTris);
}
template<class PROG> void launchTileOrdered(const DrawParams& P, const uint32_t* tileStart, const uint32_t* sortedTris, cudaStream_t s)
{
	tile_raster_shade_ordered_kernel<PROG><<<tileBlocks(P), 32 * PS_WARPS_PER_BLOCK, 0, s>>>(P, tileStart, sortedTris);
}
template<class PROG> void launchShade(const DrawParams& P, const SurvivorStream& Q, cudaStream_t s)
{
	// a flat grid-stride loop over the survivor stream: exactly one resident wave
	static int perSMs[PS_MAX_DEVICES] = { 0 }, smss[PS_MAX_DEVICES] = { 0 };
	const int dev = currentDevice();
	int& perSM = perSMs[dev];
	int& sms = smss[dev];
	if(0 == perSM)
	{
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_kernel<PROG>, 128, 0) != cudaSuccess || perSM <= 0) perSM = 4;
	}
	shade_kernel<PROG><<<sms * perSM, 128, 0, s>>>(P, Q);
}
// ---- span path launchers ----------------------------------------------------------------------------------------------
template<class PROG, int STAGED> bool launchGeomSpanStaged(const DrawParams& P, unsigned blocks, cudaStream_t s)
{
	// bytes of the block's vertex range of the staged slots (bulk copies need 16-byte aligned sources: the block's first element
	// sits at a multiple of 384 * stride bytes from the 256-byte aligned VBO base)
	size_t bytes = 0;
	bool aligned = true;
	for(int i = 0; i < 16; i++)
		if((stageMask<STAGED>(PROG::V::SLOTS) >> i) & 1)
		{
			bytes += ((size_t)PS_GEOM_THREADS * 3 * P.stride[i] + 127) & ~(size_t)127;
			aligned = aligned && 0 == ((uintptr_t)P.slot[i] & 15) && 0 == ((PS_GEOM_THREADS * 3 * P.stride[i]) & 15);
		}
	if(!aligned || 0 == bytes || bytes > 64 * 1024) return false;
	static bool attrSet[PS_MAX_DEVICES] = { false };
	const int dev = currentDevice();
	if(!attrSet[dev])
	{
		cudaFuncSetAttribute(geom_span_kernel<PROG, STAGED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(64 * 1024));
		attrSet[dev] = true;
	}
	geom_span_kernel<PROG, STAGED, false><<<blocks, PS_GEOM_THREADS, bytes, s>>>(P);
	return true;
}
template<class PROG> void launchGeomSpan(const DrawParams& P, cudaStream_t s)
{
	const unsigned blocks = (P.ntris + PS_GEOM_THREADS - 1) / PS_GEOM_THREADS;
	// sort-first band: triangles with no row in the band leave after a rows-only look at their vertices (a pre-cull kernel fills a
	// list, the geometry kernel runs over it). PS3D_GEOM_BAND=none: every rank takes every triangle through the whole position
	// half (A/B runs)
	static int bandMode = -1;
	if(bandMode < 0) { const char* e = getenv("PS3D_GEOM_BAND"); bandMode = (e && !strcmp(e, "none")) ? 0 : 1; }
	if(bandMode && P.workList && (P.band0 > 0 || P.band1 < P.vpH))
	{
		geom_precull_kernel<PROG><<<(P.ntris + PS_PRECULL_THREADS - 1) / PS_PRECULL_THREADS, PS_PRECULL_THREADS, 0, s>>>(P);
		geom_span_kernel<PROG, 0, true><<<blocks, PS_GEOM_THREADS, 0, s>>>(P);
		return;
	}
	if(P.batchTrisPerBlock)
	{
		geom_span_kernel<PROG, 0, false><<<(P.ntris + P.batchTrisPerBlock - 1) / P.batchTrisPerBlock, PS_GEOM_THREADS, 0, s>>>(P);
		return;
	}
	// PS3D_GEOM_STAGE=0: nothing staged (A/B runs); default: the position slot staged in shared memory by one TMA bulk copy per block.
	// (Every slot staged — STAGED = 2, 28 KB per block for DEF03, phase B off global memory — measured 0.194 -> 0.215 ms on C2: dropped.)
	if(geomStagingOn() && launchGeomSpanStaged<PROG, 1>(P, blocks, s)) return;
	geom_span_kernel<PROG, 0, false><<<blocks, PS_GEOM_THREADS, 0, s>>>(P);
}
template<class PROG> void launchShadeSpan(const DrawParams& P, const SurvivorStream2& Q, bool marks, cudaStream_t s)
{
	if(Q.colour)
	{
		// a draw that blends: colours left in the stream for shade_resolve_kernel (the variant with chain marks serves both cases)
		if constexpr(PROG::F::USES_WRITE4)
		{
			static int perSMo[PS_MAX_DEVICES] = { 0 }, smso[PS_MAX_DEVICES] = { 0 };
			const int dev = currentDevice();
			if(0 == perSMo[dev])
			{
				cudaDeviceGetAttribute(&smso[dev], cudaDevAttrMultiProcessorCount, dev);
				if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMo[dev], shade_span_kernel<PROG, 7, false, true, true>, PS_SHADE_THREADS, 0) != cudaSuccess || perSMo[dev] <= 0) perSMo[dev] = 4;
			}
			const BatchView none = { nullptr, nullptr, nullptr, 0 };
			shade_span_kernel<PROG, 7, false, true, true><<<smso[dev] * perSMo[dev], PS_SHADE_THREADS, 0, s>>>(P, Q, none);
		}
		return;
	}
	if(marks && PROG::NV > 0)
	{
		// long spans with chain marks (SpanStreams::markV): the variant that starts from them
		static int perSMm[PS_MAX_DEVICES] = { 0 }, smsm[PS_MAX_DEVICES] = { 0 };
		const int dev = currentDevice();
		if(0 == perSMm[dev])
		{
			cudaDeviceGetAttribute(&smsm[dev], cudaDevAttrMultiProcessorCount, dev);
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMm[dev], shade_span_kernel<PROG, 7, false, true>, PS_SHADE_THREADS, 0) != cudaSuccess || perSMm[dev] <= 0) perSMm[dev] = 4;
		}
		const BatchView none = { nullptr, nullptr, nullptr, 0 };
		shade_span_kernel<PROG, 7, false, true><<<smsm[dev] * perSMm[dev], PS_SHADE_THREADS, 0, s>>>(P, Q, none);
		return;
	}
	// a flat grid-stride loop over the survivor stream: exactly one resident wave. PS3D_SHADE_MINB=5|7|8: the variant compiled for
	// that many blocks per SM (A/B switch)
	static int perSMs[PS_MAX_DEVICES][3] = { { 0 } }, smss[PS_MAX_DEVICES] = { 0 };
	static int minb = -1;                              // 0: 5 blocks (100 registers), 1: 7 blocks (72, default), 2: 8 blocks (64)
	if(minb < 0) { const char* e = getenv("PS3D_SHADE_MINB"); const int v = e ? atoi(e) : 7; minb = v <= 5 ? 0 : (v >= 8 ? 2 : 1); }
	const int dev = currentDevice();
	int& perSM = perSMs[dev][minb];
	int& sms = smss[dev];
	if(0 == perSM)
	{
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		const cudaError_t e = 2 == minb ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_span_kernel<PROG, 8>, PS_SHADE_THREADS, 0)
		                    : (1 == minb ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_span_kernel<PROG, 7>, PS_SHADE_THREADS, 0)
		                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_span_kernel<PROG, 5>, PS_SHADE_THREADS, 0));
		if(e != cudaSuccess || perSM <= 0) perSM = 4;
	}
	const BatchView none = { nullptr, nullptr, nullptr, 0 };
	if(2 == minb) shade_span_kernel<PROG, 8><<<sms * perSM, PS_SHADE_THREADS, 0, s>>>(P, Q, none);
	else if(1 == minb) shade_span_kernel<PROG, 7><<<sms * perSM, PS_SHADE_THREADS, 0, s>>>(P, Q, none);
	else shade_span_kernel<PROG, 5><<<sms * perSM, PS_SHADE_THREADS, 0, s>>>(P, Q, none);
}
// ---- a batch of small draws (kernels_span.cuh: BatchView): one geometry and one shade launch per programme present ------------
template<class PROG> void launchGeomSpanMulti(const BatchView& B, unsigned blocks, cudaStream_t s)
{
	geom_span_multi_kernel<PROG><<<blocks, PS_GEOM_THREADS, 0, s>>>(B);
}
template<class PROG> void launchShadeSpanMulti(const DrawParams& P, const SurvivorStream2& Q, const BatchView& B, bool marks, cudaStream_t s)
{
	static int perSMs[PS_MAX_DEVICES][2] = { { 0 } }, smss[PS_MAX_DEVICES] = { 0 };
	const int dev = currentDevice();
	const int m = (marks && PROG::NV > 0) ? 1 : 0;
	int& perSM = perSMs[dev][m];
	int& sms = smss[dev];
	if(0 == perSM)
	{
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		const cudaError_t e = m ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_span_kernel<PROG, 7, true, true>, PS_SHADE_THREADS, 0)
		                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shade_span_kernel<PROG, 7, true, false>, PS_SHADE_THREADS, 0);
		if(e != cudaSuccess || perSM <= 0) perSM = 4;
	}
	if(m) shade_span_kernel<PROG, 7, true, true><<<sms * perSM, PS_SHADE_THREADS, 0, s>>>(P, Q, B);
	else shade_span_kernel<PROG, 7, true, false><<<sms * perSM, PS_SHADE_THREADS, 0, s>>>(P, Q, B);
}
// chain marks of the varyings of long spans (kernels_span.cuh: span_mark_vary_kernel); B: a batch's programme group, or NULL
template<class PROG> void launchMarkVary(const DrawParams& P, const BatchView* B, unsigned blocks, cudaStream_t s)
{
	if(0 == PROG::NV) return;
	const BatchView none = { nullptr, nullptr, nullptr, 0 };
	if(B) span_mark_vary_kernel<PROG, true><<<blocks, PS_MARK_THREADS, 0, s>>>(P, *B);
	else span_mark_vary_kernel<PROG, false><<<blocks, PS_MARK_THREADS, 0, s>>>(P, none);
}

template<class F> constexpr auto fragmentNoop(int) -> decltype(F::NOOP) { return F::NOOP; }
template<class F> constexpr bool fragmentNoop(...) { return false; }
template<class PROG> ProgEntry makeEntry(int fnV, int fnI, int fnF)
{
	ProgEntry e;
	e.fnV = fnV; e.fnI = fnI; e.fnF = fnF;
	e.nv = PROG::NV;
	e.slots = PROG::V::SLOTS;
	e.uniforms = PROG::V::UNIFORMS | PROG::F::UNIFORMS;
	e.ntex = PROG::F::NTEX;
	for(int i = 0; i < PS_MAX_BOUND_TEX; i++) e.texSlot[i] = i < e.ntex ? PROG::F::texSlot(i) : -1;
	e.mayDiscard = PROG::F::MAY_DISCARD;
	e.usesWrite4 = PROG::F::USES_WRITE4;
	e.noop = fragmentNoop<typename PROG::F>(0);
	e.geom = launchGeom<PROG>;
	e.tileImmediate = launchTileImmediate<PROG>;
	e.tileOrdered = launchTileOrdered<PROG>;
	e.shade = launchShade<PROG>;
	e.geomSpan = launchGeomSpan<PROG>;
	e.shadeSpan = launchShadeSpan<PROG>;
	e.geomSpanMulti = launchGeomSpanMulti<PROG>;
	e.shadeSpanMulti = launchShadeSpanMulti<PROG>;
	e.markVary = launchMarkVary<PROG>;
	return e;
}

const std::vector<ProgEntry>& programmeTable()
{
	static std::vector<ProgEntry> t;
	if(t.empty())
	{
		t.push_back(makeEntry<ProgDEF01>(PS3D_FN_DEF01, PS3D_FN_DEF01, PS3D_FN_DEF01));
		t.push_back(makeEntry<ProgDEF02>(PS3D_FN_DEF02, PS3D_FN_DEF02, PS3D_FN_DEF02));
		t.push_back(makeEntry<ProgDEF03>(PS3D_FN_DEF03, PS3D_FN_DEF03, PS3D_FN_DEF03));
		t.push_back(makeEntry<ProgDEF04>(PS3D_FN_DEF04, PS3D_FN_DEF04, PS3D_FN_DEF04));
		t.push_back(makeEntry<ProgDEF05>(PS3D_FN_DEF05, PS3D_FN_DEF05, PS3D_FN_DEF05));
		t.push_back(makeEntry<ProgFLATID>(PS3D_FN_FLATID, PS3D_FN_FLATID, PS3D_FN_FLATID));
		t.push_back(makeEntry<ProgTEXPROBE>(PS3D_FN_TEXPROBE, PS3D_FN_TEXPROBE, PS3D_FN_TEXPROBE));
		// demo 1 (src/test/testproc.cpp) and demo 2 (src/test2/testproc.cpp): the triples their scene objects create
		t.push_back(makeEntry<ProgEarth>(PS3D_FN_PLANET, PS3D_FN_PLANET, PS3D_FN_PLANET));
		t.push_back(makeEntry<ProgSatellite>(PS3D_FN_PLANET, PS3D_FN_PLANET, PS3D_FN_SATELLITE));
		t.push_back(makeEntry<ProgCloud>(PS3D_FN_CLOUD, PS3D_FN_CLOUD, PS3D_FN_CLOUD));
		t.push_back(makeEntry<ProgCloudShadow>(PS3D_FN_CLOUDSHADOW, PS3D_FN_CLOUDSHADOW, PS3D_FN_CLOUDSHADOW));
		t.push_back(makeEntry<ProgPositionOnly>(PS3D_FN_POSITIONONLY, PS3D_FN_POSITIONONLY, PS3D_FN_POSITIONONLY));
		t.push_back(makeEntry<ProgSingleColour>(PS3D_FN_SINGLECOLOUR, PS3D_FN_SINGLECOLOUR, PS3D_FN_SINGLECOLOUR));
		t.push_back(makeEntry<ProgDiffuseOnly>(PS3D_FN_DIFFUSEONLY, PS3D_FN_DIFFUSEONLY, PS3D_FN_DIFFUSEONLY));
		t.push_back(makeEntry<ProgShadow2>(PS3D_FN_SHADOW2, PS3D_FN_POSITIONONLY, PS3D_FN_SHADOW2));
	}
	return t;
}

bool functorKnown(int kind, int fn)
{
	for(const ProgEntry& e : programmeTable())
		if((kind == PS3D_PROC_VERTEX && e.fnV == fn) || (kind == PS3D_PROC_INTERPOLATION && e.fnI == fn) || (kind == PS3D_PROC_FRAGMENT && e.fnF == fn))
			return true;
	return false;
}

// set while a frame is being captured (ps3d_graph_begin .. ps3d_graph_end): nothing may be allocated or freed then
thread_local bool g_capturing = false;

template<typename T> struct DevBuf
{
	T* p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t n)
	{
		if(n <= cap) return cudaSuccess;
		if(g_capturing) return cudaErrorStreamCaptureUnsupported;   // run the frame once normally before capturing it
		if(p) cudaFree(p);
		p = nullptr;
		size_t want = n + n / 4 + 1024;
		cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
		cap = e == cudaSuccess ? want : 0;
		return e;
	}
	void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace

struct ps3d_pipe
{
	int device;
	int smCount;
	cudaStream_t stream;
	cudaStream_t copyStream;    // asynchronous uploads (ps3d_vbo_update_async)
	cudaStream_t readStream;    // asynchronous read-backs (ps3d_read_colour_async): its own stream, so that a read-back waiting for its frame does not hold up the next frame's uploads
	cudaEvent_t frameDone, readDone[2];
	// sort-first exchange (ps3d_comm_*): one communicator per stream that carries collectives, so that the composite of
	// frame i (pipe stream) and the upload all-gathers of frame i + 1 (gather stream) do not serialise on each other
	Ps3dNcclComm commFrame, commUpload;
	cudaStream_t gatherStream;
	int commRank, commWorld;
	bool readValid[2];
	int width, height;
	int vpW, vpH;
	int behavior;
	int band0, band1;
	uint8_t* display[2];
	int back;
	uint8_t* defaultDepth;
	int depthScanline;
	int depthTex;
	std::vector<Texture*> textures;
	std::vector<Vbo> vbos;
	std::vector<Vao> vaos;
	std::vector<Proc> procs;
	std::vector<Prog> progs;
	int curProg;
	std::vector<uint8_t> uniforms[PS3D_MAX_UNIFORMS];
	bool uniformSet[PS3D_MAX_UNIFORMS];
	// per-draw scratch
	DevBuf<TriHeader> hdr;
	DevBuf<F4> vary;
	DevBuf<uint32_t> triCount, triOffset, triRect, scanSums, keysA, valsA, keysB, valsB, tileCount, tileStart, tileFill, tileOrder, sortCounts, workList;
	DevBuf<uint32_t> svTri, svMisc, svWinner;   // survivor stream of the draw in flight (split path)
	DevBuf<int> svLeft, svRight;
	DevBuf<float> svInv;
	uint32_t* svCountDev;
	// span path (kernels_span.cuh): span records, per-triangle record index, fixed-capacity tile lists, 12-byte survivors
	DevBuf<SpanRec> spRec;
	DevBuf<TriSpan> spTri;
	DevBuf<uint32_t> tlFill, tlLen, tlIds;
	DevBuf<uint32_t> sv2Span, sv2XY;
	DevBuf<float> sv2Inv;
	DevBuf<uint32_t> sv2Next, sv2Chain, sv2Colour;   // draws that blend (device_types.cuh: SurvivorStream2)
	uint32_t* spanCountDev;
	size_t spanHigh, listHigh;     // high-water marks of earlier draws: the next speculation
	// chain marks of long spans (device_types.cuh: SpanStreams): kept once a draw has asked for some
	DevBuf<uint32_t> spMarkAt, spLongList;
	DevBuf<MarkZ> spMarkZ;
	DevBuf<F4> spMarkV;
	unsigned long long* longCountDev;
	uint32_t* longLatchedDev;
	uint32_t* bigListDev;          // TileLists::bigList: [0] the counter, from [4] the entries

	size_t markHigh;
	size_t marksMin;               // marks are kept for draws that ask for at least this many (PS3D_MARKS_MIN; two more launches per draw)
	std::vector<uint8_t> vaoLegacy; // VAOs whose last draw needed the first path (a tile list too long for the shared-memory sort)
	// batches of small draws (kernels_span.cuh: BatchView): `batch` collects draws until something other than a like draw arrives,
	// `flight` is the batch whose launches are on the stream and whose verdict settle() has yet to read
	struct BatchDraw { DrawParams P; const ProgEntry* pe; int vao; uint32_t firstBlock, nBlocks, trisPerBlock; };
	struct Batch { std::vector<BatchDraw> draws; uint32_t blocks; void clear() { draws.clear(); blocks = 0; } };
	Batch batch, flight;
	DevBuf<DrawParams> batchItems;
	DevBuf<uint32_t> batchBlockDraw, batchBlockList;
	uint64_t batchesLaunched, drawsBatched;
	// sort-first composite over peer memory (ps3d_peer_*): rank 0's display targets and flag block mapped into every rank
	struct Peer
	{
		bool active; int rank, world;
		uint8_t* display0[2];       // rank 0's display targets (ranks != 0: cudaIpc mappings; rank 0: its own)
		PeerFlags* flags0;          // rank 0's flag block (ranks != 0: mapping)
		PeerFlags* flagsOwn;        // every rank allocates one (only rank 0's is used) so that the export blob has one shape
		PeerCounters* ctr;
		bool needTake;              // no colour write since the last composite: the next one takes (rank 0: hands out) the target first
	} peer;
	// captured frames (ps3d_graph_*)
	struct Graph { cudaGraphExec_t exec; uint64_t launches; uint64_t draws, tris; std::vector<int> vaos; int back, backAfter; bool alive; uint64_t scratch; };
	std::vector<Graph> graphs;
	bool capturing, graphLaunched;
	int capBack;
	uint64_t capLaunches0, capDraws0, capTris0;
	std::vector<int> capVaos;

	uint32_t* totalDev;
	DeviceStats* statsDev;      // PS_STATS_COPIES replicas
	// asynchronous draws: the tail of a draw (binning, raster, shade) is enqueued behind the tile scan with SPECULATED
	// buffer sizes and guarded by *poisonDev; the scan's report is read at the next API call (settle)
	uint32_t* poisonDev;
	DrawReport* report;         // mapped pinned host memory, written by tile_scan_kernel
	DrawReport* reportDev;      // the device's alias of it
	cudaEvent_t scanEvent;
	bool speculate;
	size_t pairHigh, survivorHigh;   // high-water marks of earlier draws: the next speculation
	struct Pending { bool valid; DrawParams P; const ProgEntry* pe; int path; bool radix; bool span; bool tailLaunched; int vao; bool multi; bool blend; } pending;
	ps3d_stats stats;
	uint32_t* capDev;
	int capW, capH;
	uint32_t* rcpDev;
	uint32_t* rsqrtDev;
	ApproxTables approx;
	uint64_t launches;
	// per-kernel-class event timing (ps3d_profile_*)
	bool profiling;
	struct Span { cudaEvent_t a, b; int cls; };
	std::vector<Span> spans;
	std::vector<cudaEvent_t> eventPool;
	uint64_t profLaunches[4];
	uint64_t profPairs, profSurvivorBound;
	std::string err;
};

enum { CLS_GEOM = 0, CLS_BIN = 1, CLS_TILE = 2, CLS_SHADE = 3 };

static cudaEvent_t takeEvent(ps3d_pipe* p)
{
	cudaEvent_t e;
	if(!p->eventPool.empty()) { e = p->eventPool.back(); p->eventPool.pop_back(); return e; }
	cudaEventCreate(&e);
	return e;
}
struct ProfScope
{
	ps3d_pipe* p; int cls; cudaEvent_t a; uint64_t launches0;
	ProfScope(ps3d_pipe* p_, int cls_) : p(p_), cls(cls_), a(nullptr), launches0(p_->launches)
	{
		if(p->profiling) { a = takeEvent(p); cudaEventRecord(a, p->stream); }
	}
	~ProfScope()
	{
		if(!a) return;
		ps3d_pipe::Span s; s.a = a; s.b = takeEvent(p); s.cls = cls;
		cudaEventRecord(s.b, p->stream);
		p->spans.push_back(s);
		p->profLaunches[cls] += p->launches - launches0;
	}
};

#define CK(p, call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { (p)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PS3D_ERR_DEVICE; } } while(0)

static int settle(ps3d_pipe* p);
static int peerFirstWrite(ps3d_pipe* p);
static int launchSpanTail(ps3d_pipe* p, const DrawParams& P, const ProgEntry* pe, bool multi, bool blend);
static int enqueueSpan(ps3d_pipe* p, DrawParams P, const ProgEntry* pe, int vao, size_t spans, size_t longest, size_t survivors, bool multi = false, bool blend = false);
static int flushBatch(ps3d_pipe* p);
#define PS_BIG_LIST 4096u            // very large tile rectangles a draw may hand to the grid-wide append (more: their blocks append them)
#define PS_BATCH_DRAW_TRIS 16384u     // a draw with more triangles than this fills the GPU by itself
#define PS_BATCH_MAX_DRAWS 1024u
#define PS_BATCH_MAX_BLOCKS 8192u       // (every block is PS_GEOM_THREADS triangle ids: 64 B of header each)

// a small draw spreads over up to 2048 blocks (a block's rows and whole-rectangle tile appends are its own threads' work: twenty
// full-screen triangles in one block are half a millisecond of appends, in twenty blocks 25 us: kernels_span.cuh)
static uint32_t batchTrisPerBlock(size_t ntris) { return (uint32_t)std::min<size_t>(std::max<size_t>((ntris + 2047) / 2048, 1), PS_GEOM_THREADS); }
static uint32_t batchBlocks(size_t ntris) { const uint32_t per = batchTrisPerBlock(ntris); return (uint32_t)((ntris + per - 1) / per); }

static bool marksOn();
static bool spanNoShade(const ps3d_pipe* p, const DrawParams& P, const ProgEntry* pe, bool multi);
static int enqueueLegacy(ps3d_pipe* p, DrawParams P, const ProgEntry* pe, int path, int vao);

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------
// The sort-first exchange steps (band composite to rank 0, all-gather of the sharded vertex upload) are issued from inside
// the library on the pipe's own streams. libnccl is not a link-time dependency: the entry points are looked up in the
// libnccl.so.2 the process already has (torch.distributed's, when the host side is Python) or can load; the few types
// are declared here (stable NCCL 2 ABI: a 128-byte unique id by value, opaque communicator, int enums).
struct NcclApi
{
	void* lib;
	int (*GetUniqueId)(Ps3dNcclUniqueId*);
	int (*CommInitRank)(Ps3dNcclComm*, int, Ps3dNcclUniqueId, int);
	int (*CommDestroy)(Ps3dNcclComm);
	int (*GroupStart)();
	int (*GroupEnd)();
	int (*Send)(const void*, size_t, int, int, Ps3dNcclComm, cudaStream_t);
	int (*Recv)(void*, size_t, int, int, Ps3dNcclComm, cudaStream_t);
	int (*AllGather)(const void*, void*, size_t, int, Ps3dNcclComm, cudaStream_t);
	const char* (*GetErrorString)(int);
};
enum { PS_NCCL_UINT8 = 1 };   // ncclUint8 / ncclChar's unsigned sibling in nccl.h's ncclDataType_t
static const NcclApi* ncclApi()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		memset(&api, 0, sizeof(api));
		void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		if(!h) h = dlopen("libnccl.so.2", RTLD_NOW);
		if(!h) h = dlopen("libnccl.so", RTLD_NOW);
		if(!h) return;
		api.GetUniqueId = (int (*)(Ps3dNcclUniqueId*))dlsym(h, "ncclGetUniqueId");
		api.CommInitRank = (int (*)(Ps3dNcclComm*, int, Ps3dNcclUniqueId, int))dlsym(h, "ncclCommInitRank");
		api.CommDestroy = (int (*)(Ps3dNcclComm))dlsym(h, "ncclCommDestroy");
		api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
		api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
		api.Send = (int (*)(const void*, size_t, int, int, Ps3dNcclComm, cudaStream_t))dlsym(h, "ncclSend");
		api.Recv = (int (*)(void*, size_t, int, int, Ps3dNcclComm, cudaStream_t))dlsym(h, "ncclRecv");
		api.AllGather = (int (*)(const void*, void*, size_t, int, Ps3dNcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
		api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
		if(api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.AllGather) api.lib = h;
	});
	return api.lib ? &api : nullptr;
}
#define NK(p, call) do { int r_ = (call); if(r_ != 0) { const NcclApi* a_ = ncclApi(); (p)->err = std::string(#call) + ": " + ((a_ && a_->GetErrorString) ? a_->GetErrorString(r_) : "nccl error"); return PS3D_ERR_DEVICE; } } while(0)
// (collected small draws are launched first: whatever calls this is about to touch the stream, the targets or device memory)
#define SETTLE(p) do { int rc_ = flushBatch(p); if(rc_) return rc_; rc_ = settle(p); if(rc_) return rc_; } while(0)

static int fail(ps3d_pipe* p, int code, const char* msg) { p->err = msg; return code; }
static void freeVbo(Vbo& v)
{
	if(v.data) cudaFree(v.data);
	if(v.ready) cudaEventDestroy(v.ready);
	if(v.lastRead) cudaEventDestroy(v.lastRead);
	v.data = nullptr; v.ready = v.lastRead = nullptr; v.readyValid = v.readValid = false; v.alive = false;
}
static int vboEvents(ps3d_pipe* p, Vbo& v)
{
	if(!v.ready) CK(p, cudaEventCreateWithFlags(&v.ready, cudaEventDisableTiming));
	if(!v.lastRead) CK(p, cudaEventCreateWithFlags(&v.lastRead, cudaEventDisableTiming));
	return PS3D_OK;
}

// PS3D_TRACE=1 prints every C-ABI entry to stderr (debug aid)
static bool traceOn() { static int on = -1; if(on < 0) { const char* e = getenv("PS3D_TRACE"); on = (e && e[0] == '1') ? 1 : 0; } return on == 1; }
#define TRACE() do { if(traceOn()) { fprintf(stderr, "[ps3d] %s\n", __func__); fflush(stderr); } } while(0)

static const Ps3dHostApprox& hostApprox()
{
	static Ps3dHostApprox a;
	static std::once_flag once;
	std::call_once(once, [] { ps3d_measure_x86_approx(&a); });
	return a;
}

static int exclusiveScan(ps3d_pipe* p, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* totalDev)
{
	const uint32_t nb = (n + PS_SCAN_BLOCK - 1) / PS_SCAN_BLOCK;
	CK(p, p->scanSums.ensure(nb + 1));
	scan_local_kernel<<<nb, PS_SCAN_THREADS, 0, p->stream>>>(in, out, p->scanSums.p, n);
	scan_sums_kernel<<<1, 1024, 0, p->stream>>>(p->scanSums.p, nb, totalDev);
	scan_add_kernel<<<nb, PS_SCAN_THREADS, 0, p->stream>>>(out, p->scanSums.p, n);
	p->launches += 3;
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

static TargetDesc depthTarget(ps3d_pipe* p)
{
	TargetDesc d;
	if(p->depthTex < 0)
	{
		d.ptr = p->defaultDepth; d.width = p->width; d.height = p->height; d.scanline = p->depthScanline; d.topDown = 0;
	}
	else
	{
		Texture* t = p->textures[p->depthTex];
		d.ptr = t->layer[0]; d.width = t->width; d.height = t->height; d.scanline = t->scanline; d.topDown = 0;
	}
	return d;
}

static int ensureSurvivors(ps3d_pipe* p, size_t cap, const DrawParams& P)
{
	CK(p, p->svTri.ensure(cap)); CK(p, p->svLeft.ensure(cap)); CK(p, p->svRight.ensure(cap)); CK(p, p->svInv.ensure(cap)); CK(p, p->svMisc.ensure(cap));
	CK(p, p->svWinner.ensure((size_t)P.vpW * P.vpH));
	return PS3D_OK;
}

// The draw behind its tile scan: binning by atomics + per-tile sort (or, radix == true, the stable radix sort for
// lists too long for shared memory), then the tile kernels of the chosen path. Every kernel here returns at once when
// the scan raised *poison. Buffers must already be large enough (speculated, or sized exactly by settle()).
static int launchTail(ps3d_pipe* p, const DrawParams& P, const ProgEntry* pe, int path, bool radix)
{
	const uint32_t ntiles = (uint32_t)(P.tilesX * P.tilesY);
	const uint32_t* sortedTris = nullptr;
	if(!radix)
	{
		// lists filled with atomics in arrival order, then every tile's list sorted by triangle id = submission order
		ProfScope ps(p, CLS_BIN);
		bin_fill_kernel<<<(P.ntris + 127) / 128, 128, 0, p->stream>>>(p->triCount.p, p->triRect.p, p->tileStart.p, p->tileFill.p, p->valsA.p, P.ntris, P.tilesX, p->poisonDev);
		tile_list_sort_kernel<<<(ntiles + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(p->tileStart.p, p->valsA.p, ntiles, p->poisonDev, p->tileOrder.p);
		p->launches += 2;
		CK(p, cudaGetLastError());
		sortedTris = p->valsA.p;
	}
	else
	{
		// some tile list is too long for the shared-memory sort: pairs emitted in triangle order + stable LSD radix sort by tile
		ProfScope ps(p, CLS_BIN);
		const uint32_t total = p->report->pairs;
		CK(p, p->triOffset.ensure(P.ntris));
		int rc = exclusiveScan(p, p->triCount.p, p->triOffset.p, P.ntris, p->totalDev);
		if(rc) return rc;
		CK(p, p->keysA.ensure(total)); CK(p, p->valsA.ensure(total)); CK(p, p->keysB.ensure(total)); CK(p, p->valsB.ensure(total));
		emit_pairs_kernel<<<(P.ntris + 127) / 128, 128, 0, p->stream>>>(p->triCount.p, p->triOffset.p, p->triRect.p, p->keysA.p, p->valsA.p, P.ntris, P.tilesX);
		p->launches++;
		CK(p, cudaGetLastError());
		int bits = 0;
		while((1u << bits) < ntiles) bits++;
		uint32_t *kIn = p->keysA.p, *vIn = p->valsA.p, *kOut = p->keysB.p, *vOut = p->valsB.p;
		const uint32_t nwarps = (total + PS_SORT_ITEMS_PER_WARP - 1) / PS_SORT_ITEMS_PER_WARP;
		CK(p, p->sortCounts.ensure((size_t)256 * nwarps));
		for(int shift = 0; shift < bits; shift += 8)
		{
			const unsigned blocks = (nwarps + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK;
			sort_hist_kernel<<<blocks, 128, 0, p->stream>>>(kIn, total, shift, p->sortCounts.p, nwarps);
			p->launches++;
			rc = exclusiveScan(p, p->sortCounts.p, p->sortCounts.p, 256 * nwarps, p->totalDev);
			if(rc) return rc;
			sort_scatter_kernel<<<blocks, 128, 0, p->stream>>>(kIn, vIn, kOut, vOut, total, shift, p->sortCounts.p, nwarps);
			p->launches++;
			CK(p, cudaGetLastError());
			uint32_t* t;
			t = kIn; kIn = kOut; kOut = t;
			t = vIn; vIn = vOut; vOut = t;
		}
		sortedTris = vIn;
	}
	{ const int rc = peerFirstWrite(p); if(rc) return rc; }
	if(0 == path) { ProfScope ps(p, CLS_TILE); pe->tileImmediate(P, p->tileStart.p, sortedTris, p->stream); p->launches++; }
	else if(1 == path) { ProfScope ps(p, CLS_TILE); pe->tileOrdered(P, p->tileStart.p, sortedTris, p->stream); p->launches++; }
	else
	{
		SurvivorStream Q;
		Q.tri = p->svTri.p; Q.left = p->svLeft.p; Q.right = p->svRight.p; Q.inv = p->svInv.p; Q.misc = p->svMisc.p;
		Q.count = p->svCountDev; Q.winner = p->svWinner.p; Q.capacity = (uint32_t)std::min<size_t>(p->svTri.cap, 0xfffffff0u);
		CK(p, cudaMemsetAsync(p->svCountDev, 0, 4, p->stream));
		{
			ProfScope ps(p, CLS_TILE);
			// fewer tiles in play than warp slots on the GPU: cut every tile into 2 or 4 row groups, one warp each
			const int bandRows = std::max(0, std::min(P.band1, P.vpH) - std::max(P.band0, 0));
			const long long tilesInPlay = (long long)P.tilesX * ((bandRows + PS_TILE - 1) / PS_TILE);
			const long long warpSlots = (long long)p->smCount * 28;      // resident warps of this kernel (shared memory / registers)
			int parts = rasterPartsForced();
			if(parts <= 0) parts = tilesInPlay * 4 * 4 <= warpSlots * 5 ? 4 : (tilesInPlay * 2 * 4 <= warpSlots * 5 ? 2 : 1);
			const unsigned blocks = ((unsigned)(P.tilesX * P.tilesY) * (unsigned)parts + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK;
			tile_raster_depth_kernel<<<blocks, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P, Q, p->tileStart.p, sortedTris, parts);
			p->launches++;
		}
		{
			ProfScope ps(p, CLS_SHADE);
			pe->shade(P, Q, p->stream);
			p->launches++;
		}
	}
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

// Reads the verdict of the last draw's tile scan (waits for that kernel only — by the next API call it is long done).
// A draw whose speculated sizes were too small did nothing behind its scan: size the buffers exactly and enqueue its
// tail again, before anything else reaches the stream. Every entry point that touches the stream or device memory
// settles first.
static int settle(ps3d_pipe* p)
{
	if(!p->pending.valid) return PS3D_OK;
	p->pending.valid = false;
	CK(p, cudaEventSynchronize(p->scanEvent));
	const DrawReport r = *p->report;
	const DrawParams& P = p->pending.P;
	const ProgEntry* pe = p->pending.pe;
	int path = p->pending.path;
	p->profPairs += p->profiling ? r.pairs : 0;
	if(p->pending.span)
	{
		// span path: the geometry kernel itself depends on a speculated capacity (span records, list capacity), so a draw whose
		// sizes did not hold is run again from its geometry kernel with the exact sizes (its counters were not added, its lists
		// and counts were reset by the plan kernel); a list too long for the shared-memory sort sends the draw down the first path
		if(r.spans > p->spanHigh) p->spanHigh = r.spans;
		if(r.longest > p->listHigh) p->listHigh = r.longest;
		const bool multi = p->pending.multi;
		if(r.fragBound > p->survivorHigh && !spanNoShade(p, P, pe, multi)) p->survivorHigh = (size_t)r.fragBound;
		if(r.marks > p->markHigh) p->markHigh = r.marks;
		const bool blend = p->pending.blend;
		if(!r.bad)
		{
			if(p->pending.tailLaunched) return PS3D_OK;
			return launchSpanTail(p, P, pe, multi, blend);
		}
		CK(p, cudaMemsetAsync(p->poisonDev, 0, 4, p->stream));
		const int vao = p->pending.vao;
		if(r.longest > PS_SORT_LIMIT || r.fragBound >= 0xfffffff0ull || r.spans >= 0xfffffff0u)
		{
			const int path = (blend || r.fragBound >= 0xfffffff0ull) ? 1 : 2;
			if(multi)
			{
				// a batch whose lists outgrew the shared-memory sort: its draws one by one down the first path (nothing of the batch
				// has touched the targets: the poisoned tail did not run)
				const ps3d_pipe::Batch B = p->flight;
				for(const ps3d_pipe::BatchDraw& d : B.draws)
				{
					int rc = enqueueLegacy(p, d.P, d.pe, path, d.vao);
					if(!rc) rc = settle(p);
					if(rc) return rc;
				}
				return PS3D_OK;
			}
			if(vao >= 0) { if((int)p->vaoLegacy.size() <= vao) p->vaoLegacy.resize(vao + 1, 0); p->vaoLegacy[vao] = 1; }
			const DrawParams P2 = P;
			return enqueueLegacy(p, P2, pe, path, vao);
		}
		const DrawParams P2 = P;
		return enqueueSpan(p, P2, pe, vao, r.spans, r.longest, (size_t)r.fragBound, multi, blend);
	}
	if(r.pairs > p->pairHigh) p->pairHigh = r.pairs;
	if(2 == path && r.fragBound > p->survivorHigh) p->survivorHigh = (size_t)r.fragBound;
	if(p->speculate && !radixBinningForced() && !r.bad) return PS3D_OK;   // the guarded tail ran
	CK(p, cudaMemsetAsync(p->poisonDev, 0, 4, p->stream));
	if(0 == r.pairs) return PS3D_OK;
	if(2 == path && r.fragBound >= 0xfffffff0ull) path = 1;
	if(2 == path && 0 == r.fragBound) return PS3D_OK;
	const bool radix = r.longest > PS_SORT_LIMIT || radixBinningForced();
	CK(p, p->valsA.ensure(r.pairs));
	if(2 == path) { int rc = ensureSurvivors(p, (size_t)r.fragBound, P); if(rc) return rc; }
	return launchTail(p, P, pe, path, radix);
}

// VBO read events: the geometry kernels are the only readers of the vertex streams; asynchronous uploads may overwrite them behind
// the last one. (Every attached VBO, also one that has only ever been written synchronously: its first asynchronous write or
// all-gather must wait for THIS read, not for the creation-time fill.)
static int recordVboReads(ps3d_pipe* p, int vao)
{
	if(vao < 0 || vao >= (int)p->vaos.size() || !p->vaos[vao].alive) return PS3D_OK;
	if(p->capturing) { p->capVaos.push_back(vao); return PS3D_OK; }   // recorded behind every launch of the captured frame instead
	const Vao& va = p->vaos[vao];
	for(int s = 0; s < PS3D_MAX_VBOS; s++)
		if(va.vbo[s] >= 0 && p->vbos[va.vbo[s]].alive)
		{
			Vbo& v = p->vbos[va.vbo[s]];
			{ const int rc = vboEvents(p, v); if(rc) return rc; }
			CK(p, cudaEventRecord(v.lastRead, p->stream));
			v.readValid = true;
		}
	return PS3D_OK;
}

// Sort-first over peer memory: the first colour write of a frame takes rank 0's target (ranks != 0: wait until rank 0 has handed
// it out) or hands it out (rank 0). Called in front of every kernel that writes colour.
static int peerFirstWrite(ps3d_pipe* p)
{
	if(!p->peer.active || !p->peer.needTake) return PS3D_OK;
	p->peer.needTake = false;
	if(0 == p->peer.rank) peer_release_kernel<<<1, 1, 0, p->stream>>>(p->peer.ctr, p->peer.flags0, p->back);
	else peer_take_kernel<<<1, 1, 0, p->stream>>>(p->peer.ctr, p->peer.flags0, p->back);
	p->launches++;
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

// the programme groups of the batch in flight: distinct programmes in order of first appearance
static void batchGroups(const ps3d_pipe::Batch& B, std::vector<const ProgEntry*>& groups)
{
	groups.clear();
	for(const ps3d_pipe::BatchDraw& d : B.draws)
		if(std::find(groups.begin(), groups.end(), d.pe) == groups.end()) groups.push_back(d.pe);
}

// a draw (a batch) whose fragment functors do nothing — the depth-only pass of a shadow map — has no shade kernel to run and
// keeps no survivor stream; the raster kernel still counts its survivors. (Not while the per-pixel counts of the parity hook
// are on: they are taken by the shade kernel.)
static bool spanNoShade(const ps3d_pipe* p, const DrawParams& P, const ProgEntry* pe, bool multi)
{
	if(P.cap) return false;
	if(!multi) return pe->noop;
	for(const ps3d_pipe::BatchDraw& d : p->flight.draws) if(!d.pe->noop) return false;
	return true;
}

static int launchSpanTail(ps3d_pipe* p, const DrawParams& P, const ProgEntry* pe, bool multi, bool blend)
{
	const uint32_t ntiles = (uint32_t)(P.tilesX * P.tilesY);
	SurvivorStream2 Q;
	Q.span = p->sv2Span.p; Q.xy = p->sv2XY.p; Q.inv = p->sv2Inv.p; Q.count = p->svCountDev;
	Q.capacity = (uint32_t)std::min<size_t>(p->sv2Span.cap, 0xfffffff0u);
	Q.next = blend ? p->sv2Next.p : nullptr; Q.chain = blend ? p->sv2Chain.p : nullptr; Q.colour = blend ? p->sv2Colour.p : nullptr;
	const bool noShade = !blend && spanNoShade(p, P, pe, multi);
	if(noShade) { Q.span = nullptr; Q.xy = nullptr; Q.inv = nullptr; Q.capacity = 0; }
	{
		ProfScope ps(p, CLS_BIN);
		tile_list_sort_cap_kernel<<<(ntiles + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P.tl, ntiles, p->poisonDev, p->tileOrder.p);
		p->launches++;
	}
	CK(p, cudaMemsetAsync(p->svCountDev, 0, 4, p->stream));
	const bool marks = P.sp.markCap > 0;
	int partsUsed = 1;
	// (lane = long span, grid-stride: the count is on the device; a draw's spans bound it)
	const unsigned markBlocks = (unsigned)std::min<size_t>(((size_t)P.sp.capacity + PS_MARK_THREADS - 1) / PS_MARK_THREADS, (size_t)p->smCount * 8);
	{
		ProfScope ps(p, CLS_TILE);
		// fewer tiles in play than warp slots on the GPU: cut every tile into 2 or 4 row groups, one warp each
		const int bandRows = std::max(0, std::min(P.band1, P.vpH) - std::max(P.band0, 0));
		const long long tilesInPlay = (long long)P.tilesX * ((bandRows + PS_TILE - 1) / PS_TILE);
		const long long warpSlots = (long long)p->smCount * 32;      // resident warps of this kernel
		// (a tile's list is one dependent chain, ~100 us on C2 whatever the load: as soon as the tiles in play no longer fill the
		// GPU's warp slots about twice over, shorter chains win — 2, 4 or 8 row groups per tile)
		int parts = rasterPartsForced();
		if(parts <= 0) parts = tilesInPlay * 4 <= warpSlots ? 8 : (tilesInPlay * 2 <= warpSlots ? 4 : (tilesInPlay <= warpSlots * 2 ? 2 : 1));
		partsUsed = parts;
		const unsigned blocks = (ntiles * (unsigned)parts + PS_WARPS_PER_BLOCK - 1) / PS_WARPS_PER_BLOCK;
		// PS3D_RASTER_MINB=8|10|12: blocks per SM the kernel is compiled for (64 / 51 / 40 registers) — A/B switch
		static int minb = -1;
		if(minb < 0) { const char* e = getenv("PS3D_RASTER_MINB"); minb = e ? atoi(e) : 8; }
		if(marks)
		{
			// the depth chains of the long spans, walked once (kernels_span.cuh: span_mark_depth_kernel), then the variant that uses them
			span_mark_depth_kernel<<<markBlocks, PS_MARK_THREADS, 0, p->stream>>>(P);
			p->launches++;
			tile_raster_span_kernel<8, true><<<blocks, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P, Q, parts);
		}
		else if(12 == minb) tile_raster_span_kernel<12><<<blocks, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P, Q, parts);
		else if(10 == minb) tile_raster_span_kernel<10><<<blocks, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P, Q, parts);
		else tile_raster_span_kernel<8><<<blocks, 32 * PS_WARPS_PER_BLOCK, 0, p->stream>>>(P, Q, parts);
		p->launches++;
	}
	{ const int rc = peerFirstWrite(p); if(rc) return rc; }
	if(noShade) { CK(p, cudaGetLastError()); return PS3D_OK; }
	if(multi)
	{
		ProfScope ps(p, CLS_SHADE);
		std::vector<const ProgEntry*> groups;
		batchGroups(p->flight, groups);
		for(size_t g = 0; g < groups.size(); g++)
		{
			const BatchView V = { p->batchItems.p, p->batchBlockDraw.p, p->batchBlockList.p, (uint32_t)g };
			// (markV is indexed mark * NV with the programme's own NV: two programmes' marks may share bytes — each group's marks are
			// written and consumed before the next group's are written, in stream order)
			if(marks && groups[g]->nv > 0) { groups[g]->markVary(P, &V, markBlocks, p->stream); p->launches++; }
			groups[g]->shadeSpanMulti(P, Q, V, marks, p->stream);
			p->launches++;
		}
	}
	else
	{
		ProfScope ps(p, CLS_SHADE);
		if(marks && pe->nv > 0) { pe->markVary(P, nullptr, markBlocks, p->stream); p->launches++; }
		pe->shadeSpan(P, Q, marks, p->stream);
		p->launches++;
		if(blend)
		{
			// every pixel's colours applied in submission order (kernels_span.cuh: shade_resolve_kernel)
			shade_resolve_kernel<<<(ntiles * (unsigned)partsUsed + 3) / 4, 128, 0, p->stream>>>(P, Q, partsUsed);
			p->launches++;
		}
	}
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

// The span path (kernels_span.cuh). spans / longest / survivors = 0: capacities speculated from the high-water marks of earlier
// draws; otherwise the exact sizes a first attempt reported (settle()).
// multi: P is the batch-wide DrawParams of p->flight (ps3d_pipe::Batch), ntris = its blocks x PS_GEOM_THREADS ids.
static int enqueueSpan(ps3d_pipe* p, DrawParams P, const ProgEntry* pe, int vao, size_t spans, size_t longest, size_t survivors, bool multi, bool blend)
{
	const uint32_t ntiles = (uint32_t)(P.tilesX * P.tilesY);
	const size_t ntris = P.ntris;
	const bool exact = spans || longest || survivors;
	CK(p, p->hdr.ensure(ntris));
	size_t varyCount = ntris * 3 * (size_t)(pe->nv > 0 ? pe->nv : 1);
	if(multi)
	{
		varyCount = 1;
		for(const ps3d_pipe::BatchDraw& d : p->flight.draws) varyCount += (size_t)d.P.ntris * 3 * (size_t)d.pe->nv;
		CK(p, p->batchItems.ensure(p->flight.draws.size()));
		CK(p, p->batchBlockDraw.ensure(p->flight.blocks));
		CK(p, p->batchBlockList.ensure(p->flight.blocks));
	}
	CK(p, p->vary.ensure(varyCount));
	CK(p, p->spTri.ensure(ntris));
	// (a frame being captured cannot allocate: it takes the buffers the same frame ran in a moment ago as they are — they hold
	// its high-water marks — instead of asking for the usual 25 % of head-room)
	size_t spanCap = exact ? spans + 1 : std::max(p->spanHigh + p->spanHigh / 4, ntris * 6 + 4096);
	if(p->capturing && p->spRec.cap >= p->spanHigh) spanCap = std::min(spanCap, p->spRec.cap);
	CK(p, p->spRec.ensure(spanCap));
	size_t listCap = exact ? longest : std::max(p->listHigh + p->listHigh / 4, (size_t)128);
	listCap = std::min<size_t>((listCap + 31) & ~(size_t)31, PS_SORT_LIMIT);
	if(p->capturing && (size_t)ntiles * listCap > p->tlIds.cap)
	{
		const size_t fits = (p->tlIds.cap / ntiles) & ~(size_t)31;
		if(fits >= p->listHigh) listCap = fits;
	}
	CK(p, p->tlIds.ensure((size_t)ntiles * listCap));
	{
		// tile_plan_kernel leaves the per-tile fill counts zeroed behind every draw; a fresh allocation starts zeroed
		const uint32_t* before = p->tlFill.p;
		CK(p, p->tlFill.ensure(ntiles + 1)); CK(p, p->tlLen.ensure(ntiles + 1)); CK(p, p->tileOrder.ensure(ntiles + 1));
		if(p->tlFill.p != before) CK(p, cudaMemsetAsync(p->tlFill.p, 0, p->tlFill.cap * 4, p->stream));
	}
	const size_t svSlack = 0;
	const bool noShade = !blend && spanNoShade(p, P, pe, multi);
	if(!noShade)
	{
		size_t svCap = (exact ? survivors + 1 : std::max(p->survivorHigh + p->survivorHigh / 4, (size_t)P.vpW * P.vpH * 2)) + svSlack;
		if(p->capturing && p->sv2Span.cap >= p->survivorHigh) svCap = std::min(svCap, p->sv2Span.cap);
		CK(p, p->sv2Span.ensure(svCap)); CK(p, p->sv2XY.ensure(svCap)); CK(p, p->sv2Inv.ensure(svCap));
		if(blend) { CK(p, p->sv2Next.ensure(p->sv2Span.cap)); CK(p, p->sv2Colour.ensure(p->sv2Span.cap)); CK(p, p->sv2Chain.ensure((size_t)ntiles * 16)); }
	}
	P.hdr = p->hdr.p; P.vary = p->vary.p; P.tileOrder = p->tileOrder.p; P.poison = p->poisonDev;
	if(P.band0 > 0 || P.band1 < P.vpH)
	{
		CK(p, p->workList.ensure(ntris + 4));
		P.workList = p->workList.p + 4; P.workCount = p->workList.p;      // the counter lives in front of the list
		CK(p, cudaMemsetAsync(P.workCount, 0, 4, p->stream));
	}
	P.batchFirstBlock = 0;
	P.batchTrisPerBlock = (!multi && !P.workList && ntris <= PS_BATCH_DRAW_TRIS) ? batchTrisPerBlock(ntris) : 0;
	P.sp.rec = p->spRec.p; P.sp.tri = p->spTri.p; P.sp.count = p->spanCountDev;
	P.sp.capacity = (uint32_t)std::min<size_t>(p->spRec.cap, 0xfffffff0u);
	P.tl.fill = p->tlFill.p; P.tl.len = p->tlLen.p; P.tl.ids = p->tlIds.p; P.tl.cap = (uint32_t)listCap;
	// very large tile rectangles are appended by a kernel of their own — for draws that follow one with many long spans (the
	// same scenes have both; one more launch per draw does not pay elsewhere)
	const bool bigAppend = p->markHigh >= p->marksMin && p->markHigh > 0 && marksOn();
	P.tl.bigList = bigAppend ? p->bigListDev + 4 : nullptr; P.tl.bigCount = p->bigListDev; P.tl.bigCap = PS_BIG_LIST;
	// chain marks of long spans: once some draw has asked for marks (its report said how many), room for them is kept; a draw
	// whose marks do not fit replays those spans from their start, as every long span does while no marks are kept
	P.sp.longCount = p->longCountDev; P.sp.longLatched = p->longLatchedDev;
	P.sp.markAt = nullptr; P.sp.longList = nullptr; P.sp.markZ = nullptr; P.sp.markV = nullptr; P.sp.markCap = 0;
	if(p->markHigh >= p->marksMin && p->markHigh > 0 && marksOn())
	{
		size_t nvMax = (size_t)pe->nv;
		if(multi) for(const ps3d_pipe::BatchDraw& d : p->flight.draws) nvMax = std::max(nvMax, (size_t)d.pe->nv);
		size_t markCap = std::min<size_t>(p->markHigh + p->markHigh / 4 + 1024, 0xfffffff0u);
		bool room = true;
		if(p->capturing)
		{
			// (a captured frame cannot allocate: what the same frame left behind a moment ago, or no marks)
			room = p->spMarkAt.cap >= p->spRec.cap && p->spLongList.cap >= p->spRec.cap && p->spMarkZ.cap > 0 && (0 == nvMax || p->spMarkV.cap / nvMax > 0);
			if(room) { markCap = std::min(markCap, p->spMarkZ.cap); if(nvMax) markCap = std::min(markCap, p->spMarkV.cap / nvMax); }
		}
		if(room)
		{
			CK(p, p->spMarkAt.ensure(p->spRec.cap)); CK(p, p->spLongList.ensure(p->spRec.cap));
			CK(p, p->spMarkZ.ensure(markCap));
			if(nvMax) CK(p, p->spMarkV.ensure(markCap * nvMax));
			P.sp.markAt = p->spMarkAt.p; P.sp.longList = p->spLongList.p; P.sp.markZ = p->spMarkZ.p; P.sp.markV = p->spMarkV.p;
			P.sp.markCap = (uint32_t)markCap;
		}
	}
	if(multi)
	{
		// every draw's DrawParams into the table (shifted by its first id), its blocks into the block tables; then one geometry
		// launch per programme over that programme's blocks
		ProfScope ps(p, CLS_GEOM);
		std::vector<const ProgEntry*> groups;
		batchGroups(p->flight, groups);
		std::vector<uint32_t> groupBlocks(groups.size(), 0), groupAt(groups.size(), 0), groupFill(groups.size(), 0);
		for(const ps3d_pipe::BatchDraw& d : p->flight.draws)
			groupBlocks[std::find(groups.begin(), groups.end(), d.pe) - groups.begin()] += d.nBlocks;
		for(size_t g = 1; g < groups.size(); g++) groupAt[g] = groupAt[g - 1] + groupBlocks[g - 1];
		size_t varyAt = 0;
		static thread_local ItemPack pack;                         // (28 KB: not on the stack)
		uint32_t packed = 0;
		for(size_t i = 0; i < p->flight.draws.size(); i++)
		{
			const ps3d_pipe::BatchDraw& d = p->flight.draws[i];
			const size_t g = std::find(groups.begin(), groups.end(), d.pe) - groups.begin();
			if(0 == packed) pack.first = (uint32_t)i;
			DrawParams& item = pack.item[packed];
			item = P;                                               // targets, buffers, capacities: the batch's
			memcpy(item.slot, d.P.slot, sizeof(item.slot));
			memcpy(item.stride, d.P.stride, sizeof(item.stride));
			memcpy(item.u, d.P.u, sizeof(item.u));
			memcpy(item.tex, d.P.tex, sizeof(item.tex));
			item.ntris = d.P.ntris;
			item.batchFirstBlock = d.firstBlock; item.batchTrisPerBlock = d.trisPerBlock;
			item.vary = p->vary.p + varyAt;
			varyAt += (size_t)d.P.ntris * 3 * (size_t)d.pe->nv;
			pack.nBlocks[packed] = d.nBlocks; pack.listAt[packed] = groupAt[g] + groupFill[g]; pack.tag[packed] = (uint32_t)i | ((uint32_t)g << 16);
			groupFill[g] += d.nBlocks;
			packed++;
			if(PS_BATCH_PACK == packed || i + 1 == p->flight.draws.size())
			{
				batch_items_kernel<<<packed, 256, 0, p->stream>>>(pack, p->batchItems.p, p->batchBlockDraw.p, p->batchBlockList.p);
				p->launches++;
				packed = 0;
			}
		}
		for(size_t g = 0; g < groups.size(); g++)
		{
			const BatchView V = { p->batchItems.p, p->batchBlockDraw.p, p->batchBlockList.p + groupAt[g], (uint32_t)g };
			groups[g]->geomSpanMulti(V, groupBlocks[g], p->stream);
			p->launches++;
		}
		CK(p, cudaGetLastError());
		for(const ps3d_pipe::BatchDraw& d : p->flight.draws) { const int rc = recordVboReads(p, d.vao); if(rc) return rc; }
	}
	else
	{
		{
			ProfScope ps(p, CLS_GEOM);
			pe->geomSpan(P, p->stream);
			p->launches++;
		}
		CK(p, cudaGetLastError());
		{ const int rc = recordVboReads(p, vao); if(rc) return rc; }
	}
	if(bigAppend)
	{
		ProfScope ps(p, CLS_GEOM);
		tile_append_big_kernel<<<p->smCount * 4, 256, 0, p->stream>>>(P.tl, P.tilesX);
		p->launches++;
	}
	{
		ProfScope ps(p, CLS_BIN);
		tile_plan_kernel<<<1, 1024, 0, p->stream>>>(P.tl, ntiles, p->statsDev, p->spanCountDev, P.sp.capacity,
		                                          noShade ? 0xffffffffffull : (unsigned long long)(std::min<size_t>(p->sv2Span.cap, 0xfffffff0u) - svSlack), PS_SORT_LIMIT,
		                                          p->poisonDev, p->reportDev, p->tileOrder.p, p->longCountDev, p->longLatchedDev);
		p->launches++;
	}
	if(p->capturing)
	{
		// a captured frame cannot be judged by the host draw by draw: its tail is guarded by the poison word as always, and a
		// replay that did not fit is reported through DrawReport::sticky at the next ps3d_finish
		const int rc = launchSpanTail(p, P, pe, multi, blend);
		return rc;
	}
	CK(p, cudaEventRecord(p->scanEvent, p->stream));
	p->pending.valid = true; p->pending.P = P; p->pending.pe = pe; p->pending.path = 2; p->pending.span = true; p->pending.vao = vao;
	p->pending.multi = multi; p->pending.blend = blend;
	p->pending.tailLaunched = false;
	if(!p->speculate && !exact) return settle(p);      // sizes checked on the host before anything else is enqueued
	p->pending.tailLaunched = true;
	const int rc = launchSpanTail(p, P, pe, multi, blend);
	if(rc) { p->pending.valid = false; return rc; }
	return PS3D_OK;
}

// PS3D_MARKS=0: no chain marks, long spans replay from their start (A/B runs)
static bool marksOn() { static int on = -1; if(on < 0) { const char* e = getenv("PS3D_MARKS"); on = (e && e[0] == '0') ? 0 : 1; } return on == 1; }
// PS3D_BATCH=0: every draw is launched as it is submitted (A/B runs)
static bool batchingOn() { static int on = -1; if(on < 0) { const char* e = getenv("PS3D_BATCH"); on = (e && e[0] == '0') ? 0 : 1; } return on == 1; }
// a draw can join the batch being collected: same targets, viewport, band and behaviour bits (any programme of the span path)
static bool batchTakes(const ps3d_pipe* p, const DrawParams& P)
{
	if(p->batch.draws.empty()) return true;
	const DrawParams& A = p->batch.draws[0].P;
	return A.behavior == P.behavior && A.vpW == P.vpW && A.vpH == P.vpH && A.band0 == P.band0 && A.band1 == P.band1 && A.cap == P.cap
	    && 0 == memcmp(&A.colour, &P.colour, sizeof(A.colour)) && 0 == memcmp(&A.depth, &P.depth, sizeof(A.depth))
	    && p->batch.draws.size() < PS_BATCH_MAX_DRAWS && p->batch.blocks + batchBlocks(P.ntris) <= PS_BATCH_MAX_BLOCKS;
}

// Launches the collected draws: one alone as any draw, several as one batch (kernels_span.cuh: BatchView).
static int flushBatch(ps3d_pipe* p)
{
	if(p->batch.draws.empty()) return PS3D_OK;
	{ const int rc = settle(p); if(rc) { p->batch.clear(); return rc; } }   // the verdict of what is on the stream first
	if(1 == p->batch.draws.size())
	{
		const ps3d_pipe::BatchDraw d = p->batch.draws[0];
		p->batch.clear();
		return enqueueSpan(p, d.P, d.pe, d.vao, 0, 0, 0);
	}
	p->flight.draws.swap(p->batch.draws);
	p->flight.blocks = p->batch.blocks;
	p->batch.clear();
	DrawParams P = p->flight.draws[0].P;
	P.ntris = p->flight.blocks * PS_GEOM_THREADS;
	for(int sl = 0; sl < 16; sl++) { P.slot[sl] = nullptr; P.stride[sl] = 0; }
	p->batchesLaunched++; p->drawsBatched += p->flight.draws.size();
	return enqueueSpan(p, P, p->flight.draws[0].pe, -1, 0, 0, 0, true);
}

// The first path (kernels.cuh): geometry with per-thread row walks, counted binning (tile scan + fill + sort), tile kernels that
// re-derive spans from triangle headers. Draws that blend or may discard, and the fallback of the span path.
static int enqueueLegacy(ps3d_pipe* p, DrawParams P, const ProgEntry* pe, int path, int vao)
{
	P.workList = nullptr; P.workCount = nullptr;       // (a draw handed over by the span path brings its own)
	const uint32_t ntiles = (uint32_t)(P.tilesX * P.tilesY);
	const size_t ntris = P.ntris;
	CK(p, p->hdr.ensure(ntris));
	CK(p, p->vary.ensure(ntris * 3 * (size_t)(pe->nv > 0 ? pe->nv : 1)));
	CK(p, p->triCount.ensure(ntris));
	CK(p, p->triRect.ensure(ntris * 3));
	{
		// tile_scan_kernel leaves the per-tile counts zeroed behind every draw; a fresh allocation starts zeroed
		const uint32_t* before = p->tileCount.p;
		CK(p, p->tileCount.ensure(ntiles + 1)); CK(p, p->tileStart.ensure(ntiles + 1)); CK(p, p->tileFill.ensure(ntiles + 1)); CK(p, p->tileOrder.ensure(ntiles + 1));
		if(p->tileCount.p != before) CK(p, cudaMemsetAsync(p->tileCount.p, 0, p->tileCount.cap * 4, p->stream));
	}
	P.hdr = p->hdr.p; P.vary = p->vary.p; P.triCount = p->triCount.p; P.triRect = p->triRect.p; P.tileCount = p->tileCount.p; P.tileOrder = p->tileOrder.p;
	P.poison = p->poisonDev;
	// two-kernel geometry pays below about a sixth of the rows (band probe: an eighth 0.085 -> 0.067 ms, a quarter 0.092 -> 0.100 ms:
	// the list-driven half reads its vertex streams sparsely); PS3D_GEOM_SPLIT=1 forces it for any band, =0 never
	if(geomSplitWanted(P.band1 - P.band0, P.vpH) && (P.band0 > 0 || P.band1 < P.vpH))
	{
		CK(p, p->workList.ensure(ntris + 4));
		P.workList = p->workList.p + 4; P.workCount = p->workList.p;      // the counter lives in front of the list
		CK(p, cudaMemsetAsync(P.workCount, 0, 4, p->stream));
	}

	// Speculation: the draw's tail is enqueued right behind the tile scan, sized by the high-water marks of earlier draws;
	// the scan checks the sizes on the device and the host reads its verdict at the next API call (settle()).
	const bool speculate = p->speculate && !radixBinningForced();
	uint32_t pairCap = 0xffffffffu, listLimit = 0xffffffffu;
	unsigned long long survivorCap = ~0ull;
	if(speculate)
	{
		size_t pairGuess = std::max(p->pairHigh + p->pairHigh / 4, ntris + ntris / 2 + 1024);
		if(p->capturing && p->valsA.cap >= p->pairHigh) pairGuess = std::min(pairGuess, p->valsA.cap);   // (a captured frame cannot allocate)
		CK(p, p->valsA.ensure(pairGuess));
		pairCap = (uint32_t)std::min<size_t>(p->valsA.cap, 0xffffffffu);
		listLimit = PS_SORT_LIMIT;
		if(2 == path)
		{
			size_t svGuess = std::max(p->survivorHigh + p->survivorHigh / 4, (size_t)P.vpW * P.vpH * 2);
			if(p->capturing && p->svTri.cap >= p->survivorHigh) svGuess = std::min(svGuess, p->svTri.cap);
			{ const int rc = ensureSurvivors(p, svGuess, P); if(rc) return rc; }
			survivorCap = std::min<size_t>(p->svTri.cap, 0xfffffff0u);
		}
	}
	{
		ProfScope ps(p, CLS_GEOM);
		pe->geom(P, p->stream);
		p->launches++;
	}
	CK(p, cudaGetLastError());
	{ const int rc = recordVboReads(p, vao); if(rc) return rc; }
	{
		ProfScope ps(p, CLS_BIN);
		tile_scan_kernel<<<1, 1024, 0, p->stream>>>(p->tileCount.p, p->tileStart.p, p->tileFill.p, ntiles, p->statsDev,
		                                          pairCap, survivorCap, listLimit, p->poisonDev, p->reportDev, p->tileOrder.p);
		p->launches++;
	}
	if(p->capturing)
	{
		if(!speculate) return fail(p, PS3D_ERR_UNSUPPORTED, "a captured frame needs speculated capacities (PS3D_SPECULATE=0 / PS3D_BINNING=radix are set)");
		return launchTail(p, P, pe, path, false);
	}
	CK(p, cudaEventRecord(p->scanEvent, p->stream));
	p->pending.valid = true; p->pending.P = P; p->pending.pe = pe; p->pending.path = path; p->pending.span = false; p->pending.tailLaunched = true; p->pending.vao = vao; p->pending.multi = false; p->pending.blend = false;
	if(!speculate) return settle(p);          // exact sizes after a host sync in the middle of the draw
	int rc = launchTail(p, P, pe, path, false);
	if(rc) { p->pending.valid = false; return rc; }
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

extern "C" {

const char* ps3d_backend_name(void) { return "cuda-sm100a"; }
const char* ps3d_last_error(const ps3d_pipe* p) { return p ? p->err.c_str() : ""; }

int ps3d_create(int width, int height, int device, ps3d_pipe** out)
{
	TRACE();
	if(!out || width <= 0 || height <= 0) return PS3D_ERR_INVALID_ARGUMENT;
	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return PS3D_ERR_DEVICE; // no GPU, no renderer
	if(cudaSetDevice(device) != cudaSuccess) return PS3D_ERR_DEVICE;
	ps3d_pipe* p = new ps3d_pipe();
	p->device = device;
	{ cudaDeviceProp prop; p->smCount = (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ? prop.multiProcessorCount : 148; }
	p->width = width; p->height = height; p->vpW = width; p->vpH = height;
	p->behavior = PS3D_BEHAVIOR_UPDATE_DEPTH | PS3D_BEHAVIOR_TEST_DEPTH | PS3D_BEHAVIOR_FACE_CULLING; // pipeline.cpp:34
	p->band0 = 0; p->band1 = 0x7fffffff;
	p->back = 1; p->depthTex = -1; p->curProg = -1;
	p->capDev = nullptr; p->capW = p->capH = 0;
	p->launches = 0;
	p->profiling = false; p->profPairs = 0; p->profLaunches[0] = p->profLaunches[1] = p->profLaunches[2] = p->profLaunches[3] = 0;
	memset(&p->stats, 0, sizeof(p->stats));
	memset(p->uniformSet, 0, sizeof(p->uniformSet));
	p->depthScanline = ((int)(width / 4.0f + 0.5f) * 4) * (int)sizeof(float); // pipeline.cpp:31
	if(p->depthScanline < width * 4) p->depthScanline = width * 4;           // the reference under-allocates when W%4==1
	bool ok = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&p->copyStream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&p->readStream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&p->frameDone, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&p->readDone[0], cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&p->readDone[1], cudaEventDisableTiming) == cudaSuccess;
	p->readValid[0] = p->readValid[1] = false;
	p->commFrame = p->commUpload = nullptr; p->gatherStream = nullptr; p->commRank = 0; p->commWorld = 1;
	const size_t cbytes = (size_t)width * 4 * height, dbytes = (size_t)p->depthScanline * height;
	ok = ok && cudaMalloc((void**)&p->display[0], cbytes) == cudaSuccess && cudaMalloc((void**)&p->display[1], cbytes) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->defaultDepth, dbytes + 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->totalDev, 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->statsDev, sizeof(DeviceStats) * PS_STATS_COPIES) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->poisonDev, 16) == cudaSuccess;
	ok = ok && cudaHostAlloc((void**)&p->report, sizeof(DrawReport), cudaHostAllocMapped) == cudaSuccess;
	ok = ok && cudaHostGetDevicePointer((void**)&p->reportDev, p->report, 0) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&p->scanEvent, cudaEventDisableTiming) == cudaSuccess;
	p->pending.valid = false; p->pairHigh = p->survivorHigh = 0;
	{
		const char* e = getenv("PS3D_SPECULATE");   // PS3D_SPECULATE=0: size every draw exactly after a mid-draw host sync (A/B checks)
		p->speculate = !(e && e[0] == '0');
	}
	ok = ok && cudaMalloc((void**)&p->svCountDev, 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->spanCountDev, 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->longCountDev, 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->longLatchedDev, 16) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->bigListDev, (4 + 3 * PS_BIG_LIST) * sizeof(uint32_t)) == cudaSuccess;
	memset(&p->peer, 0, sizeof(p->peer));
	p->capturing = false; p->graphLaunched = false; p->capBack = 0;
	ok = ok && cudaMalloc((void**)&p->peer.flagsOwn, sizeof(PeerFlags)) == cudaSuccess;
	ok = ok && cudaMalloc((void**)&p->peer.ctr, sizeof(PeerCounters)) == cudaSuccess;
	if(ok) { cudaMemsetAsync(p->peer.flagsOwn, 0, sizeof(PeerFlags), p->stream); cudaMemsetAsync(p->peer.ctr, 0, sizeof(PeerCounters), p->stream); }
	p->spanHigh = p->listHigh = 0;
	p->markHigh = 0;
	{ const char* e = getenv("PS3D_MARKS_MIN"); p->marksMin = e ? (size_t)atoll(e) : 16384; }
	p->batch.clear(); p->flight.clear(); p->batchesLaunched = p->drawsBatched = 0; p->pending.multi = false; p->pending.blend = false;
	p->pending.span = false; p->pending.tailLaunched = false; p->pending.vao = -1;
	if(ok)
	{
		cudaMemsetAsync(p->spanCountDev, 0, 16, p->stream);
		cudaMemsetAsync(p->longCountDev, 0, 16, p->stream);
		cudaMemsetAsync(p->longLatchedDev, 0, 16, p->stream);
		cudaMemsetAsync(p->bigListDev, 0, 16, p->stream);
		cudaMemsetAsync(p->display[0], 0, cbytes, p->stream);
		cudaMemsetAsync(p->display[1], 0, cbytes, p->stream);
		cudaMemsetAsync(p->defaultDepth, 0, dbytes, p->stream);
		cudaMemsetAsync(p->statsDev, 0, sizeof(DeviceStats) * PS_STATS_COPIES, p->stream);
		cudaMemsetAsync(p->poisonDev, 0, 16, p->stream);
		memset(p->report, 0, sizeof(DrawReport));
	}
	p->rcpDev = p->rsqrtDev = nullptr;
	p->approx.rcp = p->approx.rsqrt = nullptr; p->approx.rcpBits = p->approx.rsqrtBits = 0;
	const Ps3dHostApprox& ha = hostApprox();
	if(ok && ha.rcpBits && ha.rsqrtBits)
	{
		ok = cudaMalloc((void**)&p->rcpDev, ha.rcp.size() * 4) == cudaSuccess && cudaMalloc((void**)&p->rsqrtDev, ha.rsqrt.size() * 4) == cudaSuccess;
		if(ok)
		{
			cudaMemcpyAsync(p->rcpDev, ha.rcp.data(), ha.rcp.size() * 4, cudaMemcpyHostToDevice, p->stream);
			cudaMemcpyAsync(p->rsqrtDev, ha.rsqrt.data(), ha.rsqrt.size() * 4, cudaMemcpyHostToDevice, p->stream);
			p->approx.rcp = p->rcpDev; p->approx.rsqrt = p->rsqrtDev; p->approx.rcpBits = ha.rcpBits; p->approx.rsqrtBits = ha.rsqrtBits;
		}
	}
	if(!ok || cudaStreamSynchronize(p->stream) != cudaSuccess)
	{
		delete p;
		return PS3D_ERR_DEVICE;
	}
	*out = p;
	return PS3D_OK;
}

int ps3d_destroy(ps3d_pipe* p)
{
	TRACE();
	if(!p) return PS3D_ERR_INVALID_ARGUMENT;
	cudaSetDevice(p->device);
	if(p->capturing)
	{
		// a capture nobody ended: end it here, or the thread stays in capture mode
		cudaGraph_t g = nullptr;
		cudaStreamEndCapture(p->stream, &g);
		if(g) cudaGraphDestroy(g);
		cudaGetLastError();
		p->capturing = false; g_capturing = false;
	}
	p->batch.clear();                                  // draws nobody asked the result of
	settle(p);
	cudaStreamSynchronize(p->stream);
	for(Texture* t : p->textures) if(t) { for(int i = 0; i < 6; i++) if(t->layer[i]) cudaFree(t->layer[i]); delete t; }
	cudaStreamSynchronize(p->copyStream); cudaStreamSynchronize(p->readStream);
	ps3d_comm_destroy(p);
	for(Vbo& v : p->vbos) if(v.alive) freeVbo(v);
	cudaFree(p->display[0]); cudaFree(p->display[1]); cudaFree(p->defaultDepth);
	cudaFree(p->totalDev); cudaFree(p->statsDev); cudaFree(p->svCountDev); cudaFree(p->poisonDev); cudaFreeHost(p->report); cudaEventDestroy(p->scanEvent);
	cudaFree(p->spanCountDev); cudaFree(p->longCountDev); cudaFree(p->longLatchedDev); cudaFree(p->bigListDev);
	p->spMarkAt.release(); p->spLongList.release(); p->spMarkZ.release(); p->spMarkV.release();
	for(auto& g : p->graphs) if(g.alive) cudaGraphExecDestroy(g.exec);
	if(p->peer.active && p->peer.rank != 0)
	{
		cudaIpcCloseMemHandle(p->peer.display0[0]); cudaIpcCloseMemHandle(p->peer.display0[1]); cudaIpcCloseMemHandle(p->peer.flags0);
	}
	cudaFree(p->peer.flagsOwn); cudaFree(p->peer.ctr);
	p->batchItems.release(); p->batchBlockDraw.release(); p->batchBlockList.release();
	p->spRec.release(); p->spTri.release(); p->tlFill.release(); p->tlLen.release(); p->tlIds.release();
	p->sv2Span.release(); p->sv2XY.release(); p->sv2Inv.release(); p->sv2Next.release(); p->sv2Chain.release(); p->sv2Colour.release();
	p->svTri.release(); p->svMisc.release(); p->svWinner.release(); p->svLeft.release(); p->svRight.release(); p->svInv.release();
	if(p->capDev) cudaFree(p->capDev);
	if(p->rcpDev) cudaFree(p->rcpDev);
	if(p->rsqrtDev) cudaFree(p->rsqrtDev);
	p->hdr.release(); p->vary.release(); p->triCount.release(); p->triOffset.release(); p->triRect.release(); p->scanSums.release();
	p->keysA.release(); p->valsA.release(); p->keysB.release(); p->valsB.release(); p->tileCount.release(); p->tileStart.release(); p->tileFill.release(); p->tileOrder.release(); p->sortCounts.release(); p->workList.release();
	for(auto& s : p->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
	for(auto& e : p->eventPool) cudaEventDestroy(e);
	cudaStreamDestroy(p->stream); cudaStreamDestroy(p->copyStream); cudaStreamDestroy(p->readStream);
	cudaEventDestroy(p->frameDone); cudaEventDestroy(p->readDone[0]); cudaEventDestroy(p->readDone[1]);
	delete p;
	return PS3D_OK;
}

// ---- textures (tex.cpp) ------------------------------------------------------------------------------------------

int ps3d_texture_create(ps3d_pipe* p, unsigned width, unsigned scanline, unsigned height, unsigned elemLen,
                        const void* pixels, int extraLayers, int wrapMode, int* idx)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(1 != elemLen && 4 != elemLen) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "PuresoftFBO: elemLen must be 1 or 4"); // fbo.cpp:21-24
	if(extraLayers < 0 || extraLayers > 5) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftFBO: extraLayers");
	if(!width || !height || scanline < width * elemLen) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "texture geometry");
	size_t slot = 0;
	for(; slot < p->textures.size(); slot++) if(!p->textures[slot]) break; // tex.cpp:6-16 first free slot
	Texture* t = new Texture();
	memset(t, 0, sizeof(*t));
	t->width = (int)width; t->height = (int)height; t->scanline = (int)scanline; t->elemLen = (int)elemLen; t->wrap = wrapMode; t->nLayers = 1 + extraLayers; t->filter = PS3D_FILTER_NEAREST;
	const size_t bytes = (size_t)scanline * height;
	for(int i = 0; i < t->nLayers; i++)
	{
		if(cudaMalloc((void**)&t->layer[i], bytes + 16) != cudaSuccess)
		{
			for(int j = 0; j < i; j++) cudaFree(t->layer[j]);
			delete t;
			return fail(p, PS3D_ERR_BAD_ALLOC, "texture");
		}
		cudaMemsetAsync(t->layer[i], 0, bytes, p->stream);
	}
	if(pixels) CK(p, cudaMemcpyAsync(t->layer[0], pixels, bytes, cudaMemcpyHostToDevice, p->stream)); // tex.cpp:20-23
	CK(p, cudaStreamSynchronize(p->stream));
	if(slot == p->textures.size()) p->textures.push_back(nullptr);
	p->textures[slot] = t;
	*idx = (int)slot;
	return PS3D_OK;
}

static Texture* texLayer(ps3d_pipe* p, int idx, int layer)
{
	if(idx < 0 || idx >= (int)p->textures.size() || !p->textures[idx]) return nullptr;
	if(layer < 0 || layer >= p->textures[idx]->nLayers) return nullptr;
	return p->textures[idx];
}
int ps3d_texture_upload(ps3d_pipe* p, int idx, int layer, const void* pixels)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	Texture* t = texLayer(p, idx, layer);
	if(!t) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index/layer out of range");
	CK(p, cudaMemcpyAsync(t->layer[layer], pixels, (size_t)t->scanline * t->height, cudaMemcpyHostToDevice, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_texture_download(ps3d_pipe* p, int idx, int layer, void* pixels)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	Texture* t = texLayer(p, idx, layer);
	if(!t) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index/layer out of range");
	CK(p, cudaMemcpyAsync(pixels, t->layer[layer], (size_t)t->scanline * t->height, cudaMemcpyDeviceToHost, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_texture_destroy(ps3d_pipe* p, int idx)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(idx < 0 || idx >= (int)p->textures.size()) return fail(p, PS3D_ERR_OUT_OF_RANGE, "destroyTexture: index out of range"); // tex.cpp:48-51
	if(p->textures[idx])
	{
		cudaStreamSynchronize(p->stream);
		for(int i = 0; i < 6; i++) if(p->textures[idx]->layer[i]) cudaFree(p->textures[idx]->layer[i]);
		delete p->textures[idx];
		p->textures[idx] = nullptr;
		if(p->depthTex == idx) p->depthTex = -1;
	}
	return PS3D_OK;
}

int ps3d_texture_set_filter(ps3d_pipe* p, int idx, int filter) // extension, include/ps3d.h
{
	TRACE();
	if(idx < 0 || idx >= (int)p->textures.size() || !p->textures[idx]) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index out of range");
	if(PS3D_FILTER_NEAREST != filter && PS3D_FILTER_BILINEAR != filter) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "filter");
	p->textures[idx]->filter = filter;   // latched per draw like every other texture property
	return PS3D_OK;
}

// ---- VBO / VAO (vbo.cpp, vao.cpp, pipeline.cpp:118-205) --------------------------------------------------------------

int ps3d_vbo_create(ps3d_pipe* p, size_t unitBytes, size_t unitCount, int* vbo)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	size_t slot = 0;
	for(; slot < p->vbos.size(); slot++) if(!p->vbos[slot].alive) break;
	Vbo v;
	v.unitBytes = unitBytes; v.unitCount = unitCount; v.alive = true; v.data = nullptr;
	v.ready = v.lastRead = nullptr; v.readyValid = v.readValid = false;
	if(cudaMalloc((void**)&v.data, unitBytes * unitCount + 64) != cudaSuccess) return fail(p, PS3D_ERR_BAD_ALLOC, "vbo"); // vbo.cpp:14
	cudaMemsetAsync(v.data, 0, unitBytes * unitCount + 64, p->stream);
	// the fill runs on the pipe's stream: an asynchronous upload (copy stream) must not overtake it
	{ const int rc = vboEvents(p, v); if(rc) { freeVbo(v); return rc; } }
	cudaEventRecord(v.lastRead, p->stream); v.readValid = true;
	if(slot == p->vbos.size()) p->vbos.push_back(v); else p->vbos[slot] = v;
	*vbo = (int)slot;
	return PS3D_OK;
}
static bool vboOk(ps3d_pipe* p, int v) { return v >= 0 && v < (int)p->vbos.size() && p->vbos[v].alive; }
int ps3d_vbo_update(ps3d_pipe* p, int vbo, const void* src)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	// vbo.cpp:28-31 copies synchronously: the caller may free `src` on return. (An asynchronous write still in flight on the copy
	// or gather stream lands first.)
	if(p->vbos[vbo].readyValid) CK(p, cudaStreamWaitEvent(p->stream, p->vbos[vbo].ready, 0));
	CK(p, cudaMemcpyAsync(p->vbos[vbo].data, src, p->vbos[vbo].unitBytes * p->vbos[vbo].unitCount, cudaMemcpyHostToDevice, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_vbo_update_device(ps3d_pipe* p, int vbo, const void* devSrc)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	if(p->vbos[vbo].readyValid) CK(p, cudaStreamWaitEvent(p->stream, p->vbos[vbo].ready, 0));
	CK(p, cudaMemcpyAsync(p->vbos[vbo].data, devSrc, p->vbos[vbo].unitBytes * p->vbos[vbo].unitCount, cudaMemcpyDeviceToDevice, p->stream));
	if(p->vbos[vbo].lastRead) { CK(p, cudaEventRecord(p->vbos[vbo].lastRead, p->stream)); p->vbos[vbo].readValid = true; }
	return PS3D_OK;
}
// Units [firstUnit, firstUnit + unitCount) from PINNED host memory, on the copy stream; returns at once. The copy waits for
// the last draw that read the VBO; every later draw that reads it waits for the copy. `pinnedSrc` must stay valid until
// the copy has run (ps3d_finish, or any synchronous call).
int ps3d_vbo_update_async(ps3d_pipe* p, int vbo, size_t firstUnit, size_t unitCount, const void* pinnedSrc)
{
	TRACE();
	cudaSetDevice(p->device);
	// a draw on the span path whose speculated sizes did not hold runs again from its geometry kernel, which reads the vertex
	// streams: its verdict (ready once its geometry + plan kernels are through — a fraction of the frame) is read first
	SETTLE(p);
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	Vbo& v = p->vbos[vbo];
	if(firstUnit > v.unitCount || unitCount > v.unitCount - firstUnit) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo range");
	{ const int rc = vboEvents(p, v); if(rc) return rc; }
	if(v.readValid) CK(p, cudaStreamWaitEvent(p->copyStream, v.lastRead, 0));
	if(unitCount) CK(p, cudaMemcpyAsync(v.data + firstUnit * v.unitBytes, pinnedSrc, unitCount * v.unitBytes, cudaMemcpyHostToDevice, p->copyStream));
	CK(p, cudaEventRecord(v.ready, p->copyStream));
	v.readyValid = true;
	return PS3D_OK;
}
// Somebody else (a collective, a peer copy) wrote the VBO's device memory on `cudaStream`: draws wait for what is enqueued there now.
int ps3d_vbo_device_written(ps3d_pipe* p, int vbo, void* cudaStream)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	Vbo& v = p->vbos[vbo];
	{ const int rc = vboEvents(p, v); if(rc) return rc; }
	CK(p, cudaEventRecord(v.ready, (cudaStream_t)cudaStream));
	v.readyValid = true;
	return PS3D_OK;
}
int ps3d_vbo_device_ptr(ps3d_pipe* p, int vbo, void** devPtr, size_t* bytes)
{
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	*devPtr = p->vbos[vbo].data; *bytes = p->vbos[vbo].unitBytes * p->vbos[vbo].unitCount;
	return PS3D_OK;
}
int ps3d_device_copy_stream(ps3d_pipe* p, void** s) { *s = (void*)p->copyStream; return PS3D_OK; }
// the pipe's stream waits for everything enqueued so far on the copy and read-back streams (e.g. before an event that closes a timed region)
int ps3d_device_join(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	cudaEvent_t e = takeEvent(p);
	CK(p, cudaEventRecord(e, p->copyStream)); CK(p, cudaStreamWaitEvent(p->stream, e, 0));
	CK(p, cudaEventRecord(e, p->readStream)); CK(p, cudaStreamWaitEvent(p->stream, e, 0));
	p->eventPool.push_back(e);
	return PS3D_OK;
}
int ps3d_vbo_destroy(ps3d_pipe* p, int vbo)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	cudaStreamSynchronize(p->stream);
	cudaStreamSynchronize(p->copyStream);
	if(p->gatherStream) cudaStreamSynchronize(p->gatherStream);
	freeVbo(p->vbos[vbo]);
	for(Vao& a : p->vaos) if(a.alive) for(int s = 0; s < PS3D_MAX_VBOS; s++) if(a.vbo[s] == vbo) a.vbo[s] = -1;
	return PS3D_OK;
}

int ps3d_vao_create(ps3d_pipe* p, int* vao)
{
	TRACE();
	size_t slot = 0;
	for(; slot < p->vaos.size(); slot++) if(!p->vaos[slot].alive) break; // pipeline.cpp:120-134
	Vao a;
	a.alive = true;
	for(int s = 0; s < PS3D_MAX_VBOS; s++) a.vbo[s] = -1;
	if(slot == p->vaos.size()) p->vaos.push_back(a); else p->vaos[slot] = a;
	*vao = (int)slot;
	return PS3D_OK;
}
static int vaoCheck(ps3d_pipe* p, int vao, int slot) // pipeline.cpp:140-148
{
	if(vao < 0 || vao >= (int)p->vaos.size() || !p->vaos[vao].alive) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO vao");
	if(slot < 0 || slot >= PS3D_MAX_VBOS) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO idx");
	return PS3D_OK;
}
int ps3d_vao_attach(ps3d_pipe* p, int vao, int slot, int vbo, int* displaced)
{
	TRACE();
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	int rc = vaoCheck(p, vao, slot); if(rc) return rc;
	if(displaced) *displaced = p->vaos[vao].vbo[slot];
	p->vaos[vao].vbo[slot] = vbo;
	return PS3D_OK;
}
int ps3d_vao_detach(ps3d_pipe* p, int vao, int slot, int* displaced)
{
	TRACE();
	int rc = vaoCheck(p, vao, slot); if(rc) return rc;
	if(displaced) *displaced = p->vaos[vao].vbo[slot];
	p->vaos[vao].vbo[slot] = -1;
	return PS3D_OK;
}
int ps3d_vao_get(ps3d_pipe* p, int vao, int slot, int* vbo)
{
	TRACE();
	int rc = vaoCheck(p, vao, slot); if(rc) return rc;
	*vbo = p->vaos[vao].vbo[slot];
	return PS3D_OK;
}
int ps3d_vao_destroy(ps3d_pipe* p, int vao)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(vao < 0 || vao >= (int)p->vaos.size()) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO vao"); // pipeline.cpp:185-188
	if(!p->vaos[vao].alive) return PS3D_OK;
	cudaStreamSynchronize(p->stream);
	// an asynchronous upload or all-gather may still be writing one of the VBOs about to be freed
	cudaStreamSynchronize(p->copyStream);
	if(p->gatherStream) cudaStreamSynchronize(p->gatherStream);
	for(int s = 0; s < PS3D_MAX_VBOS; s++) // pipeline.cpp:194-201: the pipeline owns attached VBOs
	{
		const int v = p->vaos[vao].vbo[s];
		if(v >= 0 && p->vbos[v].alive)
		{
			freeVbo(p->vbos[v]);
			// the handle may be reused: no other VAO may keep pointing at it (as ps3d_vbo_destroy does)
			for(Vao& a : p->vaos) if(a.alive) for(int t = 0; t < PS3D_MAX_VBOS; t++) if(a.vbo[t] == v) a.vbo[t] = -1;
		}
	}
	p->vaos[vao].alive = false;
	return PS3D_OK;
}

// ---- processors / programmes (prog.cpp) --------------------------------------------------------------------------------

int ps3d_processor_add(ps3d_pipe* p, int kind, int functor, int* idx)
{
	TRACE();
	if(kind < 0 || kind > 2) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "processor kind");
	if(!functorKnown(kind, functor)) return fail(p, PS3D_ERR_UNSUPPORTED, "no device functor with this id for this processor kind");
	size_t slot = 0;
	for(; slot < p->procs.size(); slot++) if(!p->procs[slot].alive) break; // prog.cpp:5-17
	Proc pr; pr.alive = true; pr.kind = kind; pr.functor = functor;
	if(slot == p->procs.size()) p->procs.push_back(pr); else p->procs[slot] = pr;
	*idx = (int)slot;
	return PS3D_OK;
}
int ps3d_processor_destroy(ps3d_pipe* p, int idx)
{
	TRACE();
	if(idx < 0 || idx >= (int)p->procs.size()) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::destroyProcessor"); // prog.cpp:24-27
	if(p->procs[idx].alive)
	{
		if(p->curProg >= 0 && (p->progs[p->curProg].vp == idx || p->progs[p->curProg].ip == idx || p->progs[p->curProg].fp == idx)) p->curProg = -1; // prog.cpp:32-37
		p->procs[idx].alive = false;
	}
	return PS3D_OK;
}
static const ProgEntry* findEntry(ps3d_pipe* p, const Prog& pg)
{
	for(const ProgEntry& e : programmeTable())
		if(e.fnV == p->procs[pg.vp].functor && e.fnI == p->procs[pg.ip].functor && e.fnF == p->procs[pg.fp].functor) return &e;
	return nullptr;
}
int ps3d_programme_create(ps3d_pipe* p, int vid, int iid, int fid, int* idx)
{
	TRACE();
	const int n = (int)p->procs.size();
	if(vid < 0 || vid >= n) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, vid"); // prog.cpp:45-58
	if(iid < 0 || iid >= n) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, iid");
	if(fid < 0 || fid >= n) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, fid");
	if(!p->procs[vid].alive || !p->procs[iid].alive || !p->procs[fid].alive || p->procs[vid].kind != PS3D_PROC_VERTEX ||
	   p->procs[iid].kind != PS3D_PROC_INTERPOLATION || p->procs[fid].kind != PS3D_PROC_FRAGMENT)
		return fail(p, PS3D_ERR_INVALID_ARGUMENT, "createProgramme: processor kind mismatch");
	Prog pg; pg.vp = vid; pg.ip = iid; pg.fp = fid;
	if(!findEntry(p, pg)) return fail(p, PS3D_ERR_UNSUPPORTED, "no kernel instantiation for this vertex/interpolation/fragment functor triple");
	size_t slot = 0;
	for(; slot < p->progs.size(); slot++) if(-1 == p->progs[slot].vp) break; // prog.cpp:60-71
	if(slot == p->progs.size()) p->progs.push_back(pg); else p->progs[slot] = pg;
	*idx = (int)slot;
	return PS3D_OK;
}
int ps3d_programme_destroy(ps3d_pipe* p, int idx)
{
	TRACE();
	if(idx < 0 || idx >= (int)p->progs.size()) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::destroyProgramme"); // prog.cpp:112-115
	p->progs[idx].vp = p->progs[idx].ip = p->progs[idx].fp = -1;
	if(p->curProg == idx) p->curProg = -1;
	return PS3D_OK;
}
int ps3d_programme_use(ps3d_pipe* p, int idx)
{
	TRACE();
	if(idx < 0 || idx >= (int)p->progs.size() || -1 == p->progs[idx].vp) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::useProgramme"); // prog.cpp:123-126
	p->curProg = idx;
	return PS3D_OK;
}

// ---- state (pipeline.cpp:207-342) ------------------------------------------------------------------------------------

int ps3d_set_viewport(ps3d_pipe* p, int width, int height)
{
	TRACE();
	if(width <= 0 || height <= 0 || width > 32767 || height > 32767) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "viewport");
	p->vpW = width; p->vpH = height;
	return PS3D_OK;
}
int ps3d_set_depth(ps3d_pipe* p, int textureIdx) // pipeline.cpp:218-239
{
	TRACE();
	if(-1 == textureIdx) { p->depthTex = -1; return PS3D_OK; }
	if(textureIdx < 0 || textureIdx >= (int)p->textures.size() || !p->textures[textureIdx]) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::setDepth");
	if(4 != p->textures[textureIdx]->elemLen || 0 != p->textures[textureIdx]->scanline % 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "PuresoftPipeline::setDepth");
	p->depthTex = textureIdx;
	return PS3D_OK;
}
int ps3d_set_uniform(ps3d_pipe* p, int idx, const void* data, size_t len) // pipeline.cpp:277-312
{
	TRACE();
	if(idx < 0 || idx >= PS3D_MAX_UNIFORMS) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::setUniform");
	if(!data) { p->uniforms[idx].clear(); p->uniformSet[idx] = false; return PS3D_OK; }
	if(p->uniforms[idx].size() < len) p->uniforms[idx].resize(len, 0);
	memcpy(p->uniforms[idx].data(), data, len);
	p->uniformSet[idx] = true;
	return PS3D_OK;
}
int ps3d_enable(ps3d_pipe* p, int bits) { p->behavior |= bits; return PS3D_OK; }
int ps3d_disable(ps3d_pipe* p, int bits) { p->behavior &= ~bits; return PS3D_OK; }

int ps3d_clear_depth(ps3d_pipe* p, float furthest) // pipeline.cpp:334-338 -> clear16 (fbo.cpp:348-371)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	TargetDesc d = depthTarget(p);
	const size_t quads = ((size_t)d.scanline * d.height) >> 4;
	const int blocks = (int)((quads + 255) / 256 < 148 * 16 ? (quads + 255) / 256 : 148 * 16);
	clear_depth_kernel<<<blocks ? blocks : 1, 256, 0, p->stream>>>((float4*)d.ptr, quads, furthest);
	p->launches++;
	CK(p, cudaGetLastError());
	return PS3D_OK;
}
int ps3d_clear_colour(ps3d_pipe* p, uint32_t bgra) // pipeline.cpp:340-343 -> clear4 skips the last buffer row (fbo.cpp:332-346)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(p->height < 2) return PS3D_OK;
	if(p->peer.active && p->peer.rank != 0) return PS3D_OK;   // sort-first over peer memory: the target is rank 0's, and so is its clear
	if(!p->capturing && p->readValid[p->back]) { CK(p, cudaStreamWaitEvent(p->stream, p->readDone[p->back], 0)); p->readValid[p->back] = false; }
	clear_colour_kernel<<<148 * 8, 256, 0, p->stream>>>(p->display[p->back], p->width, p->height - 1, p->width * 4, bgra);
	p->launches++;
	CK(p, cudaGetLastError());
	return peerFirstWrite(p);                          // rank 0 hands the cleared target out to the other ranks
}

// ---- the draw (drawvao.cpp:3-133) ------------------------------------------------------------------------------------

int ps3d_draw_vao(ps3d_pipe* p, int vao, int callerThread)
{
	TRACE();
	(void)callerThread;
	cudaSetDevice(p->device);
	if(p->curProg < 0 || vao < 0 || vao >= (int)p->vaos.size() || !p->vaos[vao].alive) return PS3D_OK; // drawvao.cpp:12-15
	const Prog pg = p->progs[p->curProg];
	if(pg.vp < 0) return PS3D_OK;
	const ProgEntry* pe = findEntry(p, pg);
	if(!pe) return fail(p, PS3D_ERR_UNSUPPORTED, "no kernel instantiation for this programme");

	DrawParams P;
	memset(&P, 0, sizeof(P));
	// vertex streams: all attached slots advance in lock-step; the draw ends when any runs out (vertthrd.cpp:21-31)
	const Vao& va = p->vaos[vao];
	size_t nverts = (size_t)-1;
	bool any = false;
	for(int s = 0; s < PS3D_MAX_VBOS; s++)
		if(va.vbo[s] >= 0)
		{
			const Vbo& v = p->vbos[va.vbo[s]];
			any = true;
			if(v.unitCount < nverts) nverts = v.unitCount;
			P.slot[s] = v.data;
			P.stride[s] = (uint32_t)v.unitBytes;
		}
	if(!any) return PS3D_OK;
	for(int s = 0; s < PS3D_MAX_VBOS; s++)
		if(((pe->slots >> s) & 1) && !P.slot[s]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "vertex functor reads a VBO slot that is not attached");
	// asynchronous uploads (ps3d_vbo_update_async / ps3d_vbo_device_written) of the streams this draw reads must have landed
	// (a captured frame waits for both in front of every launch instead: ps3d_graph_launch)
	if(!p->capturing)
	{
		for(int s = 0; s < PS3D_MAX_VBOS; s++)
			if(va.vbo[s] >= 0 && p->vbos[va.vbo[s]].readyValid) CK(p, cudaStreamWaitEvent(p->stream, p->vbos[va.vbo[s]].ready, 0));
		// ... and an asynchronous read-back of the target this draw writes must have left it
		if(p->readValid[p->back]) { CK(p, cudaStreamWaitEvent(p->stream, p->readDone[p->back], 0)); p->readValid[p->back] = false; }
	}
	const size_t ntris = nverts / 3;
	if(ntris > 0x7fffffffu / 3) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "too many triangles in one draw");
	// uniforms are latched now (the reference latches pointers in preprocess(), drawvao.cpp:18-20)
	for(int u = 0; u < PS_UNIFORM_SLOTS; u++)
	{
		if(((pe->uniforms >> u) & 1) && !p->uniformSet[u]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a uniform slot the programme reads is unset");
		if(p->uniformSet[u]) memcpy(P.u[u], p->uniforms[u].data(), p->uniforms[u].size() < 64 ? p->uniforms[u].size() : 64);
	}
	for(int k = 0; k < pe->ntex; k++)
	{
		const int us = pe->texSlot[k];
		int id = -1;
		if(p->uniforms[us].size() >= 4) memcpy(&id, p->uniforms[us].data(), 4);
		if(id < 0 || id >= (int)p->textures.size() || !p->textures[id]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a texture uniform does not name a live texture");
		const Texture* t = p->textures[id];
		if(4 != t->elemLen) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "sampler needs a 4-byte texture");
		for(int l = 0; l < 6; l++) P.tex[k].layer[l] = t->layer[l];
		P.tex[k].width = t->width; P.tex[k].height = t->height; P.tex[k].scanline = t->scanline; P.tex[k].wrap = t->wrap;
		P.tex[k].nLayers = t->nLayers; P.tex[k].elemLen = t->elemLen; P.tex[k].filter = t->filter;
	}
	p->stats.draws++;
	p->stats.triangles_submitted += ntris;
	if(0 == ntris) return PS3D_OK;

	P.ntris = (uint32_t)ntris;
	P.vpW = p->vpW; P.vpH = p->vpH; P.halfW = p->vpW / 2; P.halfH = p->vpH / 2; // rasterizer.cpp:28-31
	P.behavior = p->behavior;
	P.band0 = p->band0 < 0 ? 0 : p->band0;
	P.band1 = p->band1 > p->vpH ? p->vpH : p->band1;
	P.tilesX = (p->vpW + PS_TILE - 1) / PS_TILE;
	P.tilesY = (p->vpH + PS_TILE - 1) / PS_TILE;
	P.colour.ptr = (p->peer.active && p->peer.rank != 0) ? p->peer.display0[p->back] : p->display[p->back]; P.colour.width = p->width; P.colour.height = p->height; P.colour.scanline = p->width * 4; P.colour.topDown = 1;
	P.depth = depthTarget(p);
	P.approx = p->approx;
	P.stats = p->statsDev;
	P.cap = p->capDev; P.capW = p->capW; P.capH = p->capH;

	// which tile path (kernels.cuh): a functor that may discard() makes the depth write wait for the shading
	// (fragthrd.cpp:234-237) -> immediate; a draw that blends needs its colours applied in submission order -> ordered;
	// everything else -> split: raster + depth kernel, survivor stream, flat shade kernel — over span records computed once
	// by the geometry kernel (kernels_span.cuh, the default) or, PS3D_TILE_PATH=split, re-derived per tile (kernels.cuh)
	int path = pe->mayDiscard ? 0 : (((p->behavior & PS3D_BEHAVIOR_ALPHABLEND) && pe->usesWrite4) ? 1 : 2);
	// (a draw that blends takes the span path too — its colours resolved in submission order behind the shade kernel — unless
	// PS3D_SPAN_BLEND=0 keeps it on the one-kernel ordered tile path)
	static int spanBlendOn = -1;
	if(spanBlendOn < 0) { const char* e = getenv("PS3D_SPAN_BLEND"); spanBlendOn = (e && e[0] == '0') ? 0 : 1; }
	const bool blend = 1 == path && spanBlendOn && !p->peer.active;
	bool span = 2 == path || blend;
	if(tilePathForced() >= 0 && !(pe->mayDiscard)) { path = tilePathForced() == 2 && 1 == path ? 1 : tilePathForced(); span = false; }
	if(path != 0 && (P.vpW > 8191 || P.vpH > 8191)) { path = 1; span = false; }      // the survivor record packs x and y in 13 bits each
	if(radixBinningForced() || ntris >= PS_SPAN_MAX_TRIS) span = false;
	if(span && vao < (int)p->vaoLegacy.size() && p->vaoLegacy[vao]) span = false;
	// small draws of the span path are collected: consecutive ones into the same targets run as one batch (flushBatch)
	if(span && !blend && batchingOn() && p->speculate && !p->profiling && ntris <= PS_BATCH_DRAW_TRIS && 0 == P.band0 && P.band1 >= P.vpH)
	{
		if(!batchTakes(p, P)) { const int rc = flushBatch(p); if(rc) return rc; }
		ps3d_pipe::BatchDraw d;
		d.P = P; d.pe = pe; d.vao = vao; d.firstBlock = p->batch.blocks; d.trisPerBlock = batchTrisPerBlock(ntris); d.nBlocks = batchBlocks(ntris);
		p->batch.draws.push_back(d);
		p->batch.blocks += d.nBlocks;
		return PS3D_OK;
	}
	SETTLE(p);
	if(span) return enqueueSpan(p, P, pe, vao, 0, 0, 0, false, blend);
	return enqueueLegacy(p, P, pe, path, vao);
}

int ps3d_finish(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	CK(p, cudaStreamSynchronize(p->stream));
	CK(p, cudaStreamSynchronize(p->copyStream));
	CK(p, cudaStreamSynchronize(p->readStream));
	if(p->gatherStream) CK(p, cudaStreamSynchronize(p->gatherStream));
	if(p->graphLaunched && p->report->sticky)
	{
		p->report->sticky = 0; p->graphLaunched = false;
		return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a captured frame no longer fits the buffers it was captured with: run the frame normally once and re-capture");
	}
	p->graphLaunched = false;
	return PS3D_OK;
}
int ps3d_swap_buffers(ps3d_pipe* p) { p->back ^= 1; return PS3D_OK; } // pipeline.cpp:314-322

int ps3d_post_process(ps3d_pipe* p, int functor)   // post.cpp:3-19
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(functor != PS3D_POST_DEPTHOFFIELD) return fail(p, PS3D_ERR_UNSUPPORTED, "no device functor for this post-processor");
	if(p->readValid[p->back]) { CK(p, cudaStreamWaitEvent(p->stream, p->readDone[p->back], 0)); p->readValid[p->back] = false; }
	TargetDesc c;
	c.ptr = p->display[p->back]; c.width = p->width; c.height = p->height; c.scanline = p->width * 4; c.topDown = 1;
	const TargetDesc d = depthTarget(p);
	post_process_kernel<PostDepthofField><<<p->smCount * 8, 256, 0, p->stream>>>(c, d);
	p->launches++;
	CK(p, cudaGetLastError());
	return PS3D_OK;
}

int ps3d_read_colour(ps3d_pipe* p, void* bgra, size_t pitch)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	CK(p, cudaMemcpy2DAsync(bgra, pitch, p->display[p->back], (size_t)p->width * 4, (size_t)p->width * 4, p->height, cudaMemcpyDeviceToHost, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
// The colour target into PINNED host memory on the read-back stream, behind everything enqueued on the pipe's stream so far;
// returns at once. The image is complete after ps3d_finish. Later writes to the same target wait for the read-back.
int ps3d_read_colour_async(ps3d_pipe* p, void* pinnedBgra, size_t pitch)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	CK(p, cudaEventRecord(p->frameDone, p->stream));
	CK(p, cudaStreamWaitEvent(p->readStream, p->frameDone, 0));
	CK(p, cudaMemcpy2DAsync(pinnedBgra, pitch, p->display[p->back], (size_t)p->width * 4, (size_t)p->width * 4, p->height, cudaMemcpyDeviceToHost, p->readStream));
	CK(p, cudaEventRecord(p->readDone[p->back], p->readStream));
	p->readValid[p->back] = true;
	return PS3D_OK;
}
int ps3d_read_depth(ps3d_pipe* p, float* depth, size_t pitch)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	CK(p, cudaMemcpy2DAsync(depth, pitch, p->defaultDepth, p->depthScanline, (size_t)p->width * 4, p->height, cudaMemcpyDeviceToHost, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_write_colour(ps3d_pipe* p, const void* bgra, size_t pitch)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	CK(p, cudaMemcpy2DAsync(p->display[p->back], (size_t)p->width * 4, bgra, pitch, (size_t)p->width * 4, p->height, cudaMemcpyHostToDevice, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_write_depth(ps3d_pipe* p, const float* depth, size_t pitch)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	CK(p, cudaMemcpy2DAsync(p->defaultDepth, p->depthScanline, depth, pitch, (size_t)p->width * 4, p->height, cudaMemcpyHostToDevice, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}

int ps3d_get_stats(ps3d_pipe* p, ps3d_stats* out)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	DeviceStats d[PS_STATS_COPIES];                // per call: two pipes (two devices, two threads) must not share a host buffer
	CK(p, cudaMemcpyAsync(d, p->statsDev, sizeof(d), cudaMemcpyDeviceToHost, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	*out = p->stats;
	out->triangles_rasterised = out->spans = out->fragments_tested = out->fragments_shaded = 0;
	for(int i = 0; i < PS_STATS_COPIES; i++)
	{
		out->triangles_rasterised += d[i].triangles_rasterised;
		out->spans += d[i].spans;
		out->fragments_tested += d[i].fragments_tested;
		out->fragments_shaded += d[i].fragments_shaded;
	}
	return PS3D_OK;
}
int ps3d_reset_stats(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	memset(&p->stats, 0, sizeof(p->stats));
	CK(p, cudaMemsetAsync(p->statsDev, 0, sizeof(DeviceStats) * PS_STATS_COPIES, p->stream));
	return PS3D_OK;
}

int ps3d_debug_capture(ps3d_pipe* p, int width, int height)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(width < 0 || height < 0) return PS3D_ERR_INVALID_ARGUMENT;
	CK(p, cudaStreamSynchronize(p->stream));
	if(p->capDev) cudaFree(p->capDev);
	p->capDev = nullptr; p->capW = p->capH = 0;
	if(width > 0 && height > 0)
	{
		CK(p, cudaMalloc((void**)&p->capDev, (size_t)width * height * 4));
		CK(p, cudaMemsetAsync(p->capDev, 0, (size_t)width * height * 4, p->stream));
		p->capW = width; p->capH = height;
	}
	return PS3D_OK;
}
int ps3d_debug_read_shade_counts(ps3d_pipe* p, uint32_t* counts)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!p->capDev) return PS3D_ERR_INVALID_ARGUMENT;
	CK(p, cudaMemcpyAsync(counts, p->capDev, (size_t)p->capW * p->capH * 4, cudaMemcpyDeviceToHost, p->stream));
	CK(p, cudaStreamSynchronize(p->stream));
	return PS3D_OK;
}
int ps3d_debug_clear_shade_counts(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(p->capDev) CK(p, cudaMemsetAsync(p->capDev, 0, (size_t)p->capW * p->capH * 4, p->stream));
	return PS3D_OK;
}

// ---- sort-first exchange steps, issued natively on the pipe's streams ------------------------------------

int ps3d_comm_unique_id(void* id256)
{
	const NcclApi* a = ncclApi();
	if(!a || !id256) return PS3D_ERR_UNSUPPORTED;
	Ps3dNcclUniqueId ids[2];
	if(a->GetUniqueId(&ids[0]) != 0 || a->GetUniqueId(&ids[1]) != 0) return PS3D_ERR_DEVICE;
	memcpy(id256, ids, sizeof(ids));
	return PS3D_OK;
}
int ps3d_comm_init(ps3d_pipe* p, int rank, int world, const void* id256)
{
	TRACE();
	const NcclApi* a = ncclApi();
	if(!a) return fail(p, PS3D_ERR_UNSUPPORTED, "libnccl.so.2 not found");
	if(world < 1 || rank < 0 || rank >= world || !id256) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "rank / world");
	if(p->commFrame) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "communicators already initialised");
	cudaSetDevice(p->device);
	Ps3dNcclUniqueId ids[2];
	memcpy(ids, id256, sizeof(ids));
	NK(p, a->CommInitRank(&p->commFrame, world, ids[0], rank));
	NK(p, a->CommInitRank(&p->commUpload, world, ids[1], rank));
	CK(p, cudaStreamCreateWithFlags(&p->gatherStream, cudaStreamNonBlocking));
	p->commRank = rank; p->commWorld = world;
	return PS3D_OK;
}
int ps3d_comm_destroy(ps3d_pipe* p)
{
	// (only a pipe that has communicators touches the NCCL library: loading libnccl.so.2 into a process that imports torch
	// LATER would put the system's NCCL in front of the one torch bundles)
	const NcclApi* a = (p->commFrame || p->commUpload) ? ncclApi() : nullptr;
	if(p->gatherStream) { cudaStreamSynchronize(p->gatherStream); }
	if(a && p->commFrame) { cudaStreamSynchronize(p->stream); a->CommDestroy(p->commFrame); }
	if(a && p->commUpload) a->CommDestroy(p->commUpload);
	if(p->gatherStream) cudaStreamDestroy(p->gatherStream);
	p->commFrame = p->commUpload = nullptr; p->gatherStream = nullptr; p->commRank = 0; p->commWorld = 1;
	return PS3D_OK;
}
// bands: 2 * world ints, raster rows [row0, row1) rendered by each rank. The colour target is top-down (fbo.cpp:104-105):
// raster rows [r0, r1) are memory rows [H - r1, H - r0). One grouped send/recv on the pipe's stream behind the frame.
int ps3d_composite_bands(ps3d_pipe* p, const int* bands)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	const NcclApi* a = ncclApi();
	if(!a || !p->commFrame) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "ps3d_comm_init first");
	if(!bands) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "bands");
	if(1 == p->commWorld) return PS3D_OK;
	const size_t pitch = (size_t)p->width * 4;
	uint8_t* target = p->display[p->back];
	if(0 == p->commRank && p->readValid[p->back]) { CK(p, cudaStreamWaitEvent(p->stream, p->readDone[p->back], 0)); p->readValid[p->back] = false; }
	NK(p, a->GroupStart());
	for(int r = 1; r < p->commWorld; r++)
	{
		const int r0 = std::max(0, std::min(bands[2 * r], p->height)), r1 = std::max(r0, std::min(bands[2 * r + 1], p->height));
		if(r1 == r0) continue;
		uint8_t* at = target + (size_t)(p->height - r1) * pitch;
		const size_t bytes = (size_t)(r1 - r0) * pitch;
		if(0 == p->commRank) NK(p, a->Recv(at, bytes, PS_NCCL_UINT8, r, p->commFrame, p->stream));
		else if(r == p->commRank) NK(p, a->Send(at, bytes, PS_NCCL_UINT8, 0, p->commFrame, p->stream));
	}
	NK(p, a->GroupEnd());
	return PS3D_OK;           // (not counted in ps3d_device_launch_count: the kernel is NCCL's)
}
// Sharded upload's exchange: rank r has written units [r * per, (r + 1) * per) of the VBO (ps3d_vbo_update_async, per =
// unitCount / world); one in-place all-gather on the gather stream, behind that upload, makes them whole on every rank.
// Draws wait for it through the VBO's ready event.
int ps3d_vbo_all_gather(ps3d_pipe* p, int vbo)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	const NcclApi* a = ncclApi();
	if(!a || !p->commUpload) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "ps3d_comm_init first");
	if(!vboOk(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	Vbo& v = p->vbos[vbo];
	if(1 == p->commWorld) return PS3D_OK;
	{ const int rc = vboEvents(p, v); if(rc) return rc; }
	const size_t perBytes = (v.unitCount / (size_t)p->commWorld) * v.unitBytes;
	if(0 == perBytes) return PS3D_OK;
	// behind the asynchronous upload of this rank's shard (which itself waited for the last draw that read the VBO)
	if(v.readyValid) CK(p, cudaStreamWaitEvent(p->gatherStream, v.ready, 0));
	else if(v.readValid) CK(p, cudaStreamWaitEvent(p->gatherStream, v.lastRead, 0));
	NK(p, a->AllGather(v.data + (size_t)p->commRank * perBytes, v.data, perBytes, PS_NCCL_UINT8, p->commUpload, p->gatherStream));
	CK(p, cudaEventRecord(v.ready, p->gatherStream));
	v.readyValid = true;
	return PS3D_OK;
}

// ---- sort-first composite over NVLink peer memory (include/ps3d.h) ---------------------------------------------------------

int ps3d_peer_export(ps3d_pipe* p, void* blob)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!blob) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "blob");
	static_assert(3 * sizeof(cudaIpcMemHandle_t) <= PS3D_PEER_BLOB, "blob too small");
	cudaIpcMemHandle_t h[3];
	CK(p, cudaIpcGetMemHandle(&h[0], p->display[0]));
	CK(p, cudaIpcGetMemHandle(&h[1], p->display[1]));
	CK(p, cudaIpcGetMemHandle(&h[2], p->peer.flagsOwn));
	memset(blob, 0, PS3D_PEER_BLOB);
	memcpy(blob, h, sizeof(h));
	return PS3D_OK;
}
int ps3d_peer_import(ps3d_pipe* p, int rank, int world, const void* blobs)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(0 == world && !blobs)
	{
		// undo an import (a rank whose peers could not map rank 0's targets falls back to another composite with them)
		CK(p, cudaStreamSynchronize(p->stream));
		if(p->peer.active && p->peer.rank != 0)
		{
			cudaIpcCloseMemHandle(p->peer.display0[0]); cudaIpcCloseMemHandle(p->peer.display0[1]); cudaIpcCloseMemHandle(p->peer.flags0);
		}
		p->peer.active = false; p->peer.rank = 0; p->peer.world = 1; p->peer.needTake = false;
		return PS3D_OK;
	}
	if(world < 1 || world > 64 || rank < 0 || rank >= world || !blobs) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "rank / world");
	if(p->peer.active) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "peer targets already imported");
	CK(p, cudaStreamSynchronize(p->stream));
	if(0 == rank)
	{
		p->peer.display0[0] = p->display[0]; p->peer.display0[1] = p->display[1]; p->peer.flags0 = p->peer.flagsOwn;
	}
	else
	{
		cudaIpcMemHandle_t h[3];
		memcpy(h, blobs, sizeof(h));                   // rank 0's blob comes first
		void* m[3] = { nullptr, nullptr, nullptr };
		for(int i = 0; i < 3; i++)
			if(cudaIpcOpenMemHandle(&m[i], h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
			{
				const cudaError_t e = cudaGetLastError();
				for(int j = 0; j < i; j++) cudaIpcCloseMemHandle(m[j]);
				p->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) + " (no peer access to rank 0's GPU?)";
				return PS3D_ERR_DEVICE;
			}
		p->peer.display0[0] = (uint8_t*)m[0]; p->peer.display0[1] = (uint8_t*)m[1]; p->peer.flags0 = (PeerFlags*)m[2];
	}
	p->peer.rank = rank; p->peer.world = world;
	p->peer.active = world > 1;
	p->peer.needTake = true;
	return PS3D_OK;
}
int ps3d_composite_peer(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(!p->peer.active) return p->peer.world == 1 ? PS3D_OK : fail(p, PS3D_ERR_INVALID_ARGUMENT, "ps3d_peer_import first");
	// a frame that wrote no colour at all still takes part in the hand-over of the target
	{ const int rc = peerFirstWrite(p); if(rc) return rc; }
	if(0 == p->peer.rank) peer_wait_done_kernel<<<1, 64, 0, p->stream>>>(p->peer.ctr, p->peer.flags0, p->peer.world);
	else peer_signal_done_kernel<<<1, 1, 0, p->stream>>>(p->peer.ctr, p->peer.flags0, p->peer.rank);
	p->launches++;
	CK(p, cudaGetLastError());
	p->peer.needTake = true;
	return PS3D_OK;
}

// ---- captured frames (include/ps3d.h) ---------------------------------------------------------------------------------------

// The recorded kernels hold the addresses of the pipe's scratch buffers. A draw submitted normally AFTER the capture may outgrow
// one of them (freed, allocated anew): a replay would then write through a dangling pointer. The addresses are folded into one
// word at the end of the capture and compared at every launch.
static uint64_t scratchSignature(const ps3d_pipe* p)
{
	const void* ptrs[] = {
		p->hdr.p, p->vary.p, p->triCount.p, p->triOffset.p, p->triRect.p, p->scanSums.p, p->keysA.p, p->valsA.p, p->keysB.p, p->valsB.p,
		p->tileCount.p, p->tileStart.p, p->tileFill.p, p->tileOrder.p, p->sortCounts.p, p->workList.p,
		p->svTri.p, p->svMisc.p, p->svWinner.p, p->svLeft.p, p->svRight.p, p->svInv.p,
		p->spRec.p, p->spTri.p, p->tlFill.p, p->tlLen.p, p->tlIds.p, p->sv2Span.p, p->sv2XY.p, p->sv2Inv.p, p->sv2Next.p, p->sv2Chain.p, p->sv2Colour.p,
		p->spMarkAt.p, p->spLongList.p, p->spMarkZ.p, p->spMarkV.p, p->batchItems.p, p->batchBlockDraw.p, p->batchBlockList.p };
	uint64_t h = 1469598103934665603ull;
	for(const void* q : ptrs) { h ^= (uint64_t)(uintptr_t)q; h *= 1099511628211ull; }
	return h;
}

int ps3d_graph_begin(ps3d_pipe* p)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(p->capturing) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a frame is already being captured");
	if(p->profiling) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "per-kernel event timing is on: a captured frame has no such events");
	if(!p->speculate) return fail(p, PS3D_ERR_UNSUPPORTED, "PS3D_SPECULATE=0: a captured frame cannot size its buffers after a host sync");
	CK(p, cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
	p->capturing = true; g_capturing = true;
	p->capLaunches0 = p->launches; p->capDraws0 = p->stats.draws; p->capTris0 = p->stats.triangles_submitted;
	p->capBack = p->back;
	p->capVaos.clear();
	return PS3D_OK;
}
int ps3d_graph_end(ps3d_pipe* p, int* graph)
{
	TRACE();
	cudaSetDevice(p->device);
	if(!p->capturing) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "no frame is being captured");
	const int flushed = flushBatch(p);                 // small draws still being collected belong to the frame
	p->capturing = false; g_capturing = false;
	cudaGraph_t g = nullptr;
	const cudaError_t e = cudaStreamEndCapture(p->stream, &g);
	// the recorded calls did not run: what they counted belongs to the launches
	ps3d_pipe::Graph G;
	G.launches = p->launches - p->capLaunches0; G.draws = p->stats.draws - p->capDraws0; G.tris = p->stats.triangles_submitted - p->capTris0;
	p->launches = p->capLaunches0; p->stats.draws = p->capDraws0; p->stats.triangles_submitted = p->capTris0;
	G.back = p->capBack; G.backAfter = p->back;
	p->back = p->capBack;
	G.vaos = p->capVaos; G.alive = true; G.exec = nullptr; G.scratch = scratchSignature(p);
	if(flushed) { if(g) cudaGraphDestroy(g); cudaGetLastError(); return flushed; }
	if(e != cudaSuccess || !g) { cudaGetLastError(); p->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return PS3D_ERR_DEVICE; }
	const cudaError_t e2 = cudaGraphInstantiate(&G.exec, g, 0);
	cudaGraphDestroy(g);
	if(e2 != cudaSuccess) { p->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2); return PS3D_ERR_DEVICE; }
	size_t slot = 0;
	for(; slot < p->graphs.size(); slot++) if(!p->graphs[slot].alive) break;
	if(slot == p->graphs.size()) p->graphs.push_back(G); else p->graphs[slot] = G;
	if(graph) *graph = (int)slot;
	return PS3D_OK;
}
int ps3d_graph_launch(ps3d_pipe* p, int graph)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(graph < 0 || graph >= (int)p->graphs.size() || !p->graphs[graph].alive) return fail(p, PS3D_ERR_OUT_OF_RANGE, "graph");
	if(p->capturing) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a frame is being captured");
	ps3d_pipe::Graph& G = p->graphs[graph];
	if(p->back != G.back) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "the frame was captured with the other display target current");
	if(G.scratch != scratchSignature(p)) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a draw submitted since the capture outgrew a scratch buffer the captured frame points into: capture the frame again");
	// what the recorded calls would have waited for one by one: asynchronous uploads of the streams the frame reads, and an
	// asynchronous read-back still leaving the target it writes
	for(int vao : G.vaos)
		if(vao >= 0 && vao < (int)p->vaos.size() && p->vaos[vao].alive)
			for(int s = 0; s < PS3D_MAX_VBOS; s++)
			{
				const int v = p->vaos[vao].vbo[s];
				if(v >= 0 && p->vbos[v].alive && p->vbos[v].readyValid) CK(p, cudaStreamWaitEvent(p->stream, p->vbos[v].ready, 0));
			}
	if(p->readValid[p->back]) { CK(p, cudaStreamWaitEvent(p->stream, p->readDone[p->back], 0)); p->readValid[p->back] = false; }
	// (draws enqueued one by one are judged by settle(); their verdicts do not concern the captured frames)
	if(!p->graphLaunched) p->report->sticky = 0;
	CK(p, cudaGraphLaunch(G.exec, p->stream));
	p->graphLaunched = true;
	p->capturing = false;
	for(int vao : G.vaos) { const int rc = recordVboReads(p, vao); if(rc) return rc; }
	p->launches += G.launches; p->stats.draws += G.draws; p->stats.triangles_submitted += G.tris;
	p->back = G.backAfter;
	return PS3D_OK;
}
int ps3d_graph_destroy(ps3d_pipe* p, int graph)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	if(graph < 0 || graph >= (int)p->graphs.size() || !p->graphs[graph].alive) return fail(p, PS3D_ERR_OUT_OF_RANGE, "graph");
	CK(p, cudaStreamSynchronize(p->stream));
	cudaGraphExecDestroy(p->graphs[graph].exec);
	p->graphs[graph].alive = false;
	return PS3D_OK;
}

int ps3d_set_row_band(ps3d_pipe* p, int row0, int row1)
{
	TRACE();
	if(row0 == -1 && row1 == -1) { p->band0 = 0; p->band1 = 0x7fffffff; return PS3D_OK; }
	if(row0 < 0 || row1 < row0) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "row band");
	p->band0 = row0; p->band1 = row1;
	return PS3D_OK;
}

int ps3d_device_colour_ptr(ps3d_pipe* p, void** devPtr, size_t* pitch) { SETTLE(p); *devPtr = p->display[p->back]; *pitch = (size_t)p->width * 4; return PS3D_OK; }
int ps3d_device_depth_ptr(ps3d_pipe* p, void** devPtr, size_t* pitch) { SETTLE(p); *devPtr = p->defaultDepth; *pitch = (size_t)p->depthScanline; return PS3D_OK; }
int ps3d_device_stream(ps3d_pipe* p, void** s) { SETTLE(p); *s = (void*)p->stream; return PS3D_OK; }
int ps3d_device_launch_count(ps3d_pipe* p, uint64_t* n) { *n = p->launches; return PS3D_OK; }
int ps3d_debug_batch_counts(ps3d_pipe* p, uint64_t* b, uint64_t* d) { if(b) *b = p->batchesLaunched; if(d) *d = p->drawsBatched; return PS3D_OK; }

int ps3d_profile_enable(ps3d_pipe* p, int on)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	CK(p, cudaStreamSynchronize(p->stream));
	for(auto& s : p->spans) { p->eventPool.push_back(s.a); p->eventPool.push_back(s.b); }
	p->spans.clear();
	p->profiling = on != 0;
	p->profPairs = 0; p->profLaunches[0] = p->profLaunches[1] = p->profLaunches[2] = p->profLaunches[3] = 0;
	return PS3D_OK;
}
int ps3d_profile_read(ps3d_pipe* p, ps3d_profile* out)
{
	TRACE();
	cudaSetDevice(p->device);
	SETTLE(p);
	CK(p, cudaStreamSynchronize(p->stream));
	double ms[4] = { 0, 0, 0, 0 };
	for(auto& s : p->spans)
	{
		float t = 0;
		cudaEventElapsedTime(&t, s.a, s.b);
		ms[s.cls] += t;
		p->eventPool.push_back(s.a); p->eventPool.push_back(s.b);
	}
	p->spans.clear();
	out->geom_ms = ms[0]; out->bin_ms = ms[1]; out->tile_ms = ms[2]; out->shade_ms = ms[3];
	out->geom_launches = p->profLaunches[0]; out->bin_launches = p->profLaunches[1]; out->tile_launches = p->profLaunches[2];
	out->shade_launches = p->profLaunches[3];
	out->bin_pairs = p->profPairs;
	out->survivors = 0;
	p->profPairs = 0; p->profLaunches[0] = p->profLaunches[1] = p->profLaunches[2] = p->profLaunches[3] = 0;
	return PS3D_OK;
}
int ps3d_host_approx_info(int* rcpBits, int* rsqrtBits)
{
	const Ps3dHostApprox& a = hostApprox();
	*rcpBits = a.rcpBits; *rsqrtBits = a.rsqrtBits;
	return PS3D_OK;
}

} // extern "C"
