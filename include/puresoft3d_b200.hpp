// puresoft3d_b200.hpp — the thin C++ host layer over the C-ABI of include/ps3d.h.
//
// Header-only mirror of the reference's public class surface, so that the demo code of Puresoft3D
// (src/test/puresoft.cpp:113-206, src/test/scenobj.cpp:45-252, src/test2/loadscene.cpp:112-135) ports by changing one
// #include:
//     class PuresoftPipeline      src/puresoft3d/pipeline.h:24-66
//     class PuresoftVBO           src/puresoft3d/vbo.h:4-30
//     class PuresoftProcessor     src/puresoft3d/proc.h:8-71   (+ the DEF01..05 families of defproc.h)
//     PURESOFTIMGBUFF32 / PURESOFTBGRA   src/puresoft3d/defs.h:9-37
// Same member names, argument order, defaults, ownership and exception types (std::out_of_range,
// std::invalid_argument, std::bad_alloc — see the table in include/ps3d.h). What differs, and why:
//   * a processor object carries the id of a DEVICE functor (puresoft3d_b200/csrc/shaders.cuh) instead of virtual
//     process() bodies: a C++ virtual cannot be called from a CUDA kernel. The class names are kept.
//   * PuresoftVBO needs the pipeline that owns its device storage: `new PuresoftVBO(pipeline, unitBytes, unitCount)`.
//   * getTexture() cannot expose device storage as a host pointer; uploadTexture()/downloadTexture() copy a layer.
//   * drawVAO() enqueues on the pipe's CUDA stream; finish() (or any read-back) waits. swapBuffers() stays the
//     frame boundary.
//   * the presenter (PuresoftRenderer, rndr.h) is out of scope: readColour()/readDepth() return the targets.
// Link with puresoft3d_b200/libps3d_b200.so (sm_100a CUDA; no CPU fallback: the ctor throws without a GPU).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include "ps3d.h"

// ---- defs.h ------------------------------------------------------------------------------------------------------------
typedef struct
{
	union
	{
		struct { unsigned char bgra[4]; } ary;
		struct { unsigned char b, g, r, a; } elems;
		unsigned int i32;
	};
} PURESOFTBGRA;

typedef struct
{
	unsigned int width;
	unsigned int scanline;
	unsigned int height;
	unsigned int elemLen;
	void* pixels;
} PURESOFTIMGBUFF32;

const int BEHAVIOR_UPDATE_DEPTH = PS3D_BEHAVIOR_UPDATE_DEPTH; // pipeline.h:17-20
const int BEHAVIOR_TEST_DEPTH = PS3D_BEHAVIOR_TEST_DEPTH;
const int BEHAVIOR_FACE_CULLING = PS3D_BEHAVIOR_FACE_CULLING;
const int BEHAVIOR_ALPHABLEND = PS3D_BEHAVIOR_ALPHABLEND;

struct PuresoftFBO // only the two enums the pipeline API mentions (fbo.h:18-19)
{
	enum WRAPMODE { CLAMP = PS3D_WRAP_CLAMP, WRAP = PS3D_WRAP_WRAP };
	enum LAYER { LAYER_DEFAULT = 0, LAYER_XPOS = 0, LAYER_XNEG, LAYER_YPOS, LAYER_YNEG, LAYER_ZPOS, LAYER_ZNEG, LAYER_MAX };
};

// ---- proc.h: shader objects ---------------------------------------------------------------------------------------------
class PuresoftProcessor
{
public:
	PuresoftProcessor(int kind, int functor) : m_kind(kind), m_functor(functor) {}
	virtual ~PuresoftProcessor() {}
	int kind() const { return m_kind; }
	int functor() const { return m_functor; }
private:
	int m_kind, m_functor;
};
class PuresoftVertexProcessor : public PuresoftProcessor { public: explicit PuresoftVertexProcessor(int fn) : PuresoftProcessor(PS3D_PROC_VERTEX, fn) {} };
class PuresoftInterpolationProcessor : public PuresoftProcessor { public: explicit PuresoftInterpolationProcessor(int fn) : PuresoftProcessor(PS3D_PROC_INTERPOLATION, fn) {} };
class PuresoftFragmentProcessor : public PuresoftProcessor { public: explicit PuresoftFragmentProcessor(int fn) : PuresoftProcessor(PS3D_PROC_FRAGMENT, fn) {} };

#define PS3D_DECLARE_FAMILY(SUFFIX, FN) \
	class VertexProcesser##SUFFIX : public PuresoftVertexProcessor { public: VertexProcesser##SUFFIX() : PuresoftVertexProcessor(FN) {} }; \
	class InterpolationProcessor##SUFFIX : public PuresoftInterpolationProcessor { public: InterpolationProcessor##SUFFIX() : PuresoftInterpolationProcessor(FN) {} }; \
	class FragmentProcessor##SUFFIX : public PuresoftFragmentProcessor { public: FragmentProcessor##SUFFIX() : PuresoftFragmentProcessor(FN) {} };
PS3D_DECLARE_FAMILY(DEF01, PS3D_FN_DEF01) // tex1light1.h       (sic: the reference spells it "Processer")
PS3D_DECLARE_FAMILY(DEF02, PS3D_FN_DEF02) // colr1light1.h
PS3D_DECLARE_FAMILY(DEF03, PS3D_FN_DEF03) // tex1bump1light1.h
PS3D_DECLARE_FAMILY(DEF04, PS3D_FN_DEF04) // skybox.h
PS3D_DECLARE_FAMILY(DEF05, PS3D_FN_DEF05) // shadow.h
#undef PS3D_DECLARE_FAMILY

// The demos' own shader classes (src/test/testproc.h, src/test2/testproc.h) under their own names; the two demos reuse
// names (VP_Shadow, IP_Null, ...), hence one namespace per demo.
#define PS3D_DEMO_V(NAME, FN) class NAME : public PuresoftVertexProcessor { public: NAME() : PuresoftVertexProcessor(FN) {} };
#define PS3D_DEMO_I(NAME, FN) class NAME : public PuresoftInterpolationProcessor { public: NAME() : PuresoftInterpolationProcessor(FN) {} };
#define PS3D_DEMO_F(NAME, FN) class NAME : public PuresoftFragmentProcessor { public: NAME() : PuresoftFragmentProcessor(FN) {} };
namespace ps3d_demo1   // src/test/testproc.h
{
PS3D_DEMO_V(VP_Planet, PS3D_FN_PLANET) PS3D_DEMO_I(IP_Planet, PS3D_FN_PLANET) PS3D_DEMO_F(FP_Earth, PS3D_FN_PLANET) PS3D_DEMO_F(FP_Satellite, PS3D_FN_SATELLITE)
PS3D_DEMO_V(VP_Cloud, PS3D_FN_CLOUD) PS3D_DEMO_I(IP_Cloud, PS3D_FN_CLOUD) PS3D_DEMO_F(FP_Cloud, PS3D_FN_CLOUD)
PS3D_DEMO_V(VP_CloudShadow, PS3D_FN_CLOUDSHADOW) PS3D_DEMO_I(IP_CloudShadow, PS3D_FN_CLOUDSHADOW) PS3D_DEMO_F(FP_CloudShadow, PS3D_FN_CLOUDSHADOW)
}
namespace ps3d_demo2   // src/test2/testproc.h
{
PS3D_DEMO_V(VP_PositionOnly, PS3D_FN_POSITIONONLY) PS3D_DEMO_I(IP_Null, PS3D_FN_POSITIONONLY) PS3D_DEMO_F(FP_SingleColourNoLighting, PS3D_FN_POSITIONONLY)
PS3D_DEMO_V(VP_SingleColour, PS3D_FN_SINGLECOLOUR) PS3D_DEMO_I(IP_SingleColour, PS3D_FN_SINGLECOLOUR) PS3D_DEMO_F(FP_SingleColour, PS3D_FN_SINGLECOLOUR)
PS3D_DEMO_V(VP_DiffuseOnly, PS3D_FN_DIFFUSEONLY) PS3D_DEMO_I(IP_DiffuseOnly, PS3D_FN_DIFFUSEONLY) PS3D_DEMO_F(FP_DiffuseOnly, PS3D_FN_DIFFUSEONLY)
PS3D_DEMO_V(VP_Shadow, PS3D_FN_SHADOW2) PS3D_DEMO_F(FP_Null, PS3D_FN_SHADOW2)
}
#undef PS3D_DEMO_V
#undef PS3D_DEMO_I
#undef PS3D_DEMO_F

// proc.h:89-94. The reference's process(threadIndex, threadCount, frame, depth) body becomes a device functor named by id.
class PuresoftPostProcessor
{
public:
	explicit PuresoftPostProcessor(int functor) : m_functor(functor) {}
	virtual ~PuresoftPostProcessor() {}
	int functor() const { return m_functor; }
private:
	int m_functor;
};
class PP_DepthofField : public PuresoftPostProcessor { public: PP_DepthofField() : PuresoftPostProcessor(PS3D_POST_DEPTHOFFIELD) {} }; // src/test2/testpost.h

class PuresoftPipeline;

namespace ps3d_detail
{
inline void raise(ps3d_pipe* p, int rc)
{
	if(PS3D_OK == rc) return;
	const std::string msg = p ? ps3d_last_error(p) : "ps3d";
	switch(rc)
	{
	case PS3D_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
	case PS3D_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
	case PS3D_ERR_BAD_ALLOC: throw std::bad_alloc();
	default: throw std::runtime_error(msg + " (ps3d error " + std::to_string(rc) + ")");
	}
}
}

// ---- vbo.h ----------------------------------------------------------------------------------------------------------------
class PuresoftVBO
{
public:
	PuresoftVBO(PuresoftPipeline& pipeline, size_t unitBytes, size_t unitCount);
	~PuresoftVBO();
	void updateContent(const void* src);           // vbo.cpp:28-31 — whole-buffer, synchronous
	void updateContentDevice(const void* devSrc);  // source already in HBM
	// pipelined transfers (include/ps3d.h): units [firstUnit, firstUnit + unitCount) from PINNED host memory, returns at once
	void updateContentAsync(const void* pinnedSrc, size_t firstUnit, size_t unitCount);
	void allGather();                              // sharded upload's exchange step (after PuresoftPipeline::commInit)
	int handle() const { return m_handle; }
	size_t unitBytes() const { return m_unitBytes; }
	size_t unitCount() const { return m_unitCount; }
private:
	friend class PuresoftPipeline;
	ps3d_pipe* m_pipe;
	PuresoftPipeline* m_owner;                     // the pipeline that tracks this object; cleared when the pipeline dies first
	int m_handle;
	size_t m_unitBytes, m_unitCount;
	PuresoftVBO(const PuresoftVBO&);
	PuresoftVBO& operator=(const PuresoftVBO&);
};

// ---- pipeline.h -----------------------------------------------------------------------------------------------------------
class PuresoftPipeline
{
public:
	// pipeline.cpp:24-62. canvasWindow is kept for source compatibility and ignored (no presenter); `device` = CUDA ordinal.
	PuresoftPipeline(uintptr_t canvasWindow, int deviceWidth, int deviceHeight, void* rndr = NULL, int device = 0)
		: m_pipe(NULL), m_width(deviceWidth), m_height(deviceHeight)
	{
		(void)canvasWindow; (void)rndr;
		const int rc = ps3d_create(deviceWidth, deviceHeight, device, &m_pipe);
		if(PS3D_OK != rc) throw std::runtime_error(std::string("PuresoftPipeline: ps3d_create failed (") + std::to_string(rc) + "): " + ps3d_backend_name() + " needs a CUDA device; there is no CPU fallback");
	}
	~PuresoftPipeline()
	{
		// pipeline.cpp:64-116: the pipeline owns processors, textures, VAOs and the VBOs still attached to them
		for(size_t i = 0; i < m_procs.size(); i++) delete m_procs[i];
		for(size_t i = 0; i < m_vbos.size(); i++)
			if(m_vbos[i])
			{
				PuresoftVBO* v = m_vbos[i];
				m_vbos[i] = NULL;
				v->m_pipe = NULL; v->m_owner = NULL;   // a VBO the caller still holds outlives the pipeline as an empty shell
				if(m_owned[i]) delete v;
			}
		if(m_pipe) ps3d_destroy(m_pipe);
	}

	// texture api — tex.cpp:4-58
	int createTexture(const PURESOFTIMGBUFF32* image, int extraLayers = 0, PuresoftFBO::WRAPMODE mode = PuresoftFBO::CLAMP)
	{
		int idx = -1;
		check(ps3d_texture_create(m_pipe, image->width, image->scanline, image->height, image->elemLen, image->pixels, extraLayers, (int)mode, &idx));
		return idx;
	}
	void destroyTexture(int idx) { check(ps3d_texture_destroy(m_pipe, idx)); }
	// extension: bilinear filtering for PuresoftSampler2D reads of this texture (the reference is nearest-only)
	void setTextureFilter(int idx, bool bilinear) { check(ps3d_texture_set_filter(m_pipe, idx, bilinear ? PS3D_FILTER_BILINEAR : PS3D_FILTER_NEAREST)); }
	void uploadTexture(int idx, const void* pixels, PuresoftFBO::LAYER layer = PuresoftFBO::LAYER_DEFAULT) { check(ps3d_texture_upload(m_pipe, idx, (int)layer, pixels)); }
	void downloadTexture(int idx, void* pixels, PuresoftFBO::LAYER layer = PuresoftFBO::LAYER_DEFAULT) { check(ps3d_texture_download(m_pipe, idx, (int)layer, pixels)); }

	// processor api — prog.cpp:3-131. The pipeline takes ownership of `proc` (prog.cpp:5-17).
	int addProcessor(PuresoftProcessor* proc)
	{
		int idx = -1;
		const int rc = ps3d_processor_add(m_pipe, proc->kind(), proc->functor(), &idx);
		if(PS3D_OK != rc) { delete proc; check(rc); }
		if((size_t)idx >= m_procs.size()) m_procs.resize(idx + 1, NULL);
		delete m_procs[idx];
		m_procs[idx] = proc;
		return idx;
	}
	void destroyProcessor(int idx)
	{
		check(ps3d_processor_destroy(m_pipe, idx));
		if(idx >= 0 && (size_t)idx < m_procs.size()) { delete m_procs[idx]; m_procs[idx] = NULL; }
	}
	int createProgramme(int vid, int iid, int fid) { int idx = -1; check(ps3d_programme_create(m_pipe, vid, iid, fid, &idx)); return idx; }
	void destroyProgramme(int idx) { check(ps3d_programme_destroy(m_pipe, idx)); }
	void useProgramme(int idx) { check(ps3d_programme_use(m_pipe, idx)); }

	// vao api — pipeline.cpp:118-205. attachVBO returns the displaced VBO to the caller, who owns it again.
	int createVAO(void) { int idx = -1; check(ps3d_vao_create(m_pipe, &idx)); return idx; }
	PuresoftVBO* attachVBO(int vao, int idx, PuresoftVBO* vbo)
	{
		int old = -1;
		check(ps3d_vao_attach(m_pipe, vao, idx, vbo->handle(), &old));
		track(vbo, true);
		return release(old);
	}
	PuresoftVBO* detachVBO(int vao, int idx) { int old = -1; check(ps3d_vao_detach(m_pipe, vao, idx, &old)); return release(old); }
	PuresoftVBO* getVBO(int vao, int idx) { int cur = -1; check(ps3d_vao_get(m_pipe, vao, idx, &cur)); return cur >= 0 && (size_t)cur < m_vbos.size() ? m_vbos[cur] : NULL; }
	void destroyVAO(int vao)
	{
		// pipeline.cpp:194-201: attached VBOs die with the VAO
		std::vector<int> attached;
		for(int s = 0; s < PS3D_MAX_VBOS; s++) { int cur = -1; if(PS3D_OK == ps3d_vao_get(m_pipe, vao, s, &cur) && cur >= 0) attached.push_back(cur); }
		check(ps3d_vao_destroy(m_pipe, vao));
		for(size_t i = 0; i < attached.size(); i++)
		{
			const int h = attached[i];
			if((size_t)h < m_vbos.size() && m_vbos[h]) { PuresoftVBO* v = m_vbos[h]; m_vbos[h] = NULL; v->m_handle = -1; v->m_owner = NULL; if(m_owned[h]) delete v; }
		}
	}

	// rendering api — pipeline.cpp:207-342, drawvao.cpp:3-133
	void setViewport(int width, int height, uintptr_t canvasWindow = 0) { (void)canvasWindow; check(ps3d_set_viewport(m_pipe, width, height)); }
	void setDepth(int idx = -1) { check(ps3d_set_depth(m_pipe, idx)); }
	void setUniform(int idx, const void* data, size_t len) { check(ps3d_set_uniform(m_pipe, idx, data, len)); }
	void drawVAO(int vao, bool callerThrdForFragProc = false) { check(ps3d_draw_vao(m_pipe, vao, callerThrdForFragProc ? 1 : 0)); }
	void finish(void) { check(ps3d_finish(m_pipe)); }
	void swapBuffers(void) { check(ps3d_swap_buffers(m_pipe)); }
	void postProcess(PuresoftPostProcessor* processor) { check(ps3d_post_process(m_pipe, processor->functor())); } // post.cpp:3-19
	void enable(int behavior) { check(ps3d_enable(m_pipe, behavior)); }
	void disable(int behavior) { check(ps3d_disable(m_pipe, behavior)); }
	void clearDepth(float furthest = 1.0f) { check(ps3d_clear_depth(m_pipe, furthest)); }
	void clearColour(PURESOFTBGRA bkgnd = PURESOFTBGRA()) { check(ps3d_clear_colour(m_pipe, bkgnd.i32)); }

	// read-back (replaces the debug dumps of dbg.cpp) and counters
	void readColour(void* bgra, size_t pitchBytes) { check(ps3d_read_colour(m_pipe, bgra, pitchBytes)); }
	void readDepth(float* depth, size_t pitchBytes) { check(ps3d_read_depth(m_pipe, depth, pitchBytes)); }
	ps3d_stats getStats(void) { ps3d_stats s; check(ps3d_get_stats(m_pipe, &s)); return s; }
	void setRowBand(int row0 = -1, int row1 = -1) { check(ps3d_set_row_band(m_pipe, row0, row1)); }
	// asynchronous read-back into PINNED host memory (complete after finish()), and the sort-first exchange steps
	void readColourAsync(void* pinnedBgra, size_t pitchBytes) { check(ps3d_read_colour_async(m_pipe, pinnedBgra, pitchBytes)); }
	static void commUniqueId(void* id256) { ps3d_detail::raise(NULL, ps3d_comm_unique_id(id256)); }
	void commInit(int rank, int world, const void* id256) { check(ps3d_comm_init(m_pipe, rank, world, id256)); }
	void compositeBands(const int* bands) { check(ps3d_composite_bands(m_pipe, bands)); }
	// composite over NVLink peer memory: every rank renders straight into rank 0's colour target (include/ps3d.h)
	void peerExport(void* blob) { check(ps3d_peer_export(m_pipe, blob)); }
	void peerImport(int rank, int world, const void* blobs) { check(ps3d_peer_import(m_pipe, rank, world, blobs)); }
	void peerReset() { check(ps3d_peer_import(m_pipe, 0, 0, NULL)); }   // undo an import (fall back to another composite)
	void compositePeer() { check(ps3d_composite_peer(m_pipe)); }
	// captured frames: record the calls of one frame once, replay them as one launch (include/ps3d.h)
	void graphBegin() { check(ps3d_graph_begin(m_pipe)); }
	int graphEnd() { int g = -1; check(ps3d_graph_end(m_pipe, &g)); return g; }
	void graphLaunch(int graph) { check(ps3d_graph_launch(m_pipe, graph)); }
	void graphDestroy(int graph) { check(ps3d_graph_destroy(m_pipe, graph)); }
	// batches of small draws (include/ps3d.h): how many were launched, how many draws ran inside one
	void debugBatchCounts(uint64_t* batches, uint64_t* draws) { check(ps3d_debug_batch_counts(m_pipe, batches, draws)); }

	int deviceWidth() const { return m_width; }
	int deviceHeight() const { return m_height; }
	ps3d_pipe* handle() { return m_pipe; }

private:
	friend class PuresoftVBO;
	void check(int rc) { ps3d_detail::raise(m_pipe, rc); }
	void track(PuresoftVBO* v, bool owned)
	{
		const int h = v->handle();
		if((size_t)h >= m_vbos.size()) { m_vbos.resize(h + 1, NULL); m_owned.resize(h + 1, false); }
		// a handle slot the library reused: whoever sat here before is gone from the library's point of view
		if(m_vbos[h] && m_vbos[h] != v) { m_vbos[h]->m_handle = -1; m_vbos[h]->m_owner = NULL; }
		m_vbos[h] = v; m_owned[h] = owned;
	}
	void untrack(PuresoftVBO* v)
	{
		const int h = v->handle();
		if(h >= 0 && (size_t)h < m_vbos.size() && m_vbos[h] == v) { m_vbos[h] = NULL; m_owned[h] = false; }
	}
	PuresoftVBO* release(int h)
	{
		if(h < 0 || (size_t)h >= m_vbos.size()) return NULL;
		m_owned[h] = false; // back in the caller's hands
		return m_vbos[h];
	}
	ps3d_pipe* m_pipe;
	int m_width, m_height;
	std::vector<PuresoftProcessor*> m_procs;
	std::vector<PuresoftVBO*> m_vbos;
	std::vector<bool> m_owned;
	PuresoftPipeline(const PuresoftPipeline&);
	PuresoftPipeline& operator=(const PuresoftPipeline&);
};

inline PuresoftVBO::PuresoftVBO(PuresoftPipeline& pipeline, size_t unitBytes, size_t unitCount)
	: m_pipe(pipeline.m_pipe), m_owner(&pipeline), m_handle(-1), m_unitBytes(unitBytes), m_unitCount(unitCount)
{
	ps3d_detail::raise(m_pipe, ps3d_vbo_create(m_pipe, unitBytes, unitCount, &m_handle));
	pipeline.track(this, false);
}
inline PuresoftVBO::~PuresoftVBO()
{
	if(m_owner) m_owner->untrack(this);            // the pipeline must not keep (or later write through) a dangling pointer
	if(m_pipe && m_handle >= 0) ps3d_vbo_destroy(m_pipe, m_handle);
}
inline void PuresoftVBO::updateContent(const void* src) { ps3d_detail::raise(m_pipe, ps3d_vbo_update(m_pipe, m_handle, src)); }
inline void PuresoftVBO::updateContentAsync(const void* pinnedSrc, size_t firstUnit, size_t unitCount) { ps3d_detail::raise(m_pipe, ps3d_vbo_update_async(m_pipe, m_handle, firstUnit, unitCount, pinnedSrc)); }
inline void PuresoftVBO::allGather() { ps3d_detail::raise(m_pipe, ps3d_vbo_all_gather(m_pipe, m_handle)); }
inline void PuresoftVBO::updateContentDevice(const void* devSrc) { ps3d_detail::raise(m_pipe, ps3d_vbo_update_device(m_pipe, m_handle, devSrc)); }
