// include/ps3d.h implemented over the UNMODIFIED reference classes (PuresoftPipeline & co., compiled in place
// from /root/reference by build_ref.py). TEST INFRASTRUCTURE ONLY: this is how the parity tests and
// `bench.py --impl reference` drive the reference's own CPU renderer; the product never loads it.
//
// Harness-side additions (none changes what the reference computes):
//   * a headless PuresoftRenderer (two malloc'd BGRA buffers) handed to the ctor so the GDI+ fallback
//     (pipeline.cpp:255) is never taken;
//   * counting decorators around the three processors ("count it on the CPU by wrapping the FP
//     in a counting decorator") — they forward every call untouched;
//   * `#define private public` to read the default depth target back (pipeline.h:81 m_defaultDepth).
#include <vector>
#include <string>
#include <stdexcept>
#include <string.h>
#include <stdlib.h>
#include <assert.h>
#include "windows.h"
#define private public
#define protected public
#include "pipeline.h"
#include "defproc.h"
#undef private
#undef protected

#include "ps3d.h"

#ifdef PS3D_WITH_DEMO_SHADERS
PuresoftProcessor* ps3d_demo1_make_processor(int kind, int functor); // demo_procs1.cpp (src/test/testproc.cpp)
PuresoftProcessor* ps3d_demo2_make_processor(int kind, int functor); // demo_procs2.cpp (src/test2/testproc.cpp)
PuresoftPostProcessor* ps3d_demo2_make_post_processor(int functor); // demo_procs2.cpp (src/test2/testpost.cpp)
#endif

namespace
{

struct Counters
{
	struct __attribute__((aligned(64))) Slot { volatile uint64_t v; };
	Slot vertices[1];
	Slot spans[1];
	Slot tested[64];
	Slot shaded[64];
	uint64_t sum(const Slot* s, int n) const { uint64_t t = 0; for(int i = 0; i < n; i++) t += s[i].v; return t; }
	void reset() { memset(this, 0, sizeof(*this)); }
};

struct Capture
{
	int w, h;
	std::vector<uint32_t> counts;
	Capture() : w(0), h(0) {}
};

struct Shared
{
	Counters counters;
	Capture capture;
	bool counting;
};

class CountingVP : public PuresoftVertexProcessor
{
	PuresoftVertexProcessor* m_inner; Shared* m_sh;
public:
	CountingVP(PuresoftVertexProcessor* inner, Shared* sh) : m_inner(inner), m_sh(sh) {}
	~CountingVP() { delete m_inner; }
	size_t userDataBytes(void) const { return m_inner->userDataBytes(); }
	void preprocess(const PURESOFTUNIFORM* uniforms) { m_inner->preprocess(uniforms); }
	void process(const VertexProcessorInput* input, VertexProcessorOutput* output) const
	{
		m_sh->counters.vertices[0].v++; // caller thread only (vertthrd.cpp:11-12)
		m_inner->process(input, output);
	}
};

class CountingIP : public PuresoftInterpolationProcessor
{
	PuresoftInterpolationProcessor* m_inner; Shared* m_sh;
public:
	CountingIP(PuresoftInterpolationProcessor* inner, Shared* sh) : m_inner(inner), m_sh(sh) {}
	~CountingIP() { delete m_inner; }
	size_t userDataBytes(void) const { return m_inner->userDataBytes(); }
	void preprocess(const PURESOFTUNIFORM* uniforms) { m_inner->preprocess(uniforms); }
	void interpolateByContributes(void* o, const void** v, const float* c) const { m_inner->interpolateByContributes(o, v, c); }
	void calcStep(void* step, const void* start, const void* end, int n) const
	{
		m_sh->counters.spans[0].v++; // caller thread only (drawvao.cpp:96)
		m_inner->calcStep(step, start, end, n);
	}
	void correctInterpolation(void* o, const void* start, float cf2) const
	{
		// one call per visited pixel (interp.cpp:87); worker threads — spread over padded slots
		__sync_fetch_and_add(&m_sh->counters.tested[((uintptr_t)o >> 6) & 63].v, 1);
		m_inner->correctInterpolation(o, start, cf2);
	}
	void stepForward(void* start, const void* step, int n) const { m_inner->stepForward(start, step, n); }
};

class CountingFP : public PuresoftFragmentProcessor
{
	PuresoftFragmentProcessor* m_inner; Shared* m_sh;
public:
	CountingFP(PuresoftFragmentProcessor* inner, Shared* sh) : m_inner(inner), m_sh(sh) {}
	~CountingFP() { delete m_inner; }
	size_t userDataBytes(void) const { return m_inner->userDataBytes(); }
	void preprocess(const PURESOFTUNIFORM* uniforms, const void** textures) { m_inner->preprocess(uniforms, textures); }
	void process(const FragmentProcessorInput* input, FragmentProcessorOutput* output) const
	{
		int x = input->position[0], y = input->position[1];
		__sync_fetch_and_add(&m_sh->counters.shaded[y & 63].v, 1);
		Capture& cap = m_sh->capture;
		if(cap.w > 0 && x >= 0 && y >= 0 && x < cap.w && y < cap.h)
			__sync_fetch_and_add(&cap.counts[(size_t)y * cap.w + x], 1);
		m_inner->process(input, output);
	}
};

class HeadlessRenderer : public PuresoftRenderer
{
	int m_width, m_height;
	void* m_buffers[2];
	int m_back;
public:
	HeadlessRenderer(int w, int h) : m_width(w), m_height(h), m_back(0)
	{
		size_t bytes = (size_t)w * h * 4;
		m_buffers[0] = _aligned_malloc(bytes, 64);
		m_buffers[1] = _aligned_malloc(bytes, 64);
		memset(m_buffers[0], 0, bytes);
		memset(m_buffers[1], 0, bytes);
	}
	~HeadlessRenderer() { _aligned_free(m_buffers[0]); _aligned_free(m_buffers[1]); }
	void startup(uintptr_t, int, int) {}
	void shutdown(void) {}
	void setCanvas(uintptr_t) {}
	void getDesc(PURESOFTIMGBUFF32* desc)
	{
		desc->width = m_width; desc->scanline = m_width * 4; desc->height = m_height; desc->elemLen = 4; desc->pixels = NULL;
	}
	// same contract as rndrgdi.cpp:101-111: present the finished back buffer, hand out the other one
	void* swapBuffers(void) { m_back ^= 1; return m_buffers[m_back]; }
	void release(void) { delete this; }
	void* current(void) { return m_buffers[m_back]; }
};

} // namespace

struct ps3d_pipe
{
	PuresoftPipeline* pipe;
	HeadlessRenderer* rndr;
	Shared* sh;
	int width, height;
	std::vector<PuresoftVBO*> vbos;
	std::vector<size_t> vboBytes;
	std::string err;
	uint64_t draws;
	uint64_t triangles;   // complete triangles: per draw, vertex-processor calls / 3 (a dangling vertex or two never forms one)
};

#define PS3D_TRY(p) try {
#define PS3D_CATCH(p) } \
	catch(const std::out_of_range& e) { (p)->err = e.what(); return PS3D_ERR_OUT_OF_RANGE; } \
	catch(const std::invalid_argument& e) { (p)->err = e.what(); return PS3D_ERR_INVALID_ARGUMENT; } \
	catch(const std::bad_alloc& e) { (p)->err = e.what(); return PS3D_ERR_BAD_ALLOC; } \
	catch(const std::exception& e) { (p)->err = e.what(); return PS3D_ERR_INVALID_ARGUMENT; } \
	return PS3D_OK;

static int vboId(ps3d_pipe* p, PuresoftVBO* v)
{
	if(!v) return -1;
	for(size_t i = 0; i < p->vbos.size(); i++)
		if(p->vbos[i] == v) return (int)i;
	return -1;
}

extern "C" {

const char* ps3d_backend_name(void) { return "reference-shim"; }
const char* ps3d_last_error(const ps3d_pipe* p) { return p ? p->err.c_str() : ""; }

int ps3d_create(int width, int height, int, ps3d_pipe** out)
{
	if(!out || width <= 0 || height <= 0) return PS3D_ERR_INVALID_ARGUMENT;
	ps3d_pipe* p = new ps3d_pipe;
	p->width = width; p->height = height; p->draws = 0; p->triangles = 0;
	p->sh = new Shared;
	p->sh->counters.reset();
	const char* env = getenv("PS3D_REF_COUNTING");
	p->sh->counting = !(env && env[0] == '0');
	try
	{
		p->rndr = new HeadlessRenderer(width, height);
		p->pipe = new PuresoftPipeline(0, width, height, p->rndr);
	}
	catch(const std::exception& e)
	{
		delete p->sh; delete p;
		return PS3D_ERR_BAD_ALLOC;
	}
	*out = p;
	return PS3D_OK;
}

int ps3d_destroy(ps3d_pipe* p)
{
	if(!p) return PS3D_ERR_INVALID_ARGUMENT;
	// the pipeline deletes attached VBOs (pipeline.cpp:194-201); detached leftovers are ours
	std::vector<PuresoftVBO*> attached;
	for(size_t v = 0; v < p->pipe->m_vaoPool.size(); v++)
		if(p->pipe->m_vaoPool[v])
			for(size_t s = 0; s < MAX_VBOS; s++)
				if(p->pipe->m_vaoPool[v]->getVBO((unsigned)s)) attached.push_back(p->pipe->m_vaoPool[v]->getVBO((unsigned)s));
	for(size_t i = 0; i < p->vbos.size(); i++)
	{
		bool isAttached = false;
		for(size_t j = 0; j < attached.size(); j++) if(attached[j] == p->vbos[i]) isAttached = true;
		if(p->vbos[i] && !isAttached) delete p->vbos[i];
	}
	delete p->pipe; // releases the renderer (pipeline.cpp:113-114)
	delete p->sh;
	delete p;
	return PS3D_OK;
}

int ps3d_texture_create(ps3d_pipe* p, unsigned width, unsigned scanline, unsigned height, unsigned elemLen,
                        const void* pixels, int extraLayers, int wrapMode, int* idx)
{
	PS3D_TRY(p)
	PURESOFTIMGBUFF32 desc;
	desc.width = width; desc.scanline = scanline; desc.height = height; desc.elemLen = elemLen; desc.pixels = (void*)pixels;
	*idx = p->pipe->createTexture(&desc, extraLayers, wrapMode == PS3D_WRAP_WRAP ? PuresoftFBO::WRAP : PuresoftFBO::CLAMP);
	PS3D_CATCH(p)
}

static PuresoftFBO* layerOf(ps3d_pipe* p, int idx, int layer)
{
	if(idx < 0 || idx >= (int)p->pipe->m_texPool.size() || !p->pipe->m_texPool[idx])
		throw std::out_of_range("texture index");
	if(layer < 0 || layer >= PuresoftFBO::LAYER_MAX)
		throw std::out_of_range("texture layer");
	PuresoftFBO* fbo = p->pipe->m_texPool[idx]->getExtraLayer((PuresoftFBO::LAYER)layer);
	if(!fbo) throw std::out_of_range("texture layer not allocated");
	return fbo;
}

int ps3d_texture_upload(ps3d_pipe* p, int idx, int layer, const void* pixels)
{
	PS3D_TRY(p)
	PuresoftFBO* fbo = layerOf(p, idx, layer);
	memcpy(fbo->getBuffer(), pixels, fbo->getBytes());
	PS3D_CATCH(p)
}

int ps3d_texture_download(ps3d_pipe* p, int idx, int layer, void* pixels)
{
	PS3D_TRY(p)
	PuresoftFBO* fbo = layerOf(p, idx, layer);
	memcpy(pixels, fbo->getBuffer(), fbo->getBytes());
	PS3D_CATCH(p)
}

int ps3d_texture_set_filter(ps3d_pipe* p, int idx, int filter)
{
	// the reference's PuresoftSampler2D is nearest-only (samplr2d.cpp:19-25): nothing to switch
	(void)idx;
	if(PS3D_FILTER_NEAREST == filter) return PS3D_OK;
	p->err = "the reference has no bilinear sampler";
	return PS3D_ERR_UNSUPPORTED;
}

int ps3d_texture_destroy(ps3d_pipe* p, int idx)
{
	PS3D_TRY(p)
	p->pipe->destroyTexture(idx);
	PS3D_CATCH(p)
}

int ps3d_vbo_create(ps3d_pipe* p, size_t unitBytes, size_t unitCount, int* vbo)
{
	PS3D_TRY(p)
	PuresoftVBO* v = new PuresoftVBO(unitBytes, unitCount);
	size_t slot = 0;
	for(; slot < p->vbos.size(); slot++) if(!p->vbos[slot]) break;
	if(slot == p->vbos.size()) { p->vbos.push_back(NULL); p->vboBytes.push_back(0); }
	p->vbos[slot] = v;
	p->vboBytes[slot] = unitBytes * unitCount;
	*vbo = (int)slot;
	PS3D_CATCH(p)
}

int ps3d_vbo_update(ps3d_pipe* p, int vbo, const void* src)
{
	PS3D_TRY(p)
	if(vbo < 0 || vbo >= (int)p->vbos.size() || !p->vbos[vbo]) throw std::out_of_range("vbo");
	p->vbos[vbo]->updateContent(src);
	PS3D_CATCH(p)
}

int ps3d_vbo_destroy(ps3d_pipe* p, int vbo)
{
	PS3D_TRY(p)
	if(vbo < 0 || vbo >= (int)p->vbos.size() || !p->vbos[vbo]) throw std::out_of_range("vbo");
	delete p->vbos[vbo];
	p->vbos[vbo] = NULL;
	PS3D_CATCH(p)
}

int ps3d_vao_create(ps3d_pipe* p, int* vao)
{
	PS3D_TRY(p)
	*vao = p->pipe->createVAO();
	PS3D_CATCH(p)
}

int ps3d_vao_attach(ps3d_pipe* p, int vao, int slot, int vbo, int* displaced)
{
	PS3D_TRY(p)
	if(vbo < 0 || vbo >= (int)p->vbos.size() || !p->vbos[vbo]) throw std::out_of_range("vbo");
	if(vao >= 0 && vao < (int)p->pipe->m_vaoPool.size() && !p->pipe->m_vaoPool[vao]) throw std::out_of_range("vao");
	PuresoftVBO* old = p->pipe->attachVBO(vao, slot, p->vbos[vbo]);
	if(displaced) *displaced = vboId(p, old);
	PS3D_CATCH(p)
}

int ps3d_vao_detach(ps3d_pipe* p, int vao, int slot, int* displaced)
{
	PS3D_TRY(p)
	if(vao >= 0 && vao < (int)p->pipe->m_vaoPool.size() && !p->pipe->m_vaoPool[vao]) throw std::out_of_range("vao");
	PuresoftVBO* old = p->pipe->detachVBO(vao, slot);
	if(displaced) *displaced = vboId(p, old);
	PS3D_CATCH(p)
}

int ps3d_vao_get(ps3d_pipe* p, int vao, int slot, int* vbo)
{
	PS3D_TRY(p)
	if(vao >= 0 && vao < (int)p->pipe->m_vaoPool.size() && !p->pipe->m_vaoPool[vao]) throw std::out_of_range("vao");
	*vbo = vboId(p, p->pipe->getVBO(vao, slot));
	PS3D_CATCH(p)
}

int ps3d_vao_destroy(ps3d_pipe* p, int vao)
{
	PS3D_TRY(p)
	if(vao >= 0 && vao < (int)p->pipe->m_vaoPool.size() && p->pipe->m_vaoPool[vao])
	{
		for(size_t s = 0; s < MAX_VBOS; s++)
		{
			int id = vboId(p, p->pipe->m_vaoPool[vao]->getVBO((unsigned)s));
			if(id >= 0) p->vbos[id] = NULL; // destroyVAO deletes them (pipeline.cpp:194-201)
		}
	}
	p->pipe->destroyVAO(vao);
	PS3D_CATCH(p)
}

static PuresoftProcessor* makeProcessor(int kind, int functor)
{
	switch(functor)
	{
	case PS3D_FN_DEF01:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VertexProcesserDEF01 : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new InterpolationProcessorDEF01 : (PuresoftProcessor*)new FragmentProcessorDEF01;
	case PS3D_FN_DEF02:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VertexProcesserDEF02 : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new InterpolationProcessorDEF02 : (PuresoftProcessor*)new FragmentProcessorDEF02;
	case PS3D_FN_DEF03:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VertexProcesserDEF03 : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new InterpolationProcessorDEF03 : (PuresoftProcessor*)new FragmentProcessorDEF03;
	case PS3D_FN_DEF04:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VertexProcesserDEF04 : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new InterpolationProcessorDEF04 : (PuresoftProcessor*)new FragmentProcessorDEF04;
	case PS3D_FN_DEF05:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VertexProcesserDEF05 : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new InterpolationProcessorDEF05 : (PuresoftProcessor*)new FragmentProcessorDEF05;
	default:
#ifdef PS3D_WITH_DEMO_SHADERS
		return functor < 32 ? ps3d_demo1_make_processor(kind, functor) : ps3d_demo2_make_processor(kind, functor);
#else
		return NULL;
#endif
	}
}

int ps3d_processor_add(ps3d_pipe* p, int kind, int functor, int* idx)
{
	PS3D_TRY(p)
	if(kind < 0 || kind > 2) throw std::invalid_argument("processor kind");
	PuresoftProcessor* proc = makeProcessor(kind, functor);
	if(!proc) { p->err = "no such functor in the reference build"; return PS3D_ERR_UNSUPPORTED; }
	if(p->sh->counting)
	{
		if(kind == PS3D_PROC_VERTEX) proc = new CountingVP(dynamic_cast<PuresoftVertexProcessor*>(proc), p->sh);
		else if(kind == PS3D_PROC_INTERPOLATION) proc = new CountingIP(dynamic_cast<PuresoftInterpolationProcessor*>(proc), p->sh);
		else proc = new CountingFP(dynamic_cast<PuresoftFragmentProcessor*>(proc), p->sh);
	}
	*idx = p->pipe->addProcessor(proc);
	PS3D_CATCH(p)
}

int ps3d_processor_destroy(ps3d_pipe* p, int idx)
{
	PS3D_TRY(p)
	p->pipe->destroyProcessor(idx);
	PS3D_CATCH(p)
}

int ps3d_programme_create(ps3d_pipe* p, int vid, int iid, int fid, int* idx)
{
	PS3D_TRY(p)
	// the reference dereferences NULL when a slot index is valid but empty or of the wrong kind (prog.cpp:73)
	PuresoftPipeline::PROCCOLL& procs = p->pipe->m_processors;
	if(vid >= 0 && vid < (int)procs.size() && iid >= 0 && iid < (int)procs.size() && fid >= 0 && fid < (int)procs.size())
	{
		if(!dynamic_cast<PuresoftVertexProcessor*>(procs[vid]) || !dynamic_cast<PuresoftInterpolationProcessor*>(procs[iid]) ||
		   !dynamic_cast<PuresoftFragmentProcessor*>(procs[fid]))
			throw std::invalid_argument("createProgramme: processor kind mismatch");
	}
	*idx = p->pipe->createProgramme(vid, iid, fid);
	PS3D_CATCH(p)
}

int ps3d_programme_destroy(ps3d_pipe* p, int idx)
{
	PS3D_TRY(p)
	p->pipe->destroyProgramme(idx);
	PS3D_CATCH(p)
}

int ps3d_programme_use(ps3d_pipe* p, int idx)
{
	PS3D_TRY(p)
	if(idx >= 0 && idx < (int)p->pipe->m_programmes.size() && -1 == p->pipe->m_programmes[idx].vp)
		throw std::out_of_range("useProgramme: destroyed programme");
	p->pipe->useProgramme(idx);
	PS3D_CATCH(p)
}

int ps3d_set_viewport(ps3d_pipe* p, int width, int height)
{
	PS3D_TRY(p)
	if(width <= 0 || height <= 0) throw std::invalid_argument("viewport");
	p->pipe->setViewport(width, height);
	PS3D_CATCH(p)
}

int ps3d_set_depth(ps3d_pipe* p, int textureIdx)
{
	PS3D_TRY(p)
	if(textureIdx >= 0 && textureIdx < (int)p->pipe->m_texPool.size() && !p->pipe->m_texPool[textureIdx])
		throw std::out_of_range("setDepth: destroyed texture");
	p->pipe->setDepth(textureIdx);
	PS3D_CATCH(p)
}

int ps3d_set_uniform(ps3d_pipe* p, int idx, const void* data, size_t len)
{
	PS3D_TRY(p)
	p->pipe->setUniform(idx, data, len);
	PS3D_CATCH(p)
}

int ps3d_enable(ps3d_pipe* p, int bits) { p->pipe->enable(bits); return PS3D_OK; }
int ps3d_disable(ps3d_pipe* p, int bits) { p->pipe->disable(bits); return PS3D_OK; }

int ps3d_clear_depth(ps3d_pipe* p, float furthest)
{
	PS3D_TRY(p)
	p->pipe->clearDepth(furthest);
	PS3D_CATCH(p)
}

int ps3d_clear_colour(ps3d_pipe* p, uint32_t bgra)
{
	PS3D_TRY(p)
	PURESOFTBGRA c; c.i32 = bgra;
	p->pipe->clearColour(c);
	PS3D_CATCH(p)
}

int ps3d_draw_vao(ps3d_pipe* p, int vao, int callerThread)
{
	PS3D_TRY(p)
	if(vao >= 0 && vao < (int)p->pipe->m_vaoPool.size() && !p->pipe->m_vaoPool[vao]) return PS3D_OK;
	const uint64_t before = p->sh->counters.vertices[0].v;
	p->pipe->drawVAO(vao, callerThread != 0);
	p->triangles += (p->sh->counters.vertices[0].v - before) / 3;
	p->draws++;
	PS3D_CATCH(p)
}

int ps3d_finish(ps3d_pipe*) { return PS3D_OK; } // drawVAO is synchronous (drawvao.cpp:122-132)

int ps3d_swap_buffers(ps3d_pipe* p)
{
	PS3D_TRY(p)
	p->pipe->swapBuffers();
	PS3D_CATCH(p)
}

// the reference's own PP_DepthofField (src/test2/testpost.cpp, asm -> intrinsics by build_ref.py) through its own postProcess (post.cpp)
#ifdef PS3D_WITH_DEMO_SHADERS
int ps3d_post_process(ps3d_pipe* p, int functor)
{
	PS3D_TRY(p)
	PuresoftPostProcessor* pp = ps3d_demo2_make_post_processor(functor);
	if(!pp) throw std::invalid_argument("post-processor");
	p->pipe->postProcess(pp);
	delete pp;
	PS3D_CATCH(p)
}
#else
int ps3d_post_process(ps3d_pipe*, int) { return PS3D_ERR_UNSUPPORTED; }
#endif

int ps3d_read_colour(ps3d_pipe* p, void* bgra, size_t pitchBytes)
{
	PS3D_TRY(p)
	if(pitchBytes < (size_t)p->width * 4) throw std::invalid_argument("pitch");
	const PuresoftFBO* d = p->pipe->m_display;
	for(int r = 0; r < p->height; r++)
		memcpy((char*)bgra + (size_t)r * pitchBytes, (const char*)d->getBuffer() + (size_t)r * d->getScanline(), (size_t)p->width * 4);
	PS3D_CATCH(p)
}

int ps3d_read_depth(ps3d_pipe* p, float* depth, size_t pitchBytes)
{
	PS3D_TRY(p)
	if(pitchBytes < (size_t)p->width * 4) throw std::invalid_argument("pitch");
	const PuresoftFBO* d = &p->pipe->m_defaultDepth;
	for(int r = 0; r < p->height; r++)
		memcpy((char*)depth + (size_t)r * pitchBytes, (const char*)d->getBuffer() + (size_t)r * d->getScanline(), (size_t)p->width * 4);
	PS3D_CATCH(p)
}

int ps3d_write_colour(ps3d_pipe* p, const void* bgra, size_t pitchBytes)
{
	PS3D_TRY(p)
	if(pitchBytes < (size_t)p->width * 4) throw std::invalid_argument("pitch");
	PuresoftFBO* d = p->pipe->m_display;
	for(int r = 0; r < p->height; r++)
		memcpy((char*)d->getBuffer() + (size_t)r * d->getScanline(), (const char*)bgra + (size_t)r * pitchBytes, (size_t)p->width * 4);
	PS3D_CATCH(p)
}

int ps3d_write_depth(ps3d_pipe* p, const float* depth, size_t pitchBytes)
{
	PS3D_TRY(p)
	if(pitchBytes < (size_t)p->width * 4) throw std::invalid_argument("pitch");
	PuresoftFBO* d = &p->pipe->m_defaultDepth;
	for(int r = 0; r < p->height; r++)
		memcpy((char*)d->getBuffer() + (size_t)r * d->getScanline(), (const char*)depth + (size_t)r * pitchBytes, (size_t)p->width * 4);
	PS3D_CATCH(p)
}

int ps3d_get_stats(ps3d_pipe* p, ps3d_stats* out)
{
	memset(out, 0, sizeof(*out));
	const Counters& c = p->sh->counters;
	out->triangles_submitted = p->triangles;
	out->triangles_rasterised = 0; // not observable from outside the reference
	out->spans = c.spans[0].v;
	out->fragments_tested = c.sum(c.tested, 64);
	out->fragments_shaded = c.sum(c.shaded, 64);
	out->draws = p->draws;
	return PS3D_OK;
}

int ps3d_reset_stats(ps3d_pipe* p) { p->sh->counters.reset(); p->draws = 0; p->triangles = 0; return PS3D_OK; }

int ps3d_debug_capture(ps3d_pipe* p, int width, int height)
{
	if(width < 0 || height < 0) return PS3D_ERR_INVALID_ARGUMENT;
	if(!p->sh->counting && width > 0) { p->err = "PS3D_REF_COUNTING=0"; return PS3D_ERR_UNSUPPORTED; }
	p->sh->capture.w = 0;
	p->sh->capture.counts.assign((size_t)width * height, 0);
	p->sh->capture.h = height;
	p->sh->capture.w = width;
	return PS3D_OK;
}

int ps3d_debug_read_shade_counts(ps3d_pipe* p, uint32_t* counts)
{
	Capture& cap = p->sh->capture;
	if(cap.w <= 0) return PS3D_ERR_INVALID_ARGUMENT;
	memcpy(counts, &cap.counts[0], cap.counts.size() * sizeof(uint32_t));
	return PS3D_OK;
}

int ps3d_debug_clear_shade_counts(ps3d_pipe* p)
{
	Capture& cap = p->sh->capture;
	if(cap.w > 0) memset(&cap.counts[0], 0, cap.counts.size() * sizeof(uint32_t));
	return PS3D_OK;
}

int ps3d_set_row_band(ps3d_pipe* p, int row0, int row1)
{
	if(row0 == -1 && row1 == -1) return PS3D_OK;
	p->err = "row bands are a sort-first extension of the CUDA library";
	return PS3D_ERR_UNSUPPORTED;
}

int ps3d_device_colour_ptr(ps3d_pipe*, void**, size_t*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_depth_ptr(ps3d_pipe*, void**, size_t*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_stream(ps3d_pipe*, void**) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_update_device(ps3d_pipe*, int, const void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_update_async(ps3d_pipe*, int, size_t, size_t, const void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_device_ptr(ps3d_pipe*, int, void**, size_t*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_device_written(ps3d_pipe*, int, void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_copy_stream(ps3d_pipe*, void**) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_read_colour_async(ps3d_pipe*, void*, size_t) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_join(ps3d_pipe*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_unique_id(void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_init(ps3d_pipe*, int, int, const void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_destroy(ps3d_pipe*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_composite_bands(ps3d_pipe*, const int*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_peer_export(ps3d_pipe*, void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_peer_import(ps3d_pipe*, int, int, const void*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_composite_peer(ps3d_pipe*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_begin(ps3d_pipe*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_end(ps3d_pipe*, int*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_launch(ps3d_pipe*, int) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_destroy(ps3d_pipe*, int) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_all_gather(ps3d_pipe*, int) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_launch_count(ps3d_pipe*, uint64_t* n) { if(n) *n = 0; return PS3D_OK; }
int ps3d_debug_batch_counts(ps3d_pipe*, uint64_t* b, uint64_t* d) { if(b) *b = 0; if(d) *d = 0; return PS3D_OK; }
int ps3d_profile_enable(ps3d_pipe*, int) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_profile_read(ps3d_pipe*, ps3d_profile*) { return PS3D_ERR_UNSUPPORTED; }
int ps3d_host_approx_info(int* rcpBits, int* rsqrtBits) { *rcpBits = -1; *rsqrtBits = -1; return PS3D_OK; } // the hardware instructions themselves

} // extern "C"

// ---- function-level test hooks (NOT part of include/ps3d.h; tests/cpp/functors_vs_reference.cpp binds them) ---------------
// The reference's own processor classes and PuresoftFBO driven one call at a time, in plain C types, so that the product's
// device functors — compiled for the host by the test — can be compared with them input by input.
namespace
{
struct CaptureOutput : public FragmentProcessorOutput          // proc.h:51-63; FBOBridge: write never blends, write4 does (fragthrd.cpp:54-82)
{
	uint32_t bgra; int flags;                                    // bit 0 wrote, bit 1 discarded, bit 2 through write4
	CaptureOutput() : bgra(0), flags(0) {}
	void discard(void) { flags |= 2; }
	void read(int, void* d, size_t n) { memset(d, 0, n); }
	void read1(int, void* d) { memset(d, 0, 1); }
	void read4(int, void* d) { memset(d, 0, 4); }
	void read16(int, void* d) { memset(d, 0, 16); }
	void write(int, const void* d, size_t n) { memcpy(&bgra, d, n < 4 ? n : 4); flags |= 1; }
	void write1(int, const void* d) { memcpy(&bgra, d, 1); flags |= 1; }
	void write4(int, const void* d) { memcpy(&bgra, d, 4); flags |= 1 | 4; }
	void write16(int, const void* d) { memcpy(&bgra, d, 4); flags |= 1; }
};
struct RefProc
{
	PuresoftProcessor* proc;
	int kind;
	std::vector<PURESOFTUNIFORM> uniforms;
	std::vector<std::vector<unsigned char> > store;
	const void* textures[MAX_TEXTURES];
	RefProc() : proc(NULL), kind(0), uniforms(MAX_UNIFORMS), store(MAX_UNIFORMS) { memset(&uniforms[0], 0, sizeof(PURESOFTUNIFORM) * MAX_UNIFORMS); memset(textures, 0, sizeof(textures)); }
};
}
extern "C" {
void* ps3d_ref_fbo_create(int width, int height, int elemLen, int wrapMode, int layers, const void* const* pixels)
{
	PuresoftFBO* f = new PuresoftFBO(width, width * elemLen, height, elemLen, false, NULL, (PuresoftFBO::WRAPMODE)wrapMode, layers - 1);
	for(int l = 0; l < layers; l++)
	{
		PuresoftFBO* layer = 0 == l ? f : f->getExtraLayer((PuresoftFBO::LAYER)l);   // tex.cpp: extra layers are FBOs of their own
		if(layer && pixels && pixels[l]) memcpy(layer->getBuffer(), pixels[l], (size_t)width * elemLen * height);
	}
	return f;
}
void ps3d_ref_fbo_destroy(void* f) { delete (PuresoftFBO*)f; }
void* ps3d_ref_proc_create(int kind, int functor)
{
	PuresoftProcessor* pr = makeProcessor(kind, functor);
	if(!pr) return NULL;
	RefProc* r = new RefProc(); r->proc = pr; r->kind = kind;
	return r;
}
void ps3d_ref_proc_destroy(void* h) { RefProc* r = (RefProc*)h; delete r->proc; delete r; }
size_t ps3d_ref_proc_user_bytes(void* h) { return ((RefProc*)h)->proc->userDataBytes(); }
void ps3d_ref_proc_set_uniform(void* h, int slot, const void* data, size_t len)
{
	RefProc* r = (RefProc*)h;
	r->store[slot].assign((const unsigned char*)data, (const unsigned char*)data + len);
	r->store[slot].resize(len + 64);                              // room for the aligned copy below
	unsigned char* at = &r->store[slot][0];
	at += (16 - ((uintptr_t)at & 15)) & 15;                       // mcemath loads uniforms with movaps
	memmove(at, data, len);
	r->uniforms[slot].data = at; r->uniforms[slot].capacity = len;
}
void ps3d_ref_proc_set_texture(void* h, int index, void* fbo) { ((RefProc*)h)->textures[index] = fbo; }
void ps3d_ref_proc_prepare(void* h)
{
	RefProc* r = (RefProc*)h;
	if(PS3D_PROC_VERTEX == r->kind) dynamic_cast<PuresoftVertexProcessor*>(r->proc)->preprocess(&r->uniforms[0]);
	else if(PS3D_PROC_INTERPOLATION == r->kind) dynamic_cast<PuresoftInterpolationProcessor*>(r->proc)->preprocess(&r->uniforms[0]);
	else dynamic_cast<PuresoftFragmentProcessor*>(r->proc)->preprocess(&r->uniforms[0], r->textures);
}
// user: the interpolated varyings (PROCDATA_*), 16-byte aligned
int ps3d_ref_fp_process(void* h, int x, int y, void* user, uint32_t* bgra)
{
	FragmentProcessorInput in; in.position[0] = x; in.position[1] = y; in.user = user;
	CaptureOutput out;
	dynamic_cast<PuresoftFragmentProcessor*>(((RefProc*)h)->proc)->process(&in, &out);
	*bgra = out.bgra;
	return out.flags;
}
// The interpolater's use of an interpolation processor on one span (interp.cpp:26-92): start / end from the two ends' corrected
// contributions, step = calcStep, optional left-clip skip, `steps` single steps, then correctInterpolation. All buffers 16-byte
// aligned, userDataBytes() each; contributes: 4 floats per end.
void ps3d_ref_ip_span(void* h, const void* v0, const void* v1, const void* v2, const float* contribL, const float* contribR, int stepCount, int skip, int steps,
                      float correctionFactor2, void* start, void* step, void* fragment)
{
	PuresoftInterpolationProcessor* ip = dynamic_cast<PuresoftInterpolationProcessor*>(((RefProc*)h)->proc);
	const void* verts[3] = { v0, v1, v2 };
	const size_t bytes = ip->userDataBytes();
	void* end = NULL;
	if(0 != posix_memalign(&end, 16, bytes ? bytes : 16)) return;
	ip->interpolateByContributes(start, verts, contribL);
	ip->interpolateByContributes(end, verts, contribR);
	ip->calcStep(step, start, end, stepCount);
	if(skip > 0) ip->stepForward(start, step, skip);
	for(int i = 0; i < steps; i++) ip->stepForward(start, step, 1);
	ip->correctInterpolation(fragment, start, correctionFactor2);
	free(end);
}
// PuresoftFBO::blend4 (fbo.cpp:208-229, MMX asm -> intrinsics in this build) on one pixel: dst is written, src blended over it through
// the sequential cursor the fragment thread uses (fragthrd.cpp:70-82), the result read back
uint32_t ps3d_ref_blend4(uint32_t src, uint32_t dst)
{
	static PuresoftFBO* f = NULL;
	if(!f) f = new PuresoftFBO(4, 16, 2, 4);
	f->directWrite4(0, 1, &dst);
	f->setCurRow(0, 0);
	f->setCurCol(0, 1);
	f->blend4(0, (const unsigned char*)&src);
	uint32_t out = 0;
	f->directRead4(0, 1, &out);
	return out;
}
// slots: 16 pointers to this vertex's element in every attached stream (NULL = not attached); user: 16-byte aligned PROCDATA_*
void ps3d_ref_vp_process(void* h, const void* const* slots, float* position4, void* user)
{
	VertexProcessorInput in;
	for(size_t i = 0; i < MAX_VBOS; i++) in.data[i] = slots[i];
	VertexProcessorOutput out; out.user = user;
	dynamic_cast<PuresoftVertexProcessor*>(((RefProc*)h)->proc)->process(&in, &out);
	memcpy(position4, out.position, 16);
}
}
