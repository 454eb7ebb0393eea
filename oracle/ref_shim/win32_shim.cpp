// Linux stand-ins for the Win32 services src/puresoft3d links against. TEST INFRASTRUCTURE ONLY:
// this file exists so that the UNMODIFIED reference pipeline (read in place from /root/reference by
// oracle/ref_shim/build_ref.py) can run headless as the parity pin. No rendering algorithm lives here.
//
//   _beginthreadex/ResumeThread/...  <- pipeline.cpp:53-77, fragthrd.cpp:117-122
//   logicCPUs()                      <- cpu.cpp:23-72 (GetLogicalProcessorInformation)
//   PuresoftGdiRenderer              <- rndrgdi.cpp (GDI+; only a link stub: the harness always passes its
//                                       own headless PuresoftRenderer so pipeline.cpp:255 is never reached)
//   mcemaths::align_base_16/64       <- src/mcemath/wraprs.cpp:7-85 (aligned operator new/delete)
#include "windows.h"
#include <unistd.h>
#include <assert.h>
#include "mcemaths.hpp"
#include "rndrgdi.h"

static unsigned g_nextThreadId = 100;
static __thread unsigned t_threadId = 0;

unsigned ps3d_shim_current_thread_id(void)
{
	if(0 == t_threadId)
		t_threadId = __sync_add_and_fetch(&g_nextThreadId, 1);
	return t_threadId;
}

static void* ps3d_shim_trampoline(void* p)
{
	ps3d_shim_thread* t = (ps3d_shim_thread*)p;
	t_threadId = t->id;
	pthread_mutex_lock(&t->gateLock);
	while(!t->released)
		pthread_cond_wait(&t->gateCond, &t->gateLock);
	pthread_mutex_unlock(&t->gateLock);
	t->entry(t->arg);
	return NULL;
}

uintptr_t _beginthreadex(void*, unsigned, unsigned (*entry)(void*), void* arg, unsigned flags, unsigned*)
{
	ps3d_shim_thread* t = new ps3d_shim_thread;
	t->entry = entry;
	t->arg = arg;
	t->id = __sync_add_and_fetch(&g_nextThreadId, 1);
	t->released = (flags & CREATE_SUSPENDED) ? 0 : 1;
	pthread_mutex_init(&t->gateLock, NULL);
	pthread_cond_init(&t->gateCond, NULL);
	if(0 != pthread_create(&t->tid, NULL, ps3d_shim_trampoline, t))
	{
		delete t;
		return 0;
	}
	return (uintptr_t)t;
}

DWORD ResumeThread(HANDLE h)
{
	ps3d_shim_thread* t = (ps3d_shim_thread*)h;
	pthread_mutex_lock(&t->gateLock);
	t->released = 1;
	pthread_cond_broadcast(&t->gateCond);
	pthread_mutex_unlock(&t->gateLock);
	return 1;
}

DWORD WaitForMultipleObjects(DWORD n, const HANDLE* handles, BOOL, DWORD)
{
	for(DWORD i = 0; i < n; i++)
		pthread_join(((ps3d_shim_thread*)handles[i])->tid, NULL);
	return 0;
}

BOOL CloseHandle(HANDLE h)
{
	ps3d_shim_thread* t = (ps3d_shim_thread*)h;
	pthread_mutex_destroy(&t->gateLock);
	pthread_cond_destroy(&t->gateCond);
	delete t;
	return TRUE;
}

// cpu.cpp stand-in. N >= 2 is mandatory: drawvao.cpp:83 computes row % (N - 1).
int logicCPUs(void)
{
	int n = 0;
	const char* env = getenv("PS3D_REF_THREADS");
	if(env)
		n = atoi(env);
	if(n <= 0)
		n = (int)sysconf(_SC_NPROCESSORS_ONLN);
	if(n < 2)
		n = 2;
	if(n > 1024)
		n = 1024;
	return n;
}

// link stub only
PuresoftGdiRenderer::PuresoftGdiRenderer(void) { throw std::runtime_error("GDI+ renderer is not available in the headless shim"); }
PuresoftGdiRenderer::~PuresoftGdiRenderer(void) {}
void PuresoftGdiRenderer::startup(uintptr_t, int, int) {}
void PuresoftGdiRenderer::shutdown(void) {}
void PuresoftGdiRenderer::setCanvas(uintptr_t) {}
void PuresoftGdiRenderer::getDesc(PURESOFTIMGBUFF32*) {}
void* PuresoftGdiRenderer::swapBuffers(void) { return NULL; }
void PuresoftGdiRenderer::release(void) {}

namespace mcemaths
{
void* align_base_16::operator new(size_t bytes) { return _aligned_malloc(bytes, 16); }
void  align_base_16::operator delete(void* mem) { _aligned_free(mem); }
void* align_base_16::operator new(size_t, void* place) { return place; }
void  align_base_16::operator delete(void*, void*) {}
void* align_base_16::operator new[](size_t bytes) { return _aligned_malloc(bytes, 16); }
void  align_base_16::operator delete[](void* mem) { _aligned_free(mem); }
void* align_base_16::operator new[](size_t, void* place) { return place; }
void  align_base_16::operator delete[](void*, void*) {}
void* align_base_64::operator new(size_t bytes) { return _aligned_malloc(bytes, 64); }
void  align_base_64::operator delete(void* mem) { _aligned_free(mem); }
void* align_base_64::operator new(size_t, void* place) { return place; }
void  align_base_64::operator delete(void*, void*) {}
void* align_base_64::operator new[](size_t bytes) { return _aligned_malloc(bytes, 64); }
void  align_base_64::operator delete[](void* mem) { _aligned_free(mem); }
void* align_base_64::operator new[](size_t, void* place) { return place; }
void  align_base_64::operator delete[](void*, void*) {}
}
