/* Win32 stand-in for building the UNMODIFIED Puresoft3D pipeline sources on Linux (test infrastructure).
 * Only what src/puresoft3d touches: threads (pipeline.cpp:53-77, fragthrd.cpp:117-122), Interlocked*
 * (rinque.h:65,104,160-163), _aligned_malloc/_aligned_free (fbo.cpp:41,413; vbo.cpp:14; udm.cpp:30;
 * pipeline.cpp:302) and the BMP structs used by the debug dump (fbo.cpp:473-540, never called here).
 * No algorithmic content. */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <wchar.h>
#include <pthread.h>
#include <math.h>
#include <float.h>
#include <stdexcept>
#include <new>

typedef void* HANDLE;
typedef unsigned int DWORD;
typedef unsigned short WORD;
typedef int BOOL;
typedef long LONG;
typedef uintptr_t ULONG_PTR;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INFINITE 0xFFFFFFFFu
#define CREATE_SUSPENDED 0x4
#define BI_BITFIELDS 3

#pragma pack(push, 2)
typedef struct { WORD bfType; DWORD bfSize; WORD bfReserved1; WORD bfReserved2; DWORD bfOffBits; } BITMAPFILEHEADER;
#pragma pack(pop)
typedef struct { DWORD biSize; LONG biWidth; LONG biHeight; WORD biPlanes; WORD biBitCount; DWORD biCompression;
                 DWORD biSizeImage; LONG biXPelsPerMeter; LONG biYPelsPerMeter; DWORD biClrUsed; DWORD biClrImportant; } BITMAPINFOHEADER;

static inline int _wfopen_s(FILE** fp, const wchar_t*, const wchar_t*) { *fp = NULL; return 1; } /* dumps disabled */

static inline void* _aligned_malloc(size_t bytes, size_t align)
{
	void* p = NULL;
	if(align < sizeof(void*)) align = sizeof(void*);
	if(0 != posix_memalign(&p, align, bytes ? bytes : align)) return NULL;
	return p;
}
static inline void _aligned_free(void* p) { free(p); }

template<typename T> static inline T InterlockedIncrement(volatile T* p) { return __sync_add_and_fetch(p, (T)1); }
template<typename T> static inline T InterlockedDecrement(volatile T* p) { return __sync_sub_and_fetch(p, (T)1); }
template<typename T, typename U> static inline T InterlockedExchange(volatile T* p, U v) { return __sync_lock_test_and_set(p, (T)v); }

/* ---- threads: a HANDLE is a ps3d_shim_thread*; GetCurrentThread() is the Win32-style pseudo handle ---- */
struct ps3d_shim_thread
{
	pthread_t tid;
	unsigned (*entry)(void*);
	void* arg;
	unsigned id;
	pthread_mutex_t gateLock;
	pthread_cond_t gateCond;
	int released;
};
#define PS3D_SHIM_PSEUDO_HANDLE ((HANDLE)(intptr_t)-2)

unsigned ps3d_shim_current_thread_id(void);
uintptr_t _beginthreadex(void* security, unsigned stack, unsigned (*entry)(void*), void* arg, unsigned flags, unsigned* thrdaddr);
DWORD ResumeThread(HANDLE h);
static inline HANDLE GetCurrentThread(void) { return PS3D_SHIM_PSEUDO_HANDLE; }
static inline unsigned GetThreadId(HANDLE h)
{
	return (h == PS3D_SHIM_PSEUDO_HANDLE) ? ps3d_shim_current_thread_id() : ((ps3d_shim_thread*)h)->id;
}
DWORD WaitForMultipleObjects(DWORD n, const HANDLE* handles, BOOL all, DWORD ms);
BOOL CloseHandle(HANDLE h);
