/* ps3d_oracle.c — CPU restatement of Puresoft3D's per-frame rasterisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This file is the checker for the CUDA library: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it. It implements include/ps3d.h (same C-ABI as the product) with a
 * single-threaded, scalar restatement of the reference algorithm; every function cites the reference file:line
 * it follows (paths relative to /root/reference/src/puresoft3d unless stated).
 *
 * PARITY PIN: the reference ships no golden vectors or known-answer tests, so this restatement is
 * pinned against the reference ITSELF — oracle/_ref/libps3d_ref.so, the unmodified sources built through
 * oracle/ref_shim — on seeded scenes (tests/test_oracle_vs_reference.py, bit-exact colour, depth and per-pixel
 * shade counts on the same host), and against the fixtures that build wrote to tests/golden/.
 *
 * Numerics: plain IEEE binary32 operations in the reference's order (build with -ffp-contract=off -mfpmath=sse);
 * the two x86 approximations the reference uses in colour maths (rcpps, rsqrtss) are the same hardware
 * instructions here, so on one host oracle == reference bit for bit. They never feed coverage or depth.
 *
 * Single-threaded is equivalent to the reference's row-interleaved workers: each scanline is consumed FIFO by
 * one worker (drawvao.cpp:78-85), so per pixel the order is triangle submission order either way.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#include <xmmintrin.h>
#if defined(__SSE4_1__)
#include <smmintrin.h>
#endif
#include "ps3d.h"

#define MAX_PROCS 256
#define MAX_PROGS 256
#define MAX_VAOS 4096
#define MAX_VBO_OBJS 65536
#define MAX_VARY 8 /* float4 varyings per vertex (PROCDATA_PLANET has 6) */

typedef struct { float v[4]; } f4;

/* ---- the mcemath routines on the path, in their instruction order (src/mcemath/vector.cpp, matrix.cpp) ---- */

static float hsum4(float p0, float p1, float p2, float p3) { return (p0 + p1) + (p2 + p3); } /* haddps x2: vector.cpp:97-98 */
static float dot4(const float* a, const float* b) { return hsum4(a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]); } /* vector.cpp:85-112 */
#ifndef PS3D_ORACLE_IEEE_APPROX
static float rcp_approx(float x) { return _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(x))); }   /* rcpps, vector.cpp:165 */
static float rsqrt_approx(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); } /* rsqrtss, vector.cpp:245 */
#else
/* experiment switch (never the parity pin): correctly rounded 1/x and 1/sqrt(x), i.e. what a GPU computes, to
 * measure how far the x86 approximations move the colours ("approximate instructions") */
static float rcp_approx(float x) { return 1.0f / x; }
static float rsqrt_approx(float x) { return (float)(1.0 / sqrt((double)x)); }
#endif
static void sub4(float* r, const float* a, const float* b) { for(int i = 0; i < 4; i++) r[i] = a[i] - b[i]; }
static void add4(float* r, const float* a, const float* b) { for(int i = 0; i < 4; i++) r[i] = a[i] + b[i]; }
static void mul4s(float* r, float s) { for(int i = 0; i < 4; i++) r[i] = r[i] * s; }       /* vector.cpp:146-156 */
static void div4s(float* r, float s) { float q = rcp_approx(s); for(int i = 0; i < 4; i++) r[i] = r[i] * q; } /* vector.cpp:158-169 */
static float len4(const float* v) { return sqrtf(hsum4(v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3])); } /* vector.cpp:196-223 */
static void norm4(float* v) /* vector.cpp:225-250 */
{
	float q = rsqrt_approx(hsum4(v[0] * v[0], v[1] * v[1], v[2] * v[2], v[3] * v[3]));
	for(int i = 0; i < 4; i++) v[i] = v[i] * q;
}
static void clamp4(float* v, float lo, float hi) /* vector.cpp:370-383: maxps then minps, x86 NaN rule (2nd operand) */
{
	for(int i = 0; i < 4; i++)
	{
		float a = v[i];
		a = (a > lo) ? a : lo;
		a = (a < hi) ? a : hi;
		v[i] = a;
	}
}
/* matrix.cpp:515-558: ((x*c0 + y*c1) + z*c2) + w*c3, column-major, separate mul/add */
static void m4v4(float* r, const float* m, const float* v)
{
	float x = v[0], y = v[1], z = v[2], w = v[3];
	float o[4];
	for(int i = 0; i < 4; i++)
		o[i] = ((x * m[i] + y * m[4 + i]) + z * m[8 + i]) + w * m[12 + i];
	memcpy(r, o, sizeof(o));
}
static void m4m4(float* d, const float* l, const float* r) /* matrix.cpp:588-701 */
{
	float o[16];
	for(int c = 0; c < 4; c++) m4v4(o + 4 * c, l, r + 4 * c);
	memcpy(d, o, sizeof(o));
}
/* proc.h:73-86 */
static float opt_pow(float x, unsigned n)
{
	float pw = 1.0f;
	while(n > 0)
	{
		if(n & 1) pw *= x;
		x *= x;
		n >>= 1;
	}
	return pw;
}
/* x86-64 `(int)float` = cvttss2si: INT_MIN on NaN / out of range */
static int cvtt(float f)
{
	if(!(f > -2147483904.0f && f < 2147483648.0f)) return (int)0x80000000u;
	return (int)f;
}
/* x86-64 `(unsigned int)float` as gcc emits it: cvttss2si r64 then truncate to 32 bits */
static int cvtu(float f)
{
	long long q;
	if(!(f > -9223373136366403584.0f && f < 9223372036854775808.0f)) q = (long long)0x8000000000000000ull;
	else q = (long long)f;
	return (int)(uint32_t)(uint64_t)q;
}

/* ---- storage -------------------------------------------------------------------------------------------- */

typedef struct
{
	int width, height, scanline, elemLen, wrap, topDown;
	int filter;                          /* PS3D_FILTER_* (extension; 0 = the reference's nearest sampler) */
	int nLayers;
	unsigned char* layer[6];
} fbo_t;

typedef struct { size_t unitBytes, unitCount; unsigned char* data; int alive; } vbo_t;
typedef struct { int alive; int vbo[PS3D_MAX_VBOS]; } vao_t;
typedef struct { int alive, kind, functor; } proc_t;
typedef struct { int vp, ip, fp; } prog_t;
typedef struct { size_t capacity; void* data; } uniform_t;

struct ps3d_pipe
{
	int width, height;
	int vpW, vpH, halfW, halfH;          /* rasterizer.cpp:26-31 */
	int band0, band1;                    /* sort-first extension (ps3d_set_row_band): only raster rows [band0, band1) are shaded */
	int behavior;
	fbo_t display[2];
	int back;
	fbo_t defaultDepth;
	fbo_t* textures[PS3D_MAX_TEXTURES * 16];
	int nTextures;
	int depthTex;                        /* -1 = default */
	vbo_t* vbos;
	int nVbos;
	vao_t* vaos;
	int nVaos;
	proc_t procs[MAX_PROCS];
	int nProcs;
	prog_t progs[MAX_PROGS];
	int nProgs;
	int curProg;
	uniform_t uniforms[PS3D_MAX_UNIFORMS];
	ps3d_stats stats;
	int capW, capH;
	uint32_t* capCounts;
	char err[256];
};

static int fail(ps3d_pipe* p, int code, const char* msg)
{
	snprintf(p->err, sizeof(p->err), "%s", msg);
	return code;
}

static int fbo_init(fbo_t* f, int width, int scanline, int height, int elemLen, int topDown, int wrap, int extraLayers)
{
	memset(f, 0, sizeof(*f));
	f->width = width; f->scanline = scanline; f->height = height; f->elemLen = elemLen; f->topDown = topDown; f->wrap = wrap;
	f->nLayers = 1 + extraLayers;
	for(int i = 0; i < f->nLayers; i++)
	{
		void* mem = NULL;
		if(0 != posix_memalign(&mem, 64, (size_t)scanline * height + 64)) return 0;
		memset(mem, 0, (size_t)scanline * height);
		f->layer[i] = (unsigned char*)mem;
	}
	return 1;
}
static void fbo_free(fbo_t* f) { for(int i = 0; i < 6; i++) { free(f->layer[i]); f->layer[i] = NULL; } }

/* fbo.cpp:552-594 clampCoord */
static void fbo_clamp(const fbo_t* f, int* row, int* col)
{
	int maxRow = f->height - 1, maxCol = f->width - 1;
	if(f->wrap == PS3D_WRAP_CLAMP)
	{
		if(*row > maxRow) *row = maxRow; else if(*row < 0) *row = 0;
		if(*col > maxCol) *col = maxCol; else if(*col < 0) *col = 0;
	}
	else
	{
		/* WRAP is modulo (size-1), with a fix-up for negatives (fbo.cpp:582-590). size==1 would divide by zero. */
		*row = maxRow ? *row % maxRow : 0; if(*row < 0) *row += maxRow;
		*col = maxCol ? *col % maxCol : 0; if(*col < 0) *col += maxCol;
	}
}
/* fbo.cpp:287-291 directRead4 (random access is bottom-up row index regardless of topDown: m_rowEntries[row]) */
static uint32_t fbo_read4(const fbo_t* f, int layer, int row, int col)
{
	fbo_clamp(f, &row, &col);
	uint32_t v;
	memcpy(&v, f->layer[layer] + (size_t)row * f->scanline + (size_t)col * 4, 4);
	return v;
}

/* ---- samplers ------------------------------------------------------------------------------------------- */

/* samplr2d.cpp:19-25 : nearest, +0.5f, (unsigned) casts, row from v / col from u */
/* EXTENSION (include/ps3d.h, ps3d_texture_set_filter): bilinear. No reference counterpart — "parity unpinned". */
static float lerp1(float a, float b, float t) { float d = b - a; float m = d * t; return a + m; }
static uint32_t sampler2d_bilinear4(const fbo_t* t, float u, float v)
{
	float x = (float)t->width * u, y = (float)t->height * v;
	float x0 = floorf(x), y0 = floorf(y);
	float fx = x - x0, fy = y - y0;
	int col = cvtt(x0), row = cvtt(y0);
	uint32_t c00 = fbo_read4(t, 0, row, col), c10 = fbo_read4(t, 0, row, col + 1);
	uint32_t c01 = fbo_read4(t, 0, row + 1, col), c11 = fbo_read4(t, 0, row + 1, col + 1);
	uint32_t out = 0;
	for(int ch = 0; ch < 4; ch++)
	{
		float a = (float)((c00 >> (8 * ch)) & 0xff), b = (float)((c10 >> (8 * ch)) & 0xff);
		float c = (float)((c01 >> (8 * ch)) & 0xff), d = (float)((c11 >> (8 * ch)) & 0xff);
		float r = lerp1(lerp1(a, b, fx), lerp1(c, d, fx), fy);
		int q = cvtt(r + 0.5f);
		q = q < 0 ? 0 : (q > 255 ? 255 : q);
		out |= (uint32_t)q << (8 * ch);
	}
	return out;
}
static uint32_t sampler2d_get4(const fbo_t* t, float u, float v)
{
	if(t->filter) return sampler2d_bilinear4(t, u, v);
	int row = cvtu((float)t->height * v + 0.5f);
	int col = cvtu((float)t->width * u + 0.5f);
	return fbo_read4(t, 0, row, col);
}
/* samplrcube.cpp:29-97 texcoordFromDirection — including the Y-major branch choosing +-Y by the sign of Z (:83-94) */
static int cube_texcoord(float* S, float* T, const float* d)
{
	float X = d[0], Y = d[1], Z = d[2];
	if(fabsf(X) > fabsf(Y))
	{
		if(fabsf(X) > fabsf(Z))
		{
			if(X > 0) { *S = (-Z / fabsf(X) + 1.0f) / 2.0f; *T = (-Y / fabsf(X) + 1.0f) / 2.0f; return PS3D_LAYER_XPOS; }
			else      { *S = ( Z / fabsf(X) + 1.0f) / 2.0f; *T = (-Y / fabsf(X) + 1.0f) / 2.0f; return PS3D_LAYER_XNEG; }
		}
		else
		{
			if(Z > 0) { *S = ( X / fabsf(Z) + 1.0f) / 2.0f; *T = (-Y / fabsf(Z) + 1.0f) / 2.0f; return PS3D_LAYER_ZPOS; }
			else      { *S = (-X / fabsf(Z) + 1.0f) / 2.0f; *T = (-Y / fabsf(Z) + 1.0f) / 2.0f; return PS3D_LAYER_ZNEG; }
		}
	}
	else
	{
		if(fabsf(Z) > fabsf(Y))
		{
			if(Z > 0) { *S = ( X / fabsf(Z) + 1.0f) / 2.0f; *T = (-Y / fabsf(Z) + 1.0f) / 2.0f; return PS3D_LAYER_ZPOS; }
			else      { *S = (-X / fabsf(Z) + 1.0f) / 2.0f; *T = (-Y / fabsf(Z) + 1.0f) / 2.0f; return PS3D_LAYER_ZNEG; }
		}
		else
		{
			if(Z > 0) { *S = ( X / fabsf(Y) + 1.0f) / 2.0f; *T = ( Z / fabsf(Y) + 1.0f) / 2.0f; return PS3D_LAYER_YPOS; }
			else      { *S = ( X / fabsf(Y) + 1.0f) / 2.0f; *T = (-Z / fabsf(Y) + 1.0f) / 2.0f; return PS3D_LAYER_YNEG; }
		}
	}
}
/* samplrcube.cpp:119-127 : S feeds the ROW, T the COLUMN; a missing layer would be a NULL deref in the reference */
static uint32_t samplercube_get4(const fbo_t* t, const float* dir)
{
	float S, T;
	int layer = cube_texcoord(&S, &T, dir);
	if(layer >= t->nLayers) layer = 0;
	int row = cvtu((float)t->height * S + 0.5f);
	int col = cvtu((float)t->width * T + 0.5f);
	return fbo_read4(t, layer, row, col);
}
/* samplrproj.cpp:4-38 : rcpps on w, 4 Poisson taps, X-disk offsets go to v (row), Y-disk to u (col), truncation */
static float samplerproj_get(const fbo_t* t, const float* proj)
{
	static const float PX[4] = { -0.94201624f, 0.94558609f, -0.094184101f, 0.34495938f };
	static const float PY[4] = { -0.39906216f, -0.76890725f, -0.92938870f, 0.29387760f };
	float tc[4] = { proj[0], proj[1], proj[2], proj[3] };
	div4s(tc, tc[3]);
	float factor = 1.0f;
	for(int i = 0; i < 4; i++)
	{
		int y = cvtu((float)t->height * (tc[1] + PX[i] / 500.0f));
		int x = cvtu((float)t->width * (tc[0] + PY[i] / 500.0f));
		uint32_t bits = fbo_read4(t, 0, y, x);
		float d;
		memcpy(&d, &bits, 4);
		if(d < tc[2]) factor -= 0.2f;
	}
	return factor;
}

/* ---- shader functors ------------------------------------------------------------------------------------ */

typedef struct
{
	const float* u[48];       /* uniform slots 0..47 as float pointers (NULL when unset); demo 2 reads up to slot 43 */
	const fbo_t* tex[48];     /* texture named by the int in uniform slot i (NULL when not a texture id) */
} shader_env;

static int n_varyings(int functor) /* IP::userDataBytes() / 16 — float4 fields that carry data */
{
	switch(functor)
	{
	case PS3D_FN_PLANET: case PS3D_FN_SATELLITE: return 6; /* PROCDATA_PLANET, src/test/testproc.h:7-16 */
	case PS3D_FN_CLOUD: return 4;                          /* PROCDATA_CLOUD, :75-82 */
	case PS3D_FN_CLOUDSHADOW: return 1;                    /* PROCDATA_CLOUDSHADOW, :122-126 */
	case PS3D_FN_POSITIONONLY: case PS3D_FN_SHADOW2: return 0; /* IP_Null declares 16 bytes and never touches them (src/test2/testproc.cpp:26-45) */
	case PS3D_FN_SINGLECOLOUR: return 3;                   /* PROCDATA_SINGLECOLOUR, src/test2/testproc.h:78-83 */
	case PS3D_FN_DIFFUSEONLY: return 4;                    /* PROCDATA_DIFFUSEONLY, :129-135 */
	case PS3D_FN_DEF01: case PS3D_FN_DEF02: return 3; /* tex1light1.h:4-9, colr1light1.h */
	case PS3D_FN_DEF03: return 5;                      /* tex1bump1light1.h:4-11 */
	case PS3D_FN_DEF04: return 1;                      /* skybox.h:4-7 */
	case PS3D_FN_DEF05: return 0;                      /* shadow.h:4-6 */
	case PS3D_FN_FLATID: return 1;
	case PS3D_FN_TEXPROBE: return 1;
	default: return -1;
	}
}

/* Vertex processors. in[s] = pointer to this vertex's element in slot s (NULL if the slot is empty). Output:
 * clip position + varyings in PROCDATA order. Unset lanes of a varying are left 0 (the reference leaves stale
 * heap bytes there, e.g. texcoord[2..3]; no fragment functor reads them). */
static int run_vp(int functor, const shader_env* e, const unsigned char* const* in, float* pos, f4* vary)
{
	switch(functor)
	{
	case PS3D_FN_DEF01: /* tex1light1.cpp:22-41 — vary: normal, worldPos, texcoord */
	{
		if(!e->u[3] || !e->u[4] || !e->u[5] || !in[0] || !in[3] || !in[4]) return 0;
		m4v4(vary[1].v, e->u[4], (const float*)in[0]);
		memcpy(pos, vary[1].v, 16);
		vary[1].v[3] = 0;
		m4v4(pos, e->u[3], pos);
		m4v4(vary[0].v, e->u[5], (const float*)in[3]);
		vary[2].v[0] = ((const float*)in[4])[0];
		vary[2].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_DEF02: /* colr1light1.cpp:21-39 — slots 0 pos, 1 normal, 2 colour; vary: normal, worldPos, colour */
	{
		if(!e->u[3] || !e->u[4] || !e->u[5] || !in[0] || !in[1] || !in[2]) return 0;
		m4v4(vary[1].v, e->u[4], (const float*)in[0]);
		memcpy(pos, vary[1].v, 16);
		vary[1].v[3] = 0;
		m4v4(pos, e->u[3], pos);
		m4v4(vary[0].v, e->u[5], (const float*)in[1]);
		memcpy(vary[2].v, in[2], 16);
		return 1;
	}
	case PS3D_FN_DEF03: /* tex1bump1light1.cpp:22-45 — vary: tangent, binormal, normal, worldPos, texcoord */
	{
		if(!e->u[3] || !e->u[4] || !e->u[5] || !in[0] || !in[1] || !in[2] || !in[3] || !in[4]) return 0;
		m4v4(vary[3].v, e->u[4], (const float*)in[0]);
		memcpy(pos, vary[3].v, 16);
		vary[3].v[3] = 0;
		m4v4(pos, e->u[3], pos);
		m4v4(vary[0].v, e->u[5], (const float*)in[1]);
		m4v4(vary[1].v, e->u[5], (const float*)in[2]);
		m4v4(vary[2].v, e->u[5], (const float*)in[3]);
		vary[4].v[0] = ((const float*)in[4])[0];
		vary[4].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_DEF04: /* skybox.cpp:23-44 — direction = transpose(V) with last column zeroed * norm(pos - (0,0,1,0)) */
	{
		static const float observer[4] = { 0, 0, 1.0f, 0 };
		if(!e->u[1] || !in[0]) return 0;
		float d[4], inv[16];
		sub4(d, (const float*)in[0], observer);
		d[3] = 0;
		norm4(d);
		const float* V = e->u[1];
		for(int c = 0; c < 4; c++) for(int r = 0; r < 4; r++) inv[c * 4 + r] = V[r * 4 + c]; /* mat4transpose */
		inv[12] = inv[13] = inv[14] = inv[15] = 0;                                         /* zero_vec_ary(inv+12) */
		m4v4(vary[0].v, inv, d);
		memcpy(pos, in[0], 16);
		return 1;
	}
	case PS3D_FN_DEF05: /* shadow.cpp:21-29 — xyz *= 0.9f ; pvm = PV*M ; pos = pvm * p */
	{
		if(!e->u[3] || !e->u[4] || !in[0]) return 0;
		float adj[4], pvm[16];
		memcpy(adj, in[0], 16);
		adj[0] *= 0.9f; adj[1] *= 0.9f; adj[2] *= 0.9f;
		m4m4(pvm, e->u[3], e->u[4]);
		m4v4(pos, pvm, adj);
		return 1;
	}
	case PS3D_FN_PLANET: /* src/test/testproc.cpp:37-74 — vary: tangent, binormal, normal, worldPos, texcoord, shadowcoord */
	{
		if(!e->u[3] || !e->u[4] || !e->u[5] || !e->u[16] || !in[0] || !in[1] || !in[2] || !in[3] || !in[4]) return 0;
		m4v4(vary[3].v, e->u[4], (const float*)in[0]);
		m4v4(vary[5].v, e->u[16], vary[3].v);
		m4v4(pos, e->u[3], vary[3].v);
		vary[3].v[3] = 0;
		m4v4(vary[0].v, e->u[5], (const float*)in[1]);
		m4v4(vary[1].v, e->u[5], (const float*)in[2]);
		m4v4(vary[2].v, e->u[5], (const float*)in[3]);
		vary[4].v[0] = ((const float*)in[4])[0];
		vary[4].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_CLOUD: /* src/test/testproc.cpp:392-425 — vary: normal, worldPos, texcoord, shadowcoord */
	{
		if(!e->u[3] || !e->u[4] || !e->u[5] || !e->u[16] || !in[0] || !in[3] || !in[4]) return 0;
		m4v4(vary[1].v, e->u[4], (const float*)in[0]);
		m4v4(vary[3].v, e->u[16], vary[1].v);
		m4v4(pos, e->u[3], vary[1].v);
		vary[1].v[3] = 0;
		m4v4(vary[0].v, e->u[5], (const float*)in[3]);
		vary[2].v[0] = ((const float*)in[4])[0];
		vary[2].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_CLOUDSHADOW: /* src/test/testproc.cpp:595-610 — xyz *= 0.95f ; pvm = PV*M ; vary: texcoord */
	{
		if(!e->u[3] || !e->u[4] || !in[0] || !in[4]) return 0;
		float adj[4], pvm[16];
		memcpy(adj, in[0], 16);
		adj[0] *= 0.95f; adj[1] *= 0.95f; adj[2] *= 0.95f;
		m4m4(pvm, e->u[3], e->u[4]);
		m4v4(pos, pvm, adj);
		vary[0].v[0] = ((const float*)in[4])[0];
		vary[0].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_POSITIONONLY: /* src/test2/testproc.cpp:17-22 */
		if(!e->u[5] || !in[0]) return 0;
		m4v4(pos, e->u[5], (const float*)in[0]);
		return 1;
	case PS3D_FN_SHADOW2: /* src/test2/testproc.cpp:510-516 */
	{
		if(!e->u[5] || !in[0]) return 0;
		float adj[4];
		memcpy(adj, in[0], 16);
		adj[0] *= 0.9f; adj[1] *= 0.9f; adj[2] *= 0.9f;
		m4v4(pos, e->u[5], adj);
		return 1;
	}
	case PS3D_FN_SINGLECOLOUR: /* src/test2/testproc.cpp:94-120 — vary: normal, worldPos, shadowcoord */
	{
		if(!e->u[0] || !e->u[1] || !e->u[5] || !e->u[6] || !in[0] || !in[3]) return 0;
		m4v4(vary[1].v, e->u[0], (const float*)in[0]);
		m4v4(vary[2].v, e->u[6], vary[1].v);
		vary[1].v[3] = 0;
		m4v4(pos, e->u[5], (const float*)in[0]);
		m4v4(vary[0].v, e->u[1], (const float*)in[3]);
		return 1;
	}
	case PS3D_FN_DIFFUSEONLY: /* src/test2/testproc.cpp:310-342 — vary: normal, worldPos, texcoord, shadowcoord */
	{
		if(!e->u[0] || !e->u[1] || !e->u[5] || !e->u[6] || !in[0] || !in[3] || !in[4]) return 0;
		m4v4(vary[1].v, e->u[0], (const float*)in[0]);
		m4v4(vary[3].v, e->u[6], vary[1].v);
		m4v4(pos, e->u[5], (const float*)in[0]);
		vary[1].v[3] = 0;
		m4v4(vary[0].v, e->u[1], (const float*)in[3]);
		vary[2].v[0] = ((const float*)in[4])[0];
		vary[2].v[1] = ((const float*)in[4])[1];
		return 1;
	}
	case PS3D_FN_TEXPROBE: /* parity-test functor (not in the reference): clip-space position as given; vary0 = (u, v, 0, 0) */
	{
		if(!in[0] || !in[4]) return 0;
		memcpy(pos, in[0], 16);
		vary[0].v[0] = ((const float*)in[4])[0];
		vary[0].v[1] = ((const float*)in[4])[1];
		vary[0].v[2] = vary[0].v[3] = 0.0f;
		return 1;
	}
	case PS3D_FN_FLATID: /* parity-test functor (not in the reference): pos = PV*(M*p); vary0 = slot 6 (flat id colour) */
	{
		if(!e->u[3] || !e->u[4] || !in[0] || !in[6]) return 0;
		m4v4(pos, e->u[4], (const float*)in[0]);
		m4v4(pos, e->u[3], pos);
		memcpy(vary[0].v, in[6], 16);
		return 1;
	}
	}
	return 0;
}

typedef struct { int discarded; int wrote; uint32_t bgra; int blendable; } frag_out;

static void unpack_bgra(float* out, uint32_t c) /* e.g. tex1light1.cpp:158-161 */
{
	out[0] = (float)(c & 0xff); out[1] = (float)((c >> 8) & 0xff); out[2] = (float)((c >> 16) & 0xff); out[3] = (float)(c >> 24);
}
static uint32_t pack_bgr_trunc(const float* c) /* tex1light1.cpp:186-190: (unsigned char) truncation, alpha 0 */
{
	uint32_t b = (uint32_t)(cvtt(c[0]) & 0xff), g = (uint32_t)(cvtt(c[1]) & 0xff), r = (uint32_t)(cvtt(c[2]) & 0xff);
	return b | (g << 8) | (r << 16);
}
/* the Blinn-Phong tail shared by DEF01/02/03 (tex1light1.cpp:163-190) */
static uint32_t blinn_phong(const shader_env* e, float* colour, const float* worldPos, const float* normal)
{
	float L[4], E[4], H[4];
	sub4(L, e->u[7], worldPos);
	float distance = len4(L);
	div4s(L, distance);
	sub4(E, e->u[8], worldPos);
	norm4(E);
	add4(H, E, L);
	norm4(H);
	float lambert = dot4(L, normal);
	float specular = dot4(H, normal);
	specular = specular < 0 ? 0 : specular;
	specular = opt_pow(specular, 50);
	float add = 255.0f * specular;
	for(int i = 0; i < 4; i++) colour[i] = colour[i] + add; /* add_1to4 */
	mul4s(colour, lambert);
	clamp4(colour, 0, 255.0f);
	return pack_bgr_trunc(colour);
}

/* what FP_Earth and FP_Satellite share (src/test/testproc.cpp:240-262): bump normal through the TBN, L = (light-P)/len via rcpps, E, H */
static void planet_lighting(const shader_env* e, uint32_t bumpTexel, const f4* in, float* lambert, float* specular)
{
	float bump[4], tbn[16], L[4], E[4], H[4];
	bump[0] = (float)((bumpTexel >> 16) & 0xff); bump[1] = (float)((bumpTexel >> 8) & 0xff); bump[2] = (float)(bumpTexel & 0xff); bump[3] = 0;
	div4s(bump, 255.0f);
	mul4s(bump, 2.0f);
	for(int i = 0; i < 4; i++) bump[i] = bump[i] - 1.0f;
	memcpy(tbn, in[0].v, 16); memcpy(tbn + 4, in[1].v, 16); memcpy(tbn + 8, in[2].v, 16); memset(tbn + 12, 0, 16);
	m4v4(bump, tbn, bump);
	norm4(bump);
	sub4(L, e->u[7], in[3].v);
	float distance = len4(L);
	div4s(L, distance);
	sub4(E, e->u[8], in[3].v);
	norm4(E);
	add4(H, E, L);
	norm4(H);
	*lambert = dot4(L, bump);
	float sp = dot4(H, bump);
	sp = sp < 0 ? 0 : sp;
	*specular = opt_pow(sp, 50);
}

/* cvtps2dq + packusdw + packuswb on one lane (src/test/testproc.cpp:565-571) */
static uint32_t pack_rne_sat(float f)
{
#if defined(__SSE4_1__)
	__m128i v = _mm_cvtps_epi32(_mm_set1_ps(f));
	v = _mm_packus_epi32(v, v);
	v = _mm_packus_epi16(v, v);
	return (uint32_t)_mm_cvtsi128_si32(v) & 0xff;
#else
	int v = (f >= -2147483648.0f && f < 2147483648.0f) ? (int)lrintf(f) : (int)0x80000000u;
	int u16 = v < 0 ? 0 : (v > 65535 ? 65535 : v);
	int s16 = (int)(short)u16;
	return (uint32_t)(s16 < 0 ? 0 : (s16 > 255 ? 255 : s16));
#endif
}

/* the factors FP_SingleColour / FP_DiffuseOnly share (src/test2/testproc.cpp:230-256, 459-485) */
static void spot_factors(const shader_env* e, const float* worldPos, const float* normal, float* lambertOut, float* specularOut)
{
	static const float fieldOfLight = 6.283185f * (25.0f / 360.0f);
	float L[4], E[4], H[4];
	sub4(L, e->u[20], worldPos);
	norm4(L);
	sub4(E, e->u[22], worldPos);
	norm4(E);
	add4(H, E, L);
	norm4(H);
	float factors[4] = { 0, 0, 0, 0 };
	factors[0] = dot4(L, normal);
	/* <math.h> in C++ resolves acos(float) / cos(float) to the float overloads, so T of opt_pow is float too */
	float yawOfLight = acosf(dot4(L, e->u[21]));
	float cone = 1.0f;
	if(!(yawOfLight < fieldOfLight))
		cone = opt_pow(cosf(yawOfLight - fieldOfLight), 150);
	factors[0] = factors[0] * cone;
	factors[1] = opt_pow(dot4(H, normal), (unsigned)cvtu(e->u[33][0]));
	clamp4(factors, 0, 1.0f);
	*lambertOut = factors[0];
	*specularOut = factors[1];
}

static void run_fp(int functor, const shader_env* e, const f4* in, frag_out* out)
{
	out->discarded = 0; out->wrote = 0; out->blendable = 0; out->bgra = 0;
	switch(functor)
	{
	case PS3D_FN_PLANET: /* FP_Earth, src/test/testproc.cpp:208-288 */
	{
		float c[4], night[4], lambert, specular;
		unpack_bgra(c, sampler2d_get4(e->tex[9], in[4].v[0], in[4].v[1]));
		unpack_bgra(night, sampler2d_get4(e->tex[12], in[4].v[0], in[4].v[1]));
		uint32_t nb = sampler2d_get4(e->tex[10], in[4].v[0], in[4].v[1]);
		float shadowFactor = samplerproj_get(e->tex[15], in[5].v);
		uint32_t sp = sampler2d_get4(e->tex[11], in[4].v[0], in[4].v[1]);
		float specularControl = (float)((sp >> 16) & 0xff) / 255.0f;
		planet_lighting(e, nb, in, &lambert, &specular);
		specular = specular * specularControl;
		float add = 255.0f * specular;
		for(int i = 0; i < 4; i++) c[i] = c[i] + add;
		mul4s(c, lambert);
		if(lambert < 0.2f) add4(c, c, night);
		mul4s(c, shadowFactor);
		clamp4(c, 0, 255.0f);
		out->bgra = pack_bgr_trunc(c);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_SATELLITE: /* FP_Satellite, src/test/testproc.cpp:299-361 */
	{
		float c[4], lambert, specular;
		unpack_bgra(c, sampler2d_get4(e->tex[9], in[4].v[0], in[4].v[1]));
		uint32_t nb = sampler2d_get4(e->tex[10], in[4].v[0], in[4].v[1]);
		float shadowFactor = samplerproj_get(e->tex[15], in[5].v);
		planet_lighting(e, nb, in, &lambert, &specular);
		float add = 255.0f * specular;
		for(int i = 0; i < 4; i++) c[i] = c[i] + add;
		mul4s(c, lambert);
		mul4s(c, shadowFactor);
		clamp4(c, 0, 255.0f);
		out->bgra = pack_bgr_trunc(c);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_CLOUD: /* FP_Cloud, src/test/testproc.cpp:531-574 */
	{
		uint32_t tex = sampler2d_get4(e->tex[9], in[2].v[0], in[2].v[1]);
		float cloud[4] = { 255.0f, 255.0f, 255.0f, (float)((tex >> 16) & 0xff) }, L[4];
		float shadowFactor = samplerproj_get(e->tex[15], in[3].v);
		sub4(L, e->u[7], in[1].v);
		float distance = len4(L);
		div4s(L, distance);
		float lambert = 2.0f * dot4(L, in[0].v);
		float k = lambert * shadowFactor;
		cloud[0] = cloud[0] * k; cloud[1] = cloud[1] * k; cloud[2] = cloud[2] * k; /* mcemaths_mul_3 */
		out->bgra = pack_rne_sat(cloud[0]) | (pack_rne_sat(cloud[1]) << 8) | (pack_rne_sat(cloud[2]) << 16) | (pack_rne_sat(cloud[3]) << 24);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_CLOUDSHADOW: /* FP_CloudShadow, src/test/testproc.cpp:676-685 */
	{
		uint32_t tex = sampler2d_get4(e->tex[9], in[0].v[0], in[0].v[1]);
		if(((tex >> 16) & 0xff) < 150) out->discarded = 1;
		return;
	}
	case PS3D_FN_POSITIONONLY: /* FP_SingleColourNoLighting, src/test2/testproc.cpp:54-67 */
	{
		float c[4];
		memcpy(c, e->u[31], 16);
		mul4s(c, 255.0f);
		out->bgra = pack_bgr_trunc(c);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_SHADOW2: /* FP_Null, src/test2/testproc.cpp:520-524 */
		return;
	case PS3D_FN_SINGLECOLOUR: /* FP_SingleColour, src/test2/testproc.cpp:221-287 */
	{
		float lambert, specular, c[4], sc[4], ac[4];
		float shadowFactor = samplerproj_get(e->tex[23], in[2].v);
		spot_factors(e, in[1].v, in[0].v, &lambert, &specular);
		memcpy(c, e->u[31], 16);
		mul4s(c, lambert);
		for(int i = 0; i < 4; i++) sc[i] = e->u[31][i] * e->u[32][i];
		mul4s(sc, specular);
		for(int i = 0; i < 4; i++) ac[i] = e->u[31][i] * e->u[30][i];
		add4(c, c, sc);
		mul4s(c, shadowFactor);
		add4(c, c, ac);
		clamp4(c, 0, 1.0f);
		mul4s(c, 255.0f);
		out->bgra = pack_bgr_trunc(c);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_DIFFUSEONLY: /* FP_DiffuseOnly, src/test2/testproc.cpp:448-500 */
	{
		float lambert, specular, c[4], ac[4];
		float shadowFactor = samplerproj_get(e->tex[23], in[3].v);
		unpack_bgra(c, sampler2d_get4(e->tex[40], in[2].v[0], in[2].v[1]));
		for(int i = 0; i < 4; i++) ac[i] = c[i] * e->u[30][i];
		spot_factors(e, in[1].v, in[0].v, &lambert, &specular);
		mul4s(c, (lambert + specular) * shadowFactor);
		add4(c, c, ac);
		clamp4(c, 0, 255.0f);
		out->bgra = pack_bgr_trunc(c);
		out->wrote = 1; out->blendable = 1;
		return;
	}
	case PS3D_FN_DEF01: /* tex1light1.cpp:152-193 — output->write(): plain store even under ALPHABLEND */
	{
		float c[4];
		unpack_bgra(c, sampler2d_get4(e->tex[9], in[2].v[0], in[2].v[1]));
		out->bgra = blinn_phong(e, c, in[1].v, in[0].v);
		out->wrote = 1;
		return;
	}
	case PS3D_FN_DEF02: /* colr1light1.cpp:150-190 */
	{
		float c[4] = { in[2].v[0], in[2].v[1], in[2].v[2], in[2].v[3] };
		out->bgra = blinn_phong(e, c, in[1].v, in[0].v);
		out->wrote = 1;
		return;
	}
	case PS3D_FN_DEF03: /* tex1bump1light1.cpp:180-239 */
	{
		float c[4], bump[4], tbn[16];
		unpack_bgra(c, sampler2d_get4(e->tex[9], in[4].v[0], in[4].v[1]));
		uint32_t nb = sampler2d_get4(e->tex[10], in[4].v[0], in[4].v[1]);
		bump[0] = (float)((nb >> 16) & 0xff); bump[1] = (float)((nb >> 8) & 0xff); bump[2] = (float)(nb & 0xff); bump[3] = 0;
		div4s(bump, 255.0f);
		mul4s(bump, 2.0f);
		for(int i = 0; i < 4; i++) bump[i] = bump[i] - 1.0f; /* sub_4by1 */
		memcpy(tbn, in[0].v, 16); memcpy(tbn + 4, in[1].v, 16); memcpy(tbn + 8, in[2].v, 16); memset(tbn + 12, 0, 16); /* make_tbn */
		m4v4(bump, tbn, bump);
		norm4(bump);
		out->bgra = blinn_phong(e, c, in[3].v, bump);
		out->wrote = 1;
		return;
	}
	case PS3D_FN_DEF04: /* skybox.cpp:127-134 — write4(): blends under ALPHABLEND */
		out->bgra = samplercube_get4(e->tex[2], in[0].v);
		out->wrote = 1;
		out->blendable = 1;
		return;
	case PS3D_FN_DEF05: /* shadow.cpp:76-77 — writes nothing */
		return;
	case PS3D_FN_TEXPROBE: /* parity-test functor: the 2-D sampler's output, unlit */
		out->bgra = sampler2d_get4(e->tex[9], in[0].v[0], in[0].v[1]);
		out->wrote = 1;
		out->blendable = 0;
		return;
	case PS3D_FN_FLATID: /* parity-test functor: id colour rounded to bytes, write4 */
	{
		uint32_t b = (uint32_t)(cvtt(in[0].v[0] + 0.5f) & 0xff), g = (uint32_t)(cvtt(in[0].v[1] + 0.5f) & 0xff);
		uint32_t r = (uint32_t)(cvtt(in[0].v[2] + 0.5f) & 0xff), a = (uint32_t)(cvtt(in[0].v[3] + 0.5f) & 0xff);
		out->bgra = b | (g << 8) | (r << 16) | (a << 24);
		out->wrote = 1;
		out->blendable = 1;
		return;
	}
	}
}

/* fbo.cpp:208-229 blend4: per channel dst + ((sat_u16(src - dst) * srcA) >> 8), 16-bit wrap, unsigned-saturating pack */
static uint32_t blend4(uint32_t src, uint32_t dst)
{
	uint32_t a = src >> 24, out = 0;
	for(int ch = 0; ch < 4; ch++)
	{
		uint32_t s = (src >> (8 * ch)) & 0xff, d = (dst >> (8 * ch)) & 0xff;
		uint32_t diff = s > d ? s - d : 0;               /* _mm_subs_pu16 */
		uint32_t prod = (diff * a) & 0xffff;             /* _mm_mullo_pi16 */
		uint32_t sum = ((prod >> 8) + d) & 0xffff;       /* _mm_srli_pi16, _mm_add_pi16 */
		int16_t ssum = (int16_t)sum;                     /* _mm_packs_pu16: signed 16 -> unsigned 8 saturation */
		uint32_t byte = ssum < 0 ? 0 : (ssum > 255 ? 255 : (uint32_t)ssum);
		out |= byte << (8 * ch);
	}
	return out;
}

/* ---- rasteriser (rasterizer.cpp) ------------------------------------------------------------------------- */

typedef struct { float dx, dy, x0, y0; int i0, i1; } edge_t;

static edge_t make_edge(const float* vx, const float* vy, int i0, int i1) /* LineSegment ctor, rasterizer.cpp:49-64 */
{
	edge_t e;
	e.x0 = vx[i0]; e.y0 = vy[i0];
	e.dx = vx[i1] - e.x0;
	e.dy = vy[i1] - e.y0;
	e.i0 = i0; e.i1 = i1;
	if(fabsf(e.dy) < 0.000001f) e.dy = FLT_MAX;
	return e;
}
static float edge_at(const edge_t* e, float y) { return e->dx * (y - e->y0) / e->dy + e->x0; } /* rasterizer.cpp:66-69 */

typedef struct { int left, leftClamped, lv0, lv1, right, rightClamped, rv0, rv1; } row_t; /* RESULT_ROW, rasterizer.h:12-20 */

typedef struct
{
	float vx[3], vy[3];
	int firstRow, lastRow;
	int nHalves;
	edge_t eL[2], eR[2];
	int y0[2], y1[2]; /* row range of each half, second half written second (rasterizer.cpp:128-139) */
} tri_setup;

static void half_range(const ps3d_pipe* p, float yMin, float yMax, int* iy0, int* iy1) /* processTriangle, rasterizer.cpp:95-97 */
{
	*iy0 = yMin < 0 ? 0 : cvtt(yMin);
	int lastRowIdx = p->vpH - 1;
	*iy1 = yMax > (float)lastRowIdx ? lastRowIdx : cvtt(yMax);
}

/* pushTriangle, rasterizer.cpp:142-234. Returns 0 when the reference returns false (firstRow == lastRow). */
static int setup_triangle(const ps3d_pipe* p, const float pos[3][4], tri_setup* t)
{
	/* pushVertex, rasterizer.cpp:73-90 */
	for(int i = 0; i < 3; i++)
	{
		t->vy[i] = ((float)p->halfH * pos[i][1]) + (float)p->halfH;
		t->vx[i] = ((float)p->halfW * pos[i][0]) + (float)p->halfW;
		if(0 == i) t->firstRow = t->lastRow = cvtt(t->vy[0]);
		else if(t->vy[i] > (float)t->lastRow) t->lastRow = cvtt(t->vy[i]);
		else if(t->vy[i] < (float)t->firstRow) t->firstRow = cvtt(t->vy[i]);
	}
	if(t->firstRow >= p->vpH || t->lastRow < 0) { t->firstRow = 0; t->lastRow = -1; }
	if(t->firstRow < 0) t->firstRow = 0;
	if(t->lastRow >= p->vpH) t->lastRow = p->vpH - 1;
	if(t->firstRow == t->lastRow) return 0;

	const float* vx = t->vx; const float* vy = t->vy;
	int flat = -1, a = 0, b = 0, c = 0; /* a,b = the two vertices with equal y; c = the apex */
	if(vy[0] == vy[1]) { flat = 1; a = 0; b = 1; c = 2; }
	else if(vy[0] == vy[2]) { flat = 1; a = 0; b = 2; c = 1; }
	else if(vy[1] == vy[2]) { flat = 1; a = 1; b = 2; c = 0; }
	if(flat > 0)
	{
		/* rasterizer.cpp:171-203: left edge starts at whichever of a,b has the smaller x */
		t->nHalves = 1;
		if(vx[a] < vx[b]) { t->eL[0] = make_edge(vx, vy, a, c); t->eR[0] = make_edge(vx, vy, b, c); }
		else              { t->eL[0] = make_edge(vx, vy, b, c); t->eR[0] = make_edge(vx, vy, a, c); }
		half_range(p, (float)t->firstRow, (float)t->lastRow, &t->y0[0], &t->y1[0]);
		return 1;
	}
	/* rasterizer.cpp:205-231: sort */
	int top, bottom, third = 2;
	if(vy[0] > vy[1]) { top = 0; bottom = 1; } else { top = 1; bottom = 0; }
	if(vy[top] < vy[third]) { int s = top; top = third; third = s; }
	else if(vy[bottom] > vy[third]) { int s = bottom; bottom = third; third = s; }
	/* processStandingTriangle, rasterizer.cpp:121-140 */
	edge_t eTT = make_edge(vx, vy, top, third), eTB = make_edge(vx, vy, top, bottom), e3B = make_edge(vx, vy, third, bottom);
	t->nHalves = 2;
	if(edge_at(&eTB, vy[third]) > edge_at(&eTT, vy[third]))
	{
		t->eL[0] = eTT; t->eR[0] = eTB; /* upper half */
		t->eL[1] = e3B; t->eR[1] = eTB; /* lower half */
	}
	else
	{
		t->eL[0] = eTB; t->eR[0] = eTT;
		t->eL[1] = eTB; t->eR[1] = e3B;
	}
	half_range(p, vy[third], vy[top], &t->y0[0], &t->y1[0]);
	half_range(p, vy[bottom], vy[third], &t->y0[1], &t->y1[1]);
	return 1;
}

/* One RESULT_ROW, rasterizer.cpp:98-117. Returns 0 when this triangle did not write the row (the reference would
 * then read a stale row from an earlier triangle; cannot happen for rows inside [firstRow,lastRow]). */
static int row_of(const ps3d_pipe* p, const tri_setup* t, int iy, row_t* r)
{
	int h = -1;
	for(int k = t->nHalves - 1; k >= 0; k--) /* the half written last wins */
		if(iy >= t->y0[k] && iy <= t->y1[k]) { h = k; break; }
	if(h < 0) return 0;
	const edge_t* L = &t->eL[h]; const edge_t* R = &t->eR[h];
	r->left = cvtt(edge_at(L, (float)iy) + 0.5f);
	r->leftClamped = r->left < 0 ? 0 : r->left;
	r->right = cvtt(edge_at(R, (float)iy) + 0.5f);
	r->rightClamped = r->right >= p->vpW ? p->vpW - 1 : r->right;
	r->lv0 = L->i0; r->lv1 = L->i1; r->rv0 = R->i0; r->rv1 = R->i1;
	return 1;
}

/* interp.cpp:151-160 lineSegmentlinearInterpolate */
static void edge_contrib(const tri_setup* t, int v1, int v2, float x, float y, float* c)
{
	float dx = t->vx[v1] - t->vx[v2], dy = t->vy[v1] - t->vy[v2];
	c[v1] = fabsf(dx) > fabsf(dy) ? ((x - t->vx[v2]) / dx) : ((y - t->vy[v2]) / dy);
	c[v2] = 1.0f - c[v1];
	c[3 - v1 - v2] = 0;
}

/* ---- the draw ------------------------------------------------------------------------------------------- */

static fbo_t* depth_target(ps3d_pipe* p) { return p->depthTex < 0 ? &p->defaultDepth : p->textures[p->depthTex]; }

/* vertthrd.cpp:56-65 isBackFace */
static int is_back_face(const float* v0, const float* v1, const float* v2)
{
	float a[4], b[4], c[4];
	static const float test[4] = { 0, 0, 1.0f, 0 };
	sub4(a, v1, v0);
	sub4(b, v2, v1);
	/* mcemaths_cross_3, vector.cpp:114-144: a[1 2 0 3]*b[2 0 1 3] - a[2 0 1 3]*b[1 2 0 3] */
	c[0] = a[1] * b[2] - a[2] * b[1];
	c[1] = a[2] * b[0] - a[0] * b[2];
	c[2] = a[0] * b[1] - a[1] * b[0];
	c[3] = a[3] * b[3] - a[3] * b[3];
	norm4(c);
	return dot4(test, c) < 0;
}

int ps3d_draw_vao(ps3d_pipe* p, int vao, int callerThread)
{
	(void)callerThread;
	/* drawvao.cpp:12-15 */
	if(p->curProg < 0 || vao < 0 || vao >= p->nVaos || !p->vaos[vao].alive) return PS3D_OK;
	prog_t pg = p->progs[p->curProg];
	if(pg.vp < 0) return PS3D_OK;
	int fnV = p->procs[pg.vp].functor, fnI = p->procs[pg.ip].functor, fnF = p->procs[pg.fp].functor;
	int nv = n_varyings(fnI);
	if(nv < 0 || nv != n_varyings(fnV) || (n_varyings(fnF) != nv)) return fail(p, PS3D_ERR_UNSUPPORTED, "functor triple has no common varying record");
	p->stats.draws++;

	/* preprocess(): latch uniform pointers (tex1light1.cpp:15-20,145-150 etc.) */
	shader_env env;
	memset(&env, 0, sizeof(env));
	for(int i = 0; i < 48; i++)
	{
		env.u[i] = (const float*)p->uniforms[i].data;
		if(p->uniforms[i].data && p->uniforms[i].capacity >= 4)
		{
			int id;
			memcpy(&id, p->uniforms[i].data, 4);
			if(id >= 0 && id < p->nTextures && p->textures[id]) env.tex[i] = p->textures[id];
		}
	}
	/* the reference dereferences NULL for a missing uniform/texture; report it instead */
	{
		static const int needU[][8] = { {0}, {3, 4, 5, 7, 8, 9, -1}, {3, 4, 5, 7, 8, -1}, {3, 4, 5, 7, 8, 9, 10, -1}, {1, 2, -1}, {3, 4, -1} };
		if(fnV >= 1 && fnV <= 5)
			for(int k = 0; needU[fnV][k] >= 0; k++)
				if(!env.u[needU[fnV][k]]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a uniform slot the programme reads is unset");
		if((fnF == PS3D_FN_DEF01 || fnF == PS3D_FN_DEF03) && !env.tex[9]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "uniform 9 does not name a texture");
		if(fnF == PS3D_FN_DEF03 && !env.tex[10]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "uniform 10 does not name a texture");
		if(fnF == PS3D_FN_DEF04 && !env.tex[2]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "uniform 2 does not name a texture");
		if(fnV == PS3D_FN_FLATID && (!env.u[3] || !env.u[4])) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a uniform slot the programme reads is unset");
		if(fnF == PS3D_FN_TEXPROBE && !env.tex[9]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "uniform 9 does not name a texture");
		{
			/* fragment functors of the two demos: uniform slots read as vectors, then slots that must name a texture */
			static const struct { int fn; int u[8]; int t[6]; } need[] = {
				{ PS3D_FN_PLANET, { 7, 8, -1 }, { 9, 10, 11, 12, 15, -1 } },
				{ PS3D_FN_SATELLITE, { 7, 8, -1 }, { 9, 10, 15, -1 } },
				{ PS3D_FN_CLOUD, { 7, 8, -1 }, { 9, 15, -1 } },
				{ PS3D_FN_CLOUDSHADOW, { -1 }, { 9, -1 } },
				{ PS3D_FN_POSITIONONLY, { 31, -1 }, { -1 } },
				{ PS3D_FN_SINGLECOLOUR, { 20, 21, 22, 30, 31, 32, 33, -1 }, { 23, -1 } },
				{ PS3D_FN_DIFFUSEONLY, { 20, 21, 22, 30, 33, -1 }, { 23, 40, -1 } },
			};
			for(size_t k = 0; k < sizeof(need) / sizeof(need[0]); k++)
				if(need[k].fn == fnF)
				{
					for(int j = 0; need[k].u[j] >= 0; j++)
						if(!env.u[need[k].u[j]]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a uniform slot the programme reads is unset");
					for(int j = 0; need[k].t[j] >= 0; j++)
						if(!env.tex[need[k].t[j]]) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "a texture uniform does not name a live texture");
				}
		}
	}

	/* vbo cursors: all attached slots advance in lock-step, the draw ends when any runs out (vertthrd.cpp:21-31) */
	const vao_t* va = &p->vaos[vao];
	size_t nverts = (size_t)-1;
	int anySlot = 0;
	for(int s = 0; s < PS3D_MAX_VBOS; s++)
		if(va->vbo[s] >= 0)
		{
			anySlot = 1;
			if(p->vbos[va->vbo[s]].unitCount < nverts) nverts = p->vbos[va->vbo[s]].unitCount;
		}
	if(!anySlot) return PS3D_OK; /* the reference would loop forever on an empty VAO */
	size_t ntris = nverts / 3;

	fbo_t* colour = &p->display[p->back];
	fbo_t* depth = depth_target(p);
	const int behavior = p->behavior;
	/* IP_Planet::stepForward advances tangent and binormal by the NORMAL's step (src/test/testproc.cpp:178-188): replicated */
	const int planetQuirk = PS3D_FN_PLANET == fnI;

	for(size_t tri = 0; tri < ntris; tri++)
	{
		float pos[3][4];
		f4 vary[3][MAX_VARY];
		float recipW[4] = { 0, 0, 0, 0 }, projZ[4] = { 0, 0, 0, 0 }; /* drawvao.cpp:25-26 */
		memset(vary, 0, sizeof(vary));
		p->stats.triangles_submitted++;
		/* processVertices, vertthrd.cpp:14-51 */
		for(int i = 0; i < 3; i++)
		{
			const unsigned char* in[PS3D_MAX_VBOS];
			for(int s = 0; s < PS3D_MAX_VBOS; s++)
				in[s] = va->vbo[s] >= 0 ? p->vbos[va->vbo[s]].data + (tri * 3 + i) * p->vbos[va->vbo[s]].unitBytes : NULL;
			if(!run_vp(fnV, &env, in, pos[i], vary[i])) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "vertex functor: a slot or uniform it reads is missing");
			float rw = 1.0f / pos[i][3];
			mul4s(pos[i], rw);
			projZ[i] = pos[i][2];
			recipW[i] = rw;
		}
		if((behavior & PS3D_BEHAVIOR_FACE_CULLING) && is_back_face(pos[0], pos[1], pos[2])) continue; /* drawvao.cpp:46 */
		if(pos[0][2] < -1.0f || pos[0][2] > 1.0f) continue; /* drawvao.cpp:51-56 */
		if(pos[1][2] < -1.0f || pos[1][2] > 1.0f) continue;
		if(pos[2][2] < -1.0f || pos[2][2] > 1.0f) continue;

		tri_setup ts;
		if(!setup_triangle(p, pos, &ts)) continue; /* drawvao.cpp:59 */
		p->stats.triangles_rasterised++;

		for(int row = ts.firstRow; row <= ts.lastRow; row++) /* drawvao.cpp:66 */
		{
			row_t r;
			if(!row_of(p, &ts, row, &r)) continue;
			if(r.left == r.right) continue; /* drawvao.cpp:72 */
			p->stats.spans++;
			if(row < p->band0 || row >= p->band1) continue; /* not in the reference: this rank does not own the row */

			/* interpolateStartAndStep, interp.cpp:26-80 */
			float cl[4], cr[4];
			cl[3] = cr[3] = 0;
			edge_contrib(&ts, r.lv0, r.lv1, (float)r.left, (float)row, cl);
			edge_contrib(&ts, r.rv0, r.rv1, (float)r.right, (float)row, cr);
			for(int i = 0; i < 4; i++) { cl[i] = cl[i] * recipW[i]; cr[i] = cr[i] * recipW[i]; } /* mulvec_3_4, :40-41 */
			int stepCount = r.right - r.left;
			f4 start[MAX_VARY], step[MAX_VARY];
			/* IP::interpolateByContributes + calcStep (tex1light1.cpp:60-107): ((v0*c0)+(v1*c1))+(v2*c2) per lane */
			float rcpCount = 0 == stepCount ? 1.0f : 1.0f / (float)stepCount;
			for(int k = 0; k < nv; k++)
				for(int l = 0; l < 4; l++)
				{
					float s0 = (vary[0][k].v[l] * cl[0] + vary[1][k].v[l] * cl[1]) + vary[2][k].v[l] * cl[2];
					float s1 = (vary[0][k].v[l] * cr[0] + vary[1][k].v[l] * cr[1]) + vary[2][k].v[l] * cr[2];
					start[k].v[l] = s0;
					step[k].v[l] = (s1 - s0) * rcpCount;
				}
			float rcpLen = 1.0f / (float)stepCount; /* interp.cpp:47 */
			float zStart = dot4(cl, projZ);
			float zStep = dot4(cr, projZ);
			zStep = (zStep - zStart) * rcpLen;
			float cf2Start = hsum4(cl[0], cl[1], cl[2], cl[3]); /* interp.cpp:55-68 */
			float cf2Step = hsum4(cr[0], cr[1], cr[2], cr[3]);
			cf2Step = (cf2Step - cf2Start) * rcpLen;
			int skip = r.leftClamped - r.left; /* drawvao.cpp:90 */
			if(skip > 0) /* interp.cpp:74-79 */
			{
				cf2Start += cf2Step * (float)skip;
				zStart += zStep * (float)skip;
				for(int k = 0; k < nv; k++)
					for(int l = 0; l < 4; l++)
					{
						const float st = step[planetQuirk && k < 2 ? 2 : k].v[l];
						if(1 == skip) start[k].v[l] = start[k].v[l] + st;                           /* add_3_4_ip */
						else start[k].v[l] = start[k].v[l] + st * (float)skip;                      /* step_3_4_ip */
					}
			}

			/* fragmentThread, fragthrd.cpp:160-252. FBO cursors (fbo.cpp:98-173): the row is clamped into the
			 * target, the starting column too; once the cursor passes the last column writes are suppressed
			 * (overflow) while reads keep returning the last column. */
			int x1 = r.leftClamped, x2 = r.rightClamped;
			int crow = row > colour->height - 1 ? colour->height - 1 : row;  /* row >= 0 always */
			int drow = row > depth->height - 1 ? depth->height - 1 : row;
			unsigned char* crowPtr = colour->layer[0] + (size_t)(colour->topDown ? colour->height - crow - 1 : crow) * colour->scanline;
			unsigned char* drowPtr = depth->layer[0] + (size_t)(depth->topDown ? depth->height - drow - 1 : drow) * depth->scanline;
			for(int x = x1; x <= x2; x++)
			{
				/* interpolateNextStep, interp.cpp:82-92 */
				float inv = 1.0f / cf2Start;
				cf2Start += cf2Step;
				f4 frag[MAX_VARY];
				for(int k = 0; k < nv; k++)
					for(int l = 0; l < 4; l++)
					{
						frag[k].v[l] = start[k].v[l] * inv;
						start[k].v[l] = start[k].v[l] + step[planetQuirk && k < 2 ? 2 : k].v[l];
					}
				float z = zStart * inv;
				zStart += zStep;
				p->stats.fragments_tested++;

				int dcol = x > depth->width - 1 ? depth->width - 1 : x;
				int dOverflow = x > depth->width - 1;
				float cur = 1.0f;
				if(behavior & PS3D_BEHAVIOR_TEST_DEPTH) memcpy(&cur, drowPtr + (size_t)dcol * 4, 4);
				if(-1.0f < z && (z - cur < -0.0001f)) /* fragthrd.cpp:227 */
				{
					frag_out fo;
					run_fp(fnF, &env, frag, &fo);
					p->stats.fragments_shaded++;
					if(p->capCounts && x < p->capW && row < p->capH) p->capCounts[(size_t)row * p->capW + x]++;
					if(fo.wrote && x <= colour->width - 1)
					{
						uint32_t* px = (uint32_t*)(crowPtr + (size_t)x * 4);
						if(fo.blendable && (behavior & PS3D_BEHAVIOR_ALPHABLEND)) *px = blend4(fo.bgra, *px); /* fragthrd.cpp:70-82 */
						else *px = fo.bgra;
					}
					if(!fo.discarded && (behavior & PS3D_BEHAVIOR_UPDATE_DEPTH) && !dOverflow) /* fragthrd.cpp:234-237 */
						memcpy(drowPtr + (size_t)dcol * 4, &z, 4);
				}
			}
		}
	}
	return PS3D_OK;
}

/* ---- the rest of the pipeline surface (resource tables; pipeline.cpp, tex.cpp, prog.cpp, vao.cpp, vbo.cpp) ---- */

const char* ps3d_backend_name(void) { return "oracle-c"; }
const char* ps3d_last_error(const ps3d_pipe* p) { return p ? p->err : ""; }

int ps3d_create(int width, int height, int device, ps3d_pipe** out)
{
	(void)device;
	if(!out || width <= 0 || height <= 0) return PS3D_ERR_INVALID_ARGUMENT;
	ps3d_pipe* p = (ps3d_pipe*)calloc(1, sizeof(ps3d_pipe));
	if(!p) return PS3D_ERR_BAD_ALLOC;
	p->width = width; p->height = height;
	p->vpW = width; p->vpH = height; p->halfW = width / 2; p->halfH = height / 2;
	p->band0 = 0; p->band1 = 0x7fffffff;
	p->behavior = PS3D_BEHAVIOR_UPDATE_DEPTH | PS3D_BEHAVIOR_TEST_DEPTH | PS3D_BEHAVIOR_FACE_CULLING; /* pipeline.cpp:34 */
	p->depthTex = -1; p->curProg = -1;
	int depthScanline = ((int)(width / 4.0f + 0.5f) * 4) * (int)sizeof(float); /* pipeline.cpp:31 */
	if(depthScanline < width * 4) depthScanline = width * 4; /* the reference under-allocates when W%4==1 */
	if(!fbo_init(&p->defaultDepth, width, depthScanline, height, 4, 0, PS3D_WRAP_CLAMP, 0) ||
	   !fbo_init(&p->display[0], width, width * 4, height, 4, 1, PS3D_WRAP_CLAMP, 0) ||
	   !fbo_init(&p->display[1], width, width * 4, height, 4, 1, PS3D_WRAP_CLAMP, 0))
	{
		free(p);
		return PS3D_ERR_BAD_ALLOC;
	}
	p->back = 1; /* pipeline.cpp:48: the ctor takes the renderer's first swapBuffers() */
	p->vbos = (vbo_t*)calloc(MAX_VBO_OBJS, sizeof(vbo_t));
	p->vaos = (vao_t*)calloc(MAX_VAOS, sizeof(vao_t));
	*out = p;
	return PS3D_OK;
}

int ps3d_destroy(ps3d_pipe* p)
{
	if(!p) return PS3D_ERR_INVALID_ARGUMENT;
	fbo_free(&p->defaultDepth); fbo_free(&p->display[0]); fbo_free(&p->display[1]);
	for(int i = 0; i < p->nTextures; i++) if(p->textures[i]) { fbo_free(p->textures[i]); free(p->textures[i]); }
	for(int i = 0; i < p->nVbos; i++) free(p->vbos[i].data);
	for(int i = 0; i < PS3D_MAX_UNIFORMS; i++) free(p->uniforms[i].data);
	free(p->vbos); free(p->vaos); free(p->capCounts);
	free(p);
	return PS3D_OK;
}

int ps3d_texture_create(ps3d_pipe* p, unsigned width, unsigned scanline, unsigned height, unsigned elemLen,
                        const void* pixels, int extraLayers, int wrapMode, int* idx)
{
	if(1 != elemLen && 4 != elemLen) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "PuresoftFBO: elemLen must be 1 or 4"); /* fbo.cpp:21-24 */
	if(extraLayers < 0 || extraLayers > 5) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftFBO: extraLayers");          /* fbo.cpp:64-67 (LAYER_MAX-1 usable) */
	if(!width || !height || scanline < width * elemLen) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "texture geometry");
	int slot = 0;
	for(; slot < p->nTextures; slot++) if(!p->textures[slot]) break; /* tex.cpp:6-16 first free slot */
	if(slot >= (int)(sizeof(p->textures) / sizeof(p->textures[0]))) return fail(p, PS3D_ERR_BAD_ALLOC, "texture table full");
	fbo_t* f = (fbo_t*)calloc(1, sizeof(fbo_t));
	if(!f || !fbo_init(f, (int)width, (int)scanline, (int)height, (int)elemLen, 0, wrapMode, extraLayers)) { free(f); return fail(p, PS3D_ERR_BAD_ALLOC, "texture"); }
	if(pixels) memcpy(f->layer[0], pixels, (size_t)scanline * height); /* tex.cpp:20-23 */
	p->textures[slot] = f;
	if(slot == p->nTextures) p->nTextures++;
	*idx = slot;
	return PS3D_OK;
}

int ps3d_texture_set_filter(ps3d_pipe* p, int idx, int filter)
{
	if(idx < 0 || idx >= p->nTextures || !p->textures[idx]) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index out of range");
	if(PS3D_FILTER_NEAREST != filter && PS3D_FILTER_BILINEAR != filter) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "filter");
	p->textures[idx]->filter = filter;
	return PS3D_OK;
}

static fbo_t* tex_layer(ps3d_pipe* p, int idx, int layer)
{
	if(idx < 0 || idx >= p->nTextures || !p->textures[idx]) return NULL;
	if(layer < 0 || layer >= p->textures[idx]->nLayers) return NULL;
	return p->textures[idx];
}
int ps3d_texture_upload(ps3d_pipe* p, int idx, int layer, const void* pixels)
{
	fbo_t* f = tex_layer(p, idx, layer);
	if(!f) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index/layer out of range");
	memcpy(f->layer[layer], pixels, (size_t)f->scanline * f->height);
	return PS3D_OK;
}
int ps3d_texture_download(ps3d_pipe* p, int idx, int layer, void* pixels)
{
	fbo_t* f = tex_layer(p, idx, layer);
	if(!f) return fail(p, PS3D_ERR_OUT_OF_RANGE, "getTexture: index/layer out of range");
	memcpy(pixels, f->layer[layer], (size_t)f->scanline * f->height);
	return PS3D_OK;
}
int ps3d_texture_destroy(ps3d_pipe* p, int idx)
{
	if(idx < 0 || idx >= p->nTextures) return fail(p, PS3D_ERR_OUT_OF_RANGE, "destroyTexture: index out of range"); /* tex.cpp:48-51 */
	if(p->textures[idx]) { fbo_free(p->textures[idx]); free(p->textures[idx]); p->textures[idx] = NULL; if(p->depthTex == idx) p->depthTex = -1; }
	return PS3D_OK;
}

int ps3d_vbo_create(ps3d_pipe* p, size_t unitBytes, size_t unitCount, int* vbo)
{
	int slot = 0;
	for(; slot < p->nVbos; slot++) if(!p->vbos[slot].alive) break;
	if(slot >= MAX_VBO_OBJS) return fail(p, PS3D_ERR_BAD_ALLOC, "vbo table full");
	void* mem = NULL;
	if(0 != posix_memalign(&mem, 16, unitBytes * unitCount + 16)) return fail(p, PS3D_ERR_BAD_ALLOC, "vbo"); /* vbo.cpp:14 */
	memset(mem, 0, unitBytes * unitCount + 16);
	p->vbos[slot].unitBytes = unitBytes; p->vbos[slot].unitCount = unitCount; p->vbos[slot].data = (unsigned char*)mem; p->vbos[slot].alive = 1;
	if(slot == p->nVbos) p->nVbos++;
	*vbo = slot;
	return PS3D_OK;
}
static int vbo_ok(ps3d_pipe* p, int v) { return v >= 0 && v < p->nVbos && p->vbos[v].alive; }
int ps3d_vbo_update(ps3d_pipe* p, int vbo, const void* src)
{
	if(!vbo_ok(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	memcpy(p->vbos[vbo].data, src, p->vbos[vbo].unitBytes * p->vbos[vbo].unitCount); /* vbo.cpp:28-31 */
	return PS3D_OK;
}
int ps3d_vbo_destroy(ps3d_pipe* p, int vbo)
{
	if(!vbo_ok(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	free(p->vbos[vbo].data); p->vbos[vbo].data = NULL; p->vbos[vbo].alive = 0;
	for(int v = 0; v < p->nVaos; v++) for(int s = 0; s < PS3D_MAX_VBOS; s++) if(p->vaos[v].alive && p->vaos[v].vbo[s] == vbo) p->vaos[v].vbo[s] = -1;
	return PS3D_OK;
}

int ps3d_vao_create(ps3d_pipe* p, int* vao)
{
	int slot = 0;
	for(; slot < p->nVaos; slot++) if(!p->vaos[slot].alive) break; /* pipeline.cpp:120-134 */
	if(slot >= MAX_VAOS) return fail(p, PS3D_ERR_BAD_ALLOC, "vao table full");
	p->vaos[slot].alive = 1;
	for(int s = 0; s < PS3D_MAX_VBOS; s++) p->vaos[slot].vbo[s] = -1;
	if(slot == p->nVaos) p->nVaos++;
	*vao = slot;
	return PS3D_OK;
}
static int vao_check(ps3d_pipe* p, int vao, int slot) /* pipeline.cpp:140-148 */
{
	if(vao < 0 || vao >= p->nVaos || !p->vaos[vao].alive) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO vao");
	if(slot < 0 || slot >= PS3D_MAX_VBOS) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO idx");
	return PS3D_OK;
}
int ps3d_vao_attach(ps3d_pipe* p, int vao, int slot, int vbo, int* displaced)
{
	if(!vbo_ok(p, vbo)) return fail(p, PS3D_ERR_OUT_OF_RANGE, "vbo");
	int rc = vao_check(p, vao, slot); if(rc) return rc;
	if(displaced) *displaced = p->vaos[vao].vbo[slot];
	p->vaos[vao].vbo[slot] = vbo;
	return PS3D_OK;
}
int ps3d_vao_detach(ps3d_pipe* p, int vao, int slot, int* displaced)
{
	int rc = vao_check(p, vao, slot); if(rc) return rc;
	if(displaced) *displaced = p->vaos[vao].vbo[slot];
	p->vaos[vao].vbo[slot] = -1;
	return PS3D_OK;
}
int ps3d_vao_get(ps3d_pipe* p, int vao, int slot, int* vbo)
{
	int rc = vao_check(p, vao, slot); if(rc) return rc;
	*vbo = p->vaos[vao].vbo[slot];
	return PS3D_OK;
}
int ps3d_vao_destroy(ps3d_pipe* p, int vao)
{
	if(vao < 0 || vao >= p->nVaos) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::attachVBO vao"); /* pipeline.cpp:185-188 */
	if(!p->vaos[vao].alive) return PS3D_OK;
	for(int s = 0; s < PS3D_MAX_VBOS; s++) /* pipeline.cpp:194-201: the pipeline owns attached VBOs */
	{
		int v = p->vaos[vao].vbo[s];
		if(v >= 0 && p->vbos[v].alive) { free(p->vbos[v].data); p->vbos[v].data = NULL; p->vbos[v].alive = 0; }
	}
	p->vaos[vao].alive = 0;
	return PS3D_OK;
}

int ps3d_processor_add(ps3d_pipe* p, int kind, int functor, int* idx)
{
	if(kind < 0 || kind > 2) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "processor kind");
	if(n_varyings(functor) < 0) return fail(p, PS3D_ERR_UNSUPPORTED, "no such functor");
	int slot = 0;
	for(; slot < p->nProcs; slot++) if(!p->procs[slot].alive) break; /* prog.cpp:5-17 */
	if(slot >= MAX_PROCS) return fail(p, PS3D_ERR_BAD_ALLOC, "processor table full");
	p->procs[slot].alive = 1; p->procs[slot].kind = kind; p->procs[slot].functor = functor;
	if(slot == p->nProcs) p->nProcs++;
	*idx = slot;
	return PS3D_OK;
}
int ps3d_processor_destroy(ps3d_pipe* p, int idx)
{
	if(idx < 0 || idx >= p->nProcs) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::destroyProcessor"); /* prog.cpp:24-27 */
	if(p->procs[idx].alive)
	{
		/* prog.cpp:32-37: the in-use programme loses the processor -> drawVAO returns silently */
		if(p->curProg >= 0 && (p->progs[p->curProg].vp == idx || p->progs[p->curProg].ip == idx || p->progs[p->curProg].fp == idx)) p->curProg = -1;
		p->procs[idx].alive = 0;
	}
	return PS3D_OK;
}
int ps3d_programme_create(ps3d_pipe* p, int vid, int iid, int fid, int* idx)
{
	if(vid < 0 || vid >= p->nProcs) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, vid"); /* prog.cpp:45-58 */
	if(iid < 0 || iid >= p->nProcs) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, iid");
	if(fid < 0 || fid >= p->nProcs) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::createProgramme, fid");
	if(!p->procs[vid].alive || !p->procs[iid].alive || !p->procs[fid].alive ||
	   p->procs[vid].kind != PS3D_PROC_VERTEX || p->procs[iid].kind != PS3D_PROC_INTERPOLATION || p->procs[fid].kind != PS3D_PROC_FRAGMENT)
		return fail(p, PS3D_ERR_INVALID_ARGUMENT, "createProgramme: processor kind mismatch");
	int slot = 0;
	for(; slot < p->nProgs; slot++) if(-1 == p->progs[slot].vp) break; /* prog.cpp:60-71 */
	if(slot >= MAX_PROGS) return fail(p, PS3D_ERR_BAD_ALLOC, "programme table full");
	p->progs[slot].vp = vid; p->progs[slot].ip = iid; p->progs[slot].fp = fid;
	if(slot == p->nProgs) p->nProgs++;
	*idx = slot;
	return PS3D_OK;
}
int ps3d_programme_destroy(ps3d_pipe* p, int idx)
{
	if(idx < 0 || idx >= p->nProgs) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::destroyProgramme"); /* prog.cpp:112-115 */
	p->progs[idx].vp = p->progs[idx].ip = p->progs[idx].fp = -1;
	if(p->curProg == idx) p->curProg = -1;
	return PS3D_OK;
}
int ps3d_programme_use(ps3d_pipe* p, int idx)
{
	if(idx < 0 || idx >= p->nProgs || -1 == p->progs[idx].vp) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::useProgramme"); /* prog.cpp:123-126 */
	p->curProg = idx;
	return PS3D_OK;
}

int ps3d_set_viewport(ps3d_pipe* p, int width, int height) /* pipeline.cpp:207-216 -> rasterizer.cpp:26-31 */
{
	if(width <= 0 || height <= 0) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "viewport");
	p->vpW = width; p->vpH = height; p->halfW = width / 2; p->halfH = height / 2;
	return PS3D_OK;
}
int ps3d_set_depth(ps3d_pipe* p, int textureIdx) /* pipeline.cpp:218-239 */
{
	if(-1 == textureIdx) { p->depthTex = -1; return PS3D_OK; }
	if(textureIdx < 0 || textureIdx >= p->nTextures || !p->textures[textureIdx]) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::setDepth");
	if(4 != p->textures[textureIdx]->elemLen || 0 != p->textures[textureIdx]->scanline % 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "PuresoftPipeline::setDepth");
	p->depthTex = textureIdx;
	return PS3D_OK;
}
int ps3d_set_uniform(ps3d_pipe* p, int idx, const void* data, size_t len) /* pipeline.cpp:277-312 */
{
	if(idx < 0 || idx >= PS3D_MAX_UNIFORMS) return fail(p, PS3D_ERR_OUT_OF_RANGE, "PuresoftPipeline::setUniform");
	uniform_t* u = &p->uniforms[idx];
	if(!data) { free(u->data); u->data = NULL; u->capacity = 0; return PS3D_OK; }
	if(u->capacity < len)
	{
		free(u->data);
		void* mem = NULL;
		if(0 != posix_memalign(&mem, 16, len < 64 ? 64 : len)) { u->data = NULL; u->capacity = 0; return fail(p, PS3D_ERR_BAD_ALLOC, "uniform"); }
		memset(mem, 0, len < 64 ? 64 : len);
		u->data = mem; u->capacity = len;
	}
	memcpy(u->data, data, len);
	return PS3D_OK;
}
int ps3d_enable(ps3d_pipe* p, int bits) { p->behavior |= bits; return PS3D_OK; }   /* pipeline.cpp:324-327 */
int ps3d_disable(ps3d_pipe* p, int bits) { p->behavior &= ~bits; return PS3D_OK; } /* pipeline.cpp:329-332 */

int ps3d_clear_depth(ps3d_pipe* p, float furthest) /* pipeline.cpp:334-338 -> clear16 fills every byte (fbo.cpp:348-371) */
{
	fbo_t* d = depth_target(p);
	size_t n = ((size_t)d->scanline * d->height) / 16 * 4; /* whole 16-byte quads only (shr ecx,4) */
	float* f = (float*)d->layer[0];
	for(size_t i = 0; i < n; i++) f[i] = furthest;
	return PS3D_OK;
}
int ps3d_clear_colour(ps3d_pipe* p, uint32_t bgra) /* pipeline.cpp:340-343 -> clear4 skips the LAST buffer row (fbo.cpp:332-346) */
{
	fbo_t* c = &p->display[p->back];
	for(int y = 0; y < c->height - 1; y++)
	{
		uint32_t* row = (uint32_t*)(c->layer[0] + (size_t)y * c->scanline);
		for(int x = 0; x < c->width; x++) row[x] = bgra;
	}
	return PS3D_OK;
}
int ps3d_finish(ps3d_pipe* p) { (void)p; return PS3D_OK; }
int ps3d_swap_buffers(ps3d_pipe* p) { p->back ^= 1; return PS3D_OK; } /* pipeline.cpp:314-322 */

/* postProcess, post.cpp:3-19, with PP_DepthofField::process, src/test2/testpost.cpp:9-43: every worker takes the rows
 * threadIndex, threadIndex + threadCount, ... of the colour target and, two pixels (8 bytes) per step for x = 0, 2, ... < width,
 * adds 50 to each byte with wrap-around (paddb). Rows are disjoint between workers, so the row order is free; with an odd
 * width the last step of a row also covers the first pixel of the next row in memory (scanline = width * 4) — two workers
 * then read-modify-write the same 8 bytes unsynchronised in the reference; the sequential result (both additions land)
 * is what is restated, and the step that would run past the buffer's end on the last row is clipped. */
int ps3d_post_process(ps3d_pipe* p, int functor)
{
	if(functor != PS3D_POST_DEPTHOFFIELD) return fail(p, PS3D_ERR_UNSUPPORTED, "post-processor");
	fbo_t* c = &p->display[p->back];
	unsigned char* base = c->layer[0];
	const size_t bytes = (size_t)c->scanline * (size_t)p->height;
	for(int y = 0; y < p->height; y++)
	{
		unsigned char* row = base + (size_t)y * c->scanline;
		for(int x = 0; x < p->width; x += 2)
			for(int b = 0; b < 8; b++)
			{
				unsigned char* q = row + (size_t)x * 4 + b;
				if((size_t)(q - base) < bytes) *q = (unsigned char)(*q + 50);
			}
	}
	return PS3D_OK;
}

int ps3d_read_colour(ps3d_pipe* p, void* bgra, size_t pitch)
{
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	const fbo_t* c = &p->display[p->back];
	for(int r = 0; r < p->height; r++) memcpy((char*)bgra + (size_t)r * pitch, c->layer[0] + (size_t)r * c->scanline, (size_t)p->width * 4);
	return PS3D_OK;
}
int ps3d_read_depth(ps3d_pipe* p, float* depth, size_t pitch)
{
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	const fbo_t* d = &p->defaultDepth;
	for(int r = 0; r < p->height; r++) memcpy((char*)depth + (size_t)r * pitch, d->layer[0] + (size_t)r * d->scanline, (size_t)p->width * 4);
	return PS3D_OK;
}
int ps3d_write_colour(ps3d_pipe* p, const void* bgra, size_t pitch)
{
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	fbo_t* c = &p->display[p->back];
	for(int r = 0; r < p->height; r++) memcpy(c->layer[0] + (size_t)r * c->scanline, (const char*)bgra + (size_t)r * pitch, (size_t)p->width * 4);
	return PS3D_OK;
}
int ps3d_write_depth(ps3d_pipe* p, const float* depth, size_t pitch)
{
	if(pitch < (size_t)p->width * 4) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "pitch");
	fbo_t* d = &p->defaultDepth;
	for(int r = 0; r < p->height; r++) memcpy(d->layer[0] + (size_t)r * d->scanline, (const char*)depth + (size_t)r * pitch, (size_t)p->width * 4);
	return PS3D_OK;
}

int ps3d_get_stats(ps3d_pipe* p, ps3d_stats* out) { *out = p->stats; return PS3D_OK; }
int ps3d_reset_stats(ps3d_pipe* p) { memset(&p->stats, 0, sizeof(p->stats)); return PS3D_OK; }

int ps3d_debug_capture(ps3d_pipe* p, int width, int height)
{
	if(width < 0 || height < 0) return PS3D_ERR_INVALID_ARGUMENT;
	free(p->capCounts); p->capCounts = NULL; p->capW = p->capH = 0;
	if(width > 0 && height > 0)
	{
		p->capCounts = (uint32_t*)calloc((size_t)width * height, 4);
		if(!p->capCounts) return PS3D_ERR_BAD_ALLOC;
		p->capW = width; p->capH = height;
	}
	return PS3D_OK;
}
int ps3d_debug_read_shade_counts(ps3d_pipe* p, uint32_t* counts)
{
	if(!p->capCounts) return PS3D_ERR_INVALID_ARGUMENT;
	memcpy(counts, p->capCounts, (size_t)p->capW * p->capH * 4);
	return PS3D_OK;
}
int ps3d_debug_clear_shade_counts(ps3d_pipe* p)
{
	if(p->capCounts) memset(p->capCounts, 0, (size_t)p->capW * p->capH * 4);
	return PS3D_OK;
}

int ps3d_set_row_band(ps3d_pipe* p, int row0, int row1)
{
	if(row0 == -1 && row1 == -1) { p->band0 = 0; p->band1 = 0x7fffffff; return PS3D_OK; }
	if(row0 < 0 || row1 < row0) return fail(p, PS3D_ERR_INVALID_ARGUMENT, "row band");
	p->band0 = row0; p->band1 = row1;
	return PS3D_OK;
}
int ps3d_device_colour_ptr(ps3d_pipe* p, void** d, size_t* pitch) { (void)p; (void)d; (void)pitch; return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_depth_ptr(ps3d_pipe* p, void** d, size_t* pitch) { (void)p; (void)d; (void)pitch; return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_stream(ps3d_pipe* p, void** s) { (void)p; (void)s; return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_update_device(ps3d_pipe* p, int vbo, const void* src) { (void)p; (void)vbo; (void)src; return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_update_async(ps3d_pipe* p, int vbo, size_t first, size_t count, const void* src) { (void)p; (void)vbo; (void)first; (void)count; (void)src; return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_device_ptr(ps3d_pipe* p, int vbo, void** d, size_t* b) { (void)p; (void)vbo; (void)d; (void)b; return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_device_written(ps3d_pipe* p, int vbo, void* s) { (void)p; (void)vbo; (void)s; return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_copy_stream(ps3d_pipe* p, void** s) { (void)p; (void)s; return PS3D_ERR_UNSUPPORTED; }
int ps3d_read_colour_async(ps3d_pipe* p, void* dst, size_t pitch) { (void)p; (void)dst; (void)pitch; return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_join(ps3d_pipe* p) { (void)p; return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_unique_id(void* id) { (void)id; return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_init(ps3d_pipe* p, int r, int w, const void* id) { (void)p; (void)r; (void)w; (void)id; return PS3D_ERR_UNSUPPORTED; }
int ps3d_comm_destroy(ps3d_pipe* p) { (void)p; return PS3D_ERR_UNSUPPORTED; }
int ps3d_composite_bands(ps3d_pipe* p, const int* b) { (void)p; (void)b; return PS3D_ERR_UNSUPPORTED; }
int ps3d_peer_export(ps3d_pipe* p, void* b) { (void)p; (void)b; return PS3D_ERR_UNSUPPORTED; }
int ps3d_peer_import(ps3d_pipe* p, int r, int w, const void* b) { (void)p; (void)r; (void)w; (void)b; return PS3D_ERR_UNSUPPORTED; }
int ps3d_composite_peer(ps3d_pipe* p) { (void)p; return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_begin(ps3d_pipe* p) { (void)p; return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_end(ps3d_pipe* p, int* g) { (void)p; (void)g; return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_launch(ps3d_pipe* p, int g) { (void)p; (void)g; return PS3D_ERR_UNSUPPORTED; }
int ps3d_graph_destroy(ps3d_pipe* p, int g) { (void)p; (void)g; return PS3D_ERR_UNSUPPORTED; }
int ps3d_vbo_all_gather(ps3d_pipe* p, int vbo) { (void)p; (void)vbo; return PS3D_ERR_UNSUPPORTED; }
int ps3d_device_launch_count(ps3d_pipe* p, uint64_t* n) { (void)p; if(n) *n = 0; return PS3D_OK; }
int ps3d_debug_batch_counts(ps3d_pipe* p, uint64_t* b, uint64_t* d) { (void)p; if(b) *b = 0; if(d) *d = 0; return PS3D_OK; }
int ps3d_profile_enable(ps3d_pipe* p, int on) { (void)p; (void)on; return PS3D_ERR_UNSUPPORTED; }
int ps3d_profile_read(ps3d_pipe* p, ps3d_profile* out) { (void)p; (void)out; return PS3D_ERR_UNSUPPORTED; }
int ps3d_host_approx_info(int* rcpBits, int* rsqrtBits) { *rcpBits = -1; *rsqrtBits = -1; return PS3D_OK; } /* the hardware instructions themselves */
