// The product's shader functors and samplers — puresoft3d_b200/csrc/shaders.cuh, compiled here for the HOST with a few stand-ins
// for CUDA built-ins — against the reference's own processor classes (oracle/_ref/libps3d_ref.so, driven one call at a time
// through the ps3d_ref_* test hooks of oracle/ref_shim/ref_capi.cpp): every vertex functor on random vertices, every fragment
// functor on random interpolated varyings, same uniforms, same texture bytes (2-D BGRA, float shadow maps, cube maps), the
// rcpps / rsqrtss tables measured on this CPU. Outputs must be identical bit for bit (position, varyings; colour word,
// wrote / discarded / blendable). Prints "<functor> <checked> <mismatches>"; exit code = number of functors with a mismatch.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <xmmintrin.h>
// ---- stand-ins for the CUDA built-ins shaders.cuh uses (the host build of exact_math.cuh covers the arithmetic itself)
struct float4 { float x, y, z, w; }; struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = { x, y, z, w }; return r; }
template<class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float2int_rn(float f) { return _mm_cvtss_si32(_mm_set_ss(f)); }   // round to nearest even, like cvtps2dq
#define __host__
#define __device__
#define __noinline__
#include "shaders.cuh"
#include "x86_approx.h"

extern "C" {
void* ps3d_ref_fbo_create(int width, int height, int elemLen, int wrapMode, int layers, const void* const* pixels);
void ps3d_ref_fbo_destroy(void* f);
void* ps3d_ref_proc_create(int kind, int functor);
void ps3d_ref_proc_destroy(void* h);
size_t ps3d_ref_proc_user_bytes(void* h);
void ps3d_ref_proc_set_uniform(void* h, int slot, const void* data, size_t len);
void ps3d_ref_proc_set_texture(void* h, int index, void* fbo);
void ps3d_ref_proc_prepare(void* h);
int ps3d_ref_fp_process(void* h, int x, int y, void* user, uint32_t* bgra);
void ps3d_ref_vp_process(void* h, const void* const* slots, float* position4, void* user);
void ps3d_ref_ip_span(void* h, const void* v0, const void* v1, const void* v2, const float* contribL, const float* contribR, int stepCount, int skip, int steps,
                      float correctionFactor2, void* start, void* step, void* fragment);
uint32_t ps3d_ref_blend4(uint32_t src, uint32_t dst);
}
enum { KIND_V = 0, KIND_I = 1, KIND_F = 2 };   // PS3D_PROC_*

static uint64_t g_state = 0x1234567887654321ull;
static uint32_t rnd32() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return (uint32_t)(g_state >> 16); }
static float uni(float a, float b) { return a + (b - a) * (float)(rnd32() & 0xffffff) / 16777216.0f; }

struct Textures   // the same bytes on both sides
{
	std::vector<std::vector<uint8_t> > bytes[PS_MAX_BOUND_TEX];
	TexDesc desc[PS_MAX_BOUND_TEX];
	void* fbo[PS_MAX_BOUND_TEX];
	int n;
};
// kind: 'c' 2-D BGRA, 's' float shadow map, 'q' cube map (6 layers)
static void makeTextures(Textures& T, const char* kinds, int wrap)
{
	T.n = (int)strlen(kinds);
	for(int k = 0; k < T.n; k++)
	{
		const int w = kinds[k] == 's' ? 96 : (kinds[k] == 'q' ? 16 : 37), h = kinds[k] == 's' ? 96 : (kinds[k] == 'q' ? 16 : 29), layers = kinds[k] == 'q' ? 6 : 1;
		T.bytes[k].assign(layers, std::vector<uint8_t>((size_t)w * h * 4));
		const void* px[6] = { 0, 0, 0, 0, 0, 0 };
		memset(&T.desc[k], 0, sizeof(TexDesc));
		for(int l = 0; l < layers; l++)
		{
			for(int i = 0; i < w * h; i++)
			{
				if(kinds[k] == 's') { const float d = uni(0.2f, 1.0f); memcpy(&T.bytes[k][l][(size_t)i * 4], &d, 4); }
				else { const uint32_t c = rnd32(); memcpy(&T.bytes[k][l][(size_t)i * 4], &c, 4); }
			}
			px[l] = T.bytes[k][l].data();
			T.desc[k].layer[l] = T.bytes[k][l].data();
		}
		T.desc[k].width = w; T.desc[k].height = h; T.desc[k].scanline = w * 4; T.desc[k].wrap = wrap; T.desc[k].nLayers = layers; T.desc[k].elemLen = 4; T.desc[k].filter = 0;
		T.fbo[k] = ps3d_ref_fbo_create(w, h, 4, wrap, layers, px);
	}
}
static void freeTextures(Textures& T) { for(int k = 0; k < T.n; k++) ps3d_ref_fbo_destroy(T.fbo[k]); }

// random uniforms in every slot a functor may read (vectors, matrices, scalars alike), texture ids in the functor's texture slots
template<class F> static void makeUniforms(DrawParams& P, void* ref, const Textures* T)
{
	for(int s = 0; s < PS_UNIFORM_SLOTS; s++)
	{
		for(int k = 0; k < 16; k++) P.u[s][k] = uni(-2.0f, 2.0f);
		if(33 == s) P.u[s][0] = uni(5.0f, 60.0f);
		ps3d_ref_proc_set_uniform(ref, s, P.u[s], 64);
	}
	if(T)
		for(int k = 0; k < T->n; k++)
		{
			const int id = k;
			ps3d_ref_proc_set_uniform(ref, F::texSlot(k), &id, sizeof(int));
			ps3d_ref_proc_set_texture(ref, id, T->fbo[k]);
			P.tex[k] = T->desc[k];
		}
}

// lanes: one letter per varying — v vector in [-1,1], p position in [-3,3], t texcoord (x,y in [0,1.2]), s shadow coord (xyz in [0.05,0.95] * w), k colour [0,255], d direction
template<class PROG> static int fragmentCase(const char* name, int fnF, const char* lanes, const char* texKinds, int wrap, long long n, const ApproxTables& ap)
{
	constexpr int NV = PROG::NV;
	typedef typename PROG::F F;
	void* ref = ps3d_ref_proc_create(KIND_F, fnF);
	if(!ref) { printf("%s 0 1\n", name); return 1; }
	Textures T; T.n = 0;
	makeTextures(T, texKinds, wrap);
	DrawParams P; memset(&P, 0, sizeof(P)); P.approx = ap;
	long long bad = 0;
	alignas(16) F4 in[NV > 0 ? NV : 1];
	for(long long i = 0; i < n; i++)
	{
		if(0 == i % 4096) { makeUniforms<F>(P, ref, &T); ps3d_ref_proc_prepare(ref); }
		for(int k = 0; k < NV; k++)
		{
			const char c = lanes[k];
			if('t' == c) in[k] = f4(uni(0.0f, 1.2f), uni(0.0f, 1.2f), 0, 0);
			else if('s' == c) { const float w = uni(0.8f, 1.25f); in[k] = f4(uni(0.05f, 0.95f) * w, uni(0.05f, 0.95f) * w, uni(0.05f, 0.95f) * w, w); }
			else if('k' == c) in[k] = f4(uni(0, 255), uni(0, 255), uni(0, 255), 255.0f);
			else if('p' == c) in[k] = f4(uni(-3, 3), uni(-3, 3), uni(-3, 3), 0);
			else in[k] = f4(uni(-1, 1), uni(-1, 1), uni(-1, 1), 0);
		}
		FragmentProcessorOutput out; out.discarded = out.wrote = out.blendable = false; out.bgra = 0;
		F::process(in, out, P);
		alignas(16) F4 copy[NV > 0 ? NV : 1];
		memcpy(copy, in, sizeof(copy));
		uint32_t bgra = 0;
		const int flags = ps3d_ref_fp_process(ref, (int)(i & 1023), (int)((i >> 10) & 511), copy, &bgra);
		const int mine = (out.wrote ? 1 : 0) | (out.discarded ? 2 : 0) | ((out.wrote && out.blendable) ? 4 : 0);
		if(mine != flags || (out.wrote && out.bgra != bgra)) bad++;
	}
	printf("%s %lld %lld\n", name, n, bad);
	freeTextures(T);
	ps3d_ref_proc_destroy(ref);
	return bad != 0;
}

template<class PROG> static int vertexCase(const char* name, int fnV, int fnI, long long n, const ApproxTables& ap)
{
	constexpr int NV = PROG::NV;
	typedef typename PROG::V V;
	void* ref = ps3d_ref_proc_create(KIND_V, fnV);
	void* refI = ps3d_ref_proc_create(KIND_I, fnI);
	if(!ref || !refI) { printf("%s 0 1\n", name); return 1; }
	long long bad = 0;
	// PROCDATA_* is NV float4s; the reference's null interpolators still declare one unused float4 (proc.cpp, testproc.cpp IP_Null)
	if(NV > 0 ? ps3d_ref_proc_user_bytes(refI) != (size_t)NV * 16 : ps3d_ref_proc_user_bytes(refI) > 16) bad++;
	DrawParams P; memset(&P, 0, sizeof(P)); P.approx = ap;
	alignas(16) float slots[16][4];
	for(long long i = 0; i < n; i++)
	{
		if(0 == i % 4096) { makeUniforms<typename PROG::F>(P, ref, NULL); ps3d_ref_proc_prepare(ref); }
		VertexProcessorInput in;
		const void* refSlots[16];
		for(int s = 0; s < 16; s++)
		{
			for(int k = 0; k < 4; k++) slots[s][k] = uni(-2.0f, 2.0f);
			if(0 == s) slots[s][3] = 1.0f;
			in.data[s] = ((V::SLOTS >> s) & 1) ? (const uint8_t*)slots[s] : NULL;
			refSlots[s] = in.data[s];
		}
		VertexProcessorOutput<NV> out; memset(&out, 0, sizeof(out));
		V::process(in, out, P);
		alignas(16) float pos[4];
		alignas(16) float user[(NV > 0 ? NV : 1) * 4];
		memset(user, 0, sizeof(user));
		ps3d_ref_vp_process(ref, refSlots, pos, user);
		if(memcmp(pos, &out.position, 16) != 0 || (NV > 0 && memcmp(user, out.user, (size_t)NV * 16) != 0)) bad++;
	}
	printf("%s %lld %lld\n", name, n, bad);
	ps3d_ref_proc_destroy(ref); ps3d_ref_proc_destroy(refI);
	return bad != 0;
}

// the interpolation processor over one span, the way PuresoftInterpolater drives it (interp.cpp:26-92)
template<class PROG> static int interpolationCase(const char* name, int fnI, long long n)
{
	constexpr int NV = PROG::NV;
	typedef typename PROG::I IP;
	if(0 == NV) return 0;
	void* ref = ps3d_ref_proc_create(KIND_I, fnI);
	if(!ref) { printf("%s 0 1\n", name); return 1; }
	DrawParams P; memset(&P, 0, sizeof(P));
	makeUniforms<typename PROG::F>(P, ref, NULL);
	ps3d_ref_proc_prepare(ref);
	long long bad = 0;
	alignas(16) F4 v0[NV], v1[NV], v2[NV], start[NV], end[NV], step[NV], frag[NV], rStart[NV], rStep[NV], rFrag[NV];
	for(long long i = 0; i < n; i++)
	{
		for(int k = 0; k < NV; k++)
		{
			v0[k] = f4(uni(-3, 3), uni(-3, 3), uni(-3, 3), uni(-3, 3));
			v1[k] = f4(uni(-3, 3), uni(-3, 3), uni(-3, 3), uni(-3, 3));
			v2[k] = f4(uni(-3, 3), uni(-3, 3), uni(-3, 3), uni(-3, 3));
		}
		// corrected contributions of the two ends: two vertices of an edge each, times 1/w (slightly outside [0,1] happens: rounded columns)
		alignas(16) float cl[4] = { 0, 0, 0, 0 }, cr[4] = { 0, 0, 0, 0 };
		const int a = (int)(rnd32() % 3), b = (a + 1 + (int)(rnd32() % 2)) % 3, c = (int)(rnd32() % 3), d = (c + 1 + (int)(rnd32() % 2)) % 3;
		const float t = uni(-0.1f, 1.1f), u = uni(-0.1f, 1.1f);
		cl[a] = t * uni(0.3f, 2.0f); cl[b] = (1.0f - t) * uni(0.3f, 2.0f);
		cr[c] = u * uni(0.3f, 2.0f); cr[d] = (1.0f - u) * uni(0.3f, 2.0f);
		const int stepCount = (int)(rnd32() % 40) - (0 == rnd32() % 16 ? 40 : 0);       // also zero and negative (right < left never reaches the IP, zero does)
		const int skip = 0 == rnd32() % 4 ? (int)(rnd32() % 30) : 0, steps = (int)(rnd32() % 20);
		const float inv = uni(0.2f, 3.0f);
		IP::interpolateByContributes(start, v0, v1, v2, cl[0], cl[1], cl[2]);
		IP::interpolateByContributes(end, v0, v1, v2, cr[0], cr[1], cr[2]);
		IP::calcStep(step, start, end, stepCount);
		if(skip > 0) IP::stepForward(start, step, skip);
		for(int s = 0; s < steps; s++) IP::stepForward(start, step, 1);
		IP::correctInterpolation(frag, start, inv);
		ps3d_ref_ip_span(ref, v0, v1, v2, cl, cr, stepCount, skip, steps, inv, rStart, rStep, rFrag);
		if(memcmp(start, rStart, sizeof(start)) != 0 || memcmp(step, rStep, sizeof(step)) != 0 || memcmp(frag, rFrag, sizeof(frag)) != 0) bad++;
	}
	printf("%s %lld %lld\n", name, n, bad);
	ps3d_ref_proc_destroy(ref);
	return bad != 0;
}

int main(int argc, char** argv)
{
	const long long n = argc > 1 ? atoll(argv[1]) : 200000;
	Ps3dHostApprox host;
	const bool have = ps3d_measure_x86_approx(&host);
	ApproxTables ap; ap.rcp = have ? host.rcp.data() : 0; ap.rsqrt = have ? host.rsqrt.data() : 0; ap.rcpBits = have ? host.rcpBits : 0; ap.rsqrtBits = have ? host.rsqrtBits : 0;
	printf("tables %d %d\n", have ? 1 : 0, 0);
	int failing = 0;
	// ---- vertex functors (functor ids of include/ps3d.h)
	failing += vertexCase<ProgDEF01>("V_DEF01", 1, 1, n, ap);
	failing += vertexCase<ProgDEF02>("V_DEF02", 2, 2, n, ap);
	failing += vertexCase<ProgDEF03>("V_DEF03", 3, 3, n, ap);
	failing += vertexCase<ProgDEF04>("V_DEF04", 4, 4, n, ap);
	failing += vertexCase<ProgDEF05>("V_DEF05", 5, 5, n, ap);
	failing += vertexCase<ProgEarth>("V_Planet", 16, 16, n, ap);
	failing += vertexCase<ProgCloud>("V_Cloud", 18, 18, n, ap);
	failing += vertexCase<ProgCloudShadow>("V_CloudShadow", 19, 19, n, ap);
	failing += vertexCase<ProgPositionOnly>("V_PositionOnly", 32, 32, n, ap);
	failing += vertexCase<ProgSingleColour>("V_SingleColour", 33, 33, n, ap);
	failing += vertexCase<ProgDiffuseOnly>("V_DiffuseOnly", 34, 34, n, ap);
	failing += vertexCase<ProgShadow2>("V_Shadow2", 35, 32, n, ap);
	// ---- blend4 (fbo.cpp:208-229): every (source, destination, alpha) of one channel with the other channels random: 16.7 M blends
	{
		long long checked = 0, bad = 0;
		for(uint32_t a = 0; a < 256; a++)
			for(uint32_t sv = 0; sv < 256; sv++)
				for(uint32_t dv = 0; dv < 256; dv += (n >= 1000000 ? 1 : 5))
				{
					const int ch = (int)(rnd32() % 3);
					const uint32_t r = rnd32(), q = rnd32();
					const uint32_t src = ((r & 0x00ffffffu) & ~(0xffu << (8 * ch))) | (sv << (8 * ch)) | (a << 24);
					const uint32_t dst = (q & ~(0xffu << (8 * ch))) | (dv << (8 * ch));
					checked++;
					if(blend4(src, dst) != ps3d_ref_blend4(src, dst)) bad++;
				}
		printf("blend4 %lld %lld\n", checked, bad);
		failing += bad != 0;
	}
	// ---- interpolation processors
	failing += interpolationCase<ProgDEF01>("I_DEF01", 1, n);
	failing += interpolationCase<ProgDEF02>("I_DEF02", 2, n);
	failing += interpolationCase<ProgDEF03>("I_DEF03", 3, n);
	failing += interpolationCase<ProgDEF04>("I_DEF04", 4, n);
	failing += interpolationCase<ProgEarth>("I_Planet", 16, n);
	failing += interpolationCase<ProgCloud>("I_Cloud", 18, n);
	failing += interpolationCase<ProgCloudShadow>("I_CloudShadow", 19, n);
	failing += interpolationCase<ProgSingleColour>("I_SingleColour", 33, n);
	failing += interpolationCase<ProgDiffuseOnly>("I_DiffuseOnly", 34, n);
	// ---- fragment functors: varyings as the vertex functors lay them out, textures in texSlot() order
	for(int wrap = 0; wrap < 2; wrap++)
	{
		failing += fragmentCase<ProgDEF01>(wrap ? "F_DEF01_wrap" : "F_DEF01", 1, "vpt", "c", wrap, n, ap);
		failing += fragmentCase<ProgDEF03>(wrap ? "F_DEF03_wrap" : "F_DEF03", 3, "vvvpt", "cc", wrap, n, ap);
	}
	failing += fragmentCase<ProgDEF02>("F_DEF02", 2, "vpk", "", 0, n, ap);
	failing += fragmentCase<ProgDEF04>("F_DEF04_cube", 4, "d", "q", 0, n, ap);
	failing += fragmentCase<ProgDEF05>("F_DEF05", 5, "", "", 0, n / 10, ap);
	failing += fragmentCase<ProgEarth>("F_Earth", 16, "vvvpts", "ccccs", 0, n, ap);
	failing += fragmentCase<ProgSatellite>("F_Satellite", 17, "vvvpts", "ccs", 0, n, ap);
	failing += fragmentCase<ProgCloud>("F_Cloud", 18, "vpts", "cs", 0, n, ap);
	failing += fragmentCase<ProgCloudShadow>("F_CloudShadow", 19, "t", "c", 0, n, ap);
	failing += fragmentCase<ProgPositionOnly>("F_SingleColourNoLighting", 32, "", "", 0, n / 10, ap);
	failing += fragmentCase<ProgSingleColour>("F_SingleColour", 33, "vps", "s", 0, n, ap);
	failing += fragmentCase<ProgDiffuseOnly>("F_DiffuseOnly", 34, "vpts", "sc", 0, n, ap);
	failing += fragmentCase<ProgShadow2>("F_Null", 35, "", "", 0, n / 10, ap);
	return failing;
}
