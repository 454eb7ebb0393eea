// exact_math.cuh — the arithmetic contract of the hot path.
//
// Coverage and depth in Puresoft3D are defined by a specific ORDER of IEEE binary32 operations: the SSE routines of
// src/mcemath never fuse a multiply with an add, sum four lanes as (p0+p1)+(p2+p3) (haddps twice), divide with a true
// divide where the C source says `/` and with the rcpps / rsqrtss hardware approximations where the asm says so
// (SURVEY.md §2a, §9). Everything here is written with the round-to-nearest intrinsics so that nvcc can neither
// contract a*b+c into an FMA nor reassociate, whatever -fmad says.
//
// The two approximate instructions only ever feed colour (never coverage, z or 1/w) but they move 8-bit channels by
// up to 3 LSB inside specular highlights (measured: DESIGN.md "approximate instructions"), which would miss the
// 99.9 % colour gate. rcpps and rsqrtss are table machines: on the CPUs we have seen the result depends only on the
// top k mantissa bits (k=11 for rcpps; k=10 plus exponent parity for rsqrtss) and scales exactly with the exponent. The
// host side of this library measures the CPU it runs on at start-up (x86_approx.cpp), verifies that structure over
// all 2^23 mantissas, and hands the tables to the kernels; x86_rcp()/x86_rsqrt() below then reproduce the host's
// instructions bit for bit, i.e. the kernels render what the reference would render on this very machine. If the
// structure does not hold the tables are absent and correctly rounded 1/x, 1/sqrt(x) are used (PS3D_APPROX=ieee
// forces that).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#define PS_D __device__ __forceinline__
#else
#define PS_HD inline
#define PS_D inline
#endif

struct ApproxTables
{
	const uint32_t* rcp;    // 1<<rcpBits entries: bits of rcpps(1.m) for the binade [1,2)
	const uint32_t* rsqrt;  // 2<<rsqrtBits entries: [0..) binade [1,2), [1<<rsqrtBits..) binade [2,4)
	int rcpBits;            // 0 => tables absent, use IEEE
	int rsqrtBits;
};

#if defined(__CUDA_ARCH__)
PS_D float fmul(float a, float b) { return __fmul_rn(a, b); }
PS_D float fadd(float a, float b) { return __fadd_rn(a, b); }
PS_D float fsub(float a, float b) { return __fsub_rn(a, b); }
PS_D float fdiv(float a, float b) { return __fdiv_rn(a, b); }
PS_D float fsqrt(float a) { return __fsqrt_rn(a); }
PS_D uint32_t fbits(float f) { return __float_as_uint(f); }
PS_D float bitsf(uint32_t u) { return __uint_as_float(u); }
PS_D uint32_t ldtab(const uint32_t* p) { return __ldg(p); }
#else
#include <math.h>
#include <string.h>
// host build (unit tests of the emulation): compile with -ffp-contract=off -mfpmath=sse
PS_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
PS_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
PS_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
PS_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
PS_HD float fsqrt(float a) { return sqrtf(a); }
PS_HD uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PS_HD float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
PS_HD uint32_t ldtab(const uint32_t* p) { return *p; }
#endif

// haddps, haddps : vector.cpp:97-98, interp.cpp:55-66
PS_HD float hsum4(float p0, float p1, float p2, float p3) { return fadd(fadd(p0, p1), fadd(p2, p3)); }

// x86-64 `(int)f` (cvttss2si r32): 0x80000000 on NaN / out of range
PS_HD int cvtt(float f)
{
	if(!(f > -2147483904.0f && f < 2147483648.0f)) return (int)0x80000000u;
	return (int)f;
}
// x86-64 `(unsigned int)f` as gcc emits it: cvttss2si r64, then keep the low 32 bits
PS_HD int cvtu(float f)
{
	long long q;
	if(!(f > -9223373136366403584.0f && f < 9223372036854775808.0f)) q = (long long)0x8000000000000000ull;
	else q = (long long)f;
	return (int)(uint32_t)(unsigned long long)q;
}

// rcpps: vector.cpp:165,190. Specials as measured on x86: zero/denormal -> +-inf, inf -> +-0, results that would be
// denormal flush to +-0, NaN -> quiet NaN.
PS_HD float x86_rcp(float x, const ApproxTables& t)
{
	if(0 == t.rcpBits)
		return fdiv(1.0f, x);
	uint32_t u = fbits(x), sign = u & 0x80000000u;
	int e = (int)((u >> 23) & 0xff);
	uint32_t m = u & 0x7fffffu;
	if(0 == e) return bitsf(sign | 0x7f800000u);
	if(255 == e) return m ? bitsf(u | 0x00400000u) : bitsf(sign);
	uint32_t entry = ldtab(t.rcp + (m >> (23 - t.rcpBits)));
	int re = (int)(entry >> 23) - (e - 127);
	if(re <= 0) return bitsf(sign);
	return bitsf(sign | ((uint32_t)re << 23) | (entry & 0x7fffffu));
}

// rsqrtss: vector.cpp:245. zero/denormal -> +-inf (sign kept for -0), negative -> default NaN, +inf -> +0.
PS_HD float x86_rsqrt(float x, const ApproxTables& t)
{
	if(0 == t.rsqrtBits)
	{
#if defined(__CUDA_ARCH__)
		return (float)(1.0 / sqrt((double)x));
#else
		return (float)(1.0 / sqrt((double)x));
#endif
	}
	uint32_t u = fbits(x), sign = u & 0x80000000u;
	int e = (int)((u >> 23) & 0xff);
	uint32_t m = u & 0x7fffffu;
	if(255 == e && m) return bitsf(u | 0x00400000u);
	if(0 == e) return bitsf(sign | 0x7f800000u);
	if(sign) return bitsf(0xffc00000u);
	if(255 == e) return 0.0f;
	int ue = e - 127;                 // unbiased exponent
	int odd = ue & 1;                 // works for negatives in two's complement
	int j = (ue - odd) / 2;           // x = 1.m * 2^odd * 4^j
	uint32_t entry = ldtab(t.rsqrt + ((uint32_t)odd << t.rsqrtBits) + (m >> (23 - t.rsqrtBits)));
	int re = (int)(entry >> 23) - j;
	return bitsf(((uint32_t)re << 23) | (entry & 0x7fffffu));
}

// ---- float4 helpers in mcemath's order ---------------------------------------------------------------------------

struct F4 { float x, y, z, w; };

PS_HD F4 f4(float x, float y, float z, float w) { F4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PS_HD F4 f4add(F4 a, F4 b) { return f4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w)); }      // vector.cpp:4-15
PS_HD F4 f4sub(F4 a, F4 b) { return f4(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z), fsub(a.w, b.w)); }      // vector.cpp:17-28
PS_HD F4 f4muls(F4 a, float s) { return f4(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s), fmul(a.w, s)); }          // vector.cpp:146-156
PS_HD F4 f4adds(F4 a, float s) { return f4(fadd(a.x, s), fadd(a.y, s), fadd(a.z, s), fadd(a.w, s)); }          // vector.cpp:557-567
PS_HD F4 f4subs(F4 a, float s) { return f4(fsub(a.x, s), fsub(a.y, s), fsub(a.z, s), fsub(a.w, s)); }          // vector.cpp:569-579
PS_HD float f4dot(F4 a, F4 b) { return hsum4(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z), fmul(a.w, b.w)); } // vector.cpp:85-112
PS_HD float f4len(F4 a) { return fsqrt(f4dot(a, a)); }                                                          // vector.cpp:196-223
PS_HD F4 f4divs(F4 a, float s, const ApproxTables& t) { return f4muls(a, x86_rcp(s, t)); }                      // vector.cpp:158-169
PS_HD F4 f4norm(F4 a, const ApproxTables& t) { return f4muls(a, x86_rsqrt(f4dot(a, a), t)); }                   // vector.cpp:225-250
// maxps(v, lo) then minps(., hi) with the x86 rule "second operand when unordered": vector.cpp:370-383
PS_HD float clamp1(float a, float lo, float hi) { a = (a > lo) ? a : lo; a = (a < hi) ? a : hi; return a; }
PS_HD F4 f4clamp(F4 a, float lo, float hi) { return f4(clamp1(a.x, lo, hi), clamp1(a.y, lo, hi), clamp1(a.z, lo, hi), clamp1(a.w, lo, hi)); }

// M*v, column-major, ((x*c0 + y*c1) + z*c2) + w*c3 with separate multiplies and adds: matrix.cpp:515-558
PS_HD F4 m4v4(const float* m, F4 v)
{
	F4 r;
	r.x = fadd(fadd(fadd(fmul(v.x, m[0]), fmul(v.y, m[4])), fmul(v.z, m[8])), fmul(v.w, m[12]));
	r.y = fadd(fadd(fadd(fmul(v.x, m[1]), fmul(v.y, m[5])), fmul(v.z, m[9])), fmul(v.w, m[13]));
	r.z = fadd(fadd(fadd(fmul(v.x, m[2]), fmul(v.y, m[6])), fmul(v.z, m[10])), fmul(v.w, m[14]));
	r.w = fadd(fadd(fadd(fmul(v.x, m[3]), fmul(v.y, m[7])), fmul(v.z, m[11])), fmul(v.w, m[15]));
	return r;
}

// proc.h:73-86 — the multiplication order is part of the result
PS_HD float opt_pow(float x, unsigned n)
{
	float pw = 1.0f;
	while(n > 0)
	{
		if(n & 1) pw = fmul(pw, x);
		x = fmul(x, x);
		n >>= 1;
	}
	return pw;
}
