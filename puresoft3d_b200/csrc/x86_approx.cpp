// x86_approx.cpp — measure the host CPU's rcpps / rsqrtss so the kernels can reproduce them bit for bit.
//
// The reference divides and normalises in its fragment shaders with the SSE approximations (src/mcemath/vector.cpp
// :165 rcpps, :245 rsqrtss). Their results are CPU-defined, so "the reference's colours" are only defined per
// machine. This file runs the two instructions over every mantissa of one binade (two for rsqrtss: even and odd
// exponent), finds the smallest k for which the result depends on the top k mantissa bits only, and checks on a
// sample of other exponents and on the special values that exact_math.cuh's x86_rcp()/x86_rsqrt() — compiled here
// for the host — agree with the hardware. Only then are the tables handed to the device. No table ships in the
// repo; nothing here is read from the oracle.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "exact_math.cuh"
#include "x86_approx.h"

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
static float hw_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(x))); }
static float hw_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
#define PS3D_HAVE_SSE 1
#else
#define PS3D_HAVE_SSE 0
#endif

static uint32_t bitsOf(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float floatOf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#if PS3D_HAVE_SSE
// smallest k in [6,16] such that out[] is constant on every aligned run of 2^(23-k) mantissas; 0 if none
static int bucketBits(const std::vector<uint32_t>& out)
{
	for(int k = 6; k <= 16; k++)
	{
		const uint32_t run = 1u << (23 - k);
		bool ok = true;
		for(uint32_t b = 0; b < (1u << 23) && ok; b += run)
			for(uint32_t i = 1; i < run; i++)
				if(out[b + i] != out[b]) { ok = false; break; }
		if(ok) return k;
	}
	return 0;
}
#endif

bool ps3d_measure_x86_approx(Ps3dHostApprox* out)
{
	out->rcpBits = out->rsqrtBits = 0;
	out->rcp.clear();
	out->rsqrt.clear();
	const char* env = getenv("PS3D_APPROX");
	if(env && 0 == strcmp(env, "ieee")) return false;
#if PS3D_HAVE_SSE
	std::vector<uint32_t> full(1u << 23);
	for(uint32_t m = 0; m < (1u << 23); m++) full[m] = bitsOf(hw_rcp(floatOf(0x3f800000u | m)));
	int kr = bucketBits(full);
	if(!kr) return false;
	out->rcp.resize(1u << kr);
	for(uint32_t i = 0; i < (1u << kr); i++) out->rcp[i] = full[i << (23 - kr)];

	std::vector<uint32_t> full2(1u << 23);
	for(uint32_t m = 0; m < (1u << 23); m++) full[m] = bitsOf(hw_rsqrt(floatOf(0x3f800000u | m)));  // [1,2)
	for(uint32_t m = 0; m < (1u << 23); m++) full2[m] = bitsOf(hw_rsqrt(floatOf(0x40000000u | m))); // [2,4)
	int k0 = bucketBits(full), k1 = bucketBits(full2);
	if(!k0 || !k1) { out->rcp.clear(); return false; }
	int ks = k0 > k1 ? k0 : k1;
	out->rsqrt.resize(2u << ks);
	for(uint32_t i = 0; i < (1u << ks); i++)
	{
		out->rsqrt[i] = full[i << (23 - ks)];
		out->rsqrt[(1u << ks) + i] = full2[i << (23 - ks)];
	}

	// cross-check the emulation against the hardware on other exponents and the special values
	ApproxTables t;
	t.rcp = out->rcp.data(); t.rsqrt = out->rsqrt.data(); t.rcpBits = kr; t.rsqrtBits = ks;
	bool ok = true;
	uint32_t lcg = 12345u;
	for(int i = 0; i < 4000000 && ok; i++)
	{
		lcg = lcg * 1664525u + 1013904223u;
		float x = floatOf(lcg);
		uint32_t a = bitsOf(x86_rcp(x, t)), b = bitsOf(hw_rcp(x));
		if(a != b && !((a & 0x7fffffffu) > 0x7f800000u && (b & 0x7fffffffu) > 0x7f800000u)) ok = false;
		a = bitsOf(x86_rsqrt(x, t)); b = bitsOf(hw_rsqrt(x));
		if(a != b && !((a & 0x7fffffffu) > 0x7f800000u && (b & 0x7fffffffu) > 0x7f800000u)) ok = false;
	}
	static const uint32_t specials[] = { 0x00000000u, 0x80000000u, 0x00000001u, 0x807fffffu, 0x00800000u, 0x7f800000u, 0xff800000u,
	                                     0x7f7fffffu, 0x7e800000u, 0x7f000000u, 0x3f800000u, 0xbf800000u, 0x7fc00000u };
	for(size_t i = 0; i < sizeof(specials) / sizeof(specials[0]) && ok; i++)
	{
		float x = floatOf(specials[i]);
		uint32_t a = bitsOf(x86_rcp(x, t)), b = bitsOf(hw_rcp(x));
		if(a != b && !((a & 0x7fffffffu) > 0x7f800000u && (b & 0x7fffffffu) > 0x7f800000u)) ok = false;
		a = bitsOf(x86_rsqrt(x, t)); b = bitsOf(hw_rsqrt(x));
		if(a != b && !((a & 0x7fffffffu) > 0x7f800000u && (b & 0x7fffffffu) > 0x7f800000u)) ok = false;
	}
	if(!ok)
	{
		out->rcp.clear();
		out->rsqrt.clear();
		return false;
	}
	out->rcpBits = kr;
	out->rsqrtBits = ks;
	return true;
#else
	return false;
#endif
}
