// Host build of the product's arithmetic contract (puresoft3d_b200/csrc/exact_math.cuh compiles for the host as well as for
// the device) against the instructions it stands for, on the CPU the test runs on:
//   x86_rcp / x86_rsqrt with the tables x86_approx.cpp measures  ==  rcpps / rsqrtss   (bit for bit, specials included)
//   cvtt / cvtu                                                   ==  (int)f / (unsigned)f as gcc emits them on x86-64
//   hsum4, m4v4, opt_pow                                          ==  the same expressions written out with separate roundings
// Prints one "name checked mismatches" line per check; exit code = number of failing checks.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <xmmintrin.h>
#include "exact_math.cuh"
#include "x86_approx.h"

static uint32_t bitsOf(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float floatOf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static float hwRcp(float x) { return _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(x))); }
static float hwRsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
static uint64_t g_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd32() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return (uint32_t)(g_state >> 16); }
static bool sameBits(float a, float b) { return bitsOf(a) == bitsOf(b) || (a != a && b != b); }   // any NaN == any NaN

int main()
{
	int failing = 0;
	Ps3dHostApprox host;
	const bool have = ps3d_measure_x86_approx(&host);
	printf("tables %d rcpBits %d rsqrtBits %d\n", have ? 1 : 0, host.rcpBits, host.rsqrtBits);
	ApproxTables t;
	t.rcp = have ? host.rcp.data() : 0; t.rsqrt = have ? host.rsqrt.data() : 0; t.rcpBits = have ? host.rcpBits : 0; t.rsqrtBits = have ? host.rsqrtBits : 0;
	const uint32_t specials[] = { 0x00000000u, 0x80000000u, 0x00000001u, 0x807fffffu, 0x00800000u, 0x7f7fffffu, 0xff7fffffu, 0x7f800000u, 0xff800000u,
	                              0x7fc00000u, 0xffc00001u, 0x7f800001u, 0x3f800000u, 0xbf800000u, 0x7e800000u, 0x7f000000u, 0x00400000u, 0x40000000u, 0x3fffffffu };
	if(have)
	{
		long long n = 0, bad = 0;
		for(uint32_t s : specials) { n++; if(!sameBits(x86_rcp(floatOf(s), t), hwRcp(floatOf(s)))) bad++; }
		for(int i = 0; i < 20000000; i++) { const float x = floatOf(rnd32()); n++; if(!sameBits(x86_rcp(x, t), hwRcp(x))) bad++; }
		printf("x86_rcp %lld %lld\n", n, bad); failing += bad != 0;
		n = bad = 0;
		for(uint32_t s : specials) { n++; if(!sameBits(x86_rsqrt(floatOf(s), t), hwRsqrt(floatOf(s)))) bad++; }
		for(int i = 0; i < 20000000; i++) { const float x = floatOf(rnd32()); n++; if(!sameBits(x86_rsqrt(x, t), hwRsqrt(x))) bad++; }
		printf("x86_rsqrt %lld %lld\n", n, bad); failing += bad != 0;
	}
	{
		// gcc on x86-64: (int)f = cvttss2si r32 (0x80000000 when out of range / NaN); (unsigned)f = cvttss2si r64, low 32 bits
		long long n = 0, badT = 0, badU = 0;
		auto check = [&](float f) {
			volatile float vf = f;
			const int wantT = _mm_cvttss_si32(_mm_set_ss(vf));
			const uint32_t wantU = (uint32_t)(unsigned long long)_mm_cvttss_si64(_mm_set_ss(vf));
			n++;
			if(cvtt(f) != wantT) badT++;
			if((uint32_t)cvtu(f) != wantU) badU++;
		};
		for(uint32_t s : specials) check(floatOf(s));
		const float edge[] = { 2147483520.0f, 2147483648.0f, -2147483648.0f, -2147483904.0f, 4294967296.0f, 4294967040.0f, -1.0f, -0.5f, 0.5f, 1.5f, -1.5f,
		                       9223371487098961920.0f, 9223372036854775808.0f, -9223372036854775808.0f, 16777216.0f, 255.99f, -300.25f };
		for(float f : edge) check(f);
		for(int i = 0; i < 20000000; i++) check(floatOf(rnd32()));
		for(int i = 0; i < 5000000; i++) check((float)((int)(rnd32() % 8192) - 2048) + (float)(rnd32() % 1000) / 1000.0f);   // the range rasterisation lives in
		printf("cvtt %lld %lld\n", n, badT); failing += badT != 0;
		printf("cvtu %lld %lld\n", n, badU); failing += badU != 0;
	}
	{
		long long n = 0, bad = 0;
		for(int i = 0; i < 2000000; i++)
		{
			float v[4], m[16];
			for(int k = 0; k < 4; k++) v[k] = (float)((int)(rnd32() % 20001) - 10000) / 977.0f;
			for(int k = 0; k < 16; k++) m[k] = (float)((int)(rnd32() % 20001) - 10000) / 3331.0f;
			// haddps twice: (p0 + p1) + (p2 + p3), every step rounded to float
			volatile float a = v[0] + v[1], b = v[2] + v[3], s = a + b;
			n++; if(bitsOf(hsum4(v[0], v[1], v[2], v[3])) != bitsOf(s)) bad++;
			// matrix.cpp:515-558: ((x*c0 + y*c1) + z*c2) + w*c3 per row, separate multiplies and adds
			const F4 r = m4v4(m, f4(v[0], v[1], v[2], v[3]));
			const float got[4] = { r.x, r.y, r.z, r.w };
			for(int row = 0; row < 4; row++)
			{
				volatile float p0 = v[0] * m[row], p1 = v[1] * m[4 + row], p2 = v[2] * m[8 + row], p3 = v[3] * m[12 + row];
				volatile float s1 = p0 + p1, s2 = s1 + p2, s3 = s2 + p3;
				n++; if(bitsOf(got[row]) != bitsOf(s3)) bad++;
			}
			// proc.h:73-86 opt_pow(x, 50): square-and-multiply, low bit first
			volatile float x = 0.5f + (float)(rnd32() % 1000) / 2000.0f, pw = 1.0f;
			const float x0 = x;
			for(unsigned e = 50; e > 0; e >>= 1) { if(e & 1) pw = pw * x; x = x * x; }
			n++; if(bitsOf(opt_pow(x0, 50)) != bitsOf(pw)) bad++;
		}
		printf("hsum4_m4v4_optpow %lld %lld\n", n, bad); failing += bad != 0;
	}
	return failing;
}
