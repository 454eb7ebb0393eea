// ROUND 2: this file is no longer part of oracle/_ref/libps3d_ref.so. The reference build now compiles src/mcemath from its own
// text, every MSVC asm block rewritten instruction by instruction by asm_translate.py. This hand-written restatement of the 26
// `mcemaths_*` routines the pipeline calls (round 1's shim) is kept as an independent CROSS-CHECK of that translation: it is built
// into oracle/_ref/libmcemath_hand.so and compared routine by routine by tests/test_mcemath_translation.py.
// TEST INFRASTRUCTURE ONLY. Every routine keeps the instruction ORDER of the asm (separate mulps/addps, haddps pairing, hardware
// rcpps/rsqrtss) because that order defines the reference's numerics. Built with -ffp-contract=off -mfpmath=sse.
//
// Citations: src/mcemath/vector.cpp, matrix.cpp, quatern.cpp (line ranges beside each function).
#include <xmmintrin.h>
#include <pmmintrin.h>
#include <string.h>

extern "C" {

// vector.cpp:4-15
void mcemaths_add_3_4(float* r4, const float* v4_1, const float* v4_2)
{
	_mm_store_ps(r4, _mm_add_ps(_mm_load_ps(v4_1), _mm_load_ps(v4_2)));
}

// vector.cpp:17-28
void mcemaths_sub_3_4(float* r4, const float* v4_l, const float* v4_r)
{
	_mm_store_ps(r4, _mm_sub_ps(_mm_load_ps(v4_l), _mm_load_ps(v4_r)));
}

// vector.cpp:30-40
void mcemaths_add_3_4_ip(float* v4_1, const float* v4_2)
{
	_mm_store_ps(v4_1, _mm_add_ps(_mm_load_ps(v4_1), _mm_load_ps(v4_2)));
}

// vector.cpp:42-52
void mcemaths_sub_3_4_ip(float* v4_l, const float* v4_r)
{
	_mm_store_ps(v4_l, _mm_sub_ps(_mm_load_ps(v4_l), _mm_load_ps(v4_r)));
}

// vector.cpp:70-83 : target += step * n  (mulps, then addps — not fused)
void mcemaths_step_3_4_ip(float* target4, const float* step4, float n_steps)
{
	__m128 s = _mm_mul_ps(_mm_load_ps(step4), _mm_set1_ps(n_steps));
	_mm_store_ps(target4, _mm_add_ps(_mm_load_ps(target4), s));
}

// vector.cpp:85-112 : mulps ; haddps ; haddps  =>  (p0+p1)+(p2+p3)
float mcemaths_dot_3_4(const float* v4_1, const float* v4_2)
{
	__m128 p = _mm_mul_ps(_mm_load_ps(v4_1), _mm_load_ps(v4_2));
	p = _mm_hadd_ps(p, p);
	p = _mm_hadd_ps(p, p);
	return _mm_cvtss_f32(p);
}

// vector.cpp:114-144 : a[1 2 0 3]*b[2 0 1 3] - a[2 0 1 3]*b[1 2 0 3]
void mcemaths_cross_3(float* r4, const float* v4_l, const float* v4_r)
{
	__m128 a = _mm_load_ps(v4_l), b = _mm_load_ps(v4_r);
	__m128 a120 = _mm_shuffle_ps(a, a, 0xc9);
	__m128 b201 = _mm_shuffle_ps(b, b, 0xd2);
	__m128 a201 = _mm_shuffle_ps(a, a, 0xd2);
	__m128 b120 = _mm_shuffle_ps(b, b, 0xc9);
	_mm_store_ps(r4, _mm_sub_ps(_mm_mul_ps(a120, b201), _mm_mul_ps(a201, b120)));
}

// vector.cpp:146-156 : all four lanes
void mcemaths_mul_3_4(float* r4, float fac)
{
	_mm_store_ps(r4, _mm_mul_ps(_mm_load_ps(r4), _mm_set1_ps(fac)));
}

// vector.cpp:158-169 : rcpps (hardware approximation) then mulps
void mcemaths_div_3_4(float* r4, float fac)
{
	_mm_store_ps(r4, _mm_mul_ps(_mm_load_ps(r4), _mm_rcp_ps(_mm_set1_ps(fac))));
}

// vector.cpp:171-181
void mcemaths_mulvec_3_4(float* r4, const float* fac4)
{
	_mm_store_ps(r4, _mm_mul_ps(_mm_load_ps(r4), _mm_load_ps(fac4)));
}

// vector.cpp:183-194
void mcemaths_divvec_3_4(float* r4, const float* fac4)
{
	_mm_store_ps(r4, _mm_mul_ps(_mm_load_ps(r4), _mm_rcp_ps(_mm_load_ps(fac4))));
}

// vector.cpp:196-223 : mulps ; haddps x2 ; sqrtss (IEEE)
float mcemaths_len_3_4(const float* v4)
{
	__m128 v = _mm_load_ps(v4);
	__m128 p = _mm_mul_ps(v, v);
	p = _mm_hadd_ps(p, p);
	p = _mm_hadd_ps(p, p);
	return _mm_cvtss_f32(_mm_sqrt_ss(p));
}

// vector.cpp:225-250 : mulps ; haddps x2 ; rsqrtss (hardware approximation) ; broadcast ; mulps
void mcemaths_norm_3_4(float* v4)
{
	__m128 v = _mm_load_ps(v4);
	__m128 p = _mm_mul_ps(v, v);
	p = _mm_hadd_ps(p, p);
	p = _mm_hadd_ps(p, p);
	p = _mm_rsqrt_ss(p);
	p = _mm_shuffle_ps(p, p, 0x00);
	_mm_store_ps(v4, _mm_mul_ps(v, p));
}

// vector.cpp:281-306
void mcemaths_zero_vec_ary(float* ary4, int count)
{
	__m128 z = _mm_setzero_ps();
	int i = 0;
	do
	{
		_mm_store_ps(ary4 + 4 * i, z);
	} while(++i < count);
}

// vector.cpp:370-383 : maxps(v, lo) then minps(., hi)   (x86 operand order kept: result = second operand on NaN)
void mcemaths_clamp_3_4(float* v4, float min, float max)
{
	__m128 v = _mm_load_ps(v4);
	v = _mm_max_ps(v, _mm_set1_ps(min));
	v = _mm_min_ps(v, _mm_set1_ps(max));
	_mm_store_ps(v4, v);
}

// vector.cpp:517-536 : plain C on x,y,z
void mcemaths_mul_3(float* r4, float fac)
{
	r4[0] *= fac;
	r4[1] *= fac;
	r4[2] *= fac;
}

// vector.cpp:538-555 : fac = 1.0f / fac (true divide) then three multiplies
void mcemaths_div_3(float* r4, float fac)
{
	fac = 1.0f / fac;
	r4[0] *= fac;
	r4[1] *= fac;
	r4[2] *= fac;
}

// vector.cpp:557-567
void mcemaths_add_1to4(float* r4, float a)
{
	_mm_store_ps(r4, _mm_add_ps(_mm_load_ps(r4), _mm_set1_ps(a)));
}

// vector.cpp:569-579
void mcemaths_sub_4by1(float* r4, float a)
{
	_mm_store_ps(r4, _mm_sub_ps(_mm_load_ps(r4), _mm_set1_ps(a)));
}

// matrix.cpp:20-64 : the _MM_TRANSPOSE4_PS shuffle network
void mcemaths_mat4transpose(float* m44)
{
	__m128 r0 = _mm_load_ps(m44), r1 = _mm_load_ps(m44 + 4), r2 = _mm_load_ps(m44 + 8), r3 = _mm_load_ps(m44 + 12);
	__m128 t0 = _mm_shuffle_ps(r0, r1, 0x44);
	__m128 t2 = _mm_shuffle_ps(r0, r1, 0xee);
	__m128 t1 = _mm_shuffle_ps(r2, r3, 0x44);
	__m128 t3 = _mm_shuffle_ps(r2, r3, 0xee);
	_mm_store_ps(m44,      _mm_shuffle_ps(t0, t1, 0x88));
	_mm_store_ps(m44 + 4,  _mm_shuffle_ps(t0, t1, 0xdd));
	_mm_store_ps(m44 + 8,  _mm_shuffle_ps(t2, t3, 0x88));
	_mm_store_ps(m44 + 12, _mm_shuffle_ps(t2, t3, 0xdd));
}

// matrix.cpp:304-318
void mcemaths_mat4cpy(float* dest44, const float* src44)
{
	_mm_store_ps(dest44,      _mm_load_ps(src44));
	_mm_store_ps(dest44 + 4,  _mm_load_ps(src44 + 4));
	_mm_store_ps(dest44 + 8,  _mm_load_ps(src44 + 8));
	_mm_store_ps(dest44 + 12, _mm_load_ps(src44 + 12));
}

static inline __m128 m4v4(const float* trans44, __m128 v)
{
	// matrix.cpp:531-557 : ((x*c0 + y*c1) + z*c2) + w*c3, separate mulps/addps
	__m128 x = _mm_mul_ps(_mm_shuffle_ps(v, v, 0x00), _mm_load_ps(trans44));
	__m128 y = _mm_mul_ps(_mm_shuffle_ps(v, v, 0x55), _mm_load_ps(trans44 + 4));
	__m128 z = _mm_mul_ps(_mm_shuffle_ps(v, v, 0xaa), _mm_load_ps(trans44 + 8));
	__m128 w = _mm_mul_ps(_mm_shuffle_ps(v, v, 0xff), _mm_load_ps(trans44 + 12));
	x = _mm_add_ps(x, y);
	x = _mm_add_ps(x, z);
	x = _mm_add_ps(x, w);
	return x;
}

// matrix.cpp:515-558
void mcemaths_transform_m4v4(float* r4, const float* trans44, const float* v4)
{
	_mm_store_ps(r4, m4v4(trans44, _mm_load_ps(v4)));
}

// matrix.cpp:560-585
void mcemaths_transform_m4v4_ip(float* r4, const float* trans44)
{
	_mm_store_ps(r4, m4v4(trans44, _mm_load_ps(r4)));
}

// matrix.cpp:588-701 : d = l * r, one column of r at a time through the same mul/add order
void mcemaths_transform_m4m4(float* d44, const float* l44, const float* r44)
{
	__m128 c0 = m4v4(l44, _mm_load_ps(r44));
	__m128 c1 = m4v4(l44, _mm_load_ps(r44 + 4));
	__m128 c2 = m4v4(l44, _mm_load_ps(r44 + 8));
	__m128 c3 = m4v4(l44, _mm_load_ps(r44 + 12));
	_mm_store_ps(d44, c0);
	_mm_store_ps(d44 + 4, c1);
	_mm_store_ps(d44 + 8, c2);
	_mm_store_ps(d44 + 12, c3);
}

// matrix.cpp:992-1008 : columns T, B, N, 0
void mcemaths_make_tbn(float* m44, const float* tangent, const float* binormal, const float* normal)
{
	__m128 t = _mm_load_ps(tangent), b = _mm_load_ps(binormal), n = _mm_load_ps(normal);
	_mm_store_ps(m44, t);
	_mm_store_ps(m44 + 4, b);
	_mm_store_ps(m44 + 8, n);
	_mm_store_ps(m44 + 12, _mm_setzero_ps());
}

// quatern.cpp:9-17
void mcemaths_quatcpy(float* q_d, const float* q_s)
{
	_mm_store_ps(q_d, _mm_load_ps(q_s));
}

} // extern "C"
