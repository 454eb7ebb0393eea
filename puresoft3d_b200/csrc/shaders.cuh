// shaders.cuh — the reference's shader ("processor") interface re-expressed as device functors.
//
// In Puresoft3D a programme is three C++ objects with virtual methods (src/puresoft3d/proc.h:8-71):
//   PuresoftVertexProcessor::process(in.data[16]) -> position[4] + an opaque "user data" blob of float4 varyings
//   PuresoftInterpolationProcessor: interpolateByContributes / calcStep / correctInterpolation / stepForward
//   PuresoftFragmentProcessor::process(varyings) -> FragmentProcessorOutput::discard | write | write4
// Here each is a struct with the same method names, inlined into the kernels through the Programme<> template, so a
// maintainer porting a shader copies its body method by method. Uniforms are read by slot number from the block
// latched at draw time (the reference latches pointers in preprocess(), e.g. tex1light1.cpp:15-20,145-150); textures
// are the TexDesc bound for the slots named in TEX_SLOTS.
//
// All arithmetic goes through exact_math.cuh in the reference's operation order.
#pragma once
#include "device_types.cuh"

struct VertexProcessorInput { const uint8_t* data[16]; };          // proc.h:15-18
template<int NV> struct VertexProcessorOutput { F4 position; F4 user[NV > 0 ? NV : 1]; }; // proc.h:20-24

struct FragmentProcessorOutput                                     // proc.h:51-63 + FBOBridge (fragthrd.cpp:7-110)
{
	bool discarded, wrote, blendable;
	uint32_t bgra;
	PS_D void discard() { discarded = true; }
	PS_D void write(uint32_t c) { wrote = true; blendable = false; bgra = c; }   // FBOBridge::write: never blends
	PS_D void write4(uint32_t c) { wrote = true; blendable = true; bgra = c; }   // FBOBridge::write4: blend4 under ALPHABLEND
};

// fbo.cpp:208-229 blend4 : per channel dst + ((sat_u16(src - dst) * srcA) >> 8), 16-bit lanes, unsigned-saturating pack
PS_D uint32_t blend4(uint32_t src, uint32_t dst)
{
	const uint32_t a = src >> 24;
	uint32_t out = 0;
#pragma unroll
	for(int ch = 0; ch < 4; ch++)
	{
		const uint32_t s = (src >> (8 * ch)) & 0xff, d = (dst >> (8 * ch)) & 0xff;
		const uint32_t diff = s > d ? s - d : 0;           // _mm_subs_pu16
		const uint32_t prod = (diff * a) & 0xffff;          // _mm_mullo_pi16
		const uint32_t sum = ((prod >> 8) + d) & 0xffff;    // _mm_srli_pi16 + _mm_add_pi16
		const int ssum = (int)(short)sum;                   // _mm_packs_pu16 saturates a SIGNED 16-bit value to 0..255
		const uint32_t byte = ssum < 0 ? 0u : (ssum > 255 ? 255u : (uint32_t)ssum);
		out |= byte << (8 * ch);
	}
	return out;
}

// Attribute fetch. Plain (generic) loads: in.data[] points either into the VBO in global memory or into the copy of the
// block's vertex range that geom_setup staged in shared memory with one bulk copy per slot.
PS_D F4 ldF4(const uint8_t* p)
{
	if(0 == ((uintptr_t)p & 15))
	{
		const float4 v = *(const float4*)p;
		return f4(v.x, v.y, v.z, v.w);
	}
	const float* q = (const float*)p;
	return f4(q[0], q[1], q[2], q[3]);
}
PS_D void ldF2(const uint8_t* p, float& a, float& b)
{
	if(0 == ((uintptr_t)p & 7))
	{
		const float2 v = *(const float2*)p;
		a = v.x; b = v.y;
		return;
	}
	a = ((const float*)p)[0]; b = ((const float*)p)[1];
}

// ---- PuresoftFBO random access + the three samplers ---------------------------------------------------------------

// fbo.cpp:552-594 clampCoord
PS_D void clampCoord(const TexDesc& t, int& row, int& col)
{
	const int maxRow = t.height - 1, maxCol = t.width - 1;
	if(0 == t.wrap)
	{
		if(row > maxRow) row = maxRow; else if(row < 0) row = 0;
		if(col > maxCol) col = maxCol; else if(col < 0) col = 0;
	}
	else
	{
		row = maxRow ? row % maxRow : 0; if(row < 0) row += maxRow; // WRAP is modulo (size-1): fbo.cpp:582-590
		col = maxCol ? col % maxCol : 0; if(col < 0) col += maxCol;
	}
}
// fbo.cpp:287-291 directRead4
PS_D uint32_t directRead4(const TexDesc& t, int layer, int row, int col)
{
	clampCoord(t, row, col);
	return __ldg((const uint32_t*)(t.layer[layer] + (size_t)row * t.scanline) + col);
}

struct PuresoftSampler2D
{
	// samplr2d.cpp:19-25 — nearest; row from v, column from u; +0.5f then (unsigned int)
	PS_D static uint32_t get4(const TexDesc& t, float u, float v)
	{
		if(t.filter) return bilinear4(t, u, v);
		int row = cvtu(fadd(fmul((float)t.height, v), 0.5f));
		int col = cvtu(fadd(fmul((float)t.width, u), 0.5f));
		return directRead4(t, 0, row, col);
	}
	// EXTENSION (include/ps3d.h, ps3d_texture_set_filter): texel i at u = i / width, four taps through clampCoord,
	// per channel lerp(lerp(c00, c10, fx), lerp(c01, c11, fx), fy), lerp(a, b, t) = a + (b - a) * t, (int)(x + 0.5f)
	PS_D static float lerp1(float a, float b, float t) { return fadd(a, fmul(fsub(b, a), t)); }
	static __device__ __noinline__ uint32_t bilinear4(const TexDesc& t, float u, float v)
	{
		const float x = fmul((float)t.width, u), y = fmul((float)t.height, v);
		const float x0 = floorf(x), y0 = floorf(y);
		const float fx = fsub(x, x0), fy = fsub(y, y0);
		const int col = cvtt(x0), row = cvtt(y0);
		const uint32_t c00 = directRead4(t, 0, row, col), c10 = directRead4(t, 0, row, col + 1);
		const uint32_t c01 = directRead4(t, 0, row + 1, col), c11 = directRead4(t, 0, row + 1, col + 1);
		uint32_t out = 0;
#pragma unroll
		for(int ch = 0; ch < 4; ch++)
		{
			const float a = (float)((c00 >> (8 * ch)) & 0xff), b = (float)((c10 >> (8 * ch)) & 0xff);
			const float c = (float)((c01 >> (8 * ch)) & 0xff), d = (float)((c11 >> (8 * ch)) & 0xff);
			const float r = lerp1(lerp1(a, b, fx), lerp1(c, d, fx), fy);
			int q = cvtt(fadd(r, 0.5f));
			q = q < 0 ? 0 : (q > 255 ? 255 : q);
			out |= (uint32_t)q << (8 * ch);
		}
		return out;
	}
};

struct PuresoftSamplerCube
{
	// samplrcube.cpp:29-97 — GL face table, except the Y-major branch picks +-Y by the sign of Z (:83-94)
	PS_D static int texcoordFromDirection(float& S, float& T, F4 d)
	{
		const float X = d.x, Y = d.y, Z = d.z;
		const float aX = fabsf(X), aY = fabsf(Y), aZ = fabsf(Z);
		if(aX > aY)
		{
			if(aX > aZ)
			{
				if(X > 0) { S = fdiv(fadd(fdiv(-Z, aX), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aX), 1.0f), 2.0f); return 0; }
				else      { S = fdiv(fadd(fdiv( Z, aX), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aX), 1.0f), 2.0f); return 1; }
			}
			else
			{
				if(Z > 0) { S = fdiv(fadd(fdiv( X, aZ), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aZ), 1.0f), 2.0f); return 4; }
				else      { S = fdiv(fadd(fdiv(-X, aZ), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aZ), 1.0f), 2.0f); return 5; }
			}
		}
		else
		{
			if(aZ > aY)
			{
				if(Z > 0) { S = fdiv(fadd(fdiv( X, aZ), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aZ), 1.0f), 2.0f); return 4; }
				else      { S = fdiv(fadd(fdiv(-X, aZ), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Y, aZ), 1.0f), 2.0f); return 5; }
			}
			else
			{
				if(Z > 0) { S = fdiv(fadd(fdiv( X, aY), 1.0f), 2.0f); T = fdiv(fadd(fdiv( Z, aY), 1.0f), 2.0f); return 2; }
				else      { S = fdiv(fadd(fdiv( X, aY), 1.0f), 2.0f); T = fdiv(fadd(fdiv(-Z, aY), 1.0f), 2.0f); return 3; }
			}
		}
	}
	// samplrcube.cpp:119-127 — S addresses the ROW and T the COLUMN
	PS_D static uint32_t get4(const TexDesc& t, F4 direction)
	{
		float S, T;
		int layer = texcoordFromDirection(S, T, direction);
		if(layer >= t.nLayers) layer = 0; // the reference would dereference a NULL layer
		int row = cvtu(fadd(fmul((float)t.height, S), 0.5f));
		int col = cvtu(fadd(fmul((float)t.width, T), 0.5f));
		return directRead4(t, layer, row, col);
	}
};

struct PuresoftSamplerProjection
{
	// samplrproj.cpp:4-38 — tc = proj * rcpps(w); 4 Poisson taps; the X disk offsets v (row), the Y disk offsets u (col);
	// truncation without +0.5; each tap whose stored depth < tc.z costs 0.2
	PS_D static float get(const TexDesc& t, F4 projection, const ApproxTables& ap)
	{
		const float PX[4] = { -0.94201624f, 0.94558609f, -0.094184101f, 0.34495938f };
		const float PY[4] = { -0.39906216f, -0.76890725f, -0.92938870f, 0.29387760f };
		F4 tc = f4divs(projection, projection.w, ap);
		float shadowFactor = 1.0f;
#pragma unroll
		for(int i = 0; i < 4; i++)
		{
			int y = cvtu(fmul((float)t.height, fadd(tc.y, fdiv(PX[i], 500.0f))));
			int x = cvtu(fmul((float)t.width, fadd(tc.x, fdiv(PY[i], 500.0f))));
			float depthInShadowMap = __uint_as_float(directRead4(t, 0, y, x));
			if(depthInShadowMap < tc.z) shadowFactor = fsub(shadowFactor, 0.2f);
		}
		return shadowFactor;
	}
};

// ---- the interpolation processor: identical across all reference shaders up to the field list ----------------------

template<int N> struct InterpolationProcessorVec4
{
	static constexpr int NV = N;
	// e.g. tex1light1.cpp:60-91 : per lane ((v0*c0) + (v1*c1)) + (v2*c2)
	PS_D static void interpolateByContributes(F4* out, const F4* v0, const F4* v1, const F4* v2, float c0, float c1, float c2)
	{
#pragma unroll
		for(int k = 0; k < N; k++)
			out[k] = f4add(f4add(f4muls(v0[k], c0), f4muls(v1[k], c1)), f4muls(v2[k], c2));
	}
	// tex1light1.cpp:93-107 : (end - start) * (1.0f / stepCount), 1 when stepCount == 0
	PS_D static float reciprocalStepCount(int stepCount) { return 0 == stepCount ? 1.0f : fdiv(1.0f, (float)stepCount); }
	PS_D static void calcStepR(F4* step, const F4* start, const F4* end, float reciprocalStepCount)
	{
#pragma unroll
		for(int k = 0; k < N; k++)
			step[k] = f4muls(f4sub(end[k], start[k]), reciprocalStepCount);
	}
	PS_D static void calcStep(F4* step, const F4* start, const F4* end, int stepCount)
	{
		calcStepR(step, start, end, reciprocalStepCount(stepCount));
	}
	// tex1light1.cpp:109-117
	PS_D static void correctInterpolation(F4* out, const F4* start, float correctionFactor2)
	{
#pragma unroll
		for(int k = 0; k < N; k++)
			out[k] = f4muls(start[k], correctionFactor2);
	}
	// tex1light1.cpp:119-135 : += step, or += step * n as a multiply then an add (mcemaths_step_3_4_ip, vector.cpp:70-83)
	PS_D static void stepForward(F4* start, const F4* step, int stepCount)
	{
		if(1 == stepCount)
		{
#pragma unroll
			for(int k = 0; k < N; k++) start[k] = f4add(start[k], step[k]);
		}
		else
		{
			const float n = (float)stepCount;
#pragma unroll
			for(int k = 0; k < N; k++) start[k] = f4add(start[k], f4muls(step[k], n));
		}
	}
};

// ---- shared fragment code -------------------------------------------------------------------------------------------

PS_D F4 unpackBGRA(uint32_t c) // e.g. tex1light1.cpp:158-161
{
	return f4((float)(c & 0xff), (float)((c >> 8) & 0xff), (float)((c >> 16) & 0xff), (float)(c >> 24));
}
PS_D uint32_t packBGRtrunc(F4 c) // tex1light1.cpp:186-190: (unsigned char) truncation, a = 0
{
	return (uint32_t)(cvtt(c.x) & 0xff) | ((uint32_t)(cvtt(c.y) & 0xff) << 8) | ((uint32_t)(cvtt(c.z) & 0xff) << 16);
}
// the Blinn-Phong tail of DEF01/02/03 (tex1light1.cpp:163-190)
PS_D uint32_t blinnPhong(const DrawParams& P, F4 colour, F4 worldPos, F4 normal)
{
	const F4 lightPos = f4(P.u[7][0], P.u[7][1], P.u[7][2], P.u[7][3]);
	const F4 cameraPos = f4(P.u[8][0], P.u[8][1], P.u[8][2], P.u[8][3]);
	F4 L = f4sub(lightPos, worldPos);
	float distance = f4len(L);
	L = f4divs(L, distance, P.approx);
	F4 E = f4norm(f4sub(cameraPos, worldPos), P.approx);
	F4 H = f4norm(f4add(E, L), P.approx);
	float lambert = f4dot(L, normal);
	float specular = f4dot(H, normal);
	specular = specular < 0 ? 0 : specular;
	specular = opt_pow(specular, 50);
	colour = f4adds(colour, fmul(255.0f, specular));
	colour = f4muls(colour, lambert);
	colour = f4clamp(colour, 0, 255.0f);
	return packBGRtrunc(colour);
}

// ---- DEF01: textured Blinn-Phong (tex1light1.cpp) — varyings: normal, worldPos, texcoord ---------------------------

struct VertexProcesserDEF01
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 3) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1u << 3) | (1u << 4) | (1u << 5);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<3>& out, const DrawParams& P) // :22-41
	{
		F4 worldPos = m4v4(P.u[4], ldF4(in.data[0]));
		out.position = m4v4(P.u[3], worldPos);
		worldPos.w = 0;
		out.user[1] = worldPos;
		out.user[0] = m4v4(P.u[5], ldF4(in.data[3]));
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[2] = f4(tu, tv, 0, 0);
	}
};
struct FragmentProcessorDEF01
{
	static constexpr uint64_t UNIFORMS = (1u << 7) | (1u << 8) | (1u << 9);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 1;
	__host__ __device__ static constexpr int texSlot(int i) { return 9; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P) // :152-193
	{
		F4 colour = unpackBGRA(PuresoftSampler2D::get4(P.tex[0], in[2].x, in[2].y));
		out.write(blinnPhong(P, colour, in[1], in[0]));
	}
};

// ---- DEF02: vertex-colour Blinn-Phong (colr1light1.cpp) — slots 0 pos, 1 normal, 2 colour --------------------------

struct VertexProcesserDEF02
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 1) | (1u << 2);
	static constexpr uint64_t UNIFORMS = (1u << 3) | (1u << 4) | (1u << 5);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<3>& out, const DrawParams& P) // :21-39
	{
		F4 worldPos = m4v4(P.u[4], ldF4(in.data[0]));
		out.position = m4v4(P.u[3], worldPos);
		worldPos.w = 0;
		out.user[1] = worldPos;
		out.user[0] = m4v4(P.u[5], ldF4(in.data[1]));
		out.user[2] = ldF4(in.data[2]);
	}
};
struct FragmentProcessorDEF02
{
	static constexpr uint64_t UNIFORMS = (1u << 7) | (1u << 8);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 0;
	__host__ __device__ static constexpr int texSlot(int i) { return -1; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P) // :150-190
	{
		out.write(blinnPhong(P, in[2], in[1], in[0]));
	}
};

// ---- DEF03: + tangent-space normal map (tex1bump1light1.cpp) — varyings: tangent, binormal, normal, worldPos, uv ----

struct VertexProcesserDEF03
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1u << 3) | (1u << 4) | (1u << 5);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<5>& out, const DrawParams& P) // :22-45
	{
		F4 worldPos = m4v4(P.u[4], ldF4(in.data[0]));
		out.position = m4v4(P.u[3], worldPos);
		worldPos.w = 0;
		out.user[3] = worldPos;
		out.user[0] = m4v4(P.u[5], ldF4(in.data[1]));
		out.user[1] = m4v4(P.u[5], ldF4(in.data[2]));
		out.user[2] = m4v4(P.u[5], ldF4(in.data[3]));
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[4] = f4(tu, tv, 0, 0);
	}
};
struct FragmentProcessorDEF03
{
	static constexpr uint64_t UNIFORMS = (1u << 7) | (1u << 8) | (1u << 9) | (1u << 10);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 2;
	__host__ __device__ static constexpr int texSlot(int i) { return i == 0 ? 9 : 10; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P) // :180-239
	{
		F4 colour = unpackBGRA(PuresoftSampler2D::get4(P.tex[0], in[4].x, in[4].y));
		uint32_t nb = PuresoftSampler2D::get4(P.tex[1], in[4].x, in[4].y);
		F4 bump = f4((float)((nb >> 16) & 0xff), (float)((nb >> 8) & 0xff), (float)(nb & 0xff), 0); // r,g,b -> x,y,z (:193-196)
		bump = f4divs(bump, 255.0f, P.approx);
		bump = f4muls(bump, 2.0f);
		bump = f4subs(bump, 1.0f);
		// mcemaths_make_tbn (matrix.cpp:992-1008): columns T, B, N, 0 ; then M*v in m4v4's order
		const F4 T = in[0], B = in[1], N = in[2];
		F4 r;
		r.x = fadd(fadd(fadd(fmul(bump.x, T.x), fmul(bump.y, B.x)), fmul(bump.z, N.x)), fmul(bump.w, 0.0f));
		r.y = fadd(fadd(fadd(fmul(bump.x, T.y), fmul(bump.y, B.y)), fmul(bump.z, N.y)), fmul(bump.w, 0.0f));
		r.z = fadd(fadd(fadd(fmul(bump.x, T.z), fmul(bump.y, B.z)), fmul(bump.z, N.z)), fmul(bump.w, 0.0f));
		r.w = fadd(fadd(fadd(fmul(bump.x, T.w), fmul(bump.y, B.w)), fmul(bump.z, N.w)), fmul(bump.w, 0.0f));
		bump = f4norm(r, P.approx);
		out.write(blinnPhong(P, colour, in[3], bump));
	}
};

// ---- DEF04: cube-map skybox (skybox.cpp) — varying: direction -------------------------------------------------------

struct VertexProcesserDEF04
{
	static constexpr uint32_t SLOTS = (1u << 0);
	static constexpr uint64_t UNIFORMS = (1u << 1);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<1>& out, const DrawParams& P) // :23-44
	{
		const F4 position = ldF4(in.data[0]);
		F4 d = f4sub(position, f4(0, 0, 1.0f, 0));
		d.w = 0;
		d = f4norm(d, P.approx);
		// inversedView = transpose(V) with its last column zeroed (:34-37)
		const float* V = P.u[1];
		float inv[16];
#pragma unroll
		for(int c = 0; c < 4; c++)
#pragma unroll
			for(int r = 0; r < 4; r++) inv[c * 4 + r] = V[r * 4 + c];
		inv[12] = inv[13] = inv[14] = inv[15] = 0;
		out.user[0] = m4v4(inv, d);
		out.position = position;
	}
};
struct FragmentProcessorDEF04
{
	static constexpr uint64_t UNIFORMS = (1u << 2);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true; 
	static constexpr int NTEX = 1;
	__host__ __device__ static constexpr int texSlot(int i) { return 2; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P) // :127-134
	{
		out.write4(PuresoftSamplerCube::get4(P.tex[0], in[0]));
	}
};

// ---- DEF05: depth only, 0.9 shrink (shadow.cpp) ---------------------------------------------------------------------

struct VertexProcesserDEF05
{
	static constexpr uint32_t SLOTS = (1u << 0);
	static constexpr uint64_t UNIFORMS = (1u << 3) | (1u << 4);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<0>& out, const DrawParams& P) // :21-29
	{
		F4 p = ldF4(in.data[0]);
		p.x = fmul(p.x, 0.9f); p.y = fmul(p.y, 0.9f); p.z = fmul(p.z, 0.9f); // mcemaths_mul_3: xyz only
		// pvm = PV * M recomputed per vertex exactly like the reference (mcemaths_transform_m4m4, matrix.cpp:588-701)
		float pvm[16];
#pragma unroll
		for(int c = 0; c < 4; c++)
		{
			F4 col = m4v4(P.u[3], f4(P.u[4][c * 4], P.u[4][c * 4 + 1], P.u[4][c * 4 + 2], P.u[4][c * 4 + 3]));
			pvm[c * 4] = col.x; pvm[c * 4 + 1] = col.y; pvm[c * 4 + 2] = col.z; pvm[c * 4 + 3] = col.w;
		}
		out.position = m4v4(pvm, p);
	}
};
struct FragmentProcessorDEF05
{
	static constexpr uint64_t UNIFORMS = 0;
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 0;
	__host__ __device__ static constexpr int texSlot(int i) { return -1; }
	PS_D static void process(const F4*, FragmentProcessorOutput&, const DrawParams&) {} // shadow.cpp:76-77
};

// =====================================================================================================================
// demo 1 — src/test/testproc.cpp (earth + moon + cloud layer, one shadow map). Uniform slots: 3 PV, 4 M, 5 Mrot, 7 light,
// 8 camera, 9 diffuse, 10 bump, 11 specular map, 12 night map, 15 shadow map, 16 shadow PV (src/test/testproc.cpp:29-35,198-206)
// =====================================================================================================================

PS_D F4 uvec(const DrawParams& P, int slot) { return f4(P.u[slot][0], P.u[slot][1], P.u[slot][2], P.u[slot][3]); }

// the part FP_Earth / FP_Satellite share with DEF03 (testproc.cpp:240-262): bump normal through the TBN, L by length + rcp, E, H
PS_D void planetLighting(const DrawParams& P, uint32_t bumpTexel, F4 T, F4 B, F4 N, F4 worldPos, float& lambert, float& specular)
{
	F4 bump = f4((float)((bumpTexel >> 16) & 0xff), (float)((bumpTexel >> 8) & 0xff), (float)(bumpTexel & 0xff), 0);
	bump = f4divs(bump, 255.0f, P.approx);
	bump = f4muls(bump, 2.0f);
	bump = f4subs(bump, 1.0f);
	F4 r; // mcemaths_make_tbn + transform_m4v4_ip: columns T, B, N, 0
	r.x = fadd(fadd(fadd(fmul(bump.x, T.x), fmul(bump.y, B.x)), fmul(bump.z, N.x)), fmul(bump.w, 0.0f));
	r.y = fadd(fadd(fadd(fmul(bump.x, T.y), fmul(bump.y, B.y)), fmul(bump.z, N.y)), fmul(bump.w, 0.0f));
	r.z = fadd(fadd(fadd(fmul(bump.x, T.z), fmul(bump.y, B.z)), fmul(bump.z, N.z)), fmul(bump.w, 0.0f));
	r.w = fadd(fadd(fadd(fmul(bump.x, T.w), fmul(bump.y, B.w)), fmul(bump.z, N.w)), fmul(bump.w, 0.0f));
	bump = f4norm(r, P.approx);
	F4 L = f4sub(uvec(P, 7), worldPos);
	const float distance = f4len(L);
	L = f4divs(L, distance, P.approx);
	const F4 E = f4norm(f4sub(uvec(P, 8), worldPos), P.approx);
	const F4 H = f4norm(f4add(E, L), P.approx);
	lambert = f4dot(L, bump);
	specular = f4dot(H, bump);
	specular = specular < 0 ? 0 : specular;
	specular = opt_pow(specular, 50);
}

struct VP_Planet // testproc.cpp:37-74 — varyings: tangent, binormal, normal, worldPos, texcoord, shadowcoord
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1ull << 3) | (1ull << 4) | (1ull << 5) | (1ull << 16);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<6>& out, const DrawParams& P)
	{
		F4 worldPos = m4v4(P.u[4], ldF4(in.data[0]));
		out.user[5] = m4v4(P.u[16], worldPos);
		out.position = m4v4(P.u[3], worldPos);
		worldPos.w = 0;
		out.user[3] = worldPos;
		out.user[0] = m4v4(P.u[5], ldF4(in.data[1]));
		out.user[1] = m4v4(P.u[5], ldF4(in.data[2]));
		out.user[2] = m4v4(P.u[5], ldF4(in.data[3]));
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[4] = f4(tu, tv, 0, 0);   // [2],[3] are never written by the reference (stale buffer contents), never read either
	}
};
// IP_Planet::stepForward advances tangent AND binormal by the NORMAL's step (testproc.cpp:178-188) — replicated
struct IP_Planet : InterpolationProcessorVec4<6>
{
	PS_D static void stepForward(F4* start, const F4* step, int stepCount)
	{
		if(1 == stepCount)
		{
			start[0] = f4add(start[0], step[2]); start[1] = f4add(start[1], step[2]); start[2] = f4add(start[2], step[2]);
			start[3] = f4add(start[3], step[3]); start[4] = f4add(start[4], step[4]); start[5] = f4add(start[5], step[5]);
		}
		else
		{
			const float n = (float)stepCount;
			start[0] = f4add(start[0], f4muls(step[2], n)); start[1] = f4add(start[1], f4muls(step[2], n)); start[2] = f4add(start[2], f4muls(step[2], n));
			start[3] = f4add(start[3], f4muls(step[3], n)); start[4] = f4add(start[4], f4muls(step[4], n)); start[5] = f4add(start[5], f4muls(step[5], n));
		}
	}
};
struct FP_Earth // testproc.cpp:198-288
{
	static constexpr uint64_t UNIFORMS = (1ull << 7) | (1ull << 8) | (1ull << 9) | (1ull << 10) | (1ull << 11) | (1ull << 12) | (1ull << 15);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 5;
	__host__ __device__ static constexpr int texSlot(int i) { return i == 0 ? 9 : (i == 1 ? 10 : (i == 2 ? 11 : (i == 3 ? 12 : 15))); }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		F4 colour = unpackBGRA(PuresoftSampler2D::get4(P.tex[0], in[4].x, in[4].y));
		const F4 night = unpackBGRA(PuresoftSampler2D::get4(P.tex[3], in[4].x, in[4].y));
		const uint32_t nb = PuresoftSampler2D::get4(P.tex[1], in[4].x, in[4].y);
		const float shadowFactor = PuresoftSamplerProjection::get(P.tex[4], in[5], P.approx);
		const uint32_t sp = PuresoftSampler2D::get4(P.tex[2], in[4].x, in[4].y);
		const float specularControl = fdiv((float)((sp >> 16) & 0xff), 255.0f);
		float lambert, specular;
		planetLighting(P, nb, in[0], in[1], in[2], in[3], lambert, specular);
		specular = fmul(specular, specularControl);
		colour = f4adds(colour, fmul(255.0f, specular));
		colour = f4muls(colour, lambert);
		if(lambert < 0.2f) colour = f4add(colour, night);
		colour = f4muls(colour, shadowFactor);
		colour = f4clamp(colour, 0, 255.0f);
		out.write4(packBGRtrunc(colour));
	}
};
struct FP_Satellite // testproc.cpp:290-361
{
	static constexpr uint64_t UNIFORMS = (1ull << 7) | (1ull << 8) | (1ull << 9) | (1ull << 10) | (1ull << 15);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 3;
	__host__ __device__ static constexpr int texSlot(int i) { return i == 0 ? 9 : (i == 1 ? 10 : 15); }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		F4 colour = unpackBGRA(PuresoftSampler2D::get4(P.tex[0], in[4].x, in[4].y));
		const uint32_t nb = PuresoftSampler2D::get4(P.tex[1], in[4].x, in[4].y);
		const float shadowFactor = PuresoftSamplerProjection::get(P.tex[2], in[5], P.approx);
		float lambert, specular;
		planetLighting(P, nb, in[0], in[1], in[2], in[3], lambert, specular);
		colour = f4adds(colour, fmul(255.0f, specular));
		colour = f4muls(colour, lambert);
		colour = f4muls(colour, shadowFactor);
		colour = f4clamp(colour, 0, 255.0f);
		out.write4(packBGRtrunc(colour));
	}
};

// cvtps2dq + packusdw + packuswb on one lane (testproc.cpp:565-571): round to nearest even (INT_MIN when out of range), saturate
// the SIGNED 32-bit value to 0..65535, then the SIGNED 16-bit reading of that to 0..255 — so 32768..65535 become 0
PS_D uint32_t packRneSat(float f)
{
	int v;
	if(!(f >= -2147483648.0f && f < 2147483648.0f)) v = (int)0x80000000u; else v = __float2int_rn(f);
	const int u16 = v < 0 ? 0 : (v > 65535 ? 65535 : v);
	const int s16 = (int)(short)u16;
	return (uint32_t)(s16 < 0 ? 0 : (s16 > 255 ? 255 : s16));
}

struct VP_Cloud // testproc.cpp:392-425 — varyings: normal, worldPos, texcoord, shadowcoord
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 3) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1ull << 3) | (1ull << 4) | (1ull << 5) | (1ull << 16);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<4>& out, const DrawParams& P)
	{
		F4 worldPos = m4v4(P.u[4], ldF4(in.data[0]));
		out.user[3] = m4v4(P.u[16], worldPos);
		out.position = m4v4(P.u[3], worldPos);
		worldPos.w = 0;
		out.user[1] = worldPos;
		out.user[0] = m4v4(P.u[5], ldF4(in.data[3]));
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[2] = f4(tu, tv, 0, 0);
	}
};
struct FP_Cloud // testproc.cpp:531-574 — alpha = the texture's red channel; blended by the caller's ALPHABLEND (write4)
{
	static constexpr uint64_t UNIFORMS = (1ull << 7) | (1ull << 8) | (1ull << 9) | (1ull << 15);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 2;
	__host__ __device__ static constexpr int texSlot(int i) { return i == 0 ? 9 : 15; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		const uint32_t tex = PuresoftSampler2D::get4(P.tex[0], in[2].x, in[2].y);
		F4 cloud = f4(255.0f, 255.0f, 255.0f, (float)((tex >> 16) & 0xff));
		const float shadowFactor = PuresoftSamplerProjection::get(P.tex[1], in[3], P.approx);
		F4 L = f4sub(uvec(P, 7), in[1]);
		const float distance = f4len(L);
		L = f4divs(L, distance, P.approx);
		// E and H are computed by the reference and never used (:551-556)
		const float lambert = fmul(2.0f, f4dot(L, in[0]));
		const float k = fmul(lambert, shadowFactor);
		cloud.x = fmul(cloud.x, k); cloud.y = fmul(cloud.y, k); cloud.z = fmul(cloud.z, k);   // mcemaths_mul_3: xyz only
		out.write4(packRneSat(cloud.x) | (packRneSat(cloud.y) << 8) | (packRneSat(cloud.z) << 16) | (packRneSat(cloud.w) << 24));
	}
};

struct VP_CloudShadow // testproc.cpp:595-610 — 0.95 shrink, pvm = PV * M per vertex; varying: texcoord
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1ull << 3) | (1ull << 4);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<1>& out, const DrawParams& P)
	{
		F4 p = ldF4(in.data[0]);
		p.x = fmul(p.x, 0.95f); p.y = fmul(p.y, 0.95f); p.z = fmul(p.z, 0.95f);
		float pvm[16];
#pragma unroll
		for(int c = 0; c < 4; c++)
		{
			F4 col = m4v4(P.u[3], f4(P.u[4][c * 4], P.u[4][c * 4 + 1], P.u[4][c * 4 + 2], P.u[4][c * 4 + 3]));
			pvm[c * 4] = col.x; pvm[c * 4 + 1] = col.y; pvm[c * 4 + 2] = col.z; pvm[c * 4 + 3] = col.w;
		}
		out.position = m4v4(pvm, p);
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[0] = f4(tu, tv, 0, 0);
	}
};
struct FP_CloudShadow // testproc.cpp:671-685 — the only functor of the reference that discards: thin cloud casts no shadow
{
	static constexpr uint64_t UNIFORMS = (1ull << 9);
	static constexpr bool MAY_DISCARD = true;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 1;
	__host__ __device__ static constexpr int texSlot(int i) { return 9; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		const uint32_t tex = PuresoftSampler2D::get4(P.tex[0], in[0].x, in[0].y);
		if(((tex >> 16) & 0xff) < 150) out.discard();
	}
};

// =====================================================================================================================
// demo 2 — src/test2/testproc.cpp (desk scene, spot light, shadow map). Uniform slots (src/test2/testproc.h:4-37): 0 M, 1 Mrot,
// 4 PV, 5 PVM, 6 shadow PV, 20 light position, 21 light direction, 22 camera, 23 shadow map, 30 ambient, 31 diffuse colour,
// 32 specular colour, 33 specular exponent, 40 diffuse texture
// =====================================================================================================================

struct VP_PositionOnly // testproc.cpp:17-22
{
	static constexpr uint32_t SLOTS = (1u << 0);
	static constexpr uint64_t UNIFORMS = (1ull << 5);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<0>& out, const DrawParams& P)
	{
		out.position = m4v4(P.u[5], ldF4(in.data[0]));
	}
};
struct VP_Shadow // testproc.cpp:510-516 — 0.9 shrink (xyz), then PVM
{
	static constexpr uint32_t SLOTS = (1u << 0);
	static constexpr uint64_t UNIFORMS = (1ull << 5);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<0>& out, const DrawParams& P)
	{
		F4 p = ldF4(in.data[0]);
		p.x = fmul(p.x, 0.9f); p.y = fmul(p.y, 0.9f); p.z = fmul(p.z, 0.9f);
		out.position = m4v4(P.u[5], p);
	}
};
// IP_Null (testproc.cpp:26-45) declares 16 bytes of user data and never touches them: no varying reaches the fragment functor
struct FP_Null // testproc.cpp:520-524
{
	static constexpr bool NOOP = true;   // process() does nothing: a depth-only pass needs no shade kernel and no survivor stream
	static constexpr uint64_t UNIFORMS = 0;
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 0;
	__host__ __device__ static constexpr int texSlot(int i) { return -1; }
	PS_D static void process(const F4*, FragmentProcessorOutput&, const DrawParams&) {}
};
struct FP_SingleColourNoLighting // testproc.cpp:49-67
{
	static constexpr uint64_t UNIFORMS = (1ull << 31);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 0;
	__host__ __device__ static constexpr int texSlot(int i) { return -1; }
	PS_D static void process(const F4*, FragmentProcessorOutput& out, const DrawParams& P)
	{
		out.write4(packBGRtrunc(f4muls(uvec(P, 31), 255.0f)));
	}
};

static constexpr float kFieldOfLight = 6.283185f * (25.0f / 360.0f);   // testproc.cpp:9

// the lighting factors FP_SingleColour and FP_DiffuseOnly share (testproc.cpp:230-256 / 459-485): L, E, H by rsqrt, a spot
// cone through acosf / cosf / opt_pow(.,150), specular by opt_pow, both clamped to [0,1]
PS_D void spotFactors(const DrawParams& P, F4 worldPos, F4 normal, float& lambertOut, float& specularOut)
{
	const F4 L = f4norm(f4sub(uvec(P, 20), worldPos), P.approx);
	const F4 E = f4norm(f4sub(uvec(P, 22), worldPos), P.approx);
	const F4 H = f4norm(f4add(E, L), P.approx);
	float lambert = f4dot(L, normal);
	// <math.h> in C++ resolves acos(float) / cos(float) to the float overloads (acosf / cosf of the host's libm, correctly
	// rounded in all but rare cases): evaluated here in double and rounded once, which gives the same float except in those cases
	const float yawOfLight = (float)acos((double)f4dot(L, uvec(P, 21)));
	float cone = 1.0f;
	if(!(yawOfLight < kFieldOfLight))
		cone = opt_pow((float)cos((double)fsub(yawOfLight, kFieldOfLight)), 150);
	lambert = fmul(lambert, cone);
	float specular = opt_pow(f4dot(H, normal), (unsigned)cvtu(P.u[33][0]));
	lambertOut = clamp1(lambert, 0, 1.0f);
	specularOut = clamp1(specular, 0, 1.0f);
}

struct VP_SingleColour // testproc.cpp:94-120 — varyings: normal, worldPos, shadowcoord
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 3);
	static constexpr uint64_t UNIFORMS = (1ull << 0) | (1ull << 1) | (1ull << 5) | (1ull << 6);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<3>& out, const DrawParams& P)
	{
		const F4 position = ldF4(in.data[0]);
		F4 worldPos = m4v4(P.u[0], position);
		out.user[2] = m4v4(P.u[6], worldPos);
		worldPos.w = 0;
		out.user[1] = worldPos;
		out.position = m4v4(P.u[5], position);
		out.user[0] = m4v4(P.u[1], ldF4(in.data[3]));
	}
};
struct FP_SingleColour // testproc.cpp:209-287
{
	static constexpr uint64_t UNIFORMS = (1ull << 20) | (1ull << 21) | (1ull << 22) | (1ull << 23) | (1ull << 30) | (1ull << 31) | (1ull << 32) | (1ull << 33);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 1;
	__host__ __device__ static constexpr int texSlot(int i) { return 23; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		const float shadowFactor = PuresoftSamplerProjection::get(P.tex[0], in[2], P.approx);
		float lambert, specular;
		spotFactors(P, in[1], in[0], lambert, specular);
		const F4 diffuse = uvec(P, 31), specCol = uvec(P, 32), ambient = uvec(P, 30);
		F4 colour = f4muls(diffuse, lambert);
		F4 specularColour = f4(fmul(diffuse.x, specCol.x), fmul(diffuse.y, specCol.y), fmul(diffuse.z, specCol.z), fmul(diffuse.w, specCol.w));
		specularColour = f4muls(specularColour, specular);
		const F4 ambientColour = f4(fmul(diffuse.x, ambient.x), fmul(diffuse.y, ambient.y), fmul(diffuse.z, ambient.z), fmul(diffuse.w, ambient.w));
		colour = f4add(colour, specularColour);
		colour = f4muls(colour, shadowFactor);
		colour = f4add(colour, ambientColour);
		colour = f4clamp(colour, 0, 1.0f);
		colour = f4muls(colour, 255.0f);
		out.write4(packBGRtrunc(colour));
	}
};

struct VP_DiffuseOnly // testproc.cpp:310-342 — varyings: normal, worldPos, texcoord, shadowcoord
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 3) | (1u << 4);
	static constexpr uint64_t UNIFORMS = (1ull << 0) | (1ull << 1) | (1ull << 5) | (1ull << 6);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<4>& out, const DrawParams& P)
	{
		const F4 position = ldF4(in.data[0]);
		F4 worldPos = m4v4(P.u[0], position);
		out.user[3] = m4v4(P.u[6], worldPos);
		out.position = m4v4(P.u[5], position);
		worldPos.w = 0;
		out.user[1] = worldPos;
		out.user[0] = m4v4(P.u[1], ldF4(in.data[3]));
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[2] = f4(tu, tv, 0, 0);
	}
};
struct FP_DiffuseOnly // testproc.cpp:437-500
{
	static constexpr uint64_t UNIFORMS = (1ull << 20) | (1ull << 21) | (1ull << 22) | (1ull << 23) | (1ull << 30) | (1ull << 33) | (1ull << 40);
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true;
	static constexpr int NTEX = 2;
	__host__ __device__ static constexpr int texSlot(int i) { return i == 0 ? 23 : 40; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		const float shadowFactor = PuresoftSamplerProjection::get(P.tex[0], in[3], P.approx);
		F4 colour = unpackBGRA(PuresoftSampler2D::get4(P.tex[1], in[2].x, in[2].y));
		const F4 ambient = uvec(P, 30);
		const F4 ambientColour = f4(fmul(colour.x, ambient.x), fmul(colour.y, ambient.y), fmul(colour.z, ambient.z), fmul(colour.w, ambient.w));
		float lambert, specular;
		spotFactors(P, in[1], in[0], lambert, specular);
		colour = f4muls(colour, fmul(fadd(lambert, specular), shadowFactor));
		colour = f4add(colour, ambientColour);
		colour = f4clamp(colour, 0, 255.0f);
		out.write4(packBGRtrunc(colour));
	}
};

// ---- FLATID: parity-test functor, not in the reference: carries a per-triangle id colour to the pixel --------------

struct VertexProcesserFLATID
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 6);
	static constexpr uint64_t UNIFORMS = (1u << 3) | (1u << 4);
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<1>& out, const DrawParams& P)
	{
		out.position = m4v4(P.u[3], m4v4(P.u[4], ldF4(in.data[0])));
		out.user[0] = ldF4(in.data[6]);
	}
};
struct FragmentProcessorFLATID
{
	static constexpr uint64_t UNIFORMS = 0;
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = true; 
	static constexpr int NTEX = 0;
	__host__ __device__ static constexpr int texSlot(int i) { return -1; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams&)
	{
		uint32_t b = (uint32_t)(cvtt(fadd(in[0].x, 0.5f)) & 0xff), g = (uint32_t)(cvtt(fadd(in[0].y, 0.5f)) & 0xff);
		uint32_t r = (uint32_t)(cvtt(fadd(in[0].z, 0.5f)) & 0xff), a = (uint32_t)(cvtt(fadd(in[0].w, 0.5f)) & 0xff);
		out.write4(b | (g << 8) | (r << 16) | (a << 24));
	}
};

// ---- TEXPROBE: parity-test functor, not in the reference: the 2-D sampler's output, unlit, straight to the target ----
struct VertexProcesserTEXPROBE
{
	static constexpr uint32_t SLOTS = (1u << 0) | (1u << 4);
	static constexpr uint64_t UNIFORMS = 0;
	PS_D static void process(const VertexProcessorInput& in, VertexProcessorOutput<1>& out, const DrawParams&)
	{
		out.position = ldF4(in.data[0]);
		float tu, tv;
		ldF2(in.data[4], tu, tv);
		out.user[0] = f4(tu, tv, 0.0f, 0.0f);
	}
};
struct FragmentProcessorTEXPROBE
{
	static constexpr uint64_t UNIFORMS = 1u << 9;
	static constexpr bool MAY_DISCARD = false;
	static constexpr bool USES_WRITE4 = false;
	static constexpr int NTEX = 1;
	__host__ __device__ static constexpr int texSlot(int i) { return 9; }
	PS_D static void process(const F4* in, FragmentProcessorOutput& out, const DrawParams& P)
	{
		out.write(PuresoftSampler2D::get4(P.tex[0], in[0].x, in[0].y));
	}
};

// ---- programme = (V, I, F) ------------------------------------------------------------------------------------------

template<class VP, class IP, class FP> struct Programme
{
	typedef VP V;
	typedef IP I;
	typedef FP F;
	static constexpr int NV = IP::NV;
};

typedef Programme<VertexProcesserDEF01, InterpolationProcessorVec4<3>, FragmentProcessorDEF01> ProgDEF01;
typedef Programme<VertexProcesserDEF02, InterpolationProcessorVec4<3>, FragmentProcessorDEF02> ProgDEF02;
typedef Programme<VertexProcesserDEF03, InterpolationProcessorVec4<5>, FragmentProcessorDEF03> ProgDEF03;
typedef Programme<VertexProcesserDEF04, InterpolationProcessorVec4<1>, FragmentProcessorDEF04> ProgDEF04;
typedef Programme<VertexProcesserDEF05, InterpolationProcessorVec4<0>, FragmentProcessorDEF05> ProgDEF05;
typedef Programme<VertexProcesserFLATID, InterpolationProcessorVec4<1>, FragmentProcessorFLATID> ProgFLATID;
typedef Programme<VertexProcesserTEXPROBE, InterpolationProcessorVec4<1>, FragmentProcessorTEXPROBE> ProgTEXPROBE;
typedef Programme<VP_Planet, IP_Planet, FP_Earth> ProgEarth;
typedef Programme<VP_Planet, IP_Planet, FP_Satellite> ProgSatellite;
typedef Programme<VP_Cloud, InterpolationProcessorVec4<4>, FP_Cloud> ProgCloud;
typedef Programme<VP_CloudShadow, InterpolationProcessorVec4<1>, FP_CloudShadow> ProgCloudShadow;
typedef Programme<VP_PositionOnly, InterpolationProcessorVec4<0>, FP_SingleColourNoLighting> ProgPositionOnly;
typedef Programme<VP_SingleColour, InterpolationProcessorVec4<3>, FP_SingleColour> ProgSingleColour;
typedef Programme<VP_DiffuseOnly, InterpolationProcessorVec4<4>, FP_DiffuseOnly> ProgDiffuseOnly;
typedef Programme<VP_Shadow, InterpolationProcessorVec4<0>, FP_Null> ProgShadow2;
