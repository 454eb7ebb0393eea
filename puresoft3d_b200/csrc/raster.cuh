// raster.cuh — triangle setup and per-row span evaluation, bit-for-bit with PuresoftRasterizer
// (src/puresoft3d/rasterizer.cpp) and PuresoftInterpolater (interp.cpp).
//
// The reference walks a triangle once on the caller thread and leaves RESULT_ROW[firstRow..lastRow] behind
// (rasterizer.h:12-28). Here the same per-row arithmetic is a pure function of the 64-byte TriHeader, so any lane can
// evaluate any row of any triangle: geom_setup uses it to find the screen extent, the tile kernel to fill its spans.
// Quirks that define coverage and are therefore kept: rows are evaluated AT integer y, the first
// row is trunc(ymin) and may lie below the lowest vertex, spans are inclusive at both ends, flat edges are detected
// with exact float ==, |dy| < 1e-6 turns the divisor into FLT_MAX, the lower half overwrites the shared row, and a
// triangle whose clamped row range is a single row is dropped.
#pragma once
#include <float.h>
#include "device_types.cuh"

struct Edge { float dx, dy, x0, y0; };

// register-friendly v[i] for i in 0..2 (a dynamically indexed local array would live in local memory)
PS_D float sel3(const float* v, int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

// LineSegment ctor, rasterizer.cpp:49-64
PS_D Edge makeEdge(const float* vx, const float* vy, int i0, int i1)
{
	Edge e;
	e.x0 = sel3(vx, i0); e.y0 = sel3(vy, i0);
	e.dx = fsub(sel3(vx, i1), e.x0);
	e.dy = fsub(sel3(vy, i1), e.y0);
	if(fabsf(e.dy) < 0.000001f) e.dy = FLT_MAX;
	return e;
}
// LineSegment::operator(), rasterizer.cpp:66-69 : dx * (y - y0) / dy + x0
PS_D float edgeAt(const Edge& e, float y) { return fadd(fdiv(fmul(e.dx, fsub(y, e.y0)), e.dy), e.x0); }

// processTriangle's row range, rasterizer.cpp:95-97
PS_D void halfRange(int vpH, float yMin, float yMax, int& iy0, int& iy1)
{
	iy0 = yMin < 0 ? 0 : cvtt(yMin);
	const int lastRowIdx = vpH - 1;
	iy1 = yMax > (float)lastRowIdx ? lastRowIdx : cvtt(yMax);
}

PS_D uint32_t packRange(int a, int b)
{
	// row indices fit 16 bits (targets up to 32767 rows); an empty range is stored as 1..0
	if(a > b) { a = 1; b = 0; }
	return (uint32_t)(a & 0xffff) | ((uint32_t)(b & 0xffff) << 16);
}

// pushTriangle, rasterizer.cpp:142-234. ndc[i] = x,y of vertex i after the perspective divide. Returns 0 where the
// reference returns false (firstRow == lastRow), 2 where it returns true but walks no rows (entirely above/below the
// viewport: firstRow=0, lastRow=-1), 1 otherwise.
PS_D int setupTriangle(int vpW, int vpH, int halfW, int halfH, const float* ndcX, const float* ndcY, TriHeader& h, float* vx, float* vy)
{
	int firstRow = 0, lastRow = 0;
	// pushVertex, rasterizer.cpp:73-90
#pragma unroll
	for(int i = 0; i < 3; i++)
	{
		vy[i] = fadd(fmul((float)halfH, ndcY[i]), (float)halfH);
		vx[i] = fadd(fmul((float)halfW, ndcX[i]), (float)halfW);
		if(0 == i) firstRow = lastRow = cvtt(vy[0]);
		else if(vy[i] > (float)lastRow) lastRow = cvtt(vy[i]);
		else if(vy[i] < (float)firstRow) firstRow = cvtt(vy[i]);
	}
	if(firstRow >= vpH || lastRow < 0) { firstRow = 0; lastRow = -1; }
	if(firstRow < 0) firstRow = 0;
	if(lastRow >= vpH) lastRow = vpH - 1;
	if(firstRow == lastRow) return 0;
	if(firstRow > lastRow) return 2;

	h.vx0 = vx[0]; h.vy0 = vy[0]; h.vx1 = vx[1]; h.vy1 = vy[1]; h.vx2 = vx[2]; h.vy2 = vy[2];
	h.rows = packRange(firstRow, lastRow);

	int a = -1, b = 0, c = 0; // a,b: the two vertices with equal y; c: the apex
	if(vy[0] == vy[1]) { a = 0; b = 1; c = 2; }
	else if(vy[0] == vy[2]) { a = 0; b = 2; c = 1; }
	else if(vy[1] == vy[2]) { a = 1; b = 2; c = 0; }
	if(a >= 0)
	{
		// rasterizer.cpp:171-203 : the left edge starts at whichever of a,b has the smaller x
		const bool aLeft = sel3(vx, a) < sel3(vx, b);
		int l = aLeft ? a : b, r = aLeft ? b : a;
		h.plan = (uint32_t)l | ((uint32_t)c << 2) | ((uint32_t)r << 4) | ((uint32_t)c << 6);
		int y0, y1;
		halfRange(vpH, (float)firstRow, (float)lastRow, y0, y1);
		h.half0 = packRange(y0, y1);
		h.half1 = packRange(1, 0);
		return 1;
	}
	// rasterizer.cpp:205-231 : sort
	int top, bottom, third = 2;
	if(vy[0] > vy[1]) { top = 0; bottom = 1; } else { top = 1; bottom = 0; }
	if(sel3(vy, top) < vy[2]) { int s = top; top = third; third = s; }
	else if(sel3(vy, bottom) > vy[2]) { int s = bottom; bottom = third; third = s; }
	// processStandingTriangle, rasterizer.cpp:121-140
	const Edge eTT = makeEdge(vx, vy, top, third), eTB = makeEdge(vx, vy, top, bottom);
	uint32_t plan;
	const float y3 = sel3(vy, third);
	if(edgeAt(eTB, y3) > edgeAt(eTT, y3))
	{
		// upper: L = top-third, R = top-bottom ; lower: L = third-bottom, R = top-bottom
		plan = (uint32_t)top | ((uint32_t)third << 2) | ((uint32_t)top << 4) | ((uint32_t)bottom << 6)
		     | ((uint32_t)third << 8) | ((uint32_t)bottom << 10) | ((uint32_t)top << 12) | ((uint32_t)bottom << 14);
	}
	else
	{
		// upper: L = top-bottom, R = top-third ; lower: L = top-bottom, R = third-bottom
		plan = (uint32_t)top | ((uint32_t)bottom << 2) | ((uint32_t)top << 4) | ((uint32_t)third << 6)
		     | ((uint32_t)top << 8) | ((uint32_t)bottom << 10) | ((uint32_t)third << 12) | ((uint32_t)bottom << 14);
	}
	h.plan = plan | (1u << 16);
	int y0, y1;
	halfRange(vpH, y3, sel3(vy, top), y0, y1);
	h.half0 = packRange(y0, y1);
	halfRange(vpH, sel3(vy, bottom), y3, y0, y1);
	h.half1 = packRange(y0, y1);
	return 1;
}

// The row range pushTriangle derives from the three viewport y (rasterizer.cpp:73-90,149-168) and nothing else: what a
// sort-first rank needs to know to drop a triangle whose rows are all another rank's, before it pays for the rest of the
// vertex work. Same operations as setupTriangle above. Returns false where the triangle walks no row.
PS_D bool rowRangeOnly(int vpH, int halfH, const float* ndcY, int& firstRow, int& lastRow)
{
	float vy[3];
	firstRow = lastRow = 0;
#pragma unroll
	for(int i = 0; i < 3; i++)
	{
		vy[i] = fadd(fmul((float)halfH, ndcY[i]), (float)halfH);
		if(0 == i) firstRow = lastRow = cvtt(vy[0]);
		else if(vy[i] > (float)lastRow) lastRow = cvtt(vy[i]);
		else if(vy[i] < (float)firstRow) firstRow = cvtt(vy[i]);
	}
	if(firstRow >= vpH || lastRow < 0) { firstRow = 0; lastRow = -1; }
	if(firstRow < 0) firstRow = 0;
	if(lastRow >= vpH) lastRow = vpH - 1;
	return firstRow < lastRow;
}

struct RowSpan
{
	int left, right; // RESULT_ROW::left / right, unclamped (rasterizer.cpp:100,107)
	int edges;       // lv0 | lv1<<2 | rv0<<4 | rv1<<6 : the vertex pairs that produced the two ends
};

// One RESULT_ROW (rasterizer.cpp:98-117). The half written last (the lower one) wins on the shared row.
PS_D bool rowOf(const TriHeader& h, const float* vx, const float* vy, int iy, RowSpan& r)
{
	int sel;
	const int l0 = (int)(h.half1 & 0xffff), l1 = (int)(h.half1 >> 16);
	const int u0 = (int)(h.half0 & 0xffff), u1 = (int)(h.half0 >> 16);
	if(iy >= l0 && iy <= l1) sel = (int)((h.plan >> 8) & 0xff);
	else if(iy >= u0 && iy <= u1) sel = (int)(h.plan & 0xff);
	else return false;
	const Edge L = makeEdge(vx, vy, sel & 3, (sel >> 2) & 3);
	const Edge R = makeEdge(vx, vy, (sel >> 4) & 3, (sel >> 6) & 3);
	const float y = (float)iy;
	r.left = cvtt(fadd(edgeAt(L, y), 0.5f));
	r.right = cvtt(fadd(edgeAt(R, y), 0.5f));
	r.edges = sel;
	return true;
}

// lineSegmentlinearInterpolate, interp.cpp:151-160 : contribution of the two edge vertices at the rounded integer end
PS_D void edgeContrib(const float* vx, const float* vy, int v1, int v2, float x, float y, float* c)
{
	const float x1 = sel3(vx, v1), y1 = sel3(vy, v1), x2 = sel3(vx, v2), y2 = sel3(vy, v2);
	const float dx = fsub(x1, x2), dy = fsub(y1, y2);
	const bool alongX = fabsf(dx) > fabsf(dy);
	const float c1 = fdiv(alongX ? fsub(x, x2) : fsub(y, y2), alongX ? dx : dy);   // one divide, the operands of the chosen branch
	const float c2 = fsub(1.0f, c1);
	// c[v1] = c1, c[v2] = 1 - c1, c[the third] = 0 — written with selects to stay in registers
	c[0] = v1 == 0 ? c1 : (v2 == 0 ? c2 : 0.0f);
	c[1] = v1 == 1 ? c1 : (v2 == 1 ? c2 : 0.0f);
	c[2] = v1 == 2 ? c1 : (v2 == 2 ? c2 : 0.0f);
}


// ---- the same three routines reading the vertices by index from a TriHeader held in shared memory (vx0,vy0,vx1,vy1,vx2,vy2
// are its first six floats): a dynamically indexed LDS instead of a chain of selects ------------------------------------
PS_D Edge makeEdgeXY(const float* xy, int i0, int i1)
{
	Edge e;
	e.x0 = xy[2 * i0]; e.y0 = xy[2 * i0 + 1];
	e.dx = fsub(xy[2 * i1], e.x0);
	e.dy = fsub(xy[2 * i1 + 1], e.y0);
	if(fabsf(e.dy) < 0.000001f) e.dy = FLT_MAX;
	return e;
}
PS_D bool rowOfXY(const TriHeader& h, int iy, RowSpan& r)
{
	int sel;
	const int l0 = (int)(h.half1 & 0xffff), l1 = (int)(h.half1 >> 16);
	const int u0 = (int)(h.half0 & 0xffff), u1 = (int)(h.half0 >> 16);
	if(iy >= l0 && iy <= l1) sel = (int)((h.plan >> 8) & 0xff);
	else if(iy >= u0 && iy <= u1) sel = (int)(h.plan & 0xff);
	else return false;
	const float* xy = &h.vx0;
	const Edge L = makeEdgeXY(xy, sel & 3, (sel >> 2) & 3);
	const Edge R = makeEdgeXY(xy, (sel >> 4) & 3, (sel >> 6) & 3);
	const float y = (float)iy;
	r.left = cvtt(fadd(edgeAt(L, y), 0.5f));
	r.right = cvtt(fadd(edgeAt(R, y), 0.5f));
	r.edges = sel;
	return true;
}
PS_D void edgeContribXY(const float* xy, int v1, int v2, float x, float y, float* c)
{
	const float x1 = xy[2 * v1], y1 = xy[2 * v1 + 1], x2 = xy[2 * v2], y2 = xy[2 * v2 + 1];
	const float dx = fsub(x1, x2), dy = fsub(y1, y2);
	const bool alongX = fabsf(dx) > fabsf(dy);
	const float c1 = fdiv(alongX ? fsub(x, x2) : fsub(y, y2), alongX ? dx : dy);
	const float c2 = fsub(1.0f, c1);
	c[0] = v1 == 0 ? c1 : (v2 == 0 ? c2 : 0.0f);
	c[1] = v1 == 1 ? c1 : (v2 == 1 ? c2 : 0.0f);
	c[2] = v1 == 2 ? c1 : (v2 == 2 ? c2 : 0.0f);
}
