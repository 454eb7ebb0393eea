// The product's coverage arithmetic — puresoft3d_b200/csrc/raster.cuh, compiled here for the HOST — against the reference's own
// PuresoftRasterizer (rasterizer.h read in place from the reference tree, the class itself from oracle/_ref/libps3d_ref.so) and
// PuresoftInterpolater::lineSegmentlinearInterpolate, triangle by triangle: return code, row range, every RESULT_ROW
// (left, right, the vertex pairs behind both ends), which rows pushTriangle leaves untouched, and the edge weights.
// Only runs where /root/reference exists (the test skips otherwise). Prints "<check> <checked> <mismatches>"; exit code = failing checks.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include "rasterizer.h"        // the reference's header, -I/root/reference/src/puresoft3d
#include "raster.cuh"          // the product's, host build

// static member of PuresoftInterpolater (interp.h:48); declared by its mangled name to avoid the reference's other headers
extern "C" void _ZN20PuresoftInterpolater28lineSegmentlinearInterpolateEPKfiiffPf(const float* verts, int v1, int v2, float x, float y, float* contributes);

static uint64_t g_state = 0x2545F4914F6CDD1Dull;
static uint32_t rnd32() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return (uint32_t)(g_state >> 16); }
static float uni(float a, float b) { return a + (b - a) * (float)(rnd32() & 0xffffff) / 16777216.0f; }

int main(int argc, char** argv)
{
	const long long triangles = argc > 1 ? atoll(argv[1]) : 300000;
	long long nCode = 0, badCode = 0, nRows = 0, badRows = 0, nUntouched = 0, badUntouched = 0, nW = 0, badW = 0;
	const int sizes[][2] = { { 640, 480 }, { 1920, 1080 }, { 238, 131 }, { 64, 4096 }, { 17, 9 } };
	for(int si = 0; si < 5; si++)
	{
		const int W = sizes[si][0], H = sizes[si][1];
		PuresoftRasterizer ref;
		const PuresoftRasterizer::RESULT* out = ref.initialize(W, H);
		PuresoftRasterizer::RESULT_ROW* rows = const_cast<PuresoftRasterizer::RESULT_ROW*>(out->m_rows);
		for(long long t = 0; t < triangles / 5; t++)
		{
			// triangles of every kind: tiny, mid, huge, off-screen, exact flat tops/bottoms on and off pixel rows, repeated vertices
			float v[3][4];
			const int kind = (int)(rnd32() % 8);
			const float cx = uni(-1.3f, 1.3f), cy = uni(-1.3f, 1.3f);
			const float ext = kind == 0 ? 0.004f : (kind == 1 ? 3.0f : (kind == 2 ? 40.0f : 0.25f));
			for(int i = 0; i < 3; i++) { v[i][0] = cx + uni(-ext, ext); v[i][1] = cy + uni(-ext, ext); v[i][2] = 0; v[i][3] = 1; }
			if(kind == 3) v[1][1] = v[0][1];                                                        // flat, fractional y
			if(kind == 4) { const float yy = (float)((int)(rnd32() % H) - H / 2) / (float)(H / 2); v[0][1] = yy; v[2][1] = yy; }   // flat, on a pixel row
			if(kind == 5) { v[2][0] = v[1][0]; v[2][1] = v[1][1]; }                                   // repeated vertex
			if(kind == 6) { v[0][0] += 4.0f; v[1][0] += 4.0f; v[2][0] += 4.0f; }                      // off-screen in x
			for(int y = 0; y < H; y++) rows[y].left = rows[y].right = 0x7e57;                        // sentinel: rows pushTriangle does not write
			const bool pushed = ref.pushTriangle(v[0], v[1], v[2]);

			TriHeader h;
			float vx[3], vy[3];
			const float ndcX[3] = { v[0][0], v[1][0], v[2][0] }, ndcY[3] = { v[0][1], v[1][1], v[2][1] };
			const int code = setupTriangle(W, H, W / 2, H / 2, ndcX, ndcY, h, vx, vy);
			nCode++;
			if(pushed != (code != 0)) { badCode++; continue; }
			if(code != 0 && (out->firstRow != (code == 1 ? (int)(h.rows & 0xffff) : 0) || out->lastRow != (code == 1 ? (int)(h.rows >> 16) : -1))) { badCode++; continue; }
			if(code != 1) continue;
			const float refVerts[6] = { out->vertices[0].x, out->vertices[0].y, out->vertices[1].x, out->vertices[1].y, out->vertices[2].x, out->vertices[2].y };
			for(int k = 0; k < 3; k++) if(vx[k] != refVerts[2 * k] || vy[k] != refVerts[2 * k + 1]) badCode++;
			for(int y = out->firstRow; y <= out->lastRow; y++)
			{
				RowSpan r;
				const bool mine = rowOf(h, vx, vy, y, r);
				const bool theirs = rows[y].left != 0x7e57 || rows[y].right != 0x7e57;
				nUntouched++;
				if(mine != theirs) { badUntouched++; continue; }      // (a written row that happens to equal the sentinel on both ends would show up here: it does not)
				if(!mine) continue;
				nRows++;
				const int x1 = r.left < 0 ? 0 : r.left, x2 = r.right >= W ? W - 1 : r.right;
				if(r.left != rows[y].left || r.right != rows[y].right || x1 != rows[y].leftClamped || x2 != rows[y].rightClamped ||
				   (r.edges & 3) != rows[y].leftVerts[0] || ((r.edges >> 2) & 3) != rows[y].leftVerts[1] ||
				   ((r.edges >> 4) & 3) != rows[y].rightVerts[0] || ((r.edges >> 6) & 3) != rows[y].rightVerts[1]) { badRows++; continue; }
				// interp.cpp:151-160 at the rounded ends of this row
				float mineL[3], mineR[3], refL[4], refR[4];
				edgeContrib(vx, vy, r.edges & 3, (r.edges >> 2) & 3, (float)r.left, (float)y, mineL);
				edgeContrib(vx, vy, (r.edges >> 4) & 3, (r.edges >> 6) & 3, (float)r.right, (float)y, mineR);
				_ZN20PuresoftInterpolater28lineSegmentlinearInterpolateEPKfiiffPf(refVerts, rows[y].leftVerts[0], rows[y].leftVerts[1], (float)rows[y].left, (float)y, refL);
				_ZN20PuresoftInterpolater28lineSegmentlinearInterpolateEPKfiiffPf(refVerts, rows[y].rightVerts[0], rows[y].rightVerts[1], (float)rows[y].right, (float)y, refR);
				nW++;
				if(memcmp(mineL, refL, 12) != 0 || memcmp(mineR, refR, 12) != 0)
				{
					// NaN weights (a repeated vertex makes 0/0) have no defined payload: equal when both are NaN lane by lane
					bool same = true;
					for(int k = 0; k < 3; k++) same = same && ((mineL[k] == refL[k] || (mineL[k] != mineL[k] && refL[k] != refL[k])) && (mineR[k] == refR[k] || (mineR[k] != mineR[k] && refR[k] != refR[k])));
					if(!same) badW++;
				}
			}
		}
	}
	printf("return_code_rows_vertices %lld %lld\n", nCode, badCode);
	printf("rows_written %lld %lld\n", nUntouched, badUntouched);
	printf("result_rows %lld %lld\n", nRows, badRows);
	printf("edge_weights %lld %lld\n", nW, badW);
	return (badCode != 0) + (badUntouched != 0) + (badRows != 0) + (badW != 0);
}
