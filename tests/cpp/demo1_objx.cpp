// Demo 1 of the reference (src/test/puresoft.cpp:113-206, testobjs.cpp, scenobj.cpp) as a headless C++ caller: the sphere
// comes from an OBJX file through the native reader (SceneObject::findOrCreateVao, scenobj.cpp:88-163), the frame — shadow
// pass with the discarding cloud-shadow triple, cube-map skybox with depth off, earth, moon, alpha-blended cloud layer —
// goes through include/puresoft3d_b200.hpp with the demo's own class names (ps3d_demo1). Linked against the oracle library
// on CPU and against libps3d_b200.so on the GPU box; the test-suite compares what the two print.
//   usage: demo1_objx <sphere.objx> <width> <height> <shadow size>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ps3d_objx.h"
#include "puresoft3d_b200.hpp"
using namespace ps3d_demo1;

static unsigned long long fnv(const void* p, size_t n)
{
	const unsigned char* b = (const unsigned char*)p;
	unsigned long long h = 1469598103934665603ull;
	for(size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

struct Mat { float m[16]; };   // column-major like mcemath (m[col * 4 + row])
static Mat identity() { Mat r; memset(&r, 0, sizeof(r)); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
static Mat mul(const Mat& a, const Mat& b)
{
	Mat r;
	for(int c = 0; c < 4; c++) for(int rw = 0; rw < 4; rw++)
	{
		double s = 0;
		for(int k = 0; k < 4; k++) s += (double)a.m[k * 4 + rw] * (double)b.m[c * 4 + k];
		r.m[c * 4 + rw] = (float)s;
	}
	return r;
}
static Mat translation(float x, float y, float z) { Mat r = identity(); r.m[12] = x; r.m[13] = y; r.m[14] = z; return r; }
static Mat scaling(float s) { Mat r = identity(); r.m[0] = r.m[5] = r.m[10] = s; return r; }
static Mat rotationY(double a) { Mat r = identity(); r.m[0] = (float)cos(a); r.m[8] = (float)sin(a); r.m[2] = (float)-sin(a); r.m[10] = (float)cos(a); return r; }
static Mat perspective(float zn, float zf, float aspect, float fov)   // mcemaths_make_proj_perspective, matrxgl.cpp:9-22
{
	Mat r; memset(&r, 0, sizeof(r));
	const float h = (float)(1.0 / tan(fov / 2.0)), nd = zn - zf;
	r.m[0] = h / aspect; r.m[5] = h; r.m[10] = (zf + zn) / nd; r.m[11] = -1.0f; r.m[14] = 2.0f * (zn * zf) / nd;
	return r;
}
static void norm3(double* v) { const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); if(l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; } }
static void cross3(double* o, const double* a, const double* b) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
static Mat lookAt(const double* eye, const double* at)
{
	double f[3] = { eye[0] - at[0], eye[1] - at[1], eye[2] - at[2] }, up[3] = { 0, 1, 0 }, r[3], u[3];
	norm3(f); cross3(r, up, f); norm3(r); cross3(u, f, r);
	Mat m = identity();
	for(int k = 0; k < 3; k++) { m.m[k * 4 + 0] = (float)r[k]; m.m[k * 4 + 1] = (float)u[k]; m.m[k * 4 + 2] = (float)f[k]; }
	m.m[12] = (float)-(r[0] * eye[0] + r[1] * eye[1] + r[2] * eye[2]);
	m.m[13] = (float)-(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]);
	m.m[14] = (float)-(f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2]);
	return m;
}

static unsigned g_seed = 77;
static unsigned rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
// a smooth-ish seeded picture (neighbouring texels close, so that a nearest-texel flip stays inside the colour tolerance)
static int makeTexture(PuresoftPipeline& pipeline, int w, int h, bool cloud, int layers = 1)
{
	std::vector<std::vector<unsigned> > pix(layers, std::vector<unsigned>((size_t)w * h));
	for(int l = 0; l < layers; l++)
	{
		const double fx = 1 + rnd() % 4, fy = 1 + rnd() % 4, ph = (rnd() % 628) / 100.0;
		for(int y = 0; y < h; y++) for(int x = 0; x < w; x++)
		{
			const double v = 0.5 + 0.5 * sin(fx * 6.2831853 * x / w + ph) * cos(fy * 6.2831853 * y / h);
			const unsigned a = (unsigned)(40 + 200 * v), b = (unsigned)(220 - 180 * v), c = (unsigned)(128 + 100 * sin(ph + v * 3));
			pix[l][(size_t)y * w + x] = cloud ? (0xff000000u | (a << 16)) : (0xff000000u | (a << 16) | (b << 8) | (c & 0xff));   // cloud: alpha lives in red
		}
	}
	PURESOFTIMGBUFF32 img; img.width = w; img.height = h; img.elemLen = 4; img.scanline = w * 4; img.pixels = pix[0].data();
	const int idx = pipeline.createTexture(&img, layers - 1);
	for(int l = 1; l < layers; l++) pipeline.uploadTexture(idx, pix[l].data(), (PuresoftFBO::LAYER)l);
	return idx;
}

int main(int argc, char** argv)
{
	if(argc < 5) { fprintf(stderr, "usage: demo1_objx <sphere.objx> <width> <height> <shadow>\n"); return 2; }
	const int W = atoi(argv[2]), H = atoi(argv[3]), S = atoi(argv[4]);

	// ---- SceneObject::findOrCreateVao(objx), scenobj.cpp:88-163 ---------------------------------------------------------
	ps3d_objx* file = NULL;
	if(PS3D_OBJX_OK != ps3d_objx_open(argv[1], NULL, &file)) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
	ps3d_objx_mesh mi;
	if(PS3D_OBJX_OK != ps3d_objx_read_mesh_header(file, &mi) || !mi.has_normals || !mi.has_texcoords) return 2;
	const size_t nv = mi.num_vertices;
	std::vector<float> pos(nv * 4), nrm(nv * 4), tan(nv * 4), bin(nv * 4, 0.0f), uv(nv * 2);
	if(PS3D_OBJX_OK != ps3d_objx_read_mesh(file, pos.data(), nrm.data(), tan.data(), uv.data(), NULL)) return 2;
	ps3d_objx_close(file);
	for(size_t v = 0; v < nv; v++)
	{
		pos[v * 4 + 3] = 1.0f;                                                                               // :114-117
		double b[3], nn[3] = { nrm[v * 4], nrm[v * 4 + 1], nrm[v * 4 + 2] }, tt[3] = { tan[v * 4], tan[v * 4 + 1], tan[v * 4 + 2] };
		cross3(b, nn, tt); norm3(b);                                                                         // :120-129
		for(int k = 0; k < 3; k++) bin[v * 4 + k] = (float)b[k];
	}

	PuresoftPipeline pipeline(0, W, H);
	printf("backend %s\n", ps3d_backend_name());
	const int sphere = pipeline.createVAO();
	PuresoftVBO* v;
	v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(pos.data()); pipeline.attachVBO(sphere, 0, v);   // :131-158
	v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(tan.data()); pipeline.attachVBO(sphere, 1, v);
	v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(bin.data()); pipeline.attachVBO(sphere, 2, v);
	v = new PuresoftVBO(pipeline, 8, nv); v->updateContent(uv.data()); pipeline.attachVBO(sphere, 4, v);
	v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(nrm.data()); pipeline.attachVBO(sphere, 3, v);
	const float quadPos[24] = { -1, 1, 0, 1, -1, -1, 0, 1, 1, -1, 0, 1, 1, -1, 0, 1, 1, 1, 0, 1, -1, 1, 0, 1 };
	const int quad = pipeline.createVAO();
	v = new PuresoftVBO(pipeline, 16, 6); v->updateContent(quadPos); pipeline.attachVBO(quad, 0, v);

	const int diffuse = makeTexture(pipeline, 128, 64, false), bump = makeTexture(pipeline, 128, 64, false), spec = makeTexture(pipeline, 128, 64, false);
	const int night = makeTexture(pipeline, 128, 64, false), cloud = makeTexture(pipeline, 128, 64, true);
	const int moonD = makeTexture(pipeline, 64, 32, false), moonB = makeTexture(pipeline, 64, 32, false), sky = makeTexture(pipeline, 32, 32, false, 6);
	PURESOFTIMGBUFF32 sb; sb.width = S; sb.height = S; sb.elemLen = 4; sb.scanline = S * 4; sb.pixels = NULL;
	const int texShadow = pipeline.createTexture(&sb);                                                       // puresoft.cpp:141-147

	const int progShadow = pipeline.createProgramme(pipeline.addProcessor(new VertexProcesserDEF05), pipeline.addProcessor(new InterpolationProcessorDEF05), pipeline.addProcessor(new FragmentProcessorDEF05));
	const int progCloudShadow = pipeline.createProgramme(pipeline.addProcessor(new VP_CloudShadow), pipeline.addProcessor(new IP_CloudShadow), pipeline.addProcessor(new FP_CloudShadow));
	const int progSky = pipeline.createProgramme(pipeline.addProcessor(new VertexProcesserDEF04), pipeline.addProcessor(new InterpolationProcessorDEF04), pipeline.addProcessor(new FragmentProcessorDEF04));
	const int vPlanet = pipeline.addProcessor(new VP_Planet), iPlanet = pipeline.addProcessor(new IP_Planet);
	const int progEarth = pipeline.createProgramme(vPlanet, iPlanet, pipeline.addProcessor(new FP_Earth));
	const int progMoon = pipeline.createProgramme(vPlanet, iPlanet, pipeline.addProcessor(new FP_Satellite));
	const int progCloud = pipeline.createProgramme(pipeline.addProcessor(new VP_Cloud), pipeline.addProcessor(new IP_Cloud), pipeline.addProcessor(new FP_Cloud));

	const float PI = 3.14159265358979f;
	const double lightD[3] = { -2.0, 0.6, 2.4 }, cameraD[3] = { 0.0, 0.0, 2.2 }, origin[3] = { 0, 0, 0 };
	const float light[4] = { -2.0f, 0.6f, 2.4f, 0 }, camera[4] = { 0, 0, 2.2f, 0 };
	const Mat proj = perspective(0.1f, 10.0f, (float)W / H, 2 * PI * (30.0f / 360.0f)), view = translation(0, 0, -2.2f), pv = mul(proj, view);
	const Mat lproj = perspective(0.1f, 10.0f, 1.0f, 2 * PI * (30.0f / 360.0f)), lview = lookAt(lightD, origin), lpv = mul(lproj, lview);
	Mat bias = identity(); bias.m[0] = bias.m[5] = 0.5f; bias.m[12] = bias.m[13] = 0.5f;                    // puresoft.cpp:38-44
	const Mat lpvb = mul(bias, lpv);
	(void)cameraD;
	const Mat rotE = rotationY(0.7), modelE = rotE;
	const Mat rotC = rotationY(1.9), modelC = mul(rotC, scaling(1.1f));
	const Mat rotM = rotationY(2.6), modelM = mul(mul(rotM, translation(0.95f, 0, 0)), scaling(0.2f));
	auto place = [&](const Mat& model, const Mat& rot) { pipeline.setUniform(4, model.m, sizeof(Mat)); pipeline.setUniform(5, rot.m, sizeof(Mat)); };
	auto texUniform = [&](int slot, int tex) { pipeline.setUniform(slot, &tex, sizeof(int)); };

	pipeline.setUniform(7, light, 16); pipeline.setUniform(8, camera, 16);
	PURESOFTBGRA bk; bk.i32 = 0xff000000u;
	pipeline.clearColour(bk);
	// ---- shadow map, puresoft.cpp:162-190
	pipeline.setUniform(0, lproj.m, sizeof(Mat)); pipeline.setUniform(1, lview.m, sizeof(Mat)); pipeline.setUniform(3, lpv.m, sizeof(Mat));
	pipeline.setDepth(texShadow); pipeline.clearDepth(); pipeline.setViewport(S, S);
	pipeline.useProgramme(progShadow);
	place(modelE, rotE); pipeline.drawVAO(sphere);
	place(modelM, rotM); pipeline.drawVAO(sphere);
	place(modelC, rotC); texUniform(9, cloud);
	pipeline.useProgramme(progCloudShadow); pipeline.enable(BEHAVIOR_ALPHABLEND); pipeline.drawVAO(sphere); pipeline.disable(BEHAVIOR_ALPHABLEND);
	// ---- the scene, puresoft.cpp:192-206
	pipeline.setUniform(0, proj.m, sizeof(Mat)); pipeline.setUniform(1, view.m, sizeof(Mat)); pipeline.setUniform(3, pv.m, sizeof(Mat));
	pipeline.setDepth(); pipeline.clearDepth(); pipeline.setViewport(W, H);
	texUniform(15, texShadow); pipeline.setUniform(16, lpvb.m, sizeof(Mat)); texUniform(2, sky);
	pipeline.disable(BEHAVIOR_UPDATE_DEPTH | BEHAVIOR_TEST_DEPTH);
	pipeline.useProgramme(progSky); pipeline.drawVAO(quad, true);
	pipeline.enable(BEHAVIOR_UPDATE_DEPTH | BEHAVIOR_TEST_DEPTH);
	place(modelE, rotE); texUniform(9, diffuse); texUniform(10, bump); texUniform(11, spec); texUniform(12, night);
	pipeline.useProgramme(progEarth); pipeline.drawVAO(sphere);
	place(modelM, rotM); texUniform(9, moonD); texUniform(10, moonB);
	pipeline.useProgramme(progMoon); pipeline.drawVAO(sphere);
	place(modelC, rotC); texUniform(9, cloud);
	pipeline.useProgramme(progCloud); pipeline.enable(BEHAVIOR_ALPHABLEND); pipeline.drawVAO(sphere); pipeline.disable(BEHAVIOR_ALPHABLEND);
	pipeline.finish();

	std::vector<unsigned> colour((size_t)W * H);
	std::vector<float> depth((size_t)W * H), shadow((size_t)S * S);
	pipeline.readColour(colour.data(), W * 4);
	pipeline.readDepth(depth.data(), W * 4);
	pipeline.downloadTexture(texShadow, shadow.data());
	size_t covered = 0, shadowed = 0;
	for(size_t i = 0; i < depth.size(); i++) covered += depth[i] < 1.0f;
	for(size_t i = 0; i < shadow.size(); i++) shadowed += shadow[i] < 1.0f;
	const ps3d_stats st = pipeline.getStats();
	printf("stats %llu %llu %llu %llu %llu\n", (unsigned long long)st.triangles_submitted, (unsigned long long)st.spans,
	       (unsigned long long)st.fragments_tested, (unsigned long long)st.fragments_shaded, (unsigned long long)st.draws);
	printf("covered %zu shadowed %zu\n", covered, shadowed);
	printf("colour %016llx\n", fnv(colour.data(), colour.size() * 4));
	printf("depth %016llx\n", fnv(depth.data(), depth.size() * 4));
	printf("shadow %016llx\n", fnv(shadow.data(), shadow.size() * 4));
	return 0;
}
