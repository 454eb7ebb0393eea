// Demo 2 of the reference (src/test2/puresoft.cpp:100-248, loadscene.cpp:112-345) as a headless C++ caller: the scene comes
// from an OBJX file through the native reader (include/ps3d_objx.h), the frame goes through include/puresoft3d_b200.hpp —
// the host side a maintainer of the reference would keep, in the reference's own language. The test-suite links it once
// against the oracle library (CPU) and once against libps3d_b200.so (GPU) and compares what it prints.
//   usage: demo2_objx <file.objx> <width> <height> <shadow size>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>
#include "ps3d_objx.h"
#include "puresoft3d_b200.hpp"
using namespace ps3d_demo2;

static unsigned long long fnv(const void* p, size_t n)
{
	const unsigned char* b = (const unsigned char*)p;
	unsigned long long h = 1469598103934665603ull;
	for(size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

// column-major 4x4 like mcemath (m[col * 4 + row])
struct Mat { float m[16]; };
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
static Mat perspective(float zn, float zf, float aspect, float fov)   // mcemaths_make_proj_perspective, matrxgl.cpp:9-22
{
	Mat r; memset(&r, 0, sizeof(r));
	const float h = (float)(1.0 / tan(fov / 2.0)), nd = zn - zf;
	r.m[0] = h / aspect; r.m[5] = h; r.m[10] = (zf + zn) / nd; r.m[11] = -1.0f; r.m[14] = 2.0f * (zn * zf) / nd;
	return r;
}
static void norm3(double* v) { const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); if(l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; } }
static void cross3(double* o, const double* a, const double* b) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
static Mat lookAt(const double* eye, const double* at)               // mcemaths_make_view_traditional, matrxgl.cpp:36-148
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

struct Component   // SceneObject with a mesh, loadscene.cpp:263-343
{
	std::string name, programme, diffuseFile;
	std::vector<float> pos, tan, bin, nrm, uv;
	float centre[3], ambient[4], diffuse[4], specular[4], specularExponent;
	float world[3];
	int vao, prog, tex;
};

int main(int argc, char** argv)
{
	if(argc < 5) { fprintf(stderr, "usage: demo2_objx <file.objx> <width> <height> <shadow>\n"); return 2; }
	const int W = atoi(argv[2]), H = atoi(argv[3]), S = atoi(argv[4]);

	// ---- loadScene, loadscene.cpp:137-260 ------------------------------------------------------------------------------
	ps3d_objx_scene desc;
	ps3d_objx* file = NULL;
	if(PS3D_OBJX_OK != ps3d_objx_open(argv[1], &desc, &file)) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
	std::map<std::string, std::map<std::string, Component> > objects;   // std::map: the reference's iteration order
	for(int i = 0, n = ps3d_objx_mesh_count(file); i < n; i++)
	{
		ps3d_objx_mesh mh;
		if(PS3D_OBJX_OK != ps3d_objx_read_mesh_header(file, &mh)) return 2;
		const std::string full = mh.mesh_name;
		const size_t slash = full.find('/');
		if(std::string::npos == slash) { ps3d_objx_read_mesh(file, NULL, NULL, NULL, NULL, NULL); continue; }   // :150-155
		Component c;
		c.name = full; c.programme = mh.programme; c.diffuseFile = mh.diffuse_file;
		const size_t nv = mh.num_vertices;
		c.pos.resize(nv * 4); c.nrm.resize(nv * 4); c.tan.resize(nv * 4); c.uv.resize(nv * 2); c.bin.assign(nv * 4, 0.0f);
		if(PS3D_OBJX_OK != ps3d_objx_read_mesh(file, c.pos.data(), mh.has_normals ? c.nrm.data() : NULL, mh.has_texcoords ? c.tan.data() : NULL,
		                                      mh.has_texcoords ? c.uv.data() : NULL, NULL)) return 2;
		memcpy(c.ambient, mh.ambient_colour, 16); memcpy(c.diffuse, mh.diffuse_colour, 16); memcpy(c.specular, mh.specular_colour, 16);
		c.specularExponent = mh.specular_exponent;
		float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
		for(size_t v = 0; v < nv; v++)
		{
			for(int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], c.pos[v * 4 + k]); hi[k] = fmaxf(hi[k], c.pos[v * 4 + k]); }
			double b[3], nn[3] = { c.nrm[v * 4], c.nrm[v * 4 + 1], c.nrm[v * 4 + 2] }, tt[3] = { c.tan[v * 4], c.tan[v * 4 + 1], c.tan[v * 4 + 2] };
			cross3(b, nn, tt); norm3(b);                                                                          // :181-189
			for(int k = 0; k < 3; k++) c.bin[v * 4 + k] = (float)b[k];
		}
		for(int k = 0; k < 3; k++) c.centre[k] = (float)(((double)lo[k] + (double)hi[k]) / 2.0);              // :215-228
		for(size_t v = 0; v < nv; v++) { for(int k = 0; k < 3; k++) c.pos[v * 4 + k] -= c.centre[k]; c.pos[v * 4 + 3] = 1.0f; }   // :229-233
		objects[full.substr(0, slash)][full.substr(slash + 1)] = c;
	}
	ps3d_objx_close(file);
	std::vector<Component*> comps;
	double lightFrom[3] = { 0, 0, 0 }, lightTo[3] = { 0, 0, 0 };
	for(auto& o : objects)
		for(auto& m : o.second)
		{
			Component& c = m.second;
			for(int k = 0; k < 3; k++) c.world[k] = c.centre[k];      // object translation + (mesh centre - object translation), :238-260
			if(c.name == "light1/from") { for(int k = 0; k < 3; k++) lightFrom[k] = c.world[k]; continue; }   // marker meshes, :420-440
			if(c.name == "light1/to") { for(int k = 0; k < 3; k++) lightTo[k] = c.world[k]; continue; }
			comps.push_back(&c);
		}

	// ---- pipeline objects, puresoft.cpp:100-160 ------------------------------------------------------------------------
	PuresoftPipeline pipeline(0, W, H);
	printf("backend %s\n", ps3d_backend_name());
	std::map<std::string, int> progs, texs;
	unsigned seed = 2024;
	for(Component* c : comps)
	{
		c->vao = pipeline.createVAO();
		const size_t nv = c->pos.size() / 4;
		PuresoftVBO* v;
		v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(c->pos.data()); pipeline.attachVBO(c->vao, 0, v);
		v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(c->tan.data()); pipeline.attachVBO(c->vao, 1, v);
		v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(c->bin.data()); pipeline.attachVBO(c->vao, 2, v);
		v = new PuresoftVBO(pipeline, 8, nv); v->updateContent(c->uv.data()); pipeline.attachVBO(c->vao, 4, v);
		v = new PuresoftVBO(pipeline, 16, nv); v->updateContent(c->nrm.data()); pipeline.attachVBO(c->vao, 3, v);
		if(!progs.count(c->programme))   // findOrCreateProgramme: "VP_x:IP_y:FP_z"
		{
			int p = -1;
			if(c->programme == "VP_SingleColour:IP_SingleColour:FP_SingleColour")
				p = pipeline.createProgramme(pipeline.addProcessor(new VP_SingleColour), pipeline.addProcessor(new IP_SingleColour), pipeline.addProcessor(new FP_SingleColour));
			else if(c->programme == "VP_DiffuseOnly:IP_DiffuseOnly:FP_DiffuseOnly")
				p = pipeline.createProgramme(pipeline.addProcessor(new VP_DiffuseOnly), pipeline.addProcessor(new IP_DiffuseOnly), pipeline.addProcessor(new FP_DiffuseOnly));
			else if(c->programme == "VP_PositionOnly:IP_Null:FP_SingleColourNoLighting")
				p = pipeline.createProgramme(pipeline.addProcessor(new VP_PositionOnly), pipeline.addProcessor(new IP_Null), pipeline.addProcessor(new FP_SingleColourNoLighting));
			else { fprintf(stderr, "unknown programme %s\n", c->programme.c_str()); return 2; }
			progs[c->programme] = p;
		}
		c->prog = progs[c->programme];
		c->tex = -2;
		if(c->programme == "VP_DiffuseOnly:IP_DiffuseOnly:FP_DiffuseOnly")
		{
			if(!texs.count(c->diffuseFile))   // no picture decoder here: a seeded stand-in per file name
			{
				std::vector<unsigned> pix(64 * 64);
				for(size_t i = 0; i < pix.size(); i++) { seed = seed * 1664525u + 1013904223u; pix[i] = 0xff000000u | ((seed >> 8) & 0x00ffffffu); }
				PURESOFTIMGBUFF32 img; img.width = 64; img.height = 64; img.elemLen = 4; img.scanline = 64 * 4; img.pixels = pix.data();
				texs[c->diffuseFile] = pipeline.createTexture(&img);
			}
			c->tex = texs[c->diffuseFile];
		}
	}
	PURESOFTIMGBUFF32 shadowBuffer; shadowBuffer.width = S; shadowBuffer.height = S; shadowBuffer.elemLen = 4; shadowBuffer.scanline = S * 4; shadowBuffer.pixels = NULL;
	const int texShadow = pipeline.createTexture(&shadowBuffer);
	const int progShadow = pipeline.createProgramme(pipeline.addProcessor(new VP_Shadow), pipeline.addProcessor(new IP_Null), pipeline.addProcessor(new FP_Null));

	const float PI = 3.14159265358979f;
	const Mat proj = perspective(0.1f, 5.0f, (float)W / H, 2 * PI * (45.0f / 360.0f));
	const double y = desc.camera_ypr[0], p = desc.camera_ypr[1];
	const double eye[3] = { desc.camera_pos[0], desc.camera_pos[1], desc.camera_pos[2] };
	const double at[3] = { eye[0] + cos(p) * sin(y), eye[1] - sin(p), eye[2] - cos(p) * cos(y) };   // matrxgl.cpp:163-166
	const Mat view = lookAt(eye, at), projView = mul(proj, view);
	const Mat light1Proj = perspective(0.1f, 5.0f, 1.0f, 2 * PI * (90.0f / 360.0f)), light1View = lookAt(lightFrom, lightTo);
	const Mat light1pv = mul(light1Proj, light1View);
	Mat bias = identity(); bias.m[0] = bias.m[5] = 0.5f; bias.m[12] = bias.m[13] = 0.5f;          // src/test/puresoft.cpp:38-44
	const Mat light1pvb = mul(bias, light1pv);
	double rd[3] = { lightFrom[0] - lightTo[0], lightFrom[1] - lightTo[1], lightFrom[2] - lightTo[2] };
	norm3(rd);
	const float light1from[4] = { (float)lightFrom[0], (float)lightFrom[1], (float)lightFrom[2], 0 }, light1RDir[4] = { (float)rd[0], (float)rd[1], (float)rd[2], 0 };
	const float cameraPos[4] = { (float)eye[0], (float)eye[1], (float)eye[2], 0 };
	const Mat rotation = identity();

	auto drawComponent = [&](Component* c, const Mat& pv, bool privateProgramme)   // SceneObject::draw, loadscene.cpp:112-135
	{
		const Mat model = translation(c->world[0], c->world[1], c->world[2]), pvm = mul(pv, model);
		pipeline.setUniform(0, model.m, sizeof(Mat)); pipeline.setUniform(1, rotation.m, sizeof(Mat)); pipeline.setUniform(5, pvm.m, sizeof(Mat));
		pipeline.setUniform(30, c->ambient, 16); pipeline.setUniform(31, c->diffuse, 16); pipeline.setUniform(32, c->specular, 16);
		pipeline.setUniform(33, &c->specularExponent, sizeof(float));
		pipeline.setUniform(40, &c->tex, sizeof(int));
		if(privateProgramme) pipeline.useProgramme(c->prog);
		pipeline.drawVAO(c->vao);
	};

	int draws = 0;
	for(int frame = 0; frame < 2; frame++)   // puresoft.cpp:170-248, twice: the second frame starts from the first one's targets
	{
		pipeline.setUniform(2, light1View.m, sizeof(Mat)); pipeline.setUniform(3, light1Proj.m, sizeof(Mat)); pipeline.setUniform(4, light1pv.m, sizeof(Mat));
		pipeline.setDepth(texShadow);
		pipeline.clearDepth();
		pipeline.setViewport(S, S);
		pipeline.useProgramme(progShadow);
		for(Component* c : comps) if(std::string::npos == c->name.find("@noshadow")) { drawComponent(c, light1pv, false); draws++; }
		pipeline.setUniform(2, view.m, sizeof(Mat)); pipeline.setUniform(3, proj.m, sizeof(Mat)); pipeline.setUniform(4, projView.m, sizeof(Mat));
		pipeline.setUniform(6, light1pvb.m, sizeof(Mat));
		pipeline.setUniform(20, light1from, 16); pipeline.setUniform(21, light1RDir, 16); pipeline.setUniform(22, cameraPos, 16);
		pipeline.setUniform(23, &texShadow, sizeof(int));
		pipeline.setDepth();
		pipeline.clearDepth();
		pipeline.clearColour();
		pipeline.setViewport(W, H);
		for(Component* c : comps) { drawComponent(c, projView, true); draws++; }
		pipeline.finish();
		pipeline.swapBuffers();
	}
	pipeline.swapBuffers();   // back to the frame just drawn

	std::vector<unsigned> colour((size_t)W * H);
	std::vector<float> depth((size_t)W * H), shadow((size_t)S * S);
	pipeline.readColour(colour.data(), W * 4);
	pipeline.readDepth(depth.data(), W * 4);
	pipeline.downloadTexture(texShadow, shadow.data());
	size_t covered = 0;
	for(size_t i = 0; i < depth.size(); i++) covered += depth[i] < 1.0f;
	const ps3d_stats st = pipeline.getStats();
	printf("components %d draws %d\n", (int)comps.size(), draws);
	printf("stats %llu %llu %llu %llu\n", (unsigned long long)st.triangles_submitted, (unsigned long long)st.spans,
	       (unsigned long long)st.fragments_tested, (unsigned long long)st.fragments_shaded);
	printf("covered %zu\n", covered);
	printf("colour %016llx\n", fnv(colour.data(), colour.size() * 4));
	printf("depth %016llx\n", fnv(depth.data(), depth.size() * 4));
	printf("shadow %016llx\n", fnv(shadow.data(), shadow.size() * 4));
	return 0;
}
