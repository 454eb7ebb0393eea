// A caller written the way the reference's demos are (src/test/puresoft.cpp:113-206, src/test/scenobj.cpp:135-158),
// against include/puresoft3d_b200.hpp. The test-suite links it once against the oracle library (CPU: host-layer logic,
// exception mapping, ownership) and once against libps3d_b200.so (GPU: same frame, compared word for word).
// Prints:  backend <name> / errors ok / stats <submitted> <rasterised> <spans> <tested> <shaded> / colour <fnv> / depth <fnv>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "puresoft3d_b200.hpp"

static unsigned long long fnv(const void* p, size_t n)
{
	const unsigned char* b = (const unsigned char*)p;
	unsigned long long h = 1469598103934665603ull;
	for(size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
	return h;
}

template<class EXC, class FN> static bool throws(FN fn)
{
	try { fn(); } catch(const EXC&) { return true; } catch(...) { return false; }
	return false;
}

int main()
{
	const int W = 200, H = 120;
	PuresoftPipeline pipeline(0, W, H);
	printf("backend %s\n", ps3d_backend_name());

	// error behaviour of the reference: prog.cpp:45-58,123-126; pipeline.cpp:140-148,232-235; tex.cpp:48-51
	bool ok = true;
	ok &= throws<std::out_of_range>([&] { pipeline.useProgramme(7); });
	ok &= throws<std::out_of_range>([&] { pipeline.createProgramme(0, 1, 2); });
	ok &= throws<std::out_of_range>([&] { pipeline.destroyTexture(99); });
	ok &= throws<std::out_of_range>([&] { pipeline.setDepth(5); });
	PURESOFTIMGBUFF32 odd; odd.width = 8; odd.height = 8; odd.elemLen = 1; odd.scanline = 8; odd.pixels = NULL;
	const int texOdd = pipeline.createTexture(&odd);
	ok &= throws<std::invalid_argument>([&] { pipeline.setDepth(texOdd); });
	pipeline.destroyTexture(texOdd);
	pipeline.drawVAO(3); // silently ignored: no programme in use, bad vao (drawvao.cpp:12-15)
	printf("errors %s\n", ok ? "ok" : "WRONG");

	// DEF02: slots 0 position, 1 normal, 2 colour (colr1light1.cpp:21-39)
	std::vector<float> pos, nrm, col;
	unsigned s = 12345;
	auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
	for(int t = 0; t < 60; t++)
	{
		const float cx = rnd() * 2.2f - 1.1f, cy = rnd() * 2.2f - 1.1f, r = 0.05f + rnd() * 0.4f, a0 = rnd() * 6.2831853f;
		for(int k = 0; k < 3; k++)
		{
			const float a = a0 + 2.0943951f * k;
			pos.push_back(cx + r * cosf(a)); pos.push_back(cy + r * sinf(a)); pos.push_back(rnd() * 1.8f - 0.9f); pos.push_back(1.0f);
			nrm.push_back(0); nrm.push_back(0); nrm.push_back(1.0f); nrm.push_back(0);
			col.push_back(40 + 200 * rnd()); col.push_back(40 + 200 * rnd()); col.push_back(40 + 200 * rnd()); col.push_back(255.0f);
		}
	}
	const size_t nverts = pos.size() / 4;
	const int vao = pipeline.createVAO();
	PuresoftVBO* vp = new PuresoftVBO(pipeline, 16, nverts); vp->updateContent(pos.data());
	PuresoftVBO* vn = new PuresoftVBO(pipeline, 16, nverts); vn->updateContent(nrm.data());
	PuresoftVBO* vc = new PuresoftVBO(pipeline, 16, nverts); vc->updateContent(col.data());
	pipeline.attachVBO(vao, 0, vp);
	pipeline.attachVBO(vao, 1, vn);
	PuresoftVBO* displaced = pipeline.attachVBO(vao, 2, vn); // wrong on purpose ...
	if(displaced) ok = false;
	displaced = pipeline.attachVBO(vao, 2, vc);               // ... the displaced VBO comes back to the caller
	if(displaced != vn) ok = false;
	if(pipeline.getVBO(vao, 0) != vp) ok = false;

	const int prog = pipeline.createProgramme(pipeline.addProcessor(new VertexProcesserDEF02),
	                                          pipeline.addProcessor(new InterpolationProcessorDEF02),
	                                          pipeline.addProcessor(new FragmentProcessorDEF02));
	const float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	const float light[4] = { 0.3f, 0.4f, 2.0f, 0 }, camera[4] = { 0, 0, 3.0f, 0 };
	pipeline.setUniform(3, ident, sizeof(ident));
	pipeline.setUniform(4, ident, sizeof(ident));
	pipeline.setUniform(5, ident, sizeof(ident));
	pipeline.setUniform(7, light, sizeof(light));
	pipeline.setUniform(8, camera, sizeof(camera));
	pipeline.setViewport(W, H);
	pipeline.clearDepth();
	PURESOFTBGRA bk; bk.i32 = 0xff101010u;
	pipeline.clearColour(bk);
	pipeline.disable(BEHAVIOR_FACE_CULLING);
	pipeline.useProgramme(prog);
	pipeline.drawVAO(vao);
	pipeline.finish();
	pipeline.swapBuffers();
	pipeline.swapBuffers();

	std::vector<unsigned> colour((size_t)W * H);
	std::vector<float> depth((size_t)W * H);
	pipeline.readColour(colour.data(), W * 4);
	pipeline.readDepth(depth.data(), W * 4);
	const ps3d_stats st = pipeline.getStats();
	printf("ownership %s\n", ok ? "ok" : "WRONG");
	printf("stats %llu %llu %llu %llu %llu\n", (unsigned long long)st.triangles_submitted, (unsigned long long)st.triangles_rasterised,
	       (unsigned long long)st.spans, (unsigned long long)st.fragments_tested, (unsigned long long)st.fragments_shaded);
	printf("colour %016llx\n", fnv(colour.data(), colour.size() * 4));
	printf("depth %016llx\n", fnv(depth.data(), depth.size() * 4));
	// post-processing hook, as src/test2/puresoft.cpp:243 would call it
	PP_DepthofField depthofField;
	pipeline.postProcess(&depthofField);
	pipeline.readColour(colour.data(), W * 4);
	printf("post %016llx\n", fnv(colour.data(), colour.size() * 4));
	pipeline.destroyVAO(vao); // takes vp, vn(attached? no: displaced), vc with it (pipeline.cpp:194-201)
	delete vn;                // handed back by attachVBO, so it is ours
	return ok ? 0 : 1;
}
