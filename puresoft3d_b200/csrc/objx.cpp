// objx.cpp — OBJX scene files: reader and writer behind include/ps3d_objx.h (host-side I/O, no device work).
//
// Written from the format, not from the reference's code: the on-disk layout is documented in include/ps3d_objx.h
// (what /root/reference/src/objcvt/objxio.cpp:6-45 declares with #pragma pack(1)); the semantics kept are
//   * version check and "mesh header on disk may be LONGER than ours" (forward compatibility, objxio.cpp:196-203, :264)
//   * colours stored b,g,r,a, returned r,g,b,a clamped to [0,1] (:272-283)
//   * light type from the length of the stored direction (:216-222)
//   * a mesh's payload is skipped by its next_offset, whatever arrays the caller asked for (:357-359)
//   * tangents generated when the file has texcoords but no tangents (:318-326, :418-470)
// Fields are (de)serialised one by one at their byte offsets, so the code does not depend on struct packing.
#include "ps3d_objx.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

enum : size_t
{
	FILE_HEADER_BYTES = 16 + 16 + 16 + 4 * 32,                       // 176
	MH_NAME = 0, MH_NUM_VERTICES = 260, MH_NUM_INDICES = 264, MH_HAS_TEXCOORDS = 268, MH_HAS_NORMALS = 269, MH_HAS_TANGENTS = 270,
	MH_AMBIENT = 271, MH_DIFFUSE = 287, MH_SPECULAR = 303, MH_SPECULAR_EXP = 319,
	MH_DIFFUSE_FILE = 323, MH_BUMP_FILE = 583, MH_SPC_FILE = 843, MH_SPE_FILE = 1103, MH_PROGRAMME = 1363, MH_NEXT_OFFSET = 1623,
	MESH_HEADER_BYTES = 1627,
};

uint32_t getU32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
void putU32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
float getF32(const uint8_t* p) { const uint32_t u = getU32(p); float f; memcpy(&f, &u, 4); return f; }
void putF32(uint8_t* p, float f) { uint32_t u; memcpy(&u, &f, 4); putU32(p, u); }
float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
void getStr(const uint8_t* p, char* dst) { memcpy(dst, p, PS3D_OBJX_NAMELEN); dst[PS3D_OBJX_NAMELEN - 1] = 0; }
void putStr(uint8_t* p, const char* src) { const size_t n = strnlen(src, PS3D_OBJX_NAMELEN - 1); memcpy(p, src, n); }

// stored b,g,r,a -> r,g,b,a, clamped
void getColour(const uint8_t* p, float* rgba)
{
	rgba[2] = clamp01(getF32(p)); rgba[1] = clamp01(getF32(p + 4)); rgba[0] = clamp01(getF32(p + 8)); rgba[3] = clamp01(getF32(p + 12));
}
void putColour(uint8_t* p, const float* rgba) { putF32(p, rgba[2]); putF32(p + 4, rgba[1]); putF32(p + 8, rgba[0]); putF32(p + 12, rgba[3]); }

size_t payloadBytes(uint32_t nv, uint32_t ni, bool tex, bool nrm, bool tan)
{
	size_t b = (size_t)nv * 16;
	if(tex) b += (size_t)nv * 8;
	if(nrm) b += (size_t)nv * 16;
	if(tan) b += (size_t)nv * 16;
	b += (size_t)ni * 4;
	return b;
}

} // namespace

struct ps3d_objx
{
	FILE* file = nullptr;
	bool writing = false;
	uint32_t numMeshes = 0, meshHeaderBytes = MESH_HEADER_BYTES;
	uint8_t fileHeader[FILE_HEADER_BYTES];
	// the mesh whose header was read last (reader) and where its payload starts
	bool headerPending = false;
	uint32_t nv = 0, ni = 0, nextOffset = 0;
	bool hasTex = false, hasNrm = false, hasTan = false;
	long payloadStart = 0;
	long fileBytes = 0;     // reader: a header that promises more bytes than the file holds is refused before anything is allocated for it
};

namespace {

// One triangle's tangent from its positions and texcoords: the object-space direction of increasing u,
//   T = ((c - a) * (tb.v - ta.v) - (b - a) * (tc.v - ta.v)) / ((tc.u - ta.u) * (tb.v - ta.v) - (tc.v - ta.v) * (tb.u - ta.u))
// (the standard solution of  b - a = T * du1 + B * dv1,  c - a = T * du2 + B * dv2  for T, written with the reference's
// operand naming, objxio.cpp:444-470; that routine subtracts into the wrong temporary and so uses an uninitialised
// vector — a latent bug its shipped fixtures never reach, because both carry tangents. This is the intended formula.)
void triangleTangent(float* t, const float* a, const float* b, const float* c, const float* ta, const float* tb, const float* tc)
{
	const float c1x = tc[0] - ta[0], c1y = tc[1] - ta[1];
	const float c2x = tb[0] - ta[0], c2y = tb[1] - ta[1];
	const float det = c1x * c2y - c1y * c2x;
	const float f = 1.0f / det;
	for(int k = 0; k < 3; k++)
	{
		const float v1 = (c[k] - a[k]) * c2y;   // each edge scaled by the other edge's dv
		const float v2 = (b[k] - a[k]) * c1y;
		t[k] = (v1 - v2) * f;
	}
	t[3] = 0.0f;
}
void normalise3(float* v)
{
	const float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	if(l > 0.0f) { v[0] /= l; v[1] /= l; v[2] /= l; }
}
void generateTangents(float* tangents, const float* vertices, const float* texcoords, const int32_t* indices, uint32_t nv, uint32_t ni)
{
	float t[4];
	if(ni > 0 && indices)
	{
		// indexed: every vertex accumulates the tangents of the triangles that share it, then normalises
		memset(tangents, 0, (size_t)nv * 16);
		for(uint32_t i = 0; i + 2 < ni; i += 3)
		{
			const int32_t ia = indices[i], ib = indices[i + 1], ic = indices[i + 2];
			if(ia < 0 || ib < 0 || ic < 0 || (uint32_t)ia >= nv || (uint32_t)ib >= nv || (uint32_t)ic >= nv) continue;
			triangleTangent(t, vertices + 4 * ia, vertices + 4 * ib, vertices + 4 * ic, texcoords + 2 * ia, texcoords + 2 * ib, texcoords + 2 * ic);
			for(int k = 0; k < 3; k++) { tangents[4 * ia + k] += t[k]; tangents[4 * ib + k] += t[k]; tangents[4 * ic + k] += t[k]; }
		}
		for(uint32_t i = 0; i < nv; i++) normalise3(tangents + 4 * i);
	}
	else
	{
		// un-indexed: three consecutive vertices are one triangle and share its tangent
		for(uint32_t i = 0; i + 2 < nv; i += 3)
		{
			triangleTangent(t, vertices + 4 * i, vertices + 4 * (i + 1), vertices + 4 * (i + 2), texcoords + 2 * i, texcoords + 2 * (i + 1), texcoords + 2 * (i + 2));
			normalise3(t);
			for(int v = 0; v < 3; v++) memcpy(tangents + 4 * (i + v), t, 16);
		}
	}
}

bool readOrSkip(FILE* f, void* dst, size_t bytes, bool present)
{
	if(!present || 0 == bytes) return true;
	if(dst) return 1 == fread(dst, bytes, 1, f);
	return 0 == fseek(f, (long)bytes, SEEK_CUR);
}

} // namespace

extern "C" {

int ps3d_objx_open(const char* filename, ps3d_objx_scene* scene, ps3d_objx** out)
{
	if(!filename || !out) return PS3D_OBJX_ERR_ARGUMENT;
	*out = nullptr;
	FILE* f = fopen(filename, "rb");
	if(!f) return PS3D_OBJX_ERR_IO;
	ps3d_objx* h = new ps3d_objx();
	h->file = f;
	if(0 == fseek(f, 0, SEEK_END)) { h->fileBytes = ftell(f); rewind(f); }
	if(1 != fread(h->fileHeader, FILE_HEADER_BYTES, 1, f)) { fclose(f); delete h; return PS3D_OBJX_ERR_FORMAT; }
	const uint32_t version = getU32(h->fileHeader);
	h->numMeshes = getU32(h->fileHeader + 4);
	h->meshHeaderBytes = getU32(h->fileHeader + 8);
	if(version != PS3D_OBJX_VERSION || h->meshHeaderBytes < MESH_HEADER_BYTES) { fclose(f); delete h; return PS3D_OBJX_ERR_FORMAT; }
	if(scene)
	{
		const uint8_t* p = h->fileHeader + 16;
		for(int i = 0; i < 4; i++) scene->camera_pos[i] = getF32(p + 4 * i);
		for(int i = 0; i < 4; i++) scene->camera_ypr[i] = getF32(p + 16 + 4 * i);
		for(int l = 0; l < PS3D_OBJX_MAXLIGHTS; l++)
		{
			const uint8_t* q = p + 32 + 32 * l;
			for(int i = 0; i < 4; i++) { scene->light_pos[l][i] = getF32(q + 4 * i); scene->light_dir[l][i] = getF32(q + 16 + 4 * i); }
			const float* d = scene->light_dir[l];
			scene->light_types[l] = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < 0.0000001f ? PS3D_OBJX_LT_OMNI : PS3D_OBJX_LT_DIR;
		}
	}
	*out = h;
	return PS3D_OBJX_OK;
}

int ps3d_objx_mesh_count(ps3d_objx* h) { return h ? (int)h->numMeshes : PS3D_OBJX_ERR_ARGUMENT; }

int ps3d_objx_read_mesh_header(ps3d_objx* h, ps3d_objx_mesh* mesh)
{
	if(!h || !mesh) return PS3D_OBJX_ERR_ARGUMENT;
	if(h->writing || h->headerPending) return PS3D_OBJX_ERR_STATE;
	uint8_t raw[MESH_HEADER_BYTES];
	if(1 != fread(raw, MESH_HEADER_BYTES, 1, h->file)) return PS3D_OBJX_ERR_IO;
	if(h->meshHeaderBytes > MESH_HEADER_BYTES && 0 != fseek(h->file, (long)(h->meshHeaderBytes - MESH_HEADER_BYTES), SEEK_CUR)) return PS3D_OBJX_ERR_IO;
	memset(mesh, 0, sizeof(*mesh));
	getStr(raw + MH_NAME, mesh->mesh_name);
	mesh->num_vertices = h->nv = getU32(raw + MH_NUM_VERTICES);
	mesh->num_indices = h->ni = getU32(raw + MH_NUM_INDICES);
	mesh->has_texcoords = h->hasTex = 0 != raw[MH_HAS_TEXCOORDS];
	mesh->has_normals = h->hasNrm = 0 != raw[MH_HAS_NORMALS];
	mesh->has_tangents = h->hasTan = 0 != raw[MH_HAS_TANGENTS];
	getColour(raw + MH_AMBIENT, mesh->ambient_colour);
	getColour(raw + MH_DIFFUSE, mesh->diffuse_colour);
	getColour(raw + MH_SPECULAR, mesh->specular_colour);
	mesh->specular_exponent = getF32(raw + MH_SPECULAR_EXP);
	getStr(raw + MH_DIFFUSE_FILE, mesh->diffuse_file);
	getStr(raw + MH_BUMP_FILE, mesh->bump_file);
	getStr(raw + MH_SPC_FILE, mesh->spc_file);
	getStr(raw + MH_SPE_FILE, mesh->spe_file);
	getStr(raw + MH_PROGRAMME, mesh->programme);
	h->nextOffset = getU32(raw + MH_NEXT_OFFSET);
	h->payloadStart = ftell(h->file);
	h->headerPending = true;
	if(payloadBytes(h->nv, h->ni, h->hasTex, h->hasNrm, h->hasTan) > h->nextOffset) return PS3D_OBJX_ERR_FORMAT;
	if(h->fileBytes > 0 && (unsigned long long)h->payloadStart + h->nextOffset > (unsigned long long)h->fileBytes) return PS3D_OBJX_ERR_FORMAT;
	return PS3D_OBJX_OK;
}

int ps3d_objx_read_mesh(ps3d_objx* h, float* vertices, float* normals, float* tangents, float* texcoords, int32_t* indices)
{
	if(!h) return PS3D_OBJX_ERR_ARGUMENT;
	if(h->writing || !h->headerPending) return PS3D_OBJX_ERR_STATE;
	h->headerPending = false;
	FILE* f = h->file;
	const size_t v16 = (size_t)h->nv * 16;
	bool ok = readOrSkip(f, vertices, v16, true);
	ok = ok && readOrSkip(f, normals, v16, h->hasNrm);
	ok = ok && readOrSkip(f, tangents, v16, h->hasTan);
	ok = ok && readOrSkip(f, texcoords, (size_t)h->nv * 8, h->hasTex);
	const bool generate = tangents && !h->hasTan && h->hasTex && texcoords && vertices;
	std::vector<int32_t> idxTmp;
	int32_t* idx = indices;
	if(generate && !idx && h->ni) { idxTmp.resize(h->ni); idx = idxTmp.data(); }
	ok = ok && readOrSkip(f, idx, (size_t)h->ni * 4, h->ni > 0);
	if(!ok) return PS3D_OBJX_ERR_IO;
	if(generate) generateTangents(tangents, vertices, texcoords, idx, h->nv, h->ni);
	// whatever was read: the next mesh starts next_offset bytes behind this mesh's header
	if(0 != fseek(f, h->payloadStart + (long)h->nextOffset, SEEK_SET)) return PS3D_OBJX_ERR_IO;
	return PS3D_OBJX_OK;
}

int ps3d_objx_create(const char* filename, const ps3d_objx_scene* scene, ps3d_objx** out)
{
	if(!filename || !out) return PS3D_OBJX_ERR_ARGUMENT;
	*out = nullptr;
	FILE* f = fopen(filename, "w+b");
	if(!f) return PS3D_OBJX_ERR_IO;
	ps3d_objx* h = new ps3d_objx();
	h->file = f; h->writing = true;
	memset(h->fileHeader, 0, FILE_HEADER_BYTES);
	putU32(h->fileHeader, PS3D_OBJX_VERSION);
	putU32(h->fileHeader + 8, MESH_HEADER_BYTES);
	if(scene)
	{
		uint8_t* p = h->fileHeader + 16;
		for(int i = 0; i < 4; i++) { putF32(p + 4 * i, scene->camera_pos[i]); putF32(p + 16 + 4 * i, scene->camera_ypr[i]); }
		for(int l = 0; l < PS3D_OBJX_MAXLIGHTS; l++)
			for(int i = 0; i < 4; i++) { putF32(p + 32 + 32 * l + 4 * i, scene->light_pos[l][i]); putF32(p + 32 + 32 * l + 16 + 4 * i, scene->light_dir[l][i]); }
	}
	if(1 != fwrite(h->fileHeader, FILE_HEADER_BYTES, 1, f)) { fclose(f); delete h; return PS3D_OBJX_ERR_IO; }
	*out = h;
	return PS3D_OBJX_OK;
}

int ps3d_objx_write_mesh(ps3d_objx* h, const ps3d_objx_mesh* mesh, const float* vertices, const float* normals,
                         const float* tangents, const float* texcoords, const int32_t* indices)
{
	if(!h || !mesh) return PS3D_OBJX_ERR_ARGUMENT;
	if(!h->writing) return PS3D_OBJX_ERR_STATE;
	// write_mesh's argument check, objxio.cpp:112-120: vertices are mandatory; indices and their count go together
	if(0 == mesh->num_vertices || !vertices || (0 != mesh->num_indices && !indices) || (0 == mesh->num_indices && indices)) return PS3D_OBJX_ERR_ARGUMENT;
	uint8_t raw[MESH_HEADER_BYTES];
	memset(raw, 0, sizeof(raw));
	putStr(raw + MH_NAME, mesh->mesh_name);
	putU32(raw + MH_NUM_VERTICES, mesh->num_vertices);
	putU32(raw + MH_NUM_INDICES, mesh->num_indices);
	raw[MH_HAS_TEXCOORDS] = texcoords ? 1 : 0; raw[MH_HAS_NORMALS] = normals ? 1 : 0; raw[MH_HAS_TANGENTS] = tangents ? 1 : 0;
	putColour(raw + MH_AMBIENT, mesh->ambient_colour);
	putColour(raw + MH_DIFFUSE, mesh->diffuse_colour);
	putColour(raw + MH_SPECULAR, mesh->specular_colour);
	putF32(raw + MH_SPECULAR_EXP, mesh->specular_exponent);
	putStr(raw + MH_DIFFUSE_FILE, mesh->diffuse_file);
	putStr(raw + MH_BUMP_FILE, mesh->bump_file);
	putStr(raw + MH_SPC_FILE, mesh->spc_file);
	putStr(raw + MH_SPE_FILE, mesh->spe_file);
	putStr(raw + MH_PROGRAMME, mesh->programme);
	putU32(raw + MH_NEXT_OFFSET, (uint32_t)payloadBytes(mesh->num_vertices, mesh->num_indices, texcoords != nullptr, normals != nullptr, tangents != nullptr));
	FILE* f = h->file;
	const size_t v16 = (size_t)mesh->num_vertices * 16;
	bool ok = 1 == fwrite(raw, MESH_HEADER_BYTES, 1, f);
	ok = ok && 1 == fwrite(vertices, v16, 1, f);
	if(normals) ok = ok && 1 == fwrite(normals, v16, 1, f);
	if(tangents) ok = ok && 1 == fwrite(tangents, v16, 1, f);
	if(texcoords) ok = ok && 1 == fwrite(texcoords, (size_t)mesh->num_vertices * 8, 1, f);
	if(indices) ok = ok && 1 == fwrite(indices, (size_t)mesh->num_indices * 4, 1, f);
	if(!ok) return PS3D_OBJX_ERR_IO;
	h->numMeshes++;
	return PS3D_OBJX_OK;
}

int ps3d_objx_close(ps3d_objx* h)
{
	if(!h) return PS3D_OBJX_ERR_ARGUMENT;
	int rc = PS3D_OBJX_OK;
	if(h->writing)
	{
		// the mesh count is only known now
		putU32(h->fileHeader + 4, h->numMeshes);
		rewind(h->file);
		if(1 != fwrite(h->fileHeader, FILE_HEADER_BYTES, 1, h->file)) rc = PS3D_OBJX_ERR_IO;
	}
	if(0 != fclose(h->file)) rc = PS3D_OBJX_ERR_IO;
	delete h;
	return rc;
}

} // extern "C"
