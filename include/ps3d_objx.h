/* ps3d_objx.h — C-ABI of the OBJX scene-file reader/writer (the data format on the input side
 * of the hot path). Host-side I/O, no device work; implemented in puresoft3d_b200/csrc/objx.cpp inside libps3d_b200.so.
 *
 * Replaces the reference's libobjx interface, /root/reference/src/objcvt/objxio.h:42-52:
 *   open_objxA / get_mesh_count / read_mesh_header / read_mesh / close_objx   -> ps3d_objx_open / _mesh_count /
 *                                                                                _read_mesh_header / _read_mesh / _close
 *   create_objxA / write_mesh / close_objx                                    -> ps3d_objx_create / _write_mesh / _close
 * mesh_info (objxio.h:5-27) becomes ps3d_objx_mesh (fixed-size strings instead of std::string) + plain arrays,
 * scene_desc (objxio.h:31-38) becomes ps3d_objx_scene.
 *
 * File layout (objxio.cpp:6-45, packed, little endian), version 0x010003:
 *   objx_header  : u32 version, u32 num_meshes, u32 mesheader_size, u32 reserved, f32 camera_position[4],
 *                  f32 camera_ypr[4], 4 x { f32 light_position[4], f32 light_direction[4] }
 *   per mesh     : mesh_header (mesheader_size bytes on disk, at least 1627): char name[260], u32 num_vertices,
 *                  u32 num_indices, u8 has_texcoords, u8 has_normals, u8 has_tangents, f32 ambient[4], diffuse[4],
 *                  specular[4] (stored b,g,r,a), f32 specularExp, char diffuse_file[260], bump_file[260], spc_file[260],
 *                  spe_file[260], programme[260], u32 next_offset (payload bytes behind the header);
 *                  payload: vertices float4[n], normals float4[n]?, tangents float4[n]?, texcoords float2[n]?, indices i32[m]?
 */
#ifndef PS3D_OBJX_H
#define PS3D_OBJX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PS3D_OBJX_VERSION 0x010003u
#define PS3D_OBJX_NAMELEN 260
#define PS3D_OBJX_MAXLIGHTS 4
enum { PS3D_OBJX_LT_OMNI = 0, PS3D_OBJX_LT_DIR = 1 };

typedef struct ps3d_objx ps3d_objx;

typedef struct ps3d_objx_scene   /* scene_desc, objxio.h:31-38 */
{
	float camera_pos[4];
	float camera_ypr[4];
	float light_pos[PS3D_OBJX_MAXLIGHTS][4];
	float light_dir[PS3D_OBJX_MAXLIGHTS][4];
	int light_types[PS3D_OBJX_MAXLIGHTS];      /* LT_OMNI when |light_dir| < 1e-7 (objxio.cpp:216-222) */
} ps3d_objx_scene;

typedef struct ps3d_objx_mesh    /* mesh_info without its arrays, objxio.h:5-27 */
{
	char mesh_name[PS3D_OBJX_NAMELEN];
	uint32_t num_vertices, num_indices;
	int has_texcoords, has_normals, has_tangents;
	/* x,y,z,w = r,g,b,a: read_mesh_header swaps the stored b,g,r,a and clamps to [0,1] (objxio.cpp:272-283) */
	float ambient_colour[4], diffuse_colour[4], specular_colour[4];
	float specular_exponent;
	char diffuse_file[PS3D_OBJX_NAMELEN], bump_file[PS3D_OBJX_NAMELEN], spc_file[PS3D_OBJX_NAMELEN],
	     spe_file[PS3D_OBJX_NAMELEN], programme[PS3D_OBJX_NAMELEN];
} ps3d_objx_mesh;

/* 0 = ok; negative = error (PS3D_OBJX_ERR_*) */
enum { PS3D_OBJX_OK = 0, PS3D_OBJX_ERR_IO = -1, PS3D_OBJX_ERR_FORMAT = -2, PS3D_OBJX_ERR_ARGUMENT = -3, PS3D_OBJX_ERR_STATE = -4 };

/* open_objxA, objxio.cpp:182-231: fails on a wrong version or a mesh header smaller than this build's */
int ps3d_objx_open(const char* filename, ps3d_objx_scene* scene, ps3d_objx** out);
int ps3d_objx_mesh_count(ps3d_objx* h);                                         /* get_mesh_count, :233-238 */
/* read_mesh_header, :240-290: the next mesh's header; must be followed by ps3d_objx_read_mesh */
int ps3d_objx_read_mesh_header(ps3d_objx* h, ps3d_objx_mesh* mesh);
/* read_mesh, :292-362: any array may be NULL (skipped). vertices/normals/tangents: 4 floats per vertex, texcoords: 2,
 * indices: num_indices ints. If `tangents` is wanted, the file has none but has texcoords (and `texcoords` is given),
 * they are generated from positions and texcoords (generate_tangents, :418-470). Leaves the file at the next mesh. */
int ps3d_objx_read_mesh(ps3d_objx* h, float* vertices, float* normals, float* tangents, float* texcoords, int32_t* indices);
/* create_objxA + write_mesh, :64-180 (colours given r,g,b,a, stored b,g,r,a) */
int ps3d_objx_create(const char* filename, const ps3d_objx_scene* scene, ps3d_objx** out);
int ps3d_objx_write_mesh(ps3d_objx* h, const ps3d_objx_mesh* mesh, const float* vertices, const float* normals,
                         const float* tangents, const float* texcoords, const int32_t* indices);
/* close_objx, :364-376: an output file gets its header (mesh count) rewritten */
int ps3d_objx_close(ps3d_objx* h);

#ifdef __cplusplus
}
#endif
#endif
