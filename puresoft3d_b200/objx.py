"""OBJX scene files and picture loading — the data formats on the input side of the hot path.

`read_objx` / `write_objx` call the native reader/writer inside libps3d_b200.so (csrc/objx.cpp, include/ps3d_objx.h —
the replacement of the reference's libobjx, src/objcvt/objxio.h:42-52). `load_scene_meshes` is what the demos do with a
file before anything reaches a VBO (src/test2/loadscene.cpp:137-345, src/test/scenobj.cpp:88-163): w = 1, binormal =
normalise(normal x tangent), and for demo 2 the world -> model re-centring by bounding boxes with the object/mesh
translation tree. `load_picture` is the picture loader's contract (src/puresoft3d/picldr.cpp:33-118): any format the
platform decoder reads (GDI+ there, Pillow here), flipped vertically, 32-bit BGRA, scanline = width * 4.
"""
import ctypes as C
import os

import numpy as np

from . import _capi

F32 = np.float32
NAMELEN = 260


class ObjxScene(C.Structure):
    _fields_ = [("camera_pos", C.c_float * 4), ("camera_ypr", C.c_float * 4), ("light_pos", (C.c_float * 4) * 4),
                ("light_dir", (C.c_float * 4) * 4), ("light_types", C.c_int * 4)]


class ObjxMesh(C.Structure):
    _fields_ = [("mesh_name", C.c_char * NAMELEN), ("num_vertices", C.c_uint32), ("num_indices", C.c_uint32),
                ("has_texcoords", C.c_int), ("has_normals", C.c_int), ("has_tangents", C.c_int),
                ("ambient_colour", C.c_float * 4), ("diffuse_colour", C.c_float * 4), ("specular_colour", C.c_float * 4),
                ("specular_exponent", C.c_float),
                ("diffuse_file", C.c_char * NAMELEN), ("bump_file", C.c_char * NAMELEN), ("spc_file", C.c_char * NAMELEN),
                ("spe_file", C.c_char * NAMELEN), ("programme", C.c_char * NAMELEN)]


_H = C.c_void_p
_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int32)
PROTOTYPES = {
    "ps3d_objx_open": (C.c_int, [C.c_char_p, C.POINTER(ObjxScene), C.POINTER(_H)]),
    "ps3d_objx_mesh_count": (C.c_int, [_H]),
    "ps3d_objx_read_mesh_header": (C.c_int, [_H, C.POINTER(ObjxMesh)]),
    "ps3d_objx_read_mesh": (C.c_int, [_H, _FP, _FP, _FP, _FP, _IP]),
    "ps3d_objx_create": (C.c_int, [C.c_char_p, C.POINTER(ObjxScene), C.POINTER(_H)]),
    "ps3d_objx_write_mesh": (C.c_int, [_H, C.POINTER(ObjxMesh), _FP, _FP, _FP, _FP, _IP]),
    "ps3d_objx_close": (C.c_int, [_H]),
}
_lib = None


def lib():
    """The product library (the OBJX entry points are host code: they load and run without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_capi.PRODUCT_LIB):
            raise RuntimeError("puresoft3d_b200: %s is missing (build it: python -c 'import __graft_entry__ as g; g.build()')" % _capi.PRODUCT_LIB)
        l = C.CDLL(_capi.PRODUCT_LIB)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


class ObjxError(IOError):
    pass


def _ck(rc, what):
    if rc != 0:
        raise ObjxError("%s failed (%d)" % (what, rc))


def _ptr(a, t=_FP):
    return a.ctypes.data_as(t) if a is not None else None


def read_objx(path, want_tangents=True):
    """-> (scene dict, [mesh dict]). Arrays: vertices/normals/tangents (n,4) float32, texcoords (n,2), indices (m,) int32
    or None where the file has none. Tangents are generated for meshes with texcoords but no stored tangents."""
    L = lib()
    scn, h = ObjxScene(), _H()
    _ck(L.ps3d_objx_open(os.fsencode(path), C.byref(scn), C.byref(h)), "open " + str(path))
    try:
        scene = {"camera_pos": np.array(scn.camera_pos, F32), "camera_ypr": np.array(scn.camera_ypr, F32),
                 "light_pos": np.array(scn.light_pos, F32), "light_dir": np.array(scn.light_dir, F32),
                 "light_types": list(scn.light_types)}
        meshes = []
        for _ in range(L.ps3d_objx_mesh_count(h)):
            mh = ObjxMesh()
            _ck(L.ps3d_objx_read_mesh_header(h, C.byref(mh)), "read_mesh_header")
            n, m = mh.num_vertices, mh.num_indices
            v = np.empty((n, 4), F32)
            nr = np.empty((n, 4), F32) if mh.has_normals else None
            uv = np.empty((n, 2), F32) if mh.has_texcoords else None
            tg = np.empty((n, 4), F32) if (mh.has_tangents or (want_tangents and mh.has_texcoords)) else None
            ix = np.empty((m,), np.int32) if m else None
            _ck(L.ps3d_objx_read_mesh(h, _ptr(v), _ptr(nr), _ptr(tg), _ptr(uv), _ptr(ix, _IP)), "read_mesh")
            meshes.append({
                "name": mh.mesh_name.decode("latin-1"), "vertices": v, "normals": nr, "tangents": tg, "texcoords": uv, "indices": ix,
                "ambient": np.array(mh.ambient_colour, F32), "diffuse": np.array(mh.diffuse_colour, F32),
                "specular": np.array(mh.specular_colour, F32), "specular_exponent": float(mh.specular_exponent),
                "diffuse_file": mh.diffuse_file.decode("latin-1"), "bump_file": mh.bump_file.decode("latin-1"),
                "spc_file": mh.spc_file.decode("latin-1"), "spe_file": mh.spe_file.decode("latin-1"),
                "programme": mh.programme.decode("latin-1"), "stored_tangents": bool(mh.has_tangents)})
        return scene, meshes
    finally:
        L.ps3d_objx_close(h)


def write_objx(path, scene, meshes):
    """The inverse of read_objx (create_objx + write_mesh + close_objx, objxio.cpp:64-180)."""
    L = lib()
    scn = ObjxScene()
    for k in range(4):
        scn.camera_pos[k] = float(scene.get("camera_pos", (0, 0, 0, 0))[k])
        scn.camera_ypr[k] = float(scene.get("camera_ypr", (0, 0, 0, 0))[k])
        for l in range(4):
            scn.light_pos[l][k] = float(np.asarray(scene.get("light_pos", np.zeros((4, 4))))[l][k])
            scn.light_dir[l][k] = float(np.asarray(scene.get("light_dir", np.zeros((4, 4))))[l][k])
    h = _H()
    _ck(L.ps3d_objx_create(os.fsencode(path), C.byref(scn), C.byref(h)), "create " + str(path))
    try:
        for m in meshes:
            mh = ObjxMesh()
            mh.mesh_name = m["name"].encode("latin-1")
            v = np.ascontiguousarray(m["vertices"], F32)
            mh.num_vertices = v.shape[0]
            ix = None if m.get("indices") is None else np.ascontiguousarray(m["indices"], np.int32)
            mh.num_indices = 0 if ix is None else ix.shape[0]
            for key, field in (("ambient", mh.ambient_colour), ("diffuse", mh.diffuse_colour), ("specular", mh.specular_colour)):
                for k in range(4):
                    field[k] = float(np.asarray(m.get(key, (0, 0, 0, 0)))[k])
            mh.specular_exponent = float(m.get("specular_exponent", 0.0))
            for key in ("diffuse_file", "bump_file", "spc_file", "spe_file", "programme"):
                setattr(mh, key, m.get(key, "").encode("latin-1"))
            arrs = [None if m.get(k) is None else np.ascontiguousarray(m[k], F32) for k in ("normals", "tangents", "texcoords")]
            _ck(L.ps3d_objx_write_mesh(h, C.byref(mh), _ptr(v), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(ix, _IP)), "write_mesh")
    finally:
        _ck(L.ps3d_objx_close(h), "close")


# ---- what the demos do with a mesh before it reaches a VBO ---------------------------------------------------------

def _norm3(v):
    """mcemaths_norm_3_4 up to its rsqrtss approximation: host-side preprocessing, the bytes that reach the VBOs are the
    same for every backend (they are produced once, here)."""
    out = v.copy()
    l = np.sqrt((v[:, :3].astype(np.float64) ** 2).sum(axis=1))
    l[l == 0] = 1.0
    out[:, :3] = (v[:, :3] / l[:, None]).astype(F32)
    return out


def binormals(normals, tangents):
    """binormal = normalise(normal x tangent), loadscene.cpp:181-189 / scenobj.cpp:120-129 (w = 0)."""
    b = np.zeros_like(normals)
    b[:, :3] = np.cross(normals[:, :3].astype(np.float64), tangents[:, :3].astype(np.float64)).astype(F32)
    return _norm3(b)


def mesh_slots(mesh):
    """The demos' VAO layout for one mesh (scenobj.cpp:131-158 / loadscene.cpp:289-311): slot 0 position (w forced to 1),
    1 tangent, 2 binormal, 3 normal, 4 uv; slots 1, 2, 4 only with texcoords, slot 3 only with normals."""
    v = mesh["vertices"].copy()
    v[:, 3] = 1.0
    slots = {0: (16, v)}
    if mesh["texcoords"] is not None:
        slots[1] = (16, np.ascontiguousarray(mesh["tangents"], F32))
        slots[2] = (16, binormals(mesh["normals"], mesh["tangents"]))
        slots[4] = (8, np.ascontiguousarray(mesh["texcoords"], F32))
    if mesh["normals"] is not None:
        slots[3] = (16, np.ascontiguousarray(mesh["normals"], F32))
    return slots


def load_scene_meshes(path):
    """Demo 2's loader, loadscene.cpp:137-345: meshes named 'object/mesh' (others are skipped), positions moved from world
    to model space around the centre of each mesh's bounding box, object translation = mean of its meshes' centres, mesh
    translation relative to its object. -> (scene dict, [component dict]) in the reference's iteration order (std::map:
    objects by name, meshes by name); a component's model matrix is translate(object) * translate(mesh)."""
    scene, meshes = read_objx(path)
    objects = {}
    for m in meshes:
        if "/" not in m["name"]:
            continue
        obj, name = m["name"].split("/", 1)
        v = m["vertices"]
        lo, hi = v[:, :3].min(axis=0), v[:, :3].max(axis=0)
        centre = ((lo.astype(np.float64) + hi.astype(np.float64)) / 2.0).astype(F32)
        m = dict(m)
        vv = v.copy()
        vv[:, :3] = v[:, :3] - centre
        vv[:, 3] = 1.0
        m["vertices"] = vv
        m["centre"] = centre
        objects.setdefault(obj, {})[name] = m
    out = []
    for obj in sorted(objects):
        ms = objects[obj]
        otrans = (np.sum([ms[k]["centre"].astype(np.float64) for k in ms], axis=0) / len(ms)).astype(F32)
        for name in sorted(ms):
            m = ms[name]
            m["object"], m["component"] = obj, obj + "/" + name
            m["object_translation"] = otrans
            m["mesh_translation"] = (m["centre"] - otrans).astype(F32)
            m["world_translation"] = (otrans + m["mesh_translation"]).astype(F32)
            out.append(m)
    return scene, out


# ---- pictures -------------------------------------------------------------------------------------------------------

def load_picture(path_or_file):
    """PuresoftDefaultPictureLoader::loadFromFile + retrievePixel (picldr.cpp:33-118): decode, flip vertically (row 0 = the
    image's bottom row, :47), 32-bit BGRA words in memory (PixelFormat32bppARGB, :102-110), scanline = width * 4.
    -> (H, W, 4) uint8 array in B, G, R, A order."""
    from PIL import Image
    img = Image.open(path_or_file).convert("RGBA")
    a = np.asarray(img, dtype=np.uint8)[::-1]                 # RotateNoneFlipY
    return np.ascontiguousarray(a[:, :, [2, 1, 0, 3]])       # R,G,B,A -> B,G,R,A
