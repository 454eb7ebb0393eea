"""OBJX reader/writer and picture loading (include/ps3d_objx.h replaces src/objcvt/objxio.h:42-52).
CPU only: the entry points are host code inside libps3d_b200.so. Where /root/reference is present (this container) the
reference's own fixtures are read, rewritten byte for byte, and demo 2's frame built from plane.objx with the real
pictures is rendered by the oracle AND by the unmodified reference build — bit-identical, which pins the loader-side
preprocessing on the reference's real assets."""
import ctypes
import io
import os
import re
import struct

import numpy as np
import pytest

from _compare import render_all
from conftest import PRODUCT_SO, ROOT
from puresoft3d_b200 import objx, scenes

REF_TEST = "/root/reference/src/test"
REF_TEST2 = "/root/reference/src/test2"
F32 = np.float32


def test_header_symbols_are_exported_and_bound(built):
    text = open(os.path.join(ROOT, "include", "ps3d_objx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ps3d_objx_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(objx.PROTOTYPES)
    lib = ctypes.CDLL(PRODUCT_SO)
    assert not [s for s in declared if not hasattr(lib, s)]


def _tri_mesh(name, n_tris, rng, **kw):
    v = rng.standard_normal((n_tris * 3, 4)).astype(F32)
    m = {"name": name, "vertices": v, "normals": rng.standard_normal(v.shape).astype(F32),
         "tangents": rng.standard_normal(v.shape).astype(F32), "texcoords": rng.random((v.shape[0], 2)).astype(F32),
         "ambient": (0.1, 0.2, 0.3, 0.4), "diffuse": (0.5, 0.6, 0.7, 0.8), "specular": (1.0, 0.9, 0.8, 0.7), "specular_exponent": 33.5,
         "diffuse_file": "d.png", "bump_file": "b.png", "spc_file": "s.png", "spe_file": "e.png", "programme": "VP_X:IP_Y:FP_Z"}
    m.update(kw)
    return m


def test_round_trip_all_fields(tmp_path, built):
    rng = np.random.default_rng(1)
    scene = {"camera_pos": (1, 2, 3, 0), "camera_ypr": (0.1, 0.2, 0.3, 0),
             "light_pos": np.arange(16, dtype=F32).reshape(4, 4), "light_dir": np.array([[0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0], [1e-9, 0, 0, 0]], F32)}
    meshes = [_tri_mesh("obj/a", 5, rng), _tri_mesh("obj/b", 1, rng, normals=None, tangents=None, texcoords=None),
              _tri_mesh("idx/c", 4, rng, indices=np.array([0, 1, 2, 2, 1, 3, 11, 10, 9], np.int32))]
    path = tmp_path / "rt.objx"
    objx.write_objx(path, scene, meshes)
    got_scene, got = objx.read_objx(path)
    assert np.array_equal(got_scene["camera_pos"], np.array(scene["camera_pos"], F32))
    assert np.array_equal(got_scene["light_pos"], scene["light_pos"])
    assert got_scene["light_types"] == [0, 1, 0, 0]           # LT_OMNI when |dir| < 1e-7 (objxio.cpp:216-222)
    assert [m["name"] for m in got] == ["obj/a", "obj/b", "idx/c"]
    for want, have in zip(meshes, got):
        for k in ("vertices", "normals", "texcoords", "indices"):
            assert (want.get(k) is None) == (have[k] is None), k
            if want.get(k) is not None:
                assert np.array_equal(np.asarray(want[k]), have[k]), k
        if want.get("tangents") is not None:
            assert np.array_equal(want["tangents"], have["tangents"]) and have["stored_tangents"]
        for k in ("ambient", "diffuse", "specular"):
            assert np.array_equal(np.array(want[k], F32), have[k]), k
        assert have["specular_exponent"] == 33.5
        for k in ("diffuse_file", "bump_file", "spc_file", "spe_file", "programme"):
            assert have[k] == want[k]
    # on disk: packed little-endian layout of objxio.cpp:6-45; colours stored b,g,r,a
    raw = path.read_bytes()
    assert struct.unpack_from("<4I", raw, 0) == (0x010003, 3, 1627, 0)
    assert struct.unpack_from("<4f", raw, 176 + 271) == pytest.approx((0.3, 0.2, 0.1, 0.4))
    assert struct.unpack_from("<I", raw, 176 + 1623)[0] == 15 * (16 + 16 + 16 + 8)


def test_colours_are_clamped_on_read_and_longer_headers_are_skipped(tmp_path, built):
    rng = np.random.default_rng(2)
    path = tmp_path / "a.objx"
    objx.write_objx(path, {}, [_tri_mesh("o/m", 2, rng, diffuse=(1.5, -0.25, 0.5, 2.0)), _tri_mesh("o/n", 1, rng)])
    assert np.array_equal(objx.read_objx(path)[1][0]["diffuse"], np.array([1.0, 0.0, 0.5, 1.0], F32))   # objxio.cpp:272-283
    # a file written by a NEWER writer: mesh headers 40 bytes longer than ours (mesheader_size, objxio.cpp:196-203, :264)
    raw = bytearray(path.read_bytes())
    struct.pack_into("<I", raw, 8, 1627 + 40)
    first_payload = struct.unpack_from("<I", raw, 176 + 1623)[0]
    second = 176 + 1627 + first_payload
    raw[second + 1627:second + 1627] = b"\xAB" * 40
    raw[176 + 1627:176 + 1627] = b"\xCD" * 40
    newer = tmp_path / "newer.objx"
    newer.write_bytes(bytes(raw))
    a, b = objx.read_objx(path)[1], objx.read_objx(newer)[1]
    assert [m["name"] for m in b] == ["o/m", "o/n"]
    assert all(np.array_equal(x["vertices"], y["vertices"]) and np.array_equal(x["texcoords"], y["texcoords"]) for x, y in zip(a, b))


def test_bad_files_are_refused(tmp_path, built):
    rng = np.random.default_rng(3)
    good = tmp_path / "g.objx"
    objx.write_objx(good, {}, [_tri_mesh("o/m", 2, rng)])
    raw = bytearray(good.read_bytes())
    with pytest.raises(objx.ObjxError):
        objx.read_objx(tmp_path / "missing.objx")
    bad = tmp_path / "v.objx"
    struct.pack_into("<I", raw, 0, 0x010002)
    bad.write_bytes(bytes(raw))
    with pytest.raises(objx.ObjxError):
        objx.read_objx(bad)                                   # wrong version (objxio.cpp:196)
    struct.pack_into("<I", raw, 0, 0x010003)
    struct.pack_into("<I", raw, 8, 100)
    bad.write_bytes(bytes(raw))
    with pytest.raises(objx.ObjxError):
        objx.read_objx(bad)                                   # mesh header smaller than this build's (:197)
    struct.pack_into("<I", raw, 8, 1627)
    bad.write_bytes(bytes(raw[:len(raw) - 10]))
    with pytest.raises(objx.ObjxError):
        objx.read_objx(bad)                                   # truncated payload
    with pytest.raises(objx.ObjxError):
        objx.write_objx(tmp_path / "e.objx", {}, [{"name": "o/e", "vertices": np.zeros((0, 4), F32)}])   # write_mesh refuses empty meshes (:112-120)


def test_tangents_are_generated_when_the_file_has_none(tmp_path, built):
    """u runs along +x scaled by 2, v along +y: T = d(position)/du = (0.5, 0, 0) -> normalised (1, 0, 0); the rotated
    triangle has u along +z."""
    v = np.array([[0, 0, 0, 0], [2, 0, 0, 0], [0, 3, 0, 0], [0, 0, 0, 0], [0, 0, 1, 0], [0, 5, 0, 0]], F32)
    uv = np.array([[0, 0], [4, 0], [0, 1], [0.5, 0.5], [1.5, 0.5], [0.5, 2.5]], F32)
    n = np.tile(np.array([0, 0, 1, 0], F32), (6, 1))
    path = tmp_path / "t.objx"
    objx.write_objx(path, {}, [{"name": "o/flat", "vertices": v, "normals": n, "texcoords": uv},
                               {"name": "o/indexed", "vertices": v[:3], "normals": n[:3], "texcoords": uv[:3], "indices": np.array([0, 1, 2], np.int32)}])
    flat, indexed = objx.read_objx(path)[1]
    assert not flat["stored_tangents"]
    assert np.allclose(flat["tangents"][:3], [[1, 0, 0, 0]] * 3, atol=1e-6)
    assert np.allclose(flat["tangents"][3:], [[0, 0, 1, 0]] * 3, atol=1e-6)
    assert np.allclose(indexed["tangents"], [[1, 0, 0, 0]] * 3, atol=1e-6)
    assert objx.read_objx(path, want_tangents=False)[1][0]["tangents"] is None


def test_mesh_slots_follow_the_demo_layout(built):
    rng = np.random.default_rng(4)
    m = _tri_mesh("o/m", 3, rng)
    slots = objx.mesh_slots(m)
    assert sorted(slots) == [0, 1, 2, 3, 4] and [slots[k][0] for k in sorted(slots)] == [16, 16, 16, 16, 8]
    assert np.all(slots[0][1][:, 3] == 1.0)                                              # scenobj.cpp:114-117
    b = slots[2][1]
    want = np.cross(m["normals"][:, :3].astype(np.float64), m["tangents"][:, :3].astype(np.float64))
    want /= np.linalg.norm(want, axis=1)[:, None]
    assert np.allclose(b[:, :3], want, atol=1e-6) and np.all(b[:, 3] == 0)               # scenobj.cpp:120-129
    bare = objx.mesh_slots(_tri_mesh("o/p", 1, rng, normals=None, tangents=None, texcoords=None))
    assert sorted(bare) == [0]


def test_load_picture_is_flipped_bgra(built):
    from PIL import Image
    rgba = np.zeros((3, 2, 4), np.uint8)
    rgba[0, 0] = (255, 0, 0, 255)        # top-left red
    rgba[2, 1] = (0, 0, 255, 128)        # bottom-right blue, half transparent
    buf = io.BytesIO()
    Image.fromarray(rgba, "RGBA").save(buf, format="PNG")
    buf.seek(0)
    pix = objx.load_picture(buf)
    assert pix.shape == (3, 2, 4) and pix.dtype == np.uint8
    assert tuple(pix[2, 0]) == (0, 0, 255, 255)        # row 0 of the buffer is the image's BOTTOM row (picldr.cpp:47); B,G,R,A
    assert tuple(pix[0, 1]) == (255, 0, 0, 128)
    assert pix.view(np.uint32)[2, 0, 0] == 0xFFFF0000  # the 32-bit word is ARGB (PixelFormat32bppARGB, picldr.cpp:102-110)


def test_procedural_demo2_file_loads_like_plane_objx(tmp_path, built):
    path = scenes.write_demo_objx(tmp_path / "demo.objx")
    desc, comps = objx.load_scene_meshes(path)
    names = [c["component"] for c in comps]
    assert names == sorted(names) and "light1/from" in names and "light1/to" in names
    for c in comps:
        lo, hi = c["vertices"][:, :3].min(axis=0), c["vertices"][:, :3].max(axis=0)
        assert np.allclose(lo + hi, 0, atol=1e-5)                   # re-centred on the bounding box (loadscene.cpp:215-233)
        assert np.all(c["vertices"][:, 3] == 1.0)
    sc = scenes.scene_desk_objx(path, 96, 64, shadow=64, tex_size=16)
    assert "light1/from" not in sc.meta["components"]               # marker meshes are erased (loadscene.cpp:437-439)
    assert sc.meta["triangles"] > 0


needs_reference = pytest.mark.skipif(not os.path.isdir(REF_TEST2), reason="the reference's assets are only present in the build container")


@needs_reference
@pytest.mark.parametrize("path,meshes,verts", [(REF_TEST + "/sphere.objx", 1, 1584), (REF_TEST2 + "/plane.objx", 13, 31242)])
def test_reference_fixtures_read_and_rewrite_byte_for_byte(path, meshes, verts, tmp_path, built):
    desc, ms = objx.read_objx(path)
    assert len(ms) == meshes and sum(m["vertices"].shape[0] for m in ms) == verts
    assert all(m["stored_tangents"] and m["indices"] is None for m in ms)
    out = tmp_path / "copy.objx"
    objx.write_objx(out, desc, ms)
    assert out.read_bytes() == open(path, "rb").read()


@needs_reference
def test_demo2_from_plane_objx_oracle_equals_reference(oracle_lib, ref_lib, built):
    sc = scenes.scene_desk_objx(REF_TEST2 + "/plane.objx", 640, 400, shadow=512, picture_dir=REF_TEST2)
    assert [(t["width"], t["height"]) for t in sc.textures[1:]] == [(1500, 1000), (818, 460), (2048, 1356)]   # marmite, penhold, top
    a, b = render_all(oracle_lib, sc), render_all(ref_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    assert a["stats"]["fragments_shaded"] == b["stats"]["fragments_shaded"] > 100000


def test_compiled_replay_is_the_same_frame(oracle_lib, built):
    """scenes.compile_replay (the command list with its ctypes arguments prepared once) against scenes.replay."""
    from _scenes_small import SMALL
    from puresoft3d_b200.pipeline import PuresoftPipeline
    for name in ("c3_demo2_desk", "demo2_post", "c4_blend_overdraw"):
        sc = SMALL[name]()
        want = render_all(oracle_lib, sc, capture=False)
        p = PuresoftPipeline(sc.width, sc.height, lib=oracle_lib)
        up = scenes.upload(p, sc)
        frame = scenes.compile_replay(p, sc, up)
        frame()
        p.finish()
        assert np.array_equal(p.readColour(), want["colour"]), name
        assert np.array_equal(p.readDepth().view(np.uint32), want["depth"].view(np.uint32)), name
        p.close()


@needs_reference
def test_demo1_from_sphere_objx_oracle_equals_reference(oracle_lib, ref_lib, built):
    """Demo 1's frame (src/test/puresoft.cpp:162-206) with the reference's own sphere.objx and pictures: earth (diffuse, dot3,
    specular, night maps), cloud layer (blended, discarding shadow pass), moon, cube-map skybox, projective shadow lookup."""
    sc = scenes.scene_planets(640, 400, shadow=480, sphere_objx=REF_TEST + "/sphere.objx", picture_dir=REF_TEST)
    assert (sc.textures[0]["width"], sc.textures[0]["height"]) == (2048, 1024) and len(sc.textures[7]["layers"]) == 6
    a, b = render_all(oracle_lib, sc), render_all(ref_lib, sc)
    assert np.array_equal(a["colour"], b["colour"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["counts"], b["counts"])
    assert a["stats"]["fragments_shaded"] == b["stats"]["fragments_shaded"] > 300000


def test_corrupted_files_never_crash_the_reader(tmp_path, built):
    """Random byte corruption of a valid file: the native reader either reads it or refuses it (ObjxError) — a header that
    promises more than the file holds is refused before anything is allocated for it."""
    rng = np.random.default_rng(99)
    good = tmp_path / "g.objx"
    objx.write_objx(good, {"camera_pos": (1, 2, 3, 0)}, [_tri_mesh("o/a", 6, rng), _tri_mesh("o/b", 3, rng, indices=np.arange(9, dtype=np.int32))])
    raw = np.frombuffer(good.read_bytes(), dtype=np.uint8)
    refused = read = 0
    for k in range(400):
        bad = raw.copy()
        n = int(rng.integers(1, 6))
        at = rng.integers(0, 176 + 1627 if k % 2 else bad.size, size=n)     # half of the runs aim at the headers
        bad[at] = rng.integers(0, 256, size=n, dtype=np.uint8)
        if k % 7 == 0:
            bad = bad[:int(rng.integers(10, bad.size))]                     # truncation
        path = tmp_path / "bad.objx"
        path.write_bytes(bad.tobytes())
        try:
            _, meshes = objx.read_objx(path)
            read += 1
            assert all(m["vertices"].shape[0] < 10 ** 6 for m in meshes)
        except objx.ObjxError:
            refused += 1
    assert refused > 0 and read > 0
