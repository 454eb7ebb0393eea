"""Seeded synthetic scenes for the BASELINE.json configs, and the replay that feeds one scene to a pipeline.

A scene is plain data (numpy arrays + a command list) so that the CUDA library, the CPU oracle and the shimmed
reference all consume identical bytes — including the uniform matrices, which are built ONCE here in float32
 and never recomputed per backend. The command vocabulary is the reference's frame loop
(src/test/puresoft.cpp:162-206): setUniform / setDepth / clearDepth / setViewport / enable / disable /
useProgramme / drawVAO.

Vertex layout follows the demos (src/test/scenobj.cpp:135-158): slot 0 position float4 (w=1), 1 tangent float4,
2 binormal float4, 3 normal float4, 4 uv float2. Matrices are column-major like mcemath (m[col*4+row]).
"""
import math
import os

import numpy as np

from . import _capi as K

F32 = np.float32


# ---- float32 matrix helpers (host-side uniform builders; both renderers receive the same bytes) ---------------

def mat_identity():
    return np.eye(4, dtype=F32)


def mat_perspective(znear, zfar, aspect, fov_rad):
    """Same layout as mcemaths_make_proj_perspective (src/mcemath/matrxgl.cpp:9-22)."""
    h = F32(1.0 / math.tan(fov_rad / 2.0))
    nd = F32(znear - zfar)
    m = np.zeros((4, 4), dtype=F32)  # m[row, col]
    m[0, 0] = h / F32(aspect)
    m[1, 1] = h
    m[2, 2] = F32(zfar + znear) / nd
    m[3, 2] = F32(-1.0)
    m[2, 3] = F32(2.0) * F32(znear * zfar) / nd
    return m


def mat_translation(x, y, z):
    m = np.eye(4, dtype=F32)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def mat_scaling(x, y, z):
    m = np.eye(4, dtype=F32)
    m[0, 0], m[1, 1], m[2, 2] = x, y, z
    return m


def mat_rotation(axis, rad):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    c, s = math.cos(rad), math.sin(rad)
    x, y, z = a
    r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, 0],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), 0],
                  [0, 0, 0, 1]])
    return r.astype(F32)


def mat_look_at(eye, target, up=(0, 1, 0)):
    e = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - e
    f /= np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, dtype=np.float64))
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ e
    return m.astype(F32)


def colmajor(m):
    """(4,4) row/col matrix -> the 16 floats mcemath expects (column-major)."""
    return np.ascontiguousarray(np.asarray(m, dtype=F32).T).reshape(16)


def vec4(x, y, z, w=0.0):
    return np.array([x, y, z, w], dtype=F32)


def i32(v):
    return np.array([v], dtype=np.int32)


# ---- scene container ------------------------------------------------------------------------------------------

class Scene:
    def __init__(self, name, width, height):
        self.name = name
        self.width = width
        self.height = height
        self.textures = []    # dict(width, height, elemLen, layers=[arrays], wrap)
        self.vaos = []        # dict(slot -> (unitBytes, array))
        self.programmes = []  # (fnV, fnI, fnF)
        self.commands = []    # tuples, see replay()
        self.meta = {}

    def add_texture(self, width, height, elemLen=4, pixels=None, layers=None, wrap=K.WRAP_CLAMP, bilinear=False):
        lay = layers if layers is not None else [pixels]
        self.textures.append(dict(width=width, height=height, elemLen=elemLen, layers=lay, wrap=wrap, bilinear=bilinear))
        return len(self.textures) - 1

    def add_vao(self, slots):
        self.vaos.append(slots)
        return len(self.vaos) - 1

    def add_programme(self, fn_v, fn_i=None, fn_f=None):
        self.programmes.append((fn_v, fn_i if fn_i is not None else fn_v, fn_f if fn_f is not None else fn_v))
        return len(self.programmes) - 1

    def cmd(self, *c):
        self.commands.append(c)

    def vertex_bytes_read(self):
        """Σ over draws of the bytes the bound vertex functor reads."""
        slots_read = {K.FN_DEF01: (0, 3, 4), K.FN_DEF02: (0, 1, 2), K.FN_DEF03: (0, 1, 2, 3, 4), K.FN_DEF04: (0,),
                      K.FN_DEF05: (0,), K.FN_FLATID: (0, 6), K.FN_TEXPROBE: (0, 4), K.FN_PLANET: (0, 1, 2, 3, 4), K.FN_CLOUD: (0, 3, 4),
                      K.FN_CLOUDSHADOW: (0, 4), K.FN_POSITIONONLY: (0,), K.FN_SHADOW2: (0,), K.FN_SINGLECOLOUR: (0, 3),
                      K.FN_DIFFUSEONLY: (0, 3, 4)}
        total, prog = 0, None
        for c in self.commands:
            if c[0] == "use":
                prog = c[1]
            elif c[0] == "draw":
                fn = self.programmes[prog][0]
                for s in slots_read.get(fn, ()):
                    if s in self.vaos[c[1]]:
                        total += self.vaos[c[1]][s][1].nbytes
        return total


class Uploaded:
    def __init__(self):
        self.textures, self.vaos, self.programmes, self.vbos = [], [], [], []


def upload(pipe, scene):
    """Create the scene's resources on `pipe` (one-time; outside any timed region)."""
    up = Uploaded()
    for t in scene.textures:
        first = t["layers"][0]
        idx = pipe.createTexture(t["width"], t["height"], t["elemLen"], pixels=first,
                                 extraLayers=len(t["layers"]) - 1, mode=t["wrap"])
        for li in range(1, len(t["layers"])):
            pipe.uploadTexture(idx, t["layers"][li], layer=li)
        if t.get("bilinear"):
            pipe.setTextureFilter(idx, True)   # extension: the reference build refuses
        up.textures.append(idx)
    for slots in scene.vaos:
        vao = pipe.createVAO()
        for slot, (unit, arr) in sorted(slots.items()):
            a = np.ascontiguousarray(arr)
            vbo = pipe.createVBO(unit, a.nbytes // unit)
            vbo.updateContent(a)
            pipe.attachVBO(vao, slot, vbo)
            up.vbos.append((vbo, a))
        up.vaos.append(vao)
    procs = {}
    from .pipeline import PuresoftProcessor

    def proc(kind, fn):
        if (kind, fn) not in procs:
            p = PuresoftProcessor()
            p.kind, p.functor = kind, fn
            procs[(kind, fn)] = pipe.addProcessor(p)
        return procs[(kind, fn)]

    for (fv, fi, ff) in scene.programmes:
        up.programmes.append(pipe.createProgramme(proc(K.PROC_VERTEX, fv), proc(K.PROC_INTERPOLATION, fi),
                                                  proc(K.PROC_FRAGMENT, ff)))
    return up


def replay(pipe, scene, up, finish=True):
    """One frame: run the command list. ("tex_uniform", slot, sceneTexIdx) sets an int uniform to the pipe's handle.
    finish=False leaves the frame in flight on the pipe's stream (the reference's drawVAO is synchronous)."""
    for c in scene.commands:
        op = c[0]
        if op == "uniform":
            pipe.setUniform(c[1], c[2])
        elif op == "tex_uniform":
            pipe.setUniform(c[1], i32(up.textures[c[2]]))
        elif op == "viewport":
            pipe.setViewport(c[1], c[2])
        elif op == "depth":
            pipe.setDepth(-1 if c[1] < 0 else up.textures[c[1]])
        elif op == "clearDepth":
            pipe.clearDepth(c[1])
        elif op == "clearColour":
            pipe.clearColour(c[1])
        elif op == "enable":
            pipe.enable(c[1])
        elif op == "disable":
            pipe.disable(c[1])
        elif op == "use":
            pipe.useProgramme(up.programmes[c[1]])
        elif op == "draw":
            pipe.drawVAO(up.vaos[c[1]], bool(c[2]) if len(c) > 2 else False)
        elif op == "post":
            from .pipeline import PuresoftPostProcessor
            pp = PuresoftPostProcessor()
            pp.functor = c[1]
            pipe.postProcess(pp)
        else:
            raise ValueError("unknown scene command %r" % (op,))
    if finish:
        pipe.finish()


def compile_replay(pipe, scene, up):
    """The same frame as replay(pipe, scene, up, finish=False) with the host-side work done once: every command becomes a
    bound C-ABI function with its ctypes arguments prepared (uniform bytes kept alive in pinned-down numpy buffers), so a
    frame costs one ctypes call per command. Matters when the frame is short and the host must stay ahead of the GPU
    (C2 on 8 GPUs: 0.2 ms of kernels per rank). Returns frame()."""
    import ctypes as C
    L, h = pipe._lib, pipe._h
    keep, calls = [], []
    for c in scene.commands:
        op = c[0]
        if op in ("uniform", "tex_uniform"):
            a = np.ascontiguousarray(c[2]) if op == "uniform" else i32(up.textures[c[2]])
            keep.append(a)
            calls.append((L.ps3d_set_uniform, (h, int(c[1]), C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes))))
        elif op == "viewport":
            calls.append((L.ps3d_set_viewport, (h, int(c[1]), int(c[2]))))
        elif op == "depth":
            calls.append((L.ps3d_set_depth, (h, -1 if c[1] < 0 else int(up.textures[c[1]]))))
        elif op == "clearDepth":
            calls.append((L.ps3d_clear_depth, (h, C.c_float(c[1]))))
        elif op == "clearColour":
            calls.append((L.ps3d_clear_colour, (h, C.c_uint32(int(c[1]) & 0xFFFFFFFF))))
        elif op == "enable":
            calls.append((L.ps3d_enable, (h, int(c[1]))))
        elif op == "disable":
            calls.append((L.ps3d_disable, (h, int(c[1]))))
        elif op == "use":
            calls.append((L.ps3d_programme_use, (h, int(up.programmes[c[1]]))))
        elif op == "draw":
            calls.append((L.ps3d_draw_vao, (h, int(up.vaos[c[1]]), 1 if (len(c) > 2 and c[2]) else 0)))
        elif op == "post":
            calls.append((L.ps3d_post_process, (h, int(c[1]))))
        else:
            raise ValueError("unknown scene command %r" % (op,))
    check = pipe._check

    def frame():
        for fn, args in calls:
            rc = fn(*args)
            if rc:
                check(rc)
    frame._keep = keep
    return frame


def render(pipe, scene):
    up = upload(pipe, scene)
    replay(pipe, scene, up)
    return up


# ---- texture generators ---------------------------------------------------------------------------------------

def tex_random_bgra(rng, w, h):
    return rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)


def tex_smooth_bgra(rng, w, h, octaves=4, normal_map=False):
    """Smooth noise: neighbouring texels differ by a few /255 (keeps nearest-texel flips inside the colour gate)."""
    acc = np.zeros((h, w, 3), dtype=np.float64)
    amp_sum = 0.0
    for o in range(octaves):
        n = 4 << o
        g = rng.random((n + 1, n + 1, 3))
        ys = np.linspace(0, n, h, endpoint=False)
        xs = np.linspace(0, n, w, endpoint=False)
        y0, x0 = ys.astype(int), xs.astype(int)
        fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
        a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
        b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
        amp = 0.5 ** o
        acc += amp * (a * (1 - fy) + b * fy)
        amp_sum += amp
    acc /= amp_sum
    out = np.zeros((h, w, 4), dtype=np.uint8)
    if normal_map:
        nx, ny = (acc[..., 0] - 0.5) * 0.6, (acc[..., 1] - 0.5) * 0.6
        nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.0))
        # FragmentProcessorDEF03 reads r,g,b = x,y,z (tex1bump1light1.cpp:193-196); memory order is B,G,R,A
        out[..., 2] = np.clip((nx * 0.5 + 0.5) * 255.0, 0, 255)
        out[..., 1] = np.clip((ny * 0.5 + 0.5) * 255.0, 0, 255)
        out[..., 0] = np.clip((nz * 0.5 + 0.5) * 255.0, 0, 255)
    else:
        out[..., :3] = np.clip(40 + acc * 200.0, 0, 255)
    out[..., 3] = 255
    return out


# ---- geometry generators --------------------------------------------------------------------------------------

def cube_mesh():
    """Unit cube, 12 triangles / 36 un-indexed vertices, CCW outward faces."""
    faces = [  # normal, tangent(u), bitangent(v)
        ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (-1, 0, 0), (0, 1, 0)),
        ((1, 0, 0), (0, 0, -1), (0, 1, 0)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
        ((0, 1, 0), (1, 0, 0), (0, 0, -1)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
    ]
    pos, nrm, tan, bin_, uv = [], [], [], [], []
    for n, t, b in faces:
        n, t, b = np.array(n, float), np.array(t, float), np.array(b, float)
        corners = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
        quad = [0.5 * (n + cu * t + cv * b) for cu, cv in corners]
        quv = [((cu + 1) / 2, (cv + 1) / 2) for cu, cv in corners]
        for i in (0, 1, 2, 0, 2, 3):
            pos.append(list(quad[i]) + [1.0])
            nrm.append(list(n) + [0.0])
            tan.append(list(t) + [0.0])
            bin_.append(list(b) + [0.0])
            uv.append(quv[i])
    return (np.array(pos, F32), np.array(tan, F32), np.array(bin_, F32), np.array(nrm, F32), np.array(uv, F32))


def heightfield_mesh(rng, nx, ny, layers, extent=1.0, z_jitter=0.05, shuffle=True):
    """`layers` stacked height-field grids of nx*ny quads (2 triangles each) in the z=const planes facing +z,
    jittered in z, optionally shuffled so submission order is not depth order (C2 / C5 geometry)."""
    tris_pos, tris_uv = [], []
    for L in range(layers):
        z0 = -0.3 * L
        gx = np.linspace(-extent, extent, nx + 1)
        gy = np.linspace(-extent, extent, ny + 1)
        X, Y = np.meshgrid(gx, gy)
        Z = z0 + (rng.random(X.shape) * 2 - 1) * z_jitter
        U = (X + extent) / (2 * extent)
        V = (Y + extent) / (2 * extent)
        P = np.stack([X, Y, Z, np.ones_like(X)], axis=-1)
        T = np.stack([U, V], axis=-1)
        a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]
        ta, tb, tc, td = T[:-1, :-1], T[:-1, 1:], T[1:, 1:], T[1:, :-1]
        t1 = np.stack([a, b, c], axis=2).reshape(-1, 3, 4)
        t2 = np.stack([a, c, d], axis=2).reshape(-1, 3, 4)
        u1 = np.stack([ta, tb, tc], axis=2).reshape(-1, 3, 2)
        u2 = np.stack([ta, tc, td], axis=2).reshape(-1, 3, 2)
        tris_pos.append(np.concatenate([t1, t2]))
        tris_uv.append(np.concatenate([u1, u2]))
    pos = np.concatenate(tris_pos)
    uv = np.concatenate(tris_uv)
    if shuffle:
        perm = rng.permutation(pos.shape[0])
        pos, uv = pos[perm], uv[perm]
    # per-triangle geometric normal -> smooth enough for lighting; tangent along +x, binormal along +y
    e1 = pos[:, 1, :3] - pos[:, 0, :3]
    e2 = pos[:, 2, :3] - pos[:, 0, :3]
    n = np.cross(e1, e2)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    nrm = np.repeat(np.concatenate([n, np.zeros((n.shape[0], 1))], axis=1)[:, None, :], 3, axis=1)
    tan = np.zeros_like(nrm)
    tan[..., 0] = 1.0
    bin_ = np.zeros_like(nrm)
    bin_[..., 1] = 1.0
    nv = pos.shape[0] * 3
    return (pos.reshape(nv, 4).astype(F32), tan.reshape(nv, 4).astype(F32), bin_.reshape(nv, 4).astype(F32),
            nrm.reshape(nv, 4).astype(F32), uv.reshape(nv, 2).astype(F32))


def _std_slots(mesh):
    pos, tan, bin_, nrm, uv = mesh
    return {0: (16, pos), 1: (16, tan), 2: (16, bin_), 3: (16, nrm), 4: (8, uv)}


def _camera_uniforms(sc, width, height, eye, model, mrot, light):
    proj = mat_perspective(0.1, 10.0, width / height, math.radians(60.0))
    view = mat_look_at(eye, (0, 0, 0))
    pv = (proj.astype(np.float64) @ view.astype(np.float64)).astype(F32)
    sc.cmd("uniform", 0, colmajor(proj))
    sc.cmd("uniform", 1, colmajor(view))
    sc.cmd("uniform", 3, colmajor(pv))
    sc.cmd("uniform", 4, colmajor(model))
    sc.cmd("uniform", 5, colmajor(mrot))
    sc.cmd("uniform", 7, vec4(*light))
    sc.cmd("uniform", 8, vec4(*eye))
    return proj, view, pv


# ---- the BASELINE.json configs --------------------------------------------------------------------------------

def scene_cube(width=640, height=480, seed=1, functor=K.FN_DEF01, bilinear=False, wrap=K.WRAP_CLAMP, tex_size=256):
    """C1: textured cube, 12 triangles, DEF01 (per-pixel Blinn-Phong), 256² BGRA texture, depth test on.
    bilinear=True is BASELINE.json's literal C1 (bilinear filtering) — an extension the reference cannot render."""
    rng = np.random.default_rng(seed)
    sc = Scene("C1-cube" + ("-bilinear" if bilinear else ""), width, height)
    tex = sc.add_texture(tex_size, tex_size, 4, tex_random_bgra(rng, tex_size, tex_size), wrap=wrap, bilinear=bilinear)
    vao = sc.add_vao(_std_slots(cube_mesh()))
    prog = sc.add_programme(functor)
    rot = (mat_rotation((0, 1, 0), math.radians(30.0)).astype(np.float64) @ mat_rotation((1, 0, 0), math.radians(20.0)).astype(np.float64)).astype(F32)
    sc.cmd("viewport", width, height)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0xFF202020)
    _camera_uniforms(sc, width, height, (0.0, 0.0, 3.0), rot, rot, (-3.0, 2.0, 3.0))
    sc.cmd("tex_uniform", 9, tex)
    if functor == K.FN_DEF03:
        bump = sc.add_texture(256, 256, 4, tex_smooth_bgra(rng, 256, 256, normal_map=True), bilinear=bilinear)
        sc.cmd("tex_uniform", 10, bump)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    return sc


def scene_texprobe(width=64, height=48, tex=None, bilinear=True, wrap=K.WRAP_CLAMP, uv_scale=1.0, uv_offset=0.0):
    """Sampler probe (parity-test functor TEXPROBE, not in the reference): one screen-filling quad given directly in clip
    space, uv = (x / width, y / height) * uv_scale + uv_offset at pixel (x, y), the 2-D sampler's output written unlit."""
    sc = Scene("texprobe", width, height)
    t = sc.add_texture(tex.shape[1], tex.shape[0], 4, tex, wrap=wrap, bilinear=bilinear)
    lo, hi = uv_offset, uv_offset + uv_scale
    pos = np.array([[-1, -1, 0, 1], [1, -1, 0, 1], [1, 1, 0, 1], [-1, -1, 0, 1], [1, 1, 0, 1], [-1, 1, 0, 1]], dtype=F32)
    uv = np.array([[lo, lo], [hi, lo], [hi, hi], [lo, lo], [hi, hi], [lo, hi]], dtype=F32)
    vao = sc.add_vao({0: (16, pos), 4: (8, uv)})
    prog = sc.add_programme(K.FN_TEXPROBE)
    sc.cmd("viewport", width, height)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0)
    sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    sc.cmd("tex_uniform", 9, t)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    return sc


def scene_heightfield(width=1920, height=1080, grid=354, layers=4, seed=2, tex_size=2048, functor=K.FN_DEF03,
                      shuffle=True):
    """C2 (and C5 at larger sizes): `layers` stacked grids of grid² quads -> 2*grid²*layers triangles; DEF03 with
    a smooth diffuse texture and a normal map (tex_size²)."""
    rng = np.random.default_rng(seed)
    sc = Scene("C2-heightfield-%dx%d-g%d-l%d" % (width, height, grid, layers), width, height)
    diffuse = sc.add_texture(tex_size, tex_size, 4, tex_smooth_bgra(rng, tex_size, tex_size))
    bump = sc.add_texture(tex_size, tex_size, 4, tex_smooth_bgra(rng, tex_size, tex_size, normal_map=True))
    aspect = width / height
    mesh = heightfield_mesh(rng, grid, grid, layers, extent=1.0, shuffle=shuffle)
    vao = sc.add_vao(_std_slots(mesh))
    prog = sc.add_programme(functor)
    model = mat_scaling(1.55 * aspect, 1.55, 1.0)
    sc.cmd("viewport", width, height)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0xFF000000)
    _camera_uniforms(sc, width, height, (0.0, 0.0, 2.6), model, mat_identity(), (-2.0, 1.5, 3.0))
    sc.cmd("tex_uniform", 9, diffuse)
    sc.cmd("tex_uniform", 10, bump)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    sc.meta["triangles"] = mesh[0].shape[0] // 3
    return sc


def scene_blend_overdraw(width=1920, height=1080, seed=4, quads=10, randoms=2000):
    """C4: `quads` full-viewport quads at descending z (back-to-front) then `randoms` mid-size triangles, all
    through the skybox functor's write4 path with ALPHABLEND on (FBOBridge::write4 -> blend4, fragthrd.cpp:70-82),
    depth test + write on (as the cloud draw, src/test/testobjs.cpp:110-112). The cube-map texels carry random
    alpha so blend4's integer formula is exercised over its whole domain."""
    rng = np.random.default_rng(seed)
    sc = Scene("C4-blend-overdraw", width, height)
    layers = [tex_random_bgra(rng, 64, 64) for _ in range(6)]
    tex = sc.add_texture(64, 64, 4, layers=layers)
    pos = []
    for q in range(quads):
        z = 0.9 - 0.15 * q
        sx = 1.0 - 0.02 * q  # shrink slightly so every layer keeps a visible border
        for (x, y) in ((-sx, sx), (-sx, -sx), (sx, -sx), (sx, -sx), (sx, sx), (-sx, sx)):
            pos.append((x, y, z, 1.0))
    for _ in range(randoms):
        c = rng.uniform(-0.9, 0.9, size=2)
        r = rng.uniform(0.02, 0.25)
        ang = rng.uniform(0, 2 * math.pi) + np.array([0.0, 2.1, 4.2]) + rng.uniform(-0.5, 0.5, size=3)
        z = rng.uniform(-0.9, 0.9)
        tri = [(c[0] + r * math.cos(a), c[1] + r * math.sin(a), z + rng.uniform(-0.05, 0.05), 1.0) for a in ang]
        pos.extend(tri)
    pos = np.array(pos, dtype=F32)
    vao = sc.add_vao({0: (16, pos)})
    prog = sc.add_programme(K.FN_DEF04)
    sc.cmd("viewport", width, height)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0xFF404040)
    sc.cmd("uniform", 0, colmajor(mat_identity()))
    sc.cmd("uniform", 1, colmajor(mat_rotation((0, 1, 0), 0.3)))
    sc.cmd("tex_uniform", 2, tex)
    sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    sc.cmd("enable", K.BEHAVIOR_ALPHABLEND)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    sc.cmd("disable", K.BEHAVIOR_ALPHABLEND)
    sc.cmd("enable", K.BEHAVIOR_FACE_CULLING)
    sc.meta["triangles"] = pos.shape[0] // 3
    return sc


def scene_soup(width=320, height=240, seed=7, count=400, functor=K.FN_DEF02, cull=True):
    """Edge-case soup in NDC space (identity matrices): sub-pixel slivers, exact flat tops/bottoms, triangles
    straddling every screen edge, far off-screen ones, z outside [-1,1] (whole-triangle reject, drawvao.cpp:51-56),
    zero-area (NaN normal => not culled, vertthrd.cpp:59-64), mixed winding, w != 1."""
    rng = np.random.default_rng(seed)
    sc = Scene("soup-%d" % seed, width, height)
    pos, col, nrm = [], [], []

    def tri(p, w=(1.0, 1.0, 1.0)):
        for (x, y, z), ww in zip(p, w):
            pos.append((x * ww, y * ww, z * ww, ww))
            col.append(tuple(rng.uniform(30, 255, size=3)) + (255.0,))
            nrm.append((0.0, 0.0, 1.0, 0.0))

    for i in range(count):
        kind = i % 10
        c = rng.uniform(-1.1, 1.1, size=2)
        z = rng.uniform(-0.95, 0.95, size=3)
        if kind == 0:    # tiny sliver
            d = rng.uniform(-0.01, 0.01, size=(3, 2))
        elif kind == 1:  # large
            d = rng.uniform(-1.2, 1.2, size=(3, 2))
        else:
            d = rng.uniform(-0.2, 0.2, size=(3, 2))
        p = [(c[0] + d[k, 0], c[1] + d[k, 1], z[k]) for k in range(3)]
        if kind == 2:    # exact flat bottom on a pixel row
            yy = (rng.integers(0, height) - height // 2) / (height // 2)
            p[0] = (p[0][0], yy, p[0][2])
            p[1] = (p[1][0], yy, p[1][2])
        elif kind == 3:  # exact flat top, fractional y
            p[2] = (p[2][0], p[1][1], p[2][2])
        elif kind == 4:  # z out of range on one vertex
            p[1] = (p[1][0], p[1][1], rng.choice([-1.5, 1.5]))
        elif kind == 5:  # degenerate: repeated vertex
            p[2] = p[1]
        elif kind == 6:  # perspective w
            tri(p, w=tuple(rng.uniform(0.5, 3.0, size=3)))
            continue
        elif kind == 7:  # far off-screen
            p = [(x + 3.0, y, zz) for (x, y, zz) in p]
        tri(p)
    # DEF02 reads slot 0 position, 1 normal, 2 colour (colr1light1.cpp:24-26)
    vao = sc.add_vao({0: (16, np.array(pos, F32)), 1: (16, np.array(nrm, F32)), 2: (16, np.array(col, F32))})
    prog = sc.add_programme(functor)
    ident = colmajor(mat_identity())
    sc.cmd("viewport", width, height)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0xFF101010)
    sc.cmd("uniform", 3, ident)
    sc.cmd("uniform", 4, ident)
    sc.cmd("uniform", 5, ident)
    sc.cmd("uniform", 7, vec4(0.3, 0.4, 2.0))
    sc.cmd("uniform", 8, vec4(0.0, 0.0, 3.0))
    if not cull:
        sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    if not cull:
        sc.cmd("enable", K.BEHAVIOR_FACE_CULLING)
    return sc


# ---- the two demos of the reference as headless scenes (BASELINE.json configs[2]) ----------------

BIAS = np.array([[0.5, 0, 0, 0.5], [0, 0.5, 0, 0.5], [0, 0, 1.0, 0], [0, 0, 0, 1.0]], dtype=F32)   # src/test/puresoft.cpp:38-44


def sphere_mesh(stacks=16, slices=32, radius=1.0):
    """UV sphere, un-indexed triangles, with tangent (d/du), binormal (d/dv), normal and uv like sphere.objx carries them."""
    def vert(i, j):
        v, u = i / stacks, j / slices
        th, ph = math.pi * v, 2 * math.pi * u
        n = np.array([math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)])
        t = np.array([-math.sin(ph), 0.0, math.cos(ph)])
        b = np.cross(n, t)
        return n * radius, t, b, n, (u, 1.0 - v)
    pos, tan, bin_, nrm, uv = [], [], [], [], []
    for i in range(stacks):
        for j in range(slices):
            a, b, c, d = vert(i, j), vert(i + 1, j), vert(i + 1, j + 1), vert(i, j + 1)
            for tri in ((a, c, b), (a, d, c)):
                for (p, t, bb, n, q) in tri:
                    pos.append(list(p) + [1.0]); tan.append(list(t) + [0.0]); bin_.append(list(bb) + [0.0]); nrm.append(list(n) + [0.0]); uv.append(q)
    return (np.array(pos, F32), np.array(tan, F32), np.array(bin_, F32), np.array(nrm, F32), np.array(uv, F32))


def tex_cloud(rng, w, h):
    """Cloud layer: alpha lives in the red channel (FP_Cloud reads .r, FP_CloudShadow discards below 150)."""
    t = tex_smooth_bgra(rng, w, h, octaves=5)
    g = t[..., 0].astype(np.float64)
    g = (g - g.min()) / max(1.0, g.max() - g.min())
    out = np.zeros((h, w, 4), dtype=np.uint8)
    out[..., 2] = np.clip(g * 330.0 - 40.0, 0, 255)
    out[..., 3] = 255
    return out


DEMO1_PICTURES = {   # src/test/testobjs.cpp:12-18, :63-64, :119-122, :188-196 (cube faces: xpos, xneg, ypos, yneg, zpos, zneg)
    "diffuse": "earth.jpg", "bump": "earth.dot3.png", "spec": "earth.spac.png", "night": "earth.night.png", "cloud": "earth.cloud.png",
    "moon_d": "moon.jpg", "moon_b": "moon.dot3.png",
    "sky": ["purplenebula_rt.png", "purplenebula_lf.png", "purplenebula_up.png", "purplenebula_dn.png", "purplenebula_ft.png", "purplenebula_bk.png"],
}


def scene_planets(width=800, height=500, shadow=480, seed=3, stacks=16, slices=32, tex_size=256, sphere_objx=None, picture_dir=None):
    """Demo 1 (src/test/puresoft.cpp:113-206): shadow pass into a float texture (earth + moon with DEF05, cloud layer with the
    discarding CloudShadow triple), then skybox (DEF04, depth off), earth (VP/IP_Planet + FP_Earth), moon (FP_Satellite) and the
    alpha-blended cloud layer (VP/IP/FP_Cloud). Uniform slots as src/test/testobjs.cpp:45-145."""
    rng = np.random.default_rng(seed)
    sc = Scene("demo1-planets-%dx%d%s" % (width, height, "-objx" if sphere_objx else ""), width, height)

    def tex(key, synthetic):
        """The demo's own picture (picture_dir given and the file is there: SceneObject::findOrCreateTexture,
        src/test/scenobj.cpp:165-200, through objx.load_picture) or a seeded synthetic stand-in."""
        pix = synthetic()
        if picture_dir and os.path.exists(os.path.join(picture_dir, DEMO1_PICTURES[key])):
            from . import objx
            pix = objx.load_picture(os.path.join(picture_dir, DEMO1_PICTURES[key]))
        return sc.add_texture(pix.shape[1], pix.shape[0], 4, np.ascontiguousarray(pix))

    diffuse = tex("diffuse", lambda: tex_smooth_bgra(rng, tex_size, tex_size))
    bump = tex("bump", lambda: tex_smooth_bgra(rng, tex_size, tex_size, normal_map=True))
    spec = tex("spec", lambda: tex_smooth_bgra(rng, tex_size, tex_size))
    night = tex("night", lambda: (tex_smooth_bgra(rng, tex_size, tex_size) // 3).astype(np.uint8))
    cloud = tex("cloud", lambda: tex_cloud(rng, tex_size, tex_size))
    moon_d = tex("moon_d", lambda: tex_smooth_bgra(rng, tex_size // 2, tex_size // 2))
    moon_b = tex("moon_b", lambda: tex_smooth_bgra(rng, tex_size // 2, tex_size // 2, normal_map=True))
    faces = [tex_smooth_bgra(rng, 64, 64) for _ in range(6)]
    if picture_dir and all(os.path.exists(os.path.join(picture_dir, f)) for f in DEMO1_PICTURES["sky"]):
        from . import objx
        faces = [np.ascontiguousarray(objx.load_picture(os.path.join(picture_dir, f))) for f in DEMO1_PICTURES["sky"]]
    sky = sc.add_texture(faces[0].shape[1], faces[0].shape[0], 4, layers=faces)
    shadow_tex = sc.add_texture(shadow, shadow, 4, None)   # float depth, created empty (puresoft.cpp:141-147)
    if sphere_objx:
        # SceneObject::findOrCreateVao("sphere.objx"), src/test/scenobj.cpp:88-163: first mesh of the file, w = 1, binormals
        from . import objx
        sphere = sc.add_vao(objx.mesh_slots(objx.read_objx(sphere_objx)[1][0]))
    else:
        sphere = sc.add_vao(_std_slots(sphere_mesh(stacks, slices, radius=0.5)))
    quad = sc.add_vao({0: (16, np.array([(-1, 1, 0, 1), (-1, -1, 0, 1), (1, -1, 0, 1), (1, -1, 0, 1), (1, 1, 0, 1), (-1, 1, 0, 1)], F32))})
    p_shadow = sc.add_programme(K.FN_DEF05)
    p_cloudshadow = sc.add_programme(K.FN_CLOUDSHADOW)
    p_sky = sc.add_programme(K.FN_DEF04)
    p_earth = sc.add_programme(K.FN_PLANET)
    p_moon = sc.add_programme(K.FN_PLANET, K.FN_PLANET, K.FN_SATELLITE)
    p_cloud = sc.add_programme(K.FN_CLOUD)

    light, camera = (-2.0, 0.6, 2.4), (0.0, 0.0, 2.2)
    proj = mat_perspective(0.1, 10.0, width / height, 2 * math.pi * (30.0 / 360.0))
    view = mat_translation(-camera[0], -camera[1], -camera[2])
    pv = (proj.astype(np.float64) @ view.astype(np.float64)).astype(F32)
    lproj = mat_perspective(0.1, 10.0, 1.0, 2 * math.pi * (30.0 / 360.0))
    lview = mat_look_at(light, (0, 0, 0))
    lpv = (lproj.astype(np.float64) @ lview.astype(np.float64)).astype(F32)
    lpvb = (BIAS.astype(np.float64) @ lpv.astype(np.float64)).astype(F32)
    rot_e = mat_rotation((0, 1, 0), 0.7)
    model_e = rot_e
    rot_c = mat_rotation((0, 1, 0), 1.9)
    model_c = (rot_c.astype(np.float64) @ mat_scaling(1.1, 1.1, 1.1).astype(np.float64)).astype(F32)
    rot_m = mat_rotation((0, 1, 0), 2.6)
    model_m = (rot_m.astype(np.float64) @ mat_translation(0.95, 0, 0).astype(np.float64) @ mat_scaling(0.2, 0.2, 0.2).astype(np.float64)).astype(F32)

    def obj_uniforms(model, rot):
        sc.cmd("uniform", 4, colmajor(model))
        sc.cmd("uniform", 5, colmajor(rot))

    sc.cmd("uniform", 7, vec4(*light))
    sc.cmd("uniform", 8, vec4(*camera))
    sc.cmd("clearColour", 0xFF000000)
    # ---- shadow map
    sc.cmd("uniform", 0, colmajor(lproj)); sc.cmd("uniform", 1, colmajor(lview)); sc.cmd("uniform", 3, colmajor(lpv))
    sc.cmd("depth", shadow_tex)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("viewport", shadow, shadow)
    sc.cmd("use", p_shadow)
    obj_uniforms(model_e, rot_e); sc.cmd("draw", sphere)
    obj_uniforms(model_m, rot_m); sc.cmd("draw", sphere)
    obj_uniforms(model_c, rot_c); sc.cmd("tex_uniform", 9, cloud)
    sc.cmd("use", p_cloudshadow); sc.cmd("enable", K.BEHAVIOR_ALPHABLEND); sc.cmd("draw", sphere); sc.cmd("disable", K.BEHAVIOR_ALPHABLEND)
    # ---- the scene
    sc.cmd("uniform", 0, colmajor(proj)); sc.cmd("uniform", 1, colmajor(view)); sc.cmd("uniform", 3, colmajor(pv))
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("viewport", width, height)
    sc.cmd("tex_uniform", 15, shadow_tex)
    sc.cmd("uniform", 16, colmajor(lpvb))
    sc.cmd("tex_uniform", 2, sky)
    sc.cmd("disable", K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH)
    sc.cmd("use", p_sky); sc.cmd("draw", quad, True)
    sc.cmd("enable", K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH)
    obj_uniforms(model_e, rot_e)
    sc.cmd("tex_uniform", 9, diffuse); sc.cmd("tex_uniform", 10, bump); sc.cmd("tex_uniform", 11, spec); sc.cmd("tex_uniform", 12, night)
    sc.cmd("use", p_earth); sc.cmd("draw", sphere)
    obj_uniforms(model_m, rot_m)
    sc.cmd("tex_uniform", 9, moon_d); sc.cmd("tex_uniform", 10, moon_b)
    sc.cmd("use", p_moon); sc.cmd("draw", sphere)
    obj_uniforms(model_c, rot_c)
    sc.cmd("tex_uniform", 9, cloud)
    sc.cmd("use", p_cloud); sc.cmd("enable", K.BEHAVIOR_ALPHABLEND); sc.cmd("draw", sphere); sc.cmd("disable", K.BEHAVIOR_ALPHABLEND)
    sc.meta["triangles"] = 3 * stacks * slices * 2 * 2 + 2
    return sc


def box_mesh(sx, sy, sz):
    pos, tan, bin_, nrm, uv = cube_mesh()
    pos = pos.copy()
    pos[:, 0] *= sx; pos[:, 1] *= sy; pos[:, 2] *= sz
    return pos, tan, bin_, nrm, uv


def scene_desk(width=1024, height=640, shadow=1024, seed=5, clutter=24, tex_size=256, skybox=True):
    """Demo 2 / BASELINE.json configs[2] (src/test2/puresoft.cpp:185-248): a depth-only pass into a float texture through
    VP_Shadow / IP_Null / FP_Null, then the lit pass: a textured desk top and boxes (DiffuseOnly triple: spot light through
    double-precision acos/cos, 4-tap projective shadow lookup), single-colour objects (SingleColour triple), an unlit marker
    at the light (PositionOnly + FP_SingleColourNoLighting), in front of a cube-map skybox (DEF04, depth off)."""
    rng = np.random.default_rng(seed)
    sc = Scene("demo2-desk-%dx%d-s%d" % (width, height, shadow), width, height)
    wood = sc.add_texture(tex_size, tex_size, 4, tex_smooth_bgra(rng, tex_size, tex_size))
    sky = sc.add_texture(64, 64, 4, layers=[tex_smooth_bgra(rng, 64, 64) for _ in range(6)])
    shadow_tex = sc.add_texture(shadow, shadow, 4, None)
    p_shadow = sc.add_programme(K.FN_SHADOW2, K.FN_POSITIONONLY, K.FN_SHADOW2)
    p_tex = sc.add_programme(K.FN_DIFFUSEONLY)
    p_col = sc.add_programme(K.FN_SINGLECOLOUR)
    p_unlit = sc.add_programme(K.FN_POSITIONONLY)
    p_sky = sc.add_programme(K.FN_DEF04)
    quad = sc.add_vao({0: (16, np.array([(-1, 1, 0, 1), (-1, -1, 0, 1), (1, -1, 0, 1), (1, -1, 0, 1), (1, 1, 0, 1), (-1, 1, 0, 1)], F32))})

    light, camera = (1.2, 2.4, 1.6), (0.0, 1.6, 3.2)
    proj = mat_perspective(0.1, 10.0, width / height, math.radians(50.0))
    view = mat_look_at(camera, (0, 0.2, 0))
    pv = (proj.astype(np.float64) @ view.astype(np.float64)).astype(F32)
    lproj = mat_perspective(0.1, 10.0, 1.0, math.radians(60.0))
    lview = mat_look_at(light, (0, 0, 0))
    lpv = (lproj.astype(np.float64) @ lview.astype(np.float64)).astype(F32)
    lpvb = (BIAS.astype(np.float64) @ lpv.astype(np.float64)).astype(F32)
    ldir = np.array(light, dtype=np.float64)
    ldir = ldir / np.linalg.norm(ldir)   # points from the scene to the light, like L in the fragment functor

    objects = []   # (vao, model, rot, programme, colour or None, casts_shadow)
    desk = sc.add_vao(_std_slots(box_mesh(4.0, 0.1, 3.0)))
    objects.append((desk, mat_translation(0, -0.05, 0), mat_identity(), p_tex, None, True))
    for i in range(clutter):
        w, h, d = rng.uniform(0.1, 0.45, size=3)
        vao = sc.add_vao(_std_slots(box_mesh(w, h, d)))
        ang = rng.uniform(0, 2 * math.pi)
        rot = mat_rotation((0, 1, 0), ang)
        x, z = rng.uniform(-1.6, 1.6), rng.uniform(-1.1, 1.1)
        model = (mat_translation(x, h / 2, z).astype(np.float64) @ rot.astype(np.float64)).astype(F32)
        if i % 2:
            objects.append((vao, model, rot, p_col, tuple(rng.uniform(0.2, 0.95, size=3)), True))
        else:
            objects.append((vao, model, rot, p_tex, None, True))
    ball = sc.add_vao(_std_slots(sphere_mesh(12, 24, 0.3)))
    objects.append((ball, mat_translation(-0.4, 0.3, 0.5), mat_identity(), p_col, (0.9, 0.3, 0.2), True))
    marker = sc.add_vao(_std_slots(box_mesh(0.08, 0.08, 0.08)))
    objects.append((marker, mat_translation(*light), mat_identity(), p_unlit, (1.0, 1.0, 0.8), False))

    def place(model, rot, pvm_base):
        sc.cmd("uniform", 0, colmajor(model))
        sc.cmd("uniform", 1, colmajor(rot))
        sc.cmd("uniform", 5, colmajor((pvm_base.astype(np.float64) @ model.astype(np.float64)).astype(F32)))

    sc.cmd("clearColour", 0xFF000000)
    # ---- shadow map
    sc.cmd("uniform", 2, colmajor(lview)); sc.cmd("uniform", 3, colmajor(lproj)); sc.cmd("uniform", 4, colmajor(lpv))
    sc.cmd("depth", shadow_tex)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("viewport", shadow, shadow)
    sc.cmd("use", p_shadow)
    for (vao, model, rot, _prog, _col, casts) in objects:
        if casts:
            place(model, rot, lpv)
            sc.cmd("draw", vao)
    # ---- the scene
    sc.cmd("uniform", 2, colmajor(view)); sc.cmd("uniform", 3, colmajor(proj)); sc.cmd("uniform", 4, colmajor(pv))
    sc.cmd("uniform", 6, colmajor(lpvb))
    sc.cmd("uniform", 20, vec4(*light)); sc.cmd("uniform", 21, vec4(*ldir)); sc.cmd("uniform", 22, vec4(*camera))
    sc.cmd("tex_uniform", 23, shadow_tex)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("viewport", width, height)
    if skybox:
        sc.cmd("uniform", 1, colmajor(view))     # DEF04 reads the view matrix from slot 1 (skybox.cpp:17-21)
        sc.cmd("tex_uniform", 2, sky)
        sc.cmd("disable", K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH)
        sc.cmd("use", p_sky); sc.cmd("draw", quad, True)
        sc.cmd("enable", K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH)
        sc.cmd("uniform", 2, colmajor(view))     # slot 2 is the view matrix again for the demo-2 functors' callers
    sc.cmd("uniform", 30, vec4(0.15, 0.15, 0.15, 0)); sc.cmd("uniform", 32, vec4(1.0, 1.0, 1.0, 0)); sc.cmd("uniform", 33, np.array([30.0], F32))
    sc.cmd("tex_uniform", 40, wood)
    ntri = 0
    for (vao, model, rot, prog, col, _casts) in objects:
        place(model, rot, pv)
        if col is not None:
            sc.cmd("uniform", 31, vec4(col[2], col[1], col[0], 0))   # BGR order, the functors output b,g,r from [0],[1],[2]
        sc.cmd("use", prog)
        sc.cmd("draw", vao)
        ntri += sc.vaos[vao][0][1].shape[0] // 3
    sc.meta["triangles"] = 2 * ntri + 2
    return sc


# ---- demo 2 from an OBJX file -----------------------------------------------------------

OBJX_PROGRAMMES = {   # mesh_header::programme -> functor triple (src/test2/loadscene.cpp findOrCreateProgramme, procreator.cpp)
    "VP_SingleColour:IP_SingleColour:FP_SingleColour": (K.FN_SINGLECOLOUR, K.FN_SINGLECOLOUR, K.FN_SINGLECOLOUR),
    "VP_DiffuseOnly:IP_DiffuseOnly:FP_DiffuseOnly": (K.FN_DIFFUSEONLY, K.FN_DIFFUSEONLY, K.FN_DIFFUSEONLY),
    "VP_PositionOnly:IP_Null:FP_SingleColourNoLighting": (K.FN_POSITIONONLY, K.FN_POSITIONONLY, K.FN_POSITIONONLY),
}


def mat_view_ypr(eye, ypr):
    """View matrix of a camera at `eye` looking along yaw/pitch (mat4::view(from, ypr), mcemaths.hpp:133; sight vector as in
    src/mcemath/matrxgl.cpp:163-166: (cos p sin y, -sin p, -cos p cos y)). Built once; every backend gets the same bytes."""
    y, p = float(ypr[0]), float(ypr[1])
    sight = np.array([math.cos(p) * math.sin(y), -math.sin(p), -math.cos(p) * math.cos(y)])
    e = np.asarray(eye[:3], dtype=np.float64)
    return mat_look_at(tuple(e), tuple(e + sight))


def write_demo_objx(path, seed=11, clutter=6):
    """A small desk scene in OBJX form — the shape of src/test2/plane.objx (world-space positions, 'object/mesh' names,
    programme strings, '@noshadow' tags, the light1/from and light1/to marker meshes, camera position + yaw/pitch in the
    file header) — so that the OBJX path can be exercised where the reference's own assets are not present."""
    from . import objx
    rng = np.random.default_rng(seed)

    def placed(mesh, at):
        pos, tan, _bin, nrm, uv = mesh
        p = pos.copy()
        p[:, :3] += np.asarray(at, F32)
        p[:, 3] = 0.0       # the loader sets w = 1 (loadscene.cpp:231)
        return {"vertices": p, "normals": nrm, "tangents": tan, "texcoords": uv}

    meshes = []

    def add(name, mesh, at, programme, diffuse=(0.8, 0.8, 0.8, 0), spec_exp=30.0, diffuse_file=""):
        m = placed(mesh, at)
        m.update(name=name, programme=programme, ambient=(0.15, 0.15, 0.15, 0), diffuse=diffuse, specular=(1, 1, 1, 0),
                 specular_exponent=spec_exp, diffuse_file=diffuse_file)
        meshes.append(m)

    single, diffuse_only, unlit = list(OBJX_PROGRAMMES)
    add("plane/top@noshadow", box_mesh(3.0, 0.06, 2.4), (0, -0.03, 0), diffuse_only, diffuse_file="top.jpg")
    for i in range(clutter):
        w, h, d = rng.uniform(0.12, 0.4, size=3)
        x, z = rng.uniform(-1.1, 1.1), rng.uniform(-0.8, 0.8)
        if i % 2:
            add("box%d/body" % i, box_mesh(w, h, d), (x, h / 2, z), single, diffuse=tuple(rng.uniform(0.2, 0.95, size=3)) + (0,))
            add("box%d/lid" % i, box_mesh(w * 0.8, 0.03, d * 0.8), (x, h + 0.015, z), single, diffuse=(0.9, 0.9, 0.2, 0))
        else:
            add("crate%d/body" % i, box_mesh(w, h, d), (x, h / 2, z), diffuse_only, diffuse_file="marmite.jpg")
    add("ball/skin", sphere_mesh(10, 20, 0.22), (-0.35, 0.22, 0.45), single, diffuse=(0.9, 0.3, 0.2, 0))
    add("lamp/lamp@noshadow", box_mesh(0.07, 0.07, 0.07), (0.9, 1.7, 1.1), unlit, diffuse=(1.0, 1.0, 0.8, 0))
    add("light1/from", box_mesh(0.02, 0.02, 0.02), (0.9, 1.7, 1.1), diffuse_only, diffuse_file="top.jpg")
    add("light1/to", box_mesh(0.02, 0.02, 0.02), (0.0, 0.0, 0.0), single)
    scene = {"camera_pos": (-1.6, 1.3, 2.2, 0), "camera_ypr": (0.62, 0.42, 0, 0)}
    objx.write_objx(path, scene, meshes)
    return path


def scene_desk_objx(path, width=1024, height=640, shadow=1024, picture_dir=None, seed=5, tex_size=256, post=False):
    """Demo 2's frame (src/test2/puresoft.cpp:115-248) replayed headless from an OBJX file: loadScene (loadscene.cpp:137-345)
    through puresoft3d_b200.objx, the shadow pass over everything not tagged '@noshadow' with VP_Shadow/IP_Null/FP_Null into a
    float texture, then every component with its own programme, material uniforms 30-33 and texture ids 40-43
    (SceneObject::draw, loadscene.cpp:112-135), in the reference's std::map order. Pictures named by the file are loaded
    from `picture_dir` (load_picture); a name that is not found there gets a seeded synthetic texture instead (the
    reference would be left with texture id -2)."""
    from . import objx
    rng = np.random.default_rng(seed)
    desc, comps = objx.load_scene_meshes(path)
    sc = Scene("demo2-objx-%s-%dx%d-s%d" % (os.path.basename(str(path)), width, height, shadow), width, height)
    shadow_tex = sc.add_texture(shadow, shadow, 4, None)
    p_shadow = sc.add_programme(K.FN_SHADOW2, K.FN_POSITIONONLY, K.FN_SHADOW2)
    progs, texs = {}, {}

    def programme(name):
        if name not in progs:
            progs[name] = sc.add_programme(*OBJX_PROGRAMMES[name])
        return progs[name]

    def texture(name):
        if name not in texs:
            pix = None
            if picture_dir and name and os.path.exists(os.path.join(picture_dir, name)):
                pix = objx.load_picture(os.path.join(picture_dir, name))
            if pix is None:
                pix = tex_smooth_bgra(rng, tex_size, tex_size)
            texs[name] = sc.add_texture(pix.shape[1], pix.shape[0], 4, np.ascontiguousarray(pix))
        return texs[name]

    by_name = {c["component"]: c for c in comps}
    light_from = by_name["light1/from"]["world_translation"].astype(np.float64)
    light_to = by_name["light1/to"]["world_translation"].astype(np.float64)
    rdir = light_from - light_to
    rdir = rdir / np.linalg.norm(rdir)                                           # light1RDir, puresoft.cpp:133-135
    camera = desc["camera_pos"].astype(np.float64)[:3]
    proj = mat_perspective(0.1, 5.0, width / height, 2 * math.pi * (45.0 / 360.0))   # puresoft.cpp:119
    view = mat_view_ypr(camera, desc["camera_ypr"])
    pv = (proj.astype(np.float64) @ view.astype(np.float64)).astype(F32)
    lproj = mat_perspective(0.1, 5.0, 1.0, 2 * math.pi * (90.0 / 360.0))           # puresoft.cpp:146
    lview = mat_look_at(tuple(light_from), tuple(light_to))
    lpv = (lproj.astype(np.float64) @ lview.astype(np.float64)).astype(F32)
    lpvb = (BIAS.astype(np.float64) @ lpv.astype(np.float64)).astype(F32)

    drawn = []
    for c in comps:                      # light1/from and light1/to are marker meshes, erased from the scene (loadscene.cpp:437-439)
        if c["component"] in ("light1/from", "light1/to"):
            continue
        vao = sc.add_vao(objx.mesh_slots(c))
        model = mat_translation(*[float(x) for x in c["world_translation"]])
        tex = texture(c["diffuse_file"]) if OBJX_PROGRAMMES[c["programme"]][0] == K.FN_DIFFUSEONLY else None
        drawn.append((c, vao, model, programme(c["programme"]), tex))

    def place(c, model, pvm_base):
        sc.cmd("uniform", 0, colmajor(model))
        sc.cmd("uniform", 1, colmajor(mat_identity()))
        sc.cmd("uniform", 5, colmajor((pvm_base.astype(np.float64) @ model.astype(np.float64)).astype(F32)))
        sc.cmd("uniform", 30, vec4(*c["ambient"][:3])); sc.cmd("uniform", 31, vec4(*c["diffuse"][:3]))
        sc.cmd("uniform", 32, vec4(*c["specular"][:3])); sc.cmd("uniform", 33, np.array([c["specular_exponent"]], F32))

    # ---- shadow map (puresoft.cpp:185-205)
    sc.cmd("uniform", 2, colmajor(lview)); sc.cmd("uniform", 3, colmajor(lproj)); sc.cmd("uniform", 4, colmajor(lpv))
    sc.cmd("depth", shadow_tex)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("viewport", shadow, shadow)
    sc.cmd("use", p_shadow)
    ntri = 0
    for (c, vao, model, _prog, _tex) in drawn:
        if "@noshadow" in c["component"]:
            continue
        place(c, model, lpv)
        sc.cmd("draw", vao)
        ntri += c["vertices"].shape[0] // 3
    # ---- the scene (puresoft.cpp:211-240)
    sc.cmd("uniform", 2, colmajor(view)); sc.cmd("uniform", 3, colmajor(proj)); sc.cmd("uniform", 4, colmajor(pv))
    sc.cmd("uniform", 6, colmajor(lpvb))
    sc.cmd("uniform", 20, vec4(*light_from)); sc.cmd("uniform", 21, vec4(*rdir)); sc.cmd("uniform", 22, vec4(*camera))
    sc.cmd("tex_uniform", 23, shadow_tex)
    sc.cmd("depth", -1)
    sc.cmd("clearDepth", 1.0)
    sc.cmd("clearColour", 0)
    sc.cmd("viewport", width, height)
    for (c, vao, model, prog, tex) in drawn:
        place(c, model, pv)
        if tex is not None:
            sc.cmd("tex_uniform", 40, tex)
        sc.cmd("use", prog)
        sc.cmd("draw", vao)
        ntri += c["vertices"].shape[0] // 3
    if post:
        sc.cmd("post", K.POST_DEPTHOFFIELD)      # pipeline.postProcess(&depthofField), puresoft.cpp:243 (commented out there)
    sc.meta["triangles"] = ntri
    sc.meta["components"] = [c["component"] for (c, _, _, _, _) in drawn]
    return sc


# ---- edge cases of the inputs themselves -----------------------------------------------------------------------------

def scene_ragged_streams(width=200, height=120, seed=21):
    """What the draw does with awkward vertex streams (vertthrd.cpp:21-31: all attached slots advance in lock-step and the draw
    ends when ANY of them runs out): slots of different lengths, a vertex count that is not a multiple of 3 (the trailing
    vertices never form a triangle), an EMPTY stream, an attached slot the functor does not read that is shorter than the
    ones it reads, and a draw with fewer than 3 vertices — several draws of one frame, DEF02 (position, normal, colour)."""
    rng = np.random.default_rng(seed)
    sc = Scene("ragged-streams-%dx%d" % (width, height), width, height)
    prog = sc.add_programme(K.FN_DEF02)

    def streams(n_pos, n_nrm, n_col, extra=None, z=0.0):
        n = max(n_pos, n_nrm, n_col, 3)
        xy = rng.uniform(-0.9, 0.9, size=(n, 2))
        pos = np.concatenate([xy, np.full((n, 1), z), np.ones((n, 1))], axis=1).astype(F32)
        nrm = np.tile(np.array([0, 0, 1, 0], F32), (n, 1))
        col = np.concatenate([rng.uniform(40, 255, size=(n, 3)), np.full((n, 1), 255.0)], axis=1).astype(F32)
        slots = {0: (16, pos[:n_pos]), 1: (16, nrm[:n_nrm]), 2: (16, col[:n_col])}
        if extra is not None:
            slots[7] = (8, np.zeros((extra, 2), F32))      # a slot DEF02 never reads still ends the draw when it runs out
        return sc.add_vao(slots)

    vaos = [streams(30, 27, 29, z=0.1),        # 27 vertices = 9 triangles
            streams(31, 31, 31, z=0.0),        # 31 vertices = 10 triangles + 1 dangling vertex
            streams(12, 12, 0, z=-0.1),        # empty colour stream: nothing is drawn
            streams(24, 24, 24, extra=7, z=-0.2),   # the unread slot 7 has 7 units: 2 triangles
            streams(2, 2, 2, z=-0.3)]          # fewer than one triangle
    ident = colmajor(mat_identity())
    sc.cmd("viewport", width, height); sc.cmd("depth", -1); sc.cmd("clearDepth", 1.0); sc.cmd("clearColour", 0xFF202020)
    sc.cmd("uniform", 3, ident); sc.cmd("uniform", 4, ident); sc.cmd("uniform", 5, ident)
    sc.cmd("uniform", 7, vec4(0.3, 0.4, 2.0)); sc.cmd("uniform", 8, vec4(0.0, 0.0, 3.0))
    sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    sc.cmd("use", prog)
    for v in vaos:
        sc.cmd("draw", v)
    sc.cmd("enable", K.BEHAVIOR_FACE_CULLING)
    sc.meta["triangles"] = 9 + 10 + 0 + 2 + 0
    return sc


def scene_crowded_tile(width=320, height=200, seed=22, crowd=3000):
    """Extremes of the binning: `crowd` tiny triangles inside ONE 16x16 tile (a tile list longer than the shared-memory sort
    takes: the draw's speculated tail is refused on the device and the exact retry takes the radix path), triangles hundreds
    of times larger than the viewport (the whole-rectangle path of the tile mask), one with a vertex behind the eye
    (w < 0: the perspective divide mirrors it, nothing is clipped) — in submission order that matters
    (depth dead band, fragthrd.cpp:227)."""
    rng = np.random.default_rng(seed)
    sc = Scene("crowded-tile-%dx%d-%d" % (width, height, crowd), width, height)
    pos, col, nrm = [], [], []

    def tri(p, w=(1.0, 1.0, 1.0)):
        for (x, y, z), ww in zip(p, w):
            pos.append((x * ww, y * ww, z * ww, ww))
            col.append(tuple(rng.uniform(30, 255, size=3)) + (255.0,))
            nrm.append((0.0, 0.0, 1.0, 0.0))

    tri([(-60.0, -55.0, 0.9), (70.0, -50.0, 0.9), (3.0, 90.0, 0.9)])                      # covers everything, far
    cx, cy = (40.0 - width // 2) / (width // 2), (72.0 - height // 2) / (height // 2)     # tile (2, 4)
    px, py = 2.0 / width, 2.0 / height
    for i in range(crowd):
        c = np.array([cx, cy]) + rng.uniform(-6, 6, size=2) * (px, py)
        d = rng.uniform(-2.5, 2.5, size=(3, 2)) * (px, py)
        z = 0.5 - 0.0003 * (i % 1000) + rng.uniform(-0.00005, 0.00005, size=3)                # inside and around the 1e-4 dead band
        tri([(c[0] + d[k, 0], c[1] + d[k, 1], z[k]) for k in range(3)])
    tri([(-0.5, -40.0, 0.3), (0.5, -40.0, 0.3), (0.0, 45.0, 0.3)])                         # a tall spike through every tile row
    tri([(-0.8, -0.6, 0.2), (0.8, -0.6, 0.2), (0.0, 0.7, 0.2)], w=(1.0, 1.0, -0.5))        # one vertex behind the eye
    vao = sc.add_vao({0: (16, np.array(pos, F32)), 1: (16, np.array(nrm, F32)), 2: (16, np.array(col, F32))})
    prog = sc.add_programme(K.FN_DEF02)
    ident = colmajor(mat_identity())
    sc.cmd("viewport", width, height); sc.cmd("depth", -1); sc.cmd("clearDepth", 1.0); sc.cmd("clearColour", 0xFF101010)
    sc.cmd("uniform", 3, ident); sc.cmd("uniform", 4, ident); sc.cmd("uniform", 5, ident)
    sc.cmd("uniform", 7, vec4(0.3, 0.4, 2.0)); sc.cmd("uniform", 8, vec4(0.0, 0.0, 3.0))
    sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    sc.cmd("use", prog)
    sc.cmd("draw", vao)
    sc.cmd("enable", K.BEHAVIOR_FACE_CULLING)
    sc.meta["triangles"] = len(pos) // 3
    return sc


def scene_small_draws(width=320, height=200, seed=31, draws=24, tris=40, crowd=0, big_every=0, flatid=True):
    """Many small draws in a row into the same targets — what a frame of scene objects looks like (src/test2/puresoft.cpp:185-248:
    54 draws of boxes) — cycling through three programmes with 3, 1 and 0 varyings (DEF02, FLATID, PositionOnly + single colour),
    every draw with its own vertex streams and its own latched uniforms (a translation, a colour). The triangles of different
    draws overlap with depths inside and around the 1e-4 dead band (fragthrd.cpp:227), so the order of the draws decides pixels.
    `crowd` > 0: that many extra tiny triangles per draw inside one 16x16 tile (tile lists that outgrow one speculated capacity
    after the other). `big_every` > 0: every such draw is one of 20000 triangles (too big to wait for its neighbours).
    `flatid` False: without the parity-test functor FLATID the reference build does not have (DEF02 takes its turns)."""
    rng = np.random.default_rng(seed)
    sc = Scene("small-draws-%dx%d-%d-%d-%d" % (width, height, draws, tris, crowd), width, height)
    progs = [sc.add_programme(K.FN_DEF02), sc.add_programme(K.FN_FLATID if flatid else K.FN_DEF02), sc.add_programme(K.FN_POSITIONONLY)]
    ident = colmajor(mat_identity())
    px, py = 2.0 / width, 2.0 / height
    cx, cy = (40.0 - width // 2) / (width // 2), (72.0 - height // 2) / (height // 2)
    sc.cmd("viewport", width, height); sc.cmd("depth", -1); sc.cmd("clearDepth", 1.0); sc.cmd("clearColour", 0xFF202020)
    sc.cmd("uniform", 3, ident); sc.cmd("uniform", 5, ident)
    sc.cmd("uniform", 7, vec4(0.3, 0.4, 2.0)); sc.cmd("uniform", 8, vec4(0.0, 0.0, 3.0))
    sc.cmd("disable", K.BEHAVIOR_FACE_CULLING)
    total = 0
    for d in range(draws):
        n = 20000 if (big_every and d % big_every == big_every - 1) else tris + int(rng.integers(0, 5))
        c = rng.uniform(-0.9, 0.9, size=(n, 1, 2))
        ext = rng.uniform(-0.25, 0.25, size=(n, 3, 2)) if n < 1000 else rng.uniform(-0.02, 0.02, size=(n, 3, 2))
        xy = c + ext
        z = (0.4 + 0.0002 * rng.integers(0, 6, size=(n, 1)) + rng.uniform(-0.00005, 0.00005, size=(n, 3)))[..., None]
        pos = np.concatenate([xy, z, np.ones((n, 3, 1))], axis=-1)
        if crowd:
            cc = np.array([cx, cy]) + rng.uniform(-6, 6, size=(crowd, 1, 2)) * (px, py)
            ce = rng.uniform(-2.5, 2.5, size=(crowd, 3, 2)) * (px, py)
            cz = (0.3 - 0.0001 * (np.arange(crowd) % 7))[:, None, None] + rng.uniform(-0.00005, 0.00005, size=(crowd, 3, 1))
            pos = np.concatenate([pos, np.concatenate([cc + ce, cz, np.ones((crowd, 3, 1))], axis=-1)], axis=0)
            n += crowd
        pos = pos.reshape(n * 3, 4).astype(F32)
        nrm = np.tile(np.array([0.0, 0.0, 1.0, 0.0], F32), (n * 3, 1))
        col = np.concatenate([rng.uniform(30, 255, size=(n * 3, 3)), np.full((n * 3, 1), 255.0)], axis=1).astype(F32)
        ids = np.repeat(np.concatenate([rng.integers(0, 256, size=(n, 3)), np.full((n, 1), 255)], axis=1), 3, axis=0).astype(F32)
        vao = sc.add_vao({0: (16, pos), 1: (16, nrm), 2: (16, col), 6: (16, ids)})
        shift = mat_translation(float(rng.uniform(-0.05, 0.05)), float(rng.uniform(-0.05, 0.05)), 0.0)
        sc.cmd("uniform", 4, colmajor(shift))           # DEF02 / FLATID: model; PositionOnly reads slot 5 (left alone)
        sc.cmd("uniform", 31, vec4(*rng.uniform(0.1, 1.0, size=3)))
        sc.cmd("use", progs[d % 3])
        sc.cmd("draw", vao)
        total += n
    sc.cmd("enable", K.BEHAVIOR_FACE_CULLING)
    sc.meta["triangles"] = total
    return sc


def scene_state_fuzz(seed, width=None, height=None, draws=None):
    """A seeded frame of several draws with the pipeline's state changing between them: behaviour bits (depth test / depth
    write / culling / blending; never 'test on, write off', where the reference reads a depth cursor it did not position,
    SURVEY.md §9.12), viewport sizes up to the target's, DEF01 / DEF02 / DEF03 / DEF04 programmes, CLAMP and WRAP textures,
    triangle soups of all sizes with perspective w. Identity view, so positions are clip space."""
    rng = np.random.default_rng(1000 + seed)
    width = int(width or rng.integers(40, 200))
    if width % 4 == 1:
        width += 1          # pipeline.cpp:31 under-allocates the default depth rows when W % 4 == 1: not a width to pin on
    height = int(height or rng.integers(30, 150))
    sc = Scene("state-fuzz-%d-%dx%d" % (seed, width, height), width, height)
    tsz = int(rng.choice([8, 32, 64]))
    tex2d = sc.add_texture(tsz, tsz, 4, tex_random_bgra(rng, tsz, tsz), wrap=int(rng.choice([K.WRAP_CLAMP, K.WRAP_WRAP])))
    bump = sc.add_texture(tsz, tsz, 4, tex_smooth_bgra(rng, tsz, tsz, normal_map=True))
    cube = sc.add_texture(16, 16, 4, layers=[tex_random_bgra(rng, 16, 16) for _ in range(6)])
    progs = {f: sc.add_programme(f) for f in (K.FN_DEF01, K.FN_DEF02, K.FN_DEF03, K.FN_DEF04)}
    ident = colmajor(mat_identity())
    sc.cmd("depth", -1); sc.cmd("clearDepth", 1.0); sc.cmd("clearColour", int(rng.integers(0, 2 ** 32)))
    for u in (3, 4, 5):
        sc.cmd("uniform", u, ident)
    sc.cmd("uniform", 7, vec4(*rng.uniform(-2, 2, size=3))); sc.cmd("uniform", 8, vec4(0.0, 0.0, 3.0))
    sc.cmd("uniform", 1, ident)                                            # DEF04: view matrix in slot 1, cube map id in slot 2
    sc.cmd("tex_uniform", 2, cube); sc.cmd("tex_uniform", 9, tex2d); sc.cmd("tex_uniform", 10, bump)
    all_bits = K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH | K.BEHAVIOR_FACE_CULLING | K.BEHAVIOR_ALPHABLEND
    ntri = 0
    for d in range(int(draws or rng.integers(2, 6))):
        n = int(rng.integers(1, 60))
        c = rng.uniform(-1.2, 1.2, size=(n, 1, 2))
        size = rng.choice([0.02, 0.2, 0.8, 2.5], size=(n, 1, 1))
        xy = c + rng.uniform(-1, 1, size=(n, 3, 2)) * size
        z = rng.uniform(-1.05, 1.05, size=(n, 3, 1))
        w = np.where(rng.random((n, 1, 1)) < 0.3, rng.uniform(0.4, 3.0, size=(n, 3, 1)), 1.0)
        pos = (np.concatenate([xy, z, np.ones((n, 3, 1))], axis=2) * w).reshape(n * 3, 4).astype(F32)
        nrm = np.tile(np.array([0, 0, 1, 0], F32), (n * 3, 1))
        tan = np.tile(np.array([1, 0, 0, 0], F32), (n * 3, 1))
        bin_ = np.tile(np.array([0, 1, 0, 0], F32), (n * 3, 1))
        col = np.concatenate([rng.uniform(20, 255, size=(n * 3, 3)), np.full((n * 3, 1), 255.0)], axis=1).astype(F32)
        uv = rng.uniform(-0.2, 1.6, size=(n * 3, 2)).astype(F32)
        uv = np.abs(uv)                                                    # (unsigned)negative differs between x86 and CUDA; the scenes keep uv >= 0
        fn = int(rng.choice([K.FN_DEF01, K.FN_DEF02, K.FN_DEF03, K.FN_DEF04]))
        if fn == K.FN_DEF02:
            slots = {0: (16, pos), 1: (16, nrm), 2: (16, col)}
        elif fn == K.FN_DEF04:
            slots = {0: (16, pos)}
        else:
            slots = {0: (16, pos), 1: (16, tan), 2: (16, bin_), 3: (16, nrm), 4: (8, uv)}
        vao = sc.add_vao(slots)
        bits = int(rng.integers(0, 16))
        if (bits & K.BEHAVIOR_TEST_DEPTH) and not (bits & K.BEHAVIOR_UPDATE_DEPTH):
            bits |= K.BEHAVIOR_UPDATE_DEPTH
        sc.cmd("disable", all_bits); sc.cmd("enable", bits)
        sc.cmd("viewport", int(rng.integers(max(2, width // 3), width + 1)), int(rng.integers(max(2, height // 3), height + 1)))
        sc.cmd("use", progs[fn])
        sc.cmd("draw", vao)
        ntri += n
    sc.cmd("disable", all_bits)
    sc.cmd("enable", K.BEHAVIOR_UPDATE_DEPTH | K.BEHAVIOR_TEST_DEPTH | K.BEHAVIOR_FACE_CULLING)
    sc.meta["triangles"] = ntri
    return sc
