"""ctypes prototypes for include/ps3d.h.

`bind(path)` loads ONE shared library that implements the header and declares every entry point. The product
library is puresoft3d_b200/libps3d_b200.so (hand-written sm_100a CUDA); `load_product()` raises if it has not
been built — there is no CPU fallback on the product path. The test suite binds the same prototypes to the
oracle libraries (see tests/conftest.py); nothing in this package knows where those live.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "libps3d_b200.so")

OK = 0
ERR_OUT_OF_RANGE = -1
ERR_INVALID_ARGUMENT = -2
ERR_BAD_ALLOC = -3
ERR_DEVICE = -4
ERR_UNSUPPORTED = -5

BEHAVIOR_UPDATE_DEPTH = 0x1
BEHAVIOR_TEST_DEPTH = 0x2
BEHAVIOR_FACE_CULLING = 0x4
BEHAVIOR_ALPHABLEND = 0x8

WRAP_CLAMP, WRAP_WRAP = 0, 1
PROC_VERTEX, PROC_INTERPOLATION, PROC_FRAGMENT = 0, 1, 2

FN_DEF01, FN_DEF02, FN_DEF03, FN_DEF04, FN_DEF05 = 1, 2, 3, 4, 5
FN_PLANET, FN_SATELLITE, FN_CLOUD, FN_CLOUDSHADOW = 16, 17, 18, 19
FN_POSITIONONLY, FN_SINGLECOLOUR, FN_DIFFUSEONLY, FN_SHADOW2 = 32, 33, 34, 35
POST_DEPTHOFFIELD = 1
FN_FLATID = 64
FN_TEXPROBE = 65


class Stats(C.Structure):
    _fields_ = [
        ("triangles_submitted", C.c_uint64),
        ("triangles_rasterised", C.c_uint64),
        ("spans", C.c_uint64),
        ("fragments_tested", C.c_uint64),
        ("fragments_shaded", C.c_uint64),
        ("draws", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Profile(C.Structure):
    _fields_ = [("geom_ms", C.c_double), ("bin_ms", C.c_double), ("tile_ms", C.c_double), ("shade_ms", C.c_double),
                ("geom_launches", C.c_uint64), ("bin_launches", C.c_uint64), ("tile_launches", C.c_uint64),
                ("shade_launches", C.c_uint64), ("bin_pairs", C.c_uint64), ("survivors", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_INT_P = C.POINTER(C.c_int)

# name -> (restype, argtypes); the single source of truth the symbol-export test walks
PROTOTYPES = {
    "ps3d_backend_name": (C.c_char_p, []),
    "ps3d_last_error": (C.c_char_p, [_P]),
    "ps3d_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "ps3d_destroy": (C.c_int, [_P]),
    "ps3d_texture_create": (C.c_int, [_P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.c_int, _INT_P]),
    "ps3d_texture_upload": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ps3d_texture_download": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ps3d_texture_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_texture_set_filter": (C.c_int, [_P, C.c_int, C.c_int]),
    "ps3d_vbo_create": (C.c_int, [_P, C.c_size_t, C.c_size_t, _INT_P]),
    "ps3d_vbo_update": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ps3d_vbo_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_vao_create": (C.c_int, [_P, _INT_P]),
    "ps3d_vao_attach": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _INT_P]),
    "ps3d_vao_detach": (C.c_int, [_P, C.c_int, C.c_int, _INT_P]),
    "ps3d_vao_get": (C.c_int, [_P, C.c_int, C.c_int, _INT_P]),
    "ps3d_vao_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_processor_add": (C.c_int, [_P, C.c_int, C.c_int, _INT_P]),
    "ps3d_processor_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_programme_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _INT_P]),
    "ps3d_programme_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_programme_use": (C.c_int, [_P, C.c_int]),
    "ps3d_set_viewport": (C.c_int, [_P, C.c_int, C.c_int]),
    "ps3d_set_depth": (C.c_int, [_P, C.c_int]),
    "ps3d_set_uniform": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_size_t]),
    "ps3d_enable": (C.c_int, [_P, C.c_int]),
    "ps3d_disable": (C.c_int, [_P, C.c_int]),
    "ps3d_clear_depth": (C.c_int, [_P, C.c_float]),
    "ps3d_clear_colour": (C.c_int, [_P, C.c_uint32]),
    "ps3d_draw_vao": (C.c_int, [_P, C.c_int, C.c_int]),
    "ps3d_finish": (C.c_int, [_P]),
    "ps3d_post_process": (C.c_int, [_P, C.c_int]),
    "ps3d_swap_buffers": (C.c_int, [_P]),
    "ps3d_read_colour": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "ps3d_read_depth": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "ps3d_write_colour": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "ps3d_write_depth": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "ps3d_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "ps3d_reset_stats": (C.c_int, [_P]),
    "ps3d_debug_capture": (C.c_int, [_P, C.c_int, C.c_int]),
    "ps3d_debug_read_shade_counts": (C.c_int, [_P, C.c_void_p]),
    "ps3d_debug_clear_shade_counts": (C.c_int, [_P]),
    "ps3d_set_row_band": (C.c_int, [_P, C.c_int, C.c_int]),
    "ps3d_device_colour_ptr": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ps3d_device_depth_ptr": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ps3d_device_stream": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "ps3d_vbo_update_device": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ps3d_vbo_update_async": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]),
    "ps3d_vbo_device_ptr": (C.c_int, [_P, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ps3d_vbo_device_written": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "ps3d_device_copy_stream": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "ps3d_read_colour_async": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "ps3d_device_join": (C.c_int, [_P]),
    "ps3d_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ps3d_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ps3d_comm_destroy": (C.c_int, [_P]),
    "ps3d_composite_bands": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "ps3d_vbo_all_gather": (C.c_int, [_P, C.c_int]),
    "ps3d_peer_export": (C.c_int, [_P, C.c_void_p]),
    "ps3d_peer_import": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "ps3d_composite_peer": (C.c_int, [_P]),
    "ps3d_graph_begin": (C.c_int, [_P]),
    "ps3d_graph_end": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "ps3d_graph_launch": (C.c_int, [_P, C.c_int]),
    "ps3d_graph_destroy": (C.c_int, [_P, C.c_int]),
    "ps3d_device_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "ps3d_debug_batch_counts": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ps3d_profile_enable": (C.c_int, [_P, C.c_int]),
    "ps3d_profile_read": (C.c_int, [_P, C.POINTER(Profile)]),
    "ps3d_host_approx_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}


def bind(path):
    """Load `path` and attach restype/argtypes for every function of include/ps3d.h (missing symbol -> AttributeError)."""
    lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._ps3d_path = path
    return lib


_product = None


def load_product():
    """The CUDA library. Raises (loudly) when it has not been built — no fallback exists."""
    global _product
    if _product is None:
        if not os.path.exists(PRODUCT_LIB):
            raise RuntimeError(
                "puresoft3d_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback on the product path." % PRODUCT_LIB)
        _product = bind(PRODUCT_LIB)
    return _product
