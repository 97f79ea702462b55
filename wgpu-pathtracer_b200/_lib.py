"""ctypes loader for csrc/libptcore.so (the C ABI of include/ptcore.h).

There is no Python or CPU implementation behind this module: if the shared library has not been
built (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C wgpu-pathtracer_b200/csrc``)
loading fails loudly.
"""
import ctypes as C
import os

from . import _types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTC_LIB: development override used to A/B-time kernel variants built by tools/build_variant.sh (still a CUDA build of
# the same sources; there is no non-CUDA library to point it at)
LIB_PATH = os.environ.get("PTC_LIB") or os.path.join(_HERE, "csrc", "libptcore.so")
_lib = None


class PtcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"ptcore error {status}: {message}")
        self.status = status


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first (make -C wgpu-pathtracer_b200/csrc). "
                           "There is no CPU fallback for the path-tracing core.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t
    P = C.POINTER
    sig = {
        "ptc_create": (i32, [P(T.SceneView), P(T.Config), i32, P(vp)]),
        "ptc_update": (i32, [vp, P(T.Uniforms)]),
        "ptc_render": (i32, [vp]),
        "ptc_poll": (i32, [vp, i32]),
        "ptc_resize": (i32, [vp, u32, u32]),
        "ptc_set_output": (i32, [vp, vp]),
        "ptc_reset_accumulation": (i32, [vp]),
        "ptc_blit": (i32, [vp, vp]),
        "ptc_rayproject": (i32, [vp, P(T.Uniforms), P(T.Uniforms), vp]),
        "ptc_destroy": (None, [vp]),
        "ptc_read": (i32, [vp, i32, vp, sz]),
        "ptc_read_async": (i32, [vp, i32, vp]),
        "ptc_read_wait": (i32, [vp]),
        "ptc_copy_plane": (i32, [vp, i32, vp, sz]),
        "ptc_device_ptr": (vp, [vp, i32]),
        "ptc_stream": (vp, [vp]),
        "ptc_get_stats": (i32, [vp, P(T.Stats)]),
        "ptc_set_profiling": (i32, [vp, i32]),
        "ptc_get_profile": (i32, [vp, P(C.c_double), P(u64)]),
        "ptc_num_lights": (i32, [vp, P(u32)]),
        "ptc_bvh_info": (i32, [vp, P(u64)]),
        "ptc_reduce": (i32, [vp, vp, i32]),
        "ptc_set_reduce_output": (i32, [vp, vp]),
        "ptc_write": (i32, [vp, i32, vp, C.c_size_t]),
        "ptc_debug_bvh_hash": (i32, [vp, P(u64)]),
        "ptc_trace_rays": (i32, [vp, vp, u64, vp, vp]),
        "ptc_trace_rays_any": (i32, [vp, vp, u64, vp]),
        "ptc_debug_math": (i32, [i32, i32, vp, vp, vp, u64]),
        "ptc_debug_unorm8": (i32, [i32, vp]),
        "ptc_debug_l2_bandwidth": (i32, [i32, u64, C.c_uint32, P(C.c_double)]),
        "ptc_debug_sample": (i32, [vp, i32, vp, u64, vp, i32]),
        "ptc_last_error": (C.c_char_p, []),
        "ptc_version": (C.c_char_p, []),
        "ptc_device_count": (i32, []),
        "ptc_camera_new": (None, [P(T.HostCamera), u32, u32, C.c_float, P(C.c_float)]),
        "ptc_camera_update": (None, [P(T.HostCamera)]),
        "ptc_camera_uniforms": (None, [P(T.HostCamera), P(T.Uniforms)]),
        "ptc_scene_new": (i32, [C.c_char_p, P(vp)]),
        "ptc_scene_set_fix_normals": (i32, [vp, i32]),
        "ptc_scene_load_obj": (i32, [vp, C.c_char_p]),
        "ptc_scene_load_obj_text": (i32, [vp, C.c_char_p, C.c_char_p, C.c_char_p]),
        "ptc_scene_num_textures": (u64, [vp]),
        "ptc_scene_texture_info_at": (i32, [vp, u64, P(T.SceneTextureInfo)]),
        "ptc_scene_set_texels": (i32, [vp, u64, u32, u32, u32, vp]),
        "ptc_scene_decode_textures": (i32, [vp, i32, i32]),
        "ptc_decode_image": (i32, [vp, C.c_size_t, i32, P(u32), P(u32), P(u32), P(vp)]),
        "ptc_free": (None, [vp]),
        "ptc_scene_get_view": (i32, [vp, P(T.SceneView)]),
        "ptc_scene_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by tests from include/ptcore.h


def check(status):
    if status != 0:
        raise PtcError(status, lib().ptc_last_error().decode("utf-8", "replace"))
