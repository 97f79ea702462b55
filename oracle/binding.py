"""ctypes binding of the CPU oracle (oracle/libpt_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTC_ORACLE_LIB") or os.path.join(_HERE, "libpt_oracle.so")   # another build of pt_oracle.cpp: the sanitizer build of tools/fuzz.sh
_lib = None


def _types():
    return importlib.import_module("wgpu-pathtracer_b200._types")


def build():
    subprocess.check_call(["make", "-C", _HERE, "libpt_oracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    T = _types()
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
    P = C.POINTER
    L.orc_scene_create.restype, L.orc_scene_create.argtypes = vp, [P(T.SceneView), P(T.Config)]
    L.orc_scene_destroy.restype, L.orc_scene_destroy.argtypes = None, [vp]
    L.orc_num_lights.restype, L.orc_num_lights.argtypes = u32, [vp]
    L.orc_set_size.restype, L.orc_set_size.argtypes = None, [vp, u32, u32]
    L.orc_render_frame.restype = i32
    L.orc_render_frame.argtypes = [vp, P(T.Uniforms), vp, vp, vp, vp, vp, i32, i32, u32, u32, P(u64)]
    L.orc_trace_rays.restype, L.orc_trace_rays.argtypes = i32, [vp, vp, u64, i32, vp, vp, P(u64)]
    L.orc_pixel.restype, L.orc_pixel.argtypes = i32, [vp, P(T.Uniforms), u32, u32, i32, P(C.c_float), P(u32), P(u64)]
    L.orc_trace_rays_any.restype, L.orc_trace_rays_any.argtypes = i32, [vp, vp, u64, vp]
    L.orc_jenkins_hash.restype, L.orc_jenkins_hash.argtypes = u32, [u32]
    L.orc_rand_stream.restype, L.orc_rand_stream.argtypes = None, [u32, u32, u32, u32, u32, P(u32), vp, vp]
    L.orc_rand_float_of.restype, L.orc_rand_float_of.argtypes = C.c_float, [u32]
    L.orc_camera.restype = None
    L.orc_camera.argtypes = [u32, u32, C.c_float, P(C.c_float), C.c_float, C.c_float, i32, P(T.Uniforms)]
    L.orc_math.restype, L.orc_math.argtypes = None, [i32, vp, vp, vp, u64]
    L.orc_sample_texture.restype, L.orc_sample_texture.argtypes = None, [vp, u32, vp, u64, vp]
    L.orc_sky.restype, L.orc_sky.argtypes = None, [vp, vp, u64, vp]
    L.orc_bvh_info.restype, L.orc_bvh_info.argtypes = None, [vp, P(u64)]
    L.orc_rayproject.restype, L.orc_rayproject.argtypes = None, [P(T.Uniforms), P(T.Uniforms), vp, vp]
    L.orc_blit.restype, L.orc_blit.argtypes = None, [vp, u64, vp, vp]
    _lib = L
    return L


class OracleScene:
    """The oracle's copy of a scene + its own BVH2."""

    def __init__(self, scene, config):
        self.config = config
        view = scene.view()
        self._h = lib().orc_scene_create(C.byref(view), C.byref(config))
        if not self._h:
            raise RuntimeError("orc_scene_create failed")

    @property
    def num_lights(self):
        return lib().orc_num_lights(self._h)

    def render_frame(self, uniforms, color=None, brute=False, nthreads=0, rows=None, want_aux=True):
        """One frame of pathtracer.wgsl `main`.  color: (H, W, 4) float32 accumulation, updated in place (created
        zeroed if None).  Returns dict(color, depth, normal, albedo, hit_id, rays, nodes, tris)."""
        w, h = self.config.width, self.config.height
        if color is None:
            color = np.zeros((h, w, 4), dtype=np.float32)
        depth = np.zeros((h, w), dtype=np.float32) if want_aux else None
        normal = np.zeros((h, w, 4), dtype=np.float32) if want_aux else None
        albedo = np.zeros((h, w, 4), dtype=np.float32) if want_aux else None
        hit = np.full((h, w, 2), 0xFFFFFFFF, dtype=np.uint32) if want_aux else None
        cnt = (C.c_uint64 * 3)()
        y0, y1 = rows if rows is not None else (0, h)
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = lib().orc_render_frame(self._h, C.byref(uniforms), p(color), p(depth), p(normal), p(albedo), p(hit),
                                    1 if brute else 0, nthreads, y0, y1, cnt)
        assert rc == 0
        return {"color": color, "depth": depth, "normal": normal, "albedo": albedo, "hit_id": hit,
                "rays": cnt[0], "nodes": cnt[1], "tris": cnt[2]}

    def pixel(self, uniforms, x, y, brute=False):
        """pixel_color of one pixel: (rgba, RNG state after its path(s), rays traced)"""
        col, st, rays = (C.c_float * 4)(), C.c_uint32(), C.c_uint64()
        assert lib().orc_pixel(self._h, C.byref(uniforms), int(x), int(y), 1 if brute else 0, col, C.byref(st), C.byref(rays)) == 0
        return np.array(list(col), dtype=np.float32), st.value, rays.value

    def trace_rays(self, rays, brute=False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = len(rays)
        hits = np.zeros((n, 4), dtype=np.uint32)
        ids = np.zeros((n, 2), dtype=np.uint32)
        cnt = (C.c_uint64 * 3)()
        lib().orc_trace_rays(self._h, rays.ctypes.data_as(C.c_void_p), n, 1 if brute else 0,
                             hits.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p), cnt)
        return hits, ids, {"rays": cnt[0], "nodes": cnt[1], "tris": cnt[2]}

    def trace_rays_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(len(rays), dtype=np.uint8)
        lib().orc_trace_rays_any(self._h, rays.ctypes.data_as(C.c_void_p), len(rays), occ.ctypes.data_as(C.c_void_p))
        return occ

    def sample_texture(self, idx, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float32)
        lib().orc_sample_texture(self._h, idx, uv.ctypes.data_as(C.c_void_p), len(uv), out.ctypes.data_as(C.c_void_p))
        return out

    def sky(self, dirs):
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(dirs), 3), dtype=np.float32)
        lib().orc_sky(self._h, dirs.ctypes.data_as(C.c_void_p), len(dirs), out.ctypes.data_as(C.c_void_p))
        return out

    def bvh_info(self):
        a = (C.c_uint64 * 4)()
        lib().orc_bvh_info(self._h, a)
        return {"nodes": a[0], "triangles": a[1], "inner": a[2]}

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rand_stream(x, y, width, frame, n=4):
    state = C.c_uint32()
    ints = np.zeros(n, dtype=np.uint32)
    floats = np.zeros(n, dtype=np.float32)
    lib().orc_rand_stream(x, y, width, frame, n, C.byref(state), ints.ctypes.data_as(C.c_void_p), floats.ctypes.data_as(C.c_void_p))
    return state.value, ints, floats


def camera_uniforms(width, height, fov, pos, theta_x=0.0, theta_y=0.0, extra_updates=1):
    T = _types()
    u = T.Uniforms()
    p = (C.c_float * 3)(*[float(v) for v in pos])
    lib().orc_camera(width, height, fov, p, theta_x, theta_y, extra_updates, C.byref(u))
    return u


def rayproject(cur, prev, depth):
    """rayproject.wgsl: target pixel index of every source pixel (0xFFFFFFFF = store dropped)."""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    out = np.zeros(depth.size, dtype=np.uint32)
    lib().orc_rayproject(C.byref(cur), C.byref(prev), depth.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def blit(rgba):
    """blit.wgsl + sRGB surface: (..., 4) float32 -> (..., 4) uint8, plus the tonemapped linear rgb."""
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    n = rgba.size // 4
    out = np.zeros(rgba.shape, dtype=np.uint8)
    tm = np.zeros(rgba.shape[:-1] + (3,), dtype=np.float32)
    lib().orc_blit(rgba.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), tm.ctypes.data_as(C.c_void_p))
    return out, tm


def math_fn(fn, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), dtype=np.float32)
    out = np.zeros_like(a)
    code = fn if isinstance(fn, int) else {"sin": 0, "cos": 1, "atan": 2, "acos": 3, "atan2": 4}[fn]
    lib().orc_math(code, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), a.size)
    return out
