import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ptc():
    """The product package (directory name carries a hyphen, hence importlib)."""
    return importlib.import_module("wgpu-pathtracer_b200")


@pytest.fixture(scope="session")
def standins(ptc):
    return importlib.import_module("wgpu-pathtracer_b200.standins")


@pytest.fixture(scope="session")
def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding
    binding.lib()
    return binding


class HostEmu:
    """ctypes binding of tests/hostemu/libpt_hostemu.so — the product's device functions compiled for the host."""

    def __init__(self, ptc):
        here = os.path.join(ROOT, "tests", "hostemu")
        so = os.environ.get("PTC_HOSTEMU_LIB")          # another build of the same sources, e.g. the sanitizer build of `tools/fuzz.sh hostemu`
        if not so:
            so = os.path.join(here, "libpt_hostemu.so")
            subprocess.check_call(["make", "-C", here, "libpt_hostemu.so"], stdout=subprocess.DEVNULL)
        T = ptc.types
        L = C.CDLL(so)
        vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
        L.emu_create.restype, L.emu_create.argtypes = vp, [C.POINTER(T.SceneView), C.POINTER(T.Config)]
        L.emu_destroy.restype, L.emu_destroy.argtypes = None, [vp]
        L.emu_last_error.restype = C.c_char_p
        L.emu_render_frame.restype = i32
        L.emu_render_frame.argtypes = [vp, C.POINTER(T.Uniforms), vp, vp, vp, vp, vp, i32, C.POINTER(u64)]
        L.emu_trace_rays.restype, L.emu_trace_rays.argtypes = i32, [vp, vp, u64, vp, vp]
        L.emu_trace_rays_any.restype, L.emu_trace_rays_any.argtypes = i32, [vp, vp, u64, vp]
        L.emu_bvh_info.restype, L.emu_bvh_info.argtypes = None, [vp, C.POINTER(u64)]
        L.emu_bvh_hash.restype, L.emu_bvh_hash.argtypes = None, [vp, C.POINTER(u64)]
        L.emu_rayproject.restype, L.emu_rayproject.argtypes = None, [C.POINTER(T.Uniforms), C.POINTER(T.Uniforms), vp, vp]
        L.emu_blit.restype, L.emu_blit.argtypes = None, [vp, u64, vp]
        self.L = L

    def rayproject(self, cur, prev, depth):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        out = np.zeros(depth.size, dtype=np.uint32)
        self.L.emu_rayproject(C.byref(cur), C.byref(prev), depth.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def blit(self, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        out = np.zeros(rgba.shape, dtype=np.uint8)
        self.L.emu_blit(rgba.ctypes.data_as(C.c_void_p), rgba.size // 4, out.ctypes.data_as(C.c_void_p))
        return out

    def scene(self, scene, config):
        return HostEmuScene(self.L, scene, config)


class HostEmuScene:
    def __init__(self, L, scene, config):
        self.L, self.config = L, config
        view = scene.view()
        self._h = L.emu_create(C.byref(view), C.byref(config))
        if not self._h:
            raise RuntimeError("emu_create failed: " + L.emu_last_error().decode())

    def render_frame(self, uniforms, color=None, nthreads=0):
        w, h = self.config.width, self.config.height
        if color is None:
            color = np.zeros((h, w, 4), dtype=np.float32)
        depth = np.zeros((h, w), dtype=np.float32)
        normal = np.zeros((h, w, 4), dtype=np.float32)
        albedo = np.zeros((h, w, 4), dtype=np.float32)
        hit = np.full((h, w, 2), 0xFFFFFFFF, dtype=np.uint32)
        cnt = (C.c_uint64 * 3)()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert self.L.emu_render_frame(self._h, C.byref(uniforms), p(color), p(depth), p(normal), p(albedo), p(hit), nthreads, cnt) == 0
        return {"color": color, "depth": depth, "normal": normal, "albedo": albedo, "hit_id": hit, "rays": cnt[0], "nodes": cnt[1], "tris": cnt[2]}

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 4), dtype=np.uint32)
        ids = np.zeros((len(rays), 2), dtype=np.uint32)
        self.L.emu_trace_rays(self._h, rays.ctypes.data_as(C.c_void_p), len(rays), hits.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
        return hits, ids

    def trace_rays_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(len(rays), dtype=np.uint8)
        self.L.emu_trace_rays_any(self._h, rays.ctypes.data_as(C.c_void_p), len(rays), occ.ctypes.data_as(C.c_void_p))
        return occ

    def bvh_info(self):
        a = (C.c_uint64 * 4)()
        self.L.emu_bvh_info(self._h, a)
        return {"nodes": a[0], "triangles": a[1], "max_depth": a[2], "build_ms": a[3]}

    def bvh_hash(self):
        a = (C.c_uint64 * 2)()
        self.L.emu_bvh_hash(self._h, a)
        return a[0], a[1]

    def __del__(self):
        try:
            if self._h:
                self.L.emu_destroy(self._h)
                self._h = None
        except Exception:
            pass


@pytest.fixture(scope="session")
def hostemu(ptc):
    return HostEmu(ptc)


def random_rays(n, seed, extent=4.0):
    rng = np.random.RandomState(seed)
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0:3] = (rng.rand(n, 3) * 2 - 1) * extent * 1.5
    d = rng.randn(n, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    return rays
