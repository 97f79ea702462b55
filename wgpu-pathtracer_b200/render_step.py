"""``RTStep`` — the path-tracing render step, mirroring ``impl RenderStep for RTStep`` of the reference
(src/render_steps/mod.rs:20-25, src/render_steps/pt.rs:341-565,732-745) on top of the C ABI.

Every method is a thin call into libptcore.so (CUDA); there is no Python compute here.
"""
import ctypes as C

import numpy as np

from . import _types as T
from ._lib import PtcError, check, lib


def make_config(width, height, max_bounces=4, samples_per_pixel=1, send_to_light_probability=0.0,
                env_texture_index=9, accum_mode=T.ACCUM_MEAN, light_sampling=0, bvh_build=0, ploc_window=0):
    """The reference's compile-time constants (src/shaders/shader_definitions.rs:3-26) as a run-time struct."""
    c = T.Config()
    c.width, c.height, c.max_bounces, c.samples_per_pixel = width, height, max_bounces, samples_per_pixel
    c.send_to_light_probability, c.env_texture_index, c.accum_mode = send_to_light_probability, env_texture_index, accum_mode
    c.light_sampling = light_sampling          # ptc_light_sampling: 0 = reference behaviour, 1 = shadow rays + MIS (non-parity)
    c.bvh_build = bvh_build                    # ptc_bvh_build: 0 = binned SAH on the host, 1 = linear BVH, 2 = PLOC, both built on the GPU
    c.ploc_window = ploc_window                # PLOC only: merge-partner search window (0 = default)
    return c


class RTStep:
    def __init__(self, handle, config):
        self._h = handle
        self.config = config

    # RenderStep::create
    @classmethod
    def create(cls, scene, config, device=0):
        h = C.c_void_p()
        view = scene.view()
        check(lib().ptc_create(C.byref(view), C.byref(config), int(device), C.byref(h)))
        return cls(h, config)

    # RenderStep::update
    def update(self, uniforms):
        check(lib().ptc_update(self._h, C.byref(uniforms)))

    # RenderStep::render (+ queue.submit): asynchronous
    def render(self):
        check(lib().ptc_render(self._h))

    def poll(self, wait=True):
        """``device.poll(Wait)`` / ``device.poll(Poll)``: returns True when idle."""
        rc = lib().ptc_poll(self._h, 1 if wait else 0)
        if rc == T.ERR_NOT_READY:
            return False
        check(rc)
        return True

    # RenderStep::resize
    def resize(self, width, height):
        check(lib().ptc_resize(self._h, int(width), int(height)))
        self.config.width, self.config.height = int(width), int(height)

    def set_output(self, device_ptr):
        check(lib().ptc_set_output(self._h, C.c_void_p(device_ptr) if device_ptr else None))

    def write(self, image, plane=T.PLANE_COLOR):
        """Load an accumulation image previously returned by read(PLANE_COLOR): checkpoint / resume (ptc_write)."""
        image = np.ascontiguousarray(image, dtype=np.float32)
        assert image.shape == (self.config.height, self.config.width, 4)
        check(lib().ptc_write(self._h, int(plane), image.ctypes.data_as(C.c_void_p), 0))

    def reduce(self, nccl_comm, root=0):
        """Sum the per-rank SUM-mode accumulation planes into the REDUCED plane of `root` with one ncclReduce on the step's
        stream (ptc_reduce); the accumulation itself is untouched, so the call can be repeated after more frames.
        nccl_comm: the ncclComm_t as an integer (e.g. torch's ProcessGroupNCCL._comm_ptr()); root < 0 = all-reduce."""
        check(lib().ptc_reduce(self._h, C.c_void_p(int(nccl_comm)), int(root)))

    def set_reduce_output(self, device_ptr):
        check(lib().ptc_set_reduce_output(self._h, C.c_void_p(device_ptr) if device_ptr else None))

    def reset_accumulation(self):
        check(lib().ptc_reset_accumulation(self._h))

    def read(self, plane=T.PLANE_COLOR):
        w, h = self.config.width, self.config.height
        shape, dt = {T.PLANE_COLOR: ((h, w, 4), np.float32), T.PLANE_DEPTH: ((h, w), np.float32),
                     T.PLANE_NORMAL: ((h, w, 4), np.float32), T.PLANE_ALBEDO: ((h, w, 4), np.float32),
                     T.PLANE_HIT_ID: ((h, w, 2), np.uint32), T.PLANE_DISPLAY: ((h, w, 4), np.uint8),
                     T.PLANE_REPROJECTED: ((h, w, 4), np.float32), T.PLANE_REDUCED: ((h, w, 4), np.float32)}[plane]
        out = np.empty(shape, dtype=dt)
        check(lib().ptc_read(self._h, plane, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def blit(self, src_device_ptr=None):
        """BlitStep::render: tonemap + sRGB-encode the COLOR plane (or `src`) into the 8-bit DISPLAY plane."""
        check(lib().ptc_blit(self._h, C.c_void_p(src_device_ptr) if src_device_ptr else None))

    def rayproject(self, prev_uniforms, cur_uniforms, src_device_ptr=None):
        """RayprojectStep::render: reproject the last real frame from `prev` to `cur` into the REPROJECTED plane."""
        check(lib().ptc_rayproject(self._h, C.byref(prev_uniforms), C.byref(cur_uniforms), C.c_void_p(src_device_ptr) if src_device_ptr else None))

    def read_async(self, plane, pinned_ptr):
        check(lib().ptc_read_async(self._h, plane, C.c_void_p(pinned_ptr)))

    def copy_plane(self, plane, dst_device_ptr, pitch=0):
        """device-to-device copy of a plane into a caller buffer with `pitch` bytes per row (ptc_copy_plane)"""
        check(lib().ptc_copy_plane(self._h, int(plane), C.c_void_p(dst_device_ptr), int(pitch)))

    def read_wait(self):
        check(lib().ptc_read_wait(self._h))

    def device_ptr(self, plane=T.PLANE_COLOR):
        return lib().ptc_device_ptr(self._h, plane)

    @property
    def stream(self):
        return lib().ptc_stream(self._h)

    def stats(self):
        s = T.Stats()
        check(lib().ptc_get_stats(self._h, C.byref(s)))
        return {"rays_traced": s.rays_traced, "frames_rendered": s.frames_rendered, "kernel_launches": s.kernel_launches,
                "frame_parts": s.frame_parts, "packet_node_visits": s.packet_node_visits, "packet_lane_visits": s.packet_lane_visits}

    def set_profiling(self, enable=True):
        check(lib().ptc_set_profiling(self._h, 1 if enable else 0))

    def profile(self):
        """Per-stage device time (CUDA events on the render stream) since set_profiling(True)."""
        ms, n = (C.c_double * 3)(), (C.c_uint64 * 3)()
        check(lib().ptc_get_profile(self._h, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(("raygen", "trace", "shade"))}

    @property
    def num_lights(self):
        n = C.c_uint32()
        check(lib().ptc_num_lights(self._h, C.byref(n)))
        return n.value

    def bvh_info(self):
        a = (C.c_uint64 * 8)()
        check(lib().ptc_bvh_info(self._h, a))
        return {"nodes": a[0], "triangles": a[1], "node_bytes": a[2], "triangle_bytes": a[3], "max_depth": a[4], "build_ms": a[5],
                "built_on_device": bool(a[6]), "build_us": a[7]}

    def bvh_hash(self):
        a = (C.c_uint64 * 2)()
        check(lib().ptc_debug_bvh_hash(self._h, a))
        return a[0], a[1]

    def trace_rays(self, rays):
        """rays: (n, 8) float32 [ox,oy,oz,_,dx,dy,dz,_] -> (t, b1, b2, tri, ids) through the traversal kernel."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = len(rays)
        hits = np.zeros((n, 4), dtype=np.uint32)
        ids = np.zeros((n, 2), dtype=np.uint32)
        check(lib().ptc_trace_rays(self._h, rays.ctypes.data_as(C.c_void_p), n, hits.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p)))
        return hits, ids

    def trace_rays_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(len(rays), dtype=np.uint8)
        check(lib().ptc_trace_rays_any(self._h, rays.ctypes.data_as(C.c_void_p), len(rays), occ.ctypes.data_as(C.c_void_p)))
        return occ

    def sample_texture(self, index, uv, per_lane=False):
        """the kernels' bilinear lookup of texture `index` at (n, 2) uv pairs -> (n, 4) rgba (unit-test entry);
        per_lane: lookup i uses texture (index + i) % num_textures"""
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float32)
        check(lib().ptc_debug_sample(self._h, int(index), uv.ctypes.data_as(C.c_void_p), len(uv), out.ctypes.data_as(C.c_void_p), int(per_lane)))
        return out

    def sky(self, dirs):
        """the kernels' sky_color for (n, 3) directions -> (n, 3) rgb (unit-test entry)"""
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((len(dirs), 4), dtype=np.float32)
        check(lib().ptc_debug_sample(self._h, -1, dirs.ctypes.data_as(C.c_void_p), len(dirs), out.ctypes.data_as(C.c_void_p), 0))
        return out[:, :3].copy()

    def destroy(self):
        if self._h:
            lib().ptc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
