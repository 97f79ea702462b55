"""ctypes / numpy mirrors of include/ptcore_types.h (reference layouts: src/uniforms.rs:3-215, src/scene.rs:54-63)."""
import ctypes as C

import numpy as np

u32, f32, u64 = C.c_uint32, C.c_float, C.c_uint64


class Camera(C.Structure):  # 96 bytes, uniforms.rs:3-27
    _fields_ = [("position", f32 * 3), ("_pad0", u32), ("lookat", f32 * 3), ("frame", u32),
                ("pixel_space_x", f32 * 3), ("_pad1", u32), ("pixel_space_y", f32 * 3), ("_pad2", u32),
                ("first_pixel_pos", f32 * 3), ("width", u32), ("height", u32), ("fov", f32),
                ("samples_per_pixel", u32), ("max_bounces", u32)]


class Uniforms(C.Structure):  # 112 bytes, uniforms.rs:53-61
    _fields_ = [("camera", Camera), ("num_lights", u32), ("is_grabbed", u32), ("_pad", u32 * 2)]


class TextureDesc(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("format", u32), ("_pad", u32), ("texels", C.c_void_p)]


class SceneView(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("num_vertices", u64), ("indices", C.c_void_p), ("num_indices", u64),
                ("meshes", C.c_void_p), ("num_meshes", u64), ("materials", C.c_void_p), ("num_materials", u64),
                ("textures", C.POINTER(TextureDesc)), ("num_textures", u64)]


class Config(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("max_bounces", u32), ("samples_per_pixel", u32),
                ("send_to_light_probability", f32), ("env_texture_index", u32), ("accum_mode", u32),
                ("compat_flags", u32), ("bvh_width", u32), ("light_sampling", u32), ("bvh_build", u32), ("ploc_window", u32), ("reserved", u32 * 4)]


class Stats(C.Structure):
    _fields_ = [("rays_traced", u64), ("frames_rendered", u64), ("kernel_launches", u64), ("frame_parts", u64), ("packet_node_visits", u64), ("packet_lane_visits", u64), ("reserved", u64 * 2)]


class HostCamera(C.Structure):  # include/ptcore.h ptc_host_camera (camera.rs:5-28)
    _fields_ = [("width", u32), ("height", u32), ("fov", f32), ("position", f32 * 3), ("lookat", f32 * 3),
                ("frame", u32), ("samples_per_pixel", u32), ("max_bounces", u32),
                ("pixel_space_x", f32 * 3), ("pixel_space_y", f32 * 3), ("first_pixel_pos", f32 * 3),
                ("theta_x", f32), ("theta_y", f32), ("camera_speed", f32)]


class SceneTextureInfo(C.Structure):
    _fields_ = [("source", u32), ("width", u32), ("height", u32), ("format", u32), ("color", f32 * 4), ("path", C.c_char_p)]


assert C.sizeof(Camera) == 96 and C.sizeof(Uniforms) == 112 and C.sizeof(Config) == 64

VERTEX_DT = np.dtype([("position", "<f4", 3), ("_pad", "<u4"), ("normal", "<f4", 3), ("_pad2", "<u4"),
                      ("uv", "<f4", 2), ("_pad3", "<u4", 2)])                      # 48 bytes, uniforms.rs:138-147
MESH_DT = np.dtype([("vertex_offset", "<u4"), ("vertex_count", "<u4"), ("index_offset", "<u4"),
                    ("index_count", "<u4"), ("material_index", "<u4")])            # scene.rs:54-63
MATERIAL_DT = np.dtype([("albedo_texture_idx", "<u4"), ("roughness_texture_idx", "<u4"), ("metallic_texture_idx", "<u4"),
                        ("emissive_texture_idx", "<u4"), ("specular_texture_idx", "<u4"), ("sheen_texture_idx", "<u4"),
                        ("clearcoat_thickness", "<f4"), ("clearcoat_roughness", "<f4"), ("anisotropy", "<f4"),
                        ("anisotropy_rotation", "<f4"), ("normal_texture_idx", "<u4"), ("transmission_texture_idx", "<u4"),
                        ("ior_texture_idx", "<u4")])                               # 52 bytes, uniforms.rs:92-113
assert VERTEX_DT.itemsize == 48 and MESH_DT.itemsize == 20 and MATERIAL_DT.itemsize == 52

TEXEL_RGBA8, TEXEL_RGBA32F = 0, 1
ACCUM_MEAN, ACCUM_SUM = 0, 1
LIGHTS_REFERENCE, LIGHTS_MIS = 0, 1
BVH_SAH_HOST, BVH_LBVH_DEVICE, BVH_PLOC_DEVICE = 0, 1, 2
PLANE_COLOR, PLANE_DEPTH, PLANE_NORMAL, PLANE_ALBEDO, PLANE_HIT_ID, PLANE_DISPLAY, PLANE_REPROJECTED, PLANE_REDUCED = 0, 1, 2, 3, 4, 5, 6, 7
TEX_CONSTANT, TEX_FILE, TEX_TEXELS = 0, 1, 2
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_NOT_READY, ERR_IO, ERR_PARSE = 0, -1, -2, -3, -4, -5, -6, -7
