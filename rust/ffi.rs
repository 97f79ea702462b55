//! ffi.rs — the `extern "C"` surface of libptcore (include/ptcore.h) for the reference crate: EVERY `PTC_API` symbol of the
//! header, in the header's order (tests/test_rust_boundary.py diffs the two lists).
//! Text only: this image has no Rust toolchain, so this file has not been compiled here.  Layouts are the
//! crate's own `#[repr(C)]` bytemuck structs (src/uniforms.rs), passed through unchanged.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)] pub struct ptc_step { _private: [u8; 0] }
#[repr(C)] pub struct ptc_scene { _private: [u8; 0] }

#[repr(C)] #[derive(Copy, Clone)]
pub struct ptc_mesh { pub vertex_offset: u32, pub vertex_count: u32, pub index_offset: u32, pub index_count: u32, pub material_index: u32 }

#[repr(C)] #[derive(Copy, Clone)]
pub struct ptc_texture_desc { pub width: u32, pub height: u32, pub format: u32, pub _pad: u32, pub texels: *const c_void }

#[repr(C)]
pub struct ptc_scene_view {
    pub vertices: *const crate::uniforms::Vertex,     pub num_vertices: u64,
    pub indices: *const u32,                          pub num_indices: u64,
    pub meshes: *const ptc_mesh,                      pub num_meshes: u64,
    pub materials: *const crate::uniforms::Material,  pub num_materials: u64,
    pub textures: *const ptc_texture_desc,            pub num_textures: u64,
}

#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct ptc_config {
    pub width: u32, pub height: u32, pub max_bounces: u32, pub samples_per_pixel: u32,
    pub send_to_light_probability: f32, pub env_texture_index: u32, pub accum_mode: u32, pub compat_flags: u32,
    pub bvh_width: u32, pub light_sampling: u32, pub bvh_build: u32, pub ploc_window: u32, pub reserved: [u32; 4],
}

#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct ptc_stats {
    pub rays_traced: u64, pub frames_rendered: u64, pub kernel_launches: u64, pub frame_parts: u64,
    pub packet_node_visits: u64, pub packet_lane_visits: u64, pub reserved: [u64; 2],
}

#[repr(C)] #[derive(Copy, Clone)]
pub struct ptc_host_camera {
    pub width: u32, pub height: u32, pub fov: f32, pub position: [f32; 3], pub lookat: [f32; 3], pub frame: u32,
    pub samples_per_pixel: u32, pub max_bounces: u32,
    pub pixel_space_x: [f32; 3], pub pixel_space_y: [f32; 3], pub first_pixel_pos: [f32; 3],
    pub theta_x: f32, pub theta_y: f32, pub camera_speed: f32,
}

#[repr(C)]
pub struct ptc_scene_texture_info { pub source: u32, pub width: u32, pub height: u32, pub format: u32, pub color: [f32; 4], pub path: *const c_char }

pub const PTC_PLANE_COLOR: c_int = 0;
pub const PTC_PLANE_DEPTH: c_int = 1;
pub const PTC_PLANE_NORMAL: c_int = 2;
pub const PTC_PLANE_ALBEDO: c_int = 3;
pub const PTC_PLANE_HIT_ID: c_int = 4;
pub const PTC_PLANE_DISPLAY: c_int = 5;
pub const PTC_PLANE_REPROJECTED: c_int = 6;
pub const PTC_PLANE_REDUCED: c_int = 7;
pub const PTC_TEXEL_RGBA8: u32 = 0;
pub const PTC_TEXEL_RGBA32F: u32 = 1;
pub const PTC_ACCUM_MEAN: u32 = 0;
pub const PTC_ACCUM_SUM: u32 = 1;
pub const PTC_ERR_NOT_READY: c_int = -5;

type Uniforms = crate::uniforms::Uniforms;

#[link(name = "ptcore")]
extern "C" {
    // RenderStep lifecycle
    pub fn ptc_create(scene: *const ptc_scene_view, config: *const ptc_config, cuda_device: c_int, out: *mut *mut ptc_step) -> c_int;
    pub fn ptc_update(step: *mut ptc_step, uniforms: *const Uniforms) -> c_int;
    pub fn ptc_render(step: *mut ptc_step) -> c_int;
    pub fn ptc_poll(step: *mut ptc_step, wait: c_int) -> c_int;
    pub fn ptc_resize(step: *mut ptc_step, width: u32, height: u32) -> c_int;
    pub fn ptc_set_output(step: *mut ptc_step, device_rgba32f: *mut c_void) -> c_int;
    pub fn ptc_reset_accumulation(step: *mut ptc_step) -> c_int;
    pub fn ptc_blit(step: *mut ptc_step, src_device_rgba32f: *const c_void) -> c_int;                                     // BlitStep::render
    pub fn ptc_rayproject(step: *mut ptc_step, prev: *const Uniforms, cur: *const Uniforms, src_device_rgba32f: *const c_void) -> c_int;   // RayprojectStep::render
    pub fn ptc_destroy(step: *mut ptc_step);
    // results
    pub fn ptc_read(step: *mut ptc_step, plane: c_int, dst: *mut c_void, dst_pitch_bytes: usize) -> c_int;
    pub fn ptc_write(step: *mut ptc_step, plane: c_int, src: *const c_void, src_pitch_bytes: usize) -> c_int;          // checkpoint / resume
    pub fn ptc_read_async(step: *mut ptc_step, plane: c_int, dst_pinned: *mut c_void) -> c_int;
    pub fn ptc_read_wait(step: *mut ptc_step) -> c_int;
    pub fn ptc_copy_plane(step: *mut ptc_step, plane: c_int, dst_device: *mut c_void, dst_pitch_bytes: usize) -> c_int;  // denoise.rs copy_texture_to_buffer
    pub fn ptc_device_ptr(step: *mut ptc_step, plane: c_int) -> *mut c_void;
    pub fn ptc_stream(step: *mut ptc_step) -> *mut c_void;
    // multi-GPU
    pub fn ptc_reduce(step: *mut ptc_step, nccl_comm: *mut c_void, root: c_int) -> c_int;
    pub fn ptc_set_reduce_output(step: *mut ptc_step, device_rgba32f: *mut c_void) -> c_int;
    // statistics, timing, facts
    pub fn ptc_get_stats(step: *mut ptc_step, out: *mut ptc_stats) -> c_int;
    pub fn ptc_set_profiling(step: *mut ptc_step, enable: c_int) -> c_int;
    pub fn ptc_get_profile(step: *mut ptc_step, ms: *mut c_double, launches: *mut u64) -> c_int;                        // [3] each
    pub fn ptc_num_lights(step: *mut ptc_step, out: *mut u32) -> c_int;
    pub fn ptc_bvh_info(step: *mut ptc_step, out: *mut u64) -> c_int;                                                   // [8]
    pub fn ptc_debug_bvh_hash(step: *mut ptc_step, out: *mut u64) -> c_int;                                             // [2]
    // tooling / unit-test entry points
    pub fn ptc_trace_rays(step: *mut ptc_step, rays: *const c_float, n: u64, hits: *mut u32, ids: *mut u32) -> c_int;
    pub fn ptc_trace_rays_any(step: *mut ptc_step, rays: *const c_float, n: u64, occluded: *mut u8) -> c_int;
    pub fn ptc_debug_math(cuda_device: c_int, func: c_int, a: *const c_float, b: *const c_float, out: *mut c_float, n: u64) -> c_int;
    pub fn ptc_debug_sample(step: *mut ptc_step, texture_index: c_int, input: *const c_float, n: u64, rgba: *mut c_float, per_lane: c_int) -> c_int;
    pub fn ptc_debug_unorm8(cuda_device: c_int, out512: *mut c_float) -> c_int;
    pub fn ptc_debug_l2_bandwidth(cuda_device: c_int, bytes: u64, iters: u32, gbs: *mut c_double) -> c_int;
    pub fn ptc_last_error() -> *const c_char;
    pub fn ptc_version() -> *const c_char;
    pub fn ptc_device_count() -> c_int;
    // host side without the crate's own camera.rs / scene.rs (a Rust host normally keeps its own; listed for completeness)
    pub fn ptc_camera_new(cam: *mut ptc_host_camera, width: u32, height: u32, fov_deg: c_float, position: *const c_float);
    pub fn ptc_camera_update(cam: *mut ptc_host_camera);
    pub fn ptc_camera_uniforms(cam: *const ptc_host_camera, out: *mut Uniforms);
    pub fn ptc_scene_new(envmap_path: *const c_char, out: *mut *mut ptc_scene) -> c_int;
    pub fn ptc_scene_set_fix_normals(scene: *mut ptc_scene, enable: c_int) -> c_int;
    pub fn ptc_scene_load_obj(scene: *mut ptc_scene, obj_path: *const c_char) -> c_int;
    pub fn ptc_scene_load_obj_text(scene: *mut ptc_scene, obj_text: *const c_char, mtl_text: *const c_char, obj_dir: *const c_char) -> c_int;
    pub fn ptc_scene_num_textures(scene: *const ptc_scene) -> u64;
    pub fn ptc_scene_texture_info_at(scene: *const ptc_scene, index: u64, out: *mut ptc_scene_texture_info) -> c_int;
    pub fn ptc_scene_set_texels(scene: *mut ptc_scene, index: u64, width: u32, height: u32, format: u32, texels: *const c_void) -> c_int;
    pub fn ptc_scene_decode_textures(scene: *mut ptc_scene, cuda_device: c_int, n_threads: c_int) -> c_int;
    pub fn ptc_decode_image(data: *const c_void, bytes: usize, cuda_device: c_int, width: *mut u32, height: *mut u32, format: *mut u32, texels: *mut *mut c_void) -> c_int;
    pub fn ptc_free(p: *mut c_void);
    pub fn ptc_scene_get_view(scene: *const ptc_scene, out: *mut ptc_scene_view) -> c_int;
    pub fn ptc_scene_free(scene: *mut ptc_scene);
}

/// `.expect(..)`-style helper: the reference's steps panic on failure (unwrap/expect everywhere, e.g.
/// src/wgpu_util.rs:107-109); the shim keeps that behaviour at the boundary.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ptc_last_error()) }.to_string_lossy().into_owned();
        panic!("ptcore error {}: {}", rc, msg);
    }
}
