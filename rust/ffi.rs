//! ffi.rs — the `extern "C"` surface of libptcore (include/ptcore.h) for the reference crate.
//! Text only: this image has no Rust toolchain, so this file has not been compiled here.  Layouts are the
//! crate's own `#[repr(C)]` bytemuck structs (src/uniforms.rs), passed through unchanged.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct ptc_step { _private: [u8; 0] }

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
    pub bvh_width: u32, pub light_sampling: u32, pub bvh_build: u32, pub reserved: [u32; 5],
}

pub const PTC_PLANE_COLOR: c_int = 0;
pub const PTC_PLANE_DEPTH: c_int = 1;
pub const PTC_PLANE_NORMAL: c_int = 2;
pub const PTC_PLANE_ALBEDO: c_int = 3;
pub const PTC_TEXEL_RGBA8: u32 = 0;
pub const PTC_TEXEL_RGBA32F: u32 = 1;

#[link(name = "ptcore")]
extern "C" {
    pub fn ptc_create(scene: *const ptc_scene_view, config: *const ptc_config, cuda_device: c_int, out: *mut *mut ptc_step) -> c_int;
    pub fn ptc_update(step: *mut ptc_step, uniforms: *const crate::uniforms::Uniforms) -> c_int;
    pub fn ptc_render(step: *mut ptc_step) -> c_int;
    pub fn ptc_poll(step: *mut ptc_step, wait: c_int) -> c_int;
    pub fn ptc_resize(step: *mut ptc_step, width: u32, height: u32) -> c_int;
    pub fn ptc_set_output(step: *mut ptc_step, device_rgba32f: *mut c_void) -> c_int;
    pub fn ptc_reset_accumulation(step: *mut ptc_step) -> c_int;
    pub fn ptc_read(step: *mut ptc_step, plane: c_int, dst: *mut c_void, dst_pitch_bytes: usize) -> c_int;
    pub fn ptc_read_async(step: *mut ptc_step, plane: c_int, dst_pinned: *mut c_void) -> c_int;
    pub fn ptc_read_wait(step: *mut ptc_step) -> c_int;
    pub fn ptc_device_ptr(step: *mut ptc_step, plane: c_int) -> *mut c_void;
    pub fn ptc_write(step: *mut ptc_step, plane: c_int, src: *const c_void, src_pitch_bytes: usize) -> c_int;          // checkpoint / resume
    pub fn ptc_blit(step: *mut ptc_step, src_device_rgba32f: *const c_void) -> c_int;                                     // BlitStep::render
    pub fn ptc_rayproject(step: *mut ptc_step, prev: *const crate::uniforms::Uniforms, cur: *const crate::uniforms::Uniforms,
                          src_device_rgba32f: *const c_void) -> c_int;                                                  // RayprojectStep::render
    pub fn ptc_reduce(step: *mut ptc_step, nccl_comm: *mut c_void, root: c_int) -> c_int;                                // multi-GPU image sum
    pub fn ptc_stream(step: *mut ptc_step) -> *mut c_void;
    pub fn ptc_destroy(step: *mut ptc_step);
    pub fn ptc_last_error() -> *const c_char;
}

/// `.expect(..)`-style helper: the reference's steps panic on failure (unwrap/expect everywhere, e.g.
/// src/wgpu_util.rs:107-109); the shim keeps that behaviour at the boundary.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ptc_last_error()) }.to_string_lossy().into_owned();
        panic!("ptcore error {}: {}", rc, msg);
    }
}
