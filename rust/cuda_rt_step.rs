//! cuda_rt_step.rs — `impl RenderStep` over libptcore: the drop-in for src/render_steps/pt.rs (text only).
//!
//! Needs ONE addition to the reference crate, because `Texture.rgba` and `Texture.path` are private (src/texture.rs:14-24)
//! and `load()` consumes the decoded image into a BC7 / rgba32f wgpu texture (texture.rs:232-336): an accessor for the
//! pre-BC7 texels.  Add to `impl Texture` in src/texture.rs (INTEGRATION.md section 1 repeats it):
//!
//! ```rust
//! /// Decoded texels before BC7 (what `load()` starts from, texture.rs:237-243), for the CUDA path.
//! pub fn rgba32f(&mut self) -> (u32, u32, Vec<f32>) {
//!     if self.rgba.is_none() {
//!         let img = image::open(self.path.as_ref().unwrap()).unwrap();
//!         self.size = img.dimensions();
//!         self.rgba = Some(img.to_rgba32f());
//!     }
//!     (self.size.0, self.size.1, self.rgba.as_ref().unwrap().as_raw().clone())
//! }
//! ```
use crate::{ffi, scene, uniforms, wgpu_util};
use super::RenderStep;

pub struct CudaRTStep { step: *mut ffi::ptc_step }

impl RenderStep for CudaRTStep {
    // replaces RTStep::create (pt.rs:341-490): hand the Scene arrays over; the library copies them and builds its own BVH
    fn create(state: &mut wgpu_util::WGPUState, scene: &scene::Scene) -> Self {
        // scene::Mesh (scene.rs:54-63) and ffi::ptc_mesh have the same five u32 fields; Scene.materials already holds
        // uniforms::Material (scene.rs:149-166), Scene.vertices uniforms::Vertex, Scene.indices uniforms::Index (one u32 each)
        let meshes: Vec<ffi::ptc_mesh> = scene.meshes.iter().map(|m| ffi::ptc_mesh {
            vertex_offset: m.vertex_offset, vertex_count: m.vertex_count, index_offset: m.index_offset,
            index_count: m.index_count, material_index: m.material_index }).collect();
        let texels: Vec<(u32, u32, Vec<f32>)> = scene.textures.iter().map(|t| t.lock().unwrap().rgba32f()).collect();
        let descs: Vec<ffi::ptc_texture_desc> = texels.iter().map(|(w, h, px)| ffi::ptc_texture_desc {
            width: *w, height: *h, format: ffi::PTC_TEXEL_RGBA32F, _pad: 0, texels: px.as_ptr() as *const _ }).collect();
        let view = ffi::ptc_scene_view {
            vertices: scene.vertices.as_ptr(), num_vertices: scene.vertices.len() as u64,
            indices: scene.indices.as_ptr() as *const u32, num_indices: scene.indices.len() as u64,
            meshes: meshes.as_ptr(), num_meshes: meshes.len() as u64,
            materials: scene.materials.as_ptr(), num_materials: scene.materials.len() as u64,
            textures: descs.as_ptr(), num_textures: descs.len() as u64,
        };
        let cfg = ffi::ptc_config { width: state.config.width, height: state.config.height,       // pt.rs:471-472 sizes its buffers from state.config
            max_bounces: crate::shaders::shader_definitions::MAX_BOUNCES,
            samples_per_pixel: crate::shaders::shader_definitions::SAMPLES_PER_PIXEL,
            send_to_light_probability: crate::shaders::shader_definitions::SEND_TO_LIGHT_PROBABILITY,
            env_texture_index: 9, ..Default::default() };
        let mut step = std::ptr::null_mut();
        ffi::check(unsafe { ffi::ptc_create(&view, &cfg, 0, &mut step) });                         // copies everything: texels / meshes may drop now
        Self { step }
    }
    // replaces RTStep::update (pt.rs:492-521)
    fn update(&mut self, state: &mut wgpu_util::WGPUState, scene: &scene::Scene) {
        let mut u = uniforms::Uniforms::new(&scene.camera);
        u.is_grabbed = state.window_cursor_grabbed as u32;
        ffi::check(unsafe { ffi::ptc_update(self.step, &u) });
    }
    // replaces RTStep::render (pt.rs:523-565); the encoder is unused: work goes to the step's CUDA stream
    fn render(&mut self, _state: &mut wgpu_util::WGPUState, _scene: &scene::Scene, _encoder: &mut wgpu::CommandEncoder, _out: Option<&wgpu::SurfaceTexture>) {
        ffi::check(unsafe { ffi::ptc_render(self.step) });
    }
    // replaces RTStep::resize (pt.rs:732-745) + the accumulation texture re-creation (raytracer.rs:344)
    fn resize(&mut self, state: &mut wgpu_util::WGPUState, _scene: &scene::Scene) {
        ffi::check(unsafe { ffi::ptc_resize(self.step, state.config.width, state.config.height) });
    }
}

impl CudaRTStep {
    /// `rt_device.poll(Poll)` → "previous RT frame still running" (raytracer.rs:165-167)
    pub fn busy(&self) -> bool { unsafe { ffi::ptc_poll(self.step, 0) == ffi::PTC_ERR_NOT_READY } }
}

impl Drop for CudaRTStep { fn drop(&mut self) { unsafe { ffi::ptc_destroy(self.step) } } }
