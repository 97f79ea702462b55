//! cuda_rt_step.rs — `impl RenderStep` over libptcore: the drop-in for src/render_steps/pt.rs (text only).
use crate::{ffi, scene, uniforms, wgpu_util};
use super::RenderStep;

pub struct CudaRTStep { step: *mut ffi::ptc_step, texels: Vec<Vec<u8>> }

impl RenderStep for CudaRTStep {
    // replaces RTStep::create (pt.rs:341-490): hand the Scene arrays over; the library builds its own BVH
    fn create(state: &mut wgpu_util::WGPUState, scene: &scene::Scene) -> Self {
        let materials: Vec<uniforms::Material> = scene.materials.iter().map(|m| uniforms::Material::new(m.clone())).collect();
        let meshes: Vec<ffi::ptc_mesh> = scene.meshes.iter().map(|m| ffi::ptc_mesh {
            vertex_offset: m.vertex_offset, vertex_count: m.vertex_count, index_offset: m.index_offset,
            index_count: m.index_count, material_index: m.material_index }).collect();
        // texture.rs keeps the decoded `rgba` (ImageBuffer<Rgba<f32>>) until load(); pass it as RGBA32F (pre-BC7 texels)
        let texels: Vec<Vec<u8>> = scene.textures.iter().map(|t| t.lock().unwrap().rgba32f_bytes()).collect();
        let descs: Vec<ffi::ptc_texture_desc> = scene.textures.iter().zip(&texels).map(|(t, px)| {
            let t = t.lock().unwrap();
            ffi::ptc_texture_desc { width: t.size.0, height: t.size.1, format: ffi::PTC_TEXEL_RGBA32F, _pad: 0, texels: px.as_ptr() as *const _ }
        }).collect();
        let view = ffi::ptc_scene_view {
            vertices: scene.vertices.as_ptr(), num_vertices: scene.vertices.len() as u64,
            indices: scene.indices.as_ptr() as *const u32, num_indices: scene.indices.len() as u64,
            meshes: meshes.as_ptr(), num_meshes: meshes.len() as u64,
            materials: materials.as_ptr(), num_materials: materials.len() as u64,
            textures: descs.as_ptr(), num_textures: descs.len() as u64,
        };
        let cfg = ffi::ptc_config { width: state.size.0, height: state.size.1,
            max_bounces: crate::shaders::shader_definitions::MAX_BOUNCES,
            samples_per_pixel: crate::shaders::shader_definitions::SAMPLES_PER_PIXEL,
            send_to_light_probability: crate::shaders::shader_definitions::SEND_TO_LIGHT_PROBABILITY,
            env_texture_index: 9, ..Default::default() };
        let mut step = std::ptr::null_mut();
        ffi::check(unsafe { ffi::ptc_create(&view, &cfg, 0, &mut step) });
        Self { step, texels }
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
        ffi::check(unsafe { ffi::ptc_resize(self.step, state.size.0, state.size.1) });
    }
}

impl Drop for CudaRTStep { fn drop(&mut self) { unsafe { ffi::ptc_destroy(self.step) } } }
