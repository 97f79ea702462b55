// ptcore.hpp — header-only C++ face of the C ABI (include/ptcore.h), shaped like the reference's own host types so that
// compiled host code reads like the Rust it replaces:
//   ptcore::Camera      <- src/camera.rs:5-79        (Camera::new / update, Uniforms::new)
//   ptcore::Scene       <- src/scene.rs:65-128       (Scene::new / load_obj; textures decoded by the library or handed over as texels)
//   ptcore::RenderStep  <- src/render_steps/mod.rs:20-25 (trait RenderStep: create / update / render / resize)
//   ptcore::RTStep      <- src/render_steps/pt.rs:12-25, 341-565, 732-745
// Errors: the reference panics (unwrap / expect); here every failed call throws ptcore::Error carrying ptc_last_error().
#pragma once
#include "ptcore.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace ptcore {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& what) : std::runtime_error("ptcore error " + std::to_string(st) + ": " + what), status(st) {}
};
inline void check(int rc) { if (rc != PTC_OK) throw Error(rc, ptc_last_error()); }

class Camera {                                        // camera.rs:5-28
public:
    ptc_host_camera c{};
    Camera(uint32_t width, uint32_t height, float fov_deg, const float position[3]) { ptc_camera_new(&c, width, height, fov_deg, position); }
    void update() { ptc_camera_update(&c); }                                               // camera.rs:54-79
    ptc_uniforms uniforms(uint32_t frame, bool is_grabbed = false) {                      // uniforms.rs:30-50,64-71; raytracer.rs:179-180
        c.frame = frame;
        ptc_uniforms u; ptc_camera_uniforms(&c, &u);
        u.is_grabbed = is_grabbed ? 1u : 0u;
        return u;
    }
};

class Scene {                                         // scene.rs:65-74
    ptc_scene* s_ = nullptr;
public:
    explicit Scene(const std::string& envmap_path) { check(ptc_scene_new(envmap_path.c_str(), &s_)); }   // Scene::new, scene.rs:77-122
    ~Scene() { ptc_scene_free(s_); }
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    void load_obj(const std::string& path) { check(ptc_scene_load_obj(s_, path.c_str())); }              // scene.rs:124-257
    void load_obj_text(const std::string& obj, const std::string& mtl, const std::string& dir) { check(ptc_scene_load_obj_text(s_, obj.c_str(), mtl.c_str(), dir.c_str())); }
    uint64_t num_textures() const { return ptc_scene_num_textures(s_); }
    ptc_scene_texture_info texture_info(uint64_t i) const { ptc_scene_texture_info t; check(ptc_scene_texture_info_at(s_, i, &t)); return t; }
    // Texture::load's decode step (texture.rs:232-243): either the library decodes the files (PNG / PFM on the host threads, JPEG by
    // nvJPEG on the GPU; scene.rs:239-254 runs it on 32 threads), or the caller hands decoded texels over
    void decode_textures(int cuda_device = 0, int n_threads = 0) { check(ptc_scene_decode_textures(s_, cuda_device, n_threads)); }
    void set_texels(uint64_t i, uint32_t w, uint32_t h, uint32_t format, const void* texels) { check(ptc_scene_set_texels(s_, i, w, h, format, texels)); }
    ptc_scene_view view() const { ptc_scene_view v; check(ptc_scene_get_view(s_, &v)); return v; }
};

struct RenderStep {                                   // trait RenderStep, render_steps/mod.rs:20-25
    virtual ~RenderStep() = default;
    virtual void update(const ptc_uniforms& uniforms) = 0;
    virtual void render() = 0;
    virtual void resize(uint32_t width, uint32_t height) = 0;
};

class RTStep : public RenderStep {                    // render_steps/pt.rs:12-25
    ptc_step* h_ = nullptr;
    uint32_t w_ = 0, h_px_ = 0;
    explicit RTStep(ptc_step* h, uint32_t w, uint32_t hh) : h_(h), w_(w), h_px_(hh) {}
public:
    static RTStep create(const Scene& scene, const ptc_config& cfg, int cuda_device = 0) {   // RTStep::create, pt.rs:341-490
        ptc_scene_view v = scene.view();
        ptc_step* h = nullptr;
        check(ptc_create(&v, &cfg, cuda_device, &h));
        return RTStep(h, cfg.width, cfg.height);
    }
    RTStep(RTStep&& o) noexcept : h_(o.h_), w_(o.w_), h_px_(o.h_px_) { o.h_ = nullptr; }
    RTStep(const RTStep&) = delete; RTStep& operator=(const RTStep&) = delete;
    ~RTStep() override { if (h_) ptc_destroy(h_); }
    void update(const ptc_uniforms& u) override { check(ptc_update(h_, &u)); }               // pt.rs:492-521
    void render() override { check(ptc_render(h_)); }                                       // pt.rs:523-565 (+ queue.submit, raytracer.rs:195)
    void resize(uint32_t w, uint32_t h) override { check(ptc_resize(h_, w, h)); w_ = w; h_px_ = h; }   // pt.rs:732-745, raytracer.rs:344
    bool poll(bool wait) { int rc = ptc_poll(h_, wait ? 1 : 0); if (rc == PTC_ERR_NOT_READY) return false; check(rc); return true; }   // raytracer.rs:165,319
    void blit() { check(ptc_blit(h_, nullptr)); }                                            // BlitStep::render
    ptc_stats stats() { ptc_stats s; check(ptc_get_stats(h_, &s)); return s; }
    uint32_t width() const { return w_; }
    uint32_t height() const { return h_px_; }
    std::vector<float> read_color() { std::vector<float> v((size_t)w_ * h_px_ * 4); check(ptc_read(h_, PTC_PLANE_COLOR, v.data(), 0)); return v; }
    std::vector<unsigned char> read_display() { std::vector<unsigned char> v((size_t)w_ * h_px_ * 4); check(ptc_read(h_, PTC_PLANE_DISPLAY, v.data(), 0)); return v; }
    void write_color(const std::vector<float>& img) { check(ptc_write(h_, PTC_PLANE_COLOR, img.data(), 0)); }
    ptc_step* handle() { return h_; }
};

}  // namespace ptcore
