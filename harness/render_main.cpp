// render_main.cpp — plays the reference's `Raytracer` (src/raytracer.rs:40-199) headless, in compiled host code over the
// C ABI: Scene::new + load_obj, Camera::new, RTStep::create, then the per-frame loop of Raytracer::render — frame counter
// from 1 (raytracer.rs:180), step.update, step.render — and a dump of what the window would show (PPM, through the blit
// pass) plus the raw accumulation (PFM).  The textures an OBJ's materials name are decoded by the library
// (ptc_scene_decode_textures: PNG / PFM on the host threads, JPEG by nvJPEG on the GPU — Texture::load, texture.rs:232-243).
// The environment map (the reference hard-codes res/knob/envmap.exr, scene.rs:93) comes as an OpenEXR or PFM file (--env), as raw
// float32 RGBA (--env-raw), or is a constant sky.
//
//   ptrender --obj scene.obj [--env sky.exr | --env-raw sky.f32 W H] [--size WxH] [--frames N] [--bounces B] [--pos x y z] [--look x y z]
//            [--fov deg] [--lbvh] [--lights-mis] [--out prefix]
#include "../include/ptcore.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

int main(int argc, char** argv) {
    std::string obj, env_raw, env_pfm, out = "ptrender";
    uint32_t W = 512, H = 512, env_w = 0, env_h = 0, frames = 16, bounces = 4;
    float pos[3] = {0, 0, 0}, fov = 60.0f;
    double look[3] = {1, 0, 0};
    bool have_look = false, lbvh = false, mis = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) { if (i + n >= argc) { fprintf(stderr, "ptrender: %s needs %d value(s)\n", a.c_str(), n); exit(2); } };
        if (a == "--obj") { need(1); obj = argv[++i]; }
        else if (a == "--env") { need(1); env_pfm = argv[++i]; }
        else if (a == "--env-raw") { need(3); env_raw = argv[++i]; env_w = (uint32_t)atoi(argv[++i]); env_h = (uint32_t)atoi(argv[++i]); }
        else if (a == "--size") { need(1); if (sscanf(argv[++i], "%ux%u", &W, &H) != 2) { fprintf(stderr, "ptrender: --size WxH\n"); return 2; } }
        else if (a == "--frames") { need(1); frames = (uint32_t)atoi(argv[++i]); }
        else if (a == "--bounces") { need(1); bounces = (uint32_t)atoi(argv[++i]); }
        else if (a == "--pos") { need(3); for (int k = 0; k < 3; ++k) pos[k] = (float)atof(argv[++i]); }
        else if (a == "--look") { need(3); for (int k = 0; k < 3; ++k) look[k] = atof(argv[++i]); have_look = true; }
        else if (a == "--fov") { need(1); fov = (float)atof(argv[++i]); }
        else if (a == "--lbvh") lbvh = true;
        else if (a == "--lights-mis") mis = true;
        else if (a == "--out") { need(1); out = argv[++i]; }
        else if (a == "--help" || a == "-h") {
            printf("ptrender --obj scene.obj [--env sky.exr|sky.hdr|sky.pfm | --env-raw sky.f32 W H] [--size WxH] [--frames N] [--bounces B] [--pos x y z] [--look x y z] [--fov deg] [--lbvh] [--lights-mis] [--out prefix]\n");
            return 0;
        } else { fprintf(stderr, "ptrender: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (obj.empty()) { fprintf(stderr, "ptrender: --obj is required (try --help)\n"); return 2; }
    try {
        // Scene::new(envmap path): texture 9 is a file like any other and is decoded with the rest (OpenEXR or PFM give float texels)
        ptcore::Scene scene(env_pfm.empty() ? "envmap" : env_pfm);           // textures 0..8 constants, 9 = environment
        std::vector<float> env;
        if (env_pfm.empty()) {
            if (!env_raw.empty()) {
                env.resize((size_t)env_w * env_h * 4);
                std::ifstream f(env_raw, std::ios::binary);
                if (!f.read((char*)env.data(), (std::streamsize)(env.size() * 4))) { fprintf(stderr, "ptrender: cannot read %s\n", env_raw.c_str()); return 2; }
            } else { env_w = env_h = 1; env = {0.5f, 0.6f, 0.8f, 1.0f}; }
            scene.set_texels(9, env_w, env_h, PTC_TEXEL_RGBA32F, env.data());
        }
        scene.load_obj(obj);
        {
            auto td = std::chrono::steady_clock::now();
            scene.decode_textures(0, 0);                                   // Texture::load on every file the materials name
            uint64_t files = 0;
            for (uint64_t t = 0; t < scene.num_textures(); ++t) files += scene.texture_info(t).source == PTC_TEX_FILE ? 1 : 0;
            if (files) printf("ptrender: %llu texture file(s) decoded in %.3f s\n", (unsigned long long)files, std::chrono::duration<double>(std::chrono::steady_clock::now() - td).count());
        }
        ptcore::Camera cam(W, H, fov, pos);                                // Camera::new
        if (have_look) {                                                   // steer as the window's mouse handler would: theta from the target
            double d[3] = {look[0] - (double)pos[0], look[1] - (double)pos[1], look[2] - (double)pos[2]};
            double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            for (double& x : d) x /= len;
            cam.c.theta_y = (float)std::asin(std::fmax(-1.0, std::fmin(1.0, d[1]))); cam.c.theta_x = (float)std::atan2(d[2], d[0]);
            cam.update(); cam.update();                                    // lookat from theta, then the basis from that lookat (camera.rs:54-79)
        }
        ptc_config cfg; std::memset(&cfg, 0, sizeof(cfg));
        cfg.width = W; cfg.height = H; cfg.max_bounces = bounces; cfg.samples_per_pixel = 1; cfg.env_texture_index = 9;
        cfg.bvh_build = lbvh ? PTC_BVH_LBVH_DEVICE : PTC_BVH_SAH_HOST; cfg.light_sampling = mis ? PTC_LIGHTS_MIS : PTC_LIGHTS_REFERENCE;
        auto t0 = std::chrono::steady_clock::now();
        ptcore::RTStep step = ptcore::RTStep::create(scene, cfg);          // RTStep::create
        double create_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        t0 = std::chrono::steady_clock::now();
        for (uint32_t frame = 1; frame <= frames; ++frame) {               // Raytracer::render, raytracer.rs:161-199
            step.update(cam.uniforms(frame));
            step.render();
        }
        step.blit();
        step.poll(true);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ptc_stats st = step.stats();
        printf("ptrender: %ux%u, %u frames in %.3f s (create %.3f s): %.1f spp/s, %.1f Mrays/s, %llu launches\n", W, H, frames, dt, create_s,
               frames / dt, (double)st.rays_traced / dt / 1e6, (unsigned long long)st.kernel_launches);
        std::vector<float> color = step.read_color();
        std::vector<unsigned char> disp = step.read_display();
        {   // PFM stores rows bottom-up, RGB
            std::ofstream f(out + ".pfm", std::ios::binary);
            f << "PF\n" << W << " " << H << "\n-1.0\n";
            for (uint32_t y = H; y-- > 0;) for (uint32_t x = 0; x < W; ++x) f.write((const char*)&color[((size_t)y * W + x) * 4], 12);
        }
        {
            std::ofstream f(out + ".ppm", std::ios::binary);
            f << "P6\n" << W << " " << H << "\n255\n";
            for (size_t i = 0; i < (size_t)W * H; ++i) f.write((const char*)&disp[i * 4], 3);
        }
        {   // raw accumulation (rgb mean + sample count), what ptc_write takes back: checkpoint
            std::ofstream f(out + ".rgba32f", std::ios::binary);
            f.write((const char*)color.data(), (std::streamsize)(color.size() * 4));
        }
    } catch (const ptcore::Error& e) {
        fprintf(stderr, "ptrender: %s\n", e.what());
        return e.status == PTC_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
