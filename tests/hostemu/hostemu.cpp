// hostemu.cpp — TEST-ONLY host build of the product's device functions.
//
// Compiles wgpu-pathtracer_b200/csrc/pt_device.cuh (the exact per-thread code the CUDA kernels run),
// scene_prep.cpp and bvh_build.cpp with g++ and drives them one pixel at a time, so the CPU-only test
// suite (-m "not gpu") can check the product's BVH builder, traversal and shading logic against the
// oracle without a GPU.  It is NOT a fallback: nothing in libptcore links or loads this file, and
// the kernels' queueing / compaction / texture-object paths are only exercised by the -m gpu tests.
#include "../../wgpu-pathtracer_b200/csrc/pt_device.cuh"
#include "../../wgpu-pathtracer_b200/csrc/scene_prep.h"
#include "../../wgpu-pathtracer_b200/csrc/pt_error.h"

#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace ptc {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
}
using namespace ptc;

struct emu_scene {
    PreparedScene ps;
    std::vector<TexDev> texs;
    std::vector<std::vector<uint8_t>> texel_store;
    DevScene ds;
    ptc_config cfg;
};

extern "C" {

__attribute__((visibility("default")))
const char* emu_last_error() { return g_err.c_str(); }

__attribute__((visibility("default")))
emu_scene* emu_create(const ptc_scene_view* v, const ptc_config* cfg) {
    emu_scene* e = new emu_scene();
    e->cfg = *cfg;
    if (prepare_scene(v, *cfg, e->ps) != PTC_OK) { delete e; return nullptr; }
    e->texs.resize(v->num_textures); e->texel_store.resize(v->num_textures);
    for (uint64_t i = 0; i < v->num_textures; ++i) {
        const ptc_texture_desc& t = v->textures[i];
        size_t bytes = (size_t)t.width * t.height * (t.format == PTC_TEXEL_RGBA8 ? 4 : 16);
        e->texel_store[i].assign((const uint8_t*)t.texels, (const uint8_t*)t.texels + bytes);
        TexDev& d = e->texs[i]; std::memset(&d, 0, sizeof(d));
        d.w = t.width; d.h = t.height; d.fmt = t.format; d.linear = e->texel_store[i].data();
    }
    DevScene& ds = e->ds;
    ds.nodes = reinterpret_cast<const NodeQ*>(e->ps.bvh.nodes.data());
    for (int a = 0; a < 3; ++a) { ds.grid_base[a] = e->ps.bvh.grid.base[a]; ds.grid_extent[a] = e->ps.bvh.grid.extent[a]; }
    ds.tris = reinterpret_cast<const F4*>(e->ps.bvh.tris.data());
    ds.shade = e->ps.shade.data(); ds.mats = e->ps.mats.data(); ds.texs = e->texs.data(); ds.lights = e->ps.lights.data();
    ds.num_nodes = (uint32_t)e->ps.bvh.nodes.size(); ds.num_tris = e->ps.num_tris;
    ds.num_lights = (uint32_t)e->ps.lights.size(); ds.env_tex = cfg->env_texture_index;
    ds.light_of_tri = e->ps.light_of_tri.data(); ds.num_real_lights = e->ps.num_real_lights;
    return e;
}

// accessors for warpsim.cpp (development aid in the same test-only library)
__attribute__((visibility("default")))
const DevScene* emu_dev_scene(const void* e) { return &static_cast<const emu_scene*>(e)->ds; }
__attribute__((visibility("default")))
void emu_path_cfg(const void* ev, const ptc_uniforms* un_in, PathCfg* pc, ptc_uniforms* out_un) {
    const emu_scene* e = static_cast<const emu_scene*>(ev);
    *out_un = *un_in; out_un->num_lights = e->ds.num_lights;
    pc->width = e->cfg.width; pc->height = e->cfg.height; pc->max_bounces = e->cfg.max_bounces;
    pc->send_to_light_probability = e->cfg.send_to_light_probability; pc->accum_mode = e->cfg.accum_mode; pc->is_grabbed = un_in->is_grabbed;
    pc->light_sampling = e->cfg.light_sampling; pc->spp = 1; pc->sample = 0;
}

__attribute__((visibility("default")))
void emu_destroy(emu_scene* e) { delete e; }

// FNV-1a of the host-built node / triangle arrays (compared with ptc_debug_bvh_hash of the device-built tree)
__attribute__((visibility("default")))
void emu_bvh_hash(const emu_scene* e, uint64_t out[2]) {
    auto fnv = [](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; uint64_t h = 1469598103934665603ull;
                                             for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } return h; };
    out[0] = fnv(e->ps.bvh.nodes.data(), e->ps.bvh.nodes.size() * sizeof(DevNode2));
    out[1] = fnv(e->ps.bvh.tris.data(), e->ps.bvh.tris.size() * sizeof(DevTri));
}

__attribute__((visibility("default")))
void emu_bvh_info(const emu_scene* e, uint64_t out[4]) {
    out[0] = e->ps.bvh.nodes.size(); out[1] = e->ps.bvh.tris.size(); out[2] = e->ps.max_depth; out[3] = (uint64_t)e->ps.build_ms;
}

__attribute__((visibility("default")))
int emu_render_frame(const emu_scene* e, const ptc_uniforms* un_in, float* color, float* depth, float* normal, float* albedo,
                     uint32_t* hit_id, int nthreads, uint64_t counters[3]) {
    ptc_uniforms un = *un_in;
    un.num_lights = e->ds.num_lights;
    const uint32_t W = e->cfg.width, H = e->cfg.height;
    PathCfg pc; pc.width = W; pc.height = H; pc.max_bounces = e->cfg.max_bounces;
    pc.send_to_light_probability = e->cfg.send_to_light_probability; pc.accum_mode = e->cfg.accum_mode; pc.is_grabbed = un.is_grabbed;
    pc.light_sampling = e->cfg.light_sampling; pc.spp = 1; pc.sample = 0;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    std::atomic<uint32_t> next_row(0);
    std::vector<uint64_t> rays(nthreads, 0), nn(nthreads, 0), nt(nthreads, 0);
    auto work = [&](int tid) {
        uint32_t cnt[2];
        for (;;) {
            uint32_t y = next_row.fetch_add(1);
            if (y >= H) break;
            for (uint32_t x = 0; x < W; ++x) {
                size_t pi = (size_t)y * W + x;
                PathState ps; ps.mask = pt_mk3(1.0f, 1.0f, 1.0f); ps.acc = pt_mk3(0.0f, 0.0f, 0.0f);
                gen_primary(un.camera, x, y, ps.rng, ps.o, ps.d);
                for (uint32_t seg = 0;; ++seg) {
                    cnt[0] = cnt[1] = 0;
                    HitRec h = trace_closest_bvh2(e->ds, ps.o, ps.d, cnt);
                    rays[tid]++; nn[tid] += cnt[0]; nt[tid] += cnt[1];
                    BounceOut bo = shade_bounce(e->ds, pc, ps, h, seg);
                    rays[tid] += bo.shadow_rays;
                    if (seg == 0) {
                        if (depth) depth[pi] = bo.first_t;
                        if (normal) { normal[pi * 4] = bo.first_normal.x; normal[pi * 4 + 1] = bo.first_normal.y; normal[pi * 4 + 2] = bo.first_normal.z; normal[pi * 4 + 3] = 1.0f; }
                        if (albedo) { albedo[pi * 4] = bo.first_albedo.x; albedo[pi * 4 + 1] = bo.first_albedo.y; albedo[pi * 4 + 2] = bo.first_albedo.z; albedo[pi * 4 + 3] = 1.0f; }
                        if (hit_id) {
                            if (h.tri == 0xFFFFFFFFu) hit_id[pi * 2] = hit_id[pi * 2 + 1] = 0xFFFFFFFFu;
                            else { uint32_t m = e->ps.tri_mesh[h.tri]; hit_id[pi * 2] = m; hit_id[pi * 2 + 1] = h.tri - e->ps.meshes[m].index_offset / 3; }
                        }
                    }
                    if (!bo.alive) { if (color) accumulate_pixel(reinterpret_cast<F4*>(color) + pi, bo.radiance, pc.accum_mode, pc.is_grabbed); break; }
                    ps = bo.next;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    if (counters) { counters[0] = counters[1] = counters[2] = 0; for (int t = 0; t < nthreads; ++t) { counters[0] += rays[t]; counters[1] += nn[t]; counters[2] += nt[t]; } }
    return 0;
}

__attribute__((visibility("default")))
int emu_trace_rays(const emu_scene* e, const float* rays, uint64_t n, uint32_t* hits, uint32_t* ids) {
    for (uint64_t i = 0; i < n; ++i) {
        HitRec h = trace_closest_bvh2(e->ds, pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]), pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]), nullptr);
        bool hit = h.tri != 0xFFFFFFFFu;
        float f[3] = {hit ? h.t : 0.0f, hit ? h.b1 : 0.0f, hit ? h.b2 : 0.0f};
        std::memcpy(&hits[i * 4], f, 12); hits[i * 4 + 3] = h.tri;
        if (ids) {
            if (hit) { uint32_t m = e->ps.tri_mesh[h.tri]; ids[i * 2] = m; ids[i * 2 + 1] = h.tri - e->ps.meshes[m].index_offset / 3; }
            else ids[i * 2] = ids[i * 2 + 1] = 0xFFFFFFFFu;
        }
    }
    return 0;
}

__attribute__((visibility("default")))
int emu_trace_rays_any(const emu_scene* e, const float* rays, uint64_t n, uint8_t* occluded) {
    for (uint64_t i = 0; i < n; ++i) {
        float tmax = rays[i * 8 + 7] > 0.0f ? rays[i * 8 + 7] : PT_RAY_TMAX;
        occluded[i] = trace_any_bvh2(e->ds, pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]), pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]), tmax) ? 1 : 0;
    }
    return 0;
}

// display path (device functions of pt_device.cuh on the host)
__attribute__((visibility("default")))
void emu_rayproject(const ptc_uniforms* cur, const ptc_uniforms* prev, const float* depth, uint32_t* dst_of_src) {
    const uint32_t W = prev->camera.width, H = prev->camera.height;
    for (uint32_t y = 0; y < H; ++y) for (uint32_t x = 0; x < W; ++x) {
        uint32_t dx, dy, i = y * W + x;
        dst_of_src[i] = reproject_pixel(cur->camera, prev->camera, x, y, depth[i], dx, dy) ? dy * cur->camera.width + dx : 0xFFFFFFFFu;
    }
}

__attribute__((visibility("default")))
void emu_blit(const float* rgba, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i) {
        pt_v3 t = tonemap_khronos_pbr_neutral(pt_mk3(rgba[i * 4], rgba[i * 4 + 1], rgba[i * 4 + 2]));
        out[i * 4] = (uint8_t)encode_srgb8(t.x); out[i * 4 + 1] = (uint8_t)encode_srgb8(t.y); out[i * 4 + 2] = (uint8_t)encode_srgb8(t.z);
        out[i * 4 + 3] = (uint8_t)encode_unorm8(rgba[i * 4 + 3]);
    }
}

}  // extern "C"
