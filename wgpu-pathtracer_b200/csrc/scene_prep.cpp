// scene_prep.cpp — see scene_prep.h.
#include "scene_prep.h"
#include "pt_error.h"
#include "lbvh.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

namespace ptc {

// f(begin, end) over [0, n) on the host threads (scene preparation is per-triangle work on up to ~10^7 triangles)
template <class F>
static void parallel_for(uint64_t n, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (n < 65536 || nt == 1) { f((uint64_t)0, n); return; }
    nt = (unsigned)std::min<uint64_t>(nt, 32);
    std::vector<std::thread> th;
    const uint64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto& x : th) x.join();
}

static void texel_of(const ptc_texture_desc& t, float out[4]) {   // the (0,0) texel as sample_texture_* sees it
    if (t.format == PTC_TEXEL_RGBA8) {
        const uint8_t* p = (const uint8_t*)t.texels;
        for (int k = 0; k < 4; ++k) out[k] = pt_div((float)p[k], 255.0f);
    } else {
        const float* p = (const float*)t.texels;
        for (int k = 0; k < 4; ++k) out[k] = p[k];
    }
}

int prepare_scene(const ptc_scene_view* v, const ptc_config& cfg, PreparedScene& out, bool caller_builds_bvh) {
    const uint64_t ntri64 = v->num_indices / 3;
    if (v->num_indices % 3 != 0) { set_error("index count is not a multiple of 3"); return PTC_ERR_INVALID; }
    if (ntri64 >= (1ull << 28)) { set_error("too many triangles"); return PTC_ERR_INVALID; }
    if (v->num_textures > 1024) { set_error("more than MAX_TEXTURES (1024) textures"); return PTC_ERR_INVALID; }
    if (v->num_materials == 0) { set_error("scene has no materials (Scene::new always provides material 0)"); return PTC_ERR_INVALID; }
    if (cfg.env_texture_index >= v->num_textures) { set_error("env_texture_index out of range"); return PTC_ERR_INVALID; }
    const uint32_t ntri = (uint32_t)ntri64;
    out.num_tris = ntri;
    {
        std::atomic<bool> bad(false);
        parallel_for(v->num_indices, [&](uint64_t b, uint64_t e) {
            for (uint64_t i = b; i < e; ++i) if (v->indices[i] >= v->num_vertices) { bad = true; return; }
        });
        if (bad) { set_error("vertex index out of range"); return PTC_ERR_INVALID; }
    }
    for (uint64_t i = 0; i < v->num_textures; ++i) {
        const ptc_texture_desc& t = v->textures[i];
        if (!t.texels || t.width == 0 || t.height == 0 || t.format > PTC_TEXEL_RGBA32F) { set_error("texture " + std::to_string(i) + " has no texels"); return PTC_ERR_INVALID; }
    }
    // triangle -> mesh (instance); the reference builds one BLAS per mesh (pt.rs:364-433), so a triangle outside every mesh is never traced
    out.meshes.assign(v->meshes, v->meshes + v->num_meshes);
    out.tri_mesh.assign(ntri, 0xFFFFFFFFu);
    for (uint32_t m = 0; m < v->num_meshes; ++m) {
        const ptc_mesh& me = v->meshes[m];
        if (me.index_offset % 3 || me.index_count % 3 || (uint64_t)me.index_offset + me.index_count > v->num_indices) { set_error("mesh index range invalid"); return PTC_ERR_INVALID; }
        if (me.material_index >= v->num_materials) { set_error("mesh material index out of range"); return PTC_ERR_INVALID; }
        for (uint32_t t = me.index_offset / 3; t < (me.index_offset + me.index_count) / 3; ++t) out.tri_mesh[t] = m;
    }
    for (uint32_t t = 0; t < ntri; ++t) if (out.tri_mesh[t] == 0xFFFFFFFFu) { set_error("triangle outside every mesh"); return PTC_ERR_INVALID; }

    // BVH over every triangle
    std::vector<float> pos((size_t)ntri * 9);
    auto finite9 = [](const float* p) { for (int k = 0; k < 9; ++k) if (!(std::fabs(p[k]) <= 1e37f)) return false; return true; };
    std::atomic<bool> any_inactive(false);
    parallel_for(ntri, [&](uint64_t b, uint64_t e) {
        bool seen = false;
        for (uint64_t t = b; t < e; ++t) {
            for (int k = 0; k < 3; ++k) std::memcpy(&pos[(size_t)t * 9 + k * 3], v->vertices[v->indices[(size_t)t * 3 + k]].position, 12);
            seen = seen || !finite9(&pos[(size_t)t * 9]);
        }
        if (seen) any_inactive = true;
    });
    // A triangle with a non-finite corner is INACTIVE, as in the acceleration structures the reference builds (Vulkan: a
    // NaN position makes a triangle inactive, it is never hit; pt.rs:391-433 hands the vertex buffer to the driver as it is).
    // pt_ray_tri would never report it either (every comparison with NaN fails); the builders, though, bin and sort by
    // coordinates, so the tree's copy of such a triangle collapses to one point of the scene (zero edges: det == 0, no hit)
    // and its box neither poisons the bounds nor the Morton keys.  The same goes for coordinates beyond 1e37: edge vectors,
    // box extents and the quantisation grid overflow fp32 there, and no ray (tmax = 10000) resolves them.
    // Triangle numbering and shading records are unchanged.
    if (any_inactive) {
        float anchor[3] = {0.0f, 0.0f, 0.0f};
        for (uint32_t t = 0; t < ntri; ++t) if (finite9(&pos[(size_t)t * 9])) { std::memcpy(anchor, &pos[(size_t)t * 9], 12); break; }
        parallel_for(ntri, [&](uint64_t b, uint64_t e) {
            for (uint64_t t = b; t < e; ++t) if (!finite9(&pos[(size_t)t * 9])) for (int k = 0; k < 3; ++k) std::memcpy(&pos[(size_t)t * 9 + k * 3], anchor, 12);
        });
    }
    const bool ploc = cfg.bvh_build == PTC_BVH_PLOC_DEVICE && ntri > 4;
    const bool lbvh = ploc || (cfg.bvh_build == PTC_BVH_LBVH_DEVICE && ntri > PTC_LBVH_MAX_LEAF);
    if (lbvh && caller_builds_bvh) out.tri_pos = pos;
    else {
        bool sah = !lbvh;
        if (lbvh) {
            if (ploc) build_ploc_host(pos.data(), ntri, out.bvh, out.max_depth, cfg.ploc_window ? (int)cfg.ploc_window : PTC_PLOC_RADIUS);
            else build_lbvh_host(pos.data(), ntri, out.bvh, out.max_depth);
            if (out.max_depth >= PTC_STACK_SIZE) { out.bvh = DevBvh2{}; sah = true; }     // degenerate input: fall back to the SAH tree
        }
        if (sah) {
            Bvh2 bvh; build_bvh2(pos.data(), ntri, 4, 0, bvh);
            if (bvh.max_depth >= PTC_STACK_SIZE) { set_error("BVH too deep for the traversal stack"); return PTC_ERR_INVALID; }
            emit_bvh2(bvh, pos.data(), out.bvh);
            out.max_depth = bvh.max_depth; out.build_ms = bvh.build_ms;
        }
    }

    // per-triangle shading records: re-pack of vertices[] / indices[] / instance_infos[] (rt_helpers.wgsl:57-67)
    out.shade.resize(ntri);
    parallel_for(ntri, [&](uint64_t b, uint64_t e) {
        for (uint64_t t = b; t < e; ++t) {
            const ptc_vertex* vx[3];
            for (int k = 0; k < 3; ++k) vx[k] = &v->vertices[v->indices[(size_t)t * 3 + k]];
            uint32_t mat = v->meshes[out.tri_mesh[t]].material_index;
            TriShade& r = out.shade[t];
            r.q[0] = F4{vx[0]->normal[0], vx[0]->normal[1], vx[0]->normal[2], vx[0]->uv[0]};
            r.q[1] = F4{vx[1]->normal[0], vx[1]->normal[1], vx[1]->normal[2], vx[0]->uv[1]};
            r.q[2] = F4{vx[2]->normal[0], vx[2]->normal[1], vx[2]->normal[2], vx[1]->uv[0]};
            r.q[3] = F4{vx[1]->uv[1], vx[2]->uv[0], vx[2]->uv[1], u2f(mat)};
        }
    });

    // materials: 1x1 textures folded to constants (sample_material, rt_helpers.wgsl:17-42)
    out.mats.resize(v->num_materials);
    for (uint64_t i = 0; i < v->num_materials; ++i) {
        const ptc_material& m = v->materials[i];
        MatDev& d = out.mats[i]; std::memset(&d, 0, sizeof(d));
        const uint32_t idx[SLOT_COUNT] = {m.albedo_texture_idx, m.normal_texture_idx, m.roughness_texture_idx, m.metallic_texture_idx,
                                          m.emissive_texture_idx, m.transmission_texture_idx, m.ior_texture_idx};
        const uint32_t all[9] = {m.albedo_texture_idx, m.normal_texture_idx, m.roughness_texture_idx, m.metallic_texture_idx, m.emissive_texture_idx,
                                 m.transmission_texture_idx, m.ior_texture_idx, m.specular_texture_idx, m.sheen_texture_idx};
        for (uint32_t a : all) if (a >= v->num_textures) { set_error("material texture index out of range"); return PTC_ERR_INVALID; }
        d.anisotropy = m.anisotropy;
        for (int slot = 0; slot < SLOT_COUNT; ++slot) {
            const ptc_texture_desc& t = v->textures[idx[slot]];
            d.tex[slot] = idx[slot];
            if (t.width != 1 || t.height != 1) { d.tex_mask |= 1u << slot; continue; }
            float c[4]; texel_of(t, c);
            switch (slot) {
                case SLOT_ALBEDO: d.albedo[0] = c[0]; d.albedo[1] = c[1]; d.albedo[2] = c[2]; break;
                case SLOT_NORMAL: d.tsn[0] = c[0] * 2.0f - 1.0f; d.tsn[1] = c[1] * 2.0f - 1.0f; d.tsn[2] = c[2] * 2.0f - 1.0f; break;
                case SLOT_ROUGHNESS: d.roughness = c[0]; break;
                case SLOT_METALLIC: d.metallic = c[0]; break;
                case SLOT_EMISSIVE: d.emissive[0] = c[0]; d.emissive[1] = c[1]; d.emissive[2] = c[2]; break;
                case SLOT_TRANSMISSION: d.transmission = c[0]; break;
                case SLOT_IOR: d.ior = c[0]; break;
            }
        }
    }

    // emissive triangle list (pt.rs:206-243): a triangle is a light iff its material's emissive texture index != 4
    out.lights.clear();
    out.light_of_tri.assign(ntri, 0xFFFFFFFFu);
    for (uint32_t t = 0; t < ntri; ++t) {
        uint32_t mat = v->meshes[out.tri_mesh[t]].material_index;
        if (v->materials[mat].emissive_texture_idx == 4) continue;
        ptc_triangle L; std::memset(&L, 0, sizeof(L));
        const uint32_t* ix = &v->indices[(size_t)t * 3];
        std::memcpy(L.v0, v->vertices[ix[0]].position, 12); std::memcpy(L.v1, v->vertices[ix[1]].position, 12);
        std::memcpy(L.v2, v->vertices[ix[2]].position, 12); std::memcpy(L.normal, v->vertices[ix[0]].normal, 12);
        L._pad0 = t;
        out.light_of_tri[t] = (uint32_t)out.lights.size();
        out.lights.push_back(L);
    }
    out.num_real_lights = (uint32_t)out.lights.size();
    if (out.lights.empty()) { ptc_triangle z; std::memset(&z, 0, sizeof(z)); out.lights.push_back(z); }   // pt.rs:232-239
    return PTC_OK;
}

}  // namespace ptc
