// fuzz_scene.cpp — hostile scene views against the host half of ptc_create (csrc/scene_prep.cpp: validation, the three BVH
// builders of bvh_build.cpp / lbvh.cuh, shading records, light list) and against the traversal code of pt_device.cuh compiled
// for the host, all under AddressSanitizer + UBSan (tools/fuzz.sh).  Test infrastructure, not part of libptcore.
//
// The caller's arrays are untrusted: prepare_scene must refuse what it cannot render (message set, nothing read out of
// bounds) and build a sound tree for everything else, including the geometry real exports contain — zero-area triangles,
// thousands of coincident triangles, coordinates of 1e30 next to 1e-30, non-finite coordinates.  For every accepted scene:
//   * every triangle is in exactly one leaf, every node is reached exactly once, depth < PTC_STACK_SIZE
//   * rays aimed at triangle centroids terminate and hit a triangle at least as near as a brute-force loop over
//     pt_ray_tri finds (the tree may not lose hits); hostile rays (NaN, Inf, zero direction) terminate
// usage: fuzz_scene <iterations> <seed>
#include "../wgpu-pathtracer_b200/csrc/pt_device.cuh"
#include "../wgpu-pathtracer_b200/csrc/scene_prep.h"
#include "../wgpu-pathtracer_b200/csrc/pt_error.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace ptc {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
}
using namespace ptc;

namespace {

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
    float unit() { return (float)(next() & 0xFFFFFF) / 16777216.0f; }
};

float hostile_float(Rng& r) {
    static const float F[] = {0.0f, -0.0f, 1.0f, -1.0f, 1e-30f, -1e-30f, 1e30f, -1e30f, 3.4e38f, -3.4e38f, 1e-45f, 16777216.0f, 0.01f, 10000.0f,
                              std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(), std::numeric_limits<float>::quiet_NaN()};
    return F[r.below(sizeof(F) / sizeof(F[0]))];
}

struct Scene {
    std::vector<ptc_vertex> vertices; std::vector<uint32_t> indices; std::vector<ptc_mesh> meshes; std::vector<ptc_material> materials;
    std::vector<ptc_texture_desc> textures; std::vector<std::vector<float>> texels;
    bool finite = true;          // every coordinate a triangle uses is finite
    bool sound = true;           // tables are consistent: prepare_scene must accept
};

void make_scene(Rng& r, Scene& s) {
    const uint32_t shape = r.below(8);
    const uint32_t ntri = shape == 0 ? r.below(4) : shape == 1 ? 1 + r.below(12) : 1 + r.below(400);
    const uint32_t nvert = 3 + r.below(ntri * 3 + 1);
    const float scale = r.below(4) == 0 ? std::ldexp(1.0f, (int)r.below(200) - 100) : 1.0f + 10.0f * r.unit();
    s.vertices.resize(nvert);
    for (ptc_vertex& v : s.vertices) {
        std::memset(&v, 0, sizeof(v));
        for (int k = 0; k < 3; ++k) v.position[k] = (r.unit() * 2.0f - 1.0f) * scale;
        if (shape == 2) v.position[1] = 0.25f;                               // everything in one plane: flat boxes
        if (shape == 3) { v.position[0] = v.position[1] = v.position[2] = scale; }   // every triangle is the same point
        if (shape == 4) { v.position[1] = v.position[0]; v.position[2] = v.position[0]; }   // everything on one line
        v.normal[0] = 0.0f; v.normal[1] = 1.0f; v.normal[2] = 0.0f;
        if (r.below(16) == 0) v.normal[r.below(3)] = hostile_float(r);       // normals are never validated (reference F8): must not matter here
        v.uv[0] = r.unit(); v.uv[1] = r.unit();
    }
    if (shape == 5) for (uint32_t k = 0, n = 1 + r.below(6); k < n; ++k) s.vertices[r.below(nvert)].position[r.below(3)] = hostile_float(r);
    s.indices.resize((size_t)ntri * 3);
    for (uint32_t& ix : s.indices) ix = r.below(nvert);
    if (shape == 6 && ntri > 1) for (uint32_t t = 1; t < ntri; ++t) for (int k = 0; k < 3; ++k) s.indices[t * 3 + k] = s.indices[k];   // N copies of one triangle
    // meshes: consecutive ranges covering every triangle
    uint32_t nmat = 1 + r.below(4);
    for (uint32_t t0 = 0; t0 < ntri;) {
        const uint32_t n = 1 + r.below(ntri - t0);
        ptc_mesh m; m.vertex_offset = 0; m.vertex_count = nvert; m.index_offset = t0 * 3; m.index_count = n * 3; m.material_index = r.below(nmat);
        s.meshes.push_back(m); t0 += n;
    }
    // textures 0..9 as Scene::new lays them out (constants), materials over them; texture 4 = "not emissive"
    s.texels.resize(11); s.textures.resize(11);
    for (size_t t = 0; t < s.textures.size(); ++t) {
        const uint32_t w = t == 10 ? 3 : 1, h = t == 10 ? 2 : 1;
        s.texels[t].assign((size_t)w * h * 4, t == 4 ? 0.0f : 0.5f);
        s.textures[t] = ptc_texture_desc{w, h, PTC_TEXEL_RGBA32F, 0, s.texels[t].data()};
    }
    s.materials.resize(nmat);
    for (ptc_material& m : s.materials) {
        std::memset(&m, 0, sizeof(m));
        m.albedo_texture_idx = r.below(4) == 0 ? 10 : 0; m.roughness_texture_idx = 1; m.specular_texture_idx = 2; m.metallic_texture_idx = 3;
        m.emissive_texture_idx = r.below(3) == 0 ? 0 : 4; m.sheen_texture_idx = 5; m.normal_texture_idx = 6; m.transmission_texture_idx = 7; m.ior_texture_idx = 8;
    }
    // break one table entry now and then: prepare_scene has to notice
    switch (r.below(24)) {
        case 0: if (!s.indices.empty()) { s.indices[r.below((uint32_t)s.indices.size())] = nvert + r.below(2) * 0x7fffffffu; s.sound = false; } break;
        case 1: if (!s.meshes.empty()) { s.meshes[r.below((uint32_t)s.meshes.size())].material_index = nmat + r.below(2) * 0xfffffff0u; s.sound = false; } break;
        case 2: if (!s.meshes.empty()) { ptc_mesh& m = s.meshes[r.below((uint32_t)s.meshes.size())]; m.index_count += 3 * (1 + r.below(3)) + (r.below(2) ? 0xfffffff0u - m.index_count : 0); if ((uint64_t)m.index_offset + m.index_count > s.indices.size() || m.index_count % 3) s.sound = false; } break;
        case 3: if (s.meshes.size() > 1) { s.meshes.erase(s.meshes.begin() + r.below((uint32_t)s.meshes.size())); s.sound = false; } break;   // triangles outside every mesh
        case 4: s.materials[r.below(nmat)].albedo_texture_idx = 11 + r.below(2) * 0xfffffff0u; s.sound = false; break;
        case 5: if (!s.meshes.empty()) { s.meshes[r.below((uint32_t)s.meshes.size())].index_offset += 1; s.sound = false; } break;
        case 6: s.textures[r.below(11)].texels = nullptr; s.sound = false; break;
        case 7: s.textures[r.below(11)].width = 0; s.sound = false; break;
        default: break;
    }
    for (uint32_t ix : s.indices) if (ix < nvert) for (int k = 0; k < 3; ++k) if (!std::isfinite(s.vertices[ix].position[k])) s.finite = false;
}

#define FAIL(...) do { std::fprintf(stderr, "iteration %ld builder %u: ", it, builder); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); return 1; } while (0)

}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <iterations> <seed>\n", argv[0]); return 2; }
    const long iters = std::atol(argv[1]);
    Rng r{(uint64_t)std::strtoull(argv[2], nullptr, 0) * 2654435761ull + 1};
    long accepted = 0, refused = 0, nonfinite = 0; uint64_t rays = 0, hits = 0;
    for (long it = 0; it < iters; ++it) {
        Scene sc; make_scene(r, sc);
        ptc_scene_view v; std::memset(&v, 0, sizeof(v));
        v.vertices = sc.vertices.data(); v.num_vertices = sc.vertices.size(); v.indices = sc.indices.data(); v.num_indices = sc.indices.size();
        v.meshes = sc.meshes.data(); v.num_meshes = sc.meshes.size(); v.materials = sc.materials.data(); v.num_materials = sc.materials.size();
        v.textures = sc.textures.data(); v.num_textures = sc.textures.size();
        for (uint32_t builder = 0; builder < 3; ++builder) {
            ptc_config cfg; std::memset(&cfg, 0, sizeof(cfg));
            cfg.width = cfg.height = 8; cfg.max_bounces = 4; cfg.samples_per_pixel = 1; cfg.env_texture_index = 9; cfg.bvh_build = builder;
            if (builder == 2 && r.below(2)) cfg.ploc_window = 1 + r.below(64);
            PreparedScene ps; g_err.clear();
            const int rc = prepare_scene(&v, cfg, ps);
            if (rc != PTC_OK) {
                if (g_err.empty()) FAIL("refusal %d without a message", rc);
                if (sc.sound && sc.finite) FAIL("a consistent, finite scene was refused: %s", g_err.c_str());
                ++refused; continue;
            }
            if (!sc.sound) FAIL("an inconsistent scene was accepted");
            if (!sc.finite) ++nonfinite;
            ++accepted;
            const uint32_t ntri = (uint32_t)sc.indices.size() / 3, nnodes = (uint32_t)ps.bvh.nodes.size();
            if (ps.bvh.tris.size() != ntri) FAIL("%zu leaf triangles for %u triangles", ps.bvh.tris.size(), ntri);
            if (ps.max_depth >= PTC_STACK_SIZE) FAIL("depth %u", ps.max_depth);
            // walk the tree: every node once, every triangle slot once, every triangle id once
            std::vector<uint8_t> node_seen(nnodes, 0), slot_seen(ntri, 0), id_seen(ntri, 0);
            if (nnodes) {
                std::vector<std::pair<uint32_t, uint32_t>> stack; stack.push_back({0u, 1u});
                while (!stack.empty()) {
                    const uint32_t n = stack.back().first, depth = stack.back().second; stack.pop_back();
                    if (n >= nnodes) FAIL("child %u of %u nodes", n, nnodes);
                    if (node_seen[n]++) FAIL("node %u reached twice", n);
                    if (depth > PTC_STACK_SIZE) FAIL("walk deeper than the traversal stack");
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t ref = ps.bvh.nodes[n].w[c * 4 + 3];
                        const bool absent = (ps.bvh.nodes[n].w[c * 4] & 0x7fffu) > ((ps.bvh.nodes[n].w[c * 4] >> 16) & 0x7fffu);
                        if (absent) continue;
                        if ((int32_t)ref >= 0) { stack.push_back({ref, depth + 1}); continue; }
                        const uint32_t leaf = ~ref, first = leaf >> 3, count = (leaf & 7u) + 1;
                        if ((uint64_t)first + count > ntri) FAIL("leaf [%u, +%u) of %u triangles", first, count, ntri);
                        for (uint32_t k = first; k < first + count; ++k) {
                            if (slot_seen[k]++) FAIL("triangle slot %u in two leaves", k);
                            uint32_t id; std::memcpy(&id, &ps.bvh.tris[k].q[3], 4);
                            if (id >= ntri || id_seen[id]++) FAIL("triangle id %u (slot %u) out of range or twice", id, k);
                        }
                    }
                }
            }
            for (uint32_t k = 0; k < ntri; ++k) if (!slot_seen[k]) FAIL("triangle slot %u in no leaf", k);
            for (uint32_t n = 0; n < nnodes; ++n) if (!node_seen[n]) FAIL("node %u unreachable", n);
            // trace: centroid-aimed rays must find a hit no farther than brute force does; hostile rays must terminate
            DevScene ds; std::memset(&ds, 0, sizeof(ds));
            ds.nodes = reinterpret_cast<const NodeQ*>(ps.bvh.nodes.data()); ds.tris = reinterpret_cast<const F4*>(ps.bvh.tris.data());
            for (int a = 0; a < 3; ++a) { ds.grid_base[a] = ps.bvh.grid.base[a]; ds.grid_extent[a] = ps.bvh.grid.extent[a]; }
            ds.num_nodes = nnodes; ds.num_tris = ntri;
            const uint32_t nrays = ntri ? 6 : 2;
            for (uint32_t q = 0; q < nrays; ++q) {
                pt_v3 o, d;
                if (q < 4 && ntri) {
                    const uint32_t t = r.below(ntri);
                    float c[3];
                    for (int k = 0; k < 3; ++k) c[k] = (sc.vertices[sc.indices[t * 3]].position[k] + sc.vertices[sc.indices[t * 3 + 1]].position[k] + sc.vertices[sc.indices[t * 3 + 2]].position[k]) / 3.0f;
                    o = pt_mk3(c[0] + (r.unit() - 0.5f) * 4.0f, c[1] + 3.0f + r.unit(), c[2] + (r.unit() - 0.5f) * 4.0f);
                    d = pt_mk3(c[0] - o.x, c[1] - o.y, c[2] - o.z);
                    const float len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
                    if (!(len > 0.0f) || !std::isfinite(len)) continue;
                    d = pt_mk3(d.x / len, d.y / len, d.z / len);
                } else {
                    o = pt_mk3(hostile_float(r), r.unit(), hostile_float(r)); d = pt_mk3(hostile_float(r), hostile_float(r), r.unit());
                }
                uint32_t cnt[2] = {0, 0};
                const HitRec h = trace_closest_bvh2(ds, o, d, cnt);
                const bool occluded = trace_any_bvh2(ds, o, d, PT_RAY_TMAX);
                ++rays;
                if (cnt[0] > 4u * nnodes + 4u) FAIL("a ray visited %u nodes of %u", cnt[0], nnodes);
                if (!sc.finite || !ray_is_finite(o, d)) continue;            // the answer is only defined for finite geometry and rays
                // brute force over the same leaf triangles with the same intersection routine and tie rule.  Triangles without
                // area (collinear corners) are set apart: their determinant is rounding noise, pt_ray_tri may report a "hit" far
                // outside the triangle's box, and whether a tree ever tests them depends on the tree
                float bt = PT_RAY_TMAX; uint32_t bid = 0xFFFFFFFFu; bool sliver_hit = false;
                std::vector<uint8_t> sliver(ntri, 0);
                for (uint32_t k = 0; k < ntri; ++k) {
                    const F4* T = reinterpret_cast<const F4*>(&ps.bvh.tris[k]);
                    float t, b1, b2; uint32_t id; std::memcpy(&id, &T[0].w, 4);
                    const double e1[3] = {T[1].x, T[1].y, T[1].z}, e2[3] = {T[2].x, T[2].y, T[2].z};
                    const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
                    const double l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
                    const bool flat = !(cx * cx + cy * cy + cz * cz > 1e-8 * l1 * l2);
                    if (flat) sliver[id] = 1;
                    if (!pt_ray_tri(o, d, pt_mk3(T[0].x, T[0].y, T[0].z), pt_mk3(T[1].x, T[1].y, T[1].z), pt_mk3(T[2].x, T[2].y, T[2].z), PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) continue;
                    if (flat) { sliver_hit = true; continue; }
                    if (t < bt || (t == bt && id < bid)) { bt = t; bid = id; }
                }
                const bool tree_sliver = h.tri != 0xFFFFFFFFu && sliver[h.tri];
                if (tree_sliver ? !sliver_hit : (bid != h.tri || (bid != 0xFFFFFFFFu && bt != h.t))) FAIL("ray %u: tree says triangle %u at %g, brute force %u at %g", q, h.tri, (double)h.t, bid, (double)bt);
                if (bid != 0xFFFFFFFFu ? !occluded : (occluded && !sliver_hit)) FAIL("ray %u: any-hit %d, brute force %d", q, (int)occluded, (int)(bid != 0xFFFFFFFFu));
                if (bid != 0xFFFFFFFFu) ++hits;
            }
        }
    }
    std::printf("fuzz_scene: %ld scene x builder pairs accepted (%ld of them with non-finite coordinates), %ld refused with a message; %llu rays traced, %llu hits, all equal to brute force; 0 sanitizer reports\n",
                accepted, nonfinite, refused, (unsigned long long)rays, (unsigned long long)hits);
    return 0;
}
