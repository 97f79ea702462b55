// pt_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the reference's path-tracing compute pass
// (gamingdoom/wgpu-pathtracer, src/shaders/pathtracer/**, src/shaders/common/**)
// and of the host code that feeds it (src/camera.rs, src/render_steps/pt.rs).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library; the product (libptcore) never does.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and
// cannot be compiled or run here (no Rust toolchain, no Vulkan ICD),
// so this restatement is pinned only by (a) known-answer vectors derived by
// hand from random.wgsl / camera.rs (tests/test_oracle_kat.py) and (b) analytic
// fixtures (furnace, empty scene, single triangle).  Three pieces of arithmetic
// the reference delegates to un-vendored third parties are DEFINED here and in
// include/ptmath.h rather than restated: ray/triangle intersection + traversal
// (Vulkan driver via wgpu fork gamingdoom/wgpu@dc10a55), bilinear filter
// weights (GPU texture unit) and sin/cos/atan/acos ULPs (driver).
//
// Numerics primitives come from include/ptmath.h (shared with the CUDA kernels
// so both sides round identically); the algorithm below is written
// independently of wgpu-pathtracer_b200/csrc.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -mfma).

#include "../include/ptcore_types.h"
#include "../include/ptmath.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// RNG — src/shaders/common/random.wgsl:7-36 (u32 wrapping arithmetic)
// ---------------------------------------------------------------------------
struct Rng { uint32_t state; };

uint32_t jenkins_hash(uint32_t x) {           // random.wgsl:7-15
    x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15;
    return x;
}
uint32_t pcg_hash(Rng& r) {                   // random.wgsl:17-22
    r.state = r.state * 747796405u + 2891336453u;
    uint32_t state = r.state;
    uint32_t word = ((state >> ((state >> 28) + 4u)) ^ state) * 277803737u;
    return (word >> 22) ^ word;
}
void rand_init(Rng& r, uint32_t px, uint32_t py, uint32_t pz, uint32_t res_x, uint32_t frame) {  // random.wgsl:24-28
    uint32_t seed = ((px * 1u + py * res_x) * pz) ^ jenkins_hash(frame);
    r.state = jenkins_hash(seed);
}
uint32_t rand_int(Rng& r) { return pcg_hash(r); }                       // random.wgsl:30-32
float rand_float(Rng& r) {                                              // random.wgsl:34-36
    return pt_div(pt_u32_to_f32(rand_int(r)), pt_u32_to_f32(0xffffffffu));
}

// ---------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------
struct Tex { uint32_t w, h, fmt; std::vector<uint8_t> bytes; };

struct BNode {                 // BVH2 node: inner => left/right children; leaf => [first, first+count)
    float lo[3], hi[3];
    int32_t left, right;       // inner: child node ids; leaf: left = -1
    uint32_t first, count;
};

struct Scene {
    std::vector<ptc_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<ptc_mesh> meshes;
    std::vector<ptc_material> materials;
    std::vector<Tex> textures;
    std::vector<ptc_triangle> lights;     // pt.rs:206-243
    std::vector<uint32_t> light_tri;      // PTC_LIGHTS_MIS: global triangle id of lights[i]
    std::vector<uint32_t> light_of_tri;   //                 global triangle id -> light index (0xFFFFFFFF: none)
    bool has_lights = false;              //                 false when lights holds only the dummy entry
    std::vector<uint32_t> tri_mesh;       // global triangle -> mesh (instance_index)
    ptc_config cfg;
    // traversal structure
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;          // leaf-ordered global triangle ids
};

inline pt_v3 v3(const float* p) { return pt_mk3(p[0], p[1], p[2]); }

// ---------------------------------------------------------------------------
// textures — helpers.wgsl:35-48 (textureSampleLevel, lod 0), sampler
// texture.rs:312-322 (Repeat, mag Linear).  8-bit texels are value/255 with no
// sRGB decode (texture.rs:241).
// ---------------------------------------------------------------------------
pt_v4 texel(const Tex& t, int x, int y) {
    size_t i = (size_t)y * t.w + (size_t)x;
    if (t.fmt == PTC_TEXEL_RGBA8) {
        const uint8_t* p = &t.bytes[i * 4];
        return pt_mk4(pt_div((float)p[0], 255.0f), pt_div((float)p[1], 255.0f),
                      pt_div((float)p[2], 255.0f), pt_div((float)p[3], 255.0f));
    }
    const float* p = reinterpret_cast<const float*>(&t.bytes[i * 16]);
    return pt_mk4(p[0], p[1], p[2], p[3]);
}
float lerp2(float a, float b, float c, float d, float fx, float fy) {
    float wx = 1.0f - fx;
    float top = pt_fma(b, fx, a * wx);
    float bot = pt_fma(d, fx, c * wx);
    return pt_fma(bot, fy, top * (1.0f - fy));
}
pt_v4 sample_texture_rgba(const Scene& sc, uint32_t idx, float u, float v) {
    const Tex& t = sc.textures[idx];
    if (t.w == 1 && t.h == 1) return texel(t, 0, 0);      // constant textures are exact
    int x0, x1, y0, y1; float fx, fy;
    pt_bilinear_axis(u, (int)t.w, &x0, &x1, &fx);
    pt_bilinear_axis(v, (int)t.h, &y0, &y1, &fy);
    pt_v4 a = texel(t, x0, y0), b = texel(t, x1, y0), c = texel(t, x0, y1), d = texel(t, x1, y1);
    return pt_mk4(lerp2(a.x, b.x, c.x, d.x, fx, fy), lerp2(a.y, b.y, c.y, d.y, fx, fy),
                  lerp2(a.z, b.z, c.z, d.z, fx, fy), lerp2(a.w, b.w, c.w, d.w, fx, fy));
}
pt_v3 sample_texture_color(const Scene& sc, uint32_t idx, float u, float v) {
    pt_v4 r = sample_texture_rgba(sc, idx, u, v); return pt_mk3(r.x, r.y, r.z);
}
float sample_texture_float(const Scene& sc, uint32_t idx, float u, float v) {
    return sample_texture_rgba(sc, idx, u, v).x;
}

struct SampledMaterial {       // helpers.wgsl:50-70
    pt_v3 albedo; float alpha;
    pt_v3 tangent_space_normal;
    float roughness; pt_v3 specular; float metallic;
    pt_v3 emissive; float sheen;
    float clearcoat_thickness, clearcoat_roughness, anisotropy, anisotropy_rotation;
    float transmission_weight, ior;
};

SampledMaterial sample_material(const Scene& sc, const ptc_material& m, float u, float v) {   // rt_helpers.wgsl:17-42
    SampledMaterial s;
    pt_v4 rgba = sample_texture_rgba(sc, m.albedo_texture_idx, u, v);
    s.albedo = pt_mk3(rgba.x, rgba.y, rgba.z); s.alpha = rgba.w;
    pt_v3 n = sample_texture_color(sc, m.normal_texture_idx, u, v);
    s.tangent_space_normal = pt_mk3(n.x * 2.0f - 1.0f, n.y * 2.0f - 1.0f, n.z * 2.0f - 1.0f);
    s.roughness = sample_texture_float(sc, m.roughness_texture_idx, u, v);
    s.specular = sample_texture_color(sc, m.specular_texture_idx, u, v);
    s.metallic = sample_texture_float(sc, m.metallic_texture_idx, u, v);
    s.emissive = sample_texture_color(sc, m.emissive_texture_idx, u, v);
    s.sheen = sample_texture_float(sc, m.sheen_texture_idx, u, v);
    s.clearcoat_thickness = m.clearcoat_thickness; s.clearcoat_roughness = m.clearcoat_roughness;
    s.anisotropy = m.anisotropy; s.anisotropy_rotation = m.anisotropy_rotation;
    s.transmission_weight = sample_texture_float(sc, m.transmission_texture_idx, u, v);
    s.ior = sample_texture_float(sc, m.ior_texture_idx, u, v);
    return s;
}

// ---------------------------------------------------------------------------
// closest hit — rt_helpers.wgsl:44-80.  The reference calls the driver's ray
// query; here: the shared pt_ray_tri, closest t, ties broken towards the
// smaller global triangle id, over every triangle of every mesh (identity
// instance transforms, mask 0xFF, no culling; pt.rs:527-542).
// ---------------------------------------------------------------------------
struct Hit { bool hit; float t, b1, b2; uint32_t tri; };   // tri = global triangle number (index/3)

struct TraceCounters { uint64_t rays = 0, nodes = 0, tris = 0; };

inline void tri_verts(const Scene& sc, uint32_t tri, pt_v3& v0, pt_v3& e1, pt_v3& e2) {
    const uint32_t* ix = &sc.indices[(size_t)tri * 3];
    v0 = v3(sc.vertices[ix[0]].position);
    e1 = pt_sub(v3(sc.vertices[ix[1]].position), v0);
    e2 = pt_sub(v3(sc.vertices[ix[2]].position), v0);
}
inline void consider(const Scene& sc, uint32_t tri, pt_v3 o, pt_v3 d, Hit& best) {
    pt_v3 v0, e1, e2; tri_verts(sc, tri, v0, e1, e2);
    float t, b1, b2;
    if (!pt_ray_tri(o, d, v0, e1, e2, PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) return;
    if (!best.hit || t < best.t || (t == best.t && tri < best.tri)) {
        best.hit = true; best.t = t; best.b1 = b1; best.b2 = b2; best.tri = tri;
    }
}
Hit closest_brute(const Scene& sc, pt_v3 o, pt_v3 d, TraceCounters& c) {
    Hit best{false, 0, 0, 0, 0xFFFFFFFFu};
    uint32_t ntri = (uint32_t)(sc.indices.size() / 3);
    for (uint32_t i = 0; i < ntri; ++i) consider(sc, i, o, d, best);
    c.rays++; c.tris += ntri;
    return best;
}
// slab test against a padded box; returns the entry distance or -1 for a miss.  Conservative: the exit
// distance is widened by a few ulps and NaNs (0 * inf, ray inside a slab plane) drop out of fmin/fmax.
inline float box_entry(const float* lo, const float* hi, pt_v3 o, pt_v3 inv, float tbest) {
    float tx0 = (lo[0] - o.x) * inv.x, tx1 = (hi[0] - o.x) * inv.x;
    float ty0 = (lo[1] - o.y) * inv.y, ty1 = (hi[1] - o.y) * inv.y;
    float tz0 = (lo[2] - o.z) * inv.z, tz1 = (hi[2] - o.z) * inv.z;
    float tn = std::fmax(std::fmax(std::fmin(tx0, tx1), std::fmin(ty0, ty1)), std::fmax(std::fmin(tz0, tz1), 0.0f));
    float tf = std::fmin(std::fmin(std::fmax(tx0, tx1), std::fmax(ty0, ty1)), std::fmin(std::fmax(tz0, tz1), tbest));
    return tn <= tf * 1.0000005f ? tn : -1.0f;
}
// Ordered (near child first) stack traversal; `nodes` counts inner nodes visited (each visit tests the two
// child boxes = one 64-byte Aila-Laine-style node fetch), `tris` counts triangle tests.
Hit closest_bvh(const Scene& sc, pt_v3 o, pt_v3 d, TraceCounters& c) {
    Hit best{false, 0, 0, 0, 0xFFFFFFFFu};
    c.rays++;
    if (sc.nodes.empty()) return best;
    pt_v3 inv = pt_mk3(pt_div(1.0f, d.x), pt_div(1.0f, d.y), pt_div(1.0f, d.z));
    int32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const BNode& n = sc.nodes[stack[--sp]];
        if (n.left < 0) {
            for (uint32_t i = 0; i < n.count; ++i) { consider(sc, sc.order[n.first + i], o, d, best); c.tris++; }
            continue;
        }
        c.nodes++;
        float tb = best.hit ? best.t : PT_RAY_TMAX;
        const BNode& l = sc.nodes[n.left]; const BNode& r = sc.nodes[n.right];
        float el = box_entry(l.lo, l.hi, o, inv, tb), er = box_entry(r.lo, r.hi, o, inv, tb);
        if (el >= 0.0f && er >= 0.0f) {
            if (er < el) { stack[sp++] = n.left; stack[sp++] = n.right; }      // far first, near on top
            else { stack[sp++] = n.right; stack[sp++] = n.left; }
        } else if (el >= 0.0f) stack[sp++] = n.left;
        else if (er >= 0.0f) stack[sp++] = n.right;
    }
    return best;
}

// --- oracle's own BVH2 builder: recursive binned SAH (16 bins), leaves <= 4 triangles -----
struct BuildTri { float lo[3], hi[3], c[3]; uint32_t id; };

void build_rec(Scene& sc, std::vector<BuildTri>& tris, uint32_t first, uint32_t count, int32_t node_id, float pad) {
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t i = first; i < first + count; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], tris[i].lo[a]); hi[a] = std::max(hi[a], tris[i].hi[a]);
            clo[a] = std::min(clo[a], tris[i].c[a]); chi[a] = std::max(chi[a], tris[i].c[a]);
        }
    {
        BNode& n = sc.nodes[node_id];
        for (int a = 0; a < 3; ++a) { n.lo[a] = lo[a] - pad; n.hi[a] = hi[a] + pad; }
        n.left = -1; n.right = -1; n.first = first; n.count = count;
    }
    if (count <= 2) return;
    const int NB = 16;
    float best_cost = 1e30f; int best_axis = -1, best_bin = -1;
    auto area = [](const float* l, const float* h) {
        float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    };
    for (int a = 0; a < 3; ++a) {
        float ext = chi[a] - clo[a];
        if (!(ext > 0.0f)) continue;
        struct Bin { float lo[3], hi[3]; uint32_t n; } bins[NB];
        for (auto& b : bins) { for (int k = 0; k < 3; ++k) { b.lo[k] = 1e30f; b.hi[k] = -1e30f; } b.n = 0; }
        float scale = NB / ext;
        for (uint32_t i = first; i < first + count; ++i) {
            int b = std::min(NB - 1, (int)((tris[i].c[a] - clo[a]) * scale));
            bins[b].n++;
            for (int k = 0; k < 3; ++k) { bins[b].lo[k] = std::min(bins[b].lo[k], tris[i].lo[k]); bins[b].hi[k] = std::max(bins[b].hi[k], tris[i].hi[k]); }
        }
        float rarea[NB]; uint32_t rcnt[NB];
        float l[3] = {1e30f, 1e30f, 1e30f}, h[3] = {-1e30f, -1e30f, -1e30f}; uint32_t cnt = 0;
        for (int b = NB - 1; b > 0; --b) {
            for (int k = 0; k < 3; ++k) { l[k] = std::min(l[k], bins[b].lo[k]); h[k] = std::max(h[k], bins[b].hi[k]); }
            cnt += bins[b].n; rcnt[b] = cnt; rarea[b] = cnt ? area(l, h) : 0.0f;
        }
        for (int k = 0; k < 3; ++k) { l[k] = 1e30f; h[k] = -1e30f; } cnt = 0;
        for (int b = 0; b < NB - 1; ++b) {
            for (int k = 0; k < 3; ++k) { l[k] = std::min(l[k], bins[b].lo[k]); h[k] = std::max(h[k], bins[b].hi[k]); }
            cnt += bins[b].n;
            if (cnt == 0 || rcnt[b + 1] == 0) continue;
            float cost = area(l, h) * cnt + rarea[b + 1] * rcnt[b + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        if (count <= 4) return;
        mid = first + count / 2;                      // all centroids coincide: split by order
    } else {
        float leaf_cost = area(lo, hi) * count;
        if (count <= 4 && leaf_cost <= best_cost + area(lo, hi) * 1.0f) return;   // SAH says stay a leaf
        float ext = chi[best_axis] - clo[best_axis], scale = NB / ext;
        auto it = std::partition(tris.begin() + first, tris.begin() + first + count, [&](const BuildTri& t) {
            int b = std::min(NB - 1, (int)((t.c[best_axis] - clo[best_axis]) * scale));
            return b <= best_bin;
        });
        mid = (uint32_t)(it - tris.begin());
        if (mid == first || mid == first + count) mid = first + count / 2;
    }
    int32_t li = (int32_t)sc.nodes.size(); sc.nodes.push_back(BNode{});
    int32_t ri = (int32_t)sc.nodes.size(); sc.nodes.push_back(BNode{});
    sc.nodes[node_id].left = li; sc.nodes[node_id].right = ri;
    build_rec(sc, tris, first, mid - first, li, pad);
    build_rec(sc, tris, mid, first + count - mid, ri, pad);
}

void build_bvh(Scene& sc) {
    uint32_t ntri = (uint32_t)(sc.indices.size() / 3);
    sc.nodes.clear(); sc.order.clear();
    if (ntri == 0) return;
    std::vector<BuildTri> tris(ntri);
    float glo[3] = {1e30f, 1e30f, 1e30f}, ghi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t i = 0; i < ntri; ++i) {
        BuildTri& t = tris[i]; t.id = i;
        for (int a = 0; a < 3; ++a) { t.lo[a] = 1e30f; t.hi[a] = -1e30f; }
        for (int k = 0; k < 3; ++k) {
            const float* p = sc.vertices[sc.indices[(size_t)i * 3 + k]].position;
            for (int a = 0; a < 3; ++a) { t.lo[a] = std::min(t.lo[a], p[a]); t.hi[a] = std::max(t.hi[a], p[a]); }
        }
        for (int a = 0; a < 3; ++a) {
            t.c[a] = 0.5f * (t.lo[a] + t.hi[a]);
            glo[a] = std::min(glo[a], t.lo[a]); ghi[a] = std::max(ghi[a], t.hi[a]);
        }
    }
    // conservative padding: Moeller-Trumbore accepts rays that miss a triangle by O(eps * distance)
    float ext = 0.0f;
    for (int a = 0; a < 3; ++a) ext = std::max(ext, std::max(std::fabs(glo[a]), std::fabs(ghi[a])));
    for (int a = 0; a < 3; ++a) ext = std::max(ext, ghi[a] - glo[a]);
    float pad = ext * 2e-5f + 1e-30f;
    sc.nodes.reserve((size_t)ntri);
    sc.nodes.push_back(BNode{});
    build_rec(sc, tris, 0, ntri, 0, pad);
    sc.order.resize(ntri);
    for (uint32_t i = 0; i < ntri; ++i) sc.order[i] = tris[i].id;
}

// ---------------------------------------------------------------------------
// trace_ray — rt_helpers.wgsl:44-80
// ---------------------------------------------------------------------------
struct Intersection {
    Hit ri;
    ptc_vertex verts[3];
    SampledMaterial material;
    bool hit_light;
    pt_v3 uvw;
    uint32_t instance_index, geometry_index, instance_custom_data;
};

Intersection trace_ray(const Scene& sc, pt_v3 ro, pt_v3 rd, bool brute, TraceCounters& c) {
    Intersection is; std::memset(&is, 0, sizeof(is));
    is.ri = brute ? closest_brute(sc, ro, rd, c) : closest_bvh(sc, ro, rd, c);
    if (!is.ri.hit) { is.instance_index = is.geometry_index = 0xFFFFFFFFu; return is; }   // material 0 tap results are unused
    uint32_t mesh_id = sc.tri_mesh[is.ri.tri];
    const ptc_mesh& mesh = sc.meshes[mesh_id];
    is.instance_index = mesh_id;                                            // pt.rs:527-542
    is.geometry_index = is.ri.tri - mesh.index_offset / 3;                  // one geometry per triangle, pt.rs:409-427
    is.instance_custom_data = mesh.material_index;
    for (int k = 0; k < 3; ++k)                                             // rt_helpers.wgsl:57-59
        is.verts[k] = sc.vertices[sc.indices[mesh.index_offset + is.geometry_index * 3 + k]];
    float w = 1.0f - is.ri.b1 - is.ri.b2;                                   // :62
    is.uvw = pt_mk3(w, is.ri.b1, is.ri.b2);
    float u = is.verts[0].uv[0] * w + is.verts[1].uv[0] * is.ri.b1 + is.verts[2].uv[0] * is.ri.b2;   // :65
    float v = is.verts[0].uv[1] * w + is.verts[1].uv[1] * is.ri.b1 + is.verts[2].uv[1] * is.ri.b2;
    is.material = sample_material(sc, sc.materials[is.instance_custom_data], u, v);   // :67
    is.hit_light = pt_dot(is.material.emissive, pt_mk3(1.0f, 1.0f, 1.0f)) > 0.0f;     // :70
    return is;
}

// ---------------------------------------------------------------------------
// BSDF — src/shaders/common/bsdf.wgsl
// ---------------------------------------------------------------------------
float get_schlick_weight(float adotb) {                                   // bsdf.wgsl:89-95
    float m = 1.0f - adotb;
    return m * m * m * m * m;
}

struct BSDFResult { pt_v3 color; float pdf; pt_v3 ray_outgoing; pt_v3 p;
                    pt_v3 shading_normal; float alpha; bool refracted; };   // the last three feed PTC_LIGHTS_MIS only

BSDFResult disney_bsdf(Rng& rng, pt_v3 p, pt_v3 ray_incoming, pt_v3 normal_in, const SampledMaterial& m) {   // bsdf.wgsl:195-231
    // bsdf_prep_data, bsdf.wgsl:39-87
    pt_v3 v = pt_neg(ray_incoming);
    pt_v3 raw_normal = normal_in;
    pt_v3 normal = pt_faceforward(raw_normal, ray_incoming, raw_normal);
    pt_v3 tangent = pt_normalize(pt_cross(normal, v));
    pt_v3 bitangent = pt_normalize(pt_cross(normal, tangent));
    const pt_v3 tsn = m.tangent_space_normal;
    normal = pt_normalize(pt_add(pt_add(pt_scale(tangent, tsn.x), pt_scale(bitangent, tsn.y)), pt_scale(normal, tsn.z)));
    float aspect = pt_sqrt(1.0f - 0.9f * m.anisotropy);
    float alpha_x = pt_max(0.0001f, pt_div(m.roughness, aspect));
    pt_v3 diffuse_color = pt_scale(m.albedo, 1.0f - m.metallic);
    float ndotv = pt_dot(normal, v);

    (void)rand_float(rng);                                                // bsdf.wgsl:203 `let u = rand_float();` (unused)

    // bsdf_sample, bsdf.wgsl:155-187
    float u1 = rand_float(rng);
    float up_down = pt_atan(alpha_x * pt_sqrt(pt_div(u1, 1.0f - u1)));
    float left_right = rand_float(rng) * 2.0f * PT_PI;
    float sin_ud, cos_ud, sin_lr, cos_lr;
    pt_sincos(up_down, &sin_ud, &cos_ud);
    pt_sincos(left_right, &sin_lr, &cos_lr);
    float local_x = sin_ud * cos_lr, local_y = cos_ud, local_z = sin_ud * sin_lr;
    pt_v3 h = pt_normalize(pt_add(pt_add(pt_scale(bitangent, local_x), pt_scale(normal, local_y)), pt_scale(tangent, local_z)));

    // bsdf_ray_outgoing_from_h, bsdf.wgsl:120-147
    float d = pt_dot(v, raw_normal);
    float eta = m.ior * (d <= 0.0f ? 1.0f : 0.0f) + pt_div(1.0f, m.ior) * (d > 0.0f ? 1.0f : 0.0f);
    pt_v3 r = pt_refract(pt_neg(v), h, eta);
    float r0 = pt_div(1.0f - eta, 1.0f + eta);
    float r0_sq = r0 * r0;
    float reflectance_fresnel = r0_sq + (1.0f - r0_sq) * get_schlick_weight(pt_dot(h, v));
    float u = rand_float(rng);
    float transmission_factor = m.transmission_weight * (1.0f - reflectance_fresnel);
    float refracted = (transmission_factor >= u ? 1.0f : 0.0f)            // step(u, transmission_factor)
                      * pt_abs(pt_sign(pt_dot(r, pt_mk3(1.0f, 1.0f, 1.0f))));
    pt_v3 refl = pt_reflect(pt_neg(v), h);
    float keep = 1.0f - refracted;
    pt_v3 outgoing = pt_mk3(r.x * refracted + refl.x * keep, r.y * refracted + refl.y * keep, r.z * refracted + refl.z * keep);

    // back in disney_bsdf
    float pdf = 1.0f;                                                     // bsdf.wgsl:180, 221 mix(1,1,refracted)
    float off = (0.01f * pt_sign(ndotv)) * (1.0f - refracted) + 0.0f * refracted;   // mix(0.01*sign(ndotv), 0.0, f32(refracted))
    BSDFResult res;
    res.color = diffuse_color;
    res.pdf = pdf;
    res.ray_outgoing = outgoing;
    res.p = pt_add(p, pt_scale(normal, off));
    res.shading_normal = normal; res.alpha = alpha_x; res.refracted = refracted != 0.0f;
    return res;
}

// ---------------------------------------------------------------------------
// integrator — src/shaders/pathtracer/pathtracer/pathtracer.wgsl
// ---------------------------------------------------------------------------
struct ScatteredRay { pt_v3 origin, direction, attenuation, normal; bool hit_prediction;
                      pt_v3 shading_normal; float alpha; bool refracted; };

ScatteredRay scatter(Rng& rng, pt_v3 ro, pt_v3 rd, const Intersection& is) {       // :76-101
    pt_v3 p = pt_fma3(rd, is.ri.t, ro);
    pt_v3 n = pt_normalize(pt_add(pt_add(pt_scale(v3(is.verts[0].normal), is.uvw.x), pt_scale(v3(is.verts[1].normal), is.uvw.y)),
                                  pt_scale(v3(is.verts[2].normal), is.uvw.z)));
    BSDFResult result = disney_bsdf(rng, p, rd, n, is.material);
    n = pt_faceforward(n, rd, n);
    ScatteredRay s;
    s.origin = result.p; s.direction = result.ray_outgoing;
    s.attenuation = pt_mk3(pt_div(result.color.x, result.pdf), pt_div(result.color.y, result.pdf), pt_div(result.color.z, result.pdf));
    s.normal = n; s.hit_prediction = true;
    s.shading_normal = result.shading_normal; s.alpha = result.alpha; s.refracted = result.refracted;
    return s;
}

pt_v3 random_point_on_triangle(Rng& rng, const ptc_triangle& tri) {               // helpers.wgsl:7-18
    float r1 = rand_float(rng), r2 = rand_float(rng);
    float s1 = pt_sqrt(r1);
    float o[3];
    for (int a = 0; a < 3; ++a)
        o[a] = tri.v0[a] * (1.0f - s1) + tri.v1[a] * s1 * (1.0f - r2) + tri.v2[a] * s1 * r2;
    return pt_mk3(o[0], o[1], o[2]);
}

ScatteredRay light_scatter(Rng& rng, pt_v3 ro, pt_v3 rd, const ptc_triangle& tri, const Intersection& is) {   // :48-73
    pt_v3 to_point = random_point_on_triangle(rng, tri);
    pt_v3 p = pt_fma3(rd, is.ri.t, ro);
    pt_v3 n = v3(is.verts[0].normal);
    pt_v3 dir = pt_normalize(pt_sub(to_point, p));
    ScatteredRay s;
    s.origin = p; s.direction = dir; s.attenuation = is.material.albedo; s.normal = n;
    s.hit_prediction = pt_dot(dir, n) > 0.0f;
    s.shading_normal = n; s.alpha = 1.0f; s.refracted = true;
    return s;
}

pt_v3 sky_color(const Scene& sc, pt_v3 rd) {                                     // sky.wgsl:1-12
    pt_v3 d = pt_normalize(rd);
    float v = pt_div(pt_acos(d.y), PT_PI);
    float u = pt_div(pt_atan2(d.z, d.x) + PT_PI, 2.0f * PT_PI);
    return sample_texture_color(sc, sc.cfg.env_texture_index, u, v);
}

struct RayColorResult { pt_v3 color; float first_t; pt_v3 first_normal, first_albedo; uint32_t first_inst, first_geom; };

RayColorResult ray_color(const Scene& sc, Rng& rng, pt_v3 ro, pt_v3 rd, uint32_t num_lights, bool brute, TraceCounters& c) {   // :149-238
    pt_v3 curr_ro = ro, curr_rd = rd;
    pt_v3 color_mask = pt_mk3(1.0f, 1.0f, 1.0f);
    pt_v3 accumulated = pt_mk3(0.0f, 0.0f, 0.0f);
    RayColorResult res; std::memset(&res, 0, sizeof(res));
    res.first_inst = res.first_geom = 0xFFFFFFFFu;
    // PTC_LIGHTS_MIS (ptcore_types.h; not in the reference): light sampling + balance heuristic, same expectation
    const bool mis = sc.cfg.light_sampling == PTC_LIGHTS_MIS && sc.has_lights;
    float pdf_prev = -1.0f;               // density with which the previous hit's scattering produced curr_rd; < 0: weight 1
    for (uint32_t i = 0; i <= sc.cfg.max_bounces; ++i) {
        Intersection is = trace_ray(sc, curr_ro, curr_rd, brute, c);
        if (i == 0) { res.first_inst = is.instance_index; res.first_geom = is.geometry_index; }
        if (is.ri.hit) {
            ScatteredRay scattered;
            if (rand_float(rng) < 1.0f - sc.cfg.send_to_light_probability) {     // :170
                scattered = scatter(rng, curr_ro, curr_rd, is);
            } else {
                const ptc_triangle& light = sc.lights[rand_int(rng) % num_lights];   // :175
                scattered = light_scatter(rng, curr_ro, curr_rd, light, is);
                // cosine_pdf / triangle_pdf (:15-46,133-134) only feed `importance`, which ray_color never
                // reads; the probe trace has no side effect and is omitted (it consumes no random numbers).
                if (!scattered.hit_prediction) scattered = scatter(rng, curr_ro, curr_rd, is);   // :183-189
            }
            pt_v3 prev_rd = curr_rd;
            const SampledMaterial& m = is.material;
            pt_v3 emitted = m.emissive;
            pt_v3 light_sample = pt_mk3(0.0f, 0.0f, 0.0f); bool light_sample_on = false;   // added AFTER the emission term below
            if (mis) {
                if (is.hit_light && pdf_prev >= 0.0f && sc.light_of_tri[is.ri.tri] != 0xFFFFFFFFu) {
                    // an emitter found by BSDF sampling, weighted against the light sample taken at the previous hit
                    const ptc_triangle& L = sc.lights[sc.light_of_tri[is.ri.tri]];
                    float inv_pl = pt_light_inv_pdf(v3(L.v0), v3(L.v1), v3(L.v2), prev_rd, is.ri.t * is.ri.t, (float)sc.lights.size());
                    emitted = pt_scale(emitted, pt_mis_weight(pdf_prev, inv_pl));
                }
                const bool opaque = m.transmission_weight == 0.0f;
                if (!is.hit_light && i < sc.cfg.max_bounces && opaque) {
                    // one shadow ray from the scattered origin to a uniform point of a uniformly chosen emissive triangle
                    uint32_t li = rand_int(rng) % (uint32_t)sc.lights.size();
                    const ptc_triangle& L = sc.lights[li];
                    float r1 = rand_float(rng), r2 = rand_float(rng);
                    float s1 = pt_sqrt(r1);
                    float w0 = 1.0f - s1, w1 = s1 * (1.0f - r2), w2 = s1 * r2;
                    pt_v3 q = pt_mk3(L.v0[0] * w0 + L.v1[0] * w1 + L.v2[0] * w2, L.v0[1] * w0 + L.v1[1] * w1 + L.v2[1] * w2,
                                     L.v0[2] * w0 + L.v1[2] * w1 + L.v2[2] * w2);
                    pt_v3 to = pt_sub(q, scattered.origin);
                    float dist2 = pt_dot(to, to), dist = pt_sqrt(dist2);
                    pt_v3 dir = pt_scale(to, pt_div(1.0f, dist));
                    float wgt = pt_mis_weight(pt_ggx_reflect_pdf(scattered.shading_normal, pt_neg(prev_rd), dir, scattered.alpha),
                                              pt_light_inv_pdf(v3(L.v0), v3(L.v1), v3(L.v2), dir, dist2, (float)sc.lights.size()));
                    // radiance of the sampled point: the emissive tap of the light triangle's material at its uv
                    uint32_t lt = sc.light_tri[li];
                    const ptc_mesh& lm = sc.meshes[sc.tri_mesh[lt]];
                    const ptc_vertex* lv[3];
                    for (int k = 0; k < 3; ++k) lv[k] = &sc.vertices[sc.indices[(size_t)lt * 3 + k]];
                    float lu = lv[0]->uv[0] * w0 + lv[1]->uv[0] * w1 + lv[2]->uv[0] * w2;
                    float lvv = lv[0]->uv[1] * w0 + lv[1]->uv[1] * w1 + lv[2]->uv[1] * w2;
                    pt_v3 le = sample_texture_color(sc, sc.materials[lm.material_index].emissive_texture_idx, lu, lvv);
                    pt_v3 carried = pt_scale(pt_mul(le, scattered.attenuation), wgt);
                    bool wanted = !scattered.refracted && dist > PT_RAY_TMIN && pt_dot(carried, pt_mk3(1.0f, 1.0f, 1.0f)) > 0.0f;
                    if (wanted) {
                        Hit blocker = closest_bvh(sc, scattered.origin, dir, c);     // occluded iff something lies in (tmin, 0.999 dist)
                        if (!(blocker.hit && blocker.t < dist * 0.999f)) { light_sample = pt_mul(carried, color_mask); light_sample_on = true; }
                    }
                }
                pdf_prev = (!scattered.refracted && opaque)
                               ? pt_ggx_reflect_pdf(scattered.shading_normal, pt_neg(prev_rd), scattered.direction, scattered.alpha) : -1.0f;
            }
            curr_ro = scattered.origin; curr_rd = scattered.direction;
            accumulated = pt_add(accumulated, pt_mul(emitted, color_mask));          // :200
            if (light_sample_on) accumulated = pt_add(accumulated, light_sample);
            color_mask = pt_mul(color_mask, scattered.attenuation);                  // :202
            if (is.hit_light) break;                                                 // :204-206
            if (i == 0) {                                                            // :208-223
                res.first_t = is.ri.t;
                pt_v3 tangent = pt_normalize(pt_cross(scattered.normal, pt_neg(prev_rd)));
                pt_v3 bitangent = pt_normalize(pt_cross(scattered.normal, tangent));
                const pt_v3 tsn = m.tangent_space_normal;
                res.first_normal = pt_normalize(pt_add(pt_add(pt_scale(tangent, tsn.x), pt_scale(bitangent, tsn.y)), pt_scale(scattered.normal, tsn.z)));
                res.first_albedo = m.albedo;
            }
        } else {
            accumulated = pt_add(accumulated, pt_mul(sky_color(sc, curr_rd), color_mask));   // :230
            break;
        }
    }
    res.color = accumulated;
    return res;
}

struct PixelOut { pt_v4 color; float first_t; pt_v3 first_normal, first_albedo; uint32_t inst, geom; };

PixelOut pixel_color(const Scene& sc, Rng& rng, const ptc_uniforms& un, uint32_t x, uint32_t y, bool brute, TraceCounters& c) {   // :240-278
    const ptc_camera& cam = un.camera;
    pt_v3 ro = v3(cam.position);
    float jx = (float)x + 0.5f * rand_float(rng);
    float jy = (float)y + 0.5f * rand_float(rng);
    pt_v3 target = pt_add(pt_add(v3(cam.first_pixel_pos), pt_scale(v3(cam.pixel_space_x), jx)), pt_scale(v3(cam.pixel_space_y), jy));
    pt_v3 rd = pt_normalize(pt_sub(target, ro));
    pt_v4 color = pt_mk4(0.0f, 0.0f, 0.0f, 1.0f);
    PixelOut out; std::memset(&out, 0, sizeof(out));
    out.inst = out.geom = 0xFFFFFFFFu;
    for (uint32_t s = 0; s < sc.cfg.samples_per_pixel; ++s) {
        RayColorResult res = ray_color(sc, rng, ro, rd, un.num_lights, brute, c);
        color = pt_mk4(color.x + res.color.x, color.y + res.color.y, color.z + res.color.z, color.w + 1.0f);
        out.first_t = res.first_t; out.first_normal = res.first_normal; out.first_albedo = res.first_albedo;
        out.inst = res.first_inst; out.geom = res.first_geom;
    }
    float n = (float)sc.cfg.samples_per_pixel;
    out.color = pt_mk4(pt_div(color.x, n), pt_div(color.y, n), pt_div(color.z, n), 1.0f);
    return out;
}

// pathtracer.wgsl:21-43 for one pixel
void main_pixel(const Scene& sc, const ptc_uniforms& un, uint32_t x, uint32_t y, bool brute, TraceCounters& c,
                float* color, float* depth, float* normal, float* albedo, uint32_t* hit_id) {
    uint32_t W = sc.cfg.width;
    Rng rng; rand_init(rng, x, y, 1u, W, un.camera.frame);
    PixelOut po = pixel_color(sc, rng, un, x, y, brute, c);
    size_t pi = (size_t)y * W + x;
    if (color) {
        float* px = &color[pi * 4];
        float prev[4];
        for (int k = 0; k < 4; ++k) prev[k] = px[k] > 0.0f ? px[k] : 0.0f;           // max(vec4(0), textureLoad(..))
        if (sc.cfg.accum_mode == PTC_ACCUM_SUM) {
            if (un.is_grabbed == 0) { px[0] = prev[0] + po.color.x; px[1] = prev[1] + po.color.y; px[2] = prev[2] + po.color.z; px[3] = prev[3] + 1.0f; }
            else { px[0] = po.color.x; px[1] = po.color.y; px[2] = po.color.z; px[3] = 1.0f; }
        } else if (un.is_grabbed == 0) {
            float a1 = prev[3] + 1.0f;
            px[0] = pt_div(prev[0] * prev[3] + po.color.x, a1);
            px[1] = pt_div(prev[1] * prev[3] + po.color.y, a1);
            px[2] = pt_div(prev[2] * prev[3] + po.color.z, a1);
            px[3] = a1;
        } else { px[0] = po.color.x; px[1] = po.color.y; px[2] = po.color.z; px[3] = 1.0f; }
    }
    if (depth) depth[pi] = po.first_t;
    if (normal) { normal[pi * 4 + 0] = po.first_normal.x; normal[pi * 4 + 1] = po.first_normal.y; normal[pi * 4 + 2] = po.first_normal.z; normal[pi * 4 + 3] = 1.0f; }
    if (albedo) { albedo[pi * 4 + 0] = po.first_albedo.x; albedo[pi * 4 + 1] = po.first_albedo.y; albedo[pi * 4 + 2] = po.first_albedo.z; albedo[pi * 4 + 3] = 1.0f; }
    if (hit_id) { hit_id[pi * 2 + 0] = po.inst; hit_id[pi * 2 + 1] = po.geom; }
}

}  // namespace

// ===========================================================================
// C entry points (ctypes)
// ===========================================================================
extern "C" {

struct orc_scene { Scene sc; };

__attribute__((visibility("default")))
orc_scene* orc_scene_create(const ptc_scene_view* view, const ptc_config* cfg) {
    if (!view || !cfg) return nullptr;
    orc_scene* o = new orc_scene();
    Scene& sc = o->sc;
    sc.cfg = *cfg;
    sc.vertices.assign(view->vertices, view->vertices + view->num_vertices);
    sc.indices.assign(view->indices, view->indices + view->num_indices);
    sc.meshes.assign(view->meshes, view->meshes + view->num_meshes);
    sc.materials.assign(view->materials, view->materials + view->num_materials);
    for (uint64_t i = 0; i < view->num_textures; ++i) {
        const ptc_texture_desc& d = view->textures[i];
        Tex t; t.w = d.width; t.h = d.height; t.fmt = d.format;
        size_t nbytes = (size_t)d.width * d.height * (d.format == PTC_TEXEL_RGBA8 ? 4 : 16);
        t.bytes.assign((const uint8_t*)d.texels, (const uint8_t*)d.texels + nbytes);
        sc.textures.push_back(std::move(t));
    }
    uint32_t ntri = (uint32_t)(sc.indices.size() / 3);
    sc.tri_mesh.assign(ntri, 0);
    for (uint32_t m = 0; m < sc.meshes.size(); ++m) {
        const ptc_mesh& me = sc.meshes[m];
        for (uint32_t t = me.index_offset / 3; t < (me.index_offset + me.index_count) / 3 && t < ntri; ++t) sc.tri_mesh[t] = m;
    }
    // light list, pt.rs:206-243 (a triangle is a light iff its material's emissive texture index != 4)
    for (uint32_t t = 0; t < ntri; ++t) {
        const ptc_mesh& me = sc.meshes[sc.tri_mesh[t]];
        if (sc.materials[me.material_index].emissive_texture_idx == 4) continue;
        const uint32_t* ix = &sc.indices[(size_t)t * 3];
        ptc_triangle tri; std::memset(&tri, 0, sizeof(tri));
        std::memcpy(tri.v0, sc.vertices[ix[0]].position, 12);
        std::memcpy(tri.v1, sc.vertices[ix[1]].position, 12);
        std::memcpy(tri.v2, sc.vertices[ix[2]].position, 12);
        std::memcpy(tri.normal, sc.vertices[ix[0]].normal, 12);
        sc.lights.push_back(tri);
        sc.light_tri.push_back(t);
    }
    sc.has_lights = !sc.lights.empty();
    sc.light_of_tri.assign(ntri, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < sc.light_tri.size(); ++i) sc.light_of_tri[sc.light_tri[i]] = i;
    if (!sc.has_lights) sc.light_tri.push_back(0);
    if (sc.lights.empty()) { ptc_triangle z; std::memset(&z, 0, sizeof(z)); sc.lights.push_back(z); }   // pt.rs:232-239
    build_bvh(sc);
    return o;
}

__attribute__((visibility("default")))
void orc_scene_destroy(orc_scene* o) { delete o; }

__attribute__((visibility("default")))
uint32_t orc_num_lights(const orc_scene* o) { return (uint32_t)o->sc.lights.size(); }

__attribute__((visibility("default")))
void orc_set_size(orc_scene* o, uint32_t w, uint32_t h) { o->sc.cfg.width = w; o->sc.cfg.height = h; }

// Render one frame (pathtracer.wgsl `main` for every pixel).  Any output
// pointer may be NULL.  counters: [0]=rays (trace_ray calls) [1]=BVH2 inner
// nodes visited [2]=triangles tested.  y0/y1 restrict the rows (sampling).
__attribute__((visibility("default")))
int orc_render_frame(const orc_scene* o, const ptc_uniforms* un_in, float* color, float* depth, float* normal,
                     float* albedo, uint32_t* hit_id, int brute, int nthreads, uint32_t y0, uint32_t y1, uint64_t counters[3]) {
    const Scene& sc = o->sc;
    ptc_uniforms un = *un_in;
    un.num_lights = (uint32_t)sc.lights.size();                                    // pt.rs:243
    uint32_t W = sc.cfg.width, H = sc.cfg.height;
    if (y1 > H) y1 = H;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    std::atomic<uint32_t> next_row(y0);
    std::vector<TraceCounters> tc((size_t)nthreads);
    auto work = [&](int tid) {
        TraceCounters c;
        for (;;) {
            uint32_t y = next_row.fetch_add(1);
            if (y >= y1) break;
            for (uint32_t x = 0; x < W; ++x) main_pixel(sc, un, x, y, brute != 0, c, color, depth, normal, albedo, hit_id);
        }
        tc[tid] = c;
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    if (counters) {
        counters[0] = counters[1] = counters[2] = 0;
        for (auto& c : tc) { counters[0] += c.rays; counters[1] += c.nodes; counters[2] += c.tris; }
    }
    return 0;
}

// closest hit for caller-supplied rays (n x 8 floats: ox,oy,oz,_,dx,dy,dz,_).
// hits: n x 4 words (t,b1,b2 f32 bits, global triangle id); ids: n x 2 (instance, geometry)
// one pixel of one frame, with the RNG state its path(s) ended in and the rays it traced: lets a test compare the number of
// random draws per path with an independent restatement (tests/test_oracle_independent.py)
__attribute__((visibility("default")))
int orc_pixel(const orc_scene* o, const ptc_uniforms* un_in, uint32_t x, uint32_t y, int brute, float color[4], uint32_t* rng_state, uint64_t* rays) {
    if (!o || !un_in || !color) return -1;
    ptc_uniforms un = *un_in;
    un.num_lights = (uint32_t)o->sc.lights.size();
    Rng rng; rand_init(rng, x, y, 1u, o->sc.cfg.width, un.camera.frame);
    TraceCounters c;
    PixelOut po = pixel_color(o->sc, rng, un, x, y, brute != 0, c);
    color[0] = po.color.x; color[1] = po.color.y; color[2] = po.color.z; color[3] = po.color.w;
    if (rng_state) *rng_state = rng.state;
    if (rays) *rays = c.rays;
    return 0;
}

__attribute__((visibility("default")))
int orc_trace_rays(const orc_scene* o, const float* rays, uint64_t n, int brute, uint32_t* hits, uint32_t* ids, uint64_t counters[3]) {
    const Scene& sc = o->sc;
    TraceCounters c;
    for (uint64_t i = 0; i < n; ++i) {
        pt_v3 ro = v3(&rays[i * 8]), rd = v3(&rays[i * 8 + 4]);
        Hit h = brute ? closest_brute(sc, ro, rd, c) : closest_bvh(sc, ro, rd, c);
        float f[3] = {h.hit ? h.t : 0.0f, h.hit ? h.b1 : 0.0f, h.hit ? h.b2 : 0.0f};
        std::memcpy(&hits[i * 4], f, 12);
        hits[i * 4 + 3] = h.hit ? h.tri : 0xFFFFFFFFu;
        if (ids) {
            if (h.hit) { uint32_t m = sc.tri_mesh[h.tri]; ids[i * 2] = m; ids[i * 2 + 1] = h.tri - sc.meshes[m].index_offset / 3; }
            else ids[i * 2] = ids[i * 2 + 1] = 0xFFFFFFFFu;
        }
    }
    if (counters) { counters[0] = c.rays; counters[1] = c.nodes; counters[2] = c.tris; }
    return 0;
}

// any-hit semantic reference: occluded iff some triangle has tmin < t < tmax_i
__attribute__((visibility("default")))
int orc_trace_rays_any(const orc_scene* o, const float* rays, uint64_t n, uint8_t* occluded) {
    const Scene& sc = o->sc;
    uint32_t ntri = (uint32_t)(sc.indices.size() / 3);
    for (uint64_t i = 0; i < n; ++i) {
        pt_v3 ro = v3(&rays[i * 8]), rd = v3(&rays[i * 8 + 4]);
        float tmax = rays[i * 8 + 7] > 0.0f ? rays[i * 8 + 7] : PT_RAY_TMAX;
        uint8_t occ = 0;
        for (uint32_t t = 0; t < ntri && !occ; ++t) {
            pt_v3 v0, e1, e2; tri_verts(sc, t, v0, e1, e2);
            float tt, b1, b2;
            if (pt_ray_tri(ro, rd, v0, e1, e2, PT_RAY_TMIN, tmax, &tt, &b1, &b2)) occ = 1;
        }
        occluded[i] = occ;
    }
    return 0;
}

// ---- display path: rayproject.wgsl:45-74 and blit.wgsl:35-40 / color.wgsl:4-27 --------------------------------------
// Restated from the WGSL line by line.  dst_of_src[i] = target pixel index of source pixel i, or 0xFFFFFFFF when the
// textureStore is out of bounds (dropped).  Collisions (several sources, one target) are a data race in the reference;
// callers resolve them (the product lets the highest source index win).
__attribute__((visibility("default")))
void orc_rayproject(const ptc_uniforms* cur, const ptc_uniforms* prev, const float* depth, uint32_t* dst_of_src) {
    const ptc_camera& c = cur->camera; const ptc_camera& q = prev->camera;
    auto v3 = [](const float* a) { return pt_mk3(a[0], a[1], a[2]); };
    for (uint32_t y = 0; y < q.height; ++y) for (uint32_t x = 0; x < q.width; ++x) {
        uint32_t i = y * q.width + x;
        pt_v3 prev_ray_origin_at_xy = pt_add(pt_add(v3(q.first_pixel_pos), pt_scale(v3(q.pixel_space_x), (float)x)), pt_scale(v3(q.pixel_space_y), (float)y));
        pt_v3 prev_ray_dir_at_xy = pt_normalize(pt_sub(prev_ray_origin_at_xy, v3(q.position)));
        float first_t = depth[i];
        pt_v3 new_center_to_old_pix = pt_sub(pt_add(prev_ray_origin_at_xy, pt_scale(prev_ray_dir_at_xy, first_t)), v3(c.position));
        float den = pt_dot(new_center_to_old_pix, pt_sub(v3(c.lookat), v3(c.position)));
        pt_v3 on_vp = pt_mk3(pt_div(new_center_to_old_pix.x, den), pt_div(new_center_to_old_pix.y, den), pt_div(new_center_to_old_pix.z, den));
        float px = pt_div(pt_dot(pt_normalize(v3(c.pixel_space_x)), on_vp), pt_length(v3(c.pixel_space_x))) + (float)c.width * 0.5f + 0.5f;
        float py = pt_div(pt_dot(pt_normalize(v3(c.pixel_space_y)), on_vp), pt_length(v3(c.pixel_space_y))) + (float)c.height * 0.5f + 0.5f;
        dst_of_src[i] = 0xFFFFFFFFu;
        if (std::isnan(px) || std::isnan(py)) continue;                      // u32(NaN): undefined in WGSL, dropped here
        float rx = std::nearbyintf(px), ry = std::nearbyintf(py);            // round(): half to even
        uint32_t ux = rx <= 0.0f ? 0u : (rx >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)rx);   // u32(): saturating
        uint32_t uy = ry <= 0.0f ? 0u : (ry >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)ry);
        if (ux >= c.width || uy >= c.height) continue;                       // textureStore out of bounds: no-op
        dst_of_src[i] = uy * c.width + ux;
    }
}

// to_khronos_pbr_neutral (color.wgsl:4-27) followed by the *Srgb surface's encode; rgba in, 4 x u8 out per pixel
__attribute__((visibility("default")))
void orc_blit(const float* rgba, uint64_t n, uint8_t* out, float* tonemapped) {
    const float F90 = 0.04f;
    for (uint64_t i = 0; i < n; ++i) {
        float r = rgba[i * 4], g = rgba[i * 4 + 1], b = rgba[i * 4 + 2], a = rgba[i * 4 + 3];
        float k_s = 0.8f - F90, k_d = 0.15f;
        float x = pt_min(pt_min(r, g), b);
        float f = (x - pt_div(x * x, 4.0f * F90)) * (x <= 2.0f * F90 ? 1.0f : 0.0f) + F90 * (x > 2.0f * F90 ? 1.0f : 0.0f);
        float p = pt_max(pt_max(r - f, g - f), b - f);
        float p_n = 1.0f - pt_div((1.0f - k_s) * (1.0f - k_s), p + 1.0f - 2.0f * k_s);
        float gg = pt_div(1.0f, k_d * (p - p_n) + 1.0f);
        float lo = p <= k_s ? 1.0f : 0.0f, hi = p > k_s ? 1.0f : 0.0f;
        float c[3] = {r, g, b};
        for (int k = 0; k < 3; ++k) {
            float v = (c[k] - f) * lo + ((c[k] - f) * pt_div(p_n, p) * gg + p_n * (1.0f - gg)) * hi;
            if (tonemapped) tonemapped[i * 3 + k] = v;
            uint32_t q;
            if (!(v > 0.0f)) q = 0; else if (v >= 1.0f) q = 255;
            else { double e = v <= 0.0031308f ? 12.92 * v : 1.055 * std::pow((double)v, 1.0 / 2.4) - 0.055; q = (uint32_t)(e * 255.0 + 0.5); }
            out[i * 4 + k] = (uint8_t)q;
        }
        out[i * 4 + 3] = !(a > 0.0f) ? 0 : (a >= 1.0f ? 255 : (uint8_t)(a * 255.0f + 0.5f));
    }
}

// ---- unit-level entry points for known-answer tests -------------------------
__attribute__((visibility("default")))
uint32_t orc_jenkins_hash(uint32_t x) { return jenkins_hash(x); }

__attribute__((visibility("default")))
void orc_rand_stream(uint32_t x, uint32_t y, uint32_t width, uint32_t frame, uint32_t n, uint32_t* state0, uint32_t* ints, float* floats) {
    Rng r; rand_init(r, x, y, 1u, width, frame);
    *state0 = r.state;
    Rng r2 = r;
    for (uint32_t i = 0; i < n; ++i) { ints[i] = rand_int(r); floats[i] = rand_float(r2); }
}

__attribute__((visibility("default")))
float orc_rand_float_of(uint32_t u) { return pt_div(pt_u32_to_f32(u), pt_u32_to_f32(0xffffffffu)); }

// camera.rs:31-79 restated with glam 0.30.3 Vec3 semantics (scalar path: dot = x*x + y*y + z*z without
// fma, normalize = v * (1/length), Vec3 / f32 = per-component division) and libm tan/sin/cos.
struct OCam {
    uint32_t width, height; float fov; float position[3], lookat[3];
    float psx[3], psy[3], first[3]; float theta_x, theta_y;
};
static void g_normalize(float* v) {
    float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float r = 1.0f / len;
    v[0] *= r; v[1] *= r; v[2] *= r;
}
static void g_cross(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - b[1] * a[2];
    o[1] = a[2] * b[0] - b[2] * a[0];
    o[2] = a[0] * b[1] - b[0] * a[1];
}
static void cam_update(OCam& c) {                                         // camera.rs:54-79
    const float PI32 = 3.14159265358979323846f;
    float viewport_height = 2.0f * std::tan(c.fov * (PI32 / 180.0f) / 2.0f);
    float viewport_width = viewport_height * ((float)c.width / (float)c.height);
    float vz[3] = {c.position[0] - c.lookat[0], c.position[1] - c.lookat[1], c.position[2] - c.lookat[2]};
    g_normalize(vz);
    float up[3] = {0.0f, 1.0f, 0.0f}, vx[3], vy[3];
    g_cross(up, vz, vx); g_normalize(vx);
    g_cross(vz, vx, vy);
    float vxv[3], vyv[3], tl[3];
    for (int a = 0; a < 3; ++a) {
        vxv[a] = vx[a] * viewport_width;
        vyv[a] = -vy[a] * viewport_height;
        c.psx[a] = vxv[a] / (float)c.width;
        c.psy[a] = vyv[a] / (float)c.height;
        tl[a] = c.position[a] - vz[a] - vxv[a] / 2.0f - vyv[a] / 2.0f;
        c.first[a] = tl[a] + 0.5f * (c.psx[a] + c.psy[a]);
    }
    c.lookat[0] = c.position[0] + std::cos(c.theta_x) * std::cos(c.theta_y);
    c.lookat[1] = c.position[1] + std::sin(c.theta_y);
    c.lookat[2] = c.position[2] + std::sin(c.theta_x) * std::cos(c.theta_y);
}
// Camera::new (camera.rs:31-52, one update with theta = 0), then theta is set and update() runs
// `extra_updates` more times (Raytracer::update calls it once per loop iteration, raytracer.rs:215-241).
__attribute__((visibility("default")))
void orc_camera(uint32_t width, uint32_t height, float fov, const float pos[3], float theta_x, float theta_y, int extra_updates, ptc_uniforms* out) {
    OCam c; std::memset(&c, 0, sizeof(c));
    c.width = width; c.height = height; c.fov = fov;
    for (int a = 0; a < 3; ++a) { c.position[a] = pos[a]; c.lookat[a] = pos[a]; }
    c.lookat[2] = pos[2] + 1.0f;
    cam_update(c);
    c.theta_x = theta_x; c.theta_y = theta_y;
    for (int i = 0; i < extra_updates; ++i) cam_update(c);
    std::memset(out, 0, sizeof(*out));                                    // uniforms.rs:30-50,64-71
    ptc_camera& u = out->camera;
    for (int a = 0; a < 3; ++a) {
        u.position[a] = c.position[a]; u.lookat[a] = c.lookat[a];
        u.pixel_space_x[a] = c.psx[a]; u.pixel_space_y[a] = c.psy[a]; u.first_pixel_pos[a] = c.first[a];
    }
    u.width = width; u.height = height; u.fov = fov; u.frame = 0;
    u.samples_per_pixel = 1; u.max_bounces = 4;                           // shader_definitions.rs:16-17
}

__attribute__((visibility("default")))
void orc_math(int fn, const float* a, const float* b, float* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = pt_sin(a[i]); break;
            case 1: out[i] = pt_cos(a[i]); break;
            case 2: out[i] = pt_atan(a[i]); break;
            case 3: out[i] = pt_acos(a[i]); break;
            case 4: out[i] = pt_atan2(a[i], b[i]); break;
            case 5: out[i] = pt_div(a[i], b[i]); break;
            case 6: out[i] = pt_sqrt(a[i]); break;
            case 7: out[i] = pt_fma(a[i], b[i], a[i]); break;
            case 8: { Rng r; std::memcpy(&r.state, &a[i], 4); out[i] = rand_float(r); break; }
            case 9: { pt_v3 v = pt_normalize(pt_mk3(a[i], b[i], a[i] * b[i])); out[i] = v.x + v.y + v.z; break; }
            case 10: { pt_v3 v = pt_refract(pt_normalize(pt_mk3(a[i], b[i], 0.5f)), pt_normalize(pt_mk3(b[i], -a[i], 1.0f)), 1.5f); out[i] = v.x + v.y + v.z; break; }
            case 11: out[i] = a[i] * b[i] + a[i]; break;
            case 12: { int i0, i1; float f; pt_bilinear_axis(a[i], 37, &i0, &i1, &f); out[i] = f + (float)i0 + 64.0f * (float)i1; break; }
            default: out[i] = 0.0f;
        }
    }
}

// sample_texture_rgba on texture `idx` at n (u,v) pairs
__attribute__((visibility("default")))
void orc_sample_texture(const orc_scene* o, uint32_t idx, const float* uv, uint64_t n, float* rgba) {
    for (uint64_t i = 0; i < n; ++i) {
        pt_v4 r = sample_texture_rgba(o->sc, idx, uv[i * 2], uv[i * 2 + 1]);
        rgba[i * 4] = r.x; rgba[i * 4 + 1] = r.y; rgba[i * 4 + 2] = r.z; rgba[i * 4 + 3] = r.w;
    }
}

// sky_color for n directions
__attribute__((visibility("default")))
void orc_sky(const orc_scene* o, const float* dirs, uint64_t n, float* rgb) {
    for (uint64_t i = 0; i < n; ++i) {
        pt_v3 c = sky_color(o->sc, v3(&dirs[i * 3]));
        rgb[i * 3] = c.x; rgb[i * 3 + 1] = c.y; rgb[i * 3 + 2] = c.z;
    }
}

__attribute__((visibility("default")))
void orc_bvh_info(const orc_scene* o, uint64_t out[4]) {
    out[0] = o->sc.nodes.size(); out[1] = o->sc.order.size();
    uint64_t inner = 0; for (auto& n : o->sc.nodes) if (n.left >= 0) inner++;
    out[2] = inner; out[3] = 0;
}

}  // extern "C"
