// pt_device.cuh — per-thread device functions of the path-tracing wavefront:
// ray generation, BVH closest-hit / any-hit traversal, one bounce of shading.
//
// Everything here is `PT_HD` so the SAME source compiles for the GPU (nvcc,
// sm_100a) and, for the CPU-only CI in tests/hostemu, for the host with g++ —
// a test harness that lets the BVH builder, traversal and shading logic be
// checked against the oracle in a container without a GPU.  The host build is
// never part of libptcore.
//
// Algorithm restated from the reference's WGSL (file:line cited per function);
// numerics primitives from include/ptmath.h.
#pragma once
#include "../../include/ptcore_types.h"
#include "../../include/ptmath.h"
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace ptc {

struct alignas(16) F4 { float x, y, z, w; };
struct alignas(16) U4 { uint32_t x, y, z, w; };

PT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#if defined(__CUDA_ARCH__)
PT_HD F4 ld4(const F4* p) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
#else
PT_HD F4 ld4(const F4* p) { return *p; }
#endif

// ---- device scene ----------------------------------------------------------

// per-texture record (bindless table).  `obj` is a cudaTextureObject_t over a
// cudaArray (point filter, unnormalised coordinates, element read mode);
// `linear` is the same texels in linear memory (host emulation; also the
// fallback-free debug path of tests).
struct TexDev {
    unsigned long long obj;
    const void* linear;
    uint32_t w, h, fmt, _pad;
};

// One slot per material input that influences the result (specular and sheen
// taps of rt_helpers.wgsl:17-42 are dead: no consumer).  Slots whose texture is
// 1x1 are folded to constants at upload (the tap would return that texel for
// every uv).
enum MatSlot { SLOT_ALBEDO = 0, SLOT_NORMAL, SLOT_ROUGHNESS, SLOT_METALLIC, SLOT_EMISSIVE, SLOT_TRANSMISSION, SLOT_IOR, SLOT_COUNT };

struct alignas(16) MatDev {
    float albedo[3];   float roughness;
    float tsn[3];      float metallic;       // tangent-space normal, already *2-1
    float emissive[3]; float transmission;
    float ior; float anisotropy; uint32_t tex_mask; uint32_t _pad;
    uint32_t tex[8];                         // texture index per MatSlot (valid where tex_mask bit set)
};

// per-triangle shading record in GLOBAL triangle order, 64 bytes:
//  q0 = n0.xyz uv0.x | q1 = n1.xyz uv0.y | q2 = n2.xyz uv1.x | q3 = uv1.y uv2.x uv2.y material(bits)
struct alignas(16) TriShade { F4 q[4]; };

struct DevScene {
    const F4* nodes;          // BVH2 device nodes (4 x F4 each)
    const F4* tris;           // leaf-ordered traversal triangles (3 x F4 each)
    const TriShade* shade;    // global-order shading records
    const MatDev* mats;
    const TexDev* texs;
    const ptc_triangle* lights;
    uint32_t num_nodes, num_tris, num_lights, env_tex;
};

struct PathCfg {
    uint32_t width, height, max_bounces;
    float send_to_light_probability;
    uint32_t accum_mode, is_grabbed;
};

// ---- RNG: src/shaders/common/random.wgsl:7-36 ---------------------------------
PT_HD uint32_t rng_jenkins(uint32_t x) { x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15; return x; }
PT_HD uint32_t rng_next_u32(uint32_t& state) {
    state = state * 747796405u + 2891336453u;
    uint32_t s = state;
    uint32_t word = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (word >> 22) ^ word;
}
PT_HD float rng_next_f32(uint32_t& state) { return pt_div(pt_u32_to_f32(rng_next_u32(state)), 4294967296.0f); }
PT_HD uint32_t rng_seed(uint32_t x, uint32_t y, uint32_t width, uint32_t frame) {
    return rng_jenkins(((x + y * width) * 1u) ^ rng_jenkins(frame));
}

// ---- ray generation: pathtracer.wgsl:27 + pathtracer/pathtracer.wgsl:240-250 --
PT_HD void gen_primary(const ptc_camera& cam, uint32_t x, uint32_t y, uint32_t& rng, pt_v3& ro, pt_v3& rd) {
    rng = rng_seed(x, y, cam.width, cam.frame);
    float jx = (float)x + 0.5f * rng_next_f32(rng);
    float jy = (float)y + 0.5f * rng_next_f32(rng);
    ro = pt_mk3(cam.position[0], cam.position[1], cam.position[2]);
    pt_v3 fp = pt_mk3(cam.first_pixel_pos[0], cam.first_pixel_pos[1], cam.first_pixel_pos[2]);
    pt_v3 sx = pt_mk3(cam.pixel_space_x[0], cam.pixel_space_x[1], cam.pixel_space_x[2]);
    pt_v3 sy = pt_mk3(cam.pixel_space_y[0], cam.pixel_space_y[1], cam.pixel_space_y[2]);
    pt_v3 target = pt_add(pt_add(fp, pt_scale(sx, jx)), pt_scale(sy, jy));
    rd = pt_normalize(pt_sub(target, ro));
}

// ---- traversal ---------------------------------------------------------------
struct HitRec { float t, b1, b2; uint32_t tri; };   // tri = global triangle id, 0xFFFFFFFF = miss

#define PTC_STACK_SIZE 64
#define PTC_SENTINEL 0x7fffffff

// ---- slab test ------------------------------------------------------------------------------------------
// Entry/exit distances are evaluated as fma(bound, 1/d, -o/d): one FMA per plane.  Its absolute error is about
// |o/d| * 2^-24 (NOT relative to t), which for a ray nearly parallel to an axis exceeds any fixed box padding, so
// each axis carries its own error bound e = |o/d| * 2^-21 folded into the FMA constant: the plane the ray enters
// through is moved back by e and the exit plane forward by e.  The test therefore stays conservative with respect to
// pt_ray_tri (hit identity can never depend on it) at no extra instruction, and a near-parallel ray only loses
// culling along the one axis it is parallel to.  Direction components below 2^-80 are clamped so 1/d stays finite.
struct RaySlab {
    float idx, idy, idz;          // 1/d
    float lx, hx, ly, hy, lz, hz; // additive constants for the lo / hi plane of each axis (already negated)
};

PT_HD float pt_min3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fminf(fminf(a, b), c);
#endif
}
PT_HD float pt_max3(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}

PT_HD RaySlab make_ray_slab(pt_v3 o, pt_v3 d) {
    const float kTiny = 8.271806125530277e-25f;   // 2^-80
    const float kErr = 4.76837158203125e-7f;      // 2^-21
    float dx = pt_abs(d.x) > kTiny ? d.x : copysignf(kTiny, d.x);
    float dy = pt_abs(d.y) > kTiny ? d.y : copysignf(kTiny, d.y);
    float dz = pt_abs(d.z) > kTiny ? d.z : copysignf(kTiny, d.z);
    RaySlab r;
    r.idx = pt_div(1.0f, dx); r.idy = pt_div(1.0f, dy); r.idz = pt_div(1.0f, dz);
    float ox = o.x * r.idx, oy = o.y * r.idy, oz = o.z * r.idz;
    float ex = pt_abs(ox) * kErr, ey = pt_abs(oy) * kErr, ez = pt_abs(oz) * kErr;
    // positive direction: the lo plane is the entry plane (moved back: -(o/d) - e), the hi plane the exit plane
    r.lx = dx > 0.0f ? -ox - ex : -ox + ex;  r.hx = dx > 0.0f ? -ox + ex : -ox - ex;
    r.ly = dy > 0.0f ? -oy - ey : -oy + ey;  r.hy = dy > 0.0f ? -oy + ey : -oy - ey;
    r.lz = dz > 0.0f ? -oz - ez : -oz + ez;  r.hz = dz > 0.0f ? -oz + ez : -oz - ez;
    return r;
}

// Both children of one BVH2 node against the ray: hit flags and entry distances.
PT_HD void node_test(const RaySlab& r, const F4& q0, const F4& q1, const F4& q2, float tmax,
                     bool& h0, bool& h1, float& tn0, float& tn1) {
    float a0 = pt_fma(q0.x, r.idx, r.lx), b0 = pt_fma(q0.y, r.idx, r.hx);
    float c0 = pt_fma(q0.z, r.idy, r.ly), d0 = pt_fma(q0.w, r.idy, r.hy);
    float e0 = pt_fma(q2.x, r.idz, r.lz), f0 = pt_fma(q2.y, r.idz, r.hz);
    float a1 = pt_fma(q1.x, r.idx, r.lx), b1 = pt_fma(q1.y, r.idx, r.hx);
    float c1 = pt_fma(q1.z, r.idy, r.ly), d1 = pt_fma(q1.w, r.idy, r.hy);
    float e1 = pt_fma(q2.z, r.idz, r.lz), f1 = pt_fma(q2.w, r.idz, r.hz);
    // fmin/fmax drop NaNs, so a degenerate plane simply stops constraining the interval
    tn0 = pt_max3(fminf(a0, b0), fminf(c0, d0), fmaxf(fminf(e0, f0), 0.0f));
    float tf0 = pt_min3(fmaxf(a0, b0), fmaxf(c0, d0), fminf(fmaxf(e0, f0), tmax));
    tn1 = pt_max3(fminf(a1, b1), fminf(c1, d1), fmaxf(fminf(e1, f1), 0.0f));
    float tf1 = pt_min3(fmaxf(a1, b1), fmaxf(c1, d1), fminf(fmaxf(e1, f1), tmax));
    h0 = tn0 <= tf0 * 1.0000005f;
    h1 = tn1 <= tf1 * 1.0000005f;
}

// All triangles of one leaf (closest hit; ties in t go to the smaller global triangle id so the result does not
// depend on traversal order).
PT_HD void leaf_closest(const DevScene& s, int node, pt_v3 o, pt_v3 d, HitRec& best, uint32_t* counters) {
    uint32_t enc = ~(uint32_t)node;
    uint32_t first = enc >> 3, cnt = (enc & 7u) + 1u;
    for (uint32_t i = 0; i < cnt; ++i) {
        const F4* tp = s.tris + (size_t)(first + i) * 3;
        F4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
        if (counters) counters[1]++;
        float t, b1, b2;
        if (pt_ray_tri(o, d, pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z),
                       PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) {
            uint32_t id = f2u(a.w);
            if (t < best.t || (t == best.t && id < best.tri)) { best.t = t; best.b1 = b1; best.b2 = b2; best.tri = id; }
        }
    }
}

// Closest hit over the BVH2 (replaces rayQueryInitialize/Proceed/GetCommittedIntersection, rt_helpers.wgsl:49-52).
// Plain one-ray-per-call form: used by the host build of the tests and by small tools; the production kernel
// (k_trace, pt_kernels.cu) runs the same node_test / leaf_closest under a persistent, dynamically refilled warp loop.
// counters (optional): [0] nodes visited, [1] triangles tested.
PT_HD HitRec trace_closest_bvh2(const DevScene& s, pt_v3 o, pt_v3 d, uint32_t* counters) {
    HitRec best; best.t = PT_RAY_TMAX; best.b1 = 0.0f; best.b2 = 0.0f; best.tri = 0xFFFFFFFFu;
    if (s.num_nodes == 0) return best;
    const RaySlab r = make_ray_slab(o, d);
    int stack[PTC_STACK_SIZE]; int sp = 0;
    stack[sp++] = PTC_SENTINEL;
    int node = 0;
    while (node != PTC_SENTINEL) {
        if (node >= 0) {
            const F4* np = s.nodes + (size_t)node * 4;
            F4 q0 = ld4(np), q1 = ld4(np + 1), q2 = ld4(np + 2), q3 = ld4(np + 3);
            if (counters) counters[0]++;
            bool h0, h1; float tn0, tn1;
            node_test(r, q0, q1, q2, best.t, h0, h1, tn0, tn1);
            int c0 = (int)f2u(q3.x), c1 = (int)f2u(q3.y);
            if (!h0 && !h1) node = stack[--sp];
            else if (h0 && h1) {
                if (tn1 < tn0) { int t = c0; c0 = c1; c1 = t; }
                stack[sp++] = c1; node = c0;
            } else node = h0 ? c0 : c1;
        } else {
            leaf_closest(s, node, o, d, best, counters);
            node = stack[--sp];
        }
    }
    return best;
}

// Any hit in (tmin, tmax): early-out visibility query (the reference has no shadow rays in its shipped
// configuration — SEND_TO_LIGHT_PROBABILITY = 0, shader_definitions.rs:19 — this serves optional light-sampling modes).
PT_HD bool trace_any_bvh2(const DevScene& s, pt_v3 o, pt_v3 d, float tmax) {
    if (s.num_nodes == 0) return false;
    const RaySlab r = make_ray_slab(o, d);
    int stack[PTC_STACK_SIZE]; int sp = 0;
    stack[sp++] = PTC_SENTINEL;
    int node = 0;
    while (node != PTC_SENTINEL) {
        if (node >= 0) {
            const F4* np = s.nodes + (size_t)node * 4;
            F4 q0 = ld4(np), q1 = ld4(np + 1), q2 = ld4(np + 2), q3 = ld4(np + 3);
            bool h0, h1; float tn0, tn1;
            node_test(r, q0, q1, q2, tmax, h0, h1, tn0, tn1);
            int c0 = (int)f2u(q3.x), c1 = (int)f2u(q3.y);
            if (!h0 && !h1) node = stack[--sp];
            else if (h0 && h1) { stack[sp++] = c1; node = c0; }
            else node = h0 ? c0 : c1;
        } else {
            uint32_t enc = ~(uint32_t)node;
            uint32_t first = enc >> 3, cnt = (enc & 7u) + 1u;
            for (uint32_t i = 0; i < cnt; ++i) {
                const F4* tp = s.tris + (size_t)(first + i) * 3;
                F4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
                float t, b1, b2;
                if (pt_ray_tri(o, d, pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z),
                               PT_RAY_TMIN, tmax, &t, &b1, &b2)) return true;
            }
            node = stack[--sp];
        }
    }
    return false;
}

// ---- textures: helpers.wgsl:35-48, sampler texture.rs:312-322 -----------------
PT_HD pt_v4 fetch_texel(const TexDev& t, int x, int y) {
#if defined(__CUDA_ARCH__)
    if (t.fmt == PTC_TEXEL_RGBA8) {
        uchar4 p = tex2D<uchar4>((cudaTextureObject_t)t.obj, (float)x + 0.5f, (float)y + 0.5f);
        return pt_mk4(pt_div((float)p.x, 255.0f), pt_div((float)p.y, 255.0f), pt_div((float)p.z, 255.0f), pt_div((float)p.w, 255.0f));
    }
    float4 p = tex2D<float4>((cudaTextureObject_t)t.obj, (float)x + 0.5f, (float)y + 0.5f);
    return pt_mk4(p.x, p.y, p.z, p.w);
#else
    size_t i = (size_t)y * t.w + (size_t)x;
    if (t.fmt == PTC_TEXEL_RGBA8) {
        const uint8_t* p = (const uint8_t*)t.linear + i * 4;
        return pt_mk4(pt_div((float)p[0], 255.0f), pt_div((float)p[1], 255.0f), pt_div((float)p[2], 255.0f), pt_div((float)p[3], 255.0f));
    }
    const float* p = (const float*)t.linear + i * 4;
    return pt_mk4(p[0], p[1], p[2], p[3]);
#endif
}
PT_HD float bilerp(float a, float b, float c, float d, float fx, float fy) {
    float wx = 1.0f - fx;
    float top = pt_fma(b, fx, a * wx);
    float bot = pt_fma(d, fx, c * wx);
    return pt_fma(bot, fy, top * (1.0f - fy));
}
PT_HD pt_v4 sample_tex(const TexDev& t, float u, float v) {
    if (t.w == 1 && t.h == 1) return fetch_texel(t, 0, 0);
    int x0, x1, y0, y1; float fx, fy;
    pt_bilinear_axis(u, (int)t.w, &x0, &x1, &fx);
    pt_bilinear_axis(v, (int)t.h, &y0, &y1, &fy);
    pt_v4 a = fetch_texel(t, x0, y0), b = fetch_texel(t, x1, y0), c = fetch_texel(t, x0, y1), d = fetch_texel(t, x1, y1);
    return pt_mk4(bilerp(a.x, b.x, c.x, d.x, fx, fy), bilerp(a.y, b.y, c.y, d.y, fx, fy),
                  bilerp(a.z, b.z, c.z, d.z, fx, fy), bilerp(a.w, b.w, c.w, d.w, fx, fy));
}

// sky.wgsl:1-12
PT_HD pt_v3 sky_radiance(const DevScene& s, pt_v3 rd) {
    pt_v3 d = pt_normalize(rd);
    float v = pt_div(pt_acos(d.y), PT_PI);
    float u = pt_div(pt_atan2(d.z, d.x) + PT_PI, 2.0f * PT_PI);
    pt_v4 c = sample_tex(s.texs[s.env_tex], u, v);
    return pt_mk3(c.x, c.y, c.z);
}

// ---- one bounce of shading -----------------------------------------------------
struct PathState {
    pt_v3 o, d;          // current segment
    pt_v3 mask, acc;     // color_mask, accumulated_color (pathtracer/pathtracer.wgsl:155-156)
    uint32_t rng;
};

struct BounceOut {
    bool alive;                   // path continues with `next`
    PathState next;
    pt_v3 radiance;               // valid when !alive: ray_color's result
    // first-hit aux (valid when bounce == 0)
    float first_t; pt_v3 first_normal, first_albedo;
};

// The resolved material at a hit (rt_helpers.wgsl:17-42 minus its dead taps).
struct MatEval { pt_v3 albedo, tsn, emissive; float roughness, metallic, transmission, ior, anisotropy; };

PT_HD MatEval eval_material(const DevScene& s, uint32_t mat_id, float u, float v) {
    const MatDev& m = s.mats[mat_id];
    MatEval e;
    e.albedo = pt_mk3(m.albedo[0], m.albedo[1], m.albedo[2]);
    e.tsn = pt_mk3(m.tsn[0], m.tsn[1], m.tsn[2]);
    e.emissive = pt_mk3(m.emissive[0], m.emissive[1], m.emissive[2]);
    e.roughness = m.roughness; e.metallic = m.metallic; e.transmission = m.transmission; e.ior = m.ior; e.anisotropy = m.anisotropy;
    uint32_t mask = m.tex_mask;
    if (mask) {
        if (mask & (1u << SLOT_ALBEDO)) { pt_v4 c = sample_tex(s.texs[m.tex[SLOT_ALBEDO]], u, v); e.albedo = pt_mk3(c.x, c.y, c.z); }
        if (mask & (1u << SLOT_NORMAL)) {
            pt_v4 c = sample_tex(s.texs[m.tex[SLOT_NORMAL]], u, v);
            e.tsn = pt_mk3(c.x * 2.0f - 1.0f, c.y * 2.0f - 1.0f, c.z * 2.0f - 1.0f);
        }
        if (mask & (1u << SLOT_ROUGHNESS)) e.roughness = sample_tex(s.texs[m.tex[SLOT_ROUGHNESS]], u, v).x;
        if (mask & (1u << SLOT_METALLIC)) e.metallic = sample_tex(s.texs[m.tex[SLOT_METALLIC]], u, v).x;
        if (mask & (1u << SLOT_EMISSIVE)) { pt_v4 c = sample_tex(s.texs[m.tex[SLOT_EMISSIVE]], u, v); e.emissive = pt_mk3(c.x, c.y, c.z); }
        if (mask & (1u << SLOT_TRANSMISSION)) e.transmission = sample_tex(s.texs[m.tex[SLOT_TRANSMISSION]], u, v).x;
        if (mask & (1u << SLOT_IOR)) e.ior = sample_tex(s.texs[m.tex[SLOT_IOR]], u, v).x;
    }
    return e;
}

// scatter + disney_bsdf: pathtracer/pathtracer.wgsl:76-101, bsdf.wgsl:39-87,120-231.
// Returns origin/direction/attenuation; n_ff receives the face-forwarded geometric-interpolated normal.
PT_HD void scatter_regular(uint32_t& rng, pt_v3 ro, pt_v3 rd, float t, pt_v3 n, const MatEval& m,
                           pt_v3& out_o, pt_v3& out_d, pt_v3& atten, pt_v3& n_ff) {
    pt_v3 p = pt_fma3(rd, t, ro);
    pt_v3 v = pt_neg(rd);
    bool front = pt_dot(rd, n) < 0.0f;
    pt_v3 nf = front ? n : pt_neg(n);                       // faceForward(n, rd, n)
    pt_v3 tangent = pt_normalize(pt_cross(nf, v));
    pt_v3 bitangent = pt_normalize(pt_cross(nf, tangent));
    pt_v3 ns = pt_normalize(pt_add(pt_add(pt_scale(tangent, m.tsn.x), pt_scale(bitangent, m.tsn.y)), pt_scale(nf, m.tsn.z)));
    float aspect = pt_sqrt(1.0f - 0.9f * m.anisotropy);
    float alpha = pt_max(0.0001f, pt_div(m.roughness, aspect));
    float ndotv = pt_dot(ns, v);

    (void)rng_next_f32(rng);                                // bsdf.wgsl:203, drawn and unused
    float u1 = rng_next_f32(rng);                           // bsdf.wgsl:158
    float theta = pt_atan(alpha * pt_sqrt(pt_div(u1, 1.0f - u1)));
    float phi = rng_next_f32(rng) * 2.0f * PT_PI;           // bsdf.wgsl:162
    float st, ct, sp, cp;
    pt_sincos(theta, &st, &ct);
    pt_sincos(phi, &sp, &cp);
    pt_v3 h = pt_normalize(pt_add(pt_add(pt_scale(bitangent, st * cp), pt_scale(ns, ct)), pt_scale(tangent, st * sp)));

    float dvn = pt_dot(v, n);                               // raw normal, bsdf.wgsl:124
    float eta = m.ior * (dvn <= 0.0f ? 1.0f : 0.0f) + pt_div(1.0f, m.ior) * (dvn > 0.0f ? 1.0f : 0.0f);
    pt_v3 r = pt_refract(rd, h, eta);                       // refract(-v, h, eta), -v == rd
    float r0 = pt_div(1.0f - eta, 1.0f + eta);
    float r0sq = r0 * r0;
    float w = 1.0f - pt_dot(h, v);
    float fres = r0sq + (1.0f - r0sq) * (w * w * w * w * w);
    float u3 = rng_next_f32(rng);                           // bsdf.wgsl:133
    float tf = m.transmission * (1.0f - fres);
    float refr = (tf >= u3 ? 1.0f : 0.0f) * pt_abs(pt_sign(pt_dot(r, pt_mk3(1.0f, 1.0f, 1.0f))));
    pt_v3 rl = pt_reflect(rd, h);
    float keep = 1.0f - refr;
    out_d = pt_mk3(r.x * refr + rl.x * keep, r.y * refr + rl.y * keep, r.z * refr + rl.z * keep);
    float off = (0.01f * pt_sign(ndotv)) * (1.0f - refr) + 0.0f * refr;
    out_o = pt_add(p, pt_scale(ns, off));
    pt_v3 col = pt_scale(m.albedo, 1.0f - m.metallic);
    atten = pt_mk3(pt_div(col.x, 1.0f), pt_div(col.y, 1.0f), pt_div(col.z, 1.0f));   // color / pdf, pdf == 1
    n_ff = nf;
}

// One iteration of ray_color's loop body (pathtracer/pathtracer.wgsl:159-233) for a
// path whose segment has been traced to `hit`.
PT_HD BounceOut shade_bounce(const DevScene& s, const PathCfg& cfg, const PathState& ps, const HitRec& hit, uint32_t bounce) {
    BounceOut out;
    out.alive = false;
    out.first_t = 0.0f; out.first_normal = pt_mk3(0.0f, 0.0f, 0.0f); out.first_albedo = pt_mk3(0.0f, 0.0f, 0.0f);
    out.next = ps;
    if (hit.tri == 0xFFFFFFFFu) {                                          // :229-232
        out.radiance = pt_add(ps.acc, pt_mul(sky_radiance(s, ps.d), ps.mask));
        return out;
    }
    // rt_helpers.wgsl:57-67 through the re-packed per-triangle record
    const TriShade& ts = s.shade[hit.tri];
    F4 q0 = ld4(&ts.q[0]), q1 = ld4(&ts.q[1]), q2 = ld4(&ts.q[2]), q3 = ld4(&ts.q[3]);
    float w0 = 1.0f - hit.b1 - hit.b2;
    float u = q0.w * w0 + q2.w * hit.b1 + q3.y * hit.b2;
    float v = q1.w * w0 + q3.x * hit.b1 + q3.z * hit.b2;
    uint32_t mat_id = f2u(q3.w);
    MatEval m = eval_material(s, mat_id, u, v);
    bool hit_light = pt_dot(m.emissive, pt_mk3(1.0f, 1.0f, 1.0f)) > 0.0f;   // rt_helpers.wgsl:70

    uint32_t rng = ps.rng;
    pt_v3 n0 = pt_mk3(q0.x, q0.y, q0.z), n1 = pt_mk3(q1.x, q1.y, q1.z), n2 = pt_mk3(q2.x, q2.y, q2.z);
    pt_v3 n = pt_normalize(pt_add(pt_add(pt_scale(n0, w0), pt_scale(n1, hit.b1)), pt_scale(n2, hit.b2)));   // :85
    pt_v3 so, sd, atten, sn;
    bool regular = rng_next_f32(rng) < 1.0f - cfg.send_to_light_probability;   // :170
    if (!regular) {
        // light_scatter, :48-73 (the probe trace of triangle_pdf only feeds the never-read `importance`)
        const ptc_triangle& L = s.lights[rng_next_u32(rng) % s.num_lights];
        float r1 = rng_next_f32(rng), r2 = rng_next_f32(rng);               // helpers.wgsl:7-18
        float s1 = pt_sqrt(r1);
        pt_v3 q = pt_mk3(L.v0[0] * (1.0f - s1) + L.v1[0] * s1 * (1.0f - r2) + L.v2[0] * s1 * r2,
                         L.v0[1] * (1.0f - s1) + L.v1[1] * s1 * (1.0f - r2) + L.v2[1] * s1 * r2,
                         L.v0[2] * (1.0f - s1) + L.v1[2] * s1 * (1.0f - r2) + L.v2[2] * s1 * r2);
        pt_v3 p = pt_fma3(ps.d, hit.t, ps.o);
        pt_v3 dir = pt_normalize(pt_sub(q, p));
        if (pt_dot(dir, n0) > 0.0f) { so = p; sd = dir; atten = m.albedo; sn = n0; }
        else regular = true;                                                // :183-189
    }
    if (regular) scatter_regular(rng, ps.o, ps.d, hit.t, n, m, so, sd, atten, sn);

    pt_v3 acc = pt_add(ps.acc, pt_mul(m.emissive, ps.mask));               // :200
    pt_v3 mask = pt_mul(ps.mask, atten);                                    // :202
    if (hit_light) { out.radiance = acc; return out; }                      // :204-206
    if (bounce == 0) {                                                      // :208-223
        out.first_t = hit.t;
        pt_v3 tangent = pt_normalize(pt_cross(sn, pt_neg(ps.d)));
        pt_v3 bitangent = pt_normalize(pt_cross(sn, tangent));
        out.first_normal = pt_normalize(pt_add(pt_add(pt_scale(tangent, m.tsn.x), pt_scale(bitangent, m.tsn.y)), pt_scale(sn, m.tsn.z)));
        out.first_albedo = m.albedo;
    }
    if (bounce >= cfg.max_bounces) { out.radiance = acc; return out; }      // loop bound `i <= MAX_BOUNCES`, :159
    out.alive = true;
    out.next.o = so; out.next.d = sd; out.next.mask = mask; out.next.acc = acc; out.next.rng = rng;
    return out;
}

// pathtracer.wgsl:32-42 (+ the associative SUM mode used for frame-sharded multi-GPU)
PT_HD void accumulate_pixel(F4* px, pt_v3 c, uint32_t accum_mode, uint32_t is_grabbed) {
    F4 prev = *px;
    float pr = prev.x > 0.0f ? prev.x : 0.0f, pg = prev.y > 0.0f ? prev.y : 0.0f;
    float pb = prev.z > 0.0f ? prev.z : 0.0f, pa = prev.w > 0.0f ? prev.w : 0.0f;   // max(vec4(0), textureLoad)
    F4 o;
    if (is_grabbed != 0) { o.x = c.x; o.y = c.y; o.z = c.z; o.w = 1.0f; }
    else if (accum_mode == PTC_ACCUM_SUM) { o.x = pr + c.x; o.y = pg + c.y; o.z = pb + c.z; o.w = pa + 1.0f; }
    else {
        float a1 = pa + 1.0f;
        o.x = pt_div(pr * pa + c.x, a1); o.y = pt_div(pg * pa + c.y, a1); o.z = pt_div(pb * pa + c.z, a1); o.w = a1;
    }
    *px = o;
}

}  // namespace ptc
