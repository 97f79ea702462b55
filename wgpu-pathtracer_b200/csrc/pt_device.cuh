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

// Path-queue traffic (SoA ray / throughput records, ~300 MB per 1080p frame) is written once and read once, a whole
// queue later: streaming (evict-first) accesses keep it from displacing the BVH, shading records and textures in L2.
#if defined(__CUDA_ARCH__) && !defined(PTC_NO_STREAM_HINTS)
PT_HD F4 ld4_stream(const F4* p) { float4 v = __ldcs(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
PT_HD void st4_stream(F4* p, float x, float y, float z, float w) { __stcs(reinterpret_cast<float4*>(p), make_float4(x, y, z, w)); }
#else
PT_HD F4 ld4_stream(const F4* p) { return ld4(p); }
PT_HD void st4_stream(F4* p, float x, float y, float z, float w) { p->x = x; p->y = y; p->z = z; p->w = w; }
#endif

// 32 bytes (32-byte aligned) in ONE load: sm_100 has 256-bit global loads (LDG.E.256).  For divergent lanes every
// load instruction costs one L1 wavefront per lane, and k_trace is bound by exactly that (ncu, profiles/r01e:
// l1tex__data_pipe_lsu_wavefronts at 60-80 % of peak), so a BVH node is one such load and a triangle two.
PT_HD void ld8(const F4* p, F4& a, F4& b) {
#if defined(__CUDA_ARCH__) && !defined(PTC_NO_LDG256)
    // leaf triangles are read with the L1 evict-first priority: a triangle record is rarely needed again soon, the nodes above it
    // are (+1 % on Sponza, classroom and San Miguel 4K, profiles/r02zk_ab_evict.log; evict-last on the node loads: -0.5 %)
#if !defined(PTC_TRI_NO_EVICT_FIRST)
    asm("ld.global.nc.L1::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
#else
    a = ld4(p); b = ld4(p + 1);
#endif
}
struct alignas(32) NodeQ { uint32_t w[8]; };      // device BVH2 node (layout: bvh_build.h DevNode2)
PT_HD NodeQ ld_node(const NodeQ* p) {
    NodeQ n;
#if defined(__CUDA_ARCH__) && !defined(PTC_NO_LDG256)
#if defined(PTC_NODE_EVICT_LAST)
    asm("ld.global.nc.L1::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(n.w[0]), "=r"(n.w[1]), "=r"(n.w[2]), "=r"(n.w[3]), "=r"(n.w[4]), "=r"(n.w[5]), "=r"(n.w[6]), "=r"(n.w[7]) : "l"(p));
#else
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(n.w[0]), "=r"(n.w[1]), "=r"(n.w[2]), "=r"(n.w[3]), "=r"(n.w[4]), "=r"(n.w[5]), "=r"(n.w[6]), "=r"(n.w[7]) : "l"(p));
#endif
#elif defined(__CUDA_ARCH__)
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    n.w[0] = a.x; n.w[1] = a.y; n.w[2] = a.z; n.w[3] = a.w; n.w[4] = b.x; n.w[5] = b.y; n.w[6] = b.z; n.w[7] = b.w;
#else
    memcpy(&n, p, sizeof(n));      // host vectors are not 32-byte aligned
#endif
    return n;
}
// PRMT: byte k of the result is byte (sel >> 4k) & 7 of the 8-byte pool {lo = bytes 0-3, hi = bytes 4-7}
PT_HD uint32_t pt_prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel)); return r;   // (__byte_perm would first mask sel)
#else
    const uint64_t pool = (uint64_t)lo | ((uint64_t)hi << 32);
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
    return r;
#endif
}

// ---- device scene ----------------------------------------------------------

// per-texture record (bindless table).  On the device every texture lives inside an ATLAS PAGE: one cudaArray +
// one cudaTextureObject_t (point filter, unnormalised coordinates) per page, textures shelf-packed into it, `page`
// selecting the page and (ax, ay) the texture's origin inside it.  The handful of page handles travel as kernel
// parameters (DevScene::page_obj), so the handle of every TEX instruction is warp-uniform.  One texture object per
// texture would not be: lanes of a warp shade different materials, and for a per-lane handle the compiler wraps
// each TEX in a loop over the distinct handles of the warp (R2UR ... BRA.U.ANY in SASS) — one pass per texture.
// `linear` is the same texels in linear host memory (tests/hostemu only).
struct alignas(16) TexDev {
    const void* linear;
    uint32_t w, h, fmt, page;
    uint32_t ax, ay;
};
static_assert(sizeof(TexDev) == 32, "TexDev is read as two 128-bit words");
#define PTC_MAX_TEX_PAGES 8

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
    const NodeQ* nodes;       // BVH2 device nodes, 32 bytes each
    const F4* tris;           // leaf-ordered traversal triangles (4 x F4 each, the last one unused)
    const TriShade* shade;    // global-order shading records
    const MatDev* mats;
    const TexDev* texs;
    const ptc_triangle* lights;
    const uint32_t* light_of_tri;   // global triangle id -> index into lights[] (0xFFFFFFFF: not emissive); PTC_LIGHTS_MIS only
    uint32_t num_nodes, num_tris, num_lights, env_tex;
    uint32_t num_pages, num_real_lights;   // num_real_lights = 0 when lights[] holds only the dummy entry of pt.rs:232-239
    unsigned long long page_obj[PTC_MAX_TEX_PAGES];   // cudaTextureObject_t per atlas page (device build only)
    float grid_base[3], grid_extent[3];               // quantisation grid of the node boxes (bvh_build.h QuantGrid)
};

struct PathCfg {
    uint32_t width, height, max_bounces;
    float send_to_light_probability;
    uint32_t accum_mode, is_grabbed;
    uint32_t light_sampling;      // ptc_light_sampling
    uint32_t spp, sample;         // SAMPLES_PER_PIXEL and the index of the sample this pass renders (pathtracer/pathtracer.wgsl:258-266)
};

// ---- RNG: src/shaders/common/random.wgsl:7-36 ---------------------------------
PT_HD uint32_t rng_jenkins(uint32_t x) { x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15; return x; }
PT_HD uint32_t rng_next_u32(uint32_t& state) {
    state = state * 747796405u + 2891336453u;
    uint32_t s = state;
    uint32_t word = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (word >> 22) ^ word;
}
// f32(u) / f32(0xffffffff) = rn(u) / 2^32 (random.wgsl:34-36).  Scaling by a power of two is exact (no underflow: the
// smallest non-zero value is 2^-32), so the multiplication IS the division bit for bit — minus the ~10 instructions
// of an IEEE divide, 5-7 times per hit.  (The oracle keeps the division; the parity tests compare the two.)
PT_HD float rng_next_f32(uint32_t& state) { return pt_u32_to_f32(rng_next_u32(state)) * 2.3283064365386963e-10f; }
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
// A box plane is stored as q in [0, 32767] on the scene grid: position = base + q * extent / 32768; a node word holds the two
// planes of one axis as 16-bit halves `0x8000 | q`.  ONE instruction turns a half into the float F = 1 + q / 65536 and the
// plane's ray distance is one FMA with per-ray constants:
//     t = (base + q * extent / 32768 - o) / d = F * A + C,   A = 2 extent / d,   C = (base - o) / d - A.
// Which of a box's two planes per axis the ray enters through depends only on the sign of d, so the decode's selector (a
// per-ray register) picks the entry plane and the exit plane directly: no per-node min/max per axis.
// The decode is IDP.2A (dp2a.lo: c + w.h0 * sel.b0 + w.h1 * sel.b1 with sel = (128, 0) or (0, 128) and c = 0x3F800000 -
// (0x8000 << 7), i.e. float bits 0x3F800000 + (q << 7)), an FMA-pipe instruction: the kernel's busiest unit is the ALU pipe
// (ncu: 65 % of its peak at 70 % issue-active; min/max, selects, compares and logic all live there), and the same decode
// written as PRMT sat on that pipe too.  Entry and exit distance of an axis can also come from ONE FFMA2 (f32x2, sm_100):
// {F_entry, F_exit} * {A, A} + {C_entry, C_exit} (-DPTC_FFMA2).  Measured on B200 (tools/microbench/slab.cu,
// profiles/r02k_slab.txt), cycles per two-box test per scheduler: PRMT + FFMA 46.7, IDP.2A + FFMA 42.7, IDP.2A + FFMA2 39.0 —
// but k_trace itself does not follow (profiles/r02l_ab_slab.log, r02m_ab_push.log: all three within +-0.5 %): its time goes
// to waiting for the slowest lane's node load, not to the arithmetic.  FFMA2 needs aligned register pairs and costs 28 more
// bytes of spills at 48 registers, so the default is IDP.2A + FFMA; -DPTC_SLAB_PRMT restores the PRMT form (A/B).
// Error: |fma(F, A, C) - t| <= 2^-21 (|A| + |(base - o) / d|) (roundings of 1/d, A, C and the FMA itself; F < 1.5), an
// ABSOLUTE bound that is largest for a ray nearly parallel to the axis.  Twice that bound is folded into the constant:
// the entry plane moves back, the exit plane forward, so the test is conservative with respect to pt_ray_tri (hit
// identity can never depend on it) at no extra instruction, and a near-parallel ray only loses culling along the one
// axis it is parallel to.  In space units the bound is ~1e-6 of the scene extent, 1/30 of a grid step; the builder
// rounds every plane outwards by a whole extra step.  Direction components below 2^-80 are clamped so 1/d is finite.
struct RaySlab {
    float ax, ay, az;             // A per axis
    float nx, fx, ny, fy, nz, fz; // C for the entry (near) / exit (far) plane of each axis
    uint32_t sx, sy, sz;          // decode selector: entry plane in the low half, exit plane in the high half (PRMT form: sel ^ PT_SEL_FLIP)
};
#if defined(PTC_SLAB_PRMT)
#define PT_SEL_LO 0x7104u          // PRMT bytes {0x00, w.b0, w.b1, 0x3F}: the low 16 bits of the word -> float bits 0x3F000000 + (h << 8)
#define PT_SEL_HI 0x7324u          // bytes {0x00, w.b2, w.b3, 0x3F}: the high 16 bits
#define PT_SEL_FLIP 0x0220u
#define PT_SEL_POOL 0x3F000000u
#define PT_SLAB_SCALE 1.0f         // F = 1 + q / 32768
#else
// one selector word serves both planes of an axis: dp2a.lo reads its multipliers from bytes 0-1 (the entry plane's), dp2a.hi
// from bytes 2-3 (the exit plane's) — no per-step flip of the selector
#define PT_SEL_LO 0x80000080u      // entry plane = low 16 bits of the word (b0, b1) = (128, 0); exit plane = high 16 bits (b2, b3) = (0, 128)
#define PT_SEL_HI 0x00808000u      // entry plane = high 16 bits (0, 128); exit plane = low 16 bits (128, 0)
#define PT_SEL_BIAS 0x3F400000u    // + (0x8000 | q) * 128 = 0x3F800000 + (q << 7)
#define PT_SLAB_SCALE 2.0f         // F = 1 + q / 65536
#endif

// dp2a.lo.u32.u32: c + a.h0 * b.b0 + a.h1 * b.b1 (16-bit halves of a, low two bytes of b)
PT_HD uint32_t pt_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xFFFFu) * (b & 0xFFu) + (a >> 16) * ((b >> 8) & 0xFFu);
#endif
}
// dp2a.hi.u32.u32: c + a.h0 * b.b2 + a.h1 * b.b3
PT_HD uint32_t pt_dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_hi(a, b, c);
#else
    return c + (a & 0xFFFFu) * ((b >> 16) & 0xFFu) + (a >> 16) * ((b >> 24) & 0xFFu);
#endif
}

PT_HD float pt_min3(float a, float b, float c) {
#if defined(__CUDA_ARCH__) && !defined(PTC_NO_MINMAX3)
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fminf(fminf(a, b), c);
#endif
}
PT_HD float pt_max3(float a, float b, float c) {
#if defined(__CUDA_ARCH__) && !defined(PTC_NO_MINMAX3)
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}

PT_HD void slab_axis(float o, float d, float base, float extent, float& a, float& cn, float& cf, uint32_t& sel) {
    const float kTiny = 8.271806125530277e-25f;   // 2^-80
    const float kErr = 9.5367431640625e-7f;       // 2^-20
    const float dd = pt_abs(d) > kTiny ? d : copysignf(kTiny, d);
    const float inv = pt_div(1.0f, dd);
    a = (extent * PT_SLAB_SCALE) * inv;
    const float b = (base - o) * inv;
    const float c = b - a;
    const float e = (pt_abs(a) + pt_abs(b)) * kErr;
    cn = c - e; cf = c + e;
    sel = dd > 0.0f ? PT_SEL_LO : PT_SEL_HI;
}
PT_HD RaySlab make_ray_slab(const DevScene& s, pt_v3 o, pt_v3 d) {
    RaySlab r;
    slab_axis(o.x, d.x, s.grid_base[0], s.grid_extent[0], r.ax, r.nx, r.fx, r.sx);
    slab_axis(o.y, d.y, s.grid_base[1], s.grid_extent[1], r.ay, r.ny, r.fy, r.sy);
    slab_axis(o.z, d.z, s.grid_base[2], s.grid_extent[2], r.az, r.nz, r.fz, r.sz);
    return r;
}

// A ray with a non-finite origin or direction hits nothing (this project's definition; the reference leaves it to
// the Vulkan driver).  pt_ray_tri already rejects every triangle for such a ray (NaN fails each comparison), so this
// is an early-out, not a change of result — but an essential one: fmin/fmax drop NaNs, so a NaN slab would stop
// culling altogether and one such ray would walk the whole tree (measured: a 37 ms straggler per launch).
PT_HD bool ray_is_finite(pt_v3 o, pt_v3 d) {
    float z = ((o.x * 0.0f + o.y * 0.0f) + (o.z * 0.0f + d.x * 0.0f)) + (d.y * 0.0f + d.z * 0.0f);   // 0 iff all six are finite
    return z == 0.0f;
}

// Both children of one BVH2 node against the ray: hit flags and entry distances.
// slab_pair: entry and exit distance of the ray to the two planes of one axis word.
PT_HD void slab_pair(uint32_t w, uint32_t sel, float a, float cn, float cf, float& tn, float& tf) {
#if defined(PTC_SLAB_PRMT)
    const uint32_t n = pt_prmt(w, PT_SEL_POOL, sel), f = pt_prmt(w, PT_SEL_POOL, sel ^ PT_SEL_FLIP);
#else
    const uint32_t n = pt_dp2a_lo(w, sel, PT_SEL_BIAS), f = pt_dp2a_hi(w, sel, PT_SEL_BIAS);
#endif
#if defined(__CUDA_ARCH__) && defined(PTC_FFMA2)
    const float2 t = __ffma2_rn(make_float2(u2f(n), u2f(f)), make_float2(a, a), make_float2(cn, cf));   // two IEEE fmas, one instruction
    tn = t.x; tf = t.y;
#else
    tn = pt_fma(u2f(n), a, cn); tf = pt_fma(u2f(f), a, cf);
#endif
}
PT_HD void node_test(const RaySlab& r, const NodeQ& n, float tmax, bool& h0, bool& h1, float& tn0, float& tn1) {
    float nx, fx, ny, fy, nz, fz;
    // fmin/fmax drop NaNs, so a degenerate plane simply stops constraining the interval
    slab_pair(n.w[0], r.sx, r.ax, r.nx, r.fx, nx, fx); slab_pair(n.w[1], r.sy, r.ay, r.ny, r.fy, ny, fy); slab_pair(n.w[2], r.sz, r.az, r.nz, r.fz, nz, fz);
    tn0 = pt_max3(nx, ny, fmaxf(nz, 0.0f));
    const float tf0 = pt_min3(fx, fy, fminf(fz, tmax));
    slab_pair(n.w[4], r.sx, r.ax, r.nx, r.fx, nx, fx); slab_pair(n.w[5], r.sy, r.ay, r.ny, r.fy, ny, fy); slab_pair(n.w[6], r.sz, r.az, r.nz, r.fz, nz, fz);
    tn1 = pt_max3(nx, ny, fmaxf(nz, 0.0f));
    const float tf1 = pt_min3(fx, fy, fminf(fz, tmax));
    h0 = tn0 <= tf0;
    h1 = tn1 <= tf1;
}

// All triangles of one leaf (closest hit; ties in t go to the smaller global triangle id so the result does not
// depend on traversal order).
PT_HD void leaf_closest(const DevScene& s, int node, pt_v3 o, pt_v3 d, HitRec& best, uint32_t* counters) {
    uint32_t enc = ~(uint32_t)node;
    uint32_t first = enc >> 3, cnt = (enc & 7u) + 1u;
    for (uint32_t i = 0; i < cnt; ++i) {
        const F4* tp = s.tris + (size_t)(first + i) * 4;
        F4 a, b; ld8(tp, a, b); const F4 c = ld4(tp + 2);
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
    if (s.num_nodes == 0 || !ray_is_finite(o, d)) return best;
    const RaySlab r = make_ray_slab(s, o, d);
    int stack[PTC_STACK_SIZE]; int sp = 0;
    stack[sp++] = PTC_SENTINEL;
    int node = 0;
    while (node != PTC_SENTINEL) {
        if (node >= 0) {
            const NodeQ nq = ld_node(s.nodes + node);
            if (counters) counters[0]++;
            bool h0, h1; float tn0, tn1;
            node_test(r, nq, best.t, h0, h1, tn0, tn1);
            int c0 = (int)nq.w[3], c1 = (int)nq.w[7];
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
    if (s.num_nodes == 0 || !ray_is_finite(o, d)) return false;
    const RaySlab r = make_ray_slab(s, o, d);
    int stack[PTC_STACK_SIZE]; int sp = 0;
    stack[sp++] = PTC_SENTINEL;
    int node = 0;
    while (node != PTC_SENTINEL) {
        if (node >= 0) {
            const NodeQ nq = ld_node(s.nodes + node);
            bool h0, h1; float tn0, tn1;
            node_test(r, nq, tmax, h0, h1, tn0, tn1);
            int c0 = (int)nq.w[3], c1 = (int)nq.w[7];
            if (!h0 && !h1) node = stack[--sp];
            else if (h0 && h1) { stack[sp++] = c1; node = c0; }
            else node = h0 ? c0 : c1;
        } else {
            uint32_t enc = ~(uint32_t)node;
            uint32_t first = enc >> 3, cnt = (enc & 7u) + 1u;
            for (uint32_t i = 0; i < cnt; ++i) {
                const F4* tp = s.tris + (size_t)(first + i) * 4;
                F4 a, b; ld8(tp, a, b); const F4 c = ld4(tp + 2);
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
// Device: atlas pages are point-sampled with unnormalised coordinates; the four taps of one bilinear lookup are
// issued back to back (one exposed round trip, not four) and the fp32 weights of pt_bilinear_axis applied in
// registers.  8-bit pages are read through cudaReadModeNormalizedFloat: the texture unit's u8 -> f32 conversion is
// checked once per process (pt_step.cu: unorm8_hw_exact) to equal rn(v / 255) for all 256 values — the reference's
// to_rgba32f (texture.rs:241) — and if it ever did not, 8-bit textures would be expanded to f32 pages on upload.
// shade_bounce calls sample_tex from ONE place (a loop over the lookups a path needs), so the kernel holds one inlined
// copy: eight copies made k_shade 84 KB of SASS with instruction-fetch stalls as its second largest stall reason.
PT_HD float bilerp(float a, float b, float c, float d, float fx, float fy) {
    float wx = 1.0f - fx;
    float top = pt_fma(b, fx, a * wx);
    float bot = pt_fma(d, fx, c * wx);
    return pt_fma(bot, fy, top * (1.0f - fy));
}
PT_HD pt_v4 unorm8_to_f32(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    return pt_mk4(pt_div((float)x, 255.0f), pt_div((float)y, 255.0f), pt_div((float)z, 255.0f), pt_div((float)w, 255.0f));
}
// the four texels (x0,y0) (x1,y0) (x0,y1) (x1,y1) of texture t
PT_HD void fetch_quad(const DevScene& s, const TexDev& t, int x0, int x1, int y0, int y1, pt_v4& a, pt_v4& b, pt_v4& c, pt_v4& d) {
#if defined(__CUDA_ARCH__)
    const float fx0 = (float)(t.ax + (uint32_t)x0) + 0.5f, fx1 = (float)(t.ax + (uint32_t)x1) + 0.5f;
    const float fy0 = (float)(t.ay + (uint32_t)y0) + 0.5f, fy1 = (float)(t.ay + (uint32_t)y1) + 0.5f;
    float4 pa = make_float4(0.0f, 0.0f, 0.0f, 0.0f), pb = pa, pc = pa, pd = pa;
    // Fully unrolled over the page slots so that every handle is read at a LITERAL index of the kernel-parameter
    // array: with a run-time trip count the compiler rewrote `if (t.page == p) .. page_obj[p]` into page_obj[t.page] —
    // a per-lane handle after all, and each TEX became an R2UR / BRA.U.ANY loop over the warp's distinct handles
    // (120 R2UR + 369 PLOP3 in k_shade's SASS).
    // A page no lane of the (converged part of the) warp needs is skipped by a uniform branch: predicated-off TEX
    // instructions still cost their issue slots (ncu source page: 8 pages x 10 instructions per lookup before this).
    const unsigned am = __activemask();
#pragma unroll
    for (uint32_t p = 0; p < PTC_MAX_TEX_PAGES; ++p) {
        if (p < s.num_pages) {
            const bool mine = t.page == p;
            if (__ballot_sync(am, mine) != 0u) {
                if (mine) {
                    const cudaTextureObject_t obj = (cudaTextureObject_t)s.page_obj[p];
                    pa = tex2D<float4>(obj, fx0, fy0); pb = tex2D<float4>(obj, fx1, fy0);
                    pc = tex2D<float4>(obj, fx0, fy1); pd = tex2D<float4>(obj, fx1, fy1);
                }
            }
        }
    }
    a = pt_mk4(pa.x, pa.y, pa.z, pa.w); b = pt_mk4(pb.x, pb.y, pb.z, pb.w);
    c = pt_mk4(pc.x, pc.y, pc.z, pc.w); d = pt_mk4(pd.x, pd.y, pd.z, pd.w);
#else
    (void)s;
    const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
    pt_v4 out[4];
    for (int k = 0; k < 4; ++k) {
        size_t i = (size_t)ys[k] * t.w + (size_t)xs[k];
        if (t.fmt == PTC_TEXEL_RGBA8) { const uint8_t* p = (const uint8_t*)t.linear + i * 4; out[k] = unorm8_to_f32(p[0], p[1], p[2], p[3]); }
        else { const float* p = (const float*)t.linear + i * 4; out[k] = pt_mk4(p[0], p[1], p[2], p[3]); }
    }
    a = out[0]; b = out[1]; c = out[2]; d = out[3];
#endif
}
PT_HD pt_v4 sample_tex(const DevScene& s, uint32_t idx, float u, float v) {
#if defined(__CUDA_ARCH__)
    // the 32-byte record as two 128-bit loads
    const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(s.texs + idx)), r1 = __ldg(reinterpret_cast<const uint4*>(s.texs + idx) + 1);
    TexDev t; t.linear = nullptr; t.w = r0.z; t.h = r0.w; t.fmt = r1.x; t.page = r1.y; t.ax = r1.z; t.ay = r1.w;
#else
    const TexDev& t = s.texs[idx];
#endif
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0; float fx = 0.0f, fy = 0.0f;
    const bool single = t.w == 1 && t.h == 1;          // a 1x1 texture returns its texel for every uv
    if (!single) {
        pt_bilinear_axis(u, (int)t.w, &x0, &x1, &fx);
        pt_bilinear_axis(v, (int)t.h, &y0, &y1, &fy);
    }
    pt_v4 a, b, c, d;
    fetch_quad(s, t, x0, x1, y0, y1, a, b, c, d);
    if (single) return a;
    return pt_mk4(bilerp(a.x, b.x, c.x, d.x, fx, fy), bilerp(a.y, b.y, c.y, d.y, fx, fy),
                  bilerp(a.z, b.z, c.z, d.z, fx, fy), bilerp(a.w, b.w, c.w, d.w, fx, fy));
}

// sky.wgsl:1-12: direction -> equirectangular (u, v) of texture `env_tex`
PT_HD void sky_uv(pt_v3 rd, float& u, float& v) {
    pt_v3 d = pt_normalize(rd);
    v = pt_div(pt_acos(d.y), PT_PI);
    u = pt_div(pt_atan2(d.z, d.x) + PT_PI, 2.0f * PT_PI);
}
PT_HD pt_v3 sky_radiance(const DevScene& s, pt_v3 rd) {
    float u, v; sky_uv(rd, u, v);
    pt_v4 c = sample_tex(s, s.env_tex, u, v);
    return pt_mk3(c.x, c.y, c.z);
}

// ---- one bounce of shading -----------------------------------------------------
struct PathState {
    pt_v3 o, d;          // current segment
    pt_v3 mask, acc;     // color_mask, accumulated_color (pathtracer/pathtracer.wgsl:155-156)
    uint32_t rng;
    float pdf = -1.0f;   // PTC_LIGHTS_MIS: p_b of d (< 0: none)
};

struct BounceOut {
    bool alive;                   // path continues with `next`
    PathState next;
    pt_v3 radiance;               // valid when !alive: ray_color's result
    // first-hit aux (valid when bounce == 0)
    float first_t; pt_v3 first_normal, first_albedo;
    uint32_t shadow_rays;         // PTC_LIGHTS_MIS: any-hit rays this bounce traced (0 or 1)
};

// The resolved material at a hit (rt_helpers.wgsl:17-42 minus its dead taps).
struct MatEval { pt_v3 albedo, tsn, emissive; float roughness, metallic, transmission, ior, anisotropy; };

#define SLOT_SKY SLOT_COUNT      // pseudo-slot: the environment lookup of a path that escaped

PT_HD uint32_t pt_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

// scatter + disney_bsdf: pathtracer/pathtracer.wgsl:76-101, bsdf.wgsl:39-87,120-231.
// Returns origin/direction/attenuation; n_ff receives the face-forwarded geometric-interpolated normal.
struct ScatterAux { pt_v3 ns; float alpha; bool refracted; };   // shading normal, GGX alpha, lobe taken (PTC_LIGHTS_MIS)
PT_HD void scatter_regular(uint32_t& rng, pt_v3 ro, pt_v3 rd, float t, pt_v3 n, const MatEval& m,
                           pt_v3& out_o, pt_v3& out_d, pt_v3& atten, pt_v3& n_ff, ScatterAux* aux = nullptr) {
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
    if (aux) { aux->ns = ns; aux->alpha = alpha; aux->refracted = refr != 0.0f; }
}

// One iteration of ray_color's loop body (pathtracer/pathtracer.wgsl:159-233) for a
// path whose segment has been traced to `hit`.
//
// Texture work is gathered into one loop: a path that hit needs the lookups named by its material's tex_mask
// (sample_material, rt_helpers.wgsl:17-42; 1x1 textures were folded to constants at upload), a path that escaped
// needs the environment lookup (sky.wgsl:1-12).  Each lane walks its own list, so in a warp the lanes sample
// different slots of different materials in the same pass instead of each `if (slot)` serialising the warp.
//
// shade_surface is everything that does not need the path's throughput / accumulated colour (mask, acc): the kernel
// loads those two queue entries only afterwards, so they are not live across the BSDF code (k_shade is register-bound:
// at 64 registers every spilled value is local-memory traffic, measured at ~45 % of the kernel's L2 sectors before this split).
enum SurfaceKind { SURF_MISS = 0, SURF_LIGHT = 1, SURF_LAST = 2, SURF_CONTINUE = 3 };
struct SurfaceOut {
    uint32_t kind;
    pt_v3 light;                  // sky radiance (SURF_MISS) or the hit's emissive colour
    pt_v3 atten;                  // throughput factor of this bounce
    pt_v3 next_o, next_d;         // scattered ray (SURF_CONTINUE)
    uint32_t rng;
    float first_t; pt_v3 first_normal, first_albedo;   // first-hit aux (bounce 0)
    // PTC_LIGHTS_MIS (see ptcore_types.h): the shadow ray of this hit starts at next_o
    float light_q;                // SURF_LIGHT: 1 / p_l of the emitter that was hit, seen from the ray origin
    float pdf_next;               // p_b of next_d, or -1: the next emitter hit takes weight 1 (refraction, transmissive material)
    bool nee;                     // a shadow ray is wanted
    pt_v3 nee_d; float nee_tmax;  // its direction and far end
    pt_v3 nee_c;                  // radiance it carries per unit of the path's throughput BEFORE this bounce
};

// emissive colour of triangle `tri` at barycentric weights (w0, w1, w2): sample_material's emissive tap, rt_helpers.wgsl:17-42
PT_HD pt_v3 emissive_at(const DevScene& s, uint32_t tri, float w0, float w1, float w2) {
    const TriShade& ts = s.shade[tri];
    const F4 q0 = ld4(&ts.q[0]), q1 = ld4(&ts.q[1]), q2 = ld4(&ts.q[2]), q3 = ld4(&ts.q[3]);
    const MatDev* mp = s.mats + f2u(q3.w);
    if (!(mp->tex_mask & (1u << SLOT_EMISSIVE))) return pt_mk3(mp->emissive[0], mp->emissive[1], mp->emissive[2]);
    const float u = q0.w * w0 + q2.w * w1 + q3.y * w2;
    const float v = q1.w * w0 + q3.x * w1 + q3.z * w2;
    const pt_v4 c = sample_tex(s, mp->tex[SLOT_EMISSIVE], u, v);
    return pt_mk3(c.x, c.y, c.z);
}

// DEFER_SKY (kernels only): a path that escaped is returned as SURF_MISS WITHOUT its environment lookup — k_shade queues it and
// k_sky evaluates sky_radiance for all of a frame's escaped paths in one dense pass (the lookup ran with the 2-5 lanes of a warp
// that missed while the others waited: ~1/5 of k_shade's issue slots on the later segments).
template <bool MIS, bool DEFER_SKY = false>
PT_HD SurfaceOut shade_surface(const DevScene& s, const PathCfg& cfg, pt_v3 ray_o, pt_v3 ray_d, uint32_t rng_in, const HitRec& hit, uint32_t bounce) {
    SurfaceOut out;
    out.first_t = 0.0f; out.first_normal = pt_mk3(0.0f, 0.0f, 0.0f); out.first_albedo = pt_mk3(0.0f, 0.0f, 0.0f);
    out.atten = pt_mk3(0.0f, 0.0f, 0.0f); out.next_o = ray_o; out.next_d = ray_d; out.rng = rng_in;
    out.light_q = 0.0f; out.pdf_next = -1.0f; out.nee = false; out.nee_d = ray_d; out.nee_tmax = 0.0f; out.nee_c = pt_mk3(0.0f, 0.0f, 0.0f);
    const bool mis = MIS && cfg.light_sampling == PTC_LIGHTS_MIS && s.num_real_lights != 0;
    const bool miss = hit.tri == 0xFFFFFFFFu;
    pt_v3 n = pt_mk3(0.0f, 0.0f, 0.0f);
    float u = 0.0f, v = 0.0f;
    MatEval m;
    m.albedo = m.tsn = m.emissive = pt_mk3(0.0f, 0.0f, 0.0f);
    m.roughness = m.metallic = m.transmission = m.ior = m.anisotropy = 0.0f;
    const MatDev* mp = s.mats;
    uint32_t todo;
    if (miss) {                                                            // :229-232
        if (DEFER_SKY) { out.kind = SURF_MISS; out.light = pt_mk3(0.0f, 0.0f, 0.0f); return out; }
        sky_uv(ray_d, u, v);
        todo = 1u << SLOT_SKY;
    } else {
        // rt_helpers.wgsl:57-67 through the re-packed per-triangle record
        const TriShade& ts = s.shade[hit.tri];
        const F4 q0 = ld4(&ts.q[0]), q1 = ld4(&ts.q[1]), q2 = ld4(&ts.q[2]), q3 = ld4(&ts.q[3]);
        const float w0 = 1.0f - hit.b1 - hit.b2;
        u = q0.w * w0 + q2.w * hit.b1 + q3.y * hit.b2;
        v = q1.w * w0 + q3.x * hit.b1 + q3.z * hit.b2;
        // interpolated normal now (3 live values instead of the three vertex normals across the texture loop), :85
        n = pt_normalize(pt_add(pt_add(pt_scale(pt_mk3(q0.x, q0.y, q0.z), w0), pt_scale(pt_mk3(q1.x, q1.y, q1.z), hit.b1)), pt_scale(pt_mk3(q2.x, q2.y, q2.z), hit.b2)));
        mp = s.mats + f2u(q3.w);
        m.albedo = pt_mk3(mp->albedo[0], mp->albedo[1], mp->albedo[2]);
        m.tsn = pt_mk3(mp->tsn[0], mp->tsn[1], mp->tsn[2]);
        m.emissive = pt_mk3(mp->emissive[0], mp->emissive[1], mp->emissive[2]);
        m.roughness = mp->roughness; m.metallic = mp->metallic; m.transmission = mp->transmission; m.ior = mp->ior; m.anisotropy = mp->anisotropy;
        todo = mp->tex_mask;
    }
    pt_v3 sky = pt_mk3(0.0f, 0.0f, 0.0f);
    while (todo != 0) {
        const uint32_t slot = pt_ctz(todo);
        todo &= todo - 1u;
        const pt_v4 c = sample_tex(s, slot == SLOT_SKY ? s.env_tex : mp->tex[slot], u, v);
        switch (slot) {
            case SLOT_ALBEDO: m.albedo = pt_mk3(c.x, c.y, c.z); break;
            case SLOT_NORMAL: m.tsn = pt_mk3(c.x * 2.0f - 1.0f, c.y * 2.0f - 1.0f, c.z * 2.0f - 1.0f); break;
            case SLOT_ROUGHNESS: m.roughness = c.x; break;
            case SLOT_METALLIC: m.metallic = c.x; break;
            case SLOT_EMISSIVE: m.emissive = pt_mk3(c.x, c.y, c.z); break;
            case SLOT_TRANSMISSION: m.transmission = c.x; break;
            case SLOT_IOR: m.ior = c.x; break;
            default: sky = pt_mk3(c.x, c.y, c.z); break;
        }
    }
    if (miss) { out.kind = SURF_MISS; out.light = sky; return out; }
    bool hit_light = pt_dot(m.emissive, pt_mk3(1.0f, 1.0f, 1.0f)) > 0.0f;   // rt_helpers.wgsl:70

    uint32_t rng = rng_in;
    pt_v3 so, sd, atten, sn;
    bool regular = rng_next_f32(rng) < 1.0f - cfg.send_to_light_probability;   // :170
    if (!regular) {
        // light_scatter, :48-73 (the probe trace of triangle_pdf only feeds the never-read `importance`)
        const F4 qn = ld4(&s.shade[hit.tri].q[0]);                          // vertex 0's normal (re-read: rare branch)
        const pt_v3 n0 = pt_mk3(qn.x, qn.y, qn.z);
        const ptc_triangle& L = s.lights[rng_next_u32(rng) % s.num_lights];
        float r1 = rng_next_f32(rng), r2 = rng_next_f32(rng);               // helpers.wgsl:7-18
        float s1 = pt_sqrt(r1);
        pt_v3 q = pt_mk3(L.v0[0] * (1.0f - s1) + L.v1[0] * s1 * (1.0f - r2) + L.v2[0] * s1 * r2,
                         L.v0[1] * (1.0f - s1) + L.v1[1] * s1 * (1.0f - r2) + L.v2[1] * s1 * r2,
                         L.v0[2] * (1.0f - s1) + L.v1[2] * s1 * (1.0f - r2) + L.v2[2] * s1 * r2);
        pt_v3 p = pt_fma3(ray_d, hit.t, ray_o);
        pt_v3 dir = pt_normalize(pt_sub(q, p));
        if (pt_dot(dir, n0) > 0.0f) { so = p; sd = dir; atten = m.albedo; sn = n0; }
        else regular = true;                                                // :183-189
    }
    ScatterAux aux; aux.ns = n; aux.alpha = 1.0f; aux.refracted = true;
    if (regular) scatter_regular(rng, ray_o, ray_d, hit.t, n, m, so, sd, atten, sn, mis ? &aux : nullptr);

    out.light = m.emissive; out.atten = atten;
    if (mis && hit_light) {                                                 // 1 / p_l of the emitter this BSDF-sampled ray found
        const uint32_t li = s.light_of_tri[hit.tri];
        out.light_q = -1.0f;                                                // not in the light list (cannot be light-sampled): weight 1
        if (li != 0xFFFFFFFFu) {
            const ptc_triangle& L = s.lights[li];
            out.light_q = pt_light_inv_pdf(pt_mk3(L.v0[0], L.v0[1], L.v0[2]), pt_mk3(L.v1[0], L.v1[1], L.v1[2]), pt_mk3(L.v2[0], L.v2[1], L.v2[2]),
                                           ray_d, hit.t * hit.t, (float)s.num_lights);
        }
    }
    if (mis && !hit_light && bounce < cfg.max_bounces && m.transmission == 0.0f) {
        // one light sample per opaque hit whose path continues (draw order: light index, then the two of helpers.wgsl:7-18)
        const uint32_t li = rng_next_u32(rng) % s.num_lights;
        const ptc_triangle& L = s.lights[li];
        const float r1 = rng_next_f32(rng), r2 = rng_next_f32(rng);
        const float s1 = pt_sqrt(r1);
        const float w0 = 1.0f - s1, w1 = s1 * (1.0f - r2), w2 = s1 * r2;
        const pt_v3 v0 = pt_mk3(L.v0[0], L.v0[1], L.v0[2]), v1 = pt_mk3(L.v1[0], L.v1[1], L.v1[2]), v2 = pt_mk3(L.v2[0], L.v2[1], L.v2[2]);
        const pt_v3 q = pt_mk3(v0.x * w0 + v1.x * w1 + v2.x * w2, v0.y * w0 + v1.y * w1 + v2.y * w2, v0.z * w0 + v1.z * w1 + v2.z * w2);
        const pt_v3 to = pt_sub(q, so);
        const float dist2 = pt_dot(to, to), dist = pt_sqrt(dist2);
        const pt_v3 dir = pt_scale(to, pt_div(1.0f, dist));
        const float wgt = pt_mis_weight(pt_ggx_reflect_pdf(aux.ns, pt_neg(ray_d), dir, aux.alpha), pt_light_inv_pdf(v0, v1, v2, dir, dist2, (float)s.num_lights));
        const pt_v3 le = emissive_at(s, L._pad0, w0, w1, w2);               // _pad0 carries the light's global triangle id (scene_prep)
        out.nee_c = pt_scale(pt_mul(le, atten), wgt);
        out.nee_d = dir; out.nee_tmax = dist * 0.999f;
        out.nee = !aux.refracted && dist > PT_RAY_TMIN && pt_dot(out.nee_c, pt_mk3(1.0f, 1.0f, 1.0f)) > 0.0f;
    }
    if (mis && !aux.refracted && m.transmission == 0.0f) out.pdf_next = pt_ggx_reflect_pdf(aux.ns, pt_neg(ray_d), sd, aux.alpha);
    out.rng = rng;
    if (hit_light) { out.kind = SURF_LIGHT; return out; }                   // :204-206
    if (bounce == 0) {                                                      // :208-223
        out.first_t = hit.t;
        pt_v3 tangent = pt_normalize(pt_cross(sn, pt_neg(ray_d)));
        pt_v3 bitangent = pt_normalize(pt_cross(sn, tangent));
        out.first_normal = pt_normalize(pt_add(pt_add(pt_scale(tangent, m.tsn.x), pt_scale(bitangent, m.tsn.y)), pt_scale(sn, m.tsn.z)));
        out.first_albedo = m.albedo;
    }
    if (bounce >= cfg.max_bounces) { out.kind = SURF_LAST; return out; }    // loop bound `i <= MAX_BOUNCES`, :159
    out.kind = SURF_CONTINUE; out.next_o = so; out.next_d = sd;
    return out;
}

// The part of the loop body that touches the path's running values (pathtracer/pathtracer.wgsl:200-206, 229-232):
// returns true when the path continues (mask / acc updated), else `radiance` is ray_color's result.
// pdf_prev (PTC_LIGHTS_MIS): p_b of the direction that led here, < 0 = weight 1 (camera ray, refraction); an emitter found
// by BSDF sampling is weighted against the light sample the previous hit took towards it.
PT_HD bool combine_path(const SurfaceOut& so, pt_v3& mask, pt_v3& acc, pt_v3& radiance, float pdf_prev = -1.0f) {
    if (so.kind == SURF_MISS) { radiance = pt_add(acc, pt_mul(so.light, mask)); return false; }
    pt_v3 le = so.light;
    if (so.kind == SURF_LIGHT && pdf_prev >= 0.0f && so.light_q >= 0.0f) le = pt_scale(le, pt_mis_weight(pdf_prev, so.light_q));
    acc = pt_add(acc, pt_mul(le, mask));                                   // :200
    mask = pt_mul(mask, so.atten);                                          // :202
    if (so.kind != SURF_CONTINUE) { radiance = acc; return false; }
    return true;
}

// whole loop body in one call (host build of the tests)
PT_HD BounceOut shade_bounce(const DevScene& s, const PathCfg& cfg, const PathState& ps, const HitRec& hit, uint32_t bounce) {
    const SurfaceOut so = shade_surface<true>(s, cfg, ps.o, ps.d, ps.rng, hit, bounce);
    BounceOut out;
    out.next = ps; out.shadow_rays = 0;
    out.radiance = pt_mk3(0.0f, 0.0f, 0.0f);
    out.first_t = so.first_t; out.first_normal = so.first_normal; out.first_albedo = so.first_albedo;
    pt_v3 mask = ps.mask, acc = ps.acc;
    out.alive = combine_path(so, mask, acc, out.radiance, ps.pdf);
    if (out.alive) {
        out.next.o = so.next_o; out.next.d = so.next_d; out.next.mask = mask; out.next.acc = acc; out.next.rng = so.rng; out.next.pdf = so.pdf_next;
        // PTC_LIGHTS_MIS: the shadow ray is resolved here (the wavefront queues it for the any-hit kernel instead)
        if (so.nee) {
            out.shadow_rays = 1;
            if (!trace_any_bvh2(s, so.next_o, so.nee_d, so.nee_tmax)) out.next.acc = pt_add(out.next.acc, pt_mul(so.nee_c, ps.mask));
        }
    }
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

// ---- display path: the two passes the reference runs right after the path tracer -------------------------------------

// rayproject.wgsl:45-74 — where source pixel (x, y) of the last real frame lands under the current camera, given its
// first-hit distance.  Returns false when the store would be dropped (outside the target, or a non-finite position:
// WGSL leaves u32(NaN) undefined, this project drops it).  Negative coordinates saturate to 0 as WGSL's u32() does.
PT_HD bool reproject_pixel(const ptc_camera& cam, const ptc_camera& prev, uint32_t x, uint32_t y, float first_t, uint32_t& dx, uint32_t& dy) {
    const pt_v3 pfp = pt_mk3(prev.first_pixel_pos[0], prev.first_pixel_pos[1], prev.first_pixel_pos[2]);
    const pt_v3 ppx = pt_mk3(prev.pixel_space_x[0], prev.pixel_space_x[1], prev.pixel_space_x[2]);
    const pt_v3 ppy = pt_mk3(prev.pixel_space_y[0], prev.pixel_space_y[1], prev.pixel_space_y[2]);
    const pt_v3 ppos = pt_mk3(prev.position[0], prev.position[1], prev.position[2]);
    const pt_v3 cpos = pt_mk3(cam.position[0], cam.position[1], cam.position[2]);
    const pt_v3 clook = pt_mk3(cam.lookat[0], cam.lookat[1], cam.lookat[2]);
    const pt_v3 cpx = pt_mk3(cam.pixel_space_x[0], cam.pixel_space_x[1], cam.pixel_space_x[2]);
    const pt_v3 cpy = pt_mk3(cam.pixel_space_y[0], cam.pixel_space_y[1], cam.pixel_space_y[2]);
    const pt_v3 o = pt_add(pt_add(pfp, pt_scale(ppx, (float)x)), pt_scale(ppy, (float)y));
    const pt_v3 d = pt_normalize(pt_sub(o, ppos));
    const pt_v3 p = pt_sub(pt_add(o, pt_scale(d, first_t)), cpos);
    const float depth = pt_dot(p, pt_sub(clook, cpos));
    const pt_v3 on_plane = pt_mk3(pt_div(p.x, depth), pt_div(p.y, depth), pt_div(p.z, depth));
    const float rx = pt_div(pt_dot(pt_normalize(cpx), on_plane), pt_length(cpx)) + (float)cam.width * 0.5f + 0.5f;
    const float ry = pt_div(pt_dot(pt_normalize(cpy), on_plane), pt_length(cpy)) + (float)cam.height * 0.5f + 0.5f;
    if (!(rx == rx) || !(ry == ry)) return false;
    const float qx = rintf(rx), qy = rintf(ry);               // WGSL round(): ties to even
    if (qx >= (float)cam.width || qy >= (float)cam.height) return false;
    dx = qx <= 0.0f ? 0u : (uint32_t)qx;
    dy = qy <= 0.0f ? 0u : (uint32_t)qy;
    return true;
}

// color.wgsl:4-27 to_khronos_pbr_neutral, F90 = 0.04 (shader_definitions.rs:7); both arms are evaluated and blended
// with 0/1 weights as the shader does (so a black pixel yields NaN exactly like the reference's).
// The display image is 8-bit: on the device its divisions are the fast ones (2 ulp, two instructions instead of ~12 — the blit was
// compute-bound), on the host the IEEE ones; the display test allows the one step by which a result next to a rounding boundary
// may then differ.  Nothing here feeds the accumulation.
PT_HD float pt_div_display(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return pt_div(a, b);
#endif
}
PT_HD pt_v3 tonemap_khronos_pbr_neutral(pt_v3 c) {
    const float F90 = 0.04f, k_s = 0.8f - F90, k_d = 0.15f;
    const float x = pt_min(pt_min(c.x, c.y), c.z);
    const float f = (x - pt_div_display(x * x, 4.0f * F90)) * (x <= 2.0f * F90 ? 1.0f : 0.0f) + F90 * (x > 2.0f * F90 ? 1.0f : 0.0f);
    const float p = pt_max(pt_max(c.x - f, c.y - f), c.z - f);
    const float p_n = 1.0f - pt_div_display((1.0f - k_s) * (1.0f - k_s), p + 1.0f - 2.0f * k_s);
    const float g = pt_div_display(1.0f, k_d * (p - p_n) + 1.0f);
    const float lo = p <= k_s ? 1.0f : 0.0f, hi = p > k_s ? 1.0f : 0.0f;
    const float r = pt_div_display(p_n, p), off = p_n * (1.0f - g);
    return pt_mk3((c.x - f) * lo + ((c.x - f) * r * g + off) * hi,
                  (c.y - f) * lo + ((c.y - f) * r * g + off) * hi,
                  (c.z - f) * lo + ((c.z - f) * r * g + off) * hi);
}
// what writing a linear value to the reference's *Srgb surface format does (wgpu_util.rs:235-238): clamp, sRGB OETF,
// round to 8 bits; NaN -> 0.
PT_HD uint32_t encode_srgb8(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 1.0f) return 255u;
#if defined(__CUDA_ARCH__)
    // v^(1/2.4) as exp2(log2(v) / 2.4) on the special-function unit (MUFU.LG2 / MUFU.EX2, ~2 ulp): powf's software path made the
    // blit compute-bound (31 us per 1080p frame against 9 us of memory time).  The 8-bit result can differ from the host's
    // powf by one step only where the exact value sits within ~1e-6 of a rounding boundary (the display test allows 1 LSB).
    const float e = v <= 0.0031308f ? 12.92f * v : 1.055f * exp2f(__log2f(v) * (1.0f / 2.4f)) - 0.055f;
#else
    const float e = v <= 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
#endif
    return (uint32_t)(e * 255.0f + 0.5f);
}
PT_HD uint32_t encode_unorm8(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 1.0f) return 255u;
    return (uint32_t)(v * 255.0f + 0.5f);
}

}  // namespace ptc
