/* ptmath.h — deterministic fp32 numerics primitives, identical on host (g++
 * -ffp-contract=off -mfma) and device (nvcc -fmad=false, IEEE div/sqrt).
 *
 * WHY: the reference's shader arithmetic (src/shaders/common/bsdf.wgsl,
 * sky.wgsl, pathtracer/pathtracer.wgsl) runs through a Vulkan driver whose
 * sin/cos/atan/acos/normalize ULPs and FMA contraction are unspecified.  A
 * Monte-Carlo path is chaotic in those ULPs, so "radiance within 1e-3 relative
 * RMSE at matched spp" is only testable if the CUDA kernels and the CPU oracle
 * evaluate bit-identical primitives.  This header therefore contains ONLY
 * primitives (vector algebra with a fixed FMA pattern, transcendental
 * functions built from + * fma / sqrt, the WGSL builtins, and this project's
 * ray/triangle test, which the reference delegates to the Vulkan driver).  The
 * reference's *algorithm* (RNG, camera, material sampling, BSDF, integrator,
 * accumulation) is restated twice, independently: in oracle/ (CPU) and in
 * wgpu-pathtracer_b200/csrc (CUDA).
 *
 * Rules: no expression here may rely on compiler contraction; every fused
 * operation is an explicit pt_fma.  Both compilers must be told not to
 * contract (see Makefiles).
 */
#ifndef PTMATH_H
#define PTMATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD static inline
#endif

/* ---- scalar primitives ------------------------------------------------- */
PT_HD float pt_fma(float a, float b, float c) { return fmaf(a, b, c); }
PT_HD float pt_sqrt(float a) { return sqrtf(a); }          /* IEEE rn on both sides */
PT_HD float pt_div(float a, float b) { return a / b; }     /* IEEE rn on both sides */
PT_HD float pt_min(float a, float b) { return a < b ? a : b; }
PT_HD float pt_max(float a, float b) { return a > b ? a : b; }  /* returns b when a is NaN */
PT_HD float pt_abs(float a) { return fabsf(a); }
PT_HD float pt_floor(float a) { return floorf(a); }
/* WGSL sign(): 1, -1, or 0 for +-0 (NaN -> 0 here) */
PT_HD float pt_sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
/* WGSL f32(u32): round to nearest even */
PT_HD float pt_u32_to_f32(uint32_t u) { return (float)u; }

#define PT_PI      3.14159265359f   /* helpers.wgsl:4 `const PI = 3.14159265359` -> f32 0x40490FDB */
#define PT_EPSILON 0.0001f          /* helpers.wgsl:5 */

/* ---- vectors ------------------------------------------------------------ */
typedef struct pt_v3 { float x, y, z; } pt_v3;
typedef struct pt_v4 { float x, y, z, w; } pt_v4;

PT_HD pt_v3 pt_mk3(float x, float y, float z) { pt_v3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD pt_v4 pt_mk4(float x, float y, float z, float w) { pt_v4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PT_HD pt_v3 pt_add(pt_v3 a, pt_v3 b) { return pt_mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD pt_v3 pt_sub(pt_v3 a, pt_v3 b) { return pt_mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD pt_v3 pt_mul(pt_v3 a, pt_v3 b) { return pt_mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD pt_v3 pt_scale(pt_v3 a, float s) { return pt_mk3(a.x * s, a.y * s, a.z * s); }
PT_HD pt_v3 pt_neg(pt_v3 a) { return pt_mk3(-a.x, -a.y, -a.z); }
/* WGSL fma(a, vec3(t), c) */
PT_HD pt_v3 pt_fma3(pt_v3 a, float t, pt_v3 c) { return pt_mk3(pt_fma(a.x, t, c.x), pt_fma(a.y, t, c.y), pt_fma(a.z, t, c.z)); }
/* dot: fixed FMA chain x, then y, then z */
PT_HD float pt_dot(pt_v3 a, pt_v3 b) { return pt_fma(a.z, b.z, pt_fma(a.y, b.y, a.x * b.x)); }
/* cross: one rounded product, one fma per component */
PT_HD pt_v3 pt_cross(pt_v3 a, pt_v3 b) {
    return pt_mk3(pt_fma(a.y, b.z, -(a.z * b.y)),
                  pt_fma(a.z, b.x, -(a.x * b.z)),
                  pt_fma(a.x, b.y, -(a.y * b.x)));
}
PT_HD float pt_length(pt_v3 a) { return pt_sqrt(pt_dot(a, a)); }
/* WGSL normalize(): v / length(v); evaluated as v * (1/len). 0-vector -> NaN as on a GPU. */
PT_HD pt_v3 pt_normalize(pt_v3 a) { float r = pt_div(1.0f, pt_length(a)); return pt_scale(a, r); }
/* WGSL faceForward(e1,e2,e3) = select(-e1, e1, dot(e2,e3) < 0) */
PT_HD pt_v3 pt_faceforward(pt_v3 e1, pt_v3 e2, pt_v3 e3) { return pt_dot(e2, e3) < 0.0f ? e1 : pt_neg(e1); }
/* WGSL reflect(e1,e2) = e1 - 2*dot(e2,e1)*e2 */
PT_HD pt_v3 pt_reflect(pt_v3 e1, pt_v3 e2) {
    float k = 2.0f * pt_dot(e2, e1);
    return pt_mk3(e1.x - k * e2.x, e1.y - k * e2.y, e1.z - k * e2.z);
}
/* WGSL refract(e1,e2,e3): k = 1 - e3^2 (1 - dot(e2,e1)^2); k<0 -> 0; else e3*e1 - (e3*dot(e2,e1) + sqrt(k))*e2 */
PT_HD pt_v3 pt_refract(pt_v3 e1, pt_v3 e2, float e3) {
    float d = pt_dot(e2, e1);
    float k = 1.0f - (e3 * e3) * (1.0f - d * d);
    if (k < 0.0f) return pt_mk3(0.0f, 0.0f, 0.0f);
    float m = e3 * d + pt_sqrt(k);
    return pt_mk3(e3 * e1.x - m * e2.x, e3 * e1.y - m * e2.y, e3 * e1.z - m * e2.z);
}

/* ---- transcendental functions (polynomials after Cephes single precision;
 *      every step is + * fma so host and device agree bit for bit) -------- */

/* sin and cos together.  Cody-Waite reduction by pi/2 in three parts; good to
 * ~1 ulp for |x| < 1e4, which covers the shader's arguments ([0, 2*pi]). */
PT_HD void pt_sincos(float x, float* s_out, float* c_out) {
    float j = rintf(x * 0.636619772f);             /* nearest multiple of pi/2, ties to even on both sides */
    float y = pt_fma(-j, 1.5707962513e+00f, x);
    y = pt_fma(-j, 7.5497894159e-08f, y);
    y = pt_fma(-j, 5.3903029534e-15f, y);
    int q = (int)j;
    float z = y * y;
    /* sin on [-pi/4, pi/4] */
    float sp = pt_fma(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = pt_fma(z, sp, -1.6666654611e-1f);
    float s = pt_fma(y * z, sp, y);
    /* cos on [-pi/4, pi/4] */
    float cp = pt_fma(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = pt_fma(z, cp, 4.166664568298827e-2f);
    float c = pt_fma(z * z, cp, pt_fma(z, -0.5f, 1.0f));
    float rs, rc;
    switch (q & 3) {
        case 0:  rs =  s; rc =  c; break;
        case 1:  rs =  c; rc = -s; break;
        case 2:  rs = -s; rc = -c; break;
        default: rs = -c; rc =  s; break;
    }
    *s_out = rs; *c_out = rc;
}
PT_HD float pt_sin(float x) { float s, c; pt_sincos(x, &s, &c); return s; }
PT_HD float pt_cos(float x) { float s, c; pt_sincos(x, &s, &c); return c; }

PT_HD float pt_atan(float xin) {
    float x = pt_abs(xin);
    float y0;
    if (x > 2.414213562373095f) { y0 = 1.5707963267948966f; x = -pt_div(1.0f, x); }
    else if (x > 0.4142135623730950f) { y0 = 0.7853981633974483f; x = pt_div(x - 1.0f, x + 1.0f); }
    else y0 = 0.0f;
    float z = x * x;
    float p = pt_fma(z, 8.05374449538e-2f, -1.38776856032e-1f);
    p = pt_fma(z, p, 1.99777106478e-1f);
    p = pt_fma(z, p, -3.33329491539e-1f);
    float y = y0 + pt_fma(p * z, x, x);
    return xin < 0.0f ? -y : y;
}

/* asin on [0, 1] (argument already non-negative) */
PT_HD float pt_asin_pos(float a) {
    float x, z; int big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); x = pt_sqrt(z); }
    else { x = a; z = x * x; }
    float p = pt_fma(z, 4.2163199048e-2f, 2.4181311049e-2f);
    p = pt_fma(z, p, 4.5470025998e-2f);
    p = pt_fma(z, p, 7.4953002686e-2f);
    p = pt_fma(z, p, 1.6666752422e-1f);
    float r = pt_fma(p * z, x, x);
    if (big) r = 1.5707963267948966f - (r + r);
    return r;
}
/* acos with the argument clamped to [-1, 1] (a normalised direction can carry
 * |y| = 1 + 1ulp; WGSL leaves that undefined, this project defines it). */
PT_HD float pt_acos(float x) {
    if (x > 1.0f) x = 1.0f;
    if (x < -1.0f) x = -1.0f;
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * pt_asin_pos(pt_sqrt(0.5f * (1.0f + x)));
    if (x > 0.5f)  return 2.0f * pt_asin_pos(pt_sqrt(0.5f * (1.0f - x)));
    float a = pt_asin_pos(pt_abs(x));
    return 1.5707963267948966f - (x < 0.0f ? -a : a);
}
PT_HD float pt_atan2(float y, float x) {
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float a = pt_atan(pt_div(y, x));
    if (x < 0.0f) return (y < 0.0f) ? a - 3.14159265358979323846f : a + 3.14159265358979323846f;
    return a;
}

/* ---- ray / triangle test -------------------------------------------------
 * The reference has none: intersection lives in the Vulkan driver behind
 * rayQueryProceed (rt_helpers.wgsl:49-52).  This is the project's definition,
 * shared by oracle and kernels so hit identity and (t, b1, b2) are bit-equal:
 * Moeller-Trumbore on (v0, e1 = v1-v0, e2 = v2-v0), two-sided, no culling
 * (flags 0), accepted iff tmin < t < tmax.  b1/b2 are the barycentric weights
 * of vertices 1 and 2, as RayIntersection.barycentrics.
 */
PT_HD int pt_ray_tri(pt_v3 o, pt_v3 d, pt_v3 v0, pt_v3 e1, pt_v3 e2,
                     float tmin, float tmax, float* t_out, float* b1_out, float* b2_out) {
    pt_v3 p = pt_cross(d, e2);
    float det = pt_dot(e1, p);
    if (det == 0.0f) return 0;
    float inv = pt_div(1.0f, det);
    pt_v3 s = pt_sub(o, v0);
    float u = pt_dot(s, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    pt_v3 q = pt_cross(s, e1);
    float v = pt_dot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return 0;
    float t = pt_dot(e2, q) * inv;
    if (!(t > tmin && t < tmax)) return 0;
    *t_out = t; *b1_out = u; *b2_out = v;
    return 1;
}

/* ---- light sampling with multiple importance sampling (PTC_LIGHTS_MIS) -----
 * Project-defined (the reference's light branch is unfinished: it computes weights it never applies), shared by oracle
 * and kernels.  The reference scatters by sampling a GGX half vector h about the shading normal n (theta =
 * atan(alpha sqrt(u/(1-u))), bsdf.wgsl:155-187), mirroring v about it, and multiplying the throughput by
 * albedo (1 - metallic) with pdf = 1 -- so, seen as an estimator, its BSDF*cos IS  albedo (1 - metallic) * p_b(wi),
 * p_b being the solid-angle density of that sampling:  p_b(wi) = D(h) (n.h) / (4 v.h),  h = normalize(v + wi),
 * D = alpha^2 / (pi ((n.h)^2 (alpha^2 - 1) + 1)^2). */
PT_HD float pt_ggx_reflect_pdf(pt_v3 n, pt_v3 v, pt_v3 wi, float alpha) {
    pt_v3 h = pt_normalize(pt_add(v, wi));
    float nh = pt_dot(n, h), vh = pt_dot(v, h);
    if (!(nh > 0.0f) || !(vh > 1.0e-6f)) return 0.0f;
    float a2 = alpha * alpha;
    float den = pt_fma(nh * nh, a2 - 1.0f, 1.0f);
    float dd = pt_div(a2, (PT_PI * den) * den);
    return pt_div(dd * nh, 4.0f * vh);
}
/* 1 / p_l: p_l = solid-angle density, seen from the shading point, of choosing one of num_lights emissive triangles
 * uniformly and a uniform point on it = dist^2 / (cos_l * area * num_lights); emitters are two-sided (an emissive hit
 * counts from either side in the reference), hence |.|.  dir = unit vector towards the point, dist2 its squared distance. */
PT_HD float pt_light_inv_pdf(pt_v3 v0, pt_v3 v1, pt_v3 v2, pt_v3 dir, float dist2, float num_lights) {
    pt_v3 ng = pt_cross(pt_sub(v1, v0), pt_sub(v2, v0));           /* |ng| = 2 area */
    return pt_div((pt_abs(pt_dot(ng, dir)) * 0.5f) * num_lights, dist2);
}
/* balance heuristic written on x = p_b / p_l: weight of the BSDF-sampled emitter hit, p_b / (p_b + p_l) = x / (1 + x);
 * and, because BSDF*cos = a * p_b, also the whole factor of the light-sampled term, (a p_b / p_l) * p_l / (p_b + p_l).
 * 1 - 1/(1+x) stays finite for x = inf. */
PT_HD float pt_mis_weight(float pdf_b, float inv_pdf_l) { return 1.0f - pt_div(1.0f, 1.0f + pdf_b * inv_pdf_l); }

#define PT_RAY_TMIN 0.01f      /* rt_helpers.wgsl:49 RayDesc(.., 0.01, 10000.0, ..) */
#define PT_RAY_TMAX 10000.0f

/* ---- bilinear texel address arithmetic -----------------------------------
 * textureSampleLevel(.., uv, 0.0) with address mode Repeat and mag filter
 * Linear (texture.rs:312-322).  Hardware filter weights are fixed-point and
 * implementation-defined; this project defines them as exact fp32 fractions.
 * Returns the two wrapped texel coordinates and the weight of the second.
 */
PT_HD void pt_bilinear_axis(float u, int n, int* i0, int* i1, float* f) {
    float x = pt_fma(u, (float)n, -0.5f);
    if (!(x > -1.0e9f && x < 1.0e9f)) x = 0.0f;   /* NaN/huge uv: defined, not UB */
    float fl = pt_floor(x);
    *f = x - fl;
    int i = (int)fl;
    int a;
    if ((n & (n - 1)) == 0) a = i & (n - 1);        /* power-of-two size: the same residue (two's complement), without an integer division */
    else { a = i % n; if (a < 0) a += n; }
    int b = a + 1; if (b >= n) b = 0;
    *i0 = a; *i1 = b;
}

#endif /* PTMATH_H */
