// lbvh.cuh — the per-element arithmetic of the linear BVH build (scene ingest at scale), shared by the
// CUDA kernels (bvh_gpu.cu) and by the host restatement in bvh_build.cpp that the CPU-only tests run: Morton keys of
// the triangle centroids, Karras' radix-tree node construction ("Maximizing Parallelism in the Construction of BVHs,
// Octrees, and k-d Trees", HPG 2012), leaf collapsing and the quantised device-node encoding of bvh_build.h.
//
// The reference has no counterpart (its acceleration structures are built by the Vulkan driver, pt.rs:364-468); the
// binned-SAH builder of bvh_build.cpp stays the default because its trees trace faster, this one builds a 10 M-triangle
// scene in milliseconds on the GPU for time-to-first-pixel.  Hit identity does not depend on the tree (§3 of DESIGN.md),
// so both builders pass the same parity tests.
#pragma once
#include "bvh_build.h"
#include "../../include/ptmath.h"
#include <math.h>
#include <string.h>

namespace ptc {

#define PTC_LBVH_MAX_LEAF 2u      // leaves of the collapsed tree hold at most this many triangles (2 traced fastest: tools/bvh_stats.py)

PT_HD uint64_t lbvh_spread3(uint32_t v) {          // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
// centroid (as the SAH builder defines it: midpoint of the triangle's box) on a 2^21 grid over the centroid bounds
PT_HD uint32_t lbvh_quant21(float c, float cmin, float scale) {
    float f = (c - cmin) * scale;
    if (!(f > 0.0f)) return 0u;
    if (f >= 2097151.0f) return 2097151u;
    return (uint32_t)f;
}
PT_HD void lbvh_tri_box(const float* v, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(v[a], fminf(v[3 + a], v[6 + a]));
        hi[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
    }
}
PT_HD uint64_t lbvh_key(const float* v, const float cmin[3], const float scale[3]) {
    float lo[3], hi[3]; lbvh_tri_box(v, lo, hi);
    uint64_t k = 0;
    for (int a = 0; a < 3; ++a) k |= lbvh_spread3(lbvh_quant21(0.5f * (lo[a] + hi[a]), cmin[a], scale[a])) << (2 - a);
    return k;
}

PT_HD int lbvh_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
// common-prefix length of sorted keys i and j; equal keys fall back to the position, so the tree is well defined
// (and stays shallow) for any number of identical centroids
PT_HD int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a != b) return lbvh_clz64(a ^ b);
    return 64 + lbvh_clz64((uint64_t)(uint32_t)i ^ (uint64_t)(uint32_t)j);
}

// inner node i of the radix tree over n sorted keys (n >= 2): the key range it covers and its split position
// (left child covers [first, split], right child [split + 1, last])
PT_HD void lbvh_node(const uint64_t* keys, int n, int i, int& first, int& last, int& split) {
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    split = i + s * d + (d < 0 ? -1 : 0);
    first = i < j ? i : j; last = i < j ? j : i;
}

// scene box -> conservative padding and quantisation grid (same rules as build_bvh2 / emit_bvh2)
struct LbvhFrame { float pad; QuantGrid grid; double step[3]; };
void lbvh_frame(const float lo[3], const float hi[3], LbvhFrame& f);      // bvh_build.cpp

// one child box -> three node words (bvh_build.h DevNode2): planes rounded outwards by one extra grid step
PT_HD void lbvh_put_box(uint32_t* w, const float lo[3], const float hi[3], float pad, const float base[3], const double step[3]) {
    for (int a = 0; a < 3; ++a) {
        const double l = ((double)lo[a] - (double)pad - (double)base[a]) / step[a];
        const double h = ((double)hi[a] + (double)pad - (double)base[a]) / step[a];
        long long ql = (long long)floor(l) - 1, qh = (long long)ceil(h) + 1;
        ql = ql < 0 ? 0 : (ql > 32767 ? 32767 : ql); qh = qh < 0 ? 0 : (qh > 32767 ? 32767 : qh);
        w[a] = (0x8000u | (uint32_t)ql) | ((0x8000u | (uint32_t)qh) << 16);
    }
}
PT_HD void lbvh_put_tri(float* q, const float* v, uint32_t id) {           // bvh_build.h DevTri
    for (int k = 0; k < 16; ++k) q[k] = 0.0f;
    q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; memcpy(&q[3], &id, 4);
    q[4] = v[3] - v[0]; q[5] = v[4] - v[1]; q[6] = v[5] - v[2];
    q[8] = v[6] - v[0]; q[9] = v[7] - v[1]; q[10] = v[8] - v[2];
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume
// Hierarchy Construction", TVCG 2018) — the bottom-up builder behind PTC_BVH_PLOC_DEVICE.  The Morton-sorted primitives
// start as clusters; every round each cluster looks for its best partner among the PTC_PLOC_RADIUS clusters on either side
// (smallest surface area of the joint box), mutual best partners merge, the array is compacted, until one cluster is left.
// The trees are close to (here: better than) the binned-SAH builder's in traversal cost, and every round is a handful of
// data-parallel kernels.
// (window: +-4.  Swept +-2 .. +-100 on four scenes: wider windows gave WORSE trees here)
#ifndef PTC_PLOC_RADIUS
#define PTC_PLOC_RADIUS 4
#endif
#ifndef PTC_PLOC_MAX_LEAF
#define PTC_PLOC_MAX_LEAF 2u      // subtrees of at most this many triangles are collapsed into one leaf
#endif
struct PlocBox { float lo[3], hi[3]; };
PT_HD float ploc_joint_area(const PlocBox& a, const PlocBox& b) {        // half the surface area of the union, fixed operation order
    const float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]);
    const float dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]);
    const float dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
    return dx * dy + (dy * dz + dz * dx);
}
// best partner of cluster i among its neighbours in the current cluster array (m clusters, cur[] = node index per cluster).
// Candidate PAIRS are ordered by (area, index distance, parity of the lower index — even first —, lower index): a strict total
// order on pairs, so the globally best pair is always mutual and every round merges at least one pair whatever the ties; and
// among clusters with IDENTICAL boxes (coincident triangles) the order pairs neighbours (0,1) (2,3) ... so the tree stays
// balanced instead of degenerating into a chain.
PT_HD int ploc_nearest(const PlocBox* box, const uint32_t* cur, int m, int i, int radius) {
    const PlocBox bi = box[cur[i]];
    const int j0 = i - radius < 0 ? 0 : i - radius, j1 = i + radius > m - 1 ? m - 1 : i + radius;
    float best = INFINITY; int bj = -1, bd = 0, bmn = 0;
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const float c = ploc_joint_area(bi, box[cur[j]]);
        const int d = j < i ? i - j : j - i, mn = j < i ? j : i;
        bool better = c < best;
        if (c == best && bj >= 0) {
            if (d != bd) better = d < bd;
            else if ((mn & 1) != (bmn & 1)) better = (mn & 1) == 0;
            else better = mn < bmn;
        }
        if (better || bj < 0) { best = c; bj = j; bd = d; bmn = mn; }
    }
    return bj;
}

}  // namespace ptc
