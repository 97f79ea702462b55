// warpsim.cpp — DEVELOPMENT AID (test-only, part of libpt_hostemu.so, never of libptcore).
//
// A lane-accurate model of k_trace's warp scheduling (pt_kernels.cu): 32 persistent ray slots per warp, phase voting
// (refill / triangle / node), device cursor refill — running the host build of the same node_test / pt_ray_tri.  It
// counts warp iterations per phase and the lanes active in each, so scheduling policies (vote thresholds, postponed
// leaves, ...) can be compared on the CPU before spending GPU time: k_trace is issue-bound, so
// sum(iterations x instructions per iteration) tracks its run time.
#include "../../wgpu-pathtracer_b200/csrc/pt_device.cuh"
#include "../../wgpu-pathtracer_b200/csrc/scene_prep.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace ptc;

struct emu_scene_view { const DevScene* ds; };
extern "C" const DevScene* emu_dev_scene(const void* e);
extern "C" void emu_path_cfg(const void* e, const ptc_uniforms* un, PathCfg* pc, ptc_uniforms* out_un);

namespace {

struct Lane {
    int node = PTC_SENTINEL, sp = 0;
    int stack[PTC_STACK_SIZE];
    uint32_t ray = 0xFFFFFFFFu, tri_i = 0, tri_end = 0;
    int pend = 0;                 // postponed leaf (policy 1), 0 = none (leaf refs are negative)
    pt_v3 o, d;
    RaySlab r;
    HitRec best;
};

struct SimParams {
    uint32_t tri_vote, refill_vote, policy, nwarps;
    uint32_t node_steps;          // node steps per voted node phase (policy 2)
    uint32_t node_inner;          // > 0: stay in the node phase while at least this many lanes can take a node step
    uint32_t tri_inner;           // > 0: stay in the triangle phase while at least this many lanes have a triangle to test
    uint32_t _pad;
};

struct SimOut {
    uint64_t it_refill, it_tri, it_node;          // warp iterations per phase
    uint64_t ln_refill, ln_tri, ln_node;          // lanes doing useful work in those iterations
    uint64_t node_visits, tri_tests, rays;
    uint64_t mismatches;                          // hits differing from trace_closest_bvh2 (must be 0)
    uint64_t it_outer, it_inner_vote;             // outer-loop decisions (full vote) and inner-loop votes
};

inline void tri_step(const DevScene& s, Lane& L, int leaf, uint32_t& ti, uint32_t& te) {
    if (ti == te) { const uint32_t enc = ~(uint32_t)leaf; ti = enc >> 3; te = ti + (enc & 7u) + 1u; }
    const F4* tp = s.tris + (size_t)ti * 4;
    F4 a, b; ld8(tp, a, b); const F4 c = ld4(tp + 2);
    float t, b1, b2;
    if (pt_ray_tri(L.o, L.d, pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z), PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) {
        const uint32_t id = f2u(a.w);
        if (t < L.best.t || (t == L.best.t && id < L.best.tri)) { L.best.t = t; L.best.b1 = b1; L.best.b2 = b2; L.best.tri = id; }
    }
    ++ti;
}

inline void node_step(const DevScene& s, Lane& L) {
    const NodeQ nq = ld_node(s.nodes + L.node);
    bool h0, h1; float tn0, tn1;
    node_test(L.r, nq, L.best.t, h0, h1, tn0, tn1);
    const int c0 = (int)nq.w[3], c1 = (int)nq.w[7];
    const bool swap = tn1 < tn0;
    if (h0 && h1) { L.stack[L.sp++] = swap ? c0 : c1; L.node = swap ? c1 : c0; }
    else if (h0 || h1) L.node = h0 ? c0 : c1;
    else L.node = L.stack[--L.sp];
}

}  // namespace

extern "C" __attribute__((visibility("default")))
int emu_warpsim(const void* scene, const float* rays, uint64_t n, const SimParams* P, SimOut* out, uint32_t* hits_out) {
    const DevScene& s = *emu_dev_scene(scene);
    std::memset(out, 0, sizeof(*out));
    const uint32_t NW = P->nwarps ? P->nwarps : 1;
    std::vector<Lane> lanes((size_t)NW * 32);
    std::vector<uint8_t> exhausted(NW, 0), finished(NW, 0);
    std::vector<HitRec> res(n);
    uint64_t cursor = 0;
    uint32_t live = NW;
    while (live) {
        for (uint32_t w = 0; w < NW; ++w) {
            if (finished[w]) continue;
            Lane* L = &lanes[(size_t)w * 32];
            uint32_t idle = 0, at_leaf = 0, at_node = 0, pend = 0;
            for (int l = 0; l < 32; ++l) {
                if (L[l].node == PTC_SENTINEL && L[l].pend == 0) idle |= 1u << l;
                else if (L[l].node < 0) at_leaf |= 1u << l;
                else if (L[l].node != PTC_SENTINEL) at_node |= 1u << l;
                if (L[l].pend != 0) pend |= 1u << l;
            }
            const uint32_t nidle = (uint32_t)__builtin_popcount(idle);
            out->it_outer++;
            if (idle == 0xFFFFFFFFu || (!exhausted[w] && nidle >= P->refill_vote)) {
                out->it_refill++;
                for (int l = 0; l < 32; ++l) if ((idle >> l & 1) && L[l].ray != 0xFFFFFFFFu) { res[L[l].ray] = L[l].best; L[l].ray = 0xFFFFFFFFu; }
                if (!exhausted[w]) {
                    uint64_t base = cursor; cursor += nidle;
                    if (base + nidle >= n) exhausted[w] = 1;
                    uint32_t k = 0;
                    for (int l = 0; l < 32; ++l) if (idle >> l & 1) {
                        uint64_t i = base + k++;
                        if (i >= n) continue;
                        Lane& X = L[l];
                        X.ray = (uint32_t)i;
                        X.o = pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]); X.d = pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]);
                        X.r = make_ray_slab(s, X.o, X.d);
                        X.best.t = PT_RAY_TMAX; X.best.b1 = X.best.b2 = 0.0f; X.best.tri = 0xFFFFFFFFu;
                        X.stack[0] = PTC_SENTINEL; X.sp = 1; X.tri_i = X.tri_end = 0; X.pend = 0;
                        if (s.num_nodes != 0 && ray_is_finite(X.o, X.d)) X.node = 0;
                        out->ln_refill++; out->rays++;
                    }
                }
                bool any = false;
                for (int l = 0; l < 32; ++l) any |= L[l].ray != 0xFFFFFFFFu;
                if (!any) { finished[w] = 1; --live; }
                continue;
            }
            if (P->policy == 3) {
                // policy 3: greedy — run the phase that serves the most lanes per instruction (node 40, triangle 62, refill 150
                // instructions per iteration); a phase repeats while it keeps at least 3/4 of the lanes it started with
                const uint32_t nnode = (uint32_t)__builtin_popcount(at_node), nleaf = (uint32_t)__builtin_popcount(at_leaf);
                const double sn = nnode / 40.0, st = nleaf / 62.0, sr = exhausted[w] ? 0.0 : nidle / (double)(P->node_steps ? P->node_steps : 150);
                if (sr > sn && sr > st) {
                    out->it_refill++;
                    for (int l = 0; l < 32; ++l) if ((idle >> l & 1) && L[l].ray != 0xFFFFFFFFu) { res[L[l].ray] = L[l].best; L[l].ray = 0xFFFFFFFFu; }
                    uint64_t base = cursor; cursor += nidle;
                    if (base + nidle >= n) exhausted[w] = 1;
                    uint32_t k = 0;
                    for (int l = 0; l < 32; ++l) if (idle >> l & 1) {
                        uint64_t i = base + k++;
                        if (i >= n) continue;
                        Lane& X = L[l];
                        X.ray = (uint32_t)i;
                        X.o = pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]); X.d = pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]);
                        X.r = make_ray_slab(s, X.o, X.d);
                        X.best.t = PT_RAY_TMAX; X.best.b1 = X.best.b2 = 0.0f; X.best.tri = 0xFFFFFFFFu;
                        X.stack[0] = PTC_SENTINEL; X.sp = 1; X.tri_i = X.tri_end = 0; X.pend = 0;
                        if (s.num_nodes != 0 && ray_is_finite(X.o, X.d)) X.node = 0;
                        out->ln_refill++; out->rays++;
                    }
                } else if (st >= sn) {
                    const uint32_t start = nleaf;
                    for (;;) {
                        out->it_tri++;
                        uint32_t still = 0;
                        for (int l = 0; l < 32; ++l) if (L[l].node < 0) {
                            Lane& X = L[l];
                            tri_step(s, X, X.node, X.tri_i, X.tri_end); out->tri_tests++; out->ln_tri++;
                            if (X.tri_i == X.tri_end) X.node = X.stack[--X.sp];
                            if (X.node < 0) still++;
                        }
                        out->it_inner_vote++;
                        if (still * 4 < start * 3 || still == 0) break;
                    }
                } else {
                    const uint32_t start = nnode;
                    for (;;) {
                        uint32_t act = 0, still = 0;
                        for (int l = 0; l < 32; ++l) if ((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL) { node_step(s, L[l]); out->node_visits++; act++; }
                        for (int l = 0; l < 32; ++l) if ((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL) still++;
                        out->it_node++; out->ln_node += act; out->it_inner_vote++;
                        if (still * 4 < start * 3 || still == 0) break;
                    }
                }
            } else if (P->policy == 0 || P->policy == 2) {
                const uint32_t nleaf = (uint32_t)__builtin_popcount(at_leaf);
                if (nleaf >= P->tri_vote || (at_leaf | idle) == 0xFFFFFFFFu) {
                    for (;;) {
                        out->it_tri++;
                        uint32_t still = 0;
                        for (int l = 0; l < 32; ++l) if (L[l].node < 0) {
                            Lane& X = L[l];
                            tri_step(s, X, X.node, X.tri_i, X.tri_end); out->tri_tests++; out->ln_tri++;
                            if (X.tri_i == X.tri_end) X.node = X.stack[--X.sp];
                            if (X.node < 0) still++;
                        }
                        if (!P->tri_inner) break;
                        out->it_inner_vote++;
                        if (still < P->tri_inner) break;
                    }
                } else {
                    const uint32_t steps = P->policy == 2 ? (P->node_steps ? P->node_steps : 1) : 1;
                    for (uint32_t k = 0; k < steps; ++k) {
                        uint32_t act = 0;
                        for (int l = 0; l < 32; ++l) if ((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL) { node_step(s, L[l]); out->node_visits++; act++; }
                        if (!act) break;
                        out->it_node++; out->ln_node += act;
                        if (getenv("WS_DIAG")) { static uint64_t a_leaf = 0, a_idle = 0, its = 0; uint32_t nl = 0, ni = 0;
                            for (int l = 0; l < 32; ++l) { if (L[l].node == PTC_SENTINEL) ni++; else if (L[l].node < 0) nl++; }
                            a_leaf += nl; a_idle += ni; its++; if (its % 20000 == 0) fprintf(stderr, "node its %llu: after-step at_leaf %.2f idle %.2f\n", (unsigned long long)its, (double)a_leaf / its, (double)a_idle / its); }
                        if (P->node_inner) {
                            uint32_t still = 0;
                            for (int l = 0; l < 32; ++l) if ((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL) still++;
                            out->it_inner_vote++;
                            if (still >= P->node_inner) --k;      // stay
                        }
                    }
                }
            } else {
                // policy 1: postponed leaves.  A lane reaching a leaf parks it in `pend` and keeps traversing; it only
                // waits (node < 0) when it reaches a second leaf.  The triangle phase works on pend first.
                uint32_t want_tri = 0;
                for (int l = 0; l < 32; ++l) if (L[l].pend != 0) want_tri++;
                uint32_t blocked = 0;      // lanes that cannot take a node step
                for (int l = 0; l < 32; ++l) if (!((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL)) blocked++;
                if (want_tri >= P->tri_vote || blocked == 32) {
                    for (;;) {
                        out->it_tri++;
                        uint32_t still = 0;
                        for (int l = 0; l < 32; ++l) if (L[l].pend != 0) {
                            Lane& X = L[l];
                            tri_step(s, X, X.pend, X.tri_i, X.tri_end); out->tri_tests++; out->ln_tri++;
                            if (X.tri_i == X.tri_end) {
                                X.pend = 0;
                                if (X.node < 0) { X.pend = X.node; X.node = X.stack[--X.sp]; }     // the waiting leaf moves up, traversal resumes
                            }
                            if (X.pend != 0) still++;
                        }
                        if (!P->tri_inner) break;
                        out->it_inner_vote++;
                        if (still < P->tri_inner) break;
                    }
                } else {
                    for (;;) {
                        uint32_t act = 0, still = 0;
                        for (int l = 0; l < 32; ++l) {
                            Lane& X = L[l];
                            if ((uint32_t)X.node < (uint32_t)PTC_SENTINEL) {
                                node_step(s, X); out->node_visits++; act++;
                                if (X.node < 0 && X.pend == 0) { X.pend = X.node; X.node = X.stack[--X.sp]; }
                            }
                            if ((uint32_t)X.node < (uint32_t)PTC_SENTINEL) still++;
                        }
                        out->it_node++; out->ln_node += act;
                        if (!P->node_inner) break;
                        out->it_inner_vote++;
                        if (still < P->node_inner) break;
                    }
                }
            }
        }
    }
    for (uint64_t i = 0; i < n; ++i) {
        if (hits_out) { std::memcpy(&hits_out[i * 4], &res[i].t, 4); std::memcpy(&hits_out[i * 4 + 1], &res[i].b1, 4); std::memcpy(&hits_out[i * 4 + 2], &res[i].b2, 4); hits_out[i * 4 + 3] = res[i].tri; }
        HitRec h = trace_closest_bvh2(s, pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]), pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]), nullptr);
        if (h.tri != res[i].tri || std::memcmp(&h.t, &res[i].t, 4) != 0) out->mismatches++;
    }
    return 0;
}

// rays of every path alive at `segment` (pixel order), as (ox, oy, oz, 0, dx, dy, dz, 0); returns the count
extern "C" __attribute__((visibility("default")))
uint64_t emu_collect_rays(const void* scene, const ptc_uniforms* un_in, uint32_t segment, float* out, uint64_t max_rays) {
    const DevScene& s = *emu_dev_scene(scene);
    PathCfg pc; ptc_uniforms un;
    emu_path_cfg(scene, un_in, &pc, &un);
    uint64_t n = 0;
    for (uint32_t y = 0; y < pc.height; ++y) for (uint32_t x = 0; x < pc.width; ++x) {
        PathState ps; ps.mask = pt_mk3(1.0f, 1.0f, 1.0f); ps.acc = pt_mk3(0.0f, 0.0f, 0.0f);
        gen_primary(un.camera, x, y, ps.rng, ps.o, ps.d);
        for (uint32_t seg = 0;; ++seg) {
            if (seg == segment) {
                if (n < max_rays) { float* r = out + n * 8; r[0] = ps.o.x; r[1] = ps.o.y; r[2] = ps.o.z; r[3] = 0; r[4] = ps.d.x; r[5] = ps.d.y; r[6] = ps.d.z; r[7] = 0; }
                ++n; break;
            }
            HitRec h = trace_closest_bvh2(s, ps.o, ps.d, nullptr);
            BounceOut bo = shade_bounce(s, pc, ps, h, seg);
            if (!bo.alive) break;
            ps = bo.next;
        }
    }
    return n;
}

// Host model of k_trace_packet (pt_kernels.cu): 32 consecutive rays walk the tree together with one uniform node and stack,
// descending where ANY ray needs to, the near child chosen by the vote of the rays that enter both.  hits_out as
// emu_warpsim; stats: [0] node visits, [1] triangle tests (per packet), [2] hits differing from trace_closest_bvh2.
extern "C" __attribute__((visibility("default")))
int emu_packet_trace(const void* scene, const float* rays, uint64_t n, uint32_t* hits_out, uint64_t stats[3]) {
    const DevScene& s = *emu_dev_scene(scene);
    stats[0] = stats[1] = stats[2] = 0;
    for (uint64_t base = 0; base < n; base += 32) {
        pt_v3 o[32], d[32]; RaySlab r[32]; HitRec best[32]; bool live[32];
        bool any = false;
        for (int l = 0; l < 32; ++l) {
            const uint64_t i = base + l;
            o[l] = pt_mk3(0.0f, 0.0f, 0.0f); d[l] = pt_mk3(1.0f, 0.0f, 0.0f);
            if (i < n) { const float* q = rays + i * 8; o[l] = pt_mk3(q[0], q[1], q[2]); d[l] = pt_mk3(q[4], q[5], q[6]); }
            live[l] = i < n && s.num_nodes != 0 && ray_is_finite(o[l], d[l]);
            r[l] = make_ray_slab(s, o[l], d[l]);
            best[l].t = PT_RAY_TMAX; best[l].b1 = best[l].b2 = 0.0f; best[l].tri = 0xFFFFFFFFu;
            any = any || live[l];
        }
        int stack[PTC_STACK_SIZE]; int sp = 0; stack[sp++] = PTC_SENTINEL;
        int node = any ? 0 : PTC_SENTINEL;
        while (node != PTC_SENTINEL) {
            if (node >= 0) {
                const NodeQ nq = ld_node(s.nodes + node); stats[0]++;
                int m0 = 0, m1 = 0, both = 0, far_first = 0;
                for (int l = 0; l < 32; ++l) {
                    bool h0, h1; float a, b; node_test(r[l], nq, best[l].t, h0, h1, a, b);
                    h0 = h0 && live[l]; h1 = h1 && live[l];
                    m0 += h0; m1 += h1; both += h0 && h1; far_first += h0 && h1 && b < a;
                }
                const int c0 = (int)nq.w[3], c1 = (int)nq.w[7];
                if (m0 && m1) { const bool swap = 2 * far_first > both; stack[sp++] = swap ? c0 : c1; node = swap ? c1 : c0; }
                else if (m0 || m1) node = m0 ? c0 : c1;
                else node = stack[--sp];
            } else {
                const uint32_t enc = ~(uint32_t)node, first = enc >> 3, cnt = (enc & 7u) + 1u;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const F4* tp = s.tris + (size_t)(first + k) * 4;
                    F4 a, b; ld8(tp, a, b); const F4 c = ld4(tp + 2); stats[1]++;
                    for (int l = 0; l < 32; ++l) {
                        float t, b1, b2;
                        if (live[l] && pt_ray_tri(o[l], d[l], pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z), PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) {
                            const uint32_t id = f2u(a.w);
                            if (t < best[l].t || (t == best[l].t && id < best[l].tri)) { best[l].t = t; best[l].b1 = b1; best[l].b2 = b2; best[l].tri = id; }
                        }
                    }
                }
                node = stack[--sp];
            }
        }
        for (int l = 0; l < 32 && base + l < n; ++l) {
            const uint64_t i = base + l;
            if (hits_out) {                                    // a miss is reported as (0, 0, 0, 0xFFFFFFFF), as emu_trace_rays / orc_trace_rays do
                const bool hit = best[l].tri != 0xFFFFFFFFu; const float z = 0.0f;
                std::memcpy(&hits_out[i * 4], hit ? &best[l].t : &z, 4); std::memcpy(&hits_out[i * 4 + 1], hit ? &best[l].b1 : &z, 4);
                std::memcpy(&hits_out[i * 4 + 2], hit ? &best[l].b2 : &z, 4); hits_out[i * 4 + 3] = best[l].tri;
            }
            HitRec h = trace_closest_bvh2(s, o[l], d[l], nullptr);
            if (h.tri != best[l].tri || std::memcmp(&h.t, &best[l].t, 4) != 0 || std::memcmp(&h.b1, &best[l].b1, 4) != 0) stats[2]++;
        }
    }
    return 0;
}

// ---- model of a CTA-level triangle pool (development aid) ----------------------------------------------------------------
// G warps share one queue of (ray slot, triangle) requests.  A lane that reaches a leaf pushes one request per triangle and
// keeps descending (its closest hit may shrink late: a few more nodes); it blocks only when it would exceed `cap` requests in
// flight.  Any warp runs a TRIANGLE PASS (32 requests of any lane of the CTA, one triangle each, all lanes busy) when the queue
// holds a full batch, or a partial one when none of its own lanes can take a node step; otherwise it runs a node step.
// Counts warp iterations per kind so that sum(iterations x instructions) can be compared with emu_warpsim's.
struct PoolParams { uint32_t warps_per_cta, ctas, cap, refill_vote, node_min, _pad[3]; };
struct PoolOut { uint64_t it_refill, it_tri, it_node, ln_tri, ln_node, node_visits, tri_tests, rays, mismatches, it_decide, blocked_lane_steps, _pad; };

extern "C" __attribute__((visibility("default")))
int emu_warpsim_pool(const void* scene, const float* rays, uint64_t n, const PoolParams* P, PoolOut* out) {
    const DevScene& s = *emu_dev_scene(scene);
    std::memset(out, 0, sizeof(*out));
    const uint32_t G = P->warps_per_cta ? P->warps_per_cta : 4, NC = P->ctas ? P->ctas : 16, NW = G * NC;
    struct PLane : Lane { uint32_t inflight = 0; };
    struct Req { uint32_t lane; uint32_t tri; };
    std::vector<PLane> lanes((size_t)NW * 32);
    std::vector<std::vector<Req>> queue(NC);
    std::vector<size_t> qhead(NC, 0);
    std::vector<uint8_t> exhausted(NW, 0), finished(NW, 0);
    std::vector<HitRec> res(n);
    uint64_t cursor = 0; uint32_t live = NW;
    auto tri_one = [&](PLane& X, uint32_t ti) {
        const F4* tp = s.tris + (size_t)ti * 4;
        F4 a, b; ld8(tp, a, b); const F4 c = ld4(tp + 2);
        float t, b1, b2;
        if (pt_ray_tri(X.o, X.d, pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z), PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2)) {
            const uint32_t id = f2u(a.w);
            if (t < X.best.t || (t == X.best.t && id < X.best.tri)) { X.best.t = t; X.best.b1 = b1; X.best.b2 = b2; X.best.tri = id; }
        }
    };
    while (live) {
        for (uint32_t w = 0; w < NW; ++w) {
            if (finished[w]) continue;
            const uint32_t cta = w / G;
            PLane* L = &lanes[(size_t)w * 32];
            uint32_t nidle = 0, can_node = 0, blocked = 0;
            for (int l = 0; l < 32; ++l) {
                const bool done = L[l].node == PTC_SENTINEL && L[l].inflight == 0;
                if (done) nidle++;
                else if ((uint32_t)L[l].node < (uint32_t)PTC_SENTINEL || L[l].node < 0) {
                    // at a leaf: can proceed only if its requests fit
                    if (L[l].node < 0) { const uint32_t cnt = ((~(uint32_t)L[l].node) & 7u) + 1u; if (L[l].inflight + cnt <= P->cap) can_node++; else blocked++; }
                    else can_node++;
                } else blocked++;             // traversal finished, results outstanding
            }
            out->it_decide++;
            const size_t qn = queue[cta].size() - qhead[cta];
            if (nidle == 32 || (!exhausted[w] && nidle >= P->refill_vote)) {
                out->it_refill++;
                for (int l = 0; l < 32; ++l) if (L[l].node == PTC_SENTINEL && L[l].inflight == 0 && L[l].ray != 0xFFFFFFFFu) { res[L[l].ray] = L[l].best; L[l].ray = 0xFFFFFFFFu; }
                if (!exhausted[w]) {
                    for (int l = 0; l < 32; ++l) if (L[l].node == PTC_SENTINEL && L[l].inflight == 0 && L[l].ray == 0xFFFFFFFFu) {
                        const uint64_t i = cursor++;
                        if (i >= n) { exhausted[w] = 1; continue; }
                        PLane& X = L[l];
                        X.ray = (uint32_t)i;
                        X.o = pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]); X.d = pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]);
                        X.r = make_ray_slab(s, X.o, X.d);
                        X.best.t = PT_RAY_TMAX; X.best.b1 = X.best.b2 = 0.0f; X.best.tri = 0xFFFFFFFFu;
                        X.stack[0] = PTC_SENTINEL; X.sp = 1; X.inflight = 0;
                        if (s.num_nodes != 0 && ray_is_finite(X.o, X.d)) X.node = 0;
                        out->rays++;
                    }
                    if (cursor >= n) exhausted[w] = 1;
                }
                bool any = false;
                for (int l = 0; l < 32; ++l) any |= L[l].ray != 0xFFFFFFFFu;
                if (!any) {
                    // a finished warp may still be needed to drain the CTA's queue
                    bool cta_busy = qn != 0;
                    if (!cta_busy) { finished[w] = 1; --live; }
                    else goto tri_pass;
                }
                continue;
            }
            if (qn >= 32 || (can_node < P->node_min && qn > 0)) {
            tri_pass:
                const size_t qn2 = queue[cta].size() - qhead[cta];
                const size_t take = qn2 < 32 ? qn2 : 32;
                if (take == 0) continue;
                out->it_tri++; out->ln_tri += take;
                for (size_t k = 0; k < take; ++k) {
                    const Req r = queue[cta][qhead[cta]++];
                    PLane& X = lanes[r.lane];
                    tri_one(X, r.tri); X.inflight--; out->tri_tests++;
                }
                if (qhead[cta] > 4096 && qhead[cta] * 2 > queue[cta].size()) { queue[cta].erase(queue[cta].begin(), queue[cta].begin() + (long)qhead[cta]); qhead[cta] = 0; }
                continue;
            }
            if (can_node == 0) { out->blocked_lane_steps += blocked; continue; }    // nothing to do: waits for another warp's triangle pass
            out->it_node++;
            for (int l = 0; l < 32; ++l) {
                PLane& X = L[l];
                if (X.node < 0) {
                    // a leaf reached in an earlier step: push its triangles now if they fit, then go on
                    const uint32_t enc = ~(uint32_t)X.node, first = enc >> 3, cnt = (enc & 7u) + 1u;
                    if (X.inflight + cnt > P->cap) continue;
                    for (uint32_t k = 0; k < cnt; ++k) queue[cta].push_back(Req{(uint32_t)(w * 32 + l), first + k});
                    X.inflight += cnt;
                    X.node = X.stack[--X.sp];
                    out->ln_node++;
                    continue;
                }
                if ((uint32_t)X.node < (uint32_t)PTC_SENTINEL) { node_step(s, X); out->node_visits++; out->ln_node++; }
            }
            out->blocked_lane_steps += blocked;
        }
    }
    for (uint64_t i = 0; i < n; ++i) {
        HitRec h = trace_closest_bvh2(s, pt_mk3(rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2]), pt_mk3(rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6]), nullptr);
        if (h.tri != res[i].tri || std::memcmp(&h.t, &res[i].t, 4) != 0) out->mismatches++;
    }
    return 0;
}
