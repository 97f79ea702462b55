// pt_kernels.cu — sm_100a kernels of the wavefront (see pt_kernels.cuh).
#include "pt_kernels.cuh"

namespace ptc {

constexpr int kTraceThreads = 128;
constexpr int kShadeThreads = 128;
#ifndef PTC_SHADE_MIN_BLOCKS
#define PTC_SHADE_MIN_BLOCKS 8            // <= 64 registers: 32 resident warps per SM to cover the gather latency
#endif

// ---- raygen: rand_init + primary ray (pathtracer.wgsl:27, pathtracer/pathtracer.wgsl:240-250) ----
__global__ void __launch_bounds__(256) k_raygen(ptc_uniforms un, Queue q, FrameCounters* fc, uint32_t npix) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) fc->count[0] = npix;
    if (i >= npix) return;
    uint32_t x = i % un.camera.width, y = i / un.camera.width;
    uint32_t rng; pt_v3 ro, rd;
    gen_primary(un.camera, x, y, rng, ro, rd);
    reinterpret_cast<float4*>(q.ray_o)[i] = make_float4(ro.x, ro.y, ro.z, __uint_as_float(i));
    reinterpret_cast<float4*>(q.ray_d)[i] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(rng));
    reinterpret_cast<float4*>(q.mask)[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    reinterpret_cast<float4*>(q.acc)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- trace: persistent warps, warp-level work fetch -------------------------------------------
__device__ __forceinline__ uint32_t warp_fetch(uint32_t* cursor) {
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(cursor, 32u);
    return __shfl_sync(0xffffffffu, base, 0);
}

// Reference form of the closest-hit kernel: one ray per lane, 32-ray batches, every lane runs the plain
// per-ray loop of pt_device.cuh.  Kept for A/B timing (PTC_TRACE_IMPL=simple) and as the readable statement
// of what k_trace computes; its SIMT efficiency on incoherent rays is ~30 % (ncu, profiles/r01a): a warp
// pays node test + triangle loop on every iteration and runs until its slowest ray is done.
__global__ void __launch_bounds__(kTraceThreads) k_trace_simple(DevScene s, const F4* __restrict__ ray_o, const F4* __restrict__ ray_d,
                                                                 U4* __restrict__ hits, const uint32_t* __restrict__ count_ptr,
                                                                 uint32_t* cursor, unsigned long long* total_rays) {
    const uint32_t n = *count_ptr;
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)n);
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = warp_fetch(cursor);
        if (base >= n) break;
        uint32_t i = base + lane;
        if (i < n) {
            F4 o = ld4(ray_o + i), d = ld4(ray_d + i);
            HitRec h = trace_closest_bvh2(s, pt_mk3(o.x, o.y, o.z), pt_mk3(d.x, d.y, d.z), nullptr);
            reinterpret_cast<uint4*>(hits)[i] = make_uint4(__float_as_uint(h.t), __float_as_uint(h.b1), __float_as_uint(h.b2), h.tri);
        }
    }
}

// Production closest-hit kernel.  Same result as k_trace_simple ray for ray (the hit is the minimum of
// (t, global triangle id) over all accepted triangles, which does not depend on visiting order), but the warp's
// work is re-shaped so that both phases run dense:
//   * node phase: every lane that owns a ray does ONE node step per iteration (both children of a BVH2 node, slab test
//     of pt_device.cuh).  A lane that reaches a leaf does not test it: it appends the leaf's triangles to a per-warp
//     list in shared memory as (lane, triangle) pairs and pops its stack, so it is back on an inner node next iteration.
//   * triangle phase: when >= 32 pairs are pending the warp tests them 32 at a time, lane k taking pair k — the ray
//     comes from shared memory, so a test is not tied to the lane that owns the ray.  Results are combined with a
//     64-bit atomicMin on ((t bits) << 32 | triangle id) in shared memory: smaller t wins, equal t goes to the
//     smaller id — exactly the tie rule of leaf_closest.  The winner then records its barycentrics.
//   * a lane whose stack is empty and whose pairs are all resolved writes its hit and takes the next ray from the
//     device cursor (one aggregated atomicAdd per refill), so lanes do not idle until the slowest ray of a batch ends.
// The pruning distance each lane carries is refreshed after every triangle phase; between two phases it is stale
// (larger), which costs a few extra node visits and never a wrong result.
constexpr int kTraceWarps = kTraceThreads / 32;
constexpr uint32_t kTriCap = 32 + 32 * 4;         // < 32 carried over + at most 4 per lane per iteration
constexpr uint32_t kNoRay = 0xFFFFFFFFu;
#ifndef PTC_WAIT_FLUSH
#define PTC_WAIT_FLUSH 8                           // finished lanes that force a full flush
#endif

struct TraceWarp {
    float ox[32], oy[32], oz[32], dx[32], dy[32], dz[32];
    unsigned long long best[32];                  // (t bits << 32) | global triangle id
    float b1[32], b2[32];
    uint32_t pend[32];                            // unresolved pairs per lane
    uint32_t pair[kTriCap];                       // (lane << 27) | leaf-order triangle index
};

__global__ void __launch_bounds__(kTraceThreads) k_trace(DevScene s, const F4* __restrict__ ray_o, const F4* __restrict__ ray_d,
                                                          U4* __restrict__ hits, const uint32_t* __restrict__ count_ptr,
                                                          uint32_t* cursor, unsigned long long* total_rays) {
    __shared__ TraceWarp sh[kTraceWarps];
    TraceWarp& ws = sh[threadIdx.x >> 5];
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const uint32_t n = *count_ptr;
    if (total_rays && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total_rays, (unsigned long long)n);
    const unsigned long long kMissKey = ((unsigned long long)__float_as_uint(PT_RAY_TMAX) << 32) | 0xFFFFFFFFull;

    int stack[PTC_STACK_SIZE];
    int node = PTC_SENTINEL, sp = 0;
    uint32_t ray = kNoRay;
    RaySlab r = {};
    float tbest = PT_RAY_TMAX;
    uint32_t npair = 0;                            // warp-uniform
    bool exhausted = false;                        // warp-uniform: the cursor has passed n
    ws.pend[lane] = 0;
    __syncwarp();

    for (;;) {
        // ---- refill idle lanes from the device cursor ----
        const bool want = ray == kNoRay;
        const uint32_t wm = __ballot_sync(FULL, want);
        if (wm != 0 && !exhausted) {
            const uint32_t k = __popc(wm);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, k);
            base = __shfl_sync(FULL, base, 0);
            if (base + k >= n) exhausted = true;
            const uint32_t i = base + __popc(wm & lt);
            if (want && i < n) {
                ray = i;
                const F4 o = ld4(ray_o + i), d = ld4(ray_d + i);
                ws.ox[lane] = o.x; ws.oy[lane] = o.y; ws.oz[lane] = o.z;
                ws.dx[lane] = d.x; ws.dy[lane] = d.y; ws.dz[lane] = d.z;
                ws.best[lane] = kMissKey; ws.b1[lane] = 0.0f; ws.b2[lane] = 0.0f;
                const pt_v3 ov = pt_mk3(o.x, o.y, o.z), dv = pt_mk3(d.x, d.y, d.z);
                r = make_ray_slab(ov, dv);
                tbest = PT_RAY_TMAX;
                stack[0] = PTC_SENTINEL; sp = 1;
                node = (s.num_nodes != 0 && ray_is_finite(ov, dv)) ? 0 : PTC_SENTINEL;
            }
        }
        __syncwarp();
        if (__ballot_sync(FULL, ray != kNoRay) == 0) break;

        // ---- node phase ----
        bool force;
        for (;;) {
            if ((uint32_t)node < (uint32_t)PTC_SENTINEL) {                   // inner node
                const F4* np = s.nodes + (size_t)node * 4;
                const F4 q0 = ld4(np), q1 = ld4(np + 1), q2 = ld4(np + 2), q3 = ld4(np + 3);
                bool h0, h1; float tn0, tn1;
                node_test(r, q0, q1, q2, tbest, h0, h1, tn0, tn1);
                const int c0 = (int)f2u(q3.x), c1 = (int)f2u(q3.y);
                const bool swap = tn1 < tn0;
                const int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
                if (h0 && h1) { stack[sp++] = farc; node = nearc; }
                else if (h0 || h1) node = h0 ? c0 : c1;
                else node = stack[--sp];
            }
            // leaves: queue their triangles, continue with the next stack entry
            const bool leaf = node < 0;
            const uint32_t lm = __ballot_sync(FULL, leaf);
            if (lm != 0) {
                const uint32_t enc = ~(uint32_t)node;
                const uint32_t c1m = leaf ? (enc & 7u) : 0u;                  // count - 1, in 0..3
                const uint32_t m0 = __ballot_sync(FULL, (c1m & 1u) != 0), m1 = __ballot_sync(FULL, (c1m & 2u) != 0);
                if (leaf) {
                    const uint32_t pos = npair + __popc(lm & lt) + __popc(m0 & lt) + 2u * __popc(m1 & lt);
                    const uint32_t first = enc >> 3, tag = lane << 27;
                    for (uint32_t m = 0; m <= c1m; ++m) ws.pair[pos + m] = tag | (first + m);
                    ws.pend[lane] += c1m + 1u;
                    node = stack[--sp];
                }
                npair += __popc(lm) + __popc(m0) + 2u * __popc(m1);
            }
            const uint32_t busy = __ballot_sync(FULL, node != PTC_SENTINEL);
            const uint32_t waiting = __ballot_sync(FULL, ray != kNoRay && node == PTC_SENTINEL);
            force = busy == 0 || __popc(waiting) >= PTC_WAIT_FLUSH;
            if (force || npair >= 32u) break;
        }
        __syncwarp();

        // ---- triangle phase: full batches, plus the remainder when forced ----
        const uint32_t nproc = force ? npair : (npair & ~31u);
        for (uint32_t k0 = 0; k0 < nproc; k0 += 32) {
            const uint32_t k = k0 + lane;
            bool hit = false; float t = 0.0f, b1 = 0.0f, b2 = 0.0f; uint32_t src = 0; unsigned long long key = 0;
            if (k < nproc) {
                const uint32_t e = ws.pair[k];
                src = e >> 27;
                const F4* tp = s.tris + (size_t)(e & 0x07FFFFFFu) * 3;
                const F4 a = ld4(tp), b = ld4(tp + 1), c = ld4(tp + 2);
                const pt_v3 o = pt_mk3(ws.ox[src], ws.oy[src], ws.oz[src]), d = pt_mk3(ws.dx[src], ws.dy[src], ws.dz[src]);
                hit = pt_ray_tri(o, d, pt_mk3(a.x, a.y, a.z), pt_mk3(b.x, b.y, b.z), pt_mk3(c.x, c.y, c.z),
                                 PT_RAY_TMIN, PT_RAY_TMAX, &t, &b1, &b2) != 0;
                if (hit) {
                    key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)f2u(a.w);   // t > 0: bit order == value order
                    atomicMin(&ws.best[src], key);
                }
            }
            __syncwarp();
            if (hit && ws.best[src] == key) { ws.b1[src] = b1; ws.b2[src] = b2; }   // keys are unique per ray: one writer
            __syncwarp();
        }
        // carry the remainder to the front of the list and recount what each lane still waits for
        const uint32_t rem = npair - nproc;
        const uint32_t keep = lane < rem ? ws.pair[nproc + lane] : 0u;
        ws.pend[lane] = 0;
        __syncwarp();
        if (lane < rem) { ws.pair[lane] = keep; atomicAdd(&ws.pend[keep >> 27], 1u); }
        npair = rem;
        __syncwarp();

        // ---- refresh pruning distances, retire finished rays ----
        if (ray != kNoRay) {
            const unsigned long long key = ws.best[lane];
            tbest = __uint_as_float((uint32_t)(key >> 32));
            if (node == PTC_SENTINEL && ws.pend[lane] == 0) {
                reinterpret_cast<uint4*>(hits)[ray] = make_uint4((uint32_t)(key >> 32), __float_as_uint(ws.b1[lane]), __float_as_uint(ws.b2[lane]), (uint32_t)key);
                ray = kNoRay;
            }
        }
    }
}

__global__ void __launch_bounds__(kTraceThreads) k_trace_any(DevScene s, const F4* __restrict__ ray_o, const F4* __restrict__ ray_d,
                                                              uint8_t* __restrict__ occluded, uint32_t n, uint32_t* cursor) {
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = warp_fetch(cursor);
        if (base >= n) break;
        uint32_t i = base + lane;
        if (i < n) {
            F4 o = ld4(ray_o + i), d = ld4(ray_d + i);
            float tmax = d.w > 0.0f ? d.w : PT_RAY_TMAX;
            occluded[i] = trace_any_bvh2(s, pt_mk3(o.x, o.y, o.z), pt_mk3(d.x, d.y, d.z), tmax) ? 1 : 0;
        }
    }
}

// ---- shade: one bounce, ballot compaction into the next queue, fused accumulation ---------------
// Work is split statically (warp w takes 32-ray batches w, w + #warps, ...): the cost per ray is nearly uniform, so a
// device cursor would only add one dependent global atomic per batch.
__global__ void __launch_bounds__(kShadeThreads, PTC_SHADE_MIN_BLOCKS) k_shade(DevScene s, PathCfg cfg, Queue qin, Queue qout, const U4* __restrict__ hits,
                                                          FrameCounters* fc, uint32_t segment, Planes planes) {
    const uint32_t n = fc->count[segment];
    uint32_t* next_count = &fc->count[segment + 1];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n; base += nwarps * 32u) {
        uint32_t i = base + lane;
        bool alive = false;
        BounceOut bo;
        uint32_t pixel = 0;
        if (i < n) {
            F4 o = ld4(qin.ray_o + i), d = ld4(qin.ray_d + i), m = ld4(qin.mask + i), a = ld4(qin.acc + i);
            uint4 hr = __ldg(reinterpret_cast<const uint4*>(hits) + i);
            PathState ps;
            ps.o = pt_mk3(o.x, o.y, o.z); ps.d = pt_mk3(d.x, d.y, d.z);
            ps.mask = pt_mk3(m.x, m.y, m.z); ps.acc = pt_mk3(a.x, a.y, a.z);
            ps.rng = __float_as_uint(d.w);
            pixel = __float_as_uint(o.w);
            HitRec h; h.t = __uint_as_float(hr.x); h.b1 = __uint_as_float(hr.y); h.b2 = __uint_as_float(hr.z); h.tri = hr.w;
            bo = shade_bounce(s, cfg, ps, h, segment);
            alive = bo.alive;
            if (segment == 0) {                                   // pathtracer/pathtracer.wgsl:270-273
                planes.depth[pixel] = bo.first_t;
                reinterpret_cast<float4*>(planes.normal)[pixel] = make_float4(bo.first_normal.x, bo.first_normal.y, bo.first_normal.z, 1.0f);
                reinterpret_cast<float4*>(planes.albedo)[pixel] = make_float4(bo.first_albedo.x, bo.first_albedo.y, bo.first_albedo.z, 1.0f);
                planes.hit_tri[pixel] = h.tri;
            }
            if (!alive) accumulate_pixel(planes.color + pixel, bo.radiance, cfg.accum_mode, cfg.is_grabbed);   // pathtracer.wgsl:32-42
        }
        // warp-aggregated slot allocation in the next queue
        uint32_t ballot = __ballot_sync(0xffffffffu, alive);
        if (ballot) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(next_count, (uint32_t)__popc(ballot));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (alive) {
                uint32_t slot = slot0 + __popc(ballot & ((1u << lane) - 1u));
                const PathState& p = bo.next;
                reinterpret_cast<float4*>(qout.ray_o)[slot] = make_float4(p.o.x, p.o.y, p.o.z, __uint_as_float(pixel));
                reinterpret_cast<float4*>(qout.ray_d)[slot] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.rng));
                reinterpret_cast<float4*>(qout.mask)[slot] = make_float4(p.mask.x, p.mask.y, p.mask.z, 0.0f);
                reinterpret_cast<float4*>(qout.acc)[slot] = make_float4(p.acc.x, p.acc.y, p.acc.z, 0.0f);
            }
        }
    }
}

// ---- launchers ------------------------------------------------------------------------------------
void launch_raygen(cudaStream_t st, const ptc_uniforms& un, Queue q, FrameCounters* fc, uint32_t npix) {
    k_raygen<<<(npix + 255) / 256, 256, 0, st>>>(un, q, fc, npix);
}
void launch_trace(cudaStream_t st, int grid, int impl, const DevScene& s, Queue q, U4* hits, FrameCounters* fc, uint32_t segment,
                  unsigned long long* total_rays) {
    if (impl == 0) k_trace<<<grid, kTraceThreads, 0, st>>>(s, q.ray_o, q.ray_d, hits, &fc->count[segment], &fc->trace_cursor[segment], total_rays);
    else k_trace_simple<<<grid, kTraceThreads, 0, st>>>(s, q.ray_o, q.ray_d, hits, &fc->count[segment], &fc->trace_cursor[segment], total_rays);
}
void launch_shade(cudaStream_t st, int grid, const DevScene& s, const PathCfg& cfg, Queue qin, Queue qout, const U4* hits,
                  FrameCounters* fc, uint32_t segment, Planes planes) {
    k_shade<<<grid, kShadeThreads, 0, st>>>(s, cfg, qin, qout, hits, fc, segment, planes);
}

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

void launch_trace_n(cudaStream_t st, int grid, int impl, const DevScene& s, const F4* ray_o, const F4* ray_d, U4* hits, uint32_t n, uint32_t* cursor) {
    // cursor[0] = work cursor (zeroed by caller), cursor[1] = count
    k_set_u32<<<1, 1, 0, st>>>(cursor + 1, n);
    if (impl == 0) k_trace<<<grid, kTraceThreads, 0, st>>>(s, ray_o, ray_d, hits, cursor + 1, cursor, nullptr);
    else k_trace_simple<<<grid, kTraceThreads, 0, st>>>(s, ray_o, ray_d, hits, cursor + 1, cursor, nullptr);
}
void launch_trace_any_n(cudaStream_t st, int grid, const DevScene& s, const F4* ray_o, const F4* ray_d, uint8_t* occluded, uint32_t n, uint32_t* cursor) {
    k_trace_any<<<grid, kTraceThreads, 0, st>>>(s, ray_o, ray_d, occluded, n, cursor);
}

// ---- device evaluation of the shared numerics primitives (unit-test entry: host and device must agree bit for bit) ----
__global__ void k_math(int fn, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = a[i], y = b[i], r = 0.0f;
    switch (fn) {
        case 0: r = pt_sin(x); break;
        case 1: r = pt_cos(x); break;
        case 2: r = pt_atan(x); break;
        case 3: r = pt_acos(x); break;
        case 4: r = pt_atan2(x, y); break;
        case 5: r = pt_div(x, y); break;
        case 6: r = pt_sqrt(x); break;
        case 7: r = pt_fma(x, y, x); break;
        case 8: { uint32_t s = __float_as_uint(x); r = rng_next_f32(s); break; }
        case 9: { pt_v3 v = pt_normalize(pt_mk3(x, y, x * y)); r = v.x + v.y + v.z; break; }
        case 10: { pt_v3 v = pt_refract(pt_normalize(pt_mk3(x, y, 0.5f)), pt_normalize(pt_mk3(y, -x, 1.0f)), 1.5f); r = v.x + v.y + v.z; break; }
        case 11: r = x * y + x; break;                 /* must NOT be contracted (-fmad=false) */
        case 12: { int i0, i1; float f; pt_bilinear_axis(x, 37, &i0, &i1, &f); r = f + (float)i0 + 64.0f * (float)i1; break; }
        default: break;
    }
    out[i] = r;
}
void launch_math(cudaStream_t st, int fn, const float* a, const float* b, float* out, uint32_t n) {
    k_math<<<(n + 255) / 256, 256, 0, st>>>(fn, a, b, out, n);
}

// ---- sample_tex / sky on caller-supplied coordinates (unit-test entry) ------------------------------------------------
__global__ void k_sample(DevScene s, uint32_t tex, int sky, const float* __restrict__ in, float* __restrict__ out, uint32_t n, uint32_t tex_mod) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (tex_mod) tex = (tex + i) % tex_mod;               // a different texture per lane, as in k_shade
    if (sky) { pt_v3 c = sky_radiance(s, pt_mk3(in[i * 3], in[i * 3 + 1], in[i * 3 + 2])); out[i * 4] = c.x; out[i * 4 + 1] = c.y; out[i * 4 + 2] = c.z; out[i * 4 + 3] = 1.0f; }
    else { pt_v4 c = sample_tex(s, tex, in[i * 2], in[i * 2 + 1]); out[i * 4] = c.x; out[i * 4 + 1] = c.y; out[i * 4 + 2] = c.z; out[i * 4 + 3] = c.w; }
}
void launch_sample(cudaStream_t st, const DevScene& s, uint32_t tex, int sky, const float* in, float* out, uint32_t n, uint32_t tex_mod) {
    k_sample<<<(n + 127) / 128, 128, 0, st>>>(s, tex, sky, in, out, n, tex_mod);
}

// ---- probe of the texture unit's u8 -> f32 conversion (cudaReadModeNormalizedFloat) ----------------------------------
__global__ void k_unorm_probe(unsigned long long obj, float* out) {
    uint32_t i = threadIdx.x;
    float4 v = tex2D<float4>((cudaTextureObject_t)obj, (float)i + 0.5f, 0.5f);
    out[i] = v.x; out[256 + i] = v.w;
}
void launch_unorm_probe(cudaStream_t st, unsigned long long obj, float* out) { k_unorm_probe<<<1, 256, 0, st>>>(obj, out); }

static int persistent_grid(const void* kernel, int threads, int device) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    if (sms < 1) sms = 1;
    return sms * per_sm;          // a multiple of the SM count (148 on B200): every SM holds its full share of resident CTAs
}
int trace_grid_size(int device, int impl) { return persistent_grid(impl == 0 ? (const void*)k_trace : (const void*)k_trace_simple, kTraceThreads, device); }
int shade_grid_size(int device) { return persistent_grid((const void*)k_shade, kShadeThreads, device); }

}  // namespace ptc
