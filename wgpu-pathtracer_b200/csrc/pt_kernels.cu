// pt_kernels.cu — sm_100a kernels of the wavefront (see pt_kernels.cuh).
#include "pt_kernels.cuh"

namespace ptc {

constexpr int kTraceThreads = 128;
constexpr int kShadeThreads = 128;

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

__global__ void __launch_bounds__(kTraceThreads) k_trace(DevScene s, const F4* __restrict__ ray_o, const F4* __restrict__ ray_d,
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
__global__ void __launch_bounds__(kShadeThreads) k_shade(DevScene s, PathCfg cfg, Queue qin, Queue qout, const U4* __restrict__ hits,
                                                          FrameCounters* fc, uint32_t segment, Planes planes) {
    const uint32_t n = fc->count[segment];
    uint32_t* cursor = &fc->shade_cursor[segment];
    uint32_t* next_count = &fc->count[segment + 1];
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = warp_fetch(cursor);
        if (base >= n) break;
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
void launch_trace(cudaStream_t st, int grid, const DevScene& s, Queue q, U4* hits, FrameCounters* fc, uint32_t segment,
                  unsigned long long* total_rays) {
    k_trace<<<grid, kTraceThreads, 0, st>>>(s, q.ray_o, q.ray_d, hits, &fc->count[segment], &fc->trace_cursor[segment], total_rays);
}
void launch_shade(cudaStream_t st, int grid, const DevScene& s, const PathCfg& cfg, Queue qin, Queue qout, const U4* hits,
                  FrameCounters* fc, uint32_t segment, Planes planes) {
    k_shade<<<grid, kShadeThreads, 0, st>>>(s, cfg, qin, qout, hits, fc, segment, planes);
}

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

void launch_trace_n(cudaStream_t st, int grid, const DevScene& s, const F4* ray_o, const F4* ray_d, U4* hits, uint32_t n, uint32_t* cursor) {
    // cursor[0] = work cursor (zeroed by caller), cursor[1] = count
    k_set_u32<<<1, 1, 0, st>>>(cursor + 1, n);
    k_trace<<<grid, kTraceThreads, 0, st>>>(s, ray_o, ray_d, hits, cursor + 1, cursor, nullptr);
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

static int persistent_grid(const void* kernel, int threads, int device) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    if (sms < 1) sms = 1;
    return sms * per_sm;          // a multiple of the SM count (148 on B200): every SM holds its full share of resident CTAs
}
int trace_grid_size(int device) { return persistent_grid((const void*)k_trace, kTraceThreads, device); }
int shade_grid_size(int device) { return persistent_grid((const void*)k_shade, kShadeThreads, device); }

}  // namespace ptc
