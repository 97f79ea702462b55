// pt_kernels.cuh — the wavefront of CUDA kernels that replaces the reference's single WGSL
// compute dispatch (src/shaders/pathtracer/pathtracer.wgsl:21-43, dispatched at
// src/render_steps/pt.rs:547-565 as ceil(W/25) x ceil(H/25) groups of 625 threads).
//
// One frame = raygen -> (trace -> shade) x (MAX_BOUNCES + 1).  Paths live in SoA queues indexed by
// queue slot; `shade` compacts survivors into the next queue with a warp ballot and accumulates
// finished paths straight into the COLOR plane (every pixel finishes exactly once per frame, so no
// separate accumulate pass and no atomics on the image are needed).  trace and shade are persistent
// (a fixed grid sized to the SM count; trace pulls rays from a device-side cursor), and the per-segment
// ray counts live in device memory, so they never travel to the host and a frame is one uninterrupted
// stream of launches.
#pragma once
#include "pt_device.cuh"

namespace ptc {

#define PTC_MAX_SEGMENTS 64           // max_bounces + 1 <= 64

// device counters, one block per frame in flight on the stream (reset by a memset node per frame)
struct FrameCounters {
    uint32_t count[PTC_MAX_SEGMENTS + 1];   // count[b] = rays alive entering segment b
    uint32_t trace_cursor[PTC_MAX_SEGMENTS];
    uint32_t shade_cursor[PTC_MAX_SEGMENTS];
    uint32_t shadow_count[PTC_MAX_SEGMENTS];   // PTC_LIGHTS_MIS: shadow rays queued by shade(b)
    uint32_t shadow_cursor[PTC_MAX_SEGMENTS];
    uint32_t sky_count;                        // escaped paths queued for k_sky by this frame's shade launches
};

struct SkyQueue {         // escaped paths of one frame awaiting their environment lookup (at most one per pixel)
    F4* dir;              // direction.xyz, pixel index (bits)
    F4* mask;             // color_mask.rgb, 1 if `acc` was written (it is all zero bits otherwise)
    F4* acc;              // accumulated_color.rgb
};

struct ShadowQueue {      // PTC_LIGHTS_MIS: one entry per light sample awaiting its visibility test
    F4* ray_o;            // origin.xyz, slot of the path in the NEXT path queue (bits)
    F4* ray_d;            // direction.xyz, far end of the segment
    F4* contrib;          // radiance to add to that path's accumulated colour when nothing blocks the segment
};

struct Queue {            // SoA path state, one slot per live path
    F4* ray_o;            // origin.xyz, pixel index (bits)
    F4* ray_d;            // direction.xyz, rng state (bits)
    F4* mask;             // color_mask.rgb, -
    F4* acc;              // accumulated_color.rgb, -
};

struct Planes {
    F4* color;            // accumulation target (`output` storage texture)
    float* depth;         // depth_output
    F4* normal;           // normals_output
    F4* albedo;           // albedo_output
    uint32_t* hit_tri;    // primary-hit global triangle id (0xFFFFFFFF = miss)
    F4* spp_sum;          // SAMPLES_PER_PIXEL > 1 only: radiance summed over the samples of the frame so far ...
    uint32_t* spp_rng;    // ... and the pixel's RNG state after its latest sample
};

// part `part` of `num_parts`: its n queue slots are the warps t = part, part + num_parts, ... of the frame's slot order
void launch_raygen(cudaStream_t st, const ptc_uniforms& un, Queue q, FrameCounters* fc, uint32_t n, uint32_t num_parts, uint32_t part, const uint32_t* spp_rng = nullptr);
// impl 3 = packet traversal (primary rays in tile order); packet_stats (optional): [0] += node visits, [1] += lanes entering a child
void launch_trace(cudaStream_t st, int grid, int impl, const DevScene& s, Queue q, U4* hits, FrameCounters* fc, uint32_t segment,
                  unsigned long long* total_rays, unsigned long long* packet_stats = nullptr);
// sky.dir == nullptr: escaped paths are finished inside k_shade; else they are queued and launch_sky finishes them
void launch_shade(cudaStream_t st, int grid, const DevScene& s, const PathCfg& cfg, Queue qin, Queue qout, const U4* hits,
                  FrameCounters* fc, uint32_t segment, Planes planes, ShadowQueue sh, SkyQueue sky);
// the frame's escaped paths: sky lookup (sky.wgsl:1-12), radiance = acc + sky * mask, accumulation (pathtracer.wgsl:32-42)
void launch_sky(cudaStream_t st, int device, const DevScene& s, const PathCfg& cfg, SkyQueue sky, FrameCounters* fc, Planes planes);
// PTC_LIGHTS_MIS: any-hit traversal of the shadow rays shade(segment) queued; unoccluded ones add into acc_next
void launch_trace_shadow(cudaStream_t st, int grid, const DevScene& s, ShadowQueue sh, F4* acc_next, FrameCounters* fc, uint32_t segment,
                         unsigned long long* total_rays);
// standalone traversal of caller rays (tests/tools)
void launch_trace_n(cudaStream_t st, int grid, int impl, const DevScene& s, const F4* ray_o, const F4* ray_d, U4* hits, uint32_t n, uint32_t* cursor);
void launch_trace_any_n(cudaStream_t st, int grid, const DevScene& s, const F4* ray_o, const F4* ray_d, uint8_t* occluded, uint32_t n, uint32_t* cursor);

// display path: rayproject.wgsl (forward reprojection of the last real frame) and blit.wgsl (tonemap to the 8-bit surface)
void launch_reproject(cudaStream_t st, const ptc_camera& cam, const ptc_camera& prev, const float* depth, uint32_t* owner, const F4* src, F4* dst);
void launch_blit(cudaStream_t st, const F4* src, void* dst_rgba8, uint32_t n);
void launch_blit_part(cudaStream_t st, const F4* src, void* dst_rgba8, uint32_t n, uint32_t num_parts, uint32_t part, uint32_t W, uint32_t H);
void launch_unorm_probe(cudaStream_t st, unsigned long long tex_obj, float* out512);   // x and w channels of a 256x1 RGBA8 ramp
void launch_sample(cudaStream_t st, const DevScene& s, uint32_t tex, int sky, const float* in, float* out, uint32_t n, uint32_t tex_mod);
void launch_l2_read(cudaStream_t st, int device, const void* buf, uint64_t bytes, uint32_t iters, float* sink);   // L2 bandwidth probe
void launch_math(cudaStream_t st, int fn, const float* a, const float* b, float* out, uint32_t n);

int trace_grid_size(int device, int impl);   // impl 0 = k_trace (48 registers), 1 = k_trace_simple, 2 = k_trace held to 40 registers (HBM-resident BVHs)
int shade_grid_size(int device);

}  // namespace ptc
