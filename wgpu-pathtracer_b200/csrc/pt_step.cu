// pt_step.cu — the RTStep lifecycle behind the C ABI (include/ptcore.h).
// Replaces src/render_steps/pt.rs (create / update / render / resize) of the reference.
#include "../../include/ptcore.h"
#include "bvh_build.h"
#include "scene_prep.h"
#include "bvh_gpu.h"
#include "lbvh.cuh"
#include "pt_kernels.cuh"
#include "pt_error.h"

#include <algorithm>
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace ptc;

namespace ptc {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                           \
            return e_ == cudaErrorMemoryAllocation ? PTC_ERR_OOM : PTC_ERR_CUDA;                     \
        }                                                                                            \
    } while (0)

struct ptc_step {
    int device = 0;
    ptc_config cfg{};
    ptc_uniforms uniforms{};
    bool have_uniforms = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t snap_ready = nullptr, copy_done = nullptr;
    // Frame parts.  The queue slots of a frame (8x4-pixel tiles, one per warp) are dealt round-robin to `num_parts` parts;
    // each part is an independent wavefront (own queues, counters, stream) over its own pixels.  Persistent kernels leave
    // SMs idle while their longest rays finish (ncu: ~20 % of a trace launch's SM cycles); with several parts in flight
    // the block scheduler fills those slots with the other parts' kernels (measured: +8 % with 2 parts, +14 % with 3).
    // Parts own disjoint pixels, every pixel is still written exactly once per frame and a part's stream keeps frames in
    // order, so the image is bit-identical to the one-part schedule.  `stream` is the step's ordering stream: it waits for
    // every part after each frame, and anything that touches the planes between frames (blit, reduce, reset, reads) runs
    // on it; a part's first kernel that touches the planes waits for such work (`ev_main`).
    // Within a part, raygen + the primary-ray traversal of frame f+1 depend on nothing frame f produces: they run on the
    // part's second, lower-priority stream with their own ray / hit buffers (`qp`, `d_hits_p`) while frame f is still bouncing.
    struct FramePart {
        cudaStream_t stream = nullptr, prefetch_stream = nullptr;
        cudaEvent_t done = nullptr, primary_done = nullptr, primary_free = nullptr;
        Queue q[2]{}; U4* d_hits = nullptr; ShadowQueue shq{};
        Queue qp{}; U4* d_hits_p = nullptr;
        SkyQueue sky{};                            // escaped paths of the frame in flight (deferred environment lookup, k_sky)
        uint32_t n = 0;                            // queue slots (= pixels) of this part
    };
    bool overlap = true;
    static const uint32_t kMaxParts = 8;
    FramePart parts[kMaxParts + 1];              // [kMaxParts]: the whole frame as ONE part, allocated when stage timing is first enabled
    uint32_t num_parts = 1, want_parts = 0;      // want_parts: PTC_PARTS override, 0 = by frame size
    cudaEvent_t ev_main = nullptr; bool main_dirty = false;

    // scene (device)
    F4* d_nodes = nullptr; F4* d_tris = nullptr; TriShade* d_shade = nullptr;
    MatDev* d_mats = nullptr; TexDev* d_texs = nullptr; ptc_triangle* d_lights = nullptr;
    uint32_t* d_light_of_tri = nullptr;            // PTC_LIGHTS_MIS only
    std::vector<cudaArray_t> pages; std::vector<cudaTextureObject_t> page_objs;   // texture atlas pages
    uint32_t num_textures = 0; uint64_t atlas_bytes = 0;
    DevScene scene{};
    uint32_t num_lights = 0;
    // host tables for HIT_ID conversion
    std::vector<uint32_t> tri_mesh; std::vector<ptc_mesh> meshes;
    uint64_t bvh_info[8] = {0};

    // frame buffers
    FrameCounters* d_fc = nullptr;                 // ring of kFcRing x kMaxParts counter blocks
    uint32_t fc_next = 0;
    unsigned long long* d_total_rays = nullptr;
    Planes planes{}; F4* owned_color = nullptr; F4* d_snapshot = nullptr; size_t snapshot_bytes = 0;
    // display path (allocated on first use): tonemapped RGBA8 image, reprojected rgba32f image, splat ownership
    void* d_display = nullptr; F4* d_reprojected = nullptr; uint32_t* d_owner = nullptr;
    // ptc_reduce receives into its own plane (REDUCED), so the per-rank accumulation is never overwritten and a later
    // reduce — after more frames — sums each rank's whole history exactly once.  Caller-injectable like COLOR.
    F4* owned_reduced = nullptr; F4* reduced = nullptr;
    bool invalid = false;                          // a failed ptc_resize left the step without frame buffers
    cudaEvent_t ev_display_free = nullptr; bool display_busy = false;   // a read_async snapshot of DISPLAY is pending on the ordering stream
    bool blit_by_parts = true;                     // PTC_BLIT_PARTS=0: always tonemap on the ordering stream (A/B timing)
    bool defer_sky = true;                         // PTC_DEFER_SKY=0: escaped paths are finished inside k_shade (A/B timing)
    uint32_t npix = 0;
    int trace_grid = 0, shade_grid = 0;
    // Primary rays: packet traversal (k_trace_packet) pays off when the 32 rays of a tile stay together — 1.3-1.45x on
    // Sponza / San Miguel, where ~24 of 32 lanes enter a child at the average node visit, 0.7-0.86x on classroom / dragon
    // (~20 lanes: triangles smaller than a tile).  The kernel counts exactly that; the step samples the counters with an
    // asynchronous read-back every kPacketSample frames and keeps packets on while the coherence stays above kPacketMinLanes,
    // probing with one packet frame per sample period while they are off.  Either kernel produces the same hits.  A sample is
    // applied exactly kPacketLag frames after it was taken (waiting for the copy if it ever were that late), so which kernel a
    // frame uses depends on the frame sequence only, not on how fast the host thread runs.
    int packets_mode = -1;                         // PTC_PRIMARY_PACKETS: 0 never, 1 always, unset (-1) adaptive
    bool primary_packets = true;
    unsigned long long* h_packet_stats = nullptr;  // pinned: [0] node visits, [1] lane visits at the last sample
    unsigned long long packet_prev[2] = {0, 0};
    cudaEvent_t packet_ev = nullptr; bool packet_pending = false;
    uint64_t packet_frame = 0, packet_sample_frame = 0;
    int trace_impl = 0;                            // 0 = k_trace, 1 = k_trace_simple (PTC_TRACE_IMPL=simple, A/B timing only)
    uint64_t frames = 0, launches = 0;
    // optional per-stage timing (CUDA events on the render stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
    std::vector<int> ev_tag;                       // tag of the stage that STARTS at event i (-1 = end of frame)
    double prof_ms[3] = {0, 0, 0}; uint64_t prof_n[3] = {0, 0, 0};   // 0 raygen, 1 trace, 2 shade
};

static bool plane_ptr(ptc_step* s, int plane, void** p, size_t* px_bytes) {
    if (s->invalid) return false;
    switch (plane) {
        case PTC_PLANE_COLOR: *p = s->planes.color; *px_bytes = 16; return true;
        case PTC_PLANE_DEPTH: *p = s->planes.depth; *px_bytes = 4; return true;
        case PTC_PLANE_NORMAL: *p = s->planes.normal; *px_bytes = 16; return true;
        case PTC_PLANE_ALBEDO: *p = s->planes.albedo; *px_bytes = 16; return true;
        case PTC_PLANE_HIT_ID: *p = s->planes.hit_tri; *px_bytes = 4; return true;
        case PTC_PLANE_DISPLAY: *p = s->d_display; *px_bytes = 4; return s->d_display != nullptr;
        case PTC_PLANE_REPROJECTED: *p = s->d_reprojected; *px_bytes = 16; return s->d_reprojected != nullptr;
        case PTC_PLANE_REDUCED: *p = s->reduced; *px_bytes = 16; return s->reduced != nullptr;
    }
    return false;
}

static void prof_mark(ptc_step* s, int tag) {
    if (!s->profiling) return;
    if (s->ev_used == s->ev_pool.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; s->ev_pool.push_back(e); }
    cudaEventRecord(s->ev_pool[s->ev_used++], s->stream);
    s->ev_tag.push_back(tag);
}
static void prof_collect(ptc_step* s) {
    if (s->ev_used == 0) return;
    cudaStreamSynchronize(s->stream);
    for (size_t i = 0; i + 1 < s->ev_used; ++i) {
        int tag = s->ev_tag[i];
        if (tag < 0) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, s->ev_pool[i], s->ev_pool[i + 1]) == cudaSuccess) {
            s->prof_ms[tag] += ms; s->prof_n[tag]++;
            if (getenv("PTC_DEBUG_LAUNCH_TIMES")) fprintf(stderr, "[ptc] stage %d: %.3f ms\n", tag, ms);
        }
    }
    s->ev_used = 0; s->ev_tag.clear();
}

static const uint32_t kFcRing = 64;

// Reads a 256x1 RGBA8 ramp through a cudaReadModeNormalizedFloat texture object: out[0..255] = .x, out[256..511] = .w.
static int unorm8_probe(float out[512]) {
    uchar4 ramp[256];
    for (int i = 0; i < 256; ++i) ramp[i] = make_uchar4((unsigned char)i, 0, 0, (unsigned char)(255 - i));
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
    cudaArray_t arr = nullptr; cudaTextureObject_t obj = 0; float* d_out = nullptr;
    int rc = PTC_OK; cudaError_t launch_err = cudaSuccess;
    do {
        if (cudaMallocArray(&arr, &cd, 256, 1) != cudaSuccess || cudaMalloc(&d_out, 512 * sizeof(float)) != cudaSuccess) { rc = PTC_ERR_OOM; break; }
        if (cudaMemcpy2DToArray(arr, 0, 0, ramp, 256 * 4, 256 * 4, 1, cudaMemcpyHostToDevice) != cudaSuccess) { rc = PTC_ERR_CUDA; break; }
        cudaResourceDesc rd; std::memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td; std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
        if (cudaCreateTextureObject(&obj, &rd, &td, nullptr) != cudaSuccess) { rc = PTC_ERR_CUDA; break; }
        launch_unorm_probe(nullptr, (unsigned long long)obj, d_out);
        if ((launch_err = cudaGetLastError()) != cudaSuccess) { rc = PTC_ERR_CUDA; break; }   // e.g. cudaErrorNoKernelImageForDevice on a non-sm_100a GPU
        if (cudaMemcpy(out, d_out, 512 * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = PTC_ERR_CUDA; break; }
    } while (0);
    if (rc != PTC_OK) { set_error(std::string("unorm8 probe: ") + cudaGetErrorString(launch_err != cudaSuccess ? launch_err : cudaGetLastError())); }
    if (obj) cudaDestroyTextureObject(obj);
    if (arr) cudaFreeArray(arr);
    cudaFree(d_out);
    return rc;
}

// 1 iff the texture unit converts every u8 value to exactly rn(v / 255.0f) — the value texture.rs:241 (`to_rgba32f`)
// produces and the oracle uses.  Evaluated once per process on the device being created; decides the device
// encoding of 8-bit textures (PTC_DEVTEX_UNORM8 vs read-bytes-and-divide).  PTC_UNORM8=div forces the latter.
// Returns 1 / 0, or a negative ptc_status when the probe itself could not run (e.g. no kernel image for this GPU: the
// library is built for sm_100a only) — ptc_create reports that instead of failing at the first ptc_render.
static int unorm8_hw_exact(int device) {
    static std::mutex mu;
    static std::map<int, int> cached;              // per device: steps on different GPUs are created from different threads
    std::lock_guard<std::mutex> lock(mu);
    auto it = cached.find(device);
    if (it != cached.end()) return it->second;
    const char* e = getenv("PTC_UNORM8");
    if (e && std::strcmp(e, "div") == 0) { cached[device] = 0; return 0; }
    float out[512];
    int rc = unorm8_probe(out);
    if (rc != PTC_OK) return rc;                   // not cached: the error is reported, a later create may retry
    int ok = 1;
    for (int i = 0; i < 256; ++i)
        if (out[i] != pt_div((float)i, 255.0f) || out[256 + i] != pt_div((float)(255 - i), 255.0f)) { ok = 0; break; }
    cached[device] = ok;
    return ok;
}

static void free_frame_buffers(ptc_step* s) {
    for (uint32_t p = 0; p <= ptc_step::kMaxParts; ++p) {
        ptc_step::FramePart& P = s->parts[p];
        for (int i = 0; i < 2; ++i) {
            cudaFree(P.q[i].ray_o); cudaFree(P.q[i].ray_d); cudaFree(P.q[i].mask); cudaFree(P.q[i].acc);
            P.q[i] = Queue{};
        }
        cudaFree(P.d_hits); P.d_hits = nullptr;
        cudaFree(P.qp.ray_o); cudaFree(P.qp.ray_d); P.qp = Queue{}; cudaFree(P.d_hits_p); P.d_hits_p = nullptr;
        cudaFree(P.shq.ray_o); cudaFree(P.shq.ray_d); cudaFree(P.shq.contrib); P.shq = ShadowQueue{};
        cudaFree(P.sky.dir); cudaFree(P.sky.mask); cudaFree(P.sky.acc); P.sky = SkyQueue{};
        P.n = 0;
    }
    cudaFree(s->owned_color); s->owned_color = nullptr;
    cudaFree(s->planes.depth); cudaFree(s->planes.normal); cudaFree(s->planes.albedo); cudaFree(s->planes.hit_tri);
    cudaFree(s->planes.spp_sum); cudaFree(s->planes.spp_rng);
    cudaFree(s->d_snapshot); s->d_snapshot = nullptr; s->snapshot_bytes = 0;
    cudaFree(s->d_display); s->d_display = nullptr; cudaFree(s->d_reprojected); s->d_reprojected = nullptr; cudaFree(s->d_owner); s->d_owner = nullptr;
    cudaFree(s->owned_reduced); s->owned_reduced = nullptr; s->reduced = nullptr;
    s->planes = Planes{};
}

static int alloc_part(ptc_step* s, ptc_step::FramePart& P, uint32_t n) {
    P.n = n;
    const size_t m = std::max<size_t>(n, 1);
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&P.q[i].ray_o, m * 16)); CK(cudaMalloc(&P.q[i].ray_d, m * 16));
        CK(cudaMalloc(&P.q[i].mask, m * 16)); CK(cudaMalloc(&P.q[i].acc, m * 16));
    }
    CK(cudaMalloc(&P.d_hits, m * 16));
    CK(cudaMalloc(&P.qp.ray_o, m * 16)); CK(cudaMalloc(&P.qp.ray_d, m * 16)); CK(cudaMalloc(&P.d_hits_p, m * 16));
    if (s->cfg.light_sampling == PTC_LIGHTS_MIS) {
        CK(cudaMalloc(&P.shq.ray_o, m * 16)); CK(cudaMalloc(&P.shq.ray_d, m * 16)); CK(cudaMalloc(&P.shq.contrib, m * 16));
    }
    if (s->defer_sky && s->cfg.samples_per_pixel == 1) {      // every pixel's path escapes at most once per frame
        CK(cudaMalloc(&P.sky.dir, m * 16)); CK(cudaMalloc(&P.sky.mask, m * 16)); CK(cudaMalloc(&P.sky.acc, m * 16));
    }
    return PTC_OK;
}

static int alloc_frame_buffers(ptc_step* s, uint32_t w, uint32_t h) {
    if (w == 0 || h == 0 || (uint64_t)w * h > 0x7fffffffull) { set_error("bad frame size"); return PTC_ERR_INVALID; }
    uint32_t n = w * h;
    s->npix = n; s->cfg.width = w; s->cfg.height = h;
    // slots are dealt to the parts warp by warp: warp t (slots 32t .. 32t+31) belongs to part t % num_parts
    // Parts by frame size (measured, profiles/r01f): the tail a part hides is a fixed ~30-50 us per launch, so mid-size frames
    // gain most (1080p: 3 parts +8 % Sponza, +22 % classroom); at 4K launches are long and a third part only costs locality.
    // Round 2 (profiles/r02zc_ab_parts.log): a tree that lives in L1 / L2 altogether (Sponza: 6 MB of nodes + triangles) has short
    // traversal tails and is faster with 2 parts (+2.8 %), the 48-70 MB trees of classroom and dragon keep 3 (2 parts: -2 to -7 %).
    const bool small_tree = s->bvh_info[2] + s->bvh_info[3] < (16ull << 20);
    uint32_t want = s->want_parts ? s->want_parts : (n < (1u << 20) ? 2u : (n < (1u << 22) ? (small_tree ? 2u : 3u) : 2u));
    s->num_parts = std::max(1u, std::min(std::min(want, (uint32_t)ptc_step::kMaxParts), n / 8192u));
    const uint32_t nw = n >> 5, rem = n & 31u, NP = s->num_parts;
    for (uint32_t p = 0; p < NP; ++p) {
        int rc = alloc_part(s, s->parts[p], 32u * (nw > p ? (nw - p + NP - 1) / NP : 0u) + (nw % NP == p ? rem : 0u));
        if (rc != PTC_OK) return rc;
    }
    if (s->profiling) { int rc = alloc_part(s, s->parts[ptc_step::kMaxParts], n); if (rc != PTC_OK) return rc; }
    CK(cudaMalloc(&s->owned_color, (size_t)n * 16));
    CK(cudaMalloc(&s->planes.depth, (size_t)n * 4));
    CK(cudaMalloc(&s->planes.normal, (size_t)n * 16));
    CK(cudaMalloc(&s->planes.albedo, (size_t)n * 16));
    CK(cudaMalloc(&s->planes.hit_tri, (size_t)n * 4));
    if (s->cfg.samples_per_pixel > 1) { CK(cudaMalloc(&s->planes.spp_sum, (size_t)n * 16)); CK(cudaMalloc(&s->planes.spp_rng, (size_t)n * 4)); }
    s->planes.color = s->owned_color;
    CK(cudaMemsetAsync(s->owned_color, 0, (size_t)n * 16, s->stream));
    CK(cudaMemsetAsync(s->planes.depth, 0, (size_t)n * 4, s->stream));
    CK(cudaMemsetAsync(s->planes.normal, 0, (size_t)n * 16, s->stream));
    CK(cudaMemsetAsync(s->planes.albedo, 0, (size_t)n * 16, s->stream));
    CK(cudaMemsetAsync(s->planes.hit_tri, 0xFF, (size_t)n * 4, s->stream));
    s->main_dirty = true;
    return PTC_OK;
}

static int upload_scene(ptc_step* s, const ptc_scene_view* v) {
    PreparedScene ps;
    int rc = prepare_scene(v, s->cfg, ps, /*caller_builds_bvh=*/true);
    if (rc != PTC_OK) return rc;
    s->meshes = ps.meshes; s->tri_mesh = ps.tri_mesh;
    QuantGrid grid = ps.bvh.grid;
    size_t num_nodes = ps.bvh.nodes.size();
    if (!ps.tri_pos.empty()) {
        // PTC_BVH_LBVH_DEVICE: the tree is built by CUDA kernels from the triangle positions (bvh_gpu.cu)
        float* d_pos = nullptr;
        CK(cudaMalloc(&d_pos, ps.tri_pos.size() * 4));
        cudaError_t ce = cudaMemcpyAsync(d_pos, ps.tri_pos.data(), ps.tri_pos.size() * 4, cudaMemcpyHostToDevice, s->stream);
        GpuBvh g;
        rc = ce != cudaSuccess ? PTC_ERR_CUDA : (s->cfg.bvh_build == PTC_BVH_PLOC_DEVICE ? build_ploc_device(d_pos, ps.num_tris, s->stream, g, s->cfg.ploc_window ? (int)s->cfg.ploc_window : PTC_PLOC_RADIUS)
                                                                                                 : build_lbvh_device(d_pos, ps.num_tris, s->stream, g));
        cudaFree(d_pos);
        if (rc != PTC_OK) { if (ce != cudaSuccess) set_error(std::string("position upload: ") + cudaGetErrorString(ce)); return rc; }
        if (g.max_depth >= PTC_STACK_SIZE) {
            // degenerate input (tree deeper than the traversal stack): take the host SAH tree instead, as scene_prep does
            cudaFree(g.nodes); cudaFree(g.tris);
            ptc_config c2 = s->cfg; c2.bvh_build = PTC_BVH_SAH_HOST;
            ps = PreparedScene();
            if ((rc = prepare_scene(v, c2, ps, false)) != PTC_OK) return rc;
            grid = ps.bvh.grid; num_nodes = ps.bvh.nodes.size();
        } else {
            s->d_nodes = reinterpret_cast<F4*>(g.nodes); s->d_tris = reinterpret_cast<F4*>(g.tris);
            grid = g.grid; num_nodes = g.num_nodes;
            s->bvh_info[0] = g.num_nodes; s->bvh_info[1] = g.num_tris;
            s->bvh_info[2] = (uint64_t)g.num_nodes * sizeof(DevNode2); s->bvh_info[3] = (uint64_t)g.num_tris * sizeof(DevTri);
            s->bvh_info[4] = g.max_depth; s->bvh_info[5] = (uint64_t)(g.build_ms + 0.5f); s->bvh_info[6] = 1;   // [6]: 1 = built on the device
            s->bvh_info[7] = (uint64_t)(g.build_ms * 1000.0f);                                                    // [7]: build time in microseconds
        }
    }
    const DevBvh2& dev = ps.bvh;
    if (!s->d_nodes) {
        s->bvh_info[0] = dev.nodes.size(); s->bvh_info[1] = dev.tris.size();
        s->bvh_info[2] = dev.nodes.size() * sizeof(DevNode2); s->bvh_info[3] = dev.tris.size() * sizeof(DevTri);
        s->bvh_info[4] = ps.max_depth; s->bvh_info[5] = (uint64_t)ps.build_ms; s->bvh_info[7] = (uint64_t)(ps.build_ms * 1000.0);
    }
    if (!s->d_nodes && !dev.nodes.empty()) {
        CK(cudaMalloc(&s->d_nodes, dev.nodes.size() * sizeof(DevNode2)));
        CK(cudaMemcpy(s->d_nodes, dev.nodes.data(), dev.nodes.size() * sizeof(DevNode2), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&s->d_tris, dev.tris.size() * sizeof(DevTri)));
        CK(cudaMemcpy(s->d_tris, dev.tris.data(), dev.tris.size() * sizeof(DevTri), cudaMemcpyHostToDevice));
    }
    if (!ps.shade.empty()) {
        CK(cudaMalloc(&s->d_shade, ps.shade.size() * sizeof(TriShade)));
        CK(cudaMemcpy(s->d_shade, ps.shade.data(), ps.shade.size() * sizeof(TriShade), cudaMemcpyHostToDevice));
    }

    // textures (texture.rs:232-336): shelf-packed into atlas pages, one cudaArray + texture object per page
    std::vector<TexDev> texs(v->num_textures);
    {
        const int hw_unorm = unorm8_hw_exact(s->device);
        if (hw_unorm < 0) return hw_unorm;
        int max_w = 0, max_h = 0;
        CK(cudaDeviceGetAttribute(&max_w, cudaDevAttrMaxTexture2DWidth, s->device));
        CK(cudaDeviceGetAttribute(&max_h, cudaDevAttrMaxTexture2DHeight, s->device));
        const uint32_t lim_w = (uint32_t)std::min(max_w, 32768), lim_h = (uint32_t)std::min(max_h, 65536);
        // 8-bit textures stay 8-bit (read through the texture unit's exact unorm conversion); without that guarantee
        // they are expanded to f32 here with the host's rn(v / 255)
        std::vector<std::vector<float>> expanded(v->num_textures);
        struct Item { uint32_t idx, w, h; };
        std::vector<Item> group[2];
        for (uint64_t i = 0; i < v->num_textures; ++i) {
            const ptc_texture_desc& t = v->textures[i];
            if (t.width > lim_w || t.height > lim_h) { set_error("texture " + std::to_string(i) + " exceeds the device's 2D texture limits"); return PTC_ERR_INVALID; }
            TexDev& d = texs[i]; std::memset(&d, 0, sizeof(d));
            d.w = t.width; d.h = t.height; d.fmt = t.format;
            int g = t.format == PTC_TEXEL_RGBA8 ? 0 : 1;
            if (g == 0 && !hw_unorm) {
                const uint8_t* p = (const uint8_t*)t.texels; size_t n = (size_t)t.width * t.height * 4;
                expanded[i].resize(n);
                for (size_t k = 0; k < n; ++k) expanded[i][k] = pt_div((float)p[k], 255.0f);
                g = 1;
            }
            group[g].push_back(Item{(uint32_t)i, t.width, t.height});
        }
        for (int g = 0; g < 2; ++g) {
            std::vector<Item>& items = group[g];
            if (items.empty()) continue;
            std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.h > b.h; });
            uint64_t area = 0; uint32_t widest = 0;
            for (const Item& it : items) { area += (uint64_t)it.w * it.h; widest = std::max(widest, it.w); }
            uint32_t page_w = 4096;                       // roughly square pages: width = max(widest texture, ~sqrt(total area))
            while ((uint64_t)page_w * page_w < area && page_w < lim_w) page_w *= 2;
            page_w = std::min(std::max(page_w, widest), lim_w);
            const size_t texel_bytes = g == 0 ? 4 : 16;
            size_t k = 0;
            while (k < items.size()) {
                // one page: shelves of decreasing height, left to right
                if (s->pages.size() >= PTC_MAX_TEX_PAGES) { set_error("textures exceed the atlas capacity (PTC_MAX_TEX_PAGES pages)"); return PTC_ERR_INVALID; }
                const uint32_t page = (uint32_t)s->pages.size();
                uint32_t x = 0, y = 0, shelf_h = 0, used_w = 0; size_t first = k;
                for (; k < items.size(); ++k) {
                    Item& it = items[k];
                    if (x + it.w > page_w) { x = 0; y += shelf_h; shelf_h = 0; }
                    if (y + it.h > lim_h) break;
                    TexDev& d = texs[it.idx]; d.page = page; d.ax = x; d.ay = y;
                    x += it.w; shelf_h = std::max(shelf_h, it.h); used_w = std::max(used_w, x);
                }
                if (k == first) { set_error("atlas packing made no progress"); return PTC_ERR_INVALID; }
                const uint32_t used_h = y + shelf_h;
                cudaChannelFormatDesc cd = g == 0 ? cudaCreateChannelDesc<uchar4>() : cudaCreateChannelDesc<float4>();
                cudaArray_t arr = nullptr;
                CK(cudaMallocArray(&arr, &cd, used_w, used_h));
                s->pages.push_back(arr); s->page_objs.push_back(0);
                for (size_t j = first; j < k; ++j) {
                    const Item& it = items[j]; const TexDev& d = texs[it.idx];
                    const void* src = expanded[it.idx].empty() ? v->textures[it.idx].texels : (const void*)expanded[it.idx].data();
                    CK(cudaMemcpy2DToArray(arr, (size_t)d.ax * texel_bytes, d.ay, src, (size_t)it.w * texel_bytes, (size_t)it.w * texel_bytes, it.h, cudaMemcpyHostToDevice));
                }
                cudaResourceDesc rd; std::memset(&rd, 0, sizeof(rd));
                rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
                cudaTextureDesc td; std::memset(&td, 0, sizeof(td));
                td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;   // Repeat is applied in integer texel space (pt_bilinear_axis)
                td.filterMode = cudaFilterModePoint;                             // fp32 bilinear weights are applied in the kernel
                td.readMode = g == 0 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
                td.normalizedCoords = 0;
                CK(cudaCreateTextureObject(&s->page_objs[page], &rd, &td, nullptr));
                s->atlas_bytes += (uint64_t)used_w * used_h * texel_bytes;
            }
        }
    }
    s->num_textures = (uint32_t)texs.size();
    if (!texs.empty()) {
        CK(cudaMalloc(&s->d_texs, texs.size() * sizeof(TexDev)));
        CK(cudaMemcpy(s->d_texs, texs.data(), texs.size() * sizeof(TexDev), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&s->d_mats, ps.mats.size() * sizeof(MatDev)));
    CK(cudaMemcpy(s->d_mats, ps.mats.data(), ps.mats.size() * sizeof(MatDev), cudaMemcpyHostToDevice));
    s->num_lights = (uint32_t)ps.lights.size();
    CK(cudaMalloc(&s->d_lights, ps.lights.size() * sizeof(ptc_triangle)));
    CK(cudaMemcpy(s->d_lights, ps.lights.data(), ps.lights.size() * sizeof(ptc_triangle), cudaMemcpyHostToDevice));
    if (s->cfg.light_sampling == PTC_LIGHTS_MIS && !ps.light_of_tri.empty()) {
        CK(cudaMalloc(&s->d_light_of_tri, ps.light_of_tri.size() * 4));
        CK(cudaMemcpy(s->d_light_of_tri, ps.light_of_tri.data(), ps.light_of_tri.size() * 4, cudaMemcpyHostToDevice));
    }

    DevScene& ds = s->scene;
    ds.nodes = reinterpret_cast<const NodeQ*>(s->d_nodes); ds.tris = s->d_tris; ds.shade = s->d_shade; ds.mats = s->d_mats; ds.texs = s->d_texs; ds.lights = s->d_lights;
    ds.num_nodes = (uint32_t)num_nodes; ds.num_tris = ps.num_tris; ds.num_lights = s->num_lights; ds.env_tex = s->cfg.env_texture_index;
    ds.light_of_tri = s->d_light_of_tri; ds.num_real_lights = s->d_light_of_tri ? ps.num_real_lights : 0;
    ds.num_pages = (uint32_t)s->page_objs.size();
    for (int a = 0; a < 3; ++a) { ds.grid_base[a] = grid.base[a]; ds.grid_extent[a] = grid.extent[a]; }
    for (size_t p = 0; p < s->page_objs.size(); ++p) ds.page_obj[p] = (unsigned long long)s->page_objs[p];
    return PTC_OK;
}

static void destroy_step(ptc_step* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (auto& P : s->parts) { if (P.prefetch_stream) cudaStreamSynchronize(P.prefetch_stream); if (P.stream) cudaStreamSynchronize(P.stream); }
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    free_frame_buffers(s);
    for (auto o : s->page_objs) if (o) cudaDestroyTextureObject(o);
    for (auto a : s->pages) if (a) cudaFreeArray(a);
    cudaFree(s->d_nodes); cudaFree(s->d_tris); cudaFree(s->d_shade); cudaFree(s->d_mats); cudaFree(s->d_texs); cudaFree(s->d_lights); cudaFree(s->d_light_of_tri);
    cudaFree(s->d_fc); cudaFree(s->d_total_rays);
    if (s->h_packet_stats) cudaFreeHost(s->h_packet_stats);
    if (s->packet_ev) cudaEventDestroy(s->packet_ev);
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    if (s->snap_ready) cudaEventDestroy(s->snap_ready);
    if (s->copy_done) cudaEventDestroy(s->copy_done);
    if (s->ev_main) cudaEventDestroy(s->ev_main);
    if (s->ev_display_free) cudaEventDestroy(s->ev_display_free);
    for (auto& P : s->parts) {
        if (P.done) cudaEventDestroy(P.done);
        if (P.primary_done) cudaEventDestroy(P.primary_done);
        if (P.primary_free) cudaEventDestroy(P.primary_free);
        if (P.prefetch_stream) cudaStreamDestroy(P.prefetch_stream);
        if (P.stream) cudaStreamDestroy(P.stream);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

extern "C" {

PTC_API int ptc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

PTC_API int ptc_create(const ptc_scene_view* scene, const ptc_config* config, int cuda_device, ptc_step** out) {
    if (!scene || !config || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    *out = nullptr;
    if (config->samples_per_pixel == 0 || config->samples_per_pixel > 1024) { set_error("samples_per_pixel must be in 1..1024 (SAMPLES_PER_PIXEL, shader_definitions.rs:17)"); return PTC_ERR_INVALID; }
    if (config->max_bounces + 1 > PTC_MAX_SEGMENTS) { set_error("max_bounces too large"); return PTC_ERR_INVALID; }
    if (config->accum_mode > PTC_ACCUM_SUM) { set_error("bad accum_mode"); return PTC_ERR_INVALID; }
    if (config->light_sampling > PTC_LIGHTS_MIS) { set_error("bad light_sampling"); return PTC_ERR_INVALID; }
    if (config->bvh_build > PTC_BVH_PLOC_DEVICE) { set_error("bad bvh_build"); return PTC_ERR_INVALID; }
    // fields this build gives no meaning to must be zero, so a caller never believes an option took effect
    if (config->bvh_width != 0 && config->bvh_width != 2) { set_error("bad bvh_width (0 = library default, 2 = BVH2)"); return PTC_ERR_INVALID; }
    if (config->compat_flags != PTC_COMPAT_NONE) { set_error("unknown compat_flags (none are defined)"); return PTC_ERR_INVALID; }
    for (uint32_t r : config->reserved) if (r != 0) { set_error("ptc_config.reserved must be zero"); return PTC_ERR_INVALID; }
    if (config->ploc_window > 64u || (config->ploc_window != 0u && config->bvh_build != PTC_BVH_PLOC_DEVICE)) { set_error("bad ploc_window (0 = default, at most 64, PTC_BVH_PLOC_DEVICE only)"); return PTC_ERR_INVALID; }
    if (config->light_sampling == PTC_LIGHTS_MIS && config->send_to_light_probability != 0.0f) { set_error("PTC_LIGHTS_MIS requires send_to_light_probability == 0"); return PTC_ERR_INVALID; }
    int ndev = ptc_device_count();
    if (ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return PTC_ERR_NO_DEVICE; }
    if (cuda_device < 0 || cuda_device >= ndev) { set_error("cuda_device out of range"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(cuda_device));
    ptc_step* s = new (std::nothrow) ptc_step();
    if (!s) { set_error("ptc_create: out of host memory"); return PTC_ERR_OOM; }
    s->device = cuda_device; s->cfg = *config;
    int rc = PTC_OK;
    try { do {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);         // numerically lower = higher priority
        { const char* e = getenv("PTC_PARTS"); if (e && atoi(e) >= 1) s->want_parts = (uint32_t)atoi(e); }
        bool parts_ok = true;
        for (uint32_t pi = 0; pi < ptc_step::kMaxParts; ++pi) {
            ptc_step::FramePart& P = s->parts[pi];
            parts_ok = parts_ok && cudaStreamCreateWithPriority(&P.stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
                       cudaStreamCreateWithPriority(&P.prefetch_stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                       cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming) == cudaSuccess &&
                       cudaEventCreateWithFlags(&P.primary_done, cudaEventDisableTiming) == cudaSuccess &&
                       cudaEventCreateWithFlags(&P.primary_free, cudaEventDisableTiming) == cudaSuccess;
        }
        { const char* e = getenv("PTC_OVERLAP"); s->overlap = !(e && std::strcmp(e, "0") == 0); }
        { const char* e = getenv("PTC_BLIT_PARTS"); s->blit_by_parts = !(e && std::strcmp(e, "0") == 0); }
        { const char* e = getenv("PTC_DEFER_SKY"); s->defer_sky = !(e && std::strcmp(e, "0") == 0); }
        if (!parts_ok) { set_error("stream/event creation failed"); rc = PTC_ERR_CUDA; break; }
        if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_main, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_display_free, cudaEventDisableTiming) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->snap_ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming) != cudaSuccess) { set_error("stream/event creation failed"); rc = PTC_ERR_CUDA; break; }
        if ((rc = upload_scene(s, scene)) != PTC_OK) break;
        if ((rc = alloc_frame_buffers(s, config->width, config->height)) != PTC_OK) break;
        if (cudaMalloc(&s->d_fc, sizeof(FrameCounters) * kFcRing * ptc_step::kMaxParts) != cudaSuccess || cudaMalloc(&s->d_total_rays, 24) != cudaSuccess) { set_error("counter allocation failed"); rc = PTC_ERR_OOM; break; }
        cudaMemsetAsync(s->d_total_rays, 0, 24, s->stream);     // [0] rays traced, [1] packet node visits, [2] lanes entering a child at them
        {   // traversal kernel variant: by BVH size (pt_kernels.cu), PTC_TRACE_IMPL = simple | std | dense overrides (A/B timing)
            const char* e = getenv("PTC_TRACE_IMPL");
            s->trace_impl = (s->bvh_info[2] + s->bvh_info[3]) > (96ull << 20) ? 2 : 0;
            if (e && std::strcmp(e, "simple") == 0) s->trace_impl = 1;
            if (e && std::strcmp(e, "std") == 0) s->trace_impl = 0;
            if (e && std::strcmp(e, "dense") == 0) s->trace_impl = 2;
            if (e && std::strcmp(e, "postpone") == 0) s->trace_impl = 4;
        }
        { const char* e = getenv("PTC_PRIMARY_PACKETS"); s->packets_mode = e ? (std::strcmp(e, "0") == 0 ? 0 : 1) : -1; s->primary_packets = s->packets_mode != 0; }
        if (cudaMallocHost(&s->h_packet_stats, 16) != cudaSuccess || cudaEventCreateWithFlags(&s->packet_ev, cudaEventDisableTiming) != cudaSuccess) { set_error("packet statistics allocation failed"); rc = PTC_ERR_OOM; break; }
        s->trace_grid = trace_grid_size(cuda_device, s->trace_impl);
        s->shade_grid = shade_grid_size(cuda_device);
        {   // persistent-grid sizes in CTAs per SM (A/B timing): smaller grids leave room for another part's kernel on the same SM
            int sms = 1; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cuda_device);
            const char* e = getenv("PTC_TRACE_CTAS"); if (e && atoi(e) >= 1) s->trace_grid = std::min(s->trace_grid, sms * atoi(e));
            e = getenv("PTC_SHADE_CTAS"); if (e && atoi(e) >= 1) s->shade_grid = std::min(s->shade_grid, sms * atoi(e));
        }
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) { set_error("create: stream sync failed"); rc = PTC_ERR_CUDA; break; }
    } while (0); }
    // host-side allocations (scene preparation, BVH build) must not throw across the C ABI
    catch (const std::bad_alloc&) { set_error("ptc_create: out of host memory"); rc = PTC_ERR_OOM; }
    catch (const std::exception& e) { set_error(std::string("ptc_create: ") + e.what()); rc = PTC_ERR_INVALID; }
    if (rc != PTC_OK) { destroy_step(s); return rc; }
    *out = s;
    return PTC_OK;
}

PTC_API int ptc_update(ptc_step* s, const ptc_uniforms* u) {
    if (!s || !u) { set_error("null argument"); return PTC_ERR_INVALID; }
    s->uniforms = *u;
    s->uniforms.num_lights = s->num_lights;            // pt.rs:243
    s->have_uniforms = true;
    return PTC_OK;
}

PTC_API int ptc_render(ptc_step* s) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    if (!s->have_uniforms) { set_error("ptc_render before ptc_update"); return PTC_ERR_INVALID; }
    if (s->uniforms.camera.width != s->cfg.width || s->uniforms.camera.height != s->cfg.height) { set_error("uniform size differs from step size; call ptc_resize"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    PathCfg pc; pc.width = s->cfg.width; pc.height = s->cfg.height; pc.max_bounces = s->cfg.max_bounces;
    pc.send_to_light_probability = s->cfg.send_to_light_probability; pc.accum_mode = s->cfg.accum_mode; pc.is_grabbed = s->uniforms.is_grabbed;
    pc.light_sampling = s->cfg.light_sampling;
    const bool mis = pc.light_sampling == PTC_LIGHTS_MIS && s->scene.num_real_lights != 0;
    if (s->profiling && s->ev_used > 60000) prof_collect(s);
    // primary rays: packets or per-ray traversal, by the measured coherence (see ptc_step::packets_mode)
    bool use_packets = s->primary_packets && s->trace_impl != 1;
    if (s->packets_mode < 0 && s->trace_impl != 1) {
        static const uint64_t kPacketSample = 32, kPacketLag = 4; static const double kPacketMinLanes = 21.5;
        if (s->packet_pending && s->packet_frame >= s->packet_sample_frame + kPacketLag) {
            CK(cudaEventSynchronize(s->packet_ev));                               // four frames old: long complete
            s->packet_pending = false;
            const unsigned long long dn = s->h_packet_stats[0] - s->packet_prev[0], dl = s->h_packet_stats[1] - s->packet_prev[1];
            s->packet_prev[0] = s->h_packet_stats[0]; s->packet_prev[1] = s->h_packet_stats[1];
            if (dn) s->primary_packets = (double)dl / (double)dn >= kPacketMinLanes;
        }
        const bool probe = s->packet_frame % kPacketSample == 0;             // frame 0, 32, 64, ...: always a packet frame, then sampled
        use_packets = s->primary_packets || probe;
    }
    // work enqueued on the ordering stream since the last frame (blit, reduce, reset, an asynchronous read's snapshot) must
    // be over before this frame touches the planes; raygen and the primary-ray traversal do not, and are not held back
    const bool wait_main = s->main_dirty && !s->profiling;
    if (wait_main) CK(cudaEventRecord(s->ev_main, s->stream));
    s->main_dirty = false;
    // stage timing renders the frame as ONE part on the ordering stream, so that the per-stage CUDA-event times are exclusive
    // and each launch covers a whole segment of the frame (same image: pixels do not care which part they are in)
    const uint32_t nparts = s->profiling ? 1u : s->num_parts;
    const uint32_t spp = s->cfg.samples_per_pixel;
    pc.spp = spp;
    for (uint32_t p = 0; p < nparts; ++p) {
        ptc_step::FramePart& P = s->parts[s->profiling ? ptc_step::kMaxParts : p];
        cudaStream_t st = s->profiling ? s->stream : P.stream;
        // SAMPLES_PER_PIXEL > 1: sample k of a pixel starts from the RNG state sample k-1 ended with, so the samples of a part
        // run back to back on the part's stream (no primary-ray prefetch: its queue is reused by every sample)
        const bool prefetch = s->overlap && !s->profiling && spp == 1;
        cudaStream_t st0 = prefetch ? P.prefetch_stream : st;
        if (prefetch) CK(cudaStreamWaitEvent(st0, P.primary_free, 0));      // the previous frame's shade(0) has consumed qp / d_hits_p
        for (uint32_t sample = 0; sample < spp; ++sample) {
            pc.sample = sample;
            const uint32_t ring = s->fc_next++ % kFcRing;
            FrameCounters* fc = s->d_fc + ((size_t)ring * ptc_step::kMaxParts + p);
            CK(cudaMemsetAsync(fc, 0, sizeof(FrameCounters), st0));
            prof_mark(s, 0);
            launch_raygen(st0, s->uniforms, P.qp, fc, P.n, nparts, p, sample ? s->planes.spp_rng : nullptr);
            prof_mark(s, 1);
            // primary rays arrive in 8x4-pixel tiles, one per warp: packet traversal (pt_kernels.cu); PTC_PRIMARY_PACKETS=0 uses k_trace
            launch_trace(st0, s->trace_grid, use_packets ? 3 : s->trace_impl, s->scene, P.qp, P.d_hits_p, fc, 0, s->d_total_rays, s->d_total_rays + 1);
            if (prefetch) { CK(cudaEventRecord(P.primary_done, st0)); CK(cudaStreamWaitEvent(st, P.primary_done, 0)); }
            if (wait_main && sample == 0) CK(cudaStreamWaitEvent(st, s->ev_main, 0));
            s->launches += 2;
            for (uint32_t seg = 0; seg <= s->cfg.max_bounces; ++seg) {
                if (seg) { prof_mark(s, 1); launch_trace(st, s->trace_grid, s->trace_impl, s->scene, P.q[seg & 1], P.d_hits, fc, seg, s->d_total_rays); s->launches++; }
                prof_mark(s, 2);
                launch_shade(st, s->shade_grid, s->scene, pc, seg ? P.q[seg & 1] : P.qp, P.q[(seg + 1) & 1], seg ? P.d_hits : P.d_hits_p, fc, seg, s->planes, P.shq, P.sky);
                if (seg == 0 && prefetch) CK(cudaEventRecord(P.primary_free, st));
                s->launches++;
                // PTC_LIGHTS_MIS: the shadow rays shade(b) queued are traced (any-hit) before shade(b+1) reads the accumulated colours
                if (mis && seg < s->cfg.max_bounces) { prof_mark(s, 1); launch_trace_shadow(st, s->trace_grid, s->scene, P.shq, P.q[(seg + 1) & 1].acc, fc, seg, s->d_total_rays); s->launches++; }
            }
            // the frame's escaped paths: one dense environment pass (counted with the shade stage)
            if (P.sky.dir) { prof_mark(s, 2); launch_sky(st, s->device, s->scene, pc, P.sky, fc, s->planes); s->launches++; }
        }
        if (!s->profiling) { CK(cudaEventRecord(P.done, st)); CK(cudaStreamWaitEvent(s->stream, P.done, 0)); }
    }
    if (s->packets_mode < 0 && s->trace_impl != 1) {
        // sample the coherence counters after a probing frame: an asynchronous read-back behind this frame on the ordering stream
        if (s->packet_frame % 32 == 0 && !s->packet_pending) {
            CK(cudaMemcpyAsync(s->h_packet_stats, s->d_total_rays + 1, 16, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaEventRecord(s->packet_ev, s->stream));
            s->packet_pending = true; s->packet_sample_frame = s->packet_frame;
        }
        s->packet_frame++;
    }
    prof_mark(s, -1);
    if (s->profiling) s->main_dirty = true;           // the next unprofiled frame's parts must wait for this one
    CK(cudaGetLastError());
    s->frames++;
    return PTC_OK;
}

PTC_API int ptc_poll(ptc_step* s, int wait) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    if (wait) { CK(cudaStreamSynchronize(s->stream)); CK(cudaStreamSynchronize(s->copy_stream)); return PTC_OK; }
    cudaError_t e = cudaStreamQuery(s->stream);
    if (e == cudaErrorNotReady) return PTC_ERR_NOT_READY;
    CK(e);
    e = cudaStreamQuery(s->copy_stream);
    if (e == cudaErrorNotReady) return PTC_ERR_NOT_READY;
    CK(e);
    return PTC_OK;
}

PTC_API int ptc_resize(ptc_step* s, uint32_t w, uint32_t h) {
    return ptc::guarded("ptc_resize", [&]() -> int {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    // validate BEFORE anything is freed: a bad size leaves the step exactly as it was
    if (w == 0 || h == 0 || (uint64_t)w * h > 0x7fffffffull) { set_error("bad frame size"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    for (auto& P : s->parts) { if (P.prefetch_stream) CK(cudaStreamSynchronize(P.prefetch_stream)); if (P.stream) CK(cudaStreamSynchronize(P.stream)); }
    CK(cudaStreamSynchronize(s->stream)); CK(cudaStreamSynchronize(s->copy_stream));
    const uint32_t old_w = s->cfg.width, old_h = s->cfg.height;
    free_frame_buffers(s);                       // drops buffers injected with ptc_set_output / ptc_set_reduce_output too (their size is the old one)
    int rc = alloc_frame_buffers(s, w, h);
    if (rc != PTC_OK) {
        // out of memory midway: go back to the old size (which fitted a moment ago); if even that fails the step is marked
        // invalid and every entry point that touches the planes reports it instead of dereferencing null buffers
        const std::string why = g_last_error;
        free_frame_buffers(s);
        if (alloc_frame_buffers(s, old_w, old_h) != PTC_OK) { free_frame_buffers(s); s->invalid = true; s->npix = 0; }
        cudaStreamSynchronize(s->stream);
        set_error("ptc_resize: " + why + (s->invalid ? " (and the previous size could not be restored: step invalid)" : " (previous size restored, accumulation cleared)"));
        return rc;
    }
    s->invalid = false;
    CK(cudaStreamSynchronize(s->stream));
    return PTC_OK;
    });
}

PTC_API int ptc_set_output(ptc_step* s, void* device_rgba32f) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->planes.color = device_rgba32f ? (F4*)device_rgba32f : s->owned_color;
    return PTC_OK;
}

PTC_API int ptc_reset_accumulation(ptc_step* s) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    CK(cudaMemsetAsync(s->planes.color, 0, (size_t)s->npix * 16, s->stream));
    s->main_dirty = true;
    return PTC_OK;
}

PTC_API int ptc_blit(ptc_step* s, const void* src_device_rgba32f) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    if (!s->d_display) CK(cudaMalloc(&s->d_display, (size_t)s->npix * 4));
    if (!src_device_rgba32f && !s->main_dirty && !s->profiling && s->frames > 0 && s->blit_by_parts) {
        // Presenting the accumulated image right after a frame: every frame part tonemaps ITS pixels on its own stream, behind
        // its own last shade — the parts never wait for each other and the next frame is not held back (the reference runs
        // BlitStep on its second queue while the RT queue renders on, wgpu_util.rs:111-114).  The DISPLAY plane may still be
        // feeding the previous ptc_read_async snapshot: that copy (ordering stream) is awaited first.
        if (s->display_busy) for (uint32_t p = 0; p < s->num_parts; ++p) CK(cudaStreamWaitEvent(s->parts[p].stream, s->ev_display_free, 0));
        s->display_busy = false;
        for (uint32_t p = 0; p < s->num_parts; ++p) {
            ptc_step::FramePart& P = s->parts[p];
            launch_blit_part(P.stream, s->planes.color, s->d_display, P.n, s->num_parts, p, s->cfg.width, s->cfg.height);
            CK(cudaEventRecord(P.done, P.stream)); CK(cudaStreamWaitEvent(s->stream, P.done, 0));
            s->launches += P.n ? 1 : 0;
        }
        CK(cudaGetLastError());
        return PTC_OK;
    }
    s->main_dirty = true;
    launch_blit(s->stream, src_device_rgba32f ? (const F4*)src_device_rgba32f : s->planes.color, s->d_display, s->npix);
    s->launches += 1;
    CK(cudaGetLastError());
    return PTC_OK;
}

// The analogue of denoise.rs's copy_texture_to_buffer (src/render_steps/denoise.rs:84-100): a device-to-device copy of one
// plane into a caller-owned buffer with the caller's row pitch (wgpu needs 256-byte-aligned rows there, pt.rs:470-472), after
// the frames already enqueued.
PTC_API int ptc_copy_plane(ptc_step* s, int plane, void* dst_device, size_t dst_pitch_bytes) {
    if (!s || !dst_device) { set_error("null argument"); return PTC_ERR_INVALID; }
    void* src; size_t pb;
    if (plane == PTC_PLANE_HIT_ID || !plane_ptr(s, plane, &src, &pb)) { set_error("plane not copyable (or not produced yet)"); return PTC_ERR_INVALID; }
    const size_t row = (size_t)s->cfg.width * pb;
    if (dst_pitch_bytes && dst_pitch_bytes < row) { set_error("ptc_copy_plane: pitch smaller than a row"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    if (plane != PTC_PLANE_DISPLAY && plane != PTC_PLANE_REPROJECTED && plane != PTC_PLANE_REDUCED) s->main_dirty = true;   // the next frame writes these planes
    CK(cudaMemcpy2DAsync(dst_device, dst_pitch_bytes ? dst_pitch_bytes : row, src, row, row, s->cfg.height, cudaMemcpyDeviceToDevice, s->stream));
    return PTC_OK;
}

PTC_API int ptc_rayproject(ptc_step* s, const ptc_uniforms* prev, const ptc_uniforms* cur, const void* src_device_rgba32f) {
    if (!s || !prev || !cur) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    if (prev->camera.width != s->cfg.width || prev->camera.height != s->cfg.height || cur->camera.width != s->cfg.width || cur->camera.height != s->cfg.height) {
        set_error("rayproject: camera size differs from step size"); return PTC_ERR_INVALID;
    }
    CK(cudaSetDevice(s->device));
    if (!s->d_reprojected) {
        CK(cudaMalloc(&s->d_reprojected, (size_t)s->npix * 16)); CK(cudaMalloc(&s->d_owner, (size_t)s->npix * 4));
        CK(cudaMemsetAsync(s->d_reprojected, 0, (size_t)s->npix * 16, s->stream)); CK(cudaMemsetAsync(s->d_owner, 0, (size_t)s->npix * 4, s->stream));
    }
    s->main_dirty = true;
    launch_reproject(s->stream, cur->camera, prev->camera, s->planes.depth, s->d_owner, src_device_rgba32f ? (const F4*)src_device_rgba32f : s->planes.color, s->d_reprojected);
    s->launches += 2;
    CK(cudaGetLastError());
    return PTC_OK;
}

PTC_API void ptc_destroy(ptc_step* s) { destroy_step(s); }

PTC_API int ptc_write(ptc_step* s, int plane, const void* src, size_t pitch) {
    if (!s || !src) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (plane != PTC_PLANE_COLOR) { set_error("ptc_write: only the accumulation plane (PTC_PLANE_COLOR) can be loaded"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    const size_t row = (size_t)s->cfg.width * 16;
    s->main_dirty = true;
    CK(cudaMemcpy2DAsync(s->planes.color, row, src, pitch ? pitch : row, row, s->cfg.height, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return PTC_OK;
}

PTC_API int ptc_read(ptc_step* s, int plane, void* dst, size_t pitch) {
    return ptc::guarded("ptc_read", [&]() -> int {
    if (!s || !dst) { set_error("null argument"); return PTC_ERR_INVALID; }
    void* src; size_t pb;
    if (!plane_ptr(s, plane, &src, &pb)) { set_error("unknown plane, or a display-path plane that has not been produced yet"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    const uint32_t W = s->cfg.width, H = s->cfg.height;
    if (plane == PTC_PLANE_HIT_ID) {
        // device keeps the global triangle id; (instance_index, geometry_index) = (mesh, triangle inside the mesh), pt.rs:409-427,527-542
        std::vector<uint32_t> tri((size_t)s->npix);
        CK(cudaMemcpyAsync(tri.data(), src, (size_t)s->npix * 4, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        size_t row = pitch ? pitch : (size_t)W * 8;
        for (uint32_t y = 0; y < H; ++y) {
            uint32_t* o = (uint32_t*)((char*)dst + (size_t)y * row);
            for (uint32_t x = 0; x < W; ++x) {
                uint32_t t = tri[(size_t)y * W + x];
                if (t == 0xFFFFFFFFu || t >= s->tri_mesh.size()) { o[x * 2] = o[x * 2 + 1] = 0xFFFFFFFFu; continue; }
                uint32_t m = s->tri_mesh[t];
                o[x * 2] = m; o[x * 2 + 1] = t - s->meshes[m].index_offset / 3;
            }
        }
        return PTC_OK;
    }
    size_t row = (size_t)W * pb;
    CK(cudaMemcpy2DAsync(dst, pitch ? pitch : row, src, row, row, H, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return PTC_OK;
    });
}

PTC_API int ptc_read_async(ptc_step* s, int plane, void* dst_pinned) {
    if (!s || !dst_pinned) { set_error("null argument"); return PTC_ERR_INVALID; }
    void* src; size_t pb;
    if (plane == PTC_PLANE_HIT_ID || !plane_ptr(s, plane, &src, &pb)) { set_error("plane not readable asynchronously (or not produced yet)"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    size_t bytes = (size_t)s->npix * pb;
    if (s->snapshot_bytes < bytes) {
        CK(cudaStreamSynchronize(s->copy_stream));
        cudaFree(s->d_snapshot); s->d_snapshot = nullptr; s->snapshot_bytes = 0;
        CK(cudaMalloc(&s->d_snapshot, bytes)); s->snapshot_bytes = bytes;
    }
    // the previous host copy must have drained the snapshot before it is overwritten
    const bool frame_plane = plane != PTC_PLANE_DISPLAY && plane != PTC_PLANE_REPROJECTED && plane != PTC_PLANE_REDUCED;
    if (frame_plane) s->main_dirty = true;         // the next frame writes this plane: it must wait for the snapshot copy
    CK(cudaStreamWaitEvent(s->stream, s->copy_done, 0));
    CK(cudaMemcpyAsync(s->d_snapshot, src, bytes, cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaEventRecord(s->snap_ready, s->stream));
    if (plane == PTC_PLANE_DISPLAY) { CK(cudaEventRecord(s->ev_display_free, s->stream)); s->display_busy = true; }   // the next per-part blit waits for this
    CK(cudaStreamWaitEvent(s->copy_stream, s->snap_ready, 0));
    CK(cudaMemcpyAsync(dst_pinned, s->d_snapshot, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->copy_done, s->copy_stream));
    return PTC_OK;
}

PTC_API int ptc_read_wait(ptc_step* s) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->copy_stream));
    return PTC_OK;
}

PTC_API void* ptc_device_ptr(ptc_step* s, int plane) {
    if (!s) return nullptr;
    void* p; size_t pb;
    return plane_ptr(s, plane, &p, &pb) ? p : nullptr;
}

PTC_API void* ptc_stream(ptc_step* s) { return s ? (void*)s->stream : nullptr; }

PTC_API int ptc_get_stats(ptc_step* s, ptc_stats* out) {
    if (!s || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    std::memset(out, 0, sizeof(*out));
    unsigned long long c[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(c, s->d_total_rays, 24, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const unsigned long long rays = c[0];
    out->packet_node_visits = c[1]; out->packet_lane_visits = c[2];
    out->rays_traced = rays; out->frames_rendered = s->frames; out->kernel_launches = s->launches; out->frame_parts = s->num_parts;
    return PTC_OK;
}

PTC_API int ptc_set_profiling(ptc_step* s, int enable) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    prof_collect(s);
    if (enable && s->parts[ptc_step::kMaxParts].n == 0) {
        CK(cudaStreamSynchronize(s->stream));
        int rc = alloc_part(s, s->parts[ptc_step::kMaxParts], s->npix);
        if (rc != PTC_OK) return rc;
    }
    s->profiling = enable != 0;
    for (int k = 0; k < 3; ++k) { s->prof_ms[k] = 0.0; s->prof_n[k] = 0; }
    return PTC_OK;
}

PTC_API int ptc_get_profile(ptc_step* s, double ms[3], uint64_t launches[3]) {
    if (!s || !ms || !launches) { set_error("null argument"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    prof_collect(s);
    for (int k = 0; k < 3; ++k) { ms[k] = s->prof_ms[k]; launches[k] = s->prof_n[k]; }
    return PTC_OK;
}

PTC_API int ptc_num_lights(ptc_step* s, uint32_t* out) {
    if (!s || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    *out = s->num_lights; return PTC_OK;
}

PTC_API int ptc_bvh_info(ptc_step* s, uint64_t out[8]) {
    if (!s || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    std::memcpy(out, s->bvh_info, sizeof(s->bvh_info)); return PTC_OK;
}

static int upload_rays(ptc_step* s, const float* rays, uint64_t n, F4** d_o, F4** d_d) {
    std::vector<F4> o(n), d(n);
    for (uint64_t i = 0; i < n; ++i) {
        o[i] = F4{rays[i * 8], rays[i * 8 + 1], rays[i * 8 + 2], rays[i * 8 + 3]};
        d[i] = F4{rays[i * 8 + 4], rays[i * 8 + 5], rays[i * 8 + 6], rays[i * 8 + 7]};
    }
    CK(cudaMalloc(d_o, n * 16)); CK(cudaMalloc(d_d, n * 16));
    CK(cudaMemcpyAsync(*d_o, o.data(), n * 16, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(*d_d, d.data(), n * 16, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return PTC_OK;
}

PTC_API int ptc_trace_rays(ptc_step* s, const float* rays, uint64_t n, uint32_t* hits, uint32_t* ids) {
    return ptc::guarded("ptc_trace_rays", [&]() -> int {
    if (!s || !rays || !hits) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (n == 0) return PTC_OK;
    if (n > 0x7fffffffull) { set_error("too many rays"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    F4 *d_o = nullptr, *d_d = nullptr; U4* d_h = nullptr; uint32_t* d_cur = nullptr;
    int rc = upload_rays(s, rays, n, &d_o, &d_d);
    if (rc == PTC_OK) {
        do {
            if (cudaMalloc(&d_h, n * 16) != cudaSuccess || cudaMalloc(&d_cur, 8) != cudaSuccess) { set_error("trace_rays: allocation failed"); rc = PTC_ERR_OOM; break; }
            cudaMemsetAsync(d_cur, 0, 8, s->stream);
            launch_trace_n(s->stream, s->trace_grid, s->trace_impl, s->scene, d_o, d_d, d_h, (uint32_t)n, d_cur);
            s->launches += 2;
            if (cudaMemcpyAsync(hits, d_h, n * 16, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) {
                set_error(std::string("trace_rays: ") + cudaGetErrorString(cudaGetLastError())); rc = PTC_ERR_CUDA; break;
            }
        } while (0);
    }
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_h); cudaFree(d_cur);
    if (rc != PTC_OK) return rc;
    if (ids) {
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t t = hits[i * 4 + 3];
            if (t == 0xFFFFFFFFu || t >= s->tri_mesh.size()) { ids[i * 2] = ids[i * 2 + 1] = 0xFFFFFFFFu; continue; }
            uint32_t m = s->tri_mesh[t]; ids[i * 2] = m; ids[i * 2 + 1] = t - s->meshes[m].index_offset / 3;
        }
    }
    for (uint64_t i = 0; i < n; ++i)
        if (hits[i * 4 + 3] == 0xFFFFFFFFu) hits[i * 4] = hits[i * 4 + 1] = hits[i * 4 + 2] = 0;   // miss: t,b1,b2 reported as 0
    return PTC_OK;
    });
}

PTC_API int ptc_trace_rays_any(ptc_step* s, const float* rays, uint64_t n, uint8_t* occluded) {
    return ptc::guarded("ptc_trace_rays_any", [&]() -> int {
    if (!s || !rays || !occluded) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (n == 0) return PTC_OK;
    if (n > 0x7fffffffull) { set_error("too many rays"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    F4 *d_o = nullptr, *d_d = nullptr; uint8_t* d_occ = nullptr; uint32_t* d_cur = nullptr;
    int rc = upload_rays(s, rays, n, &d_o, &d_d);
    if (rc == PTC_OK) {
        do {
            if (cudaMalloc(&d_occ, n) != cudaSuccess || cudaMalloc(&d_cur, 8) != cudaSuccess) { set_error("trace_rays_any: allocation failed"); rc = PTC_ERR_OOM; break; }
            cudaMemsetAsync(d_cur, 0, 8, s->stream);
            launch_trace_any_n(s->stream, s->trace_grid, s->scene, d_o, d_d, d_occ, (uint32_t)n, d_cur);
            s->launches += 1;
            if (cudaMemcpyAsync(occluded, d_occ, n, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) {
                set_error(std::string("trace_rays_any: ") + cudaGetErrorString(cudaGetLastError())); rc = PTC_ERR_CUDA; break;
            }
        } while (0);
    }
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_occ); cudaFree(d_cur);
    return rc;
    });
}

PTC_API int ptc_debug_math(int cuda_device, int fn, const float* a, const float* b, float* out, uint64_t n) {
    if (!a || !b || !out || n == 0 || n > 0x7fffffffull) { set_error("bad argument"); return PTC_ERR_INVALID; }
    if (ptc_device_count() <= 0) { set_error("no CUDA device: this library has no CPU path"); return PTC_ERR_NO_DEVICE; }
    CK(cudaSetDevice(cuda_device));
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    int rc = PTC_OK;
    if (cudaMalloc(&da, n * 4) != cudaSuccess || cudaMalloc(&db, n * 4) != cudaSuccess || cudaMalloc(&dout, n * 4) != cudaSuccess) { set_error("debug_math: allocation failed"); rc = PTC_ERR_OOM; }
    if (rc == PTC_OK) {
        cudaMemcpy(da, a, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b, n * 4, cudaMemcpyHostToDevice);
        launch_math(nullptr, fn, da, db, dout, (uint32_t)n);
        if (cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error(std::string("debug_math: ") + cudaGetErrorString(cudaGetLastError())); rc = PTC_ERR_CUDA; }
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return rc;
}

PTC_API int ptc_debug_sample(ptc_step* s, int texture_index, const float* in, uint64_t n, float* rgba, int per_lane) {
    if (!s || !in || !rgba || n == 0 || n > 0x7fffffffull) { set_error("bad argument"); return PTC_ERR_INVALID; }
    const int sky = texture_index < 0;
    if (!sky && (uint64_t)texture_index >= s->num_textures) { set_error("texture index out of range"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    const size_t in_bytes = n * (sky ? 3 : 2) * sizeof(float);
    float *d_in = nullptr, *d_out = nullptr;
    int rc = PTC_OK;
    if (cudaMalloc(&d_in, in_bytes) != cudaSuccess || cudaMalloc(&d_out, n * 16) != cudaSuccess) { set_error("debug_sample: allocation failed"); rc = PTC_ERR_OOM; }
    if (rc == PTC_OK) {
        cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, s->stream);
        launch_sample(s->stream, s->scene, sky ? 0u : (uint32_t)texture_index, sky, d_in, d_out, (uint32_t)n, (!sky && per_lane) ? s->num_textures : 0u);
        if (cudaMemcpyAsync(rgba, d_out, n * 16, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) {
            set_error(std::string("debug_sample: ") + cudaGetErrorString(cudaGetLastError())); rc = PTC_ERR_CUDA;
        }
    }
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

// ---- ptc_reduce: NCCL bound at run time (no link dependency; the communicator's own library is the one called) ---------
typedef int (*nccl_reduce_fn)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);   // ncclReduce(send, recv, count, dtype, op, root, comm, stream)
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);     // ncclAllReduce(send, recv, count, dtype, op, comm, stream)
typedef const char* (*nccl_errstr_fn)(int);
static nccl_reduce_fn g_nccl_reduce = nullptr;
static nccl_allreduce_fn g_nccl_allreduce = nullptr;
static nccl_errstr_fn g_nccl_errstr = nullptr;
static bool bind_nccl() {
    if (g_nccl_reduce) return true;
    // Only an NCCL that is ALREADY part of this process can own the caller's ncclComm_t: look the shared object up without
    // loading it (RTLD_NOLOAD), then the global symbol scope (an NCCL linked statically into the host program or into
    // torch).  Loading another copy — e.g. the system libnccl next to torch's bundled one — would hand the communicator
    // to a different NCCL instance, so there is no such fallback.
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = RTLD_DEFAULT;
    nccl_reduce_fn red = (nccl_reduce_fn)dlsym(h, "ncclReduce");
    if (!red) return false;
    g_nccl_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    g_nccl_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    g_nccl_reduce = red;
    return true;
}
PTC_API int ptc_set_reduce_output(ptc_step* s, void* device_rgba32f) {
    if (!s) { set_error("null step"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (device_rgba32f) s->reduced = (F4*)device_rgba32f;
    else s->reduced = s->owned_reduced;          // may be null: allocated by the next ptc_reduce
    return PTC_OK;
}
PTC_API int ptc_reduce(ptc_step* s, void* nccl_comm, int root) {
    if (!s || !nccl_comm) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (s->invalid) { set_error("step has no frame buffers (a ptc_resize failed)"); return PTC_ERR_INVALID; }
    // running means do not add: rank images are only summable as (sum of radiance, sample count)
    if (s->cfg.accum_mode != PTC_ACCUM_SUM) { set_error("ptc_reduce needs ptc_config.accum_mode == PTC_ACCUM_SUM (per-rank running means cannot be summed)"); return PTC_ERR_INVALID; }
    if (!bind_nccl()) { set_error("ptc_reduce: no NCCL in this process (neither a loaded libnccl.so.2 nor a global ncclReduce symbol); the caller's communicator must come from the NCCL the call goes to"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    if (!s->reduced) {
        if (!s->owned_reduced) CK(cudaMalloc(&s->owned_reduced, (size_t)s->npix * 16));
        s->reduced = s->owned_reduced;
    }
    const int kNcclFloat32 = 7, kNcclSum = 0;                     // nccl.h: ncclFloat32 = 7, ncclSum = 0 (stable since NCCL 2.0)
    s->main_dirty = true;                                         // the next frame must not write COLOR while NCCL reads it
    int r;
    if (root < 0) {                                               // every rank wants the image (each one displays): all-reduce
        if (!g_nccl_allreduce) { set_error("ptc_reduce: ncclAllReduce not found"); return PTC_ERR_INVALID; }
        r = g_nccl_allreduce(s->planes.color, s->reduced, (size_t)s->npix * 4, kNcclFloat32, kNcclSum, nccl_comm, s->stream);
    } else r = g_nccl_reduce(s->planes.color, s->reduced, (size_t)s->npix * 4, kNcclFloat32, kNcclSum, root, nccl_comm, s->stream);
    if (r != 0) { set_error(std::string("ncclReduce: ") + (g_nccl_errstr ? g_nccl_errstr(r) : "error")); return PTC_ERR_CUDA; }
    return PTC_OK;
}

PTC_API int ptc_debug_l2_bandwidth(int cuda_device, uint64_t bytes, uint32_t iters, double* gbs) {
    if (!gbs || bytes < 4096 || iters == 0) { set_error("bad argument"); return PTC_ERR_INVALID; }
    if (ptc_device_count() <= 0) { set_error("no CUDA device: this library has no CPU path"); return PTC_ERR_NO_DEVICE; }
    CK(cudaSetDevice(cuda_device));
    bytes &= ~(uint64_t)15;
    void* buf = nullptr; float* sink = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = PTC_OK; float ms = 0.0f;
    do {
        if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { rc = PTC_ERR_OOM; set_error("l2 probe: allocation failed"); break; }
        if (cudaMemset(buf, 0, bytes) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { rc = PTC_ERR_CUDA; set_error("l2 probe: setup failed"); break; }
        launch_l2_read(nullptr, cuda_device, buf, bytes, 2, sink);                 // warm-up: brings the buffer into L2
        cudaEventRecord(e0, nullptr);
        launch_l2_read(nullptr, cuda_device, buf, bytes, iters, sink);
        cudaEventRecord(e1, nullptr);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) { rc = PTC_ERR_CUDA; set_error("l2 probe: kernel failed"); break; }
        *gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return rc;
}

static uint64_t fnv1a(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
PTC_API int ptc_debug_bvh_hash(ptc_step* s, uint64_t out[2]) {
    return ptc::guarded("ptc_debug_bvh_hash", [&]() -> int {
    if (!s || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    CK(cudaSetDevice(s->device));
    std::vector<unsigned char> nodes((size_t)s->bvh_info[2]), tris((size_t)s->bvh_info[3]);
    if (!nodes.empty()) CK(cudaMemcpy(nodes.data(), s->d_nodes, nodes.size(), cudaMemcpyDeviceToHost));
    if (!tris.empty()) CK(cudaMemcpy(tris.data(), s->d_tris, tris.size(), cudaMemcpyDeviceToHost));
    out[0] = fnv1a(nodes.data(), nodes.size()); out[1] = fnv1a(tris.data(), tris.size());
    return PTC_OK;
    });
}

PTC_API int ptc_debug_unorm8(int cuda_device, float* out512) {
    if (!out512) { set_error("null argument"); return PTC_ERR_INVALID; }
    if (ptc_device_count() <= 0) { set_error("no CUDA device: this library has no CPU path"); return PTC_ERR_NO_DEVICE; }
    CK(cudaSetDevice(cuda_device));
    int rc = unorm8_probe(out512);
    if (rc != PTC_OK) return rc;
    const int hw = unorm8_hw_exact(cuda_device);
    return hw < 0 ? hw : (hw ? 1 : 0);
}

PTC_API const char* ptc_last_error(void) { return g_last_error.c_str(); }
PTC_API const char* ptc_version(void) { return "ptcore 0.1 (sm_100a)"; }

}  // extern "C"
