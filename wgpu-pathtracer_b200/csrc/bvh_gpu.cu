// bvh_gpu.cu — linear BVH build on the device: Morton keys -> radix sort (CUB) -> Karras radix tree ->
// bottom-up box fit -> leaf collapse (<= PTC_LBVH_MAX_LEAF triangles) -> quantised 32-byte nodes + leaf-ordered triangles, i.e. exactly
// the arrays k_trace consumes (bvh_build.h), without the scene's geometry making a second trip through the host.
// Per-element arithmetic lives in lbvh.cuh and is restated for the host in bvh_build.cpp (build_lbvh_host), which the
// CPU-only tests run and against which tests/test_gpu_parity.py compares these kernels' output word for word.
#include "bvh_gpu.h"
#include "lbvh.cuh"
#include "pt_error.h"
#include "../../include/ptcore_types.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace ptc {
namespace {

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(std::string("bvh_gpu: ") + cudaGetErrorString(e_)); rc = PTC_ERR_CUDA; goto done; } } while (0)

struct Box6 { float lo[3], hi[3]; };

// order-preserving float <-> uint map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2o(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float o2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// bounds[0..5] = scene lo/hi, bounds[6..11] = centroid lo/hi (ordered-uint encoded)
__global__ void k_bounds(const float* __restrict__ pos, uint32_t n, Box6* __restrict__ prim, uint32_t* bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, c[3];
    const bool on = i < n;
    if (on) {
        lbvh_tri_box(pos + (size_t)i * 9, lo, hi);
        for (int a = 0; a < 3; ++a) { prim[i].lo[a] = lo[a]; prim[i].hi[a] = hi[a]; }
    }
    for (int a = 0; a < 3; ++a) c[a] = 0.5f * (lo[a] + hi[a]);
    // warp reduction, then one atomic per warp and quantity
    for (int a = 0; a < 3; ++a) {
        float l = lo[a], h = hi[a], cl = on ? c[a] : INFINITY, ch = on ? c[a] : -INFINITY;
        for (int o = 16; o; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
            cl = fminf(cl, __shfl_xor_sync(0xffffffffu, cl, o)); ch = fmaxf(ch, __shfl_xor_sync(0xffffffffu, ch, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&bounds[a], f2o(l)); atomicMax(&bounds[3 + a], f2o(h));
            atomicMin(&bounds[6 + a], f2o(cl)); atomicMax(&bounds[9 + a], f2o(ch));
        }
    }
}

struct KeyFrame { float cmin[3], scale[3]; };

__global__ void k_keys(const float* __restrict__ pos, uint32_t n, KeyFrame kf, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = lbvh_key(pos + (size_t)i * 9, kf.cmin, kf.scale);
    ids[i] = i;
}

// one thread per inner node of the radix tree
__global__ void k_tree(const uint64_t* __restrict__ keys, int n, int* __restrict__ first, int* __restrict__ last, int* __restrict__ split,
                       int* __restrict__ node_parent, int* __restrict__ leaf_parent, uint32_t* __restrict__ kept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int f, l, s;
    lbvh_node(keys, n, i, f, l, s);
    first[i] = f; last[i] = l; split[i] = s;
    if (f == s) leaf_parent[s] = i; else node_parent[s] = i;
    if (s + 1 == l) leaf_parent[s + 1] = i; else node_parent[s + 1] = i;
    if (i == 0) node_parent[0] = -1;
    kept[i] = (uint32_t)(l - f + 1) > PTC_LBVH_MAX_LEAF ? 1u : 0u;
}

__device__ __forceinline__ Box6 ld_box_cg(const Box6* p) {
    Box6 b;
    for (int a = 0; a < 3; ++a) { b.lo[a] = __ldcg(&p->lo[a]); b.hi[a] = __ldcg(&p->hi[a]); }
    return b;
}

// bottom-up box fit: one thread per sorted primitive climbs; the second thread to reach a node computes its box
__global__ void k_fit(const uint32_t* __restrict__ ids, const Box6* __restrict__ prim, int n, const int* __restrict__ first, const int* __restrict__ last,
                      const int* __restrict__ split, const int* __restrict__ node_parent, const int* __restrict__ leaf_parent,
                      Box6* node_box, uint32_t* arrived, const uint32_t* __restrict__ kept, uint32_t* max_depth) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int cur = leaf_parent[k];
    // depth of the collapsed tree above this primitive = kept ancestors (the traversal stack must hold that many entries)
    uint32_t depth = 0;
    for (int a = cur; a >= 0; a = node_parent[a]) depth += kept[a];
    atomicMax(max_depth, depth);
    for (;;) {
        __threadfence();
        if (atomicAdd(&arrived[cur], 1u) == 0u) return;
        const int f = first[cur], l = last[cur], s = split[cur];
        const Box6 a = (f == s) ? prim[ids[s]] : ld_box_cg(&node_box[s]);
        const Box6 b = (s + 1 == l) ? prim[ids[s + 1]] : ld_box_cg(&node_box[s + 1]);
        Box6 u;
        for (int x = 0; x < 3; ++x) { u.lo[x] = fminf(a.lo[x], b.lo[x]); u.hi[x] = fmaxf(a.hi[x], b.hi[x]); }
        for (int x = 0; x < 3; ++x) { __stcg(&node_box[cur].lo[x], u.lo[x]); __stcg(&node_box[cur].hi[x], u.hi[x]); }
        if (cur == 0) return;
        cur = node_parent[cur];
    }
}

struct EmitFrame { float pad; float base[3]; double step[3]; };

__global__ void k_emit_nodes(int n, const uint32_t* __restrict__ ids, const Box6* __restrict__ prim, const Box6* __restrict__ node_box,
                             const int* __restrict__ first, const int* __restrict__ last, const int* __restrict__ split,
                             const uint32_t* __restrict__ kept, const uint32_t* __restrict__ new_id, EmitFrame ef, DevNode2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !kept[i]) return;
    const int f = first[i], l = last[i], s = split[i];
    DevNode2 d;
    const int cf[2] = {f, s + 1}, cl[2] = {s, l}, ci[2] = {s, s + 1};
    for (int c = 0; c < 2; ++c) {
        const uint32_t cnt = (uint32_t)(cl[c] - cf[c] + 1);
        const Box6 b = (cnt == 1) ? prim[ids[cf[c]]] : node_box[ci[c]];
        lbvh_put_box(&d.w[c * 4], b.lo, b.hi, ef.pad, ef.base, ef.step);
        d.w[c * 4 + 3] = cnt <= PTC_LBVH_MAX_LEAF ? ~(((uint32_t)cf[c] << 3) | (cnt - 1u)) : new_id[ci[c]];
    }
    out[new_id[i]] = d;
}

__global__ void k_emit_tris(uint32_t n, const uint32_t* __restrict__ ids, const float* __restrict__ pos, DevTri* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids[i];
    lbvh_put_tri(out[i].q, pos + (size_t)id * 9, id);
}

// ---- PLOC (lbvh.cuh): rounds of nearest-partner search, mutual merge and compaction over the Morton-sorted clusters --------
__global__ void k_ploc_init(const uint32_t* __restrict__ ids, const Box6* __restrict__ prim, uint32_t n, PlocBox* __restrict__ box, uint32_t* __restrict__ cur,
                            uint32_t* __restrict__ count, uint32_t* __restrict__ kc) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Box6 b = prim[ids[k]];
    PlocBox o; for (int a = 0; a < 3; ++a) { o.lo[a] = b.lo[a]; o.hi[a] = b.hi[a]; }
    box[k] = o; cur[k] = k; count[k] = 1u; kc[k] = 0u;
}
__global__ void k_ploc_nn(const PlocBox* __restrict__ box, const uint32_t* __restrict__ cur, int m, int* __restrict__ nn, int radius) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) nn[i] = ploc_nearest(box, cur, m, i, radius);
}
// merged[i] = 1 when cluster i is the LEFT partner of a mutual pair (it becomes the merged cluster); valid[i] = 0 for the right partner
__global__ void k_ploc_flags(const int* __restrict__ nn, int m, uint32_t* __restrict__ merged, uint32_t* __restrict__ valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    merged[i] = (mutual && i < j) ? 1u : 0u;
    valid[i] = (mutual && i > j) ? 0u : 1u;
}
__global__ void k_ploc_merge(const int* __restrict__ nn, int m, const uint32_t* __restrict__ merged, const uint32_t* __restrict__ valid,
                             const uint32_t* __restrict__ mpos, const uint32_t* __restrict__ vpos, uint32_t base, const uint32_t* __restrict__ cur,
                             uint32_t* __restrict__ nxt, PlocBox* box, uint32_t* left, uint32_t* right, uint32_t* count, uint32_t* kc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !valid[i]) return;
    if (merged[i]) {
        const uint32_t a = cur[i], b = cur[nn[i]], nd = base + mpos[i];
        const PlocBox ba = box[a], bb = box[b];
        PlocBox u; for (int x = 0; x < 3; ++x) { u.lo[x] = fminf(ba.lo[x], bb.lo[x]); u.hi[x] = fmaxf(ba.hi[x], bb.hi[x]); }
        const uint32_t cnt = count[a] + count[b];
        box[nd] = u; left[nd] = a; right[nd] = b; count[nd] = cnt;
        kc[nd] = kc[a] + kc[b] + (cnt > PTC_PLOC_MAX_LEAF ? 1u : 0u);       // emitted nodes in the subtree
        nxt[vpos[i]] = nd;
    } else nxt[vpos[i]] = cur[i];
}
// top-down over the nodes one round created (their parents, if any, were created in later rounds and are done): depth-first
// position of the children's first triangle, the number of emitted ancestors above them, and their node number in DEPTH-FIRST
// order — a child follows its parent directly, so a descent walks forward through memory (numbering the nodes in order of
// creation, i.e. level by level, cost 15-25 % of the traversal speed on the 890 MB San Miguel tree)
__global__ void k_ploc_down(uint32_t first, uint32_t end, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                            const uint32_t* __restrict__ count, const uint32_t* __restrict__ kc, uint32_t* offset, uint32_t* depth, uint32_t* new_id) {
    const uint32_t nd = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= end) return;
    const uint32_t a = left[nd], b = right[nd], k = count[nd] > PTC_PLOC_MAX_LEAF ? 1u : 0u;
    const uint32_t o = offset[nd], d = depth[nd] + k, id = new_id[nd];
    offset[a] = o; offset[b] = o + count[a];
    depth[a] = d; depth[b] = d;
    new_id[a] = id + 1u; new_id[b] = id + 1u + kc[a];
}
__global__ void k_ploc_depth(uint32_t n, const uint32_t* __restrict__ depth, uint32_t* max_depth) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t d = k < n ? depth[k] : 0u;
    for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, d);
}
__global__ void k_ploc_emit_nodes(uint32_t n, uint32_t total, const PlocBox* __restrict__ box, const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                  const uint32_t* __restrict__ count, const uint32_t* __restrict__ offset, const uint32_t* __restrict__ new_id, EmitFrame ef,
                                  DevNode2* __restrict__ out) {
    const uint32_t nd = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= total || count[nd] <= PTC_PLOC_MAX_LEAF) return;
    DevNode2 d;
    const uint32_t ch[2] = {left[nd], right[nd]};
    for (int c = 0; c < 2; ++c) {
        const PlocBox b = box[ch[c]];
        lbvh_put_box(&d.w[c * 4], b.lo, b.hi, ef.pad, ef.base, ef.step);
        const uint32_t cnt = count[ch[c]];
        d.w[c * 4 + 3] = cnt <= PTC_PLOC_MAX_LEAF ? ~((offset[ch[c]] << 3) | (cnt - 1u)) : new_id[ch[c]];
    }
    out[new_id[nd]] = d;
}
__global__ void k_ploc_emit_tris(uint32_t n, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ offset, const float* __restrict__ pos, DevTri* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = ids[k];
    lbvh_put_tri(out[offset[k]].q, pos + (size_t)id * 9, id);
}

}  // namespace

int build_lbvh_device(const float* d_pos, uint32_t ntri, cudaStream_t st, GpuBvh& out) {
    int rc = PTC_OK;
    const int n = (int)ntri;
    out = GpuBvh{};
    if (ntri <= PTC_LBVH_MAX_LEAF || ntri >= (1u << 28)) { set_error("build_lbvh_device: triangle count out of range"); return PTC_ERR_INVALID; }
    Box6 *prim = nullptr, *node_box = nullptr;
    uint32_t *bounds = nullptr, *ids = nullptr, *ids2 = nullptr, *kept = nullptr, *new_id = nullptr, *arrived = nullptr, *d_depth = nullptr;
    uint64_t *keys = nullptr, *keys2 = nullptr;
    int *first = nullptr, *last = nullptr, *split = nullptr, *node_parent = nullptr, *leaf_parent = nullptr;
    void* tmp = nullptr; size_t tmp_bytes = 0, tb2 = 0;
    DevNode2* nodes = nullptr; DevTri* tris = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint32_t hb[12], num_kept_last[2] = {0, 0}, depth = 0;
    const uint32_t T = 256, gb = (ntri + T - 1) / T;
    {
        CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
        CKB(cudaEventRecord(e0, st));
        CKB(cudaMalloc(&prim, sizeof(Box6) * (size_t)n)); CKB(cudaMalloc(&node_box, sizeof(Box6) * (size_t)n));
        CKB(cudaMalloc(&bounds, 48)); CKB(cudaMalloc(&d_depth, 4));
        CKB(cudaMalloc(&keys, 8 * (size_t)n)); CKB(cudaMalloc(&keys2, 8 * (size_t)n)); CKB(cudaMalloc(&ids, 4 * (size_t)n)); CKB(cudaMalloc(&ids2, 4 * (size_t)n));
        CKB(cudaMalloc(&first, 4 * (size_t)n)); CKB(cudaMalloc(&last, 4 * (size_t)n)); CKB(cudaMalloc(&split, 4 * (size_t)n));
        CKB(cudaMalloc(&node_parent, 4 * (size_t)n)); CKB(cudaMalloc(&leaf_parent, 4 * (size_t)n));
        CKB(cudaMalloc(&kept, 4 * (size_t)n)); CKB(cudaMalloc(&new_id, 4 * (size_t)n)); CKB(cudaMalloc(&arrived, 4 * (size_t)n));
        // 1. boxes and bounds
        uint32_t init[12];
        for (int a = 0; a < 3; ++a) { init[a] = init[6 + a] = 0xFFFFFFFFu; init[3 + a] = init[9 + a] = 0u; }
        CKB(cudaMemcpyAsync(bounds, init, 48, cudaMemcpyHostToDevice, st));
        k_bounds<<<gb, T, 0, st>>>(d_pos, ntri, prim, bounds);
        CKB(cudaMemcpyAsync(hb, bounds, 48, cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        float slo[3], shi[3];
        KeyFrame kf;
        for (int a = 0; a < 3; ++a) {
            slo[a] = o2f(hb[a]); shi[a] = o2f(hb[3 + a]);
            const float cl = o2f(hb[6 + a]), ch = o2f(hb[9 + a]), ext = ch - cl;
            kf.cmin[a] = cl; kf.scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f;
        }
        // 2. keys + stable radix sort (ties keep triangle-id order)
        k_keys<<<gb, T, 0, st>>>(d_pos, ntri, kf, keys, ids);
        CKB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, ids2, n, 0, 63, st));
        CKB(cub::DeviceScan::ExclusiveSum(nullptr, tb2, kept, new_id, n - 1, st));
        if (tb2 > tmp_bytes) tmp_bytes = tb2;
        CKB(cudaMalloc(&tmp, tmp_bytes));
        CKB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ids2, n, 0, 63, st));
        // 3. radix tree, collapse flags, compacted node numbering
        k_tree<<<(n - 1 + T - 1) / T, T, 0, st>>>(keys2, n, first, last, split, node_parent, leaf_parent, kept);
        CKB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, kept, new_id, n - 1, st));
        // 4. boxes bottom-up (+ depth of the collapsed tree)
        CKB(cudaMemsetAsync(arrived, 0, 4 * (size_t)n, st)); CKB(cudaMemsetAsync(d_depth, 0, 4, st));
        k_fit<<<gb, T, 0, st>>>(ids2, prim, n, first, last, split, node_parent, leaf_parent, node_box, arrived, kept, d_depth);
        CKB(cudaMemcpyAsync(&num_kept_last[0], new_id + (n - 2), 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(&num_kept_last[1], kept + (n - 2), 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(&depth, d_depth, 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        const uint32_t num_nodes = num_kept_last[0] + num_kept_last[1];
        // 5. emit
        LbvhFrame fr; lbvh_frame(slo, shi, fr);
        EmitFrame ef; ef.pad = fr.pad;
        for (int a = 0; a < 3; ++a) { ef.base[a] = fr.grid.base[a]; ef.step[a] = fr.step[a]; }
        CKB(cudaMalloc(&nodes, sizeof(DevNode2) * (size_t)num_nodes)); CKB(cudaMalloc(&tris, sizeof(DevTri) * (size_t)n));
        k_emit_nodes<<<(n - 1 + T - 1) / T, T, 0, st>>>(n, ids2, prim, node_box, first, last, split, kept, new_id, ef, nodes);
        k_emit_tris<<<gb, T, 0, st>>>(ntri, ids2, d_pos, tris);
        CKB(cudaEventRecord(e1, st));
        CKB(cudaStreamSynchronize(st));
        CKB(cudaGetLastError());
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1);
        out.nodes = nodes; out.tris = tris; nodes = nullptr; tris = nullptr;
        out.num_nodes = num_nodes; out.num_tris = ntri; out.grid = fr.grid; out.max_depth = depth + 1; out.build_ms = ms;
    }
done:
    cudaFree(prim); cudaFree(node_box); cudaFree(bounds); cudaFree(d_depth); cudaFree(keys); cudaFree(keys2); cudaFree(ids); cudaFree(ids2);
    cudaFree(first); cudaFree(last); cudaFree(split); cudaFree(node_parent); cudaFree(leaf_parent); cudaFree(kept); cudaFree(new_id); cudaFree(arrived);
    cudaFree(tmp); cudaFree(nodes); cudaFree(tris);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

int build_ploc_device(const float* d_pos, uint32_t ntri, cudaStream_t st, GpuBvh& out, int radius) {
    int rc = PTC_OK;
    const int n = (int)ntri;
    out = GpuBvh{};
    if (ntri <= 4 || ntri >= (1u << 28)) { set_error("build_ploc_device: triangle count out of range"); return PTC_ERR_INVALID; }
    Box6* prim = nullptr; PlocBox* box = nullptr;
    uint32_t *bounds = nullptr, *ids = nullptr, *ids2 = nullptr, *cur = nullptr, *nxt = nullptr, *merged = nullptr, *valid = nullptr, *mpos = nullptr, *vpos = nullptr;
    uint32_t *left = nullptr, *right = nullptr, *count = nullptr, *offset = nullptr, *depth = nullptr, *kc = nullptr, *new_id = nullptr, *d_depth = nullptr;
    uint64_t *keys = nullptr, *keys2 = nullptr;
    int* nn = nullptr;
    void* tmp = nullptr; size_t tmp_bytes = 0, tb2 = 0;
    DevNode2* nodes = nullptr; DevTri* tris = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    uint32_t hb[12], tail[2] = {0, 0}, dmax = 0;
    const uint32_t T = 256, gb = (ntri + T - 1) / T;
    const size_t total_cap = (size_t)2 * n - 1;
    std::vector<uint32_t> round_base;
    {
        CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
        CKB(cudaEventRecord(e0, st));
        CKB(cudaMalloc(&prim, sizeof(Box6) * (size_t)n)); CKB(cudaMalloc(&box, sizeof(PlocBox) * total_cap));
        CKB(cudaMalloc(&bounds, 48)); CKB(cudaMalloc(&d_depth, 4));
        CKB(cudaMalloc(&keys, 8 * (size_t)n)); CKB(cudaMalloc(&keys2, 8 * (size_t)n)); CKB(cudaMalloc(&ids, 4 * (size_t)n)); CKB(cudaMalloc(&ids2, 4 * (size_t)n));
        CKB(cudaMalloc(&cur, 4 * (size_t)n)); CKB(cudaMalloc(&nxt, 4 * (size_t)n)); CKB(cudaMalloc(&nn, 4 * (size_t)n));
        CKB(cudaMalloc(&merged, 4 * (size_t)n)); CKB(cudaMalloc(&valid, 4 * (size_t)n)); CKB(cudaMalloc(&mpos, 4 * (size_t)n)); CKB(cudaMalloc(&vpos, 4 * (size_t)n));
        CKB(cudaMalloc(&left, 4 * total_cap)); CKB(cudaMalloc(&right, 4 * total_cap)); CKB(cudaMalloc(&count, 4 * total_cap));
        CKB(cudaMalloc(&offset, 4 * total_cap)); CKB(cudaMalloc(&depth, 4 * total_cap));
        CKB(cudaMalloc(&kc, 4 * total_cap)); CKB(cudaMalloc(&new_id, 4 * total_cap));
        // 1. boxes, bounds, Morton keys, stable sort: as the linear BVH
        uint32_t init[12];
        for (int a = 0; a < 3; ++a) { init[a] = init[6 + a] = 0xFFFFFFFFu; init[3 + a] = init[9 + a] = 0u; }
        CKB(cudaMemcpyAsync(bounds, init, 48, cudaMemcpyHostToDevice, st));
        k_bounds<<<gb, T, 0, st>>>(d_pos, ntri, prim, bounds);
        CKB(cudaMemcpyAsync(hb, bounds, 48, cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        float slo[3], shi[3];
        KeyFrame kf;
        for (int a = 0; a < 3; ++a) {
            slo[a] = o2f(hb[a]); shi[a] = o2f(hb[3 + a]);
            const float cl = o2f(hb[6 + a]), ch = o2f(hb[9 + a]), ext = ch - cl;
            kf.cmin[a] = cl; kf.scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f;
        }
        k_keys<<<gb, T, 0, st>>>(d_pos, ntri, kf, keys, ids);
        CKB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, ids2, n, 0, 63, st));
        CKB(cub::DeviceScan::ExclusiveSum(nullptr, tb2, merged, mpos, n, st));
        if (tb2 > tmp_bytes) tmp_bytes = tb2;
        CKB(cudaMalloc(&tmp, tmp_bytes));
        CKB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, ids2, n, 0, 63, st));
        // 2. clustering rounds
        k_ploc_init<<<gb, T, 0, st>>>(ids2, prim, ntri, box, cur, count, kc);
        uint32_t next_node = ntri;
        int m = n;
        while (m > 1) {
            const uint32_t g = ((uint32_t)m + T - 1) / T;
            k_ploc_nn<<<g, T, 0, st>>>(box, cur, m, nn, radius);
            k_ploc_flags<<<g, T, 0, st>>>(nn, m, merged, valid);
            CKB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, merged, mpos, m, st));
            CKB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, valid, vpos, m, st));
            k_ploc_merge<<<g, T, 0, st>>>(nn, m, merged, valid, mpos, vpos, next_node, cur, nxt, box, left, right, count, kc);
            CKB(cudaMemcpyAsync(&tail[0], mpos + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            CKB(cudaMemcpyAsync(&tail[1], merged + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            CKB(cudaStreamSynchronize(st));
            const uint32_t merges = tail[0] + tail[1];
            if (merges == 0) { set_error("build_ploc_device: a clustering round merged nothing"); rc = PTC_ERR_CUDA; goto done; }
            round_base.push_back(next_node);
            next_node += merges; m -= (int)merges;
            std::swap(cur, nxt);
        }
        const uint32_t total = next_node;                         // = 2n - 1; the root is node total - 1
        round_base.push_back(total);
        // 3. top-down: depth-first triangle positions, emitted-ancestor counts and depth-first node numbers, round by round from the root's
        CKB(cudaMemsetAsync(offset + (total - 1), 0, 4, st)); CKB(cudaMemsetAsync(depth + (total - 1), 0, 4, st));
        CKB(cudaMemsetAsync(new_id + (total - 1), 0, 4, st)); CKB(cudaMemsetAsync(d_depth, 0, 4, st));
        for (int r = (int)round_base.size() - 2; r >= 0; --r) {
            const uint32_t cnt = round_base[r + 1] - round_base[r];
            k_ploc_down<<<(cnt + T - 1) / T, T, 0, st>>>(round_base[r], round_base[r + 1], left, right, count, kc, offset, depth, new_id);
        }
        k_ploc_depth<<<gb, T, 0, st>>>(ntri, depth, d_depth);
        const uint32_t inner = total - ntri;
        CKB(cudaMemcpyAsync(&tail[0], kc + (total - 1), 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaMemcpyAsync(&dmax, d_depth, 4, cudaMemcpyDeviceToHost, st));
        CKB(cudaStreamSynchronize(st));
        const uint32_t num_nodes = tail[0];
        LbvhFrame fr; lbvh_frame(slo, shi, fr);
        EmitFrame ef; ef.pad = fr.pad;
        for (int a = 0; a < 3; ++a) { ef.base[a] = fr.grid.base[a]; ef.step[a] = fr.step[a]; }
        CKB(cudaMalloc(&nodes, sizeof(DevNode2) * (size_t)num_nodes)); CKB(cudaMalloc(&tris, sizeof(DevTri) * (size_t)n));
        k_ploc_emit_nodes<<<(inner + T - 1) / T, T, 0, st>>>(ntri, total, box, left, right, count, offset, new_id, ef, nodes);
        k_ploc_emit_tris<<<gb, T, 0, st>>>(ntri, ids2, offset, d_pos, tris);
        CKB(cudaEventRecord(e1, st));
        CKB(cudaStreamSynchronize(st));
        CKB(cudaGetLastError());
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1);
        out.nodes = nodes; out.tris = tris; nodes = nullptr; tris = nullptr;
        out.num_nodes = num_nodes; out.num_tris = ntri; out.grid = fr.grid; out.max_depth = dmax + 1; out.build_ms = ms;
    }
done:
    cudaFree(prim); cudaFree(box); cudaFree(bounds); cudaFree(d_depth); cudaFree(keys); cudaFree(keys2); cudaFree(ids); cudaFree(ids2);
    cudaFree(cur); cudaFree(nxt); cudaFree(nn); cudaFree(merged); cudaFree(valid); cudaFree(mpos); cudaFree(vpos);
    cudaFree(left); cudaFree(right); cudaFree(count); cudaFree(offset); cudaFree(depth); cudaFree(kc); cudaFree(new_id);
    cudaFree(tmp); cudaFree(nodes); cudaFree(tris);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

}  // namespace ptc
