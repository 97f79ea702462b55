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

}  // namespace ptc
