// bvh_build.cpp — binned-SAH BVH2 builder (host, multi-threaded) and device-layout emission.
// See bvh_build.h for why the library owns this (reference: src/render_steps/pt.rs:364-468).
#include "bvh_build.h"
#include "lbvh.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <limits>
#include <mutex>
#include <thread>

namespace ptc {
namespace {

constexpr int kBins = 16;               // 32 / 64 bins and traversal costs 0.6 .. 2.0 move nodes visited per ray by < 2 % (tools/bvh_stats.py)
constexpr float kTraversalCost = 1.2f;   // relative to one triangle test
constexpr uint32_t kTaskGrain = 8192;
constexpr size_t kTopNodes = 128;        // inner nodes numbered breadth-first at the top of the tree    // subtrees below this many triangles are built inline by one thread

// Four floats at a time (GCC / Clang vector extension: SSE on x86-64, NEON on aarch64).  vmin / vmax are std::min(acc, x) /
// std::max(acc, x) lane by lane — the same value in every case, signed zeros and NaN included — so the vectorised build
// makes exactly the tree the scalar one made (order and node hashes compared on five scenes, 36 to 10.7 M triangles).
typedef float v4f __attribute__((vector_size(16)));
inline v4f vmin(v4f acc, v4f x) { return x < acc ? x : acc; }
inline v4f vmax(v4f acc, v4f x) { return acc < x ? x : acc; }
inline v4f vset(float x) { return v4f{x, x, x, x}; }

// 48 bytes, three aligned 16-byte rows; the fourth lane of each row is never read as a coordinate (lo's carries the id)
struct alignas(16) Prim { float lo[3]; uint32_t id; float hi[3]; uint32_t pad0; float c[3]; uint32_t pad1; };
inline v4f row(const float* p) { return *reinterpret_cast<const v4f*>(p); }
struct Box4 {
    v4f lo, hi;
    void reset() { lo = vset(std::numeric_limits<float>::infinity()); hi = vset(-std::numeric_limits<float>::infinity()); }
    void grow(v4f l, v4f h) { lo = vmin(lo, l); hi = vmax(hi, h); }
    float half_area() const { const v4f d = hi - lo; return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]; }
};

// bin of a scaled centroid offset; coordinates near FLT_MAX make the scale 0 or the offset infinite, and a NaN or an
// out-of-range value must not reach the float -> int conversion (undefined): those land in bin 0 / the last bin
inline int bin_of(float x) { return x >= 0.0f ? (x < (float)kBins ? (int)x : kBins - 1) : 0; }

// f(begin, end) over [0, n) on the host threads: the per-triangle and per-node loops around the build proper
template <class F>
void par_for(uint64_t n, int nthreads, F f) {
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (n < 65536 || nthreads <= 1) { f((uint64_t)0, n); return; }
    const uint64_t nt = (uint64_t)std::min(nthreads, 32), chunk = (n + nt - 1) / nt;
    std::vector<std::thread> th;
    for (uint64_t t = 0; t < nt; ++t) {
        const uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto& x : th) x.join();
}

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const float* l, const float* h) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); } }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Builder {
    std::vector<Prim, NoInitAlloc<Prim>> prims;
    std::vector<BuildNode, NoInitAlloc<BuildNode>>* nodes;
    std::atomic<uint32_t> next_node{1};
    std::atomic<uint32_t> max_depth{0};
    uint32_t max_leaf;

    struct Task { uint32_t node, first, count, depth; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Task> queue;
    int pending = 0;     // tasks queued or running
    bool done = false;

    void note_depth(uint32_t d) {
        uint32_t cur = max_depth.load(std::memory_order_relaxed);
        while (d > cur && !max_depth.compare_exchange_weak(cur, d)) {}
    }

    // Splits [first, first+count); returns mid or 0 when the range stays a leaf.
    // Two passes over the range: bounds and centroid bounds, then the bins of all three axes at once (a primitive is
    // loaded once and its box goes to one bin per axis with two vector min / max each); the sweeps over the 16 bins per
    // axis evaluate the same costs in the same order as ever.  (Splitting the top ranges of a large scene with all threads
    // together — chunked bounds / bins, stable partition through a scratch array — was built and measured: the root split
    // of 4 M primitives went from 156 to 82 ms on 8 cores, the next three levels did not move, 1.13 s against 1.16 s for the
    // whole build: those passes are bound by memory bandwidth, not by cores.  Not kept.)
    uint32_t split(uint32_t node, uint32_t first, uint32_t count) {
        Box4 bounds4, cb4; bounds4.reset(); cb4.reset();
        for (uint32_t i = first; i < first + count; ++i) { const Prim& p = prims[i]; bounds4.grow(row(p.lo), row(p.hi)); const v4f c = row(p.c); cb4.grow(c, c); }
        Box bounds, cb;
        for (int a = 0; a < 3; ++a) { bounds.lo[a] = bounds4.lo[a]; bounds.hi[a] = bounds4.hi[a]; cb.lo[a] = cb4.lo[a]; cb.hi[a] = cb4.hi[a]; }
        BuildNode& n = (*nodes)[node];
        std::memcpy(n.lo, bounds.lo, 12); std::memcpy(n.hi, bounds.hi, 12);
        n.left = 0xFFFFFFFFu; n.first = first; n.count = count;
        if (count == 1) return 0;

        bool valid[3]; float scale[3];
        for (int a = 0; a < 3; ++a) { const float ext = cb.hi[a] - cb.lo[a]; valid[a] = ext > 0.0f; scale[a] = valid[a] ? (float)kBins / ext : 0.0f; }
        Box4 bb[3][kBins]; uint32_t bn[3][kBins];
        for (int a = 0; a < 3; ++a) if (valid[a]) for (int b = 0; b < kBins; ++b) { bb[a][b].reset(); bn[a][b] = 0; }
        if (valid[0] && valid[1] && valid[2]) {
            for (uint32_t i = first; i < first + count; ++i) {
                const Prim& p = prims[i]; const v4f l = row(p.lo), h = row(p.hi);
                const int b0 = bin_of((p.c[0] - cb.lo[0]) * scale[0]), b1 = bin_of((p.c[1] - cb.lo[1]) * scale[1]), b2 = bin_of((p.c[2] - cb.lo[2]) * scale[2]);
                bb[0][b0].grow(l, h); bn[0][b0]++; bb[1][b1].grow(l, h); bn[1][b1]++; bb[2][b2].grow(l, h); bn[2][b2]++;
            }
        } else {
            for (uint32_t i = first; i < first + count; ++i) {
                const Prim& p = prims[i]; const v4f l = row(p.lo), h = row(p.hi);
                for (int a = 0; a < 3; ++a) if (valid[a]) { const int b = bin_of((p.c[a] - cb.lo[a]) * scale[a]); bb[a][b].grow(l, h); bn[a][b]++; }
            }
        }
        float best = std::numeric_limits<float>::infinity(); int best_axis = -1, best_split = -1;
        for (int a = 0; a < 3; ++a) {
            if (!valid[a]) continue;
            float right_area[kBins]; uint32_t right_n[kBins];
            Box4 acc; acc.reset(); uint32_t cnt = 0;
            for (int b = kBins - 1; b >= 1; --b) { acc.grow(bb[a][b].lo, bb[a][b].hi);          // an empty bin is (+inf, -inf): growing by it changes nothing
                cnt += bn[a][b]; right_n[b] = cnt; right_area[b] = cnt ? acc.half_area() : 0.0f; }
            acc.reset(); cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(bb[a][b].lo, bb[a][b].hi);
                cnt += bn[a][b];
                if (!cnt || !right_n[b + 1]) continue;
                float cost = acc.half_area() * (float)cnt + right_area[b + 1] * (float)right_n[b + 1];
                if (cost < best) { best = cost; best_axis = a; best_split = b; }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {                                    // every centroid identical
            if (count <= max_leaf) return 0;
            mid = first + count / 2;
        } else {
            float parent = bounds.half_area();
            float split_cost = kTraversalCost * parent + best;   // SAH, un-normalised by parent area
            float leaf_cost = parent * (float)count;
            if (count <= max_leaf && leaf_cost <= split_cost) return 0;
            float ext = cb.hi[best_axis] - cb.lo[best_axis], k = (float)kBins / ext, base = cb.lo[best_axis];
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim& p) {
                return bin_of((p.c[best_axis] - base) * k) <= best_split;
            });
            mid = (uint32_t)(it - prims.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        uint32_t left = next_node.fetch_add(2);
        (*nodes)[node].left = left;
        return mid;
    }

    void build_inline(uint32_t node, uint32_t first, uint32_t count, uint32_t depth) {
        note_depth(depth);
        uint32_t mid = split(node, first, count);
        if (!mid) return;
        uint32_t left = (*nodes)[node].left;
        build_inline(left, first, mid - first, depth + 1);
        build_inline(left + 1, mid, first + count - mid, depth + 1);
    }

    void push(Task t) { { std::lock_guard<std::mutex> l(mu); queue.push_back(t); pending++; } cv.notify_one(); }

    void worker() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [&] { return done || !queue.empty(); });
                if (queue.empty()) return;
                t = queue.front(); queue.pop_front();
            }
            if (t.count <= kTaskGrain) build_inline(t.node, t.first, t.count, t.depth);
            else {
                note_depth(t.depth);
                uint32_t mid = split(t.node, t.first, t.count);
                if (mid) {
                    uint32_t left = (*nodes)[t.node].left;
                    push(Task{left, t.first, mid - t.first, t.depth + 1});
                    push(Task{left + 1, mid, t.first + t.count - mid, t.depth + 1});
                }
            }
            {
                std::lock_guard<std::mutex> l(mu);
                if (--pending == 0) { done = true; cv.notify_all(); }
            }
        }
    }
};

}  // namespace

void build_bvh2(const float* tri_pos, uint32_t ntri, uint32_t max_leaf, int nthreads, Bvh2& out) {
    auto t0 = std::chrono::steady_clock::now();
    out.nodes.clear(); out.order.clear(); out.max_depth = 0; out.pad = 0.0f;
    if (ntri == 0) { out.build_ms = 0.0; return; }
    if (max_leaf < 1) max_leaf = 1;
    if (max_leaf > 8) max_leaf = 8;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;

    Builder b;
    b.max_leaf = max_leaf;
    b.prims.resize(ntri);
    Box scene; scene.reset();
    {
        std::mutex mu;
        par_for(ntri, nthreads, [&](uint64_t i0, uint64_t i1) {
            Box local; local.reset();
            for (uint64_t i = i0; i < i1; ++i) {
                Prim& p = b.prims[i]; p.id = (uint32_t)i; p.pad0 = p.pad1 = 0u;
                const float* v = tri_pos + (size_t)i * 9;
                for (int a = 0; a < 3; ++a) {
                    p.lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a]));
                    p.hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a]));
                    p.c[a] = 0.5f * (p.lo[a] + p.hi[a]);
                }
                local.grow(p.lo, p.hi);
            }
            std::lock_guard<std::mutex> l(mu); scene.grow(local.lo, local.hi);      // min / max: the order of the merges does not matter
        });
    }
    // Padding that makes every box conservative with respect to pt_ray_tri: Moeller-Trumbore can accept a
    // ray passing O(eps * distance) outside a triangle edge, so boxes grow by 2e-5 of the scene magnitude.
    float mag = 0.0f;
    for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])), scene.hi[a] - scene.lo[a]));
    out.pad = mag * 2e-5f + 1e-30f;

    out.nodes.resize((size_t)2 * ntri);
    b.nodes = &out.nodes;
    b.push(Builder::Task{0, 0, ntri, 1});
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&b] { b.worker(); });
    for (auto& t : th) t.join();

    out.nodes.resize(b.next_node.load());
    out.order.resize(ntri);
    par_for(ntri, nthreads, [&](uint64_t i0, uint64_t i1) { for (uint64_t i = i0; i < i1; ++i) out.order[i] = b.prims[i].id; });
    out.max_depth = b.max_depth.load();
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void emit_bvh2(const Bvh2& bvh, const float* tri_pos, DevBvh2& out) {
    out.nodes.clear(); out.tris.clear(); out.grid = QuantGrid{};
    if (bvh.nodes.empty()) return;
    uint32_t ntri = (uint32_t)bvh.order.size();
    out.tris.resize(ntri);
    par_for(ntri, 0, [&](uint64_t i0, uint64_t i1) {
        for (uint64_t i = i0; i < i1; ++i) {
            uint32_t id = bvh.order[i];
            const float* v = tri_pos + (size_t)id * 9;
            DevTri& t = out.tris[i]; std::memset(&t, 0, sizeof(t));
            t.q[0] = v[0]; t.q[1] = v[1]; t.q[2] = v[2]; std::memcpy(&t.q[3], &id, 4);
            t.q[4] = v[3] - v[0]; t.q[5] = v[4] - v[1]; t.q[6] = v[5] - v[2];
            t.q[8] = v[6] - v[0]; t.q[9] = v[7] - v[1]; t.q[10] = v[8] - v[2];
        }
    });
    // quantisation grid: the padded scene box, widened by a few steps so that outward rounding never clamps
    const BuildNode& root = bvh.nodes[0];
    double step[3];
    for (int a = 0; a < 3; ++a) {
        double ext = (double)root.hi[a] - (double)root.lo[a] + 2.0 * (double)bvh.pad;
        double S = ext * (1.0 + 16.0 / 32768.0);
        if (!(S > 1e-20)) S = 1e-20;
        float Sf = std::nextafter((float)S, std::numeric_limits<float>::infinity());
        float bf = std::nextafter((float)((double)root.lo[a] - (double)bvh.pad - 0.5 * ((double)Sf - ext)), -std::numeric_limits<float>::infinity());
        out.grid.base[a] = bf; out.grid.extent[a] = Sf;
        step[a] = (double)Sf / 32768.0;
    }
    // inner build nodes -> device node ids in depth-first order
    std::vector<int32_t> dev_id(bvh.nodes.size(), -1);
    std::vector<uint32_t> stack;
    uint32_t n_inner = 0;
    auto is_leaf = [&](uint32_t n) { return bvh.nodes[n].left == 0xFFFFFFFFu; };
    if (is_leaf(0)) {
        // single-leaf tree: synthesise a root whose second child is absent
        out.nodes.resize(1);
    } else {
        // numbering: the first kTopNodes inner nodes in breadth-first order (the top of the tree is contiguous: every ray
        // starts there; used by the shared-memory staging experiment PTC_SMEM_TOP), the rest depth-first below them
        std::vector<uint32_t> bfs; bfs.push_back(0);
        for (size_t h = 0; h < bfs.size() && bfs.size() < kTopNodes; ++h) {
            uint32_t l = bvh.nodes[bfs[h]].left;
            if (!is_leaf(l) && bfs.size() < kTopNodes) bfs.push_back(l);
            if (!is_leaf(l + 1) && bfs.size() < kTopNodes) bfs.push_back(l + 1);
        }
        for (uint32_t n : bfs) dev_id[n] = (int32_t)n_inner++;
        for (size_t h = bfs.size(); h-- > 0;) stack.push_back(bfs[h]);
        while (!stack.empty()) {
            uint32_t n = stack.back(); stack.pop_back();
            if (dev_id[n] < 0) dev_id[n] = (int32_t)n_inner++;
            uint32_t l = bvh.nodes[n].left;
            if (!is_leaf(l + 1) && dev_id[l + 1] < 0) stack.push_back(l + 1);
            if (!is_leaf(l) && dev_id[l] < 0) stack.push_back(l);
        }
        out.nodes.resize(n_inner);
    }
    auto child_ref = [&](uint32_t n) -> uint32_t {
        if (!is_leaf(n)) return (uint32_t)dev_id[n];
        const BuildNode& b = bvh.nodes[n];
        uint32_t enc = (b.first << 3) | (b.count - 1);
        return ~enc;
    };
    // planes rounded outwards by one extra step: the kernel's plane distance carries up to ~1/30 step of rounding error
    auto put_box = [&](DevNode2& d, int c, const BuildNode& b) {
        for (int a = 0; a < 3; ++a) {
            double lo = ((double)b.lo[a] - (double)bvh.pad - (double)out.grid.base[a]) / step[a];
            double hi = ((double)b.hi[a] + (double)bvh.pad - (double)out.grid.base[a]) / step[a];
            if (!(lo > -8.0)) lo = -8.0;                           // out of the grid or NaN: the whole axis (conservative), and the
            if (!(hi < 40000.0)) hi = 40000.0;                     // conversion below stays defined
            long ql = (long)std::floor(lo) - 1, qh = (long)std::ceil(hi) + 1;
            ql = std::min(std::max(ql, 0l), 32767l); qh = std::min(std::max(qh, 0l), 32767l);
            d.w[c * 4 + a] = (0x8000u | (uint32_t)ql) | ((0x8000u | (uint32_t)qh) << 16);
        }
    };
    if (is_leaf(0)) {
        DevNode2& d = out.nodes[0];
        put_box(d, 0, bvh.nodes[0]);
        for (int a = 0; a < 3; ++a) d.w[4 + a] = 0xFFFFu | (0x8000u << 16);   // absent child: inverted box (lo = 32767, hi = 0), never entered
        d.w[3] = d.w[7] = child_ref(0);
        return;
    }
    par_for(bvh.nodes.size(), 0, [&](uint64_t n0, uint64_t n1) {
        for (uint64_t n = n0; n < n1; ++n) {
            if (dev_id[n] < 0) continue;
            DevNode2& d = out.nodes[dev_id[n]];
            uint32_t l = bvh.nodes[n].left;
            put_box(d, 0, bvh.nodes[l]); put_box(d, 1, bvh.nodes[l + 1]);
            d.w[3] = child_ref(l); d.w[7] = child_ref(l + 1);
        }
    });
}

// ---- linear BVH: host restatement of bvh_gpu.cu --------------------------------------------------------------------
void lbvh_frame(const float lo[3], const float hi[3], LbvhFrame& f) {
    float mag = 0.0f;
    for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::max(std::fabs(lo[a]), std::fabs(hi[a])), hi[a] - lo[a]));
    f.pad = mag * 2e-5f + 1e-30f;                                  // as build_bvh2
    for (int a = 0; a < 3; ++a) {                                  // as emit_bvh2
        double ext = (double)hi[a] - (double)lo[a] + 2.0 * (double)f.pad;
        double S = ext * (1.0 + 16.0 / 32768.0);
        if (!(S > 1e-20)) S = 1e-20;
        float Sf = std::nextafter((float)S, std::numeric_limits<float>::infinity());
        float bf = std::nextafter((float)((double)lo[a] - (double)f.pad - 0.5 * ((double)Sf - ext)), -std::numeric_limits<float>::infinity());
        f.grid.base[a] = bf; f.grid.extent[a] = Sf;
        f.step[a] = (double)Sf / 32768.0;
    }
}

void build_lbvh_host(const float* tri_pos, uint32_t ntri, DevBvh2& out, uint32_t& max_depth) {
    const int n = (int)ntri;
    struct B6 { float lo[3], hi[3]; };
    std::vector<B6> prim(ntri), node_box(ntri);
    float slo[3], shi[3], clo[3], chi[3];
    for (int a = 0; a < 3; ++a) { slo[a] = clo[a] = std::numeric_limits<float>::infinity(); shi[a] = chi[a] = -slo[a]; }
    for (uint32_t i = 0; i < ntri; ++i) {
        lbvh_tri_box(tri_pos + (size_t)i * 9, prim[i].lo, prim[i].hi);
        for (int a = 0; a < 3; ++a) {
            float c = 0.5f * (prim[i].lo[a] + prim[i].hi[a]);
            slo[a] = fminf(slo[a], prim[i].lo[a]); shi[a] = fmaxf(shi[a], prim[i].hi[a]);
            clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
        }
    }
    float scale[3];
    for (int a = 0; a < 3; ++a) { float ext = chi[a] - clo[a]; scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f; }
    std::vector<uint64_t> key(ntri), keys(ntri);
    std::vector<uint32_t> ids(ntri);
    for (uint32_t i = 0; i < ntri; ++i) { key[i] = lbvh_key(tri_pos + (size_t)i * 9, clo, scale); ids[i] = i; }
    std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (uint32_t i = 0; i < ntri; ++i) keys[i] = key[ids[i]];
    std::vector<int> first(n), last(n), split(n), node_parent(n, -1), leaf_parent(n, -1);
    std::vector<uint32_t> kept(n, 0), new_id(n, 0);
    uint32_t num_nodes = 0;
    for (int i = 0; i < n - 1; ++i) {
        lbvh_node(keys.data(), n, i, first[i], last[i], split[i]);
        const int f = first[i], l = last[i], s = split[i];
        if (f == s) leaf_parent[s] = i; else node_parent[s] = i;
        if (s + 1 == l) leaf_parent[s + 1] = i; else node_parent[s + 1] = i;
        kept[i] = (uint32_t)(l - f + 1) > PTC_LBVH_MAX_LEAF ? 1u : 0u;
    }
    node_parent[0] = -1;
    for (int i = 0; i < n - 1; ++i) { new_id[i] = num_nodes; num_nodes += kept[i]; }
    // boxes bottom-up: a node's range strictly contains its children's, so increasing range size is a valid order
    std::vector<int> order(n - 1);
    for (int i = 0; i < n - 1; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return last[a] - first[a] < last[b] - first[b]; });
    for (int i : order) {
        const int f = first[i], l = last[i], s = split[i];
        const B6& a = (f == s) ? prim[ids[s]] : node_box[s];
        const B6& b = (s + 1 == l) ? prim[ids[s + 1]] : node_box[s + 1];
        for (int x = 0; x < 3; ++x) { node_box[i].lo[x] = fminf(a.lo[x], b.lo[x]); node_box[i].hi[x] = fmaxf(a.hi[x], b.hi[x]); }
    }
    uint32_t depth = 0;
    for (int k = 0; k < n; ++k) {
        uint32_t d = 0;
        for (int a = leaf_parent[k]; a >= 0; a = node_parent[a]) d += kept[a];
        depth = std::max(depth, d);
    }
    max_depth = depth + 1;
    LbvhFrame fr; lbvh_frame(slo, shi, fr);
    out.grid = fr.grid;
    out.nodes.assign(num_nodes, DevNode2{});
    for (int i = 0; i < n - 1; ++i) {
        if (!kept[i]) continue;
        const int f = first[i], l = last[i], s = split[i];
        DevNode2 d;
        const int cf[2] = {f, s + 1}, cl[2] = {s, l}, ci[2] = {s, s + 1};
        for (int c = 0; c < 2; ++c) {
            const uint32_t cnt = (uint32_t)(cl[c] - cf[c] + 1);
            const B6& b = (cnt == 1) ? prim[ids[cf[c]]] : node_box[ci[c]];
            lbvh_put_box(&d.w[c * 4], b.lo, b.hi, fr.pad, fr.grid.base, fr.step);
            d.w[c * 4 + 3] = cnt <= PTC_LBVH_MAX_LEAF ? ~(((uint32_t)cf[c] << 3) | (cnt - 1u)) : new_id[ci[c]];
        }
        out.nodes[new_id[i]] = d;
    }
    out.tris.resize(ntri);
    for (uint32_t i = 0; i < ntri; ++i) lbvh_put_tri(out.tris[i].q, tri_pos + (size_t)ids[i] * 9, ids[i]);
}

// ---- PLOC: host restatement of the device build (bvh_gpu.cu build_ploc_device; shared arithmetic in lbvh.cuh) -------------
// Same keys and sort as the linear BVH, then rounds of nearest-partner search / mutual merge / compaction, a top-down pass
// that numbers the triangles in depth-first order (so that a collapsed leaf's triangles are consecutive), and the emission
// of the quantised nodes.  Every step is a deterministic function of its inputs, so the arrays equal the device build's
// word for word (ptc_debug_bvh_hash).
void build_ploc_host(const float* tri_pos, uint32_t ntri, DevBvh2& out, uint32_t& max_depth, int radius) {
    const int n = (int)ntri;
    std::vector<PlocBox> box((size_t)2 * n - 1);
    std::vector<PlocBox> prim(ntri);
    float slo[3], shi[3], clo[3], chi[3];
    for (int a = 0; a < 3; ++a) { slo[a] = clo[a] = std::numeric_limits<float>::infinity(); shi[a] = chi[a] = -slo[a]; }
    for (uint32_t i = 0; i < ntri; ++i) {
        lbvh_tri_box(tri_pos + (size_t)i * 9, prim[i].lo, prim[i].hi);
        for (int a = 0; a < 3; ++a) {
            float c = 0.5f * (prim[i].lo[a] + prim[i].hi[a]);
            slo[a] = fminf(slo[a], prim[i].lo[a]); shi[a] = fmaxf(shi[a], prim[i].hi[a]);
            clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
        }
    }
    float scale[3];
    for (int a = 0; a < 3; ++a) { float ext = chi[a] - clo[a]; scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f; }
    std::vector<uint64_t> key(ntri);
    std::vector<uint32_t> ids(ntri);
    for (uint32_t i = 0; i < ntri; ++i) { key[i] = lbvh_key(tri_pos + (size_t)i * 9, clo, scale); ids[i] = i; }
    std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    // nodes 0 .. n-1: the sorted primitives; n .. 2n-2: merged clusters in order of creation (the last one is the root)
    std::vector<uint32_t> left((size_t)2 * n - 1, 0xFFFFFFFFu), right((size_t)2 * n - 1, 0xFFFFFFFFu), count((size_t)2 * n - 1, 1u);
    std::vector<uint32_t> kc((size_t)2 * n - 1, 0u);       // emitted (count > PTC_PLOC_MAX_LEAF) nodes in the subtree, the node included
    for (int k = 0; k < n; ++k) box[k] = prim[ids[k]];
    std::vector<uint32_t> cur(n), nxt(n);
    std::vector<int> nn(n);
    for (int k = 0; k < n; ++k) cur[k] = (uint32_t)k;
    std::vector<uint32_t> round_base;                 // first node created in each round (top-down passes walk the rounds backwards)
    uint32_t next_node = (uint32_t)n;
    int m = n;
    const int nthreads = std::max(1u, std::thread::hardware_concurrency());
    auto par = [&](int cnt, auto&& fn) {
        if (cnt < 4096 || nthreads == 1) { fn(0, cnt); return; }
        std::vector<std::thread> th; const int chunk = (cnt + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) { const int b = t * chunk, e = std::min(cnt, b + chunk); if (b < e) th.emplace_back([&, b, e] { fn(b, e); }); }
        for (auto& x : th) x.join();
    };
    while (m > 1) {
        par(m, [&](int b, int e) { for (int i = b; i < e; ++i) nn[i] = ploc_nearest(box.data(), cur.data(), m, i, radius); });
        round_base.push_back(next_node);
        int w = 0;
        for (int i = 0; i < m; ++i) {
            const int j = nn[i];
            const bool mutual = nn[j] == i;
            if (mutual && i > j) continue;                                   // the right partner disappears into the merged cluster
            if (mutual) {
                const uint32_t a = cur[i], b = cur[j], nd = next_node++;
                left[nd] = a; right[nd] = b; count[nd] = count[a] + count[b];
                kc[nd] = kc[a] + kc[b] + (count[nd] > PTC_PLOC_MAX_LEAF ? 1u : 0u);
                for (int x = 0; x < 3; ++x) { box[nd].lo[x] = fminf(box[a].lo[x], box[b].lo[x]); box[nd].hi[x] = fmaxf(box[a].hi[x], box[b].hi[x]); }
                nxt[w++] = nd;
            } else nxt[w++] = cur[i];
        }
        m = w;
        cur.swap(nxt);
    }
    const uint32_t root = cur[0], total = next_node;
    round_base.push_back(total);
    // top-down: position of every subtree's first triangle in depth-first order, emitted ancestors above it, and the node's
    // number in DEPTH-FIRST (pre-)order — a child follows its parent directly, so a descent walks forward through memory
    // (numbering the nodes in order of creation, i.e. level by level, cost 15-25 % of the traversal speed on the 890 MB tree)
    std::vector<uint32_t> offset(total, 0), depth(total, 0), new_id(total, 0);
    auto kept = [&](uint32_t nd) { return nd >= (uint32_t)n && count[nd] > PTC_PLOC_MAX_LEAF; };
    for (int r = (int)round_base.size() - 2; r >= 0; --r)
        for (uint32_t nd = round_base[r + 1]; nd-- > round_base[r];) {
            const uint32_t a = left[nd], b = right[nd], k = kept(nd) ? 1u : 0u;
            offset[a] = offset[nd]; offset[b] = offset[nd] + count[a];
            depth[a] = depth[b] = depth[nd] + k;
            new_id[a] = new_id[nd] + 1u; new_id[b] = new_id[nd] + 1u + kc[a];
        }
    uint32_t dmax = 0;
    for (int k = 0; k < n; ++k) dmax = std::max(dmax, depth[k]);
    max_depth = dmax + 1;
    const uint32_t num_nodes = kc[root];
    LbvhFrame fr; lbvh_frame(slo, shi, fr);
    out.grid = fr.grid;
    out.nodes.assign(num_nodes, DevNode2{});
    for (uint32_t nd = (uint32_t)n; nd < total; ++nd) {
        if (!kept(nd)) continue;
        DevNode2 d;
        const uint32_t ch[2] = {left[nd], right[nd]};
        for (int c = 0; c < 2; ++c) {
            lbvh_put_box(&d.w[c * 4], box[ch[c]].lo, box[ch[c]].hi, fr.pad, fr.grid.base, fr.step);
            d.w[c * 4 + 3] = count[ch[c]] <= PTC_PLOC_MAX_LEAF ? ~((offset[ch[c]] << 3) | (count[ch[c]] - 1u)) : new_id[ch[c]];
        }
        out.nodes[new_id[nd]] = d;
    }
    out.tris.resize(ntri);
    for (int k = 0; k < n; ++k) lbvh_put_tri(out.tris[offset[k]].q, tri_pos + (size_t)ids[k] * 9, ids[k]);
}

}  // namespace ptc
