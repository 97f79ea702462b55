// bvh_build.h — host-side BVH construction for the CUDA traversal kernels.
//
// The reference never builds a tree: it hands per-mesh triangle lists to the
// Vulkan driver (one BLAS per mesh with one geometry per triangle, one TLAS of
// identity instances; reference src/render_steps/pt.rs:364-468,524-545) and the
// RT hardware does the rest.  B200 has no RT cores, so this library owns the
// acceleration structure: one tree over every triangle of every mesh (the
// instances are all identity transforms, so flattening loses nothing).
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <vector>

namespace ptc {

struct BuildNode {
    float lo[3], hi[3];      // unpadded bounds
    uint32_t left;           // inner: index of the first child (second is left+1); leaf: 0xFFFFFFFF
    uint32_t first, count;   // leaf: range in `order`
};

// std::allocator that leaves trivially constructible elements uninitialised: resize() of a 10 M-triangle node or
// primitive array is then an allocation, not a single-threaded pass over hundreds of megabytes that the build threads
// overwrite a moment later (the pages are first touched by the threads that fill them).
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(static_cast<A&&>(a)...);
    }
};

struct Bvh2 {
    std::vector<BuildNode, NoInitAlloc<BuildNode>> nodes;    // nodes[0] is the root
    std::vector<uint32_t> order;     // leaf-ordered global triangle ids
    uint32_t max_depth = 0;
    float pad = 0.0f;                // conservative padding to apply to every box when emitting device nodes
    double build_ms = 0.0;
};

// positions: per-triangle 9 floats (v0, v1, v2), ntri triangles.
// Binned SAH (16 bins x 3 axes), leaves of at most `max_leaf` triangles,
// multi-threaded over subtrees.
void build_bvh2(const float* tri_pos, uint32_t ntri, uint32_t max_leaf, int nthreads, Bvh2& out);

// ---- device layouts --------------------------------------------------------

// BVH2 device node, 32 bytes = ONE 256-bit load (Aila-Laine style: the node carries BOTH children's boxes so one
// visit decides both descents).  Box planes are quantised to a 15-bit grid over the (padded) scene box, rounded
// outwards — the traversal kernel is bound by L1 wavefronts (one per lane per load instruction for divergent lanes),
// so halving the node from two loads to one matters more than the ~1e-4 relative growth of the boxes:
//   w[0..2] = child 0, axis x/y/z: lo16 | hi16 << 16      w[3] = child 0 reference
//   w[4..6] = child 1, axis x/y/z                         w[7] = child 1 reference
// each 16-bit plane is 0x8000 | q (q in [0, 32767]): with 0x3F as the top byte and 0x00 as the bottom byte (one PRMT)
// it IS the float 1 + q / 32768, so decoding costs one instruction and the plane distance one FMA (pt_device.cuh).
// plane position = grid.base[axis] + q * grid.extent[axis] / 32768.
// child >= 0: inner node index.  child < 0: leaf, ~child = (first_tri << 3) | (count - 1), first_tri indexing the
// leaf-ordered triangle array.  An absent child (single-leaf tree)
// has an inverted box (lo = 32767, hi = 0) and is never entered.
struct DevNode2 { uint32_t w[8]; };

// Leaf-ordered triangle for traversal, 64-byte stride so that (v0, id, e1) is one aligned 256-bit load:
//   (v0.xyz, global triangle id bits) (e1.xyz, 0) | (e2.xyz, 0) (unused)
// e1 = v1 - v0, e2 = v2 - v0 in fp32.
struct DevTri { float q[16]; };

// quantisation grid of the node boxes (per axis)
struct QuantGrid { float base[3]; float extent[3]; };

struct DevBvh2 {
    std::vector<DevNode2> nodes;
    std::vector<DevTri> tris;
    QuantGrid grid{};
};

void emit_bvh2(const Bvh2& bvh, const float* tri_pos, DevBvh2& out);

// Host restatement of the device LBVH build (bvh_gpu.cu; shared arithmetic in lbvh.cuh): same keys, same radix tree,
// same collapse, same node numbering, so the output equals the device build word for word.  Used by the CPU-only test
// harness; the product builds on the GPU.  Needs ntri > 4; max_depth = levels of the collapsed tree.
void build_lbvh_host(const float* tri_pos, uint32_t ntri, DevBvh2& out, uint32_t& max_depth);
// the same for the PLOC build (parallel locally-ordered clustering, lbvh.cuh): ntri > PTC_PLOC_MAX_LEAF
void build_ploc_host(const float* tri_pos, uint32_t ntri, DevBvh2& out, uint32_t& max_depth, int radius);

}  // namespace ptc
