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
#include <vector>

namespace ptc {

struct BuildNode {
    float lo[3], hi[3];      // unpadded bounds
    uint32_t left;           // inner: index of the first child (second is left+1); leaf: 0xFFFFFFFF
    uint32_t first, count;   // leaf: range in `order`
};

struct Bvh2 {
    std::vector<BuildNode> nodes;    // nodes[0] is the root
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

// BVH2 device node, 64 bytes, four 128-bit loads (Aila-Laine style: the node
// carries BOTH children's boxes so one visit decides both descents):
//   q0 = c0.lo.x c0.hi.x c0.lo.y c0.hi.y
//   q1 = c1.lo.x c1.hi.x c1.lo.y c1.hi.y
//   q2 = c0.lo.z c0.hi.z c1.lo.z c1.hi.z
//   q3 = child0 child1 (int bits) 0 0
// child >= 0: inner node index.  child < 0: leaf, ~child = (first_tri << 3) | (count - 1),
// first_tri indexing the leaf-ordered triangle array.  An absent child has an
// inverted box (lo = +inf, hi = -inf) and is never entered.
struct DevNode2 { float q[16]; };

// Leaf-ordered triangle for traversal, 48 bytes, three 128-bit loads:
//   (v0.xyz, global triangle id bits) (e1.xyz, 0) (e2.xyz, 0), e1 = v1 - v0, e2 = v2 - v0 in fp32
struct DevTri { float q[12]; };

struct DevBvh2 {
    std::vector<DevNode2> nodes;
    std::vector<DevTri> tris;
};

void emit_bvh2(const Bvh2& bvh, const float* tri_pos, DevBvh2& out);

}  // namespace ptc
