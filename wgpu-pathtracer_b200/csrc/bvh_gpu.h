// bvh_gpu.h — device-side linear BVH build (see bvh_gpu.cu).
#pragma once
#include "bvh_build.h"
#include <cuda_runtime.h>

namespace ptc {

struct GpuBvh {
    DevNode2* nodes = nullptr;      // device memory, owned by the caller after a successful build (cudaFree)
    DevTri* tris = nullptr;
    uint32_t num_nodes = 0, num_tris = 0, max_depth = 0;
    QuantGrid grid{};
    float build_ms = 0.0f;          // CUDA-event time of the whole build on the stream
};

// d_pos: ntri x 9 floats (v0, v1, v2) in device memory; ntri > 4.  Returns PTC_OK or a negative ptc_status.
int build_lbvh_device(const float* d_pos, uint32_t ntri, cudaStream_t st, GpuBvh& out);
// the same interface for the PLOC build (parallel locally-ordered clustering, lbvh.cuh); ntri > 4
int build_ploc_device(const float* d_pos, uint32_t ntri, cudaStream_t st, GpuBvh& out, int radius);

}  // namespace ptc
