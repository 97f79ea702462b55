// scene_prep.h — host-side preparation of the device scene: validation, BVH build, per-triangle
// shading records, constant-folded materials, emissive-triangle list.  Produces host vectors that
// pt_step.cu uploads verbatim (and that the CPU-only test harness in tests/hostemu consumes directly).
// Restates the host half of RTStep::create / create_static_bind_groups (reference:
// src/render_steps/pt.rs:188-243,341-433).
#pragma once
#include "../../include/ptcore_types.h"
#include "bvh_build.h"
#include "pt_device.cuh"

#include <string>
#include <vector>

namespace ptc {

struct PreparedScene {
    DevBvh2 bvh;                          // device-layout nodes + leaf-ordered triangles
    std::vector<TriShade, NoInitAlloc<TriShade>> shade;   // global triangle order; every word is written by prepare_scene (no zero-fill first)
    std::vector<MatDev> mats;
    std::vector<ptc_triangle> lights;     // never empty (dummy entry, pt.rs:232-239); _pad0 = the light's global triangle id
    std::vector<uint32_t> light_of_tri;   // global triangle -> index into lights (0xFFFFFFFF: not emissive)
    uint32_t num_real_lights = 0;         // 0 when lights holds only the dummy entry
    std::vector<uint32_t> tri_mesh;       // global triangle -> mesh (instance_index)
    std::vector<ptc_mesh> meshes;
    std::vector<float> tri_pos;           // 9 floats per triangle; filled only when the caller builds the BVH itself
    uint32_t num_tris = 0;
    uint32_t max_depth = 0;
    double build_ms = 0.0;
};

// returns PTC_OK or a negative ptc_status with set_error() called
// caller_builds_bvh: leave out.bvh empty and hand the triangle positions over in out.tri_pos (the product's device build);
// otherwise cfg.bvh_build selects the host builder (binned SAH, or the host restatement of the device LBVH build).
int prepare_scene(const ptc_scene_view* v, const ptc_config& cfg, PreparedScene& out, bool caller_builds_bvh = false);

}  // namespace ptc
