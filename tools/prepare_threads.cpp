// prepare_threads.cpp — prepare_scene (csrc/scene_prep.cpp) on a scene large enough for every threaded host path: the parallel
// per-triangle loops, the SAH task pool, emit, and the LBVH / PLOC host builds; one triangle is inactive (NaN corner).  Built
// with ThreadSanitizer by `tools/fuzz.sh tsan`; prints the tree sizes per builder.  Test infrastructure, not part of libptcore.
#include "../wgpu-pathtracer_b200/csrc/pt_device.cuh"
#include "../wgpu-pathtracer_b200/csrc/scene_prep.h"
#include "../wgpu-pathtracer_b200/csrc/pt_error.h"
#include <cstdio>
#include <cstring>
#include <random>
namespace ptc { static thread_local std::string g_err; void set_error(const std::string& s) { g_err = s; } }
using namespace ptc;
int main() {
    const uint32_t ntri = 120000;
    std::mt19937 rng(3); std::uniform_real_distribution<float> U(-1, 1);
    std::vector<ptc_vertex> V(ntri * 3); std::vector<uint32_t> I(ntri * 3);
    for (uint32_t t = 0; t < ntri; ++t) { float c[3] = {U(rng) * 10, U(rng) * 10, U(rng) * 10};
        for (int k = 0; k < 3; ++k) { ptc_vertex& v = V[t * 3 + k]; std::memset(&v, 0, sizeof(v)); for (int a = 0; a < 3; ++a) v.position[a] = c[a] + U(rng) * 0.1f; v.normal[1] = 1; I[t * 3 + k] = t * 3 + k; } }
    V[5].position[0] = NAN;   // one inactive triangle: exercises the parallel canonicalisation
    ptc_mesh m{0, ntri * 3, 0, ntri * 3, 0};
    ptc_material mat; std::memset(&mat, 0, sizeof(mat)); mat.albedo_texture_idx = 0; mat.roughness_texture_idx = 1; mat.specular_texture_idx = 2; mat.metallic_texture_idx = 3; mat.emissive_texture_idx = 4; mat.sheen_texture_idx = 5; mat.normal_texture_idx = 6; mat.transmission_texture_idx = 7; mat.ior_texture_idx = 8;
    float tex[10][4]; ptc_texture_desc T[10]; for (int k = 0; k < 10; ++k) { for (int j = 0; j < 4; ++j) tex[k][j] = 0.5f; T[k] = ptc_texture_desc{1, 1, PTC_TEXEL_RGBA32F, 0, tex[k]}; }
    ptc_scene_view v; std::memset(&v, 0, sizeof(v)); v.vertices = V.data(); v.num_vertices = V.size(); v.indices = I.data(); v.num_indices = I.size(); v.meshes = &m; v.num_meshes = 1; v.materials = &mat; v.num_materials = 1; v.textures = T; v.num_textures = 10;
    for (uint32_t b = 0; b < 3; ++b) { ptc_config cfg; std::memset(&cfg, 0, sizeof(cfg)); cfg.width = cfg.height = 8; cfg.max_bounces = 4; cfg.samples_per_pixel = 1; cfg.env_texture_index = 9; cfg.bvh_build = b;
        PreparedScene ps; int rc = prepare_scene(&v, cfg, ps); printf("builder %u rc %d nodes %zu depth %u\n", b, rc, ps.bvh.nodes.size(), ps.max_depth); }
}
