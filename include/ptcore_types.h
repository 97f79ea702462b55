/* ptcore_types.h — POD layouts that cross the C ABI.
 *
 * Every struct here is byte-for-byte the layout the reference uploads to its
 * WGSL compute pass (reference: src/uniforms.rs:3-215, mirrored by
 * src/shaders/common/uniforms.wgsl:3-79) or keeps in `scene::Scene`
 * (reference: src/scene.rs:54-74).  A Rust caller can hand the library a
 * `bytemuck::cast_slice(&scene.vertices)` pointer unchanged.
 *
 * Plain C, no CUDA or torch types.
 */
#ifndef PTCORE_TYPES_H
#define PTCORE_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference: src/uniforms.rs:3-27 (Camera, 96 bytes) */
typedef struct ptc_camera {
    float    position[3];        /* @0  */
    uint32_t _pad0;              /* @12 */
    float    lookat[3];          /* @16 */
    uint32_t frame;              /* @28  RNG stream id; first rendered frame is 1 (raytracer.rs:180) */
    float    pixel_space_x[3];   /* @32 */
    uint32_t _pad1;
    float    pixel_space_y[3];   /* @48 */
    uint32_t _pad2;
    float    first_pixel_pos[3]; /* @64 */
    uint32_t width;              /* @76 */
    uint32_t height;             /* @80 */
    float    fov;                /* @84 */
    uint32_t samples_per_pixel;  /* @88  uploaded but ignored by the shader (pathtracer/pathtracer.wgsl:258) */
    uint32_t max_bounces;        /* @92  ditto (:159) */
} ptc_camera;

/* reference: src/uniforms.rs:53-61 (Uniforms, 112 bytes) */
typedef struct ptc_uniforms {
    ptc_camera camera;
    uint32_t   num_lights;       /* @96  */
    uint32_t   is_grabbed;       /* @100 non-zero resets the accumulation (pathtracer.wgsl:36-40) */
    uint32_t   _pad[2];
} ptc_uniforms;

/* reference: src/uniforms.rs:92-113 (Material, 52 bytes, 13 words) */
typedef struct ptc_material {
    uint32_t albedo_texture_idx;
    uint32_t roughness_texture_idx;
    uint32_t metallic_texture_idx;
    uint32_t emissive_texture_idx;
    uint32_t specular_texture_idx;
    uint32_t sheen_texture_idx;
    float    clearcoat_thickness;
    float    clearcoat_roughness;
    float    anisotropy;
    float    anisotropy_rotation;
    uint32_t normal_texture_idx;
    uint32_t transmission_texture_idx;
    uint32_t ior_texture_idx;
} ptc_material;

/* reference: src/uniforms.rs:138-147 (Vertex, 48 bytes) */
typedef struct ptc_vertex {
    float    position[3];
    uint32_t _pad;
    float    normal[3];
    uint32_t _pad2;
    float    uv[2];
    uint32_t _pad3[2];
} ptc_vertex;

/* reference: src/uniforms.rs:204-215 (Triangle, 64 bytes; the NEE light list, pt.rs:206-243) */
typedef struct ptc_triangle {
    float v0[3]; uint32_t _pad0;
    float v1[3]; uint32_t _pad1;
    float v2[3]; uint32_t _pad2;
    float normal[3]; uint32_t _pad3;
} ptc_triangle;

/* reference: src/scene.rs:54-63 (scene::Mesh).  One mesh == one BLAS == one
 * TLAS instance (pt.rs:364-433,527-542): instance_index = position in this
 * array, instance_custom_data = material_index, geometry_index = triangle
 * number inside the mesh. */
typedef struct ptc_mesh {
    uint32_t vertex_offset;
    uint32_t vertex_count;
    uint32_t index_offset;   /* into the global index array; also InstanceInfo.index_offset (pt.rs:199-203) */
    uint32_t index_count;
    uint32_t material_index;
} ptc_mesh;

/* Decoded texels handed over by the caller (the reference decodes with the
 * `image` crate and uploads rgba32f or BC7, src/texture.rs:232-322; this
 * library takes the pre-BC7 texels). */
enum ptc_texel_format {
    PTC_TEXEL_RGBA8   = 0,   /* 4 x u8, value/255, NO sRGB decode (texture.rs:241 to_rgba32f) */
    PTC_TEXEL_RGBA32F = 1    /* 4 x f32 (EXR, constant colours) */
};

typedef struct ptc_texture_desc {
    uint32_t    width;
    uint32_t    height;
    uint32_t    format;      /* enum ptc_texel_format */
    uint32_t    _pad;
    const void* texels;      /* row-major, tightly packed, row 0 first */
} ptc_texture_desc;

/* Borrowed view of the arrays `scene::Scene` owns (src/scene.rs:65-74). */
typedef struct ptc_scene_view {
    const ptc_vertex*       vertices;   uint64_t num_vertices;
    const uint32_t*         indices;    uint64_t num_indices;    /* global (already rebased, scene.rs:195-199) */
    const ptc_mesh*         meshes;     uint64_t num_meshes;
    const ptc_material*     materials;  uint64_t num_materials;
    const ptc_texture_desc* textures;   uint64_t num_textures;   /* <= MAX_TEXTURES = 1024 (shader_definitions.rs:9) */
} ptc_scene_view;

/* Accumulation semantics. */
enum ptc_accum_mode {
    PTC_ACCUM_MEAN = 0,  /* reference: running mean in rgb, sample count in a (pathtracer.wgsl:32-42) */
    PTC_ACCUM_SUM  = 1   /* rgb += c, a += 1: associative, used for frame-sharded multi-GPU + reduce */
};

/* Light sampling.  The reference ships SEND_TO_LIGHT_PROBABILITY = 0 (shader_definitions.rs:19): emitters are found by
 * BSDF sampling alone, and its light branch (pathtracer/pathtracer.wgsl:21-73,175-191) never applies the weights it
 * computes.  PTC_LIGHTS_MIS is this library's completion of that direction, a NON-PARITY quality mode with
 * the same expected image: at every opaque hit whose path continues, one shadow ray (any-hit traversal, early out) to a
 * uniform point of a uniformly chosen emissive triangle, combined with the BSDF-sampled emitter hits by the balance
 * heuristic.  Definition: include/ptmath.h (pt_ggx_reflect_pdf, pt_light_inv_pdf, pt_mis_weight), DESIGN.md section 9. */
enum ptc_light_sampling {
    PTC_LIGHTS_REFERENCE = 0,
    PTC_LIGHTS_MIS       = 1
};

/* Who builds the acceleration structure (the reference leaves it to the Vulkan driver, pt.rs:364-468).  Images do not
 * depend on the choice: hit identity is defined on triangles, never on the tree. */
enum ptc_bvh_build {
    PTC_BVH_SAH_HOST    = 0,  /* binned SAH on the host cores: best trees, seconds for 10 M triangles (default) */
    PTC_BVH_LBVH_DEVICE = 1,  /* Morton-order linear BVH built by CUDA kernels: milliseconds, slower to trace */
    PTC_BVH_PLOC_DEVICE = 2   /* bottom-up clustering of the Morton-sorted triangles (PLOC) by CUDA kernels: tens of milliseconds */
};

/* compat flags */
#define PTC_COMPAT_NONE 0u

/* The reference's compile-time knobs (src/shaders/shader_definitions.rs:3-26)
 * as run-time configuration. */
typedef struct ptc_config {
    uint32_t width;
    uint32_t height;
    uint32_t max_bounces;              /* MAX_BOUNCES, reference value 4 */
    uint32_t samples_per_pixel;        /* SAMPLES_PER_PIXEL, reference value 1 */
    float    send_to_light_probability;/* SEND_TO_LIGHT_PROBABILITY, reference value 0.0 */
    uint32_t env_texture_index;        /* sky.wgsl:10 hard-codes 9 */
    uint32_t accum_mode;               /* enum ptc_accum_mode */
    uint32_t compat_flags;
    uint32_t bvh_width;                /* 0 = library default; 2 = BVH2 */
    uint32_t light_sampling;           /* enum ptc_light_sampling; PTC_LIGHTS_MIS requires send_to_light_probability == 0 */
    uint32_t bvh_build;                /* enum ptc_bvh_build */
    uint32_t ploc_window;              /* PTC_BVH_PLOC_DEVICE: clusters searched on either side for a merge partner; 0 = default (4), at most 64.
                                          Regular tessellations (the stand-in scenes) build best with 3-4: Morton-aligned blocks; irregular
                                          meshes are reported to gain up to 8-25.  Must be 0 for the other builders. */
    uint32_t reserved[4];
} ptc_config;

/* Planes readable from a step. */
enum ptc_plane {
    PTC_PLANE_COLOR  = 0,  /* float4 per pixel: accumulated rgb + sample count (the `output` storage texture) */
    PTC_PLANE_DEPTH  = 1,  /* float  per pixel: first-hit t (depth_output, r32float) */
    PTC_PLANE_NORMAL = 2,  /* float4 per pixel: first-hit normal-mapped normal, w=1 (normals_output) */
    PTC_PLANE_ALBEDO = 3,  /* float4 per pixel: first-hit albedo, w=1 (albedo_output) */
    PTC_PLANE_HIT_ID = 4,  /* uint32x2 per pixel: primary (instance_index, geometry_index); 0xFFFFFFFF,0xFFFFFFFF = miss */
    PTC_PLANE_DISPLAY = 5,     /* uint8x4 per pixel: tonemapped, sRGB-encoded RGBA (what ptc_blit produced; the swap-chain image) */
    PTC_PLANE_REPROJECTED = 6, /* float4 per pixel: what ptc_rayproject produced (blit_storage_texture) */
    PTC_PLANE_REDUCED = 7      /* float4 per pixel: what ptc_reduce received (sum over ranks of rgb sums + sample counts) */
};

/* Per-step statistics (device counters, read at poll time). */
typedef struct ptc_stats {
    uint64_t rays_traced;      /* trace_ray invocations since create/reset (path segments incl. the escaping one) */
    uint64_t frames_rendered;
    uint64_t kernel_launches;  /* CUDA kernels this library launched */
    uint64_t frame_parts;      /* independent wavefronts a frame is split into (each runs raygen + (trace, shade) x segments) */
    uint64_t packet_node_visits;  /* primary-ray packet traversal: node visits of all packets ... */
    uint64_t packet_lane_visits;  /* ... and lanes (rays) that entered a child at them: coherence = lane / (32 x node) */
    uint64_t reserved[2];
} ptc_stats;

/* status codes */
enum ptc_status {
    PTC_OK              =  0,
    PTC_ERR_INVALID     = -1,
    PTC_ERR_CUDA        = -2,
    PTC_ERR_NO_DEVICE   = -3,
    PTC_ERR_OOM         = -4,
    PTC_ERR_NOT_READY   = -5,   /* ptc_poll(wait=0) while a frame is still in flight */
    PTC_ERR_IO          = -6,
    PTC_ERR_PARSE       = -7
};

#ifdef __cplusplus
}
#endif
#endif /* PTCORE_TYPES_H */
