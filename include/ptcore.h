/* ptcore.h — C ABI of the B200 path-tracing core.
 *
 * This is the drop-in boundary for the reference's path-tracing compute pass:
 * `impl RenderStep for RTStep` (reference: src/render_steps/mod.rs:20-25,
 * src/render_steps/pt.rs:341-565,732-745) as called by `Raytracer`
 * (src/raytracer.rs:84-91,177-195,351-357).  One opaque `ptc_step` per GPU.
 * Plain pointers and sizes only; no CUDA, torch or C++ types.  All functions
 * return 0 (PTC_OK) or a negative `enum ptc_status`; the message for the last
 * failure on the calling thread is `ptc_last_error()`.  Nothing here ever
 * falls back to a CPU implementation: without a CUDA device, ptc_create
 * fails with PTC_ERR_NO_DEVICE.  No C++ exception crosses this boundary and
 * nothing aborts: host allocation failures come back as PTC_ERR_OOM, damaged
 * scene files and inconsistent arrays as PTC_ERR_INVALID / PTC_ERR_PARSE with
 * a message (the Rust shim turns them into the reference's panics).
 * Triangles with a non-finite corner, or a coordinate beyond 1e37, are
 * inactive (never hit), as in the reference's Vulkan acceleration structures.
 *
 * The host-side helpers at the bottom (camera, scene ingest) restate
 * src/camera.rs and src/scene.rs so a caller without the Rust crate can
 * produce the same arrays.
 */
#ifndef PTCORE_H
#define PTCORE_H

#include "ptcore_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PTC_API
#else
#define PTC_API __attribute__((visibility("default")))
#endif

typedef struct ptc_step ptc_step;

/* ---- RenderStep lifecycle ------------------------------------------------ */

/* replaces RTStep::create (pt.rs:341-490) + create_static_bind_groups
 * (pt.rs:188-337): copies the scene arrays to the device, builds this
 * library's own BVH over every mesh (the reference builds one driver BLAS per
 * mesh and one TLAS, pt.rs:364-468), builds the emissive-triangle list
 * (pt.rs:206-243), creates bindless texture objects (texture.rs:232-336) and
 * allocates the accumulation + aux planes.  The scene view is not retained. */
PTC_API int ptc_create(const ptc_scene_view* scene, const ptc_config* config, int cuda_device, ptc_step** out);

/* replaces RTStep::update (pt.rs:492-521): latch the 112-byte Uniforms for the
 * next render.  `num_lights` is overwritten with the step's own light count
 * exactly as pt.rs:243 does.  Host pointer; copied before return. */
PTC_API int ptc_update(ptc_step* step, const ptc_uniforms* uniforms);

/* replaces RTStep::render (pt.rs:523-565) + rt_queue.submit (raytracer.rs:195):
 * enqueue one frame (SAMPLES_PER_PIXEL samples for every pixel, accumulated
 * into the COLOR plane, aux planes overwritten) on the step's stream and
 * return immediately. */
PTC_API int ptc_render(ptc_step* step);

/* replaces rt_device.poll(Poll / Wait) (raytracer.rs:165-167,319): wait != 0
 * blocks until every enqueued frame has finished; wait == 0 returns
 * PTC_ERR_NOT_READY while work is in flight. */
PTC_API int ptc_poll(ptc_step* step, int wait);

/* replaces RTStep::resize (pt.rs:732-745) + the accumulation-texture
 * re-creation in Raytracer::resize (raytracer.rs:344): reallocates and zeroes
 * every plane.  Buffers injected with ptc_set_output / ptc_set_reduce_output are
 * dropped (their size is the old one): the step renders into its own planes
 * until the caller injects new ones.  A bad size fails without touching the
 * step; if allocation fails the previous size is restored (accumulation
 * cleared), and if even that fails every later call reports the step invalid. */
PTC_API int ptc_resize(ptc_step* step, uint32_t width, uint32_t height);

/* the analogue of `step.output_texture_view = ...` (raytracer.rs:86): the
 * caller injects the accumulation target, a device pointer to width*height
 * float4 (row-major, tightly packed).  NULL restores the step-owned buffer. */
PTC_API int ptc_set_output(ptc_step* step, void* device_rgba32f);

/* zero the accumulation plane (what recreating the texture does) */
PTC_API int ptc_reset_accumulation(ptc_step* step);

/* ---- display path: the two passes that follow the path tracer in Raytracer::update (raytracer.rs:215-241) ---- */

/* replaces BlitStep::render (src/render_steps/blit.rs:176-205, shaders/blit/blit.wgsl:35-40, shaders/common/color.wgsl:4-27):
 * Khronos PBR-neutral tonemap of an rgba32f image + the surface's sRGB encode (wgpu_util.rs:235-238) into the 8-bit
 * DISPLAY plane.  src = device pointer to width*height float4, NULL = the COLOR plane.  Ordered after the frames already
 * enqueued.  With src = NULL right after ptc_render the tonemap runs per frame part on the parts' own streams, so it
 * does not hold the next frame back (the reference blits on its second queue while the RT queue renders on). */
PTC_API int ptc_blit(ptc_step* step, const void* src_device_rgba32f);

/* replaces RayprojectStep::render (src/render_steps/rayproject.rs:205-249, shaders/rayproject/rayproject.wgsl:45-74):
 * forward-splats every pixel of the last real frame (src, NULL = COLOR) by its first-hit distance (DEPTH plane, rendered
 * with `prev`) to where it lands under camera `cur`, into the REPROJECTED plane.  Targets nothing lands on keep their
 * previous content (the reference never clears blit_storage_texture); where several sources land on one target the
 * reference lets the stores race — here the highest source pixel index wins. */
PTC_API int ptc_rayproject(ptc_step* step, const ptc_uniforms* prev, const ptc_uniforms* cur, const void* src_device_rgba32f);

PTC_API void ptc_destroy(ptc_step* step);

/* ---- results -------------------------------------------------------------- */

/* synchronous device->host copy of one plane (enum ptc_plane) after all
 * enqueued frames; dst_pitch_bytes = 0 means tightly packed. */
PTC_API int ptc_read(ptc_step* step, int plane, void* dst, size_t dst_pitch_bytes);
/* the reverse for PTC_PLANE_COLOR: load an accumulation image (rgb mean or sum + sample count, as ptc_read returned it)
 * after all enqueued frames.  Checkpoint / resume of a progressive render (the reference keeps its
 * accumulation only in a GPU texture and cannot); continuing with the following frame numbers reproduces the uninterrupted
 * run bit for bit.  src_pitch_bytes = 0 means tightly packed. */
PTC_API int ptc_write(ptc_step* step, int plane, const void* src, size_t src_pitch_bytes);

/* asynchronous variant for display / streaming: snapshots the plane on the
 * render stream (device->device), then copies the snapshot to `dst_pinned` on
 * a second stream so following frames are not held up.  Completion is
 * observed with ptc_poll(step, 1) or ptc_read_wait. */
PTC_API int ptc_read_async(ptc_step* step, int plane, void* dst_pinned);
PTC_API int ptc_read_wait(ptc_step* step);

/* device-to-device copy of a plane into a caller-owned buffer with the caller's row pitch, enqueued on the step's stream —
 * what DenoiseStep does before OIDN (src/render_steps/denoise.rs:84-100 copy_texture_to_buffer into a buffer whose rows
 * are rounded up to 256 bytes, pt.rs:470-472).  dst_pitch_bytes = 0 means tightly packed.  The NORMAL / ALBEDO planes
 * themselves are tightly packed float4 rows, exactly as the reference's shader writes them (index y*W + x,
 * pathtracer/pathtracer.wgsl:272-273) and as OIDN reads them there (pixel stride 16, row stride 0, denoise.rs:137-160). */
PTC_API int ptc_copy_plane(ptc_step* step, int plane, void* dst_device, size_t dst_pitch_bytes);

/* zero-copy interop: device address of a plane (valid until resize/destroy) */
PTC_API void* ptc_device_ptr(ptc_step* step, int plane);

/* the CUDA stream frames are enqueued on (a cudaStream_t), for callers that
 * order their own work after it */
PTC_API void* ptc_stream(ptc_step* step);

/* Multi-GPU: frames are sharded over one step per GPU, each accumulating with PTC_ACCUM_SUM (any other mode
 * is refused with PTC_ERR_INVALID: running means cannot be summed).  This enqueues, on the step's stream and after the
 * frames already enqueued, ONE ncclReduce(ncclFloat32, ncclSum) of the COLOR planes (rgb sums + sample counts,
 * width*height*4 floats) of every rank of `nccl_comm` into the REDUCED plane of rank `root` (PTC_PLANE_REDUCED: read it
 * with ptc_read / ptc_device_ptr, or inject the receive buffer with ptc_set_reduce_output).  The COLOR plane — each
 * rank's own accumulation — is left untouched, so the call may be repeated: a reduce after more frames receives every
 * rank's whole history exactly once (an in-place reduce would add the earlier ranks' sums again).  On ranks other than
 * the root the REDUCED plane's content is unspecified.  root < 0: ncclAllReduce instead (every rank receives the sum,
 * e.g. to display it).  The reference has no counterpart (single GPU).
 * `nccl_comm` is the caller's ncclComm_t.  The call binds to the NCCL that is ALREADY in the process — the loaded
 * libnccl.so.2 (dlopen RTLD_NOLOAD), else a global ncclReduce symbol (static NCCL) — and never loads another copy, so
 * communicator and call always come from the same NCCL; PTC_ERR_INVALID when there is none. */
PTC_API int ptc_reduce(ptc_step* step, void* nccl_comm, int root);
/* inject the receive buffer of ptc_reduce: a device pointer to width*height float4 (NULL restores the step-owned plane,
 * allocated by the first ptc_reduce).  Dropped by ptc_resize, like the buffer of ptc_set_output. */
PTC_API int ptc_set_reduce_output(ptc_step* step, void* device_rgba32f);

PTC_API int ptc_get_stats(ptc_step* step, ptc_stats* out);

/* per-stage device timing with CUDA events recorded on the render stream around every launch
 * (the reference has no counterpart: all its passes set timestamp_writes: None, pt.rs:550).
 * ms / launches: [0] raygen, [1] trace (closest-hit traversal), [2] shade, summed since enabling. */
PTC_API int ptc_set_profiling(ptc_step* step, int enable);
PTC_API int ptc_get_profile(ptc_step* step, double ms[3], uint64_t launches[3]);
PTC_API int ptc_num_lights(ptc_step* step, uint32_t* out);

/* BVH facts for reports: out[0]=nodes, out[1]=triangles, out[2]=node bytes,
 * out[3]=triangle bytes, out[4]=max depth, out[5]=build milliseconds, out[6]=1 when built by CUDA kernels
 * (ptc_config.bvh_build = PTC_BVH_LBVH_DEVICE), out[7]=build microseconds */
PTC_API int ptc_bvh_info(ptc_step* step, uint64_t out[8]);
/* FNV-1a hashes of the device-resident node array (out[0]) and leaf-ordered triangle array (out[1]), read back from
 * the GPU (unit-test entry: the device-built tree must equal the host restatement of the same build word for word). */
PTC_API int ptc_debug_bvh_hash(ptc_step* step, uint64_t out[2]);

/* trace caller-supplied rays (closest hit) — unit-test / tooling entry point
 * for the traversal kernel.  rays: n x 8 floats (ox,oy,oz,_,dx,dy,dz,_) host
 * memory; hits: n x 4 words (t, b1, b2 as f32; global triangle id or
 * 0xFFFFFFFF as u32); ids: optional n x 2 (instance_index, geometry_index). */
PTC_API int ptc_trace_rays(ptc_step* step, const float* rays, uint64_t n, uint32_t* hits, uint32_t* ids);
/* same through the any-hit kernel: occluded[i] = 1 iff some triangle lies in (tmin, tmax_i);
 * rays[i*8+7] carries tmax_i (<= 0 means PT_RAY_TMAX) */
PTC_API int ptc_trace_rays_any(ptc_step* step, const float* rays, uint64_t n, uint8_t* occluded);

/* evaluate one of the shared numerics primitives of include/ptmath.h ON THE DEVICE for n inputs (unit-test
 * entry: device results must equal the host's bit for bit).  fn: 0 sin, 1 cos, 2 atan, 3 acos, 4 atan2(a,b),
 * 5 a/b, 6 sqrt, 7 fma(a,b,a), 8 rand_float of state bits a, 9 normalize probe, 10 refract probe,
 * 11 a*b+a (uncontracted), 12 bilinear address probe. */
PTC_API int ptc_debug_math(int cuda_device, int fn, const float* a, const float* b, float* out, uint64_t n);

/* evaluate the kernels' texture lookup on the device (unit-test entry): texture_index >= 0 samples that texture at
 * n (u, v) pairs (`in` = n x 2 floats) exactly as sample_texture_rgba does (helpers.wgsl:35-37); texture_index < 0
 * evaluates sky_color (sky.wgsl:1-12) for n directions (`in` = n x 3 floats).  rgba = n x 4 floats.  per_lane != 0:
 * lookup i uses texture (texture_index + i) mod num_textures, so the lanes of a warp hold different texture handles,
 * as they do in the shading kernel. */
PTC_API int ptc_debug_sample(ptc_step* step, int texture_index, const float* in, uint64_t n, float* rgba, int per_lane);

/* the texture unit's u8 -> f32 conversion (cudaReadModeNormalizedFloat) for a 256x1 RGBA8 ramp: out[i] = .x of texel i
 * (value i), out[256+i] = .w (value 255-i).  Returns 1 when the library reads 8-bit textures through it (every value
 * equals i/255 rounded to nearest, the reference's to_rgba32f, texture.rs:241), 0 when it divides in the kernel, < 0 on error. */
PTC_API int ptc_debug_unorm8(int cuda_device, float* out512);

/* measured L2 read bandwidth of the device in GB/s: a persistent grid streams an L2-resident buffer (`bytes`, e.g. 32 MiB
 * on a 126 MB L2) `iters` times with 128-bit loads; CUDA-event timed.  The second denominator of the traversal roofline:
 * every configuration but San Miguel keeps its BVH in L2. */
PTC_API int ptc_debug_l2_bandwidth(int cuda_device, uint64_t bytes, uint32_t iters, double* gbs);

PTC_API const char* ptc_last_error(void);
PTC_API const char* ptc_version(void);
PTC_API int ptc_device_count(void);

/* ---- host side: camera (reference: src/camera.rs:31-79) ------------------- */

typedef struct ptc_host_camera {
    uint32_t width, height;
    float    fov;
    float    position[3];
    float    lookat[3];
    uint32_t frame;
    uint32_t samples_per_pixel, max_bounces;
    float    pixel_space_x[3], pixel_space_y[3], first_pixel_pos[3];
    float    theta_x, theta_y, camera_speed;
} ptc_host_camera;

/* Camera::new (camera.rs:31-52), including its one call to update() */
PTC_API void ptc_camera_new(ptc_host_camera* cam, uint32_t width, uint32_t height, float fov_deg, const float position[3]);
/* Camera::update (camera.rs:54-79): viewport basis from the CURRENT lookat, then lookat from theta */
PTC_API void ptc_camera_update(ptc_host_camera* cam);
/* uniforms::Uniforms::new (uniforms.rs:30-50,64-71) */
PTC_API void ptc_camera_uniforms(const ptc_host_camera* cam, ptc_uniforms* out);

/* ---- host side: scene ingest (reference: src/scene.rs:77-394) -------------- */

typedef struct ptc_scene ptc_scene;

enum ptc_texture_source { PTC_TEX_CONSTANT = 0, PTC_TEX_FILE = 1, PTC_TEX_TEXELS = 2 };

typedef struct ptc_scene_texture_info {
    uint32_t source;         /* enum ptc_texture_source */
    uint32_t width, height;  /* 1x1 for constants; 0x0 for files not decoded yet */
    uint32_t format;
    float    color[4];       /* constants */
    const char* path;        /* files: joined with the OBJ's directory (scene.rs:281-283) */
} ptc_scene_texture_info;

/* Scene::new (scene.rs:77-122): textures 0..8 = default constants, texture 9 =
 * the environment map (reference: "res/knob/envmap.exr"), material 0 = default. */
PTC_API int ptc_scene_new(const char* envmap_path, ptc_scene** out);
/* scene.rs:202-215 computes missing normals with mesh-local indices into the global vertex array (a bug
 * that corrupts every mesh after the first).  Default 0 = reproduce it (parity); 1 = apply the offset. */
PTC_API int ptc_scene_set_fix_normals(ptc_scene* scene, int enable);
/* Scene::load_obj (scene.rs:124-257) with tobj 4.0.3 GPU_LOAD_OPTIONS semantics */
PTC_API int ptc_scene_load_obj(ptc_scene* scene, const char* obj_path);
/* same from memory (obj_dir is used to resolve texture paths) */
PTC_API int ptc_scene_load_obj_text(ptc_scene* scene, const char* obj_text, const char* mtl_text, const char* obj_dir);
PTC_API uint64_t ptc_scene_num_textures(const ptc_scene* scene);
PTC_API int ptc_scene_texture_info_at(const ptc_scene* scene, uint64_t index, ptc_scene_texture_info* out);
/* hand decoded texels for a FILE texture (what Texture::load does with `image::open`, texture.rs:232-243);
 * the scene copies them */
PTC_API int ptc_scene_set_texels(ptc_scene* scene, uint64_t index, uint32_t width, uint32_t height, uint32_t format, const void* texels);
/* Texture::load for every FILE texture without texels (texture.rs:232-243: image::open + to_rgba32f), on n_threads host
 * threads (0 = all; the reference spawns 32, scene.rs:239-254).  PNG (all colour types, bit depths, Adam7), OpenEXR
 * (single-part scan-line, HALF / FLOAT, NONE / ZIPS / ZIP / PIZ), Radiance HDR (RGBE, run-length or flat) and PFM are decoded on the host; JPEG (baseline and
 * progressive) by nvJPEG on `cuda_device`, bound at run time — without libnvjpeg or a GPU only JPEG textures fail.
 * 8-bit sources stay RGBA8, 16-bit become RGBA32F = v / 65535, float sources stay float.  Returns the
 * first failure (ptc_last_error names the file); textures that did decode keep their texels. */
PTC_API int ptc_scene_decode_textures(ptc_scene* scene, int cuda_device, int n_threads);
/* the same decoders on one image held in memory; *texels is allocated by the library: release it with ptc_free */
PTC_API int ptc_decode_image(const void* data, size_t bytes, int cuda_device, uint32_t* width, uint32_t* height, uint32_t* format, void** texels);
PTC_API void ptc_free(void* p);
/* borrow the arrays; every FILE texture must have texels by now */
PTC_API int ptc_scene_get_view(const ptc_scene* scene, ptc_scene_view* out);
PTC_API void ptc_scene_free(ptc_scene* scene);

#ifdef __cplusplus
}
#endif
#endif /* PTCORE_H */
