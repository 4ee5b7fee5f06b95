/* btrc_b200 -- C ABI of the B200-native wavefront render path (libbtrc_b200.so).
 *
 * The reference (AirGuanZ/Btrc) has no FFI: its CLI creates C++ objects through the factory and calls
 * Renderer::render() (src/cli/main.cpp:17-81).  This header is the boundary a binding would use instead; every
 * entry point names the reference interface it stands for.  Conventions: plain pointers and sizes only, no C++
 * or torch types; every function returns 0 on success and non-zero on error, with the chained message (one
 * context per line, like the reference's nested BtrcException printout, src/cli/main.cpp:97-104) available from
 * btrc_last_error() on the calling thread.  One host thread drives one GPU (the current CUDA device).
 *
 * Buffers are caller-owned; `memory` says where they live: BTRC_HOST (pageable or pinned host memory) or
 * BTRC_DEVICE (device memory of the current device, e.g. a torch tensor's data_ptr).
 */
#ifndef BTRC_B200_H
#define BTRC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTRC_HOST 0
#define BTRC_DEVICE 1

typedef struct btrc_scene btrc_scene_t;

typedef struct btrc_scene_info {
    int32_t width, height;        /* root["width"], root["height"]            (cli/main.cpp:50-51)          */
    int32_t spp;                  /* renderer params                          (renderer/wavefront.h:13-27)  */
    int32_t min_depth, max_depth;
    int32_t state_count;
    int32_t tile, albedo, normal;
    int32_t n_instances, n_lights, n_triangles, n_materials;
    int32_t bvh_depth;
    int32_t n_media;              /* media incl. the volume-primitive medium  (core/scene.cpp:136-161)     */
    int32_t has_medium;           /* Scene::has_medium                        (core/scene.cpp:346-356)     */
    int32_t n_textures;
    float world_diagonal;         /* 1.2 * |bbox(scene + eye)|                (wavefront.cpp:138-139)       */
    float bvh_build_seconds;
    uint64_t bvh_bytes;           /* nodes + triangles the traversal kernels walk (compare with the 126 MB L2)           */
    int32_t bvh_instanced;        /* 1: two-level layout (one object-space tree per mesh under a tree over the instances) */
    int32_t bvh_wide;             /* 1: compressed 8-wide layout ("bvh_layout": "wide8"; 80-byte nodes, 8-bit child boxes)   */
} btrc_scene_info_t;

typedef struct btrc_render_stats {
    uint64_t n_gen;          /* camera paths generated                                         */
    uint64_t n_closest;      /* closest-hit rays traced                                        */
    uint64_t n_hit;          /* ... of which hit something                                     */
    uint64_t n_next;         /* continuation states emitted by shade                           */
    uint64_t n_shadow;       /* shadow rays emitted by shade                                   */
    uint64_t n_shadow_miss;  /* ... of which reached the light                                 */
    uint64_t n_splat;        /* radiance splats (3 float atomics each)                         */
    uint64_t n_iter;         /* wavefront iterations                                           */
    double render_ms;        /* device time of the last btrc_render_range (CUDA events)        */
    double stage_ms[6];      /* with btrc_set_profiling(1): device time per stage of the last  */
                             /* range -- generate, closest-hit trace, shade, shadow trace,     */
                             /* medium, iteration bookkeeping (counter snapshot + prepare)     */
    uint64_t n_launches;     /* kernels launched by the last btrc_render_range                 */
} btrc_render_stats_t;

/* Chained error text of the last failing call on this thread ("" if none). */
const char *btrc_last_error(void);

/* parse + create + commit: cli/main.cpp:28-71 (JSONParser, create_scene, camera, renderer, ObjectDAG::commit).
 * Loads OBJ files, builds the BVH and uploads all tables to the current device. */
int btrc_scene_load_json(const char *path, btrc_scene_t **out);
/* Same from memory; scene_dir is what "$env{scene_directory}" and "$include{}" resolve against. */
int btrc_scene_load_json_string(const char *json, const char *scene_dir, btrc_scene_t **out);
void btrc_scene_free(btrc_scene_t *scene);
int btrc_scene_info_get(btrc_scene_t *scene, btrc_scene_info_t *out);

/* Renderer parameter overrides (WavefrontPathTracer::set_params, wavefront.cpp:67-70): spp, state_count,
 * min_depth, max_depth.  Call before the first render; pass -1 to keep a value. */
int btrc_set_params(btrc_scene_t *scene, int32_t spp, int32_t state_count, int32_t min_depth, int32_t max_depth);

/* Record one CUDA event between consecutive stage kernels so that btrc_render_stats_t::stage_ms is filled
 * (the reference has no profiling hooks: SURVEY 5). */
int btrc_set_profiling(btrc_scene_t *scene, int32_t on);

/* Reporter (core/reporter.h:7-36, Renderer::set_reporter, core/renderer.h:28) + cancellation (Renderer::stop / should_stop,
 * core/renderer.h:44-48): `fn` is called from the rendering thread only -- new_stage as fn(0), then after every wave whose
 * counters have arrived with the percentage of finished camera paths (wavefront.cpp:286-287), complete_stage as fn(100).
 * A non-zero return value stops the render after the waves in flight: btrc_render / btrc_render_range then return
 * BTRC_STOPPED (2) and write no image, like the reference's empty RenderResult (wavefront.cpp:308-316).  NULL removes it. */
#define BTRC_STOPPED 2
typedef int (*btrc_progress_fn)(float percentage, void *user);
int btrc_set_reporter(btrc_scene_t *scene, btrc_progress_fn fn, void *user);

/* Renderer::render() (wavefront.cpp:205-325): clears the film, traces all spp*W*H camera paths and resolves
 * value/weight (preview.cpp:28-80).  color is W*H float4 (row 0 = top, .w = 1); albedo / normal may be NULL and
 * are written only when the renderer was created with "albedo"/"normal": true. */
int btrc_render(btrc_scene_t *scene, int memory, float *color, float *albedo, float *normal, btrc_render_stats_t *stats);

/* The same render split into its parts, for multi-GPU sample partitioning: each rank renders a
 * contiguous range of camera-path indices into its own film, the films are summed (NCCL) and resolved once.
 * Path index space: Uniform mode index = sample * W*H + pixel, Tile mode index = pixel * spp + sample
 * (generate.cpp:49-60). */
int btrc_film_clear(btrc_scene_t *scene);
int btrc_render_range(btrc_scene_t *scene, int64_t path_begin, int64_t path_end, btrc_render_stats_t *stats);
/* This device's share of the render when n_parts devices split it (the reference is single-GPU, cli/main.cpp:23).
 * The partition follows the renderer's "tile" flag: sample-index ranges of every pixel ("tile": false, generate.cpp:49-55), or
 * image tiles dealt round-robin to the parts ("tile": true, the pixel-major index space of generate.cpp:56-60): tiles of
 * 64 x rows_per_strip pixels when they divide the image, strips of rows_per_strip whole pixel rows otherwise; rows_per_strip <= 0
 * picks a default (at most 16).  Adds to this device's film like btrc_render_range. */
int btrc_render_part(btrc_scene_t *scene, int32_t n_parts, int32_t part, int32_t rows_per_strip, btrc_render_stats_t *stats);
/* The path's one collective: ncclReduce(sum) of the film planes (radiance, weight [, albedo, normal]) onto rank `root`, in
 * place, on the renderer's stream -- btrc_resolve on the root then follows in stream order.  btrc_nccl_unique_id fills the 128
 * bytes of an ncclUniqueId on one rank; every rank passes them to btrc_comm_create (ncclCommInitRank).  NCCL is loaded at first
 * use; with n_ranks == 1 no NCCL is needed and the reduce is a no-op. */
typedef struct btrc_comm btrc_comm_t;
int btrc_nccl_unique_id(uint8_t *out128);
int btrc_comm_create(int32_t n_ranks, int32_t rank, const uint8_t *id128, btrc_comm_t **out);
void btrc_comm_destroy(btrc_comm_t *comm);
int btrc_reduce_films(btrc_scene_t *scene, btrc_comm_t *comm, int32_t root);
/* Film::get_float3_output / get_float_output (core/film.cpp): radiance W*H*4 floats, weight W*H floats,
 * albedo / normal W*H*4 floats or NULL. */
int btrc_film_export(btrc_scene_t *scene, int memory, float *radiance, float *weight, float *albedo, float *normal);
int btrc_film_import(btrc_scene_t *scene, int memory, const float *radiance, const float *weight, const float *albedo, const float *normal);
int btrc_resolve(btrc_scene_t *scene, int memory, float *color, float *albedo, float *normal);

/* PostProcessor::process for every entry of root["post_processors"] (cli/main.cpp:79-80; "save": gamma on colour,
 * format by extension -- postprocess/save_to_image.cpp:30-62; "tonemap": ACES * exposure -- tonemap.cpp:14-51). */
int btrc_run_post_processors(btrc_scene_t *scene, int memory, float *color, float *albedo, float *normal);

/* ImageDynamic::load (utils/image.cpp:116-215; formats by extension: .png .jpg/.jpeg .hdr): the decoder the texture creators use.
 * First call with out = NULL to get the size, then with a buffer of width*height*channels bytes (8-bit formats) or
 * width*height*3 floats (.hdr, *is_float = 1).  Host memory, no device involved. */
int btrc_image_load(const char *path, int32_t *width, int32_t *height, int32_t *channels, int32_t *is_float, void *out, uint64_t out_bytes);

/* The scene-file layer on its own (factory/node/parser.cpp:16-252: comments, "$include{..}", "$reference{..}", the std::to_string(float)
 * round trip; factory/node/node.cpp:14-325: typed values), host only - what the parity tests compare with the reference's own parser
 * built under oracle/_ref.  btrc_scene_file_dump writes the resolved tree as text: groups {"key":v,...} in first-insertion key order,
 * arrays [v,...], values "text"; call with out = NULL to get the size in *needed.  btrc_scene_file_parse reads the node at node_path
 * ("a/b/0/c" from the root) as Node::parse<T>: kind 0 int, 1 float, 2 bool, 3 Vec3f, 4 Spectrum, 5 Degree, 6 Radian,
 * 7 Transform3D (16 floats, row-major forward matrix), 8 Transform2D (9 floats, row-major); *count = floats written. */
int btrc_scene_file_dump(const char *path, char *out, uint64_t out_bytes, uint64_t *needed);
int btrc_scene_file_parse(const char *path, const char *node_path, int32_t kind, float *out, int32_t *count);

/* ---- stage-level entry points on explicit batches, in the reference's SoA layout (wavefront/soa.h:8-37) ----
 * ray: o_med = float4(o.xyz, bit-cast medium id), d_t1 = float4(d.xyz, t1).
 * hit: flag = instance id | bit30 on hit (common.h:12-14); detail = uint4(t bits, primitive id, u bits, v bits).
 * path: pixel uint2, beta_depth float4(beta.rgb, bit-cast depth), path_rad float4(L.rgb, 1), rng 2 x uint64
 *       (PCG state, inc), ble float4(beta_le.rgb, bsdf pdf; negative = delta). */

/* TracePipeline::trace (trace.cpp:127-142) */
int btrc_closest_hit(btrc_scene_t *scene, int memory, const float *o_med, const float *d_t1, int32_t n, uint32_t *flag, uint32_t *detail);
/* occlusion part of ShadowPipeline::test (shadow.cpp:24-36): occluded[i] = 1 when any triangle has t in [0, t1] */
int btrc_any_hit(btrc_scene_t *scene, int memory, const float *o_med, const float *d_t1, int32_t n, uint8_t *occluded);
/* GeneratePipeline::generate (generate.cpp:140-177) for camera-path indices [first, first + n); also adds the
 * film weights. */
int btrc_generate(btrc_scene_t *scene, int memory, int64_t first, int32_t n, float *o_med, float *d_t1, uint32_t *pixel,
                  float *beta_depth, float *path_rad, uint64_t *rng, float *ble);
/* ShadePipeline::shade (shade.cpp:263-283) on n states: appends survivors to the next_* arrays and shadow rays to
 * the shadow_* arrays (any order; counts returned), splats terminated paths into the film. */
int btrc_shade(btrc_scene_t *scene, int memory, int32_t n,
               const float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth, const float *path_rad,
               const uint64_t *rng, const float *ble, const uint32_t *flag, const uint32_t *detail,
               int32_t *n_next, float *next_o_med, float *next_d_t1, uint32_t *next_pixel, float *next_beta_depth,
               float *next_path_rad, uint64_t *next_rng, float *next_ble,
               int32_t *n_shadow, uint32_t *shadow_pixel, float *shadow_beta_li, float *shadow_o_med, float *shadow_d_t1);
/* MediumPipeline::sample_scattering (medium.cpp:204-211) on n states, run between closest-hit and shade when the scene
 * has media.  Updated IN PLACE like the kernel updates its own slot: o_med.w (resolved medium id), ble (beta_le * Tr of
 * the unscattered segment), rng (advanced sampler; only where no scattering happened), flag (bit 31 = scattered).
 * Scattered paths append their continuation to the next_* arrays and their NEE ray to the shadow_* arrays. */
int btrc_medium(btrc_scene_t *scene, int memory, int32_t n,
                float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth, const float *path_rad,
                uint64_t *rng, float *ble, uint32_t *flag, const uint32_t *detail,
                int32_t *n_next, float *next_o_med, float *next_d_t1, uint32_t *next_pixel, float *next_beta_depth,
                float *next_path_rad, uint64_t *next_rng, float *next_ble,
                int32_t *n_shadow, uint32_t *shadow_pixel, float *shadow_beta_li, float *shadow_o_med, float *shadow_d_t1);
/* ShadowPipeline::test (shadow.cpp:143-158): traces the shadow rays and splats the unoccluded ones into the film. */
int btrc_shadow(btrc_scene_t *scene, int memory, int32_t n, const uint32_t *shadow_pixel, const float *shadow_beta_li,
                const float *shadow_o_med, const float *shadow_d_t1);

/* Table access for parity tests: the six float4 shading tables of geometry `geometry_id` (24 floats per triangle,
 * geometry/triangle_mesh.cpp:40-106) and the instance transforms (local_to_world, world_to_local; 3x4 row-major). */
int btrc_geometry_tables(btrc_scene_t *scene, int32_t geometry_id, int32_t *n_triangles, float *out24);
/* TriangleMeshLoader::get_positions (utils/triangle_mesh_loader.cpp:61-123, after transform_to_unit_cube :193-213 when the geometry asks for it):
 * the local-space corner positions of the mesh, 9 floats per triangle in primitive order.  Host side only. */
int btrc_geometry_positions(btrc_scene_t *scene, int32_t geometry_id, int32_t *n_triangles, float *out9);
int btrc_instance_transform(btrc_scene_t *scene, int32_t instance_id, float *mat12, float *inv12);
/* InstanceInfo::geometry_id (core/scene.h:21-30): which mesh the entity instantiates.  Host side only. */
int btrc_instance_geometry(btrc_scene_t *scene, int32_t instance_id, int32_t *geometry_id);
/* EnvirLightSampler::preprocess result (light/env_sampler.cpp:72-122): the normalised lut_w * lut_h tile table of the "ibl"
 * environment light, computed on the device at commit.  probs may be NULL to query the size. */
int btrc_ibl_table(btrc_scene_t *scene, int32_t *lut_w, int32_t *lut_h, float *probs);

/* The CUDA stream (cudaStream_t) all of this scene's kernels are launched on, so that a caller can bracket calls
 * with its own events on the launching stream. */
void *btrc_stream(btrc_scene_t *scene);

/* Library / device facts.  btrc_set_device = cuda::Context(device_index) of cli/main.cpp:23 (one process per GPU: call it
 * before the scene is loaded). */
int btrc_device_sm_count(void);
int btrc_device_count(void);
int btrc_wave_ctl_bytes(void);   /* size of the wave control block: the only per-render upload, and what comes back per progress snapshot */
int btrc_set_device(int32_t device);
const char *btrc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BTRC_B200_H */
