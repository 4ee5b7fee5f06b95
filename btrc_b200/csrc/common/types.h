// Plain-old-data structures shared by the host runtime and the sm_100a kernels.
//
// Queue layout ("path-state pool"): structure of arrays of 16-byte vectors so that every load and
// store in the wavefront kernels is one coalesced LDG.128/STG.128 per lane.  The reference's SoA
// (src/builtin/btrc/builtin/renderer/wavefront/soa.h:8-37) is kept field for field, except that the
// pixel coordinate (uint2 there) rides in the spare .w lane of the radiance vector.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace btrc {

constexpr float kPi = 3.1415926535f;          // reference: utils/math/scalar.h:8
constexpr float kMaxFloat = 3.402823466e+38f;  // reference: utils/math/scalar.h:7

// reference: wavefront/common.h:12-14
constexpr uint32_t kFlagHasIntersection = 0b01u << 30;
constexpr uint32_t kFlagHasScattering = 0b10u << 30;
constexpr uint32_t kFlagInstanceMask = ~0u << 2 >> 2;

enum MaterialType : int {
    MAT_DIFFUSE = 0,
    MAT_BLACK = 1,
    MAT_DISNEY = 2,
    MAT_METAL = 3,
    MAT_GLASS = 4,
    MAT_MIRROR = 5,
    MAT_INVISIBLE = 6,
    MAT_COUNT
};

// 2-D textures are small postfix programs over a value stack (constant / array leaves, add sub mul div), with the
// chain of "transform" wrappers folded into each array leaf's uv matrix.  Replaces the reference's per-object
// generated sampling code (builtin/texture2d/{array2d,math,transform}.cpp, core/texture2d.cpp).
enum TexOpCode : int { TEXOP_CONST = 0, TEXOP_ARRAY = 1, TEXOP_ADD = 2, TEXOP_SUB = 3, TEXOP_MUL = 4, TEXOP_DIV = 5 };

struct TexOp {
    int op;
    float value[3];                 // TEXOP_CONST
    unsigned long long tex;         // TEXOP_ARRAY: cudaTextureObject_t
    float m[6];                     // TEXOP_ARRAY: uv' = (m0 u + m1 v + m2, m3 u + m4 v + m5)
};

struct TextureDesc {
    int first_op, n_ops;            // program in SceneDevice::tex_ops (n_ops == 1 for a plain constant / array)
};

// texture slots of a material (argument order of the reference's creators: diffuse.cpp:36-48, disney.cpp:844-894,
// metal.cpp:50-77, glass.cpp:141-157, mirror.cpp:91-103)
enum MatSlot : int {
    SLOT_COLOR = 0, SLOT_METALLIC, SLOT_ROUGHNESS, SLOT_TRANSMISSION, SLOT_IOR, SLOT_SPECULAR, SLOT_SPECULAR_TINT, SLOT_ANISOTROPIC,
    SLOT_SHEEN, SLOT_SHEEN_TINT, SLOT_CLEARCOAT, SLOT_CLEARCOAT_GLOSS, SLOT_TRANSMISSION_ROUGHNESS, SLOT_COUNT
};

struct MaterialDesc {
    int type;
    int normal_tex;                 // -1: no normal map
    int shadow_terminator;
    int tex[SLOT_COUNT];            // texture ids, -1 where the type has no such parameter
    // Parameters whose texture is a literal / "constant" node are folded into the table at upload: bit s of const_mask set
    // <=> slot s needs no texture program (its value is scalar_const[s]; colour: color_const).  One struct read replaces a
    // chain of three dependent loads per parameter (material -> texture desc -> program op).
    uint32_t const_mask;
    float color_const[3];
    float scalar_const[SLOT_COUNT];
};

// 3-D textures: constant, a CUDA 3-D texture object (texture3d/array3d.cpp, core/texture3d.cpp) or the sum / product of two
// (texture3d/binary.cpp).  Node table entry: TEX3D_CONST / TEX3D_ARRAY, or TEX3D_PROGRAM with tex = first | count << 32 naming a
// postfix program of TEX3D_CONST / ARRAY / ADD / MUL entries in the same table (behind the nodes).
enum Tex3DKind : int { TEX3D_CONST = 0, TEX3D_ARRAY = 1, TEX3D_ADD = 2, TEX3D_MUL = 3, TEX3D_PROGRAM = 4 };
constexpr int TEX3D_STACK = 4;
struct Texture3DDesc {
    int is_array;  // Tex3DKind
    float value[3];
    unsigned long long tex;
};

enum MediumType : int { MED_HOMOGENEOUS = 0, MED_HETEROGENEOUS = 1, MED_VOLPRIM = 2 };

// builtin/medium/homogeneous.cpp, hetergeneous.cpp; TransformMedium chains (core/medium.cpp:98-129) are composed into
// one world_to_local (3x4 row-major).  The last medium of the table is the volume-primitive medium (empty: Tr = 1).
struct MediumDesc {
    int type;
    float sigma_t, albedo[3], g;            // homogeneous
    int sigma_t_tex, albedo_tex, g_tex;     // heterogeneous
    int has_transform;
    float world_to_local[12];
    float inv_max_density;                  // 1 / max(max sigma_t, 0.001)   (hetergeneous.cpp:33-34)
};

// One entry of the scene file's "volumes" array (factory/scene.cpp:57-80): a parallelepiped o + u x + v y + w z, u,v,w in
// [0,1], with 3-D density and albedo textures addressed by (u,v,w).  core/volume.h:9-45 + core/volume/bvh.h:24-30.
struct VolumePrim {
    float o[3];
    float x[3], y[3], z[3];
    float x_div_x2[3], y_div_y2[3], z_div_z2[3];   // x / |x|^2 ... (inside test, boundary crossings)
    float inv_x2, inv_y2, inv_z2;                  // 1 / |x|^2 ... (world_pos_to_uvw, core/volume.cpp:42-49)
    int sigma_t_tex, albedo_tex;                   // 3-D texture ids
    float max_sigma_t;                             // Texture3D::get_max_float of sigma_t
};

enum EnvType : int { ENV_NONE = 0, ENV_GRADIENT_SKY = 1, ENV_IBL = 2 };

struct EnvLight {  // builtin/light/gradient_sky.cpp, ibl.cpp + env_sampler.cpp
    int type;
    int light_index;                // index in the uniform light sampler (= number of area lights), -1 without env light
    float lower[3], upper[3], up[3];        // gradient sky (up normalised)
    int ibl_tex;                    // texture id (IBL)
    float frame[9];                 // IBL frame x, y, z
    int lut_w, lut_h;
    const float *tile_probs;        // lut_w * lut_h
    const float2 *tile_alias;
};

// reference: core/geometry.h:10-19 -- six float4 tables per mesh
struct GeometryInfo {
    const float4 *gx_ua;    // geometry ex | tex u of vertex a
    const float4 *gy_uba;   // geometry ey | u_b - u_a
    const float4 *gz_uca;   // geometry ez | u_c - u_a
    const float4 *sn_va;    // shading normal a | tex v of a
    const float4 *sn_vba;   // n_b - n_a | v_b - v_a
    const float4 *sn_vca;   // n_c - n_a | v_c - v_a
    // light sampling tables (geometry/triangle_mesh.cpp:108-164)
    const float *positions;        // 9 floats per triangle: a, b-a, c-a
    const float2 *area_alias;      // (accept_prob, bit-cast another_idx)
    int prim_count;
    float total_area;
};

// reference: core/scene.h:21-30.  Matrices are row-major 3x4 (the last row of an affine transform is 0 0 0 1).
struct InstanceInfo {
    int geometry_id;
    int material_id;
    int light_id;
    float mat[12];  // local_to_world
    float inv[12];  // world_to_local
    uint32_t inner_medium_id;
    uint32_t outer_medium_id;
    uint32_t flag;
    // copies of the mesh's table pointers (GeometryInfo): hit reconstruction goes flag -> instance -> tables, one dependent
    // load less than through scene.geometries[geometry_id]
    const float4 *gx_ua, *gy_uba, *gz_uca, *sn_va, *sn_vba, *sn_vca;
};

struct LightInfo {  // MeshLight (builtin/light/mesh_light.cpp)
    int instance_id;
    float intensity[3];
    float inv_scale2;  // 1 / (scale_ * scale_)
    // copies of what light sampling needs from the instance and its mesh (sampling goes light -> alias entry -> triangle:
    // three dependent loads instead of five through instances[] and geometries[])
    int prim_count;
    float total_area;
    const float2 *area_alias;
    const float *positions;
    const float4 *gz_uca;
    float mat[12];  // local_to_world of the instance
    float inv[12];  // world_to_local
};

struct CameraInfo {  // builtin/camera/pinhole.cpp:55-77
    float eye[3];
    float left_bottom_corner[3];
    float film_x[3];
    float film_y[3];
};

struct RenderParams {  // builtin/renderer/wavefront.h:13-27
    int width, height;
    int spp;
    int min_depth, max_depth;
    float rr_threshold, rr_cont_prob;
    int tile;
    int albedo, normal;
    int halton;       // 0: IndependentSampler (PCG), 1: HaltonSampler -- renderer "sampler": "independent" | "halton"
    int filter_type;  // 0 box, 1 gaussian
    float filter_radius;
    float filter_step;  // 2 * radius / GRID_SIZE
};

// Binary BVH over world-space triangles, two children per 64-byte node:
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = same for child 1
//   nz = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   meta = (ref0, ref1, -, -) as int bits
// ref >= 0: inner node index; ref < 0: leaf, ~ref = (first_triangle << 3) | (count - 1).
struct BvhDevice {
    const float4 *nodes;
    const float4 *tris;      // 3 float4 per triangle: (p0 | n.x) (e1 | n.y) (e2 | n.z)
    const uint2 *tri_ids;    // (instance, primitive) per triangle slot
    int root_ref;            // reference of the root (a leaf ref when the scene has <= leaf-size triangles)
    int n_tris;
    // two-level layout (host/bvh.h TwoLevelBvh): `nodes` then holds the object-space trees of the meshes, `tl_nodes` the tree over
    // the instances (root_ref refers to it; its leaf reference ~(instance << 3) names one InstanceBlas record of 4 float4),
    // `tris` / `tri_ids` the world-space triangles per instance in its mesh tree's leaf order
    const float4 *tl_nodes;
    const float4 *inst_blas;
    int two_level;
    // compressed 8-wide layout (host/bvh.h WideBvh8, "bvh_layout": "wide8"): 5 uint4 per node, node 0 = root; `tris` / `tri_ids` are
    // then in the wide tree's leaf order and `nodes` is unused
    const uint4 *wide_nodes;
    int wide;
};

struct HaltonConsts {  // HaltonSampler::initialize_consts (builtin/sampler/halton.cpp:116-133)
    int base_scales[2], base_exponents[2], mult_inverse[2];
    const int *primes;  // the first 1000 primes (utils/prime.cpp)
};

struct SceneDevice {
    BvhDevice bvh;
    const InstanceInfo *instances;
    const GeometryInfo *geometries;
    const MaterialDesc *materials;
    const TextureDesc *textures;
    const TexOp *tex_ops;
    const Texture3DDesc *textures3d;
    const MediumDesc *media;
    const VolumePrim *vol_prims;    // members of the volume-primitive medium (the last entry of `media`)
    int n_vol_prims;
    const LightInfo *lights;        // area lights
    const uint8_t *instance_class;  // shade class of each instance (1 + material type)
    int n_instances;
    int n_lights;                   // area lights + (environment light ? 1 : 0)   (core/light_sampler.cpp:24-33)
    int has_medium;                 // Scene::has_medium (core/scene.cpp:346-356); 2 = and some 3-D texture is an add / mul program
    EnvLight env;
    CameraInfo camera;
    const float2 *filter_alias;  // gaussian filter alias table (256*256)
    uint32_t vol_prim_medium_id;
    float world_diagonal;
    float sort_lo[3], sort_scale[3];  // ray-sort grid: cell = (p - sort_lo) * sort_scale, 8 cells per axis over the world box
    HaltonConsts halton;
};

// Path-state pool (one set = "current", one set = "next").
struct PathQueue {
    float4 *ray_o;   // o.xyz | medium id bits
    float4 *ray_d;   // d.xyz | t1
    float4 *beta;    // beta.rgb | depth bits
    float4 *rad;     // path radiance rgb | pixel index bits (y * W + x)
    ulonglong2 *rng; // PCG state, inc
    float4 *ble;     // beta_le.rgb | bsdf pdf (negative: delta)
    uint32_t *key;   // ray-sort key of the state: direction octant << 9 | 8x8x8 cell of the origin (scratch of the sort kernels)
};

// Shade classes: 0 = miss, 1 + MaterialType = hit on an instance of that material type.  The closest-hit kernel
// appends every state index to its class list; shade runs one kernel instantiation per class (material-sorted queues).
constexpr int kNumShadeClasses = 1 + MAT_COUNT;

struct HitQueue {
    uint32_t *flag;  // instance id | kFlagHasIntersection | kFlagHasScattering
    uint4 *detail;   // t bits, primitive id, u bits, v bits
    uint32_t *class_list;  // kNumShadeClasses lists of `class_stride` state indices each
    int class_stride;
};

struct ShadowQueue {
    float4 *ray_o;   // o.xyz | medium id bits
    float4 *ray_d;   // d.xyz | t1
    float4 *beta_li; // beta_li.rgb | pixel index bits
    uint32_t *key;   // ray-sort key (as PathQueue::key)
};

// Binned ray order (counting sort by key, 4096 bins): the traversal kernels visit surviving paths and shadow rays bin by bin, so that
// the 32 rays of a warp start in the same region of the scene and head into the same octant.
constexpr int kSortBins = 4096;
struct RaySort {
    uint32_t *order;   // sorted ray indices
    int *bin;          // kSortBins + 1 counters: histogram, then exclusive offsets / scatter cursors
};

struct FilmDevice {
    float *radiance;  // W*H*4 (float4 stride, .w unused -- core/film.cpp:55-72)
    float *weight;    // W*H
    float *albedo;    // W*H*4 or null
    float *normal;    // W*H*4 or null
};

// Device-resident wave control block: lets the host enqueue whole iterations without reading anything back.
// BTRC_CTL_LINE: the counters that every warp of a persistent kernel hits with an atomic per 32-entry batch (work cursors, output
// counters, class counts) each start their own 128-byte line, so that the L2 serves them from different lines (and slices) instead
// of queueing every atomic of a kernel -- and every warp's read of n_active -- on ONE line.
// Measured (profiles/r2ab_ab_ctl_lines.md): C2 +1.3 %, C3 +1.5 %, C4 +-0; closest-hit kernel -3.4 % / -5 %.
#ifndef BTRC_CTL_PAD
#define BTRC_CTL_PAD 1
#endif
#if BTRC_CTL_PAD
#define BTRC_CTL_LINE alignas(128)
#else
#define BTRC_CTL_LINE
#endif
struct WaveCtl {
    int n_active;       // states processed by trace/shade this iteration
    int n_survivors;    // states carried over from the previous iteration (slots [0, n_survivors))
    int n_new;          // camera paths generated this iteration (slots [n_survivors, n_active))
    int n_shadow_cur;   // shadow rays to trace this iteration
    // the two output counters share one aligned 64-bit word: a warp reserves its slots in both queues with ONE atomic
    BTRC_CTL_LINE alignas(8) int n_next;  // shade's output counter (reference: StateCounters::active_state_counter)  -- low half
    int n_shadow;       // shade's shadow-ray counter (reference: StateCounters::shadow_ray_counter)       -- high half
    BTRC_CTL_LINE int trace_cursor;   // dynamic work cursors for the persistent traversal kernels
    BTRC_CTL_LINE int shadow_cursor;
    BTRC_CTL_LINE int class_count[8]; // entries of each shade class list this iteration (kNumShadeClasses <= 8); one line per list measured no
                                      // different from one line for all eight (profiles/r2ab_ab_ctl_lines.md)
    BTRC_CTL_LINE int class_cursor[8]; // dynamic work cursors of the shade kernels (one 32-entry batch per grab)
    BTRC_CTL_LINE long long gen_base;   // first camera-path index generated this iteration
    long long cursor;     // next camera-path index
    long long end;        // one past the last camera-path index of this render
    int pool;             // state_count
    int done;
    int sort_enabled;     // the renderer sorts surviving paths and shadow rays before tracing them
    int n_sorted;         // surviving paths whose visiting order comes from RaySort::order (0: queue order)
    int n_shadow_sorted;  // shadow rays visited through the shadow RaySort (0: queue order)
    // Strided partition of the path index space (multi-GPU image tiles, interleaved): the render walks VIRTUAL indices
    // [cursor, end); virtual index v stands for the camera path ((v / stride_block) * stride_parts + stride_part) * stride_block
    // + v % stride_block.  stride_block == 0: identity (one contiguous range).
    long long stride_block;
    int stride_parts, stride_part;
    // 2-D form of the same partition: stride_bw > 0 makes a block a TILE of stride_bw x stride_rows pixels (both divide the image,
    // stride_block = their pixels * spp); tiles are numbered row-major over the tile grid and dealt round-robin, and inside a tile
    // pixels run row-major with all samples of a pixel consecutively.  stride_bw == 0: a block is a contiguous index range (strip).
    int stride_bw, stride_rows;
    int weight_prefilled;  // the film's weight plane was filled in closed form for this render: generate skips its atomic
    // Generation ORDER (not a partition): when the range is whole sample planes [order_s0, order_s0 + order_ns) of every pixel in
    // Uniform indexing, the walk visits them pixel block by pixel block (order_bw x order_bh pixels, all samples of a pixel
    // consecutively) instead of plane by plane.  Every path keeps its (pixel, sample index) and with it its random stream
    // (results are order independent up to the film's float additions); what changes is that the rays of a wave --
    // and the 32 rays of a warp -- start in one image region.  order_ns == 0: queue order = index order.
    int order_s0, order_ns;
    int order_bw, order_bh;   // pixel block of the generation order
    int order_sample_major;   // inside a block: 0 = all samples of a pixel consecutively, 1 = all pixels of the block sample by sample,
                              // 2 = 32 samples of a pixel (one warp), next pixel, ..., then the next 32 samples (needs order_ns % 32 == 0)
    // statistics (SURVEY 8d event counts)
    unsigned long long n_gen, n_closest, n_hit, n_next_total, n_shadow_total, n_shadow_miss, n_splat, n_iter;
};

static_assert(offsetof(WaveCtl, n_next) % 8 == 0 && offsetof(WaveCtl, n_shadow) == offsetof(WaveCtl, n_next) + 4,
              "n_next | n_shadow must form one aligned 64-bit word (emit_outputs reserves both with one atomic)");

}  // namespace btrc
