// Scene creation from a Btrc scene description and commit to device tables.
//
// Mirrors the behaviour of: cli/main.cpp:36-71 (root: width/height/camera/renderer/scene/post_processors),
// factory/scene.cpp:5-94 (entities, envir_light), factory/context.h:205-236 (objects pooled per node, so
// "$reference"d nodes share one object; literal numbers/arrays are constant textures), builtin/register.cpp:31-113
// (the creator names), core/scene.cpp:92-238 (Scene::commit: light list, id maps, instance table, bbox).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../common/types.h"
#include "bvh.h"
#include "mesh.h"
#include "node.h"

namespace btrc {

void cuda_check(cudaError_t e, const char *what);
#define BTRC_CUDA(x) ::btrc::cuda_check((x), #x)

// Owning device allocation.
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t n) { alloc(n); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    DeviceBuffer(DeviceBuffer &&o) noexcept : ptr(o.ptr), bytes(o.bytes) { o.ptr = nullptr; o.bytes = 0; }
    DeviceBuffer &operator=(DeviceBuffer &&o) noexcept { release(); ptr = o.ptr; bytes = o.bytes; o.ptr = nullptr; o.bytes = 0; return *this; }
    ~DeviceBuffer() { release(); }
    void alloc(size_t n);
    void release();
    void upload(const void *src, size_t n);
    void zero();
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

struct PostProcessorDesc {
    std::string type;  // "save" | "tonemap"
    std::string color, albedo, normal;
    float gamma = 1.0f;
    float exposure = 1.0f;
};

struct HostInstance {
    int geometry_id = 0, material_id = 0, light_id = -1;
    Transform3D transform;
    bool has_light = false;
    Vec3 intensity;
    int inner_medium = -1, outer_medium = -1;  // handles into Scene::media (-1: none), ids are assigned by commit()
    uint32_t flag = 0;
};

// An "array" texture: decoded texels + the sampling description (builtin/texture2d/description.cpp:28-80,
// utils/cuda/texture.cpp:63-105, utils/cuda/array.cpp:104-172,313-385).  2-D when depth == 0.
struct TextureSource {
    int width = 0, height = 0, depth = 0;
    int channels = 4;            // 1 or 4
    bool is_float = false;       // false: 8-bit texels in u8 -- UNorm8 (images), or SNorm8 when `snorm` ("byte" text grids)
    bool snorm = false;
    std::vector<uint8_t> u8;
    std::vector<float> f32;
    int address[3] = {0, 0, 0};  // 0 clamp, 1 wrap, 2 mirror, 3 border
    bool linear = true, srgb = false;
    float border[4] = {0, 0, 0, 0};
    float max_value[3] = {0, 0, 0}, min_value[3] = {0, 0, 0};  // Array::find_minmax_values
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
};

// host texture graph node (flattened into TexOp programs at upload)
struct Tex2DNode {
    int type = 0;  // TexOpCode, or 6 = transform wrapper
    Vec3 value;
    int source = -1;
    int lhs = -1, rhs = -1;
    Mat3 m = Mat3::identity();
};

struct Tex3DNode {
    bool is_array = false;
    Vec3 value;
    int source = -1;
    int op = 0;  // 0: leaf (constant / array); TEX3D_ADD / TEX3D_MUL: texture3d/binary.cpp
    int lhs = -1, rhs = -1;
};

struct HostMedium {
    MediumDesc desc{};
    float priority = 0;
};

struct EnvLightDesc {
    int type = 0;  // EnvType
    Vec3 lower, upper, up;
    int tex = -1;
    int lut_w = 400, lut_h = 200;
};

struct RendererDesc {  // builtin/renderer/wavefront.cpp:375-398
    std::string type = "wfpt";
    bool tile = false;
    bool lbvh = false;    // "bvh_builder": "sah" (host binned SAH, default) | "lbvh" (device-side Morton-code build at upload)
    // "bvh_layout": "flat" (instances flattened into one world-space tree) | "instanced" (one object-space tree per mesh under a tree
    // over the instances: bvh.h; saves the per-instance copies of a mesh's nodes, measured slower to trace) | "wide8" (the flat tree
    // collapsed into compressed 8-wide nodes: bvh.h WideBvh8; a fifth of the node bytes, measured slower to trace) | "auto" (default: flat)
    int bvh_layout = 0;   // 0 auto, 1 flat, 2 instanced, 3 wide8
    bool halton = false;  // "sampler": "independent" (default, what the reference's wavefront renderer is compiled with) | "halton"
    int spp = 128;
    int min_depth = 5, max_depth = 10;
    float rr_threshold = 0.1f, rr_cont_prob = 0.5f;
    int state_count = 1000000;
    bool albedo = false, normal = false;
    int filter_type = 0;
    float filter_radius = 1.3f, filter_alpha = 1.6f;
};

class Scene {
public:
    // --- description (filled by the factory)
    int width = 0, height = 0;
    Vec3 cam_eye, cam_dst, cam_up;
    float cam_fov_y_deg = 60;
    RendererDesc renderer;
    std::vector<PostProcessorDesc> post_processors;
    std::vector<Tex2DNode> textures;
    std::vector<Tex3DNode> textures3d;
    std::vector<TextureSource> sources;
    std::vector<MaterialDesc> materials;
    std::vector<HostMedium> media;      // creation order
    EnvLightDesc env;
    std::vector<MeshData> meshes;
    std::vector<HostInstance> instances;
    std::string scene_dir;

    // --- committed state
    std::vector<InstanceInfo> instance_info;
    std::vector<LightInfo> lights;
    std::vector<MediumDesc> medium_table;
    std::vector<VolumePrim> volume_prims;   // "volumes" (core/volume.h), in file order   // id order: descending priority, volume-primitive medium last (core/scene.cpp:136-161)
    bool has_medium = false;
    std::vector<WorldTri> world_tris;
    AABB bbox;
    float world_diagonal = 0;
    SceneDevice dev{};
    FlatBvh2 flat_bvh;
    TwoLevelBvh two_level;       // used instead of flat_bvh when `instanced_bvh`
    bool instanced_bvh = false;
    WideBvh8 wide_bvh;           // used instead of flat_bvh when `wide_bvh8`
    bool wide_bvh8 = false;
    std::vector<AliasUnit> filter_alias;
    std::vector<float> ibl_tile_probs;      // EnvirLightSampler tile table (light/env_sampler.cpp:72-122)
    std::vector<AliasUnit> ibl_alias;
    double bvh_build_seconds = 0;
    float scene_coord_max = 0;   // largest |coordinate| of the world box (BVH box padding)
    AABB scene_box;              // world box incl. the camera (Morton grid of the device builder)
    int lbvh_nodes = 0;          // nodes produced by the device builder
    int bvh_depth = 0;

    Scene() = default;
    ~Scene();
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    static std::unique_ptr<Scene> from_file(const std::string &filename);
    static std::unique_ptr<Scene> from_string(const std::string &json, const std::string &scene_dir);
    void commit();          // host side: tables, world triangles, BVH (no GPU needed)
    void upload_device();   // device side: uploads everything committed (needs a CUDA device)
    RenderParams render_params() const;
    std::string resolve_path(const std::string &p) const;  // factory/path_resolver.cpp:6-17

private:
    void create(const NodeP &root);
    int create_texture2d(const NodeP &node);
    int constant_texture2d(float v);
    int create_texture3d(const NodeP &node);
    float texture3d_extremum(int tex, bool want_max) const;
    int create_material(const NodeP &node);
    int create_medium(const NodeP &node, const Mat4 *outer_world_to_local);
    void create_envir_light(const NodeP &node);
    void parse_texture_desc(const NodeP &node, TextureSource &src) const;
    void release_textures();
    void emit_tex_ops(int node, const Mat3 &m, std::vector<TexOp> &ops) const;
    int create_geometry(const NodeP &node);
    std::map<const Node *, int> tex_pool_, tex3d_pool_, mat_pool_, geo_pool_, med_pool_;
    std::vector<DeviceBuffer> buffers_;
    bool device_ready_ = false;
    template <typename T> const T *upload(const std::vector<T> &v);
};

}  // namespace btrc
