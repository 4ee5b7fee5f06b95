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
    uint32_t inner_medium_id = 0, outer_medium_id = 0;
    uint32_t flag = 0;
};

struct RendererDesc {  // builtin/renderer/wavefront.cpp:375-398
    std::string type = "wfpt";
    bool tile = false;
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
    std::vector<TextureDesc> textures;
    std::vector<MaterialDesc> materials;
    std::vector<MeshData> meshes;
    std::vector<HostInstance> instances;
    std::string scene_dir;

    // --- committed state
    std::vector<InstanceInfo> instance_info;
    std::vector<LightInfo> lights;
    std::vector<WorldTri> world_tris;
    AABB bbox;
    float world_diagonal = 0;
    SceneDevice dev{};
    FlatBvh2 flat_bvh;
    std::vector<AliasUnit> filter_alias;
    double bvh_build_seconds = 0;
    int bvh_depth = 0;

    static std::unique_ptr<Scene> from_file(const std::string &filename);
    static std::unique_ptr<Scene> from_string(const std::string &json, const std::string &scene_dir);
    void commit();          // host side: tables, world triangles, BVH (no GPU needed)
    void upload_device();   // device side: uploads everything committed (needs a CUDA device)
    RenderParams render_params() const;
    std::string resolve_path(const std::string &p) const;  // factory/path_resolver.cpp:6-17

private:
    void create(const NodeP &root);
    int create_texture2d(const NodeP &node);
    int create_material(const NodeP &node);
    int create_geometry(const NodeP &node);
    std::map<const Node *, int> tex_pool_, mat_pool_, geo_pool_;
    std::vector<DeviceBuffer> buffers_;
    bool device_ready_ = false;
    template <typename T> const T *upload(const std::vector<T> &v);
};

}  // namespace btrc
