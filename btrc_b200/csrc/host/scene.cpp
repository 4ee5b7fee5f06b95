#include "scene.h"

#include <chrono>
#include <filesystem>
#include <functional>

namespace btrc {

void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw Error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}

void DeviceBuffer::alloc(size_t n) {
    release();
    bytes = n;
    if (n) BTRC_CUDA(cudaMalloc(&ptr, n));
}
void DeviceBuffer::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}
void DeviceBuffer::upload(const void *src, size_t n) {
    if (n > bytes) throw Error("DeviceBuffer::upload: out of range");
    if (n) BTRC_CUDA(cudaMemcpy(ptr, src, n, cudaMemcpyHostToDevice));
}
void DeviceBuffer::zero() {
    if (bytes) BTRC_CUDA(cudaMemset(ptr, 0, bytes));
}

template <typename T>
const T *Scene::upload(const std::vector<T> &v) {
    buffers_.emplace_back(std::max<size_t>(v.size() * sizeof(T), 16));
    buffers_.back().upload(v.data(), v.size() * sizeof(T));
    return buffers_.back().as<T>();
}

// Runs f and prefixes any error with a context line (reference: BTRC_HI_TRY / BTRC_HI_WRAP)
template <typename F>
static auto with_context(const std::string &ctx, F &&f) -> decltype(f()) {
    try {
        return f();
    } catch (const std::exception &e) {
        throw Error(ctx + "\n" + e.what());
    }
}

std::string Scene::resolve_path(const std::string &p) const {
    namespace fs = std::filesystem;
    std::string s = p;
    const std::string key = "$env{scene_directory}";
    for (size_t pos; (pos = s.find(key)) != std::string::npos;) s.replace(pos, key.size(), scene_dir);
    return fs::absolute(fs::path(s)).lexically_normal().string();
}

std::unique_ptr<Scene> Scene::from_file(const std::string &filename) {
    namespace fs = std::filesystem;
    auto s = std::make_unique<Scene>();
    s->scene_dir = fs::path(filename).parent_path().string();
    if (s->scene_dir.empty()) s->scene_dir = ".";
    NodeP root = parse_scene_file(filename);
    s->create(root);
    return s;
}

std::unique_ptr<Scene> Scene::from_string(const std::string &json, const std::string &scene_dir) {
    auto s = std::make_unique<Scene>();
    s->scene_dir = scene_dir.empty() ? "." : scene_dir;
    NodeP root = parse_scene_text(json, s->scene_dir);
    s->create(root);
    return s;
}

int Scene::create_texture2d(const NodeP &node) {
    return with_context("in creating texture2d", [&]() -> int {
        auto it = tex_pool_.find(node.get());
        if (it != tex_pool_.end()) return it->second;
        TextureDesc t{};
        if (node->is_value() || node->is_array()) {  // context.h:211-220: literal => Constant2D (not pooled)
            const Vec3 v = node->as_vec3();
            t.type = TEX_CONSTANT;
            t.value[0] = v.x; t.value[1] = v.y; t.value[2] = v.z;
            textures.push_back(t);
            return (int)textures.size() - 1;
        }
        const std::string type = node->child("type")->as_string();
        if (type == "constant") {
            const Vec3 v = node->child("value")->as_vec3();
            t.type = TEX_CONSTANT;
            t.value[0] = v.x; t.value[1] = v.y; t.value[2] = v.z;
        } else
            throw Error("unknown texture2d type: " + type);
        textures.push_back(t);
        tex_pool_[node.get()] = (int)textures.size() - 1;
        return (int)textures.size() - 1;
    });
}

int Scene::create_material(const NodeP &node) {
    return with_context("in creating material", [&]() -> int {
        auto it = mat_pool_.find(node.get());
        if (it != mat_pool_.end()) return it->second;
        const std::string type = node->child("type")->as_string();
        MaterialDesc m{};
        m.color_tex = m.normal_tex = -1;
        m.shadow_terminator = 1;
        if (type == "diffuse") {  // material/diffuse.cpp:36-48
            m.type = MAT_DIFFUSE;
            m.shadow_terminator = node->bool_or("shadow_terminator_term", true) ? 1 : 0;
            m.color_tex = create_texture2d(node->child("color"));
            if (auto n = node->find("normal")) m.normal_tex = create_texture2d(n);
        } else if (type == "black") {
            m.type = MAT_BLACK;
        } else
            throw Error("unknown material type: " + type);
        materials.push_back(m);
        mat_pool_[node.get()] = (int)materials.size() - 1;
        return (int)materials.size() - 1;
    });
}

int Scene::create_geometry(const NodeP &node) {
    return with_context("in creating geometry", [&]() -> int {
        auto it = geo_pool_.find(node.get());
        if (it != geo_pool_.end()) return it->second;
        const std::string type = node->child("type")->as_string();
        if (type != "triangle_mesh") throw Error("unknown geometry type: " + type);
        const std::string filename = resolve_path(node->child("filename")->as_string());  // triangle_mesh.cpp:245-254
        const bool unit = node->bool_or("transform_to_unit_cube", false);
        meshes.push_back(load_triangle_mesh(filename, unit));
        geo_pool_[node.get()] = (int)meshes.size() - 1;
        return (int)meshes.size() - 1;
    });
}

void Scene::create(const NodeP &root) {
    // cli/main.cpp:45-66
    with_context("in creating scene", [&]() {
        NodeP scene_node = root->child("scene");
        NodeP ents = scene_node->child("entities");
        if (!ents->is_array()) throw Error("entity array expected");
        for (auto &ent : ents->elements) {  // factory/scene.cpp:12-55
            if (!ent->is_group()) throw Error("entity group expected");
            HostInstance inst;
            inst.geometry_id = create_geometry(ent->child("geometry"));
            inst.material_id = create_material(ent->child("material"));
            if (auto t = ent->find("local_to_world")) inst.transform = t->as_transform3d();
            if (ent->find("inner_medium") || ent->find("outer_medium")) throw Error("unknown medium type: (media are not supported yet)");
            if (auto ln = ent->find("light")) {
                with_context("in creating light", [&]() {
                    const std::string type = ln->child("type")->as_string();
                    if (type == "mesh_light") {
                        inst.has_light = true;
                        inst.intensity = ln->child("intensity")->as_vec3();
                    } else if (type == "ibl" || type == "gradient_sky")
                        throw Error("area light expected");
                    else
                        throw Error("unknown light type: " + type);
                    return 0;
                });
            }
            if (ent->bool_or("caustics_projector", false)) inst.flag |= 1u;
            if (ent->bool_or("caustics_receiver", false)) inst.flag |= 2u;
            instances.push_back(inst);
        }
        if (scene_node->find("volumes")) throw Error("'volumes' are not supported yet");
        if (scene_node->find("envir_light")) throw Error("unknown light type: (environment lights are not supported yet)");
        return 0;
    });
    width = root->child("width")->as_int();
    height = root->child("height")->as_int();
    with_context("in creating camera", [&]() {
        NodeP cam = root->child("camera");
        const std::string type = cam->child("type")->as_string();
        if (type != "pinhole") throw Error("unknown camera type: " + type);
        cam_eye = cam->child("eye")->as_vec3();  // camera/pinhole.cpp:101-114
        cam_dst = cam->child("dst")->as_vec3();
        cam_up = cam->child("up")->as_vec3();
        cam_fov_y_deg = cam->child("fov_y")->as_degree();
        return 0;
    });
    with_context("in creating renderer", [&]() {
        NodeP r = root->child("renderer");
        renderer.type = r->child("type")->as_string();
        if (renderer.type != "wfpt" && renderer.type != "pt") throw Error("unknown renderer type: " + renderer.type);
        renderer.tile = r->bool_or("tile", renderer.tile);
        renderer.spp = r->int_or("spp", renderer.spp);
        renderer.min_depth = r->int_or("min_depth", renderer.min_depth);
        renderer.max_depth = r->int_or("max_depth", renderer.max_depth);
        renderer.rr_threshold = r->float_or("rr_threshold", renderer.rr_threshold);
        renderer.rr_cont_prob = r->float_or("rr_cont_prob", renderer.rr_cont_prob);
        renderer.state_count = r->int_or("state_count", renderer.state_count);
        renderer.albedo = r->bool_or("albedo", renderer.albedo);
        renderer.normal = r->bool_or("normal", renderer.normal);
        if (auto f = r->find("filter")) {
            const std::string ft = f->child("type")->as_string();
            if (ft == "box") renderer.filter_type = 0;
            else if (ft == "gaussian") {
                renderer.filter_type = 1;
                renderer.filter_radius = f->float_or("radius", 1.3f);
                renderer.filter_alpha = f->float_or("alpha", 1.6f);
            } else
                throw Error("unknown film_filter type: " + ft);
        }
        return 0;
    });
    if (auto pp = root->find("post_processors")) {  // factory/post_processor.cpp:5-25: group or array
        auto add = [&](const NodeP &n) {
            with_context("in creating post_processor", [&]() {
                PostProcessorDesc d;
                d.type = n->child("type")->as_string();
                if (d.type == "save") {  // postprocess/save_to_image.cpp:64-76
                    d.gamma = n->float_or("gamma", 1.0f);
                    const std::string c = n->string_or("color", ""), a = n->string_or("albedo", ""), nn = n->string_or("normal", "");
                    d.color = c.empty() ? "" : resolve_path(c);
                    d.albedo = a.empty() ? "" : resolve_path(a);
                    d.normal = nn.empty() ? "" : resolve_path(nn);
                } else if (d.type == "tonemap") {
                    d.exposure = n->float_or("exposure", 1.0f);
                } else
                    throw Error("unknown post_processor type: " + d.type);
                post_processors.push_back(d);
                return 0;
            });
        };
        if (pp->is_array()) for (auto &e : pp->elements) add(e);
        else add(pp);
    }
}

RenderParams Scene::render_params() const {
    RenderParams p{};
    p.width = width; p.height = height;
    p.spp = renderer.spp;
    p.min_depth = renderer.min_depth; p.max_depth = renderer.max_depth;
    p.rr_threshold = renderer.rr_threshold; p.rr_cont_prob = renderer.rr_cont_prob;
    p.tile = renderer.tile ? 1 : 0;
    p.albedo = renderer.albedo ? 1 : 0; p.normal = renderer.normal ? 1 : 0;
    p.filter_type = renderer.filter_type;
    p.filter_radius = renderer.filter_radius;
    p.filter_step = 2 * renderer.filter_radius / 256;
    return p;
}

static void to_3x4(const Mat4 &m, float *out) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = m.m[r][c];
}

static int tree_depth(const BinaryBvh &b) {
    int best = 0;
    std::vector<std::pair<int, int>> st = {{0, 1}};
    while (!st.empty()) {
        auto [ni, d] = st.back();
        st.pop_back();
        best = std::max(best, d);
        if (b.nodes[ni].left >= 0) { st.push_back({b.nodes[ni].left, d + 1}); st.push_back({b.nodes[ni].right, d + 1}); }
    }
    return best;
}

void Scene::commit() {
    if (width <= 0 || height <= 0) throw Error("invalid film size");
    if (instances.empty()) throw Error("scene has no entities");
    buffers_.clear();
    dev = SceneDevice{};
    device_ready_ = false;

    // ---- lights + instance table (core/scene.cpp:94-100, 163-207)
    lights.clear();
    instance_info.clear();
    for (size_t i = 0; i < instances.size(); ++i) {
        HostInstance &inst = instances[i];
        InstanceInfo info{};
        info.geometry_id = inst.geometry_id;
        info.material_id = inst.material_id;
        info.light_id = -1;
        if (inst.has_light) {
            info.light_id = (int)lights.size();
            LightInfo li{};
            li.instance_id = (int)i;
            li.intensity[0] = inst.intensity.x; li.intensity[1] = inst.intensity.y; li.intensity[2] = inst.intensity.z;
            // MeshLight::set_geometry (mesh_light.cpp:11-15)
            const float scale = length(inst.transform.point(Vec3(1, 0, 0)) - inst.transform.point(Vec3(0, 0, 0)));
            li.inv_scale2 = 1 / (scale * scale);
            lights.push_back(li);
        }
        inst.light_id = info.light_id;
        to_3x4(inst.transform.mat, info.mat);
        to_3x4(inst.transform.inv, info.inv);
        // no media: the only medium is the (empty) volume-primitive medium with id 0 (scene.cpp:136-161)
        info.inner_medium_id = info.outer_medium_id = 0;
        info.flag = inst.flag;
        instance_info.push_back(info);
    }

    // ---- world-space triangles (flattened instances) and bbox (scene.cpp:224-237)
    world_tris.clear();
    bbox = AABB();
    for (size_t i = 0; i < instances.size(); ++i) {
        const HostInstance &inst = instances[i];
        const MeshData &m = meshes[inst.geometry_id];
        for (int k = 0; k < m.ntri; ++k) {
            WorldTri t;
            t.p0 = inst.transform.point(m.pos[3 * k]);
            t.p1 = inst.transform.point(m.pos[3 * k + 1]);
            t.p2 = inst.transform.point(m.pos[3 * k + 2]);
            t.e1 = t.p0 - t.p1;
            t.e2 = t.p2 - t.p0;
            t.n = cross(t.e1, t.e2);
            t.inst = (uint32_t)i;
            t.prim = (uint32_t)k;
            world_tris.push_back(t);
        }
        for (int c = 0; c < 8; ++c) {  // Transform3D::apply_to_aabb
            Vec3 p(c & 4 ? m.bbox.hi.x : m.bbox.lo.x, c & 2 ? m.bbox.hi.y : m.bbox.lo.y, c & 1 ? m.bbox.hi.z : m.bbox.lo.z);
            bbox.extend(inst.transform.point(p));
        }
    }
    AABB wb = bbox;
    wb.extend(cam_eye);  // wavefront.cpp:138-139
    world_diagonal = 1.2f * length(wb.hi - wb.lo);

    // ---- BVH
    auto t0 = std::chrono::steady_clock::now();
    {
        const size_t n = world_tris.size();
        std::vector<AABB> boxes(n);
        std::vector<Vec3> centers(n);
        float coord_max = 0;
        for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
#pragma omp parallel for
        for (long long i = 0; i < (long long)n; ++i) {
            const WorldTri &t = world_tris[i];
            AABB b;
            b.extend(t.p0); b.extend(t.p1); b.extend(t.p2);
            // the kernels test the "effective" triangle (p0, p0 - e1, p0 + e2): include it, then pad so that the
            // fp32 slab test never culls a triangle the exact test accepts ("conservative boxes")
            b.extend(t.p0 - t.e1); b.extend(t.p0 + t.e2);
            const Vec3 ext = b.hi - b.lo;
            const float pad = 1e-5f * std::max(ext.x, std::max(ext.y, ext.z)) + 4e-7f * coord_max;
            b.lo = b.lo - Vec3(pad); b.hi = b.hi + Vec3(pad);
            boxes[i] = b;
            centers[i] = 0.5f * (b.lo + b.hi);
        }
        BinaryBvh bb = build_binary_bvh(boxes, centers, 4);
        bvh_depth = tree_depth(bb);
        if (bvh_depth > 60) throw Error("bvh too deep for the traversal stack: " + std::to_string(bvh_depth));
        flat_bvh = flatten_bvh2(bb, world_tris);
    }
    bvh_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // ---- camera (camera/pinhole.cpp:55-77)
    {
        const float w_over_h = static_cast<float>(width) / height;
        const Vec3 forward = normalize(cam_dst - cam_eye);
        const Vec3 ex = normalize(cross(forward, cam_up));
        const Vec3 ey = normalize(cross(forward, ex));
        const float y_len = 2 * std::tan(0.5f * cam_fov_y_deg * 3.1415926535f / 180);
        const float x_len = w_over_h * y_len;
        const Vec3 lbc = cam_eye + forward - 0.5f * (x_len * ex + y_len * ey);
        const Vec3 fx = x_len * ex, fy = y_len * ey;
        for (int a = 0; a < 3; ++a) {
            dev.camera.eye[a] = cam_eye[a];
            dev.camera.left_bottom_corner[a] = lbc[a];
            dev.camera.film_x[a] = fx[a];
            dev.camera.film_y[a] = fy[a];
        }
    }

    // ---- gaussian filter alias table (film_filter/gaussian.cpp:27-52)
    if (renderer.filter_type == 1) {
        constexpr int GRID = 256;
        const float radius = renderer.filter_radius, alpha = renderer.filter_alpha;
        const float expv = std::exp(-alpha * radius * radius);
        auto gaussian = [&](float d) { return std::max(0.0f, std::exp(-alpha * d * d) - expv); };
        std::vector<float> table(GRID * GRID, 0.0f);
        for (int y = 0; y < GRID; ++y) {
            const float yf = 2 * radius * (y + 0.5f) / GRID - radius;
            for (int x = 0; x < GRID; ++x) {
                const float xf = 2 * radius * (x + 0.5f) / GRID - radius;
                const float d2 = xf * xf + yf * yf;
                if (d2 < radius * radius) table[y * GRID + x] = gaussian(xf) * gaussian(yf);
            }
        }
        filter_alias = build_alias_table(table);
    }
}

void Scene::upload_device() {
    if (device_ready_) return;
    buffers_.clear();
    // ---- geometry tables
    std::vector<GeometryInfo> geos;
    for (auto &m : meshes) {
        GeometryInfo g{};
        g.gx_ua = reinterpret_cast<const float4 *>(upload(m.gx_ua));
        g.gy_uba = reinterpret_cast<const float4 *>(upload(m.gy_uba));
        g.gz_uca = reinterpret_cast<const float4 *>(upload(m.gz_uca));
        g.sn_va = reinterpret_cast<const float4 *>(upload(m.sn_va));
        g.sn_vba = reinterpret_cast<const float4 *>(upload(m.sn_vba));
        g.sn_vca = reinterpret_cast<const float4 *>(upload(m.sn_vca));
        g.positions = upload(m.a_ba_ca);
        g.area_alias = reinterpret_cast<const float2 *>(upload(m.area_alias));
        g.prim_count = m.ntri;
        g.total_area = m.total_area;
        geos.push_back(g);
    }
    dev.bvh.nodes = reinterpret_cast<const float4 *>(upload(flat_bvh.nodes));
    dev.bvh.tris = reinterpret_cast<const float4 *>(upload(flat_bvh.tris));
    dev.bvh.tri_ids = reinterpret_cast<const uint2 *>(upload(flat_bvh.tri_ids));
    dev.bvh.root_ref = flat_bvh.root_ref;
    dev.bvh.n_tris = (int)world_tris.size();
    if (!filter_alias.empty()) dev.filter_alias = reinterpret_cast<const float2 *>(upload(filter_alias));
    dev.instances = upload(instance_info);
    dev.geometries = upload(geos);
    dev.materials = upload(materials);
    dev.textures = upload(textures);
    dev.lights = upload(lights);
    dev.n_instances = (int)instance_info.size();
    dev.n_lights = (int)lights.size();
    dev.vol_prim_medium_id = 0;
    dev.world_diagonal = world_diagonal;
    device_ready_ = true;
}

}  // namespace btrc
