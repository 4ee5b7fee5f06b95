#include "scene.h"

#include "../device/kernels.h"
#include "image_io.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

Scene::~Scene() { release_textures(); }

void Scene::release_textures() {
    for (auto &src : sources) {
        if (src.tex) cudaDestroyTextureObject(src.tex);
        if (src.array) cudaFreeArray(src.array);
        src.tex = 0;
        src.array = nullptr;
    }
}

static int parse_address_mode(const std::string &s) {  // description.cpp:7-18
    if (s == "clamp") return 0;
    if (s == "wrap") return 1;
    if (s == "mirror") return 2;
    if (s == "border") return 3;
    throw Error("unknown address mode: " + s);
}

void Scene::parse_texture_desc(const NodeP &node, TextureSource &src) const {  // description.cpp:28-80
    with_context("in parsing texture description", [&]() {
        src.address[0] = parse_address_mode(node->string_or("address_mode_u", "clamp"));
        src.address[1] = parse_address_mode(node->string_or("address_mode_v", "clamp"));
        src.address[2] = parse_address_mode(node->string_or("address_mode_w", "clamp"));
        const std::string filter = node->string_or("filter", "linear");
        if (filter == "point") src.linear = false;
        else if (filter == "linear") src.linear = true;
        else throw Error("unknown filter mode: " + filter);
        src.srgb = node->bool_or("srgb_to_linear", false);
        if (auto n = node->find("border")) {
            if (n->is_array()) {
                if (n->elements.size() == 1) {
                    const float v = n->elements[0]->as_float();
                    for (float &b : src.border) b = v;
                } else if (n->elements.size() == 3) {
                    for (int k = 0; k < 3; ++k) src.border[k] = n->elements[k]->as_float();
                } else
                    throw Error("invalid border array size");
            } else if (n->is_value()) {
                const float v = n->as_float();
                for (float &b : src.border) b = v;
            } else
                throw Error("border: expect array or value node");
        }
        return 0;
    });
}

int Scene::constant_texture2d(float v) {
    Tex2DNode t;
    t.type = TEXOP_CONST;
    t.value = Vec3(v);
    textures.push_back(t);
    return (int)textures.size() - 1;
}

int Scene::create_texture2d(const NodeP &node) {
    return with_context("in creating texture2d", [&]() -> int {
        auto it = tex_pool_.find(node.get());
        if (it != tex_pool_.end()) return it->second;
        Tex2DNode t;
        if (node->is_value() || node->is_array()) {  // context.h:211-220: literal => Constant2D
            t.type = TEXOP_CONST;
            t.value = node->as_vec3();
        } else {
            const std::string type = node->child("type")->as_string();
            if (type == "constant") {
                t.type = TEXOP_CONST;
                t.value = node->child("value")->as_vec3();
            } else if (type == "array") {  // array2d.cpp:32-39; utils/image.cpp:299-335; cuda/array.cpp:136-149
                TextureSource src;
                const std::string filename = resolve_path(node->child("filename")->as_string());
                parse_texture_desc(node, src);
                LoadedImage img = load_image(filename);
                src.width = img.width; src.height = img.height;
                const size_t npx = (size_t)img.width * img.height;
                if (img.is_float) {  // Image<Vec3f> -> F32x4
                    src.is_float = true; src.channels = 4;
                    src.f32.resize(4 * npx);
                    for (size_t i = 0; i < npx; ++i) {
                        for (int k = 0; k < 3; ++k) src.f32[4 * i + k] = img.f32[3 * i + k];
                        src.f32[4 * i + 3] = 1.0f;
                    }
                } else if (img.channels == 1) {  // UNorm8x1
                    src.channels = 1;
                    src.u8 = std::move(img.u8);
                } else if (img.channels == 3 || img.channels == 4) {  // UNorm8x4 (alpha 255 for rgb)
                    src.channels = 4;
                    src.u8.resize(4 * npx);
                    for (size_t i = 0; i < npx; ++i) {
                        for (int k = 0; k < 3; ++k) src.u8[4 * i + k] = img.u8[img.channels * i + k];
                        src.u8[4 * i + 3] = img.channels == 4 ? img.u8[4 * i + 3] : 255;
                    }
                } else
                    throw Error("ImageDynamic::load: unsupported channels: " + std::to_string(img.channels));
                sources.push_back(std::move(src));
                t.type = TEXOP_ARRAY;
                t.source = (int)sources.size() - 1;
            } else if (type == "add" || type == "sub" || type == "mul" || type == "div") {  // math.cpp
                t.type = type == "add" ? TEXOP_ADD : type == "sub" ? TEXOP_SUB : type == "mul" ? TEXOP_MUL : TEXOP_DIV;
                t.lhs = create_texture2d(node->child("lhs"));
                t.rhs = create_texture2d(node->child("rhs"));
            } else if (type == "transform") {  // transform.cpp:43-53
                t.type = 6;
                t.m = node->child("transform")->as_transform2d();
                t.lhs = create_texture2d(node->child("tex"));
            } else
                throw Error("unknown texture2d type: " + type);
        }
        textures.push_back(t);
        tex_pool_[node.get()] = (int)textures.size() - 1;
        return (int)textures.size() - 1;
    });
}

// Text grid "<byte|float> <channels> <W> <H> <D>" + values, x fastest (utils/cuda/array.cpp:313-385)
static void load_text_grid(const std::string &filename, TextureSource &src) {
    const std::string text = read_text_file(filename);
    const char *p = text.c_str();
    char comp[16] = {0};
    int channels = 0, w = 0, h = 0, d = 0, used = 0;
    if (std::sscanf(p, "%15s %d %d %d %d%n", comp, &channels, &w, &h, &d, &used) != 5) throw Error("Array::load_from_text: failed to parse file head");
    p += used;
    if (channels != 1 && channels != 3 && channels != 4) throw Error("unsupported channel count: " + std::to_string(channels));
    const bool bytes = std::string(comp) == "byte";
    if (!bytes && std::string(comp) != "float") throw Error(std::string("unknown component type: ") + comp);
    if (w <= 0 || h <= 0 || d <= 0) throw Error("Array::load_from_text: failed to parse file head");
    const size_t n = (size_t)w * h * d;
    if (bytes) {
        // array.cpp:331-355 reads each component with `uint8_t comp; fin >> comp;` -- formatted extraction into an unsigned char
        // takes ONE non-blank CHARACTER, not a number -- and uploads the grid as SNorm8 (texel value = max(int8 / 127, -1), the
        // extrema scaled by 1 / 127: find_minmax_impl<int8_t, N, true>, :436-510).  A three-channel grid gets 255 (= -1 / 127) as
        // its fourth component.  Reproduced as is: a "byte" file means the same grid here as there.
        src.width = w; src.height = h; src.depth = d;
        src.is_float = false; src.snorm = true;
        src.channels = channels == 1 ? 1 : 4;
        src.u8.resize(n * src.channels);
        int lo[3] = {127, 127, 127}, hi[3] = {-128, -128, -128};
        for (size_t i = 0; i < n; ++i) {
            for (int c = 0; c < channels; ++c) {
                while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\v' || *p == '\f') ++p;
                if (!*p) throw Error("failed to load " + filename);
                const uint8_t v = (uint8_t)*p++;
                src.u8[i * src.channels + c] = v;
                if (c < 3) { lo[c] = std::min(lo[c], (int)(int8_t)v); hi[c] = std::max(hi[c], (int)(int8_t)v); }
            }
            if (channels == 3) src.u8[i * 4 + 3] = 255;
        }
        const float factor = static_cast<float>(1.0 / 127);
        for (int k = 0; k < 3; ++k) {
            const bool have = k < channels;  // components a texel does not have keep find_minmax_impl's start values
            src.min_value[k] = static_cast<float>(have ? lo[k] : 127) * factor;
            src.max_value[k] = static_cast<float>(have ? hi[k] : -128) * factor;
        }
        return;
    }
    if (n * (size_t)channels > text.size()) throw Error("failed to load " + filename);  // every value takes at least two characters: the file cannot hold that grid
    src.width = w; src.height = h; src.depth = d;
    src.is_float = true;
    src.channels = channels == 1 ? 1 : 4;
    src.f32.resize(n * src.channels);
    for (int k = 0; k < 3; ++k) { src.max_value[k] = -3.402823466e+38f; src.min_value[k] = 3.402823466e+38f; }
    for (size_t i = 0; i < n; ++i) {
        for (int c = 0; c < channels; ++c) {
            char *end = nullptr;
            const float v = std::strtof(p, &end);
            if (end == p) throw Error("failed to load " + filename);
            p = end;
            src.f32[i * src.channels + c] = v;
            if (c < 3) { src.max_value[c] = std::max(src.max_value[c], v); src.min_value[c] = std::min(src.min_value[c], v); }
        }
        if (channels == 3) src.f32[i * 4 + 3] = 1.0f;
    }
}

// Texture3D::get_min_float / get_max_float: constant, array (texel extrema of the grid), add / mul by interval arithmetic on the
// operands' extrema (texture3d/binary.cpp:95-137)
float Scene::texture3d_extremum(int tex, bool want_max) const {
    const Tex3DNode &t = textures3d[tex];
    if (t.op == 0) {
        if (!t.is_array) return t.value.x;
        return want_max ? sources[t.source].max_value[0] : sources[t.source].min_value[0];
    }
    const float lmax = texture3d_extremum(t.lhs, true), lmin = texture3d_extremum(t.lhs, false);
    const float rmax = texture3d_extremum(t.rhs, true), rmin = texture3d_extremum(t.rhs, false);
    if (t.op == TEX3D_ADD) return want_max ? lmax + rmax : lmin + rmin;
    if (want_max) return std::max(std::max(lmax * rmax, lmax * rmin), std::max(lmin * rmax, lmin * rmin));
    return std::min(std::min(lmax * rmax, lmax * rmin), std::min(lmin * rmax, lmin * rmin));
}

int Scene::create_texture3d(const NodeP &node) {
    return with_context("in creating texture3d", [&]() -> int {
        auto it = tex3d_pool_.find(node.get());
        if (it != tex3d_pool_.end()) return it->second;
        Tex3DNode t;
        if (node->is_value() || node->is_array()) {
            t.value = node->as_vec3();
        } else {
            const std::string type = node->child("type")->as_string();
            if (type == "constant") {
                t.value = node->child("value")->as_vec3();
            } else if (type == "array") {  // texture3d/array3d.cpp:53-62
                TextureSource src;
                parse_texture_desc(node, src);
                load_text_grid(resolve_path(node->child("text")->as_string()), src);
                sources.push_back(std::move(src));
                t.is_array = true;
                t.source = (int)sources.size() - 1;
            } else if (type == "add" || type == "mul") {  // texture3d/binary.cpp:150-160 (lhs is created first)
                t.op = type == "add" ? TEX3D_ADD : TEX3D_MUL;
                t.lhs = create_texture3d(node->child("lhs"));
                t.rhs = create_texture3d(node->child("rhs"));
            } else
                throw Error("unknown texture3d type: " + type);
        }
        textures3d.push_back(t);
        tex3d_pool_[node.get()] = (int)textures3d.size() - 1;
        return (int)textures3d.size() - 1;
    });
}

int Scene::create_material(const NodeP &node) {
    return with_context("in creating material", [&]() -> int {
        auto it = mat_pool_.find(node.get());
        if (it != mat_pool_.end()) return it->second;
        const std::string type = node->child("type")->as_string();
        MaterialDesc m{};
        for (int &t : m.tex) t = -1;
        m.normal_tex = -1;
        m.shadow_terminator = 1;
        auto req = [&](int slot, const char *name) { m.tex[slot] = create_texture2d(node->child(name)); };
        auto opt = [&](int slot, const char *name, float d) {
            if (auto n = node->find(name)) m.tex[slot] = create_texture2d(n);
            else m.tex[slot] = constant_texture2d(d);
        };
        bool has_normal = true, has_st = true;
        if (type == "diffuse") {  // material/diffuse.cpp:36-48
            m.type = MAT_DIFFUSE;
            req(SLOT_COLOR, "color");
        } else if (type == "disney") {  // material/disney.cpp:844-894
            m.type = MAT_DISNEY;
            req(SLOT_COLOR, "color"); req(SLOT_METALLIC, "metallic"); req(SLOT_ROUGHNESS, "roughness");
            opt(SLOT_TRANSMISSION, "transmission", 0); opt(SLOT_IOR, "ior", 1.5f); opt(SLOT_SPECULAR, "specular", 1);
            opt(SLOT_SPECULAR_TINT, "specular_tint", 0); opt(SLOT_ANISOTROPIC, "anisotropic", 0); opt(SLOT_SHEEN, "sheen", 0);
            opt(SLOT_SHEEN_TINT, "sheen_tint", 0); opt(SLOT_CLEARCOAT, "clearcoat", 0); opt(SLOT_CLEARCOAT_GLOSS, "clearcoat_gloss", 1);
            if (auto n = node->find("transmission_roughness")) m.tex[SLOT_TRANSMISSION_ROUGHNESS] = create_texture2d(n);
            else m.tex[SLOT_TRANSMISSION_ROUGHNESS] = m.tex[SLOT_ROUGHNESS];
        } else if (type == "metal") {  // material/metal.cpp:50-77
            m.type = MAT_METAL;
            req(SLOT_COLOR, "color"); req(SLOT_ROUGHNESS, "roughness"); opt(SLOT_ANISOTROPIC, "anisotropic", 0);
        } else if (type == "glass") {  // material/glass.cpp:141-157
            m.type = MAT_GLASS; has_st = false;
            req(SLOT_COLOR, "color"); req(SLOT_IOR, "ior");
        } else if (type == "mirror") {  // material/mirror.cpp:91-103
            m.type = MAT_MIRROR; has_st = false;
            req(SLOT_COLOR, "color");
        } else if (type == "black") {
            m.type = MAT_BLACK; has_normal = has_st = false;
        } else if (type == "invisible") {
            m.type = MAT_INVISIBLE; has_normal = has_st = false;
        } else
            throw Error("unknown material type: " + type);
        if (has_st) m.shadow_terminator = node->bool_or("shadow_terminator_term", true) ? 1 : 0;
        if (has_normal) if (auto n = node->find("normal")) m.normal_tex = create_texture2d(n);  // utils/normal_map.cpp:5-11
        materials.push_back(m);
        mat_pool_[node.get()] = (int)materials.size() - 1;
        return (int)materials.size() - 1;
    });
}

static void mat4_to_3x4(const Mat4 &m, float *out) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = m.m[r][c];
}

// Returns a handle into `media`.  "transform" wrappers (factory/context.h:146-165, core/medium.cpp:98-129) are folded
// into the wrapped medium's world_to_local: inner(outer(p)).
int Scene::create_medium(const NodeP &node, const Mat4 *outer) {
    return with_context("in creating medium", [&]() -> int {
        if (!outer) {
            auto it = med_pool_.find(node.get());
            if (it != med_pool_.end()) return it->second;
        }
        const std::string type = node->child("type")->as_string();
        int handle = -1;
        if (type == "transform") {
            Mat4 w2l;
            if (auto t = node->find("world_to_local")) w2l = t->as_transform3d().mat;
            else w2l = node->child("local_to_world")->as_transform3d().inv;
            if (outer) w2l = w2l * *outer;
            handle = create_medium(node->child("medium"), &w2l);
        } else if (type == "homogeneous") {  // medium/homogeneous.cpp:80-92
            HostMedium m;
            m.priority = node->float_or("priority", 0.0f);
            m.desc.type = MED_HOMOGENEOUS;
            m.desc.sigma_t = node->child("sigma_t")->as_float();
            const Vec3 a = node->child("albedo")->as_vec3();
            m.desc.albedo[0] = a.x; m.desc.albedo[1] = a.y; m.desc.albedo[2] = a.z;
            m.desc.g = node->float_or("g", 0.0f);
            m.desc.sigma_t_tex = m.desc.albedo_tex = m.desc.g_tex = -1;
            media.push_back(m);
            handle = (int)media.size() - 1;
        } else if (type == "hetergeneous") {  // medium/hetergeneous.cpp:113-135
            HostMedium m;
            m.priority = node->float_or("priority", 0.0f);
            m.desc.type = MED_HETEROGENEOUS;
            m.desc.sigma_t_tex = create_texture3d(node->child("sigma_t"));
            m.desc.albedo_tex = create_texture3d(node->child("albedo"));
            if (auto g = node->find("g")) m.desc.g_tex = create_texture3d(g);
            else {
                textures3d.push_back(Tex3DNode());  // Constant3D(0)
                m.desc.g_tex = (int)textures3d.size() - 1;
            }
            if (outer) { m.desc.has_transform = 1; mat4_to_3x4(*outer, m.desc.world_to_local); }
            const float mx = texture3d_extremum(m.desc.sigma_t_tex, true);
            m.desc.inv_max_density = 1.0f / std::max(mx, 0.001f);  // hetergeneous.cpp:33-34
            media.push_back(m);
            handle = (int)media.size() - 1;
        } else
            throw Error("unknown medium type: " + type);
        if (!outer) med_pool_[node.get()] = handle;
        return handle;
    });
}

void Scene::create_envir_light(const NodeP &node) {
    with_context("in creating light", [&]() {
        const std::string type = node->child("type")->as_string();
        if (type == "gradient_sky") {  // light/gradient_sky.cpp:56-67
            env.type = ENV_GRADIENT_SKY;
            env.upper = node->child("upper")->as_vec3();
            env.lower = node->child("lower")->as_vec3();
            env.up = normalize(node->child("up")->as_vec3());
        } else if (type == "ibl") {  // light/ibl.cpp:89-102
            env.type = ENV_IBL;
            env.tex = create_texture2d(node->child("tex"));
            env.up = node->find("up") ? node->child("up")->as_vec3() : Vec3(0, 0, 1);
            env.lut_w = node->int_or("lut_width", 400);
            env.lut_h = node->int_or("lut_height", 200);
        } else if (type == "mesh_light")
            throw Error("'envir_light' is not an environment light");
        else
            throw Error("unknown light type: " + type);
        return 0;
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
            if (auto mn = ent->find("inner_medium")) inst.inner_medium = create_medium(mn, nullptr);
            if (auto mn = ent->find("outer_medium")) inst.outer_medium = create_medium(mn, nullptr);
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
        if (auto v = scene_node->find("volumes")) {
            if (!v->is_array()) throw Error("'volumes' is not an array node");
            for (const NodeP &vol : v->elements) {  // factory/scene.cpp:64-79
                VolumePrim vp{};
                const Vec3 o = vol->child("o")->as_vec3(), x = vol->child("x")->as_vec3(), y = vol->child("y")->as_vec3(), z = vol->child("z")->as_vec3();
                const float x2 = dot(x, x), y2 = dot(y, y), z2 = dot(z, z);
                for (int a = 0; a < 3; ++a) {
                    vp.o[a] = o[a]; vp.x[a] = x[a]; vp.y[a] = y[a]; vp.z[a] = z[a];
                    vp.x_div_x2[a] = x[a] / x2; vp.y_div_y2[a] = y[a] / y2; vp.z_div_z2[a] = z[a] / z2;  // volume/bvh.cpp:96-102
                }
                vp.inv_x2 = 1.0f / x2; vp.inv_y2 = 1.0f / y2; vp.inv_z2 = 1.0f / z2;
                vp.sigma_t_tex = create_texture3d(vol->child("sigma_t"));
                vp.albedo_tex = create_texture3d(vol->child("albedo"));
                vp.max_sigma_t = texture3d_extremum(vp.sigma_t_tex, true);
                volume_prims.push_back(vp);
            }
        }
        if (auto e = scene_node->find("envir_light")) create_envir_light(e);
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
        if (auto bn = r->find("bvh_builder")) {  // the reference always builds with OptiX; here: host SAH or device LBVH
            const std::string bt = bn->as_string();
            if (bt != "sah" && bt != "lbvh") throw Error("unknown bvh builder: " + bt);
            renderer.lbvh = bt == "lbvh";
        }
        if (auto ln = r->find("bvh_layout")) {
            const std::string lt = ln->as_string();
            if (lt != "auto" && lt != "flat" && lt != "instanced" && lt != "wide8") throw Error("unknown bvh layout: " + lt);
            renderer.bvh_layout = lt == "flat" ? 1 : lt == "instanced" ? 2 : lt == "wide8" ? 3 : 0;
        }
        if (auto sn = r->find("sampler")) {  // wavefront/common.h:36-37 chooses at compile time; here it is a renderer parameter
            const std::string st = sn->as_string();
            if (st != "independent" && st != "halton") throw Error("unknown sampler type: " + st);
            renderer.halton = st == "halton";
        }
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
                } else if (d.type == "denoise") {
                    // postprocess/optix_denoiser.cpp: NVIDIA's closed-source OptiX AI denoiser.  There is no OptiX in this renderer
                    // (out of scope); say so instead of "unknown", the name is a registered creator in the reference.
                    throw Error("post_processor 'denoise' is the OptiX AI denoiser, which this OptiX-free renderer does not provide; "
                                "remove it from \"post_processors\" (albedo / normal outputs for an external denoiser: \"save\")");
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
    p.halton = renderer.halton ? 1 : 0;
    p.albedo = renderer.albedo ? 1 : 0; p.normal = renderer.normal ? 1 : 0;
    p.filter_type = renderer.filter_type;
    p.filter_radius = renderer.filter_radius;
    p.filter_step = 2 * renderer.filter_radius / 256;
    return p;
}

static void to_3x4(const Mat4 &m, float *out) { mat4_to_3x4(m, out); }

void Scene::commit() {
    if (width <= 0 || height <= 0) throw Error("invalid film size");
    if (instances.empty()) throw Error("scene has no entities");
    buffers_.clear();
    dev = SceneDevice{};
    device_ready_ = false;

    // ---- medium ids (core/scene.cpp:136-161): real media by descending priority (ties: creation order), then the
    // volume-primitive medium.  A null medium means the volume-primitive medium.
    std::vector<int> order(media.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return media[a].priority > media[b].priority; });
    std::vector<int> medium_id_of(media.size());
    medium_table.clear();
    for (size_t i = 0; i < order.size(); ++i) {
        medium_id_of[order[i]] = (int)i;
        medium_table.push_back(media[order[i]].desc);
    }
    const uint32_t vol_prim_id = (uint32_t)medium_table.size();
    {
        MediumDesc vp{};
        vp.type = MED_VOLPRIM;
        medium_table.push_back(vp);
    }
    has_medium = !volume_prims.empty();  // core/scene.cpp:346-349

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
        info.inner_medium_id = inst.inner_medium >= 0 ? (uint32_t)medium_id_of[inst.inner_medium] : vol_prim_id;
        info.outer_medium_id = inst.outer_medium >= 0 ? (uint32_t)medium_id_of[inst.outer_medium] : vol_prim_id;
        if (inst.inner_medium >= 0 || inst.outer_medium >= 0) has_medium = true;
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
    for (const VolumePrim &v : volume_prims) {  // core/scene.cpp:233-237; VolumePrimitive::get_bounding_box (core/volume.cpp:51-66)
        const Vec3 o(v.o[0], v.o[1], v.o[2]);
        bbox.extend(o);
        bbox.extend(o + Vec3(v.x[0], v.x[1], v.x[2]));
        bbox.extend(o + Vec3(v.y[0], v.y[1], v.y[2]));
        bbox.extend(o + Vec3(v.z[0], v.z[1], v.z[2]));
    }
    AABB wb = bbox;
    wb.extend(cam_eye);  // wavefront.cpp:138-139
    world_diagonal = 1.2f * length(wb.hi - wb.lo);

    // ---- BVH
    auto t0 = std::chrono::steady_clock::now();
    scene_coord_max = 0;
    for (int a = 0; a < 3; ++a) scene_coord_max = std::max(scene_coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    scene_box = wb;
    // layout: the two-level tree (host/bvh.h) is built on request only ("bvh_layout": "instanced" / BTRC_BVH_LAYOUT).  Measured on B200
    // (profiles/r2c_ab_bvh_layout.md): it keeps a mesh's nodes once (C4: 0.6 GB of nodes -> 4.5 MB) and passes every parity test bit
    // for bit, but traces SLOWER than the flat tree -- C4 129 vs 144 Msamples/s, C2 991 vs 1068: the 250 instances of C4 overlap in
    // space, so a ray enters many mesh trees from their roots, where the flat SAH tree separates the triangles of all instances at once.
    instanced_bvh = false;
    two_level = TwoLevelBvh();
    {
        int layout = renderer.bvh_layout;
        if (const char *e = std::getenv("BTRC_BVH_LAYOUT")) layout = std::string(e) == "flat" ? 1 : std::string(e) == "instanced" ? 2 : std::string(e) == "wide8" ? 3 : 0;
        instanced_bvh = !renderer.lbvh && layout == 2;
        wide_bvh8 = !renderer.lbvh && layout == 3;
    }
    wide_bvh = WideBvh8();
    if (renderer.lbvh) {  // built on the device at upload (device/lbvh.cu): only the unsorted triangle arrays are prepared here
        flat_bvh = FlatBvh2();
        bvh_depth = 0;
    } else if (instanced_bvh) {
        flat_bvh = FlatBvh2();
        std::vector<std::vector<WorldTri>> mesh_tris(meshes.size());
        for (size_t m = 0; m < meshes.size(); ++m) {
            const MeshData &md = meshes[m];
            mesh_tris[m].resize((size_t)md.ntri);
            for (int k = 0; k < md.ntri; ++k) {
                WorldTri &t = mesh_tris[m][k];
                t.p0 = md.pos[3 * k]; t.p1 = md.pos[3 * k + 1]; t.p2 = md.pos[3 * k + 2];
                t.e1 = t.p0 - t.p1; t.e2 = t.p2 - t.p0; t.n = cross(t.e1, t.e2);
                t.inst = 0; t.prim = (uint32_t)k;
            }
        }
        std::vector<InstanceRef> refs(instances.size());
        size_t first = 0;
        for (size_t i = 0; i < instances.size(); ++i) {
            refs[i].mesh = instances[i].geometry_id;
            to_3x4(instances[i].transform.inv, refs[i].inv);
            refs[i].first_tri = first;
            first += (size_t)meshes[instances[i].geometry_id].ntri;
        }
        float coord_max = 0;
        for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
        two_level = build_two_level_bvh(mesh_tris, refs, world_tris, coord_max);
        bvh_depth = two_level.depth;
        if (bvh_depth > 60) throw Error("bvh too deep for the traversal stack: " + std::to_string(bvh_depth));
    } else {
        std::vector<AABB> boxes;
        std::vector<Vec3> centers;
        float coord_max = 0;
        for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
        padded_triangle_boxes(world_tris, coord_max, boxes, centers);  // DESIGN.md, "conservative boxes"
        if (wide_bvh8) {  // "bvh_layout": "wide8" (bvh.h): leaves of at most 3 triangles, collapsed 8-wide, 8-bit child boxes.  Measured on B200
            // (profiles/r2v_ab_wide_bvh8.md): a fifth of the node bytes, but a step costs ~5x the instructions of a binary step for ~0.6x
            // the steps -- slower than the binary tree on every BASELINE scene, hence on request only.
            flat_bvh = FlatBvh2();
            const BinaryBvh bb = build_binary_bvh(boxes, centers, 3, true);
            wide_bvh = build_wide_bvh8(bb, world_tris);
            bvh_depth = wide_bvh.depth;
            if (bvh_depth > 30) throw Error("wide bvh too deep for the traversal stack: " + std::to_string(bvh_depth));
        } else {
            BinaryBvh bb = build_binary_bvh(boxes, centers, 4);
            bvh_depth = tree_depth(bb);
            if (bvh_depth > 60) throw Error("bvh too deep for the traversal stack: " + std::to_string(bvh_depth));
            flat_bvh = flatten_bvh2(bb, world_tris);
        }
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

// Postfix program of a texture graph node; `m` = uv transform accumulated from the "transform" wrappers above it.
void Scene::emit_tex_ops(int node, const Mat3 &m, std::vector<TexOp> &ops) const {
    const Tex2DNode &t = textures[node];
    if (t.type == 6) {  // uv' = t.m.apply_to_point(m.apply_to_point(uv)): affine composition of rows 0 and 1
        Mat3 c = Mat3::identity();
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 3; ++k) c.m[r][k] = t.m.m[r][0] * m.m[0][k] + t.m.m[r][1] * m.m[1][k] + (k == 2 ? t.m.m[r][2] : 0.0f);
        emit_tex_ops(t.lhs, c, ops);
        return;
    }
    if (t.type == TEXOP_ADD || t.type == TEXOP_SUB || t.type == TEXOP_MUL || t.type == TEXOP_DIV) {
        emit_tex_ops(t.lhs, m, ops);
        emit_tex_ops(t.rhs, m, ops);
    }
    TexOp op{};
    op.op = t.type;
    op.value[0] = t.value.x; op.value[1] = t.value.y; op.value[2] = t.value.z;
    if (t.type == TEXOP_ARRAY) {
        op.tex = (unsigned long long)sources[t.source].tex;
        for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) op.m[3 * r + c] = m.m[r][c];
    }
    ops.push_back(op);
}

static void create_texture_object(TextureSource &src) {  // utils/cuda/array.cpp:104-172, texture.cpp:63-105
    if (src.tex) return;
    cudaChannelFormatDesc fmt;
    size_t texel;
    if (src.is_float) {
        fmt = src.channels == 1 ? cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat) : cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
        texel = 4 * (size_t)src.channels;
    } else {
        const cudaChannelFormatKind kind = src.snorm ? cudaChannelFormatKindSigned : cudaChannelFormatKindUnsigned;
        fmt = src.channels == 1 ? cudaCreateChannelDesc(8, 0, 0, 0, kind) : cudaCreateChannelDesc(8, 8, 8, 8, kind);
        texel = (size_t)src.channels;
    }
    const void *data = src.is_float ? (const void *)src.f32.data() : (const void *)src.u8.data();
    if (src.depth > 0) {
        const cudaExtent ext = make_cudaExtent((size_t)src.width, (size_t)src.height, (size_t)src.depth);
        BTRC_CUDA(cudaMalloc3DArray(&src.array, &fmt, ext));
        cudaMemcpy3DParms p{};
        p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(data), src.width * texel, (size_t)src.width, (size_t)src.height);
        p.dstArray = src.array;
        p.extent = ext;
        p.kind = cudaMemcpyHostToDevice;
        BTRC_CUDA(cudaMemcpy3D(&p));
    } else {
        BTRC_CUDA(cudaMallocArray(&src.array, &fmt, (size_t)src.width, (size_t)src.height));
        BTRC_CUDA(cudaMemcpy2DToArray(src.array, 0, 0, data, src.width * texel, src.width * texel, (size_t)src.height, cudaMemcpyHostToDevice));
    }
    static const cudaTextureAddressMode modes[4] = {cudaAddressModeClamp, cudaAddressModeWrap, cudaAddressModeMirror, cudaAddressModeBorder};
    cudaTextureDesc td{};
    for (int k = 0; k < 3; ++k) td.addressMode[k] = modes[src.address[k]];
    td.filterMode = src.linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.normalizedCoords = 1;
    for (int k = 0; k < 4; ++k) td.borderColor[k] = src.border[k];
    td.sRGB = src.srgb ? 1 : 0;
    td.readMode = src.is_float ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = src.array;
    BTRC_CUDA(cudaCreateTextureObject(&src.tex, &rd, &td, nullptr));
}

void Scene::upload_device() {
    if (device_ready_) return;
    buffers_.clear();
    release_textures();
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
    if (renderer.lbvh) {
        const size_t n = world_tris.size();
        std::vector<float> tin(12 * n);
        std::vector<uint32_t> iin(2 * n);
#pragma omp parallel for
        for (long long i = 0; i < (long long)n; ++i) {
            const WorldTri &t = world_tris[i];
            float *o = &tin[12 * (size_t)i];
            o[0] = t.p0.x; o[1] = t.p0.y; o[2] = t.p0.z; o[3] = t.n.x;
            o[4] = t.e1.x; o[5] = t.e1.y; o[6] = t.e1.z; o[7] = t.n.y;
            o[8] = t.e2.x; o[9] = t.e2.y; o[10] = t.e2.z; o[11] = t.n.z;
            iin[2 * (size_t)i] = t.inst; iin[2 * (size_t)i + 1] = t.prim;
        }
        DeviceBuffer d_tin(tin.size() * 4), d_iin(iin.size() * 4);
        d_tin.upload(tin.data(), tin.size() * 4);
        d_iin.upload(iin.data(), iin.size() * 4);
        buffers_.emplace_back(tin.size() * 4);
        float4 *d_tris = buffers_.back().as<float4>();
        buffers_.emplace_back(iin.size() * 4);
        uint2 *d_ids = buffers_.back().as<uint2>();
        float4 *d_nodes = nullptr;
        int n_nodes = 0;
        const float lo[3] = {scene_box.lo.x, scene_box.lo.y, scene_box.lo.z}, hi[3] = {scene_box.hi.x, scene_box.hi.y, scene_box.hi.z};
        const auto t0 = std::chrono::steady_clock::now();
        try {
            dev.bvh.root_ref = build_lbvh_device((int)n, d_tin.as<float4>(), d_iin.as<uint2>(), lo, hi, scene_coord_max, d_tris, d_ids, &d_nodes,
                                                 &n_nodes, nullptr);
        } catch (const std::exception &e) { throw Error(e.what()); }
        bvh_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        buffers_.emplace_back();
        buffers_.back().ptr = d_nodes;  // adopted: freed with the scene's other device buffers
        buffers_.back().bytes = 64 * (size_t)std::max(n_nodes, 1);
        if (!d_nodes) buffers_.back().alloc(64);
        dev.bvh.nodes = buffers_.back().as<float4>();
        dev.bvh.tris = d_tris;
        dev.bvh.tri_ids = d_ids;
        dev.bvh.n_tris = (int)n;
        lbvh_nodes = n_nodes;
    } else if (instanced_bvh) {
        dev.bvh.nodes = reinterpret_cast<const float4 *>(upload(two_level.bl_nodes));
        dev.bvh.tl_nodes = reinterpret_cast<const float4 *>(upload(two_level.tl_nodes));
        static_assert(sizeof(InstanceBlas) == 64, "InstanceBlas is read as 4 float4");
        dev.bvh.inst_blas = reinterpret_cast<const float4 *>(upload(two_level.inst));
        dev.bvh.tris = reinterpret_cast<const float4 *>(upload(two_level.tris));
        dev.bvh.tri_ids = reinterpret_cast<const uint2 *>(upload(two_level.tri_ids));
        dev.bvh.root_ref = two_level.tl_root_ref;
        dev.bvh.two_level = 1;
        dev.bvh.n_tris = (int)world_tris.size();
    } else if (wide_bvh8) {
        dev.bvh.wide_nodes = reinterpret_cast<const uint4 *>(upload(wide_bvh.nodes));
        dev.bvh.tris = reinterpret_cast<const float4 *>(upload(wide_bvh.tris));
        dev.bvh.tri_ids = reinterpret_cast<const uint2 *>(upload(wide_bvh.tri_ids));
        dev.bvh.root_ref = 0;
        dev.bvh.wide = 1;
        dev.bvh.n_tris = (int)world_tris.size();
    } else {
        dev.bvh.nodes = reinterpret_cast<const float4 *>(upload(flat_bvh.nodes));
        dev.bvh.tris = reinterpret_cast<const float4 *>(upload(flat_bvh.tris));
        dev.bvh.tri_ids = reinterpret_cast<const uint2 *>(upload(flat_bvh.tri_ids));
        dev.bvh.root_ref = flat_bvh.root_ref;
        dev.bvh.n_tris = (int)world_tris.size();
    }
    if (!filter_alias.empty()) dev.filter_alias = reinterpret_cast<const float2 *>(upload(filter_alias));
    for (InstanceInfo &info : instance_info) {  // table pointers of the instance's mesh, denormalised (types.h)
        const GeometryInfo &g = geos[info.geometry_id];
        info.gx_ua = g.gx_ua; info.gy_uba = g.gy_uba; info.gz_uca = g.gz_uca;
        info.sn_va = g.sn_va; info.sn_vba = g.sn_vba; info.sn_vca = g.sn_vca;
    }
    for (LightInfo &li : lights) {
        const InstanceInfo &info = instance_info[li.instance_id];
        const GeometryInfo &g = geos[info.geometry_id];
        li.prim_count = g.prim_count; li.total_area = g.total_area;
        li.area_alias = g.area_alias; li.positions = g.positions; li.gz_uca = g.gz_uca;
        std::memcpy(li.mat, info.mat, sizeof(li.mat));
        std::memcpy(li.inv, info.inv, sizeof(li.inv));
    }
    dev.instances = upload(instance_info);
    {
        std::vector<uint8_t> cls(instance_info.size());
        for (size_t i = 0; i < cls.size(); ++i) cls[i] = (uint8_t)(1 + materials[instance_info[i].material_id].type);
        dev.instance_class = upload(cls);
    }
    dev.geometries = upload(geos);
    // ---- textures: CUDA texture objects for the "array" sources, postfix programs for the graph
    for (auto &src : sources) create_texture_object(src);
    std::vector<TextureDesc> tex_descs;
    std::vector<TexOp> tex_ops;
    for (size_t i = 0; i < textures.size(); ++i) {
        TextureDesc d{};
        d.first_op = (int)tex_ops.size();
        emit_tex_ops((int)i, Mat3::identity(), tex_ops);
        d.n_ops = (int)tex_ops.size() - d.first_op;
        tex_descs.push_back(d);
    }
    dev.textures = upload(tex_descs);
    dev.tex_ops = upload(tex_ops);
    for (MaterialDesc &m : materials) {  // fold literal / "constant" parameters into the material table
        m.const_mask = 0;
        for (int slot = 0; slot < SLOT_COUNT; ++slot) {
            m.scalar_const[slot] = 0.0f;
            if (m.tex[slot] < 0) continue;
            const TextureDesc &d = tex_descs[m.tex[slot]];
            const TexOp &op = tex_ops[d.first_op];
            if (d.n_ops != 1 || op.op != TEXOP_CONST) continue;
            m.const_mask |= 1u << slot;
            m.scalar_const[slot] = op.value[0];
            if (slot == SLOT_COLOR) for (int c = 0; c < 3; ++c) m.color_const[c] = op.value[c];
        }
    }
    dev.materials = upload(materials);
    // 3-D textures: one descriptor per node; an add / mul node points at its postfix program, stored behind the node table
    std::vector<Texture3DDesc> t3, t3_ops;
    auto leaf3 = [&](const Tex3DNode &t) {
        Texture3DDesc d{};
        d.is_array = t.is_array ? TEX3D_ARRAY : TEX3D_CONST;
        d.value[0] = t.value.x; d.value[1] = t.value.y; d.value[2] = t.value.z;
        d.tex = t.is_array ? (unsigned long long)sources[t.source].tex : 0ull;
        return d;
    };
    std::function<int(int, int)> emit3 = [&](int i, int depth) -> int {  // returns the value-stack depth the subtree needs
        const Tex3DNode &t = textures3d[i];
        if (t.op == 0) { t3_ops.push_back(leaf3(t)); return depth + 1; }
        const int dl = emit3(t.lhs, depth), dr = emit3(t.rhs, depth + 1);
        Texture3DDesc d{};
        d.is_array = t.op;
        t3_ops.push_back(d);
        return std::max(dl, dr);
    };
    for (size_t i = 0; i < textures3d.size(); ++i) {
        const Tex3DNode &t = textures3d[i];
        if (t.op == 0) { t3.push_back(leaf3(t)); continue; }
        const size_t first = t3_ops.size();
        if (emit3((int)i, 0) > TEX3D_STACK) throw Error("texture3d expression is nested too deeply");
        Texture3DDesc d{};
        d.is_array = TEX3D_PROGRAM;
        d.tex = (unsigned long long)(textures3d.size() + first) | ((unsigned long long)(t3_ops.size() - first) << 32);
        t3.push_back(d);
    }
    t3.insert(t3.end(), t3_ops.begin(), t3_ops.end());
    dev.textures3d = upload(t3);
    {   // HaltonSampler::initialize_consts (builtin/sampler/halton.cpp:116-133) + the prime table (utils/prime.cpp)
        const int res[2] = {width, height};
        for (int i = 0; i < 2; ++i) {
            const int base = i == 0 ? 2 : 3;
            int scale = 1, exp = 0;
            while (scale < std::min(res[i], 128)) { scale *= base; ++exp; }
            dev.halton.base_scales[i] = scale;
            dev.halton.base_exponents[i] = exp;
        }
        auto mult_inverse = [](long long a, long long n) {  // extended Euclid (halton.cpp:12-37)
            long long x0 = 1, x1 = 0, r0 = a, r1 = n;
            while (r1 != 0) { const long long q = r0 / r1, t = r0 - q * r1; r0 = r1; r1 = t; const long long u = x0 - q * x1; x0 = x1; x1 = u; }
            const long long m = x0 - x0 / n * n;
            return (int)(m < 0 ? m + n : m);
        };
        dev.halton.mult_inverse[0] = mult_inverse(dev.halton.base_scales[1], dev.halton.base_scales[0]);
        dev.halton.mult_inverse[1] = mult_inverse(dev.halton.base_scales[0], dev.halton.base_scales[1]);
        std::vector<int> primes;
        for (int v = 2; (int)primes.size() < 1000; ++v) {
            bool is_prime = true;
            for (int q : primes) { if (q * q > v) break; if (v % q == 0) { is_prime = false; break; } }
            if (is_prime) primes.push_back(v);
        }
        dev.halton.primes = upload(primes);
    }
    for (int a = 0; a < 3; ++a) {  // ray-sort grid: 8 cells per axis over the world box
        const float ext = scene_box.hi[a] - scene_box.lo[a];
        dev.sort_lo[a] = scene_box.lo[a];
        dev.sort_scale[a] = ext > 0 ? 8.0f / ext : 0.0f;
    }
    dev.media = upload(medium_table);
    dev.vol_prims = volume_prims.empty() ? nullptr : upload(volume_prims);
    dev.n_vol_prims = (int)volume_prims.size();
    dev.has_medium = has_medium ? 1 : 0;
    if (has_medium)  // 2: some 3-D texture is an "add" / "mul" node, the medium / shadow kernels must be the instantiations that run programs
        for (const Tex3DNode &t : textures3d) if (t.op != 0) dev.has_medium = 2;
    dev.lights = upload(lights);
    dev.n_instances = (int)instance_info.size();
    // ---- environment light (core/light_sampler.cpp:24-33: appended after the area lights)
    dev.env = EnvLight{};
    dev.env.type = env.type;
    dev.env.light_index = env.type != ENV_NONE ? (int)lights.size() : -1;
    for (int k = 0; k < 3; ++k) { dev.env.lower[k] = env.lower[k]; dev.env.upper[k] = env.upper[k]; dev.env.up[k] = env.up[k]; }
    dev.n_lights = (int)lights.size() + (env.type != ENV_NONE ? 1 : 0);
    dev.vol_prim_medium_id = (uint32_t)medium_table.size() - 1;
    if (env.type == ENV_IBL) {  // IBL::commit -> EnvirLightSampler::preprocess (ibl.cpp:27-31, env_sampler.cpp:72-122)
        dev.env.ibl_tex = env.tex;
        dev.env.lut_w = env.lut_w; dev.env.lut_h = env.lut_h;
        {  // Frame::from_z(up) (utils/math/frame.cpp:45-57)
            const Vec3 nz = normalize(env.up);
            const Vec3 ny = 1 - std::fabs(nz.x) < 0.1f ? cross(nz, Vec3(0, 1, 0)) : cross(nz, Vec3(1, 0, 0));
            const Vec3 fx = normalize(cross(nz, ny)), fy = normalize(ny), fz = normalize(nz);
            for (int k = 0; k < 3; ++k) { dev.env.frame[k] = fx[k]; dev.env.frame[3 + k] = fy[k]; dev.env.frame[6 + k] = fz[k]; }
        }
        const size_t nt = (size_t)env.lut_w * env.lut_h;
        if (nt == 0) throw Error("ibl: empty lookup table");
        DeviceBuffer d_lum(nt * sizeof(float)), d_area(nt * sizeof(float));
        launch_ibl_lut(dev, env.tex, env.lut_w, env.lut_h, 256, d_lum.as<float>(), d_area.as<float>(), nullptr);
        BTRC_CUDA(cudaDeviceSynchronize());
        BTRC_CUDA(cudaGetLastError());
        std::vector<float> lum(nt), area(nt);
        BTRC_CUDA(cudaMemcpy(lum.data(), d_lum.ptr, nt * sizeof(float), cudaMemcpyDeviceToHost));
        BTRC_CUDA(cudaMemcpy(area.data(), d_area.ptr, nt * sizeof(float), cudaMemcpyDeviceToHost));
        float lum_area_sum = 0.0f, area_sum = 0.0f;
        for (size_t i = 0; i < nt; ++i) { lum_area_sum += lum[i] * area[i]; area_sum += area[i]; }
        const float avg_lum = lum_area_sum / area_sum;
        for (size_t i = 0; i < nt; ++i) lum[i] = std::max(0.0f, area[i] * (lum[i] - avg_lum));
        float table_sum = 0.0f;
        for (float v : lum) table_sum += v;
        if (table_sum > 1e-3f) { const float ratio = 1 / table_sum; for (float &v : lum) v *= ratio; }
        ibl_tile_probs = lum;
        ibl_alias = build_alias_table(lum);
        dev.env.tile_probs = upload(ibl_tile_probs);
        dev.env.tile_alias = reinterpret_cast<const float2 *>(upload(ibl_alias));
    }
    dev.world_diagonal = world_diagonal;
    device_ready_ = true;
}

}  // namespace btrc
