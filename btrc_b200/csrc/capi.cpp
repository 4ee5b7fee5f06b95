// C ABI of libbtrc_b200.so (declared in include/btrc_b200.h).  No exception crosses this boundary.
#include "../../include/btrc_b200.h"

#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "host/image_io.h"
#include "host/comm.h"
#include "host/renderer.h"
#include "host/scene.h"

using namespace btrc;

struct btrc_scene {
    std::unique_ptr<Scene> scene;
    std::unique_ptr<Renderer> renderer;
    int state_count_override = 0;
    btrc_progress_fn progress_fn = nullptr;
    void *progress_user = nullptr;
    void apply_reporter() {
        if (progress_fn) {
            btrc_progress_fn fn = progress_fn;
            void *user = progress_user;
            renderer->set_reporter([fn, user](float pct) { return fn(pct, user) != 0; });
        } else {
            renderer->set_reporter(nullptr);
        }
    }
    Renderer &r() {
        scene->upload_device();
        if (!renderer) {
            renderer = std::make_unique<Renderer>(*scene, state_count_override);
            apply_reporter();
        }
        return *renderer;
    }
};

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F &&f) {
    try {
        f();
        g_last_error.clear();
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    } catch (...) {
        g_last_error = "unknown error";
        return 1;
    }
}

void fill_stats(const RenderStats &s, btrc_render_stats_t *out) {
    if (!out) return;
    out->n_gen = s.n_gen; out->n_closest = s.n_closest; out->n_hit = s.n_hit; out->n_next = s.n_next;
    out->n_shadow = s.n_shadow; out->n_shadow_miss = s.n_shadow_miss; out->n_splat = s.n_splat; out->n_iter = s.n_iter;
    out->render_ms = s.render_ms;
    for (int k = 0; k < 6; ++k) out->stage_ms[k] = s.stage_ms[k];
    out->n_launches = s.n_launches;
}

// copies between any two address spaces (unified virtual addressing)
void copy_any(void *dst, const void *src, size_t bytes) {
    if (!bytes) return;
    auto on_host = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }  // no device at all: plain host memory
        return a.type == cudaMemoryTypeUnregistered || a.type == cudaMemoryTypeHost;
    };
    if (on_host(dst) && on_host(src)) { std::memcpy(dst, src, bytes); return; }  // post-processors on host images need no device
    // The renderer's stream is non-blocking: it does not order itself against the legacy default stream, and a pageable
    // host-to-device cudaMemcpy may return once the data is staged.  Copy, then wait for the device, so that a stage kernel
    // launched next on the renderer's stream can never read a queue before its upload has landed.
    BTRC_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    BTRC_CUDA(cudaStreamSynchronize(0));
}

// stage entry points write caller batches into pool-sized device queues: refuse before the first copy
void check_batch(Renderer &r, long long n) {
    if (n > (long long)r.state_count()) throw Error("stage batch larger than state_count");
}

void check_scene(btrc_scene_t *s) {
    if (!s || !s->scene) throw Error("null scene handle");
}

void resolve_to(btrc_scene_t *s, float *color, float *albedo, float *normal) { s->r().resolve_any(color, albedo, normal); }

struct HostPaths {
    std::vector<float4> ray_o, ray_d, beta, rad, ble;
    std::vector<ulonglong2> rng;
    explicit HostPaths(size_t n) : ray_o(n), ray_d(n), beta(n), rad(n), ble(n), rng(n) {}
};

void upload_paths(Renderer &r, int which, int n, int width, const float *o_med, const float *d_t1, const uint32_t *pixel,
                  const float *beta_depth, const float *path_rad, const uint64_t *rng, const float *ble) {
    const size_t N = (size_t)n;
    std::vector<uint32_t> px(2 * N);
    std::vector<float4> rad(N);
    copy_any(px.data(), pixel, 8 * N);
    copy_any(rad.data(), path_rad, 16 * N);
    for (size_t i = 0; i < N; ++i) {
        const uint32_t idx = px[2 * i + 1] * (uint32_t)width + px[2 * i];
        std::memcpy(&rad[i].w, &idx, 4);
    }
    PathQueue &q = r.queue(which);
    copy_any(q.ray_o, o_med, 16 * N);
    copy_any(q.ray_d, d_t1, 16 * N);
    copy_any(q.beta, beta_depth, 16 * N);
    copy_any(q.rad, rad.data(), 16 * N);
    copy_any(q.rng, rng, 16 * N);
    copy_any(q.ble, ble, 16 * N);
}

void download_paths(Renderer &r, int which, int n, int width, float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth,
                    float *path_rad, uint64_t *rng, float *ble) {
    const size_t N = (size_t)n;
    PathQueue &q = r.queue(which);
    std::vector<float4> rad(N);
    std::vector<uint32_t> px(2 * N);
    copy_any(rad.data(), q.rad, 16 * N);
    for (size_t i = 0; i < N; ++i) {
        uint32_t idx;
        std::memcpy(&idx, &rad[i].w, 4);
        px[2 * i] = idx % (uint32_t)width;
        px[2 * i + 1] = idx / (uint32_t)width;
        rad[i].w = 1.0f;  // reference layout: path_radiance.w = 1 (soa.cpp:49)
    }
    copy_any(o_med, q.ray_o, 16 * N);
    copy_any(d_t1, q.ray_d, 16 * N);
    copy_any(beta_depth, q.beta, 16 * N);
    copy_any(path_rad, rad.data(), 16 * N);
    copy_any(pixel, px.data(), 8 * N);
    copy_any(rng, q.rng, 16 * N);
    copy_any(ble, q.ble, 16 * N);
}

}  // namespace

extern "C" {

const char *btrc_last_error(void) { return g_last_error.c_str(); }
const char *btrc_version(void) { return "btrc_b200 0.1 (sm_100a)"; }

int btrc_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

int btrc_wave_ctl_bytes(void) { return (int)sizeof(btrc::WaveCtl); }

int btrc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int btrc_set_device(int32_t device) {
    return guarded([&] { BTRC_CUDA(cudaSetDevice(device)); });
}

int btrc_scene_load_json(const char *path, btrc_scene_t **out) {
    return guarded([&] {
        if (!path || !out) throw Error("null argument");
        auto h = std::make_unique<btrc_scene>();
        h->scene = Scene::from_file(path);
        h->scene->commit();
        *out = h.release();
    });
}

int btrc_scene_load_json_string(const char *json, const char *scene_dir, btrc_scene_t **out) {
    return guarded([&] {
        if (!json || !out) throw Error("null argument");
        auto h = std::make_unique<btrc_scene>();
        h->scene = Scene::from_string(json, scene_dir ? scene_dir : ".");
        h->scene->commit();
        *out = h.release();
    });
}

void btrc_scene_free(btrc_scene_t *scene) { delete scene; }

int btrc_scene_info_get(btrc_scene_t *s, btrc_scene_info_t *o) {
    return guarded([&] {
        check_scene(s);
        const Scene &sc = *s->scene;
        o->width = sc.width; o->height = sc.height;
        o->spp = sc.renderer.spp; o->min_depth = sc.renderer.min_depth; o->max_depth = sc.renderer.max_depth;
        o->state_count = s->state_count_override > 0 ? s->state_count_override : sc.renderer.state_count;
        o->tile = sc.renderer.tile; o->albedo = sc.renderer.albedo; o->normal = sc.renderer.normal;
        o->n_instances = (int)sc.instances.size(); o->n_lights = (int)sc.lights.size() + (sc.env.type != 0 ? 1 : 0);
        o->n_media = (int)sc.medium_table.size(); o->has_medium = sc.has_medium ? 1 : 0; o->n_textures = (int)sc.textures.size();
        o->n_triangles = (int)sc.world_tris.size(); o->n_materials = (int)sc.materials.size();
        o->bvh_depth = sc.bvh_depth;
        o->world_diagonal = sc.world_diagonal;
        o->bvh_build_seconds = (float)sc.bvh_build_seconds;
        o->bvh_bytes = sc.renderer.lbvh ? 64ull * (uint64_t)sc.lbvh_nodes + 48ull * sc.world_tris.size()
                       : sc.instanced_bvh ? (uint64_t)(sc.two_level.bl_nodes.size() + sc.two_level.tl_nodes.size() + sc.two_level.tris.size()) * sizeof(float)
                       : sc.wide_bvh8     ? (uint64_t)(sc.wide_bvh.nodes.size() + sc.wide_bvh.tris.size()) * sizeof(float)
                                          : (uint64_t)(sc.flat_bvh.nodes.size() + sc.flat_bvh.tris.size()) * sizeof(float);
        o->bvh_instanced = sc.instanced_bvh ? 1 : 0;
        o->bvh_wide = sc.wide_bvh8 ? 1 : 0;
    });
}

int btrc_set_params(btrc_scene_t *s, int32_t spp, int32_t state_count, int32_t min_depth, int32_t max_depth) {
    return guarded([&] {
        check_scene(s);
        if (s->renderer && state_count > 0) throw Error("state_count must be set before the first render");
        if (spp > 0) s->scene->renderer.spp = spp;
        if (state_count > 0) s->state_count_override = state_count;
        if (min_depth >= 0) s->scene->renderer.min_depth = min_depth;
        if (max_depth >= 0) s->scene->renderer.max_depth = max_depth;
    });
}

int btrc_set_profiling(btrc_scene_t *s, int32_t on) {
    return guarded([&] { check_scene(s); s->r().set_profiling(on != 0); });
}

void *btrc_stream(btrc_scene_t *s) {
    void *out = nullptr;
    guarded([&] { check_scene(s); out = (void *)s->r().stream(); });
    return out;
}

int btrc_film_clear(btrc_scene_t *s) {
    return guarded([&] { check_scene(s); s->r().clear_film(); });
}

int btrc_render_range(btrc_scene_t *s, int64_t begin, int64_t end, btrc_render_stats_t *stats) {
    bool stopped = false;
    const int rc = guarded([&] {
        check_scene(s);
        s->r().render_range(begin, end);
        stopped = s->r().stopped();
        fill_stats(s->r().stats(), stats);
    });
    if (rc == 0 && stopped) { g_last_error = "render stopped by the reporter"; return BTRC_STOPPED; }
    return rc;
}

int btrc_render_part(btrc_scene_t *s, int32_t n_parts, int32_t part, int32_t rows_per_strip, btrc_render_stats_t *stats) {
    bool stopped = false;
    const int rc = guarded([&] {
        check_scene(s);
        s->r().render_part(n_parts, part, rows_per_strip);
        stopped = s->r().stopped();
        fill_stats(s->r().stats(), stats);
    });
    if (rc == 0 && stopped) { g_last_error = "render stopped by the reporter"; return BTRC_STOPPED; }
    return rc;
}

int btrc_nccl_unique_id(uint8_t *out128) {
    return guarded([&] { if (!out128) throw Error("null id buffer"); nccl_unique_id(out128); });
}

int btrc_comm_create(int32_t n_ranks, int32_t rank, const uint8_t *id128, btrc_comm_t **out) {
    return guarded([&] {
        if (!out) throw Error("null output handle");
        *out = nullptr;
        if (n_ranks > 1 && !id128) throw Error("null NCCL id");
        static const uint8_t zero[kNcclIdBytes] = {0};
        *out = reinterpret_cast<btrc_comm_t *>(comm_create(n_ranks, rank, id128 ? id128 : zero));
    });
}

void btrc_comm_destroy(btrc_comm_t *comm) { comm_destroy(reinterpret_cast<Comm *>(comm)); }

int btrc_reduce_films(btrc_scene_t *s, btrc_comm_t *comm, int32_t root) {
    return guarded([&] {
        check_scene(s);
        Renderer &r = s->r();
        const size_t npix = (size_t)s->scene->width * s->scene->height;
        FilmDevice &f = r.film();
        float *bufs[4] = {f.radiance, f.weight, f.albedo, f.normal};
        const size_t counts[4] = {4 * npix, npix, 4 * npix, 4 * npix};
        comm_reduce_sum(reinterpret_cast<Comm *>(comm), bufs, counts, 4, root, r.stream());
    });
}

int btrc_set_reporter(btrc_scene_t *s, btrc_progress_fn fn, void *user) {
    return guarded([&] {
        check_scene(s);
        s->progress_fn = fn;
        s->progress_user = user;
        if (s->renderer) s->apply_reporter();
    });
}

int btrc_render(btrc_scene_t *s, int memory, float *color, float *albedo, float *normal, btrc_render_stats_t *stats) {
    (void)memory;
    bool stopped = false;
    const int rc = guarded([&] {
        check_scene(s);
        Renderer &r = s->r();
        r.clear_film();
        r.render_range(0, r.total_paths());
        stopped = r.stopped();
        if (!stopped) resolve_to(s, color, albedo, normal);
        fill_stats(r.stats(), stats);
    });
    if (rc == 0 && stopped) { g_last_error = "render stopped by the reporter"; return BTRC_STOPPED; }
    return rc;
}

int btrc_resolve(btrc_scene_t *s, int memory, float *color, float *albedo, float *normal) {
    (void)memory;
    return guarded([&] { check_scene(s); resolve_to(s, color, albedo, normal); });
}

int btrc_film_export(btrc_scene_t *s, int memory, float *radiance, float *weight, float *albedo, float *normal) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        Renderer &r = s->r();
        r.export_film(radiance, weight, albedo, normal);
    });
}

int btrc_film_import(btrc_scene_t *s, int memory, const float *radiance, const float *weight, const float *albedo, const float *normal) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        Renderer &r = s->r();
        r.import_film(radiance, weight, albedo, normal);
    });
}

int btrc_image_load(const char *path, int32_t *width, int32_t *height, int32_t *channels, int32_t *is_float, void *out, uint64_t out_bytes) {
    return guarded([&] {
        if (!path) throw Error("null path");
        const LoadedImage img = load_image(path);
        if (width) *width = img.width;
        if (height) *height = img.height;
        if (channels) *channels = img.is_float ? 3 : img.channels;
        if (is_float) *is_float = img.is_float ? 1 : 0;
        if (!out) return;
        const size_t need = img.is_float ? img.f32.size() * sizeof(float) : img.u8.size();
        if (out_bytes < need) throw Error("btrc_image_load: buffer too small");
        std::memcpy(out, img.is_float ? (const void *)img.f32.data() : (const void *)img.u8.data(), need);
    });
}

namespace {
// Canonical text of a parsed + resolved scene file: groups in first-insertion key order {"k":v,...}, arrays [v,...], values "text"
void dump_node(const NodeP &n, std::string &out, int depth) {
    if (depth > 200) { out += "<deep>"; return; }
    if (n->is_group()) {
        out += "{";
        bool first = true;
        for (const auto &kv : n->children) {
            if (!first) out += ",";
            first = false;
            out += "\"" + kv.first + "\":";
            dump_node(kv.second, out, depth + 1);
        }
        out += "}";
    } else if (n->is_array()) {
        out += "[";
        for (size_t i = 0; i < n->elements.size(); ++i) {
            if (i) out += ",";
            dump_node(n->elements[i], out, depth + 1);
        }
        out += "]";
    } else
        out += "\"" + n->value + "\"";
}
}  // namespace

int btrc_scene_file_dump(const char *path, char *out, uint64_t out_bytes, uint64_t *needed) {
    return guarded([&] {
        if (!path) throw Error("null path");
        std::string text;
        dump_node(parse_scene_file(path), text, 0);
        if (needed) *needed = text.size() + 1;
        if (!out) return;
        if (out_bytes < text.size() + 1) throw Error("btrc_scene_file_dump: buffer too small");
        std::memcpy(out, text.c_str(), text.size() + 1);
    });
}

int btrc_scene_file_parse(const char *path, const char *node_path, int32_t kind, float *out, int32_t *count) {
    return guarded([&] {
        if (!path || !node_path || !out) throw Error("null argument");
        NodeP n = parse_scene_file(path);
        std::string cur;
        auto step = [&](const std::string &sec) {
            if (sec.empty()) return;
            if (n->is_array()) {
                const size_t i = std::stoul(sec);
                if (i >= n->elements.size()) throw Error("invalid node path: " + std::string(node_path));
                n = n->elements[i];
            } else
                n = n->child(sec);
        };
        for (const char *p = node_path; *p; ++p) {
            if (*p == '/') { step(cur); cur.clear(); }
            else cur += *p;
        }
        step(cur);
        int c = 1;
        switch (kind) {
        case 0: out[0] = (float)n->as_int(); break;
        case 1: out[0] = n->as_float(); break;
        case 2: out[0] = n->as_bool() ? 1.0f : 0.0f; break;
        case 3: case 4: { const Vec3 v = n->as_vec3(); out[0] = v.x; out[1] = v.y; out[2] = v.z; c = 3; break; }
        case 5: out[0] = n->as_degree(); break;
        case 6: out[0] = n->as_radian(); break;
        case 7: { const Transform3D t = n->as_transform3d(); for (int r = 0; r < 4; ++r) for (int k = 0; k < 4; ++k) out[4 * r + k] = t.mat.m[r][k]; c = 16; break; }
        case 8: { const Mat3 t = n->as_transform2d(); for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) out[3 * r + k] = t.m[r][k]; c = 9; break; }
        default: throw Error("btrc_scene_file_parse: unknown kind");
        }
        if (count) *count = c;
    });
}

int btrc_run_post_processors(btrc_scene_t *s, int memory, float *color, float *albedo, float *normal) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        const int w = s->scene->width, h = s->scene->height;
        const size_t n = 4 * (size_t)w * h;
        std::vector<float> c(color ? n : 0), a(albedo ? n : 0), nn(normal ? n : 0);
        if (color) copy_any(c.data(), color, n * 4);
        if (albedo) copy_any(a.data(), albedo, n * 4);
        if (normal) copy_any(nn.data(), normal, n * 4);
        for (const PostProcessorDesc &p : s->scene->post_processors) {
            if (p.type == "tonemap") {  // postprocess/tonemap.cpp:14-22,35-51 (ACES fit on exposure-scaled colour, in place)
                if (!color) continue;
                auto aces = [](float x) {
                    constexpr float tA = 2.51f, tB = 0.03f, tC = 2.43f, tD = 0.59f, tE = 0.14f;
                    return std::fmin(std::fmax((x * (tA * x + tB)) / (x * (tC * x + tD) + tE), 0.0f), 1.0f);
                };
                for (size_t i = 0; i < n; i += 4)
                    for (int k = 0; k < 3; ++k) c[i + k] = aces(p.exposure * c[i + k]);
                copy_any(color, c.data(), n * 4);
            } else if (p.type == "save") {  // postprocess/save_to_image.cpp:30-62
                if (color && !p.color.empty()) {
                    std::vector<float> g = c;
                    const float e = 1 / p.gamma;
                    for (float &v : g) v = std::pow(v, e);  // Image<Vec4f>::pow_ (all four components)
                    save_image_rgba32f(p.color, w, h, g.data());
                }
                if (albedo && !p.albedo.empty()) save_image_rgba32f(p.albedo, w, h, a.data());
                if (normal && !p.normal.empty()) save_image_rgba32f(p.normal, w, h, nn.data());
            }
        }
    });
}

int btrc_closest_hit(btrc_scene_t *s, int memory, const float *o_med, const float *d_t1, int32_t n, uint32_t *flag, uint32_t *detail) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        if (n <= 0) return;
        const size_t N = (size_t)n;
        s->scene->upload_device();
        DeviceBuffer o(16 * N), d(16 * N), f(4 * N), det(16 * N);
        copy_any(o.ptr, o_med, 16 * N);
        copy_any(d.ptr, d_t1, 16 * N);
        LaunchCfg cfg;
        cfg.sm_count = btrc_device_sm_count();
        launch_trace_batch(cfg, s->scene->dev, o.as<float4>(), d.as<float4>(), n, f.as<uint32_t>(), det.as<uint4>(), nullptr);
        BTRC_CUDA(cudaDeviceSynchronize());
        BTRC_CUDA(cudaGetLastError());
        copy_any(flag, f.ptr, 4 * N);
        copy_any(detail, det.ptr, 16 * N);
    });
}

int btrc_any_hit(btrc_scene_t *s, int memory, const float *o_med, const float *d_t1, int32_t n, uint8_t *occluded) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        if (n <= 0) return;
        const size_t N = (size_t)n;
        s->scene->upload_device();
        DeviceBuffer o(16 * N), d(16 * N), occ(N);
        copy_any(o.ptr, o_med, 16 * N);
        copy_any(d.ptr, d_t1, 16 * N);
        LaunchCfg cfg;
        cfg.sm_count = btrc_device_sm_count();
        launch_anyhit_batch(cfg, s->scene->dev, o.as<float4>(), d.as<float4>(), n, occ.as<uint8_t>(), nullptr);
        BTRC_CUDA(cudaDeviceSynchronize());
        BTRC_CUDA(cudaGetLastError());
        copy_any(occluded, occ.ptr, N);
    });
}

int btrc_generate(btrc_scene_t *s, int memory, int64_t first, int32_t n, float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth,
                  float *path_rad, uint64_t *rng, float *ble) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        if (n <= 0) return;
        Renderer &r = s->r();
        r.stage_generate(first, n);
        download_paths(r, 0, n, s->scene->width, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble);
    });
}

// shared body of btrc_shade / btrc_medium: upload the batch, run the stage, download compacted outputs
static void run_state_stage(btrc_scene_t *s, bool medium, int32_t n, float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth,
                            const float *path_rad, uint64_t *rng, float *ble, uint32_t *flag, const uint32_t *detail, int32_t *n_next,
                            float *next_o_med, float *next_d_t1, uint32_t *next_pixel, float *next_beta_depth, float *next_path_rad,
                            uint64_t *next_rng, float *next_ble, int32_t *n_shadow, uint32_t *shadow_pixel, float *shadow_beta_li,
                            float *shadow_o_med, float *shadow_d_t1) {
    check_scene(s);
    *n_next = 0; *n_shadow = 0;
    if (n <= 0) return;
    Renderer &r = s->r();
    check_batch(r, n);
    const int W = s->scene->width;
    upload_paths(r, 0, n, W, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble);
    copy_any(r.hits().flag, flag, 4 * (size_t)n);
    copy_any(r.hits().detail, detail, 16 * (size_t)n);
    int nn = 0, ns = 0;
    if (medium) {
        r.stage_medium(n, &nn, &ns);
        // in-place results of the medium stage: resolved medium id, beta_le * Tr, advanced sampler, scattered flag
        copy_any(o_med, r.queue(0).ray_o, 16 * (size_t)n);
        copy_any(ble, r.queue(0).ble, 16 * (size_t)n);
        copy_any(rng, r.queue(0).rng, 16 * (size_t)n);
        copy_any(flag, r.hits().flag, 4 * (size_t)n);
    } else
        r.stage_shade(n, &nn, &ns);
    *n_next = nn; *n_shadow = ns;
    if (nn) download_paths(r, 1, nn, W, next_o_med, next_d_t1, next_pixel, next_beta_depth, next_path_rad, next_rng, next_ble);
    if (ns) {
        const size_t N = (size_t)ns;
        std::vector<float4> b(N);
        std::vector<uint32_t> px(2 * N);
        copy_any(b.data(), r.shadows().beta_li, 16 * N);
        for (size_t i = 0; i < N; ++i) {
            uint32_t idx;
            std::memcpy(&idx, &b[i].w, 4);
            px[2 * i] = idx % (uint32_t)W; px[2 * i + 1] = idx / (uint32_t)W;
            b[i].w = 1.0f;  // soa.cpp:112
        }
        copy_any(shadow_beta_li, b.data(), 16 * N);
        copy_any(shadow_pixel, px.data(), 8 * N);
        copy_any(shadow_o_med, r.shadows().ray_o, 16 * N);
        copy_any(shadow_d_t1, r.shadows().ray_d, 16 * N);
    }
}

int btrc_shade(btrc_scene_t *s, int memory, int32_t n, const float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth,
               const float *path_rad, const uint64_t *rng, const float *ble, const uint32_t *flag, const uint32_t *detail, int32_t *n_next,
               float *next_o_med, float *next_d_t1, uint32_t *next_pixel, float *next_beta_depth, float *next_path_rad, uint64_t *next_rng,
               float *next_ble, int32_t *n_shadow, uint32_t *shadow_pixel, float *shadow_beta_li, float *shadow_o_med, float *shadow_d_t1) {
    (void)memory;
    return guarded([&] {
        run_state_stage(s, false, n, const_cast<float *>(o_med), d_t1, pixel, beta_depth, path_rad, const_cast<uint64_t *>(rng), const_cast<float *>(ble),
                        const_cast<uint32_t *>(flag), detail, n_next, next_o_med, next_d_t1, next_pixel, next_beta_depth, next_path_rad, next_rng,
                        next_ble, n_shadow, shadow_pixel, shadow_beta_li, shadow_o_med, shadow_d_t1);
    });
}

int btrc_medium(btrc_scene_t *s, int memory, int32_t n, float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth,
                const float *path_rad, uint64_t *rng, float *ble, uint32_t *flag, const uint32_t *detail, int32_t *n_next,
                float *next_o_med, float *next_d_t1, uint32_t *next_pixel, float *next_beta_depth, float *next_path_rad, uint64_t *next_rng,
                float *next_ble, int32_t *n_shadow, uint32_t *shadow_pixel, float *shadow_beta_li, float *shadow_o_med, float *shadow_d_t1) {
    (void)memory;
    return guarded([&] {
        run_state_stage(s, true, n, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble, flag, detail, n_next, next_o_med, next_d_t1, next_pixel,
                        next_beta_depth, next_path_rad, next_rng, next_ble, n_shadow, shadow_pixel, shadow_beta_li, shadow_o_med, shadow_d_t1);
    });
}

int btrc_shadow(btrc_scene_t *s, int memory, int32_t n, const uint32_t *shadow_pixel, const float *shadow_beta_li, const float *shadow_o_med,
                const float *shadow_d_t1) {
    (void)memory;
    return guarded([&] {
        check_scene(s);
        if (n <= 0) return;
        Renderer &r = s->r();
        check_batch(r, n);
        const size_t N = (size_t)n;
        const int W = s->scene->width;
        std::vector<float4> b(N);
        std::vector<uint32_t> px(2 * N);
        copy_any(b.data(), shadow_beta_li, 16 * N);
        copy_any(px.data(), shadow_pixel, 8 * N);
        for (size_t i = 0; i < N; ++i) {
            const uint32_t idx = px[2 * i + 1] * (uint32_t)W + px[2 * i];
            std::memcpy(&b[i].w, &idx, 4);
        }
        copy_any(r.shadows().beta_li, b.data(), 16 * N);
        copy_any(r.shadows().ray_o, shadow_o_med, 16 * N);
        copy_any(r.shadows().ray_d, shadow_d_t1, 16 * N);
        r.stage_shadow(n);
    });
}

int btrc_geometry_tables(btrc_scene_t *s, int32_t geometry_id, int32_t *n_triangles, float *out24) {
    return guarded([&] {
        check_scene(s);
        if (geometry_id < 0 || geometry_id >= (int)s->scene->meshes.size()) throw Error("geometry id out of range");
        const MeshData &m = s->scene->meshes[geometry_id];
        if (n_triangles) *n_triangles = m.ntri;
        if (!out24) return;
        for (int i = 0; i < m.ntri; ++i) {
            float *o = out24 + 24 * (size_t)i;
            std::memcpy(o, &m.gx_ua[4 * i], 16); std::memcpy(o + 4, &m.gy_uba[4 * i], 16); std::memcpy(o + 8, &m.gz_uca[4 * i], 16);
            std::memcpy(o + 12, &m.sn_va[4 * i], 16); std::memcpy(o + 16, &m.sn_vba[4 * i], 16); std::memcpy(o + 20, &m.sn_vca[4 * i], 16);
        }
    });
}

int btrc_geometry_positions(btrc_scene_t *s, int32_t geometry_id, int32_t *n_triangles, float *out9) {
    return guarded([&] {
        check_scene(s);
        if (geometry_id < 0 || geometry_id >= (int)s->scene->meshes.size()) throw Error("geometry id out of range");
        const MeshData &m = s->scene->meshes[geometry_id];
        if (n_triangles) *n_triangles = m.ntri;
        if (!out9) return;
        for (size_t i = 0; i < m.pos.size(); ++i) { out9[3 * i] = m.pos[i].x; out9[3 * i + 1] = m.pos[i].y; out9[3 * i + 2] = m.pos[i].z; }
    });
}

int btrc_ibl_table(btrc_scene_t *s, int32_t *lut_w, int32_t *lut_h, float *probs) {
    return guarded([&] {
        check_scene(s);
        s->scene->upload_device();
        if (s->scene->env.type != ENV_IBL) throw Error("the scene has no ibl environment light");
        if (lut_w) *lut_w = s->scene->env.lut_w;
        if (lut_h) *lut_h = s->scene->env.lut_h;
        if (probs) std::memcpy(probs, s->scene->ibl_tile_probs.data(), sizeof(float) * s->scene->ibl_tile_probs.size());
    });
}

int btrc_instance_geometry(btrc_scene_t *s, int32_t instance_id, int32_t *geometry_id) {
    return guarded([&] {
        check_scene(s);
        if (instance_id < 0 || instance_id >= (int)s->scene->instance_info.size()) throw Error("instance id out of range");
        *geometry_id = s->scene->instance_info[instance_id].geometry_id;
    });
}

int btrc_instance_transform(btrc_scene_t *s, int32_t instance_id, float *mat12, float *inv12) {
    return guarded([&] {
        check_scene(s);
        if (instance_id < 0 || instance_id >= (int)s->scene->instance_info.size()) throw Error("instance id out of range");
        std::memcpy(mat12, s->scene->instance_info[instance_id].mat, 48);
        std::memcpy(inv12, s->scene->instance_info[instance_id].inv, 48);
    });
}

}  // extern "C"
