// ORACLE / reference build (test infrastructure only; compiled only where /root/reference exists, into oracle/_ref/libbtrcrefcpu.so).
//
// The REFERENCE'S OWN wavefront path tracer, executed on the host.  Everything on the path below is reference code compiled
// from the sources where they lie (oracle/Makefile, `refcpu` target):
//   * scene files: factory::JSONParser + factory::Context + every builtin creator (src/factory, src/builtin/register.cpp:31-99);
//   * scene commit: Scene::commit, TriangleMesh::commit, textures, media, lights, alias tables (src/core, src/builtin);
//   * device code: GeneratePipeline / MediumPipeline / ShadePipeline::record_device_code (wavefront/generate.cpp:26-88,
//     medium.cpp:19-165, shade.cpp:20-236) with all the materials, media, lights, filters and the camera they call, recorded
//     with CUJ and printed by CUJ's own LLVM-free CPPCodeGenerator (shim_ptxgen.cpp), compiled for the host;
//   * host loop: GeneratePipeline::generate / MediumPipeline::sample_scattering / ShadePipeline::shade launch code,
//     PreviewImageGenerator (preview.cpp), Film, the SoA buffers (soa_buffer.cpp).
// What is NOT reference code, because it cannot exist here:
//   * OptiX closest hit / occlusion -> the reference's vendored lib/bvh (shim_optix.cpp);
//   * the shadow pipeline's OptiX raygen/miss programs (wavefront/shadow.cpp:24-76) -> `shadow_kernel` below, which calls the
//     same reference functions (Medium::tr, IndependentSampler, Film::splat_atomic, CShadowRaySOA) in the same order;
//   * WavefrontPathTracer::commit / render (wavefront.cpp:98-325, needs OptiX pipelines) -> commit_renderer() / render() below,
//     the same calls in the same order;
//   * the CUDA runtime / texture unit -> shim_cuda.cpp (host memory; texture filtering per the CUDA programming guide).
// Floating point is IEEE fp32 without contraction (the reference's device build is fast-math): same contract as the oracle.
#include <btrc/builtin/camera/pinhole.h>
#include <btrc/builtin/film_filter/box.h>
#include <btrc/builtin/film_filter/gaussian.h>
#include <btrc/builtin/geometry/triangle_mesh.h>
#include <btrc/builtin/light/gradient_sky.h>
#include <btrc/builtin/light/ibl.h>
#include <btrc/builtin/light/mesh_light.h>
#include <btrc/builtin/material/black.h>
#include <btrc/builtin/material/diffuse.h>
#include <btrc/builtin/material/disney.h>
#include <btrc/builtin/material/glass.h>
#include <btrc/builtin/material/invisible.h>
#include <btrc/builtin/material/metal.h>
#include <btrc/builtin/material/mirror.h>
#include <btrc/builtin/medium/hetergeneous.h>
#include <btrc/builtin/medium/homogeneous.h>
#include <btrc/builtin/renderer/wavefront/generate.h>
#include <btrc/builtin/renderer/wavefront/medium.h>
#include <btrc/builtin/renderer/wavefront/preview.h>
#include <btrc/builtin/renderer/wavefront/shade.h>
#include <btrc/builtin/renderer/wavefront/soa_buffer.h>
#include <btrc/builtin/texture2d/array2d.h>
#include <btrc/builtin/texture2d/math.h>
#include <btrc/builtin/texture2d/transform.h>
#include <btrc/builtin/texture3d/array3d.h>
#include <btrc/builtin/texture3d/binary.h>
#include <btrc/core/object_dag.h>
#include <btrc/core/scene.h>
#include <btrc/factory/context.h>
#include <btrc/factory/node/parser.h>
#include <btrc/factory/scene.h>
#include <btrc/utils/exception.h>
#include <btrc/utils/file.h>

#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>

#include <unistd.h>

#include "refcpu_internal.h"

// utils/file.cpp does not compile with gcc (a bare `#elif`); its read_txt_file (file.cpp:13-22) reads the whole file as text
namespace btrc {
std::string read_txt_file(const std::string &filename)
{
    std::ifstream fin(filename, std::ios_base::in);
    if(!fin)
        throw BtrcException("failed to open text file: " + filename);
    std::stringstream sst;
    sst << fin.rdbuf();
    return sst.str();
}
// the reference keeps PTX caches next to its executable (utils/ptx_cache.cpp); here they would name host libraries that a later
// build may have removed, so every process gets a private directory
std::filesystem::path get_executable_filename()
{
    return std::filesystem::temp_directory_path() / ("refcpu_" + std::to_string(getpid())) / "refcpu";
}
}  // namespace btrc

using namespace btrc;
using namespace btrc::builtin;

namespace {

thread_local std::string last_error;

struct ShadowParams
{
    wfpt::ShadowRaySOA         shadow_ray;
    IndependentSampler::State *sampler_state;
    uint8_t                   *occluded;
};
CUJ_PROXY_CLASS(CShadowParams, ShadowParams, shadow_ray, sampler_state, occluded);

constexpr char SHADOW_KERNEL_NAME[] = "shadow_kernel";

// builtin/register.cpp:31-99 without the renderers and post-processors (their creators need the OptiX pipelines)
void register_creators(factory::Context &context)
{
    context.get_factory<Camera>().add_creator(newBox<PinholeCameraCreator>());
    context.get_factory<FilmFilter>().add_creator(newBox<BoxFilterCreator>());
    context.get_factory<FilmFilter>().add_creator(newBox<GaussianFilterCreator>());
    context.get_factory<Geometry>().add_creator(newBox<TriangleMeshCreator>());
    context.get_factory<Light>().add_creator(newBox<GradientSkyCreator>());
    context.get_factory<Light>().add_creator(newBox<IBLCreator>());
    context.get_factory<Light>().add_creator(newBox<MeshLightCreator>());
    auto &mat = context.get_factory<Material>();
    mat.add_creator(newBox<BlackCreator>());
    mat.add_creator(newBox<DiffuseCreator>());
    mat.add_creator(newBox<DisneyMaterialCreator>());
    mat.add_creator(newBox<GlassCreator>());
    mat.add_creator(newBox<InvisibleSurfaceCreator>());
    mat.add_creator(newBox<MetalCreator>());
    mat.add_creator(newBox<MirrorCreator>());
    context.get_factory<Medium>().add_creator(newBox<HomogeneousMediumCreator>());
    context.get_factory<Medium>().add_creator(newBox<HetergeneousMediumCreator>());
    auto &t2 = context.get_factory<Texture2D>();
    t2.add_creator(newBox<Array2DCreator>());
    t2.add_creator(newBox<Texture2DBinaryOperatorCreator>("add", Texture2DBinaryOperator::Type::Add));
    t2.add_creator(newBox<Texture2DBinaryOperatorCreator>("sub", Texture2DBinaryOperator::Type::Sub));
    t2.add_creator(newBox<Texture2DBinaryOperatorCreator>("mul", Texture2DBinaryOperator::Type::Mul));
    t2.add_creator(newBox<Texture2DBinaryOperatorCreator>("div", Texture2DBinaryOperator::Type::Div));
    t2.add_creator(newBox<TransformTexture2DCreator>());
    auto &t3 = context.get_factory<Texture3D>();
    t3.add_creator(newBox<Array3DCreator>());
    t3.add_creator(newBox<Texture3DBinaryOperatorCreator<BinaryOp3D::Add>>());
    t3.add_creator(newBox<Texture3DBinaryOperatorCreator<BinaryOp3D::Mul>>());
}

struct Params
{
    bool tile = false;
    int spp = 128;
    int min_depth = 5;
    int max_depth = 10;
    float rr_threshold = 0.1f;
    float rr_cont_prob = 0.5f;
    int state_count = 1000000;
    bool albedo = false;
    bool normal = false;
};

// the renderer's place in the object DAG (WavefrontPathTracer::get_dependent_objects, wavefront.cpp:197-204)
struct RendererRoot : Object
{
    RC<Scene> scene;
    RC<Camera> camera;
    RC<FilmFilter> filter;
    std::vector<RC<Object>> get_dependent_objects() override { return { scene, camera, filter }; }
};

struct Stats { uint64_t n_gen = 0, n_closest = 0, n_hit = 0, n_next = 0, n_shadow = 0, n_shadow_miss = 0, n_splat = 0; };

}  // namespace

struct refcpu_scene
{
    Box<optix::Context> optix_ctx;
    Box<factory::Context> ctx;
    RC<Scene> scene;
    RC<Camera> camera;
    RC<FilmFilter> filter;
    RC<RendererRoot> root;
    Params params;
    int width = 0, height = 0;
    bool has_medium = false;
    float world_diagonal = 0;

    Film film;
    RC<cuda::Buffer<wfpt::StateCounters>> counters;
    wfpt::GeneratePipeline generate;
    wfpt::MediumPipeline medium;
    wfpt::ShadePipeline shade;
    wfpt::PreviewImageGenerator preview;
    RC<cuda::Module> module;
    RC<cuda::Module> shadow_module;

    RC<wfpt::RayBuffer> ray, next_ray;
    RC<wfpt::PathBuffer> path, next_path;
    RC<wfpt::BSDFLeBuffer> ble, next_ble;
    RC<wfpt::IntersectionBuffer> inct;
    RC<wfpt::ShadowRayBuffer> shadow_ray;
    RC<wfpt::ShadowSamplerBuffer> shadow_sampler;
    cuda::Buffer<uint8_t> occluded;
    double jit_seconds = 0;
    std::string kernel_source;  // the generated C++ (PTXGenerator::get_llvm_ir of the shim)
};

namespace {

// shadow.cpp:38-74 with the OptiX getters replaced by the SoA loads of the same values; runs for unoccluded rays only
void record_shadow_kernel(CompileContext &cc, const Scene &scene, Film &film, float world_diagonal)
{
    using namespace cuj;
    kernel(SHADOW_KERNEL_NAME, [&](i32 shadow_ray_count, CShadowParams launch_params)
    {
        var launch_idx = cstd::block_dim_x() * cstd::block_idx_x() + cstd::thread_idx_x();
        $if(launch_idx >= shadow_ray_count)
        {
            $return();
        };
        $if(launch_params.occluded[launch_idx] != 0)
        {
            $return();
        };

        auto [ray, ray_medium_id] = launch_params.shadow_ray.load_ray(i32(launch_idx));
        auto [pixel_coord, beta] = launch_params.shadow_ray.load_beta(i32(launch_idx));

        IndependentSampler sampler({ film.width(), film.height() }, launch_params.sampler_state[launch_idx]);

        var ray_o = ray.o;
        var ray_d = ray.d;
        var ray_t1 = ray.t;

        var tr = CSpectrum::one();
        CMediumID medium_id = ray_medium_id;

        $if(ray_t1 > 1)
        {
            tr = scene.get_volume_primitive_medium()->tr(
                cc, ray_o, ray_o + normalize(ray_d) * world_diagonal, sampler);
        }
        $else
        {
            var end_pnt = ray_o + ray_d * ray_t1;
            scene.access_medium(i32(medium_id), [&](const Medium *medium)
            {
                tr = medium->tr(cc, ray_o, end_pnt, ray_o, end_pnt, sampler);
            });
        };
        beta = beta * tr;

        sampler.save(launch_params.sampler_state + launch_idx);
        film.splat_atomic(pixel_coord, Film::OUTPUT_RADIANCE, beta.to_rgb());
    });
}

// WavefrontPathTracer::commit, wavefront.cpp:98-195
void commit_renderer(refcpu_scene &s)
{
    CompileContext cc;
    auto &params = s.params;
    s.has_medium = s.scene->has_medium();

    s.film = Film(s.width, s.height);
    s.film.add_output(Film::OUTPUT_RADIANCE, Film::Float3);
    s.film.add_output(Film::OUTPUT_WEIGHT, Film::Float);
    if(params.albedo)
        s.film.add_output(Film::OUTPUT_ALBEDO, Film::Float3);
    if(params.normal)
        s.film.add_output(Film::OUTPUT_NORMAL, Film::Float3);

    s.counters = newRC<cuda::Buffer<wfpt::StateCounters>>(1);

    const auto shade_params = wfpt::ShadeParams{
        .min_depth = params.min_depth,
        .max_depth = params.max_depth,
        .rr_threshold = params.rr_threshold,
        .rr_cont_prob = params.rr_cont_prob
    };

    s.generate.set_mode(params.tile ? wfpt::GeneratePipeline::Mode::Tile : wfpt::GeneratePipeline::Mode::Uniform);

    const AABB3f world_bbox = union_aabb(s.camera->get_bounding_box(), s.scene->get_bbox());
    s.world_diagonal = 1.2f * length(world_bbox.upper - world_bbox.lower);

    const auto t0 = std::chrono::steady_clock::now();
    {
        cuj::ScopedModule cuj_module;

        s.generate.record_device_code(cc, *s.scene, *s.camera, s.film, *s.filter, params.spp);
        if(s.has_medium)
            s.medium.record_device_code(cc, s.film, *s.scene, shade_params, s.world_diagonal);
        s.shade.record_device_code(cc, s.film, *s.scene, shade_params, s.world_diagonal);

        cuj::PTXGenerator ptx_gen;
        ptx_gen.set_options(cuj::Options{
            .opt_level = cuj::OptimizationLevel::O3,
            .fast_math = true,
            .approx_math_func = true
        });
        ptx_gen.generate(cuj_module);
        s.kernel_source = ptx_gen.get_llvm_ir();

        auto &ptx = ptx_gen.get_ptx();
        auto cuda_module = newRC<cuda::Module>();
        cuda_module->load_ptx_from_memory(ptx.data(), ptx.size());
        cuda_module->link();

        s.generate.initialize(cuda_module, params.spp, params.state_count, { s.width, s.height });
        if(s.has_medium)
            s.medium.initialize(cuda_module, s.counters, *s.scene);
        s.shade.initialize(cuda_module, s.counters, *s.scene);
        s.module = cuda_module;
    }
    {
        // ShadowPipeline's constructor (shadow.cpp:98-122): own CompileContext, own module
        CompileContext shadow_cc;
        cuj::ScopedModule cuj_module;
        record_shadow_kernel(shadow_cc, *s.scene, s.film, s.world_diagonal);
        cuj::PTXGenerator ptx_gen;
        ptx_gen.generate(cuj_module);
        auto &ptx = ptx_gen.get_ptx();
        s.shadow_module = newRC<cuda::Module>();
        s.shadow_module->load_ptx_from_memory(ptx.data(), ptx.size());
        s.shadow_module->link();
    }
    s.jit_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    s.ray = newRC<wfpt::RayBuffer>(params.state_count);
    s.path = newRC<wfpt::PathBuffer>(params.state_count);
    s.ble = newRC<wfpt::BSDFLeBuffer>(params.state_count);
    s.inct = newRC<wfpt::IntersectionBuffer>(params.state_count);
    s.next_ray = newRC<wfpt::RayBuffer>(params.state_count);
    s.next_path = newRC<wfpt::PathBuffer>(params.state_count);
    s.next_ble = newRC<wfpt::BSDFLeBuffer>(params.state_count);
    s.shadow_ray = newRC<wfpt::ShadowRayBuffer>(params.state_count);
    s.shadow_sampler = newRC<wfpt::ShadowSamplerBuffer>(params.state_count);
    s.occluded.initialize(params.state_count);
}

void clear_film(refcpu_scene &s)
{
    s.film.clear_output(Film::OUTPUT_RADIANCE);
    s.film.clear_output(Film::OUTPUT_WEIGHT);
    if(s.film.has_output(Film::OUTPUT_ALBEDO))
        s.film.clear_output(Film::OUTPUT_ALBEDO);
    if(s.film.has_output(Film::OUTPUT_NORMAL))
        s.film.clear_output(Film::OUTPUT_NORMAL);
}

void trace_closest(refcpu_scene &s, int n, wfpt::RaySOA ray, wfpt::IntersectionSOA inct)
{
    refcpu::closest_hit(
        s.scene->get_tlas(), n,
        reinterpret_cast<const float *>(ray.o_med_id_buffer), reinterpret_cast<const float *>(ray.d_t1_buffer),
        inct.path_flag_buffer, reinterpret_cast<uint32_t *>(inct.t_prim_id_buffer));
}

// ShadowPipeline::test (shadow.cpp:143-158): occlusion by the lib/bvh stand-in, then the miss program for unoccluded rays
int shadow_test(refcpu_scene &s, int n)
{
    wfpt::ShadowRaySOA soa = *s.shadow_ray;
    refcpu::any_hit(
        s.scene->get_tlas(), n,
        reinterpret_cast<const float *>(soa.ray.o_med_id_buffer), reinterpret_cast<const float *>(soa.ray.d_t1_buffer),
        s.occluded.get());
    int miss = 0;
    for(int i = 0; i < n; ++i) miss += s.occluded.get()[i] ? 0 : 1;
    ShadowParams p{ soa, *s.shadow_sampler, s.occluded.get() };
    constexpr int BLOCK_DIM = 256;
    s.shadow_module->launch(SHADOW_KERNEL_NAME, { (n + BLOCK_DIM - 1) / BLOCK_DIM, 1, 1 }, { BLOCK_DIM, 1, 1 }, n, p);
    return miss;
}

// WavefrontPathTracer::render, wavefront.cpp:205-325 (no reporter, no preview)
void render(refcpu_scene &s, Stats &st)
{
    s.generate.clear();
    s.shadow_sampler->clear();
    clear_film(s);

    int active_state_count = 0;
    while(!s.generate.is_done() || active_state_count > 0)
    {
        const int new_state_count = s.generate.generate(
            active_state_count,
            wfpt::GeneratePipeline::SOAParams{ .path = *s.path, .ray = *s.ray, .bsdf_le = *s.ble },
            (std::numeric_limits<int64_t>::max)());
        active_state_count += new_state_count;
        st.n_gen += new_state_count;

        trace_closest(s, active_state_count, *s.ray, *s.inct);
        st.n_closest += active_state_count;
        {
            wfpt::IntersectionSOA in = *s.inct;
            for(int i = 0; i < active_state_count; ++i)
                st.n_hit += (in.path_flag_buffer[i] & wfpt::PATH_FLAG_HAS_INTERSECTION) ? 1 : 0;
        }

        s.counters->clear_bytes_async(0);

        if(s.has_medium)
        {
            s.medium.sample_scattering(
                active_state_count,
                wfpt::MediumPipeline::SOAParams{
                    .path = *s.path, .ray = *s.ray, .bsdf_le = *s.ble, .inct = *s.inct,
                    .output_path = *s.next_path, .output_ray = *s.next_ray, .output_bsdf_le = *s.next_ble,
                    .shadow_ray = *s.shadow_ray });
        }

        s.shade.shade(
            active_state_count,
            wfpt::ShadePipeline::SOAParams{
                .path = *s.path, .ray = *s.ray, .bsdf_le = *s.ble, .inct = *s.inct,
                .output_path = *s.next_path, .output_ray = *s.next_ray, .output_bsdf_le = *s.next_ble,
                .shadow_ray = *s.shadow_ray });

        wfpt::StateCounters state_counters;
        s.counters->to_cpu(&state_counters);
        active_state_count = state_counters.active_state_counter;
        st.n_next += state_counters.active_state_counter;
        st.n_shadow += state_counters.shadow_ray_counter;

        if(state_counters.shadow_ray_counter)
            st.n_shadow_miss += shadow_test(s, state_counters.shadow_ray_counter);

        std::swap(s.path, s.next_path);
        std::swap(s.ray, s.next_ray);
        std::swap(s.ble, s.next_ble);
    }
}

// ---- stage-level access on caller batches (uncompacted outputs, the convention of oracle/wfpt_oracle.cpp orc_shade/orc_medium)
struct StateArrays { float *o_med, *d_t1; uint32_t *pixel; float *beta_depth, *path_rad; uint64_t *rng; float *ble; };

void upload_state(refcpu_scene &s, int n, const StateArrays &a)
{
    wfpt::RaySOA r = *s.ray; wfpt::PathSOA p = *s.path; wfpt::BSDFLeSOA b = *s.ble;
    std::memcpy(r.o_med_id_buffer, a.o_med, 16ull * n);
    std::memcpy(r.d_t1_buffer, a.d_t1, 16ull * n);
    std::memcpy(p.pixel_coord_buffer, a.pixel, 8ull * n);
    std::memcpy(p.beta_depth_buffer, a.beta_depth, 16ull * n);
    std::memcpy(p.path_radiance_buffer, a.path_rad, 16ull * n);
    static_assert(sizeof(wfpt::GlobalSampler::State) == 16);
    std::memcpy(p.sampler_state_buffer, a.rng, 16ull * n);
    std::memcpy(b.beta_le_bsdf_pdf_buffer, a.ble, 16ull * n);
}

void copy_state_slot(wfpt::RaySOA r, wfpt::PathSOA p, wfpt::BSDFLeSOA b, int src, const StateArrays &a, int dst)
{
    std::memcpy(a.o_med + 4 * dst, r.o_med_id_buffer + src, 16);
    std::memcpy(a.d_t1 + 4 * dst, r.d_t1_buffer + src, 16);
    std::memcpy(a.pixel + 2 * dst, p.pixel_coord_buffer + src, 8);
    std::memcpy(a.beta_depth + 4 * dst, p.beta_depth_buffer + src, 16);
    std::memcpy(a.path_rad + 4 * dst, p.path_radiance_buffer + src, 16);
    std::memcpy(a.rng + 2 * dst, p.sampler_state_buffer + src, 16);
    std::memcpy(a.ble + 4 * dst, b.beta_le_bsdf_pdf_buffer + src, 16);
}

struct StageHook
{
    refcpu_scene *s;
    int n;
    const uint32_t *pixel;  // input pixel coordinates (2 per state)
    uint32_t *status;
    StateArrays next;
    float *sh_o_med, *sh_d_t1, *sh_beta_li, *splat, *albedo, *normal;
    wfpt::StateCounters before;
    float *film_rad, *film_alb, *film_nor;
};

void stage_hook(int i, int phase, void *user)
{
    auto &h = *static_cast<StageHook *>(user);
    if(i >= h.n) return;
    refcpu_scene &s = *h.s;
    const size_t px = (size_t)h.pixel[2 * i + 1] * s.width + h.pixel[2 * i];
    auto *c = s.counters->get();
    if(phase == 0)
    {
        h.before = *c;
        std::memset(h.film_rad + 4 * px, 0, 16);
        if(h.film_alb) std::memset(h.film_alb + 4 * px, 0, 16);
        if(h.film_nor) std::memset(h.film_nor + 4 * px, 0, 16);
        return;
    }
    uint32_t st = 0;
    if(c->active_state_counter != h.before.active_state_counter)
    {
        st |= 1;
        copy_state_slot(*s.next_ray, *s.next_path, *s.next_ble, h.before.active_state_counter, h.next, i);
    }
    if(c->shadow_ray_counter != h.before.shadow_ray_counter)
    {
        st |= 2;
        wfpt::ShadowRaySOA sr = *s.shadow_ray;
        const int k = h.before.shadow_ray_counter;
        std::memcpy(h.sh_o_med + 4 * i, sr.ray.o_med_id_buffer + k, 16);
        std::memcpy(h.sh_d_t1 + 4 * i, sr.ray.d_t1_buffer + k, 16);
        std::memcpy(h.sh_beta_li + 4 * i, sr.beta_li_buffer + k, 16);
    }
    std::memcpy(h.splat + 3 * i, h.film_rad + 4 * px, 12);
    if(h.film_alb && h.albedo) std::memcpy(h.albedo + 3 * i, h.film_alb + 4 * px, 12);
    if(h.film_nor && h.normal) std::memcpy(h.normal + 3 * i, h.film_nor + 4 * px, 12);
    h.status[i] = st;
}

float *film3(refcpu_scene &s, const char *name)
{
    return s.film.has_output(name) ? const_cast<float *>(s.film.get_float3_output(name).get()) : nullptr;
}

template<typename F>
int guarded(F &&f)
{
    try { f(); return 0; }
    catch(const std::exception &err)
    {
        std::vector<std::string> msgs;
        extract_hierarchy_exceptions(err, std::back_inserter(msgs));
        last_error.clear();
        for(auto &m : msgs) { last_error += m; last_error += '\n'; }
        return -1;
    }
}

}  // namespace

extern "C" {

const char *refcpu_last_error() { return last_error.c_str(); }

// cli/main.cpp:17-71 up to and including `dag.commit()`
refcpu_scene *refcpu_open(const char *scene_filename)
{
    refcpu_scene *s = new refcpu_scene();
    const int rc = guarded([&]
    {
        s->optix_ctx = newBox<optix::Context>(nullptr);
        s->ctx = newBox<factory::Context>(*s->optix_ctx);
        register_creators(*s->ctx);

        const auto scene_dir = std::filesystem::path(scene_filename).parent_path();
        s->ctx->add_path_mapping("scene_directory", scene_dir.string());

        factory::JSONParser parser;
        std::string json_source = read_txt_file(scene_filename);
        parser.set_source(std::move(json_source));
        parser.add_include_directory(scene_dir);
        parser.parse();
        auto root_node = parser.get_result();

        s->scene = create_scene(root_node->child_node("scene"), *s->ctx);

        s->width = root_node->parse_child<int>("width");
        s->height = root_node->parse_child<int>("height");
        s->camera = s->ctx->create<Camera>(root_node->child_node("camera"));
        s->camera->set_w_over_h(static_cast<float>(s->width) / s->height);

        // WavefrontPathTracerCreator::create, wavefront.cpp:375-398
        auto node = root_node->child_node("renderer");
        auto &params = s->params;
        params.tile         = node->parse_child_or("tile", params.tile);
        params.spp          = node->parse_child_or("spp", params.spp);
        params.min_depth    = node->parse_child_or("min_depth", params.min_depth);
        params.max_depth    = node->parse_child_or("max_depth", params.max_depth);
        params.rr_threshold = node->parse_child_or("rr_threshold", params.rr_threshold);
        params.rr_cont_prob = node->parse_child_or("rr_cont_prob", params.rr_cont_prob);
        params.state_count  = node->parse_child_or("state_count", params.state_count);
        params.albedo       = node->parse_child_or("albedo", params.albedo);
        params.normal       = node->parse_child_or("normal", params.normal);
        if(auto n = node->find_child_node("filter"))
            s->filter = s->ctx->create<FilmFilter>(n);
        else
            s->filter = newRC<BoxFilter>();

        s->root = newRC<RendererRoot>();
        s->root->scene = s->scene; s->root->camera = s->camera; s->root->filter = s->filter;
        ObjectDAG dag(s->root);
        dag.commit();
        commit_renderer(*s);
    });
    if(rc != 0) { delete s; return nullptr; }
    return s;
}

void refcpu_close(refcpu_scene *s) { delete s; }

// out[0..8] = width, height, spp, state_count, has_medium, triangle count, instance count, volume-primitive medium id, medium count
void refcpu_info(refcpu_scene *s, int64_t *out)
{
    out[0] = s->width; out[1] = s->height; out[2] = s->params.spp; out[3] = s->params.state_count; out[4] = s->has_medium;
    out[5] = (int64_t)refcpu::world_triangles(s->scene->get_tlas()).inst_prim.size() / 2;
    out[6] = s->scene->get_instance_count();
    out[7] = (int64_t)s->scene->get_volume_primitive_medium_id();
    out[8] = s->scene->get_medium_count();
}
double refcpu_jit_seconds(refcpu_scene *s) { return s->jit_seconds; }
float refcpu_world_diagonal(refcpu_scene *s) { return s->world_diagonal; }
size_t refcpu_kernel_source(refcpu_scene *s, char *buf, size_t cap)
{
    if(buf && cap) { const size_t n = std::min(cap - 1, s->kernel_source.size()); std::memcpy(buf, s->kernel_source.data(), n); buf[n] = 0; }
    return s->kernel_source.size();
}

void refcpu_get_triangles(refcpu_scene *s, float *verts, uint32_t *inst_prim)
{
    auto &w = refcpu::world_triangles(s->scene->get_tlas());
    std::memcpy(verts, w.verts.data(), w.verts.size() * 4);
    std::memcpy(inst_prim, w.inst_prim.data(), w.inst_prim.size() * 4);
}

// the reference's per-triangle shading tables (core/geometry.h:10-19) of instance `inst`'s geometry: 6 float4 per triangle
int refcpu_get_geometry_tables(refcpu_scene *s, int inst, int ntri, float *out24)
{
    const InstanceInfo &ii = s->scene->get_host_instance_info()[inst];
    const GeometryInfo &g = s->scene->get_host_geometry_info()[ii.geometry_id];
    const Vec4f *src[6] = { g.geometry_ex_tex_coord_u_a, g.geometry_ey_tex_coord_u_ba, g.geometry_ez_tex_coord_u_ca,
                            g.shading_normal_tex_coord_v_a, g.shading_normal_tex_coord_v_ba, g.shading_normal_tex_coord_v_ca };
    for(int t = 0; t < ntri; ++t)
        for(int k = 0; k < 6; ++k)
            std::memcpy(out24 + 24 * t + 4 * k, src[k] + t, 16);
    return 0;
}

void refcpu_closest_hit(refcpu_scene *s, int n, const float *o_med, const float *d_t1, uint32_t *flag, uint32_t *detail)
{
    refcpu::closest_hit(s->scene->get_tlas(), n, o_med, d_t1, flag, detail);
}

// the reference's generate kernel for path indices [first, first + n) (generate.cpp:39-86; launch as in :140-177)
int refcpu_generate(refcpu_scene *s, int64_t first, int n, float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth,
                    float *path_rad, uint64_t *rng, float *ble)
{
    return guarded([&]
    {
        if(n > s->params.state_count) throw BtrcException("batch larger than state_count");
        clear_film(*s);
        constexpr int BLOCK_DIM = 256;
        wfpt::GeneratePipeline::SOAParams soa{ .path = *s->path, .ray = *s->ray, .bsdf_le = *s->ble };
        s->module->launch("generate_kernel", { (n + BLOCK_DIM - 1) / BLOCK_DIM, 1, 1 }, { BLOCK_DIM, 1, 1 }, soa, first, n, 0);
        StateArrays a{ o_med, d_t1, pixel, beta_depth, path_rad, rng, ble };
        for(int i = 0; i < n; ++i) copy_state_slot(*s->ray, *s->path, *s->ble, i, a, i);
    });
}

static int run_stage(refcpu_scene *s, bool medium, int n, StateArrays in, uint32_t *flag, const uint32_t *detail, uint32_t *status,
                     StateArrays next, float *sh_o_med, float *sh_d_t1, float *sh_beta_li, float *splat, float *albedo, float *normal)
{
    return guarded([&]
    {
        if(n > s->params.state_count) throw BtrcException("batch larger than state_count");
        upload_state(*s, n, in);
        wfpt::IntersectionSOA ic = *s->inct;
        std::memcpy(ic.path_flag_buffer, flag, 4ull * n);
        std::memcpy(ic.t_prim_id_buffer, detail, 16ull * n);
        clear_film(*s);
        s->counters->clear_bytes(0);
        StageHook h{ s, n, in.pixel, status, next, sh_o_med, sh_d_t1, sh_beta_li, splat, albedo, normal, {},
                     film3(*s, Film::OUTPUT_RADIANCE), film3(*s, Film::OUTPUT_ALBEDO), film3(*s, Film::OUTPUT_NORMAL) };
        refcpu::set_launch_mode({ true, &stage_hook, &h });
        BTRC_SCOPE_EXIT{ refcpu::set_launch_mode({}); };
        if(medium)
        {
            s->medium.sample_scattering(n, wfpt::MediumPipeline::SOAParams{
                .path = *s->path, .ray = *s->ray, .bsdf_le = *s->ble, .inct = *s->inct,
                .output_path = *s->next_path, .output_ray = *s->next_ray, .output_bsdf_le = *s->next_ble, .shadow_ray = *s->shadow_ray });
            // what the kernel wrote back to its own slots (medium.cpp:121-124,160-163)
            wfpt::RaySOA r = *s->ray; wfpt::PathSOA p = *s->path; wfpt::BSDFLeSOA b = *s->ble;
            for(int i = 0; i < n; ++i)
            {
                in.o_med[4 * i + 3] = r.o_med_id_buffer[i].w;
                std::memcpy(in.ble + 4 * i, b.beta_le_bsdf_pdf_buffer + i, 16);
                std::memcpy(in.rng + 2 * i, p.sampler_state_buffer + i, 16);
                flag[i] = ic.path_flag_buffer[i];
            }
        }
        else
        {
            s->shade.shade(n, wfpt::ShadePipeline::SOAParams{
                .path = *s->path, .ray = *s->ray, .bsdf_le = *s->ble, .inct = *s->inct,
                .output_path = *s->next_path, .output_ray = *s->next_ray, .output_bsdf_le = *s->next_ble, .shadow_ray = *s->shadow_ray });
        }
    });
}

// signatures = oracle/wfpt_oracle.cpp orc_shade / orc_medium (status bit 4 is not observable from the reference kernel: a
// radiance splat of zero leaves no trace; `splat` is the film's change for the state's pixel, zero when nothing was splatted)
int refcpu_shade(refcpu_scene *s, int n,
                 float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth, float *path_rad, uint64_t *rng, float *ble,
                 uint32_t *flag, const uint32_t *detail, uint32_t *status,
                 float *n_o_med, float *n_d_t1, uint32_t *n_pixel, float *n_beta_depth, float *n_path_rad, uint64_t *n_rng, float *n_ble,
                 float *sh_o_med, float *sh_d_t1, float *sh_beta_li, float *splat_rgb, float *gbuf_albedo, float *gbuf_normal)
{
    return run_stage(s, false, n, { o_med, d_t1, pixel, beta_depth, path_rad, rng, ble }, flag, detail, status,
                     { n_o_med, n_d_t1, n_pixel, n_beta_depth, n_path_rad, n_rng, n_ble }, sh_o_med, sh_d_t1, sh_beta_li, splat_rgb,
                     gbuf_albedo, gbuf_normal);
}

int refcpu_medium(refcpu_scene *s, int n,
                  float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth, float *path_rad, uint64_t *rng, float *ble,
                  uint32_t *flag, const uint32_t *detail, uint32_t *status,
                  float *n_o_med, float *n_d_t1, uint32_t *n_pixel, float *n_beta_depth, float *n_path_rad, uint64_t *n_rng, float *n_ble,
                  float *sh_o_med, float *sh_d_t1, float *sh_beta_li, float *splat_rgb)
{
    if(!s->has_medium) { last_error = "scene has no medium"; return -1; }
    return run_stage(s, true, n, { o_med, d_t1, pixel, beta_depth, path_rad, rng, ble }, flag, detail, status,
                     { n_o_med, n_d_t1, n_pixel, n_beta_depth, n_path_rad, n_rng, n_ble }, sh_o_med, sh_d_t1, sh_beta_li, splat_rgb,
                     nullptr, nullptr);
}

// shadow stage on a caller batch: contribution[i] = what the film received for ray i (0 when occluded).  The per-slot
// sampler states are the reference's (ShadowSamplerBuffer::clear, soa_buffer.cpp:86-96) after `clear`.
int refcpu_shadow(refcpu_scene *s, int n, const float *sh_o_med, const float *sh_d_t1, const float *sh_beta_li, const uint32_t *pixel,
                  uint8_t *occluded, float *contribution)
{
    return guarded([&]
    {
        if(n > s->params.state_count) throw BtrcException("batch larger than state_count");
        wfpt::ShadowRaySOA sr = *s->shadow_ray;
        std::memcpy(sr.ray.o_med_id_buffer, sh_o_med, 16ull * n);
        std::memcpy(sr.ray.d_t1_buffer, sh_d_t1, 16ull * n);
        std::memcpy(sr.beta_li_buffer, sh_beta_li, 16ull * n);
        std::memcpy(sr.pixel_coord_buffer, pixel, 8ull * n);
        s->shadow_sampler->clear();
        clear_film(*s);
        std::vector<uint32_t> status(n);
        StageHook h{ s, n, pixel, status.data(), {}, nullptr, nullptr, nullptr, contribution, nullptr, nullptr, {},
                     film3(*s, Film::OUTPUT_RADIANCE), nullptr, nullptr };
        refcpu::set_launch_mode({ true, &stage_hook, &h });
        BTRC_SCOPE_EXIT{ refcpu::set_launch_mode({}); };
        shadow_test(*s, n);
        if(occluded) std::memcpy(occluded, s->occluded.get(), n);
    });
}

// full render = WavefrontPathTracer::render + the preview resolve (preview.cpp:28-49).  color/albedo/normal: h*w*4 floats or null;
// raw_radiance h*w*4, raw_weight h*w (the film before the resolve) or null; stats: 7 counters as in orc_render
int refcpu_render(refcpu_scene *s, float *color, float *albedo, float *normal, float *raw_radiance, float *raw_weight, uint64_t *stats,
                  double *seconds)
{
    return guarded([&]
    {
        Stats st;
        const auto t0 = std::chrono::steady_clock::now();
        render(*s, st);
        const size_t texels = (size_t)s->width * s->height;
        auto rad = s->film.get_float3_output(Film::OUTPUT_RADIANCE).as<Vec4f>();
        auto wgt = s->film.get_float_output(Film::OUTPUT_WEIGHT).get();
        cuda::Buffer<Vec4f> out(texels);
        if(color)
        {
            s->preview.generate(s->width, s->height, rad, wgt, out.get());
            std::memcpy(color, out.get(), texels * 16);
        }
        if(seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if(albedo && s->params.albedo)
        {
            s->preview.generate_albedo(s->width, s->height, s->film.get_float3_output(Film::OUTPUT_ALBEDO).as<Vec4f>(), wgt, out.get());
            std::memcpy(albedo, out.get(), texels * 16);
        }
        if(normal && s->params.normal)
        {
            s->preview.generate_normal(s->width, s->height, s->film.get_float3_output(Film::OUTPUT_NORMAL).as<Vec4f>(), wgt, out.get());
            std::memcpy(normal, out.get(), texels * 16);
        }
        if(raw_radiance) std::memcpy(raw_radiance, rad, texels * 16);
        if(raw_weight) std::memcpy(raw_weight, wgt, texels * 4);
        if(stats)
        {
            stats[0] = st.n_gen; stats[1] = st.n_closest; stats[2] = st.n_hit; stats[3] = st.n_next;
            stats[4] = st.n_shadow; stats[5] = st.n_shadow_miss; stats[6] = st.n_splat;
        }
    });
}

void refcpu_set_spp(refcpu_scene *s, int spp)
{
    // GeneratePipeline keeps spp for its cursor; re-initialise it (generate.cpp:90-104 asserts a fresh object)
    s->params.spp = spp;
    s->generate = wfpt::GeneratePipeline();
    s->generate.set_mode(s->params.tile ? wfpt::GeneratePipeline::Mode::Tile : wfpt::GeneratePipeline::Mode::Uniform);
    s->generate.initialize(s->module, spp, s->params.state_count, { s->width, s->height });
}

}  // extern "C"
