// Wavefront path-tracing kernels for sm_100a (see kernels.h for the reference stage each one replaces).
//
// Execution model: every stage is a persistent, fixed-size grid (a multiple of the SM count) that reads its work
// count from the device-resident WaveCtl block, so the host can enqueue whole iterations back to back without the
// per-iteration StateCounters read-back of the reference (wavefront.cpp:283-284).  Queues are SoA float4 arrays:
// each lane issues 16-byte coalesced loads/stores.  Output slots are allocated with one atomic per warp
// (ballot + popc prefix) instead of one per thread (shade.cpp:165,225).
#include "kernels.h"

#include "dev_math.cuh"
#include "rng.cuh"
#include "traverse.cuh"

namespace btrc {

namespace {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Warp-aggregated slot allocation: returns this lane's slot (valid only where pred is true).
__device__ __forceinline__ int warp_alloc(int *counter, bool pred) {
    const unsigned m = __ballot_sync(kFull, pred);
    if (m == 0) return -1;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(kFull, base, leader);
    return base + __popc(m & ((1u << lane_id()) - 1u));
}

__device__ __forceinline__ void warp_count(unsigned long long *counter, bool pred) {
    const unsigned m = __ballot_sync(kFull, pred);
    if (m && lane_id() == __ffs(m) - 1) atomicAdd(counter, (unsigned long long)__popc(m));
}

// core/film.cpp:131-165: bounds are guaranteed by construction, the isfinite test is kept
__device__ __forceinline__ void film_splat3(float *buf, uint32_t pixel, f3 v) {
    if (isfinite3(v)) {
        float *p = buf + 4 * (size_t)pixel;
        atomicAdd(p + 0, v.x);
        atomicAdd(p + 1, v.y);
        atomicAdd(p + 2, v.z);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// prepare: device-side bookkeeping of one iteration (wavefront.cpp:227-306, generate.cpp:140-177)
// ---------------------------------------------------------------------------------------------------------------
__global__ void prepare_kernel(WaveCtl *ctl) {
    const int survivors = ctl->n_next;
    ctl->n_next_total += (unsigned long long)survivors;
    ctl->n_shadow_total += (unsigned long long)ctl->n_shadow;
    const long long remaining = ctl->end - ctl->cursor;
    const long long room = (long long)ctl->pool - survivors;
    const int n_new = (int)(remaining < room ? remaining : room);
    ctl->gen_base = ctl->cursor;
    ctl->cursor += n_new;
    ctl->n_survivors = survivors;
    ctl->n_new = n_new;
    ctl->n_active = survivors + n_new;
    ctl->n_next = 0;
    ctl->n_shadow = 0;
    ctl->trace_cursor = 0;
    ctl->shadow_cursor = 0;
    ctl->n_gen += (unsigned long long)n_new;
    ctl->n_closest += (unsigned long long)(survivors + n_new);
    ctl->n_iter += (survivors + n_new) > 0 ? 1 : 0;
    ctl->done = (survivors + n_new) == 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// generate (generate.cpp:39-86, film_filter/box.cpp:5-10, gaussian.cpp:54-66, camera/pinhole.cpp:84-99)
// ---------------------------------------------------------------------------------------------------------------
struct GenOut {
    f3 o, d;
    uint32_t px, py;
    Sampler sampler;
};

__device__ __forceinline__ GenOut generate_one(const SceneDevice &scene, const RenderParams &P, long long accu) {
    GenOut g;
    const int W = P.width, H = P.height;
    int pixel_index, sample_index;
    if (!P.tile) {
        pixel_index = int(accu % (long long)(W * H));
        sample_index = int(accu / (long long)(W * H));
    } else {
        pixel_index = int(accu / P.spp);
        sample_index = int(accu % P.spp);
    }
    const int pixel_x = pixel_index % W, pixel_y = pixel_index / W;
    g.px = pixel_x; g.py = pixel_y;
    g.sampler.init((uint32_t)pixel_x, (uint32_t)pixel_y, sample_index);
    float fx, fy;
    if (P.filter_type == 0) {
        fx = g.sampler.get1d() - 0.5f;
        fy = g.sampler.get1d() - 0.5f;
    } else {
        constexpr int GRID = 256;
        const uint32_t gi = alias_sample(scene.filter_alias, GRID * GRID, g.sampler.get1d());
        const uint32_t gy = gi / GRID, gx = gi % GRID;
        const float r = P.filter_radius, step = P.filter_step;
        const float x_beg = __fadd_rn(-r, __fmul_rn(step, float(gx))), x_end = __fadd_rn(-r, __fmul_rn(step, float(gx + 1)));
        const float y_beg = __fadd_rn(-r, __fmul_rn(step, float(gy))), y_end = __fadd_rn(-r, __fmul_rn(step, float(gy + 1)));
        fx = lerpf(x_beg, x_end, g.sampler.get1d());
        fy = lerpf(y_beg, y_end, g.sampler.get1d());
    }
    const float pixel_xf = float(pixel_x) + 0.5f + fx;
    const float pixel_yf = float(pixel_y) + 0.5f + fy;
    const float film_x = pixel_xf / float(W);
    const float film_y = pixel_yf / float(H);
    (void)g.sampler.get1d();  // time_sample: drawn, unused by the pinhole camera (generate.cpp:76)
    const CameraInfo &c = scene.camera;
    const f3 x = film_x * f3(c.film_x[0], c.film_x[1], c.film_x[2]);
    const f3 y = film_y * f3(c.film_y[0], c.film_y[1], c.film_y[2]);
    const f3 dst = f3(c.left_bottom_corner[0], c.left_bottom_corner[1], c.left_bottom_corner[2]) + x + y;
    g.o = f3(c.eye[0], c.eye[1], c.eye[2]);
    g.d = normalize(dst - g.o);
    return g;
}

__global__ void __launch_bounds__(kBlock) generate_kernel(SceneDevice scene, RenderParams P, PathQueue q, FilmDevice film, const WaveCtl *ctl) {
    const int n_new = ctl->n_new;
    const int base_slot = ctl->n_survivors;
    const long long gen_base = ctl->gen_base;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x) {
        GenOut g = generate_one(scene, P, gen_base + i);
        const uint32_t pixel = g.py * (uint32_t)P.width + g.px;
        atomicAdd(film.weight + pixel, 1.0f);
        const int s = base_slot + i;
        q.ray_o[s] = make4(g.o, __uint_as_float(scene.vol_prim_medium_id));
        q.ray_d[s] = make4(g.d, kMaxFloat);
        q.beta[s] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
        q.rad[s] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(pixel));
        q.rng[s] = make_ulonglong2(g.sampler.rng.state, g.sampler.rng.inc);
        q.ble[s] = make_float4(1.0f, 1.0f, 1.0f, -1.0f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// closest hit (trace.cpp:38-78)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_hit(const BvhDevice &bvh, const HitRec &h, int i, uint32_t *flag, uint4 *detail) {
    if (h.slot < 0) {
        flag[i] = 0;  // miss: flag only (save_flag(false, false, 0))
    } else {
        const uint2 id = __ldg(bvh.tri_ids + h.slot);
        flag[i] = id.x | kFlagHasIntersection;
        detail[i] = make_uint4(__float_as_uint(h.t), id.y, __float_as_uint(h.u), __float_as_uint(h.v));
    }
}

__global__ void __launch_bounds__(kBlock) trace_kernel(SceneDevice scene, PathQueue q, HitQueue hq, WaveCtl *ctl) {
    const int n = ctl->n_active;
    unsigned long long hits = 0;
    for (;;) {
        int base = 0;
        if (lane_id() == 0) base = atomicAdd(&ctl->trace_cursor, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const int i = base + lane_id();
        if (i < n) {
            const float4 o4 = q.ray_o[i], d4 = q.ray_d[i];
            const HitRec h = traverse_bvh2<false>(scene.bvh, xyz(o4), xyz(d4), d4.w);
            store_hit(scene.bvh, h, i, hq.flag, hq.detail);
            hits += h.slot >= 0 ? 1 : 0;
        }
        __syncwarp();
    }
    for (int off = 16; off; off >>= 1) hits += __shfl_down_sync(kFull, hits, off);
    if (lane_id() == 0 && hits) atomicAdd(&ctl->n_hit, hits);
}

__global__ void __launch_bounds__(kBlock) trace_batch_kernel(SceneDevice scene, const float4 *ray_o, const float4 *ray_d, int n,
                                                            uint32_t *flag, uint4 *detail) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ray_o[i], d4 = ray_d[i];
        const HitRec h = traverse_bvh2<false>(scene.bvh, xyz(o4), xyz(d4), d4.w);
        if (h.slot < 0) detail[i] = make_uint4(0, 0, 0, 0);
        store_hit(scene.bvh, h, i, flag, detail);
    }
}

__global__ void __launch_bounds__(kBlock) anyhit_batch_kernel(SceneDevice scene, const float4 *ray_o, const float4 *ray_d, int n, uint8_t *occluded) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ray_o[i], d4 = ray_d[i];
        occluded[i] = traverse_bvh2<true>(scene.bvh, xyz(o4), xyz(d4), d4.w).slot >= 0 ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shade (shade.cpp:38-235, helper.cpp:10-207, core/scene.cpp:9-64, material/*)
// ---------------------------------------------------------------------------------------------------------------
struct SurfacePoint {
    f3 position;
    Frame3 frame;
    f3 interp_z;
    float2 uv, tex_coord;
};

// core/scene.cpp:9-64
__device__ __forceinline__ SurfacePoint get_hitinfo(f3 o, f3 d, const InstanceInfo &inst, const GeometryInfo &geo, float t, uint32_t prim, float2 uv) {
    SurfacePoint sp;
    sp.position = o + t * d;
    const float4 gx = ldg4(geo.gx_ua + prim), gy = ldg4(geo.gy_uba + prim), gz = ldg4(geo.gz_uca + prim);
    const float4 sa = ldg4(geo.sn_va + prim), sba = ldg4(geo.sn_vba + prim), sca = ldg4(geo.sn_vca + prim);
    Frame3 gf(xyz(gx), xyz(gy), xyz(gz));
    gf.x = normalize(xf_vector(inst.mat, gf.x));
    gf.y = normalize(xf_vector(inst.mat, gf.y));
    gf.z = normalize(xf_normal(inst.inv, gf.z));
    f3 interp = xyz(sa) + xyz(sba) * uv.x + xyz(sca) * uv.y;
    interp = normalize(xf_normal(inst.inv, interp));
    sp.frame = gf;
    sp.interp_z = interp;
    sp.uv = uv;
    sp.tex_coord = make_float2(gx.w + gy.w * uv.x + gz.w * uv.y, sa.w + sba.w * uv.x + sca.w * uv.y);
    return sp;
}

__device__ __forceinline__ f3 eval_texture(const SceneDevice &scene, int tex, const SurfacePoint &) {
    const TextureDesc &t = scene.textures[tex];
    return f3(t.value[0], t.value[1], t.value[2]);
}

// helper.cpp:174-207 + mesh_light.cpp:17-29,60-76
__device__ __forceinline__ f3 handle_intersected_light(const SceneDevice &scene, f3 o, f3 d, const SurfacePoint &inct, f3 beta_le, float bsdf_pdf, int light_index) {
    f3 result(0.0f);
    if (light_index >= 0) {
        const float select_pdf = 1.0f / float(scene.n_lights);
        const LightInfo &li = scene.lights[light_index];
        const f3 intensity(li.intensity[0], li.intensity[1], li.intensity[2]);
        const f3 le = dot(inct.frame.z, -d) >= 0.0f ? intensity : f3(0.0f);
        if (bsdf_pdf < 0.0f) {
            result = beta_le * le / -bsdf_pdf;
        } else {
            float light_dir_pdf;
            const f3 nor = inct.frame.z, pos = inct.position;
            if (dot(nor, o - pos) <= 0.0f) light_dir_pdf = 0.0f;
            else {
                const GeometryInfo &geo = scene.geometries[scene.instances[li.instance_id].geometry_id];
                const float pdf_area = 1.0f / geo.total_area;
                const f3 spt_to_ref = o - pos;
                const float dist2 = length_square(spt_to_ref);
                const float dist3 = sqrtf(dist2) * dist2;
                light_dir_pdf = li.inv_scale2 * pdf_area * dist3 / fabsf(dot(nor, spt_to_ref));
            }
            const float light_pdf = select_pdf * light_dir_pdf;
            result = beta_le * le / (bsdf_pdf + light_pdf);
        }
    }
    return result;
}

struct LiSample {
    bool success;
    f3 o, d, li;
    float t1, light_pdf;
};

// helper.cpp:10-65,125-142; light_sampler.cpp:41-63; mesh_light.cpp:31-58; triangle_mesh.cpp:188-228
__device__ __forceinline__ LiSample sample_surface_light_li(const SceneDevice &scene, f3 ref_pos, f3 ref_nor, Sampler &sampler) {
    LiSample r;
    r.success = false;
    const float s0 = sampler.get1d(), s1 = sampler.get1d(), s2 = sampler.get1d();
    const float sel = sampler.get1d();
    const int n = scene.n_lights;
    if (n == 0) return r;
    const int idx = min(int(sel * n), n - 1);
    const float select_pdf = 1.0f / float(n);
    const LightInfo &li = scene.lights[idx];
    const InstanceInfo &linst = scene.instances[li.instance_id];
    const GeometryInfo &geo = scene.geometries[linst.geometry_id];
    const uint32_t prim = alias_sample(geo.area_alias, (uint32_t)geo.prim_count, s0);
    const float2 uv = sample_triangle_uniform(s1, s2);
    const float *pp = geo.positions + 9 * (size_t)prim;
    const f3 a(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)), ba(__ldg(pp + 3), __ldg(pp + 4), __ldg(pp + 5)), ca(__ldg(pp + 6), __ldg(pp + 7), __ldg(pp + 8));
    const f3 pos = a + ba * uv.x + ca * uv.y;
    const f3 ez = normalize(xyz(ldg4(geo.gz_uca + prim)));  // CFrame(ex, ey, ez).z
    const float pdf_area = 1.0f / geo.total_area;
    const f3 spos = xf_point(linst.mat, pos);
    const f3 snor = normalize(xf_normal(linst.inv, ez));
    const f3 pos_to_ref = ref_pos - spos;
    const float dist2 = length_square(pos_to_ref);
    const float pdf_sa = li.inv_scale2 * pdf_area * dist2 / fabsf(dot(snor, normalize(pos_to_ref)));
    const f3 rad = dot(pos_to_ref, snor) > 0.0f ? f3(li.intensity[0], li.intensity[1], li.intensity[2]) : f3(0.0f);
    const f3 diff = spos - ref_pos;
    const f3 shadow_dst = intersection_offset(spos, snor, -diff);
    r.o = intersection_offset(ref_pos, ref_nor, diff);
    r.d = shadow_dst - r.o;
    r.t1 = 1.0f;
    r.li = rad;
    r.light_pdf = select_pdf * pdf_sa;
    r.success = true;
    return r;
}

// helper.cpp:67-95
__device__ __forceinline__ bool simple_russian_roulette(const RenderParams &P, f3 &beta, int depth, Sampler &sampler) {
    bool ret = false;
    if (depth >= P.min_depth) {
        if (depth >= P.max_depth) ret = true;
        else {
            const float sam = sampler.get1d();
            if (lum(beta) < P.rr_threshold) {
                if (sam > P.rr_cont_prob) ret = true;
                else beta = beta / P.rr_cont_prob;
            }
        }
    }
    return ret;
}

// material/utils/shader_frame.cpp:5-95
struct ShaderFrame {
    Frame3 geometry, shading;
    __device__ __forceinline__ bool is_black_fringes(f3 w) const {
        return geometry.in_positive_z_hemisphere(w) != shading.in_positive_z_hemisphere(w);
    }
    __device__ __forceinline__ float correct_shading_energy(f3 wi) const {
        const float down = fabsf(dot(geometry.z, wi));
        return down > 1e-4f ? fabsf(dot(shading.z, wi)) / down : 1.0f;
    }
};

struct BsdfSample {
    f3 bsdf, dir;
    float pdf;
    bool is_delta;
};

// BSDFAggregate with the closures of one material (material/utils/aggregate.cpp:32-230).  The per-material
// closure sets are compile-time specialisations selected by MaterialDesc::type -- this replaces the reference's
// per-object run-time code generation (core/context.h:321-382).
struct Shader {
    int type;
    ShaderFrame frame;
    bool shadow_terminator;
    f3 albedo_value;

    __device__ __forceinline__ float shadow_terminator_term(f3 wi) const {  // aggregate.cpp:219-230
        if (!shadow_terminator) return 1.0f;
        const float wgwi = dot(frame.geometry.z, wi);
        const float wswi = dot(frame.shading.z, wi);
        const float wgws = dot(frame.geometry.z, frame.shading.z);
        const float g = saturate(wgwi / (wswi * wgws));
        return -g * g * g + g * g + g;
    }
    __device__ __forceinline__ f3 albedo() const { return type == MAT_BLACK ? f3(0.0f) : albedo_value; }
    __device__ __forceinline__ f3 normal() const { return frame.shading.z; }

    __device__ __forceinline__ Frame3 flipped_shading(f3 wo) const {  // flip_for_black_fringes
        return frame.is_black_fringes(wo) ? frame.geometry : frame.shading;
    }

    __device__ __forceinline__ BsdfSample sample(f3 wo, float s0, float s1, float s2) const {
        BsdfSample r;
        r.bsdf = f3(0.0f); r.dir = f3(0.0f); r.pdf = 0.0f; r.is_delta = false;
        if (type == MAT_BLACK) return r;
        const Frame3 sh = flipped_shading(wo);
        const f3 lwo = normalize(sh.global_to_local(wo));
        // diffuse closure (material/component/diffuse.cpp:10-30); single closure: selector = sam[0]
        if (lwo.z <= 0.0f) return r;
        const f3 lwi = sample_hemisphere_zweighted(s0, s1);
        const float cpdf = pdf_sample_hemisphere_zweighted(lwi);
        if (cpdf <= 0.0f) return r;
        const f3 dir = sh.local_to_global(lwi);
        if (frame.is_black_fringes(dir)) return r;
        ShaderFrame fl; fl.geometry = frame.geometry; fl.shading = sh;
        const float corr = fl.correct_shading_energy(dir);
        const float stt = shadow_terminator_term(dir);
        r.dir = dir;
        r.bsdf = albedo_value / kPi * corr * stt;
        r.pdf = cpdf;
        return r;
    }
    __device__ __forceinline__ f3 eval(f3 wi, f3 wo) const {  // aggregate.cpp:185-197
        if (type == MAT_BLACK) return f3(0.0f);
        const Frame3 sh = flipped_shading(wo);
        const f3 lwi = normalize(sh.global_to_local(wi));
        const f3 lwo = normalize(sh.global_to_local(wo));
        f3 result(0.0f);
        if (!(lwi.z <= 0.0f || lwo.z <= 0.0f)) result = result + albedo_value / kPi;
        ShaderFrame fl; fl.geometry = frame.geometry; fl.shading = sh;
        return result * fl.correct_shading_energy(wi) * shadow_terminator_term(wi);
    }
    __device__ __forceinline__ float pdf(f3 wi, f3 wo) const {  // aggregate.cpp:199-209
        if (type == MAT_BLACK) return 0.0f;
        const Frame3 sh = flipped_shading(wo);
        const f3 lwi = normalize(sh.global_to_local(wi));
        const f3 lwo = normalize(sh.global_to_local(wo));
        return (lwi.z <= 0.0f || lwo.z <= 0.0f) ? 0.0f : pdf_sample_hemisphere_zweighted(lwi);
    }
};

__device__ __forceinline__ Shader create_shader(const SceneDevice &scene, const MaterialDesc &mat, const SurfacePoint &inct) {
    Shader sh;
    sh.type = mat.type;
    sh.frame.geometry = inct.frame;
    sh.frame.shading = inct.frame.rotate_to_new_z(inct.interp_z);
    if (mat.normal_tex >= 0) {  // material/utils/normal_map.cpp:13-21
        const f3 c = eval_texture(scene, mat.normal_tex, inct);
        const f3 local_nor = 2.0f * c - f3(1.0f);
        const f3 world_nor = sh.frame.shading.local_to_global(local_nor);
        sh.frame.shading = sh.frame.shading.rotate_to_new_z(normalize(world_nor));
    }
    sh.shadow_terminator = mat.shadow_terminator != 0;
    sh.albedo_value = mat.type == MAT_DIFFUSE ? eval_texture(scene, mat.color_tex, inct) : f3(0.0f);
    return sh;
}

__global__ void __launch_bounds__(kBlock) shade_kernel(SceneDevice scene, RenderParams P, PathQueue q, HitQueue hq, PathQueue nq, ShadowQueue sq,
                                                      FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->n_active;
    const int n_round = (n + 31) & ~31;  // whole warps take part in the aggregated allocations
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool emit_next = false, emit_shadow = false, splat = false;
        f3 splat_value(0.0f);
        uint32_t pixel = 0;
        // next state
        f3 n_o, n_d, n_beta, n_L, n_ble;
        float n_pdf = 0.0f;
        uint32_t n_med = 0;
        int n_depth = 0;
        ulonglong2 n_rng = make_ulonglong2(0, 0);
        // shadow ray
        f3 s_o, s_d, s_beta;
        float s_t1 = 0.0f;
        uint32_t s_med = 0;

        if (i < n) {
            const uint32_t flag = hq.flag[i];
            const float4 beta4 = q.beta[i], rad4 = q.rad[i], o4 = q.ray_o[i], d4 = q.ray_d[i], ble4 = q.ble[i];
            const ulonglong2 rng2 = q.rng[i];
            pixel = __float_as_uint(rad4.w);
            f3 beta = xyz(beta4), L = xyz(rad4);
            const int depth = __float_as_int(beta4.w);
            const f3 o = xyz(o4), d = xyz(d4);
            const uint32_t ray_medium = __float_as_uint(o4.w);
            const f3 beta_le = xyz(ble4);
            const float bsdf_pdf = ble4.w;
            Sampler sampler;
            sampler.rng.state = rng2.x; sampler.rng.inc = rng2.y;

            if (!(flag & kFlagHasIntersection)) {
                // miss: no environment light yet => Le = 0; (shade.cpp:54-69)
                splat = true; splat_value = L;
            } else {
                const uint4 det = hq.detail[i];
                const InstanceInfo &inst = scene.instances[flag & kFlagInstanceMask];
                const GeometryInfo &geo = scene.geometries[inst.geometry_id];
                const SurfacePoint inct = get_hitinfo(o, d, inst, geo, __uint_as_float(det.x), det.y,
                                                      make_float2(__uint_as_float(det.z), __uint_as_float(det.w)));
                const f3 le_rad = handle_intersected_light(scene, o, d, inct, beta_le, bsdf_pdf, inst.light_id);
                L = L + le_rad;
                if (simple_russian_roulette(P, beta, depth, sampler)) {
                    splat = true; splat_value = L;
                } else {
                    LiSample ls = sample_surface_light_li(scene, inct.position, inct.frame.z, sampler);
                    const Shader shader = create_shader(scene, scene.materials[inst.material_id], inct);
                    if (depth == 0) {
                        if (film.albedo) film_splat3(film.albedo, pixel, shader.albedo());
                        if (film.normal) film_splat3(film.normal, pixel, shader.normal());
                    }
                    const f3 wo = -d;
                    const float b0 = sampler.get1d(), b1 = sampler.get1d(), b2 = sampler.get1d();
                    BsdfSample bs = shader.sample(wo, b0, b1, b2);
                    f3 shadow_bsdf_val(0.0f);
                    float shadow_bsdf_pdf = 0.0f;
                    if (ls.success) {
                        shadow_bsdf_val = shader.eval(ls.d, wo);
                        ls.success = !is_zero(shadow_bsdf_val);
                        if (ls.success) shadow_bsdf_pdf = shader.pdf(ls.d, wo);
                    }
                    if (ls.success) {
                        const bool shadow_d_out = dot(inct.frame.z, ls.d) > 0.0f;
                        const bool last_ray_out = dot(inct.frame.z, d) < 0.0f;
                        if (shadow_d_out == last_ray_out) s_med = ray_medium;
                        else if (shadow_d_out) s_med = inst.outer_medium_id;
                        else s_med = inst.inner_medium_id;
                        const float cosv = fabsf(dot(normalize(ls.d), inct.frame.z));
                        s_beta = ls.li * beta * shadow_bsdf_val * cosv / (shadow_bsdf_pdf + ls.light_pdf);
                        s_o = ls.o; s_d = ls.d; s_t1 = ls.t1;
                        emit_shadow = true;
                    }
                    if (!is_zero(bs.bsdf)) {
                        const f3 dir = normalize(bs.dir);
                        const float cosv = fabsf(dot(dir, inct.frame.z));
                        const f3 new_beta_le = beta * cosv * bs.bsdf;
                        beta = new_beta_le / bs.pdf;
                        n_o = intersection_offset(inct.position, inct.frame.z, dir);
                        n_d = dir;
                        n_med = dot(inct.frame.z, dir) > 0.0f ? inst.outer_medium_id : inst.inner_medium_id;
                        n_pdf = bs.is_delta ? -bs.pdf : bs.pdf;
                        n_beta = beta; n_L = L; n_ble = new_beta_le; n_depth = depth + 1;
                        n_rng = make_ulonglong2(sampler.rng.state, sampler.rng.inc);
                        emit_next = true;
                    } else {
                        splat = true; splat_value = L;
                    }
                }
            }
        }

        // ---- converged: warp-aggregated queue appends
        const int next_slot = warp_alloc(&ctl->n_next, emit_next);
        if (emit_next) {
            nq.ray_o[next_slot] = make4(n_o, __uint_as_float(n_med));
            nq.ray_d[next_slot] = make4(n_d, kMaxFloat);
            nq.beta[next_slot] = make4(n_beta, __int_as_float(n_depth));
            nq.rad[next_slot] = make4(n_L, __uint_as_float(pixel));
            nq.rng[next_slot] = n_rng;
            nq.ble[next_slot] = make4(n_ble, n_pdf);
        }
        const int shadow_slot = warp_alloc(&ctl->n_shadow, emit_shadow);
        if (emit_shadow) {
            sq.ray_o[shadow_slot] = make4(s_o, __uint_as_float(s_med));
            sq.ray_d[shadow_slot] = make4(s_d, s_t1);
            sq.beta_li[shadow_slot] = make4(s_beta, __uint_as_float(pixel));
        }
        if (splat) film_splat3(film.radiance, pixel, splat_value);
        warp_count(&ctl->n_splat, splat);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shadow rays (shadow.cpp:24-74): occlusion with early exit; unoccluded => splat beta_li * Tr (Tr = 1 without media)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) shadow_kernel(SceneDevice scene, RenderParams P, ShadowQueue sq, FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->n_shadow;
    unsigned long long misses = 0;
    for (;;) {
        int base = 0;
        if (lane_id() == 0) base = atomicAdd(&ctl->shadow_cursor, 32);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const int i = base + lane_id();
        if (i < n) {
            const float4 o4 = sq.ray_o[i], d4 = sq.ray_d[i];
            const HitRec h = traverse_bvh2<true>(scene.bvh, xyz(o4), xyz(d4), d4.w);
            if (h.slot < 0) {
                const float4 b4 = sq.beta_li[i];
                film_splat3(film.radiance, __float_as_uint(b4.w), xyz(b4));
                misses += 1;
            }
        }
        __syncwarp();
    }
    for (int off = 16; off; off >>= 1) misses += __shfl_down_sync(kFull, misses, off);
    if (lane_id() == 0 && misses) {
        atomicAdd(&ctl->n_shadow_miss, misses);
        atomicAdd(&ctl->n_splat, misses);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// resolve (preview.cpp:28-80)
// ---------------------------------------------------------------------------------------------------------------
__global__ void resolve_kernel(int npix, const float4 *value, const float *weight, float4 *out, int normalise) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float4 v = value[i];
        const float w = weight[i];
        f3 o(0.0f);
        if (w > 0.0f) {
            o = xyz(v) / w;
            if (normalise) o = length(o) > 1e-3f ? normalize(o) : f3(0.0f);
        }
        out[i] = make4(o, 1.0f);
    }
}

inline int persistent_blocks(const LaunchCfg &cfg, int per_sm) { return cfg.sm_count * per_sm; }

}  // namespace

void launch_prepare(WaveCtl *ctl, cudaStream_t stream) { prepare_kernel<<<1, 1, 0, stream>>>(ctl); }

void launch_generate(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    generate_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, params, cur, film, ctl);
}

void launch_trace(const LaunchCfg &cfg, const SceneDevice &scene, PathQueue cur, HitQueue hit, WaveCtl *ctl, cudaStream_t stream) {
    trace_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, cur, hit, ctl);
}

void launch_shade(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, HitQueue hit, PathQueue next,
                  ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    shade_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
}

void launch_shadow(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, ShadowQueue shadow, FilmDevice film, WaveCtl *ctl,
                   cudaStream_t stream) {
    shadow_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, params, shadow, film, ctl);
}

void launch_resolve(int width, int height, const float *value, const float *weight, float *out, int normalise, cudaStream_t stream) {
    const int npix = width * height;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(npix, reinterpret_cast<const float4 *>(value), weight,
                                                          reinterpret_cast<float4 *>(out), normalise);
}

void launch_trace_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n, uint32_t *flag, uint4 *detail,
                        cudaStream_t stream) {
    trace_batch_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, ray_o, ray_d, n, flag, detail);
}

void launch_anyhit_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n, uint8_t *occluded,
                         cudaStream_t stream) {
    anyhit_batch_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, ray_o, ray_d, n, occluded);
}

}  // namespace btrc
