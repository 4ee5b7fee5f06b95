// Scene evaluation shared by the shade / medium / shadow kernels: textures, hit reconstruction, light sampling,
// russian roulette, shader creation, participating media.
//
// Reference semantics followed (src/...):
//   core/scene.cpp:9-64                                   get_hitinfo
//   builtin/renderer/wavefront/helper.cpp:10-207          light sampling, russian roulette, Le at hits and misses
//   core/light_sampler.cpp:41-63                          uniform light selection
//   builtin/light/mesh_light.cpp:17-76, geometry/triangle_mesh.cpp:188-228   mesh area lights
//   builtin/light/gradient_sky.cpp:20-54, ibl.cpp:27-87, env_sampler.cpp:124-187   environment lights
//   builtin/medium/homogeneous.cpp:25-73, hetergeneous.cpp:25-106, core/medium.cpp:32-129   media, HG phase function
#pragma once
#include "bsdf.cuh"
#include "rng.cuh"

namespace btrc {

// ---- textures ----------------------------------------------------------------------------------------------------------
// General path: postfix program.  Kept out of line: the shade kernel evaluates up to 15 textures per hit and one copy of
// this loop is enough (the instruction footprint of the kernel is what limits its issue rate, see profiles/).
__device__ __noinline__ f3 eval_texture_program(const TexOp *__restrict__ ops, int n_ops, float2 uv) {
    f3 stack[4];
    int sp = 0;
    for (int k = 0; k < n_ops; ++k) {
        const TexOp &op = ops[k];
        const int code = op.op;
        if (code == TEXOP_CONST) {
            stack[sp++] = f3(op.value[0], op.value[1], op.value[2]);
        } else if (code == TEXOP_ARRAY) {  // builtin/texture2d/array2d.cpp:18-30
            const float u = op.m[0] * uv.x + op.m[1] * uv.y + op.m[2];
            const float v = op.m[3] * uv.x + op.m[4] * uv.y + op.m[5];
            const float4 t = tex2D<float4>((cudaTextureObject_t)op.tex, u, v);
            stack[sp++] = f3(t.x, t.y, t.z);
        } else {  // builtin/texture2d/math.cpp:20-96
            const f3 b = stack[--sp], a = stack[--sp];
            stack[sp++] = code == TEXOP_ADD ? a + b : code == TEXOP_SUB ? a - b : code == TEXOP_MUL ? a * b : a / b;
        }
    }
    return stack[0];
}

__device__ __forceinline__ f3 eval_texture(const SceneDevice &scene, int tex, float2 uv) {
    const TextureDesc td = scene.textures[tex];
    const TexOp &op = scene.tex_ops[td.first_op];
    if (td.n_ops == 1 && op.op == TEXOP_CONST) return f3(op.value[0], op.value[1], op.value[2]);  // literal / "constant"
    return eval_texture_program(&op, td.n_ops, uv);
}
__device__ __forceinline__ float eval_texture_float(const SceneDevice &scene, int tex, float2 uv) { return eval_texture(scene, tex, uv).x; }

// texture3d/binary.cpp:17-42 as a postfix program; out of line, only reached by scenes that use "add" / "mul" 3-D textures
__device__ __noinline__ f3 eval_texture3d_program(const Texture3DDesc *__restrict__ ops, int n_ops, f3 uvw) {
    f3 stack[TEX3D_STACK];
    int sp = 0;
    for (int k = 0; k < n_ops; ++k) {
        const Texture3DDesc &op = ops[k];
        const int code = op.is_array;
        if (code == TEX3D_CONST) {
            stack[sp++] = f3(op.value[0], op.value[1], op.value[2]);
        } else if (code == TEX3D_ARRAY) {
            const float4 v = tex3D<float4>((cudaTextureObject_t)op.tex, uvw.x, uvw.y, uvw.z);
            stack[sp++] = f3(v.x, v.y, v.z);
        } else {
            const f3 b = stack[--sp], a = stack[--sp];
            stack[sp++] = code == TEX3D_ADD ? a + b : a * b;
        }
    }
    return stack[0];
}

// T3P: the scene has "add" / "mul" 3-D textures (SceneDevice::has_medium == 2).  The kernels are instantiated for both cases so
// that scenes without them run the code they ran before the programs existed (no call, same registers).
template <bool T3P = false>
__device__ __forceinline__ f3 eval_texture3d(const SceneDevice &scene, int tex, f3 uvw) {
    const Texture3DDesc &t = scene.textures3d[tex];
    if (t.is_array == TEX3D_CONST) return f3(t.value[0], t.value[1], t.value[2]);
    if (T3P && t.is_array == TEX3D_PROGRAM) return eval_texture3d_program(scene.textures3d + (unsigned)t.tex, (int)(t.tex >> 32), uvw);
    const float4 v = tex3D<float4>((cudaTextureObject_t)t.tex, uvw.x, uvw.y, uvw.z);
    return f3(v.x, v.y, v.z);
}

// Texture fetches issued early and consumed later: the common texture is a single `array` node (one tex2D); issuing the
// fetches of a hit's normal map, colour and first textured scalars back to back overlaps their latencies instead of paying
// one texture round trip after the other (13 % of the Disney kernel's stall samples sat on those waits, profiles/r1i).
struct TexPending {
    f3 v;
    int tex;       // texture id, -1: nothing to do
    bool fetched;  // v holds the result; otherwise tex_finish runs the general program
};
__device__ __forceinline__ TexPending tex_issue(const SceneDevice &scene, int tex, float2 uv) {
    TexPending p;
    p.v = f3(0.0f); p.tex = tex; p.fetched = false;
    if (tex < 0) return p;
    const TextureDesc td = scene.textures[tex];
    const TexOp &op = scene.tex_ops[td.first_op];
    if (td.n_ops == 1 && op.op == TEXOP_ARRAY) {
        const float u = op.m[0] * uv.x + op.m[1] * uv.y + op.m[2];
        const float v = op.m[3] * uv.x + op.m[4] * uv.y + op.m[5];
        const float4 t = tex2D<float4>((cudaTextureObject_t)op.tex, u, v);
        p.v = f3(t.x, t.y, t.z);
        p.fetched = true;
    }
    return p;
}
__device__ __forceinline__ f3 tex_finish(const SceneDevice &scene, const TexPending &p, float2 uv) {
    return p.fetched ? p.v : eval_texture(scene, p.tex, uv);
}

// ---- hit reconstruction ------------------------------------------------------------------------------------------------
struct SurfacePoint {
    f3 position;
    Frame3 frame;
    f3 interp_z;
    float2 tex_coord;
};

__device__ __forceinline__ SurfacePoint get_hitinfo(f3 o, f3 d, const InstanceInfo &inst, float t, uint32_t prim, float2 uv) {
    SurfacePoint sp;
    sp.position = o + t * d;
    const float4 gx = ldg4(inst.gx_ua + prim), gy = ldg4(inst.gy_uba + prim), gz = ldg4(inst.gz_uca + prim);
    const float4 sa = ldg4(inst.sn_va + prim), sba = ldg4(inst.sn_vba + prim), sca = ldg4(inst.sn_vca + prim);
    Frame3 gf(xyz(gx), xyz(gy), xyz(gz));
    gf.x = normalize(xf_vector(inst.mat, gf.x));
    gf.y = normalize(xf_vector(inst.mat, gf.y));
    gf.z = normalize(xf_normal(inst.inv, gf.z));
    f3 interp = xyz(sa) + xyz(sba) * uv.x + xyz(sca) * uv.y;
    interp = normalize(xf_normal(inst.inv, interp));
    sp.frame = gf;
    sp.interp_z = interp;
    sp.tex_coord = make_float2(gx.w + gy.w * uv.x + gz.w * uv.y, sa.w + sba.w * uv.x + sca.w * uv.y);
    return sp;
}

// ---- environment lights -----------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 ibl_eval_local(const SceneDevice &scene, f3 dir) {  // ibl.cpp:80-87
    const float u = saturate(1.0f / (2.0f * kPi) * phi_of(dir));
    const float v = saturate(1.0f / kPi * theta_of(dir));
    return eval_texture(scene, scene.env.ibl_tex, make_float2(u, v));
}
__device__ __forceinline__ f3 ibl_to_local(const EnvLight &e, f3 g) {
    return f3(g.x * e.frame[0] + g.y * e.frame[1] + g.z * e.frame[2], g.x * e.frame[3] + g.y * e.frame[4] + g.z * e.frame[5],
              g.x * e.frame[6] + g.y * e.frame[7] + g.z * e.frame[8]);
}
__device__ __forceinline__ f3 env_eval_le(const SceneDevice &scene, f3 to_light) {
    const EnvLight &e = scene.env;
    if (e.type == ENV_GRADIENT_SKY) {
        const float cos_theta = dot(f3(e.up[0], e.up[1], e.up[2]), normalize(to_light));
        const float s = saturate(0.5f * (cos_theta + 1.0f));
        return f3(e.lower[0], e.lower[1], e.lower[2]) * (1.0f - s) + f3(e.upper[0], e.upper[1], e.upper[2]) * s;
    }
    return ibl_eval_local(scene, normalize(ibl_to_local(e, to_light)));
}
__device__ __forceinline__ float ibl_tile_pdf(const EnvLight &e, int tile_x, int tile_y, bool use_abs) {
    const float tile_pdf = __ldg(e.tile_probs + tile_y * e.lut_w + tile_x);
    const float u0 = float(tile_x) / e.lut_w, u1 = float(tile_x + 1) / e.lut_w;
    const float v0 = float(tile_y) / e.lut_h, v1 = float(tile_y + 1) / e.lut_h;
    const float c0 = __cosf(kPi * v0), c1 = __cosf(kPi * v1);
    (void)use_abs;
    return tile_pdf / fabsf(2.0f * kPi * (u1 - u0) * (c1 - c0));
}
__device__ __forceinline__ float env_pdf_li(const SceneDevice &scene, f3 to_light) {
    const EnvLight &e = scene.env;
    if (e.type == ENV_GRADIENT_SKY) return pdf_sample_sphere_uniform();
    const f3 dir = normalize(ibl_to_local(e, to_light));  // env_sampler.cpp:159-187
    const float u = phi_of(dir) / (2.0f * kPi), v = acosf(dir.z) / kPi;
    const int tile_x = min(max(int(floorf(u * e.lut_w)), 0), e.lut_w - 1);
    const int tile_y = min(max(int(floorf(v * e.lut_h)), 0), e.lut_h - 1);
    return ibl_tile_pdf(e, tile_x, tile_y, true);
}
struct EnvSample { f3 dir, radiance; float pdf; };
__device__ __forceinline__ EnvSample env_sample_li(const SceneDevice &scene, float s0, float s1, float s2) {
    const EnvLight &e = scene.env;
    EnvSample r;
    if (e.type == ENV_GRADIENT_SKY) {
        r.dir = sample_sphere_uniform(s0, s1);
        r.radiance = env_eval_le(scene, r.dir);
        r.pdf = pdf_sample_sphere_uniform();
        return r;
    }
    // EnvirLightSampler::sample (env_sampler.cpp:124-157)
    const uint32_t tile_idx = alias_sample(e.tile_alias, uint32_t(e.lut_w * e.lut_h), s0);
    const int tile_y = int(tile_idx) / e.lut_w, tile_x = int(tile_idx) % e.lut_w;
    const float u0 = float(tile_x) / e.lut_w, u1 = float(tile_x + 1) / e.lut_w;
    const float v0 = float(tile_y) / e.lut_h, v1 = float(tile_y + 1) / e.lut_h;
    float cv0 = __cosf(kPi * v0), cv1 = __cosf(kPi * v1);
    if (cv0 > cv1) { const float t = cv0; cv0 = cv1; cv1 = t; }
    const float cos_theta = cv0 + s2 * (cv1 - cv0);
    const float sin_theta = cos2sin(cos_theta);
    float sp, cp;
    fast_sincos(2.0f * kPi * lerpf(u0, u1, s1), sp, cp);
    const f3 dir(sin_theta * cp, sin_theta * sp, cos_theta);
    r.dir = f3(e.frame[0], e.frame[1], e.frame[2]) * dir.x + f3(e.frame[3], e.frame[4], e.frame[5]) * dir.y + f3(e.frame[6], e.frame[7], e.frame[8]) * dir.z;
    r.radiance = ibl_eval_local(scene, dir);
    r.pdf = __ldg(e.tile_probs + tile_idx) / (2.0f * kPi * (u1 - u0) * (cv1 - cv0));
    return r;
}

// ---- light sampling (helper.cpp:10-65, 125-142) -----------------------------------------------------------------------
struct LiSample {
    bool success;
    f3 o, d, li;
    float t1, light_pdf;
};

// SURFACE: o = intersection_offset(ref_pos, ref_nor, dir) (sample_surface_light_li); otherwise o = ref_pos (medium)
template <bool SURFACE, typename S>
__device__ __forceinline__ LiSample sample_light_li(const SceneDevice &scene, f3 ref_pos, f3 ref_nor, S &sampler) {
    LiSample r;
    r.success = false;
    sampler.begin3d();  // light sample: get3d, then the light choice get1d (helper.cpp:21-22)
    const float s0 = sampler.get1d(), s1 = sampler.get1d(), s2 = sampler.get1d();
    const float sel = sampler.get1d();
    const int n = scene.n_lights;
    if (n == 0) return r;
    const int idx = min(int(sel * n), n - 1);
    const float select_pdf = 1.0f / float(n);
    if (idx == scene.env.light_index) {  // process_envir_light
        const EnvSample es = env_sample_li(scene, s0, s1, s2);
        r.o = SURFACE ? intersection_offset(ref_pos, ref_nor, es.dir) : ref_pos;
        r.d = es.dir;
        r.t1 = kMaxFloat;
        r.li = es.radiance;
        r.light_pdf = select_pdf * es.pdf;
        r.success = true;
        return r;
    }
    const LightInfo &li = scene.lights[idx];
    const uint32_t prim = alias_sample(li.area_alias, (uint32_t)li.prim_count, s0);
    const float2 uv = sample_triangle_uniform(s1, s2);
    const float *pp = li.positions + 9 * (size_t)prim;
    const f3 a(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)), ba(__ldg(pp + 3), __ldg(pp + 4), __ldg(pp + 5)), ca(__ldg(pp + 6), __ldg(pp + 7), __ldg(pp + 8));
    const f3 pos = a + ba * uv.x + ca * uv.y;
    const f3 ez = normalize(xyz(ldg4(li.gz_uca + prim)));  // CFrame(ex, ey, ez).z
    const float pdf_area = 1.0f / li.total_area;
    const f3 spos = xf_point(li.mat, pos);
    const f3 snor = normalize(xf_normal(li.inv, ez));
    const f3 pos_to_ref = ref_pos - spos;
    const float dist2 = length_square(pos_to_ref);
    const float pdf_sa = li.inv_scale2 * pdf_area * dist2 / fabsf(dot(snor, normalize(pos_to_ref)));
    const f3 rad = dot(pos_to_ref, snor) > 0.0f ? f3(li.intensity[0], li.intensity[1], li.intensity[2]) : f3(0.0f);
    const f3 diff = spos - ref_pos;
    const f3 shadow_dst = intersection_offset(spos, snor, -diff);
    r.o = SURFACE ? intersection_offset(ref_pos, ref_nor, diff) : ref_pos;
    r.d = shadow_dst - r.o;
    r.t1 = 1.0f;
    r.li = rad;
    r.light_pdf = select_pdf * pdf_sa;
    r.success = true;
    return r;
}

// helper.cpp:144-172
__device__ __forceinline__ f3 eval_miss_le(const SceneDevice &scene, f3 d, f3 beta_le, float bsdf_pdf) {
    if (scene.env.type == ENV_NONE) return f3(0.0f);
    const f3 le = env_eval_le(scene, d);
    if (bsdf_pdf < 0.0f) return beta_le * le / -bsdf_pdf;
    const float light_pdf = env_pdf_li(scene, d) / float(scene.n_lights);
    return beta_le * le / (bsdf_pdf + light_pdf);
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
                const float pdf_area = 1.0f / li.total_area;
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

// helper.cpp:67-95
template <typename S>
__device__ __forceinline__ bool simple_russian_roulette(const RenderParams &P, f3 &beta, int depth, S &sampler) {
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

// Material::create_shader of every material type (see bsdf.cuh for the citations); normal map: utils/normal_map.cpp:13-21
template <int MAT>
__device__ __forceinline__ Shader<MAT> create_shader(const SceneDevice &scene, const MaterialDesc &mat, const SurfacePoint &inct) {
    Shader<MAT> sh;
    sh.shadow_terminator = mat.shadow_terminator != 0;
    sh.frame.geometry = inct.frame;
    if (MAT == MAT_BLACK || MAT == MAT_INVISIBLE) {
        sh.frame.shading = inct.frame;
        return sh;
    }
    sh.frame.shading = inct.frame.rotate_to_new_z(inct.interp_z);
    const float2 uv = inct.tex_coord;
    const uint32_t cmask = mat.const_mask;
    // issue first, consume below: normal map, colour, and (Disney) the first two textured scalar parameters
    const TexPending pend_normal = tex_issue(scene, mat.normal_tex, uv);
    const TexPending pend_color = tex_issue(scene, (cmask & (1u << SLOT_COLOR)) ? -1 : mat.tex[SLOT_COLOR], uv);
    uint32_t todo = MAT == MAT_DISNEY ? (~cmask & ((1u << SLOT_COUNT) - 2u)) : 0u;
    const int slot_a = todo ? __ffs(todo) - 1 : -1;
    todo &= todo - 1u;
    const int slot_b = todo ? __ffs(todo) - 1 : -1;
    todo &= todo - 1u;
    const TexPending pend_a = tex_issue(scene, slot_a >= 0 ? mat.tex[slot_a] : -1, uv);
    const TexPending pend_b = tex_issue(scene, slot_b >= 0 ? mat.tex[slot_b] : -1, uv);
    if (mat.normal_tex >= 0) {
        const f3 c = tex_finish(scene, pend_normal, uv);
        const f3 local_nor = 2.0f * c - f3(1.0f);
        const f3 world_nor = sh.frame.shading.local_to_global(local_nor);
        sh.frame.shading = sh.frame.shading.rotate_to_new_z(normalize(world_nor));
    }
    const f3 color = (cmask & (1u << SLOT_COLOR)) ? f3(mat.color_const[0], mat.color_const[1], mat.color_const[2]) : tex_finish(scene, pend_color, uv);
    if (MAT == MAT_DISNEY) {
        const f3 base = color;
        float v[SLOT_COUNT];
#pragma unroll
        for (int slot = 0; slot < SLOT_COUNT; ++slot) v[slot] = mat.scalar_const[slot];  // independent loads, registers
        // the few parameters that do have a texture graph: the first two were issued above, any further one runs the program here
        if (slot_a >= 0) {
            const float val = tex_finish(scene, pend_a, uv).x;
#pragma unroll
            for (int k = SLOT_METALLIC; k < SLOT_COUNT; ++k) if (k == slot_a) v[k] = val;
        }
        if (slot_b >= 0) {
            const float val = tex_finish(scene, pend_b, uv).x;
#pragma unroll
            for (int k = SLOT_METALLIC; k < SLOT_COUNT; ++k) if (k == slot_b) v[k] = val;
        }
        for (; todo; todo &= todo - 1u) {
            const int slot = __ffs(todo) - 1;
            const float val = eval_texture_float(scene, mat.tex[slot], uv);
#pragma unroll
            for (int k = SLOT_METALLIC; k < SLOT_COUNT; ++k) if (k == slot) v[k] = val;
        }
        sh.disney = disney_make(base, v[SLOT_METALLIC], v[SLOT_ROUGHNESS], v[SLOT_SPECULAR], v[SLOT_SPECULAR_TINT], v[SLOT_ANISOTROPIC], v[SLOT_SHEEN],
                                v[SLOT_SHEEN_TINT], v[SLOT_CLEARCOAT], v[SLOT_CLEARCOAT_GLOSS], v[SLOT_TRANSMISSION], v[SLOT_TRANSMISSION_ROUGHNESS],
                                v[SLOT_IOR]);
    } else if (MAT == MAT_METAL) {
        sh.metal.R0 = color;
        const float roughness = (cmask & (1u << SLOT_ROUGHNESS)) ? mat.scalar_const[SLOT_ROUGHNESS] : eval_texture_float(scene, mat.tex[SLOT_ROUGHNESS], uv);
        const float anisotropic = (cmask & (1u << SLOT_ANISOTROPIC)) ? mat.scalar_const[SLOT_ANISOTROPIC] : eval_texture_float(scene, mat.tex[SLOT_ANISOTROPIC], uv);
        float aspect = 1.0f;
        if (anisotropic > 0.0f) aspect = sqrtf(1.0f - 0.9f * anisotropic);
        sh.metal.ax = fmaxf(0.001f, roughness * roughness / aspect);
        sh.metal.ay = fmaxf(0.001f, roughness * roughness * aspect);
    } else {
        sh.color = color;
        if (MAT == MAT_GLASS) sh.ior = (cmask & (1u << SLOT_IOR)) ? mat.scalar_const[SLOT_IOR] : eval_texture_float(scene, mat.tex[SLOT_IOR], uv);
    }
    return sh;
}

// ---- media -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float henyey_greenstein(float g, float u) {  // utils/henyey_greenstein.h:10-18
    const float g2 = g * g;
    const float dem = 1.0f + g2 - 2.0f * (g * u);
    return (1.0f - g2) / (4.0f * kPi * dem * sqrtf(dem));
}
__device__ __forceinline__ float cos_between(f3 a, f3 b) { return dot(normalize(a), normalize(b)); }

struct PhaseShader {  // HenyeyGreensteinPhaseShader (core/medium.cpp:22-76)
    float g;
    f3 color;
    __device__ __forceinline__ float pdf(f3 wi, f3 wo) const { return henyey_greenstein(g, -cos_between(wi, wo)); }
    __device__ __forceinline__ void sample(f3 wo, float s0, float s1, f3 &phase, f3 &dir, float &pdf_out) const {
        const float sv = s0 + s0 - 1.0f;
        float u;
        if (fabsf(g) < 0.001f) u = sv;
        else {
            const float g2 = g * g;
            u = (1.0f + g2 - sqr((1.0f - g2) / (1.0f + g * sv))) / (g + g);
        }
        const float cos_theta = -u;
        const float sin_theta = cos2sin(cos_theta);
        float sp, cp;
        fast_sincos(2.0f * kPi * s1, sp, cp);
        const f3 local_wi(sin_theta * sp, sin_theta * cp, cos_theta);
        const float phase_val = henyey_greenstein(g, u);
        phase = color * phase_val;
        // CFrame::from_z(wo) (cframe.cpp:52-64)
        const f3 nz = normalize(wo);
        const f3 ny = 1.0f - fabsf(nz.x) < 0.1f ? cross(nz, f3(0.0f, 1.0f, 0.0f)) : cross(nz, f3(1.0f, 0.0f, 0.0f));
        const Frame3 fr(cross(nz, ny), ny, nz);
        dir = fr.local_to_global(local_wi);
        pdf_out = phase_val;
    }
};

__device__ __forceinline__ f3 medium_uvw(const MediumDesc &m, f3 p) {  // TransformMedium: world_to_local.apply_to_point
    return m.has_transform ? xf_point(m.world_to_local, p) : p;
}

// ---------------------------------------------------------------------------------------------------------------
// Volume primitives: VolumePrimitiveMedium::sample / tr (core/volume.cpp:112-195), boundary crossings and overlap
// queries (core/volume/bvh.cpp:146-326), delta / ratio tracking inside an overlap (core/volume/aggregate.cpp:129-253).
// The reference's per-overlap generated code + trie (volume/indexing.cpp) and its box BVH only select which
// primitives contain a point; closest boundary (a minimum) and overlap (a set) are order independent, so the kernel
// loops over the primitive table.  Overlap members are visited in creation order; more than 4 overlapping primitives:
// the four smallest ids (the reference: first four in its BVH's traversal order).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxVolOverlap = 4;  // volume/bvh.h:11

__device__ __forceinline__ f3 ld3(const float *p) { return f3(p[0], p[1], p[2]); }

__device__ __forceinline__ f3 vol_world_to_uvw(const VolumePrim &v, f3 p) {  // core/volume.cpp:42-49
    const f3 op = p - ld3(v.o);
    return f3(dot(op, ld3(v.x)) * v.inv_x2, dot(op, ld3(v.y)) * v.inv_y2, dot(op, ld3(v.z)) * v.inv_z2);
}

// volume/bvh.cpp:281-326: nearest boundary crossing (entry, else exit) on p + t d with t in (0, t_max); -1 if none
__device__ __forceinline__ float vol_prim_intersect(f3 p, f3 d, float t_max, const VolumePrim &prim) {
    const f3 op = p - ld3(prim.o);
    float t0 = 0.0f, t1 = t_max;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const f3 ax = ld3(a == 0 ? prim.x_div_x2 : (a == 1 ? prim.y_div_y2 : prim.z_div_z2));
        const float opd = dot(op, ax), dd = dot(d, ax);
        const float nd = (0.0f - opd) / dd, fd = (1.0f - opd) / dd;
        t0 = fmaxf(t0, fminf(nd, fd));
        t1 = fminf(t1, fmaxf(nd, fd));
    }
    if (t0 > t1) return -1.0f;
    if (t0 > 0.0f) return t0;
    if (t1 < t_max) return t1;
    return -1.0f;
}

__device__ __forceinline__ bool vol_find_closest_intersection(const SceneDevice &scene, f3 a, f3 b, f3 &out) {  // volume/bvh.cpp:146-214
    const f3 dir = b - a;
    float final_t = kMaxFloat, t_max = 1.0f;
    for (int i = 0; i < scene.n_vol_prims; ++i) {
        const float t = vol_prim_intersect(a, dir, t_max, scene.vol_prims[i]);
        if (t >= 0.0f) { final_t = fminf(final_t, t); t_max = final_t; }
    }
    if (final_t > 10.0f) return false;
    out = a * (1.0f - final_t) + b * final_t;
    return true;
}

struct VolOverlap {
    int data[kMaxVolOverlap];
    int count;
};

__device__ __forceinline__ VolOverlap vol_get_overlap(const SceneDevice &scene, f3 pos) {  // volume/bvh.cpp:216-279
    VolOverlap r;
    r.count = 0;
    for (int i = 0; i < scene.n_vol_prims && r.count < kMaxVolOverlap; ++i) {
        const VolumePrim &prim = scene.vol_prims[i];
        const f3 op = pos - ld3(prim.o);
        const float u = dot(op, ld3(prim.x_div_x2)), v = dot(op, ld3(prim.y_div_y2)), w = dot(op, ld3(prim.z_div_z2));
        if (0.0f <= u && u <= 1.0f && 0.0f <= v && v <= 1.0f && 0.0f <= w && w <= 1.0f) {
#pragma unroll
            for (int k = 0; k < kMaxVolOverlap; ++k) if (k == r.count) r.data[k] = i;  // register-resident list
            r.count++;
        }
    }
    return r;
}

__device__ __forceinline__ int vol_member(const VolOverlap &ov, int k) {
    int id = ov.data[0];
#pragma unroll
    for (int j = 1; j < kMaxVolOverlap; ++j) if (j == k) id = ov.data[j];
    return id;
}

__device__ __forceinline__ float vol_overlap_inv_max_density(const SceneDevice &scene, const VolOverlap &ov) {  // aggregate.cpp:121-127
    float m = 0.0f;
    for (int k = 0; k < ov.count; ++k) m += scene.vol_prims[vol_member(ov, k)].max_sigma_t;
    return 1.0f / fmaxf(m, 0.01f);
}

// density sum of the overlap's members at parameter tf of segment a -> b (members' uvw are affine in the position)
template <bool T3P = false>
__device__ __forceinline__ float vol_density_sum(const SceneDevice &scene, const VolOverlap &ov, f3 a, f3 b, float tf) {
    float density_sum = 0.0f;
    for (int k = 0; k < ov.count; ++k) {
        const VolumePrim &v = scene.vol_prims[vol_member(ov, k)];
        const f3 uvw = vol_world_to_uvw(v, a) * (1.0f - tf) + vol_world_to_uvw(v, b) * tf;
        density_sum = density_sum + eval_texture3d<T3P>(scene, v.sigma_t_tex, uvw).x;
    }
    return density_sum;
}

struct MediumSample {
    bool scattered;
    f3 position;
    PhaseShader shader;
};

// aggregate.cpp:129-199 (delta tracking over the summed densities; albedo = density-weighted mix, g = 0)
template <bool T3P = false, typename S>
__device__ __forceinline__ void vol_sample_scattering(const SceneDevice &scene, const VolOverlap &ov, f3 a, f3 b, S &sampler, MediumSample &r) {
    const float inv_max_density = vol_overlap_inv_max_density(scene, ov);
    const float t_max = length(b - a);
    float t = 0.0f;
    for (;;) {
        const float dt = -logf(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) { r.scattered = false; return; }
        const float tf = t / t_max;
        const float density_sum = vol_density_sum<T3P>(scene, ov, a, b, tf);
        if (sampler.get1d() < density_sum * inv_max_density) {
            r.scattered = true;
            r.position = a * (1.0f - tf) + b * tf;
            f3 albedo(0.0f);
            for (int k = 0; k < ov.count; ++k) {
                const VolumePrim &v = scene.vol_prims[vol_member(ov, k)];
                const f3 uvw = vol_world_to_uvw(v, a) * (1.0f - tf) + vol_world_to_uvw(v, b) * tf;
                const float weight = eval_texture3d<T3P>(scene, v.sigma_t_tex, uvw).x / density_sum;
                albedo = albedo + weight * eval_texture3d<T3P>(scene, v.albedo_tex, uvw);
            }
            r.shader.g = 0.0f;
            r.shader.color = albedo;
            return;
        }
    }
}

// aggregate.cpp:201-253 (ratio tracking)
template <bool T3P = false, typename S>
__device__ __forceinline__ float vol_tr_in_overlap(const SceneDevice &scene, const VolOverlap &ov, f3 a, f3 b, S &sampler) {
    const float inv_max_density = vol_overlap_inv_max_density(scene, ov);
    const float t_max = length(b - a);
    float result = 1.0f, t = 0.0f;
    for (;;) {
        const float dt = -logf(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        result = result * (1.0f - vol_density_sum<T3P>(scene, ov, a, b, t / t_max) * inv_max_density);
    }
    return result;
}

template <bool T3P = false, typename S>
__device__ __noinline__ MediumSample volprim_sample(const SceneDevice &scene, f3 a, f3 b, S &sampler) {  // core/volume.cpp:112-160
    MediumSample r;
    r.scattered = false;
    if (scene.n_vol_prims == 0) return r;
    f3 o = a;
    const f3 d = normalize(b - a);
    for (;;) {
        if (dot(b - o, d) <= 0.0f || length_square(b - o) < 0.00001f) { r.scattered = false; break; }
        f3 inct_pos;
        if (!vol_find_closest_intersection(scene, o, b, inct_pos)) inct_pos = b;
        const VolOverlap ov = vol_get_overlap(scene, 0.5f * (o + inct_pos));
        if (ov.count != 0) {
            vol_sample_scattering<T3P>(scene, ov, o, inct_pos, sampler, r);
            if (r.scattered) break;
        }
        o = inct_pos + 0.001f * d;
    }
    return r;
}

template <bool T3P = false, typename S>
__device__ __noinline__ float volprim_tr(const SceneDevice &scene, f3 a, f3 b, S &sampler) {  // core/volume.cpp:162-195
    float result = 1.0f;
    if (scene.n_vol_prims == 0) return result;
    f3 o = a;
    const f3 d = normalize(b - a);
    for (;;) {
        if (dot(b - o, d) <= 0.0f || length_square(b - o) < 0.0001f) break;
        f3 inct_pos;
        if (!vol_find_closest_intersection(scene, o, b, inct_pos)) inct_pos = b;
        const VolOverlap ov = vol_get_overlap(scene, 0.5f * (o + inct_pos));
        if (ov.count != 0) result = result * vol_tr_in_overlap<T3P>(scene, ov, o, inct_pos, sampler);
        o = inct_pos + 0.001f * d;
    }
    return result;
}

template <bool T3P = false, typename S>
__device__ __forceinline__ MediumSample medium_sample(const SceneDevice &scene, const MediumDesc &m, f3 a, f3 b, S &sampler) {
    MediumSample r;
    r.scattered = false;
    if (m.type == MED_VOLPRIM) return volprim_sample<T3P>(scene, a, b, sampler);
    if (m.type == MED_HOMOGENEOUS) {  // homogeneous.cpp:25-58
        const float t_max = length(b - a);
        const float st = -logf(1.0f - sampler.get1d()) / m.sigma_t;
        if (st < t_max) {
            const float t = st / t_max;
            r.scattered = true;
            r.position = a * (1.0f - t) + b * t;
            r.shader.g = m.g;
            r.shader.color = f3(m.albedo[0], m.albedo[1], m.albedo[2]);
        }
        return r;
    }
    // hetergeneous.cpp:25-73: delta tracking against the global majorant
    const float inv_max_density = m.inv_max_density;
    const float t_max = length(b - a);
    float t = 0.0f;
    const f3 local_a = f3(0.5f) + 0.5f * medium_uvw(m, a);
    const f3 local_b = f3(0.5f) + 0.5f * medium_uvw(m, b);
    const f3 local_ba_div_t_max = (local_b - local_a) / t_max;
    const f3 ba_div_t_max = (b - a) / t_max;
    for (;;) {
        const float dt = -logf(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        const f3 uvw = local_a + t * local_ba_div_t_max;
        const float density = eval_texture3d<T3P>(scene, m.sigma_t_tex, uvw).x;
        if (sampler.get1d() < density * inv_max_density) {
            r.scattered = true;
            r.position = a + t * ba_div_t_max;
            r.shader.color = eval_texture3d<T3P>(scene, m.albedo_tex, uvw);
            r.shader.g = eval_texture3d<T3P>(scene, m.g_tex, uvw).x;
            break;
        }
    }
    return r;
}

template <bool T3P = false, typename S>
__device__ __forceinline__ f3 medium_tr(const SceneDevice &scene, const MediumDesc &m, f3 a, f3 b, S &sampler) {
    if (m.type == MED_VOLPRIM) return f3(volprim_tr<T3P>(scene, a, b, sampler));
    if (m.type == MED_HOMOGENEOUS) {  // homogeneous.cpp:60-73: sigma_a = sigma_t * (1 - albedo)
        const f3 sigma_a = m.sigma_t * (f3(1.0f) - f3(m.albedo[0], m.albedo[1], m.albedo[2]));
        const f3 e = -length(a - b) * sigma_a;
        return f3(expf(e.x), expf(e.y), expf(e.z));
    }
    // hetergeneous.cpp:75-106: ratio tracking
    float result = 1.0f, t = 0.0f;
    const float t_max = length(b - a);
    const float inv_max_density = m.inv_max_density;
    const f3 local_a = f3(0.5f) + 0.5f * medium_uvw(m, a);
    const f3 local_b = f3(0.5f) + 0.5f * medium_uvw(m, b);
    const f3 local_ba_div_t_max = (local_b - local_a) / t_max;
    for (;;) {
        const float dt = -logf(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        const f3 uvw = local_a + t * local_ba_div_t_max;
        const float density = eval_texture3d<T3P>(scene, m.sigma_t_tex, uvw).x;
        result = result * (1.0f - density * inv_max_density);
    }
    return f3(result);
}

// Shadow-ray transmittance stream: keyed on the ray (murmur64A over its 32 bytes) instead of the reference's per-SLOT
// generators (soa_buffer.cpp:86-96), which make media renders irreproducible.  Same distribution, deterministic.
__device__ __forceinline__ Sampler shadow_sampler(float4 o_med, float4 d_t1) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = 0 ^ (32ull * m);
    const uint64_t blocks[4] = {uint64_t(__float_as_uint(o_med.x)) | (uint64_t(__float_as_uint(o_med.y)) << 32),
                                uint64_t(__float_as_uint(o_med.z)) | (uint64_t(__float_as_uint(o_med.w)) << 32),
                                uint64_t(__float_as_uint(d_t1.x)) | (uint64_t(__float_as_uint(d_t1.y)) << 32),
                                uint64_t(__float_as_uint(d_t1.z)) | (uint64_t(__float_as_uint(d_t1.w)) << 32)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint64_t k = blocks[i];
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    h ^= h >> r; h *= m; h ^= h >> r;
    Sampler s;
    s.rng.seed(h);
    return s;
}

// shadow.cpp:39-74: transmittance of an unoccluded shadow ray
template <bool T3P = false>
__device__ __forceinline__ f3 shadow_tr(const SceneDevice &scene, float4 o_med, float4 d_t1) {
    Sampler sm = shadow_sampler(o_med, d_t1);
    const f3 o = xyz(o_med), d = xyz(d_t1);
    if (d_t1.w > 1.0f) return medium_tr<T3P>(scene, scene.media[scene.vol_prim_medium_id], o, o + normalize(d) * scene.world_diagonal, sm);
    return medium_tr<T3P>(scene, scene.media[__float_as_uint(o_med.w)], o, o + d * d_t1.w, sm);
}

}  // namespace btrc
