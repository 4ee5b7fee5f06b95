// Material evaluation on the device: one `Shader<MAT>` value per hit.  The material type is a template parameter: the
// closest-hit kernel bins path states by material type and each bin is shaded by its own kernel instantiation
// (kernels.cu), so a warp runs one material's code.  This replaces the reference's per-object run-time code
// generation (core/context.h:321-382) with ordinary device functions whose parameters are data.
//
// Reference semantics followed (src/builtin/btrc/builtin/material/...):
//   utils/aggregate.cpp:32-230        BSDFAggregate (black-fringe flip, shading-energy correction, shadow terminator)
//   utils/shader_frame.cpp:5-95       ShaderFrame
//   component/diffuse.cpp:5-60        diffuse closure
//   component/microfacet_reflection.h:10-121, metal.cpp:31-48   metal
//   disney.cpp:9-736                  Disney BSDF (diffuse / sheen / specular / clearcoat / transmission / inner reflection)
//   glass.cpp:11-118, mirror.cpp:8-72, invisible.cpp:6-52, black.cpp:5-54
//   utils/microfacet.cpp:9-110, utils/fresnel.cpp:5-49, common/btrc/utils/local_angle.h
// Floating-point contract: sin/cos are the fast SFU forms, division and sqrt are the approximate ones
// (nvcc -prec-div=false -prec-sqrt=false -ftz=true); log/exp/pow/atan2/acos are the full-precision CUDA math functions.
#pragma once
#include "dev_math.cuh"

namespace btrc {

__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ f3 lerp3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ void fast_sincos(float x, float &s, float &c) { __sincosf(x, &s, &c); }

__device__ __forceinline__ f3 sample_sphere_uniform(float u1, float u2) {  // cdistribution.cpp:5-13
    const float z = 1.0f - (u1 + u1);
    const float phi = 2.0f * kPi * u2;
    const float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    fast_sincos(phi, s, c);
    return f3(r * c, r * s, z);
}
__device__ __forceinline__ float pdf_sample_sphere_uniform() { return 1.0f / (4.0f * kPi); }

// ---- local angles --------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cos2sin(float c) { return sqrtf(fmaxf(0.0f, 1.0f - c * c)); }
__device__ __forceinline__ float tan_theta(f3 w) {
    const float t = 1.0f - w.z * w.z;
    return t > 0.0f ? sqrtf(t) / w.z : 0.0f;
}
// (sin phi, cos phi) of the azimuth of w.  The reference computes phi = atan2(y, x) and then sin/cos of it; the pair
// below is the same quantity without the round trip through the angle (y / r, x / r), within the SFU's own error.
__device__ __forceinline__ void sincos_phi(f3 w, float &s, float &c) {
    const float r2 = w.x * w.x + w.y * w.y;
    if (r2 > 0.0f) {
        const float inv = rsqrtf(r2);
        s = w.y * inv; c = w.x * inv;
    } else {
        s = 0.0f; c = 1.0f;  // phi = 0
    }
}
__device__ __forceinline__ float phi_of(f3 w) {
    float r = 0.0f;
    if (w.x != 0.0f || w.y != 0.0f) {
        r = atan2f(w.y, w.x);
        if (r < 0.0f) r += 2.0f * kPi;
    }
    return r;
}
__device__ __forceinline__ float theta_of(f3 w) { return acosf(clampf(w.z, -1.0f, 1.0f)); }

// ---- fresnel -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dielectric_fresnel(float eta_i, float eta_o, float cos_theta_i) {
    if (cos_theta_i < 0.0f) { const float t = eta_i; eta_i = eta_o; eta_o = t; cos_theta_i = -cos_theta_i; }
    const float sin_theta_i = sqrtf(fmaxf(0.0f, 1.0f - cos_theta_i * cos_theta_i));
    const float sin_theta_t = eta_o / eta_i * sin_theta_i;
    if (sin_theta_t >= 1.0f) return 1.0f;
    const float cos_theta_t = sqrtf(fmaxf(0.0f, 1.0f - sin_theta_t * sin_theta_t));
    const float para = (eta_i * cos_theta_i - eta_o * cos_theta_t) / (eta_i * cos_theta_i + eta_o * cos_theta_t);
    const float perp = (eta_o * cos_theta_i - eta_i * cos_theta_t) / (eta_o * cos_theta_i + eta_i * cos_theta_t);
    return 0.5f * (para * para + perp * perp);
}
__device__ __forceinline__ float one_minus_5(float x) { const float t = 1.0f - x, t2 = t * t; return t2 * t2 * t; }
__device__ __forceinline__ float schlick1(float R0, float c) { return R0 + (1.0f - R0) * one_minus_5(c); }
__device__ __forceinline__ f3 schlick3(f3 R0, float c) { return R0 + (f3(1.0f) - R0) * one_minus_5(c); }

// ---- microfacet ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mf_compute_a(float sin_phi_h, float cos_phi_h, float ax, float ay) { return sqr(cos_phi_h / ax) + sqr(sin_phi_h / ay); }
__device__ __forceinline__ float smith_gtr2(float tan_t, float alpha) {
    if (tan_t == 0.0f) return 1.0f;
    const float root = alpha * tan_t;
    return 2.0f / (1.0f + sqrtf(1.0f + root * root));
}
__device__ __forceinline__ float anisotropic_gtr2(float sin_phi_h, float cos_phi_h, float sin_theta_h, float cos_theta_h, float ax, float ay) {
    const float A = mf_compute_a(sin_phi_h, cos_phi_h, ax, ay);
    const float RD = sqr(sin_theta_h) * A + sqr(cos_theta_h);
    return 1.0f / (kPi * ax * ay * sqr(RD));
}
__device__ __forceinline__ float smith_anisotropic_gtr2(float cos_phi, float sin_phi, float ax, float ay, float tan_t) {
    const float t = sqr(ax * cos_phi) + sqr(ay * sin_phi);
    const float lambda = -0.5f + 0.5f * sqrtf(1.0f + t * sqr(tan_t));
    return 1.0f / (1.0f + lambda);
}
// D of the half vector and the two Smith terms for (wi, wo), the block every GGX lobe of the reference repeats
struct GgxTerms { float D, Gi, Go; };
__device__ __forceinline__ GgxTerms ggx_terms(f3 lwi, f3 lwo, f3 lwh, float ax, float ay) {
    GgxTerms g;
    float sh, ch, si, ci, so, co;
    sincos_phi(lwh, sh, ch);
    sincos_phi(lwi, si, ci);
    sincos_phi(lwo, so, co);
    const float cos_theta_h = lwh.z, sin_theta_h = cos2sin(cos_theta_h);
    g.D = anisotropic_gtr2(sh, ch, sin_theta_h, cos_theta_h, ax, ay);
    g.Gi = smith_anisotropic_gtr2(ci, si, ax, ay, tan_theta(lwi));
    g.Go = smith_anisotropic_gtr2(co, so, ax, ay, tan_theta(lwo));
    return g;
}
__device__ __forceinline__ f3 sample_anisotropic_gtr2(float ax, float ay, float s0, float s1) {
    float sp, cp;
    fast_sincos(2.0f * kPi * s0, sp, cp);
    float sin_phi_h = ay * sp, cos_phi_h = ax * cp;
    const float nor = rsqrtf(sqr(sin_phi_h) + sqr(cos_phi_h));
    sin_phi_h *= nor; cos_phi_h *= nor;
    const float A = mf_compute_a(sin_phi_h, cos_phi_h, ax, ay);
    const float cos_theta_h = sqrtf(A * (1.0f - s1) / ((1.0f - A) * s1 + A));
    const float sin_theta_h = cos2sin(cos_theta_h);
    return normalize(f3(sin_theta_h * cos_phi_h, sin_theta_h * sin_phi_h, cos_theta_h));
}
__device__ __forceinline__ float gtr1(float sin_theta_h, float cos_theta_h, float alpha) {
    const float U = sqr(alpha) - 1.0f;
    const float LD = 2.0f * kPi * logf(alpha);
    const float RD = sqr(alpha * cos_theta_h) + sqr(sin_theta_h);
    return U / (LD * RD);
}
__device__ __forceinline__ f3 sample_gtr1(float alpha, float s0, float s1) {
    const float phi = 2.0f * kPi * s0;
    const float cos_theta = sqrtf((powf(alpha, 2.0f - 2.0f * s1) - 1.0f) / (sqr(alpha) - 1.0f));
    const float sin_theta = cos2sin(cos_theta);
    float s, c;
    fast_sincos(phi, s, c);
    return normalize(f3(sin_theta * c, sin_theta * s, cos_theta));
}
__device__ __forceinline__ f3 sample_anisotropic_gtr2_vnor(f3 ve, float ax, float ay, float s0, float s1) {
    const f3 vh = normalize(f3(ax * ve.x, ay * ve.y, ve.z));
    const float lensq = vh.x * vh.x + vh.y * vh.y;
    f3 t1(1.0f, 0.0f, 0.0f);
    if (lensq > 3e-4f) t1 = f3(-vh.y, vh.x, 0.0f) / sqrtf(lensq);
    const f3 t2 = cross(vh, t1);
    const float r = sqrtf(s0);
    float sp, cp;
    fast_sincos(2.0f * kPi * s1, sp, cp);
    const float t_1 = r * cp, _t_2 = r * sp;
    const float s = 0.5f * (1.0f + vh.z);
    const float t_2 = (1.0f - s) * sqrtf(1.0f - t_1 * t_1) + s * _t_2;
    const f3 nh = t_1 * t1 + t_2 * t2 + sqrtf(fmaxf(0.0f, 1.0f - t_1 * t_1 - t_2 * t_2)) * vh;
    return normalize(f3(ax * nh.x, ay * nh.y, fmaxf(0.0f, nh.z)));
}

struct CompSample { f3 bsdf, dir; float pdf; };
__device__ __forceinline__ CompSample comp_none() { CompSample r; r.bsdf = f3(0.0f); r.dir = f3(0.0f); r.pdf = 0.0f; return r; }

// ---- metal: MicrofacetReflectionComponentImpl<ConductorFresnelPoint> ----------------------------------------------
struct MetalParams { f3 R0; float ax, ay; };
__device__ __forceinline__ void metal_eval_pdf(const MetalParams &m, f3 lwi, f3 lwo, f3 &eval_out, float &pdf_out) {
    if (lwi.z <= 0.0f || lwo.z <= 0.0f) { eval_out = f3(0.0f); pdf_out = 0.0f; return; }
    const f3 lwh = normalize(lwi + lwo);
    const float cos_theta_d = dot(lwi, lwh);
    const GgxTerms g = ggx_terms(lwi, lwo, lwh, m.ax, m.ay);
    const f3 F = schlick3(m.R0, cos_theta_d);
    eval_out = F * g.D * (g.Gi * g.Go) / fabsf(4.0f * lwi.z * lwo.z);
    pdf_out = g.Go * g.D / (4.0f * lwo.z);
}
__device__ __forceinline__ f3 metal_sample_dir(const MetalParams &m, f3 lwo, float s1, float s2) {  // microfacet_reflection.h:33-56
    if (lwo.z > 0.0f) {
        const f3 lwh = normalize(sample_anisotropic_gtr2_vnor(lwo, m.ax, m.ay, s1, s2));
        if (lwh.z > 0.0f) {
            const f3 lwi = normalize(2.0f * dot(lwo, lwh) * lwh - lwo);
            if (lwi.z > 0.0f) return lwi;
        }
    }
    return f3(0.0f);
}

// ---- Disney --------------------------------------------------------------------------------------------------------
struct DisneyParams {
    f3 C, C_tint;
    float metallic, roughness, specular, specular_tint, sheen, sheen_tint, clearcoat, clearcoat_roughness;
    float transmission, ior, trans_ax, trans_ay, ax, ay;
    float diffuse_weight, specular_weight, clearcoat_weight, transmission_weight;
};

__device__ __forceinline__ DisneyParams disney_make(f3 base_color, float metallic, float roughness, float specular, float specular_tint,
                                                    float anisotropic, float sheen, float sheen_tint, float clearcoat, float clearcoat_gloss,
                                                    float transmission, float transmission_roughness, float ior) {  // disney.cpp:426-475
    DisneyParams p;
    p.C = base_color;
    const float l = lum(base_color);
    p.C_tint = l > 0.0f ? base_color / l : f3(1.0f);
    p.metallic = metallic; p.roughness = roughness; p.specular = specular; p.specular_tint = specular_tint;
    p.sheen = sheen; p.sheen_tint = sheen_tint; p.transmission = transmission;
    p.ior = fmaxf(1.01f, ior);
    float aspect = 1.0f;
    if (anisotropic > 0.0f) aspect = sqrtf(1.0f - 0.9f * anisotropic);
    p.ax = fmaxf(0.001f, sqr(roughness) / aspect);
    p.ay = fmaxf(0.001f, sqr(roughness) * aspect);
    p.trans_ax = fmaxf(0.001f, sqr(transmission_roughness) / aspect);
    p.trans_ay = fmaxf(0.001f, sqr(transmission_roughness) * aspect);
    p.clearcoat = clearcoat;
    p.clearcoat_roughness = fmaxf(0.0001f, sqr(lerpf(0.1f, 0.0f, clearcoat_gloss)));
    const float A = clampf(l * (1.0f - metallic), 0.3f, 0.7f);
    const float B = 1.0f - A;
    p.diffuse_weight = A * (1.0f - transmission);
    p.transmission_weight = A * transmission;
    p.specular_weight = B * 2.0f / (2.0f + clearcoat);
    p.clearcoat_weight = B * clearcoat / (2.0f + clearcoat);
    return p;
}

__device__ __forceinline__ bool disney_refract(f3 w, f3 n, float eta, f3 &out) {  // disney.cpp:57-74
    const float cos_theta_i = dot(w, n);
    const float cos2_theta_t = 1.0f - sqr(eta) * (1.0f - sqr(cos_theta_i));
    if (cos2_theta_t <= 0.0f) return false;
    out = normalize((eta * cos_theta_i - sqrtf(cos2_theta_t)) * n - eta * w);
    return true;
}

__device__ __forceinline__ f3 disney_eval_trans(const DisneyParams &p, f3 lwi, f3 lwo) {  // disney.cpp:105-152 (Radiance)
    const float eta = lwo.z > 0.0f ? p.ior : 1.0f / p.ior;
    f3 lwh = normalize(lwo + eta * lwi);
    if (lwh.z < 0.0f) lwh = -lwh;
    const float cos_theta_d = dot(lwo, lwh);
    const float F = dielectric_fresnel(p.ior, 1.0f, cos_theta_d);
    const GgxTerms g = ggx_terms(lwi, lwo, lwh, p.trans_ax, p.trans_ay);
    const float ih = dot(lwi, lwh);
    const float sdem = cos_theta_d + eta * ih;
    const float corr = 1.0f / eta;
    const float val = (1.0f - F) * g.D * (g.Gi * g.Go) * eta * eta * ih * cos_theta_d * corr * corr / (lwi.z * lwo.z * sdem * sdem);
    const float trans_factor = lwo.z > 0.0f ? p.transmission : 1.0f;
    return (1.0f - p.metallic) * trans_factor * f3(sqrtf(p.C.x), sqrtf(p.C.y), sqrtf(p.C.z)) * fabsf(val);
}
__device__ __forceinline__ f3 disney_eval_inner_refl(const DisneyParams &p, f3 lwi, f3 lwo) {  // disney.cpp:154-186
    const f3 lwh = -normalize(lwi + lwo);
    const float F = dielectric_fresnel(p.ior, 1.0f, dot(lwo, lwh));
    const GgxTerms g = ggx_terms(lwi, lwo, lwh, p.trans_ax, p.trans_ay);
    return p.transmission * p.C * fabsf(F * g.D * (g.Gi * g.Go) / (4.0f * lwi.z * lwo.z));
}
__device__ __forceinline__ f3 disney_sample_transmission(const DisneyParams &p, f3 lwo, float s0, float s1) {  // disney.cpp:277-308
    const f3 lwh = sample_anisotropic_gtr2(p.trans_ax, p.trans_ay, s0, s1);
    if (lwh.z <= 0.0f) return f3(0.0f);
    const float oh = dot(lwh, lwo);
    if ((lwo.z > 0.0f) != (oh > 0.0f)) return f3(0.0f);
    const float eta = lwo.z > 0.0f ? 1.0f / p.ior : p.ior;
    const f3 owh = oh > 0.0f ? lwh : -lwh;
    f3 lwi;
    if (!disney_refract(lwo, owh, eta, lwi)) return f3(0.0f);
    if (lwi.z * lwo.z > 0.0f || ((lwi.z > 0.0f) != (dot(lwh, lwi) > 0.0f))) return f3(0.0f);
    return normalize(lwi);
}
__device__ __forceinline__ f3 disney_sample_inner_refl(const DisneyParams &p, f3 lwo, float s0, float s1) {  // disney.cpp:310-331
    const f3 lwh = sample_anisotropic_gtr2(p.trans_ax, p.trans_ay, s0, s1);
    if (lwh.z <= 0.0f) return f3(0.0f);
    const f3 lwi = 2.0f * dot(lwo, lwh) * lwh - lwo;
    if (lwi.z > 0.0f) return f3(0.0f);
    return normalize(lwi);
}
__device__ __forceinline__ f3 reflect_about(f3 lwo, f3 lwh) {  // disney.cpp:230-275 (specular / clearcoat tail)
    if (lwh.z <= 0.0f) return f3(0.0f);
    const f3 lwi = 2.0f * dot(lwo, lwh) * lwh - lwo;
    if (lwi.z <= 0.0f) return f3(0.0f);
    return normalize(lwi);
}
__device__ __forceinline__ float disney_pdf_transmission(const DisneyParams &p, f3 lwi, f3 lwo) {  // disney.cpp:369-404
    const float eta = lwo.z > 0.0f ? p.ior : 1.0f / p.ior;
    f3 lwh = normalize(lwo + eta * lwi);
    if (lwh.z < 0.0f) lwh = -lwh;
    const float oh = dot(lwo, lwh), ih = dot(lwi, lwh);
    if (((lwo.z > 0.0f) != (oh > 0.0f)) || ((lwi.z > 0.0f) != (ih > 0.0f))) return 0.0f;
    const float sdem = oh + eta * ih;
    const float dwh_to_dwi = eta * eta * ih / (sdem * sdem);
    float sh, ch;
    sincos_phi(lwh, sh, ch);
    const float D = anisotropic_gtr2(sh, ch, cos2sin(lwh.z), lwh.z, p.trans_ax, p.trans_ay);
    return fabsf(ih * D * dwh_to_dwi);
}
__device__ __forceinline__ float disney_pdf_inner_refl(const DisneyParams &p, f3 lwi, f3 lwo) {  // disney.cpp:406-422
    const f3 lwh = -normalize(lwi + lwo);
    float sh, ch;
    sincos_phi(lwh, sh, ch);
    const float D = anisotropic_gtr2(sh, ch, cos2sin(lwh.z), lwh.z, p.trans_ax, p.trans_ay);
    return fabsf(lwh.z * D / (4.0f * dot(lwi, lwh)));
}

// eval and pdf share the half vector and the GGX terms, so they are computed together (the reference evaluates
// eval() and pdf() back to back for both the BSDF sample and the light sample: shade.cpp:143-160)
__device__ __forceinline__ void disney_eval_pdf(const DisneyParams &p, f3 lwi, f3 lwo, f3 &f, float &pdf) {  // disney.cpp:477-558, 682-730
    f = f3(0.0f); pdf = 0.0f;
    if (fabsf(lwi.z) < 1e-4f || fabsf(lwo.z) < 1e-4f) return;
    if (lwo.z < 0.0f) {  // from inside: refraction out or inner reflection
        if (p.transmission == 0.0f) return;
        const float macro_F = clampf(dielectric_fresnel(p.ior, 1.0f, lwo.z), 0.1f, 0.9f);
        if (lwi.z > 0.0f) { f = disney_eval_trans(p, lwi, lwo); pdf = (1.0f - macro_F) * disney_pdf_transmission(p, lwi, lwo); }
        else { f = disney_eval_inner_refl(p, lwi, lwo); pdf = macro_F * disney_pdf_inner_refl(p, lwi, lwo); }
        return;
    }
    if (lwi.z < 0.0f) {  // from outside, refracted in
        if (p.transmission != 0.0f) f = disney_eval_trans(p, lwi, lwo);
        pdf = p.transmission_weight * disney_pdf_transmission(p, lwi, lwo);
        return;
    }
    // reflection
    const float cos_theta_i = lwi.z, cos_theta_o = lwo.z;
    const f3 lwh = normalize(lwi + lwo);
    const float cos_theta_d = dot(lwi, lwh);
    f3 diffuse(0.0f), sheen(0.0f);
    if (p.metallic < 1.0f) {
        const f3 lambert = p.C / kPi;
        const float fl = one_minus_5(cos_theta_i), fv = one_minus_5(cos_theta_o);
        const float rr = 2.0f * p.roughness * cos_theta_d * cos_theta_d;
        const f3 retro = p.C / kPi * rr * (fl + fv + fl * fv * (rr - 1.0f));
        diffuse = lambert * (1.0f - 0.5f * fl) * (1.0f - 0.5f * fv) + retro;
        if (p.sheen > 0.0f) sheen = 4.0f * p.sheen * lerp3(f3(1.0f), p.C_tint, p.sheen_tint) * one_minus_5(cos_theta_d);
    }
    const GgxTerms g = ggx_terms(lwi, lwo, lwh, p.ax, p.ay);
    const f3 Cspec = lerp3(lerp3(f3(1.0f), p.C_tint, p.specular_tint), p.C, p.metallic);
    const f3 Fd = Cspec * dielectric_fresnel(p.ior, 1.0f, cos_theta_d);
    const f3 Fc = schlick3(Cspec, cos_theta_d);
    const f3 F = lerp3(p.specular * Fd, Fc, p.metallic);
    const f3 specular = F * g.D * (g.Gi * g.Go) / fabsf(4.0f * cos_theta_i * cos_theta_o);
    const float cos_theta_h = lwh.z, sin_theta_h = cos2sin(cos_theta_h);
    float cc_D = 0.0f;  // clearcoat == 0 => clearcoat_weight == 0: the reference's term is 0 * finite
    f3 clearcoat(0.0f);
    if (p.clearcoat > 0.0f) {
        cc_D = gtr1(sin_theta_h, cos_theta_h, p.clearcoat_roughness);
        const float Fcc = schlick1(0.04f, cos_theta_d);
        const float G = smith_gtr2(tan_theta(lwi), 0.25f) * smith_gtr2(tan_theta(lwo), 0.25f);
        clearcoat = f3(p.clearcoat * cc_D * Fcc * G / fabsf(4.0f * cos_theta_i * cos_theta_o));
    }
    f = (1.0f - p.metallic) * (1.0f - p.transmission) * (diffuse + sheen) + specular + clearcoat;
    // pdf_specular_clearcoat (disney.cpp:339-367): sin_phi_o there is cos2sin(cos_phi_o), i.e. |sin phi_o|; it enters squared
    const float pdf_specular = g.Go * g.D / (4.0f * lwo.z);
    const float pdf_clearcoat = cos_theta_h * cc_D / (4.0f * cos_theta_d);
    pdf = p.diffuse_weight * pdf_sample_hemisphere_zweighted(lwi) + p.specular_weight * pdf_specular + p.clearcoat_weight * pdf_clearcoat;
}

// Lobe selection and direction sampling of DisneyBSDF::sample (disney.cpp:560-658) without the evaluation: the caller
// evaluates eval / pdf for the sampled and the light direction through ONE copy of disney_eval_pdf (code size, see shade_kernel).
__device__ __forceinline__ f3 disney_sample_dir(const DisneyParams &p, f3 lwo, float s0, float s1, float s2) {
    if (fabsf(lwo.z) < 1e-4f) return f3(0.0f);
    if (lwo.z < 0.0f) {
        if (p.transmission == 0.0f) return f3(0.0f);
        const float macro_F = clampf(dielectric_fresnel(p.ior, 1.0f, lwo.z), 0.1f, 0.9f);
        return s0 >= macro_F ? disney_sample_transmission(p, lwo, s1, s2) : disney_sample_inner_refl(p, lwo, s1, s2);
    }
    float sel = s0;
    if (sel < p.diffuse_weight) return sample_hemisphere_zweighted(s1, s2);
    sel -= p.diffuse_weight;
    if (sel < p.transmission_weight) return disney_sample_transmission(p, lwo, s1, s2);
    sel -= p.transmission_weight;
    if (sel < p.specular_weight) return reflect_about(lwo, sample_anisotropic_gtr2_vnor(lwo, p.ax, p.ay, s1, s2));
    return reflect_about(lwo, sample_gtr1(p.clearcoat_roughness, s1, s2));
}

// ---- shader ----------------------------------------------------------------------------------------------------------
struct ShaderFrame {
    Frame3 geometry, shading;
    __device__ __forceinline__ bool is_black_fringes(f3 w) const {
        return geometry.in_positive_z_hemisphere(w) != shading.in_positive_z_hemisphere(w);
    }
};
__device__ __forceinline__ float correct_shading_energy(f3 geometry_z, f3 shading_z, f3 wi) {
    const float down = fabsf(dot(geometry_z, wi));
    return down > 1e-4f ? fabsf(dot(shading_z, wi)) / down : 1.0f;
}

struct BsdfSample {
    f3 bsdf, dir;
    float pdf;
    bool is_delta;
};
__device__ __forceinline__ BsdfSample bsdf_none() { BsdfSample r; r.bsdf = f3(0.0f); r.dir = f3(0.0f); r.pdf = 0.0f; r.is_delta = false; return r; }

// MAT is a compile-time constant: every shade kernel instantiation carries the code of one material type only
// (this is the template specialisation that replaces CUJ's per-object code generation; core/context.h:321-382).
template <int MAT>
struct Shader {
    static constexpr int type = MAT;
    bool shadow_terminator;
    ShaderFrame frame;
    f3 color;          // diffuse albedo / glass & mirror colour
    float ior;         // glass
    MetalParams metal;
    DisneyParams disney;

    __device__ __forceinline__ bool is_aggregate() const { return type == MAT_DIFFUSE || type == MAT_DISNEY || type == MAT_METAL; }

    __device__ __forceinline__ float shadow_terminator_term(f3 wi) const {  // aggregate.cpp:219-230
        if (!shadow_terminator) return 1.0f;
        const float wgwi = dot(frame.geometry.z, wi);
        const float wswi = dot(frame.shading.z, wi);
        const float wgws = dot(frame.geometry.z, frame.shading.z);
        const float g = saturate(wgwi / (wswi * wgws));
        return -g * g * g + g * g + g;
    }
    __device__ __forceinline__ f3 albedo() const {
        switch (type) {
        case MAT_DIFFUSE: case MAT_GLASS: case MAT_MIRROR: return color;
        case MAT_METAL: return schlick3(metal.R0, 1.0f);
        case MAT_DISNEY: return disney.C;
        case MAT_INVISIBLE: return f3(1.0f);
        default: return f3(0.0f);
        }
    }
    __device__ __forceinline__ f3 normal() const { return frame.shading.z; }
    __device__ __forceinline__ Frame3 flipped_shading(f3 wo) const {  // flip_for_black_fringes
        return frame.is_black_fringes(wo) ? frame.geometry : frame.shading;
    }

    __device__ __forceinline__ void comp_eval_pdf(f3 lwi, f3 lwo, f3 &f, float &pdf) const {
        if (type == MAT_DIFFUSE) {
            const bool ok = !(lwi.z <= 0.0f || lwo.z <= 0.0f);
            f = ok ? color / kPi : f3(0.0f);
            pdf = ok ? pdf_sample_hemisphere_zweighted(lwi) : 0.0f;
        } else if (type == MAT_METAL) {
            metal_eval_pdf(metal, lwi, lwo, f, pdf);
        } else {
            disney_eval_pdf(disney, lwi, lwo, f, pdf);
        }
    }

    // ---- aggregate materials (diffuse / metal / disney), BSDFAggregate with one closure of weight 1 (aggregate.cpp:32-230).
    // The kernel drives them in three steps so that the closure's eval+pdf code exists once:
    //   lwi = sample_local_dir(lwo, sam)                       direction of BSDFComponent::sample
    //   comp_eval_pdf(lwi | local light direction, lwo)        BSDFComponent::eval / pdf
    //   wrap(global dir)                                       black fringes, shading-energy correction, shadow terminator
    __device__ __forceinline__ f3 sample_local_dir(f3 lwo, float s0, float s1, float s2) const {
        if (type == MAT_DIFFUSE) return lwo.z > 0.0f ? sample_hemisphere_zweighted(s0, s1) : f3(0.0f);  // component/diffuse.cpp:10-30
        if (type == MAT_METAL) return metal_sample_dir(metal, lwo, s1, s2);
        return disney_sample_dir(disney, lwo, s0, s1, s2);
    }
    // DisneyBSDF::sample's handle_bad_sample (disney.cpp:566-573)
    __device__ __forceinline__ bool reject_sample(f3 f, f3 lwi, float pdf) const {
        if (type != MAT_DISNEY) return false;
        const f3 v = f * fabsf(lwi.z) / pdf;
        return fmaxf(fmaxf(v.x, v.y), v.z) > 10.0f;
    }
    __device__ __forceinline__ float wrap_factor(f3 shading_z, f3 w) const {
        return correct_shading_energy(frame.geometry.z, shading_z, w) * shadow_terminator_term(w);
    }

    // ---- delta / empty shaders: the whole Shader::sample (eval and pdf are zero)
    __device__ __forceinline__ BsdfSample sample_delta(f3 wo, float s0) const {
        BsdfSample r = bsdf_none();
        if (type == MAT_BLACK) return r;
        if (type == MAT_INVISIBLE) {  // invisible.cpp:19-28
            r.bsdf = f3(1.0f) / fabsf(dot(frame.geometry.z, normalize(wo)));
            r.dir = -wo; r.pdf = 1.0f; r.is_delta = true;
            return r;
        }
        const Frame3 sh = flipped_shading(wo);
        if (type == MAT_MIRROR) {  // mirror.cpp:17-42
            f3 lwo = sh.global_to_local(wo);
            if (lwo.z > 0.0f) {
                lwo = normalize(lwo);
                const f3 wi = sh.local_to_global(f3(-lwo.x, -lwo.y, lwo.z));
                if (!frame.is_black_fringes(wi)) {
                    r.bsdf = schlick3(color, lwo.z) / lwo.z * correct_shading_energy(frame.geometry.z, sh.z, wi);
                    r.dir = wi; r.pdf = 1.0f; r.is_delta = true;
                }
            }
            return r;
        }
        // glass.cpp:41-98
        const f3 nwo = normalize(sh.global_to_local(wo));
        const float fr = dielectric_fresnel(ior, 1.0f, nwo.z);
        if (s0 < fr) {
            const f3 lwi(-nwo.x, -nwo.y, nwo.z);
            const f3 wi = sh.local_to_global(lwi);
            if (frame.is_black_fringes(wi)) return r;
            r.bsdf = color * fr / fabsf(lwi.z) * correct_shading_energy(frame.geometry.z, sh.z, wi);
            r.dir = wi; r.pdf = fr; r.is_delta = true;
            return r;
        }
        const f3 nor(0.0f, 0.0f, nwo.z > 0.0f ? 1.0f : -1.0f);
        const float eta = nwo.z > 0.0f ? 1.0f / ior : ior;
        const float cos_theta_i = fabsf(nwo.z);
        const float sin_theta_t_2 = eta * eta * fmaxf(0.0f, 1.0f - cos_theta_i * cos_theta_i);
        if (sin_theta_t_2 >= 1.0f) return r;
        const f3 nwi = normalize((eta * cos_theta_i - sqrtf(1.0f - sin_theta_t_2)) * nor - eta * nwo);
        const f3 wi = sh.local_to_global(nwi);
        if (frame.is_black_fringes(wi)) return r;
        r.bsdf = (eta * eta) * color * (1.0f - fr) / fabsf(nwi.z) * correct_shading_energy(frame.geometry.z, sh.z, wi);
        r.dir = wi;
        r.pdf = 1.0f - dielectric_fresnel(ior, 1.0f, nwi.z);  // glass.cpp:91-92: the second assignment wins
        r.is_delta = true;
        return r;
    }
};

}  // namespace btrc
