// ORACLE (test infrastructure only) -- CPU restatement of Btrc's materials (src/builtin/btrc/builtin/material).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
//
// Each function cites the reference file:line it follows (paths relative to /root/reference/src/builtin/btrc/builtin
// unless stated).  IEEE fp32, no FMA contraction (-ffp-contract=off); sin/cos/sqrt are the exact libm ones where
// the reference build uses .approx PTX -- the image tolerances absorb that.
#pragma once
#include <algorithm>
#include <cmath>

#include "orc_math.h"

namespace orc {

enum MaterialType { MAT_DIFFUSE = 0, MAT_BLACK = 1, MAT_DISNEY = 2, MAT_METAL = 3, MAT_GLASS = 4, MAT_MIRROR = 5, MAT_INVISIBLE = 6 };

inline float sqr(float x) { return x * x; }
inline V3 lerp3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }  // utils/cmath/cscalar.h lerp: a*(1-t)+b*t
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }

// ---- sampling warps (common/btrc/utils/cmath/cdistribution.cpp) ----------------------------------------------
inline V3 sample_hemisphere_zweighted(float u1, float u2) {  // :31-54
    float sx = 0, sy = 0;
    u1 = 2.0f * u1 - 1.0f;
    u2 = 2.0f * u2 - 1.0f;
    if (u1 != 0.0f || u2 != 0.0f) {
        float theta, r;
        if (std::fabs(u1) > std::fabs(u2)) { r = u1; theta = 0.25f * PI * (u2 / u1); }
        else { r = u2; theta = 0.5f * PI - 0.25f * PI * (u1 / u2); }
        sx = r * std::cos(theta); sy = r * std::sin(theta);
    }
    float z = std::sqrt(std::fmax(0.0f, 1.0f - (sx * sx + sy * sy)));
    return V3(sx, sy, z);
}
inline float pdf_sample_hemisphere_zweighted(V3 v) { return v.z >= 0 ? v.z / PI : 0.0f; }  // :66-69
inline V2 sample_triangle_uniform(float u1, float u2) { float t = std::sqrt(u1); return V2{1.0f - t, t * u2}; }  // :71-75
inline V3 sample_sphere_uniform(float u1, float u2) {  // :5-13
    float z = 1.0f - (u1 + u1);
    float phi = 2.0f * PI * u2;
    float r = std::sqrt(std::fmax(0.0f, 1.0f - z * z));
    return V3(r * std::cos(phi), r * std::sin(phi), z);
}
inline float pdf_sample_sphere_uniform() { return 1 / (4 * PI); }  // :25-28

// ---- local angles (common/btrc/utils/local_angle.h) ------------------------------------------------------------
namespace local_angle {
inline float cos_theta(V3 w) { return w.z; }
inline float cos2sin(float c) { return std::sqrt(std::fmax(0.0f, 1.0f - c * c)); }
inline float tan_theta(V3 w) { float t = 1.0f - w.z * w.z; return t > 0 ? std::sqrt(t) / w.z : 0.0f; }
inline float theta(V3 w) { return std::acos(clampf(w.z, -1.0f, 1.0f)); }
inline float phi(V3 w) {
    float result = 0.0f;
    if (w.x != 0 || w.y != 0) { result = std::atan2(w.y, w.x); if (result < 0) result = result + 2 * PI; }
    return result;
}
}  // namespace local_angle

// ---- fresnel (material/utils/fresnel.cpp:5-49) -------------------------------------------------------------------
inline float dielectric_fresnel(float eta_i, float eta_o, float cos_theta_i) {
    if (cos_theta_i < 0) { std::swap(eta_i, eta_o); cos_theta_i = -cos_theta_i; }
    float sin_theta_i = std::sqrt(std::fmax(0.0f, 1.0f - cos_theta_i * cos_theta_i));
    float sin_theta_t = eta_o / eta_i * sin_theta_i;
    if (sin_theta_t >= 1) return 1;
    float cos_theta_t = std::sqrt(std::fmax(0.0f, 1.0f - sin_theta_t * sin_theta_t));
    float para = (eta_i * cos_theta_i - eta_o * cos_theta_t) / (eta_i * cos_theta_i + eta_o * cos_theta_t);
    float perp = (eta_o * cos_theta_i - eta_i * cos_theta_t) / (eta_o * cos_theta_i + eta_i * cos_theta_t);
    return 0.5f * (para * para + perp * perp);
}
inline float schlick_approx(float R0, float cos_theta_i) {
    float t = 1.0f - cos_theta_i, t2 = t * t, t5 = t2 * t2 * t;
    return R0 + (1.0f - R0) * t5;
}
inline V3 schlick_approx(V3 R0, float cos_theta_i) {
    float t = 1.0f - cos_theta_i, t2 = t * t, t5 = t2 * t2 * t;
    return R0 + (V3(1, 1, 1) - R0) * t5;
}

// ---- microfacet (material/utils/microfacet.cpp:9-110) ------------------------------------------------------------
namespace microfacet {
inline float compute_a(float sin_phi_h, float cos_phi_h, float ax, float ay) { return sqr(cos_phi_h / ax) + sqr(sin_phi_h / ay); }
inline float smith_gtr2(float tan_theta, float alpha) {
    float result = 1.0f;
    if (tan_theta != 0) { float root = alpha * tan_theta; result = 2.0f / (1.0f + std::sqrt(1.0f + root * root)); }
    return result;
}
inline float anisotropic_gtr2(float sin_phi_h, float cos_phi_h, float sin_theta_h, float cos_theta_h, float ax, float ay) {
    float A = compute_a(sin_phi_h, cos_phi_h, ax, ay);
    float RD = sqr(sin_theta_h) * A + sqr(cos_theta_h);
    return 1.0f / (PI * ax * ay * sqr(RD));
}
inline float smith_anisotropic_gtr2(float cos_phi, float sin_phi, float ax, float ay, float tan_theta) {
    float t = sqr(ax * cos_phi) + sqr(ay * sin_phi);
    float sqr_val = 1.0f + t * sqr(tan_theta);
    float lambda = -0.5f + 0.5f * std::sqrt(sqr_val);
    return 1.0f / (1.0f + lambda);
}
inline V3 sample_anisotropic_gtr2(float ax, float ay, float s0, float s1) {
    float sin_phi_h = ay * std::sin(2 * PI * s0);
    float cos_phi_h = ax * std::cos(2 * PI * s0);
    float nor = 1.0f / std::sqrt(sqr(sin_phi_h) + sqr(cos_phi_h));
    sin_phi_h = sin_phi_h * nor;
    cos_phi_h = cos_phi_h * nor;
    float A = compute_a(sin_phi_h, cos_phi_h, ax, ay);
    float cos_theta_h = std::sqrt(A * (1.0f - s1) / ((1.0f - A) * s1 + A));
    float sin_theta_h = local_angle::cos2sin(cos_theta_h);
    return normalize(V3(sin_theta_h * cos_phi_h, sin_theta_h * sin_phi_h, cos_theta_h));
}
inline float gtr1(float sin_theta_h, float cos_theta_h, float alpha) {
    float U = sqr(alpha) - 1;
    float LD = 2.0f * PI * std::log(alpha);
    float RD = sqr(alpha * cos_theta_h) + sqr(sin_theta_h);
    return U / (LD * RD);
}
inline V3 sample_gtr1(float alpha, float s0, float s1) {
    float phi = 2 * PI * s0;
    float cos_theta = std::sqrt((std::pow(alpha, 2.0f - 2.0f * s1) - 1.0f) / (sqr(alpha) - 1.0f));
    float sin_theta = local_angle::cos2sin(cos_theta);
    return normalize(V3(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta));
}
inline V3 sample_anisotropic_gtr2_vnor(V3 ve, float ax, float ay, float s0, float s1) {
    V3 vh = normalize(V3(ax * ve.x, ay * ve.y, ve.z));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    V3 t1(1, 0, 0);
    if (lensq > 3e-4f) t1 = V3(-vh.y, vh.x, 0) / std::sqrt(lensq);
    V3 t2 = cross(vh, t1);
    float r = std::sqrt(s0);
    float phi = 2 * PI * s1;
    float t_1 = r * std::cos(phi);
    float _t_2 = r * std::sin(phi);
    float s = 0.5f * (1.0f + vh.z);
    float t_2 = (1.0f - s) * std::sqrt(1.0f - t_1 * t_1) + s * _t_2;
    V3 nh = t_1 * t1 + t_2 * t2 + std::sqrt(std::fmax(0.0f, 1.0f - t_1 * t_1 - t_2 * t_2)) * vh;
    return normalize(V3(ax * nh.x, ay * nh.y, std::fmax(0.0f, nh.z)));
}
}  // namespace microfacet

struct CompSample { V3 bsdf{0, 0, 0}, dir{0, 0, 0}; float pdf = 0; };

// ---- DiffuseComponentImpl (material/component/diffuse.cpp:5-60) --------------------------------------------------
struct DiffuseComponent {
    V3 albedo_value;
    CompSample sample(V3 lwo, V3 sam) const {
        CompSample r;
        if (lwo.z <= 0) return r;
        V3 lwi = sample_hemisphere_zweighted(sam.x, sam.y);
        r.bsdf = albedo_value / PI; r.dir = lwi; r.pdf = pdf_sample_hemisphere_zweighted(lwi);
        return r;
    }
    V3 eval(V3 lwi, V3 lwo) const { return (lwi.z <= 0 || lwo.z <= 0) ? V3(0) : albedo_value / PI; }
    float pdf(V3 lwi, V3 lwo) const { return (lwi.z <= 0 || lwo.z <= 0) ? 0.0f : pdf_sample_hemisphere_zweighted(lwi); }
    V3 albedo() const { return albedo_value; }
};

// ---- MicrofacetReflectionComponentImpl<ConductorFresnelPoint> (material/component/microfacet_reflection.h:10-121) --
struct MicrofacetReflection {
    V3 R0; float ax = 0, ay = 0;
    MicrofacetReflection() = default;
    MicrofacetReflection(V3 r0, float roughness, float anisotropic) : R0(r0) {
        float aspect = 1.0f;
        if (anisotropic > 0) aspect = std::sqrt(1.0f - 0.9f * anisotropic);
        ax = std::fmax(0.001f, roughness * roughness / aspect);
        ay = std::fmax(0.001f, roughness * roughness * aspect);
    }
    void eval_and_pdf(V3 lwi, V3 lwo, V3 &eval_out, float &pdf_out) const {
        if (lwi.z <= 0 || lwo.z <= 0) { eval_out = V3(0); pdf_out = 0; return; }
        float cos_theta_i = lwi.z, cos_theta_o = lwo.z;
        V3 lwh = normalize(lwi + lwo);
        float cos_theta_d = dot(lwi, lwh);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, ax, ay);
        float phi_i = local_angle::phi(lwi), phi_o = local_angle::phi(lwo);
        float sin_phi_i = std::sin(phi_i), cos_phi_i = std::cos(phi_i);
        float sin_phi_o = std::sin(phi_o), cos_phi_o = std::cos(phi_o);
        float tan_theta_i = local_angle::tan_theta(lwi), tan_theta_o = local_angle::tan_theta(lwo);
        float Go = microfacet::smith_anisotropic_gtr2(cos_phi_o, sin_phi_o, ax, ay, tan_theta_o);
        float Gi = microfacet::smith_anisotropic_gtr2(cos_phi_i, sin_phi_i, ax, ay, tan_theta_i);
        float G = Gi * Go;
        V3 F = schlick_approx(R0, cos_theta_d);
        eval_out = F * D * G / std::fabs(4.0f * cos_theta_i * cos_theta_o);
        pdf_out = Go * D / (4.0f * lwo.z);
    }
    CompSample sample(V3 lwo, V3 sam) const {
        CompSample r;
        if (lwo.z > 0) {
            V3 lwh = normalize(microfacet::sample_anisotropic_gtr2_vnor(lwo, ax, ay, sam.y, sam.z));
            if (lwh.z > 0) {
                V3 lwi = normalize(2.0f * dot(lwo, lwh) * lwh - lwo);
                if (lwi.z > 0) { r.dir = lwi; eval_and_pdf(lwi, lwo, r.bsdf, r.pdf); }
            }
        }
        return r;
    }
    V3 eval(V3 lwi, V3 lwo) const { V3 e; float p; eval_and_pdf(lwi, lwo, e, p); return e; }
    float pdf(V3 lwi, V3 lwo) const { V3 e; float p; eval_and_pdf(lwi, lwo, e, p); return p; }
    V3 albedo() const { return schlick_approx(R0, 1.0f); }
};

// ---- DisneyBSDF (material/disney.cpp:9-736), TransportMode::Radiance ---------------------------------------------
struct DisneyBSDF {
    V3 C, C_tint;
    float metallic, roughness, specular, specular_tint, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_roughness;
    float transmission, ior, transmission_roughness, trans_ax, trans_ay, ax, ay;
    float diffuse_weight, specular_weight, clearcoat_weight, transmission_weight;

    static float one_minus_5(float x) { float t = 1.0f - x, t2 = t * t; return t2 * t2 * t; }
    static V3 to_tint(V3 color) { float l = lum(color); return l > 0 ? color / l : V3(1, 1, 1); }
    static bool refract(V3 w, V3 n, float eta, V3 &output) {  // :57-74
        float cos_theta_i = dot(w, n);
        float sin2_theta_t = sqr(eta) * (1.0f - sqr(cos_theta_i));
        float cos2_theta_t = 1.0f - sin2_theta_t;
        if (cos2_theta_t <= 0) return false;
        float cos_theta_t = std::sqrt(cos2_theta_t);
        output = normalize((eta * cos_theta_i - cos_theta_t) * n - eta * w);
        return true;
    }

    DisneyBSDF() = default;
    DisneyBSDF(V3 base_color, float metallic_, float roughness_, float specular_, float specular_tint_, float anisotropic_, float sheen_,
               float sheen_tint_, float clearcoat_, float clearcoat_gloss, float transmission_, float transmission_roughness_, float ior_) {  // :426-475
        C = base_color; C_tint = to_tint(C);
        metallic = metallic_; roughness = roughness_; specular = specular_; specular_tint = specular_tint_; anisotropic = anisotropic_;
        sheen = sheen_; sheen_tint = sheen_tint_; transmission = transmission_; transmission_roughness = transmission_roughness_;
        ior = std::fmax(1.01f, ior_);
        float aspect = 1.0f;
        if (anisotropic_ > 0) aspect = std::sqrt(1.0f - 0.9f * anisotropic_);
        ax = std::fmax(0.001f, sqr(roughness_) / aspect);
        ay = std::fmax(0.001f, sqr(roughness_) * aspect);
        trans_ax = std::fmax(0.001f, sqr(transmission_roughness_) / aspect);
        trans_ay = std::fmax(0.001f, sqr(transmission_roughness_) * aspect);
        clearcoat = clearcoat_;
        clearcoat_roughness = std::fmax(0.0001f, sqr(lerp(0.1f, 0.0f, clearcoat_gloss)));
        float A = clampf(lum(base_color) * (1.0f - metallic_), 0.3f, 0.7f);
        float B = 1.0f - A;
        diffuse_weight = A * (1.0f - transmission_);
        transmission_weight = A * transmission_;
        specular_weight = B * 2.0f / (2.0f + clearcoat_);
        clearcoat_weight = B * clearcoat_ / (2.0f + clearcoat_);
    }

    V3 eval_diffuse(float cos_theta_i, float cos_theta_o, float cos_theta_d) const {  // :76-84
        V3 lambert = C / PI;
        float fl = one_minus_5(cos_theta_i), fv = one_minus_5(cos_theta_o);
        float rr = 2.0f * roughness * cos_theta_d * cos_theta_d;
        V3 retro_refl = C / PI * rr * (fl + fv + fl * fv * (rr - 1));
        return lambert * (1.0f - 0.5f * fl) * (1.0f - 0.5f * fv) + retro_refl;
    }
    V3 eval_sheen(float cos_theta_d) const {  // :86-90
        return 4.0f * sheen * lerp3(V3(1, 1, 1), C_tint, sheen_tint) * one_minus_5(cos_theta_d);
    }
    V3 eval_clearcoat(float cos_theta_i, float cos_theta_o, float tan_theta_i, float tan_theta_o, float sin_theta_h, float cos_theta_h,
                      float cos_theta_d) const {  // :92-103
        float D = microfacet::gtr1(sin_theta_h, cos_theta_h, clearcoat_roughness);
        float F = schlick_approx(0.04f, cos_theta_d);
        float G = microfacet::smith_gtr2(tan_theta_i, 0.25f) * microfacet::smith_gtr2(tan_theta_o, 0.25f);
        float val = clearcoat * D * F * G / std::fabs(4.0f * cos_theta_i * cos_theta_o);
        return V3(val, val, val);
    }
    V3 eval_trans(V3 lwi, V3 lwo) const {  // :105-152 (mode == Radiance)
        float cos_theta_i = lwi.z, cos_theta_o = lwo.z;
        float eta = cos_theta_o > 0 ? ior : 1.0f / ior;
        V3 lwh = normalize(lwo + eta * lwi);
        if (lwh.z < 0) lwh = -lwh;
        float cos_theta_d = dot(lwo, lwh);
        float F = dielectric_fresnel(ior, 1, cos_theta_d);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, trans_ax, trans_ay);
        float phi_i = local_angle::phi(lwi), phi_o = local_angle::phi(lwo);
        float sin_phi_i = std::sin(phi_i), cos_phi_i = std::cos(phi_i);
        float sin_phi_o = std::sin(phi_o), cos_phi_o = std::cos(phi_o);
        float tan_theta_i = local_angle::tan_theta(lwi), tan_theta_o = local_angle::tan_theta(lwo);
        float G = microfacet::smith_anisotropic_gtr2(cos_phi_i, sin_phi_i, trans_ax, trans_ay, tan_theta_i) *
                  microfacet::smith_anisotropic_gtr2(cos_phi_o, sin_phi_o, trans_ax, trans_ay, tan_theta_o);
        float sdem = cos_theta_d + eta * dot(lwi, lwh);
        float corr_factor = 1.0f / eta;
        V3 sqrtC(std::sqrt(C.x), std::sqrt(C.y), std::sqrt(C.z));
        float val = (1.0f - F) * D * G * eta * eta * dot(lwi, lwh) * dot(lwo, lwh) * corr_factor * corr_factor /
                    (cos_theta_i * cos_theta_o * sdem * sdem);
        float trans_factor = cos_theta_o > 0 ? transmission : 1.0f;
        return (1.0f - metallic) * trans_factor * sqrtC * std::fabs(val);
    }
    V3 eval_inner_refl(V3 lwi, V3 lwo) const {  // :154-186
        V3 lwh = -normalize(lwi + lwo);
        float cos_theta_d = dot(lwo, lwh);
        float F = dielectric_fresnel(ior, 1, cos_theta_d);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, trans_ax, trans_ay);
        float phi_i = local_angle::phi(lwi), phi_o = local_angle::phi(lwo);
        float sin_phi_i = std::sin(phi_i), cos_phi_i = std::cos(phi_i);
        float sin_phi_o = std::sin(phi_o), cos_phi_o = std::cos(phi_o);
        float tan_theta_i = local_angle::tan_theta(lwi), tan_theta_o = local_angle::tan_theta(lwo);
        float G = microfacet::smith_anisotropic_gtr2(cos_phi_i, sin_phi_i, trans_ax, trans_ay, tan_theta_i) *
                  microfacet::smith_anisotropic_gtr2(cos_phi_o, sin_phi_o, trans_ax, trans_ay, tan_theta_o);
        return transmission * C * std::fabs(F * D * G / (4.0f * lwi.z * lwo.z));
    }
    V3 eval_specular(V3 lwi, V3 lwo) const {  // :188-223
        float cos_theta_i = lwi.z, cos_theta_o = lwo.z;
        V3 lwh = normalize(lwi + lwo);
        float cos_theta_d = dot(lwi, lwh);
        V3 Cspec = lerp3(lerp3(V3(1, 1, 1), C_tint, specular_tint), C, metallic);
        V3 dielectric = Cspec * dielectric_fresnel(ior, 1.0f, cos_theta_d);
        V3 conductor = schlick_approx(Cspec, cos_theta_d);
        V3 F = lerp3(specular * dielectric, conductor, metallic);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, ax, ay);
        float phi_i = local_angle::phi(lwi), phi_o = local_angle::phi(lwo);
        float sin_phi_i = std::sin(phi_i), cos_phi_i = std::cos(phi_i);
        float sin_phi_o = std::sin(phi_o), cos_phi_o = std::cos(phi_o);
        float tan_theta_i = local_angle::tan_theta(lwi), tan_theta_o = local_angle::tan_theta(lwo);
        float G = microfacet::smith_anisotropic_gtr2(cos_phi_i, sin_phi_i, ax, ay, tan_theta_i) *
                  microfacet::smith_anisotropic_gtr2(cos_phi_o, sin_phi_o, ax, ay, tan_theta_o);
        return F * D * G / std::fabs(4.0f * cos_theta_i * cos_theta_o);
    }

    V3 sample_specular(V3 lwo, float s0, float s1) const {  // :230-252
        V3 lwh = microfacet::sample_anisotropic_gtr2_vnor(lwo, ax, ay, s0, s1);
        if (lwh.z <= 0) return V3(0);
        V3 lwi = 2.0f * dot(lwo, lwh) * lwh - lwo;
        if (lwi.z <= 0) return V3(0);
        return normalize(lwi);
    }
    V3 sample_clearcoat(V3 lwo, float s0, float s1) const {  // :254-275
        V3 lwh = microfacet::sample_gtr1(clearcoat_roughness, s0, s1);
        if (lwh.z <= 0) return V3(0);
        V3 lwi = 2.0f * dot(lwo, lwh) * lwh - lwo;
        if (lwi.z <= 0) return V3(0);
        return normalize(lwi);
    }
    V3 sample_transmission(V3 lwo, float s0, float s1) const {  // :277-308
        V3 lwh = microfacet::sample_anisotropic_gtr2(trans_ax, trans_ay, s0, s1);
        if (lwh.z <= 0) return V3(0);
        if ((lwo.z > 0) != (dot(lwh, lwo) > 0)) return V3(0);
        float eta = lwo.z > 0 ? 1.0f / ior : ior;
        V3 owh = dot(lwh, lwo) > 0 ? lwh : -lwh;
        V3 lwi;
        if (!refract(lwo, owh, eta, lwi)) return V3(0);
        if (lwi.z * lwo.z > 0 || ((lwi.z > 0) != (dot(lwh, lwi) > 0))) return V3(0);
        return normalize(lwi);
    }
    V3 sample_inner_refl(V3 lwo, float s0, float s1) const {  // :310-331
        V3 lwh = microfacet::sample_anisotropic_gtr2(trans_ax, trans_ay, s0, s1);
        if (lwh.z <= 0) return V3(0);
        V3 lwi = 2.0f * dot(lwo, lwh) * lwh - lwo;
        if (lwi.z > 0) return V3(0);
        return normalize(lwi);
    }

    void pdf_specular_clearcoat(V3 lwi, V3 lwo, float &pdf_specular, float &pdf_clearcoat) const {  // :339-367
        V3 lwh = normalize(lwi + lwo);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float cos_theta_d = dot(lwi, lwh);
        float cos_phi_o = std::cos(local_angle::phi(lwo));
        float sin_phi_o = local_angle::cos2sin(cos_phi_o);
        float tan_theta_o = local_angle::tan_theta(lwo);
        float specular_D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, ax, ay);
        pdf_specular = microfacet::smith_anisotropic_gtr2(cos_phi_o, sin_phi_o, ax, ay, tan_theta_o) * specular_D / (4.0f * lwo.z);
        float clearcoat_D = microfacet::gtr1(sin_theta_h, cos_theta_h, clearcoat_roughness);
        pdf_clearcoat = cos_theta_h * clearcoat_D / (4.0f * cos_theta_d);
    }
    float pdf_transmission(V3 lwi, V3 lwo) const {  // :369-404
        float eta = lwo.z > 0 ? ior : 1.0f / ior;
        V3 lwh = normalize(lwo + eta * lwi);
        if (lwh.z < 0) lwh = -lwh;
        if (((lwo.z > 0) != (dot(lwh, lwo) > 0)) || ((lwi.z > 0) != (dot(lwh, lwi) > 0))) return 0;
        float sdem = dot(lwo, lwh) + eta * dot(lwi, lwh);
        float dwh_to_dwi = eta * eta * dot(lwi, lwh) / (sdem * sdem);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, trans_ax, trans_ay);
        return std::fabs(dot(lwi, lwh) * D * dwh_to_dwi);
    }
    float pdf_inner_refl(V3 lwi, V3 lwo) const {  // :406-422
        V3 lwh = -normalize(lwi + lwo);
        float phi_h = local_angle::phi(lwh);
        float sin_phi_h = std::sin(phi_h), cos_phi_h = std::cos(phi_h);
        float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
        float cos_theta_d = dot(lwi, lwh);
        float D = microfacet::anisotropic_gtr2(sin_phi_h, cos_phi_h, sin_theta_h, cos_theta_h, trans_ax, trans_ay);
        return std::fabs(cos_theta_h * D / (4.0f * cos_theta_d));
    }

    V3 eval(V3 lwi, V3 lwo) const {  // :477-558
        if (std::fabs(lwi.z) < 1e-4f || std::fabs(lwo.z) < 1e-4f) return V3(0);
        if (lwi.z * lwo.z < 0) { if (transmission == 0.0f) return V3(0); return eval_trans(lwi, lwo); }
        if (lwi.z < 0 && lwo.z < 0) { if (transmission == 0.0f) return V3(0); return eval_inner_refl(lwi, lwo); }
        if (lwi.z <= 0 || lwo.z <= 0) return V3(0);
        float cos_theta_i = lwi.z, cos_theta_o = lwo.z;
        V3 lwh = normalize(lwi + lwo);
        float cos_theta_d = dot(lwi, lwh);
        V3 diffuse(0), sheen_v(0);
        if (metallic < 1.0f) {
            diffuse = eval_diffuse(cos_theta_i, cos_theta_o, cos_theta_d);
            if (sheen > 0.0f) sheen_v = eval_sheen(cos_theta_d);
        }
        V3 spec = eval_specular(lwi, lwo);
        V3 cc(0);
        if (clearcoat > 0.0f) {
            float tan_theta_i = local_angle::tan_theta(lwi), tan_theta_o = local_angle::tan_theta(lwo);
            float cos_theta_h = lwh.z, sin_theta_h = local_angle::cos2sin(cos_theta_h);
            cc = eval_clearcoat(cos_theta_i, cos_theta_o, tan_theta_i, tan_theta_o, sin_theta_h, cos_theta_h, cos_theta_d);
        }
        return (1.0f - metallic) * (1.0f - transmission) * (diffuse + sheen_v) + spec + cc;
    }
    float pdf(V3 lwi, V3 lwo) const {  // :682-730
        if (std::fabs(lwi.z) < 1e-4f || std::fabs(lwo.z) < 1e-4f) return 0;
        if (lwo.z < 0) {
            if (transmission == 0.0f) return 0;
            float macro_F = clampf(dielectric_fresnel(ior, 1, lwo.z), 0.1f, 0.9f);
            if (lwi.z > 0) return (1.0f - macro_F) * pdf_transmission(lwi, lwo);
            return macro_F * pdf_inner_refl(lwi, lwo);
        }
        if (lwi.z < 0) return transmission_weight * pdf_transmission(lwi, lwo);
        float diffuse = pdf_sample_hemisphere_zweighted(lwi);
        float spec, cc;
        pdf_specular_clearcoat(lwi, lwo, spec, cc);
        return diffuse_weight * diffuse + specular_weight * spec + clearcoat_weight * cc;
    }
    CompSample sample(V3 lwo, V3 sam) const {  // :560-658
        CompSample result;
        auto finish = [&](V3 lwi) {
            if (lwi.x == 0 && lwi.y == 0 && lwi.z == 0) return CompSample();
            CompSample r;
            r.dir = lwi; r.bsdf = eval(lwi, lwo); r.pdf = pdf(lwi, lwo);
            V3 v = r.bsdf * std::fabs(r.dir.z) / r.pdf;  // handle_bad_sample (:566-573)
            if (std::fmax(std::fmax(v.x, v.y), v.z) > 10) return CompSample();
            return r;
        };
        if (std::fabs(lwo.z) < 1e-4f) return result;
        if (lwo.z < 0) {
            if (transmission == 0.0f) return result;
            float macro_F = clampf(dielectric_fresnel(ior, 1, lwo.z), 0.1f, 0.9f);
            V3 lwi = sam.x >= macro_F ? sample_transmission(lwo, sam.y, sam.z) : sample_inner_refl(lwo, sam.y, sam.z);
            return finish(lwi);
        }
        float sel = sam.x;
        V3 lwi;
        if (sel < diffuse_weight) lwi = sample_hemisphere_zweighted(sam.y, sam.z);
        else {
            sel = sel - diffuse_weight;
            if (sel < transmission_weight) lwi = sample_transmission(lwo, sam.y, sam.z);
            else {
                sel = sel - transmission_weight;
                if (sel < specular_weight) lwi = sample_specular(lwo, sam.y, sam.z);
                else lwi = sample_clearcoat(lwo, sam.y, sam.z);
            }
        }
        return finish(lwi);
    }
    V3 albedo() const { return C; }
};

// ---- shader frame (material/utils/shader_frame.cpp:5-95) ---------------------------------------------------------
struct ShaderFrame {
    Frame geometry, shading;
    bool is_black_fringes(V3 w) const { return geometry.in_positive_z_hemisphere(w) != shading.in_positive_z_hemisphere(w); }
    float correct_shading_energy(V3 wi) const {
        float down = std::fabs(dot(geometry.z, wi));
        if (down > 1e-4f) return std::fabs(dot(shading.z, wi)) / down;
        return 1;
    }
    ShaderFrame flip_for_black_fringes(V3 wo) const {
        ShaderFrame f; f.geometry = geometry;
        f.shading = is_black_fringes(wo) ? geometry : shading;
        return f;
    }
};

struct BsdfSample { V3 bsdf{0, 0, 0}, dir{0, 0, 0}; float pdf = 0; bool is_delta = false; };

// Material parameters evaluated at the hit point (what every create_shader() samples from its textures).
struct ShaderParams {
    V3 color{0, 0, 0};
    float metallic = 0, roughness = 0, transmission = 0, ior = 1.5f, specular = 1, specular_tint = 0, anisotropic = 0, sheen = 0,
          sheen_tint = 0, clearcoat = 0, clearcoat_gloss = 1, transmission_roughness = 0;
};

// One shader object for all material types: BSDFAggregate with its single closure (diffuse.cpp:21-34, disney.cpp:813-842,
// metal.cpp:31-48; aggregate.cpp:32-230) or the closed-form glass / mirror / invisible / black shaders.
struct Shader {
    int type = MAT_DIFFUSE;
    ShaderFrame frame;
    bool shadow_terminator = true;
    DiffuseComponent diffuse;
    MicrofacetReflection metal;
    DisneyBSDF disney;
    V3 color;       // glass / mirror
    float ior = 1;  // glass

    bool is_aggregate() const { return type == MAT_DIFFUSE || type == MAT_DISNEY || type == MAT_METAL; }

    CompSample comp_sample(V3 lwo, V3 sam) const {
        return type == MAT_DIFFUSE ? diffuse.sample(lwo, sam) : type == MAT_METAL ? metal.sample(lwo, sam) : disney.sample(lwo, sam);
    }
    V3 comp_eval(V3 lwi, V3 lwo) const {
        return type == MAT_DIFFUSE ? diffuse.eval(lwi, lwo) : type == MAT_METAL ? metal.eval(lwi, lwo) : disney.eval(lwi, lwo);
    }
    float comp_pdf(V3 lwi, V3 lwo) const {
        return type == MAT_DIFFUSE ? diffuse.pdf(lwi, lwo) : type == MAT_METAL ? metal.pdf(lwi, lwo) : disney.pdf(lwi, lwo);
    }

    float shadow_terminator_term(V3 wi) const {  // aggregate.cpp:219-230
        if (!shadow_terminator) return 1.0f;
        float wgwi = dot(frame.geometry.z, wi);
        float wswi = dot(frame.shading.z, wi);
        float wgws = dot(frame.geometry.z, frame.shading.z);
        float g = saturate(wgwi / (wswi * wgws));
        return -g * g * g + g * g + g;
    }

    V3 albedo() const {
        switch (type) {
        case MAT_DIFFUSE: return diffuse.albedo();
        case MAT_METAL: return metal.albedo();
        case MAT_DISNEY: return disney.albedo();
        case MAT_GLASS: case MAT_MIRROR: return color;
        case MAT_INVISIBLE: return V3(1, 1, 1);
        default: return V3(0);
        }
    }
    // black / invisible use inct.frame.z (black.cpp:50-53, invisible.cpp:46-51): frame.shading is set to the geometry frame for them
    V3 normal() const { return frame.shading.z; }

    BsdfSample sample(V3 wo, V3 sam) const {
        BsdfSample result;
        if (type == MAT_BLACK) return result;
        if (type == MAT_INVISIBLE) {  // invisible.cpp:19-28
            result.bsdf = V3(1, 1, 1) / std::fabs(dot(frame.geometry.z, normalize(wo)));
            result.dir = -wo; result.pdf = 1; result.is_delta = true;
            return result;
        }
        if (type == MAT_MIRROR) {  // mirror.cpp:17-42
            ShaderFrame fr = frame.flip_for_black_fringes(wo);
            V3 lwo = fr.shading.global_to_local(wo);
            if (lwo.z > 0) {
                lwo = normalize(lwo);
                V3 lwi(-lwo.x, -lwo.y, lwo.z);
                V3 wi = fr.shading.local_to_global(lwi);
                if (!frame.is_black_fringes(wi)) {
                    V3 fres = schlick_approx(color, lwo.z);
                    V3 bsdf = fres / lwo.z;
                    float norm_factor = fr.correct_shading_energy(wi);
                    result.bsdf = bsdf * norm_factor; result.dir = wi; result.pdf = 1; result.is_delta = true;
                }
            }
            return result;
        }
        if (type == MAT_GLASS) {  // glass.cpp:41-98
            ShaderFrame fr = frame.flip_for_black_fringes(wo);
            V3 nwo = normalize(fr.shading.global_to_local(wo));
            float fres = dielectric_fresnel(ior, 1, nwo.z);
            if (sam.x < fres) {
                V3 lwi(-nwo.x, -nwo.y, nwo.z);
                V3 wi = fr.shading.local_to_global(lwi);
                if (frame.is_black_fringes(wi)) return BsdfSample();
                V3 bsdf = color * fres / std::fabs(lwi.z);
                float norm = fr.correct_shading_energy(wi);
                result.dir = wi; result.bsdf = bsdf * norm; result.pdf = fres; result.is_delta = true;
                return result;
            }
            V3 nor(0, 0, nwo.z > 0 ? 1.0f : -1.0f);
            float eta = nwo.z > 0 ? 1.0f / ior : ior;
            V3 nwi;
            {  // refr (glass.cpp:11-28)
                float cos_theta_i = std::fabs(nwo.z);
                float sin_theta_i_2 = std::fmax(0.0f, 1.0f - cos_theta_i * cos_theta_i);
                float sin_theta_t_2 = eta * eta * sin_theta_i_2;
                if (sin_theta_t_2 >= 1) return BsdfSample();
                float cosThetaT = std::sqrt(1.0f - sin_theta_t_2);
                nwi = normalize((eta * cos_theta_i - cosThetaT) * nor - eta * nwo);
            }
            V3 wi = fr.shading.local_to_global(nwi);
            if (frame.is_black_fringes(wi)) return BsdfSample();
            float corr = eta * eta;  // TransportMode::Radiance
            V3 f = corr * color * (1.0f - fres) / std::fabs(nwi.z);
            float norm = fr.correct_shading_energy(wi);
            result.bsdf = f * norm; result.dir = wi;
            result.pdf = 1.0f - dielectric_fresnel(ior, 1, nwi.z);  // glass.cpp:91-92: the second assignment wins
            result.is_delta = true;
            return result;
        }
        // BSDFAggregate::sample (aggregate.cpp:32-119) with one component of weight 1
        ShaderFrame fr = frame.flip_for_black_fringes(wo);
        V3 lwo = normalize(fr.shading.global_to_local(wo));
        float comp_selector = 1.0f * sam.x;
        V3 new_sam(comp_selector / 1.0f, sam.y, sam.z);
        CompSample cs = comp_sample(lwo, new_sam);
        if (cs.pdf <= 0) return result;
        V3 bsdf = cs.bsdf, lwi = cs.dir;
        float pdf = 1.0f * cs.pdf;
        result.dir = fr.shading.local_to_global(lwi);
        if (frame.is_black_fringes(result.dir)) return BsdfSample();
        result.bsdf = bsdf;
        result.pdf = pdf / 1.0f;
        float corr_factor = fr.correct_shading_energy(result.dir);
        float stt = shadow_terminator_term(result.dir);
        result.bsdf = result.bsdf * corr_factor * stt;
        result.is_delta = false;
        return result;
    }
    V3 eval(V3 wi, V3 wo) const {  // aggregate.cpp:185-197; glass/mirror/invisible/black: zero
        if (!is_aggregate()) return V3(0);
        ShaderFrame fr = frame.flip_for_black_fringes(wo);
        V3 lwi = normalize(fr.shading.global_to_local(wi));
        V3 lwo = normalize(fr.shading.global_to_local(wo));
        V3 result = V3(0) + comp_eval(lwi, lwo);
        float corr_factor = fr.correct_shading_energy(wi);
        float stt = shadow_terminator_term(wi);
        return result * corr_factor * stt;
    }
    float pdf(V3 wi, V3 wo) const {  // aggregate.cpp:199-209
        if (!is_aggregate()) return 0;
        ShaderFrame fr = frame.flip_for_black_fringes(wo);
        V3 lwi = normalize(fr.shading.global_to_local(wi));
        V3 lwo = normalize(fr.shading.global_to_local(wo));
        float result = 0 + 1.0f * comp_pdf(lwi, lwo);
        return result / 1.0f;
    }
};

}  // namespace orc
