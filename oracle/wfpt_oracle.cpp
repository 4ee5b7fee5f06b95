// ORACLE (test infrastructure only) -- CPU restatement of Btrc's wavefront path tracer ("wfpt").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
// this library; the product (btrc_b200/csrc) never links, includes or calls it.
//
// Parity pins: closest-hit is pinned against the reference's own vendored lib/bvh (oracle/_ref,
// built from /root/reference/lib/bvh by oracle/Makefile) and against lib/bvh's golden image
// (tests/golden/cornell_box_reference.png).  Everything else on the path (generate / shade /
// shadow / film) is *parity unpinned by the reference* (Btrc ships no tests or fixtures for it,
// and its device code cannot be built here: it needs LLVM 11 + OptiX 7); this file restates the
// reference source line by line, each function citing the file:line it follows (paths relative to
// /root/reference/src unless they start with lib/).
//
// Paths are independent (each carries its own PCG stream), so the oracle walks
// one path at a time instead of in waves; per-path arithmetic and RNG draw order are those of
// the wavefront kernels.  All arithmetic is IEEE fp32, compiled with -ffp-contract=off.
#include "orc_math.h"
#include "orc_bsdf.h"
#include "orc_texture.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

// ------------------------------------------------------------------------------------------
// RNG: cstd::PCG (lib/cuj/cuj/src/cstd/random.cpp:83-161), murmur64A (utils/hash.h:12-99),
// IndependentSampler (builtin/sampler/independent.cpp:6-25)
// ------------------------------------------------------------------------------------------
namespace {

constexpr uint64_t PCG_DEFAULT_STATE = 0x853c49e6748fea9bULL;
constexpr uint64_t PCG_MULT = 0x5851f42d4c957f2dULL;

struct PCG {
    uint64_t state = 0, inc = 0;
    uint32_t next() {  // random.cpp:153-160
        uint64_t old = state;
        state = old * PCG_MULT + inc;
        uint32_t xs = uint32_t(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = uint32_t(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    void seed(uint64_t s) {  // random.cpp:113-120
        state = 0;
        inc = (s << 1u) | 1u;
        next();
        state = state + PCG_DEFAULT_STATE;
        next();
    }
    void advance(int64_t idelta) {  // random.cpp:122-141
        uint64_t cur_mult = PCG_MULT, cur_plus = inc, acc_mult = 1, acc_plus = 0;
        uint64_t delta = uint64_t(idelta);
        while (delta > 0) {
            if (delta & 1) {
                acc_mult = acc_mult * cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult = cur_mult * cur_mult;
            delta = delta / 2;
        }
        state = acc_mult * state + acc_plus;
    }
    float uniform_float() {  // random.cpp:143-146
        return std::fmin(float(next()) * 0x1p-32f, 0x1.fffffep-1f);
    }
};

uint64_t murmur_hash64A(const uint8_t *key, uint64_t len, uint64_t seed) {  // hash.h:12-70
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ (len * m);
    const uint8_t *end = key + 8 * (len / 8);
    while (key != end) {
        uint64_t k;
        std::memcpy(&k, key, 8);
        key += 8;
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    switch (len & 7) {
    case 7: h ^= uint64_t(key[6]) << 48; [[fallthrough]];
    case 6: h ^= uint64_t(key[5]) << 40; [[fallthrough]];
    case 5: h ^= uint64_t(key[4]) << 32; [[fallthrough]];
    case 4: h ^= uint64_t(key[3]) << 24; [[fallthrough]];
    case 3: h ^= uint64_t(key[2]) << 16; [[fallthrough]];
    case 2: h ^= uint64_t(key[1]) << 8; [[fallthrough]];
    case 1: h ^= uint64_t(key[0]); h *= m;
    }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}

uint64_t hash_pixel(uint32_t px, uint32_t py) {  // hash.h:88-96 with (u32,u32) => 8 bytes
    uint8_t buf[8];
    std::memcpy(buf, &px, 4);
    std::memcpy(buf + 4, &py, 4);
    return murmur_hash64A(buf, 8, 0);
}

// ------------------------------------------------------------------------------------------
// Halton sampler with Owen scrambling: builtin/sampler/halton.cpp:9-133, utils/low_discrepancy.cpp:10-120, utils/prime.cpp.
// The reference's wavefront renderer is compiled with IndependentSampler (wavefront/common.h:36-37); "sampler": "halton" in the
// renderer node selects this one here.  Its state rides in the PCG's two words: state = halton index, inc = dimension << 1 (a
// PCG increment is odd), so the SoA layout of the stages is unchanged.
// ------------------------------------------------------------------------------------------
constexpr int PRIME_TABLE_SIZE = 1000;      // utils/prime.h:9
constexpr int MAX_HALTON_RESOLUTION = 128;  // halton.cpp:10

const std::vector<int> &prime_table() {  // the first 1000 primes (utils/prime.cpp lists them; pinned by tests against the reference's array)
    static const std::vector<int> primes = [] {
        std::vector<int> p;
        for (int v = 2; (int)p.size() < PRIME_TABLE_SIZE; ++v) {
            bool ok = true;
            for (int q : p) { if (q * q > v) break; if (v % q == 0) { ok = false; break; } }
            if (ok) p.push_back(v);
        }
        return p;
    }();
    return primes;
}

uint64_t mix_bits(uint64_t v) {  // low_discrepancy.cpp:47-55
    v = v ^ (v >> 31); v = v * 0x7fb5d329728ea185ull;
    v = v ^ (v >> 27); v = v * 0x81dadef4bc2dd44dull;
    v = v ^ (v >> 33);
    return v;
}

int32_t permutation_element(uint32_t i, uint32_t l, uint32_t p) {  // low_discrepancy.cpp:10-43
    uint32_t w = l - 1;
    w = w | (w >> 1); w = w | (w >> 2); w = w | (w >> 4); w = w | (w >> 8); w = w | (w >> 16);
    for (;;) {
        i = i ^ p; i = i * 0xe170893du;
        i = i ^ (p >> 16);
        i = i ^ ((i & w) >> 4);
        i = i ^ (p >> 8); i = i * 0x0929eb3fu;
        i = i ^ (p >> 23);
        i = i ^ ((i & w) >> 1); i = i * (1u | p >> 27);
        i = i * 0x6935fa69u;
        i = i ^ ((i & w) >> 11); i = i * 0x74dcb303u;
        i = i ^ ((i & w) >> 2); i = i * 0x9e501cc3u;
        i = i ^ ((i & w) >> 2); i = i * 0xc860a3dfu;
        i = i & w;
        i = i ^ (i >> 5);
        if (i < l) break;
    }
    return int32_t((i + p) % l);
}

float owen_scrambled_radical_inverse(int base_index, uint64_t a, uint64_t hash) {  // low_discrepancy.cpp:82-103
    const uint64_t base = (uint64_t)prime_table()[base_index];
    const float inv_base = 1.0f / float(base);
    float inv_base_m = 1.0f;
    uint64_t reversed_digits = 0;
    while (1.0f - inv_base_m < 1) {
        uint64_t next = a / base;
        uint64_t digit_value = a - next * base;
        uint64_t digit_hash = mix_bits(hash ^ reversed_digits);
        digit_value = (uint64_t)permutation_element(uint32_t(digit_value), uint32_t(base), uint32_t(digit_hash));
        reversed_digits = reversed_digits * base + digit_value;
        inv_base_m = inv_base_m * inv_base;
        a = next;
    }
    return std::fmin(inv_base_m * float(reversed_digits), 0x1.fffffep-1f);
}

uint64_t inverse_radical_inverse(uint64_t inverse, int base, int digits) {  // low_discrepancy.cpp:105-117
    uint64_t index = 0;
    for (int i = 0; i < digits; ++i) {
        uint64_t digit = inverse % uint64_t(base);
        inverse = inverse / uint64_t(base);
        index = index * uint64_t(base) + digit;
    }
    return index;
}

struct HaltonConsts {  // HaltonSampler::initialize_consts (halton.cpp:116-133)
    int base_scales[2] = {1, 1}, base_exponents[2] = {0, 0}, mult_inverse[2] = {0, 0};
    void initialize(int res_x, int res_y) {
        const int res[2] = {res_x, res_y};
        for (int i = 0; i < 2; ++i) {
            const int base = i == 0 ? 2 : 3;
            int scale = 1, exp = 0;
            while (scale < std::min(res[i], MAX_HALTON_RESOLUTION)) { scale *= base; ++exp; }
            base_scales[i] = scale; base_exponents[i] = exp;
        }
        mult_inverse[0] = (int)multiplicative_inverse(base_scales[1], base_scales[0]);
        mult_inverse[1] = (int)multiplicative_inverse(base_scales[0], base_scales[1]);
    }
    static int64_t mod(int64_t a, int64_t b) { const int64_t r = a - a / b * b; return r < 0 ? r + b : r; }  // halton.cpp:12-16
    static void extended_gcd(uint64_t a, uint64_t b, int64_t *x, int64_t *y) {                                  // halton.cpp:18-30
        if (b == 0) { *x = 1; *y = 0; return; }
        int64_t d = a / b, xp, yp;
        extended_gcd(b, a % b, &xp, &yp);
        *x = yp; *y = xp - (d * yp);
    }
    static uint64_t multiplicative_inverse(int64_t a, int64_t n) { int64_t x, y; extended_gcd(a, n, &x, &y); return mod(x, n); }
};

struct Sampler {  // independent.cpp:11-25, core/sampler.h:22-28; halton.cpp:46-114 when rng.inc is even
    PCG rng;
    void init(uint32_t px, uint32_t py, int sample_index) {
        rng.seed(hash_pixel(px, py));
        rng.advance(int64_t(sample_index) * 65536);
    }
    void init_halton(const HaltonConsts &hc, uint32_t px, uint32_t py, int sample_index) {  // halton.cpp:46-67
        uint64_t halton_index = 0;
        const int sample_stride = hc.base_scales[0] * hc.base_scales[1];
        if (sample_stride > 1) {
            const uint32_t pm[2] = {px % MAX_HALTON_RESOLUTION, py % MAX_HALTON_RESOLUTION};
            for (int i = 0; i < 2; ++i) {
                uint64_t dim_offset = inverse_radical_inverse(uint64_t(pm[i]), i == 0 ? 2 : 3, hc.base_exponents[i]);
                halton_index = halton_index + dim_offset * uint64_t(sample_stride / hc.base_scales[i]) * uint64_t(hc.mult_inverse[i]);
            }
            halton_index = halton_index % uint64_t(sample_stride);
        }
        halton_index = halton_index + uint64_t(int64_t(sample_index * sample_stride));
        rng.state = halton_index;
        rng.inc = uint64_t(2) << 1;  // dimension 2: the first two are skipped (filter importance sampling)
    }
    float get1d() {
        if (rng.inc & 1) return rng.uniform_float();
        uint32_t dimension = uint32_t(rng.inc >> 1);  // halton.cpp:80-89
        if (dimension >= (uint32_t)PRIME_TABLE_SIZE) dimension = 2;
        float r = owen_scrambled_radical_inverse(int(dimension), rng.state, mix_bits(uint64_t(1u + (dimension << 4))));
        rng.inc = uint64_t(dimension + 1) << 1;
        return r;
    }
    V3 get3d() {  // independent: three get1d (core/sampler.h:22-28); halton.cpp:106-118 wraps the dimension before the triple
        if (!(rng.inc & 1) && uint32_t(rng.inc >> 1) + 2 >= (uint32_t)PRIME_TABLE_SIZE) rng.inc = uint64_t(2) << 1;
        float a = get1d(); float b = get1d(); float c = get1d(); return {a, b, c};
    }
};

// ------------------------------------------------------------------------------------------
// Alias table (utils/math/alias.cpp:13-71, utils/cmath/calias.cpp:10-23)
// ------------------------------------------------------------------------------------------
struct AliasUnit { float accept_prob; uint32_t another_idx; };

// 0: sequential sum (MSVC /fp:precise, the reference's toolchain); 1: libstdc++'s std::reduce (groups of four) -- the order
// the reference has when it is compiled HERE (oracle/_ref/libbtrcrefcpu.so), selected by tests that compare against it
static int g_reduce_order = 0;

struct AliasTable {
    std::vector<AliasUnit> units;
    void initialize(const std::vector<float> &probs) {
        float sum = 0.0f;
        // std::reduce(.., 0.0f, plus) leaves the association order unspecified: MSVC, the reference's toolchain, adds sequentially
        // under /fp:precise (libstdc++ adds in groups of four; tests/test_reference_host_pins_cpu.py pins both regimes)
        size_t first = 0;
        if (g_reduce_order == 1)  // libstdc++ <numeric> std::reduce, random-access branch
            for (; probs.size() - first >= 4; first += 4) {
                const float v1 = probs[first] + probs[first + 1], v2 = probs[first + 2] + probs[first + 3];
                const float v3 = v1 + v2;
                sum = sum + v3;
            }
        for (; first < probs.size(); ++first) sum = sum + probs[first];
        const float ratio = probs.size() / sum;
        std::vector<uint32_t> overs, unders;
        units.resize(probs.size());
        for (uint32_t i = 0; i < (uint32_t)units.size(); ++i) {
            const float p = probs[i] * ratio;
            units[i].accept_prob = p;
            units[i].another_idx = i;
            if (p > 1) overs.push_back(i);
            else if (p < 1) unders.push_back(i);
        }
        while (!overs.empty() && !unders.empty()) {
            const uint32_t over = overs.back(), under = unders.back();
            overs.pop_back(); unders.pop_back();
            units[over].accept_prob -= 1 - units[under].accept_prob;
            units[under].another_idx = over;
            if (units[over].accept_prob > 1) overs.push_back(over);
            else if (units[over].accept_prob < 1) unders.push_back(over);
        }
        for (auto i : overs) units[i].accept_prob = 1;
        for (auto i : unders) units[i].accept_prob = 1;
    }
    uint32_t sample(float u) const {  // calias.cpp:14-20
        const uint32_t n = (uint32_t)units.size();
        float nu = float(n) * u;
        uint32_t i = std::min(uint32_t(nu), n - 1);
        float s = nu - float(i);
        return s <= units[i].accept_prob ? i : units[i].another_idx;
    }
};

// ------------------------------------------------------------------------------------------
// intersection_offset (utils/intersection.cpp:5-36)
// ------------------------------------------------------------------------------------------
V3 intersection_offset(V3 pos, V3 nor) {
    constexpr float ORIGIN = 1.0f / 32.0f;
    constexpr float FLOAT_SCALE = 1.0f / 65536.0f;
    constexpr float INT_SCALE = 256.0f;
    int32_t of_i_x = int32_t(2 * INT_SCALE * nor.x);
    int32_t of_i_y = int32_t(2 * INT_SCALE * nor.y);
    int32_t of_i_z = int32_t(2 * INT_SCALE * nor.z);
    float p_i_x = i2f(f2i(pos.x) + (pos.x < 0 ? -of_i_x : of_i_x));
    float p_i_y = i2f(f2i(pos.y) + (pos.y < 0 ? -of_i_y : of_i_y));
    float p_i_z = i2f(f2i(pos.z) + (pos.z < 0 ? -of_i_z : of_i_z));
    float r_x = std::fabs(pos.x) < ORIGIN ? pos.x + FLOAT_SCALE * nor.x : p_i_x;
    float r_y = std::fabs(pos.y) < ORIGIN ? pos.y + FLOAT_SCALE * nor.y : p_i_y;
    float r_z = std::fabs(pos.z) < ORIGIN ? pos.z + FLOAT_SCALE * nor.z : p_i_z;
    return {r_x, r_y, r_z};
}
V3 intersection_offset(V3 pos, V3 nor, V3 next_dir) {  // intersection.cpp:30-34
    return intersection_offset(pos, dot(nor, next_dir) > 0 ? nor : -nor);
}

// ------------------------------------------------------------------------------------------
// Scene description
// ------------------------------------------------------------------------------------------
// Texture slots of a material (creator argument order: diffuse.cpp:36-48, disney.cpp:844-894, metal.cpp:50-77,
// glass.cpp:141-157, mirror.cpp:91-103)
enum MatSlot { SLOT_COLOR = 0, SLOT_METALLIC, SLOT_ROUGHNESS, SLOT_TRANSMISSION, SLOT_IOR, SLOT_SPECULAR, SLOT_SPECULAR_TINT,
               SLOT_ANISOTROPIC, SLOT_SHEEN, SLOT_SHEEN_TINT, SLOT_CLEARCOAT, SLOT_CLEARCOAT_GLOSS, SLOT_TRANSMISSION_ROUGHNESS, SLOT_COUNT };

struct Material {
    int type = MAT_DIFFUSE;
    int tex[SLOT_COUNT];
    int normal_tex = -1;
    bool shadow_terminator = true;
    Material() { for (int &t : tex) t = -1; }
};

enum MediumType { MED_HOMOGENEOUS = 0, MED_HETEROGENEOUS = 1, MED_VOLPRIM = 2 };

struct Medium {
    int type = MED_VOLPRIM;
    float sigma_t = 0; V3 albedo{0, 0, 0}; float g = 0;  // homogeneous
    int sigma_t_tex = -1, albedo_tex = -1, g_tex = -1;   // heterogeneous (3-D textures)
    bool has_transform = false; M4 world_to_local{};     // TransformMedium chain, composed
    float max_density = 0.001f;
};

enum EnvType { ENV_NONE = 0, ENV_GRADIENT_SKY = 1, ENV_IBL = 2 };

struct Mesh {
    int ntri = 0;
    std::vector<V3> pos;  // 3 per tri (local space)
    // per-triangle shading tables (builtin/geometry/triangle_mesh.cpp:40-67), 6 x float4
    std::vector<float> gx_ua, gy_uba, gz_uca, sn_va, sn_vba, sn_vca;  // 4 floats per tri each
    // light sampling tables (triangle_mesh.cpp:108-164)
    std::vector<V3> a_ba_ca;  // 3 per tri
    AliasTable area_alias;
    float total_area = 0;
    V3 bb_lo{MAX_FLOAT, MAX_FLOAT, MAX_FLOAT}, bb_hi{-MAX_FLOAT, -MAX_FLOAT, -MAX_FLOAT};
};

struct Instance {
    int mesh = 0, material = 0, light_id = -1;
    Xf xf;
    uint32_t inner_medium = 0, outer_medium = 0;
    V3 intensity{0, 0, 0};
    float light_scale = 1;
    int tri_offset = 0;  // offset into flattened world triangle list
};

struct Tri {  // lib/bvh/include/bvh/triangle.hpp:27-34 (p0, e1 = p0 - p1, e2 = p2 - p0, n = cross(e1, e2))
    V3 p0, e1, e2, n;
    V3 p1, p2;  // world-space vertices as transformed (kept so tests can hand them to oracle/_ref)
    uint32_t inst, prim;
};

struct BNode { V3 lo, hi; int left, right, first, count; };

struct Hit { bool hit = false; float t = 0, u = 0, v = 0; uint32_t inst = 0, prim = 0; };

struct Stats {
    uint64_t n_gen = 0, n_closest = 0, n_hit = 0, n_next = 0, n_shadow = 0, n_shadow_miss = 0, n_splat = 0;
};

}  // namespace

// core/volume.h:9-45 (VolumePrimitive) + core/volume/bvh.h:24-30 (BVHPrimitive)
struct VolumePrim {
    V3 o, x, y, z;
    V3 x_div_x2, y_div_y2, z_div_z2;     // x / |x|^2 ...            (volume/bvh.cpp:96-102)
    float inv_x2, inv_y2, inv_z2;        // 1 / |x|^2 ...            (volume.cpp:42-49: world_pos_to_uvw)
    int sigma_t_tex = -1, albedo_tex = -1;
    float max_sigma_t = 0;               // Texture3D::get_max_float
};

struct orc_scene {
    TextureSet tex;
    std::vector<Texture3D> tex3d;
    std::vector<Material> materials;
    std::vector<Medium> media;           // id order = Scene::commit order (core/scene.cpp:136-161); last = volume-primitive medium
    bool has_medium = false;             // Scene::has_medium (core/scene.cpp:346-356)
    std::vector<struct VolumePrim> volumes;  // "volumes" of the scene file (factory/scene.cpp:57-80), members of the volume-primitive medium
    // environment light (core/light_sampler.cpp:24-33: appended after the area lights)
    int env_type = ENV_NONE; int env_light_index = -1;
    V3 sky_lower{0, 0, 0}, sky_upper{0, 0, 0}, sky_up{0, 0, 1};
    int ibl_tex = -1; Frame ibl_frame; int lut_w = 0, lut_h = 0;
    std::vector<float> ibl_tile_probs; AliasTable ibl_alias;
    std::vector<Mesh> meshes;
    std::vector<Instance> instances;
    std::vector<int> lights;  // instance index per area-light id
    int n_lights() const { return (int)lights.size() + (env_type != ENV_NONE ? 1 : 0); }
    // camera (builtin/camera/pinhole.cpp:55-77)
    V3 eye, dst, up; float fov_y_deg = 60, w_over_h = 1;
    V3 cam_eye, cam_lbc, cam_fx, cam_fy;
    int width = 64, height = 64;
    // renderer params (builtin/renderer/wavefront.h:13-27)
    int spp = 1, min_depth = 5, max_depth = 10; float rr_threshold = 0.1f, rr_cont_prob = 0.5f;
    bool tile = false, want_albedo = false, want_normal = false;
    bool halton = false; HaltonConsts halton_consts;  // renderer "sampler": "halton"
    int filter_type = 0; float filter_radius = 1.3f, filter_alpha = 1.6f;
    AliasTable gauss_alias;
    // accel
    std::vector<Tri> tris;
    std::vector<BNode> nodes;
    std::vector<int> tri_order;
    V3 bb_lo, bb_hi; float world_diagonal = 0;
    uint32_t vol_prim_medium_id = 0;
    bool committed = false;
};

namespace {

// ------------------------------------------------------------------------------------------
// Mesh tables: utils/triangle_mesh_loader.cpp:61-123, utils/math/triangle.h:8-22,
// utils/math/frame.cpp:45-57, builtin/geometry/triangle_mesh.cpp:40-171
// ------------------------------------------------------------------------------------------
V3 host_frame_from_z_x(V3 z) {  // Frame::from_z(nor).x  (frame.cpp:45-57); Frame's constructor normalises the axes (frame.cpp:11-15) --
                                // pinned against the reference's own code by tests/test_reference_host_pins_cpu.py
    V3 new_z = host_normalize(z), new_y;
    if (1 - std::fabs(new_z.x) < 0.1f) new_y = cross(new_z, V3(0, 1, 0));
    else new_y = cross(new_z, V3(1, 0, 0));
    return host_normalize(cross(new_z, new_y));
}

V3 triangle_dpdu(V3 B_A, V3 C_A, V2 b_a, V2 c_a, V3 nor) {  // triangle.h:8-22
    const float m00 = b_a.x, m01 = b_a.y, m10 = c_a.x, m11 = c_a.y;
    const float det = m00 * m11 - m01 * m10;
    if (det == 0.0f) return host_frame_from_z_x(nor);
    const float inv_det = 1 / det;
    return host_normalize(m11 * inv_det * B_A - m01 * inv_det * C_A);
}

// frame_pos: the positions as read from the file when `pos` was rescaled by transform_to_unit_cube -- the reference computes the
// geometry frames in the loader's constructor, before the rescale (triangle_mesh_loader.cpp:80-123 vs 193-213); null = same as pos
void build_mesh(Mesh &m, int ntri, const float *pos, const float *nor, const uint8_t *has_nor, const float *uv, const float *frame_pos) {
    m.ntri = ntri;
    m.pos.resize(3 * (size_t)ntri);
    for (size_t i = 0; i < 3 * (size_t)ntri; ++i) m.pos[i] = V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    m.gx_ua.resize(4 * (size_t)ntri); m.gy_uba.resize(4 * (size_t)ntri); m.gz_uca.resize(4 * (size_t)ntri);
    m.sn_va.resize(4 * (size_t)ntri); m.sn_vba.resize(4 * (size_t)ntri); m.sn_vca.resize(4 * (size_t)ntri);
    m.a_ba_ca.resize(3 * (size_t)ntri);
    std::vector<float> areas(ntri);
    for (int i = 0; i < ntri; ++i) {
        const V3 a = m.pos[3 * i], b = m.pos[3 * i + 1], c = m.pos[3 * i + 2];
        const float *fp = frame_pos ? frame_pos + 9 * (size_t)i : nullptr;
        const V3 fa = fp ? V3(fp[0], fp[1], fp[2]) : a, fb = fp ? V3(fp[3], fp[4], fp[5]) : b, fc = fp ? V3(fp[6], fp[7], fp[8]) : c;
        const V3 geo_z = host_normalize(cross(fb - fa, fc - fa));  // loader.cpp:86
        V3 sz[3]; V2 tc[3];
        for (int k = 0; k < 3; ++k) {
            const size_t j = 3 * (size_t)i + k;
            if (!has_nor || !has_nor[j]) sz[k] = geo_z;
            else {
                sz[k] = V3(nor[3 * j], nor[3 * j + 1], nor[3 * j + 2]);
                if (sz[k].x == 0.0f && sz[k].y == 0.0f && sz[k].z == 0.0f) sz[k] = geo_z;  // loader.cpp:100-102
            }
            tc[k] = uv ? V2{uv[2 * j], uv[2 * j + 1]} : V2{0, 0};
        }
        const V3 gx = triangle_dpdu(fb - fa, fc - fa, V2{tc[1].x - tc[0].x, tc[1].y - tc[0].y},
                                    V2{tc[2].x - tc[0].x, tc[2].y - tc[0].y}, geo_z);
        const V3 gz = geo_z;
        const V3 gy = host_normalize(cross(gz, gx));  // triangle_mesh.cpp:52
        auto put = [&](std::vector<float> &dst, V3 v, float w) {
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = w;
        };
        put(m.gx_ua, gx, tc[0].x);
        put(m.gy_uba, gy, tc[1].x - tc[0].x);
        put(m.gz_uca, gz, tc[2].x - tc[0].x);
        put(m.sn_va, sz[0], tc[0].y);
        put(m.sn_vba, sz[1] - sz[0], tc[1].y - tc[0].y);
        put(m.sn_vca, sz[2] - sz[0], tc[2].y - tc[0].y);
        const V3 ba = b - a, ca = c - a;
        areas[i] = 0.5f * length(cross(ba, ca));  // triangle.h:24-27
        m.a_ba_ca[3 * i] = a; m.a_ba_ca[3 * i + 1] = ba; m.a_ba_ca[3 * i + 2] = ca;
    }
    m.area_alias.initialize(areas);
    float total = 0.0f;
    for (float a : areas) total += a;  // std::accumulate(.., 0.0f)
    m.total_area = total;
    for (auto &p : m.pos) {
        m.bb_lo = V3(std::fmin(m.bb_lo.x, p.x), std::fmin(m.bb_lo.y, p.y), std::fmin(m.bb_lo.z, p.z));
        m.bb_hi = V3(std::fmax(m.bb_hi.x, p.x), std::fmax(m.bb_hi.y, p.y), std::fmax(m.bb_hi.z, p.z));
    }
}

// ------------------------------------------------------------------------------------------
// Closest hit. Reference: OptiX (closed source) behind wavefront/trace.cpp:38-78; arithmetic oracle =
// lib/bvh Moeller-Trumbore (lib/bvh/include/bvh/triangle.hpp:95-116), double-sided, tmin <= t <= tmax.
// Geometry is flattened to world space: vertex = local_to_world.apply_to_point(v) (transform.cpp:66-72).
// Tie rule (ours, deterministic): smallest t, then smallest (instance, primitive).
// ------------------------------------------------------------------------------------------
inline bool tri_intersect(const Tri &tr, V3 o, V3 d, float tmin, float tmax, float &t, float &u, float &v) {
    V3 c = tr.p0 - o;
    V3 r = cross(d, c);
    float inv_det = 1.0f / dot(tr.n, d);
    u = dot(r, tr.e2) * inv_det;
    v = dot(r, tr.e1) * inv_det;
    float w = 1.0f - u - v;
    if (u >= 0 && v >= 0 && w >= 0) {
        t = dot(tr.n, c) * inv_det;
        if (t >= tmin && t <= tmax) return true;
    }
    return false;
}

inline bool better(float t, uint32_t inst, uint32_t prim, const Hit &h) {
    if (!h.hit) return true;
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst != h.inst) return inst < h.inst;
    return prim < h.prim;
}

Hit closest_hit_brute(const orc_scene &s, V3 o, V3 d, float tmax) {
    Hit best;
    for (const Tri &tr : s.tris) {
        float t, u, v;
        if (tri_intersect(tr, o, d, 0.0f, tmax, t, u, v) && better(t, tr.inst, tr.prim, best)) {
            best.hit = true; best.t = t; best.u = u; best.v = v; best.inst = tr.inst; best.prim = tr.prim;
        }
    }
    return best;
}

int build_node(orc_scene &s, std::vector<V3> &cent, std::vector<V3> &lo, std::vector<V3> &hi, int first, int count) {
    BNode n;
    n.lo = V3(MAX_FLOAT); n.hi = V3(-MAX_FLOAT);
    V3 clo(MAX_FLOAT), chi(-MAX_FLOAT);
    for (int i = first; i < first + count; ++i) {
        int t = s.tri_order[i];
        n.lo = V3(std::fmin(n.lo.x, lo[t].x), std::fmin(n.lo.y, lo[t].y), std::fmin(n.lo.z, lo[t].z));
        n.hi = V3(std::fmax(n.hi.x, hi[t].x), std::fmax(n.hi.y, hi[t].y), std::fmax(n.hi.z, hi[t].z));
        clo = V3(std::fmin(clo.x, cent[t].x), std::fmin(clo.y, cent[t].y), std::fmin(clo.z, cent[t].z));
        chi = V3(std::fmax(chi.x, cent[t].x), std::fmax(chi.y, cent[t].y), std::fmax(chi.z, cent[t].z));
    }
    n.left = n.right = -1; n.first = first; n.count = count;
    int idx = (int)s.nodes.size();
    s.nodes.push_back(n);
    if (count > 4) {
        V3 ext = chi - clo;
        int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
        if (ext[axis] > 0) {
            int mid = first + count / 2;
            std::nth_element(s.tri_order.begin() + first, s.tri_order.begin() + mid, s.tri_order.begin() + first + count,
                             [&](int a, int b) { return cent[a][axis] < cent[b][axis]; });
            int l = build_node(s, cent, lo, hi, first, mid - first);
            int r = build_node(s, cent, lo, hi, mid, first + count - mid);
            s.nodes[idx].left = l; s.nodes[idx].right = r; s.nodes[idx].count = 0;
        }
    }
    return idx;
}

void build_accel(orc_scene &s) {
    s.tris.clear(); s.nodes.clear();
    for (size_t ii = 0; ii < s.instances.size(); ++ii) {
        Instance &inst = s.instances[ii];
        const Mesh &m = s.meshes[inst.mesh];
        inst.tri_offset = (int)s.tris.size();
        for (int i = 0; i < m.ntri; ++i) {
            V3 p0 = inst.xf.point(m.pos[3 * i]), p1 = inst.xf.point(m.pos[3 * i + 1]), p2 = inst.xf.point(m.pos[3 * i + 2]);
            Tri t;
            t.p0 = p0; t.e1 = p0 - p1; t.e2 = p2 - p0; t.n = cross(t.e1, t.e2); t.p1 = p1; t.p2 = p2;
            t.inst = (uint32_t)ii; t.prim = (uint32_t)i;
            s.tris.push_back(t);
        }
    }
    const size_t n = s.tris.size();
    std::vector<V3> cent(n), lo(n), hi(n);
    for (size_t i = 0; i < n; ++i) {
        const Tri &t = s.tris[i];
        V3 p1 = t.p0 - t.e1, p2 = t.p0 + t.e2;
        lo[i] = V3(std::fmin(t.p0.x, std::fmin(p1.x, p2.x)), std::fmin(t.p0.y, std::fmin(p1.y, p2.y)), std::fmin(t.p0.z, std::fmin(p1.z, p2.z)));
        hi[i] = V3(std::fmax(t.p0.x, std::fmax(p1.x, p2.x)), std::fmax(t.p0.y, std::fmax(p1.y, p2.y)), std::fmax(t.p0.z, std::fmax(p1.z, p2.z)));
        // generous conservative padding: the oracle must never cull a triangle the exact test accepts
        V3 ext = hi[i] - lo[i];
        float pad = 1e-4f * std::fmax(std::fmax(std::fabs(lo[i].x), std::fabs(hi[i].x)),
                                      std::fmax(std::fmax(std::fabs(lo[i].y), std::fabs(hi[i].y)),
                                                std::fmax(std::fabs(lo[i].z), std::fabs(hi[i].z)))) +
                    1e-4f * std::fmax(ext.x, std::fmax(ext.y, ext.z)) + 1e-30f;
        lo[i] = lo[i] - V3(pad); hi[i] = hi[i] + V3(pad);
        cent[i] = (lo[i] + hi[i]) * 0.5f;
    }
    s.tri_order.resize(n);
    std::iota(s.tri_order.begin(), s.tri_order.end(), 0);
    if (n) build_node(s, cent, lo, hi, 0, (int)n);
}

inline bool box_hit(const BNode &n, V3 o, V3 inv, float tmax) {
    float t0 = 0.0f, t1 = tmax;
    for (int a = 0; a < 3; ++a) {
        float lo = (n.lo[a] - o[a]) * inv[a], hi = (n.hi[a] - o[a]) * inv[a];
        if (std::isnan(lo) || std::isnan(hi)) continue;  // 0 * inf: origin on slab plane, be conservative
        if (lo > hi) std::swap(lo, hi);
        hi *= 1.0000004f;  // robust (1 + 2*gamma3) widening
        t0 = std::fmax(t0, lo); t1 = std::fmin(t1, hi);
    }
    return t0 <= t1;
}

template <bool ANY>
Hit traverse(const orc_scene &s, V3 o, V3 d, float tmax) {
    Hit best;
    if (s.nodes.empty()) return best;
    V3 inv(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0;
    stack[sp++] = 0;
    float cur_tmax = tmax;
    while (sp) {
        const BNode &n = s.nodes[stack[--sp]];
        if (!box_hit(n, o, inv, cur_tmax)) continue;
        if (n.left < 0) {
            for (int i = n.first; i < n.first + n.count; ++i) {
                const Tri &tr = s.tris[s.tri_order[i]];
                float t, u, v;
                if (tri_intersect(tr, o, d, 0.0f, tmax, t, u, v)) {
                    if (ANY) { best.hit = true; return best; }
                    if (better(t, tr.inst, tr.prim, best)) {
                        best.hit = true; best.t = t; best.u = u; best.v = v; best.inst = tr.inst; best.prim = tr.prim;
                        cur_tmax = t;
                    }
                }
            }
        } else {
            stack[sp++] = n.left; stack[sp++] = n.right;
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------
// Filters (builtin/film_filter/box.cpp:5-10, gaussian.cpp:27-66)
// ------------------------------------------------------------------------------------------
constexpr int GRID_SIZE = 256;
void build_gaussian(orc_scene &s) {
    const float radius = s.filter_radius, alpha = s.filter_alpha;
    const float expv = std::exp(-alpha * radius * radius);
    auto gaussian = [&](float d) { return std::max(0.0f, std::exp(-alpha * d * d) - expv); };
    std::vector<float> table(GRID_SIZE * GRID_SIZE, 0.0f);
    for (int y = 0; y < GRID_SIZE; ++y) {
        const float yf = 2 * radius * (y + 0.5f) / GRID_SIZE - radius;
        for (int x = 0; x < GRID_SIZE; ++x) {
            const float xf = 2 * radius * (x + 0.5f) / GRID_SIZE - radius;
            const float d2 = xf * xf + yf * yf;
            if (d2 < radius * radius) table[y * GRID_SIZE + x] = gaussian(xf) * gaussian(yf);
        }
    }
    s.gauss_alias.initialize(table);
}

V2 filter_sample(const orc_scene &s, Sampler &sampler) {
    if (s.filter_type == 0) {
        float x = sampler.get1d() - 0.5f;
        float y = sampler.get1d() - 0.5f;
        return {x, y};
    }
    const float radius = s.filter_radius;
    uint32_t grid_index = s.gauss_alias.sample(sampler.get1d());
    uint32_t grid_y = grid_index / GRID_SIZE, grid_x = grid_index % GRID_SIZE;
    const float step = 2 * radius / GRID_SIZE;  // host-folded constant in the reference expression
    float x_beg = -radius + step * float(grid_x);
    float x_end = -radius + step * float(grid_x + 1);
    float y_beg = -radius + step * float(grid_y);
    float y_end = -radius + step * float(grid_y + 1);
    float xf = lerp(x_beg, x_end, sampler.get1d());
    float yf = lerp(y_beg, y_end, sampler.get1d());
    return {xf, yf};
}

// ------------------------------------------------------------------------------------------
// Path state, mirroring wavefront/soa.h:8-37
// ------------------------------------------------------------------------------------------
struct PathState {
    V3 o, d; float t1; uint32_t medium_id;
    uint32_t px, py;
    V3 beta; int depth;
    V3 L;
    Sampler sampler;
    V3 beta_le; float bsdf_pdf;
};

struct ShadowRay {
    bool valid = false;
    uint32_t px = 0, py = 0; V3 beta_li; V3 o, d; float t1 = 0; uint32_t medium_id = 0;
};

struct Splat { bool valid = false; V3 value; };

void camera_commit(orc_scene &s) {  // pinhole.cpp:55-77
    const V3 forward = host_normalize(s.dst - s.eye);
    const V3 ex = host_normalize(cross(forward, s.up));
    const V3 ey = host_normalize(cross(forward, ex));
    const float y_len = 2 * std::tan(0.5f * s.fov_y_deg * PI / 180);
    const float x_len = s.w_over_h * y_len;
    s.cam_lbc = s.eye + forward - 0.5f * (x_len * ex + y_len * ey);
    s.cam_fx = x_len * ex;
    s.cam_fy = y_len * ey;
    s.cam_eye = s.eye;
}

// wavefront/generate.cpp:39-86
void generate_path(const orc_scene &s, int64_t accu_state_index, PathState &ps) {
    const int W = s.width, H = s.height;
    int pixel_index, sample_index;
    if (!s.tile) {
        pixel_index = int(accu_state_index % (W * H));
        sample_index = int(accu_state_index / (W * H));
    } else {
        pixel_index = int(accu_state_index / s.spp);
        sample_index = int(accu_state_index % s.spp);
    }
    int pixel_x = pixel_index % W, pixel_y = pixel_index / W;
    if (s.halton) ps.sampler.init_halton(s.halton_consts, (uint32_t)pixel_x, (uint32_t)pixel_y, sample_index);
    else ps.sampler.init((uint32_t)pixel_x, (uint32_t)pixel_y, sample_index);
    V2 fs = filter_sample(s, ps.sampler);
    float pixel_xf = float(pixel_x) + 0.5f + fs.x;
    float pixel_yf = float(pixel_y) + 0.5f + fs.y;
    float film_x = pixel_xf / float(W);
    float film_y = pixel_yf / float(H);
    float time_sample = ps.sampler.get1d();
    (void)time_sample;
    // pinhole.cpp:84-99
    V3 x = film_x * s.cam_fx;
    V3 y = film_y * s.cam_fy;
    V3 dst = s.cam_lbc + x + y;
    V3 dir = normalize(dst - s.cam_eye);
    ps.px = (uint32_t)pixel_x; ps.py = (uint32_t)pixel_y;
    ps.depth = 0; ps.beta = V3(1); ps.L = V3(0);
    ps.o = s.cam_eye; ps.d = dir; ps.t1 = MAX_FLOAT; ps.medium_id = s.vol_prim_medium_id;
    ps.beta_le = V3(1); ps.bsdf_pdf = -1;
}

// core/scene.cpp:9-64
struct SurfacePoint { V3 position; Frame frame; V3 interp_z; V2 uv, tex_coord; };

SurfacePoint get_hitinfo(const orc_scene &s, V3 o, V3 d, const Instance &inst, float t, uint32_t prim, V2 uv) {
    const Mesh &m = s.meshes[inst.mesh];
    SurfacePoint sp;
    sp.position = o + t * d;
    const float *gx = &m.gx_ua[4 * prim], *gy = &m.gy_uba[4 * prim], *gz = &m.gz_uca[4 * prim];
    const float *sa = &m.sn_va[4 * prim], *sba = &m.sn_vba[4 * prim], *sca = &m.sn_vca[4 * prim];
    Frame gf(V3(gx[0], gx[1], gx[2]), V3(gy[0], gy[1], gy[2]), V3(gz[0], gz[1], gz[2]));
    gf.x = inst.xf.vector(gf.x);
    gf.y = inst.xf.vector(gf.y);
    gf.z = inst.xf.normal(gf.z);
    gf.x = normalize(gf.x); gf.y = normalize(gf.y); gf.z = normalize(gf.z);
    V3 interp = V3(sa[0], sa[1], sa[2]) + V3(sba[0], sba[1], sba[2]) * uv.x + V3(sca[0], sca[1], sca[2]) * uv.y;
    interp = normalize(inst.xf.normal(interp));
    float tu = gx[3] + gy[3] * uv.x + gz[3] * uv.y;
    float tv = sa[3] + sba[3] * uv.x + sca[3] * uv.y;
    sp.frame = gf; sp.interp_z = interp; sp.uv = uv; sp.tex_coord = V2{tu, tv};
    return sp;
}

V3 eval_texture(const orc_scene &s, int tex, const SurfacePoint &sp) {  // Texture2D::sample_spectrum(spt): uses spt.tex_coord
    return s.tex.eval2d(tex, sp.tex_coord);
}
float eval_texture_float(const orc_scene &s, int tex, const SurfacePoint &sp) { return eval_texture(s, tex, sp).x; }

V3 eval_texture3d(const orc_scene &s, int tex, V3 uvw) {  // texture3d/array3d.cpp:29-41, core/texture3d.cpp (constant)
    const Texture3D &t = s.tex3d[tex];
    if (t.type == TEX_CONSTANT) return t.value;
    if (t.type == TEX_ADD || t.type == TEX_MUL) {  // texture3d/binary.cpp:17-29
        V3 a = eval_texture3d(s, t.lhs, uvw), b = eval_texture3d(s, t.rhs, uvw);
        return t.type == TEX_ADD ? a + b : a * b;
    }
    float o[4];
    sample_image(s.tex.images[t.image], uvw.x, uvw.y, uvw.z, o);
    return V3(o[0], o[1], o[2]);
}

// ---- environment lights: builtin/light/gradient_sky.cpp:20-54, ibl.cpp:27-78, env_sampler.cpp:124-187 ----------
V3 ibl_eval_local(const orc_scene &s, V3 dir) {  // ibl.cpp:80-87
    float phi = local_angle::phi(dir);
    float theta = local_angle::theta(dir);
    float u = saturate(1 / (2 * PI) * phi);
    float v = saturate(1 / PI * theta);
    return s.tex.eval2d(s.ibl_tex, V2{u, v});
}
V3 env_eval_le(const orc_scene &s, V3 to_light) {
    if (s.env_type == ENV_GRADIENT_SKY) {
        float cos_theta = dot(s.sky_up, normalize(to_light));
        float t = saturate(0.5f * (cos_theta + 1.0f));
        return s.sky_lower * (1.0f - t) + s.sky_upper * t;
    }
    return ibl_eval_local(s, normalize(s.ibl_frame.global_to_local(to_light)));
}
float ibl_pdf_local(const orc_scene &s, V3 to_light) {  // env_sampler.cpp:159-187
    V3 dir = normalize(to_light);
    float cos_theta = dir.z;
    float theta = std::acos(cos_theta);
    float phi = local_angle::phi(dir);
    float u = phi / (2 * PI), v = theta / PI;
    int tile_x = std::min(std::max(int(std::floor(u * s.lut_w)), 0), s.lut_w - 1);
    int tile_y = std::min(std::max(int(std::floor(v * s.lut_h)), 0), s.lut_h - 1);
    int tile_idx = tile_y * s.lut_w + tile_x;
    float tile_pdf = s.ibl_tile_probs[tile_idx];
    float u0 = float(tile_x) / s.lut_w, u1 = float(tile_x + 1) / s.lut_w;
    float v0 = float(tile_y) / s.lut_h, v1 = float(tile_y + 1) / s.lut_h;
    float c0 = std::cos(PI * v0), c1 = std::cos(PI * v1);
    float in_tile_pdf = 1.0f / std::fabs(2 * PI * (u1 - u0) * (c1 - c0));
    return tile_pdf * in_tile_pdf;
}
float env_pdf_li(const orc_scene &s, V3 to_light) {
    if (s.env_type == ENV_GRADIENT_SKY) return pdf_sample_sphere_uniform();
    return ibl_pdf_local(s, s.ibl_frame.global_to_local(to_light));  // ibl.cpp:48-52 (not normalised before the frame)
}
struct EnvSample { V3 dir, radiance; float pdf; };
EnvSample env_sample_li(const orc_scene &s, V3 sam) {
    EnvSample r;
    if (s.env_type == ENV_GRADIENT_SKY) {
        r.dir = sample_sphere_uniform(sam.x, sam.y);
        r.radiance = env_eval_le(s, r.dir);
        r.pdf = pdf_sample_sphere_uniform();
        return r;
    }
    // EnvirLightSampler::sample (env_sampler.cpp:124-157)
    uint32_t tile_idx = s.ibl_alias.sample(sam.x);
    int tile_y = int(tile_idx) / s.lut_w, tile_x = int(tile_idx) % s.lut_w;
    float tile_pdf = s.ibl_tile_probs[tile_idx];
    float u0 = float(tile_x) / s.lut_w, u1 = float(tile_x + 1) / s.lut_w;
    float v0 = float(tile_y) / s.lut_h, v1 = float(tile_y + 1) / s.lut_h;
    float cv0 = std::cos(PI * v0), cv1 = std::cos(PI * v1);
    if (cv0 > cv1) std::swap(cv0, cv1);
    float cos_theta = cv0 + sam.z * (cv1 - cv0);
    float sin_theta = local_angle::cos2sin(cos_theta);
    float phi = 2 * PI * lerp(u0, u1, sam.y);
    V3 dir(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta);
    float in_tile_pdf = 1.0f / (2 * PI * (u1 - u0) * (cv1 - cv0));
    r.dir = s.ibl_frame.local_to_global(dir);
    r.radiance = ibl_eval_local(s, dir);
    r.pdf = tile_pdf * in_tile_pdf;
    return r;
}

// ---- lights: builtin/light/mesh_light.cpp:17-76, geometry/triangle_mesh.cpp:188-228,
//      core/light_sampler.cpp:41-63, wavefront/helper.cpp:10-65,125-142,174-207
struct LiSample { bool success = false; V3 o, d; float t1 = 0, light_pdf = 0; V3 li; };

// surface = true: o = intersection_offset(ref_pos, ref_nor, dir) (sample_surface_light_li); false: o = ref_pos (medium)
LiSample sample_light_li(const orc_scene &s, V3 ref_pos, V3 ref_nor, bool surface, Sampler &sampler) {
    LiSample r;
    V3 sam = sampler.get3d();
    float sel = sampler.get1d();
    const int n = s.n_lights();
    if (n == 0) return r;
    int idx = std::min(int(sel * n), n - 1);
    float select_pdf = 1.0f / float(n);
    if (idx == s.env_light_index) {  // process_envir_light (helper.cpp:36-44)
        EnvSample es = env_sample_li(s, sam);
        r.o = surface ? intersection_offset(ref_pos, ref_nor, es.dir) : ref_pos;
        r.d = es.dir;
        r.t1 = MAX_FLOAT;
        r.li = es.radiance;
        r.light_pdf = select_pdf * es.pdf;
        r.success = true;
        return r;
    }
    const Instance &li = s.instances[s.lights[idx]];
    const Mesh &m = s.meshes[li.mesh];
    // TriangleMesh::sample_inline
    uint32_t prim = m.area_alias.sample(sam.x);
    V2 uv = sample_triangle_uniform(sam.y, sam.z);
    V3 a = m.a_ba_ca[3 * prim], ba = m.a_ba_ca[3 * prim + 1], ca = m.a_ba_ca[3 * prim + 2];
    V3 pos = a + ba * uv.x + ca * uv.y;
    Frame fr(V3(m.gx_ua[4 * prim], m.gx_ua[4 * prim + 1], m.gx_ua[4 * prim + 2]),
             V3(m.gy_uba[4 * prim], m.gy_uba[4 * prim + 1], m.gy_uba[4 * prim + 2]),
             V3(m.gz_uca[4 * prim], m.gz_uca[4 * prim + 1], m.gz_uca[4 * prim + 2]));
    float pdf_area = 1 / m.total_area;
    // MeshLight::sample_li_inline
    V3 spos = li.xf.point(pos);
    V3 snor = normalize(li.xf.normal(fr.z));
    V3 pos_to_ref = ref_pos - spos;
    float dist2 = length_square(pos_to_ref);
    const float inv_scale2 = 1 / (li.light_scale * li.light_scale);
    float pdf_sa = inv_scale2 * pdf_area * dist2 / std::fabs(dot(snor, normalize(pos_to_ref)));
    V3 rad = dot(pos_to_ref, snor) > 0 ? li.intensity : V3(0);
    // helper.cpp:24-34
    V3 diff = spos - ref_pos;
    V3 shadow_dst = intersection_offset(spos, snor, -diff);
    r.o = surface ? intersection_offset(ref_pos, ref_nor, diff) : ref_pos;
    r.d = shadow_dst - r.o;
    r.t1 = 1;
    r.li = rad;
    r.light_pdf = select_pdf * pdf_sa;
    r.success = true;
    return r;
}

// helper.cpp:144-172
V3 eval_miss_le(const orc_scene &s, V3 o, V3 d, V3 beta_le, float bsdf_pdf) {
    V3 result(0);
    if (s.env_type != ENV_NONE) {
        V3 le = env_eval_le(s, d);
        if (bsdf_pdf < 0) result = beta_le * le / -bsdf_pdf;
        else {
            float select_light_pdf = 1.0f / float(s.n_lights());
            float envir_light_pdf = env_pdf_li(s, d);
            float light_pdf = select_light_pdf * envir_light_pdf;
            result = beta_le * le / (bsdf_pdf + light_pdf);
        }
    }
    return result;
}

V3 handle_intersected_light(const orc_scene &s, V3 o, V3 d, const SurfacePoint &inct, V3 beta_le, float bsdf_pdf, int light_index) {
    V3 result(0);
    if (light_index >= 0) {
        const int n = s.n_lights();
        float select_pdf = 1.0f / float(n);
        const Instance &li = s.instances[s.lights[light_index]];
        const Mesh &m = s.meshes[li.mesh];
        V3 le = dot(inct.frame.z, -d) >= 0 ? li.intensity : V3(0);
        if (bsdf_pdf < 0) result = beta_le * le / -bsdf_pdf;
        else {
            float light_dir_pdf;
            V3 nor = inct.frame.z, pos = inct.position;
            if (dot(nor, o - pos) <= 0) light_dir_pdf = 0;
            else {
                float pdf_area = 1 / m.total_area;
                V3 spt_to_ref = o - pos;
                float dist2 = length_square(spt_to_ref);
                float dist3 = std::sqrt(dist2) * dist2;
                const float inv_scale2 = 1 / (li.light_scale * li.light_scale);
                light_dir_pdf = inv_scale2 * pdf_area * dist3 / std::fabs(dot(nor, spt_to_ref));
            }
            float light_pdf = select_pdf * light_dir_pdf;
            result = beta_le * le / (bsdf_pdf + light_pdf);
        }
    }
    return result;
}

// wavefront/helper.cpp:67-95
bool simple_russian_roulette(const orc_scene &s, V3 &beta, int depth, Sampler &sampler) {
    bool ret = false;
    if (depth >= s.min_depth) {
        if (depth >= s.max_depth) ret = true;
        else {
            float sam = sampler.get1d();
            if (lum(beta) < s.rr_threshold) {
                if (sam > s.rr_cont_prob) ret = true;
                else beta = beta / s.rr_cont_prob;
            }
        }
    }
    return ret;
}

// Material::create_shader: diffuse.cpp:21-34, disney.cpp:813-842, metal.cpp:31-48, glass.cpp:120-139, mirror.cpp:78-89,
// black.cpp:50-53, invisible.cpp:46-51; normal map: material/utils/normal_map.cpp:13-21
Shader create_shader(const orc_scene &s, const Material &mat, const SurfacePoint &inct) {
    Shader sh;
    sh.type = mat.type;
    sh.frame.geometry = inct.frame;
    if (mat.type == MAT_BLACK || mat.type == MAT_INVISIBLE) { sh.frame.shading = inct.frame; return sh; }
    sh.frame.shading = inct.frame.rotate_to_new_z(inct.interp_z);
    if (mat.normal_tex >= 0) {
        V3 c = eval_texture(s, mat.normal_tex, inct);
        V3 local_nor = 2.0f * c - V3(1.0f);
        V3 world_nor = sh.frame.shading.local_to_global(local_nor);
        sh.frame.shading = sh.frame.shading.rotate_to_new_z(normalize(world_nor));
    }
    sh.shadow_terminator = mat.shadow_terminator;
    auto F = [&](int slot) { return eval_texture_float(s, mat.tex[slot], inct); };
    switch (mat.type) {
    case MAT_DIFFUSE: sh.diffuse.albedo_value = eval_texture(s, mat.tex[SLOT_COLOR], inct); break;
    case MAT_METAL:
        sh.metal = MicrofacetReflection(eval_texture(s, mat.tex[SLOT_COLOR], inct), F(SLOT_ROUGHNESS), F(SLOT_ANISOTROPIC));
        break;
    case MAT_DISNEY: {
        V3 base_color = eval_texture(s, mat.tex[SLOT_COLOR], inct);
        float metallic = F(SLOT_METALLIC), roughness = F(SLOT_ROUGHNESS), transmission = F(SLOT_TRANSMISSION);
        float transmission_roughness = F(SLOT_TRANSMISSION_ROUGHNESS), ior = F(SLOT_IOR), specular = F(SLOT_SPECULAR);
        float specular_tint = F(SLOT_SPECULAR_TINT), anisotropic = F(SLOT_ANISOTROPIC), sheen = F(SLOT_SHEEN), sheen_tint = F(SLOT_SHEEN_TINT);
        float clearcoat = F(SLOT_CLEARCOAT), clearcoat_gloss = F(SLOT_CLEARCOAT_GLOSS);
        sh.disney = DisneyBSDF(base_color, metallic, roughness, specular, specular_tint, anisotropic, sheen, sheen_tint, clearcoat,
                               clearcoat_gloss, transmission, transmission_roughness, ior);
        break;
    }
    case MAT_GLASS: sh.color = eval_texture(s, mat.tex[SLOT_COLOR], inct); sh.ior = F(SLOT_IOR); break;
    case MAT_MIRROR: sh.color = eval_texture(s, mat.tex[SLOT_COLOR], inct); break;
    default: break;
    }
    return sh;
}

struct ShadeOut {
    bool has_next = false; PathState next;
    ShadowRay shadow;
    Splat radiance;                      // terminal / Le-only splat
    bool has_gbuffer = false; V3 albedo, normal;
};

// wavefront/shade.cpp:38-235.  `scattered` = bit 31 of the intersection flag, set by the medium stage.
void shade(const orc_scene &s, const PathState &in, const Hit &hit, bool scattered, ShadeOut &out) {
    PathState path = in;
    Sampler &sampler = path.sampler;
    if (!hit.hit) {
        V3 splat = eval_miss_le(s, in.o, in.d, in.beta_le, in.bsdf_pdf);
        if (!scattered) splat = splat + path.L;
        out.radiance.valid = true; out.radiance.value = splat;
        return;
    }
    const Instance &inst = s.instances[hit.inst];
    SurfacePoint inct = get_hitinfo(s, in.o, in.d, inst, hit.t, hit.prim, V2{hit.u, hit.v});
    V3 le_rad = handle_intersected_light(s, in.o, in.d, inct, in.beta_le, in.bsdf_pdf, inst.light_id);
    if (scattered) { out.radiance.valid = true; out.radiance.value = le_rad; return; }
    path.L = path.L + le_rad;
    if (simple_russian_roulette(s, path.beta, path.depth, sampler)) {
        out.radiance.valid = true; out.radiance.value = path.L;
        return;
    }
    V3 shadow_bsdf_val(0); float shadow_bsdf_pdf = 0;
    LiSample ls = sample_light_li(s, inct.position, inct.frame.z, true, sampler);
    Shader shader = create_shader(s, s.materials[inst.material], inct);
    if (path.depth == 0) { out.has_gbuffer = true; out.albedo = shader.albedo(); out.normal = shader.normal(); }
    V3 wo = -in.d;
    BsdfSample bs = shader.sample(wo, sampler.get3d());
    if (ls.success) {
        shadow_bsdf_val = shader.eval(ls.d, wo);
        ls.success = !is_zero(shadow_bsdf_val);
        if (ls.success) shadow_bsdf_pdf = shader.pdf(ls.d, wo);
    }
    if (ls.success) {
        bool shadow_d_out = dot(inct.frame.z, ls.d) > 0;
        bool last_ray_out = dot(inct.frame.z, in.d) < 0;
        uint32_t med;
        if (shadow_d_out == last_ray_out) med = in.medium_id;
        else if (shadow_d_out) med = inst.outer_medium;
        else med = inst.inner_medium;
        float cosv = std::fabs(dot(normalize(ls.d), inct.frame.z));
        V3 beta_li = ls.li * path.beta * shadow_bsdf_val * cosv / (shadow_bsdf_pdf + ls.light_pdf);
        out.shadow.valid = true; out.shadow.px = path.px; out.shadow.py = path.py;
        out.shadow.beta_li = beta_li; out.shadow.o = ls.o; out.shadow.d = ls.d; out.shadow.t1 = ls.t1; out.shadow.medium_id = med;
    }
    if (!is_zero(bs.bsdf)) {
        bs.dir = normalize(bs.dir);
        float cosv = std::fabs(dot(bs.dir, inct.frame.z));
        V3 new_beta_le = path.beta * cosv * bs.bsdf;
        path.beta = new_beta_le / bs.pdf;
        V3 next_o = intersection_offset(inct.position, inct.frame.z, bs.dir);
        bool next_out = dot(inct.frame.z, bs.dir) > 0;
        uint32_t next_med = next_out ? inst.outer_medium : inst.inner_medium;
        float stored_pdf = bs.is_delta ? -bs.pdf : bs.pdf;
        out.has_next = true;
        out.next = path;
        out.next.depth = path.depth + 1;
        out.next.beta_le = new_beta_le; out.next.bsdf_pdf = stored_pdf;
        out.next.o = next_o; out.next.d = bs.dir; out.next.t1 = MAX_FLOAT; out.next.medium_id = next_med;
    } else {
        out.radiance.valid = true; out.radiance.value = path.L;
    }
}

// ---- media: builtin/medium/homogeneous.cpp:25-73, hetergeneous.cpp:25-106, core/medium.cpp:32-129 -----------------
float henyey_greenstein(float g, float u) {  // utils/henyey_greenstein.h:10-18
    float g2 = g * g;
    float dem = 1.0f + g2 - 2.0f * (g * u);
    return (1.0f - g2) / (4.0f * PI * dem * std::sqrt(dem));
}
float cos_between(V3 a, V3 b) { return dot(normalize(a), normalize(b)); }  // cvec3.h:182-185

struct PhaseShader {  // HenyeyGreensteinPhaseShader (core/medium.cpp:22-76)
    float g = 0; V3 color{0, 0, 0};
    V3 eval(V3 wi, V3 wo) const { float f = henyey_greenstein(g, -cos_between(wi, wo)); return color * V3(f, f, f); }
    float pdf(V3 wi, V3 wo) const { return henyey_greenstein(g, -cos_between(wi, wo)); }
    void sample(V3 wo, V3 sam, V3 &phase, V3 &dir, float &pdf_out) const {
        float sv = sam.x + sam.x - 1;
        float u;
        if (std::fabs(g) < 0.001f) u = sv;
        else { float g2 = g * g; u = (1.0f + g2 - sqr((1.0f - g2) / (1.0f + g * sv))) / (g + g); }
        float cos_theta = -u;
        float sin_theta = local_angle::cos2sin(cos_theta);
        float phi = 2 * PI * sam.y;
        V3 local_wi(sin_theta * std::sin(phi), sin_theta * std::cos(phi), cos_theta);
        float phase_val = henyey_greenstein(g, u);
        phase = color * V3(phase_val, phase_val, phase_val);
        dir = Frame::from_z(wo).local_to_global(local_wi);
        pdf_out = phase_val;
    }
};

V3 medium_uvw(const Medium &m, V3 p) {  // TransformMedium (core/medium.cpp:98-129): world_to_local.apply_to_point
    if (!m.has_transform) return p;
    const M4 &w = m.world_to_local;
    return {w.m[0][0] * p.x + w.m[0][1] * p.y + w.m[0][2] * p.z + w.m[0][3],
            w.m[1][0] * p.x + w.m[1][1] * p.y + w.m[1][2] * p.z + w.m[1][3],
            w.m[2][0] * p.x + w.m[2][1] * p.y + w.m[2][2] * p.z + w.m[2][3]};
}

struct MediumSample { bool scattered = false; V3 position; V3 throughput{1, 1, 1}; PhaseShader shader; };

// ------------------------------------------------------------------------------------------
// Volume primitives: core/volume.cpp:112-195 (VolumePrimitiveMedium::sample / tr), core/volume/bvh.cpp:146-326,
// core/volume/aggregate.cpp:129-253.
// The reference walks a small BVH over the primitives' boxes and picks per-overlap generated code through a trie
// (volume/indexing.cpp, overlap.cpp); both only select WHICH primitives contain a point.  The results they produce are
// order independent (closest boundary = a minimum, overlap = a set), so this restatement loops over the primitives.
// Two documented choices where the reference depends on incidental order: the members of an overlap are visited in
// creation order (the reference iterates a std::set ordered by object address), and when more than
// MAX_OVERLAP_COUNT = 4 primitives contain a point the four with the smallest ids are kept (the reference keeps the
// first four in ITS BVH's traversal order).
// ------------------------------------------------------------------------------------------
constexpr int MAX_OVERLAP_COUNT = 4;  // volume/bvh.h:11

V3 vol_world_to_uvw(const VolumePrim &v, V3 p) {  // volume.cpp:42-49
    V3 op = p - v.o;
    return V3(dot(op, v.x) * v.inv_x2, dot(op, v.y) * v.inv_y2, dot(op, v.z) * v.inv_z2);
}

// volume/bvh.cpp:281-326: nearest boundary crossing (entry, else exit) of one primitive on p + t d, t in (0, t_max); -1 if none
float vol_prim_intersect(V3 p, V3 d, float t_max, const VolumePrim &prim) {
    V3 op = p - prim.o;
    float t0 = 0.0f, t1 = t_max;
    const V3 axes[3] = {prim.x_div_x2, prim.y_div_y2, prim.z_div_z2};
    for (const V3 &a : axes) {
        float opd = dot(op, a), dd = dot(d, a);
        float nd = (0.0f - opd) / dd, fd = (1.0f - opd) / dd;
        t0 = std::fmax(t0, std::fmin(nd, fd));
        t1 = std::fmin(t1, std::fmax(nd, fd));
    }
    if (t0 > t1) return -1;
    if (t0 > 0) return t0;
    if (t1 < t_max) return t1;
    return -1;
}

// volume/bvh.cpp:146-214
bool vol_find_closest_intersection(const orc_scene &s, V3 a, V3 b, V3 &out) {
    V3 dir = b - a;
    float final_t = MAX_FLOAT, t_max = 1.0f;
    for (const VolumePrim &prim : s.volumes) {
        float t = vol_prim_intersect(a, dir, t_max, prim);
        if (t >= 0.0f) { final_t = std::fmin(final_t, t); t_max = final_t; }
    }
    if (final_t > 10) return false;
    out = a * (1.0f - final_t) + b * final_t;
    return true;
}

struct VolOverlap { uint32_t data[MAX_OVERLAP_COUNT]; int count = 0; };

VolOverlap vol_get_overlap(const orc_scene &s, V3 pos) {  // volume/bvh.cpp:216-263, is_in_prim :270-279
    VolOverlap r;
    for (size_t i = 0; i < s.volumes.size() && r.count < MAX_OVERLAP_COUNT; ++i) {
        const VolumePrim &prim = s.volumes[i];
        V3 op = pos - prim.o;
        float u = dot(op, prim.x_div_x2), v = dot(op, prim.y_div_y2), w = dot(op, prim.z_div_z2);
        if (0.0f <= u && u <= 1.0f && 0.0f <= v && v <= 1.0f && 0.0f <= w && w <= 1.0f) r.data[r.count++] = (uint32_t)i;
    }
    return r;
}

float vol_overlap_max_density(const orc_scene &s, const VolOverlap &ov) {  // aggregate.cpp:121-127
    float m = 0.0f;
    for (int i = 0; i < ov.count; ++i) m += s.volumes[ov.data[i]].max_sigma_t;
    return std::fmax(m, 0.01f);
}

// aggregate.cpp:129-199
void vol_sample_scattering(const orc_scene &s, const VolOverlap &ov, V3 a, V3 b, Sampler &sampler, MediumSample &r) {
    const float inv_max_density = 1.0f / vol_overlap_max_density(s, ov);
    V3 uvw_a[MAX_OVERLAP_COUNT], uvw_b[MAX_OVERLAP_COUNT];
    for (int i = 0; i < ov.count; ++i) {
        uvw_a[i] = vol_world_to_uvw(s.volumes[ov.data[i]], a);
        uvw_b[i] = vol_world_to_uvw(s.volumes[ov.data[i]], b);
    }
    float t_max = length(b - a), t = 0.0f;
    for (;;) {
        float dt = -std::log(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) { r.scattered = false; return; }
        float tf = t / t_max;
        V3 uvws[MAX_OVERLAP_COUNT]; float densities[MAX_OVERLAP_COUNT];
        float density_sum = 0.0f;
        for (int i = 0; i < ov.count; ++i) {
            uvws[i] = uvw_a[i] * (1.0f - tf) + uvw_b[i] * tf;
            densities[i] = eval_texture3d(s, s.volumes[ov.data[i]].sigma_t_tex, uvws[i]).x;
            density_sum = density_sum + densities[i];
        }
        if (sampler.get1d() < density_sum * inv_max_density) {
            r.scattered = true;
            r.position = a * (1.0f - tf) + b * tf;
            V3 albedo(0, 0, 0);
            for (int i = 0; i < ov.count; ++i) {
                float weight = densities[i] / density_sum;
                albedo = albedo + weight * eval_texture3d(s, s.volumes[ov.data[i]].albedo_tex, uvws[i]);
            }
            r.shader.g = 0.0f; r.shader.color = albedo;
            return;
        }
    }
}

// aggregate.cpp:201-253
float vol_tr_in_overlap(const orc_scene &s, const VolOverlap &ov, V3 a, V3 b, Sampler &sampler) {
    const float inv_max_density = 1.0f / vol_overlap_max_density(s, ov);
    V3 uvw_a[MAX_OVERLAP_COUNT], uvw_b[MAX_OVERLAP_COUNT];
    for (int i = 0; i < ov.count; ++i) {
        uvw_a[i] = vol_world_to_uvw(s.volumes[ov.data[i]], a);
        uvw_b[i] = vol_world_to_uvw(s.volumes[ov.data[i]], b);
    }
    float result = 1.0f, t_max = length(b - a), t = 0.0f;
    for (;;) {
        float dt = -std::log(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        float tf = t / t_max;
        float density_sum = 0.0f;
        for (int i = 0; i < ov.count; ++i) {
            V3 uvw = uvw_a[i] * (1.0f - tf) + uvw_b[i] * tf;
            density_sum = density_sum + eval_texture3d(s, s.volumes[ov.data[i]].sigma_t_tex, uvw).x;
        }
        result = result * (1.0f - density_sum * inv_max_density);
    }
    return result;
}

MediumSample volprim_sample(const orc_scene &s, V3 a, V3 b, Sampler &sampler) {  // volume.cpp:112-160
    MediumSample r;
    if (s.volumes.empty()) return r;
    V3 o = a, d = normalize(b - a);
    for (;;) {
        if (dot(b - o, d) <= 0.0f || length_square(b - o) < 0.00001f) { r.scattered = false; break; }
        V3 inct_pos;
        if (!vol_find_closest_intersection(s, o, b, inct_pos)) inct_pos = b;
        VolOverlap ov = vol_get_overlap(s, 0.5f * (o + inct_pos));
        if (ov.count != 0) {
            vol_sample_scattering(s, ov, o, inct_pos, sampler, r);
            if (r.scattered) break;
        }
        o = inct_pos + 0.001f * d;
    }
    return r;
}

V3 volprim_tr(const orc_scene &s, V3 a, V3 b, Sampler &sampler) {  // volume.cpp:162-195
    float result = 1.0f;
    if (s.volumes.empty()) return V3(1, 1, 1);
    V3 o = a, d = normalize(b - a);
    for (;;) {
        if (dot(b - o, d) <= 0.0f || length_square(b - o) < 0.0001f) break;
        V3 inct_pos;
        if (!vol_find_closest_intersection(s, o, b, inct_pos)) inct_pos = b;
        VolOverlap ov = vol_get_overlap(s, 0.5f * (o + inct_pos));
        if (ov.count != 0) result = result * vol_tr_in_overlap(s, ov, o, inct_pos, sampler);
        o = inct_pos + 0.001f * d;
    }
    return V3(result, result, result);
}

MediumSample medium_sample(const orc_scene &s, const Medium &m, V3 a, V3 b, Sampler &sampler) {
    MediumSample r;
    if (m.type == MED_VOLPRIM) return volprim_sample(s, a, b, sampler);
    if (m.type == MED_HOMOGENEOUS) {  // homogeneous.cpp:25-58
        float t_max = length(b - a);
        float st = -std::log(1.0f - sampler.get1d()) / m.sigma_t;
        if (st < t_max) {
            float t = st / t_max;
            r.scattered = true;
            r.position = a * (1.0f - t) + b * t;
            r.shader.g = m.g; r.shader.color = m.albedo;
        }
        return r;
    }
    // hetergeneous.cpp:25-73 (delta tracking against the global majorant)
    const float inv_max_density = 1.0f / m.max_density;
    float t_max = length(b - a), t = 0.0f;
    V3 local_a = V3(0.5f) + 0.5f * medium_uvw(m, a);
    V3 local_b = V3(0.5f) + 0.5f * medium_uvw(m, b);
    V3 local_ba_div_t_max = (local_b - local_a) / t_max;
    V3 ba_div_t_max = (b - a) / t_max;
    for (;;) {
        float dt = -std::log(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        V3 uvw = local_a + t * local_ba_div_t_max;
        float density = eval_texture3d(s, m.sigma_t_tex, uvw).x;
        if (sampler.get1d() < density * inv_max_density) {
            r.scattered = true;
            r.position = a + t * ba_div_t_max;
            r.shader.color = eval_texture3d(s, m.albedo_tex, uvw);
            r.shader.g = eval_texture3d(s, m.g_tex, uvw).x;
            break;
        }
    }
    return r;
}

V3 medium_tr(const orc_scene &s, const Medium &m, V3 a, V3 b, Sampler &sampler) {
    if (m.type == MED_VOLPRIM) return volprim_tr(s, a, b, sampler);
    if (m.type == MED_HOMOGENEOUS) {  // homogeneous.cpp:60-73: sigma_a = sigma_t * (1 - albedo)
        V3 sigma_a = m.sigma_t * (V3(1, 1, 1) - m.albedo);
        V3 e = -length(a - b) * sigma_a;
        return V3(std::exp(e.x), std::exp(e.y), std::exp(e.z));
    }
    // hetergeneous.cpp:75-106 (ratio tracking)
    float result = 1.0f, t = 0.0f, t_max = length(b - a);
    const float inv_max_density = 1.0f / m.max_density;
    V3 local_a = V3(0.5f) + 0.5f * medium_uvw(m, a);
    V3 local_b = V3(0.5f) + 0.5f * medium_uvw(m, b);
    V3 local_ba_div_t_max = (local_b - local_a) / t_max;
    for (;;) {
        float dt = -std::log(1.0f - sampler.get1d()) * inv_max_density;
        t = t + dt;
        if (t >= t_max) break;
        V3 uvw = local_a + t * local_ba_div_t_max;
        float density = eval_texture3d(s, m.sigma_t_tex, uvw).x;
        result = result * (1.0f - density * inv_max_density);
    }
    return V3(result, result, result);
}

struct MediumOut {
    bool scattered = false;
    bool has_next = false; PathState next;
    ShadowRay shadow;
    Splat radiance;
};

// wavefront/medium.cpp:36-164 + get_medium_segment (:214-249).  `ps` is updated in place the way the kernel updates its
// slot: resolved medium id, beta_le *= unscatter_tr, and (only when not scattered) the advanced sampler.
void medium_stage(const orc_scene &s, PathState &ps, const Hit &hit, MediumOut &out) {
    PathState path = ps;
    Sampler &sampler = path.sampler;
    uint32_t medium_id; V3 medium_end;
    if (hit.hit) {
        const Instance &inst = s.instances[hit.inst];
        const Mesh &m = s.meshes[inst.mesh];
        V3 local_normal(m.gz_uca[4 * hit.prim], m.gz_uca[4 * hit.prim + 1], m.gz_uca[4 * hit.prim + 2]);
        V3 inct_nor = inst.xf.normal(local_normal);
        uint32_t inct_medium_id = dot(ps.d, inct_nor) < 0 ? inst.outer_medium : inst.inner_medium;
        medium_id = std::min(inct_medium_id, ps.medium_id);  // resolve_mediums (wavefront/medium.cpp:12-15): smaller id = higher priority
        medium_end = intersection_offset(ps.o + hit.t * ps.d, inct_nor, -ps.d);
    } else {
        medium_id = s.vol_prim_medium_id;
        medium_end = ps.o + normalize(ps.d) * s.world_diagonal;
    }
    ps.medium_id = medium_id;
    const Medium &medium = s.media[medium_id];
    MediumSample ms = medium_sample(s, medium, ps.o, medium_end, sampler);
    V3 unscatter_tr = medium_tr(s, medium, ps.o, medium_end, sampler);
    ps.beta_le = ps.beta_le * unscatter_tr;
    if (!ms.scattered) { ps.sampler = sampler; return; }
    out.scattered = true;
    path.beta = path.beta * ms.throughput;
    if (simple_russian_roulette(s, path.beta, path.depth, sampler)) {
        out.radiance.valid = true; out.radiance.value = path.L;
        return;
    }
    LiSample ls = sample_light_li(s, ms.position, V3(0), false, sampler);
    if (ls.success) {
        V3 shadow_phase_val = ms.shader.eval(ls.d, -ps.d);
        float shadow_phase_pdf = ms.shader.pdf(ls.d, -ps.d);
        ls.li = ls.li * path.beta * shadow_phase_val / (shadow_phase_pdf + ls.light_pdf);
    }
    V3 phase, pdir; float ppdf;
    ms.shader.sample(-ps.d, sampler.get3d(), phase, pdir, ppdf);
    if (ls.success) {
        out.shadow.valid = true; out.shadow.px = path.px; out.shadow.py = path.py;
        out.shadow.beta_li = ls.li; out.shadow.o = ls.o; out.shadow.d = ls.d; out.shadow.t1 = ls.t1;
        out.shadow.medium_id = ls.t1 > 1 ? s.vol_prim_medium_id : medium_id;
    }
    if (!is_zero(phase)) {
        pdir = normalize(pdir);
        V3 beta_le = path.beta * phase;
        path.beta = beta_le / ppdf;
        out.has_next = true;
        out.next = path;
        out.next.depth = path.depth + 1;
        out.next.beta_le = beta_le; out.next.bsdf_pdf = ppdf;
        out.next.o = ms.position; out.next.d = pdir; out.next.t1 = MAX_FLOAT; out.next.medium_id = medium_id;
    } else {
        out.radiance.valid = true; out.radiance.value = path.L;
    }
}

// Shadow-ray transmittance sampler.  The reference keeps one PCG per shadow-queue SLOT (soa_buffer.cpp:86-96), so its
// media renders are not reproducible run to run.  Both this oracle and the CUDA path key the stream on
// the ray itself instead: seed = murmur64A over the 32 bytes (o.xyz, medium id, d.xyz, t1).  Same distribution.
Sampler shadow_sampler(const ShadowRay &r) {
    uint8_t buf[32];
    float f[8] = {r.o.x, r.o.y, r.o.z, u2f(r.medium_id), r.d.x, r.d.y, r.d.z, r.t1};
    std::memcpy(buf, f, 32);
    Sampler sm;
    sm.rng.seed(murmur_hash64A(buf, 32, 0));
    return sm;
}

// shadow.cpp:39-74: transmittance of an unoccluded shadow ray
V3 shadow_tr(const orc_scene &s, const ShadowRay &r) {
    if (!s.has_medium) return V3(1, 1, 1);
    Sampler sm = shadow_sampler(r);
    if (r.t1 > 1) return medium_tr(s, s.media[s.vol_prim_medium_id], r.o, r.o + normalize(r.d) * s.world_diagonal, sm);
    V3 end_pnt = r.o + r.d * r.t1;
    return medium_tr(s, s.media[r.medium_id], r.o, end_pnt, sm);
}

// core/film.cpp:131-165 (bounds + isfinite checks)
inline void film_add(float *rgba, V3 v) { if (isfinite3(v)) { rgba[0] += v.x; rgba[1] += v.y; rgba[2] += v.z; } }

}  // namespace

// ==========================================================================================
// C ABI (ctypes)
// ==========================================================================================
extern "C" {

orc_scene *orc_scene_new() { return new orc_scene(); }
void orc_scene_free(orc_scene *s) { delete s; }

int orc_add_texture_constant(orc_scene *s, float r, float g, float b) {
    Texture2D t; t.type = TEX_CONSTANT; t.value = V3(r, g, b);
    s->tex.tex2d.push_back(t);
    return (int)s->tex.tex2d.size() - 1;
}

// Decoded image (2-D when depth == 0).  data: w*h*max(d,1)*channels floats, already normalised to [0,1] for 8-bit sources;
// srgb != 0 applies the sRGB->linear decode to the first three channels (cudaTextureDesc::sRGB; 8-bit 4-channel only).
int orc_add_image(orc_scene *s, int w, int h, int d, int channels, const float *data, int srgb, int addr_u, int addr_v, int addr_w,
                  int linear, const float *border) {
    TexImage im; im.w = w; im.h = h; im.d = d; im.channels = channels;
    const size_t n = (size_t)w * h * (d > 0 ? d : 1);
    im.data.assign(data, data + n * channels);
    // Array::find_minmax_values runs on the raw texel data (before any sRGB decode)
    for (int c = 0; c < 3; ++c) { im.max_value[c] = -MAX_FLOAT; im.min_value[c] = MAX_FLOAT; }
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < std::min(channels, 3); ++c) {
            im.max_value[c] = std::fmax(im.max_value[c], im.data[i * channels + c]);
            im.min_value[c] = std::fmin(im.min_value[c], im.data[i * channels + c]);
        }
    if (srgb && channels == 4)
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) im.data[4 * i + c] = srgb_to_linear(im.data[4 * i + c]);
    im.addr[0] = addr_u; im.addr[1] = addr_v; im.addr[2] = addr_w; im.linear = linear != 0;
    if (border) for (int c = 0; c < 4; ++c) im.border[c] = border[c];
    s->tex.images.push_back(std::move(im));
    return (int)s->tex.images.size() - 1;
}

// Array::find_minmax_values of an integer format scales the integer extrema by float(1 / max) (array.cpp:470-490), which is not
// always the float the texel itself becomes (c / max): front ends that decode such grids hand the reference's extrema over.
void orc_set_image_minmax(orc_scene *s, int image, const float *min3, const float *max3) {
    for (int c = 0; c < 3; ++c) { s->tex.images[image].min_value[c] = min3[c]; s->tex.images[image].max_value[c] = max3[c]; }
}

// storage format of the image's texels as far as the texture unit's border colour is concerned (orc_texture.h border_component)
void orc_set_image_border_levels(orc_scene *s, int image, int levels) { s->tex.images[image].border_levels = levels; }

int orc_add_texture_array(orc_scene *s, int image) {
    Texture2D t; t.type = TEX_ARRAY; t.image = image;
    s->tex.tex2d.push_back(t);
    return (int)s->tex.tex2d.size() - 1;
}

// op: TEX_ADD / TEX_SUB / TEX_MUL / TEX_DIV
int orc_add_texture_binary(orc_scene *s, int op, int lhs, int rhs) {
    Texture2D t; t.type = op; t.lhs = lhs; t.rhs = rhs;
    s->tex.tex2d.push_back(t);
    return (int)s->tex.tex2d.size() - 1;
}

int orc_add_texture_transform(orc_scene *s, int inner, const float *m6) {
    Texture2D t; t.type = TEX_TRANSFORM; t.lhs = inner;
    for (int i = 0; i < 6; ++i) t.m[i] = m6[i];
    s->tex.tex2d.push_back(t);
    return (int)s->tex.tex2d.size() - 1;
}

void orc_eval_texture2d(orc_scene *s, int tex, int n, const float *uv, float *out_rgb) {
    for (int i = 0; i < n; ++i) {
        V3 v = s->tex.eval2d(tex, V2{uv[2 * i], uv[2 * i + 1]});
        out_rgb[3 * i] = v.x; out_rgb[3 * i + 1] = v.y; out_rgb[3 * i + 2] = v.z;
    }
}

int orc_add_texture3d_constant(orc_scene *s, float r, float g, float b) {
    Texture3D t; t.type = TEX_CONSTANT; t.value = V3(r, g, b);
    s->tex3d.push_back(t);
    return (int)s->tex3d.size() - 1;
}

int orc_add_texture3d_array(orc_scene *s, int image) {
    Texture3D t; t.type = TEX_ARRAY; t.image = image;
    s->tex3d.push_back(t);
    return (int)s->tex3d.size() - 1;
}

// Texture3DBinaryOperatorCreator (texture3d/binary.cpp:150-160): op 0 = "add", 1 = "mul"
int orc_add_texture3d_binary(orc_scene *s, int op, int lhs, int rhs) {
    Texture3D t; t.type = op == 0 ? TEX_ADD : TEX_MUL; t.lhs = lhs; t.rhs = rhs;
    s->tex3d.push_back(t);
    return (int)s->tex3d.size() - 1;
}

// Texture3D::get_min_float / get_max_float: constant (core/texture3d.cpp), array (texel extrema of the grid), add / mul by
// interval arithmetic on the operands' extrema (texture3d/binary.cpp:95-137).  which = 0: min, 1: max.
float orc_texture3d_extremum(const orc_scene *s, int tex, int which) {
    const Texture3D &t = s->tex3d[tex];
    if (t.type == TEX_CONSTANT) return t.value.x;
    if (t.type == TEX_ARRAY) return which ? s->tex.images[t.image].max_value[0] : s->tex.images[t.image].min_value[0];
    const float lmax = orc_texture3d_extremum(s, t.lhs, 1), lmin = orc_texture3d_extremum(s, t.lhs, 0);
    const float rmax = orc_texture3d_extremum(s, t.rhs, 1), rmin = orc_texture3d_extremum(s, t.rhs, 0);
    if (t.type == TEX_ADD) return which ? lmax + rmax : lmin + rmin;
    if (which) return std::max(std::max(lmax * rmax, lmax * rmin), std::max(lmin * rmax, lmin * rmin));
    return std::min(std::min(lmax * rmax, lmax * rmin), std::min(lmin * rmax, lmin * rmin));
}

// tex13: texture id per MatSlot (-1 where the material type has no such parameter)
int orc_add_material(orc_scene *s, int type, const int *tex13, int normal_tex, int shadow_terminator) {
    Material m; m.type = type; m.normal_tex = normal_tex; m.shadow_terminator = shadow_terminator != 0;
    for (int i = 0; i < SLOT_COUNT; ++i) m.tex[i] = tex13[i];
    s->materials.push_back(m);
    return (int)s->materials.size() - 1;
}

// Media are added in final id order (the front-end sorts by priority like Scene::commit); the volume-primitive medium is
// appended by orc_commit.  Returns the medium id.
int orc_add_medium_homogeneous(orc_scene *s, float sigma_t, const float *albedo, float g) {
    Medium m; m.type = MED_HOMOGENEOUS; m.sigma_t = sigma_t; m.albedo = V3(albedo[0], albedo[1], albedo[2]); m.g = g;
    s->media.push_back(m);
    return (int)s->media.size() - 1;
}

int orc_add_medium_heterogeneous(orc_scene *s, int sigma_t_tex, int albedo_tex, int g_tex, const float *world_to_local16) {
    Medium m; m.type = MED_HETEROGENEOUS; m.sigma_t_tex = sigma_t_tex; m.albedo_tex = albedo_tex; m.g_tex = g_tex;
    if (world_to_local16) { m.has_transform = true; std::memcpy(m.world_to_local.m, world_to_local16, 64); }
    const float mx = orc_texture3d_extremum(s, sigma_t_tex, 1);
    m.max_density = std::fmax(mx, 0.001f);  // hetergeneous.cpp:33
    s->media.push_back(m);
    return (int)s->media.size() - 1;
}

// factory/scene.cpp:57-80 ("volumes": o, x, y, z, sigma_t, albedo)
int orc_add_volume(orc_scene *s, const float *o, const float *x, const float *y, const float *z, int sigma_t_tex, int albedo_tex) {
    VolumePrim v;
    v.o = V3(o[0], o[1], o[2]); v.x = V3(x[0], x[1], x[2]); v.y = V3(y[0], y[1], y[2]); v.z = V3(z[0], z[1], z[2]);
    const float x2 = dot(v.x, v.x), y2 = dot(v.y, v.y), z2 = dot(v.z, v.z);
    v.x_div_x2 = v.x / x2; v.y_div_y2 = v.y / y2; v.z_div_z2 = v.z / z2;
    v.inv_x2 = 1.0f / x2; v.inv_y2 = 1.0f / y2; v.inv_z2 = 1.0f / z2;
    v.sigma_t_tex = sigma_t_tex; v.albedo_tex = albedo_tex;
    v.max_sigma_t = orc_texture3d_extremum(s, sigma_t_tex, 1);
    s->volumes.push_back(v);
    s->has_medium = true;  // core/scene.cpp:346-349
    return (int)s->volumes.size() - 1;
}

// IBL + EnvirLightSampler::preprocess (light/ibl.cpp:27-36,89-102; light/env_sampler.cpp:22-122): per-tile mean luminance from
// 256 Hammersley taps (utils/math/hammersley.h), tile solid angle, table = max(0, area * (lum - avg)), normalised; alias table.
void orc_set_ibl(orc_scene *s, int tex, const float *up, int lut_w, int lut_h) {
    s->env_type = ENV_IBL;
    s->ibl_tex = tex;
    {  // Frame::from_z (utils/math/frame.cpp:45-57), host normalisation
        V3 nz = host_normalize(V3(up[0], up[1], up[2])), ny;
        if (1 - std::fabs(nz.x) < 0.1f) ny = cross(nz, V3(0, 1, 0));
        else ny = cross(nz, V3(1, 0, 0));
        V3 nx = cross(nz, ny);
        s->ibl_frame.x = host_normalize(nx); s->ibl_frame.y = host_normalize(ny); s->ibl_frame.z = host_normalize(nz);
    }
    s->lut_w = lut_w; s->lut_h = lut_h;
    const int n_samples = 256;
    auto radical_inverse = [](uint32_t bits) {
        bits = (bits << 16u) | (bits >> 16u);
        bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
        bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
        bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
        bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
        return bits * 2.3283064365386963e-10f;
    };
    std::vector<float> lum((size_t)lut_w * lut_h), area((size_t)lut_w * lut_h);
#pragma omp parallel for
    for (int yi = 0; yi < lut_h; ++yi)
        for (int xi = 0; xi < lut_w; ++xi) {
            float x0 = float(xi) / lut_w, x1 = float(xi + 1) / lut_w, y0 = float(yi) / lut_h, y1 = float(yi + 1) / lut_h;
            float lum_sum = 0.0f;
            for (int i = 0; i < n_samples; ++i) {
                V2 h{float(i) / n_samples, radical_inverse((uint32_t)i)};
                V2 ls{1.2f * h.x - 0.1f, 1.2f * h.y - 0.1f};
                float x = lerp(x0, x1, ls.x), y = lerp(y0, y1, ls.y);
                lum_sum = lum_sum + orc::lum(s->tex.eval2d(tex, V2{x, y}));
            }
            lum[(size_t)yi * lut_w + xi] = lum_sum / n_samples;
            area[(size_t)yi * lut_w + xi] = std::fabs(2 * PI * (x1 - x0) * (std::cos(PI * y1) - std::cos(PI * y0)));
        }
    float lum_area_sum = 0.0f, area_sum = 0.0f;
    for (size_t i = 0; i < lum.size(); ++i) { lum_area_sum += lum[i] * area[i]; area_sum += area[i]; }
    const float avg_lum = lum_area_sum / area_sum;
    for (size_t i = 0; i < lum.size(); ++i) lum[i] = std::max(0.0f, area[i] * (lum[i] - avg_lum));
    float table_sum = 0.0f;
    for (float v : lum) table_sum += v;
    if (table_sum > 1e-3f) { const float ratio = 1 / table_sum; for (auto &v : lum) v *= ratio; }
    s->ibl_tile_probs = lum;
    s->ibl_alias.initialize(lum);
}

void orc_get_ibl_table(orc_scene *s, float *probs) { std::memcpy(probs, s->ibl_tile_probs.data(), sizeof(float) * s->ibl_tile_probs.size()); }

void orc_set_gradient_sky(orc_scene *s, const float *lower, const float *upper, const float *up) {
    s->env_type = ENV_GRADIENT_SKY;
    s->sky_lower = V3(lower[0], lower[1], lower[2]); s->sky_upper = V3(upper[0], upper[1], upper[2]);
    s->sky_up = host_normalize(V3(up[0], up[1], up[2]));  // gradient_sky.cpp:15-18
}

int orc_add_mesh(orc_scene *s, int ntri, const float *pos, const float *nor, const uint8_t *has_nor, const float *uv, const float *frame_pos) {
    s->meshes.emplace_back();
    build_mesh(s->meshes.back(), ntri, pos, nor, has_nor, uv, frame_pos);
    return (int)s->meshes.size() - 1;
}

// inner_medium / outer_medium: medium id, or -1 for "no medium" (= the volume-primitive medium)
int orc_add_instance(orc_scene *s, int mesh, int material, const float *mat16, const float *inv16,
                     int has_light, const float *intensity, int inner_medium, int outer_medium) {
    Instance in; in.mesh = mesh; in.material = material;
    in.inner_medium = (uint32_t)inner_medium; in.outer_medium = (uint32_t)outer_medium;
    if (inner_medium >= 0 || outer_medium >= 0) s->has_medium = true;
    std::memcpy(in.xf.mat.m, mat16, 64); std::memcpy(in.xf.inv.m, inv16, 64);
    if (has_light) {
        in.light_id = (int)s->lights.size();
        in.intensity = V3(intensity[0], intensity[1], intensity[2]);
        // MeshLight::set_geometry (mesh_light.cpp:11-15)
        in.light_scale = length(in.xf.point(V3(1, 0, 0)) - in.xf.point(V3(0, 0, 0)));
        s->lights.push_back((int)s->instances.size());
    }
    s->instances.push_back(in);
    return (int)s->instances.size() - 1;
}

void orc_set_camera(orc_scene *s, const float *eye, const float *dst, const float *up, float fov_y_deg) {
    s->eye = V3(eye[0], eye[1], eye[2]); s->dst = V3(dst[0], dst[1], dst[2]); s->up = V3(up[0], up[1], up[2]);
    s->fov_y_deg = fov_y_deg;
}

void orc_set_film(orc_scene *s, int w, int h) { s->width = w; s->height = h; s->w_over_h = float(w) / h; }

void orc_set_renderer(orc_scene *s, int spp, int min_depth, int max_depth, float rr_threshold, float rr_cont_prob,
                      int tile, int albedo, int normal, int filter_type, float radius, float alpha) {
    s->spp = spp; s->min_depth = min_depth; s->max_depth = max_depth; s->rr_threshold = rr_threshold;
    s->rr_cont_prob = rr_cont_prob; s->tile = tile != 0; s->want_albedo = albedo != 0; s->want_normal = normal != 0;
    s->filter_type = filter_type; s->filter_radius = radius; s->filter_alpha = alpha;
}

void orc_set_sampler(orc_scene *s, int halton) { s->halton = halton != 0; }

// known-answer access to the Halton machinery
float orc_owen_scrambled_radical_inverse(int base_index, uint64_t a, uint64_t hash) { return owen_scrambled_radical_inverse(base_index, a, hash); }
int orc_prime(int i) { return prime_table()[i]; }

void orc_commit(orc_scene *s) {
    s->halton_consts.initialize(s->width, s->height);
    camera_commit(*s);
    if (s->filter_type == 1) build_gaussian(*s);
    build_accel(*s);
    // Scene::commit bbox (core/scene.cpp:224-237) + wavefront.cpp:138-139
    V3 lo(MAX_FLOAT), hi(-MAX_FLOAT);
    for (auto &inst : s->instances) {
        const Mesh &m = s->meshes[inst.mesh];
        for (int c = 0; c < 8; ++c) {
            V3 p(c & 4 ? m.bb_hi.x : m.bb_lo.x, c & 2 ? m.bb_hi.y : m.bb_lo.y, c & 1 ? m.bb_hi.z : m.bb_lo.z);
            V3 q = inst.xf.point(p);
            lo = V3(std::fmin(lo.x, q.x), std::fmin(lo.y, q.y), std::fmin(lo.z, q.z));
            hi = V3(std::fmax(hi.x, q.x), std::fmax(hi.y, q.y), std::fmax(hi.z, q.z));
        }
    }
    for (const VolumePrim &v : s->volumes) {  // core/scene.cpp:233-237, core/volume.cpp:51-66 (box of o, o+x, o+y, o+z)
        const V3 c[4] = {v.o, v.o + v.x, v.o + v.y, v.o + v.z};
        for (const V3 &q : c) {
            lo = V3(std::fmin(lo.x, q.x), std::fmin(lo.y, q.y), std::fmin(lo.z, q.z));
            hi = V3(std::fmax(hi.x, q.x), std::fmax(hi.y, q.y), std::fmax(hi.z, q.z));
        }
    }
    lo = V3(std::fmin(lo.x, s->eye.x), std::fmin(lo.y, s->eye.y), std::fmin(lo.z, s->eye.z));
    hi = V3(std::fmax(hi.x, s->eye.x), std::fmax(hi.y, s->eye.y), std::fmax(hi.z, s->eye.z));
    s->bb_lo = lo; s->bb_hi = hi;
    s->world_diagonal = 1.2f * length(hi - lo);
    // medium ids (scene.cpp:136-161): real media in priority order, then the volume-primitive medium
    s->vol_prim_medium_id = (uint32_t)s->media.size();
    s->media.emplace_back();  // MED_VOLPRIM
    for (auto &inst : s->instances) {
        if ((int)inst.inner_medium < 0) inst.inner_medium = s->vol_prim_medium_id;
        if ((int)inst.outer_medium < 0) inst.outer_medium = s->vol_prim_medium_id;
    }
    s->env_light_index = s->env_type != ENV_NONE ? (int)s->lights.size() : -1;
    s->committed = true;
}

int orc_num_triangles(orc_scene *s) { return (int)s->tris.size(); }
float orc_world_diagonal(orc_scene *s) { return s->world_diagonal; }
uint32_t orc_vol_prim_medium_id(orc_scene *s) { return s->vol_prim_medium_id; }

// Flattened world-space triangles as (p0, p1, p2) 9 floats + (inst, prim): lets tests feed the very
// same triangles to oracle/_ref (lib/bvh).
void orc_get_triangles(orc_scene *s, float *verts, uint32_t *inst_prim) {
    for (size_t i = 0; i < s->tris.size(); ++i) {
        const Tri &t = s->tris[i];
        float *o = verts + 9 * i;
        o[0] = t.p0.x; o[1] = t.p0.y; o[2] = t.p0.z; o[3] = t.p1.x; o[4] = t.p1.y; o[5] = t.p1.z; o[6] = t.p2.x; o[7] = t.p2.y; o[8] = t.p2.z;
        inst_prim[2 * i] = t.inst; inst_prim[2 * i + 1] = t.prim;
    }
}

// Per-mesh shading tables (6 x float4 per triangle) for table parity tests.
void orc_get_mesh_tables(orc_scene *s, int mesh, float *out24) {
    const Mesh &m = s->meshes[mesh];
    for (int i = 0; i < m.ntri; ++i) {
        float *o = out24 + 24 * (size_t)i;
        std::memcpy(o, &m.gx_ua[4 * i], 16); std::memcpy(o + 4, &m.gy_uba[4 * i], 16); std::memcpy(o + 8, &m.gz_uca[4 * i], 16);
        std::memcpy(o + 12, &m.sn_va[4 * i], 16); std::memcpy(o + 16, &m.sn_vba[4 * i], 16); std::memcpy(o + 20, &m.sn_vca[4 * i], 16);
    }
}

// Closest hit on a ray batch in the reference SoA layout (soa.cpp:5-23,55-91): o_med/d_t1 are float4,
// flag = inst | bit30 on hit, detail = (t bits, prim, u bits, v bits).  mode 0 = BVH, 1 = brute force.
void orc_closest_hit(orc_scene *s, int n, const float *o_med, const float *d_t1, uint32_t *flag, uint32_t *detail, int mode) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        V3 o(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]), d(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]);
        float t1 = d_t1[4 * i + 3];
        Hit h = mode == 1 ? closest_hit_brute(*s, o, d, t1) : traverse<false>(*s, o, d, t1);
        if (!h.hit) { flag[i] = 0; detail[4 * i] = detail[4 * i + 1] = detail[4 * i + 2] = detail[4 * i + 3] = 0; }
        else {
            flag[i] = h.inst | (1u << 30);
            detail[4 * i] = f2u(h.t); detail[4 * i + 1] = h.prim; detail[4 * i + 2] = f2u(h.u); detail[4 * i + 3] = f2u(h.v);
        }
    }
}

// Occlusion test: out[i] = 1 if any triangle has t in [0, t1] (shadow.cpp:24-36).
void orc_any_hit(orc_scene *s, int n, const float *o_med, const float *d_t1, uint8_t *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        V3 o(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]), d(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]);
        out[i] = traverse<true>(*s, o, d, d_t1[4 * i + 3]).hit ? 1 : 0;
    }
}

static void store_state(const PathState &ps, int i, float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth,
                        float *path_rad, uint64_t *rng, float *ble) {
    o_med[4 * i] = ps.o.x; o_med[4 * i + 1] = ps.o.y; o_med[4 * i + 2] = ps.o.z; o_med[4 * i + 3] = u2f(ps.medium_id);
    d_t1[4 * i] = ps.d.x; d_t1[4 * i + 1] = ps.d.y; d_t1[4 * i + 2] = ps.d.z; d_t1[4 * i + 3] = ps.t1;
    pixel[2 * i] = ps.px; pixel[2 * i + 1] = ps.py;
    beta_depth[4 * i] = ps.beta.x; beta_depth[4 * i + 1] = ps.beta.y; beta_depth[4 * i + 2] = ps.beta.z; beta_depth[4 * i + 3] = i2f(ps.depth);
    path_rad[4 * i] = ps.L.x; path_rad[4 * i + 1] = ps.L.y; path_rad[4 * i + 2] = ps.L.z; path_rad[4 * i + 3] = 1;
    rng[2 * i] = ps.sampler.rng.state; rng[2 * i + 1] = ps.sampler.rng.inc;
    ble[4 * i] = ps.beta_le.x; ble[4 * i + 1] = ps.beta_le.y; ble[4 * i + 2] = ps.beta_le.z; ble[4 * i + 3] = ps.bsdf_pdf;
}

static void load_state(PathState &ps, int i, const float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth,
                       const float *path_rad, const uint64_t *rng, const float *ble) {
    ps.o = V3(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]); ps.medium_id = f2u(o_med[4 * i + 3]);
    ps.d = V3(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]); ps.t1 = d_t1[4 * i + 3];
    ps.px = pixel[2 * i]; ps.py = pixel[2 * i + 1];
    ps.beta = V3(beta_depth[4 * i], beta_depth[4 * i + 1], beta_depth[4 * i + 2]); ps.depth = f2i(beta_depth[4 * i + 3]);
    ps.L = V3(path_rad[4 * i], path_rad[4 * i + 1], path_rad[4 * i + 2]);
    ps.sampler.rng.state = rng[2 * i]; ps.sampler.rng.inc = rng[2 * i + 1];
    ps.beta_le = V3(ble[4 * i], ble[4 * i + 1], ble[4 * i + 2]); ps.bsdf_pdf = ble[4 * i + 3];
}

// generate stage in the reference SoA layout (generate.cpp:84-86)
void orc_generate(orc_scene *s, int64_t first, int n, float *o_med, float *d_t1, uint32_t *pixel, float *beta_depth,
                  float *path_rad, uint64_t *rng, float *ble) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        PathState ps;
        generate_path(*s, first + i, ps);
        store_state(ps, i, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble);
    }
}

// shade stage on a batch (uncompacted outputs: slot i of every output belongs to input i).
// status bits: 1 = next state emitted, 2 = shadow ray emitted, 4 = radiance splat, 8 = g-buffer splat
void orc_shade(orc_scene *s, int n,
               const float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth, const float *path_rad,
               const uint64_t *rng, const float *ble, const uint32_t *flag, const uint32_t *detail,
               uint32_t *status,
               float *n_o_med, float *n_d_t1, uint32_t *n_pixel, float *n_beta_depth, float *n_path_rad, uint64_t *n_rng, float *n_ble,
               float *sh_o_med, float *sh_d_t1, float *sh_beta_li,
               float *splat_rgb, float *gbuf_albedo, float *gbuf_normal) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        PathState ps;
        load_state(ps, i, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble);
        Hit h;
        h.hit = (flag[i] >> 30) & 1;
        if (h.hit) { h.inst = flag[i] & 0x3fffffffu; h.t = u2f(detail[4 * i]); h.prim = detail[4 * i + 1]; h.u = u2f(detail[4 * i + 2]); h.v = u2f(detail[4 * i + 3]); }
        ShadeOut out;
        shade(*s, ps, h, ((flag[i] >> 31) & 1) != 0, out);
        uint32_t st = 0;
        if (out.has_next) { st |= 1; store_state(out.next, i, n_o_med, n_d_t1, n_pixel, n_beta_depth, n_path_rad, n_rng, n_ble); }
        if (out.shadow.valid) {
            st |= 2;
            sh_o_med[4 * i] = out.shadow.o.x; sh_o_med[4 * i + 1] = out.shadow.o.y; sh_o_med[4 * i + 2] = out.shadow.o.z; sh_o_med[4 * i + 3] = u2f(out.shadow.medium_id);
            sh_d_t1[4 * i] = out.shadow.d.x; sh_d_t1[4 * i + 1] = out.shadow.d.y; sh_d_t1[4 * i + 2] = out.shadow.d.z; sh_d_t1[4 * i + 3] = out.shadow.t1;
            sh_beta_li[4 * i] = out.shadow.beta_li.x; sh_beta_li[4 * i + 1] = out.shadow.beta_li.y; sh_beta_li[4 * i + 2] = out.shadow.beta_li.z; sh_beta_li[4 * i + 3] = 1;
        }
        if (out.radiance.valid) { st |= 4; splat_rgb[3 * i] = out.radiance.value.x; splat_rgb[3 * i + 1] = out.radiance.value.y; splat_rgb[3 * i + 2] = out.radiance.value.z; }
        if (out.has_gbuffer) {
            st |= 8;
            gbuf_albedo[3 * i] = out.albedo.x; gbuf_albedo[3 * i + 1] = out.albedo.y; gbuf_albedo[3 * i + 2] = out.albedo.z;
            gbuf_normal[3 * i] = out.normal.x; gbuf_normal[3 * i + 1] = out.normal.y; gbuf_normal[3 * i + 2] = out.normal.z;
        }
        status[i] = st;
    }
}

// medium stage on a batch (wavefront/medium.cpp:36-164).  In-place outputs mirror what the kernel writes back to its own
// slot: o_med.w (resolved medium id), ble (beta_le * unscatter_tr), flag (bit 31 when scattered), rng (advanced only when
// not scattered).  Uncompacted per-slot outputs like orc_shade: status bits 1 = next state, 2 = shadow ray, 4 = radiance splat.
void orc_medium(orc_scene *s, int n,
                float *o_med, const float *d_t1, const uint32_t *pixel, const float *beta_depth, const float *path_rad,
                uint64_t *rng, float *ble, uint32_t *flag, const uint32_t *detail,
                uint32_t *status,
                float *n_o_med, float *n_d_t1, uint32_t *n_pixel, float *n_beta_depth, float *n_path_rad, uint64_t *n_rng, float *n_ble,
                float *sh_o_med, float *sh_d_t1, float *sh_beta_li, float *splat_rgb) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        PathState ps;
        load_state(ps, i, o_med, d_t1, pixel, beta_depth, path_rad, rng, ble);
        Hit h;
        h.hit = (flag[i] >> 30) & 1;
        if (h.hit) { h.inst = flag[i] & 0x3fffffffu; h.t = u2f(detail[4 * i]); h.prim = detail[4 * i + 1]; h.u = u2f(detail[4 * i + 2]); h.v = u2f(detail[4 * i + 3]); }
        MediumOut out;
        medium_stage(*s, ps, h, out);
        o_med[4 * i + 3] = u2f(ps.medium_id);
        ble[4 * i] = ps.beta_le.x; ble[4 * i + 1] = ps.beta_le.y; ble[4 * i + 2] = ps.beta_le.z;
        uint32_t st = 0;
        if (out.scattered) flag[i] |= 1u << 31;
        else { rng[2 * i] = ps.sampler.rng.state; rng[2 * i + 1] = ps.sampler.rng.inc; }
        if (out.has_next) { st |= 1; store_state(out.next, i, n_o_med, n_d_t1, n_pixel, n_beta_depth, n_path_rad, n_rng, n_ble); }
        if (out.shadow.valid) {
            st |= 2;
            sh_o_med[4 * i] = out.shadow.o.x; sh_o_med[4 * i + 1] = out.shadow.o.y; sh_o_med[4 * i + 2] = out.shadow.o.z; sh_o_med[4 * i + 3] = u2f(out.shadow.medium_id);
            sh_d_t1[4 * i] = out.shadow.d.x; sh_d_t1[4 * i + 1] = out.shadow.d.y; sh_d_t1[4 * i + 2] = out.shadow.d.z; sh_d_t1[4 * i + 3] = out.shadow.t1;
            sh_beta_li[4 * i] = out.shadow.beta_li.x; sh_beta_li[4 * i + 1] = out.shadow.beta_li.y; sh_beta_li[4 * i + 2] = out.shadow.beta_li.z; sh_beta_li[4 * i + 3] = 1;
        }
        if (out.radiance.valid) { st |= 4; splat_rgb[3 * i] = out.radiance.value.x; splat_rgb[3 * i + 1] = out.radiance.value.y; splat_rgb[3 * i + 2] = out.radiance.value.z; }
        status[i] = st;
    }
}

// Transmittance factor applied to unoccluded shadow rays (shadow.cpp:39-74); tr_rgb = 1 in scenes without media.
void orc_shadow_tr(orc_scene *s, int n, const float *o_med, const float *d_t1, float *tr_rgb) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        ShadowRay r;
        r.o = V3(o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2]); r.medium_id = f2u(o_med[4 * i + 3]);
        r.d = V3(d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2]); r.t1 = d_t1[4 * i + 3];
        V3 tr = shadow_tr(*s, r);
        tr_rgb[3 * i] = tr.x; tr_rgb[3 * i + 1] = tr.y; tr_rgb[3 * i + 2] = tr.z;
    }
}

int orc_has_medium(orc_scene *s) { return s->has_medium ? 1 : 0; }

// BSDF known-answer access: evaluates one material at a synthetic hit (geometry frame = identity axes, shading normal
// `interp_z`, texture coordinate uv) for n (wo, wi, sam) triples.  out_* per triple: sample {bsdf rgb, dir xyz, pdf, is_delta},
// eval rgb, pdf.
void orc_bsdf_probe(orc_scene *s, int material, const float *interp_z, const float *uv, int n, const float *wo, const float *wi,
                    const float *sam, float *out_sample8, float *out_eval3, float *out_pdf, float *out_albedo3, float *out_normal3) {
    SurfacePoint sp;
    sp.position = V3(0);
    sp.frame = Frame(V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1));
    sp.interp_z = normalize(V3(interp_z[0], interp_z[1], interp_z[2]));
    sp.uv = V2{0, 0}; sp.tex_coord = V2{uv[0], uv[1]};
    Shader sh = create_shader(*s, s->materials[material], sp);
    V3 al = sh.albedo(), nr = sh.normal();
    out_albedo3[0] = al.x; out_albedo3[1] = al.y; out_albedo3[2] = al.z;
    out_normal3[0] = nr.x; out_normal3[1] = nr.y; out_normal3[2] = nr.z;
    for (int i = 0; i < n; ++i) {
        V3 o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), sm(sam[3 * i], sam[3 * i + 1], sam[3 * i + 2]);
        BsdfSample bs = sh.sample(o, sm);
        float *p = out_sample8 + 8 * i;
        p[0] = bs.bsdf.x; p[1] = bs.bsdf.y; p[2] = bs.bsdf.z; p[3] = bs.dir.x; p[4] = bs.dir.y; p[5] = bs.dir.z; p[6] = bs.pdf; p[7] = bs.is_delta ? 1.0f : 0.0f;
        V3 e = sh.eval(w, o);
        out_eval3[3 * i] = e.x; out_eval3[3 * i + 1] = e.y; out_eval3[3 * i + 2] = e.z;
        out_pdf[i] = sh.pdf(w, o);
    }
}

// Whole render (wavefront.cpp:205-325 + preview.cpp:28-80).  Renders sample indices [spp_begin, spp_end) of every
// pixel; out_* are W*H float4 (resolved: value/weight, w = 1).  raw_* (optional) receive unresolved sums
// radiance(float4 stride) / weight(float).  stats (optional): 7 x uint64 event counters.
void orc_render(orc_scene *s, int spp_begin, int spp_end, float *out_color, float *out_albedo, float *out_normal,
                float *raw_radiance, float *raw_weight, uint64_t *stats_out) {
    const int W = s->width, H = s->height;
    const size_t npix = (size_t)W * H;
    std::vector<float> rad(4 * npix, 0.0f), wgt(npix, 0.0f), alb(4 * npix, 0.0f), nor(4 * npix, 0.0f);
    Stats total;
#pragma omp parallel
    {
        Stats st;
#pragma omp for schedule(dynamic, 64)
        for (long long pix = 0; pix < (long long)npix; ++pix) {
            for (int sidx = spp_begin; sidx < spp_end; ++sidx) {
                // state index for Uniform mode = sidx * W*H + pix, for Tile mode = pix * spp + sidx
                int64_t accu = s->tile ? (int64_t)pix * s->spp + sidx : (int64_t)sidx * (int64_t)npix + pix;
                PathState ps;
                generate_path(*s, accu, ps);
                wgt[pix] += 1.0f; st.n_gen++;
                const size_t p = (size_t)ps.py * W + ps.px;
                auto do_shadow = [&](const ShadowRay &sr) {  // shadow.cpp:24-74
                    st.n_shadow++;
                    bool occluded = traverse<true>(*s, sr.o, sr.d, sr.t1).hit;
                    if (!occluded) { st.n_shadow_miss++; st.n_splat++; film_add(&rad[4 * p], sr.beta_li * shadow_tr(*s, sr)); }
                };
                for (;;) {
                    Hit h = traverse<false>(*s, ps.o, ps.d, ps.t1);
                    st.n_closest++; if (h.hit) st.n_hit++;
                    bool scattered = false;
                    MediumOut mo;
                    if (s->has_medium) {  // wavefront.cpp:254-268
                        medium_stage(*s, ps, h, mo);
                        scattered = mo.scattered;
                        if (mo.shadow.valid) do_shadow(mo.shadow);
                        if (mo.radiance.valid) { st.n_splat++; film_add(&rad[4 * p], mo.radiance.value); }
                    }
                    ShadeOut out;
                    shade(*s, ps, h, scattered, out);
                    if (out.shadow.valid) do_shadow(out.shadow);
                    if (out.has_gbuffer) {
                        if (s->want_albedo) film_add(&alb[4 * p], out.albedo);
                        if (s->want_normal) film_add(&nor[4 * p], out.normal);
                    }
                    if (out.radiance.valid) { st.n_splat++; film_add(&rad[4 * p], out.radiance.value); }
                    if (scattered) {
                        if (!mo.has_next) break;
                        st.n_next++;
                        ps = mo.next;
                        continue;
                    }
                    if (!out.has_next) break;
                    st.n_next++;
                    ps = out.next;
                }
            }
        }
#pragma omp critical
        {
            total.n_gen += st.n_gen; total.n_closest += st.n_closest; total.n_hit += st.n_hit; total.n_next += st.n_next;
            total.n_shadow += st.n_shadow; total.n_shadow_miss += st.n_shadow_miss; total.n_splat += st.n_splat;
        }
    }
    auto resolve = [&](const std::vector<float> &src, float *dst, bool normalise) {  // preview.cpp:28-80
        if (!dst) return;
        for (size_t i = 0; i < npix; ++i) {
            V3 o(0);
            if (wgt[i] > 0) {
                o = V3(src[4 * i], src[4 * i + 1], src[4 * i + 2]) / wgt[i];
                if (normalise) o = length(o) > 1e-3f ? normalize(o) : V3(0);
            }
            dst[4 * i] = o.x; dst[4 * i + 1] = o.y; dst[4 * i + 2] = o.z; dst[4 * i + 3] = 1;
        }
    };
    resolve(rad, out_color, false);
    if (s->want_albedo) resolve(alb, out_albedo, false);
    if (s->want_normal) resolve(nor, out_normal, true);
    if (raw_radiance) std::memcpy(raw_radiance, rad.data(), sizeof(float) * 4 * npix);
    if (raw_weight) std::memcpy(raw_weight, wgt.data(), sizeof(float) * npix);
    if (stats_out) {
        stats_out[0] = total.n_gen; stats_out[1] = total.n_closest; stats_out[2] = total.n_hit; stats_out[3] = total.n_next;
        stats_out[4] = total.n_shadow; stats_out[5] = total.n_shadow_miss; stats_out[6] = total.n_splat;
    }
}

// Renders only the listed pixels (index = y * width + x): out_rad[4 * k .. + 2] = the film's radiance sum of pixel k over sample
// indices [spp_begin, spp_end), out_rad[4 * k + 3] = its weight.  Every path of the render is a function of (pixel, sample index)
// alone and splats into its own pixel only, so this equals those pixels of orc_render's raw film: it lets tests
// compare a full-size GPU render with the oracle on a strided pixel subset in seconds.
void orc_render_pixels(orc_scene *s, int n, const uint32_t *pixel_index, int spp_begin, int spp_end, float *out_rad, uint64_t *stats_out) {
    const int W = s->width, H = s->height;
    const size_t npix = (size_t)W * H;
    Stats total;
#pragma omp parallel
    {
        Stats st;
#pragma omp for schedule(dynamic, 16)
        for (int k = 0; k < n; ++k) {
            const long long pix = pixel_index[k];
            float *rad = out_rad + 4 * (size_t)k;
            rad[0] = rad[1] = rad[2] = rad[3] = 0.0f;
            for (int sidx = spp_begin; sidx < spp_end; ++sidx) {
                int64_t accu = s->tile ? (int64_t)pix * s->spp + sidx : (int64_t)sidx * (int64_t)npix + pix;
                PathState ps;
                generate_path(*s, accu, ps);
                rad[3] += 1.0f; st.n_gen++;
                auto do_shadow = [&](const ShadowRay &sr) {
                    st.n_shadow++;
                    bool occluded = traverse<true>(*s, sr.o, sr.d, sr.t1).hit;
                    if (!occluded) { st.n_shadow_miss++; st.n_splat++; film_add(rad, sr.beta_li * shadow_tr(*s, sr)); }
                };
                for (;;) {
                    Hit h = traverse<false>(*s, ps.o, ps.d, ps.t1);
                    st.n_closest++; if (h.hit) st.n_hit++;
                    bool scattered = false;
                    MediumOut mo;
                    if (s->has_medium) {
                        medium_stage(*s, ps, h, mo);
                        scattered = mo.scattered;
                        if (mo.shadow.valid) do_shadow(mo.shadow);
                        if (mo.radiance.valid) { st.n_splat++; film_add(rad, mo.radiance.value); }
                    }
                    ShadeOut out;
                    shade(*s, ps, h, scattered, out);
                    if (out.shadow.valid) do_shadow(out.shadow);
                    if (out.radiance.valid) { st.n_splat++; film_add(rad, out.radiance.value); }
                    if (scattered) {
                        if (!mo.has_next) break;
                        st.n_next++;
                        ps = mo.next;
                        continue;
                    }
                    if (!out.has_next) break;
                    st.n_next++;
                    ps = out.next;
                }
            }
        }
#pragma omp critical
        {
            total.n_gen += st.n_gen; total.n_closest += st.n_closest; total.n_hit += st.n_hit; total.n_next += st.n_next;
            total.n_shadow += st.n_shadow; total.n_shadow_miss += st.n_shadow_miss; total.n_splat += st.n_splat;
        }
    }
    if (stats_out) {
        stats_out[0] = total.n_gen; stats_out[1] = total.n_closest; stats_out[2] = total.n_hit; stats_out[3] = total.n_next;
        stats_out[4] = total.n_shadow; stats_out[5] = total.n_shadow_miss; stats_out[6] = total.n_splat;
    }
}

// --- known-answer helpers for RNG / hash / alias / offset tests -------------------------------
uint64_t orc_murmur_pixel(uint32_t px, uint32_t py) { return hash_pixel(px, py); }
void orc_pcg_stream(uint32_t px, uint32_t py, int sample_index, int n, uint32_t *out_u32, float *out_f32, uint64_t *state_inc) {
    Sampler sm; sm.init(px, py, sample_index);
    if (state_inc) { state_inc[0] = sm.rng.state; state_inc[1] = sm.rng.inc; }
    for (int i = 0; i < n; ++i) {
        PCG copy = sm.rng;
        uint32_t r = copy.next();
        out_u32[i] = r;
        out_f32[i] = sm.get1d();
    }
}
void orc_set_reduce_order(int mode) { g_reduce_order = mode; }
void orc_alias_build(int n, const float *probs, float *accept, uint32_t *alias) {
    AliasTable t; t.initialize(std::vector<float>(probs, probs + n));
    for (int i = 0; i < n; ++i) { accept[i] = t.units[i].accept_prob; alias[i] = t.units[i].another_idx; }
}
void orc_intersection_offset(int n, const float *pos, const float *nor, const float *dir, float *out) {
    for (int i = 0; i < n; ++i) {
        V3 r = intersection_offset(V3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), V3(nor[3 * i], nor[3 * i + 1], nor[3 * i + 2]),
                                   V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
// all host cores regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its ranks: the CPU arm of bench.py must not inherit it)
void orc_use_all_cores() {
#ifdef _OPENMP
    omp_set_num_threads(omp_get_num_procs());
#endif
}

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
