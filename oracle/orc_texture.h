// ORACLE (test infrastructure only) -- CPU restatement of Btrc's 2-D / 3-D textures.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
//
// Reference: builtin/texture2d/{array2d,constant,math,transform}.cpp, builtin/texture3d/{array3d,binary}.cpp,
// common/btrc/utils/cuda/{texture,array}.cpp.  The reference samples CUDA texture objects (normalised coordinates,
// per-axis address mode, point / linear filter, optional sRGB->linear, 8-bit images as UNorm8): that arithmetic is the
// GPU's texture unit, not reference source.  This file restates the filtering rule the CUDA programming guide publishes
// ("Texture Fetching": x = N*u, xB = x - 0.5, i = floor(xB), alpha = frac(xB) kept in 9-bit fixed point with 8 fractional
// bits; sRGB decoded per texel before filtering).  Parity unpinned at that boundary: tests compare against the hardware
// with a tolerance of one 8-bit filter-weight step times the local texel contrast.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_math.h"

namespace orc {

enum AddressMode { ADDR_CLAMP = 0, ADDR_WRAP = 1, ADDR_MIRROR = 2, ADDR_BORDER = 3 };

struct TexImage {
    int w = 0, h = 0, d = 0;      // d == 0: 2-D
    int channels = 4;             // 1 or 4
    std::vector<float> data;      // w*h*max(d,1)*channels, decoded (UNorm8 / 255, sRGB->linear applied when requested)
    int addr[3] = {ADDR_CLAMP, ADDR_CLAMP, ADDR_CLAMP};
    bool linear = true;
    float border[4] = {0, 0, 0, 0};
    int border_levels = 0;        // 255: UNorm8 storage, 127: SNorm8, 0: float texels (the border colour is used as given)
    float max_value[3] = {0, 0, 0}, min_value[3] = {0, 0, 0};  // Array::find_minmax_values
};

inline float srgb_to_linear(float c) {  // IEC 61966-2-1, what cudaTextureDesc::sRGB applies per texel
    return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f);
}

// maps texel index i (possibly out of range) to a valid index or -1 (border)
inline int address_index(int i, int n, int mode) {
    switch (mode) {
    case ADDR_CLAMP: return i < 0 ? 0 : (i >= n ? n - 1 : i);
    case ADDR_WRAP: { int m = i % n; return m < 0 ? m + n : m; }
    case ADDR_MIRROR: {
        int p = 2 * n;
        int m = i % p; if (m < 0) m += p;
        return m < n ? m : p - 1 - m;
    }
    default: return (i < 0 || i >= n) ? -1 : i;
    }
}

// normalised coordinate -> unnormalised texel-space coordinate according to the address mode
inline float address_coord(float u, int n, int mode) {
    switch (mode) {
    case ADDR_WRAP: return (u - std::floor(u)) * n;
    case ADDR_MIRROR: {
        float f = std::floor(u);
        float fr = u - f;
        bool odd = (((long long)f) & 1) != 0;
        return (odd ? 1.0f - fr : fr) * n;
    }
    default: return u * n;  // clamp / border are applied on the texel index
    }
}

// Border colour as the texture unit returns it -- measured on B200 (scripts/diag_border.py, round 2): it is stored in the TEXEL
// FORMAT, so for 8-bit textures it is quantised (0.1 -> 25/255, 0.25 -> 64/255, 0.3 -> 76/255, 0.5 -> 127/255, 0.7 -> 178/255: the fp32
// product b * 255 rounded to nearest with ties DOWN, clamped to the format's range), and a one-channel texture has no second and
// third component to put it in: (b, 0, 0, 1), like any of its texels.
inline float border_component(const TexImage &t, int c) {
    const float b = t.border[c];
    if (t.border_levels == 0) return b;
    const float lv = (float)t.border_levels;
    float q = std::ceil(b * lv - 0.5f);
    const float lo = t.border_levels == 127 ? -127.0f : 0.0f;
    q = q < lo ? lo : (q > lv ? lv : q);
    return q / lv;
}

inline void fetch_texel(const TexImage &t, int x, int y, int z, float out[4]) {
    if (x < 0 || y < 0 || z < 0) {
        if (t.channels == 1) { out[0] = border_component(t, 0); out[1] = 0; out[2] = 0; out[3] = 1; }
        else for (int c = 0; c < 4; ++c) out[c] = border_component(t, c);
        return;
    }
    const size_t idx = ((size_t)z * t.h + y) * t.w + x;
    if (t.channels == 1) { out[0] = t.data[idx]; out[1] = 0; out[2] = 0; out[3] = 1; }
    else for (int c = 0; c < 4; ++c) out[c] = t.data[4 * idx + c];
}

inline float quantise_weight(float a) { return std::floor(a * 256.0f + 0.5f) / 256.0f; }

inline void sample_image(const TexImage &t, float u, float v, float w, float out[4]) {
    const bool is3d = t.d > 0;
    const float x = address_coord(u, t.w, t.addr[0]);
    const float y = address_coord(v, t.h, t.addr[1]);
    const float z = is3d ? address_coord(w, t.d, t.addr[2]) : 0.0f;
    if (!t.linear) {
        int ix = address_index((int)std::floor(x), t.w, t.addr[0]);
        int iy = address_index((int)std::floor(y), t.h, t.addr[1]);
        int iz = is3d ? address_index((int)std::floor(z), t.d, t.addr[2]) : 0;
        fetch_texel(t, ix, iy, iz, out);
        return;
    }
    const float xb = x - 0.5f, yb = y - 0.5f, zb = z - 0.5f;
    const int i = (int)std::floor(xb), j = (int)std::floor(yb), k = is3d ? (int)std::floor(zb) : 0;
    const float a = quantise_weight(xb - std::floor(xb)), b = quantise_weight(yb - std::floor(yb));
    const float g = is3d ? quantise_weight(zb - std::floor(zb)) : 0.0f;
    for (int c = 0; c < 4; ++c) out[c] = 0;
    for (int dz = 0; dz <= (is3d ? 1 : 0); ++dz)
        for (int dy = 0; dy <= 1; ++dy)
            for (int dx = 0; dx <= 1; ++dx) {
                float wgt = (dx ? a : 1 - a) * (dy ? b : 1 - b) * (is3d ? (dz ? g : 1 - g) : 1.0f);
                if (wgt == 0.0f) continue;
                float tx[4];
                fetch_texel(t, address_index(i + dx, t.w, t.addr[0]), address_index(j + dy, t.h, t.addr[1]),
                            is3d ? address_index(k + dz, t.d, t.addr[2]) : 0, tx);
                for (int c = 0; c < 4; ++c) out[c] += wgt * tx[c];
            }
}

// ---- 2-D texture graph -------------------------------------------------------------------------------------------
enum Tex2DType { TEX_CONSTANT = 0, TEX_ARRAY = 1, TEX_ADD = 2, TEX_SUB = 3, TEX_MUL = 4, TEX_DIV = 5, TEX_TRANSFORM = 6 };

struct Texture2D {
    int type = TEX_CONSTANT;
    V3 value{0, 0, 0};
    int image = -1;
    int lhs = -1, rhs = -1;  // TEX_TRANSFORM: lhs = wrapped texture
    float m[6] = {1, 0, 0, 0, 1, 0};  // Transform2D rows 0,1 of the 3x3 matrix (CTransform2D::apply_to_point)
};

struct TextureSet {
    std::vector<Texture2D> tex2d;
    std::vector<TexImage> images;

    V3 eval2d(int id, V2 uv) const {
        const Texture2D &t = tex2d[id];
        switch (t.type) {
        case TEX_CONSTANT: return t.value;  // core/texture2d.cpp:20-28
        case TEX_ARRAY: {  // array2d.cpp:18-30
            float o[4];
            sample_image(images[t.image], uv.x, uv.y, 0, o);
            return V3(o[0], o[1], o[2]);
        }
        case TEX_TRANSFORM: {  // transform.cpp:15-53
            V2 q{t.m[0] * uv.x + t.m[1] * uv.y + t.m[2], t.m[3] * uv.x + t.m[4] * uv.y + t.m[5]};
            return eval2d(t.lhs, q);
        }
        default: {  // math.cpp:20-96
            V3 a = eval2d(t.lhs, uv), b = eval2d(t.rhs, uv);
            return t.type == TEX_ADD ? a + b : t.type == TEX_SUB ? a - b : t.type == TEX_MUL ? a * b : a / b;
        }
        }
    }
};

// ---- 3-D textures (constant, array, add / mul of two; texture3d/array3d.cpp:29-41, texture3d/binary.cpp) -------
struct Texture3D {
    int type = TEX_CONSTANT;  // TEX_CONSTANT, TEX_ARRAY, TEX_ADD, TEX_MUL
    V3 value{0, 0, 0};
    int image = -1;
    int lhs = -1, rhs = -1;
};

}  // namespace orc
