// Device-side fp32 helpers.  Operation order mirrors the reference's CUJ math
// (src/common/btrc/utils/cmath/cvec3.h:153-180, cframe.cpp, ctransform.cpp) so that the kernels and the CPU
// oracle walk the same paths; the compiler is free to contract a*b+c into FMA here (the reference build
// does too: fast_math in lib/cuj/cuj/src/gen/ptx.cpp:57-63).  The triangle test in traverse.cuh is the
// exception: it uses explicit round-to-nearest intrinsics and is bit-exact against the oracle.
#pragma once
#include <cuda_runtime.h>

#include "../common/types.h"

namespace btrc {

struct f3 {
    float x, y, z;
    __device__ __forceinline__ f3() {}
    __device__ __forceinline__ f3(float a, float b, float c) : x(a), y(b), z(c) {}
    __device__ __forceinline__ explicit f3(float a) : x(a), y(a), z(a) {}
};
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator/(f3 a, f3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ f3 operator-(f3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length_square(f3 v) { return dot(v, v); }
__device__ __forceinline__ float length(f3 v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ f3 normalize(f3 v) { return rsqrtf(dot(v, v)) * v; }  // cvec3.h:163-167
__device__ __forceinline__ bool is_zero(f3 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f; }
__device__ __forceinline__ bool isfinite3(f3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }
__device__ __forceinline__ float lum(f3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }
__device__ __forceinline__ float saturate(float x) { return __saturatef(x); }
__device__ __forceinline__ float lerpf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ f3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 make4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

// read-only 16-byte loads (LDG.E.128.CONSTANT)
__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// 3x4 row-major affine transform helpers (ctransform.cpp:36-58)
__device__ __forceinline__ f3 xf_point(const float *m, f3 p) {
    return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
__device__ __forceinline__ f3 xf_vector(const float *m, f3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// normal transform: transpose of the inverse
__device__ __forceinline__ f3 xf_normal(const float *inv, f3 n) {
    return f3(inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
              inv[2] * n.x + inv[6] * n.y + inv[10] * n.z);
}

// CFrame (cframe.cpp:17-23, 52-86): the constructor normalises all three axes
struct Frame3 {
    f3 x, y, z;
    __device__ __forceinline__ Frame3() {}
    __device__ __forceinline__ Frame3(f3 _x, f3 _y, f3 _z) : x(normalize(_x)), y(normalize(_y)), z(normalize(_z)) {}
    __device__ __forceinline__ f3 local_to_global(f3 l) const { return x * l.x + y * l.y + z * l.z; }
    __device__ __forceinline__ f3 global_to_local(f3 g) const { return f3(dot(g, x), dot(g, y), dot(g, z)); }
    __device__ __forceinline__ Frame3 rotate_to_new_z(f3 new_z) const {
        f3 nx = cross(y, new_z);
        f3 ny = cross(new_z, nx);
        return Frame3(nx, ny, new_z);
    }
    __device__ __forceinline__ bool in_positive_z_hemisphere(f3 v) const { return dot(v, z) > 0.0f; }
};

// utils/intersection.cpp:5-36 (Waechter-Binder integer-ulp ray offset, doubled int scale)
__device__ __forceinline__ f3 intersection_offset(f3 pos, f3 nor) {
    constexpr float ORIGIN = 1.0f / 32.0f;
    constexpr float FLOAT_SCALE = 1.0f / 65536.0f;
    constexpr float INT_SCALE = 256.0f;
    int of_i_x = int(2 * INT_SCALE * nor.x);
    int of_i_y = int(2 * INT_SCALE * nor.y);
    int of_i_z = int(2 * INT_SCALE * nor.z);
    float p_i_x = __int_as_float(__float_as_int(pos.x) + (pos.x < 0 ? -of_i_x : of_i_x));
    float p_i_y = __int_as_float(__float_as_int(pos.y) + (pos.y < 0 ? -of_i_y : of_i_y));
    float p_i_z = __int_as_float(__float_as_int(pos.z) + (pos.z < 0 ? -of_i_z : of_i_z));
    float r_x = fabsf(pos.x) < ORIGIN ? __fadd_rn(pos.x, __fmul_rn(FLOAT_SCALE, nor.x)) : p_i_x;
    float r_y = fabsf(pos.y) < ORIGIN ? __fadd_rn(pos.y, __fmul_rn(FLOAT_SCALE, nor.y)) : p_i_y;
    float r_z = fabsf(pos.z) < ORIGIN ? __fadd_rn(pos.z, __fmul_rn(FLOAT_SCALE, nor.z)) : p_i_z;
    return f3(r_x, r_y, r_z);
}
__device__ __forceinline__ f3 intersection_offset(f3 pos, f3 nor, f3 next_dir) {
    return intersection_offset(pos, dot(nor, next_dir) > 0.0f ? nor : -nor);
}

// utils/cmath/cdistribution.cpp:31-75
__device__ __forceinline__ f3 sample_hemisphere_zweighted(float u1, float u2) {
    float sx = 0.0f, sy = 0.0f;
    u1 = 2.0f * u1 - 1.0f;
    u2 = 2.0f * u2 - 1.0f;
    if (u1 != 0.0f || u2 != 0.0f) {
        float theta, r;
        if (fabsf(u1) > fabsf(u2)) {
            r = u1;
            theta = 0.25f * kPi * (u2 / u1);
        } else {
            r = u2;
            theta = 0.5f * kPi - 0.25f * kPi * (u1 / u2);
        }
        float s, c;
        __sincosf(theta, &s, &c);  // |theta| <= pi: the SFU form, like the reference's sin/cos.approx
        sx = r * c;
        sy = r * s;
    }
    float z = sqrtf(fmaxf(0.0f, 1.0f - (sx * sx + sy * sy)));
    return f3(sx, sy, z);
}
__device__ __forceinline__ float pdf_sample_hemisphere_zweighted(f3 v) { return v.z >= 0.0f ? v.z / kPi : 0.0f; }
__device__ __forceinline__ float2 sample_triangle_uniform(float u1, float u2) {
    float t = sqrtf(u1);
    return make_float2(1.0f - t, t * u2);
}

// utils/cmath/calias.cpp:10-23
__device__ __forceinline__ uint32_t alias_sample(const float2 *table, uint32_t n, float u) {
    float nu = float(n) * u;
    uint32_t i = min(uint32_t(nu), n - 1);
    float s = nu - float(i);
    float2 unit = __ldg(table + i);
    return s <= unit.x ? i : __float_as_uint(unit.y);
}

}  // namespace btrc
