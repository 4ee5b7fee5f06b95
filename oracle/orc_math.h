// ORACLE (test infrastructure only) -- CPU restatement of Btrc's wavefront path tracer.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
//
// Small fp32 vector helpers whose operation ORDER follows the reference's CUJ math
// (src/common/btrc/utils/cmath/cvec3.h:153-180): dot = x*x' + y*y' + z*z' left to right,
// normalize = rsqrt(dot(v,v)) * v.  Compiled with -ffp-contract=off so no FMA is formed.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

constexpr float PI = 3.1415926535f;            // utils/math/scalar.h:8
constexpr float MAX_FLOAT = 3.402823466e+38f;  // utils/math/scalar.h:7

struct V2 { float x = 0, y = 0; };
struct V3 {
    float x = 0, y = 0, z = 0;
    V3() = default;
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length_square(V3 v) { return dot(v, v); }
inline float length(V3 v) { return std::sqrt(dot(v, v)); }
// device-side normalize (cvec3.h:163-167): rsqrt * v
inline V3 normalize(V3 v) { float inv = 1.0f / std::sqrt(dot(v, v)); return inv * v; }
// host-side normalize (utils/math/vec3.h:110-114): T(1) / length(v) * v
inline V3 host_normalize(V3 v) { return 1.0f / length(v) * v; }
inline bool is_zero(V3 v) { return v.x == 0 && v.y == 0 && v.z == 0; }
inline bool isfinite3(V3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }
inline float lum(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }  // core/spectrum.cpp:84-87
inline float saturate(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline float lerp(float a, float b, float t) { return a * (1.0f - t) + b * t; }

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int32_t f2i(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
inline float i2f(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Row-major 4x4; Transform3D keeps mat and inv (utils/math/transform.h:35-63).
struct M4 { float m[4][4]; };
struct Xf {
    M4 mat, inv;
    // ctransform.cpp:36-58
    V3 point(V3 p) const {
        return {mat.m[0][0] * p.x + mat.m[0][1] * p.y + mat.m[0][2] * p.z + mat.m[0][3],
                mat.m[1][0] * p.x + mat.m[1][1] * p.y + mat.m[1][2] * p.z + mat.m[1][3],
                mat.m[2][0] * p.x + mat.m[2][1] * p.y + mat.m[2][2] * p.z + mat.m[2][3]};
    }
    V3 vector(V3 v) const {
        return {mat.m[0][0] * v.x + mat.m[0][1] * v.y + mat.m[0][2] * v.z,
                mat.m[1][0] * v.x + mat.m[1][1] * v.y + mat.m[1][2] * v.z,
                mat.m[2][0] * v.x + mat.m[2][1] * v.y + mat.m[2][2] * v.z};
    }
    V3 normal(V3 n) const {
        return {inv.m[0][0] * n.x + inv.m[1][0] * n.y + inv.m[2][0] * n.z,
                inv.m[0][1] * n.x + inv.m[1][1] * n.y + inv.m[2][1] * n.z,
                inv.m[0][2] * n.x + inv.m[1][2] * n.y + inv.m[2][2] * n.z};
    }
    // inverse transform of a point (used by TransformMedium, core/medium.cpp:98-129)
    V3 inv_point(V3 p) const {
        return {inv.m[0][0] * p.x + inv.m[0][1] * p.y + inv.m[0][2] * p.z + inv.m[0][3],
                inv.m[1][0] * p.x + inv.m[1][1] * p.y + inv.m[1][2] * p.z + inv.m[1][3],
                inv.m[2][0] * p.x + inv.m[2][1] * p.y + inv.m[2][2] * p.z + inv.m[2][3]};
    }
};

// CFrame (cframe.cpp:17-23): constructor normalises all three axes.
struct Frame {
    V3 x{1, 0, 0}, y{0, 1, 0}, z{0, 0, 1};
    Frame() = default;
    Frame(V3 _x, V3 _y, V3 _z) : x(normalize(_x)), y(normalize(_y)), z(normalize(_z)) {}
    V3 local_to_global(V3 l) const { return x * l.x + y * l.y + z * l.z; }
    V3 global_to_local(V3 g) const { return {dot(g, x), dot(g, y), dot(g, z)}; }
    Frame rotate_to_new_z(V3 new_z) const {  // cframe.cpp:74-79
        V3 nx = cross(y, new_z);
        V3 ny = cross(new_z, nx);
        return Frame(nx, ny, new_z);
    }
    bool in_positive_z_hemisphere(V3 v) const { return dot(v, z) > 0; }
    static Frame from_z(V3 _z) {  // cframe.cpp:52-64
        V3 nz = normalize(_z), ny;
        if (1.0f - std::fabs(nz.x) < 0.1f) ny = cross(nz, V3(0, 1, 0));
        else ny = cross(nz, V3(1, 0, 0));
        return Frame(cross(nz, ny), ny, nz);
    }
};

}  // namespace orc
