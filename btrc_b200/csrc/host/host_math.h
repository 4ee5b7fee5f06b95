// Host-side fp32 math for scene preparation.  Operation order follows the reference's host math
// (src/common/btrc/utils/math/{vec3.h,mat4.cpp,transform.cpp,frame.cpp}) so that tables built here
// are bit-identical to the ones the oracle front-end builds independently; host code is compiled
// with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace btrc {

struct Vec2 { float x = 0, y = 0; };

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit Vec3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(Vec3 v) { return std::sqrt(dot(v, v)); }
inline Vec3 normalize(Vec3 v) { return 1.0f / length(v) * v; }  // vec3.h:110-114
// plain comparisons (minss / maxss), not std::fmin / std::fmax: without -ffast-math those are libm calls, and the BVH builder
// spends its time here (10 M triangles: 62 s -> see DESIGN.md); the inputs are never NaN
inline float fmin_(float a, float b) { return b < a ? b : a; }
inline float fmax_(float a, float b) { return a < b ? b : a; }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }

struct AABB {
    Vec3 lo{3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    Vec3 hi{-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    void extend(Vec3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    void extend(const AABB &b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    bool empty() const { return lo.x > hi.x; }
    float half_area() const {
        Vec3 d = hi - lo;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
};

struct Mat4 {
    float m[4][4];
    static Mat4 identity() {
        Mat4 r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0f : 0.0f;
        return r;
    }
    static Mat4 rows(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3,
                     float c0, float c1, float c2, float c3, float d0, float d1, float d2, float d3) {
        Mat4 r = {{{a0, a1, a2, a3}, {b0, b1, b2, b3}, {c0, c1, c2, c3}, {d0, d1, d2, d3}}};
        return r;
    }
};

inline Mat4 operator*(const Mat4 &a, const Mat4 &b) {  // mat4.cpp:186-199
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 4; ++c)
            r.m[i][c] = a.m[i][0] * b.m[0][c] + a.m[i][1] * b.m[1][c] + a.m[i][2] * b.m[2][c] + a.m[i][3] * b.m[3][c];
    return r;
}

struct Mat3 {
    float m[3][3];
    static Mat3 identity() {
        Mat3 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = i == j ? 1.0f : 0.0f;
        return r;
    }
};

inline Mat3 operator*(const Mat3 &a, const Mat3 &b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c)
            r.m[i][c] = a.m[i][0] * b.m[0][c] + a.m[i][1] * b.m[1][c] + a.m[i][2] * b.m[2][c];
    return r;
}

// reference: utils/math/transform.h:35-63 -- a matrix and its inverse, composed analytically
struct Transform3D {
    Mat4 mat = Mat4::identity();
    Mat4 inv = Mat4::identity();

    Vec3 point(Vec3 p) const {  // transform.cpp:66-72
        return {mat.m[0][0] * p.x + mat.m[0][1] * p.y + mat.m[0][2] * p.z + mat.m[0][3],
                mat.m[1][0] * p.x + mat.m[1][1] * p.y + mat.m[1][2] * p.z + mat.m[1][3],
                mat.m[2][0] * p.x + mat.m[2][1] * p.y + mat.m[2][2] * p.z + mat.m[2][3]};
    }
    static Transform3D translate(float x, float y, float z);
    static Transform3D rotate(Vec3 axis, float rad);
    static Transform3D rotate_x(float rad);
    static Transform3D rotate_y(float rad);
    static Transform3D rotate_z(float rad);
    static Transform3D scale(float x, float y, float z);
};

inline Transform3D operator*(const Transform3D &a, const Transform3D &b) {  // transform.cpp:127-130
    Transform3D r;
    r.mat = a.mat * b.mat;
    r.inv = b.inv * a.inv;
    return r;
}

}  // namespace btrc
