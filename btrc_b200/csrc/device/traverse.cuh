// Ray / BVH traversal for sm_100a: closest hit and occlusion (early exit).
//
// Replaces the reference's OptiX pipelines (wavefront/trace.cpp:38-78, shadow.cpp:24-36; accel build in
// utils/optix/context.cpp:158-215).  Semantics kept: double-sided triangles, no any-hit program, t in [0, ray.t],
// closest hit returns (instance id, primitive id, barycentrics (u -> v1, v -> v2), t).
//
// Triangle test: Moeller-Trumbore in the exact operation order of lib/bvh/include/bvh/triangle.hpp:95-116, with
// round-to-nearest intrinsics (no FMA contraction), so t/u/v are bit-identical to the CPU oracle.
// Tie rule: smallest t, then smallest (instance, primitive) -- independent of traversal order.
#pragma once
#include "dev_math.cuh"

namespace btrc {

struct HitRec {
    float t, u, v;
    int slot;  // triangle slot in BvhDevice::tris, -1 = miss
};

__device__ __forceinline__ float dot_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// returns true when the triangle is hit with t in [0, tmax]
__device__ __forceinline__ bool tri_test(const float4 *__restrict__ tri, f3 o, f3 d, float tmax, float &t, float &u, float &v) {
    const float4 a = ldg4(tri), b = ldg4(tri + 1), c4 = ldg4(tri + 2);  // (p0|n.x) (e1|n.y) (e2|n.z)
    const float cx = __fsub_rn(a.x, o.x), cy = __fsub_rn(a.y, o.y), cz = __fsub_rn(a.z, o.z);
    // r = cross(d, c)
    const float rx = __fsub_rn(__fmul_rn(d.y, cz), __fmul_rn(d.z, cy));
    const float ry = __fsub_rn(__fmul_rn(d.z, cx), __fmul_rn(d.x, cz));
    const float rz = __fsub_rn(__fmul_rn(d.x, cy), __fmul_rn(d.y, cx));
    const float det = dot_rn(a.w, b.w, c4.w, d.x, d.y, d.z);
    const float inv_det = __fdiv_rn(1.0f, det);
    u = __fmul_rn(dot_rn(rx, ry, rz, c4.x, c4.y, c4.z), inv_det);
    v = __fmul_rn(dot_rn(rx, ry, rz, b.x, b.y, b.z), inv_det);
    const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
    if (u >= 0.0f && v >= 0.0f && w >= 0.0f) {
        t = __fmul_rn(dot_rn(a.w, b.w, c4.w, cx, cy, cz), inv_det);
        return t >= 0.0f && t <= tmax;
    }
    return false;
}

__device__ __forceinline__ bool id_less(const uint2 *__restrict__ ids, int a, int b) {
    const uint2 ia = __ldg(ids + a), ib = __ldg(ids + b);
    return ia.x != ib.x ? ia.x < ib.x : ia.y < ib.y;
}

constexpr int kStackSize = 64;

// CONVERGE: all 32 lanes of the warp call this together (lanes without a ray pass active = false); the vote at the top of
// every descend-to-leaf / process-leaf round reconverges the warp (classic while-while), otherwise lanes drift apart and
// the node loop and the leaf loop are issued separately for a handful of lanes each.
template <bool ANY_HIT, bool CONVERGE = false>
__device__ __forceinline__ HitRec traverse_bvh2(const BvhDevice &bvh, f3 o, f3 d, float tmax, bool active = true) {
    HitRec best;
    best.t = tmax; best.u = 0.0f; best.v = 0.0f; best.slot = -1;

    // A zero direction component would give inf * 0 = NaN in the slab test; clamp it to a tiny finite slope instead
    // (the slab then evaluates to +-huge with the sign of (plane - origin), which is the exact answer for a parallel ray).
    constexpr float kTiny = 1e-20f;
    const float idx = 1.0f / (fabsf(d.x) > kTiny ? d.x : copysignf(kTiny, d.x));
    const float idy = 1.0f / (fabsf(d.y) > kTiny ? d.y : copysignf(kTiny, d.y));
    const float idz = 1.0f / (fabsf(d.z) > kTiny ? d.z : copysignf(kTiny, d.z));
    const float oox = o.x * idx, ooy = o.y * idy, ooz = o.z * idz;

    int stack[kStackSize];
    int sp = 0;
    constexpr int SENTINEL = 0x7fffffff;
    int cur = active ? bvh.root_ref : SENTINEL;
    stack[sp++] = SENTINEL;

    while (CONVERGE ? __any_sync(0xffffffffu, cur != SENTINEL) : (cur != SENTINEL)) {
        // descend through inner nodes
        while (cur >= 0 && cur != SENTINEL) {
            const float4 *node = bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = ldg4(node), n1 = ldg4(node + 1), nz = ldg4(node + 2), meta = ldg4(node + 3);
            // slabs; NaN (0 * inf) lanes drop out of fminf/fmaxf, which is the conservative outcome
            const float c0lox = fmaf(n0.x, idx, -oox), c0hix = fmaf(n0.y, idx, -oox);
            const float c0loy = fmaf(n0.z, idy, -ooy), c0hiy = fmaf(n0.w, idy, -ooy);
            const float c0loz = fmaf(nz.x, idz, -ooz), c0hiz = fmaf(nz.y, idz, -ooz);
            const float c1lox = fmaf(n1.x, idx, -oox), c1hix = fmaf(n1.y, idx, -oox);
            const float c1loy = fmaf(n1.z, idy, -ooy), c1hiy = fmaf(n1.w, idy, -ooy);
            const float c1loz = fmaf(nz.z, idz, -ooz), c1hiz = fmaf(nz.w, idz, -ooz);
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), best.t));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), best.t));
            const bool h0 = c0min <= c0max, h1 = c1min <= c1max;
            const int r0 = __float_as_int(meta.x), r1 = __float_as_int(meta.y);
            if (!h0 && !h1) {
                cur = stack[--sp];
            } else if (h0 && h1) {
                const bool swp = c1min < c0min;
                cur = swp ? r1 : r0;
                stack[sp++] = swp ? r0 : r1;
            } else {
                cur = h0 ? r0 : r1;
            }
        }
        // leaf
        if (cur < 0) {
            const int ref = ~cur;
            const int first = ref >> 3, count = (ref & 7) + 1;
            for (int s = first; s < first + count; ++s) {
                float t, u, v;
                if (tri_test(bvh.tris + 3 * (size_t)s, o, d, best.t, t, u, v)) {
                    if (ANY_HIT) {
                        best.slot = s;
                        if (!CONVERGE) return best;
                        sp = 1;  // the next pop yields the sentinel
                        break;
                    }
                    if (t < best.t || best.slot < 0 || id_less(bvh.tri_ids, s, best.slot)) {
                        best.t = t; best.u = u; best.v = v; best.slot = s;
                    }
                }
            }
            cur = stack[--sp];
        }
    }
    return best;
}


}  // namespace btrc
