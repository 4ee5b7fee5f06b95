// Ray / BVH traversal for sm_100a: closest hit and occlusion (early exit).
//
// Replaces the reference's OptiX pipelines (wavefront/trace.cpp:38-78, shadow.cpp:24-36; accel build in
// utils/optix/context.cpp:158-215).  Semantics kept: double-sided triangles, no any-hit program, t in [0, ray.t],
// closest hit returns (instance id, primitive id, barycentrics (u -> v1, v -> v2), t).
//
// Triangle test: Moeller-Trumbore in the exact operation order of lib/bvh/include/bvh/triangle.hpp:95-116, with
// round-to-nearest intrinsics (no FMA contraction), so t/u/v are bit-identical to the CPU oracle.
// Tie rule: smallest t, then smallest (instance, primitive) -- independent of traversal order.
//
// The file only depends on <cuda_runtime.h> vector types, common/types.h and a handful of intrinsics, so that tests/ can
// compile the very same function for the host (tests/host/host_traverse.cpp supplies the intrinsics) and check it
// against the oracle without a GPU.  BTRC_TRAV_QUAL is the only switch; the product build never defines it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../common/types.h"

#ifndef BTRC_TRAV_QUAL
#define BTRC_TRAV_QUAL __device__ __forceinline__
#endif
#ifndef BTRC_TRAV_COUNT
#define BTRC_TRAV_COUNT(which)  // host test harness only: traversal-step / triangle-test statistics
#endif

namespace btrc {

struct HitRec {
    float t, u, v;
    int slot;  // triangle slot in BvhDevice::tris, -1 = miss
};

BTRC_TRAV_QUAL float dot_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// returns true when the triangle is hit with t in [0, tmax]
BTRC_TRAV_QUAL bool tri_test(const float4 *__restrict__ tri, float ox, float oy, float oz, float dx, float dy, float dz, float tmax,
                             float &t, float &u, float &v) {
    const float4 a = __ldg(tri), b = __ldg(tri + 1), c4 = __ldg(tri + 2);  // (p0|n.x) (e1|n.y) (e2|n.z)
    const float cx = __fsub_rn(a.x, ox), cy = __fsub_rn(a.y, oy), cz = __fsub_rn(a.z, oz);
    // r = cross(d, c)
    const float rx = __fsub_rn(__fmul_rn(dy, cz), __fmul_rn(dz, cy));
    const float ry = __fsub_rn(__fmul_rn(dz, cx), __fmul_rn(dx, cz));
    const float rz = __fsub_rn(__fmul_rn(dx, cy), __fmul_rn(dy, cx));
    const float det = dot_rn(a.w, b.w, c4.w, dx, dy, dz);
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

BTRC_TRAV_QUAL bool id_less(const uint2 *__restrict__ ids, int a, int b) {
    const uint2 ia = __ldg(ids + a), ib = __ldg(ids + b);
    return ia.x != ib.x ? ia.x < ib.x : ia.y < ib.y;
}

constexpr int kStackSize = 64;

// CONVERGE: all 32 lanes of the warp call this together (lanes without a ray pass active = false); the vote at the top of
// every descend-to-leaf / process-leaf round reconverges the warp (classic while-while), otherwise lanes drift apart and
// the node loop and the leaf loop are issued separately for a handful of lanes each.
template <bool ANY_HIT, bool CONVERGE = false>
BTRC_TRAV_QUAL HitRec traverse_bvh2(const BvhDevice &bvh, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, bool active = true) {
    HitRec best;
    best.t = tmax; best.u = 0.0f; best.v = 0.0f; best.slot = -1;

    // A zero direction component would give inf * 0 = NaN in the slab test; clamp it to a tiny finite slope instead
    // (the slab then evaluates to +-huge with the sign of (plane - origin), which is the exact answer for a parallel ray).
    constexpr float kTiny = 1e-20f;
    const float idx = 1.0f / (fabsf(dx) > kTiny ? dx : copysignf(kTiny, dx));
    const float idy = 1.0f / (fabsf(dy) > kTiny ? dy : copysignf(kTiny, dy));
    const float idz = 1.0f / (fabsf(dz) > kTiny ? dz : copysignf(kTiny, dz));
    const float oox = ox * idx, ooy = oy * idy, ooz = oz * idz;

    int stack[kStackSize];
    int sp = 0;
    constexpr int SENTINEL = 0x7fffffff;
    int cur = active ? bvh.root_ref : SENTINEL;
    stack[sp++] = SENTINEL;

    while (CONVERGE ? __any_sync(0xffffffffu, cur != SENTINEL) : (cur != SENTINEL)) {
        // descend through inner nodes
        while (cur >= 0 && cur != SENTINEL) {
            const float4 *node = bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = __ldg(node), n1 = __ldg(node + 1), nz = __ldg(node + 2), meta = __ldg(node + 3);
            BTRC_TRAV_COUNT(0);
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
                BTRC_TRAV_COUNT(1);
                if (tri_test(bvh.tris + 3 * (size_t)s, ox, oy, oz, dx, dy, dz, best.t, t, u, v)) {
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


// "if-if" form of the same walk: every iteration a lane does ONE thing -- a triangle of its current leaf, or one node step --
// and all lanes vote once per iteration.  Lanes that are in a leaf no longer wait for the slowest lane to finish its descent
// (and vice versa); a warp-level cost model on real C2 ray batches (experiments/simt_sim) gives 11-19 % fewer issue slots for
// bounce and shadow rays than the while-while form.  Results are identical (same boxes, same triangle test, same tie rule).
template <bool ANY_HIT, bool CONVERGE = false>
BTRC_TRAV_QUAL HitRec traverse_bvh2_ifif(const BvhDevice &bvh, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, bool active = true) {
    HitRec best;
    best.t = tmax; best.u = 0.0f; best.v = 0.0f; best.slot = -1;
    constexpr float kTiny = 1e-20f;
    const float idx = 1.0f / (fabsf(dx) > kTiny ? dx : copysignf(kTiny, dx));
    const float idy = 1.0f / (fabsf(dy) > kTiny ? dy : copysignf(kTiny, dy));
    const float idz = 1.0f / (fabsf(dz) > kTiny ? dz : copysignf(kTiny, dz));
    const float oox = ox * idx, ooy = oy * idy, ooz = oz * idz;

    int stack[kStackSize];
    int sp = 0;
    constexpr int SENTINEL = 0x7fffffff;
    stack[sp++] = SENTINEL;
    int cur = active ? bvh.root_ref : SENTINEL;
    int tri = 0, tri_end = 0;  // triangles [tri, tri_end) of the current leaf are still to be tested
    if (cur < 0) { const int ref = ~cur; tri = ref >> 3; tri_end = tri + (ref & 7) + 1; cur = stack[--sp]; }  // the root is a leaf

    while (CONVERGE ? __any_sync(0xffffffffu, cur != SENTINEL || tri < tri_end) : (cur != SENTINEL || tri < tri_end)) {
        if (tri < tri_end) {
            const int s = tri++;
            float t, u, v;
            BTRC_TRAV_COUNT(1);
            if (tri_test(bvh.tris + 3 * (size_t)s, ox, oy, oz, dx, dy, dz, best.t, t, u, v)) {
                if (ANY_HIT) {
                    best.slot = s;
                    if (!CONVERGE) return best;
                    tri = tri_end; cur = SENTINEL;  // finished; keeps voting until the warp is done
                } else if (t < best.t || best.slot < 0 || id_less(bvh.tri_ids, s, best.slot)) {
                    best.t = t; best.u = u; best.v = v; best.slot = s;
                }
            }
        } else if (cur != SENTINEL) {
            const float4 *node = bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = __ldg(node), n1 = __ldg(node + 1), nz = __ldg(node + 2), meta = __ldg(node + 3);
            BTRC_TRAV_COUNT(0);
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
        if (tri >= tri_end && cur < 0) {  // arrived at a leaf: take its triangle range and the next node from the stack
            const int ref = ~cur;
            tri = ref >> 3; tri_end = tri + (ref & 7) + 1;
            cur = stack[--sp];
        }
    }
    return best;
}

// Two-level walk (host/bvh.h TwoLevelBvh; replaces the reference's IAS -> GAS traversal, utils/optix/context.cpp:217-294): the
// instance tree in world space, then the mesh's tree in OBJECT space -- the ray is transformed for the box tests only
// (o' = W2O * o, d' = W2O * d without renormalising, so t is the same parameter in both spaces and best.t culls directly) --
// while every triangle is still tested in WORLD space on the instance's own world-space record.  t / u / v, the accepted hit
// and the tie rule are therefore the flat tree's, bit for bit; the mesh trees' boxes carry the extra padding that makes the
// fp32 transform conservative.  Same if-if iteration shape as traverse_bvh2_ifif: one triangle, one node step or one level
// switch per lane and iteration, one vote per iteration.
template <bool ANY_HIT, bool CONVERGE = false>
BTRC_TRAV_QUAL HitRec traverse_two_level(const BvhDevice &bvh, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, bool active = true) {
    HitRec best;
    best.t = tmax; best.u = 0.0f; best.v = 0.0f; best.slot = -1;
    constexpr float kTiny = 1e-20f;
    // slab constants of the space the walk is in (world at the top, the instance's object space below)
    float idx = 1.0f / (fabsf(dx) > kTiny ? dx : copysignf(kTiny, dx));
    float idy = 1.0f / (fabsf(dy) > kTiny ? dy : copysignf(kTiny, dy));
    float idz = 1.0f / (fabsf(dz) > kTiny ? dz : copysignf(kTiny, dz));
    float oox = ox * idx, ooy = oy * idy, ooz = oz * idz;

    int stack[kStackSize];
    int sp = 0;
    constexpr int SENTINEL = 0x7fffffff, LEAVE = 0x7ffffffe;  // LEAVE: everything above it on the stack belongs to a mesh tree
    stack[sp++] = SENTINEL;
    int cur = active ? bvh.root_ref : SENTINEL;
    bool below = false;        // walking a mesh tree
    int tri = 0, tri_end = 0;  // world triangle slots [tri, tri_end) of the current leaf are still to be tested
    int tri_base = 0;
    const float4 *nodes = bvh.tl_nodes;

    while (CONVERGE ? __any_sync(0xffffffffu, cur != SENTINEL || tri < tri_end) : (cur != SENTINEL || tri < tri_end)) {
        if (tri < tri_end) {
            const int s = tri++;
            float t, u, v;
            BTRC_TRAV_COUNT(1);
            if (tri_test(bvh.tris + 3 * (size_t)s, ox, oy, oz, dx, dy, dz, best.t, t, u, v)) {
                if (ANY_HIT) {
                    best.slot = s;
                    if (!CONVERGE) return best;
                    tri = tri_end; cur = SENTINEL;
                } else if (t < best.t || best.slot < 0 || id_less(bvh.tri_ids, s, best.slot)) {
                    best.t = t; best.u = u; best.v = v; best.slot = s;
                }
            }
        } else if (cur == LEAVE) {  // the mesh tree is exhausted: back to world space
            idx = 1.0f / (fabsf(dx) > kTiny ? dx : copysignf(kTiny, dx));
            idy = 1.0f / (fabsf(dy) > kTiny ? dy : copysignf(kTiny, dy));
            idz = 1.0f / (fabsf(dz) > kTiny ? dz : copysignf(kTiny, dz));
            oox = ox * idx; ooy = oy * idy; ooz = oz * idz;
            below = false;
            nodes = bvh.tl_nodes;
            cur = stack[--sp];
        } else if (cur >= 0 && cur != SENTINEL) {
            const float4 *node = nodes + 4 * (size_t)cur;
            const float4 n0 = __ldg(node), n1 = __ldg(node + 1), nz = __ldg(node + 2), meta = __ldg(node + 3);
            BTRC_TRAV_COUNT(0);
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
        if (tri >= tri_end && cur < 0) {  // arrived at a leaf
            const int ref = ~cur;
            if (below) {  // triangles of the mesh: slots of this instance's world-space records
                tri = tri_base + (ref >> 3); tri_end = tri + (ref & 7) + 1;
                cur = stack[--sp];
            } else {      // an instance: enter its mesh tree
                const float4 *rec = bvh.inst_blas + 4 * (size_t)(ref >> 3);
                const float4 m0 = __ldg(rec), m1 = __ldg(rec + 1), m2 = __ldg(rec + 2), meta = __ldg(rec + 3);
                const float lox = fmaf(m0.x, ox, fmaf(m0.y, oy, fmaf(m0.z, oz, m0.w)));
                const float loy = fmaf(m1.x, ox, fmaf(m1.y, oy, fmaf(m1.z, oz, m1.w)));
                const float loz = fmaf(m2.x, ox, fmaf(m2.y, oy, fmaf(m2.z, oz, m2.w)));
                const float ldx = fmaf(m0.x, dx, fmaf(m0.y, dy, m0.z * dz));
                const float ldy = fmaf(m1.x, dx, fmaf(m1.y, dy, m1.z * dz));
                const float ldz = fmaf(m2.x, dx, fmaf(m2.y, dy, m2.z * dz));
                idx = 1.0f / (fabsf(ldx) > kTiny ? ldx : copysignf(kTiny, ldx));
                idy = 1.0f / (fabsf(ldy) > kTiny ? ldy : copysignf(kTiny, ldy));
                idz = 1.0f / (fabsf(ldz) > kTiny ? ldz : copysignf(kTiny, ldz));
                oox = lox * idx; ooy = loy * idy; ooz = loz * idz;
                tri_base = __float_as_int(meta.y);
                stack[sp++] = LEAVE;
                below = true;
                nodes = bvh.nodes;
                cur = __float_as_int(meta.x);
            }
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------------------------
// Compressed 8-wide BVH (host/bvh.h WideBvh8; selected by "bvh_layout": "wide8").  One traversal step = one 80-byte node (five
// LDG.128) and eight quantised child boxes tested in registers; the children that were hit are kept as a bit mask ordered front to
// back by the ray octant, so the stack holds one 8-byte entry per level instead of one per child.  Node boxes are conservative
// supersets of the padded triangle boxes, the triangle test and the tie rule are the binary walk's: results are bit-identical.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStack8 = 32;  // one entry per level of wide nodes (the host refuses deeper trees)

// byte j of x as the float 2^23 + q: one PRMT builds 0x4B0000qq (no I2F); the 2^23 is folded into the plane offset
BTRC_TRAV_QUAL float u8f(unsigned x, unsigned sel) { return __uint_as_float(__byte_perm(x, 0x4B000000u, sel)); }

// CONVERGE: all 32 lanes of the warp call this together (lanes without a ray pass active = false) and vote before
// every node step, which keeps the warp on one instruction stream (node step, then its triangles).
template <bool ANY_HIT, bool CONVERGE>
BTRC_TRAV_QUAL HitRec traverse_bvh8(const BvhDevice &bvh, float ox, float oy, float oz, float dx, float dy, float dz, float tmax, bool active = true) {
    HitRec best;
    best.t = tmax; best.u = 0.0f; best.v = 0.0f; best.slot = -1;

    // A zero direction component would give inf * 0 = NaN in the slab test; clamp it to a tiny finite slope instead
    // (the slab then evaluates to +-huge with the sign of (plane - origin), which is the exact answer for a parallel ray).
    constexpr float kTiny = 1e-20f;
    const float idx = 1.0f / (fabsf(dx) > kTiny ? dx : copysignf(kTiny, dx));
    const float idy = 1.0f / (fabsf(dy) > kTiny ? dy : copysignf(kTiny, dy));
    const float idz = 1.0f / (fabsf(dz) > kTiny ? dz : copysignf(kTiny, dz));
    const bool negx = idx < 0.0f, negy = idy < 0.0f, negz = idz < 0.0f;
    // slot bit a set <=> child on the + side of axis a; (slot ^ octinv) is large for children on the near side
    const unsigned octinv = (negx ? 0u : 1u) | (negy ? 0u : 2u) | (negz ? 0u : 4u);
    const unsigned octinv4 = octinv * 0x01010101u;

    uint2 stack[kStack8];
    int sp = 0;
    // node group: x = index of the first inner child, y = hit bits of the children still to visit (bits 31..24,
    // priority order) | imask of the parent (bits 7..0).  The root is "the only child" of a virtual parent.
    uint2 ngroup = make_uint2(0u, 0x80000000u);
    bool alive = active;

    while (CONVERGE ? __any_sync(0xffffffffu, alive) : alive) {
        if (alive) {
            // ---- next child in front-to-back order
            const unsigned hits = ngroup.y;
            const unsigned bit = 31u - (unsigned)__clz(hits);
            ngroup.y = hits & ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            const unsigned slot = (bit - 24u) ^ octinv;
            const unsigned rel = (unsigned)__popc(hits & ~(0xffffffffu << slot));  // inner children stored before this one
            const uint4 *node = bvh.wide_nodes + 5 * (size_t)(ngroup.x + rel);
            BTRC_TRAV_COUNT(0);
            const uint4 v0 = __ldg(node), v1 = __ldg(node + 1), v2 = __ldg(node + 2), v3 = __ldg(node + 3), v4 = __ldg(node + 4);

            // ---- eight quantised child boxes: plane = p + q * 2^e  =>  t = q * (2^e / d) + (p - o) / d
            //      evaluated as (2^23 + q) * s + (b - 2^23 * s): the rounding of the offset moves a plane by at most half a
            //      quantum, which the builder's one-quantum widening of every child box covers
            const float sx = __uint_as_float((v0.w & 0xffu) << 23) * idx;
            const float sy = __uint_as_float(((v0.w >> 8) & 0xffu) << 23) * idy;
            const float sz = __uint_as_float(((v0.w >> 16) & 0xffu) << 23) * idz;
            const float bx = fmaf(-8388608.0f, sx, (__uint_as_float(v0.x) - ox) * idx);
            const float by = fmaf(-8388608.0f, sy, (__uint_as_float(v0.y) - oy) * idy);
            const float bz = fmaf(-8388608.0f, sz, (__uint_as_float(v0.z) - oz) * idz);
            unsigned hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const unsigned meta4 = half ? v1.w : v1.z;
                const unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const unsigned inner_mask4 = (is_inner4 >> 4) * 0xffu;
                const unsigned bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
                const unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
                const unsigned qlox = half ? v2.y : v2.x, qloy = half ? v2.w : v2.z, qloz = half ? v3.y : v3.x;
                const unsigned qhix = half ? v3.w : v3.z, qhiy = half ? v4.y : v4.x, qhiz = half ? v4.w : v4.z;
                const unsigned nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
                const unsigned ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
                const unsigned nz = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned sel = 0x7650u + (unsigned)j;
                    const float tnx = fmaf(u8f(nx, sel), sx, bx), tfx = fmaf(u8f(fx, sel), sx, bx);
                    const float tny = fmaf(u8f(ny, sel), sy, by), tfy = fmaf(u8f(fy, sel), sy, by);
                    const float tnz = fmaf(u8f(nz, sel), sz, bz), tfz = fmaf(u8f(fz, sel), sz, bz);
                    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
                    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, best.t));
                    if (tn <= tf) hitmask |= ((child_bits4 >> (8 * j)) & 0xffu) << ((bit_index4 >> (8 * j)) & 0xffu);
                }
            }
            ngroup.x = v1.x;
            ngroup.y = (hitmask & 0xff000000u) | (v0.w >> 24);
            unsigned tmask = hitmask & 0x00ffffffu;

            // ---- triangles of the leaf children that were hit
            while (tmask) {
                const int j = 31 - __clz(tmask);
                tmask &= ~(1u << j);
                const int s = (int)v1.y + j;
                float t, u, v;
                BTRC_TRAV_COUNT(1);
                if (tri_test(bvh.tris + 3 * (size_t)s, ox, oy, oz, dx, dy, dz, best.t, t, u, v)) {
                    if (ANY_HIT) {
                        best.slot = s;
                        alive = false;
                        break;
                    }
                    if (t < best.t || best.slot < 0 || id_less(bvh.tri_ids, s, best.slot)) {
                        best.t = t; best.u = u; best.v = v; best.slot = s;
                    }
                }
            }
            if (alive && ngroup.y <= 0x00ffffffu) {
                if (sp == 0) alive = false;
                else ngroup = stack[--sp];
            }
        }
    }
    return best;
}


}  // namespace btrc
