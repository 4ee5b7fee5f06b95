// Device-side BVH construction (LBVH): Morton codes of the padded triangle boxes' centres, radix sort, the binary radix tree
// of Karras 2012 ("Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees") built with one thread per
// internal node, bottom-up box fit with one atomic counter per node, and emission of the traversal kernels' 64-byte nodes
// with every subtree of <= kLbvhLeaf triangles collapsed into one leaf reference (a radix-tree node covers a contiguous range
// of the sorted triangles, so a collapsed subtree is just (first, count)).
//
// Replaces the reference's OptiX accel build (utils/optix/context.cpp:158-215) for scenes where the host SAH builder
// (host/bvh.cpp) is too slow; selected with the renderer parameter "bvh_builder": "lbvh".  Any conservative BVH gives the same
// hits (traverse.cuh: the result depends only on the triangle test and the tie rule), so parity does not depend on the builder.
//
// The per-element functions below are plain functions of their index; lbvh.cu launches them as kernels, and
// tests/host/host_traverse.cpp runs the very same functions in host loops so that the builder is checked without a GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../common/types.h"

#ifndef BTRC_TRAV_QUAL
#define BTRC_TRAV_QUAL __device__ __forceinline__
#endif
#ifndef BTRC_LBVH_FENCE  // the host harness runs the functions in sequential loops and defines these three itself
#define BTRC_LBVH_FENCE() __threadfence()
#define BTRC_LBVH_ATOMIC_INC(p) atomicAdd(p, 1)
#define BTRC_LBVH_LOAD(p) __ldcg(p)  // boxes written by other threads of the same launch: bypass the (non-coherent) L1
#endif

namespace btrc {

constexpr int kLbvhLeaf = 4;  // largest leaf (the node format allows 8)

struct LbvhBox { float lo[3], hi[3]; };

struct LbvhBuild {
    int n;                         // triangles
    const float4 *tris_in;         // 3 float4 per triangle, input order: (p0 | n.x) (e1 | n.y) (e2 | n.z)
    const uint2 *ids_in;           // (instance, primitive), input order
    float scene_lo[3], scene_inv_ext[3];  // Morton grid
    float coord_max;               // largest |coordinate| of the scene (box padding, as host/bvh.cpp padded_triangle_boxes)
    // per triangle
    uint32_t *code;                // Morton code (input order; sorted in place together with `index`)
    uint32_t *index;               // triangle index
    LbvhBox *box_in;               // padded box, input order
    // sorted order
    float4 *tris;                  // output triangle array of the BVH
    uint2 *ids;
    // radix tree: internal nodes [0, n-1)
    int2 *child;                   // (left, right): >= 0 internal node, < 0: ~leaf (sorted triangle slot)
    int2 *range;                   // (first, last) sorted triangle slots covered, inclusive
    int *parent;                   // [0, n-1): parent of internal node i; [n-1, 2n-1): parent of leaf (i - (n-1))
    LbvhBox *node_box;             // box of internal node i
    int *visit;                    // arrival counters of the bottom-up pass
    int *new_index;                // compact index of the live internal nodes (exclusive scan of `live`)
    int *live;                     // 1 if the node's range holds more than kLbvhLeaf triangles
    float4 *nodes_out;             // 4 float4 per live node (types.h: BvhDevice layout)
};

// spreads the low 10 bits of v over every third bit
BTRC_TRAV_QUAL uint32_t lbvh_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// step 1: padded box and 30-bit Morton code of triangle i
BTRC_TRAV_QUAL void lbvh_box_and_code(const LbvhBuild &b, int i) {
    const float4 a = b.tris_in[3 * (size_t)i], e1 = b.tris_in[3 * (size_t)i + 1], e2 = b.tris_in[3 * (size_t)i + 2];
    // the triangle the kernels test: p0, p0 - e1, p0 + e2
    const float p[3][3] = {{a.x, a.y, a.z}, {a.x - e1.x, a.y - e1.y, a.z - e1.z}, {a.x + e2.x, a.y + e2.y, a.z + e2.z}};
    LbvhBox bx;
    float ext = 0.0f;
    for (int k = 0; k < 3; ++k) {
        bx.lo[k] = fminf(p[0][k], fminf(p[1][k], p[2][k]));
        bx.hi[k] = fmaxf(p[0][k], fmaxf(p[1][k], p[2][k]));
        ext = fmaxf(ext, bx.hi[k] - bx.lo[k]);
    }
    // p1 / p2 recomputed from the edges may differ from the mesh vertices in the last bit: twice the host builder's padding
    const float pad = 2.0f * (1e-5f * ext + 4e-7f * b.coord_max);
    uint32_t code = 0;
    for (int k = 0; k < 3; ++k) {
        bx.lo[k] -= pad; bx.hi[k] += pad;
        const float c = (0.5f * (bx.lo[k] + bx.hi[k]) - b.scene_lo[k]) * b.scene_inv_ext[k];
        const uint32_t q = (uint32_t)fminf(fmaxf(c * 1024.0f, 0.0f), 1023.0f);
        code |= lbvh_expand_bits(q) << (2 - k);
    }
    b.box_in[i] = bx;
    b.code[i] = code;
    b.index[i] = (uint32_t)i;
}

// step 3 (after the sort): triangle data in sorted order
BTRC_TRAV_QUAL void lbvh_gather(const LbvhBuild &b, int s) {
    const uint32_t i = b.index[s];
    b.tris[3 * (size_t)s] = b.tris_in[3 * (size_t)i];
    b.tris[3 * (size_t)s + 1] = b.tris_in[3 * (size_t)i + 1];
    b.tris[3 * (size_t)s + 2] = b.tris_in[3 * (size_t)i + 2];
    b.ids[s] = b.ids_in[i];
}

// length of the common prefix of the keys at sorted positions i and j; equal codes are told apart by the position
BTRC_TRAV_QUAL int lbvh_delta(const LbvhBuild &b, int i, int j) {
    if (j < 0 || j >= b.n) return -1;
    const uint32_t ci = b.code[i], cj = b.code[j];
    if (ci == cj) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(ci ^ cj);
}

// step 4: internal node i of the binary radix tree (Karras 2012, section 4)
BTRC_TRAV_QUAL void lbvh_build_node(const LbvhBuild &b, int i) {
    const int d = lbvh_delta(b, i, i + 1) - lbvh_delta(b, i, i - 1) >= 0 ? 1 : -1;
    const int delta_min = lbvh_delta(b, i, i - d);
    int lmax = 2;
    while (lbvh_delta(b, i, i + lmax * d) > delta_min) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(b, i, i + (l + t) * d) > delta_min) l += t;
    const int j = i + l * d;
    const int delta_node = lbvh_delta(b, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(b, i, i + (s + t) * d) > delta_node) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int first = i < j ? i : j, last = i < j ? j : i;
    const int left = first == gamma ? ~gamma : gamma;
    const int right = last == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    b.child[i] = make_int2(left, right);
    b.range[i] = make_int2(first, last);
    b.parent[left >= 0 ? left : (b.n - 1) + ~left] = i;
    b.parent[right >= 0 ? right : (b.n - 1) + ~right] = i;
    if (i == 0) b.parent[0] = -1;
    b.visit[i] = 0;
    b.live[i] = (last - first + 1) > kLbvhLeaf ? 1 : 0;
}

BTRC_TRAV_QUAL LbvhBox lbvh_child_box(const LbvhBuild &b, int ref) {
    const float *p = ref >= 0 ? b.node_box[ref].lo : b.box_in[b.index[~ref]].lo;  // lo[3] and hi[3] are contiguous
    LbvhBox r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = BTRC_LBVH_LOAD(p + k); r.hi[k] = BTRC_LBVH_LOAD(p + 3 + k); }
    return r;
}

// step 5: bottom-up fit starting at sorted triangle s; the second thread to arrive at a node computes its box and goes on
BTRC_TRAV_QUAL void lbvh_fit_from_leaf(const LbvhBuild &b, int s) {
    int node = b.parent[(b.n - 1) + s];
    while (node >= 0) {
        BTRC_LBVH_FENCE();  // the box written below (or by an earlier kernel) is visible before the arrival is counted
        if (BTRC_LBVH_ATOMIC_INC(&b.visit[node]) == 0) return;
        BTRC_LBVH_FENCE();
        const int2 c = b.child[node];
        const LbvhBox x = lbvh_child_box(b, c.x), y = lbvh_child_box(b, c.y);
        LbvhBox u;
        for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(x.lo[k], y.lo[k]); u.hi[k] = fmaxf(x.hi[k], y.hi[k]); }
        b.node_box[node] = u;
        node = b.parent[node];
    }
}

// reference of a child as the traversal kernels expect it, and its box
BTRC_TRAV_QUAL int lbvh_child_ref(const LbvhBuild &b, int c, LbvhBox &box) {
    if (c < 0) { box = b.box_in[b.index[~c]]; return ~((~c) << 3); }                                       // one triangle
    box = b.node_box[c];
    if (!b.live[c]) { const int2 r = b.range[c]; return ~((r.x << 3) | (r.y - r.x)); }                      // collapsed subtree
    return b.new_index[c];
}

// step 6: the 64-byte node of live internal node i (types.h: BvhDevice; boxes widened by one ulp like host/bvh.cpp flatten_bvh2)
BTRC_TRAV_QUAL void lbvh_emit_node(const LbvhBuild &b, int i) {
    if (!b.live[i]) return;
    const int2 c = b.child[i];
    LbvhBox x, y;
    const int r0 = lbvh_child_ref(b, c.x, x), r1 = lbvh_child_ref(b, c.y, y);
    float4 *o = b.nodes_out + 4 * (size_t)b.new_index[i];
    const float dn = -3.402823466e+38f, up = 3.402823466e+38f;
    o[0] = make_float4(nextafterf(x.lo[0], dn), nextafterf(x.hi[0], up), nextafterf(x.lo[1], dn), nextafterf(x.hi[1], up));
    o[1] = make_float4(nextafterf(y.lo[0], dn), nextafterf(y.hi[0], up), nextafterf(y.lo[1], dn), nextafterf(y.hi[1], up));
    o[2] = make_float4(nextafterf(x.lo[2], dn), nextafterf(x.hi[2], up), nextafterf(y.lo[2], dn), nextafterf(y.hi[2], up));
    o[3] = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
}

}  // namespace btrc
