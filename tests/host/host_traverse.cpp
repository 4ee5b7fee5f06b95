// TEST INFRASTRUCTURE (never shipped, never on the product path): compiles the product's BVH builder
// (btrc_b200/csrc/host/bvh.cpp) and the product's traversal function (btrc_b200/csrc/device/traverse.cuh) for the HOST,
// so that the exact code the sm_100a kernels run can be checked against the oracle in the CPU-only test suite.
// The intrinsics the device header uses are supplied here with their IEEE meaning (g++ -ffp-contract=off).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>  // vector types only

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __uint_as_float(unsigned x) { float f; std::memcpy(&f, &x, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned x; std::memcpy(&x, &f, 4); return x; }
static inline int __float_as_int(float f) { int x; std::memcpy(&x, &f, 4); return x; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }

static thread_local uint64_t g_count[2];
#define BTRC_TRAV_COUNT(which) (++g_count[which])
#define BTRC_TRAV_QUAL static inline
#define BTRC_LBVH_FENCE() ((void)0)
#define BTRC_LBVH_ATOMIC_INC(p) ((*(p))++)
#define BTRC_LBVH_LOAD(p) (*(p))
static inline float __int_as_float(int x) { float f; std::memcpy(&f, &x, 4); return f; }
#include "../../btrc_b200/csrc/device/traverse.cuh"
#include "../../btrc_b200/csrc/device/lbvh.cuh"
#include "../../btrc_b200/csrc/host/bvh.h"
#include <algorithm>
#include <numeric>

using namespace btrc;

#ifndef HT_TRAVERSE
#define HT_TRAVERSE traverse_bvh2   // -DHT_TRAVERSE=traverse_bvh2_ifif builds the if-if form
#endif

struct HostBvh {
    FlatBvh2 w;
    int n_inner = 0;
};

static BvhDevice view(const HostBvh *h) {
    BvhDevice b{};
    b.nodes = reinterpret_cast<const float4 *>(h->w.nodes.data());
    b.tris = reinterpret_cast<const float4 *>(h->w.tris.data());
    b.tri_ids = reinterpret_cast<const uint2 *>(h->w.tri_ids.data());
    b.root_ref = h->w.root_ref;
    b.n_tris = (int)(h->w.tri_ids.size() / 2);
    return b;
}

extern "C" {

// verts: n x 9 floats (p0, p1, p2 world space); ids: n x 2 (instance, primitive)
void *ht_build(int n, const float *verts, const uint32_t *ids) {
    std::vector<WorldTri> tris(n);
    AABB wb;
    for (int i = 0; i < n; ++i) {
        WorldTri &t = tris[i];
        t.p0 = Vec3(verts[9 * i], verts[9 * i + 1], verts[9 * i + 2]);
        t.p1 = Vec3(verts[9 * i + 3], verts[9 * i + 4], verts[9 * i + 5]);
        t.p2 = Vec3(verts[9 * i + 6], verts[9 * i + 7], verts[9 * i + 8]);
        t.e1 = t.p0 - t.p1;
        t.e2 = t.p2 - t.p0;
        t.n = cross(t.e1, t.e2);
        t.inst = ids[2 * i];
        t.prim = ids[2 * i + 1];
        wb.extend(t.p0); wb.extend(t.p1); wb.extend(t.p2);
    }
    float coord_max = 0;
    for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    std::vector<AABB> boxes;
    std::vector<Vec3> centers;
    padded_triangle_boxes(tris, coord_max, boxes, centers);
    BinaryBvh bb = build_binary_bvh(boxes, centers, 4);  // as Scene::commit does
    auto *h = new HostBvh;
    h->w = flatten_bvh2(bb, tris);
    h->n_inner = (int)(h->w.nodes.size() / 16);
    return h;
}

// The device LBVH builder (device/lbvh.cuh) run in host loops: same per-element functions, std::stable_sort for the radix sort
void *ht_build_lbvh(int n, const float *verts, const uint32_t *ids) {
    std::vector<float4> tris_in(3 * (size_t)n);
    std::vector<uint2> ids_in(n);
    AABB wb;
    for (int i = 0; i < n; ++i) {
        const Vec3 p0(verts[9 * i], verts[9 * i + 1], verts[9 * i + 2]), p1(verts[9 * i + 3], verts[9 * i + 4], verts[9 * i + 5]),
            p2(verts[9 * i + 6], verts[9 * i + 7], verts[9 * i + 8]);
        const Vec3 e1 = p0 - p1, e2 = p2 - p0, nn = cross(e1, e2);
        tris_in[3 * i] = make_float4(p0.x, p0.y, p0.z, nn.x);
        tris_in[3 * i + 1] = make_float4(e1.x, e1.y, e1.z, nn.y);
        tris_in[3 * i + 2] = make_float4(e2.x, e2.y, e2.z, nn.z);
        ids_in[i] = make_uint2(ids[2 * i], ids[2 * i + 1]);
        wb.extend(p0); wb.extend(p1); wb.extend(p2);
    }
    float cm = 0;
    for (int a = 0; a < 3; ++a) cm = std::max(cm, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    const int ni = std::max(n - 1, 1);
    std::vector<uint32_t> code(n), index(n);
    std::vector<LbvhBox> box_in(n), node_box(ni);
    std::vector<int2> child(ni), range(ni);
    std::vector<int> parent(2 * (size_t)n), visit(ni), live(ni), new_index(ni);
    auto *h = new HostBvh;
    h->w.tris.resize(12 * (size_t)n);
    h->w.tri_ids.resize(2 * (size_t)n);
    LbvhBuild b{};
    b.n = n; b.tris_in = tris_in.data(); b.ids_in = ids_in.data(); b.coord_max = cm;
    for (int k = 0; k < 3; ++k) { b.scene_lo[k] = wb.lo[k]; const float e = wb.hi[k] - wb.lo[k]; b.scene_inv_ext[k] = e > 0 ? 1.0f / e : 0.0f; }
    b.code = code.data(); b.index = index.data(); b.box_in = box_in.data();
    b.tris = reinterpret_cast<float4 *>(h->w.tris.data()); b.ids = reinterpret_cast<uint2 *>(h->w.tri_ids.data());
    b.child = child.data(); b.range = range.data(); b.parent = parent.data(); b.node_box = node_box.data();
    b.visit = visit.data(); b.live = live.data(); b.new_index = new_index.data();
    for (int i = 0; i < n; ++i) lbvh_box_and_code(b, i);
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) { return code[x] < code[y]; });
    std::vector<uint32_t> code2(n), index2(n);
    for (int s = 0; s < n; ++s) { code2[s] = code[perm[s]]; index2[s] = index[perm[s]]; }
    b.code = code2.data(); b.index = index2.data();
    for (int s = 0; s < n; ++s) lbvh_gather(b, s);
    if (n <= kLbvhLeaf) { h->w.root_ref = ~(n - 1); h->w.nodes.assign(16, 0.0f); return h; }
    for (int i = 0; i < n - 1; ++i) lbvh_build_node(b, i);
    for (int s = 0; s < n; ++s) lbvh_fit_from_leaf(b, s);
    int acc = 0;
    for (int i = 0; i < n - 1; ++i) { new_index[i] = acc; acc += live[i]; }
    h->w.nodes.assign(16 * (size_t)std::max(acc, 1), 0.0f);
    b.nodes_out = reinterpret_cast<float4 *>(h->w.nodes.data());
    for (int i = 0; i < n - 1; ++i) lbvh_emit_node(b, i);
    h->w.root_ref = 0;
    h->n_inner = acc;
    return h;
}

// The two-level layout of instanced scenes (host/bvh.h build_two_level_bvh + traverse_two_level), from what Scene::commit hands
// the builder: object-space meshes, the instances' mesh ids and world -> object matrices, and the world-space triangles.
struct HostTwoLevel { TwoLevelBvh w; };

static WorldTri make_tri(const float *v, uint32_t inst, uint32_t prim) {
    WorldTri t;
    t.p0 = Vec3(v[0], v[1], v[2]); t.p1 = Vec3(v[3], v[4], v[5]); t.p2 = Vec3(v[6], v[7], v[8]);
    t.e1 = t.p0 - t.p1; t.e2 = t.p2 - t.p0; t.n = cross(t.e1, t.e2);
    t.inst = inst; t.prim = prim;
    return t;
}

void *ht2_build(int n_mesh, const int *mesh_ntri, const float *mesh_verts, int n_inst, const int *inst_mesh, const float *inst_inv,
                int n_world, const float *world_verts, const uint32_t *world_ids) {
    std::vector<std::vector<WorldTri>> mesh_tris(n_mesh);
    size_t off = 0;
    for (int m = 0; m < n_mesh; ++m) {
        for (int k = 0; k < mesh_ntri[m]; ++k) mesh_tris[m].push_back(make_tri(mesh_verts + 9 * (off + k), 0, (uint32_t)k));
        off += (size_t)mesh_ntri[m];
    }
    std::vector<WorldTri> world(n_world);
    AABB wb;
    for (int i = 0; i < n_world; ++i) {
        world[i] = make_tri(world_verts + 9 * (size_t)i, world_ids[2 * i], world_ids[2 * i + 1]);
        wb.extend(world[i].p0); wb.extend(world[i].p1); wb.extend(world[i].p2);
    }
    std::vector<InstanceRef> refs(n_inst);
    size_t first = 0;
    for (int i = 0; i < n_inst; ++i) {
        refs[i].mesh = inst_mesh[i];
        std::memcpy(refs[i].inv, inst_inv + 12 * (size_t)i, 48);
        refs[i].first_tri = first;
        first += mesh_tris[inst_mesh[i]].size();
    }
    if ((int)first != n_world) return nullptr;
    float coord_max = 0;
    for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    auto *h = new HostTwoLevel;
    h->w = build_two_level_bvh(mesh_tris, refs, world, coord_max);
    return h;
}
void ht2_free(void *h) { delete static_cast<HostTwoLevel *>(h); }
int ht2_info(void *hh, int which) {
    const HostTwoLevel *h = static_cast<HostTwoLevel *>(hh);
    return which == 0 ? (int)(h->w.tl_nodes.size() / 16) : which == 1 ? (int)(h->w.bl_nodes.size() / 16) : h->w.depth;
}

static BvhDevice view2(const HostTwoLevel *h) {
    BvhDevice b{};
    b.nodes = reinterpret_cast<const float4 *>(h->w.bl_nodes.data());
    b.tl_nodes = reinterpret_cast<const float4 *>(h->w.tl_nodes.data());
    b.inst_blas = reinterpret_cast<const float4 *>(h->w.inst.data());
    b.tris = reinterpret_cast<const float4 *>(h->w.tris.data());
    b.tri_ids = reinterpret_cast<const uint2 *>(h->w.tri_ids.data());
    b.root_ref = h->w.tl_root_ref;
    b.n_tris = (int)(h->w.tri_ids.size() / 2);
    b.two_level = 1;
    return b;
}

void ht2_closest(void *hh, int n, const float *o_med, const float *d_t1, uint32_t *out_ids, float *out_tuv, uint64_t *steps) {
    const HostTwoLevel *h = static_cast<HostTwoLevel *>(hh);
    const BvhDevice b = view2(h);
    uint64_t n_nodes = 0, n_tests = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_nodes, n_tests)
    for (int i = 0; i < n; ++i) {
        g_count[0] = g_count[1] = 0;
        const HitRec r = traverse_two_level<false, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1],
                                                          d_t1[4 * i + 2], d_t1[4 * i + 3]);
        if (r.slot < 0) {
            out_ids[2 * i] = out_ids[2 * i + 1] = 0xffffffffu;
            out_tuv[3 * i] = out_tuv[3 * i + 1] = out_tuv[3 * i + 2] = 0;
        } else {
            out_ids[2 * i] = h->w.tri_ids[2 * r.slot];
            out_ids[2 * i + 1] = h->w.tri_ids[2 * r.slot + 1];
            out_tuv[3 * i] = r.t; out_tuv[3 * i + 1] = r.u; out_tuv[3 * i + 2] = r.v;
        }
        n_nodes += g_count[0];
        n_tests += g_count[1];
    }
    if (steps) { steps[0] = n_nodes; steps[1] = n_tests; }
}

void ht2_any(void *hh, int n, const float *o_med, const float *d_t1, uint8_t *occluded) {
    const BvhDevice b = view2(static_cast<HostTwoLevel *>(hh));
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i)
        occluded[i] = traverse_two_level<true, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2],
                                                      d_t1[4 * i + 3]).slot >= 0;
}

// The compressed 8-wide layout ("bvh_layout": "wide8"): build_binary_bvh(.., 3, true) + build_wide_bvh8 as Scene::commit does,
// walked by traverse_bvh8 (the product's device function compiled for the host).
struct HostWide8 { WideBvh8 w; };

void *ht8_build(int n, const float *verts, const uint32_t *ids) {
    std::vector<WorldTri> tris(n);
    AABB wb;
    for (int i = 0; i < n; ++i) {
        tris[i] = make_tri(verts + 9 * (size_t)i, ids[2 * i], ids[2 * i + 1]);
        wb.extend(tris[i].p0); wb.extend(tris[i].p1); wb.extend(tris[i].p2);
    }
    float coord_max = 0;
    for (int a = 0; a < 3; ++a) coord_max = std::max(coord_max, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    std::vector<AABB> boxes;
    std::vector<Vec3> centers;
    padded_triangle_boxes(tris, coord_max, boxes, centers);
    const BinaryBvh bb = build_binary_bvh(boxes, centers, 3, true);
    auto *h = new HostWide8;
    h->w = build_wide_bvh8(bb, tris);
    return h;
}
void ht8_free(void *h) { delete static_cast<HostWide8 *>(h); }
int ht8_info(void *hh, int which) {  // 0: wide nodes, 1: levels of wide nodes
    const HostWide8 *h = static_cast<HostWide8 *>(hh);
    return which == 0 ? (int)(h->w.nodes.size() / 20) : h->w.depth;
}
static BvhDevice view8(const HostWide8 *h) {
    BvhDevice b{};
    b.wide_nodes = reinterpret_cast<const uint4 *>(h->w.nodes.data());
    b.tris = reinterpret_cast<const float4 *>(h->w.tris.data());
    b.tri_ids = reinterpret_cast<const uint2 *>(h->w.tri_ids.data());
    b.n_tris = (int)(h->w.tri_ids.size() / 2);
    b.wide = 1;
    return b;
}
void ht8_closest(void *hh, int n, const float *o_med, const float *d_t1, uint32_t *out_ids, float *out_tuv, uint64_t *steps) {
    const HostWide8 *h = static_cast<HostWide8 *>(hh);
    const BvhDevice b = view8(h);
    uint64_t n_nodes = 0, n_tests = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_nodes, n_tests)
    for (int i = 0; i < n; ++i) {
        g_count[0] = g_count[1] = 0;
        const HitRec r = traverse_bvh8<false, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1],
                                                     d_t1[4 * i + 2], d_t1[4 * i + 3]);
        if (r.slot < 0) {
            out_ids[2 * i] = out_ids[2 * i + 1] = 0xffffffffu;
            out_tuv[3 * i] = out_tuv[3 * i + 1] = out_tuv[3 * i + 2] = 0;
        } else {
            out_ids[2 * i] = h->w.tri_ids[2 * r.slot];
            out_ids[2 * i + 1] = h->w.tri_ids[2 * r.slot + 1];
            out_tuv[3 * i] = r.t; out_tuv[3 * i + 1] = r.u; out_tuv[3 * i + 2] = r.v;
        }
        n_nodes += g_count[0];
        n_tests += g_count[1];
    }
    if (steps) { steps[0] = n_nodes; steps[1] = n_tests; }
}
void ht8_any(void *hh, int n, const float *o_med, const float *d_t1, uint8_t *occluded) {
    const BvhDevice b = view8(static_cast<HostWide8 *>(hh));
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i)
        occluded[i] = traverse_bvh8<true, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2],
                                                 d_t1[4 * i + 3]).slot >= 0;
}

void ht_free(void *h) { delete static_cast<HostBvh *>(h); }
int ht_num_nodes(void *h) { return static_cast<HostBvh *>(h)->n_inner; }

// closest hit: out_ids n x 2 (instance, primitive; 0xffffffff on a miss), out_tuv n x 3; steps: node visits in total
void ht_closest(void *hh, int n, const float *o_med, const float *d_t1, uint32_t *out_ids, float *out_tuv, uint64_t *steps) {
    const HostBvh *h = static_cast<HostBvh *>(hh);
    const BvhDevice b = view(h);
    uint64_t n_nodes = 0, n_tests = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_nodes, n_tests)
    for (int i = 0; i < n; ++i) {
        g_count[0] = g_count[1] = 0;
        const HitRec r = HT_TRAVERSE<false, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1],
                                                     d_t1[4 * i + 2], d_t1[4 * i + 3]);
        if (r.slot < 0) {
            out_ids[2 * i] = out_ids[2 * i + 1] = 0xffffffffu;
            out_tuv[3 * i] = out_tuv[3 * i + 1] = out_tuv[3 * i + 2] = 0;
        } else {
            out_ids[2 * i] = h->w.tri_ids[2 * r.slot];
            out_ids[2 * i + 1] = h->w.tri_ids[2 * r.slot + 1];
            out_tuv[3 * i] = r.t; out_tuv[3 * i + 1] = r.u; out_tuv[3 * i + 2] = r.v;
        }
        n_nodes += g_count[0];
        n_tests += g_count[1];
    }
    if (steps) { steps[0] = n_nodes; steps[1] = n_tests; }
}

void ht_any(void *hh, int n, const float *o_med, const float *d_t1, uint8_t *occluded) {
    const HostBvh *h = static_cast<HostBvh *>(hh);
    const BvhDevice b = view(h);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i)
        occluded[i] = HT_TRAVERSE<true, false>(b, o_med[4 * i], o_med[4 * i + 1], o_med[4 * i + 2], d_t1[4 * i], d_t1[4 * i + 1], d_t1[4 * i + 2],
                                                 d_t1[4 * i + 3]).slot >= 0;
}

}  // extern "C"
