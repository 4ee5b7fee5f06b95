// Host BVH construction over the flattened world-space triangle list.
//
// The reference hands triangles to OptiX (utils/optix/context.cpp:158-215: GAS per mesh, IAS over instances,
// PREFER_FAST_TRACE); there is no reference BVH layout to match.  We build our own binary SAH hierarchy
// (binned sweep, OpenMP tasks) and lay it out for the sm_100a traversal kernels (types.h: BvhDevice).
#pragma once
#include <cstdint>
#include <vector>

#include "host_math.h"

namespace btrc {

struct BuildNode {
    AABB box;
    int left = -1, right = -1;  // children (inner) or -1
    int first = 0, count = 0;   // primitive range in `order` (leaf)
};

struct BinaryBvh {
    std::vector<BuildNode> nodes;  // nodes[0] is the root
    std::vector<int> order;        // primitive permutation
};

// boxes/centers: one per primitive.  max_leaf: largest allowed leaf.
// force_leaf: every range of <= max_leaf primitives becomes a leaf without asking the SAH (wide layout).
BinaryBvh build_binary_bvh(const std::vector<AABB> &boxes, const std::vector<Vec3> &centers, int max_leaf, bool force_leaf = false);

// World-space triangle as the traversal kernels consume it (lib/bvh/include/bvh/triangle.hpp:27-34 convention:
// e1 = p0 - p1, e2 = p2 - p0, n = cross(e1, e2)).
struct WorldTri {
    Vec3 p0, e1, e2, n;
    Vec3 p1, p2;
    uint32_t inst, prim;
};

// Conservative triangle boxes for the builder: the box of the "effective" triangle the kernels test
// (p0, p0 - e1, p0 + e2) padded by 1e-5 of its extent + 4e-7 of the largest scene coordinate, so that the fp32 slab
// test can never cull a triangle the exact test accepts.
void padded_triangle_boxes(const std::vector<WorldTri> &tris, float coord_max, std::vector<AABB> &boxes, std::vector<Vec3> &centers);

struct FlatBvh2 {
    std::vector<float> nodes;     // 16 floats per inner node
    std::vector<float> tris;      // 12 floats per triangle slot
    std::vector<uint32_t> tri_ids;  // 2 per triangle slot
    int root_ref = 0;
};

FlatBvh2 flatten_bvh2(const BinaryBvh &bvh, const std::vector<WorldTri> &tris);

// ---- compressed 8-wide layout ("bvh_layout": "wide8") ---------------------------------------------------------------------
// The "compact wide BVH" of BASELINE's north_star, after Ylitie, Karras, Laine 2017: the binary SAH tree above, collapsed into
// 8-wide nodes by the SAH-optimal dynamic programme, child boxes quantised to 8 bits in a per-node frame.
// 80-byte nodes = five 16-byte vectors
//   v0 = (p.x, p.y, p.z, ex | ey << 8 | ez << 16 | imask << 24)    quantisation origin, per-axis scale exponents (biased),
//                                                                     bit i of imask set <=> slot i holds an inner node
//   v1 = (child_base, tri_base, meta[0..3], meta[4..7])             first inner child / first triangle slot of this node
//   v2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])        8-bit child boxes: plane = p + q * 2^(e - 127)
//   v3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])
//   v4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// meta[i]: 0 = empty slot; inner child: 0b001'11000 | i; leaf: (unary triangle count 0b001/0b011/0b111) << 5 | offset of
// its first triangle from tri_base (< 24).  Inner children are stored contiguously in slot order from child_base,
// triangles of the leaf children contiguously from tri_base.  Children are assigned to slots by octant so that
// (slot ^ ray octant) orders them front to back.  10 M triangles: 0.6 GB of binary nodes -> 0.11 GB.
struct WideBvh8 {
    std::vector<uint32_t> nodes;    // 20 words per node, node 0 = root
    std::vector<float> tris;        // 12 floats per triangle slot
    std::vector<uint32_t> tri_ids;  // 2 per triangle slot
    int depth = 0;                  // levels of wide nodes
};

// `bvh` must have leaves of at most 3 primitives (build_binary_bvh(boxes, centers, 3, true)).
WideBvh8 build_wide_bvh8(const BinaryBvh &bvh, const std::vector<WorldTri> &tris);

// ---- two-level hierarchy for instanced scenes --------------------------------------------------------------------------------
// The reference builds one GAS per mesh and an IAS over the instances (utils/optix/context.cpp:158-294, core/scene.cpp:196-209).
// Flattening instances to world space multiplies the node array by the instance count (BASELINE config 4: 250 x 40,000 triangles
// = 0.6 GB of nodes, five times the L2); here every mesh keeps ONE object-space tree (config 4: 4.5 MB, L2-resident) under a small
// tree over the instances' world boxes.  The walk transforms the ray into object space for the BOX tests only; triangles are
// still tested in world space on the flattened world-space triangle records (48 B per instance triangle, laid out per instance in
// the mesh tree's leaf order), so t / u / v and the tie rule stay bit-identical to the flat tree and to the oracle.
struct InstanceBlas {      // 64 bytes = 4 float4 on the device
    float w2o[12];         // rows of the world -> object matrix (3 x 4)
    int root_ref;          // root of the mesh's tree (inner index into `bl_nodes`, or a leaf reference)
    int tri_base;          // first world-space triangle slot of this instance
    int pad[2];
};

struct InstanceRef {
    int mesh;
    float inv[12];         // world -> object, 3 x 4 rows
    size_t first_tri;      // index of the instance's first triangle in the world triangle list (mesh order)
};

struct TwoLevelBvh {
    std::vector<float> tl_nodes;   // 16 floats per inner node of the instance tree
    int tl_root_ref = 0;           // leaf reference = ~(instance << 3)
    std::vector<float> bl_nodes;   // mesh trees, concatenated
    std::vector<InstanceBlas> inst;
    std::vector<float> tris;       // 12 floats per world-space triangle slot
    std::vector<uint32_t> tri_ids;
    int depth = 0;                 // instance-tree depth + deepest mesh tree + 1 (the level marker)
};

// mesh_tris[m]: object-space triangles of mesh m (prim order); world_tris: all instances' world-space triangles, instance-major.
TwoLevelBvh build_two_level_bvh(const std::vector<std::vector<WorldTri>> &mesh_tris, const std::vector<InstanceRef> &insts,
                                const std::vector<WorldTri> &world_tris, float coord_max);
int tree_depth(const BinaryBvh &bvh);

}  // namespace btrc
