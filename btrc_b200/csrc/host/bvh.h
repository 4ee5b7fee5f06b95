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
BinaryBvh build_binary_bvh(const std::vector<AABB> &boxes, const std::vector<Vec3> &centers, int max_leaf);

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

}  // namespace btrc
