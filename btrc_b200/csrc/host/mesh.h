// Triangle-mesh loading and per-triangle shading tables ("triangle_mesh" geometry).
//
// Behaviour mirrors: utils/triangle_mesh_loader.cpp:10-149 (tinyobj with triangulate=true, face order kept,
// per-corner normals / tex-coords with the geometric normal as fall-back), builtin/geometry/triangle_mesh.cpp:24-171
// (six float4 tables per triangle, area alias table, total area, bbox) and utils/math/alias.cpp:13-71.
#pragma once
#include <string>
#include <vector>

#include "host_math.h"

namespace btrc {

struct AliasUnit { float accept_prob; uint32_t another_idx; };
std::vector<AliasUnit> build_alias_table(const std::vector<float> &probs);

struct MeshData {
    int ntri = 0;
    std::vector<Vec3> pos;  // 3 per triangle, local space
    std::vector<float> gx_ua, gy_uba, gz_uca, sn_va, sn_vba, sn_vca;  // float4 per triangle each
    std::vector<float> a_ba_ca;  // 9 floats per triangle
    std::vector<AliasUnit> area_alias;
    float total_area = 0;
    AABB bbox;
};

// Loads an OBJ file (v / vt / vn / f; triangles and quads; negative indices).
MeshData load_triangle_mesh(const std::string &filename, bool transform_to_unit_cube);

}  // namespace btrc
