// ORACLE (test infrastructure only): thin C-ABI driver around the REFERENCE'S OWN host-side sources, compiled from
// /root/reference where they lie by oracle/Makefile (target `ref`) into oracle/_ref/libbtrcref.so.  No reference source is
// copied into this repo.  What it exposes pins the parts of the path that are plain host C++ in the reference:
//   * AliasTable::initialize                    (src/common/btrc/utils/math/alias.cpp:13-71)   -> filter / light / IBL tables
//   * TriangleMeshLoader                        (src/common/btrc/utils/triangle_mesh_loader.cpp:10-149, vendored tinyobj)
//                                               -> triangle order (= primitive ids), positions, geometry ex / ez, normals, uv
//   * PRIME_TABLE                              (src/common/btrc/utils/prime.cpp)              -> Halton bases
//   * Transform3D factories and operator*        (src/common/btrc/utils/math/{transform,mat4,mat3}.cpp) -> instance transforms
// (factory/node/node.cpp does not compile with gcc -- it reaches CUJ class macros with in-class explicit specialisations -- so the
// JSON parser stays pinned by the oracle's own restatement and tests.)
// The reference's device code (CUJ) and OptiX cannot be built here; those stay pinned by known-answer vectors.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <btrc/utils/math/alias.h>
#include <btrc/utils/math/transform.h>
#include <btrc/utils/prime.h>
#include <btrc/utils/triangle_mesh_loader.h>

using namespace btrc;


static thread_local std::string g_err;

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

int ref_prime_table_size() { return (int)PRIME_TABLE_SIZE; }
int ref_prime(int i) { return PRIME_TABLE[i]; }  // utils/prime.cpp (bases of the Halton sampler)

void ref_alias_build(int n, const float *probs, float *accept, uint32_t *another) {
    std::vector<float> p(probs, probs + n);
    AliasTable t{std::span<float>(p)};
    auto tab = t.get_table();
    for (int i = 0; i < n; ++i) { accept[i] = tab[i].accept_prob; another[i] = tab[i].another_idx; }
}

struct ref_mesh { TriangleMeshLoader loader; };

ref_mesh *ref_mesh_load(const char *filename, int transform_to_unit_cube) {
    try {
        auto *m = new ref_mesh{TriangleMeshLoader(filename)};
        if (transform_to_unit_cube) m->loader.transform_to_unit_cube();   // builtin/geometry/triangle_mesh.cpp:24-38 order
        m->loader.remove_indices();
        return m;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void ref_mesh_free(ref_mesh *m) { delete m; }
int ref_mesh_prim_count(ref_mesh *m) { return (int)m->loader.get_primitive_count(); }
// positions / tex coords / interpolated normals: 3 per primitive; geometry ex / ez: 1 per primitive
void ref_mesh_get(ref_mesh *m, float *pos9, float *uv6, float *interp_ez9, float *geo_ex3, float *geo_ez3) {
    const size_t n = m->loader.get_primitive_count();
    auto P = m->loader.get_positions(); auto T = m->loader.get_tex_coords(); auto N = m->loader.get_interp_ezs();
    auto X = m->loader.get_geometry_exs(); auto Z = m->loader.get_geometry_ezs();
    for (size_t i = 0; i < 3 * n; ++i) {
        pos9[3 * i] = P[i].x; pos9[3 * i + 1] = P[i].y; pos9[3 * i + 2] = P[i].z;
        uv6[2 * i] = T[i].x; uv6[2 * i + 1] = T[i].y;
        interp_ez9[3 * i] = N[i].x; interp_ez9[3 * i + 1] = N[i].y; interp_ez9[3 * i + 2] = N[i].z;
    }
    for (size_t i = 0; i < n; ++i) {
        geo_ex3[3 * i] = X[i].x; geo_ex3[3 * i + 1] = X[i].y; geo_ex3[3 * i + 2] = X[i].z;
        geo_ez3[3 * i] = Z[i].x; geo_ez3[3 * i + 1] = Z[i].y; geo_ez3[3 * i + 2] = Z[i].z;
    }
}

// Transform3D factories and composition (utils/math/transform.cpp:66-130): op 0 translate(x,y,z), 1 rotate(axis xyz, rad), 2 rotate_x(rad),
// 3 rotate_y, 4 rotate_z, 5 scale(x,y,z); the ops are multiplied in the order the scene file lists them (node.cpp:188-201: t = op * t)
void ref_transform_compose(int n_ops, const int *ops, const float *args4, float *mat16, float *inv16) {
    Transform3D t;
    for (int i = 0; i < n_ops; ++i) {
        const float *a = args4 + 4 * i;
        Transform3D op;
        switch (ops[i]) {
        case 0: op = Transform3D::translate(a[0], a[1], a[2]); break;
        case 1: op = Transform3D::rotate(Vec3f(a[0], a[1], a[2]), a[3]); break;
        case 2: op = Transform3D::rotate_x(a[0]); break;
        case 3: op = Transform3D::rotate_y(a[0]); break;
        case 4: op = Transform3D::rotate_z(a[0]); break;
        default: op = Transform3D::scale(a[0], a[1], a[2]); break;
        }
        t = op * t;
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { mat16[4 * r + c] = t.mat.at(r, c); inv16[4 * r + c] = t.inv.at(r, c); }
}

}  // extern "C"
