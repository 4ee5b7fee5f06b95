#include "mesh.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

#include "node.h"

namespace btrc {

std::vector<AliasUnit> build_alias_table(const std::vector<float> &probs) {  // utils/math/alias.cpp:13-57
    if (probs.empty()) throw Error("alias table: empty input");
    float sum = 0.0f;
    for (float p : probs) sum = sum + p;
    const float ratio = probs.size() / sum;
    std::vector<uint32_t> overs, unders;
    std::vector<AliasUnit> units(probs.size());
    for (uint32_t i = 0; i < (uint32_t)units.size(); ++i) {
        const float p = probs[i] * ratio;
        units[i].accept_prob = p;
        units[i].another_idx = i;
        if (p > 1) overs.push_back(i);
        else if (p < 1) unders.push_back(i);
    }
    while (!overs.empty() && !unders.empty()) {
        const uint32_t over = overs.back(), under = unders.back();
        overs.pop_back();
        unders.pop_back();
        units[over].accept_prob -= 1 - units[under].accept_prob;
        units[under].another_idx = over;
        if (units[over].accept_prob > 1) overs.push_back(over);
        else if (units[over].accept_prob < 1) unders.push_back(over);
    }
    for (auto i : overs) units[i].accept_prob = 1;
    for (auto i : unders) units[i].accept_prob = 1;
    return units;
}

namespace {

struct Corner { int v, t, n; };

inline void skip_space(const char *&p) { while (*p == ' ' || *p == '\t') ++p; }

bool parse_corner(const char *&p, Corner &c, int nv, int nt, int nn) {
    skip_space(p);
    if (*p == '\0' || *p == '\n' || *p == '\r' || *p == '#') return false;
    char *e;
    long v = std::strtol(p, &e, 10);
    if (e == p) return false;
    p = e;
    c.v = v > 0 ? int(v - 1) : nv + int(v);
    c.t = c.n = -1;
    if (*p == '/') {
        ++p;
        if (*p != '/') {
            long t = std::strtol(p, &e, 10);
            if (e != p) { c.t = t > 0 ? int(t - 1) : nt + int(t); p = e; }
        }
        if (*p == '/') {
            ++p;
            long n = std::strtol(p, &e, 10);
            if (e != p) { c.n = n > 0 ? int(n - 1) : nn + int(n); p = e; }
        }
    }
    return true;
}

Vec3 frame_from_z_x(Vec3 z) {  // Frame::from_z(nor).x -- utils/math/frame.cpp:45-57; the Frame constructor normalises its axes (frame.cpp:11-15)
    Vec3 nz = normalize(z), ny;
    if (1 - std::fabs(nz.x) < 0.1f) ny = cross(nz, Vec3(0, 1, 0));
    else ny = cross(nz, Vec3(1, 0, 0));
    return normalize(cross(nz, ny));
}

Vec3 triangle_dpdu(Vec3 B_A, Vec3 C_A, Vec2 b_a, Vec2 c_a, Vec3 nor) {  // utils/math/triangle.h:8-22
    const float m00 = b_a.x, m01 = b_a.y, m10 = c_a.x, m11 = c_a.y;
    const float det = m00 * m11 - m01 * m10;
    if (det == 0.0f) return frame_from_z_x(nor);
    const float inv_det = 1 / det;
    return normalize(m11 * inv_det * B_A - m01 * inv_det * C_A);
}

// Polygons with more than four corners: tinyobj's built-in ear clipping (lib/tinyobjloader/tiny_obj_loader.h:1521-1813), restated
// so that triangle order and corner order - the primitive ids - are the reference's.  The polygon is projected on two coordinate
// axes chosen from its first non-degenerate corner; a candidate ear (three consecutive corners from a moving start) is taken
// when its turn agrees in sign with tinyobj's "area" term and no other remaining corner lies inside it (even-odd test); the
// middle corner is then removed.  The search gives up after a full round without progress, as tinyobj does (the rest of such a
// polygon is dropped).  Arithmetic in float, in tinyobj's order.
bool point_in_triangle_2d(const float *vx, const float *vy, float tx, float ty) {  // pnpoly, tiny_obj_loader.h:1387-1398
    bool inside = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) inside = !inside;
    return inside;
}

void ear_clip(const std::vector<Corner> &face, const std::vector<Vec3> &vs, std::vector<Corner> &out) {
    const size_t n = face.size();
    int ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const Vec3 a = vs[face[k].v], b = vs[face[(k + 1) % n].v], c = vs[face[(k + 2) % n].v];
        const Vec3 e0 = b - a, e1 = c - b;
        const float cx = std::fabs(e0.y * e1.z - e0.z * e1.y), cy = std::fabs(e0.z * e1.x - e0.x * e1.z), cz = std::fabs(e0.x * e1.y - e0.y * e1.x);
        const float eps = 1.1920928955078125e-07f;
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                ax0 = 0;
                if (cz > cx && cz > cy) ax1 = 1;
            }
            break;
        }
    }
    auto coord = [&](const Corner &c, int axis) { const Vec3 &v = vs[c.v]; return axis == 0 ? v.x : axis == 1 ? v.y : v.z; };
    std::vector<Corner> rem = face;
    size_t start = 0, budget = n, last_count = n;
    while (rem.size() > 3 && budget > 0) {
        const size_t m = rem.size();
        if (start >= m) start -= m;
        if (last_count != m) { last_count = m; budget = m; }
        else --budget;
        Corner tri[3];
        float vx[3], vy[3];
        for (int k = 0; k < 3; ++k) {
            tri[k] = rem[(start + k) % m];
            vx[k] = coord(tri[k], ax0); vy[k] = coord(tri[k], ax1);
        }
        const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
        const float turn = e0x * e1y - e0y * e1x;
        const float area = (vx[0] * vy[1] - vy[0] * vx[1]) * 0.5f;
        if (turn * area < 0.0f) { ++start; continue; }
        bool blocked = false;
        for (size_t other = 3; other < m && !blocked; ++other) {
            const Corner &q = rem[(start + other) % m];
            blocked = point_in_triangle_2d(vx, vy, coord(q, ax0), coord(q, ax1));
        }
        if (blocked) { ++start; continue; }
        out.push_back(tri[0]); out.push_back(tri[1]); out.push_back(tri[2]);
        rem.erase(rem.begin() + (long)((start + 1) % m));
    }
    if (rem.size() == 3) out.insert(out.end(), rem.begin(), rem.end());
}

}  // namespace

MeshData load_triangle_mesh(const std::string &filename, bool to_unit_cube) {
    const std::string text = read_text_file(filename);
    std::vector<Vec3> vs, vns;
    std::vector<Vec2> vts;
    std::vector<Corner> corners;  // 3 per triangle, face order
    std::vector<Corner> face;
    const char *p = text.c_str();
    while (*p) {
        skip_space(p);
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            char *e;
            float x = (float)std::strtod(p, &e); p = e;
            float y = (float)std::strtod(p, &e); p = e;
            float z = (float)std::strtod(p, &e); p = e;
            vs.emplace_back(x, y, z);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            p += 3;
            char *e;
            float x = (float)std::strtod(p, &e); p = e;
            float y = (float)std::strtod(p, &e); p = e;
            float z = (float)std::strtod(p, &e); p = e;
            vns.emplace_back(x, y, z);
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            p += 3;
            char *e;
            float x = (float)std::strtod(p, &e); p = e;
            float y = (float)std::strtod(p, &e); p = e;
            vts.push_back(Vec2{x, y});
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            face.clear();
            Corner c;
            while (parse_corner(p, c, (int)vs.size(), (int)vts.size(), (int)vns.size())) face.push_back(c);
            for (auto &fc : face)
                if (fc.v < 0 || fc.v >= (int)vs.size()) throw Error("invalid obj vertex index: out of range");
            if (face.size() == 3) {
                corners.insert(corners.end(), face.begin(), face.end());
            } else if (face.size() == 4) {
                // tinyobj quad rule (lib/tinyobjloader/tiny_obj_loader.h:1429-1520): split along the shorter diagonal
                const Vec3 v0 = vs[face[0].v], v1 = vs[face[1].v], v2 = vs[face[2].v], v3 = vs[face[3].v];
                const Vec3 e02 = v2 - v0, e13 = v3 - v1;
                const float sqr02 = e02.x * e02.x + e02.y * e02.y + e02.z * e02.z;
                const float sqr13 = e13.x * e13.x + e13.y * e13.y + e13.z * e13.z;
                if (sqr02 < sqr13) {
                    for (int k : {0, 1, 2, 0, 2, 3}) corners.push_back(face[k]);
                } else {
                    for (int k : {0, 1, 3, 1, 2, 3}) corners.push_back(face[k]);
                }
            } else if (face.size() > 4) {
                ear_clip(face, vs, corners);
            }
        }
        while (*p && *p != '\n') ++p;
        if (*p == '\n') ++p;
    }

    // geometry frames come from the positions as read: the reference computes them in the loader's constructor and
    // rescales the positions afterwards (triangle_mesh_loader.cpp:80-123 vs 193-213; builtin/geometry/triangle_mesh.cpp:26-28)
    const std::vector<Vec3> vs_read = to_unit_cube ? vs : std::vector<Vec3>();
    if (to_unit_cube) {  // triangle_mesh_loader.cpp:193-213
        Vec3 lower(3.402823466e+38f), upper(-3.402823466e+38f);
        for (auto &v : vs) { lower = vmin(lower, v); upper = vmax(upper, v); }
        const Vec3 extent = upper - lower;
        const float extent_max = std::max(extent.x, std::max(extent.y, extent.z));
        const float scale = extent_max > 0 ? 1 / extent_max : 1.0f;
        for (auto &v : vs) v = scale * (v - 0.5f * (lower + upper));
    }

    MeshData m;
    m.ntri = (int)(corners.size() / 3);
    if (m.ntri == 0) throw Error("obj file has no faces: " + filename);
    const size_t nt = (size_t)m.ntri;
    m.pos.resize(3 * nt);
    m.gx_ua.resize(4 * nt); m.gy_uba.resize(4 * nt); m.gz_uca.resize(4 * nt);
    m.sn_va.resize(4 * nt); m.sn_vba.resize(4 * nt); m.sn_vca.resize(4 * nt);
    m.a_ba_ca.resize(9 * nt);
    std::vector<float> areas(nt);
    for (size_t i = 0; i < nt; ++i) {
        const Corner *c = &corners[3 * i];
        const Vec3 a = vs[c[0].v], b = vs[c[1].v], cc = vs[c[2].v];
        m.pos[3 * i] = a; m.pos[3 * i + 1] = b; m.pos[3 * i + 2] = cc;
        const Vec3 fa = to_unit_cube ? vs_read[c[0].v] : a, fb = to_unit_cube ? vs_read[c[1].v] : b, fc = to_unit_cube ? vs_read[c[2].v] : cc;
        const Vec3 geo_z = normalize(cross(fb - fa, fc - fa));  // loader.cpp:86
        Vec3 sz[3];
        Vec2 tc[3];
        for (int k = 0; k < 3; ++k) {
            if (c[k].n < 0) sz[k] = geo_z;
            else {
                if (c[k].n >= (int)vns.size()) throw Error("invalid obj normal index: out of range");
                sz[k] = vns[c[k].n];
                if (sz[k].x == 0.0f && sz[k].y == 0.0f && sz[k].z == 0.0f) sz[k] = geo_z;
            }
            if (c[k].t < 0) tc[k] = Vec2{0, 0};
            else {
                if (c[k].t >= (int)vts.size()) throw Error("invalid obj tex coord index: out of range");
                tc[k] = vts[c[k].t];
            }
        }
        const Vec3 gx = triangle_dpdu(fb - fa, fc - fa, Vec2{tc[1].x - tc[0].x, tc[1].y - tc[0].y},
                                      Vec2{tc[2].x - tc[0].x, tc[2].y - tc[0].y}, geo_z);
        const Vec3 gz = geo_z;
        const Vec3 gy = normalize(cross(gz, gx));  // triangle_mesh.cpp:52
        auto put = [&](std::vector<float> &dst, Vec3 v, float w) {
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = w;
        };
        put(m.gx_ua, gx, tc[0].x);
        put(m.gy_uba, gy, tc[1].x - tc[0].x);
        put(m.gz_uca, gz, tc[2].x - tc[0].x);
        put(m.sn_va, sz[0], tc[0].y);
        put(m.sn_vba, sz[1] - sz[0], tc[1].y - tc[0].y);
        put(m.sn_vca, sz[2] - sz[0], tc[2].y - tc[0].y);
        const Vec3 ba = b - a, ca = cc - a;
        areas[i] = 0.5f * length(cross(ba, ca));  // triangle.h:24-27
        float *o = &m.a_ba_ca[9 * i];
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = ba.x; o[4] = ba.y; o[5] = ba.z; o[6] = ca.x; o[7] = ca.y; o[8] = ca.z;
    }
    m.area_alias = build_alias_table(areas);
    float total = 0.0f;
    for (float a : areas) total += a;
    m.total_area = total;
    for (auto &v : vs) m.bbox.extend(v);  // bbox over all file positions (triangle_mesh.cpp:168-170)
    return m;
}

}  // namespace btrc
