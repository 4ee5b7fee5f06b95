// Experiment (not part of the product build): warp-level cost model of BVH2 traversal loop structures on real ray batches.
// 32 rays per warp run in lock step; a warp-wide node step costs C_NODE issue slots, a triangle test C_TRI, whether 1 or 32
// lanes take part.  Reports issue slots per ray and SIMT efficiency for:
//   ww    while-while with a warp vote (the product kernel: all lanes descend to their next leaf, then all leaves are tested)
//   spec  speculative while-while: a lane that reaches a leaf postpones it and keeps descending until it holds two leaves
//   ifif  if-if: every iteration each lane does one node step OR one triangle test
// Build: g++ -O2 -std=c++17 -fopenmp -I/usr/local/cuda/include simt_sim.cpp ../../btrc_b200/csrc/host/bvh.cpp -shared -fPIC -o _simt.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../btrc_b200/csrc/host/bvh.h"

using namespace btrc;

namespace {

struct Sim {
    FlatBvh2 b;
};

constexpr int SENT = 0x7fffffff;

struct Lane {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, oox, ooy, ooz, best;
    int cur, sp, hit;
    int stack[64];
    int leaf_first, leaf_left;   // triangles still to test in the current leaf
    int post_first, post_left;   // postponed leaf (speculative)
    bool any_hit, done;
};

void init(Lane &l, const float *o, const float *d, int root, bool any_hit) {
    l.ox = o[0]; l.oy = o[1]; l.oz = o[2]; l.dx = d[0]; l.dy = d[1]; l.dz = d[2]; l.best = d[3];
    auto inv = [](float v) { const float t = 1e-20f; return 1.0f / (std::fabs(v) > t ? v : std::copysign(t, v)); };
    l.idx = inv(l.dx); l.idy = inv(l.dy); l.idz = inv(l.dz);
    l.oox = l.ox * l.idx; l.ooy = l.oy * l.idy; l.ooz = l.oz * l.idz;
    l.sp = 0; l.stack[l.sp++] = SENT; l.cur = root; l.hit = -1;
    l.leaf_left = l.post_left = 0; l.any_hit = any_hit; l.done = false;
}

inline bool in_node(const Lane &l) { return !l.done && l.cur >= 0 && l.cur != SENT; }

void node_step(const Sim &s, Lane &l) {
    const float *n = &s.b.nodes[16 * (size_t)l.cur];
    auto slab = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, float &tmin, float &tmax) {
        const float ax = std::fma(lox, l.idx, -l.oox), bx = std::fma(hix, l.idx, -l.oox);
        const float ay = std::fma(loy, l.idy, -l.ooy), by = std::fma(hiy, l.idy, -l.ooy);
        const float az = std::fma(loz, l.idz, -l.ooz), bz = std::fma(hiz, l.idz, -l.ooz);
        tmin = std::fmax(std::fmax(std::fmin(ax, bx), std::fmin(ay, by)), std::fmax(std::fmin(az, bz), 0.0f));
        tmax = std::fmin(std::fmin(std::fmax(ax, bx), std::fmax(ay, by)), std::fmin(std::fmax(az, bz), l.best));
    };
    float a0, a1, b0, b1;
    slab(n[0], n[1], n[2], n[3], n[8], n[9], a0, a1);
    slab(n[4], n[5], n[6], n[7], n[10], n[11], b0, b1);
    int r0, r1;
    std::memcpy(&r0, &n[12], 4); std::memcpy(&r1, &n[13], 4);
    const bool h0 = a0 <= a1, h1 = b0 <= b1;
    if (!h0 && !h1) l.cur = l.stack[--l.sp];
    else if (h0 && h1) { const bool sw = b0 < a0; l.cur = sw ? r1 : r0; l.stack[l.sp++] = sw ? r0 : r1; }
    else l.cur = h0 ? r0 : r1;
}

// one triangle of the lane's current leaf (plain Moeller-Trumbore: hit identity does not matter for the cost model)
void tri_step(const Sim &s, Lane &l, int &first, int &left) {
    const float *t = &s.b.tris[12 * (size_t)first];
    ++first; --left;
    const float cx = t[0] - l.ox, cy = t[1] - l.oy, cz = t[2] - l.oz;
    const float rx = l.dy * cz - l.dz * cy, ry = l.dz * cx - l.dx * cz, rz = l.dx * cy - l.dy * cx;
    const float det = t[3] * l.dx + t[7] * l.dy + t[11] * l.dz, inv = 1.0f / det;
    const float u = (rx * t[8] + ry * t[9] + rz * t[10]) * inv, v = (rx * t[4] + ry * t[5] + rz * t[6]) * inv;
    if (u >= 0 && v >= 0 && 1.0f - u - v >= 0) {
        const float tt = (t[3] * cx + t[7] * cy + t[11] * cz) * inv;
        if (tt >= 0 && tt <= l.best) {
            l.best = tt; l.hit = first - 1;
            if (l.any_hit) { l.done = true; left = 0; l.leaf_left = l.post_left = 0; l.cur = SENT; }
        }
    }
}

inline void enter_leaf(Lane &l) {  // cur < 0: set up the triangle range and pop the next node
    const int ref = ~l.cur;
    l.leaf_first = ref >> 3; l.leaf_left = (ref & 7) + 1;
}

}  // namespace

extern "C" {

void *sim_build(int n, const float *verts) {
    std::vector<WorldTri> tris(n);
    AABB wb;
    for (int i = 0; i < n; ++i) {
        WorldTri &t = tris[i];
        t.p0 = Vec3(verts[9 * i], verts[9 * i + 1], verts[9 * i + 2]);
        t.p1 = Vec3(verts[9 * i + 3], verts[9 * i + 4], verts[9 * i + 5]);
        t.p2 = Vec3(verts[9 * i + 6], verts[9 * i + 7], verts[9 * i + 8]);
        t.e1 = t.p0 - t.p1; t.e2 = t.p2 - t.p0; t.n = cross(t.e1, t.e2); t.inst = 0; t.prim = (uint32_t)i;
        wb.extend(t.p0); wb.extend(t.p1); wb.extend(t.p2);
    }
    float cm = 0;
    for (int a = 0; a < 3; ++a) cm = std::max(cm, std::max(std::fabs(wb.lo[a]), std::fabs(wb.hi[a])));
    std::vector<AABB> boxes; std::vector<Vec3> centers;
    padded_triangle_boxes(tris, cm, boxes, centers);
    auto *s = new Sim;
    s->b = flatten_bvh2(build_binary_bvh(boxes, centers, 4), tris);
    return s;
}
void sim_free(void *s) { delete static_cast<Sim *>(s); }

// mode 0 = ww, 1 = spec, 2 = ifif.  out[0] = warp node steps, out[1] = warp triangle steps, out[2] = lane node steps, out[3] = lane triangle steps
void sim_run(void *sp, int n, const float *o_med, const float *d_t1, int any_hit, int mode, double *out) {
    const Sim &s = *static_cast<Sim *>(sp);
    double wn = 0, wt = 0, ln = 0, lt = 0;
    const bool refill = mode >= 3;       // 3: ww + refill (threshold in mode - 3 ... see below)
    const int thresh = refill ? (mode == 3 ? 32 : (mode == 4 ? 24 : 16)) : 0;   // refill when fewer than `thresh` lanes are busy
    if (refill) mode = 0;
    const int per_warp = refill ? 32 * 64 : 32;  // a persistent warp works through a chunk of rays
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : wn, wt, ln, lt)
    for (int w = 0; w < (n + per_warp - 1) / per_warp; ++w) {
        Lane L[32];
        int next = per_warp * w;
        const int end = std::min(n, per_warp * (w + 1));
        for (int i = 0; i < 32; ++i) {
            if (next < end) { init(L[i], o_med + 4 * (size_t)next, d_t1 + 4 * (size_t)next, s.b.root_ref, any_hit != 0); ++next; }
            else { L[i].done = true; L[i].cur = SENT; L[i].leaf_left = L[i].post_left = 0; }
        }
        auto alive = [&](const Lane &l) { return !l.done && (l.cur != SENT || l.leaf_left > 0 || l.post_left > 0); };
        for (;;) {
            bool any = false;
            for (int i = 0; i < 32; ++i) any |= alive(L[i]);
            if (!any && next >= end) break;
            if (refill && next < end) {
                int busy = 0;
                for (int i = 0; i < 32; ++i) busy += alive(L[i]) ? 1 : 0;
                if (busy < thresh) {
                    for (int i = 0; i < 32 && next < end; ++i)
                        if (!alive(L[i])) { init(L[i], o_med + 4 * (size_t)next, d_t1 + 4 * (size_t)next, s.b.root_ref, any_hit != 0); ++next; }
                    wt += 0.5;  // the refill itself: ballot, cursor, two 16-byte loads, set-up (counted as half a triangle step)
                }
            }
            if (mode == 0 || mode == 1) {
                // ---- descend phase
                for (;;) {
                    int active = 0;
                    for (int i = 0; i < 32; ++i) {
                        Lane &l = L[i];
                        if (l.done) continue;
                        if (mode == 1 && l.cur < 0 && l.cur != SENT && l.post_left == 0 && l.leaf_left == 0) {
                            // speculative: postpone this leaf and continue with the next node from the stack
                            const int ref = ~l.cur;
                            l.post_first = ref >> 3; l.post_left = (ref & 7) + 1;
                            l.cur = l.stack[--l.sp];
                        }
                        if (in_node(l)) { node_step(s, l); ++active; }
                    }
                    if (!active) break;
                    wn += 1; ln += active;
                }
                // ---- leaf phase: current leaf (+ postponed leaf)
                for (int i = 0; i < 32; ++i) if (!L[i].done && L[i].cur < 0 && L[i].cur != SENT && L[i].leaf_left == 0) { enter_leaf(L[i]); L[i].cur = L[i].stack[--L[i].sp]; }
                for (;;) {
                    int active = 0;
                    for (int i = 0; i < 32; ++i) {
                        Lane &l = L[i];
                        if (l.done) continue;
                        if (l.post_left > 0) { tri_step(s, l, l.post_first, l.post_left); ++active; }
                        else if (l.leaf_left > 0) { tri_step(s, l, l.leaf_first, l.leaf_left); ++active; }
                    }
                    if (!active) break;
                    wt += 1; lt += active;
                }
            } else {
                int an = 0, at = 0;
                for (int i = 0; i < 32; ++i) {
                    Lane &l = L[i];
                    if (l.done) continue;
                    if (l.leaf_left > 0) { tri_step(s, l, l.leaf_first, l.leaf_left); ++at; }
                    else if (in_node(l)) { node_step(s, l); ++an; }
                    if (!l.done && l.leaf_left == 0 && l.cur < 0 && l.cur != SENT) { enter_leaf(l); l.cur = l.stack[--l.sp]; }
                }
                if (an) { wn += 1; ln += an; }
                if (at) { wt += 1; lt += at; }
            }
            for (int i = 0; i < 32; ++i) if (!L[i].done && L[i].cur == SENT && L[i].leaf_left == 0 && L[i].post_left == 0) L[i].done = true;
        }
    }
    out[0] = wn; out[1] = wt; out[2] = ln; out[3] = lt;
}

}  // extern "C"
