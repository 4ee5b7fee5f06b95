#include "bvh.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <numeric>

#include "node.h"

namespace btrc {

namespace {

constexpr int kBins = 16;

struct Builder {
    const std::vector<AABB> &boxes;
    const std::vector<Vec3> &centers;
    int max_leaf;
    std::vector<BuildNode> nodes;
    std::vector<int> order;
    std::atomic<int> next_node{1};

    Builder(const std::vector<AABB> &b, const std::vector<Vec3> &c, int ml) : boxes(b), centers(c), max_leaf(ml) {}

    void make_leaf(int ni, int first, int count, const AABB &box) {
        nodes[ni].box = box;
        nodes[ni].left = nodes[ni].right = -1;
        nodes[ni].first = first;
        nodes[ni].count = count;
    }

    void build(int ni, int first, int count) {
        AABB box, cbox;
        for (int i = first; i < first + count; ++i) {
            box.extend(boxes[order[i]]);
            cbox.extend(centers[order[i]]);
        }
        if (count <= 1) { make_leaf(ni, first, count, box); return; }

        // binned SAH over the three axes
        float best_cost = 3.402823466e+38f;
        int best_axis = -1, best_bin = -1;
        const Vec3 cext = cbox.hi - cbox.lo;
        for (int axis = 0; axis < 3; ++axis) {
            if (!(cext[axis] > 0)) continue;
            AABB bin_box[kBins];
            int bin_cnt[kBins] = {0};
            const float scale = kBins / cext[axis];
            for (int i = first; i < first + count; ++i) {
                const int p = order[i];
                int b = (int)((centers[p][axis] - cbox.lo[axis]) * scale);
                b = std::min(std::max(b, 0), kBins - 1);
                bin_box[b].extend(boxes[p]);
                bin_cnt[b]++;
            }
            float right_area[kBins];
            int right_cnt[kBins];
            AABB acc;
            int cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                acc.extend(bin_box[b]);
                cnt += bin_cnt[b];
                right_area[b] = acc.empty() ? 0.0f : acc.half_area();
                right_cnt[b] = cnt;
            }
            acc = AABB();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.extend(bin_box[b]);
                cnt += bin_cnt[b];
                if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                const float cost = acc.half_area() * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }

        int mid = -1;
        if (best_axis >= 0) {
            const float leaf_cost = box.half_area() * count;
            const float split_cost = box.half_area() * 1.0f + best_cost;  // traversal cost 1, intersection cost 1
            if (count <= max_leaf && leaf_cost <= split_cost) { make_leaf(ni, first, count, box); return; }
            const float scale = kBins / cext[best_axis];
            const float lo = cbox.lo[best_axis];
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int p) {
                int b = (int)((centers[p][best_axis] - lo) * scale);
                b = std::min(std::max(b, 0), kBins - 1);
                return b <= best_bin;
            });
            mid = (int)(it - order.begin());
        }
        if (mid <= first || mid >= first + count) {
            if (count <= max_leaf) { make_leaf(ni, first, count, box); return; }
            // degenerate (coincident centroids): median split by index
            mid = first + count / 2;
        }
        const int c = next_node.fetch_add(2);
        nodes[ni].box = box;
        nodes[ni].left = c;
        nodes[ni].right = c + 1;
        nodes[ni].first = 0;
        nodes[ni].count = 0;
        const int lc = mid - first, rc = first + count - mid;
        if (count > 8192) {
#pragma omp task shared(nodes, order)
            build(c, first, lc);
#pragma omp task shared(nodes, order)
            build(c + 1, mid, rc);
#pragma omp taskwait
        } else {
            build(c, first, lc);
            build(c + 1, mid, rc);
        }
    }
};

}  // namespace

BinaryBvh build_binary_bvh(const std::vector<AABB> &boxes, const std::vector<Vec3> &centers, int max_leaf) {
    const int n = (int)boxes.size();
    if (n == 0) throw Error("bvh: no primitives");
    Builder b(boxes, centers, max_leaf);
    b.nodes.resize(2 * (size_t)n);
    b.order.resize(n);
    std::iota(b.order.begin(), b.order.end(), 0);
#pragma omp parallel
#pragma omp single
    b.build(0, 0, n);
    BinaryBvh out;
    b.nodes.resize(b.next_node.load());
    out.nodes = std::move(b.nodes);
    out.order = std::move(b.order);
    return out;
}

void padded_triangle_boxes(const std::vector<WorldTri> &tris, float coord_max, std::vector<AABB> &boxes, std::vector<Vec3> &centers) {
    const size_t n = tris.size();
    boxes.resize(n);
    centers.resize(n);
#pragma omp parallel for
    for (long long i = 0; i < (long long)n; ++i) {
        const WorldTri &t = tris[i];
        AABB b;
        b.extend(t.p0); b.extend(t.p1); b.extend(t.p2);
        b.extend(t.p0 - t.e1); b.extend(t.p0 + t.e2);
        const Vec3 ext = b.hi - b.lo;
        const float pad = 1e-5f * std::max(ext.x, std::max(ext.y, ext.z)) + 4e-7f * coord_max;
        b.lo = b.lo - Vec3(pad); b.hi = b.hi + Vec3(pad);
        boxes[i] = b;
        centers[i] = 0.5f * (b.lo + b.hi);
    }
}

// Conservative widening of a box so that the fp32 slab test in the kernel can never cull a triangle that the
// exact (lib/bvh-arithmetic) triangle test accepts.
static inline float next_down(float x) { return std::nextafter(x, -3.402823466e+38f); }
static inline float next_up(float x) { return std::nextafter(x, 3.402823466e+38f); }

FlatBvh2 flatten_bvh2(const BinaryBvh &bvh, const std::vector<WorldTri> &tris) {
    FlatBvh2 out;
    const size_t n = bvh.order.size();
    out.tris.resize(12 * n);
    out.tri_ids.resize(2 * n);
    for (size_t s = 0; s < n; ++s) {
        const WorldTri &t = tris[bvh.order[s]];
        float *o = &out.tris[12 * s];
        o[0] = t.p0.x; o[1] = t.p0.y; o[2] = t.p0.z; o[3] = t.n.x;
        o[4] = t.e1.x; o[5] = t.e1.y; o[6] = t.e1.z; o[7] = t.n.y;
        o[8] = t.e2.x; o[9] = t.e2.y; o[10] = t.e2.z; o[11] = t.n.z;
        out.tri_ids[2 * s] = t.inst;
        out.tri_ids[2 * s + 1] = t.prim;
    }
    auto leaf_ref = [](const BuildNode &nd) {
        if (nd.count < 1 || nd.count > 8) throw Error("bvh: leaf size out of range");
        return ~((nd.first << 3) | (nd.count - 1));
    };
    // inner nodes are renumbered densely in DFS order
    std::vector<int> inner_index(bvh.nodes.size(), -1);
    std::vector<int> stack = {0};
    int n_inner = 0;
    std::vector<int> dfs;
    while (!stack.empty()) {
        int ni = stack.back();
        stack.pop_back();
        const BuildNode &nd = bvh.nodes[ni];
        if (nd.left < 0) continue;
        inner_index[ni] = n_inner++;
        dfs.push_back(ni);
        stack.push_back(nd.right);
        stack.push_back(nd.left);
    }
    out.nodes.assign(16 * (size_t)std::max(n_inner, 1), 0.0f);
    auto ref_of = [&](int ni) { return bvh.nodes[ni].left < 0 ? leaf_ref(bvh.nodes[ni]) : inner_index[ni]; };
    for (int ni : dfs) {
        const BuildNode &nd = bvh.nodes[ni];
        const AABB &a = bvh.nodes[nd.left].box, &b = bvh.nodes[nd.right].box;
        float *o = &out.nodes[16 * (size_t)inner_index[ni]];
        o[0] = next_down(a.lo.x); o[1] = next_up(a.hi.x); o[2] = next_down(a.lo.y); o[3] = next_up(a.hi.y);
        o[4] = next_down(b.lo.x); o[5] = next_up(b.hi.x); o[6] = next_down(b.lo.y); o[7] = next_up(b.hi.y);
        o[8] = next_down(a.lo.z); o[9] = next_up(a.hi.z); o[10] = next_down(b.lo.z); o[11] = next_up(b.hi.z);
        int r0 = ref_of(nd.left), r1 = ref_of(nd.right);
        std::memcpy(&o[12], &r0, 4);
        std::memcpy(&o[13], &r1, 4);
    }
    out.root_ref = ref_of(0);
    return out;
}

}  // namespace btrc
