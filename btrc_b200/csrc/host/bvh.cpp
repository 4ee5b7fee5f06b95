#include "bvh.h"

#include <algorithm>
#include <array>
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
    bool force_leaf = false;  // every range of <= max_leaf primitives becomes a leaf (wide BVH: the triangles of a node's leaf children are tested in one step)
    std::vector<BuildNode> nodes;
    // primitive records are partitioned physically (box, centre, id together): every pass over a range streams memory
    // instead of gathering boxes[order[i]] through a shuffled index array (10 M triangles: 14 s -> a few seconds)
    struct Rec { AABB box; Vec3 c; int id; };
    std::vector<Rec> recs;
    std::atomic<int> next_node{1};

    Builder(const std::vector<AABB> &b, const std::vector<Vec3> &c, int ml) : boxes(b), centers(c), max_leaf(ml) {}

    void make_leaf(int ni, int first, int count, const AABB &box) {
        nodes[ni].box = box;
        nodes[ni].left = nodes[ni].right = -1;
        nodes[ni].first = first;
        nodes[ni].count = count;
    }

    // Bounds and per-axis bins of [first, first + count).  Large ranges (the top of the tree) are binned by all threads with
    // thread-local bins; small ones serially inside their task.
    struct Bins {
        AABB box, cbox;
        AABB bin_box[3][kBins];
        int bin_cnt[3][kBins];
    };
    void bin_range(int first, int count, bool parallel, Bins &out) const {
        auto bounds = [&](int lo, int hi, AABB &box, AABB &cbox) {
            for (int i = lo; i < hi; ++i) { box.extend(recs[i].box); cbox.extend(recs[i].c); }
        };
        out.box = AABB(); out.cbox = AABB();
        if (!parallel) bounds(first, first + count, out.box, out.cbox);
        else {
#pragma omp parallel
            {
                AABB b, c;
#pragma omp for nowait schedule(static)
                for (int i = first; i < first + count; ++i) { b.extend(recs[i].box); c.extend(recs[i].c); }
#pragma omp critical
                { out.box.extend(b); out.cbox.extend(c); }
            }
        }
        for (int a = 0; a < 3; ++a) for (int k = 0; k < kBins; ++k) { out.bin_box[a][k] = AABB(); out.bin_cnt[a][k] = 0; }
        const Vec3 cext = out.cbox.hi - out.cbox.lo;
        float scale[3];
        for (int a = 0; a < 3; ++a) scale[a] = cext[a] > 0 ? kBins / cext[a] : 0.0f;
        auto bin = [&](int lo, int hi, Bins &dst) {
            for (int i = lo; i < hi; ++i) {
                const Rec &r = recs[i];
                for (int a = 0; a < 3; ++a) {
                    if (!(cext[a] > 0)) continue;
                    int b = (int)((r.c[a] - out.cbox.lo[a]) * scale[a]);
                    b = std::min(std::max(b, 0), kBins - 1);
                    dst.bin_box[a][b].extend(r.box);
                    dst.bin_cnt[a][b]++;
                }
            }
        };
        if (!parallel) bin(first, first + count, out);
        else {
#pragma omp parallel
            {
                Bins local;
                for (int a = 0; a < 3; ++a) for (int k = 0; k < kBins; ++k) { local.bin_box[a][k] = AABB(); local.bin_cnt[a][k] = 0; }
#pragma omp for nowait schedule(static)
                for (int i = first; i < first + count; ++i) bin(i, i + 1, local);
#pragma omp critical
                for (int a = 0; a < 3; ++a) for (int k = 0; k < kBins; ++k) { out.bin_box[a][k].extend(local.bin_box[a][k]); out.bin_cnt[a][k] += local.bin_cnt[a][k]; }
            }
        }
    }

    // ranges below this size are built by one task each; above it the node itself is processed by all threads
    static constexpr int kParallelNode = 1 << 18;
    std::vector<std::array<int, 3>> pending;  // (node, first, count) subtrees collected by the parallel top phase

    void build(int ni, int first, int count, bool top = false) {
        if (count <= 1 || (force_leaf && count <= max_leaf)) {
            AABB box;
            for (int i = first; i < first + count; ++i) box.extend(recs[i].box);
            make_leaf(ni, first, count, box);
            return;
        }
        if (top && count < kParallelNode) { pending.push_back({ni, first, count}); return; }
        Bins bins;
        bin_range(first, count, top, bins);
        const AABB &box = bins.box, &cbox = bins.cbox;

        // binned SAH over the three axes
        float best_cost = 3.402823466e+38f;
        int best_axis = -1, best_bin = -1;
        const Vec3 cext = cbox.hi - cbox.lo;
        for (int axis = 0; axis < 3; ++axis) {
            if (!(cext[axis] > 0)) continue;
            const AABB *bin_box = bins.bin_box[axis];
            const int *bin_cnt = bins.bin_cnt[axis];
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
            auto it = std::partition(recs.begin() + first, recs.begin() + first + count, [&](const Rec &r) {
                int b = (int)((r.c[best_axis] - lo) * scale);
                b = std::min(std::max(b, 0), kBins - 1);
                return b <= best_bin;
            });
            mid = (int)(it - recs.begin());
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
        if (top) {
            build(c, first, lc, true);
            build(c + 1, mid, rc, true);
        } else if (count > 8192) {
#pragma omp task shared(nodes, recs)
            build(c, first, lc);
#pragma omp task shared(nodes, recs)
            build(c + 1, mid, rc);
#pragma omp taskwait
        } else {
            build(c, first, lc);
            build(c + 1, mid, rc);
        }
    }
};

}  // namespace

BinaryBvh build_binary_bvh(const std::vector<AABB> &boxes, const std::vector<Vec3> &centers, int max_leaf, bool force_leaf) {
    const int n = (int)boxes.size();
    if (n == 0) throw Error("bvh: no primitives");
    Builder b(boxes, centers, max_leaf);
    b.force_leaf = force_leaf;
    b.nodes.resize(2 * (size_t)n);
    b.recs.resize(n);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) { b.recs[i].box = boxes[i]; b.recs[i].c = centers[i]; b.recs[i].id = i; }
    // top of the tree: every node above kParallelNode primitives is processed by all threads (bounds, binning); the subtrees
    // below are then built by independent tasks
    b.build(0, 0, n, true);
#pragma omp parallel
#pragma omp single
    for (size_t k = 0; k < b.pending.size(); ++k) {
        const std::array<int, 3> t = b.pending[k];
#pragma omp task firstprivate(t) shared(b)
        b.build(t[0], t[1], t[2]);
    }
    BinaryBvh out;
    b.nodes.resize(b.next_node.load());
    out.nodes = std::move(b.nodes);
    out.order.resize(n);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) out.order[i] = b.recs[i].id;
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
#pragma omp parallel for
    for (long long s = 0; s < (long long)n; ++s) {
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
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)dfs.size(); ++k) {
        const int ni = dfs[k];
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

int tree_depth(const BinaryBvh &bvh) {
    if (bvh.nodes.empty()) return 0;
    int best = 0;
    std::vector<std::pair<int, int>> stack = {{0, 1}};
    while (!stack.empty()) {
        auto [ni, d] = stack.back();
        stack.pop_back();
        best = std::max(best, d);
        const BuildNode &nd = bvh.nodes[ni];
        if (nd.left >= 0) { stack.push_back({nd.left, d + 1}); stack.push_back({nd.right, d + 1}); }
    }
    return best;
}

// nodes of a binary tree over arbitrary primitives in the kernels' 64-byte layout; leaf_ref(node) encodes the leaf
template <typename LeafRef>
static std::vector<float> flatten_nodes(const BinaryBvh &bvh, LeafRef leaf_ref, int &root_ref) {
    std::vector<int> inner_index(bvh.nodes.size(), -1), stack = {0}, dfs;
    int n_inner = 0;
    while (!stack.empty()) {
        const int ni = stack.back();
        stack.pop_back();
        const BuildNode &nd = bvh.nodes[ni];
        if (nd.left < 0) continue;
        inner_index[ni] = n_inner++;
        dfs.push_back(ni);
        stack.push_back(nd.right);
        stack.push_back(nd.left);
    }
    std::vector<float> out(16 * (size_t)std::max(n_inner, 1), 0.0f);
    auto ref_of = [&](int ni) { return bvh.nodes[ni].left < 0 ? leaf_ref(bvh.nodes[ni]) : inner_index[ni]; };
    for (size_t k = 0; k < dfs.size(); ++k) {
        const BuildNode &nd = bvh.nodes[dfs[k]];
        const AABB &a = bvh.nodes[nd.left].box, &b = bvh.nodes[nd.right].box;
        float *o = &out[16 * (size_t)inner_index[dfs[k]]];
        o[0] = next_down(a.lo.x); o[1] = next_up(a.hi.x); o[2] = next_down(a.lo.y); o[3] = next_up(a.hi.y);
        o[4] = next_down(b.lo.x); o[5] = next_up(b.hi.x); o[6] = next_down(b.lo.y); o[7] = next_up(b.hi.y);
        o[8] = next_down(a.lo.z); o[9] = next_up(a.hi.z); o[10] = next_down(b.lo.z); o[11] = next_up(b.hi.z);
        const int r0 = ref_of(nd.left), r1 = ref_of(nd.right);
        std::memcpy(&o[12], &r0, 4);
        std::memcpy(&o[13], &r1, 4);
    }
    root_ref = ref_of(0);
    return out;
}

TwoLevelBvh build_two_level_bvh(const std::vector<std::vector<WorldTri>> &mesh_tris, const std::vector<InstanceRef> &insts,
                                const std::vector<WorldTri> &world_tris, float coord_max) {
    TwoLevelBvh out;
    const size_t n_mesh = mesh_tris.size();
    // Object-space boxes must contain the pre-image of the WORLD triangle the kernel tests, seen by a ray that was itself
    // transformed in fp32: both errors are a few ulp of the world coordinates, mapped back by the inverse transform.  Pad every
    // mesh's boxes by 4e-6 * (largest world coordinate) * (largest row norm of any of its instances' world -> object matrices).
    std::vector<float> extra_pad(n_mesh, 0.0f);
    std::vector<char> used(n_mesh, 0);
    for (const InstanceRef &in : insts) {
        float norm = 0;
        for (int r = 0; r < 3; ++r) norm = std::max(norm, std::fabs(in.inv[4 * r]) + std::fabs(in.inv[4 * r + 1]) + std::fabs(in.inv[4 * r + 2]));
        extra_pad[in.mesh] = std::max(extra_pad[in.mesh], 4e-6f * coord_max * norm);
        used[in.mesh] = 1;
    }
    std::vector<BinaryBvh> trees(n_mesh);
    std::vector<int> mesh_root(n_mesh, 0), mesh_depth(n_mesh, 0);
    int deepest = 0;
    for (size_t m = 0; m < n_mesh; ++m) {
        if (!used[m] || mesh_tris[m].empty()) continue;
        std::vector<AABB> boxes;
        std::vector<Vec3> centers;
        float obj_max = 0;
        for (const WorldTri &t : mesh_tris[m])
            for (const Vec3 &p : {t.p0, t.p1, t.p2})
                obj_max = std::max(obj_max, std::max(std::fabs(p.x), std::max(std::fabs(p.y), std::fabs(p.z))));
        padded_triangle_boxes(mesh_tris[m], obj_max, boxes, centers);
        for (AABB &b : boxes) { b.lo = b.lo - Vec3(extra_pad[m]); b.hi = b.hi + Vec3(extra_pad[m]); }
        trees[m] = build_binary_bvh(boxes, centers, 4);
        mesh_depth[m] = tree_depth(trees[m]);
        deepest = std::max(deepest, mesh_depth[m]);
        int root = 0;
        std::vector<float> nodes = flatten_nodes(trees[m], [](const BuildNode &nd) {
            if (nd.count < 1 || nd.count > 8) throw Error("bvh: leaf size out of range");
            return ~((nd.first << 3) | (nd.count - 1));
        }, root);
        const int base = (int)(out.bl_nodes.size() / 16);
        for (size_t k = 0; k + 15 < nodes.size(); k += 16)  // inner references become indices into the concatenated array
            for (int c = 12; c <= 13; ++c) {
                int r;
                std::memcpy(&r, &nodes[k + c], 4);
                if (r >= 0) { r += base; std::memcpy(&nodes[k + c], &r, 4); }
            }
        mesh_root[m] = root >= 0 ? root + base : root;
        out.bl_nodes.insert(out.bl_nodes.end(), nodes.begin(), nodes.end());
    }
    // world-space triangle records per instance, in its mesh tree's leaf order
    size_t total = 0;
    for (const InstanceRef &in : insts) total += mesh_tris[in.mesh].size();
    if (total >= (1ull << 31)) throw Error("bvh: more than 2^31 - 1 instanced triangles");
    out.tris.resize(12 * total);
    out.tri_ids.resize(2 * total);
    out.inst.resize(insts.size());
    size_t base = 0;
    for (size_t i = 0; i < insts.size(); ++i) {
        const InstanceRef &in = insts[i];
        InstanceBlas &ib = out.inst[i];
        std::memcpy(ib.w2o, in.inv, sizeof(ib.w2o));
        ib.root_ref = mesh_root[in.mesh];
        ib.tri_base = (int)base;
        ib.pad[0] = ib.pad[1] = 0;
        const std::vector<int> &order = trees[in.mesh].order;
        const long long n = (long long)order.size();
#pragma omp parallel for if (n > 4096)
        for (long long s = 0; s < n; ++s) {
            const WorldTri &t = world_tris[in.first_tri + order[s]];
            float *o = &out.tris[12 * (base + s)];
            o[0] = t.p0.x; o[1] = t.p0.y; o[2] = t.p0.z; o[3] = t.n.x;
            o[4] = t.e1.x; o[5] = t.e1.y; o[6] = t.e1.z; o[7] = t.n.y;
            o[8] = t.e2.x; o[9] = t.e2.y; o[10] = t.e2.z; o[11] = t.n.z;
            out.tri_ids[2 * (base + s)] = t.inst;
            out.tri_ids[2 * (base + s) + 1] = t.prim;
        }
        base += (size_t)n;
    }
    // instance tree over the world boxes of the instances' (padded) world triangles
    std::vector<AABB> tri_boxes;
    std::vector<Vec3> tri_centers;
    padded_triangle_boxes(world_tris, coord_max, tri_boxes, tri_centers);
    std::vector<AABB> ibox(insts.size());
    std::vector<Vec3> icen(insts.size());
    std::vector<int> live;  // instances with geometry
    for (size_t i = 0; i < insts.size(); ++i) {
        AABB b;
        const size_t n = mesh_tris[insts[i].mesh].size();
        for (size_t k = 0; k < n; ++k) { b.extend(tri_boxes[insts[i].first_tri + k].lo); b.extend(tri_boxes[insts[i].first_tri + k].hi); }
        ibox[i] = b;
        icen[i] = 0.5f * (b.lo + b.hi);
        if (n) live.push_back((int)i);
    }
    if (live.empty()) throw Error("bvh: no geometry");
    std::vector<AABB> lb;
    std::vector<Vec3> lc;
    for (int i : live) { lb.push_back(ibox[i]); lc.push_back(icen[i]); }
    BinaryBvh top = build_binary_bvh(lb, lc, 1);
    out.tl_nodes = flatten_nodes(top, [&](const BuildNode &nd) {
        if (nd.count != 1) throw Error("bvh: instance leaves hold one instance");
        return ~(live[top.order[nd.first]] << 3);
    }, out.tl_root_ref);
    out.depth = tree_depth(top) + deepest + 1;
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// 8-wide collapse + quantisation
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct WideTask {
    int binary;  // binary node this wide node is collapsed from
    int index;   // wide node index
    int depth;
};

// SAH-optimal collapse of the binary hierarchy into 8-wide nodes (dynamic programme of Ylitie et al. 2017, section 4.1,
// restricted to the binary leaves as the only leaves): cost[n][i-1] = cheapest way to represent the subtree of n by at
// most i wide-node children; i = 1 means "n becomes a wide node of its own" (or is a leaf).
struct Collapse {
    static constexpr float kNodeCost = 1.0f, kPrimCost = 0.3f;
    const BinaryBvh &bvh;
    std::vector<float> cost;      // 7 per binary node
    std::vector<uint8_t> split;   // 8 per binary node: [i-1] for i = 2..7: children given to the left subtree (0 = same as i-1);
                                  // [7]: left share when the node is opened as a wide node (8 slots)
    explicit Collapse(const BinaryBvh &b) : bvh(b), cost(7 * b.nodes.size()), split(8 * b.nodes.size(), 0) {
        // children are always created after their parent (Builder::build), so a reverse sweep sees children first
        for (long long ni = (long long)b.nodes.size() - 1; ni >= 0; --ni) {
            const BuildNode &nd = b.nodes[ni];
            float *c = &cost[7 * ni];
            const float area = nd.box.half_area();
            if (nd.left < 0) {
                for (int i = 0; i < 7; ++i) c[i] = area * nd.count * kPrimCost;
                continue;
            }
            const float *cl = &cost[7 * (size_t)nd.left], *cr = &cost[7 * (size_t)nd.right];
            auto distribute = [&](int j, int &best_k) {
                float best = 3.402823466e+38f;
                for (int k = 1; k < j; ++k) {
                    if (k > 7 || j - k > 7) continue;
                    const float v = cl[k - 1] + cr[j - k - 1];
                    if (v < best) { best = v; best_k = k; }
                }
                return best;
            };
            int k = 1;
            c[0] = distribute(8, k) + area * kNodeCost;
            split[8 * ni + 7] = (uint8_t)k;
            for (int i = 2; i <= 7; ++i) {
                const float d = distribute(i, k);
                if (d < c[i - 2]) { c[i - 1] = d; split[8 * ni + i - 1] = (uint8_t)k; }
                else { c[i - 1] = c[i - 2]; split[8 * ni + i - 1] = 0; }
            }
        }
    }
    void expand(int ni, int j, int *out, int &n) const {
        const BuildNode &nd = bvh.nodes[ni];
        if (nd.left < 0 || j == 1) { out[n++] = ni; return; }
        const int k = split[8 * (size_t)ni + j - 1];
        if (k == 0) { expand(ni, j - 1, out, n); return; }
        expand(nd.left, k, out, n);
        expand(nd.right, j - k, out, n);
    }
    // children of the wide node made from binary node ni
    int children(int ni, int *out) const {
        const BuildNode &root = bvh.nodes[ni];
        int n = 0;
        if (root.left < 0) { out[n++] = ni; return n; }  // a scene of <= 3 triangles: the root is a leaf
        const int k = split[8 * (size_t)ni + 7];
        expand(root.left, k, out, n);
        expand(root.right, 8 - k, out, n);
        return n;
    }
};

// Greedy octant assignment: slot bit a set <=> the child lies on the + side of the node centre along axis a.
void assign_slots(const BinaryBvh &bvh, const AABB &box, const int *children, int n, int *slot_child) {
    const Vec3 centre = 0.5f * (box.lo + box.hi);
    float cost[8][8];
    for (int c = 0; c < n; ++c) {
        const AABB &b = bvh.nodes[children[c]].box;
        const Vec3 d = 0.5f * (b.lo + b.hi) - centre;
        for (int s = 0; s < 8; ++s) cost[c][s] = (s & 1 ? d.x : -d.x) + (s & 2 ? d.y : -d.y) + (s & 4 ? d.z : -d.z);
    }
    bool child_done[8] = {false}, slot_done[8] = {false};
    for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    for (int k = 0; k < n; ++k) {
        int bc = -1, bs = -1;
        float best = -3.402823466e+38f;
        for (int c = 0; c < n; ++c) {
            if (child_done[c]) continue;
            for (int s = 0; s < 8; ++s)
                if (!slot_done[s] && cost[c][s] > best) { best = cost[c][s]; bc = c; bs = s; }
        }
        child_done[bc] = true;
        slot_done[bs] = true;
        slot_child[bs] = children[bc];
    }
}

}  // namespace

WideBvh8 build_wide_bvh8(const BinaryBvh &bvh, const std::vector<WorldTri> &tris) {
    WideBvh8 out;
    const size_t n_tris = bvh.order.size();
    out.tris.resize(12 * n_tris);
    out.tri_ids.resize(2 * n_tris);
    out.nodes.assign(20, 0u);
    size_t tri_cursor = 0;
    const Collapse collapse(bvh);
    std::vector<WideTask> queue = {{0, 0, 1}};
    for (size_t qi = 0; qi < queue.size(); ++qi) {  // breadth first: siblings are contiguous, the top levels are packed together
        const WideTask task = queue[qi];
        out.depth = std::max(out.depth, task.depth);
        int children[8], slot_child[8];
        const int n = collapse.children(task.binary, children);
        AABB box;
        for (int c = 0; c < n; ++c) box.extend(bvh.nodes[children[c]].box);
        assign_slots(bvh, box, children, n, slot_child);

        // quantisation frame: plane = p + q * 2^e with q in [0, 255].  The kernel evaluates t = (2^23 + q) * s + (b - 2^23 * s)
        // (one PRMT + one FFMA per plane), whose rounding can move a plane by half a quantum: every child box is widened
        // by one quantum on each side, and the frame leaves room for that (children occupy quanta 1..254).
        int e[3];
        float p[3];
        for (int a = 0; a < 3; ++a) {
            const double lo = box.lo[a], hi = box.hi[a];
            int ea = hi > lo ? (int)std::ceil(std::log2((hi - lo) / 253.0)) : -100;
            ea = std::min(std::max(ea, -100), 100);
            for (;;) {
                const double sc = std::ldexp(1.0, ea);
                float pf = (float)(lo - sc);
                if ((double)pf > lo - sc) pf = std::nextafter(pf, -3.402823466e+38f);
                if (std::floor((lo - (double)pf) / sc) >= 1.0 && std::ceil((hi - (double)pf) / sc) <= 254.0) { p[a] = pf; break; }
                ++ea;
            }
            e[a] = ea;
        }
        uint32_t imask = 0;
        uint8_t meta[8] = {0}, qlo[3][8], qhi[3][8];
        const uint32_t child_base = (uint32_t)(out.nodes.size() / 20);
        const uint32_t tri_base = (uint32_t)tri_cursor;
        int n_inner = 0;
        for (int s = 0; s < 8; ++s) {
            for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }  // empty slots can never be hit
            const int ci = slot_child[s];
            if (ci < 0) continue;
            const BuildNode &c = bvh.nodes[ci];
            for (int a = 0; a < 3; ++a) {
                const double sc = std::ldexp(1.0, e[a]);
                const double lo = std::floor(((double)c.box.lo[a] - (double)p[a]) / sc) - 1.0;
                const double hi = std::ceil(((double)c.box.hi[a] - (double)p[a]) / sc) + 1.0;
                if (lo < 0.0 || hi > 255.0) throw Error("bvh8: child box outside the quantisation frame");
                qlo[a][s] = (uint8_t)lo;
                qhi[a][s] = (uint8_t)hi;
            }
            if (c.left >= 0) {
                imask |= 1u << s;
                meta[s] = (uint8_t)(0x20 | (24 + s));
                ++n_inner;
            } else {
                if (c.count < 1 || c.count > 3) throw Error("bvh8: leaf size out of range");
                const size_t off = tri_cursor - tri_base;
                if (off + c.count > 24) throw Error("bvh8: more than 24 triangles under one node");
                meta[s] = (uint8_t)((((1u << c.count) - 1u) << 5) | (uint32_t)off);
                for (int k = 0; k < c.count; ++k, ++tri_cursor) {
                    const WorldTri &t = tris[bvh.order[c.first + k]];
                    float *o = &out.tris[12 * tri_cursor];
                    o[0] = t.p0.x; o[1] = t.p0.y; o[2] = t.p0.z; o[3] = t.n.x;
                    o[4] = t.e1.x; o[5] = t.e1.y; o[6] = t.e1.z; o[7] = t.n.y;
                    o[8] = t.e2.x; o[9] = t.e2.y; o[10] = t.e2.z; o[11] = t.n.z;
                    out.tri_ids[2 * tri_cursor] = t.inst;
                    out.tri_ids[2 * tri_cursor + 1] = t.prim;
                }
            }
        }
        out.nodes.resize(out.nodes.size() + 20 * (size_t)n_inner, 0u);
        int k = 0;
        for (int s = 0; s < 8; ++s)
            if (imask & (1u << s)) queue.push_back({slot_child[s], (int)child_base + k++, task.depth + 1});

        uint32_t *w = &out.nodes[20 * (size_t)task.index];
        std::memcpy(w, p, 12);
        w[3] = (uint32_t)(e[0] + 127) | (uint32_t)(e[1] + 127) << 8 | (uint32_t)(e[2] + 127) << 16 | imask << 24;
        w[4] = child_base;
        w[5] = tri_base;
        auto pack = [](const uint8_t *b) { return (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24; };
        w[6] = pack(meta); w[7] = pack(meta + 4);
        w[8] = pack(qlo[0]); w[9] = pack(qlo[0] + 4); w[10] = pack(qlo[1]); w[11] = pack(qlo[1] + 4);
        w[12] = pack(qlo[2]); w[13] = pack(qlo[2] + 4); w[14] = pack(qhi[0]); w[15] = pack(qhi[0] + 4);
        w[16] = pack(qhi[1]); w[17] = pack(qhi[1] + 4); w[18] = pack(qhi[2]); w[19] = pack(qhi[2] + 4);
    }
    if (tri_cursor != n_tris) throw Error("bvh8: triangle count mismatch");
    return out;
}

}  // namespace btrc
