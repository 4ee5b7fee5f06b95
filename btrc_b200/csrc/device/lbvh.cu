// Device-side LBVH build: kernels around the per-element functions of lbvh.cuh, CUB radix sort / scan in between.
#include "lbvh.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdexcept>
#include <string>

#include "kernels.h"

namespace btrc {

namespace {

#define LBVH_CUDA(x)                                                                                     \
    do {                                                                                                 \
        cudaError_t e_ = (x);                                                                            \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("lbvh: ") + cudaGetErrorString(e_)); \
    } while (0)

template <void (*F)(const LbvhBuild &, int)>
__global__ void lbvh_for_each(LbvhBuild b, int count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) F(b, i);
}

__device__ void k_box_and_code(const LbvhBuild &b, int i) { lbvh_box_and_code(b, i); }
__device__ void k_gather(const LbvhBuild &b, int i) { lbvh_gather(b, i); }
__device__ void k_build_node(const LbvhBuild &b, int i) { lbvh_build_node(b, i); }
__device__ void k_fit(const LbvhBuild &b, int i) { lbvh_fit_from_leaf(b, i); }
__device__ void k_emit(const LbvhBuild &b, int i) { lbvh_emit_node(b, i); }

struct Scratch {
    void *p = nullptr;
    explicit Scratch(size_t bytes) { if (bytes) LBVH_CUDA(cudaMalloc(&p, bytes)); }
    ~Scratch() { if (p) cudaFree(p); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

// tris_in / ids_in: device arrays in input order.  On return *nodes_out (cudaMalloc'ed, 64 bytes per node), tris_out / ids_out (caller
// allocated, sorted order) describe the BVH; the function returns the root reference.
int build_lbvh_device(int n, const float4 *tris_in, const uint2 *ids_in, const float scene_lo[3], const float scene_hi[3], float coord_max,
                      float4 *tris_out, uint2 *ids_out, float4 **nodes_out, int *n_nodes_out, cudaStream_t stream) {
    *nodes_out = nullptr;
    *n_nodes_out = 0;
    if (n <= 0) throw std::runtime_error("lbvh: no primitives");
    // Leaf references pack `first << 3` into an int, and the traversal kernels walk with a 64-entry stack (traverse.cuh kStackSize).
    // A radix tree over 30-bit codes with the input index as tie-break is at most 30 + ceil(log2 n) levels deep: below 2^28
    // primitives that is 58 levels plus the sentinel, inside the stack; beyond it both bounds fail, so refuse (the SAH builder's
    // counterpart is the `bvh_depth > 60` check in host/scene.cpp).
    if (n >= (1 << 28)) throw std::runtime_error("lbvh: more than 2^28 - 1 primitives are not supported (leaf reference width, traversal stack depth)");
    const int ni = n - 1;  // internal nodes of the radix tree
    Scratch code(4 * (size_t)n), code2(4 * (size_t)n), index(4 * (size_t)n), index2(4 * (size_t)n), box_in(sizeof(LbvhBox) * (size_t)n);
    Scratch child(sizeof(int2) * (size_t)std::max(ni, 1)), range(sizeof(int2) * (size_t)std::max(ni, 1)), parent(4 * (size_t)(2 * n)),
        node_box(sizeof(LbvhBox) * (size_t)std::max(ni, 1)), visit(4 * (size_t)std::max(ni, 1)), live(4 * (size_t)std::max(ni, 1)),
        new_index(4 * (size_t)std::max(ni, 1));
    LbvhBuild b{};
    b.n = n; b.tris_in = tris_in; b.ids_in = ids_in; b.coord_max = coord_max;
    for (int k = 0; k < 3; ++k) {
        b.scene_lo[k] = scene_lo[k];
        const float ext = scene_hi[k] - scene_lo[k];
        b.scene_inv_ext[k] = ext > 0 ? 1.0f / ext : 0.0f;
    }
    b.code = code.as<uint32_t>(); b.index = index.as<uint32_t>(); b.box_in = box_in.as<LbvhBox>();
    b.tris = tris_out; b.ids = ids_out;
    b.child = child.as<int2>(); b.range = range.as<int2>(); b.parent = parent.as<int>(); b.node_box = node_box.as<LbvhBox>();
    b.visit = visit.as<int>(); b.live = live.as<int>(); b.new_index = new_index.as<int>();
    const int block = 256, grid = 148 * 8;

    lbvh_for_each<k_box_and_code><<<grid, block, 0, stream>>>(b, n);
    {   // stable radix sort of (code, index): equal codes keep their input order, which lbvh_delta relies on
        size_t tmp_bytes = 0;
        LBVH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, code.as<uint32_t>(), code2.as<uint32_t>(), index.as<uint32_t>(), index2.as<uint32_t>(), n, 0, 30, stream));
        Scratch tmp(tmp_bytes);
        LBVH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, code.as<uint32_t>(), code2.as<uint32_t>(), index.as<uint32_t>(), index2.as<uint32_t>(), n, 0, 30, stream));
        LBVH_CUDA(cudaStreamSynchronize(stream));
    }
    b.code = code2.as<uint32_t>(); b.index = index2.as<uint32_t>();
    lbvh_for_each<k_gather><<<grid, block, 0, stream>>>(b, n);
    if (n <= kLbvhLeaf) {  // the whole scene is one leaf
        LBVH_CUDA(cudaStreamSynchronize(stream));
        LBVH_CUDA(cudaGetLastError());
        return ~(n - 1);  // first = 0, count = n
    }
    lbvh_for_each<k_build_node><<<grid, block, 0, stream>>>(b, ni);
    lbvh_for_each<k_fit><<<grid, block, 0, stream>>>(b, n);
    {
        size_t tmp_bytes = 0;
        LBVH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b.live, b.new_index, ni, stream));
        Scratch tmp(tmp_bytes);
        LBVH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, b.live, b.new_index, ni, stream));
        LBVH_CUDA(cudaStreamSynchronize(stream));
    }
    int last_live = 0, last_index = 0;
    LBVH_CUDA(cudaMemcpy(&last_live, b.live + (ni - 1), 4, cudaMemcpyDeviceToHost));
    LBVH_CUDA(cudaMemcpy(&last_index, b.new_index + (ni - 1), 4, cudaMemcpyDeviceToHost));
    const int n_nodes = last_index + last_live;
    LBVH_CUDA(cudaMalloc(nodes_out, 64 * (size_t)std::max(n_nodes, 1)));
    b.nodes_out = *nodes_out;
    lbvh_for_each<k_emit><<<grid, block, 0, stream>>>(b, ni);
    LBVH_CUDA(cudaStreamSynchronize(stream));
    LBVH_CUDA(cudaGetLastError());
    *n_nodes_out = n_nodes;
    return 0;  // the root of the radix tree is live (n > kLbvhLeaf) and the exclusive scan numbers it 0
}

}  // namespace btrc
