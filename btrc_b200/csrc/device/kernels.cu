// Wavefront path-tracing kernels for sm_100a (see kernels.h for the reference stage each one replaces).
//
// Execution model: every stage is a persistent, fixed-size grid (a multiple of the SM count) that reads its work
// count from the device-resident WaveCtl block, so the host can enqueue whole iterations back to back without the
// per-iteration StateCounters read-back of the reference (wavefront.cpp:283-284).  Queues are SoA float4 arrays:
// each lane issues 16-byte coalesced loads/stores.  Output slots are allocated with one atomic per warp
// (ballot + popc prefix) instead of one per thread (shade.cpp:165,225).
#include "kernels.h"

#include "dev_math.cuh"
#include "rng.cuh"
#include "shading.cuh"
#include "path_index.cuh"
#include "traverse.cuh"

namespace btrc {

namespace {

constexpr int kBlock = 256;
// 1: the dynamic work cursors are advanced one round ahead of their use, to hide the atomic's round trip.  Measured on B200 and NOT
// adopted (profiles/r2h_ab_kernel_variants.md: C2 1123 with, 1159 Msamples/s without -- the value carried across the round costs the
// register-capped kernels more than the hidden latency returns); kept as a build switch for the record.
#ifndef BTRC_CURSOR_AHEAD
#define BTRC_CURSOR_AHEAD 0
#endif
constexpr unsigned kFull = 0xffffffffu;
// TRAV: loop form / hierarchy of the walk (kernels.h LaunchCfg::traverse): 0 while-while, 1 if-if, 2 two-level (instanced scenes), 3 compressed 8-wide
template <bool ANY_HIT, int TRAV>
__device__ __forceinline__ HitRec traverse_warp(const BvhDevice &bvh, float4 o4, float4 d4, bool valid) {  // all 32 lanes call this together
    if (TRAV == 3) return traverse_bvh8<ANY_HIT, true>(bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w, valid);
    if (TRAV == 2) return traverse_two_level<ANY_HIT, true>(bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w, valid);
    if (TRAV == 1) return traverse_bvh2_ifif<ANY_HIT, true>(bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w, valid);
    return traverse_bvh2<ANY_HIT, true>(bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w, valid);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Hand-out of 32-entry batches to the warps of a persistent kernel.  BTRC_STATIC_SHARE (percent, default 0 = all dynamic): the first
// share of the batches is dealt statically (warp w of T takes batches w, w + T, ...: no atomic, neighbouring batches still run on
// neighbouring warps at the same time), the rest through the atomic cursor, which keeps the tail balanced.
#ifndef BTRC_STATIC_SHARE
#define BTRC_STATIC_SHARE 0
#endif
struct BatchCursor {
    int round, static_rounds, total_warps, warp_gid;
    __device__ __forceinline__ void init(int n) {
        total_warps = gridDim.x * (blockDim.x >> 5);
        warp_gid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int n_batches = (n + 31) >> 5;
        static_rounds = BTRC_STATIC_SHARE ? int((long long)n_batches * BTRC_STATIC_SHARE / 100 / total_warps) : 0;
        round = 0;
    }
    // first entry of this warp's next batch, or a value >= n when the work is gone
    __device__ __forceinline__ int next(int *cursor) {
        if (BTRC_STATIC_SHARE && round < static_rounds) return (round++ * total_warps + warp_gid) * 32;
        int base = 0;
        if (lane_id() == 0) base = atomicAdd(cursor, 32);
        return __shfl_sync(0xffffffffu, base, 0) + (BTRC_STATIC_SHARE ? static_rounds * total_warps * 32 : 0);
    }
};

// core/film.cpp:131-165: bounds are guaranteed by construction, the isfinite test is kept
__device__ __forceinline__ void film_splat3(float *buf, uint32_t pixel, f3 v) {
    if (isfinite3(v)) {
        float *p = buf + 4 * (size_t)pixel;
        atomicAdd(p + 0, v.x);
        atomicAdd(p + 1, v.y);
        atomicAdd(p + 2, v.z);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// prepare: device-side bookkeeping of one iteration (wavefront.cpp:227-306, generate.cpp:140-177)
// ---------------------------------------------------------------------------------------------------------------
__global__ void prepare_kernel(WaveCtl *ctl) {
    const int survivors = ctl->n_next;
    ctl->n_next_total += (unsigned long long)survivors;
    ctl->n_shadow_total += (unsigned long long)ctl->n_shadow;
    const long long remaining = ctl->end - ctl->cursor;
    const long long room = (long long)ctl->pool - survivors;
    const int n_new = (int)(remaining < room ? remaining : room);
    ctl->gen_base = ctl->cursor;
    ctl->cursor += n_new;
    ctl->n_survivors = survivors;
    ctl->n_sorted = ctl->sort_enabled ? survivors : 0;
    ctl->n_shadow_sorted = 0;
    ctl->n_new = n_new;
    ctl->n_active = survivors + n_new;
    ctl->n_next = 0;
    ctl->n_shadow = 0;
    ctl->trace_cursor = 0;
    ctl->shadow_cursor = 0;
    for (int c = 0; c < 8; ++c) { ctl->class_count[c] = 0; ctl->class_cursor[c] = 0; }
    ctl->n_gen += (unsigned long long)n_new;
    ctl->n_closest += (unsigned long long)(survivors + n_new);
    ctl->n_iter += (survivors + n_new) > 0 ? 1 : 0;
    ctl->done = (survivors + n_new) == 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// generate (generate.cpp:39-86, film_filter/box.cpp:5-10, gaussian.cpp:54-66, camera/pinhole.cpp:84-99)
// ---------------------------------------------------------------------------------------------------------------
template <bool HALTON>
struct GenOut {
    f3 o, d;
    uint32_t px, py;
    SamplerT<HALTON> sampler;
};

template <bool HALTON>
__device__ __forceinline__ GenOut<HALTON> generate_one(const SceneDevice &scene, const RenderParams &P, int pixel_x, int pixel_y, int sample_index) {
    GenOut<HALTON> g;
    const int W = P.width, H = P.height;
    g.px = pixel_x; g.py = pixel_y;
    g.sampler.init(scene.halton, (uint32_t)pixel_x, (uint32_t)pixel_y, sample_index);
    float fx, fy;
    if (P.filter_type == 0) {
        fx = g.sampler.get1d() - 0.5f;
        fy = g.sampler.get1d() - 0.5f;
    } else {
        constexpr int GRID = 256;
        const uint32_t gi = alias_sample(scene.filter_alias, GRID * GRID, g.sampler.get1d());
        const uint32_t gy = gi / GRID, gx = gi % GRID;
        const float r = P.filter_radius, step = P.filter_step;
        const float x_beg = __fadd_rn(-r, __fmul_rn(step, float(gx))), x_end = __fadd_rn(-r, __fmul_rn(step, float(gx + 1)));
        const float y_beg = __fadd_rn(-r, __fmul_rn(step, float(gy))), y_end = __fadd_rn(-r, __fmul_rn(step, float(gy + 1)));
        fx = lerpf(x_beg, x_end, g.sampler.get1d());
        fy = lerpf(y_beg, y_end, g.sampler.get1d());
    }
    const float pixel_xf = float(pixel_x) + 0.5f + fx;
    const float pixel_yf = float(pixel_y) + 0.5f + fy;
    const float film_x = pixel_xf / float(W);
    const float film_y = pixel_yf / float(H);
    (void)g.sampler.get1d();  // time_sample: drawn, unused by the pinhole camera (generate.cpp:76)
    const CameraInfo &c = scene.camera;
    const f3 x = film_x * f3(c.film_x[0], c.film_x[1], c.film_x[2]);
    const f3 y = film_y * f3(c.film_y[0], c.film_y[1], c.film_y[2]);
    const f3 dst = f3(c.left_bottom_corner[0], c.left_bottom_corner[1], c.left_bottom_corner[2]) + x + y;
    g.o = f3(c.eye[0], c.eye[1], c.eye[2]);
    g.d = normalize(dst - g.o);
    return g;
}

// generate.cpp:39-86.  The kernel's own part is WHICH (pixel, sample) the i-th new path of the wave is: path_index.cuh.
template <bool HALTON>
__global__ void __launch_bounds__(kBlock) generate_kernel(SceneDevice scene, RenderParams P, PathQueue q, FilmDevice film, const WaveCtl *ctl) {
    const int n_new = ctl->n_new;
    const int base_slot = ctl->n_survivors;
    const long long gen_base = ctl->gen_base;
    const bool splat_weight = ctl->weight_prefilled == 0;
    const PathIndexMap map = make_path_index_map(*ctl, P.width, P.height, P.spp, P.tile);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x) {
        uint32_t x, y, sample;
        decode_path_index(map, gen_base + i, x, y, sample);
        GenOut<HALTON> g = generate_one<HALTON>(scene, P, (int)x, (int)y, (int)sample);
        const uint32_t pixel = g.py * (uint32_t)P.width + g.px;
        // generate.cpp:79: film.splat_atomic(pixel, weight, 1).  A render over a known index range fills the plane beforehand
        // (weight_fill_kernel: the filter keeps every sample inside its own pixel, so the count per pixel is a closed form)
        if (splat_weight) atomicAdd(film.weight + pixel, 1.0f);
        const int s = base_slot + i;
        q.ray_o[s] = make4(g.o, __uint_as_float(scene.vol_prim_medium_id));
        q.ray_d[s] = make4(g.d, kMaxFloat);
        q.beta[s] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
        q.rad[s] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(pixel));
        q.rng[s] = g.sampler.store();
        q.ble[s] = make_float4(1.0f, 1.0f, 1.0f, -1.0f);
    }
}

// Weight plane of a render over the (virtual, see WaveCtl) path index range [begin, end): adds to every pixel the number of
// camera paths of that range that start in it -- the sum of generate.cpp:79's splats of 1.0f, exact in fp32 below 2^24.
__global__ void __launch_bounds__(kBlock) weight_fill_kernel(float *weight, int npix, int spp, int tile, long long begin, long long end,
                                                             long long sblock, int sparts, int spart, int sbw, int srows, int width) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        long long count = 0;
        if (sblock <= 0) {
            if (!tile) {  // index = sample * npix + pixel: samples s with begin <= s * npix + p < end
                const long long lo = begin - p, hi = end - p;  // s * npix in [lo, hi)
                const long long s0 = lo <= 0 ? 0 : (lo + npix - 1) / npix, s1 = hi <= 0 ? 0 : (hi + npix - 1) / npix;
                count = s1 > s0 ? s1 - s0 : 0;
            } else {      // index = pixel * spp + sample
                const long long a = (long long)p * spp, b = a + spp;
                const long long lo = a > begin ? a : begin, hi = b < end ? b : end;
                count = hi > lo ? hi - lo : 0;
            }
        } else {
            // strided partitions are made of whole pixels in tile mode (stride_block is a multiple of spp): pixel p is ours when
            // its block belongs to this part and its virtual index lies inside [begin, end)
            long long a = (long long)p * spp;
            long long blk = a / sblock;
            if (sbw > 0) {  // 2-D tiles: tile of the pixel and the pixel's ordinal inside it
                const long long x = p % width, y = p / width;
                blk = (y / srows) * (width / sbw) + x / sbw;
                a = blk * sblock + ((y % srows) * sbw + x % sbw) * (long long)spp;
            }
            if (blk % sparts == spart) {
                const long long v = (blk / sparts) * sblock + a % sblock;
                const long long lo = v > begin ? v : begin, hi = v + spp < end ? v + spp : end;
                count = hi > lo ? hi - lo : 0;
            }
        }
        if (count) weight[p] += (float)count;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// closest hit (trace.cpp:38-78)
// ---------------------------------------------------------------------------------------------------------------
// returns the instance id of the hit, or -1 on a miss
__device__ __forceinline__ int store_hit(const BvhDevice &bvh, const HitRec &h, int i, uint32_t *flag, uint4 *detail) {
    if (h.slot < 0) {
        flag[i] = 0;  // miss: flag only (save_flag(false, false, 0))
        return -1;
    }
    const uint2 id = __ldg(bvh.tri_ids + h.slot);
    flag[i] = id.x | kFlagHasIntersection;
    detail[i] = make_uint4(__float_as_uint(h.t), id.y, __float_as_uint(h.u), __float_as_uint(h.v));
    return (int)id.x;
}

// Appends state index i to the list of shade class c: one atomic per warp and class (match.any groups the lanes).
// All lanes of the warp must call this (inactive ones with valid = false).
__device__ __forceinline__ void class_append(const HitQueue &hq, WaveCtl *ctl, bool valid, int c, int i) {
    const unsigned peers = __match_any_sync(kFull, valid ? c : -1);
    if (!valid) return;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(&ctl->class_count[c], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    hq.class_list[(size_t)c * hq.class_stride + base + __popc(peers & ((1u << lane_id()) - 1u))] = (uint32_t)i;
}

// Occupancy of the traversal kernels: 6 resident CTAs per SM (40 registers, 48 warps) measured 2.4 % faster on C2 than the
// compiler's own 46 registers / 5 CTAs, 8 CTAs (32 registers, spills) 3 % slower (profiles/r1h_ab_occupancy.log)
#ifndef BTRC_TRACE_MIN_BLOCKS
#define BTRC_TRACE_MIN_BLOCKS 6
#endif
// the two-level walk keeps the world ray next to the object-space slab constants: 5 CTAs (51 registers) instead of 6 (40)
#ifndef BTRC_TRACE_MIN_BLOCKS_2L
#define BTRC_TRACE_MIN_BLOCKS_2L 5
#endif
// the 8-wide walk tests eight quantised boxes per step in registers: 4 CTAs (64 registers)
#ifndef BTRC_TRACE_MIN_BLOCKS_W8
#define BTRC_TRACE_MIN_BLOCKS_W8 4
#endif
template <int TRAV> constexpr int trace_min_blocks() { return TRAV == 3 ? BTRC_TRACE_MIN_BLOCKS_W8 : TRAV == 2 ? BTRC_TRACE_MIN_BLOCKS_2L : BTRC_TRACE_MIN_BLOCKS; }
// direction octant << 9 | 8x8x8 cell of the origin over the world box
__device__ __forceinline__ uint32_t ray_sort_key(const SceneDevice &scene, f3 o, f3 d) {
    const int cx = min(max(int((o.x - scene.sort_lo[0]) * scene.sort_scale[0]), 0), 7);
    const int cy = min(max(int((o.y - scene.sort_lo[1]) * scene.sort_scale[1]), 0), 7);
    const int cz = min(max(int((o.z - scene.sort_lo[2]) * scene.sort_scale[2]), 0), 7);
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    return (oct << 9) | uint32_t(cx | (cy << 3) | (cz << 6));
}

// ---- binned ray order: histogram of the keys, exclusive scan of the 4096 bins, scatter of the indices --------------------------
// block-private histograms in shared memory: rays cluster in a few bins, global atomics on those would serialise
constexpr int kSortBlock = 512;
// (the keys are computed here from the rays, not by the shade kernels that wrote them: those run at their register limit)
__global__ void __launch_bounds__(kSortBlock) sort_hist_kernel(SceneDevice scene, const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d,
                                                              uint32_t *__restrict__ key, const int *count, int *bin) {
    __shared__ int sh[kSortBins];
    for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) sh[b] = 0;
    __syncthreads();
    const int n = *count;
    for (int i = blockIdx.x * kSortBlock + threadIdx.x; i < n; i += gridDim.x * kSortBlock) {
        const uint32_t k = ray_sort_key(scene, xyz(ray_o[i]), xyz(ray_d[i]));
        key[i] = k;
        atomicAdd(&sh[k], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) if (sh[b]) atomicAdd(&bin[b], sh[b]);
}
__global__ void __launch_bounds__(1024) sort_scan_kernel(int *bin) {  // one block: exclusive scan of kSortBins counters (4 per thread)
    __shared__ int part[1024];
    const int t = threadIdx.x;
    int v[4], sum = 0;
    for (int k = 0; k < 4; ++k) { v[k] = bin[4 * t + k]; sum += v[k]; }
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int add = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    int base = part[t] - sum;
    for (int k = 0; k < 4; ++k) { bin[4 * t + k] = base; base += v[k]; }
}
// every block ranks one contiguous chunk of the keys: counts in shared memory, one global reservation per (block, bin), local ranks
__global__ void __launch_bounds__(kSortBlock) sort_scatter_kernel(const uint32_t *__restrict__ key, const int *count, int *bin, uint32_t *order, int *n_sorted_out) {
    __shared__ int sh_cnt[kSortBins];
    __shared__ int sh_base[kSortBins];
    const int n = *count;
    const int chunk = (n + gridDim.x - 1) / gridDim.x;
    const int lo = min(blockIdx.x * chunk, n), hi = min(lo + chunk, n);
    for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) sh_cnt[b] = 0;
    __syncthreads();
    for (int i = lo + threadIdx.x; i < hi; i += kSortBlock) atomicAdd(&sh_cnt[key[i]], 1);
    __syncthreads();
    for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) {
        const int c = sh_cnt[b];
        sh_base[b] = c ? atomicAdd(&bin[b], c) : 0;
        sh_cnt[b] = 0;
    }
    __syncthreads();
    for (int i = lo + threadIdx.x; i < hi; i += kSortBlock) {
        const uint32_t k = key[i];
        order[sh_base[k] + atomicAdd(&sh_cnt[k], 1)] = (uint32_t)i;
    }
    if (n_sorted_out && blockIdx.x == 0 && threadIdx.x == 0) *n_sorted_out = n;
}

// SORTED: surviving paths are visited through RaySort::order and the class lists are left to classify_kernel (state order)
template <int TRAV, bool SORTED>
__global__ void __launch_bounds__(kBlock, trace_min_blocks<TRAV>()) trace_kernel(SceneDevice scene, PathQueue q, HitQueue hq, RaySort rs, WaveCtl *ctl) {
    const int n = ctl->n_active;
    const int n_sorted = SORTED ? ctl->n_sorted : 0;
    const bool append = !SORTED || ctl->sort_enabled != 2;
    unsigned long long hits = 0;
    BatchCursor bc;
    bc.init(n);
    for (;;) {
        const int base = bc.next(&ctl->trace_cursor);
        if (base >= n) break;
        const int j = base + lane_id();
        const bool valid = j < n;
        const int i = (SORTED && valid && j < n_sorted) ? (int)rs.order[j] : j;  // survivors bin by bin, new camera paths in pixel order
        float4 o4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d4 = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
        if (valid) { o4 = q.ray_o[i]; d4 = q.ray_d[i]; }
        const HitRec h = traverse_warp<false, TRAV>(scene.bvh, o4, d4, valid);
        int cls = 0;
        if (valid) {
            const int inst = store_hit(scene.bvh, h, i, hq.flag, hq.detail);
            if (inst >= 0) { cls = __ldg(scene.instance_class + inst); hits += 1; }
        }
        if (append) class_append(hq, ctl, valid, cls, i);
    }
    for (int off = 16; off; off >>= 1) hits += __shfl_down_sync(kFull, hits, off);
    if (lane_id() == 0 && hits) atomicAdd(&ctl->n_hit, hits);
}

// PROFILING ONLY (BTRC_PROBE_TOP_LEVELS=<depth>, never launched otherwise): replays the closest-hit walk of the current wave but
// touches only the nodes of the first `max_depth` levels of the tree -- subtrees below are skipped, triangles never fetched, nothing
// is written.  Under ncu its lts__t_sector_hit_rate / l1tex__t_sector_hit_rate are the cache hit rates of the TOP BVH LEVELS on real
// ray batches (the counters cannot be split by address inside trace_kernel itself); profiles/README.md.
__global__ void __launch_bounds__(kBlock) top_levels_probe_kernel(SceneDevice scene, PathQueue q, const WaveCtl *ctl, int max_depth, unsigned long long *sink) {
    const int n = ctl->n_active;
    unsigned long long visited = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = q.ray_o[i], d4 = q.ray_d[i];
        constexpr float kTiny = 1e-20f;
        const float idx = 1.0f / (fabsf(d4.x) > kTiny ? d4.x : copysignf(kTiny, d4.x));
        const float idy = 1.0f / (fabsf(d4.y) > kTiny ? d4.y : copysignf(kTiny, d4.y));
        const float idz = 1.0f / (fabsf(d4.z) > kTiny ? d4.z : copysignf(kTiny, d4.z));
        const float oox = o4.x * idx, ooy = o4.y * idy, ooz = o4.z * idz;
        int stack[64], depth_of[64];
        int sp = 0;
        if (scene.bvh.root_ref >= 0) { stack[sp] = scene.bvh.root_ref; depth_of[sp++] = 1; }
        while (sp > 0) {
            const int cur = stack[--sp], depth = depth_of[sp];
            const float4 *node = scene.bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = __ldg(node), n1 = __ldg(node + 1), nz = __ldg(node + 2), meta = __ldg(node + 3);
            ++visited;
            const float c0min = fmaxf(fmaxf(fminf(fmaf(n0.x, idx, -oox), fmaf(n0.y, idx, -oox)), fminf(fmaf(n0.z, idy, -ooy), fmaf(n0.w, idy, -ooy))),
                                      fmaxf(fminf(fmaf(nz.x, idz, -ooz), fmaf(nz.y, idz, -ooz)), 0.0f));
            const float c0max = fminf(fminf(fmaxf(fmaf(n0.x, idx, -oox), fmaf(n0.y, idx, -oox)), fmaxf(fmaf(n0.z, idy, -ooy), fmaf(n0.w, idy, -ooy))),
                                      fminf(fmaxf(fmaf(nz.x, idz, -ooz), fmaf(nz.y, idz, -ooz)), d4.w));
            const float c1min = fmaxf(fmaxf(fminf(fmaf(n1.x, idx, -oox), fmaf(n1.y, idx, -oox)), fminf(fmaf(n1.z, idy, -ooy), fmaf(n1.w, idy, -ooy))),
                                      fmaxf(fminf(fmaf(nz.z, idz, -ooz), fmaf(nz.w, idz, -ooz)), 0.0f));
            const float c1max = fminf(fminf(fmaxf(fmaf(n1.x, idx, -oox), fmaf(n1.y, idx, -oox)), fmaxf(fmaf(n1.z, idy, -ooy), fmaf(n1.w, idy, -ooy))),
                                      fminf(fmaxf(fmaf(nz.z, idz, -ooz), fmaf(nz.w, idz, -ooz)), d4.w));
            const int r0 = __float_as_int(meta.x), r1 = __float_as_int(meta.y);
            if (depth < max_depth) {
                if (c1min <= c1max && r1 >= 0) { stack[sp] = r1; depth_of[sp++] = depth + 1; }
                if (c0min <= c0max && r0 >= 0) { stack[sp] = r0; depth_of[sp++] = depth + 1; }
            }
        }
    }
    if (visited == 0xffffffffffffffffull) *sink = visited;  // keeps the loads alive
}

// Builds the class lists from hit flags (stage-level entry points: the hits come from the caller, not from trace_kernel)
__global__ void __launch_bounds__(kBlock) classify_kernel(SceneDevice scene, HitQueue hq, WaveCtl *ctl) {
    const int n = ctl->n_active;
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        int cls = 0;
        if (i < n) {
            const uint32_t flag = hq.flag[i];
            if (flag & kFlagHasIntersection) cls = __ldg(scene.instance_class + (flag & kFlagInstanceMask));
        }
        class_append(hq, ctl, i < n, cls, i);
    }
}

__global__ void __launch_bounds__(kBlock) trace_batch_kernel(SceneDevice scene, const float4 *ray_o, const float4 *ray_d, int n,
                                                            uint32_t *flag, uint4 *detail) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ray_o[i], d4 = ray_d[i];
        const HitRec h = scene.bvh.wide        ? traverse_bvh8<false, false>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w)
                         : scene.bvh.two_level ? traverse_two_level<false>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w)
                                               : traverse_bvh2<false>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w);
        if (h.slot < 0) detail[i] = make_uint4(0, 0, 0, 0);
        store_hit(scene.bvh, h, i, flag, detail);
    }
}

__global__ void __launch_bounds__(kBlock) anyhit_batch_kernel(SceneDevice scene, const float4 *ray_o, const float4 *ray_d, int n, uint8_t *occluded) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ray_o[i], d4 = ray_d[i];
        const HitRec h = scene.bvh.wide        ? traverse_bvh8<true, false>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w)
                         : scene.bvh.two_level ? traverse_two_level<true>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w)
                                               : traverse_bvh2<true>(scene.bvh, o4.x, o4.y, o4.z, d4.x, d4.y, d4.z, d4.w);
        occluded[i] = h.slot >= 0 ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Queue appends shared by the medium and shade kernels: one atomic per warp and queue (ballot + popc prefix)
// instead of the reference's one atomic per thread (shade.cpp:165,225; medium.cpp:131,151).
// ---------------------------------------------------------------------------------------------------------------
struct NextState {
    f3 o, d, beta, L, ble;
    float pdf;
    uint32_t med;
    int depth;
    ulonglong2 rng;
};
struct ShadowOut {
    f3 o, d, beta;
    float t1;
    uint32_t med;
};

__device__ __forceinline__ void emit_outputs(PathQueue &nq, ShadowQueue &sq, FilmDevice &film, WaveCtl *ctl, unsigned long long &n_splat, uint32_t pixel, bool emit_next,
                                             const NextState &ns, bool emit_shadow, const ShadowOut &so, bool splat, f3 splat_value) {
    // one 64-bit atomic per warp reserves the slots of both queues (n_next | n_shadow << 32 share an aligned word in WaveCtl)
    const unsigned m_next = __ballot_sync(kFull, emit_next), m_shadow = __ballot_sync(kFull, emit_shadow);
    unsigned long long both = 0;
    if ((m_next | m_shadow) && lane_id() == 0)
        both = atomicAdd(reinterpret_cast<unsigned long long *>(&ctl->n_next), (unsigned long long)__popc(m_next) | ((unsigned long long)__popc(m_shadow) << 32));
    both = __shfl_sync(kFull, both, 0);
    const unsigned below = (1u << lane_id()) - 1u;
    const int next_slot = int(both & 0xffffffffull) + __popc(m_next & below);
    if (emit_next) {
        nq.ray_o[next_slot] = make4(ns.o, __uint_as_float(ns.med));
        nq.ray_d[next_slot] = make4(ns.d, kMaxFloat);
        nq.beta[next_slot] = make4(ns.beta, __int_as_float(ns.depth));
        nq.rad[next_slot] = make4(ns.L, __uint_as_float(pixel));
        nq.rng[next_slot] = ns.rng;
        nq.ble[next_slot] = make4(ns.ble, ns.pdf);
    }
    const int shadow_slot = int(both >> 32) + __popc(m_shadow & below);
    if (emit_shadow) {
        sq.ray_o[shadow_slot] = make4(so.o, __uint_as_float(so.med));
        sq.ray_d[shadow_slot] = make4(so.d, so.t1);
        sq.beta_li[shadow_slot] = make4(so.beta, __uint_as_float(pixel));
    }
    if (splat) { film_splat3(film.radiance, pixel, splat_value); n_splat += 1; }  // the statistic is summed per thread, one atomic per warp at kernel end
}

__device__ __forceinline__ void flush_splat_count(WaveCtl *ctl, unsigned long long n_splat) {
    for (int off = 16; off; off >>= 1) n_splat += __shfl_down_sync(kFull, n_splat, off);
    if (lane_id() == 0 && n_splat) atomicAdd(&ctl->n_splat, n_splat);
}

// ---------------------------------------------------------------------------------------------------------------
// medium (medium.cpp:36-164, 214-249): only launched when the scene has media (wavefront.cpp:254-268)
// ---------------------------------------------------------------------------------------------------------------
// CTAs per SM the medium kernel is compiled for (A/B: profiles/r2o_ab_medium_occupancy.md)
#ifndef BTRC_MEDIUM_MIN_BLOCKS
#define BTRC_MEDIUM_MIN_BLOCKS 2
#endif
template <bool HALTON, bool T3P>
__global__ void __launch_bounds__(kBlock, BTRC_MEDIUM_MIN_BLOCKS) medium_kernel(SceneDevice scene, RenderParams P, PathQueue q, HitQueue hq, PathQueue nq, ShadowQueue sq,
                                                       FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->n_active;
    const int n_round = (n + 31) & ~31;
    unsigned long long n_splat = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool emit_next = false, emit_shadow = false, splat = false;
        f3 splat_value(0.0f);
        uint32_t pixel = 0;
        NextState ns;
        ShadowOut so;
        if (i < n) {
            const uint32_t flag = hq.flag[i];
            const float4 o4 = q.ray_o[i], d4 = q.ray_d[i];
            const f3 o = xyz(o4), d = xyz(d4);
            // get_medium_segment
            uint32_t medium_id;
            f3 medium_end;
            if (flag & kFlagHasIntersection) {
                const uint4 det = hq.detail[i];
                const InstanceInfo &inst = scene.instances[flag & kFlagInstanceMask];
                const f3 inct_nor = xf_normal(inst.inv, xyz(ldg4(inst.gz_uca + det.y)));
                const uint32_t inct_medium_id = dot(d, inct_nor) < 0.0f ? inst.outer_medium_id : inst.inner_medium_id;
                medium_id = min(inct_medium_id, __float_as_uint(o4.w));  // resolve_mediums: smaller id = higher priority
                medium_end = intersection_offset(o + __uint_as_float(det.x) * d, inct_nor, -d);
            } else {
                medium_id = scene.vol_prim_medium_id;
                medium_end = o + normalize(d) * scene.world_diagonal;
            }
            q.ray_o[i] = make4(o, __uint_as_float(medium_id));
            const MediumDesc &medium = scene.media[medium_id];
            const bool tracked = medium.type != MED_VOLPRIM || scene.n_vol_prims > 0;  // an empty volume-primitive medium: Tr = 1, no draws
            // a path outside every medium (the common case) touches 52 of its 148 queue bytes: throughput, radiance and sampler are read only
            // when the segment is tracked
            float4 beta4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rad4 = beta4;
            SamplerT<HALTON> sampler;
            MediumSample ms;
            ms.scattered = false;
            f3 unscatter_tr(1.0f);
            if (tracked) {
                beta4 = q.beta[i]; rad4 = q.rad[i];
                sampler.load(q.rng[i], scene.halton);
                ms = medium_sample<T3P>(scene, medium, o, medium_end, sampler);
                unscatter_tr = medium_tr<T3P>(scene, medium, o, medium_end, sampler);
            }
            pixel = __float_as_uint(rad4.w);
            if (tracked) {
                const float4 ble4 = q.ble[i];
                q.ble[i] = make4(xyz(ble4) * unscatter_tr, ble4.w);
            }
            if (!ms.scattered) {
                if (tracked) q.rng[i] = sampler.store();
            } else {
                hq.flag[i] = flag | kFlagHasScattering;
                f3 beta = xyz(beta4);  // throughput of both media is one
                const f3 L = xyz(rad4);
                const int depth = __float_as_int(beta4.w);
                if (simple_russian_roulette(P, beta, depth, sampler)) {
                    splat = true; splat_value = L;
                } else {
                    LiSample ls = sample_light_li<false>(scene, ms.position, f3(0.0f), sampler);
                    const f3 wo = -d;
                    if (ls.success) {
                        const float ph = ms.shader.pdf(ls.d, wo);  // eval = color * pdf
                        so.beta = ls.li * beta * (ms.shader.color * ph) / (ph + ls.light_pdf);
                        so.o = ls.o; so.d = ls.d; so.t1 = ls.t1;
                        so.med = ls.t1 > 1.0f ? scene.vol_prim_medium_id : medium_id;
                        emit_shadow = true;
                    }
                    f3 phase, pdir;
                    float ppdf;
                    sampler.begin3d();  // phase sample: get3d (medium.cpp:117)
                    const float p0 = sampler.get1d(), p1 = sampler.get1d();
                    (void)sampler.get1d();
                    ms.shader.sample(wo, p0, p1, phase, pdir, ppdf);
                    if (!is_zero(phase)) {
                        ns.d = normalize(pdir);
                        ns.ble = beta * phase;
                        ns.beta = ns.ble / ppdf;
                        ns.pdf = ppdf;
                        ns.o = ms.position; ns.med = medium_id; ns.L = L; ns.depth = depth + 1;
                        ns.rng = sampler.store();
                        emit_next = true;
                    } else {
                        splat = true; splat_value = L;
                    }
                }
            }
        }
        emit_outputs(nq, sq, film, ctl, n_splat, pixel, emit_next, ns, emit_shadow, so, splat, splat_value);
    }
    flush_splat_count(ctl, n_splat);
}

// ---------------------------------------------------------------------------------------------------------------
// shade (shade.cpp:38-235), one kernel per shade class: shade_miss_kernel for rays that left the scene and
// shade_kernel<MAT> for hits on each material type.  The reference dispatches with a switch over per-object generated
// code inside one kernel (core/scene.cpp:358-374); here the closest-hit kernel has already binned the states.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) shade_miss_kernel(SceneDevice scene, PathQueue q, HitQueue hq, FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->class_count[0];
    const uint32_t *list = hq.class_list;
    unsigned long long splats = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {  // shade.cpp:54-69
        const int i = (int)list[j];
        const float4 rad4 = q.rad[i];
        f3 v(0.0f);
        if (scene.env.type != ENV_NONE) {
            const float4 d4 = q.ray_d[i], ble4 = q.ble[i];
            v = eval_miss_le(scene, xyz(d4), xyz(ble4), ble4.w);
        }
        if (!(hq.flag[i] & kFlagHasScattering)) v = v + xyz(rad4);
        film_splat3(film.radiance, __float_as_uint(rad4.w), v);
        splats += 1;
    }
    for (int off = 16; off; off >>= 1) splats += __shfl_down_sync(kFull, splats, off);
    if (lane_id() == 0 && splats) atomicAdd(&ctl->n_splat, splats);
}

#ifndef BTRC_SHADE_BLOCKS_HEAVY
#define BTRC_SHADE_BLOCKS_HEAVY 2
#endif
#ifndef BTRC_SHADE_BLOCKS_LIGHT
#define BTRC_SHADE_BLOCKS_LIGHT 3
#endif
#ifndef BTRC_SHADE_CTA_SYNC
#define BTRC_SHADE_CTA_SYNC 0
#endif
#ifndef BTRC_SHADE_BLOCK_HEAVY
#define BTRC_SHADE_BLOCK_HEAVY 256
#endif
template <int MAT> constexpr bool shade_is_heavy() { return MAT == MAT_DISNEY || MAT == MAT_METAL; }
template <int MAT> constexpr int shade_block() { return shade_is_heavy<MAT>() ? BTRC_SHADE_BLOCK_HEAVY : kBlock; }
template <int MAT, bool HALTON>
__global__ void __launch_bounds__(shade_block<MAT>(), shade_is_heavy<MAT>() ? BTRC_SHADE_BLOCKS_HEAVY : BTRC_SHADE_BLOCKS_LIGHT)
shade_kernel(SceneDevice scene, RenderParams P, PathQueue q, HitQueue hq, PathQueue nq, ShadowQueue sq, FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->class_count[1 + MAT];
    const uint32_t *list = hq.class_list + (size_t)(1 + MAT) * hq.class_stride;
    // Whole warps take part in the aggregated allocations.  The Disney kernel hands out batches of 32 list entries dynamically: its
    // states differ a lot in cost (a path ended by roulette against a full evaluation) and a fixed grid-stride split leaves a tail
    // (C2 -2.1 %); the light kernels keep the static split (dynamic: C4 +4 %, the atomic per batch shows).  profiles/r1k_ab_shade_fetch.log
    constexpr bool kDynamic = MAT == MAT_DISNEY;
    const int n_round = (n + 31) & ~31;
    unsigned long long n_splat = 0;
    int next_base = 0;
#if BTRC_SHADE_CTA_SYNC
    // experiment (profiles/r2h_ab_kernel_variants.md, not adopted: no gain): the CTA takes 256 states per round and its warps start every round together, so
    // that they run the same stretch of this 64 KB kernel at about the same time (instruction-cache sharing)
    __shared__ int s_base[2];
    if (kDynamic && threadIdx.x == 0) s_base[0] = atomicAdd(&ctl->class_cursor[1 + MAT], (int)blockDim.x);
    int round = 0;
#else
    if (BTRC_CURSOR_AHEAD && kDynamic && lane_id() == 0) next_base = atomicAdd(&ctl->class_cursor[1 + MAT], 32);
#endif
    for (int j = blockIdx.x * blockDim.x + threadIdx.x;; j += gridDim.x * blockDim.x) {
        if (kDynamic) {
#if BTRC_SHADE_CTA_SYNC
            __syncthreads();
            const int base = s_base[round & 1];
            if (base >= n) break;
            if (threadIdx.x == 0) s_base[(round + 1) & 1] = atomicAdd(&ctl->class_cursor[1 + MAT], (int)blockDim.x);
            ++round;
            j = base + (int)threadIdx.x;
            (void)next_base;
#else
            if (!BTRC_CURSOR_AHEAD && lane_id() == 0) next_base = atomicAdd(&ctl->class_cursor[1 + MAT], 32);
            const int base = __shfl_sync(kFull, next_base, 0);
            if (base >= n) break;
            if (BTRC_CURSOR_AHEAD && lane_id() == 0) next_base = atomicAdd(&ctl->class_cursor[1 + MAT], 32);
            j = base + lane_id();
#endif
        } else if (j >= n_round) {
            break;
        }
        bool emit_next = false, emit_shadow = false, splat = false;
        f3 splat_value(0.0f);
        uint32_t pixel = 0;
        NextState ns;
        ShadowOut so;

        if (j < n) {
            const int i = (int)list[j];
            const uint32_t flag = hq.flag[i];
            const bool scattered = (flag & kFlagHasScattering) != 0;
            const float4 rad4 = q.rad[i], o4 = q.ray_o[i], d4 = q.ray_d[i], ble4 = q.ble[i];
            pixel = __float_as_uint(rad4.w);
            f3 L = xyz(rad4);
            const f3 o = xyz(o4), d = xyz(d4);
            const f3 beta_le = xyz(ble4);
            const float bsdf_pdf = ble4.w;
            const uint4 det = hq.detail[i];
            const InstanceInfo &inst = scene.instances[flag & kFlagInstanceMask];
            const SurfacePoint inct = get_hitinfo(o, d, inst, __uint_as_float(det.x), det.y,
                                                  make_float2(__uint_as_float(det.z), __uint_as_float(det.w)));
            const f3 le_rad = handle_intersected_light(scene, o, d, inct, beta_le, bsdf_pdf, inst.light_id);
            if (scattered) {
                splat = true; splat_value = le_rad;
            } else {
                const float4 beta4 = q.beta[i];
                const ulonglong2 rng2 = q.rng[i];
                f3 beta = xyz(beta4);
                const int depth = __float_as_int(beta4.w);
                const uint32_t ray_medium = __float_as_uint(o4.w);
                SamplerT<HALTON> sampler;
                sampler.load(rng2, scene.halton);
                L = L + le_rad;
                if (simple_russian_roulette(P, beta, depth, sampler)) {
                    splat = true; splat_value = L;
                } else {
                    LiSample ls = sample_light_li<true>(scene, inct.position, inct.frame.z, sampler);
                    const Shader<MAT> shader = create_shader<MAT>(scene, scene.materials[inst.material_id], inct);
                    if (depth == 0) {
                        if (film.albedo) film_splat3(film.albedo, pixel, shader.albedo());
                        if (film.normal) film_splat3(film.normal, pixel, shader.normal());
                    }
                    const f3 wo = -d;
                    sampler.begin3d();  // bsdf sample: get3d (shade.cpp:143-144)
                    const float b0 = sampler.get1d(), b1 = sampler.get1d(), b2 = sampler.get1d();
                    BsdfSample bs = bsdf_none();
                    if (!shader.is_aggregate()) {
                        bs = shader.sample_delta(wo, b0);  // eval == 0: no shadow ray (shade.cpp:151)
                    } else {
                        const Frame3 shf = shader.flipped_shading(wo);
                        const f3 lwo = normalize(shf.global_to_local(wo));
                        const f3 l_sample = shader.sample_local_dir(lwo, b0, b1, b2);
                        const f3 l_light = ls.success ? normalize(shf.global_to_local(ls.d)) : f3(0.0f);
                        f3 f_sample(0.0f), f_light(0.0f);
                        float p_sample = 0.0f, p_light = 0.0f;
#pragma unroll 1
                        for (int k = 0; k < 2; ++k) {  // one copy of the closure's eval + pdf for both directions
                            const f3 lw = k == 0 ? l_sample : l_light;
                            f3 f(0.0f);
                            float p = 0.0f;
                            if (!is_zero(lw)) shader.comp_eval_pdf(lw, lwo, f, p);
                            if (k == 0) { f_sample = f; p_sample = p; } else { f_light = f; p_light = p; }
                        }
                        if (p_sample > 0.0f && !shader.reject_sample(f_sample, l_sample, p_sample)) {  // aggregate.cpp:84-117
                            const f3 dir = shf.local_to_global(l_sample);
                            if (!shader.frame.is_black_fringes(dir)) {
                                bs.dir = dir;
                                bs.bsdf = f_sample * shader.wrap_factor(shf.z, dir);
                                bs.pdf = p_sample;
                            }
                        }
                        if (ls.success) {
                            const f3 shadow_bsdf_val = f_light * shader.wrap_factor(shf.z, ls.d);  // aggregate.cpp:185-197
                            if (!is_zero(shadow_bsdf_val)) {
                                const bool shadow_d_out = dot(inct.frame.z, ls.d) > 0.0f;
                                const bool last_ray_out = dot(inct.frame.z, d) < 0.0f;
                                if (shadow_d_out == last_ray_out) so.med = ray_medium;
                                else if (shadow_d_out) so.med = inst.outer_medium_id;
                                else so.med = inst.inner_medium_id;
                                const float cosv = fabsf(dot(normalize(ls.d), inct.frame.z));
                                so.beta = ls.li * beta * shadow_bsdf_val * cosv / (p_light + ls.light_pdf);
                                so.o = ls.o; so.d = ls.d; so.t1 = ls.t1;
                                emit_shadow = true;
                            }
                        }
                    }
                    if (!is_zero(bs.bsdf)) {
                        const f3 dir = normalize(bs.dir);
                        const float cosv = fabsf(dot(dir, inct.frame.z));
                        ns.ble = beta * cosv * bs.bsdf;
                        ns.beta = ns.ble / bs.pdf;
                        ns.o = intersection_offset(inct.position, inct.frame.z, dir);
                        ns.d = dir;
                        ns.med = dot(inct.frame.z, dir) > 0.0f ? inst.outer_medium_id : inst.inner_medium_id;
                        ns.pdf = bs.is_delta ? -bs.pdf : bs.pdf;
                        ns.L = L; ns.depth = depth + 1;
                        ns.rng = sampler.store();
                        emit_next = true;
                    } else {
                        splat = true; splat_value = L;
                    }
                }
            }
        }
        emit_outputs(nq, sq, film, ctl, n_splat, pixel, emit_next, ns, emit_shadow, so, splat, splat_value);
    }
    flush_splat_count(ctl, n_splat);
}

// ---------------------------------------------------------------------------------------------------------------
// shadow rays (shadow.cpp:24-74): occlusion with early exit; unoccluded => splat beta_li * Tr
// ---------------------------------------------------------------------------------------------------------------
template <int TRAV, bool SORTED, bool T3P>
__global__ void __launch_bounds__(kBlock, trace_min_blocks<TRAV>()) shadow_kernel(SceneDevice scene, RenderParams P, ShadowQueue sq, RaySort rs, FilmDevice film, WaveCtl *ctl) {
    const int n = ctl->n_shadow;
    const int n_sorted = SORTED ? ctl->n_shadow_sorted : 0;
    unsigned long long misses = 0;
    BatchCursor bc;
    bc.init(n);
    for (;;) {
        const int base = bc.next(&ctl->shadow_cursor);
        if (base >= n) break;
        const int j = base + lane_id();
        const bool valid = j < n;
        const int i = (SORTED && valid && j < n_sorted) ? (int)rs.order[j] : j;
        float4 o4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d4 = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
        if (valid) { o4 = sq.ray_o[i]; d4 = sq.ray_d[i]; }
        const HitRec h = traverse_warp<true, TRAV>(scene.bvh, o4, d4, valid);
        if (valid && h.slot < 0) {  // unoccluded: the reference's miss program (shadow.cpp:39-74)
            const float4 b4 = sq.beta_li[i];
            f3 v = xyz(b4);
            if (scene.has_medium) v = v * shadow_tr<T3P>(scene, o4, d4);
            film_splat3(film.radiance, __float_as_uint(b4.w), v);
            misses += 1;
        }
    }
    for (int off = 16; off; off >>= 1) misses += __shfl_down_sync(kFull, misses, off);
    if (lane_id() == 0 && misses) {
        atomicAdd(&ctl->n_shadow_miss, misses);
        atomicAdd(&ctl->n_splat, misses);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// resolve (preview.cpp:28-80)
// ---------------------------------------------------------------------------------------------------------------
__global__ void resolve_kernel(int npix, const float4 *value, const float *weight, float4 *out, int normalise) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float4 v = value[i];
        const float w = weight[i];
        f3 o(0.0f);
        if (w > 0.0f) {
            o = xyz(v) / w;
            if (normalise) o = length(o) > 1e-3f ? normalize(o) : f3(0.0f);
        }
        out[i] = make4(o, 1.0f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// IBL set-up: generate_lum_sum_table (light/env_sampler.cpp:22-57) -- per LUT tile, mean luminance of the environment
// texture over 256 Hammersley taps (stretched by 1.2 around the tile) and the tile's solid angle
// ---------------------------------------------------------------------------------------------------------------
__global__ void ibl_lut_kernel(SceneDevice scene, int tex, int lut_w, int lut_h, int n_samples, float *lum_table, float *area_table) {
    const int xi = blockDim.x * blockIdx.x + threadIdx.x;
    const int yi = blockDim.y * blockIdx.y + threadIdx.y;
    if (xi >= lut_w || yi >= lut_h) return;
    const float x0 = float(xi) / lut_w, x1 = float(xi + 1) / lut_w;
    const float y0 = float(yi) / lut_h, y1 = float(yi + 1) / lut_h;
    float lum_sum = 0.0f;
    for (int i = 0; i < n_samples; ++i) {
        const float hx = float(i) / n_samples;
        const float hy = __brev((unsigned)i) * 2.3283064365386963e-10f;  // hammersley_radical_inverse (utils/math/hammersley.h:7-15)
        const float x = lerpf(x0, x1, 1.2f * hx - 0.1f), y = lerpf(y0, y1, 1.2f * hy - 0.1f);
        lum_sum += lum(eval_texture(scene, tex, make_float2(x, y)));
    }
    lum_table[yi * lut_w + xi] = lum_sum / n_samples;
    area_table[yi * lut_w + xi] = fabsf(2.0f * kPi * (x1 - x0) * (__cosf(kPi * y1) - __cosf(kPi * y0)));
}

inline int persistent_blocks(const LaunchCfg &cfg, int per_sm) { return cfg.sm_count * per_sm; }

}  // namespace

void launch_prepare(WaveCtl *ctl, cudaStream_t stream) { prepare_kernel<<<1, 1, 0, stream>>>(ctl); }

void launch_generate(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    if (params.halton) generate_kernel<true><<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, params, cur, film, ctl);
    else generate_kernel<false><<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, params, cur, film, ctl);
}

void launch_weight_fill(const LaunchCfg &cfg, const RenderParams &params, FilmDevice film, long long begin, long long end, long long stride_block,
                        int stride_parts, int stride_part, int stride_bw, int stride_rows, cudaStream_t stream) {
    weight_fill_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(film.weight, params.width * params.height, params.spp, params.tile ? 1 : 0,
                                                                         begin, end, stride_block, stride_parts, stride_part, stride_bw, stride_rows, params.width);
}

// counting sort of `count` keys into rs.order; *n_sorted_out (a WaveCtl field) is set to the count by the scatter kernel
void launch_ray_sort(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, uint32_t *key, const int *count, RaySort rs,
                     int *n_sorted_out, cudaStream_t stream) {
    cudaMemsetAsync(rs.bin, 0, sizeof(int) * (kSortBins + 1), stream);
    sort_hist_kernel<<<persistent_blocks(cfg, 2), kSortBlock, 0, stream>>>(scene, ray_o, ray_d, key, count, rs.bin);
    sort_scan_kernel<<<1, 1024, 0, stream>>>(rs.bin);
    sort_scatter_kernel<<<persistent_blocks(cfg, 2), kSortBlock, 0, stream>>>(key, count, rs.bin, rs.order, n_sorted_out);
}

void launch_trace(const LaunchCfg &cfg, const SceneDevice &scene, PathQueue cur, HitQueue hit, RaySort rs, WaveCtl *ctl, cudaStream_t stream) {
    const int blocks = persistent_blocks(cfg, 8);
    if (cfg.probe_top_levels > 0 && !scene.bvh.two_level && !scene.bvh.wide)
        top_levels_probe_kernel<<<blocks, kBlock, 0, stream>>>(scene, cur, ctl, cfg.probe_top_levels, reinterpret_cast<unsigned long long *>(&ctl->n_iter));
#define BTRC_TRACE(T, S) trace_kernel<T, S><<<blocks, kBlock, 0, stream>>>(scene, cur, hit, rs, ctl)
    const int trav = scene.bvh.wide ? 3 : scene.bvh.two_level ? 2 : (cfg.traverse_ifif ? 1 : 0);
    if (cfg.sorted_rays) { if (trav == 3) BTRC_TRACE(3, true); else if (trav == 2) BTRC_TRACE(2, true); else if (trav == 1) BTRC_TRACE(1, true); else BTRC_TRACE(0, true); }
    else { if (trav == 3) BTRC_TRACE(3, false); else if (trav == 2) BTRC_TRACE(2, false); else if (trav == 1) BTRC_TRACE(1, false); else BTRC_TRACE(0, false); }
#undef BTRC_TRACE
}

template <int MAT>
static void launch_shade_class(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, HitQueue hit, PathQueue next,
                               ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    const int blocks = persistent_blocks(cfg, 8 * kBlock / shade_block<MAT>());
    if (params.halton) shade_kernel<MAT, true><<<blocks, shade_block<MAT>(), 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
    else shade_kernel<MAT, false><<<blocks, shade_block<MAT>(), 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
}

int launch_shade(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, uint32_t material_mask, PathQueue cur, HitQueue hit,
                 PathQueue next, ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    int launches = 1;
    shade_miss_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, cur, hit, film, ctl);
#define BTRC_SHADE_CLASS(M) \
    if (material_mask & (1u << M)) { launch_shade_class<M>(cfg, scene, params, cur, hit, next, shadow, film, ctl, stream); ++launches; }
    BTRC_SHADE_CLASS(MAT_DISNEY)
    BTRC_SHADE_CLASS(MAT_DIFFUSE)
    BTRC_SHADE_CLASS(MAT_METAL)
    BTRC_SHADE_CLASS(MAT_GLASS)
    BTRC_SHADE_CLASS(MAT_MIRROR)
    BTRC_SHADE_CLASS(MAT_INVISIBLE)
    BTRC_SHADE_CLASS(MAT_BLACK)
#undef BTRC_SHADE_CLASS
    return launches;
}

void launch_ibl_lut(const SceneDevice &scene, int tex, int lut_w, int lut_h, int n_samples, float *lum_table, float *area_table, cudaStream_t stream) {
    const dim3 block(8, 8), grid((lut_w + 7) / 8, (lut_h + 7) / 8);
    ibl_lut_kernel<<<grid, block, 0, stream>>>(scene, tex, lut_w, lut_h, n_samples, lum_table, area_table);
}

void launch_classify(const LaunchCfg &cfg, const SceneDevice &scene, HitQueue hit, WaveCtl *ctl, cudaStream_t stream) {
    classify_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, hit, ctl);
}

void launch_medium(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, HitQueue hit, PathQueue next,
                   ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream) {
    const int blocks = persistent_blocks(cfg, 8);
    if (scene.has_medium == 2) {  // the scene has "add" / "mul" 3-D textures (postfix programs): instantiations that can run them
        if (params.halton) medium_kernel<true, true><<<blocks, kBlock, 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
        else medium_kernel<false, true><<<blocks, kBlock, 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
        return;
    }
    if (params.halton) medium_kernel<true, false><<<blocks, kBlock, 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
    else medium_kernel<false, false><<<blocks, kBlock, 0, stream>>>(scene, params, cur, hit, next, shadow, film, ctl);
}

void launch_shadow(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, ShadowQueue shadow, RaySort rs, FilmDevice film, WaveCtl *ctl,
                   cudaStream_t stream) {
    const int blocks = persistent_blocks(cfg, 8);
#define BTRC_SHADOW(T, S, P) shadow_kernel<T, S, P><<<blocks, kBlock, 0, stream>>>(scene, params, shadow, rs, film, ctl)
#define BTRC_SHADOW_T(S, P) do { if (trav == 3) BTRC_SHADOW(3, S, P); else if (trav == 2) BTRC_SHADOW(2, S, P); else if (trav == 1) BTRC_SHADOW(1, S, P); else BTRC_SHADOW(0, S, P); } while (0)
    const int trav = scene.bvh.wide ? 3 : scene.bvh.two_level ? 2 : (cfg.traverse_ifif ? 1 : 0);
    if (scene.has_medium == 2) {  // 3-D texture programs in the transmittance: see launch_medium
        if (cfg.sorted_rays) BTRC_SHADOW_T(true, true); else BTRC_SHADOW_T(false, true);
    } else {
        if (cfg.sorted_rays) BTRC_SHADOW_T(true, false); else BTRC_SHADOW_T(false, false);
    }
#undef BTRC_SHADOW_T
#undef BTRC_SHADOW
}

void launch_resolve(int width, int height, const float *value, const float *weight, float *out, int normalise, cudaStream_t stream) {
    const int npix = width * height;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(npix, reinterpret_cast<const float4 *>(value), weight,
                                                          reinterpret_cast<float4 *>(out), normalise);
}

void launch_trace_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n, uint32_t *flag, uint4 *detail,
                        cudaStream_t stream) {
    trace_batch_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, ray_o, ray_d, n, flag, detail);
}

void launch_anyhit_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n, uint8_t *occluded,
                         cudaStream_t stream) {
    anyhit_batch_kernel<<<persistent_blocks(cfg, 8), kBlock, 0, stream>>>(scene, ray_o, ray_d, n, occluded);
}

}  // namespace btrc
