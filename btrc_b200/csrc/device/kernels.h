// Launch wrappers for the wavefront stages (definitions in the .cu files of this directory).
// One stage of the reference = one kernel here:
//   generate_kernel (wavefront/generate.cpp:33-87)            -> launch_generate
//   __raygen__/__closesthit__/__miss__trace (trace.cpp:38-78)  -> launch_trace
//   shade_kernel (shade.cpp:29-235)                            -> launch_shade
//   __raygen__/__miss__shadow (shadow.cpp:24-76)               -> launch_shadow
//   preview kernels (preview.cpp:28-80)                        -> launch_resolve
// plus launch_prepare, the device-side replacement of the host bookkeeping in wavefront.cpp:227-306.
#pragma once
#include <cuda_runtime.h>

#include "../common/types.h"

namespace btrc {

struct LaunchCfg {
    int sm_count = 148;
    // Loop structure of the traversal kernels (traverse.cuh).  While-while is faster while the BVH lives in L2 (C2: 66.5 vs
    // 67.8 ms), if-if when node fetches go to DRAM and the extra loads in flight per warp pay (C4, 0.7 GB of BVH: 136 vs 153 ms);
    // profiles/r1h_ab_ifif.log.  The renderer picks by BVH size against the L2 capacity.
    bool traverse_ifif = false;
    int probe_top_levels = 0;  // profiling only (BTRC_PROBE_TOP_LEVELS): also launch top_levels_probe_kernel on every wave
    bool sorted_rays = false;  // the traversal kernels visit rays through RaySort::order (types.h); set with the renderer's sort mode
};

// Device-side LBVH build (lbvh.cu): tris_in / ids_in are device arrays in input order; tris_out / ids_out (caller allocated) receive
// the sorted triangles, *nodes_out (cudaMalloc'ed here, 64 bytes per node) the hierarchy.  Returns the root reference.
int build_lbvh_device(int n, const float4 *tris_in, const uint2 *ids_in, const float scene_lo[3], const float scene_hi[3], float coord_max,
                      float4 *tris_out, uint2 *ids_out, float4 **nodes_out, int *n_nodes_out, cudaStream_t stream);

void launch_prepare(WaveCtl *ctl, cudaStream_t stream);
void launch_generate(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, FilmDevice film,
                     WaveCtl *ctl, cudaStream_t stream);
void launch_weight_fill(const LaunchCfg &cfg, const RenderParams &params, FilmDevice film, long long begin, long long end, long long stride_block,
                        int stride_parts, int stride_part, int stride_bw, int stride_rows, cudaStream_t stream);
void launch_ray_sort(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, uint32_t *key, const int *count, RaySort rs,
                     int *n_sorted_out, cudaStream_t stream);
void launch_trace(const LaunchCfg &cfg, const SceneDevice &scene, PathQueue cur, HitQueue hit, RaySort rs, WaveCtl *ctl, cudaStream_t stream);
// One kernel per shade class: misses + one instantiation per material type present in the scene (bit m of material_mask
// = a material of MaterialType m exists).  Returns the number of kernels launched.
int launch_shade(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, uint32_t material_mask, PathQueue cur,
                 HitQueue hit, PathQueue next, ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream);
// Builds the shade class lists from hit flags (stage-level entry points; in a render trace_kernel builds them)
void launch_classify(const LaunchCfg &cfg, const SceneDevice &scene, HitQueue hit, WaveCtl *ctl, cudaStream_t stream);
// sample_medium_scattering_kernel (medium.cpp:28-164): runs between trace and shade when the scene has media
void launch_medium(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, PathQueue cur, HitQueue hit,
                   PathQueue next, ShadowQueue shadow, FilmDevice film, WaveCtl *ctl, cudaStream_t stream);
void launch_shadow(const LaunchCfg &cfg, const SceneDevice &scene, const RenderParams &params, ShadowQueue shadow, RaySort rs, FilmDevice film,
                   WaveCtl *ctl, cudaStream_t stream);
// out = value.xyz / weight (normalise == 1: renormalise as a direction); out.w = 1
void launch_resolve(int width, int height, const float *value, const float *weight, float *out, int normalise, cudaStream_t stream);

// IBL preprocess (env_sampler.cpp:22-57): per-tile mean luminance and solid angle into two lut_w*lut_h device tables
void launch_ibl_lut(const SceneDevice &scene, int tex, int lut_w, int lut_h, int n_samples, float *lum_table, float *area_table, cudaStream_t stream);

// Stage-level entry points for tests on explicit batches (count given by the host; no WaveCtl).
void launch_trace_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n,
                        uint32_t *flag, uint4 *detail, cudaStream_t stream);
void launch_anyhit_batch(const LaunchCfg &cfg, const SceneDevice &scene, const float4 *ray_o, const float4 *ray_d, int n,
                         uint8_t *occluded, cudaStream_t stream);

}  // namespace btrc
