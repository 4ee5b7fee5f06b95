// Wavefront path tracer host loop ("wfpt" renderer).
//
// Replaces builtin/renderer/wavefront.cpp:98-194 (commit: film + queue allocation) and :205-325 (render loop).
// Differences by design (results are unaffected, SURVEY B.1: every path owns its RNG stream):
//   * the pool is refilled every iteration and all bookkeeping lives in a device-resident WaveCtl block, so the
//     host enqueues iterations ahead of the GPU and never blocks on a per-iteration counter read-back;
//   * a render can cover any contiguous range of camera-path indices, which is how sample streams are
//     partitioned across GPUs (the caller sum-reduces the film buffers and resolves once).
#pragma once
#include <cstdint>
#include <functional>

#include "../device/kernels.h"
#include "scene.h"

namespace btrc {

struct RenderStats {
    uint64_t n_gen = 0, n_closest = 0, n_hit = 0, n_next = 0, n_shadow = 0, n_shadow_miss = 0, n_splat = 0, n_iter = 0;
    double render_ms = 0;  // device time of the last render_range (CUDA events)
    // per-stage device time of the last render_range (only when profiling is on): generate, trace, shade, shadow, medium,
    // iteration bookkeeping (counter snapshot + prepare kernel)
    double stage_ms[6] = {0, 0, 0, 0, 0, 0};
    uint64_t n_launches = 0;  // kernels launched by the last render_range
};

class Renderer {
public:
    explicit Renderer(Scene &scene, int state_count_override = 0);
    ~Renderer();

    void clear_film();
    void set_profiling(bool on) { profiling_ = on; }
    // progress(percentage) -> true to stop (core/reporter.h + Renderer::should_stop); called from the rendering thread
    void set_reporter(std::function<bool(float)> fn) { reporter_ = std::move(fn); }
    bool stopped() const { return stopped_; }
    // Traces camera paths [begin, end) of the scene's path index space (Uniform mode: index = sample * W*H + pixel;
    // Tile mode: index = pixel * spp + sample) and accumulates them into the film.
    // renders camera paths [begin, end) of the path index space; with stride_block > 0 the range is VIRTUAL (types.h WaveCtl:
    // interleaved blocks of `stride_block` paths, this renderer taking block `stride_part` of every `stride_parts`)
    void render_range(long long begin, long long end, long long stride_block = 0, int stride_parts = 1, int stride_part = 0, int stride_bw = 0, int stride_rows = 0);
    // this renderer's share of the whole render when `parts` devices split it (partition chosen by the renderer's "tile" flag:
    // sample-index ranges, or interleaved strips of `rows_per_strip` image rows)
    void render_part(int parts, int part, int rows_per_strip);
    FilmDevice &film() { return film_; }
    // value / weight into caller-provided DEVICE buffers of W*H float4 (albedo/normal may be null).
    void resolve(float *d_color, float *d_albedo, float *d_normal);
    // The same into buffers in ANY address space: device destinations are written by the resolve kernel itself, host
    // destinations through a scratch image owned by the renderer (allocated once) and an async copy on the renderer's
    // stream -- no per-call cudaMalloc/cudaFree and nothing on the legacy default stream.
    void resolve_any(float *color, float *albedo, float *normal);
    // Film buffers to / from caller memory in any address space (multi-GPU accumulation): async copies on the renderer's
    // stream, one device-wide fence first (the caller's buffers may be in use on torch / NCCL streams).
    void export_film(float *radiance, float *weight, float *albedo, float *normal);
    void import_film(const float *radiance, const float *weight, const float *albedo, const float *normal);

    // ---- stage-level runs on the internal queues (tests / C-ABI stage entry points); n <= state_count
    PathQueue &queue(int k) { return q_[k]; }
    HitQueue &hits() { return hit_; }
    ShadowQueue &shadows() { return shadow_; }
    void stage_generate(long long first, int n);                  // fills queue(0)[0, n)
    void stage_trace(int n);                                      // queue(0) -> hits()
    void stage_shade(int n, int *n_next, int *n_shadow);          // queue(0) + hits() -> queue(1), shadows()
    void stage_medium(int n, int *n_next, int *n_shadow);         // same, medium stage (updates queue(0) / hits() in place)
    void stage_shadow(int n);                                     // shadows()[0, n) -> film
    const LaunchCfg &launch_cfg() const { return cfg_; }

    FilmDevice film() const { return film_; }
    const RenderStats &stats() const { return stats_; }
    cudaStream_t stream() const { return stream_; }
    int state_count() const { return pool_; }
    long long total_paths() const { return (long long)scene_.renderer.spp * scene_.width * scene_.height; }

private:
    Scene &scene_;
    int pool_ = 0;
    LaunchCfg cfg_;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_begin_ = nullptr, ev_end_ = nullptr;
    std::vector<DeviceBuffer> bufs_;
    PathQueue q_[2]{};
    HitQueue hit_{};
    ShadowQueue shadow_{};
    RaySort sort_path_{}, sort_shadow_{};  // binned visiting order of surviving paths / shadow rays (types.h)
    int sort_rays_ = 1;  // 0: off; 1: sorted traversal, shade lists in traversal order; 2: sorted traversal, shade lists in state order
    FilmDevice film_{};
    uint32_t material_mask_ = 0;    // bit m: a material of MaterialType m exists (one shade kernel per set bit)
    WaveCtl *ctl_ = nullptr;        // device
    WaveCtl *ctl_host_ = nullptr;   // pinned ring of snapshots
    cudaEvent_t snap_ev_[8]{};
    RenderStats stats_;
    bool profiling_ = false;
    std::function<bool(float)> reporter_;
    bool stopped_ = false;
    std::vector<cudaEvent_t> prof_events_;
    float *resolve_scratch_[3]{};   // lazily allocated W*H float4 images for host destinations
    template <typename T> T *alloc(size_t n);
};

}  // namespace btrc
