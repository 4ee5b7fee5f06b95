#include "renderer.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace btrc {

namespace {
constexpr int kLag = 4;  // iterations kept in flight ahead of the host's view of the counters

// Diagnostics (BTRC_HOST_TRACE=1): reports host-side spans of the render calls that took longer than 5 ms
struct HostSpan {
    const char *name;
    std::chrono::steady_clock::time_point t0;
    static bool enabled() {
        static const bool on = std::getenv("BTRC_HOST_TRACE") != nullptr;
        return on;
    }
    explicit HostSpan(const char *n) : name(n) { if (enabled()) t0 = std::chrono::steady_clock::now(); }
    ~HostSpan() {
        if (!enabled()) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 5.0) std::fprintf(stderr, "[btrc host trace] %s: %.1f ms\n", name, ms);
    }
};
}

template <typename T>
T *Renderer::alloc(size_t n) {
    bufs_.emplace_back(n * sizeof(T));
    return bufs_.back().as<T>();
}

Renderer::Renderer(Scene &scene, int state_count_override) : scene_(scene) {
    int dev = 0;
    BTRC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    BTRC_CUDA(cudaGetDeviceProperties(&prop, dev));
    cfg_.sm_count = prop.multiProcessorCount;
    {   // traversal loop form by BVH footprint against the L2 (kernels.h: LaunchCfg::traverse_ifif); BTRC_TRAVERSE=ww|ifif overrides
        const size_t bvh_bytes = scene.renderer.lbvh ? 64 * (size_t)scene.lbvh_nodes + 48 * scene.world_tris.size()
                                 : scene.wide_bvh8 ? (scene.wide_bvh.nodes.size() + scene.wide_bvh.tris.size()) * sizeof(float)
                                                     : scene.flat_bvh.nodes.size() * sizeof(float) + scene.flat_bvh.tris.size() * sizeof(float);
        cfg_.traverse_ifif = bvh_bytes > (size_t)prop.l2CacheSize;
        if (const char *e = std::getenv("BTRC_TRAVERSE")) cfg_.traverse_ifif = std::string(e) == "ifif";
        // Binned ray order (types.h: RaySort) under the same criterion: on C4 (0.7 GB of BVH) visiting surviving paths and shadow rays
        // bin by bin makes traversal 9-13 % faster for 2 ms of sorting per render (-7.7 % overall); while the BVH sits in L2 the gain
        // does not pay for the sort and the gather (C2 +4 %, C3 +11 %, C1 +17 %; profiles/r1j_ab_ray_sort.log).  BTRC_SORT_RAYS=0|1|2 overrides.
        sort_rays_ = bvh_bytes > (size_t)prop.l2CacheSize ? 2 : 0;
    }
    if (const char *e = std::getenv("BTRC_PROBE_TOP_LEVELS")) cfg_.probe_top_levels = std::atoi(e);
    pool_ = state_count_override > 0 ? state_count_override : scene.renderer.state_count;
    if (pool_ <= 0) throw Error("state_count must be positive");
    BTRC_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    BTRC_CUDA(cudaEventCreate(&ev_begin_));
    BTRC_CUDA(cudaEventCreate(&ev_end_));
    for (auto &e : snap_ev_) BTRC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    const size_t P = (size_t)pool_;
    for (int k = 0; k < 2; ++k) {
        q_[k].ray_o = alloc<float4>(P);
        q_[k].ray_d = alloc<float4>(P);
        q_[k].beta = alloc<float4>(P);
        q_[k].rad = alloc<float4>(P);
        q_[k].rng = alloc<ulonglong2>(P);
        q_[k].ble = alloc<float4>(P);
        q_[k].key = alloc<uint32_t>(P);
    }
    hit_.flag = alloc<uint32_t>(P);
    hit_.detail = alloc<uint4>(P);
    hit_.class_list = alloc<uint32_t>(P * kNumShadeClasses);
    hit_.class_stride = pool_;
    for (const MaterialDesc &m : scene.materials) material_mask_ |= 1u << m.type;
    shadow_.ray_o = alloc<float4>(P);
    shadow_.ray_d = alloc<float4>(P);
    shadow_.beta_li = alloc<float4>(P);
    shadow_.key = alloc<uint32_t>(P);
    sort_path_.order = alloc<uint32_t>(P); sort_path_.bin = alloc<int>(kSortBins + 1);
    sort_shadow_.order = alloc<uint32_t>(P); sort_shadow_.bin = alloc<int>(kSortBins + 1);
    if (const char *e = std::getenv("BTRC_SORT_RAYS")) sort_rays_ = std::atoi(e);
    cfg_.sorted_rays = sort_rays_ != 0;

    const size_t npix = (size_t)scene.width * scene.height;
    film_.radiance = alloc<float>(4 * npix);
    film_.weight = alloc<float>(npix);
    film_.albedo = scene.renderer.albedo ? alloc<float>(4 * npix) : nullptr;
    film_.normal = scene.renderer.normal ? alloc<float>(4 * npix) : nullptr;
    ctl_ = alloc<WaveCtl>(1);
    BTRC_CUDA(cudaMallocHost(&ctl_host_, sizeof(WaveCtl) * 10));  // 8 snapshot slots + upload staging + final counters
    clear_film();
}

Renderer::~Renderer() {
    if (ctl_host_) cudaFreeHost(ctl_host_);
    for (auto &e : snap_ev_) if (e) cudaEventDestroy(e);
    for (auto &e : prof_events_) cudaEventDestroy(e);
    if (ev_begin_) cudaEventDestroy(ev_begin_);
    if (ev_end_) cudaEventDestroy(ev_end_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Renderer::clear_film() {
    HostSpan span("clear_film");
    const size_t npix = (size_t)scene_.width * scene_.height;
    BTRC_CUDA(cudaMemsetAsync(film_.radiance, 0, 4 * npix * sizeof(float), stream_));
    BTRC_CUDA(cudaMemsetAsync(film_.weight, 0, npix * sizeof(float), stream_));
    if (film_.albedo) BTRC_CUDA(cudaMemsetAsync(film_.albedo, 0, 4 * npix * sizeof(float), stream_));
    if (film_.normal) BTRC_CUDA(cudaMemsetAsync(film_.normal, 0, 4 * npix * sizeof(float), stream_));
    stats_ = RenderStats();
}

void Renderer::render_part(int parts, int part, int rows_per_strip) {
    if (parts < 1 || part < 0 || part >= parts) throw Error("render_part: part out of range");
    const long long W = scene_.width, H = scene_.height, spp = scene_.renderer.spp;
    if (!scene_.renderer.tile) {  // sample streams: sample indices [part * spp / parts, ...) of every pixel (partition A)
        const long long base = spp / parts, rem = spp % parts;
        const long long b = part * base + std::min<long long>(part, rem), e = b + base + (part < rem ? 1 : 0);
        render_range(b * W * H, e * W * H);
        return;
    }
    // image tiles (partition B), dealt round-robin so that cheap and expensive image regions are spread over all devices.  When the
    // image divides into 64 (128, 256, 32) x rows tiles the split is 2-D (compact tiles: the rays of a wave start in one image region, like the
    // generation order of uniform-mode renders); otherwise strips of `rows` whole image rows, which tile-mode indexing makes
    // contiguous blocks of the index space.  rows: at most 16, at least 8 strips per device, a divisor of the height if one is near.
    if (rows_per_strip <= 0) {
        rows_per_strip = (int)std::max<long long>(1, std::min<long long>(16, H / (8LL * parts)));
        for (int r = rows_per_strip; r >= std::max(1, rows_per_strip / 2); --r)
            if (H % r == 0) { rows_per_strip = r; break; }
    }
    const long long rows = rows_per_strip;
    // tile width: 64 pixels, or the first of 128 / 256 / 32 that makes the number of tiles per row coprime to the number of parts --
    // then "tile k belongs to part k % parts" walks diagonally over the image instead of giving every part fixed columns
    long long kTileW = 64;
    {
        auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
        for (long long tw : {64LL, 128LL, 256LL, 32LL})
            if (W % tw == 0 && gcd(W / tw, parts) == 1) { kTileW = tw; break; }
        if (const char *e = std::getenv("BTRC_TILE_W")) if (std::atoi(e) > 0) kTileW = std::atoi(e);
    }
    static const bool tiles_2d = !(std::getenv("BTRC_TILES_2D") && std::atoi(std::getenv("BTRC_TILES_2D")) == 0);
    if (tiles_2d && W % kTileW == 0 && H % rows == 0 && kTileW * rows * spp < (1LL << 31)) {  // the kernel's in-tile arithmetic is 32-bit
        const long long block = kTileW * rows * spp, tiles = (W / kTileW) * (H / rows);
        const long long owned = part < tiles ? (tiles - 1 - part) / parts + 1 : 0;
        render_range(0, owned * block, block, parts, part, (int)kTileW, (int)rows);
        return;
    }
    const long long block = rows * W * spp, total = W * H * spp;
    const long long strips = (H + rows - 1) / rows;
    const long long owned = part < strips ? (strips - 1 - part) / parts + 1 : 0;
    if (owned == 0) { render_range(0, 0); return; }
    const bool owns_last = (strips - 1) % parts == part;
    const long long last_paths = total - (strips - 1) * block;
    const long long n_virtual = owns_last ? (owned - 1) * block + last_paths : owned * block;
    render_range(0, n_virtual, block, parts, part);
}

void Renderer::render_range(long long begin, long long end, long long stride_block, int stride_parts, int stride_part, int stride_bw, int stride_rows) {
    if (stride_bw > 0 && (stride_rows <= 0 || scene_.width % stride_bw || scene_.height % stride_rows || stride_block >= (1LL << 31) ||
                          stride_block != (long long)stride_bw * stride_rows * scene_.renderer.spp))
        throw Error("render_range: 2-D tiles must divide the image");
    // inside a tile the paths are walked warp by warp (path_index.cuh), so only whole tiles have a closed-form weight per pixel
    if (stride_bw > 0 && (begin % stride_block != 0 || end % stride_block != 0)) throw Error("render_range: a 2-D tile partition renders whole tiles");
    if (stride_block < 0 || stride_parts < 1 || stride_part < 0 || stride_part >= stride_parts) throw Error("render_range: bad stride");
    if (stride_block > 0 && (!scene_.renderer.tile || stride_block % scene_.renderer.spp != 0))
        throw Error("render_range: a strided partition needs tile-mode indexing and whole pixels per block");
    if (begin < 0 || end < begin || end > total_paths()) throw Error("render_range: path range out of bounds");
    HostSpan span_all("render_range");
    const RenderParams params = scene_.render_params();
    WaveCtl init;
    std::memset(&init, 0, sizeof(init));
    init.cursor = begin;
    init.end = end;
    init.pool = pool_;
    init.sort_enabled = sort_rays_;
    init.stride_block = stride_block;
    init.stride_parts = stride_parts;
    init.stride_part = stride_part;
    init.stride_bw = stride_bw;
    init.stride_rows = stride_rows;
    init.weight_prefilled = 1;
    {   // block-coherent generation order for ranges made of whole sample planes (types.h WaveCtl::order_ns); BTRC_GEN_ORDER=0 turns it off
        static const bool coherent = !(std::getenv("BTRC_GEN_ORDER") && std::atoi(std::getenv("BTRC_GEN_ORDER")) == 0);
        const long long npix = (long long)scene_.width * scene_.height;
        if (coherent && !scene_.renderer.tile && stride_block == 0 && end > begin && begin % npix == 0 && end % npix == 0) {
            init.order_s0 = (int)(begin / npix);
            init.order_ns = (int)((end - begin) / npix);
            // profiles/r2i_sweep_gen_order.md, r2r_ab_chunked_order.md: 64 x 32 blocks with all samples of a pixel in a row; when the
            // sample count is a multiple of the warp size, 32 x 32 blocks walked warp by warp (32 samples of a pixel, next pixel, ...)
            init.order_bw = 64; init.order_bh = 32; init.order_sample_major = 0;
            if (init.order_ns % 32 == 0) { init.order_bw = 32; init.order_sample_major = 2; }
            if (const char *e = std::getenv("BTRC_GEN_BLOCK")) {  // tuning: "<width>x<height>[s|c]" (s: sample planes inside a block, c: 32-sample chunks)
                int bw = 0, bh = 0;
                char mode = 0;
                if (std::sscanf(e, "%dx%d%c", &bw, &bh, &mode) >= 2 && bw > 0 && bh > 0) {
                    init.order_bw = bw; init.order_bh = bh;
                    init.order_sample_major = mode == 's' ? 1 : (mode == 'c' && init.order_ns % 32 == 0) ? 2 : 0;  // c: 32-sample chunks
                }
            }
            // the kernel divides inside a row of blocks in 32 bits; a row that large (16 K-wide image at 4096 spp) keeps index order
            if ((long long)scene_.width * init.order_bh * init.order_ns >= (1LL << 31)) init.order_ns = 0;
            else { init.cursor = 0; init.end = end - begin; }
        }
    }
    ctl_host_[8] = init;  // pinned staging slot: the upload is a true async copy on the renderer's stream
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &ctl_host_[8], sizeof(init), cudaMemcpyHostToDevice, stream_));
    BTRC_CUDA(cudaEventRecord(ev_begin_, stream_));
    launch_weight_fill(cfg_, params, film_, begin, end, stride_block, stride_parts, stride_part, stride_bw, stride_rows, stream_);

    const bool has_medium = scene_.dev.has_medium != 0;
    const int kernels_per_iter = has_medium ? 5 : 4;  // prepare, generate, trace, [medium], shadow; shade adds its own count
    unsigned long long n_kernels = 0;
    int cur = 0;
    long long iter = 0;
    bool done = false;
    WaveCtl last = init;
    size_t n_prof = 0;
    auto mark = [&]() {  // profiling: one event between consecutive kernels
        if (!profiling_) return;
        if (n_prof == prof_events_.size()) {
            cudaEvent_t e;
            BTRC_CUDA(cudaEventCreate(&e));
            prof_events_.push_back(e);
        }
        BTRC_CUDA(cudaEventRecord(prof_events_[n_prof++], stream_));
    };
    // Diagnostics (BTRC_ITER_LOG=1): work counts of every launched iteration, in launch order across renders, so that a
    // profiler capture of "the k-th trace_kernel launch" can be matched with the number of rays it processed
    static const bool iter_log = std::getenv("BTRC_ITER_LOG") != nullptr;
    static long long global_iter = 0;
    long long logged = 0;
    auto log_iter = [&](long long idx) {
        const WaveCtl &c = ctl_host_[idx % 8];
        std::fprintf(stderr, "[btrc iter] %lld n_active=%d n_new=%d n_next=%d n_shadow=%d classes=%d,%d,%d,%d,%d,%d,%d,%d\n", global_iter++, c.n_active,
                     c.n_new, c.n_next, c.n_shadow, c.class_count[0], c.class_count[1], c.class_count[2], c.class_count[3], c.class_count[4],
                     c.class_count[5], c.class_count[6], c.class_count[7]);
    };
    stopped_ = false;
    float reported = 0.0f;
    if (reporter_ && reporter_(0.0f)) stopped_ = true;  // Reporter::new_stage
    while (!done && !stopped_) {
        launch_prepare(ctl_, stream_);
        mark();
        if (sort_rays_) launch_ray_sort(cfg_, scene_.dev, q_[cur].ray_o, q_[cur].ray_d, q_[cur].key, &ctl_->n_survivors, sort_path_, nullptr, stream_);
        launch_generate(cfg_, scene_.dev, params, q_[cur], film_, ctl_, stream_);
        mark();
        launch_trace(cfg_, scene_.dev, q_[cur], hit_, sort_path_, ctl_, stream_);
        if (sort_rays_ == 2) launch_classify(cfg_, scene_.dev, hit_, ctl_, stream_);
        mark();
        if (has_medium) launch_medium(cfg_, scene_.dev, params, q_[cur], hit_, q_[cur ^ 1], shadow_, film_, ctl_, stream_);
        mark();
        n_kernels += launch_shade(cfg_, scene_.dev, params, material_mask_, q_[cur], hit_, q_[cur ^ 1], shadow_, film_, ctl_, stream_);
        mark();
        if (sort_rays_) launch_ray_sort(cfg_, scene_.dev, shadow_.ray_o, shadow_.ray_d, shadow_.key, &ctl_->n_shadow, sort_shadow_, &ctl_->n_shadow_sorted, stream_);
        launch_shadow(cfg_, scene_.dev, params, shadow_, sort_shadow_, film_, ctl_, stream_);
        mark();
        const int slot = (int)(iter % 8);
        BTRC_CUDA(cudaMemcpyAsync(&ctl_host_[slot], ctl_, sizeof(WaveCtl), cudaMemcpyDeviceToHost, stream_));
        BTRC_CUDA(cudaEventRecord(snap_ev_[slot], stream_));
        cur ^= 1;
        ++iter;
        if (iter >= kLag) {
            const int old = (int)((iter - kLag) % 8);
            BTRC_CUDA(cudaEventSynchronize(snap_ev_[old]));
            last = ctl_host_[old];
            if (iter_log) log_iter(logged++);
            // the snapshot is taken after shade: the wave is finished when nothing survived and nothing is left to generate
            if (last.n_next == 0 && last.cursor >= last.end) done = true;
            if (reporter_ && !done) {  // wavefront.cpp:286-287: finished = generated - still active; reported in steps of >= 1 % (console.cpp:26-33)
                const float pct = 100.0f * float(double(last.n_gen - (unsigned long long)last.n_next) / double(std::max<long long>(end - begin, 1)));
                if (pct - reported >= 1.0f) {
                    reported = pct;
                    if (reporter_(pct)) stopped_ = true;  // should_stop(): the waves in flight finish, nothing new is enqueued
                }
            }
        }
    }
    BTRC_CUDA(cudaEventRecord(ev_end_, stream_));
    HostSpan span_tail("render_range tail (final sync + counters)");
    // final counters (the drained iterations after `done` are empty)
    BTRC_CUDA(cudaMemcpyAsync(&ctl_host_[9], ctl_, sizeof(WaveCtl), cudaMemcpyDeviceToHost, stream_));
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
    const WaveCtl fin = ctl_host_[9];
    if (reporter_ && !stopped_) reporter_(100.0f);  // Reporter::complete_stage
    if (iter_log) while (logged < iter) log_iter(logged++);
    float ms = 0;
    BTRC_CUDA(cudaEventElapsedTime(&ms, ev_begin_, ev_end_));
    stats_.n_gen += fin.n_gen;
    stats_.n_closest += fin.n_closest;
    stats_.n_hit += fin.n_hit;
    stats_.n_next += fin.n_next_total + (unsigned long long)fin.n_next;
    stats_.n_shadow += fin.n_shadow_total + (unsigned long long)fin.n_shadow;
    stats_.n_shadow_miss += fin.n_shadow_miss;
    stats_.n_splat += fin.n_splat;
    stats_.n_iter += fin.n_iter;
    stats_.render_ms = ms;
    stats_.n_launches = (unsigned long long)kernels_per_iter * (unsigned long long)iter + n_kernels + 1;  // + weight_fill_kernel
    for (double &v : stats_.stage_ms) v = 0;
    // events per iteration: after prepare, generate, trace, medium, shade, shadow  => intervals gen, trace, medium, shade, shadow
    // the sixth interval (after shadow -> after the next prepare) is the iteration bookkeeping: counter snapshot + prepare kernel
    static const int kStageOfInterval[6] = {0, 1, 4, 2, 3, 5};
    for (size_t k = 0; k + 5 < n_prof; k += 6) {
        for (int iv = 0; iv < 6; ++iv) {
            if (k + iv + 1 >= n_prof) break;
            float t = 0;
            BTRC_CUDA(cudaEventElapsedTime(&t, prof_events_[k + iv], prof_events_[k + iv + 1]));
            stats_.stage_ms[kStageOfInterval[iv]] += t;
        }
    }
}

void Renderer::stage_generate(long long first, int n) {
    if (n > pool_) throw Error("stage batch larger than state_count");
    WaveCtl c;
    std::memset(&c, 0, sizeof(c));
    c.n_new = n; c.n_survivors = 0; c.gen_base = first; c.n_active = n; c.pool = pool_;
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, stream_));
    launch_generate(cfg_, scene_.dev, scene_.render_params(), q_[0], film_, ctl_, stream_);
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
}

void Renderer::stage_trace(int n) {
    if (n > pool_) throw Error("stage batch larger than state_count");
    WaveCtl c;
    std::memset(&c, 0, sizeof(c));
    c.n_active = n; c.pool = pool_;
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, stream_));
    launch_trace(cfg_, scene_.dev, q_[0], hit_, sort_path_, ctl_, stream_);
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
}

void Renderer::stage_shade(int n, int *n_next, int *n_shadow) {
    if (n > pool_) throw Error("stage batch larger than state_count");
    WaveCtl c;
    std::memset(&c, 0, sizeof(c));
    c.n_active = n; c.pool = pool_;
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, stream_));
    launch_classify(cfg_, scene_.dev, hit_, ctl_, stream_);
    launch_shade(cfg_, scene_.dev, scene_.render_params(), material_mask_, q_[0], hit_, q_[1], shadow_, film_, ctl_, stream_);
    BTRC_CUDA(cudaMemcpyAsync(&c, ctl_, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
    *n_next = c.n_next;
    *n_shadow = c.n_shadow;
}

void Renderer::stage_medium(int n, int *n_next, int *n_shadow) {
    if (n > pool_) throw Error("stage batch larger than state_count");
    WaveCtl c;
    std::memset(&c, 0, sizeof(c));
    c.n_active = n; c.pool = pool_;
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, stream_));
    launch_medium(cfg_, scene_.dev, scene_.render_params(), q_[0], hit_, q_[1], shadow_, film_, ctl_, stream_);
    BTRC_CUDA(cudaMemcpyAsync(&c, ctl_, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
    *n_next = c.n_next;
    *n_shadow = c.n_shadow;
}

void Renderer::stage_shadow(int n) {
    if (n > pool_) throw Error("stage batch larger than state_count");
    WaveCtl c;
    std::memset(&c, 0, sizeof(c));
    c.n_shadow = n; c.pool = pool_;
    BTRC_CUDA(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, stream_));
    launch_shadow(cfg_, scene_.dev, scene_.render_params(), shadow_, sort_shadow_, film_, ctl_, stream_);
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
}

void Renderer::resolve(float *d_color, float *d_albedo, float *d_normal) {
    if (d_color) launch_resolve(scene_.width, scene_.height, film_.radiance, film_.weight, d_color, 0, stream_);
    if (d_albedo && film_.albedo) launch_resolve(scene_.width, scene_.height, film_.albedo, film_.weight, d_albedo, 0, stream_);
    if (d_normal && film_.normal) launch_resolve(scene_.width, scene_.height, film_.normal, film_.weight, d_normal, 1, stream_);
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
}

static bool is_device_pointer(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

void Renderer::resolve_any(float *color, float *albedo, float *normal) {
    HostSpan span("resolve_any");
    const size_t bytes = sizeof(float) * 4 * (size_t)scene_.width * scene_.height;
    float *dst[3] = {color, film_.albedo ? albedo : nullptr, film_.normal ? normal : nullptr};
    const float *src[3] = {film_.radiance, film_.albedo, film_.normal};
    for (int k = 0; k < 3; ++k) {
        if (!dst[k]) continue;
        float *target = dst[k];
        const bool direct = is_device_pointer(target);
        if (!direct) {
            if (!resolve_scratch_[k]) resolve_scratch_[k] = alloc<float>(bytes / sizeof(float));
            target = resolve_scratch_[k];
        }
        launch_resolve(scene_.width, scene_.height, src[k], film_.weight, target, k == 2 ? 1 : 0, stream_);
        if (!direct) BTRC_CUDA(cudaMemcpyAsync(dst[k], target, bytes, cudaMemcpyDeviceToHost, stream_));
    }
    BTRC_CUDA(cudaStreamSynchronize(stream_));
    BTRC_CUDA(cudaGetLastError());
}

void Renderer::export_film(float *radiance, float *weight, float *albedo, float *normal) {
    HostSpan span("export_film");
    BTRC_CUDA(cudaDeviceSynchronize());
    const size_t npix = (size_t)scene_.width * scene_.height;
    if (radiance) BTRC_CUDA(cudaMemcpyAsync(radiance, film_.radiance, 16 * npix, cudaMemcpyDefault, stream_));
    if (weight) BTRC_CUDA(cudaMemcpyAsync(weight, film_.weight, 4 * npix, cudaMemcpyDefault, stream_));
    if (albedo && film_.albedo) BTRC_CUDA(cudaMemcpyAsync(albedo, film_.albedo, 16 * npix, cudaMemcpyDefault, stream_));
    if (normal && film_.normal) BTRC_CUDA(cudaMemcpyAsync(normal, film_.normal, 16 * npix, cudaMemcpyDefault, stream_));
    BTRC_CUDA(cudaStreamSynchronize(stream_));
}

void Renderer::import_film(const float *radiance, const float *weight, const float *albedo, const float *normal) {
    HostSpan span("import_film");
    BTRC_CUDA(cudaDeviceSynchronize());
    const size_t npix = (size_t)scene_.width * scene_.height;
    if (radiance) BTRC_CUDA(cudaMemcpyAsync(film_.radiance, radiance, 16 * npix, cudaMemcpyDefault, stream_));
    if (weight) BTRC_CUDA(cudaMemcpyAsync(film_.weight, weight, 4 * npix, cudaMemcpyDefault, stream_));
    if (albedo && film_.albedo) BTRC_CUDA(cudaMemcpyAsync(film_.albedo, albedo, 16 * npix, cudaMemcpyDefault, stream_));
    if (normal && film_.normal) BTRC_CUDA(cudaMemcpyAsync(film_.normal, normal, 16 * npix, cudaMemcpyDefault, stream_));
    BTRC_CUDA(cudaStreamSynchronize(stream_));
}

}  // namespace btrc
