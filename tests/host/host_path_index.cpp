// Host build of the product's path-index maps (btrc_b200/csrc/device/path_index.cuh) for tests/test_path_index_host.py.
#include <cstdint>
#include <cstring>

#include "../../btrc_b200/csrc/device/path_index.cuh"

using namespace btrc;

extern "C" {

// decodes running indices idx[0..n) of a render described by the WaveCtl fields below into the reference's own path index
// (uniform: sample * W*H + pixel; tile: pixel * spp + sample); out[k] = -1 when (x, y, sample) is out of range
void hpi_decode(int width, int height, int spp, int tile, long long end, long long stride_block, int stride_parts, int stride_part, int stride_bw,
                int stride_rows, int order_ns, int order_s0, int order_bw, int order_bh, int order_mode, long long n, const long long *idx, long long *out) {
    WaveCtl ctl;
    std::memset(&ctl, 0, sizeof(ctl));
    ctl.end = end;
    ctl.stride_block = stride_block; ctl.stride_parts = stride_parts; ctl.stride_part = stride_part;
    ctl.stride_bw = stride_bw; ctl.stride_rows = stride_rows;
    ctl.order_ns = order_ns; ctl.order_s0 = order_s0; ctl.order_bw = order_bw; ctl.order_bh = order_bh; ctl.order_sample_major = order_mode;
    const PathIndexMap m = make_path_index_map(ctl, width, height, spp, tile);
#pragma omp parallel for
    for (long long k = 0; k < n; ++k) {
        uint32_t x, y, s;
        decode_path_index(m, idx[k], x, y, s);
        if (x >= (uint32_t)width || y >= (uint32_t)height) { out[k] = -1; continue; }
        const long long pixel = (long long)y * width + x;
        out[k] = tile ? pixel * spp + s : (long long)s * width * height + pixel;
    }
}

}  // extern "C"
