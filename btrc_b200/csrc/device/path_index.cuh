// Which (pixel, sample) the i-th camera path of a render is.
//
// The reference numbers its camera paths in one index space (generate.cpp:49-60): Uniform mode index = sample * W*H + pixel,
// Tile mode index = pixel * spp + sample.  A path is a function of (pixel, sample) alone (sampler seed: independent.cpp:11-15), so a
// renderer is free to VISIT the index space in any order and to split it between devices; this header holds the closed-form maps
// from the render's running ("virtual") index to (x, y, sample) for
//   * index order (both modes, any sub-range),
//   * the multi-GPU partitions of the tile-mode space: interleaved strips of whole rows and interleaved 2-D tiles (types.h
//     WaveCtl::stride_*), walked warp by warp inside a tile when spp is a multiple of 32,
//   * the block-coherent order over whole sample planes of the uniform space (WaveCtl::order_*).
// One division of the running index -- 64-bit only when the render has more than 2^32 paths -- and 32-bit arithmetic after it
// (the host keeps every inner extent below 2^31: renderer.cpp).  Integer arithmetic only; compiled for the device by kernels.cu
// and for the host by tests/host/host_path_index.cpp, which checks every map for being one-to-one onto its index set.
#pragma once
#include <cstdint>

#include "../common/types.h"

#ifndef BTRC_HD
#ifdef __CUDACC__
#define BTRC_HD __host__ __device__ __forceinline__
#else
#define BTRC_HD inline
#endif
#endif

namespace btrc {

struct PathIndexMap {  // uniform over a launch
    long long sblock;                       // WaveCtl::stride_block
    uint32_t sblock32, sparts, spart, sbw, srows, tiles_x;
    uint32_t order_ns, order_s0, obw, obh, per_block_row;
    int osm;                                // WaveCtl::order_sample_major
    uint32_t W, H, spp;
    int tile;
    bool small;                             // every running index of this render fits 32 bits
};

BTRC_HD PathIndexMap make_path_index_map(const WaveCtl &ctl, int width, int height, int spp, int tile) {
    PathIndexMap m;
    m.sblock = ctl.stride_block;
    m.sblock32 = (uint32_t)ctl.stride_block;
    m.sparts = (uint32_t)ctl.stride_parts; m.spart = (uint32_t)ctl.stride_part;
    m.sbw = (uint32_t)ctl.stride_bw; m.srows = (uint32_t)ctl.stride_rows;
    m.W = (uint32_t)width; m.H = (uint32_t)height; m.spp = (uint32_t)spp; m.tile = tile;
    m.tiles_x = m.sbw ? m.W / m.sbw : 1;
    m.order_ns = (uint32_t)ctl.order_ns; m.order_s0 = (uint32_t)ctl.order_s0;
    m.obw = (uint32_t)ctl.order_bw; m.obh = (uint32_t)ctl.order_bh;
    m.per_block_row = m.W * m.obh * m.order_ns;   // block-coherent order: paths of one row of blocks
    m.osm = ctl.order_sample_major;
    m.small = ctl.end <= 0xffffffffLL;
    return m;
}

// a / d and the remainder: in 32-bit arithmetic when `small`, else one 64-bit division.  Quotient and remainder fit 32 bits wherever
// this is used.
BTRC_HD uint32_t index_divmod(long long a, uint32_t d, bool small, uint32_t &rem) {
    if (small) {
        const uint32_t a32 = (uint32_t)a, q = a32 / d;
        rem = a32 - q * d;
        return q;
    }
    const long long q = a / (long long)d;
    rem = (uint32_t)(a - q * (long long)d);
    return (uint32_t)q;
}

BTRC_HD void decode_path_index(const PathIndexMap &m, long long accu, uint32_t &x, uint32_t &y, uint32_t &sample) {
    if (m.sblock > 0 && m.sbw > 0) {  // 2-D tiles of the tile-mode index space (types.h WaveCtl::stride_bw)
        uint32_t within;
        const uint32_t blk = index_divmod(accu, m.sblock32, m.small, within) * m.sparts + m.spart;
        uint32_t pq;  // pixel ordinal inside the tile
        if (m.spp % 32 == 0) {  // 32 samples of a pixel (one warp), then the next pixel of the tile, ... then the next 32 samples
            const uint32_t per_chunk = m.sbw * m.srows * 32, chunk = within / per_chunk, rem = within - chunk * per_chunk;
            pq = rem >> 5; sample = chunk * 32 + (rem & 31);
        } else {
            pq = within / m.spp; sample = within - pq * m.spp;
        }
        const uint32_t ty = blk / m.tiles_x, tx = blk - ty * m.tiles_x, qy = pq / m.sbw, qx = pq - qy * m.sbw;
        x = tx * m.sbw + qx; y = ty * m.srows + qy;
    } else if (m.order_ns > 0) {  // block-coherent order over whole sample planes (types.h WaveCtl::order_ns)
        uint32_t in_row;
        const uint32_t br = index_divmod(accu, m.per_block_row, m.small, in_row);
        const uint32_t hr = (br + 1) * m.obh <= m.H ? m.obh : m.H - br * m.obh;   // the last block row may be lower
        uint32_t bc, within;  // block column, pixel ordinal inside the block (row-major)
        if (!m.osm) {    // all samples of a pixel consecutively
            const uint32_t pq = in_row / m.order_ns;
            sample = m.order_s0 + (in_row - pq * m.order_ns);
            bc = pq / (m.obw * hr);
            within = pq - bc * m.obw * hr;
        } else {
            const uint32_t full = m.obw * hr * m.order_ns;                  // paths of a full-width block of this row
            bc = in_row / full;
            const uint32_t in_block = in_row - bc * full;
            const uint32_t npb = ((bc + 1) * m.obw <= m.W ? m.obw : m.W - bc * m.obw) * hr;   // pixels of this block
            if (m.osm == 2) {  // 32 samples of a pixel (one warp), then the next pixel, ... then the next 32 samples
                const uint32_t chunk = in_block / (npb * 32), rem = in_block - chunk * npb * 32;
                sample = m.order_s0 + chunk * 32 + (rem & 31);
                within = rem >> 5;
            } else {           // sample plane by sample plane
                const uint32_t plane = in_block / npb;
                sample = m.order_s0 + plane;
                within = in_block - plane * npb;
            }
        }
        const uint32_t wc = (bc + 1) * m.obw <= m.W ? m.obw : m.W - bc * m.obw;
        const uint32_t qy = within / wc;
        x = bc * m.obw + (within - qy * wc); y = br * m.obh + qy;
    } else {
        if (m.sblock > 0) accu = ((accu / m.sblock) * m.sparts + m.spart) * m.sblock + accu % m.sblock;  // strips of whole rows (WaveCtl::stride_block)
        uint32_t pixel_index;
        if (!m.tile) {   // generate.cpp:49-55
            sample = index_divmod(accu, m.W * m.H, accu <= 0xffffffffLL, pixel_index);
        } else {         // generate.cpp:56-60
            pixel_index = index_divmod(accu, m.spp, accu <= 0xffffffffLL, sample);
        }
        y = pixel_index / m.W; x = pixel_index - y * m.W;
    }
}

}  // namespace btrc
