#include "image_io.h"

#include <zlib.h>
#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "node.h"

namespace btrc {

namespace {

std::string lower_ext(const std::string &f) {
    const size_t p = f.find_last_of('.');
    std::string e = p == std::string::npos ? "" : f.substr(p + 1);
    std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return e;
}

void put_u32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

void png_chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data) {
    put_u32(out, (uint32_t)data.size());
    std::vector<uint8_t> body(type, type + 4);
    body.insert(body.end(), data.begin(), data.end());
    out.insert(out.end(), body.begin(), body.end());
    put_u32(out, (uint32_t)crc32(0, body.data(), (uInt)body.size()));
}

void write_png_rgba8(const std::string &filename, int w, int h, const std::vector<uint8_t> &rgba) {
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (4 * w + 1));
    for (int y = 0; y < h; ++y) {
        raw.push_back(0);  // filter: none
        raw.insert(raw.end(), rgba.begin() + (size_t)y * 4 * w, rgba.begin() + (size_t)(y + 1) * 4 * w);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw Error("failed to write png file: " + filename);
    z.resize(zlen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_u32(ihdr, (uint32_t)w); put_u32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    std::ofstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to write png file: " + filename);
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
}

void write_hdr(const std::string &filename, int w, int h, const float *rgba) {
    std::ofstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to write hdr file: " + filename);
    f << "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " << h << " +X " << w << "\n";
    std::vector<uint8_t> row(4 * (size_t)w);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const float *p = rgba + 4 * ((size_t)y * w + x);
            const float m = std::max(p[0], std::max(p[1], p[2]));
            uint8_t *o = &row[4 * (size_t)x];
            if (!(m > 1e-32f)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            int e;
            const float s = std::frexp(m, &e) * 256.0f / m;
            o[0] = (uint8_t)(p[0] * s); o[1] = (uint8_t)(p[1] * s); o[2] = (uint8_t)(p[2] * s); o[3] = (uint8_t)(e + 128);
        }
        f.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)row.size());
    }
}

// OpenEXR, what tinyexr's SaveEXRImageToFile writes for the reference (utils/image.cpp:50-111): single-part scanline file, three
// 32-bit float channels named B, G, R, no compression (InitEXRHeader zero-fills compression_type), increasing-y line order.
void write_exr(const std::string &filename, int w, int h, const float *rgba) {
    std::vector<uint8_t> out;
    auto put = [&](const void *p, size_t n) { const uint8_t *b = static_cast<const uint8_t *>(p); out.insert(out.end(), b, b + n); };
    auto put_i32 = [&](int32_t v) { put(&v, 4); };
    auto put_f32 = [&](float v) { put(&v, 4); };
    auto put_str = [&](const char *z) { put(z, std::strlen(z) + 1); };
    auto attr = [&](const char *name, const char *type, int32_t size) { put_str(name); put_str(type); put_i32(size); };
    put_i32(20000630);  // magic
    put_i32(2);         // version 2, no flags
    attr("channels", "chlist", 3 * 18 + 1);
    for (const char *c : {"B", "G", "R"}) {
        put_str(c);
        put_i32(2);  // FLOAT
        const uint8_t plinear_reserved[4] = {0, 0, 0, 0};
        put(plinear_reserved, 4);
        put_i32(1); put_i32(1);  // sampling
    }
    out.push_back(0);
    attr("compression", "compression", 1); out.push_back(0);
    attr("dataWindow", "box2i", 16); put_i32(0); put_i32(0); put_i32(w - 1); put_i32(h - 1);
    attr("displayWindow", "box2i", 16); put_i32(0); put_i32(0); put_i32(w - 1); put_i32(h - 1);
    attr("lineOrder", "lineOrder", 1); out.push_back(0);
    attr("pixelAspectRatio", "float", 4); put_f32(1.0f);
    attr("screenWindowCenter", "v2f", 8); put_f32(0.0f); put_f32(0.0f);
    attr("screenWindowWidth", "float", 4); put_f32(1.0f);
    out.push_back(0);  // end of header
    const uint64_t row_bytes = 8 + 12ull * (uint64_t)w;
    uint64_t offset = out.size() + 8ull * (uint64_t)h;
    for (int y = 0; y < h; ++y, offset += row_bytes) put(&offset, 8);
    std::vector<float> plane((size_t)w);
    for (int y = 0; y < h; ++y) {
        put_i32(y);
        put_i32(12 * w);
        for (int c : {2, 1, 0}) {  // B, G, R planes
            for (int x = 0; x < w; ++x) plane[x] = rgba[4 * ((size_t)y * w + x) + c];
            put(plane.data(), 4 * (size_t)w);
        }
    }
    std::ofstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to save exr file: " + filename);
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
}

// Baseline JPEG as stb_image_write produces it for the reference's quality argument 0 (utils/image.cpp:28-38): quality 90, which
// in stb_image_write also means 2x2 chroma subsampling (16x16 MCUs, chroma averaged over 2x2 pixels); ITU T.81 Annex K tables, one scan.
void write_jpg_rgb8(const std::string &filename, int w, int h, const std::vector<uint8_t> &rgb) {
    static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    static const uint8_t qy[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                   18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const uint8_t qc[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                   99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    static const uint8_t dc_y_bits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, dc_c_bits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
    static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
    static const uint8_t ac_y_bits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d}, ac_c_bits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
    static const uint8_t ac_y_vals[162] = {
        0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
        0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
        0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
        0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
        0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
        0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
    static const uint8_t ac_c_vals[162] = {
        0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
        0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
        0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
        0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
        0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
        0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
    struct Huff { uint16_t code[256]; uint8_t len[256]; };
    auto build = [](const uint8_t *bits, const uint8_t *vals, Huff &t) {
        std::memset(&t, 0, sizeof(t));
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l, code <<= 1)
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) { t.code[vals[k]] = (uint16_t)code; t.len[vals[k]] = (uint8_t)l; }
    };
    Huff hdc_y, hdc_c, hac_y, hac_c;
    build(dc_y_bits, dc_vals, hdc_y); build(dc_c_bits, dc_vals, hdc_c); build(ac_y_bits, ac_y_vals, hac_y); build(ac_c_bits, ac_c_vals, hac_c);
    const int quality = 90, scale = 200 - 2 * quality;
    uint8_t tq[2][64];
    for (int i = 0; i < 64; ++i) {
        tq[0][i] = (uint8_t)std::min(std::max((qy[i] * scale + 50) / 100, 1), 255);
        tq[1][i] = (uint8_t)std::min(std::max((qc[i] * scale + 50) / 100, 1), 255);
    }
    std::vector<uint8_t> out;
    auto put = [&](std::initializer_list<int> b) { for (int v : b) out.push_back((uint8_t)v); };
    put({0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; ++t) {
        put({0xff, 0xdb, 0, 67, t});
        for (int i = 0; i < 64; ++i) out.push_back(tq[t][zz[i]]);
    }
    put({0xff, 0xc0, 0, 17, 8, h >> 8, h & 255, w >> 8, w & 255, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    auto dht = [&](int cls_id, const uint8_t *bits, const uint8_t *vals, int n) {
        put({0xff, 0xc4, (19 + n) >> 8, (19 + n) & 255, cls_id});
        for (int l = 1; l <= 16; ++l) out.push_back(bits[l]);
        for (int i = 0; i < n; ++i) out.push_back(vals[i]);
    };
    dht(0x00, dc_y_bits, dc_vals, 12); dht(0x10, ac_y_bits, ac_y_vals, 162); dht(0x01, dc_c_bits, dc_vals, 12); dht(0x11, ac_c_bits, ac_c_vals, 162);
    put({0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    uint32_t acc = 0;
    int nbits = 0;
    auto emit = [&](uint32_t code, int len) {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        nbits += len;
        while (nbits >= 8) {
            const uint8_t b = (uint8_t)(acc >> (nbits - 8));
            out.push_back(b);
            if (b == 0xff) out.push_back(0);
            nbits -= 8;
        }
    };
    auto magnitude = [](int v, int &size, uint32_t &bits) {
        int a = v < 0 ? -v : v;
        size = 0;
        while (a) { ++size; a >>= 1; }
        bits = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << size) - 1u);
    };
    float cosv[8][8];
    for (int u = 0; u < 8; ++u)
        for (int x = 0; x < 8; ++x) cosv[u][x] = std::cos((2 * x + 1) * u * 3.14159265358979f / 16.0f) * (u == 0 ? 0.353553391f : 0.5f);
    int pred[3] = {0, 0, 0};
    // one 8x8 data unit: separable DCT, quantisation, Huffman coding (stride = row pitch of `src`)
    auto code_block = [&](const float *src, int stride, int c) {
        float tmp[64], dct[64];
        for (int y = 0; y < 8; ++y)
            for (int u = 0; u < 8; ++u) {
                float a = 0;
                for (int x = 0; x < 8; ++x) a += src[stride * y + x] * cosv[u][x];
                tmp[8 * y + u] = a;
            }
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u) {
                float a = 0;
                for (int y = 0; y < 8; ++y) a += tmp[8 * y + u] * cosv[v][y];
                dct[8 * v + u] = a;
            }
        int q[64];
        const uint8_t *tab = tq[c ? 1 : 0];
        for (int i = 0; i < 64; ++i) q[i] = (int)std::lround(dct[zz[i]] / tab[zz[i]]);
        const Huff &hd = c ? hdc_c : hdc_y, &ha = c ? hac_c : hac_y;
        int size; uint32_t bits;
        magnitude(q[0] - pred[c], size, bits);
        pred[c] = q[0];
        emit(hd.code[size], hd.len[size]);
        if (size) emit(bits, size);
        int run = 0;
        for (int i = 1; i < 64; ++i) {
            if (q[i] == 0) { ++run; continue; }
            while (run > 15) { emit(ha.code[0xf0], ha.len[0xf0]); run -= 16; }
            magnitude(q[i], size, bits);
            emit(ha.code[(run << 4) | size], ha.len[(run << 4) | size]);
            emit(bits, size);
            run = 0;
        }
        if (run) emit(ha.code[0], ha.len[0]);
    };
    // 16x16 MCUs: four luma units, then Cb and Cr averaged over 2x2 pixels; pixels beyond the image repeat the last row / column
    for (int my = 0; my < h; my += 16)
        for (int mx = 0; mx < w; mx += 16) {
            float Y[256], U[256], V[256];
            for (int y = 0; y < 16; ++y)
                for (int x = 0; x < 16; ++x) {
                    const uint8_t *p = &rgb[3 * ((size_t)std::min(my + y, h - 1) * w + std::min(mx + x, w - 1))];
                    const float r = p[0], g = p[1], b = p[2];
                    Y[16 * y + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
                    U[16 * y + x] = -0.16874f * r - 0.33126f * g + 0.5f * b;
                    V[16 * y + x] = 0.5f * r - 0.41869f * g - 0.08131f * b;
                }
            code_block(Y, 16, 0); code_block(Y + 8, 16, 0); code_block(Y + 128, 16, 0); code_block(Y + 136, 16, 0);
            float su[64], sv[64];
            for (int y = 0; y < 8; ++y)
                for (int x = 0; x < 8; ++x) {
                    const int j = 32 * y + 2 * x;
                    su[8 * y + x] = (U[j] + U[j + 1] + U[j + 16] + U[j + 17]) * 0.25f;
                    sv[8 * y + x] = (V[j] + V[j + 1] + V[j + 16] + V[j + 17]) * 0.25f;
                }
            code_block(su, 8, 1); code_block(sv, 8, 2);
        }
    if (nbits) emit(0x7f, 8 - nbits);
    put({0xff, 0xd9});
    std::ofstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to write jpg file: " + filename);
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
}

uint32_t get_u32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// PNG reader for ImageDynamic::load (utils/image.cpp:116-147 reads through stb_image's 8-bit interface).  Covers what that
// interface covers: colour types 0/2/3/4/6, bit depths 1/2/4/8/16, Adam7 interlacing, tRNS.  Conventions follow stb_image:
// 16-bit samples are reduced to their high byte, 1/2/4-bit grey is scaled to 0..255 (x 255 / (2^depth - 1)), a palette image
// becomes RGB (RGBA when it has a tRNS chunk), a colour-keyed grey / RGB image gains an alpha channel (0 for the key, 255 otherwise).
LoadedImage load_png(const std::string &filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to open file: " + filename);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const Error bad("failed to load image: " + filename);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) throw bad;
    size_t p = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (p + 8 <= buf.size()) {
        const uint32_t len = get_u32(&buf[p]);
        const std::string type(reinterpret_cast<const char *>(&buf[p + 4]), 4);
        const uint8_t *d = &buf[p + 8];
        if (p + 12 + (size_t)len > buf.size()) throw bad;
        if (type == "IHDR") {
            if (len < 13) throw bad;
            w = (int)get_u32(d); h = (int)get_u32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        }
        else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "tRNS") trns.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        p += 12 + len;
    }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (!ch || !depth_ok || w <= 0 || h <= 0 || interlace > 1 || (ctype == 3 && plte.empty())) throw bad;

    // the passes of the image: one for a sequential file, seven for Adam7 (x0, y0, dx, dy)
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole = {0, 0, 1, 1};
    const int npass = interlace ? 7 : 1;
    const int bits = depth * ch;            // per pixel
    const size_t fbpp = (size_t)std::max(1, bits / 8);  // filter distance in bytes
    auto pass_dim = [&](const Pass &ps, int &pw, int &ph) { pw = (w - ps.x0 + ps.dx - 1) / ps.dx; ph = (h - ps.y0 + ps.dy - 1) / ps.dy; };
    size_t total = 0;
    for (int i = 0; i < npass; ++i) {
        int pw, ph;
        pass_dim(interlace ? adam7[i] : whole, pw, ph);
        if (pw > 0 && ph > 0) total += (((size_t)pw * bits + 7) / 8 + 1) * ph;
    }
    // deflate expands by at most ~1030x: a size the IDAT data cannot fill is a damaged header, not a reason to allocate it
    if (total > 1100ull * idat.size() + 64 || (uint64_t)w * (uint64_t)h > (1ull << 32)) throw bad;
    std::vector<uint8_t> raw(total);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) throw bad;

    std::vector<uint16_t> smp((size_t)w * h * ch);  // raw sample values, unscaled
    size_t at = 0;
    std::vector<uint8_t> cur, prev;
    for (int i = 0; i < npass; ++i) {
        const Pass &ps = interlace ? adam7[i] : whole;
        int pw, ph;
        pass_dim(ps, pw, ph);
        if (pw <= 0 || ph <= 0) continue;
        const size_t rb = ((size_t)pw * bits + 7) / 8;
        cur.assign(rb, 0); prev.assign(rb, 0);
        for (int y = 0; y < ph; ++y) {
            const uint8_t ft = raw[at];
            const uint8_t *src = &raw[at + 1];
            at += rb + 1;
            if (ft > 4) throw bad;
            for (size_t x = 0; x < rb; ++x) {
                const int a = x >= fbpp ? cur[x - fbpp] : 0, b = prev[x], c = x >= fbpp ? prev[x - fbpp] : 0;
                int v = src[x];
                switch (ft) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
                }
                cur[x] = (uint8_t)v;
            }
            const int oy = ps.y0 + y * ps.dy;
            for (int x = 0; x < pw; ++x) {
                uint16_t *o = &smp[((size_t)oy * w + ps.x0 + (size_t)x * ps.dx) * ch];
                for (int k = 0; k < ch; ++k) {
                    const size_t s = (size_t)x * ch + k;
                    if (depth == 8) o[k] = cur[s];
                    else if (depth == 16) o[k] = (uint16_t)(cur[2 * s] << 8 | cur[2 * s + 1]);
                    else { const size_t bit = s * depth; o[k] = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1); }
                }
            }
            cur.swap(prev);
        }
    }

    LoadedImage img;
    img.width = w; img.height = h;
    const size_t npx = (size_t)w * h;
    if (ctype == 3) {  // palette -> rgb(a)
        img.channels = trns.empty() ? 3 : 4;
        img.u8.resize(npx * img.channels);
        for (size_t i = 0; i < npx; ++i) {
            const size_t e = smp[i];
            for (int k = 0; k < 3; ++k) img.u8[img.channels * i + k] = 3 * e + k < plte.size() ? plte[3 * e + k] : 0;
            if (img.channels == 4) img.u8[4 * i + 3] = e < trns.size() ? trns[e] : 255;
        }
        return img;
    }
    const bool keyed = (ctype == 0 && trns.size() >= 2) || (ctype == 2 && trns.size() >= 6);
    uint16_t key[3] = {0, 0, 0};
    if (keyed) for (int k = 0; k < ch; ++k) key[k] = (uint16_t)(trns[2 * k] << 8 | trns[2 * k + 1]);
    img.channels = ch + (keyed ? 1 : 0);
    img.u8.resize(npx * img.channels);
    const int scale = depth < 8 ? 255 / ((1 << depth) - 1) : 1;
    for (size_t i = 0; i < npx; ++i) {
        bool is_key = keyed;
        for (int k = 0; k < ch; ++k) {
            const uint16_t v = smp[i * ch + k];
            if (keyed && v != key[k]) is_key = false;
            img.u8[i * img.channels + k] = depth == 16 ? (uint8_t)(v >> 8) : (uint8_t)(v * scale);
        }
        if (keyed) img.u8[i * img.channels + ch] = is_key ? 0 : 255;
    }
    return img;
}

}  // namespace

void save_image_rgba32f(const std::string &filename, int width, int height, const float *data) {
    const std::string ext = lower_ext(filename);
    if (ext == "png") {
        std::vector<uint8_t> rgba(4 * (size_t)width * height);
        for (size_t i = 0; i < rgba.size(); ++i) {
            const float v = data[i];
            const int q = std::isfinite(v) ? std::min(static_cast<int>(v * 256), 255) : 0;  // image.h:212-216
            rgba[i] = (uint8_t)std::max(q, 0);
        }
        write_png_rgba8(filename, width, height, rgba);
    } else if (ext == "hdr") {
        write_hdr(filename, width, height, data);
    } else if (ext == "exr") {
        write_exr(filename, width, height, data);
    } else if (ext == "jpg" || ext == "jpeg") {
        std::vector<uint8_t> rgb(3 * (size_t)width * height);
        for (size_t i = 0; i < (size_t)width * height; ++i)
            for (int c = 0; c < 3; ++c) {
                const float v = data[4 * i + c];
                rgb[3 * i + c] = (uint8_t)std::max(std::isfinite(v) ? std::min(static_cast<int>(v * 256), 255) : 0, 0);  // image.h:212-216
            }
        write_jpg_rgb8(filename, width, height, rgb);
    } else
        throw Error("unknown image format: " + filename);
}

// Radiance .hdr (RGBE; flat or new-style RLE scanlines), "-Y h +X w" orientation -- what stbi_loadf reads
static LoadedImage load_hdr(const std::string &filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to open file: " + filename);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t p = 0;
    auto getline = [&]() {
        std::string l;
        while (p < buf.size() && buf[p] != '\n') l.push_back((char)buf[p++]);
        ++p;
        return l;
    };
    std::string l = getline();
    if (l.rfind("#?", 0) != 0) throw Error("failed to load image: " + filename);
    while (p < buf.size() && !(l = getline()).empty()) {}
    l = getline();
    int w = 0, h = 0;
    if (std::sscanf(l.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) throw Error("unsupported hdr orientation: " + filename);
    // a scanline takes at least 8 bytes per 127 pixels (run-length coded): sizes the file cannot hold are a damaged header
    if ((uint64_t)w * (uint64_t)h > 64ull * buf.size() + 64) throw Error("failed to load image: " + filename);
    LoadedImage img;
    img.width = w; img.height = h; img.channels = 3; img.is_float = true;
    img.f32.resize((size_t)w * h * 3);
    std::vector<uint8_t> row(4 * (size_t)w);
    for (int y = 0; y < h; ++y) {
        bool rle = false;
        if (w >= 8 && w < 32768 && p + 4 <= buf.size() && buf[p] == 2 && buf[p + 1] == 2 && (buf[p + 2] & 0x80) == 0) {
            rle = ((int(buf[p + 2]) << 8) | buf[p + 3]) == w;
        }
        if (rle) {
            p += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < w) {
                    if (p >= buf.size()) throw Error("failed to load image: " + filename);
                    int count = buf[p++];
                    if (count > 128) {
                        count -= 128;
                        if (p >= buf.size() || x + count > w) throw Error("failed to load image: " + filename);
                        const uint8_t v = buf[p++];
                        for (int k = 0; k < count; ++k) row[4 * (size_t)(x++) + c] = v;
                    } else {
                        if (p + count > buf.size() || x + count > w) throw Error("failed to load image: " + filename);
                        for (int k = 0; k < count; ++k) row[4 * (size_t)(x++) + c] = buf[p++];
                    }
                }
            }
        } else {
            if (p + 4 * (size_t)w > buf.size()) throw Error("failed to load image: " + filename);
            std::memcpy(row.data(), &buf[p], 4 * (size_t)w);
            p += 4 * (size_t)w;
        }
        for (int x = 0; x < w; ++x) {
            const uint8_t *q = &row[4 * (size_t)x];
            float *o = &img.f32[3 * ((size_t)y * w + x)];
            if (q[3] == 0) { o[0] = o[1] = o[2] = 0; continue; }
            const float s = std::ldexp(1.0f, int(q[3]) - (128 + 8));
            o[0] = q[0] * s; o[1] = q[1] * s; o[2] = q[2] * s;
        }
    }
    return img;
}

// Huffman-coded JPEG (ITU T.81: SOF0 baseline, SOF1 extended sequential, SOF2 progressive; 8-bit samples), 1 or 3 components, any
// 1x/2x sampling factors, restart intervals, interleaved and single-component scans.  The reference reads .jpg textures through
// stb_image (utils/image.cpp:135-143), which covers the same set; arithmetic-coded and lossless files are reported as unsupported
// (stb_image rejects them too).  Every scan is decoded into per-component coefficient arrays (spectral selection and successive
// approximation as in T.81 annex G); dequantisation and the IDCT run once after the last scan, in stb_image's integer arithmetic:
// the decoded bytes are those of the reference's loader (tests/test_reference_host_pins_cpu.py) and within 1-3 eight-bit levels of
// libjpeg (tests/test_postprocess_cpu.py).
static LoadedImage load_jpg(const std::string &filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to open file: " + filename);
    const std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto fail = [&](const std::string &why) -> Error { return Error("failed to load image: " + filename + " (" + why + ")"); };
    if (buf.size() < 4 || buf[0] != 0xff || buf[1] != 0xd8) throw fail("not a JPEG file");
    static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    struct Huff {
        uint8_t bits[17] = {0};
        uint8_t vals[256] = {0};
        int mincode[17], maxcode[18], valptr[17];
        bool ok = false;
        void build() {
            int code = 0, k = 0;
            for (int l = 1; l <= 16; ++l) {
                valptr[l] = k; mincode[l] = code;
                code += bits[l]; k += bits[l];
                maxcode[l] = bits[l] ? code - 1 : -1;
                code <<= 1;
            }
            maxcode[17] = 0x7fffffff;
            ok = true;
        }
    };
    struct Comp {
        int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0;
        int bw = 0, bh = 0;    // padded size in pixels (whole MCUs)
        int nbx = 0, nby = 0;  // padded size in blocks: the layout of coef
        int cbx = 0, cby = 0;  // blocks that cover the component itself: what a single-component scan codes (T.81 A.2.2)
        std::vector<int16_t> coef;
        std::vector<uint8_t> px;
    };
    uint16_t qt[4][64] = {{0}};
    Huff hdc[4], hac[4];
    std::vector<Comp> comps;
    int width = 0, height = 0, restart = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, scans = 0;
    bool progressive = false, jfif = false;
    int adobe_transform = -1;
    auto u16 = [&](size_t at) { if (at + 1 >= buf.size()) throw fail("truncated"); return (int)buf[at] << 8 | buf[at + 1]; };

    // entropy-coded segment reader: stops feeding at a marker (zeros from there on), never steps over one
    size_t bp = 0;
    uint32_t acc = 0;
    int nb = 0;
    auto fill = [&]() {
        while (nb <= 24) {
            int b = 0;
            if (bp < buf.size()) {
                b = buf[bp];
                if (b == 0xff) {
                    const int n2 = bp + 1 < buf.size() ? buf[bp + 1] : 0xd9;
                    if (n2 == 0) bp += 2;
                    else b = 0;  // a marker: feed zeros, do not advance
                } else ++bp;
            }
            acc |= (uint32_t)b << (24 - nb);
            nb += 8;
        }
    };
    auto getbits = [&](int n) -> int {
        if (n == 0) return 0;
        if (n > 16) throw fail("bad bit count");  // a magnitude category from a damaged Huffman table
        fill();
        const int v = (int)(acc >> (32 - n));
        acc <<= n; nb -= n;
        return v;
    };
    auto decode = [&](const Huff &h) -> int {
        fill();
        int code = 0;
        for (int l = 1; l <= 16; ++l) {
            code = (code << 1) | (int)(acc >> 31);
            acc <<= 1; --nb;
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        throw fail("bad Huffman code");
    };
    auto extend = [](int v, int t) { return t && v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; };

    // one scan: `sc` = the components it codes, spectral band [ss, se], successive-approximation bits ah / al
    auto decode_scan = [&](const std::vector<Comp *> &sc, int ss, int se, int ah, int al) {
        acc = 0; nb = 0;
        int eobrun = 0;
        for (Comp *c : sc) c->pred = 0;
        const int bit = 1 << al;
        auto refine = [&](int16_t &q) {  // correction bit of an already non-zero coefficient (G.1.2.3)
            if (getbits(1) && (q & bit) == 0) q = (int16_t)(q > 0 ? q + bit : q - bit);
        };
        auto block = [&](Comp &c, int16_t *q) {
            if (!progressive) {
                if (!hdc[c.td].ok || !hac[c.ta].ok) throw fail("missing Huffman table");
                const int t = decode(hdc[c.td]);
                c.pred += extend(getbits(t), t);
                q[0] = (int16_t)c.pred;
                for (int k = 1; k < 64;) {
                    const int rs = decode(hac[c.ta]), r = rs >> 4, sz = rs & 15;
                    if (sz == 0) { if (r != 15) break; k += 16; continue; }
                    k += r;
                    if (k > 63) throw fail("bad coefficient index");
                    q[zz[k]] = (int16_t)extend(getbits(sz), sz);
                    ++k;
                }
                return;
            }
            if (ss == 0) {  // DC scan: first pass sends the value >> al, refinement one more bit
                if (ah == 0) {
                    if (!hdc[c.td].ok) throw fail("missing Huffman table");
                    const int t = decode(hdc[c.td]);
                    c.pred += extend(getbits(t), t);
                    q[0] = (int16_t)(c.pred * bit);
                } else if (getbits(1)) q[0] = (int16_t)(q[0] | bit);
                return;
            }
            if (!hac[c.ta].ok) throw fail("missing Huffman table");
            if (ah == 0) {  // AC first pass (G.1.2.2): run/size symbols with end-of-band runs
                if (eobrun) { --eobrun; return; }
                for (int k = ss; k <= se;) {
                    const int rs = decode(hac[c.ta]), r = rs >> 4, sz = rs & 15;
                    if (sz == 0) {
                        if (r < 15) { eobrun = (1 << r) - 1 + (r ? getbits(r) : 0); break; }
                        k += 16;
                        continue;
                    }
                    k += r;
                    if (k > se) throw fail("bad coefficient index");
                    q[zz[k]] = (int16_t)(extend(getbits(sz), sz) * bit);
                    ++k;
                }
                return;
            }
            // AC refinement (G.1.2.3): new +-1 coefficients are placed after skipping r *zero* coefficients; every non-zero
            // coefficient passed on the way takes one correction bit
            if (eobrun) {
                --eobrun;
                for (int k = ss; k <= se; ++k) if (q[zz[k]]) refine(q[zz[k]]);
                return;
            }
            for (int k = ss; k <= se;) {
                const int rs = decode(hac[c.ta]), sz = rs & 15;
                int r = rs >> 4, val = 0;
                if (sz == 0) {
                    if (r < 15) { eobrun = (1 << r) - 1 + (r ? getbits(r) : 0); r = 64; }  // rest of the band: corrections only
                } else {
                    if (sz != 1) throw fail("bad refinement symbol");
                    val = getbits(1) ? bit : -bit;
                }
                while (k <= se) {
                    int16_t &cq = q[zz[k++]];
                    if (cq) refine(cq);
                    else if (r == 0) { cq = (int16_t)val; break; }
                    else --r;
                }
            }
        };
        auto at_restart = [&](int done) {
            if (!(restart && done && done % restart == 0)) return;
            acc = 0; nb = 0; eobrun = 0;  // RSTn: byte align, skip the marker, reset the predictors
            while (bp + 1 < buf.size() && !(buf[bp] == 0xff && buf[bp + 1] >= 0xd0 && buf[bp + 1] <= 0xd7)) ++bp;
            bp += 2;
            for (Comp *c : sc) c->pred = 0;
        };
        int done = 0;
        if (sc.size() == 1) {
            Comp &c = *sc[0];
            for (int j = 0; j < c.cby; ++j)
                for (int i = 0; i < c.cbx; ++i) {
                    at_restart(done++);
                    block(c, &c.coef[((size_t)j * c.nbx + i) * 64]);
                }
        } else {
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    at_restart(done++);
                    for (Comp *c : sc)
                        for (int by = 0; by < c->v; ++by)
                            for (int bx = 0; bx < c->h; ++bx) block(*c, &c->coef[((size_t)(my * c->v + by) * c->nbx + mx * c->h + bx) * 64]);
                }
        }
    };

    size_t p = 2;
    bool eoi = false;
    while (!eoi && p + 4 <= buf.size()) {
        if (buf[p] != 0xff) throw fail("marker expected");
        const int m = buf[p + 1];
        p += 2;
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01 || m == 0xff) { if (m == 0xff) --p; continue; }
        if (m == 0xd9) break;
        const int len = u16(p);
        if (len < 2 || p + len > buf.size()) throw fail("bad segment length");
        const size_t q = p + 2, end = p + len;
        if (m == 0xdb) {
            for (size_t a = q; a < end;) {
                const int pq = buf[a] >> 4, id = buf[a] & 15;
                ++a;
                if (id > 3) throw fail("bad quantisation table");
                for (int i = 0; i < 64; ++i) { qt[id][zz[i]] = pq ? (uint16_t)u16(a) : buf[a]; a += pq ? 2 : 1; }
            }
        } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {
            if (!comps.empty()) throw fail("more than one frame");
            progressive = m == 0xc2;
            if (buf[q] != 8) throw fail("only 8-bit samples");
            height = u16(q + 1); width = u16(q + 3);
            const int nc = buf[q + 5];
            if (nc != 1 && nc != 3) throw fail("unsupported component count " + std::to_string(nc));
            if (width <= 0 || height <= 0) throw fail("no image data");
            comps.resize(nc);
            for (int c = 0; c < nc; ++c) {
                comps[c].id = buf[q + 6 + 3 * c];
                comps[c].h = buf[q + 7 + 3 * c] >> 4; comps[c].v = buf[q + 7 + 3 * c] & 15;
                comps[c].tq = buf[q + 8 + 3 * c];
                if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2 || comps[c].tq > 3) throw fail("unsupported sampling");
            }
            if (nc == 1) comps[0].h = comps[0].v = 1;  // a single component is never interleaved: its sampling factors do not matter
            for (auto &c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
            mcux = (width + 8 * hmax - 1) / (8 * hmax); mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            // every block costs at least one bit in its first scan: more blocks than bits is a damaged header
            if ((uint64_t)mcux * (uint64_t)mcuy > 8ull * buf.size()) throw fail("image size does not fit the file");
            for (auto &c : comps) {
                c.nbx = mcux * c.h; c.nby = mcuy * c.v;
                c.bw = c.nbx * 8; c.bh = c.nby * 8;
                c.cbx = ((width * c.h + hmax - 1) / hmax + 7) / 8; c.cby = ((height * c.v + vmax - 1) / vmax + 7) / 8;
                c.coef.assign((size_t)c.nbx * c.nby * 64, 0);
            }
        } else if (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
            throw fail("lossless / hierarchical / arithmetic-coded JPEG is not supported");
        } else if (m == 0xc4) {
            for (size_t a = q; a < end;) {
                const int tc = buf[a] >> 4, id = buf[a] & 15;
                ++a;
                if (id > 3 || tc > 1 || a + 16 > end) throw fail("bad Huffman table");
                Huff &h = tc ? hac[id] : hdc[id];
                int n = 0;
                for (int l = 1; l <= 16; ++l) { h.bits[l] = buf[a + l - 1]; n += h.bits[l]; }
                a += 16;
                if (n > 256 || a + n > end) throw fail("bad Huffman table");
                for (int i = 0; i < n; ++i) h.vals[i] = buf[a + i];
                a += n;
                h.build();
            }
        } else if (m == 0xe0 && len >= 7) {  // APP0 "JFIF"
            if (std::memcmp(&buf[q], "JFIF\0", 5) == 0) jfif = true;
        } else if (m == 0xee && len >= 14) {  // APP14 "Adobe": colour transform flag
            if (std::memcmp(&buf[q], "Adobe\0", 6) == 0) adobe_transform = buf[q + 11];
        } else if (m == 0xdd) {
            restart = u16(q);
        } else if (m == 0xda) {
            if (comps.empty()) throw fail("scan before frame header");
            const int ns = buf[q];
            if (ns < 1 || ns > (int)comps.size() || q + 1 + 2 * (size_t)ns + 3 > end) throw fail("bad scan header");
            std::vector<Comp *> sc;
            for (int i = 0; i < ns; ++i) {
                const int cid = buf[q + 1 + 2 * i], t = buf[q + 2 + 2 * i];
                Comp *found = nullptr;
                for (auto &c : comps) if (c.id == cid) found = &c;
                if (!found || (t >> 4) > 3 || (t & 15) > 3) throw fail("bad scan header");
                found->td = t >> 4; found->ta = t & 15;
                sc.push_back(found);
            }
            int ss = buf[q + 1 + 2 * ns], se = buf[q + 2 + 2 * ns];
            const int ah = buf[q + 3 + 2 * ns] >> 4, al = buf[q + 3 + 2 * ns] & 15;
            if (!progressive) { ss = 0; se = 63; }
            else if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13 || ah > 13) throw fail("bad progressive scan parameters");
            bp = end;
            decode_scan(sc, ss, se, progressive ? ah : 0, progressive ? al : 0);
            ++scans;
            // next marker that is not a restart marker / stuffed byte
            size_t n = bp;
            while (n + 1 < buf.size() && !(buf[n] == 0xff && buf[n + 1] != 0 && buf[n + 1] != 0xff && !(buf[n + 1] >= 0xd0 && buf[n + 1] <= 0xd7))) ++n;
            if (n + 1 >= buf.size()) eoi = true;  // truncated file: keep what was decoded (stb_image does the same)
            p = n;
            continue;
        }
        p = end;
    }
    if (!scans || comps.empty()) throw fail("no image data");

    // ---- reconstruction, in the integer arithmetic of stb_image (the decoder the reference calls: lib/stb/stb/stb_image.h) so that a
    // texture decodes to the bytes the reference uploads.  Dequantisation in 16 bits (stbi__jpeg_dequantize), the jidctint-style
    // 12-bit fixed-point IDCT with 2 extra bits between the passes (stbi__idct_block), chroma rows / columns by the 3:1 and 9:3:3:1
    // filters with stb's edge rules (stbi__resample_row_*), 20-bit fixed-point YCbCr -> RGB (stbi__YCbCr_to_RGB_row).
    auto fx = [](double x) { return (int)(x * 4096 + 0.5); };  // stbi__f2f
    const int c0 = fx(0.5411961f), c1 = fx(-1.847759065f), c2 = fx(0.765366865f), c3 = fx(1.175875602f), c4 = fx(0.298631336f), c5 = fx(2.053119869f),
              c6 = fx(3.072711026f), c7 = fx(1.501321110f), c8 = fx(-0.899976223f), c9 = fx(-2.562915447f), c10 = fx(-1.961570560f), c11 = fx(-0.390180644f);
    struct Pass { int x0, x1, x2, x3, t0, t1, t2, t3; };
    auto idct_1d = [&](int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {  // even part -> x*, odd part -> t*
        Pass r;
        int p1 = (s2 + s6) * c0;
        int e2 = p1 + s6 * c1, e3 = p1 + s2 * c2;
        int e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
        r.x0 = e0 + e3; r.x3 = e0 - e3; r.x1 = e1 + e2; r.x2 = e1 - e2;
        int t0 = s7, t1 = s5, t2 = s3, t3 = s1;
        int p3 = t0 + t2, p4 = t1 + t3, q1 = t0 + t3, q2 = t1 + t2;
        const int p5 = (p3 + p4) * c3;
        t0 *= c4; t1 *= c5; t2 *= c6; t3 *= c7;
        q1 = p5 + q1 * c8; q2 = p5 + q2 * c9; p3 *= c10; p4 *= c11;
        r.t3 = t3 + q1 + p4; r.t2 = t2 + q2 + p3; r.t1 = t1 + q2 + p4; r.t0 = t0 + q1 + p3;
        return r;
    };
    auto clamp8 = [](int x) { return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x); };
    for (auto &c : comps) {
        c.px.assign((size_t)c.bw * c.bh, 0);
        for (int bj = 0; bj < c.nby; ++bj)
            for (int bi = 0; bi < c.nbx; ++bi) {
                const int16_t *q = &c.coef[((size_t)bj * c.nbx + bi) * 64];
                int16_t d[64];
                for (int i = 0; i < 64; ++i) d[i] = (int16_t)(q[i] * qt[c.tq][i]);
                int v[64];
                for (int i = 0; i < 8; ++i) {  // columns
                    if (!(d[8 + i] | d[16 + i] | d[24 + i] | d[32 + i] | d[40 + i] | d[48 + i] | d[56 + i])) {
                        for (int k = 0; k < 8; ++k) v[8 * k + i] = d[i] * 4;
                        continue;
                    }
                    Pass r = idct_1d(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i]);
                    r.x0 += 512; r.x1 += 512; r.x2 += 512; r.x3 += 512;
                    v[i] = (r.x0 + r.t3) >> 10; v[56 + i] = (r.x0 - r.t3) >> 10;
                    v[8 + i] = (r.x1 + r.t2) >> 10; v[48 + i] = (r.x1 - r.t2) >> 10;
                    v[16 + i] = (r.x2 + r.t1) >> 10; v[40 + i] = (r.x2 - r.t1) >> 10;
                    v[24 + i] = (r.x3 + r.t0) >> 10; v[32 + i] = (r.x3 - r.t0) >> 10;
                }
                for (int y = 0; y < 8; ++y) {  // rows: >> 17 removes 12 + 2 + 3 bits; + 128 level shift, rounded
                    const int *w = &v[8 * y];
                    Pass r = idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
                    const int bias = 65536 + (128 << 17);
                    r.x0 += bias; r.x1 += bias; r.x2 += bias; r.x3 += bias;
                    uint8_t *o = &c.px[(size_t)(bj * 8 + y) * c.bw + bi * 8];
                    o[0] = clamp8((r.x0 + r.t3) >> 17); o[7] = clamp8((r.x0 - r.t3) >> 17);
                    o[1] = clamp8((r.x1 + r.t2) >> 17); o[6] = clamp8((r.x1 - r.t2) >> 17);
                    o[2] = clamp8((r.x2 + r.t1) >> 17); o[5] = clamp8((r.x2 - r.t1) >> 17);
                    o[3] = clamp8((r.x3 + r.t0) >> 17); o[4] = clamp8((r.x3 - r.t0) >> 17);
                }
            }
        std::vector<int16_t>().swap(c.coef);
    }
    // One output row at a time: every component's row is brought to full width from its near / far source rows.
    // Vertical 2x: output row j reads source row j >> 1 ("near") and the row above (even j) or below (odd j), held at the first row
    // and at the component's own last row.  Horizontal 2x over w = ceil(width / 2) samples.
    std::vector<std::vector<uint8_t>> line(comps.size(), std::vector<uint8_t>((size_t)width + 8));
    LoadedImage img;
    img.width = width; img.height = height;
    img.channels = comps.size() == 1 ? 1 : 3;
    img.u8.resize((size_t)width * height * img.channels);
    // an RGB file: component ids 'R' 'G' 'B', or an Adobe APP14 marker with transform 0 and no JFIF header (load_jpeg_image's is_rgb)
    const bool is_rgb = comps.size() == 3 && ((comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') || (adobe_transform == 0 && !jfif));
    for (int j = 0; j < height; ++j) {
        for (size_t k = 0; k < comps.size(); ++k) {
            const Comp &c = comps[k];
            const int hs = hmax / c.h, vs = vmax / c.v;
            const int w = (width + hs - 1) / hs, rows = (height * c.v + vmax - 1) / vmax;
            int rn = j, rf = j;
            if (vs == 2) {
                rn = j >> 1;
                rf = (j & 1) ? std::min(rn + 1, rows - 1) : std::max(rn - 1, 0);
                // stb advances its far pointer only while rows remain: beyond the last row near and far coincide
                if (rn > rows - 1) rn = rows - 1;
            }
            const uint8_t *near_ = &c.px[(size_t)rn * c.bw], *far_ = &c.px[(size_t)rf * c.bw];
            uint8_t *out = line[k].data();
            if (hs == 1 && vs == 1) {
                std::memcpy(out, near_, (size_t)w);
            } else if (hs == 1) {  // stbi__resample_row_v_2
                for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * near_[i] + far_[i] + 2) >> 2);
            } else if (vs == 1) {  // stbi__resample_row_h_2, including its asymmetric last pair
                if (w == 1) { out[0] = out[1] = near_[0]; continue; }
                out[0] = near_[0];
                out[1] = (uint8_t)((near_[0] * 3 + near_[1] + 2) >> 2);
                for (int i = 1; i < w - 1; ++i) {
                    const int n3 = 3 * near_[i] + 2;
                    out[2 * i] = (uint8_t)((n3 + near_[i - 1]) >> 2);
                    out[2 * i + 1] = (uint8_t)((n3 + near_[i + 1]) >> 2);
                }
                out[2 * (w - 1)] = (uint8_t)((near_[w - 2] * 3 + near_[w - 1] + 2) >> 2);
                out[2 * (w - 1) + 1] = near_[w - 1];
            } else {  // stbi__resample_row_hv_2
                if (w == 1) { out[0] = out[1] = (uint8_t)((3 * near_[0] + far_[0] + 2) >> 2); continue; }
                int t1 = 3 * near_[0] + far_[0];
                out[0] = (uint8_t)((t1 + 2) >> 2);
                for (int i = 1; i < w; ++i) {
                    const int t0 = t1;
                    t1 = 3 * near_[i] + far_[i];
                    out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
                    out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
                }
                out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
            }
        }
        uint8_t *o = &img.u8[(size_t)j * width * img.channels];
        if (img.channels == 1) { std::memcpy(o, line[0].data(), (size_t)width); continue; }
        for (int x = 0; x < width; ++x, o += 3) {
            if (is_rgb) { o[0] = line[0][x]; o[1] = line[1][x]; o[2] = line[2][x]; continue; }
            auto f2f = [](float v) { return ((int)(v * 4096.0f + 0.5f)) << 8; };  // stbi__float2fixed
            const int y_fixed = (line[0][x] << 20) + (1 << 19);
            const int cr = line[2][x] - 128, cb = line[1][x] - 128;
            int r = y_fixed + cr * f2f(1.40200f);
            int g = y_fixed + (cr * -f2f(0.71414f)) + ((cb * -f2f(0.34414f)) & 0xffff0000);
            int bl = y_fixed + cb * f2f(1.77200f);
            r >>= 20; g >>= 20; bl >>= 20;
            o[0] = clamp8(r); o[1] = clamp8(g); o[2] = clamp8(bl);
        }
    }
    return img;
}

LoadedImage load_image(const std::string &filename) {
    const std::string ext = lower_ext(filename);
    if (ext == "png" || ext == "jpg" || ext == "jpeg") {
        LoadedImage img = ext == "png" ? load_png(filename) : load_jpg(filename);
        // utils/image.cpp:313-335: one, three or four channels become Image<uint8_t / Vec3b / Vec4b>; grey + alpha is refused
        if (img.channels != 1 && img.channels != 3 && img.channels != 4) throw Error("ImageDynamic::load: unsupported channels: " + std::to_string(img.channels));
        return img;
    }
    if (ext == "hdr") return load_hdr(filename);
    throw Error("unsupported image format (use .png, .jpg or .hdr): " + filename);
}

}  // namespace btrc
