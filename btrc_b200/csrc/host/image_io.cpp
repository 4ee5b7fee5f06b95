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

uint32_t get_u32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

LoadedImage load_png(const std::string &filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw Error("failed to open file: " + filename);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) throw Error("failed to load image: " + filename);
    size_t p = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    while (p + 8 <= buf.size()) {
        const uint32_t len = get_u32(&buf[p]);
        const std::string type(reinterpret_cast<const char *>(&buf[p + 4]), 4);
        const uint8_t *d = &buf[p + 8];
        if (p + 12 + len > buf.size()) throw Error("failed to load image: " + filename);
        if (type == "IHDR") { w = (int)get_u32(d); h = (int)get_u32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        p += 12 + len;
    }
    if (interlace != 0 || depth != 8) throw Error("unsupported png (only 8-bit non-interlaced): " + filename);
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw Error("failed to load image: " + filename);
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size())
        throw Error("failed to load image: " + filename);
    std::vector<uint8_t> px(stride * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *src = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &px[stride * y];
        const uint8_t *up = y ? &px[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = src[x];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) / 2; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            dst[x] = (uint8_t)v;
        }
    }
    LoadedImage img;
    img.width = w; img.height = h;
    if (ctype == 3) {  // palette -> rgb
        img.channels = 3;
        img.u8.resize((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; ++i)
            for (int k = 0; k < 3; ++k) img.u8[3 * i + k] = 3 * (size_t)px[i] + k < plte.size() ? plte[3 * px[i] + k] : 0;
    } else {
        img.channels = ch;
        img.u8 = std::move(px);
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
    } else if (ext == "jpg" || ext == "jpeg" || ext == "exr") {
        throw Error("unsupported image format (use .png or .hdr): " + filename);
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

LoadedImage load_image(const std::string &filename) {
    const std::string ext = lower_ext(filename);
    if (ext == "png") return load_png(filename);
    if (ext == "hdr") return load_hdr(filename);
    throw Error("unsupported image format (use .png or .hdr): " + filename);
}

}  // namespace btrc
