// ORACLE build (test infrastructure): C entry points over the reference's OWN image layer - utils/image.cpp with the vendored
// stb_image / stb_image_write / tinyexr, compiled where they lie into oracle/_ref/libbtrcref.so.  tests/test_reference_host_pins_cpu.py
// uses them to pin the product's texture readers (PNG / JPEG / HDR) and the `save` post-processor's writers.  Nothing here is shipped.
#include <cstring>
#include <string>
#include <type_traits>
#include <variant>

#include <btrc/utils/image.h>
#include <tinyexr.h>

using namespace btrc;

static thread_local std::string g_img_err;

extern "C" {

const char *ref_image_last_error() { return g_img_err.c_str(); }

// ImageDynamic::load (utils/image.cpp:302-349).  First call with out = NULL for the size.  8-bit images: width*height*channels bytes
// (channels 1, 3 or 4); .hdr: width*height*3 floats, *is_float = 1.  Returns 0, or -1 (ref_image_last_error()).
int ref_image_load(const char *path, int *width, int *height, int *channels, int *is_float, void *out, unsigned long long out_bytes) {
    try {
        const ImageDynamic img = ImageDynamic::load(path);
        int c = 0, f = 0;
        const void *data = nullptr;
        size_t bytes = 0;
        const size_t n = (size_t)img.width() * img.height();
        img.match(   // the visitor form utils/cuda/array.cpp:166-172 uses
            [](std::monostate) {},
            [&](const auto &im) {
                using Texel = typename std::remove_cvref_t<decltype(im)>::Texel;
                data = im.data();
                if constexpr (std::is_same_v<Texel, uint8_t>) { c = 1; bytes = n; }
                else if constexpr (std::is_same_v<Texel, Vec3b>) { c = 3; bytes = 3 * n; }
                else if constexpr (std::is_same_v<Texel, Vec4b>) { c = 4; bytes = 4 * n; }
                else if constexpr (std::is_same_v<Texel, Vec3f>) { c = 3; f = 1; bytes = 12 * n; }
                else data = nullptr;
            });
        if (!data) { g_img_err = "unexpected texel type"; return -1; }
        if (width) *width = img.width();
        if (height) *height = img.height();
        if (channels) *channels = c;
        if (is_float) *is_float = f;
        if (!out) return 0;
        if (out_bytes < bytes) { g_img_err = "buffer too small"; return -1; }
        std::memcpy(out, data, bytes);
        return 0;
    } catch (const std::exception &e) { g_img_err = e.what(); return -1; }
}

// SaveToImage::process for one output (postprocess/save_to_image.cpp:30-62): Image<Vec4f> from rgba, pow_(1 / gamma) when
// apply_gamma (the colour output; albedo / normal are saved as they are), save(filename) with the format taken from the extension.
int ref_image_save(const char *path, int width, int height, const float *rgba, int apply_gamma, float gamma) {
    try {
        Image<Vec4f> image(width, height);
        std::memcpy(image.data(), rgba, sizeof(Vec4f) * (size_t)width * height);
        if (apply_gamma) image.pow_(1 / gamma);
        image.save(path);
        return 0;
    } catch (const std::exception &e) { g_img_err = e.what(); return -1; }
}

}  // extern "C"

// tinyexr's deflate: the vendored miniz (lib/tinyexr/deps/miniz/miniz.c), compiled where it lies as part of this translation unit
// (a .c file on the g++ command line would be compiled as C with the C++-only force-includes of the `ref` recipe).
#include <miniz.c>

// The reference cannot *read* EXR through ImageDynamic::load (utils/image.cpp:344), but its vendored tinyexr can: used to compare
// what the reference's save_exr wrote with what the product's EXR writer wrote.  out: width*height*4 floats (RGBA).
extern "C" int ref_exr_load(const char *path, int *width, int *height, float *out, unsigned long long out_floats) {
    float *rgba = nullptr;
    const char *err = nullptr;
    int w = 0, h = 0;
    if (LoadEXR(&rgba, &w, &h, path, &err) != TINYEXR_SUCCESS) {
        g_img_err = err ? err : "LoadEXR failed";
        if (err) FreeEXRErrorMessage(err);
        return -1;
    }
    if (width) *width = w;
    if (height) *height = h;
    int rc = 0;
    if (out) {
        if (out_floats < 4ull * w * h) { g_img_err = "buffer too small"; rc = -1; }
        else std::memcpy(out, rgba, sizeof(float) * 4 * (size_t)w * h);
    }
    free(rgba);
    return rc;
}
