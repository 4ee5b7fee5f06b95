// Image files: the "save" post-processor's writers and the texture readers.
//
// Behaviour mirrors utils/image.h:207-224,495-548 and utils/image.cpp:16-114: format inferred from the extension;
// PNG is written as 8-bit RGBA with u8 = min(int(x * 256), 255); HDR as Radiance RGBE.  The reference uses the
// vendored stb / tinyexr; here PNG goes through zlib directly and JPG / EXR are reported as unsupported.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace btrc {

// data: width*height float4 (row 0 = top)
void save_image_rgba32f(const std::string &filename, int width, int height, const float *data);

struct LoadedImage {
    int width = 0, height = 0, channels = 0;
    bool is_float = false;
    std::vector<uint8_t> u8;   // 8-bit images: width*height*channels
    std::vector<float> f32;    // float images (hdr): width*height*3
};
LoadedImage load_image(const std::string &filename);

}  // namespace btrc
