// ORACLE build shim (test infrastructure): utils/image.cpp calls the MSVC-only strcpy_s(dst, size, src) for three one-letter EXR
// channel names; force-included (-include) so that the source compiles unmodified with gcc.
#pragma once
#include <cstring>
static inline int strcpy_s(char *dst, size_t size, const char *src) {
    if (!dst || !src || size == 0) return 22;
    std::strncpy(dst, src, size - 1);
    dst[size - 1] = 0;
    return 0;
}
