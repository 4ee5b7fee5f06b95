// ORACLE / reference build (test infrastructure only).
//
// Prelude for the C++ that the REFERENCE'S OWN code generator emits (lib/cuj/cuj/src/gen/cpp.cpp, CPPCodeGenerator with
// Target::PTX, i.e. "CUDA C++" text).  The reference records its kernels (generate / medium / shade, materials, media, lights,
// filters, camera ...) into a cuj::Module; oracle/refcpu/shim_ptxgen.cpp turns that module into C++ text with the reference's
// generator and compiles it with g++ for the HOST.  This header gives that text the handful of CUDA names it mentions
// (cpp_prefix.inl:369-768): qualifiers, vector types, thread indices, atomics, rsqrtf/exp10f, tex2D/tex3D.
// Nothing here is reference source; nothing here is used by the product.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define __global__
#define __device__

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float3 make_float3(float a, float b, float c) { return {a, b, c}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int3 make_int3(int a, int b, int c) { return {a, b, c}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline uint3 make_uint3(unsigned a, unsigned b, unsigned c) { return {a, b, c}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }

struct refcpu_dim3 { int x, y, z; };
static thread_local refcpu_dim3 threadIdx, blockIdx, blockDim;

static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline double rsqrt(double a) { return 1.0 / sqrt(a); }
#ifndef _GNU_SOURCE
static inline float exp10f(float a) { return powf(10.0f, a); }
static inline double exp10(double a) { return pow(10.0, a); }
#endif

// relaxed atomics on plain memory (the launcher may run blocks on several host threads)
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float *p, float v) {
    unsigned *u = reinterpret_cast<unsigned *>(p);
    unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        float f; memcpy(&f, &old, 4);
        float n = f + v;
        unsigned nu; memcpy(&nu, &n, 4);
        if (__atomic_compare_exchange_n(u, &old, nu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
    }
}
template <typename T> static inline T atomicCAS(T *p, T cmp, T val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return cmp;
}

static inline void __assertfail(const char *message, const char *file, unsigned line, const char *function, size_t) {
    fprintf(stderr, "device assertion failed. file: %s, line: %u, func: %s, message: %s\n", file, line, function, message);
    abort();
}

// texture objects live in the harness (oracle/refcpu/shim_cuda.cpp); the harness hands its sampler in after dlopen
struct refcpu_callbacks {
    void (*tex2d)(unsigned long long tex, float u, float v, float out[4]);
    void (*tex3d)(unsigned long long tex, float u, float v, float w, float out[4]);
};
static refcpu_callbacks refcpu_cb;
extern "C" void refcpu_set_callbacks(const refcpu_callbacks *cb) { refcpu_cb = *cb; }

template <typename T> static inline T tex2D(unsigned long long tex, float u, float v);
template <> inline float4 tex2D<float4>(unsigned long long tex, float u, float v) {
    float o[4]; refcpu_cb.tex2d(tex, u, v, o); return {o[0], o[1], o[2], o[3]};
}
template <> inline int4 tex2D<int4>(unsigned long long, float, float) { abort(); }
template <typename T> static inline T tex3D(unsigned long long tex, float u, float v, float w);
template <> inline float4 tex3D<float4>(unsigned long long tex, float u, float v, float w) {
    float o[4]; refcpu_cb.tex3d(tex, u, v, w, o); return {o[0], o[1], o[2], o[3]};
}
template <> inline int4 tex3D<int4>(unsigned long long, float, float, float) { abort(); }

// per-thread hook of the sequential launch mode (stage-level parity runs read counters / film after every thread)
typedef void (*refcpu_thread_hook)(int linear_thread, int phase, void *user);
