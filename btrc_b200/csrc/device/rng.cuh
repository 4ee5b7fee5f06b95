// Per-path random numbers: the reference's IndependentSampler, draw for draw.
//   builtin/sampler/independent.cpp:11-25   seed = PCG(murmur64A(pixel.x, pixel.y)), advance(sample_index * 65536)
//   lib/cuj/cuj/src/cstd/random.cpp:113-161 PCG32 (XSH-RR), LCG skip-ahead, uniform_float
//   utils/hash.h:12-99                      MurmurHash64A over the 8 bytes (x:u32, y:u32), seed 0
// Integer arithmetic: bit-exact against the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace btrc {

constexpr uint64_t kPcgDefaultState = 0x853c49e6748fea9bULL;
constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dULL;

struct Pcg {
    uint64_t state, inc;

    __device__ __forceinline__ uint32_t next() {
        uint64_t old = state;
        state = old * kPcgMult + inc;
        uint32_t xs = uint32_t(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = uint32_t(old >> 59u);
        return __funnelshift_r(xs, xs, rot);  // rotate right
    }
    __device__ __forceinline__ void seed(uint64_t s) {
        state = 0;
        inc = (s << 1u) | 1u;
        next();
        state += kPcgDefaultState;
        next();
    }
    __device__ __forceinline__ void advance(uint64_t delta) {
        uint64_t cur_mult = kPcgMult, cur_plus = inc, acc_mult = 1, acc_plus = 0;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta >>= 1;
        }
        state = acc_mult * state + acc_plus;
    }
    // min(f32(u32) * 2^-32, 0x1.fffffep-1)
    __device__ __forceinline__ float uniform() { return fminf(__uint2float_rn(next()) * 0x1p-32f, 0x1.fffffep-1f); }
};

// MurmurHash64A specialised for one 8-byte block, seed 0 (hash.h:12-70 with len = 8)
__device__ __forceinline__ uint64_t murmur_pixel(uint32_t px, uint32_t py) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = 0 ^ (8ull * m);
    uint64_t k = uint64_t(px) | (uint64_t(py) << 32);
    k *= m; k ^= k >> r; k *= m;
    h ^= k; h *= m;
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}

struct Sampler {
    Pcg rng;
    __device__ __forceinline__ void init(uint32_t px, uint32_t py, int sample_index) {
        rng.seed(murmur_pixel(px, py));
        rng.advance(uint64_t(int64_t(sample_index)) * 65536ull);
    }
    __device__ __forceinline__ float get1d() { return rng.uniform(); }
};

}  // namespace btrc
