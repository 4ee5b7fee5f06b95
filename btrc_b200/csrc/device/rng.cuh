// Per-path random numbers: the reference's IndependentSampler, draw for draw.
//   builtin/sampler/independent.cpp:11-25   seed = PCG(murmur64A(pixel.x, pixel.y)), advance(sample_index * 65536)
//   lib/cuj/cuj/src/cstd/random.cpp:113-161 PCG32 (XSH-RR), LCG skip-ahead, uniform_float
//   utils/hash.h:12-99                      MurmurHash64A over the 8 bytes (x:u32, y:u32), seed 0
// Integer arithmetic: bit-exact against the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../common/types.h"

namespace btrc {

constexpr uint64_t kPcgDefaultState = 0x853c49e6748fea9bULL;
constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dULL;

// LCG skip-ahead constants for 2^j steps at once: the step map is x -> M x + c; 2^j steps are x -> M^(2^j) x + c S_j with
// S_0 = 1, S_(j+1) = (M^(2^j) + 1) S_j (all mod 2^64) -- what the loop of Pcg::advance reaches after j halvings of delta
constexpr uint64_t pcg_mult_pow2(int j) { uint64_t m = 0x5851f42d4c957f2dULL; for (int i = 0; i < j; ++i) m *= m; return m; }
constexpr uint64_t pcg_plus_sum2(int j) { uint64_t m = 0x5851f42d4c957f2dULL, s = 1; for (int i = 0; i < j; ++i) { s = (m + 1) * s; m *= m; } return s; }

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
    // advance(blocks << 16) (independent.cpp:11-15 advances by sample_index * 65536).  2^(16+j) steps of the generator are the map
    // x -> M_j x + S_j inc with compile-time M_j, S_j; the maps of the set bits of `blocks` are applied one after the other (they
    // commute: powers of one affine map) -- two 64-bit multiplies per SET bit instead of the loop above's two per bit for squaring the
    // map plus three per set bit.  All arithmetic is exact mod 2^64, so the state is bit-identical to advance(blocks << 16).
    template <int J>
    __device__ __forceinline__ void skip_bit(uint64_t blocks) {
        constexpr uint64_t m = pcg_mult_pow2(16 + J), s = pcg_plus_sum2(16 + J);
        if ((blocks >> J) & 1) state = m * state + s * inc;
    }
    __device__ __forceinline__ void advance_blocks_of_65536(uint64_t blocks) {
        skip_bit<0>(blocks); skip_bit<1>(blocks); skip_bit<2>(blocks); skip_bit<3>(blocks); skip_bit<4>(blocks);
        if (blocks >> 5) {
            skip_bit<5>(blocks); skip_bit<6>(blocks); skip_bit<7>(blocks); skip_bit<8>(blocks); skip_bit<9>(blocks);
            if (blocks >> 10) {
                skip_bit<10>(blocks); skip_bit<11>(blocks); skip_bit<12>(blocks); skip_bit<13>(blocks); skip_bit<14>(blocks); skip_bit<15>(blocks);
                constexpr uint64_t kM32 = pcg_mult_pow2(32), kS32 = pcg_plus_sum2(32);
                if (blocks >> 16) advance_blocks_loop(blocks >> 16, kM32, kS32);
            }
        }
    }
    // the generic loop, starting from the map of 2^k steps (mult0 = M^(2^k), sum0 = S_k)
    __device__ __noinline__ void advance_blocks_loop(uint64_t blocks, uint64_t mult0, uint64_t sum0) {
        uint64_t cur_mult = mult0, cur_plus = sum0 * inc, acc_mult = 1, acc_plus = 0;
        while (blocks > 0) {
            if (blocks & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            blocks >>= 1;
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

// ---------------------------------------------------------------------------------------------------------------
// Halton sampler with Owen scrambling (builtin/sampler/halton.cpp:9-133, utils/low_discrepancy.cpp:10-120, after pbrt-v4).
// The reference wires its wavefront renderer to IndependentSampler at compile time (wavefront/common.h:36-37); here the
// renderer node's "sampler": "halton" selects this one (kernels are instantiated for both).  The state shares the PCG's
// 16-byte queue slot: state = halton index, inc = dimension << 1 (a PCG increment is always odd).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPrimeTableSize = 1000;     // utils/prime.h:9
constexpr int kMaxHaltonResolution = 128;  // halton.cpp:10

__device__ __forceinline__ uint64_t mix_bits(uint64_t v) {  // low_discrepancy.cpp:47-55
    v ^= v >> 31; v *= 0x7fb5d329728ea185ull;
    v ^= v >> 27; v *= 0x81dadef4bc2dd44dull;
    v ^= v >> 33;
    return v;
}

__device__ __forceinline__ int permutation_element(uint32_t i, uint32_t l, uint32_t p) {  // low_discrepancy.cpp:10-43
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return int((i + p) % l);
}

__device__ __noinline__ float owen_scrambled_radical_inverse(const int *primes, int base_index, uint64_t a, uint64_t hash) {  // low_discrepancy.cpp:82-103
    const uint64_t base = (uint64_t)__ldg(primes + base_index);
    const float inv_base = __fdiv_rn(1.0f, float(base));
    float inv_base_m = 1.0f;
    uint64_t reversed_digits = 0;
    while (__fsub_rn(1.0f, inv_base_m) < 1.0f) {
        const uint64_t next = a / base;
        uint64_t digit_value = a - next * base;
        const uint64_t digit_hash = mix_bits(hash ^ reversed_digits);
        digit_value = (uint64_t)permutation_element(uint32_t(digit_value), uint32_t(base), uint32_t(digit_hash));
        reversed_digits = reversed_digits * base + digit_value;
        inv_base_m = __fmul_rn(inv_base_m, inv_base);
        a = next;
    }
    return fminf(__fmul_rn(inv_base_m, __ull2float_rn(reversed_digits)), 0x1.fffffep-1f);
}

__device__ __forceinline__ uint64_t inverse_radical_inverse(uint64_t inverse, int base, int digits) {  // low_discrepancy.cpp:105-117
    uint64_t index = 0;
    for (int i = 0; i < digits; ++i) {
        const uint64_t digit = inverse % (uint64_t)base;
        inverse = inverse / (uint64_t)base;
        index = index * (uint64_t)base + digit;
    }
    return index;
}

template <bool HALTON>
struct SamplerT;

template <>
struct SamplerT<false> {
    Pcg rng;
    __device__ __forceinline__ void init(const HaltonConsts &, uint32_t px, uint32_t py, int sample_index) {
        rng.seed(murmur_pixel(px, py));
        // advance(sample_index * 65536); a negative index (never generated) would wrap like the reference's i64 product
        if (sample_index >= 0 && sample_index < (1 << 30)) rng.advance_blocks_of_65536(uint64_t(sample_index));
        else rng.advance(uint64_t(int64_t(sample_index)) * 65536ull);
    }
    __device__ __forceinline__ void load(ulonglong2 v, const HaltonConsts &) { rng.state = v.x; rng.inc = v.y; }
    __device__ __forceinline__ ulonglong2 store() const { return make_ulonglong2(rng.state, rng.inc); }
    __device__ __forceinline__ float get1d() { return rng.uniform(); }
    __device__ __forceinline__ void begin3d() {}  // get3d = three get1d (core/sampler.h:22-28); see SamplerT<true>
};

template <>
struct SamplerT<true> {
    uint64_t halton_index;
    uint32_t dimension;
    const int *primes;
    __device__ __forceinline__ void init(const HaltonConsts &hc, uint32_t px, uint32_t py, int sample_index) {  // halton.cpp:46-67
        primes = hc.primes;
        halton_index = 0;
        const int sample_stride = hc.base_scales[0] * hc.base_scales[1];
        if (sample_stride > 1) {
            const uint32_t pm[2] = {px % kMaxHaltonResolution, py % kMaxHaltonResolution};
            for (int i = 0; i < 2; ++i) {
                const uint64_t dim_offset = inverse_radical_inverse(uint64_t(pm[i]), i == 0 ? 2 : 3, hc.base_exponents[i]);
                halton_index = halton_index + dim_offset * uint64_t(sample_stride / hc.base_scales[i]) * uint64_t(hc.mult_inverse[i]);
            }
            halton_index = halton_index % uint64_t(sample_stride);
        }
        halton_index = halton_index + uint64_t(int64_t(sample_index * sample_stride));
        dimension = 2;  // the first two dimensions are left to the film position (filter importance sampling takes over)
    }
    __device__ __forceinline__ void load(ulonglong2 v, const HaltonConsts &hc) { halton_index = v.x; dimension = uint32_t(v.y >> 1); primes = hc.primes; }
    __device__ __forceinline__ ulonglong2 store() const { return make_ulonglong2(halton_index, uint64_t(dimension) << 1); }
    // HaltonSampler::get3d (halton.cpp:106-118) wraps the dimension counter BEFORE the triple when it would run past the prime
    // table, so that the three values never straddle the wrap; call this before three consecutive get1d() that stand for a get3d
    __device__ __forceinline__ void begin3d() {
        if (dimension + 2 >= (uint32_t)kPrimeTableSize) dimension = 2;
    }
    __device__ __forceinline__ float get1d() {  // halton.cpp:80-89
        if (dimension >= (uint32_t)kPrimeTableSize) dimension = 2;
        const float r = owen_scrambled_radical_inverse(primes, int(dimension), halton_index, mix_bits(uint64_t(1u + (dimension << 4))));
        dimension = dimension + 1;
        return r;
    }
};

using Sampler = SamplerT<false>;  // shadow-ray transmittance streams stay PCG in both modes (keyed on the ray, shading.cuh)

}  // namespace btrc
