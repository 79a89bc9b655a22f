// rfm_rng.cuh -- random streams of the SGD kernel.
//
//  * Philox4x32-10: counter-based generator for production-mode negative sampling.  Replaces the reference's
//    global MT19937 stream (rankfm/_rankfm.pyx:182,251 -> mt19937ar.c:105-140) with a stream every warp can index
//    independently: counter = (row.lo, epoch, attempt/4, row.hi), key = seed; attempt a uses word a%4.
//  * Feistel permutation: on-device replacement of the per-epoch np.random.shuffle (rankfm/_rankfm.pyx:227).
//  * MT19937 (warp-cooperative twist, state in shared memory): only for the serial replay mode, where the kernel
//    has to consume exactly the stream the reference consumes (init_genrand(1492), genrand_int32() % I).
//
// The Philox/Feistel definitions are a contract shared with oracle/rankfm_oracle.c (orc_philox4x32,
// orc_feistel_perm); tests/test_rng_contract.py holds both sides to it.
#pragma once
#include <cstdint>

namespace rfm {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// Per-epoch permutation of [0,N): 4-round balanced Feistel network + cycle walking.  The round keys and the bit
// split are computed on the host once per epoch (make_feistel) and passed by value.
struct Feistel {
    uint32_t key[4];
    uint32_t mask;
    int32_t  half;
    int64_t  n;
};

inline Feistel make_feistel(int64_t n, uint64_t seed, int epoch)
{
    Feistel f;
    int bits = 2;
    while (((int64_t)1 << bits) < n) ++bits;
    if (bits & 1) ++bits;
    f.half = bits / 2;
    f.mask = (uint32_t)(((uint64_t)1 << f.half) - 1);
    f.n = n;
    for (int k = 0; k < 4; ++k)
        f.key[k] = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(epoch + 1) + 0x85EBCA6Bu * (uint32_t)(k + 1)));
    return f;
}

__host__ __device__ __forceinline__ int64_t feistel_perm(const Feistel& f, int64_t r)
{
    uint64_t x = (uint64_t)r;
    do {
        uint32_t L = (uint32_t)(x >> f.half), R = (uint32_t)x & f.mask;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = L ^ (mix32(R ^ f.key[k]) & f.mask);
            L = R; R = t;
        }
        x = ((uint64_t)L << f.half) | R;
    } while ((int64_t)x >= f.n);
    return (int64_t)x;
}

// ---------------------------------------------------------------------------------------------------------------
// MT19937, one warp.  State: 624 words + cursor in shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMtN = 624;
constexpr int kMtM = 397;

struct MtState {
    uint32_t s[kMtN];
    int pos;
};

#ifdef __CUDACC__
// whole warp calls; sequential recurrence is done by lane 0 (runs once per fit)
__device__ inline void mt_seed_warp(MtState* g, uint32_t seed)
{
    if ((threadIdx.x & 31) == 0) {
        uint32_t prev = seed;
        g->s[0] = prev;
        for (int k = 1; k < kMtN; ++k) {
            prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)k;
            g->s[k] = prev;
        }
        g->pos = kMtN;
    }
    __syncwarp();
}

// regenerate all 624 words; 32 lanes per batch, read-then-write so every lane sees the values the sequential
// recurrence would (s[k+1] old, s[k+397 mod 624] old for k<227 and new afterwards)
__device__ inline void mt_twist_warp(MtState* g)
{
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < kMtN; base += 32) {
        const int k = base + lane;
        uint32_t v = 0;
        if (k < kMtN) {
            const uint32_t y = (g->s[k] & 0x80000000u) | (g->s[(k + 1) % kMtN] & 0x7fffffffu);
            v = g->s[(k + kMtM) % kMtN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (k < kMtN) g->s[k] = v;
        __syncwarp();
    }
    if (lane == 0) g->pos = 0;
    __syncwarp();
}

// next tempered word, uniform across the warp (all lanes read the same slot)
__device__ inline uint32_t mt_next_warp(MtState* g)
{
    if (g->pos >= kMtN) mt_twist_warp(g);
    uint32_t y = g->s[g->pos];
    __syncwarp();
    if ((threadIdx.x & 31) == 0) g->pos = g->pos + 1;
    __syncwarp();
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
#endif

}  // namespace rfm
