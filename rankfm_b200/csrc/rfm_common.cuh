// rfm_common.cuh -- HBM data layout and the lane-group primitives shared by the SGD, predict and recommend kernels.
//
// HBM layout ("fat rows": everything one entity needs arrives with ONE contiguous, 16-byte aligned gather)
//   user table UT [U, ldu]   row = [ v_u[u, 0..F) padded to Fp | x_uf[u, 0..P) padded to Pp ]
//   item table IT [I, ldi]   row = [ v_i[i, 0..F) padded to Fp | w_i[i], 0, 0, 0 | x_if[i, 0..Q) padded to Qp ]
//   globals    GP            [ w_if padded to Qp | v_uf [P, Fp] | v_if [Q, Fp] ]
// Fp/Pp/Qp are multiples of 4 floats so every row is a whole number of float4 "quads"; pads are zero and stay zero
// (every update of a pad lane is eta*(c*0 - reg*0) = 0).  Pp / Qp are 0 when the feature block is absent or all
// zero (the reference's x_uf_any / x_if_any, rankfm/_rankfm.pyx:193-194).
//
// Work decomposition: a "group" of G lanes (G = 4, 8, 16 or 32, a power of two chosen so that 4*G*QPL >= the widest
// row section) owns one (user, item) pair; lane `sub` of the group owns quads sub, sub+G, ... of every section.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rfm {

struct Tables {
    float *UT, *IT, *GP;
    int32_t U, I, F, Fp, NQ;     // NQ = Fp/4 factor quads
    int32_t u0, Un;              // the user table holds the rows of users [u0, u0+Un) only (multi-GPU: this rank's users); UT is the
                                 // VIRTUAL base of the full table, so UT + u*ldu is the row of global user u for every owned u
    int32_t P, Pp, Q, Qp;        // Pp/Qp = 0 when that feature block is inactive
    int32_t ldu, ldi;            // row strides in floats
    int32_t x_uf_any, x_if_any;
    int32_t gp_vuf, gp_vif;      // float offsets into GP (w_if sits at 0)
};

__device__ __forceinline__ float4 ld_cg4(const float* p)          // L2-coherent: rows are updated by other SMs' reds
{
    return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float ld_cg1(const float* p) { return __ldcg(p); }

__device__ __forceinline__ void red_add4(float* p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc)
{
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float  get4(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

// all-lanes butterfly sum over the G lanes of each group (every lane of the warp must call)
template <int G>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, G);
    return v;
}

template <int G>
__device__ __forceinline__ unsigned group_ballot(bool pred, int gw)
{
    const unsigned b = __ballot_sync(0xffffffffu, pred);
    if constexpr (G == 32) return b;
    else return (b >> (gw * G)) & ((1u << G) - 1u);
}

// bit index of `item` in the 32*deg-bit membership filter of a user with `deg` observed items (deg < 2^27)
__host__ __device__ __forceinline__ uint32_t bloom_slot(int item, int deg)
{
    uint32_t x = (uint32_t)item * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
#ifdef __CUDA_ARCH__
    return __umulhi(x, (uint32_t)deg << 5);
#else
    return (uint32_t)(((uint64_t)x * ((uint32_t)deg << 5)) >> 32);
#endif
}

// Membership of `cand` in the sorted list items[0..deg): (G+1)-ary search, one probe per lane and round.
// Replaces the reference's linear scan (`lsearch`, rankfm/_rankfm.pyx:20-27; `bsearch` :30-45 is dead code there).
// Warp-uniform control flow: every lane of the warp calls with its group's arguments; `active` masks groups out.
// 32-bit pivot arithmetic: needs G*deg < 2^32, i.e. deg < 1.3e8 (the API rejects larger catalogues).
template <int G>
__device__ __forceinline__ bool group_member(int cand, const int32_t* __restrict__ items, int deg, bool active, int sub, int gw)
{
    int lo = 0, hi = active ? deg : 0;
    bool found = false;
    // NOTE: the ballots sit at ONE program point for all 32 lanes -- groups of the same warp are in different
    // phases of their searches (or idle), and *_sync primitives must never be reached through divergent branches.
    while (__any_sync(0xffffffffu, hi > lo && !found)) {
        const int len = hi - lo;
        const bool live = len > 0 && !found;
        const bool small = len <= G;
        // small: element lo+sub ; large: pivot p_sub = lo + (sub+1)*len/(G+1), strictly increasing because len > G
        const int idx = small ? lo + sub : lo + (int)(((unsigned)(sub + 1) * (unsigned)len) / (unsigned)(G + 1));
        const bool probe = live && (!small || sub < len);
        const int e = probe ? __ldg(items + idx) : 0;
        const unsigned eq = group_ballot<G>(probe && e == cand, gw);
        const unsigned lt = group_ballot<G>(probe && e < cand, gw);
        if (eq) {
            found = true;
        } else if (live) {
            if (small) {
                hi = lo;
            } else {
                const int c = __popc(lt);                       // pivots below cand: p_0..p_{c-1}
                const int nlo = c == 0 ? lo : lo + (int)(((unsigned)c * (unsigned)len) / (unsigned)(G + 1)) + 1;
                const int nhi = c == G ? hi : lo + (int)(((unsigned)(c + 1) * (unsigned)len) / (unsigned)(G + 1));
                lo = nlo; hi = nhi;
            }
        }
    }
    return found;
}

}  // namespace rfm
