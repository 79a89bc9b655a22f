// rfm_gemm.cu -- tensor-core candidate generation for `_recommend` (rankfm/_rankfm.pyx:393-460).
//
// The reference scores every item for every requested user with a scalar loop (:440-441) and fully sorts the scores
// (:444).  Mathematically that is S = A . B^T + bias with
//     A[u]  = [ v_u + x_uf.v_uf  |  v_u ]            (second half only with item features)
//     B[i]  = [ v_i              |  x_if[i].v_if ]
//     bias  = w_i + x_if[i].w_if                     (fp32)
// (derived from compute_ui_utility :48-89; there is no user-feature x item-feature cross term), i.e. a dense GEMM whose
// output (U x I) can never be materialised at cfg5 scale (1M x 1M).  So:
//
//   pack_gemm_items_kernel  B / bias / order in DESCENDING BIAS ORDER (once per weight state): the epilogues below work on
//                           raw dot products and touch the bias once per run of consecutive positions.
//   score_filter_kernel     bf16 operands, fp32 accumulation on the 5th-generation tensor cores:
//       TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory -> tcgen05.mma (cta_group::1, M=128, N=128|256,
//       K=16 per instruction, issued by one elected thread) -> accumulators in TMEM (2 stages) -> tcgen05.ld in the
//       epilogue warps (double-buffered in registers).  One CTA owns 128 or 256 users (A stays resident in shared memory)
//       and streams its share of the item tiles; warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
//       allocator, warp 3 bias producer, warps 4-11 epilogue (thread <-> user row).  The epilogue never writes scores:
//         pass 1 (MODE_ROWMAX)  a lower bound of the best score of every 8-item block of the highest-bias 1/k of the
//                               catalogue; row_threshold_*_kernel takes the n'-th largest bound of a row: a valid lower
//                               bound tau_r of the row's n'-th best score.
//         pass 2 (MODE_FILTER)  (dot, position) is appended to the row's candidate buffer wherever the score can reach
//                               tau_r: a superset of the row's n' best items, ~2-4 n' entries.
//   shortlist_kernel        per user: n' best candidates by bf16 score, their exact fp32 utility (same code as predict),
//                           seen items dropped, bitonic sort, final top-n row.
//
// n' = 2*n_items + 16 (+ the user's number of seen items when filtering) absorbs the bf16 rounding of the candidate
// scores; the final ranking is exact fp32.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include "rfm_host.h"
#include "rfm_kernels.h"
#include "rfm_pair.cuh"

namespace rfm {

// ---------------------------------------------------------------------------------------------------------------
// operand packing: bf16 A (requested users) / B (all items), fp32 bias
// ---------------------------------------------------------------------------------------------------------------
// The item operand is laid out in DESCENDING BIAS ORDER: position pos of B / bias holds item order[pos].  Inside any run of
// consecutive positions the largest bias is the first and the smallest the last, which lets the GEMM epilogues work on
// the raw dot products (one bias per 32- or 64-item run instead of one per score) -- see score_filter_kernel.
__global__ void item_bias_kernel(const Tables T, float* __restrict__ bias, int32_t* __restrict__ iota)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T.I; i += gridDim.x * blockDim.x) {
        const float* row = T.IT + (size_t)i * T.ldi;
        float b = row[T.Fp];
        if (T.x_if_any) for (int q = 0; q < T.Q; ++q) b += row[T.Fp + 4 + q] * T.GP[q];
        bias[i] = b;
        iota[i] = i;
    }
}

// bf16 operand [I_pad, Kp] (Kp = factor columns padded to a multiple of 64) in bias order; pads: zero rows, bias -1e30, order -1
__global__ void pack_gemm_items_kernel(const Tables T, int Kp, int I_pad, const int32_t* __restrict__ order, __nv_bfloat16* __restrict__ B,
                                       float* __restrict__ bias)
{
    const long long n = (long long)I_pad * Kp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int pos = (int)(e / Kp), c = (int)(e % Kp);
        float v = 0.f;
        if (pos < T.I) {
            const float* row = T.IT + (size_t)order[pos] * T.ldi;
            if (c < T.F) v = row[c];
            else if (T.x_if_any && c >= T.Fp && c - T.Fp < T.F) {             // second half: x_if[i] . v_if[:, f]
                const int f = c - T.Fp;
                for (int q = 0; q < T.Q; ++q) v += row[T.Fp + 4 + q] * T.GP[T.gp_vif + (size_t)q * T.Fp + f];
            }
        }
        B[e] = __float2bfloat16(v);
    }
    for (int pos = T.I + blockIdx.x * blockDim.x + threadIdx.x; pos < I_pad; pos += gridDim.x * blockDim.x) bias[pos] = -1e30f;
}

__global__ void pack_gemm_users_kernel(const Tables T, const int32_t* __restrict__ users, int n_users, int M_pad, int Kp, __nv_bfloat16* __restrict__ A)
{
    const long long n = (long long)M_pad * Kp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / Kp), c = (int)(e % Kp);
        float v = 0.f;
        const int u = r < n_users ? users[r] : -1;
        if (u >= 0) {
            const float* row = T.UT + (size_t)u * T.ldu;
            if (c < T.F) {
                v = row[c];
                if (T.x_uf_any) for (int p = 0; p < T.P; ++p) v += row[T.Fp + p] * T.GP[T.gp_vuf + (size_t)p * T.Fp + c];
            } else if (T.x_if_any && c >= T.Fp && c - T.Fp < T.F) v = row[c - T.Fp];
        }
        A[e] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane: asynchronous issue ...
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// ... and the wait that makes the registers valid.  The registers are in/out operands so that no use of them can be
// scheduled above the wait (the load of the NEXT 32 columns is issued before this chunk is processed).
__device__ __forceinline__ void tc_ld32_wait(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// three-input maximum (SASS FMNMX3): 32 scores fold in 16 instructions instead of 31
__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float max8(const uint32_t* r)
{
    float m = max3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
    m = max3(m, __uint_as_float(r[3]), __uint_as_float(r[4]));
    m = max3(m, __uint_as_float(r[5]), __uint_as_float(r[6]));
    return fmaxf(m, __uint_as_float(r[7]));
}

// shared-memory matrix descriptor: K-major tile of 128-byte rows, 128B swizzle, 8-row groups 1024 B apart (sm_100 format)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ uint32_t ord_key(float s) { const uint32_t b = __float_as_uint(s); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }

// Radix-select helpers shared by the threshold / shortlist kernels (256-bin histograms in shared memory).
// hist_add: warp-aggregated increment -- keys of one row share their leading byte (sign + 7 exponent bits), so plain
// shared-memory atomics of the first pass serialise on two or three addresses.  Every lane of the warp must call.
__device__ __forceinline__ void hist_add(uint32_t* hist, bool valid, uint32_t bin)
{
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? bin : 0x100u);
    if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}
// radix_pick (one full warp): the bin, scanning from 255 down, where the running count first reaches `remaining`, and how
// many of that bin's keys are still wanted.  Lane l owns bins [8l, 8l+8); suffix sums over lanes by shuffles.
__device__ __forceinline__ void radix_pick(const uint32_t* hist, uint32_t remaining, uint32_t& bin, uint32_t& left)
{
    const int lane = threadIdx.x & 31;
    uint32_t c[8], mine = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[8 * lane + k]; mine += c[k]; }
    uint32_t suf = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, suf, off);
        if (lane + off < 32) suf += t;
    }
    const uint32_t above = suf - mine;                                   // keys in bins owned by higher lanes
    const bool here = above < remaining && suf >= remaining;
    const unsigned m = __ballot_sync(0xffffffffu, here);
    uint32_t b = 0u, l = 0u;
    if (here) {
        uint32_t rem = remaining - above;
#pragma unroll
        for (int k = 7; k >= 0; --k) {
            if (l == 0u) { if (c[k] >= rem) { b = (uint32_t)(8 * lane + k); l = rem; } else rem -= c[k]; }
        }
    }
    const int src = m ? __ffs(m) - 1 : 0;                                // m == 0: fewer keys than wanted -> bin 0, nothing left
    bin = __shfl_sync(0xffffffffu, b, src);
    left = __shfl_sync(0xffffffffu, l, src);
}

// ---------------------------------------------------------------------------------------------------------------
// the GEMM + running-threshold filter
// ---------------------------------------------------------------------------------------------------------------
struct GemmParams {
    const float* bias;           // [I_pad] fp32 item bias (w_i + x_if.w_if) in descending order (position order of B); -1e30 for pads
    int kblocks;                 // Kp / 64
    int n_tiles;                 // item tiles this launch visits (pass 1 may visit only every tile_stride-th tile)
    int tile_stride;             // visited tile k is item tile k * tile_stride
    int n_splits;                // item-range splits (grid.y)
    int n_users;                 // valid rows of A
    int nstage;
    // MODE_FILTER
    float2* cand;                // [M_pad * n_slots, cap]  (raw dot product, item POSITION as int bits); n_slots = n_splits * gemm_slots_per_split
    int* cand_cnt;               // [M_pad * n_slots]; cap+1 flags an overflowing slot
    const float* tau;            // [M_pad] per-row threshold from pass 1
    int cap;
    // MODE_ROWMAX
    float* rowmax;               // [M_pad, n_tiles * BLOCK_N/kTauBlock] lower bounds of the best score of every visited 8-item block
    // MODE_DUMP
    float* S;                    // [M_pad, I_pad]
    long long ldS;
};

// warps: 0 TMA producer (A, B), 1 MMA issuer, 2 TMEM allocator, 3 bias producer, 4 .. 4+EW-1 epilogue (EW = 8 or 16)
constexpr int MODE_DUMP = 0, MODE_ROWMAX = 1, MODE_FILTER = 2;

// MSUB = 128-row user sub-tiles per CTA.  MSUB == 2 halves the L2 -> shared-memory traffic per MMA: every B tile that
// lands in shared memory feeds two M=128 MMAs (K is only 64..128 here, so with one sub-tile the kernel asks L2 for
// 64 B/clk/SM -- 148 SMs x 64 B = 9.5 KB/clk against the ~6.3 KB/clk the L2 delivers chip-wide).
// What bounds this kernel: the TMEM READ of the accumulators, not the tensor core.  K is only 64..128, so one 128 x 128 fp32
// accumulator tile (64 KB) is produced in 512 clk (4,096 MAC/clk/SM) but takes 1,024 clk to drain at the ~64 B/clk/SM the
// TMEM delivers to registers (B300_MICROARCH.md "LDTM throughput").  ncu: ~1,980 clk per pair of tiles, 30 % issue
// activity, and the time does not move between EW = 8 and EW = 16 epilogue warps (round 2 A/B: 6.227 vs 6.218 ms per
// 65,536 users) -- more warps hide latency, they do not add TMEM bandwidth.  With every score read once as fp32 the
// ceiling is therefore 2K/(4 B / 64 B/clk) = 0.5 of the tensor peak at K = 128; pass 2 runs at ~0.9 of that.
// EW = epilogue warps (8; RANKFM_B200_GEMM_EW=16 with two user sub-tiles).  A TMEM lane quarter may be read by any warp
// with the same (warp % 4); warps sharing a row split its tile columns.
template <int BLOCK_N, int MSUB, int MODE, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p)
{
    static_assert(2 * MSUB * BLOCK_N <= 512, "two accumulator stages must fit the 512 TMEM columns");
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int A_KB_BYTES = 128 * 128;              // one 64-wide k-block of the A tile (128 rows x 128 B)
    constexpr int B_KB_BYTES = BLOCK_N * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_n = p.kblocks, nstage = p.nstage;
    unsigned char* sA = smem;                          // [MSUB][kb_n] k-blocks
    unsigned char* sB = sA + (size_t)MSUB * kb_n * A_KB_BYTES;
    float* sBias = reinterpret_cast<float*>(sB + (size_t)nstage * kb_n * B_KB_BYTES);      // [2][BLOCK_N], one slot per accumulator stage
    unsigned char* sBar = reinterpret_cast<unsigned char*>(sBias + 2 * BLOCK_N);
    const uint32_t bar_full = s32(sBar), bar_empty = bar_full + 8u * nstage, bar_a = bar_empty + 8u * nstage;
    const uint32_t bar_tfull = bar_a + 8u, bar_tempty = bar_tfull + 16u, bar_bias = bar_tempty + 16u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sBar + 8 * (2 * nstage + 7));

    // this CTA: user tile blockIdx.x, visited item tiles blockIdx.y, blockIdx.y + n_splits, ... (round robin: in bias order
    // a contiguous range would hand one split all of a row's best items and overflow its candidate slot)
    const int m0 = blockIdx.x * (128 * MSUB);
    constexpr int ACC_COLS = MSUB * BLOCK_N;           // TMEM columns of one accumulator stage
    const int t0 = blockIdx.y, tstep = p.n_splits;
    const int my_tiles = t0 < p.n_tiles ? (p.n_tiles - t0 + tstep - 1) / tstep : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < nstage; ++s) { bar_init(bar_full + 8u * s, 1); bar_init(bar_empty + 8u * s, 1); }
        bar_init(bar_a, 1);
        for (int a = 0; a < 2; ++a) { bar_init(bar_tfull + 8u * a, 1); bar_init(bar_tempty + 8u * a, EW); bar_init(bar_bias + 8u * a, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {            // TMEM: 2 accumulator stages of BLOCK_N fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(2 * ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            bar_expect_tx(bar_a, (uint32_t)(MSUB * kb_n) * A_KB_BYTES);
            for (int h = 0; h < MSUB; ++h)
                for (int kb = 0; kb < kb_n; ++kb) tma_load_2d(s32(sA + (size_t)(h * kb_n + kb) * A_KB_BYTES), &tmA, kb * 64, m0 + h * 128, bar_a);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % nstage;
                const uint32_t ph = (uint32_t)(it / nstage) & 1u;
                bar_wait(bar_empty + 8u * s, ph ^ 1u);
                bar_expect_tx(bar_full + 8u * s, (uint32_t)kb_n * B_KB_BYTES);
                const int n0 = (t0 + it * tstep) * p.tile_stride * BLOCK_N;
                for (int kb = 0; kb < kb_n; ++kb)
                    tma_load_2d(s32(sB + ((size_t)s * kb_n + kb) * B_KB_BYTES), &tmB, kb * 64, n0, bar_full + 8u * s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread drives the tensor core =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16_f32(BLOCK_N);
            bar_wait(bar_a, 0);
            for (int it = 0; it < my_tiles; ++it) {
                const int s = it % nstage, as = it & 1;
                const uint32_t ph = (uint32_t)(it / nstage) & 1u, aph = (uint32_t)(it >> 1) & 1u;
                bar_wait(bar_tempty + 8u * as, aph ^ 1u);           // epilogue has drained this accumulator stage
                bar_wait(bar_full + 8u * s, ph);                    // B tile has landed
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < MSUB; ++h) {                    // the same B tile against each 128-row user sub-tile
                    const uint32_t d = tmem_base + (uint32_t)(as * ACC_COLS + h * BLOCK_N);
                    for (int kb = 0; kb < kb_n; ++kb) {
                        const uint64_t ad = smem_desc_sw128(s32(sA + (size_t)(h * kb_n + kb) * A_KB_BYTES));
                        const uint64_t bd = smem_desc_sw128(s32(sB + ((size_t)s * kb_n + kb) * B_KB_BYTES));
#pragma unroll
                        for (int k = 0; k < 4; ++k)                 // 4 x (K=16) per 64-wide k-block: +32 B per step
                            tc_mma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                tc_commit(bar_empty + 8u * s);                      // smem stage reusable once these MMAs retire
                tc_commit(bar_tfull + 8u * as);                     // accumulator ready for the epilogue
            }
        }
    } else if (warp == 3) {
        // ===== bias producer: the tile's fp32 biases ride next to the accumulator stage they belong to =====
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int as = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                bar_wait(bar_tempty + 8u * as, aph ^ 1u);           // the epilogue is done with this slot
                bar_expect_tx(bar_bias + 8u * as, BLOCK_N * 4u);
                bulk_load_1d(s32(sBias + as * BLOCK_N), p.bias + (size_t)(t0 + it * tstep) * p.tile_stride * BLOCK_N, BLOCK_N * 4u, bar_bias + 8u * as);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread <-> user row (TMEM lane).  Two warps share each TMEM lane quarter: with one user sub-tile
        // they split the tile's columns, with two sub-tiles each takes all columns of one sub-tile.
        // The epilogue is the issue-bound part of this kernel (K is only 64..128: 2.3 instructions per score would cost more
        // issue slots than the MMAs take cycles), so it works on the RAW dot products and touches the bias once per run of
        // items -- possible because positions are in descending bias order (run maximum = first, minimum = last):
        //   pass 1  lower bound of an 8-item block's best score = max(dot) + smallest bias of the block  (FMNMX3 tree).
        //           Small blocks because bias order CLUSTERS a row's best items (one bound per block must not hide them)
        //   pass 2  score >= tau can only hold where dot >= tau - largest bias of the 32-item chunk; the four 8-item
        //           sub-maxima of the tree gate the (rare per row) append of (dot, position)
        // Both are conservative (candidates form a superset); exact fp32 scores are recomputed from the shortlist. =====
        const int wq = warp & 3;                                    // TMEM lane quarter this warp may access
        const int part = (warp - 4) >> 2;                           // 0 .. EW/4 - 1: (user sub-tile, column split)
        constexpr int CSPLIT = (EW / 4) / MSUB;                     // warps sharing a row split its tile columns
        static_assert(CSPLIT >= 1 && CSPLIT * MSUB * 4 == EW, "epilogue warps = 4 lane quarters x sub-tiles x column splits");
        constexpr int COLS = BLOCK_N / CSPLIT;                      // tile columns this warp reads
        constexpr int NCH = COLS / 32;
        const int h = part % MSUB, cs = part / MSUB;
        const int col0 = cs * COLS;                                 // ... starting at this tile column
        const int tcol0 = h * BLOCK_N + col0;                       // ... found at this column of the accumulator stage
        const int row = m0 + h * 128 + wq * 32 + lane;
        const bool row_ok = row < p.n_users;
        const float tau = MODE == MODE_FILTER ? (row_ok ? p.tau[row] : INFINITY) : 0.f;
        const int cap = p.cap;
        const int n_slots = CSPLIT * p.n_splits;
        const int slot = blockIdx.y * CSPLIT + cs;
        float2* const my = MODE == MODE_FILTER ? p.cand + ((size_t)row * n_slots + slot) * cap : nullptr;
        float2* wp = my;
        float2* const wp_room = my + (cap - 8);                     // last write position that still leaves room for 8 entries
        bool over = false;
        float* rmax = MODE == MODE_ROWMAX ? p.rowmax + (size_t)row * (p.n_tiles * (BLOCK_N / kTauBlock)) : nullptr;

        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            bar_wait(bar_bias + 8u * as, aph);
            bar_wait(bar_tfull + 8u * as, aph);
            tc_fence_after();
            const int n0 = (t0 + it * tstep) * p.tile_stride * BLOCK_N + col0;
            const float* sb = sBias + as * BLOCK_N + col0;
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * ACC_COLS + tcol0);
            uint32_t ra[32], rb[32];                                // double-buffered: chunk c+1 is in flight while chunk c is processed
            float4 lb_even = zero4();
            static_assert(NCH % 2 == 0, "block bounds are stored per pair of 32-column chunks");
            tc_ld32_issue(taddr, ra);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t (&v)[32] = (c & 1) ? rb : ra;
                tc_ld32_wait(v);
                if (c + 1 < NCH) tc_ld32_issue(taddr + (uint32_t)((c + 1) * 32), (c & 1) ? ra : rb);
                if (MODE == MODE_DUMP) {
                    float* out = p.S + (size_t)row * p.ldS + n0 + c * 32;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = reinterpret_cast<const float4*>(sb + c * 32)[q];
                        reinterpret_cast<float4*>(out)[q] = make_float4(__uint_as_float(v[4 * q]) + b.x, __uint_as_float(v[4 * q + 1]) + b.y,
                                                                        __uint_as_float(v[4 * q + 2]) + b.z, __uint_as_float(v[4 * q + 3]) + b.w);
                    }
                } else {
                    const float m0_ = max8(v), m1_ = max8(v + 8), m2_ = max8(v + 16), m3_ = max8(v + 24);
                    if (MODE == MODE_ROWMAX) {
                        // one bound per 8-item block: its best dot product + its smallest (= last) bias
                        const float4 lb = make_float4(m0_ + sb[c * 32 + 7], m1_ + sb[c * 32 + 15], m2_ + sb[c * 32 + 23], m3_ + sb[c * 32 + 31]);
                        if (c & 1) {                                 // two chunks' bounds leave together: one full 32-byte sector per row
                            float4* dst = reinterpret_cast<float4*>(rmax + (size_t)(t0 + it * tstep) * (BLOCK_N / kTauBlock) + (col0 + (c - 1) * 32) / kTauBlock);
                            dst[0] = lb_even; dst[1] = lb;
                        } else lb_even = lb;
                    } else {
                        // fl(fl(m + b) - b) can exceed m by an ulp: the slack keeps every item that set a pass-1 bound
                        const float bmax = sb[c * 32];
                        const float thr = (tau - bmax) - 4.8e-7f * (fabsf(tau) + fabsf(bmax));
                        const float msub[4] = {m0_, m1_, m2_, m3_};
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (msub[g] >= thr) {                  // rare per row (a warp takes it when any of its 32 rows does)
                                if (wp > wp_room) over = true;
                                else {
#pragma unroll
                                    for (int k = 0; k < 8; ++k) {
                                        const float d = __uint_as_float(v[8 * g + k]);
                                        if (d >= thr) *wp++ = make_float2(d, __int_as_float(n0 + c * 32 + 8 * g + k));
                                    }
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(bar_tempty + 8u * as);
        }
        if (MODE == MODE_FILTER) p.cand_cnt[(size_t)row * n_slots + slot] = !row_ok ? 0 : (over ? cap + 1 : (int)(wp - my));
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// shortlist_kernel: one block per requested user, from the pass-2 candidates to the final recommendation row.
//   1. gather the row's candidates (raw dot, position) from its slots into shared memory, bf16-GEMM score =
//      dot + bias[position]
//   2. keep the n' best by that score (exact radix select of the n'-th largest + ordered compaction): pass 2
//      over-collects when the pass-1 threshold came from a subset of the item tiles, and only the n' best need more work
//   3. exact fp32 utility of the kept items (same lane-group code as predict); seen items are dropped (:450)
//   4. bitonic sort of the (at most kShortWidth) exact scores, best n_items out -- `np.argsort(...)[::-1]` + the walk
//      of `_rankfm.pyx:444-456`, on the shortlist
// flag[row] = 1 when the row must be redone on the exact path (a candidate slot overflowed, or more than kShortWidth
// candidates tie at the cut), 2 when fewer than n' candidates reached the row threshold (an estimated threshold that came
// out too high: redone with the provable one); unknown users (-1) get the reference's NaN row (:437-438).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kShortThreads = 256;

// Blocks per SM: the kernel is latency-bound (block-wide barriers between the select passes, the compaction and the sort
// stages), so residency is what it needs.  Unconstrained it took 64 registers = 4 blocks per SM; the common variant (no side
// features, one quad per lane, narrow tier) fits 40 registers (8 B of spills) = 6 blocks per SM: -8 % on a whole recommend
// call (round 2 A/B on one box: 8.10 -> 7.42 ms per 65,536 users together with the 2,048-entry staging below).
#ifndef RFM_SHORT_MINB
#define RFM_SHORT_MINB 6
#endif
template <int G, int QPL, bool FEAT, int SW>
__global__ void __launch_bounds__(kShortThreads, (!FEAT && QPL == 1 && SW <= 512) ? RFM_SHORT_MINB : 4) shortlist_kernel(const Tables T, const int32_t* __restrict__ users, const float2* __restrict__ cand,
                                                                  const int* __restrict__ cand_cnt, int slots, int cap, const float* __restrict__ bias,
                                                                  const int32_t* __restrict__ order, const int* __restrict__ n_target,
                                                                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                                  int filter_previous, int n_items, float* __restrict__ rec, int* __restrict__ flag,
                                                                  const float* __restrict__ tau, int I_pad, int guard, int stage_cap)
{
    extern __shared__ __align__(16) unsigned char short_smem[];
    __shared__ float s_norm2;
    uint2* ent = reinterpret_cast<uint2*>(short_smem);               // [min(slots * cap, stage_cap)] (ordered key of the bf16 score, position)
    __shared__ int32_t kept[SW];                            // item ids of the shortlist, in position order
    __shared__ unsigned long long sel[SW];                  // (ordered key of the exact score << 32) | item index
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining;
    __shared__ int s_off[65], s_over, s_base, wsum[kShortThreads / 32];
    constexpr int GPW = 32 / G;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane % G, gw = lane / G;
    const int u = __ldg(users + b);
    const bool known = u >= 0;
    float* out = rec + (size_t)b * n_items;
    const int* cc = cand_cnt + (size_t)b * slots;
    if (tid == 0) {
        int total = 0, over = 0;
        for (int sl = 0; sl < slots; ++sl) { const int k = cc[sl]; s_off[sl] = total; over |= k > cap; total += min(k, cap); }
        s_off[slots] = total; s_over = over; s_prefix = 0u; s_base = 0;
        s_remaining = (uint32_t)n_target[b];
    }
    __syncthreads();
    const int total = s_off[slots], want = n_target[b];
    if (!known || s_over || total > stage_cap) {                      // overflow (of a slot, or of this kernel's staging): the caller redoes the row
        if (!known) for (int k = tid; k < n_items; k += kShortThreads) out[k] = __int_as_float(0x7fc00000);
        if (tid == 0) flag[b] = known ? 1 : 0;
        return;
    }
    // 1. gather
    for (int sl = 0; sl < slots; ++sl) {
        const int n = s_off[sl + 1] - s_off[sl];
        const float2* src = cand + ((size_t)b * slots + sl) * cap;
        for (int k = tid; k < n; k += kShortThreads) {
            const float2 ce = src[k];
            const int pos = __float_as_int(ce.y);
            const bool ok = pos >= 0 && pos < T.I;                    // positions of padded items carry no item
            ent[s_off[sl] + k] = make_uint2(ok ? max(ord_key(ce.x + __ldg(bias + (ok ? pos : 0))), 1u) : 0u, (uint32_t)pos);
        }
    }
    __syncthreads();
    // 2. the n'-th largest key (1 = keep every real entry)
    uint32_t cut = 1u;
    if (want < total) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            hist[tid] = 0u;
            __syncthreads();
            const uint32_t prefix = s_prefix, pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
            for (int e0 = 0; e0 < total; e0 += kShortThreads) {      // warp-uniform trip count (hist_add is warp-collective)
                const int e = e0 + tid;
                const uint32_t key = e < total ? ent[e].x : 0u;
                const bool match = e < total && (key & pmask) == prefix;
                if (pass == 0) hist_add(hist, match, (key >> shift) & 0xffu);
                else if (match) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                uint32_t bin, left;
                radix_pick(hist, s_remaining, bin, left);
                if (tid == 0) { s_prefix = prefix | (bin << shift); s_remaining = left; }
            }
            __syncthreads();
        }
        cut = max(s_prefix, 1u);
    }
    // 2b. was the row threshold low enough?  Pass 2 collected EVERY item whose bf16 score reaches tau, so the n' best
    //     candidates are the n' best items of the catalogue iff n' candidates reach tau, i.e. iff the cut is >= tau.  With the
    //     provable threshold (tau_rank: z = 0) that holds by construction; with the estimated one it is what makes the
    //     estimate safe to use: a row whose estimate came out too high is flagged 2 and redone with the provable threshold.
    {
        const float tau_b = tau[b];
        const bool short_of = tau_b != -INFINITY && (want < total ? ord_key(tau_b) > cut : want > total);
        if (short_of) {
            if (tid == 0) flag[b] = 2;
            return;
        }
    }
    //    ordered compaction (position order keeps the result independent of thread scheduling)
    for (int base = 0; base < total; base += kShortThreads) {
        const int e = base + tid;
        const bool keep = e < total && ent[e].x >= cut;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = s_base, all = 0;
        for (int w = 0; w < kShortThreads / 32; ++w) { if (w < warp) before += wsum[w]; all += wsum[w]; }
        const int mine = before + __popc(bal & ((1u << lane) - 1u));
        if (keep && mine < SW) kept[mine] = __ldg(order + ent[e].y);
        __syncthreads();
        if (tid == 0) s_base += all;
        __syncthreads();
    }
    const int n_kept = s_base;
    if (n_kept > SW) {                                        // > SW - n' ties at the cut: exact path
        if (tid == 0) flag[b] = 1;
        return;
    }
    if (tid == 0) flag[b] = 0;
    const int n_sort = n_kept <= SW / 2 ? SW / 2 : SW;             // power of two >= n_kept
    for (int e = tid; e < n_sort; e += kShortThreads) sel[e] = 0ull;                            // key 0 sorts last
    __syncthreads();
    // 3. exact fp32 re-score
    UserCtx<QPL> uc;
    load_user<G, QPL, FEAT>(T, u, true, sub, uc);
    user_precompute<G, QPL, FEAT>(T, T.GP, true, sub, uc);
    {   // |A_u|^2 of the GEMM's user operand [ a | v_u ] (second half only with item features)
        float n2 = 0.f;
#pragma unroll
        for (int k = 0; k < QPL; ++k) { n2 = dot4(uc.a[k], uc.a[k], n2); if (FEAT && T.x_if_any) n2 = dot4(uc.vu[k], uc.vu[k], n2); }
        n2 = group_sum<G>(n2);
        if (tid == 0) s_norm2 = n2;
    }
    long long seg = 0; int deg = 0;
    if (filter_previous) { seg = __ldg(indptr + u); deg = (int)(__ldg(indptr + u + 1) - seg); }
    constexpr int STRIDE = (kShortThreads / 32) * GPW;
    constexpr int INFL = QPL <= 2 ? 4 : 2;                            // item rows in flight per lane group
    for (int e0 = 0; e0 < n_kept; e0 += INFL * STRIDE) {              // warp-uniform trip count
        int ee[INFL], ii[INFL];
        ItemRow<QPL> rr[INFL];
#pragma unroll
        for (int w = 0; w < INFL; ++w) {
            ee[w] = e0 + w * STRIDE + warp * GPW + gw;
            ii[w] = ee[w] < n_kept ? kept[ee[w]] : 0;
            load_item<G, QPL, FEAT>(T, ii[w], ee[w] < n_kept, sub, rr[w]);
        }
#pragma unroll
        for (int w = 0; w < INFL; ++w) {
            const bool ok = ee[w] < n_kept;
            const float sc = utility<G, QPL, FEAT>(uc, rr[w]);
            const bool seen = filter_previous ? group_member<G>(ii[w], indices + seg, deg, ok, sub, gw) : false;
            if (sub == 0 && ok && !seen) sel[ee[w]] = ((unsigned long long)max(ord_key(sc), 1u) << 32) | (uint32_t)ii[w];
        }
    }
    __syncthreads();
    // 4. bitonic sort, descending by (exact score, item index): equal scores come out larger index first, like the reversed
    //    argsort of the reference and the exact path -- and independent of how pass 2 laid the candidates out
    for (int k = 2; k <= n_sort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n_sort; t += kShortThreads) {
                const int x = t ^ j;
                if (x > t) {
                    const unsigned long long a = sel[t], c = sel[x];
                    const bool desc = (t & k) == 0;
                    if (desc ? (a < c) : (a > c)) { sel[t] = c; sel[x] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int k = tid; k < n_items; k += kShortThreads) {
        const unsigned long long e = k < n_sort ? sel[k] : 0ull;
        out[k] = (e >> 32) == 0ull ? __int_as_float(0x7fc00000) : (float)(uint32_t)(e & 0xffffffffull);
    }
    // 5. is the bf16 shortlist PROVABLY a superset of the exact top n_items?  Every item outside it has a bf16 score <= c
    //    (the cut of step 2, or the row threshold tau when every candidate was kept), hence an exact score <= c + delta with
    //    delta = 2^-8 |A_u| max_i |B_i| (both operands rounded to bf16: 2^-9 relative each, Cauchy-Schwarz over the sum).
    //    If the n_items-th exact score found clears that, nothing outside can displace it; otherwise the row is redone on
    //    the exact fp32 path (flag), like an overflowing candidate slot.
    if (guard && tid == 0 && n_items <= n_sort) {
        const uint32_t kn = (uint32_t)(sel[n_items - 1] >> 32);
        if (kn != 0u) {
            auto key_to_float = [](uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); };
            const float e_n = key_to_float(kn);
            const float c = want < total ? key_to_float(cut) : tau[b];
            const float delta = 1.05f * 0.00390625f * sqrtf(s_norm2) * __ldg(bias + I_pad) + 1e-6f * fabsf(c);
            if (!(e_n >= c + delta)) flag[b] = 1;
        }
    }
}

static bool shortlist_guard()
{
    const char* e = getenv("RANKFM_B200_TC_GUARD");            // experiments: 0 = trust the 2n+16 shortlist without the bf16 error bound
    return !(e && !strcmp(e, "0"));
}

template <int G, int QPL, int SW>
static cudaError_t shortlist_launch(const Tables& T, const int32_t* users, int n_users, const float2* cand, const int* cnt, int slots, int cap, const float* bias,
                                    const int32_t* order, const int* n_target, const int64_t* indptr, const int32_t* indices, int filt, int n_items, float* rec,
                                    int* flag, const float* tau, int I_pad, int guard, int stage_cap, size_t smem, cudaStream_t st)
{
    cudaError_t e;
    if (T.x_uf_any || T.x_if_any) {
        e = cudaFuncSetAttribute(shortlist_kernel<G, QPL, true, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        shortlist_kernel<G, QPL, true, SW><<<n_users, kShortThreads, smem, st>>>(T, users, cand, cnt, slots, cap, bias, order, n_target, indptr, indices, filt, n_items, rec, flag, tau, I_pad, guard, stage_cap);
    } else {
        e = cudaFuncSetAttribute(shortlist_kernel<G, QPL, false, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        shortlist_kernel<G, QPL, false, SW><<<n_users, kShortThreads, smem, st>>>(T, users, cand, cnt, slots, cap, bias, order, n_target, indptr, indices, filt, n_items, rec, flag, tau, I_pad, guard, stage_cap);
    }
    return cudaGetLastError();
}

template <int G, int QPL>
static cudaError_t shortlist_gq(const Tables& T, const int32_t* users, int n_users, const float2* cand, const int* cnt, int slots, int cap, const float* bias,
                                const int32_t* order, const int* n_target, const int64_t* indptr, const int32_t* indices, int filt, int n_items, float* rec,
                                int* flag, const float* tau, int I_pad, int short_width, int stage_cap, cudaStream_t st)
{
    stage_cap = stage_cap > 0 ? min(stage_cap, slots * cap) : slots * cap;
    const size_t smem = (size_t)stage_cap * sizeof(uint2);
    const int guard = shortlist_guard() ? 1 : 0;
    if (short_width > kShortWidth) return shortlist_launch<G, QPL, kShortWidthWide>(T, users, n_users, cand, cnt, slots, cap, bias, order, n_target, indptr, indices, filt, n_items, rec, flag, tau, I_pad, guard, stage_cap, smem, st);
    return shortlist_launch<G, QPL, kShortWidth>(T, users, n_users, cand, cnt, slots, cap, bias, order, n_target, indptr, indices, filt, n_items, rec, flag, tau, I_pad, guard, stage_cap, smem, st);
}

// rec [n_users, n_items]: final rows (float item indexes, NaN-padded like topn_select_kernel); flag [n_users]
cudaError_t launch_shortlist(const Tables& T, const int32_t* users, int n_users, const float2* cand, const int* cnt, int slots, int cap, const float* bias,
                             const int32_t* order, const int* n_target, const int64_t* indptr, const int32_t* indices, int filt, int n_items, float* rec,
                             int* flag, const float* tau, int I_pad, int short_width, int stage_cap, cudaStream_t st)
{
    int qpl = 1;
    const int G = train_group_size(T, &qpl);
    if (max(T.Pp, T.Qp) > 4 * G || qpl > 4 || slots > 64 || (size_t)slots * cap * sizeof(uint2) > 160 * 1024) return cudaErrorInvalidValue;
#define RFM_SHORT(GG, QQ) return shortlist_gq<GG, QQ>(T, users, n_users, cand, cnt, slots, cap, bias, order, n_target, indptr, indices, filt, n_items, rec, flag, tau, I_pad, short_width, stage_cap, st)
    switch (G) {
        case 4:  RFM_SHORT(4, 1);
        case 8:  RFM_SHORT(8, 1);
        case 16: RFM_SHORT(16, 1);
        default:
            if (qpl == 1) RFM_SHORT(32, 1);
            if (qpl == 2) RFM_SHORT(32, 2);
            RFM_SHORT(32, 4);
    }
#undef RFM_SHORT
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// row-major bf16 [rows, Kp] -> boxes of 64 (K) x box_rows, 128B swizzle
static bool make_map(CUtensorMap* map, const void* base, long long rows, int Kp, int box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)Kp * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool gemm_encode_available() { return encode_tiled_fn() != nullptr; }
int gemm_kraw(const Tables& T) { return T.x_if_any ? 2 * T.Fp : T.Fp; }
int gemm_kp(const Tables& T) { return (gemm_kraw(T) + 63) / 64 * 64; }
// user sub-tiles per CTA: 2 whenever A (2 x 128 rows) and >= 2 B stages fit shared memory (RANKFM_B200_GEMM_MSUB=1 forces 1)
int gemm_msub(const Tables& T)
{
    const char* e = getenv("RANKFM_B200_GEMM_MSUB");
    if (e && atoi(e) == 1) return 1;
    return gemm_kp(T) <= 128 ? 2 : 1;
}
int gemm_block_n(const Tables& T) { return gemm_msub(T) == 2 ? 128 : (gemm_kp(T) <= 128 ? 256 : 128); }
int gemm_m_tile(const Tables& T) { return 128 * gemm_msub(T); }
// epilogue warps: 16 with two user sub-tiles (RANKFM_B200_GEMM_EW=8 forces the round-1 layout), else 8
int gemm_epi_warps(const Tables& T)
{
    const char* e = getenv("RANKFM_B200_GEMM_EW");
    return (e && atoi(e) == 16 && gemm_msub(T) == 2) ? 16 : 8;
}
int gemm_slots_per_split(const Tables& T) { return (gemm_epi_warps(T) / 4) / gemm_msub(T); }
bool gemm_supported(const Tables& T) { return gemm_kp(T) <= 256; }

// largest Euclidean norm of a bf16 item operand row: |score_bf16 - score_fp32| <= 2^-8 |A_u| |B_i| bounds what the shortlist
// can miss (see shortlist_kernel, step 5)
__global__ void item_norm_max_kernel(const __nv_bfloat16* __restrict__ B, int I_pad, int Kp, float* __restrict__ out)
{
    float best = 0.f;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < I_pad; pos += gridDim.x * blockDim.x) {
        const uint4* row = reinterpret_cast<const uint4*>(B + (size_t)pos * Kp);
        float n2 = 0.f;
        for (int c = 0; c < Kp / 8; ++c) {
            const uint4 v = row[c];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xffff0000u);
                n2 = fmaf(lo, lo, fmaf(hi, hi, n2));
            }
        }
        best = fmaxf(best, n2);
    }
    for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(sqrtf(best)));      // non-negative floats order like their bits
}

// bias[I_pad] (sorted, descending; bias[I_pad] = largest operand-row norm), order[I_pad] (position -> item), B[I_pad, Kp]; once per weight state
cudaError_t launch_pack_gemm_items(const Tables& T, int Kp, int I_pad, void* B, float* bias, int32_t* order, cudaStream_t st)
{
    // scratch from the library's device block cache (rfm_host.h): a one-shot `_recommend` creates a session per call and
    // would otherwise pay three cudaMalloc/cudaFree pairs (each a device synchronisation) per weight state
    float* raw = nullptr; int32_t* iota = nullptr; void* tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = rfmh::dev_malloc(reinterpret_cast<void**>(&raw), (size_t)T.I * 4);
    if (e == cudaSuccess) e = rfmh::dev_malloc(reinterpret_cast<void**>(&iota), (size_t)T.I * 4);
    if (e == cudaSuccess) {
        item_bias_kernel<<<148 * 4, 256, 0, st>>>(T, raw, iota);
        e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, raw, bias, iota, order, T.I, 0, 32, st);
    }
    if (e == cudaSuccess) e = rfmh::dev_malloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, raw, bias, iota, order, T.I, 0, 32, st);   // stable: ties stay in item order
    if (e == cudaSuccess) e = cudaMemsetAsync(order + T.I, 0xff, (size_t)(I_pad - T.I) * 4, st);
    if (e == cudaSuccess) {
        pack_gemm_items_kernel<<<148 * 8, 256, 0, st>>>(T, Kp, I_pad, order, reinterpret_cast<__nv_bfloat16*>(B), bias);
        cudaMemsetAsync(bias + I_pad, 0, 4 * sizeof(float), st);
        item_norm_max_kernel<<<148 * 4, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(B), I_pad, Kp, bias + I_pad);
        e = cudaGetLastError();
    }
    cudaStreamSynchronize(st);                  // the scratch goes back to the cache: nothing queued may still use it
    rfmh::dev_free(raw); rfmh::dev_free(iota); rfmh::dev_free(tmp);
    return e;
}
cudaError_t launch_pack_gemm_users(const Tables& T, const int32_t* users, int n_users, int M_pad, int Kp, void* A, cudaStream_t st)
{
    pack_gemm_users_kernel<<<148 * 4, 256, 0, st>>>(T, users, n_users, M_pad, Kp, reinterpret_cast<__nv_bfloat16*>(A));
    return cudaGetLastError();
}

// n_target[row]-th largest of the row's block bounds -> tau[row]  (one block per row, exact radix select over ordered keys;
// the row is staged in shared memory once when it fits, so HBM sees it once instead of four times)
constexpr int kThrThreads = 1024;
__global__ void __launch_bounds__(kThrThreads) row_threshold_kernel(const float* __restrict__ rowmax, int n_blocks, const int* __restrict__ n_target,
                                                                    float* __restrict__ tau, int staged, int sample_k, float z)
{
    extern __shared__ __align__(16) unsigned char thr_smem[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(thr_smem);
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining;
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* v = rowmax + (size_t)row * n_blocks;
    const int want = tau_rank(n_target[row], sample_k, z);
    if (want > n_blocks) { if (tid == 0) tau[row] = -INFINITY; return; }
    if (tid == 0) { s_prefix = 0u; s_remaining = (uint32_t)want; }
    if (staged) {
        const float4* v4 = reinterpret_cast<const float4*>(v);               // n_blocks is a multiple of 4 (16 bounds per 128-item tile)
        for (int k = tid; k < n_blocks / 4; k += kThrThreads) {
            const float4 x = __ldg(v4 + k);
            reinterpret_cast<uint4*>(keys)[k] = make_uint4(ord_key(x.x), ord_key(x.y), ord_key(x.z), ord_key(x.w));
        }
    }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix, pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int k0 = 0; k0 < n_blocks; k0 += kThrThreads) {               // warp-uniform trip count (hist_add is warp-collective)
            const int k = k0 + tid;
            const bool inb = k < n_blocks;
            const uint32_t key = !inb ? 0u : (staged ? keys[k] : ord_key(v[k]));
            const bool match = inb && (key & pmask) == prefix;
            if (pass == 0) hist_add(hist, match, (key >> shift) & 0xffu);
            else if (match) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t bin, left;
            radix_pick(hist, s_remaining, bin, left);
            if (tid == 0) { s_prefix = prefix | (bin << shift); s_remaining = left; }
        }
        __syncthreads();
    }
    if (tid == 0) { const uint32_t kb = s_prefix; tau[row] = __uint_as_float((kb & 0x80000000u) ? (kb & 0x7fffffffu) : ~kb); }
}

// Register-resident variant (the common case): every thread keeps its KPT keys in registers for all passes, so a pass
// costs ~5 instructions per key.  The kernel is latency-bound (barriers and the serial bin pick between passes), hence
// small blocks -- several rows in flight per SM -- and three 8-bit passes instead of four: tau is the lower edge of the
// 24-bit bucket of the n'-th largest bound, i.e. smaller by < 2^-15 relative, which only makes it more conservative.
// Pass 0 folds a thread's run of equal leading bytes locally and counts into per-warp private histograms: the leading
// byte (sign + 7 exponent bits) is shared by almost all keys of a row and would serialise block-wide atomics.
template <int THREADS, int KPT>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) row_threshold_reg_kernel(const float* __restrict__ rowmax, int n_blocks, const int* __restrict__ n_target,
                                                                    float* __restrict__ tau, int sample_k, float z)
{
    __shared__ uint32_t hist[256];
    __shared__ uint32_t whist[(THREADS / 32) * 256];
    __shared__ uint32_t s_prefix, s_remaining;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int want = tau_rank(n_target[row], sample_k, z);
    if (want > n_blocks) { if (tid == 0) tau[row] = -INFINITY; return; }
    const float4* v4 = reinterpret_cast<const float4*>(rowmax + (size_t)row * n_blocks);     // n_blocks % 4 == 0
    uint32_t key[KPT];
#pragma unroll
    for (int j = 0; j < KPT / 4; ++j) {
        const int q = j * THREADS + tid;
        const bool inb = q < n_blocks / 4;
        const float4 x = inb ? __ldg(v4 + q) : zero4();
        key[4 * j] = inb ? ord_key(x.x) : 0u; key[4 * j + 1] = inb ? ord_key(x.y) : 0u;      // 0 sorts below every real key
        key[4 * j + 2] = inb ? ord_key(x.z) : 0u; key[4 * j + 3] = inb ? ord_key(x.w) : 0u;
    }
    if (tid == 0) { s_prefix = 0u; s_remaining = (uint32_t)want; }
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int k = tid; k < 256; k += THREADS) hist[k] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix, pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        if (pass == 0) {
            uint32_t* mine = whist + (tid >> 5) * 256;
            for (int k = lane; k < 256; k += 32) mine[k] = 0u;
            __syncwarp();
            uint32_t cur = key[0] >> 24, n = 1u;
#pragma unroll
            for (int j = 1; j < KPT; ++j) {
                const uint32_t bin = key[j] >> 24;
                if (bin == cur) ++n; else { atomicAdd(&mine[cur], n); cur = bin; n = 1u; }
            }
            atomicAdd(&mine[cur], n);
            __syncthreads();
            for (int k = tid; k < 256; k += THREADS) {
                uint32_t sum = 0u;
#pragma unroll 8
                for (int w = 0; w < THREADS / 32; ++w) sum += whist[w * 256 + k];
                hist[k] = sum;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KPT; ++j)
                if ((key[j] & pmask) == prefix) atomicAdd(&hist[(key[j] >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t bin, left;
            radix_pick(hist, s_remaining, bin, left);
            if (tid == 0) { s_prefix = prefix | (bin << shift); s_remaining = left; }
        }
        __syncthreads();
    }
    // low 8 key bits left at zero: for either sign that decodes to a value <= every float of the bucket
    if (tid == 0) { const uint32_t kb = s_prefix; tau[row] = __uint_as_float((kb & 0x80000000u) ? (kb & 0x7fffffffu) : ~kb); }
}

cudaError_t launch_row_threshold(const float* rowmax, int n_rows, int n_blocks, const int* n_target, float* tau, int sample_k, float z, cudaStream_t st)
{
    if (n_blocks % 4 == 0 && n_blocks <= 32 * 1024) {
        // keys per thread sized to the row: with a 1/32 head subset a row has ~1,000 bounds, and a thread that loops over
        // 32 mostly absent keys made this kernel 13 % of the GEMM + filter time (ncu, round 2: 0.199 ms per 18,944 rows)
        if (n_blocks <= 8 * 128) row_threshold_reg_kernel<128, 8><<<n_rows, 128, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        else if (n_blocks <= 8 * 256) row_threshold_reg_kernel<256, 8><<<n_rows, 256, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        else if (n_blocks <= 16 * 256) row_threshold_reg_kernel<256, 16><<<n_rows, 256, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        else if (n_blocks <= 32 * 256) row_threshold_reg_kernel<256, 32><<<n_rows, 256, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        else if (n_blocks <= 32 * 512) row_threshold_reg_kernel<512, 32><<<n_rows, 512, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        else row_threshold_reg_kernel<1024, 32><<<n_rows, 1024, 0, st>>>(rowmax, n_blocks, n_target, tau, sample_k, z);
        return cudaGetLastError();
    }
    const size_t smem = (size_t)n_blocks * 4;
    const int staged = smem <= 200 * 1024 && n_blocks % 4 == 0;
    if (staged) {
        cudaError_t e = cudaFuncSetAttribute(row_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    row_threshold_kernel<<<n_rows, kThrThreads, staged ? smem : 0, st>>>(rowmax, n_blocks, n_target, tau, staged, sample_k, z);
    return cudaGetLastError();
}

template <int BN, int MSUB, int MODE, int EW>
static cudaError_t launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(score_filter_kernel<BN, MSUB, MODE, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    score_filter_kernel<BN, MSUB, MODE, EW><<<grid, 128 + 32 * EW, smem, st>>>(tmA, tmB, p);
    return cudaGetLastError();
}

template <int BN, int MSUB, int EW>
static cudaError_t launch_modes(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid, size_t smem, cudaStream_t st)
{
    if (mode == MODE_DUMP) return launch_mode<BN, MSUB, MODE_DUMP, EW>(tmA, tmB, p, grid, smem, st);
    if (mode == MODE_ROWMAX) return launch_mode<BN, MSUB, MODE_ROWMAX, EW>(tmA, tmB, p, grid, smem, st);
    return launch_mode<BN, MSUB, MODE_FILTER, EW>(tmA, tmB, p, grid, smem, st);
}

// mode 0: dump dense scores into S [M_pad, I_pad]; 1: lower bounds of the best score of every 64-item block of a subset of
// the item tiles (see tile_stride below) into rowmax [M_pad, ceil(n_tiles/|tile_stride|) * BN/64]; 2: candidates that can
// reach tau[row] into cand/cand_cnt.
// M_pad must be a multiple of gemm_m_tile(T), I_pad of gemm_block_n(T).
cudaError_t launch_score_filter(const Tables& T, int mode, const void* A, const void* B, const float* bias, int n_users, int M_pad, int I_pad, int n_splits,
                                int tile_stride, float2* cand, int* cand_cnt, const float* tau, int cap, float* rowmax, float* S, cudaStream_t st)
{
    const int Kp = gemm_kp(T), BN = gemm_block_n(T), MSUB = gemm_msub(T);
    alignas(64) CUtensorMap tmA, tmB;
    if (!make_map(&tmA, A, M_pad, Kp, 128) || !make_map(&tmB, B, I_pad, Kp, BN)) return cudaErrorNotSupported;
    if (tile_stride == 0 || M_pad % (128 * MSUB) || I_pad % BN) return cudaErrorInvalidValue;
    GemmParams p{};
    // tile_stride = k > 0: every k-th item tile; k < 0: the first 1/|k| of the tiles (the highest-bias items of the catalogue)
    const int frac = tile_stride < 0 ? -tile_stride : tile_stride;
    p.bias = bias; p.kblocks = Kp / 64; p.n_tiles = (I_pad / BN + frac - 1) / frac; p.tile_stride = tile_stride < 0 ? 1 : tile_stride;
    p.n_splits = n_splits; p.n_users = n_users;
    p.cand = cand; p.cand_cnt = cand_cnt; p.tau = tau; p.cap = cap; p.rowmax = rowmax; p.S = S; p.ldS = I_pad;
    const size_t a_bytes = (size_t)MSUB * p.kblocks * 128 * 128, stage_bytes = (size_t)p.kblocks * BN * 128;
    int nstage = (int)((196 * 1024 - a_bytes) / stage_bytes);
    nstage = nstage > 4 ? 4 : (nstage < 2 ? 2 : nstage);
    p.nstage = nstage;
    const size_t smem = a_bytes + nstage * stage_bytes + 2 * BN * 4 + 8 * (2 * nstage + 7) + 16 + 1024;
    const dim3 grid(M_pad / (128 * MSUB), n_splits);
    if (MSUB == 2) return gemm_epi_warps(T) == 16 ? launch_modes<128, 2, 16>(mode, tmA, tmB, p, grid, smem, st) : launch_modes<128, 2, 8>(mode, tmA, tmB, p, grid, smem, st);
    if (BN == 256) return launch_modes<256, 1, 8>(mode, tmA, tmB, p, grid, smem, st);
    return launch_modes<128, 1, 8>(mode, tmA, tmB, p, grid, smem, st);
}

}  // namespace rfm
