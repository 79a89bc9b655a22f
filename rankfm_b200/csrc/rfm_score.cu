// rfm_score.cu -- scoring kernels on the fat-row tables (fp32, exact-parity path).
//
//   predict_kernel      replaces `_predict`   (rankfm/_rankfm.pyx:345-390): one lane group per (user,item) pair
//   score_users_kernel  replaces the all-item scoring loop of `_recommend` (:440-441): a lane group keeps one user's
//                       hoisted context (a[], b[]) in registers and streams item rows past it
//   topn_select_kernel  replaces `np.argsort(item_scores)[::-1]` + the seen-item walk (:444-456): per user an exact
//                       radix select of the n-th largest score, then a bitonic sort of the n survivors.  O(I) reads
//                       instead of an O(I log I) sort; ties broken towards the larger item index.
#include <cfloat>
#include "rfm_kernels.h"
#include "rfm_pair.cuh"

namespace rfm {

// ---------------------------------------------------------------------------------------------------------------
template <int G, int QPL, bool FEAT>
__global__ void __launch_bounds__(256) predict_kernel(const Tables T, const float2* __restrict__ pairs, long long n, float* __restrict__ scores)
{
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane % G, gw = lane / G;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
    for (long long base = warp_global * GPW; base < n; base += stride) {
        const long long r = base + gw;
        const bool inb = r < n;
        float2 pr = make_float2(0.f, 0.f);
        if (inb) pr = __ldg(pairs + r);
        // NaN index = cold-start id (:383-384); users whose rows this session does not hold (multi-GPU shard) score NaN too
        const bool known = inb && !isnan(pr.x) && !isnan(pr.y) && (int)pr.x >= T.u0 && (int)pr.x < T.u0 + T.Un;
        const int u = known ? (int)pr.x : 0, i = known ? (int)pr.y : 0;
        UserCtx<QPL> uc;
        ItemRow<QPL> it;
        load_user<G, QPL, FEAT>(T, u, known, sub, uc);
        load_item<G, QPL, FEAT>(T, i, known, sub, it);
        user_precompute<G, QPL, FEAT>(T, T.GP, known, sub, uc);
        const float s = utility<G, QPL, FEAT>(uc, it);
        if (inb && sub == 0) scores[r] = known ? s : __int_as_float(0x7fc00000);
    }
}

template <int G, int QPL>
static cudaError_t predict_gq(const Tables& T, const float2* pairs, long long n, float* scores, int grid, cudaStream_t st)
{
    if (T.x_uf_any || T.x_if_any) predict_kernel<G, QPL, true><<<grid, 256, 0, st>>>(T, pairs, n, scores);
    else predict_kernel<G, QPL, false><<<grid, 256, 0, st>>>(T, pairs, n, scores);
    return cudaGetLastError();
}

#define RFM_DISPATCH_GQ(FN, ...)                                                             \
    do {                                                                                     \
        int qpl = 1;                                                                         \
        const int G = train_group_size(T, &qpl);                                             \
        if (max(T.Pp, T.Qp) > 4 * G || qpl > 4) return cudaErrorInvalidValue;                \
        switch (G) {                                                                         \
            case 4:  return FN<4, 1>(__VA_ARGS__);                                           \
            case 8:  return FN<8, 1>(__VA_ARGS__);                                           \
            case 16: return FN<16, 1>(__VA_ARGS__);                                          \
            default:                                                                         \
                if (qpl == 1) return FN<32, 1>(__VA_ARGS__);                                 \
                if (qpl == 2) return FN<32, 2>(__VA_ARGS__);                                 \
                return FN<32, 4>(__VA_ARGS__);                                               \
        }                                                                                    \
    } while (0)

cudaError_t launch_predict(const Tables& T, const float2* pairs, long long n, float* scores, int grid, cudaStream_t st)
{
    RFM_DISPATCH_GQ(predict_gq, T, pairs, n, scores, grid, st);
}

// ---------------------------------------------------------------------------------------------------------------
// all-item scores of a batch of users: S[b, i] = u(users[b], i).  grid = (item chunks, users)
// ---------------------------------------------------------------------------------------------------------------
template <int G, int QPL, bool FEAT>
__global__ void __launch_bounds__(256) score_users_kernel(const Tables T, const int32_t* __restrict__ users, int n_users, float* __restrict__ S)
{
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane % G, gw = lane / G;
    const int b = blockIdx.y;
    const int u = __ldg(users + b);
    const bool known = u >= 0;
    UserCtx<QPL> uc;
    load_user<G, QPL, FEAT>(T, known ? u : 0, known, sub, uc);
    user_precompute<G, QPL, FEAT>(T, T.GP, known, sub, uc);
    const long long group_global = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gw;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
    float* out = S + (size_t)b * T.I;
    const long long span = ((long long)T.I + stride - 1) / stride * stride;   // warp-uniform trip count
    for (long long i = group_global; i < span; i += stride) {
        const bool ok = known && i < T.I;
        ItemRow<QPL> it;
        load_item<G, QPL, FEAT>(T, ok ? (int)i : 0, ok, sub, it);
        const float s = utility<G, QPL, FEAT>(uc, it);
        if (ok && sub == 0) out[i] = s;
    }
}

template <int G, int QPL>
static cudaError_t score_users_gq(const Tables& T, const int32_t* users, int n_users, float* S, int chunks, cudaStream_t st)
{
    const dim3 grid(chunks, n_users);
    if (T.x_uf_any || T.x_if_any) score_users_kernel<G, QPL, true><<<grid, 256, 0, st>>>(T, users, n_users, S);
    else score_users_kernel<G, QPL, false><<<grid, 256, 0, st>>>(T, users, n_users, S);
    return cudaGetLastError();
}

cudaError_t launch_score_users(const Tables& T, const int32_t* users, int n_users, float* S, int chunks, cudaStream_t st)
{
    RFM_DISPATCH_GQ(score_users_gq, T, users, n_users, S, chunks, st);
}

// ---------------------------------------------------------------------------------------------------------------
// exact top-n of each score row
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float s)   // larger score -> larger key; NaN sorts above +inf like np.argsort
{
    const uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

constexpr int kSelThreads = 512;

// One block per user row.  `S` row is consumed destructively when filter_previous is set (seen items -> -inf key 0).
// rec [n_users, n_items] receives item indexes as float32 (reference: rec_items, :426,453), NaN rows for unknown users.
__global__ void __launch_bounds__(kSelThreads) topn_select_kernel(float* __restrict__ S, int I, const int32_t* __restrict__ users,
                                                                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                                  int filter_previous, int n_items, float* __restrict__ rec,
                                                                  const int32_t* __restrict__ exclude, const int32_t* __restrict__ idxmap)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem_raw);        // [npow2] (key<<32 | index)
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining, s_count_gt, s_count_eq;
    __shared__ uint32_t warp_cnt[kSelThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int u = __ldg(users + b);
    float* out = rec + (size_t)b * n_items;
    if (u < 0) {
        for (int k = tid; k < n_items; k += kSelThreads) out[k] = __int_as_float(0x7fc00000);
        return;
    }
    float* row = S + (size_t)b * I;
    const uint32_t kRemoved = 0u;    // below every real key (order_key(-inf) = 0x007fffff)
    if (filter_previous) {
        const long long s0 = __ldg(indptr + u), s1 = __ldg(indptr + u + 1);
        for (long long k = s0 + tid; k < s1; k += kSelThreads) row[__ldg(indices + k)] = __uint_as_float(0xffffffffu);   // marker: -NaN payload
    }
    if (exclude) { if (tid == 0) row[__ldg(exclude + b)] = __uint_as_float(0xffffffffu); }
    __syncthreads();
    auto key_of = [&](int i) -> uint32_t {
        const uint32_t bits = __float_as_uint(row[i]);
        return bits == 0xffffffffu ? kRemoved : order_key(__uint_as_float(bits));
    };

    // radix select (MSB first, 8 bits per pass) of the n_items-th largest key
    __shared__ uint32_t s_short, s_eq_total;
    if (tid == 0) { s_prefix = 0u; s_remaining = (uint32_t)n_items; s_short = 0u; s_eq_total = 0u; }
    __syncthreads();
    for (int pass = 0; pass < 4 && !s_short; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int k = tid; k < 256; k += kSelThreads) hist[k] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < I; i += kSelThreads) {
            const uint32_t k = key_of(i);
            if (k != kRemoved && (k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t remaining = s_remaining, bin = 0u;
            bool hit = false;
            for (int d = 255; d >= 0; --d) {
                const uint32_t c = hist[d];
                if (c >= remaining) { bin = (uint32_t)d; hit = true; break; }
                remaining -= c;
            }
            if (!hit) { s_short = 1u; s_prefix = 0u; s_remaining = 0u; }   // fewer than n_items candidates: all are selected
            else { s_prefix = prefix | (bin << shift); s_remaining = remaining; s_eq_total = hist[bin]; }
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;                // key of the n-th largest (0 when the row is short: every key > 0)
    const uint32_t need_eq = s_remaining;         // how many of the keys == thr belong to the top n
    const bool ties_exact = s_eq_total == need_eq;   // no surplus ties: the claim order cannot matter
    if (tid == 0) { s_count_gt = 0u; s_count_eq = 0u; }
    int npow2 = 1;
    while (npow2 < n_items) npow2 <<= 1;
    for (int k = tid; k < npow2; k += kSelThreads) sel[k] = 0ull;          // key 0 sorts last
    __syncthreads();
    // ties at the threshold: prefer larger item index -> scan indexes descending so the first `need_eq` claimed are the largest
    for (int base = I - 1; base >= 0; base -= kSelThreads) {
        const int i = base - tid;
        if (i >= 0) {
            const uint32_t k = key_of(i);
            if (k != kRemoved) {
                if (k > thr || (ties_exact && k == thr)) {
                    const uint32_t slot = atomicAdd(&s_count_gt, 1u);
                    if (slot < (uint32_t)npow2) sel[slot] = ((unsigned long long)k << 32) | (uint32_t)i;
                }
            }
        }
    }
    __syncthreads();
    const uint32_t n_gt = min(s_count_gt, (uint32_t)n_items);
    for (int base = I - 1; base >= 0 && !ties_exact && !s_short; base -= kSelThreads) {   // surplus ties: ordered claim, largest index first
        const int i = base - tid;
        const bool eq = i >= 0 && key_of(i) == thr && thr != kRemoved;
        // ordered rank among the threads of this step: lower tid = larger index
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if ((tid & 31) == 0) warp_cnt[tid >> 5] = __popc(bal);
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < (tid >> 5); ++w) before += warp_cnt[w];
        uint32_t total = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) total += warp_cnt[w];
        const uint32_t my = s_count_eq + before + __popc(bal & ((1u << (tid & 31)) - 1u));
        const uint32_t cap = need_eq;
        if (eq && my < cap && n_gt + my < (uint32_t)npow2) sel[n_gt + my] = ((unsigned long long)thr << 32) | (uint32_t)i;
        __syncthreads();
        if (tid == 0) s_count_eq += total;
        __syncthreads();
        if (s_count_eq >= cap) break;
    }
    __syncthreads();
    // bitonic sort, descending by (key, index)
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < npow2; t += kSelThreads) {
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
    for (int k = tid; k < n_items; k += kSelThreads) {
        const unsigned long long e = sel[k];
        const uint32_t pos = (uint32_t)(e & 0xffffffffull);              // position in the row; shortlist rows map it to an item id
        out[k] = (e >> 32) == 0ull ? __int_as_float(0x7fc00000) : (float)(idxmap ? (uint32_t)idxmap[(size_t)b * I + pos] : pos);
    }
}

cudaError_t launch_topn_select(float* S, int I, const int32_t* users, int n_users, const int64_t* indptr, const int32_t* indices,
                               int filter_previous, int n_items, float* rec, const int32_t* exclude, cudaStream_t st, const int32_t* idxmap)
{
    int npow2 = 1;
    while (npow2 < n_items) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * sizeof(unsigned long long);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(topn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
    topn_select_kernel<<<n_users, kSelThreads, smem, st>>>(S, I, users, indptr, indices, filter_previous, n_items, rec, exclude, idxmap);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// similar_items / similar_users (rankfm/rankfm.py:405-454): rep(row) = v[row] + x[row].v_f ; score = rep(all).rep(query)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float latent_elem(const Tables& T, int which, int row, int f)
{
    if (which == 0) {
        const float* r = T.IT + (size_t)row * T.ldi;
        float v = r[f];
        if (T.x_if_any) for (int q = 0; q < T.Q; ++q) v += r[T.Fp + 4 + q] * T.GP[T.gp_vif + (size_t)q * T.Fp + f];
        return v;
    }
    const float* r = T.UT + (size_t)row * T.ldu;
    float v = r[f];
    if (T.x_uf_any) for (int p = 0; p < T.P; ++p) v += r[T.Fp + p] * T.GP[T.gp_vuf + (size_t)p * T.Fp + f];
    return v;
}

// qc = [ q (Fp floats): latent representation of the query row | c (Pp or Qp floats): c[k] = v_f[k] . q ]
// so that  rep(row) . rep(query) = v[row] . q + x[row] . c  -- one fat-row gather and one group reduction per row,
// exactly the shape of the FM utility without the item bias.
__global__ void latent_query_kernel(const Tables T, int which, int index, float* __restrict__ qc)
{
    for (int f = threadIdx.x; f < T.Fp; f += blockDim.x) qc[f] = f < T.F ? latent_elem(T, which, index, f) : 0.f;
    __syncthreads();
    const int nc = which == 0 ? T.Qp : T.Pp;
    const int n_real = which == 0 ? (T.x_if_any ? T.Q : 0) : (T.x_uf_any ? T.P : 0);
    const int base = which == 0 ? T.gp_vif : T.gp_vuf;
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        float acc = 0.f;
        if (k < n_real) for (int f = 0; f < T.F; ++f) acc += T.GP[base + (size_t)k * T.Fp + f] * qc[f];
        qc[T.Fp + k] = acc;
    }
}

template <int G, int QPL, bool FEAT>
__global__ void __launch_bounds__(256) latent_scores_kernel(const Tables T, int which, const float* __restrict__ qc, float* __restrict__ S)
{
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane % G, gw = lane / G;
    float4 a[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) { const int q = sub + k * G; a[k] = q < T.NQ ? reinterpret_cast<const float4*>(qc)[q] : zero4(); }
    float4 c = zero4();
    if (FEAT) { const int nc4 = (which == 0 ? T.Qp : T.Pp) / 4; if (sub < nc4) c = reinterpret_cast<const float4*>(qc + T.Fp)[sub]; }
    const long long n = which == 0 ? T.I : T.U;
    const long long group_global = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gw;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * GPW;
    const long long span = (n + stride - 1) / stride * stride;          // warp-uniform trip count
    for (long long row = group_global; row < span; row += stride) {
        const bool ok = row < n;
        float part = 0.f;
        if (which == 0) {
            ItemRow<QPL> r;
            load_item<G, QPL, FEAT>(T, ok ? (int)row : 0, ok, sub, r);
#pragma unroll
            for (int k = 0; k < QPL; ++k) part = dot4(a[k], r.v[k], part);
            if (FEAT) part = dot4(c, r.x, part);
        } else {
            UserCtx<QPL> u;
            load_user<G, QPL, FEAT>(T, ok ? (int)row : 0, ok, sub, u);
#pragma unroll
            for (int k = 0; k < QPL; ++k) part = dot4(a[k], u.vu[k], part);
            if (FEAT) part = dot4(c, u.xu, part);
        }
        const float sc = group_sum<G>(part);
        if (ok && sub == 0) S[row] = sc;
    }
}

template <int G, int QPL>
static cudaError_t latent_scores_gq(const Tables& T, int which, const float* qc, float* S, cudaStream_t st)
{
    const long long n = which == 0 ? T.I : T.U;
    const int grid = (int)max(1LL, min(148LL * 8, (n * G + 255) / 256));
    if (T.x_uf_any || T.x_if_any) latent_scores_kernel<G, QPL, true><<<grid, 256, 0, st>>>(T, which, qc, S);
    else latent_scores_kernel<G, QPL, false><<<grid, 256, 0, st>>>(T, which, qc, S);
    return cudaGetLastError();
}

// qc: scratch of Fp + max(Pp, Qp) floats
cudaError_t launch_latent_scores(const Tables& T, int which, int index, float* qc, float* S, cudaStream_t st)
{
    latent_query_kernel<<<1, 128, 0, st>>>(T, which, index, qc);
    RFM_DISPATCH_GQ(latent_scores_gq, T, which, qc, S, st);
}

// ---------------------------------------------------------------------------------------------------------------
// hold-out ranking metrics straight from the device top-k (rankfm/evaluation.py:9-143): one warp per evaluated user
// tests each recommended item against the user's sorted test items and adds the user's terms of the five metrics
//   out5 += [ any hit, 1/(rank of first hit + 1), sum_hits 1/log2(rank+2), hits/k, hits/|test items| ]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eval_topk_kernel(const float* __restrict__ rec, const int64_t* __restrict__ order, int n_users, int k,
                                                        const int64_t* __restrict__ test_indptr, const int32_t* __restrict__ test_items,
                                                        const int32_t* __restrict__ n_test, double* __restrict__ out5, uint8_t* __restrict__ hits_out)
{
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    double acc[5] = {0, 0, 0, 0, 0};
    for (long long b = warp_global; b < n_users; b += n_warps) {
        const long long orig = order ? order[b] : b;                  // row b of `rec` belongs to the orig-th requested user
        const long long s0 = test_indptr[orig], s1 = test_indptr[orig + 1];
        int cnt = 0, first = 0x7fffffff;
        double dcg = 0.0;
        for (int r = lane; r < k; r += 32) {
            const float item_f = rec[(size_t)b * k + r];
            bool hit = false;
            if (!isnan(item_f)) {
                const int item = (int)item_f;
                long long lo = s0, hi = s1 - 1;
                while (lo <= hi) {
                    const long long md = (lo + hi) >> 1;
                    const int e = __ldg(test_items + md);
                    if (e == item) { hit = true; break; }
                    if (e < item) lo = md + 1; else hi = md - 1;
                }
            }
            if (hit) { ++cnt; first = min(first, r); dcg += 1.0 / log2((double)r + 2.0); }
            if (hits_out) hits_out[(size_t)orig * k + r] = hit ? 1 : 0;
        }
        for (int off = 16; off > 0; off >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
            first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
            dcg += __shfl_xor_sync(0xffffffffu, dcg, off);
        }
        if (lane == 0 && cnt > 0) {
            const int nt = n_test[orig];
            acc[0] += 1.0; acc[1] += 1.0 / (double)(first + 1); acc[2] += dcg; acc[3] += (double)cnt / (double)k;
            acc[4] += nt > 0 ? (double)cnt / (double)nt : 0.0;
        }
    }
    if (lane == 0)
        for (int m = 0; m < 5; ++m) if (acc[m] != 0.0) atomicAdd(out5 + m, acc[m]);
}

cudaError_t launch_eval_topk(const float* rec, const int64_t* order, int n_users, int k, const int64_t* test_indptr, const int32_t* test_items,
                             const int32_t* n_test, double* out5, uint8_t* hits_out, cudaStream_t st)
{
    const int grid = max(1, min(148 * 8, (n_users + 7) / 8));
    eval_topk_kernel<<<grid, 256, 0, st>>>(rec, order, n_users, k, test_indptr, test_items, n_test, out5, hits_out);
    return cudaGetLastError();
}

// rows of `src` (in the planner's order: narrow tier, wide tier, exact path) -> rows order[k] of `dst` (the caller's order)
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ order, long long n_rows, int n_items, float* __restrict__ dst)
{
    const long long total = n_rows * n_items;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / n_items;
        dst[(size_t)__ldg(order + k) * n_items + (e - k * n_items)] = src[e];
    }
}

cudaError_t launch_scatter_rows(const float* src, const int64_t* order, long long n_rows, int n_items, float* dst, cudaStream_t st)
{
    scatter_rows_kernel<<<148 * 8, 256, 0, st>>>(src, order, n_rows, n_items, dst);
    return cudaGetLastError();
}

}  // namespace rfm
