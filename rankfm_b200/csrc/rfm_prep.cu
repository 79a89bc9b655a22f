// rfm_prep.cu -- data preparation on the device.
//
// The reference prepares its inputs on the host with pandas: id -> index maps (`rankfm.py:114-128`), index pairs
// (`:150-155`), and `user_items` through a groupby + Python set/sort per user (`:165-174`), then `_fit` copies that dict
// into a ragged C table element by element (`_rankfm.pyx:201-212`): 28-44 s at ~1 M users / 4.7 M interactions (SURVEY
// Appendix C).  Here the same objects are three radix sorts on the GPU:
//   rfm_prep_index_ids    sorted unique ids + the index of every id in that list            (np.unique + Series.map)
//   rfm_prep_user_items   CSR of each user's items, ascending, duplicates kept             (groupby -> dict of sorted arrays)
//   rfm_synth_zipf        synthetic Zipf interactions of the BASELINE.json shapes (bench tooling: the NumPy generator in
//                         rankfm_b200/synthetic.py needs 50-80 s for the 50-62 M-interaction configurations)
// CUB (device radix sort / select / scan) is library code; none of this is on the training or scoring path.
#include <cmath>
#include <cstdio>
#include <vector>

#include <cub/cub.cuh>

#include "rfm_host.h"
#include "rfm_rng.cuh"

using rfmh::fail;

namespace {

struct Buf {                                 // scope-owned device allocation (plain cudaMalloc: sizes vary from call to call)
    void* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) cudaStreamDestroy(s); }
};

int bits_for(uint64_t max_value) { int b = 1; while (b < 64 && (max_value >> b)) ++b; return b; }

int pick_device(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail(RFM_ERR_NO_DEVICE, "no CUDA device: rankfm_b200 has no CPU fallback"); }
    if (device < 0 || device >= n) return fail(RFM_ERR_ARG, "device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    return RFM_OK;
}

template <typename K>
int sort_keys(K* in, K* out, int64_t n, int end_bit, cudaStream_t st)
{
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n, 0, end_bit, st));
    Buf tmp;
    CU(tmp.alloc(tmp_bytes));
    CU(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, in, out, n, 0, end_bit, st));
    CU(cudaStreamSynchronize(st));
    return RFM_OK;
}

template <typename K, typename V>
int sort_pairs(K* kin, K* kout, V* vin, V* vout, int64_t n, int end_bit, cudaStream_t st)
{
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, n, 0, end_bit, st));
    Buf tmp;
    CU(tmp.alloc(tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kin, kout, vin, vout, n, 0, end_bit, st));
    CU(cudaStreamSynchronize(st));
    return RFM_OK;
}

template <typename K>
int unique_sorted(K* in, K* out, int64_t n, int64_t* n_out_host, cudaStream_t st)
{
    size_t tmp_bytes = 0;
    Buf cnt, tmp;
    CU(cnt.alloc(8));
    CU(cub::DeviceSelect::Unique(nullptr, tmp_bytes, in, out, cnt.as<int64_t>(), n, st));
    CU(tmp.alloc(tmp_bytes));
    CU(cub::DeviceSelect::Unique(tmp.p, tmp_bytes, in, out, cnt.as<int64_t>(), n, st));
    CU(cudaMemcpyAsync(n_out_host, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RFM_OK;
}

template <typename T>
int exclusive_sum(T* in, T* out, int64_t n, cudaStream_t st)
{
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, st));
    Buf tmp;
    CU(tmp.alloc(tmp_bytes));
    CU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, st));
    return RFM_OK;
}

constexpr int kThreads = 256;
inline int grid_for(int64_t n) { const int64_t g = (n + kThreads - 1) / kThreads; return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); }

// ---- user_items ----
__global__ void pair_keys_kernel(const int2* __restrict__ inter, int64_t n, int32_t U, int32_t I, uint64_t* __restrict__ keys, unsigned long long* __restrict__ counts, int* __restrict__ bad)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int2 ui = inter[e];
        if (ui.x < 0 || ui.x >= U || ui.y < 0 || ui.y >= I) { *bad = 1; keys[e] = ~0ull; continue; }
        keys[e] = (uint64_t)ui.x * (uint64_t)I + (uint64_t)ui.y;
        atomicAdd(counts + ui.x, 1ull);
    }
}
__global__ void key_items_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t I, int32_t* __restrict__ items)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) items[e] = (int32_t)(keys[e] % (uint64_t)I);
}

// ---- id maps ----
__global__ void lower_bound_kernel(const int64_t* __restrict__ ids, int64_t n, const int64_t* __restrict__ uniq, int64_t m, int32_t* __restrict__ index)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ids[e];
        int64_t lo = 0, hi = m;
        while (lo < hi) { const int64_t md = (lo + hi) >> 1; if (uniq[md] < v) lo = md + 1; else hi = md; }
        index[e] = (lo < m && uniq[lo] == v) ? (int32_t)lo : -1;
    }
}

// ---- synthetic Zipf interactions ----
__device__ __forceinline__ int32_t cdf_rank(const double* __restrict__ cdf, int32_t n, double x)       // np.searchsorted(cdf, x), clipped
{
    int32_t lo = 0, hi = n;
    while (lo < hi) { const int32_t md = (lo + hi) >> 1; if (cdf[md] < x) lo = md + 1; else hi = md; }
    return lo < n ? lo : n - 1;
}
__global__ void zipf_draw_kernel(const double* __restrict__ cdf_u, int32_t U, const double* __restrict__ cdf_i, int32_t I, int64_t n, uint32_t k0, uint32_t k1, uint32_t round,
                                 uint64_t* __restrict__ keys)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const rfm::Philox4 r = rfm::philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), round, 0x5a17u, k0, k1);
        const double xu = ((double)r.x * 4294967296.0 + (double)r.y + 0.5) * (1.0 / 18446744073709551616.0);
        const double xi = ((double)r.z * 4294967296.0 + (double)r.w + 0.5) * (1.0 / 18446744073709551616.0);
        keys[e] = (uint64_t)cdf_rank(cdf_u, U, xu) * (uint64_t)I + (uint64_t)cdf_rank(cdf_i, I, xi);
    }
}
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__global__ void tag_kernel(const uint64_t* __restrict__ keys, int64_t n, uint64_t seed, uint64_t* __restrict__ tags)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) tags[e] = mix64(keys[e] ^ seed);
}
__global__ void iota_tag_kernel(int32_t n, uint64_t seed, uint64_t* __restrict__ tags, int32_t* __restrict__ idx)
{
    for (int32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) { tags[e] = mix64((uint64_t)e * 0x9E3779B97F4A7C15ull + seed); idx[e] = e; }
}
__global__ void mark_present_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t I, const int32_t* __restrict__ perm_u, const int32_t* __restrict__ perm_i,
                                    int32_t* __restrict__ present_u, int32_t* __restrict__ present_i)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[e];
        present_u[perm_u[k / (uint64_t)I]] = 1;
        present_i[perm_i[k % (uint64_t)I]] = 1;
    }
}
__global__ void emit_pairs_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t I, const int32_t* __restrict__ perm_u, const int32_t* __restrict__ perm_i,
                                  const int32_t* __restrict__ rank_u, const int32_t* __restrict__ rank_i, int32_t reindex, int32_t user_offset, int2* __restrict__ out)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[e];
        const int32_t u = perm_u[k / (uint64_t)I], i = perm_i[k % (uint64_t)I];
        out[e] = reindex ? make_int2(rank_u[u] + user_offset, rank_i[i]) : make_int2(u + user_offset, i);
    }
}

std::vector<double> zipf_cdf(int32_t n, double a)
{
    std::vector<double> c((size_t)n);
    double tot = 0.0;
    for (int32_t k = 0; k < n; ++k) { tot += 1.0 / std::pow((double)(k + 1), a); c[(size_t)k] = tot; }
    for (auto& v : c) v /= tot;
    return c;
}

// random permutation of [0, n) on the device: sort the indexes by a hash tag
int random_permutation(int32_t n, uint64_t seed, Buf& perm, cudaStream_t st)
{
    Buf t0, t1, i0;
    CU(t0.alloc((size_t)n * 8)); CU(t1.alloc((size_t)n * 8)); CU(i0.alloc((size_t)n * 4)); CU(perm.alloc((size_t)n * 4));
    iota_tag_kernel<<<grid_for(n), kThreads, 0, st>>>(n, seed, t0.as<uint64_t>(), i0.as<int32_t>());
    return sort_pairs(t0.as<uint64_t>(), t1.as<uint64_t>(), i0.as<int32_t>(), perm.as<int32_t>(), n, 64, st);
}

}  // namespace

extern "C" int rfm_prep_user_items(const int32_t* interactions, int64_t n, int32_t U, int32_t I, int32_t device, int64_t* indptr_out, int32_t* indices_out)
{
    if (!interactions || !indptr_out || (n > 0 && !indices_out) || n < 0 || U < 1 || I < 1) return fail(RFM_ERR_ARG, "bad argument");
    int rc = pick_device(device);
    if (rc) return rc;
    Stream st;
    CU(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    Buf inter, k0, k1, counts, ptr, items, bad;
    CU(inter.alloc((size_t)n * 8)); CU(k0.alloc((size_t)n * 8)); CU(k1.alloc((size_t)n * 8));
    CU(counts.alloc(((size_t)U + 1) * 8)); CU(ptr.alloc(((size_t)U + 1) * 8)); CU(items.alloc((size_t)n * 4)); CU(bad.alloc(4));
    CU(cudaMemcpyAsync(inter.p, interactions, (size_t)n * 8, cudaMemcpyHostToDevice, st.s));
    CU(cudaMemsetAsync(counts.p, 0, ((size_t)U + 1) * 8, st.s));
    CU(cudaMemsetAsync(bad.p, 0, 4, st.s));
    pair_keys_kernel<<<grid_for(n), kThreads, 0, st.s>>>(inter.as<int2>(), n, U, I, k0.as<uint64_t>(), counts.as<unsigned long long>(), bad.as<int>());
    CU(cudaGetLastError());
    int bad_h = 0;
    CU(cudaMemcpyAsync(&bad_h, bad.p, 4, cudaMemcpyDeviceToHost, st.s));
    if ((rc = sort_keys(k0.as<uint64_t>(), k1.as<uint64_t>(), n, bits_for((uint64_t)U * (uint64_t)I), st.s))) return rc;
    if (bad_h) return fail(RFM_ERR_ARG, "interactions hold an index outside [0,%d) x [0,%d)", U, I);
    key_items_kernel<<<grid_for(n), kThreads, 0, st.s>>>(k1.as<uint64_t>(), n, I, items.as<int32_t>());
    if ((rc = exclusive_sum(counts.as<int64_t>(), ptr.as<int64_t>(), (int64_t)U + 1, st.s))) return rc;
    CU(cudaMemcpyAsync(indptr_out, ptr.p, ((size_t)U + 1) * 8, cudaMemcpyDeviceToHost, st.s));
    CU(cudaMemcpyAsync(indices_out, items.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st.s));
    CU(cudaStreamSynchronize(st.s));
    return RFM_OK;
}

extern "C" int rfm_prep_index_ids(const int64_t* ids, int64_t n, int32_t device, int64_t* unique_out, int64_t* n_unique_out, int32_t* index_out)
{
    if (!ids || !unique_out || !n_unique_out || !index_out || n < 1) return fail(RFM_ERR_ARG, "bad argument");
    int rc = pick_device(device);
    if (rc) return rc;
    Stream st;
    CU(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    Buf raw, sorted, uniq, index;
    CU(raw.alloc((size_t)n * 8)); CU(sorted.alloc((size_t)n * 8)); CU(uniq.alloc((size_t)n * 8)); CU(index.alloc((size_t)n * 4));
    CU(cudaMemcpyAsync(raw.p, ids, (size_t)n * 8, cudaMemcpyHostToDevice, st.s));
    if ((rc = sort_keys(raw.as<int64_t>(), sorted.as<int64_t>(), n, 64, st.s))) return rc;          // signed keys: CUB orders them numerically
    int64_t m = 0;
    if ((rc = unique_sorted(sorted.as<int64_t>(), uniq.as<int64_t>(), n, &m, st.s))) return rc;
    lower_bound_kernel<<<grid_for(n), kThreads, 0, st.s>>>(raw.as<int64_t>(), n, uniq.as<int64_t>(), m, index.as<int32_t>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(unique_out, uniq.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st.s));
    CU(cudaMemcpyAsync(index_out, index.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st.s));
    CU(cudaStreamSynchronize(st.s));
    *n_unique_out = m;
    return RFM_OK;
}

extern "C" int rfm_synth_zipf(int32_t U, int32_t I, int64_t N, double a_u, double a_i, uint64_t seed, uint64_t perm_seed, int32_t reindex, int32_t user_offset,
                              int32_t device, int32_t* out, int64_t* n_out, int32_t* users_out, int32_t* items_out)
{
    if (!out || !n_out || !users_out || !items_out || U < 1 || I < 1 || N < 1) return fail(RFM_ERR_ARG, "bad argument");
    int rc = pick_device(device);
    if (rc) return rc;
    Stream st;
    CU(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    const std::vector<double> cu = zipf_cdf(U, a_u), ci = zipf_cdf(I, a_i);
    Buf d_cu, d_ci;
    CU(d_cu.alloc((size_t)U * 8)); CU(d_ci.alloc((size_t)I * 8));
    CU(cudaMemcpyAsync(d_cu.p, cu.data(), (size_t)U * 8, cudaMemcpyHostToDevice, st.s));
    CU(cudaMemcpyAsync(d_ci.p, ci.data(), (size_t)I * 8, cudaMemcpyHostToDevice, st.s));
    // draw, de-duplicate, repeat until N distinct (user, item) pairs exist
    const int64_t draw = N + N / 4 + 1024;
    const int key_bits = bits_for((uint64_t)U * (uint64_t)I);
    Buf uniq;                                 // distinct keys so far, sorted
    int64_t n_uniq = 0;
    for (uint32_t round = 0; round < 8 && n_uniq < N; ++round) {
        const int64_t tot = n_uniq + draw;
        Buf a, b, u2;
        CU(a.alloc((size_t)tot * 8)); CU(b.alloc((size_t)tot * 8)); CU(u2.alloc((size_t)tot * 8));
        if (n_uniq) CU(cudaMemcpyAsync(a.p, uniq.p, (size_t)n_uniq * 8, cudaMemcpyDeviceToDevice, st.s));
        zipf_draw_kernel<<<grid_for(draw), kThreads, 0, st.s>>>(d_cu.as<double>(), U, d_ci.as<double>(), I, draw, (uint32_t)seed, (uint32_t)(seed >> 32), round, a.as<uint64_t>() + n_uniq);
        CU(cudaGetLastError());
        if ((rc = sort_keys(a.as<uint64_t>(), b.as<uint64_t>(), tot, key_bits, st.s))) return rc;
        if ((rc = unique_sorted(b.as<uint64_t>(), u2.as<uint64_t>(), tot, &n_uniq, st.s))) return rc;
        if (uniq.p) { cudaFree(uniq.p); }
        uniq.p = u2.p; u2.p = nullptr;
    }
    const int64_t n_keep = n_uniq < N ? n_uniq : N;
    // unbiased trim to N + final shuffle in one pass: order the distinct pairs by a hash tag, keep the first N
    Buf t0, t1, kk;
    CU(t0.alloc((size_t)n_uniq * 8)); CU(t1.alloc((size_t)n_uniq * 8)); CU(kk.alloc((size_t)n_uniq * 8));
    tag_kernel<<<grid_for(n_uniq), kThreads, 0, st.s>>>(uniq.as<uint64_t>(), n_uniq, seed * 0x9E3779B97F4A7C15ull + 77ull, t0.as<uint64_t>());
    if ((rc = sort_pairs(t0.as<uint64_t>(), t1.as<uint64_t>(), uniq.as<uint64_t>(), kk.as<uint64_t>(), n_uniq, 64, st.s))) return rc;
    // popularity must not be index-ordered: random permutation of both id spaces, then re-index to the observed uniques
    // like rankfm.py:115-116 does for real data
    Buf perm_u, perm_i, pres_u, pres_i, rank_u, rank_i, pairs;
    if ((rc = random_permutation(U, perm_seed + 101, perm_u, st.s))) return rc;
    if ((rc = random_permutation(I, perm_seed + 202, perm_i, st.s))) return rc;
    CU(pres_u.alloc(((size_t)U + 1) * 4)); CU(pres_i.alloc(((size_t)I + 1) * 4)); CU(rank_u.alloc(((size_t)U + 1) * 4)); CU(rank_i.alloc(((size_t)I + 1) * 4));
    CU(cudaMemsetAsync(pres_u.p, 0, ((size_t)U + 1) * 4, st.s));
    CU(cudaMemsetAsync(pres_i.p, 0, ((size_t)I + 1) * 4, st.s));
    mark_present_kernel<<<grid_for(n_keep), kThreads, 0, st.s>>>(kk.as<uint64_t>(), n_keep, I, perm_u.as<int32_t>(), perm_i.as<int32_t>(), pres_u.as<int32_t>(), pres_i.as<int32_t>());
    if ((rc = exclusive_sum(pres_u.as<int32_t>(), rank_u.as<int32_t>(), (int64_t)U + 1, st.s))) return rc;
    if ((rc = exclusive_sum(pres_i.as<int32_t>(), rank_i.as<int32_t>(), (int64_t)I + 1, st.s))) return rc;
    CU(pairs.alloc((size_t)n_keep * 8));
    emit_pairs_kernel<<<grid_for(n_keep), kThreads, 0, st.s>>>(kk.as<uint64_t>(), n_keep, I, perm_u.as<int32_t>(), perm_i.as<int32_t>(), rank_u.as<int32_t>(), rank_i.as<int32_t>(),
                                                              reindex, user_offset, pairs.as<int2>());
    CU(cudaGetLastError());
    int32_t nu = 0, ni = 0;
    CU(cudaMemcpyAsync(&nu, rank_u.as<int32_t>() + U, 4, cudaMemcpyDeviceToHost, st.s));
    CU(cudaMemcpyAsync(&ni, rank_i.as<int32_t>() + I, 4, cudaMemcpyDeviceToHost, st.s));
    CU(cudaMemcpyAsync(out, pairs.p, (size_t)n_keep * 8, cudaMemcpyDeviceToHost, st.s));
    CU(cudaStreamSynchronize(st.s));
    *n_out = n_keep; *users_out = nu; *items_out = ni;
    return RFM_OK;
}
