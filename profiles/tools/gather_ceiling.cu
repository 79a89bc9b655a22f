// gather_ceiling.cu -- what HBM3e sustains for the SGD kernel's ACCESS PATTERN (not for a streaming copy):
//   mode 0: random fat-row gather        (TMA bulk copy global->shared, D stages per warp in flight)
//   mode 1: random fat-row read-modify-write (gather, then TMA bulk reduce-add of the same row back)
// Rows are `row_bytes` long and uniformly random over a table much larger than L2.  The result is the practical
// ceiling that `roofline.frac` of sgd_pipe_kernel should be read against next to the streaming-copy peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_ceiling gather_ceiling.cu && ./gather_ceiling
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(float* table, const int* __restrict__ idx, long long n, int row_floats, int D)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row_bytes = row_floats * 4u;
    unsigned char* wbase = smem + (size_t)warp * (128 + (size_t)D * row_bytes);
    const uint32_t bars = smem_u32(wbase), slots = smem_u32(wbase + 128);
    if (lane == 0) {
        for (int d = 0; d < D; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * d));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // each warp owns a contiguous chunk of the index stream; indices are fetched 32 at a time (one per lane) and handed
    // to the issuing lane by shuffle, so no global load sits on the issue path
    const long long wg = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
    const long long per = (n + nw - 1) / nw, lo = wg * per, hi = lo + per < n ? lo + per : n;
    uint32_t si = 0, sc = 0, par = 0;
    long long issue_i = lo, cons_i = lo;
    int cur_idx = 0, cons_idx = 0;
    auto next_issue_row = [&]() -> int {
        const int o = (int)((issue_i - lo) & 31);
        if (o == 0) cur_idx = (issue_i + lane < hi) ? idx[issue_i + lane] : 0;
        return __shfl_sync(0xffffffffu, cur_idx, o);
    };
    auto issue = [&]() {
        const int row = next_issue_row();
        if (lane == 0) {
            const float* src = table + (size_t)row * row_floats;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8u * si), "r"(row_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(slots + si * row_bytes), "l"(src), "r"(row_bytes), "r"(bars + 8u * si) : "memory");
        }
        ++issue_i; si = si + 1 == (uint32_t)D ? 0 : si + 1;
    };
    for (int d = 0; d < D && issue_i < hi; ++d) issue();
    while (cons_i < hi) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bars + 8u * sc), "r"(par) : "memory");
        if (MODE == 1) {
            const int o = (int)((cons_i - lo) & 31);
            if (o == 0) cons_idx = (cons_i + lane < hi) ? idx[cons_i + lane] : 0;
            const int row = __shfl_sync(0xffffffffu, cons_idx, o);
            if (lane == 0) {
                float* dst = table + (size_t)row * row_floats;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(slots + sc * row_bytes), "r"(row_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
        }
        ++cons_i;
        if (++sc == (uint32_t)D) { sc = 0; par ^= 1u; }
        if (issue_i < hi) issue();
    }
    if (MODE == 1 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv)
{
    const int row_floats = argc > 1 ? atoi(argv[1]) : 132;          // 528-byte item rows of factors=128
    const long long rows = argc > 2 ? atoll(argv[2]) : 2400000;     // 1.27 GB table
    const long long n = argc > 3 ? atoll(argv[3]) : 48000000;       // row touches per launch (= 16M positives x 3 rows)
    const int D = argc > 4 ? atoi(argv[4]) : 8;
    float* table; int* idx;
    CK(cudaMalloc(&table, (size_t)rows * row_floats * 4));
    CK(cudaMemset(table, 0, (size_t)rows * row_floats * 4));
    std::vector<int> h(n);
    uint64_t s = 88172645463325252ull;
    for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)rows); }
    CK(cudaMalloc(&idx, n * 4));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 8 * (128 + (size_t)D * row_floats * 4);
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int mode = 0; mode < 2; ++mode) {
        for (int bps = 1; bps <= 3; ++bps) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(a));
                if (mode == 0) k<0><<<sms * bps, 256, smem>>>(table, idx, n, row_floats, D);
                else k<1><<<sms * bps, 256, smem>>>(table, idx, n, row_floats, D);
                CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
                float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
            }
            const double bytes = (double)n * row_floats * 4 * (mode == 0 ? 1 : 2);
            printf("{\"mode\": \"%s\", \"row_bytes\": %d, \"blocks_per_sm\": %d, \"depth\": %d, \"ms\": %.3f, \"row_touches_per_s\": %.3e, \"dram_GBps\": %.1f}\n",
                   mode == 0 ? "gather" : "gather+reduce", row_floats * 4, bps, D, best, n / (best * 1e-3), bytes / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
