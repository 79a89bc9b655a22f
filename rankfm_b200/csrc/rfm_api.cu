// rfm_api.cu -- the C ABI (include/rankfm_b200.h): sessions, H2D/D2H, epoch loop, multi-GPU delta exchange.
//
// Host-side driver of the kernels; replaces the Python/Cython set-up and epoch loop of `_fit`
// (rankfm/_rankfm.pyx:182-228, 328-342), `_predict` (:345-390) and `_recommend` (:393-460).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "rfm_host.h"
#include <cuda_bf16.h>

using namespace rfm;
using rfmh::fail;
using rfmh::dev_malloc;
using rfmh::dev_free;

namespace rfm { bool gemm_encode_available(); }
static bool encode_ok() { return rfm::gemm_encode_available(); }

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

int rfmh::fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" const char* rfm_version(void) { return "rankfm_b200 0.1.0 (sm_100a)"; }
extern "C" const char* rfm_last_error(void) { return g_err.c_str(); }

extern "C" int rfm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
// Device block cache.  The one-shot entry points (`rfm_fit`, `rfm_predict`, `rfm_recommend`: the reference's stateless
// `_fit` / `_predict` / `_recommend`) create and destroy a session per call; cudaMalloc + cudaFree of its ~15 buffers cost
// 6-40 ms per call on the cfg2 workload and cudaFree stalled for 100-230 ms every few calls (profiles/: e2e probe,
// RANKFM_B200_TIMING=1) -- against 22 ms of training.  Freed blocks are therefore kept (exact size classes, up to
// RANKFM_B200_CACHE_MB, default 16384 MiB per process -- a tenth of a B200's HBM; a cfg4-sized shard needs ~4 GB and
// releasing its blocks to the driver cost 0.5 s per call) and handed back to the next call of the same shape.
// dev_free() synchronises the device like cudaFree does, so a cached block is never still in use by queued work.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct BlockCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> idle;                   // (device, bytes) -> block
    std::unordered_map<void*, std::pair<int, size_t>> owner;             // every block handed out or idle
    size_t idle_bytes = 0;
    size_t limit() const
    {
        static const size_t v = [] { const char* e = getenv("RANKFM_B200_CACHE_MB"); return (size_t)(e ? atoll(e) : 16384) << 20; }();
        return v;
    }
};
BlockCache g_blocks;

size_t size_class(size_t bytes) { const size_t g = bytes >= ((size_t)1 << 20) ? ((size_t)1 << 20) : 4096; return (bytes + g - 1) / g * g; }

void trim_locked(int dev)
{
    for (auto it = g_blocks.idle.begin(); it != g_blocks.idle.end();) {
        if (dev >= 0 && it->first.first != dev) { ++it; continue; }
        g_blocks.idle_bytes -= it->first.second;
        g_blocks.owner.erase(it->second);
        cudaFree(it->second);
        it = g_blocks.idle.erase(it);
    }
}
}  // namespace

cudaError_t rfmh::dev_malloc(void** out, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t cls = size_class(bytes ? bytes : 1);
    std::lock_guard<std::mutex> lock(g_blocks.mu);
    auto it = g_blocks.idle.find({dev, cls});
    if (it != g_blocks.idle.end()) {
        *out = it->second;
        g_blocks.idle_bytes -= cls;
        g_blocks.idle.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, cls);
    if (e == cudaErrorMemoryAllocation) {                                // give the idle blocks back and try once more
        cudaGetLastError();
        trim_locked(dev);
        e = cudaMalloc(out, cls);
    }
    if (e == cudaSuccess) g_blocks.owner[*out] = {dev, cls};
    return e;
}

void rfmh::dev_free(void* p)
{
    if (!p) return;
    std::unique_lock<std::mutex> lock(g_blocks.mu);
    auto it = g_blocks.owner.find(p);
    if (it == g_blocks.owner.end()) { lock.unlock(); cudaFree(p); return; }
    const int dev = it->second.first;
    const size_t cls = it->second.second;
    if (g_blocks.idle_bytes + cls > g_blocks.limit()) { g_blocks.owner.erase(it); lock.unlock(); cudaFree(p); return; }
    lock.unlock();
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    cudaDeviceSynchronize();                                              // what cudaFree would have waited for
    if (cur != dev) cudaSetDevice(cur);
    lock.lock();
    g_blocks.idle.insert({{dev, cls}, p});
    g_blocks.idle_bytes += cls;
}

extern "C" int rfm_trim_device_cache(void)
{
    std::lock_guard<std::mutex> lock(g_blocks.mu);
    trim_locked(-1);
    return RFM_OK;
}

extern "C" int rfm_nccl_unique_id(uint8_t* out128)
{
    if (!out128) return fail(RFM_ERR_ARG, "out128 is NULL");
    return rfmh::nccl_unique_id(out128);
}

extern "C" int rfm_host_register(void* ptr, uint64_t bytes)
{
    if (!ptr || bytes == 0) return fail(RFM_ERR_ARG, "NULL buffer");
    if (rfm_device_count() == 0) return fail(RFM_ERR_NO_DEVICE, "no CUDA device");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RFM_OK; }
    if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
    return RFM_OK;
}

extern "C" int rfm_host_unregister(void* ptr)
{
    if (!ptr) return fail(RFM_ERR_ARG, "NULL buffer");
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RFM_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return RFM_OK;
}

extern "C" int rfm_debug_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4)
{
    if (!out4) return fail(RFM_ERR_ARG, "out4 is NULL");
    const Philox4 b = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out4[0] = b.x; out4[1] = b.y; out4[2] = b.z; out4[3] = b.w;
    return RFM_OK;
}

extern "C" int rfm_debug_feat8(int32_t F, int32_t P, int32_t Q, uint32_t seed, float* out5)
{
    if (!out5 || F < 1 || P < 0 || Q < 0 || (P == 0 && Q == 0)) return fail(RFM_ERR_ARG, "bad argument");
    if (rfm_device_count() == 0) return fail(RFM_ERR_NO_DEVICE, "no CUDA device: rankfm_b200 has no CPU fallback");
    Tables T{};
    T.U = 1; T.I = 2; T.Un = 1; T.F = F; T.P = std::max(P, 1); T.Q = std::max(Q, 1);
    T.x_uf_any = P > 0; T.x_if_any = Q > 0;
    T.Fp = (F + 3) & ~3; T.NQ = T.Fp / 4;
    T.Pp = P > 0 ? (P + 3) & ~3 : 0; T.Qp = Q > 0 ? (Q + 3) & ~3 : 0;
    T.ldu = T.Fp + T.Pp; T.ldi = T.Fp + 4 + T.Qp;
    T.gp_vuf = (T.Q + 3) & ~3; T.gp_vif = T.gp_vuf + T.P * T.Fp;
    float* d = nullptr;
    const float init[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.9f};      // [6..7]: the WARP multiplier table the step reads
    CU(cudaMalloc(&d, sizeof init));
    CU(cudaMemcpy(d, init, sizeof init, cudaMemcpyHostToDevice));
    cudaError_t e = launch_feat8_selftest(T, seed, 0.05f, 0.2f, d, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out5, d, 5 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorInvalidValue ? RFM_ERR_UNSUPPORTED : RFM_ERR_CUDA, "feat8 self-test: %s", cudaGetErrorString(e)); }
    return RFM_OK;
}

// rank of the block bound that becomes the row threshold of the tensor-core recommend path (rfm_kernels.h: tau_rank)
extern "C" int rfm_debug_tau_rank(int32_t want, int32_t sample_k, float z) { return rfm::tau_rank(want, sample_k, z); }

extern "C" int rfm_debug_feistel(int64_t n, uint64_t seed, int32_t epoch, int64_t r0, int64_t count, int64_t* out)
{
    if (!out || n < 1 || r0 < 0 || r0 + count > n) return fail(RFM_ERR_ARG, "bad feistel range");
    const Feistel f = make_feistel(n, seed, epoch);
    for (int64_t k = 0; k < count; ++k) out[k] = feistel_perm(f, r0 + k);
    return RFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------------------------------
__global__ void item_histogram_kernel(const int2* __restrict__ inter, long long n, int32_t* __restrict__ count);

struct rfm_session {
    rfm_problem p{};            // host pointers are NOT retained past create (copied fields only)
    Tables T{};
    int device = 0, n_sm = 148;
    cudaStream_t st = nullptr;
    cudaStream_t st_side = nullptr;     // lower-priority helper stream of the tensor-core recommend pipeline (created on first use)
    // data
    int2* d_inter = nullptr; float* d_sw = nullptr; int64_t* d_indptr = nullptr; int32_t* d_indices = nullptr;
    int64_t N = 0, nnz = 0;
    int32_t* d_perm = nullptr;
    uint32_t* d_bitmap = nullptr; int bitmap_words = 0;
    float* d_mult = nullptr;
    MtState* d_mt = nullptr;
    EpochAcc* d_acc = nullptr; int acc_cap = 0;
    size_t gp_floats = 0;
    int epochs_done = 0;
    int grid = 148;
    int64_t launches = 0;
    // real allocations behind the tables.  T.UT, d_indptr, d_indices and d_bitmap are VIRTUAL bases when the session holds
    // only the users [T.u0, T.u0+T.Un) of a multi-GPU job: indexing them with a global user id lands in
    // the allocation; T.IT / T.GP / d_item_touch sit in the communicator's peer window when `p2p`
    float* ut_alloc = nullptr; int64_t* indptr_alloc = nullptr; int32_t* indices_alloc = nullptr; uint32_t* bitmap_alloc = nullptr;
    uint32_t *bloom_alloc = nullptr, *d_bloom = nullptr;   // membership pre-filter when there is no bitmap (one word per CSR entry)
    // multi-GPU
    rfmh::Comm* comm = nullptr;
    bool p2p = false;
    float *d_it_snap = nullptr, *d_gp_snap = nullptr;
    double* d_red = nullptr; int red_cap = 0;       // end-of-training reduction of the epoch records over the ranks
    float* d_gp_acc = nullptr;
    float *d_xuf = nullptr, *d_xif = nullptr;   // device copies of x_uf [U,P] / x_if [I,Q] (only when that block is active)
    int32_t* d_item_touch = nullptr;    // multi-GPU: how often each item occurs as a positive in this rank's shard
    // scratch
    float *d_snap_ut = nullptr, *d_snap_it = nullptr, *d_snap_gp = nullptr; int snap_epochs = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int32_t* d_trace = nullptr;
    // tensor-core recommend: bf16 item operand + bias, rebuilt lazily whenever the weights change
    // (items in descending bias order: d_gemm_order maps a position of B / bias back to the item index)
    void* d_gemm_B = nullptr; float* d_gemm_bias = nullptr; int32_t* d_gemm_order = nullptr; int gemm_I_pad = 0; bool gemm_valid = false;
    void* scratch[12] = {nullptr}; size_t scratch_bytes[12] = {0};   // grow-only device scratch of the recommend paths
    int64_t tc_rows = 0, tc_redo = 0;   // tensor-core recommend: rows served / rows redone on the exact path (candidate overflow)
    int64_t tc_retry = 0;               // ... rows redone with the provable threshold (estimated threshold too high)
    bool tau_spec_off = false;          // the speculative row threshold failed too often on this catalogue: conservative one until the weights change
    bool tau_spec_ok = false;           // ... or was verified on a batch of this weight state
    std::vector<int64_t> h_indptr;      // host copy of the CSR row pointers (degrees for the recommend planner)
    float* d_flush = nullptr; size_t flush_bytes = 0;
    std::vector<cudaEvent_t> ev;
    rfm_epoch_callback epoch_cb = nullptr; void* epoch_cb_user = nullptr;
};

static size_t round4(size_t x) { return (x + 3) & ~(size_t)3; }

static bool any_nonzero(const float* x, size_t n)
{
    for (size_t k = 0; k < n; ++k) if (x[k] != 0.0f) return true;
    return false;
}

template <typename T>
static int dev_alloc(T** out, size_t n)
{
    *out = nullptr;
    if (n == 0) n = 1;
    CU(dev_malloc((void**)out, n * sizeof(T)));
    return RFM_OK;
}

// scope-owned device buffer: freed on every exit path
template <typename T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) dev_free(p); }
    int alloc(size_t n) { return dev_alloc(&p, n); }
    operator T*() const { return p; }
};

static int upload_weights(rfm_session* s, const float* w_i, const float* w_if, const float* v_u, const float* v_i,
                          const float* v_uf, const float* v_if, const float* x_uf, const float* x_if)
{
    const Tables& T = s->T;
    DevBuf<float> st_vu, st_vi, st_wi, st_g;
    int rc;
    if ((rc = st_vu.alloc((size_t)T.Un * T.F))) return rc;
    if ((rc = st_vi.alloc((size_t)T.I * T.F))) return rc;
    if ((rc = st_wi.alloc((size_t)T.I))) return rc;
    CU(cudaMemcpyAsync(st_vu, v_u + (size_t)T.u0 * T.F, (size_t)T.Un * T.F * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(st_vi, v_i, (size_t)T.I * T.F * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(st_wi, w_i, (size_t)T.I * 4, cudaMemcpyHostToDevice, s->st));
    // the (read-only) feature matrices are uploaded once per session and kept, so later weight uploads need no host copy
    if (T.Pp && !s->d_xuf) { if ((rc = dev_alloc(&s->d_xuf, (size_t)T.Un * T.P))) return rc; CU(cudaMemcpyAsync(s->d_xuf, x_uf + (size_t)T.u0 * T.P, (size_t)T.Un * T.P * 4, cudaMemcpyHostToDevice, s->st)); }
    if (T.Qp && !s->d_xif) { if ((rc = dev_alloc(&s->d_xif, (size_t)T.I * T.Q))) return rc; CU(cudaMemcpyAsync(s->d_xif, x_if, (size_t)T.I * T.Q * 4, cudaMemcpyHostToDevice, s->st)); }
    CU(launch_pack_users(T, st_vu, s->d_xuf, s->st));
    CU(launch_pack_items(T, st_vi, st_wi, s->d_xif, s->st));
    s->launches += 2;
    // globals
    const size_t n_wif = (size_t)T.Q, n_vuf = (size_t)T.P * T.F, n_vif = (size_t)T.Q * T.F;
    if ((rc = st_g.alloc(n_wif + n_vuf + n_vif))) return rc;
    CU(cudaMemcpyAsync(st_g, w_if, n_wif * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(st_g + n_wif, v_uf, n_vuf * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(st_g + n_wif + n_vuf, v_if, n_vif * 4, cudaMemcpyHostToDevice, s->st));
    CU(launch_pack_globals(T, st_g, st_g + n_wif, st_g + n_wif + n_vuf, (int)s->gp_floats, s->st));
    s->launches += 1;
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

// Streams and events are pooled per device like the device blocks: a stateless `_fit` call would otherwise create and
// destroy one stream and 4 x epochs events (160+ driver calls at 20 epochs), each of which takes the driver's locks and
// is exposed to whatever else the host is doing -- measured as 10-70 ms spikes of the call on busy hosts.
namespace {
struct StreamPool {
    std::mutex mu;
    std::vector<std::pair<int, cudaStream_t>> streams;
    std::vector<std::pair<int, cudaEvent_t>> events;
};
StreamPool g_pool;
}  // namespace

static cudaError_t pool_stream(int dev, cudaStream_t* out)
{
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        for (size_t k = 0; k < g_pool.streams.size(); ++k)
            if (g_pool.streams[k].first == dev) { *out = g_pool.streams[k].second; g_pool.streams.erase(g_pool.streams.begin() + (long)k); return cudaSuccess; }
    }
    // one step above the default (= lowest) priority: helper streams created with the default priority (the threshold select
    // of the recommend pipeline) then only fill what the kernels of a session's main stream leave free
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { least = 0; greatest = 0; }
    return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, greatest < least ? least - 1 : least);
}
static cudaError_t pool_event(int dev, cudaEvent_t* out)
{
    {
        std::lock_guard<std::mutex> lock(g_pool.mu);
        for (size_t k = g_pool.events.size(); k-- > 0;)
            if (g_pool.events[k].first == dev) { *out = g_pool.events[k].second; g_pool.events.erase(g_pool.events.begin() + (long)k); return cudaSuccess; }
    }
    return cudaEventCreate(out);
}
static void pool_return(int dev, cudaStream_t st, std::vector<cudaEvent_t>& ev)
{
    std::lock_guard<std::mutex> lock(g_pool.mu);
    if (st) g_pool.streams.push_back({dev, st});
    for (auto e : ev) g_pool.events.push_back({dev, e});
    ev.clear();
}

// 32-bit pivot arithmetic of group_member (rfm_common.cuh): G * degree < 2^32 with G <= 32
constexpr int64_t kMaxUserDegree = ((int64_t)1 << 32) / 32;

// (re)seed the device MT19937 state on the host exactly like init_genrand (mt19937ar.c:60-73); pos = N -> first draw twists
static int seed_mt(rfm_session* s)
{
    MtState h;
    uint32_t prev = s->p.mt_seed;
    h.s[0] = prev;
    for (int k = 1; k < kMtN; ++k) { prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)k; h.s[k] = prev; }
    h.pos = kMtN;
    if (!s->d_mt) { int rc = dev_alloc(&s->d_mt, 1); if (rc) return rc; }
    CU(cudaMemcpy(s->d_mt, &h, sizeof h, cudaMemcpyHostToDevice));
    return RFM_OK;
}

extern "C" int rfm_session_destroy(rfm_session* s)
{
    if (!s) return RFM_OK;
    cudaSetDevice(s->device);
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->comm) {                     // the communicator (and its peer window) stays cached in the library for the next session
        rfmh::comm_window_detach(s->comm, s);
        rfmh::comm_release(s->comm);
    }
    dev_free(s->ut_alloc);
    if (!s->p2p) { dev_free(s->T.IT); dev_free(s->T.GP); dev_free(s->d_item_touch); }
    dev_free(s->d_inter); dev_free(s->d_sw); dev_free(s->indptr_alloc); dev_free(s->indices_alloc);
    dev_free(s->bitmap_alloc); dev_free(s->bloom_alloc); dev_free(s->d_perm); dev_free(s->d_mult); dev_free(s->d_mt); dev_free(s->d_acc);
    dev_free(s->d_it_snap); dev_free(s->d_gp_snap); dev_free(s->d_red); dev_free(s->d_flush); dev_free(s->d_gp_acc); dev_free(s->d_xuf); dev_free(s->d_xif);
    dev_free(s->d_snap_ut); dev_free(s->d_snap_it); dev_free(s->d_snap_gp); dev_free(s->d_trace);
    dev_free(s->d_gemm_B); dev_free(s->d_gemm_bias); dev_free(s->d_gemm_order);
    for (void* q : s->scratch) dev_free(q);
    if (s->t0) s->ev.push_back(s->t0);
    if (s->t1) s->ev.push_back(s->t1);
    if (s->st_side) cudaStreamDestroy(s->st_side);
    pool_return(s->device, s->st, s->ev);               // the stream is idle: every dev_free above synchronised the device
    delete s;
    return RFM_OK;
}

extern "C" int rfm_session_create(const rfm_problem* p, rfm_session** out)
{
    if (!p || !out) return fail(RFM_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (p->U <= 0 || p->I <= 0 || p->F <= 0 || p->P <= 0 || p->Q <= 0) return fail(RFM_ERR_ARG, "U, I, P, Q, F must be positive");
    if (!p->w_i || !p->w_if || !p->v_u || !p->v_i || !p->v_uf || !p->v_if || !p->x_uf || !p->x_if) return fail(RFM_ERR_ARG, "weight / feature pointer is NULL");
    if (p->n_interactions < 0) return fail(RFM_ERR_ARG, "n_interactions < 0");
    if (p->n_interactions > 0 && (!p->interactions || !p->sample_weight || !p->csr_indptr || !p->csr_indices))
        return fail(RFM_ERR_ARG, "interaction data pointer is NULL");
    if (p->world < 1 || p->rank < 0 || p->rank >= p->world) return fail(RFM_ERR_ARG, "bad rank/world %d/%d", p->rank, p->world);
    if (p->sampler == RFM_SAMPLER_MT && p->sched != RFM_SCHED_SERIAL) return fail(RFM_ERR_ARG, "the MT19937 sampler needs the serial schedule");
    if (p->max_samples < 1) return fail(RFM_ERR_ARG, "max_samples must be >= 1");
    if (p->epoch_offset < 0) return fail(RFM_ERR_ARG, "epoch_offset < 0");
    const bool ranged = p->user_lo != 0 || p->user_hi != 0;
    if (ranged && (p->user_lo < 0 || p->user_hi <= p->user_lo || p->user_hi > p->U))
        return fail(RFM_ERR_ARG, "bad user range [%d, %d) for %d users", p->user_lo, p->user_hi, p->U);
    const int ndev = rfm_device_count();
    if (ndev == 0) return fail(RFM_ERR_NO_DEVICE, "no CUDA device: rankfm_b200 has no CPU fallback");
    if (p->device < 0 || p->device >= ndev) return fail(RFM_ERR_ARG, "device %d out of range (%d devices)", p->device, ndev);

    rfm_session* s = new rfm_session();
    s->p = *p;
    s->device = p->device;
    s->epochs_done = p->epoch_offset;
    int rc = RFM_OK;
    auto bail = [&](int code) { rfm_session_destroy(s); return code; };
#define TRY(x) do { rc = (x); if (rc) return bail(rc); } while (0)
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(RFM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    CUB(cudaSetDevice(s->device));
    {
        static int sm_count_cache[64] = {0};                       // cudaGetDeviceProperties costs ~1 ms per call
        if (s->device < 64 && sm_count_cache[s->device] > 0) s->n_sm = sm_count_cache[s->device];
        else {
            int n = 0;
            CUB(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, s->device));
            s->n_sm = n;
            if (s->device < 64) sm_count_cache[s->device] = n;
        }
    }
    CUB(pool_stream(s->device, &s->st));

    Tables& T = s->T;
    T.U = p->U; T.I = p->I; T.F = p->F; T.P = p->P; T.Q = p->Q;
    T.u0 = ranged ? p->user_lo : 0;
    T.Un = ranged ? p->user_hi - p->user_lo : p->U;
    // the flags decide the table layout and which kernel variant runs: every rank of a multi-GPU job must reach the same
    // answer, so they are computed over ALL users / items like the reference does (_rankfm.pyx:193-194)
    T.x_uf_any = any_nonzero(p->x_uf, (size_t)p->U * p->P) ? 1 : 0;
    T.x_if_any = any_nonzero(p->x_if, (size_t)p->I * p->Q) ? 1 : 0;
    T.Fp = (int)round4(p->F); T.NQ = T.Fp / 4;
    T.Pp = T.x_uf_any ? (int)round4(p->P) : 0;
    T.Qp = T.x_if_any ? (int)round4(p->Q) : 0;
    T.ldu = T.Fp + T.Pp;
    T.ldi = T.Fp + 4 + T.Qp;
    const size_t wif_p = round4(p->Q);
    T.gp_vuf = (int)wif_p;
    T.gp_vif = T.gp_vuf + p->P * T.Fp;
    s->gp_floats = (size_t)T.gp_vif + (size_t)p->Q * T.Fp;
    {
        int qpl = 1;
        const int G = train_group_size(T, &qpl);
        if (std::max(T.Pp, T.Qp) > 4 * G || qpl > 4)
            return bail(fail(RFM_ERR_UNSUPPORTED, "unsupported shape: factors=%d (max 512), active user/item features %d/%d (max 128)", p->F, T.Pp ? p->P : 0, T.Qp ? p->Q : 0));
    }
    if (p->world > 1) {
        // communicator of this job: created by the first session, kept by the library for the following ones
        TRY(rfmh::comm_acquire(p->nccl_id, p->rank, p->world, p->device, &s->comm));
        const rfmh::ExchangeShape shape{T.I, T.ldi, T.NQ, T.Fp, s->gp_floats};
        TRY(rfmh::comm_window_attach(s->comm, shape, s->st, s, &T.IT, &T.GP, &s->d_item_touch, &s->p2p));
    }
    TRY(dev_alloc(&s->ut_alloc, (size_t)T.Un * T.ldu));
    T.UT = s->ut_alloc - (ptrdiff_t)T.u0 * T.ldu;                   // virtual base: T.UT + u*ldu is the row of GLOBAL user u
    if (!s->p2p) {
        TRY(dev_alloc(&T.IT, (size_t)T.I * T.ldi));
        TRY(dev_alloc(&T.GP, s->gp_floats));
    }
    TRY(upload_weights(s, p->w_i, p->w_if, p->v_u, p->v_i, p->v_uf, p->v_if, p->x_uf, p->x_if));

    s->N = p->n_interactions;
    if (s->N > 0) {
        // user_items of the owned users only: indptr[u0 .. u0+Un] and the slice of indices they point into
        const int64_t nz0 = p->csr_indptr[T.u0], nz1 = p->csr_indptr[T.u0 + T.Un];
        s->nnz = nz1 - nz0;
        TRY(dev_alloc(&s->d_inter, (size_t)s->N));
        TRY(dev_alloc(&s->d_sw, (size_t)s->N));
        TRY(dev_alloc(&s->indptr_alloc, (size_t)T.Un + 1));
        TRY(dev_alloc(&s->indices_alloc, (size_t)s->nnz));
        s->d_indptr = s->indptr_alloc - T.u0;
        s->d_indices = s->indices_alloc - nz0;
        CUB(cudaMemcpyAsync(s->d_inter, p->interactions, (size_t)s->N * 8, cudaMemcpyHostToDevice, s->st));
        CUB(cudaMemcpyAsync(s->d_sw, p->sample_weight, (size_t)s->N * 4, cudaMemcpyHostToDevice, s->st));
        CUB(cudaMemcpyAsync(s->indptr_alloc, p->csr_indptr + T.u0, ((size_t)T.Un + 1) * 8, cudaMemcpyHostToDevice, s->st));
        CUB(cudaMemcpyAsync(s->indices_alloc, p->csr_indices + nz0, (size_t)s->nnz * 4, cudaMemcpyHostToDevice, s->st));
        s->h_indptr.assign(p->csr_indptr, p->csr_indptr + p->U + 1);
        // WARP multiplier by number of draws: log((I-1)//sampled)/log(I)  (_rankfm.pyx:269, integer quotient)
        std::vector<float> mult((size_t)p->max_samples + 1, 0.f);
        for (int k = 1; k <= p->max_samples; ++k)
            mult[k] = (float)(std::log((double)((long)(p->I - 1) / (long)k)) / std::log((double)p->I));
        TRY(dev_alloc(&s->d_mult, mult.size()));
        CUB(cudaMemcpyAsync(s->d_mult, mult.data(), mult.size() * 4, cudaMemcpyHostToDevice, s->st));
        if (p->sampler == RFM_SAMPLER_MT) TRY(seed_mt(s));
        if (p->order == RFM_ORDER_HOST) TRY(dev_alloc(&s->d_perm, (size_t)s->N));
        {
            // the (G+1)-ary membership search does 32-bit pivot arithmetic on G * (a user's degree), G <= 32
            int64_t max_deg = 0;
            for (int u = T.u0; u < T.u0 + T.Un; ++u) max_deg = std::max(max_deg, p->csr_indptr[u + 1] - p->csr_indptr[u]);
            if (max_deg >= kMaxUserDegree)
                return bail(fail(RFM_ERR_UNSUPPORTED, "a user with %lld observed items: more than %lld per user are not supported", (long long)max_deg, (long long)kMaxUserDegree - 1));
        }
        {
            // membership bitmap when (owned users) x I bits fit the budget: default min(8 GiB, a quarter of the free HBM);
            // RANKFM_B200_BITMAP_MB overrides, 0 disables (the kernels then search the CSR)
            const char* env = getenv("RANKFM_B200_BITMAP_MB");
            const int words = (p->I + 31) / 32;
            const double need_mb = (double)T.Un * words * 4.0 / (1024.0 * 1024.0);
            size_t free_b = (size_t)64 << 30, total_b = 0;
            if (!env && need_mb > 64.0) cudaMemGetInfo(&free_b, &total_b);       // small bitmaps always fit: skip the driver call
            const double budget_mb = env ? atof(env) : std::min(8192.0, (double)free_b / (4.0 * 1024.0 * 1024.0));
            if (need_mb <= budget_mb) {
                s->bitmap_words = words;
                TRY(dev_alloc(&s->bitmap_alloc, (size_t)T.Un * words));
                s->d_bitmap = s->bitmap_alloc - (ptrdiff_t)T.u0 * words;
                CUB(cudaMemsetAsync(s->bitmap_alloc, 0, (size_t)T.Un * words * 4, s->st));
                CUB(launch_build_bitmap(s->indptr_alloc, s->d_indices, T.Un, s->bitmap_alloc, words, s->st));
                s->launches += 1;
            } else if (s->nnz > 0 && !(getenv("RANKFM_B200_BLOOM") && !strcmp(getenv("RANKFM_B200_BLOOM"), "0"))) {
                // catalogue too large for a bitmap: a filter word per CSR entry lets the sampler dismiss ~97 % of the candidates
                // with one load instead of a search of the user's item list (exact: set bits are verified by the search)
                TRY(dev_alloc(&s->bloom_alloc, (size_t)s->nnz));
                s->d_bloom = s->bloom_alloc - nz0;
                CUB(cudaMemsetAsync(s->bloom_alloc, 0, (size_t)s->nnz * 4, s->st));
                CUB(launch_build_bloom(s->indptr_alloc, s->d_indices, T.Un, s->d_bloom, s->st));
                s->launches += 1;
            }
        }
        CUB(cudaStreamSynchronize(s->st));
    }
    if (s->N > 0 && (T.x_uf_any || T.x_if_any)) {
        TRY(dev_alloc(&s->d_gp_acc, s->gp_floats));
        CUB(cudaMemsetAsync(s->d_gp_acc, 0, s->gp_floats * 4, s->st));
        if (p->sched == RFM_SCHED_PARALLEL && sgd_pipe_smem_bytes(T) > 200 * 1024)
            return bail(fail(RFM_ERR_UNSUPPORTED, "side-feature blocks too large for the production schedule (warp-private copies need %zu KiB of shared memory per block); use fewer feature columns/factors or the replay mode",
                             sgd_pipe_smem_bytes(T) / 1024));
    }
    {
        TrainParams tp{};
        tp.T = T;
        tp.max_samples = p->max_samples;
        int per_sm = std::max(1, sgd_epoch_blocks_per_sm(tp));
        if (const char* e = getenv("RANKFM_B200_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // experiments
        s->grid = s->n_sm * per_sm;
    }
    if (s->comm) {
        TRY(dev_alloc(&s->d_it_snap, (size_t)T.I * T.ldi));
        TRY(dev_alloc(&s->d_gp_snap, s->gp_floats));
        if (!s->p2p) TRY(dev_alloc(&s->d_item_touch, (size_t)T.I));
        CUB(cudaMemsetAsync(s->d_item_touch, 0, (size_t)T.I * 4, s->st));
        if (s->N > 0) { item_histogram_kernel<<<s->n_sm * 4, 256, 0, s->st>>>(s->d_inter, s->N, s->d_item_touch); s->launches += 1; }
        CUB(cudaStreamSynchronize(s->st));
    }
#undef TRY
#undef CUB
    *out = s;
    return RFM_OK;
}

extern "C" int rfm_session_set_weights(rfm_session* s, const float* w_i, const float* w_if, const float* v_u, const float* v_i,
                                       const float* v_uf, const float* v_if)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    CU(cudaSetDevice(s->device));
    int rc = upload_weights(s, w_i, w_if, v_u, v_i, v_uf, v_if, nullptr, nullptr);   // features stay as uploaded at creation
    if (rc) return rc;
    s->gemm_valid = false;
    // the Philox / Feistel keys keep counting over the life of the session (`epochs_done`): new weights are a warm start
    // (`fit_partial`), not a replay; the MT19937 stream restarts like the reference re-seeds it in every `_fit` (:182)
    if (s->d_mt) { rc = seed_mt(s); if (rc) return rc; }
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_snapshot(rfm_session* s)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    CU(cudaSetDevice(s->device));
    const size_t nu = (size_t)s->T.Un * s->T.ldu, ni = (size_t)s->T.I * s->T.ldi;
    int rc;
    if (!s->d_snap_ut) {
        if ((rc = dev_alloc(&s->d_snap_ut, nu))) return rc;
        if ((rc = dev_alloc(&s->d_snap_it, ni))) return rc;
        if ((rc = dev_alloc(&s->d_snap_gp, s->gp_floats))) return rc;
    }
    CU(cudaMemcpyAsync(s->d_snap_ut, s->ut_alloc, nu * 4, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_snap_it, s->T.IT, ni * 4, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_snap_gp, s->T.GP, s->gp_floats * 4, cudaMemcpyDeviceToDevice, s->st));
    s->snap_epochs = s->epochs_done;
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_restore(rfm_session* s)
{
    if (!s || !s->d_snap_ut) return fail(RFM_ERR_ARG, "no snapshot to restore");
    CU(cudaSetDevice(s->device));
    const size_t nu = (size_t)s->T.Un * s->T.ldu, ni = (size_t)s->T.I * s->T.ldi;
    CU(cudaMemcpyAsync(s->ut_alloc, s->d_snap_ut, nu * 4, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaMemcpyAsync(s->T.IT, s->d_snap_it, ni * 4, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaMemcpyAsync(s->T.GP, s->d_snap_gp, s->gp_floats * 4, cudaMemcpyDeviceToDevice, s->st));
    s->epochs_done = s->snap_epochs;
    s->gemm_valid = false;
    return RFM_OK;
}

extern "C" int rfm_session_timer_start(rfm_session* s)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    CU(cudaSetDevice(s->device));
    if (!s->t0) { CU(pool_event(s->device, &s->t0)); CU(pool_event(s->device, &s->t1)); }
    CU(cudaStreamSynchronize(s->st));
    CU(cudaEventRecord(s->t0, s->st));
    return RFM_OK;
}

extern "C" int rfm_session_timer_stop(rfm_session* s, float* ms_out)
{
    if (!s || !ms_out || !s->t0) return fail(RFM_ERR_ARG, "timer not started");
    CU(cudaSetDevice(s->device));
    CU(cudaEventRecord(s->t1, s->st));
    CU(cudaEventSynchronize(s->t1));
    CU(cudaEventElapsedTime(ms_out, s->t0, s->t1));
    return RFM_OK;
}

extern "C" int rfm_session_download(rfm_session* s, float* w_i, float* w_if, float* v_u, float* v_i, float* v_uf, float* v_if)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    CU(cudaSetDevice(s->device));
    const Tables& T = s->T;
    DevBuf<float> st_vu, st_vi, st_wi, st_g;
    int rc;
    if ((rc = st_vu.alloc((size_t)T.Un * T.F))) return rc;
    if ((rc = st_vi.alloc((size_t)T.I * T.F))) return rc;
    if ((rc = st_wi.alloc((size_t)T.I))) return rc;
    const size_t n_wif = (size_t)T.Q, n_vuf = (size_t)T.P * T.F, n_vif = (size_t)T.Q * T.F;
    if ((rc = st_g.alloc(n_wif + n_vuf + n_vif))) return rc;
    CU(launch_unpack_users(T, st_vu, s->st));
    CU(launch_unpack_items(T, st_vi, st_wi, s->st));
    CU(launch_unpack_globals(T, st_g, st_g + n_wif, st_g + n_wif + n_vuf, s->st));
    s->launches += 3;
    // only the rows this session owns are written: with a user-partitioned job the other rows belong to other ranks
    CU(cudaMemcpyAsync(v_u + (size_t)T.u0 * T.F, st_vu, (size_t)T.Un * T.F * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(v_i, st_vi, (size_t)T.I * T.F * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(w_i, st_wi, (size_t)T.I * 4, cudaMemcpyDeviceToHost, s->st));
    if (T.x_if_any) CU(cudaMemcpyAsync(w_if, st_g, n_wif * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(v_uf, st_g + n_wif, n_vuf * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(v_if, st_g + n_wif + n_vuf, n_vif * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU: replicated item table / globals, per-epoch sum of deltas (see DESIGN.md, section multi-GPU)
// ---------------------------------------------------------------------------------------------------------------
__global__ void delta_kernel(float* __restrict__ cur, const float* __restrict__ snap, size_t n, float scale)     // cur <- scale*(cur - snap)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) cur[e] = scale * (cur[e] - snap[e]);
}
__global__ void apply_kernel(float* __restrict__ cur, float* __restrict__ snap, size_t n)           // cur <- snap + cur ; snap <- cur
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const float v = snap[e] + cur[e];
        cur[e] = v; snap[e] = v;
    }
}

static int exchange_deltas(rfm_session* s, float* cur, float* snap, size_t n, float scale = 1.0f)
{
    const int grid = s->n_sm * 4;
    delta_kernel<<<grid, 256, 0, s->st>>>(cur, snap, n, scale);
    int rc = rfmh::comm_allreduce_f32(s->comm, cur, n, s->st);
    if (rc) return rc;
    apply_kernel<<<grid, 256, 0, s->st>>>(cur, snap, n);
    s->launches += 2;
    CU(cudaGetLastError());
    return RFM_OK;
}

// Item table: every rank moved its replica of row i as if it were alone.  How the C replica deltas combine depends on how
// far the row travelled towards its own equilibrium within the epoch: rows touched rarely barely moved (deltas add up,
// gain 1), rows touched thousands of times have each forgotten their start (the replicas are C samples of the same
// quasi-stationary value: average, gain 1/C).  Same fold rule as the feature-parameter chains (fold_gain), evaluated per
// row from its touch count n_i = positives in this rank's shard + expected uniform negatives, with the per-touch
// contraction rate eta*(2*alpha + curvature); the logistic curvature is ~0.15 for the bias and ~0.01 for a factor.
__global__ void item_histogram_kernel(const int2* __restrict__ inter, long long n, int32_t* __restrict__ count)
{
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) atomicAdd(count + inter[e].y, 1);
}

__global__ void item_delta_kernel(float* __restrict__ cur, const float* __restrict__ snap, int I, int ldi, int Fp, const int32_t* __restrict__ touch,
                                  const EpochAcc* __restrict__ acc, float lam_factor, float lam_bias, float C)
{
    const float neg_per_item = (float)((double)acc->draws / (double)I);
    const size_t n = (size_t)I * ldi;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / ldi), col = (int)(e % ldi);
        const float touches = (float)touch[row] + neg_per_item;
        const float x = (col == Fp ? lam_bias : lam_factor) * touches;       // -log of the row's own contraction over the epoch
        float g = 1.0f;
        if (x > 1e-4f) { const float d = __expf(-x); g = (1.0f - __expf(-x * C)) / (C * (1.0f - d)); }
        cur[e] = g * (cur[e] - snap[e]);
    }
}

// NCCL fallback of the per-epoch exchange (the default is the fused peer-memory kernel, rfm_comm.cu)
static int exchange_item_deltas_nccl(rfm_session* s, const EpochAcc* acc, float lam_factor, float lam_bias)
{
    const Tables& T = s->T;
    const int grid = s->n_sm * 4;
    const size_t n = (size_t)T.I * T.ldi;
    item_delta_kernel<<<grid, 256, 0, s->st>>>(T.IT, s->d_it_snap, T.I, T.ldi, T.Fp, s->d_item_touch, acc, lam_factor, lam_bias, (float)s->p.world);
    int rc = rfmh::comm_allreduce_f32(s->comm, T.IT, n, s->st);
    if (rc) return rc;
    apply_kernel<<<grid, 256, 0, s->st>>>(T.IT, s->d_it_snap, n);
    s->launches += 2;
    CU(cudaGetLastError());
    return RFM_OK;
}

// Folding C independent chains of a parameter with per-step decay (1-lambda), each run for n steps from the same start:
// theta = start + gain * sum_c (theta_c - start).  gain = (1 - d^C) / (C (1 - d)), d = (1-lambda)^n, is exact for the
// decay part: 1 (sum of deltas) when the chains barely move, 1/C (average) when each chain has forgotten its start.
// logistic curvature per touch assumed by the per-row fold gain of the item table (DESIGN.md "Per-row fold gain")
constexpr float kCurvatureFactor = 0.01f, kCurvatureBias = 0.15f;

static float fold_gain(double lambda, double n, double C)
{
    if (C <= 1.0) return 1.0f;
    const double d = std::pow(std::max(0.0, 1.0 - lambda), n);
    if (1.0 - d < 1e-12) return 1.0f;
    return (float)((1.0 - std::pow(d, C)) / (C * (1.0 - d)));
}

// Read back the records of epochs [e0, e1) of the current rfm_session_train call, fold them over the ranks of a multi-GPU
// job, fill `stats`, run the epoch callback and turn non-finite weights into RFM_ERR_NONFINITE (`assert_finite`,
// _rankfm.pyx:95-103,329) -- `status` keeps the first failure.
static int collect_epochs(rfm_session* s, int e0, int e1, const std::vector<float>& etas, rfm_epoch_stats* stats, int& status)
{
    if (e1 <= e0) return status;
    const rfm_problem& p = s->p;
    const int n = e1 - e0;
    std::vector<EpochAcc> acc((size_t)n);
    CU(cudaMemcpyAsync(acc.data(), s->d_acc + e0, (size_t)n * sizeof(EpochAcc), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (s->comm) {
        // user rows are owned by exactly one rank and stay there (no collective on v_u).  The epoch records
        // are per shard: sum log-likelihood / draws / the v_u part of the penalty over the ranks (one tiny allreduce)
        int rc = rfmh::comm_check(s->comm, s->st);
        if (rc) return rc;
        constexpr int kRed = 5;
        if (s->red_cap < n) {
            dev_free(s->d_red); s->d_red = nullptr;
            if ((rc = dev_alloc(&s->d_red, (size_t)n * kRed))) return rc;
            s->red_cap = n;
        }
        std::vector<double> red((size_t)n * kRed);
        for (int e = 0; e < n; ++e) {
            const EpochAcc& a = acc[(size_t)e];
            double* r = red.data() + (size_t)e * kRed;
            r[0] = a.ll; r[1] = (double)a.draws; r[2] = a.bad ? 1.0 : 0.0; r[3] = a.wstats[2]; r[4] = a.wstats[6 + 2];
        }
        CU(cudaMemcpyAsync(s->d_red, red.data(), red.size() * 8, cudaMemcpyHostToDevice, s->st));
        if ((rc = rfmh::comm_allreduce_f64(s->comm, s->d_red, red.size(), s->st))) return rc;
        CU(cudaMemcpyAsync(red.data(), s->d_red, red.size() * 8, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        for (int e = 0; e < n; ++e) {
            EpochAcc& a = acc[(size_t)e];
            const double* r = red.data() + (size_t)e * kRed;
            a.ll = r[0]; a.draws = (long long)(r[1] + 0.5); a.bad = r[2] > 0.5 ? 1 : 0; a.wstats[2] = r[3]; a.wstats[6 + 2] = r[4];
        }
    }
    for (int e = e0; e < e1; ++e) {
        const EpochAcc& a = acc[(size_t)(e - e0)];
        rfm_epoch_stats st{};
        st.log_likelihood = a.ll;
        st.draws = a.draws;
        st.eta = etas[(size_t)e];
        st.penalty = (double)p.alpha * (a.wstats[6 + 0] + a.wstats[6 + 2] + a.wstats[6 + 3]) + (double)p.beta * (a.wstats[6 + 1] + a.wstats[6 + 4] + a.wstats[6 + 5]);
        for (int k = 0; k < 6; ++k) st.finite[k] = std::isfinite(a.wstats[k]) ? 1 : 0;
        cudaEventElapsedTime(&st.kernel_ms, s->ev[4 * e + 0], s->ev[4 * e + 1]);
        cudaEventElapsedTime(&st.sync_ms, s->ev[4 * e + 2], s->ev[4 * e + 3]);
        if (stats) stats[e] = st;
        if (s->epoch_cb) s->epoch_cb(e, &st, s->epoch_cb_user);
        if (status == RFM_OK) {
            static const char* names[6] = {"item weights [w_i]", "item feature weights [w_if]", "user factors [v_u]", "item factors [v_i]",
                                           "user-feature factors [v_uf]", "item-feature factors [v_if]"};
            for (int k = 0; k < 6; ++k)
                if (!st.finite[k]) { status = fail(RFM_ERR_NONFINITE, "%s are not finite - try decreasing feature/sample_weight magnitudes", names[k]); break; }
            if (status == RFM_OK && a.bad) status = fail(RFM_ERR_NONFINITE, "pairwise utilities are not finite - try decreasing feature/sample_weight magnitudes");
        }
    }
    return status;
}

extern "C" int rfm_session_set_epoch_callback(rfm_session* s, rfm_epoch_callback cb, void* user)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    s->epoch_cb = cb; s->epoch_cb_user = user;
    return RFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// training
// ---------------------------------------------------------------------------------------------------------------
extern "C" int rfm_session_train(rfm_session* s, int32_t epochs, const int32_t* perms, rfm_epoch_stats* stats)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    if (epochs < 1) return fail(RFM_ERR_ARG, "epochs must be >= 1");
    if (s->N <= 0) return fail(RFM_ERR_ARG, "session holds no interactions");
    const rfm_problem& p = s->p;
    if (p.order == RFM_ORDER_HOST && !perms) return fail(RFM_ERR_ARG, "order=HOST needs perms [epochs,N]");
    if (p.schedule != RFM_SCHEDULE_CONSTANT && p.schedule != RFM_SCHEDULE_INVSCALING) return fail(RFM_ERR_ARG, "unknown [learning_schedule]");
    CU(cudaSetDevice(s->device));
    if (s->acc_cap < epochs) {
        dev_free(s->d_acc);
        int rc = dev_alloc(&s->d_acc, (size_t)epochs);
        if (rc) return rc;
        s->acc_cap = epochs;
    }
    CU(cudaMemsetAsync(s->d_acc, 0, (size_t)epochs * sizeof(EpochAcc), s->st));
    while ((int)s->ev.size() < 4 * epochs) { cudaEvent_t e; CU(pool_event(s->device, &e)); s->ev.push_back(e); }
    if (s->comm) {
        CU(cudaMemcpyAsync(s->d_it_snap, s->T.IT, (size_t)s->T.I * s->T.ldi * 4, cudaMemcpyDeviceToDevice, s->st));
        CU(cudaMemcpyAsync(s->d_gp_snap, s->T.GP, s->gp_floats * 4, cudaMemcpyDeviceToDevice, s->st));
    }

    TrainParams tp{};
    tp.T = s->T;
    tp.interactions = s->d_inter; tp.sample_weight = s->d_sw; tp.indptr = s->d_indptr; tp.indices = s->d_indices;
    tp.mult = s->d_mult;
    tp.bitmap = s->d_bitmap; tp.bitmap_words = s->bitmap_words;
    tp.bloom = s->d_bloom;
    tp.N = s->N;
    tp.reg_a = (float)(2.0 * p.alpha);                      // d_reg_a / d_reg_b, _rankfm.pyx:171-172
    tp.reg_b = (float)(2.0 * p.beta);
    tp.max_samples = p.max_samples;
    tp.max_rejects = p.max_rejects > 0 ? p.max_rejects : (p.sampler == RFM_SAMPLER_MT ? 0x7fffffff : 64);
    tp.serial = p.sched == RFM_SCHED_SERIAL ? 1 : 0;
    tp.k0 = (uint32_t)p.seed; tp.k1 = (uint32_t)(p.seed >> 32) ^ (0x9E3779B9u * (uint32_t)p.rank);   // ranks index their shards locally: decorrelate
    tp.mt = p.sampler == RFM_SAMPLER_MT ? s->d_mt : nullptr;
    tp.trace = s->d_trace;
    std::vector<float> etas((size_t)epochs);
    int status = RFM_OK, cb_done = 0;
    const char* spec_s = getenv("RANKFM_B200_SPEC");          // 0 (default) = adaptive, else force 1 / 2 / 4
    const int spec_env = spec_s ? atoi(spec_s) : 0;
    // Hogwild staleness cap: never keep more than 1/8 of an epoch in flight, so that on small inputs the schedule
    // degrades towards sequential SGD instead of one giant stale batch (large inputs always get the full machine).
    // A warp of the pipelined kernel holds one batch of 32 positives.
    int grid = s->grid;
    {
        const long long in_flight_per_block = (long long)(kTrainThreads / 32) * 32;
        long long div = 8;
        if (const char* e = getenv("RANKFM_B200_INFLIGHT_DIV")) div = std::max(1, atoi(e));      // experiments
        const long long cap_blocks = std::max<long long>(1, (s->N / div) / in_flight_per_block);
        grid = (int)std::min<long long>(grid, cap_blocks);
    }

    for (int e = 0; e < epochs; ++e) {
        // the LR schedule restarts with every call, like the reference's per-`_fit` epoch counter (_rankfm.pyx:218-223);
        // the Philox/Feistel keys keep counting (epochs_done) so a warm start does not replay the same randomness
        float eta = p.learning_rate;
        if (p.schedule == RFM_SCHEDULE_INVSCALING) eta = (float)(((double)p.learning_rate) / std::pow((double)(e + 1), (double)p.learning_exponent));
        etas[e] = eta;
        tp.eta = eta;
        tp.epoch_key = (uint32_t)(s->epochs_done + e);                 // epochs_done starts at the problem's epoch_offset
        tp.acc = s->d_acc + e;
        tp.prev_acc = e > 0 ? s->d_acc + (e - 1) : nullptr;
        tp.spec = tp.serial ? 4 : spec_env;
        if (p.order == RFM_ORDER_HOST) {
            CU(cudaMemcpyAsync(s->d_perm, perms + (size_t)e * s->N, (size_t)s->N * 4, cudaMemcpyHostToDevice, s->st));
            tp.perm = s->d_perm;
        } else {
            tp.perm = nullptr;
            tp.feistel = make_feistel(s->N, p.seed, s->epochs_done + e);
        }
        const bool feat_parallel = !tp.serial && (s->T.x_uf_any || s->T.x_if_any);
        double chains = 1.0, steps_per_chain = (double)s->N;
        if (feat_parallel) {
            // private feature-parameter chains: one per lane group (or warp) that owns >= 1 batch
            const double n_batches = std::ceil((double)s->N / 32.0);
            chains = std::min((double)grid * (kTrainThreads / 32), n_batches) * sgd_pipe_chains_per_warp(s->T);
            steps_per_chain = (double)s->N / chains / sgd_pipe_groups_per_chain(s->T);   // a racing per-warp chain advances once per warp step
            tp.gp_acc = s->d_gp_acc;
            tp.gp_gain = fold_gain((double)tp.reg_b * eta, steps_per_chain, chains);
        }
        CU(cudaEventRecord(s->ev[4 * e + 0], s->st));
        cudaError_t le = launch_sgd_epoch(tp, grid, s->st);
        if (le != cudaSuccess) return fail(RFM_ERR_CUDA, "sgd_epoch launch failed: %s", cudaGetErrorString(le));
        if (feat_parallel) { CU(launch_gp_apply(s->T.GP, s->d_gp_acc, (int)s->gp_floats, s->st)); s->launches += 1; }
        CU(cudaEventRecord(s->ev[4 * e + 1], s->st));
        s->launches += 1;
        CU(cudaEventRecord(s->ev[4 * e + 2], s->st));
        if (s->comm) {
            // fold the ranks' replicas of the item table (and of the feature parameters: every rank ran its own chains, same rule)
            const bool has_gp = s->T.x_uf_any || s->T.x_if_any;
            const float rank_gain = feat_parallel ? fold_gain((double)tp.reg_b * eta, (double)s->N, (double)p.world) : 1.0f;
            const float lam_factor = eta * (2.0f * p.alpha + kCurvatureFactor), lam_bias = eta * (2.0f * p.alpha + kCurvatureBias);
            int rc;
            if (s->p2p) {
                rc = rfmh::comm_exchange_p2p(s->comm, s->st, s->d_it_snap, has_gp ? s->d_gp_snap : nullptr, s->d_acc + e, lam_factor, lam_bias, rank_gain);
                s->launches += 1;
            } else {
                rc = exchange_item_deltas_nccl(s, s->d_acc + e, lam_factor, lam_bias);
                if (!rc && has_gp) rc = exchange_deltas(s, s->T.GP, s->d_gp_snap, s->gp_floats, rank_gain);
            }
            if (rc) return rc;
        }
        CU(cudaEventRecord(s->ev[4 * e + 3], s->st));
        CU(launch_weight_stats(s->T, (s->d_acc + e)->wstats, s->n_sm * 2, s->st));
        s->launches += 1;
        if (s->epoch_cb) {                      // verbose callers see every epoch as it completes, like the reference's prints (:332-336)
            int rc = collect_epochs(s, e, e + 1, etas, stats, status);
            if (rc != RFM_OK && rc != RFM_ERR_NONFINITE) return rc;
            cb_done = e + 1;
        }
    }
    s->epochs_done += epochs;
    s->gemm_valid = false;
    return collect_epochs(s, cb_done, epochs, etas, stats, status);
}

// ---------------------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------------------
static int predict_dev(rfm_session* s, const float2* d_pairs, int64_t n, float* d_scores)
{
    cudaError_t e = launch_predict(s->T, d_pairs, n, d_scores, s->n_sm * 8, s->st);
    if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "predict launch failed: %s", cudaGetErrorString(e));
    s->launches += 1;
    return RFM_OK;
}

extern "C" int rfm_session_predict(rfm_session* s, const float* pairs, int64_t n, float* scores)
{
    if (!s || (n > 0 && (!pairs || !scores))) return fail(RFM_ERR_ARG, "NULL argument");
    if (n == 0) return RFM_OK;
    CU(cudaSetDevice(s->device));
    DevBuf<float2> d_pairs; DevBuf<float> d_scores;
    int rc;
    if ((rc = d_pairs.alloc((size_t)n))) return rc;
    if ((rc = d_scores.alloc((size_t)n))) return rc;
    CU(cudaMemcpyAsync(d_pairs, pairs, (size_t)n * 8, cudaMemcpyHostToDevice, s->st));
    if ((rc = predict_dev(s, d_pairs, n, d_scores))) return rc;
    CU(cudaMemcpyAsync(scores, d_scores, (size_t)n * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_time_predict(rfm_session* s, const float* pairs, int64_t n, int32_t iters, float* ms_out)
{
    if (!s || !pairs || !ms_out || n <= 0 || iters < 1) return fail(RFM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    DevBuf<float2> d_pairs; DevBuf<float> d_scores;
    int rc;
    if ((rc = d_pairs.alloc((size_t)n))) return rc;
    if ((rc = d_scores.alloc((size_t)n))) return rc;
    CU(cudaMemcpyAsync(d_pairs, pairs, (size_t)n * 8, cudaMemcpyHostToDevice, s->st));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    if ((rc = predict_dev(s, d_pairs, n, d_scores))) return rc;      // warm-up
    CU(cudaEventRecord(a, s->st));
    for (int k = 0; k < iters; ++k) if ((rc = predict_dev(s, d_pairs, n, d_scores))) return rc;
    CU(cudaEventRecord(b, s->st));
    CU(cudaStreamSynchronize(s->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / iters;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return RFM_OK;
}

// users as float32 indexes (NaN = unknown) -> int32 (-1 = unknown)
static void users_to_int(const float* users, int64_t n, std::vector<int32_t>& out)
{
    out.resize((size_t)n);
    for (int64_t k = 0; k < n; ++k) out[(size_t)k] = std::isnan(users[k]) ? -1 : (int32_t)users[k];
}

// grow-only scratch buffers (cudaMalloc/cudaFree of hundreds of MB per call would dominate a recommend() call)
template <typename T>
static int scratch_get(rfm_session* s, int slot, size_t count, T** out)
{
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (s->scratch_bytes[slot] < bytes) {
        dev_free(s->scratch[slot]);
        s->scratch[slot] = nullptr; s->scratch_bytes[slot] = 0;
        CU(dev_malloc(&s->scratch[slot], bytes));
        s->scratch_bytes[slot] = bytes;
    }
    *out = reinterpret_cast<T*>(s->scratch[slot]);
    return RFM_OK;
}

// exact fp32 path: score every item, exact radix select
static int recommend_exact(rfm_session* s, const int32_t* d_users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                           float* d_rec, float* gemm_ms)
{
    const Tables& T = s->T;
    if (n_users <= 0) return RFM_OK;
    // batch users so the score matrix stays under ~8 GB
    const int64_t max_batch = std::max<int64_t>(1, std::min<int64_t>(n_users, ((int64_t)2 << 30) / std::max(1, T.I)));
    float* S = nullptr;
    int rc;
    if ((rc = scratch_get(s, 0, (size_t)max_batch * T.I, &S))) return rc;
    const int chunks = std::max(1, std::min(64, (T.I + 2047) / 2048));
    cudaEvent_t a = nullptr, b = nullptr;
    float acc_ms = 0.f;
    if (gemm_ms) { CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b)); }
    for (int64_t off = 0; off < n_users; off += max_batch) {
        const int nb = (int)std::min<int64_t>(max_batch, n_users - off);
        if (gemm_ms) CU(cudaEventRecord(a, s->st));
        for (int y0 = 0; y0 < nb; y0 += 32768) {   // gridDim.y limit is 65535
            const int ny = std::min(32768, nb - y0);
            cudaError_t e = launch_score_users(T, d_users + off + y0, ny, S + (size_t)y0 * T.I, chunks, s->st);
            if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "score_users launch failed: %s", cudaGetErrorString(e));
            s->launches += 1;
        }
        if (gemm_ms) { CU(cudaEventRecord(b, s->st)); }
        cudaError_t e = launch_topn_select(S, T.I, d_users + off, nb, s->d_indptr, s->d_indices, filter_previous, n_items,
                                           d_rec + (size_t)off * n_items, nullptr, s->st);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "topn_select launch failed: %s", cudaGetErrorString(e));
        s->launches += 1;
        if (gemm_ms) { CU(cudaStreamSynchronize(s->st)); float ms = 0.f; cudaEventElapsedTime(&ms, a, b); acc_ms += ms; }
    }
    CU(cudaStreamSynchronize(s->st));
    if (gemm_ms) { *gemm_ms += acc_ms; cudaEventDestroy(a); cudaEventDestroy(b); }
    return RFM_OK;
}

// Shortlist tiers of the tensor-core path.  A user's shortlist n' = 2 n_items + 16 (+ the items it has seen when filtering)
// must fit the candidate buffers: users up to kCandCap go through the narrow tier (small buffers, 512-entry shortlist
// kernel), users with long histories up to kCandCapWide through the wide tier (4x the buffers, 2048-entry shortlist) --
// round 1 sent everybody above 256 to the exact fp32 path, 150x slower per user.
constexpr int kCandCap = 256, kCandCapWide = 1024;

static int ensure_gemm_items(rfm_session* s)
{
    const Tables& T = s->T;
    const int Kp = gemm_kp(T), BN = gemm_block_n(T);
    const int I_pad = (T.I + BN - 1) / BN * BN;
    if (s->gemm_valid && s->gemm_I_pad == I_pad) return RFM_OK;
    if (!s->d_gemm_B || s->gemm_I_pad != I_pad) {
        dev_free(s->d_gemm_B); dev_free(s->d_gemm_bias); dev_free(s->d_gemm_order);
        s->d_gemm_B = nullptr; s->d_gemm_bias = nullptr; s->d_gemm_order = nullptr;
        CU(dev_malloc(&s->d_gemm_B, (size_t)I_pad * Kp * 2));
        int rc = dev_alloc(&s->d_gemm_bias, (size_t)I_pad + 4);          // [I_pad]: largest operand-row norm (shortlist guard)
        if (rc) return rc;
        if ((rc = dev_alloc(&s->d_gemm_order, (size_t)I_pad))) return rc;
        s->gemm_I_pad = I_pad;
    }
    CU(launch_pack_gemm_items(T, Kp, I_pad, s->d_gemm_B, s->d_gemm_bias, s->d_gemm_order, s->st));
    s->launches += 3;
    s->gemm_valid = true;
    s->tau_spec_off = false; s->tau_spec_ok = false;
    return RFM_OK;
}

static int shortlist_target(rfm_session* s, int32_t u, int32_t n_items, int32_t filter_previous)
{
    int need = 2 * n_items + 16;                                   // 2x absorbs the bf16 rounding of the candidate scores
    if (filter_previous && u >= 0) need += (int)(s->h_indptr[(size_t)u + 1] - s->h_indptr[(size_t)u]);
    return need;
}

// Pass 1 (block bounds -> per-row threshold) visits only 1/k of the item tiles.  Three ways to choose them and turn the
// bounds into tau (tau_mode); all of them end in the same verified shortlist, so results never depend on the choice:
//
//  0  CONSERVATIVE   tau = the n'-th largest block bound of the subset: a PROVABLE lower bound of the row's n'-th best
//      score, just a looser one than the whole catalogue would give.  Subset: the catalogue is in descending bias order,
//      so the FIRST 1/k of the tiles hold the items with the largest biases (RANKFM_B200_TAU_SUBSET=stride takes every
//      k-th tile instead); k is the largest of {8, 4, 2, 1} that leaves the subset >= 256 n' items, so that even a
//      catalogue whose biases say nothing about the ranking yields at most ~k n' candidates per row.
//  2  HEAD (default)  the same provable threshold from a much smaller head subset: the largest k <= 32 that leaves
//      4 block bounds per wanted candidate.  Pass 2 is bound by the TMEM read of the accumulators (see rfm_gemm.cu), so
//      pass 1 + threshold are pure overhead: k = 4 -> 32 takes them from 28 % to 5 % of the GEMM + filter time at 262 k
//      items.  Where item biases carry part of the ranking (any popularity signal) the head's n'-th best bound is as
//      tight as the whole catalogue's; where they carry none, the threshold is loose, rows collect ~k n' candidates and
//      overflow their slots (flag 1).  That is detected, not assumed away: flagged rows are served again in mode 0, and
//      if more than 1/32 of the first batch's rows are flagged the rest of the call -- and the session, until the weights
//      change -- runs in mode 0.  One verified batch per weight state switches the check (a stream sync) off.
//  1  ESTIMATED (RANKFM_B200_TAU_MODE=estimate)   every k-th tile (bias order makes that a stratified sample) and tau =
//      the m-th largest bound with m ~ n'/k plus z-sigma head room (tau_rank in rfm_kernels.h): an ESTIMATE of the n'-th
//      best score of the whole catalogue, ~2 n' candidates even without any bias signal.  It can come out too high:
//      shortlist_kernel checks the one thing that matters (did n' candidates reach tau?) and flags the row (2) otherwise,
//      with the same second serving and switch-off as mode 2.  Measured no better than mode 2 on catalogues with a bias
//      signal (the sample sees fewer head items than the head subset), hence opt-in.
//  RANKFM_B200_TAU_MODE=head|safe|estimate, RANKFM_B200_TAU_Z (default 4.5), RANKFM_B200_TAU_STRIDE caps k.
static int tau_mode_default()
{
    const char* e = getenv("RANKFM_B200_TAU_MODE");
    if (e && !strcmp(e, "estimate")) return 1;
    if (e && !strcmp(e, "safe")) return 0;
    return 2;
}
static float tau_z()
{
    const char* e = getenv("RANKFM_B200_TAU_Z");
    const float z = e ? (float)atof(e) : 4.5f;
    return z > 0.f ? z : 1e-3f;
}
// block bounds (one per kTauBlock items) pass 1 produces per row
static int tau_blocks(const Tables& T, int stride)
{
    const int BN = gemm_block_n(T), n_tiles = (T.I + BN - 1) / BN;
    return (n_tiles + stride - 1) / stride * (BN / kTauBlock);
}
static int tau_stride(const Tables& T, int32_t n_items)
{
    const char* e = getenv("RANKFM_B200_TAU_STRIDE");
    const int cap = e ? std::max(1, atoi(e)) : 8;
    const int BN = gemm_block_n(T), n_tiles = (T.I + BN - 1) / BN, want = 2 * n_items + 16;
    const char* m = getenv("RANKFM_B200_TAU_MIN");                   // experiments: items the subset keeps per wanted candidate
    const int per = m ? std::max(16, atoi(m)) : 256;
    int best = 1;
    for (int k = 2; k <= std::min(cap, 16); k *= 2)
        if ((int64_t)((n_tiles + k - 1) / k) * BN >= (int64_t)per * want) best = k;          // the subset keeps >= 256 n' items
    return best;
}
// estimated threshold: the m-th largest bound must sit in the sparse upper tail of the sample (<= 1/32 of its blocks),
// where block maxima and item scores rank alike; `want_max` = the largest n' of the candidate tier
static int tau_stride_estimate(const Tables& T, int want_max, float z)
{
    const char* e = getenv("RANKFM_B200_TAU_STRIDE");
    const int cap = e ? std::max(1, atoi(e)) : 8;
    const char* t = getenv("RANKFM_B200_TAU_TAIL");                  // tests: smaller catalogues through the estimate
    const int tail = t ? std::max(2, atoi(t)) : 32;
    int best = 1;
    for (int k = 2; k <= std::min(cap, 16); k *= 2)
        if ((int64_t)tau_blocks(T, k) >= (int64_t)tail * tau_rank(want_max, k, z)) best = k;
    return best;
}

// mode 2: the smallest head subset that still has 4 block bounds per wanted candidate of the tier
static int tau_stride_head(const Tables& T, int want_max)
{
    const char* e = getenv("RANKFM_B200_TAU_STRIDE");
    const int cap = e ? std::max(1, atoi(e)) : 32;
    int best = 1;
    for (int k = 2; k <= std::min(cap, 32); k *= 2)
        if ((int64_t)tau_blocks(T, k) >= (int64_t)4 * want_max) best = k;
    return best;
}

static bool tau_subset_head()
{
    const char* e = getenv("RANKFM_B200_TAU_SUBSET");
    return !(e && !strcmp(e, "stride"));
}

// rows [0, n) of `src` -> rows `rows[k]` of d_rec
static int scatter_rows(rfm_session* s, float* d_rec, const float* src, const std::vector<int64_t>& rows, int32_t n_items)
{
    for (size_t k = 0; k < rows.size(); ++k)
        CU(cudaMemcpyAsync(d_rec + (size_t)rows[k] * n_items, src + k * n_items, (size_t)n_items * 4, cudaMemcpyDeviceToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

// Where a top-level recommend_tc call may put finished rows while later batches still compute: the caller's host buffer,
// filled batch by batch on a copy stream (a 1 M x 100 result is 400 MB: copied after the last kernel it cost 110 ms on
// top of 290 ms of GPU work).  Rows rewritten afterwards (second servings, exact path) are listed in `dirty`.
struct HostSink {
    float* out = nullptr;               // [n_users, n_items], rows in the order of the call's device rows
    cudaStream_t cs = nullptr;
    std::vector<int64_t> dirty;
    bool whole = false;                 // nothing was copied row-exactly: the caller copies everything
};

// tensor-core path (rfm_gemm.cu): pass 1 block bounds -> per-row threshold -> pass 2 candidates -> shortlist (n' best by
// bf16 score, exact fp32 re-score) -> top-n.  `top_level`: the caller's rows (counted in tc_rows), not a redo.
static int recommend_tc(rfm_session* s, const int32_t* d_users, const int32_t* h_users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                        float* d_rec, float* gemm_ms, int cand_cap, int tau_mode, bool top_level = true, HostSink* sink = nullptr)
{
    const Tables& T = s->T;
    if (n_users <= 0) return RFM_OK;
    int rc = ensure_gemm_items(s);
    if (rc) return rc;
    const float z = tau_z();
    const int Kp = gemm_kp(T), MT = gemm_m_tile(T), SPS = gemm_slots_per_split(T), I_pad = s->gemm_I_pad;
    int stride = tau_stride(T, n_items);
    if (tau_mode == 1) {
        const int k = tau_stride_estimate(T, cand_cap, z);
        if (k > 1) stride = k; else tau_mode = 0;                    // nothing to estimate from a full pass 1
    } else if (tau_mode == 2) {
        const int k = tau_stride_head(T, cand_cap);
        if (k > stride) stride = k; else tau_mode = 0;               // the conservative subset is already that small
    }
    const int n_sub1 = tau_blocks(T, stride), n_tiles1 = n_sub1 / (gemm_block_n(T) / kTauBlock);
    // candidate entries per row (the shortlist kernel stages a row's candidates in shared memory: at most 16384 entries =
    // 128 KB): provable threshold ~1-2 n' with a full pass 1, ~stride x n' with a partial one; estimated threshold ~k m
    // expected, 6x head room (bias-dominated rows collect more: one bound per 8-item block hides up to 7 items)
    int width = std::min(16384, stride == 1 ? 8 * cand_cap : 4 * cand_cap * stride);
    if (tau_mode) {                                                  // speculative modes: sized for what they expect, not for the worst case
        const int expect = tau_mode == 1 ? 6 * stride * tau_rank(cand_cap, stride, z) : 12 * cand_cap;
        int w = 2048;
        while (w < 16384 && w < expect) w *= 2;
        width = w;
    }
    // The shortlist kernel stages a row's candidates in shared memory; sized for every slot's capacity that is 32 KB per block at
    // width 4096 and caps it at 5 blocks per SM.  The speculative modes stage only what they expect -- 2,048 entries in the
    // narrow tier, 16 KB, so that the kernel's 6 blocks per SM fit (rows with more take the second serving like an
    // overflowing slot).
    const char* stg = getenv("RANKFM_B200_TC_STAGE");
    const int stage_cap = tau_mode ? (stg ? std::max(256, atoi(stg)) : std::min(width, std::max(2048, 4 * cand_cap))) : 0;
    // one wave: at most n_sm CTAs (one resident per SM), user tiles x item splits; >= 2 splits keep a partial last batch balanced
    int64_t max_rows = (int64_t)std::max(1, s->n_sm / 2) * MT;
    max_rows = std::min<int64_t>(max_rows, std::max<int64_t>(MT, (((int64_t)4 << 30) / ((int64_t)n_sub1 * 4)) / MT * MT));
    const int64_t rows_alloc = std::min<int64_t>(max_rows, (n_users + MT - 1) / MT * MT);
    const int split_cap = std::max(1, std::min(n_tiles1, width / (cand_cap * SPS)));
    // The user batches run back to back on the session's stream; targets of all batches are uploaded once and the redo
    // flags of all rows are read once, so the loop never synchronises with the host (except once after the first batch of
    // a speculative-threshold call, to see whether the speculation works on this catalogue).
    // Software pipeline over the batches: the FRONT of batch b+1 (pack the user operand, pass 1, threshold select) is
    // issued before pass 2 of batch b, with the threshold select on a second, lower-priority stream -- it needs ~5 KB of
    // shared memory per block and co-resides with the pass-2 CTAs (which leave ~28 KB per SM and 70 % of the issue slots
    // free), so its 0.054 ms per batch disappear under pass 2.  A, the block bounds and tau are double-buffered.
    // (The same trick does NOT work for the shortlist kernel: its blocks cannot co-reside with a GEMM CTA that holds
    // ~200 KB of an SM's shared memory and only delay the next wave -- measured in round 2: GEMM + filter 6.2 -> 8.3 ms
    // per 65,536 users for the same 9.0 ms total.)
    __nv_bfloat16_raw* d_A = nullptr; int* d_ntgt = nullptr; float2* d_cand = nullptr; int* d_cnt = nullptr;
    float *d_rowmax = nullptr, *d_tau = nullptr; int* d_flag = nullptr;
    const int64_t n_batches = (n_users + rows_alloc - 1) / rows_alloc;
    const bool piped = n_batches > 1 && !(getenv("RANKFM_B200_TC_PIPE") && !strcmp(getenv("RANKFM_B200_TC_PIPE"), "0"));
    const size_t nbuf = piped ? 2 : 1;
    struct Events : std::vector<cudaEvent_t> { ~Events() { for (auto e : *this) cudaEventDestroy(e); } } ev;    // timing pairs (GEMM + filter kernels)
    if ((rc = scratch_get(s, 1, nbuf * rows_alloc * Kp, &d_A))) return rc;
    if ((rc = scratch_get(s, 2, (size_t)n_batches * rows_alloc, &d_ntgt))) return rc;
    if ((rc = scratch_get(s, 3, nbuf * rows_alloc, &d_tau))) return rc;
    if ((rc = scratch_get(s, 4, nbuf * rows_alloc * n_sub1, &d_rowmax))) return rc;
    if ((rc = scratch_get(s, 5, (size_t)rows_alloc * width, &d_cand))) return rc;
    if ((rc = scratch_get(s, 6, (size_t)rows_alloc * split_cap * SPS, &d_cnt))) return rc;
    if ((rc = scratch_get(s, 9, (size_t)n_users, &d_flag))) return rc;
    std::vector<int> ntgt((size_t)(n_batches * rows_alloc));
    for (int64_t bi = 0; bi < n_batches; ++bi)
        for (int64_t r = 0; r < rows_alloc; ++r) {
            const int64_t k = bi * rows_alloc + r;
            ntgt[(size_t)k] = shortlist_target(s, k < n_users ? h_users[k] : -1, n_items, filter_previous);
        }
    CU(cudaMemcpyAsync(d_ntgt, ntgt.data(), ntgt.size() * 4, cudaMemcpyHostToDevice, s->st));
    std::vector<int> flag_h((size_t)n_users);
    struct PooledEvents : std::vector<cudaEvent_t> { int dev; explicit PooledEvents(int d) : dev(d) {} ~PooledEvents() { pool_return(dev, nullptr, *this); } };
    PooledEvents batch_done(s->device), order_ev(s->device);
    std::vector<cudaEvent_t> thr_done((size_t)n_batches, nullptr);
    cudaStream_t side = s->st;
    if (piped) {
        if (!s->st_side) CU(cudaStreamCreateWithFlags(&s->st_side, cudaStreamNonBlocking));      // default (= lowest) priority; pooled streams run above it
        side = s->st_side;
    }
    struct SideSync { cudaStream_t st; ~SideSync() { if (st) cudaStreamSynchronize(st); } } side_sync{piped ? side : nullptr};   // nothing of this call outlives it
    struct Geom { int64_t off; int nb, M_pad, n_splits, slots, cap; };
    auto geom = [&](int64_t bi) {
        Geom g;
        g.off = bi * rows_alloc;
        g.nb = (int)std::min<int64_t>(rows_alloc, n_users - g.off);
        g.M_pad = (g.nb + MT - 1) / MT * MT;
        g.n_splits = std::max(1, std::min(split_cap, s->n_sm / (g.M_pad / MT)));
        g.slots = g.n_splits * SPS; g.cap = width / g.slots;
        return g;
    };
    auto mark = [&]() -> int {                                       // opens / closes a timing bracket on the main stream (events pair up in order)
        if (!gemm_ms) return RFM_OK;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return fail(RFM_ERR_CUDA, "cudaEventCreate failed");
        ev.push_back(e);
        CU(cudaEventRecord(e, s->st));
        return RFM_OK;
    };
    // front of a batch: user operand, pass 1 (block bounds), threshold select.  In the pipelined loop pass 1 is bracketed by
    // itself and the select runs on the side stream; otherwise the caller's bracket spans pass 1 .. pass 2.
    auto front = [&](int64_t bi) -> int {
        const Geom g = geom(bi);
        const size_t par = piped ? (size_t)(bi & 1) : 0;
        __nv_bfloat16_raw* A = d_A + par * rows_alloc * Kp;
        float* rowmax = d_rowmax + par * rows_alloc * n_sub1;
        float* tau = d_tau + par * rows_alloc;
        CU(launch_pack_gemm_users(T, d_users + g.off, g.nb, g.M_pad, Kp, A, s->st));
        int r = mark();
        if (r) return r;
        const int subset = (tau_mode == 1 || (tau_mode == 0 && !tau_subset_head())) ? stride : -stride;
        cudaError_t e = launch_score_filter(T, 1, A, s->d_gemm_B, s->d_gemm_bias, g.nb, g.M_pad, I_pad, g.n_splits, subset, nullptr, nullptr, nullptr, 0, rowmax, nullptr, s->st);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "score_filter pass 1 (tcgen05) launch failed: %s", cudaGetErrorString(e));
        if (piped) {
            if ((r = mark())) return r;
            cudaEvent_t p1;
            CU(pool_event(s->device, &p1)); order_ev.push_back(p1);
            CU(cudaEventRecord(p1, s->st));
            CU(cudaStreamWaitEvent(side, p1, 0));
        }
        e = launch_row_threshold(rowmax, g.M_pad, n_sub1, d_ntgt + (size_t)g.off, tau, tau_mode == 1 ? stride : 1, tau_mode == 1 ? z : 0.f, side);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "row_threshold launch failed: %s", cudaGetErrorString(e));
        if (piped) {
            cudaEvent_t td;
            CU(pool_event(s->device, &td)); order_ev.push_back(td);
            CU(cudaEventRecord(td, side));
            thr_done[(size_t)bi] = td;
        }
        return RFM_OK;
    };
    int64_t served = n_users;                                        // rows this invocation's loop handles
    if ((rc = front(0))) return rc;
    for (int64_t bi = 0; bi < n_batches; ++bi) {
        const Geom g = geom(bi);
        const int64_t off = g.off;
        const int nb = g.nb;
        const size_t par = piped ? (size_t)(bi & 1) : 0;
        const __nv_bfloat16_raw* A = d_A + par * rows_alloc * Kp;
        const float* tau = d_tau + par * rows_alloc;
        const int* tgt = d_ntgt + (size_t)off;
        if (piped) {
            if (bi + 1 < n_batches && (rc = front(bi + 1))) return rc;
            if ((rc = mark())) return rc;                       // a select that is not done yet stalls inside this bracket
            CU(cudaStreamWaitEvent(s->st, thr_done[(size_t)bi], 0));
        }
        cudaError_t e = launch_score_filter(T, 2, A, s->d_gemm_B, s->d_gemm_bias, nb, g.M_pad, I_pad, g.n_splits, 1, d_cand, d_cnt, tau, g.cap, nullptr, nullptr, s->st);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "score_filter pass 2 (tcgen05) launch failed: %s", cudaGetErrorString(e));
        if ((rc = mark())) return rc;
        e = launch_shortlist(T, d_users + off, nb, d_cand, d_cnt, g.slots, g.cap, s->d_gemm_bias, s->d_gemm_order, tgt, s->d_indptr, s->d_indices, filter_previous,
                             n_items, d_rec + (size_t)off * n_items, d_flag + off, tau, I_pad, 2 * cand_cap, stage_cap, s->st);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "shortlist launch failed: %s", cudaGetErrorString(e));
        s->launches += 5;
        if (sink) {
            cudaEvent_t done;
            CU(pool_event(s->device, &done));
            batch_done.push_back(done);                              // back to the pool when this call returns
            CU(cudaEventRecord(done, s->st));
        }
        if (!piped && bi + 1 < n_batches && (rc = front(bi + 1))) return rc;
        if (tau_mode && bi == 0 && n_batches > 1 && !s->tau_spec_ok) {      // does the speculation work on this catalogue?
            CU(cudaMemcpyAsync(flag_h.data(), d_flag, (size_t)nb * 4, cudaMemcpyDeviceToHost, s->st));
            CU(cudaStreamSynchronize(s->st));
            int64_t missed = 0;
            for (int r = 0; r < nb; ++r) missed += flag_h[(size_t)r] != 0;
            if (missed * 32 > nb) { s->tau_spec_off = true; served = off + nb; if (piped) cudaStreamSynchronize(side); break; }
            s->tau_spec_ok = true;
        }
    }
    if (sink) {                                                      // every batch is enqueued: drain finished ones while the rest compute
        for (size_t bi = 0; bi < batch_done.size(); ++bi) {
            const int64_t off = (int64_t)bi * rows_alloc;
            const int64_t nb = std::min<int64_t>(rows_alloc, n_users - off);
            CU(cudaStreamWaitEvent(sink->cs, batch_done[bi], 0));
            CU(cudaMemcpyAsync(sink->out + (size_t)off * n_items, d_rec + (size_t)off * n_items, (size_t)nb * n_items * 4, cudaMemcpyDeviceToHost, sink->cs));
        }
        CU(cudaStreamSynchronize(sink->cs));
    }
    CU(cudaStreamSynchronize(s->st));
    if (gemm_ms)
        for (size_t k = 0; k + 1 < ev.size(); k += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); *gemm_ms += ms; }
    CU(cudaMemcpyAsync(flag_h.data(), d_flag, (size_t)served * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    // flag 1: candidates overflowed (pathological ties / clustered scores) or the bf16 guard could not prove the shortlist;
    // flag 2: the estimated threshold came out too high.  Provable threshold: 1 -> exact path (2 cannot happen).  Estimated
    // threshold: both -> once more with the provable threshold and its wider candidate buffers, which sends what it
    // cannot serve either to the exact path.
    std::vector<int32_t> redo_users, retry_users; std::vector<int64_t> redo_rows, retry_rows;
    for (int64_t r = 0; r < served; ++r) {
        const int f = flag_h[(size_t)r];
        if (f == 0) continue;
        if (tau_mode || f == 2) { retry_users.push_back(h_users[r]); retry_rows.push_back(r); }
        else { redo_users.push_back(h_users[r]); redo_rows.push_back(r); }
    }
    if (top_level) s->tc_rows += n_users;
    s->tc_redo += (int64_t)redo_users.size(); s->tc_retry += (int64_t)retry_users.size();
    if (sink) {
        if (served < n_users) sink->whole = true;
        sink->dirty.insert(sink->dirty.end(), retry_rows.begin(), retry_rows.end());
        sink->dirty.insert(sink->dirty.end(), redo_rows.begin(), redo_rows.end());
    }
    if (tau_mode && n_batches == 1 && top_level) {
        if ((int64_t)retry_users.size() * 32 > n_users) { s->tau_spec_off = true; s->tau_spec_ok = false; }
        else if (n_users >= 1024) s->tau_spec_ok = true;
    }
    if (!tau_mode && !retry_users.empty()) return fail(RFM_ERR_CUDA, "recommend: %zu rows fell short of a provable threshold (internal error)", retry_users.size());
    // (the scratch buffers above are reused by the calls below: everything this invocation needed from them is on the host)
    if (served < n_users) {                                          // the estimate was switched off after the first batch
        rc = recommend_tc(s, d_users + served, h_users + served, n_users - served, n_items, filter_previous, d_rec + (size_t)served * n_items, gemm_ms,
                          cand_cap, 0, false);
        if (rc) return rc;
    }
    if (!retry_users.empty()) {
        DevBuf<float> d_fix; DevBuf<int32_t> d_fix_users;
        if ((rc = d_fix_users.alloc(retry_users.size()))) return rc;
        if ((rc = d_fix.alloc(retry_users.size() * n_items))) return rc;
        CU(cudaMemcpyAsync(d_fix_users, retry_users.data(), retry_users.size() * 4, cudaMemcpyHostToDevice, s->st));
        if ((rc = recommend_tc(s, d_fix_users, retry_users.data(), (int64_t)retry_users.size(), n_items, filter_previous, d_fix, gemm_ms, cand_cap, 0, false))) return rc;
        if ((rc = scatter_rows(s, d_rec, d_fix, retry_rows, n_items))) return rc;
    }
    if (!redo_users.empty()) {
        DevBuf<float> d_fix; DevBuf<int32_t> d_fix_users;
        if ((rc = d_fix_users.alloc(redo_users.size()))) return rc;
        if ((rc = d_fix.alloc(redo_users.size() * n_items))) return rc;
        CU(cudaMemcpyAsync(d_fix_users, redo_users.data(), redo_users.size() * 4, cudaMemcpyHostToDevice, s->st));
        if ((rc = recommend_exact(s, d_fix_users, (int64_t)redo_users.size(), n_items, filter_previous, d_fix, nullptr))) return rc;
        if ((rc = scatter_rows(s, d_rec, d_fix, redo_rows, n_items))) return rc;
    }
    return RFM_OK;
}

// 0 = auto, 1 = force tensor-core path where it applies, 2 = force the exact path       (RANKFM_B200_RECOMMEND=auto|tc|exact)
static int recommend_mode()
{
    const char* e = getenv("RANKFM_B200_RECOMMEND");
    if (!e) return 0;
    return !strcmp(e, "tc") ? 1 : (!strcmp(e, "exact") ? 2 : 0);
}

// Plan: users whose shortlist (2n+16 [+ seen items]) fits a candidate tier go through the tensor cores, the rest (and
// everything when the shape is not supported / too small to pay off) through the exact path.  `order` receives a
// permutation of [0,n_users): narrow-tier users first, then wide-tier users, then the exact-path users.
struct RecommendPlan { int64_t n_narrow = 0, n_wide = 0; };

static RecommendPlan recommend_plan(rfm_session* s, const std::vector<int32_t>& hu, int32_t n_items, int32_t filter_previous, std::vector<int64_t>& order)
{
    const int64_t n = (int64_t)hu.size();
    order.resize((size_t)n);
    const int mode = recommend_mode();
    const Tables& T = s->T;
    bool tc = mode != 2 && gemm_supported(T) && encode_ok() && (mode == 1 || ((int64_t)T.I >= 32768 && n * (int64_t)T.I >= ((int64_t)1 << 26)));
    // the per-row threshold is the n'-th largest bound over 8-item blocks: needs comfortably more blocks than n'
    const int blocks_half = tau_blocks(T, tau_stride(T, n_items)) / 2;
    const int limit_narrow = std::min(kCandCap, blocks_half), limit_wide = std::min(kCandCapWide, blocks_half);
    if (2 * n_items + 16 > limit_wide) tc = false;
    std::vector<int64_t> wide, exact;
    RecommendPlan plan;
    for (int64_t k = 0; k < n; ++k) {
        const int need = tc ? shortlist_target(s, hu[(size_t)k], n_items, filter_previous) : 0;
        if (tc && need <= limit_narrow) order[(size_t)plan.n_narrow++] = k;
        else if (tc && need <= limit_wide) wide.push_back(k);
        else exact.push_back(k);
    }
    plan.n_wide = (int64_t)wide.size();
    std::copy(wide.begin(), wide.end(), order.begin() + plan.n_narrow);
    std::copy(exact.begin(), exact.end(), order.begin() + plan.n_narrow + plan.n_wide);
    return plan;
}

static int recommend_dev(rfm_session* s, const int32_t* d_users, const int32_t* h_users, const RecommendPlan& plan, int64_t n_users, int32_t n_items,
                         int32_t filter_previous, float* d_rec, float* gemm_ms, HostSink* sink = nullptr)
{
    if (gemm_ms) *gemm_ms = 0.f;
    const int tau_mode = s->tau_spec_off ? 0 : tau_mode_default();
    int rc = recommend_tc(s, d_users, h_users, plan.n_narrow, n_items, filter_previous, d_rec, gemm_ms, kCandCap, tau_mode, true, sink);
    if (rc) return rc;
    const int64_t o1 = plan.n_narrow, o2 = plan.n_narrow + plan.n_wide;
    rc = recommend_tc(s, d_users + o1, h_users + o1, plan.n_wide, n_items, filter_previous, d_rec + (size_t)o1 * n_items, gemm_ms, kCandCapWide, tau_mode);
    if (rc) return rc;
    return recommend_exact(s, d_users + o2, n_users - o2, n_items, filter_previous, d_rec + (size_t)o2 * n_items, gemm_ms);
}

static int recommend_checks(rfm_session* s, int64_t n_users, int32_t n_items, int32_t filter_previous)
{
    if (n_items < 1) return fail(RFM_ERR_ARG, "n_items must be >= 1");
    if (n_items > 16384) return fail(RFM_ERR_UNSUPPORTED, "n_items > 16384 is not supported");
    if (filter_previous && !s->d_indptr) return fail(RFM_ERR_ARG, "filter_previous needs a session created with user_items (CSR)");
    (void)n_users;
    return RFM_OK;
}

extern "C" int rfm_session_recommend(rfm_session* s, const float* users, int64_t n_users, int32_t n_items, int32_t filter_previous, float* rec_items)
{
    if (!s || (n_users > 0 && (!users || !rec_items))) return fail(RFM_ERR_ARG, "NULL argument");
    if (n_users == 0) return RFM_OK;
    int rc = recommend_checks(s, n_users, n_items, filter_previous);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    std::vector<int32_t> hu;
    users_to_int(users, n_users, hu);
    for (auto u : hu) if (u >= 0 && (u < s->T.u0 || u >= s->T.u0 + s->T.Un)) return fail(RFM_ERR_ARG, "user index %d out of range [%d, %d)", u, s->T.u0, s->T.u0 + s->T.Un);
    std::vector<int64_t> order;
    const RecommendPlan plan = recommend_plan(s, hu, n_items, filter_previous, order);
    std::vector<int32_t> hp((size_t)n_users);
    for (int64_t k = 0; k < n_users; ++k) hp[(size_t)k] = hu[(size_t)order[(size_t)k]];
    DevBuf<int32_t> d_users; DevBuf<float> d_rec;
    if ((rc = d_users.alloc((size_t)n_users))) return rc;
    if ((rc = d_rec.alloc((size_t)n_users * n_items))) return rc;
    CU(cudaMemcpyAsync(d_users, hp.data(), (size_t)n_users * 4, cudaMemcpyHostToDevice, s->st));
    // the common large call -- every user in the narrow tensor-core tier, hence already in the caller's order -- streams
    // finished batches into the caller's buffer while later ones compute
    HostSink sink;
    // (from 64 MB of results: below that one copy after the last kernel measured no slower; RANKFM_B200_STREAM_MIN = floats)
    const char* smin = getenv("RANKFM_B200_STREAM_MIN");
    const bool stream_out = plan.n_narrow == n_users && (int64_t)n_users * n_items >= (smin ? atoll(smin) : ((int64_t)1 << 24));
    if (stream_out) { sink.out = rec_items; CU(pool_stream(s->device, &sink.cs)); }
    rc = recommend_dev(s, d_users, hp.data(), plan, n_users, n_items, filter_previous, d_rec, nullptr, stream_out ? &sink : nullptr);
    if (stream_out) {
        std::vector<cudaEvent_t> none;
        cudaStreamSynchronize(sink.cs);
        pool_return(s->device, sink.cs, none);
    }
    if (rc) return rc;
    if (stream_out && !sink.whole && (int64_t)sink.dirty.size() * 64 <= n_users) {      // few rewritten rows: copy just those again
        for (int64_t r : sink.dirty)
            CU(cudaMemcpyAsync(rec_items + (size_t)r * n_items, d_rec + (size_t)r * n_items, (size_t)n_items * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        return RFM_OK;
    }
    // back to the caller's order on the device (nothing to do when every user took the same path), then ONE copy straight
    // into the caller's buffer: a host staging vector + row-wise memcpy cost more than the GPU work at 1 M users x 100
    bool identity = true;
    for (int64_t k = 0; k < n_users && identity; ++k) identity = order[(size_t)k] == k;
    const float* d_final = d_rec;
    DevBuf<float> d_perm; DevBuf<int64_t> d_order;
    if (!identity) {
        if ((rc = d_perm.alloc((size_t)n_users * n_items))) return rc;
        if ((rc = d_order.alloc((size_t)n_users))) return rc;
        CU(cudaMemcpyAsync(d_order, order.data(), (size_t)n_users * 8, cudaMemcpyHostToDevice, s->st));
        CU(launch_scatter_rows(d_rec, d_order, n_users, n_items, d_perm, s->st));
        s->launches += 1;
        d_final = d_perm;
    }
    CU(cudaMemcpyAsync(rec_items, d_final, (size_t)n_users * n_items * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

// evaluation.py:9-143 on the device: recommend top-k for the users, then test every recommendation against the user's
// hold-out items; only the five sums (and optionally the hit matrix) travel back to the host
extern "C" int rfm_session_evaluate(rfm_session* s, const float* users, int64_t n_users, int32_t k, int32_t filter_previous,
                                    const int64_t* test_indptr, const int32_t* test_items, const int32_t* n_test, double* out5, uint8_t* hits_out)
{
    if (!s || !users || !test_indptr || !n_test || !out5 || n_users < 1) return fail(RFM_ERR_ARG, "bad argument");
    int rc = recommend_checks(s, n_users, k, filter_previous);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    std::vector<int32_t> hu;
    users_to_int(users, n_users, hu);
    for (auto u : hu) if (u < s->T.u0 || u >= s->T.u0 + s->T.Un) return fail(RFM_ERR_ARG, "user index %d out of range [%d, %d)", u, s->T.u0, s->T.u0 + s->T.Un);
    std::vector<int64_t> order;
    const RecommendPlan plan = recommend_plan(s, hu, k, filter_previous, order);
    std::vector<int32_t> hp((size_t)n_users);
    for (int64_t r = 0; r < n_users; ++r) hp[(size_t)r] = hu[(size_t)order[(size_t)r]];
    const int64_t nnz = test_indptr[n_users];
    if (nnz < 0 || (nnz > 0 && !test_items)) return fail(RFM_ERR_ARG, "bad test CSR");
    DevBuf<int32_t> d_users, d_items, d_ntest; DevBuf<float> d_rec; DevBuf<int64_t> d_order, d_ptr; DevBuf<double> d_out; DevBuf<uint8_t> d_hits;
    if ((rc = d_users.alloc((size_t)n_users))) return rc;
    if ((rc = d_rec.alloc((size_t)n_users * k))) return rc;
    if ((rc = d_order.alloc((size_t)n_users))) return rc;
    if ((rc = d_ptr.alloc((size_t)n_users + 1))) return rc;
    if ((rc = d_items.alloc((size_t)nnz))) return rc;
    if ((rc = d_ntest.alloc((size_t)n_users))) return rc;
    if ((rc = d_out.alloc(5))) return rc;
    if (hits_out && (rc = d_hits.alloc((size_t)n_users * k))) return rc;
    CU(cudaMemcpyAsync(d_users, hp.data(), (size_t)n_users * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(d_order, order.data(), (size_t)n_users * 8, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(d_ptr, test_indptr, ((size_t)n_users + 1) * 8, cudaMemcpyHostToDevice, s->st));
    if (nnz) CU(cudaMemcpyAsync(d_items, test_items, (size_t)nnz * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(d_ntest, n_test, (size_t)n_users * 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemsetAsync(d_out, 0, 5 * sizeof(double), s->st));
    if ((rc = recommend_dev(s, d_users, hp.data(), plan, n_users, k, filter_previous, d_rec, nullptr))) return rc;
    cudaError_t e = launch_eval_topk(d_rec, d_order, (int)n_users, k, d_ptr, d_items, d_ntest, d_out, hits_out ? (uint8_t*)d_hits : nullptr, s->st);
    if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "eval_topk launch failed: %s", cudaGetErrorString(e));
    s->launches += 1;
    CU(cudaMemcpyAsync(out5, d_out, 5 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    if (hits_out) CU(cudaMemcpyAsync(hits_out, d_hits, (size_t)n_users * k, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int m = 0; m < 5; ++m) out5[m] /= (double)n_users;
    return RFM_OK;
}

extern "C" int rfm_session_time_recommend(rfm_session* s, const float* users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                                          int32_t iters, float* ms_out, float* gemm_ms_out)
{
    if (!s || !users || !ms_out || n_users <= 0 || iters < 1) return fail(RFM_ERR_ARG, "bad argument");
    int rc = recommend_checks(s, n_users, n_items, filter_previous);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    std::vector<int32_t> hu;
    users_to_int(users, n_users, hu);
    std::vector<int64_t> order;
    const RecommendPlan plan = recommend_plan(s, hu, n_items, filter_previous, order);
    std::vector<int32_t> hp((size_t)n_users);
    for (int64_t k = 0; k < n_users; ++k) hp[(size_t)k] = hu[(size_t)order[(size_t)k]];
    DevBuf<int32_t> d_users; DevBuf<float> d_rec;
    if ((rc = d_users.alloc((size_t)n_users))) return rc;
    if ((rc = d_rec.alloc((size_t)n_users * n_items))) return rc;
    CU(cudaMemcpyAsync(d_users, hp.data(), (size_t)n_users * 4, cudaMemcpyHostToDevice, s->st));
    if ((rc = recommend_dev(s, d_users, hp.data(), plan, n_users, n_items, filter_previous, d_rec, nullptr))) return rc;   // warm-up
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    float gemm_total = 0.f;
    CU(cudaEventRecord(a, s->st));
    for (int k = 0; k < iters; ++k) {
        float g = 0.f;
        if ((rc = recommend_dev(s, d_users, hp.data(), plan, n_users, n_items, filter_previous, d_rec, gemm_ms_out ? &g : nullptr))) return rc;
        gemm_total += g;
    }
    CU(cudaEventRecord(b, s->st));
    CU(cudaStreamSynchronize(s->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / iters;
    if (gemm_ms_out) *gemm_ms_out = gemm_total / iters;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return RFM_OK;
}

// debugging / parity: the bf16 tensor-core scores S = A.B^T + bias of the requested users against ALL items, dense
extern "C" int rfm_session_debug_gemm(rfm_session* s, const float* users, int64_t n_users, float* scores_out /* [n_users, I] */)
{
    if (!s || !users || !scores_out || n_users <= 0) return fail(RFM_ERR_ARG, "bad argument");
    if (!gemm_supported(s->T) || !encode_ok()) return fail(RFM_ERR_UNSUPPORTED, "tensor-core scoring not available for this shape");
    CU(cudaSetDevice(s->device));
    int rc = ensure_gemm_items(s);
    if (rc) return rc;
    std::vector<int32_t> hu;
    users_to_int(users, n_users, hu);
    const Tables& T = s->T;
    const int Kp = gemm_kp(T), I_pad = s->gemm_I_pad, MT = gemm_m_tile(T), M_pad = (int)((n_users + MT - 1) / MT * MT);
    DevBuf<int32_t> d_users; DevBuf<__nv_bfloat16_raw> d_A; DevBuf<float> d_S;
    if ((rc = d_users.alloc((size_t)n_users))) return rc;
    if ((rc = d_A.alloc((size_t)M_pad * Kp))) return rc;
    if ((rc = d_S.alloc((size_t)M_pad * I_pad))) return rc;
    CU(cudaMemcpyAsync(d_users, hu.data(), (size_t)n_users * 4, cudaMemcpyHostToDevice, s->st));
    CU(launch_pack_gemm_users(T, d_users, (int)n_users, M_pad, Kp, d_A, s->st));
    cudaError_t e = launch_score_filter(T, 0, d_A, s->d_gemm_B, s->d_gemm_bias, (int)n_users, M_pad, I_pad, 1, 1, nullptr, nullptr, nullptr, 0, nullptr, d_S, s->st);
    if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "score_filter (tcgen05) launch failed: %s", cudaGetErrorString(e));
    // the kernel's columns are positions of the bias-ordered catalogue: scatter them back to item indexes
    std::vector<float> tmp((size_t)n_users * I_pad);
    std::vector<int32_t> order((size_t)T.I);
    CU(cudaMemcpyAsync(tmp.data(), d_S, (size_t)n_users * I_pad * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(order.data(), s->d_gemm_order, (size_t)T.I * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int64_t r = 0; r < n_users; ++r)
        for (int pos = 0; pos < T.I; ++pos) scores_out[(size_t)r * T.I + order[(size_t)pos]] = tmp[(size_t)r * I_pad + pos];
    return RFM_OK;
}

extern "C" int rfm_session_trace_enable(rfm_session* s)
{
    if (!s || s->N <= 0) return fail(RFM_ERR_ARG, "trace needs a training session");
    CU(cudaSetDevice(s->device));
    if (!s->d_trace) { int rc = dev_alloc(&s->d_trace, (size_t)s->N * 2); if (rc) return rc; }
    CU(cudaMemsetAsync(s->d_trace, 0xff, (size_t)s->N * 8, s->st));
    return RFM_OK;
}

extern "C" int rfm_session_trace_read(rfm_session* s, int32_t* out)
{
    if (!s || !out || !s->d_trace) return fail(RFM_ERR_ARG, "trace not enabled");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(out, s->d_trace, (size_t)s->N * 8, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_recommend_retried(rfm_session* s, int64_t* tc_retried)
{
    if (!s || !tc_retried) return fail(RFM_ERR_ARG, "NULL argument");
    *tc_retried = s->tc_retry;
    return RFM_OK;
}

extern "C" int rfm_session_recommend_stats(rfm_session* s, int64_t* tc_rows, int64_t* tc_redone)
{
    if (!s || !tc_rows || !tc_redone) return fail(RFM_ERR_ARG, "NULL argument");
    *tc_rows = s->tc_rows; *tc_redone = s->tc_redo;
    return RFM_OK;
}

extern "C" int rfm_session_flush_l2(rfm_session* s)
{
    if (!s) return fail(RFM_ERR_ARG, "NULL session");
    CU(cudaSetDevice(s->device));
    if (!s->d_flush) {
        s->flush_bytes = (size_t)512 << 20;   // 4x the 126 MB L2
        CU(dev_malloc((void**)&s->d_flush, s->flush_bytes));
    }
    CU(cudaMemsetAsync(s->d_flush, 0x5a, s->flush_bytes, s->st));
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_exchange_path(rfm_session* s, int32_t* path)
{
    if (!s || !path) return fail(RFM_ERR_ARG, "NULL argument");
    *path = !s->comm ? 0 : (s->p2p ? 1 : 2);
    return RFM_OK;
}

extern "C" int rfm_session_launch_count(rfm_session* s, int64_t* launches)
{
    if (!s || !launches) return fail(RFM_ERR_ARG, "NULL argument");
    *launches = s->launches;
    return RFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// one-shot entry points on host buffers
// ---------------------------------------------------------------------------------------------------------------
static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static thread_local double g_fit_phases[4] = {0, 0, 0, 0};

extern "C" int rfm_last_fit_phases(double* out4)
{
    if (!out4) return fail(RFM_ERR_ARG, "out4 is NULL");
    for (int k = 0; k < 4; ++k) out4[k] = g_fit_phases[k];
    return RFM_OK;
}

extern "C" int rfm_fit(const rfm_problem* p, int32_t epochs, const int32_t* perms, rfm_epoch_stats* stats)
{
    const bool timing = getenv("RANKFM_B200_TIMING") != nullptr;
    const double t0 = now_ms();
    rfm_session* s = nullptr;
    int rc = rfm_session_create(p, &s);
    if (rc) return rc;
    const double t1 = now_ms();
    rc = rfm_session_train(s, epochs, perms, stats);
    const double t2 = now_ms();
    // like the reference, weights are written back even when they went non-finite (_rankfm.pyx:329 raises after mutation)
    if (rc == RFM_OK || rc == RFM_ERR_NONFINITE) {
        const std::string keep = g_err;
        int rc2 = rfm_session_download(s, p->w_i, p->w_if, p->v_u, p->v_i, p->v_uf, p->v_if);
        if (rc2) rc = rc2; else g_err = keep;
    }
    const double t3 = now_ms();
    rfm_session_destroy(s);
    const double t4 = now_ms();
    g_fit_phases[0] = t1 - t0; g_fit_phases[1] = t2 - t1; g_fit_phases[2] = t3 - t2; g_fit_phases[3] = t4 - t3;
    if (timing) fprintf(stderr, "[rfm_fit] create+H2D %.2f ms, train %.2f ms, D2H %.2f ms, destroy %.2f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    return rc;
}

extern "C" int rfm_predict(const rfm_problem* p, const float* pairs, int64_t n, float* scores)
{
    rfm_problem q = *p;
    q.n_interactions = 0;
    rfm_session* s = nullptr;
    int rc = rfm_session_create(&q, &s);
    if (rc) return rc;
    rc = rfm_session_predict(s, pairs, n, scores);
    rfm_session_destroy(s);
    return rc;
}

static int attach_csr(rfm_session* s, const rfm_problem* p)
{
    if (!p->csr_indptr || !p->csr_indices) return fail(RFM_ERR_ARG, "filter_previous needs csr_indptr / csr_indices");
    const int64_t nnz = p->csr_indptr[p->U];
    int rc;
    if (s->T.Un != s->T.U) return fail(RFM_ERR_UNSUPPORTED, "scoring with filter_previous needs a session that holds all users");
    if ((rc = dev_alloc(&s->indptr_alloc, (size_t)p->U + 1))) return rc;
    if ((rc = dev_alloc(&s->indices_alloc, (size_t)nnz))) return rc;
    s->d_indptr = s->indptr_alloc; s->d_indices = s->indices_alloc;
    CU(cudaMemcpyAsync(s->d_indptr, p->csr_indptr, ((size_t)p->U + 1) * 8, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->d_indices, p->csr_indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, s->st));
    s->h_indptr.assign(p->csr_indptr, p->csr_indptr + p->U + 1);
    CU(cudaStreamSynchronize(s->st));
    return RFM_OK;
}

extern "C" int rfm_session_attach_csr(rfm_session* s, const int64_t* indptr, const int32_t* indices)
{
    if (!s || !indptr || !indices) return fail(RFM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(s->device));
    dev_free(s->indptr_alloc); dev_free(s->indices_alloc);
    s->indptr_alloc = nullptr; s->indices_alloc = nullptr;
    s->d_indptr = nullptr; s->d_indices = nullptr;
    rfm_problem q = s->p;
    q.csr_indptr = indptr; q.csr_indices = indices;
    return attach_csr(s, &q);
}

extern "C" int rfm_recommend(const rfm_problem* p, const float* users, int64_t n_users, int32_t n_items, int32_t filter_previous, float* rec_items)
{
    rfm_problem q = *p;
    q.n_interactions = 0;
    rfm_session* s = nullptr;
    int rc = rfm_session_create(&q, &s);
    if (rc) return rc;
    if (filter_previous) rc = attach_csr(s, p);
    if (!rc) rc = rfm_session_recommend(s, users, n_users, n_items, filter_previous, rec_items);
    rfm_session_destroy(s);
    return rc;
}

extern "C" int rfm_session_similar(rfm_session* s, int32_t which, int32_t index, int32_t n, int32_t* out)
{
    if (!s || !out) return fail(RFM_ERR_ARG, "NULL argument");
    if (which != 0 && which != 1) return fail(RFM_ERR_ARG, "which must be 0 (items) or 1 (users)");
    const int rows = which == 0 ? s->T.I : s->T.U;
    if (index < 0 || index >= rows) return fail(RFM_ERR_ARG, "index out of range");
    if (which == 1 && s->T.Un != s->T.U) return fail(RFM_ERR_UNSUPPORTED, "similar_users needs a session that holds all users");
    if (n < 1 || n > 16384) return fail(RFM_ERR_ARG, "n out of range");
    CU(cudaSetDevice(s->device));
    DevBuf<float> qc, S, d_rec; DevBuf<int32_t> d_one, d_ex;
    int rc;
    if ((rc = qc.alloc((size_t)s->T.Fp + (size_t)std::max(s->T.Pp, s->T.Qp) + 4))) return rc;
    if ((rc = S.alloc((size_t)rows))) return rc;
    if ((rc = d_rec.alloc((size_t)n))) return rc;
    if ((rc = d_one.alloc(1))) return rc;
    if ((rc = d_ex.alloc(1))) return rc;
    const int32_t zero = 0;
    CU(cudaMemcpyAsync(d_one, &zero, 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(d_ex, &index, 4, cudaMemcpyHostToDevice, s->st));
    cudaError_t e = launch_latent_scores(s->T, which, index, qc, S, s->st);
    if (e == cudaSuccess) e = launch_topn_select(S, rows, d_one, 1, nullptr, nullptr, 0, n, d_rec, d_ex, s->st);
    if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "similar launch failed: %s", cudaGetErrorString(e));
    s->launches += 3;
    std::vector<float> rec((size_t)n);
    CU(cudaMemcpyAsync(rec.data(), d_rec, (size_t)n * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int k = 0; k < n; ++k) out[k] = std::isnan(rec[(size_t)k]) ? -1 : (int32_t)rec[(size_t)k];
    return RFM_OK;
}

// similar_items / similar_users for MANY query rows in one call (SURVEY 8(f)4 "batched all-items variant"): the queries
// are scored in chunks (one latent_scores pass per query into a [chunk, rows] score block) and every chunk gets ONE
// top-n launch; out int32 [n_queries, n], -1 pads rows with fewer than n other rows
extern "C" int rfm_session_similar_batch(rfm_session* s, int32_t which, const int32_t* indexes, int64_t n_queries, int32_t n, int32_t* out)
{
    if (!s || !indexes || !out || n_queries < 1) return fail(RFM_ERR_ARG, "bad argument");
    if (which != 0 && which != 1) return fail(RFM_ERR_ARG, "which must be 0 (items) or 1 (users)");
    const int rows = which == 0 ? s->T.I : s->T.U;
    if (which == 1 && s->T.Un != s->T.U) return fail(RFM_ERR_UNSUPPORTED, "similar_users needs a session that holds all users");
    if (n < 1 || n > 16384) return fail(RFM_ERR_ARG, "n out of range");
    for (int64_t q = 0; q < n_queries; ++q) if (indexes[q] < 0 || indexes[q] >= rows) return fail(RFM_ERR_ARG, "index out of range");
    CU(cudaSetDevice(s->device));
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_queries, ((int64_t)1 << 28) / std::max(1, rows)));      // <= 1 GiB of scores
    DevBuf<float> qc, S, d_rec; DevBuf<int32_t> d_zero, d_idx;
    int rc;
    const size_t qc_floats = (size_t)s->T.Fp + (size_t)std::max(s->T.Pp, s->T.Qp) + 4;
    if ((rc = qc.alloc(qc_floats))) return rc;
    if ((rc = S.alloc((size_t)chunk * rows))) return rc;
    if ((rc = d_rec.alloc((size_t)chunk * n))) return rc;
    if ((rc = d_zero.alloc((size_t)chunk))) return rc;
    if ((rc = d_idx.alloc((size_t)n_queries))) return rc;
    CU(cudaMemsetAsync(d_zero, 0, (size_t)chunk * 4, s->st));
    CU(cudaMemcpyAsync(d_idx, indexes, (size_t)n_queries * 4, cudaMemcpyHostToDevice, s->st));
    std::vector<float> rec((size_t)chunk * n);
    for (int64_t off = 0; off < n_queries; off += chunk) {
        const int nb = (int)std::min<int64_t>(chunk, n_queries - off);
        for (int b = 0; b < nb; ++b) {
            cudaError_t e = launch_latent_scores(s->T, which, indexes[off + b], qc, S + (size_t)b * rows, s->st);
            if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "similar launch failed: %s", cudaGetErrorString(e));
        }
        cudaError_t e = launch_topn_select(S, rows, d_zero, nb, nullptr, nullptr, 0, n, d_rec, d_idx + off, s->st);
        if (e != cudaSuccess) return fail(RFM_ERR_CUDA, "similar top-n launch failed: %s", cudaGetErrorString(e));
        s->launches += 2 * nb + 1;
        CU(cudaMemcpyAsync(rec.data(), d_rec, (size_t)nb * n * 4, cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        for (int64_t e2 = 0; e2 < (int64_t)nb * n; ++e2) out[(size_t)off * n + e2] = std::isnan(rec[(size_t)e2]) ? -1 : (int32_t)rec[(size_t)e2];
    }
    return RFM_OK;
}

extern "C" int rfm_similar(const rfm_problem* p, int32_t which, int32_t index, int32_t n, int32_t* out)
{
    if (!p || !out) return fail(RFM_ERR_ARG, "NULL argument");
    rfm_problem q = *p;
    q.n_interactions = 0;
    rfm_session* s = nullptr;
    int rc = rfm_session_create(&q, &s);
    if (rc) return rc;
    rc = rfm_session_similar(s, which, index, n, out);
    rfm_session_destroy(s);
    return rc;
}
