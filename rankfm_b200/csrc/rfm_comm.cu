// rfm_comm.cu -- multi-GPU communicator of the training path (one process per GPU): NCCL bootstrap + the fused
// peer-memory exchange of the replicated item table.
//
// The reference trains one model in one thread (`_fit`, rankfm/_rankfm.pyx:218-326); SURVEY.md 8(e) shards that loop by
// user: user rows are owned by one rank, the item table (v_i, w_i) and the feature parameters (w_if, v_uf, v_if) are
// replicated and folded once per epoch:   table <- snapshot + sum_r gain_r(row) * (table_r - snapshot).
//
// Data path (default): every rank's replica lives in a cudaMalloc'ed WINDOW that the other ranks map through cudaIpc
// (NVLink 5 / NVSwitch peer access).  One kernel per epoch -- exchange_kernel -- replaces
//     delta kernel -> ncclAllReduce -> apply kernel          (three passes over the table + a library collective)
// by a reduce-scatter / all-gather written directly over peer memory: rank r owns rows [I r/C, I (r+1)/C), reads those
// rows of every replica (peer loads), folds them against the (identical) start-of-epoch snapshot, and stores the result
// into every replica (peer stores).  Only the factor + bias columns travel (not the read-only x_if block or the pads).
// Ranks meet at two in-kernel barriers (release/acquire flags at system scope in each other's window headers).
// NCCL is used for the bootstrap (exchange of the 64-byte IPC handles), for tiny end-of-training reductions, and as the
// fallback data path when peer mapping is not available.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <vector>

#include "rfm_host.h"

using namespace rfm;

namespace rfmh {

// ---------------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU use has no NCCL dependency
// ---------------------------------------------------------------------------------------------------------------
struct Id128 { char b[128]; };   // ncclUniqueId is 128 opaque bytes, passed by value
namespace {
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_mu;
}  // namespace

static int nccl_load()
{
    if (g_nccl.h) return RFM_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(RFM_ERR_NCCL, "libnccl.so.2 not found: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy)
        return fail(RFM_ERR_NCCL, "libnccl.so.2 lacks required symbols");
    g_nccl.h = h;
    return RFM_OK;
}
#define NC(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0) return fail(RFM_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
    } while (0)
constexpr int kNcclChar = 0, kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;   // nccl.h enum values, stable across 2.x

int nccl_unique_id(uint8_t* out128)
{
    int rc = nccl_load();
    if (rc) return rc;
    NC(g_nccl.GetUniqueId(out128));
    return RFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// peer window
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;                 // one NVSwitch domain
constexpr size_t kHeaderBytes = 4096;

struct ExchHeader {                          // first kHeaderBytes of every rank's window
    uint32_t flag[kMaxPeers * 32];           // flag[32 r]: last barrier sequence number rank r signalled HERE (one 128-byte line each)
    uint32_t grid_count;                     // blocks of the local exchange grid that reached a barrier (cumulative)
    uint32_t err;                            // 1: a barrier wait timed out
    float neg_per_item;                      // this rank's uniform negatives per item in the epoch being folded
    uint32_t pad;
};
static_assert(sizeof(ExchHeader) <= kHeaderBytes, "window header too large");

struct Comm {
    uint8_t id[128];
    int rank = 0, world = 1, device = 0;
    void* nccl = nullptr;
    int refs = 0;
    // window
    void* win = nullptr; size_t win_bytes = 0;
    void* peer[kMaxPeers] = {nullptr};       // peer[r]: rank r's window in this process' address space (peer[rank] == win)
    bool p2p = false;                        // windows are mapped both ways on every rank
    const void* owner = nullptr;             // the session whose tables sit in the window
    ExchangeShape shape{};
    size_t off_touch = 0, off_gp = 0, off_gpnew = 0, off_it = 0;
    bool p2p_failed = false;                 // peer mapping was tried and is not possible on this machine: stay on NCCL
    uint32_t seq = 0, arrivals = 0;          // barrier sequence numbers / grid arrivals consumed so far
    int grid = 0;
    void* d_scratch = nullptr;               // small device buffer for bootstrap / reductions
};
static std::vector<Comm*> g_comms;

int comm_world(const Comm* c) { return c ? c->world : 1; }

static bool p2p_wanted()
{
    const char* e = getenv("RANKFM_B200_EXCHANGE");       // p2p (default) | nccl
    return !(e && !strcmp(e, "nccl"));
}

int comm_acquire(const uint8_t* id128, int rank, int world, int device, Comm** out)
{
    *out = nullptr;
    if (!id128) return fail(RFM_ERR_ARG, "world>1 needs nccl_id");
    int rc = nccl_load();
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    for (Comm* c : g_comms)
        if (c->rank == rank && c->world == world && c->device == device && !memcmp(c->id, id128, 128)) { ++c->refs; *out = c; return RFM_OK; }
    Comm* c = new Comm();
    memcpy(c->id, id128, 128);
    c->rank = rank; c->world = world; c->device = device;
    Id128 id;
    memcpy(id.b, id128, 128);
    int r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
    if (r != 0) { delete c; return fail(RFM_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); }
    if (cudaMalloc(&c->d_scratch, 64 * (kMaxPeers + 1) + 4096) != cudaSuccess) { g_nccl.CommDestroy(c->nccl); delete c; return fail(RFM_ERR_CUDA, "cudaMalloc failed"); }
    c->refs = 1;
    g_comms.push_back(c);
    *out = c;
    return RFM_OK;
}

void comm_release(Comm* c)
{
    if (!c) return;
    std::lock_guard<std::mutex> lock(g_mu);
    if (c->refs > 0) --c->refs;
}

int comm_allreduce_f32(Comm* c, float* buf, size_t n, cudaStream_t st)
{
    NC(g_nccl.AllReduce(buf, buf, n, kNcclFloat, kNcclSum, c->nccl, st));
    return RFM_OK;
}
int comm_allreduce_f64(Comm* c, double* buf, size_t n, cudaStream_t st)
{
    NC(g_nccl.AllReduce(buf, buf, n, kNcclDouble, kNcclSum, c->nccl, st));
    return RFM_OK;
}

// every rank learns whether ALL ranks succeeded (collective)
static int all_ok(Comm* c, bool mine, cudaStream_t st, bool* everyone)
{
    float v = mine ? 1.0f : 0.0f;
    float* d = reinterpret_cast<float*>(static_cast<char*>(c->d_scratch) + 64 * (kMaxPeers + 1));
    CU(cudaMemcpyAsync(d, &v, 4, cudaMemcpyHostToDevice, st));
    NC(g_nccl.AllReduce(d, d, 1, kNcclFloat, kNcclSum, c->nccl, st));
    CU(cudaMemcpyAsync(&v, d, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *everyone = v > (float)c->world - 0.5f;
    return RFM_OK;
}

static void window_close(Comm* c)
{
    for (int r = 0; r < c->world && r < kMaxPeers; ++r) {
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
        c->peer[r] = nullptr;
    }
    c->p2p = false;
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int comm_window_attach(Comm* c, const ExchangeShape& shape, cudaStream_t st, const void* owner, float** IT, float** GP, int32_t** touch, bool* p2p)
{
    *p2p = false;
    if (!p2p_wanted() || c->world > kMaxPeers || c->owner || c->p2p_failed) return RFM_OK;      // same decision on every rank: sessions are created collectively
    const size_t off_touch = kHeaderBytes;
    const size_t off_gp = off_touch + round_up((size_t)shape.I * 4, 256);
    const size_t off_gpnew = off_gp + round_up(shape.gp_floats * 4 + 16, 256);
    const size_t off_it = off_gpnew + round_up(shape.gp_floats * 4 + 16, 256);
    const size_t need = off_it + (size_t)shape.I * shape.ldi * 4;
    if (c->win_bytes < need || !c->p2p) {
        // (re)build the window: quiesce, unmap the peers' old windows, free ours, allocate, exchange IPC handles over NCCL
        bool dummy;
        int rc = all_ok(c, true, st, &dummy);                                  // nobody is still inside an exchange on the old window
        if (rc) return rc;
        window_close(c);
        if ((rc = all_ok(c, true, st, &dummy))) return rc;                     // every rank has unmapped before anyone frees
        if (c->win) { cudaFree(c->win); c->win = nullptr; c->win_bytes = 0; }
        bool ok = cudaMalloc(&c->win, need) == cudaSuccess;
        if (!ok) { cudaGetLastError(); c->win = nullptr; }
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof mine);
        if (ok) ok = cudaIpcGetMemHandle(&mine, c->win) == cudaSuccess;
        if (ok) ok = cudaMemsetAsync(c->win, 0, kHeaderBytes, st) == cudaSuccess;
        if (!ok) cudaGetLastError();
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        char* d_send = static_cast<char*>(c->d_scratch);
        char* d_recv = d_send + 64;
        CU(cudaMemcpyAsync(d_send, &mine, 64, cudaMemcpyHostToDevice, st));
        NC(g_nccl.AllGather(d_send, d_recv, 64, kNcclChar, c->nccl, st));
        std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
        CU(cudaMemcpyAsync(all.data(), d_recv, 64 * (size_t)c->world, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        bool everyone = false;
        if ((rc = all_ok(c, ok, st, &everyone))) return rc;                    // every rank exported a window
        if (everyone) {
            for (int r = 0; r < c->world && ok; ++r) {
                if (r == c->rank) { c->peer[r] = c->win; continue; }
                if (cudaIpcOpenMemHandle(&c->peer[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); c->peer[r] = nullptr; ok = false; }
            }
        }
        if ((rc = all_ok(c, ok && everyone, st, &everyone))) return rc;        // every rank mapped every window
        if (!everyone) {
            window_close(c);
            if (c->win) { cudaFree(c->win); c->win = nullptr; }
            c->win_bytes = 0;
            c->p2p_failed = true;
            return RFM_OK;                                                     // NCCL data path
        }
        c->win_bytes = need;
        c->p2p = true;
        c->seq = 0; c->arrivals = 0;
        if (!c->grid) {
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device);
            c->grid = n_sm * 2;
        }
    }
    c->shape = shape;
    c->off_touch = off_touch; c->off_gp = off_gp; c->off_gpnew = off_gpnew; c->off_it = off_it;
    c->owner = owner;
    char* w = static_cast<char*>(c->win);
    *touch = reinterpret_cast<int32_t*>(w + off_touch);
    *GP = reinterpret_cast<float*>(w + off_gp);
    *IT = reinterpret_cast<float*>(w + off_it);
    *p2p = true;
    return RFM_OK;
}

void comm_window_detach(Comm* c, const void* owner)
{
    if (c && c->owner == owner) c->owner = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// the fused exchange kernel
// ---------------------------------------------------------------------------------------------------------------
struct ExchParams {
    float* it[kMaxPeers];
    const int32_t* touch[kMaxPeers];
    float* gp[kMaxPeers];
    ExchHeader* hdr[kMaxPeers];
    float *snap_it, *snap_gp, *gp_new;
    int32_t rank, world, I, ldi, NQ, gp_floats;
    float lam_factor, lam_bias, gp_gain;
    uint32_t seq, arrivals;
    const EpochAcc* acc;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys4(const float* p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys4(float* p, float4 v)
{
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_sys1(const float* p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_sys_i32(const int32_t* p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr long long kBarrierTimeoutCycles = 60000000000ll;       // ~30 s: a rank that never arrives must not hang the others forever

// All blocks of the local grid AND all ranks: nobody passes before every block of every rank has arrived, and everything
// written before the barrier (by this grid, or by earlier kernels of any rank's stream) is visible after it.
__device__ void exchange_barrier(const ExchParams& p, uint32_t seq, uint32_t arrive_target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        ExchHeader* me = p.hdr[p.rank];
        __threadfence_system();
        const uint32_t prev = atomicAdd(&me->grid_count, 1u);
        if (prev + 1u == arrive_target) {                 // last local block: signal every rank (this one included)
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(&p.hdr[r]->flag[32 * p.rank], seq);
        }
        const long long t0 = clock64();
        for (int r = 0; r < p.world; ++r) {
            while ((int32_t)(ld_acquire_sys(&me->flag[32 * r]) - seq) < 0) {
                if (clock64() - t0 > kBarrierTimeoutCycles) { me->err = 1u; break; }
                __nanosleep(20);
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}

// gain of one replica's delta when C replicas of a row are folded; x = -log of the row's own contraction over the epoch
// (DESIGN.md "Per-row fold gain"): 1 (sum) for rows that barely moved, 1/C (average) for rows that forgot their start
__device__ __forceinline__ float fold_gain_dev(float x, float C)
{
    if (x <= 1e-4f) return 1.0f;
    return (1.0f - __expf(-x * C)) / (C * (1.0f - __expf(-x)));
}

template <int W>
__global__ void __launch_bounds__(256) exchange_kernel(const ExchParams p)
{
    float* gp_new = p.gp_new;
    ExchHeader* me = p.hdr[p.rank];
    if (blockIdx.x == 0 && threadIdx.x == 0) me->neg_per_item = (float)((double)p.acc->draws / (double)p.I);
    const uint32_t G = gridDim.x;
    exchange_barrier(p, p.seq + 1u, p.arrivals + G);                         // A: every rank finished its SGD epoch

    const float C = (float)p.world;
    float neg[W];
#pragma unroll
    for (int r = 0; r < W; ++r) neg[r] = r < p.world ? ld_sys1(&p.hdr[r]->neg_per_item) : 0.f;
    const int row0 = (int)((long long)p.I * p.rank / p.world), row1 = (int)((long long)p.I * (p.rank + 1) / p.world);
    const int nq = p.NQ + 1;                                                  // factor quads + the bias quad
    const long long total = (long long)(row1 - row0) * nq;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)G * blockDim.x;
    for (long long e = tid; e < total; e += nth) {
        const int row = row0 + (int)(e / nq), q = (int)(e % nq);
        const size_t off = (size_t)row * p.ldi + 4 * (size_t)q;
        const float4 s = *reinterpret_cast<const float4*>(p.snap_it + off);
        float4 v[W];
        int t[W];
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < p.world) { v[r] = ld_sys4(p.it[r] + off); t[r] = ld_sys_i32(p.touch[r] + row); }
        const float lam = q == p.NQ ? p.lam_bias : p.lam_factor;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < p.world) {
                const float g = fold_gain_dev(lam * ((float)t[r] + neg[r]), C);
                a.x += g * (v[r].x - s.x); a.y += g * (v[r].y - s.y); a.z += g * (v[r].z - s.z); a.w += g * (v[r].w - s.w);
            }
        const float4 nv = make_float4(s.x + a.x, s.y + a.y, s.z + a.z, s.w + a.w);
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < p.world) st_sys4(p.it[r] + off, nv);
        *reinterpret_cast<float4*>(p.snap_it + off) = nv;
    }
    if (blockIdx.x == 0 && p.gp_floats > 0) {        // feature parameters: every rank folds all replicas itself (a few KB)
        for (int e = threadIdx.x; e < p.gp_floats; e += blockDim.x) {
            const float s = p.snap_gp[e];
            float a = 0.f;
            for (int r = 0; r < p.world; ++r) a += ld_sys1(p.gp[r] + e) - s;
            gp_new[e] = s + p.gp_gain * a;
        }
    }
    exchange_barrier(p, p.seq + 2u, p.arrivals + 2u * G);                    // B: every replica holds the folded rows; all reads of GP are done

    // the rows folded by the other ranks have landed in the local table: refresh their snapshot
    float* mine = p.it[p.rank];
    const long long all = (long long)p.I * nq, own0 = (long long)row0 * nq, own1 = (long long)row1 * nq;
    for (long long e = tid; e < all; e += nth) {
        if (e >= own0 && e < own1) continue;
        const size_t off = (size_t)(e / nq) * p.ldi + 4 * (size_t)(e % nq);
        *reinterpret_cast<float4*>(p.snap_it + off) = ld_sys4(mine + off);
    }
    if (blockIdx.x == 0 && p.gp_floats > 0) {
        float* gp = p.gp[p.rank];
        for (int e = threadIdx.x; e < p.gp_floats; e += blockDim.x) { gp[e] = gp_new[e]; p.snap_gp[e] = gp_new[e]; }
    }
}

int comm_exchange_p2p(Comm* c, cudaStream_t st, float* snap_it, float* snap_gp, const EpochAcc* acc, float lam_factor, float lam_bias, float gp_gain)
{
    if (!c->p2p || !c->owner) return fail(RFM_ERR_NCCL, "peer window not attached");
    ExchParams p{};
    for (int r = 0; r < c->world; ++r) {
        char* w = static_cast<char*>(c->peer[r]);
        p.hdr[r] = reinterpret_cast<ExchHeader*>(w);
        p.touch[r] = reinterpret_cast<const int32_t*>(w + c->off_touch);
        p.gp[r] = reinterpret_cast<float*>(w + c->off_gp);
        p.it[r] = reinterpret_cast<float*>(w + c->off_it);
    }
    p.snap_it = snap_it; p.snap_gp = snap_gp;
    p.rank = c->rank; p.world = c->world; p.I = c->shape.I; p.ldi = c->shape.ldi; p.NQ = c->shape.NQ;
    p.gp_floats = snap_gp ? (int)c->shape.gp_floats : 0;
    p.lam_factor = lam_factor; p.lam_bias = lam_bias; p.gp_gain = gp_gain;
    p.seq = c->seq; p.arrivals = c->arrivals;
    p.acc = acc;
    p.gp_new = reinterpret_cast<float*>(static_cast<char*>(c->win) + c->off_gpnew);
    // every block spins at the barriers: the whole grid must be resident (at most 2 blocks of 256 threads and no shared
    // memory per SM always are, and nothing else runs on the session's stream at this point).  Small tables get a small
    // grid: the exchange of a 350 KB table is two barriers and a few microseconds of copies, and every block more is one
    // more arrival to count and one more spinner.
    const long long quads = (long long)p.I * (p.NQ + 1);
    const int grid = (int)std::min<long long>(c->grid, std::max<long long>(8, (quads + 1023) / 1024));
    if (c->world <= 2) exchange_kernel<2><<<grid, 256, 0, st>>>(p);
    else if (c->world <= 4) exchange_kernel<4><<<grid, 256, 0, st>>>(p);
    else exchange_kernel<8><<<grid, 256, 0, st>>>(p);
    CU(cudaGetLastError());
    c->seq += 2u;
    c->arrivals += 2u * (uint32_t)grid;
    return RFM_OK;
}

int comm_check(Comm* c, cudaStream_t st)
{
    if (!c || !c->p2p || !c->win) return RFM_OK;
    uint32_t err = 0;
    CU(cudaMemcpyAsync(&err, &reinterpret_cast<ExchHeader*>(c->win)->err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (err) return fail(RFM_ERR_NCCL, "multi-GPU exchange: a peer did not reach the epoch barrier within 30 s");
    return RFM_OK;
}

}  // namespace rfmh

extern "C" int rfm_comm_release_all(void)
{
    using namespace rfmh;
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto it = g_comms.begin(); it != g_comms.end();) {
        Comm* c = *it;
        if (c->refs > 0) { ++it; continue; }
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        bool dummy;
        all_ok(c, true, nullptr, &dummy);            // collective: nobody unmaps while a peer is still inside an exchange ...
        window_close(c);
        all_ok(c, true, nullptr, &dummy);            // ... and nobody frees a window a peer still maps
        if (c->win) cudaFree(c->win);
        if (c->d_scratch) cudaFree(c->d_scratch);
        if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
        delete c;
        it = g_comms.erase(it);
    }
    return RFM_OK;
}
