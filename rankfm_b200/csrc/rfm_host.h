// rfm_host.h -- host-side plumbing shared by the translation units behind the C ABI (internal; the public ABI is
// include/rankfm_b200.h): error reporting, the device block cache, and the multi-GPU communicator (rfm_comm.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rankfm_b200.h"
#include "rfm_kernels.h"

namespace rfmh {

// sets the thread-local message returned by rfm_last_error() and hands `code` back
int fail(int code, const char* fmt, ...);

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return rfmh::fail(RFM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// device block cache (rfm_api.cu): freed blocks are kept per size class for the next call of the same shape
cudaError_t dev_malloc(void** out, size_t bytes);
void dev_free(void* p);

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU communicator: one per (nccl_id, rank, world, device), kept by the library across sessions.
//
// Two data paths for the per-epoch exchange of the replicated item table (DESIGN.md "Multi-GPU"):
//   peer window  every rank places its item table / feature parameters / touch histogram in ONE cudaMalloc'ed window,
//                the windows are mapped into each other's address space (cudaIpc*, handles exchanged once over NCCL) and a
//                single kernel per epoch does delta -> gain-weighted sum over ranks -> apply, reading and writing peer
//                memory over NVLink/NVSwitch directly (rfm_comm.cu: exchange_kernel).  No NCCL call on the data path.
//   NCCL         fallback when peer mapping is not possible (devices not peer-accessible, world > 8, several live
//                sessions on one communicator): delta kernel + ncclAllReduce + apply kernel.
// ---------------------------------------------------------------------------------------------------------------
struct Comm;

struct ExchangeShape {
    int32_t I, ldi, NQ, Fp;
    size_t gp_floats;
};

int comm_acquire(const uint8_t* id128, int rank, int world, int device, Comm** out);    // collective on first use
void comm_release(Comm* c);                                                              // drops one reference, keeps the communicator cached
int comm_world(const Comm* c);

// Place a session's replicated tables in the communicator's peer window (collective).  On success *p2p tells whether the
// fused peer-memory exchange is available; if so IT / GP / touch point INTO the window (owned by the communicator, never
// to be freed by the session), otherwise they are left untouched and the session allocates its own.
int comm_window_attach(Comm* c, const ExchangeShape& shape, cudaStream_t st, const void* owner, float** IT, float** GP, int32_t** touch, bool* p2p);
void comm_window_detach(Comm* c, const void* owner);

// fused exchange of one epoch: table <- snapshot + sum_r gain_r(row) * (table_r - snapshot) on every rank, snapshot <- table
int comm_exchange_p2p(Comm* c, cudaStream_t st, float* snap_it, float* snap_gp, const rfm::EpochAcc* acc, float lam_factor, float lam_bias, float gp_gain);
int comm_check(Comm* c, cudaStream_t st);        // after the stream drained: did a peer barrier time out?

// NCCL collectives (fallback data path, end-of-training statistics)
int comm_allreduce_f32(Comm* c, float* buf, size_t n, cudaStream_t st);
int comm_allreduce_f64(Comm* c, double* buf, size_t n, cudaStream_t st);

int nccl_unique_id(uint8_t* out128);

}  // namespace rfmh
