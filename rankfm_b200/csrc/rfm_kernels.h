// rfm_kernels.h -- host-visible declarations of the kernel launchers (internal; the public ABI is include/rankfm_b200.h)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rfm_common.cuh"
#include "rfm_rng.cuh"

namespace rfm {

constexpr int kTrainThreads = 256;

struct EpochAcc {            // zeroed before each epoch, read back after
    double ll;
    long long draws;
    int bad;
    int pad;
    double wstats[12];       // weight_stats_kernel output
};

struct TrainParams {
    Tables T;
    const int2* interactions;   // [N] (user, item); N < 2^31
    const float* sample_weight;
    const int64_t* indptr;
    const int32_t* indices;
    const uint32_t* bitmap;  // optional [U, bitmap_words]: bit i of row u set <=> item i in user_items[u]; nullptr -> search the CSR
    int32_t bitmap_words;
    // optional pre-filter of the CSR membership search when there is no bitmap: one 32-bit word per CSR entry, aligned
    // with `indices` (user u owns the words [indptr[u], indptr[u+1]) = 32*deg bits); bit bloom_slot(item, deg) is set for
    // every observed item.  A clear bit proves "not observed" (the common case, ~97 % of the candidates) with ONE load;
    // a set bit is verified by the exact search, so the sampler's results are unchanged.
    const uint32_t* bloom;
    const int32_t* perm;     // this epoch's order, or nullptr -> Feistel
    const float* mult;       // [max_samples+1]  WARP multiplier by number of draws
    Feistel feistel;
    long long N;
    float eta, reg_a, reg_b;
    int32_t max_samples, max_rejects, serial;
    int32_t depth;           // stages of the TMA row pipeline (set by the launcher)
    int32_t spec;            // WARP sampler look-ahead: 1, 2, 4 attempts per round; 0 = choose from prev_acc
    const EpochAcc* prev_acc;   // previous epoch's record (device), or nullptr
    float* gp_acc;           // [gp_floats] accumulator of the warps' feature-parameter deltas (production, FEAT)
    float gp_gain;           // weight of each warp's delta when the chains are folded (see rfm_session_train)
    int32_t gp_floats;
    int32_t gp_private;      // 1: one feature-parameter chain per lane group (plain RMW), 0: one per warp (atomics)
    int32_t gp_race;         // one chain per warp, one update lands per step: 1 = the lane groups' plain stores race, 2 = the warp applies the winner group's update with all lanes
    uint32_t k0, k1, epoch_key;
    MtState* mt;             // non-null -> MT19937 sampler (serial only)
    EpochAcc* acc;
    int32_t* trace;          // optional [N,2]: (chosen negative, draws used) per position of the epoch
};

int train_group_size(const Tables& T, int* qpl_out);
cudaError_t launch_sgd_epoch(const TrainParams& p, int grid, cudaStream_t st);
cudaError_t launch_build_bloom(const int64_t* indptr, const int32_t* indices, int U, uint32_t* bloom, cudaStream_t st);
cudaError_t launch_build_bitmap(const int64_t* indptr, const int32_t* indices, int U, uint32_t* bitmap, int words, cudaStream_t st);
cudaError_t launch_gp_apply(float* gp, float* acc, int n, cudaStream_t st);
size_t sgd_pipe_smem_bytes(const Tables& T);
int sgd_pipe_chains_per_warp(const Tables& T);
int sgd_pipe_groups_per_chain(const Tables& T);
// self-test of the P, Q <= 8 specialisation of the side-feature math (rfm_feat8.cuh) against the generic code: out5 =
// max |difference| of a[], b[], the chain copy after one step, the row deltas; and the largest chain movement (non-zero)
cudaError_t launch_feat8_selftest(const Tables& T, uint32_t seed, float eta, float reg_b, float* out5, cudaStream_t st);
cudaError_t launch_weight_stats(const Tables& T, double* out12, int grid, cudaStream_t st);

// scoring (rfm_score.cu)
cudaError_t launch_predict(const Tables& T, const float2* pairs, long long n, float* scores, int grid, cudaStream_t st);

cudaError_t launch_score_users(const Tables& T, const int32_t* users, int n_users, float* S, int chunks, cudaStream_t st);
cudaError_t launch_topn_select(float* S, int I, const int32_t* users, int n_users, const int64_t* indptr, const int32_t* indices,
                               int filter_previous, int n_items, float* rec, const int32_t* exclude, cudaStream_t st,
                               const int32_t* idxmap = nullptr);

// tensor-core candidate generation for recommend (rfm_gemm.cu)
int gemm_kp(const Tables& T);
int gemm_block_n(const Tables& T);
int gemm_m_tile(const Tables& T);             // user rows per CTA (128 x sub-tiles); M_pad must be a multiple
int gemm_slots_per_split(const Tables& T);    // candidate slots per (row, item split)
bool gemm_supported(const Tables& T);
cudaError_t launch_pack_gemm_items(const Tables& T, int Kp, int I_pad, void* B, float* bias, int32_t* order, cudaStream_t st);
cudaError_t launch_pack_gemm_users(const Tables& T, const int32_t* users, int n_users, int M_pad, int Kp, void* A, cudaStream_t st);
cudaError_t launch_score_filter(const Tables& T, int mode, const void* A, const void* B, const float* bias, int n_users, int M_pad, int I_pad, int n_splits,
                                int tile_stride, float2* cand, int* cand_cnt, const float* tau, int cap, float* rowmax, float* S, cudaStream_t st);
// Rank of the block bound that becomes the row threshold when pass 1 saw a 1-in-k sample of the item tiles.
//   z == 0, k == 1: the n'-th largest bound -- a PROVABLE lower bound of the row's n'-th best score (the "safe" threshold).
//   z  > 0:         the m-th largest bound of the sample, m ~ n'/k + z-sigma head room: an ESTIMATE of the n'-th best score
//                   of the whole catalogue.  With m sample items above tau the catalogue holds ~ k m of them, standard
//                   deviation ~ k sqrt(m) (negative binomial); m solves k m - z k sqrt(m) = n'.  The estimate can be
//                   too high; shortlist_kernel VERIFIES it (>= n' candidates reached tau) and flags the row otherwise,
//                   so results never depend on it.
__host__ __device__ inline int tau_rank(int want, int k, float z)
{
    if (k <= 1 || !(z > 0.f)) return want;
    const float r = 0.5f * (z + sqrtf(z * z + 4.f * (float)want / (float)k));
    const int m = (int)ceilf(r * r) + 1;
    return m < want ? m : want;
}
cudaError_t launch_row_threshold(const float* rowmax, int n_rows, int n_blocks, const int* n_target, float* tau, int sample_k, float z, cudaStream_t st);
constexpr int kTauBlock = 8;                   // items per pass-1 block bound
constexpr int kShortWidth = 512;               // shortlist entries per row, narrow tier (n' <= 256 plus ties at the cut)
constexpr int kShortWidthWide = 2048;          // wide tier (n' <= 1024: users with long histories under filter_previous)
cudaError_t launch_shortlist(const Tables& T, const int32_t* users, int n_users, const float2* cand, const int* cnt, int slots, int cap, const float* bias,
                             const int32_t* order, const int* n_target, const int64_t* indptr, const int32_t* indices, int filt, int n_items, float* rec,
                             int* flag, const float* tau, int I_pad, int short_width, int stage_cap /* candidates staged per row; 0 = every slot's capacity */, cudaStream_t st);
cudaError_t launch_eval_topk(const float* rec, const int64_t* order, int n_users, int k, const int64_t* test_indptr, const int32_t* test_items,
                             const int32_t* n_test, double* out5, uint8_t* hits_out, cudaStream_t st);
cudaError_t launch_scatter_rows(const float* src, const int64_t* order, long long n_rows, int n_items, float* dst, cudaStream_t st);
cudaError_t launch_latent_scores(const Tables& T, int which, int index, float* qvec, float* S, cudaStream_t st);
int sgd_epoch_blocks_per_sm(const TrainParams& p);

// packing between the reference's array layout and the fat-row tables (rfm_pack.cu)
cudaError_t launch_pack_users(const Tables& T, const float* v_u, const float* x_uf, cudaStream_t st);
cudaError_t launch_pack_items(const Tables& T, const float* v_i, const float* w_i, const float* x_if, cudaStream_t st);
cudaError_t launch_unpack_users(const Tables& T, float* v_u, cudaStream_t st);
cudaError_t launch_unpack_items(const Tables& T, float* v_i, float* w_i, cudaStream_t st);
cudaError_t launch_pack_globals(const Tables& T, const float* w_if, const float* v_uf, const float* v_if, int n_total, cudaStream_t st);
cudaError_t launch_unpack_globals(const Tables& T, float* w_if, float* v_uf, float* v_if, cudaStream_t st);

}  // namespace rfm
