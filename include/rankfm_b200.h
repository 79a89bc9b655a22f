/*
 * rankfm_b200.h -- C ABI of librankfm_b200.so: the B200 (sm_100a) replacement for RankFM's native hot path.
 *
 * The reference's native boundary is the Cython module `rankfm/_rankfm.pyx`, imported at `rankfm/rankfm.py:8`
 * (`from rankfm._rankfm import _fit, _predict, _recommend`).  Each entry point below names the reference
 * interface it replaces.  Plain pointers and sizes only: no Python objects, no torch types.  All pointers are
 * HOST pointers unless a name ends in `_dev`.  Every function returns 0 on success or an RFM_ERR_* code;
 * `rfm_last_error()` then holds a human-readable message (thread-local).
 *
 * There is no CPU fallback: with no CUDA device the calls fail with RFM_ERR_NO_DEVICE.
 *
 * Layout conventions (identical to what `rankfm.py:140-244` hands to the Cython functions):
 *   interactions int32 [N,2] C-contiguous (user_idx, item_idx); sample_weight f32 [N];
 *   user_items as CSR: csr_indptr int64 [U+1], csr_indices int32 [nnz], each user's items sorted ascending;
 *   x_uf f32 [U,P], x_if f32 [I,Q] (P=Q=1 all-zero when absent, rankfm.py:199,211);
 *   w_i f32 [I], w_if f32 [Q], v_u f32 [U,F], v_i f32 [I,F], v_uf f32 [P,F], v_if f32 [Q,F].
 */
#ifndef RANKFM_B200_H
#define RANKFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFM_OK              0
#define RFM_ERR_ARG         1   /* bad argument (message says which) */
#define RFM_ERR_CUDA        2   /* CUDA runtime error */
#define RFM_ERR_NO_DEVICE   3   /* no usable CUDA device: this library has no CPU path */
#define RFM_ERR_NCCL        4   /* NCCL missing or failed */
#define RFM_ERR_NONFINITE   5   /* weights went non-finite (reference: AssertionError from assert_finite) */
#define RFM_ERR_UNSUPPORTED 6

/* learning-rate schedules, `_rankfm.pyx:220-225` */
#define RFM_SCHEDULE_CONSTANT   0
#define RFM_SCHEDULE_INVSCALING 1

/* row order inside an epoch */
#define RFM_ORDER_FEISTEL 0     /* on-device keyed permutation (production) */
#define RFM_ORDER_HOST    1     /* caller supplies the permutation of each epoch (what np.random.shuffle produced,
                                   `_rankfm.pyx:227`) */
/* negative sampler */
#define RFM_SAMPLER_PHILOX 0    /* counter-based, any schedule (production) */
#define RFM_SAMPLER_MT     1    /* MT19937 stream of the reference (`_rankfm.pyx:182,251`); serial schedule only */
/* schedule */
#define RFM_SCHED_PARALLEL 0    /* Hogwild: one lane-group per positive, whole GPU */
#define RFM_SCHED_SERIAL   1    /* one lane-group, positives strictly in order: reproduces sequential SGD */

typedef struct rfm_session rfm_session;

/* Model + data description.  Mirrors the positional arguments of `_fit` (`_rankfm.pyx:122-142`). */
typedef struct rfm_problem {
    /* interaction data (may be NULL / 0 for a scoring-only session) */
    const int32_t *interactions;
    const float   *sample_weight;
    int64_t        n_interactions;
    const int64_t *csr_indptr;       /* user_items, [U+1] */
    const int32_t *csr_indices;      /* [csr_indptr[U]] */
    /* side features */
    const float *x_uf;               /* [U,P] */
    const float *x_if;               /* [I,Q] */
    /* weights: read at session creation, written back by rfm_session_download / rfm_fit */
    float *w_i, *w_if, *v_u, *v_i, *v_uf, *v_if;
    int32_t U, I, P, Q, F;
    /* hyper-parameters */
    float   alpha, beta, learning_rate, learning_exponent;
    int32_t schedule;                /* RFM_SCHEDULE_* */
    int32_t max_samples;             /* 1 for BPR (`rankfm.py:294-295`) */
    /* execution */
    int32_t  order;                  /* RFM_ORDER_* */
    int32_t  sampler;                /* RFM_SAMPLER_* */
    int32_t  sched;                  /* RFM_SCHED_* */
    uint32_t mt_seed;                /* 1492 in the reference */
    uint64_t seed;                   /* Philox / Feistel key */
    int32_t  max_rejects;            /* rejection-sampling bound per draw (reference: unbounded); 0 -> 64 */
    int32_t  device;                 /* CUDA device ordinal */
    /* multi-GPU (one process per GPU).  world==1: single GPU, nccl_id ignored. */
    int32_t  rank, world;
    const uint8_t *nccl_id;          /* 128 bytes from rfm_nccl_unique_id() on rank 0, broadcast by the caller */
    /* multi-GPU partition: this rank owns users [user_lo, user_hi) -- only their rows of v_u / x_uf are
     * uploaded, trained and written back (the host arrays keep their global [U, .] shape; rows of other users are never
     * touched) and every interaction passed to this rank must belong to one of them.  0,0 = all users (world == 1). */
    int32_t  user_lo, user_hi;
    /* epochs this model has already been trained for by EARLIER calls: offsets the Philox / Feistel keys so that a
     * sequence of warm-start calls (`fit_partial`, rankfm.py:269-327) does not replay the same order and negatives */
    int32_t  epoch_offset;
} rfm_problem;

/* Per-epoch report, the device-side equivalent of `_rankfm.pyx:328-336` (assert_finite, reg_penalty, log-lik). */
typedef struct rfm_epoch_stats {
    double  log_likelihood;          /* sum of log sigma(pairwise utility), measured before each update */
    double  penalty;                 /* alpha*sum(w_i^2,v_u^2,v_i^2) + beta*sum(w_if^2,v_uf^2,v_if^2) */
    int64_t draws;                   /* negatives evaluated (sum of `sampled`) */
    int32_t finite[6];               /* w_i, w_if, v_u, v_i, v_uf, v_if: 1 = finite */
    float   eta;                     /* learning rate used */
    float   kernel_ms;               /* CUDA-event time of the SGD kernel launch(es) of this epoch */
    float   sync_ms;                 /* CUDA-event time of the multi-GPU delta exchange (0 when world==1) */
} rfm_epoch_stats;
/* With world > 1 log_likelihood, draws, penalty and finite[] are those of the WHOLE job (summed over the ranks' shards
 * after the last epoch), like the single `ll` the reference prints (`_rankfm.pyx:332-336`). */

/* ---- library / device ---- */
const char *rfm_version(void);
const char *rfm_last_error(void);
int rfm_device_count(void);                                      /* number of CUDA devices, 0 if none */
int rfm_nccl_unique_id(uint8_t *out128);                         /* rank 0 calls, caller broadcasts */
/* Multi-GPU communicators (NCCL communicator + the peer-memory window the fused delta exchange runs over) are created by
 * the first session that names a (nccl_id, rank, world, device) and are KEPT by the library for later sessions of the same
 * job -- a one-shot `rfm_fit` per `fit_partial` call pays ncclCommInitRank + the IPC handshake once.  This call destroys
 * every cached communicator that no live session uses (collective: every rank of the job must call it). */
int rfm_comm_release_all(void);

/* page-lock / unlock a caller-owned host buffer (cudaHostRegister) so that the one-shot calls below copy it at full PCIe
 * speed; optional -- every entry point also accepts pageable memory */
int rfm_host_register(void *ptr, uint64_t bytes);
int rfm_host_unregister(void *ptr);

/* host-side evaluation of the kernels' RNG definitions (csrc/rfm_rng.cuh), so CPU tests can hold them to the
 * contract shared with the oracle: Philox4x32-10 block and the per-epoch Feistel permutation of [0,n) */
int rfm_debug_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out4);
int rfm_debug_feistel(int64_t n, uint64_t seed, int32_t epoch, int64_t r0, int64_t count, int64_t *out);
/* device self-test of the production kernel's side-feature code specialised for <= 8 + 8 feature columns (csrc/rfm_feat8.cuh)
 * against the generic code, on pseudo-random rows: out5 = max |difference| of the hoisted user vectors a[], b[]
 * (`_rankfm.pyx:67-87`), of the feature parameters after one gradient step (`:283-326`), of the row deltas, and the largest
 * parameter movement of that step (so a caller can see the step was not a no-op).  P or Q = 0: that block is inactive. */
int rfm_debug_feat8(int32_t F, int32_t P, int32_t Q, uint32_t seed, float *out5);
/* host arithmetic of the estimated row threshold of the tensor-core recommend path (`RANKFM_B200_TAU_MODE=estimate`): which
 * block bound of a 1-in-k sample of the item tiles estimates the n'-th best score of the catalogue with z sigma of head room
 * (z = 0 or k = 1: the n'-th itself).  No GPU needed. */
int rfm_debug_tau_rank(int32_t want, int32_t sample_k, float z);

/* ---- data preparation on the device: what `RankFM._init_all` / `_init_interactions` do with pandas on
 * the host (`rankfm.py:114-177`) and `_fit` with a Python loop (`_rankfm.pyx:201-212`), as device radix sorts ---- */

/* sorted unique ids and the index of every id in that list: `np.unique` + the `user_to_index` / `item_to_index` Series maps
 * (`rankfm.py:114-128,150-155`) for integer ids.  unique_out must have room for n values. */
int rfm_prep_index_ids(const int64_t *ids, int64_t n, int32_t device, int64_t *unique_out, int64_t *n_unique_out, int32_t *index_out);
/* `user_items` (`rankfm.py:165-174`: {user_idx: sorted item_idx array}, duplicates kept) as CSR: indptr_out [U+1],
 * indices_out [n], each user's items ascending */
int rfm_prep_user_items(const int32_t *interactions /* [n,2] */, int64_t n, int32_t U, int32_t I, int32_t device,
                        int64_t *indptr_out, int32_t *indices_out);
/* bench tooling: synthetic interactions of the BASELINE.json shapes: user ~ Zipf(a_u) over U ranks, item ~
 * Zipf(a_i) over I ranks, de-duplicated, trimmed to N pairs in random order, ids randomly permuted (perm_seed: ranks of a
 * multi-GPU job share the item popularity ranking) and, if `reindex`, re-indexed to the observed uniques like
 * `rankfm.py:115-116`.  out int32 [N,2] (user index + user_offset, item index); n_out <= N pairs were written;
 * users_out / items_out = number of distinct users / items observed. */
int rfm_synth_zipf(int32_t U, int32_t I, int64_t N, double a_u, double a_i, uint64_t seed, uint64_t perm_seed, int32_t reindex,
                   int32_t user_offset, int32_t device, int32_t *out, int64_t *n_out, int32_t *users_out, int32_t *items_out);

/* ---- one-shot calls on host buffers: the drop-in boundary ---- */

/* replaces `_fit` (`_rankfm.pyx:122-342`): trains `epochs` epochs, updates the six weight arrays in place.
 * `perms` = int32 [epochs,N] when order==RFM_ORDER_HOST, else NULL.  `stats` = [epochs] or NULL. */
int rfm_fit(const rfm_problem *p, int32_t epochs, const int32_t *perms, rfm_epoch_stats *stats);

/* wall-clock phases of this thread's last rfm_fit call, milliseconds: { session create + H2D, training, D2H, destroy }
 * (bench.py reports them next to the end-to-end time of the stateless plug-in call) */
int rfm_last_fit_phases(double *out4);

/* replaces `_predict` (`_rankfm.pyx:345-390`): pairs f32 [n,2] hold indexes as floats, NaN = unknown id */
int rfm_predict(const rfm_problem *p, const float *pairs, int64_t n, float *scores);

/* replaces `_recommend` (`_rankfm.pyx:393-460`): users f32 [n_users] (NaN = unknown), rec_items f32
 * [n_users,n_items] receives item indexes as floats; rows of unknown users are NaN.  When fewer than n_items
 * candidates survive `filter_previous` the tail is NaN (the reference leaves it uninitialised). */
int rfm_recommend(const rfm_problem *p, const float *users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                  float *rec_items);

/* similar_items / similar_users (`rankfm.py:405-454`): top-n rows of (v + x.v_f) by inner product with row
 * `index`, the query row itself excluded.  which = 0 items, 1 users.  out int32 [n]. */
int rfm_similar(const rfm_problem *p, int32_t which, int32_t index, int32_t n, int32_t *out);

/* ---- resident sessions: upload once, train / score many times (bench.py, RankFM class) ---- */
int rfm_session_create(const rfm_problem *p, rfm_session **out);             /* allocates HBM, copies H2D */
int rfm_session_train(rfm_session *s, int32_t epochs, const int32_t *perms, rfm_epoch_stats *stats);
/* called on the training thread after every epoch of rfm_session_train with that epoch's record (the reference prints the
 * log-likelihood of every epoch as it completes when verbose, `_rankfm.pyx:332-336`); costs one stream synchronisation per
 * epoch, so it is only installed by verbose callers.  NULL removes it. */
typedef void (*rfm_epoch_callback)(int32_t epoch, const rfm_epoch_stats *stats, void *user);
int rfm_session_set_epoch_callback(rfm_session *s, rfm_epoch_callback cb, void *user);
int rfm_session_set_weights(rfm_session *s, const float *w_i, const float *w_if, const float *v_u, const float *v_i,
                            const float *v_uf, const float *v_if);           /* H2D of the weights only */
/* device-side copy of the current weights, and restoring it (D2D only): lets a benchmark restart training from
 * identical weights without a host round trip */
int rfm_session_snapshot(rfm_session *s);
int rfm_session_restore(rfm_session *s);
/* CUDA events on the session's stream: start .. stop brackets whatever was enqueued in between */
int rfm_session_timer_start(rfm_session *s);
int rfm_session_timer_stop(rfm_session *s, float *ms_out);
int rfm_session_download(rfm_session *s, float *w_i, float *w_if, float *v_u, float *v_i, float *v_uf, float *v_if);
int rfm_session_predict(rfm_session *s, const float *pairs, int64_t n, float *scores);
int rfm_session_recommend(rfm_session *s, const float *users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                          float *rec_items);
/* device-resident timing probes used by bench.py: run the op `iters` times on inputs already in HBM and return
 * the mean CUDA-event milliseconds per iteration (no host<->device copies inside the timed region) */
int rfm_session_time_predict(rfm_session *s, const float *pairs, int64_t n, int32_t iters, float *ms_out);
int rfm_session_time_recommend(rfm_session *s, const float *users, int64_t n_users, int32_t n_items, int32_t filter_previous,
                               int32_t iters, float *ms_out, float *gemm_ms_out);
/* debugging / parity: record, for every position r of an epoch, the negative item the sampler settled on and the
 * number of draws it used (`min_index`, `sampled` of `_rankfm.pyx:244-268`); read back the last epoch's record */
int rfm_session_trace_enable(rfm_session *s);
int rfm_session_trace_read(rfm_session *s, int32_t *neg_and_sampled /* [N,2] */);
/* debugging / parity: dense bf16 tensor-core scores (A.B^T + bias) of the requested users against all items */
int rfm_session_debug_gemm(rfm_session *s, const float *users, int64_t n_users, float *scores_out /* [n_users, I] */);
/* `rfm_similar` on a resident session (same arguments and result) */
int rfm_session_similar(rfm_session *s, int32_t which, int32_t index, int32_t n, int32_t *out);
/* many queries in one call (the batched all-items variant): out int32 [n_queries, n], -1 = no such row */
int rfm_session_similar_batch(rfm_session *s, int32_t which, const int32_t *indexes, int64_t n_queries, int32_t n, int32_t *out);
/* hold-out ranking metrics (`rankfm/evaluation.py:9-143`) computed on the device from the top-k of `users` (float32 indexes,
 * all known) against each user's test items: test_indptr int64 [n_users+1] / test_items int32 (item INDEXES, ascending per
 * user, unknown items left out) in the order of `users`; n_test [n_users] = the user's number of distinct test items
 * (unknown ones included: the recall denominator).  out5 = mean over the users of { hit, reciprocal rank, DCG, precision,
 * recall }; hits_out (optional) uint8 [n_users, k] = 1 where the r-th recommendation of the user is a test item. */
int rfm_session_evaluate(rfm_session *s, const float *users, int64_t n_users, int32_t k, int32_t filter_previous,
                         const int64_t *test_indptr, const int32_t *test_items, const int32_t *n_test, double *out5, uint8_t *hits_out);
/* (re)attach the user_items CSR a scoring session filters with (`filter_previous`, `_rankfm.pyx:450`): lets a caller keep
 * one session resident across `_predict` / `_recommend` calls instead of re-uploading the weights every call */
int rfm_session_attach_csr(rfm_session *s, const int64_t *csr_indptr /* [U+1] */, const int32_t *csr_indices /* [nnz] */);
/* tensor-core recommend bookkeeping since session creation: rows served by the tcgen05 path, and how many of those had to
 * be redone on the exact fp32 path because a candidate slot overflowed */
int rfm_session_recommend_stats(rfm_session *s, int64_t *tc_rows, int64_t *tc_redone);
/* ... and how many rows were served a second time with the provable row threshold because the estimated one (taken from a
 * 1-in-k sample of the item tiles, see rfm_api.cu "ESTIMATED") came out too high for them; results never depend on it */
int rfm_session_recommend_retried(rfm_session *s, int64_t *tc_retried);
/* the library keeps freed device blocks for the next call of the same shape (one-shot calls create and destroy a session
 * each; see rfm_api.cu "Device block cache"; limit RANKFM_B200_CACHE_MB, default 16384): give them back to the driver */
int rfm_trim_device_cache(void);
int rfm_session_flush_l2(rfm_session *s);                                    /* overwrite a >L2-sized scratch buffer */
int rfm_session_launch_count(rfm_session *s, int64_t *launches);             /* kernels launched by this session */
/* multi-GPU: which path the per-epoch item-delta exchange of this session takes: 0 = none (world == 1), 1 = fused
 * peer-memory kernel over NVLink (cudaIpc-mapped windows, no NCCL on the data path), 2 = ncclAllReduce (fallback) */
int rfm_session_exchange_path(rfm_session *s, int32_t *path);
int rfm_session_destroy(rfm_session *s);

#ifdef __cplusplus
}
#endif
#endif /* RANKFM_B200_H */
