// rfm_sgd.cuh -- the two halves of one SGD step shared by both epoch kernels (rfm_train.cu):
//   sample_negatives  WARP/BPR rejection sampler  (rankfm/_rankfm.pyx:244-264)
//   apply_update      gradient step               (rankfm/_rankfm.pyx:267-326)
// Both are written for lane groups (rfm_common.cuh); every lane of the warp must call them (warp-uniform control flow,
// *_sync primitives are never reached through divergent branches).
#pragma once
#include "rfm_kernels.h"
#include "rfm_pair.cuh"
#include "rfm_feat8.cuh"
#include "rfm_rng.cuh"

namespace rfm {

struct StepAcc {
    double ll = 0.0;          // sum log sigma(pairwise utility)
    float ll_f = 0.f;         // short-run float partial of the same sum, folded into `ll` by fold()
    int draws = 0;            // negatives evaluated (folded into draws64 by fold())
    long long draws64 = 0;
    int bad = 0;              // a positive whose pairwise utilities were all NaN
    __device__ __forceinline__ void fold() { ll += (double)ll_f; ll_f = 0.f; draws64 += draws; draws = 0; }
};

// is item `cj` in user u's observed set?  bitmap (small catalogues) or (G+1)-ary search of the sorted CSR segment
template <int G>
__device__ __forceinline__ bool is_member(const TrainParams& p, int u, int cj, long long seg, int deg, bool active, int sub, int gw)
{
    if (p.bitmap) return active && ((__ldg(p.bitmap + (size_t)u * p.bitmap_words + (cj >> 5)) >> (cj & 31)) & 1u);
    return group_member<G>(cj, p.indices + seg, deg, active, sub, gw);
}

// Draws s = s_begin .. max_samples for every group that is not `done`, tracking the hardest negative seen so far
// (min pairwise utility) in (neg, min_pu, min_j, sampled) exactly like the reference's loop.  `attempt` counts Philox
// words consumed for this positive (rejected ones included), so a caller that already made draw 1 can continue.
template <int G, int QPL, bool FEAT, bool MT>
__device__ __forceinline__ void sample_negatives(const TrainParams& p, const UserCtx<QPL>& uc, float ut_ui, int row, int u, long long seg, int deg,
                                                 int s_begin, bool& done, uint32_t& attempt, MtState* mt, int sub, int gw,
                                                 ItemRow<QPL>& neg, float& min_pu, int& min_j, int& sampled)
{
    const Tables& T = p.T;
    ItemRow<QPL> cand;
    Philox4 blk = {0u, 0u, 0u, 0u};
    if (!MT && (attempt & 3u) != 0u) blk = philox4x32_10((uint32_t)row, p.epoch_key, attempt >> 2, 0u, p.k0, p.k1);
    for (int s = s_begin; s <= p.max_samples; ++s) {
        if (!__any_sync(0xffffffffu, !done)) break;
        // rejection-sample an unobserved item: `while True: j = genrand_int32() % I` (:250-253)
        int j = 0, rejects = 0;
        bool need = !done;
        while (__any_sync(0xffffffffu, need)) {
            uint32_t word;
            if (MT) {
                word = mt_next_warp(mt);
            } else {
                if ((attempt & 3u) == 0u) blk = philox4x32_10((uint32_t)row, p.epoch_key, attempt >> 2, 0u, p.k0, p.k1);
                const uint32_t c = attempt & 3u;
                word = c == 0u ? blk.x : (c == 1u ? blk.y : (c == 2u ? blk.z : blk.w));
                if (need) ++attempt;
            }
            // MT replay: `genrand_int32() % I` like the reference; Philox: multiply-shift range reduction (1 IMAD.HI)
            const int cj = MT ? (int)(word % (uint32_t)T.I) : (int)__umulhi(word, (uint32_t)T.I);
            if (need) load_item<G, QPL, FEAT>(T, cj, true, sub, cand);      // speculative: in flight during the membership test
            const bool member = is_member<G>(p, u, cj, seg, deg, need, sub, gw);
            if (need && (!member || ++rejects >= p.max_rejects)) { j = cj; need = false; }
        }
        const float pu = ut_ui - utility<G, QPL, FEAT>(uc, cand);
        if (!done) {
            sampled = s;
            if (pu < min_pu) { min_pu = pu; min_j = j; neg = cand; }
            if (pu < 1.0f) done = true;                                      // MARGIN (:149,263)
        }
    }
}

// Philox-only variant of the loop above with speculative look-ahead: up to `spec` (1, 2 or 4) consecutive attempts of
// one Philox block are resolved together -- their membership words and candidate rows are all in flight at once, then
// they are consumed strictly in attempt order, so the accepted draws, their count and the arg-min are exactly those of
// the sequential loop (attempts past the terminating draw are simply discarded, like unread words of the block).
// Cuts the dependent-latency chain of a WARP positive from one L2/HBM round trip per draw to one per `spec` draws.
// `s` = draws already made for this positive (the caller made draw 1 from the prefetched candidate).
template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void sample_negatives_spec(const TrainParams& p, const UserCtx<QPL>& uc, float ut_ui, int row, int u, long long seg, int deg,
                                                      int s, bool& done, uint32_t& attempt, int spec, int sub, int gw,
                                                      ItemRow<QPL>& neg, float& min_pu, int& min_j, int& sampled)
{
    constexpr int KMAX = (QPL == 1) ? 4 : 2;
    const Tables& T = p.T;
    int rejects = 0;
    // CSR membership without a memory access per candidate: a user with at most 4*G observed items is held entirely
    // in the registers of its lane group (lane `sub` keeps items sub, sub+G, sub+2G, sub+3G); longer lists are searched
    constexpr int OWN = 4;
    int own[OWN];
    const bool listed = !p.bitmap && deg <= OWN * G;
#pragma unroll
    for (int k = 0; k < OWN; ++k) own[k] = (!done && listed && sub + k * G < deg) ? __ldg(p.indices + seg + sub + k * G) : -1;
    while (__any_sync(0xffffffffu, !done)) {
        const Philox4 blk = philox4x32_10((uint32_t)row, p.epoch_key, attempt >> 2, 0u, p.k0, p.k1);
        const int off = (int)(attempt & 3u);
        const int navail = min(4 - off, spec);                     // attempts this group resolves in this round
        ItemRow<QPL> cand[KMAX];
        int cj[KMAX];
        uint32_t mword[KMAX];
#pragma unroll
        for (int w = 0; w < KMAX; ++w) {
            if (w >= spec) break;                                  // warp-uniform
            const bool live = !done && w < navail;
            const int c = off + w;
            const uint32_t word = c == 0 ? blk.x : (c == 1 ? blk.y : (c == 2 ? blk.z : blk.w));
            cj[w] = (int)__umulhi(word, (uint32_t)T.I);
            load_item<G, QPL, FEAT>(T, cj[w], live, sub, cand[w]);
            mword[w] = (live && p.bitmap) ? __ldg(p.bitmap + (size_t)u * p.bitmap_words + (cj[w] >> 5)) : 0u;
            if (!p.bitmap && p.bloom) {                        // filter word of the candidate (users held in registers need none)
                const uint32_t h = bloom_slot(cj[w], deg);
                mword[w] = (live && !listed && deg > 0) ? ((__ldg(p.bloom + seg + (h >> 5)) >> (h & 31)) & 1u) : 1u;
            }
        }
#pragma unroll
        for (int w = 0; w < KMAX; ++w) {
            if (w >= spec) break;
            const bool live = !done && w < navail;
            bool member;
            if (p.bitmap) member = ((mword[w] >> (cj[w] & 31)) & 1u) != 0u;
            else {
                const int c = cj[w];
                const bool hit = group_ballot<G>(own[0] == c || own[1] == c || own[2] == c || own[3] == c, gw) != 0u;
                // without a filter every candidate is searched; with one only those whose bit is set (mword = 1)
                const bool searched = group_member<G>(c, p.indices + seg, deg, live && !listed && (!p.bloom || mword[w] != 0u), sub, gw);
                member = listed ? hit : searched;
            }
            const float pu = ut_ui - utility<G, QPL, FEAT>(uc, cand[w]);
            if (live) {
                ++attempt;
                if (member && ++rejects < p.max_rejects) {
                    // observed item: rejected, the next attempt belongs to the same draw
                } else {
                    rejects = 0;
                    sampled = ++s;
                    if (pu < min_pu) { min_pu = pu; min_j = cj[w]; neg = cand[w]; }
                    if (pu < 1.0f || s >= p.max_samples) done = true;          // MARGIN (:149,263) / loop bound (:247)
                }
            }
        }
    }
}

// Pipelined variant of sample_negatives_spec (same results, attempt for attempt):
//   * the Philox words of a positive are computed ONCE, one block per lane of the group (lane `sub` holds block b0 + sub:
//     4 G consecutive attempts), and fetched with a shuffle -- the look-ahead loop above recomputes the same block on
//     every lane in every round;
//   * attempts are resolved in batches of K, and the NEXT batch's candidate rows / membership words are already in flight
//     while the current batch is consumed (two batches = the same 4 candidate rows in registers as one round of the
//     look-ahead loop), so a positive that needs several draws pays one L2/HBM round trip plus its arithmetic instead of
//     one round trip per round.  Attempts issued past the terminating draw are discarded, like unread words of a block.
template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void sample_negatives_pipe(const TrainParams& p, const UserCtx<QPL>& uc, float ut_ui, int row, int u, long long seg, int deg,
                                                      int s, bool& done, uint32_t& attempt, int spec, int sub, int gw,
                                                      ItemRow<QPL>& neg, float& min_pu, int& min_j, int& sampled)
{
    constexpr int K = (QPL == 1) ? 2 : 1;
    const Tables& T = p.T;
    int rejects = 0;
    constexpr int OWN = 4;
    int own[OWN];
    const bool listed = !p.bitmap && deg <= OWN * G;
#pragma unroll
    for (int k = 0; k < OWN; ++k) own[k] = (!done && listed && sub + k * G < deg) ? __ldg(p.indices + seg + sub + k * G) : -1;

    uint32_t b0 = attempt >> 2;                                  // first block of the word window held by the group
    Philox4 blk = philox4x32_10((uint32_t)row, p.epoch_key, b0 + (uint32_t)sub, 0u, p.k0, p.k1);
    uint32_t a_issue = attempt;                                  // next attempt to put in flight (>= attempt, the next to consume)

    struct Batch { ItemRow<QPL> cand[K]; int cj[K]; uint32_t mw[K]; };
    auto issue = [&](Batch& B) {
        const bool refill = !done && ((a_issue + (uint32_t)(K - 1)) >> 2) - b0 >= (uint32_t)G;      // the batch leaves the window (rare)
        if (__any_sync(0xffffffffu, refill)) {
            const uint32_t nb0 = refill ? (a_issue >> 2) : b0;
            const Philox4 nb = philox4x32_10((uint32_t)row, p.epoch_key, nb0 + (uint32_t)sub, 0u, p.k0, p.k1);
            if (refill) { blk = nb; b0 = nb0; }
        }
#pragma unroll
        for (int w = 0; w < K; ++w) {
            const uint32_t a = a_issue + (uint32_t)w;
            const uint32_t c = a & 3u;
            const uint32_t mine = c == 0u ? blk.x : (c == 1u ? blk.y : (c == 2u ? blk.z : blk.w));
            const uint32_t word = __shfl_sync(0xffffffffu, mine, (int)(((a >> 2) - b0) & (uint32_t)(G - 1)), G);
            const bool live = !done;
            B.cj[w] = (int)__umulhi(word, (uint32_t)T.I);
            load_item<G, QPL, FEAT>(T, B.cj[w], live, sub, B.cand[w]);
            // membership word (bitmap word, or filter word of the candidate); stored RAW: the bit is extracted when the
            // batch is consumed, so issuing a batch never waits for its own loads
            B.mw[w] = (live && p.bitmap) ? __ldg(p.bitmap + (size_t)u * p.bitmap_words + (B.cj[w] >> 5)) : 0u;
            if (!p.bitmap && p.bloom) {
                const uint32_t h = bloom_slot(B.cj[w], deg);
                B.mw[w] = (live && !listed && deg > 0) ? __ldg(p.bloom + seg + (h >> 5)) : 0xffffffffu;
            }
        }
        a_issue += (uint32_t)K;
    };
    auto consume = [&](const Batch& B) {
#pragma unroll
        for (int w = 0; w < K; ++w) {
            const bool live = !done;
            bool member;
            if (p.bitmap) member = ((B.mw[w] >> (B.cj[w] & 31)) & 1u) != 0u;
            else {
                const int c = B.cj[w];
                const bool hit = group_ballot<G>(own[0] == c || own[1] == c || own[2] == c || own[3] == c, gw) != 0u;
                // without a filter every candidate is searched; with one only those whose bit is set
                const bool maybe = !p.bloom || ((B.mw[w] >> (bloom_slot(c, deg) & 31u)) & 1u) != 0u;
                const bool searched = group_member<G>(c, p.indices + seg, deg, live && !listed && maybe, sub, gw);
                member = listed ? hit : searched;
            }
            const float pu = ut_ui - utility<G, QPL, FEAT>(uc, B.cand[w]);
            if (live) {
                ++attempt;
                if (member && ++rejects < p.max_rejects) {
                    // observed item: rejected, the next attempt belongs to the same draw
                } else {
                    rejects = 0;
                    sampled = ++s;
                    if (pu < min_pu) { min_pu = pu; min_j = B.cj[w]; neg = B.cand[w]; }
                    if (pu < 1.0f || s >= p.max_samples) done = true;          // MARGIN (:149,263) / loop bound (:247)
                }
            }
        }
    };
    Batch A, B;
    const bool ahead = spec > 1;                                 // early epochs stop at the first draws: no look-ahead then
    issue(A);
    for (;;) {
        if (ahead) issue(B);
        consume(A);
        if (!__any_sync(0xffffffffu, !done)) break;
        if (ahead) {
            issue(A);
            consume(B);
            if (!__any_sync(0xffffffffu, !done)) break;
        } else {
            issue(A);
        }
    }
}

// sampler of the Philox schedules: the pipelined loop; -DRFM_SAMPLER_ROUNDS builds the round-based look-ahead loop instead (A/B)
template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void sample_negatives_philox(const TrainParams& p, const UserCtx<QPL>& uc, float ut_ui, int row, int u, long long seg, int deg,
                                                        int s, bool& done, uint32_t& attempt, int spec, int sub, int gw,
                                                        ItemRow<QPL>& neg, float& min_pu, int& min_j, int& sampled)
{
#ifdef RFM_SAMPLER_ROUNDS
    sample_negatives_spec<G, QPL, FEAT>(p, uc, ut_ui, row, u, seg, deg, s, done, attempt, spec, sub, gw, neg, min_pu, min_j, sampled);
#else
    sample_negatives_pipe<G, QPL, FEAT>(p, uc, ut_ui, row, u, seg, deg, s, done, attempt, spec, sub, gw, neg, min_pu, min_j, sampled);
#endif
}

// Where the row deltas of one step go.
//   RedSink   straight to HBM/L2 as per-lane vector reductions (REDG.E.ADD.F32x4)           -- serial schedule
//   SmemSink  into the shared-memory slot the rows were staged in; the caller then ships each delta row with ONE TMA
//             bulk reduction (cp.reduce.async.bulk ... .add.f32 -> UBLKRED) instead of ~33 lane reductions per row
struct RedSink {
    float *urow, *irow, *jrow;
    int Fp;
    __device__ __forceinline__ void user(int q, float4 d) const { red_add4(urow + 4 * q, d); }
    __device__ __forceinline__ void item_i(int q, float4 d) const { red_add4(irow + 4 * q, d); }
    __device__ __forceinline__ void item_j(int q, float4 d) const { red_add4(jrow + 4 * q, d); }
    __device__ __forceinline__ void bias(float di, float dj) const { red_add1(irow + Fp, di); red_add1(jrow + Fp, dj); }
};
struct SmemSink {
    float4 *su, *si, *sj;     // the tuple's three row slots in shared memory (deltas overwrite the staged rows)
    int nq;                   // Fp / 4
    __device__ __forceinline__ void user(int q, float4 d) const { su[q] = d; }
    __device__ __forceinline__ void item_i(int q, float4 d) const { si[q] = d; }
    __device__ __forceinline__ void item_j(int q, float4 d) const { sj[q] = d; }
    __device__ __forceinline__ void bias(float di, float dj) const { si[nq] = make_float4(di, 0.f, 0.f, 0.f); sj[nq] = make_float4(dj, 0.f, 0.f, 0.f); }
};

// One gradient step on (u, i, j): reads nothing but GP (feature parameters) from memory, everything else arrives in
// registers; writes go out as vector reductions.  Update formula and operand association follow the generated C of the
// reference:  w += eta * (((sw*mult) * (d_outer*d)) - (2reg * w)).
// add a delta quad to a feature parameter: HBM -> vector reduction; shared-memory chain -> plain store when the chain
// belongs to this lane group alone, shared-memory atomics only when several groups of a warp share one copy
template <int G, bool GPS>
__device__ __forceinline__ void gp_add4(float* q, const float4& w, const float4& d, bool exclusive)
{
    if (!GPS) { red_add4(q, d); return; }
    if (G == 32 || exclusive) { *reinterpret_cast<float4*>(q) = make_float4(w.x + d.x, w.y + d.y, w.z + d.z, w.w + d.w); return; }
    atomicAdd(q + 0, d.x); atomicAdd(q + 1, d.y); atomicAdd(q + 2, d.z); atomicAdd(q + 3, d.w);
}

// F8 (production kernel, P and Q <= 8, G >= 8, private chain copy): the feature loops run on the compile-time unrolled
// code of rfm_feat8.cuh with the feature values in `f8` (f8->xu filled by the caller; f8->dx is filled here)
template <int G, int QPL, bool FEAT, bool EXACT, bool GPS, typename Sink, bool F8 = false>
__device__ __forceinline__ void apply_update(const TrainParams& p, float* gp, const UserCtx<QPL>& uc, const ItemRow<QPL>& pos, const ItemRow<QPL>& neg,
                                             int min_j, float sw, int sampled, float min_pu, bool valid, long long r,
                                             int sub, StepAcc& acc, const Sink& sink, Feat8* f8 = nullptr)
{
    static_assert(!F8 || (FEAT && !EXACT && GPS), "the feat8 path belongs to the production kernel");
    const Tables& T = p.T;
    const bool upd = valid && min_j >= 0;
    if (valid && min_j < 0) acc.bad = 1;
    const float mult = upd ? __ldg(p.mult + sampled) : 0.f;                  // log((I-1)//sampled)/log(I), host table (:269)
    // d_outer = 1/(exp(pu)+1) (:276).  Serial/replay schedules keep the reference's double-precision libm path;
    // the Hogwild schedule uses the SFU (ex2 + rcp, ~1e-6 relative), far below its own scheduling noise.
    float d_outer, ll_term;
    if (EXACT) {
        d_outer = (float)(1.0 / (exp((double)min_pu) + 1.0));
        ll_term = -(fmaxf(-min_pu, 0.f) + log1pf(__expf(-fabsf(min_pu))));       // log sigma(pu) without cancellation (:270)
    } else {
        const float t = __expf(-fabsf(min_pu));               // one ex2 shared by the sigmoid and the log-likelihood
        const float rcp = __frcp_rn(1.0f + t);
        d_outer = min_pu > 0.f ? t * rcp : rcp;
        ll_term = fminf(min_pu, 0.f) - __logf(1.0f + t);
    }
    const float smul = sw * mult;
    if (upd) { acc.ll_f += ll_term; acc.draws += sampled; }  // every lane of the group adds the same value; flush_acc keeps lane `sub==0`
    if (p.trace && upd && sub == 0) { p.trace[2 * r] = min_j; p.trace[2 * r + 1] = sampled; }
    const float eta = p.eta, ra = p.reg_a, rb = p.reg_b;
    // EXACT keeps the reference's operand association; the Hogwild schedule folds it into two FMAs per element
    const float ec = eta * (smul * d_outer);
#define RFM_G(d, w) (EXACT ? (eta * ((smul * (d_outer * (d))) - (rb_or_ra * (w)))) : fmaf(ec, (d), -(eta * rb_or_ra) * (w)))

    float4 dx = zero4();
    if (FEAT) { dx.x = pos.x.x - neg.x.x; dx.y = pos.x.y - neg.x.y; dx.z = pos.x.z - neg.x.z; dx.w = pos.x.w - neg.x.w; }

    // d u / d v_u = (v_i - v_j) + sum_q v_if[q] (x_if[i,q] - x_if[j,q])   (:292,303-305)
    float4 dvu[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        dvu[k].x = pos.v[k].x - neg.v[k].x; dvu[k].y = pos.v[k].y - neg.v[k].y;
        dvu[k].z = pos.v[k].z - neg.v[k].z; dvu[k].w = pos.v[k].w - neg.v[k].w;
    }
    if constexpr (F8) {
        feat8_item_diff<G>(dx, *f8);
        feat8_dvu<G, QPL>(T, gp, upd, sub, *f8, dvu);
    } else if (FEAT && T.x_if_any) {
        for (int q = 0; q < T.Q; ++q) {
            const float dxq = __shfl_sync(0xffffffffu, get4(dx, q & 3), q >> 2, G);
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                const int qq = sub + k * G;
                if (upd && qq < T.NQ) {
                    const float4 w = gp_ld4<GPS>(gp + T.gp_vif + (size_t)q * T.Fp + 4 * qq);
                    dvu[k].x += w.x * dxq; dvu[k].y += w.y * dxq; dvu[k].z += w.z * dxq; dvu[k].w += w.w * dxq;
                }
            }
        }
    }

    float4 vu_new[QPL], dij_new[QPL];   // updated v_u and (v_i - v_j), needed by the feature-factor updates
    {
        const float rb_or_ra = ra;
        if (upd && sub == 0) {           // item biases (:279-280)
            sink.bias(RFM_G(1.0f, pos.w), RFM_G(-1.0f, neg.w));
        }
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const int q = sub + k * G;
            float4 du, di, dj;
            du.x = RFM_G(dvu[k].x, uc.vu[k].x); du.y = RFM_G(dvu[k].y, uc.vu[k].y);
            du.z = RFM_G(dvu[k].z, uc.vu[k].z); du.w = RFM_G(dvu[k].w, uc.vu[k].w);
            di.x = RFM_G(uc.a[k].x, pos.v[k].x); di.y = RFM_G(uc.a[k].y, pos.v[k].y);
            di.z = RFM_G(uc.a[k].z, pos.v[k].z); di.w = RFM_G(uc.a[k].w, pos.v[k].w);
            dj.x = RFM_G(-uc.a[k].x, neg.v[k].x); dj.y = RFM_G(-uc.a[k].y, neg.v[k].y);
            dj.z = RFM_G(-uc.a[k].z, neg.v[k].z); dj.w = RFM_G(-uc.a[k].w, neg.v[k].w);
            if (upd && q < T.NQ) {
                sink.user(q, du);
                sink.item_i(q, di);
                sink.item_j(q, dj);
            }
            if (FEAT) {
                vu_new[k].x = uc.vu[k].x + du.x; vu_new[k].y = uc.vu[k].y + du.y;
                vu_new[k].z = uc.vu[k].z + du.z; vu_new[k].w = uc.vu[k].w + du.w;
                dij_new[k].x = (pos.v[k].x + di.x) - (neg.v[k].x + dj.x); dij_new[k].y = (pos.v[k].y + di.y) - (neg.v[k].y + dj.y);
                dij_new[k].z = (pos.v[k].z + di.z) - (neg.v[k].z + dj.z); dij_new[k].w = (pos.v[k].w + di.w) - (neg.v[k].w + dj.w);
            }
        }
    }
    if constexpr (F8) {
        if constexpr (G >= 8 && G < 32) {
            if (p.gp_race == 2) feat8_update_chains_warp<G, QPL>(T, gp, upd, sub, *f8, dx, ec, eta * rb, vu_new, dij_new);   // warp-uniform
            else feat8_update_chains<G, QPL>(T, gp, upd, sub, *f8, dx, ec, eta * rb, vu_new, dij_new);
        } else {
            feat8_update_chains<G, QPL>(T, gp, upd, sub, *f8, dx, ec, eta * rb, vu_new, dij_new);
        }
    } else if (FEAT) {
        const float rb_or_ra = rb;
        if (T.x_if_any) {
            if (upd && 4 * sub < T.Qp) {                                   // w_if, every q (:283-286)
                const float4 w = gp_ld4<GPS>(gp + 4 * sub);
                float4 d;
                d.x = RFM_G(dx.x, w.x); d.y = RFM_G(dx.y, w.y); d.z = RFM_G(dx.z, w.z); d.w = RFM_G(dx.w, w.w);
                gp_add4<G, GPS>(gp + 4 * sub, w, d, (p.gp_private | p.gp_race) != 0);
            }
        }
        if (T.x_uf_any) {                                                  // v_uf[p] for x_uf[u,p] != 0 (:313-318)
            for (int pp = 0; pp < T.P; ++pp) {
                const float xp = __shfl_sync(0xffffffffu, get4(uc.xu, pp & 3), pp >> 2, G);
                const bool nz = xp != 0.0f;               // predicate, not `continue`: other groups of the warp still need the shuffle
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    const int q = sub + k * G;
                    if (upd && nz && q < T.NQ) {
                        float* wp = gp + T.gp_vuf + (size_t)pp * T.Fp + 4 * q;
                        const float4 w = gp_ld4<GPS>(wp);
                        float4 d;
                        d.x = RFM_G(xp * dij_new[k].x, w.x); d.y = RFM_G(xp * dij_new[k].y, w.y);
                        d.z = RFM_G(xp * dij_new[k].z, w.z); d.w = RFM_G(xp * dij_new[k].w, w.w);
                        gp_add4<G, GPS>(wp, w, d, (p.gp_private | p.gp_race) != 0);
                    }
                }
            }
        }
        if (T.x_if_any) {                                                  // v_if[q] for dx[q] != 0 (:321-326)
            for (int q = 0; q < T.Q; ++q) {
                const float dxq = __shfl_sync(0xffffffffu, get4(dx, q & 3), q >> 2, G);
                const bool nz = dxq != 0.0f;
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    const int qq = sub + k * G;
                    if (upd && nz && qq < T.NQ) {
                        float* wp = gp + T.gp_vif + (size_t)q * T.Fp + 4 * qq;
                        const float4 w = gp_ld4<GPS>(wp);
                        float4 d;
                        d.x = RFM_G(dxq * vu_new[k].x, w.x); d.y = RFM_G(dxq * vu_new[k].y, w.y);
                        d.z = RFM_G(dxq * vu_new[k].z, w.z); d.w = RFM_G(dxq * vu_new[k].w, w.w);
                        gp_add4<G, GPS>(wp, w, d, (p.gp_private | p.gp_race) != 0);
                    }
                }
            }
        }
    }
#undef RFM_G
}

// fold the per-lane accumulators of a warp into the epoch record; lanes with sub != 0 hold copies of their group's sums
template <int G>
__device__ __forceinline__ void flush_acc(StepAcc& a, EpochAcc* out)
{
    a.fold();
    if ((threadIdx.x & 31) % G != 0) { a.ll = 0.0; a.draws64 = 0; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a.ll += __shfl_xor_sync(0xffffffffu, a.ll, off);
        a.draws64 += __shfl_xor_sync(0xffffffffu, a.draws64, off);
        a.bad |= __shfl_xor_sync(0xffffffffu, a.bad, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out->ll, a.ll);
        atomicAdd(reinterpret_cast<unsigned long long*>(&out->draws), (unsigned long long)a.draws64);
        if (a.bad) atomicOr(&out->bad, 1);
    }
}

}  // namespace rfm
