// rfm_train.cu -- the SGD epoch kernels: B200 replacement of the per-epoch loop of `_fit`
// (rankfm/_rankfm.pyx:218-326: shuffle order, compute_ui_utility, WARP/BPR rejection sampler, gradient step).
//
// HBM-bound gather/scatter (arithmetic intensity ~1 FLOP/B, no tensor cores by design): per positive one gather of
// the user's fat row, one of the positive item's, one per negative candidate; updates leave as vector reductions
// (REDG.E.ADD.F32x4) so concurrent groups never lose an update.
//
//   sgd_pipe_kernel    production (Hogwild).  Each warp alternates between
//                        front end  -- 32 positives at once, one per lane: epoch permutation, (u,i) fetch, Philox draw
//                                      and membership test of the first negative (32 independent small-load chains)
//                        back end   -- lane groups walk those 32 tuples; the three fat rows of every tuple are staged
//                                      into shared memory by TMA bulk copies (cp.async.bulk -> UBLKCP) that run `depth`
//                                      steps ahead and complete on per-stage mbarriers, so row traffic for several
//                                      positives per warp is in flight without costing registers.
//   sgd_serial_kernel  one lane group, positives strictly in order, optionally the reference's MT19937 stream:
//                      exact replay of sequential SGD, used to pin the arithmetic against the oracle / the reference.
// Both call the same sample_negatives / apply_update (rfm_sgd.cuh).
#include <cstdlib>
#include <cstring>
#include "rfm_sgd.cuh"

namespace rfm {

// ---------------------------------------------------------------------------------------------------------------
// serial schedule
// ---------------------------------------------------------------------------------------------------------------
template <int G, int QPL, bool FEAT, bool MT>
__global__ void __launch_bounds__(32) sgd_serial_kernel(const TrainParams p)
{
    __shared__ MtState mt_smem;
    const Tables& T = p.T;
    const int lane = threadIdx.x & 31;
    const int sub = lane % G, gw = lane / G;
    if (MT) {
        for (int k = lane; k < kMtN; k += 32) mt_smem.s[k] = p.mt->s[k];
        if (lane == 0) mt_smem.pos = p.mt->pos;
        __syncwarp();
    }
    StepAcc acc;
    for (long long r = 0; r < p.N; ++r) {
        const bool valid = gw == 0;
        // ---- locate the observed (user, item, sample weight): _rankfm.pyx:233-236 ----
        int row = 0;
        if (valid) row = p.perm ? __ldg(p.perm + r) : (int)feistel_perm(p.feistel, r);
        int2 ui = make_int2(0, 0);
        float sw = 0.f;
        if (valid) { ui = __ldg(p.interactions + row); sw = __ldg(p.sample_weight + row); }
        const int u = ui.x, i = ui.y;
        UserCtx<QPL> uc;
        ItemRow<QPL> pos, neg;
        load_user<G, QPL, FEAT>(T, u, valid, sub, uc);
        load_item<G, QPL, FEAT>(T, i, valid, sub, pos);
        long long seg = 0; int deg = 0;
        if (valid) { seg = __ldg(p.indptr + u); deg = (int)(__ldg(p.indptr + u + 1) - seg); }
        user_precompute<G, QPL, FEAT>(T, T.GP, valid, sub, uc);
        const float ut_ui = utility<G, QPL, FEAT>(uc, pos);
        // ---- WARP / BPR sampling loop: _rankfm.pyx:244-264 ----
        int sampled = 0, min_j = -1;
        float min_pu = 1e6f;
        bool done = !valid;
        uint32_t attempt = 0;
#pragma unroll
        for (int k = 0; k < QPL; ++k) neg.v[k] = zero4();
        neg.x = zero4(); neg.w = 0.f;
        if (MT) sample_negatives<G, QPL, FEAT, true>(p, uc, ut_ui, row, u, seg, deg, 1, done, attempt, &mt_smem, sub, gw, neg, min_pu, min_j, sampled);
        else    sample_negatives_philox<G, QPL, FEAT>(p, uc, ut_ui, row, u, seg, deg, 0, done, attempt, p.spec, sub, gw, neg, min_pu, min_j, sampled);
        // ---- gradient step: _rankfm.pyx:267-326 ----
        const RedSink sink{T.UT + (size_t)u * T.ldu, T.IT + (size_t)i * T.ldi, T.IT + (size_t)(min_j >= 0 ? min_j : 0) * T.ldi, T.Fp};
        apply_update<G, QPL, FEAT, true, false>(p, T.GP, uc, pos, neg, min_j, sw, sampled, min_pu, valid, r, sub, acc, sink);
        acc.fold();
        __threadfence(); __syncwarp();          // the next step must observe this step's reductions
    }
    flush_acc<G>(acc, p.acc);
    if (MT) {
        __syncwarp();
        for (int k = lane; k < kMtN; k += 32) p.mt->s[k] = mt_smem.s[k];
        if (lane == 0) p.mt->pos = mt_smem.pos;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// production schedule: TMA-staged pipeline
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// TMA bulk reduction shared -> global: global[0..bytes) += shared[0..bytes) as f32 (SASS: UBLKRED), bulk-group completion
__device__ __forceinline__ void bulk_red_add_f32(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Tuple {            // one positive with its first negative candidate, produced by the front end (one per lane)
    int u, i, j, row;     // row < 0: past the end of the epoch
    float sw;
    uint32_t attempt;     // Philox words consumed so far for this positive
};

__device__ __forceinline__ bool lane_member(int cand, const int32_t* __restrict__ items, int deg)
{
    int lo = 0, hi = deg - 1;
    while (lo <= hi) {
        const int md = (lo + hi) >> 1;
        const int e = __ldg(items + md);
        if (e == cand) return true;
        if (e < cand) lo = md + 1; else hi = md - 1;
    }
    return false;
}

__device__ __forceinline__ Tuple front_end(const TrainParams& p, long long r)
{
    Tuple t;
    t.u = t.i = t.j = 0; t.row = -1; t.sw = 0.f; t.attempt = 0u;
    if (r < p.N) {
        const int row = p.perm ? __ldg(p.perm + r) : (int)feistel_perm(p.feistel, r);
        const int2 ui = __ldg(p.interactions + row);
        t.row = row; t.u = ui.x; t.i = ui.y;
        t.sw = __ldg(p.sample_weight + row);
        long long seg = 0; int deg = 0;
        if (!p.bitmap) { seg = __ldg(p.indptr + ui.x); deg = (int)(__ldg(p.indptr + ui.x + 1) - seg); }
        Philox4 blk = {0u, 0u, 0u, 0u};
        int rejects = 0;
        for (;;) {                                           // first draw of the positive (:250-253)
            if ((t.attempt & 3u) == 0u) blk = philox4x32_10((uint32_t)row, p.epoch_key, t.attempt >> 2, 0u, p.k0, p.k1);
            const uint32_t c = t.attempt & 3u;
            const uint32_t word = c == 0u ? blk.x : (c == 1u ? blk.y : (c == 2u ? blk.z : blk.w));
            ++t.attempt;
            const int cj = (int)__umulhi(word, (uint32_t)p.T.I);
            bool member;
            if (p.bitmap) member = ((__ldg(p.bitmap + (size_t)ui.x * p.bitmap_words + (cj >> 5)) >> (cj & 31)) & 1u) != 0u;
            else {
                bool maybe = deg > 0;
                if (p.bloom && maybe) { const uint32_t h = bloom_slot(cj, deg); maybe = ((__ldg(p.bloom + seg + (h >> 5)) >> (h & 31)) & 1u) != 0u; }
                member = maybe && lane_member(cj, p.indices + seg, deg);      // the filter has no false negatives: a clear bit is final
            }
            t.j = cj;
            if (!member || ++rejects >= p.max_rejects) break;
        }
    }
    return t;
}

template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void smem_user(const Tables& T, const float* s, int sub, UserCtx<QPL>& c)
{
    const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
    for (int k = 0; k < QPL; ++k) { const int q = sub + k * G; c.vu[k] = q < T.NQ ? s4[q] : zero4(); }
    if (FEAT) c.xu = 4 * sub < T.Pp ? s4[T.NQ + sub] : zero4();
}
template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void smem_item(const Tables& T, const float* s, int sub, ItemRow<QPL>& r)
{
    const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
    for (int k = 0; k < QPL; ++k) { const int q = sub + k * G; r.v[k] = q < T.NQ ? s4[q] : zero4(); }
    r.w = s[T.Fp];
    if (FEAT) r.x = 4 * sub < T.Qp ? s4[T.NQ + 1 + sub] : zero4();
}

constexpr int kPipeBarBytes = 128;      // up to 16 mbarriers per warp, keeps the stages 128-byte aligned

// resident blocks per SM the register allocator is told to aim for: BPR fits 3 without spilling; the WARP look-ahead
// sampler and the feature variants need the 128-register budget of 2
template <int QPL, bool FEAT, bool WARP>
constexpr int pipe_min_blocks() { return QPL > 2 ? 1 : ((FEAT || WARP || QPL == 2) ? 2 : 3); }

// F8: side-feature math specialised for P, Q <= 8 (rfm_feat8.cuh); needs G >= 8 and a chain copy per lane group
// MINB > 0 overrides the resident-blocks target (experiments: the WARP kernel at 3 blocks/SM trades ~60 spilled words for
// 50 % more warps to hide the L2 latency of the sampler's candidate rows)
template <int G, int QPL, bool FEAT, bool WARP, bool TRED, bool F8 = false, int MINB = 0>
__global__ void __launch_bounds__(kTrainThreads, MINB > 0 ? MINB : pipe_min_blocks<QPL, FEAT, WARP>()) sgd_pipe_kernel(const TrainParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Tables& T = p.T;
    constexpr int GPW = 32 / G;                       // tuples processed per step (one per lane group)
    constexpr int STEPS = 32 / GPW;                   // steps per batch of 32 tuples
    const int lane = threadIdx.x & 31, sub = lane % G, gw = lane / G;
    const int D = p.depth;
    const int tuple_floats = T.ldu + 2 * T.ldi;
    const uint32_t tuple_bytes = (uint32_t)tuple_floats * 4u;
    const int stage_floats = GPW * tuple_floats;
    // per warp: [mbarriers | D stages | chain copies of the feature parameters GP (FEAT only)]: one copy per lane group
    // (plain read-modify-write, p.gp_private) when that fits shared memory, else one per warp (shared-memory atomics)
    const int gp_floats = FEAT ? p.gp_floats : 0;
    const int gp_copies = FEAT ? (p.gp_private ? GPW : 1) : 0;
    unsigned char* wbase = smem_raw + (size_t)(threadIdx.x >> 5) * (kPipeBarBytes + ((size_t)D * stage_floats + (size_t)gp_copies * gp_floats) * 4);
    const uint32_t bars = smem_u32(wbase);
    float* stages = reinterpret_cast<float*>(wbase + kPipeBarBytes);
    float* gp_all = FEAT ? stages + (size_t)D * stage_floats : nullptr;
    float* gp = FEAT ? gp_all + (size_t)(p.gp_private ? gw : 0) * gp_floats : nullptr;
    if (FEAT) {          // every chain starts the epoch from the same parameters and evolves over its own positives
        for (int e = lane; e < gp_copies * gp_floats; e += 32) gp_all[e] = __ldcg(T.GP + (e % gp_floats));
        __syncwarp();
    }
    if (lane == 0) {
        for (int d = 0; d < D; ++d) mbar_init(bars + 8u * d, 1u);      // one arrival: the lane that posts expect_tx
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long n_batches = (p.N + 31) / 32;
    // look-ahead of the WARP sampler: follow the previous epoch's mean number of draws per positive (device-side,
    // no host round trip); early epochs violate the margin at the first draw and would waste the look-ahead
    int spec = p.spec;
    if (spec == 0) {
        spec = 1;
        if (p.prev_acc) {
            const long long prev = p.prev_acc->draws;
            spec = prev >= 3 * p.N ? 4 : (prev >= 3 * p.N / 2 ? 2 : 1);
        }
    }
    StepAcc acc;
    uint32_t st_issue = 0, st_cons = 0, par_cons = 0;   // ring positions of the next stage to fill / to drain, and its parity
    const uint32_t stage0 = smem_u32(stages);
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;

    // stage the rows of step `step` of the batch held in `cur` (warp-collective)
    auto issue = [&](const Tuple& cur, int step) {
        const int k = step * GPW + gw;
        const int tu = __shfl_sync(0xffffffffu, cur.u, k), ti = __shfl_sync(0xffffffffu, cur.i, k);
        const int tj = __shfl_sync(0xffffffffu, cur.j, k), trow = __shfl_sync(0xffffffffu, cur.row, k);
        const bool ok = trow >= 0;
        const uint32_t bar = bars + 8u * st_issue;
        const uint32_t dst = stage0 + st_issue * stage_bytes + (uint32_t)gw * tuple_bytes;
        const float* su = T.UT + (size_t)tu * T.ldu;
        const float* si = T.IT + (size_t)ti * T.ldi;
        const float* sj = T.IT + (size_t)tj * T.ldi;
        const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, ok && sub == 0));
        if (lane == 0) mbar_expect_tx(bar, nvalid * tuple_bytes);
        __syncwarp();
        if (ok && sub < 3) {
            if (sub == 0)      bulk_g2s(dst, su, (uint32_t)T.ldu * 4u, bar);
            else if (sub == 1) bulk_g2s(dst + (uint32_t)T.ldu * 4u, si, (uint32_t)T.ldi * 4u, bar);
            else               bulk_g2s(dst + (uint32_t)(T.ldu + T.ldi) * 4u, sj, (uint32_t)T.ldi * 4u, bar);
        }
        st_issue = st_issue + 1 == (uint32_t)D ? 0u : st_issue + 1;
    };

    Tuple cur = front_end(p, warp_global * 32 + lane);
    for (long long b = warp_global; b < n_batches; b += n_warps) {
        const int ahead = D < STEPS ? D : STEPS;
        for (int d = 0; d < ahead; ++d) issue(cur, d);
        // next batch's front end overlaps with the copies just issued
        const Tuple nxt = front_end(p, (b + n_warps) * 32 + lane);

        for (int step = 0; step < STEPS; ++step) {
            const int k = step * GPW + gw;
            const int u = __shfl_sync(0xffffffffu, cur.u, k), i = __shfl_sync(0xffffffffu, cur.i, k);
            const int j1 = __shfl_sync(0xffffffffu, cur.j, k), row = __shfl_sync(0xffffffffu, cur.row, k);
            const float sw = __shfl_sync(0xffffffffu, cur.sw, k);
            uint32_t attempt = __shfl_sync(0xffffffffu, cur.attempt, k);
            const bool valid = row >= 0;
            while (!mbar_try_wait(bars + 8u * st_cons, par_cons)) { }
            const float* base = stages + (size_t)st_cons * stage_floats + (size_t)gw * tuple_floats;
            if (++st_cons == (uint32_t)D) { st_cons = 0; par_cons ^= 1u; }
            UserCtx<QPL> uc;
            ItemRow<QPL> pos, neg;
            if (valid) {
                smem_user<G, QPL, FEAT>(T, base, sub, uc);
                smem_item<G, QPL, FEAT>(T, base + T.ldu, sub, pos);
                smem_item<G, QPL, FEAT>(T, base + T.ldu + T.ldi, sub, neg);
            } else {
#pragma unroll
                for (int q = 0; q < QPL; ++q) { uc.vu[q] = zero4(); pos.v[q] = zero4(); neg.v[q] = zero4(); }
                uc.xu = zero4(); pos.x = zero4(); neg.x = zero4(); pos.w = 0.f; neg.w = 0.f;
            }
            Feat8 f8;
            if constexpr (F8) {
                feat8_user(T, base, valid, f8);
                user_precompute8<G, QPL>(T, gp, valid, sub, f8, uc);
            } else {
                user_precompute<G, QPL, FEAT, true>(T, gp, valid, sub, uc);
            }
            float ut_ui = 0.f, pu1;
            if (!WARP) {                       // BPR: only the difference is needed -> one group reduction instead of two
                float part = 0.f;
#pragma unroll
                for (int q = 0; q < QPL; ++q) {
                    const float4 d = make_float4(pos.v[q].x - neg.v[q].x, pos.v[q].y - neg.v[q].y, pos.v[q].z - neg.v[q].z, pos.v[q].w - neg.v[q].w);
                    part = dot4(uc.a[q], d, part);
                }
                if (FEAT) part = dot4(uc.b, make_float4(pos.x.x - neg.x.x, pos.x.y - neg.x.y, pos.x.z - neg.x.z, pos.x.w - neg.x.w), part);
                pu1 = (pos.w - neg.w) + group_sum<G>(part);
            } else {
                ut_ui = utility<G, QPL, FEAT>(uc, pos);
                pu1 = ut_ui - utility<G, QPL, FEAT>(uc, neg);
            }                                  // the group reductions also order every smem read of this stage
            // draw 1 of the reference's loop (:247-264); NaN leaves min_index at -1 like `pu < 1e6` failing
            int sampled = 1;
            float min_pu = 1e6f;
            int min_j = -1;
            if (pu1 < min_pu) { min_pu = pu1; min_j = j1; }
            bool done = !valid || pu1 < 1.0f;
            __syncwarp();
            if (!TRED) { if (step + D < STEPS) issue(cur, step + D); }      // refill the stage just drained
            if (WARP && __any_sync(0xffffffffu, !done)) {
                long long seg = 0; int deg = 0;
                if (!done && !p.bitmap) { seg = __ldg(p.indptr + u); deg = (int)(__ldg(p.indptr + u + 1) - seg); }
                sample_negatives_philox<G, QPL, FEAT>(p, uc, ut_ui, row, u, seg, deg, 1, done, attempt, spec, sub, gw, neg, min_pu, min_j, sampled);
            }
            const int jj = min_j >= 0 ? min_j : 0;
            if (TRED) {
                // deltas overwrite the staged rows of this tuple, then three bulk reductions ship them (one per row)
                float4* slot = reinterpret_cast<float4*>(const_cast<float*>(base));
                const int nu4 = T.ldu >> 2, ni4 = T.ldi >> 2;
                const SmemSink sink{slot, slot + nu4, slot + nu4 + ni4, T.NQ};
                if (FEAT && valid) {      // the read-only feature blocks must add 0
                    if (4 * sub < T.Pp) slot[T.NQ + sub] = zero4();
                    if (4 * sub < T.Qp) { slot[nu4 + T.NQ + 1 + sub] = zero4(); slot[nu4 + ni4 + T.NQ + 1 + sub] = zero4(); }
                }
                apply_update<G, QPL, FEAT, false, true, SmemSink, F8>(p, gp, uc, pos, neg, min_j, sw, sampled, min_pu, valid, b * 32 + k, sub, acc, sink, &f8);
                fence_async_smem();
                __syncwarp();
                const bool upd = valid && min_j >= 0;
                if (upd && sub < 3) {
                    const uint32_t src = smem_u32(slot);
                    if (sub == 0)      bulk_red_add_f32(T.UT + (size_t)u * T.ldu, src, (uint32_t)T.ldu * 4u);
                    else if (sub == 1) bulk_red_add_f32(T.IT + (size_t)i * T.ldi, src + (uint32_t)T.ldu * 4u, (uint32_t)T.ldi * 4u);
                    else               bulk_red_add_f32(T.IT + (size_t)jj * T.ldi, src + (uint32_t)(T.ldu + T.ldi) * 4u, (uint32_t)T.ldi * 4u);
                }
                if (sub < 3) bulk_commit();
                // refill the stage drained one step ago: its reductions (the group before the one just committed) have had
                // a whole step to read their source; the group just committed may still be reading
                if (step >= 1 && step - 1 + D < STEPS) { if (sub < 3) bulk_wait_read1(); __syncwarp(); issue(cur, step - 1 + D); }
            } else {
                const RedSink sink{T.UT + (size_t)u * T.ldu, T.IT + (size_t)i * T.ldi, T.IT + (size_t)jj * T.ldi, T.Fp};
                apply_update<G, QPL, FEAT, false, true, RedSink, F8>(p, gp, uc, pos, neg, min_j, sw, sampled, min_pu, valid, b * 32 + k, sub, acc, sink, &f8);
            }
        }
        if (TRED) { if (sub < 3) bulk_wait_read0(); __syncwarp(); }   // the next batch refills every stage
        acc.fold();
        cur = nxt;
    }
    if (TRED) bulk_wait_all();
    if (FEAT && warp_global < n_batches) {
        // fold this warp's chain into the epoch result: GP_new = GP_start + gain * sum over active warps of (copy - GP_start)
        __syncwarp();
        for (int e = lane; e < gp_copies * gp_floats; e += 32) {
            const float d = gp_all[e] - __ldcg(T.GP + (e % gp_floats));
            if (d != 0.f) red_add1(p.gp_acc + (e % gp_floats), p.gp_gain * d);
        }
    }
    flush_acc<G>(acc, p.acc);
}

// ---------------------------------------------------------------------------------------------------------------
// launch: pick the group geometry from the widest row section
// ---------------------------------------------------------------------------------------------------------------
int train_group_size(const Tables& T, int* qpl_out)
{
    const int widest = max(T.Fp, max(T.Pp, T.Qp)) / 4;   // quads
    int G = 4;
    while (G < 32 && G < widest) G <<= 1;
    int qpl = (T.NQ + G - 1) / G;
    if (qpl_out) *qpl_out = qpl;
    return G;
}

// Lane-group geometry of the SGD kernels.  Without side features: HALF the narrowest group that covers a row, two quads
// per lane -- twice the positives per warp instruction, so the scalar work per step and per sampler round (Philox, address
// arithmetic, bookkeeping, TMA issue) is shared by twice as many positives: cfg2 0.97 -> 0.86 ms, cfg3n 27.9 -> 21.5 ms,
// cfg4s 10.85 -> 10.00 ms per launch (profiles/r02_ab_sampler_occupancy.md).  RANKFM_B200_GROUP_SHIFT=0 restores one quad
// per lane.
// Side features with at most 8 + 8 columns (the feat8 code path, rfm_feat8.cuh) keep ONE feature-parameter chain per warp
// instead of one per lane group.  All lane groups of the warp read it, compute their update of the same step and store it
// with plain stores: one of them lands (per element), i.e. the chain advances by ONE positive per warp step -- the same
// dynamics as a group-private chain (which also sees one positive per warp step), with 1/GPW of the shared memory.  That is
// what lets this kernel use half-width lane groups too (two quads per lane, four positives per warp step at F=64) at two
// blocks per SM: cfg3 56.3 -> 44.0 ms, cfg3m 21.4 -> 16.7 ms per launch (profiles/r02_ab_sampler_occupancy.md).
// RANKFM_B200_CHAIN=group restores one chain per lane group, RANKFM_B200_FEAT_HALVE=0 the wide groups.
static bool feat8_shape(const Tables& T)
{
    const char* e = getenv("RANKFM_B200_FEAT8");                                   // experiments / tests: RANKFM_B200_FEAT8=0
    const bool off = e && !strcmp(e, "0");
    return !off && (T.x_uf_any || T.x_if_any) && T.P <= kFeat8 && T.Q <= kFeat8;
}
static bool chain_per_warp(const Tables& T, int G)
{
    const char* e = getenv("RANKFM_B200_CHAIN");
    if (e && !strcmp(e, "group")) return false;
    return feat8_shape(T) && G >= 8 && G < 32;
}
static bool feat_halve()
{
    const char* e = getenv("RANKFM_B200_FEAT_HALVE");
    return !(e && atoi(e) == 0);
}

static int sgd_group_size(const Tables& T, int* qpl_out)
{
    int qpl = 1;
    int G = train_group_size(T, &qpl);
    const char* e = getenv("RANKFM_B200_GROUP_SHIFT");
    const bool halve = !(e && atoi(e) == 0);
    if (halve && G >= 8 && qpl == 1 && !(T.x_uf_any || T.x_if_any)) { G >>= 1; qpl = (T.NQ + G - 1) / G; }
    // the feat8 side-feature kernel: half-width groups as well (its chain is per warp, see chain_per_warp)
    if (feat_halve() && G >= 16 && qpl == 1 && feat8_shape(T) && max(T.Pp, T.Qp) <= 2 * G && chain_per_warp(T, G >> 1)) { G >>= 1; qpl = (T.NQ + G - 1) / G; }
    if (qpl_out) *qpl_out = qpl;
    return G;
}

// depth of the staging pipeline: ~8 KB of rows in flight per warp, 2..8 stages
static int pipe_depth(const Tables& T, int G)
{
    const int stage_bytes = (32 / G) * (T.ldu + 2 * T.ldi) * 4;
    int d = 4096 / stage_bytes;                                        // measured: shallow rings win (profiles/r01_depth_sweep.md)
    if (const char* e = getenv("RANKFM_B200_DEPTH")) d = atoi(e);      // experiments
    return d < 2 ? 2 : (d > 8 ? 8 : d);
}
static size_t gp_floats_of(const Tables& T) { return (T.x_uf_any || T.x_if_any) ? (size_t)T.gp_vif + (size_t)T.Q * T.Fp : 0; }
static size_t pipe_smem_bytes_copies(const Tables& T, int G, int depth, int copies)
{
    const size_t stage_bytes = (size_t)(32 / G) * (T.ldu + 2 * T.ldi) * 4;
    return (size_t)(kTrainThreads / 32) * (kPipeBarBytes + (size_t)depth * stage_bytes + (size_t)copies * gp_floats_of(T) * 4);
}
// group-private feature-parameter chains when two blocks of them still fit an SM, else one (atomic) chain per warp
static int gp_private_of(const Tables& T, int G, int depth)
{
    if (chain_per_warp(T, G)) return 0;                                // one racing copy per warp (see chain_per_warp)
    return (32 / G) > 1 && pipe_smem_bytes_copies(T, G, depth, 32 / G) <= 100 * 1024 ? 1 : 0;
}
static size_t pipe_smem_bytes(const Tables& T, int G, int depth)
{
    return pipe_smem_bytes_copies(T, G, depth, gp_private_of(T, G, depth) ? 32 / G : 1);
}
// lane groups of a warp that share one chain copy and race for its update: the chain advances once per warp step
int sgd_pipe_groups_per_chain(const Tables& T)
{
    int qpl = 1;
    const int G = sgd_group_size(T, &qpl);
    return chain_per_warp(T, G) ? 32 / G : 1;
}

int sgd_pipe_chains_per_warp(const Tables& T)
{
    int qpl = 1;
    const int G = sgd_group_size(T, &qpl);
    return gp_private_of(T, G, pipe_depth(T, G)) ? 32 / G : 1;
}
size_t sgd_pipe_smem_bytes(const Tables& T)
{
    int qpl = 1;
    const int G = sgd_group_size(T, &qpl);
    return pipe_smem_bytes(T, G, pipe_depth(T, G));
}

template <typename K>
static cudaError_t launch_kernel(K kernel, const TrainParams& p, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kTrainThreads, smem, st>>>(p);
    return cudaGetLastError();
}

// delta path of the production kernel: TMA bulk reductions (default) or per-lane REDG; RANKFM_B200_UPDATE=redg|tma
static bool use_tred()
{
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("RANKFM_B200_UPDATE"); cached = (e && !strcmp(e, "redg")) ? 0 : 1; }
    return cached == 1;
}

// the feat8 specialisation applies to: at most 8 + 8 active feature columns, lane groups of >= 8, and a chain copy that a
// lane group may update with plain stores (its own, or the warp's racing one)
static bool feat8_ok(const Tables& T, int G, int gp_private)
{
    return feat8_shape(T) && G >= 8 && (gp_private || G == 32 || chain_per_warp(T, G));
}

static bool warp_occ3()
{
    const char* e = getenv("RANKFM_B200_WARP_OCC");                                // experiments: 2 = the 128-register build at 2 blocks per SM
    return !(e && !strcmp(e, "2"));
}

template <int G, int QPL, typename F>
static void with_pipe_kernel(bool feat, bool warp, bool tred, bool f8, F&& f)
{
    if constexpr (G >= 8) {
        if (feat && f8 && tred) {
            if constexpr (QPL == 2) {          // half-width groups (only chosen together with the per-warp chain): 2 blocks/SM
                if (warp) f(sgd_pipe_kernel<G, QPL, true, true, true, true, 2>); else f(sgd_pipe_kernel<G, QPL, true, false, true, true, 2>);
            } else {
                if (warp) f(sgd_pipe_kernel<G, QPL, true, true, true, true>); else f(sgd_pipe_kernel<G, QPL, true, false, true, true>);
            }
            return;
        }
    }
    if constexpr (QPL <= 2) {
        if (!feat && warp && tred && warp_occ3()) { f(sgd_pipe_kernel<G, QPL, false, true, true, false, 3>); return; }
    }
    if (feat) {
        if (warp) { if (tred) f(sgd_pipe_kernel<G, QPL, true, true, true>); else f(sgd_pipe_kernel<G, QPL, true, true, false>); }
        else      { if (tred) f(sgd_pipe_kernel<G, QPL, true, false, true>); else f(sgd_pipe_kernel<G, QPL, true, false, false>); }
    } else {
        if (warp) { if (tred) f(sgd_pipe_kernel<G, QPL, false, true, true>); else f(sgd_pipe_kernel<G, QPL, false, true, false>); }
        else      { if (tred) f(sgd_pipe_kernel<G, QPL, false, false, true>); else f(sgd_pipe_kernel<G, QPL, false, false, false>); }
    }
}

template <int G, int QPL>
static cudaError_t launch_pipe(const TrainParams& p, bool feat, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaSuccess;
    with_pipe_kernel<G, QPL>(feat, p.max_samples > 1, use_tred(), feat8_ok(p.T, G, p.gp_private), [&](auto kernel) { e = launch_kernel(kernel, p, grid, smem, st); });
    return e;
}

template <int G, int QPL>
static cudaError_t launch_gq(const TrainParams& p0, int grid, cudaStream_t st)
{
    TrainParams p = p0;
    const bool feat = p.T.x_uf_any || p.T.x_if_any;
    if (p.serial) {
        const bool mt = p.mt != nullptr;
        if (feat) { if (mt) sgd_serial_kernel<G, QPL, true, true><<<1, 32, 0, st>>>(p); else sgd_serial_kernel<G, QPL, true, false><<<1, 32, 0, st>>>(p); }
        else      { if (mt) sgd_serial_kernel<G, QPL, false, true><<<1, 32, 0, st>>>(p); else sgd_serial_kernel<G, QPL, false, false><<<1, 32, 0, st>>>(p); }
        return cudaGetLastError();
    }
    p.depth = pipe_depth(p.T, G);
    p.gp_floats = (int)gp_floats_of(p.T);
    p.gp_private = gp_private_of(p.T, G, p.depth);
    {
        const char* e = getenv("RANKFM_B200_CHAIN");                        // group | race | winner (default)
        p.gp_race = chain_per_warp(p.T, G) ? ((e && !strcmp(e, "race")) ? 1 : 2) : 0;
    }
    const size_t smem = pipe_smem_bytes(p.T, G, p.depth);
    return launch_pipe<G, QPL>(p, feat, grid, smem, st);
}

template <int G, int QPL>
static int occ_gq(const TrainParams& p)
{
    const bool feat = p.T.x_uf_any || p.T.x_if_any;
    const size_t smem = pipe_smem_bytes(p.T, G, pipe_depth(p.T, G));
    int n = 0;
    with_pipe_kernel<G, QPL>(feat, p.max_samples > 1, use_tred(), feat8_ok(p.T, G, gp_private_of(p.T, G, pipe_depth(p.T, G))), [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kTrainThreads, smem);
    });
    return n;
}

cudaError_t launch_sgd_epoch(const TrainParams& p, int grid, cudaStream_t st)
{
    int qpl = 1;
    const int G = sgd_group_size(p.T, &qpl);
    if (max(p.T.Pp, p.T.Qp) > 4 * G || qpl > 4) return cudaErrorInvalidValue;   // caller reports RFM_ERR_UNSUPPORTED
    switch (G) {
        case 4:  return qpl == 2 ? launch_gq<4, 2>(p, grid, st) : launch_gq<4, 1>(p, grid, st);
        case 8:  return qpl == 2 ? launch_gq<8, 2>(p, grid, st) : launch_gq<8, 1>(p, grid, st);
        case 16: return qpl == 2 ? launch_gq<16, 2>(p, grid, st) : launch_gq<16, 1>(p, grid, st);
        default:
            if (qpl == 1) return launch_gq<32, 1>(p, grid, st);
            if (qpl == 2) return launch_gq<32, 2>(p, grid, st);
            return launch_gq<32, 4>(p, grid, st);
    }
}

int sgd_epoch_blocks_per_sm(const TrainParams& p)
{
    int qpl = 1;
    const int G = sgd_group_size(p.T, &qpl);
    switch (G) {
        case 4:  return qpl == 2 ? occ_gq<4, 2>(p) : occ_gq<4, 1>(p);
        case 8:  return qpl == 2 ? occ_gq<8, 2>(p) : occ_gq<8, 1>(p);
        case 16: return qpl == 2 ? occ_gq<16, 2>(p) : occ_gq<16, 1>(p);
        default: return qpl == 1 ? occ_gq<32, 1>(p) : (qpl == 2 ? occ_gq<32, 2>(p) : occ_gq<32, 4>(p));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// self-test of the feat8 specialisation: one warp runs the generic and the specialised feature code on identical
// pseudo-random inputs (user row, two item rows, chain copy) and reports the largest differences
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float selftest_value(uint32_t seed, uint32_t k) { return ((mix32(seed ^ (k * 0x9E3779B9u)) >> 8) * (1.0f / 16777216.0f) - 0.5f); }

template <int G, int QPL>
__global__ void __launch_bounds__(32) feat8_selftest_kernel(const Tables T, int gp_floats, uint32_t seed, float eta, float reg_b, float* __restrict__ out /* [8]: 0-4 results, 6-7 multiplier table */)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x, sub = lane % G, gw = lane / G;
    // per group: [user row | item row i | item row j | chain copy A | chain copy B | delta rows A (3) | delta rows B (3)]
    const int rows = T.ldu + 2 * T.ldi;
    float* mine = sm + (size_t)gw * (3 * rows + 2 * gp_floats);
    float *urow = mine, *irow = urow + T.ldu, *jrow = irow + T.ldi, *gpa = jrow + T.ldi, *gpb = gpa + gp_floats, *da = gpb + gp_floats, *db = da + rows;
    for (int e = sub; e < rows; e += G) {
        const int col_u = e, col_i = e - T.ldu, col_j = e - T.ldu - T.ldi;
        float v = selftest_value(seed + gw, e);
        // pads must be zero, and some feature values exactly zero (the skip-if-zero branches)
        if (e < T.ldu) { if ((col_u >= T.F && col_u < T.Fp) || col_u >= T.Fp + T.P || (col_u >= T.Fp && ((col_u + gw) % 3 == 0))) v = 0.f; }
        else {
            const int c = e < T.ldu + T.ldi ? col_i : col_j;
            if ((c >= T.F && c < T.Fp) || (c > T.Fp && c < T.Fp + 4) || c >= T.Fp + 4 + T.Q) v = 0.f;
            if (c >= T.Fp + 4 && (c % 4 == 1)) v = e < T.ldu + T.ldi ? 0.25f : 0.25f;          // equal on both items: dx == 0 there
        }
        mine[e] = v;
    }
    for (int e = sub; e < gp_floats; e += G) {
        float v = 0.3f * selftest_value(seed * 7u + gw, e);
        const int o = e >= T.gp_vif ? (e - T.gp_vif) % T.Fp : (e >= T.gp_vuf ? (e - T.gp_vuf) % T.Fp : 0);
        if (o >= T.F || (e < T.gp_vuf && e >= T.Q)) v = 0.f;
        gpa[e] = v; gpb[e] = v;
    }
    // warp-shared chain copy for the winner-update path: starts as the LAST group's chain (the winner = highest valid group)
    float* gpc = sm + (size_t)GPW * (3 * rows + 2 * gp_floats);
    if (G < 32) {
        for (int e = lane; e < gp_floats; e += 32) {
            float v = 0.3f * selftest_value(seed * 7u + (GPW - 1), e);
            const int o = e >= T.gp_vif ? (e - T.gp_vif) % T.Fp : (e >= T.gp_vuf ? (e - T.gp_vuf) % T.Fp : 0);
            if (o >= T.F || (e < T.gp_vuf && e >= T.Q)) v = 0.f;
            gpc[e] = v;
        }
    }
    __syncwarp();
    UserCtx<QPL> ua, ub;
    ItemRow<QPL> pos, neg;
    smem_user<G, QPL, true>(T, urow, sub, ua);
    smem_item<G, QPL, true>(T, irow, sub, pos);
    smem_item<G, QPL, true>(T, jrow, sub, neg);
    ub = ua;
    Feat8 f8;
    user_precompute<G, QPL, true, true>(T, gpa, true, sub, ua);
    feat8_user(T, urow, true, f8);
    user_precompute8<G, QPL>(T, gpb, true, sub, f8, ub);
    float d_a = 0.f, d_b = 0.f;
#pragma unroll
    for (int k = 0; k < QPL; ++k)
        d_a = fmaxf(d_a, fmaxf(fmaxf(fabsf(ua.a[k].x - ub.a[k].x), fabsf(ua.a[k].y - ub.a[k].y)), fmaxf(fabsf(ua.a[k].z - ub.a[k].z), fabsf(ua.a[k].w - ub.a[k].w))));
    d_b = fmaxf(fmaxf(fabsf(ua.b.x - ub.b.x), fabsf(ua.b.y - ub.b.y)), fmaxf(fabsf(ua.b.z - ub.b.z), fabsf(ua.b.w - ub.b.w)));
    // one gradient step with each code path on its own chain copy
    TrainParams p{};
    p.T = T; p.eta = eta; p.reg_a = 0.02f; p.reg_b = reg_b; p.gp_private = 1;
    p.mult = out + 6;                        // WARP multiplier table {-, 0.9} in global memory (apply_update reads it with __ldg)
    StepAcc acc_a, acc_b;
    const int nu4 = T.ldu >> 2, ni4 = T.ldi >> 2;
    const SmemSink sa{reinterpret_cast<float4*>(da), reinterpret_cast<float4*>(da) + nu4, reinterpret_cast<float4*>(da) + nu4 + ni4, T.NQ};
    const SmemSink sb{reinterpret_cast<float4*>(db), reinterpret_cast<float4*>(db) + nu4, reinterpret_cast<float4*>(db) + nu4 + ni4, T.NQ};
    for (int e = sub; e < rows; e += G) { da[e] = 0.f; db[e] = 0.f; }
    __syncwarp();
    const float pu = utility<G, QPL, true>(ua, pos) - utility<G, QPL, true>(ua, neg);
    apply_update<G, QPL, true, false, true>(p, gpa, ua, pos, neg, 1, 1.3f, 1, pu, true, 0, sub, acc_a, sa);
    apply_update<G, QPL, true, false, true, SmemSink, true>(p, gpb, ua, pos, neg, 1, 1.3f, 1, pu, true, 0, sub, acc_b, sb, &f8);
    __syncwarp();
    float d_gp = 0.f, d_rows = 0.f, moved = 0.f;
    for (int e = sub; e < gp_floats; e += G) { d_gp = fmaxf(d_gp, fabsf(gpa[e] - gpb[e])); moved = fmaxf(moved, fabsf(gpa[e] - 0.3f * selftest_value(seed * 7u + gw, e))); }
    for (int e = sub; e < rows; e += G) d_rows = fmaxf(d_rows, fabsf(da[e] - db[e]));
    if constexpr (G >= 8 && G < 32) {
        // the warp-wide winner update (feat8_update_chains_warp) on the shared copy must leave exactly what the last group's
        // own feat8 update left on its private copy
        TrainParams p2 = p;
        p2.gp_private = 0; p2.gp_race = 2;
        __syncwarp();
        for (int e = sub; e < rows; e += G) da[e] = 0.f;                    // scratch sink
        __syncwarp();
        apply_update<G, QPL, true, false, true, SmemSink, true>(p2, gpc, ua, pos, neg, 1, 1.3f, 1, pu, true, 0, sub, acc_b, sa, &f8);
        __syncwarp();
        const float* win = sm + (size_t)(GPW - 1) * (3 * rows + 2 * gp_floats) + rows + gp_floats;      // gpb of the last group
        for (int e = lane; e < gp_floats; e += 32) d_gp = fmaxf(d_gp, fabsf(gpc[e] - win[e]));
    }
    for (int off = 16; off > 0; off >>= 1) {
        d_a = fmaxf(d_a, __shfl_xor_sync(0xffffffffu, d_a, off)); d_b = fmaxf(d_b, __shfl_xor_sync(0xffffffffu, d_b, off));
        d_gp = fmaxf(d_gp, __shfl_xor_sync(0xffffffffu, d_gp, off)); d_rows = fmaxf(d_rows, __shfl_xor_sync(0xffffffffu, d_rows, off));
        moved = fmaxf(moved, __shfl_xor_sync(0xffffffffu, moved, off));
    }
    if (lane == 0) {                                          // max with what an earlier launch (another group shape) left
        out[0] = fmaxf(out[0], d_a); out[1] = fmaxf(out[1], d_b); out[2] = fmaxf(out[2], d_gp); out[3] = fmaxf(out[3], d_rows); out[4] = fmaxf(out[4], moved);
    }
    (void)GPW;
}

template <int G, int QPL>
static cudaError_t feat8_selftest_gq(const Tables& T, int gp_floats, uint32_t seed, float eta, float reg_b, float* out, cudaStream_t st)
{
    if constexpr (G >= 8) {
        const size_t smem = ((size_t)(32 / G) * (3 * (T.ldu + 2 * T.ldi) + 2 * gp_floats) + (size_t)gp_floats) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(feat8_selftest_kernel<G, QPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        feat8_selftest_kernel<G, QPL><<<1, 32, smem, st>>>(T, gp_floats, seed, eta, reg_b, out);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidValue;
    }
}

cudaError_t launch_feat8_selftest(const Tables& T, uint32_t seed, float eta, float reg_b, float* out5, cudaStream_t st)
{
    int qpl = 1;
    const int G = train_group_size(T, &qpl);
    const int gpf = (int)gp_floats_of(T);
    if (T.P > kFeat8 || T.Q > kFeat8 || G < 8) return cudaErrorInvalidValue;
    {   // the production kernel's half-width shape (sgd_group_size), when it differs: same checks, results folded by max
        int hq = 1;
        const int hg = sgd_group_size(T, &hq);
        if (hg != G || hq != qpl) {
            cudaError_t e = cudaSuccess;
            if (hg == 8 && hq == 2) e = feat8_selftest_gq<8, 2>(T, gpf, seed, eta, reg_b, out5, st);
            else if (hg == 16 && hq == 2) e = feat8_selftest_gq<16, 2>(T, gpf, seed, eta, reg_b, out5, st);
            if (e != cudaSuccess) return e;
        }
    }
    switch (G) {
        case 8:  return feat8_selftest_gq<8, 1>(T, gpf, seed, eta, reg_b, out5, st);
        case 16: return feat8_selftest_gq<16, 1>(T, gpf, seed, eta, reg_b, out5, st);
        default:
            if (qpl == 1) return feat8_selftest_gq<32, 1>(T, gpf, seed, eta, reg_b, out5, st);
            if (qpl == 2) return feat8_selftest_gq<32, 2>(T, gpf, seed, eta, reg_b, out5, st);
            return feat8_selftest_gq<32, 4>(T, gpf, seed, eta, reg_b, out5, st);
    }
}

// GP += accumulated (already gain-weighted) warp deltas; clears the accumulator for the next epoch
__global__ void gp_apply_kernel(float* __restrict__ gp, float* __restrict__ acc, int n)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) { gp[e] += acc[e]; acc[e] = 0.f; }
}
cudaError_t launch_gp_apply(float* gp, float* acc, int n, cudaStream_t st)
{
    gp_apply_kernel<<<(n + 255) / 256 > 64 ? 64 : (n + 255) / 256, 256, 0, st>>>(gp, acc, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// per-user membership bitmap (small catalogues): replaces the search of user_items by one bit test per draw
// ---------------------------------------------------------------------------------------------------------------
// indptr [U+1] holds offsets into `indices` that need not start at 0 (a rank's slice of a global CSR); row u of the bitmap
// belongs to the u-th user of that slice
__global__ void build_bitmap_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int U, uint32_t* __restrict__ bitmap, int words)
{
    const long long first = indptr[0], nnz = indptr[U];
    for (long long e = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = U;                                   // owner of CSR slot e: last u with indptr[u] <= e
        while (hi - lo > 1) { const int md = (lo + hi) >> 1; if (indptr[md] <= e) lo = md; else hi = md; }
        const int it = indices[e];
        atomicOr(bitmap + (size_t)lo * words + (it >> 5), 1u << (it & 31));
    }
}

// membership filter for large catalogues (see TrainParams::bloom): word e of `bloom` belongs to the owner of CSR slot e
__global__ void build_bloom_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int U, uint32_t* __restrict__ bloom)
{
    const long long first = indptr[0], nnz = indptr[U];
    for (long long e = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = U;                                   // owner of CSR slot e: last u with indptr[u] <= e
        while (hi - lo > 1) { const int md = (lo + hi) >> 1; if (indptr[md] <= e) lo = md; else hi = md; }
        const long long seg = indptr[lo];
        const int deg = (int)(indptr[lo + 1] - seg);
        const uint32_t h = bloom_slot(indices[e], deg);
        atomicOr(bloom + seg + (h >> 5), 1u << (h & 31));
    }
}

cudaError_t launch_build_bloom(const int64_t* indptr, const int32_t* indices, int U, uint32_t* bloom, cudaStream_t st)
{
    build_bloom_kernel<<<148 * 8, 256, 0, st>>>(indptr, indices, U, bloom);
    return cudaGetLastError();
}

cudaError_t launch_build_bitmap(const int64_t* indptr, const int32_t* indices, int U, uint32_t* bitmap, int words, cudaStream_t st)
{
    build_bitmap_kernel<<<148 * 4, 256, 0, st>>>(indptr, indices, U, bitmap, words);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// epoch-end reductions: assert_finite + reg_penalty (rankfm/_rankfm.pyx:95-116,328-336) in one pass over the tables
// out[0..5] = sum(w) of w_i, w_if, v_u, v_i, v_uf, v_if ; out[6..11] = sum(w^2) of the same
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_stats_kernel(const Tables T, double* __restrict__ out)
{
    double s[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const long long uq = (long long)T.Un * T.NQ, iq = (long long)T.I * T.NQ;      // owned user rows only
    for (long long e = tid; e < uq; e += nth) {
        const float4 v = ld_cg4(T.UT + (T.u0 + e / T.NQ) * T.ldu + 4 * (e % T.NQ));
        s[2] += (double)v.x + v.y + v.z + v.w; s2[2] += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long e = tid; e < iq; e += nth) {
        const float4 v = ld_cg4(T.IT + (e / T.NQ) * T.ldi + 4 * (e % T.NQ));
        s[3] += (double)v.x + v.y + v.z + v.w; s2[3] += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long e = tid; e < T.I; e += nth) {
        const float w = ld_cg1(T.IT + e * T.ldi + T.Fp);
        s[0] += w; s2[0] += (double)w * w;
    }
    if (T.GP) {
        // all six arrays count, active or not, like the reference's assert_finite / reg_penalty (:95-116); pads are zero
        const long long n_wif = T.gp_vuf, n_vuf = (long long)T.P * T.Fp, n_vif = (long long)T.Q * T.Fp;
        for (long long e = tid; e < n_wif; e += nth) { const float w = ld_cg1(T.GP + e); s[1] += w; s2[1] += (double)w * w; }
        for (long long e = tid; e < n_vuf; e += nth) { const float w = ld_cg1(T.GP + T.gp_vuf + e); s[4] += w; s2[4] += (double)w * w; }
        for (long long e = tid; e < n_vif; e += nth) { const float w = ld_cg1(T.GP + T.gp_vif + e); s[5] += w; s2[5] += (double)w * w; }
    }
    __shared__ double red[12][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double a = s[k], b = s2[k];
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
        if (lane == 0) { red[k][wid] = a; red[6 + k][wid] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double a = 0;
        for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, a);
    }
}

cudaError_t launch_weight_stats(const Tables& T, double* out12, int grid, cudaStream_t st)
{
    weight_stats_kernel<<<grid, 256, 0, st>>>(T, out12);
    return cudaGetLastError();
}

}  // namespace rfm
