// rfm_train.cu -- the SGD epoch kernel: B200 replacement of the per-epoch loop of `_fit`
// (rankfm/_rankfm.pyx:218-326: shuffle order, compute_ui_utility, WARP/BPR rejection sampler, gradient step).
//
// One lane group (G lanes) per positive (u,i); groups stride over the positions r of the epoch's permutation.
// HBM-bound gather/scatter: per positive one gather of the user's fat row, one of the positive item's, one per
// negative candidate; updates go out as vector reductions (REDG.E.ADD.F32x4) so concurrent groups never lose an
// update.  No tensor cores here by design (arithmetic intensity ~1 FLOP/B).
//
// Two schedules share this code (template/param switches, no second implementation):
//   parallel  Hogwild over the whole GPU, Philox negatives, Feistel (or host) order           -> production
//   serial    one group, positions strictly in order, optionally the reference's MT19937      -> exact replay of
//             sequential SGD, used to pin the kernel arithmetic against the oracle / the reference
#include "rfm_kernels.h"
#include "rfm_pair.cuh"
#include "rfm_rng.cuh"

namespace rfm {

template <int G, int QPL, bool FEAT, bool MT>
__global__ void __launch_bounds__(kTrainThreads) sgd_epoch_kernel(const TrainParams p)
{
    __shared__ MtState mt_smem;
    const Tables& T = p.T;
    constexpr int GPW = 32 / G;                       // groups per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane % G, gw = lane / G;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const bool serial = p.serial != 0;
    const long long stride = serial ? 1 : n_warps * GPW;

    if (MT) {   // serial launch is <<<1,32>>>: the warp owns the generator
        for (int k = lane; k < kMtN; k += 32) mt_smem.s[k] = p.mt->s[k];
        if (lane == 0) mt_smem.pos = p.mt->pos;
        __syncwarp();
    }

    double ll_acc = 0.0;
    long long draws_acc = 0;
    int bad = 0;

    for (long long base = serial ? 0 : warp_global * GPW; base < p.N; base += stride) {
        const long long r = base + (serial ? 0 : gw);
        const bool valid = r < p.N && (!serial || gw == 0);

        // ---- locate the observed (user, item, sample weight): _rankfm.pyx:233-236 ----
        long long row = 0;
        if (valid) row = p.perm ? (long long)__ldg(p.perm + r) : feistel_perm(p.feistel, r);
        int2 ui = make_int2(0, 0);
        float sw = 0.f;
        if (valid) { ui = __ldg(p.interactions + row); sw = __ldg(p.sample_weight + row); }
        const int u = ui.x, i = ui.y;

        UserCtx<QPL> uc;
        ItemRow<QPL> pos, cand, neg;
        load_user<G, QPL, FEAT>(T, u, valid, sub, uc);
        load_item<G, QPL, FEAT>(T, i, valid, sub, pos);
        long long seg = 0; int deg = 0;
        if (valid) { seg = __ldg(p.indptr + u); deg = (int)(__ldg(p.indptr + u + 1) - seg); }
        user_precompute<G, QPL, FEAT>(T, valid, sub, uc);
        const float ut_ui = utility<G, QPL, FEAT>(uc, pos);

        // ---- WARP / BPR sampling loop: _rankfm.pyx:244-264 ----
        int sampled = 0, min_j = -1;
        float min_pu = 1e6f;
        bool done = !valid;
        uint32_t attempt = 0;
        Philox4 blk = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < QPL; ++k) neg.v[k] = zero4();
        neg.x = zero4(); neg.w = 0.f;

        for (int s = 1; s <= p.max_samples; ++s) {
            if (!__any_sync(0xffffffffu, !done)) break;
            // rejection-sample an unobserved item: `while True: j = genrand_int32() % I` (:250-253)
            int j = 0, rejects = 0;
            bool need = !done;
            while (__any_sync(0xffffffffu, need)) {
                uint32_t word;
                if (MT) {
                    word = mt_next_warp(&mt_smem);
                } else {
                    if ((attempt & 3u) == 0u)
                        blk = philox4x32_10((uint32_t)row, p.epoch_key, attempt >> 2, (uint32_t)((unsigned long long)row >> 32), p.k0, p.k1);
                    const uint32_t c = attempt & 3u;
                    word = c == 0u ? blk.x : (c == 1u ? blk.y : (c == 2u ? blk.z : blk.w));
                    if (need) ++attempt;
                }
                const int cj = (int)(word % (uint32_t)T.I);
                if (need) load_item<G, QPL, FEAT>(T, cj, true, sub, cand);      // speculative: in flight during the search
                const bool member = group_member<G>(cj, p.indices + seg, deg, need, sub, gw);
                if (need && (!member || ++rejects >= p.max_rejects)) { j = cj; need = false; }
            }
            const float ut_uj = utility<G, QPL, FEAT>(uc, cand);
            const float pu = ut_ui - ut_uj;
            if (!done) {
                sampled = s;
                if (pu < min_pu) { min_pu = pu; min_j = j; neg = cand; }
                if (pu < 1.0f) done = true;                                      // MARGIN (:149,263)
            }
        }

        // ---- gradient step: _rankfm.pyx:267-326 ----
        const bool upd = valid && min_j >= 0;
        if (valid && min_j < 0) bad = 1;
        const float mult = upd ? __ldg(p.mult + sampled) : 0.f;                  // log((I-1)//sampled)/log(I), host table
        const float d_outer = (float)(1.0 / (exp((double)min_pu) + 1.0));
        const float smul = sw * mult;
        if (upd && sub == 0) {   // log sigma(pu) = -softplus(-pu), evaluated without cancellation (:270)
            ll_acc -= (double)(fmaxf(-min_pu, 0.f) + log1pf(__expf(-fabsf(min_pu))));
            draws_acc += sampled;
        }
        const int j = upd ? min_j : 0;
        if (p.trace && upd && sub == 0) { p.trace[2 * r] = min_j; p.trace[2 * r + 1] = sampled; }
        float* urow = T.UT + (size_t)u * T.ldu;
        float* irow = T.IT + (size_t)i * T.ldi;
        float* jrow = T.IT + (size_t)j * T.ldi;
        const float eta = p.eta, ra = p.reg_a, rb = p.reg_b;
#define RFM_G(d, w) (eta * ((smul * (d_outer * (d))) - (rb_or_ra * (w))))

        float4 dx = zero4();
        if (FEAT) { dx.x = pos.x.x - neg.x.x; dx.y = pos.x.y - neg.x.y; dx.z = pos.x.z - neg.x.z; dx.w = pos.x.w - neg.x.w; }

        // d u / d v_u = (v_i - v_j) + sum_q v_if[q] (x_if[i,q] - x_if[j,q])   (:292,303-305)
        float4 dvu[QPL];
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            dvu[k].x = pos.v[k].x - neg.v[k].x; dvu[k].y = pos.v[k].y - neg.v[k].y;
            dvu[k].z = pos.v[k].z - neg.v[k].z; dvu[k].w = pos.v[k].w - neg.v[k].w;
        }
        if (FEAT && T.x_if_any) {
            for (int q = 0; q < T.Q; ++q) {
                const float dxq = __shfl_sync(0xffffffffu, get4(dx, q & 3), q >> 2, G);
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    const int qq = sub + k * G;
                    if (upd && qq < T.NQ) {
                        const float4 w = ld_cg4(T.GP + T.gp_vif + (size_t)q * T.Fp + 4 * qq);
                        dvu[k].x += w.x * dxq; dvu[k].y += w.y * dxq; dvu[k].z += w.z * dxq; dvu[k].w += w.w * dxq;
                    }
                }
            }
        }

        float4 vu_new[QPL], dij_new[QPL];   // updated v_u and (v_i - v_j), needed by the feature-factor updates
        {
            const float rb_or_ra = ra;
            if (upd && sub == 0) {           // item biases (:279-280)
                red_add1(irow + T.Fp, RFM_G(1.0f, pos.w));
                red_add1(jrow + T.Fp, RFM_G(-1.0f, neg.w));
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
                    red_add4(urow + 4 * q, du);
                    red_add4(irow + 4 * q, di);
                    red_add4(jrow + 4 * q, dj);
                }
                if (FEAT) {
                    vu_new[k].x = uc.vu[k].x + du.x; vu_new[k].y = uc.vu[k].y + du.y;
                    vu_new[k].z = uc.vu[k].z + du.z; vu_new[k].w = uc.vu[k].w + du.w;
                    dij_new[k].x = (pos.v[k].x + di.x) - (neg.v[k].x + dj.x); dij_new[k].y = (pos.v[k].y + di.y) - (neg.v[k].y + dj.y);
                    dij_new[k].z = (pos.v[k].z + di.z) - (neg.v[k].z + dj.z); dij_new[k].w = (pos.v[k].w + di.w) - (neg.v[k].w + dj.w);
                }
            }
        }
        if (FEAT) {
            const float rb_or_ra = rb;
            if (T.x_if_any) {
                if (upd && 4 * sub < T.Qp) {                                   // w_if, every q (:283-286)
                    const float4 w = ld_cg4(T.GP + 4 * sub);
                    float4 d;
                    d.x = RFM_G(dx.x, w.x); d.y = RFM_G(dx.y, w.y); d.z = RFM_G(dx.z, w.z); d.w = RFM_G(dx.w, w.w);
                    red_add4(T.GP + 4 * sub, d);
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
                            float* wp = T.GP + T.gp_vuf + (size_t)pp * T.Fp + 4 * q;
                            const float4 w = ld_cg4(wp);
                            float4 d;
                            d.x = RFM_G(xp * dij_new[k].x, w.x); d.y = RFM_G(xp * dij_new[k].y, w.y);
                            d.z = RFM_G(xp * dij_new[k].z, w.z); d.w = RFM_G(xp * dij_new[k].w, w.w);
                            red_add4(wp, d);
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
                            float* wp = T.GP + T.gp_vif + (size_t)q * T.Fp + 4 * qq;
                            const float4 w = ld_cg4(wp);
                            float4 d;
                            d.x = RFM_G(dxq * vu_new[k].x, w.x); d.y = RFM_G(dxq * vu_new[k].y, w.y);
                            d.z = RFM_G(dxq * vu_new[k].z, w.z); d.w = RFM_G(dxq * vu_new[k].w, w.w);
                            red_add4(wp, d);
                        }
                    }
                }
            }
        }
#undef RFM_G
        if (serial) { __threadfence(); __syncwarp(); }   // next step must observe this step's reductions
    }

    // ---- epoch accumulators ----
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, off);
        draws_acc += __shfl_xor_sync(0xffffffffu, draws_acc, off);
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (lane == 0) {
        atomicAdd(&p.acc->ll, ll_acc);
        atomicAdd(reinterpret_cast<unsigned long long*>(&p.acc->draws), (unsigned long long)draws_acc);
        if (bad) atomicOr(&p.acc->bad, 1);
    }
    if (MT) {
        __syncwarp();
        for (int k = lane; k < kMtN; k += 32) p.mt->s[k] = mt_smem.s[k];
        if (lane == 0) p.mt->pos = mt_smem.pos;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launch: pick the group geometry from the widest row section
// ---------------------------------------------------------------------------------------------------------------
template <int G, int QPL>
static cudaError_t launch_gq(const TrainParams& p, int grid, cudaStream_t st)
{
    const bool feat = p.T.x_uf_any || p.T.x_if_any;
    const bool mt = p.mt != nullptr;
    const dim3 g(p.serial ? 1 : grid), b(p.serial ? 32 : kTrainThreads);
    if (feat) { if (mt) sgd_epoch_kernel<G, QPL, true, true><<<g, b, 0, st>>>(p); else sgd_epoch_kernel<G, QPL, true, false><<<g, b, 0, st>>>(p); }
    else      { if (mt) sgd_epoch_kernel<G, QPL, false, true><<<g, b, 0, st>>>(p); else sgd_epoch_kernel<G, QPL, false, false><<<g, b, 0, st>>>(p); }
    return cudaGetLastError();
}

template <int G, int QPL>
static int occ_gq(const TrainParams& p)
{
    const bool feat = p.T.x_uf_any || p.T.x_if_any;
    int n = 0;
    if (feat) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sgd_epoch_kernel<G, QPL, true, false>, kTrainThreads, 0);
    else      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sgd_epoch_kernel<G, QPL, false, false>, kTrainThreads, 0);
    return n;
}

int train_group_size(const Tables& T, int* qpl_out)
{
    const int widest = max(T.Fp, max(T.Pp, T.Qp)) / 4;   // quads
    int G = 4;
    while (G < 32 && G < widest) G <<= 1;
    int qpl = (T.NQ + G - 1) / G;
    if (qpl_out) *qpl_out = qpl;
    return G;
}

cudaError_t launch_sgd_epoch(const TrainParams& p, int grid, cudaStream_t st)
{
    int qpl = 1;
    const int G = train_group_size(p.T, &qpl);
    if (max(p.T.Pp, p.T.Qp) > 4 * G || qpl > 4) return cudaErrorInvalidValue;   // caller reports RFM_ERR_UNSUPPORTED
    switch (G) {
        case 4:  return launch_gq<4, 1>(p, grid, st);
        case 8:  return launch_gq<8, 1>(p, grid, st);
        case 16: return launch_gq<16, 1>(p, grid, st);
        default:
            if (qpl == 1) return launch_gq<32, 1>(p, grid, st);
            if (qpl == 2) return launch_gq<32, 2>(p, grid, st);
            return launch_gq<32, 4>(p, grid, st);
    }
}

int sgd_epoch_blocks_per_sm(const TrainParams& p)
{
    int qpl = 1;
    const int G = train_group_size(p.T, &qpl);
    switch (G) {
        case 4:  return occ_gq<4, 1>(p);
        case 8:  return occ_gq<8, 1>(p);
        case 16: return occ_gq<16, 1>(p);
        default: return qpl == 1 ? occ_gq<32, 1>(p) : (qpl == 2 ? occ_gq<32, 2>(p) : occ_gq<32, 4>(p));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// epoch-end reductions: assert_finite + reg_penalty (rankfm/_rankfm.pyx:95-116,328-336) in one pass over the tables
// out[0..5] = sum(w) of w_i, w_if, v_u, v_i, v_uf, v_if ; out[6..11] = sum(w^2) of the same
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_stats_kernel(const Tables T, double* __restrict__ out)
{
    double s[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const long long uq = (long long)T.U * T.NQ, iq = (long long)T.I * T.NQ;
    for (long long e = tid; e < uq; e += nth) {
        const float4 v = ld_cg4(T.UT + (e / T.NQ) * T.ldu + 4 * (e % T.NQ));
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
        const long long n_wif = T.Qp, n_vuf = (long long)T.P * T.Fp * (T.x_uf_any ? 1 : 0), n_vif = (long long)T.Q * T.Fp * (T.x_if_any ? 1 : 0);
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
