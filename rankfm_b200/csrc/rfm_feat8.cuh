// rfm_feat8.cuh -- side-feature math of the production SGD kernel specialised for at most 8 user and 8 item feature columns
// (BASELINE.json configs[2]: 8 + 8 dense side features).
//
// The generic code (rfm_pair.cuh: user_precompute, rfm_sgd.cuh: apply_update) walks P and Q with run-time loops and moves
// every feature value with a shuffle + select chain per column and loop: ~2,400 warp instructions per positive at F=64,
// P=Q=8 -- the kernel is bound by issue slots, not by HBM.  Here
//   * the loops over the columns are unrolled at compile time (kFeat8 = 8 slots, columns >= P / Q are warp-uniformly skipped);
//   * the feature values x_uf[u, :] and x_if[i, :] - x_if[j, :] are broadcast ONCE per positive into registers of every lane
//     (two LDS.128 broadcasts from the staged user row; eight shuffles for the item-feature difference) instead of once
//     per column and loop;
//   * the Q reductions of b[q] = w_if[q] + v_if[q].v_u become one multi-value reduce-scatter butterfly (8 + 4 shuffles
//     instead of 8 x log2(G));
//   * the chain updates of v_uf / v_if (`w += eta*(c*d - 2 beta w)`, _rankfm.pyx:313-326) are two FMAs per element on the
//     shared-memory copy: w <- w*(1 - eta*2beta) + (eta*c*x)*d.
// Same arithmetic as the generic production path up to float reassociation (the Hogwild schedule is not bit-reproducible
// anyway); the serial / replay kernels keep the reference's operand association and never come here.
#pragma once
#include "rfm_pair.cuh"

namespace rfm {

constexpr int kFeat8 = 8;

struct Feat8 {
    float xu[kFeat8];     // x_uf[u, p]            (0 beyond P)
    float dx[kFeat8];     // x_if[i, q] - x_if[j, q] (0 beyond Q); filled by feat8_item_diff
};

__device__ __forceinline__ void feat8_unpack(float (&dst)[kFeat8], const float4& a, const float4& b)
{
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
    dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}

// x_uf[u, :] from the staged user row in shared memory: every lane reads the same two quads (LDS broadcast)
__device__ __forceinline__ void feat8_user(const Tables& T, const float* s_user_row, bool valid, Feat8& f)
{
    const float4* s4 = reinterpret_cast<const float4*>(s_user_row) + T.NQ;
    const float4 a = (valid && T.Pp >= 4) ? s4[0] : zero4();
    const float4 b = (valid && T.Pp >= 8) ? s4[1] : zero4();
    feat8_unpack(f.xu, a, b);
}

// dx = x_if[i] - x_if[j] lives as quads on lanes sub = 0, 1 of the group: broadcast the eight values to every lane
template <int G>
__device__ __forceinline__ void feat8_item_diff(const float4& dxq, Feat8& f)
{
    f.dx[0] = __shfl_sync(0xffffffffu, dxq.x, 0, G); f.dx[1] = __shfl_sync(0xffffffffu, dxq.y, 0, G);
    f.dx[2] = __shfl_sync(0xffffffffu, dxq.z, 0, G); f.dx[3] = __shfl_sync(0xffffffffu, dxq.w, 0, G);
    f.dx[4] = __shfl_sync(0xffffffffu, dxq.x, 1, G); f.dx[5] = __shfl_sync(0xffffffffu, dxq.y, 1, G);
    f.dx[6] = __shfl_sync(0xffffffffu, dxq.z, 1, G); f.dx[7] = __shfl_sync(0xffffffffu, dxq.w, 1, G);
}

// Sum each of v[0..8) over the G lanes of a group with a reduce-scatter butterfly: after the three halving rounds a lane
// holds the total of ONE value, index feat8_owned_index(sub) (duplicated on the G/8 lanes that share those three bits).
// Needs G >= 8.  Every lane of the warp must call.
template <int G>
__device__ __forceinline__ float feat8_reduce_scatter(float (&v)[kFeat8], int sub)
{
    static_assert(G >= 8, "feat8 needs lane groups of at least 8");
    {
        const bool up = (sub & (G / 2)) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float keep = up ? v[k + 4] : v[k], send = up ? v[k] : v[k + 4];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, G / 2, G);
        }
    }
    {
        const bool up = (sub & (G / 4)) != 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float keep = up ? v[k + 2] : v[k], send = up ? v[k] : v[k + 2];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, G / 4, G);
        }
    }
    {
        const bool up = (sub & (G / 8)) != 0;
        const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, G / 8, G);
    }
    float r = v[0];
#pragma unroll
    for (int off = G / 16; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off, G);
    return r;
}
template <int G>
__device__ __forceinline__ int feat8_owner_lane(int q)        // a lane of the group that holds the total of value q
{
    return ((q >> 2) & 1) * (G / 2) + ((q >> 1) & 1) * (G / 4) + (q & 1) * (G / 8);
}

// a[] and b[] of rfm_pair.cuh (user_precompute) for P, Q <= 8; gp = this group's chain copy in shared memory
template <int G, int QPL>
__device__ __forceinline__ void user_precompute8(const Tables& T, const float* gp, bool valid, int sub, const Feat8& f, UserCtx<QPL>& c)
{
#pragma unroll
    for (int k = 0; k < QPL; ++k) c.a[k] = c.vu[k];
    c.b = zero4();
    if (T.x_uf_any) {
        const float* base = gp + T.gp_vuf + 4 * sub;
#pragma unroll
        for (int p = 0; p < kFeat8; ++p) {
            if (p < T.P) {                                                   // warp-uniform
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (valid && sub + k * G < T.NQ) {
                        const float4 w = *reinterpret_cast<const float4*>(base + (size_t)p * T.Fp + 4 * k * G);
                        c.a[k].x = fmaf(w.x, f.xu[p], c.a[k].x); c.a[k].y = fmaf(w.y, f.xu[p], c.a[k].y);
                        c.a[k].z = fmaf(w.z, f.xu[p], c.a[k].z); c.a[k].w = fmaf(w.w, f.xu[p], c.a[k].w);
                    }
                }
            }
        }
    }
    if (T.x_if_any) {
        const float* base = gp + T.gp_vif + 4 * sub;
        float part[kFeat8];
#pragma unroll
        for (int q = 0; q < kFeat8; ++q) {
            part[q] = 0.f;
            if (q < T.Q) {
#pragma unroll
                for (int k = 0; k < QPL; ++k)
                    if (valid && sub + k * G < T.NQ) part[q] = dot4(*reinterpret_cast<const float4*>(base + (size_t)q * T.Fp + 4 * k * G), c.vu[k], part[q]);
            }
        }
        const float mine = feat8_reduce_scatter<G>(part, sub);                // total of value feat8_owned(sub)
        // lanes sub = 0, 1 own the quads of b[]: gather their four totals, add w_if
        const int q0 = 4 * (sub & 1);
        float4 s;
        s.x = __shfl_sync(0xffffffffu, mine, feat8_owner_lane<G>(q0 + 0), G);
        s.y = __shfl_sync(0xffffffffu, mine, feat8_owner_lane<G>(q0 + 1), G);
        s.z = __shfl_sync(0xffffffffu, mine, feat8_owner_lane<G>(q0 + 2), G);
        s.w = __shfl_sync(0xffffffffu, mine, feat8_owner_lane<G>(q0 + 3), G);
        if (valid && 4 * sub < T.Qp) {
            const float4 w = *reinterpret_cast<const float4*>(gp + 4 * sub);
            c.b = make_float4(s.x + w.x, s.y + w.y, s.z + w.z, s.w + w.w);
        }
    }
}

// The feature-parameter part of one gradient step (rfm_sgd.cuh: apply_update, blocks "FEAT") for P, Q <= 8 on this
// group's private chain copy `gp` in shared memory.  dvu enters as v_i - v_j and leaves with the item-feature term added
// (_rankfm.pyx:303-305); feat8_update_chains runs after the row deltas are known (:313-326).
template <int G, int QPL>
__device__ __forceinline__ void feat8_dvu(const Tables& T, const float* gp, bool upd, int sub, const Feat8& f, float4 (&dvu)[QPL])
{
    if (!T.x_if_any) return;
    const float* base = gp + T.gp_vif + 4 * sub;
#pragma unroll
    for (int q = 0; q < kFeat8; ++q) {
        if (q < T.Q) {
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                if (upd && sub + k * G < T.NQ) {
                    const float4 w = *reinterpret_cast<const float4*>(base + (size_t)q * T.Fp + 4 * k * G);
                    dvu[k].x = fmaf(w.x, f.dx[q], dvu[k].x); dvu[k].y = fmaf(w.y, f.dx[q], dvu[k].y);
                    dvu[k].z = fmaf(w.z, f.dx[q], dvu[k].z); dvu[k].w = fmaf(w.w, f.dx[q], dvu[k].w);
                }
            }
        }
    }
}

// w <- w + eta*((sw*mult*d_outer) * x * d - 2 beta w)  ==  w*(1 - eta*2beta) + (ec*x)*d          ec = eta*sw*mult*d_outer
template <int G, int QPL>
__device__ __forceinline__ void feat8_update_chains(const Tables& T, float* gp, bool upd, int sub, const Feat8& f, const float4& dxq, float ec, float eta_rb,
                                                    const float4 (&vu_new)[QPL], const float4 (&dij_new)[QPL])
{
    const float keep = 1.0f - eta_rb;
    if (T.x_if_any && upd && 4 * sub < T.Qp) {                              // w_if, every q (:283-286)
        float4* wp = reinterpret_cast<float4*>(gp + 4 * sub);
        const float4 w = *wp;
        *wp = make_float4(fmaf(ec, dxq.x, w.x * keep), fmaf(ec, dxq.y, w.y * keep), fmaf(ec, dxq.z, w.z * keep), fmaf(ec, dxq.w, w.w * keep));
    }
    if (T.x_uf_any) {                                                       // v_uf[p] for x_uf[u,p] != 0 (:313-318)
        float* base = gp + T.gp_vuf + 4 * sub;
#pragma unroll
        for (int p = 0; p < kFeat8; ++p) {
            if (p < T.P) {
                const float cx = ec * f.xu[p];
                const bool nz = f.xu[p] != 0.0f;
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (upd && nz && sub + k * G < T.NQ) {
                        float4* wp = reinterpret_cast<float4*>(base + (size_t)p * T.Fp + 4 * k * G);
                        const float4 w = *wp;
                        *wp = make_float4(fmaf(cx, dij_new[k].x, w.x * keep), fmaf(cx, dij_new[k].y, w.y * keep),
                                          fmaf(cx, dij_new[k].z, w.z * keep), fmaf(cx, dij_new[k].w, w.w * keep));
                    }
                }
            }
        }
    }
    if (T.x_if_any) {                                                       // v_if[q] for dx[q] != 0 (:321-326)
        float* base = gp + T.gp_vif + 4 * sub;
#pragma unroll
        for (int q = 0; q < kFeat8; ++q) {
            if (q < T.Q) {
                const float cx = ec * f.dx[q];
                const bool nz = f.dx[q] != 0.0f;
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (upd && nz && sub + k * G < T.NQ) {
                        float4* wp = reinterpret_cast<float4*>(base + (size_t)q * T.Fp + 4 * k * G);
                        const float4 w = *wp;
                        *wp = make_float4(fmaf(cx, vu_new[k].x, w.x * keep), fmaf(cx, vu_new[k].y, w.y * keep),
                                          fmaf(cx, vu_new[k].z, w.z * keep), fmaf(cx, vu_new[k].w, w.w * keep));
                    }
                }
            }
        }
    }
}

// One chain per warp (TrainParams::gp_race): the chain advances by ONE group's update per warp step.  With racing plain
// stores every group computes and stores its update and all but the last writer's are lost -- G/32 of the work lands.
// Here the warp picks the winner up front (the highest valid group) and applies ITS update with all 32 lanes: lane l takes
// parameter row (l >> 2) of v_uf and of v_if and the quads (l & 3) + 4 j of that row.  The winner's scalars and vectors
// arrive by shuffle (2 + 1 + 4 for the scalars and the w_if quad, 8 per quad pair for dij_new / vu_new).  Same result as
// the race would give when that group wrote last; ~half the instructions (rows at a 256-byte stride cost a two-way bank
// conflict per quarter warp, still half the shared-memory wavefronts).  Needs 8 <= G < 32.  Every lane must call.
template <int G, int QPL>
__device__ __forceinline__ void feat8_update_chains_warp(const Tables& T, float* gp, bool upd, int sub, const Feat8& f, const float4& dxq, float ec, float eta_rb,
                                                         const float4 (&vu_new)[QPL], const float4 (&dij_new)[QPL])
{
    static_assert(G >= 8 && G < 32, "a warp-wide winner update needs several groups of at least 8 lanes");
    const int lane = threadIdx.x & 31;
    const unsigned lead = __ballot_sync(0xffffffffu, upd && sub == 0);
    if (lead == 0u) return;                                                 // warp-uniform: no group has an update
    const int L = (31 - __clz(lead)) / G * G;                               // first lane of the winner group
    // the winner's x_uf[u, row] / dx[row] for this lane's row: lane s of a group offers value s, this lane fetches value `row`
    float su = f.xu[0], sd = f.dx[0];
#pragma unroll
    for (int t = 1; t < kFeat8; ++t) { if ((sub & 7) == t) { su = f.xu[t]; sd = f.dx[t]; } }
    const int row = lane >> 2, qsub = lane & 3;
    const float xu_r = __shfl_sync(0xffffffffu, su, L + row), dx_r = __shfl_sync(0xffffffffu, sd, L + row);
    const float ec_w = __shfl_sync(0xffffffffu, ec, L);
    const float keep = 1.0f - eta_rb;
    {                                                                       // w_if, every q (:283-286): lanes 0, 1 take the two quads
        float4 d;
        d.x = __shfl_sync(0xffffffffu, dxq.x, L + (lane & 1)); d.y = __shfl_sync(0xffffffffu, dxq.y, L + (lane & 1));
        d.z = __shfl_sync(0xffffffffu, dxq.z, L + (lane & 1)); d.w = __shfl_sync(0xffffffffu, dxq.w, L + (lane & 1));
        if (T.x_if_any && lane < 2 && 4 * lane < T.Qp) {
            float4* wp = reinterpret_cast<float4*>(gp + 4 * lane);
            const float4 w = *wp;
            *wp = make_float4(fmaf(ec_w, d.x, w.x * keep), fmaf(ec_w, d.y, w.y * keep), fmaf(ec_w, d.z, w.z * keep), fmaf(ec_w, d.w, w.w * keep));
        }
    }
    const float cu = ec_w * xu_r, cd = ec_w * dx_r;
    const bool do_u = T.x_uf_any && row < T.P && xu_r != 0.0f;               // v_uf[p] for x_uf[u,p] != 0 (:313-318)
    const bool do_d = T.x_if_any && row < T.Q && dx_r != 0.0f;               // v_if[q] for dx[q] != 0 (:321-326)
    float* base_u = gp + T.gp_vuf + (size_t)row * T.Fp;
    float* base_d = gp + T.gp_vif + (size_t)row * T.Fp;
    constexpr int J = G * QPL / 4;                                          // quads per lane and row
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = (4 * j) / G;                                          // compile-time after unrolling: which of the winner's quads-per-lane
        const int src = L + qsub + (4 * j) % G;                             // ... held by which of its lanes
        const int q = qsub + 4 * j;                                         // quad of the row
        float4 dj, vn;
        dj.x = __shfl_sync(0xffffffffu, dij_new[k].x, src); dj.y = __shfl_sync(0xffffffffu, dij_new[k].y, src);
        dj.z = __shfl_sync(0xffffffffu, dij_new[k].z, src); dj.w = __shfl_sync(0xffffffffu, dij_new[k].w, src);
        vn.x = __shfl_sync(0xffffffffu, vu_new[k].x, src); vn.y = __shfl_sync(0xffffffffu, vu_new[k].y, src);
        vn.z = __shfl_sync(0xffffffffu, vu_new[k].z, src); vn.w = __shfl_sync(0xffffffffu, vu_new[k].w, src);
        if (do_u && q < T.NQ) {
            float4* wp = reinterpret_cast<float4*>(base_u + 4 * q);
            const float4 w = *wp;
            *wp = make_float4(fmaf(cu, dj.x, w.x * keep), fmaf(cu, dj.y, w.y * keep), fmaf(cu, dj.z, w.z * keep), fmaf(cu, dj.w, w.w * keep));
        }
        if (do_d && q < T.NQ) {
            float4* wp = reinterpret_cast<float4*>(base_d + 4 * q);
            const float4 w = *wp;
            *wp = make_float4(fmaf(cd, vn.x, w.x * keep), fmaf(cd, vn.y, w.y * keep), fmaf(cd, vn.z, w.z * keep), fmaf(cd, vn.w, w.w * keep));
        }
    }
    __syncwarp();                                                           // the next step's lanes read rows other lanes wrote
}

}  // namespace rfm
