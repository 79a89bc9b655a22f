// rfm_pair.cuh -- the pointwise FM utility (reference: compute_ui_utility, rankfm/_rankfm.pyx:48-89) on lane groups.
//
// The reference evaluates  u(u,i) = w_i + v_u.v_i + sum_p x_uf[u,p] (v_uf[p].v_i) + sum_q x_if[i,q] (w_if[q] + v_if[q].v_u)
// from scratch for every (u,i).  Here the user-only parts are hoisted once per user:
//     a[f] = v_u[f] + sum_p x_uf[u,p] v_uf[p,f]          ("effective user vector"; also d u/d v_i, _rankfm.pyx:293-300)
//     b[q] = w_if[q] + v_if[q].v_u
// so that  u(u,i) = w_i + a.v_i + b.x_if[i]  costs one gather of the item's fat row and ONE group reduction.
#pragma once
#include "rfm_common.cuh"

namespace rfm {

template <int QPL>
struct UserCtx {
    float4 vu[QPL];   // own quads of v_u[u]
    float4 a[QPL];    // own quads of the effective user vector
    float4 xu;        // own quad of x_uf[u]   (lane sub < Pp/4)
    float4 b;         // own quad of b[]       (lane sub < Qp/4)
};

template <int QPL>
struct ItemRow {
    float4 v[QPL];    // own quads of v_i[i]
    float4 x;         // own quad of x_if[i]   (lane sub < Qp/4)
    float  w;         // w_i[i] (every lane)
};

template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void load_user(const Tables& T, int u, bool valid, int sub, UserCtx<QPL>& c)
{
    const float* row = T.UT + (size_t)u * T.ldu;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int q = sub + k * G;
        c.vu[k] = (valid && q < T.NQ) ? ld_cg4(row + 4 * q) : zero4();
    }
    if (FEAT) c.xu = (valid && 4 * sub < T.Pp) ? __ldg(reinterpret_cast<const float4*>(row + T.Fp + 4 * sub)) : zero4();
}

// feature parameters live either in HBM (shared by everybody, read through L2) or in a warp-private shared-memory copy
template <bool GPS>
__device__ __forceinline__ float4 gp_ld4(const float* q) { return GPS ? *reinterpret_cast<const float4*>(q) : ld_cg4(q); }
template <bool GPS>
__device__ __forceinline__ float gp_ld1(const float* q) { return GPS ? *q : ld_cg1(q); }

// a[] and b[]: warp-uniform loops over P and Q; every lane of the warp must call.  gp = base of [w_if | v_uf | v_if]
template <int G, int QPL, bool FEAT, bool GPS = false>
__device__ __forceinline__ void user_precompute(const Tables& T, const float* gp, bool valid, int sub, UserCtx<QPL>& c)
{
#pragma unroll
    for (int k = 0; k < QPL; ++k) c.a[k] = c.vu[k];
    if (!FEAT) return;
    c.b = zero4();
    if (T.x_uf_any) {
        for (int p = 0; p < T.P; ++p) {
            const float xp = __shfl_sync(0xffffffffu, get4(c.xu, p & 3), p >> 2, G);
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                const int q = sub + k * G;
                if (valid && q < T.NQ) {
                    const float4 w = gp_ld4<GPS>(gp + T.gp_vuf + (size_t)p * T.Fp + 4 * q);
                    c.a[k].x += w.x * xp; c.a[k].y += w.y * xp; c.a[k].z += w.z * xp; c.a[k].w += w.w * xp;
                }
            }
        }
    }
    if (T.x_if_any) {
        for (int q = 0; q < T.Q; ++q) {
            float part = 0.f;
#pragma unroll
            for (int k = 0; k < QPL; ++k) {
                const int qq = sub + k * G;
                if (valid && qq < T.NQ) part = dot4(gp_ld4<GPS>(gp + T.gp_vif + (size_t)q * T.Fp + 4 * qq), c.vu[k], part);
            }
            part = group_sum<G>(part);
            if ((q >> 2) == sub) {
                const float s = part + (valid ? gp_ld1<GPS>(gp + q) : 0.f);
                const int cidx = q & 3;
                if (cidx == 0) c.b.x = s; else if (cidx == 1) c.b.y = s; else if (cidx == 2) c.b.z = s; else c.b.w = s;
            }
        }
    }
}

template <int G, int QPL, bool FEAT>
__device__ __forceinline__ void load_item(const Tables& T, int i, bool valid, int sub, ItemRow<QPL>& r)
{
    const float* row = T.IT + (size_t)i * T.ldi;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int q = sub + k * G;
        r.v[k] = (valid && q < T.NQ) ? ld_cg4(row + 4 * q) : zero4();
    }
    r.w = valid ? ld_cg1(row + T.Fp) : 0.f;
    if (FEAT) r.x = (valid && 4 * sub < T.Qp) ? __ldg(reinterpret_cast<const float4*>(row + T.Fp + 4 + 4 * sub)) : zero4();
}

// u(u,i); all lanes of the group return the same value; every lane of the warp must call
template <int G, int QPL, bool FEAT>
__device__ __forceinline__ float utility(const UserCtx<QPL>& c, const ItemRow<QPL>& r)
{
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < QPL; ++k) part = dot4(c.a[k], r.v[k], part);
    if (FEAT) part = dot4(c.b, r.x, part);
    return r.w + group_sum<G>(part);
}

}  // namespace rfm
