/*
 * rankfm_oracle.c -- CPU restatement of the RankFM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 kernels in rankfm_b200/csrc/.  It is a plain, sequential,
 * strict-IEEE (no -ffast-math, no FMA contraction) C restatement of the algorithm the reference implements in
 * Cython; nothing under rankfm_b200/ links, loads or calls it.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker or the timed CPU baseline.
 *
 * Reference lines followed (all relative to /root/reference/):
 *   orc_mt_seed / orc_mt_next    rankfm/mt19937ar/mt19937ar.c:60-73, 105-140   (MT19937, seeded 1492 per _fit)
 *   orc_utility                  rankfm/_rankfm.pyx:48-89                      (compute_ui_utility)
 *   orc_fit                      rankfm/_rankfm.pyx:122-342                    (_fit: epoch loop, sampler, update)
 *   orc_predict                  rankfm/_rankfm.pyx:345-390                    (_predict)
 *   orc_scores_user              rankfm/_rankfm.pyx:440-441                    (_recommend's all-item scoring)
 * C-level types (double intermediates, integer division) follow the generated code the reference compiles to:
 *   eta invscaling  (double)lr / pow(epoch+1, exponent) -> float
 *   multiplier      log((long)(I-1) / (long)sampled) / log(I)  -> float   [integer quotient, cdivision=True]
 *   log-likelihood  ll = (float)(ll + log(1.0/(1.0+exp(-pu))))
 *   d_outer         (float)(1.0/(exp(pu)+1.0))
 *   update term     w += eta * (((sw*mult) * (d_outer*d)) - (2reg * w))   [this association]
 *
 * Pinning: tests/test_oracle.py checks this restatement against (i) the MT19937 known-answer vector for seed
 * 1492, (ii) golden vectors minted from the compiled reference (tests/golden/, generator committed beside them)
 * and (iii) the compiled reference itself (oracle/_ref) whenever it is present.
 *
 * Two extra "schedules" that the reference does not have are restated here so that the GPU production kernel
 * (Philox negatives, on-device Feistel shuffle) can be checked draw-for-draw in its serial configuration:
 *   sampler 1  = Philox4x32-10 keyed by (seed), counter (row.lo, epoch, attempt/4, row.hi); attempt a uses word a%4;
 *                item = (word * I) >> 32
 *   perms NULL = position r of epoch e maps to row orc_feistel_perm(r, N, seed, e)
 * Their definitions are the contract shared with rankfm_b200/csrc/rfm_rng.cuh.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------ */
/* MT19937 (mt19937ar.c:60-73,105-140) -- re-entrant: state lives in a struct instead of file statics           */
/* ------------------------------------------------------------------------------------------------------------ */
#define ORC_MT_N 624
#define ORC_MT_M 397

typedef struct {
    uint32_t s[ORC_MT_N];
    int pos;
} orc_mt;

void orc_mt_seed(orc_mt *g, uint32_t seed)
{
    g->s[0] = seed;
    for (int k = 1; k < ORC_MT_N; ++k) {
        uint32_t prev = g->s[k - 1];
        g->s[k] = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)k;
    }
    g->pos = ORC_MT_N;
}

static void orc_mt_twist(orc_mt *g)
{
    for (int k = 0; k < ORC_MT_N; ++k) {
        uint32_t y = (g->s[k] & 0x80000000u) | (g->s[(k + 1) % ORC_MT_N] & 0x7fffffffu);
        uint32_t v = g->s[(k + ORC_MT_M) % ORC_MT_N] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        g->s[k] = v;
    }
    g->pos = 0;
}

uint32_t orc_mt_next(orc_mt *g)
{
    if (g->pos >= ORC_MT_N) orc_mt_twist(g);
    uint32_t y = g->s[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* fill out[0..n) with the stream for `seed` -- used by the known-answer test */
void orc_mt_stream(uint32_t seed, uint32_t *out, int n)
{
    orc_mt g;
    orc_mt_seed(&g, seed);
    for (int k = 0; k < n; ++k) out[k] = orc_mt_next(&g);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 and the Feistel permutation: the production-mode RNG contract (shared with rfm_rng.cuh)        */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t orc_mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* bijection of [0,N): 4-round balanced Feistel network on an even number of bits, cycle-walked into range */
int64_t orc_feistel_perm(int64_t r, int64_t N, uint64_t seed, int epoch)
{
    int bits = 2;
    while (((int64_t)1 << bits) < N) ++bits;
    if (bits & 1) ++bits;
    const int half = bits / 2;
    const uint32_t mask = (uint32_t)(((uint64_t)1 << half) - 1);
    uint32_t key[4];
    for (int k = 0; k < 4; ++k)
        key[k] = orc_mix32((uint32_t)seed ^ orc_mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(epoch + 1) + 0x85EBCA6Bu * (uint32_t)(k + 1)));
    uint64_t x = (uint64_t)r;
    do {
        uint32_t L = (uint32_t)(x >> half), R = (uint32_t)x & mask;
        for (int k = 0; k < 4; ++k) {
            uint32_t f = orc_mix32(R ^ key[k]) & mask;
            uint32_t t = L ^ f;
            L = R; R = t;
        }
        x = ((uint64_t)L << half) | R;
    } while ((int64_t)x >= N);
    return (int64_t)x;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* pointwise FM utility (_rankfm.pyx:48-89): float32 accumulate, this exact order                               */
/* ------------------------------------------------------------------------------------------------------------ */
float orc_utility(int F, int P, int Q, const float *x_uf, const float *x_if, float w_i, const float *w_if,
                  const float *v_u, const float *v_i, const float *v_uf, const float *v_if, int x_uf_any, int x_if_any)
{
    float res = w_i;
    for (int f = 0; f < F; ++f) res += v_u[f] * v_i[f];
    if (x_uf_any) {
        for (int p = 0; p < P; ++p) {
            if (x_uf[p] == 0.0f) continue;
            for (int f = 0; f < F; ++f) res += x_uf[p] * (v_uf[(size_t)p * F + f] * v_i[f]);
        }
    }
    if (x_if_any) {
        for (int q = 0; q < Q; ++q) {
            if (x_if[q] == 0.0f) continue;
            res += x_if[q] * w_if[q];
            for (int f = 0; f < F; ++f) res += x_if[q] * (v_if[(size_t)q * F + f] * v_u[f]);
        }
    }
    return res;
}

static int orc_any_nonzero(const float *x, size_t n)
{
    for (size_t k = 0; k < n; ++k) if (x[k] != 0.0f) return 1;
    return 0;
}

/* membership of item in the user's sorted item list; same truth value as the reference's lsearch (:20-27) */
static int orc_member(int item, const int32_t *items, int n)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int md = lo + (hi - lo) / 2;
        if (items[md] == item) return 1;
        if (items[md] < item) lo = md + 1; else hi = md - 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* _fit (_rankfm.pyx:122-342)                                                                                   */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    /* data */
    const int32_t *interactions;   /* [N,2] */
    const float   *sample_weight;  /* [N]   */
    const int64_t *indptr;         /* [U+1]  CSR of user_items (sorted per user) */
    const int32_t *indices;        /* [nnz]  */
    const float   *x_uf;           /* [U,P] */
    const float   *x_if;           /* [I,Q] */
    /* weights, updated in place */
    float *w_i, *w_if, *v_u, *v_i, *v_uf, *v_if;
    int64_t N;
    int32_t U, I, P, Q, F;
    float alpha, beta, learning_rate, learning_exponent;
    int32_t schedule;              /* 0 constant, 1 invscaling */
    int32_t max_samples, epochs;
    /* order + sampler */
    const int32_t *perms;          /* [epochs,N] row order per epoch (the reference's np.random.shuffle), or NULL */
    int32_t sampler;               /* 0 = MT19937 seeded mt_seed (reference), 1 = Philox contract */
    uint32_t mt_seed;              /* 1492 in the reference (_rankfm.pyx:182) */
    uint64_t seed;                 /* Philox / Feistel key */
    int32_t epoch_offset;          /* added to the epoch index for Philox/Feistel keys only */
    int32_t max_rejects;           /* 0 = unbounded like the reference; else accept after this many rejections */
    /* outputs */
    float  *out_ll;                /* [epochs] float32 running-sum log-likelihood, as the reference prints it */
    int64_t *out_draws;            /* [epochs] total negatives evaluated (sum of `sampled`) */
    int32_t *out_neg;              /* optional [epochs,N]: chosen negative per position (debug / replay) */
} orc_fit_args;

int orc_fit(orc_fit_args *a)
{
    const int F = a->F, P = a->P, Q = a->Q, I = a->I;
    const int64_t N = a->N;
    const float MARGIN = 1.0f;
    const float d_reg_a = (float)(2.0 * a->alpha);
    const float d_reg_b = (float)(2.0 * a->beta);
    const int x_uf_any = orc_any_nonzero(a->x_uf, (size_t)a->U * P);
    const int x_if_any = orc_any_nonzero(a->x_if, (size_t)a->I * Q);
    orc_mt mt;
    orc_mt_seed(&mt, a->mt_seed);

    for (int epoch = 0; epoch < a->epochs; ++epoch) {
        float eta;
        if (a->schedule == 0) eta = a->learning_rate;
        else if (a->schedule == 1) eta = (float)(((double)a->learning_rate) / pow((double)(epoch + 1), (double)a->learning_exponent));
        else return 1;
        float ll = 0.0f;
        int64_t draws = 0;
        const int key_epoch = epoch + a->epoch_offset;

        for (int64_t r = 0; r < N; ++r) {
            const int64_t row = a->perms ? (int64_t)a->perms[(size_t)epoch * N + r] : orc_feistel_perm(r, N, a->seed, key_epoch);
            const int u = a->interactions[2 * row], i = a->interactions[2 * row + 1];
            const float sw = a->sample_weight[row];
            const float *xu = a->x_uf + (size_t)u * P;
            float *vu = a->v_u + (size_t)u * F;
            const int32_t *items = a->indices + a->indptr[u];
            const int n_items = (int)(a->indptr[u + 1] - a->indptr[u]);

            const float ut_ui = orc_utility(F, P, Q, xu, a->x_if + (size_t)i * Q, a->w_i[i], a->w_if, vu,
                                            a->v_i + (size_t)i * F, a->v_uf, a->v_if, x_uf_any, x_if_any);
            int min_index = -1;
            float min_pu = 1e6f;
            int sampled = 0;
            uint32_t attempt = 0;
            for (sampled = 1; sampled <= a->max_samples; ++sampled) {
                int j, rejects = 0;
                for (;;) {
                    uint32_t word;
                    if (a->sampler == 0) word = orc_mt_next(&mt);
                    else {
                        uint32_t blk[4];
                        orc_philox4x32((uint32_t)row, (uint32_t)key_epoch, attempt >> 2, (uint32_t)((uint64_t)row >> 32),
                                       (uint32_t)a->seed, (uint32_t)(a->seed >> 32), blk);
                        word = blk[attempt & 3u];
                        ++attempt;
                    }
                    j = a->sampler == 0 ? (int)(word % (uint32_t)I)                  /* genrand_int32() % I, _rankfm.pyx:251 */
                                        : (int)(((uint64_t)word * (uint64_t)(uint32_t)I) >> 32);  /* Philox contract: multiply-shift */
                    if (!orc_member(j, items, n_items)) break;
                    if (a->max_rejects > 0 && ++rejects >= a->max_rejects) break;
                }
                const float ut_uj = orc_utility(F, P, Q, xu, a->x_if + (size_t)j * Q, a->w_i[j], a->w_if, vu,
                                                a->v_i + (size_t)j * F, a->v_uf, a->v_if, x_uf_any, x_if_any);
                const float pu = ut_ui - ut_uj;
                if (pu < min_pu) { min_index = j; min_pu = pu; }
                if (pu < MARGIN) break;
            }
            if (sampled > a->max_samples) sampled = a->max_samples;   /* loop ran to completion: C `for` leaves max+1, Cython `range` leaves max */
            draws += sampled;
            const int j = min_index;
            if (j < 0) return 2;                                     /* every pairwise utility was NaN */
            const float pu = min_pu;
            const float multiplier = (float)(log((double)((long)(I - 1) / (long)sampled)) / log((double)I));
            ll = (float)((double)ll + log(1.0 / (1.0 + exp(-(double)pu))));
            const float d_outer = (float)(1.0 / (exp((double)pu) + 1.0));
            if (a->out_neg) a->out_neg[(size_t)epoch * N + r] = j;

            const float smul = sw * multiplier;
            float *vi = a->v_i + (size_t)i * F, *vj = a->v_i + (size_t)j * F;
            const float *xi = a->x_if + (size_t)i * Q, *xj = a->x_if + (size_t)j * Q;

            a->w_i[i] += eta * ((smul * (d_outer * 1.0f)) - (d_reg_a * a->w_i[i]));
            a->w_i[j] += eta * ((smul * (d_outer * -1.0f)) - (d_reg_a * a->w_i[j]));
            if (x_if_any) {
                for (int q = 0; q < Q; ++q) {
                    const float d = xi[q] - xj[q];
                    a->w_if[q] += eta * ((smul * (d_outer * d)) - (d_reg_b * a->w_if[q]));
                }
            }
            for (int f = 0; f < F; ++f) {
                float d_v_u = vi[f] - vj[f];
                float d_v_i = vu[f];
                float d_v_j = -vu[f];
                if (x_uf_any) {
                    for (int p = 0; p < P; ++p) {
                        d_v_i += a->v_uf[(size_t)p * F + f] * xu[p];
                        d_v_j -= a->v_uf[(size_t)p * F + f] * xu[p];
                    }
                }
                if (x_if_any) {
                    for (int q = 0; q < Q; ++q) d_v_u += a->v_if[(size_t)q * F + f] * (xi[q] - xj[q]);
                }
                vu[f] += eta * ((smul * (d_outer * d_v_u)) - (d_reg_a * vu[f]));
                vi[f] += eta * ((smul * (d_outer * d_v_i)) - (d_reg_a * vi[f]));
                vj[f] += eta * ((smul * (d_outer * d_v_j)) - (d_reg_a * vj[f]));
                if (x_uf_any) {
                    for (int p = 0; p < P; ++p) {
                        if (xu[p] == 0.0f) continue;
                        const float d = xu[p] * (vi[f] - vj[f]);
                        float *w = a->v_uf + (size_t)p * F + f;
                        *w += eta * ((smul * (d_outer * d)) - (d_reg_b * *w));
                    }
                }
                if (x_if_any) {
                    for (int q = 0; q < Q; ++q) {
                        if (xi[q] - xj[q] == 0.0f) continue;
                        const float d = (xi[q] - xj[q]) * vu[f];
                        float *w = a->v_if + (size_t)q * F + f;
                        *w += eta * ((smul * (d_outer * d)) - (d_reg_b * *w));
                    }
                }
            }
        }
        if (a->out_ll) a->out_ll[epoch] = ll;
        if (a->out_draws) a->out_draws[epoch] = draws;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* _predict (_rankfm.pyx:345-390): indexes arrive as float32, NaN = unknown id                                  */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_predict(const float *pairs, int64_t N, const float *x_uf, const float *x_if, const float *w_i, const float *w_if,
                 const float *v_u, const float *v_i, const float *v_uf, const float *v_if,
                 int U, int I, int P, int Q, int F, float *scores)
{
    const int x_uf_any = orc_any_nonzero(x_uf, (size_t)U * P);
    const int x_if_any = orc_any_nonzero(x_if, (size_t)I * Q);
    for (int64_t row = 0; row < N; ++row) {
        const float uf = pairs[2 * row], itf = pairs[2 * row + 1];
        if (isnan(uf) || isnan(itf)) { scores[row] = NAN; continue; }
        const int u = (int)uf, i = (int)itf;
        scores[row] = orc_utility(F, P, Q, x_uf + (size_t)u * P, x_if + (size_t)i * Q, w_i[i], w_if, v_u + (size_t)u * F,
                                  v_i + (size_t)i * F, v_uf, v_if, x_uf_any, x_if_any);
    }
}

/* all-item scores of one user (_rankfm.pyx:440-441); ranking (np.argsort, :444) stays in numpy in oracle.py */
void orc_scores_user(int u, const float *x_uf, const float *x_if, const float *w_i, const float *w_if,
                     const float *v_u, const float *v_i, const float *v_uf, const float *v_if,
                     int U, int I, int P, int Q, int F, float *scores)
{
    const int x_uf_any = orc_any_nonzero(x_uf, (size_t)U * P);
    const int x_if_any = orc_any_nonzero(x_if, (size_t)I * Q);
    for (int i = 0; i < I; ++i)
        scores[i] = orc_utility(F, P, Q, x_uf + (size_t)u * P, x_if + (size_t)i * Q, w_i[i], w_if, v_u + (size_t)u * F,
                                v_i + (size_t)i * F, v_uf, v_if, x_uf_any, x_if_any);
}
