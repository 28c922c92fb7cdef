/* Tuned CPU implementation of half 2 (get_3dmfv_n_est + MuPS assembly), the second CPU denominator of bench.py.
 *
 * TEST / BENCHMARK INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * oracle/mups_oracle.c is a literal port of utils/tf_util.py:655-753 (one divide, one powf and one expf per pair,
 * constants recomputed in the pair loop): faithful, and a weak baseline.  This file is what a competent CPU
 * implementation of the same arithmetic looks like: per-Gaussian constants hoisted (1/sigma, w * prefactor, 1/sqrt(w)),
 * structure-of-arrays Gaussians, reciprocal multiplies instead of divides, every inner loop a unit-stride loop over the
 * Gaussians that gcc vectorises (AVX2 / AVX-512, vector expf from libmvec under -ffast-math), OpenMP over the
 * (query, scale) patches.  It keeps the general (non-separable) algorithm of the reference -- 1 exp per pair -- so it is
 * a fair CPU counterpart of the 46-op / 1-exp algorithmic count of SURVEY.md 8(d).  Results agree with the literal
 * port to ~1e-5 (checked in tests/test_oracle.py); it is a timing baseline, not a parity oracle.
 *
 * Build: gcc -O3 -march=<native|x86-64-v3> -ffast-math -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_tuned_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#ifndef VL
#define VL 16        /* Gaussians per register block: one AVX-512 vector, two AVX2 vectors */
#endif

typedef struct {
    int G;
    float *mux, *muy, *muz, *isx, *isy, *isz, *cw, *w, *rsw, *rs2w;   /* [G] each */
} gmm_soa;

static int soa_init(gmm_soa* g, const float* w, const float* mu, const float* sigma, int G) {
    g->G = G;
    float* base = (float*)aligned_alloc(64, sizeof(float) * 10 * (size_t)((G + 15) & ~15));
    if (!base) return 1;
    const size_t st = (size_t)((G + 15) & ~15);
    g->mux = base; g->muy = base + st; g->muz = base + 2 * st; g->isx = base + 3 * st; g->isy = base + 4 * st;
    g->isz = base + 5 * st; g->cw = base + 6 * st; g->w = base + 7 * st; g->rsw = base + 8 * st; g->rs2w = base + 9 * st;
    const float two_pi_pow = powf((float)(2.0 * M_PI), 1.5f);
    for (int i = 0; i < G; ++i) {
        g->mux[i] = mu[3 * i]; g->muy[i] = mu[3 * i + 1]; g->muz[i] = mu[3 * i + 2];
        g->isx[i] = 1.0f / sigma[3 * i]; g->isy[i] = 1.0f / sigma[3 * i + 1]; g->isz[i] = 1.0f / sigma[3 * i + 2];
        g->cw[i] = w[i] / (two_pi_pow * sigma[3 * i] * sigma[3 * i] * sigma[3 * i]);   /* tf_util.py:687 prefactor x w (:700) */
        g->w[i] = w[i];
        g->rsw[i] = 1.0f / sqrtf(w[i]);
        g->rs2w[i] = 1.0f / sqrtf(2.0f * w[i]);
    }
    return 0;
}

/* one patch -> st[20][G] (channel-major); e: scratch [P][G padded to 16] + [P] */
static void fv_one_tuned(const float* pts, int P, int n_eff, const gmm_soa* gm, float* restrict st, float* restrict e) {
    const int G = gm->G;
    const int m = n_eff + 1 < P ? n_eff + 1 : P;
    const int any_masked = m < P;
    float* restrict pi_max = st;            float* restrict pi_sum = st + G;
    float* restrict mu_max[3] = {st + 2 * G, st + 3 * G, st + 4 * G};
    float* restrict mu_min[3] = {st + 5 * G, st + 6 * G, st + 7 * G};
    float* restrict mu_sum[3] = {st + 8 * G, st + 9 * G, st + 10 * G};
    float* restrict sg_max[3] = {st + 11 * G, st + 12 * G, st + 13 * G};
    float* restrict sg_min[3] = {st + 14 * G, st + 15 * G, st + 16 * G};
    float* restrict sg_sum[3] = {st + 17 * G, st + 18 * G, st + 19 * G};
    for (int g = 0; g < G; ++g) {
        pi_max[g] = -INFINITY; pi_sum[g] = 0.f;
        for (int k = 0; k < 3; ++k) {
            mu_max[k][g] = -INFINITY; mu_min[k][g] = INFINITY; mu_sum[k][g] = 0.f;
            sg_max[k][g] = -INFINITY; sg_min[k][g] = INFINITY; sg_sum[k][g] = 0.f;
        }
    }
    const float* restrict mux = gm->mux; const float* restrict muy = gm->muy; const float* restrict muz = gm->muz;
    const float* restrict isx = gm->isx; const float* restrict isy = gm->isy; const float* restrict isz = gm->isz;
    const float* restrict cw = gm->cw;
    const int Gp = (G + 15) & ~15;
    /* phase 1: unnormalised posteriors e[n][g] and 1 / sum_g per point (one exp per pair) */
    float* restrict inv = e + (size_t)P * Gp;
    for (int n = 0; n < m; ++n) {
        const float x = pts[3 * n], y = pts[3 * n + 1], z = pts[3 * n + 2];
        float* restrict en = e + (size_t)n * Gp;
        float denom = 0.f;
#pragma omp simd reduction(+ : denom)
        for (int g = 0; g < G; ++g) {
            const float tx = (x - mux[g]) * isx[g], ty = (y - muy[g]) * isy[g], tz = (z - muz[g]) * isz[g];
            const float v = cw[g] * expf(-0.5f * (tx * tx + ty * ty + tz * tz));
            en[g] = v;
            denom += v;
        }
        inv[n] = 1.0f / denom;
    }
    /* phase 2: a block of VL Gaussians keeps its 20 running reductions in vector registers while the points stream by */
    for (int g0 = 0; g0 < G; g0 += VL) {
        float bmx[VL], bmy[VL], bmz[VL], bix[VL], biy[VL], biz[VL];
        float a0[VL], a1[VL], amx[3][VL], amn[3][VL], asm_[3][VL], bmxs[3][VL], bmns[3][VL], bsms[3][VL];
        for (int j = 0; j < VL; ++j) {
            const int g = g0 + j < G ? g0 + j : G - 1;
            bmx[j] = mux[g]; bmy[j] = muy[g]; bmz[j] = muz[g]; bix[j] = isx[g]; biy[j] = isy[g]; biz[j] = isz[g];
            a0[j] = -INFINITY; a1[j] = 0.f;
            for (int k = 0; k < 3; ++k) {
                amx[k][j] = -INFINITY; amn[k][j] = INFINITY; asm_[k][j] = 0.f;
                bmxs[k][j] = -INFINITY; bmns[k][j] = INFINITY; bsms[k][j] = 0.f;
            }
        }
        for (int n = 0; n < m; ++n) {
            const float x = pts[3 * n], y = pts[3 * n + 1], z = pts[3 * n + 2], r = inv[n];
            const float* restrict en = e + (size_t)n * Gp + g0;
#pragma omp simd
            for (int j = 0; j < VL; ++j) {
                const float Q = en[j] * r;
                const float tx = (x - bmx[j]) * bix[j], ty = (y - bmy[j]) * biy[j], tz = (z - bmz[j]) * biz[j];
                a0[j] = fmaxf(a0[j], Q);            /* (Q - w)/sqrt(w) is monotone in Q: finished in the epilogue */
                a1[j] += Q;
                const float ax = Q * tx, ay = Q * ty, az = Q * tz;
                amx[0][j] = fmaxf(amx[0][j], ax); amn[0][j] = fminf(amn[0][j], ax); asm_[0][j] += ax;
                amx[1][j] = fmaxf(amx[1][j], ay); amn[1][j] = fminf(amn[1][j], ay); asm_[1][j] += ay;
                amx[2][j] = fmaxf(amx[2][j], az); amn[2][j] = fminf(amn[2][j], az); asm_[2][j] += az;
                const float bx = ax * tx - Q, by = ay * ty - Q, bz = az * tz - Q;
                bmxs[0][j] = fmaxf(bmxs[0][j], bx); bmns[0][j] = fminf(bmns[0][j], bx); bsms[0][j] += bx;
                bmxs[1][j] = fmaxf(bmxs[1][j], by); bmns[1][j] = fminf(bmns[1][j], by); bsms[1][j] += by;
                bmxs[2][j] = fmaxf(bmxs[2][j], bz); bmns[2][j] = fminf(bmns[2][j], bz); bsms[2][j] += bz;
            }
        }
        for (int j = 0; j < VL && g0 + j < G; ++j) {
            const int g = g0 + j;
            pi_max[g] = a0[j]; pi_sum[g] = a1[j];
            for (int k = 0; k < 3; ++k) {
                mu_max[k][g] = amx[k][j]; mu_min[k][g] = amn[k][j]; mu_sum[k][g] = asm_[k][j];
                sg_max[k][g] = bmxs[k][j]; sg_min[k][g] = bmns[k][j]; sg_sum[k][g] = bsms[k][j];
            }
        }
    }
    const float inv_n = 1.0f / (float)n_eff;
    for (int g = 0; g < G; ++g) {
        const float w = gm->w[g], rsw = gm->rsw[g];
        pi_max[g] = (pi_max[g] - w) * rsw;
        pi_sum[g] = (pi_sum[g] - (float)m * w) * rsw;
        if (any_masked && pi_max[g] < 0.f) pi_max[g] = 0.f;
    }
    for (int c = 0; c < 20; ++c) {
        float* restrict v = st + c * G;
        double sq = 0.0;                     /* tree-accurate like Eigen / numpy reductions (see mups_oracle.c) */
        for (int g = 0; g < G; ++g) {
            float t = v[g];
            if (c >= 2) {
                const int is_max = (c >= 2 && c < 5) || (c >= 11 && c < 14), is_min = (c >= 5 && c < 8) || (c >= 14 && c < 17);
                if (any_masked && is_max && t < 0.f) t = 0.f;
                if (any_masked && is_min && t > 0.f) t = 0.f;
                t *= c < 11 ? gm->rsw[g] : gm->rs2w[g];
            }
            t *= inv_n;
            t = t > 0.f ? sqrtf(t) : (t < 0.f ? -sqrtf(-t) : 0.f);
            v[g] = t;
            sq += (double)(t * t);
        }
        const float nrm = 1.0f / sqrtf((float)sq > 1e-12f ? (float)sq : 1e-12f);
        for (int g = 0; g < G; ++g) v[g] *= nrm;
    }
}

/* points [B,S*P,3]; n_eff [B,S]; out [B,G,20*S] (the MuPS layout, models/experts_n_est.py:59-76) */
int oracle_mups_tuned(const float* points, const int32_t* n_eff, const float* w, const float* mu, const float* sigma,
                      int64_t B, int S, int P, int G, float* out) {
    gmm_soa gm;
    if (soa_init(&gm, w, mu, sigma, G)) return 1;
    int err = 0;
#pragma omp parallel
    {
        float* e = (float*)aligned_alloc(64, sizeof(float) * ((size_t)P * ((G + 15) & ~15) + (size_t)((P + 15) & ~15)));
        float* fv = (float*)aligned_alloc(64, sizeof(float) * 20 * (size_t)((G + 15) & ~15));
        if (!e || !fv) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (int64_t b = 0; b < B; ++b)
                for (int s = 0; s < S; ++s) {
                    fv_one_tuned(points + (b * S + s) * (int64_t)P * 3, P, n_eff[b * S + s], &gm, fv, e);
                    float* o = out + b * (int64_t)G * 20 * S + s * 20;
                    for (int g = 0; g < G; ++g)
                        for (int c = 0; c < 20; ++c) o[(int64_t)g * 20 * S + c] = fv[c * G + g];
                }
        }
        free(e);
        free(fv);
    }
    free(gm.mux);
    return err;
}
