/* CPU oracle (plain C port) for half 2 of the MuPS hot path of sitzikbs/Nesti-Net.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product
 * library.  Used by tests/ (as a fast checker at sizes the numpy
 * transliteration is too slow for) and by bench.py's cpu_baseline /
 * --impl reference legs (kind "port": the thing timed on the host cores).
 *
 * PARITY UNPINNED: the reference computes these statistics with TensorFlow
 * 1.12 (utils/tf_util.py:655-753 get_3dmfv_n_est, :578-652 get_3dmfv), which
 * cannot be installed here and has no golden vectors.  This file restates that
 * arithmetic in fp32 and is itself checked against oracle/mups_oracle.py (the
 * op-by-op numpy transliteration) in tests/test_oracle.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static float signed_sqrt(float x) {           /* tf_util.py:733-736: sign(x)*pow(|x|,0.5) */
    if (x > 0.f) return sqrtf(x);
    if (x < 0.f) return -sqrtf(-x);
    return x;                                  /* tf.sign(0)=0; NaN propagates */
}

/* One patch [P,3] -> fv [20,G] (channel-major, i.e. flatten=True order).
 * masked != 0: get_3dmfv_n_est with n_eff = n_original_points (tf_util.py:655-753)
 * masked == 0: get_3dmfv (tf_util.py:578-652), static P, full-sigma prefactor.
 * scratch: 21*G floats + G floats. */
static void fv_one(const float* pts, int P, int n_eff, const float* w, const float* mu,
                   const float* sigma, int G, int masked, float* out, float* e) {
    float* st = out;                           /* [20][G] accumulators */
    const int m = masked ? (n_eff + 1 < P ? n_eff + 1 : P) : P;   /* unmasked slots: r > n_eff is masked (:696) */
    const int any_masked = m < P;
    for (int g = 0; g < G; ++g) {
        st[0 * G + g] = -INFINITY; st[1 * G + g] = 0.f;
        for (int k = 0; k < 3; ++k) {
            st[(2 + k) * G + g] = -INFINITY; st[(5 + k) * G + g] = INFINITY; st[(8 + k) * G + g] = 0.f;
            st[(11 + k) * G + g] = -INFINITY; st[(14 + k) * G + g] = INFINITY; st[(17 + k) * G + g] = 0.f;
        }
    }
    const float two_pi_pow = powf((float)(2.0 * M_PI), 1.5f);       /* :687 */
    const float log_2pi_term = (float)(0.5 * 3 * log(2.0 * M_PI));
    for (int n = 0; n < m; ++n) {
        const float x = pts[3 * n], y = pts[3 * n + 1], z = pts[3 * n + 2];
        float denom = 0.f;
        for (int g = 0; g < G; ++g) {
            const float tx = (x - mu[3 * g]) / sigma[3 * g];
            const float ty = (y - mu[3 * g + 1]) / sigma[3 * g + 1];
            const float tz = (z - mu[3 * g + 2]) / sigma[3 * g + 2];
            const float ss = (tx * tx + ty * ty) + tz * tz;
            float p;
            if (masked) {                      /* :687-688, prefactor uses sigma[:,0]^D only */
                const float pref = 1.0f / (two_pi_pow * powf(sigma[3 * g], 3.0f));
                p = pref * expf(-0.5f * ss);
            } else {                           /* MultivariateNormalDiag.prob (:606-608) */
                const float ls = (logf(sigma[3 * g]) + logf(sigma[3 * g + 1])) + logf(sigma[3 * g + 2]);
                p = expf((-0.5f * ss - ls) - log_2pi_term);
            }
            const float wp = p * w[g];         /* :700 */
            e[g] = wp;
            denom += wp;                       /* :701 reduce_sum over Gaussians */
        }
        for (int g = 0; g < G; ++g) {
            const float Q = e[g] / denom;      /* :701 */
            const float dpi = masked ? (Q - w[g]) / sqrtf(w[g])                 /* :710 */
                                     : (Q - w[g]) / (sqrtf(w[g]) * (float)P);   /* :618 */
            if (dpi > st[0 * G + g] || dpi != dpi) st[0 * G + g] = dpi;
            st[1 * G + g] += dpi;
            for (int k = 0; k < 3; ++k) {
                const float diff = pts[3 * n + k] - mu[3 * g + k];
                const float dmu = Q * diff / sigma[3 * g + k];                   /* :714 */
                const float t = diff / sigma[3 * g + k];
                const float dsg = Q * (powf(t, 2.0f) - 1.0f);                    /* :718 */
                if (dmu > st[(2 + k) * G + g] || dmu != dmu) st[(2 + k) * G + g] = dmu;
                if (dmu < st[(5 + k) * G + g] || dmu != dmu) st[(5 + k) * G + g] = dmu;
                st[(8 + k) * G + g] += dmu;
                if (dsg > st[(11 + k) * G + g] || dsg != dsg) st[(11 + k) * G + g] = dsg;
                if (dsg < st[(14 + k) * G + g] || dsg != dsg) st[(14 + k) * G + g] = dsg;
                st[(17 + k) * G + g] += dsg;
            }
        }
    }
    /* masked slots contribute exact zeros to every reduction (:698,703,710-720) */
    if (any_masked) {
        for (int g = 0; g < G; ++g) {
            if (st[0 * G + g] < 0.f) st[0 * G + g] = 0.f;
            for (int k = 0; k < 3; ++k) {
                if (st[(2 + k) * G + g] < 0.f) st[(2 + k) * G + g] = 0.f;
                if (st[(5 + k) * G + g] > 0.f) st[(5 + k) * G + g] = 0.f;
                if (st[(11 + k) * G + g] < 0.f) st[(11 + k) * G + g] = 0.f;
                if (st[(14 + k) * G + g] > 0.f) st[(14 + k) * G + g] = 0.f;
            }
        }
    }
    const float npts = masked ? (float)n_eff : 1.0f;   /* :722-730; get_3dmfv folds 1/P into the scale factors */
    for (int c = 0; c < 20; ++c) {
        float sq = 0.f;
        for (int g = 0; g < G; ++g) {
            float v = st[c * G + g];
            if (c >= 2 && c < 11) v = (masked ? 1.0f / sqrtf(w[g]) : 1.0f / ((float)P * sqrtf(w[g]))) * v;            /* :715 / :623 */
            else if (c >= 11) v = (masked ? 1.0f / sqrtf(2.0f * w[g]) : 1.0f / ((float)P * sqrtf(2.0f * w[g]))) * v;  /* :719 / :627 */
            v = v / npts;                      /* :728-730 */
            v = signed_sqrt(v);                /* :733-736 */
            st[c * G + g] = v;
            sq += v * v;
        }
        const float inv = 1.0f / sqrtf(sq > 1e-12f ? sq : 1e-12f);   /* tf.nn.l2_normalize, :739-741 */
        for (int g = 0; g < G; ++g) st[c * G + g] *= inv;
    }
}

/* points [B,P,3]; n_eff [B] (ignored when masked==0); out [B,20,G]. */
int oracle_3dmfv(const float* points, const int32_t* n_eff, const float* w, const float* mu,
                 const float* sigma, int64_t B, int P, int G, int masked, float* out) {
    int err = 0;
#pragma omp parallel
    {
        float* e = (float*)malloc(sizeof(float) * (size_t)G);
        if (!e) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t b = 0; b < B; ++b)
                fv_one(points + b * (int64_t)P * 3, P, masked ? n_eff[b] : P, w, mu, sigma, G, masked,
                       out + b * 20 * (int64_t)G, e);
            free(e);
        }
    }
    return err;
}

/* MuPS assembly, models/experts_n_est.py:59-76.
 * points [B,S*P,3]; n_eff [B,S]; out [B,res,res,res,20*S] with res^3 == G:
 * out[b, g, s*20 + c] = fv_s[b, c, g]. */
int oracle_mups(const float* points, const int32_t* n_eff, const float* w, const float* mu,
                const float* sigma, int64_t B, int S, int P, int G, float* out) {
    int err = 0;
#pragma omp parallel
    {
        float* e = (float*)malloc(sizeof(float) * (size_t)G);
        float* fv = (float*)malloc(sizeof(float) * 20 * (size_t)G);
        if (!e || !fv) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (int64_t b = 0; b < B; ++b)
                for (int s = 0; s < S; ++s) {
                    fv_one(points + (b * S + s) * (int64_t)P * 3, P, n_eff[b * S + s], w, mu, sigma, G, 1, fv, e);
                    float* o = out + b * (int64_t)G * 20 * S + s * 20;
                    for (int g = 0; g < G; ++g)
                        for (int c = 0; c < 20; ++c) o[(int64_t)g * 20 * S + c] = fv[c * G + g];
                }
        }
        free(e);
        free(fv);
    }
    return err;
}

/* The shared seeded selection key, see oracle/mups_oracle.py::selection_keys:
 * (a, b) = first two words of Philox4x32-10(counter = (center, scale, 0, 0), key = seed);
 * key(j) = fmix32((j ^ a) * (b | 1)).  out[i] = key of neighbour nbr[i]. */
static uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

void oracle_selection_keys(uint64_t seed, uint32_t center, uint32_t scale, const int32_t* nbr, int64_t n,
                           uint32_t* out) {
    uint32_t c0 = center, c1 = scale, c2 = 0, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t a = c0, b = c1 | 1u;
    for (int64_t i = 0; i < n; ++i) out[i] = fmix32(((uint32_t)nbr[i] ^ a) * b);
}
