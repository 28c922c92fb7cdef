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
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile): optimised, but no
 * FMA contraction and no fast-math, so the arithmetic stays the literal fp32 op sequence.  oracle/mups_oracle_tuned.c
 * is the tuned CPU implementation bench.py times beside it.
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
        /* sum of squares over the Gaussians: tf.nn.l2_normalize reduces with Eigen's tree / packet reduction (numpy:
         * pairwise), whose error stays ~1e-7; a sequential fp32 loop over thousands of nearly equal terms drifts
         * systematically (2.5e-5 at G = 4096), an artefact a port must not add -- accumulated in double, rounded once */
        double sq = 0.0;
        for (int g = 0; g < G; ++g) {
            float v = st[c * G + g];
            if (c >= 2 && c < 11) v = (masked ? 1.0f / sqrtf(w[g]) : 1.0f / ((float)P * sqrtf(w[g]))) * v;            /* :715 / :623 */
            else if (c >= 11) v = (masked ? 1.0f / sqrtf(2.0f * w[g]) : 1.0f / ((float)P * sqrtf(2.0f * w[g]))) * v;  /* :719 / :627 */
            v = v / npts;                      /* :728-730 */
            v = signed_sqrt(v);                /* :733-736 */
            st[c * G + g] = v;
            sq += (double)(v * v);
        }
        const float sqf = (float)sq;
        const float inv = 1.0f / sqrtf(sqf > 1e-12f ? sqf : 1e-12f);   /* tf.nn.l2_normalize, :739-741 */
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

/* float64 evaluation of get_3dmfv_n_est (tf_util.py:655-753) in the MuPS layout + a per-element forward-error bound
 * of an fp32 evaluation of the same formula.  For every (patch, Gaussian, channel):
 *
 *   raw statistic      u = reduce_n term_n                         (sum / max / min over the unmasked slots)
 *   fp32 error model   |fl(term_n) - term_n| <= eps (c0 + c1 (ss_ng + ss_min,n)) |term_n|,  eps = 2^-24
 *                      ss_ng  = squared standardised distance of point n to Gaussian g (the exp argument is
 *                               -ss/2: its rounding error is relative to ss, and exp turns it into a relative
 *                               error of the pdf); ss_min,n = the same for the Gaussian nearest to n, which
 *                               dominates the posterior's normaliser;
 *                      sums add the accumulation error  eps csum sqrt(m) sum_n |term_n|   (probabilistic bound,
 *                               Higham & Mary 2019: error of an m-term sum grows like sqrt(m) eps, not m eps)
 *   bound_u            = eps [ sum_n (c0 + c1 (ss_ng + ss_min,n) + csum sqrt(m)) |term_n| ]        for the 7 sum channels
 *                      = eps max_n (c0 + c1 (ss_ng + ss_min,n)) |term_n|                           for max / min
 *                        (for d_pi the -w / sqrt(w) term of every slot enters with 3 eps, and in the sum channel with the
 *                        accumulation factor as well: tf.reduce_sum adds m terms of size sqrt(w))
 *   y = u k / n_eff, by = bound_u k / n_eff + 2 eps |y|
 *   x = sign(y) sqrt|y|:  bx = by / sqrt|y| + eps |x|  if |y| > by,  else  2.5 sqrt(by)
 *   z = x / N, N = sqrt(max(sum_g x^2, 1e-12)):  bz = bx / N + |x| sqrt(sum_g bx^2) / N^2 + 3 eps |z|
 *
 * out / bound: [B, G, 20 S] doubles.  c0, c1, csum are passed in (tests state them).  masked == 0 evaluates get_3dmfv
 * (tf_util.py:578-652) instead: nothing masked, full-sigma prefactor, 1/P. */
int oracle_mups_f64(const float* points, const int32_t* n_eff, const float* w, const float* mu, const float* sigma,
                    int64_t B, int S, int P, int G, int masked, double c0, double c1, double csum, double* out, double* bound) {
    int err = 0;
    const double eps = ldexp(1.0, -24);
    const double pref0 = pow(2.0 * M_PI, 1.5);
#pragma omp parallel
    {
        double* e = (double*)malloc(sizeof(double) * (size_t)G * 2);          /* e_g, ss_g */
        double* u = (double*)malloc(sizeof(double) * 40 * (size_t)G);         /* u[20][G], bu[20][G] */
        if (!e || !u) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (int64_t b = 0; b < B; ++b)
                for (int s = 0; s < S; ++s) {
                    const float* pts = points + (b * S + s) * (int64_t)P * 3;
                    const int ne = masked ? n_eff[b * S + s] : P;       /* get_3dmfv: static P, nothing masked (:618-628) */
                    const int m = ne + 1 < P ? ne + 1 : P;
                    const int any_masked = m < P;
                    double* ss = e + G;
                    double* bu = u + 20 * (size_t)G;
                    for (int g = 0; g < G; ++g) {
                        for (int c = 0; c < 20; ++c) { u[c * G + g] = 0.0; bu[c * G + g] = 0.0; }
                        u[0 * G + g] = -INFINITY;
                        for (int k = 0; k < 3; ++k) {
                            u[(2 + k) * G + g] = -INFINITY; u[(5 + k) * G + g] = INFINITY;
                            u[(11 + k) * G + g] = -INFINITY; u[(14 + k) * G + g] = INFINITY;
                        }
                    }
                    const double acc = csum * sqrt((double)m);
                    for (int n = 0; n < m; ++n) {
                        double denom = 0.0, ssmin = INFINITY;
                        for (int g = 0; g < G; ++g) {
                            double q = 0.0;
                            for (int k = 0; k < 3; ++k) {
                                const double t = ((double)pts[3 * n + k] - (double)mu[3 * g + k]) / (double)sigma[3 * g + k];
                                q += t * t;
                            }
                            ss[g] = q;
                            if (q < ssmin) ssmin = q;
                            const double s0 = (double)sigma[3 * g];
                            const double vol = masked ? s0 * s0 * s0                      /* :687 sigma_0^D */
                                                      : s0 * (double)sigma[3 * g + 1] * (double)sigma[3 * g + 2];   /* MultivariateNormalDiag */
                            e[g] = (double)w[g] / (pref0 * vol) * exp(-0.5 * q);
                            denom += e[g];
                        }
                        for (int g = 0; g < G; ++g) {
                            const double Q = e[g] / denom;
                            const double rho = c0 + c1 * (ss[g] + ssmin);
                            const double rsw = 1.0 / sqrt((double)w[g]);
                            const double dpi = (Q - (double)w[g]) * rsw;
                            if (dpi > u[0 * G + g]) u[0 * G + g] = dpi;
                            { const double bb = rho * Q * rsw + 3.0 * (double)w[g] * rsw; if (bb > bu[0 * G + g]) bu[0 * G + g] = bb; }
                            u[1 * G + g] += dpi;
                            bu[1 * G + g] += (rho + acc) * Q * rsw + (3.0 + acc) * (double)w[g] * rsw;   /* the -w/sqrt(w) terms accumulate too */
                            for (int k = 0; k < 3; ++k) {
                                const double t = ((double)pts[3 * n + k] - (double)mu[3 * g + k]) / (double)sigma[3 * g + k];
                                const double dm = Q * t, ds = Q * (t * t - 1.0);
                                const double am = fabs(dm), as = Q * (t * t + 1.0);   /* |Q t^2| + |Q|: the cancelling operands */
                                if (dm > u[(2 + k) * G + g]) u[(2 + k) * G + g] = dm;
                                if (dm < u[(5 + k) * G + g]) u[(5 + k) * G + g] = dm;
                                u[(8 + k) * G + g] += dm;
                                if (rho * am > bu[(2 + k) * G + g]) bu[(2 + k) * G + g] = rho * am;
                                if (rho * am > bu[(5 + k) * G + g]) bu[(5 + k) * G + g] = rho * am;
                                bu[(8 + k) * G + g] += (rho + acc) * am;
                                if (ds > u[(11 + k) * G + g]) u[(11 + k) * G + g] = ds;
                                if (ds < u[(14 + k) * G + g]) u[(14 + k) * G + g] = ds;
                                u[(17 + k) * G + g] += ds;
                                if (rho * as > bu[(11 + k) * G + g]) bu[(11 + k) * G + g] = rho * as;
                                if (rho * as > bu[(14 + k) * G + g]) bu[(14 + k) * G + g] = rho * as;
                                bu[(17 + k) * G + g] += (rho + acc) * as;
                            }
                        }
                    }
                    for (int c = 0; c < 20; ++c) {
                        const int is_max = c == 0 || (c >= 2 && c < 5) || (c >= 11 && c < 14);
                        const int is_min = (c >= 5 && c < 8) || (c >= 14 && c < 17);
                        double sq = 0.0, sqb = 0.0;
                        for (int g = 0; g < G; ++g) {
                            double v = u[c * G + g], bv = eps * bu[c * G + g];
                            if (any_masked && is_max && v < 0.0) v = 0.0;     /* masked slots: exact zeros (:698,703) */
                            if (any_masked && is_min && v > 0.0) v = 0.0;
                            const double k = c < 2 ? 1.0 : (c < 11 ? 1.0 / sqrt((double)w[g]) : 1.0 / sqrt(2.0 * (double)w[g]));
                            const double y = v * k / (double)ne;
                            const double by = bv * k / (double)ne + 2.0 * eps * fabs(y);
                            const double x = y > 0.0 ? sqrt(y) : (y < 0.0 ? -sqrt(-y) : 0.0);
                            const double bx = fabs(y) > by ? by / sqrt(fabs(y)) + eps * fabs(x) : 2.5 * sqrt(by);
                            u[c * G + g] = x;
                            bu[c * G + g] = bx;
                            sq += x * x;
                            sqb += bx * bx;
                        }
                        const double N = sqrt(sq > 1e-12 ? sq : 1e-12);
                        for (int g = 0; g < G; ++g) {
                            const double x = u[c * G + g], z = x / N;
                            const int64_t o = (b * (int64_t)G + g) * 20 * S + s * 20 + c;
                            out[o] = z;
                            bound[o] = bu[c * G + g] / N + fabs(x) * sqrt(sqb) / (N * N) + 3.0 * eps * fabs(z);
                        }
                    }
                }
        }
        free(e);
        free(u);
    }
    return err;
}

/* The shared seeded selection key, see oracle/mups_oracle.py::selection_keys:
 * (a, b) = first two words of Philox4x32-10(counter = (center, scale, 0, 0), key = seed);
 * key(j) = (fmix32(j) ^ a) * (b | 1) mod 2^32.  out[i] = key of neighbour nbr[i]. */
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
    for (int64_t i = 0; i < n; ++i) out[i] = (fmix32((uint32_t)nbr[i]) ^ a) * b;
}

/* ---- half 1 after the kd-tree query: shared seeded selection + gather + centre + normalise (pcpnet_dataset.py:310-343)
 * for one radius, OpenMP over the queries.  nbr_flat / nbr_off: the neighbour lists of cKDTree.query_ball_point in CSR
 * form (any order).  Same results as oracle/mups_oracle.py::gather_patches (checked in tests/test_oracle.py). */
/* rearranges v so that v[0..k) are the k smallest (quickselect, median-of-three) */
static void select_smallest(uint64_t* v, int64_t n, int64_t k) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        uint64_t a = v[lo], b = v[mid], c = v[hi];
        const uint64_t pivot = a < b ? (b < c ? b : (a < c ? c : a)) : (a < c ? a : (b < c ? c : b));
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (v[i] < pivot) ++i;
            while (v[j] > pivot) --j;
            if (i <= j) { const uint64_t t = v[i]; v[i] = v[j]; v[j] = t; ++i; --j; }
        }
        if (k - 1 <= j) hi = j;
        else if (k - 1 >= i) lo = i;
        else return;
    }
}
static int cmp_i64(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int oracle_half1_gather(const float* pts, const int64_t* query_idx, int64_t B, float rad_f32, const int64_t* nbr_flat,
                        const int64_t* nbr_off, int P, int S, int scale, uint64_t seed, float* patches /* [B,S*P,3] */,
                        int32_t* n_eff /* [B,S] */) {
    int err = 0;
#pragma omp parallel
    {
        uint64_t* keyed = NULL;
        int64_t* sel = (int64_t*)malloc(sizeof(int64_t) * (size_t)P);
        size_t cap = 0;
        if (!sel) {
#pragma omp atomic write
            err = 1;
        }
#pragma omp for schedule(dynamic, 8)
        for (int64_t b = 0; b < B; ++b) {
            if (err) continue;
            const int64_t* nb = nbr_flat + nbr_off[b];
            const int64_t n = nbr_off[b + 1] - nbr_off[b];
            const int64_t c = query_idx[b];
            const int64_t take = n < P ? n : P;
            n_eff[b * S + scale] = (int32_t)take;
            float* out = patches + ((size_t)b * S + scale) * (size_t)P * 3;
            const int64_t* chosen = nb;
            if ((size_t)n > cap) {
                free(keyed);
                cap = (size_t)n * 2;
                keyed = (uint64_t*)malloc(sizeof(uint64_t) * cap);
                if (!keyed) {
#pragma omp atomic write
                    err = 1;
                    cap = 0;
                    continue;
                }
            }
            if (n > P) {
                /* Philox4x32-10(counter = (centre, scale, 0, 0), key = seed) -> salt (a, b | 1); key(j) = (fmix32(j) ^ a) * b */
                uint32_t c0 = (uint32_t)c, c1 = (uint32_t)scale, c2 = 0, c3 = 0;
                uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
                for (int r = 0; r < 10; ++r) {
                    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
                    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
                    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
                    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
                    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
                }
                const uint32_t sa = c0, sb = c1 | 1u;
                for (int64_t i = 0; i < n; ++i)
                    keyed[i] = ((uint64_t)((fmix32((uint32_t)nb[i]) ^ sa) * sb) << 32) | (uint32_t)nb[i];
                select_smallest(keyed, n, P);                                /* the P smallest (key, index) pairs */
                for (int64_t i = 0; i < P; ++i) sel[i] = (int64_t)(keyed[i] & 0xFFFFFFFFull);
                qsort(sel, (size_t)P, sizeof(int64_t), cmp_i64);
                chosen = sel;
            } else {
                int64_t* tmp = (int64_t*)keyed;
                memcpy(tmp, nb, sizeof(int64_t) * (size_t)n);
                qsort(tmp, (size_t)n, sizeof(int64_t), cmp_i64);
                chosen = tmp;
            }
            const float cx = pts[3 * c], cy = pts[3 * c + 1], cz = pts[3 * c + 2];
            for (int64_t i = 0; i < take; ++i) {
                const int64_t j = chosen[i];
                out[3 * i] = (pts[3 * j] - cx) / rad_f32;
                out[3 * i + 1] = (pts[3 * j + 1] - cy) / rad_f32;
                out[3 * i + 2] = (pts[3 * j + 2] - cz) / rad_f32;
            }
        }
        free(keyed);
        free(sel);
    }
    return err;
}
