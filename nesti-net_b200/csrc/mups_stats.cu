// K5: fused 3DmFV statistics -> MuPS layout.  Replaces the un-fused TF1 op chain of
// get_3dmfv_n_est (reference utils/tf_util.py:655-753), get_3dmfv (:578-652) and the MuPS
// assembly (models/experts_n_est.py:59-76).
//
// Two kernels, one CTA per (query, scale), FP32-issue bound (no tensor cores: the K=3 contraction
// is not a GEMM; HBM traffic is just the 80*G bytes written per (query, scale)):
//
//  * stats_separable_kernel -- the lattice GMM of get_3d_grid_gmm (tensor-product means, one sigma
//    per axis, uniform w) factorises: Q(n,(i,j,k)) = qx(n,i) qy(n,j) qz(n,k).  Per tile of points the
//    per-axis factors (q, q t, q (t^2-1)) are staged in shared memory (3*res exps per point instead
//    of res^3), then every thread owns 4 Gaussians along z and keeps their 20 running reductions in
//    registers, consuming two points per step: products and sums as packed f32x2, max/min folded
//    with 3-input FMNMX3.  Patches whose points leave the lattice's 5-sigma box (where the
//    reference's single exp would underflow differently) are pushed to a worklist and redone by
//    the general kernel, so the fast path never changes semantics.
//  * stats_general_kernel   -- arbitrary (w, mu, sigma): phase 1 computes every point's posterior
//    normaliser (thread per point, Gaussians broadcast from shared memory), phase 2 is thread per
//    Gaussian with the 20 reductions in registers.
//
// Both end with the same epilogue: masked-slot zeros, scale factors, /n_eff, signed square root,
// per-channel L2 norm over the Gaussians (CTA-wide reduction), direct store into
// [B, res, res, res, 20*S] (or channel-major).
#include <cstdlib>
#include <cooperative_groups.h>

#include "mups_common.cuh"

namespace mups {

namespace cg = cooperative_groups;

struct StatsArgs {
    const float4* A;
    const float4* Bv;
    const float4* C;
    int G;
    const float* patches;
    const int32_t* n_eff;
    int S, P;
    uint32_t flags;
    float* out;
    // separable lattice
    const float* axis_mu;      // [3][64]
    float isig[3];
    float w_uniform;
    int res[3];
    int shift[3];              // log2(res)
    float guard_lo[3], guard_hi[3];   // coordinates outside [lattice - 5 sigma, lattice + 5 sigma] -> general kernel
    // fallback worklist (separable kernel pushes, general kernel pops)
    int* worklist;             // [B*S] item ids
    int* work_count;           // [1]
    int use_worklist;          // general kernel: take items from the worklist
    int gmm_in_smem;           // general kernel: stage A / Bv in shared memory (else read them through L1/L2)
    // K6: patches never materialise -- the staging step gathers the selected neighbours from the index and centres /
    // normalises them itself (same __fsub_rn / __fdiv_rn as K4: bit-identical to reading the patch tensor)
    const float4* g_sorted;    // index->sorted, or nullptr: read `patches`
    const int32_t* g_pos_of;   // index->pos_of
    const int64_t* g_q;        // [B] centre indices
    const int32_t* g_pos;      // [B,S,P] positions in `sorted` of the selected neighbours, -1 beyond n_eff
    int64_t g_n;               // points in the index
    float g_rf[MUPS_MAX_SCALES];   // float32(r_s): the divisor of pcpnet_dataset.py:343
};

// One patch slot (query b, scale s, slot t) of the K6 path: (p - c) / float32(r) in IEEE fp32, zeros beyond n_eff.
__device__ __forceinline__ float3 gathered_point(const StatsArgs& a, int64_t b, int s, int t, const float4& c) {
    const int32_t ps = __ldg(a.g_pos + (b * a.S + s) * (int64_t)a.P + t);
    if (ps < 0) return make_float3(0.f, 0.f, 0.f);
    const float4 p = __ldg(a.g_sorted + ps);
    const float rf = a.g_rf[s];
    return make_float3(__fdiv_rn(__fsub_rn(p.x, c.x), rf), __fdiv_rn(__fsub_rn(p.y, c.y), rf), __fdiv_rn(__fsub_rn(p.z, c.z), rf));
}
__device__ __forceinline__ float4 gather_centre(const StatsArgs& a, int64_t b) {
    const int64_t q = a.g_q[b];
    if (q < 0 || q >= a.g_n) return make_float4(0.f, 0.f, 0.f, 0.f);      // invalid centre: the ball query left an empty row
    return __ldg(a.g_sorted + __ldg(a.g_pos_of + q));
}

typedef unsigned long long u64;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ float signed_sqrt(float v) {   // tf.sign(x) * tf.pow(tf.abs(x), 0.5)
    const float r = sqrtf(fabsf(v));
    return v < 0.f ? -r : (v > 0.f ? r : v);
}

constexpr float kNegHalfLog2e = -0.72134752044448170368f;   // -0.5 * log2(e)

// Shared per-Gaussian epilogue: v[0..19] hold the raw reductions over the unmasked slots
// (v[0] = max d_pi, v[1] = sum d_pi already in d_pi units).  Applies the masked slots' exact zeros
// (tf_util.py:698,703), the 1/sqrt(w), 1/sqrt(2w) factors (:715,:719), /n_eff (:728-730) and the
// signed square root (:733-736); accumulates the squares for the channel norms.
__device__ __forceinline__ void finalize_gaussian(float v[20], bool any_masked, float smu, float ssg, float npts,
                                                  bool valid, float sq[20]) {
    if (any_masked) {
        v[0] = fmaxf(v[0], 0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[2 + k] = fmaxf(v[2 + k], 0.f); v[5 + k] = fminf(v[5 + k], 0.f);
            v[11 + k] = fmaxf(v[11 + k], 0.f); v[14 + k] = fminf(v[14 + k], 0.f);
        }
    }
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        float x = v[c];
        if (c >= 2) x *= (c < 11 ? smu : ssg);
        x = signed_sqrt(x / npts);
        v[c] = x;
        if (valid) sq[c] = fmaf(x, x, sq[c]);
    }
}

// tf.nn.l2_normalize over the Gaussians per channel: inv_norm[c] = rsqrt(max(sum_g x^2, 1e-12)) (:739-741)
template <int NT>
__device__ __forceinline__ void channel_norms(const float sq[20], float* red /*[NT/32][20]*/, float* inv_norm /*[20]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        float x = sq[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp * 20 + c] = x;
    }
    __syncthreads();
    if (tid < 20) {
        float x = 0.f;
        for (int w = 0; w < NT / 32; ++w) x += red[w * 20 + tid];
        inv_norm[tid] = 1.0f / sqrtf(fmaxf(x, 1e-12f));
    }
    __syncthreads();
}

__device__ __forceinline__ void store_gaussian(const StatsArgs& a, int64_t b, int s, int g, const float v[20],
                                               const float* scale /* nullptr: raw */) {
    const int S = a.S, G = a.G;
    if (a.flags & MUPS_LAYOUT_CHANNEL) {
#pragma unroll
        for (int c = 0; c < 20; ++c) a.out[((b * S + s) * 20 + c) * (int64_t)G + g] = scale ? v[c] * scale[c] : v[c];
    } else {
        float4* o = reinterpret_cast<float4*>(a.out + ((b * G + g) * (int64_t)S + s) * 20);
#pragma unroll
        for (int c = 0; c < 20; c += 4)
            o[c >> 2] = scale ? make_float4(v[c] * scale[c], v[c + 1] * scale[c + 1], v[c + 2] * scale[c + 2], v[c + 3] * scale[c + 3])
                              : make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    }
}

__device__ __forceinline__ void rescale_gaussian(const StatsArgs& a, int64_t b, int s, int g, const float* scale) {
    const int S = a.S, G = a.G;
    if (a.flags & MUPS_LAYOUT_CHANNEL) {
#pragma unroll
        for (int c = 0; c < 20; ++c) a.out[((b * S + s) * 20 + c) * (int64_t)G + g] *= scale[c];
    } else {
        float* o = a.out + ((b * G + g) * (int64_t)S + s) * 20;
#pragma unroll
        for (int c = 0; c < 20; ++c) o[c] *= scale[c];
    }
}

// =====================================================================================================
// general kernel
// =====================================================================================================

// w_g * p_g(x) = 2^(c_g - 0.5*log2(e) * sum_k t_k^2), c_g = log2(w_g * prefactor_g)
__device__ __forceinline__ float weighted_pdf(float x, float y, float z, const float4& A, const float4& Bq, float cg) {
    const float tx = (x - A.x) * Bq.x, ty = (y - A.y) * Bq.y, tz = (z - A.z) * Bq.z;
    const float ss = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
    return ex2_approx(fmaf(ss, kNegHalfLog2e, cg));
}

template <int GPT, int NT>
__global__ void __launch_bounds__(NT) stats_general_kernel(const StatsArgs a) {
    extern __shared__ __align__(16) float4 smem4[];
    const int G = a.G, S = a.S, P = a.P;
    // the Gaussians are staged in shared memory when they fit (G <= ~6500); very fine lattices read them through L1
    const float4* gA = a.gmm_in_smem ? smem4 : a.A;
    const float4* gB = a.gmm_in_smem ? smem4 + G : a.Bv;
    float4* pt = smem4 + (a.gmm_in_smem ? 2 * G : 0);
    float* part = reinterpret_cast<float*>(pt + P);       // [max(NT, P)]
    float* red = part + (NT > P ? NT : P);                // [NT/32][20]
    float* inv_norm = red + (NT / 32) * 20;               // [20]

    const int tid = threadIdx.x;
    const bool masked = (a.flags & MUPS_FLAG_MASKED) != 0;
    if (a.gmm_in_smem)
        for (int i = tid; i < G; i += NT) { smem4[i] = __ldg(a.A + i); smem4[G + i] = __ldg(a.Bv + i); }

    const int n_items = a.use_worklist ? *a.work_count : (int)gridDim.x;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int item = a.use_worklist ? a.worklist[it] : it;
        const int64_t b = item / S;
        const int s = item % S;
        int n_eff = masked ? a.n_eff[b * S + s] : P;
        if (n_eff < 0) n_eff = 0;
        const int m = masked ? min(n_eff + 1, P) : P;     // slots with r > n_eff are masked (tf_util.py:696)
        const bool any_masked = m < P;
        if (a.g_sorted) {
            const float4 cq = gather_centre(a, b);
            for (int n = tid; n < m; n += NT) {
                const float3 v = gathered_point(a, b, s, n, cq);
                pt[n] = make_float4(v.x, v.y, v.z, 0.f);
            }
        } else {
            const float* src = a.patches + (b * S + s) * (int64_t)P * 3;
            for (int n = tid; n < m; n += NT)
                pt[n] = make_float4(__ldg(src + 3 * n), __ldg(src + 3 * n + 1), __ldg(src + 3 * n + 2), 0.f);
        }
        __syncthreads();

        // ---- phase 1: normaliser per point ------------------------------------------------------------
        {
            int PT = 1;
            while (PT < m && PT < NT) PT <<= 1;
            const int chunks = NT / PT;
            const int pl = tid & (PT - 1), ch = tid / PT;
            const int Gc = (G + chunks - 1) / chunks;
            const int g0 = ch * Gc, g1 = min(G, g0 + Gc);
            for (int n0 = pl; n0 < m; n0 += 2 * PT) {
                const int n1 = n0 + PT;
                const bool has1 = n1 < m;
                const float4 p0 = pt[n0];
                const float4 p1 = has1 ? pt[n1] : p0;
                float acc0 = 0.f, acc1 = 0.f;
                for (int g = g0; g < g1; ++g) {
                    const float4 A = gA[g], Bq = gB[g];
                    const float cg = masked ? A.w : Bq.w;
                    acc0 += weighted_pdf(p0.x, p0.y, p0.z, A, Bq, cg);
                    acc1 += weighted_pdf(p1.x, p1.y, p1.z, A, Bq, cg);
                }
                part[ch * m + n0] = acc0;
                if (has1) part[ch * m + n1] = acc1;
            }
            __syncthreads();
            for (int n = tid; n < m; n += NT) {
                float d = 0.f;
                for (int c = 0; c < chunks; ++c) d += part[c * m + n];
                pt[n].w = 1.0f / d;
            }
            __syncthreads();
        }

        // ---- phase 2: 20 reductions per Gaussian ------------------------------------------------------
        const float npts = masked ? (float)n_eff : 1.0f;                     // tf_util.py:722-730
        const float inv_static = masked ? 1.0f : 1.0f / (float)P;            // get_3dmfv folds 1/n_points into the scales
        const int tiles = (G + GPT * NT - 1) / (GPT * NT);
        float sq[20];
#pragma unroll
        for (int c = 0; c < 20; ++c) sq[c] = 0.f;
        float v[GPT][20];

        for (int tile = 0; tile < tiles; ++tile) {
            float mux[GPT], muy[GPT], muz[GPT], isx[GPT], isy[GPT], isz[GPT], cg[GPT], pis[GPT], pio[GPT];
#pragma unroll
            for (int i = 0; i < GPT; ++i) {
                const int g = (tile * GPT + i) * NT + tid;
                const int gc = g < G ? g : G - 1;
                const float4 A = gA[gc], Bq = gB[gc], Cq = __ldg(a.C + gc);
                mux[i] = A.x; muy[i] = A.y; muz[i] = A.z;
                isx[i] = Bq.x; isy[i] = Bq.y; isz[i] = Bq.z;
                cg[i] = masked ? A.w : Bq.w;
                pis[i] = Cq.y * inv_static;          // 1/sqrt(w) [/P]
                pio[i] = -Cq.x * pis[i];             // d_pi = (Q - w) * pis   (tf_util.py:710 / :618)
#pragma unroll
                for (int c = 0; c < 20; ++c) v[i][c] = 0.f;
                v[i][0] = -INFINITY;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    v[i][2 + k] = -INFINITY; v[i][5 + k] = INFINITY;
                    v[i][11 + k] = -INFINITY; v[i][14 + k] = INFINITY;
                }
            }
            for (int n = 0; n < m; ++n) {
                const float4 p = pt[n];
#pragma unroll
                for (int i = 0; i < GPT; ++i) {
                    const float tx = (p.x - mux[i]) * isx[i], ty = (p.y - muy[i]) * isy[i], tz = (p.z - muz[i]) * isz[i];
                    const float ss = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
                    const float Q = ex2_approx(fmaf(ss, kNegHalfLog2e, cg[i])) * p.w;
                    const float d = fmaf(Q, pis[i], pio[i]);
                    v[i][0] = fmaxf(v[i][0], d);
                    v[i][1] += d;
                    const float ax = Q * tx, ay = Q * ty, az = Q * tz;                        // d_mu  (:714)
                    v[i][2] = fmaxf(v[i][2], ax); v[i][3] = fmaxf(v[i][3], ay); v[i][4] = fmaxf(v[i][4], az);
                    v[i][5] = fminf(v[i][5], ax); v[i][6] = fminf(v[i][6], ay); v[i][7] = fminf(v[i][7], az);
                    v[i][8] += ax; v[i][9] += ay; v[i][10] += az;
                    const float bx = fmaf(ax, tx, -Q), by = fmaf(ay, ty, -Q), bz = fmaf(az, tz, -Q);   // d_sigma (:718)
                    v[i][11] = fmaxf(v[i][11], bx); v[i][12] = fmaxf(v[i][12], by); v[i][13] = fmaxf(v[i][13], bz);
                    v[i][14] = fminf(v[i][14], bx); v[i][15] = fminf(v[i][15], by); v[i][16] = fminf(v[i][16], bz);
                    v[i][17] += bx; v[i][18] += by; v[i][19] += bz;
                }
            }
#pragma unroll
            for (int i = 0; i < GPT; ++i) {
                const int g = (tile * GPT + i) * NT + tid;
                const bool valid = g < G;
                const float4 Cq = __ldg(a.C + (valid ? g : G - 1));
                finalize_gaussian(v[i], any_masked, Cq.y * inv_static, Cq.z * inv_static, npts, valid, sq);
                if (tiles > 1 && valid) store_gaussian(a, b, s, g, v[i], nullptr);   // raw; rescaled below
            }
        }

        channel_norms<NT>(sq, red, inv_norm);

        if (tiles == 1) {
#pragma unroll
            for (int i = 0; i < GPT; ++i) {
                const int g = i * NT + tid;
                if (g < G) store_gaussian(a, b, s, g, v[i], inv_norm);
            }
        } else {
            // each thread rescales the raw values it wrote itself (same thread: no fence needed)
            for (int tile = 0; tile < tiles; ++tile) {
#pragma unroll
                for (int i = 0; i < GPT; ++i) {
                    const int g = (tile * GPT + i) * NT + tid;
                    if (g < G) rescale_gaussian(a, b, s, g, inv_norm);
                }
            }
        }
        __syncthreads();   // pt / part / red are reused by the next item
    }
}

// =====================================================================================================
// separable (lattice) kernel
// =====================================================================================================

constexpr int kSepThreads = 128;
constexpr int kSepMinBlocks = 4;        // 128 registers per thread: 80 running reductions + operands; 4 CTAs per SM
constexpr int kSepKPT = 4;              // Gaussians per thread, consecutive along z
constexpr int kSepTilePoints = 128;     // points staged per tile (64 point pairs)
constexpr int kSepClusterTilePoints = 64;   // cluster variant: smaller tiles keep 4 CTAs per SM at 16^3

__device__ __forceinline__ float sqrt_approx(float x) {   // max relative error 2^-23 (PTX ISA)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Shared-memory factor tables of one tile, per axis a with n_a lattice points:
//   FA[a][pp][i] = float4(q(n0,i), q(n1,i), qt(n0,i), qt(n1,i))       pp = point pair (n0 = 2pp, n1 = 2pp+1)
//   FB[a][pp][i] = float2(q(t^2-1)(n0,i), q(t^2-1)(n1,i))
// A thread fetches both points' factors as packed f32x2 operands with one LDS.128 + one LDS.64 per
// axis entry; consecutive lattice indices are consecutive 16 B / 8 B words, so a warp's loads
// (<= 8 distinct j, <= 4 distinct i, one k-quad) are bank-conflict free.
// LOG_RES: the lattice is RES x RES x RES with RES = 1 << LOG_RES in {4, 8, 16, 32} (get_3d_grid_gmm is always
// called with [n, n, n]); other lattices take the general kernel.
// MODE: how the 7 products per (point, Gaussian) and their sums are issued (all variants measured within 5 % of
// each other: the loop is bound by operand delivery / dispatch, not by one pipe -- profiles/README.md):
//   0  packed FMUL2 products, sums as (s + lo) + hi
//   1  hybrid: Q t_z and Q (t_z^2 - 1) as scalar FMUL, the other five packed; sums as s + (lo + hi) -- the inner add
//      reads an (even, odd) register pair, so neither add has a register-bank conflict.  Default.
//   2  all products scalar, pairwise sums
//
// CL > 1 (16^3 lattice: 1024 thread-tasks = 8 groups of 128): a thread-block CLUSTER of CL CTAs shares one
// (query, scale).  CTA r owns task group r; the factor tables are staged once per cluster -- every CTA computes
// 1/CL of the entries and writes them into all CL shared memories through distributed shared memory -- and the
// per-channel sums of squares are exchanged the same way, so nothing is recomputed and nothing is re-read from HBM.
// STG 1 (small lattices, no cluster): the staging lane-task is one (point pair, axis) and walks the lattice axis
// serially -- no shuffle reductions, ~1/3 fewer instructions and much shorter dependency chains than one lane per
// (pair, axis, lattice index); table rows are padded to RES + 1 entries so that these strided writers stay
// bank-conflict free (the readers touch one row per step and do not care).
// ---- bulk-copy (TMA, 1-D) helpers of the two-items-per-CTA variant -------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_wait(uint32_t bar, uint32_t parity) {          // bounded: a stuck copy traps instead of hanging
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

// One (query, scale) item of the lattice kernel.  coords: shared-memory buffer of the patch; wait_bar != 0: the patch is
// being delivered by a bulk copy that completes on that mbarrier (two-items-per-CTA variant), else it is loaded here.
template <int MODE, int LOG_RES, int CL, int STG>
__device__ __forceinline__ void stats_separable_item(const StatsArgs& a, unsigned char* smem_raw, const int item, float* coords,
                                                     const uint32_t wait_bar) {
    constexpr int TILE = CL > 1 ? kSepClusterTilePoints : kSepTilePoints;
    constexpr int NT = kSepThreads, TPP = TILE / 2;
    constexpr bool PAIRSUM = MODE != 0;
    constexpr int RES = 1 << LOG_RES;
    constexpr int nx = RES, ny = RES, nz = RES;
    static_assert(STG == 0 || (CL == 1 && RES <= 8), "serial staging: small lattices, single CTA");
    constexpr int PITCH = STG ? RES + 1 : RES;   // entries per (axis, point pair) table row
    float4* FA[3];
    float2* FB[3];
    FA[0] = reinterpret_cast<float4*>(smem_raw);
    FA[1] = FA[0] + TPP * PITCH;
    FA[2] = FA[1] + TPP * PITCH;
    FB[0] = reinterpret_cast<float2*>(FA[2] + TPP * PITCH);
    FB[1] = FB[0] + TPP * PITCH;
    FB[2] = FB[1] + TPP * PITCH;
    float* lat = reinterpret_cast<float*>(FB[2] + TPP * PITCH);   // [3][64] lattice coordinates
    float* red = lat + 3 * 64;                                 // [NT/32][20]
    float* inv_norm = red + (NT / 32) * 20;                    // [20] (+12 pad)
    float* axis_par = inv_norm + 20;                           // [3][4]: 1/sigma, guard lo, guard hi
    float* cl_sq = inv_norm + 32;                              // [CL][20] per-CTA sums of squares (cluster exchange)
    __shared__ int s_fallback;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = CL > 1 ? cluster.block_rank() : 0u;

    const int tid = threadIdx.x;
    const int S = a.S, P = a.P;
    const int64_t b = item / S;
    const int s = item % S;
    const bool masked = (a.flags & MUPS_FLAG_MASKED) != 0;
    int n_eff = masked ? a.n_eff[b * S + s] : P;
    if (n_eff < 0) n_eff = 0;
    const int m = masked ? min(n_eff + 1, P) : P;       // slots with r > n_eff are masked (tf_util.py:696)
    const bool any_masked = m < P;
    for (int i = tid; i < 3 * 64; i += NT) lat[i] = __ldg(a.axis_mu + i);
    if (wait_bar) {
        bulk_wait(wait_bar, 0);                    // the patch was requested before the previous item's main loop
    } else if (a.g_sorted) {
        const float4 cq = gather_centre(a, b);
        for (int t = tid; t < m; t += NT) {
            const float3 v = gathered_point(a, b, s, t, cq);
            coords[3 * t] = v.x; coords[3 * t + 1] = v.y; coords[3 * t + 2] = v.z;
        }
    } else {
        const float* src = a.patches + (b * S + s) * (int64_t)P * 3;
        for (int i = tid; i < 3 * m; i += NT) coords[i] = __ldg(src + i);
    }
    if (tid < 3) { axis_par[4 * tid] = a.isig[tid]; axis_par[4 * tid + 1] = a.guard_lo[tid]; axis_par[4 * tid + 2] = a.guard_hi[tid]; }
    if (tid == 0) s_fallback = 0;

    const float w = a.w_uniform;
    const float rsw = 1.0f / sqrtf(w), rs2w = 1.0f / sqrtf(2.0f * w);
    const float inv_static = masked ? 1.0f : 1.0f / (float)P;           // get_3dmfv folds 1/n_points into the scales
    const float inv_npts = masked ? 1.0f / (float)n_eff : 1.0f;         // tf_util.py:722-730
    const float pis = rsw * inv_static, smu = rsw * inv_static, ssg = rs2w * inv_static;

    constexpr int nzq = nz / kSepKPT;
    constexpr int nxy = nx * ny;
    constexpr int tasks = nxy * nzq;
    constexpr int groups = (tasks + NT - 1) / NT;
    static_assert(CL == 1 || CL == groups, "a cluster covers all task groups of one (query, scale)");
    auto block_or_cluster_sync = [&]() {
        if (CL > 1) cluster.sync(); else __syncthreads();
    };
    float sq[20];
#pragma unroll
    for (int c = 0; c < 20; ++c) sq[c] = 0.f;
    float v[kSepKPT][20];
    block_or_cluster_sync();   // s_fallback is initialised in every CTA before anyone may raise it remotely

    for (int group = (CL > 1 ? (int)crank : 0); group < (CL > 1 ? (int)crank + 1 : groups); ++group) {
        // task -> (k-quad, i, j) with j fastest: a warp spans <= 8 j, <= 4 i and (nx*ny >= 32) one k-quad
        const int task = group * NT + tid;
        const bool valid = task < tasks;
        const int tk = valid ? task : 0;
        const int j = tk % ny, i = (tk / ny) % nx, kq = tk / nxy;
        const int k0 = kq * kSepKPT;

        float ss[kSepKPT][7], mx[kSepKPT][7], mn[kSepKPT][6];     // 28 sums, 28 max, 24 min
#pragma unroll
        for (int g = 0; g < kSepKPT; ++g) {
#pragma unroll
            for (int c = 0; c < 7; ++c) { ss[g][c] = 0.f; mx[g][c] = -INFINITY; }
#pragma unroll
            for (int c = 0; c < 6; ++c) mn[g][c] = INFINITY;
        }

        for (int tile0 = 0; tile0 < m; tile0 += TILE) {
            const int tile_pts = min(TILE, m - tile0);
            const int npairs = (tile_pts + 1) >> 1;
            if (STG) {
                // ---- stage the per-axis factors of this tile: one lane per (point pair, axis) ----
                const int total = 3 * npairs;
                const float* ctile = coords + 3 * tile0;
                const int last = tile_pts - 1;                                     // last real point of the tile
                bool bad = false;
                for (int t = tid; t < total; t += NT) {
                    const int ax = (t >= npairs) + (t >= 2 * npairs);              // axis-major: a warp's lanes share the axis
                    const int pp = t - ax * npairs;
                    const bool real1 = 2 * pp + 1 <= last;                          // the odd tail has no second point
                    const float c0 = ctile[6 * pp + ax];
                    const float c1 = ctile[3 * min(2 * pp + 1, last) + ax];
                    const float isg = axis_par[4 * ax];
                    const float4* lat4 = reinterpret_cast<const float4*>(lat + ax * 64);
                    float4* fa = FA[0] + ax * (TPP * PITCH) + pp * PITCH;
                    float2* fb = FB[0] + ax * (TPP * PITCH) + pp * PITCH;
                    // pass 1: unnormalised factors and t parked in the table row itself (keeps the register
                    // footprint of this phase small: the 80 running reductions stay live across it)
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int v = 0; v < RES / 4; ++v) {
                        const float4 m4 = lat4[v];
                        const float mu4[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float t0 = (c0 - mu4[u]) * isg, t1 = (c1 - mu4[u]) * isg;
                            const float e0 = ex2_approx(kNegHalfLog2e * t0 * t0), e1 = ex2_approx(kNegHalfLog2e * t1 * t1);
                            s0 += e0;
                            s1 += e1;
                            fa[4 * v + u] = make_float4(e0, e1, t0, t1);
                        }
                    }
                    const float r0 = __fdividef(1.0f, s0);
                    const float r1 = real1 ? __fdividef(1.0f, s1) : 0.f;            // a missing second point contributes exact zeros
                    // pass 2: normalise, q t and q (t^2 - 1)
#pragma unroll
                    for (int li = 0; li < RES; ++li) {
                        const float4 f = fa[li];
                        const float q0 = f.x * r0, q1 = f.y * r1;
                        const float a0 = q0 * f.z, a1 = q1 * f.w;
                        fa[li] = make_float4(q0, q1, a0, a1);
                        fb[li] = make_float2(fmaf(a0, f.z, -q0), fmaf(a1, f.w, -q1));
                    }
                    const float glo = axis_par[4 * ax + 1], ghi = axis_par[4 * ax + 2];
                    bad |= !(c0 >= glo && c0 <= ghi) || !(c1 >= glo && c1 <= ghi);   // outside the 5-sigma box (or NaN)
                }
                if (bad) s_fallback = 1;
            } else
            // ---- stage the per-axis factors of this tile: one lane per (point pair, axis, lattice index) ----
            {
                const int total = (npairs * 3) << LOG_RES;
                // a cluster splits the lane-tasks: CTA r stages [r * share, (r + 1) * share) for everyone
                const int share = CL > 1 ? ((total + CL * NT - 1) / (CL * NT)) * NT : total;
                const int first = CL > 1 ? (int)crank * share : 0;
                const int stop = min(total, first + share);
                const float* ctile = coords + 3 * tile0;
                const int last = tile_pts - 1;                                     // last real point of the tile
                bool bad = false;
#pragma unroll 3
                for (int base = first; base < stop; base += NT) {
                    const int idx = base + tid;
                    const int li = idx & (RES - 1);
                    const int r = min(idx, total - 1) >> LOG_RES;                   // r = 3 * pair + axis (idle lanes clamp)
                    const int pp = r / 3, ax = r - 3 * pp;
                    const bool real1 = 2 * pp + 1 <= last;                          // the odd tail has no second point
                    const float c0 = ctile[6 * pp + ax];
                    const float c1 = ctile[3 * min(2 * pp + 1, last) + ax];
                    const float mu_l = lat[ax * 64 + li], isg = axis_par[4 * ax];
                    const float t0 = (c0 - mu_l) * isg, t1 = (c1 - mu_l) * isg;
                    const float e0 = ex2_approx(kNegHalfLog2e * t0 * t0), e1 = ex2_approx(kNegHalfLog2e * t1 * t1);
                    float s0 = e0, s1 = e1;
#pragma unroll
                    for (int o = RES >> 1; o > 0; o >>= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    }
                    const float q0 = __fdividef(e0, s0);
                    const float q1 = real1 ? __fdividef(e1, s1) : 0.f;              // a missing second point contributes exact zeros
                    const float a0 = q0 * t0, a1 = q1 * t1;
                    if (idx < total) {
                        const int slot = ax * (TPP * RES) + (pp << LOG_RES) + li;   // FA[ax] / FB[ax] are contiguous
                        const float4 fa = make_float4(q0, q1, a0, a1);
                        const float2 fb = make_float2(fmaf(a0, t0, -q0), fmaf(a1, t1, -q1));
                        if (CL > 1) {
#pragma unroll
                            for (int r = 0; r < CL; ++r) {       // distributed shared memory: every CTA of the cluster
                                cluster.map_shared_rank(FA[0], r)[slot] = fa;
                                cluster.map_shared_rank(FB[0], r)[slot] = fb;
                            }
                        } else {
                            FA[0][slot] = fa;
                            FB[0][slot] = fb;
                        }
                    }
                    const float glo = axis_par[4 * ax + 1], ghi = axis_par[4 * ax + 2];
                    bad |= !(c0 >= glo && c0 <= ghi) || !(c1 >= glo && c1 <= ghi);   // outside the 5-sigma box (or NaN)
                }
                if (bad) {
                    if (CL > 1) {
                        for (int r = 0; r < CL; ++r) *cluster.map_shared_rank(&s_fallback, r) = 1;
                    } else {
                        s_fallback = 1;
                    }
                }
            }
            block_or_cluster_sync();
            if (s_fallback) break;

            // ---- 20 reductions x 4 Gaussians, two points per step -------------------------------------------
            if (valid) {
                const float4* pxa = FA[0] + i;
                const float4* pya = FA[1] + j;
                const float4* pza = FA[2] + k0;
                const float2* pxb = FB[0] + i;
                const float2* pyb = FB[1] + j;
                const float2* pzb = FB[2] + k0;
#pragma unroll 1
                for (int pp = 0; pp < npairs; ++pp, pxa += PITCH, pya += PITCH, pza += PITCH, pxb += PITCH, pyb += PITCH, pzb += PITCH) {
                    const float4 fx = *pxa, fy = *pya;
                    const float2 bx = *pxb, by = *pyb;
                    if (MODE == 2) {
                        // all scalar; products grouped so that consecutive multiplies share a source register
                        float u0[5], u1[5];      // qq, aq, qa, bq, qb of point 0 / point 1
                        u0[0] = fx.x * fy.x; u0[1] = fx.z * fy.x; u0[3] = bx.x * fy.x; u0[2] = fx.x * fy.z; u0[4] = fx.x * by.x;
                        u1[0] = fx.y * fy.y; u1[1] = fx.w * fy.y; u1[3] = bx.y * fy.y; u1[2] = fx.y * fy.w; u1[4] = fx.y * by.y;
#pragma unroll
                        for (int g = 0; g < kSepKPT; ++g) {
                            const float4 fz = pza[g];
                            const float2 bz = pzb[g];
                            float lo[7], hi[7];
                            lo[0] = u0[0] * fz.x; lo[3] = u0[0] * fz.z; lo[6] = u0[0] * bz.x;
                            lo[1] = u0[1] * fz.x; lo[2] = u0[2] * fz.x; lo[4] = u0[3] * fz.x; lo[5] = u0[4] * fz.x;
                            hi[0] = u1[0] * fz.y; hi[3] = u1[0] * fz.w; hi[6] = u1[0] * bz.y;
                            hi[1] = u1[1] * fz.y; hi[2] = u1[2] * fz.y; hi[4] = u1[3] * fz.y; hi[5] = u1[4] * fz.y;
#pragma unroll
                            for (int c = 0; c < 7; ++c) {
                                ss[g][c] += lo[c] + hi[c];
                                mx[g][c] = fmax3(mx[g][c], lo[c], hi[c]);
                                if (c > 0) mn[g][c - 1] = fmin3(mn[g][c - 1], lo[c], hi[c]);
                            }
                        }
                    } else {
                        const u64 qx = pack2(fx.x, fx.y), ax_ = pack2(fx.z, fx.w), bx_ = pack2(bx.x, bx.y);
                        const u64 qy = pack2(fy.x, fy.y), ay_ = pack2(fy.z, fy.w), by_ = pack2(by.x, by.y);
                        const u64 u_qq = mul2(qx, qy), u_aq = mul2(ax_, qy), u_qa = mul2(qx, ay_), u_bq = mul2(bx_, qy),
                                  u_qb = mul2(qx, by_);
                        float q0, q1;
                        unpack2(u_qq, q0, q1);
#pragma unroll
                        for (int g = 0; g < kSepKPT; ++g) {
                            const float4 fz = pza[g];
                            const float2 bz = pzb[g];
                            const u64 qz = pack2(fz.x, fz.y);
                            float lo[7], hi[7];
                            unpack2(mul2(u_qq, qz), lo[0], hi[0]);     // Q
                            unpack2(mul2(u_aq, qz), lo[1], hi[1]);     // Q t_x
                            unpack2(mul2(u_qa, qz), lo[2], hi[2]);     // Q t_y
                            unpack2(mul2(u_bq, qz), lo[4], hi[4]);     // Q (t_x^2 - 1)
                            unpack2(mul2(u_qb, qz), lo[5], hi[5]);     // Q (t_y^2 - 1)
                            if (MODE == 1) {
                                lo[3] = q0 * fz.z; hi[3] = q1 * fz.w;  // Q t_z
                                lo[6] = q0 * bz.x; hi[6] = q1 * bz.y;  // Q (t_z^2 - 1)
                            } else {
                                unpack2(mul2(u_qq, pack2(fz.z, fz.w)), lo[3], hi[3]);
                                unpack2(mul2(u_qq, pack2(bz.x, bz.y)), lo[6], hi[6]);
                            }
#pragma unroll
                            for (int c = 0; c < 7; ++c) {
                                ss[g][c] = PAIRSUM ? ss[g][c] + (lo[c] + hi[c]) : (ss[g][c] + lo[c]) + hi[c];
                                mx[g][c] = fmax3(mx[g][c], lo[c], hi[c]);
                                if (c > 0) mn[g][c - 1] = fmin3(mn[g][c - 1], lo[c], hi[c]);
                            }
                        }
                    }
                }
            }
            block_or_cluster_sync();   // the factor tables are rewritten by the next tile
        }
        if (s_fallback) break;

        // ---- per-Gaussian epilogue (same steps as finalize_gaussian, with fast reciprocal/sqrt) -------------
#pragma unroll
        for (int g = 0; g < kSepKPT; ++g) {
            const float* sums = ss[g];
            // d_pi = (Q - w)/sqrt(w): max and sum over the m unmasked slots (tf_util.py:710-712)
            v[g][0] = (mx[g][0] - w) * pis;
            v[g][1] = fmaf(-(float)m, w, sums[0]) * pis;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                v[g][2 + k] = mx[g][1 + k] * smu; v[g][5 + k] = mn[g][k] * smu; v[g][8 + k] = sums[1 + k] * smu;
                v[g][11 + k] = mx[g][4 + k] * ssg; v[g][14 + k] = mn[g][3 + k] * ssg; v[g][17 + k] = sums[4 + k] * ssg;
            }
            if (any_masked) {      // masked slots contribute exact zeros to max / min (tf_util.py:698,703)
                v[g][0] = fmaxf(v[g][0], 0.f);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    v[g][2 + k] = fmaxf(v[g][2 + k], 0.f); v[g][5 + k] = fminf(v[g][5 + k], 0.f);
                    v[g][11 + k] = fmaxf(v[g][11 + k], 0.f); v[g][14 + k] = fminf(v[g][14 + k], 0.f);
                }
            }
#pragma unroll
            for (int c = 0; c < 20; ++c) {
                const float x = v[g][c] * inv_npts;                         // :728-730
                const float r = copysignf(sqrt_approx(fabsf(x)), x);        // :733-736 (sign(0) = 0: sqrt(0) = 0)
                v[g][c] = r;
                if (valid) sq[c] = fmaf(r, r, sq[c]);
            }
            if (CL == 1 && groups > 1 && valid) store_gaussian(a, b, s, (i * ny + j) * nz + k0 + g, v[g], nullptr);
        }
    }

    if (s_fallback) {   // uniform over the CTA (and the cluster): everyone leaves, the general kernel redoes this item
        if (tid == 0 && crank == 0) a.worklist[atomicAdd(a.work_count, 1)] = item;
        return;
    }

    channel_norms<NT>(sq, red, inv_norm);
    if (CL > 1) {
        // inv_norm holds rsqrt(max(own sum, eps)); exchange the raw per-CTA sums instead and redo the norm over the cluster
        if (tid < 20) {
            float x = 0.f;
            for (int wq = 0; wq < NT / 32; ++wq) x += red[wq * 20 + tid];
            for (int r = 0; r < CL; ++r) cluster.map_shared_rank(cl_sq, r)[crank * 20 + tid] = x;
        }
        cluster.sync();
        if (tid < 20) {
            float x = 0.f;
            for (int r = 0; r < CL; ++r) x += cl_sq[r * 20 + tid];
            inv_norm[tid] = 1.0f / sqrtf(fmaxf(x, 1e-12f));
        }
        __syncthreads();
    }

    constexpr bool kWideFits = CL == 1 && groups == 1 &&
                               (size_t)TPP * 3 * PITCH * (sizeof(float4) + sizeof(float2)) >= (size_t)tasks * kSepKPT * 80;
    bool wide = false;
    if constexpr (kWideFits) wide = (a.flags & MUPS_FLAG_WIDE_STORES) && !(a.flags & MUPS_LAYOUT_CHANNEL);
    if (wide) {
        // out lives in a peer GPU's memory: 16-byte stores at a 320-byte stride travel badly over NVLink, so the
        // (query, scale) result is transposed through the (now idle) factor-table region and written as 80-byte runs
        constexpr int G = tasks * kSepKPT;
        float4* tile = reinterpret_cast<float4*>(smem_raw);
        if (tid < tasks) {
            const int j = tid % ny, i = (tid / ny) % nx, kq = tid / nxy;
#pragma unroll
            for (int g = 0; g < kSepKPT; ++g) {
                const int gi = (i * ny + j) * nz + kq * kSepKPT + g;
#pragma unroll
                for (int c = 0; c < 20; c += 4)
                    tile[gi * 5 + (c >> 2)] = make_float4(v[g][c] * inv_norm[c], v[g][c + 1] * inv_norm[c + 1],
                                                          v[g][c + 2] * inv_norm[c + 2], v[g][c + 3] * inv_norm[c + 3]);
            }
        }
        __syncthreads();
        float4* o = reinterpret_cast<float4*>(a.out);
        for (int f = tid; f < G * 5; f += NT) {
            const int gi = f / 5, c4 = f - 5 * gi;
            o[((b * G + gi) * (int64_t)S + s) * 5 + c4] = tile[f];
        }
    } else if (groups == 1 || CL > 1) {
        const int task = (CL > 1 ? (int)crank * NT : 0) + tid;
        if (task < tasks) {
            const int j = task % ny, i = (task / ny) % nx, kq = task / nxy;
#pragma unroll
            for (int g = 0; g < kSepKPT; ++g) store_gaussian(a, b, s, (i * ny + j) * nz + kq * kSepKPT + g, v[g], inv_norm);
        }
    } else {
        for (int group = 0; group < groups; ++group) {
            const int task = group * NT + tid;
            if (task < tasks) {
                const int j = task % ny, i = (task / ny) % nx, kq = task / nxy;
#pragma unroll
                for (int g = 0; g < kSepKPT; ++g) rescale_gaussian(a, b, s, (i * ny + j) * nz + kq * kSepKPT + g, inv_norm);
            }
        }
    }
}

template <int MODE, int LOG_RES, int CL, int STG>
__global__ void __launch_bounds__(kSepThreads, kSepMinBlocks) stats_separable_kernel(const StatsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = CL > 1 ? kSepClusterTilePoints : kSepTilePoints;
    constexpr int PITCH = STG ? (1 << LOG_RES) + 1 : (1 << LOG_RES);
    float* coords = reinterpret_cast<float*>(smem_raw + (size_t)(TILE / 2) * 3 * PITCH * (sizeof(float4) + sizeof(float2))) +
                    3 * 64 + (kSepThreads / 32) * 20 + 32 + (CL > 1 ? CL * 20 : 0);
    stats_separable_item<MODE, LOG_RES, CL, STG>(a, smem_raw, CL > 1 ? blockIdx.x / CL : blockIdx.x, coords, 0u);
}

// Variant 3 (mups_set_option "stats_variant"): TWO (query, scale) items per CTA; both patches are requested at kernel start
// with cp.async.bulk (TMA, 1-D) completing on an mbarrier each, so the second item's patch arrives under the first item's
// main loop and the lattice / per-axis parameters are staged once.  Patch tensor path only (the K6 gather cannot be a bulk copy).
template <int MODE, int LOG_RES>
__global__ void __launch_bounds__(kSepThreads, kSepMinBlocks) stats_separable_pair_kernel(const StatsArgs a, const int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PITCH = (1 << LOG_RES) + 1;
    float* coords = reinterpret_cast<float*>(smem_raw + (size_t)(kSepTilePoints / 2) * 3 * PITCH * (sizeof(float4) + sizeof(float2))) +
                    3 * 64 + (kSepThreads / 32) * 20 + 32;
    __shared__ __align__(8) unsigned long long bars[2];
    const int first = 2 * blockIdx.x;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr32(bars + k)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int k = 0; k < 2 && first + k < n_items; ++k) {
            const int item = first + k;
            const int64_t b = item / a.S;
            const int sc = item % a.S;
            int n_eff = (a.flags & MUPS_FLAG_MASKED) ? a.n_eff[b * a.S + sc] : a.P;
            if (n_eff < 0) n_eff = 0;
            const int m = (a.flags & MUPS_FLAG_MASKED) ? min(n_eff + 1, a.P) : a.P;
            const uint32_t bytes = (uint32_t)((12 * m + 15) & ~15);
            const float* src = a.patches + (b * a.S + sc) * (int64_t)a.P * 3;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr32(bars + k)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr32(coords + (size_t)k * 3 * a.P)), "l"(src), "r"(bytes), "r"(smem_addr32(bars + k)) : "memory");
        }
    }
    __syncthreads();
    for (int k = 0; k < 2 && first + k < n_items; ++k) {
        stats_separable_item<MODE, LOG_RES, 1, 1>(a, smem_raw, first + k, coords + (size_t)k * 3 * a.P, smem_addr32(bars + k));
        __syncthreads();                           // the tables, the norms and s_fallback are reused by the next item
    }
}

// =====================================================================================================
// launchers
// =====================================================================================================

template <int GPT, int NT>
static int launch_general(const StatsArgs& a_in, int64_t items, int grid, cudaStream_t st) {
    StatsArgs a = a_in;
    const size_t rest = sizeof(float4) * (size_t)a.P + sizeof(float) * ((NT > a.P ? NT : a.P) + (NT / 32) * 20 + 32);
    a.gmm_in_smem = sizeof(float4) * 2 * (size_t)a.G + rest <= 200 * 1024;
    const size_t smem = rest + (a.gmm_in_smem ? sizeof(float4) * 2 * (size_t)a.G : 0);
    if (smem > 220 * 1024) {
        set_error("3dmfv: G=%d, P=%d needs %zu bytes of shared memory", a.G, a.P, smem);
        return MUPS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        MUPS_CUDA_TRY(cudaFuncSetAttribute(stats_general_kernel<GPT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)items;
    stats_general_kernel<GPT, NT><<<(unsigned)grid, NT, smem, st>>>(a);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

static int dispatch_general(const StatsArgs& a, int64_t items, int grid, cudaStream_t st) {
    if (a.G <= 128) return launch_general<1, 128>(a, items, grid, st);
    if (a.G <= 256) return launch_general<1, 256>(a, items, grid, st);
    return launch_general<2, 256>(a, items, grid, st);
}

static bool pow2_in(int v, int lo, int hi) { return v >= lo && v <= hi && (v & (v - 1)) == 0; }

int launch_3dmfv(const mups_gmm* gmm, const float* patches, const int32_t* n_eff, int64_t B, int S, int P,
                 uint32_t flags, float* out, int* work /* [1 + B*S] ints of scratch, or nullptr */, const GatherSource* src,
                 cudaStream_t st) {
    StatsArgs a;
    a.g_sorted = nullptr; a.g_pos_of = nullptr; a.g_q = nullptr; a.g_pos = nullptr; a.g_n = 0;
    for (int k = 0; k < MUPS_MAX_SCALES; ++k) a.g_rf[k] = 1.0f;
    if (src) {
        a.g_sorted = src->index->sorted; a.g_pos_of = src->index->pos_of; a.g_n = src->index->n;
        a.g_q = src->q; a.g_pos = src->nbr_pos;
        for (int k = 0; k < S; ++k) a.g_rf[k] = (float)src->r_abs[k];
    }
    a.A = gmm->A; a.Bv = gmm->Bv; a.C = gmm->C; a.G = gmm->G;
    a.patches = patches; a.n_eff = n_eff; a.S = S; a.P = P; a.flags = flags & ~MUPS_FLAG_NO_FASTPATH; a.out = out;
    a.axis_mu = gmm->axis_mu;
    for (int k = 0; k < 3; ++k) {
        a.isig[k] = gmm->axis_isig[k]; a.res[k] = gmm->res[k];
        a.guard_lo[k] = gmm->guard_lo[k]; a.guard_hi[k] = gmm->guard_hi[k];
        a.shift[k] = 0;
        while ((1 << a.shift[k]) < a.res[k]) ++a.shift[k];
    }
    a.w_uniform = gmm->w_uniform;
    a.worklist = nullptr; a.work_count = nullptr; a.use_worklist = 0; a.gmm_in_smem = 1;
    if (B == 0) return MUPS_OK;
    const int64_t items = B * (int64_t)S;
    if (items > 0x7FFFFFFFll) {
        set_error("3dmfv: B*S = %lld exceeds the grid limit; split the batch", (long long)items);
        return MUPS_ERR_UNSUPPORTED;
    }
    // the lattice fast path: power-of-two axes (shuffle reductions), 4 | nz, even P (points are consumed in pairs)
    const bool fast = gmm->separable && !(flags & MUPS_FLAG_NO_FASTPATH) && work != nullptr && (P % 2 == 0) &&
                      pow2_in(gmm->res[0], 4, 32) && gmm->res[1] == gmm->res[0] && gmm->res[2] == gmm->res[0];
    if (!fast) return dispatch_general(a, items, (int)items, st);

    a.work_count = work;
    a.worklist = work + 1;
    MUPS_CUDA_TRY(cudaMemsetAsync(work, 0, sizeof(int), st));
    const int variant = g_stats_variant.load();      // 0 automatic; 1 round-1 loop; 2 all-scalar loop; 8 no cluster at 16^3
    const int log_res = a.shift[0];
    if (log_res == 4 && variant != 8 && items * 8 <= 0x7FFFFFFFll) {
        // 16^3 lattice: one thread-block cluster of 8 CTAs per (query, scale), tables shared through DSMEM
        constexpr int kCl = 8;
        auto kern = variant == 1 ? stats_separable_kernel<0, 4, kCl, 0> : stats_separable_kernel<1, 4, kCl, 0>;
        const size_t smem_cl = (size_t)(kSepClusterTilePoints / 2) * 3 * 16 * (sizeof(float4) + sizeof(float2)) +
                               sizeof(float) * (3 * 64 + (kSepThreads / 32) * 20 + 32 + 3 * (size_t)a.P + kCl * 20);
        MUPS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cl));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(items * kCl));
        cfg.blockDim = dim3(kSepThreads);
        cfg.dynamicSmemBytes = smem_cl;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MUPS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        // one CTA per (query, scale); res <= 8: serial staging with padded table rows (pitch res + 1)
        const int TPP = kSepTilePoints / 2;
        const bool serial = log_res <= 3 && variant != 1;
        const int pitch = a.res[0] + (serial ? 1 : 0);
        size_t smem = (size_t)TPP * 3 * pitch * (sizeof(float4) + sizeof(float2)) +
                      sizeof(float) * (3 * 64 + (kSepThreads / 32) * 20 + 32 + 3 * (size_t)a.P);
        // experiment knob (profiles/README.md): extra dynamic shared memory lowers the CTAs per SM of this kernel so that a
        // ball-query CTA of the next cloud can co-reside on the side stream
        static const int smem_pad = getenv("MUPS_STATS_SMEM_PAD") ? atoi(getenv("MUPS_STATS_SMEM_PAD")) : 0;
        if (smem_pad > 0 && smem + smem_pad <= 200 * 1024) smem += smem_pad;
#define MUPS_LAUNCH_SEP(MODE, LOG, STG)                                                                             \
    do {                                                                                                            \
        MUPS_CUDA_TRY(cudaFuncSetAttribute(stats_separable_kernel<MODE, LOG, 1, STG>,                               \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        stats_separable_kernel<MODE, LOG, 1, STG><<<(unsigned)items, kSepThreads, smem, st>>>(a);                   \
    } while (0)
        if (variant == 3 && log_res == 3 && patches != nullptr && a.g_sorted == nullptr && P % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(patches) & 15) == 0) {
            // two items per CTA, patches by cp.async.bulk + mbarrier (benchmarking: profiles/README.md)
            const size_t smem2 = smem + sizeof(float) * 3 * (size_t)a.P;
            MUPS_CUDA_TRY(cudaFuncSetAttribute(stats_separable_pair_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            stats_separable_pair_kernel<1, 3><<<(unsigned)((items + 1) / 2), kSepThreads, smem2, st>>>(a, (int)items);
        } else if (variant == 1) {                      // the round-1 kernel: packed products, sequential sums, shuffle staging
            if (log_res == 2) MUPS_LAUNCH_SEP(0, 2, 0);
            else if (log_res == 3) MUPS_LAUNCH_SEP(0, 3, 0);
            else if (log_res == 4) MUPS_LAUNCH_SEP(0, 4, 0);
            else MUPS_LAUNCH_SEP(0, 5, 0);
        } else if (variant == 2 && log_res == 3) MUPS_LAUNCH_SEP(2, 3, 1);
        else if (log_res == 2) MUPS_LAUNCH_SEP(1, 2, 1);
        else if (log_res == 3) MUPS_LAUNCH_SEP(1, 3, 1);
        else if (log_res == 4) MUPS_LAUNCH_SEP(1, 4, 0);
        else MUPS_LAUNCH_SEP(1, 5, 0);
#undef MUPS_LAUNCH_SEP
    }
    MUPS_CHECK_LAUNCH();
    // patches that left the lattice's 5-sigma box (none for real patches, which live in the unit ball)
    a.use_worklist = 1;
    const int grid = (int)(items < 2 * kNumSMs ? items : 2 * kNumSMs);
    return dispatch_general(a, items, grid, st);
}

}  // namespace mups
