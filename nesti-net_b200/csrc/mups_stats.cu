// K5: fused 3DmFV statistics -> MuPS layout.  Replaces the un-fused TF1 op chain of
// get_3dmfv_n_est (reference utils/tf_util.py:655-753), get_3dmfv (:578-652) and the MuPS
// assembly (models/experts_n_est.py:59-76).
//
// One CTA per (query, scale).  The patch (<= P x 3 fp32) and the GMM are staged in shared memory;
// phase 1 computes every point's posterior normaliser sum_g w_g p_g(x) (thread per point, Gaussian
// parameters broadcast from shared memory); phase 2 is thread-per-Gaussian: the 20 running
// reductions (7 sums, 7 max, 6 min) of each owned Gaussian live in registers while the points
// are broadcast from shared memory; the epilogue applies /n_eff, the signed square root and the
// per-channel L2 norm over the Gaussians (a CTA-wide reduction) and stores straight into the
// [B, res, res, res, 20*S] tensor.  FP32-issue/MUFU bound (no tensor cores: the K=3 contraction
// is not a GEMM); HBM traffic is the 80*G bytes written per (query, scale).
#include "mups_common.cuh"

namespace mups {

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
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float signed_sqrt(float v) {   // tf.sign(x) * tf.pow(tf.abs(x), 0.5)
    const float r = sqrtf(fabsf(v));
    return v < 0.f ? -r : (v > 0.f ? r : v);
}

constexpr float kNegHalfLog2e = -0.72134752044448170368f;   // -0.5 * log2(e)

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
    float4* gA = smem4;
    float4* gB = gA + G;
    float4* pt = gB + G;
    float* part = reinterpret_cast<float*>(pt + P);       // [max(NT, P)]
    float* red = part + (NT > P ? NT : P);                // [NT/32][20] then inv[20]
    float* inv_norm = red + (NT / 32) * 20;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.x / S;
    const int s = blockIdx.x % S;
    const bool masked = (a.flags & MUPS_FLAG_MASKED) != 0;
    int n_eff = masked ? a.n_eff[b * S + s] : P;
    if (n_eff < 0) n_eff = 0;
    const int m = masked ? min(n_eff + 1, P) : P;         // slots with r > n_eff are masked (tf_util.py:696)
    const bool any_masked = m < P;

    for (int i = tid; i < G; i += NT) { gA[i] = __ldg(a.A + i); gB[i] = __ldg(a.Bv + i); }
    {
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
    const bool layout_channel = (a.flags & MUPS_LAYOUT_CHANNEL) != 0;
    float sq[20];
#pragma unroll
    for (int c = 0; c < 20; ++c) sq[c] = 0.f;
    float v[GPT][20];

    for (int tile = 0; tile < tiles; ++tile) {
        float mux[GPT], muy[GPT], muz[GPT], isx[GPT], isy[GPT], isz[GPT], cg[GPT], pis[GPT], pio[GPT];
        bool valid[GPT];
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = (tile * GPT + i) * NT + tid;
            valid[i] = g < G;
            const int gc = valid[i] ? g : G - 1;
            const float4 A = gA[gc], Bq = gB[gc], Cq = __ldg(a.C + gc);
            mux[i] = A.x; muy[i] = A.y; muz[i] = A.z;
            isx[i] = Bq.x; isy[i] = Bq.y; isz[i] = Bq.z;
            cg[i] = masked ? A.w : Bq.w;
            pis[i] = Cq.y * inv_static;          // 1/sqrt(w) [/P]
            pio[i] = -Cq.x * pis[i];             // d_pi = (Q - w) * pis
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
                const float ax = Q * tx, ay = Q * ty, az = Q * tz;
                v[i][2] = fmaxf(v[i][2], ax); v[i][3] = fmaxf(v[i][3], ay); v[i][4] = fmaxf(v[i][4], az);
                v[i][5] = fminf(v[i][5], ax); v[i][6] = fminf(v[i][6], ay); v[i][7] = fminf(v[i][7], az);
                v[i][8] += ax; v[i][9] += ay; v[i][10] += az;
                const float bx = fmaf(ax, tx, -Q), by = fmaf(ay, ty, -Q), bz = fmaf(az, tz, -Q);
                v[i][11] = fmaxf(v[i][11], bx); v[i][12] = fmaxf(v[i][12], by); v[i][13] = fmaxf(v[i][13], bz);
                v[i][14] = fminf(v[i][14], bx); v[i][15] = fminf(v[i][15], by); v[i][16] = fminf(v[i][16], bz);
                v[i][17] += bx; v[i][18] += by; v[i][19] += bz;
            }
        }
        // per-Gaussian epilogue: masked slots' zeros, scale factors, /n_eff, signed sqrt
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = (tile * GPT + i) * NT + tid;
            const int gc = valid[i] ? g : G - 1;
            const float4 Cq = __ldg(a.C + gc);
            const float smu = Cq.y * inv_static, ssg = Cq.z * inv_static;
            if (any_masked) {
                v[i][0] = fmaxf(v[i][0], 0.f);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    v[i][2 + k] = fmaxf(v[i][2 + k], 0.f); v[i][5 + k] = fminf(v[i][5 + k], 0.f);
                    v[i][11 + k] = fmaxf(v[i][11 + k], 0.f); v[i][14 + k] = fminf(v[i][14 + k], 0.f);
                }
            }
#pragma unroll
            for (int c = 0; c < 20; ++c) {
                float x = v[i][c];
                if (c >= 2) x *= (c < 11 ? smu : ssg);
                x = signed_sqrt(x / npts);
                v[i][c] = x;
                if (valid[i]) sq[c] = fmaf(x, x, sq[c]);
            }
            if (tiles > 1 && valid[i]) {   // raw values out; rescaled after the norm is known
                if (layout_channel) {
#pragma unroll
                    for (int c = 0; c < 20; ++c) a.out[((b * S + s) * 20 + c) * (int64_t)G + g] = v[i][c];
                } else {
                    float* o = a.out + ((b * G + g) * (int64_t)S + s) * 20;
#pragma unroll
                    for (int c = 0; c < 20; ++c) o[c] = v[i][c];
                }
            }
        }
    }

    // ---- channel-wise L2 norm over the Gaussians: x * rsqrt(max(sum x^2, 1e-12)) -------------------
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

    if (tiles == 1) {
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = i * NT + tid;
            if (g >= G) continue;
            if (layout_channel) {
#pragma unroll
                for (int c = 0; c < 20; ++c) a.out[((b * S + s) * 20 + c) * (int64_t)G + g] = v[i][c] * inv_norm[c];
            } else {
                float4* o = reinterpret_cast<float4*>(a.out + ((b * G + g) * (int64_t)S + s) * 20);
#pragma unroll
                for (int c = 0; c < 20; c += 4)
                    o[c >> 2] = make_float4(v[i][c] * inv_norm[c], v[i][c + 1] * inv_norm[c + 1],
                                            v[i][c + 2] * inv_norm[c + 2], v[i][c + 3] * inv_norm[c + 3]);
            }
        }
    } else {
        // each thread rescales the raw values it wrote itself (same thread: no fence needed)
        for (int tile = 0; tile < tiles; ++tile) {
#pragma unroll
            for (int i = 0; i < GPT; ++i) {
                const int g = (tile * GPT + i) * NT + tid;
                if (g >= G) continue;
                if (layout_channel) {
#pragma unroll
                    for (int c = 0; c < 20; ++c) a.out[((b * S + s) * 20 + c) * (int64_t)G + g] *= inv_norm[c];
                } else {
                    float* o = a.out + ((b * G + g) * (int64_t)S + s) * 20;
#pragma unroll
                    for (int c = 0; c < 20; ++c) o[c] *= inv_norm[c];
                }
            }
        }
    }
}

template <int GPT, int NT>
static int launch_general(const StatsArgs& a, int64_t B, cudaStream_t st) {
    const size_t smem = sizeof(float4) * (2 * (size_t)a.G + a.P) + sizeof(float) * ((NT > a.P ? NT : a.P) + (NT / 32) * 20 + 32);
    if (smem > 220 * 1024) {
        set_error("3dmfv: G=%d, P=%d needs %zu bytes of shared memory", a.G, a.P, smem);
        return MUPS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        MUPS_CUDA_TRY(cudaFuncSetAttribute(stats_general_kernel<GPT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stats_general_kernel<GPT, NT><<<(unsigned)(B * a.S), NT, smem, st>>>(a);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

int launch_3dmfv(const mups_gmm* gmm, const float* patches, const int32_t* n_eff, int64_t B, int S, int P,
                 uint32_t flags, float* out, cudaStream_t st) {
    StatsArgs a;
    a.A = gmm->A; a.Bv = gmm->Bv; a.C = gmm->C; a.G = gmm->G;
    a.patches = patches; a.n_eff = n_eff; a.S = S; a.P = P; a.flags = flags; a.out = out;
    if (B == 0) return MUPS_OK;
    if (B * (int64_t)S > 0x7FFFFFFFll) {
        set_error("3dmfv: B*S = %lld exceeds the grid limit; split the batch", (long long)(B * S));
        return MUPS_ERR_UNSUPPORTED;
    }
    if (a.G <= 128) return launch_general<1, 128>(a, B, st);
    if (a.G <= 256) return launch_general<1, 256>(a, B, st);
    return launch_general<2, 256>(a, B, st);
}

}  // namespace mups
