// Upper-bound micro-benchmarks for two K5 (stats_separable_kernel) variants that round 1 only "counted out":
//
//   mix_fp32<4>      the loop the shipped kernel runs: 4 Gaussians per thread, per 2 points 5 packed pair products, then per
//                    Gaussian 5 FMUL2 + 4 FMUL, 14 FADD (s + (lo + hi)), 13 FMNMX3; factor tables read from shared memory;
//                    128 threads, 4 CTAs per SM (128 registers).
//   mix_fp32<8>      8 Gaussians per thread (two z-quads): the pair products and the x / y table reads are amortised over
//                    twice the Gaussians; 160 accumulators -> 255 registers, 64-thread CTAs, 8 warps per SM.
//   mix_mma          4 Gaussians per thread, the 7 SUM channels moved to the tensor cores as 3xTF32 mma.sync.m16n8k8
//                    (D[16 (i,j) rows x 8 z-columns] += A[rows x 8 points] * B[8 points x cols], hi*hi + hi*lo + lo*hi):
//                    per 8 points and warp 7 tiles x 3 = 21 MMAs replace 224 FADDs, and cost the A fragments (20 products
//                    for the (row, point) pairs the fragment layout wants + 40 split instructions, their table reads) and
//                    the B fragments (6 reads + 12 split instructions).  Max / min stay on the FP32 pipes unchanged.
//
// The kernels compute nothing meaningful (tables are synthetic) -- they measure how fast each instruction mix issues.
// Output: T pairs/s per variant, comparable with the shipped kernel's 1.17-1.19 T pairs/s and with
// r01_microbench.jsonl::mix_separable_f32x2_fmnmx3 (register-only mix, 1.43).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench3 microbench3.cu && ./microbench3
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned tf32_hi(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kTilePairs = 64;      // 128 points per tile, as in the shipped kernel
constexpr int kRes = 8;
// tables: [pair][axis index] -> float4 (q_even, q_odd, qt_even, qt_odd) and float2 (qs_even, qs_odd); rows padded to 9
struct Tables { float4 a[3][kTilePairs][kRes + 1]; float2 b[3][kTilePairs][kRes + 1]; };

__device__ __forceinline__ void fill_tables(Tables& t, const float* seed) {
    for (int e = threadIdx.x; e < 3 * kTilePairs * (kRes + 1); e += blockDim.x) {
        const float s = seed[e & 31] + 1e-3f * (e % 97);
        (&t.a[0][0][0])[e] = make_float4(s, s * 0.99f, s - 1.f, 1.01f - s);
        (&t.b[0][0][0])[e] = make_float2(s * s - 1.f, 0.5f - s);
    }
    __syncthreads();
}

// One thread: NG Gaussians (fixed (i, j), NG consecutive z indices), `tiles` tiles of 128 points.
// SUMS_FP32 = false drops the 14 FADDs per Gaussian and point pair (the MMA variant adds its own work around it).
template <int NG, bool SUMS_FP32>
__device__ __forceinline__ void pair_step(const Tables& t, int p, int i, int j, int k0, float (&sum)[NG][7], float (&mx)[NG][7],
                                          float (&mn)[NG][6]) {
    const float4 xa = t.a[0][p][i], ya = t.a[1][p][j];
    const float2 xb = t.b[0][p][i], yb = t.b[1][p][j];
    const u64 qx = pack(xa.x, xa.y), qxt = pack(xa.z, xa.w), qxs = pack(xb.x, xb.y);
    const u64 qy = pack(ya.x, ya.y), qyt = pack(ya.z, ya.w), qys = pack(yb.x, yb.y);
    const u64 pxy = mul2(qx, qy), pxt = mul2(qxt, qy), pyt = mul2(qx, qyt), pxs = mul2(qxs, qy), pys = mul2(qx, qys);
    float pxy0, pxy1;
    unpack(pxy, pxy0, pxy1);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const float4 za = t.a[2][p][k0 + g];
        const float2 zb = t.b[2][p][k0 + g];
        const u64 qz = pack(za.x, za.y);
        float v[7][2];
        unpack(mul2(pxy, qz), v[0][0], v[0][1]);
        unpack(mul2(pxt, qz), v[1][0], v[1][1]);
        unpack(mul2(pyt, qz), v[2][0], v[2][1]);
        unpack(mul2(pxs, qz), v[4][0], v[4][1]);
        unpack(mul2(pys, qz), v[5][0], v[5][1]);
        v[3][0] = pxy0 * za.z; v[3][1] = pxy1 * za.w;
        v[6][0] = pxy0 * zb.x; v[6][1] = pxy1 * zb.y;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (SUMS_FP32) sum[g][c] = sum[g][c] + (v[c][0] + v[c][1]);
            mx[g][c] = fmax3(mx[g][c], v[c][0], v[c][1]);
            if (c) mn[g][c - 1] = fmin3(mn[g][c - 1], v[c][0], v[c][1]);
        }
    }
}

template <int NG, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) mix_fp32(float* out, const float* seed, int tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tables& t = *reinterpret_cast<Tables*>(smem_raw);
    fill_tables(t, seed);
    // thread -> (i, j, z-group): 64 (i, j) x (8 / NG) groups
    const int zg = threadIdx.x % (kRes / NG), ij = (threadIdx.x / (kRes / NG)) % 64, i = ij >> 3, j = ij & 7, k0 = zg * NG;
    float sum[NG][7], mx[NG][7], mn[NG][6];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int c = 0; c < 7; ++c) { sum[g][c] = 0.f; mx[g][c] = -1e30f; if (c) mn[g][c - 1] = 1e30f; }
    for (int tile = 0; tile < tiles; ++tile) {
#pragma unroll 1
        for (int p = 0; p < kTilePairs; ++p) pair_step<NG, true>(t, p, i, j, k0, sum, mx, mn);
        __syncthreads();            // the shipped kernel re-stages the tables here
    }
    float r = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int c = 0; c < 7; ++c) r += sum[g][c] + mx[g][c] + (c ? mn[g][c - 1] : 0.f);
    out[blockIdx.x * THREADS + threadIdx.x] = r;
}

// Sums on the tensor cores.  Warp = 16 (i, j) rows x 2 z-quads (lane = row * 2 + quad), as in the shipped kernel's thread map.
__global__ void __launch_bounds__(128, 4) mix_mma(float* out, const float* seed, int tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Tables& t = *reinterpret_cast<Tables*>(smem_raw);
    fill_tables(t, seed);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zg = lane & 1, ij = warp * 16 + (lane >> 1), i = ij >> 3, j = ij & 7, k0 = zg * 4;
    // fragment coordinates: rows fg, fg + 8 of the warp's 16 (i, j); points ft, ft + 4 of an 8-point group; column fg of B
    const int fg = lane >> 2, ft = lane & 3;
    const int r0 = warp * 16 + fg, r1 = r0 + 8;
    float sum[4][7], mx[4][7], mn[4][6];
    float acc[7][4];                 // 7 D tiles (pxy x {qz, qzt, qzs}; pxt, pyt, pxs, pys x qz), 4 registers each
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int c = 0; c < 7; ++c) { sum[g][c] = 0.f; mx[g][c] = -1e30f; if (c) mn[g][c - 1] = 1e30f; }
#pragma unroll
    for (int c = 0; c < 7; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    for (int tile = 0; tile < tiles; ++tile) {
#pragma unroll 1
        for (int p4 = 0; p4 < kTilePairs; p4 += 4) {          // 8 points
#pragma unroll
            for (int p = 0; p < 4; ++p) pair_step<4, false>(t, p4 + p, i, j, k0, sum, mx, mn);
            // ---- A fragments: 5 row types x (rows r0, r1) x (points ft, ft + 4), split hi / lo ----
            unsigned ahi[5][4], alo[5][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = (e & 1) ? r1 : r0, pt = ft + ((e >> 1) << 2);      // a0 (r0,t) a1 (r1,t) a2 (r0,t+4) a3 (r1,t+4)
                const int pr = p4 + (pt >> 1), odd = pt & 1;
                const float* xa = reinterpret_cast<const float*>(&t.a[0][pr][row >> 3]);
                const float* ya = reinterpret_cast<const float*>(&t.a[1][pr][row & 7]);
                const float* xb = reinterpret_cast<const float*>(&t.b[0][pr][row >> 3]);
                const float* yb = reinterpret_cast<const float*>(&t.b[1][pr][row & 7]);
                const float qx = xa[odd], qxt = xa[2 + odd], qxs = xb[odd], qy = ya[odd], qyt = ya[2 + odd], qys = yb[odd];
                const float v[5] = {qx * qy, qxt * qy, qx * qyt, qxs * qy, qx * qys};
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    ahi[c][e] = tf32_hi(v[c]);
                    alo[c][e] = __float_as_uint(v[c] - __uint_as_float(ahi[c][e]));
                }
            }
            // ---- B fragments: 3 z-factor types, b0 (point ft, column fg), b1 (point ft + 4, column fg) ----
            unsigned bhi[3][2], blo[3][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int pt = ft + (e << 2), pr = p4 + (pt >> 1), odd = pt & 1;
                const float* za = reinterpret_cast<const float*>(&t.a[2][pr][fg]);
                const float* zb = reinterpret_cast<const float*>(&t.b[2][pr][fg]);
                const float v[3] = {za[odd], za[2 + odd], zb[odd]};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    bhi[c][e] = tf32_hi(v[c]);
                    blo[c][e] = __float_as_uint(v[c] - __uint_as_float(bhi[c][e]));
                }
            }
            // ---- 7 tiles x (hi*hi + hi*lo + lo*hi) ----
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const int at = c < 3 ? 0 : c - 2, bt = c < 3 ? c : 0;
                mma_tf32(acc[c], alo[at], bhi[bt]);
                mma_tf32(acc[c], ahi[at], blo[bt]);
                mma_tf32(acc[c], ahi[at], bhi[bt]);
            }
        }
        __syncthreads();
    }
    float r = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int c = 0; c < 7; ++c) r += sum[g][c] + mx[g][c] + (c ? mn[g][c - 1] : 0.f);
#pragma unroll
    for (int c = 0; c < 7; ++c) r += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <typename K>
static void run(const char* name, K kern, int threads, int gaussians_per_cta, int grid, int tiles, float* out, const float* seed) {
    const size_t smem = sizeof(Tables);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    kern<<<grid, threads, smem>>>(out, seed, tiles);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<grid, threads, smem>>>(out, seed, tiles);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double pairs = (double)grid * gaussians_per_cta * tiles * 128.0;
    printf("{\"test\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"registers\": %d, \"local_bytes\": %d, \"grid\": %d, \"ms\": %.3f, "
           "\"Tpairs_per_s\": %.4f}\n", name, threads, occ, fa.numRegs, (int)fa.localSizeBytes, grid, best, pairs / (best * 1e9));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    float *out, *seed;
    CK(cudaMalloc(&out, sizeof(float) * sms * 64 * 128));
    float hs[32];
    for (int i = 0; i < 32; ++i) hs[i] = 0.999f + 1e-4f * i;
    CK(cudaMalloc(&seed, sizeof(hs)));
    CK(cudaMemcpy(seed, hs, sizeof(hs), cudaMemcpyHostToDevice));
    const int tiles = 256, waves = 8;
    run("k5_mix_fp32_4_gaussians_per_thread (shipped loop)", mix_fp32<4, 128, 4>, 128, 512, sms * 4 * waves, tiles, out, seed);
    run("k5_mix_fp32_8_gaussians_per_thread_64_threads", mix_fp32<8, 64, 4>, 64, 512, sms * 4 * waves, tiles, out, seed);
    run("k5_mix_fp32_8_gaussians_per_thread_128_threads_2_items", mix_fp32<8, 128, 2>, 128, 1024, sms * 2 * waves, tiles, out, seed);
    run("k5_mix_sums_on_mma_sync_3xtf32", mix_mma, 128, 512, sms * 4 * waves, tiles, out, seed);
    return 0;
}
