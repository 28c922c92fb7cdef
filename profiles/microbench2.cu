// Pipe-interaction micro-benchmarks (B200): do FMA-pipe and ALU-pipe instructions of the MuPS
// statistics loop overlap, and what does a 3-source FMNMX3 / a packed FMUL2 cost when all source
// registers are distinct (no operand-reuse cache hits)?  Decides the issue-rate ceiling of K5.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu && ./microbench2
// Output: JSON lines {"test", "warp_inst_per_clk_per_smsp", ...}; clocks from clock64 of resident warps.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 128;     // 4 warps per CTA: one per SMSP
constexpr int kIters = 65536;

typedef unsigned long long u64;
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmaxs(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

enum Test { T_FADD, T_FMNMX, T_FMNMX3, T_FFMA3, T_FMUL2, T_FADD_FMNMX, T_FADD_FMNMX3, T_FMUL2_FMNMX3, T_FADD2_FMNMX3, T_LOOPMIX, T_LOOPMIX_PACKED, T_LOOPMIX_SCALAR, T_LOOPMIX_SCALAR_FFMA, T_COUNT };

// a[]: 16 accumulators; b[], c[]: 16 + 16 read-only distinct source registers; p*: packed versions
template <int T>
__global__ void __launch_bounds__(kThreads) k(float* out, const float* seed, long long* cycles) {
    float a[16], b[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed[(threadIdx.x + i) & 63]; b[i] = seed[(threadIdx.x + 2 * i + 1) & 63]; c[i] = seed[(threadIdx.x + 3 * i + 2) & 63]; }
    u64 pa[8], pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pa[i] = pack(a[2 * i], a[2 * i + 1]); pb[i] = pack(b[2 * i], b[2 * i + 1]); }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (T == T_FADD) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fadd(a[i], b[i]);
        } else if (T == T_FMNMX) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaxs(a[i], b[i]);
        } else if (T == T_FMNMX3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmax3(a[i], b[i], c[i]);
        } else if (T == T_FFMA3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ffma(b[i], c[i], a[i]);
        } else if (T == T_FMUL2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pa[i] = mul2(pa[i], pb[i]);
        } else if (T == T_FADD_FMNMX) {         // 8 + 8
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fadd(a[i], b[i]); a[8 + i] = fmaxs(a[8 + i], c[i]); }
        } else if (T == T_FADD_FMNMX3) {        // 8 + 8
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fadd(a[i], b[i]); a[8 + i] = fmax3(a[8 + i], b[8 + i], c[i]); }
        } else if (T == T_FMUL2_FMNMX3) {       // 8 + 8
#pragma unroll
            for (int i = 0; i < 8; ++i) { pa[i] = mul2(pa[i], pb[i]); a[i] = fmax3(a[i], b[8 + (i & 7)], c[i]); }
        } else if (T == T_FADD2_FMNMX3) {       // 8 + 8
#pragma unroll
            for (int i = 0; i < 8; ++i) { pa[i] = add2(pa[i], pb[i]); a[i] = fmax3(a[i], b[8 + (i & 7)], c[i]); }
        } else if (T == T_LOOPMIX) {
            // one Gaussian x two points of the K5 main loop: 7 FMUL2, 14 FADD, 13 FMNMX3 (+ the amortised 5/4 FMUL2)
            u64 t[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) t[i] = mul2(pb[i], pb[(i + 1) & 7]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                float lo, hi; unpack(t[i], lo, hi);
                a[i] = fadd(fadd(a[i], lo), hi);
                a[7 + i] = fmax3(a[7 + i], lo, hi);
                if (i > 0) c[i] = fmin3(c[i], lo, hi);
            }
            pb[7] = mul2(pb[7], pa[0]);
        } else if (T == T_LOOPMIX_PACKED) {
            u64 t[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) t[i] = mul2(pb[i], pb[(i + 1) & 7]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                float lo, hi; unpack(t[i], lo, hi);
                pa[i] = add2(pa[i], t[i]);
                a[i] = fmax3(a[i], lo, hi);
                if (i > 0) c[i] = fmin3(c[i], lo, hi);
            }
            pb[7] = mul2(pb[7], pa[7]);
        } else if (T == T_LOOPMIX_SCALAR) {
            // all scalar, products ordered so that consecutive FMULs share an operand (reuse cache):
            // u_qq * {qz, az, bz}, then {u_aq, u_qa, u_bq, u_qb} * qz -- for both points
            // b[0..4] = u_* of point 0, b[5..9] = u_* of point 1; c[0..2] = qz, az, bz of point 0, c[3..5] of point 1
            float v0[7], v1[7];
            v0[0] = fmul(b[0], c[0]); v0[3] = fmul(b[0], c[1]); v0[6] = fmul(b[0], c[2]);
            v0[1] = fmul(b[1], c[0]); v0[2] = fmul(b[2], c[0]); v0[4] = fmul(b[3], c[0]); v0[5] = fmul(b[4], c[0]);
            v1[0] = fmul(b[5], c[3]); v1[3] = fmul(b[5], c[4]); v1[6] = fmul(b[5], c[5]);
            v1[1] = fmul(b[6], c[3]); v1[2] = fmul(b[7], c[3]); v1[4] = fmul(b[8], c[3]); v1[5] = fmul(b[9], c[3]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                a[i] = fadd(fadd(a[i], v0[i]), v1[i]);
                a[7 + i] = fmax3(a[7 + i], v0[i], v1[i]);
                if (i > 0) c[8 + i] = fmin3(c[8 + i], v0[i], v1[i]);
            }
            b[10] = fmul(b[10], a[0]);
        } else if (T == T_LOOPMIX_SCALAR_FFMA) {
            // as above, sums fused into FFMA (acc += u * z) next to the FMUL that makes the same product
            float v0[7], v1[7];
            v0[0] = fmul(b[0], c[0]); a[0] = ffma(b[0], c[0], a[0]); v0[3] = fmul(b[0], c[1]); a[3] = ffma(b[0], c[1], a[3]);
            v0[6] = fmul(b[0], c[2]); a[6] = ffma(b[0], c[2], a[6]);
            v0[1] = fmul(b[1], c[0]); a[1] = ffma(b[1], c[0], a[1]); v0[2] = fmul(b[2], c[0]); a[2] = ffma(b[2], c[0], a[2]);
            v0[4] = fmul(b[3], c[0]); a[4] = ffma(b[3], c[0], a[4]); v0[5] = fmul(b[4], c[0]); a[5] = ffma(b[4], c[0], a[5]);
            v1[0] = fmul(b[5], c[3]); a[0] = ffma(b[5], c[3], a[0]); v1[3] = fmul(b[5], c[4]); a[3] = ffma(b[5], c[4], a[3]);
            v1[6] = fmul(b[5], c[5]); a[6] = ffma(b[5], c[5], a[6]);
            v1[1] = fmul(b[6], c[3]); a[1] = ffma(b[6], c[3], a[1]); v1[2] = fmul(b[7], c[3]); a[2] = ffma(b[7], c[3], a[2]);
            v1[4] = fmul(b[8], c[3]); a[4] = ffma(b[8], c[3], a[4]); v1[5] = fmul(b[9], c[3]); a[5] = ffma(b[9], c[3], a[5]);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                a[7 + i] = fmax3(a[7 + i], v0[i], v1[i]);
                if (i > 0) c[8 + i] = fmin3(c[8 + i], v0[i], v1[i]);
            }
            b[10] = fmul(b[10], a[0]);
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i] + b[i] + c[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; unpack(pa[i], x, y); r += x + y; unpack(pb[i], x, y); r += x + y; }
    out[blockIdx.x * kThreads + threadIdx.x] = r;
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int T>
static void run(const char* name, double inst_per_iter, int sms, int ctas_per_sm, float* out, const float* seed, long long* cyc) {
    const int grid = sms * ctas_per_sm;
    k<T><<<grid, kThreads>>>(out, seed, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 4; ++r) k<T><<<grid, kThreads>>>(out, seed, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 4;
    // warp instructions per SM-clock per SMSP, at the nominal boost clock (1965 MHz, what NVML reports under this load)
    const double warp_inst = inst_per_iter * kIters * (double)grid * (kThreads / 32);
    const double rate = warp_inst / (ms * 1e-3 * 1.965e9 * sms * 4);
    printf("{\"test\": \"%s\", \"warps_per_smsp\": %d, \"warp_inst_per_clk_per_smsp\": %.3f, \"ms\": %.3f}\n", name, ctas_per_sm, rate, ms);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float *out, *seed; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * kThreads));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 8 * 4));
    float hs[64];
    for (int i = 0; i < 64; ++i) hs[i] = 0.999f + 1e-4f * i;
    CK(cudaMalloc(&seed, sizeof(hs)));
    CK(cudaMemcpy(seed, hs, sizeof(hs), cudaMemcpyHostToDevice));
    for (int w : {2, 4, 8}) {
        run<T_FADD>("fadd", 16, sms, w, out, seed, cyc);
        run<T_FMNMX>("fmnmx", 16, sms, w, out, seed, cyc);
        run<T_FMNMX3>("fmnmx3", 16, sms, w, out, seed, cyc);
        run<T_FFMA3>("ffma_3src", 16, sms, w, out, seed, cyc);
        run<T_FMUL2>("fmul2", 8, sms, w, out, seed, cyc);
        run<T_FADD_FMNMX>("fadd+fmnmx", 16, sms, w, out, seed, cyc);
        run<T_FADD_FMNMX3>("fadd+fmnmx3", 16, sms, w, out, seed, cyc);
        run<T_FMUL2_FMNMX3>("fmul2+fmnmx3", 16, sms, w, out, seed, cyc);
        run<T_FADD2_FMNMX3>("fadd2+fmnmx3", 16, sms, w, out, seed, cyc);
        run<T_LOOPMIX>("k5_loop_scalar_sums(35 inst = 2 pair-warps)", 35, sms, w, out, seed, cyc);
        run<T_LOOPMIX_PACKED>("k5_loop_packed_sums(28 inst = 2 pair-warps)", 28, sms, w, out, seed, cyc);
        run<T_LOOPMIX_SCALAR>("k5_loop_all_scalar(42 inst = 2 pair-warps)", 42, sms, w, out, seed, cyc);
        run<T_LOOPMIX_SCALAR_FFMA>("k5_loop_scalar_ffma_sums(42 inst = 2 pair-warps)", 42, sms, w, out, seed, cyc);
    }
    return 0;
}
