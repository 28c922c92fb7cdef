// Instruction-throughput micro-benchmarks for the pipes the MuPS statistics kernel lives on
// (FP32 FMA pipe, ALU pipe min/max, MUFU ex2, packed f32x2, 3-input min/max) on the B200 this
// runs on.  The measured per-SM rates are the denominators of the FP32/MUFU roofline in bench.py
// (SURVEY.md section 8d asks for measured micro-peaks next to the HBM peak).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
//
// Output: one JSON object per line: {"test": ..., "lane_ops_per_clk_per_sm": ..., "gops": ...}
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kIters = 4096;

__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// every kernel: out[tid] = f(state) so nothing is dead; `seed` defeats constant folding
enum Test { T_FFMA, T_FMUL_FADD, T_FMNMX, T_FMNMX3, T_FFMA2, T_MUL2_ADD2, T_MUFU, T_MIX_PLAIN, T_MIX_FAST, T_MIX_GENERAL, T_COUNT };

template <int T>
__global__ void __launch_bounds__(kThreads) bench_kernel(float* out, const float* seed, long long* cycles) {
    const float s0 = seed[threadIdx.x & 31], s1 = seed[(threadIdx.x + 7) & 31], s2 = seed[(threadIdx.x + 13) & 31];
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = s0 + i;
    u64 pa[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pa[i] = pack(s0 + i, s1 + i);
    const u64 pb = pack(s1, s2), pc = pack(s2, s0);
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        if (T == T_FFMA) {            // 16 independent FFMA, 3 distinct source registers
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s1, s2);
        } else if (T == T_FMUL_FADD) { // 8 FMUL + 8 FADD
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = a[i] * s1; a[8 + i] = a[8 + i] + s2; }
        } else if (T == T_FMNMX) {    // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = (i & 1) ? fmaxf(a[i], a[(i + 5) & 15]) : fminf(a[i], a[(i + 3) & 15]);
        } else if (T == T_FMNMX3) {   // 16 FMNMX3
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = (i & 1) ? fmax3(a[i], a[(i + 5) & 15], s1) : fmin3(a[i], a[(i + 3) & 15], s2);
        } else if (T == T_FFMA2) {    // 8 FFMA2 = 16 lane-FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) pa[i] = fma2(pa[i], pb, pc);
        } else if (T == T_MUL2_ADD2) { // 4 FMUL2 + 4 FADD2 = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 4; ++i) { pa[i] = mul2(pa[i], pb); pa[4 + i] = add2(pa[4 + i], pc); }
        } else if (T == T_MUFU) {     // 16 MUFU.EX2
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);
        } else if (T == T_MIX_PLAIN) {
            // separable-path inner loop, plain: per (point, Gaussian) 7 FMUL + 7 FADD + 7 max + 6 min
            // two Gaussians per iteration; a[0..6] products are recomputed from a rotating q
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float q = a[14 + g];
                const float v0 = q * s0, v1 = q * s1, v2 = q * s2, v3 = v0 * s1, v4 = v1 * s2, v5 = v2 * s0, v6 = v3 * s2;
                a[0] += v0; a[1] += v1; a[2] += v2; a[3] += v3; a[4] += v4; a[5] += v5; a[6] += v6;
                a[7] = fmaxf(a[7], v0); a[8] = fmaxf(a[8], v1); a[9] = fmaxf(a[9], v2); a[10] = fmaxf(a[10], v3);
                a[11] = fmaxf(a[11], v4); a[12] = fmaxf(a[12], v5); a[13] = fmaxf(a[13], v6);
                a[7] = fminf(a[7], v1); a[8] = fminf(a[8], v2); a[9] = fminf(a[9], v3); a[10] = fminf(a[10], v4);
                a[11] = fminf(a[11], v5); a[12] = fminf(a[12], v6);
                a[14 + g] = q + 1e-9f;
            }
        } else if (T == T_MIX_FAST) {
            // same work for 2 points x 2 Gaussians: packed f32x2 products/sums (point pair in one
            // register pair) + 3-input min/max folding both points at once
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const u64 q = pa[6 + g];
                const u64 v0 = mul2(q, pb), v1 = mul2(q, pc), v2 = mul2(q, pa[5]), v3 = mul2(v0, pc), v4 = mul2(v1, pb),
                          v5 = mul2(v2, pb), v6 = mul2(v3, pc);
                pa[0] = add2(pa[0], v0); pa[1] = add2(pa[1], v1); pa[2] = add2(pa[2], v2); pa[3] = add2(pa[3], v3);
                pa[4] = add2(pa[4], v4); a[0] += __uint_as_float((unsigned)v5); a[1] += __uint_as_float((unsigned)(v5 >> 32));
                a[2] += __uint_as_float((unsigned)v6); a[3] += __uint_as_float((unsigned)(v6 >> 32));
                float x, y;
                unpack(v0, x, y); a[4] = fmax3(a[4], x, y); a[5] = fmin3(a[5], x, y);
                unpack(v1, x, y); a[6] = fmax3(a[6], x, y); a[7] = fmin3(a[7], x, y);
                unpack(v2, x, y); a[8] = fmax3(a[8], x, y); a[9] = fmin3(a[9], x, y);
                unpack(v3, x, y); a[10] = fmax3(a[10], x, y); a[11] = fmin3(a[11], x, y);
                unpack(v4, x, y); a[12] = fmax3(a[12], x, y); a[13] = fmin3(a[13], x, y);
                unpack(v5, x, y); a[14] = fmax3(a[14], x, y); a[15] = fmin3(a[15], x, y);
                unpack(v6, x, y); a[4] = fmax3(a[4], x, y);
                pa[6 + g] = add2(q, pc);
            }
        } else if (T == T_MIX_GENERAL) {
            // general-path inner loop for one (point, Gaussian): 3 sub, 3 mul, 3 fma-ish, ex2, 20 reductions
            const float tx = (a[15] - s0) * s1, ty = (a[15] - s1) * s2, tz = (a[15] - s2) * s0;
            const float ss = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
            const float Q = ex2(fmaf(ss, -0.72134752f, s1)) * s2;
            const float d = fmaf(Q, s0, s1);
            a[0] = fmaxf(a[0], d); a[1] += d;
            const float ax = Q * tx, ay = Q * ty, az = Q * tz;
            a[2] = fmaxf(a[2], ax); a[3] = fmaxf(a[3], ay); a[4] = fmaxf(a[4], az);
            a[5] = fminf(a[5], ax); a[6] = fminf(a[6], ay); a[7] = fminf(a[7], az);
            a[8] += ax; a[9] += ay; a[10] += az;
            const float bx = fmaf(ax, tx, -Q), by = fmaf(ay, ty, -Q), bz = fmaf(az, tz, -Q);
            a[11] = fmaxf(a[11], bx); a[12] = fmaxf(a[12], by); a[13] = fmaxf(a[13], bz);
            a[11] = fminf(a[11], by); a[12] = fminf(a[12], bz); a[13] = fminf(a[13], bx);
            a[14] += bx + by + bz;
            a[15] += 1e-9f;
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; unpack(pa[i], x, y); r += x + y; }
    out[blockIdx.x * kThreads + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Spec { const char* name; double lane_ops_per_iter; const char* unit; };

template <int T>
static void run(const Spec& sp, int sms, int blocks_per_sm, float* out, const float* seed, long long* cyc) {
    const int grid = sms * blocks_per_sm;
    bench_kernel<T><<<grid, kThreads>>>(out, seed, cyc);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    bench_kernel<T><<<grid, kThreads>>>(out, seed, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; long long mx = 0;
    for (long long c : h) { mean += (double)c; if (c > mx) mx = c; }
    mean /= grid;
    // per SM: blocks_per_sm * kThreads threads each doing lane_ops_per_iter * kIters lane-ops in `mean` cycles
    const double per_clk_per_sm = sp.lane_ops_per_iter * kIters * kThreads * blocks_per_sm / mean;
    const double total = sp.lane_ops_per_iter * kIters * (double)kThreads * grid;
    printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"%s_per_clk_per_sm\": %.2f, \"g%s_per_s\": %.1f, \"ms\": %.4f, "
           "\"mean_cycles\": %.0f, \"max_cycles\": %lld, \"implied_mhz\": %.0f}\n",
           sp.name, blocks_per_sm, sp.unit, per_clk_per_sm, sp.unit, total / (ms * 1e6), ms, mean, mx, mx / (ms * 1e3));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    float *out, *seed; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * kThreads));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 8));
    float hs[32];
    for (int i = 0; i < 32; ++i) hs[i] = 0.999f + 1e-4f * i;
    CK(cudaMalloc(&seed, sizeof(hs)));
    CK(cudaMemcpy(seed, hs, sizeof(hs), cudaMemcpyHostToDevice));
    for (int bps : {2, 4, 8}) {
        run<T_FFMA>({"ffma", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_FMUL_FADD>({"fmul_fadd", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_FMNMX>({"fmnmx", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_FMNMX3>({"fmnmx3", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_FFMA2>({"ffma2", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_MUL2_ADD2>({"fmul2_fadd2", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_MUFU>({"mufu_ex2", 16, "lane_op"}, sms, bps, out, seed, cyc);
        run<T_MIX_PLAIN>({"mix_separable_plain", 2, "pair"}, sms, bps, out, seed, cyc);
        run<T_MIX_FAST>({"mix_separable_f32x2_fmnmx3", 4, "pair"}, sms, bps, out, seed, cyc);
        run<T_MIX_GENERAL>({"mix_general", 1, "pair"}, sms, bps, out, seed, cyc);
    }
    return 0;
}
