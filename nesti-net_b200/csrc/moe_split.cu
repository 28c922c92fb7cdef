// bf16x3 ("split") mode of the consumer: fp32-grade normals from the same tcgen05 kernels.
//
// A value v is carried as TWO bf16 numbers, hi = bf16(v) and lo = bf16(v - hi) (16 significant bits), and a product
// a * w is evaluated as a_hi w_hi + a_lo w_hi + a_hi w_lo with fp32 accumulation in tensor memory -- the dropped
// a_lo w_lo term is 2^-18 of the product.  No new convolution kernel is needed for that: the three terms are ONE
// convolution over a three times longer channel axis,
//
//     activations   [ hi | lo | hi ]            (a "triplet": 3 w channels for w logical ones)
//     weights       [ w_hi | w_hi | w_lo ]      (expanded once on the host, moe_engine.PackedConv)
//
// so the convolution kernels of moe_conv.cu (TMA, tcgen05.mma, TMEM) run with an unchanged main loop on 3 x the K extent; their
// epilogue writes the next layer's triplets itself (mups_conv3d_bn_relu_x3: the X3 instantiation of the kernels).  The kernels
// here are the rest of the mode:
//
//   split3_kernel      fp32 [rows, src_stride] columns [src_off, +w_src) -> triplet at channels [dst_off, +3 w_dst) of
//                      an NDHWC bf16 tensor (w_src <= w_dst: the tail is zero padding, e.g. MuPS' 20 channels per scale
//                      in a 32-wide slot)
//   pool3_x3_kernel    tf_util.avg_pool3d ('SAME', stride 1, mean over the valid cells; utils/tf_util.py:432-455) and
//                      max_pool3d (2, stride 2; :406-430) from a triplet to a triplet: the window is reduced on
//                      hi + lo in fp32 and split again
//   avgpool8_f32_x3_kernel / avgpool_f32_x3_kernel   the inception module's pool branch: its 1^3 convolution commutes with the
//                      average pool (both linear), so the convolution runs first (raw fp32 output, n_filters instead of C_in
//                      channels) and this kernel pools the fp32 tensor, applies the folded bias + batch norm + ReLU and writes
//                      the triplet -- 8^3 volumes as a separable box sum on a shared-memory tile (every byte read once)
// All are memory-bound kernels (16-byte loads / stores).
#include <mutex>

#include <cuda_bf16.h>

#include "mups_common.cuh"

namespace mups {

// v -> (hi, lo) for 8 values, packed as two 16-byte vectors
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
    __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        const float2 hf = __bfloat1622float2(h);
        float r0 = f[2 * j] - hf.x, r1 = f[2 * j + 1] - hf.y;        // exact in fp32 (Sterbenz / short mantissas)
        if (!(fabsf(r0) <= 3.0e38f)) r0 = 0.f;                         // inf - inf, NaN: the high part carries it alone
        if (!(fabsf(r1) <= 3.0e38f)) r1 = 0.f;
        h2[j] = h;
        l2[j] = __floats2bfloat162_rn(r0, r1);
    }
}

__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&f)[8]) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
    const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 h = __bfloat1622float2(h2[j]), l = __bfloat1622float2(l2[j]);
        f[2 * j] = h.x + l.x;
        f[2 * j + 1] = h.y + l.y;
    }
}

__device__ __forceinline__ void store_triplet(__nv_bfloat16* d, int w, const float (&f)[8]) {
    uint4 hi, lo;
    split8(f, hi, lo);
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + w) = lo;
    *reinterpret_cast<uint4*>(d + 2 * w) = hi;
}

__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long rows, int src_stride, int src_off, int w_src,
                                                     __nv_bfloat16* __restrict__ dst, int dst_stride, int dst_off, int w_dst) {
    const int chunks = w_dst >> 3;
    const long long n = rows * chunks;
    const bool vec = ((src_stride | src_off) & 3) == 0 && (w_src & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long r = i / chunks;
        const float* s = src + r * src_stride + src_off + ch * 8;
        float f[8];
        if (vec && ch * 8 < w_src) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = ch * 8 + j < w_src ? __ldg(s + j) : 0.f;
        }
        store_triplet(dst + r * dst_stride + dst_off + ch * 8, w_dst, f);
    }
}

__global__ void __launch_bounds__(256) pool3_x3_kernel(const __nv_bfloat16* __restrict__ x, long long B, int D, int ct, int x_off, int w, int k,
                                                       int is_max, __nv_bfloat16* __restrict__ y, int y_ct, int y_off) {
    const int chunks = w >> 3;
    const int Do = is_max ? D / 2 : D;
    const long long n = B * Do * Do * Do * chunks;
    const int pl = (k - 1) / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long v = i / chunks;
        const int xo = (int)(v % Do), yo = (int)((v / Do) % Do), zo = (int)((v / ((long long)Do * Do)) % Do);
        const long long b = v / ((long long)Do * Do * Do);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = is_max ? -INFINITY : 0.f;
        int cnt = 0;
        const int z0 = is_max ? 2 * zo : zo - pl, y0 = is_max ? 2 * yo : yo - pl, x0 = is_max ? 2 * xo : xo - pl;
        const int kk = is_max ? 2 : k;
        for (int dz = 0; dz < kk; ++dz) {
            const int z = z0 + dz;
            if (z < 0 || z >= D) continue;
            for (int dy = 0; dy < kk; ++dy) {
                const int yy = y0 + dy;
                if (yy < 0 || yy >= D) continue;
                for (int dx = 0; dx < kk; ++dx) {
                    const int xx = x0 + dx;
                    if (xx < 0 || xx >= D) continue;
                    const __nv_bfloat16* p = x + (((b * D + z) * D + yy) * D + xx) * (long long)ct + x_off + ch * 8;
                    float f[8];
                    join8(__ldg(reinterpret_cast<const uint4*>(p)), __ldg(reinterpret_cast<const uint4*>(p + w)), f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = is_max ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
                    ++cnt;
                }
            }
        }
        if (!is_max) {
            const float inv = 1.f / (float)cnt;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= inv;
        }
        store_triplet(y + v * (long long)y_ct + y_off + ch * 8, w, acc);
    }
}

// tf.nn.max_pool3d(2, stride 2) from a triplet to a triplet (D even): one thread per (output voxel, 8-channel chunk), the sixteen
// 16-byte loads of its window (hi and lo of eight cells) issued back to back, 32-bit index arithmetic.  The maximum of hi + lo
// is one of the window's values, so the output pair is exact.
__global__ void __launch_bounds__(256) maxpool2_x3_kernel(const __nv_bfloat16* __restrict__ x, unsigned n, int D, int ct, int x_off, int w,
                                                          __nv_bfloat16* __restrict__ y, int y_ct, int y_off) {
    const unsigned chunks = (unsigned)w >> 3, Do = (unsigned)D >> 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned ch = i % chunks, v = i / chunks;
        const unsigned xo = v % Do, yo = (v / Do) % Do, zo = (v / (Do * Do)) % Do, b = v / (Do * Do * Do);
        const size_t vin = (((size_t)b * D + 2 * zo) * D + 2 * yo) * D + 2 * xo;
        const uint4* p = reinterpret_cast<const uint4*>(x + vin * ct + x_off) + ch;
        const size_t sx = (size_t)ct / 8, sy = sx * D, sz = sy * D, sl = (size_t)w / 8;      // strides in 16-byte units
        uint4 hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const uint4* q = p + (t & 1) * sx + ((t >> 1) & 1) * sy + (t >> 2) * sz;
            hi[t] = __ldg(q);
            lo[t] = __ldg(q + sl);
        }
        float m[8];
        join8(hi[0], lo[0], m);
#pragma unroll
        for (int t = 1; t < 8; ++t) {
            float f[8];
            join8(hi[t], lo[t], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
        }
        store_triplet(y + (size_t)v * y_ct + y_off + ch * 8, w, m);
    }
}

// fp32 [B, 8, 8, 8, c] -> TF 'SAME' average pool (window K, mean over the valid cells) -> act(scale * . + shift) -> triplet.
// One CTA = one sample x 32 channels: the 512 x 32 tile is staged once in shared memory (64 KB) and the box sum is done
// separably in place (three passes of 64 lines of 8 voxels, one warp per line, lane = channel).
template <int K>
__global__ void __launch_bounds__(256) avgpool8_f32_x3_kernel(const float* __restrict__ x, int c, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int relu, __nv_bfloat16* __restrict__ y,
                                                              int y_total, int y_off) {
    extern __shared__ __align__(16) float tile[];                  // [512 voxels][32 channels]
    constexpr int D = 8, PL = (K - 1) / 2;
    const int chunks = c >> 5;
    const long long b = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = x + b * 512 * (long long)c + chunk * 32;
#pragma unroll
    for (int half = 0; half < 2; ++half) {                         // 4096 float4 per tile: two batches of eight loads in flight
        float4 raw[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + (half * 8 + r) * 256;
            raw[r] = __ldg(reinterpret_cast<const float4*>(src + (i >> 3) * (long long)c) + (i & 7));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + (half * 8 + r) * 256;
            *reinterpret_cast<float4*>(tile + (i >> 3) * 32 + (i & 7) * 4) = raw[r];
        }
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const int step = pass == 0 ? 1 : pass == 1 ? 8 : 64;
        for (int l = warp; l < 64; l += 8) {
            const int base = pass == 0 ? l * 8 : pass == 1 ? (l >> 3) * 64 + (l & 7) : l;
            float r[D], o[D];
#pragma unroll
            for (int i = 0; i < D; ++i) r[i] = tile[(base + i * step) * 32 + lane];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    const int j = i + d - PL;
                    if (j >= 0 && j < D) acc += r[j];
                }
                o[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < D; ++i) tile[(base + i * step) * 32 + lane] = o[i];
        }
        __syncthreads();
    }
    __nv_bfloat16* dst = y + b * 512 * (long long)y_total + y_off + chunk * 32;
    for (int i = tid; i < 512 * 4; i += 256) {
        const int v = i >> 2, part = i & 3;
        const int vx = v & 7, vy = (v >> 3) & 7, vz = v >> 6;
        auto valid = [](int q) { return min(q - PL + K - 1, D - 1) - max(q - PL, 0) + 1; };
        const float inv = 1.f / (float)(valid(vx) * valid(vy) * valid(vz));
        const float4* sp = reinterpret_cast<const float4*>(tile + v * 32 + part * 8);
        const float4 a0 = sp[0], a1 = sp[1];
        float f[8] = {a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv, a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv};
        const int c0 = chunk * 32 + part * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(f[j], __ldg(scale + c0 + j), __ldg(shift + c0 + j));
            f[j] = relu ? fmaxf(t, 0.f) : t;
        }
        store_triplet(dst + v * (long long)y_total + part * 8, c, f);
    }
}

// the same for the 4^3 / 2^3 volumes (and channel counts that are not multiples of 32): one thread per (voxel, 8 channels)
__global__ void __launch_bounds__(256) avgpool_f32_x3_kernel(const float* __restrict__ x, long long B, int D, int c, int k,
                                                             const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                             __nv_bfloat16* __restrict__ y, int y_total, int y_off) {
    const int chunks = c >> 3;
    const long long n = B * D * D * D * chunks;
    const int pl = (k - 1) / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long v = i / chunks;
        const int xo = (int)(v % D), yo = (int)((v / D) % D), zo = (int)((v / ((long long)D * D)) % D);
        const long long b = v / ((long long)D * D * D);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int cnt = 0;
        for (int dz = 0; dz < k; ++dz) {
            const int z = zo - pl + dz;
            if (z < 0 || z >= D) continue;
            for (int dy = 0; dy < k; ++dy) {
                const int yy = yo - pl + dy;
                if (yy < 0 || yy >= D) continue;
                for (int dx = 0; dx < k; ++dx) {
                    const int xx = xo - pl + dx;
                    if (xx < 0 || xx >= D) continue;
                    const float4* p = reinterpret_cast<const float4*>(x + (((b * D + z) * D + yy) * D + xx) * (long long)c + ch * 8);
                    const float4 a0 = __ldg(p), a1 = __ldg(p + 1);
                    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                    acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
                    ++cnt;
                }
            }
        }
        const float inv = 1.f / (float)cnt;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(acc[j] * inv, __ldg(scale + ch * 8 + j), __ldg(shift + ch * 8 + j));
            acc[j] = relu ? fmaxf(t, 0.f) : t;
        }
        store_triplet(y + v * (long long)y_total + y_off + ch * 8, c, acc);
    }
}

template <int K>
static cudaError_t launch_avgpool8_f32_x3(const float* x, long long B, int c, const float* scale, const float* shift, int relu,
                                          __nv_bfloat16* y, int y_total, int y_off, cudaStream_t st) {
    constexpr int smem = 512 * 32 * 4;
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(avgpool8_f32_x3_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    avgpool8_f32_x3_kernel<K><<<(unsigned)(B * (c >> 5)), 256, smem, st>>>(x, c, scale, shift, relu, y, y_total, y_off);
    return cudaGetLastError();
}

}  // namespace mups

using namespace mups;

extern "C" {

int mups_split_bf16x3(const float* src_dev, int64_t rows, int src_stride, int src_off, int w_src, void* dst_bf16_dev, int dst_stride,
                      int dst_off, int w_dst, mups_stream stream) {
    MUPS_REQUIRE(rows >= 0 && rows < (1ll << 40), "mups_split_bf16x3: rows=%lld out of range", (long long)rows);
    MUPS_REQUIRE(w_src >= 1 && w_dst >= w_src && w_dst % 8 == 0, "mups_split_bf16x3: widths %d -> %d (destination: a multiple of 8)", w_src, w_dst);
    MUPS_REQUIRE(src_off >= 0 && src_off + w_src <= src_stride, "mups_split_bf16x3: source columns (%d of %d at %d)", w_src, src_stride, src_off);
    MUPS_REQUIRE(dst_stride % 8 == 0 && dst_off >= 0 && dst_off % 8 == 0 && dst_off + 3 * w_dst <= dst_stride,
                 "mups_split_bf16x3: destination triplet (3 x %d of %d at %d)", w_dst, dst_stride, dst_off);
    MUPS_REQUIRE(rows == 0 || (src_dev && dst_bf16_dev), "mups_split_bf16x3: NULL buffer");
    MUPS_REQUIRE((reinterpret_cast<uintptr_t>(dst_bf16_dev) & 15) == 0, "mups_split_bf16x3: destination must be 16-byte aligned");
    if (rows == 0) return MUPS_OK;
    const long long n = (long long)rows * (w_dst / 8);
    const int grid = (int)((n + 255) / 256 < 32 * kNumSMs ? (n + 255) / 256 : 32 * kNumSMs);
    split3_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src_dev, rows, src_stride, src_off, w_src,
                                                                         static_cast<__nv_bfloat16*>(dst_bf16_dev), dst_stride, dst_off, w_dst);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

int mups_pool3d_bf16x3(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int w, int k, int is_max, void* y_bf16_dev,
                       int y_total, int y_off, mups_stream stream) {
    MUPS_REQUIRE(x_bf16_dev && y_bf16_dev, "mups_pool3d_bf16x3: NULL buffer");
    MUPS_REQUIRE(B >= 1 && B < (1ll << 30) && (D == 2 || D == 4 || D == 8), "mups_pool3d_bf16x3: B=%lld, volume edge %d", (long long)B, D);
    MUPS_REQUIRE(w >= 8 && w % 8 == 0 && c_total % 8 == 0 && c_off >= 0 && c_off % 8 == 0 && c_off + 3 * w <= c_total,
                 "mups_pool3d_bf16x3: input triplet (3 x %d of %d at %d) must be multiples of 8", w, c_total, c_off);
    MUPS_REQUIRE(y_total % 8 == 0 && y_off >= 0 && y_off % 8 == 0 && y_off + 3 * w <= y_total,
                 "mups_pool3d_bf16x3: output triplet (3 x %d of %d at %d)", w, y_total, y_off);
    MUPS_REQUIRE(is_max ? k == 2 : (k >= 1 && k <= 5), "mups_pool3d_bf16x3: window %d", k);
    const int Do = is_max ? D / 2 : D;
    const long long n = (long long)B * Do * Do * Do * (w / 8);
    if (is_max && (long long)B * D * D * D * (c_total / 8) < 0xFFFFFFFFll) {
        maxpool2_x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(x_bf16_dev), (unsigned)n, D, c_total, c_off, w, static_cast<__nv_bfloat16*>(y_bf16_dev), y_total, y_off);
        MUPS_CHECK_LAUNCH();
        return MUPS_OK;
    }
    const int grid = (int)((n + 255) / 256 < 32 * kNumSMs ? (n + 255) / 256 : 32 * kNumSMs);
    pool3_x3_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16_dev), B, D, c_total, c_off, w, k,
                                                                           is_max, static_cast<__nv_bfloat16*>(y_bf16_dev), y_total, y_off);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

int mups_avgpool3d_f32_bn_relu_x3(const float* x_f32_dev, int64_t B, int D, int c, int k, const float* scale_dev, const float* shift_dev,
                                  int relu, void* y_bf16_dev, int y_total, int y_off, mups_stream stream) {
    MUPS_REQUIRE(x_f32_dev && y_bf16_dev && scale_dev && shift_dev, "mups_avgpool3d_f32_bn_relu_x3: NULL buffer");
    MUPS_REQUIRE(B >= 1 && B < (1ll << 30) && (D == 2 || D == 4 || D == 8), "mups_avgpool3d_f32_bn_relu_x3: B=%lld, volume edge %d", (long long)B, D);
    MUPS_REQUIRE(c >= 8 && c % 8 == 0, "mups_avgpool3d_f32_bn_relu_x3: channels %d must be a multiple of 8", c);
    MUPS_REQUIRE(k >= 2 && k <= 5, "mups_avgpool3d_f32_bn_relu_x3: window %d", k);
    MUPS_REQUIRE(y_total % 8 == 0 && y_off >= 0 && y_off % 8 == 0 && y_off + 3 * c <= y_total,
                 "mups_avgpool3d_f32_bn_relu_x3: output triplet (3 x %d of %d at %d)", c, y_total, y_off);
    MUPS_REQUIRE((reinterpret_cast<uintptr_t>(x_f32_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_bf16_dev) & 15) == 0,
                 "mups_avgpool3d_f32_bn_relu_x3: buffers must be 16-byte aligned");
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* y = static_cast<__nv_bfloat16*>(y_bf16_dev);
    if (D == 8 && c % 32 == 0 && (long long)B * (c >> 5) <= 0x7FFFFFFFll) {
        const cudaError_t e = k == 2 ? launch_avgpool8_f32_x3<2>(x_f32_dev, B, c, scale_dev, shift_dev, relu, y, y_total, y_off, st)
                            : k == 3 ? launch_avgpool8_f32_x3<3>(x_f32_dev, B, c, scale_dev, shift_dev, relu, y, y_total, y_off, st)
                            : k == 4 ? launch_avgpool8_f32_x3<4>(x_f32_dev, B, c, scale_dev, shift_dev, relu, y, y_total, y_off, st)
                                     : launch_avgpool8_f32_x3<5>(x_f32_dev, B, c, scale_dev, shift_dev, relu, y, y_total, y_off, st);
        if (e != cudaSuccess) { set_error("mups_avgpool3d_f32_bn_relu_x3: %s", cudaGetErrorString(e)); return MUPS_ERR_CUDA; }
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return MUPS_OK;
    }
    const long long n = (long long)B * D * D * D * (c / 8);
    const int grid = (int)((n + 255) / 256 < 32 * kNumSMs ? (n + 255) / 256 : 32 * kNumSMs);
    avgpool_f32_x3_kernel<<<grid, 256, 0, st>>>(x_f32_dev, B, D, c, k, scale_dev, shift_dev, relu, y, y_total, y_off);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

}  // extern "C"
