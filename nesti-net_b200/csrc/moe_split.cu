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
// so mups_conv3d_bn_relu (moe_conv.cu: TMA, tcgen05.mma, TMEM) runs unchanged on 3 x the K extent and writes its fp32
// output (scale / shift / ReLU applied) to a scratch tensor; the two kernels here turn fp32 back into triplets:
//
//   split3_kernel      fp32 [rows, src_stride] columns [src_off, +w_src) -> triplet at channels [dst_off, +3 w_dst) of
//                      an NDHWC bf16 tensor (w_src <= w_dst: the tail is zero padding, e.g. MuPS' 20 channels per scale
//                      in a 32-wide slot)
//   pool3_x3_kernel    tf_util.avg_pool3d ('SAME', stride 1, mean over the valid cells; utils/tf_util.py:432-455) and
//                      max_pool3d (2, stride 2; :406-430) from a triplet to a triplet: the window is reduced on
//                      hi + lo in fp32 and split again
// Both are memory-bound elementwise kernels (16-byte loads / stores, one thread per 8 channels).
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
    const int grid = (int)((n + 255) / 256 < 32 * kNumSMs ? (n + 255) / 256 : 32 * kNumSMs);
    pool3_x3_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16_dev), B, D, c_total, c_off, w, k,
                                                                           is_max, static_cast<__nv_bfloat16*>(y_bf16_dev), y_total, y_off);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

}  // extern "C"
