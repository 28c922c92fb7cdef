// Consumer of the MuPS tensor: the 3D-Inception / Mixture-of-Experts convolutions of
// models/experts_n_est.py:155-314 (tf_util.conv3d :254-311 and fully_connected :314-351, both with batch norm and ReLU)
// as hand-written implicit GEMMs on the 5th-generation tensor cores:
//
//   y[b, z, y, x, co] = act( scale[co] * sum_{dz,dy,dx,ci} x[b, z+dz-p, y+dy-p, x+dx-p, ci] * w[dz,dy,dx][co][ci] + shift[co] )
//
// * GEMM view: voxels x output channels x (taps x input channels, 64 per stage).
// * Activations: TMA loads of the NDHWC bf16 tensor through a 5-D tiled tensor map whose box is (64 channels, W, H, z-slices,
//   samples); a tap only shifts the box coordinates and the TMA unit zero-fills what falls outside the volume -- TF 'SAME'
//   padding (asymmetric for even kernels: the smaller half first) without an im2col buffer and without a single predicate.
//   The box lands in shared memory as rows of 128 bytes in the 128-byte-swizzled K-major layout tcgen05.mma reads.
// * Weights [tap][Cout][Cin] bf16 through a 3-D tensor map, same layout.
// * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) issued by one elected thread, accumulators in tensor memory, tcgen05.commit
//   arrives on the mbarriers that free the shared-memory stages / announce the finished accumulator.
// * warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (tcgen05.ld, folded bias + batch norm as scale /
//   shift, ReLU, bf16 pack, stores into a channel slice of the NDHWC output, so the inception module's concat costs nothing).
// Kernels:
//   conv3d_tcgen05_kernel   one 128-voxel activation tile (two when N <= 128) + one weight tile per (tap, channel block):
//                           the 1^3 layers (two CTAs per SM), the 4^3 / 2^3 volumes, the fully connected layers
//   conv3d_zhalo_kernel     the 8^3 volumes with k > 1 (85 % of the arithmetic): one activation box per (dy, dx, channel
//                           block) serves all k dz-taps; 128-channel tiles with the operand roles swapped (weights = M,
//                           voxels = N = 256), out-of-volume slices left out of the MMAs, whole sample per CTA for large batches
//   avgpool8_tile_kernel, maxpool2_kernel, pool3d_kernel, pack_mups_bf16_kernel: the pools and the input conversion
//   conv3d_pair_kernel      (conv_variant 8) / the pair mode of the z-halo kernel (conv_variant 9): CTA pairs that share every weight
//                           tile through TMA multicast -- bit-identical, measured, not faster: kept as tested variants
// The <true> instantiations of the two kernels write bf16x3 triplets (hi, lo pairs; moe_split.cu) instead of plain bf16.
// Every mbarrier wait is bounded: a barrier that never completes traps instead of hanging the GPU.
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>

#include "mups_common.cuh"

namespace mups {

constexpr int kConvThreads = 192;
constexpr int kTileM = 128;
constexpr int kTileK = 64;                 // bf16 elements per stage = one 128-byte swizzle row
constexpr int kABytes = kTileM * kTileK * 2;
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 8;

struct ConvArgs {
    int B, D, k, pl;                 // batch, volume edge, kernel edge, padding before (TF 'SAME': (k - 1) / 2)
    int kblocks;                     // ceil(Cin / 64)
    int n_tile;                      // output channels per CTA (multiple of 16, <= 256)
    int stages;
    int m_sub;                       // 128-voxel tiles per CTA (2 when N <= 128: the weight tile is loaded once for both)
    int m_tiles;                     // 128-voxel tiles of the whole batch
    int dz_box, b_box;               // z-slices / samples per 128-voxel tile
    int Cout;                        // output channels of this launch (rows of the weight tensor per tap)
    const float* scale;              // [Cout]
    const float* shift;              // [Cout]
    int relu;
    __nv_bfloat16* y;                // NDHWC
    long long y_stride;              // channels per voxel of the output tensor
    int cout_off;                    // first output channel written
    float* y_f32;                    // optional fp32 output [rows, Cout] (last layers), or nullptr
    // split destination (1^3 layers that compute two branches of an inception module from one read of the input): output
    // channels [split, Cout) go to channels [y2_off, ...) of y2 instead; ReLU only on channels below relu_upto
    __nv_bfloat16* y2;
    long long y2_stride;
    int y2_off, split, relu_upto;
    // bf16x3 output (moe_split.cu): x3_split > 0 -> every value leaves as hi = bf16(v), lo = bf16(v - hi); output channels
    // [0, x3_split) form the triplet [hi | lo | hi] of part width x3_split at cout_off, channels [x3_split, Cout) a second triplet
    // of part width Cout - x3_split right behind it
    int x3_split;
    int n_tiles, m_ctas, n_fast;     // channel tiles per voxel tile, voxel-tile CTAs; the grid is 1-D: with n_fast the channel tile is
};                                   // the FAST index, so that the CTAs that read the same activation tile run together (L2 hits)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: ~2^22 probes (each probe itself blocks for a hardware-defined interval) before giving up
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// K-major, 128-byte swizzle: 8-row atoms of 1024 bytes, SBO = 1024; version 1 (sm_100), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// bf16x3 output: channel co of this launch -> (first channel of its triplet's hi part relative to cout_off, part width)
__device__ __forceinline__ void x3_dest(const ConvArgs& a, int co, int& ych, int& w) {
    if (co < a.x3_split) { ych = co; w = a.x3_split; }
    else { ych = 3 * a.x3_split + (co - a.x3_split); w = a.Cout - a.x3_split; }
}
__device__ __forceinline__ uint32_t pack_lo2(float f0, float f1, uint32_t hi_packed) {
    const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi_packed));
    float r0 = f0 - h.x, r1 = f1 - h.y;                   // exact in fp32
    if (!(fabsf(r0) <= 3.0e38f)) r0 = 0.f;                // inf / NaN: the high part carries it alone (as moe_split.cu::split8)
    if (!(fabsf(r1) <= 3.0e38f)) r1 = 0.f;
    const __nv_bfloat162 p = __floats2bfloat162_rn(r0, r1);
    return *reinterpret_cast<const uint32_t*>(&p);
}

// Epilogue of both convolution kernels: the four epilogue warps read their 32 TMEM lanes x 16 columns at a time, apply
// scale / shift (+ ReLU), pack to bf16 and store into the channel slice of the NDHWC output (and / or the fp32 output).
template <bool X3>      // X3: bf16x3 (triplet) output; its own instantiation, so that the plain kernels' code is untouched by it
__device__ __forceinline__ void conv_epilogue(const ConvArgs& a, uint32_t bar_acc, uint32_t tmem_base, int warp, int lane, int n0, int tile0,
                                              const int (&tb0)[2], const int (&tz0)[2]) {
    // warp w may touch TMEM lanes [32 (w % 4), +32)
    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;                 // row of the tile = voxel in box order (w fastest, then h, z, sample)
    const int W = a.D, HW = a.D * a.D;
    const int x = m % W, yy = (m / W) % a.D, zl = (m / HW) % a.dz_box, bl = m / (HW * a.dz_box);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
    if (u >= a.m_sub) break;
    const long long bsample = (long long)tb0[u] + bl;
    const bool live = bsample < a.B && tile0 + u < a.m_tiles;
    const long long voxel = ((bsample * a.D + (tz0[u] + zl)) * a.D + yy) * a.D + x;
    __nv_bfloat16* yrow = a.y ? (a.y2 && n0 >= a.split ? a.y2 + voxel * a.y2_stride + a.y2_off + (n0 - a.split)
                                                        : a.y + voxel * a.y_stride + a.cout_off + n0) : nullptr;
    float* frow = a.y_f32 ? a.y_f32 + voxel * (long long)a.Cout + n0 : nullptr;
    for (int c0 = 0; c0 < a.n_tile; c0 += 32) {
        // two 16-column TMEM loads in flight before the wait (a 1^3 layer's epilogue is as long as its main loop: its latency
        // chain -- load, wait, scale / shift fetch, pack, store -- is what bounds those layers)
        uint32_t v[2][16];
        const bool second = c0 + 16 < a.n_tile;
        tmem_ld16_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(u * a.n_tile + c0), v[0]);
        if (second) tmem_ld16_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(u * a.n_tile + c0 + 16), v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !second) break;
            const int cc = c0 + 16 * h;
            float f[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int co = n0 + cc + 4 * q4;           // n0, cc multiples of 16; scale / shift are padded to Cout (a multiple of 16)
                float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
                if (co < a.Cout) {
                    sc = __ldg(reinterpret_cast<const float4*>(a.scale + co));
                    sh = __ldg(reinterpret_cast<const float4*>(a.shift + co));
                }
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float t = fmaf(__uint_as_float(v[h][4 * q4 + e]), scv[e], shv[e]);
                    f[4 * q4 + e] = (a.relu && co + e < a.relu_upto) ? fmaxf(t, 0.f) : t;
                }
            }
            if (live) {
                if (yrow) {
                    uint4 lo, hi;
                    __nv_bfloat162 p;
                    p = __floats2bfloat162_rn(f[0], f[1]);   lo.x = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[2], f[3]);   lo.y = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[4], f[5]);   lo.z = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[6], f[7]);   lo.w = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[8], f[9]);   hi.x = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[10], f[11]); hi.y = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[12], f[13]); hi.z = *reinterpret_cast<uint32_t*>(&p);
                    p = __floats2bfloat162_rn(f[14], f[15]); hi.w = *reinterpret_cast<uint32_t*>(&p);
                    if (X3) {                   // triplet output: n0 + cc is a multiple of 16, as is x3_split -- one triplet per group
                        int ych, w3;
                        x3_dest(a, n0 + cc, ych, w3);
                        __nv_bfloat16* d3 = a.y + voxel * a.y_stride + a.cout_off + ych;
                        uint4 l0, l1;
                        l0.x = pack_lo2(f[0], f[1], lo.x);   l0.y = pack_lo2(f[2], f[3], lo.y);
                        l0.z = pack_lo2(f[4], f[5], lo.z);   l0.w = pack_lo2(f[6], f[7], lo.w);
                        l1.x = pack_lo2(f[8], f[9], hi.x);   l1.y = pack_lo2(f[10], f[11], hi.y);
                        l1.z = pack_lo2(f[12], f[13], hi.z); l1.w = pack_lo2(f[14], f[15], hi.w);
                        reinterpret_cast<uint4*>(d3)[0] = lo;            reinterpret_cast<uint4*>(d3)[1] = hi;
                        reinterpret_cast<uint4*>(d3 + w3)[0] = l0;       reinterpret_cast<uint4*>(d3 + w3)[1] = l1;
                        reinterpret_cast<uint4*>(d3 + 2 * w3)[0] = lo;   reinterpret_cast<uint4*>(d3 + 2 * w3)[1] = hi;
                    } else {
                    uint4* dst = reinterpret_cast<uint4*>(yrow + cc);
                    dst[0] = lo;
                    dst[1] = hi;
                    }
                }
                if (frow) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (n0 + cc + e < a.Cout) frow[cc + e] = f[e];
                }
            }
        }
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

template <bool X3>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // 1024-byte alignment of the stages (the swizzle pattern is a function of the shared-memory address bits)
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.n_tile * kTileK * 2;
    const int a_bytes = a.m_sub * kABytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);      // full[stages], empty[stages], accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages), bar_acc = smem_u32(bars + 2 * kMaxStages);
    const uint32_t tmem_cols = a.m_sub * a.n_tile > kTmemCols ? 512u : (uint32_t)kTmemCols;   // two 256-channel accumulators: all of TMEM
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's tiles: m_sub x 128 voxels x n_tile output channels
    const int n0 = (int)(a.n_fast ? blockIdx.x % (unsigned)a.n_tiles : blockIdx.x / (unsigned)a.m_ctas) * a.n_tile;
    const int tiles_per_sample = (a.D * a.D * a.D + kTileM - 1) / kTileM;          // 4 for 8^3, else 1
    const int tile0 = (int)(a.n_fast ? blockIdx.x / (unsigned)a.n_tiles : blockIdx.x % (unsigned)a.m_ctas) * a.m_sub;
    int tb0[2], tz0[2];                          // first sample / first z-slice of each of this CTA's tiles (computed once: the
    for (int u = 0; u < 2; ++u) {                // producer issues a load every ~100 ns and cannot afford divisions)
        const int t = tile0 + u;
        tb0[u] = tiles_per_sample > 1 ? t / tiles_per_sample : t * a.b_box;
        tz0[u] = tiles_per_sample > 1 ? (t % tiles_per_sample) * a.dz_box : 0;
    }
    const int taps = a.k * a.k * a.k;
    const int iters = taps * a.kblocks;

    if (warp == 0) {
        if (elect_one()) {
            int s = 0, ph = 1, tap = 0, kb = 0, dz = 0, dy = 0, dx = 0;      // all counters advance incrementally
            for (int it = 0; it < iters; ++it) {
                mbar_wait(bar_empty + 8 * s, ph);
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(bar_full + 8 * s, (uint32_t)stage_bytes);
                tma_load_5d(dst, &map_x, bar_full + 8 * s, kb * kTileK, dx - a.pl, dy - a.pl, tz0[0] + dz - a.pl, tb0[0]);
                if (a.m_sub == 2)                       // a tile past the end of the batch is all out of bounds: zero-filled
                    tma_load_5d(dst + kABytes, &map_x, bar_full + 8 * s, kb * kTileK, dx - a.pl, dy - a.pl, tz0[1] + dz - a.pl, tb0[1]);
                tma_load_3d(dst + a_bytes, &map_w, bar_full + 8 * s, kb * kTileK, n0, tap);
                if (++s == a.stages) { s = 0; ph ^= 1; }
                if (++kb == a.kblocks) {
                    kb = 0; ++tap;
                    if (++dx == a.k) { dx = 0; if (++dy == a.k) { dy = 0; ++dz; } }
                }
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        int s = 0, ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t db = umma_desc(sa + a_bytes);
#pragma unroll
                for (int u = 0; u < 2; ++u) {                // accumulator of tile u: TMEM columns [u n_tile, (u + 1) n_tile)
                    if (u >= a.m_sub) break;
                    const uint64_t da = umma_desc(sa + u * kABytes);
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)   // UMMA_K = 16 bf16 = 32 bytes: +2 in the descriptor's 16-byte units
                        umma_bf16(tmem_base + (uint32_t)(u * a.n_tile), da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);              // frees the stage once these MMAs have read it
                if (it == iters - 1) umma_commit(bar_acc);   // accumulator complete
            }
            __syncwarp();
            if (++s == a.stages) { s = 0; ph ^= 1; }
        }
    } else {
        conv_epilogue<X3>(a, bar_acc, tmem_base, warp, lane, n0, tile0, tb0, tz0);
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- CTA-pair variant for the 1^3 layers -----------------------------------------------------------------------------------
// Those layers are bound by L2 -> SM traffic: every 128-voxel tile streams the layer's whole weight matrix again (two thirds
// of the bytes).  Here a thread-block CLUSTER of two CTAs works on two voxel tiles and ONE channel tile: each CTA loads half
// of every weight tile and the TMA unit multicasts it into both shared memories (.multicast::cluster), so a CTA fetches
// 16 KB of activations + half a weight tile per stage instead of a whole one.  A stage may be refilled once BOTH CTAs' MMAs have read
// it: the empty barriers count two arrivals, tcgen05.commit arrives on the peer's barrier too (.multicast::cluster).  The two
// CTAs stay ordinary 100 KB CTAs, so two of them still share an SM and overlap their epilogues.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w_half, const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.n_tile * kTileK * 2, half_bytes = b_bytes / 2;
    const int stage_bytes = kABytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);      // full[stages], empty[stages], accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages), bar_acc = smem_u32(bars + 2 * kMaxStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_half) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers exist before anything is multicast into this CTA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // grid: channel tile slow, voxel tile fast, so that a cluster = two consecutive voxel tiles of one channel tile
    const int n0 = (int)(blockIdx.x / (unsigned)a.m_ctas) * a.n_tile;
    const int tiles_per_sample = (a.D * a.D * a.D + kTileM - 1) / kTileM;
    const int tile0 = (int)(blockIdx.x % (unsigned)a.m_ctas);
    const int tb0[2] = {tiles_per_sample > 1 ? tile0 / tiles_per_sample : tile0 * a.b_box, 0};
    const int tz0[2] = {tiles_per_sample > 1 ? (tile0 % tiles_per_sample) * a.dz_box : 0, 0};
    const int taps = a.k * a.k * a.k;
    const int iters = taps * a.kblocks;

    if (warp == 0) {
        if (elect_one()) {
            int s = 0, ph = 1, tap = 0, kb = 0, dz = 0, dy = 0, dx = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(bar_empty + 8 * s, ph);                          // both CTAs have read this stage
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(bar_full + 8 * s, (uint32_t)stage_bytes);   // own activations + both halves of the weight tile
                tma_load_5d(dst, &map_x, bar_full + 8 * s, kb * kTileK, dx - a.pl, dy - a.pl, tz0[0] + dz - a.pl, tb0[0]);
                tma_load_3d_multicast(dst + kABytes + rank * half_bytes, &map_w_half, bar_full + 8 * s, kb * kTileK,
                                      n0 + (int)rank * (a.n_tile / 2), tap, (uint16_t)3);
                if (++s == a.stages) { s = 0; ph ^= 1; }
                if (++kb == a.kblocks) {
                    kb = 0; ++tap;
                    if (++dx == a.k) { dx = 0; if (++dy == a.k) { dy = 0; ++dz; } }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        int s = 0, ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = umma_desc(sa), db = umma_desc(sa + kABytes);
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k)
                    umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                umma_commit_multicast(bar_empty + 8 * s, (uint16_t)3);     // frees the stage in BOTH CTAs (each refills half of it)
                if (it == iters - 1) umma_commit(bar_acc);
            }
            __syncwarp();
            if (++s == a.stages) { s = 0; ph ^= 1; }
        }
    } else {
        conv_epilogue<false>(a, bar_acc, tmem_base, warp, lane, n0, tile0, tb0, tz0);
    }
    __syncthreads();
    cluster_sync_all();                 // nobody leaves while the peer may still write into this CTA's shared memory / barriers
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- z-halo variant for the 8^3 volumes (the layers that carry 85 % of the network's arithmetic) -----------------------------
// The kernel above re-reads the activation tile from L2 once per tap and sits at the L2 -> SM throughput cap.  A shift of
// the window by one z-slice is a shift by 64 rows = 8 swizzle atoms of the K-major shared-memory tile, i.e. a LEGAL operand
// start address.  So for every (dy, dx, 64-channel block) this kernel loads ONE box of 4 + k - 1 z-slices (the CTA's four
// slices = two 128-voxel tiles, plus the halo; out-of-volume slices / rows / columns zero-filled by the TMA unit) and issues
// the MMAs of all k dz-taps of both tiles from it, the descriptor start advanced by (2 u + dz) * 8 KB.  Activation traffic per
// tap falls from 32 KB to (4 + k - 1) * 8 / k KB (12.8 KB at k = 5); the weights stream through their own ring.
struct HaloArgs { int na, nb, a_box_bytes, swap, ext, nz, pair; };   // A stages, B stages, bytes of one activation box, operand roles swapped,
                                                                 // z-slices per box, output z-slices per CTA (4, or 8 = the whole sample),
                                                                 // CTA pair sharing every weight tile by TMA multicast (cluster of 2)

// Epilogue of the z-halo kernel with SWAPPED operand roles (accumulator = [128 output channels (TMEM lanes)] x [256 voxels
// (columns)]): a thread owns one output channel, a warp's store covers 32 consecutive channels of one voxel (64 bytes).
template <bool X3>
__device__ __forceinline__ void conv_epilogue_swapped(const ConvArgs& a, uint32_t bar_acc, uint32_t tmem_base, int warp, int lane, int n0,
                                                      long long voxel0, int voxels, bool live) {
    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;
    const int co = n0 + quarter * 32 + lane;
    const bool real = co < a.Cout;
    const float sc = real ? __ldg(a.scale + co) : 0.f, sh = real ? __ldg(a.shift + co) : 0.f;
    // live == false: the all-out-of-bounds partner of an odd batch's last CTA pair -- it waits for its MMAs, stores nothing
    int ych = co, w3 = 0;
    if (X3) x3_dest(a, co < a.Cout ? co : a.Cout - 1, ych, w3);
    __nv_bfloat16* ycol = (a.y && live && (real || !X3)) ? a.y + voxel0 * a.y_stride + a.cout_off + ych : nullptr;
    float* fcol = (a.y_f32 && live) ? a.y_f32 + voxel0 * (long long)a.Cout + co : nullptr;
    for (int j0 = 0; j0 < (live ? voxels : 0); j0 += 64) {           // voxels is 256 or 512: four TMEM loads in flight per wait
        uint32_t v[4][16];
#pragma unroll
        for (int h = 0; h < 4; ++h) tmem_ld16_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j0 + 16 * h), v[h]);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 4; ++h) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float t = fmaf(__uint_as_float(v[h][j]), sc, sh);
                t = a.relu ? fmaxf(t, 0.f) : t;
                if (ycol) {
                    const __nv_bfloat16 hb = __float2bfloat16_rn(t);
                    __nv_bfloat16* d = ycol + (long long)(j0 + 16 * h + j) * a.y_stride;
                    d[0] = hb;
                    if (X3) {
                        float r = t - __bfloat162float(hb);
                        if (!(fabsf(r) <= 3.0e38f)) r = 0.f;
                        d[w3] = __float2bfloat16_rn(r);
                        d[2 * w3] = hb;
                    }
                }
                if (fcol && real) fcol[(long long)(j0 + 16 * h + j) * a.Cout] = t;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

template <bool X3>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_zhalo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_half,
                    const ConvArgs a, const HaloArgs h) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.n_tile * kTileK * 2;
    unsigned char* smem_b = smem + (size_t)h.na * h.a_box_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)h.nb * b_bytes);   // a_full[4], a_empty[4], b_full[8], b_empty[8], accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kMaxStages + 1);
    const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + 4), bar_bfull = smem_u32(bars + 8),
                   bar_bempty = smem_u32(bars + 8 + kMaxStages), bar_acc = smem_u32(bars + 8 + 2 * kMaxStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = h.nz == 8 ? 512u : (uint32_t)kTmemCols;      // whole sample: two 256-column accumulators

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
        // pair mode: a weight stage may be refilled once BOTH CTAs' MMAs have read it (each CTA writes half of it into both)
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, h.pair ? 2 : 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (h.pair) cluster_sync_all();     // the peer's barriers exist before anything is multicast into this CTA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = h.pair ? cluster_ctarank() : 0u;

    // this CTA: sample blockIdx.x / 2, z-slices [4 (blockIdx.x & 1), +4) = tiles 2 blockIdx.x, 2 blockIdx.x + 1; n_tile channels at n0
    const int n0 = blockIdx.y * a.n_tile;
    // (whole-sample mode, h.nz == 8: sample blockIdx.x, all eight slices -- every weight tile then serves 512 voxels)
    const int tile0 = h.nz == 8 ? blockIdx.x * 4 : blockIdx.x * 2;
    const int sample = h.nz == 8 ? (int)blockIdx.x : (int)(blockIdx.x >> 1);
    const int zfirst = h.nz == 8 ? 0 : (int)(blockIdx.x & 1) * 4;
    const int tb0[2] = {sample, sample};
    const int tz0[2] = {zfirst, zfirst + 2};
    const int steps = a.kblocks * a.k * a.k;            // (64-channel block, dy, dx)
    // the box: h.ext z-slices starting at input slice z_box.  Untrimmed (ext = 4 + k - 1): z_box = z0 - pl, the out-of-volume
    // slices are zero-filled.  Trimmed (swapped path, whose MMAs never read an out-of-volume slice): only the in-volume slices
    // this half of the sample can reach -- [0, ext) for the lower half, [8 - ext, 8) for the upper -- 48 instead of 64 KB at
    // k = 5, which buys a third box in the ring.  Logical slice s (input slice z0 - pl + s) lies at box slice s + soff.
    const bool trimmed = h.ext < h.nz + a.k - 1;
    const int z_box = trimmed ? (tz0[0] == 0 ? 0 : a.D - h.ext) : tz0[0] - a.pl;
    const int soff = tz0[0] - a.pl - z_box;

    if (warp == 0) {
        if (elect_one()) {
            int sa = 0, pa = 1, sb = 0, pb = 1, kb = 0, dy = 0, dx = 0;
            for (int st = 0; st < steps; ++st) {
                mbar_wait(bar_aempty + 8 * sa, pa);
                mbar_expect_tx(bar_afull + 8 * sa, (uint32_t)h.a_box_bytes);
                tma_load_5d(smem_u32(smem + (size_t)sa * h.a_box_bytes), &map_x, bar_afull + 8 * sa, kb * kTileK, dx - a.pl, dy - a.pl,
                            z_box, tb0[0]);
                if (++sa == h.na) { sa = 0; pa ^= 1; }
                for (int i = 0, dz = a.pl; i < a.k; ++i, dz = dz + 1 == a.k ? 0 : dz + 1) {    // dz = pl first (see the MMA loop)
                    mbar_wait(bar_bempty + 8 * sb, pb);
                    mbar_expect_tx(bar_bfull + 8 * sb, (uint32_t)b_bytes);
                    if (h.pair)         // this CTA fetches its half of the tile's rows; the TMA unit writes it into both shared memories
                        tma_load_3d_multicast(smem_u32(smem_b + (size_t)sb * b_bytes) + crank * (uint32_t)(b_bytes >> 1), &map_w_half, bar_bfull + 8 * sb,
                                              kb * kTileK, n0 + (int)crank * (a.n_tile >> 1), (dz * a.k + dy) * a.k + dx, (uint16_t)3);
                    else
                    tma_load_3d(smem_u32(smem_b + (size_t)sb * b_bytes), &map_w, bar_bfull + 8 * sb, kb * kTileK, n0, (dz * a.k + dy) * a.k + dx);
                    if (++sb == h.nb) { sb = 0; pb ^= 1; }
                }
                if (++dx == a.k) { dx = 0; if (++dy == a.k) { dy = 0; ++kb; } }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const uint32_t idesc_swapped = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int sa = 0, pa = 0, sb = 0, pb = 0;
        for (int st = 0; st < steps; ++st) {
            mbar_wait(bar_afull + 8 * sa, pa);
            const uint32_t abase = smem_u32(smem + (size_t)sa * h.a_box_bytes);
            // taps in the order dz = pl, pl + 1, ..., k - 1, 0, ..., pl - 1: the first one reads only in-volume slices, so it
            // initialises every accumulator column and the later taps may skip their out-of-volume slices
            for (int i = 0, dz = a.pl; i < a.k; ++i, dz = dz + 1 == a.k ? 0 : dz + 1) {
                mbar_wait(bar_bfull + 8 * sb, pb);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t db = umma_desc(smem_u32(smem_b + (size_t)sb * b_bytes));
                    if (h.swap) {
                        // D^T[128 channels x 256 voxels] += W[128 x 64] * X[256 x 64]^T: ONE N = 256 MMA per K step covers both
                        // voxel tiles (box slices dz .. dz + 3 are 256 contiguous rows): 12 KB of operand reads per 128
                        // tensor cycles instead of 16 KB for two N = 128 MMAs
                        // box slice s holds input slice z0 - pl + s; tap dz reads s in [dz, dz + 4): the slices outside the
                        // volume are zeros -- leave them (whole 64-voxel column groups) out of the MMA: 15 % of the work at k = 5
                        for (int half = 0; half < (h.nz >> 2); ++half) {       // one N <= 256 MMA per group of four output slices
                            const int lo = max(dz + 4 * half, a.pl - tz0[0]), hi = min(dz + 4 * half + 4, a.D + a.pl - tz0[0]);
                            if (hi <= lo) continue;
                            const uint64_t dxv = umma_desc(abase + (uint32_t)(lo + soff) * 8192u);
                            const uint32_t idesc_n = (idesc_swapped & ~(0x3Fu << 17)) | ((uint32_t)((hi - lo) * 64 >> 3) << 17);
#pragma unroll
                            for (int k = 0; k < kTileK / 16; ++k)
                                umma_bf16(tmem_base + (uint32_t)((lo - dz) * 64), db + 2 * k, dxv + 2 * k, idesc_n, (st > 0 || i > 0 || k > 0) ? 1u : 0u);
                        }
                    } else
#pragma unroll
                    for (int u = 0; u < 2; ++u) {            // tile u = slices 2u, 2u + 1 of the CTA; tap dz reads box slices 2u + dz, +1
                        const uint64_t da = umma_desc(abase + (uint32_t)(2 * u + dz + soff) * 8192u);
#pragma unroll
                        for (int k = 0; k < kTileK / 16; ++k)
                            umma_bf16(tmem_base + (uint32_t)(u * a.n_tile), da + 2 * k, db + 2 * k, idesc, (st > 0 || i > 0 || k > 0) ? 1u : 0u);
                    }
                    if (h.pair) umma_commit_multicast(bar_bempty + 8 * sb, (uint16_t)3);   // frees the stage in both CTAs
                    else umma_commit(bar_bempty + 8 * sb);
                    if (i == a.k - 1) {
                        umma_commit(bar_aempty + 8 * sa);    // the box is free once the MMAs of its last tap have read it
                        if (st == steps - 1) umma_commit(bar_acc);
                    }
                }
                __syncwarp();
                if (++sb == h.nb) { sb = 0; pb ^= 1; }
            }
            if (++sa == h.na) { sa = 0; pa ^= 1; }
        }
    } else if (h.swap) {
        conv_epilogue_swapped<X3>(a, bar_acc, tmem_base, warp, lane, n0, (long long)tb0[0] * 512 + tz0[0] * 64, h.nz * 64, sample < a.B);
    } else {
        conv_epilogue<X3>(a, bar_acc, tmem_base, warp, lane, n0, tile0, tb0, tz0);
    }
    __syncthreads();
    if (h.pair) cluster_sync_all();     // nobody leaves while the peer may still write into this CTA's shared memory / barriers
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// MuPS fp32 [B, G, 20 S] -> bf16 [B, G, 32 S]: every scale's 20 channels padded to 32 (16-byte aligned channel slices for
// the single-scale experts' tensor maps; the pad channels meet zero weights)
__global__ void __launch_bounds__(256) pack_mups_bf16_kernel(const float* __restrict__ in, long long rows, int S, __nv_bfloat16* __restrict__ out) {
    const long long n = rows * S * 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 31);
        const long long rs = i >> 5;                       // row * S + scale
        out[i] = __float2bfloat16_rn(c < 20 ? __ldg(in + rs * 20 + c) : 0.f);
    }
}

// tf.nn.avg_pool3d(ksize k, stride 1, 'SAME') on NDHWC bf16: mean over the VALID cells of each window (the padding does not
// count, utils/tf_util.py:432-455), fp32 inside; and tf.nn.max_pool3d(2, stride 2) (:406-430).  One thread per (output voxel,
// 8-channel chunk): 16-byte loads / stores, memory bound, the window re-reads hit L1 / L2.
struct PoolEpi { const float* scale; const float* shift; int relu; int y_total, y_off; };   // scale == nullptr: plain pool, contiguous y

__global__ void __launch_bounds__(256) pool3d_kernel(const __nv_bfloat16* __restrict__ x, long long B, int D, int ct, int c_off, int c,
                                                     int k, int is_max, __nv_bfloat16* __restrict__ y, const PoolEpi ep) {
    const int chunks = c >> 3;
    const int Do = is_max ? D / 2 : D;
    const long long n = B * Do * Do * Do * chunks;
    const int pl = (k - 1) / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        long long v = i / chunks;
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
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (((b * D + z) * D + yy) * D + xx) * (long long)ct + c_off) + ch);
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __bfloat1622float2(h[j]);
                        if (is_max) { acc[2 * j] = fmaxf(acc[2 * j], f.x); acc[2 * j + 1] = fmaxf(acc[2 * j + 1], f.y); }
                        else { acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
                    }
                    ++cnt;
                }
            }
        }
        const float inv = is_max ? 1.f : 1.f / (float)cnt;
        uint4 out;
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&out);
        if (ep.scale) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(acc[j] * inv, __ldg(ep.scale + ch * 8 + j), __ldg(ep.shift + ch * 8 + j));
                acc[j] = ep.relu ? fmaxf(t, 0.f) : t;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
            reinterpret_cast<uint4*>(y + v * (long long)ep.y_total + ep.y_off)[ch] = out;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
            reinterpret_cast<uint4*>(y + v * (long long)c)[ch] = out;
        }
    }
}

// The same average pool for the 8^3 volumes of the reference's networks (the 27-tap version above re-reads every input
// voxel 27 times through L2: 20 % of the whole forward pass).  One CTA = one sample x 32 channels: the 512 x 32 tile is staged
// once in shared memory as fp32 (64 KB) and the box sum is done SEPARABLY in place -- three passes of 64 lines of 8 voxels,
// one warp per line, lane = channel (bank-conflict free) -- then scaled by 1 / (valid cells) = 1 / (cx cy cz) and stored.
// HBM traffic: every input and output byte once.
template <int K>
__global__ void __launch_bounds__(256) avgpool8_tile_kernel(const __nv_bfloat16* __restrict__ x, int ct, int c_off, int c,
                                                            __nv_bfloat16* __restrict__ y, const PoolEpi ep) {
    extern __shared__ __align__(16) float tile[];                  // [512 voxels][32 channels]
    constexpr int D = 8, PL = (K - 1) / 2;
    const int chunks = c >> 5;
    const long long b = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __nv_bfloat16* src = x + b * 512 * (long long)ct + c_off + chunk * 32;
    uint4 raw[8];                                                  // 16-byte loads, 4 per voxel: all eight in flight at once
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = tid + r * 256;
        raw[r] = __ldg(reinterpret_cast<const uint4*>(src + (i >> 2) * (long long)ct) + (i & 3));
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = tid + r * 256, v = i >> 2, part = i & 3;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[r]);
        const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
        float4* dst = reinterpret_cast<float4*>(tile + v * 32 + part * 8);
        dst[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
        dst[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        // line l of pass 0 (along x): voxels (l * 8 + i); pass 1 (along y): z = l / 8, x = l % 8; pass 2 (along z): y = l / 8, x = l % 8
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
    const int y_stride = ep.scale ? ep.y_total : c;
    __nv_bfloat16* dst = y + b * 512 * (long long)y_stride + (ep.scale ? ep.y_off : 0) + chunk * 32;
    for (int i = tid; i < 512 * 4; i += 256) {
        const int v = i >> 2, part = i & 3;
        const int vx = v & 7, vy = (v >> 3) & 7, vz = v >> 6;
        auto valid = [](int q) { return min(q - PL + K - 1, D - 1) - max(q - PL, 0) + 1; };
        const float inv = 1.f / (float)(valid(vx) * valid(vy) * valid(vz));
        const float4* sp = reinterpret_cast<const float4*>(tile + v * 32 + part * 8);
        const float4 a0 = sp[0], a1 = sp[1];
        float f[8] = {a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv, a1.x * inv, a1.y * inv, a1.z * inv, a1.w * inv};
        if (ep.scale) {                       // folded bias + batch norm (+ ReLU) of the 1^3 convolution that ran BEFORE this pool
            const int c0 = chunk * 32 + part * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(f[j], __ldg(ep.scale + c0 + j), __ldg(ep.shift + c0 + j));
                f[j] = ep.relu ? fmaxf(t, 0.f) : t;
            }
        }
        uint4 out;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
        for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        reinterpret_cast<uint4*>(dst + v * (long long)y_stride)[part] = out;
    }
}

template <int K>
static cudaError_t launch_avgpool8(const __nv_bfloat16* x, long long B, int ct, int c_off, int c, __nv_bfloat16* y, const PoolEpi& ep, cudaStream_t st) {
    constexpr int smem = 512 * 32 * 4;
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(avgpool8_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    avgpool8_tile_kernel<K><<<(unsigned)(B * (c >> 5)), 256, smem, st>>>(x, ct, c_off, c, y, ep);
    return cudaGetLastError();
}

// tf.nn.max_pool3d(2, stride 2) on NDHWC bf16 (utils/tf_util.py:406-430; D even, so 'SAME' needs no padding): one thread per
// (output voxel, 8-channel chunk), the eight 16-byte loads of its window issued back to back, 32-bit index arithmetic.
__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16* __restrict__ x, unsigned n, int D, int ct, int c_off, int c,
                                                       __nv_bfloat16* __restrict__ y) {
    const unsigned chunks = (unsigned)c >> 3, Do = (unsigned)D >> 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned ch = i % chunks, v = i / chunks;
        const unsigned xo = v % Do, yo = (v / Do) % Do, zo = (v / (Do * Do)) % Do, b = v / (Do * Do * Do);
        const size_t vin = (((size_t)b * D + 2 * zo) * D + 2 * yo) * D + 2 * xo;
        const uint4* p = reinterpret_cast<const uint4*>(x + vin * ct + c_off) + ch;
        const size_t sx = (size_t)ct / 8, sy = sx * D, sz = sy * D;            // strides in 16-byte units
        uint4 r[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) r[t] = __ldg(p + (t & 1) * sx + ((t >> 1) & 1) * sy + (t >> 2) * sz);
        uint4 out;
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 m = reinterpret_cast<const __nv_bfloat162*>(&r[0])[j];
#pragma unroll
            for (int t = 1; t < 8; ++t) m = __hmax2(m, reinterpret_cast<const __nv_bfloat162*>(&r[t])[j]);
            o[j] = m;
        }
        reinterpret_cast<uint4*>(y + (size_t)v * c)[ch] = out;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace mups

using namespace mups;

extern "C" {

int mups_moe_pack_input(const float* mups_dev, int64_t rows, int S, void* out_bf16_dev, mups_stream stream) {
    MUPS_REQUIRE(rows >= 0 && S >= 1 && S <= MUPS_MAX_SCALES, "mups_moe_pack_input: rows=%lld, S=%d out of range", (long long)rows, S);
    MUPS_REQUIRE(rows == 0 || (mups_dev && out_bf16_dev), "mups_moe_pack_input: NULL buffer");
    if (rows == 0) return MUPS_OK;
    const long long n = (long long)rows * S * 32;
    const int grid = (int)((n + 255) / 256 < 16 * kNumSMs ? (n + 255) / 256 : 16 * kNumSMs);
    pack_mups_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mups_dev, rows, S, static_cast<__nv_bfloat16*>(out_bf16_dev));
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

static int pool_launch(const char* who, const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int c, int k, int is_max,
                       void* y_bf16_dev, const PoolEpi& ep, mups_stream stream) {
    MUPS_REQUIRE(x_bf16_dev && y_bf16_dev, "%s: NULL buffer", who);
    MUPS_REQUIRE(B >= 1 && (D == 2 || D == 4 || D == 8), "%s: B=%lld, volume edge %d", who, (long long)B, D);
    MUPS_REQUIRE(c >= 8 && c % 8 == 0 && c_total % 8 == 0 && c_off % 8 == 0 && c_off + c <= c_total, "%s: channels (%d of %d at %d) must be multiples of 8", who, c, c_total, c_off);
    MUPS_REQUIRE(is_max ? k == 2 : (k >= 1 && k <= 5), "%s: window %d", who, k);
    if (!is_max && D == 8 && k >= 2 && c % 32 == 0 && (long long)B * (c >> 5) <= 0x7FFFFFFFll && g_pool_variant.load() != 1) {
        // shared-memory tile, separable box sum (every byte read once)
        const auto* xs = static_cast<const __nv_bfloat16*>(x_bf16_dev);
        auto* ys = static_cast<__nv_bfloat16*>(y_bf16_dev);
        const cudaStream_t st = static_cast<cudaStream_t>(stream);
        cudaError_t e = k == 2 ? launch_avgpool8<2>(xs, B, c_total, c_off, c, ys, ep, st) : k == 3 ? launch_avgpool8<3>(xs, B, c_total, c_off, c, ys, ep, st)
                      : k == 4 ? launch_avgpool8<4>(xs, B, c_total, c_off, c, ys, ep, st) : launch_avgpool8<5>(xs, B, c_total, c_off, c, ys, ep, st);
        if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return MUPS_ERR_CUDA; }
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return MUPS_OK;
    }
    if (is_max && (long long)B * D * D * D * (c_total / 8) < 0xFFFFFFFFll && g_pool_variant.load() != 1) {
        const unsigned n = (unsigned)((long long)B * (D / 2) * (D / 2) * (D / 2) * (c / 8));
        const unsigned grid = (n + 255) / 256;
        maxpool2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16_dev), n, D, c_total, c_off, c,
                                                                               static_cast<__nv_bfloat16*>(y_bf16_dev));
        MUPS_CHECK_LAUNCH();
        return MUPS_OK;
    }
    const int Do = is_max ? D / 2 : D;
    const long long n = (long long)B * Do * Do * Do * (c / 8);
    const int grid = (int)((n + 255) / 256 < 32 * kNumSMs ? (n + 255) / 256 : 32 * kNumSMs);
    pool3d_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16_dev), B, D, c_total, c_off, c, k,
                                                                         is_max, static_cast<__nv_bfloat16*>(y_bf16_dev), ep);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

int mups_pool3d(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int c, int k, int is_max, void* y_bf16_dev,
                mups_stream stream) {
    return pool_launch("mups_pool3d", x_bf16_dev, B, D, c_total, c_off, c, k, is_max, y_bf16_dev, PoolEpi{nullptr, nullptr, 0, 0, 0}, stream);
}

int mups_avgpool3d_bn_relu(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int c, int k, const float* scale_dev,
                           const float* shift_dev, int relu, void* y_bf16_dev, int y_total, int y_off, mups_stream stream) {
    MUPS_REQUIRE(scale_dev && shift_dev, "mups_avgpool3d_bn_relu: NULL scale / shift");
    MUPS_REQUIRE(k >= 2, "mups_avgpool3d_bn_relu: window %d (a 1-wide pool is the identity: use the convolution's own epilogue)", k);
    MUPS_REQUIRE(y_total % 8 == 0 && y_off % 8 == 0 && y_off + c <= y_total, "mups_avgpool3d_bn_relu: output channel slice (%d of %d at %d)", c, y_total, y_off);
    return pool_launch("mups_avgpool3d_bn_relu", x_bf16_dev, B, D, c_total, c_off, c, k, 0, y_bf16_dev,
                       PoolEpi{scale_dev, shift_dev, relu, y_total, y_off}, stream);
}

struct ConvSplit { void* y2; int y2_total, y2_off, split, relu_upto; int x3_split; };   // y2 == nullptr, x3_split > 0: bf16x3 output

static int conv_launch(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                       int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev, int relu, void* y_bf16_dev,
                       int cout_total, int cout_off, float* y_f32_dev, const ConvSplit* sp, mups_stream stream);

int mups_conv3d_bn_relu(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                        int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev, int relu, void* y_bf16_dev,
                        int cout_total, int cout_off, float* y_f32_dev, mups_stream stream) {
    return conv_launch(x_bf16_dev, B, D, cin_total, cin_off, cin, w_bf16_dev, cin_w, cout, k, scale_dev, shift_dev, relu, y_bf16_dev,
                       cout_total, cout_off, y_f32_dev, nullptr, stream);
}

int mups_conv3d_bn_relu_x3(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                           int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev, int relu, void* y_bf16_dev,
                           int cout_total, int cout_off, int split, mups_stream stream) {
    MUPS_REQUIRE(y_bf16_dev, "mups_conv3d_bn_relu_x3: NULL output");
    MUPS_REQUIRE(split >= 16, "mups_conv3d_bn_relu_x3: split %d of %d output channels (multiples of 16)", split, cout);
    const ConvSplit sp{nullptr, 0, 0, 0, 0, split};
    return conv_launch(x_bf16_dev, B, D, cin_total, cin_off, cin, w_bf16_dev, cin_w, cout, k, scale_dev, shift_dev, relu, y_bf16_dev,
                       cout_total, cout_off, nullptr, &sp, stream);
}

int mups_conv1_split_bn_relu(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                             int cin_w, int cout, const float* scale_dev, const float* shift_dev, int relu_upto, void* y1_bf16_dev,
                             int y1_total, int y1_off, int split, void* y2_bf16_dev, int y2_total, int y2_off, mups_stream stream) {
    MUPS_REQUIRE(y1_bf16_dev && y2_bf16_dev, "mups_conv1_split_bn_relu: NULL output");
    MUPS_REQUIRE(split >= 16 && split % 16 == 0 && split < cout && (cout - split) % 16 == 0, "mups_conv1_split_bn_relu: split %d of %d channels", split, cout);
    MUPS_REQUIRE(y2_total % 8 == 0 && y2_off % 8 == 0 && y2_off + (cout - split) <= y2_total && y1_off + split <= y1_total,
                 "mups_conv1_split_bn_relu: output channel slices");
    MUPS_REQUIRE(relu_upto >= 0 && relu_upto <= cout, "mups_conv1_split_bn_relu: relu_upto %d", relu_upto);
    const ConvSplit sp{y2_bf16_dev, y2_total, y2_off, split, relu_upto, 0};
    return conv_launch(x_bf16_dev, B, D, cin_total, cin_off, cin, w_bf16_dev, cin_w, cout, 1, scale_dev, shift_dev, relu_upto > 0, y1_bf16_dev,
                       y1_total, y1_off, nullptr, &sp, stream);
}

static int conv_launch(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                       int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev, int relu, void* y_bf16_dev,
                       int cout_total, int cout_off, float* y_f32_dev, const ConvSplit* sp_in, mups_stream stream) {
    const int x3_split = (sp_in && !sp_in->y2) ? sp_in->x3_split : 0;
    const ConvSplit* sp = (sp_in && sp_in->y2) ? sp_in : nullptr;
    MUPS_REQUIRE(x_bf16_dev && w_bf16_dev && scale_dev && shift_dev && (y_bf16_dev || y_f32_dev), "mups_conv3d_bn_relu: NULL buffer");
    MUPS_REQUIRE(B >= 1 && B < (1ll << 30), "mups_conv3d_bn_relu: B=%lld out of range", (long long)B);
    MUPS_REQUIRE(D == 1 || D == 2 || D == 4 || D == 8, "mups_conv3d_bn_relu: volume edge %d (1, 2, 4 or 8)", D);
    MUPS_REQUIRE(k >= 1 && k <= 5 && (D > 1 || k == 1), "mups_conv3d_bn_relu: kernel edge %d", k);
    MUPS_REQUIRE(cin >= 8 && cin % 8 == 0 && cin_total % 8 == 0 && cin_off % 8 == 0 && cin_off + cin <= cin_total,
                 "mups_conv3d_bn_relu: input channels (%d of %d at %d) must be multiples of 8", cin, cin_total, cin_off);
    MUPS_REQUIRE(cin_w >= cin && cin_w % 8 == 0, "mups_conv3d_bn_relu: weight inner dimension %d", cin_w);
    MUPS_REQUIRE(cout >= 16 && cout % 16 == 0, "mups_conv3d_bn_relu: output channels %d must be a multiple of 16", cout);
    MUPS_REQUIRE(!y_bf16_dev || (cout_total % 8 == 0 && cout_off % 8 == 0 && cout_off + (x3_split ? 3 * cout : sp ? sp->split : cout) <= cout_total),
                 "mups_conv3d_bn_relu: output channel slice (%d of %d at %d)", x3_split ? 3 * cout : cout, cout_total, cout_off);
    MUPS_REQUIRE(!x3_split || (y_bf16_dev && x3_split >= 16 && x3_split % 16 == 0 && x3_split <= cout),
                 "mups_conv3d_bn_relu_x3: split %d of %d output channels (multiples of 16)", x3_split, cout);
    MUPS_REQUIRE(((reinterpret_cast<uintptr_t>(scale_dev) | reinterpret_cast<uintptr_t>(shift_dev)) & 15) == 0,
                 "mups_conv3d_bn_relu: scale / shift must be 16-byte aligned");
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("mups_conv3d_bn_relu: cuTensorMapEncodeTiled is not available (driver too old?)"); return MUPS_ERR_CUDA; }

    ConvArgs a;
    a.B = (int)B; a.D = D; a.k = k; a.pl = (k - 1) / 2;
    a.kblocks = (cin + kTileK - 1) / kTileK;
    int n_tile = cout;
    if (n_tile > 256) { n_tile = 256; while (cout % n_tile) n_tile -= 16; }
    if (sp) {                                   // a channel tile must not straddle the two destinations
        n_tile = sp->split < 256 ? sp->split : 256;
        while (sp->split % n_tile || (cout - sp->split) % n_tile) n_tile -= 16;
    }
    if (const char* e = getenv("MUPS_CONV_N1")) {          // experiment: narrower channel tiles (deeper pipelines) for the 1^3 layers
        const int want = atoi(e);
        if (k == 1 && D > 1 && want >= 16 && want < n_tile && n_tile % want == 0 && (!sp || sp->split % want == 0)) n_tile = want;
    }
    a.n_tile = n_tile;
    a.y2 = sp ? static_cast<__nv_bfloat16*>(sp->y2) : nullptr;
    a.y2_stride = sp ? sp->y2_total : 0; a.y2_off = sp ? sp->y2_off : 0; a.split = sp ? sp->split : cout;
    a.relu_upto = sp ? sp->relu_upto : cout;
    a.x3_split = x3_split;
    const int vox = D * D * D;
    a.dz_box = vox >= kTileM ? kTileM / (D * D) : D;
    a.b_box = vox >= kTileM ? 1 : kTileM / vox;
    a.Cout = cout;
    a.scale = scale_dev; a.shift = shift_dev; a.relu = relu;
    a.y = static_cast<__nv_bfloat16*>(y_bf16_dev); a.y_stride = cout_total; a.cout_off = cout_off; a.y_f32 = y_f32_dev;
    const long long m_tiles = vox >= kTileM ? (long long)B * (vox / kTileM) : (B + a.b_box - 1) / a.b_box;
    MUPS_REQUIRE(m_tiles <= 0x7FFFFFFFll, "mups_conv3d_bn_relu: batch too large for one launch");
    a.m_tiles = (int)m_tiles;
    // two voxel tiles per CTA share one weight tile when both accumulators fit the 256 TMEM columns and the layer is deep
    // enough for operand delivery to matter (g_conv_m_sub: benchmarking override)
    a.m_sub = (n_tile <= 128 && m_tiles >= 2 * kNumSMs && k > 1) ? 2 : 1;
    // conv_variant 7 (experiment): 1^3 layers with 256-channel tiles take two voxel tiles per CTA (all 512 TMEM columns, one CTA per
    // SM): a third less L2 -> SM traffic per output (the weight tile is streamed once per 256 voxels), no epilogue overlap
    const bool wide1 = k == 1 && n_tile == 256 && m_tiles >= 4 * kNumSMs && g_conv_variant.load() == 7;
    if (wide1) a.m_sub = 2;
    if (k > 1 && D < 8 && ((m_tiles + a.m_sub - 1) / a.m_sub) * (cout / n_tile) < kNumSMs / 4 && g_conv_variant.load() != 5) {
        // nearly empty grids (a 2^3 layer of a 256-query batch is 16 tiles -- 16 CTAs on 148 SMs; measured: +2.5 % on the whole
        // forward at batch 256, while splitting grids of 64 or more CTAs gains nothing -- their K loop is latency-bound).  Pick the
        // (channel tile, voxel tiles per CTA) with the lowest estimated time = waves x per-CTA cost (relative time per
        // 64-channel stage: N = 256 1.0, N = 128 0.62, N = 64 0.5 -- the narrow tiles pay more operand reads and more L2
        // traffic per FLOP, measured: splitting a grid that already fills the machine loses), and only when it wins by 15 %
        const int cands[3] = {n_tile, 128, 64};
        const double rel[3] = {n_tile >= 256 ? 1.0 : n_tile >= 128 ? 0.62 : 0.5 * n_tile / 64.0, 0.62, 0.5};
        int best_n = n_tile, best_ms = a.m_sub;
        const long long ctas0 = ((m_tiles + a.m_sub - 1) / a.m_sub) * (cout / n_tile);
        double best = 0.85 * (double)((ctas0 + kNumSMs - 1) / kNumSMs) * a.m_sub * rel[0] * (a.m_sub == 2 ? 0.95 : 1.0);
        for (int c = 1; c < 3; ++c) {
            const int n = cands[c];
            if (n >= n_tile || cout % n) continue;
            for (int ms = 1; ms <= 2; ++ms) {
                const long long ctas = ((m_tiles + ms - 1) / ms) * (cout / n);
                const double t = (double)((ctas + kNumSMs - 1) / kNumSMs) * ms * rel[c] * (ms == 2 ? 0.95 : 1.0);
                if (t < best - 1e-9) { best = t; best_n = n; best_ms = ms; }
            }
        }
        n_tile = best_n; a.n_tile = n_tile; a.m_sub = best_ms;
    }
    if (const char* e = getenv("MUPS_CONV_M_SUB")) a.m_sub = (atoi(e) == 2 && n_tile <= 128) ? 2 : 1;
    const int stage_bytes = a.m_sub * kABytes + n_tile * kTileK * 2;
    int stages = (200 * 1024) / stage_bytes;
    // short K (the 1^3 layers: 2-24 stages of work per tile): the epilogue is as long as the main loop and nothing overlaps it
    // inside one CTA, so leave room for TWO CTAs per SM (<= 100 KB of shared memory and 256 TMEM columns each) -- one CTA's
    // epilogue then runs under the other's loads and MMAs
    const int iters_total = k * k * k * a.kblocks;
    if (iters_total <= 24 && m_tiles > kNumSMs && g_conv_variant.load() != 1 && !wide1) stages = (100 * 1024) / stage_bytes < 2 ? 2 : (100 * 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > iters_total) stages = iters_total < 2 ? 2 : iters_total;
    a.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * kMaxStages + 1) * 8 + 16;
    if (smem < 80 * 1024) smem = 80 * 1024;      // never more than two CTAs per SM: each allocates 256 of the 512 TMEM columns

    // z-halo kernel: 8^3 volumes, k > 1, both accumulators in TMEM (conv_variant 2 forces the per-tap kernel)
    const bool zhalo = D == 8 && k >= 2 && n_tile <= 128 && g_conv_variant.load() != 2;
    HaloArgs h{0, 0, 0, 0, 0, 4, 0};
    if (zhalo) {
        a.m_sub = 2;
        h.swap = (n_tile == 128 && g_conv_variant.load() != 3) ? 1 : 0;     // conv_variant 3: z-halo without the operand swap
        // whole sample per CTA (weights read once per 512 voxels: L2 -> SM traffic -40 %) when the batch still fills the
        // machine four times over; conv_variant 4 forces the half-sample CTAs
        h.nz = (h.swap && B >= 4 * kNumSMs && g_conv_variant.load() != 4) ? 8 : 4;
        h.ext = h.nz == 8 ? 8 : h.swap ? 4 + (k - 1 - a.pl) : 4 + k - 1;    // swapped path: in-volume slices only (see the kernel)
        h.a_box_bytes = h.ext * 8192;
        const int b_bytes = n_tile * kTileK * 2;
        const size_t budget = 224 * 1024;                                   // of the 227 KB a CTA may have
        h.na = (3 * (size_t)h.a_box_bytes + 4 * (size_t)b_bytes <= budget) ? 3 : 2;
        if (const char* e = getenv("MUPS_CONV_NA")) h.na = atoi(e) == 3 ? 3 : 2;       // benchmarking override
        h.nb = (int)((budget - (size_t)h.na * h.a_box_bytes) / b_bytes);
        if (const char* e = getenv("MUPS_CONV_NB")) h.nb = atoi(e) < h.nb ? (atoi(e) < 2 ? 2 : atoi(e)) : h.nb;
        if (h.nb > kMaxStages) h.nb = kMaxStages;
        smem = (size_t)h.na * h.a_box_bytes + (size_t)h.nb * b_bytes + 1024 + (8 + 2 * kMaxStages + 1) * 8 + 16;
    }

    CUtensorMap map_x, map_w;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)cin, (cuuint64_t)D, (cuuint64_t)D, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t strides[4] = {(cuuint64_t)cin_total * 2, (cuuint64_t)cin_total * 2 * D, (cuuint64_t)cin_total * 2 * D * D,
                                       (cuuint64_t)cin_total * 2 * D * D * D};
        const cuuint32_t box[5] = {(cuuint32_t)kTileK, (cuuint32_t)D, (cuuint32_t)D, (cuuint32_t)(zhalo ? h.ext : a.dz_box), (cuuint32_t)a.b_box};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        void* base = const_cast<unsigned char*>(static_cast<const unsigned char*>(x_bf16_dev)) + (size_t)cin_off * 2;
        const CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("mups_conv3d_bn_relu: activation tensor map rejected (CUresult %d)", (int)r); return MUPS_ERR_CUDA; }
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)cin_w, (cuuint64_t)cout, (cuuint64_t)(k * k * k)};
        const cuuint64_t strides[2] = {(cuuint64_t)cin_w * 2, (cuuint64_t)cin_w * 2 * cout};
        const cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)n_tile, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        const CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_bf16_dev), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("mups_conv3d_bn_relu: weight tensor map rejected (CUresult %d)", (int)r); return MUPS_ERR_CUDA; }
    }
    if (zhalo) {
        const auto zkernel = x3_split ? conv3d_zhalo_kernel<true> : conv3d_zhalo_kernel<false>;
        MUPS_CUDA_TRY(cudaFuncSetAttribute(zkernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unsigned gx = (unsigned)(h.nz == 8 ? m_tiles / 4 : m_tiles / 2);
        // conv_variant 9: CTA pairs (clusters of two CTAs: two samples, or the two halves of one) that share every weight tile -- each
        // CTA fetches half of its rows and the TMA unit multicasts them into both shared memories: a fifth to a quarter less L2 -> SM
        // traffic per CTA (the weights are 3 x 16 of 112 KB per activation box at k = 3, 5 x 16 of 144 KB at k = 5)
        h.pair = (g_conv_variant.load() == 9 && gx >= 2 * kNumSMs) ? 1 : 0;
        if (!h.pair) {
            zkernel<<<dim3(gx, (unsigned)(cout / n_tile)), kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(map_x, map_w, map_w, a, h);
            MUPS_CHECK_LAUNCH();
            return MUPS_OK;
        }
        CUtensorMap map_wh;
        {
            const cuuint64_t dims[3] = {(cuuint64_t)cin_w, (cuuint64_t)cout, (cuuint64_t)(k * k * k)};
            const cuuint64_t strides[2] = {(cuuint64_t)cin_w * 2, (cuuint64_t)cin_w * 2 * cout};
            const cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)(n_tile / 2), 1};
            const cuuint32_t es[3] = {1, 1, 1};
            const CUresult r = enc(&map_wh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_bf16_dev), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("mups_conv3d_bn_relu: half weight tensor map rejected (CUresult %d)", (int)r); return MUPS_ERR_CUDA; }
        }
        gx = (gx + 1) & ~1u;                                 // an odd batch's last CTA gets an all-out-of-bounds partner
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(gx, (unsigned)(cout / n_tile));
        cfg.blockDim = dim3(kConvThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        MUPS_CUDA_TRY(cudaLaunchKernelEx(&cfg, zkernel, map_x, map_w, map_wh, a, h));
        MUPS_CHECK_LAUNCH();
        return MUPS_OK;
    }
    const auto tkernel = x3_split ? conv3d_tcgen05_kernel<true> : conv3d_tcgen05_kernel<false>;
    MUPS_CUDA_TRY(cudaFuncSetAttribute(tkernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    a.n_tiles = cout / n_tile;
    a.m_ctas = (int)((m_tiles + a.m_sub - 1) / a.m_sub);
    if (k == 1 && a.m_sub == 1 && n_tile >= 32 && n_tile % 32 == 0 && m_tiles >= 4 * kNumSMs && g_conv_variant.load() == 8 && !x3_split) {
        // CTA pairs sharing the weight tile by TMA multicast (see conv3d_pair_kernel)
        CUtensorMap map_wh;
        const cuuint64_t dims[3] = {(cuuint64_t)cin_w, (cuuint64_t)cout, (cuuint64_t)(k * k * k)};
        const cuuint64_t strides[2] = {(cuuint64_t)cin_w * 2, (cuuint64_t)cin_w * 2 * cout};
        const cuuint32_t box[3] = {(cuuint32_t)kTileK, (cuuint32_t)(n_tile / 2), 1};
        const cuuint32_t es[3] = {1, 1, 1};
        const CUresult r = enc(&map_wh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_bf16_dev), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("mups_conv3d_bn_relu: half weight tensor map rejected (CUresult %d)", (int)r); return MUPS_ERR_CUDA; }
        a.m_ctas = (a.m_ctas + 1) & ~1;                      // an odd tail tile gets an all-out-of-bounds partner
        MUPS_CUDA_TRY(cudaFuncSetAttribute(conv3d_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((long long)a.m_ctas * a.n_tiles));
        cfg.blockDim = dim3(kConvThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        MUPS_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3d_pair_kernel, map_x, map_wh, a));
        MUPS_CHECK_LAUNCH();
        return MUPS_OK;
    }
    a.n_fast = getenv("MUPS_CONV_NSLOW") ? 0 : 1;            // benchmarking override: channel tile as the slow index
    MUPS_REQUIRE(((m_tiles + a.m_sub - 1) / a.m_sub) * a.n_tiles <= 0x7FFFFFFFll, "mups_conv3d_bn_relu: grid too large");
    tkernel<<<(unsigned)(((m_tiles + a.m_sub - 1) / a.m_sub) * a.n_tiles), kConvThreads, smem,
                            static_cast<cudaStream_t>(stream)>>>(map_x, map_w, a);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

}  // extern "C"
