// Shared declarations of the MuPS sm_100a library (internal; the public ABI is include/mups.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/mups.h"

namespace mups {

constexpr int kNumSMs = 148;  // B200

// ---- error plumbing (thread-local message, never throws across the ABI) -------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;
extern std::atomic<int> g_boundary_cap;
extern std::atomic<int> g_fuse_candidates;
extern std::atomic<int> g_stats_variant;
extern std::atomic<int> g_query_kernel;     // 0 automatic, 1 flat cell scan, 2 hierarchical (octree over the Morton codes)
extern std::atomic<int> g_hier_margin;      // -1 automatic (6 sqrt(P) + 8); tests set 0 to exercise the worklist hand-over
extern std::atomic<int> g_pool_variant;     // 0 automatic (shared-memory tile for the 8^3 average pools), 1 per-voxel kernel everywhere
extern std::atomic<int> g_conv_variant;     // 0 automatic (two CTAs per SM for short-K layers), 1 always one CTA per SM
extern std::atomic<int> g_query_order;      // 0 automatic, 1 CTAs in caller order, 2 CTAs in Morton order of the centres

#define MUPS_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::mups::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                              __LINE__);                                                      \
            return (_e == cudaErrorMemoryAllocation) ? MUPS_ERR_NOMEM : MUPS_ERR_CUDA;        \
        }                                                                                     \
    } while (0)

#define MUPS_CHECK_LAUNCH()                                                                   \
    do {                                                                                      \
        ::mups::g_launch_count.fetch_add(1, std::memory_order_relaxed);                       \
        MUPS_CUDA_TRY(cudaGetLastError());                                                    \
    } while (0)

#define MUPS_REQUIRE(cond, ...)                                                               \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::mups::set_error(__VA_ARGS__);                                                   \
            return MUPS_ERR_INVALID;                                                          \
        }                                                                                     \
    } while (0)

// ---- spatial index ---------------------------------------------------------------------------
// Device-resident grid description, filled by the bbox kernel (the host never needs it to
// enqueue the build, so index creation is fully asynchronous).
struct GridDesc {
    float bb_min[3];
    float bb_max[3];
    float origin[3];   // == bb_min
    float inv_cell;    // 1 / cell edge
    float cell;        // cell edge (>= cell_frac * |bbox diagonal|)
    int dims[3];       // cells per axis (<= max_dim)
};

}  // namespace mups

struct mups_index {
    int device = 0;
    int64_t n = 0;
    int bits = 0;             // Morton bits per axis; code space = 1 << (3*bits)
    int max_dim = 0;
    double cell_frac = 0.07;
    mups::GridDesc* grid = nullptr;      // device
    float4* sorted = nullptr;            // device [n]: (x, y, z, original index as int bits), Morton-cell order
    uint32_t* cell_start = nullptr;      // device [ncode + 1]: exclusive prefix of cell populations
    int32_t* pos_of = nullptr;           // device [n]: position in `sorted` of original point i
    uint32_t* codes = nullptr;           // device [n]: Morton cell code per original point (build scratch, kept)
    uint32_t* hash_sorted = nullptr;     // device [n]: point_hash(original index of sorted[i]) (4 B/point: all the hierarchical
                                         // query reads of a point in a cell wholly inside a ball)
    cudaStream_t build_stream = nullptr;
    cudaEvent_t built = nullptr;
    // host copy of the bbox, fetched lazily by mups_index_bbox
    mutable bool have_bbox = false;
    mutable float bb[6] = {0, 0, 0, 0, 0, 0};
    // streams that have queried the index (one re-recorded event each): destroy orders the frees after them
    mutable std::mutex use_mutex;
    mutable std::vector<std::pair<cudaStream_t, cudaEvent_t>> uses;
};

struct mups_gmm {
    int device = 0;
    int G = 0;
    int separable = 0;
    int res[3] = {0, 0, 0};   // lattice resolution per axis when separable
    // general path, per Gaussian (device, float4 each):
    float4* A = nullptr;      // (mu_x, mu_y, mu_z, log2(w / ((2pi)^1.5 sigma_x^3)))        [masked prefactor, tf_util.py:687]
    float4* Bv = nullptr;     // (1/sigma_x, 1/sigma_y, 1/sigma_z, log2(w / ((2pi)^1.5 sx sy sz))) [MultivariateNormalDiag]
    float4* C = nullptr;      // (w, 1/sqrt(w), 1/sqrt(2w), -w/sqrt(w))
    // separable path: per axis lattice coordinates and 1/sigma (device, [3][max res] floats) + uniform weight
    float* axis_mu = nullptr;   // [3*64]
    float axis_isig[3] = {0, 0, 0};
    float guard_lo[3] = {0, 0, 0}, guard_hi[3] = {0, 0, 0};   // lattice hull widened by 5 sigma per axis
    float w_uniform = 0;
};

namespace mups {

// library-private stream-ordered memory pool of a device (release threshold = keep everything): index
// buffers and scratch are allocated and freed without touching the OS or synchronising the device
int library_pool(int device, cudaMemPool_t* pool);

// kernels' host launchers (defined in the .cu files)
int launch_index_build(mups_index* ix, const float* xyz, cudaStream_t st);
int launch_ball_query(const mups_index* ix, const int64_t* q, int64_t B, const double* r_abs, int S, int P,
                      uint64_t seed, int32_t* nbr_idx, int32_t* nbr_total, float* patches, int32_t* n_eff,
                      int32_t* nbr_pos, cudaStream_t st);
// in-place exclusive scan of n uint32 counters; tile_sums: scratch of (n + 2047) / 2048 entries
int launch_exclusive_scan(uint32_t* data, int64_t n, uint32_t* tile_sums, cudaStream_t st);
// K6: where the statistics kernel finds the patch points when no patch tensor exists
struct GatherSource {
    const mups_index* index;
    const int64_t* q;          // [B] centre indices (device)
    const int32_t* nbr_pos;    // [B,S,P] positions in index->sorted (device), -1 beyond n_eff
    const double* r_abs;       // [S] absolute radii (host)
};
int launch_3dmfv(const mups_gmm* gmm, const float* patches, const int32_t* n_eff, int64_t B, int S, int P,
                 uint32_t flags, float* out, int* work, const GatherSource* src, cudaStream_t st);

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton_expand(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return morton_expand(x) | (morton_expand(y) << 1) | (morton_expand(z) << 2);
}
__device__ __forceinline__ uint32_t morton_compact(uint32_t v) {  // every third bit -> 10 bits
    v &= 0x09249249u;
    v = (v | (v >> 2)) & 0x030C30C3u;
    v = (v | (v >> 4)) & 0x0300F00Fu;
    v = (v | (v >> 8)) & 0x030000FFu;
    v = (v | (v >> 16)) & 0x3FFu;
    return v;
}
__device__ __forceinline__ int cell_coord(float p, float origin, float inv_cell, int dim) {
    int c = (int)floorf((p - origin) * inv_cell);
    return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

// Philox4x32-10 (Random123 constants); counter (c0..c3), key (k0,k1)
// murmur3's 32-bit finaliser (a bijection): the patch-independent hash of a point index in the shared seeded selection
// (oracle/mups_oracle.py::selection_keys); stored per point by the index build (mups_index::hash_sorted)
__device__ __forceinline__ uint32_t point_hash(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

}  // namespace mups
