// K1 (bounding box) + K2 (uniform-grid spatial hash build, Morton-ordered cells).
// Replaces load_shape -> spatial.cKDTree(pts, 10) (reference utils/pcpnet_dataset.py:13-39) and
// the pts.max(0)/pts.min(0) of utils/pcpnet_dataset.py:281.
//
// Layout in HBM: `sorted` holds the cloud as float4 (x, y, z, original index) ordered by the
// Morton code of the point's grid cell, so that a cell is one contiguous, 16-byte aligned run
// (coalesced LDG.128 in the query kernel) and spatially close cells are close in memory.
// `cell_start[code] .. cell_start[code+1]` is the run of cell `code`.
#include "mups_common.cuh"

namespace mups {

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// K1: grid-stride min/max, warp shuffle + one atomic per warp on order-preserving integer keys.
// Memory-bound: 12 B/point read once.
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ xyz, int64_t n, uint32_t* __restrict__ mm) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = __ldg(xyz + 3 * i + k);
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(mm + k, float_to_ordered(lo[k]));
            atomicMax(mm + 3 + k, float_to_ordered(hi[k]));
        }
    }
}

__global__ void grid_desc_kernel(const uint32_t* __restrict__ mm, GridDesc* __restrict__ g, double cell_frac, int max_dim) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float ext[3];
    double d2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        g->bb_min[k] = ordered_to_float(mm[k]);
        g->bb_max[k] = ordered_to_float(mm[3 + k]);
        g->origin[k] = g->bb_min[k];
        ext[k] = g->bb_max[k] - g->bb_min[k];
        d2 += (double)ext[k] * (double)ext[k];
    }
    double cell = cell_frac * sqrt(d2) * (1.0 + 1e-5);
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    if (!(cell > 0.0)) cell = 1.0;                                       // degenerate cloud (single location)
    if (emax / cell > (double)(max_dim - 1)) cell = (double)emax / (double)(max_dim - 1);
    g->cell = (float)cell;
    g->inv_cell = (float)(1.0 / cell);
    for (int k = 0; k < 3; ++k) {
        int d = (int)floorf(ext[k] * g->inv_cell) + 1;
        g->dims[k] = d < 1 ? 1 : (d > max_dim ? max_dim : d);
    }
}

// K2a: Morton cell code per point + cell population; the atomic's return value is the point's
// rank inside its cell (kept in rank_out), so the scatter needs no second counter pass.
__global__ void __launch_bounds__(256) cell_code_kernel(const float* __restrict__ xyz, int64_t n,
                                                        const GridDesc* __restrict__ g, uint32_t* __restrict__ codes,
                                                        int32_t* __restrict__ rank_out, uint32_t* __restrict__ cell_count) {
    const GridDesc gd = *g;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
        const uint32_t code = morton3(cell_coord(x, gd.origin[0], gd.inv_cell, gd.dims[0]),
                                      cell_coord(y, gd.origin[1], gd.inv_cell, gd.dims[1]),
                                      cell_coord(z, gd.origin[2], gd.inv_cell, gd.dims[2]));
        codes[i] = code;
        rank_out[i] = (int32_t)atomicAdd(cell_count + code, 1u);
    }
}

// Exclusive scan of cell populations: tile sums -> scan of tile sums -> tile-local scan + offset.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*[kScanThreads/32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? smem[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) smem[lane] = w;
    }
    __syncthreads();
    const uint32_t warp_off = warp ? smem[warp - 1] : 0u;
    *total = smem[kScanThreads / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sm[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) s += (base + k < n) ? in[base + k] : 0u;
    uint32_t total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(uint32_t* __restrict__ tile_sums, int n_tiles) {
    __shared__ uint32_t sm[kScanThreads / 32];
    uint32_t carry = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total, sm);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t* __restrict__ data, int64_t n,
                                                                  const uint32_t* __restrict__ tile_offsets) {
    __shared__ uint32_t sm[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, &total, sm) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// K2c: scatter into Morton-cell order as float4 (x, y, z, original index).  12 B read + 16 B write
// per point (+ code/rank/pos 12 B).
__global__ void __launch_bounds__(256) scatter_kernel(const float* __restrict__ xyz, int64_t n,
                                                      const uint32_t* __restrict__ codes,
                                                      const uint32_t* __restrict__ cell_start,
                                                      int32_t* __restrict__ pos_of /* in: rank, out: position */,
                                                      float4* __restrict__ sorted, uint32_t* __restrict__ hash_sorted) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t pos = (int32_t)(cell_start[codes[i]] + (uint32_t)pos_of[i]);
        pos_of[i] = pos;
        sorted[pos] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2),
                                  __int_as_float((int)i));
        hash_sorted[pos] = point_hash((uint32_t)i);      // patch-independent half of the selection key
    }
}

int launch_exclusive_scan(uint32_t* data, int64_t n, uint32_t* tile_sums, cudaStream_t st) {
    const int n_tiles = (int)((n + kScanTile - 1) / kScanTile);
    scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, st>>>(data, n, tile_sums);
    MUPS_CHECK_LAUNCH();
    scan_spine_kernel<<<1, kScanThreads, 0, st>>>(tile_sums, n_tiles);
    MUPS_CHECK_LAUNCH();
    scan_apply_kernel<<<n_tiles, kScanThreads, 0, st>>>(data, n, tile_sums);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

int launch_index_build(mups_index* ix, const float* xyz, cudaStream_t st) {
    const int64_t n = ix->n;
    const int64_t ncode = (int64_t)1 << (3 * ix->bits);
    const int grid = (int)((n + 255) / 256 < 8 * kNumSMs ? (n + 255) / 256 : 8 * kNumSMs);

    // scratch: 6 ordered min/max keys live at the head of cell_start until the scan overwrites them
    uint32_t* mm = reinterpret_cast<uint32_t*>(ix->codes);   // reuse: codes is written after grid_desc has read mm
    static const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    MUPS_CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    bbox_kernel<<<grid, 256, 0, st>>>(xyz, n, mm);
    MUPS_CHECK_LAUNCH();
    grid_desc_kernel<<<1, 32, 0, st>>>(mm, ix->grid, ix->cell_frac, ix->max_dim);
    MUPS_CHECK_LAUNCH();

    MUPS_CUDA_TRY(cudaMemsetAsync(ix->cell_start, 0, sizeof(uint32_t) * (size_t)(ncode + 1), st));
    cell_code_kernel<<<grid, 256, 0, st>>>(xyz, n, ix->grid, ix->codes, ix->pos_of, ix->cell_start);
    MUPS_CHECK_LAUNCH();

    const int64_t n_scan = ncode + 1;
    uint32_t* tile_sums = ix->cell_start + n_scan;            // allocated with n_tiles extra entries
    if (int rc = launch_exclusive_scan(ix->cell_start, n_scan, tile_sums, st)) return rc;

    scatter_kernel<<<grid, 256, 0, st>>>(xyz, n, ix->codes, ix->cell_start, ix->pos_of, ix->sorted, ix->hash_sorted);
    MUPS_CHECK_LAUNCH();
    return MUPS_OK;
}

}  // namespace mups
