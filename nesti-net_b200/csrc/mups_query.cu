// K3 (multi-radius ball query + seeded subsample) and K4 (gather, centre, normalise).
// Replaces PointcloudPatchDataset.__getitem__ (reference utils/pcpnet_dataset.py:286-343):
//   kdtree.query_ball_point (:304) -> uniform-grid scan with cKDTree's float64 predicate
//   rng.choice (:320-321)          -> shared seeded selection: P smallest (philox key, index)
//   gather / centre / /rad (:330-343) -> IEEE fp32 subtract and true division
//
// One CTA per centre point; all S radii are classified in the same scan of the candidate cells.
//   scan     the points of the (2R+1)^3 candidate cells are one flat range list, strided by the
//            whole CTA (coalesced float4 loads, balanced); neighbours are counted per radius and
//            appended to a shared-memory hit list (position + radius mask)
//   select   only for radii with more than P neighbours: keys of the listed hits (a bijection of the
//            point index salted per patch by Philox; dense loop, no divergence) -> 9-bit radix histogram -> threshold bin -> everything below is taken,
//            the threshold group is resolved by rank counting (deeper radix levels if it is large)
//   order    the <= P selected neighbours of each radius are bucket-sorted by point index
//   K4       gather, centre on the query point, divide by float32(r), zero padding
// A ball with more neighbours than the hit list holds (dense scans) re-scans instead of listing.
//
// Dense clouds (fine grids) take ball_query_hier_kernel instead: the Morton order of the cells makes every aligned
// 2^k-cell block one contiguous run of `sorted`, i.e. cell_start is an implicit octree.  A breadth-first descent
// classifies nodes against all radii by their nearest / farthest box distance: outside -> dropped, wholly inside ->
// accepted as a whole (its population comes from cell_start, no per-point predicate: cKDTree's rectangle shortcut),
// straddling a sphere -> descended, and only straddling leaf cells are tested point by point.  The subsample is
// one pass: the population bounds known after the descent give a key threshold that keeps P + 6 sqrt(P) expected
// candidates, and the P smallest keys are resolved among those few; anything this cannot decide is handed to the
// flat kernel through a worklist, so the result is the same function of (seed, centre, scale, index set).
#include <cmath>
#include <cstdlib>

#include "mups_common.cuh"

namespace mups {

constexpr int kQT = 256;            // threads per query CTA
constexpr int kBinBits = 9;
constexpr int kBins = 1 << kBinBits;
constexpr int kBoundaryCap = 128;   // max entries of the threshold radix group resolved by rank counting
constexpr int kHitCap = 4096;       // neighbours (any radius) kept in the shared-memory hit list
constexpr int kRangeCap = 64;       // candidate cells per scan batch
// first-batch candidates above which the first scan also builds the level-1 key histograms (default 3 * kHitCap:
// balls whose hit list will overflow and that re-scan the cells in every pass); mups_set_option("fuse_candidates")

struct QueryArgs {
    const float4* sorted;
    const uint32_t* cell_start;
    const int32_t* pos_of;
    const GridDesc* grid;
    const int64_t* q;
    int64_t n;
    int S, P, Ppad;
    int cap;                        // threshold groups up to this size are resolved by rank counting
    uint32_t fuse;                  // see g_fuse_candidates
    double r2[MUPS_MAX_SCALES];     // r*r in float64 (cKDTree's upper bound for p=2)
    float r2_lo[MUPS_MAX_SCALES];   // fp32 guard band around r2: below -> inside, above hi -> outside
    float r2_hi[MUPS_MAX_SCALES];
    float rf[MUPS_MAX_SCALES];      // float32(r): the divisor of pcpnet_dataset.py:343
    int orig[MUPS_MAX_SCALES];      // the kernels see the radii in ASCENDING order (nested balls); orig[s] = the caller's scale index:
                                    // it salts the selection (Philox counter) and addresses the output rows
    float r_max;
    float r2_hi_max;
    uint32_t k0, k1;                // philox key = seed
    // hierarchical kernel
    const uint32_t* hash_sorted;
    int bits;                       // Morton bits per axis of the leaf level
    uint32_t ccap;                  // candidate positions kept per radius
    uint32_t want;                  // P + 6 sqrt(P) + 8: expected candidates under the key threshold
    const int32_t* order;           // CTA -> batch row (Morton order of the centres), or nullptr
    int32_t* worklist;              // hier: rows it could not decide; flat: rows to process when work_count != nullptr
    int32_t* work_count;
    int32_t* nbr_pos;               // [B,S,P] position in `sorted` of each selected neighbour, -1 beyond n_eff (may be NULL)
    int32_t* nbr_idx;
    int32_t* nbr_total;
    float* patches;
    int32_t* n_eff;
};

struct QueryCtx {
    float cx, cy, cz;
    double cxd, cyd, czd;
    int x0, x1, y0, y1, z0, z1;
    float ox, oy, oz, cell;
    uint32_t center;
};

struct ScanTables {
    uint32_t start[kRangeCap];
    uint32_t prefix[kRangeCap + 1];
    uint32_t first_total;           // candidates of the first batch of the last scan (all 27 cells when R = 1)
    uint32_t n_ranges, total;       // non-empty cells of the batch (start / prefix hold only those), candidates of the batch
};

// cKDTree leaf predicate: s = 0; s += d*d for x, y, z in float64 without FMA contraction; s <= r*r
// (scipy/spatial/ckdtree/src/distance_base.h sqeuclidean_distance_double, m = 3).
__device__ __forceinline__ bool inside_exact(const QueryCtx& c, const float4& p, double r2) {
    const double dx = __dsub_rn((double)p.x, c.cxd), dy = __dsub_rn((double)p.y, c.cyd), dz = __dsub_rn((double)p.z, c.czd);
    double s = __dmul_rn(dx, dx);
    s = __dadd_rn(s, __dmul_rn(dy, dy));
    s = __dadd_rn(s, __dmul_rn(dz, dz));
    return s <= r2;
}

// Visits every neighbour of the centre (any radius): f(position in sorted, original index, bitmask of radii).
// Block-wide: contains __syncthreads(); every thread of the CTA must call it.
template <int NS, class F>
__device__ __forceinline__ void for_each_hit(const QueryArgs& a, const QueryCtx& c, ScanTables& st, F&& f) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int nx = c.x1 - c.x0 + 1, ny = c.y1 - c.y0 + 1, nz = c.z1 - c.z0 + 1;
    const int ncell = nx * ny * nz;
    const float slack = 1e-3f * c.cell;
    for (int cbase = 0; cbase < ncell; cbase += kRangeCap) {
        if (tid < 32) {   // warp 0 builds the range table of this batch: two cells per lane
            uint32_t cnt[2], beg[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ci = cbase + 2 * lane + h;
                cnt[h] = 0u; beg[h] = 0u;
                if (ci < ncell) {
                    const int ix = c.x0 + ci % nx, iy = c.y0 + (ci / nx) % ny, iz = c.z0 + ci / (nx * ny);
                    // distance from the centre to the cell's box (shrunk by the rounding slack of the cell assignment)
                    const float lx = c.ox + ix * c.cell, ly = c.oy + iy * c.cell, lz = c.oz + iz * c.cell;
                    const float gx = fmaxf(0.f, fmaxf(lx - c.cx, c.cx - (lx + c.cell)) - slack);
                    const float gy = fmaxf(0.f, fmaxf(ly - c.cy, c.cy - (ly + c.cell)) - slack);
                    const float gz = fmaxf(0.f, fmaxf(lz - c.cz, c.cz - (lz + c.cell)) - slack);
                    if (gx * gx + gy * gy + gz * gz <= a.r2_hi_max) {
                        const uint32_t code = morton3((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
                        beg[h] = __ldg(a.cell_start + code);
                        cnt[h] = __ldg(a.cell_start + code + 1) - beg[h];
                    }
                }
            }
            uint32_t inc = cnt[0] + cnt[1];
            uint32_t slots = (cnt[0] ? 1u : 0u) + (cnt[1] ? 1u : 0u);       // only non-empty cells get a table entry: the
#pragma unroll                                                              // scan's walk over the table is a tenth of the
            for (int o = 1; o < 32; o <<= 1) {                              // kernel's instructions on surface clouds, where
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);     // two thirds of the 27 cells are empty
                const uint32_t u = __shfl_up_sync(0xffffffffu, slots, o);
                if (lane >= o) { inc += t; slots += u; }
            }
            const uint32_t exc = inc - cnt[0] - cnt[1];
            uint32_t e = slots - (cnt[0] ? 1u : 0u) - (cnt[1] ? 1u : 0u);
            if (cnt[0]) { st.start[e] = beg[0]; st.prefix[e] = exc; ++e; }
            if (cnt[1]) { st.start[e] = beg[1]; st.prefix[e] = exc + cnt[0]; }
            if (lane == 31) {
                st.prefix[slots] = inc;          // end sentinel of the compacted table
                st.n_ranges = slots;
                st.total = inc;
                if (cbase == 0) st.first_total = inc;
            }
        }
        __syncthreads();
        const uint32_t total = st.total;
        int k = 0;
        for (uint32_t fi = tid; fi < total; fi += kQT) {
            while (fi >= st.prefix[k + 1]) ++k;
            const uint32_t i = st.start[k] + (fi - st.prefix[k]);
            const float4 p = __ldg(a.sorted + i);
            const float dx = p.x - c.cx, dy = p.y - c.cy, dz = p.z - c.cz;
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d2 > a.r2_hi_max) continue;
            uint32_t in = 0, band = 0;      // inside for sure / inside the fp32 guard band of a radius
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                in |= (d2 < a.r2_lo[s] ? 1u : 0u) << s;
                band |= (d2 >= a.r2_lo[s] && d2 <= a.r2_hi[s] ? 1u : 0u) << s;
            }
            if (band) {                     // rare: decide with cKDTree's float64 predicate
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    if ((band >> s) & 1u) in |= (inside_exact(c, p, a.r2[s]) ? 1u : 0u) << s;
            }
            if (in) f(i, (uint32_t)__float_as_int(p.w), in);
        }
        __syncthreads();
    }
}

// key of neighbour `idx` for radius s: (fmix32(idx) ^ a_s) * b_s, a bijection of idx keyed by the salt
// (a_s, b_s | 1) = Philox4x32-10(counter = (centre, s, 0, 0), key = seed) computed once per CTA and radius; fmix32(idx) is
// patch-independent: once per neighbour here, read from the index (hash_sorted) in the hierarchical kernel
struct Salts {
    uint32_t a[MUPS_MAX_SCALES], b[MUPS_MAX_SCALES];
};
__device__ __forceinline__ uint32_t fmix32(uint32_t h) { return point_hash(h); }
template <int NS>
__device__ __forceinline__ void selection_keys(const Salts& salt, uint32_t idx, uint32_t need_mask,
                                               uint32_t key[MUPS_MAX_SCALES]) {
    const uint32_t h = fmix32(idx);
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (need_mask & (1u << s)) key[s] = (h ^ salt.a[s]) * salt.b[s];
}

// One warp finds, in hist[0..nbins), the bin T holding the `need`-th smallest element.
// Returns T, the population below T and the population of T (valid on all lanes).
__device__ __forceinline__ void warp_find_threshold(const uint32_t* hist, int nbins, uint32_t need, int lane,
                                                    uint32_t* T, uint32_t* below, uint32_t* group) {
    const int per = (nbins + 31) / 32;
    uint32_t mine = 0;
    for (int k = 0; k < per; ++k) {
        const int b = lane * per + k;
        mine += b < nbins ? hist[b] : 0u;
    }
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t exc = inc - mine;
    const bool owner = exc < need && need <= inc;
    uint32_t t_bin = 0, t_below = 0, t_group = 0;
    if (owner) {
        uint32_t run = exc;
        for (int k = 0; k < per; ++k) {
            const int b = lane * per + k;
            const uint32_t h = b < nbins ? hist[b] : 0u;
            if (run < need && need <= run + h) { t_bin = b; t_below = run; t_group = h; break; }
            run += h;
        }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, owner);
    const int src = ballot ? (__ffs(ballot) - 1) : 0;
    *T = __shfl_sync(0xffffffffu, t_bin, src);
    *below = __shfl_sync(0xffffffffu, t_below, src);
    *group = __shfl_sync(0xffffffffu, t_group, src);
}

// Exclusive scan of kBins counters by the whole CTA (kBins / NT per thread), in place.
template <int NT>
__device__ __forceinline__ void block_scan_bins(uint32_t* bins, uint32_t* warp_sums /*[NT/32]*/) {
    constexpr int PER = kBins / NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = bins[tid * PER + k]; s += v[k]; }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += warp_sums[w];
    uint32_t run = off + inc - s;
#pragma unroll
    for (int k = 0; k < PER; ++k) { bins[tid * PER + k] = run; run += v[k]; }
    __syncthreads();
}

// Shared-memory state both query kernels hand to the common tail.
struct PatchLists {
    uint32_t* hist;                 // [S][kBins]
    unsigned long long* sel;        // [S][Ppad]  (idx << 32 | pos)
    unsigned long long* tmp;        // [S][Ppad]  second buffer of the bucket sort
    unsigned long long* bndl;       // [S][bcap]  threshold group: (key << 32 | idx)
    uint32_t* bndp;                 // [S][bcap]  its positions in `sorted`
    uint32_t bcap;
    uint32_t* s_cnt;                // [S] neighbours per radius
    uint32_t* s_nsel;               // [S] entries already in sel
    uint32_t* s_nb;                 // [S] entries of the threshold group
    uint32_t* s_need;               // [S] how many of the threshold group are kept
    uint32_t* s_min;                // [S] scratch (index range of the selection)
    uint32_t* s_max;
    uint32_t* s_warp_sums;          // [NT/32]
};

// An invalid centre index yields an empty patch with total = -1 (no error crosses the ABI for a bad row).
template <int NT>
__device__ __forceinline__ void write_invalid_row(const QueryArgs& a, int64_t b) {
    const int S = a.S, P = a.P, tid = threadIdx.x;
    for (int i = tid; i < S * P; i += NT) {
        if (a.patches) { float* o = a.patches + (b * S * P + i) * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
        if (a.nbr_idx) a.nbr_idx[b * S * P + i] = -1;
        if (a.nbr_pos) a.nbr_pos[b * S * P + i] = -1;
    }
    if (tid < S) { a.n_eff[b * S + tid] = 0; if (a.nbr_total) a.nbr_total[b * S + tid] = -1; }
}

// Tail of both kernels.  On entry (after a barrier): sel holds the neighbours already known to be selected, the
// threshold group of every over-full radius is in bndl / bndp with s_need of them still to be taken.  Resolves the
// group by rank counting on (key, index), orders every radius' selection by point index (bucket sort over the index
// range: monotone buckets, so bucket order + order inside a bucket = index order) and runs K4.
template <int NS, int NT>
__device__ __forceinline__ void finish_patch(const QueryArgs& a, const QueryCtx& c, int64_t b, const PatchLists& L) {
    constexpr int S = NS;
    const int P = a.P, Ppad = a.Ppad;
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t* hist = L.hist;
    unsigned long long* sel = L.sel;
    unsigned long long* tmp = L.tmp;

    // ---- threshold group: keep the `need` smallest (key, index) pairs ------------------------------------
    for (int s = 0; s < S; ++s) {
        const uint32_t m = min(L.s_nb[s], L.bcap);
        const uint32_t need = L.s_need[s];
        for (uint32_t i = tid; i < m; i += NT) {
            const unsigned long long mine = L.bndl[s * L.bcap + i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) rank += L.bndl[s * L.bcap + j] < mine ? 1u : 0u;
            if (rank < need) {
                const uint32_t slot = atomicAdd(L.s_nsel + s, 1u);
                if (slot < (uint32_t)Ppad)
                    sel[(size_t)s * Ppad + slot] = ((mine & 0xFFFFFFFFull) << 32) | L.bndp[s * L.bcap + i];
            }
        }
    }
    __syncthreads();

    // ---- order every radius' selection by point index ----------------------------------------------------
    for (int i = tid; i < S * kBins; i += NT) hist[i] = 0u;
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(L.s_cnt[s], (uint32_t)P);
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        for (uint32_t t = tid; t < ne; t += NT) {
            const uint32_t id = (uint32_t)(sel[(size_t)s * Ppad + t] >> 32);
            lo = min(lo, id); hi = max(hi, id);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && lo <= hi) { atomicMin(L.s_min + s, lo); atomicMax(L.s_max + s, hi); }
    }
    __syncthreads();
    auto bucket_of = [&](int s, uint32_t id) {
        const float scale = (float)kBins / ((float)(L.s_max[s] - L.s_min[s]) + 1.0f);
        return min((uint32_t)(kBins - 1), (uint32_t)((float)(id - L.s_min[s]) * scale));
    };
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(L.s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += NT)
            atomicAdd(hist + s * kBins + bucket_of(s, (uint32_t)(sel[(size_t)s * Ppad + t] >> 32)), 1u);
    }
    __syncthreads();
    for (int s = 0; s < S; ++s) block_scan_bins<NT>(hist + s * kBins, L.s_warp_sums);     // counts -> bucket starts
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(L.s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += NT) {
            const unsigned long long e = sel[(size_t)s * Ppad + t];
            const uint32_t slot = atomicAdd(hist + s * kBins + bucket_of(s, (uint32_t)(e >> 32)), 1u);   // starts -> ends
            tmp[(size_t)s * Ppad + slot] = e;
        }
    }
    __syncthreads();
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(L.s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += NT) {
            const unsigned long long e = tmp[(size_t)s * Ppad + t];
            const uint32_t bk = bucket_of(s, (uint32_t)(e >> 32));
            const uint32_t beg = bk ? hist[s * kBins + bk - 1] : 0u, end = hist[s * kBins + bk];
            uint32_t rank = 0;
            for (uint32_t j = beg; j < end; ++j) rank += tmp[(size_t)s * Ppad + j] < e ? 1u : 0u;
            sel[(size_t)s * Ppad + beg + rank] = e;
        }
    }
    __syncthreads();

    // ---- K4: gather, centre on the query point, divide by float32(r)  (pcpnet_dataset.py:330-343) ----------
    for (int s = 0; s < S; ++s) {
        const uint32_t total = L.s_cnt[s];
        const uint32_t ne = min(total, (uint32_t)P);
        const unsigned long long* v = sel + (size_t)s * Ppad;
        const float rf = a.rf[s];
        const int os = a.orig[s];                           // output row of this radius
        for (uint32_t t = tid; t < (uint32_t)P; t += NT) {
            float ox = 0.f, oy = 0.f, oz = 0.f;
            int32_t id = -1, ps = -1;
            if (t < ne) {
                const unsigned long long e = v[t];
                id = (int32_t)(e >> 32);
                ps = (int32_t)(uint32_t)(e & 0xFFFFFFFFull);
                if (a.patches) {
                    const float4 p = __ldg(a.sorted + (uint32_t)ps);
                    ox = __fdiv_rn(__fsub_rn(p.x, c.cx), rf);
                    oy = __fdiv_rn(__fsub_rn(p.y, c.cy), rf);
                    oz = __fdiv_rn(__fsub_rn(p.z, c.cz), rf);
                }
            }
            if (a.patches) {
                float* o = a.patches + ((b * S + os) * (int64_t)P + t) * 3;
                o[0] = ox; o[1] = oy; o[2] = oz;
            }
            if (a.nbr_idx) a.nbr_idx[(b * S + os) * (int64_t)P + t] = id;
            if (a.nbr_pos) a.nbr_pos[(b * S + os) * (int64_t)P + t] = ps;
        }
        if (tid == 0) {
            a.n_eff[b * S + os] = (int32_t)ne;
            if (a.nbr_total) a.nbr_total[b * S + os] = (int32_t)total;
        }
    }
}

// Centre point, its cell and the candidate cell box of the largest radius.
__device__ __forceinline__ void load_centre(const QueryArgs& a, int64_t q, QueryCtx& c, int* R_out) {
    const GridDesc g = *a.grid;
    const float4 pc = __ldg(a.sorted + __ldg(a.pos_of + q));
    c.cx = pc.x; c.cy = pc.y; c.cz = pc.z;
    c.cxd = (double)pc.x; c.cyd = (double)pc.y; c.czd = (double)pc.z;
    c.ox = g.origin[0]; c.oy = g.origin[1]; c.oz = g.origin[2]; c.cell = g.cell;
    c.center = (uint32_t)q;
    // |floor(u) - floor(v)| <= floor(|u - v|) + 1; 1e-4 covers the fp32 rounding of the cell assignment
    const int R = (int)floorf(a.r_max * g.inv_cell + 1e-4f) + 1;
    const int ix = cell_coord(pc.x, g.origin[0], g.inv_cell, g.dims[0]);
    const int iy = cell_coord(pc.y, g.origin[1], g.inv_cell, g.dims[1]);
    const int iz = cell_coord(pc.z, g.origin[2], g.inv_cell, g.dims[2]);
    c.x0 = max(ix - R, 0); c.x1 = min(ix + R, g.dims[0] - 1);
    c.y0 = max(iy - R, 0); c.y1 = min(iy + R, g.dims[1] - 1);
    c.z0 = max(iz - R, 0); c.z1 = min(iz + R, g.dims[2] - 1);
    *R_out = R;
}

// =====================================================================================================
// flat kernel: every cell of the candidate box is scanned point by point (PCPNet-size clouds, one cell per radius)
// =====================================================================================================

template <int NS>
__device__ __forceinline__ void ball_query_flat_one(const QueryArgs& a, const int64_t b, unsigned char* smem_raw) {
    constexpr int S = NS;
    const int P = a.P, Ppad = a.Ppad;
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);                                  // [S][kBins]
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(hist + S * kBins);       // [S][Ppad]  (idx << 32 | pos)
    unsigned long long* bnd = sel + (size_t)S * Ppad;                                        // [S][kBoundaryCap] (key << 32 | idx)
    uint32_t* bnd_pos = reinterpret_cast<uint32_t*>(bnd + (size_t)S * kBoundaryCap);         // [S][kBoundaryCap] position in sorted
    // the hit list (scan .. select) and the sort's second buffer (order) share one region
    unsigned char* shared_region = reinterpret_cast<unsigned char*>(bnd_pos + (size_t)S * kBoundaryCap);
    uint32_t* hit_pos = reinterpret_cast<uint32_t*>(shared_region);                           // [kHitCap]
    unsigned char* hit_mask = reinterpret_cast<unsigned char*>(hit_pos + kHitCap);            // [kHitCap]
    unsigned long long* tmp = reinterpret_cast<unsigned long long*>(shared_region);           // [S][Ppad]
    __shared__ ScanTables st;
    __shared__ uint32_t s_cnt[MUPS_MAX_SCALES], s_nsel[MUPS_MAX_SCALES], s_nb[MUPS_MAX_SCALES];
    __shared__ uint32_t s_prefix[MUPS_MAX_SCALES], s_bits[MUPS_MAX_SCALES], s_need[MUPS_MAX_SCALES];
    __shared__ uint32_t s_min[MUPS_MAX_SCALES], s_max[MUPS_MAX_SCALES];
    __shared__ uint32_t s_unresolved, s_nhits, s_fused;
    __shared__ uint32_t s_warp_sums[kQT / 32];
    __shared__ Salts salt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = a.q[b];

    for (int i = tid; i < S * kBins; i += kQT) hist[i] = 0u;
    if (tid < MUPS_MAX_SCALES) {
        s_cnt[tid] = 0u; s_nsel[tid] = 0u; s_nb[tid] = 0u; s_prefix[tid] = 0u; s_bits[tid] = 0u; s_need[tid] = 0u;
        s_min[tid] = 0xFFFFFFFFu; s_max[tid] = 0u;
    }
    if (tid == 0) { s_unresolved = 0u; s_nhits = 0u; }

    if (q < 0 || q >= a.n) {   // invalid centre: empty patch, total = -1
        write_invalid_row<kQT>(a, b);
        return;
    }

    QueryCtx c;
    int R_unused;
    load_centre(a, q, c, &R_unused);
    if (tid < S) {   // per-patch randomness: one Philox call per radius
        const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)a.orig[tid], 0u, 0u, a.k0, a.k1);
        salt.a[tid] = w.x;
        salt.b[tid] = w.y | 1u;
    }
    __syncthreads();

    // ---- scan: neighbour count per radius + hit list ---------------------------------------------
    {
        uint32_t cnt[MUPS_MAX_SCALES];
#pragma unroll
        for (int s = 0; s < NS; ++s) cnt[s] = 0u;
        for_each_hit<NS>(a, c, st, [&](uint32_t pos, uint32_t idx, uint32_t in) {
#pragma unroll
            for (int s = 0; s < NS; ++s) cnt[s] += (in >> s) & 1u;
            // one shared-memory atomic per warp instead of one per hit (half of the candidates are hits; the counter is a
            // single address)
            const unsigned peers = __activemask();
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&s_nhits, (uint32_t)__popc(peers));
            const uint32_t slot = __shfl_sync(peers, base, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (slot < (uint32_t)kHitCap) { hit_pos[slot] = pos; hit_mask[slot] = (unsigned char)in; }
            if (st.first_total > a.fuse) {
                // dense neighbourhood: the hit list will overflow and every later pass re-scans the cells, so
                // the level-1 key histograms of all radii are built here and one full scan is saved
                uint32_t key[MUPS_MAX_SCALES];
                selection_keys<NS>(salt, idx, in, key);
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    if (in & (1u << s)) atomicAdd(hist + s * kBins + (key[s] >> (32 - kBinBits)), 1u);
            }
        });
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            {
                uint32_t v = cnt[s];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v) atomicAdd(s_cnt + s, v);
            }
        }
        // latched here: later scans rewrite st.first_total (with the same value) without a barrier in between
        if (tid == 0) s_fused = st.first_total > a.fuse ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t nhits = s_nhits;
    const bool listed = nhits <= (uint32_t)kHitCap;     // else: every pass re-scans the cells
    // threshold-group storage: the small dedicated lists, or -- when the hit list is unused (re-scan mode: dense
    // balls, large groups) -- the hit-list region, which holds ~3x more entries and saves a refinement scan
    uint32_t bcap = (uint32_t)kBoundaryCap;
    unsigned long long* bndl = bnd;
    uint32_t* bndp = bnd_pos;
    if (!listed) {
        bcap = ((uint32_t)(kHitCap * 5) / (12u * (uint32_t)S)) & ~7u;
        bndl = reinterpret_cast<unsigned long long*>(shared_region);
        bndp = reinterpret_cast<uint32_t*>(bndl + (size_t)S * bcap);
    }
    const uint32_t gcap = min((uint32_t)a.cap, bcap);   // largest group resolved without another radix level
    uint32_t over = 0;                                   // radii with more than P neighbours need keys
    for (int s = 0; s < S; ++s) over |= (s_cnt[s] > (uint32_t)P) ? (1u << s) : 0u;

    // visits the neighbours that belong to at least one radius of `want`
    auto visit = [&](uint32_t want, auto&& f) {
        if (listed) {
            for (uint32_t h = tid; h < nhits; h += kQT) {
                const uint32_t in = hit_mask[h];
                if (in & want) {
                    const uint32_t pos = hit_pos[h];
                    f(pos, (uint32_t)__float_as_int(__ldg(&a.sorted[pos].w)), in);
                }
            }
            __syncthreads();
        } else {
            for_each_hit<NS>(a, c, st, [&](uint32_t pos, uint32_t idx, uint32_t in) { if (in & want) f(pos, idx, in); });
        }
    };

    const bool fused_hist = s_fused != 0u;               // uniform
    if (over) {
        // ---- first-level key histogram of the over-full radii ------------------------------------------
        if (!fused_hist) visit(over, [&](uint32_t, uint32_t idx, uint32_t in) {
            in &= over;
            uint32_t key[MUPS_MAX_SCALES];
            selection_keys<NS>(salt, idx, in, key);
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (in & (1u << s)) atomicAdd(hist + s * kBins + (key[s] >> (32 - kBinBits)), 1u);
        });
        // ---- radix threshold (warp s handles radius s) ----------------------------------------------------
        if (warp < S && (over & (1u << warp))) {
            uint32_t T, below, group;
            warp_find_threshold(hist + warp * kBins, kBins, (uint32_t)P, lane, &T, &below, &group);
            if (lane == 0) {
                s_prefix[warp] = T; s_bits[warp] = kBinBits; s_need[warp] = (uint32_t)P - below;
                if (group > gcap) atomicOr(&s_unresolved, 1u << warp);
            }
        }
        __syncthreads();
        // ---- refinement levels (only when a threshold group exceeds the cap: > ~60k neighbours) ----------
        while (s_unresolved) {
            const uint32_t unresolved = s_unresolved;
            __syncthreads();
            for (int i = tid; i < S * kBins; i += kQT)
                if (unresolved & (1u << (i / kBins))) hist[i] = 0u;
            if (tid == 0) s_unresolved = 0u;
            __syncthreads();
            visit(unresolved, [&](uint32_t, uint32_t idx, uint32_t in) {
                in &= unresolved;
                uint32_t key[MUPS_MAX_SCALES];
                selection_keys<NS>(salt, idx, in, key);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (in & (1u << s)) {
                        const uint32_t bits = s_bits[s];
                        const uint32_t nb = min((uint32_t)kBinBits, 32u - bits);
                        if ((key[s] >> (32u - bits)) == s_prefix[s])
                            atomicAdd(hist + s * kBins + ((key[s] >> (32u - bits - nb)) & ((1u << nb) - 1u)), 1u);
                    }
                }
            });
            if (warp < S && (unresolved & (1u << warp))) {
                const uint32_t bits = s_bits[warp];
                const uint32_t nb = min((uint32_t)kBinBits, 32u - bits);
                uint32_t T, below, group;
                warp_find_threshold(hist + warp * kBins, 1 << nb, s_need[warp], lane, &T, &below, &group);
                __syncwarp();                 // every lane has read s_bits / s_need before lane 0 updates them
                if (lane == 0) {
                    s_prefix[warp] = (s_prefix[warp] << nb) | T; s_bits[warp] = bits + nb; s_need[warp] -= below;
                    // with all 32 key bits fixed the group is a set of exact key ties; more than the cap of
                    // them cannot be told apart here (the keys are a bijection of the index, so this cannot happen)
                    if (group > gcap && bits + nb < 32u) atomicOr(&s_unresolved, 1u << warp);
                }
            }
            __syncthreads();
        }
    }

    // ---- collect the selection ----------------------------------------------------------------------------
    visit(0xFFu, [&](uint32_t pos, uint32_t idx, uint32_t in) {
        uint32_t key[MUPS_MAX_SCALES];
        selection_keys<NS>(salt, idx, in & over, key);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (!(in & (1u << s))) continue;
            bool take = true;
            if (over & (1u << s)) {
                const uint32_t hp = key[s] >> (32u - s_bits[s]);
                take = hp < s_prefix[s];
                if (hp == s_prefix[s]) {
                    const uint32_t slot = atomicAdd(s_nb + s, 1u);
                    if (slot < bcap) {
                        bndl[s * bcap + slot] = ((unsigned long long)key[s] << 32) | idx;
                        bndp[s * bcap + slot] = pos;
                    }
                }
            }
            if (take) {
                const uint32_t slot = atomicAdd(s_nsel + s, 1u);
                if (slot < (uint32_t)Ppad) sel[(size_t)s * Ppad + slot] = ((unsigned long long)idx << 32) | pos;
            }
        }
    });
    if (!listed) __syncthreads();

    PatchLists L;
    L.hist = hist; L.sel = sel; L.tmp = tmp; L.bndl = bndl; L.bndp = bndp; L.bcap = bcap;
    L.s_cnt = s_cnt; L.s_nsel = s_nsel; L.s_nb = s_nb; L.s_need = s_need; L.s_min = s_min; L.s_max = s_max;
    L.s_warp_sums = s_warp_sums;
    if (!listed) {
        // the threshold group lives in the region the sort's second buffer uses: resolve it first (finish_patch
        // does exactly that before it touches tmp)
    }
    finish_patch<NS, kQT>(a, c, b, L);
}

template <int NS>
__global__ void __launch_bounds__(kQT) ball_query_kernel(const QueryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ball_query_flat_one<NS>(a, a.order ? (int64_t)a.order[blockIdx.x] : (int64_t)blockIdx.x, smem_raw);
}

// worklist mode: the rows the hierarchical kernel could not decide, a few resident CTAs looping over them
template <int NS>
__global__ void __launch_bounds__(kQT) ball_query_worklist_kernel(const QueryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_items = *a.work_count;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        ball_query_flat_one<NS>(a, (int64_t)a.worklist[it], smem_raw);
        __syncthreads();
    }
}

// =====================================================================================================
// hierarchical kernel: octree descent over the Morton-ordered cells (dense clouds, fine grids)
// =====================================================================================================

constexpr int kHT = 256;            // threads per CTA
constexpr int kSegCap = 1280;       // segments (contiguous runs of `sorted` with one classification) per query
constexpr int kQueueCap = 1024;     // straddling nodes per octree level
constexpr int kSegSplit = 2048;     // longer runs are cut so that the warps share them

// nearest / farthest squared distance from the centre to the box of cells [X0, X1) x [Y0, Y1) x [Z0, Z1), widened by
// the rounding slack of the cell assignment
__device__ __forceinline__ void box_dist2(const QueryCtx& c, int X0, int X1, int Y0, int Y1, int Z0, int Z1, float* near2,
                                          float* far2) {
    const float slack = 1e-3f * c.cell;
    const float lx = c.ox + X0 * c.cell - slack, hx = c.ox + X1 * c.cell + slack;
    const float ly = c.oy + Y0 * c.cell - slack, hy = c.oy + Y1 * c.cell + slack;
    const float lz = c.oz + Z0 * c.cell - slack, hz = c.oz + Z1 * c.cell + slack;
    const float nx = fmaxf(0.f, fmaxf(lx - c.cx, c.cx - hx)), fx = fmaxf(c.cx - lx, hx - c.cx);
    const float ny = fmaxf(0.f, fmaxf(ly - c.cy, c.cy - hy)), fy = fmaxf(c.cy - ly, hy - c.cy);
    const float nz = fmaxf(0.f, fmaxf(lz - c.cz, c.cz - hz)), fz = fmaxf(c.cz - lz, hz - c.cz);
    *near2 = nx * nx + ny * ny + nz * nz;
    *far2 = fx * fx + fy * fy + fz * fz;
}

struct NodeClass {
    uint32_t start, pop;
    uint32_t inside, straddle;      // bitmasks over the radii
};

// node `code` of octree level `level` (level = bits per axis; the leaf level is a.bits)
template <int NS>
__device__ __forceinline__ NodeClass classify_node(const QueryArgs& a, const QueryCtx& c, uint32_t code, int level) {
    NodeClass r;
    const int sh = 3 * (a.bits - level);
    const uint32_t lo = code << sh;
    r.start = __ldg(a.cell_start + lo);
    r.pop = __ldg(a.cell_start + lo + (1u << sh)) - r.start;
    r.inside = 0u; r.straddle = 0u;
    if (r.pop == 0u) return r;
    const int e = a.bits - level;
    const int X = (int)morton_compact(code), Y = (int)morton_compact(code >> 1), Z = (int)morton_compact(code >> 2);
    float near2, far2;
    box_dist2(c, X << e, (X + 1) << e, Y << e, (Y + 1) << e, Z << e, (Z + 1) << e, &near2, &far2);
    // conservative on both sides (2e-5 relative covers the fp32 evaluation of the box distances): a node counted as
    // inside holds only points the float64 predicate accepts, a dropped node only points it rejects
    const float n2 = near2 * (1.0f - 2e-5f), f2 = far2 * (1.0f + 2e-5f);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (f2 <= a.r2_lo[s]) r.inside |= 1u << s;
        else if (n2 <= a.r2_hi[s]) r.straddle |= 1u << s;
    }
    return r;
}

template <int NS>
__global__ void __launch_bounds__(kHT, 3) ball_query_hier_kernel(const QueryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int S = NS;
    const int P = a.P, Ppad = a.Ppad;
    const uint32_t ccap = a.ccap;
    // layout (launch_ball_query sizes it the same way)
    uint32_t* seg = reinterpret_cast<uint32_t*>(smem_raw);                                      // [kSegCap][3]
    constexpr int kHQ = (S * kBins > 2 * kQueueCap) ? S * kBins : 2 * kQueueCap;
    uint32_t* hist = seg + kSegCap * 3;                                                         // [S][kBins]; the two level queues before
    uint32_t* queue0 = hist;
    uint32_t* queue1 = hist + kQueueCap;
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(hist + kHQ);                // [S][Ppad]
    unsigned long long* bnd = sel + (size_t)S * Ppad;                                           // [S][kBoundaryCap]
    uint32_t* bnd_pos = reinterpret_cast<uint32_t*>(bnd + (size_t)S * kBoundaryCap);            // [S][kBoundaryCap]
    unsigned char* region = reinterpret_cast<unsigned char*>(bnd_pos + (size_t)S * kBoundaryCap);
    uint32_t* cand = reinterpret_cast<uint32_t*>(region);                                       // [S][ccap] positions
    unsigned long long* tmp = reinterpret_cast<unsigned long long*>(region);                    // [S][Ppad] (after the selection)
    __shared__ uint32_t s_cnt[MUPS_MAX_SCALES], s_nsel[MUPS_MAX_SCALES], s_nb[MUPS_MAX_SCALES], s_need[MUPS_MAX_SCALES];
    __shared__ uint32_t s_min[MUPS_MAX_SCALES], s_max[MUPS_MAX_SCALES];
    __shared__ uint32_t s_nlo[MUPS_MAX_SCALES], s_nbd[MUPS_MAX_SCALES];        // population wholly inside / in straddling segments
    __shared__ uint32_t s_ncand[MUPS_MAX_SCALES], s_T[MUPS_MAX_SCALES], s_scale[MUPS_MAX_SCALES], s_tbin[MUPS_MAX_SCALES];
    __shared__ uint32_t s_nseg, s_qn[2], s_emit, s_expand, s_ticket, s_fail, s_collect, s_late;
    __shared__ uint32_t s_warp_sums[kHT / 32];
    __shared__ Salts salt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = a.order ? (int64_t)a.order[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t q = a.q[b];

    if (tid < MUPS_MAX_SCALES) {
        s_cnt[tid] = 0u; s_nsel[tid] = 0u; s_nb[tid] = 0u; s_need[tid] = 0u; s_min[tid] = 0xFFFFFFFFu; s_max[tid] = 0u;
        s_nlo[tid] = 0u; s_nbd[tid] = 0u; s_ncand[tid] = 0u; s_T[tid] = 0u; s_scale[tid] = 0u; s_tbin[tid] = 0u;
    }
    if (tid == 0) { s_nseg = 0u; s_qn[0] = 0u; s_qn[1] = 0u; s_emit = 0u; s_expand = 0u; s_ticket = 0u; s_fail = 0u; s_collect = 0u; s_late = 0u; }

    if (q < 0 || q >= a.n) {
        write_invalid_row<kHT>(a, b);
        return;
    }
    QueryCtx c;
    int R;
    load_centre(a, q, c, &R);
    if (tid < S) {
        const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)a.orig[tid], 0u, 0u, a.k0, a.k1);
        salt.a[tid] = w.x;
        salt.b[tid] = w.y | 1u;
    }
    __syncthreads();

    // ---- descent ----------------------------------------------------------------------------------------------
    // start level: nodes of edge 2^e >= R + 1 cells, so the candidate box spans at most 3 of them per axis
    int e0 = 0;
    while ((1 << e0) < R + 1 && e0 < a.bits) ++e0;
    const int X0 = c.x0 >> e0, Y0 = c.y0 >> e0, Z0 = c.z0 >> e0;
    const int NX = (c.x1 >> e0) - X0 + 1, NY = (c.y1 >> e0) - Y0 + 1, NZ = (c.z1 >> e0) - Z0 + 1;

    auto emit_segments = [&](const NodeClass& nc) {      // cut into pieces of <= kSegSplit points
        const uint32_t pieces = (nc.pop + kSegSplit - 1) / kSegSplit;
        const uint32_t slot = atomicAdd(&s_nseg, pieces);
        for (uint32_t k = 0; k < pieces; ++k) {
            if (slot + k < (uint32_t)kSegCap) {
                seg[3 * (slot + k)] = nc.start + k * kSegSplit;
                seg[3 * (slot + k) + 1] = min((uint32_t)kSegSplit, nc.pop - k * kSegSplit);
                seg[3 * (slot + k) + 2] = nc.inside | (nc.straddle << 8);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (nc.inside & (1u << s)) atomicAdd(s_nlo + s, nc.pop);
            if (nc.straddle & (1u << s)) atomicAdd(s_nbd + s, nc.pop);
        }
    };

    int level = a.bits - e0;
    int cur = 0;                                         // queue holding the straddling nodes of `level - 1`
    bool first = true;
    while (true) {
        // candidates of this level: the start nodes, or the 8 children of every queued node
        const uint32_t* qin = cur ? queue1 : queue0;
        uint32_t* qout = cur ? queue0 : queue1;
        const uint32_t n_cand = first ? (uint32_t)(NX * NY * NZ) : 8u * s_qn[cur];
        auto candidate = [&](uint32_t t) -> uint32_t {
            if (first) {
                const int dx = (int)(t % (uint32_t)NX), dy = (int)((t / (uint32_t)NX) % (uint32_t)NY), dz = (int)(t / (uint32_t)(NX * NY));
                return morton3((uint32_t)(X0 + dx), (uint32_t)(Y0 + dy), (uint32_t)(Z0 + dz));
            }
            return (qin[t >> 3] << 3) | (t & 7u);
        };
        const bool leaf = level == a.bits;
        // pass 1: what would this level add?
        uint32_t my_emit = 0, my_expand = 0;
        for (uint32_t t = tid; t < n_cand; t += kHT) {
            const NodeClass nc = classify_node<NS>(a, c, candidate(t), level);
            if (nc.pop == 0u || (nc.inside | nc.straddle) == 0u) continue;
            const uint32_t pieces = (nc.pop + kSegSplit - 1) / kSegSplit;
            if (nc.straddle == 0u || leaf || nc.pop <= 32u) my_emit += pieces;
            else { my_expand += 1u; my_emit += pieces; }      // budget: a queued node may still have to be emitted whole
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_emit += __shfl_xor_sync(0xffffffffu, my_emit, o);
            my_expand += __shfl_xor_sync(0xffffffffu, my_expand, o);
        }
        if (lane == 0 && (my_emit | my_expand)) { atomicAdd(&s_emit, my_emit); atomicAdd(&s_expand, my_expand); }
        __syncthreads();
        const bool fits = s_nseg + s_emit <= (uint32_t)kSegCap && s_expand <= (uint32_t)kQueueCap;
        const uint32_t n_parents = first ? 0u : s_qn[cur];
        __syncthreads();
        if (tid == 0) { s_emit = 0u; s_expand = 0u; s_qn[cur ^ 1] = 0u; }
        __syncthreads();
        if (!fits) {
            if (first) { if (tid == 0) s_fail = 1u; }
            else {
                // the children do not fit: the queued nodes become segments as they are (their points are tested one by one)
                for (uint32_t t = tid; t < n_parents; t += kHT) emit_segments(classify_node<NS>(a, c, qin[t], level - 1));
            }
            __syncthreads();
            break;
        }
        // pass 2: commit
        for (uint32_t t = tid; t < n_cand; t += kHT) {
            const uint32_t code = candidate(t);
            const NodeClass nc = classify_node<NS>(a, c, code, level);
            if (nc.pop == 0u || (nc.inside | nc.straddle) == 0u) continue;
            if (nc.straddle == 0u || leaf || nc.pop <= 32u) emit_segments(nc);
            else qout[atomicAdd(&s_qn[cur ^ 1], 1u)] = code;
        }
        __syncthreads();
        cur ^= 1;
        first = false;
        ++level;
        if (s_qn[cur] == 0u) break;
    }
    if (s_fail || s_nseg > (uint32_t)kSegCap) {          // cannot happen after the budget check; kept as a guard
        if (tid == 0) a.worklist[atomicAdd(a.work_count, 1)] = (int32_t)b;
        return;
    }
    const uint32_t nseg = s_nseg;

    // ---- key thresholds from the population bounds ----------------------------------------------------------------
    // n_lo <= n <= n_lo + n_bd.  A radius keeps every hit when all of them fit the candidate list; otherwise the keys
    // are thinned to `want` = P + 6 sqrt(P) + 8 expected survivors among the n_lo sure neighbours, provided the
    // straddling cells cannot overflow the list; otherwise the radius waits for its exact count (second pass).
    if (tid < S) {
        const uint32_t nlo = s_nlo[tid], nhi = nlo + s_nbd[tid];
        uint32_t T = 0u, collect = 0u;
        if (nhi <= ccap) { T = 0xFFFFFFFFu; collect = 1u; }
        else if (nlo > a.want) {
            const float expect = (float)a.want * (float)nhi / (float)nlo;
            if (expect + 6.0f * sqrtf(expect) + 16.0f <= (float)ccap) {
                T = (uint32_t)min(4294967295.0, ceil((double)a.want * 4294967296.0 / (double)nlo));
                collect = 1u;
            }
        }
        s_T[tid] = T;
        if (collect) atomicOr(&s_collect, 1u << tid); else atomicOr(&s_late, 1u << tid);
    }
    __syncthreads();

    // ---- scan of the segments: one warp per segment (its classification is warp-uniform) -------------------------
    auto scan = [&](const uint32_t collect, const bool count) {
        uint32_t cnt[MUPS_MAX_SCALES];
#pragma unroll
        for (int s = 0; s < NS; ++s) cnt[s] = 0u;
        uint32_t Ts[MUPS_MAX_SCALES];
#pragma unroll
        for (int s = 0; s < NS; ++s) Ts[s] = s_T[s];
        while (true) {
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(&s_ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= nseg) break;
            const uint32_t start = seg[3 * t], n = seg[3 * t + 1], masks = seg[3 * t + 2];
            const uint32_t inside = masks & 0xFFu, straddle = (masks >> 8) & 0xFFu;
            if (!count && !((inside | straddle) & collect)) continue;
            if (count && lane == 0) {
#pragma unroll
                for (int s = 0; s < NS; ++s) cnt[s] += (inside >> s) & 1u ? n : 0u;
            }
            const uint32_t test = count ? straddle : (straddle & collect);
            if (test == 0u) {
                if (!(inside & collect)) continue;
                // wholly inside: only the points' patch-independent hashes are read (4 bytes per neighbour)
                for (uint32_t i0 = 0; i0 < n; i0 += 128) {
                    uint32_t idx[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = i0 + u * 32 + lane;
                        idx[u] = i < n ? __ldg(a.hash_sorted + start + i) : 0xFFFFFFFFu;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = i0 + u * 32 + lane;
                        if (i >= n) continue;
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            if (!((inside & collect) & (1u << s))) continue;
                            const uint32_t key = (idx[u] ^ salt.a[s]) * salt.b[s];
                            if (key <= Ts[s]) {
                                const uint32_t slot = atomicAdd(s_ncand + s, 1u);
                                if (slot < ccap) cand[s * ccap + slot] = start + i;
                            }
                        }
                    }
                }
            } else {
                for (uint32_t i = lane; i < n; i += 32) {
                    const float4 p = __ldg(a.sorted + start + i);
                    const float dx = p.x - c.cx, dy = p.y - c.cy, dz = p.z - c.cz;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    uint32_t in = inside, band = 0;
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (!(test & (1u << s))) continue;
                        in |= (d2 < a.r2_lo[s] ? 1u : 0u) << s;
                        band |= (d2 >= a.r2_lo[s] && d2 <= a.r2_hi[s] ? 1u : 0u) << s;
                    }
                    if (band) {                     // rare: decide with cKDTree's float64 predicate
#pragma unroll
                        for (int s = 0; s < NS; ++s)
                            if ((band >> s) & 1u) in |= (inside_exact(c, p, a.r2[s]) ? 1u : 0u) << s;
                    }
                    if (count) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) cnt[s] += (in & test) >> s & 1u;
                    }
                    in &= collect;
                    if (!in) continue;
                    const uint32_t hsh = fmix32((uint32_t)__float_as_int(p.w));
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (!(in & (1u << s))) continue;
                        const uint32_t key = (hsh ^ salt.a[s]) * salt.b[s];
                        if (key <= Ts[s]) {
                            const uint32_t slot = atomicAdd(s_ncand + s, 1u);
                            if (slot < ccap) cand[s * ccap + slot] = start + i;
                        }
                    }
                }
            }
        }
        if (count) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                uint32_t v = cnt[s];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v) atomicAdd(s_cnt + s, v);
            }
        }
    };
    scan(s_collect, true);
    __syncthreads();
    const uint32_t late = s_late;
    if (late) {
        // radii whose population bounds were too loose: now the count is exact; second pass over their segments only
        if (tid < S && (late & (1u << tid))) {
            const uint32_t n = s_cnt[tid];
            s_T[tid] = n <= ccap ? 0xFFFFFFFFu
                                 : (uint32_t)min(4294967295.0, ceil((double)a.want * 4294967296.0 / (double)n));
        }
        if (tid == 0) s_ticket = 0u;
        __syncthreads();
        scan(late, false);
        __syncthreads();
    }
    // every over-full radius needs at least P candidates, none may have overflowed
    if (tid < S) {
        const uint32_t nc = s_ncand[tid], n = s_cnt[tid];
        if (nc > ccap || nc < min((uint32_t)P, n)) s_fail = 1u;
    }
    for (int i = tid; i < S * kBins; i += kHT) hist[i] = 0u;      // (the level queues are dead)
    __syncthreads();
    uint32_t over = 0;
    for (int s = 0; s < S; ++s) over |= (s_cnt[s] > (uint32_t)P) ? (1u << s) : 0u;

    // ---- the P smallest keys among the candidates: 9-bit histogram of the key scaled to the threshold ------------
    if (!s_fail && over) {
        if (tid < S) {
            const uint32_t T = s_T[tid];
            s_scale[tid] = T == 0xFFFFFFFFu ? (uint32_t)kBins
                                            : (uint32_t)min(4294967295ull, ((unsigned long long)kBins << 32) / ((unsigned long long)T + 1ull));
        }
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            if (!(over & (1u << s))) continue;
            const uint32_t nc = s_ncand[s], sc = s_scale[s];
            for (uint32_t i = tid; i < nc; i += kHT) {
                const uint32_t key = (__ldg(a.hash_sorted + cand[s * ccap + i]) ^ salt.a[s]) * salt.b[s];
                atomicAdd(hist + s * kBins + min((uint32_t)(kBins - 1), __umulhi(key, sc)), 1u);
            }
        }
        __syncthreads();
        if (warp < S && (over & (1u << warp))) {
            uint32_t T, below, group;
            warp_find_threshold(hist + warp * kBins, kBins, (uint32_t)P, lane, &T, &below, &group);
            if (lane == 0) {
                s_tbin[warp] = T; s_need[warp] = (uint32_t)P - below;
                if (group > (uint32_t)kBoundaryCap) s_fail = 1u;
            }
        }
        __syncthreads();
    }
    if (s_fail) {
        if (tid == 0) a.worklist[atomicAdd(a.work_count, 1)] = (int32_t)b;
        return;
    }
    for (int s = 0; s < S; ++s) {
        const uint32_t nc = s_ncand[s];
        const bool thin = (over >> s) & 1u;
        const uint32_t sc = s_scale[s], tb = s_tbin[s];
        for (uint32_t i = tid; i < nc; i += kHT) {
            const uint32_t pos = cand[s * ccap + i];
            const uint32_t idx = (uint32_t)__float_as_int(__ldg(&a.sorted[pos].w));
            bool take = true;
            if (thin) {
                const uint32_t key = (__ldg(a.hash_sorted + pos) ^ salt.a[s]) * salt.b[s];
                const uint32_t bin = min((uint32_t)(kBins - 1), __umulhi(key, sc));
                take = bin < tb;
                if (bin == tb) {
                    const uint32_t slot = atomicAdd(s_nb + s, 1u);
                    if (slot < (uint32_t)kBoundaryCap) {
                        bnd[s * kBoundaryCap + slot] = ((unsigned long long)key << 32) | idx;
                        bnd_pos[s * kBoundaryCap + slot] = pos;
                    }
                }
            }
            if (take) {
                const uint32_t slot = atomicAdd(s_nsel + s, 1u);
                if (slot < (uint32_t)Ppad) sel[(size_t)s * Ppad + slot] = ((unsigned long long)idx << 32) | pos;
            }
        }
    }
    __syncthreads();                                       // cand is dead from here: tmp may overwrite it

    PatchLists L;
    L.hist = hist; L.sel = sel; L.tmp = tmp; L.bndl = bnd; L.bndp = bnd_pos; L.bcap = (uint32_t)kBoundaryCap;
    L.s_cnt = s_cnt; L.s_nsel = s_nsel; L.s_nb = s_nb; L.s_need = s_need; L.s_min = s_min; L.s_max = s_max;
    L.s_warp_sums = s_warp_sums;
    finish_patch<NS, kHT>(a, c, b, L);
}

// ---- CTA order: the centres' positions in `sorted` are Morton-ordered, so a counting sort of the batch rows by
// (position / bucket width) makes consecutive CTAs work on overlapping balls (L2 reuse on clouds larger than L2) ----
constexpr int kOrderBuckets = 1 << 16;

__global__ void __launch_bounds__(256) order_count_kernel(const int64_t* __restrict__ q, int64_t B, int64_t n,
                                                          const int32_t* __restrict__ pos_of, uint32_t width,
                                                          uint32_t* __restrict__ bucket) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = q[b];
        const uint32_t k = (qi >= 0 && qi < n) ? (uint32_t)__ldg(pos_of + qi) / width : 0u;
        atomicAdd(bucket + k, 1u);
    }
}

__global__ void __launch_bounds__(256) order_fill_kernel(const int64_t* __restrict__ q, int64_t B, int64_t n,
                                                         const int32_t* __restrict__ pos_of, uint32_t width,
                                                         uint32_t* __restrict__ bucket, int32_t* __restrict__ order) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = q[b];
        const uint32_t k = (qi >= 0 && qi < n) ? (uint32_t)__ldg(pos_of + qi) / width : 0u;
        order[atomicAdd(bucket + k, 1u)] = (int32_t)b;
    }
}

int launch_ball_query(const mups_index* ix, const int64_t* q, int64_t B, const double* r_abs, int S, int P,
                      uint64_t seed, int32_t* nbr_idx, int32_t* nbr_total, float* patches, int32_t* n_eff,
                      int32_t* nbr_pos, cudaStream_t st) {
    QueryArgs a;
    a.sorted = ix->sorted; a.cell_start = ix->cell_start; a.pos_of = ix->pos_of; a.grid = ix->grid;
    a.q = q; a.n = ix->n; a.S = S; a.P = P;
    int ppad = 1;
    while (ppad < P) ppad <<= 1;
    a.Ppad = ppad;
    a.cap = g_boundary_cap.load();
    a.fuse = (uint32_t)g_fuse_candidates.load();
    double rmax = 0.0;
    float hi_max = 0.f;
    int ord[MUPS_MAX_SCALES];
    for (int s = 0; s < MUPS_MAX_SCALES; ++s) ord[s] = s;
    for (int i = 1; i < S; ++i)                                  // stable insertion sort of the scales by radius
        for (int j = i; j > 0 && r_abs[ord[j]] < r_abs[ord[j - 1]]; --j) { const int t = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = t; }
    for (int s = 0; s < MUPS_MAX_SCALES; ++s) {
        a.orig[s] = ord[s];
        const double r = s < S ? r_abs[ord[s]] : 0.0;
        a.r2[s] = r * r;
        a.r2_lo[s] = (float)(a.r2[s] * (1.0 - 4e-6));
        a.r2_hi[s] = (float)(a.r2[s] * (1.0 + 4e-6));
        a.rf[s] = (float)r;
        if (r > rmax) rmax = r;
        if (a.r2_hi[s] > hi_max) hi_max = a.r2_hi[s];
    }
    a.r_max = (float)(rmax * (1.0 + 1e-6));
    a.r2_hi_max = hi_max;
    a.k0 = (uint32_t)(seed & 0xFFFFFFFFull); a.k1 = (uint32_t)(seed >> 32);
    a.nbr_idx = nbr_idx; a.nbr_total = nbr_total; a.patches = patches; a.n_eff = n_eff; a.nbr_pos = nbr_pos;
    a.hash_sorted = ix->hash_sorted; a.bits = ix->bits;
    const int margin = g_hier_margin.load();
    a.want = (uint32_t)P + (margin >= 0 ? (uint32_t)margin : (uint32_t)std::ceil(6.0 * std::sqrt((double)P)) + 8u);
    a.ccap = (uint32_t)(P <= 512 ? (3 * P > 1024 ? 3 * P : 1024) : (5 * P) / 2);
    a.order = nullptr; a.worklist = nullptr; a.work_count = nullptr;

    const size_t region = (size_t)kHitCap * 5 > (size_t)S * ppad * 8 ? (size_t)kHitCap * 5 : (size_t)S * ppad * 8;
    const size_t smem = (size_t)S * kBins * 4 + (size_t)S * ppad * 8 + (size_t)S * kBoundaryCap * 12 + region;
    if (smem > 200 * 1024) {
        set_error("ball query: S=%d, P=%d needs %zu bytes of shared memory", S, P, smem);
        return MUPS_ERR_UNSUPPORTED;
    }
    if (B == 0) return MUPS_OK;

    // hierarchical kernel: fine grids (the host picks them for dense clouds); its lists must fit next to 2 other CTAs
    const size_t hq = (size_t)(S * kBins > 2 * kQueueCap ? S * kBins : 2 * kQueueCap) * 4;
    const size_t region_h = (size_t)S * a.ccap * 4 > (size_t)S * ppad * 8 ? (size_t)S * a.ccap * 4 : (size_t)S * ppad * 8;
    const size_t smem_h = (size_t)kSegCap * 12 + hq + (size_t)S * ppad * 8 + (size_t)S * kBoundaryCap * 12 + region_h;
    const int mode = g_query_kernel.load();
    const bool hier = mode == 2 || (mode == 0 && ix->bits >= 6);
    const bool use_hier = hier && smem_h <= 112 * 1024 && ix->hash_sorted != nullptr;

    cudaMemPool_t pool = nullptr;
    int32_t* scratch = nullptr;      // [order: B][worklist: B][work_count: 1][buckets + scan tiles]
    const int order_mode = g_query_order.load();
    const bool reorder = order_mode == 2 || (order_mode == 0 && use_hier && B >= 8192 && ix->n >= (1 << 20));
    if (use_hier || reorder) {
        if (int rc = library_pool(ix->device, &pool)) return rc;
        const size_t words = (size_t)2 * B + 8 + kOrderBuckets + 64;
        MUPS_CUDA_TRY(cudaMallocFromPoolAsync((void**)&scratch, sizeof(int32_t) * words, pool, st));
    }
    if (reorder) {
        int32_t* order = scratch;
        uint32_t* bucket = reinterpret_cast<uint32_t*>(scratch + 2 * B + 8);
        const uint32_t width = (uint32_t)((ix->n + kOrderBuckets - 1) / kOrderBuckets);
        MUPS_CUDA_TRY(cudaMemsetAsync(bucket, 0, sizeof(uint32_t) * kOrderBuckets, st));
        const int grid = (int)((B + 255) / 256 < 4 * kNumSMs ? (B + 255) / 256 : 4 * kNumSMs);
        order_count_kernel<<<grid, 256, 0, st>>>(q, B, ix->n, ix->pos_of, width, bucket);
        MUPS_CHECK_LAUNCH();
        if (int rc = launch_exclusive_scan(bucket, kOrderBuckets, bucket + kOrderBuckets, st)) return rc;
        order_fill_kernel<<<grid, 256, 0, st>>>(q, B, ix->n, ix->pos_of, width, bucket, order);
        MUPS_CHECK_LAUNCH();
        a.order = order;
    }

#define MUPS_LAUNCH_QUERY(KERNEL, NS, GRID, THREADS, SMEM)                                                            \
    case NS:                                                                                                           \
        if ((SMEM) > 48 * 1024)                                                                                        \
            MUPS_CUDA_TRY(cudaFuncSetAttribute(KERNEL<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM))); \
        KERNEL<NS><<<(unsigned)(GRID), THREADS, SMEM, st>>>(a);                                                        \
        break;
    if (use_hier) {
        a.worklist = scratch + B;
        a.work_count = scratch + 2 * B;
        MUPS_CUDA_TRY(cudaMemsetAsync(a.work_count, 0, sizeof(int32_t), st));
        switch (S) {
            MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 1, B, kHT, smem_h) MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 2, B, kHT, smem_h)
            MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 3, B, kHT, smem_h) MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 4, B, kHT, smem_h)
            MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 5, B, kHT, smem_h) MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 6, B, kHT, smem_h)
            MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 7, B, kHT, smem_h) MUPS_LAUNCH_QUERY(ball_query_hier_kernel, 8, B, kHT, smem_h)
        }
        MUPS_CHECK_LAUNCH();
        if (getenv("MUPS_DEBUG_WORKLIST")) {              // diagnostics only: synchronises
            int32_t handed = -1;
            cudaMemcpyAsync(&handed, a.work_count, sizeof(handed), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            fprintf(stderr, "[mups] hierarchical ball query: %d of %lld rows handed to the flat kernel\n", handed, (long long)B);
        }
        // the rows it could not decide (statistical tail of the key threshold): the flat kernel, persistent over the worklist
        a.order = nullptr;
        const int64_t grid = B < 2 * kNumSMs ? B : 2 * kNumSMs;
        switch (S) {
            MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 1, grid, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 2, grid, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 3, grid, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 4, grid, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 5, grid, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 6, grid, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 7, grid, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_worklist_kernel, 8, grid, kQT, smem)
        }
        MUPS_CHECK_LAUNCH();
    } else {
        switch (S) {
            MUPS_LAUNCH_QUERY(ball_query_kernel, 1, B, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_kernel, 2, B, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_kernel, 3, B, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_kernel, 4, B, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_kernel, 5, B, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_kernel, 6, B, kQT, smem)
            MUPS_LAUNCH_QUERY(ball_query_kernel, 7, B, kQT, smem) MUPS_LAUNCH_QUERY(ball_query_kernel, 8, B, kQT, smem)
        }
        MUPS_CHECK_LAUNCH();
    }
#undef MUPS_LAUNCH_QUERY
    if (scratch) MUPS_CUDA_TRY(cudaFreeAsync(scratch, st));
    return MUPS_OK;
}

}  // namespace mups
