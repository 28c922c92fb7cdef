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
    float r_max;
    float r2_hi_max;
    uint32_t k0, k1;                // philox key = seed
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
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const uint32_t exc = inc - cnt[0] - cnt[1];
            st.start[2 * lane] = beg[0];
            st.start[2 * lane + 1] = beg[1];
            st.prefix[2 * lane] = exc;
            st.prefix[2 * lane + 1] = exc + cnt[0];
            if (lane == 31) {
                st.prefix[kRangeCap] = inc;
                if (cbase == 0) st.first_total = inc;
            }
        }
        __syncthreads();
        const uint32_t total = st.prefix[kRangeCap];
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

// key of neighbour `idx` for radius s: fmix32((idx ^ a_s) * b_s), a bijection of idx keyed by the salt
// (a_s, b_s | 1) = Philox4x32-10(counter = (centre, s, 0, 0), key = seed) computed once per CTA and radius
struct Salts {
    uint32_t a[MUPS_MAX_SCALES], b[MUPS_MAX_SCALES];
};
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
template <int NS>
__device__ __forceinline__ void selection_keys(const Salts& salt, uint32_t idx, uint32_t need_mask,
                                               uint32_t key[MUPS_MAX_SCALES]) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
        if (need_mask & (1u << s)) key[s] = fmix32((idx ^ salt.a[s]) * salt.b[s]);
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

// Exclusive scan of kBins counters by the whole CTA (kBins / kQT per thread), in place.
__device__ __forceinline__ void block_scan_bins(uint32_t* bins, uint32_t* warp_sums /*[kQT/32]*/) {
    constexpr int PER = kBins / kQT;
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

template <int NS>
__global__ void __launch_bounds__(kQT) ball_query_kernel(const QueryArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
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

    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = a.q[b];

    for (int i = tid; i < S * kBins; i += kQT) hist[i] = 0u;
    if (tid < MUPS_MAX_SCALES) {
        s_cnt[tid] = 0u; s_nsel[tid] = 0u; s_nb[tid] = 0u; s_prefix[tid] = 0u; s_bits[tid] = 0u; s_need[tid] = 0u;
        s_min[tid] = 0xFFFFFFFFu; s_max[tid] = 0u;
    }
    if (tid == 0) { s_unresolved = 0u; s_nhits = 0u; }

    if (q < 0 || q >= a.n) {   // invalid centre: empty patch, total = -1
        for (int i = tid; i < S * P; i += kQT) {
            if (a.patches) { float* o = a.patches + (b * S * P + i) * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
            if (a.nbr_idx) a.nbr_idx[b * S * P + i] = -1;
        }
        if (tid < S) { a.n_eff[b * S + tid] = 0; if (a.nbr_total) a.nbr_total[b * S + tid] = -1; }
        return;
    }

    QueryCtx c;
    {
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
    }
    if (tid < S) {   // per-patch randomness: one Philox call per radius
        const uint4 w = philox4x32_10((uint32_t)q, (uint32_t)tid, 0u, 0u, a.k0, a.k1);
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
            const uint32_t slot = atomicAdd(&s_nhits, 1u);
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

    // ---- threshold group: keep the `need` smallest (key, index) pairs ------------------------------------
    for (int s = 0; s < S; ++s) {
        const uint32_t m = min(s_nb[s], bcap);
        const uint32_t need = s_need[s];
        for (uint32_t i = tid; i < m; i += kQT) {
            const unsigned long long mine = bndl[s * bcap + i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) rank += bndl[s * bcap + j] < mine ? 1u : 0u;
            if (rank < need) {
                const uint32_t slot = atomicAdd(s_nsel + s, 1u);
                if (slot < (uint32_t)Ppad)
                    sel[(size_t)s * Ppad + slot] = ((mine & 0xFFFFFFFFull) << 32) | bndp[s * bcap + i];
            }
        }
    }
    __syncthreads();

    // ---- order every radius' selection by point index: bucket sort over the index range --------------------
    // bucket(e) is monotone in the index, so bucket order + order inside a bucket = index order
    for (int i = tid; i < S * kBins; i += kQT) hist[i] = 0u;
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(s_cnt[s], (uint32_t)P);
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        for (uint32_t t = tid; t < ne; t += kQT) {
            const uint32_t id = (uint32_t)(sel[(size_t)s * Ppad + t] >> 32);
            lo = min(lo, id); hi = max(hi, id);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0 && lo <= hi) { atomicMin(s_min + s, lo); atomicMax(s_max + s, hi); }
    }
    __syncthreads();
    auto bucket_of = [&](int s, uint32_t id) {
        const float scale = (float)kBins / ((float)(s_max[s] - s_min[s]) + 1.0f);
        return min((uint32_t)(kBins - 1), (uint32_t)((float)(id - s_min[s]) * scale));
    };
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += kQT)
            atomicAdd(hist + s * kBins + bucket_of(s, (uint32_t)(sel[(size_t)s * Ppad + t] >> 32)), 1u);
    }
    __syncthreads();
    for (int s = 0; s < S; ++s) block_scan_bins(hist + s * kBins, s_warp_sums);     // counts -> bucket starts
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += kQT) {
            const unsigned long long e = sel[(size_t)s * Ppad + t];
            const uint32_t slot = atomicAdd(hist + s * kBins + bucket_of(s, (uint32_t)(e >> 32)), 1u);   // starts -> ends
            tmp[(size_t)s * Ppad + slot] = e;
        }
    }
    __syncthreads();
    for (int s = 0; s < S; ++s) {
        const uint32_t ne = min(s_cnt[s], (uint32_t)P);
        for (uint32_t t = tid; t < ne; t += kQT) {
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
        const uint32_t total = s_cnt[s];
        const uint32_t ne = min(total, (uint32_t)P);
        const unsigned long long* v = sel + (size_t)s * Ppad;
        const float rf = a.rf[s];
        for (uint32_t t = tid; t < (uint32_t)P; t += kQT) {
            float ox = 0.f, oy = 0.f, oz = 0.f;
            int32_t id = -1;
            if (t < ne) {
                const unsigned long long e = v[t];
                id = (int32_t)(e >> 32);
                const float4 p = __ldg(a.sorted + (uint32_t)(e & 0xFFFFFFFFull));
                ox = __fdiv_rn(__fsub_rn(p.x, c.cx), rf);
                oy = __fdiv_rn(__fsub_rn(p.y, c.cy), rf);
                oz = __fdiv_rn(__fsub_rn(p.z, c.cz), rf);
            }
            if (a.patches) {
                float* o = a.patches + ((b * S + s) * (int64_t)P + t) * 3;
                o[0] = ox; o[1] = oy; o[2] = oz;
            }
            if (a.nbr_idx) a.nbr_idx[(b * S + s) * (int64_t)P + t] = id;
        }
        if (tid == 0) {
            a.n_eff[b * S + s] = (int32_t)ne;
            if (a.nbr_total) a.nbr_total[b * S + s] = (int32_t)total;
        }
    }
}

int launch_ball_query(const mups_index* ix, const int64_t* q, int64_t B, const double* r_abs, int S, int P,
                      uint64_t seed, int32_t* nbr_idx, int32_t* nbr_total, float* patches, int32_t* n_eff,
                      cudaStream_t st) {
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
    for (int s = 0; s < MUPS_MAX_SCALES; ++s) {
        const double r = s < S ? r_abs[s] : 0.0;
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
    a.nbr_idx = nbr_idx; a.nbr_total = nbr_total; a.patches = patches; a.n_eff = n_eff;
    const size_t region = (size_t)kHitCap * 5 > (size_t)S * ppad * 8 ? (size_t)kHitCap * 5 : (size_t)S * ppad * 8;
    const size_t smem = (size_t)S * kBins * 4 + (size_t)S * ppad * 8 + (size_t)S * kBoundaryCap * 12 + region;
    if (smem > 200 * 1024) {
        set_error("ball query: S=%d, P=%d needs %zu bytes of shared memory", S, P, smem);
        return MUPS_ERR_UNSUPPORTED;
    }
    if (B > 0) {
#define MUPS_LAUNCH_QUERY(NS)                                                                                          \
    case NS:                                                                                                           \
        if (smem > 48 * 1024)                                                                                          \
            MUPS_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ball_query_kernel<NS><<<(unsigned)B, kQT, smem, st>>>(a);                                                      \
        break;
        switch (S) {
            MUPS_LAUNCH_QUERY(1) MUPS_LAUNCH_QUERY(2) MUPS_LAUNCH_QUERY(3) MUPS_LAUNCH_QUERY(4)
            MUPS_LAUNCH_QUERY(5) MUPS_LAUNCH_QUERY(6) MUPS_LAUNCH_QUERY(7) MUPS_LAUNCH_QUERY(8)
        }
#undef MUPS_LAUNCH_QUERY
        MUPS_CHECK_LAUNCH();
    }
    return MUPS_OK;
}

}  // namespace mups
