// C ABI of libmups_b200.so (declared in include/mups.h): argument validation, handle lifetime,
// error strings.  All compute lives in the kernels of mups_index.cu / mups_query.cu / mups_stats.cu.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "mups_common.cuh"

namespace mups {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launch_count{0};
std::atomic<int> g_boundary_cap{512};
std::atomic<int> g_fuse_candidates{3 * 4096};
std::atomic<int> g_stats_variant{0};
std::atomic<int> g_pool_variant{0};
std::atomic<int> g_conv_variant{0};
std::atomic<int> g_query_kernel{0};
std::atomic<int> g_query_order{0};
std::atomic<int> g_hier_margin{-1};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int library_pool(int device, cudaMemPool_t* pool) {
    static std::mutex mtx;
    static cudaMemPool_t pools[64] = {nullptr};
    MUPS_REQUIRE(device >= 0 && device < 64, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(mtx);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t p = nullptr;
        MUPS_CUDA_TRY(cudaMemPoolCreate(&p, &props));
        uint64_t keep = ~0ull;
        MUPS_CUDA_TRY(cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep));
        pools[device] = p;
    }
    *pool = pools[device];
    return MUPS_OK;
}

// remembers that `st` has work queued that reads the index (the event is re-recorded per call)
static int note_use(const mups_index* ix, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(ix->use_mutex);
    for (auto& u : ix->uses) {
        if (u.first == st) {
            MUPS_CUDA_TRY(cudaEventRecord(u.second, st));
            return MUPS_OK;
        }
    }
    cudaEvent_t ev = nullptr;
    MUPS_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MUPS_CUDA_TRY(cudaEventRecord(ev, st));
    ix->uses.emplace_back(st, ev);
    return MUPS_OK;
}

static int check_device(int handle_device, const char* what) {
    int dev = -1;
    MUPS_CUDA_TRY(cudaGetDevice(&dev));
    MUPS_REQUIRE(dev == handle_device, "%s: handle belongs to device %d but device %d is current", what, handle_device, dev);
    return MUPS_OK;
}

}  // namespace mups

using namespace mups;

extern "C" {

int mups_abi_version(void) { return MUPS_ABI_VERSION; }
const char* mups_last_error(void) { return t_error; }
int64_t mups_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int mups_set_option(const char* name, int64_t value) {
    MUPS_REQUIRE(name != nullptr, "mups_set_option: name is NULL");
    if (!strcmp(name, "boundary_cap")) {
        MUPS_REQUIRE(value >= 1 && value <= 512, "mups_set_option: boundary_cap=%lld out of range [1, 512]", (long long)value);
        g_boundary_cap.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "fuse_candidates")) {
        MUPS_REQUIRE(value >= 0 && value <= 0x7FFFFFFF, "mups_set_option: fuse_candidates=%lld out of range", (long long)value);
        g_fuse_candidates.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "stats_variant")) {
        MUPS_REQUIRE(value >= 0 && value <= 64, "mups_set_option: stats_variant=%lld out of range", (long long)value);
        g_stats_variant.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "query_kernel")) {
        MUPS_REQUIRE(value >= 0 && value <= 2, "mups_set_option: query_kernel=%lld out of range [0, 2]", (long long)value);
        g_query_kernel.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "hier_margin")) {
        MUPS_REQUIRE(value >= -1 && value <= 4096, "mups_set_option: hier_margin=%lld out of range [-1, 4096]", (long long)value);
        g_hier_margin.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "query_order")) {
        MUPS_REQUIRE(value >= 0 && value <= 2, "mups_set_option: query_order=%lld out of range [0, 2]", (long long)value);
        g_query_order.store((int)value);
        return MUPS_OK;
    }
    if (!strcmp(name, "pool_variant") || !strcmp(name, "conv_variant")) {
        MUPS_REQUIRE(value >= 0 && value <= (name[0] == 'p' ? 1 : 9), "mups_set_option: %s=%lld out of range", name, (long long)value);
        (name[0] == 'p' ? g_pool_variant : g_conv_variant).store((int)value);
        return MUPS_OK;
    }
    set_error("mups_set_option: unknown option '%s'", name);
    return MUPS_ERR_INVALID;
}

// ---- index ------------------------------------------------------------------------------------
void mups_index_destroy(mups_index* ix) {
    if (!ix) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ix->device);
    // stream-ordered: the frees are queued on the build stream behind every stream that queried the index;
    // nothing here blocks the host or synchronises the device
    for (auto& u : ix->uses) {
        if (u.first != ix->build_stream) cudaStreamWaitEvent(ix->build_stream, u.second, 0);
        cudaEventDestroy(u.second);
    }
    if (ix->built) cudaEventDestroy(ix->built);
    if (ix->grid) cudaFreeAsync(ix->grid, ix->build_stream);
    if (ix->sorted) cudaFreeAsync(ix->sorted, ix->build_stream);
    if (ix->cell_start) cudaFreeAsync(ix->cell_start, ix->build_stream);
    if (ix->pos_of) cudaFreeAsync(ix->pos_of, ix->build_stream);
    if (ix->codes) cudaFreeAsync(ix->codes, ix->build_stream);
    if (ix->hash_sorted) cudaFreeAsync(ix->hash_sorted, ix->build_stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete ix;
}

int mups_index_create(mups_index** out, const float* xyz_dev, int64_t n, double cell_frac, mups_stream stream) {
    MUPS_REQUIRE(out != nullptr, "mups_index_create: out is NULL");
    *out = nullptr;
    MUPS_REQUIRE(xyz_dev != nullptr, "mups_index_create: xyz is NULL");
    MUPS_REQUIRE(n >= 1 && n < (int64_t)0x7FFFFFFF, "mups_index_create: n=%lld out of range [1, 2^31)", (long long)n);
    if (!(cell_frac > 0.0)) cell_frac = 0.07;
    if (cell_frac > 1.0) cell_frac = 1.0;
    mups_index* ix = new (std::nothrow) mups_index();
    if (!ix) { set_error("mups_index_create: out of host memory"); return MUPS_ERR_NOMEM; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = MUPS_OK;
    auto fail = [&](int code) { mups_index_destroy(ix); return code; };
    {
        cudaError_t e = cudaGetDevice(&ix->device);
        if (e != cudaSuccess) { set_error("no usable CUDA device: %s", cudaGetErrorString(e)); return fail(MUPS_ERR_CUDA); }
    }
    ix->n = n;
    ix->cell_frac = cell_frac;
    // extent <= diagonal, so floor(1/cell_frac) + 2 cells per axis always suffice; capped at 256 (8 Morton bits)
    int max_dim = (int)std::floor(1.0 / cell_frac) + 2;
    if (max_dim > 256) max_dim = 256;
    ix->max_dim = max_dim;
    int bits = 1;
    while ((1 << bits) < max_dim) ++bits;
    ix->bits = bits;
    const int64_t ncode = (int64_t)1 << (3 * bits);
    const int64_t n_scan = ncode + 1;
    const int64_t n_tiles = (n_scan + 2047) / 2048;
    cudaMemPool_t pool = nullptr;
    ix->build_stream = st;
    if ((rc = library_pool(ix->device, &pool))) return fail(rc);
    auto alloc = [&](void** p, size_t bytes) -> int {
        cudaError_t e = cudaMallocFromPoolAsync(p, bytes, pool, st);
        if (e != cudaSuccess) {
            set_error("mups_index_create: allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? MUPS_ERR_NOMEM : MUPS_ERR_CUDA;
        }
        return MUPS_OK;
    };
    if ((rc = alloc((void**)&ix->grid, sizeof(GridDesc)))) return fail(rc);
    if ((rc = alloc((void**)&ix->sorted, sizeof(float4) * (size_t)n))) return fail(rc);
    if ((rc = alloc((void**)&ix->cell_start, sizeof(uint32_t) * (size_t)(n_scan + n_tiles + 8)))) return fail(rc);
    if ((rc = alloc((void**)&ix->pos_of, sizeof(int32_t) * (size_t)n))) return fail(rc);
    if ((rc = alloc((void**)&ix->codes, sizeof(uint32_t) * (size_t)(n < 8 ? 8 : n)))) return fail(rc);
    if ((rc = alloc((void**)&ix->hash_sorted, sizeof(uint32_t) * (size_t)n))) return fail(rc);
    ix->build_stream = st;
    if ((rc = launch_index_build(ix, xyz_dev, st))) return fail(rc);
    if (cudaEventCreateWithFlags(&ix->built, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventRecord(ix->built, st) != cudaSuccess) {
        set_error("mups_index_create: event setup failed");
        return fail(MUPS_ERR_CUDA);
    }
    *out = ix;
    return MUPS_OK;
}

int mups_index_bbox(const mups_index* ix, float min3_host[3], float max3_host[3]) {
    MUPS_REQUIRE(ix && min3_host && max3_host, "mups_index_bbox: NULL argument");
    if (int rc = check_device(ix->device, "mups_index_bbox")) return rc;
    if (!ix->have_bbox) {
        MUPS_CUDA_TRY(cudaEventSynchronize(ix->built));
        GridDesc g;
        MUPS_CUDA_TRY(cudaMemcpy(&g, ix->grid, sizeof(g), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 3; ++k) { ix->bb[k] = g.bb_min[k]; ix->bb[3 + k] = g.bb_max[k]; }
        ix->have_bbox = true;
    }
    for (int k = 0; k < 3; ++k) { min3_host[k] = ix->bb[k]; max3_host[k] = ix->bb[3 + k]; }
    return MUPS_OK;
}

int64_t mups_index_size(const mups_index* ix) { return ix ? ix->n : 0; }

// ---- half 1 -------------------------------------------------------------------------------------
static int query_common(const char* what, const mups_index* ix, const int64_t* query_idx_dev, int64_t B, const double* r_abs_host,
                        int S, int P, uint64_t seed, int32_t* nbr_idx_dev, int32_t* nbr_total_dev, float* patches_dev,
                        int32_t* n_eff_dev, int32_t* nbr_pos_dev, cudaStream_t st);

int mups_ball_query(const mups_index* ix, const int64_t* query_idx_dev, int64_t B, const double* r_abs_host, int S,
                    int P, uint64_t seed, int32_t* nbr_idx_dev, int32_t* nbr_total_dev, float* patches_dev,
                    int32_t* n_eff_dev, mups_stream stream) {
    return query_common("mups_ball_query", ix, query_idx_dev, B, r_abs_host, S, P, seed, nbr_idx_dev, nbr_total_dev, patches_dev,
                        n_eff_dev, nullptr, static_cast<cudaStream_t>(stream));
}

// ---- GMM ---------------------------------------------------------------------------------------
void mups_gmm_destroy(mups_gmm* g) {
    if (!g) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    cudaFree(g->A);
    cudaFree(g->Bv);
    cudaFree(g->C);
    cudaFree(g->axis_mu);
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
}

int mups_gmm_create(mups_gmm** out, const float* w, const float* mu, const float* sigma, int G) {
    MUPS_REQUIRE(out != nullptr, "mups_gmm_create: out is NULL");
    *out = nullptr;
    MUPS_REQUIRE(w && mu && sigma, "mups_gmm_create: NULL parameter array");
    MUPS_REQUIRE(G >= 1 && G <= 32768, "mups_gmm_create: G=%d out of range [1, 32768]", G);
    for (int g = 0; g < G; ++g) {
        MUPS_REQUIRE(w[g] > 0.f && std::isfinite(w[g]), "mups_gmm_create: w[%d]=%g must be positive", g, w[g]);
        for (int k = 0; k < 3; ++k) {
            MUPS_REQUIRE(sigma[3 * g + k] > 0.f && std::isfinite(sigma[3 * g + k]), "mups_gmm_create: sigma[%d][%d]=%g must be positive",
                         g, k, sigma[3 * g + k]);
            MUPS_REQUIRE(std::isfinite(mu[3 * g + k]), "mups_gmm_create: mu[%d][%d] is not finite", g, k);
        }
    }
    mups_gmm* gm = new (std::nothrow) mups_gmm();
    if (!gm) { set_error("mups_gmm_create: out of host memory"); return MUPS_ERR_NOMEM; }
    auto fail = [&](int code) { mups_gmm_destroy(gm); return code; };
    {
        cudaError_t e = cudaGetDevice(&gm->device);
        if (e != cudaSuccess) { set_error("no usable CUDA device: %s", cudaGetErrorString(e)); return fail(MUPS_ERR_CUDA); }
    }
    gm->G = G;
    std::vector<float4> A(G), Bv(G), C(G);
    const float two_pi_pow = powf((float)(2.0 * M_PI), 1.5f);
    for (int g = 0; g < G; ++g) {
        const float sx = sigma[3 * g], sy = sigma[3 * g + 1], sz = sigma[3 * g + 2];
        // masked variant: prefactor 1/((2pi)^1.5 * sigma_x^3) (tf_util.py:687 uses batch_sig[...,0] only);
        // plain variant: MultivariateNormalDiag -> product of the three sigmas (tf_util.py:606-608)
        const double pm = 1.0 / ((double)two_pi_pow * (double)sx * (double)sx * (double)sx);
        const double pp = 1.0 / ((double)two_pi_pow * (double)sx * (double)sy * (double)sz);
        A[g] = make_float4(mu[3 * g], mu[3 * g + 1], mu[3 * g + 2], (float)std::log2((double)w[g] * pm));
        Bv[g] = make_float4(1.0f / sx, 1.0f / sy, 1.0f / sz, (float)std::log2((double)w[g] * pp));
        const float rsw = 1.0f / sqrtf(w[g]);
        C[g] = make_float4(w[g], rsw, 1.0f / sqrtf(2.0f * w[g]), -w[g] * rsw);
    }
    // separable lattice detection: mu_g = (X_i, Y_j, Z_k) with g = (i*ny + j)*nz + k (np.mgrid order,
    // utils/utils.py:83-86), one sigma per axis shared by all Gaussians, uniform w
    {
        int nz = 1;
        while (nz < G && mu[3 * nz] == mu[0] && mu[3 * nz + 1] == mu[1]) ++nz;
        int ny = 1;
        while (ny * nz < G && mu[3 * ny * nz] == mu[0]) ++ny;
        const int nx = (ny * nz > 0 && G % (ny * nz) == 0) ? G / (ny * nz) : 0;
        bool sep = nx >= 1 && nx <= 64 && ny <= 64 && nz <= 64 && nx * ny * nz == G;
        for (int g = 0; sep && g < G; ++g) {
            const int i = g / (ny * nz), j = (g / nz) % ny, k = g % nz;
            sep = mu[3 * g] == mu[3 * (i * ny * nz)] && mu[3 * g + 1] == mu[3 * (j * nz) + 1] && mu[3 * g + 2] == mu[3 * k + 2] &&
                  sigma[3 * g] == sigma[0] && sigma[3 * g + 1] == sigma[1] && sigma[3 * g + 2] == sigma[2] && w[g] == w[0];
        }
        // the fast path's guard assumes that inside the lattice hull some centre is always within
        // 5 sigma per axis: consecutive lattice coordinates at most 10 sigma apart, ascending
        for (int ax = 0; sep && ax < 3; ++ax) {
            const int na = ax == 0 ? nx : (ax == 1 ? ny : nz);
            const int stride = ax == 0 ? ny * nz : (ax == 1 ? nz : 1);
            for (int t = 1; sep && t < na; ++t) {
                const float gap = mu[3 * (t * stride) + ax] - mu[3 * ((t - 1) * stride) + ax];
                sep = gap > 0.f && gap <= 10.0f * sigma[ax];
            }
            if (sep) {
                gm->guard_lo[ax] = mu[ax] - 5.0f * sigma[ax];
                gm->guard_hi[ax] = mu[3 * ((na - 1) * stride) + ax] + 5.0f * sigma[ax];
            }
        }
        gm->separable = sep ? 1 : 0;
        if (sep) {
            gm->res[0] = nx; gm->res[1] = ny; gm->res[2] = nz;
            gm->w_uniform = w[0];
            for (int k = 0; k < 3; ++k) gm->axis_isig[k] = 1.0f / sigma[k];
        }
    }
    auto upload = [&](float4** dst, const std::vector<float4>& src) -> int {
        MUPS_CUDA_TRY(cudaMalloc((void**)dst, sizeof(float4) * src.size()));
        MUPS_CUDA_TRY(cudaMemcpy(*dst, src.data(), sizeof(float4) * src.size(), cudaMemcpyHostToDevice));
        return MUPS_OK;
    };
    int rc;
    if ((rc = upload(&gm->A, A)) || (rc = upload(&gm->Bv, Bv)) || (rc = upload(&gm->C, C))) return fail(rc);
    if (gm->separable) {
        float axis[3 * 64] = {0};
        const int nx = gm->res[0], ny = gm->res[1], nz = gm->res[2];
        for (int i = 0; i < nx; ++i) axis[i] = mu[3 * (i * ny * nz)];
        for (int j = 0; j < ny; ++j) axis[64 + j] = mu[3 * (j * nz) + 1];
        for (int k = 0; k < nz; ++k) axis[128 + k] = mu[3 * k + 2];
        cudaError_t e = cudaMalloc((void**)&gm->axis_mu, sizeof(axis));
        if (e == cudaSuccess) e = cudaMemcpy(gm->axis_mu, axis, sizeof(axis), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("mups_gmm_create: %s", cudaGetErrorString(e)); return fail(MUPS_ERR_CUDA); }
    }
    *out = gm;
    return MUPS_OK;
}

int mups_gmm_size(const mups_gmm* g) { return g ? g->G : 0; }
int mups_gmm_is_separable(const mups_gmm* g) { return g ? g->separable : 0; }

// ---- half 2 -------------------------------------------------------------------------------------
static int stats_common(const char* what, const mups_gmm* gmm, const float* patches_dev, const int32_t* n_eff_dev, int64_t B,
                        int S, int P, uint32_t flags, float* out_dev, const GatherSource* src, cudaStream_t st) {
    MUPS_REQUIRE(gmm != nullptr, "%s: gmm is NULL", what);
    MUPS_REQUIRE(B >= 0, "%s: B=%lld is negative", what, (long long)B);
    MUPS_REQUIRE(S >= 1 && S <= MUPS_MAX_SCALES, "%s: S=%d out of range [1, %d]", what, S, MUPS_MAX_SCALES);
    MUPS_REQUIRE(P >= 1 && P <= MUPS_MAX_POINTS_PER_PATCH, "%s: P=%d out of range [1, %d]", what, P, MUPS_MAX_POINTS_PER_PATCH);
    MUPS_REQUIRE(B == 0 || ((patches_dev || src) && out_dev), "%s: patches / out is NULL", what);
    MUPS_REQUIRE(!(flags & MUPS_FLAG_MASKED) || B == 0 || n_eff_dev, "%s: n_eff is required with MUPS_FLAG_MASKED "
                 "(the reference fails on n_original_points=None, tf_util.py:665)", what);
    MUPS_REQUIRE((flags & ~(MUPS_FLAG_MASKED | MUPS_LAYOUT_CHANNEL | MUPS_FLAG_NO_FASTPATH | MUPS_FLAG_WIDE_STORES)) == 0, "%s: unknown flags 0x%x", what, flags);
    if (int rc = check_device(gmm->device, what)) return rc;
    // stream-ordered scratch for the fast path's fallback worklist (1 counter + one id per (query, scale))
    int* work = nullptr;
    if (gmm->separable && !(flags & MUPS_FLAG_NO_FASTPATH) && B > 0)
    {
        cudaMemPool_t pool = nullptr;
        if (int prc = library_pool(gmm->device, &pool)) return prc;
        MUPS_CUDA_TRY(cudaMallocFromPoolAsync((void**)&work, sizeof(int) * (size_t)(B * S + 1), pool, st));
    }
    const int rc = launch_3dmfv(gmm, patches_dev, n_eff_dev, B, S, P, flags, out_dev, work, src, st);
    if (work) cudaFreeAsync(work, st);
    return rc;
}

int mups_3dmfv(const mups_gmm* gmm, const float* patches_dev, const int32_t* n_eff_dev, int64_t B, int S, int P,
               uint32_t flags, float* out_dev, mups_stream stream) {
    return stats_common("mups_3dmfv", gmm, patches_dev, n_eff_dev, B, S, P, flags, out_dev, nullptr, static_cast<cudaStream_t>(stream));
}

// ---- K6: selection hand-off (patches never in HBM) ------------------------------------------------------------------
static int query_common(const char* what, const mups_index* ix, const int64_t* query_idx_dev, int64_t B, const double* r_abs_host,
                        int S, int P, uint64_t seed, int32_t* nbr_idx_dev, int32_t* nbr_total_dev, float* patches_dev,
                        int32_t* n_eff_dev, int32_t* nbr_pos_dev, cudaStream_t st) {
    MUPS_REQUIRE(ix != nullptr, "%s: index is NULL", what);
    MUPS_REQUIRE(B >= 0 && B <= 0x7FFFFFFFll, "%s: B=%lld out of range", what, (long long)B);
    MUPS_REQUIRE(S >= 1 && S <= MUPS_MAX_SCALES, "%s: S=%d out of range [1, %d]", what, S, MUPS_MAX_SCALES);
    MUPS_REQUIRE(P >= 1 && P <= MUPS_MAX_POINTS_PER_PATCH, "%s: P=%d out of range [1, %d]", what, P, MUPS_MAX_POINTS_PER_PATCH);
    MUPS_REQUIRE(r_abs_host != nullptr, "%s: radii are NULL", what);
    for (int s = 0; s < S; ++s)
        MUPS_REQUIRE(std::isfinite(r_abs_host[s]) && r_abs_host[s] >= 0.0, "%s: radius %d is %g", what, s, r_abs_host[s]);
    MUPS_REQUIRE(B == 0 || (query_idx_dev && n_eff_dev), "%s: query_idx / n_eff is NULL", what);
    if (int rc = check_device(ix->device, what)) return rc;
    if (st != ix->build_stream) MUPS_CUDA_TRY(cudaStreamWaitEvent(st, ix->built, 0));
    if (int rc = launch_ball_query(ix, query_idx_dev, B, r_abs_host, S, P, seed, nbr_idx_dev, nbr_total_dev, patches_dev,
                                   n_eff_dev, nbr_pos_dev, st))
        return rc;
    return note_use(ix, st);
}

int mups_ball_query_select(const mups_index* index, const int64_t* query_idx_dev, int64_t B, const double* r_abs_host, int S,
                           int P, uint64_t seed, int32_t* nbr_pos_dev, int32_t* nbr_total_dev, int32_t* n_eff_dev,
                           mups_stream stream) {
    MUPS_REQUIRE(B == 0 || nbr_pos_dev, "mups_ball_query_select: nbr_pos is NULL");
    return query_common("mups_ball_query_select", index, query_idx_dev, B, r_abs_host, S, P, seed, nullptr, nbr_total_dev, nullptr,
                        n_eff_dev, nbr_pos_dev, static_cast<cudaStream_t>(stream));
}

int mups_3dmfv_selected(const mups_gmm* gmm, const mups_index* index, const int64_t* query_idx_dev, int64_t B,
                        const double* r_abs_host, int S, int P, const int32_t* nbr_pos_dev, const int32_t* n_eff_dev,
                        uint32_t flags, float* out_dev, mups_stream stream) {
    MUPS_REQUIRE(index != nullptr, "mups_3dmfv_selected: index is NULL");
    MUPS_REQUIRE(r_abs_host != nullptr, "mups_3dmfv_selected: radii are NULL");
    MUPS_REQUIRE(B == 0 || (query_idx_dev && nbr_pos_dev), "mups_3dmfv_selected: query_idx / nbr_pos is NULL");
    if (int rc = check_device(index->device, "mups_3dmfv_selected")) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (st != index->build_stream) MUPS_CUDA_TRY(cudaStreamWaitEvent(st, index->built, 0));
    GatherSource src{index, query_idx_dev, nbr_pos_dev, r_abs_host};
    if (int rc = stats_common("mups_3dmfv_selected", gmm, nullptr, n_eff_dev, B, S, P, flags | MUPS_FLAG_MASKED, out_dev, &src, st))
        return rc;
    return note_use(index, st);
}

int mups_features(const mups_index* index, const mups_gmm* gmm, const int64_t* query_idx_dev, int64_t B,
                  const double* r_abs_host, int S, int P, uint64_t seed, uint32_t flags, float* patches_dev,
                  int32_t* n_eff_dev, int32_t* nbr_total_dev, float* out_dev, mups_stream stream) {
    if (patches_dev) {
        int rc = mups_ball_query(index, query_idx_dev, B, r_abs_host, S, P, seed, nullptr, nbr_total_dev, patches_dev,
                                 n_eff_dev, stream);
        if (rc) return rc;
        return mups_3dmfv(gmm, patches_dev, n_eff_dev, B, S, P, flags | MUPS_FLAG_MASKED, out_dev, stream);
    }
    // K6: no patch tensor.  The ball query hands over the positions of the selected neighbours (4 bytes per slot, library
    // scratch); the statistics kernel gathers, centres and normalises them while it stages a patch.
    MUPS_REQUIRE(index != nullptr, "mups_features: index is NULL");
    if (B <= 0 || S < 1 || S > MUPS_MAX_SCALES || P < 1 || P > MUPS_MAX_POINTS_PER_PATCH)       // let the callee report it
        return mups_ball_query_select(index, query_idx_dev, B, r_abs_host, S, P, seed, nullptr, nbr_total_dev, n_eff_dev, stream);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemPool_t pool = nullptr;
    if (int rc = library_pool(index->device, &pool)) return rc;
    int32_t* pos = nullptr;
    MUPS_CUDA_TRY(cudaMallocFromPoolAsync((void**)&pos, sizeof(int32_t) * (size_t)B * S * P, pool, st));
    int rc = mups_ball_query_select(index, query_idx_dev, B, r_abs_host, S, P, seed, pos, nbr_total_dev, n_eff_dev, stream);
    if (!rc) rc = mups_3dmfv_selected(gmm, index, query_idx_dev, B, r_abs_host, S, P, pos, n_eff_dev, flags, out_dev, stream);
    cudaFreeAsync(pos, st);
    return rc;
}

}  // extern "C"
