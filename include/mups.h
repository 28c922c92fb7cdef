/* mups.h -- C ABI of the B200-native MuPS (Multi-scale Point Statistics) hot path.
 *
 * Drop-in boundary for the one data-parallel path of sitzikbs/Nesti-Net this
 * repository accelerates.  The reference has no FFI layer
 * of its own (pure Python + TF1 graph), so each entry point cites the Python
 * interface it replaces (paths relative to the reference root):
 *
 *   mups_index_*     <- utils/pcpnet_dataset.py:13-39   load_shape -> spatial.cKDTree(pts, 10)
 *                       utils/pcpnet_dataset.py:281     bbdiag from pts.max(0) - pts.min(0)
 *   mups_ball_query  <- utils/pcpnet_dataset.py:286-343 PointcloudPatchDataset.__getitem__
 *                       (query_ball_point :304, subsample :320-321, centre :335-336, /rad :343)
 *   mups_gmm_*       <- utils/utils.py:70-95 get_3d_grid_gmm as fed by
 *                       train_n_est_w_experts.py:284-286 (w, mu, sqrt(cov) as float32)
 *   mups_3dmfv       <- utils/tf_util.py:655-753 get_3dmfv_n_est (MUPS_FLAG_MASKED)
 *                       utils/tf_util.py:578-652 get_3dmfv        (no MUPS_FLAG_MASKED)
 *                       models/experts_n_est.py:59-76 MuPS assembly (MUPS_LAYOUT_MUPS)
 *   mups_features    <- the two halves back to back (what one DataLoader item + one
 *                       sess.run of the MuPS sub-graph compute, test_n_est_w_experts.py:129-148)
 *
 * Conventions
 *   - plain C: pointers and sizes only; "dev" pointers are CUDA device pointers on the
 *     device that is current when the call is made, "host" pointers are ordinary memory.
 *   - every function returns MUPS_OK (0) or a negative mups_status; nothing throws or aborts
 *     across the ABI.  mups_last_error() returns a thread-local description of the last failure.
 *   - work is enqueued on the CUDA stream passed in (a cudaStream_t cast to void*; NULL = legacy
 *     default stream).  No call synchronises unless its comment says so.
 *   - an index / gmm handle is immutable after creation: concurrent queries from several host
 *     threads or streams are legal.  One handle belongs to the device it was created on.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with MUPS_ERR_CUDA.
 */
#ifndef MUPS_H_
#define MUPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUPS_ABI_VERSION 1

typedef enum mups_status {
    MUPS_OK = 0,
    MUPS_ERR_INVALID = -1,   /* bad argument (the reference raises ValueError) */
    MUPS_ERR_CUDA = -2,      /* CUDA runtime error, no device, launch failure  */
    MUPS_ERR_NOMEM = -3,     /* allocation failed                              */
    MUPS_ERR_UNSUPPORTED = -4/* shape outside the compiled limits              */
} mups_status;

/* flags of mups_3dmfv / mups_features */
#define MUPS_FLAG_MASKED 1u        /* get_3dmfv_n_est semantics (n_eff mask, /n_eff); else get_3dmfv */
#define MUPS_LAYOUT_MUPS 0u        /* out[b][g][s*20+c]  == [B,res,res,res,20*S] (experts_n_est.py:71-76) */
#define MUPS_LAYOUT_CHANNEL 2u     /* out[b][s][c][g]    == per scale [B,20*G] flatten=True / [B,20,G] */
#define MUPS_FLAG_NO_FASTPATH 4u   /* force the general (non-separable) statistics kernel */
#define MUPS_FLAG_WIDE_STORES 8u   /* lattice kernel, MUPS layout: transpose a (query, scale) result through shared memory
                                      and write it as 80-byte runs (5 lanes per Gaussian) instead of 16-byte pieces at a
                                      320-byte stride -- for `out_dev` in a PEER GPU's memory (NVLink write efficiency) */

/* limits compiled into the kernels */
#define MUPS_MAX_SCALES 8
#define MUPS_MAX_POINTS_PER_PATCH 2048

typedef struct mups_index mups_index;   /* uniform-grid spatial hash of one cloud (replaces cKDTree) */
typedef struct mups_gmm mups_gmm;       /* device copy of (w, mu, sigma) + derived constants       */
typedef void* mups_stream;              /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------------- */
int mups_abi_version(void);
const char* mups_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (monotonic). */
int64_t mups_launch_count(void);
/* Process-wide tuning knobs.  "boundary_cap" (1..512, default 512): largest radix threshold group
 * the subsample resolves without another refinement level (tests lower it to exercise the
 * refinement path, which otherwise needs > ~500k neighbours in one ball).
 * "fuse_candidates" (default 12288): candidate points in a ball's first cell batch above which the
 * first scan also builds the subsample's key histograms (dense balls; results never depend on it).
 * "stats_variant" (0 = automatic; 1 round-1 loop and staging, 2 all-scalar loop, 3 two (query, scale) items per CTA with the
 * patches delivered by cp.async.bulk + mbarrier, 8 no cluster at 16^3):
 * force a statistics-kernel variant (benchmarking only).
 * "query_kernel" (0 = automatic: hierarchical for indices with >= 6 Morton bits per axis, i.e. fine grids; 1 = flat
 * cell scan; 2 = hierarchical), "query_order" (0 = automatic; 1 = CTAs in caller order; 2 = CTAs in Morton order of the
 * centres) and "hier_margin" (-1 = automatic; extra expected candidates above P kept by the hierarchical kernel's key
 * threshold; tests set 0 to exercise the hand-over to the flat kernel): results never depend on any of them.
 * "pool_variant" (0 = automatic: shared-memory tile + separable box sum for the 8^3 average pools; 1 = the per-voxel
 * kernel everywhere) and "conv_variant" (0 = automatic; 1 = one CTA per SM with the deepest pipeline also for the short-K 1^3
 * layers; 2 = the per-tap kernel also for the 8^3 layers; 3 = z-halo kernel without the operand swap; 4 = half-sample CTAs
 * at every batch size; 5 = no channel-tile splitting for nearly empty grids; 7 = two voxel tiles per 256-channel tile in the 1^3
 * layers): benchmarking only. */
int mups_set_option(const char* name, int64_t value);

/* ---- spatial index (K1 bbox + K2 grid build) --------------------------------------------- */
/* Builds the index of xyz_dev [n,3] float32 (row-major, device).  cell_frac: cell edge as a
 * fraction of the bounding-box diagonal -- pass the largest patch radius that will be queried
 * (<= 0 selects 0.07); any radius is still answered correctly, larger ones scan more cells.
 * Asynchronous on `stream`; the handle may be used on the same stream immediately.
 * The point data is copied: xyz_dev may be freed once the stream has passed the call. */
int mups_index_create(mups_index** out, const float* xyz_dev, int64_t n, double cell_frac, mups_stream stream);
/* Bounding box (float32 min/max per axis, what pts.min(0)/pts.max(0) return).  Synchronises the
 * build stream on first use. */
int mups_index_bbox(const mups_index* index, float min3_host[3], float max3_host[3]);
int64_t mups_index_size(const mups_index* index);
void mups_index_destroy(mups_index* index);

/* ---- half 1: multi-radius ball query + seeded subsample + centre + normalise (K3 + K4) ---- */
/* For each of the B centre points query_idx_dev[b] (indices into the indexed cloud) and each of
 * the S radii r_abs_host[s] (absolute, float64, exactly the reference's bbdiag*patch_radius[s]):
 *   nbr_total[b][s] = |{ j : sum_k (double(p_jk) - double(p_ck))^2 <= r*r }|  (cKDTree predicate)
 *   n_eff[b][s]     = min(P, nbr_total)
 *   the shared seeded selection when nbr_total > P, else all neighbours, in ascending index order:
 *   (a, b) = Philox4x32-10(counter=(centre, s, 0, 0), key=seed)[0..1], key(j) = (fmix32(j) ^ a) * (b | 1) mod 2^32;
 *   the P smallest (key, j) pairs are kept (oracle/mups_oracle.py::select_subset)
 *   nbr_idx[b][s][t]      = selected point index, -1 beyond n_eff           (may be NULL)
 *   patches[b][s*P+t][k]  = (p_jk - p_ck) / float(r_s) in IEEE fp32, 0 beyond n_eff (may be NULL)
 */
int mups_ball_query(const mups_index* index, const int64_t* query_idx_dev, int64_t B,
                    const double* r_abs_host, int S, int P, uint64_t seed,
                    int32_t* nbr_idx_dev, int32_t* nbr_total_dev, float* patches_dev, int32_t* n_eff_dev,
                    mups_stream stream);

/* ---- GMM ----------------------------------------------------------------------------------- */
/* w [G], mu [G,3], sigma [G,3] (std-dev) float32 on the HOST.  Detects the separable lattice
 * (tensor-product means, one sigma per axis, uniform w) produced by get_3d_grid_gmm.
 * w <= 0, sigma <= 0 or non-finite parameters are rejected with MUPS_ERR_INVALID (the reference would propagate
 * NaN / Inf through sqrt(w) and 1/sigma; get_3d_grid_gmm cannot produce them). */
int mups_gmm_create(mups_gmm** out, const float* w_host, const float* mu_host, const float* sigma_host, int G);
int mups_gmm_size(const mups_gmm* gmm);
int mups_gmm_is_separable(const mups_gmm* gmm);
void mups_gmm_destroy(mups_gmm* gmm);

/* ---- half 2: 3DmFV statistics (K5) ---------------------------------------------------------- */
/* patches_dev [B, S*P, 3] float32; n_eff_dev [B,S] int32 (required with MUPS_FLAG_MASKED, else
 * ignored and may be NULL); out_dev: B*S*20*G float32 in the layout selected by `flags`.
 * Rows with n_eff == 0 are last-batch padding (test_n_est_w_experts.py:134-140): the reference divides by zero there
 * and so does this kernel (NaN / Inf rows, to be dropped by the caller as the reference's loop does).  A NEGATIVE n_eff
 * cannot come out of the reference's dataset (effective_points_num is min(P, len) >= 1); it is treated as 0, where the
 * reference would mask nothing and divide by the negative count. */
int mups_3dmfv(const mups_gmm* gmm, const float* patches_dev, const int32_t* n_eff_dev,
               int64_t B, int S, int P, uint32_t flags, float* out_dev, mups_stream stream);

/* ---- K6: selection hand-off, patches never in HBM ------------------------------------------------ */
/* mups_ball_query without the patch tensor: nbr_pos[b][s][t] = position of the t-th selected neighbour (ascending point
 * index) in the index's Morton-ordered point array, -1 beyond n_eff -- 4 bytes per slot instead of 12.  Opaque to the
 * caller; only mups_3dmfv_selected with the SAME index, query list and radii can consume it. */
int mups_ball_query_select(const mups_index* index, const int64_t* query_idx_dev, int64_t B,
                           const double* r_abs_host, int S, int P, uint64_t seed,
                           int32_t* nbr_pos_dev, int32_t* nbr_total_dev, int32_t* n_eff_dev, mups_stream stream);
/* get_3dmfv_n_est + MuPS assembly of the patches a selection describes: the statistics kernel gathers the selected
 * points from the index while it stages a patch and centres / normalises them itself ((p - c) / float32(r) in IEEE
 * fp32: K4 folded into K5).  Bit-identical to mups_3dmfv on the patch tensor mups_ball_query would have written.
 * MUPS_FLAG_MASKED is implied. */
int mups_3dmfv_selected(const mups_gmm* gmm, const mups_index* index, const int64_t* query_idx_dev, int64_t B,
                        const double* r_abs_host, int S, int P, const int32_t* nbr_pos_dev, const int32_t* n_eff_dev,
                        uint32_t flags, float* out_dev, mups_stream stream);

/* ---- both halves ---------------------------------------------------------------------------- */
/* mups_ball_query followed by mups_3dmfv(MUPS_FLAG_MASKED) for the same B centres.
 * patches_dev / n_eff_dev are caller-provided scratch of the sizes above (kept so that the
 * caller can also read the patches); nbr_total_dev may be NULL.
 * patches_dev == NULL selects the K6 path: mups_ball_query_select + mups_3dmfv_selected through stream-ordered library
 * scratch (4 B*S*P bytes) -- same features bit for bit, a third of the intermediate HBM traffic, no patch tensor. */
int mups_features(const mups_index* index, const mups_gmm* gmm, const int64_t* query_idx_dev, int64_t B,
                  const double* r_abs_host, int S, int P, uint64_t seed, uint32_t flags,
                  float* patches_dev, int32_t* n_eff_dev, int32_t* nbr_total_dev, float* out_dev,
                  mups_stream stream);

/* ---- the consumer on the tensor cores (replaces the TF layers of models/experts_n_est.py:155-314) ---- */
/* MuPS fp32 [rows, 20 S] (rows = B * res^3, as mups_3dmfv wrote it) -> bf16 [rows, 32 S]: each scale's 20 channels padded
 * to 32, the activation layout of the kernels below (channel slices of the single-scale experts 16-byte aligned). */
int mups_moe_pack_input(const float* mups_dev, int64_t rows, int S, void* out_bf16_dev, mups_stream stream);
/* tf_util.avg_pool3d (utils/tf_util.py:432-455; window k, stride 1, 'SAME': mean over the valid cells, is_max = 0) and
 * tf_util.max_pool3d (:406-430; window 2, stride 2, is_max = 1) on channels [c_off, c_off + c) of x_bf16_dev NDHWC bf16
 * [B, D, D, D, c_total]; y_bf16_dev: [B, D', D', D', c] contiguous (D' = D, or D / 2 for the max pool). */
int mups_pool3d(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int c, int k, int is_max, void* y_bf16_dev,
                mups_stream stream);
/* tf_util.conv3d (utils/tf_util.py:254-311: stride 1, 'SAME', bias, batch norm, ReLU) and tf_util.fully_connected
 * (:314-351; D = 1, k = 1) as one tcgen05 implicit GEMM:
 *   y[b,z,y,x,co] = act(scale[co] * sum_{dz,dy,dx,ci} x[b, z+dz-p, y+dy-p, x+dx-p, cin_off+ci] * w[(dz*k+dy)*k+dx][co][ci] + shift[co])
 * x_bf16_dev: NDHWC bf16 [B, D, D, D, cin_total] (D in {1, 2, 4, 8}); channels [cin_off, cin_off + cin) are read, zero outside the
 * volume with p = (k - 1) / 2 cells before (TF 'SAME': the smaller half first).  w_bf16_dev: [k^3][cout][cin_w] bf16.
 * scale / shift: [cout] fp32 (bias and the batch norm's moving statistics folded by the caller); relu: 0 / 1.
 * Output: bf16 into channels [cout_off, cout_off + cout) of y_bf16_dev [B, D, D, D, cout_total] (NULL to skip) and / or
 * fp32 [B*D^3, cout] into y_f32_dev (NULL to skip).  Channel counts: cin, cin_total, cin_off, cin_w, cout_total, cout_off
 * multiples of 8, cout a multiple of 16.  bf16 products, fp32 accumulation (tensor memory). */
int mups_conv3d_bn_relu(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin,
                        const void* w_bf16_dev, int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev,
                        int relu, void* y_bf16_dev, int cout_total, int cout_off, float* y_f32_dev, mups_stream stream);
/* The two 1^3 convolutions of an inception module that read the module's input -- `one` (utils/tf_util.py conv3d, scope
 * '_conv1') and the convolution after the average pool ('_conv4', models/experts_n_est.py:291-310) -- from ONE read of it.
 * A 1^3 convolution without bias commutes with the average pool (both are linear, one over channels, one over voxels), so
 * conv4(avgpool(x)) = avgpool(conv4_nobias(x)) + bias: the pool then runs on n_filters instead of C_in channels.
 * w_bf16_dev: [1][cout][cin_w], rows [0, split) = conv1, rows [split, cout) = conv4; scale / shift per output channel (identity
 * for the rows whose batch norm is applied after the pool); ReLU on output channels below relu_upto.  Channels [0, split) go to
 * channels [y1_off, ...) of y1_bf16_dev [.., y1_total], channels [split, cout) to [y2_off, ...) of y2_bf16_dev [.., y2_total]. */
int mups_conv1_split_bn_relu(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                             int cin_w, int cout, const float* scale_dev, const float* shift_dev, int relu_upto, void* y1_bf16_dev,
                             int y1_total, int y1_off, int split, void* y2_bf16_dev, int y2_total, int y2_off, mups_stream stream);
/* tf_util.avg_pool3d (window k >= 2, stride 1, 'SAME', mean over the valid cells) of channels [c_off, c_off + c) of x_bf16_dev
 * followed by y = act(scale[ch] * pooled + shift[ch]) into channels [y_off, y_off + c) of y_bf16_dev [B, D, D, D, y_total]: the
 * second half of the pool branch above. */
int mups_avgpool3d_bn_relu(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int c, int k, const float* scale_dev,
                           const float* shift_dev, int relu, void* y_bf16_dev, int y_total, int y_off, mups_stream stream);

/* ---- bf16x3 ("split") mode of the consumer: fp32-grade results from the same tensor-core kernels ------------------------
 * A value is carried as hi = bf16(v), lo = bf16(v - hi); a tensor of w logical channels is stored as the TRIPLET
 * [hi | lo | hi] (3 w channels), the weights are expanded to [w_hi | w_hi | w_lo], and the convolution over the 3 x longer
 * channel axis accumulates a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32.  The entry points below are that convolution with
 * triplets out, the conversion of an fp32 tensor (the MuPS input) into triplets, and the pools between triplets; they replace
 * nothing in the reference (which computes in fp32 throughout, models/experts_n_est.py:155-314) -- they are what lets the
 * bf16 tensor cores reproduce it to ~1e-5 relative (normals within 0.001 degrees). */
/* mups_conv3d_bn_relu with its output written as triplets straight from the epilogue (no fp32 round trip): output channels
 * [0, split) form the triplet [hi | lo | hi] of part width `split` at channels [cout_off, cout_off + 3 split) of y_bf16_dev, channels
 * [split, cout) a second triplet of part width cout - split right behind it (split == cout: one triplet; split a multiple of 16).
 * Bit-identical to the fp32 output followed by mups_split_bf16x3. */
int mups_conv3d_bn_relu_x3(const void* x_bf16_dev, int64_t B, int D, int cin_total, int cin_off, int cin, const void* w_bf16_dev,
                           int cin_w, int cout, int k, const float* scale_dev, const float* shift_dev, int relu, void* y_bf16_dev,
                           int cout_total, int cout_off, int split, mups_stream stream);
/* fp32 src_dev [rows, src_stride], columns [src_off, src_off + w_src) -> triplet at channels [dst_off, dst_off + 3 w_dst) of
 * dst_bf16_dev [rows, dst_stride] (w_dst >= w_src, a multiple of 8; channels [w_src, w_dst) of every part are zero). */
int mups_split_bf16x3(const float* src_dev, int64_t rows, int src_stride, int src_off, int w_src, void* dst_bf16_dev, int dst_stride,
                      int dst_off, int w_dst, mups_stream stream);
/* tf_util.avg_pool3d / max_pool3d (as mups_pool3d) from the triplet at channels [c_off, c_off + 3 w) of x_bf16_dev
 * [B, D, D, D, c_total] to the triplet at [y_off, y_off + 3 w) of y_bf16_dev [B, D', D', D', y_total]; the window is
 * reduced on hi + lo in fp32. */
int mups_pool3d_bf16x3(const void* x_bf16_dev, int64_t B, int D, int c_total, int c_off, int w, int k, int is_max, void* y_bf16_dev,
                       int y_total, int y_off, mups_stream stream);
/* The pool branch of an inception module in bf16x3 mode (models/experts_n_est.py:291-310): its 1^3 convolution has run first
 * (raw fp32 output x_f32_dev [B, D, D, D, c], no bias -- it commutes with the pool); this is tf_util.avg_pool3d (window k >= 2,
 * stride 1, 'SAME', mean over the valid cells) on that tensor followed by act(scale[ch] * pooled + shift[ch]) (the folded bias +
 * batch norm + ReLU) written as the triplet at channels [y_off, y_off + 3 c) of y_bf16_dev [B, D, D, D, y_total]. */
int mups_avgpool3d_f32_bn_relu_x3(const float* x_f32_dev, int64_t B, int D, int c, int k, const float* scale_dev, const float* shift_dev,
                                  int relu, void* y_bf16_dev, int y_total, int y_off, mups_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MUPS_H_ */
