/*
 * gpet_b200.h - C ABI of libgpet_b200.so: the sm_100a (B200) implementation of the data-parallel hot
 * path of jaburke166/gaussian_process_edge_trace.
 *
 * The reference has no FFI layer (pure Python); each entry point below replaces one seam of
 * /root/reference/gp_edge_tracing (file:line cited per function) and is what a ctypes binding in the
 * reference would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host, nothing allocates:
 *     scratch is passed in by the caller (sizes from the *_workspace_bytes queries);
 *   - a "batch" is B independent traces that share image shape (M rows, N columns), edge span
 *     x_st..x_st+n-1, sample count S and the GP kernel; per-trace data is indexed by b;
 *   - posterior curves are stored as Y[b][j][s] (float64, s contiguous) = the reference's
 *     y_samples[n, N_samples] (gpet.py:260-261) per trace;
 *   - return value: 0 = OK, otherwise a GPET_ERR_* code; gpet_last_error() gives the text.
 */
#ifndef GPET_B200_H
#define GPET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPET_OK 0
#define GPET_ERR_INVALID 1     /* bad argument */
#define GPET_ERR_CUDA 2        /* a CUDA call failed */
#define GPET_ERR_UNSUPPORTED 3 /* shape outside what the kernels are built for */

#define GPET_MAX_TRAIN 224     /* max training points m of the shared-memory posterior / final-fit kernels (packed triangle); larger
                                  training sets take the HBM-resident blocked path (gpet_dense.cu, "_big" entry points) */
#define GPET_MAX_RANK 160      /* max padded rank rp of the low-rank factor path */

const char* gpet_last_error(void);
#define GPET_ABI_VERSION 8
int gpet_abi_version(void);   /* == GPET_ABI_VERSION of the header the library was built from */

/* launch-shape / variant knobs (defaults = measured best on B200; used by the tuning benchmarks) */
#define GPET_TUNE_SCORE_THREADS 0   /* 128 | 256 | 512 threads per CTA of the scoring kernel */
#define GPET_TUNE_SCORE_SCAN 1      /* 1: Simpson abscissa = running sum of segment lengths (reference); 0: h = segment */
#define GPET_TUNE_EIG_THREADS 2     /* 0: Householder + QL eigensolver (default); else threads per CTA of the
                                       parallel Jacobi eigensolver (256..1024) */
#define GPET_TUNE_LML_THREADS 3     /* 0: blocked LML objective kernel (default); else threads per CTA of the
                                       column-at-a-time kernel (multiple of 32, <= 1024) */
#define GPET_TUNE_SCORE_STAGES 4    /* ring depth of the bulk-copy staged scoring kernel (4 or 8); 0: register-prefetch kernel */
#define GPET_TUNE_SCORE_MINBLOCKS 5 /* register cap of the staged scoring kernel as CTAs/SM: 4, 5 (80 regs), 6 (64) */
#define GPET_TUNE_SAMPLE_ROWS 6     /* 1: sampling GEMM with the A tile resident per CTA and cp.async-staged Z tiles */
#define GPET_TUNE_SCORE_CPT 7       /* curves per consumer thread of the staged scoring kernel: 0 = by launch size (default), 1, 2 or 3 */
#define GPET_TUNE_LBFGSB_THREADS 8  /* L-BFGS-B advance kernel: 0 = one run per warp, state of run e contiguous; 32 | 64 | 128 = one run per
                                       thread with that CTA size, state interleaved with stride E (must not change during a fit) */
#define GPET_TUNE_POSTERIOR_PACKED 9 /* 1: always the packed-triangle posterior kernels (default 0: only when the sizes need them) */
#define GPET_TUNE_JACOBI_BLOCK 10    /* block Jacobi eigensolver of the full covariance: 32 (64 x 64 pivots, default) or 64 (128 x 128) */
#define GPET_TUNE_JACOBI_PIVOT 11    /* solver of the 64 x 64 pivots: threads per CTA of the parallel Jacobi kernel (default 512), 0 = Householder + QL */
#define GPET_TUNE_JACOBI_INNER 12    /* 0: pivots diagonalised to rounding level; k > 0: at most k inner sweeps per pivot (inexact block Jacobi; default 2) */
#define GPET_TUNE_JACOBI_SYM 13      /* 1 (default): A <- J^T A J as one fused kernel on the lower half (blocks of 32); 0: column pass + row pass */
#define GPET_TUNE_COUNT 14
int gpet_set_tuning(int knob, int value);

/* ---- gpet_utils.comp_grad_img (gpet_utils.py:95-119) + normalise (:65-91) -------------------------
 * img[B][M][N] f64 -> out[B][M][N] f32 in [0,1].  taps[kh][kw] f64 = the kernel as passed to comp_grad_img
 * (true convolution: flipped inside, edge-replicated border, fp64 accumulation in scipy.ndimage's tap order
 * without FMA => the pre-normalisation map is bit-identical to scipy).  minmax[B][2] u32 scratch. */
int gpet_comp_grad_img_f64(const double* img, int B, int M, int N, const double* taps, int kh, int kw,
                           float* out, uint32_t* minmax, void* stream);
/* opt-in fast form: the same map accumulated with float32 fused multiply-adds from a float32 tile (within ~1e-6 relative
 * of the exact one; the stencil's bar is 1e-4) - HBM bound (12 B per pixel) instead of FP64-pipe bound. */
int gpet_comp_grad_img_fast_f32(const double* img, int B, int M, int N, const double* taps, int kh, int kw,
                                float* out, uint32_t* minmax, void* stream);

/* normalise(img, (0,1), float32) in place (gpet_utils.py:81-91) for a float32 map: a -= min; a /= max. */
int gpet_normalise_f32(float* img, int B, int M, int N, uint32_t* minmax, void* stream);

/* ---- GP_Edge_Tracing.__init__: grad_kde = kernel_density_estimate(None, None) (gpet.py:127, 503-529) --
 * grad[B][M][N] f32 (normalised gradient) -> grad_kde[B][M][N] f32.  work: B*M*N f64 + 4096 B. */
int64_t gpet_grad_kde_workspace_bytes(int B, int M, int N);
int gpet_grad_kde_f32(const float* grad, int B, int M, int N, float* grad_kde, void* work, void* stream);

/* transpose to the column-major layout used by the scoring gather, one guard entry at each end of a column:
 * gradT[b][x][r+1] = src[b][clamp(r,0,M-1)][x] for r = -1..M, i.e. dst is f32[B][N][M+2] */
int gpet_transpose_f32(const float* src, int B, int M, int N, float* dst, void* stream);

/* ---- fit_predict_GP(converged=False): GaussianProcessRegressor.fit + predict (gpet.py:182-268,
 * sklearn_gpr.py:183-321, 379-407), low-rank form of the posterior covariance --------------------------
 * Per trace b: m[b] training points sorted by x: xi[b][mmax] = x - x_st (int32), y[b][mmax] raw rows (f64),
 * w[b][mmax] noise weights.  kd[n] = unit kernel k(d) at integer distance d (f64).  Ur[n][rp], lam[rp]:
 * leading eigenpairs of the unit kernel matrix on the grid (rows zero-padded to rp).
 * Outputs: mean[b][n] (posterior mean in scaled units, sklearn_gpr.py:381-385), ys[b] (= std(y)+1,
 * gpet.py:228), Mr[b][rp][rp] = U_r^T Sigma U_r (reduced posterior covariance), status[b] (0 ok, 1 = Cholesky
 * failed).  Needs rp <= GPET_MAX_RANK.  mmax > GPET_MAX_TRAIN: the training matrices live in `work` (HBM) and are factored
 * in 64 x 64 blocks (gpet_dense.cu: blocked Cholesky / forward substitution / Gram product on the fp64 tensor
 * instruction); the workspace query covers it.  Small batches (roughly mmax <= 160 with rp <= 76) run in
 * one all-in-shared-memory kernel and need no workspace (the query returns 0, work may be NULL); larger ones keep the
 * packed triangle of K in shared memory and solve G = L^-1 U_r[I,:] column by column through `work`.
 * m_cap: an upper bound on the m[b] of THIS call (<= 0: mmax).  The shared-memory working set is laid out for m_cap
 * points, so the early iterations of a trace (a handful of observations) share an SM between several traces; a trace
 * with m[b] > m_cap gets status 2.  The tracing loop passes K + ctrl[3] (gpet_training_sets_f64). */
int64_t gpet_posterior_lowrank_workspace_bytes(int B, int mmax, int rp);
int gpet_posterior_lowrank_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax,
                               int m_cap, int B, int n, const double* sigma_f, double noise_y, double gp_alpha,
                               const double* kd, const double* Ur, const double* lam, int rp,
                               double* mean, double* ys, double* Mr, int32_t* status, void* work, void* stream);

/* Full posterior covariance Sigma[b][n][n] (sklearn_gpr.py:392-407) for the host-SVD parity mode and for
 * full-rank (Matern) kernels.  work: gpet_posterior_full_workspace_bytes.  Same inputs as above; any mmax (beyond
 * GPET_MAX_TRAIN: blocked path in HBM, as above). */
int64_t gpet_posterior_full_workspace_bytes(int B, int mmax, int n);
int gpet_posterior_full_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax,
                            int B, int n, const double* sigma_f, double noise_y, double gp_alpha,
                            const double* kd, double* mean, double* ys, double* cov, int32_t* status,
                            void* work, void* stream);

/* ---- factor of the posterior covariance: numpy multivariate_normal's svd (sklearn_gpr.py:464) --------
 * Batched symmetric eigen-decomposition (parallel cyclic Jacobi, one CTA per matrix) of Mr[b][rp][rp];
 * eigenvalues sorted descending into d[b][rp], eigenvectors as columns of Q[b][rp][rp] (row-major). */
int64_t gpet_sym_eig_workspace_bytes(int B, int rp);
int gpet_sym_eig_f64(double* Mr, int B, int rp, double* d, double* Q, int32_t* sweeps, void* work, void* stream);

/* A[b][k][j] = sign_k * sqrt(max(d_k,0)) * sum_i Q[b][i][k] * Ur[j][i], sign_k chosen so that
 * <Vt[k], w> > 0 with w_j = 1 + j/n (canonical sign rule); uw[rp] = Ur^T w. */
int gpet_factor_assemble_f64(const double* d, const double* Q, const double* Ur, const double* uw, int B,
                             int rp, int n, double* A, void* stream);

/* ---- sample_y (sklearn_gpr.py:440-473) * y_s (gpet.py:261) ----------------------------------------------
 * Y[b][j][s] = ys[b] * (sum_k Zt[k][s] * A[b][k][j] + mean[b][j]);  Zt[rp][S] = first rp columns of the
 * RandomState(seed).standard_normal((S, n)) draw, transposed; A[b][rp][n].  fp64 DMMA (mma.sync m8n8k4). */
int gpet_sample_f64(const double* Zt, const double* A, const double* mean, const double* ys, int B, int rp,
                    int n, int S, double* Y, void* stream);

/* ---- sample_y + cost_funct fused (sklearn_gpr.py:460-464, gpet.py:261, 437-440, 371-410) ---------------------------------
 * cost[b][s] of the curve Y[b][:][s] = ys[b] (Zt[:, s] . A[b] + mean[b]) without storing Y: a CTA owns 64 curves of one
 * trace, its tensor warps form 32-column chunks of them (fp64 DMMA) into shared memory, its scoring warps consume the
 * chunks with the arithmetic of gpet_score_f64.  Arguments as gpet_sample_f64 + gpet_score_f64.  Supported when
 * gpet_sample_score_supported(rp, n, S) != 0 (rp % 4 == 0, rp <= 80, even n); otherwise GPET_ERR_UNSUPPORTED - use the
 * unfused pair.
 * gpet_sample_keep_f64: the kept curves only, Ykeep[b][j][c] = curve idx[b][c] (c < Kp; idx < 0: zero column), computed
 * with the same accumulation chain (identical bits), in the layout gpet_density_f64 reads with S = Kp. */
int gpet_sample_score_supported(int rp, int n, int S);
int gpet_sample_score_f64(const double* Zt, const double* A, const double* mean, const double* ys, const float* gradT,
                          const int32_t* img_index, int B, int rp, int n, int S, int M, int N, int x_st, double* cost,
                          void* stream);
int gpet_sample_keep_f64(const double* Zt, const double* A, const double* mean, const double* ys, const int32_t* idx,
                         int B, int rp, int n, int S, int Kp, double* Ykeep, void* stream);

/* ---- numpy legacy standard normals on the device (sklearn_gpr.py:460-464 -> RandomState(seed)
 * .standard_normal((S, n)): MT19937 + polar method).  Writes the first kcols grid columns of the samples
 * s0 .. s0+S_loc-1, transposed: Zt[j][s - s0], j < kcols (the layout gpet_sample_f64 consumes).  Same accepted
 * attempts as numpy.  ok[0] = 1 unless the (8 sigma) attempt budget was too small.
 * Values: every operation of the polar method is correctly rounded on both sides except log(r2) - glibc's log errs by
 * < 0.52 ulp.  The device rounds a double-double logarithm correctly and lists the attempts whose logarithm lies within
 * 0.03 ulp of a rounding boundary (~6 %) in `fixups` (device, gpet_standard_normal_fixup_bytes(S, n) bytes: i64 count,
 * i64 0, then with cap = (bytes - 16) / 40 the arrays f64 r2[cap], f64 lg[cap] and three private ones).  The caller
 * reads count and r2[0 .. count), stores libm's own log of them (gpet_host_log_f64, HOST pointers) in lg[0 .. count) and
 * calls gpet_standard_normal_fixup_apply_f64, which rewrites those normals: the result is numpy's array bit for bit
 * (engine.device_standard_normal does exactly this).  fixups == NULL: no
 * list: the logarithm is the correctly rounded one, which is glibc's in all but ~0.05 % of the draws (those normals
 * differ by 1-3 ulp). */
int64_t gpet_standard_normal_workspace_bytes(int64_t S, int n);
int64_t gpet_standard_normal_fixup_bytes(int64_t S, int n);
int gpet_standard_normal_t_f64(uint32_t seed, int64_t S, int n, int kcols, int64_t s0, int64_t S_loc, double* Zt,
                               int32_t* ok, void* fixups, void* work, void* stream);
int gpet_host_log_f64(const double* x, double* out, int64_t count);
int gpet_standard_normal_fixup_apply_f64(int64_t S, int n, int kcols, int64_t s0, int64_t S_loc, double* Zt, void* fixups,
                                         int64_t count, void* stream);

/* ---- get_best_curves / cost_funct (gpet.py:371-451) --------------------------------------------------------
 * cost[b][s] = arc_length / line_integral of curve s over the gradient image (bilinear gather, composite
 * non-uniform Simpson).  gradT[b][N][M+2] f32 guarded column-major copy of the normalised gradient image
 * (gpet_transpose_f32).  img_index (may be NULL): batch item b reads image img_index[b] of gradT instead of
 * image b, so that a caller can pass only the traces that are still active. */
int gpet_score_f64(const double* Y, const float* gradT, const int32_t* img_index, int B, int n, int S, int M, int N,
                   int x_st, double* cost, void* stream);

/* argsort(cost)[:Kp] ascending (gpet.py:443) + KDE weights (1/cost)/sum(1/cost) (gpet.py:492-493).
 * idx[b][Kp] i32, best_cost[b][Kp] f64, wts[b][Kp] f64.  S <= 8192: one shared-memory sort; larger S: radix select
 * of the Kp smallest followed by the sort of those (Kp <= 16384). */
int gpet_topk_f64(const double* cost, int B, int S, int Kp, int32_t* idx, double* best_cost, double* wts,
                  void* stream);

/* ---- kernel_density_estimate(best_curves, costs) (gpet.py:455-529) --------------------------------------
 * Linear binning of the kept curves' points (fixed-point u64 accumulation => order independent), 9x9
 * Gaussian blur, float32 cast and min/max.  dens[b][M][N] f32 is the UN-normalised float32 density; the
 * float32 min-max normalisation (gpet_utils.py:84-85) is applied on the fly by gpet_select_f64 from
 * minmax[b][2].  work: gpet_density_workspace_bytes(B, M, N, Kp). */
int64_t gpet_density_workspace_bytes(int B, int M, int N, int Kp);
int gpet_density_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp,
                     int M, int N, int x_st, float* dens, uint32_t* minmax, void* work, void* stream);
/* The two halves of gpet_density_f64 for sample-sharded runs: every rank splats the kept curves it owns
 * (idx[b][c] = local sample index, or < 0 for a curve of another rank) into work = u64 grid[B][M][N] | f64 scale[B] |
 * i32 dropped[B][Kp]; the ranks all-reduce (sum) grid and dropped - exact integer sums, order independent - and
 * every rank finishes (weight renormalisation, blur, float32 cast, min/max). */
int gpet_density_splat_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp,
                           int M, int N, int x_st, void* work, void* stream);
int gpet_density_finish_f64(const double* wts, int B, int n, int Kp, int M, int N, float* dens, uint32_t* minmax,
                            void* work, void* stream);

/* ---- get_best_pixels / compute_new_obs (gpet.py:532-662), collapsed per-bin form --------------------------
 * For every bin = np.round((x - x_st)/delta_x) (gpet.py:606) the max score 1/3*(kde*gk + kde + gk) (:582) over
 * candidates (kde > 1e-3, :651; old observations first, then row-major pixels; first max wins, :613-616).
 * The host supplies the bin of every image column: col_bin[x] >= 0 for candidate columns, -(bin+1) for columns
 * whose new pixels are excluded (fix_endpoints, :655-657); group_cols[n_groups+1] splits the columns into runs of
 * <= 64 columns (one CTA each) such that no bin straddles two runs.  old_yx[b][max_old][2] i32 (row, col),
 * n_old[b].  Outputs bin_score[b][nb] f64 (-1 when the bin is empty) and bin_pos[b][nb] i32
 * (k < max_old: old observation k; otherwise max_old + y*N + x; -1 when empty). */
int gpet_select_f64(const float* dens, const uint32_t* minmax, const float* grad_kde, const int32_t* img_index, int B,
                    int M, int N, const int32_t* col_bin, const int32_t* group_cols, int n_groups, const int32_t* old_yx,
                    const int32_t* n_old, int max_old, int nb, double* bin_score, int32_t* bin_pos,
                    void* stream);

/* ---- band-limited form of the two stages above (single-rank runs, the default of the tracing loop) -----------------
 * The kept curves cover a narrow band of rows, so the density of one column group (group_cols, as for gpet_select_f64,
 * every group at most max_width <= 56 columns wide) of one trace is built entirely in ONE CTA's shared memory: compact
 * copy of the kept curves, fixed-point histogram of the band, both 9-tap passes in place, float32 cast, min/max.  Only
 * the band rows bands[b][g] = [r_lo, r_hi) of the group's columns are written to dens[b][M][N] (every other pixel is
 * exactly zero and is NOT stored): 4 B per band pixel instead of 28 B per image pixel and iteration.  The stored values
 * and minmax are bit-identical to gpet_density_f64's.  gpet_density_bands_supported: whether (M + 8) rows x
 * (max_width + 8) columns of 64-bit cells fit one CTA's shared memory (M <= 662 at max_width 32); otherwise use
 * gpet_density_f64.  work: gpet_density_bands_workspace_bytes(B, n, Kp); bands i32[B][n_groups][2].
 * gpet_select_bands_f64 = gpet_select_f64 reading such band-limited densities; gpet_kde_bands_f32 expands them to the
 * normalised float32 kde map (inspection / tests). */
int gpet_density_bands_supported(int M, int N, int max_width);
int64_t gpet_density_bands_workspace_bytes(int B, int n, int Kp);
int gpet_density_bands_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp, int M,
                           int N, int x_st, const int32_t* group_cols, int n_groups, int max_width, float* dens,
                           uint32_t* minmax, int32_t* bands, void* work, void* stream);
int gpet_select_bands_f64(const float* dens, const uint32_t* minmax, const float* grad_kde, const int32_t* img_index,
                          const int32_t* bands, int B, int M, int N, const int32_t* col_bin, const int32_t* group_cols,
                          int n_groups, const int32_t* old_yx, const int32_t* n_old, int max_old, int nb,
                          double* bin_score, int32_t* bin_pos, void* stream);
int gpet_kde_bands_f32(const float* dens, const uint32_t* minmax, const int32_t* bands, const int32_t* group_cols,
                       int n_groups, int B, int M, int N, float* kde, void* stream);

/* ---- loop-carried state of __call__ (gpet.py:829-870) on the device -------------------------------------------------
 * The observation sets obs_xy[B][max_old][2] i32 (x, y; the reference's pre_fobs), their sizes n_obs[B], the decaying
 * score thresholds thr[B] f64 (self.score_thresh, gpet.py:595) and the iteration counters n_iter[B] live on the device;
 * the host reads one control block per iteration: ctrl i32[4] = {active traces, error code (0 none, 1 Cholesky of the
 * training kernel matrix failed, 2 the threshold loop of gpet.py:591-609 cannot end), a trace that raised it,
 * the largest observation count among the active traces}.
 *
 * gpet_update_obs_f64: compute_new_obs (gpet.py:589-616) for the B_active compacted traces rows[k] from the per-bin
 * maxima of gpet_select_f64 (bin_score/bin_pos[k][nb]): the threshold is multiplied by 0.95 (by 1.0 on the first pass)
 * until min(n_pre + pixel_thresh, algo_thresh) bins reach it; the accepted bins in ascending order become the new
 * observation set.  post_status[k] (may be NULL) != 0 flags error 1.  Needs max_old >= nb.
 *
 * gpet_training_sets_f64: (1) rows[0..ctrl[0]) = traces with n_obs < algo_thresh (gpet.py:829), ascending, and ctrl[3] =
 * the largest n_obs among them, i.e. K + ctrl[3] bounds the training sets built here (gpet_posterior_lowrank_f64's m_cap;
 * bump_iter is ignored); (2) for every such trace the training set of fit_predict_GP (gpet.py:209-224):
 * stable sort by x of concat(init_xy[b][K][2], obs) -> xi[k][mmax] = x - x_st, y[k][mmax], w[k][mmax] = alpha_init[K] for the
 * initial points and 1 for observations, m[k]; old_yx[k][max_old][2] (row, col) and n_old[k] for gpet_select_f64;
 * (3) ctrl is copied to ctrl_host (pinned, may be NULL).  B_launch >= the number of active traces (e.g. the previous
 * count; B the first time). */
int gpet_update_obs_f64(const double* bin_score, const int32_t* bin_pos, const int32_t* rows, const int32_t* post_status,
                        int B_active, int nb, int N, int max_old, int pixel_thresh, int algo_thresh, int32_t* obs_xy,
                        int32_t* n_obs, double* thr, int32_t* n_iter, int32_t* ctrl, void* stream);
int gpet_training_sets_f64(const int32_t* init_xy, const double* alpha_init, int K, const int32_t* obs_xy,
                           const int32_t* n_obs, int B, int B_launch, int max_old, int algo_thresh, int x_st, int mmax,
                           int bump_iter, int32_t* rows, int32_t* ctrl, int32_t* xi, double* y, double* w, int32_t* m,
                           int32_t* old_yx, int32_t* n_old, int32_t* ctrl_host, void* stream);

/* ---- fit_predict_GP(converged=True): objective of the final hyper-parameter fit ------------------------------
 * E evaluations of -(log marginal likelihood) and its gradient (sklearn_gpr.py:257-262, 475-585) for the kernel
 * Constant*(RBF|Matern) + WeightedWhite, theta[e][3] = log[constant, length_scale, noise_level]; evaluation e
 * uses the training set of trace trace_of[e]: X[t][mmax] (standardised x), y[t][mmax] (standardised y), w[t][mmax],
 * m[t].  kind: 0 RBF, 1/2/3 Matern nu = 0.5/1.5/2.5.  f[e] = +inf and g = 0 when the Cholesky fails (:521-522).
 * xcol (may be NULL): xcol[t][mmax] i32, the integer pixel columns X was standardised from (ascending); with it the
 * RBF kernel values are tabulated per distinct pixel distance instead of evaluated per matrix entry.
 * trace_of[e] < 0: slot e is skipped (f, g untouched).  The L-BFGS-B iterations run on the device
 * (gpet_lbfgsb_*, below) or on the host (scipy's setulb, one instance per start). */
int gpet_lml_f64(const double* X, const double* y, const double* w, const int32_t* xcol, const int32_t* m, int mmax,
                 const int32_t* trace_of, const double* theta, int E, int kind, double gp_alpha, double* f, double* g,
                 void* stream);

/* ---- L-BFGS-B state machines of the final fit, on the device -------------------------------------------------------
 * Replaces the host loop around scipy.optimize.minimize(method='L-BFGS-B', jac=True, bounds=...) that the reference
 * runs 13 times per trace (sklearn_gpr.py:254-295, 587-607; scipy's _minimize_lbfgsb / setulb): E independent runs over
 * theta[3] with both bounds finite, m = 10 corrections, ftol = 2.22e-9, gtol = 1e-5, maxls = 20, advanced in lock step
 * with gpet_lml_f64.  State: dstate[gpet_lbfgsb_state_doubles()][E] f64 and istate[gpet_lbfgsb_state_ints()][E] i32
 * (run e at column e), caller-owned.
 *   init:    x0[E][3] is clipped to [lo, up] (lo[3], up[3] device arrays shared by all runs).
 *   advance: first != 0 on the first round; otherwise every run with trace_eval[e] >= 0 takes f[e], g[e][3] (the
 *            objective at the theta[e][3] it asked for) and advances.  On return theta[e][3] is the next evaluation
 *            point and trace_eval[e] = trace_of[e] for runs that wait for an evaluation, trace_eval[e] = -1 for runs
 *            that have ended; n_active (device i32[3]): [0] = number of waiting runs (reset here), [1] += that number,
 *            [2] += 1 if it is non-zero.  gpet_lml_f64 skips slots whose
 *            trace_of entry is negative, so one round is advance -> gpet_lml_f64(trace_of = trace_eval, E).
 *   result:  x[E][3] final point, fval[E] objective there, nfev[E] evaluations, task[E] (4 converged, 5 abnormal
 *            line-search termination, 6 iteration/evaluation limit).
 * gpet_lbfgsb_host_init / _host_advance run the same code on host memory (run e at dstate + e*doubles,
 * istate + e*ints; give[e] != 0: take f[e], g[e]; need[e] = 1: waits for an evaluation at x[e]); they exist so the
 * algorithm can be checked against scipy's own setulb without a GPU. */
int64_t gpet_lbfgsb_state_doubles(void);
int64_t gpet_lbfgsb_state_ints(void);
int gpet_lbfgsb_init_f64(double* dstate, int32_t* istate, int E, const double* x0, const double* lo, const double* up,
                         void* stream);
int gpet_lbfgsb_advance_f64(double* dstate, int32_t* istate, int E, int first, const int32_t* trace_of, const double* f,
                            const double* g, double* theta, int32_t* trace_eval, int32_t* n_active, void* stream);
/* n_rounds x [gpet_lbfgsb_advance_f64 -> gpet_lml_f64(trace_of = trace_eval)] enqueued back to back (first != 0: the
 * very first round of the fit), then counters[3] is copied to counters_host (pinned).  counters (device i32[3], zeroed by
 * the caller before the first call): [0] runs waiting after the last round of this call, [1] evaluations requested so
 * far, [2] rounds so far that had at least one waiting run.  Rounds after the last run has ended are empty. */
int gpet_fit_rounds_f64(const double* X, const double* y, const double* w, const int32_t* xcol, const int32_t* m,
                        int mmax, int kind, double gp_alpha, double* dstate, int32_t* istate, int E, int first,
                        int n_rounds, const int32_t* trace_of, double* f, double* g, double* theta,
                        int32_t* trace_eval, int32_t* counters, int32_t* counters_host, void* stream);
int gpet_lbfgsb_result_f64(const double* dstate, const int32_t* istate, int E, double* x, double* fval, int32_t* nfev,
                           int32_t* task, void* stream);
int gpet_lbfgsb_host_init(double* dstate, int32_t* istate, int E, const double* x0, const double* lo, const double* up);
int gpet_lbfgsb_host_advance(double* dstate, int32_t* istate, int E, const int32_t* give, const double* f,
                             const double* g, int32_t* need, double* x);

/* predict(return_std) at the optimum (sklearn_gpr.py:379-436, gpet.py:263-266): xq[t][n] standardised grid,
 * tm_ts[t][2] = (mean, std) removed from y by the regressor; mean[t][n] = ts*(K* alpha)+tm, sd[t][n]. */
int gpet_final_predict_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int T,
                           const double* theta, int kind, double gp_alpha, const double* xq, int n,
                           const double* tm_ts, double* mean, double* sd, int32_t* status, void* stream);

/* ---- training sets beyond GPET_MAX_TRAIN (BASELINE config 3: delta_x = 2 on 4096 columns -> m up to 2046) -------------------
 * Same seams and arithmetic as gpet_lml_f64 / gpet_fit_rounds_f64 / gpet_final_predict_f64 (sklearn_gpr.py:254-295, 379-436,
 * 475-585), with the kernel matrices in `work` (HBM): blocked Cholesky, alpha by substitution, T = L^-1 by blocked forward
 * substitution on the identity, K^-1 = T^T T consumed tile by tile in the gradient sums (fixed summation order).
 * gpet_lml_big_f64 evaluates the E slots in chunks of as many evaluations as `work_bytes` holds
 * (gpet_lml_big_workspace_bytes(E, mmax) = all at once; at least gpet_lml_big_workspace_bytes(1, mmax)). */
int64_t gpet_lml_big_workspace_bytes(int E, int mmax);
int gpet_lml_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax,
                     const int32_t* trace_of, const double* theta, int E, int kind, double gp_alpha, double* f, double* g,
                     void* work, int64_t work_bytes, void* stream);
int gpet_fit_rounds_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int kind,
                            double gp_alpha, double* dstate, int32_t* istate, int E, int first, int n_rounds,
                            const int32_t* trace_of, double* f, double* g, double* theta, int32_t* trace_eval,
                            int32_t* counters, int32_t* counters_host, void* work, int64_t work_bytes, void* stream);
int64_t gpet_final_predict_big_workspace_bytes(int T, int mmax, int n);
int gpet_final_predict_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int T,
                               const double* theta, int kind, double gp_alpha, const double* xq, int n,
                               const double* tm_ts, double* mean, double* sd, int32_t* status, void* work, void* stream);
/* The two blocked primitives on their own (LAPACK dpotrf 'L' / dtrsm 'L','L','N','N' batched over matrices of different
 * sizes): A[b][ld][ld] row-major, ld a multiple of 64, m[b] <= m_cap <= ld rows in use; rows from m[b] up to the next
 * multiple of 64 must hold an identity block (zeros left of the diagonal).  status[b] = 1: not positive definite.
 * trsm: R[b][ld][ldr] <- L^-1 R (ldr a multiple of 64); ident != 0: R is not read, its lower 64 x 64 tiles receive L^-1
 * (ldr == ld). */
int gpet_dense_potrf_f64(double* A, int ld, const int32_t* m, int B, int m_cap, int32_t* status, void* stream);
int gpet_dense_trsm_f64(const double* L, int ld, const int32_t* m, int B, int m_cap, double* R, int ldr, int ident,
                        void* stream);

/* ---- eigen-decomposition of the full n x n posterior covariance (full-rank kernels: Matern) ----------------------------
 * numpy's multivariate_normal factors the covariance by SVD (sklearn_gpr.py:460-464); for a symmetric positive
 * semi-definite matrix that is its eigen-decomposition.  Two-sided block Jacobi in HBM (csrc/gpet_jacobi.cu): blocks of
 * 64 columns, round-robin pairs, 128 x 128 pivots solved by gpet_sym_eig_f64, updates as fp64 tensor-instruction tiles.
 * np = n rounded up to a multiple of 128.  init: A[b][np][np] = padded copy of cov[b][n][n] (lower triangle mirrored),
 * V[b][np][np] = I.  sweep: one pass over all block pairs; off[b][2] = (off-diagonal, total) squared Frobenius norms of
 * A[b] afterwards - the caller repeats sweeps until off[b][0] <= tol^2 off[b][1] (5-8 sweeps).  factor: rows of
 * F[b][rp][n] = sign sqrt(max(d_r, 0)) v_r^T for the eigenpairs in descending order, sign such that <v_r, w> > 0
 * (canonical signs); rows n..rp-1 zero.  work: gpet_block_jacobi_workspace_bytes(B, np) for both calls. */
int64_t gpet_block_jacobi_workspace_bytes(int B, int np);
int gpet_block_jacobi_init_f64(const double* cov, int B, int n, int np, double* A, double* V, void* stream);
int gpet_block_jacobi_sweep_f64(double* A, double* V, int B, int np, double* off, void* work, void* stream);
/* Instead of init, when V[b] holds the eigenvectors of a nearby matrix (the covariance of the previous iteration of a trace):
 * V <- V (1.5 I - 0.5 V^T V) (one Newton-Schulz step, removes the orthogonality drift of earlier starts), A = V^T Sigma V
 * (nearly diagonal: 2-4 sweeps instead of 8).  tmp: 2 * B * np * np doubles. */
int gpet_block_jacobi_warm_f64(const double* cov, int B, int n, int np, double* A, double* V, void* tmp, void* stream);
int gpet_block_jacobi_factor_f64(const double* A, const double* V, int B, int n, int np, int rp, const double* w,
                                 double* F, void* work, void* stream);

/* ---- bench inputs and trace-quality metrics on the device (gpet_utils.py:163-253, 256-313) ---------------------------------
 * gpet_test_img_f64: img[b][y][x] = intensity for y >= rows[b][x] (1 - intensity for y >= rows2[b][x] when rows2 != NULL:
 * the multi-sinusoidal types), 0 above and in the four gap column runs when gaps != 0; then, when noise != NULL,
 * clip(img + noise_sd * noise[b][y][x], 0, 1).  The edge rows are evaluated by the caller (N integers per image).
 * gpet_trace_metrics_f64: out[b][3] = (trace_MSE, trace_relarea, Jaccard index; DICE = 2J / (J + 1)) of
 * edge_pred[b][n][2] int64 (y, x) against true_rows[b][n], unrounded. */
int gpet_test_img_f64(const int32_t* rows, const int32_t* rows2, int B, int M, int N, double intensity, int gaps,
                      const double* noise, double noise_sd, double* img, void* stream);
int gpet_trace_metrics_f64(const int64_t* edge_pred, const int32_t* true_rows, int B, int n, double* out, void* stream);

/* normalised kde map (float32) for inspection / tests: kde = (dens - min) / (max - min) in float32 */
int gpet_kde_normalised_f32(const float* dens, const uint32_t* minmax, int B, int M, int N, float* kde,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPET_B200_H */
