/*
 * verde_b200.h -- C ABI of the B200-native Green's-function spline engine.
 *
 * This is the drop-in boundary for ONE hot path of fatiando/verde: the numeric
 * kernels under verde.Spline / verde.SplineCV / verde.VectorSpline2D.  The
 * reference reaches that path through caller-allocates-output free functions on
 * contiguous float64 1-D arrays (numba-jitted) plus `least_squares`; each entry
 * point below names the reference function it replaces (paths relative to the
 * reference checkout).  A maintainer binds them with ctypes exactly as
 * INTEGRATION.md shows.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++/torch types; all arrays float64, contiguous.
 *  - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE
 *    pointer of the current CUDA device; the library checks with
 *    cudaPointerGetAttributes and stages host data itself.  The caller owns all
 *    buffers, outputs are written in place.
 *  - calls are synchronous: results are complete when the function returns.
 *  - return value: 0 = ok; > 0 = numerical failure (1-based index of the first
 *    non-positive Cholesky pivot, LAPACK `info` style); < 0 = VB200_E_* below.
 *    Nothing throws across the ABI.  vb200_last_error() gives a message.
 *  - the library keeps per-process state (cached device workspace, options, last error per thread):
 *    drive it from one host thread per process (the intended deployment is one process per GPU).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    returns VB200_E_CUDA.
 */
#ifndef VERDE_B200_H
#define VERDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_E_BADARG (-1)
#define VB200_E_CUDA (-2)
#define VB200_E_NOMEM (-3)
#define VB200_E_NOCONV (-4) /* the Jacobi SVD of a *_lstsq call did not converge within its sweep limit */

/* Diagnostics of one fit (all optional; pass NULL to skip). */
typedef struct vb200_fit_info {
    int32_t info;            /* 0, or 1-based index of the failed pivot                        */
    int32_t block_rows;      /* T: 128-wide block rows of the padded normal matrix             */
    int64_t rows;            /* rows of the least-squares system (n, or 2n for the 2-D spline) */
    int64_t cols;            /* unknowns (m, or 2m)                                            */
    double gram_ms;          /* device time: panel generation + Gram accumulation              */
    double factor_ms;        /* device time: scaling + Cholesky                                */
    double solve_ms;         /* device time: forward/back substitution + unscaling             */
    double gemm_ms;          /* device time inside the DMMA tile kernel (launches that own the GPU; see shared_gemm_*) */
    double gemm_flops;       /* algorithmic flops those launches performed (2*128^3 per tile-chunk) */
    int64_t gemm_launches;   /* launches of the DMMA tile kernel                               */
    int64_t kernel_launches; /* all kernel launches of this fit                                */
    double diag_min;         /* min / max of diag(L)^2: a cheap conditioning indicator         */
    double diag_max;
    /* --- since ABI version 200 (SURVEY 8b out_diag{cond, backward_err}); 0 when "fit_diagnostics" is off --- */
    double lambda_max;       /* largest eigenvalue of the damped, column-scaled normal matrix A (power iteration on L L^T) */
    double lambda_min_est;   /* Rayleigh-quotient estimate (>= the true value) of its smallest eigenvalue (inverse iteration) */
    double cond_est;         /* lambda_max / lambda_min_est  <= cond_2(A)                       */
    double cond_upper;       /* lambda_max / damping         >= cond_2(A)  (G is positive semi-definite) */
    double backward_err;     /* ||Js^T W (d - Js c) - damping c|| / (lambda_max ||c|| + ||Js^T W d||), matrix-free, exact
                                Green's functions; only the fused fits (vb200_spline_fit*, vb200_least_squares) and
                                vb200_backward_error fill it, otherwise -1                       */
    double diag_ms;          /* device + host time spent on these diagnostics                   */
    /* --- the damping=None entry points (vb200_*_lstsq) only; 0 otherwise --- */
    int64_t svd_rank;        /* singular values kept: sigma > rcond * sigma_max (LinearRegression.rank_) */
    int64_t svd_sweeps;      /* one-sided Jacobi sweeps until a sweep without rotations        */
    double sigma_max;        /* largest / smallest KEPT singular value of sqrt(W) Js           */
    double sigma_min_kept;
    int32_t power_its;       /* iterations the condition estimate took: power (lambda_max) / inverse (lambda_min) */
    int32_t inverse_its;
    /* --- since ABI version 201.  Look-ahead Cholesky: DMMA launches that ran BESIDE the factorisation of the next panel (on SMs - R of the
     * SMs, joined later by a second grid on the same tile queue).  Their event spans time a kernel that owns only part
     * of the machine, so they are reported apart from gemm_ms / gemm_flops (the roofline figures). --- */
    double shared_gemm_ms;
    double shared_gemm_flops;
    int64_t shared_gemm_launches;
} vb200_fit_info;

/* ---- library / device -------------------------------------------------------------- */
/* 201: vb200_fit_info grew shared_gemm_*; added vb200_get_option, vb200_chol_update_ranges, vb200_shared_*,
 * vb200_peer_copy_async, vb200_chol_wait_panel, vb200_stream_wait_flag, vb200_block_reduce_stream */
int vb200_version(void);
const char *vb200_last_error(void);
int vb200_device_count(void);
int vb200_set_device(int device);
/* tunables: "predict_exact" (0/1), "chol_panel_tiles" (1..16), "gram_slab_chunks" (1..16),
 * "gram_fused" (0/1: generate Jacobian rows in shared memory inside the Gram kernel instead of staging
 * a panel), "gemm_variant" (0/1/2/3/4; 4 = persistent kernel with a dynamic tile queue), "gemm_grid_limit", "gemm_strip", "predict_pts", "predict_minb", "predict_wide" (0/1),
 * "sweep_group" (systems per batch of vb200_gram_solve_multi, 0 = as many as fit), "solve_variant",
 * "fit_diagnostics" (0/1, default 1: condition estimate + backward error after every fit), "cond_inverse_its" (inverse
 * iterations of the condition estimate; 0 = automatic), "chol_lookahead" (SMs that factor panel p+1 beside the trailing
 * update of panel p; 0 = off), "chol_lookahead_min_rows" */
int vb200_set_option(const char *key, double value);
/* current value of a tunable (so that callers can restore it) */
int vb200_get_option(const char *key, double *value);
/* frees the cached device workspace (Gram matrix, panels) */
int vb200_release_workspace(void);
/* kernels launched by this library since it was loaded (bench.py's gpu_launches) */
int64_t vb200_launch_count(void);
/* FP64 issue-rate microbenchmark on the current device, TFLOP/s: kind 0 = DMMA.8x8x4
 * (mma.sync f64), kind 1 = DFMA.  The roofline denominator of the fit kernels;
 * returns < 0 on CUDA errors. */
double vb200_measure_fp64_peak_tflops(int kind);

/* ---- K0: Jacobian materialisation --------------------------------------------------- */
/* replaces verde/spline.py:642-650 jacobian_numba(east, north, force_east, force_north,
 * mindist, jac): jac is row-major (n, m). */
int vb200_jacobian(const double *east, const double *north, int64_t n, const double *force_east,
                   const double *force_north, int64_t m, double mindist, double *jac);
/* replaces verde/vector.py:461-475 jacobian_2d_numba(..., mindist, poisson, jac):
 * jac is row-major (2n, 2m) = [[ee, ne], [ne, nn]]. */
int vb200_jacobian_2d(const double *east, const double *north, int64_t n,
                      const double *force_east, const double *force_north, int64_t m,
                      double mindist, double poisson, double *jac);

/* ---- K3: matrix-free predict -------------------------------------------------------- */
/* replaces verde/spline.py:629-639 predict_numba(east, north, force_east, force_north,
 * mindist, forces, result). */
int vb200_predict(const double *east, const double *north, int64_t n, const double *force_east,
                  const double *force_north, const double *forces, int64_t m, double mindist,
                  double *result);
/* several coefficient vectors on the SAME forces (a verde.Vector of Splines on shared coordinates,
 * verde/vector.py:134-141: `tuple(comp.predict(coordinates) for comp in self.components)` evaluates
 * every Green's function once per component; here once per pair).  forces: n_components x m,
 * result: n_components x n, both row-major; each row equals its own vb200_predict bit for bit. */
int vb200_predict_multi(const double *east, const double *north, int64_t n, const double *force_east,
                        const double *force_north, const double *forces, int64_t m,
                        int32_t n_components, double mindist, double *result);
/* replaces verde/vector.py:443-458 predict_2d_numba(..., mindist, poisson, forces, vec_east,
 * vec_north); forces has length 2m (east forces, then north forces). */
int vb200_predict_2d(const double *east, const double *north, int64_t n,
                     const double *force_east, const double *force_north, const double *forces,
                     int64_t m, double mindist, double poisson, double *vec_east,
                     double *vec_north);

/* ---- K1 + K2: fused fit -------------------------------------------------------------- */
/* replaces, in one call, verde/spline.py:462-463:
 *     jacobian = self.jacobian(coordinates[:2], self.force_coords_)
 *     self.force_ = least_squares(jacobian, data, weights, self.damping)
 * (verde/base/least_squares.py:17-71, damped branch).  weights may be NULL.
 * out_force has length m. */
int vb200_spline_fit(const double *east, const double *north, const double *data,
                     const double *weights, int64_t n, const double *force_east,
                     const double *force_north, int64_t m, double mindist, double damping,
                     double *out_force, vb200_fit_info *info);
/* replaces verde/vector.py:287-288 for VectorSpline2D.  data / weights have length 2n (east
 * components then north components, verde/vector.py:279-284); out_force has length 2m. */
int vb200_spline_fit_2d(const double *east, const double *north, const double *data,
                        const double *weights, int64_t n, const double *force_east,
                        const double *force_north, int64_t m, double mindist, double poisson,
                        double damping, double *out_force, vb200_fit_info *info);

/* ---- staged fit, for row-sharded multi-GPU and for SplineCV's damping sweeps -------------
 * gram:   accumulates this rank's rows into G (tile-packed lower triangle, DEVICE memory,
 *         vb200_gram_doubles(cols) doubles) and stats (DEVICE, vb200_stats_doubles(cols)
 *         doubles = {J^T W d, J^T 1, diag(J^T J), reserved}).  accumulate = 0 overwrites.
 *         kind 0 = scalar spline, 1 = 2-D elastic spline.
 * (the caller all-reduces G and stats across ranks: both are plain sums over data rows)
 * solve:  scales (StandardScaler algebra), adds damping, factors and solves.  G is destroyed
 *         unless keep_gram != 0 (then a device copy is factored instead).  total_rows is the
 *         number of system rows summed over all ranks. */
int64_t vb200_gram_doubles(int64_t cols);
int64_t vb200_stats_doubles(int64_t cols);
int vb200_gram_accumulate(int kind, const double *east, const double *north, const double *data,
                          const double *weights, int64_t n, const double *force_east,
                          const double *force_north, int64_t m, double mindist, double poisson,
                          double *gram_dev, double *stats_dev, int accumulate,
                          vb200_fit_info *info);
int vb200_gram_solve(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                     double damping, int keep_gram, double *out_params, vb200_fit_info *info);

/* solve_multi: SplineCV's damping sweep (verde/spline.py:243-255 refits every (mindist, damping, fold)
 *         from scratch; the Gram matrix only depends on (mindist, fold)).  Scales G once, writes one
 *         damped copy per damping and factors / solves all of them SIDE BY SIDE: every panel kernel
 *         and every DMMA update launch covers the whole batch.  G is not modified.  dampings: HOST
 *         array of n_dampings positive values; out_params: n_dampings x cols, row-major (host or
 *         device); out_info (HOST, may be NULL): per-damping 0 or 1-based failed pivot.  As many
 *         systems as fit in free device memory are batched at a time. */
int vb200_gram_solve_multi(const double *gram_dev, const double *stats_dev, int64_t cols,
                           int64_t total_rows, const double *dampings, int32_t n_dampings,
                           double *out_params, int32_t *out_info, vb200_fit_info *info);

/* ---- stepwise Cholesky for the multi-GPU panel-cyclic factorisation ---------------------------
 * The replicated m x m Cholesky is the Amdahl term of the row-sharded fit.  These entry points expose
 * its outer steps so that N ranks, each holding the all-reduced G, share the m^3/3 flops: outer
 * panel p (pw block columns of 128) is owned by rank p % N, which factors it (chol_panel) and
 * broadcasts the panel -- ONE contiguous range of the tile-packed triangle, chol_panel_range -- to
 * the others (NCCL over NVLink, done by the caller); every rank then applies the panel to the
 * block columns IT owns (chol_update).  After the last panel every rank holds the complete factor
 * and chol_finish runs the substitution.  begin/panel/update only enqueue work on the library
 * stream (no host synchronisation); finish synchronises and reports like vb200_gram_solve.
 * One factorisation at a time per process.  verde_b200/kernels.py::GramSystem.solve_distributed
 * is the driver. */
int vb200_chol_begin(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                     double damping);
int vb200_chol_panel_range(int64_t cols, int32_t k0, int32_t pw, int64_t *offset, int64_t *count);
int vb200_chol_panel(double *chol_dev, int64_t cols, int32_t k0, int32_t pw);
int vb200_chol_update(double *chol_dev, int64_t cols, int32_t k0, int32_t pw, int32_t j_lo,
                      int32_t j_hi);
/* the same update for several ascending, disjoint block-column ranges [j_lo[r], j_hi[r]) in ONE launch of the DMMA
 * kernel (dynamic tile queue): the panels one rank owns are strided, and one launch has one tail instead of one per
 * panel.  j_lo / j_hi are host arrays. */
int vb200_chol_update_ranges(double *chol_dev, int64_t cols, int32_t k0, int32_t pw, int32_t n_ranges,
                             const int32_t *j_lo, const int32_t *j_hi);
/* Device buffers that the other processes of the node can map (CUDA IPC), for the peer-memory panel exchange:
 * shared_alloc = cudaMalloc + cudaIpcGetMemHandle (handle: 64 bytes, sent to the peers by the host layer);
 * shared_open  = cudaIpcOpenMemHandle in the CALLER's device context (the peer's buffer becomes addressable from
 *                this GPU over NVLink); shared_close / shared_free undo them (close every mapping before freeing). */
int vb200_shared_alloc(int64_t bytes, void **ptr, void *handle);
int vb200_shared_free(void *ptr);
int vb200_shared_open(const void *handle, void **ptr);
int vb200_shared_close(void *ptr);
/* Peer-memory panel exchange (one process per GPU on one node, buffers shared over CUDA IPC by the host layer): parks
 * the library stream until the 64-bit word *flag_dev, which the panel's owner writes AFTER pushing the panel into this
 * rank's matrix with a copy engine, is >= epoch.  Asynchronous; a panel that does not arrive within ~10 s makes
 * vb200_chol_finish return VB200_E_CUDA instead of hanging. */
int vb200_chol_wait_panel(const int64_t *flag_dev, int64_t epoch);
/* the same wait on a stream of the caller (cudaStream_t; NULL = the library stream): an owner parks its copy stream
 * on a peer's "ready" word before the first push of a factorisation, so that no panel can land in a matrix whose
 * scaling pass (vb200_chol_begin) is still running */
int vb200_stream_wait_flag(const int64_t *flag_dev, int64_t epoch, void *stream);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream) for the pushes above: dst is a peer GPU's memory mapped
 * into this process (CUDA IPC), stream a cudaStream_t of the caller (NULL = the library stream).  Asynchronous. */
int vb200_peer_copy_async(void *dst, const void *src, int64_t bytes, void *stream);
int vb200_chol_finish(const double *chol_dev, int64_t cols, double *out_params,
                      vb200_fit_info *info);

/* ---- factor once, solve many right-hand sides (Vector of Splines on shared coordinates) ------
 * factor:  like vb200_gram_solve but stops after the Cholesky factorisation: gram_dev is overwritten
 *          by L (tile-packed) and inv_scale_dev (DEVICE, vb200_padded_cols(cols) doubles) receives
 *          the column scaling 1/s_j.
 * solve:   out_params = D L^-T L^-1 D rhs for one UNSCALED right-hand side rhs = J^T W d
 *          (length cols, host or device).
 * jacobian_transpose_apply: out[j] = sum_i g(e_i - fe_j, n_i - fn_j) v[i], i.e. J^T v for the
 *          scalar spline without forming J (v = W d gives the right-hand side of another data
 *          component on the same coordinates). */
int64_t vb200_padded_cols(int64_t cols);
int vb200_gram_factor(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                      double damping, double *inv_scale_dev, vb200_fit_info *info);
int vb200_chol_solve(const double *chol_dev, const double *inv_scale_dev, int64_t cols,
                     const double *rhs, double *out_params, vb200_fit_info *info);
int vb200_jacobian_transpose_apply(const double *east, const double *north, int64_t n,
                                   const double *force_east, const double *force_north, int64_t m,
                                   double mindist, const double *v, double *out);

/* ---- damping=None ------------------------------------------------------------
 * replaces verde/base/least_squares.py:65-70 with damping=None: StandardScaler column scaling, then
 * LinearRegression(fit_intercept=False).fit(Js, d, sample_weight=w) = scipy.linalg.lstsq(sqrt(W) Js, sqrt(W) d,
 * cond=rcond) -- the minimum-norm solution with every singular value <= rcond * sigma_max dropped (LAPACK gelsd) --
 * unscaled on return.  rcond is the caller's: scikit-learn 1.9 passes LinearRegression.tol = 1e-6.  Computed by a
 * one-sided Jacobi SVD of the explicit scaled Jacobian on the device (vb_svd.cu); arguments otherwise as the damped
 * twins above.  info (optional) receives svd_rank, svd_sweeps, sigma_max, sigma_min_kept, cond_est =
 * sigma_max / sigma_min_kept and solve_ms.  Returns VB200_E_NOCONV if the sweeps do not converge. */
int vb200_spline_fit_lstsq(const double *east, const double *north, const double *data,
                           const double *weights, int64_t n, const double *force_east,
                           const double *force_north, int64_t m, double mindist, double rcond,
                           double *out_force, vb200_fit_info *info);
int vb200_spline_fit_2d_lstsq(const double *east, const double *north, const double *data,
                              const double *weights, int64_t n, const double *force_east,
                              const double *force_north, int64_t m, double mindist, double poisson,
                              double rcond, double *out_force, vb200_fit_info *info);
int vb200_least_squares_lstsq(const double *jac, int64_t n, int64_t m, const double *data,
                              const double *weights, double rcond, double *out_params,
                              vb200_fit_info *info);

/* Diagnostics for the staged (multi-GPU) fit, SURVEY 8b out_diag{cond, backward_err}:
 * vb200_normal_residual: grad = J^T W (d - J force) over THIS caller's rows, matrix-free with the exact Green's
 *         functions (kind 0 scalar / 1 elastic, as vb200_gram_accumulate; data and weights stacked for kind 1;
 *         grad has m or 2m entries, host or device).  Ranks sum their partial gradients (m doubles).
 * vb200_backward_error: eta = ||D grad - damping c|| / (lambda_max ||c|| + ||D J^T W d||), c = force / D, from
 *         the summed gradient, the all-reduced statistics vector of vb200_gram_accumulate and lambda_max of
 *         vb200_fit_info; the same quantity the fused fits report as vb200_fit_info.backward_err. */
int vb200_normal_residual(int kind, const double *east, const double *north, const double *data,
                          const double *weights, int64_t n, const double *force_east,
                          const double *force_north, int64_t m, double mindist, double poisson,
                          const double *force, double *grad);
int vb200_backward_error(const double *grad, const double *stats_dev, int64_t cols, int64_t total_rows,
                         double damping, const double *force, double lambda_max, double *eta);

/* ---- either side of the spline path ---------------------------------
 * block_split: replaces verde/coordinates.py:848-946 block_split(coordinates, ...): every point gets the
 *         index of the nearest block centre, label = i_north * n_east + i_east over the centre grid the
 *         host builds exactly like the reference (grid_coordinates(..., pixel_register=True)).  The
 *         reference asks a KD-tree; centres are a product grid, so the nearest centre is the nearest
 *         per axis.  Exact ties go to the lower index (the KD-tree's choice depends on its traversal).
 *         labels (n int64, host or device) may be NULL.  n_blocks receives the number of NON-EMPTY
 *         blocks.  The sorted (block, point) order stays cached for the calls below.
 * block_ids: ascending labels of the non-empty blocks (n_blocks int64).
 * block_aggregate: replaces the pandas groupby().aggregate(...) of verde/blockreduce.py:175-186 and
 *         406-506 for one column: out[k] = reduction over the points of the k-th non-empty block,
 *         visited in original point order.  weights is only read by the weighted reductions. */
#define VB200_REDUCE_MEAN 0
#define VB200_REDUCE_SUM 1
#define VB200_REDUCE_MIN 2
#define VB200_REDUCE_MAX 3
#define VB200_REDUCE_COUNT 4
#define VB200_REDUCE_MEDIAN 5
#define VB200_REDUCE_VAR 6       /* sample variance, ddof = 1 (pandas "var") */
#define VB200_REDUCE_WMEAN 7     /* numpy.average(values, weights=w) */
#define VB200_REDUCE_WVAR 8      /* numpy.average((values - wmean)**2, weights=w) */
#define VB200_REDUCE_INV_SUM_W 9 /* 1 / sum(w) */
#define VB200_REDUCE_VAR0 10     /* population variance, ddof = 0 (numpy.var applied as a callable, pandas >= 3) */
#define VB200_REDUCE_STD 11      /* sample standard deviation, ddof = 1 (pandas "std") */
#define VB200_REDUCE_STD0 12     /* population standard deviation, ddof = 0 (numpy.std) */
int vb200_block_split(const double *east, const double *north, int64_t n, const double *centers_east,
                      int32_t n_east, const double *centers_north, int32_t n_north, int64_t *labels,
                      int64_t *n_blocks);
int vb200_block_ids(int64_t *ids);
int vb200_block_aggregate(const double *values, const double *weights, int64_t n, int32_t op,
                          double *out);
/* Order-free block reduction in ONE pass (no grouping, 24 B read per point and column): every point is labelled
 * exactly as vb200_block_split labels it and added into dense per-block accumulators with atomics; the non-empty blocks
 * come back in label order.  op: VB200_REDUCE_SUM / MEAN / COUNT, or WMEAN with weights ([n], applied to every column).
 * values: [n_columns][n] (n_columns may be 0 for ids / counts only); out_ids: [capacity]; out_values:
 * [n_columns][capacity] (column c, block j at c * capacity + j); capacity >= min(n, n_east * n_north) always suffices.
 * Block grids up to 2^26 blocks.  Sums differ from the ordered path (vb200_block_aggregate, pandas' order) by the
 * rounding order of the additions only; replaces the same reference lines as vb200_block_split + block_aggregate
 * (verde/blockreduce.py:175-186). */
int vb200_block_reduce_stream(const double *east, const double *north, int64_t n, const double *centers_east,
                              int32_t n_east, const double *centers_north, int32_t n_north, const double *values,
                              int32_t n_columns, const double *weights, int32_t op, int64_t capacity, int64_t *out_ids,
                              double *out_values, int64_t *n_blocks);

/* distance_mask: replaces verde/mask.py:104-109 (KD-tree nearest-neighbour distance <= maxdist):
 *         mask[i] = 1 when some data point lies within maxdist of (east[i], north[i]), else 0.
 *         mask: n_points bytes, host or device. */
int vb200_distance_mask(const double *data_east, const double *data_north, int64_t n, double maxdist,
                        const double *east, const double *north, int64_t n_points, uint8_t *mask);

/* ---- least squares on an explicit Jacobian ----------------------------------------------
 * replaces verde/base/least_squares.py:17 least_squares(jacobian, data, weights, damping):
 * jac row-major (n, m); damped branch only (damping > 0). */
int vb200_least_squares(const double *jac, int64_t n, int64_t m, const double *data,
                        const double *weights, double damping, double *out_params,
                        vb200_fit_info *info);

#ifdef __cplusplus
}
#endif
#endif /* VERDE_B200_H */
