// vb_internal.h -- internal (non-ABI) declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VB_MAX_KCHUNKS 16
#define VB_MAX_RANGES 32

// One launch of the DMMA tile kernel:
//   for every lower-triangle tile (bi, bj), j_lo <= bj < j_hi, bj <= bi < T:
//     C(bi,bj)[r][c] (+)= alpha * sum_{kk < nk} sum_{k < 128}
//                          Bop_kk[bi][k][r] * Aop_kk[bj][k][c]
// where Xop_kk[b] = x_base + x_chunk_off[kk] + b * 16384 is a k-major operand
// chunk chunk[k][i] of 128 x 128 doubles.
struct VbGemmParams {
    const double *a_base;
    const double *b_base;
    long long a_chunk_off[VB_MAX_KCHUNKS];
    long long b_chunk_off[VB_MAX_KCHUNKS];
    int nk;
    double *c_base;
    int T;
    int j_lo, j_hi;
    double alpha;
    int accumulate;  // 0: C = alpha*acc, 1: C += alpha*acc
    int strip;       // rasterisation strip width in block columns (filled by vb_launch_gemm)
    // batch > 1: the same update on `batch` independent matrices (SplineCV's damping sweep); matrix b
    // has every base pointer advanced by b * batch_stride doubles.  0 or 1 = a single matrix.
    int batch;
    long long batch_stride;
    // dynamic tile queue (gemm variant 4 only).  queue == nullptr: the launch gets a fresh, zeroed counter.  Several
    // launches given the SAME counter (vb_gemm_new_queue) drain one queue -- SMs freed by a concurrent panel
    // factorisation join an update that is already running.  grid_limit > 0 caps the persistent grid.
    unsigned long long *queue;
    int grid_limit;
    // nranges > 0 (dynamic kernel only): the launch covers the block-column ranges [range_lo[r], range_hi[r]) instead of
    // [j_lo, j_hi) -- the panels ONE rank owns in the panel-cyclic multi-GPU factorisation are strided, and one launch
    // over all of them has one tail instead of one per panel.  range_first[r] = tiles before range r.
    int nranges;
    int range_lo[VB_MAX_RANGES], range_hi[VB_MAX_RANGES];
    long long range_first[VB_MAX_RANGES];
    int untimed;  // 1: keep this launch out of the CUDA-event profile (the second grid of a shared queue);
                  // 2: time it, but as a launch that shares the GPU with a panel factorisation (shared_gemm_*)
};

long long vb_gemm_num_tiles(int T, int j_lo, int j_hi);
// tiles of one matrix of a launch (all ranges); fills range_first when the launch has ranges
long long vb_gemm_launch_tiles(VbGemmParams &p);
cudaError_t vb_launch_gemm(const VbGemmParams &p, cudaStream_t stream);
// a zeroed tile-queue counter (zeroed in `stream` order) for launches that share one queue
cudaError_t vb_gemm_new_queue(unsigned long long **queue, cudaStream_t stream);
// same, but recorded by the API layer's CUDA-event profile (vb_api.cu)
cudaError_t vb_gemm_dispatch(const VbGemmParams &p, cudaStream_t stream);
// SM count of the CURRENT device (cached per device ordinal: a process may switch devices)
cudaError_t vb_sm_count(int *sms);
// every kernel launch of the library bumps this (the bench's gpu_launches claim)
extern long long vb_launch_counter;
// DMMA tile kernel flavour (see vb_gemm.cu)
extern int vb_gemm_variant, vb_gemm_strip, vb_gemm_grid_limit;
// predict kernel tuning knobs (vb_greens.cu)
extern int vb_predict_minb, vb_predict_pts, vb_predict_wide;

cudaError_t vb_init_tables();
cudaError_t vb_launch_jacobian(const double *e, const double *n, long long npts, const double *fe,
                               const double *fn, long long m, double mindist, double *jac,
                               cudaStream_t stream);
cudaError_t vb_launch_jacobian2d(const double *e, const double *n, long long npts,
                                 const double *fe, const double *fn, long long m, double mindist,
                                 double poisson, double *jac, cudaStream_t stream);
void vb_predict_plan(long long npts, long long m, int *pts, int *nsplit);
cudaError_t vb_launch_predict(const double *e, const double *n, long long npts, const double *fe,
                              const double *fn, const double *f, long long m, double mindist,
                              int exact, int pts, int nsplit, double *scratch, double *out,
                              cudaStream_t stream);
cudaError_t vb_launch_predict_multi(const double *e, const double *n, long long npts, const double *fe,
                                    const double *fn, const double *f, long long m, int nrhs, double mindist,
                                    int exact, int nsplit, double *scratch, double *out, cudaStream_t stream);
cudaError_t vb_launch_predict2d(const double *e, const double *n, long long npts, const double *fe,
                                const double *fn, const double *f, long long m, double mindist,
                                double poisson, int exact, int pts, int nsplit, double *scratch,
                                double *out_e, double *out_n, cudaStream_t stream);

// Gram panel generator (vb_gram.cu)
//   kind 0: scalar spline     (columns = m forces,   rows = n data)
//   kind 1: elastic 2-D spline (columns = 2m,         rows = 2n; block structure of
//           verde/vector.py:461-475)
//   kind 2: explicit Jacobian  (verde.base.least_squares on a caller-built matrix)
struct VbGramProblem {
    int kind;
    const double *jac;           // kind 2 only: explicit row-major (n, m) Jacobian
    const double *east, *north;  // n data coordinates
    const double *data;          // n (kind 0) or 2n (kind 1: east comps then north comps)
    const double *weights;       // same length as data, or null
    long long n;                 // data points
    const double *fe, *fn;       // m force coordinates
    long long m;                 // forces
    double mindist, poisson;
};
// rows/cols of the least-squares system
static inline long long vb_sys_rows(const VbGramProblem &p) { return p.kind == 1 ? 2 * p.n : p.n; }
static inline long long vb_sys_cols(const VbGramProblem &p) { return p.kind == 1 ? 2 * p.m : p.m; }

// Generates rows [row0, row0 + nchunk*128) of sqrt(W) J as k-major chunks into
// `panel` (layout [chunk][T][128][128]) and ADDS this slab's contribution to
// stats[0..3][Mpad] = {J^T W d, J^T 1, diag(J^T J), unused}.
cudaError_t vb_launch_column_scale(const double *stats, int T, long long cols, long long nrows,
                                   double *inv_scale, double *rhs, cudaStream_t stream);
cudaError_t vb_launch_scale_system(double *G, int T, long long cols, const double *inv_scale,
                                   double damping, cudaStream_t stream);
cudaError_t vb_launch_gram_panel(const VbGramProblem &prob, long long row0, int nchunk, int T,
                                 double *panel, double *stats, cudaStream_t stream);

cudaError_t vb_launch_sqrt(const double *w, long long n, double *out, cudaStream_t stream);

// Fused in-shared-memory Gram variant (vb_gemm.cu): the Jacobian rows are GENERATED into the
// DMMA operand stages by four generator warps instead of being streamed from a panel.
struct VbFusedGramParams {
    const double *east, *north;  // n data coordinates
    const double *sqrt_w;        // n, or null
    long long n;
    const double *fe, *fn;       // m force coordinates
    long long m;
    double mindist;
    double *c_base;
    int T;
    int accumulate;
    int strip;
};
cudaError_t vb_launch_fused_gram(const VbFusedGramParams &p, cudaStream_t stream);

// Cholesky and solves on tile-packed storage (vb_chol.cu)
// batch matrices of stride `stride` doubles are factored side by side (d_info: one int per matrix)
// lookahead_sms > 0: factor panel p+1 on that many SMs beside the trailing update of panel p (vb_chol.cu)
cudaError_t vb_cholesky_tiled(double *L, int T, int panel_tiles, int *d_info, cudaStream_t stream,
                              int batch = 1, long long stride = 0, int lookahead_sms = 0);
extern int vb_chol_lookahead_min_rows;
// the two halves of one outer step of vb_cholesky_tiled, exposed for the multi-GPU panel-cyclic
// factorisation (the owner of a panel factors it, everybody updates the columns it owns)
cudaError_t vb_chol_factor_panel(double *L, int T, int k0, int pw, int *d_info, cudaStream_t stream,
                                 int batch = 1, long long stride = 0);
cudaError_t vb_chol_trailing_update(double *L, int T, int k0, int pw, int j_lo, int j_hi,
                                    cudaStream_t stream, int batch = 1, long long stride = 0);
// the same for several block-column ranges in ONE launch of the dynamic kernel (the panels a rank owns)
cudaError_t vb_chol_trailing_update_ranges(double *L, int T, int k0, int pw, int nranges, const int *j_lo,
                                           const int *j_hi, cudaStream_t stream);
// park `stream` until *flag >= value (written by a peer GPU); a timeout stores -1 in *d_info
cudaError_t vb_chol_wait_flag(const long long *flag, long long value, int *d_info, cudaStream_t stream);
// offset (doubles) of the diagonal tile of block column c: columns [c0, c1) occupy the contiguous
// range [vb_chol_column_offset(c0), vb_chol_column_offset(c1)) of the tile-packed triangle
long long vb_chol_column_offset(int c, int T);
// flags: 2*T*batch ints of device scratch for the single-launch sweeps (null -> per-block-row kernels);
// x holds batch right-hand sides of T*128 doubles each
cudaError_t vb_solve_tiled(const double *L, int T, double *x, int *flags, cudaStream_t stream,
                           int batch = 1, long long stride = 0);
// L_b = D G D + damping_b I for b < batch (reads G once)
cudaError_t vb_launch_scale_system_batched(const double *G, int T, long long cols, const double *inv_scale,
                                           const double *dampings_dev, int batch, double *L,
                                           long long stride, cudaStream_t stream);
extern int vb_solve_variant, vb_trsm_tiles;

// ---- per-fit diagnostics (vb_diag.cu) ----------------------------------------------------------------
struct VbCondEstimate {
    double lambda_max, lambda_min;  // Rayleigh-quotient estimates of the extreme eigenvalues of L L^T
    int power_its, inverse_its;
};
long long vb_diag_scratch_doubles(int T);
// flags: 2*T ints (the sweeps' scratch); synchronises the stream (early stopping is decided on the host)
cudaError_t vb_cond_estimate(const double *L, int T, long long cols, double *scratch, int *flags, cudaStream_t st,
                             int max_power, int max_inverse, const double *x0, VbCondEstimate *out);
cudaError_t vb_launch_weighted_residual(const double *d, const double *pred, const double *w, long long n, double *r,
                                        cudaStream_t st);
// out3 = {||D grad - damping c||^2, ||c||^2, ||D J^T W d||^2} with c = force / inv_scale
cudaError_t vb_launch_backward_norms(const double *grad, const double *jtwd, const double *inv_scale,
                                     const double *force, long long cols, double damping, double *out3,
                                     cudaStream_t st);
// explicit row-major (n, m) Jacobian times a vector (transpose = 0) or its transpose times a vector
cudaError_t vb_launch_gemv(const double *jac, long long n, long long m, const double *v, int transpose, double *out,
                           cudaStream_t st);

// ---- damping=None: truncated minimum-norm least squares by one-sided Jacobi SVD (vb_svd.cu) ----------
struct VbSvdResult {
    long long rank;       // singular values kept: sigma > rcond * sigma_max
    double sigma_max, sigma_min_kept;
    int sweeps, converged;
    long long rotations;
};
long long vb_svd_stack_doubles(long long rows, long long cols);
cudaError_t vb_svd_inv_scale(const double *jac, long long rows, long long cols, long long Mpad, double *inv_scale,
                             cudaStream_t st);
cudaError_t vb_svd_solve(const double *jac, const double *data, const double *weights, const double *inv_scale,
                         long long rows, long long cols, double rcond, int max_sweeps, double *A, double *work,
                         unsigned *counter, double *params, VbSvdResult *res, cudaStream_t st);

// ---- block reductions and the distance mask (vb_block.cu) --------------------------------------
// reduction codes of vb200_block_aggregate (mirrored in include/verde_b200.h)
#define VB_REDUCE_MEAN 0
#define VB_REDUCE_SUM 1
#define VB_REDUCE_MIN 2
#define VB_REDUCE_MAX 3
#define VB_REDUCE_COUNT 4
#define VB_REDUCE_MEDIAN 5
#define VB_REDUCE_VAR 6        // sample variance, ddof = 1 (pandas "var")
#define VB_REDUCE_WMEAN 7      // sum(w v) / sum(w)          (numpy.average(values, weights=w))
#define VB_REDUCE_WVAR 8       // sum(w (v - wmean)^2) / sum(w)
#define VB_REDUCE_INV_SUM_W 9  // 1 / sum(w)                 (propagated uncertainty of a weighted mean)
#define VB_REDUCE_VAR0 10      // population variance, ddof = 0 (numpy.var as pandas >= 3 applies it)
#define VB_REDUCE_STD 11       // sample standard deviation, ddof = 1 (pandas "std")
#define VB_REDUCE_STD0 12      // population standard deviation, ddof = 0 (numpy.std)
#define VB_REDUCE_LAST VB_REDUCE_STD0

struct VbBlockWorkspace {
    void *keys[2];                 // (label, point index) pairs, ping-pong (uint32 labels; sized for uint64)
    unsigned *vals[2];
    void *mkeys[2];                // second pair of buffers for the (label, value) order of the median
    unsigned *mvals[2];
    unsigned *counts;              // 256 x sort tiles
    unsigned *scratch;             // scan scratch
    unsigned *flags;               // n
    unsigned *seg_start;           // nseg + 1
    long long *seg_label;          // nseg
    unsigned *nseg_dev;            // 1
    int sorted;                    // which of keys/vals holds the label-sorted order
    int label_bits;
};
long long vb_block_scan_scratch_elems(long long n);
cudaError_t vb_block_split(const double *east, const double *north, long long n, const double *ce, int n_east,
                           const double *cn, int n_north, VbBlockWorkspace &ws, long long *labels_dev,
                           cudaStream_t st);
cudaError_t vb_block_heads(long long n, long long nseg, VbBlockWorkspace &ws, cudaStream_t st);
// order-free one-pass reduction (sum / mean / count / weighted mean) into dense per-block accumulators; *nseg_dev
// receives the number of non-empty blocks, ids / out their labels and values [ncol][capacity]
cudaError_t vb_block_stream(const double *east, const double *north, long long n, const double *ce, int n_east,
                            const double *cn, int n_north, const double *values, int ncol, const double *weights,
                            int op, unsigned *cnt, double *sums, double *sumw, unsigned *flags, unsigned *scratch,
                            unsigned *nseg_dev, long long capacity, long long *ids, double *out, cudaStream_t st);
cudaError_t vb_block_aggregate(const double *values, const double *weights, long long n, long long nseg, int op,
                               const long long *labels_dev, VbBlockWorkspace &ws, double *out, cudaStream_t st);
cudaError_t vb_distance_mask(const double *data_e, const double *data_n, long long n, double x0, double y0,
                             double pitch, int nx, int ny, unsigned *cell_of, unsigned *cell_start,
                             unsigned *cursor, unsigned *scratch, double *sorted_xy, const double *qe,
                             const double *qn, long long nq, double maxdist, unsigned char *mask, cudaStream_t st);
