// vb_svd.cu -- the reference's damping=None branch: truncated minimum-norm least squares.
//
// verde/base/least_squares.py:65-70 with damping=None builds
//   LinearRegression(fit_intercept=False).fit(Js, d, sample_weight=w)
// which (sklearn/linear_model/_base.py, scikit-learn 1.9) rescales rows by sqrt(w) and calls
//   scipy.linalg.lstsq(X, y, cond=tol)          tol = 1e-6 (the estimator's default)
// i.e. LAPACK gelsd: the minimum-norm solution with every singular value <= cond * sigma_max dropped.  On the
// spline's Jacobians that truncation is the regulariser (rank 210 of 240, 270 of 1500 on the goldens), so the
// solution is only reproducible if the singular values NEAR the cut-off (1e-6 sigma_max) are right to a small
// fraction of their 1 % spacing.  The normal equations cannot deliver that (eigenvalues of Js^T Js at
// 1e-12 lambda_max carry ~m eps / 1e-12 relative noise), so this path does not touch the Gram machinery:
//
//   one-sided Jacobi (Hestenes) on the stacked matrix [B; I], B = sqrt(W) J D:
//     every rotation orthogonalises two columns of B and applies the same rotation to the identity below, so at
//     convergence the top block is U Sigma and the bottom block is V.  Column norms are the singular values,
//     accurate to eps * sigma_max or better; x = V Sigma^+ U^T d follows column by column.
//
//   * pairs are visited in round-robin (tournament) order: m - 1 rounds of m / 2 DISJOINT pairs per sweep, one
//     CTA per pair, one launch per round; a CTA keeps its two columns in shared memory (<= 13 000 rows) between
//     the dot-product pass and the rotation pass;
//   * a pair is rotated when |a_p . a_q| > max(rows, cols) eps ||a_p|| ||a_q|| (LAPACK dgesvj's criterion, CTOL = M)
//     unless BOTH columns are below 1e-9 of the largest column: those belong to the discarded part of the spectrum
//     (the cut is at 1e-6) and their mutual orthogonality cannot matter;
//   * strongly graded pairs (norm ratio < sqrt(eps)) update the small column only (dgesvj's graded branch), so
//     numerically-zero columns -- every rank-deficient spline Jacobian has them -- cannot keep the large ones rotating;
//   * sweeps repeat until one passes without a rotation.
//
// Bound: HBM/L2 streaming of the column pairs (each round reads and writes the whole stacked matrix once):
// 16 m^2 (rows + m) bytes per sweep, ~20 sweeps on the spline Jacobians.  This branch is for the reference's
// default constructor on modest systems; the damped Cholesky path is the production one.
#include <math.h>

#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

constexpr int JAC_THREADS = 256;

// pair of round `r` for CTA `k` of mp/2 (mp even): players 0..mp-2 rotate, player mp-1 stays
__device__ __forceinline__ void tournament_pair(int mp, int r, int k, int &p, int &q)
{
    int a, b;
    if (k == 0) {
        a = mp - 1;
        b = r;
    } else {
        a = (r + k) % (mp - 1);
        b = (r - k + (mp - 1)) % (mp - 1);
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    // fixed tree: warp shuffles, then the first warp over the per-warp results
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < (JAC_THREADS >> 5)) ? sh[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

// One round of the sweep.  A is column-major, leading dimension L = rows + cols; dots over the first `rows`.
template <bool CACHE>
__global__ void __launch_bounds__(JAC_THREADS)
vb_jacobi_round_kernel(double *__restrict__ A, long long L, long long rows, int mp, int round, double tol,
                       const double *__restrict__ tiny2_ptr, unsigned *__restrict__ rotations)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[JAC_THREADS / 32];
    int p, q;
    tournament_pair(mp, round, blockIdx.x, p, q);
    double *ap = A + (long long)p * L, *aq = A + (long long)q * L;
    double *sp = sm, *sq = sm + L;
    double al = 0.0, be = 0.0, ga = 0.0;
    for (long long i = threadIdx.x; i < rows; i += JAC_THREADS) {
        const double x = ap[i], y = aq[i];
        if (CACHE) {
            sp[i] = x;
            sq[i] = y;
        }
        al = fma(x, x, al);
        be = fma(y, y, be);
        ga = fma(x, y, ga);
    }
    al = block_sum(al, red);
    be = block_sum(be, red);
    ga = block_sum(ga, red);
    const double tiny2 = *tiny2_ptr;
    if (!(al > 0.0) || !(be > 0.0)) return;                  // a zero (or padding) column
    if (al <= tiny2 && be <= tiny2) return;                  // both far below the truncation level
    if (!(fabs(ga) > tol * sqrt(al) * sqrt(be))) return;     // already orthogonal
    if (threadIdx.x == 0) atomicAdd(rotations, 1u);
    // Strongly graded pair (norm ratio below sqrt(eps)): a rotation would change the large column only by
    // rounding noise -- thousands of such no-op updates per sweep are what keeps numerically-zero columns from
    // ever settling.  Like LAPACK dgesvj's graded branch, only the SMALL column is updated (a Gram-Schmidt step:
    // the rotation with its O(ratio^2) <= eps effect on the large column dropped).
    const double eps = 2.220446049250313e-16;
    if (be < eps * al || al < eps * be) {
        const bool p_large = be < al;
        const double mu = p_large ? ga / al : ga / be;
        for (long long i = threadIdx.x; i < L; i += JAC_THREADS) {
            double x, y;
            if (CACHE && i < rows) {
                x = sp[i];
                y = sq[i];
            } else {
                x = ap[i];
                y = aq[i];
            }
            if (p_large) aq[i] = fma(-mu, x, y);
            else ap[i] = fma(-mu, y, x);
        }
        return;
    }
    // Rutishauser's formulas: t = tan of the rotation angle, |t| <= 1
    const double zeta = (be - al) / (2.0 * ga);
    const double t = (zeta == 0.0) ? 1.0 : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
    for (long long i = threadIdx.x; i < L; i += JAC_THREADS) {
        double x, y;
        if (CACHE && i < rows) {
            x = sp[i];
            y = sq[i];
        } else {
            x = ap[i];
            y = aq[i];
        }
        ap[i] = c * x - s * y;
        aq[i] = s * x + c * y;
    }
}

// A[c * L + r] = sqrt(w[r]) * jac[r * cols + c] * inv_scale[c]   (r < rows);  identity below; padding columns zero
__global__ void __launch_bounds__(256)
vb_svd_stack_kernel(const double *__restrict__ jac, const double *__restrict__ w, const double *__restrict__ inv_scale,
                    long long rows, long long cols, int mp, long long L, double *__restrict__ A)
{
    // 32 x 32 transposing tiles through shared memory: coalesced reads along c, coalesced writes along r
    __shared__ double tile[32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of threads
    for (int k = ty; k < 32; k += 8) {
        const long long r = r0 + k, c = c0 + tx;
        double v = 0.0;
        if (r < rows && c < cols) v = jac[r * cols + c] * inv_scale[c] * (w ? sqrt(w[r]) : 1.0);
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const long long c = c0 + k, r = r0 + tx;
        if (c < mp && r < rows) A[c * L + r] = tile[tx][k];
    }
}
__global__ void vb_svd_identity_kernel(long long rows, long long cols, int mp, long long L, double *__restrict__ A)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)mp * cols) return;
    const long long c = idx / cols, i = idx - c * cols;
    A[c * L + rows + i] = (c == i) ? 1.0 : 0.0;
}

// per column: sigma^2 = ||a_c||^2 and a_c . dw over the first `rows` rows   (dw = sqrt(w) d)
__global__ void __launch_bounds__(JAC_THREADS)
vb_svd_column_kernel(const double *__restrict__ A, long long L, long long rows, const double *__restrict__ d,
                     const double *__restrict__ w, double *__restrict__ sig2, double *__restrict__ proj)
{
    __shared__ double red[JAC_THREADS / 32];
    const double *a = A + (long long)blockIdx.x * L;
    double s2 = 0.0, pr = 0.0;
    for (long long i = threadIdx.x; i < rows; i += JAC_THREADS) {
        const double x = a[i];
        s2 = fma(x, x, s2);
        if (d) pr = fma(x, d[i] * (w ? sqrt(w[i]) : 1.0), pr);
    }
    s2 = block_sum(s2, red);
    pr = block_sum(pr, red);
    if (threadIdx.x == 0) {
        sig2[blockIdx.x] = s2;
        if (proj) proj[blockIdx.x] = pr;
    }
}

// out[0] = max sig2, out[1] = tiny2 = (1e-9)^2 * max sig2   (single CTA)
__global__ void __launch_bounds__(1024)
vb_svd_max_kernel(const double *__restrict__ sig2, int mp, double *__restrict__ out)
{
    __shared__ double s[1024];
    double v = 0.0;
    for (int j = threadIdx.x; j < mp; j += blockDim.x) v = fmax(v, sig2[j]);
    s[threadIdx.x] = v;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s[0];
        out[1] = 1e-18 * s[0];
    }
}

// coefficient of column j in x = V Sigma^+ U^T dw: proj_j / sigma_j^2 when sigma_j > rcond sigma_max, else 0;
// out3 = {rank, min kept sigma^2, max sigma^2}
__global__ void __launch_bounds__(1024)
vb_svd_truncate_kernel(const double *__restrict__ sig2, const double *__restrict__ proj, int mp, double rcond,
                       const double *__restrict__ max_ptr, double *__restrict__ coef, double *__restrict__ out3)
{
    __shared__ double smin[1024];
    __shared__ int scount[1024];
    const double smax2 = max_ptr[0];
    const double cut2 = rcond * rcond * smax2;
    double lo = 1e300;
    int count = 0;
    for (int j = threadIdx.x; j < mp; j += blockDim.x) {
        const double s2 = sig2[j];
        const bool keep = s2 > cut2 && s2 > 0.0;  // sigma > rcond * sigma_max  (gelsd drops sigma <= rcond * sigma_max)
        coef[j] = keep ? proj[j] / s2 : 0.0;
        if (keep) {
            lo = fmin(lo, s2);
            ++count;
        }
    }
    smin[threadIdx.x] = lo;
    scount[threadIdx.x] = count;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + st]);
            scount[threadIdx.x] += scount[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out3[0] = (double)scount[0];
        out3[1] = smin[0];
        out3[2] = smax2;
    }
}

// params[i] = inv_scale[i] * sum_j V(i, j) coef[j],  V(i, j) = A[j * L + rows + i]
__global__ void __launch_bounds__(256)
vb_svd_expand_kernel(const double *__restrict__ A, long long L, long long rows, long long cols, int mp,
                     const double *__restrict__ coef, const double *__restrict__ inv_scale, double *__restrict__ params)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cols) return;
    const double *v = A + rows + i;
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
    for (; j + 1 < mp; j += 2) {
        a0 = fma(v[(long long)j * L], coef[j], a0);
        a1 = fma(v[(long long)(j + 1) * L], coef[j + 1], a1);
    }
    if (j < mp) a0 = fma(v[(long long)j * L], coef[j], a0);
    params[i] = (a0 + a1) * inv_scale[i];
}

// StandardScaler(with_mean=False).scale_ of an explicit row-major Jacobian, as 1 / scale: population standard
// deviation about the column mean by the corrected two-pass formula (sklearn/utils/extmath.py:1203-1240) and the
// constant-column rule of sklearn/preprocessing/_data.py:83-133.  One thread per column, coalesced across columns.
__global__ void __launch_bounds__(256)
vb_svd_inv_scale_kernel(const double *__restrict__ jac, long long rows, long long cols, long long Mpad,
                        double *__restrict__ inv_scale)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Mpad) return;
    if (c >= cols) {
        inv_scale[c] = 1.0;
        return;
    }
    const double n = (double)rows;
    double s = 0.0;
    for (long long r = 0; r < rows; ++r) s += jac[r * cols + c];
    const double mean = s / n;
    double q = 0.0, corr = 0.0;
    for (long long r = 0; r < rows; ++r) {
        const double t = jac[r * cols + c] - mean;
        q = fma(t, t, q);
        corr += t;
    }
    q -= corr * corr / n;
    double var = q / n;
    const double eps = 2.220446049250313e-16;
    const double nm = n * mean * eps;
    const bool constant = var <= n * eps * var + nm * nm;
    double scale = constant ? 0.0 : sqrt(var);
    if (constant || scale < 10.0 * eps) scale = 1.0;
    inv_scale[c] = 1.0 / scale;
}

}  // namespace

long long vb_svd_stack_doubles(long long rows, long long cols)
{
    const long long mp = cols + (cols & 1);
    return (rows + cols) * mp;
}

cudaError_t vb_svd_inv_scale(const double *jac, long long rows, long long cols, long long Mpad, double *inv_scale,
                             cudaStream_t st)
{
    vb_svd_inv_scale_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, st>>>(jac, rows, cols, Mpad, inv_scale);
    ++vb_launch_counter;
    return cudaGetLastError();
}

// jac: row-major (rows, cols) on the device; A: vb_svd_stack_doubles(rows, cols) doubles of scratch;
// work: 3 * mp + 8 doubles; counter: one unsigned.  Synchronises the stream once per sweep.
cudaError_t vb_svd_solve(const double *jac, const double *data, const double *weights, const double *inv_scale,
                         long long rows, long long cols, double rcond, int max_sweeps, double *A, double *work,
                         unsigned *counter, double *params, VbSvdResult *res, cudaStream_t st)
{
    const int mp = (int)(cols + (cols & 1));
    const long long L = rows + cols;
    double *sig2 = work, *proj = sig2 + mp, *coef = proj + mp, *sc = coef + mp;  // sc[0] = max sig2, sc[1] = tiny2, sc[2..4] = out3
    cudaError_t e;
    // ---- stack [sqrt(W) J D; I] column-major ----
    if ((e = cudaMemsetAsync(A, 0, sizeof(double) * (size_t)L * mp, st)) != cudaSuccess) return e;
    vb_svd_stack_kernel<<<dim3((unsigned)((mp + 31) / 32), (unsigned)((rows + 31) / 32)), 256, 0, st>>>(
        jac, weights, inv_scale, rows, cols, mp, L, A);
    vb_svd_identity_kernel<<<(unsigned)(((long long)mp * cols + 255) / 256), 256, 0, st>>>(rows, cols, mp, L, A);
    vb_launch_counter += 2;
    // ---- sweeps ----
    // LAPACK dgesvj's default: CTOL = number of rows, a pair counts as orthogonal below CTOL * eps (here the longer
    // side, because the under-determined case has more columns than rows)
    const double tol = (double)(rows > cols ? rows : cols) * 2.220446049250313e-16;
    const size_t cache_bytes = sizeof(double) * 2 * (size_t)L;
    const bool cache = cache_bytes <= 200 * 1024;
    if (cache) {
        if ((e = cudaFuncSetAttribute(vb_jacobi_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cache_bytes)) != cudaSuccess)
            return e;
    }
    res->sweeps = 0;
    res->converged = 0;
    res->rotations = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        vb_svd_column_kernel<<<mp, JAC_THREADS, 0, st>>>(A, L, rows, nullptr, nullptr, sig2, nullptr);
        vb_svd_max_kernel<<<1, 1024, 0, st>>>(sig2, mp, sc);
        vb_launch_counter += 2;
        if ((e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st)) != cudaSuccess) return e;
        for (int round = 0; round < mp - 1; ++round) {
            if (cache)
                vb_jacobi_round_kernel<true><<<mp / 2, JAC_THREADS, cache_bytes, st>>>(A, L, rows, mp, round, tol, sc + 1, counter);
            else
                vb_jacobi_round_kernel<false><<<mp / 2, JAC_THREADS, 0, st>>>(A, L, rows, mp, round, tol, sc + 1, counter);
        }
        vb_launch_counter += mp - 1;
        unsigned h = 0;
        if ((e = cudaMemcpyAsync(&h, counter, sizeof(unsigned), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        res->sweeps = sweep + 1;
        res->rotations += h;
        if (h == 0) {
            res->converged = 1;
            break;
        }
    }
    // ---- x = V Sigma^+ U^T (sqrt(W) d), unscaled ----
    vb_svd_column_kernel<<<mp, JAC_THREADS, 0, st>>>(A, L, rows, data, weights, sig2, proj);
    vb_svd_max_kernel<<<1, 1024, 0, st>>>(sig2, mp, sc);
    vb_svd_truncate_kernel<<<1, 1024, 0, st>>>(sig2, proj, mp, rcond, sc, coef, sc + 2);
    vb_svd_expand_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(A, L, rows, cols, mp, coef, inv_scale, params);
    vb_launch_counter += 4;
    double h3[3] = {0, 0, 0};
    if ((e = cudaMemcpyAsync(h3, sc + 2, sizeof(h3), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    res->rank = (long long)h3[0];
    res->sigma_min_kept = h3[0] > 0 ? sqrt(h3[1]) : 0.0;
    res->sigma_max = sqrt(h3[2]);
    return cudaGetLastError();
}
