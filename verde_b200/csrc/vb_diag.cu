// vb_diag.cu -- per-fit diagnostics: condition-number estimate of the damped normal matrix and the
// normwise backward error of the solved forces (out_diag{cond, backward_err}).
//
// north_star asks for results "with the condition number stated"; the reference itself never computes it
// (scipy.linalg.solve(assume_a="pos") only warns through LAPACK's rcond), so there is nothing to match
// bit for bit here -- tests pin the estimate against numpy.linalg.eigvalsh on the golden systems.
//
// * cond_est = lambda_max / lambda_min of A = D G D + damping I, both from the Cholesky factor that every
//   solve path leaves in HBM (A = L L^T up to the factorisation's backward error):
//     lambda_max -- power iteration on L L^T: two triangular matrix-vector products per step, each ONE
//                   streaming pass over the tile-packed factor (HBM-bound, 10 GB at m = 50k);
//     lambda_min -- inverse iteration: the existing single-launch triangular sweeps (vb_solve_tiled).
//   The Rayleigh quotients approach the extreme eigenvalues from inside, so cond_est <= cond(A); since
//   G is positive semi-definite, lambda_max / damping is the matching upper bound (reported beside it).
// * backward error -- eta = || Js^T W (d - Js c) - damping c || / (lambda_max ||c|| + ||Js^T W d||) for the
//   scaled unknowns c = force / inv_scale: the residual of the normal equations the reference solves,
//   evaluated MATRIX-FREE with the exact Green's function (J f through the predict kernel, J^T r through
//   the same kernel with points and forces swapped -- every Green's function of the path is even), so it
//   also sees the rounding of the Gram accumulation, not just the triangular solves.
#include <math.h>

#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

// partial[t][c] = sum_r L_t(r, c) x[bi*128 + r]     (TRANS:  y = L^T x, reduced over block columns)
// partial[t][r] = sum_c L_t(r, c) y[bj*128 + c]     (!TRANS: z = L y,   reduced over block rows)
// grid (T, T): blockIdx.x = bi, blockIdx.y = bj; the upper half exits at once.
template <bool TRANS>
__global__ void __launch_bounds__(256)
vb_tri_partial_kernel(const double *__restrict__ Lmat, int T, const double *__restrict__ v, double *__restrict__ partial)
{
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi < bj) return;
    const long long t = vb_tile_index(bi, bj, T);
    const double *tile = Lmat + t * VB_TILE_ELEMS;
    double *out = partial + t * VB_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (TRANS) {
        // warp w owns columns 16w .. 16w+15; a lane holds rows 2*lane + 64q (q = 0, 1) as double2
        const double2 x0 = *reinterpret_cast<const double2 *>(v + (long long)bi * VB_TILE + 2 * lane);
        const double2 x1 = *reinterpret_cast<const double2 *>(v + (long long)bi * VB_TILE + 64 + 2 * lane);
        double acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double *col = tile + (warp * 16 + c) * VB_TILE;
            const double2 a0 = __ldcs(reinterpret_cast<const double2 *>(col + 2 * lane));
            const double2 a1 = __ldcs(reinterpret_cast<const double2 *>(col + 64 + 2 * lane));
            acc[c] = fma(a0.x, x0.x, fma(a0.y, x0.y, fma(a1.x, x1.x, a1.y * x1.y)));
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double s = acc[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) out[warp * 16 + c] = s;
        }
    } else {
        __shared__ double sy[VB_TILE];
        __shared__ double shalf[VB_TILE];
        if (tid < VB_TILE) sy[tid] = v[(long long)bj * VB_TILE + tid];
        __syncthreads();
        const int r = tid & (VB_TILE - 1), h = tid >> 7;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
        for (int c = h * 64; c < h * 64 + 64; c += 4) {
            a0 = fma(__ldcs(tile + (c + 0) * VB_TILE + r), sy[c + 0], a0);
            a1 = fma(__ldcs(tile + (c + 1) * VB_TILE + r), sy[c + 1], a1);
            a2 = fma(__ldcs(tile + (c + 2) * VB_TILE + r), sy[c + 2], a2);
            a3 = fma(__ldcs(tile + (c + 3) * VB_TILE + r), sy[c + 3], a3);
        }
        const double s = (a0 + a1) + (a2 + a3);
        if (h == 1) shalf[r] = s;
        __syncthreads();
        if (h == 0) out[r] = s + shalf[r];
    }
}

// y[bj] = sum_{bi >= bj} partial[tile(bi, bj)]   (TRANS)   /   z[bi] = sum_{bj <= bi} partial[tile(bi, bj)]
template <bool TRANS>
__global__ void __launch_bounds__(VB_TILE)
vb_tri_reduce_kernel(const double *__restrict__ partial, int T, double *__restrict__ out)
{
    const int b = blockIdx.x, r = threadIdx.x;
    double s = 0.0;
    if (TRANS) {
        for (int bi = b; bi < T; ++bi) s += partial[vb_tile_index(bi, b, T) * VB_TILE + r];
    } else {
        for (int bj = 0; bj <= b; ++bj) s += partial[vb_tile_index(b, bj, T) * VB_TILE + r];
    }
    out[(long long)b * VB_TILE + r] = s;
}

// deterministic start vectors: kind 0 = ones, kind 1 = +-1 from a fixed integer hash; padding stays 0
__global__ void vb_diag_start_kernel(double *__restrict__ x, long long cols, long long Mpad, int kind)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    double v = 0.0;
    if (j < cols) {
        if (kind == 0) v = 1.0;
        else {
            unsigned h = (unsigned)j * 2654435761u;
            h ^= h >> 15;
            h *= 2246822519u;
            h ^= h >> 13;
            v = (h & 1u) ? 1.0 : -1.0;
        }
    }
    x[j] = v;
}

// out[0] = a . b, out[1] = a . a  (single CTA, fixed order -> deterministic)
__global__ void __launch_bounds__(1024)
vb_dot2_kernel(const double *__restrict__ a, const double *__restrict__ b, long long n, double *__restrict__ out)
{
    __shared__ double s0[1024], s1[1024];
    double d0 = 0.0, d1 = 0.0;
    for (long long j = threadIdx.x; j < n; j += blockDim.x) {
        const double av = a[j];
        d0 = fma(av, b[j], d0);
        d1 = fma(av, av, d1);
    }
    s0[threadIdx.x] = d0;
    s1[threadIdx.x] = d1;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) {
            s0[threadIdx.x] += s0[threadIdx.x + st];
            s1[threadIdx.x] += s1[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s0[0];
        out[1] = s1[0];
    }
}

// dst = src / sqrt(norm2[1])
__global__ void vb_normalise_kernel(const double *__restrict__ src, const double *__restrict__ norm2, long long n,
                                    double *__restrict__ dst)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double nn = norm2[1];
    dst[j] = nn > 0.0 ? src[j] * rsqrt(nn) : 0.0;
}

// r[i] = w[i] * (d[i] - pred[i])
__global__ void vb_weighted_residual_kernel(const double *__restrict__ d, const double *__restrict__ pred,
                                            const double *__restrict__ w, long long n, double *__restrict__ r)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = (w ? w[i] : 1.0) * (d[i] - pred[i]);
}

// Normal-equation residual in the scaled unknowns (single CTA):
//   c_j = f_j / inv_j,  res_j = inv_j * grad_j - damping * c_j,  rhs_j = inv_j * jtwd_j
// out = {||res||^2, ||c||^2, ||rhs||^2}
__global__ void __launch_bounds__(1024)
vb_backward_norms_kernel(const double *__restrict__ grad, const double *__restrict__ jtwd,
                         const double *__restrict__ inv_scale, const double *__restrict__ force, long long cols,
                         double damping, double *__restrict__ out3)
{
    __shared__ double s[3][1024];
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long j = threadIdx.x; j < cols; j += blockDim.x) {
        const double inv = inv_scale[j];
        const double cj = force[j] / inv;
        const double res = inv * grad[j] - damping * cj;
        const double rh = inv * jtwd[j];
        a = fma(res, res, a);
        b = fma(cj, cj, b);
        c = fma(rh, rh, c);
    }
    s[0][threadIdx.x] = a;
    s[1][threadIdx.x] = b;
    s[2][threadIdx.x] = c;
    __syncthreads();
    for (int st = 512; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st)
            for (int q = 0; q < 3; ++q) s[q][threadIdx.x] += s[q][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x < 3) out3[threadIdx.x] = s[threadIdx.x][0];
}

// explicit-Jacobian matvecs for vb200_least_squares (kind 2), row-major (n, m)
__global__ void __launch_bounds__(256)
vb_gemv_rows_kernel(const double *__restrict__ jac, long long n, long long m, const double *__restrict__ f,
                    double *__restrict__ out)
{
    // one warp per row
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (long long j = lane; j < m; j += 32) s = fma(jac[row * m + j], f[j], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[row] = s;
}
__global__ void __launch_bounds__(256)
vb_gemv_cols_kernel(const double *__restrict__ jac, long long n, long long m, const double *__restrict__ v,
                    double *__restrict__ out)
{
    // one thread per column, coalesced across the row
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double s = 0.0;
    for (long long i = 0; i < n; ++i) s = fma(jac[i * m + j], v[i], s);
    out[j] = s;
}

cudaError_t tri_matvec(bool trans, const double *L, int T, const double *v, double *partial, double *out,
                       cudaStream_t st)
{
    if (trans) {
        vb_tri_partial_kernel<true><<<dim3(T, T), 256, 0, st>>>(L, T, v, partial);
        vb_tri_reduce_kernel<true><<<T, VB_TILE, 0, st>>>(partial, T, out);
    } else {
        vb_tri_partial_kernel<false><<<dim3(T, T), 256, 0, st>>>(L, T, v, partial);
        vb_tri_reduce_kernel<false><<<T, VB_TILE, 0, st>>>(partial, T, out);
    }
    vb_launch_counter += 2;
    return cudaGetLastError();
}

}  // namespace

long long vb_diag_scratch_doubles(int T)
{
    // partial sums (one 128-vector per tile) + 3 work vectors + 8 scalars
    return vb_num_tiles(T) * VB_TILE + 3LL * T * VB_TILE + 8;
}

// x0 (optional, T*128 doubles, padding zero): start vector of the inverse iteration.  The solve paths pass the
// scaled solution c = A^-1 rhs they just computed -- already one inverse-iteration step away from rhs, and on
// the ill-conditioned spline systems dominated by the smallest eigenvectors -- which saves sweeps over L.
cudaError_t vb_cond_estimate(const double *L, int T, long long cols, double *scratch, int *flags, cudaStream_t st,
                             int max_power, int max_inverse, const double *x0, VbCondEstimate *out)
{
    const long long Mpad = (long long)T * VB_TILE;
    double *partial = scratch;
    double *x = partial + vb_num_tiles(T) * VB_TILE, *y = x + Mpad, *z = y + Mpad, *sc = z + Mpad;
    const unsigned gv = (unsigned)((Mpad + 255) / 256);
    cudaError_t e;
    double h[2];
    out->lambda_max = out->lambda_min = 0.0;
    out->power_its = out->inverse_its = 0;

    // ---- lambda_max: power iteration on L L^T ----
    vb_diag_start_kernel<<<gv, 256, 0, st>>>(x, cols, Mpad, 0);
    vb_dot2_kernel<<<1, 1024, 0, st>>>(x, x, Mpad, sc);
    vb_normalise_kernel<<<gv, 256, 0, st>>>(x, sc, Mpad, x);
    vb_launch_counter += 3;
    double lam = 0.0;
    for (int it = 0; it < max_power; ++it) {
        if ((e = tri_matvec(true, L, T, x, partial, y, st)) != cudaSuccess) return e;   // y = L^T x
        vb_dot2_kernel<<<1, 1024, 0, st>>>(y, y, Mpad, sc);                              // ||y||^2 = x^T A x
        if ((e = tri_matvec(false, L, T, y, partial, z, st)) != cudaSuccess) return e;  // z = L y = A x
        vb_dot2_kernel<<<1, 1024, 0, st>>>(z, z, Mpad, sc + 2);
        vb_normalise_kernel<<<gv, 256, 0, st>>>(z, sc + 2, Mpad, x);
        vb_launch_counter += 3;
        if ((e = cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        const double prev = lam;
        lam = h[1];
        out->power_its = it + 1;
        if (!(lam == lam) || lam <= 0.0) break;
        if (it > 0 && fabs(lam - prev) <= 1e-3 * lam) break;
    }
    out->lambda_max = lam;

    // ---- lambda_min: inverse iteration, (L L^T) z = x by the two triangular sweeps ----
    if (x0) {
        if ((e = cudaMemcpyAsync(x, x0, sizeof(double) * Mpad, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    } else {
        vb_diag_start_kernel<<<gv, 256, 0, st>>>(x, cols, Mpad, 1);
    }
    vb_dot2_kernel<<<1, 1024, 0, st>>>(x, x, Mpad, sc);
    vb_normalise_kernel<<<gv, 256, 0, st>>>(x, sc, Mpad, x);
    vb_launch_counter += 3;
    double mu = 0.0;
    for (int it = 0; it < max_inverse; ++it) {
        if ((e = cudaMemcpyAsync(z, x, sizeof(double) * Mpad, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        if ((e = vb_solve_tiled(L, T, z, flags, st)) != cudaSuccess) return e;  // z = A^-1 x
        vb_dot2_kernel<<<1, 1024, 0, st>>>(z, x, Mpad, sc);                     // {x^T A^-1 x, ||z||^2}
        vb_normalise_kernel<<<gv, 256, 0, st>>>(z, sc, Mpad, x);
        vb_launch_counter += 2;
        if ((e = cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        const double prev = mu;
        mu = h[0];
        out->inverse_its = it + 1;
        if (!(mu == mu) || mu <= 0.0) break;
        if (it > 0 && fabs(mu - prev) <= 1e-2 * mu) break;
    }
    out->lambda_min = mu > 0.0 ? 1.0 / mu : 0.0;
    // The Rayleigh quotient of A itself at the last (normalised) iterate is a second upper bound on lambda_min and
    // one half-step ahead of 1/mu: x^T A x = ||L^T x||^2 costs ONE streaming pass over the factor (a sweep pair costs
    // ten), which is what lets the large systems stop after three inverse iterations.
    if (out->inverse_its > 0 && mu > 0.0) {
        if ((e = tri_matvec(true, L, T, x, partial, y, st)) != cudaSuccess) return e;
        vb_dot2_kernel<<<1, 1024, 0, st>>>(y, y, Mpad, sc);
        ++vb_launch_counter;
        if ((e = cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        if (h[1] > 0.0 && h[1] < out->lambda_min) out->lambda_min = h[1];
    }
    return cudaGetLastError();
}

cudaError_t vb_launch_weighted_residual(const double *d, const double *pred, const double *w, long long n, double *r,
                                        cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    vb_weighted_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, pred, w, n, r);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_backward_norms(const double *grad, const double *jtwd, const double *inv_scale,
                                     const double *force, long long cols, double damping, double *out3,
                                     cudaStream_t st)
{
    vb_backward_norms_kernel<<<1, 1024, 0, st>>>(grad, jtwd, inv_scale, force, cols, damping, out3);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_gemv(const double *jac, long long n, long long m, const double *v, int transpose, double *out,
                           cudaStream_t st)
{
    if (n <= 0 || m <= 0) return cudaSuccess;
    if (transpose) vb_gemv_cols_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(jac, n, m, v, out);
    else vb_gemv_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(jac, n, m, v, out);
    ++vb_launch_counter;
    return cudaGetLastError();
}
