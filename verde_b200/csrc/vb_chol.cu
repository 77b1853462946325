// vb_chol.cu -- K2: blocked right-looking Cholesky on the tile-packed lower
// triangle, plus forward/back substitution.
//
// Replaces scipy.linalg.solve(A, b, assume_a="pos") as reached from
// sklearn/linear_model/_ridge.py:215-227 (LAPACK dpotrf + dpotrs).
//
// Structure (T = block rows of 128, outer panel = `panel_tiles` block columns):
//   for each outer panel [k0, k0+P):
//     for k in panel:  potrf_tile(k)           one CTA, 128x128 in shared memory
//                      trsm_tiles(k)           one CTA per tile below the diagonal,
//                                              thread-per-row substitution (backward
//                                              stable: no explicit inverses)
//                      gemm(K=128)             panel-internal columns (k, k0+P)
//     gemm(K=128*P)                            trailing update, the bulk of m^3/3 flops,
//                                              the DMMA tile kernel of vb_gemm.cu
// A failed pivot (matrix not positive definite) is recorded as its 1-based
// global index in *d_info (first failure wins); later kernels then propagate
// NaN harmlessly and the host reports the index, like LAPACK's info > 0.
#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

constexpr int POTRF_THREADS = 256;
constexpr int POTRF_LD = VB_TILE + 1;

// Factor one diagonal tile in place (lower triangle), column by column.
__global__ void __launch_bounds__(POTRF_THREADS)
vb_potrf_tile_kernel(double *__restrict__ tile, int k, int *__restrict__ d_info)
{
    extern __shared__ double s[];  // s[c * POTRF_LD + r]
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    if (tid == 0) s_fail = 0;
    for (int idx = tid; idx < VB_TILE_ELEMS; idx += POTRF_THREADS) {
        const int c = idx >> 7, r = idx & 127;
        s[c * POTRF_LD + r] = tile[idx];
    }
    __syncthreads();
    for (int c = 0; c < VB_TILE; ++c) {
        const double piv = s[c * POTRF_LD + c];
        if (!(piv > 0.0)) {
            if (tid == 0) {
                s_fail = 1;
                atomicCAS(d_info, 0, k * VB_TILE + c + 1);
            }
        }
        const double l = sqrt(piv);
        const double inv = 1.0 / l;
        __syncthreads();  // everyone has read the pivot
        if (tid == 0) s[c * POTRF_LD + c] = l;
        for (int r = c + 1 + tid; r < VB_TILE; r += POTRF_THREADS) s[c * POTRF_LD + r] *= inv;
        __syncthreads();
        // trailing update of columns c+1..127 (lower part): s[c2][r] -= s[c][r]*s[c][c2]
        const int rem = VB_TILE - 1 - c;
        for (int idx = tid; idx < rem * rem; idx += POTRF_THREADS) {
            const int c2 = c + 1 + idx / rem, r = c + 1 + idx % rem;
            if (r >= c2) s[c2 * POTRF_LD + r] = fma(-s[c * POTRF_LD + r], s[c * POTRF_LD + c2], s[c2 * POTRF_LD + r]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < VB_TILE_ELEMS; idx += POTRF_THREADS) {
        const int c = idx >> 7, r = idx & 127;
        tile[idx] = (r >= c) ? s[c * POTRF_LD + r] : 0.0;
    }
}

// X * L_kk^T = A for the tiles below the diagonal of block column k.
// One CTA per tile, thread r owns row r.  L_kk (column-major) sits in shared
// memory; row blocks of 32 unknowns live in registers:
//   a[c] -= x_prev * L[c][prev]   (prev columns: x re-read from the tile itself)
//   in-block forward substitution, fully unrolled.
constexpr int TRSM_THREADS = 128;
constexpr int TRSM_BLK = 32;

__global__ void __launch_bounds__(TRSM_THREADS)
vb_trsm_tiles_kernel(double *__restrict__ Lmat, int T, int k)
{
    extern __shared__ __align__(16) double sL[];  // sL[c * 128 + r] = L_kk(r, c)
    __shared__ double s_inv[VB_TILE];
    const double *diag = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    double *tile = Lmat + vb_tile_index(k + 1 + blockIdx.x, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    for (int idx = threadIdx.x; idx < VB_TILE_ELEMS / 2; idx += TRSM_THREADS)
        reinterpret_cast<double2 *>(sL)[idx] = reinterpret_cast<const double2 *>(diag)[idx];
    __syncthreads();
    s_inv[r] = 1.0 / sL[r * VB_TILE + r];
    __syncthreads();

    for (int cb = 0; cb < VB_TILE / TRSM_BLK; ++cb) {
        double a[TRSM_BLK];
#pragma unroll
        for (int c = 0; c < TRSM_BLK; ++c) a[c] = tile[(cb * TRSM_BLK + c) * VB_TILE + r];
        // contributions of already-solved columns
        for (int kp = 0; kp < cb * TRSM_BLK; ++kp) {
            const double xk = tile[kp * VB_TILE + r];
            const double2 *lcol = reinterpret_cast<const double2 *>(sL + kp * VB_TILE + cb * TRSM_BLK);
#pragma unroll
            for (int c2 = 0; c2 < TRSM_BLK / 2; ++c2) {
                const double2 l = lcol[c2];
                a[2 * c2] = fma(-xk, l.x, a[2 * c2]);
                a[2 * c2 + 1] = fma(-xk, l.y, a[2 * c2 + 1]);
            }
        }
        // in-block substitution
#pragma unroll
        for (int kk = 0; kk < TRSM_BLK; ++kk) {
            const int kg = cb * TRSM_BLK + kk;
            const double xk = a[kk] * s_inv[kg];
            a[kk] = xk;
#pragma unroll
            for (int c = kk + 1; c < TRSM_BLK; ++c) a[c] = fma(-xk, sL[kg * VB_TILE + cb * TRSM_BLK + c], a[c]);
        }
#pragma unroll
        for (int c = 0; c < TRSM_BLK; ++c) tile[(cb * TRSM_BLK + c) * VB_TILE + r] = a[c];
    }
}

// ---- triangular solves with one right-hand side ---------------------------------
// forward: L y = b.  Step k: (1) y_k = L_kk^{-1} b_k (one CTA), (2) b_i -= L(i,k) y_k.
__global__ void __launch_bounds__(VB_TILE)
vb_fwd_diag_kernel(const double *__restrict__ Lmat, int T, int k, double *__restrict__ x)
{
    __shared__ double sx[VB_TILE];
    const double *tile = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    sx[r] = x[k * VB_TILE + r];
    __syncthreads();
    for (int c = 0; c < VB_TILE; ++c) {
        if (r == c) sx[c] = sx[c] / tile[c * VB_TILE + c];
        __syncthreads();
        if (r > c) sx[r] = fma(-tile[c * VB_TILE + r], sx[c], sx[r]);
        __syncthreads();
    }
    x[k * VB_TILE + r] = sx[r];
}

__global__ void __launch_bounds__(VB_TILE)
vb_fwd_update_kernel(const double *__restrict__ Lmat, int T, int k, double *__restrict__ x)
{
    __shared__ double sy[VB_TILE];
    const int bi = k + 1 + blockIdx.x;
    const double *tile = Lmat + vb_tile_index(bi, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    sy[r] = x[k * VB_TILE + r];
    __syncthreads();
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll 4
    for (int c = 0; c < VB_TILE; c += 4) {
        acc0 = fma(tile[(c + 0) * VB_TILE + r], sy[c + 0], acc0);
        acc1 = fma(tile[(c + 1) * VB_TILE + r], sy[c + 1], acc1);
        acc2 = fma(tile[(c + 2) * VB_TILE + r], sy[c + 2], acc2);
        acc3 = fma(tile[(c + 3) * VB_TILE + r], sy[c + 3], acc3);
    }
    x[bi * VB_TILE + r] -= (acc0 + acc1) + (acc2 + acc3);
}

// backward: L^T x = y.  Step k (descending): (1) x_k = L_kk^{-T} y_k, (2) y_j -= L(k,j)^T x_k, j < k.
__global__ void __launch_bounds__(VB_TILE)
vb_bwd_diag_kernel(const double *__restrict__ Lmat, int T, int k, double *__restrict__ x)
{
    __shared__ double sx[VB_TILE];
    const double *tile = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    sx[r] = x[k * VB_TILE + r];
    __syncthreads();
    for (int c = VB_TILE - 1; c >= 0; --c) {
        if (r == c) sx[c] = sx[c] / tile[c * VB_TILE + c];
        __syncthreads();
        // (L^T)(r, c) = L(c, r), r < c: stored at tile[r * 128 + c]
        if (r < c) sx[r] = fma(-tile[r * VB_TILE + c], sx[c], sx[r]);
        __syncthreads();
    }
    x[k * VB_TILE + r] = sx[r];
}

__global__ void __launch_bounds__(256)
vb_bwd_update_kernel(const double *__restrict__ Lmat, int T, int k, double *__restrict__ x)
{
    __shared__ double sxk[VB_TILE];
    const int bj = blockIdx.x;  // 0..k-1
    const double *tile = Lmat + vb_tile_index(k, bj, T) * VB_TILE_ELEMS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < VB_TILE) sxk[threadIdx.x] = x[k * VB_TILE + threadIdx.x];
    __syncthreads();
    for (int c = warp; c < VB_TILE; c += 8) {
        const double *col = tile + c * VB_TILE;
        double v = col[lane] * sxk[lane];
        v = fma(col[lane + 32], sxk[lane + 32], v);
        v = fma(col[lane + 64], sxk[lane + 64], v);
        v = fma(col[lane + 96], sxk[lane + 96], v);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) x[bj * VB_TILE + c] -= v;
    }
}

long long col_base(int c, int T)
{
    return ((long long)c * T - ((long long)c * (c - 1)) / 2 - c) * VB_TILE_ELEMS;
}

}  // namespace

cudaError_t vb_cholesky_tiled(double *L, int T, int panel_tiles, int *d_info, cudaStream_t stream)
{
    cudaError_t err;
    const int potrf_smem = VB_TILE * POTRF_LD * (int)sizeof(double);
    const int trsm_smem = VB_TILE_ELEMS * (int)sizeof(double);
    if ((err = cudaFuncSetAttribute(vb_potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potrf_smem)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(vb_trsm_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem)) != cudaSuccess) return err;
    if (panel_tiles < 1) panel_tiles = 1;
    if (panel_tiles > VB_MAX_KCHUNKS) panel_tiles = VB_MAX_KCHUNKS;
    if ((err = cudaMemsetAsync(d_info, 0, sizeof(int), stream)) != cudaSuccess) return err;

    for (int k0 = 0; k0 < T; k0 += panel_tiles) {
        const int pw = (T - k0 < panel_tiles) ? (T - k0) : panel_tiles;
        for (int kk = 0; kk < pw; ++kk) {
            const int k = k0 + kk;
            vb_potrf_tile_kernel<<<1, POTRF_THREADS, potrf_smem, stream>>>(
                L + vb_tile_index(k, k, T) * VB_TILE_ELEMS, k, d_info);
            ++vb_launch_counter;
            if (k + 1 < T) {
                vb_trsm_tiles_kernel<<<T - k - 1, TRSM_THREADS, trsm_smem, stream>>>(L, T, k);
                ++vb_launch_counter;
            }
            if (kk + 1 < pw) {
                VbGemmParams p = {};
                p.a_base = p.b_base = L;
                p.a_chunk_off[0] = p.b_chunk_off[0] = col_base(k, T);
                p.nk = 1;
                p.c_base = L;
                p.T = T;
                p.j_lo = k + 1;
                p.j_hi = k0 + pw;
                p.alpha = -1.0;
                p.accumulate = 1;
                if ((err = vb_gemm_dispatch(p, stream)) != cudaSuccess) return err;
            }
        }
        if (k0 + pw < T) {
            VbGemmParams p = {};
            p.a_base = p.b_base = L;
            for (int kk = 0; kk < pw; ++kk) p.a_chunk_off[kk] = p.b_chunk_off[kk] = col_base(k0 + kk, T);
            p.nk = pw;
            p.c_base = L;
            p.T = T;
            p.j_lo = k0 + pw;
            p.j_hi = T;
            p.alpha = -1.0;
            p.accumulate = 1;
            if ((err = vb_gemm_dispatch(p, stream)) != cudaSuccess) return err;
        }
    }
    return cudaGetLastError();
}

// x (length T*128) is overwritten by the solution of L L^T x = b.
cudaError_t vb_solve_tiled(const double *L, int T, double *x, cudaStream_t stream)
{
    for (int k = 0; k < T; ++k) {
        vb_fwd_diag_kernel<<<1, VB_TILE, 0, stream>>>(L, T, k, x);
        if (k + 1 < T) vb_fwd_update_kernel<<<T - k - 1, VB_TILE, 0, stream>>>(L, T, k, x);
        vb_launch_counter += (k + 1 < T) ? 2 : 1;
    }
    for (int k = T - 1; k >= 0; --k) {
        vb_bwd_diag_kernel<<<1, VB_TILE, 0, stream>>>(L, T, k, x);
        if (k > 0) vb_bwd_update_kernel<<<k, 256, 0, stream>>>(L, T, k, x);
        vb_launch_counter += (k > 0) ? 2 : 1;
    }
    return cudaGetLastError();
}
