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

constexpr int POTRF_THREADS = 128;
constexpr int POTRF_BLK = 16;

// Factor one diagonal tile in place (lower triangle).  Thread r owns row r; the tile lives in shared
// memory (column-major, so a thread's accesses s[c*128 + r] are conflict-free and L(c0.., k) rows are
// contiguous for LDS.128 broadcasts).  Left-looking over 16-column blocks held in registers:
//   1. the block is updated with all columns to its left (16 independent FMAs per own-row load);
//   2. the 16 x 16 diagonal block is factored by the ONE warp whose lanes own those rows, exchanging
//      pivots and multipliers with shuffles -- no CTA barrier inside the 16 pivot steps;
//   3. every row below solves its 16 entries against that block (forward substitution, independent per row).
// Two CTA barriers per block instead of two per column (32 per block): the kernel is pure latency, and this
// took it from 65 us to the time of the dependent chains themselves (54 us); replacing sqrt + per-row division by
// one rsqrt per column and multiplications (round 2) removes the longest of those chains.
__global__ void __launch_bounds__(POTRF_THREADS)
vb_potrf_tile_kernel(double *__restrict__ tile, int k, int *__restrict__ d_info, long long stride)
{
    extern __shared__ __align__(16) double s[];  // s[c * 128 + r]
    __shared__ double s_rinv[VB_TILE];           // 1 / sqrt(pivot) of every column
    const int r = threadIdx.x;
    const int lane = r & 31;
    tile += (long long)blockIdx.x * stride;  // one CTA per matrix of a batch
    d_info += blockIdx.x;
    // 128 KB through one 128-thread CTA: keep 16 independent 16-byte loads in flight per thread, or the
    // staging (64 dependent round trips to L2) costs more than the factorisation
#pragma unroll 16
    for (int idx = r; idx < VB_TILE_ELEMS / 2; idx += POTRF_THREADS)
        reinterpret_cast<double2 *>(s)[idx] = __ldcg(reinterpret_cast<const double2 *>(tile) + idx);
    __syncthreads();
    for (int c0 = 0; c0 < VB_TILE; c0 += POTRF_BLK) {
        double a[POTRF_BLK];
#pragma unroll
        for (int c = 0; c < POTRF_BLK; ++c) a[c] = s[(c0 + c) * VB_TILE + r];
        if (r >= c0) {  // rows above the block only hold the (ignored) upper triangle
#pragma unroll 4
            for (int kp = 0; kp < c0; ++kp) {
                const double lrk = s[kp * VB_TILE + r];
                const double2 *lc = reinterpret_cast<const double2 *>(s + kp * VB_TILE + c0);
#pragma unroll
                for (int c2 = 0; c2 < POTRF_BLK / 2; ++c2) {
                    const double2 l = lc[c2];
                    a[2 * c2] = fma(-lrk, l.x, a[2 * c2]);
                    a[2 * c2 + 1] = fma(-lrk, l.y, a[2 * c2 + 1]);
                }
            }
        }
        if ((r >> 5) == (c0 >> 5)) {
            // the warp that owns rows c0 .. c0+15 (lanes lo .. lo+15) factors the diagonal block
            const int lo = c0 & 31;
#pragma unroll
            for (int kk = 0; kk < POTRF_BLK; ++kk) {
                const double piv = __shfl_sync(0xffffffffu, a[kk], lo + kk);
                if (lane == lo + kk && !(piv > 0.0)) atomicCAS(d_info, 0, k * VB_TILE + c0 + kk + 1);
                // 1/sqrt(pivot) once per column instead of a square root followed by a division per row: the
                // dependent sqrt + divide sequences (~350 cycles x 128 columns) were half of this kernel.  The
                // diagonal entry gets one Newton step (l = p r, l += r/2 (p - l^2)); rows multiply by r.
                const double rs = rsqrt(piv);
                double x = a[kk] * rs;
                if (lane == lo + kk) {
                    x = fma(0.5 * rs, fma(-x, x, piv), x);
                    s_rinv[c0 + kk] = rs;
                }
                a[kk] = x;
#pragma unroll
                for (int c = kk + 1; c < POTRF_BLK; ++c) {
                    const double lck = __shfl_sync(0xffffffffu, x, lo + c);  // L(c0+c, c0+kk)
                    a[c] = fma(-x, lck, a[c]);
                }
            }
            // lanes of this warp BELOW the block ran the same recurrence on their own rows: that is exactly
            // their forward substitution, so they are finished too (lanes above the block hold don't-cares)
            if (r >= c0) {
#pragma unroll
                for (int c = 0; c < POTRF_BLK; ++c) s[(c0 + c) * VB_TILE + r] = a[c];
            }
        }
        __syncthreads();
        if (r >= c0 + POTRF_BLK && (r >> 5) != (c0 >> 5)) {
#pragma unroll
            for (int kk = 0; kk < POTRF_BLK; ++kk) {
                const double *lcol = s + (c0 + kk) * VB_TILE + c0;  // column c0+kk of the factored block
                const double x = a[kk] * s_rinv[c0 + kk];
                a[kk] = x;
#pragma unroll
                for (int c = kk + 1; c < POTRF_BLK; ++c) a[c] = fma(-x, lcol[c], a[c]);
            }
#pragma unroll
            for (int c = 0; c < POTRF_BLK; ++c) s[(c0 + c) * VB_TILE + r] = a[c];
        }
        __syncthreads();
    }
#pragma unroll 8
    for (int idx = r; idx < VB_TILE_ELEMS / 2; idx += POTRF_THREADS) {
        const int c = idx >> 6, rr = (idx & 63) * 2;
        double2 v = reinterpret_cast<const double2 *>(s)[idx];
        if (rr < c) v.x = 0.0;
        if (rr + 1 < c) v.y = 0.0;
        reinterpret_cast<double2 *>(tile)[idx] = v;
    }
}

// X * L_kk^T = A for the tiles below the diagonal of block column k.
// One CTA per tile, thread r owns row r.  L_kk (column-major) sits in shared
// memory; row blocks of 32 unknowns live in registers:
//   a[c] -= x_prev * L[c][prev]   (prev columns: x re-read from the tile itself)
//   in-block forward substitution, fully unrolled.
// 128^3/2 FMAs per tile is 8 us of one SM's FP64 pipe, and L_kk takes 128 KB of shared memory, so only one
// CTA fits per SM: with one tile per CTA the ~390 tiles of an early block column ran as three waves of
// one-warp-per-scheduler CTAs.  A CTA now solves up to TRSM_MAX_TILES tiles against the same staged L_kk
// (128 threads per tile); the launcher picks ceil(tiles / SMs) so that a block column is ONE wave with the
// tiles of an SM running concurrently, and small problems keep one tile per CTA.
constexpr int TRSM_MAX_TILES = 4;
constexpr int TRSM_BLK = 32;

__global__ void __launch_bounds__(128 * TRSM_MAX_TILES)
vb_trsm_tiles_kernel(double *__restrict__ Lmat, int T, int k, long long stride)
{
    extern __shared__ __align__(16) double sL[];  // sL[c * 128 + r] = L_kk(r, c)
    __shared__ double s_inv[VB_TILE];
    Lmat += (long long)blockIdx.y * stride;  // grid.y = matrix of a batch
    const double *diag = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x & (VB_TILE - 1);
    const int tiles_per_cta = blockDim.x >> 7;
    const int below = blockIdx.x * tiles_per_cta + (threadIdx.x >> 7);  // which tile under the diagonal
#pragma unroll 8
    for (int idx = threadIdx.x; idx < VB_TILE_ELEMS / 2; idx += blockDim.x)
        reinterpret_cast<double2 *>(sL)[idx] = __ldcg(reinterpret_cast<const double2 *>(diag) + idx);
    __syncthreads();
    if (threadIdx.x < VB_TILE) s_inv[r] = 1.0 / sL[r * VB_TILE + r];
    __syncthreads();
    if (below >= T - k - 1) return;  // ragged last CTA (no barrier follows)
    double *tile = Lmat + vb_tile_index(k + 1 + below, k, T) * VB_TILE_ELEMS;

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
    extern __shared__ __align__(16) double st[];  // st[c * 128 + r] = L_kk(r, c)
    __shared__ double sy[VB_TILE];
    const double *tile = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    for (int idx = r; idx < VB_TILE_ELEMS / 2; idx += VB_TILE)
        reinterpret_cast<double2 *>(st)[idx] = reinterpret_cast<const double2 *>(tile)[idx];
    double v = x[k * VB_TILE + r];
    __syncthreads();
    for (int c = 0; c < VB_TILE; ++c) {
        if (r == c) sy[c] = v / st[c * VB_TILE + c];
        __syncthreads();
        if (r > c) v = fma(-st[c * VB_TILE + r], sy[c], v);
    }
    x[k * VB_TILE + r] = sy[r];
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
    extern __shared__ __align__(16) double st[];  // st[r * 129 + c] = L_kk(r, c)  (row-major, padded)
    __shared__ double sx[VB_TILE];
    const double *tile = Lmat + vb_tile_index(k, k, T) * VB_TILE_ELEMS;
    const int r = threadIdx.x;
    for (int c = 0; c < VB_TILE; ++c) st[r * (VB_TILE + 1) + c] = tile[c * VB_TILE + r];
    double v = x[k * VB_TILE + r];
    __syncthreads();
    for (int c = VB_TILE - 1; c >= 0; --c) {
        if (r == c) sx[c] = v / st[c * (VB_TILE + 1) + c];
        __syncthreads();
        // (L^T)(r, c) = L(c, r) for r < c
        if (r < c) v = fma(-st[c * (VB_TILE + 1) + r], sx[c], v);
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

// ---- single-launch triangular sweeps ------------------------------------------------------
// The per-block-row kernels above cost two launches and ~20 us of latency per block row
// (18 ms per sweep at T = 391).  vb_trisweep_kernel does a whole sweep in ONE cooperative
// launch: CTA c owns block rows c, c+G, ... (descending for the backward sweep); for a row it
// streams the row's off-diagonal tiles into REGISTERS before it needs the matching solution
// block (128 KB in flight per CTA -> HBM-bound, not latency-bound), waits on a per-block flag
// published by the CTA that solved that block, accumulates, then solves its diagonal tile from
// shared memory and publishes.  Co-residency is guaranteed by the cooperative launch, and a CTA
// only ever waits for lower-numbered (forward) / higher-numbered (backward) blocks, which are
// always ahead of it, so the flags cannot deadlock.
constexpr int SWEEP_THREADS = 256;

__device__ __forceinline__ void flag_wait(const int *flag)
{
    int v;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    } while (v == 0);
}
__device__ __forceinline__ void flag_set(int *flag)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}

// Diagonal-tile solves of the sweeps.  The old form (one pivot per step: divide, publish through shared memory, CTA
// barrier -- ~300 cycles x 128 steps = 19 us) WAS the critical path of a sweep: block row i cannot finish before
// block row i-1 has, so T x 19 us = 7.5 ms at T = 391 against 1.5 ms of HBM time for the 10 GB factor.  Now: the tile
// is solved in four 32-row blocks; warp b solves its 32 x 32 diagonal block by substitution with the pivots exchanged
// by shuffles (no barrier, reciprocal pivots computed once per tile: ~50 cycles per step), and one CTA barrier per
// block lets the rows below (above, for L^T) subtract the block's 32 solved unknowns.  Still plain substitution -- no
// explicit inverse, backward stable -- with x * (1/l) in place of x / l.
//   forward:  st[c * 128 + r] = L(r, c);  v = right-hand side of row `tid` (threads < 128), returns y_tid
__device__ __forceinline__ double sweep_diag_forward(const double *st, double *sv, double v, int tid)
{
    const int lane = tid & 31, wb = tid >> 5;
    const double rinv = tid < VB_TILE ? 1.0 / st[tid * VB_TILE + tid] : 0.0;
#pragma unroll 1
    for (int b = 0; b < VB_TILE / 32; ++b) {
        const int b0 = 32 * b;
        if (wb == b) {
            double lcol[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) lcol[c] = st[(b0 + c) * VB_TILE + tid];  // L(tid, b0 + c), used where lane > c
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const double yc = __shfl_sync(0xffffffffu, v * rinv, c);
                if (lane == c) v = yc;
                else if (lane > c) v = fma(-lcol[c], yc, v);
            }
            sv[tid] = v;
        }
        __syncthreads();
        if (tid < VB_TILE && wb > b) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                a0 = fma(st[(b0 + c + 0) * VB_TILE + tid], sv[b0 + c + 0], a0);
                a1 = fma(st[(b0 + c + 1) * VB_TILE + tid], sv[b0 + c + 1], a1);
                a2 = fma(st[(b0 + c + 2) * VB_TILE + tid], sv[b0 + c + 2], a2);
                a3 = fma(st[(b0 + c + 3) * VB_TILE + tid], sv[b0 + c + 3], a3);
            }
            v -= (a0 + a1) + (a2 + a3);
        }
    }
    return v;
}
//   backward: st[r * 129 + c] = L(r, c);  solves L^T x = v, blocks and pivots in descending order
__device__ __forceinline__ double sweep_diag_backward(const double *st, double *sv, double v, int tid)
{
    constexpr int LD = VB_TILE + 1;
    const int lane = tid & 31, wb = tid >> 5;
    const double rinv = tid < VB_TILE ? 1.0 / st[tid * LD + tid] : 0.0;
#pragma unroll 1
    for (int b = VB_TILE / 32 - 1; b >= 0; --b) {
        const int b0 = 32 * b;
        if (wb == b) {
            double lrow[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) lrow[c] = st[(b0 + c) * LD + tid];  // L(b0 + c, tid) = L^T(tid, b0 + c), used where lane < c
#pragma unroll
            for (int c = 31; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, v * rinv, c);
                if (lane == c) v = xc;
                else if (lane < c) v = fma(-lrow[c], xc, v);
            }
            sv[tid] = v;
        }
        __syncthreads();
        if (wb < b) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                a0 = fma(st[(b0 + c + 0) * LD + tid], sv[b0 + c + 0], a0);
                a1 = fma(st[(b0 + c + 1) * LD + tid], sv[b0 + c + 1], a1);
                a2 = fma(st[(b0 + c + 2) * LD + tid], sv[b0 + c + 2], a2);
                a3 = fma(st[(b0 + c + 3) * LD + tid], sv[b0 + c + 3], a3);
            }
            v -= (a0 + a1) + (a2 + a3);
        }
    }
    return v;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
vb_trisweep_kernel(const double *__restrict__ Lmat, int T, double *x, int *flags, int batch, long long stride)
{
    extern __shared__ __align__(16) double sm[];
    // a batch of independent systems shares the launch: CTAs [b*G, (b+1)*G) sweep matrix b
    const int G = gridDim.x / batch;
    const int mat = blockIdx.x / G, cta = blockIdx.x - mat * G;
    Lmat += (long long)mat * stride;
    x += (long long)mat * T * VB_TILE;
    flags += mat * T;
    double *st = sm;                                    // diagonal tile: fwd [c*128 + r], bwd [r*129 + c]
    double *sv = sm + VB_TILE * (VB_TILE + 1);          // solution block of the tile being applied
    double *sacc = sv + VB_TILE;                        // 2 x 128 partial sums
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int it = cta; it < T; it += G) {
        const int i = BACKWARD ? T - 1 - it : it;
        // stage the diagonal tile (needed last) while the off-diagonal stream runs
        const double *dtile = Lmat + vb_tile_index(i, i, T) * VB_TILE_ELEMS;
        if (!BACKWARD) {
            for (int idx = tid; idx < VB_TILE_ELEMS / 2; idx += SWEEP_THREADS)
                reinterpret_cast<double2 *>(st)[idx] = reinterpret_cast<const double2 *>(dtile)[idx];
        } else {
            for (int idx = tid; idx < VB_TILE_ELEMS; idx += SWEEP_THREADS) {
                const int c = idx >> 7, r = idx & 127;
                st[r * (VB_TILE + 1) + c] = dtile[idx];
            }
        }
        // this row's right-hand side block: nobody else writes it, so it is fetched ahead of the critical path
        const double xi = tid < VB_TILE ? x[i * VB_TILE + tid] : 0.0;
        if (!BACKWARD) {
            // y_i = L_ii^{-1} (b_i - sum_{k<i} L(i,k) y_k): thread (r, h) covers columns c = 2j + h
            const int r = tid & 127, h = tid >> 7;
            double acc = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            for (int k = 0; k < i; ++k) {
                const double *tile = Lmat + vb_tile_index(i, k, T) * VB_TILE_ELEMS;
                double l[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) l[j] = __ldcs(tile + (2 * j + h) * VB_TILE + r);
                if (tid == 0) flag_wait(flags + k);
                __syncthreads();
                if (tid < VB_TILE) sv[tid] = __ldcg(x + k * VB_TILE + tid);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                    acc = fma(l[j], sv[2 * j + h], acc);
                    acc1 = fma(l[j + 1], sv[2 * j + 2 + h], acc1);
                    acc2 = fma(l[j + 2], sv[2 * j + 4 + h], acc2);
                    acc3 = fma(l[j + 3], sv[2 * j + 6 + h], acc3);
                }
            }
            sacc[tid] = (acc + acc1) + (acc2 + acc3);
            __syncthreads();
            double v = 0.0;
            if (tid < VB_TILE) v = xi - (sacc[tid] + sacc[tid + VB_TILE]);
            v = sweep_diag_forward(st, sv, v, tid);
            if (tid < VB_TILE) x[i * VB_TILE + tid] = v;
        } else {
            // x_i = L_ii^{-T} (y_i - sum_{k>i} L(k,i)^T x_k): warp w covers columns 16w..16w+15, lanes run along rows
            double acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = 0.0;
            for (int k = T - 1; k > i; --k) {
                const double *tile = Lmat + vb_tile_index(k, i, T) * VB_TILE_ELEMS;
                double l[64];
#pragma unroll
                for (int c = 0; c < 16; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) l[c * 4 + q] = __ldcs(tile + (warp * 16 + c) * VB_TILE + lane + 32 * q);
                if (tid == 0) flag_wait(flags + k);
                __syncthreads();
                if (tid < VB_TILE) sv[tid] = __ldcg(x + k * VB_TILE + tid);
                __syncthreads();
#pragma unroll
                for (int c = 0; c < 16; ++c)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[c] = fma(l[c * 4 + q], sv[lane + 32 * q], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double v = acc[c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) sacc[warp * 16 + c] = v;
            }
            __syncthreads();
            double v = 0.0;
            if (tid < VB_TILE) v = xi - sacc[tid];
            v = sweep_diag_backward(st, sv, v, tid);
            if (tid < VB_TILE) x[i * VB_TILE + tid] = v;
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            flag_set(flags + i);
        }
        __syncthreads();
    }
}

long long col_base(int c, int T)
{
    return ((long long)c * T - ((long long)c * (c - 1)) / 2 - c) * VB_TILE_ELEMS;
}

}  // namespace

static cudaError_t chol_set_attributes(int *potrf_smem, int *trsm_smem)
{
    cudaError_t err;
    *potrf_smem = VB_TILE_ELEMS * (int)sizeof(double);
    *trsm_smem = VB_TILE_ELEMS * (int)sizeof(double);
    if ((err = cudaFuncSetAttribute(vb_potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, *potrf_smem)) != cudaSuccess) return err;
    return cudaFuncSetAttribute(vb_trsm_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, *trsm_smem);
}

// Factor the outer panel of block columns [k0, k0 + pw): potrf + trsm per column and the rank-128
// updates of the panel's own remaining columns.  Everything left of k0 must already be applied.
static cudaError_t factor_panel_ex(double *L, int T, int k0, int pw, int *d_info, cudaStream_t stream, int batch,
                                   long long stride, int gemm_grid_limit);

cudaError_t vb_chol_factor_panel(double *L, int T, int k0, int pw, int *d_info, cudaStream_t stream, int batch,
                                 long long stride)
{
    return factor_panel_ex(L, T, k0, pw, d_info, stream, batch, stride, 0);
}

// gemm_grid_limit > 0: the panel's own rank-128 updates may only use that many SMs (look-ahead: the rest of the
// machine runs the previous panel's trailing update, and a persistent grid wider than the free SMs could not finish
// before that update does)
static cudaError_t factor_panel_ex(double *L, int T, int k0, int pw, int *d_info, cudaStream_t stream, int batch,
                                   long long stride, int gemm_grid_limit)
{
    cudaError_t err;
    int potrf_smem, trsm_smem;
    if ((err = chol_set_attributes(&potrf_smem, &trsm_smem)) != cudaSuccess) return err;
    if (batch < 1) batch = 1;
    int sms = 0;
    if ((err = vb_sm_count(&sms)) != cudaSuccess) return err;
    for (int kk = 0; kk < pw; ++kk) {
        const int k = k0 + kk;
        vb_potrf_tile_kernel<<<batch, POTRF_THREADS, potrf_smem, stream>>>(
            L + vb_tile_index(k, k, T) * VB_TILE_ELEMS, k, d_info, stride);
        ++vb_launch_counter;
        if (k + 1 < T) {
            const int ntiles = T - k - 1;
            int tpc = (int)(((long long)ntiles * batch + sms - 1) / sms);
            // one wave when it fits in <= 3 tiles per SM (37 us at T = 391); beyond that several waves are
            // inevitable and 2 tiles per CTA measured best (batched sweep: 11.6 ms per job vs 12.9 / 12.3 / 13.2)
            tpc = tpc < 1 ? 1 : (tpc > 3 ? 2 : tpc);
            if (vb_trsm_tiles >= 1 && vb_trsm_tiles <= TRSM_MAX_TILES) tpc = vb_trsm_tiles;
            vb_trsm_tiles_kernel<<<dim3((ntiles + tpc - 1) / tpc, batch), 128 * tpc, trsm_smem, stream>>>(L, T, k, stride);
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
            p.batch = batch;
            p.batch_stride = stride;
            p.grid_limit = gemm_grid_limit;
            // beside a running trailing update these launches own only a few SMs and overlap the update's own
            // CUDA-event span: they stay out of the roofline profile (time and flops) instead of skewing it
            p.untimed = gemm_grid_limit > 0;
            if ((err = vb_gemm_dispatch(p, stream)) != cudaSuccess) return err;
        }
    }
    return cudaGetLastError();
}

// Trailing update C(:, [j_lo, j_hi)) -= L(:, panel) L([j_lo, j_hi), panel)^T with the factored panel
// [k0, k0 + pw): the bulk of the m^3/3 flops, one launch of the DMMA tile kernel (K = 128 * pw).
static cudaError_t trailing_update_ex(double *L, int T, int k0, int pw, int j_lo, int j_hi, cudaStream_t stream,
                                      int batch, long long stride, unsigned long long *queue, int grid_limit, int untimed)
{
    if (j_lo < k0 + pw) j_lo = k0 + pw;
    if (j_hi > T) j_hi = T;
    if (j_lo >= j_hi) return cudaSuccess;
    VbGemmParams p = {};
    p.a_base = p.b_base = L;
    for (int kk = 0; kk < pw; ++kk) p.a_chunk_off[kk] = p.b_chunk_off[kk] = col_base(k0 + kk, T);
    p.nk = pw;
    p.c_base = L;
    p.T = T;
    p.j_lo = j_lo;
    p.j_hi = j_hi;
    p.alpha = -1.0;
    p.accumulate = 1;
    p.batch = batch < 1 ? 1 : batch;
    p.batch_stride = stride;
    p.queue = queue;
    p.grid_limit = grid_limit;
    p.untimed = untimed;
    return vb_gemm_dispatch(p, stream);
}

cudaError_t vb_chol_trailing_update(double *L, int T, int k0, int pw, int j_lo, int j_hi, cudaStream_t stream,
                                    int batch, long long stride)
{
    return trailing_update_ex(L, T, k0, pw, j_lo, j_hi, stream, batch, stride, nullptr, 0, 0);
}

cudaError_t vb_chol_trailing_update_ranges(double *L, int T, int k0, int pw, int nranges, const int *j_lo,
                                           const int *j_hi, cudaStream_t stream)
{
    for (int r0 = 0; r0 < nranges; r0 += VB_MAX_RANGES) {
        VbGemmParams p = {};
        p.a_base = p.b_base = L;
        p.c_base = L;
        p.T = T;
        p.nk = pw;
        for (int kk = 0; kk < pw; ++kk) p.a_chunk_off[kk] = p.b_chunk_off[kk] = col_base(k0 + kk, T);
        p.alpha = -1.0;
        p.accumulate = 1;
        p.batch = 1;
        for (int r = r0; r < nranges && r < r0 + VB_MAX_RANGES; ++r) {
            const int lo = j_lo[r] < k0 + pw ? k0 + pw : j_lo[r];
            const int hi = j_hi[r] > T ? T : j_hi[r];
            if (lo >= hi) continue;
            p.range_lo[p.nranges] = lo;
            p.range_hi[p.nranges] = hi;
            ++p.nranges;
        }
        if (p.nranges == 0) continue;
        p.j_lo = p.range_lo[0];
        p.j_hi = p.range_hi[p.nranges - 1];
        cudaError_t e = vb_gemm_dispatch(p, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Multi-GPU panel exchange over peer memory: the owner of a panel PUSHES it into every peer's copy of the matrix
// with copy engines (no SM, no collective) and then writes the fit's epoch into that peer's per-panel flag word.
// The receiver's stream parks on the flag with this one-thread kernel (system-scope acquire: the writer is another
// GPU).  Bounded: ~10 s without the flag records -1 in *status instead of hanging the box.
__global__ void vb_wait_flag_kernel(const long long *flag, long long value, int *status)
{
    long long v = 0;
    for (unsigned spins = 0; spins < 50000000u; ++spins) {
        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= value) return;
        __nanosleep(200);
    }
    atomicCAS(status, 0, -1);
}

cudaError_t vb_chol_wait_flag(const long long *flag, long long value, int *d_info, cudaStream_t stream)
{
    vb_wait_flag_kernel<<<1, 1, 0, stream>>>(flag, value, d_info);
    ++vb_launch_counter;
    return cudaGetLastError();
}

long long vb_chol_column_offset(int c, int T) { return col_base(c, T) + (long long)c * VB_TILE_ELEMS; }

// Look-ahead (single GPU).  Without it every outer panel costs ~1 ms in which 147 SMs wait for one CTA's potrf chain.
// With `lookahead_sms` = R > 0 (needs the dynamic tile queue, gemm variant 4) panel p+1 is factored BESIDE the
// trailing update of panel p:
//   main stream:  U1 = update of panel p+1's own block columns (all SMs, short)
//                 factor panel p+1                                  -> runs on the R SMs the big update leaves free
//                 U2b = R more CTAs on the big update's tile queue  -> the freed SMs join as soon as the panel is done
//   side stream:  U2a = update of everything right of panel p+1, grid capped at SMs - R, same tile queue as U2b
// A 128 KB panel CTA and a 203 KB GEMM CTA cannot share an SM, so the split is exact.  Every tile is still updated
// by exactly one CTA with the same operands in the same order: the factor is bit-identical to the plain schedule.
int vb_chol_lookahead_min_rows = 48;  // look-ahead only while at least this many block rows remain (below, the
                                      // panel chain is longer than the update that should hide it)

static cudaError_t chol_side_stream(cudaStream_t *side, cudaEvent_t *e1, cudaEvent_t *e2)
{
    static cudaStream_t streams[256] = {nullptr};
    static cudaEvent_t ev1[256] = {nullptr}, ev2[256] = {nullptr};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 256) return cudaErrorInvalidDevice;
    if (!streams[dev]) {
        if ((e = cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev1[dev], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev2[dev], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    *side = streams[dev];
    *e1 = ev1[dev];
    *e2 = ev2[dev];
    return cudaSuccess;
}

cudaError_t vb_cholesky_tiled(double *L, int T, int panel_tiles, int *d_info, cudaStream_t stream, int batch,
                              long long stride, int lookahead_sms)
{
    cudaError_t err;
    if (batch < 1) batch = 1;
    if (panel_tiles < 1) panel_tiles = 1;
    if (panel_tiles > VB_MAX_KCHUNKS) panel_tiles = VB_MAX_KCHUNKS;
    if ((err = cudaMemsetAsync(d_info, 0, sizeof(int) * batch, stream)) != cudaSuccess) return err;
    int sms = 0;
    if ((err = vb_sm_count(&sms)) != cudaSuccess) return err;
    const bool can_look_ahead = lookahead_sms > 0 && lookahead_sms < sms && vb_gemm_variant == 4;
    cudaStream_t side = nullptr;
    cudaEvent_t e1 = nullptr, e2 = nullptr;
    if (can_look_ahead && (err = chol_side_stream(&side, &e1, &e2)) != cudaSuccess) return err;
    bool factored = false;  // is the current panel already factored (by the previous iteration's look-ahead)?
    for (int k0 = 0; k0 < T; k0 += panel_tiles) {
        const int pw = (T - k0 < panel_tiles) ? (T - k0) : panel_tiles;
        if (!factored && (err = vb_chol_factor_panel(L, T, k0, pw, d_info, stream, batch, stride)) != cudaSuccess) return err;
        factored = false;
        const int k1 = k0 + pw;
        if (k1 >= T) break;
        const int pw1 = (T - k1 < panel_tiles) ? (T - k1) : panel_tiles;
        if (can_look_ahead && T - k1 >= vb_chol_lookahead_min_rows && k1 + pw1 < T) {
            // U1: the next panel's own columns, then its factorisation beside the rest of the update
            if ((err = trailing_update_ex(L, T, k0, pw, k1, k1 + pw1, stream, batch, stride, nullptr, 0, 0)) != cudaSuccess) return err;
            unsigned long long *queue = nullptr;
            if ((err = vb_gemm_new_queue(&queue, stream)) != cudaSuccess) return err;  // zeroed before e1 in stream order
            if ((err = cudaEventRecord(e1, stream)) != cudaSuccess) return err;
            if ((err = cudaStreamWaitEvent(side, e1, 0)) != cudaSuccess) return err;
            if ((err = trailing_update_ex(L, T, k0, pw, k1 + pw1, T, side, batch, stride, queue, sms - lookahead_sms, 2)) != cudaSuccess) return err;
            if ((err = cudaEventRecord(e2, side)) != cudaSuccess) return err;
            if ((err = factor_panel_ex(L, T, k1, pw1, d_info, stream, batch, stride, lookahead_sms)) != cudaSuccess) return err;
            if ((err = trailing_update_ex(L, T, k0, pw, k1 + pw1, T, stream, batch, stride, queue, lookahead_sms, 1)) != cudaSuccess) return err;
            if ((err = cudaStreamWaitEvent(stream, e2, 0)) != cudaSuccess) return err;
            factored = true;
        } else {
            if ((err = vb_chol_trailing_update(L, T, k0, pw, k1, T, stream, batch, stride)) != cudaSuccess) return err;
        }
    }
    return cudaGetLastError();
}

// x (length T*128) is overwritten by the solution of L L^T x = b.
int vb_trsm_tiles = 0;    // tiles per trsm CTA: 0 = ceil(tiles / SMs) capped at 4
int vb_solve_variant = 1;  // 1: one cooperative launch per sweep, 0: two launches per block row

static cudaError_t solve_cooperative(const double *L, int T, double *x, int *flags, cudaStream_t stream, int batch,
                                     long long stride)
{
    const int smem = (VB_TILE * (VB_TILE + 1) + VB_TILE + 2 * SWEEP_THREADS) * (int)sizeof(double);
    cudaError_t err;
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return err;
    if (!coop) return cudaErrorNotSupported;
    if ((err = cudaFuncSetAttribute(vb_trisweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(vb_trisweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vb_trisweep_kernel<false>, SWEEP_THREADS, smem)) != cudaSuccess) return err;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    // every CTA of the launch must be co-resident (the flags spin): one CTA per SM, shared by the batch
    int per_mat = sms * per_sm / batch;
    if (per_mat < 1) return cudaErrorLaunchOutOfResources;
    if (per_mat > T) per_mat = T;
    const int grid = per_mat * batch;
    if ((err = cudaMemsetAsync(flags, 0, sizeof(int) * 2 * T * batch, stream)) != cudaSuccess) return err;
    int *fl_f = flags, *fl_b = flags + (long long)T * batch;
    void *args_f[] = {(void *)&L, (void *)&T, (void *)&x, (void *)&fl_f, (void *)&batch, (void *)&stride};
    void *args_b[] = {(void *)&L, (void *)&T, (void *)&x, (void *)&fl_b, (void *)&batch, (void *)&stride};
    if ((err = cudaLaunchCooperativeKernel((void *)vb_trisweep_kernel<false>, dim3(grid), dim3(SWEEP_THREADS), args_f, smem, stream)) != cudaSuccess) return err;
    if ((err = cudaLaunchCooperativeKernel((void *)vb_trisweep_kernel<true>, dim3(grid), dim3(SWEEP_THREADS), args_b, smem, stream)) != cudaSuccess) return err;
    vb_launch_counter += 2;
    return cudaGetLastError();
}

cudaError_t vb_solve_tiled(const double *L, int T, double *x, int *flags, cudaStream_t stream, int batch,
                           long long stride)
{
    if (batch < 1) batch = 1;
    if (vb_solve_variant == 1 && flags != nullptr) return solve_cooperative(L, T, x, flags, stream, batch, stride);
    if (batch > 1) {
        for (int b = 0; b < batch; ++b) {
            cudaError_t eb = vb_solve_tiled(L + (long long)b * stride, T, x + (long long)b * T * VB_TILE, nullptr, stream, 1, 0);
            if (eb != cudaSuccess) return eb;
        }
        return cudaSuccess;
    }
    const int fwd_smem = VB_TILE_ELEMS * (int)sizeof(double);
    const int bwd_smem = VB_TILE * (VB_TILE + 1) * (int)sizeof(double);
    cudaError_t err;
    if ((err = cudaFuncSetAttribute(vb_fwd_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(vb_bwd_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem)) != cudaSuccess) return err;
    for (int k = 0; k < T; ++k) {
        vb_fwd_diag_kernel<<<1, VB_TILE, fwd_smem, stream>>>(L, T, k, x);
        if (k + 1 < T) vb_fwd_update_kernel<<<T - k - 1, VB_TILE, 0, stream>>>(L, T, k, x);
        vb_launch_counter += (k + 1 < T) ? 2 : 1;
    }
    for (int k = T - 1; k >= 0; --k) {
        vb_bwd_diag_kernel<<<1, VB_TILE, bwd_smem, stream>>>(L, T, k, x);
        if (k > 0) vb_bwd_update_kernel<<<k, 256, 0, stream>>>(L, T, k, x);
        vb_launch_counter += (k > 0) ? 2 : 1;
    }
    return cudaGetLastError();
}
