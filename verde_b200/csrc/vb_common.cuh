// vb_common.cuh -- shared device code for the verde_b200 CUDA library (sm_100a).
//
// Layout conventions used by every kernel in this directory
// ---------------------------------------------------------
// * All arithmetic is IEEE float64 (the reference's dtype, verde/spline.py:529).
// * The normal matrix G (m x m, symmetric) is stored as its LOWER triangle in
//   "tile-packed" form: T = ceil(m/128) block rows; tile (bi, bj), bi >= bj,
//   is a contiguous 128x128 COLUMN-major block (element (r, c) at c*128 + r)
//   at tile index  bj*T - bj*(bj-1)/2 + (bi - bj)   (block-column-major).
//   A column-major tile is, read the other way, a "k-major operand chunk"
//   chunk[k][i] (k = column, i = row) -- exactly what the DMMA mainloop
//   consumes, so Cholesky panels feed the tile GEMM without any transposition.
// * Jacobian panels are stored as k-major operand chunks too:
//   chunk[k][i], k = data row inside a 128-row chunk, i = force index inside a
//   128-wide force block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VB_TILE 128
#define VB_TILE_ELEMS (VB_TILE * VB_TILE)

__host__ __device__ __forceinline__ long long vb_tile_index(int bi, int bj, int T)
{
    return (long long)bj * T - ((long long)bj * (bj - 1)) / 2 + (bi - bj);
}

__host__ __device__ __forceinline__ long long vb_num_tiles(int T)
{
    return ((long long)T * (T + 1)) / 2;
}

// ---------------------------------------------------------------------------
// Green's functions, "exact" flavour: same operation order as the reference
// (no FMA contraction in r^2, IEEE sqrt, <=1 ulp log).  Used wherever the
// entry count is O(n*m) once (Jacobian materialisation, Gram panel generator).
// ---------------------------------------------------------------------------

// verde/spline.py:587-605.  The reference's d<1 branch d*(log(d**d) - d) is the
// same function as d^2 (log d - 1) evaluated with ~eps/(d |log d|) relative
// noise from the pow/log round trip; the closed form is used for both branches
// (difference <= ~4e-16 absolute on entries of magnitude < 1).
// d == 0 gives exactly 0 as in the reference (0**0 = 1, log 1 = 0).
__device__ __forceinline__ double vb_greens_exact(double dx, double dy, double mindist)
{
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double d = __dadd_rn(__dsqrt_rn(r2), mindist);
    if (d == 0.0) return 0.0;
    const double lg = log(d);
    return __dmul_rn(__dmul_rn(d, d), __dadd_rn(lg, -1.0));
}

// verde/vector.py:393-405 (compiled with fastmath in the reference, so only the
// formulas are binding, not the association order).
__device__ __forceinline__ void vb_greens2d_exact(double dx, double dy, double mindist,
                                                  double poisson, double &g_ee, double &g_nn,
                                                  double &g_ne)
{
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const double d = __dadd_rn(__dsqrt_rn(r2), mindist);
    const double ln_r = (3.0 - poisson) * log(d);
    const double over_r2 = (1.0 + poisson) / (d * d);
    g_ee = ln_r + over_r2 * (dy * dy);
    g_nn = ln_r + over_r2 * (dx * dx);
    g_ne = -over_r2 * dx * dy;
}

// ---------------------------------------------------------------------------
// Green's functions, "fast" flavour for the O(N*M) matrix-free predict, where
// the FP64 pipe is the roofline (36.8 TFLOP/s DFMA measured; libdevice log costs
// ~50 FP64 issue slots, sqrt ~16).  Table-driven log (no division) + Goldschmidt
// sqrt seeded by MUFU.RSQ64H: ~11 + ~7 FP64 slots.  Accuracy: |log error| <=
// ~2e-16 * max(1, |log d|) absolute, which bounds the Green's value error by a
// few eps * d^2 * max(1, |log d|)  (tests/test_gpu_greens.py pins this against
// the exact flavour and the oracle).
// ---------------------------------------------------------------------------
#define VB_LOGTAB_ENTRIES 128
// smem layout: entry idx is replicated 8x at byte offset idx*128 + (lane%8)*16
// so that the 8 lanes of a quarter-warp always hit 8 different 16-byte slots:
// one LDS.128 per evaluation, never a bank conflict.
#define VB_LOGTAB_SMEM_DOUBLES (VB_LOGTAB_ENTRIES * 16)

// gtab: global copy of the base table {rcp_c, log_c} (filled by vb_init_tables()).
__device__ __forceinline__ void vb_logtab_stage(double *smem_tab, const double2 *gtab)
{
    // called by all threads of a CTA; caller syncs afterwards
    for (int t = threadIdx.x; t < VB_LOGTAB_ENTRIES * 8; t += blockDim.x) {
        const double2 v = gtab[t >> 3];
        reinterpret_cast<double2 *>(smem_tab)[t] = v;
    }
}

// log(d) for normal positive d.  `tab` points at this lane's replica column:
// reinterpret_cast<const double2*>(smem_tab) + (lane & 7).
__device__ __forceinline__ double vb_log_fast(double d, const double2 *tab)
{
    const int hi = __double2hiint(d);
    const int lo = __double2loint(d);
    const int idx = (hi >> 13) & 0x7F;
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    // double(e) without an I2F: 2^52 + 2^31 + (e ^ 0x80000000) trick
    const int e = (hi >> 20) - 1023;
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    const double2 t = tab[idx * 8];
    const double r = fma(m, t.x, -1.0);
    // log1p(r), |r| <= 2^-8: r - r^2/2 + ... + r^7/7
    double q = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    q = fma(r, q, 1.0 / 5.0);
    q = fma(r, q, -1.0 / 4.0);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double r2 = r * r;
    const double l1p = fma(r2, q, r);
    return fma(ed, 0.6931471805599453094, t.y) + l1p;
}

// sqrt(x) for x >= 0 via rsqrt.approx (MUFU.RSQ64H, 2^-22) + 2 Goldschmidt steps.
__device__ __forceinline__ double vb_sqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y;
    double h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    // x == 0 -> y = inf -> NaN chain; x tiny/denormal is flushed by .ftz likewise
    return (__double2hiint(x) < 0x00200000) ? 0.0 : g;
}

// d^2 (log d - 1) with d = sqrt(dx^2+dy^2) + mindist.  HAS_MINDIST=false folds the
// sqrt away: r2 * (0.5*log(r2) - 1).
template <bool HAS_MINDIST>
__device__ __forceinline__ double vb_greens_fast(double dx, double dy, double mindist,
                                                 const double2 *tab)
{
    const double r2 = fma(dx, dx, dy * dy);
    double d, dd, half;
    if (HAS_MINDIST) {
        d = vb_sqrt_fast(r2) + mindist;
        dd = d * d;
        half = 1.0;
    } else {
        d = r2;
        dd = r2;
        half = 0.5;
    }
    const int hi = __double2hiint(d);
    const double lg = vb_log_fast(d, tab);
    double g = dd * fma(half, lg, -1.0);
    // d < 2^-511 (incl. 0, where the reference gives exactly 0): d^2 underflows anyway
    if (hi < 0x20000000) g = 0.0;
    // inf / NaN propagate like the reference's arithmetic would
    if (hi >= 0x7FF00000) g = d;
    return g;
}

template <bool HAS_MINDIST>
__device__ __forceinline__ void vb_greens2d_fast(double dx, double dy, double mindist,
                                                 double three_minus_nu, double one_plus_nu,
                                                 const double2 *tab, double &g_ee, double &g_nn,
                                                 double &g_ne)
{
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double r2 = dx2 + dy2;
    double lg, inv_d2;
    if (HAS_MINDIST) {
        const double d = vb_sqrt_fast(r2) + mindist;
        lg = vb_log_fast(d, tab);
        inv_d2 = 1.0 / (d * d);
        if (__double2hiint(d) >= 0x7FF00000 || __double2hiint(d) < 0x00100000) lg = log(d);
    } else {
        lg = 0.5 * vb_log_fast(r2, tab);
        inv_d2 = 1.0 / r2;
        if (__double2hiint(r2) >= 0x7FF00000 || __double2hiint(r2) < 0x00100000) lg = 0.5 * log(r2);
    }
    const double ln_r = three_minus_nu * lg;
    const double q = one_plus_nu * inv_d2;
    g_ee = fma(q, dy2, ln_r);
    g_nn = fma(q, dx2, ln_r);
    g_ne = -q * dx * dy;
}
