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
// ~50 FP64 issue slots, sqrt ~16).  Table-driven log (no division, 8 slots) + a
// third-order sqrt correction of the MUFU.RSQ64H seed (5 slots).  Accuracy: |log error| <=
// ~2e-16 * max(1, |log d|) absolute, which bounds the Green's value error by a
// few eps * d^2 * max(1, |log d|)  (tests/test_gpu_greens.py pins this against
// the exact flavour and the oracle).
// ---------------------------------------------------------------------------
#define VB_LOGTAB_ENTRIES 1024
// smem layout: the plain table, 16 bytes per entry {1/c_k, log c_k}, c_k = 1 + (k + 0.5)/1024.
// One LDS.128 per evaluation at a data-dependent index: ~2.5-way bank conflicts on average,
// which the (otherwise idle) LSU pipe absorbs; in exchange |r| <= 2^-11 and log1p needs only
// a degree-4 polynomial (truncation r^5/5 <= 2^-57 absolute).
#define VB_LOGTAB_SMEM_DOUBLES (VB_LOGTAB_ENTRIES * 2)

// gtab: global copy of the table (filled by vb_init_tables()).  REP > 1 interleaves REP copies
// (entry idx, copy r at double2 index idx*REP + r); lane l reads copy l % REP, which spreads the
// 8 lanes of a quarter-warp over more bank groups (REP = 4: ~1.5-way instead of ~2.5-way conflicts).
template <int REP = 1>
__device__ __forceinline__ void vb_logtab_stage(double *smem_tab, const double2 *gtab)
{
    // called by all threads of a CTA; caller syncs afterwards
    for (int t = threadIdx.x; t < VB_LOGTAB_ENTRIES * REP; t += blockDim.x)
        reinterpret_cast<double2 *>(smem_tab)[t] = gtab[t / REP];
}

// log(d) for normal positive d; `tab` is the shared-memory table.  Never produces inf/NaN:
// d = 0 reads as 2^-1023 (finite log), so r2 * log(r2) is an exact 0 at r2 = 0.
// 8 FP64 issue slots: 1 (exponent) + 1 (r) + 2 (Horner) + 1 (r^2) + 1 + 2 (assembly).
template <int REP = 1>
__device__ __forceinline__ double vb_log_fast(double d, const double2 *tab)
{
    static_assert(REP == 1 || REP == 2 || REP == 4 || REP == 8, "replication factor");
    const int hi = __double2hiint(d);
    const int lo = __double2loint(d);
    // byte offset of table entry idx = mantissa bits 19..10: idx * 16 * REP (tab already points at
    // this lane's copy)
    const int off = REP == 1 ? ((hi >> 6) & 0x3FF0) : REP == 2 ? ((hi >> 5) & 0x7FE0) : REP == 4 ? ((hi >> 4) & 0xFFC0)
                                                                                                  : ((hi >> 3) & 0x1FF80);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    // double(biased exponent - 1023) without an I2F: 2^52 + biased as a bit pattern
    const double ed = __hiloint2double(0x43300000, (unsigned)hi >> 20) - 4503599627371519.0;
    const double2 t = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(tab) + off);
    const double r = fma(m, t.x, -1.0);
    // log1p(r), |r| <= 2^-11: r - r^2/2 + r^3/3 - r^4/4
    double q = fma(r, -0.25, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double r2 = r * r;
    const double l1p = fma(r2, q, r);
    return fma(ed, 0.6931471805599453094, t.y) + l1p;
}

// sqrt(x) for x > 0 (normal): MUFU.RSQ64H seed (rel. error d ~ 2^-22) and ONE third-order
// correction: with t = x*y, e = 1 - t*y,  sqrt(x) = t (1-e)^(-1/2) = t (1 + e/2 + 3e^2/8) + O(e^3),
// i.e. relative error ~2.5 d^3 < 2^-60.  5 FP64 issue slots; callers keep x away from 0.
__device__ __forceinline__ double vb_sqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = e * p;
    return fma(t, q, t);
}

// 1/x for normal x: MUFU.RCP64H seed + two Newton steps (4 FP64 issue slots)
__device__ __forceinline__ double vb_rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// d^2 (log d - 1) with d = sqrt(dx^2+dy^2) + mindist, for FINITE inputs (the predict
// kernels screen non-finite coordinates / forces outside their inner loop).
//   HAS_MINDIST = false folds the sqrt away: r2 * (0.5*log(r2) - 1), exactly 0 at r2 = 0
//   (15 FP64 slots per pair); HAS_MINDIST = true adds 1e-300 to r2 inside the FMA chain so the
//   rsqrt seed never sees 0 (sqrt -> 1e-150, absorbed by any mindist > 1e-134): 22 slots.
#define VB_R2_FLOOR 1e-300
template <bool HAS_MINDIST, int REP = 1>
__device__ __forceinline__ double vb_greens_fast(double dx, double dy, double mindist,
                                                 const double2 *tab)
{
    if (HAS_MINDIST) {
        const double r2 = fma(dx, dx, fma(dy, dy, VB_R2_FLOOR));
        const double d = vb_sqrt_fast(r2) + mindist;
        const double lg = vb_log_fast<REP>(d, tab);
        return (d * d) * (lg - 1.0);
    } else {
        const double r2 = fma(dx, dx, dy * dy);
        const double lg = vb_log_fast<REP>(r2, tab);
        return r2 * fma(0.5, lg, -1.0);
    }
}

// elastic Green's functions (verde/vector.py:393-405) for finite inputs with d > 0
template <bool HAS_MINDIST>
__device__ __forceinline__ void vb_greens2d_fast(double dx, double dy, double mindist,
                                                 double three_minus_nu, double one_plus_nu,
                                                 const double2 *tab, double &g_ee, double &g_nn,
                                                 double &g_ne)
{
    const double dx2 = dx * dx, dy2 = dy * dy;
    double lg, inv_d2;
    if (HAS_MINDIST) {
        const double r2 = dx2 + (dy2 + VB_R2_FLOOR);
        const double d = vb_sqrt_fast(r2) + mindist;
        lg = vb_log_fast(d, tab);
        inv_d2 = vb_rcp_fast(d * d);
    } else {
        const double r2 = dx2 + dy2;  // r2 = 0 -> log = -inf, 1/r2 = inf in the reference: screened by caller
        lg = 0.5 * vb_log_fast(r2, tab);
        inv_d2 = vb_rcp_fast(r2);
    }
    const double ln_r = three_minus_nu * lg;
    const double q = one_plus_nu * inv_d2;
    g_ee = fma(q, dy2, ln_r);
    g_nn = fma(q, dx2, ln_r);
    g_ne = -q * dx * dy;
}
