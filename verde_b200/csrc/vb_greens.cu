// vb_greens.cu -- K0 (Jacobian materialisation) and K3 (matrix-free predict).
//
// K0  vb_jacobian_kernel / vb_jacobian2d_kernel
//     replaces verde/spline.py:642-650 jacobian_numba and
//     verde/vector.py:461-475 jacobian_2d_numba.  HBM-write bound
//     (8 B per entry, coalesced along the force index).
// K3  vb_predict_kernel / vb_predict2d_kernel
//     replaces verde/spline.py:629-639 predict_numba and
//     verde/vector.py:443-458 predict_2d_numba.  N-body style, FP64-issue bound:
//     forces (coords + coefficient) are staged in shared memory chunk by chunk;
//     a warp owns PTS prediction points, its 32 lanes split the forces of each
//     chunk, partial sums are reduced with warp shuffles, and the CTA's results
//     are written back coalesced through shared memory.
#include "vb_common.cuh"
#include "vb_internal.h"

__device__ double2 vb_logtab_global[VB_LOGTAB_ENTRIES];

// ---------------------------------------------------------------------------
// K0: Jacobian
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vb_jacobian_kernel(const double *__restrict__ east, const double *__restrict__ north, long long n,
                   const double *__restrict__ fe, const double *__restrict__ fn, long long m,
                   double mindist, double *__restrict__ jac)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double fx = fe[j], fy = fn[j];
    for (long long i = blockIdx.y; i < n; i += gridDim.y) {
        jac[i * m + j] = vb_greens_exact(east[i] - fx, north[i] - fy, mindist);
    }
}

__global__ void __launch_bounds__(256)
vb_jacobian2d_kernel(const double *__restrict__ east, const double *__restrict__ north,
                     long long n, const double *__restrict__ fe, const double *__restrict__ fn,
                     long long m, double mindist, double poisson, double *__restrict__ jac)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double fx = fe[j], fy = fn[j];
    const long long ld = 2 * m;
    for (long long i = blockIdx.y; i < n; i += gridDim.y) {
        double ee, nn, ne;
        vb_greens2d_exact(east[i] - fx, north[i] - fy, mindist, poisson, ee, nn, ne);
        jac[i * ld + j] = ee;
        jac[i * ld + j + m] = ne;
        jac[(i + n) * ld + j] = ne;
        jac[(i + n) * ld + j + m] = nn;
    }
}

// ---------------------------------------------------------------------------
// K3: predict
// ---------------------------------------------------------------------------
constexpr int PRED_THREADS = 256;
constexpr int PRED_WARPS = PRED_THREADS / 32;
constexpr int PRED_CHUNK = 1024;  // forces staged per pass: 3 * 8 KB
constexpr int PRED_TAB_REP = 4;    // copies of the log table in the scalar kernel's dynamic smem
constexpr int PRED_TAB_BYTES = VB_LOGTAB_ENTRIES * 16 * PRED_TAB_REP;
constexpr int PRED2_CHUNK = 512;  // 2-D kernel stages 4 arrays; static smem limit is 48 KB

// MODE 0: fast Green's, mindist == 0; MODE 1: fast, mindist != 0; MODE 2: exact.
// THREADS / REP: 256 threads with 4 interleaved table copies (64 KB, two CTAs per SM), or the "wide" flavour
// 512 threads with 8 copies (128 KB, one CTA per SM): copy c = lane % 8 of entry idx sits at double2 index
// idx*8 + c, i.e. in banks 4c..4c+3 whatever idx is, so the 8 lanes of every LDS.128 quarter-warp hit 8
// disjoint bank groups: the data-dependent table lookup becomes bank-conflict free.
template <int PTS, int MODE, int MINB, int THREADS = PRED_THREADS, int REP = PRED_TAB_REP>
__global__ void __launch_bounds__(THREADS, MINB)
vb_predict_kernel(const double *__restrict__ east, const double *__restrict__ north, long long npts,
                  const double *__restrict__ fe, const double *__restrict__ fn,
                  const double *__restrict__ force, int m_begin_stride, int m_total,
                  double mindist, double *__restrict__ out)
{
    extern __shared__ __align__(16) double s_tab[];  // REP interleaved copies of the log table
    __shared__ double s_fe[PRED_CHUNK], s_fn[PRED_CHUNK], s_f[PRED_CHUNK];
    constexpr int WARPS = THREADS / 32;
    __shared__ double s_out[WARPS * PTS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE != 2) vb_logtab_stage<REP>(s_tab, vb_logtab_global);
    const double2 *tab = reinterpret_cast<const double2 *>(s_tab) + (lane & (REP - 1));

    // force range of this split (blockIdx.y): [m0, m1)
    const int m0 = min(m_total, (int)blockIdx.y * m_begin_stride);
    const int m1 = min(m_total, m0 + m_begin_stride);

    const long long p0 = ((long long)blockIdx.x * WARPS + warp) * PTS;
    double px[PTS], py[PTS], acc[PTS];
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        const long long idx = min(p0 + p, npts - 1);
        px[p] = east[idx];
        py[p] = north[idx];
        acc[p] = 0.0;
    }

    // The fast Green's function assumes finite inputs; non-finite forces / force coordinates
    // poison every output of the CTA, non-finite points poison their own output (NaN, as the
    // reference's arithmetic would give).
    int poisoned = 0;
    for (int c0 = m0; c0 < m1; c0 += PRED_CHUNK) {
        const int cn = min(PRED_CHUNK, m1 - c0);
        __syncthreads();
        int bad = 0;
        for (int t = threadIdx.x; t < cn; t += THREADS) {
            const double a = fe[c0 + t], b = fn[c0 + t], c = force[c0 + t];
            s_fe[t] = a;
            s_fn[t] = b;
            s_f[t] = c;
            bad |= !(isfinite(a) && isfinite(b) && isfinite(c));
        }
        poisoned |= __syncthreads_or(bad);
#pragma unroll 2
        for (int j = lane; j < cn; j += 32) {
            const double fx = s_fe[j], fy = s_fn[j], fc = s_f[j];
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                double g;
                if (MODE == 0) g = vb_greens_fast<false, REP>(px[p] - fx, py[p] - fy, mindist, tab);
                else if (MODE == 1) g = vb_greens_fast<true, REP>(px[p] - fx, py[p] - fy, mindist, tab);
                else g = vb_greens_exact(px[p] - fx, py[p] - fy, mindist);
                acc[p] = fma(g, fc, acc[p]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        double v = acc[p];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (MODE != 2 && (poisoned || !(isfinite(px[p]) && isfinite(py[p])))) v = __longlong_as_double(0x7ff8000000000000LL);
        if (lane == 0) s_out[warp * PTS + p] = v;
    }
    __syncthreads();
    // coalesced write of the CTA's WARPS*PTS consecutive points
    const long long base = (long long)blockIdx.x * WARPS * PTS;
    double *dst = out + (long long)blockIdx.y * npts;
    for (int t = threadIdx.x; t < WARPS * PTS; t += THREADS) {
        if (base + t < npts) dst[base + t] = s_out[t];
    }
}

template <int PTS, int MODE>
__global__ void __launch_bounds__(PRED_THREADS)
vb_predict2d_kernel(const double *__restrict__ east, const double *__restrict__ north,
                    long long npts, const double *__restrict__ fe, const double *__restrict__ fn,
                    const double *__restrict__ force, int m_begin_stride, int m_total,
                    double mindist, double poisson, double *__restrict__ out_e,
                    double *__restrict__ out_n)
{
    __shared__ __align__(16) double s_tab[MODE == 2 ? 2 : VB_LOGTAB_SMEM_DOUBLES];
    __shared__ double s_fe[PRED2_CHUNK], s_fn[PRED2_CHUNK], s_f0[PRED2_CHUNK], s_f1[PRED2_CHUNK];
    __shared__ double s_out[2 * PRED_WARPS * PTS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE != 2) vb_logtab_stage(s_tab, vb_logtab_global);
    const double2 *tab = reinterpret_cast<const double2 *>(s_tab);
    const double c3 = 3.0 - poisson, c1 = 1.0 + poisson;

    const int m0 = min(m_total, (int)blockIdx.y * m_begin_stride);
    const int m1 = min(m_total, m0 + m_begin_stride);

    const long long p0 = ((long long)blockIdx.x * PRED_WARPS + warp) * PTS;
    double px[PTS], py[PTS], ae[PTS], an[PTS];
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        const long long idx = min(p0 + p, npts - 1);
        px[p] = east[idx];
        py[p] = north[idx];
        ae[p] = 0.0;
        an[p] = 0.0;
    }
    int poisoned = 0;
    for (int c0 = m0; c0 < m1; c0 += PRED2_CHUNK) {
        const int cn = min(PRED2_CHUNK, m1 - c0);
        __syncthreads();
        int bad = 0;
        for (int t = threadIdx.x; t < cn; t += PRED_THREADS) {
            const double a = fe[c0 + t], b = fn[c0 + t];
            const double c = force[c0 + t];            // east force j
            const double d = force[m_total + c0 + t];  // north force j (verde/vector.py:456-457)
            s_fe[t] = a;
            s_fn[t] = b;
            s_f0[t] = c;
            s_f1[t] = d;
            bad |= !(isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d));
        }
        poisoned |= __syncthreads_or(bad);
        for (int j = lane; j < cn; j += 32) {
            const double fx = s_fe[j], fy = s_fn[j], f0 = s_f0[j], f1 = s_f1[j];
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                double ee, nn, ne;
                if (MODE == 0) vb_greens2d_fast<false>(px[p] - fx, py[p] - fy, mindist, c3, c1, tab, ee, nn, ne);
                else if (MODE == 1) vb_greens2d_fast<true>(px[p] - fx, py[p] - fy, mindist, c3, c1, tab, ee, nn, ne);
                else vb_greens2d_exact(px[p] - fx, py[p] - fy, mindist, poisson, ee, nn, ne);
                ae[p] = fma(ee, f0, fma(ne, f1, ae[p]));
                an[p] = fma(ne, f0, fma(nn, f1, an[p]));
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        double v = ae[p], w = an[p];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            w += __shfl_xor_sync(0xffffffffu, w, off);
        }
        if (MODE != 2 && (poisoned || !(isfinite(px[p]) && isfinite(py[p])))) v = w = __longlong_as_double(0x7ff8000000000000LL);
        if (lane == 0) {
            s_out[warp * PTS + p] = v;
            s_out[PRED_WARPS * PTS + warp * PTS + p] = w;
        }
    }
    __syncthreads();
    const long long base = (long long)blockIdx.x * PRED_WARPS * PTS;
    double *de = out_e + (long long)blockIdx.y * npts;
    double *dn = out_n + (long long)blockIdx.y * npts;
    for (int t = threadIdx.x; t < PRED_WARPS * PTS; t += PRED_THREADS) {
        if (base + t < npts) {
            de[base + t] = s_out[t];
            dn[base + t] = s_out[PRED_WARPS * PTS + t];
        }
    }
}

// ---------------------------------------------------------------------------
// K3 for several coefficient vectors on ONE set of force coordinates: a Vector of Splines on shared
// coordinates (verde/vector.py:114-141 predicts component by component, i.e. re-evaluates every Green's
// function per component).  The Green's function is evaluated once per pair and fed to NRHS accumulators:
// 21 + NRHS FP64 slots per pair instead of NRHS * 22.  Same lane striding and chunking as
// vb_predict_kernel, so every component is bit-identical to its own single predict.
// ---------------------------------------------------------------------------
constexpr int PREDM_PTS = 4;

template <int MODE, int NRHS>
__global__ void __launch_bounds__(PRED_THREADS, 2)
vb_predict_multi_kernel(const double *__restrict__ east, const double *__restrict__ north, long long npts,
                        const double *__restrict__ fe, const double *__restrict__ fn,
                        const double *__restrict__ force /* [NRHS][m_total] */, int m_begin_stride, int m_total,
                        double mindist, double *__restrict__ out /* [NRHS][gridDim.y][npts] */)
{
    constexpr int PTS = PREDM_PTS, REP = PRED_TAB_REP;
    extern __shared__ __align__(16) double s_tab[];
    __shared__ double s_fe[PRED_CHUNK], s_fn[PRED_CHUNK], s_f[NRHS][PRED_CHUNK];
    __shared__ double s_out[NRHS][PRED_WARPS * PTS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE != 2) vb_logtab_stage<REP>(s_tab, vb_logtab_global);
    const double2 *tab = reinterpret_cast<const double2 *>(s_tab) + (lane & (REP - 1));
    const int m0 = min(m_total, (int)blockIdx.y * m_begin_stride);
    const int m1 = min(m_total, m0 + m_begin_stride);
    const long long p0 = ((long long)blockIdx.x * PRED_WARPS + warp) * PTS;
    double px[PTS], py[PTS], acc[PTS][NRHS];
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        const long long idx = min(p0 + p, npts - 1);
        px[p] = east[idx];
        py[p] = north[idx];
#pragma unroll
        for (int q = 0; q < NRHS; ++q) acc[p][q] = 0.0;
    }
    int poisoned = 0;
    for (int c0 = m0; c0 < m1; c0 += PRED_CHUNK) {
        const int cn = min(PRED_CHUNK, m1 - c0);
        __syncthreads();
        int bad = 0;
        for (int t = threadIdx.x; t < cn; t += PRED_THREADS) {
            const double a = fe[c0 + t], b = fn[c0 + t];
            s_fe[t] = a;
            s_fn[t] = b;
            if (!(isfinite(a) && isfinite(b))) bad |= 128;  // non-finite force coordinates poison every component
#pragma unroll
            for (int q = 0; q < NRHS; ++q) {
                const double c = force[(long long)q * m_total + c0 + t];
                s_f[q][t] = c;
                if (!isfinite(c)) bad |= 1 << q;          // a non-finite coefficient only its own component
            }
        }
        // __syncthreads_or only says "some thread saw one": one vote per flag
        if (__syncthreads_or(bad & 128)) poisoned |= 128;
#pragma unroll
        for (int q = 0; q < NRHS; ++q)
            if (__syncthreads_or(bad & (1 << q))) poisoned |= 1 << q;
#pragma unroll 2
        for (int j = lane; j < cn; j += 32) {
            const double fx = s_fe[j], fy = s_fn[j];
            double fc[NRHS];
#pragma unroll
            for (int q = 0; q < NRHS; ++q) fc[q] = s_f[q][j];
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                double g;
                if (MODE == 0) g = vb_greens_fast<false, REP>(px[p] - fx, py[p] - fy, mindist, tab);
                else if (MODE == 1) g = vb_greens_fast<true, REP>(px[p] - fx, py[p] - fy, mindist, tab);
                else g = vb_greens_exact(px[p] - fx, py[p] - fy, mindist);
#pragma unroll
                for (int q = 0; q < NRHS; ++q) acc[p][q] = fma(g, fc[q], acc[p][q]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PTS; ++p) {
        const bool bad_point = !(isfinite(px[p]) && isfinite(py[p]));
#pragma unroll
        for (int q = 0; q < NRHS; ++q) {
            double v = acc[p][q];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (MODE != 2 && (bad_point || (poisoned & (128 | (1 << q))))) v = __longlong_as_double(0x7ff8000000000000LL);
            if (lane == 0) s_out[q][warp * PTS + p] = v;
        }
    }
    __syncthreads();
    const long long base = (long long)blockIdx.x * PRED_WARPS * PTS;
    for (int t = threadIdx.x; t < NRHS * PRED_WARPS * PTS; t += PRED_THREADS) {
        const int q = t / (PRED_WARPS * PTS), i = t % (PRED_WARPS * PTS);
        if (base + i < npts) out[((long long)q * gridDim.y + blockIdx.y) * npts + base + i] = s_out[q][i];
    }
}

// sum the force-split partials in fixed order: out[i] = sum_s part[s][i]
__global__ void vb_reduce_splits_kernel(const double *__restrict__ part, long long npts, int nsplit,
                                        double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += part[(long long)k * npts + i];
    out[i] = s;
}

// ---------------------------------------------------------------------------
// host-side launchers (device pointers)
// ---------------------------------------------------------------------------
// one copy of the __device__ table exists PER DEVICE: readiness is tracked per device ordinal so that a process
// which switches devices (vb200_set_device / torch.cuda.set_device) never reads an all-zero table
static unsigned long long g_tables_ready_mask[4] = {0, 0, 0, 0};  // up to 256 devices

const double2 *vb_logtab_device_ptr()
{
    void *ptr = nullptr;
    cudaGetSymbolAddress(&ptr, vb_logtab_global);
    return static_cast<const double2 *>(ptr);
}

cudaError_t vb_init_tables()
{
    int dev = 0;
    cudaError_t derr = cudaGetDevice(&dev);
    if (derr != cudaSuccess) return derr;
    const bool tracked = dev >= 0 && dev < 256;
    if (tracked && (g_tables_ready_mask[dev >> 6] >> (dev & 63)) & 1ull) return cudaSuccess;
    double2 h[VB_LOGTAB_ENTRIES];
    for (int i = 0; i < VB_LOGTAB_ENTRIES; ++i) {
        const long double c = 1.0L + ((long double)i + 0.5L) / VB_LOGTAB_ENTRIES;
        const double rc = (double)(1.0L / c);
        h[i].x = rc;
        h[i].y = (double)(-logl((long double)rc));
    }
    cudaError_t err = cudaMemcpyToSymbol(vb_logtab_global, h, sizeof(h));
    if (err == cudaSuccess && tracked) g_tables_ready_mask[dev >> 6] |= 1ull << (dev & 63);
    return err;
}

cudaError_t vb_launch_jacobian(const double *e, const double *n, long long npts, const double *fe,
                               const double *fn, long long m, double mindist, double *jac,
                               cudaStream_t stream)
{
    if (npts == 0 || m == 0) return cudaSuccess;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)(npts < 4096 ? npts : 4096));
    vb_jacobian_kernel<<<grid, 256, 0, stream>>>(e, n, npts, fe, fn, m, mindist, jac);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_jacobian2d(const double *e, const double *n, long long npts,
                                 const double *fe, const double *fn, long long m, double mindist,
                                 double poisson, double *jac, cudaStream_t stream)
{
    if (npts == 0 || m == 0) return cudaSuccess;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)(npts < 4096 ? npts : 4096));
    vb_jacobian2d_kernel<<<grid, 256, 0, stream>>>(e, n, npts, fe, fn, m, mindist, poisson, jac);
    ++vb_launch_counter;
    return cudaGetLastError();
}

// Choose points-per-warp and the force split so that small problems still fill
// the 148 SMs; returns the number of splits (partials needed when > 1).
void vb_predict_plan(long long npts, long long m, int *pts, int *nsplit)
{
    int p = 8;
    while (p > 1 && (npts + (long long)PRED_WARPS * p - 1) / ((long long)PRED_WARPS * p) < 2 * 148) p >>= 1;
    long long blocks = (npts + (long long)PRED_WARPS * p - 1) / ((long long)PRED_WARPS * p);
    int s = 1;
    while (blocks * s < 2 * 148 && m / (s * 2) >= 2 * PRED_CHUNK && s < 64) s *= 2;
    if (vb_predict_pts == 1 || vb_predict_pts == 2 || vb_predict_pts == 4 || vb_predict_pts == 8) {
        p = vb_predict_pts;
        blocks = (npts + (long long)PRED_WARPS * p - 1) / ((long long)PRED_WARPS * p);
        s = 1;
        while (blocks * s < 2 * 148 && m / (s * 2) >= 2 * PRED_CHUNK && s < 64) s *= 2;
    }
    *pts = p;
    *nsplit = s;
}

int vb_predict_minb = 2;  // resident CTAs per SM the scalar predict kernel is compiled for (2..4)
int vb_predict_pts = 0;   // 0 = choose from the problem size
int vb_predict_wide = 0;  // 1: 512-thread CTAs with the conflict-free 8-copy log table for the large-problem path

template <int PTS, int MINB>
static cudaError_t launch_predict_pts_minb(int mode, dim3 grid, const double *e, const double *n,
                                           long long npts, const double *fe, const double *fn,
                                           const double *f, int stride, int m, double mindist,
                                           double *out, cudaStream_t stream)
{
    cudaError_t err;
    if (mode == 0) {
        err = cudaFuncSetAttribute(vb_predict_kernel<PTS, 0, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_TAB_BYTES);
        if (err != cudaSuccess) return err;
        vb_predict_kernel<PTS, 0, MINB><<<grid, PRED_THREADS, PRED_TAB_BYTES, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, out);
    } else if (mode == 1) {
        err = cudaFuncSetAttribute(vb_predict_kernel<PTS, 1, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_TAB_BYTES);
        if (err != cudaSuccess) return err;
        vb_predict_kernel<PTS, 1, MINB><<<grid, PRED_THREADS, PRED_TAB_BYTES, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, out);
    } else {
        vb_predict_kernel<PTS, 2, 2><<<grid, PRED_THREADS, 16, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, out);
    }
    return cudaGetLastError();
}

template <int PTS>
static cudaError_t launch_predict_pts(int mode, dim3 grid, const double *e, const double *n,
                                      long long npts, const double *fe, const double *fn,
                                      const double *f, int stride, int m, double mindist,
                                      double *out, cudaStream_t stream)
{
    if (vb_predict_minb >= 4)
        return launch_predict_pts_minb<PTS, 4>(mode, grid, e, n, npts, fe, fn, f, stride, m, mindist, out, stream);
    if (vb_predict_minb == 3)
        return launch_predict_pts_minb<PTS, 3>(mode, grid, e, n, npts, fe, fn, f, stride, m, mindist, out, stream);
    return launch_predict_pts_minb<PTS, 2>(mode, grid, e, n, npts, fe, fn, f, stride, m, mindist, out, stream);
}

// scratch: nsplit*npts doubles when nsplit > 1 (may be null otherwise)
cudaError_t vb_launch_predict(const double *e, const double *n, long long npts, const double *fe,
                              const double *fn, const double *f, long long m, double mindist,
                              int exact, int pts, int nsplit, double *scratch, double *out,
                              cudaStream_t stream)
{
    if (npts == 0) return cudaSuccess;
    if (m == 0) return cudaMemsetAsync(out, 0, sizeof(double) * npts, stream);
    cudaError_t err = vb_init_tables();
    if (err != cudaSuccess) return err;
    const int mode = (exact || !(mindist >= 0.0)) ? 2 : (mindist != 0.0 ? 1 : 0);
    const int stride = (int)((m + nsplit - 1) / nsplit);
    dim3 grid((unsigned)((npts + (long long)PRED_WARPS * pts - 1) / ((long long)PRED_WARPS * pts)), nsplit);
    double *dst = nsplit > 1 ? scratch : out;
    if (vb_predict_wide && pts == 8 && mode != 2) {
        // large problems: one 512-thread CTA per SM with the conflict-free 8-copy table
        constexpr int WT = 512, WREP = 8, WBYTES = VB_LOGTAB_ENTRIES * 16 * WREP;
        dim3 wgrid((unsigned)((npts + (long long)(WT / 32) * 8 - 1) / ((long long)(WT / 32) * 8)), nsplit);
        if (mode == 0) {
            err = cudaFuncSetAttribute(vb_predict_kernel<8, 0, 1, WT, WREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, WBYTES);
            if (err != cudaSuccess) return err;
            vb_predict_kernel<8, 0, 1, WT, WREP><<<wgrid, WT, WBYTES, stream>>>(e, n, npts, fe, fn, f, stride, (int)m, mindist, dst);
        } else {
            err = cudaFuncSetAttribute(vb_predict_kernel<8, 1, 1, WT, WREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, WBYTES);
            if (err != cudaSuccess) return err;
            vb_predict_kernel<8, 1, 1, WT, WREP><<<wgrid, WT, WBYTES, stream>>>(e, n, npts, fe, fn, f, stride, (int)m, mindist, dst);
        }
        err = cudaGetLastError();
    } else
    switch (pts) {
    case 8: err = launch_predict_pts<8>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream); break;
    case 4: err = launch_predict_pts<4>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream); break;
    case 2: err = launch_predict_pts<2>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream); break;
    default: err = launch_predict_pts<1>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream); break;
    }
    if (err != cudaSuccess) return err;
    ++vb_launch_counter;
    if (nsplit > 1) {
        vb_reduce_splits_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, stream>>>(scratch, npts, nsplit, out);
        ++vb_launch_counter;
        err = cudaGetLastError();
    }
    return err;
}

template <int PTS>
static cudaError_t launch_predict2d_pts(int mode, dim3 grid, const double *e, const double *n,
                                        long long npts, const double *fe, const double *fn,
                                        const double *f, int stride, int m, double mindist,
                                        double poisson, double *oe, double *on, cudaStream_t stream)
{
    if (mode == 0)
        vb_predict2d_kernel<PTS, 0><<<grid, PRED_THREADS, 0, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, poisson, oe, on);
    else if (mode == 1)
        vb_predict2d_kernel<PTS, 1><<<grid, PRED_THREADS, 0, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, poisson, oe, on);
    else
        vb_predict2d_kernel<PTS, 2><<<grid, PRED_THREADS, 0, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, poisson, oe, on);
    return cudaGetLastError();
}

// scratch: 2*nsplit*npts doubles when nsplit > 1
cudaError_t vb_launch_predict2d(const double *e, const double *n, long long npts, const double *fe,
                                const double *fn, const double *f, long long m, double mindist,
                                double poisson, int exact, int pts, int nsplit, double *scratch,
                                double *out_e, double *out_n, cudaStream_t stream)
{
    if (npts == 0) return cudaSuccess;
    if (m == 0) {
        cudaError_t e0 = cudaMemsetAsync(out_e, 0, sizeof(double) * npts, stream);
        if (e0 != cudaSuccess) return e0;
        return cudaMemsetAsync(out_n, 0, sizeof(double) * npts, stream);
    }
    cudaError_t err = vb_init_tables();
    if (err != cudaSuccess) return err;
    const int mode = (exact || !(mindist >= 0.0)) ? 2 : (mindist != 0.0 ? 1 : 0);
    const int stride = (int)((m + nsplit - 1) / nsplit);
    if (pts > 4) pts = 4;  // two accumulators per point: keep register pressure bounded
    dim3 grid((unsigned)((npts + (long long)PRED_WARPS * pts - 1) / ((long long)PRED_WARPS * pts)), nsplit);
    double *de = nsplit > 1 ? scratch : out_e;
    double *dn = nsplit > 1 ? scratch + (long long)nsplit * npts : out_n;
    switch (pts) {
    case 4: err = launch_predict2d_pts<4>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, poisson, de, dn, stream); break;
    case 2: err = launch_predict2d_pts<2>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, poisson, de, dn, stream); break;
    default: err = launch_predict2d_pts<1>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, poisson, de, dn, stream); break;
    }
    if (err != cudaSuccess) return err;
    ++vb_launch_counter;
    if (nsplit > 1) {
        vb_launch_counter += 2;
        const unsigned nb = (unsigned)((npts + 255) / 256);
        vb_reduce_splits_kernel<<<nb, 256, 0, stream>>>(de, npts, nsplit, out_e);
        vb_reduce_splits_kernel<<<nb, 256, 0, stream>>>(dn, npts, nsplit, out_n);
        err = cudaGetLastError();
    }
    return err;
}

template <int NRHS>
static cudaError_t launch_predict_multi(int mode, dim3 grid, const double *e, const double *n, long long npts,
                                        const double *fe, const double *fn, const double *f, int stride, int m,
                                        double mindist, double *dst, cudaStream_t stream)
{
    cudaError_t err;
    if (mode == 0) {
        err = cudaFuncSetAttribute(vb_predict_multi_kernel<0, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_TAB_BYTES);
        if (err != cudaSuccess) return err;
        vb_predict_multi_kernel<0, NRHS><<<grid, PRED_THREADS, PRED_TAB_BYTES, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, dst);
    } else if (mode == 1) {
        err = cudaFuncSetAttribute(vb_predict_multi_kernel<1, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_TAB_BYTES);
        if (err != cudaSuccess) return err;
        vb_predict_multi_kernel<1, NRHS><<<grid, PRED_THREADS, PRED_TAB_BYTES, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, dst);
    } else {
        vb_predict_multi_kernel<2, NRHS><<<grid, PRED_THREADS, 16, stream>>>(e, n, npts, fe, fn, f, stride, m, mindist, dst);
    }
    return cudaGetLastError();
}

// f: [nrhs][m] coefficient vectors on the same forces; out: [nrhs][npts]; nrhs = 2 or 3.
// scratch: nrhs*nsplit*npts doubles when nsplit > 1.
cudaError_t vb_launch_predict_multi(const double *e, const double *n, long long npts, const double *fe,
                                    const double *fn, const double *f, long long m, int nrhs, double mindist,
                                    int exact, int nsplit, double *scratch, double *out, cudaStream_t stream)
{
    if (npts == 0) return cudaSuccess;
    if (m == 0) return cudaMemsetAsync(out, 0, sizeof(double) * npts * nrhs, stream);
    if (nrhs != 2 && nrhs != 3) return cudaErrorInvalidValue;
    cudaError_t err = vb_init_tables();
    if (err != cudaSuccess) return err;
    const int mode = (exact || !(mindist >= 0.0)) ? 2 : (mindist != 0.0 ? 1 : 0);
    const int stride = (int)((m + nsplit - 1) / nsplit);
    dim3 grid((unsigned)((npts + (long long)PRED_WARPS * PREDM_PTS - 1) / ((long long)PRED_WARPS * PREDM_PTS)), nsplit);
    double *dst = nsplit > 1 ? scratch : out;
    err = nrhs == 2 ? launch_predict_multi<2>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream)
                    : launch_predict_multi<3>(mode, grid, e, n, npts, fe, fn, f, stride, (int)m, mindist, dst, stream);
    if (err != cudaSuccess) return err;
    ++vb_launch_counter;
    if (nsplit > 1) {
        for (int q = 0; q < nrhs; ++q) {
            vb_reduce_splits_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, stream>>>(scratch + (long long)q * nsplit * npts, npts,
                                                                                   nsplit, out + (long long)q * npts);
            ++vb_launch_counter;
        }
        err = cudaGetLastError();
    }
    return err;
}
