// vb_gram.cu -- K1: Jacobian panel generator + normal-equation assembly.
//
// Replaces the front half of verde/base/least_squares.py:57-70 together with
// the Jacobian build of verde/spline.py:501-538 / verde/vector.py:347-390:
// instead of materialising J (n x m, 20-160 GB at the BASELINE configs), then
// StandardScaler (2-3 passes over J), then Ridge's `X.T @ X` (another copy),
// a slab of <= 2048 rows of sqrt(W) J is generated into an L2/HBM scratch panel
// in DMMA operand layout and immediately consumed by the tile GEMM
// (vb_gemm.cu), which accumulates G = J^T W J.  The same generator pass
// accumulates the vectors the scaler and the right-hand side need:
//   b = J^T W d,  c1 = J^T 1,  c2 = diag(J^T J)   (c1, c2 UNWEIGHTED: they feed
//   sklearn's StandardScaler, which ignores sample weights here).
// Afterwards vb_scale_system_kernel applies the reference's algebra
// (SURVEY.md fact 5):  s_j = sqrt(c2_j/n - (c1_j/n)^2) with sklearn's
// constant-column rule, A = D G D + damping*I, rhs = D b, D = diag(1/s).
//
// Generation cost is n*m Green's evaluations ONCE (about 0.3% of the FP64 work of
// the Gram product at m = 50k), which is why the exact (reference-ordered)
// Green's function is affordable here.
#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

constexpr int GEN_THREADS = 256;

// entry (row r, column c) of the system Jacobian, before weighting
__device__ __forceinline__ double jac_entry(const VbGramProblem &pb, long long r, long long c,
                                            double ex, double ny, double fx, double fy)
{
    if (pb.kind == 0) return vb_greens_exact(ex - fx, ny - fy, pb.mindist);
    double ee, nn, ne;
    vb_greens2d_exact(ex - fx, ny - fy, pb.mindist, pb.poisson, ee, nn, ne);
    const bool north_row = r >= pb.n, north_col = c >= pb.m;
    return north_row ? (north_col ? nn : ne) : (north_col ? ne : ee);
}

// grid.x = T (force block), one CTA loops over the slab's chunks so that each
// stats column is owned by exactly one CTA per launch (deterministic sums).
__global__ void __launch_bounds__(GEN_THREADS)
vb_gram_panel_kernel(const VbGramProblem pb, long long row0, int nchunk, int T, long long Mpad,
                     double *__restrict__ panel, double *__restrict__ stats)
{
    __shared__ double s_red[3][GEN_THREADS];
    const long long rows = pb.kind == 1 ? 2 * pb.n : pb.n;
    const long long cols = pb.kind == 1 ? 2 * pb.m : pb.m;
    const int bcol = blockIdx.x;
    const int i = threadIdx.x & (VB_TILE - 1);   // column inside the block
    const int khalf = threadIdx.x >> 7;          // 0/1: which k parity this thread generates
    const long long c = (long long)bcol * VB_TILE + i;
    const bool col_ok = c < cols;
    const long long cf = col_ok ? (pb.kind == 1 ? (c >= pb.m ? c - pb.m : c) : c) : 0;
    const double fx = pb.kind == 2 ? 0.0 : pb.fe[cf], fy = pb.kind == 2 ? 0.0 : pb.fn[cf];

    double sb = 0.0, s1 = 0.0, s2 = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
        double *dst = panel + ((long long)ch * T + bcol) * VB_TILE_ELEMS;
#pragma unroll 4
        for (int kk = 0; kk < VB_TILE / 2; ++kk) {
            const int k = kk * 2 + khalf;
            const long long r = row0 + (long long)ch * VB_TILE + k;
            double v = 0.0;
            if (col_ok && r < rows) {
                const long long rp = pb.kind == 1 ? (r >= pb.n ? r - pb.n : r) : r;
                const double j = pb.kind == 2 ? pb.jac[r * cols + c]
                                              : jac_entry(pb, r, c, pb.east[rp], pb.north[rp], fx, fy);
                const double w = pb.weights ? pb.weights[r] : 1.0;
                s1 += j;
                s2 = fma(j, j, s2);
                sb = fma(w * j, pb.data[r], sb);
                v = pb.weights ? sqrt(w) * j : j;
            }
            if (panel) dst[k * VB_TILE + i] = v;  // panel == nullptr: statistics-only pass (fused Gram variant)
        }
    }
    s_red[0][threadIdx.x] = sb;
    s_red[1][threadIdx.x] = s1;
    s_red[2][threadIdx.x] = s2;
    __syncthreads();
    if (khalf == 0 && col_ok) {
        stats[0 * Mpad + c] += s_red[0][i] + s_red[0][i + VB_TILE];
        stats[1 * Mpad + c] += s_red[1][i] + s_red[1][i + VB_TILE];
        stats[2 * Mpad + c] += s_red[2][i] + s_red[2][i + VB_TILE];
    }
}

// inv_scale[j] = 1/s_j following sklearn/preprocessing/_data.py:83-133,1042-1078
__global__ void vb_column_scale_kernel(const double *__restrict__ stats, long long Mpad,
                                       long long cols, double nrows, double *__restrict__ inv_scale,
                                       double *__restrict__ rhs)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    double inv = 1.0, b = 0.0;
    if (j < cols) {
        const double mean = stats[1 * Mpad + j] / nrows;
        double var = stats[2 * Mpad + j] / nrows - mean * mean;
        if (var < 0.0) var = 0.0;
        const double eps = 2.220446049250313e-16;
        const double nm = nrows * mean * eps;
        const bool constant = var <= nrows * eps * var + nm * nm;
        double s = constant ? 0.0 : sqrt(var);
        if (constant || s < 10.0 * eps) s = 1.0;
        inv = 1.0 / s;
        b = stats[0 * Mpad + j] * inv;
    }
    inv_scale[j] = inv;
    rhs[j] = b;
}

// A = D G D + damping*I on the tile-packed lower triangle; padded columns get a
// unit diagonal so the factorisation of the padded system is the identity there.
__global__ void __launch_bounds__(256)
vb_scale_system_kernel(double *__restrict__ G, int T, long long cols,
                       const double *__restrict__ inv_scale, double damping)
{
    // blockIdx.x = tile index (block-column-major), decode (bi, bj)
    long long t = blockIdx.x;
    int bj = 0;
    while (t >= T - bj) {
        t -= T - bj;
        ++bj;
    }
    const int bi = bj + (int)t;
    double *tile = G + (long long)blockIdx.x * VB_TILE_ELEMS;
    const int r = threadIdx.x & (VB_TILE - 1);
    const long long gr = (long long)bi * VB_TILE + r;
    const double sr = inv_scale[gr];
    for (int cc = threadIdx.x >> 7; cc < VB_TILE; cc += 2) {
        const long long gc = (long long)bj * VB_TILE + cc;
        double v = tile[cc * VB_TILE + r] * sr * inv_scale[gc];
        if (gr == gc) v = (gr < cols) ? v + damping : 1.0;
        tile[cc * VB_TILE + r] = v;
    }
}

// Same algebra for a damping sweep: G is read ONCE and one scaled, damped copy is written per
// damping (L + b * stride); G itself is left untouched (SplineCV re-uses it).
__global__ void __launch_bounds__(256)
vb_scale_system_batched_kernel(const double *__restrict__ G, int T, long long cols,
                               const double *__restrict__ inv_scale, const double *__restrict__ dampings,
                               int batch, double *__restrict__ L, long long stride)
{
    long long t = blockIdx.x;
    int bj = 0;
    while (t >= T - bj) {
        t -= T - bj;
        ++bj;
    }
    const int bi = bj + (int)t;
    const long long toff = (long long)blockIdx.x * VB_TILE_ELEMS;
    const int r = threadIdx.x & (VB_TILE - 1);
    const long long gr = (long long)bi * VB_TILE + r;
    const double sr = inv_scale[gr];
    for (int cc = threadIdx.x >> 7; cc < VB_TILE; cc += 2) {
        const long long gc = (long long)bj * VB_TILE + cc;
        const double v = G[toff + cc * VB_TILE + r] * sr * inv_scale[gc];
        for (int b = 0; b < batch; ++b) {
            double vb = v;
            if (gr == gc) vb = (gr < cols) ? v + dampings[b] : 1.0;
            L[(long long)b * stride + toff + cc * VB_TILE + r] = vb;
        }
    }
}

}  // namespace

cudaError_t vb_launch_gram_panel(const VbGramProblem &prob, long long row0, int nchunk, int T,
                                 double *panel, double *stats, cudaStream_t stream)
{
    vb_gram_panel_kernel<<<T, GEN_THREADS, 0, stream>>>(prob, row0, nchunk, T, (long long)T * VB_TILE,
                                                        panel, stats);
    ++vb_launch_counter;
    return cudaGetLastError();
}

__global__ void vb_sqrt_kernel(const double *__restrict__ w, long long n, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sqrt(w[i]);
}

cudaError_t vb_launch_sqrt(const double *w, long long n, double *out, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    vb_sqrt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, n, out);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_column_scale(const double *stats, int T, long long cols, long long nrows,
                                   double *inv_scale, double *rhs, cudaStream_t stream)
{
    const long long Mpad = (long long)T * VB_TILE;
    vb_column_scale_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, stream>>>(stats, Mpad, cols, (double)nrows,
                                                                               inv_scale, rhs);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_scale_system(double *G, int T, long long cols, const double *inv_scale,
                                   double damping, cudaStream_t stream)
{
    vb_scale_system_kernel<<<(unsigned)vb_num_tiles(T), 256, 0, stream>>>(G, T, cols, inv_scale, damping);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_launch_scale_system_batched(const double *G, int T, long long cols, const double *inv_scale,
                                           const double *dampings_dev, int batch, double *L,
                                           long long stride, cudaStream_t stream)
{
    vb_scale_system_batched_kernel<<<(unsigned)vb_num_tiles(T), 256, 0, stream>>>(G, T, cols, inv_scale,
                                                                                  dampings_dev, batch, L, stride);
    ++vb_launch_counter;
    return cudaGetLastError();
}
