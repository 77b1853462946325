// vb_gemm.cu -- the FP64 DMMA tile kernel shared by the Gram accumulation
// (G += P^T P over Jacobian panels) and the Cholesky updates (C -= L L^T).
//
// Replaces the BLAS-3 the reference reaches through scikit-learn
// (sklearn/linear_model/_ridge.py:215-227 `safe_sparse_dot(X.T, X)` -> dsyrk/dgemm)
// and LAPACK potrf's trailing updates (scipy.linalg.solve(assume_a="pos")).
//
// Roofline: FP64 pipe.  On sm_100a every f64 mma.sync shape lowers to
// DMMA.8x8x4 and the DMMA and DFMA paths share one 64-FMA/clk/SM pipe
// (tools/fp64_peak.cu: 37.2 TFLOP/s DMMA, 36.8 DFMA, no concurrency), so the
// job of this kernel is to keep that pipe issuing DMMA back to back:
//   * CTA tile 128x128 (half the register file holds the accumulators),
//     8 warps as 2x4, warp tile 64x32 = 8x4 DMMA.8x8x4 per k4 step;
//   * operands are k-major 128-wide chunks -> a k16 stage is 2 x 16 rows of 1 KB,
//     brought in by cp.async into a 4-stage ring of padded rows (stride 132
//     doubles => the 8x4 / 4x8 fragment loads are bank-conflict free);
//   * per CTA the L2->SM operand traffic is 8 B/clk (2 KB per 256 DMMA-cycles),
//     ~19% of the L2 bandwidth chip-wide; tiles are enumerated in 8-column
//     strips, row-major inside a strip, so the ~148 concurrently resident CTAs
//     share ~26 operand panels in L2.
#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

constexpr int KC = 16;                 // k rows per pipeline stage
constexpr int LDS_STRIDE = VB_TILE + 4;  // doubles per padded smem row
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int STRIP = 8;  // block columns per rasterisation strip
constexpr int STAGE_DOUBLES = 2 * KC * LDS_STRIDE;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_DOUBLES * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// tiles in the strip starting at block column c0 with width w (rows c0..T-1)
__host__ __device__ __forceinline__ long long strip_tiles(int T, int c0, int w)
{
    return (long long)w * (w + 1) / 2 + (long long)w * (T - c0 - w);
}

// linear tile id -> (bi, bj)
__device__ __forceinline__ void decode_tile(long long t, int T, int j_lo, int j_hi, int &bi, int &bj)
{
    int c0 = j_lo;
    int w = min(STRIP, j_hi - c0);
    long long cnt = strip_tiles(T, c0, w);
    while (t >= cnt) {
        t -= cnt;
        c0 += w;
        w = min(STRIP, j_hi - c0);
        cnt = strip_tiles(T, c0, w);
    }
    const long long head = (long long)w * (w + 1) / 2;
    if (t < head) {
        // triangular head: row r (0-based inside the strip) has r+1 tiles
        int r = 0;
        while (t >= r + 1) {
            t -= r + 1;
            ++r;
        }
        bi = c0 + r;
        bj = c0 + (int)t;
    } else {
        t -= head;
        bi = c0 + w + (int)(t / w);
        bj = c0 + (int)(t % w);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) vb_gemm_kernel(const VbGemmParams p)
{
    extern __shared__ __align__(16) double smem[];

    int bi, bj;
    decode_tile(blockIdx.x, p.T, p.j_lo, p.j_hi, bi, bj);
    const bool diag = (bi == bj) && (p.a_base == p.b_base);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int p0 = (warp >> 2) * 64;  // A-operand (C column) offset of the warp tile
    const int q0 = (warp & 3) * 32;   // B-operand (C row) offset of the warp tile
    const int lk = lane & 3, lr = lane >> 2;

    const int nsteps = p.nk * (VB_TILE / KC);

    auto issue_stage = [&](int step) {
        const int slot = step % STAGES;
        const int kk = step / (VB_TILE / KC);
        const int krow0 = (step % (VB_TILE / KC)) * KC;
        double *sA = smem + slot * STAGE_DOUBLES;
        double *sB = sA + KC * LDS_STRIDE;
        const double *gA = p.a_base + p.a_chunk_off[kk] + (long long)bj * VB_TILE_ELEMS + krow0 * VB_TILE;
        const double *gB = p.b_base + p.b_chunk_off[kk] + (long long)bi * VB_TILE_ELEMS + krow0 * VB_TILE;
        // KC rows x 64 sixteen-byte segments per operand
#pragma unroll
        for (int r = 0; r < (KC * 64) / GEMM_THREADS; ++r) {
            const int lin = tid + r * GEMM_THREADS;
            const int k = lin >> 6, seg = lin & 63;
            cp_async16(sA + k * LDS_STRIDE + seg * 2, gA + k * VB_TILE + seg * 2);
        }
        if (!diag) {
#pragma unroll
            for (int r = 0; r < (KC * 64) / GEMM_THREADS; ++r) {
                const int lin = tid + r * GEMM_THREADS;
                const int k = lin >> 6, seg = lin & 63;
                cp_async16(sB + k * LDS_STRIDE + seg * 2, gB + k * VB_TILE + seg * 2);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nsteps) issue_stage(s);
        cp_async_commit();
    }

    for (int step = 0; step < nsteps; ++step) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (step + STAGES - 1 < nsteps) issue_stage(step + STAGES - 1);
        cp_async_commit();

        const double *sA = smem + (step % STAGES) * STAGE_DOUBLES;
        const double *sB = diag ? sA : sA + KC * LDS_STRIDE;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            const double *ra = sA + (k4 * 4 + lk) * LDS_STRIDE + p0 + lr;
            const double *rb = sB + (k4 * 4 + lk) * LDS_STRIDE + q0 + lr;
            double fa[8], fb[4];
#pragma unroll
            for (int a = 0; a < 8; ++a) fa[a] = ra[a * 8];
#pragma unroll
            for (int b = 0; b < 4; ++b) fb[b] = rb[b * 8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
    }
    cp_async_wait<0>();

    // epilogue: acc[a][b][0..1] = D[p][q], p = p0 + 8a + lr (C column), q = q0 + 8b + 2lk (+1) (C row)
    double *ctile = p.c_base + vb_tile_index(bi, bj, p.T) * VB_TILE_ELEMS;
    const double alpha = p.alpha;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        double2 *dst[4];
        double2 old[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            dst[b] = reinterpret_cast<double2 *>(ctile + (p0 + 8 * a + lr) * VB_TILE + q0 + 8 * b + 2 * lk);
            if (p.accumulate) old[b] = *dst[b];
            else old[b] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            old[b].x = fma(alpha, acc[a][b][0], old[b].x);
            old[b].y = fma(alpha, acc[a][b][1], old[b].y);
            *dst[b] = old[b];
        }
    }
}

}  // namespace

long long vb_launch_counter = 0;

long long vb_gemm_num_tiles(int T, int j_lo, int j_hi)
{
    long long total = 0;
    for (int c0 = j_lo; c0 < j_hi; c0 += STRIP) {
        const int w = (j_hi - c0 < STRIP) ? (j_hi - c0) : STRIP;
        total += strip_tiles(T, c0, w);
    }
    return total;
}

cudaError_t vb_launch_gemm(const VbGemmParams &p, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(vb_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (p.nk <= 0 || p.nk > VB_MAX_KCHUNKS) return cudaErrorInvalidValue;
    const long long tiles = vb_gemm_num_tiles(p.T, p.j_lo, p.j_hi);
    if (tiles <= 0) return cudaSuccess;
    vb_gemm_kernel<<<(unsigned)tiles, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    ++vb_launch_counter;
    return cudaGetLastError();
}
