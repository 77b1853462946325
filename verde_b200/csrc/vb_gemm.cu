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
__device__ __forceinline__ void decode_tile(long long t, int T, int j_lo, int j_hi, int STRIP, int &bi, int &bj)
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

// tile id of a launch that may cover several block-column ranges (VbGemmParams::nranges)
__device__ __forceinline__ void decode_tile_ranges(const VbGemmParams &p, long long t, int &bi, int &bj)
{
    if (p.nranges > 0) {
        int r = 0;
        while (r + 1 < p.nranges && t >= p.range_first[r + 1]) ++r;
        decode_tile(t - p.range_first[r], p.T, p.range_lo[r], p.range_hi[r], p.strip, bi, bj);
    } else {
        decode_tile(t, p.T, p.j_lo, p.j_hi, p.strip, bi, bj);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) vb_gemm_kernel(const VbGemmParams p)
{
    extern __shared__ __align__(16) double smem[];

    int bi, bj;
    decode_tile(blockIdx.x, p.T, p.j_lo, p.j_hi, p.strip, bi, bj);
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

// ===========================================================================
// v2: Blackwell-style asynchronous pipeline.
//   * one PRODUCER warp streams operand rows with TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a ring of
//     padded stages guarded by full/empty mbarriers;
//   * eight CONSUMER warps never meet a CTA-wide barrier in the main loop: each
//     waits on full[s], issues its 32 DMMA per k4 step, and arrives on empty[s];
//   * epilogue: accumulators are staged in shared memory in the tile's own
//     layout and handed to the TMA engine as ONE asynchronous reduce-add into
//     the tile in HBM/L2 (cp.reduce.async.bulk ... add.f64): no read-modify-write
//     round trip through the SM, which was ~8 us per tile in v1 (12% of a K=512
//     Cholesky update).
// ===========================================================================
constexpr int V2_KC = 32;
constexpr int V2_STAGES = 3;
constexpr int V2_STAGE_DOUBLES = 2 * V2_KC * LDS_STRIDE;
constexpr int V2_RING_BYTES = V2_STAGES * V2_STAGE_DOUBLES * (int)sizeof(double);
constexpr int V2_SMEM_BYTES = V2_RING_BYTES + 64;  // + mbarriers
static_assert(V2_RING_BYTES >= VB_TILE_ELEMS * (int)sizeof(double), "epilogue staging reuses the ring");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 28)) __trap();  // a byte-count bug must fault, not hang the box
    }
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f64(void *gdst, unsigned ssrc, unsigned bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, unsigned ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
}

// NCW consumer warps: 8 -> warp tile 64x32 (2x4 warps), 16 -> warp tile 32x32 (4x4 warps)
template <int NCW>
__global__ void __launch_bounds__(NCW * 32 + 32, 1) vb_gemm_tma_kernel(const VbGemmParams p)
{
    constexpr int V2_CONSUMERS = NCW * 32;
    constexpr int MI = (NCW == 8) ? 8 : 4;  // A fragments (8 columns of C each) per warp
    extern __shared__ __align__(128) double smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + V2_STAGES * V2_STAGE_DOUBLES);
    // bars[0..S) = full, bars[S..2S) = empty

    int bi, bj;
    decode_tile(blockIdx.x, p.T, p.j_lo, p.j_hi, p.strip, bi, bj);
    const bool diag = (bi == bj) && (p.a_base == p.b_base);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int STEPS_PER_CHUNK = VB_TILE / V2_KC;
    const int nsteps = p.nk * STEPS_PER_CHUNK;

    if (tid == 0) {
        for (int s = 0; s < V2_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                          // producer's arrive.expect_tx
            mbar_init(smem_u32(&bars[V2_STAGES + s]), V2_CONSUMERS / 32);  // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == V2_CONSUMERS / 32) {
        // ---------------- producer warp ----------------
        const unsigned stage_bytes = (diag ? 1u : 2u) * V2_KC * VB_TILE * (unsigned)sizeof(double);
        for (int step = 0; step < nsteps; ++step) {
            const int slot = step % V2_STAGES;
            const unsigned ph = (unsigned)(step / V2_STAGES) & 1u;
            mbar_wait(smem_u32(&bars[V2_STAGES + slot]), ph ^ 1u);
            const unsigned full = smem_u32(&bars[slot]);
            if (lane == 0) mbar_expect_tx(full, stage_bytes);
            __syncwarp();
            const int kk = step / STEPS_PER_CHUNK;
            const int krow0 = (step % STEPS_PER_CHUNK) * V2_KC;
            double *sA = smem + slot * V2_STAGE_DOUBLES;
            double *sB = sA + V2_KC * LDS_STRIDE;
            const double *gA = p.a_base + p.a_chunk_off[kk] + (long long)bj * VB_TILE_ELEMS + krow0 * VB_TILE;
            const double *gB = p.b_base + p.b_chunk_off[kk] + (long long)bi * VB_TILE_ELEMS + krow0 * VB_TILE;
            // lane r copies row r of A (1 KB) and, off the diagonal, row r of B
            bulk_load(smem_u32(sA + lane * LDS_STRIDE), gA + lane * VB_TILE, VB_TILE * sizeof(double), full);
            if (!diag)
                bulk_load(smem_u32(sB + lane * LDS_STRIDE), gB + lane * VB_TILE, VB_TILE * sizeof(double), full);
        }
        // the ring is reused as epilogue staging: make sure every load has landed and been consumed
        // (consumers only leave the main loop after their last full-wait) -- nothing more to do here.
    } else {
        // ---------------- consumer warps ----------------
        const int p0 = (warp >> 2) * (MI * 8);
        const int q0 = (warp & 3) * 32;
        const int lk = lane & 3, lr = lane >> 2;
        double acc[MI][4][2];
#pragma unroll
        for (int a = 0; a < MI; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

        for (int step = 0; step < nsteps; ++step) {
            const int slot = step % V2_STAGES;
            const unsigned ph = (unsigned)(step / V2_STAGES) & 1u;
            mbar_wait(smem_u32(&bars[slot]), ph);
            const double *sA = smem + slot * V2_STAGE_DOUBLES;
            const double *sB = diag ? sA : sA + V2_KC * LDS_STRIDE;
#pragma unroll
            for (int k4 = 0; k4 < V2_KC / 4; ++k4) {
                const double *ra = sA + (k4 * 4 + lk) * LDS_STRIDE + p0 + lr;
                const double *rb = sB + (k4 * 4 + lk) * LDS_STRIDE + q0 + lr;
                double fa[MI], fb[4];
#pragma unroll
                for (int a = 0; a < MI; ++a) fa[a] = ra[a * 8];
#pragma unroll
                for (int b = 0; b < 4; ++b) fb[b] = rb[b * 8];
#pragma unroll
                for (int a = 0; a < MI; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[V2_STAGES + slot]));
        }

        // ---- epilogue: stage alpha*acc in the tile's own column-major layout, then one TMA reduce ----
        asm volatile("bar.sync 1, %0;" ::"n"(V2_CONSUMERS) : "memory");  // all consumers are done reading the ring
        const double alpha = p.alpha;
#pragma unroll
        for (int a = 0; a < MI; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double2 v = make_double2(alpha * acc[a][b][0], alpha * acc[a][b][1]);
                *reinterpret_cast<double2 *>(smem + (p0 + 8 * a + lr) * VB_TILE + q0 + 8 * b + 2 * lk) = v;
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(V2_CONSUMERS) : "memory");
        if (tid == 0) {
            double *ctile = p.c_base + vb_tile_index(bi, bj, p.T) * VB_TILE_ELEMS;
            constexpr unsigned PIECE = 16384;  // bytes per bulk operation
#pragma unroll 1
            for (unsigned off = 0; off < VB_TILE_ELEMS * sizeof(double); off += PIECE) {
                if (p.accumulate) bulk_reduce_add_f64(reinterpret_cast<char *>(ctile) + off, smem_u32(smem) + off, PIECE);
                else bulk_store(reinterpret_cast<char *>(ctile) + off, smem_u32(smem) + off, PIECE);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

// ===========================================================================
// v3: persistent version of v2.  One CTA per SM loops over tiles (t = blockIdx.x,
// += gridDim.x, so concurrently resident CTAs still work on neighbouring tiles).
// The producer warp keeps streaming stages ACROSS tile boundaries, so the pipeline
// never drains: no per-tile prologue bubble.  The epilogue cannot borrow the ring
// (it is already filling with the next tile), so each consumer warp commits its
// accumulators straight from registers with fire-and-forget RED.ADD.F64 (exactly
// one add per element per launch -> still deterministic) and moves on.
// ===========================================================================
__global__ void __launch_bounds__(8 * 32 + 32, 1)
vb_gemm_persistent_kernel(const VbGemmParams p, long long ntiles)
{
    constexpr int NCW = 8, MI = 8;
    extern __shared__ __align__(128) double smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + V2_STAGES * V2_STAGE_DOUBLES);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int STEPS_PER_CHUNK = VB_TILE / V2_KC;
    const int nsteps = p.nk * STEPS_PER_CHUNK;
    const long long total = ntiles * (p.batch > 1 ? p.batch : 1);  // ntiles = tiles of ONE matrix

    if (tid == 0) {
        for (int s = 0; s < V2_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[V2_STAGES + s]), NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long gstep = 0;  // ring position, continues across tiles
    if (warp == NCW) {
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            int bi, bj;
            const long long mat = t / ntiles;
            decode_tile(t - mat * ntiles, p.T, p.j_lo, p.j_hi, p.strip, bi, bj);
            const long long moff = mat * p.batch_stride;
            const bool diag = (bi == bj) && (p.a_base == p.b_base);
            const unsigned stage_bytes = (diag ? 1u : 2u) * V2_KC * VB_TILE * (unsigned)sizeof(double);
            for (int step = 0; step < nsteps; ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[V2_STAGES + slot]), ph ^ 1u);
                const unsigned full = smem_u32(&bars[slot]);
                if (lane == 0) mbar_expect_tx(full, stage_bytes);
                __syncwarp();
                const int kk = step / STEPS_PER_CHUNK;
                const int krow0 = (step % STEPS_PER_CHUNK) * V2_KC;
                double *sA = smem + slot * V2_STAGE_DOUBLES;
                double *sB = sA + V2_KC * LDS_STRIDE;
                const double *gA = p.a_base + moff + p.a_chunk_off[kk] + (long long)bj * VB_TILE_ELEMS + krow0 * VB_TILE;
                const double *gB = p.b_base + moff + p.b_chunk_off[kk] + (long long)bi * VB_TILE_ELEMS + krow0 * VB_TILE;
                bulk_load(smem_u32(sA + lane * LDS_STRIDE), gA + lane * VB_TILE, VB_TILE * sizeof(double), full);
                if (!diag)
                    bulk_load(smem_u32(sB + lane * LDS_STRIDE), gB + lane * VB_TILE, VB_TILE * sizeof(double), full);
            }
        }
    } else {
        const int p0 = (warp >> 2) * (MI * 8);
        const int q0 = (warp & 3) * 32;
        const int lk = lane & 3, lr = lane >> 2;
        const double alpha = p.alpha;
        for (long long t = blockIdx.x; t < total; t += gridDim.x) {
            int bi, bj;
            const long long mat = t / ntiles;
            decode_tile(t - mat * ntiles, p.T, p.j_lo, p.j_hi, p.strip, bi, bj);
            const bool diag = (bi == bj) && (p.a_base == p.b_base);
            double acc[MI][4][2];
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            for (int step = 0; step < nsteps; ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[slot]), ph);
                const double *sA = smem + slot * V2_STAGE_DOUBLES;
                const double *sB = diag ? sA : sA + V2_KC * LDS_STRIDE;
#pragma unroll
                for (int k4 = 0; k4 < V2_KC / 4; ++k4) {
                    const double *ra = sA + (k4 * 4 + lk) * LDS_STRIDE + p0 + lr;
                    const double *rb = sB + (k4 * 4 + lk) * LDS_STRIDE + q0 + lr;
                    double fa[MI], fb[4];
#pragma unroll
                    for (int a = 0; a < MI; ++a) fa[a] = ra[a * 8];
#pragma unroll
                    for (int b = 0; b < 4; ++b) fb[b] = rb[b * 8];
#pragma unroll
                    for (int a = 0; a < MI; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[V2_STAGES + slot]));
            }
            double *ctile = p.c_base + mat * p.batch_stride + vb_tile_index(bi, bj, p.T) * VB_TILE_ELEMS;
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double *dst = ctile + (p0 + 8 * a + lr) * VB_TILE + q0 + 8 * b + 2 * lk;
                    if (p.accumulate) {
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(alpha * acc[a][b][0]) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 1), "d"(alpha * acc[a][b][1]) : "memory");
                    } else {
                        *reinterpret_cast<double2 *>(dst) = make_double2(alpha * acc[a][b][0], alpha * acc[a][b][1]);
                    }
                }
        }
    }
}

// ===========================================================================
// v4: v3 with a DYNAMIC tile scheduler.  Tiles are handed out by one atomic counter in the same rasterised order
// (so the L2 behaviour of v3 is kept); the producer warp draws the next tile id and passes it to the consumer
// warps through a small shared-memory ring published by the full-barrier of the tile's first stage.  What it buys:
//   * a CTA that starts late -- because a panel kernel or an NCCL kernel of another stream still holds its SM --
//     simply draws fewer tiles instead of stalling the launch on its static share (look-ahead, multi-GPU overlap);
//   * several launches (grids) can drain ONE queue: SMs released by a concurrent panel factorisation join in.
// Each tile is still computed by exactly one CTA with the same arithmetic: results are bit-identical to v3.
// ===========================================================================
constexpr int V4_RING = 4;  // tile ids in flight between producer and consumers (the producer runs <= 1 tile ahead)

__global__ void __launch_bounds__(8 * 32 + 32, 1)
vb_gemm_dynamic_kernel(const VbGemmParams p, long long ntiles, unsigned long long *__restrict__ counter)
{
    constexpr int NCW = 8, MI = 8;
    extern __shared__ __align__(128) double smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + V2_STAGES * V2_STAGE_DOUBLES);
    long long *tile_ring = reinterpret_cast<long long *>(bars + 2 * V2_STAGES);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int STEPS_PER_CHUNK = VB_TILE / V2_KC;
    const int nsteps = p.nk * STEPS_PER_CHUNK;
    const long long total = ntiles * (p.batch > 1 ? p.batch : 1);

    if (tid == 0) {
        for (int s = 0; s < V2_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[V2_STAGES + s]), NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long gstep = 0;  // ring position, continues across tiles
    if (warp == NCW) {
        for (long long tcount = 0;; ++tcount) {
            long long t = 0;
            if (lane == 0) t = (long long)atomicAdd(counter, 1ull);
            t = __shfl_sync(0xffffffffu, t, 0);
            const bool done = t >= total;
            int bi = 0, bj = 0;
            long long moff = 0;
            bool diag = false;
            if (!done) {
                const long long mat = t / ntiles;
                decode_tile_ranges(p, t - mat * ntiles, bi, bj);
                moff = mat * p.batch_stride;
                diag = (bi == bj) && (p.a_base == p.b_base);
            }
            const unsigned stage_bytes = (diag ? 1u : 2u) * V2_KC * VB_TILE * (unsigned)sizeof(double);
            for (int step = 0; step < (done ? 1 : nsteps); ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[V2_STAGES + slot]), ph ^ 1u);
                const unsigned full = smem_u32(&bars[slot]);
                if (step == 0 && lane == 0) tile_ring[tcount % V4_RING] = done ? -1 : t;  // published by the arrive below
                if (done) {
                    if (lane == 0) mbar_arrive(full);  // an empty stage: only carries the end-of-queue marker
                    break;
                }
                if (lane == 0) mbar_expect_tx(full, stage_bytes);
                __syncwarp();
                const int kk = step / STEPS_PER_CHUNK;
                const int krow0 = (step % STEPS_PER_CHUNK) * V2_KC;
                double *sA = smem + slot * V2_STAGE_DOUBLES;
                double *sB = sA + V2_KC * LDS_STRIDE;
                const double *gA = p.a_base + moff + p.a_chunk_off[kk] + (long long)bj * VB_TILE_ELEMS + krow0 * VB_TILE;
                const double *gB = p.b_base + moff + p.b_chunk_off[kk] + (long long)bi * VB_TILE_ELEMS + krow0 * VB_TILE;
                bulk_load(smem_u32(sA + lane * LDS_STRIDE), gA + lane * VB_TILE, VB_TILE * sizeof(double), full);
                if (!diag)
                    bulk_load(smem_u32(sB + lane * LDS_STRIDE), gB + lane * VB_TILE, VB_TILE * sizeof(double), full);
            }
            if (done) break;
        }
    } else {
        const int p0 = (warp >> 2) * (MI * 8);
        const int q0 = (warp & 3) * 32;
        const int lk = lane & 3, lr = lane >> 2;
        const double alpha = p.alpha;
        for (long long tcount = 0;; ++tcount) {
            // the first stage of a tile also publishes its id
            {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[slot]), ph);
            }
            const long long t = tile_ring[tcount % V4_RING];
            if (t < 0) break;
            int bi, bj;
            const long long mat = t / ntiles;
            decode_tile_ranges(p, t - mat * ntiles, bi, bj);
            const bool diag = (bi == bj) && (p.a_base == p.b_base);
            double acc[MI][4][2];
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            for (int step = 0; step < nsteps; ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                if (step > 0) mbar_wait(smem_u32(&bars[slot]), ph);
                const double *sA = smem + slot * V2_STAGE_DOUBLES;
                const double *sB = diag ? sA : sA + V2_KC * LDS_STRIDE;
#pragma unroll
                for (int k4 = 0; k4 < V2_KC / 4; ++k4) {
                    const double *ra = sA + (k4 * 4 + lk) * LDS_STRIDE + p0 + lr;
                    const double *rb = sB + (k4 * 4 + lk) * LDS_STRIDE + q0 + lr;
                    double fa[MI], fb[4];
#pragma unroll
                    for (int a = 0; a < MI; ++a) fa[a] = ra[a * 8];
#pragma unroll
                    for (int b = 0; b < 4; ++b) fb[b] = rb[b * 8];
#pragma unroll
                    for (int a = 0; a < MI; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[V2_STAGES + slot]));
            }
            double *ctile = p.c_base + mat * p.batch_stride + vb_tile_index(bi, bj, p.T) * VB_TILE_ELEMS;
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double *dst = ctile + (p0 + 8 * a + lr) * VB_TILE + q0 + 8 * b + 2 * lk;
                    if (p.accumulate) {
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(alpha * acc[a][b][0]) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 1), "d"(alpha * acc[a][b][1]) : "memory");
                    } else {
                        *reinterpret_cast<double2 *>(dst) = make_double2(alpha * acc[a][b][0], alpha * acc[a][b][1]);
                    }
                }
        }
    }
}

// ===========================================================================
// Fused in-shared-memory Gram (the literal north_star form): no Jacobian panel
// anywhere in HBM.  Same persistent consumer pipeline as v3, but the stages are
// PRODUCED by four generator warps that evaluate the Green's function
// (fast flavour: 21 FP64 slots per entry) straight into the padded operand
// rows; full[] barriers count generator-thread arrivals instead of TMA bytes.
// Kept as a measured alternative: DMMA and DFMA share one pipe on sm_100a, so
// regenerating 2*128 entries per 128x128 MACs per data row costs >= 33% extra
// FP64 issue on top of the matrix product.
// ===========================================================================
constexpr int FG_GEN_WARPS = 4;
constexpr int FG_THREADS = 8 * 32 + FG_GEN_WARPS * 32;
constexpr int FG_SMEM_BYTES = V2_RING_BYTES + VB_LOGTAB_SMEM_DOUBLES * (int)sizeof(double) + 64;

template <bool HAS_MINDIST>
__global__ void __launch_bounds__(FG_THREADS, 1)
vb_fused_gram_kernel(const VbFusedGramParams p, long long ntiles, const double2 *__restrict__ gtab)
{
    constexpr int NCW = 8, MI = 8;
    extern __shared__ __align__(128) double smem[];
    double *s_tab = smem + V2_STAGES * V2_STAGE_DOUBLES;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(s_tab + VB_LOGTAB_SMEM_DOUBLES);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nsteps = (int)((p.n + V2_KC - 1) / V2_KC);

    vb_logtab_stage(s_tab, gtab);
    if (tid == 0) {
        for (int s = 0; s < V2_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), FG_GEN_WARPS * 32);  // every generator thread arrives
            mbar_init(smem_u32(&bars[V2_STAGES + s]), NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long gstep = 0;
    if (warp >= NCW) {
        // ---------------- generator warps: thread g owns column g of both operand blocks ----------------
        const int g = tid - NCW * 32;
        const double2 *tab = reinterpret_cast<const double2 *>(s_tab);
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int bi, bj;
            decode_tile(t, p.T, 0, p.T, p.strip, bi, bj);
            const bool diag = bi == bj;
            const long long ca = (long long)bj * VB_TILE + g, cb = (long long)bi * VB_TILE + g;
            const bool va = ca < p.m, vb = cb < p.m;
            const double fxa = va ? p.fe[ca] : 0.0, fya = va ? p.fn[ca] : 0.0;
            const double fxb = vb ? p.fe[cb] : 0.0, fyb = vb ? p.fn[cb] : 0.0;
            for (int step = 0; step < nsteps; ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[V2_STAGES + slot]), ph ^ 1u);
                double *sA = smem + slot * V2_STAGE_DOUBLES;
                double *sB = sA + V2_KC * LDS_STRIDE;
                const long long r0 = (long long)step * V2_KC;
#pragma unroll 4
                for (int k = 0; k < V2_KC; ++k) {
                    const long long row = r0 + k;
                    const bool vr = row < p.n;
                    const double ex = vr ? p.east[row] : 0.0, ny = vr ? p.north[row] : 0.0;
                    const double sw = vr ? (p.sqrt_w ? p.sqrt_w[row] : 1.0) : 0.0;
                    double ga = vb_greens_fast<HAS_MINDIST>(ex - fxa, ny - fya, p.mindist, tab) * sw;
                    sA[k * LDS_STRIDE + g] = va ? ga : 0.0;
                    if (!diag) {
                        double gb = vb_greens_fast<HAS_MINDIST>(ex - fxb, ny - fyb, p.mindist, tab) * sw;
                        sB[k * LDS_STRIDE + g] = vb ? gb : 0.0;
                    }
                }
                mbar_arrive(smem_u32(&bars[slot]));
            }
        }
    } else {
        const int p0 = (warp >> 2) * (MI * 8);
        const int q0 = (warp & 3) * 32;
        const int lk = lane & 3, lr = lane >> 2;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int bi, bj;
            decode_tile(t, p.T, 0, p.T, p.strip, bi, bj);
            const bool diag = bi == bj;
            double acc[MI][4][2];
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            for (int step = 0; step < nsteps; ++step, ++gstep) {
                const int slot = (int)(gstep % V2_STAGES);
                const unsigned ph = (unsigned)(gstep / V2_STAGES) & 1u;
                mbar_wait(smem_u32(&bars[slot]), ph);
                const double *sA = smem + slot * V2_STAGE_DOUBLES;
                const double *sB = diag ? sA : sA + V2_KC * LDS_STRIDE;
#pragma unroll
                for (int k4 = 0; k4 < V2_KC / 4; ++k4) {
                    const double *ra = sA + (k4 * 4 + lk) * LDS_STRIDE + p0 + lr;
                    const double *rb = sB + (k4 * 4 + lk) * LDS_STRIDE + q0 + lr;
                    double fa[MI], fb[4];
#pragma unroll
                    for (int a = 0; a < MI; ++a) fa[a] = ra[a * 8];
#pragma unroll
                    for (int b = 0; b < 4; ++b) fb[b] = rb[b * 8];
#pragma unroll
                    for (int a = 0; a < MI; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[V2_STAGES + slot]));
            }
            double *ctile = p.c_base + vb_tile_index(bi, bj, p.T) * VB_TILE_ELEMS;
#pragma unroll
            for (int a = 0; a < MI; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double *dst = ctile + (p0 + 8 * a + lr) * VB_TILE + q0 + 8 * b + 2 * lk;
                    if (p.accumulate) {
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(acc[a][b][0]) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 1), "d"(acc[a][b][1]) : "memory");
                    } else {
                        *reinterpret_cast<double2 *>(dst) = make_double2(acc[a][b][0], acc[a][b][1]);
                    }
                }
        }
    }
}

}  // namespace

long long vb_launch_counter = 0;

cudaError_t vb_sm_count(int *sms)
{
    static int cached[256] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 256 && cached[dev] > 0) {
        *sms = cached[dev];
        return cudaSuccess;
    }
    if ((e = cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (dev >= 0 && dev < 256) cached[dev] = *sms;
    return cudaSuccess;
}
// 0: v1 (cp.async + __syncthreads); 1: v2 (TMA bulk + mbarrier + bulk reduce), 8 consumer warps;
// 2: v2 with 16 consumer warps (32x32 warp tiles); 3: v3 persistent (producer run-ahead, RED epilogue);
// 4: v3 with the dynamic tile queue (default: 35.72 vs 35.58 TFLOP/s over the config-2 fit, and what look-ahead needs)
int vb_gemm_variant = 4;
int vb_gemm_strip = 16;  // block columns per rasterisation strip (L2 reuse of the column panels)

// Tile-queue counters of the dynamic kernel: a per-device ring of zero-initialised slots; every launch takes the
// next slot and re-zeroes the one it will meet again 512 launches later (stream-ordered, so concurrent launches
// on different streams never share a slot).
constexpr int V4_SMEM_BYTES = V2_SMEM_BYTES + V4_RING * (int)sizeof(long long);
constexpr int V4_COUNTERS = 1024;
int vb_gemm_grid_limit = 0;  // > 0: cap the persistent grid (leave SMs to a concurrent stream); 0 = all SMs

cudaError_t vb_gemm_new_queue(unsigned long long **out, cudaStream_t stream)
{
    static unsigned long long *ring[256] = {nullptr};
    static unsigned next[256] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 256) return cudaErrorInvalidDevice;
    if (!ring[dev]) {
        if ((e = cudaMalloc(&ring[dev], sizeof(unsigned long long) * V4_COUNTERS)) != cudaSuccess) return e;
        if ((e = cudaMemset(ring[dev], 0, sizeof(unsigned long long) * V4_COUNTERS)) != cudaSuccess) return e;
    }
    const unsigned slot = next[dev]++ % V4_COUNTERS;
    *out = ring[dev] + slot;
    // this launch's slot must be zero when its kernel starts: zero it in stream order right before the launch
    return cudaMemsetAsync(*out, 0, sizeof(unsigned long long), stream);
}

long long vb_gemm_num_tiles(int T, int j_lo, int j_hi)
{
    // sum over block columns j_lo..j_hi-1 of the column heights (T - c): independent of the strip width
    long long total = 0;
    for (int c = j_lo; c < j_hi; ++c) total += T - c;
    return total;
}

long long vb_gemm_launch_tiles(VbGemmParams &p)
{
    if (p.nranges <= 0) return vb_gemm_num_tiles(p.T, p.j_lo, p.j_hi);
    long long total = 0;
    for (int r = 0; r < p.nranges; ++r) {
        p.range_first[r] = total;
        total += vb_gemm_num_tiles(p.T, p.range_lo[r], p.range_hi[r]);
    }
    return total;
}

cudaError_t vb_launch_gemm(const VbGemmParams &p_in, cudaStream_t stream)
{
    if (p_in.nranges > 0 && vb_gemm_variant != 4) {
        // only the dynamic kernel walks several ranges in one launch
        for (int r = 0; r < p_in.nranges; ++r) {
            VbGemmParams q = p_in;
            q.nranges = 0;
            q.j_lo = p_in.range_lo[r];
            q.j_hi = p_in.range_hi[r];
            cudaError_t er = vb_launch_gemm(q, stream);
            if (er != cudaSuccess) return er;
        }
        return cudaSuccess;
    }
    VbGemmParams p = p_in;
    p.strip = vb_gemm_strip < 1 ? 1 : vb_gemm_strip;
    if (p.nk <= 0 || p.nk > VB_MAX_KCHUNKS) return cudaErrorInvalidValue;
    const long long tiles = vb_gemm_launch_tiles(p);
    if (tiles <= 0) return cudaSuccess;
    cudaError_t e;
    if (p.batch > 1 && vb_gemm_variant != 3 && vb_gemm_variant != 4) {
        // only the persistent kernel walks a batch in one launch; the older variants go matrix by matrix
        for (int b = 0; b < p.batch; ++b) {
            VbGemmParams q = p_in;
            q.batch = 1;
            q.a_base += (long long)b * p.batch_stride;
            q.b_base += (long long)b * p.batch_stride;
            q.c_base += (long long)b * p.batch_stride;
            if ((e = vb_launch_gemm(q, stream)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (vb_gemm_variant == 1) {
        e = cudaFuncSetAttribute(vb_gemm_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        vb_gemm_tma_kernel<8><<<(unsigned)tiles, 8 * 32 + 32, V2_SMEM_BYTES, stream>>>(p);
    } else if (vb_gemm_variant == 3) {
        int sms = 0;
        if ((e = vb_sm_count(&sms)) != cudaSuccess) return e;
        e = cudaFuncSetAttribute(vb_gemm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        const long long all_tiles = tiles * (p.batch > 1 ? p.batch : 1);
        const unsigned grid = (unsigned)(all_tiles < sms ? all_tiles : sms);
        vb_gemm_persistent_kernel<<<grid, 8 * 32 + 32, V2_SMEM_BYTES, stream>>>(p, tiles);
    } else if (vb_gemm_variant == 4) {
        int sms = 0;
        if ((e = vb_sm_count(&sms)) != cudaSuccess) return e;
        e = cudaFuncSetAttribute(vb_gemm_dynamic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V4_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        unsigned long long *counter = p.queue;
        if (!counter && (e = vb_gemm_new_queue(&counter, stream)) != cudaSuccess) return e;
        const long long all_tiles = tiles * (p.batch > 1 ? p.batch : 1);
        const int limit = p.grid_limit > 0 ? p.grid_limit : vb_gemm_grid_limit;
        long long want = limit > 0 && limit < sms ? limit : sms;
        const unsigned grid = (unsigned)(all_tiles < want ? all_tiles : want);
        vb_gemm_dynamic_kernel<<<grid, 8 * 32 + 32, V4_SMEM_BYTES, stream>>>(p, tiles, counter);
    } else if (vb_gemm_variant == 2) {
        e = cudaFuncSetAttribute(vb_gemm_tma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        vb_gemm_tma_kernel<16><<<(unsigned)tiles, 16 * 32 + 32, V2_SMEM_BYTES, stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(vb_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        vb_gemm_kernel<<<(unsigned)tiles, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(p);
    }
    ++vb_launch_counter;
    return cudaGetLastError();
}

// defined in vb_greens.cu: device address of the log table used by the fast Green's function
const double2 *vb_logtab_device_ptr();

cudaError_t vb_launch_fused_gram(const VbFusedGramParams &p_in, cudaStream_t stream)
{
    VbFusedGramParams p = p_in;
    p.strip = vb_gemm_strip < 1 ? 1 : vb_gemm_strip;
    cudaError_t e = vb_init_tables();
    if (e != cudaSuccess) return e;
    const long long tiles = vb_gemm_num_tiles(p.T, 0, p.T);
    if (tiles <= 0 || p.n <= 0) return cudaSuccess;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    const double2 *gtab = vb_logtab_device_ptr();
    if (p.mindist != 0.0) {
        e = cudaFuncSetAttribute(vb_fused_gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FG_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        vb_fused_gram_kernel<true><<<grid, FG_THREADS, FG_SMEM_BYTES, stream>>>(p, tiles, gtab);
    } else {
        e = cudaFuncSetAttribute(vb_fused_gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FG_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        vb_fused_gram_kernel<false><<<grid, FG_THREADS, FG_SMEM_BYTES, stream>>>(p, tiles, gtab);
    }
    ++vb_launch_counter;
    return cudaGetLastError();
}
