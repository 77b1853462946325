// tools/fp64_peak.cu -- FP64 issue-rate microbenchmarks for B200 (sm_100a).
//
// MEASURED_PEAKS.json has no FP64 entry, so the roofline denominator for the
// Gram / Cholesky / predict kernels is measured here: DFMA (FP64 pipe), DMMA
// (mma.sync f64 shapes), both mixed in one warp, and the libdevice log/sqrt
// costs that bound the Green's function evaluation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
//   ./tools/fp64_peak            (prints one JSON object)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

// ---------------- DFMA: 16 independent chains per thread ----------------
__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- DMMA m8n8k4: NACC independent accumulators per warp ----------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double *out, double a, double b)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, double a0, double b0)
{
    double c[NACC][4];
    double a[4] = {a0, a0 + 1, a0 + 2, a0 + 3};
    double b[2] = {b0, b0 + 1};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, double a0, double b0)
{
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 0; c[i][3] = 1; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- mixed: per iteration 8 DMMA.884 + NF DFMA in the same warp ----------------
template <int NF>
__global__ void __launch_bounds__(256) k_mixed(double *out, double a, double b)
{
    double c[8][2];
    double acc[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int k = 0; k < NF / 8; ++k) {
                const int q = (i * (NF / 8) + k) % 16;
                acc[q] = fma(acc[q], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp-specialised mix: even warps DMMA only, odd warps DFMA only
__global__ void __launch_bounds__(256) k_mixed_ws(double *out, double a, double b)
{
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
    } else {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- libdevice costs: log, sqrt, rsqrt, div ----------------
template <int OP>
__global__ void __launch_bounds__(256) k_func(double *out, double x0)
{
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x0 + threadIdx.x * 1e-3 + i;
    double s = 0;
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v;
            if (OP == 0) v = log(x[i]);
            else if (OP == 1) v = sqrt(x[i]);
            else if (OP == 2) v = rsqrt(x[i]);
            else if (OP == 3) v = 1.0 / x[i];
            else v = __dsqrt_rn(x[i]) ;
            s += v;
            x[i] += 1.25;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int blocks = sms * 8, threads = 256;
    double *out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    const double nthreads = double(blocks) * threads, nwarps = nthreads / 32;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);

    double ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf(" \"dfma_tflops\": %.3f,\n", nthreads * ITERS * 16 * 2 / ms * 1e-9);

#define DM(NAME, KERN, NACC, MACS) \
    ms = time_ms([&] { KERN<NACC><<<blocks, threads>>>(out, 1.0000001, 1e-9); }); \
    printf(" \"" NAME "_acc%d_tflops\": %.3f,\n", NACC, nwarps * ITERS * NACC * MACS * 2.0 / ms * 1e-9);
    DM("dmma884", k_dmma884, 1, 256) DM("dmma884", k_dmma884, 4, 256) DM("dmma884", k_dmma884, 8, 256) DM("dmma884", k_dmma884, 16, 256)
    DM("dmma1688", k_dmma1688, 1, 1024) DM("dmma1688", k_dmma1688, 4, 1024) DM("dmma1688", k_dmma1688, 8, 1024)
    DM("dmma16816", k_dmma16816, 1, 2048) DM("dmma16816", k_dmma16816, 4, 2048) DM("dmma16816", k_dmma16816, 8, 2048)

#define MX(NF) \
    ms = time_ms([&] { k_mixed<NF><<<blocks, threads>>>(out, 1.0000001, 1e-9); }); \
    printf(" \"mixed_8dmma_%ddfma\": {\"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"ms\": %.4f},\n", NF, \
           nwarps * ITERS * 8 * 256 * 2.0 / ms * 1e-9, nthreads * ITERS * NF * 2.0 / ms * 1e-9, ms);
    MX(8) MX(16) MX(32) MX(64)
    ms = time_ms([&] { k_mixed_ws<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    printf(" \"mixed_warp_specialised\": {\"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f, \"ms\": %.4f},\n",
           nwarps / 2 * ITERS * 8 * 256 * 2.0 / ms * 1e-9, nthreads / 2 * ITERS * 16 * 2.0 / ms * 1e-9, ms);

#define FN(NAME, OP) \
    ms = time_ms([&] { k_func<OP><<<blocks, threads>>>(out, 3.7); }); \
    printf(" \"" NAME "_gops\": %.3f,\n", nthreads * 512 * 4 / ms * 1e-6);
    FN("log", 0) FN("sqrt", 1) FN("rsqrt", 2) FN("div", 3)
    printf(" \"done\": true}\n");
    return 0;
}
