// vb_api.cu -- the C ABI declared in include/verde_b200.h.
//
// Host orchestration only: staging of host buffers, the cached device
// workspace, the slab loop of the Gram accumulation, CUDA-event timing for the
// roofline report.  All arithmetic is in the kernels of vb_greens.cu,
// vb_gram.cu, vb_gemm.cu and vb_chol.cu.
#include "../../include/verde_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(expr)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(e_ == cudaErrorMemoryAllocation ? VB200_E_NOMEM : VB200_E_CUDA,          \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- options -------------------------------------------------------------------------
struct Options {
    int predict_exact = 0;
    int chol_panel_tiles = 8;
    int gram_slab_chunks = 16;
    int gram_fused = 0;
    int sweep_group = 0;  // damped systems factored side by side in vb200_gram_solve_multi (0 = as many as fit)
    int fit_diagnostics = 1;  // condition estimate + backward error after every fit (vb_diag.cu)
    int svd_max_sweeps = 80;  // sweep limit of the damping=None Jacobi SVD (vb_svd.cu)
    int cond_inverse_its = 0;  // inverse iterations of the condition estimate (0 = 10 up to 8192 unknowns, else 3)
    int chol_lookahead = 24;  // SMs on which panel p+1 is factored beside the trailing update of panel p (0 = off)
} g_opt;

// ---- cached workspace (grow-only, per process; one process per GPU) --------------------
struct Slot {
    void *ptr = nullptr;
    size_t bytes = 0;
};
std::map<std::string, Slot> g_ws;

cudaError_t ws_get(const char *name, size_t bytes, void **out)
{
    // slots are per device: a process that switches devices must not see another device's pointers
    int dev = 0;
    cudaError_t de = cudaGetDevice(&dev);
    if (de != cudaSuccess) return de;
    Slot &s = g_ws[std::to_string(dev) + ":" + name];
    if (s.bytes < bytes) {
        if (s.ptr) cudaFree(s.ptr);
        s.ptr = nullptr;
        s.bytes = 0;
        cudaError_t e = cudaMalloc(&s.ptr, bytes ? bytes : 8);
        if (e != cudaSuccess) return e;
        s.bytes = bytes ? bytes : 8;
    }
    *out = s.ptr;
    return cudaSuccess;
}

bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input staging: host pointers are copied into a named workspace slot.
cudaError_t stage_in(const char *slot, const double *p, long long count, const double **dev,
                     cudaStream_t st)
{
    if (p == nullptr || count == 0) {
        *dev = nullptr;
        return cudaSuccess;
    }
    if (is_device_ptr(p)) {
        *dev = p;
        return cudaSuccess;
    }
    void *d;
    cudaError_t e = ws_get(slot, sizeof(double) * count, &d);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d, p, sizeof(double) * count, cudaMemcpyHostToDevice, st);
    *dev = static_cast<const double *>(d);
    return e;
}

struct OutBuf {
    double *user = nullptr;
    double *dev = nullptr;
    long long count = 0;
    bool staged = false;
};

cudaError_t stage_out(const char *slot, double *p, long long count, OutBuf *ob)
{
    ob->user = p;
    ob->count = count;
    if (count == 0) return cudaSuccess;
    if (is_device_ptr(p)) {
        ob->dev = p;
        ob->staged = false;
        return cudaSuccess;
    }
    void *d;
    cudaError_t e = ws_get(slot, sizeof(double) * count, &d);
    ob->dev = static_cast<double *>(d);
    ob->staged = true;
    return e;
}

cudaError_t finish_out(const OutBuf &ob, cudaStream_t st)
{
    if (ob.staged && ob.count)
        return cudaMemcpyAsync(ob.user, ob.dev, sizeof(double) * ob.count, cudaMemcpyDeviceToHost, st);
    return cudaSuccess;
}

// ---- event timing ------------------------------------------------------------------------
struct Span {
    cudaEvent_t a, b;
};
struct Profile {
    std::vector<Span> gemm;
    double gemm_flops = 0.0;
    long long launches = 0;
    // launches that share the GPU with a concurrently running panel factorisation (look-ahead, untimed == 2):
    // their spans measure a kernel that owns only part of the machine, so they are kept apart from the roofline spans
    std::vector<Span> shared;
    double shared_flops = 0.0;
    long long shared_launches = 0;
    bool on = false;
} g_prof;

cudaError_t span_begin(Span *s, cudaStream_t st)
{
    cudaError_t e = cudaEventCreate(&s->a);
    if (e != cudaSuccess) return e;
    e = cudaEventCreate(&s->b);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(s->a, st);
}
double span_ms(Span &s)
{
    float ms = 0.f;
    cudaEventSynchronize(s.b);
    cudaEventElapsedTime(&ms, s.a, s.b);
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
    return ms;
}

cudaError_t timed_gemm(const VbGemmParams &p, cudaStream_t st)
{
    if (p.untimed == 1) return vb_launch_gemm(p, st);
    Span s;
    cudaError_t e = span_begin(&s, st);
    if (e != cudaSuccess) return e;
    e = vb_launch_gemm(p, st);
    cudaEventRecord(s.b, st);
    VbGemmParams counted = p;
    const double flops = 2.0 * VB_TILE * VB_TILE * VB_TILE * (double)p.nk * (double)vb_gemm_launch_tiles(counted) *
                         (double)(p.batch > 1 ? p.batch : 1);
    if (p.untimed == 2) {
        g_prof.shared.push_back(s);
        g_prof.shared_flops += flops;
        g_prof.shared_launches += 1;
    } else {
        g_prof.gemm.push_back(s);
        g_prof.gemm_flops += flops;
        g_prof.launches += 1;
    }
    return e;
}

void prof_reset()
{
    g_prof.gemm.clear();
    g_prof.gemm_flops = 0.0;
    g_prof.launches = 0;
    g_prof.shared.clear();
    g_prof.shared_flops = 0.0;
    g_prof.shared_launches = 0;
}

void prof_collect(vb200_fit_info *info)
{
    double ms = 0.0;
    for (auto &s : g_prof.gemm) ms += span_ms(s);
    double shared_ms = 0.0;
    for (auto &s : g_prof.shared) shared_ms += span_ms(s);
    if (info) {
        info->gemm_ms += ms;
        info->gemm_flops += g_prof.gemm_flops;
        info->gemm_launches += g_prof.launches;
        info->shared_gemm_ms += shared_ms;
        info->shared_gemm_flops += g_prof.shared_flops;
        info->shared_gemm_launches += g_prof.shared_launches;
    }
    prof_reset();
}

__global__ void vb_unscale_kernel(const double *__restrict__ x, const double *__restrict__ inv_scale,
                                  long long cols, double *__restrict__ out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cols) out[j] = x[j] * inv_scale[j];
}

// batched flavours for the damping sweep: grid.y = system
__global__ void vb_replicate_kernel(const double *__restrict__ src, long long count, double *__restrict__ dst)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) dst[(long long)blockIdx.y * count + j] = src[j];
}
__global__ void vb_unscale_batched_kernel(const double *__restrict__ x, long long xstride,
                                          const double *__restrict__ inv_scale, long long cols,
                                          double *__restrict__ out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cols) out[(long long)blockIdx.y * cols + j] = x[(long long)blockIdx.y * xstride + j] * inv_scale[j];
}

// min / max of diag(L)^2 over the first `cols` unknowns (single CTA; T*128 values)
__global__ void vb_diag_minmax_kernel(const double *__restrict__ L, int T, long long cols,
                                      double *__restrict__ out2)
{
    __shared__ double smin[256], smax[256];
    double lo = 1e300, hi = 0.0;
    for (long long j = threadIdx.x; j < cols; j += blockDim.x) {
        const int b = (int)(j / VB_TILE), c = (int)(j % VB_TILE);
        const double d = L[vb_tile_index(b, b, T) * VB_TILE_ELEMS + c * VB_TILE + c];
        const double v = d * d;
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    smin[threadIdx.x] = lo;
    smax[threadIdx.x] = hi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + s]);
            smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out2[0] = smin[0];
        out2[1] = smax[0];
    }
}

// ---- FP64 issue-rate microbenchmark (roofline denominator) -----------------------------
constexpr int PEAK_ITERS = 2048;
__global__ void __launch_bounds__(256) vb_peak_dmma_kernel(double *out, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i][0] = threadIdx.x;
        c[i][1] = i;
    }
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) vb_peak_dfma_kernel(double *out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ---- per-fit diagnostics (kernels in vb_diag.cu) -------------------------------------------------------
double wall_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

// condition estimate of A = L L^T (every solve path leaves L in `L`); fills the cond fields of *info
int diag_condition(const double *L, int T, long long cols, double damping, vb200_fit_info *info, cudaStream_t st,
                   const double *solution = nullptr)
{
    if (!info || !g_opt.fit_diagnostics) return 0;
    const double t0 = wall_ms();
    double *scratch;
    int *d_flags;
    CU(ws_get("diag_scratch", sizeof(double) * (size_t)vb_diag_scratch_doubles(T), (void **)&scratch));
    CU(ws_get("sweep_flags", sizeof(int) * 2 * T, (void **)&d_flags));
    VbCondEstimate ce;
    const long long launches0 = vb_launch_counter;
    // inverse iterations: each is a pair of triangular sweeps over the whole factor.  Seeded with the solution the
    // estimate is within 2x of the true lambda_min after two steps and within 1.4x after three (measured on the
    // 3k/5k systems against eigvalsh); small systems, where a step costs microseconds, run to 1 % stagnation.
    const int inverse_its = g_opt.cond_inverse_its > 0 ? g_opt.cond_inverse_its : (T <= 64 ? 10 : 3);
    CU(vb_cond_estimate(L, T, cols, scratch, d_flags, st, 30, inverse_its, solution, &ce));
    info->power_its = ce.power_its;
    info->inverse_its = ce.inverse_its;
    info->lambda_max = ce.lambda_max;
    info->lambda_min_est = ce.lambda_min;
    info->cond_upper = damping > 0.0 ? ce.lambda_max / damping : 0.0;
    info->cond_est = ce.lambda_min > 0.0 ? ce.lambda_max / ce.lambda_min : 0.0;
    // When damping < eps * lambda_max the factor's own rounding (||A - L L^T|| ~ eps lambda_max) can push the smallest
    // eigenvalue of L L^T below the damping; cond_2(A) itself is bounded by lambda_max / damping, so the estimate is
    // clamped there (lambda_min_est keeps the raw Rayleigh quotient).
    if (info->cond_upper > 0.0 && (info->cond_est > info->cond_upper || info->cond_est == 0.0)) info->cond_est = info->cond_upper;
    info->kernel_launches += vb_launch_counter - launches0;
    info->diag_ms += wall_ms() - t0;
    return 0;
}

// grad = J^T W (d - J force) for the rows of `pb` (all pointers on the device), matrix-free, exact Green's functions
int normal_residual_dev(const VbGramProblem &pb, const double *force, double *grad, cudaStream_t st)
{
    const long long rows = vb_sys_rows(pb), cols = vb_sys_cols(pb);
    double *pred, *resid;
    CU(ws_get("diag_pred", sizeof(double) * (size_t)(rows > 0 ? rows : 1), (void **)&pred));
    CU(ws_get("diag_resid", sizeof(double) * (size_t)(rows > 0 ? rows : 1), (void **)&resid));
    if (rows == 0) {
        CU(cudaMemsetAsync(grad, 0, sizeof(double) * cols, st));
        return 0;
    }
    int pts, nsplit;
    double *scratch = nullptr;
    if (pb.kind == 2) {
        CU(vb_launch_gemv(pb.jac, rows, cols, force, 0, pred, st));
        CU(vb_launch_weighted_residual(pb.data, pred, pb.weights, rows, resid, st));
        CU(vb_launch_gemv(pb.jac, rows, cols, resid, 1, grad, st));
    } else if (pb.kind == 0) {
        vb_predict_plan(pb.n, pb.m, &pts, &nsplit);
        if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * (size_t)nsplit * pb.n, (void **)&scratch));
        CU(vb_launch_predict(pb.east, pb.north, pb.n, pb.fe, pb.fn, force, pb.m, pb.mindist, 1, pts, nsplit, scratch,
                             pred, st));
        CU(vb_launch_weighted_residual(pb.data, pred, pb.weights, rows, resid, st));
        // J^T r: the same kernel with points and forces swapped (g is even in the coordinate differences)
        vb_predict_plan(pb.m, pb.n, &pts, &nsplit);
        scratch = nullptr;
        if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * (size_t)nsplit * pb.m, (void **)&scratch));
        CU(vb_launch_predict(pb.fe, pb.fn, pb.m, pb.east, pb.north, resid, pb.n, pb.mindist, 1, pts, nsplit, scratch,
                             grad, st));
    } else {
        vb_predict_plan(pb.n, pb.m, &pts, &nsplit);
        if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * 2 * (size_t)nsplit * pb.n, (void **)&scratch));
        CU(vb_launch_predict2d(pb.east, pb.north, pb.n, pb.fe, pb.fn, force, pb.m, pb.mindist, pb.poisson, 1, pts, nsplit,
                               scratch, pred, pred + pb.n, st));
        CU(vb_launch_weighted_residual(pb.data, pred, pb.weights, rows, resid, st));
        // the block Jacobian [[ee, ne], [ne, nn]] is symmetric under the swap of points and forces as well
        vb_predict_plan(pb.m, pb.n, &pts, &nsplit);
        scratch = nullptr;
        if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * 2 * (size_t)nsplit * pb.m, (void **)&scratch));
        CU(vb_launch_predict2d(pb.fe, pb.fn, pb.m, pb.east, pb.north, resid, pb.n, pb.mindist, pb.poisson, 1, pts, nsplit,
                               scratch, grad, grad + pb.m, st));
    }
    return 0;
}

// eta from the summed gradient (all device pointers); synchronises the stream
int backward_error_dev(const double *grad, const double *jtwd, const double *inv_scale, const double *force,
                       long long cols, double damping, double lambda_max, double *eta, cudaStream_t st)
{
    double *d3, h3[3] = {0, 0, 0};
    CU(ws_get("diag_norms", sizeof(double) * 3, (void **)&d3));
    CU(vb_launch_backward_norms(grad, jtwd, inv_scale, force, cols, damping, d3, st));
    CU(cudaMemcpyAsync(h3, d3, sizeof(h3), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double den = lambda_max * sqrt(h3[1]) + sqrt(h3[2]);
    *eta = den > 0.0 ? sqrt(h3[0]) / den : 0.0;
    return 0;
}

VbGramProblem g_last_problem = {};  // the inputs of the last vb200_gram_accumulate, as device pointers

// stages (host pointers) or passes through (device pointers) the inputs of a fit into *pb
int stage_problem(int kind, const double *jac, const double *east, const double *north, const double *data,
                  const double *weights, int64_t n, const double *force_east, const double *force_north, int64_t m,
                  double mindist, double poisson, VbGramProblem *pb, cudaStream_t st)
{
    *pb = VbGramProblem{};
    pb->kind = kind;
    pb->n = n;
    pb->m = m;
    pb->mindist = mindist;
    pb->poisson = poisson;
    const long long rows = vb_sys_rows(*pb), cols = vb_sys_cols(*pb);
    if (kind == 2) {
        CU(stage_in("in_jac", jac, rows * cols, &pb->jac, st));
    } else {
        CU(stage_in("in_e", east, n, &pb->east, st));
        CU(stage_in("in_n", north, n, &pb->north, st));
        CU(stage_in("in_fe", force_east, m, &pb->fe, st));
        CU(stage_in("in_fn", force_north, m, &pb->fn, st));
    }
    CU(stage_in("in_d", data, rows, &pb->data, st));
    CU(stage_in("in_w", weights, rows, &pb->weights, st));
    return 0;
}
}  // namespace

// vb_chol.cu calls the GEMM through this hook so that its launches are timed too.
cudaError_t vb_gemm_dispatch(const VbGemmParams &p, cudaStream_t st) { return timed_gemm(p, st); }

extern "C" {

int vb200_version(void) { return 201; }
const char *vb200_last_error(void) { return g_error.c_str(); }

int vb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vb200_set_device(int device)
{
    CU(cudaSetDevice(device));
    return 0;
}

int vb200_set_option(const char *key, double value)
{
    if (!key) return fail(VB200_E_BADARG, "null option key");
    const std::string k(key);
    if (k == "predict_exact") g_opt.predict_exact = value != 0.0;
    else if (k == "chol_panel_tiles") g_opt.chol_panel_tiles = value < 1 ? 1 : (value > VB_MAX_KCHUNKS ? VB_MAX_KCHUNKS : (int)value);
    else if (k == "gram_slab_chunks") g_opt.gram_slab_chunks = value < 1 ? 1 : (value > VB_MAX_KCHUNKS ? VB_MAX_KCHUNKS : (int)value);
    else if (k == "solve_variant") vb_solve_variant = (int)value;
    else if (k == "trsm_tiles") vb_trsm_tiles = (int)value;
    else if (k == "gram_fused") g_opt.gram_fused = value != 0.0;
    else if (k == "sweep_group") g_opt.sweep_group = value < 0 ? 0 : (int)value;
    else if (k == "fit_diagnostics") g_opt.fit_diagnostics = value != 0.0;
    else if (k == "svd_max_sweeps") g_opt.svd_max_sweeps = value < 1 ? 1 : (int)value;
    else if (k == "chol_lookahead") g_opt.chol_lookahead = value < 0 ? 0 : (int)value;
    else if (k == "cond_inverse_its") g_opt.cond_inverse_its = value < 0 ? 0 : (int)value;
    else if (k == "chol_lookahead_min_rows") vb_chol_lookahead_min_rows = value < 1 ? 1 : (int)value;
    else if (k == "gemm_variant") vb_gemm_variant = (int)value;
    else if (k == "gemm_strip") vb_gemm_strip = (int)value;
    else if (k == "gemm_grid_limit") vb_gemm_grid_limit = value < 0 ? 0 : (int)value;
    else if (k == "predict_minb") vb_predict_minb = (int)value;
    else if (k == "predict_pts") vb_predict_pts = (int)value;
    else if (k == "predict_wide") vb_predict_wide = value != 0.0;
    else return fail(VB200_E_BADARG, "unknown option '%s'", key);
    return 0;
}

int vb200_get_option(const char *key, double *value)
{
    if (!key || !value) return fail(VB200_E_BADARG, "null pointer");
    const std::string k(key);
    if (k == "predict_exact") *value = g_opt.predict_exact;
    else if (k == "chol_panel_tiles") *value = g_opt.chol_panel_tiles;
    else if (k == "gram_slab_chunks") *value = g_opt.gram_slab_chunks;
    else if (k == "solve_variant") *value = vb_solve_variant;
    else if (k == "trsm_tiles") *value = vb_trsm_tiles;
    else if (k == "gram_fused") *value = g_opt.gram_fused;
    else if (k == "sweep_group") *value = g_opt.sweep_group;
    else if (k == "fit_diagnostics") *value = g_opt.fit_diagnostics;
    else if (k == "svd_max_sweeps") *value = g_opt.svd_max_sweeps;
    else if (k == "gemm_variant") *value = vb_gemm_variant;
    else if (k == "gemm_strip") *value = vb_gemm_strip;
    else if (k == "gemm_grid_limit") *value = vb_gemm_grid_limit;
    else if (k == "predict_minb") *value = vb_predict_minb;
    else if (k == "predict_pts") *value = vb_predict_pts;
    else if (k == "predict_wide") *value = vb_predict_wide;
    else if (k == "chol_lookahead") *value = g_opt.chol_lookahead;
    else if (k == "cond_inverse_its") *value = g_opt.cond_inverse_its;
    else if (k == "chol_lookahead_min_rows") *value = vb_chol_lookahead_min_rows;
    else return fail(VB200_E_BADARG, "unknown option '%s'", key);
    return 0;
}

int vb200_release_workspace(void)
{
    for (auto &kv : g_ws)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    g_ws.clear();
    return 0;
}

int64_t vb200_launch_count(void) { return vb_launch_counter; }

double vb200_measure_fp64_peak_tflops(int kind)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1.0;
    const int blocks = sms * 8, threads = 256;
    double *out = nullptr;
    if (ws_get("peak", sizeof(double) * blocks * threads, (void **)&out) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, 0);
        if (kind == 0) vb_peak_dmma_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
        else vb_peak_dfma_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    const double threads_total = (double)blocks * threads;
    const double flops = kind == 0 ? threads_total / 32 * PEAK_ITERS * 8 * 256 * 2.0
                                   : threads_total * PEAK_ITERS * 16 * 2.0;
    return flops / best * 1e-9;
}

int vb200_jacobian(const double *east, const double *north, int64_t n, const double *force_east,
                   const double *force_north, int64_t m, double mindist, double *jac)
{
    if (n < 0 || m < 0) return fail(VB200_E_BADARG, "negative size");
    if (n == 0 || m == 0) return 0;
    if (!east || !north || !force_east || !force_north || !jac) return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    OutBuf ob;
    CU(stage_out("out_jac", jac, (long long)n * m, &ob));
    CU(vb_launch_jacobian(de, dn, n, dfe, dfn, m, mindist, ob.dev, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_jacobian_2d(const double *east, const double *north, int64_t n,
                      const double *force_east, const double *force_north, int64_t m,
                      double mindist, double poisson, double *jac)
{
    if (n < 0 || m < 0) return fail(VB200_E_BADARG, "negative size");
    if (n == 0 || m == 0) return 0;
    if (!east || !north || !force_east || !force_north || !jac) return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    OutBuf ob;
    CU(stage_out("out_jac", jac, 4LL * n * m, &ob));
    CU(vb_launch_jacobian2d(de, dn, n, dfe, dfn, m, mindist, poisson, ob.dev, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_predict(const double *east, const double *north, int64_t n, const double *force_east,
                  const double *force_north, const double *forces, int64_t m, double mindist,
                  double *result)
{
    if (n < 0 || m < 0) return fail(VB200_E_BADARG, "negative size");
    if (n == 0) return 0;
    if (!east || !north || !result || (m > 0 && (!force_east || !force_north || !forces)))
        return fail(VB200_E_BADARG, "null pointer");
    if (m > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many forces");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn, *df;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    CU(stage_in("in_f", forces, m, &df, st));
    OutBuf ob;
    CU(stage_out("out_pred", result, n, &ob));
    int pts, nsplit;
    vb_predict_plan(n, m, &pts, &nsplit);
    double *scratch = nullptr;
    if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * nsplit * n, (void **)&scratch));
    CU(vb_launch_predict(de, dn, n, dfe, dfn, df, m, mindist, g_opt.predict_exact, pts, nsplit, scratch, ob.dev, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_predict_multi(const double *east, const double *north, int64_t n, const double *force_east,
                        const double *force_north, const double *forces, int64_t m, int32_t n_components,
                        double mindist, double *result)
{
    if (n < 0 || m < 0 || n_components < 1) return fail(VB200_E_BADARG, "bad sizes");
    if (n == 0) return 0;
    if (!east || !north || !result || (m > 0 && (!force_east || !force_north || !forces)))
        return fail(VB200_E_BADARG, "null pointer");
    if (m > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many forces");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn, *df;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    CU(stage_in("in_f_multi", forces, m * n_components, &df, st));
    OutBuf ob;
    CU(stage_out("out_pred_multi", result, n * n_components, &ob));
    int pts, nsplit;
    vb_predict_plan(n, m, &pts, &nsplit);
    // components go through the kernel three (or two) at a time; a single leftover uses the plain kernel
    for (int q0 = 0; q0 < n_components;) {
        const int left = n_components - q0;
        const int k = left >= 3 && left != 4 ? 3 : (left >= 2 ? 2 : 1);
        double *scratch = nullptr;
        if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * (size_t)k * nsplit * n, (void **)&scratch));
        if (k == 1)
            CU(vb_launch_predict(de, dn, n, dfe, dfn, df + (long long)q0 * m, m, mindist, g_opt.predict_exact, pts, nsplit,
                                 scratch, ob.dev + (long long)q0 * n, st));
        else
            CU(vb_launch_predict_multi(de, dn, n, dfe, dfn, df + (long long)q0 * m, m, k, mindist, g_opt.predict_exact, nsplit,
                                       scratch, ob.dev + (long long)q0 * n, st));
        q0 += k;
    }
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_predict_2d(const double *east, const double *north, int64_t n,
                     const double *force_east, const double *force_north, const double *forces,
                     int64_t m, double mindist, double poisson, double *vec_east,
                     double *vec_north)
{
    if (n < 0 || m < 0) return fail(VB200_E_BADARG, "negative size");
    if (n == 0) return 0;
    if (!east || !north || !vec_east || !vec_north || (m > 0 && (!force_east || !force_north || !forces)))
        return fail(VB200_E_BADARG, "null pointer");
    if (m > 0x3fffffffLL) return fail(VB200_E_BADARG, "too many forces");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn, *df;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    CU(stage_in("in_f", forces, 2 * m, &df, st));
    OutBuf oe, on;
    CU(stage_out("out_pred", vec_east, n, &oe));
    CU(stage_out("out_pred2", vec_north, n, &on));
    int pts, nsplit;
    vb_predict_plan(n, m, &pts, &nsplit);
    double *scratch = nullptr;
    if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * 2 * nsplit * n, (void **)&scratch));
    CU(vb_launch_predict2d(de, dn, n, dfe, dfn, df, m, mindist, poisson, g_opt.predict_exact, pts, nsplit,
                           scratch, oe.dev, on.dev, st));
    CU(finish_out(oe, st));
    CU(finish_out(on, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int64_t vb200_gram_doubles(int64_t cols)
{
    const int64_t T = (cols + VB_TILE - 1) / VB_TILE;
    return T * (T + 1) / 2 * VB_TILE_ELEMS;
}

int64_t vb200_stats_doubles(int64_t cols)
{
    const int64_t T = (cols + VB_TILE - 1) / VB_TILE;
    return 4 * T * VB_TILE;
}

static int gram_accumulate_impl(int kind, const double *jac, const double *east, const double *north,
                                const double *data, const double *weights, int64_t n,
                                const double *force_east, const double *force_north, int64_t m,
                                double mindist, double poisson, double *gram_dev, double *stats_dev,
                                int accumulate, vb200_fit_info *info)
{
    if (n < 0 || m <= 0) return fail(VB200_E_BADARG, "bad sizes n=%lld m=%lld", (long long)n, (long long)m);
    if (!gram_dev || !stats_dev) return fail(VB200_E_BADARG, "null gram/stats buffer");
    if (!is_device_ptr(gram_dev) || !is_device_ptr(stats_dev))
        return fail(VB200_E_BADARG, "gram/stats buffers must be device memory");
    if (n > 0 && (!data || (kind != 2 && (!east || !north)))) return fail(VB200_E_BADARG, "null data pointer");
    if (kind != 2 && (!force_east || !force_north)) return fail(VB200_E_BADARG, "null force coordinates");
    cudaStream_t st = 0;
    VbGramProblem pb;
    {
        const int rc = stage_problem(kind, jac, east, north, data, weights, n, force_east, force_north, m, mindist,
                                     poisson, &pb, st);
        if (rc != 0) return rc;
    }
    g_last_problem = pb;
    const long long rows = vb_sys_rows(pb), cols = vb_sys_cols(pb);
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;

    if (info) {
        info->block_rows = T;
        info->rows = rows;
        info->cols = cols;
    }
    Span whole;
    CU(span_begin(&whole, st));
    prof_reset();
    if (!accumulate) CU(cudaMemsetAsync(stats_dev, 0, sizeof(double) * 4 * Mpad, st));
    const long long chunks = (rows + VB_TILE - 1) / VB_TILE;
    if (chunks == 0 && !accumulate)
        CU(cudaMemsetAsync(gram_dev, 0, sizeof(double) * vb200_gram_doubles(cols), st));
    const int slab = g_opt.gram_slab_chunks;
    const bool fused = g_opt.gram_fused && kind == 0 && chunks > 0;
    double *panel = nullptr;
    if (chunks > 0 && !fused)
        CU(ws_get("panel", sizeof(double) * (size_t)(chunks < slab ? chunks : slab) * T * VB_TILE_ELEMS, (void **)&panel));
    const long long launches0 = vb_launch_counter;
    if (fused) {
        // measured alternative: statistics in a generation-only pass, then ONE
        // persistent kernel that generates Jacobian rows in shared memory and feeds the DMMA tiles
        for (long long c0 = 0; c0 < chunks; c0 += slab) {
            const int nch = (int)((chunks - c0 < slab) ? (chunks - c0) : slab);
            CU(vb_launch_gram_panel(pb, c0 * VB_TILE, nch, T, nullptr, stats_dev, st));
        }
        VbFusedGramParams f = {};
        f.east = pb.east;
        f.north = pb.north;
        f.n = pb.n;
        f.fe = pb.fe;
        f.fn = pb.fn;
        f.m = pb.m;
        f.mindist = pb.mindist;
        f.c_base = gram_dev;
        f.T = T;
        f.accumulate = accumulate ? 1 : 0;
        if (pb.weights) {
            double *sw;
            CU(ws_get("sqrt_w", sizeof(double) * rows, (void **)&sw));
            CU(vb_launch_sqrt(pb.weights, rows, sw, st));
            f.sqrt_w = sw;
        }
        Span sp;
        CU(span_begin(&sp, st));
        CU(vb_launch_fused_gram(f, st));
        CU(cudaEventRecord(sp.b, st));
        g_prof.gemm.push_back(sp);
        g_prof.gemm_flops += 2.0 * VB_TILE * VB_TILE * (double)rows * (double)vb_gemm_num_tiles(T, 0, T);
        g_prof.launches += 1;
    }
    for (long long c0 = 0; c0 < chunks && !fused; c0 += slab) {
        const int nch = (int)((chunks - c0 < slab) ? (chunks - c0) : slab);
        CU(vb_launch_gram_panel(pb, c0 * VB_TILE, nch, T, panel, stats_dev, st));
        VbGemmParams g = {};
        g.a_base = g.b_base = panel;
        for (int kk = 0; kk < nch; ++kk) g.a_chunk_off[kk] = g.b_chunk_off[kk] = (long long)kk * T * VB_TILE_ELEMS;
        g.nk = nch;
        g.c_base = gram_dev;
        g.T = T;
        g.j_lo = 0;
        g.j_hi = T;
        g.alpha = 1.0;
        g.accumulate = (accumulate || c0 > 0) ? 1 : 0;
        CU(timed_gemm(g, st));
    }
    CU(cudaEventRecord(whole.b, st));
    CU(cudaStreamSynchronize(st));
    if (info) {
        info->gram_ms += span_ms(whole);
        info->kernel_launches += vb_launch_counter - launches0;
    } else {
        span_ms(whole);
    }
    prof_collect(info);
    return 0;
}

int vb200_gram_accumulate(int kind, const double *east, const double *north, const double *data,
                          const double *weights, int64_t n, const double *force_east,
                          const double *force_north, int64_t m, double mindist, double poisson,
                          double *gram_dev, double *stats_dev, int accumulate,
                          vb200_fit_info *info)
{
    if (kind != 0 && kind != 1) return fail(VB200_E_BADARG, "kind must be 0 (scalar) or 1 (elastic 2-D)");
    return gram_accumulate_impl(kind, nullptr, east, north, data, weights, n, force_east, force_north, m,
                                mindist, poisson, gram_dev, stats_dev, accumulate, info);
}

int vb200_gram_solve(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                     double damping, int keep_gram, double *out_params, vb200_fit_info *info)
{
    if (cols <= 0 || total_rows <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (!gram_dev || !stats_dev || !out_params) return fail(VB200_E_BADARG, "null pointer");
    if (!(damping > 0.0)) return fail(VB200_E_BADARG, "damping must be > 0 (the damping=None branch of the reference is out of scope)");
    if (!is_device_ptr(gram_dev) || !is_device_ptr(stats_dev))
        return fail(VB200_E_BADARG, "gram/stats buffers must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    double *inv_scale, *rhs, *mm;
    int *d_info;
    CU(ws_get("inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(ws_get("minmax", sizeof(double) * 2, (void **)&mm));
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    double *L = gram_dev;
    if (keep_gram) {
        CU(ws_get("chol", sizeof(double) * vb200_gram_doubles(cols), (void **)&L));
        CU(cudaMemcpyAsync(L, gram_dev, sizeof(double) * vb200_gram_doubles(cols), cudaMemcpyDeviceToDevice, st));
    }
    OutBuf ob;
    CU(stage_out("out_params", out_params, cols, &ob));
    if (info) {
        info->block_rows = T;
        info->cols = cols;
        if (info->rows == 0) info->rows = total_rows;
    }
    prof_reset();
    const long long launches0 = vb_launch_counter;
    Span fac, sol;
    CU(span_begin(&fac, st));
    CU(vb_launch_column_scale(stats_dev, T, cols, total_rows, inv_scale, rhs, st));
    CU(vb_launch_scale_system(L, T, cols, inv_scale, damping, st));
    CU(vb_cholesky_tiled(L, T, g_opt.chol_panel_tiles, d_info, st, 1, 0, g_opt.chol_lookahead));
    CU(cudaEventRecord(fac.b, st));
    CU(span_begin(&sol, st));
    int *d_flags;
    CU(ws_get("sweep_flags", sizeof(int) * 2 * T, (void **)&d_flags));
    CU(vb_solve_tiled(L, T, rhs, d_flags, st));
    vb_unscale_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(rhs, inv_scale, cols, ob.dev);
    vb_diag_minmax_kernel<<<1, 256, 0, st>>>(L, T, cols, mm);
    vb_launch_counter += 2;
    CU(cudaGetLastError());
    CU(cudaEventRecord(sol.b, st));
    CU(finish_out(ob, st));
    int h_info = 0;
    double h_mm[2] = {0, 0};
    CU(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double fms = span_ms(fac), sms = span_ms(sol);
    if (info) {
        info->factor_ms += fms;
        info->solve_ms += sms;
        info->info = h_info;
        info->diag_min = h_mm[0];
        info->diag_max = h_mm[1];
        info->kernel_launches += vb_launch_counter - launches0;
    }
    prof_collect(info);
    if (h_info > 0) {
        fail(h_info, "normal matrix not positive definite: pivot %d <= 0", h_info);
        return h_info;
    }
    if (info) info->backward_err = -1.0;
    return diag_condition(L, T, cols, damping, info, st, rhs);  // rhs now holds the scaled solution
}

int vb200_gram_solve_multi(const double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                           const double *dampings, int32_t n_dampings, double *out_params, int32_t *out_info,
                           vb200_fit_info *info)
{
    if (cols <= 0 || total_rows <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (n_dampings <= 0) return fail(VB200_E_BADARG, "need at least one damping");
    if (!gram_dev || !stats_dev || !dampings || !out_params) return fail(VB200_E_BADARG, "null pointer");
    if (is_device_ptr(dampings)) return fail(VB200_E_BADARG, "dampings must be a host array");
    for (int b = 0; b < n_dampings; ++b)
        if (!(dampings[b] > 0.0)) return fail(VB200_E_BADARG, "damping must be > 0 (the damping=None branch of the reference is out of scope)");
    if (!is_device_ptr(gram_dev) || !is_device_ptr(stats_dev))
        return fail(VB200_E_BADARG, "gram/stats buffers must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    const long long stride = vb200_gram_doubles(cols);
    // how many factors fit side by side: at most 60 % of what is free (the workspace is cached, and the
    // caller still needs room for its own tensors, panels and predict scratch afterwards)
    long long want = n_dampings > 64 ? 64 : n_dampings;
    if (g_opt.sweep_group > 0 && want > g_opt.sweep_group) want = g_opt.sweep_group;
    size_t cached_b = 0;
    {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        auto it = g_ws.find(std::to_string(dev) + ":chol_batch");
        if (it != g_ws.end()) cached_b = it->second.bytes;  // re-usable
    }
    long long group = want;
    if (cached_b < sizeof(double) * (size_t)stride * (size_t)want) {
        // the cached batch workspace does not hold the whole sweep yet: ask the driver what is free (cudaMemGetInfo
        // costs milliseconds on a 180 GB device, so a sweep whose workspace already exists -- every (fold, mindist)
        // job of a SplineCV after the first -- does not ask again)
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        free_b += cached_b;
        group = (long long)((size_t)(0.6 * (double)free_b) / (sizeof(double) * (size_t)stride));
        if (group > want) group = want;
    }
    if (group < 1) return fail(VB200_E_NOMEM, "not enough device memory for one %lld x %lld factor", (long long)cols, (long long)cols);
    double *Lb, *inv_scale, *rhs, *xb, *damp_dev, *mm;
    int *d_info, *d_flags;
    CU(ws_get("chol_batch", sizeof(double) * (size_t)stride * group, (void **)&Lb));
    CU(ws_get("inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(ws_get("rhs_batch", sizeof(double) * Mpad * group, (void **)&xb));
    CU(ws_get("damp_batch", sizeof(double) * group, (void **)&damp_dev));
    CU(ws_get("minmax", sizeof(double) * 2, (void **)&mm));
    CU(ws_get("info_batch", sizeof(int) * group, (void **)&d_info));
    CU(ws_get("sweep_flags_batch", sizeof(int) * 2 * T * group, (void **)&d_flags));
    OutBuf ob;
    CU(stage_out("out_params_batch", out_params, (long long)cols * n_dampings, &ob));
    if (info) {
        info->block_rows = T;
        info->cols = cols;
        if (info->rows == 0) info->rows = total_rows;
    }
    prof_reset();
    const long long launches0 = vb_launch_counter;
    std::vector<int> h_info(n_dampings, 0);
    double fms = 0.0, sms = 0.0;
    double h_mm[2] = {0, 0};
    CU(vb_launch_column_scale(stats_dev, T, cols, total_rows, inv_scale, rhs, st));
    for (int b0 = 0; b0 < n_dampings; b0 += (int)group) {
        const int nb = (int)((n_dampings - b0 < group) ? (n_dampings - b0) : group);
        Span fac, sol;
        CU(cudaMemcpyAsync(damp_dev, dampings + b0, sizeof(double) * nb, cudaMemcpyHostToDevice, st));
        CU(span_begin(&fac, st));
        CU(vb_launch_scale_system_batched(gram_dev, T, cols, inv_scale, damp_dev, nb, Lb, stride, st));
        CU(vb_cholesky_tiled(Lb, T, g_opt.chol_panel_tiles, d_info, st, nb, stride, g_opt.chol_lookahead));
        CU(cudaEventRecord(fac.b, st));
        CU(span_begin(&sol, st));
        vb_replicate_kernel<<<dim3((unsigned)((Mpad + 255) / 256), nb), 256, 0, st>>>(rhs, Mpad, xb);
        ++vb_launch_counter;
        CU(vb_solve_tiled(Lb, T, xb, d_flags, st, nb, stride));
        vb_unscale_batched_kernel<<<dim3((unsigned)((cols + 255) / 256), nb), 256, 0, st>>>(xb, Mpad, inv_scale, cols,
                                                                                            ob.dev + (long long)b0 * cols);
        vb_diag_minmax_kernel<<<1, 256, 0, st>>>(Lb, T, cols, mm);  // first system of the group (smallest damping first in SplineCV)
        vb_launch_counter += 2;
        CU(cudaGetLastError());
        CU(cudaEventRecord(sol.b, st));
        CU(cudaMemcpyAsync(h_info.data() + b0, d_info, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
        if (b0 == 0) CU(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        fms += span_ms(fac);
        sms += span_ms(sol);
    }
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    int first_bad = 0;
    for (int b = 0; b < n_dampings; ++b) {
        if (out_info) out_info[b] = h_info[b];
        if (h_info[b] > 0 && first_bad == 0) first_bad = h_info[b];
    }
    if (info) {
        info->factor_ms += fms;
        info->solve_ms += sms;
        info->info = first_bad;
        info->diag_min = h_mm[0];
        info->diag_max = h_mm[1];
        info->kernel_launches += vb_launch_counter - launches0;
    }
    prof_collect(info);
    if (first_bad > 0) {
        fail(first_bad, "normal matrix not positive definite: pivot %d <= 0", first_bad);
        return first_bad;
    }
    return 0;
}

// ---- stepwise factorisation for the multi-GPU panel-cyclic Cholesky ------------------------------
// State between begin / panel / update / finish lives in the cached workspace (inv_scale, rhs, info)
// and in g_step; all launches are asynchronous on the library stream until finish.
namespace {
struct StepState {
    bool open = false;
    int T = 0;
    long long cols = 0;
    double damping = 0.0;
    Span fac;
    long long launches0 = 0;
} g_step;
}  // namespace

int vb200_chol_begin(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows, double damping)
{
    if (cols <= 0 || total_rows <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (!gram_dev || !stats_dev) return fail(VB200_E_BADARG, "null pointer");
    if (!(damping > 0.0)) return fail(VB200_E_BADARG, "damping must be > 0 (the damping=None branch of the reference is out of scope)");
    if (!is_device_ptr(gram_dev) || !is_device_ptr(stats_dev))
        return fail(VB200_E_BADARG, "gram/stats buffers must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    double *inv_scale, *rhs;
    int *d_info;
    CU(ws_get("inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    prof_reset();
    g_step.open = true;
    g_step.T = T;
    g_step.cols = cols;
    g_step.damping = damping;
    g_step.launches0 = vb_launch_counter;
    CU(span_begin(&g_step.fac, st));
    CU(cudaMemsetAsync(d_info, 0, sizeof(int), st));
    CU(vb_launch_column_scale(stats_dev, T, cols, total_rows, inv_scale, rhs, st));
    CU(vb_launch_scale_system(gram_dev, T, cols, inv_scale, damping, st));
    return 0;
}

int vb200_chol_panel_range(int64_t cols, int32_t k0, int32_t pw, int64_t *offset, int64_t *count)
{
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    if (cols <= 0 || k0 < 0 || pw < 1 || k0 + pw > T || !offset || !count) return fail(VB200_E_BADARG, "bad panel range");
    *offset = vb_chol_column_offset(k0, T);
    *count = vb_chol_column_offset(k0 + pw, T) - *offset;
    return 0;
}

int vb200_chol_panel(double *chol_dev, int64_t cols, int32_t k0, int32_t pw)
{
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    if (!g_step.open || g_step.cols != cols) return fail(VB200_E_BADARG, "vb200_chol_begin was not called for this system");
    if (!chol_dev || k0 < 0 || pw < 1 || pw > VB_MAX_KCHUNKS || k0 + pw > T) return fail(VB200_E_BADARG, "bad panel");
    int *d_info;
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    CU(vb_chol_factor_panel(chol_dev, T, k0, pw, d_info, 0));
    return 0;
}

int vb200_chol_update(double *chol_dev, int64_t cols, int32_t k0, int32_t pw, int32_t j_lo, int32_t j_hi)
{
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    if (!g_step.open || g_step.cols != cols) return fail(VB200_E_BADARG, "vb200_chol_begin was not called for this system");
    if (!chol_dev || k0 < 0 || pw < 1 || pw > VB_MAX_KCHUNKS || k0 + pw > T || j_lo < k0 + pw || j_hi > T)
        return fail(VB200_E_BADARG, "bad update range");
    CU(vb_chol_trailing_update(chol_dev, T, k0, pw, j_lo, j_hi, 0));
    return 0;
}

int vb200_chol_update_ranges(double *chol_dev, int64_t cols, int32_t k0, int32_t pw, int32_t n_ranges,
                             const int32_t *j_lo, const int32_t *j_hi)
{
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    if (!g_step.open || g_step.cols != cols) return fail(VB200_E_BADARG, "vb200_chol_begin was not called for this system");
    if (!chol_dev || k0 < 0 || pw < 1 || pw > VB_MAX_KCHUNKS || k0 + pw > T || n_ranges < 0 || (n_ranges > 0 && (!j_lo || !j_hi)))
        return fail(VB200_E_BADARG, "bad update ranges");
    for (int r = 0; r < n_ranges; ++r)
        if (j_lo[r] < k0 + pw || j_hi[r] > T || j_hi[r] < j_lo[r] || (r > 0 && j_lo[r] < j_hi[r - 1]))
            return fail(VB200_E_BADARG, "update ranges must be ascending, disjoint and right of the panel");
    if (n_ranges == 0) return 0;
    CU(vb_chol_trailing_update_ranges(chol_dev, T, k0, pw, n_ranges, j_lo, j_hi, 0));
    return 0;
}

// ---- device buffers shared between the processes of one node (CUDA IPC) ---------------------------------------
int vb200_shared_alloc(int64_t bytes, void **ptr, void *handle)
{
    if (bytes <= 0 || !ptr || !handle) return fail(VB200_E_BADARG, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI documents a 64-byte handle");
    void *p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(VB200_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return 0;
}

int vb200_shared_free(void *ptr)
{
    if (ptr) CU(cudaFree(ptr));
    return 0;
}

int vb200_shared_open(const void *handle, void **ptr)
{
    if (!handle || !ptr) return fail(VB200_E_BADARG, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    // opened in THIS rank's device context: the peer's memory becomes addressable from this GPU over NVLink
    CU(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int vb200_shared_close(void *ptr)
{
    if (ptr) CU(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int vb200_peer_copy_async(void *dst, const void *src, int64_t bytes, void *stream)
{
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(VB200_E_BADARG, "bad copy");
    if (bytes == 0) return 0;
    // unified addressing resolves both sides; between two GPUs this is an NVLink peer copy on a copy engine
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}

int vb200_chol_wait_panel(const int64_t *flag_dev, int64_t epoch)
{
    if (!flag_dev || !is_device_ptr(flag_dev)) return fail(VB200_E_BADARG, "the flag must be device memory");
    int *d_info;
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    CU(vb_chol_wait_flag(reinterpret_cast<const long long *>(flag_dev), (long long)epoch, d_info, 0));
    return 0;
}

int vb200_stream_wait_flag(const int64_t *flag_dev, int64_t epoch, void *stream)
{
    if (!flag_dev || !is_device_ptr(flag_dev)) return fail(VB200_E_BADARG, "the flag must be device memory");
    int *d_info;
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    CU(vb_chol_wait_flag(reinterpret_cast<const long long *>(flag_dev), (long long)epoch, d_info,
                         static_cast<cudaStream_t>(stream)));
    return 0;
}

int vb200_chol_finish(const double *chol_dev, int64_t cols, double *out_params, vb200_fit_info *info)
{
    if (!g_step.open || g_step.cols != cols) return fail(VB200_E_BADARG, "vb200_chol_begin was not called for this system");
    if (!chol_dev || !out_params) return fail(VB200_E_BADARG, "null pointer");
    g_step.open = false;
    cudaStream_t st = 0;
    const int T = g_step.T;
    const long long Mpad = (long long)T * VB_TILE;
    double *inv_scale, *rhs, *mm;
    int *d_info, *d_flags;
    CU(ws_get("inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(ws_get("minmax", sizeof(double) * 2, (void **)&mm));
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    CU(ws_get("sweep_flags", sizeof(int) * 2 * T, (void **)&d_flags));
    OutBuf ob;
    CU(stage_out("out_params", out_params, cols, &ob));
    CU(cudaEventRecord(g_step.fac.b, st));
    Span sol;
    CU(span_begin(&sol, st));
    CU(vb_solve_tiled(chol_dev, T, rhs, d_flags, st));
    vb_unscale_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(rhs, inv_scale, cols, ob.dev);
    vb_diag_minmax_kernel<<<1, 256, 0, st>>>(chol_dev, T, cols, mm);
    vb_launch_counter += 2;
    CU(cudaGetLastError());
    CU(cudaEventRecord(sol.b, st));
    CU(finish_out(ob, st));
    int h_info = 0;
    double h_mm[2] = {0, 0};
    CU(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double fms = span_ms(g_step.fac), sms = span_ms(sol);
    if (info) {
        info->block_rows = T;
        info->cols = cols;
        info->factor_ms += fms;
        info->solve_ms += sms;
        info->info = h_info;
        info->diag_min = h_mm[0];
        info->diag_max = h_mm[1];
        info->kernel_launches += vb_launch_counter - g_step.launches0;
    }
    prof_collect(info);
    if (h_info > 0) {
        fail(h_info, "normal matrix not positive definite: pivot %d <= 0", h_info);
        return h_info;
    }
    if (h_info < 0) return fail(VB200_E_CUDA, "a panel pushed by a peer GPU did not arrive within the time limit");
    if (info) info->backward_err = -1.0;
    return diag_condition(chol_dev, T, cols, g_step.damping, info, st, rhs);
}

// ---- block reductions and the distance mask (kernels in vb_block.cu) -------
namespace {
struct BlockState {
    bool open = false;
    long long n = 0, nseg = 0;
    VbBlockWorkspace ws = {};
    long long *labels = nullptr;
} g_block;

__global__ void vb_bbox_partial_kernel(const double *__restrict__ e, const double *__restrict__ n_, long long n,
                                       double *__restrict__ out4)
{
    __shared__ double s[4][256];
    double xmin = 1.0 / 0.0, xmax = -1.0 / 0.0, ymin = 1.0 / 0.0, ymax = -1.0 / 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        xmin = fmin(xmin, e[i]);
        xmax = fmax(xmax, e[i]);
        ymin = fmin(ymin, n_[i]);
        ymax = fmax(ymax, n_[i]);
    }
    s[0][threadIdx.x] = xmin;
    s[1][threadIdx.x] = xmax;
    s[2][threadIdx.x] = ymin;
    s[3][threadIdx.x] = ymax;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            s[0][threadIdx.x] = fmin(s[0][threadIdx.x], s[0][threadIdx.x + st]);
            s[1][threadIdx.x] = fmax(s[1][threadIdx.x], s[1][threadIdx.x + st]);
            s[2][threadIdx.x] = fmin(s[2][threadIdx.x], s[2][threadIdx.x + st]);
            s[3][threadIdx.x] = fmax(s[3][threadIdx.x], s[3][threadIdx.x + st]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) out4[blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x][0];
}
}  // namespace

int vb200_block_split(const double *east, const double *north, int64_t n, const double *centers_east,
                      int32_t n_east, const double *centers_north, int32_t n_north, int64_t *labels,
                      int64_t *n_blocks)
{
    g_block.open = false;
    if (n < 0 || n_east < 1 || n_north < 1) return fail(VB200_E_BADARG, "bad sizes");
    if (n > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many points (limit 2^31 - 1)");
    if ((long long)n_east * n_north > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many blocks (limit 2^31 - 1)");
    if (!n_blocks || !centers_east || !centers_north || (n > 0 && (!east || !north))) return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    if (n == 0) {
        *n_blocks = 0;
        g_block.open = true;
        g_block.n = 0;
        g_block.nseg = 0;
        return 0;
    }
    const double *de, *dn, *dce, *dcn;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("blk_ce", centers_east, n_east, &dce, st));
    CU(stage_in("blk_cn", centers_north, n_north, &dcn, st));
    VbBlockWorkspace &ws = g_block.ws;
    const long long ntiles = (n + 2047) / 2048;
    const long long scan_elems = vb_block_scan_scratch_elems(256 * ntiles > n ? 256 * ntiles : n) + 16;
    CU(ws_get("blk_keys0", sizeof(uint64_t) * n, (void **)&ws.keys[0]));
    CU(ws_get("blk_keys1", sizeof(uint64_t) * n, (void **)&ws.keys[1]));
    CU(ws_get("blk_vals0", sizeof(unsigned) * n, (void **)&ws.vals[0]));
    CU(ws_get("blk_vals1", sizeof(unsigned) * n, (void **)&ws.vals[1]));
    CU(ws_get("blk_counts", sizeof(unsigned) * 256 * ntiles, (void **)&ws.counts));
    CU(ws_get("blk_scratch", sizeof(unsigned) * scan_elems, (void **)&ws.scratch));
    CU(ws_get("blk_flags", sizeof(unsigned) * n, (void **)&ws.flags));
    CU(ws_get("blk_nseg", sizeof(unsigned), (void **)&ws.nseg_dev));
    CU(ws_get("blk_labels", sizeof(long long) * n, (void **)&g_block.labels));
    CU(vb_block_split(de, dn, n, dce, n_east, dcn, n_north, ws, g_block.labels, st));
    unsigned nseg = 0;
    CU(cudaMemcpyAsync(&nseg, ws.nseg_dev, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(ws_get("blk_seg_start", sizeof(unsigned) * ((size_t)nseg + 1), (void **)&ws.seg_start));
    CU(ws_get("blk_seg_label", sizeof(long long) * (size_t)nseg, (void **)&ws.seg_label));
    CU(vb_block_heads(n, nseg, ws, st));
    if (labels) CU(cudaMemcpyAsync(labels, g_block.labels, sizeof(long long) * n, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    g_block.open = true;
    g_block.n = n;
    g_block.nseg = nseg;
    *n_blocks = nseg;
    return 0;
}

int vb200_block_ids(int64_t *ids)
{
    if (!g_block.open) return fail(VB200_E_BADARG, "vb200_block_split was not called");
    if (g_block.nseg == 0) return 0;
    if (!ids) return fail(VB200_E_BADARG, "null pointer");
    CU(cudaMemcpy(ids, g_block.ws.seg_label, sizeof(long long) * g_block.nseg, cudaMemcpyDefault));
    return 0;
}

int vb200_block_aggregate(const double *values, const double *weights, int64_t n, int32_t op, double *out)
{
    if (!g_block.open) return fail(VB200_E_BADARG, "vb200_block_split was not called");
    if (n != g_block.n) return fail(VB200_E_BADARG, "values must have one entry per point of the last vb200_block_split (%lld given, %lld expected)", (long long)n, g_block.n);
    if (op < VB_REDUCE_MEAN || op > VB_REDUCE_LAST) return fail(VB200_E_BADARG, "unknown reduction %d", (int)op);
    if (g_block.nseg == 0) return 0;
    const bool needs_w = op == VB_REDUCE_WMEAN || op == VB_REDUCE_WVAR || op == VB_REDUCE_INV_SUM_W;
    const bool needs_v = op != VB_REDUCE_COUNT && op != VB_REDUCE_INV_SUM_W;
    if (!out || (needs_v && !values) || (needs_w && !weights)) return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    const double *dv = nullptr, *dw = nullptr;
    if (needs_v) CU(stage_in("in_d", values, n, &dv, st));
    if (needs_w) CU(stage_in("in_w", weights, n, &dw, st));
    VbBlockWorkspace &ws = g_block.ws;
    if (op == VB_REDUCE_MEDIAN) {
        CU(ws_get("blk_mkeys0", sizeof(uint64_t) * n, (void **)&ws.mkeys[0]));
        CU(ws_get("blk_mkeys1", sizeof(uint64_t) * n, (void **)&ws.mkeys[1]));
        CU(ws_get("blk_mvals0", sizeof(unsigned) * n, (void **)&ws.mvals[0]));
        CU(ws_get("blk_mvals1", sizeof(unsigned) * n, (void **)&ws.mvals[1]));
    }
    OutBuf ob;
    CU(stage_out("blk_out", out, g_block.nseg, &ob));
    CU(vb_block_aggregate(dv, dw, n, g_block.nseg, op, g_block.labels, ws, ob.dev, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_block_reduce_stream(const double *east, const double *north, int64_t n, const double *centers_east,
                              int32_t n_east, const double *centers_north, int32_t n_north, const double *values,
                              int32_t n_columns, const double *weights, int32_t op, int64_t capacity, int64_t *out_ids,
                              double *out_values, int64_t *n_blocks)
{
    if (n < 0 || n_east < 1 || n_north < 1 || n_columns < 0 || capacity < 0) return fail(VB200_E_BADARG, "bad sizes");
    const long long nblk = (long long)n_east * n_north;
    if (nblk > (1LL << 26)) return fail(VB200_E_BADARG, "too many blocks for dense accumulators (limit 2^26): use vb200_block_split");
    if (op != VB_REDUCE_SUM && op != VB_REDUCE_MEAN && op != VB_REDUCE_COUNT && op != VB_REDUCE_WMEAN)
        return fail(VB200_E_BADARG, "the one-pass reduction offers sum, mean, count and weighted mean (got %d)", (int)op);
    if ((op == VB_REDUCE_WMEAN) != (weights != nullptr)) return fail(VB200_E_BADARG, "weights go with VB200_REDUCE_WMEAN, and only with it");
    if (!n_blocks || !centers_east || !centers_north || (n > 0 && (!east || !north)) || (n_columns > 0 && n > 0 && !values))
        return fail(VB200_E_BADARG, "null pointer");
    if (n == 0) {
        *n_blocks = 0;
        return 0;
    }
    if (!out_ids || (n_columns > 0 && !out_values)) return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    const double *de, *dn, *dce, *dcn, *dv = nullptr, *dw = nullptr;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("blk_ce", centers_east, n_east, &dce, st));
    CU(stage_in("blk_cn", centers_north, n_north, &dcn, st));
    if (n_columns > 0) CU(stage_in("blk_cols", values, (long long)n_columns * n, &dv, st));
    if (weights) CU(stage_in("in_w", weights, n, &dw, st));
    unsigned *cnt, *flags, *scratch, *nseg_dev;
    double *sums, *sumw;
    CU(ws_get("blk_s_cnt", sizeof(unsigned) * nblk, (void **)&cnt));
    CU(ws_get("blk_s_flags", sizeof(unsigned) * nblk, (void **)&flags));
    CU(ws_get("blk_s_scratch", sizeof(unsigned) * (vb_block_scan_scratch_elems(nblk) + 16), (void **)&scratch));
    CU(ws_get("blk_nseg", sizeof(unsigned), (void **)&nseg_dev));
    CU(ws_get("blk_s_sums", sizeof(double) * nblk * (n_columns > 0 ? n_columns : 1), (void **)&sums));
    CU(ws_get("blk_s_sumw", sizeof(double) * nblk, (void **)&sumw));
    // outputs: device pointers are written in place, host pointers through a staged copy
    long long *ids_dev;
    const bool ids_staged = !is_device_ptr(out_ids);
    if (ids_staged) CU(ws_get("blk_s_ids", sizeof(long long) * (capacity > 0 ? capacity : 1), (void **)&ids_dev));
    else ids_dev = reinterpret_cast<long long *>(out_ids);
    OutBuf ob;
    CU(stage_out("blk_s_out", out_values, (long long)n_columns * capacity, &ob));
    CU(vb_block_stream(de, dn, n, dce, n_east, dcn, n_north, dv, n_columns, dw, op, cnt, sums, sumw, flags, scratch, nseg_dev,
                       capacity, ids_dev, ob.dev, st));
    unsigned nseg = 0;
    CU(cudaMemcpyAsync(&nseg, nseg_dev, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_blocks = nseg;
    if ((long long)nseg > capacity) return fail(VB200_E_BADARG, "%u non-empty blocks do not fit the output capacity %lld", nseg, (long long)capacity);
    if (ids_staged && nseg) CU(cudaMemcpyAsync(out_ids, ids_dev, sizeof(long long) * nseg, cudaMemcpyDeviceToHost, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_distance_mask(const double *data_east, const double *data_north, int64_t n, double maxdist,
                        const double *east, const double *north, int64_t n_points, uint8_t *mask)
{
    if (n < 0 || n_points < 0) return fail(VB200_E_BADARG, "negative size");
    if (n > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many data points (limit 2^31 - 1)");
    if (n_points == 0) return 0;
    if (!mask || !east || !north || (n > 0 && (!data_east || !data_north))) return fail(VB200_E_BADARG, "null pointer");
    if (maxdist != maxdist) return fail(VB200_E_BADARG, "maxdist is NaN");
    cudaStream_t st = 0;
    const bool mask_on_device = is_device_ptr(mask);
    unsigned char *dmask = mask;
    if (!mask_on_device) CU(ws_get("mask_out", (size_t)n_points, (void **)&dmask));
    if (n == 0) {
        CU(cudaMemsetAsync(dmask, 0, (size_t)n_points, st));
    } else if (maxdist >= 1e300) {
        CU(cudaMemsetAsync(dmask, 1, (size_t)n_points, st));  // everything is within an infinite distance
    } else {
        const double *de, *dn, *qe, *qn;
        CU(stage_in("in_fe", data_east, n, &de, st));
        CU(stage_in("in_fn", data_north, n, &dn, st));
        CU(stage_in("in_e", east, n_points, &qe, st));
        CU(stage_in("in_n", north, n_points, &qn, st));
        // bounding box of the data -> uniform cell grid of pitch >= maxdist, at most 4096 x 4096 cells
        const int parts = 512;
        double *d_bbox, h_bbox[4 * 512];
        CU(ws_get("mask_bbox", sizeof(double) * 4 * parts, (void **)&d_bbox));
        vb_bbox_partial_kernel<<<parts, 256, 0, st>>>(de, dn, n, d_bbox);
        ++vb_launch_counter;
        CU(cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        double xmin = h_bbox[0], xmax = h_bbox[1], ymin = h_bbox[2], ymax = h_bbox[3];
        for (int p = 1; p < parts; ++p) {
            xmin = h_bbox[4 * p] < xmin ? h_bbox[4 * p] : xmin;
            xmax = h_bbox[4 * p + 1] > xmax ? h_bbox[4 * p + 1] : xmax;
            ymin = h_bbox[4 * p + 2] < ymin ? h_bbox[4 * p + 2] : ymin;
            ymax = h_bbox[4 * p + 3] > ymax ? h_bbox[4 * p + 3] : ymax;
        }
        if (!(xmax >= xmin) || !(ymax >= ymin) || !(xmax - xmin < 1e300) || !(ymax - ymin < 1e300))
            return fail(VB200_E_BADARG, "data coordinates must be finite");
        const double extent = (xmax - xmin) > (ymax - ymin) ? (xmax - xmin) : (ymax - ymin);
        // a hair wider than maxdist: the cell index is (int)((x - x0) * (1 / pitch)) with two roundings, and a data
        // point and a node at distance <= maxdist must never land two cells apart (3 x 3 scan)
        double pitch = maxdist > 0.0 ? maxdist * (1.0 + 1e-9) : 0.0;
        if (pitch < extent / 4096.0) pitch = extent / 4096.0;
        if (!(pitch > 0.0) || !(pitch < 1e300)) pitch = 1.0;
        const int nx = (int)((xmax - xmin) / pitch) + 1, ny = (int)((ymax - ymin) / pitch) + 1;
        const long long ncell = (long long)nx * ny;
        unsigned *cell_of, *cell_start, *cursor, *scratch;
        double *sorted_xy;
        CU(ws_get("mask_cell_of", sizeof(unsigned) * n, (void **)&cell_of));
        CU(ws_get("mask_cell_start", sizeof(unsigned) * (ncell + 1), (void **)&cell_start));
        CU(ws_get("mask_cursor", sizeof(unsigned) * ncell, (void **)&cursor));
        CU(ws_get("mask_scratch", sizeof(unsigned) * (vb_block_scan_scratch_elems(ncell + 1) + 16), (void **)&scratch));
        CU(ws_get("mask_sorted_xy", sizeof(double) * 2 * n, (void **)&sorted_xy));
        CU(vb_distance_mask(de, dn, n, xmin, ymin, pitch, nx, ny, cell_of, cell_start, cursor, scratch, sorted_xy, qe, qn,
                            n_points, maxdist, dmask, st));
    }
    if (!mask_on_device) CU(cudaMemcpyAsync(mask, dmask, (size_t)n_points, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int64_t vb200_padded_cols(int64_t cols) { return (cols + VB_TILE - 1) / VB_TILE * VB_TILE; }

int vb200_gram_factor(double *gram_dev, const double *stats_dev, int64_t cols, int64_t total_rows,
                      double damping, double *inv_scale_dev, vb200_fit_info *info)
{
    if (cols <= 0 || total_rows <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (!gram_dev || !stats_dev || !inv_scale_dev) return fail(VB200_E_BADARG, "null pointer");
    if (!(damping > 0.0)) return fail(VB200_E_BADARG, "damping must be > 0 (the damping=None branch of the reference is out of scope)");
    if (!is_device_ptr(gram_dev) || !is_device_ptr(stats_dev) || !is_device_ptr(inv_scale_dev))
        return fail(VB200_E_BADARG, "gram/stats/inv_scale buffers must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    double *rhs, *mm;
    int *d_info;
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(ws_get("minmax", sizeof(double) * 2, (void **)&mm));
    CU(ws_get("info", sizeof(int), (void **)&d_info));
    if (info) {
        info->block_rows = T;
        info->cols = cols;
        if (info->rows == 0) info->rows = total_rows;
    }
    prof_reset();
    const long long launches0 = vb_launch_counter;
    Span fac;
    CU(span_begin(&fac, st));
    CU(vb_launch_column_scale(stats_dev, T, cols, total_rows, inv_scale_dev, rhs, st));
    CU(vb_launch_scale_system(gram_dev, T, cols, inv_scale_dev, damping, st));
    CU(vb_cholesky_tiled(gram_dev, T, g_opt.chol_panel_tiles, d_info, st, 1, 0, g_opt.chol_lookahead));
    vb_diag_minmax_kernel<<<1, 256, 0, st>>>(gram_dev, T, cols, mm);
    ++vb_launch_counter;
    CU(cudaGetLastError());
    CU(cudaEventRecord(fac.b, st));
    int h_info = 0;
    double h_mm[2] = {0, 0};
    CU(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double fms = span_ms(fac);
    if (info) {
        info->factor_ms += fms;
        info->info = h_info;
        info->diag_min = h_mm[0];
        info->diag_max = h_mm[1];
        info->kernel_launches += vb_launch_counter - launches0;
    }
    prof_collect(info);
    if (h_info > 0) {
        fail(h_info, "normal matrix not positive definite: pivot %d <= 0", h_info);
        return h_info;
    }
    if (info) info->backward_err = -1.0;
    return diag_condition(gram_dev, T, cols, damping, info, st);
}

int vb200_chol_solve(const double *chol_dev, const double *inv_scale_dev, int64_t cols, const double *rhs,
                     double *out_params, vb200_fit_info *info)
{
    if (cols <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (!chol_dev || !inv_scale_dev || !rhs || !out_params) return fail(VB200_E_BADARG, "null pointer");
    if (!is_device_ptr(chol_dev) || !is_device_ptr(inv_scale_dev))
        return fail(VB200_E_BADARG, "factor / inv_scale buffers must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    const double *drhs;
    CU(stage_in("in_rhs", rhs, cols, &drhs, st));
    double *x;
    int *d_flags;
    CU(ws_get("rhs", sizeof(double) * Mpad, (void **)&x));
    CU(ws_get("sweep_flags", sizeof(int) * 2 * T, (void **)&d_flags));
    OutBuf ob;
    CU(stage_out("out_params", out_params, cols, &ob));
    const long long launches0 = vb_launch_counter;
    Span sol;
    CU(span_begin(&sol, st));
    CU(cudaMemsetAsync(x, 0, sizeof(double) * Mpad, st));
    vb_unscale_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(drhs, inv_scale_dev, cols, x);  // D rhs
    CU(vb_solve_tiled(chol_dev, T, x, d_flags, st));
    vb_unscale_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(x, inv_scale_dev, cols, ob.dev);  // D c
    vb_launch_counter += 2;
    CU(cudaGetLastError());
    CU(cudaEventRecord(sol.b, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    const double sms = span_ms(sol);
    if (info) {
        info->solve_ms += sms;
        info->kernel_launches += vb_launch_counter - launches0;
    }
    return 0;
}

int vb200_jacobian_transpose_apply(const double *east, const double *north, int64_t n,
                                   const double *force_east, const double *force_north, int64_t m,
                                   double mindist, const double *v, double *out)
{
    if (n < 0 || m < 0) return fail(VB200_E_BADARG, "negative size");
    if (m == 0) return 0;
    if (!force_east || !force_north || !out || (n > 0 && (!east || !north || !v)))
        return fail(VB200_E_BADARG, "null pointer");
    if (n > 0x7fffffffLL) return fail(VB200_E_BADARG, "too many data points");
    cudaStream_t st = 0;
    const double *de, *dn, *dfe, *dfn, *dv;
    CU(stage_in("in_e", east, n, &de, st));
    CU(stage_in("in_n", north, n, &dn, st));
    CU(stage_in("in_fe", force_east, m, &dfe, st));
    CU(stage_in("in_fn", force_north, m, &dfn, st));
    CU(stage_in("in_f", v, n, &dv, st));
    OutBuf ob;
    CU(stage_out("out_pred", out, m, &ob));
    // (J^T v)_j = sum_i g(e_i - fe_j, n_i - fn_j) v_i: the predict kernel with the roles of points and
    // forces swapped (g is even), with the exact, reference-ordered Green's function
    int pts, nsplit;
    vb_predict_plan(m, n, &pts, &nsplit);
    double *scratch = nullptr;
    if (nsplit > 1) CU(ws_get("pred_split", sizeof(double) * nsplit * m, (void **)&scratch));
    CU(vb_launch_predict(dfe, dfn, m, de, dn, dv, n, mindist, 1, pts, nsplit, scratch, ob.dev, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

static int fit_impl(int kind, const double *jac, const double *east, const double *north,
                    const double *data, const double *weights, int64_t n, const double *force_east,
                    const double *force_north, int64_t m, double mindist, double poisson,
                    double damping, double *out_force, vb200_fit_info *info)
{
    if (info) memset(info, 0, sizeof(*info));
    if (n <= 0 || m <= 0) return fail(VB200_E_BADARG, "bad sizes n=%lld m=%lld", (long long)n, (long long)m);
    if (!(damping > 0.0)) return fail(VB200_E_BADARG, "damping must be > 0 (the damping=None branch of the reference is out of scope)");
    const int64_t cols = kind == 1 ? 2 * m : m;
    const int64_t rows = kind == 1 ? 2 * n : n;
    double *G, *stats;
    CU(ws_get("gram", sizeof(double) * vb200_gram_doubles(cols), (void **)&G));
    CU(ws_get("stats", sizeof(double) * vb200_stats_doubles(cols), (void **)&stats));
    int rc = gram_accumulate_impl(kind, jac, east, north, data, weights, n, force_east, force_north, m,
                                  mindist, poisson, G, stats, 0, info);
    if (rc != 0) return rc;
    rc = vb200_gram_solve(G, stats, cols, rows, damping, 0, out_force, info);
    if (rc != 0 || !info || !g_opt.fit_diagnostics) return rc;
    // backward error of the forces just computed (the inputs are still staged in their workspace slots)
    const double t0 = wall_ms();
    const long long launches0 = vb_launch_counter;
    cudaStream_t st = 0;
    const VbGramProblem pb = g_last_problem;  // device pointers staged by gram_accumulate_impl just above
    const double *force_dev = out_force;
    if (!is_device_ptr(out_force)) {
        void *staged;
        CU(ws_get("out_params", sizeof(double) * cols, &staged));  // vb200_gram_solve's staging copy
        force_dev = static_cast<const double *>(staged);
    }
    double *grad, *inv_scale;
    CU(ws_get("diag_grad", sizeof(double) * cols, (void **)&grad));
    CU(ws_get("inv_scale", sizeof(double) * vb200_padded_cols(cols), (void **)&inv_scale));
    rc = normal_residual_dev(pb, force_dev, grad, st);
    if (rc != 0) return rc;
    rc = backward_error_dev(grad, stats, inv_scale, force_dev, cols, damping, info->lambda_max, &info->backward_err, st);
    info->kernel_launches += vb_launch_counter - launches0;
    info->diag_ms += wall_ms() - t0;
    return rc;
}

int vb200_normal_residual(int kind, const double *east, const double *north, const double *data,
                          const double *weights, int64_t n, const double *force_east,
                          const double *force_north, int64_t m, double mindist, double poisson,
                          const double *force, double *grad)
{
    if (kind != 0 && kind != 1) return fail(VB200_E_BADARG, "kind must be 0 (scalar) or 1 (elastic 2-D)");
    if (n < 0 || m <= 0) return fail(VB200_E_BADARG, "bad sizes n=%lld m=%lld", (long long)n, (long long)m);
    if (!force_east || !force_north || !force || !grad || (n > 0 && (!east || !north || !data)))
        return fail(VB200_E_BADARG, "null pointer");
    if (n > 0x7fffffffLL || m > 0x3fffffffLL) return fail(VB200_E_BADARG, "too many points");
    cudaStream_t st = 0;
    VbGramProblem pb;
    int rc = stage_problem(kind, nullptr, east, north, data, weights, n, force_east, force_north, m, mindist, poisson,
                           &pb, st);
    if (rc != 0) return rc;
    const long long cols = vb_sys_cols(pb);
    const double *dforce;
    CU(stage_in("in_f", force, cols, &dforce, st));
    OutBuf ob;
    CU(stage_out("diag_grad", grad, cols, &ob));
    rc = normal_residual_dev(pb, dforce, ob.dev, st);
    if (rc != 0) return rc;
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int vb200_backward_error(const double *grad, const double *stats_dev, int64_t cols, int64_t total_rows,
                         double damping, const double *force, double lambda_max, double *eta)
{
    if (cols <= 0 || total_rows <= 0) return fail(VB200_E_BADARG, "bad sizes");
    if (!grad || !stats_dev || !force || !eta) return fail(VB200_E_BADARG, "null pointer");
    if (!is_device_ptr(stats_dev)) return fail(VB200_E_BADARG, "stats buffer must be device memory");
    cudaStream_t st = 0;
    const int T = (int)((cols + VB_TILE - 1) / VB_TILE);
    const long long Mpad = (long long)T * VB_TILE;
    const double *dgrad, *dforce;
    CU(stage_in("diag_grad_in", grad, cols, &dgrad, st));
    CU(stage_in("in_f", force, cols, &dforce, st));
    double *inv_scale, *rhs;
    CU(ws_get("diag_inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("diag_rhs", sizeof(double) * Mpad, (void **)&rhs));
    CU(vb_launch_column_scale(stats_dev, T, cols, total_rows, inv_scale, rhs, st));
    return backward_error_dev(dgrad, stats_dev, inv_scale, dforce, cols, damping, lambda_max, eta, st);
}

static int lstsq_impl(int kind, const double *jac, const double *east, const double *north, const double *data,
                      const double *weights, int64_t n, const double *force_east, const double *force_north,
                      int64_t m, double mindist, double poisson, double rcond, double *out_force, vb200_fit_info *info)
{
    if (info) memset(info, 0, sizeof(*info));
    if (n <= 0 || m <= 0) return fail(VB200_E_BADARG, "bad sizes n=%lld m=%lld", (long long)n, (long long)m);
    if (!(rcond >= 0.0) || !(rcond < 1.0)) return fail(VB200_E_BADARG, "rcond must be in [0, 1)");
    if (!data || !out_force || (kind == 2 ? !jac : (!east || !north || !force_east || !force_north)))
        return fail(VB200_E_BADARG, "null pointer");
    cudaStream_t st = 0;
    VbGramProblem pb;
    int rc = stage_problem(kind, jac, east, north, data, weights, n, force_east, force_north, m, mindist, poisson, &pb, st);
    if (rc != 0) return rc;
    const long long rows = vb_sys_rows(pb), cols = vb_sys_cols(pb);
    if (cols > 0x3fffffffLL) return fail(VB200_E_BADARG, "too many unknowns");
    const long long Mpad = vb200_padded_cols(cols), mp = cols + (cols & 1);
    const long long launches0 = vb_launch_counter;
    Span whole;
    CU(span_begin(&whole, st));
    const double *jdev = pb.jac;
    if (kind != 2) {
        double *jbuf;
        CU(ws_get("svd_jac", sizeof(double) * (size_t)rows * cols, (void **)&jbuf));
        if (kind == 0) CU(vb_launch_jacobian(pb.east, pb.north, n, pb.fe, pb.fn, m, mindist, jbuf, st));
        else CU(vb_launch_jacobian2d(pb.east, pb.north, n, pb.fe, pb.fn, m, mindist, poisson, jbuf, st));
        jdev = jbuf;
    }
    double *inv_scale, *stack, *work;
    unsigned *counter;
    CU(ws_get("inv_scale", sizeof(double) * Mpad, (void **)&inv_scale));
    CU(ws_get("svd_stack", sizeof(double) * (size_t)vb_svd_stack_doubles(rows, cols), (void **)&stack));
    CU(ws_get("svd_work", sizeof(double) * (size_t)(3 * mp + 8), (void **)&work));
    CU(ws_get("svd_counter", sizeof(unsigned), (void **)&counter));
    OutBuf ob;
    CU(stage_out("out_params", out_force, cols, &ob));
    CU(vb_svd_inv_scale(jdev, rows, cols, Mpad, inv_scale, st));
    VbSvdResult res = {};
    CU(vb_svd_solve(jdev, pb.data, pb.weights, inv_scale, rows, cols, rcond, g_opt.svd_max_sweeps, stack, work, counter,
                    ob.dev, &res, st));
    CU(cudaEventRecord(whole.b, st));
    CU(finish_out(ob, st));
    CU(cudaStreamSynchronize(st));
    const double ms = span_ms(whole);
    if (info) {
        info->block_rows = (int32_t)(Mpad / VB_TILE);
        info->rows = rows;
        info->cols = cols;
        info->solve_ms = ms;
        info->kernel_launches = vb_launch_counter - launches0;
        info->svd_rank = res.rank;
        info->svd_sweeps = res.sweeps;
        info->sigma_max = res.sigma_max;
        info->sigma_min_kept = res.sigma_min_kept;
        info->lambda_max = res.sigma_max * res.sigma_max;
        info->lambda_min_est = res.sigma_min_kept * res.sigma_min_kept;
        info->cond_est = res.sigma_min_kept > 0.0 ? res.sigma_max / res.sigma_min_kept : 0.0;
        info->cond_upper = rcond > 0.0 ? 1.0 / rcond : 0.0;
        info->backward_err = -1.0;
    }
    if (!res.converged)
        return fail(VB200_E_NOCONV, "one-sided Jacobi SVD did not converge in %d sweeps (%lld x %lld system)", res.sweeps,
                    rows, cols);
    return 0;
}

int vb200_spline_fit_lstsq(const double *east, const double *north, const double *data,
                           const double *weights, int64_t n, const double *force_east,
                           const double *force_north, int64_t m, double mindist, double rcond,
                           double *out_force, vb200_fit_info *info)
{
    return lstsq_impl(0, nullptr, east, north, data, weights, n, force_east, force_north, m, mindist, 0.0, rcond,
                      out_force, info);
}

int vb200_spline_fit_2d_lstsq(const double *east, const double *north, const double *data,
                              const double *weights, int64_t n, const double *force_east,
                              const double *force_north, int64_t m, double mindist, double poisson,
                              double rcond, double *out_force, vb200_fit_info *info)
{
    return lstsq_impl(1, nullptr, east, north, data, weights, n, force_east, force_north, m, mindist, poisson, rcond,
                      out_force, info);
}

int vb200_least_squares_lstsq(const double *jac, int64_t n, int64_t m, const double *data,
                              const double *weights, double rcond, double *out_params,
                              vb200_fit_info *info)
{
    return lstsq_impl(2, jac, nullptr, nullptr, data, weights, n, nullptr, nullptr, m, 0.0, 0.0, rcond, out_params, info);
}

int vb200_spline_fit(const double *east, const double *north, const double *data,
                     const double *weights, int64_t n, const double *force_east,
                     const double *force_north, int64_t m, double mindist, double damping,
                     double *out_force, vb200_fit_info *info)
{
    return fit_impl(0, nullptr, east, north, data, weights, n, force_east, force_north, m, mindist, 0.0,
                    damping, out_force, info);
}

int vb200_spline_fit_2d(const double *east, const double *north, const double *data,
                        const double *weights, int64_t n, const double *force_east,
                        const double *force_north, int64_t m, double mindist, double poisson,
                        double damping, double *out_force, vb200_fit_info *info)
{
    return fit_impl(1, nullptr, east, north, data, weights, n, force_east, force_north, m, mindist,
                    poisson, damping, out_force, info);
}

int vb200_least_squares(const double *jac, int64_t n, int64_t m, const double *data,
                        const double *weights, double damping, double *out_params,
                        vb200_fit_info *info)
{
    if (!jac) return fail(VB200_E_BADARG, "null jacobian");
    return fit_impl(2, jac, nullptr, nullptr, data, weights, n, nullptr, nullptr, m, 0.0, 0.0, damping,
                    out_params, info);
}

}  // extern "C"
