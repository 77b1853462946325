// vb_block.cu -- the steps either side of the spline path:
//
//   * block (window) reductions: verde.BlockReduce / verde.BlockMean (verde/blockreduce.py:99-243,
//     360-506) on top of verde.block_split (verde/coordinates.py:848-946).  The reference labels every
//     point with the nearest block centre through a KD-tree and aggregates with a pandas groupby;
//   * the distance mask: verde.distance_mask (verde/mask.py:17-113), a KD-tree nearest-neighbour
//     query per grid node followed by `distance <= maxdist`.
//
// Both are HBM-bound integer / byte work (no tensor cores): the design rules are coalesced streams,
// shared-memory staging of the per-CTA histograms and grids sized in multiples of the SM count.
//
//   block_split      label kernel (nearest centre per axis, 16 B in / 16 B out per point)
//                    -> stable LSD radix sort of (label, point index) pairs, 8 bits per pass, only the
//                       passes the label range needs (2 passes for 224 x 224 blocks)
//                    -> segment heads by a flag + exclusive scan (no dense per-block array: a
//                       10 000 x 10 000 block grid with few points costs nothing)
//   block_aggregate  one warp per non-empty block walks its segment IN ORIGINAL POINT ORDER (the sort is
//                    stable), compensated (Kahan) sums per lane, fixed shuffle tree across lanes:
//                    deterministic, and within 1e-15 relative of the pandas result.  The median sorts
//                    (label, value) pairs with the same radix sort (value pass first, stable label
//                    pass second).
//   distance_mask    data points binned into a uniform cell grid of pitch >= maxdist (counting sort:
//                    histogram, scan, scatter), coordinates gathered into cell order; one thread per
//                    query point scans the 3 x 3 neighbouring cells and stops at the first hit.
#include <stdint.h>

#include "vb_common.cuh"
#include "vb_internal.h"

namespace {

// ------------------------------------------------------------------------------------------
// exclusive scan of a uint32 array (in place), any length: per-CTA tiles of 2048, recursive
// ------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ unsigned cta_exclusive_scan(unsigned v, unsigned *total)
{
    // 256 threads: warp scan by shuffles, then the 8 warp totals
    __shared__ unsigned warp_tot[SC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    unsigned base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        const unsigned t = warp_tot[w];
        if (w < warp) base += t;
        all += t;
    }
    __syncthreads();
    *total = all;
    return base + x - v;
}

__global__ void __launch_bounds__(SC_THREADS)
vb_scan_tiles_kernel(unsigned *__restrict__ data, long long n, unsigned *__restrict__ partial)
{
    const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
    unsigned v[SC_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        sum += v[k];
    }
    unsigned total;
    unsigned run = cta_exclusive_scan(sum, &total);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS)
vb_scan_add_kernel(unsigned *__restrict__ data, long long n, const unsigned *__restrict__ partial)
{
    const unsigned add = partial[blockIdx.x];
    const long long base = (long long)blockIdx.x * SC_TILE;
    for (int k = threadIdx.x; k < SC_TILE; k += SC_THREADS)
        if (base + k < n) data[base + k] += add;
}

// scratch must hold scan_scratch_elems(n) uint32; *total_dev (optional) receives the grand total
long long scan_scratch_elems(long long n)
{
    long long total = 0;
    while (n > 1) {
        n = (n + SC_TILE - 1) / SC_TILE;
        total += n;
        if (n == 1) break;
    }
    return total + 1;
}

cudaError_t scan_u32(unsigned *data, long long n, unsigned *scratch, unsigned *total_dev, cudaStream_t st)
{
    if (n <= 0) {
        if (total_dev) return cudaMemsetAsync(total_dev, 0, sizeof(unsigned), st);
        return cudaSuccess;
    }
    const long long tiles = (n + SC_TILE - 1) / SC_TILE;
    vb_scan_tiles_kernel<<<(unsigned)tiles, SC_THREADS, 0, st>>>(data, n, scratch);
    ++vb_launch_counter;
    if (tiles > 1) {
        cudaError_t e = scan_u32(scratch, tiles, scratch + tiles, total_dev, st);
        if (e != cudaSuccess) return e;
        vb_scan_add_kernel<<<(unsigned)tiles, SC_THREADS, 0, st>>>(data, n, scratch);
        ++vb_launch_counter;
    } else if (total_dev) {
        cudaError_t e = cudaMemcpyAsync(total_dev, scratch, sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, uint32 payload) pairs, 8 bits per pass; KeyT = uint32 (block
// labels) or uint64 (order-preserving image of a double, for the median)
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_BINS = 256;
constexpr int RS_WARPS = RS_THREADS / 32, RS_WARP_ELEMS = RS_TILE / RS_WARPS;

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
vb_rsort_hist_kernel(const KeyT *__restrict__ keys, long long n, int shift, unsigned *__restrict__ counts,
                     long long ntiles)
{
    __shared__ unsigned h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(long long)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// Stable scatter of one 2048-element tile.  Warp w owns elements [256 w, 256 (w+1)) of the tile and ranks
// them against its OWN digit counters (8 rounds of 32 consecutive elements, __match_any_sync inside a
// round), so the ranking loop needs no CTA-wide barrier; one barrier pair turns the per-warp counts into
// exclusive bases (digit d = thread d: prefix over the 8 warps).  The tile is then re-ordered IN SHARED
// MEMORY (position = start of its digit inside the tile + rank) and written out in tile order, so the
// elements of one digit leave as one contiguous run: full 32-byte sectors instead of one sector per
// element (ncu, 20M points: 30 sectors per store request before, profiles/r01_ncu_blocks_summary.json).
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
vb_rsort_scatter_kernel(const KeyT *__restrict__ keys, const unsigned *__restrict__ vals, long long n, int shift,
                        const unsigned *__restrict__ offsets, long long ntiles, KeyT *__restrict__ keys_out,
                        unsigned *__restrict__ vals_out)
{
    __shared__ unsigned cnt[RS_WARPS][RS_BINS];  // per-warp digit counts -> per-warp base inside the tile
    __shared__ unsigned dig_delta[RS_BINS];      // global offset of the digit minus its tile start
    __shared__ KeyT s_key[RS_TILE];
    __shared__ unsigned s_val[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const long long tile_base = (long long)blockIdx.x * RS_TILE;
    const long long wbase = tile_base + warp * RS_WARP_ELEMS;
    const int tile_n = (int)((n - tile_base < RS_TILE) ? (n - tile_base) : RS_TILE);
    KeyT key[RS_ITEMS];
    unsigned val[RS_ITEMS], rank[RS_ITEMS], digit[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys[i] : (KeyT)0;
        val[r] = ok ? vals[i] : 0u;
        digit[r] = ok ? ((unsigned)(key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, digit[r]);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));
        const unsigned seen = ok ? cnt[warp][digit[r]] : 0u;
        __syncwarp();
        if (ok && before == 0) cnt[warp][digit[r]] = seen + __popc(peers);
        __syncwarp();
        rank[r] = seen + before;
    }
    __syncthreads();
    // digit d = thread d: total of the digit in this tile, exclusive scan over digits, per-warp bases
    {
        unsigned total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) total += cnt[w][threadIdx.x];
        unsigned dummy;
        const unsigned start = cta_exclusive_scan(total, &dummy);
        dig_delta[threadIdx.x] = offsets[(long long)threadIdx.x * ntiles + blockIdx.x] - start;
        unsigned run = start;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const unsigned t = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        if (digit[r] < 256u) {
            const unsigned pos = cnt[warp][digit[r]] + rank[r];
            s_key[pos] = key[r];
            s_val[pos] = val[r];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tile_n; t += RS_THREADS) {
        const KeyT k = s_key[t];
        const unsigned pos = t + dig_delta[(unsigned)(k >> shift) & 255u];
        keys_out[pos] = k;
        vals_out[pos] = s_val[t];
    }
}

struct SortBuffers {
    void *keys[2];      // n keys each (allocated for uint64)
    unsigned *vals[2];
    unsigned *counts;   // 256 * ntiles
    unsigned *scratch;  // scan scratch
};

// sorts on bits [0, nbits) rounded up to whole bytes; result ends in buffers[*which]
template <typename KeyT>
cudaError_t radix_sort_pairs(SortBuffers &b, long long n, int nbits, int *which, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const long long ntiles = (n + RS_TILE - 1) / RS_TILE;
    const int passes = (nbits + 7) / 8;
    int cur = *which;
    for (int p = 0; p < passes; ++p) {
        vb_rsort_hist_kernel<KeyT><<<(unsigned)ntiles, RS_THREADS, 0, st>>>((const KeyT *)b.keys[cur], n, 8 * p, b.counts, ntiles);
        ++vb_launch_counter;
        cudaError_t e = scan_u32(b.counts, (long long)RS_BINS * ntiles, b.scratch, nullptr, st);
        if (e != cudaSuccess) return e;
        vb_rsort_scatter_kernel<KeyT><<<(unsigned)ntiles, RS_THREADS, 0, st>>>((const KeyT *)b.keys[cur], b.vals[cur], n, 8 * p,
                                                                              b.counts, ntiles, (KeyT *)b.keys[cur ^ 1],
                                                                              b.vals[cur ^ 1]);
        ++vb_launch_counter;
        cur ^= 1;
    }
    *which = cur;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// block labels
// ------------------------------------------------------------------------------------------
// index of the nearest of nc ascending, evenly spaced centres; exact ties go to the lower index (the
// reference's KD-tree breaks ties by traversal order, which is not a property of the data)
__device__ __forceinline__ int nearest_centre(double x, const double *__restrict__ c, int nc)
{
    if (nc <= 1) return 0;
    const double step = (c[nc - 1] - c[0]) / (double)(nc - 1);
    double t = (x - c[0]) / step + 0.5;
    int j = (t > 0.0) ? ((t < (double)nc) ? (int)t : nc - 1) : 0;  // NaN -> 0
    while (j > 0 && fabs(x - c[j - 1]) <= fabs(x - c[j])) --j;
    while (j < nc - 1 && fabs(x - c[j + 1]) < fabs(x - c[j])) ++j;
    return j;
}

__global__ void __launch_bounds__(256)
vb_block_label_kernel(const double *__restrict__ east, const double *__restrict__ north, long long n,
                      const double *__restrict__ ce, int n_east, const double *__restrict__ cn, int n_north,
                      unsigned *__restrict__ keys, unsigned *__restrict__ vals, long long *__restrict__ labels)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int je = nearest_centre(east[i], ce, n_east);
        const int jn = nearest_centre(north[i], cn, n_north);
        const long long label = (long long)jn * n_east + je;  // raveled C order of the (n_north, n_east) centre grid
        keys[i] = (unsigned)label;  // n_east * n_north < 2^31 (checked by the caller)
        vals[i] = (unsigned)i;
        if (labels) labels[i] = label;
    }
}

__global__ void __launch_bounds__(256)
vb_segment_flag_kernel(const unsigned *__restrict__ sorted_keys, long long n, unsigned *__restrict__ flags)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1u : 0u;
}

// flags_scanned[i] = index of the segment that starts at i (valid where a head is)
__global__ void __launch_bounds__(256)
vb_segment_heads_kernel(const unsigned *__restrict__ sorted_keys, const unsigned *__restrict__ scanned, long long n,
                        unsigned *__restrict__ seg_start, long long *__restrict__ seg_label, long long nseg)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i == 0 || sorted_keys[i] != sorted_keys[i - 1])) {
        seg_start[scanned[i]] = (unsigned)i;
        seg_label[scanned[i]] = (long long)sorted_keys[i];
    }
    if (i == 0) seg_start[nseg] = (unsigned)n;
}

// ------------------------------------------------------------------------------------------
// aggregation: one warp per non-empty block
// ------------------------------------------------------------------------------------------
struct Kahan {
    double s = 0.0, c = 0.0;
    __device__ __forceinline__ void add(double x)
    {
        const double y = x - c;
        const double t = s + y;
        c = (t - s) - y;
        s = t;
    }
    __device__ __forceinline__ double value() const { return s - c; }
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// GROUP = 32: one warp per block (the usual case, a handful of points per block);
// GROUP = 256: one CTA per block, for block grids so coarse that a block holds thousands of points.
// Both reduce in a fixed order (lane-strided compensated partial sums, shuffle tree, then the 8 warp
// totals in warp order), so results do not depend on scheduling.
template <int GROUP>
__device__ __forceinline__ double group_sum(double v, double *sh)
{
    v = warp_sum(v);
    if (GROUP == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    return t;
}
template <int GROUP>
__device__ __forceinline__ double group_minmax(double v, bool is_min, double *sh)
{
    v = is_min ? warp_min(v) : warp_max(v);
    if (GROUP == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = is_min ? fmin(t, sh[w]) : fmax(t, sh[w]);
    return t;
}

template <int GROUP>
__global__ void __launch_bounds__(256)
vb_block_aggregate_kernel(const unsigned *__restrict__ perm, const unsigned *__restrict__ seg_start, long long nseg,
                          const double *__restrict__ v, const double *__restrict__ w, int op, double *__restrict__ out)
{
    __shared__ double sh[8];
    const long long seg = GROUP == 32 ? (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (long long)blockIdx.x;
    const int lane = GROUP == 32 ? (threadIdx.x & 31) : threadIdx.x;
    if (seg >= nseg) return;  // whole warps (GROUP 32) or whole CTAs (GROUP 256) leave together
    const unsigned a = seg_start[seg], b = seg_start[seg + 1];
    const double cnt = (double)(b - a);
    double res = 0.0;
    if (op == VB_REDUCE_COUNT) {
        res = cnt;
    } else if (op == VB_REDUCE_MIN || op == VB_REDUCE_MAX) {
        const bool is_min = op == VB_REDUCE_MIN;
        double m = is_min ? 1.0 / 0.0 : -1.0 / 0.0;
        double nan = 0.0;
        for (unsigned k = a + lane; k < b; k += GROUP) {
            const double x = v[perm[k]];
            if (x != x) nan = 1.0;
            m = is_min ? fmin(m, x) : fmax(m, x);
        }
        m = group_minmax<GROUP>(m, is_min, sh);
        nan = group_minmax<GROUP>(nan, false, sh);
        res = nan > 0.0 ? (0.0 / 0.0) : m;  // numpy's min/max propagate NaN
    } else if (op == VB_REDUCE_MEDIAN) {
        // perm is sorted by (block, value) for this op
        const unsigned len = b - a;
        const double lo = v[perm[a + (len - 1) / 2]], hi = v[perm[a + len / 2]];
        const double last = v[perm[b - 1]];  // NaNs sort last: numpy's median is NaN when any value is
        res = (last != last) ? last : ((len & 1u) ? lo : 0.5 * (lo + hi));
    } else {
        // sums: SUM, MEAN, WMEAN, VAR (ddof = 1), WVAR, INV_SUM_W
        const bool weighted = (op == VB_REDUCE_WMEAN || op == VB_REDUCE_WVAR);
        Kahan sv, sw;
        for (unsigned k = a + lane; k < b; k += GROUP) {
            const unsigned i = perm[k];
            if (op == VB_REDUCE_INV_SUM_W) {
                sw.add(w[i]);
            } else if (weighted) {
                const double wi = w[i];
                sv.add(wi * v[i]);
                sw.add(wi);
            } else {
                sv.add(v[i]);
            }
        }
        const double tv = group_sum<GROUP>(sv.value(), sh);
        const double tw = group_sum<GROUP>(sw.value(), sh);
        if (op == VB_REDUCE_SUM) res = tv;
        else if (op == VB_REDUCE_INV_SUM_W) res = 1.0 / tw;
        else {
            const double mean = weighted ? tv / tw : tv / cnt;
            if (op == VB_REDUCE_MEAN || op == VB_REDUCE_WMEAN) {
                res = mean;
            } else {
                Kahan s2;
                for (unsigned k = a + lane; k < b; k += GROUP) {
                    const unsigned i = perm[k];
                    const double dlt = v[i] - mean;
                    s2.add(weighted ? w[i] * dlt * dlt : dlt * dlt);
                }
                const double t2 = group_sum<GROUP>(s2.value(), sh);
                // pandas "var"/"std": ddof = 1 (NaN for single points); the numpy callables np.var / np.std, which
                // pandas >= 3 applies as they are: ddof = 0
                const bool population = (op == VB_REDUCE_VAR0 || op == VB_REDUCE_STD0);
                res = weighted ? t2 / tw : t2 / (population ? cnt : cnt - 1.0);
                if (op == VB_REDUCE_STD || op == VB_REDUCE_STD0) res = sqrt(res);
            }
        }
    }
    if (lane == 0) out[seg] = res;
}

// order-preserving map double -> uint64 (NaN sorts after +inf)
__global__ void __launch_bounds__(256)
vb_value_keys_kernel(const double *__restrict__ v, long long n, uint64_t *__restrict__ keys, unsigned *__restrict__ vals)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = v[i];
    uint64_t u = (uint64_t)__double_as_longlong(x);
    if (x != x) u = 0xffffffffffffffffull;
    else u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    keys[i] = u;
    vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256)
vb_label_keys_kernel(const long long *__restrict__ labels, const unsigned *__restrict__ vals, long long n,
                     unsigned *__restrict__ keys)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (unsigned)labels[vals[i]];
}

// ------------------------------------------------------------------------------------------
// distance mask
// ------------------------------------------------------------------------------------------
struct CellGrid {
    double x0, y0, inv_pitch;
    int nx, ny;
};

__device__ __forceinline__ int cell_coord(double x, double x0, double inv_pitch, int n)
{
    const double t = (x - x0) * inv_pitch;
    return (t > 0.0) ? ((t < (double)n) ? (int)t : n - 1) : 0;
}

__global__ void __launch_bounds__(256)
vb_cell_count_kernel(const double *__restrict__ e, const double *__restrict__ n_, long long n, CellGrid g,
                     unsigned *__restrict__ cell_of, unsigned *__restrict__ counts)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned c = (unsigned)cell_coord(n_[i], g.y0, g.inv_pitch, g.ny) * (unsigned)g.nx +
                           (unsigned)cell_coord(e[i], g.x0, g.inv_pitch, g.nx);
        cell_of[i] = c;
        atomicAdd(&counts[c], 1u);
    }
}

// cursor starts as a copy of the cell offsets; the order inside a cell does not matter for a minimum
__global__ void __launch_bounds__(256)
vb_cell_fill_kernel(const double *__restrict__ e, const double *__restrict__ n_, long long n,
                    const unsigned *__restrict__ cell_of, unsigned *__restrict__ cursor,
                    double2 *__restrict__ sorted_xy)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned pos = atomicAdd(&cursor[cell_of[i]], 1u);
        sorted_xy[pos] = make_double2(e[i], n_[i]);
    }
}

__global__ void __launch_bounds__(256)
vb_distance_mask_kernel(const double *__restrict__ qe, const double *__restrict__ qn, long long nq, CellGrid g,
                        const unsigned *__restrict__ cell_start, const double2 *__restrict__ sorted_xy,
                        double maxdist, unsigned char *__restrict__ mask)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long)gridDim.x * blockDim.x) {
        const double x = qe[i], y = qn[i];
        bool hit = false;
        // a point within maxdist lies inside the bounding box of the data grown by maxdist
        const double tx = (x - g.x0) * g.inv_pitch, ty = (y - g.y0) * g.inv_pitch;
        if (tx >= -1.0 && ty >= -1.0 && tx <= (double)g.nx + 1.0 && ty <= (double)g.ny + 1.0) {
            const int cx = cell_coord(x, g.x0, g.inv_pitch, g.nx), cy = cell_coord(y, g.y0, g.inv_pitch, g.ny);
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1) && !hit; ++yy) {
                const int xa = max(cx - 1, 0), xb = min(cx + 1, g.nx - 1);
                // the three cells of a row are contiguous in the cell-sorted point array
                const unsigned a = cell_start[yy * g.nx + xa], b = cell_start[yy * g.nx + xb + 1];
                for (unsigned k = a; k < b; ++k) {
                    const double2 p = sorted_xy[k];
                    const double dx = x - p.x, dy = y - p.y;
                    // unfused, like the reference's sqrt(sum of squares): the comparison is exact at ties
                    const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    if (d <= maxdist) {
                        hit = true;
                        break;
                    }
                }
            }
        }
        mask[i] = hit ? 1 : 0;
    }
}

int grid_for(long long n)
{
    long long g = (n + 255) / 256;
    const long long cap = 148LL * 16;  // 16 resident CTAs of 256 threads per SM cover the grid-stride loops
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// order-free ("streaming") block reduction: ONE pass over the points, no grouping
// ------------------------------------------------------------------------------------------
// The sort-based path above moves ~68 B per point to give every block its points in pandas' order (needed by the
// median, by exact reproduction of pandas' summation order, and shared by all columns).  Sums, means, counts and
// weighted means do not need an order: each point is read once (24 B: easting, northing, value), labelled, and added
// into dense per-block accumulators with fire-and-forget atomics (RED.ADD.F64 / RED.ADD.U32, resolved in L2: the
// 224 x 224 accumulators of config 3 are 1 MB), then the non-empty blocks are compacted in label order.  The labels
// are the bit-exact ones of vb_block_label_kernel; the sums differ from the ordered path only by the rounding order
// of the additions (and may differ in the last bits from run to run).
__global__ void __launch_bounds__(256)
vb_block_stream_kernel(const double *__restrict__ east, const double *__restrict__ north, long long n,
                       const double *__restrict__ ce, int n_east, const double *__restrict__ cn, int n_north,
                       const double *__restrict__ values, int ncol, const double *__restrict__ weights,
                       long long nblk, unsigned *__restrict__ cnt, double *__restrict__ sums, double *__restrict__ sumw)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int je = nearest_centre(__ldcs(east + i), ce, n_east);
        const int jn = nearest_centre(__ldcs(north + i), cn, n_north);
        const long long label = (long long)jn * n_east + je;
        atomicAdd(cnt + label, 1u);
        double w = 1.0;
        if (weights) {
            w = __ldcs(weights + i);
            atomicAdd(sumw + label, w);
        }
        for (int c = 0; c < ncol; ++c) {
            const double v = __ldcs(values + (long long)c * n + i);
            atomicAdd(sums + (long long)c * nblk + label, weights ? w * v : v);
        }
    }
}

__global__ void __launch_bounds__(256)
vb_block_stream_flag_kernel(const unsigned *__restrict__ cnt, long long nblk, unsigned *__restrict__ flags)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk) flags[b] = cnt[b] ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
vb_block_stream_emit_kernel(const unsigned *__restrict__ cnt, const unsigned *__restrict__ slot, long long nblk,
                            const double *__restrict__ sums, const double *__restrict__ sumw, int ncol, int op,
                            long long capacity, long long *__restrict__ ids, double *__restrict__ out)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const unsigned c0 = cnt[b];
    if (!c0) return;
    const long long j = slot[b];
    if (j >= capacity) return;
    ids[j] = b;
    for (int c = 0; c < ncol; ++c) {
        const double s = sums[(long long)c * nblk + b];
        double r;
        if (op == VB_REDUCE_SUM) r = s;
        else if (op == VB_REDUCE_COUNT) r = (double)c0;
        else if (op == VB_REDUCE_WMEAN) r = s / sumw[b];
        else r = s / (double)c0;
        out[(long long)c * capacity + j] = r;
    }
}

// ------------------------------------------------------------------------------------------
// host entry points used by vb_api.cu
// ------------------------------------------------------------------------------------------
cudaError_t vb_block_stream(const double *east, const double *north, long long n, const double *ce, int n_east,
                            const double *cn, int n_north, const double *values, int ncol, const double *weights,
                            int op, unsigned *cnt, double *sums, double *sumw, unsigned *flags, unsigned *scratch,
                            unsigned *nseg_dev, long long capacity, long long *ids, double *out, cudaStream_t st)
{
    const long long nblk = (long long)n_east * n_north;
    cudaError_t e;
    if ((e = cudaMemsetAsync(cnt, 0, sizeof(unsigned) * nblk, st)) != cudaSuccess) return e;
    if (ncol > 0 && (e = cudaMemsetAsync(sums, 0, sizeof(double) * nblk * ncol, st)) != cudaSuccess) return e;
    if (weights && (e = cudaMemsetAsync(sumw, 0, sizeof(double) * nblk, st)) != cudaSuccess) return e;
    vb_block_stream_kernel<<<grid_for(n), 256, 0, st>>>(east, north, n, ce, n_east, cn, n_north, values, ncol, weights,
                                                        nblk, cnt, sums, sumw);
    const unsigned blocks = (unsigned)((nblk + 255) / 256);
    vb_block_stream_flag_kernel<<<blocks, 256, 0, st>>>(cnt, nblk, flags);
    vb_launch_counter += 2;
    if ((e = scan_u32(flags, nblk, scratch, nseg_dev, st)) != cudaSuccess) return e;
    vb_block_stream_emit_kernel<<<blocks, 256, 0, st>>>(cnt, flags, nblk, sums, sumw, ncol, op, capacity, ids, out);
    ++vb_launch_counter;
    return cudaGetLastError();
}

long long vb_block_scan_scratch_elems(long long n) { return scan_scratch_elems(n); }

cudaError_t vb_block_split(const double *east, const double *north, long long n, const double *ce, int n_east,
                           const double *cn, int n_north, VbBlockWorkspace &ws, long long *labels_dev,
                           cudaStream_t st)
{
    SortBuffers sb = {{ws.keys[0], ws.keys[1]}, {ws.vals[0], ws.vals[1]}, ws.counts, ws.scratch};
    vb_block_label_kernel<<<grid_for(n), 256, 0, st>>>(east, north, n, ce, n_east, cn, n_north, (unsigned *)sb.keys[0],
                                                       sb.vals[0], labels_dev);
    ++vb_launch_counter;
    int nbits = 1;
    while (nbits < 31 && (1LL << nbits) < (long long)n_east * n_north) ++nbits;
    int which = 0;
    cudaError_t e = radix_sort_pairs<unsigned>(sb, n, nbits, &which, st);
    if (e != cudaSuccess) return e;
    ws.sorted = which;
    ws.label_bits = nbits;
    // segment heads
    const unsigned blocks = (unsigned)((n + 255) / 256);
    vb_segment_flag_kernel<<<blocks, 256, 0, st>>>((const unsigned *)sb.keys[which], n, ws.flags);
    ++vb_launch_counter;
    return scan_u32(ws.flags, n, ws.scratch, ws.nseg_dev, st);
}

cudaError_t vb_block_heads(long long n, long long nseg, VbBlockWorkspace &ws, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((n + 255) / 256);
    vb_segment_heads_kernel<<<blocks, 256, 0, st>>>((const unsigned *)ws.keys[ws.sorted], ws.flags, n, ws.seg_start, ws.seg_label, nseg);
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_block_aggregate(const double *values, const double *weights, long long n, long long nseg, int op,
                               const long long *labels_dev, VbBlockWorkspace &ws, double *out, cudaStream_t st)
{
    const unsigned *perm = ws.vals[ws.sorted];
    if (op == VB_REDUCE_MEDIAN) {
        // (block, value) order: value pass first, stable block pass second, in the spare buffers
        SortBuffers sb = {{ws.mkeys[0], ws.mkeys[1]}, {ws.mvals[0], ws.mvals[1]}, ws.counts, ws.scratch};
        const unsigned blocks = (unsigned)((n + 255) / 256);
        vb_value_keys_kernel<<<blocks, 256, 0, st>>>(values, n, (uint64_t *)sb.keys[0], sb.vals[0]);
        ++vb_launch_counter;
        int which = 0;
        cudaError_t e = radix_sort_pairs<uint64_t>(sb, n, 64, &which, st);
        if (e != cudaSuccess) return e;
        // the value keys are spent: the same buffer now takes the 32-bit block labels in value order
        vb_label_keys_kernel<<<blocks, 256, 0, st>>>(labels_dev, sb.vals[which], n, (unsigned *)sb.keys[which]);
        ++vb_launch_counter;
        e = radix_sort_pairs<unsigned>(sb, n, ws.label_bits, &which, st);
        if (e != cudaSuccess) return e;
        perm = sb.vals[which];
    }
    if (n / nseg >= 2048) {
        vb_block_aggregate_kernel<256><<<(unsigned)nseg, 256, 0, st>>>(perm, ws.seg_start, nseg, values, weights, op, out);
    } else {
        const long long threads = nseg * 32;
        vb_block_aggregate_kernel<32><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(perm, ws.seg_start, nseg, values,
                                                                                        weights, op, out);
    }
    ++vb_launch_counter;
    return cudaGetLastError();
}

cudaError_t vb_distance_mask(const double *data_e, const double *data_n, long long n, double x0, double y0,
                             double pitch, int nx, int ny, unsigned *cell_of, unsigned *cell_start /* nx*ny + 1 */,
                             unsigned *cursor, unsigned *scratch, double *sorted_xy, const double *qe,
                             const double *qn, long long nq, double maxdist, unsigned char *mask, cudaStream_t st)
{
    CellGrid g = {x0, y0, 1.0 / pitch, nx, ny};
    const long long ncell = (long long)nx * ny;
    cudaError_t e;
    if ((e = cudaMemsetAsync(cell_start, 0, sizeof(unsigned) * (ncell + 1), st)) != cudaSuccess) return e;
    vb_cell_count_kernel<<<grid_for(n), 256, 0, st>>>(data_e, data_n, n, g, cell_of, cell_start);
    ++vb_launch_counter;
    if ((e = scan_u32(cell_start, ncell + 1, scratch, nullptr, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(cursor, cell_start, sizeof(unsigned) * ncell, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    vb_cell_fill_kernel<<<grid_for(n), 256, 0, st>>>(data_e, data_n, n, cell_of, cursor, reinterpret_cast<double2 *>(sorted_xy));
    ++vb_launch_counter;
    vb_distance_mask_kernel<<<grid_for(nq), 256, 0, st>>>(qe, qn, nq, g, cell_start, reinterpret_cast<const double2 *>(sorted_xy),
                                                         maxdist, mask);
    ++vb_launch_counter;
    return cudaGetLastError();
}
