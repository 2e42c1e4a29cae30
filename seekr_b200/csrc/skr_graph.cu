// Similarity graph of a Pearson matrix: the numeric front half of seekr/kmer_leiden.py (lines 91-107).
//
//   ld_sim[ld_sim < pearsoncutoff] = 0 ; np.fill_diagonal(ld_sim, 0)            -> skr_sim_threshold (in place)
//   (df.values > 0) adjacency + df.values[df.values > 0] weights (row-major)   -> skr_sim_edge_offsets + skr_sim_edge_fill
//
// An entry (i, j) of the thresholded matrix is positive exactly when the ORIGINAL value x satisfies
// !(x < cutoff) && x > 0 && i != j (NaN < cutoff is false, so NaN survives the threshold, and NaN > 0 is false, so it
// is no edge), hence the edge kernels work on the untouched r matrix and never need the dense thresholded copy.
// The comparison is made in the matrix's own type with the cutoff converted to it (numpy: a Python scalar is weak).
//
// All three passes are HBM-bound streams over the m x n matrix with 16-byte loads, several issued before the first
// is used.  Edges come out in row-major order (np.nonzero order): counts per (row, column slice) -> one exclusive
// scan (two small launches) -> ordered compaction inside each slice by the warp that owns it (shuffle scan of per-lane counts, no
// CTA-wide barrier; steps without an edge cost one ballot).
#include <cuda_runtime.h>

#include "skr_common.h"

namespace {

constexpr int kThreads = 256;

template <typename T>
struct Vec;
template <>
struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, T (&v)[Vec<T>::N]) {
    typename Vec<T>::type q = *reinterpret_cast<const typename Vec<T>::type*>(p);
    const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int u = 0; u < Vec<T>::N; ++u) v[u] = e[u];
}

__device__ __forceinline__ void store_vec(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

template <typename T>
__device__ __forceinline__ bool is_edge(T x, T cut, long long i, long long j, long long jmin) {
    return !(x < cut) && x > T(0) && i != j && j >= jmin;
}

// !(x < cut) && x > 0 as ONE comparison x >= thr: for cut > 0 it is x >= cut, otherwise (cut <= 0 or NaN) it is x > 0,
// i.e. x >= the smallest positive subnormal (comparisons are IEEE here: no flush-to-zero).  NaN fails both forms.
template <typename T>
__host__ __device__ inline T edge_threshold(T cut) {
    return cut > T(0) ? cut : (sizeof(T) == 4 ? T(1.40129846432481707e-45) : T(4.9406564584124654e-324));
}

// Row i, elements [j0, j0 + N): vector load when the whole vector is inside the row and aligned, else scalars
// (out-of-range lanes read as 0, which is never an edge).
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_row(const T* row, long long j0, long long n, T (&v)[Vec<T>::N]) {
    constexpr int N = Vec<T>::N;
    if (ALIGNED && j0 + N <= n) {
        load_vec<T>(row + j0, v);
    } else {
#pragma unroll
        for (int u = 0; u < N; ++u) v[u] = (j0 + u < n) ? row[j0 + u] : T(0);
    }
}

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) sim_threshold_kernel(T* c, long long m, long long n, long long ld, long long row0, T cut,
                                                                 int zero_diag) {
    constexpr int N = Vec<T>::N;
    for (long long i = blockIdx.x; i < m; i += gridDim.x) {
        T* row = c + i * ld;
        for (long long j0 = (long long)threadIdx.x * N; j0 < n; j0 += (long long)kThreads * N) {
            T v[N];
            load_row<T, ALIGNED>(row, j0, n, v);
            bool touched = false;
#pragma unroll
            for (int u = 0; u < N; ++u) {
                const bool z = (v[u] < cut) || (zero_diag && j0 + u == row0 + i);
                if (z) v[u] = T(0);
                touched |= z;
            }
            if (!touched) continue;  // rows of a dense similarity matrix above the cutoff are left unwritten
            if (ALIGNED && j0 + N <= n) {
                store_vec(row + j0, v);
            } else {
#pragma unroll
                for (int u = 0; u < N; ++u)
                    if (j0 + u < n) row[j0 + u] = v[u];
            }
        }
    }
}

__device__ __forceinline__ int warp_sum(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Work item = (row, slice): a row is cut into kSlices column slices of `width` elements (a multiple of the
// 32 * N elements a warp loads at once), and every warp of the grid takes items round-robin on its own -- no CTA-wide
// barrier anywhere.  Item q = row * kSlices + slice counts into counts[q]; the scan over all items gives every
// item its place in the edge arrays, rows in order and slices in column order inside a row.
constexpr int kSlices = SKR_SIM_SLICES;

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) sim_edge_count_kernel(const T* __restrict__ c, long long m, long long n,
                                                                  long long ld, long long row0, long long width, T cut,
                                                                  int upper_only, long long* __restrict__ counts) {
    constexpr int N = Vec<T>::N;
    constexpr long long STEP = 32ll * N;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * kThreads) >> 5;
    const T thr = edge_threshold<T>(cut);
    for (long long q0 = warp0; q0 < m * kSlices; q0 += nwarps) {
        // the grid holds a multiple of kSlices warps, so q0 % kSlices is fixed per warp: rotate the slice by the row
        // (with upper_only the left slices of a row are empty -- every warp gets its share of them)
        const long long i = q0 / kSlices, sl = (q0 + i) % kSlices, q = i * kSlices + sl;
        const long long g = row0 + i;  // row of the whole matrix: diagonal column, first column of the upper half
        const long long jmin = upper_only ? g + 1 : 0;
        const long long jend = (sl + 1) * width < n ? (sl + 1) * width : n;
        long long jbeg = sl * width;
        if (jbeg < jmin) jbeg = jmin / STEP * STEP;  // whole steps before jmin hold no edge (jmin >= jbeg: same grid)
        int cnt = 0;
        const T* row = c + i * ld;
        for (long long j0 = jbeg + (long long)lane * N; j0 - (long long)lane * N < jend; j0 += 4 * STEP) {
            T v[4][N];
#pragma unroll
            for (int w = 0; w < 4; ++w) load_row<T, ALIGNED>(row, j0 + w * STEP, jend, v[w]);
            const long long base = j0 - (long long)lane * N;
            if (base >= jmin && (g < base || g >= base + 4 * STEP)) {
                // no diagonal element and nothing left of jmin in these four steps: one comparison per element
                // (lanes beyond jend hold zeros)
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int u = 0; u < N; ++u) cnt += v[w][u] >= thr;
            } else {
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int u = 0; u < N; ++u) cnt += is_edge<T>(v[w][u], cut, g, j0 + w * STEP + u, jmin);
            }
        }
        cnt = warp_sum(cnt);
        if (lane == 0) counts[q] = cnt;
    }
}

// offsets[0] = 0, offsets[i+1] = offsets[i] + counts[i]; counts and offsets + 1 may alias (in-place inclusive scan).
// Two small launches over kScanCtas chunks of consecutive counts: chunk sums, then every CTA adds up the sums of
// the chunks before its own and scans its chunk in coalesced rounds of 256 (the single-CTA version took 0.29 ms
// for the 320 000 counts of a 40 000-row matrix, a quarter of the pass that produces them).
constexpr int kScanCtas = 296;

__device__ __forceinline__ long long block_sum_256(long long x, long long* warp_tot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    long long t = 0;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) t += warp_tot[q];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kThreads) sim_chunk_sum_kernel(const long long* __restrict__ counts, long long total,
                                                                 long long* __restrict__ partial) {
    __shared__ long long warp_tot[kThreads / 32];
    const long long chunk = (total + gridDim.x - 1) / gridDim.x;
    const long long beg = blockIdx.x * chunk < total ? blockIdx.x * chunk : total;
    const long long end = beg + chunk < total ? beg + chunk : total;
    long long sum = 0;
    for (long long k = beg + threadIdx.x; k < end; k += kThreads) sum += counts[k];
    sum = block_sum_256(sum, warp_tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = sum;
}

__global__ void __launch_bounds__(kThreads) sim_chunk_scan_kernel(const long long* counts, long long total,
                                                                  const long long* __restrict__ partial,
                                                                  long long* offsets) {
    __shared__ long long warp_tot[kThreads / 32];
    const long long chunk = (total + gridDim.x - 1) / gridDim.x;
    const long long beg = blockIdx.x * chunk < total ? blockIdx.x * chunk : total;
    const long long end = beg + chunk < total ? beg + chunk : total;
    long long before = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kThreads) before += partial[b];
    long long carry = block_sum_256(before, warp_tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[0] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = beg; base < end; base += kThreads) {
        const long long k = base + threadIdx.x;
        const long long x = k < end ? counts[k] : 0;
        long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long add = carry, round = 0;
#pragma unroll
        for (int q = 0; q < kThreads / 32; ++q) {
            const long long t = warp_tot[q];
            if (q < warp) add += t;
            round += t;
        }
        if (k < end) offsets[k + 1] = incl + add;
        carry += round;
        __syncthreads();
    }
}

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(kThreads) sim_edge_fill_kernel(const T* __restrict__ c, long long m, long long n,
                                                                 long long ld, long long row0, long long width, T cut,
                                                                 int upper_only, const long long* __restrict__ offsets,
                                                                 int* __restrict__ src, int* __restrict__ dst,
                                                                 T* __restrict__ weight) {
    constexpr int N = Vec<T>::N;
    constexpr long long STEP = 32ll * N;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * kThreads) >> 5;
    const T thr = edge_threshold<T>(cut);
    __shared__ int stage_col[kThreads / 32][2 * 32 * N];
    __shared__ T stage_w[kThreads / 32][2 * 32 * N];
    int* const st_col = stage_col[threadIdx.x >> 5];
    T* const st_w = stage_w[threadIdx.x >> 5];
    for (long long q0 = warp0; q0 < m * kSlices; q0 += nwarps) {
        const long long i = q0 / kSlices, sl = (q0 + i) % kSlices, q = i * kSlices + sl;  // as in the count kernel
        long long out = offsets[q];
        if (offsets[q + 1] == out) continue;  // uniform over the warp
        const long long g = row0 + i;  // row of the whole matrix: diagonal column, first column of the upper half
        const long long jmin = upper_only ? g + 1 : 0;
        const long long jend = (sl + 1) * width < n ? (sl + 1) * width : n;
        long long jbeg = sl * width;
        if (jbeg < jmin) jbeg = jmin / STEP * STEP;
        const T* row = c + i * ld;
        long long j0 = jbeg + (long long)lane * N;
        // column order inside a double step: the 32 a-vectors (j0 ...), then the 32 b-vectors (j0 + STEP ...);
        // the next double step is in flight while this one is compacted
        T a[N], b[N];
        load_row<T, ALIGNED>(row, j0, jend, a);
        load_row<T, ALIGNED>(row, j0 + STEP, jend, b);
        for (; j0 - (long long)lane * N < jend; j0 += 2 * STEP) {
            T na[N], nb[N];
            load_row<T, ALIGNED>(row, j0 + 2 * STEP, jend, na);
            load_row<T, ALIGNED>(row, j0 + 3 * STEP, jend, nb);
            const long long base = j0 - (long long)lane * N;
            unsigned ma = 0, mb = 0;
            if (base >= jmin && (g < base || g >= base + 2 * STEP)) {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    ma |= (unsigned)(a[u] >= thr) << u;
                    mb |= (unsigned)(b[u] >= thr) << u;
                }
            } else {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    ma |= (unsigned)is_edge<T>(a[u], cut, g, j0 + u, jmin) << u;
                    mb |= (unsigned)is_edge<T>(b[u], cut, g, j0 + STEP + u, jmin) << u;
                }
            }
            if (__ballot_sync(0xffffffffu, (ma | mb) != 0) != 0) {  // sparse graphs: most steps hold no edge
                // packed warp scan: low half counts the a-parts, high half the b-parts (each at most 32 * N)
                const unsigned own = (unsigned)__popc(ma) | ((unsigned)__popc(mb) << 16);
                unsigned incl = own;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
                const unsigned excl = incl - own;
                // ranks inside the double step: a-parts first, then b-parts.  The edges are scattered into the warp's
                // shared staging rows by rank and written out by consecutive lanes: full 128-byte store
                // instructions instead of 4-byte stores at per-lane offsets (8 half-used sectors each).
                const int ta = (int)(total & 0xffffu), tt = ta + (int)(total >> 16);
                int ra = (int)(excl & 0xffffu), rb = ta + (int)(excl >> 16);
#pragma unroll
                for (int u = 0; u < N; ++u)
                    if (ma >> u & 1) {
                        st_col[ra] = (int)(j0 + u);
                        st_w[ra] = a[u];
                        ++ra;
                    }
#pragma unroll
                for (int u = 0; u < N; ++u)
                    if (mb >> u & 1) {
                        st_col[rb] = (int)(j0 + STEP + u);
                        st_w[rb] = b[u];
                        ++rb;
                    }
                __syncwarp();
                for (int r = lane; r < tt; r += 32) {
                    if (src) src[out + r] = (int)g;
                    dst[out + r] = st_col[r];
                    weight[out + r] = st_w[r];
                }
                __syncwarp();
                out += tt;
            }
#pragma unroll
            for (int u = 0; u < N; ++u) {
                a[u] = na[u];
                b[u] = nb[u];
            }
        }
    }
}

// slice width: ceil(n / kSlices) rounded up to the 256 elements a warp covers with two float4 (four double2) loads
long long slice_width(long long n) {
    const long long w = (n + kSlices - 1) / kSlices;
    return (w + 255) / 256 * 256;
}

int row_grid(long long m) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride over rows
    return (int)(grid < m ? grid : m);
}

bool vec_aligned(const void* p, long long ld, int elem) {
    return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && ((ld * elem) % 16 == 0);
}

int check_matrix(const char* who, const void* d_c, int64_t m, int64_t n, int64_t ld) {
    if (!d_c) return skr::fail(SKR_ERR_ARG, "%s: null matrix", who);
    if (ld < n) return skr::fail(SKR_ERR_ARG, "%s: leading dimension smaller than n", who);
    if (m >= (1ll << 31) || n >= (1ll << 31)) return skr::fail(SKR_ERR_ARG, "%s: more than 2^31-1 rows or columns", who);
    return SKR_OK;
}

}  // namespace

extern "C" int skr_sim_threshold(void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0, double cutoff,
                                 int zero_diagonal, void* stream) {
    if (m <= 0 || n <= 0) return SKR_OK;
    if (int rc = check_matrix("skr_sim_threshold", d_c, m, n, ld)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = row_grid(m);
    const bool al = vec_aligned(d_c, ld, c_is_f64 ? 8 : 4);
    const long long M = m, N = n, LD = ld, R0 = row0;
    if (c_is_f64) {
        if (al) sim_threshold_kernel<double, true><<<grid, kThreads, 0, s>>>((double*)d_c, M, N, LD, R0, cutoff, zero_diagonal);
        else sim_threshold_kernel<double, false><<<grid, kThreads, 0, s>>>((double*)d_c, M, N, LD, R0, cutoff, zero_diagonal);
    } else {
        const float cut = (float)cutoff;
        if (al) sim_threshold_kernel<float, true><<<grid, kThreads, 0, s>>>((float*)d_c, M, N, LD, R0, cut, zero_diagonal);
        else sim_threshold_kernel<float, false><<<grid, kThreads, 0, s>>>((float*)d_c, M, N, LD, R0, cut, zero_diagonal);
    }
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_sim_edge_offsets(const void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0,
                                    double cutoff, int upper_only, int64_t* d_offsets, void* stream) {
    if (!d_offsets) return skr::fail(SKR_ERR_ARG, "skr_sim_edge_offsets: null offsets");
    cudaStream_t s = (cudaStream_t)stream;
    if (m <= 0 || n <= 0) {
        SKR_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, sizeof(int64_t) * (size_t)((m > 0 ? m : 0) * kSlices + 1), s));
        return SKR_OK;
    }
    if (int rc = check_matrix("skr_sim_edge_offsets", d_c, m, n, ld)) return rc;
    const int grid = row_grid(m);
    const bool al = vec_aligned(d_c, ld, c_is_f64 ? 8 : 4);
    const long long M = m, N = n, LD = ld, R0 = row0, W = slice_width(n);
    long long* counts = (long long*)d_offsets + 1;  // scanned in place
    if (c_is_f64) {
        if (al) sim_edge_count_kernel<double, true><<<grid, kThreads, 0, s>>>((const double*)d_c, M, N, LD, R0, W, cutoff, upper_only, counts);
        else sim_edge_count_kernel<double, false><<<grid, kThreads, 0, s>>>((const double*)d_c, M, N, LD, R0, W, cutoff, upper_only, counts);
    } else {
        const float cut = (float)cutoff;
        if (al) sim_edge_count_kernel<float, true><<<grid, kThreads, 0, s>>>((const float*)d_c, M, N, LD, R0, W, cut, upper_only, counts);
        else sim_edge_count_kernel<float, false><<<grid, kThreads, 0, s>>>((const float*)d_c, M, N, LD, R0, W, cut, upper_only, counts);
    }
    SKR_LAUNCH_CHECK();
    long long* partial = nullptr;  // stream-ordered scratch: concurrent calls on other streams get their own
    SKR_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(long long) * kScanCtas, s));
    sim_chunk_sum_kernel<<<kScanCtas, kThreads, 0, s>>>(counts, M * kSlices, partial);
    SKR_LAUNCH_CHECK();
    sim_chunk_scan_kernel<<<kScanCtas, kThreads, 0, s>>>(counts, M * kSlices, partial, (long long*)d_offsets);
    SKR_LAUNCH_CHECK();
    SKR_CUDA_CHECK(cudaFreeAsync(partial, s));
    return SKR_OK;
}

extern "C" int64_t skr_sim_slice_width(int64_t n) { return slice_width(n); }

// offsets[1 ..] holds the per-(row, slice) counts (e.g. written by the Pearson GEMM's epilogue): the scan alone
extern "C" int skr_sim_offsets_scan(int64_t* d_offsets, int64_t m, void* stream) {
    if (!d_offsets) return skr::fail(SKR_ERR_ARG, "skr_sim_offsets_scan: null offsets");
    if (m <= 0) return SKR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    long long* counts = (long long*)d_offsets + 1;
    long long* partial = nullptr;
    SKR_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(long long) * kScanCtas, s));
    sim_chunk_sum_kernel<<<kScanCtas, kThreads, 0, s>>>(counts, (long long)m * kSlices, partial);
    SKR_LAUNCH_CHECK();
    sim_chunk_scan_kernel<<<kScanCtas, kThreads, 0, s>>>(counts, (long long)m * kSlices, partial, (long long*)d_offsets);
    SKR_LAUNCH_CHECK();
    SKR_CUDA_CHECK(cudaFreeAsync(partial, s));
    return SKR_OK;
}

extern "C" int skr_sim_edge_fill(const void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0,
                                 double cutoff, int upper_only, const int64_t* d_offsets, int32_t* d_src,
                                 int32_t* d_dst, void* d_weight, void* stream) {
    if (m <= 0 || n <= 0) return SKR_OK;
    if (int rc = check_matrix("skr_sim_edge_fill", d_c, m, n, ld)) return rc;
    if (!d_offsets || !d_dst || !d_weight) return skr::fail(SKR_ERR_ARG, "skr_sim_edge_fill: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = row_grid(m);
    const bool al = vec_aligned(d_c, ld, c_is_f64 ? 8 : 4);
    const long long M = m, N = n, LD = ld, R0 = row0, W = slice_width(n);
    const long long* off = (const long long*)d_offsets;
    if (c_is_f64) {
        if (al) sim_edge_fill_kernel<double, true><<<grid, kThreads, 0, s>>>((const double*)d_c, M, N, LD, R0, W, cutoff, upper_only, off, d_src, d_dst, (double*)d_weight);
        else sim_edge_fill_kernel<double, false><<<grid, kThreads, 0, s>>>((const double*)d_c, M, N, LD, R0, W, cutoff, upper_only, off, d_src, d_dst, (double*)d_weight);
    } else {
        const float cut = (float)cutoff;
        if (al) sim_edge_fill_kernel<float, true><<<grid, kThreads, 0, s>>>((const float*)d_c, M, N, LD, R0, W, cut, upper_only, off, d_src, d_dst, (float*)d_weight);
        else sim_edge_fill_kernel<float, false><<<grid, kThreads, 0, s>>>((const float*)d_c, M, N, LD, R0, W, cut, upper_only, off, d_src, d_dst, (float*)d_weight);
    }
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}
