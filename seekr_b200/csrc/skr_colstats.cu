// K2: column statistics for mean=True / std=True and seekr_norm_vectors.
// Replaces np.mean / np.std(axis=0) at seekr/kmer_counts.py:168,174.
//
// Order-exact passes (skr_col_pass).  numpy reduces axis 0 of the C-ordered float32 matrix row
// after row into an fp32 accumulator, so the reference's vectors carry a specific rounding
// history (at 250k rows its std is ~1e-3 relative away from the binary64 value).  To reproduce
// them bit for bit each column is summed sequentially in row order with plain fp32 adds.  The
// dependent add chain is 4 cycles per row, about the time HBM needs to deliver the row anyway,
// provided the loads never stall it (no pass writes the matrix: the centred values are recomputed
// where needed, with the same single rounding): a CTA owns a strip of 32 columns; one producer thread
// streams [128 rows x 32 columns] boxes of the strip through a 8-deep shared-memory ring with
// TMA (cp.async.bulk.tensor + mbarrier), and one consumer warp (lane = column) walks the rows.
// The running sums enter and leave through d_acc, so row shards on several GPUs can be chained.
//
// Scalable passes (skr_col_partial_f64): row-parallel binary64 partial sums for one all-reduce.
#include <cuda.h>
#include <cuda_runtime.h>

#include "skr_common.h"
#include "skr_device.cuh"
#include "skr_tma.h"

namespace {

constexpr int kGroups = 1;                  // independent columns (add chains) per consumer lane (2 was slower: half the CTAs)
constexpr int kStripCols = 32 * kGroups;
constexpr int kTileRows = 256;
constexpr int kStages = 6;
constexpr int kTileBytes = kTileRows * kStripCols * 4;  // 16 KB

template <int KIND, bool kVecF64>
__global__ void __launch_bounds__(64) col_pass_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ a,
                                                      long long m, long long cols, long long ld, const void* vec,
                                                      const float* vec2, float* acc_io) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];
    float(*tiles)[kTileRows][kStripCols] = reinterpret_cast<float(*)[kTileRows][kStripCols]>(smem_raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long col0 = (long long)blockIdx.x * kStripCols;
    const long long ntiles = (m + kTileRows - 1) / kTileRows;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            skr::mbar_init(&full_bar[s], 1);
            skr::mbar_init(&empty_bar[s], 1);
        }
        skr::fence_mbar_init();
    }
    __syncthreads();

    if (warp == 1) {
        if (lane == 0) {
            skr::tma_prefetch_desc(&tmap);
            for (long long t = 0; t < ntiles; ++t) {
                const int s = (int)(t % kStages);
                const uint32_t ph = (uint32_t)((t / kStages) & 1);
                skr::mbar_wait(&empty_bar[s], ph ^ 1u);
                skr::mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
                skr::tma_load_2d(&tiles[s][0][0], &tmap, (int)col0, (int)(t * kTileRows), &full_bar[s]);
            }
        }
        return;
    }

    // consumer warp: lane owns columns col0 + lane + 32*g (g < kGroups).  The adds of one column form a
    // dependent fp32 chain (4 cycles each) and the warp issues in order, so a single chain leaves the issue
    // slot idle most of the time; kGroups independent chains per lane are interleaved to fill it.
    float acc[kGroups], vf[kGroups], v2[kGroups];
    double vd[kGroups];
    bool active[kGroups];
    const bool has_vec = vec != nullptr;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const long long col = col0 + lane + 32 * g;
        active[g] = col < cols;
        acc[g] = active[g] ? acc_io[col] : 0.0f;
        vf[g] = v2[g] = 0.0f;
        vd[g] = 0.0;
        if (active[g]) {
            if (has_vec) {
                if (kVecF64) vd[g] = reinterpret_cast<const double*>(vec)[col];
                else vf[g] = reinterpret_cast<const float*>(vec)[col];
            }
            if (KIND == SKR_COLPASS_SQDEV) v2[g] = vec2[col];
        }
    }
    // one IEEE operation per step, exactly the reference's sequence: counts -= mean (one rounding),
    // np.std: x - arrmean, square, sequential fp32 sum (kmer_counts.py:169,174; numpy _methods.py:_var).
    // value() is the per-element work that does not depend on the running sum.
    auto value = [&](float x, int g) -> float {
        float y = x;
        if (KIND != SKR_COLPASS_SUM && has_vec)
            y = kVecF64 ? __double2float_rn(__dsub_rn((double)x, vd[g])) : __fsub_rn(x, vf[g]);
        if (KIND == SKR_COLPASS_SQDEV) {
            const float d = __fsub_rn(y, v2[g]);
            y = __fmul_rn(d, d);
        }
        return y;
    };
    // Register window of kWin rows: right after row r has been added, its register is refilled with row
    // r + kWin of the same tile, so every dependent add (4 cycles) has an independent shared load next to it
    // in program order and the in-order warp never waits for a load (the refill is consumed kWin adds later).
    constexpr int kWin = 16;
    static_assert(kTileRows % kWin == 0, "tile rows must be a multiple of the window");
    for (long long t = 0; t < ntiles; ++t) {
        const int s = (int)(t % kStages);
        const uint32_t ph = (uint32_t)((t / kStages) & 1);
        skr::mbar_wait(&full_bar[s], ph);
        const int rows = (int)min((long long)kTileRows, m - t * kTileRows);
        if (rows == kTileRows) {
            float y[kWin][kGroups];
#pragma unroll
            for (int u = 0; u < kWin; ++u)
#pragma unroll
                for (int g = 0; g < kGroups; ++g) y[u][g] = value(tiles[s][u][lane + 32 * g], g);
            for (int r = 0; r < kTileRows - kWin; r += kWin) {
#pragma unroll
                for (int u = 0; u < kWin; ++u)
#pragma unroll
                    for (int g = 0; g < kGroups; ++g) {
                        acc[g] = __fadd_rn(acc[g], y[u][g]);
                        y[u][g] = value(tiles[s][r + kWin + u][lane + 32 * g], g);
                    }
            }
#pragma unroll
            for (int u = 0; u < kWin; ++u)
#pragma unroll
                for (int g = 0; g < kGroups; ++g) acc[g] = __fadd_rn(acc[g], y[u][g]);
        } else {
            for (int r = 0; r < rows; ++r)
#pragma unroll
                for (int g = 0; g < kGroups; ++g) acc[g] = __fadd_rn(acc[g], value(tiles[s][r][lane + 32 * g], g));
        }
        __syncwarp();
        if (lane == 0) skr::mbar_arrive(&empty_bar[s]);
    }
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
        if (active[g]) acc_io[col0 + lane + 32 * g] = acc[g];
}

// flag (optional): bit 0 set if any result is not finite, bit 1 if any result is <= 0
__device__ __forceinline__ void note_flag(float v, int* flag) {
    if (!flag) return;
    int f = 0;
    if (!(v - v == 0.0f)) f |= 1;
    if (!(v > 0.0f)) f |= 2;
    if (f) atomicOr(flag, f);
}

__global__ void col_finish_kernel(const float* acc, long long cols, long long rows, int take_sqrt, float* out, int* flag) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    // numpy divides the fp32 sum by the row count in binary64 and rounds to fp32 (_methods.py:_mean/_var)
    float v = __double2float_rn(__ddiv_rn((double)acc[j], (double)rows));
    if (take_sqrt) v = __fsqrt_rn(v);
    out[j] = v;
    note_flag(v, flag);
}

__global__ void col_finish_f64_kernel(const double* acc, long long cols, long long rows, int take_sqrt, float* out,
                                      int* flag) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    double v = acc[j] / (double)rows;
    if (take_sqrt) v = sqrt(v);
    out[j] = (float)v;
    note_flag((float)v, flag);
}

// Row-parallel binary64 partial sums: a CTA covers 128 columns x a slab of rows; thread = column.
template <int KIND, bool kVecF64>
__global__ void __launch_bounds__(128) col_partial_kernel(const float* __restrict__ a, long long m, long long cols,
                                                          long long ld, const void* vec, const float* vec2,
                                                          long long rows_per_cta, double* acc) {
    const long long col = (long long)blockIdx.x * 128 + threadIdx.x;
    if (col >= cols) return;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = min(m, r0 + rows_per_cta);
    double v = 0.0, v2 = 0.0;
    if (KIND != SKR_COLPASS_SUM && vec) v = kVecF64 ? ((const double*)vec)[col] : (double)((const float*)vec)[col];
    if (KIND == SKR_COLPASS_SQDEV) v2 = (double)vec2[col];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long r = r0;
    auto term = [&](float x) -> double {
        if (KIND == SKR_COLPASS_SUM) return (double)x;
        const double d = (double)x - v;
        if (KIND == SKR_COLPASS_CENTERED) return d;
        return (d - v2) * (d - v2);
    };
    for (; r + 4 <= r1; r += 4) {
        const float x0 = a[(r + 0) * ld + col], x1 = a[(r + 1) * ld + col];
        const float x2 = a[(r + 2) * ld + col], x3 = a[(r + 3) * ld + col];
        s0 += term(x0); s1 += term(x1); s2 += term(x2); s3 += term(x3);
    }
    for (; r < r1; ++r) s0 += term(a[r * ld + col]);
    atomicAdd(&acc[col], (s0 + s1) + (s2 + s3));
}

}  // namespace

extern "C" int skr_col_pass(int kind, const float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec,
                            int vec_is_f64, const float* d_vec2, float* d_acc, void* stream) {
    if (m <= 0 || cols <= 0) return SKR_OK;
    if (!d_a || !d_acc) return skr::fail(SKR_ERR_ARG, "skr_col_pass: null argument");
    if (kind == SKR_COLPASS_CENTERED && !d_vec) return skr::fail(SKR_ERR_ARG, "skr_col_pass: CENTERED needs the mean vector");
    if (kind == SKR_COLPASS_SQDEV && !d_vec2) return skr::fail(SKR_ERR_ARG, "skr_col_pass: SQDEV needs vec2");
    if (ld < cols || (ld % 4) != 0) return skr::fail(SKR_ERR_ARG, "skr_col_pass: ld must be >= cols and a multiple of 4");
    if (m > 0x7FFFFFFFll || cols > 0x7FFFFFFFll) return skr::fail(SKR_ERR_ARG, "skr_col_pass: matrix too large");
    CUtensorMap tmap;
    int rc = skr::make_tmap_2d(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_a, (uint64_t)cols, (uint64_t)m,
                               (uint64_t)ld * 4, kStripCols, kTileRows, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != SKR_OK) return rc;
    const unsigned grid = (unsigned)((cols + kStripCols - 1) / kStripCols);
    const size_t smem = (size_t)kStages * kTileBytes;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f64 = d_vec && vec_is_f64;
#define SKR_COL_LAUNCH(KIND, F64)                                                                              \
    do {                                                                                                       \
        auto kern = col_pass_kernel<KIND, F64>;                                                                \
        SKR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, 64, smem, s>>>(tmap, d_a, m, cols, ld, d_vec, d_vec2, d_acc);                             \
    } while (0)
    switch (kind) {
        case SKR_COLPASS_SUM: SKR_COL_LAUNCH(SKR_COLPASS_SUM, false); break;
        case SKR_COLPASS_CENTERED:
            if (f64) SKR_COL_LAUNCH(SKR_COLPASS_CENTERED, true); else SKR_COL_LAUNCH(SKR_COLPASS_CENTERED, false);
            break;
        case SKR_COLPASS_SQDEV:
            if (f64) SKR_COL_LAUNCH(SKR_COLPASS_SQDEV, true); else SKR_COL_LAUNCH(SKR_COLPASS_SQDEV, false);
            break;
        default: return skr::fail(SKR_ERR_ARG, "skr_col_pass: unknown pass kind %d", kind);
    }
#undef SKR_COL_LAUNCH
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_col_finish(const float* d_acc, int64_t cols, int64_t total_rows, int take_sqrt, float* d_out,
                              int* d_flag, void* stream) {
    if (cols <= 0) return SKR_OK;
    if (!d_acc || !d_out || total_rows <= 0) return skr::fail(SKR_ERR_ARG, "skr_col_finish: bad argument");
    if (d_flag) SKR_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), (cudaStream_t)stream));
    col_finish_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_acc, cols, total_rows, take_sqrt,
                                                                                      d_out, d_flag);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_col_partial_f64(int kind, const float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec,
                                   int vec_is_f64, const float* d_vec2, double* d_acc, void* stream) {
    if (m <= 0 || cols <= 0) return SKR_OK;
    if (!d_a || !d_acc) return skr::fail(SKR_ERR_ARG, "skr_col_partial_f64: null argument");
    if (kind == SKR_COLPASS_CENTERED && !d_vec) return skr::fail(SKR_ERR_ARG, "skr_col_partial_f64: CENTERED needs the mean");
    if (kind == SKR_COLPASS_SQDEV && !d_vec2) return skr::fail(SKR_ERR_ARG, "skr_col_partial_f64: SQDEV needs vec2");
    int dev = 0, sms = 0;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    SKR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long gx = (cols + 127) / 128;
    long long gy = ((long long)sms * 16 + gx - 1) / gx;  // ~16 CTAs of 128 threads per SM
    if (gy > m) gy = m;
    if (gy > 65535) gy = 65535;
    const long long rows_per_cta = (m + gy - 1) / gy;
    gy = (m + rows_per_cta - 1) / rows_per_cta;
    dim3 grid((unsigned)gx, (unsigned)gy);
    cudaStream_t s = (cudaStream_t)stream;
    const bool f64 = d_vec && vec_is_f64;
#define SKR_PART_LAUNCH(KIND)                                                                                         \
    do {                                                                                                              \
        if (f64) col_partial_kernel<KIND, true><<<grid, 128, 0, s>>>(d_a, m, cols, ld, d_vec, d_vec2, rows_per_cta, d_acc);  \
        else col_partial_kernel<KIND, false><<<grid, 128, 0, s>>>(d_a, m, cols, ld, d_vec, d_vec2, rows_per_cta, d_acc);     \
    } while (0)
    switch (kind) {
        case SKR_COLPASS_SUM: SKR_PART_LAUNCH(SKR_COLPASS_SUM); break;
        case SKR_COLPASS_CENTERED: SKR_PART_LAUNCH(SKR_COLPASS_CENTERED); break;
        case SKR_COLPASS_SQDEV: SKR_PART_LAUNCH(SKR_COLPASS_SQDEV); break;
        default: return skr::fail(SKR_ERR_ARG, "skr_col_partial_f64: unknown pass kind %d", kind);
    }
#undef SKR_PART_LAUNCH
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_col_finish_f64(const double* d_acc, int64_t cols, int64_t total_rows, int take_sqrt, float* d_out,
                                  int* d_flag, void* stream) {
    if (cols <= 0) return SKR_OK;
    if (!d_acc || !d_out || total_rows <= 0) return skr::fail(SKR_ERR_ARG, "skr_col_finish_f64: bad argument");
    if (d_flag) SKR_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), (cudaStream_t)stream));
    col_finish_f64_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_acc, cols, total_rows,
                                                                                          take_sqrt, d_out, d_flag);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}
