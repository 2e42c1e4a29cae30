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
// streams [192 rows x 32 columns] boxes of the strip through an 8-deep shared-memory ring with
// TMA (cp.async.bulk.tensor + mbarrier), and three consumer warps (lane = column) take turns: load a
// tile into registers, then run its adds when the running sums arrive from the previous tile.
// The running sums enter and leave through d_acc, so row shards on several GPUs can be chained.
//
// Scalable passes (skr_col_partial_f64): row-parallel binary64 partial sums for one all-reduce.
#include <cuda.h>
#include <cuda_runtime.h>

#include "skr_common.h"
#include "skr_device.cuh"
#include "skr_tma.h"

namespace {

constexpr int kStripCols = 32;
constexpr int kTileRows = 192;
constexpr int kStages = 8;
constexpr int kRelay = 3;  // consumer warps taking turns on the add chain
constexpr int kTileBytes = kTileRows * kStripCols * 4;  // 24 KB
constexpr int kColThreads = 32 * (1 + kRelay);

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int KIND, bool kVecF64>
__global__ void __launch_bounds__(kColThreads) col_pass_kernel(const __grid_constant__ CUtensorMap tmap,
                                                               const float* __restrict__ a, long long m, long long cols,
                                                               long long ld, const void* vec, const float* vec2,
                                                               float* acc_io) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];
    __shared__ float relay_acc[32];
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

    if (warp == 0) {
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

    // Consumer warps (lane = column col0 + lane).  The adds of one column are a dependent fp32 chain, 4 cycles
    // per row, and that chain is the whole critical path.  Shared loads cannot feed it directly: a warp has six
    // scoreboards, so at most a handful of loads are individually tracked and an add that waits for its operand
    // also waits for younger loads on the same scoreboard (measured: 7.6 cycles per row with loads and adds
    // interleaved).  So kRelay warps take turns: a warp copies its whole tile into registers (value() applied,
    // off the chain), gives the slot back to the producer, then waits for the running sums from the warp that
    // holds the previous tile, runs kTileRows register-only adds and passes the sums on through shared memory.
    const int cw = warp - 1;
    const long long col = col0 + lane;
    const bool active = col < cols;
    const bool has_vec = vec != nullptr;
    float vf = 0.0f, v2 = 0.0f;
    double vd = 0.0;
    if (active) {
        if (has_vec) {
            if (kVecF64) vd = reinterpret_cast<const double*>(vec)[col];
            else vf = reinterpret_cast<const float*>(vec)[col];
        }
        if (KIND == SKR_COLPASS_SQDEV) v2 = vec2[col];
    }
    // one IEEE operation per step, exactly the reference's sequence: counts -= mean (one rounding),
    // np.std: x - arrmean, square, sequential fp32 sum (kmer_counts.py:169,174; numpy _methods.py:_var).
    auto value = [&](float x) -> float {
        float y = x;
        if (KIND != SKR_COLPASS_SUM && has_vec)
            y = kVecF64 ? __double2float_rn(__dsub_rn((double)x, vd)) : __fsub_rn(x, vf);
        if (KIND == SKR_COLPASS_SQDEV) {
            const float d = __fsub_rn(y, v2);
            y = __fmul_rn(d, d);
        }
        return y;
    };
    float acc = 0.0f;
    const uint32_t relay_slot = (uint32_t)__cvta_generic_to_shared(&relay_acc[lane]);
    auto relay_load = [&]() -> float {  // address formed before the barrier, so the hand-over is barrier + one load
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(relay_slot) : "memory");
        return v;
    };
    for (long long t = cw; t < ntiles; t += kRelay) {
        const int s = (int)(t % kStages);
        const uint32_t ph = (uint32_t)((t / kStages) & 1);
        const int rows = (int)min((long long)kTileRows, m - t * kTileRows);
        skr::mbar_wait(&full_bar[s], ph);
        if (rows == kTileRows) {
            float y[kTileRows];
#pragma unroll
            for (int r = 0; r < kTileRows; ++r) y[r] = value(tiles[s][r][lane]);
            __syncwarp();
            if (lane == 0) skr::mbar_arrive(&empty_bar[s]);
            if (t == 0) {
                acc = active ? acc_io[col] : 0.0f;
            } else {
                named_bar_sync(1 + cw, 64);
                acc = relay_load();
            }
#pragma unroll
            for (int r = 0; r < kTileRows; ++r) acc = __fadd_rn(acc, y[r]);
        } else {  // ragged last tile
            if (t == 0) {
                acc = active ? acc_io[col] : 0.0f;
            } else {
                named_bar_sync(1 + cw, 64);
                acc = relay_load();
            }
#pragma unroll 8
            for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, value(tiles[s][r][lane]));
            __syncwarp();
            if (lane == 0) skr::mbar_arrive(&empty_bar[s]);
        }
        if (t + 1 < ntiles) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(relay_slot), "f"(acc) : "memory");
            named_bar_arrive(1 + (cw + 1) % kRelay, 64);
        } else if (active) {
            acc_io[col] = acc;
        }
    }
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
        kern<<<grid, kColThreads, smem, s>>>(tmap, d_a, m, cols, ld, d_vec, d_vec2, d_acc);                             \
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
