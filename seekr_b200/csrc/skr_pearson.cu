// K3 + K4: Pearson as a dense contraction on the 5th-generation tensor cores.
// Replaces seekr/pearson.py:32-44 (row standardise, np.inner / K).
//
// K3  skr_pearson_prepare: one CTA per row.  Row mean and std (ddof=0) with binary64 accumulation,
//     y = (x - mean) / std, then y is scaled by a power of two so that max|y| lies in [2^14, 2^15)
//     and split into two fp16 planes  hi = fp16(y'),  lo = fp16(y' - hi)  (22 significant bits;
//     the products of two fp16 values are exact in the fp32 accumulator).  Planes are K-major,
//     zero padded to [rows % 128 == 0][K % 64 == 0], i.e. ready for 128-byte-swizzled TMA boxes.
//
// K4  skr_pearson_gemm: C = alpha * sA_i * sB_j * (Ahi.Bhi' + Ahi.Blo' + Alo.Bhi'), a persistent,
//     warp-specialised tcgen05 kernel.  Per CTA: warp 0 = TMA producer (cp.async.bulk.tensor, SW128
//     boxes into a multi-stage shared-memory ring), warp 1 = MMA issuer (tcgen05.mma kind::f16,
//     fp32 accumulators in TMEM, 12 MMAs per 64-wide k-block: 3 products x 4 k-steps), warp 2 owns
//     the TMEM allocation, warps 4-11 = epilogue.  kCG = 2 pairs two CTAs (cta_group::2): a 256 x 256
//     output tile per pair, each CTA stages its own 128 rows of A and its half of the 256 rows of B,
//     the leader issues the MMAs and its commits are multicast to both CTAs' barriers.
//
//     Accumulator promotion.  The tensor core adds into its fp32 accumulator with truncation, so a
//     long K loop drifts low by ~3.5e-8 per MMA (measured: 2.7e-5 on the diagonal at K = 4096, far
//     outside the 1e-5 parity band, and growing with K).  The K loop is therefore cut into chunks of
//     chunk_kb k-blocks: each chunk accumulates from zero into one of two TMEM buffers (2 x 256
//     columns) and the epilogue warps drain the finished chunk with tcgen05.ld and add it, with
//     round-to-nearest fp32 adds, to running sums they keep in registers (8 warps x 128 columns,
//     register budget moved from the control warps with setmaxnreg) while the next chunk's MMAs run
//     into the other buffer.  The error no longer depends on K.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "skr_common.h"
#include "skr_device.cuh"
#include "skr_tma.h"

namespace {

// ---------------------------------------------------------------------------------------------
// K3: row standardise + split
// ---------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kPrepThreads / 32; ++i) t += s_buf[i];
    return t;
}

__device__ __forceinline__ float block_max(float v, float* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < kPrepThreads / 32; ++i) t = fmaxf(t, s_buf[i]);
    return t;
}

// T = float: the reference's float32 path (pearson.py:35-38 on float32 arrays keeps float32);
// T = double: everything else (numpy promotes ints / DataFrames to float64).
template <typename T>
__global__ void __launch_bounds__(kPrepThreads) prepare_kernel(const T* __restrict__ a, long long rows, long long K,
                                                               long long ld, long long kp, int standardize,
                                                               __half* __restrict__ hi, __half* __restrict__ lo,
                                                               float* __restrict__ row_scale) {
    __shared__ double s_d[kPrepThreads / 32];
    __shared__ float s_f[kPrepThreads / 32];
    const long long row = blockIdx.x;
    __half* hrow = hi + row * kp;
    __half* lrow = lo + row * kp;
    if (row >= rows) {  // padding rows: zeros
        for (long long j = threadIdx.x; j < kp; j += kPrepThreads) {
            hrow[j] = __float2half_rn(0.0f);
            lrow[j] = __float2half_rn(0.0f);
        }
        if (threadIdx.x == 0) row_scale[row] = 0.0f;
        return;
    }
    const T* x = a + row * ld;
    T mean = (T)0, mean2 = (T)0, sd = (T)1;
    float amax;
    if (standardize) {
        double s = 0.0;
        for (long long j = threadIdx.x; j < K; j += kPrepThreads) s += (double)x[j];
        mean = (T)(block_sum(s, s_d) / (double)K);
        // np.std of the centred row: its (tiny) mean is removed again before squaring (_methods.py:_var)
        double s2 = 0.0;
        for (long long j = threadIdx.x; j < K; j += kPrepThreads) s2 += (double)(T)(x[j] - mean);
        mean2 = (T)(block_sum(s2, s_d) / (double)K);
        double q = 0.0;
        float mx = 0.0f;
        for (long long j = threadIdx.x; j < K; j += kPrepThreads) {
            const T d = (T)(x[j] - mean);
            const T e = (T)(d - mean2);
            q += (double)(T)(e * e);
            mx = fmaxf(mx, fabsf((float)d));
        }
        sd = (T)sqrt((double)(T)(block_sum(q, s_d) / (double)K));
        amax = block_max(mx, s_f) / fabsf((float)sd);
    } else {
        float mx = 0.0f;
        for (long long j = threadIdx.x; j < K; j += kPrepThreads) mx = fmaxf(mx, fabsf((float)x[j]));
        amax = block_max(mx, s_f);
    }
    // power-of-two row scale: max |y| * 2^e in [2^14, 2^15); NaN / inf / zero rows are left unscaled
    int e = 0;
    if (amax > 0.0f && amax < INFINITY) {
        int ex;
        frexpf(amax, &ex);  // amax = f * 2^ex, f in [0.5, 1)
        e = 15 - ex;
        e = max(-100, min(100, e));
    }
    const T up = (T)ldexp(1.0, e);
    for (long long j = threadIdx.x; j < kp; j += kPrepThreads) {
        float h = 0.0f, l = 0.0f;
        if (j < K) {
            T y = standardize ? (T)((T)(x[j] - mean) / sd) : x[j];
            y = (T)(y * up);  // exact: a power of two
            const __half hh = __double2half(static_cast<double>(y));
            h = __half2float(hh);
            l = (float)(y - (T)h);  // exact in T
            hrow[j] = hh;
            lrow[j] = __float2half_rn(l);
        } else {
            hrow[j] = __float2half_rn(0.0f);
            lrow[j] = __float2half_rn(0.0f);
        }
    }
    if (threadIdx.x == 0) row_scale[row] = (float)ldexp(1.0, -e);
}

// float32 rows of at most 16 * kPrepThreads columns (k <= 6): the row is read ONCE with 128-bit loads and stays in
// registers for the three statistics passes and the split; hi / lo leave as 64-bit stores.  Same operations per
// element as prepare_kernel<float>; only the order of the binary64 partial sums differs.
template <int kVecs>
__global__ void __launch_bounds__(kPrepThreads) prepare_rows_kernel(const float* __restrict__ a, long long rows, long long K,
                                                                    long long ld, long long kp, int standardize,
                                                                    __half* __restrict__ hi, __half* __restrict__ lo,
                                                                    float* __restrict__ row_scale) {
    __shared__ double s_d[kPrepThreads / 32];
    __shared__ float s_f[kPrepThreads / 32];
    const long long row = blockIdx.x;
    uint2* hrow = reinterpret_cast<uint2*>(hi + row * kp);
    uint2* lrow = reinterpret_cast<uint2*>(lo + row * kp);
    const long long nv = kp / 4;  // 4-element groups of the padded row
    if (row >= rows) {
        for (long long g = threadIdx.x; g < nv; g += kPrepThreads) hrow[g] = lrow[g] = make_uint2(0u, 0u);
        if (threadIdx.x == 0) row_scale[row] = 0.0f;
        return;
    }
    const float4* x4 = reinterpret_cast<const float4*>(a + row * ld);
    const long long kv = K / 4;
    float4 x[kVecs];
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
        const long long g = (long long)v * kPrepThreads + threadIdx.x;
        x[v] = g < kv ? __ldg(x4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    auto valid = [&](int v) { return (long long)v * kPrepThreads + threadIdx.x < kv; };
    float mean = 0.0f, mean2 = 0.0f, sd = 1.0f, amax;
    if (standardize) {
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < kVecs; ++v)
            if (valid(v)) s += ((double)x[v].x + (double)x[v].y) + ((double)x[v].z + (double)x[v].w);
        mean = (float)(block_sum(s, s_d) / (double)K);
        double s2 = 0.0;
#pragma unroll
        for (int v = 0; v < kVecs; ++v)
            if (valid(v)) {
                x[v].x = __fsub_rn(x[v].x, mean); x[v].y = __fsub_rn(x[v].y, mean);
                x[v].z = __fsub_rn(x[v].z, mean); x[v].w = __fsub_rn(x[v].w, mean);
                s2 += ((double)x[v].x + (double)x[v].y) + ((double)x[v].z + (double)x[v].w);
            }
        mean2 = (float)(block_sum(s2, s_d) / (double)K);
        double q = 0.0;
        float mx = 0.0f;
#pragma unroll
        for (int v = 0; v < kVecs; ++v)
            if (valid(v)) {
                const float e0 = __fsub_rn(x[v].x, mean2), e1 = __fsub_rn(x[v].y, mean2);
                const float e2 = __fsub_rn(x[v].z, mean2), e3 = __fsub_rn(x[v].w, mean2);
                q += ((double)__fmul_rn(e0, e0) + (double)__fmul_rn(e1, e1)) + ((double)__fmul_rn(e2, e2) + (double)__fmul_rn(e3, e3));
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x[v].x), fabsf(x[v].y))), fmaxf(fabsf(x[v].z), fabsf(x[v].w)));
            }
        sd = (float)sqrt((double)(float)(block_sum(q, s_d) / (double)K));
        amax = block_max(mx, s_f) / fabsf(sd);
    } else {
        float mx = 0.0f;
#pragma unroll
        for (int v = 0; v < kVecs; ++v)
            if (valid(v)) mx = fmaxf(fmaxf(mx, fmaxf(fabsf(x[v].x), fabsf(x[v].y))), fmaxf(fabsf(x[v].z), fabsf(x[v].w)));
        amax = block_max(mx, s_f);
    }
    int e = 0;
    if (amax > 0.0f && amax < INFINITY) {
        int ex;
        frexpf(amax, &ex);
        e = max(-100, min(100, 15 - ex));
    }
    const float up = (float)ldexp(1.0, e);
    auto split = [&](float d, __half& hh, __half& ll) {
        float y = standardize ? __fdiv_rn(d, sd) : d;
        y = __fmul_rn(y, up);  // exact: a power of two
        hh = __float2half_rn(y);
        ll = __float2half_rn(__fsub_rn(y, __half2float(hh)));  // y - hi is exact in fp32
    };
#pragma unroll
    for (int v = 0; v < kVecs; ++v) {
        const long long g = (long long)v * kPrepThreads + threadIdx.x;
        if (g >= nv) continue;
        __half h[4], l[4];
        if (g < kv) {
            split(x[v].x, h[0], l[0]); split(x[v].y, h[1], l[1]);
            split(x[v].z, h[2], l[2]); split(x[v].w, h[3], l[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = l[i] = __float2half_rn(0.0f);
        }
        hrow[g] = *reinterpret_cast<const uint2*>(h);
        lrow[g] = *reinterpret_cast<const uint2*>(l);
    }
    // padded tail beyond kVecs * kPrepThreads groups cannot exist: the host checks kp / 4 <= kVecs * kPrepThreads
    if (threadIdx.x == 0) row_scale[row] = (float)ldexp(1.0, -e);
}

// ---------------------------------------------------------------------------------------------
// K4: tcgen05 GEMM
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128;        // A rows per CTA (UMMA M = 128 * kCG)
constexpr int kBN = 256;        // UMMA N
constexpr int kBK = 64;         // fp16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;  // warps 0-3: control (TMA, MMA, TMEM alloc, spare); warps 4-11: epilogue
constexpr int kEpiThreads = 256;
constexpr int kChunkKbDefault = 2;        // k-blocks accumulated in TMEM before promotion to the fp32 register sums
constexpr int kCtrlRegs = 56;
constexpr int kEpiRegs = 224;
constexpr int kTmemCols = 512;  // two 256-column accumulators
constexpr int kGroupM = 8;      // tile rasterisation: 8 tile-rows share their B tiles in L2

template <int kCG>
struct GemmCfg {
    static constexpr int kBNLocal = kBN / kCG;                         // B rows staged per CTA
    static constexpr int kABytes = kBM * kBK * 2;                      // 16 KB per plane
    static constexpr int kBBytes = kBNLocal * kBK * 2;                 // 16 / 32 KB per plane
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;      // 64 / 96 KB
    static constexpr int kStages = kCG == 2 ? 3 : 2;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct GemmParams {
    long long m, n;            // valid rows of A / B
    int num_kb;                // k-blocks of 64
    int chunk_kb;              // k-blocks accumulated in TMEM before promotion (see the header comment)
    int accumulate;            // add to what C already holds (K segments after the first)
    int do_mirror;             // last K segment: the values this launch stores are final
    int mirror_epi;            // symmetric mode, final values: a tile above the diagonal also stores its transpose
    int tiles_m, tiles_n;
    float alpha;
    const float* a_scale;
    const float* b_scale;
    void* c;
    long long ldc;
    int c_is_f64;
    int symmetric;  // A == B, square tiles: only tiles with tn >= tm are computed, the rest is mirrored
    // similarity-graph edge counts fused into the epilogue (kmer_leiden.py:91-104; skr_pearson_gemm_edges): the
    // finished r values are compared with the threshold where they are produced, so the offsets pass of the edge
    // extraction never reads the matrix.  counts[row * SKR_SIM_SLICES + column slice]; a 256-wide tile lies inside
    // one slice (the slice width is a multiple of 256)
    unsigned long long* edge_counts;
    float edge_thr;          // edge iff r >= edge_thr (skr_graph.cu: edge_threshold), off the diagonal
    int edge_upper;          // only columns right of the diagonal
    long long edge_row0;     // whole-matrix index of row 0 of this block
    long long edge_width;    // slice width in columns
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n.reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}"
        ::"r"(skr::smem_u32(bar)), "r"(rank)
        : "memory");
}

template <int kCG>
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    if constexpr (kCG == 2) {
        // both CTAs of the pair signal the leader's barrier: clear the CTA-rank bit of the cluster address
        const uint32_t bar_addr = skr::smem_u32(bar) & 0xFEFFFFFFu;
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4}], [%2];"
            ::"r"(skr::smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1)
            : "memory");
    } else {
        skr::tma_load_2d(dst, map, c0, c1, bar);
    }
}

// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;           // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32; // stride byte offset
    d |= (uint64_t)1 << 46;           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;           // SWIZZLE_128B
    return d;
}

template <int kCG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (kCG == 2) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}

template <int kCG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (kCG == 2) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(skr::smem_u32(bar)), "h"((uint16_t)3)
            : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(skr::smem_u32(bar))
                     : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: several can be in flight, tmem_wait_ld() covers all of them
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Tile order: groups of kGroupM tile-rows sweep the tile-columns, so the CTAs running at the same time
// share a handful of A and B slabs in L2.  Symmetric mode keeps the order but skips tiles below the diagonal.
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int symmetric, int& tm, int& tn) {
    if (!symmetric) {
        const int per_group = kGroupM * tiles_n;
        const int group = t / per_group;
        const int first_m = group * kGroupM;
        const int gsize = min(kGroupM, tiles_m - first_m);
        const int r = t - group * per_group;
        tm = first_m + r % gsize;
        tn = r / gsize;
        return;
    }
    int first_m = 0, r = t;
    for (;;) {  // at most tiles_m / kGroupM iterations
        const int R = min(kGroupM, tiles_m - first_m);
        const int ncols = tiles_n - first_m;                      // tile-columns first_m .. tiles_n-1
        const int size = R * (R + 1) / 2 + (ncols - R) * R;       // columns 0..R-1 hold 1..R tiles, the rest R each
        if (r < size) {
            const int tri = R * (R + 1) / 2;
            int c, lm;
            if (r < tri) {
                c = 0;
                while ((c + 1) * (c + 2) / 2 <= r) ++c;
                lm = r - c * (c + 1) / 2;
            } else {
                const int rr = r - tri;
                c = R + rr / R;
                lm = rr % R;
            }
            tm = first_m + lm;
            tn = first_m + c;
            return;
        }
        r -= size;
        first_m += kGroupM;
    }
}

__host__ __device__ inline long long symmetric_tile_count(int tiles) { return (long long)tiles * (tiles + 1) / 2; }

template <int kCG>
__global__ void __launch_bounds__(kGemmThreads, 1)
pearson_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    const GemmParams p) {
    using Cfg = GemmCfg<kCG>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
    uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;      // [2]
    uint64_t* tmem_empty_bar = bars + 2 * Cfg::kStages + 2; // [2]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = kCG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / kCG;
    const int num_clusters = gridDim.x / kCG;
    const int num_tiles = p.symmetric ? (int)symmetric_tile_count(p.tiles_m) : p.tiles_m * p.tiles_n;

    if (warp == 0 && lane == 0) {
        skr::tma_prefetch_desc(&map_a_hi);
        skr::tma_prefetch_desc(&map_a_lo);
        skr::tma_prefetch_desc(&map_b_hi);
        skr::tma_prefetch_desc(&map_b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            skr::mbar_init(&full_bar[s], kCG);   // leader's arrive.expect_tx (+ the peer's remote arrive)
            skr::mbar_init(&empty_bar[s], 1);    // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            skr::mbar_init(&tmem_full_bar[a], 1);
            skr::mbar_init(&tmem_empty_bar[a], kCG * kEpiThreads);
        }
        skr::fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kCG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(skr::smem_u32(tmem_base_slot)), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(skr::smem_u32(tmem_base_slot)), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
    if (warp == 0) {
        // ===================== TMA producer (every CTA stages its own halves) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                int tm, tn;
                tile_coords(t, p.tiles_m, p.tiles_n, p.symmetric, tm, tn);
                const int a_row = (tm * kCG + (int)cta_rank) * kBM;
                const int b_row = tn * kBN + (int)cta_rank * Cfg::kBNLocal;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    skr::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    unsigned char* st = smem + stage * Cfg::kStageBytes;
                    if (leader) skr::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * kCG);
                    tma_load_tile<kCG>(st, &map_a_hi, kb * kBK, a_row, &full_bar[stage]);
                    tma_load_tile<kCG>(st + Cfg::kABytes, &map_a_lo, kb * kBK, a_row, &full_bar[stage]);
                    tma_load_tile<kCG>(st + 2 * Cfg::kABytes, &map_b_hi, kb * kBK, b_row, &full_bar[stage]);
                    tma_load_tile<kCG>(st + 2 * Cfg::kABytes + Cfg::kBBytes, &map_b_lo, kb * kBK, b_row, &full_bar[stage]);
                    if (kCG == 2 && !leader) mbar_arrive_cluster(&full_bar[stage], 0);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only, one thread) ===========================
        if (leader && lane == 0) {
            // instruction descriptor: D = F32, A = B = F16, both K-major, N = 256, M = 128 * kCG
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)((kBM * kCG) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.chunk_kb) {
                    skr::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kBN);
                    const int kb1 = min(p.num_kb, kb0 + p.chunk_kb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        skr::mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t st = skr::smem_u32(smem + stage * Cfg::kStageBytes);
                        const uint64_t a_hi = make_smem_desc(st);
                        const uint64_t a_lo = make_smem_desc(st + Cfg::kABytes);
                        const uint64_t b_hi = make_smem_desc(st + 2 * Cfg::kABytes);
                        const uint64_t b_lo = make_smem_desc(st + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
                        for (int k4 = 0; k4 < kBK / kUmmaK; ++k4) {
                            const uint64_t adv = (uint64_t)((k4 * kUmmaK * 2) >> 4);  // 32 bytes per k-step inside the swizzle row
                            umma_f16<kCG>(tmem_d, a_lo + adv, b_hi + adv, idesc, (kb != kb0 || k4 != 0) ? 1u : 0u);
                            umma_f16<kCG>(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_f16<kCG>(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
                        }
                        umma_commit<kCG>(&empty_bar[stage]);  // frees the stage in both CTAs once the MMAs have read it
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit<kCG>(&tmem_full_bar[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: promote chunks, then scale and store =======================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        const int ew = warp - 4;
        const int quarter = ew & 3;   // TMEM lanes 32*quarter .. +31 (a warp may only touch lanes 32*(warp%4)..)
        const int half = ew >> 2;     // columns 128*half .. +127 of the 256-wide tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters) {
            int tm, tn;
            tile_coords(t, p.tiles_m, p.tiles_n, p.symmetric, tm, tn);
            float sum[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) sum[i] = 0.0f;
            for (int kb0 = 0; kb0 < p.num_kb; kb0 += p.chunk_kb) {
                skr::mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBN + half * 128);
                // The drain paces the MMA pipe (the accumulator buffer of chunk i is the one chunk i + 2 needs): two
                // loads in flight per wait instead of one, and the buffer is handed back as soon as the last load has
                // landed in registers, before the additions (4 round trips + 128 adds -> 2 round trips before the
                // arrive).
                uint32_t v0[32], v1[32];
                tmem_ld32_nowait(taddr, v0);
                tmem_ld32_nowait(taddr + 32u, v1);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[i] = __fadd_rn(sum[i], __uint_as_float(v0[i]));
                tmem_ld32_nowait(taddr + 64u, v0);
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[32 + i] = __fadd_rn(sum[32 + i], __uint_as_float(v1[i]));
                tmem_ld32_nowait(taddr + 96u, v1);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive_cluster(&tmem_empty_bar[acc], 0);  // the leader's MMA warp owns this barrier
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[64 + i] = __fadd_rn(sum[64 + i], __uint_as_float(v0[i]));
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[96 + i] = __fadd_rn(sum[96 + i], __uint_as_float(v1[i]));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            const long long row = (long long)(tm * kCG + (int)cta_rank) * kBM + quarter * 32 + lane;
            if (row < p.m) {
                const float rs = p.alpha * __ldg(p.a_scale + row);
                const long long colbase = (long long)tn * kBN + half * 128;
                // fused edge count: the diagonal column and the first column that may hold an edge of this row
                const long long gdiag = p.edge_row0 + row;
                const long long jmin = p.edge_upper ? gdiag + 1 : 0;
                int ecnt = 0;
                // symmetric mode: the tile below the diagonal is this tile's transpose.  Lane = row, so the 32 lanes
                // of a warp store one column of the tile as 128 consecutive bytes of the mirrored row.
                const bool mir = p.mirror_epi && tn > tm;
#pragma unroll
                for (int c = 0; c < 128; c += 4) {
                    const long long col0 = colbase + c;
                    if (col0 >= p.n) break;
                    if (!p.c_is_f64) {
                        float* dst = reinterpret_cast<float*>(p.c) + row * p.ldc + col0;
                        if (col0 + 4 <= p.n && ((p.ldc & 3) == 0)) {
                            const float4 bs = __ldg(reinterpret_cast<const float4*>(p.b_scale + col0));
                            float4 o;
                            o.x = sum[c + 0] * rs * bs.x;
                            o.y = sum[c + 1] * rs * bs.y;
                            o.z = sum[c + 2] * rs * bs.z;
                            o.w = sum[c + 3] * rs * bs.w;
                            if (p.accumulate) {
                                const float4 prev = *reinterpret_cast<const float4*>(dst);
                                o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                            }
                            *reinterpret_cast<float4*>(dst) = o;
                            if (mir) {
                                float* mt = reinterpret_cast<float*>(p.c) + col0 * p.ldc + row;
                                mt[0] = o.x;
                                mt[p.ldc] = o.y;
                                mt[2 * p.ldc] = o.z;
                                mt[3 * p.ldc] = o.w;
                            }
                            if (p.edge_counts) {
                                if (col0 >= jmin && (gdiag < col0 || gdiag >= col0 + 4)) {
                                    ecnt += (o.x >= p.edge_thr) + (o.y >= p.edge_thr) + (o.z >= p.edge_thr) + (o.w >= p.edge_thr);
                                } else {
                                    ecnt += (o.x >= p.edge_thr && col0 + 0 != gdiag && col0 + 0 >= jmin);
                                    ecnt += (o.y >= p.edge_thr && col0 + 1 != gdiag && col0 + 1 >= jmin);
                                    ecnt += (o.z >= p.edge_thr && col0 + 2 != gdiag && col0 + 2 >= jmin);
                                    ecnt += (o.w >= p.edge_thr && col0 + 3 != gdiag && col0 + 3 >= jmin);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (col0 + i < p.n) {
                                    const float o = sum[c + i] * rs * __ldg(p.b_scale + col0 + i) + (p.accumulate ? dst[i] : 0.0f);
                                    dst[i] = o;
                                    if (mir) reinterpret_cast<float*>(p.c)[(col0 + i) * p.ldc + row] = o;
                                    if (p.edge_counts) ecnt += (o >= p.edge_thr && col0 + i != gdiag && col0 + i >= jmin);
                                }
                        }
                    } else {
                        double* dst = reinterpret_cast<double*>(p.c) + row * p.ldc + col0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (col0 + i < p.n) {
                                const double o = (double)(sum[c + i] * rs * __ldg(p.b_scale + col0 + i)) + (p.accumulate ? dst[i] : 0.0);
                                dst[i] = o;
                                if (mir) reinterpret_cast<double*>(p.c)[(col0 + i) * p.ldc + row] = o;
                            }
                    }
                }
                if (ecnt) atomicAdd(p.edge_counts + row * SKR_SIM_SLICES + colbase / p.edge_width, (unsigned long long)ecnt);
            }
        }
    }

    tc_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        if constexpr (kCG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}


template <int kCG>
int launch_gemm(const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mb_hi, const CUtensorMap& mb_lo,
                GemmParams p, cudaStream_t stream) {
    using Cfg = GemmCfg<kCG>;
    auto kern = pearson_gemm_kernel<kCG>;
    SKR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    int dev = 0, sms = 0;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    SKR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    p.tiles_m = (int)((p.m + kBM * kCG - 1) / (kBM * kCG));
    p.tiles_n = (int)((p.n + kBN - 1) / kBN);
    if (p.symmetric && (kBM * kCG != kBN || p.tiles_m != p.tiles_n))
        return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: symmetric mode needs square tiles (cta_group 2) and m == n");
    long long clusters = p.symmetric ? symmetric_tile_count(p.tiles_m) : (long long)p.tiles_m * p.tiles_n;
    if (clusters > sms / kCG) clusters = sms / kCG;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * kCG));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SKR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, p));
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

}  // namespace

extern "C" int64_t skr_pearson_rows_padded(int64_t rows) { return (rows + 127) / 128 * 128; }
extern "C" int64_t skr_pearson_k_padded(int64_t K) { return (K + 63) / 64 * 64; }

extern "C" int skr_pearson_prepare(const void* d_a, int a_is_f64, int64_t rows, int64_t K, int64_t ld,
                                   int row_standardize, uint16_t* d_hi, uint16_t* d_lo, float* d_row_scale,
                                   void* stream) {
    if (rows <= 0 || K <= 0) return SKR_OK;
    if (!d_a || !d_hi || !d_lo || !d_row_scale) return skr::fail(SKR_ERR_ARG, "skr_pearson_prepare: null argument");
    if (ld < K) return skr::fail(SKR_ERR_ARG, "skr_pearson_prepare: ld < K");
    const int64_t rp = skr_pearson_rows_padded(rows), kp = skr_pearson_k_padded(K);
    if (rp > 0x7FFFFFFFll) return skr::fail(SKR_ERR_ARG, "skr_pearson_prepare: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    const bool reg_path = !a_is_f64 && K % 4 == 0 && ld % 4 == 0 && (((uintptr_t)d_a & 15) == 0) && kp / 4 <= 4 * kPrepThreads &&
                          !getenv("SEEKR_B200_PREPARE_GENERIC");
    if (reg_path)
        prepare_rows_kernel<4><<<(unsigned)rp, kPrepThreads, 0, s>>>((const float*)d_a, rows, K, ld, kp, row_standardize,
                                                                     (__half*)d_hi, (__half*)d_lo, d_row_scale);
    else if (a_is_f64)
        prepare_kernel<double><<<(unsigned)rp, kPrepThreads, 0, s>>>((const double*)d_a, rows, K, ld, kp, row_standardize,
                                                                     (__half*)d_hi, (__half*)d_lo, d_row_scale);
    else
        prepare_kernel<float><<<(unsigned)rp, kPrepThreads, 0, s>>>((const float*)d_a, rows, K, ld, kp, row_standardize,
                                                                    (__half*)d_hi, (__half*)d_lo, d_row_scale);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

struct EdgeArgs {
    unsigned long long* counts = nullptr;
    float thr = 0.0f;
    int upper = 0;
    long long row0 = 0, width = 1;
};

static int pearson_gemm_impl(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                             const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n,
                             int64_t K, double alpha, void* d_c, int c_is_f64, int64_t ldc, int symmetric,
                             const EdgeArgs& edges, void* stream);

extern "C" int skr_pearson_gemm(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                                const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n,
                                int64_t K, double alpha, void* d_c, int c_is_f64, int64_t ldc, int symmetric, void* stream) {
    return pearson_gemm_impl(d_a_hi, d_a_lo, d_a_scale, m, d_b_hi, d_b_lo, d_b_scale, n, K, alpha, d_c, c_is_f64, ldc,
                             symmetric, EdgeArgs{}, stream);
}

extern "C" int skr_pearson_gemm_edges(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                                      const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n,
                                      int64_t K, double alpha, float* d_c, int64_t ldc, int symmetric, int64_t row0,
                                      double cutoff, int upper_only, int64_t* d_offsets, void* stream) {
    if (!d_offsets) return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm_edges: null offsets");
    if (symmetric && !upper_only)
        return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm_edges: the symmetric GEMM computes the upper tiles only; count both "
                                      "orientations with skr_sim_edge_offsets on the finished matrix");
    cudaStream_t s = (cudaStream_t)stream;
    SKR_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, sizeof(int64_t) * (size_t)((m > 0 ? m : 0) * SKR_SIM_SLICES + 1), s));
    if (m <= 0 || n <= 0) return SKR_OK;
    EdgeArgs e;
    e.counts = reinterpret_cast<unsigned long long*>(d_offsets) + 1;  // scanned in place afterwards
    const float cut = (float)cutoff;
    e.thr = cut > 0.0f ? cut : 1.40129846432481707e-45f;  // !(x < cut) && x > 0 as one comparison (skr_graph.cu)
    e.upper = upper_only;
    e.row0 = row0;
    e.width = skr_sim_slice_width(n);
    int rc = pearson_gemm_impl(d_a_hi, d_a_lo, d_a_scale, m, d_b_hi, d_b_lo, d_b_scale, n, K, alpha, d_c, 0, ldc, symmetric, e,
                               stream);
    if (rc != SKR_OK) return rc;
    return skr_sim_offsets_scan(d_offsets, m, stream);
}

static int pearson_gemm_impl(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                             const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n,
                             int64_t K, double alpha, void* d_c, int c_is_f64, int64_t ldc, int symmetric,
                             const EdgeArgs& edges, void* stream) {
    if (m <= 0 || n <= 0) return SKR_OK;
    if (symmetric && (d_a_hi != d_b_hi || d_a_lo != d_b_lo || m != n))
        return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: symmetric mode needs identical operands");
    if (!d_a_hi || !d_a_lo || !d_a_scale || !d_b_hi || !d_b_lo || !d_b_scale || !d_c || K <= 0)
        return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: null argument");
    if (ldc < n) return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: ldc < n");
    if (((uintptr_t)d_c & 15)) return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: output must be 16-byte aligned");
    const int64_t kp = skr_pearson_k_padded(K);
    const int64_t mp = skr_pearson_rows_padded(m), np_ = skr_pearson_rows_padded(n);
    if (kp / kBK > 0x7FFFFFFF || mp > 0x7FFFFFFFll || np_ > 0x7FFFFFFFll)
        return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: problem too large");
    int cg = 2;
    if (const char* env = getenv("SEEKR_B200_GEMM_CTA_GROUP")) cg = atoi(env) == 1 ? 1 : 2;
    if (cg == 1) symmetric = 0;  // 128 x 256 tiles are not square: compute the full matrix
    const uint32_t b_box_rows = cg == 2 ? 128 : 256;
    // K segments: beyond kSegKb k-blocks the contraction is cut into launches that add into C, so the running
    // sums of the promotion never see more than kSegKb / chunk_kb chunk values (their systematic rounding on
    // count-like rows grows with that number); the segment results are combined by 2 ... 8 fp32 adds in C.
    int seg_kb = 128;  // 8192 columns
    if (const char* env = getenv("SEEKR_B200_GEMM_SEGMENT")) seg_kb = atoi(env) > 0 ? atoi(env) : 0x7FFFFFFF;
    const int total_kb = (int)(kp / kBK);
    int rc = SKR_OK;
    for (int kb_first = 0; kb_first < total_kb; kb_first += seg_kb) {
    const int kb_count = total_kb - kb_first < seg_kb ? total_kb - kb_first : seg_kb;
    const size_t koff = (size_t)kb_first * kBK;  // elements into every plane row (a multiple of 64: 128-byte aligned)
    const uint64_t kw = (uint64_t)kb_count * kBK;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    if ((rc = skr::make_tmap_2d(&ma_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_a_hi + koff, kw, (uint64_t)mp,
                                (uint64_t)kp * 2, kBK, kBM, CU_TENSOR_MAP_SWIZZLE_128B)) != SKR_OK) return rc;
    if ((rc = skr::make_tmap_2d(&ma_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_a_lo + koff, kw, (uint64_t)mp,
                                (uint64_t)kp * 2, kBK, kBM, CU_TENSOR_MAP_SWIZZLE_128B)) != SKR_OK) return rc;
    if ((rc = skr::make_tmap_2d(&mb_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_b_hi + koff, kw, (uint64_t)np_,
                                (uint64_t)kp * 2, kBK, b_box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) != SKR_OK) return rc;
    if ((rc = skr::make_tmap_2d(&mb_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_b_lo + koff, kw, (uint64_t)np_,
                                (uint64_t)kp * 2, kBK, b_box_rows, CU_TENSOR_MAP_SWIZZLE_128B)) != SKR_OK) return rc;
    GemmParams p{};
    p.m = m;
    p.n = n;
    p.num_kb = kb_count;
    p.accumulate = kb_first > 0;
    p.do_mirror = kb_first + kb_count >= total_kb;
    // Promotion interval.  Two error sources pull in opposite directions on count-like rows (a few big z-scores
    // among thousands of small ones): inside a chunk the tensor core truncates every addend to the accumulator's
    // ulp, so small products that share a chunk with a large one lose low bits (grows with the chunk); across
    // chunks the fp32 running sum, once it holds the large product, rounds each of the many similar small chunk
    // values the same way (grows with the number of chunks).  Measured max |r - binary64| over sparsity 0.8 ... 0.01
    // and K = 4096 ... 65 536 (tools/pearson_error_sweep.py, profiles/r01_pearson_error_sweep.txt):
    // 1 k-block 1.4e-5, 2 k-blocks 7.4e-6, 4 k-blocks 8.2e-6 -- hence 2.
    p.chunk_kb = kChunkKbDefault;
    if (const char* env = getenv("SEEKR_B200_GEMM_CHUNK"))  // experiment knob
        if (atoi(env) >= 1 && atoi(env) <= 64) p.chunk_kb = atoi(env);
    p.alpha = (float)alpha;
    p.a_scale = d_a_scale;
    p.b_scale = d_b_scale;
    p.c = d_c;
    p.ldc = ldc;
    p.c_is_f64 = c_is_f64;
    p.symmetric = symmetric;
    // The separate mirror pass this replaces (read the upper tiles, transpose through shared memory, write) took
    // 2.2 ms at 50k x 50k; from the epilogue the GEMM is 0.5 ms longer (profiles/r02_gemm_mirror.txt).
    p.mirror_epi = symmetric && p.do_mirror;
    if (kb_first + kb_count >= total_kb) {  // the finished values exist in the last K segment only
        p.edge_counts = edges.counts;
        p.edge_thr = edges.thr;
        p.edge_upper = edges.upper;
        p.edge_row0 = edges.row0;
        p.edge_width = edges.width;
    }
    if (p.edge_width <= 0) p.edge_width = 1;
    cudaStream_t s = (cudaStream_t)stream;
    rc = cg == 2 ? launch_gemm<2>(ma_hi, ma_lo, mb_hi, mb_lo, p, s) : launch_gemm<1>(ma_hi, ma_lo, mb_hi, mb_lo, p, s);
    if (rc != SKR_OK) return rc;
    }  // K segments
    return rc;
}
