// K1: sliding-window k-mer counting with a fused normalisation epilogue, plus the element-wise
// kernels of the get_counts() tail.  Replaces seekr/kmer_counts.py:140-151 and :189-209.
//
// One CTA owns one record at a time (records are handed out through a device counter, so a
// long transcript does not stall a fixed partner set).  The record's overlapping k-mers are
// counted into a shared-memory histogram of 16-bit sub-counters (two bins per 32-bit word, one
// shared atomicAdd per window, consecutive equal k-mers merged per thread).  A record with more
// than 65 520 windows is processed in segments whose partial histograms are spilled into a
// per-CTA 32-bit row in global memory, so no sub-counter can overflow.  The epilogue turns the
// integer count c of every bin into the reference's float32 value: the c-fold binary64 sum of
// 1000/(L-k+1) (kmer_counts.py:144-150 adds the increment once per window in Python floats),
// rounded once, then log2(x+1) / -mean / /std as requested, and writes the dense output row
// with 128-bit stores.  HBM traffic per record is its packed codes + mask + one output row.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "skr_common.h"
#include "skr_device.cuh"

namespace {

constexpr int kSegChunks = 4095;  // 4095 chunks x 16 windows = 65 520 < 65 536 increments per segment
constexpr int kTab = 32;          // counts below this are looked up, larger ones take the generic chain_sum

template <int K>
struct CountCfg {
    static constexpr int kBins = 1 << (2 * K);
    static constexpr int kWords = kBins / 2;  // 16-bit sub-counters, two per word
    static constexpr int kThreads = K <= 4 ? 64 : (K == 5 ? 128 : (K <= 7 ? 256 : 1024));
    static constexpr size_t kSmem = (size_t)kWords * 4;
};

struct CountParams {
    const uint32_t* codes;
    const uint32_t* mask;
    const uint64_t* blk_off;
    const uint32_t* len;
    long long m;
    int log2_pre;
    const void* mean;
    const void* std_;
    void* out;
    long long ld_out;
    SkrMinCell* min_cell;
    unsigned int* work_counter;
    uint32_t* spill;  // [gridDim.x][kBins] 32-bit counts for records longer than one segment
    // records the warp kernel hands over to the CTA kernel (too long for one warp): list + its length
    uint32_t* long_list;
    unsigned int* long_count;
    // per-column minimum of the values written (float bits; values are >= 0 on this path), for the
    // Log2.post shift when normalisation is deferred to the element-wise pass
    uint32_t* colmin;
    const float* rstd;            // RN(1/std) per column: enables the 5-instruction exact division (fp32 vectors only)
    int no_store;                 // column-minimum pass only: nothing is written to `out`
    const SkrMinCell* post_cell;  // Log2.post fused into the epilogue: + |min|, + 1, log2 with this cell's minimum
    // speculative Log2.post (skr_post_spec): the shift in post_cell was derived from the vectors alone and is the
    // true matrix minimum iff some record has a zero count in column spec->zero_col; the kernel reports that
    SkrPostSpec* spec;
    // ... and with it the whole tail folded into one multiply-add per value: a_j = 1/std_j, b_j = shift + 1 -
    // mean_j/std_j (skr_post_spec_affine, binary64 then one rounding); out = log2(x * a_j + b_j)
    const float* post_a;
    const float* post_b;
    const uint32_t* skip_flag;    // the launch does nothing when *skip_flag == skip_value (device-side choice of the route)
    uint32_t skip_value;
    uint32_t spec_epoch;          // written to spec->zero_seen when a zero count is met in the arg-min column
    SkrMinCell* min_reset;        // reset at the start of the launch (the cell the two-pass route behind it will track)
    // k >= 7: the finished values of a ZERO count in every column (vectors and Log2.post applied); a record of a few
    // thousand windows leaves 80-95 % of 16 384 / 65 536 bins empty, and a quad of empty bins is copied from here
    // (one 16-byte load from L2) instead of loading three vector quads and redoing the arithmetic
    const float4* zero_row;
    uint32_t max_length;          // longest record of this launch when the caller knows it (0 = unknown)
    // accurate column statistics (norm_vectors in one pass): per-column sum and sum of squares of the values
    // written, accumulated in fp32 per thread over its records and added here in binary64 at the end
    double* colsum;
    double* colsq;
};

// Column minima for the two-pass Log2.post path (values are >= 0 there, so float bits order like
// unsigned integers).  Almost every column holds a zero in some record; a thread therefore keeps one
// bit per column it owns ("a zero was seen", at most 64 columns = one 64-bit mask; the thread ->
// column mapping is the same for every record) and only columns that have not shown a zero yet take
// the slow road: compare with the global minimum and atomicMin when smaller.  The mask is rotated by
// kRot bits per epilogue step, so the nibble of step s is always the low one when step s runs and the
// mask is back in place after the kSteps steps of a record.
template <int kSteps>
__device__ __forceinline__ void colmin_note(unsigned long long& seen, const uint32_t (&c4)[4], const float (&r)[4],
                                            uint32_t* colmin, int q) {
    constexpr int kRot = 64 / kSteps;
    static_assert(kSteps >= 1 && kSteps <= 16 && 64 % kSteps == 0, "one nibble per step must fit the mask");
    const uint32_t z4 = (c4[0] == 0 ? 1u : 0u) | (c4[1] == 0 ? 2u : 0u) | (c4[2] == 0 ? 4u : 0u) | (c4[3] == 0 ? 8u : 0u);
    const uint32_t known = (uint32_t)seen & 0xFu;
    if ((z4 | known) != 0xFu) {  // a non-zero value in a column with no zero so far (rare after the first records)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!((z4 | known) & (1u << e))) {
                const uint32_t b = __float_as_uint(r[e]);
                if (b < colmin[4 * q + e]) atomicMin(&colmin[4 * q + e], b);
            }
        }
    }
    seen |= z4;
    if constexpr (kRot < 64) seen = (seen >> kRot) | (seen << (64 - kRot));
}

// store 0 into colmin[] for every column of this thread whose mask bit is set
template <int kSteps>
__device__ __forceinline__ void colmin_flush(unsigned long long seen, uint32_t* colmin, int first_q, int stride_q) {
    constexpr int kRot = 64 / kSteps;
#pragma unroll
    for (int st = 0; st < kSteps; ++st)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if ((seen >> ((kRot * st) % 64 + e)) & 1ull) colmin[4 * (first_q + st * stride_q) + e] = 0u;
}

// shared-memory accesses by 32-bit shared-window address: the generic-pointer forms make the compiler
// rebuild the window base (S2UR SR_CgaCtaId, ...) at every access, 4 extra instructions per window
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t inc, uint32_t pred) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p red.shared.add.u32 [%0], %1;\n}" ::"r"(addr), "r"(inc), "r"(pred) : "memory");
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// log2 of the Log2.post tail, log2((z + |min|) + 1): the argument is >= 1 (or NaN), where MUFU.LG2 is within
// 2^-22.6 absolute of the exact value per unit of magnitude (measured on B200: tools/log2_error.py, profiles/);
// libdevice's log2f is a 25-instruction polynomial per element, which would triple the count kernel's epilogue.
// Every Log2.post path of the library (fused epilogues and the element-wise passes) uses this one function, so
// they agree bit for bit with each other; against numpy's log2 the values stay inside the 1e-5 parity band.
__device__ __forceinline__ float log2_post(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// value of a bin whose count is beyond the table (rare: low-complexity records); kept out of line so the
// hot epilogue stays small in the instruction cache
__device__ __noinline__ float chain_bin_value(double inc, uint32_t c, int log2_pre) {
    float v = __double2float_rn(skr::chain_sum(inc, c));
    if (log2_pre) v = log2f(__fadd_rn(v, 1.0f));
    return v;
}

// Counts beyond the table.  The c-fold binary64 sum s differs from p = RN(c * inc) by less than c + 1 units in the
// last place of p (every add rounds by at most half an ulp of a partial sum no larger than p, the product by half an
// ulp), and fp32(s) = fp32(p) unless a float32 rounding boundary -- a binary64 whose low 29 mantissa bits are exactly
// 2^28 -- lies that close to p.  So p decides in ~8 instructions whenever its low 29 bits are further than c + 2
// from 2^28 (all but a fraction c * 2^-27 of the values); only the rest walk the exact chain.  Counts of 32 and
// more are the rule for k <= 5 (256 / 1 024 bins for thousands of windows), where the chain made up 40 % of the
// kernel's instructions (profiles/r02_ncu_count_k4_raw.txt).
__device__ __forceinline__ float slow_bin_value(double inc, uint32_t c, int log2_pre) {
    const double prod = __dmul_rn(inc, (double)c);
    const int low = (int)((unsigned long long)__double_as_longlong(prod) & 0x1FFFFFFFull);
    const uint32_t dist = (uint32_t)abs(low - 0x10000000);
    if (c < 0x08000000u && dist > c + 2u) {
        float v = __double2float_rn(prod);
        if (log2_pre) v = log2f(__fadd_rn(v, 1.0f));
        return v;
    }
    return chain_bin_value(inc, c, log2_pre);
}

// Correctly rounded a / b from y = RN(1/b) with two Newton corrections on the quotient (Markstein: when y is
// the correctly rounded reciprocal and q is within one ulp of a/b, q + (a - b*q)*y rounds to RN(a/b); the first
// correction makes q faithful, the second makes it exact).  Valid when no intermediate over/underflows: the
// caller only enables it for |b| in [2^-40, 2^40] and |a| in {0} U [2^-60, 2^40] (checked on the host for the
// vectors, true by construction for per-kb counts).  5 instructions instead of the 14 of the generic
// div.rn.f32 sequence (MUFU.RCP, FCHK, 5 FFMA, branch, slow-path call); verified against __fdiv_rn on
// 2^32 operand pairs by skr_selftest_division (tests/test_gpu_counts.py).
__device__ __forceinline__ float div_by_rcp(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

// two lanes of (x - mean) / std at once: nm = -mean, nb = -std, y = RN(1/std); the same operations as
// __fsub_rn + div_by_rcp on each lane (x + (-m) is x - m; packed fp32x2 instructions round to nearest, no ftz)
__device__ __forceinline__ uint64_t sub_div_by_rcp2(uint64_t x, uint64_t nm, uint64_t nb, uint64_t y) {
    const uint64_t a = skr::f2_add(x, nm);
    uint64_t q = skr::f2_mul(a, y);
    uint64_t r = skr::f2_fma(nb, q, a);
    q = skr::f2_fma(r, y, q);
    r = skr::f2_fma(nb, q, a);
    return skr::f2_fma(r, y, q);
}

__device__ __forceinline__ uint64_t f2_neg(uint64_t v) { return v ^ 0x8000000080000000ull; }

// -mean, /std of 4 consecutive columns starting at 4 * q, with the reference's roundings (kmer_counts.py:169,175)
template <bool kVecF64>
__device__ __forceinline__ void normalize4(const CountParams& p, int q, float (&r)[4]) {
    if constexpr (!kVecF64) {
        if (p.mean && p.std_ && p.rstd) {  // two columns per instruction (packed fp32x2), exact reciprocal division
            const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mean) + q);
            const float4 sv = __ldg(reinterpret_cast<const float4*>(p.std_) + q);
            const float4 yv = __ldg(reinterpret_cast<const float4*>(p.rstd) + q);
            const uint64_t z0 = sub_div_by_rcp2(skr::f2_pack(r[0], r[1]), f2_neg(skr::f2_pack(mv.x, mv.y)),
                                                f2_neg(skr::f2_pack(sv.x, sv.y)), skr::f2_pack(yv.x, yv.y));
            const uint64_t z1 = sub_div_by_rcp2(skr::f2_pack(r[2], r[3]), f2_neg(skr::f2_pack(mv.z, mv.w)),
                                                f2_neg(skr::f2_pack(sv.z, sv.w)), skr::f2_pack(yv.z, yv.w));
            skr::f2_unpack(z0, r[0], r[1]);
            skr::f2_unpack(z1, r[2], r[3]);
            return;
        }
    }
    if (p.mean) {
        if constexpr (kVecF64) {
            const double* mv = reinterpret_cast<const double*>(p.mean) + 4 * q;
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = __double2float_rn(__dsub_rn((double)r[e], __ldg(mv + e)));
        } else {
            const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mean) + q);
            r[0] = __fsub_rn(r[0], mv.x); r[1] = __fsub_rn(r[1], mv.y);
            r[2] = __fsub_rn(r[2], mv.z); r[3] = __fsub_rn(r[3], mv.w);
        }
    }
    if (p.std_) {
        if constexpr (kVecF64) {
            const double* sv = reinterpret_cast<const double*>(p.std_) + 4 * q;
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = __double2float_rn(__ddiv_rn((double)r[e], __ldg(sv + e)));
        } else {
            const float4 sv = __ldg(reinterpret_cast<const float4*>(p.std_) + q);
            if (p.rstd) {
                const float4 yv = __ldg(reinterpret_cast<const float4*>(p.rstd) + q);
                r[0] = div_by_rcp(r[0], sv.x, yv.x); r[1] = div_by_rcp(r[1], sv.y, yv.y);
                r[2] = div_by_rcp(r[2], sv.z, yv.z); r[3] = div_by_rcp(r[3], sv.w, yv.w);
            } else {
                r[0] = __fdiv_rn(r[0], sv.x); r[1] = __fdiv_rn(r[1], sv.y);
                r[2] = __fdiv_rn(r[2], sv.z); r[3] = __fdiv_rn(r[3], sv.w);
            }
        }
    }
}

// counts of 4 consecutive bins -> the reference's float32 values (per-kb chain, log2.pre, -mean, /std).
template <bool kVecF64>
__device__ __forceinline__ void finish4(const uint32_t (&c4)[4], uint32_t tab_addr, double inc, const CountParams& p, int q,
                                        float (&r)[4]) {
    if (((c4[0] | c4[1] | c4[2] | c4[3]) & ~(uint32_t)(kTab - 1)) == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = lds_f32(tab_addr + c4[e] * 4);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            r[e] = c4[e] < kTab ? lds_f32(tab_addr + c4[e] * 4) : slow_bin_value(inc, c4[e], p.log2_pre);
    }
    normalize4<kVecF64>(p, q, r);
}

// 16 window starts of one chunk -> shared-memory histogram (two 16-bit sub-counters per word).
// x: 32 bases (2 bits each, first base in the top bits); mb: the chunk's 16+K-1 mask bits; nv: valid windows.
// hist_addr: shared-window address of the histogram.  Bin b lives in word b>>1, half b&1.
__device__ __forceinline__ void red_add_shared_always(uint32_t addr, uint32_t inc) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
}

// kAligned: hist_addr is a multiple of the histogram size, so base | offset replaces base + offset and the
// word address is a single three-input logic operation on the shifted k-mer.
template <int K, bool kAligned = false, int KH = K>
__device__ __forceinline__ void count_chunk(uint32_t hist_addr, uint64_t x, uint32_t mb, int nv, uint32_t pass = 0) {
    constexpr uint32_t kMask = (1u << (2 * K)) - 1;
    constexpr uint32_t kOffMask = (kMask >> 1) << 2;  // (kmer >> 1) * 4 out of t = kmer << 1
    if constexpr (KH != K) {
        // split histogram (see BatchCfg): every window through the predicated path, filtered by its leading base(s)
        constexpr uint32_t kOffH = (((1u << (2 * KH)) - 1) >> 1) << 2;
        uint32_t bad = mb;
#pragma unroll
        for (int i = 1; i < K; ++i) bad |= mb << i;
        const uint32_t ok = ~(bad >> (K - 1)) & 0xFFFFu & ~(0xFFFFu >> nv);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t t = (uint32_t)(x >> (64 - 2 * (j + K) - 1));
            const uint32_t sel = ((t >> (2 * KH + 1)) & ((1u << (2 * (K - KH))) - 1)) == pass;
            red_add_shared((kAligned ? ((t & kOffH) | hist_addr) : ((t & kOffH) + hist_addr)), (t & 2u) ? 0x10000u : 1u,
                           (ok & (0x8000u >> j)) && sel);
        }
        return;
    }
    if (mb == 0 && nv == 16) {
        // all 16 + K - 1 bases equal (homopolymer run): one add of 16 instead of 16 colliding adds
        const uint64_t same = (x ^ (x << 2)) >> (64 - 2 * (16 + K - 2));
        if (same == 0) {
            const uint32_t kmer = (uint32_t)(x >> (64 - 2 * K)) & kMask;
            red_add_shared_always(hist_addr + (kmer >> 1) * 4, 16u << ((kmer & 1) * 16));
            return;
        }
        // the common case: every window counts, no predicates (5 instructions per window:
        // funnel shift, address, half select, increment, red)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t t = (uint32_t)(x >> (64 - 2 * (j + K) - 1));
            const uint32_t addr = kAligned ? ((t & kOffMask) | hist_addr) : hist_addr + (t & kOffMask);
            red_add_shared_always(addr, (t & 2u) ? 0x10000u : 1u);
        }
        return;
    }
    // window j is bad if any of its K mask bits is set: OR the mask with itself shifted by 1..K-1
    uint32_t bad = mb;
#pragma unroll
    for (int i = 1; i < K; ++i) bad |= mb << i;
    const uint32_t ok = ~(bad >> (K - 1)) & 0xFFFFu & ~(0xFFFFu >> nv);  // bit 15-j: window j is counted
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        // t = k-mer shifted left by one: bits [2K..1] = k-mer, so (t & mask) is the byte offset of its word / 2
        const uint32_t t = (uint32_t)(x >> (64 - 2 * (j + K) - 1));
        const uint32_t off = t & kOffMask;
        const uint32_t inc = 1u << ((t << 3) & 16);                  // kmer & 1 ? 0x10000 : 1
        red_add_shared(hist_addr + off, inc, ok & (0x8000u >> j));
    }
}

template <int K, bool kVecF64, typename OutT>
__global__ void __launch_bounds__(CountCfg<K>::kThreads) count_kernel(const CountParams p) {
    using Cfg = CountCfg<K>;
    constexpr int T = Cfg::kThreads;
    extern __shared__ __align__(16) uint32_t hist[];
    __shared__ long long s_rec;
    __shared__ float s_wmin[T / 32];
    __shared__ int s_wnan[T / 32];
    // per-record value table: s_tab[c] = float32 value of a bin seen c times (c < kTab), so the
    // epilogue does one shared load per bin instead of a binary64 add chain
    __shared__ float s_tab[kTab];
    __shared__ double s_tab64[sizeof(OutT) == 8 ? kTab : 1];

    const int tid = threadIdx.x;
    if (p.skip_flag && *p.skip_flag == p.skip_value) return;
    if (p.min_reset && blockIdx.x == 0 && threadIdx.x == 0) { p.min_reset->min_ordered = skr::ordered_encode(INFINITY); p.min_reset->nan_seen = 0; }
    float tmin = INFINITY;
    int tnan = 0;
    unsigned long long zero_seen = 0;
    constexpr int kSteps = (Cfg::kBins / 4 + T - 1) / T;  // epilogue steps per thread and record
    float shift = 0.0f;
    if (p.post_cell) shift = p.post_cell->nan_seen ? __int_as_float(0x7FC00000) : fabsf(skr::ordered_decode(p.post_cell->min_ordered));
    int zq = -1, ze = 0;  // speculative Log2.post: quad / element of the arg-min column (skr_post_spec)
    uint32_t zseen = 0;
    if (p.spec && p.spec->zero_col >= 0) { zq = p.spec->zero_col >> 2; ze = p.spec->zero_col & 3; }
    uint32_t* spill = p.spill + (size_t)blockIdx.x * Cfg::kBins;

    for (;;) {
        if (tid == 0) {
            long long r = (long long)atomicAdd(p.work_counter, 1u);
            if (p.long_list) r = r < (long long)*p.long_count ? (long long)p.long_list[r] : p.m;
            s_rec = r;
        }
        __syncthreads();  // also orders the previous record's histogram reads before the zeroing below
        const long long rec = s_rec;
        if (rec >= p.m) break;

        if constexpr (Cfg::kWords % 4 == 0) {
            for (int i = tid; i < Cfg::kWords / 4; i += T) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
        } else {
            for (int i = tid; i < Cfg::kWords; i += T) hist[i] = 0;
        }
        const uint32_t L = p.len[rec];
        const long long nwin = (long long)L - K + 1;  // > 0 or the row is all zero counts (L == k-1 is rejected on the host)
        const double inc = nwin > 0 ? 1000.0 / (double)nwin : 0.0;
        if (tid >= T - kTab) {  // the last kTab threads fill the table while the others start counting
            const int c = tid - (T - kTab);
            const double v64 = skr::chain_sum(inc, (uint32_t)c);
            if constexpr (sizeof(OutT) == 8) {
                s_tab64[c] = v64;
            } else {
                float v = __double2float_rn(v64);
                if (p.log2_pre) v = log2f(__fadd_rn(v, 1.0f));
                s_tab[c] = v;
            }
        }
        __syncthreads();

        const uint64_t b0 = p.blk_off[rec];
        const uint32_t* __restrict__ cw = p.codes + b0 * 4;
        const uint32_t* __restrict__ mw = p.mask + b0 * 2;
        const long long nchunks = nwin > 0 ? (nwin + 15) / 16 : 0;
        const int nseg = (int)((nchunks + kSegChunks - 1) / kSegChunks);

        for (int seg = 0; seg < (nseg > 0 ? nseg : 1); ++seg) {
            const long long c_begin = (long long)seg * kSegChunks;
            const long long c_end = min(nchunks, c_begin + kSegChunks);
            for (long long c = c_begin + tid; c < c_end; c += T) {
                const uint64_t x = ((uint64_t)__ldg(cw + c) << 32) | __ldg(cw + c + 1);
                const uint64_t m64 = ((uint64_t)__ldg(mw + (c >> 1)) << 32) | __ldg(mw + (c >> 1) + 1);
                // the 16 + K - 1 mask bits of this chunk, first base in the top bit
                const uint32_t mb = (uint32_t)((m64 << (16 * (int)(c & 1))) >> (64 - (16 + K - 1)));
                const long long left = nwin - c * 16;
                const int nv = left < 16 ? (int)left : 16;
                count_chunk<K>(skr::smem_u32(hist), x, mb, nv);
            }
            __syncthreads();
            if (nseg > 1) {
                // spill this segment's 16-bit partial counts into the CTA's 32-bit row and start over
                for (int w = tid; w < Cfg::kWords; w += T) {
                    const uint32_t v = hist[w];
                    uint32_t lo = v & 0xFFFFu, hi = v >> 16;
                    if (seg > 0) {
                        lo += spill[2 * w];
                        hi += spill[2 * w + 1];
                    }
                    spill[2 * w] = lo;
                    spill[2 * w + 1] = hi;
                    hist[w] = 0;
                }
                __syncthreads();
            }
        }

        // ---- epilogue: 4 bins per thread per step ----------------------------------------------
        OutT* __restrict__ orow = reinterpret_cast<OutT*>(p.out) + (size_t)rec * (size_t)p.ld_out;
        for (int q = tid; q < Cfg::kBins / 4; q += T) {
            uint32_t c4[4];
            if (nseg > 1) {
                const uint4 v = reinterpret_cast<const uint4*>(spill)[q];
                c4[0] = v.x; c4[1] = v.y; c4[2] = v.z; c4[3] = v.w;
            } else {
                const uint2 v = reinterpret_cast<const uint2*>(hist)[q];
                c4[0] = v.x & 0xFFFFu; c4[1] = v.x >> 16; c4[2] = v.y & 0xFFFFu; c4[3] = v.y >> 16;
            }
            if constexpr (sizeof(OutT) == 8) {
                double r[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) r[e] = c4[e] < kTab ? s_tab64[c4[e]] : skr::chain_sum(inc, c4[e]);
                reinterpret_cast<double2*>(orow)[2 * q] = make_double2(r[0], r[1]);
                reinterpret_cast<double2*>(orow)[2 * q + 1] = make_double2(r[2], r[3]);
            } else {
                float r[4];
                if (p.zero_row && (c4[0] | c4[1] | c4[2] | c4[3]) == 0) {
                    const float4 z = __ldg(p.zero_row + q);  // normalised and Log2.post-shifted already
                    r[0] = z.x; r[1] = z.y; r[2] = z.z; r[3] = z.w;
                } else {
                    finish4<kVecF64>(c4, skr::smem_u32(s_tab), inc, p, q, r);
                    if (p.post_cell && !p.colmin) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) r[e] = log2_post(__fadd_rn(__fadd_rn(r[e], shift), 1.0f));
                    }
                }
                if (q == zq && (ze == 0 ? c4[0] : ze == 1 ? c4[1] : ze == 2 ? c4[2] : c4[3]) == 0) zseen = 1;
                if (p.colmin) colmin_note<kSteps>(zero_seen, c4, r, p.colmin, q);
                if (p.min_cell) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) skr::min_update(r[e], tmin, tnan);
                }
                if (!p.no_store) reinterpret_cast<float4*>(orow)[q] = make_float4(r[0], r[1], r[2], r[3]);
            }
        }
        // the __syncthreads at the top of the loop separates these reads from the next zeroing
    }

    if (zseen) p.spec->zero_seen = p.spec_epoch;
    if (p.min_cell) skr::min_commit<T>(tmin, tnan, s_wmin, s_wnan, p.min_cell);
    if (p.colmin && tid < Cfg::kBins / 4) colmin_flush<kSteps>(zero_seen, p.colmin, tid, T);
}


// ---------------------------------------------------------------------------------------------
// Team-per-record variant for k <= 6 (the headline path).  A record of a few thousand bases is too
// little work for 256 threads: the CTA kernel above spends its time in barriers and serialized
// latencies (work counter, lengths, offsets, codes).  Here a small team owns a private histogram
// and runs records end to end; the next record's index is fetched while the current one is
// processed.  k <= 5: team = one warp (histogram <= 2 KB, 8 teams per CTA, __syncwarp only).
// k = 6: team = one 64-thread CTA (8 KB histogram -> 27 CTAs = 54 warps per SM; a one-warp team would
// cap the SM at 27 warps and leave the issue slots idle).  Records longer than kLongWin windows are
// pushed to a list that the CTA kernel drains afterwards (256 threads and the 32-bit spill path).
// ---------------------------------------------------------------------------------------------
constexpr long long kLongWin = 32768;

template <int K>
struct WarpCfg {
    static constexpr int kBins = 1 << (2 * K);
    static constexpr int kWords = kBins / 2;
    static constexpr int kTeam = K == 6 ? 64 : 32;       // threads per record
    static constexpr int kTeams = K == 6 ? 1 : 8;        // records in flight per CTA
    static constexpr int kThreads = kTeam * kTeams;
    static constexpr size_t kSmem = (size_t)kTeams * (kWords * 4 + kTab * 4);
    static constexpr int kMinCtas = K == 6 ? 24 : 8;     // register cap: 48 / 64 resident warps per SM
};

template <int K, bool kVecF64>
__global__ void __launch_bounds__(WarpCfg<K>::kThreads, WarpCfg<K>::kMinCtas) count_warp_kernel(const CountParams p) {
    using Cfg = WarpCfg<K>;
    constexpr int TT = Cfg::kTeam;
    extern __shared__ __align__(16) uint32_t smem_w[];
    __shared__ long long s_next;
    __shared__ float s_wmin[2];
    __shared__ int s_wnan[2];
    const int team = threadIdx.x / TT, lane = threadIdx.x % TT;  // lane: index inside the team
    if (p.skip_flag && *p.skip_flag == p.skip_value) return;
    if (p.min_reset && blockIdx.x == 0 && threadIdx.x == 0) { p.min_reset->min_ordered = skr::ordered_encode(INFINITY); p.min_reset->nan_seen = 0; }
    int zq = -1, ze = 0;  // speculative Log2.post: quad / element of the arg-min column (skr_post_spec)
    uint32_t zseen = 0;
    if (p.spec && p.spec->zero_col >= 0) { zq = p.spec->zero_col >> 2; ze = p.spec->zero_col & 3; }
    uint32_t* hist = smem_w + team * Cfg::kWords;
    float* tab = reinterpret_cast<float*>(smem_w + Cfg::kTeams * Cfg::kWords) + team * kTab;
    const uint32_t hist_addr = skr::smem_u32(hist), tab_addr = skr::smem_u32(tab);
    float tmin = INFINITY;
    int tnan = 0;
    unsigned long long zero_seen = 0;
    constexpr int kSteps = (Cfg::kBins / 4 + TT - 1) / TT;  // epilogue steps per thread and record
    float shift = 0.0f;
    if (p.post_cell) shift = p.post_cell->nan_seen ? __int_as_float(0x7FC00000) : fabsf(skr::ordered_decode(p.post_cell->min_ordered));

    auto team_sync = [&]() {
        if constexpr (TT == 32) __syncwarp(); else __syncthreads();
    };
    // the team leader draws the next record index; everybody learns it at the next team_sync
    auto draw = [&]() -> long long {
        return lane == 0 ? (long long)atomicAdd(p.work_counter, 1u) : 0;
    };
    auto share = [&](long long mine) -> long long {
        if constexpr (TT == 32) {
            return (long long)__shfl_sync(0xFFFFFFFFu, (unsigned int)mine, 0);
        } else {
            if (lane == 0) s_next = mine;
            __syncthreads();
            return s_next;
        }
    };
    long long rec = share(draw());

    while (rec < p.m) {
        const long long mine_next = draw();  // in flight while this record is processed
        const uint32_t L = __ldg(p.len + rec);
        const uint64_t b0 = __ldg(p.blk_off + rec);
        const long long nwin = (long long)L - K + 1;
        if (nwin > kLongWin) {
            if (lane == 0) p.long_list[atomicAdd(p.long_count, 1u)] = (uint32_t)rec;
        } else {
            const uint32_t* __restrict__ cw = p.codes + b0 * 4;
            const uint32_t* __restrict__ mw = p.mask + b0 * 2;
            const int nchunks = nwin > 0 ? (int)((nwin + 15) / 16) : 0;
            // first chunk's words are requested before the histogram is cleared
            uint32_t w0 = 0, w1 = 0, m0 = 0, m1 = 0;
            if (lane < nchunks) {
                w0 = __ldg(cw + lane); w1 = __ldg(cw + lane + 1);
                m0 = __ldg(mw + (lane >> 1)); m1 = __ldg(mw + (lane >> 1) + 1);
            }
            if constexpr (Cfg::kWords % 4 == 0) {
                for (int i = lane; i < Cfg::kWords / 4; i += TT) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
            } else {
                for (int i = lane; i < Cfg::kWords; i += TT) hist[i] = 0;
            }
            const double inc = nwin > 0 ? 1000.0 / (double)nwin : 0.0;
            if (lane == TT - 1) {  // the literal chain: kTab-1 dependent binary64 adds, cheap in issue slots
                double acc = 0.0;
                tab[0] = p.log2_pre ? log2f(1.0f) : 0.0f;
#pragma unroll 8
                for (int c = 1; c < kTab; ++c) {
                    acc = __dadd_rn(acc, inc);
                    float v = __double2float_rn(acc);
                    if (p.log2_pre) v = log2f(__fadd_rn(v, 1.0f));
                    tab[c] = v;
                }
            }
            team_sync();
            for (int c = lane; c < nchunks; c += TT) {
                const int cn = c + TT;
                uint32_t n0 = 0, n1 = 0, q0 = 0, q1 = 0;
                if (cn < nchunks) {  // next iteration's words
                    n0 = __ldg(cw + cn); n1 = __ldg(cw + cn + 1);
                    q0 = __ldg(mw + (cn >> 1)); q1 = __ldg(mw + (cn >> 1) + 1);
                }
                const uint64_t x = ((uint64_t)w0 << 32) | w1;
                const uint64_t m64 = ((uint64_t)m0 << 32) | m1;
                const uint32_t mb = (uint32_t)((m64 << (16 * (c & 1))) >> (64 - (16 + K - 1)));
                const long long left = nwin - (long long)c * 16;
                count_chunk<K>(hist_addr, x, mb, left < 16 ? (int)left : 16);
                w0 = n0; w1 = n1; m0 = q0; m1 = q1;
            }
            team_sync();
            float* __restrict__ orow = reinterpret_cast<float*>(p.out) + (size_t)rec * (size_t)p.ld_out;
            for (int q = lane; q < Cfg::kBins / 4; q += TT) {
                const uint2 v = lds_v2(hist_addr + q * 8);
                const uint32_t c4[4] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16};
                float r[4];
                finish4<kVecF64>(c4, tab_addr, inc, p, q, r);
                if (q == zq && (ze == 0 ? c4[0] : ze == 1 ? c4[1] : ze == 2 ? c4[2] : c4[3]) == 0) zseen = 1;
                if (p.colmin) colmin_note<kSteps>(zero_seen, c4, r, p.colmin, q);
                if (p.post_cell) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) r[e] = log2_post(__fadd_rn(__fadd_rn(r[e], shift), 1.0f));
                }
                if (p.min_cell) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) skr::min_update(r[e], tmin, tnan);
                }
                if (!p.no_store) reinterpret_cast<float4*>(orow)[q] = make_float4(r[0], r[1], r[2], r[3]);
            }
        }
        rec = share(mine_next);  // also separates this record's histogram reads from the next clearing
        if constexpr (TT == 32) __syncwarp();
    }
    if (zseen) p.spec->zero_seen = p.spec_epoch;
    if (p.colmin && lane < Cfg::kBins / 4) colmin_flush<kSteps>(zero_seen, p.colmin, lane, TT);
    if (p.min_cell) {
        if constexpr (TT == 32) {
            if (tmin != tmin) { tnan = 1; tmin = INFINITY; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                tmin = fminf(tmin, __shfl_xor_sync(0xFFFFFFFFu, tmin, o));
                tnan |= __shfl_xor_sync(0xFFFFFFFFu, tnan, o);
            }
            if (lane == 0) {
                if (tmin < INFINITY) atomicMin(&p.min_cell->min_ordered, skr::ordered_encode(tmin));
                if (tnan) atomicOr(&p.min_cell->nan_seen, 1u);
            }
        } else {
            skr::min_commit<TT>(tmin, tnan, s_wmin, s_wnan, p.min_cell);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Batch kernel (k = 6, float32 output): a persistent CTA of 16 worker warps + 1 bookkeeping warp takes
// kB consecutive records at a time.
//   count phase   the batch's windows are cut into units of 32 (two chunks that share their code and mask
//                 words).  Whole units of all records come first in the unit order, the (at most kB) ragged
//                 tail units last, and the units are dealt round-robin to the 512 worker threads whatever
//                 record they belong to: the threads stay busy on short and long records alike, and the
//                 predicated path of a ragged unit is executed by one warp per batch instead of by every
//                 warp that happens to hold the end of a record.  Each record has its own histogram.
//   epilogue      thread w owns bin quads w and w + 512 of EVERY record.  Its slices of mean / std / 1/std
//                 are loaded once per kernel and stay in registers, so the normalisation costs no memory
//                 traffic at all (the warp kernel re-reads 48 KB of vectors from L2 per record, which made
//                 it latency-bound: 0.32 ms fused against 0.21 ms raw).  The value table of a record (32
//                 floats) sits in one register per lane and is read with an indexed warp shuffle: one
//                 instruction per bin instead of address arithmetic + a shared load.  Each warp writes 512
//                 contiguous bytes per step; the histogram words are zeroed as they are read.
//   bookkeeping   warp 16 draws the next batch, reads its lengths / offsets, runs the kTab-entry binary64
//                 value chains and the unit prefix while the workers are busy (double-buffered BatchMeta).
// ---------------------------------------------------------------------------------------------
// kSplit (k = 8 with vectors): the 65 536 columns are produced in kPasses = 4 passes of 16 384 -- pass p counts the
// windows whose first base is p into a k = 7 sized histogram and finishes columns [p * 16 384, (p + 1) * 16 384) --
// so that five records share a batch and the vector slices are re-read once per batch instead of once per record
// (a single 128 KB histogram per SM left the fused flavours L2-bound at 0.38-0.42 of the HBM roofline).
template <int K, bool kSplit = false>
struct BatchCfg {
    static constexpr int kHK = kSplit ? K - 1 : K;     // k of the histogram a CTA holds
    static constexpr int kPasses = kSplit ? 4 : 1;
    static constexpr int kBins = 1 << (2 * kHK);
    static constexpr int kWords = kBins / 2;
    static constexpr int kHistBytes = kWords * 4;
    // 512 threads, two CTAs per SM = 1024 threads = 64 registers each (with a 17th warp the limit was 56 and the
    // epilogue spilled).  Every thread owns bin quads in the epilogue; in the count phase warp 15 keeps the books
    // and the other 15 warps count.
    // k = 7: 32 KB per histogram, so one CTA of 1024 threads per SM with five records per batch
    static constexpr int kWorkers = kHK >= 7 ? 1024 : 512;
    static constexpr int kThreads = kWorkers;
    static constexpr int kCtasPerSm = kHK >= 7 ? 1 : 2;
    static constexpr int kCounters = kWorkers - 32;
    static constexpr int kQuads = kBins / 4;
    // k = 6: a thread owns kQ = 2 quads of EVERY record.  k = 4, 5: a histogram has fewer quads than the CTA has
    // threads, so the threads form kG groups of kQuads threads; group g takes records g, g + kG, ... of the batch
    // in the epilogue (a thread still owns fixed columns, so its vector slices stay in registers), and a batch holds
    // many more records, which also evens out the count phase (units per thread 6.7 instead of 1.7).
    // k = 7: kQ = 4 quads per thread; their vector slices (48 registers) do not fit beside the epilogue, so the
    // epilogue runs in kChunks = 2 passes over the batch with kQc = 2 quads' slices re-read (L2) before each pass:
    // 192 KB of vector reads per batch of five records instead of 192 KB per record.
    static constexpr int kQ = kQuads >= kWorkers ? kQuads / kWorkers : 1;
    static constexpr int kG = kQuads >= kWorkers ? 1 : kWorkers / kQuads;
    static constexpr int kQc = kQ > 2 ? 2 : kQ;       // quads whose vector slices are in registers at a time
    static constexpr int kChunks = kQ / kQc;
    static constexpr int kB = kHK >= 8 ? 1 : (kHK == 7 ? 5 : (kHK == 6 ? 8 : 32));  // records per batch
    // k = 8: one 128 KB histogram is all an SM holds: no slack for the size-aligned trick (the word address is an add)
    static constexpr bool kAligned = kHK <= 7;
    // histograms start at a multiple of their size (count_chunk_full): one histogram of slack
    static constexpr size_t kSmem = (size_t)(kB + (kAligned ? 1 : 0)) * kHistBytes;
    static_assert(kQuads % 32 == 0, "a warp works on one record at a time in the epilogue (table lookups are warp shuffles)");
    static_assert((kQuads >= kWorkers && kQuads % kWorkers == 0) || kWorkers % kQuads == 0, "whole quads per thread / whole groups per CTA");
    static_assert(kB % 32 == 0 || kB < 32, "the bookkeeping warp handles the records of a batch 32 at a time");
    static_assert(kQ % kQc == 0, "whole chunks");
};

template <int kB>
struct BatchMeta {
    alignas(128) float tab[kB][kTab];  // 128-byte rows: table address | 4*count
    long long rec0;                    // first record of the batch, < 0: no more work
    int nrec;
    int pass;                          // kSplit: which quarter of the columns this batch produces
    uint32_t prefix[kB + 1];           // whole units (32 windows) before record r
    long long nwin[kB];
    unsigned long long b0[kB];         // first 64-base block
    double inc[kB];
    int store[kB];                     // 0: the row is produced elsewhere (long record list)
};
static_assert(kTab * 4 == 128, "table rows are addressed by OR");

__device__ __forceinline__ void sts_zero_v2(uint32_t addr) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// 16 windows without a masked base and without a ragged end: no predicates
template <int K, bool kAligned = true, int KH = K>
__device__ __forceinline__ void count_chunk_full(uint32_t hist_addr, uint64_t x, uint32_t pass = 0) {
    constexpr uint32_t kMask = (1u << (2 * K)) - 1;
    constexpr uint32_t kOffMask = (((1u << (2 * KH)) - 1) >> 1) << 2;  // word offset of the histogram's 2 * KH key bits
    const uint64_t same = (x ^ (x << 2)) >> (64 - 2 * (16 + K - 2));
    if (same == 0) {  // homopolymer run: one add of 16 instead of 16 colliding adds
        const uint32_t kmer = (uint32_t)(x >> (64 - 2 * K)) & kMask;
        if (KH == K || (kmer >> (2 * KH)) == pass) {
            const uint32_t key = kmer & ((1u << (2 * KH)) - 1);
            red_add_shared_always(hist_addr + ((key >> 1) * 4), 16u << ((key & 1) * 16));
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t t = (uint32_t)(x >> (64 - 2 * (j + K) - 1));  // k-mer << 1
        const uint32_t addr = kAligned ? ((t & kOffMask) | hist_addr) : ((t & kOffMask) + hist_addr);
        if constexpr (KH == K) {
            red_add_shared_always(addr, (t & 2u) ? 0x10000u : 1u);
        } else {  // only the windows whose leading base(s) select this pass
            red_add_shared(addr, (t & 2u) ? 0x10000u : 1u, ((t >> (2 * KH + 1)) & ((1u << (2 * (K - KH))) - 1)) == pass);
        }
    }
}

// The count phase of one batch: out of line so that its registers (three code words and two mask words of the
// current and of the next unit, the unit bookkeeping) are allocated apart from the epilogue's, which holds the
// thread's slices of the mean / std / 1/std vectors; in one body the two sets together exceeded the 56 registers a
// thread may have with 2 x 544 threads per SM, and the compiler spilled inside both hot loops.
template <int K, int kB, int kW, int KH, bool kAligned>
__device__ __noinline__ void count_phase(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ mask,
                                         uint32_t mt_addr, uint32_t hist_addr, int tid, uint32_t pass) {
    // mt_addr: shared-window address of the batch's BatchMeta (explicit ld.shared: a generic reference would make
    // every access a generic load in an out-of-line function)
    using Meta = BatchMeta<kB>;
    constexpr uint32_t kHistBytes = (1u << (2 * KH)) * 2;
    constexpr uint32_t oPrefix = offsetof(Meta, prefix), oNwin = offsetof(Meta, nwin), oB0 = offsetof(Meta, b0);
    const uint32_t whole = lds_u32(mt_addr + oPrefix + 4 * kB);
    const uint32_t total = whole + kB;  // whole units first, then one tail slot per record
    // cursor of the incremental search (large batches): record cr holds the whole units [cbase, cnext)
    int cr = 0;
    uint32_t cbase = 0, cnext = lds_u32(mt_addr + oPrefix + 4);
    auto locate = [&](uint32_t g, int& r, uint32_t& u) {
        if (g >= whole) {  // tail slot of record g - whole: the unit after its whole units (may be empty)
            r = (int)(g - whole);
            u = (uint32_t)(lds_u64(mt_addr + oNwin + 8 * (uint32_t)r) >> 5);
            return;
        }
        if constexpr (kB <= 8) {
            r = 0;
            uint32_t before = 0;
#pragma unroll
            for (int i = 1; i < kB; ++i) {
                const uint32_t pf = lds_u32(mt_addr + oPrefix + 4 * i);  // broadcast shared loads
                if (g >= pf) { r = i; before = pf; }
            }
            u = g - before;
        } else {
            // a thread's unit indices only grow, so the record is found by walking on from the previous one
            while (g >= cnext) {  // g < whole = prefix[kB], so cr stays below kB
                ++cr;
                cbase = cnext;
                cnext = lds_u32(mt_addr + oPrefix + 4 * (uint32_t)(cr + 1));
            }
            r = cr;
            u = g - cbase;
        }
    };
    uint32_t w0 = 0, w1 = 0, w2 = 0, m0 = 0, m1 = 0, u = 0;
    int r = 0;
    uint32_t g = (uint32_t)tid;
    if (g < total) {
        locate(g, r, u);
        const unsigned long long b0 = lds_u64(mt_addr + oB0 + 8 * (uint32_t)r);
        const uint32_t* cw = codes + b0 * 4 + 2 * u;
        const uint32_t* mw = mask + b0 * 2 + u;
        w0 = __ldg(cw); w1 = __ldg(cw + 1); w2 = __ldg(cw + 2);
        m0 = __ldg(mw); m1 = __ldg(mw + 1);
    }
    while (g < total) {
        const uint32_t gn = g + kW;
        uint32_t n0 = 0, n1 = 0, n2 = 0, q0 = 0, q1 = 0, un = 0;
        int rn = 0;
        if (gn < total) {  // next unit's words are in flight while this one is counted
            locate(gn, rn, un);
            const unsigned long long b0 = lds_u64(mt_addr + oB0 + 8 * (uint32_t)rn);
            const uint32_t* cw = codes + b0 * 4 + 2 * un;
            const uint32_t* mw = mask + b0 * 2 + un;
            n0 = __ldg(cw); n1 = __ldg(cw + 1); n2 = __ldg(cw + 2);
            q0 = __ldg(mw); q1 = __ldg(mw + 1);
        }
        const uint32_t h = hist_addr + (uint32_t)r * kHistBytes;
        const uint64_t m64 = ((uint64_t)m0 << 32) | m1;
        if (g < whole && (m64 >> (64 - (32 + K - 1))) == 0) {
            count_chunk_full<K, kAligned, KH>(h, ((uint64_t)w0 << 32) | w1, pass);
            count_chunk_full<K, kAligned, KH>(h, ((uint64_t)w1 << 32) | w2, pass);
        } else {
            // tail slot: nwin % 32 windows, 0 = nothing to do
            const long long left = (long long)lds_u64(mt_addr + oNwin + 8 * (uint32_t)r) - (long long)u * 32;
            if (left > 0)
                count_chunk<K, kAligned, KH>(h, ((uint64_t)w0 << 32) | w1, (uint32_t)(m64 >> (64 - (16 + K - 1))),
                                             left < 16 ? (int)left : 16, pass);
            if (left > 16)
                count_chunk<K, kAligned, KH>(h, ((uint64_t)w1 << 32) | w2, (uint32_t)((m64 << 16) >> (64 - (16 + K - 1))),
                                             left < 32 ? (int)(left - 16) : 16, pass);
        }
        w0 = n0; w1 = n1; w2 = n2; m0 = q0; m1 = q1; u = un; r = rn; g = gn;
    }
}

// Epilogue flavours, so that the common cases carry no dead predicated instructions (the kernel is bound by
// instruction issue and the shared-memory pipe):
//   kBatchPlain  no vectors / fused Log2.post  (optional: column minima, column sums for the statistics),
//   kBatchFast   fp32 mean, std and 1/std, no column minima / fused Log2.post,
//   kBatchPost   kBatchFast + the Log2.post tail with a shift known before the launch (skr_post_spec),
//   kBatchAny    everything decided at run time.
enum { kBatchAny = 0, kBatchPlain = 1, kBatchFast = 2, kBatchPost = 3, kBatchAffine = 4 };

template <int K, bool kVecF64, int kMode, bool kMin, bool kColmin = false, bool kStats = false, bool kSplit = false>
__global__ void __launch_bounds__(BatchCfg<K, kSplit>::kThreads, BatchCfg<K, kSplit>::kCtasPerSm) count_batch_kernel(const CountParams p) {
    using Cfg = BatchCfg<K, kSplit>;
    constexpr int kB = Cfg::kB, kW = Cfg::kWorkers, kQ = Cfg::kQ, kG = Cfg::kG, kQc = Cfg::kQc, kChunks = Cfg::kChunks;
    static_assert(kChunks == 1 || (!kColmin && !kStats && kMode != kBatchAny),
                  "k = 7 runs the plain, fast and post flavours here; the others stay with the CTA-per-record kernel");
    constexpr bool kRegVec = kMode == kBatchFast || kMode == kBatchPost;  // -mean, -std, 1/std in registers
    constexpr bool kAffine = kMode == kBatchAffine;                       // a, b of the folded Log2.post tail in registers
    static_assert(!(kVecF64 && kMode != kBatchAny), "binary64 vectors take the generic epilogue");
    static_assert(!kColmin || kMode == kBatchPlain, "kColmin: column minima of the plain values (kBatchAny decides at run time)");
    static_assert(!kStats || kMode == kBatchPlain, "kStats: column sums of the plain values");
    static_assert(!(kMin && (kMode == kBatchPost || kAffine)), "the speculative Log2.post epilogue needs no running minimum");
    static_assert(!(kAffine && (kVecF64 || kColmin || kStats)), "the folded tail: fp32 vectors, nothing else fused");
    if (p.skip_flag && *p.skip_flag == p.skip_value) return;
    if (p.min_reset && blockIdx.x == 0 && threadIdx.x == 0) { p.min_reset->min_ordered = skr::ordered_encode(INFINITY); p.min_reset->nan_seen = 0; }
    extern __shared__ __align__(16) uint32_t smem_b[];
    __shared__ BatchMeta<kB> s_meta[2];
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr bool worker = true;                 // every thread takes part in the epilogue
    const bool counter = tid < Cfg::kCounters;    // count phase: warps 0..14 count, warp 15 prepares the next batch
    const int gi = kG > 1 ? tid / Cfg::kQuads : 0;    // epilogue group: records gi, gi + kG, ... of a batch
    const int qb = kG > 1 ? tid % Cfg::kQuads : tid;  // first bin quad of this thread
    const uint32_t raw_addr = skr::smem_u32(smem_b);
    const uint32_t hist_addr = Cfg::kAligned ? ((raw_addr + Cfg::kHistBytes - 1) & ~(uint32_t)(Cfg::kHistBytes - 1)) : raw_addr;

    float tmin = INFINITY;
    int tnan = 0;
    float shift = 0.0f;
    if ((kMode == kBatchAny || kMode == kBatchPost) && p.post_cell)
        shift = p.post_cell->nan_seen ? __int_as_float(0x7FC00000) : fabsf(skr::ordered_decode(p.post_cell->min_ordered));
    // speculative Log2.post: the thread that owns the arg-min column looks for a zero count there
    int zoff = -1;       // byte offset of that column's histogram word inside a record's histogram
    uint32_t zsh = 0;
    uint32_t zseen = 0;
    int zpass = 0;  // kSplit: the pass whose columns hold the arg-min column
    if ((kMode == kBatchAny || kMode == kBatchPost || kAffine) && p.spec) {
        int zc = p.spec->zero_col;
        if (zc >= 0) {
            zpass = zc / Cfg::kBins;
            zc -= zpass * Cfg::kBins;  // column inside the pass
            if (kG > 1 ? (zc >> 2) == qb : ((zc >> 2) % kW) == tid) { zoff = (zc >> 1) * 4; zsh = (uint32_t)(zc & 1) * 16u; }
        }
    }

    // this thread's slices of the vectors (fp32 vectors only; binary64 vectors are read in the epilogue)
    float4 mv[kQc], sv[kQc], yv[kQc];
    uint32_t cmin[kQc][4];
    float sx[kQc][4], sq[kQc][4];
#pragma unroll
    for (int j = 0; j < kQc; ++j) {
        if constexpr (!kRegVec && !kAffine) mv[j] = sv[j] = yv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) { cmin[j][e] = 0xFFFFFFFFu; sx[j][e] = 0.f; sq[j][e] = 0.f; }
    }
    // the thread's slices of the vectors for the quads of chunk `ch` (fp32 vectors)
    auto load_vectors = [&](int ch, int pass) {
#pragma unroll
        for (int j = 0; j < kQc; ++j) {
            const int q = pass * Cfg::kQuads + qb + (ch * kQc + j) * kW;
            if constexpr (kAffine) {  // mv = a, sv = b
                mv[j] = __ldg(reinterpret_cast<const float4*>(p.post_a) + q);
                sv[j] = __ldg(reinterpret_cast<const float4*>(p.post_b) + q);
            } else if constexpr (kRegVec) {  // all three vectors are there (dispatch): unconditional, negated once
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mean) + q);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.std_) + q);
                yv[j] = __ldg(reinterpret_cast<const float4*>(p.rstd) + q);
                mv[j] = make_float4(-m4.x, -m4.y, -m4.z, -m4.w);
                sv[j] = make_float4(-s4.x, -s4.y, -s4.z, -s4.w);
            } else {
                if (p.mean) mv[j] = __ldg(reinterpret_cast<const float4*>(p.mean) + q);
                if (p.std_) sv[j] = __ldg(reinterpret_cast<const float4*>(p.std_) + q);
                if (p.std_ && p.rstd) yv[j] = __ldg(reinterpret_cast<const float4*>(p.rstd) + q);
            }
        }
    };
    if (worker) {
        if constexpr (!kVecF64 && kMode != kBatchPlain && kChunks == 1) load_vectors(0, 0);  // once per kernel
        uint4* h4 = reinterpret_cast<uint4*>(smem_b + (hist_addr - raw_addr) / 4);
        for (int i = tid; i < kB * Cfg::kWords / 4; i += kW) h4[i] = make_uint4(0, 0, 0, 0);
    }

    // bookkeeping warp: fill s_meta[buf] for the next batch
    auto produce = [&](int buf) {
        BatchMeta<kB>& mt = s_meta[buf];
        long long b = 0;
        if (lane == 0) b = (long long)atomicAdd(p.work_counter, 1u);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        // kSplit: work item = (pass, batch), passes outermost
        const long long nbatches = (p.m + kB - 1) / kB;
        const int pass = Cfg::kPasses > 1 ? (int)(b / nbatches) : 0;
        if (Cfg::kPasses > 1) b = pass < Cfg::kPasses ? b - pass * nbatches : nbatches;
        const long long rec0 = b * kB;
        const int nrec = rec0 < p.m ? (int)min((long long)kB, p.m - rec0) : 0;
        uint32_t carry = 0;  // whole units of the records handled in earlier rounds
#pragma unroll 1
        for (int r0 = 0; r0 < kB; r0 += 32) {
            const int rr = r0 + lane;  // this lane's record of the round
            uint32_t units = 0;
            if (rr < kB) {
                long long nwin = 0;
                unsigned long long b0 = 0;
                int store = 0;
                if (rr < nrec) {
                    const long long rec = rec0 + rr;
                    nwin = (long long)__ldg(p.len + rec) - K + 1;
                    b0 = __ldg(p.blk_off + rec);
                    store = 1;
                    if (nwin > kLongWin) {
                        if (pass == 0) p.long_list[atomicAdd(p.long_count, 1u)] = (uint32_t)rec;  // listed once
                        store = 0;
                        nwin = 0;
                    }
                    if (nwin < 0) nwin = 0;
                }
                const double inc = nwin > 0 ? 1000.0 / (double)nwin : 0.0;
                mt.nwin[rr] = nwin;
                mt.b0[rr] = b0;
                mt.inc[rr] = inc;
                mt.store[rr] = store;
                units = (uint32_t)(nwin >> 5);  // whole units; the ragged rest is the record's tail unit
                if constexpr (kMode == kBatchAffine) {
                    // folded Log2.post tail: the value of a bin seen c times is c * inc in fp32 (one rounding of the
                    // product; 1.2e-7 from the chain below, far inside what the log2 behind it lets through), taken
                    // from the counter with the 2^23 trick -- no table, no shuffles in the epilogue
                    const float incf = __double2float_rn(inc);
                    mt.tab[rr][0] = incf;
                    mt.tab[rr][1] = -8388608.0f * incf;  // exact (a power of two)
                } else {
                double acc = 0.0;  // the literal chain of kmer_counts.py:144-150 for counts below kTab
                mt.tab[rr][0] = p.log2_pre ? log2f(1.0f) : 0.0f;
#pragma unroll 8
                for (int c = 1; c < kTab; ++c) {
                    acc = __dadd_rn(acc, inc);
                    float v = __double2float_rn(acc);
                    if (p.log2_pre) v = log2f(__fadd_rn(v, 1.0f));
                    mt.tab[rr][c] = v;
                }
                }
            }
            uint32_t incl = units;  // inclusive scan over the lanes of the round
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += up;
            }
            if (rr < kB) mt.prefix[rr + 1] = carry + incl;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) {
            mt.prefix[0] = 0;
            mt.rec0 = nrec > 0 ? rec0 : -1;
            mt.nrec = nrec;
            mt.pass = pass;
        }
    };

    if (!counter) produce(0);
    __syncthreads();

    for (int it = 0;; ++it) {
        const BatchMeta<kB>& mt = s_meta[it & 1];
        if (mt.rec0 < 0) break;
        if (!counter) {
            produce((it + 1) & 1);
        } else {
            count_phase<K, kB, Cfg::kCounters, Cfg::kHK, Cfg::kAligned>(p.codes, p.mask, skr::smem_u32(&mt), hist_addr, tid, (uint32_t)mt.pass);
        }
        __syncthreads();
        if (worker) {
            // ---- epilogue ----
            const int nrec = mt.nrec;
#pragma unroll 1
            for (int ch = 0; ch < kChunks; ++ch) {
            if constexpr (!kVecF64 && kMode != kBatchPlain && kChunks > 1) load_vectors(ch, mt.pass);  // once per batch and chunk
            for (int r = gi; r < nrec; r += kG) {
                if (!mt.store[r]) continue;
                const uint32_t tab_addr = skr::smem_u32(&mt.tab[r][0]);
                // lane c holds the value of a bin seen c times (kAffine: lanes 0 / 1 hold inc and -2^23 inc)
                const float treg = lds_f32(tab_addr + (kAffine ? 0u : (uint32_t)lane * 4));
                const float tneg = kAffine ? lds_f32(tab_addr + 4) : 0.0f;
                const uint32_t hrec = hist_addr + (uint32_t)r * Cfg::kHistBytes;
                float* __restrict__ orow = reinterpret_cast<float*>(p.out) + (size_t)(mt.rec0 + r) * (size_t)p.ld_out +
                                           (Cfg::kPasses > 1 ? (size_t)mt.pass * Cfg::kBins : 0);
                if (ch == 0 && zoff >= 0 && (Cfg::kPasses == 1 || mt.pass == zpass) &&
                    ((lds_u32(hrec + (uint32_t)zoff) >> zsh) & 0xFFFFu) == 0)
                    zseen = 1;
#pragma unroll
                for (int j = 0; j < kQc; ++j) {
                    const int q = qb + (ch * kQc + j) * kW;
                    const uint32_t a = hrec + (uint32_t)q * 8;
                    const uint2 v = lds_v2(a);
                    sts_zero_v2(a);
                    float x[4];
                    if constexpr (kAffine) {
                        // 0x4B000000 | c is the float 2^23 + c: (2^23 + c) * inc - 2^23 * inc = RN(c * inc) in one FFMA2
                        // per two columns, then the folded tail in another, then the hardware log2
                        const uint64_t inc2 = skr::f2_pack(treg, treg), neg2 = skr::f2_pack(tneg, tneg);
                        const uint64_t m0 = skr::f2_pack(__uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7610)),
                                                         __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7632)));
                        const uint64_t m1 = skr::f2_pack(__uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7610)),
                                                         __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7632)));
                        const uint64_t z0 = skr::f2_fma(skr::f2_fma(m0, inc2, neg2), skr::f2_pack(mv[j].x, mv[j].y), skr::f2_pack(sv[j].x, sv[j].y));
                        const uint64_t z1 = skr::f2_fma(skr::f2_fma(m1, inc2, neg2), skr::f2_pack(mv[j].z, mv[j].w), skr::f2_pack(sv[j].z, sv[j].w));
                        skr::f2_unpack(z0, x[0], x[1]);
                        skr::f2_unpack(z1, x[2], x[3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = log2_post(x[e]);
                    } else {
                    // indexed shuffles take the source lane modulo 32; counts of kTab and more are fixed up below
                    x[0] = __shfl_sync(0xFFFFFFFFu, treg, (int)v.x);
                    x[1] = __shfl_sync(0xFFFFFFFFu, treg, (int)(v.x >> 16));
                    x[2] = __shfl_sync(0xFFFFFFFFu, treg, (int)v.y);
                    x[3] = __shfl_sync(0xFFFFFFFFu, treg, (int)(v.y >> 16));
                    if (((v.x | v.y) & ~(uint32_t)((kTab - 1) * 0x10001u)) != 0) {
                        const uint32_t c4[4] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16};
                        const double inc = mt.inc[r];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c4[e] >= kTab) x[e] = slow_bin_value(inc, c4[e], p.log2_pre);
                    }
                    }
                    if constexpr (kStats) {  // column sums of the plain values (accurate norm_vectors)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            sx[j][e] = __fadd_rn(sx[j][e], x[e]);
                            sq[j][e] = __fmaf_rn(x[e], x[e], sq[j][e]);
                        }
                    }
                    if constexpr (kAffine) {
                        // (done above)
                    } else if constexpr (kRegVec) {
                        // mv / sv hold -mean / -std in these flavours (negated once, after the load); FADD2 / FMUL2 /
                        // FFMA2 do two columns per issue slot
                        uint64_t z0 = sub_div_by_rcp2(skr::f2_pack(x[0], x[1]), skr::f2_pack(mv[j].x, mv[j].y),
                                                      skr::f2_pack(sv[j].x, sv[j].y), skr::f2_pack(yv[j].x, yv[j].y));
                        uint64_t z1 = sub_div_by_rcp2(skr::f2_pack(x[2], x[3]), skr::f2_pack(mv[j].z, mv[j].w),
                                                      skr::f2_pack(sv[j].z, sv[j].w), skr::f2_pack(yv[j].z, yv[j].w));
                        if constexpr (kMode == kBatchPost) {  // (z + |min|) + 1: two roundings, kmer_counts.py:208-209
                            const uint64_t s2 = skr::f2_pack(shift, shift), one2 = skr::f2_pack(1.0f, 1.0f);
                            z0 = skr::f2_add(skr::f2_add(z0, s2), one2);
                            z1 = skr::f2_add(skr::f2_add(z1, s2), one2);
                        }
                        skr::f2_unpack(z0, x[0], x[1]);
                        skr::f2_unpack(z1, x[2], x[3]);
                        if constexpr (kMode == kBatchPost) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = log2_post(x[e]);
                        }
                    } else if constexpr (kMode == kBatchAny) {
                        if constexpr (kVecF64) {
                            if (p.mean) {
                                const double* mp = reinterpret_cast<const double*>(p.mean) + 4 * q;
#pragma unroll
                                for (int e = 0; e < 4; ++e) x[e] = __double2float_rn(__dsub_rn((double)x[e], __ldg(mp + e)));
                            }
                            if (p.std_) {
                                const double* sp = reinterpret_cast<const double*>(p.std_) + 4 * q;
#pragma unroll
                                for (int e = 0; e < 4; ++e) x[e] = __double2float_rn(__ddiv_rn((double)x[e], __ldg(sp + e)));
                            }
                        } else {
                            if (p.mean) {
                                x[0] = __fsub_rn(x[0], mv[j].x); x[1] = __fsub_rn(x[1], mv[j].y);
                                x[2] = __fsub_rn(x[2], mv[j].z); x[3] = __fsub_rn(x[3], mv[j].w);
                            }
                            if (p.std_) {
                                if (p.rstd) {
                                    x[0] = div_by_rcp(x[0], sv[j].x, yv[j].x); x[1] = div_by_rcp(x[1], sv[j].y, yv[j].y);
                                    x[2] = div_by_rcp(x[2], sv[j].z, yv[j].z); x[3] = div_by_rcp(x[3], sv[j].w, yv[j].w);
                                } else {
                                    x[0] = __fdiv_rn(x[0], sv[j].x); x[1] = __fdiv_rn(x[1], sv[j].y);
                                    x[2] = __fdiv_rn(x[2], sv[j].z); x[3] = __fdiv_rn(x[3], sv[j].w);
                                }
                            }
                        }
                        if (p.colmin) {  // values are >= 0 on this path: float bits order like unsigned integers
#pragma unroll
                            for (int e = 0; e < 4; ++e) cmin[j][e] = min(cmin[j][e], __float_as_uint(x[e]));
                        }
                        if (p.post_cell) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = log2_post(__fadd_rn(__fadd_rn(x[e], shift), 1.0f));
                        }
                    }
                    if constexpr (kColmin) {  // plain values are >= 0: float bits order like unsigned integers
#pragma unroll
                        for (int e = 0; e < 4; ++e) cmin[j][e] = min(cmin[j][e], __float_as_uint(x[e]));
                    }
                    if constexpr (kMin) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) skr::min_update(x[e], tmin, tnan);
                    }
                    if (kMode != kBatchAny || !p.no_store)
                        reinterpret_cast<float4*>(orow)[q] = make_float4(x[0], x[1], x[2], x[3]);
                }
            }
            }
        }
        __syncthreads();
    }

    if (worker) {
        if (zseen) p.spec->zero_seen = p.spec_epoch;
        if (kColmin || (kMode == kBatchAny && p.colmin)) {
#pragma unroll
            for (int j = 0; j < kQc; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = 4 * (qb + j * kW) + e;
                    if (cmin[j][e] < p.colmin[col]) atomicMin(&p.colmin[col], cmin[j][e]);
                }
        }
        if constexpr (kStats) {
#pragma unroll
            for (int j = 0; j < kQc; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = 4 * (qb + j * kW) + e;
                    atomicAdd(&p.colsum[col], (double)sx[j][e]);
                    atomicAdd(&p.colsq[col], (double)sq[j][e]);
                }
        }
        if constexpr (kMin) {
            if (tmin != tmin) { tnan = 1; tmin = INFINITY; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                tmin = fminf(tmin, __shfl_xor_sync(0xFFFFFFFFu, tmin, o));
                tnan |= __shfl_xor_sync(0xFFFFFFFFu, tnan, o);
            }
            if (lane == 0) {
                if (tmin < INFINITY) atomicMin(&p.min_cell->min_ordered, skr::ordered_encode(tmin));
                if (tnan) atomicOr(&p.min_cell->nan_seen, 1u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// count_ws_kernel: the folded Log2.post flavour with the two phases of count_batch_kernel given to different warps.
// In count_batch_kernel every warp counts, meets the others at a barrier, runs its share of the epilogue and meets
// them again: a quarter of the warp time is spent at those two barriers (ncu, profiles/r02_ncu_count_details.txt).
// Here warps 0-6 only count, warp 7 only keeps the books and warps 8-15 only run the epilogue; two histogram sets
// of kB records circulate between them through mbarriers (meta ready -> counted -> empty), so the counters fill
// one set while the epilogue drains the other and nobody waits for a phase change.  An epilogue thread owns four
// bin quads of every record (1024 quads / 256 threads); that is affordable only because the folded tail needs two
// vectors (a, b: 32 registers) instead of three.  Raw counts (kPlain: table lookups, no vectors) run here too; the
// flavour with per-thread column sums was tried and is slower than in count_batch_kernel (0.266 against 0.223 ms:
// 32 more registers spill and the eight epilogue warps become the bottleneck).
template <int K, int kB_ = 4, int kSets_ = 2>
struct WsCfg {
    static constexpr int kBins = 1 << (2 * K);
    static constexpr int kHistBytes = kBins * 2;
    static constexpr int kQuads = kBins / 4;
    static constexpr int kB = kB_;    // records per histogram set
    static constexpr int kSets = kSets_;
    static constexpr int kThreads = 512;
    static constexpr int kCountWarps = 7;
    static constexpr int kCounters = kCountWarps * 32;
    static constexpr int kEpi = 256;
    static constexpr int kEpiWarps = kEpi / 32;
    static constexpr int kQc = kQuads / kEpi;
    static constexpr size_t kSmem = (size_t)(kSets * kB + 1) * kHistBytes;  // + one histogram of alignment slack
    static_assert(kQuads % kEpi == 0 && kQc >= 1 && kQc <= 4, "whole quads per epilogue thread, vectors in registers");
};

template <int K, int kB_, int kSets_, bool kPlain = false>
__global__ void __launch_bounds__(WsCfg<K, kB_, kSets_>::kThreads, 2) count_ws_kernel(const CountParams p) {
    using Cfg = WsCfg<K, kB_, kSets_>;
    constexpr int kB = Cfg::kB;
    constexpr int kSets = Cfg::kSets;
    if (p.skip_flag && *p.skip_flag == p.skip_value) return;
    if (p.min_reset && blockIdx.x == 0 && threadIdx.x == 0) { p.min_reset->min_ordered = skr::ordered_encode(INFINITY); p.min_reset->nan_seen = 0; }
    extern __shared__ __align__(16) uint32_t smem_b[];
    __shared__ BatchMeta<kB> s_meta[kSets];
    __shared__ __align__(8) uint64_t s_full[kSets], s_counted[kSets], s_empty[kSets];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t raw_addr = skr::smem_u32(smem_b);
    const uint32_t hist_addr = (raw_addr + Cfg::kHistBytes - 1) & ~(uint32_t)(Cfg::kHistBytes - 1);
    {
        uint4* h4 = reinterpret_cast<uint4*>(smem_b + (hist_addr - raw_addr) / 4);
        for (int i = tid; i < Cfg::kSets * kB * Cfg::kHistBytes / 16; i += Cfg::kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int s = 0; s < kSets; ++s) {
            skr::mbar_init(&s_full[s], 1);
            skr::mbar_init(&s_counted[s], Cfg::kCountWarps);
            skr::mbar_init(&s_empty[s], Cfg::kEpiWarps);
        }
        skr::fence_mbar_init();
    }
    __syncthreads();

    if (warp < Cfg::kCountWarps) {
        // ---- counters ----
        for (int it = 0;; ++it) {
            const int s = it % kSets;
            skr::mbar_wait(&s_full[s], (uint32_t)(it / kSets) & 1u);
            if (s_meta[s].rec0 < 0) break;
            count_phase<K, kB, Cfg::kCounters, K, true>(p.codes, p.mask, skr::smem_u32(&s_meta[s]),
                                                       hist_addr + (uint32_t)(s * kB) * Cfg::kHistBytes, tid, 0u);
            __syncwarp();
            if (lane == 0) skr::mbar_arrive(&s_counted[s]);
        }
    } else if (warp == Cfg::kCountWarps) {
        // ---- bookkeeping ----
        for (int it = 0;; ++it) {
            const int s = it % kSets;
            if (it >= kSets) skr::mbar_wait(&s_empty[s], (uint32_t)(it / kSets - 1) & 1u);  // the epilogue is done with this slot
            BatchMeta<kB>& mt = s_meta[s];
            long long b = 0;
            if (lane == 0) b = (long long)atomicAdd(p.work_counter, 1u);
            b = __shfl_sync(0xFFFFFFFFu, b, 0);
            const long long rec0 = b * kB;
            const int nrec = rec0 < p.m ? (int)min((long long)kB, p.m - rec0) : 0;
            uint32_t units = 0;
            if (lane < kB) {
                long long nwin = 0;
                unsigned long long b0 = 0;
                int store = 0;
                if (lane < nrec) {
                    const long long rec = rec0 + lane;
                    nwin = (long long)__ldg(p.len + rec) - K + 1;
                    b0 = __ldg(p.blk_off + rec);
                    store = 1;
                    if (nwin > kLongWin) {
                        p.long_list[atomicAdd(p.long_count, 1u)] = (uint32_t)rec;
                        store = 0;
                        nwin = 0;
                    }
                    if (nwin < 0) nwin = 0;
                }
                const double inc = nwin > 0 ? 1000.0 / (double)nwin : 0.0;
                mt.nwin[lane] = nwin;
                mt.b0[lane] = b0;
                mt.inc[lane] = inc;
                mt.store[lane] = store;
                if constexpr (kPlain) {
                    double acc = 0.0;  // the literal chain of kmer_counts.py:144-150 for counts below kTab
                    mt.tab[lane][0] = p.log2_pre ? log2f(1.0f) : 0.0f;
#pragma unroll 8
                    for (int c = 1; c < kTab; ++c) {
                        acc = __dadd_rn(acc, inc);
                        float v = __double2float_rn(acc);
                        if (p.log2_pre) v = log2f(__fadd_rn(v, 1.0f));
                        mt.tab[lane][c] = v;
                    }
                } else {
                    const float incf = __double2float_rn(inc);
                    mt.tab[lane][0] = incf;
                    mt.tab[lane][1] = -8388608.0f * incf;
                }
                units = (uint32_t)(nwin >> 5);
            }
            uint32_t incl = units;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane < kB) mt.prefix[lane + 1] = incl;
            if (lane == 0) {
                mt.prefix[0] = 0;
                mt.rec0 = nrec > 0 ? rec0 : -1;
                mt.nrec = nrec;
                mt.pass = 0;
            }
            __syncwarp();
            if (lane == 0) skr::mbar_arrive(&s_full[s]);
            if (nrec == 0) break;
        }
    } else {
        // ---- epilogue ----
        const int et = tid - (Cfg::kThreads - Cfg::kEpi);
        float4 av[Cfg::kQc], bv[Cfg::kQc];
        if constexpr (!kPlain) {
#pragma unroll
            for (int j = 0; j < Cfg::kQc; ++j) {
                av[j] = __ldg(reinterpret_cast<const float4*>(p.post_a) + et + j * Cfg::kEpi);
                bv[j] = __ldg(reinterpret_cast<const float4*>(p.post_b) + et + j * Cfg::kEpi);
            }
        }
        int zoff = -1;
        uint32_t zsh = 0, zseen = 0;
        if (!kPlain && p.spec) {
            const int zc = p.spec->zero_col;
            if (zc >= 0 && ((zc >> 2) % Cfg::kEpi) == et) { zoff = (zc >> 1) * 4; zsh = (uint32_t)(zc & 1) * 16u; }
        }
        for (int it = 0;; ++it) {
            const int s = it % kSets;
            const uint32_t ph = (uint32_t)(it / kSets) & 1u;
            skr::mbar_wait(&s_full[s], ph);
            const BatchMeta<kB>& mt = s_meta[s];
            if (mt.rec0 < 0) break;
            skr::mbar_wait(&s_counted[s], ph);
            const int nrec = mt.nrec;
            for (int r = 0; r < nrec; ++r) {
                if (!mt.store[r]) continue;
                // plain: lane c holds the value of a bin seen c times; folded: lanes 0 / 1 of the row hold inc, -2^23 inc
                const float incf = mt.tab[r][kPlain ? lane : 0], nincf = kPlain ? 0.0f : mt.tab[r][1];
                const uint64_t inc2 = skr::f2_pack(incf, incf), neg2 = skr::f2_pack(nincf, nincf);
                const uint32_t hrec = hist_addr + (uint32_t)(s * kB + r) * Cfg::kHistBytes;
                float* __restrict__ orow = reinterpret_cast<float*>(p.out) + (size_t)(mt.rec0 + r) * (size_t)p.ld_out;
                if (zoff >= 0 && ((lds_u32(hrec + (uint32_t)zoff) >> zsh) & 0xFFFFu) == 0) zseen = 1;
#pragma unroll
                for (int j = 0; j < Cfg::kQc; ++j) {
                    const int q = et + j * Cfg::kEpi;
                    const uint32_t a = hrec + (uint32_t)q * 8;
                    const uint2 v = lds_v2(a);
                    sts_zero_v2(a);
                    float x[4];
                    if constexpr (kPlain) {
                        // indexed shuffles take the source lane modulo 32; counts of kTab and more are fixed up below
                        x[0] = __shfl_sync(0xFFFFFFFFu, incf, (int)v.x);
                        x[1] = __shfl_sync(0xFFFFFFFFu, incf, (int)(v.x >> 16));
                        x[2] = __shfl_sync(0xFFFFFFFFu, incf, (int)v.y);
                        x[3] = __shfl_sync(0xFFFFFFFFu, incf, (int)(v.y >> 16));
                        if (((v.x | v.y) & ~(uint32_t)((kTab - 1) * 0x10001u)) != 0) {
                            const uint32_t c4[4] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16};
                            const double inc = mt.inc[r];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c4[e] >= kTab) x[e] = slow_bin_value(inc, c4[e], p.log2_pre);
                        }
                        reinterpret_cast<float4*>(orow)[q] = make_float4(x[0], x[1], x[2], x[3]);
                    } else {
                    const uint64_t m0 = skr::f2_pack(__uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7610)),
                                                     __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7632)));
                    const uint64_t m1 = skr::f2_pack(__uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7610)),
                                                     __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7632)));
                    const uint64_t z0 = skr::f2_fma(skr::f2_fma(m0, inc2, neg2), skr::f2_pack(av[j].x, av[j].y), skr::f2_pack(bv[j].x, bv[j].y));
                    const uint64_t z1 = skr::f2_fma(skr::f2_fma(m1, inc2, neg2), skr::f2_pack(av[j].z, av[j].w), skr::f2_pack(bv[j].z, bv[j].w));
                    skr::f2_unpack(z0, x[0], x[1]);
                    skr::f2_unpack(z1, x[2], x[3]);
                    reinterpret_cast<float4*>(orow)[q] = make_float4(log2_post(x[0]), log2_post(x[1]), log2_post(x[2]), log2_post(x[3]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) skr::mbar_arrive(&s_empty[s]);
        }
        if (zseen) p.spec->zero_seen = p.spec_epoch;
    }
}

template <int K, int kB_, int kSets_, bool kPlain = false>
int launch_ws(const CountParams& wp, int sms, cudaStream_t stream) {
    using W = WsCfg<K, kB_, kSets_>;
    auto kern = count_ws_kernel<K, kB_, kSets_, kPlain>;
    SKR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::kSmem));
    int per_sm = 0;
    SKR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W::kThreads, W::kSmem));
    if (per_sm < 1) return skr::fail(SKR_ERR_CUDA, "count_ws_kernel for k=%d does not fit on this device", K);
    long long grid = (long long)sms * per_sm;
    const long long need = (wp.m + W::kB - 1) / W::kB;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, W::kThreads, W::kSmem, stream>>>(wp);
    return SKR_OK;
}

template <int K, bool kVecF64, int kMode, bool kMin, bool kColmin = false, bool kStats = false, bool kSplit = false>
int launch_batch(const CountParams& wp, int sms, cudaStream_t stream) {
    using B = BatchCfg<K, kSplit>;
    auto bkern = count_batch_kernel<K, kVecF64, kMode, kMin, kColmin, kStats, kSplit>;
    SKR_CUDA_CHECK(cudaFuncSetAttribute(bkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::kSmem));
    int bper_sm = 0;
    SKR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bper_sm, bkern, B::kThreads, B::kSmem));
    if (bper_sm < 1) return skr::fail(SKR_ERR_CUDA, "batch count kernel for k=%d does not fit on this device", K);
    long long bgrid = (long long)sms * bper_sm;
    const long long bneed = (wp.m + B::kB - 1) / B::kB * B::kPasses;
    if (bgrid > bneed) bgrid = bneed;
    bkern<<<(unsigned)bgrid, B::kThreads, B::kSmem, stream>>>(wp);
    return SKR_OK;
}

// k = 7: the fast and post flavours only (vectors in registers, two chunk passes); false = not handled here
template <int K>
bool batch_handles(const CountParams& wp, int vec_is_f64) {
    if (K <= 6) return true;
    if (wp.colmin || wp.no_store || wp.colsum) return false;
    if (!wp.mean && !wp.std_ && !wp.post_cell) return true;  // plain counts
    return !vec_is_f64 && wp.mean && wp.std_ && wp.rstd && !(wp.post_cell && wp.min_cell);
}

template <int K, bool kVecF64>
int dispatch_batch(const CountParams& wp, int sms, cudaStream_t stream) {
    if constexpr (K >= 7) {
        if constexpr (kVecF64) {
            return skr::fail(SKR_ERR_ARG, "skr_count: internal dispatch error");
        } else {
            const bool mn7 = wp.min_cell != nullptr;
            if (!wp.mean && !wp.std_ && !wp.post_cell)
                return mn7 ? launch_batch<K, false, kBatchPlain, true>(wp, sms, stream)
                           : launch_batch<K, false, kBatchPlain, false>(wp, sms, stream);
            constexpr bool kSp = K == 8;  // k = 8 with vectors: four column passes over a k = 7 sized histogram
            if (wp.post_cell && wp.post_a && !wp.log2_pre) return launch_batch<K, false, kBatchAffine, false, false, false, kSp>(wp, sms, stream);
            if (wp.post_cell) return launch_batch<K, false, kBatchPost, false, false, false, kSp>(wp, sms, stream);
            return mn7 ? launch_batch<K, false, kBatchFast, true, false, false, kSp>(wp, sms, stream)
                       : launch_batch<K, false, kBatchFast, false, false, false, kSp>(wp, sms, stream);
        }
    } else {
    const bool plain = !wp.mean && !wp.std_ && !wp.post_cell && !wp.no_store;
    const bool regvec = !kVecF64 && wp.mean && wp.std_ && wp.rstd && !wp.colmin && !wp.no_store;
    const bool fast = regvec && !wp.post_cell;
    const bool post = regvec && wp.post_cell && !wp.min_cell;
    const bool mn = wp.min_cell != nullptr;
    if constexpr (!kVecF64) {
        if (plain && wp.colsum) {
            if (mn) return skr::fail(SKR_ERR_ARG, "skr_count: column sums and a running minimum are not combined");
            return wp.colmin ? launch_batch<K, false, kBatchPlain, false, true, true>(wp, sms, stream)
                             : launch_batch<K, false, kBatchPlain, false, false, true>(wp, sms, stream);
        }
        if (plain && wp.colmin) return mn ? launch_batch<K, false, kBatchPlain, true, true>(wp, sms, stream)
                                          : launch_batch<K, false, kBatchPlain, false, true>(wp, sms, stream);
        if (plain && !mn) {
            if constexpr (K == 6) {  // raw counts: the warp-specialised kernel (the table lookups need no vector registers)
                const char* e = getenv("SEEKR_B200_COUNT_WS");  // read per launch: tests switch it
                if (!e || atoi(e) != 0) return launch_ws<K, 4, 3, true>(wp, sms, stream);
            }
        }
        if (plain) return mn ? launch_batch<K, false, kBatchPlain, true>(wp, sms, stream)
                             : launch_batch<K, false, kBatchPlain, false>(wp, sms, stream);
        if (fast) return mn ? launch_batch<K, false, kBatchFast, true>(wp, sms, stream)
                            : launch_batch<K, false, kBatchFast, false>(wp, sms, stream);
        if (post && wp.post_a && !wp.log2_pre) {
            if constexpr (K == 6) {
                // three sets of four records measured best (S50k, alone): 0.178 ms against 0.195 (two sets of four),
                // 0.184 (two of six), 0.180 (four of three), 0.193 (six of two) and 0.199 for count_batch_kernel
                const char* e = getenv("SEEKR_B200_COUNT_WS");  // read per launch: tests switch it
                const int ws = e ? atoi(e) : 3;
                if (ws == 1) return launch_ws<K, 4, 2>(wp, sms, stream);
                if (ws == 2) return launch_ws<K, 6, 2>(wp, sms, stream);
                if (ws == 3) return launch_ws<K, 4, 3>(wp, sms, stream);
                if (ws == 4) return launch_ws<K, 3, 4>(wp, sms, stream);
            }
            return launch_batch<K, false, kBatchAffine, false>(wp, sms, stream);
        }
        if (post) return launch_batch<K, false, kBatchPost, false>(wp, sms, stream);
    }
    if (wp.colsum) return skr::fail(SKR_ERR_ARG, "skr_count: column sums go with plain counts (no vectors, no Log2.post)");
    return mn ? launch_batch<K, kVecF64, kBatchAny, true>(wp, sms, stream)
              : launch_batch<K, kVecF64, kBatchAny, false>(wp, sms, stream);
    }
}

// ---------------------------------------------------------------------------------------------
// per-device scratch: work counters (a ring, so launches on different streams do not share one).
// The per-CTA spill rows and the long-record list are stream-ordered allocations of each launch
// (cudaMallocAsync / cudaFreeAsync), so concurrent launches on different streams never share them;
// the default memory pool is told to keep its memory across synchronisations.
// ---------------------------------------------------------------------------------------------
struct DeviceScratch {
    unsigned int* counters = nullptr;
    int next_counter = 0;
};
constexpr int kCounterRing = 1024;   // 4 counters per launch
constexpr int kCounterBlock = 256;   // zeroed together, on the stream that is about to use them
std::mutex g_scratch_mu;
std::map<std::pair<int, cudaStream_t>, DeviceScratch> g_scratch;  // one ring per (device, stream)

// Four zeroed work counters for a launch on `stream`.  A ring belongs to one stream, so everything that touches it
// is ordered by the stream: a block of 64 launches' worth of counters is cleared by ONE memset when the ring enters
// it (its previous users are 192+ launches behind on the same stream), instead of one memset per launch.
int next_counters(int dev, cudaStream_t stream, unsigned int** out) {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    DeviceScratch& s = g_scratch[std::make_pair(dev, stream)];
    if (!s.counters) {
        SKR_CUDA_CHECK(cudaMalloc(&s.counters, kCounterRing * sizeof(unsigned int)));
        cudaMemPool_t pool;
        SKR_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        SKR_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    if (s.next_counter % kCounterBlock == 0)
        SKR_CUDA_CHECK(cudaMemsetAsync(s.counters + s.next_counter, 0, kCounterBlock * sizeof(unsigned int), stream));
    *out = s.counters + s.next_counter;
    s.next_counter = (s.next_counter + 4) % kCounterRing;
    return SKR_OK;
}

// zero_row[q] = what the epilogue makes of four empty bins in columns 4q .. 4q + 3: the same device functions, so
// the same bits (value of a zero count is 0, also after log2(0 + 1))
template <bool kVecF64>
__global__ void __launch_bounds__(256) zero_row_kernel(const CountParams p, float4* zero_row, int quads) {
    if (p.skip_flag && *p.skip_flag == p.skip_value) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    normalize4<kVecF64>(p, q, r);
    if (p.post_cell) {
        const float shift = p.post_cell->nan_seen ? __int_as_float(0x7FC00000) : fabsf(skr::ordered_decode(p.post_cell->min_ordered));
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = log2_post(__fadd_rn(__fadd_rn(r[e], shift), 1.0f));
    }
    zero_row[q] = make_float4(r[0], r[1], r[2], r[3]);
}

template <int K, bool kVecF64, typename OutT>
int launch_count(CountParams p, cudaStream_t stream) {
    using Cfg = CountCfg<K>;
    auto kern = count_kernel<K, kVecF64, OutT>;
    int dev = 0, sms = 0;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    SKR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SKR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    int per_sm = 0;
    SKR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::kThreads, Cfg::kSmem));
    if (per_sm < 1) return skr::fail(SKR_ERR_CUDA, "count kernel for k=%d does not fit on this device", K);
    // the team / batch kernels take k <= 6 with float output (and k = 7 with fp32 vectors in the fast / post
    // flavours); they leave records that are too long for them on a list which the CTA kernel then drains
    constexpr bool kTeamK = K <= 6 && sizeof(OutT) == 4;
    constexpr bool kBatchOnly = K >= 7 && sizeof(OutT) == 4 && !kVecF64;
    const bool use_list = kTeamK || (kBatchOnly && batch_handles<K>(p, kVecF64));
    if (use_list && p.m > 0xFFFFFFFFll) return skr::fail(SKR_ERR_ARG, "skr_count: more than 2^32 records");
    unsigned int* ctr = nullptr;
    int rc = next_counters(dev, stream, &ctr);
    if (rc != SKR_OK) return rc;
    // no record is too long for the team / batch kernels (the caller knows the longest record): nothing to hand over
    const bool no_long = use_list && p.max_length > 0 && (long long)p.max_length - K + 1 <= kLongWin;
    uint32_t* long_list = nullptr;
    if (use_list && !no_long) SKR_CUDA_CHECK(cudaMallocAsync(&long_list, (size_t)p.m * sizeof(uint32_t), stream));
    long long grid = (long long)sms * per_sm;
    if (use_list) {
        CountParams wp = p;
        wp.work_counter = ctr;
        wp.long_list = long_list;
        wp.long_count = ctr + 1;
        bool batch = false;
        if constexpr (kBatchOnly || (kTeamK && K >= 4)) {
            const char* env = getenv("SEEKR_B200_COUNT_KERNEL");  // experiment knob: "warp" selects the team-per-record kernel
            batch = kBatchOnly || !(env && env[0] == 'w');
            if (batch) {
                rc = dispatch_batch<K, kVecF64>(wp, sms, stream);
                if (rc != SKR_OK) return rc;
            }
        }
        if constexpr (kTeamK) {
            if (!batch) {
                using W = WarpCfg<K>;
                auto wkern = count_warp_kernel<K, kVecF64>;
                SKR_CUDA_CHECK(cudaFuncSetAttribute(wkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::kSmem));
                if (const char* env = getenv("SEEKR_B200_COUNT_CARVEOUT"))  // experiment knob: % of the SM's 228 KB given to smem
                    SKR_CUDA_CHECK(cudaFuncSetAttribute(wkern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env)));
                int wper_sm = 0;
                SKR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wper_sm, wkern, W::kThreads, W::kSmem));
                if (wper_sm < 1) return skr::fail(SKR_ERR_CUDA, "warp count kernel for k=%d does not fit on this device", K);
                if (const char* env = getenv("SEEKR_B200_COUNT_CTAS_PER_SM"))  // experiment knob: resident CTAs per SM
                    if (atoi(env) > 0 && atoi(env) < wper_sm) wper_sm = atoi(env);
                long long wgrid = (long long)sms * wper_sm;
                const long long need = (p.m + W::kTeams - 1) / W::kTeams;
                if (wgrid > need) wgrid = need;
                wkern<<<(unsigned)wgrid, W::kThreads, W::kSmem, stream>>>(wp);
            }
        }
        SKR_LAUNCH_CHECK();
        if (no_long) return SKR_OK;
        // long records (rare): the CTA kernel reads the list length on the device
        p.work_counter = ctr + 2;
        p.long_list = long_list;
        p.long_count = ctr + 1;
        if (grid > sms) grid = sms;
    } else {
        p.work_counter = ctr;
        p.long_list = nullptr;
        p.long_count = nullptr;
        if (grid > p.m) grid = p.m;
    }
    SKR_CUDA_CHECK(cudaMallocAsync(&p.spill, (size_t)grid * Cfg::kBins * sizeof(uint32_t), stream));
    float4* zero_row = nullptr;
    if constexpr (K >= 7 && sizeof(OutT) == 4) {
        if (!p.colmin && (p.mean || p.std_ || p.post_cell)) {
            SKR_CUDA_CHECK(cudaMallocAsync(&zero_row, (size_t)Cfg::kBins * sizeof(float), stream));
            zero_row_kernel<kVecF64><<<(Cfg::kBins / 4 + 255) / 256, 256, 0, stream>>>(p, zero_row, Cfg::kBins / 4);
            SKR_LAUNCH_CHECK();
            p.zero_row = zero_row;
        }
    }
    kern<<<(unsigned)grid, Cfg::kThreads, Cfg::kSmem, stream>>>(p);
    SKR_LAUNCH_CHECK();
    if (zero_row) SKR_CUDA_CHECK(cudaFreeAsync(zero_row, stream));
    SKR_CUDA_CHECK(cudaFreeAsync(p.spill, stream));
    if (long_list) SKR_CUDA_CHECK(cudaFreeAsync(long_list, stream));
    return SKR_OK;
}

template <int K>
int dispatch_count(const CountParams& p, int vec_is_f64, int out_is_f64, cudaStream_t stream) {
    if (out_is_f64) return launch_count<K, false, double>(p, stream);
    if (vec_is_f64) return launch_count<K, true, float>(p, stream);
    return launch_count<K, false, float>(p, stream);
}

// ---------------------------------------------------------------------------------------------
// element-wise kernels of the get_counts() tail
// ---------------------------------------------------------------------------------------------
enum { OP_LOG2 = 0, OP_POST = 1, OP_SUB = 2, OP_DIV = 3, OP_SCAN = 4, OP_NORM = 5, OP_NORMPOST = 6 };

template <int OP, bool kVecF64>
__device__ __forceinline__ float ew_apply(float v, const void* vec, const void* vec2, long long col, float shift) {
    if constexpr (OP == OP_NORM || OP == OP_NORMPOST) {  // fl(fl(v - mean) / std), either vector optional
        if (vec) v = kVecF64 ? __double2float_rn(__dsub_rn((double)v, __ldg((const double*)vec + col)))
                             : __fsub_rn(v, __ldg((const float*)vec + col));
        if (vec2) v = kVecF64 ? __double2float_rn(__ddiv_rn((double)v, __ldg((const double*)vec2 + col)))
                              : __fdiv_rn(v, __ldg((const float*)vec2 + col));
        if constexpr (OP == OP_NORMPOST) v = log2_post(__fadd_rn(__fadd_rn(v, shift), 1.0f));  // then the Log2.post tail
        return v;
    }
    if constexpr (OP == OP_LOG2) return log2f(__fadd_rn(v, 1.0f));
    if constexpr (OP == OP_POST) return log2_post(__fadd_rn(__fadd_rn(v, shift), 1.0f));
    if constexpr (OP == OP_SUB) {
        if constexpr (kVecF64) return __double2float_rn(__dsub_rn((double)v, __ldg((const double*)vec + col)));
        else return __fsub_rn(v, __ldg((const float*)vec + col));
    }
    if constexpr (OP == OP_DIV) {
        if constexpr (kVecF64) return __double2float_rn(__ddiv_rn((double)v, __ldg((const double*)vec + col)));
        else return __fdiv_rn(v, __ldg((const float*)vec + col));
    }
    return v;
}

// Element-wise kernels walk the matrix column-quad by column-quad: a thread owns four consecutive
// columns (blockIdx.x * 256 + threadIdx.x) and strides over rows (blockIdx.y, gridDim.y), so its mean /
// std values are loaded once into registers and no index division is needed; a warp reads 512
// contiguous bytes of a row per step.  The vector path needs cols % 4 == 0 and 16-byte aligned rows.
template <int OP, bool kVecF64>
__device__ __forceinline__ float ew_apply_reg(float v, double m64, double s64, float m32, float s32, bool has_mean,
                                              bool has_std, float shift) {
    if constexpr (OP == OP_LOG2) return log2f(__fadd_rn(v, 1.0f));
    if constexpr (OP == OP_POST) return log2_post(__fadd_rn(__fadd_rn(v, shift), 1.0f));
    if constexpr (OP == OP_SUB) return kVecF64 ? __double2float_rn(__dsub_rn((double)v, m64)) : __fsub_rn(v, m32);
    if constexpr (OP == OP_DIV) return kVecF64 ? __double2float_rn(__ddiv_rn((double)v, m64)) : __fdiv_rn(v, m32);
    if constexpr (OP == OP_NORM || OP == OP_NORMPOST) {
        if (has_mean) v = kVecF64 ? __double2float_rn(__dsub_rn((double)v, m64)) : __fsub_rn(v, m32);
        if (has_std) v = kVecF64 ? __double2float_rn(__ddiv_rn((double)v, s64)) : __fdiv_rn(v, s32);
        if constexpr (OP == OP_NORMPOST) v = log2_post(__fadd_rn(__fadd_rn(v, shift), 1.0f));
        return v;
    }
    return v;
}

template <int OP, bool kVecF64, bool kVec4>
__global__ void __launch_bounds__(256) ew_kernel(float* a, long long m, long long cols, long long ld, const void* vec,
                                                 const void* vec2, const SkrMinCell* min_in, SkrMinCell* min_out,
                                                 const uint32_t* skip, uint32_t skip_value) {
    __shared__ float s_wmin[8];
    __shared__ int s_wnan[8];
    if (skip && *skip == skip_value) return;
    float shift = 0.0f;
    if constexpr (OP == OP_POST || OP == OP_NORMPOST) {
        // np.abs(np.min(counts)): NaN anywhere makes the shift NaN, hence the whole matrix (kmer_counts.py:208)
        shift = min_in->nan_seen ? __int_as_float(0x7FC00000) : fabsf(skr::ordered_decode(min_in->min_ordered));
    }
    float tmin = INFINITY;
    int tnan = 0;
    constexpr int W = kVec4 ? 4 : 1;
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x;  // column group
    const long long col0 = q * W;
    if (col0 < cols) {
        // this thread's slice of the vectors (vec: mean or the single operand; vec2: std)
        float m32[W], s32[W];
        double m64[W], s64[W];
#pragma unroll
        for (int e = 0; e < W; ++e) {
            m32[e] = s32[e] = 0.0f;
            m64[e] = s64[e] = 0.0;
            if (vec) { if (kVecF64) m64[e] = ((const double*)vec)[col0 + e]; else m32[e] = ((const float*)vec)[col0 + e]; }
            if (vec2) { if (kVecF64) s64[e] = ((const double*)vec2)[col0 + e]; else s32[e] = ((const float*)vec2)[col0 + e]; }
        }
        const bool has_mean = vec != nullptr, has_std = vec2 != nullptr;
        for (long long r = blockIdx.y; r < m; r += gridDim.y) {
            if constexpr (kVec4) {
                float4* ptr = reinterpret_cast<float4*>(a + r * ld) + q;
                float4 v = *ptr;
                v.x = ew_apply_reg<OP, kVecF64>(v.x, m64[0], s64[0], m32[0], s32[0], has_mean, has_std, shift);
                v.y = ew_apply_reg<OP, kVecF64>(v.y, m64[1], s64[1], m32[1], s32[1], has_mean, has_std, shift);
                v.z = ew_apply_reg<OP, kVecF64>(v.z, m64[2], s64[2], m32[2], s32[2], has_mean, has_std, shift);
                v.w = ew_apply_reg<OP, kVecF64>(v.w, m64[3], s64[3], m32[3], s32[3], has_mean, has_std, shift);
                if (min_out) {
                    skr::min_update(v.x, tmin, tnan); skr::min_update(v.y, tmin, tnan);
                    skr::min_update(v.z, tmin, tnan); skr::min_update(v.w, tmin, tnan);
                }
                if constexpr (OP != OP_SCAN) *ptr = v;
            } else {
                float v = ew_apply_reg<OP, kVecF64>(a[r * ld + col0], m64[0], s64[0], m32[0], s32[0], has_mean, has_std, shift);
                if (min_out) skr::min_update(v, tmin, tnan);
                if constexpr (OP != OP_SCAN) a[r * ld + col0] = v;
            }
        }
    }
    if (min_out) skr::min_commit<256>(tmin, tnan, s_wmin, s_wnan, min_out);
}

template <int OP>
int launch_ew(float* a, long long m, long long cols, long long ld, const void* vec, int vec_is_f64,
              const SkrMinCell* min_in, SkrMinCell* min_out, cudaStream_t stream, const void* vec2 = nullptr,
              const uint32_t* skip = nullptr, uint32_t skip_value = 1) {
    if (m <= 0 || cols <= 0) return SKR_OK;
    if (!a) return skr::fail(SKR_ERR_ARG, "null matrix");
    if (ld < cols) return skr::fail(SKR_ERR_ARG, "ld < cols");
    const bool vec4 = (cols % 4 == 0) && (ld % 4 == 0) && (((uintptr_t)a & 15) == 0);
    int dev = 0, sms = 0;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    SKR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long groups = vec4 ? cols / 4 : cols;
    const long long gx = (groups + 255) / 256;
    long long gy = ((long long)sms * 8 + gx - 1) / gx;  // ~8 resident CTAs of 256 threads per SM
    if (gy > m) gy = m;
    if (gy > 65535) gy = 65535;
    if (gx > 0x7FFFFFFFll) return skr::fail(SKR_ERR_ARG, "matrix too wide");
    dim3 grid((unsigned)gx, (unsigned)gy);
    auto go = [&](auto kern) {
        kern<<<grid, 256, 0, stream>>>(a, m, cols, ld, vec, vec2, min_in, min_out, skip, skip_value);
    };
    if (vec_is_f64) {
        if (vec4) go(ew_kernel<OP, true, true>); else go(ew_kernel<OP, true, false>);
    } else {
        if (vec4) go(ew_kernel<OP, false, true>); else go(ew_kernel<OP, false, false>);
    }
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

// Columns whose minimum is still unknown (+inf: no record had a zero count there) are reduced over all
// rows here: one warp per strip of 32 columns, coalesced 128-byte reads; strips without such a column exit
// at once, which is the normal case for k >= 5 (every k-mer is absent from some transcript).
__global__ void __launch_bounds__(256) colmin_scan_kernel(const float* __restrict__ a, long long m, long long cols,
                                                          long long ld, long long rows_per_slab, uint32_t* colmin) {
    const long long strip = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const long long col = strip * 32 + lane;
    const bool need = col < cols && colmin[col] != 0u;  // +inf, or a minimum other slabs are still lowering
    if (!__any_sync(0xFFFFFFFFu, need)) return;
    const long long r0 = (long long)blockIdx.y * rows_per_slab, r1 = min(m, r0 + rows_per_slab);
    uint32_t best = 0x7F800000u;
    if (col < cols) {
        long long r = r0;
        for (; r + 8 <= r1; r += 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __float_as_uint(a[(r + u) * ld + col]);
#pragma unroll
            for (int u = 0; u < 8; ++u) best = min(best, v[u]);
        }
        for (; r < r1; ++r) best = min(best, __float_as_uint(a[r * ld + col]));
    }
    if (need && best != 0x7F800000u) atomicMin(&colmin[col], best);
}

// min over columns of fl(fl(colmin_j - mean_j) / std_j): the matrix-wide minimum of the normalised values
// when every std_j is positive and finite (rounded subtraction and division are monotone)
template <bool kVecF64>
__global__ void __launch_bounds__(256) colmin_finish_kernel(const uint32_t* colmin, long long cols, const void* mean,
                                                            const void* std_, SkrMinCell* cell) {
    __shared__ float s_wmin[8];
    __shared__ int s_wnan[8];
    float tmin = INFINITY;
    int tnan = 0;
    for (long long j = threadIdx.x; j < cols; j += 256) {
        float v = __uint_as_float(colmin[j]);
        v = ew_apply<OP_NORM, kVecF64>(v, mean, std_, j, 0.0f);
        skr::min_update(v, tmin, tnan);
    }
    skr::min_commit<256>(tmin, tnan, s_wmin, s_wnan, cell);
}

__global__ void colmin_reset_kernel(uint32_t* colmin, long long cols) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cols) colmin[j] = 0x7F800000u;  // +inf
}

// flag |= 1 if any element is not finite, |= 2 if any element is <= 0
template <typename T>
__global__ void vec_check_kernel(const T* v, long long n, int* flag) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double x = (double)v[j];
    int f = 0;
    if (!(x - x == 0.0)) f |= 1;
    if (!(x > 0.0)) f |= 2;
    if (f) atomicOr(flag, f);
}

__global__ void reciprocal_kernel(const float* v, long long n, float* out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = __frcp_rn(v[j]);
}

// operand pairs from a counter-based hash: a = +-m1 * 2^ea (ea in [-60, 40], 1/64 of them exactly 0), b = m2 * 2^eb
// (eb in [-40, 40]); counts results of div_by_rcp that differ from __fdiv_rn
__global__ void selftest_division_kernel(unsigned long long n, unsigned long long seed, unsigned long long* bad) {
    unsigned long long local = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long h = (i + seed) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; h *= 0x94D049BB133111EBull; h ^= h >> 29;
        const uint32_t ma = (uint32_t)h & 0x7FFFFFu, mb = (uint32_t)(h >> 23) & 0x7FFFFFu;
        const int ea = (int)((h >> 46) % 101) - 60, eb = (int)((h >> 53) % 81) - 40;
        float a = __uint_as_float(((uint32_t)(ea + 127) << 23) | ma);
        if ((h >> 60) & 1) a = -a;
        if (((h >> 40) & 63) == 0) a = 0.0f;
        const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | mb);
        const float want = __fdiv_rn(a, b);
        const float got = div_by_rcp(a, b, __frcp_rn(b));
        if (__float_as_uint(want) != __float_as_uint(got)) ++local;
    }
    if (local) atomicAdd(bad, local);
}

__global__ void min_reset_kernel(SkrMinCell* cell) {
    cell->min_ordered = skr::ordered_encode(INFINITY);
    cell->nan_seen = 0;
}

}  // namespace

extern "C" int skr_min_reset(SkrMinCell* d_cell, void* stream) {
    if (!d_cell) return skr::fail(SKR_ERR_ARG, "skr_min_reset: null cell");
    min_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_cell);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_count_ex(const SkrCountArgs* a, void* stream) {
    if (!a) return skr::fail(SKR_ERR_ARG, "skr_count_ex: null arguments");
    const int64_t m = a->m;
    const int k = a->k;
    if (m == 0) return SKR_OK;
    if (a->d_rstd && (!a->d_std || a->vec_is_f64 || ((uintptr_t)a->d_rstd & 15)))
        return skr::fail(SKR_ERR_ARG, "skr_count: the reciprocal vector goes with an aligned fp32 std vector");
    if (!a->d_codes || !a->d_mask || !a->d_block_offsets || !a->d_lengths || m < 0)
        return skr::fail(SKR_ERR_ARG, "skr_count: null or negative argument");
    if (!a->d_out && !a->d_colmin) return skr::fail(SKR_ERR_ARG, "skr_count: no output matrix (only a column-minimum pass may omit it)");
    if (k < 1 || k > 8) return skr::fail(SKR_ERR_ARG, "skr_count: k=%d not supported (1 <= k <= 8)", k);
    const int64_t bins = (int64_t)1 << (2 * k);
    if (a->d_out && a->ld_out < bins) return skr::fail(SKR_ERR_ARG, "skr_count: ld_out < 4^k");
    if (a->out_is_f64 && (a->log2_pre || a->d_mean || a->d_std || a->d_min || a->d_post || a->d_colmin || a->d_colsum || a->d_spec))
        return skr::fail(SKR_ERR_ARG, "skr_count: float64 output carries raw counts only");
    const size_t esz = a->out_is_f64 ? 8 : 4;
    if (a->d_out && (((uintptr_t)a->d_out & 15) || ((size_t)a->ld_out * esz) % 16))
        return skr::fail(SKR_ERR_ARG, "skr_count: output rows must be 16-byte aligned");
    if ((a->d_mean && ((uintptr_t)a->d_mean & 15)) || (a->d_std && ((uintptr_t)a->d_std & 15)))
        return skr::fail(SKR_ERR_ARG, "skr_count: mean/std vectors must be 16-byte aligned");
    if (a->d_colmin && (a->d_mean || a->d_std || a->d_post || ((uintptr_t)a->d_colmin & 15)))
        return skr::fail(SKR_ERR_ARG, "skr_count: column minima are those of the un-normalised values (16-byte aligned array)");
    if ((a->d_colsum == nullptr) != (a->d_colsq == nullptr) || (a->d_colsum && (a->d_mean || a->d_std || a->d_post || !a->d_out)))
        return skr::fail(SKR_ERR_ARG, "skr_count: column sums come in pairs and go with plain counts");
    if (a->d_colsum && (k < 4 || k > 6)) return skr::fail(SKR_ERR_ARG, "skr_count: in-kernel column sums are implemented for k = 4, 5, 6");
    if (a->d_spec && !a->d_post) return skr::fail(SKR_ERR_ARG, "skr_count: a speculation cell goes with a Log2.post shift");
    if ((a->d_post_a == nullptr) != (a->d_post_b == nullptr) || (a->d_post_a && !a->d_post))
        return skr::fail(SKR_ERR_ARG, "skr_count: the folded Log2.post tail needs both arrays and the shift they were built from");
    CountParams p{};
    p.codes = a->d_codes;
    p.mask = a->d_mask;
    p.blk_off = a->d_block_offsets;
    p.len = a->d_lengths;
    p.m = m;
    p.log2_pre = a->log2_pre;
    p.mean = a->d_mean;
    p.std_ = a->d_std;
    p.out = a->d_out;
    p.ld_out = a->ld_out;
    p.min_cell = a->d_min;
    p.post_cell = a->d_post;
    p.rstd = a->d_rstd;
    p.colmin = a->d_colmin;
    p.no_store = a->d_out == nullptr;
    p.spec = a->d_spec;
    p.post_a = a->d_post_a;
    p.post_b = a->d_post_b;
    p.skip_flag = a->d_skip;
    p.skip_value = a->skip_value ? a->skip_value : 1u;
    p.spec_epoch = a->spec_epoch ? a->spec_epoch : 1u;
    p.min_reset = a->d_min_reset;
    p.max_length = a->max_length;
    p.colsum = a->d_colsum;
    p.colsq = a->d_colsq;
    cudaStream_t s = (cudaStream_t)stream;
    switch (k) {
        case 1: return dispatch_count<1>(p, a->vec_is_f64, a->out_is_f64, s);
        case 2: return dispatch_count<2>(p, a->vec_is_f64, a->out_is_f64, s);
        case 3: return dispatch_count<3>(p, a->vec_is_f64, a->out_is_f64, s);
        case 4: return dispatch_count<4>(p, a->vec_is_f64, a->out_is_f64, s);
        case 5: return dispatch_count<5>(p, a->vec_is_f64, a->out_is_f64, s);
        case 6: return dispatch_count<6>(p, a->vec_is_f64, a->out_is_f64, s);
        case 7: return dispatch_count<7>(p, a->vec_is_f64, a->out_is_f64, s);
        default: return dispatch_count<8>(p, a->vec_is_f64, a->out_is_f64, s);
    }
}

extern "C" int skr_count(const uint32_t* d_codes, const uint32_t* d_mask, const uint64_t* d_block_offsets,
                         const uint32_t* d_lengths, int64_t m, int k, int log2_pre, const void* d_mean,
                         const void* d_std, int vec_is_f64, void* d_out, int out_is_f64, int64_t ld_out,
                         SkrMinCell* d_min, const SkrMinCell* d_post, const float* d_rstd, void* stream) {
    if (!d_out && m != 0) return skr::fail(SKR_ERR_ARG, "skr_count: null or negative argument");
    SkrCountArgs a{};
    a.d_codes = d_codes;
    a.d_mask = d_mask;
    a.d_block_offsets = d_block_offsets;
    a.d_lengths = d_lengths;
    a.m = m;
    a.k = k;
    a.log2_pre = log2_pre;
    a.d_mean = d_mean;
    a.d_std = d_std;
    a.d_rstd = d_rstd;
    a.vec_is_f64 = vec_is_f64;
    a.out_is_f64 = out_is_f64;
    a.d_out = d_out;
    a.ld_out = ld_out;
    a.d_min = d_min;
    a.d_post = d_post;
    return skr_count_ex(&a, stream);
}

// Speculative Log2.post shift (see SkrPostSpec): min over the columns of the z-score of a ZERO count,
// fl(fl(0 - mean_j) / std_j), and the first column that attains it.
template <bool kVecF64>
__global__ void __launch_bounds__(256) post_spec_kernel(const void* mean, const void* std_, long long cols, SkrPostSpec* spec,
                                                        float* post_a, float* post_b) {
    __shared__ float s_v[256];
    __shared__ long long s_j[256];
    float best = INFINITY;
    long long bj = -1;
    for (long long j = threadIdx.x; j < cols; j += 256) {
        const float v = ew_apply<OP_NORM, kVecF64>(0.0f, mean, std_, j, 0.0f);
        if (v < best) { best = v; bj = j; }  // strict: the first column wins inside a thread (ascending j)
    }
    s_v[threadIdx.x] = best;
    s_j[threadIdx.x] = bj;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 256; ++t)
            if (s_j[t] >= 0 && (bj < 0 || s_v[t] < best || (s_v[t] == best && s_j[t] < bj))) { best = s_v[t]; bj = s_j[t]; }
        spec->shift.min_ordered = skr::ordered_encode(bj >= 0 ? best : 0.0f);
        spec->shift.nan_seen = 0;
        spec->zero_col = (int32_t)bj;  // -1: nothing to speculate on (a NaN / +inf in every column): callers fall back
        spec->zero_seen = 0;
        s_v[0] = bj >= 0 ? fabsf(best) : 0.0f;
    }
    if (!post_a || !post_b || !mean || !std_) return;
    // the tail ((x - mean) / std + shift) + 1 as x * a + b.  b_j is the reference's own value of a ZERO count in
    // column j -- fl(fl(fl(fl(0 - mean_j) / std_j) + shift) + 1), the same fp32 operations -- so every empty bin
    // (almost half of a 6-mer matrix) comes out with the reference's bits and the matrix minimum is exactly
    // log2(1) = 0; a_j = RN(1 / std_j).  Every term of x * a + b is >= 0 (b_j >= 1), nothing cancels: a counted
    // bin differs from the step-by-step tail by a few ulp before the log2.
    __syncthreads();
    const float shift = s_v[0];
    for (long long j = threadIdx.x; j < cols; j += 256) {
        const double sd = kVecF64 ? reinterpret_cast<const double*>(std_)[j] : (double)reinterpret_cast<const float*>(std_)[j];
        const float z0 = ew_apply<OP_NORM, kVecF64>(0.0f, mean, std_, j, 0.0f);
        post_a[j] = (float)(1.0 / sd);
        post_b[j] = __fadd_rn(__fadd_rn(z0, shift), 1.0f);
    }
}

extern "C" int skr_post_spec_affine(const void* d_mean, const void* d_std, int vec_is_f64, int64_t cols, SkrPostSpec* d_spec,
                                    float* d_post_a, float* d_post_b, void* stream) {
    if (!d_spec || cols <= 0 || (!d_mean && !d_std)) return skr::fail(SKR_ERR_ARG, "skr_post_spec: bad argument");
    if ((d_post_a == nullptr) != (d_post_b == nullptr) || (d_post_a && (((uintptr_t)d_post_a | (uintptr_t)d_post_b) & 15)))
        return skr::fail(SKR_ERR_ARG, "skr_post_spec_affine: a and b come together, 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (vec_is_f64) post_spec_kernel<true><<<1, 256, 0, s>>>(d_mean, d_std, cols, d_spec, d_post_a, d_post_b);
    else post_spec_kernel<false><<<1, 256, 0, s>>>(d_mean, d_std, cols, d_spec, d_post_a, d_post_b);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_post_spec(const void* d_mean, const void* d_std, int vec_is_f64, int64_t cols, SkrPostSpec* d_spec,
                             void* stream) {
    return skr_post_spec_affine(d_mean, d_std, vec_is_f64, cols, d_spec, nullptr, nullptr, stream);
}

// mean / std from the column sums the count kernel accumulated (accurate norm_vectors): binary64 throughout,
// mean = S1 / rows, var = S2 / rows - mean^2 (clamped at 0), one rounding to fp32 each
__global__ void colstat_finish_kernel(const double* colsum, const double* colsq, long long cols, long long rows,
                                      float* mean, float* std_, int* flags) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const double mu = colsum[j] / (double)rows;
    double var = colsq[j] / (double)rows - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mf = (float)mu, sf = (float)sqrt(var);
    if (mean) mean[j] = mf;
    if (std_) std_[j] = sf;
    if (flags) {
        int fm = 0, fs = 0;
        if (!(mf - mf == 0.0f)) fm |= 1;
        if (!(sf - sf == 0.0f)) fs |= 1;
        if (!(sf > 0.0f)) fs |= 2;
        if (fm) atomicOr(&flags[0], fm);
        if (fs) atomicOr(&flags[1], fs);
    }
}

extern "C" int skr_colstat_finish(const double* d_colsum, const double* d_colsq, int64_t cols, int64_t total_rows,
                                  float* d_mean, float* d_std, int* d_flags, void* stream) {
    if (cols <= 0) return SKR_OK;
    if (!d_colsum || !d_colsq || total_rows <= 0) return skr::fail(SKR_ERR_ARG, "skr_colstat_finish: bad argument");
    colstat_finish_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_colsum, d_colsq, cols, total_rows,
                                                                                             d_mean, d_std, d_flags);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_colmin_reset(uint32_t* d_colmin, int64_t cols, void* stream) {
    if (cols <= 0) return SKR_OK;
    if (!d_colmin) return skr::fail(SKR_ERR_ARG, "skr_colmin_reset: null");
    colmin_reset_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_colmin, cols);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_count_colmin(const uint32_t* d_codes, const uint32_t* d_mask, const uint64_t* d_block_offsets,
                                const uint32_t* d_lengths, int64_t m, int k, int log2_pre, float* d_out, int64_t ld_out,
                                uint32_t* d_colmin, void* stream) {
    if (m == 0) return SKR_OK;
    if (!d_colmin) return skr::fail(SKR_ERR_ARG, "skr_count_colmin: null or negative argument");
    SkrCountArgs a{};
    a.d_codes = d_codes;
    a.d_mask = d_mask;
    a.d_block_offsets = d_block_offsets;
    a.d_lengths = d_lengths;
    a.m = m;
    a.k = k;
    a.log2_pre = log2_pre;
    a.d_out = d_out;
    a.ld_out = ld_out;
    a.d_colmin = d_colmin;
    return skr_count_ex(&a, stream);
}

extern "C" int skr_colmin_scan(const float* d_a, int64_t m, int64_t cols, int64_t ld, uint32_t* d_colmin, void* stream) {
    if (m <= 0 || cols <= 0) return SKR_OK;
    if (!d_a || !d_colmin) return skr::fail(SKR_ERR_ARG, "skr_colmin_scan: null argument");
    const long long strips = (cols + 31) / 32;
    const long long rows_per_slab = 1024;
    long long slabs = (m + rows_per_slab - 1) / rows_per_slab;
    if (slabs > 65535) return skr::fail(SKR_ERR_ARG, "skr_colmin_scan: too many rows");
    dim3 grid((unsigned)((strips + 7) / 8), (unsigned)slabs);
    colmin_scan_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_a, m, cols, ld, rows_per_slab, d_colmin);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_colmin_finish(const uint32_t* d_colmin, int64_t cols, const void* d_mean, const void* d_std,
                                 int vec_is_f64, SkrMinCell* d_min, void* stream) {
    if (!d_colmin || !d_min || cols <= 0) return skr::fail(SKR_ERR_ARG, "skr_colmin_finish: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    min_reset_kernel<<<1, 1, 0, s>>>(d_min);
    SKR_LAUNCH_CHECK();
    if (vec_is_f64) colmin_finish_kernel<true><<<1, 256, 0, s>>>(d_colmin, cols, d_mean, d_std, d_min);
    else colmin_finish_kernel<false><<<1, 256, 0, s>>>(d_colmin, cols, d_mean, d_std, d_min);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_normalize_post_log2(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_mean,
                                       const void* d_std, int vec_is_f64, const SkrMinCell* d_min, void* stream) {
    if (!d_min) return skr::fail(SKR_ERR_ARG, "skr_normalize_post_log2: null min cell");
    return launch_ew<OP_NORMPOST>(d_a, m, cols, ld, d_mean, vec_is_f64, d_min, nullptr, (cudaStream_t)stream, d_std);
}

extern "C" int skr_reciprocal(const float* d_vec, int64_t n, float* d_out, void* stream) {
    if (n <= 0) return SKR_OK;
    if (!d_vec || !d_out) return skr::fail(SKR_ERR_ARG, "skr_reciprocal: null argument");
    reciprocal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_vec, n, d_out);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_selftest_division(uint64_t n, uint64_t seed, uint64_t* d_mismatches, void* stream) {
    if (!d_mismatches) return skr::fail(SKR_ERR_ARG, "skr_selftest_division: null argument");
    selftest_division_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(n, seed, (unsigned long long*)d_mismatches);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_vec_check(const void* d_vec, int vec_is_f64, int64_t n, int* d_flag, void* stream) {
    if (n <= 0) return SKR_OK;
    if (!d_vec || !d_flag) return skr::fail(SKR_ERR_ARG, "skr_vec_check: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (vec_is_f64) vec_check_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const double*)d_vec, n, d_flag);
    else vec_check_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float*)d_vec, n, d_flag);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_log2_norm(float* d_a, int64_t m, int64_t cols, int64_t ld, void* stream) {
    return launch_ew<OP_LOG2>(d_a, m, cols, ld, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int skr_post_log2(float* d_a, int64_t m, int64_t cols, int64_t ld, const SkrMinCell* d_min, void* stream) {
    if (!d_min) return skr::fail(SKR_ERR_ARG, "skr_post_log2: null min cell");
    return launch_ew<OP_POST>(d_a, m, cols, ld, nullptr, 0, d_min, nullptr, (cudaStream_t)stream);
}

extern "C" int skr_post_log2_skip(float* d_a, int64_t m, int64_t cols, int64_t ld, const SkrMinCell* d_min,
                                  const uint32_t* d_skip, uint32_t skip_value, void* stream) {
    if (!d_min) return skr::fail(SKR_ERR_ARG, "skr_post_log2_skip: null min cell");
    return launch_ew<OP_POST>(d_a, m, cols, ld, nullptr, 0, d_min, nullptr, (cudaStream_t)stream, nullptr, d_skip,
                              skip_value ? skip_value : 1u);
}

extern "C" int skr_sub_vec(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec, int vec_is_f64,
                           SkrMinCell* d_min, void* stream) {
    if (!d_vec) return skr::fail(SKR_ERR_ARG, "skr_sub_vec: null vector");
    return launch_ew<OP_SUB>(d_a, m, cols, ld, d_vec, vec_is_f64, nullptr, d_min, (cudaStream_t)stream);
}

extern "C" int skr_div_vec(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec, int vec_is_f64,
                           SkrMinCell* d_min, void* stream) {
    if (!d_vec) return skr::fail(SKR_ERR_ARG, "skr_div_vec: null vector");
    return launch_ew<OP_DIV>(d_a, m, cols, ld, d_vec, vec_is_f64, nullptr, d_min, (cudaStream_t)stream);
}

extern "C" int skr_normalize(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_mean, const void* d_std,
                             int vec_is_f64, SkrMinCell* d_min, void* stream) {
    if (!d_mean && !d_std) return skr::fail(SKR_ERR_ARG, "skr_normalize: no vector given");
    return launch_ew<OP_NORM>(d_a, m, cols, ld, d_mean, vec_is_f64, nullptr, d_min, (cudaStream_t)stream, d_std);
}

extern "C" int skr_min_scan(const float* d_a, int64_t m, int64_t cols, int64_t ld, SkrMinCell* d_min, void* stream) {
    if (!d_min) return skr::fail(SKR_ERR_ARG, "skr_min_scan: null min cell");
    return launch_ew<OP_SCAN>(const_cast<float*>(d_a), m, cols, ld, nullptr, 0, nullptr, d_min, (cudaStream_t)stream);
}

extern "C" double skr_chain_sum_host(double increment, uint32_t count) { return skr::chain_sum(increment, count); }
