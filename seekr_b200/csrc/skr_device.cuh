// Device helpers shared by the kernels: the exact c-fold binary64 sum, the NaN-propagating
// running minimum, and small PTX wrappers (mbarrier, TMA, tcgen05) used by the column and
// Pearson kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "seekr_b200.h"

namespace skr {

// ---------------------------------------------------------------------------------------------
// chain_sum(inc, c): the value of `acc = 0; repeat c times: acc += inc` in IEEE binary64
// (round-to-nearest-even), which is what kmer_counts.py:148 computes for a k-mer seen c times.
//
// A literal loop costs c dependent adds (a 100 kb homopolymer would serialise 10^5 of them), so
// beyond a few terms the sum is advanced one binade at a time.  Inside a binade [2^E, 2^(E+1))
// the accumulator is a multiple of u = 2^(E-52) and every step adds inc rounded to a multiple of
// u: the same amount d each time, except that when inc's remainder is exactly u/2 (a tie) the
// first step of the binade may differ, after which the accumulator is even and the step is
// constant again.  So: take real steps until two consecutive results lie in the same binade (the
// second one is then "settled"), measure d from one more real step, and jump
// j = min(remaining, steps that stay below 2^(E+1)) steps at once with exact integer arithmetic.
// The crossing into the next binade is always done by a real add.  Checked against the literal
// loop in tests/test_chain_sum.py (host) and through the count kernel (device).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int f64_exponent(double x) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(x) >> 20) & 0x7FF;
#else
    union { double d; uint64_t u; } v;
    v.d = x;
    return (int)((v.u >> 52) & 0x7FF);
#endif
}

__host__ __device__ inline double f64_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;  // one IEEE add, not contracted or reassociated
    return r;
#endif
}

__host__ __device__ inline double f64_from_parts(long long mant, int biased_exp) {
    // mant in [2^52, 2^53), value = mant * 2^(biased_exp - 1075)
    unsigned long long bits = ((unsigned long long)biased_exp << 52) | ((unsigned long long)mant & 0xFFFFFFFFFFFFFull);
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)bits);
#else
    union { double d; uint64_t u; } v;
    v.u = bits;
    return v.d;
#endif
}

__host__ __device__ inline long long f64_mantissa(double x) {
#ifdef __CUDA_ARCH__
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
#else
    union { double d; uint64_t u; } v;
    v.d = x;
    unsigned long long bits = v.u;
#endif
    return (long long)((bits & 0xFFFFFFFFFFFFFull) | (1ull << 52));
}

__host__ __device__ inline double chain_sum(double inc, uint32_t c) {
    if (c == 0) return 0.0;
    if (c <= 6) {
        double acc = inc;
        for (uint32_t i = 1; i < c; ++i) acc = f64_add(acc, inc);
        return acc;
    }
    double acc = inc;  // 0 + inc is exact
    uint32_t rem = c - 1;
    while (rem) {
        // a1: one real step
        double a1 = f64_add(acc, inc);
        --rem;
        if (!rem || f64_exponent(a1) != f64_exponent(acc)) { acc = a1; continue; }
        // acc -> a1 stayed inside the binade, so a1 is settled; measure the steady step from it
        double a2 = f64_add(a1, inc);
        --rem;
        if (!rem || f64_exponent(a2) != f64_exponent(a1)) { acc = a2; continue; }
        const int e = f64_exponent(a2);
        const long long A = f64_mantissa(a2);
        const long long D = A - f64_mantissa(a1);  // steady step in units of 2^(e-1075)
        if (D <= 0) { acc = a2; continue; }        // cannot happen for c < 2^32 (inc >= 2^-32 * acc); stay safe
        // largest j with A + j*D <= 2^53 - 1 (still inside the binade)
        long long j = ((1ll << 53) - 1 - A) / D;
        if (j > (long long)rem) j = rem;
        acc = f64_from_parts(A + j * D, e);
        rem -= (uint32_t)j;
    }
    return acc;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// NaN-propagating minimum (np.min semantics, kmer_counts.py:208)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_encode(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_decode(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// tmin carries the NaN itself (min.NaN propagates it), so one instruction per value suffices;
// tnan is only folded in when the running value is committed.
__device__ __forceinline__ void min_update(float v, float& tmin, int& tnan) {
    (void)tnan;
    asm("min.NaN.f32 %0, %0, %1;" : "+f"(tmin) : "f"(v));
}

template <int T>
__device__ __forceinline__ void min_commit(float tmin, int tnan, float* s_wmin, int* s_wnan, SkrMinCell* cell) {
    if (tmin != tmin) { tnan = 1; tmin = INFINITY; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tmin = fminf(tmin, __shfl_xor_sync(0xFFFFFFFFu, tmin, o));
        tnan |= __shfl_xor_sync(0xFFFFFFFFu, tnan, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_wmin[w] = tmin; s_wnan[w] = tnan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < T / 32; ++i) { tmin = fminf(tmin, s_wmin[i]); tnan |= s_wnan[i]; }
        if (tmin < INFINITY) atomicMin(&cell->min_ordered, ordered_encode(tmin));
        if (tnan) atomicOr(&cell->nan_seen, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, one issue slot for two IEEE round-to-nearest
// operations, no flush-to-zero).  A pair lives in an aligned 64-bit register pair.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {  // bar_addr: shared-window address
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 2-D tiled TMA load: coordinates are (inner, outer) element indices of the box origin.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

#endif  // __CUDACC__

}  // namespace skr
