// Consumers of the Pearson matrix that the reference runs as Python loops over every r value
// (SURVEY section 8f, rows 1 and 2):
//   skr_pval_empirical   find_pval.py:157-159   p[i,j] = count(background > r[i,j]) / N
//   skr_pval_dist        find_pval.py:126-128   p[i,j] = 1 - CDF((r[i,j] - loc) / scale)  (closed-form families)
//   skr_triu_extract     find_dist.py:163       sim[np.triu_indices(n, k=1)], row-major
//   skr_pearson_pairs    find_dist.py:160-169   r of sampled (i, j) pairs without forming the n x n matrix
// All of them are HBM-bound element-wise / gather kernels over device-resident matrices.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>

#include "skr_common.h"
#include "skr_device.cuh"

namespace {

// count(background > x) through a fine uniform-bin table.  bin() is a monotone function of the value, evaluated
// by the same instructions for background values and for queries, so the table needs no floating-point edges:
// lut[b] = number of background values whose bin is below b.  Every value in a lower bin is < x and every value
// in a higher bin is > x, hence upper_bound(x) lies inside [lut[b], lut[b+1]) -- 0.8 values on average for 100 000
// background values, a handful at the mode.  Keys are float32 when both sides are float32 (exact), else binary64.
constexpr int kBins = 1 << 17;
constexpr float kBinLo = -1.0078125f, kBinScale = (float)kBins / 2.015625f;  // [-1.0078, 1.0078): Pearson r

__device__ __forceinline__ int bin_of(float x) {
    const float t = __fmul_rn(__fsub_rn(x, kBinLo), kBinScale);
    return t < 0.0f ? 0 : (t >= (float)(kBins - 1) ? kBins - 1 : (int)t);  // NaN is handled by the caller
}
__device__ __forceinline__ int bin_of(double x) {
    const double t = __dmul_rn(__dsub_rn(x, (double)kBinLo), (double)kBinScale);
    return t < 0.0 ? 0 : (t >= (double)(kBins - 1) ? kBins - 1 : (int)t);
}

// KT: the type comparisons are made in
template <typename BG, typename KT>
__global__ void pval_lut_kernel(const BG* __restrict__ sorted, long long N, uint32_t* __restrict__ lut) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > kBins) return;
    long long lo = 0, hi = N;  // first value whose bin is >= b
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (bin_of((KT)sorted[mid]) < b) lo = mid + 1; else hi = mid;
    }
    lut[b] = (uint32_t)lo;
}

// p = count(bg > r) / N with the division in binary64 and one rounding to T (numpy: int64 / int -> float64,
// stored into zeros_like(sim)).  NaN r compares false with everything: p = 0.
template <typename T, typename BG, typename KT>
__global__ void __launch_bounds__(256) pval_empirical_kernel(const T* __restrict__ r, long long m, long long n, long long ld,
                                                             const BG* __restrict__ sorted, long long N,
                                                             const uint32_t* __restrict__ lut, T* __restrict__ p,
                                                             long long ldp) {
    // float32 output with N < 2^24: count and N are exact float32 values and fl32(count / N) equals
    // fl32(fl64(count / N)) -- a quotient of integers below 2^24 is never within 2^-49 (relative) of a float32
    // rounding boundary unless it lies on it, so the intermediate binary64 rounding of the reference cannot change
    // the result -- which keeps the binary64 divider out of the inner loop.
    const bool small_n = sizeof(T) == 4 && N < (1ll << 24);
    const float nf = (float)N;
    // The kernel is bound by the latency of its dependent loads (r -> table -> a few background values), so every
    // thread walks kU independent columns in lock step: each step of the search issues kU independent loads.
    constexpr int kU = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint32_t last = (uint32_t)(N - 1);
    for (long long row = blockIdx.y; row < m; row += gridDim.y) {
        const T* rr = r + row * ld;
        T* pr = p + row * ldp;
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride * kU) {
            KT x[kU];
            uint32_t lo[kU], hi[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) x[u] = j + u * stride < n ? (KT)rr[j + u * stride] : (KT)NAN;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                lo[u] = hi[u] = (uint32_t)N;  // NaN: nothing is greater
                if (x[u] == x[u]) {
                    const int b = bin_of(x[u]);
                    lo[u] = __ldg(lut + b);
                    hi[u] = __ldg(lut + b + 1);
                }
            }
            for (;;) {  // first value > x inside the bin, all kU searches advance together
                bool more = false;
                uint32_t mid[kU];
                KT v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    mid[u] = lo[u] + ((hi[u] - lo[u]) >> 1);
                    v[u] = (KT)__ldg(sorted + min(mid[u], last));
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    if (lo[u] < hi[u]) {
                        if (v[u] <= x[u]) lo[u] = mid[u] + 1; else hi[u] = mid[u];
                        more = more || lo[u] < hi[u];
                    }
                }
                if (!more) break;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (j + u * stride < n) {
                    const uint32_t cnt = (uint32_t)N - lo[u];
                    pr[j + u * stride] = small_n ? (T)__fdiv_rn((float)cnt, nf) : (T)((double)cnt / (double)N);
                }
            }
        }
    }
}

// scipy.stats closed-form families of find_dist's 'common10' list (scipy/stats/_continuous_distns.py _cdf
// bodies; rv_continuous.cdf supplies the support handling: 0 at or below the lower end, 1 at or above the
// upper end, NaN for NaN input or invalid parameters).
__device__ __forceinline__ double ndtr(double a) {
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 1.0) return 0.5 + 0.5 * erf(x);
    const double y = 0.5 * erfc(z);
    return x > 0 ? 1.0 - y : y;
}

// Regularised lower incomplete gamma P(a, x) in binary64 (scipy.special.gammainc behind gamma._cdf, chdtr behind
// chi2._cdf): power series for x < a + 1, modified-Lentz continued fraction of Q = 1 - P otherwise.  The factor
// x^a e^-x / Gamma(a) is formed as exp(c0 - a (mu - log1p(mu))), mu = (x - a) / a, with
// c0 = a ln a - a - lgamma(a) computed once on the host in extended precision: the large terms a ln x, x and
// lgamma(a) never meet in binary64, so the relative error stays near 1e-15 (a |mu| eps) instead of a ln(a) eps.
__device__ double igam_p(double a, double x, double c0) {
    if (x <= 0.0) return 0.0;
    if (x == INFINITY) return 1.0;
    const double mu = (x - a) / a;
    const double pre = exp(c0 - a * (mu - log1p(mu)));
    const int kMaxIter = 1 << 20;
    // Both expansions run as numerator / denominator recurrences WITHOUT a division per term (a binary64 division is
    // ~40 instructions, the recurrences are 4-6): the first version, `del *= x / ap` and Lentz' two divisions per
    // step, took 234 ms per 1e9 values at a = 120.
    if (x < a + 1.0) {
        // sum_n x^n / (a (a+1) ... (a+n)) = S_n / D_n:  N_n = N_{n-1} x,  S_n = S_{n-1} (a+n) + N_n,  D_n = D_{n-1} (a+n)
        double ap = a, N = 1.0, S = 1.0, D = a;
        for (int i = 0; i < kMaxIter; ++i) {
            ap += 1.0;
            N *= x;
            S = fma(S, ap, N);
            D *= ap;
            if (N < S * 1e-17) break;
            if (D > 0x1p600) {  // exact rescaling by a power of two
                N *= 0x1p-600;
                S *= 0x1p-600;
                D *= 0x1p-600;
            }
        }
        return fmin(S / D * pre, 1.0);
    }
    // Q = pre * continued fraction, convergents p_k / q_k by the forward recurrence (the form cephes' igamc uses),
    // compared every fourth step
    double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
    double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
    double ans = pkm1 / qkm1;
    for (int i = 1; i <= kMaxIter; ++i) {
        c += 1.0;
        y += 1.0;
        z += 2.0;
        const double yc = y * c;
        const double pk = fma(pkm1, z, -(pkm2 * yc));
        const double qk = fma(qkm1, z, -(qkm2 * yc));
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;
        if (fabs(pk) > 0x1p500) {
            pkm2 *= 0x1p-500;
            pkm1 *= 0x1p-500;
            qkm2 *= 0x1p-500;
            qkm1 *= 0x1p-500;
        }
        if ((i & 3) == 0 && qk != 0.0) {
            const double r = pk / qk;
            const bool done = fabs(ans - r) <= fabs(r) * 1.2e-16;
            ans = r;
            if (done) break;
        }
    }
    return 1.0 - pre * ans;
}

__device__ __forceinline__ double dist_cdf(int kind, double x, double shape, double aux) {
    switch (kind) {
        case SKR_DIST_GAMMA: return igam_p(shape, x, aux);
        case SKR_DIST_CHI2: return igam_p(0.5 * shape, 0.5 * x, aux);
        case SKR_DIST_NORM: return ndtr(x);
        case SKR_DIST_LOGNORM: return x <= 0.0 ? 0.0 : ndtr(log(x) / shape);
        case SKR_DIST_CAUCHY: return atan2(1.0, -x) / 3.14159265358979323846;
        case SKR_DIST_EXPON: return x <= 0.0 ? 0.0 : -expm1(-x);
        case SKR_DIST_RAYLEIGH: return x <= 0.0 ? 0.0 : -expm1(-0.5 * (x * x));
        case SKR_DIST_UNIFORM: return x <= 0.0 ? 0.0 : (x >= 1.0 ? 1.0 : x);
        case SKR_DIST_PARETO: return x <= 1.0 ? 0.0 : 1.0 - pow(x, -shape);
        case SKR_DIST_EXPONPOW: return x <= 0.0 ? 0.0 : -expm1(-expm1(pow(x, shape)));
        default: return NAN;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pval_dist_kernel(const T* __restrict__ r, long long m, long long n, long long ld,
                                                        int kind, double shape, double aux, double loc, double scale,
                                                        int valid, T* __restrict__ p, long long ldp) {
    for (long long row = blockIdx.y; row < m; row += gridDim.y) {
        const T* rr = r + row * ld;
        T* pr = p + row * ldp;
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
            const double x = ((double)rr[j] - loc) / scale;
            double c;
            if (!valid || x != x) c = NAN;
            else if (x == INFINITY) c = 1.0;
            else c = dist_cdf(kind, x, shape, aux);
            pr[j] = (T)(1.0 - c);
        }
    }
}

// out[off(i) + j - i - 1] = c[i][j] for j > i, off(i) = i*(n-1) - i*(i-1)/2  (np.triu_indices(n, k=1) order)
template <typename T>
__global__ void __launch_bounds__(256) triu_extract_kernel(const T* __restrict__ c, long long n, long long ld,
                                                           T* __restrict__ out) {
    for (long long i = blockIdx.x; i < n - 1; i += gridDim.x) {
        const long long off = i * (n - 1) - i * (i - 1) / 2;
        const T* row = c + i * ld + i + 1;
        T* dst = out + off;
        for (long long j = threadIdx.x; j < n - 1 - i; j += blockDim.x) dst[j] = row[j];
    }
}

// One warp per pair (row scales as skr_pearson_prepare stores them, 2^-e): r = alpha * sa[i] * sb[j] * sum_k (hi+lo)_a[i][k] * (hi+lo)_b[j][k].  hi + lo is a 22-bit
// value (exact in fp32), the product of two is exact in binary64, and the sum runs in binary64.
__global__ void __launch_bounds__(256) pearson_pairs_kernel(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo,
                                                            const float* __restrict__ a_scale,
                                                            const __half* __restrict__ b_hi, const __half* __restrict__ b_lo,
                                                            const float* __restrict__ b_scale, long long kp,
                                                            const long long* __restrict__ pi, const long long* __restrict__ pj,
                                                            long long npairs, double alpha, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long q = warp0; q < npairs; q += nwarps) {
        const long long i = pi[q], j = pj[q];
        const uint4* ah = reinterpret_cast<const uint4*>(a_hi + i * kp);
        const uint4* al = reinterpret_cast<const uint4*>(a_lo + i * kp);
        const uint4* bh = reinterpret_cast<const uint4*>(b_hi + j * kp);
        const uint4* bl = reinterpret_cast<const uint4*>(b_lo + j * kp);
        double acc = 0.0;
        for (long long v = lane; v < kp / 8; v += 32) {  // 8 halves per 16-byte load; padding columns are zero
            const uint4 xh = __ldg(ah + v), xl = __ldg(al + v), yh = __ldg(bh + v), yl = __ldg(bl + v);
            const __half2* xh2 = reinterpret_cast<const __half2*>(&xh);
            const __half2* xl2 = reinterpret_cast<const __half2*>(&xl);
            const __half2* yh2 = reinterpret_cast<const __half2*>(&yh);
            const __half2* yl2 = reinterpret_cast<const __half2*>(&yl);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a0 = __half22float2(xh2[e]), a1 = __half22float2(xl2[e]);
                const float2 b0 = __half22float2(yh2[e]), b1 = __half22float2(yl2[e]);
                acc = fma((double)(a0.x + a1.x), (double)(b0.x + b1.x), acc);
                acc = fma((double)(a0.y + a1.y), (double)(b0.y + b1.y), acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (lane == 0) out[q] = (float)(alpha * acc * (double)a_scale[i] * (double)b_scale[j]);  // scales are 2^-e: exact
    }
}

int grid_for(long long total, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sms * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// rows x column tiles: x covers a row in 256-thread blocks (at most 64 of them), y strides over rows
dim3 grid_2d(long long m, long long n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long gx = (n + 255) / 256;
    if (gx > 64) gx = 64;
    long long gy = ((long long)sms * 8 + gx - 1) / gx;
    if (gy > m) gy = m;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace

extern "C" int skr_pval_empirical(const void* d_r, int r_is_f64, int64_t m, int64_t n, int64_t ld, const void* d_sorted_bg,
                                  int bg_is_f64, int64_t N, void* d_p, int64_t ldp, void* stream) {
    if (m <= 0 || n <= 0) return SKR_OK;
    if (!d_r || !d_p || !d_sorted_bg || N <= 0) return skr::fail(SKR_ERR_ARG, "skr_pval_empirical: bad argument");
    if (ld < n || ldp < n) return skr::fail(SKR_ERR_ARG, "skr_pval_empirical: leading dimension smaller than n");
    if (N >= 0xFFFFFFFFll) return skr::fail(SKR_ERR_ARG, "skr_pval_empirical: more than 2^32 - 1 background values");
    // the table is stream-ordered scratch: concurrent calls on other streams get their own
    uint32_t* lut = nullptr;
    SKR_CUDA_CHECK(cudaMallocAsync(&lut, (size_t)(kBins + 1) * sizeof(uint32_t), (cudaStream_t)stream));
    cudaStream_t s = (cudaStream_t)stream;
    const bool f32 = !r_is_f64 && !bg_is_f64;
    const int lut_blocks = (kBins + 1 + 255) / 256;
    if (f32) pval_lut_kernel<float, float><<<lut_blocks, 256, 0, s>>>((const float*)d_sorted_bg, N, lut);
    else if (bg_is_f64) pval_lut_kernel<double, double><<<lut_blocks, 256, 0, s>>>((const double*)d_sorted_bg, N, lut);
    else pval_lut_kernel<float, double><<<lut_blocks, 256, 0, s>>>((const float*)d_sorted_bg, N, lut);
    SKR_LAUNCH_CHECK();
    const dim3 grid = grid_2d(m, n);
#define SKR_PVAL_LAUNCH(T, BG, KT)                                                                                   \
    pval_empirical_kernel<T, BG, KT><<<grid, 256, 0, s>>>((const T*)d_r, m, n, ld, (const BG*)d_sorted_bg, N, lut, (T*)d_p, ldp)
    if (r_is_f64) { if (bg_is_f64) SKR_PVAL_LAUNCH(double, double, double); else SKR_PVAL_LAUNCH(double, float, double); }
    else { if (bg_is_f64) SKR_PVAL_LAUNCH(float, double, double); else SKR_PVAL_LAUNCH(float, float, float); }
#undef SKR_PVAL_LAUNCH
    SKR_LAUNCH_CHECK();
    SKR_CUDA_CHECK(cudaFreeAsync(lut, s));
    return SKR_OK;
}

extern "C" int skr_pval_dist(const void* d_r, int r_is_f64, int64_t m, int64_t n, int64_t ld, int kind, double shape,
                             double loc, double scale, void* d_p, int64_t ldp, void* stream) {
    if (m <= 0 || n <= 0) return SKR_OK;
    if (!d_r || !d_p) return skr::fail(SKR_ERR_ARG, "skr_pval_dist: null argument");
    if (ld < n || ldp < n) return skr::fail(SKR_ERR_ARG, "skr_pval_dist: leading dimension smaller than n");
    if (kind < SKR_DIST_NORM || kind > SKR_DIST_CHI2) return skr::fail(SKR_ERR_ARG, "skr_pval_dist: unknown family %d", kind);
    const bool needs_shape = kind == SKR_DIST_LOGNORM || kind == SKR_DIST_PARETO || kind == SKR_DIST_EXPONPOW ||
                             kind == SKR_DIST_GAMMA || kind == SKR_DIST_CHI2;
    // rv_continuous._argcheck: shape parameters must be > 0; scale must be > 0; otherwise every value is NaN
    const int valid = (scale > 0.0) && (!needs_shape || shape > 0.0) && std::isfinite(loc);
    double aux = 0.0;
    if (valid && (kind == SKR_DIST_GAMMA || kind == SKR_DIST_CHI2)) {
        // c0 = a ln a - a - lgamma(a) in extended precision (see igam_p)
        const long double a = kind == SKR_DIST_CHI2 ? 0.5L * (long double)shape : (long double)shape;
        aux = (double)(a * logl(a) - a - lgammal(a));
    }
    const dim3 grid = grid_2d(m, n);
    cudaStream_t s = (cudaStream_t)stream;
    if (r_is_f64) pval_dist_kernel<double><<<grid, 256, 0, s>>>((const double*)d_r, m, n, ld, kind, shape, aux, loc, scale, valid, (double*)d_p, ldp);
    else pval_dist_kernel<float><<<grid, 256, 0, s>>>((const float*)d_r, m, n, ld, kind, shape, aux, loc, scale, valid, (float*)d_p, ldp);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int64_t skr_triu_count(int64_t n) { return n > 1 ? n * (n - 1) / 2 : 0; }

extern "C" int skr_triu_extract(const void* d_c, int c_is_f64, int64_t n, int64_t ld, void* d_out, void* stream) {
    if (n <= 1) return SKR_OK;
    if (!d_c || !d_out) return skr::fail(SKR_ERR_ARG, "skr_triu_extract: null argument");
    if (ld < n) return skr::fail(SKR_ERR_ARG, "skr_triu_extract: leading dimension smaller than n");
    int dev = 0, sms = 148;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    SKR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long grid = (long long)sms * 8;
    if (grid > n - 1) grid = n - 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (c_is_f64) triu_extract_kernel<double><<<(unsigned)grid, 256, 0, s>>>((const double*)d_c, n, ld, (double*)d_out);
    else triu_extract_kernel<float><<<(unsigned)grid, 256, 0, s>>>((const float*)d_c, n, ld, (float*)d_out);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int skr_pearson_pairs(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale,
                                 const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t K,
                                 const int64_t* d_i, const int64_t* d_j, int64_t npairs, double alpha, float* d_out,
                                 void* stream) {
    if (npairs <= 0) return SKR_OK;
    if (!d_a_hi || !d_a_lo || !d_a_scale || !d_b_hi || !d_b_lo || !d_b_scale || !d_i || !d_j || !d_out || K <= 0)
        return skr::fail(SKR_ERR_ARG, "skr_pearson_pairs: bad argument");
    const long long kp = skr_pearson_k_padded(K);
    const int grid = grid_for(npairs * 32, 256);
    pearson_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __half*)d_a_hi, (const __half*)d_a_lo, d_a_scale, (const __half*)d_b_hi, (const __half*)d_b_lo, d_b_scale, kp,
        (const long long*)d_i, (const long long*)d_j, npairs, alpha, d_out);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}
