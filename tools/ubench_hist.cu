// Micro-benchmark: shared-memory histogram update strategies on random 12-bit keys (dev tool).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_hist tools/ubench_hist.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kBins = 4096;
constexpr int kIters = 4096;   // steps per warp, 32 keys per step

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 20; }

// V0: packed 16-bit sub-counters, one 32-bit atomic per key, CTA-shared histogram
__global__ void v0_atomic_cta(uint32_t* out) {
    __shared__ uint32_t h[kBins / 2];
    for (int i = threadIdx.x; i < kBins / 2; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); atomicAdd(&h[k >> 1], 1u << ((k & 1) * 16)); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[5];
}

// V1: warp-private u16 histogram, plain read-modify-write (drops duplicates: speed of light for RMW)
template <int WARPS> __global__ void v1_rmw16(uint32_t* out) {
    __shared__ uint16_t h[WARPS][kBins];
    uint16_t* my = h[threadIdx.x >> 5];
    for (int i = threadIdx.x & 31; i < kBins; i += 32) my[i] = 0;
    __syncwarp();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); my[k] = my[k] + 1; __syncwarp(); }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * WARPS + (threadIdx.x >> 5)] = my[5];
}

// V2: match_any, group leader adds the group size
template <int WARPS> __global__ void v2_match(uint32_t* out) {
    __shared__ uint16_t h[WARPS][kBins];
    uint16_t* my = h[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < kBins; i += 32) my[i] = 0;
    __syncwarp();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) {
        uint32_t k = lcg(s);
        uint32_t grp = __match_any_sync(0xFFFFFFFFu, k);
        if ((grp & ((1u << lane) - 1)) == 0) my[k] = my[k] + __popc(grp);
        __syncwarp();
    }
    if (lane == 0) out[blockIdx.x * WARPS + (threadIdx.x >> 5)] = my[5];
}

// V3: owner tag (u8) write/read-back detects duplicates; conflict-free steps use plain RMW,
//     the others fall back to match_any
template <int WARPS> __global__ void v3_tag(uint32_t* out) {
    __shared__ uint16_t h[WARPS][kBins];
    __shared__ uint8_t tag[WARPS][kBins];
    uint16_t* my = h[threadIdx.x >> 5];
    uint8_t* tg = tag[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < kBins; i += 32) my[i] = 0;
    __syncwarp();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) {
        uint32_t k = lcg(s);
        tg[k] = (uint8_t)lane;
        __syncwarp();
        bool lost = tg[k] != lane;
        if (!__any_sync(0xFFFFFFFFu, lost)) {
            my[k] = my[k] + 1;
        } else {
            uint32_t grp = __match_any_sync(0xFFFFFFFFu, k);
            if ((grp & ((1u << lane) - 1)) == 0) my[k] = my[k] + __popc(grp);
        }
        __syncwarp();
    }
    if (lane == 0) out[blockIdx.x * WARPS + (threadIdx.x >> 5)] = my[5];
}

// V4: warp-private histogram but atomics (no dedupe needed) on packed words
template <int WARPS> __global__ void v4_atomic_warp(uint32_t* out) {
    __shared__ uint32_t h[WARPS][kBins / 2];
    uint32_t* my = h[threadIdx.x >> 5];
    for (int i = threadIdx.x & 31; i < kBins / 2; i += 32) my[i] = 0;
    __syncwarp();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); atomicAdd(&my[k >> 1], 1u << ((k & 1) * 16)); }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * WARPS + (threadIdx.x >> 5)] = my[5];
}

// V5: only the key generation + match (cost of MATCH.ANY alone); V6: only LCG (loop overhead)
template <int WARPS> __global__ void v5_match_only(uint32_t* out) {
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1, acc = 0;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); acc += __match_any_sync(0xFFFFFFFFu, k); }
    if (acc == 0x12345) out[0] = acc;
}
template <int WARPS> __global__ void v6_lcg_only(uint32_t* out) {
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1, acc = 0;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); acc += k; }
    if (acc == 0x12345) out[0] = acc;
}
// V7: RMW on 32-bit counters, warp-private (16 KB per warp)
template <int WARPS> __global__ void v7_rmw32(uint32_t* out) {
    extern __shared__ uint32_t hd[];
    uint32_t* my = hd + (threadIdx.x >> 5) * kBins;
    for (int i = threadIdx.x & 31; i < kBins; i += 32) my[i] = 0;
    __syncwarp();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    for (int it = 0; it < kIters; ++it) { uint32_t k = lcg(s); my[k] = my[k] + 1; __syncwarp(); }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * WARPS + (threadIdx.x >> 5)] = my[5];
}

// V8: key-width sweep for the small histograms of k = 4, 5 (256 / 1024 bins): red.shared on a CTA-shared histogram
// of 2^bits bins, packed (two 16-bit sub-counters per word) or one 32-bit word per bin, optionally `rep` replicas
// selected by lane (same-word collisions inside a warp are what slows small histograms down)
__global__ void v8_bits(uint32_t* out, int bits, int packed, int rep) {
    __shared__ uint32_t h[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
    const uint32_t mask = (1u << bits) - 1;
    const uint32_t words = packed ? (1u << bits) / 2 : (1u << bits);
    const uint32_t base = (threadIdx.x % rep) * words;
    for (int it = 0; it < kIters; ++it) {
        uint32_t k = lcg(s) & mask;
        if (packed) atomicAdd(&h[base + (k >> 1)], 1u << ((k & 1) * 16));
        else atomicAdd(&h[base + k], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[5];
}

template <class F> void run(const char* name, F launch, int ctas, int threads, int sms, double ghz) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double keys = (double)ctas * threads * kIters;
    cudaError_t e = cudaGetLastError();
    printf("%-28s ctas %5d x %4d thr  %.3f ms  %.1f Gkeys/s  %.2f keys/clk/SM @%.2f GHz  %s\n", name, ctas, threads, ms,
           keys / ms / 1e6, keys / (ms * 1e-3) / (sms * ghz * 1e9), ghz, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz / 1e6;
    uint32_t* out; cudaMalloc(&out, 1 << 22);
    run("V0 atomic CTA-shared 256thr", [&] { v0_atomic_cta<<<sms * 8, 256>>>(out); }, sms * 8, 256, sms, ghz);
    run("V4 atomic warp-private x4", [&] { v4_atomic_warp<4><<<sms * 7, 128>>>(out); }, sms * 7, 128, sms, ghz);
    run("V1 rmw16 warp-private x4", [&] { v1_rmw16<4><<<sms * 7, 128>>>(out); }, sms * 7, 128, sms, ghz);
    run("V1 rmw16 warp-private x1", [&] { v1_rmw16<1><<<sms * 24, 32>>>(out); }, sms * 24, 32, sms, ghz);
    cudaFuncSetAttribute(v7_rmw32<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kBins * 4);
    run("V7 rmw32 warp-private x4", [&] { v7_rmw32<4><<<sms * 3, 128, 4 * kBins * 4>>>(out); }, sms * 3, 128, sms, ghz);
    run("V2 match+leader rmw16 x4", [&] { v2_match<4><<<sms * 7, 128>>>(out); }, sms * 7, 128, sms, ghz);
    run("V3 tag check + rmw16 x4", [&] { v3_tag<4><<<sms * 4, 128>>>(out); }, sms * 4, 128, sms, ghz);
    run("V5 match only x4", [&] { v5_match_only<4><<<sms * 8, 128>>>(out); }, sms * 8, 128, sms, ghz);
    run("V6 lcg only x4", [&] { v6_lcg_only<4><<<sms * 8, 128>>>(out); }, sms * 8, 128, sms, ghz);
    for (int bits : {12, 10, 8, 6}) {
        for (int packed : {1, 0}) {
            for (int rep : {1, 2, 4, 8}) {
                if (((1 << bits) >> packed) * rep > 4096) continue;
                char name[64];
                snprintf(name, sizeof name, "V8 %2d-bit keys %s x%d", bits, packed ? "packed16" : "word32  ", rep);
                run(name, [&] { v8_bits<<<sms * 4, 512>>>(out, bits, packed, rep); }, sms * 4, 512, sms, ghz);
            }
        }
    }
    return 0;
}
