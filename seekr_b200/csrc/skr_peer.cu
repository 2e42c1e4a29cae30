// Small collectives over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC).
//
// skr_min_exchange is the all-reduce of the Log2.post minimum cell (kmer_counts.py:207-208 needs the minimum over
// ALL rows, which are sharded over ranks).  Every rank owns one 64-bit word per epoch parity in every peer's
// exchange buffer.  One warp: lane t stores this rank's cell, tagged with the call's epoch, straight into peer t's
// buffer (a P2P store through NVLink / NVSwitch), then polls its own buffer until the word of rank t carries the
// same epoch, and the warp reduces the `world` cells with shuffles.  One launch, no host round trip, no library
// collective.  A word is (epoch << 33) | (nan_seen << 32) | min_ordered, written with a single 64-bit store, so a
// reader can never see a torn cell.  Two parities: a rank can run at most one exchange ahead of a peer (its next
// publish needs that peer's previous one), so the slot of epoch e is free again when epoch e + 2 is written.
#include <cuda_runtime.h>

#include "skr_common.h"
#include "skr_device.cuh"

namespace {

constexpr unsigned long long kSpinLimitNs = 4000000000ull;  // a missing peer must not hang the GPU: 4 s, then error

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// peers[t]: address of rank t's exchange buffer as mapped into this process; layout [2 parities][world] words
__global__ void min_exchange_kernel(SkrMinCell* cell, unsigned long long* const* peers, int world, int rank,
                                    unsigned long long epoch, int* err) {
    const int lane = threadIdx.x;
    const unsigned long long parity = epoch & 1ull;
    unsigned long long mine = 0;
    if (lane < world) {
        const unsigned long long word = (epoch << 33) | ((unsigned long long)(cell->nan_seen ? 1u : 0u) << 32) |
                                        (unsigned long long)cell->min_ordered;
        st_sys_u64(peers[lane] + parity * world + rank, word);
        const unsigned long long* slot = peers[rank] + parity * world + lane;
        const unsigned long long t0 = globaltimer_ns();
        for (;;) {
            mine = ld_sys_u64(slot);
            if ((mine >> 33) == epoch) break;
            if (globaltimer_ns() - t0 > kSpinLimitNs) {
                atomicExch(err, 1);
                mine = ~0ull;
                break;
            }
        }
    }
    uint32_t mn = lane < world ? (uint32_t)mine : 0xFFFFFFFFu;
    uint32_t nan = lane < world ? (uint32_t)(mine >> 32) & 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
        nan |= __shfl_xor_sync(0xFFFFFFFFu, nan, o);
    }
    if (lane == 0) {
        cell->min_ordered = mn;
        cell->nan_seen = nan;
    }
}

}  // namespace

extern "C" int skr_peer_alloc(size_t bytes, void** d_out, unsigned char* handle64) {
    if (!d_out || !handle64 || bytes == 0) return skr::fail(SKR_ERR_ARG, "skr_peer_alloc: bad argument");
    void* p = nullptr;
    SKR_CUDA_CHECK(cudaMalloc(&p, bytes));
    SKR_CUDA_CHECK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return skr::fail(SKR_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    *d_out = p;
    return SKR_OK;
}

extern "C" int skr_peer_open(const unsigned char* handle64, void** d_out) {
    if (!handle64 || !d_out) return skr::fail(SKR_ERR_ARG, "skr_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SKR_CUDA_CHECK(cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SKR_OK;
}

extern "C" int skr_peer_close(void* d_ptr) {
    if (d_ptr) SKR_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return SKR_OK;
}

extern "C" int skr_peer_free(void* d_ptr) {
    if (d_ptr) SKR_CUDA_CHECK(cudaFree(d_ptr));
    return SKR_OK;
}

extern "C" int skr_min_exchange(SkrMinCell* d_cell, void* const* d_peers, int world, int rank, uint64_t epoch, int* d_err,
                                void* stream) {
    if (!d_cell || !d_peers || !d_err) return skr::fail(SKR_ERR_ARG, "skr_min_exchange: null argument");
    if (world < 1 || world > 32 || rank < 0 || rank >= world)
        return skr::fail(SKR_ERR_ARG, "skr_min_exchange: world must be 1..32 and 0 <= rank < world");
    if (epoch == 0 || epoch >= (1ull << 31)) return skr::fail(SKR_ERR_ARG, "skr_min_exchange: epoch must be 1 .. 2^31 - 1");
    min_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_cell, (unsigned long long* const*)d_peers, world, rank, epoch, d_err);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}
