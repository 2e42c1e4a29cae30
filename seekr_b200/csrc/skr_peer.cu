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

#include <cstdlib>
#include <cstring>

#include "skr_common.h"
#include "skr_device.cuh"

namespace {

// A missing peer must not hang the GPU for ever: after the spin limit the kernel gives up and raises the error
// flag (the caller then falls back to the library all-reduce or raises).  The default is generous (ranks of one
// job may be seconds apart: FASTA parsing, a first pinned allocation); SEEKR_B200_PEER_TIMEOUT_S changes it.
unsigned long long spin_limit_ns() {
    static const unsigned long long v = [] {
        const char* env = getenv("SEEKR_B200_PEER_TIMEOUT_S");
        double s = env ? atof(env) : 60.0;
        if (!(s > 0.0)) s = 60.0;
        return (unsigned long long)(s * 1e9);
    }();
    return v;
}

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
                                    unsigned long long epoch, int* err, unsigned long long kSpinLimitNs,
                                    const uint32_t* skip, uint32_t skip_value, uint32_t flag_value) {
    // skip: the same value on every rank (callers exchange the flag first).  flag_value != 0: the second word of
    // the cell is a flag that counts as set when it EQUALS flag_value (the epoch scheme of the speculative
    // Log2.post route), and the OR over the ranks is written back as flag_value / 0
    if (skip && *skip == skip_value) return;
    const int lane = threadIdx.x;
    const unsigned long long parity = epoch & 1ull;
    unsigned long long mine = 0;
    if (lane < world) {
        const uint32_t bit = flag_value ? (cell->nan_seen == flag_value ? 1u : 0u) : (cell->nan_seen ? 1u : 0u);
        const unsigned long long word = (epoch << 33) | ((unsigned long long)bit << 32) |
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
        cell->nan_seen = flag_value ? (nan ? flag_value : 0u) : nan;
    }
}

// OR of one flag per rank WITHOUT waiting in the common case (the speculative Log2.post route: "some record had a
// zero count in the arg-min column").  A rank whose own flag is set knows the OR already and returns after storing its
// word into every peer; only a rank whose flag is clear waits for the others.  Because ranks no longer wait for each
// other every epoch, a waiting rank can find a peer's slot already overwritten by a LATER epoch; every word therefore
// carries, beside the writer's own flag for its epoch, the OUTCOMES (global ORs) of the writer's previous 31 epochs --
// a rank only moves on once it knows the outcome of its epoch (own flag set, or learnt by waiting), so a peer that is
// d epochs ahead can tell a straggler what epoch - d turned out to be.  word = epoch << 32 | outcomes(epoch-31 ..
// epoch-1) << 1 | own flag; slots [world] after the two parities of the minimum exchange; `state` = this rank's
// outcome history.  A straggler more than 31 epochs behind sets *err (callers raise).
__global__ void flag_or_kernel(uint32_t* flag, uint32_t flag_value, unsigned long long* const* peers, int world, int rank,
                               unsigned long long epoch, uint32_t* state, int* err, unsigned long long kSpinLimitNs) {
    const int lane = threadIdx.x;
    const uint32_t mine = *flag == flag_value ? 1u : 0u;
    const uint32_t hist = *state & 0x7FFFFFFFu;  // bit d - 1 = outcome of epoch - d
    const unsigned long long word = (epoch << 32) | ((unsigned long long)hist << 1) | mine;
    const size_t base = (size_t)2 * world;  // behind the minimum exchange's [2][world] words
    if (lane < world && lane != rank) st_sys_u64(peers[lane] + base + rank, word);
    uint32_t outcome = mine;
    if (!mine) {  // nobody here saw it: ask the others
        uint32_t bit = 0;
        if (lane < world && lane != rank) {
            const unsigned long long* slot = peers[rank] + base + lane;
            const unsigned long long t0 = globaltimer_ns();
            for (;;) {
                const unsigned long long w = ld_sys_u64(slot);
                const unsigned long long we = w >> 32;
                if (we >= epoch) {
                    const unsigned long long d = we - epoch;
                    if (d > 31) atomicExch(err, 1);  // too far ahead to tell
                    else bit = (uint32_t)(w >> d) & 1u;  // d = 0: its own flag; d > 0: the outcome it recorded for our epoch
                    break;
                }
                if (globaltimer_ns() - t0 > kSpinLimitNs) {
                    atomicExch(err, 1);
                    break;
                }
            }
        }
        outcome = __any_sync(0xFFFFFFFFu, bit) ? 1u : 0u;
    }
    if (lane == 0) {
        if (!mine) *flag = outcome ? flag_value : 0u;
        *state = ((hist << 1) | outcome) & 0x7FFFFFFFu;
    }
}

// All-reduce(sum) of n binary64 column partials fused with the finishing step of the column statistics
// (skr_col_finish_f64: / total rows, optional sqrt, one rounding to fp32, quality flag).  Exchange buffer of a rank:
// [2 parities][world][n_cap] doubles, then [2][world] 64-bit flags, then one 32-bit CTA counter.  Every CTA stores
// its slice of this rank's partials into every peer (P2P stores), fences, and counts itself done; the last CTA
// of the rank publishes the epoch flag to all peers.  Each CTA then waits until the flags of all ranks carry the
// epoch and adds the `world` slices in rank order, so every rank computes the same bits.  A rank's CTAs never
// wait for their own rank's later CTAs (stores come first), and peers do not depend on us: no circular wait.
__global__ void __launch_bounds__(256) colstat_exchange_kernel(const double* __restrict__ acc, unsigned char* const* peers,
                                                               int world, int rank, unsigned long long epoch, long long n,
                                                               long long n_cap, long long total_rows, int take_sqrt,
                                                               float* __restrict__ out, int* flag, int* err,
                                                               unsigned long long kSpinLimitNs) {
    const unsigned long long parity = epoch & 1ull;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t flags_off = (size_t)2 * world * n_cap * sizeof(double);
    const size_t counter_off = flags_off + (size_t)2 * world * sizeof(unsigned long long);
    if (j < n) {
        const double v = acc[j];
        for (int t = 0; t < world; ++t) {
            double* dst = reinterpret_cast<double*>(peers[t]) + (parity * world + rank) * n_cap + j;
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        s_bad = 0;
        unsigned int* counter = reinterpret_cast<unsigned int*>(peers[rank] + counter_off);
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {  // every CTA of this rank has stored and fenced
            *counter = 0u;
            __threadfence_system();
            for (int t = 0; t < world; ++t) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(peers[t] + flags_off) + parity * world + rank;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
            }
        }
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank] + flags_off) + parity * world;
        const unsigned long long t0 = globaltimer_ns();
        for (int t = 0; t < world; ++t) {
            for (;;) {
                unsigned long long f;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine + t) : "memory");
                if (f == epoch) break;
                if (globaltimer_ns() - t0 > kSpinLimitNs) {
                    atomicExch(err, 1);
                    s_bad = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (j >= n || s_bad) return;
    const double* slices = reinterpret_cast<const double*>(peers[rank]) + parity * world * n_cap;
    double sum = 0.0;
    for (int t = 0; t < world; ++t) {
        double v;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(slices + (size_t)t * n_cap + j) : "memory");
        sum += v;
    }
    double r = sum / (double)total_rows;
    if (take_sqrt) r = sqrt(r);
    const float o = (float)r;
    out[j] = o;
    if (flag) {
        int bits = 0;
        if (!(o - o == 0.0f)) bits |= 1;
        if (!(o > 0.0f)) bits |= 2;
        if (bits) atomicOr(flag, bits);
    }
}

// The ONE exchange of the accurate column statistics (skr_count_ex's d_colsum / d_colsq) fused with their finish:
// acc = [2][cols] binary64 (sum, sum of squares).  Same protocol and buffer layout as colstat_exchange_kernel with
// n = 2 * cols values; thread j finishes column j: mean = S1 / rows, std = sqrt(max(S2 / rows - mean^2, 0)).
__global__ void __launch_bounds__(256) colsum_exchange_kernel(const double* __restrict__ acc, unsigned char* const* peers,
                                                              int world, int rank, unsigned long long epoch, long long cols,
                                                              long long n_cap, long long total_rows, float* __restrict__ mean,
                                                              float* __restrict__ std_, int* flags, int* err,
                                                              unsigned long long kSpinLimitNs) {
    const unsigned long long parity = epoch & 1ull;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t flags_off = (size_t)2 * world * n_cap * sizeof(double);
    const size_t counter_off = flags_off + (size_t)2 * world * sizeof(unsigned long long);
    if (j < cols) {
        const double s1 = acc[j], s2 = acc[cols + j];
        for (int t = 0; t < world; ++t) {
            double* dst = reinterpret_cast<double*>(peers[t]) + (parity * world + rank) * n_cap;
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + j), "d"(s1) : "memory");
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + cols + j), "d"(s2) : "memory");
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        s_bad = 0;
        unsigned int* counter = reinterpret_cast<unsigned int*>(peers[rank] + counter_off);
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            for (int t = 0; t < world; ++t) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(peers[t] + flags_off) + parity * world + rank;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
            }
        }
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank] + flags_off) + parity * world;
        const unsigned long long t0 = globaltimer_ns();
        for (int t = 0; t < world; ++t) {
            for (;;) {
                unsigned long long f;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine + t) : "memory");
                if (f == epoch) break;
                if (globaltimer_ns() - t0 > kSpinLimitNs) {
                    atomicExch(err, 1);
                    s_bad = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (j >= cols || s_bad) return;
    const double* slices = reinterpret_cast<const double*>(peers[rank]) + parity * world * n_cap;
    double s1 = 0.0, s2 = 0.0;
    for (int t = 0; t < world; ++t) {
        double a, b;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(a) : "l"(slices + (size_t)t * n_cap + j) : "memory");
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(b) : "l"(slices + (size_t)t * n_cap + cols + j) : "memory");
        s1 += a;
        s2 += b;
    }
    const double mu = s1 / (double)total_rows;
    double var = s2 / (double)total_rows - mu * mu;
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

}  // namespace

extern "C" int skr_colsum_exchange(const double* d_acc, void* const* d_peers, int world, int rank, uint64_t epoch, int64_t cols,
                                   int64_t n_cap, int64_t total_rows, float* d_mean, float* d_std, int* d_flags, int* d_err,
                                   void* stream) {
    if (cols <= 0) return SKR_OK;
    if (!d_acc || !d_peers || !d_err || total_rows <= 0) return skr::fail(SKR_ERR_ARG, "skr_colsum_exchange: bad argument");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || 2 * cols > n_cap)
        return skr::fail(SKR_ERR_ARG, "skr_colsum_exchange: bad world / rank / capacity");
    if (epoch == 0) return skr::fail(SKR_ERR_ARG, "skr_colsum_exchange: epoch starts at 1");
    const long long blocks = (cols + 255) / 256;
    if (blocks > 148 * 4) return skr::fail(SKR_ERR_ARG, "skr_colsum_exchange: vector too long for a co-resident grid");
    colsum_exchange_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_acc, (unsigned char* const*)d_peers, world, rank,
                                                                               epoch, cols, n_cap, total_rows, d_mean, d_std,
                                                                               d_flags, d_err, spin_limit_ns());
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

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
    return skr_min_exchange_skip(d_cell, d_peers, world, rank, epoch, nullptr, 0, 0, d_err, stream);
}

extern "C" int skr_min_exchange_skip(SkrMinCell* d_cell, void* const* d_peers, int world, int rank, uint64_t epoch,
                                     const uint32_t* d_skip, uint32_t skip_value, uint32_t flag_value, int* d_err,
                                     void* stream) {
    if (!d_cell || !d_peers || !d_err) return skr::fail(SKR_ERR_ARG, "skr_min_exchange: null argument");
    if (world < 1 || world > 32 || rank < 0 || rank >= world)
        return skr::fail(SKR_ERR_ARG, "skr_min_exchange: world must be 1..32 and 0 <= rank < world");
    if (epoch == 0 || epoch >= (1ull << 31)) return skr::fail(SKR_ERR_ARG, "skr_min_exchange: epoch must be 1 .. 2^31 - 1");
    min_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_cell, (unsigned long long* const*)d_peers, world, rank, epoch, d_err,
                                                            spin_limit_ns(), d_skip, skip_value ? skip_value : 1u, flag_value);
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int64_t skr_min_exchange_bytes(int world) { return (int64_t)3 * world * 8; }

extern "C" int skr_flag_or_exchange(uint32_t* d_flag, uint32_t flag_value, void* const* d_peers, int world, int rank,
                                    uint64_t epoch, uint32_t* d_state, int* d_err, void* stream) {
    if (!d_flag || !d_peers || !d_state || !d_err) return skr::fail(SKR_ERR_ARG, "skr_flag_or_exchange: null argument");
    if (world < 1 || world > 32 || rank < 0 || rank >= world)
        return skr::fail(SKR_ERR_ARG, "skr_flag_or_exchange: world must be 1..32 and 0 <= rank < world");
    if (epoch == 0 || epoch >= (1ull << 32) || flag_value == 0)
        return skr::fail(SKR_ERR_ARG, "skr_flag_or_exchange: epoch must be 1 .. 2^32 - 1 and flag_value non-zero");
    flag_or_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_flag, flag_value, (unsigned long long* const*)d_peers, world, rank, epoch,
                                                       d_state, d_err, spin_limit_ns());
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}

extern "C" int64_t skr_colstat_exchange_bytes(int world, int64_t n_cap) {
    return (int64_t)2 * world * n_cap * 8 + (int64_t)2 * world * 8 + 64;
}

extern "C" int skr_colstat_exchange(const double* d_acc, void* const* d_peers, int world, int rank, uint64_t epoch, int64_t n,
                                    int64_t n_cap, int64_t total_rows, int take_sqrt, float* d_out, int* d_flag, int* d_err,
                                    void* stream) {
    if (n <= 0) return SKR_OK;
    if (!d_acc || !d_peers || !d_out || !d_err || total_rows <= 0) return skr::fail(SKR_ERR_ARG, "skr_colstat_exchange: bad argument");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || n > n_cap)
        return skr::fail(SKR_ERR_ARG, "skr_colstat_exchange: bad world / rank / capacity");
    if (epoch == 0) return skr::fail(SKR_ERR_ARG, "skr_colstat_exchange: epoch starts at 1");
    const long long blocks = (n + 255) / 256;
    if (blocks > 148 * 4) return skr::fail(SKR_ERR_ARG, "skr_colstat_exchange: vector too long for a co-resident grid");
    if (d_flag) SKR_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), (cudaStream_t)stream));
    colstat_exchange_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_acc, (unsigned char* const*)d_peers, world, rank,
                                                                                 epoch, n, n_cap, total_rows, take_sqrt, d_out,
                                                                                 d_flag, d_err, spin_limit_ns());
    SKR_LAUNCH_CHECK();
    return SKR_OK;
}
